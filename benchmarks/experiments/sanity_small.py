"""Small run of every kernel family/variant (for compute-sanitizer memcheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
import cuda_dct_idct_b200 as m
from oracle import oracle as o
ok = True
for path in (2, 1):
    for (H, W) in ((8, 32), (72, 544), (40, 256 + 32)):
        img = o.rand_image(H, W, 1)
        for T in (None, o.dct2_T()):
            for keep in (m.ALL_COEFFS, o.zigzag_mask(6)):
                for Q in (None, o.jpeg_Q() * 0.37):
                    plan = m.Plan(T=T, Q=Q, keep=keep, path=path)
                    want_out, want_coef = o.roundtrip(img, T=T, Q=Q, keep=keep, want_coef=True)
                    d = torch.from_numpy(img).cuda()
                    for cdt in (None, torch.float32, torch.int16):
                        coef = None if cdt is None else torch.empty(H, W, dtype=cdt, device="cuda")
                        out = m.roundtrip(d, coef=coef, plan=plan)
                        ok &= np.array_equal(out.cpu().numpy().view(np.uint32), want_out.view(np.uint32))
                    c = m.forward(d, plan=plan); r = m.inverse(c, plan=plan)
                    ok &= np.array_equal(c.cpu().numpy(), want_coef) and np.array_equal(r.cpu().numpy().view(np.uint32), want_out.view(np.uint32))
                    c16 = m.forward(d, plan=plan, coef_dtype=torch.int16); r8 = m.inverse(c16, plan=plan, img_dtype=torch.uint8)
                    ok &= np.array_equal(r8.cpu().numpy(), o.to_u8(want_out))
                    u8 = torch.from_numpy(img.astype(np.uint8)).cuda()
                    ok &= np.array_equal(m.roundtrip(u8, plan=plan).cpu().numpy(), o.roundtrip(img.astype(np.uint8), T=T, Q=Q, keep=keep))
h = o.rand_image(64, 64, 2); ok &= np.array_equal(m.roundtrip_host(h).view(np.uint32), o.roundtrip(h).view(np.uint32))
a = torch.from_numpy(h).cuda(); print(m.metrics(a, m.roundtrip(a)))
torch.cuda.synchronize(); print("all ok" if ok else "MISMATCH"); sys.exit(0 if ok else 1)
