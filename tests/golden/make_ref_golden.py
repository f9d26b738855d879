"""Run ON THE GPU BOX: writes gpurun_out/golden/refgpu_*.npz from the UNMODIFIED reference
kernels (oracle/_ref/libref_newappr.so, compiled from /root/reference/main_newAppr.cu).
Copy the files into tests/golden/ and commit them: they pin the CPU oracle to real
reference outputs, so the CPU-only suite can check the oracle against the reference.

    gpurun -- python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

import inputs  # noqa: E402
import refgpu  # noqa: E402
from oracle import oracle as o  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    T = torch.from_numpy(o.haweel_T()).cuda()
    refgpu.set_quant("newappr", o.jpeg_Q())
    cases = {
        "rand128": o.rand_image(128, 128, 42),
        "adversarial": inputs.adversarial(16),
        "floatnoise": inputs.float_noise(32, 64),
    }
    for name, img in cases.items():
        d = torch.from_numpy(img).cuda()
        coef, _ = refgpu.dct("newappr", d, T)
        rec, _ = refgpu.idct("newappr", coef, T)
        torch.cuda.synchronize()
        np.savez_compressed(os.path.join(out_dir, f"refgpu_{name}.npz"), img=img, coef=coef.cpu().numpy(),
                            shifted=d.cpu().numpy(), rec=rec.cpu().numpy())
        print("wrote", name)


if __name__ == "__main__":
    main()
