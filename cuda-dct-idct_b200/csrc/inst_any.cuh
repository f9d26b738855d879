// inst_any.cuh -- instantiates k_any (single-pass round trip for any image size and alignment)
// for one pixel type: INST_PIX = DT_F32 / DT_U8, INST_TAG = name suffix.
#include "dct_kernels.cuh"

namespace b200dct {

#define B200_CAT_(a, b) a##b
#define B200_CAT(a, b) B200_CAT_(a, b)

template <bool SPARSE, int QM, int PIX>
static cudaError_t launch_one(const AnyParams &P, dim3 grid, dim3 block, cudaStream_t s, bool pdl)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_any<SPARSE, QM, PIX>, P);
}

template <int PIX>
static cudaError_t launch_pix(bool sparse, int qm, const AnyParams &P, dim3 grid, dim3 block, cudaStream_t s, bool pdl)
{
    if (sparse) {
        if (qm == Q_IMM) return launch_one<true, Q_IMM, PIX>(P, grid, block, s, pdl);
        if (qm == Q_PARAM) return launch_one<true, Q_PARAM, PIX>(P, grid, block, s, pdl);
        return launch_one<true, Q_PARAM_DIV, PIX>(P, grid, block, s, pdl);
    }
    if (qm == Q_PARAM) return launch_one<false, Q_PARAM, PIX>(P, grid, block, s, pdl);
    return launch_one<false, Q_PARAM_DIV, PIX>(P, grid, block, s, pdl);
}

cudaError_t B200_CAT(launch_any_, INST_TAG)(bool sparse, int qm, const AnyParams &P, dim3 grid, dim3 block, cudaStream_t s, bool pdl)
{
    return launch_pix<INST_PIX>(sparse, qm, P, grid, block, s, pdl);
}

} // namespace b200dct
