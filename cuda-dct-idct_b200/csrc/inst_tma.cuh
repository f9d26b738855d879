// inst_tma.cuh -- instantiates k_tma for one (T family, quantiser) pair (see inst_direct.cuh).
#include "dct_kernels.cuh"

namespace b200dct {

#define B200_CAT_(a, b) a##b
#define B200_CAT(a, b) B200_CAT_(a, b)

#define B200_TMA_CASE(M, X, F) B200_TMA_CASE_M(M, X, F, false)
#define B200_TMA_CASE_M(M, X, F, MET)                                                                 \
    if (mode == (M) && pix == (X) && finv == (F) && (P.macc != nullptr) == (MET)) {                   \
        auto kern = k_tma<M, INST_SPARSE, INST_Q, X, F, MET>;                                         \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        cudaLaunchConfig_t cfg = {};                                                                  \
        cfg.gridDim = dim3((unsigned)grid);                                                           \
        cfg.blockDim = dim3((unsigned)block);                                                         \
        cfg.dynamicSmemBytes = smem;                                                                  \
        cfg.stream = s;                                                                               \
        cudaLaunchAttribute attr[1];                                                                  \
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                              \
        attr[0].val.programmaticStreamSerializationAllowed = 1;                                       \
        cfg.attrs = attr;                                                                             \
        cfg.numAttrs = pdl ? 1 : 0;                                                                   \
        return cudaLaunchKernelEx(&cfg, kern, P);                                                     \
    }

cudaError_t B200_CAT(launch_tma_, INST_TAG)(int mode, int pix, bool finv, const TmaParams &P, int grid, int block, size_t smem, cudaStream_t s, bool pdl)
{
    B200_TMA_CASE(MODE_RT, DT_F32, false)
    B200_TMA_CASE(MODE_RT, DT_U8, false)
#if INST_SPARSE == 1
    B200_TMA_CASE(MODE_RT, DT_U8, true)
    B200_TMA_CASE_M(MODE_RT, DT_F32, false, true) // fused MSE / PEEN / non-zero count (P.macc != NULL)
#endif
#ifdef B200DCT_FAST_BUILD
#if INST_Q == 0
    B200_TMA_CASE(MODE_FWD, DT_F32, false)
    B200_TMA_CASE(MODE_INV, DT_F32, false)
#endif
    return cudaErrorInvalidValue;
#else
    B200_TMA_CASE(MODE_FWD, DT_F32, false)
    B200_TMA_CASE(MODE_FWD, DT_U8, false)
    B200_TMA_CASE(MODE_INV, DT_F32, false)
    B200_TMA_CASE(MODE_INV, DT_U8, false)
#if INST_SPARSE == 1
    B200_TMA_CASE(MODE_INV, DT_U8, true)
#endif
    return cudaErrorInvalidValue;
#endif
}

} // namespace b200dct
