// inst_tma.cuh -- instantiates k_tma for one T family (INST_SPARSE = true/false).
#include "dct_kernels.cuh"

namespace b200dct {

#define B200_TMA_CASE(M, Q, X)                                                                        \
    if (mode == (M) && qmode == (Q) && pix == (X)) {                                                  \
        auto kern = k_tma<M, INST_SPARSE, Q, X>;                                                      \
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        kern<<<grid, block, smem, s>>>(P);                                                            \
        return cudaGetLastError();                                                                    \
    }
#define B200_TMA_MODES(Q, X) B200_TMA_CASE(MODE_FWD, Q, X) B200_TMA_CASE(MODE_INV, Q, X) B200_TMA_CASE(MODE_RT, Q, X)

cudaError_t INST_NAME(int mode, int qmode, int pix, const TmaParams &P, int grid, int block, size_t smem, cudaStream_t s)
{
#ifdef B200DCT_FAST_BUILD /* experiment builds: headline kernels only */
#if INST_SPARSE
    B200_TMA_MODES(Q_IMM, DT_F32)
    B200_TMA_CASE(MODE_RT, Q_IMM, DT_U8)
#endif
    return cudaErrorInvalidValue;
#endif
#if INST_SPARSE
    B200_TMA_MODES(Q_IMM, DT_F32)
    B200_TMA_MODES(Q_IMM, DT_U8)
#endif
    B200_TMA_MODES(Q_PARAM, DT_F32)
    B200_TMA_MODES(Q_PARAM, DT_U8)
    B200_TMA_MODES(Q_PARAM_DIV, DT_F32)
    B200_TMA_MODES(Q_PARAM_DIV, DT_U8)
    return cudaErrorInvalidValue;
}

} // namespace b200dct
