// inst_direct.cuh -- instantiates k_direct for one T family (INST_SPARSE = true/false).
#include "dct_kernels.cuh"

namespace b200dct {

#define B200_DIRECT_CASE(M, Q, X)                                                   \
    if (mode == (M) && qmode == (Q) && pix == (X)) {                                \
        k_direct<M, INST_SPARSE, Q, X><<<grid, block, 0, s>>>(P);                   \
        return cudaGetLastError();                                                  \
    }
#define B200_DIRECT_MODES(Q, X) \
    B200_DIRECT_CASE(MODE_FWD, Q, X) B200_DIRECT_CASE(MODE_INV, Q, X) B200_DIRECT_CASE(MODE_RT, Q, X)

cudaError_t INST_NAME(int mode, int qmode, int pix, const DirectParams &P, dim3 grid, dim3 block, cudaStream_t s)
{
#ifdef B200DCT_FAST_BUILD /* experiment builds: headline kernels only */
#if INST_SPARSE
    B200_DIRECT_MODES(Q_IMM, DT_F32)
    B200_DIRECT_CASE(MODE_RT, Q_IMM, DT_U8)
#endif
    return cudaErrorInvalidValue;
#endif
#if INST_SPARSE
    B200_DIRECT_MODES(Q_IMM, DT_F32)
    B200_DIRECT_MODES(Q_IMM, DT_U8)
#endif
    B200_DIRECT_MODES(Q_PARAM, DT_F32)
    B200_DIRECT_MODES(Q_PARAM, DT_U8)
    B200_DIRECT_MODES(Q_PARAM_DIV, DT_F32)
    B200_DIRECT_MODES(Q_PARAM_DIV, DT_U8)
    return cudaErrorInvalidValue;
}

} // namespace b200dct
