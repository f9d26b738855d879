#include "dct_kernels.cuh"
namespace b200dct { template __global__ void k_tma<MODE_RT, true, Q_IMM, DT_F32>(const __grid_constant__ TmaParams); 
template __global__ void k_direct<MODE_RT, true, Q_IMM, DT_F32>(const __grid_constant__ DirectParams); }
