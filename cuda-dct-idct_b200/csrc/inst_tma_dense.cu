#define INST_SPARSE 0
#define INST_NAME launch_tma_dense
#include "inst_tma.cuh"
