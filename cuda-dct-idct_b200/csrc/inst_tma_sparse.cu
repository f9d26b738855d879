#define INST_SPARSE 1
#define INST_NAME launch_tma_sparse
#include "inst_tma.cuh"
