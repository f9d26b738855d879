#define INST_SPARSE 1
#define INST_NAME launch_direct_sparse
#include "inst_direct.cuh"
