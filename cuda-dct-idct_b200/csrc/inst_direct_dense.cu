#define INST_SPARSE 0
#define INST_NAME launch_direct_dense
#include "inst_direct.cuh"
