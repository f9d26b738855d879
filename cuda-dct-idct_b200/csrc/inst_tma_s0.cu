#define INST_SPARSE 1
#define INST_Q 0
#define INST_TAG s0
#include "inst_tma.cuh"
