#define INST_SPARSE 1
#define INST_Q 2
#define INST_TAG s2
#include "inst_tma.cuh"
