#define INST_SPARSE 1
#define INST_Q 1
#define INST_TAG s1
#include "inst_tma.cuh"
