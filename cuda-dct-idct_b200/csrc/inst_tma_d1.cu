#define INST_SPARSE 0
#define INST_Q 1
#define INST_TAG d1
#include "inst_tma.cuh"
