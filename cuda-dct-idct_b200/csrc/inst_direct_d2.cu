#define INST_SPARSE 0
#define INST_Q 2
#define INST_TAG d2
#include "inst_direct.cuh"
