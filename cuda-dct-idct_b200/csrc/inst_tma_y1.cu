// dense T with symmetric even / antisymmetric odd rows (TK_DENSE_SYM), quantiser variant 1
#define INST_SPARSE 2
#define INST_Q 1
#define INST_TAG y1
#include "inst_tma.cuh"
