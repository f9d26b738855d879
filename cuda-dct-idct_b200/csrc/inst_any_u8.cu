#define INST_PIX DT_U8
#define INST_TAG u8
#include "inst_any.cuh"
