#define INST_PIX DT_F32
#define INST_TAG f32
#include "inst_any.cuh"
