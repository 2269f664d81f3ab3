// mppi_f32x2_L.cu — FP32 fast path with two samples per thread and packed f32x2 arithmetic (f32x2.cuh), model L.
#define MPCB_INST_MODEL ModelL
#define MPCB_INST_REAL float
#define MPCB_INST_FN mppi_kernel_f32x2_L
#define MPCB_INST_SPT 2
#define MPCB_INST_SAMPLES512 1
#include "mppi_inst.cuh"
