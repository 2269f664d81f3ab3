// mppi_f32_L.cu — FP32 fast path (FMA contraction on, constants pre-folded on the host, see models.cuh), model L.
#define MPCB_INST_MODEL ModelL
#define MPCB_INST_REAL float
#define MPCB_INST_FN mppi_kernel_f32_L
#define MPCB_INST_SAMPLES512 1
#include "mppi_inst.cuh"
