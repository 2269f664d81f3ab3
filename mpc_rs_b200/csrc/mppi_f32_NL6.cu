// mppi_f32_NL6.cu — FP32 fast path (FMA contraction on, constants pre-folded on the host, see models.cuh), model NL6.
#define MPCB_INST_MODEL ModelNL6
#define MPCB_INST_REAL float
#define MPCB_INST_FN mppi_kernel_f32_NL6
#define MPCB_INST_SAMPLES512 1
#include "mppi_inst.cuh"
