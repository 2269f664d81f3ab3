// mppi_f64fast_L.cu — FP64 fast path (MPCB_F64_FAST), model L: the folded formulas of the FP32 kernels in double, with
// FMA contraction (this TU is NOT compiled with -fmad=false; see models.cuh, ModelLF).
#define MPCB_INST_MODEL ModelLF
#define MPCB_INST_REAL double
#define MPCB_INST_FN mppi_kernel_f64fast_L
#include "mppi_inst.cuh"
