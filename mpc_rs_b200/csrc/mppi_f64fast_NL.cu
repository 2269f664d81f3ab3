// mppi_f64fast_NL.cu — FP64 fast path (MPCB_F64_FAST), model NL: the folded formulas of the FP32 kernels in double, with
// FMA contraction (this TU is NOT compiled with -fmad=false; see models.cuh, ModelNLF).
#define MPCB_INST_MODEL ModelNLF
#define MPCB_INST_REAL double
#define MPCB_INST_FN mppi_kernel_f64fast_NL
#include "mppi_inst.cuh"
