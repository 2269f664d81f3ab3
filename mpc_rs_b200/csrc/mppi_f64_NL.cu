// mppi_f64_NL.cu — FP64 reference-order path, model NL.  Compiled with -fmad=false: the reference (Rust f64)
// never contracts a*b+c, and this path reproduces it.
#define MPCB_INST_MODEL ModelNL
#define MPCB_INST_REAL double
#define MPCB_INST_FN mppi_kernel_f64_NL
#include "mppi_inst.cuh"
