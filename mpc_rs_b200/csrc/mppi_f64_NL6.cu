// mppi_f64_NL6.cu — FP64 reference-order path, model NL6.  Compiled with -fmad=false: the reference (Rust f64)
// never contracts a*b+c, and this path reproduces it.
#define MPCB_INST_MODEL ModelNL6
#define MPCB_INST_REAL double
#define MPCB_INST_FN mppi_kernel_f64_NL6
#include "mppi_inst.cuh"
