// mppi_f64_L.cu — FP64 reference-order path, model L.  Compiled with -fmad=false: the reference (Rust f64)
// never contracts a*b+c, and this path reproduces it.
#define MPCB_INST_MODEL ModelL
#define MPCB_INST_REAL double
#define MPCB_INST_FN mppi_kernel_f64_L
#include "mppi_inst.cuh"
