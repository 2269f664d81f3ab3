// mppi_ws_L.cu — warp-specialised FP32 kernels (mppi_ws_kernel.cuh), model L.
#define MPCB_INST_MODEL ModelL
#define MPCB_INST_FN mppi_kernel_ws_L
#include "mppi_ws_inst.cuh"
