// mppi_f64_short.cu — the short-horizon FP64 control step (mppi_short_kernel.cuh) in reference order: compiled with
// -fmad=false like the other mppi_f64_*.cu (Makefile rule mppi_f64_%.o).
#include "mppi_short_kernel.cuh"

namespace mpcb {
MPCB_SHORT_TABLE(mppi_kernel_f64_short, ModelL, ModelNL, ModelNL6)
}  // namespace mpcb
