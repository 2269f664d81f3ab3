// mppi_f64fast_short.cu — the short-horizon FP64 control step (mppi_short_kernel.cuh) for MPCB_F64_FAST: the folded
// model forms with FMA contraction (Makefile rule mppi_f64fast_%.o).
#include "mppi_short_kernel.cuh"

namespace mpcb {
MPCB_SHORT_TABLE(mppi_kernel_f64fast_short, ModelLF, ModelNLF, ModelNL6F)
}  // namespace mpcb
