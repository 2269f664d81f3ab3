// ukf_n4_fast.cu — the same kernels as ukf_n4.cu in their FAST form: FMA contraction on (this TU is compiled
// without -fmad=false), symmetric-half covariance accumulation, shared sigma weight factored out.
#define MPCB_UKF_FAST true
#define MPCB_UKF_ENTRY ukf_kernel_n4_fast
#include "ukf_n4.cu"
