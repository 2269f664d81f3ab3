// ukf_n6_fast.cu — the same kernels as ukf_n6.cu in their FAST form: FMA contraction on (this TU is compiled
// without -fmad=false), symmetric-half covariance accumulation, shared sigma weight factored out.
#define MPCB_UKF_FAST true
#define MPCB_UKF_ENTRY ukf_kernel_n6_fast
#include "ukf_n6.cu"
