// Register-resident FFT: slice 4 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans4[] = {
#include "fft_reg_sizes_p4.inc"
};
RegPlanSlice fft_reg_slice4() { return {kPlans4, (int)(sizeof(kPlans4) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
