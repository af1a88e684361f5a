// Register-resident FFT: slice 2 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans2[] = {
#include "fft_reg_sizes_p2.inc"
};
RegPlanSlice fft_reg_slice2() { return {kPlans2, (int)(sizeof(kPlans2) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
