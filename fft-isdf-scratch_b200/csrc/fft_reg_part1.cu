// Register-resident FFT: slice 1 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans1[] = {
#include "fft_reg_sizes_p1.inc"
};
RegPlanSlice fft_reg_slice1() { return {kPlans1, (int)(sizeof(kPlans1) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
