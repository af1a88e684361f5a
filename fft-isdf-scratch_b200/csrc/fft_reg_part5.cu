// Register-resident FFT: slice 5 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans5[] = {
#include "fft_reg_sizes_p5.inc"
};
RegPlanSlice fft_reg_slice5() { return {kPlans5, (int)(sizeof(kPlans5) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
