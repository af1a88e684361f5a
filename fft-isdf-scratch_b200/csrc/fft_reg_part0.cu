// Register-resident FFT: slice 0 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans0[] = {
#include "fft_reg_sizes_p0.inc"
};
RegPlanSlice fft_reg_slice0() { return {kPlans0, (int)(sizeof(kPlans0) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
