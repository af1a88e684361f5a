// Register-resident FFT: slice 3 of the size table (see fft_reg.cu).
#include "fft_reg_kernels.cuh"
namespace isdf {
namespace fftreg {
static const RegPlan kPlans3[] = {
#include "fft_reg_sizes_p3.inc"
};
RegPlanSlice fft_reg_slice3() { return {kPlans3, (int)(sizeof(kPlans3) / sizeof(RegPlan))}; }
}  // namespace fftreg
}  // namespace isdf
