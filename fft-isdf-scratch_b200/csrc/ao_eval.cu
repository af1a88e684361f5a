// Periodic Gaussian AO evaluation on the device (input producer for synthetic cells; SURVEY.md section 8
// next-row f-3: replaces pcell.pbc_eval_gto("GTOval", coords, kpts=...) at /root/reference/fftisdf.py:367-370
// and NumInt.block_loop at :350-352 for `SyntheticCell`):
//     out[k][g][mu] = sum_T e^{i k.T} * N_mu * P_mu(r_g - c_mu - T) * exp(-alpha_mu |r_g - c_mu - T|^2)
// with P_mu a polynomial of up to 3 cartesian monomials (s, p, five d combinations).
#include "common.cuh"

namespace isdf {

constexpr int AO_KC = 8;      // k-points accumulated per thread
constexpr int AO_MAXTERM = 3;

struct AoDesc {               // one per AO, device array
  double cx, cy, cz, alpha, norm;
  double coef[AO_MAXTERM];
  int pw[AO_MAXTERM][3];
  int nterm;
};

__device__ __forceinline__ double ipow(double x, int p) { return p == 0 ? 1.0 : (p == 1 ? x : x * x); }

__global__ void __launch_bounds__(128)
ao_eval_kernel(const double* __restrict__ coords, long npts, const AoDesc* __restrict__ aos, int nao,
               const double* __restrict__ images, int nimg, const cplx* __restrict__ kphase /*[nk][nimg]*/, int nk,
               cplx* __restrict__ out) {
  const long w = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (g, mu) pair, mu fastest
  if (w >= npts * nao) return;
  const int mu = (int)(w % nao);
  const long g = w / nao;
  const int k0 = blockIdx.y * AO_KC;
  const AoDesc ao = aos[mu];
  const double rx = coords[3 * g] - ao.cx, ry = coords[3 * g + 1] - ao.cy, rz = coords[3 * g + 2] - ao.cz;
  cplx acc[AO_KC];
#pragma unroll
  for (int kk = 0; kk < AO_KC; ++kk) acc[kk] = make_double2(0.0, 0.0);
  for (int t = 0; t < nimg; ++t) {
    const double dx = rx - images[3 * t], dy = ry - images[3 * t + 1], dz = rz - images[3 * t + 2];
    const double a_r2 = ao.alpha * (dx * dx + dy * dy + dz * dz);
    if (a_r2 > 46.0) continue;                       // exp(-46) = 1e-20: below FP64 resolution of the sum
    double ang = 0.0;
    for (int i = 0; i < ao.nterm; ++i)
      ang += ao.coef[i] * ipow(dx, ao.pw[i][0]) * ipow(dy, ao.pw[i][1]) * ipow(dz, ao.pw[i][2]);
    const double chi = ao.norm * ang * exp(-a_r2);
#pragma unroll
    for (int kk = 0; kk < AO_KC; ++kk) {
      if (k0 + kk < nk) {
        const cplx ph = kphase[(long)(k0 + kk) * nimg + t];
        acc[kk].x = fma(chi, ph.x, acc[kk].x);
        acc[kk].y = fma(chi, ph.y, acc[kk].y);
      }
    }
  }
#pragma unroll
  for (int kk = 0; kk < AO_KC; ++kk)
    if (k0 + kk < nk) out[((long)(k0 + kk) * npts + g) * nao + mu] = acc[kk];
}

}  // namespace isdf

using namespace isdf;

extern "C" int isdf_ao_desc_bytes(void) { return (int)sizeof(AoDesc); }

// coords [npts][3] f64, aos [nao] AoDesc (layout: 5 doubles, 3 doubles, 9 ints, 1 int), images [nimg][3] f64,
// kphase [nk][nimg] c128 = exp(i k.T): all device.  out [nk][npts][nao] c128.
extern "C" int isdf_eval_ao(void* hv, const double* coords, long npts, const void* aos, int nao, const double* images,
                            int nimg, const void* kphase, int nk, void* out, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, coords && aos && images && kphase && out, "null pointer");
  ISDF_CHECK_ARG(h, npts >= 0 && nao >= 1 && nimg >= 1 && nk >= 1, "shape");
  if (npts == 0) return ISDF_OK;
  const long tot = npts * nao;
  dim3 grid((unsigned)((tot + 127) / 128), (nk + AO_KC - 1) / AO_KC);
  ao_eval_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(coords, npts, (const AoDesc*)aos, nao, images, nimg,
                                                         (const cplx*)kphase, nk, (cplx*)out);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
