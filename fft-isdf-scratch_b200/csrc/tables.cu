// Device-side generation of the per-q tables of the Coulomb stage (no host loops, no H2D):
//   * post-weight  sqrt(v(q+G) vol)/ng  with PySCF's get_coulG(cell, k=q, mesh) semantics for
//     exxdiv=None, wrap_around=True  (/root/reference/fftisdf.py:114-115; PySCF pbc/tools/pbc.py):
//     v = 4 pi/|q+G|^2, v(0) = 0, q+G folded into the first-zone box and box-boundary terms zeroed.
//   * pre-phase    fq(r) = exp(-i q.r)                                  (fftisdf.py:99)
#include "common.cuh"

namespace isdf {

struct CoulParams {
  double b[9];     // reciprocal vectors, rows
  double ks[3];    // q in units of b (scaled k-point)
  int mesh[3];
  int knz;         // |q| != 0 (PySCF: abs(k).sum() > 1e-9)
  double vol_over_ng2;  // vol / ng^2
};

__global__ void coulomb_weight_kernel(CoulParams p, double* __restrict__ out, long ng) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng) return;
  const int n3 = p.mesh[2], n2 = p.mesh[1];
  int idx[3];
  idx[2] = (int)(i % n3);
  idx[1] = (int)((i / n3) % n2);
  idx[0] = (int)(i / ((long)n3 * n2));
  double x[3];
  bool eq = false;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const int m = p.mesh[a];
    const int n = (idx[a] < (m + 1) / 2) ? idx[a] : idx[a] - m;   // numpy.fft.fftfreq order
    x[a] = (double)n + (p.knz ? p.ks[a] : 0.0);
  }
  if (p.knz) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double edge = (double)(p.mesh[a] / 2) + 0.5;
      double red = x[a] / edge;
      red = rint(red * 1e9) / 1e9;          // numpy .round(9)
      eq = eq || (red == 1.0) || (red == -1.0);
      const int on = (int)red;              // astype(int): truncation toward zero
      if (on == 1) x[a] -= 2.0 * edge;
      else if (on == -1) x[a] += 2.0 * edge;
    }
  }
  const double gx = x[0] * p.b[0] + x[1] * p.b[3] + x[2] * p.b[6];
  const double gy = x[0] * p.b[1] + x[1] * p.b[4] + x[2] * p.b[7];
  const double gz = x[0] * p.b[2] + x[1] * p.b[5] + x[2] * p.b[8];
  const double g2 = gx * gx + gy * gy + gz * gz;
  double v = (g2 == 0.0 || eq) ? 0.0 : 4.0 * 3.14159265358979323846 / g2;
  out[i] = sqrt(v * p.vol_over_ng2);
}

__global__ void phase_table_kernel(const double* __restrict__ coords, double qx, double qy, double qz,
                                   cplx* __restrict__ out, long ng) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ng) return;
  const double t = coords[3 * i] * qx + coords[3 * i + 1] * qy + coords[3 * i + 2] * qz;
  double s, c;
  sincos(-t, &s, &c);
  out[i] = make_double2(c, s);
}

}  // namespace isdf

using namespace isdf;

// out[G] = sqrt(v(q+G) * vol) / ng,  G in numpy fftfreq / C order over mesh.  b, kscaled: host arrays.
extern "C" int isdf_coulomb_weights(void* hv, const double* b_host, const double* kscaled_host, const int* mesh,
                                    double vol, double* out_dev, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, b_host && kscaled_host && mesh && out_dev, "null pointer");
  CoulParams p;
  for (int i = 0; i < 9; ++i) p.b[i] = b_host[i];
  double kabs = 0.0;
  for (int a = 0; a < 3; ++a) {
    p.ks[a] = kscaled_host[a];
    p.mesh[a] = mesh[a];
    // |k| in cartesian components, as PySCF tests abs(k).sum()
  }
  for (int c = 0; c < 3; ++c) {
    double kc = 0.0;
    for (int a = 0; a < 3; ++a) kc += kscaled_host[a] * b_host[3 * a + c];
    kabs += fabs(kc);
  }
  p.knz = kabs > 1e-9;
  const long ng = (long)mesh[0] * mesh[1] * mesh[2];
  p.vol_over_ng2 = vol / ((double)ng * (double)ng);
  coulomb_weight_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, out_dev, ng);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

// out[i] = exp(-i q . coords[i]),  coords [ng][3] device, q host (cartesian).
extern "C" int isdf_phase_table(void* hv, const double* coords_dev, const double* q_host, long ng, void* out_dev,
                                void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, coords_dev && q_host && out_dev && ng >= 0, "args");
  if (ng == 0) return ISDF_OK;
  phase_table_kernel<<<(unsigned)((ng + 255) / 256), 256, 0, (cudaStream_t)stream>>>(coords_dev, q_host[0], q_host[1],
                                                                                  q_host[2], (cplx*)out_dev, ng);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
