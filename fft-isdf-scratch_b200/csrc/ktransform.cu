// k <-> R transform + Hadamard square + k <-> R transform  (SURVEY K3/K4 epilogue).
//
// Replaces, for a block of elements e=(g,I):
//     s   = phase @ v_k            (/root/reference/fftisdf.py:41 and :79)
//     assert |Im s| < 1e-10        (:43, :81)   -> reported through diag[0..1]
//     y_s = s * s                  (:45, :83)
//     out = phase^H @ y_s   (:46, metric)   or   phase^T @ y_s   (:84, right-hand side)
// For the Gamma-centred regular k-mesh the reference always uses (fftisdf.py:322), the
// Bloch phase matrix factors as P = U1 (x) U2 (x) U3 with U_a[m,j] = exp(2 pi i m j/N_a)/sqrt(N_a),
// so each transform is three passes of tiny dense N_a-point DFTs held in shared memory.
// The result is written with arbitrary output strides (transposed for the RHS: Y^T[q][I][g]),
// optionally only for selected q (time-reversal partners skipped) and with the I index
// scattered through a per-q row map (pivot order of the truncated Cholesky factor; -1 drops).
#include "common.cuh"

namespace isdf {

constexpr int KT_NMAX = 8;     // max k-mesh points per axis
constexpr int KT_IT = 8;       // tile width along the input's contiguous index
constexpr int KT_THREADS = 256;

struct KtParams {
  const cplx* in; long in_sk; long in_sg;  // in[k*in_sk + g*in_sg + i]
  cplx* out; long out_sq; long out_sg; long out_si;  // out[slot*out_sq + g*out_sg + irow*out_si]
  long out_g0;                  // offset added to g on output (grid block origin)
  int ng, ni;                   // extents of the (g, i) element block
  int n1, n2, n3;               // k-mesh
  int gt;                       // tile height along g
  int conj2;                    // second transform uses conj(U) (metric) or U (rhs)
  int out_g_fast;               // 1: g is the unit-stride output index
  const cplx* uax;              // 3 * KT_NMAX*KT_NMAX packed U_a (row-major [m][j])
  const int* qslot;             // [nk] output slot per q, -1 = skip (may be null = identity)
  const int* rowmap; long rowmap_sq;  // [nslot][ni] output row per i, -1 = drop (may be null)
  double* diag;                 // diag[0] = max |Im s|, diag[1] = max |Re s|
  // mode 0: y = Re(s)^2 (metric / right-hand side).  mode 1: y = scale * Re(s) * table[R][g][i] (exchange:
  // vs = ws * rhos^T, fftisdf.py:215-223).  mode 2: write scale * Re(s) as a REAL table out[R][g][i] and stop
  // (ws = Re(phase @ wq) sqrt(nk), fftisdf.py:205-207).
  int mode; const double* table; long tab_sk; long tab_sg; double scale;
};

// one pass of N-point DFTs along one k-mesh axis, in place in shared memory.
// s[k][e] with k=(j1,j2,j3) C-order; element pitch EP.
template <bool CONJ>
__device__ __forceinline__ void kt_axis_pass(cplx* s, const cplx* U, int N, int stride_axis, int nk, int ne, int EP) {
  // lines: all (k with axis index 0) x e
  const int nlines_k = nk / N;
  const int total = nlines_k * ne;
  for (int w = threadIdx.x; w < total; w += KT_THREADS) {
    const int e = w % ne;
    const int lk = w / ne;
    // decompose lk into (outer, inner) around the axis
    const int inner = lk % stride_axis;
    const int outer = lk / stride_axis;
    const int kbase = outer * stride_axis * N + inner;
    cplx x[KT_NMAX];
#pragma unroll
    for (int j = 0; j < KT_NMAX; ++j)
      if (j < N) x[j] = s[(long)(kbase + j * stride_axis) * EP + e];
#pragma unroll
    for (int m = 0; m < KT_NMAX; ++m) {
      if (m < N) {
        cplx acc = make_double2(0.0, 0.0);
#pragma unroll
        for (int j = 0; j < KT_NMAX; ++j) {
          if (j < N) {
            cplx u = U[m * KT_NMAX + j];
            if (CONJ) u.y = -u.y;
            cfma(acc, u, x[j]);
          }
        }
        s[(long)(kbase + m * stride_axis) * EP + e] = acc;
      }
    }
  }
}

__global__ void __launch_bounds__(KT_THREADS) ktransform_square_kernel(KtParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int nk = p.n1 * p.n2 * p.n3;
  const int GT = p.gt;
  const int EP = GT * (KT_IT + 1);  // padded element pitch per k
  cplx* s = reinterpret_cast<cplx*>(smem_raw);
  cplx* U = s + (long)nk * EP;      // 3 * NMAX*NMAX

  for (int i = threadIdx.x; i < 3 * KT_NMAX * KT_NMAX; i += KT_THREADS) U[i] = p.uax[i];

  const int i0 = blockIdx.x * KT_IT;
  const int g0 = blockIdx.y * GT;
  const int gcnt = min(GT, p.ng - g0);
  const int icnt = min(KT_IT, p.ni - i0);

  // ---- load tile: for each k, gcnt rows of icnt contiguous elements
  const int per_k = GT * KT_IT;
  for (int w = threadIdx.x; w < nk * per_k; w += KT_THREADS) {
    const int k = w / per_k;
    const int r = w % per_k;
    const int gg = r / KT_IT, ii = r % KT_IT;
    cplx v = make_double2(0.0, 0.0);
    if (gg < gcnt && ii < icnt) v = p.in[(long)k * p.in_sk + (long)(g0 + gg) * p.in_sg + (i0 + ii)];
    s[(long)k * EP + gg * (KT_IT + 1) + ii] = v;
  }
  __syncthreads();

  // element count incl. padding column (harmless extra work on a zero column is avoided by using
  // the padded pitch only for addressing): iterate e over the padded range but skip pad slots
  const int ne = EP;  // pad slots hold garbage-free zeros? they are never written -> treat separately
  // zero the pad slots once so the passes can run over the dense range [0, EP)
  for (int w = threadIdx.x; w < nk * GT; w += KT_THREADS) {
    const int k = w / GT, gg = w % GT;
    s[(long)k * EP + gg * (KT_IT + 1) + KT_IT] = make_double2(0.0, 0.0);
  }
  __syncthreads();

  const cplx* U1 = U;
  const cplx* U2 = U + KT_NMAX * KT_NMAX;
  const cplx* U3 = U + 2 * KT_NMAX * KT_NMAX;

  // ---- first transform: s = P v   (U, no conjugation)
  if (p.n3 > 1) { kt_axis_pass<false>(s, U3, p.n3, 1, nk, ne, EP); __syncthreads(); }
  if (p.n2 > 1) { kt_axis_pass<false>(s, U2, p.n2, p.n3, nk, ne, EP); __syncthreads(); }
  if (p.n1 > 1) { kt_axis_pass<false>(s, U1, p.n1, p.n2 * p.n3, nk, ne, EP); __syncthreads(); }

  // ---- reality check + square
  double mx_im = 0.0, mx_re = 0.0;
  for (int w = threadIdx.x; w < nk * EP; w += KT_THREADS) {
    cplx v = s[w];
    mx_im = fmax(mx_im, fabs(v.y));
    mx_re = fmax(mx_re, fabs(v.x));
    if (p.mode == 0) {
      s[w] = make_double2(v.x * v.x, 0.0);
    } else {
      const int R = w / EP, r = w - R * EP;
      const int gg = r / (KT_IT + 1), ii = r - gg * (KT_IT + 1);
      double y = 0.0;
      if (gg < gcnt && ii < icnt) {
        if (p.mode == 1) {
          y = p.scale * v.x * p.table[(long)R * p.tab_sk + (long)(g0 + gg) * p.tab_sg + (i0 + ii)];
        } else {   // mode 2: real output table, no second transform
          reinterpret_cast<double*>(p.out)[(long)R * p.out_sq + (p.out_g0 + g0 + gg) * p.out_sg + (long)(i0 + ii) * p.out_si] =
              p.scale * v.x;
        }
      }
      s[w] = make_double2(y, 0.0);
    }
  }
  if (p.diag != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx_im = fmax(mx_im, __shfl_xor_sync(0xffffffffu, mx_im, o));
      mx_re = fmax(mx_re, __shfl_xor_sync(0xffffffffu, mx_re, o));
    }
    if ((threadIdx.x & 31) == 0) {
      atomic_max_nonneg(p.diag + 0, mx_im);
      atomic_max_nonneg(p.diag + 1, mx_re);
    }
  }
  __syncthreads();
  if (p.mode == 2) return;

  // ---- second transform
  if (p.conj2) {
    if (p.n3 > 1) { kt_axis_pass<true>(s, U3, p.n3, 1, nk, ne, EP); __syncthreads(); }
    if (p.n2 > 1) { kt_axis_pass<true>(s, U2, p.n2, p.n3, nk, ne, EP); __syncthreads(); }
    if (p.n1 > 1) { kt_axis_pass<true>(s, U1, p.n1, p.n2 * p.n3, nk, ne, EP); __syncthreads(); }
  } else {
    if (p.n3 > 1) { kt_axis_pass<false>(s, U3, p.n3, 1, nk, ne, EP); __syncthreads(); }
    if (p.n2 > 1) { kt_axis_pass<false>(s, U2, p.n2, p.n3, nk, ne, EP); __syncthreads(); }
    if (p.n1 > 1) { kt_axis_pass<false>(s, U1, p.n1, p.n2 * p.n3, nk, ne, EP); __syncthreads(); }
  }

  // ---- store
  for (int w = threadIdx.x; w < nk * per_k; w += KT_THREADS) {
    const int q = w / per_k;
    const int r = w % per_k;
    int gg, ii;
    if (p.out_g_fast) { ii = r / GT; gg = r % GT; } else { gg = r / KT_IT; ii = r % KT_IT; }
    if (gg >= gcnt || ii >= icnt) continue;
    const int slot = p.qslot ? p.qslot[q] : q;
    if (slot < 0) continue;
    int irow = i0 + ii;
    if (p.rowmap) {
      irow = p.rowmap[(long)slot * p.rowmap_sq + irow];
      if (irow < 0) continue;
    }
    p.out[(long)slot * p.out_sq + (p.out_g0 + g0 + gg) * p.out_sg + (long)irow * p.out_si] =
        s[(long)q * EP + gg * (KT_IT + 1) + ii];
  }
}

}  // namespace isdf

using namespace isdf;

extern "C" int isdf_ktransform_ex(void* hv, const void* in, long in_sk, long in_sg, void* out, long out_sq,
                                  long out_sg, long out_si, long out_g0, int ng, int ni, const int* kmesh,
                                  const void* uaxes_dev, int conj2, int out_g_fast, const int* qslot_dev,
                                  const int* rowmap_dev, long rowmap_sq, double* diag_dev, int mode,
                                  const double* table, long tab_sk, long tab_sg, double scale, void* stream);

extern "C" int isdf_ktransform_square(void* hv, const void* in, long in_sk, long in_sg, void* out, long out_sq,
                                      long out_sg, long out_si, long out_g0, int ng, int ni, const int* kmesh,
                                      const void* uaxes_dev, int conj2, int out_g_fast, const int* qslot_dev,
                                      const int* rowmap_dev, long rowmap_sq, double* diag_dev, void* stream) {
  return isdf_ktransform_ex(hv, in, in_sk, in_sg, out, out_sq, out_sg, out_si, out_g0, ng, ni, kmesh, uaxes_dev, conj2,
                            out_g_fast, qslot_dev, rowmap_dev, rowmap_sq, diag_dev, 0, nullptr, 0, 0, 1.0, stream);
}

extern "C" int isdf_ktransform_ex(void* hv, const void* in, long in_sk, long in_sg, void* out, long out_sq,
                                  long out_sg, long out_si, long out_g0, int ng, int ni, const int* kmesh,
                                  const void* uaxes_dev, int conj2, int out_g_fast, const int* qslot_dev,
                                  const int* rowmap_dev, long rowmap_sq, double* diag_dev, int mode,
                                  const double* table, long tab_sk, long tab_sg, double scale, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, mode >= 0 && mode <= 2 && (mode != 1 || table != nullptr), "mode/table");
  ISDF_CHECK_ARG(h, in && out && kmesh && uaxes_dev, "null pointer");
  ISDF_CHECK_ARG(h, kmesh[0] >= 1 && kmesh[1] >= 1 && kmesh[2] >= 1, "kmesh");
  ISDF_CHECK_ARG(h, kmesh[0] <= KT_NMAX && kmesh[1] <= KT_NMAX && kmesh[2] <= KT_NMAX, "kmesh axis > 8 unsupported");
  if (ng <= 0 || ni <= 0) return ISDF_OK;
  const int nk = kmesh[0] * kmesh[1] * kmesh[2];
  // pick the tallest tile that fits in ~200 KB of shared memory
  int gt = 16;
  auto bytes = [&](int g) { return (size_t)nk * g * (KT_IT + 1) * sizeof(cplx) + 3 * KT_NMAX * KT_NMAX * sizeof(cplx); };
  while (gt > 1 && bytes(gt) > (size_t)200 * 1024) gt >>= 1;
  ISDF_CHECK_ARG(h, bytes(gt) <= (size_t)h->max_smem_optin, "k-mesh too large for the shared-memory tile");
  KtParams p;
  p.in = (const cplx*)in; p.in_sk = in_sk; p.in_sg = in_sg;
  p.out = (cplx*)out; p.out_sq = out_sq; p.out_sg = out_sg; p.out_si = out_si; p.out_g0 = out_g0;
  p.ng = ng; p.ni = ni; p.n1 = kmesh[0]; p.n2 = kmesh[1]; p.n3 = kmesh[2];
  p.gt = gt; p.conj2 = conj2; p.out_g_fast = out_g_fast;
  p.uax = (const cplx*)uaxes_dev; p.qslot = qslot_dev; p.rowmap = rowmap_dev; p.rowmap_sq = rowmap_sq;
  p.diag = diag_dev;
  p.mode = mode; p.table = table; p.tab_sk = tab_sk; p.tab_sg = tab_sg; p.scale = scale;
  const size_t smem = bytes(gt);
  ISDF_CUDA(h, cudaFuncSetAttribute(ktransform_square_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((ni + KT_IT - 1) / KT_IT, (ng + gt - 1) / gt);
  ISDF_CHECK_ARG(h, grid.y <= 65535, "too many g tiles; split the block");
  ktransform_square_kernel<<<grid, KT_THREADS, smem, (cudaStream_t)stream>>>(p);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
