// Register-resident variant of the k<->R transform / square / k<->R transform for small k-meshes
// (every axis <= 4): one thread (nk < 27) or three / four lanes (27 <= nk <= 64) own all nk values of one element, so there is no shared
// memory, no barrier and the input/output accesses are plain coalesced streams.  Same arithmetic and
// reference lines as ktransform.cu (fftisdf.py:41-47, :79-85).  Input is [nk][rows][cols] with cols
// contiguous; for the right-hand side the GEMM already produces the transposed fx^T[k][i][g], so the
// output Y^T[q][row(i)][g] needs no transpose either.
#include "common.cuh"

namespace isdf {

__constant__ cplx c_uax[3][8][8];  // U_a[m][j], a = 0,1,2 (pitch 8, as packed by the host)

struct KtRegParams {
  const cplx* in; long in_sk; long in_sr;      // in[k*in_sk + r*in_sr + c]
  cplx* out; long out_sq; long out_sr; long out_c0;  // out[slot*out_sq + row*out_sr + out_c0 + c]
  int nrows, ncols;
  int conj2;
  const int* qslot;                  // [nk] or null
  const int* rowmap; long rowmap_sq; // [nslot][nrows] or null
  double* diag;
  // mode 0: y = s*s (metric / right-hand side).  mode 1: y = scale * Re(s) * table[R][r][c] (exchange build,
  // fftisdf.py:215-223).  mode 2: write scale * Re(s) as a real table [R][r][c] and stop (fftisdf.py:205-207).
  int mode; const double* table; long tab_sk; long tab_sr; double scale;
};

template <int N, int STRIDE, int NK, int AX, bool CONJ>
__device__ __forceinline__ void dft_axis(cplx (&x)[NK]) {
  if constexpr (N > 1) {
#pragma unroll
  for (int base = 0; base < NK; ++base) {
    // lines start where the axis index is 0:  (base / STRIDE) % N == 0
    if ((base / STRIDE) % N != 0) continue;
    cplx t[N];
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = x[base + j * STRIDE];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      cplx acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        cplx u = c_uax[AX][m][j];
        if (CONJ) u.y = -u.y;
        cfma(acc, u, t[j]);
      }
      x[base + m * STRIDE] = acc;
    }
  }
  }
}

// First pass of the second transform: the input is real (y_s = s*s), half the FMAs.
template <int N, int STRIDE, int NK, int AX, bool CONJ>
__device__ __forceinline__ void dft_axis_real_in(cplx (&x)[NK]) {
  if constexpr (N > 1) {
#pragma unroll
  for (int base = 0; base < NK; ++base) {
    if ((base / STRIDE) % N != 0) continue;
    double t[N];
#pragma unroll
    for (int j = 0; j < N; ++j) t[j] = x[base + j * STRIDE].x;
#pragma unroll
    for (int m = 0; m < N; ++m) {
      double re = 0.0, im = 0.0;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        const cplx u = c_uax[AX][m][j];
        re = fma(u.x, t[j], re);
        im = fma(CONJ ? -u.y : u.y, t[j], im);
      }
      x[base + m * STRIDE] = make_double2(re, im);
    }
  }
  }
}

// Columns per thread: tiny k-meshes (nk <= 4, e.g. the Gamma point, where the "transform" is just the square) leave a
// thread with one or two 16-byte loads in flight -- the stage is then latency-bound at ~0.1 of the HBM rate; several
// independent column chunks per thread restore the memory-level parallelism.
__host__ __device__ constexpr int kt_cpt(int nk) { return nk >= 8 ? 1 : (nk >= 4 ? 2 : (nk >= 2 ? 4 : 8)); }

template <int N1, int N2, int N3>
__global__ void __launch_bounds__(128) ktransform_reg_kernel(KtRegParams p) {
  constexpr int NK = N1 * N2 * N3;
  constexpr int CPT = kt_cpt(NK);
  const int r = blockIdx.y;
  double mx_im = 0.0, mx_re = 0.0;
  // all loads first (the stores below may alias the input as far as the compiler knows)
  cplx xs[CPT][NK];
#pragma unroll
  for (int cu = 0; cu < CPT; ++cu) {
    const int c = (blockIdx.x * CPT + cu) * 128 + threadIdx.x;
    const cplx* src = p.in + (long)r * p.in_sr + c;
#pragma unroll
    for (int k = 0; k < NK; ++k) xs[cu][k] = (c < p.ncols) ? src[(long)k * p.in_sk] : make_double2(0.0, 0.0);
  }
#pragma unroll
  for (int cu = 0; cu < CPT; ++cu) {
  const int c = (blockIdx.x * CPT + cu) * 128 + threadIdx.x;
  if (c < p.ncols) {
    cplx (&x)[NK] = xs[cu];
    dft_axis<N3, 1, NK, 2, false>(x);
    dft_axis<N2, N3, NK, 1, false>(x);
    dft_axis<N1, N2 * N3, NK, 0, false>(x);
#pragma unroll
    for (int k = 0; k < NK; ++k) {
      mx_im = fmax(mx_im, fabs(x[k].y));
      mx_re = fmax(mx_re, fabs(x[k].x));
      if (p.mode == 0) x[k] = make_double2(x[k].x * x[k].x, 0.0);
      else if (p.mode == 1) x[k] = make_double2(p.scale * x[k].x * p.table[(long)k * p.tab_sk + (long)r * p.tab_sr + c], 0.0);
      else reinterpret_cast<double*>(p.out)[(long)k * p.out_sq + (long)r * p.out_sr + p.out_c0 + c] = p.scale * x[k].x;
    }
    // second transform: its first non-trivial pass sees real input (y_s is real) -> half the FMAs there
    constexpr bool R3 = N3 > 1, R2 = !R3 && N2 > 1, R1 = !R3 && !R2;
    if (p.mode == 2) {
    } else if (p.conj2) {
      if (R3) dft_axis_real_in<N3, 1, NK, 2, true>(x); else dft_axis<N3, 1, NK, 2, true>(x);
      if (R2) dft_axis_real_in<N2, N3, NK, 1, true>(x); else dft_axis<N2, N3, NK, 1, true>(x);
      if (R1) dft_axis_real_in<N1, N2 * N3, NK, 0, true>(x); else dft_axis<N1, N2 * N3, NK, 0, true>(x);
    } else {
      if (R3) dft_axis_real_in<N3, 1, NK, 2, false>(x); else dft_axis<N3, 1, NK, 2, false>(x);
      if (R2) dft_axis_real_in<N2, N3, NK, 1, false>(x); else dft_axis<N2, N3, NK, 1, false>(x);
      if (R1) dft_axis_real_in<N1, N2 * N3, NK, 0, false>(x); else dft_axis<N1, N2 * N3, NK, 0, false>(x);
    }
#pragma unroll
    for (int q = 0; q < NK; ++q) {
      if (p.mode == 2) continue;
      const int slot = p.qslot ? p.qslot[q] : q;
      if (slot < 0) continue;
      int row = r;
      if (p.rowmap) {
        row = p.rowmap[(long)slot * p.rowmap_sq + r];
        if (row < 0) continue;
      }
      p.out[(long)slot * p.out_sq + (long)row * p.out_sr + p.out_c0 + c] = x[q];
    }
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
}

// Lane-split variant: NS = 3 or 4 lanes (the length of the split axis SA) share an element, each holding the
// slab with split-axis index `sub`; the two local axes are transformed in registers, the split axis with warp
// shuffles.  Lane = sub*E + e with E = 32/NS elements per warp (NS = 3 leaves two lanes idle), so the E elements
// of a warp stay contiguous in memory per slab.  A third / quarter of the registers of the one-thread kernel,
// i.e. 3-4x the resident warps and loads in flight: the stage is latency-bound, not flop-bound.
template <int N1, int N2, int N3, int SA>
struct KtSplit {
  static constexpr int NS = (SA == 0) ? N1 : (SA == 1) ? N2 : N3;
  static constexpr int E = 32 / NS;
  static constexpr int NA = (SA == 0) ? N2 : N1;
  static constexpr int NB = (SA == 2) ? N2 : N3;
  static constexpr int NL = NA * NB;
};

template <int N1, int N2, int N3, int SA, bool CONJ>
__device__ __forceinline__ void kt_split_transform(cplx (&x)[KtSplit<N1, N2, N3, SA>::NL], int sub, int e) {
  using T = KtSplit<N1, N2, N3, SA>;
  constexpr int NA = T::NA, NB = T::NB, NL = T::NL, NS = T::NS, E = T::E;
  constexpr int AXA = (SA == 0) ? 1 : 0;
  constexpr int AXB = (SA == 2) ? 1 : 2;
  dft_axis<NB, 1, NL, AXB, CONJ>(x);
  dft_axis<NA, NB, NL, AXA, CONJ>(x);
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    cplx acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      cplx v;
      v.x = __shfl_sync(0xffffffffu, x[l].x, j * E + e);
      v.y = __shfl_sync(0xffffffffu, x[l].y, j * E + e);
      cplx u = c_uax[SA][sub][j];
      if (CONJ) u.y = -u.y;
      cfma(acc, u, v);
    }
    x[l] = acc;
  }
}

template <int N1, int N2, int N3, int SA>
__global__ void __launch_bounds__(128) ktransform_split_kernel(KtRegParams p) {
  using T = KtSplit<N1, N2, N3, SA>;
  static_assert(T::NS == 3 || T::NS == 4, "split axis must have length 3 or 4");
  constexpr int NB = T::NB, NL = T::NL, NS = T::NS, E = T::E;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool lane_on = lane < NS * E;                 // NS = 3: lanes 30, 31 only take part in the shuffles
  const int sub = lane_on ? lane / E : NS - 1, e = lane_on ? lane % E : E - 1;
  const int c = (blockIdx.x * 4 + warp) * E + e;
  const int r = blockIdx.y;
  const bool valid = lane_on && c < p.ncols;
  auto kof = [&](int l) {
    const int a = l / NB, b = l % NB;
    return (SA == 0) ? ((sub * N2 + a) * N3 + b) : (SA == 1) ? ((a * N2 + sub) * N3 + b) : ((a * N2 + b) * N3 + sub);
  };
  cplx x[NL];
  const cplx* src = p.in + (long)r * p.in_sr + (valid ? c : 0);
#pragma unroll
  for (int l = 0; l < NL; ++l) x[l] = valid ? src[(long)kof(l) * p.in_sk] : make_double2(0.0, 0.0);
  kt_split_transform<N1, N2, N3, SA, false>(x, sub, e);
  double mx_im = 0.0, mx_re = 0.0;
#pragma unroll
  for (int l = 0; l < NL; ++l) {
    mx_im = fmax(mx_im, fabs(x[l].y));
    mx_re = fmax(mx_re, fabs(x[l].x));
    if (p.mode == 0) {
      x[l] = make_double2(x[l].x * x[l].x, 0.0);
    } else if (p.mode == 1) {
      const double tv = valid ? p.table[(long)kof(l) * p.tab_sk + (long)r * p.tab_sr + c] : 0.0;
      x[l] = make_double2(p.scale * x[l].x * tv, 0.0);
    } else if (valid) {
      reinterpret_cast<double*>(p.out)[(long)kof(l) * p.out_sq + (long)r * p.out_sr + p.out_c0 + c] = p.scale * x[l].x;
    }
  }
  if (p.mode != 2) {
    if (p.conj2) kt_split_transform<N1, N2, N3, SA, true>(x, sub, e);
    else kt_split_transform<N1, N2, N3, SA, false>(x, sub, e);
  }
  if (valid && p.mode != 2) {
#pragma unroll
    for (int l = 0; l < NL; ++l) {
      const int q = kof(l);
      const int slot = p.qslot ? p.qslot[q] : q;
      if (slot < 0) continue;
      int row = r;
      if (p.rowmap) {
        row = p.rowmap[(long)slot * p.rowmap_sq + r];
        if (row < 0) continue;
      }
      p.out[(long)slot * p.out_sq + (long)row * p.out_sr + p.out_c0 + c] = x[l];
    }
  }
  if (p.diag != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mx_im = fmax(mx_im, __shfl_xor_sync(0xffffffffu, mx_im, o));
      mx_re = fmax(mx_re, __shfl_xor_sync(0xffffffffu, mx_re, o));
    }
    if (lane == 0) {
      atomic_max_nonneg(p.diag + 0, mx_im);
      atomic_max_nonneg(p.diag + 1, mx_re);
    }
  }
}

template <int N1, int N2, int N3, int SA>
static cudaError_t launch_split(const KtRegParams& p, cudaStream_t st) {
  constexpr int CB = 4 * KtSplit<N1, N2, N3, SA>::E;   // columns per 4-warp block
  dim3 grid((p.ncols + CB - 1) / CB, p.nrows);
  ktransform_split_kernel<N1, N2, N3, SA><<<grid, 128, 0, st>>>(p);
  return cudaGetLastError();
}

template <int N1, int N2, int N3>
static cudaError_t launch_reg(const KtRegParams& p, cudaStream_t st) {
  constexpr int CPT = kt_cpt(N1 * N2 * N3);
  dim3 grid((p.ncols + 128 * CPT - 1) / (128 * CPT), p.nrows);
  ktransform_reg_kernel<N1, N2, N3><<<grid, 128, 0, st>>>(p);
  return cudaGetLastError();
}

}  // namespace isdf

using namespace isdf;

#define KT_CASE(a, b, c) \
  if (n1 == a && n2 == b && n3 == c) { e = launch_reg<a, b, c>(p, st); hit = true; }

// Returns ISDF_ESIZE (-2) without launching when the mesh has no register instantiation; the caller
// then uses isdf_ktransform_square (shared-memory kernel).
extern "C" int isdf_ktransform_rows_ex(void* hv, const void* in, long in_sk, long in_sr, void* out, long out_sq,
                                       long out_sr, long out_c0, int nrows, int ncols, const int* kmesh,
                                       const void* uaxes_host, int conj2, const int* qslot, const int* rowmap,
                                       long rowmap_sq, double* diag, int mode, const double* table, long tab_sk,
                                       long tab_sr, double scale, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, in && out && kmesh && uaxes_host, "null pointer");
  const int n1 = kmesh[0], n2 = kmesh[1], n3 = kmesh[2];
  ISDF_CHECK_ARG(h, n1 >= 1 && n2 >= 1 && n3 >= 1, "kmesh");
  if (n1 > 4 || n2 > 4 || n3 > 4) return ISDF_ESIZE;
  ISDF_CHECK_ARG(h, nrows <= 65535, "too many rows for one launch");
  if (nrows <= 0 || ncols <= 0) return ISDF_OK;
  ISDF_CUDA(h, cudaMemcpyToSymbolAsync(c_uax, uaxes_host, sizeof(cplx) * 3 * 64, 0, cudaMemcpyHostToDevice, st));
  KtRegParams p;
  p.in = (const cplx*)in; p.in_sk = in_sk; p.in_sr = in_sr;
  p.out = (cplx*)out; p.out_sq = out_sq; p.out_sr = out_sr; p.out_c0 = out_c0;
  p.nrows = nrows; p.ncols = ncols; p.conj2 = conj2;
  p.qslot = qslot; p.rowmap = rowmap; p.rowmap_sq = rowmap_sq; p.diag = diag;
  ISDF_CHECK_ARG(h, mode >= 0 && mode <= 2 && (mode != 1 || table != nullptr), "mode/table");
  p.mode = mode; p.table = table; p.tab_sk = tab_sk; p.tab_sr = tab_sr; p.scale = scale;
  cudaError_t e = cudaSuccess;
  bool hit = false;
  KT_CASE(1, 1, 1) KT_CASE(1, 1, 2) KT_CASE(1, 2, 1) KT_CASE(2, 1, 1) KT_CASE(1, 2, 2) KT_CASE(2, 1, 2)
  KT_CASE(2, 2, 1) KT_CASE(2, 2, 2) KT_CASE(1, 1, 3) KT_CASE(1, 3, 1) KT_CASE(3, 1, 1) KT_CASE(1, 3, 3)
  KT_CASE(3, 1, 3) KT_CASE(3, 3, 1) KT_CASE(2, 2, 3) KT_CASE(2, 3, 2) KT_CASE(3, 2, 2)
  KT_CASE(2, 3, 3) KT_CASE(3, 2, 3) KT_CASE(3, 3, 2) KT_CASE(1, 2, 3) KT_CASE(1, 3, 2) KT_CASE(2, 1, 3)
  KT_CASE(2, 3, 1) KT_CASE(3, 1, 2) KT_CASE(3, 2, 1) KT_CASE(1, 1, 4) KT_CASE(1, 4, 1) KT_CASE(4, 1, 1)
  KT_CASE(2, 2, 4) KT_CASE(2, 4, 2) KT_CASE(4, 2, 2) KT_CASE(1, 4, 4) KT_CASE(4, 1, 4) KT_CASE(4, 4, 1)
  KT_CASE(2, 4, 4) KT_CASE(4, 2, 4) KT_CASE(4, 4, 2)
  KT_CASE(1, 2, 4) KT_CASE(1, 4, 2) KT_CASE(2, 1, 4) KT_CASE(2, 4, 1) KT_CASE(4, 1, 2) KT_CASE(4, 2, 1)
  KT_CASE(1, 3, 4) KT_CASE(1, 4, 3) KT_CASE(3, 1, 4) KT_CASE(3, 4, 1) KT_CASE(4, 1, 3) KT_CASE(4, 3, 1)
  KT_CASE(2, 3, 4) KT_CASE(2, 4, 3) KT_CASE(3, 2, 4) KT_CASE(3, 4, 2) KT_CASE(4, 2, 3) KT_CASE(4, 3, 2)
#define KT_SPLIT(a, b, c, sa) \
  if (n1 == a && n2 == b && n3 == c) { e = launch_split<a, b, c, sa>(p, st); hit = true; }
  KT_SPLIT(4, 4, 4, 0) KT_SPLIT(3, 4, 4, 1) KT_SPLIT(4, 3, 4, 0) KT_SPLIT(4, 4, 3, 0) KT_SPLIT(4, 3, 3, 0)
  KT_SPLIT(3, 4, 3, 1) KT_SPLIT(3, 3, 4, 2) KT_SPLIT(3, 3, 3, 0)
  if (!hit) return ISDF_ESIZE;
  ISDF_CUDA(h, e);
  return ISDF_OK;
}

extern "C" int isdf_ktransform_square_rows(void* hv, const void* in, long in_sk, long in_sr, void* out, long out_sq,
                                           long out_sr, long out_c0, int nrows, int ncols, const int* kmesh,
                                           const void* uaxes_host, int conj2, const int* qslot, const int* rowmap,
                                           long rowmap_sq, double* diag, void* stream) {
  return isdf_ktransform_rows_ex(hv, in, in_sk, in_sr, out, out_sq, out_sr, out_c0, nrows, ncols, kmesh, uaxes_host,
                                 conj2, qslot, rowmap, rowmap_sq, diag, 0, nullptr, 0, 0, 1.0, stream);
}
