// Batched 3-D complex128 FFT over the dense grid: shared-memory Stockham autosort kernels with hard-coded
// radix-2/3/4/5/7/8/11/13 butterflies, fused with the e^{-iq.r} pre-phase and the sqrt(v(q+G) vol)/ng post-weight
// (replaces pbctools.fft / get_coulG / ifft at /root/reference/fftisdf.py:113-119 -- the inverse FFT is
// removed analytically).
//
//   fft_plane_kernel : one CTA per (vector, x-plane).  The n2 x n3 plane is loaded once (coalesced 16-byte accesses,
//                      phase fused), transformed along z and then along y entirely in shared memory (the last z stage
//                      writes transposed, so the y stages see unit-stride lines again), and stored once.
//   fft_lines_kernel : tiles of 32 lines along one axis (the x pass: 512-byte coalesced segments per x; also the z / y
//                      passes of planes that do not fit in shared memory); weight fused into the store.
// Both kernels hold a tile as [point][line] with an odd pitch: a warp works on 32 consecutive LINES of the same
// butterfly, so every shared-memory access of a stage is conflict-free and the twiddle factor is warp-uniform.
// Each thread takes the R inputs of one butterfly into registers, runs the hard-coded R-point DFT, multiplies by the
// stage twiddles and scatters to the other buffer (Stockham: no bit reversal, natural order out).  Primes above 13
// fall back to a direct R-point DFT inside the same stage code; meshes made of such primes are better served by the
// tensor-core DFT (dft_dmma.cu), which kernels.py selects for axes <= 48.
// The host runs the plane pass and the x pass back to back on L2-sized groups of vectors, so that only the first read
// and the last write of every vector reach HBM.
#include <map>
#include <vector>
#include <math.h>
#include "common.cuh"

namespace isdf {

constexpr int FFT_THREADS = 256;          // lines kernel; the plane kernel runs 256 or 512 threads (blockDim.x)
constexpr int FFT_MAXSTAGES = 10;
constexpr int FFT_LB = 2;            // lines per thread (twiddle reuse)
constexpr int FFT_RMAX = 64;         // largest radix of the generic fallback

template <int R> struct PrimeTab;
template <> struct PrimeTab<3> {
  static __device__ __forceinline__ double c(int m) { constexpr double t[3] = {1.0, -0.5, -0.5}; return t[m]; }
  static __device__ __forceinline__ double s(int m) { constexpr double t[3] = {0.0, 0.8660254037844386, -0.8660254037844386}; return t[m]; }
};
template <> struct PrimeTab<5> {
  static __device__ __forceinline__ double c(int m) { constexpr double t[5] = {1.0, 0.30901699437494745, -0.8090169943749475, -0.8090169943749475, 0.30901699437494745}; return t[m]; }
  static __device__ __forceinline__ double s(int m) { constexpr double t[5] = {0.0, 0.9510565162951535, 0.5877852522924731, -0.5877852522924731, -0.9510565162951535}; return t[m]; }
};
template <> struct PrimeTab<7> {
  static __device__ __forceinline__ double c(int m) { constexpr double t[7] = {1.0, 0.6234898018587335, -0.2225209339563144, -0.9009688679024191, -0.9009688679024191, -0.2225209339563144, 0.6234898018587335}; return t[m]; }
  static __device__ __forceinline__ double s(int m) { constexpr double t[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.4338837391175581, -0.4338837391175581, -0.9749279121818236, -0.7818314824680298}; return t[m]; }
};
template <> struct PrimeTab<11> {
  static __device__ __forceinline__ double c(int m) { constexpr double t[11] = {1.0, 0.8412535328311812, 0.41541501300188644, -0.14231483827328514, -0.6548607339452851, -0.9594929736144974, -0.9594929736144974, -0.6548607339452851, -0.14231483827328514, 0.41541501300188644, 0.8412535328311812}; return t[m]; }
  static __device__ __forceinline__ double s(int m) { constexpr double t[11] = {0.0, 0.5406408174555976, 0.9096319953545183, 0.9898214418809327, 0.7557495743542583, 0.28173255684142967, -0.28173255684142967, -0.7557495743542583, -0.9898214418809327, -0.9096319953545183, -0.5406408174555976}; return t[m]; }
};
template <> struct PrimeTab<13> {
  static __device__ __forceinline__ double c(int m) { constexpr double t[13] = {1.0, 0.8854560256532099, 0.5680647467311558, 0.12053668025532305, -0.3546048870425356, -0.7485107481711011, -0.970941817426052, -0.970941817426052, -0.7485107481711011, -0.3546048870425356, 0.12053668025532305, 0.5680647467311558, 0.8854560256532099}; return t[m]; }
  static __device__ __forceinline__ double s(int m) { constexpr double t[13] = {0.0, 0.46472317204376856, 0.8229838658936564, 0.992708874098054, 0.9350162426854148, 0.6631226582407952, 0.23931566428755777, -0.23931566428755777, -0.6631226582407952, -0.9350162426854148, -0.992708874098054, -0.8229838658936564, -0.46472317204376856}; return t[m]; }
};

// ---- R-point DFTs in registers, forward sign e^{-2 pi i jk/R}, natural order in and out ----------------------------
__device__ __forceinline__ cplx mul_mi(cplx a) { return make_double2(a.y, -a.x); }   // a * (-i)

template <int R>
__device__ __forceinline__ void dft_prime(cplx (&x)[R]) {
  constexpr int H = (R - 1) / 2;
  cplx s[H], d[H];
#pragma unroll
  for (int j = 0; j < H; ++j) { s[j] = cadd(x[j + 1], x[R - 1 - j]); d[j] = csub(x[j + 1], x[R - 1 - j]); }
  cplx y0 = x[0];
#pragma unroll
  for (int j = 0; j < H; ++j) y0 = cadd(y0, s[j]);
  cplx out[R];
  out[0] = y0;
#pragma unroll
  for (int r = 1; r <= H; ++r) {
    double ar = x[0].x, ai = x[0].y, br = 0.0, bi = 0.0;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      const double c = PrimeTab<R>::c((j * r) % R), sn = PrimeTab<R>::s((j * r) % R);
      ar = fma(c, s[j - 1].x, ar); ai = fma(c, s[j - 1].y, ai);
      br = fma(sn, d[j - 1].x, br); bi = fma(sn, d[j - 1].y, bi);
    }
    // y_r = A - i B,  y_{R-r} = A + i B
    out[r] = make_double2(ar + bi, ai - br);
    out[R - r] = make_double2(ar - bi, ai + br);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = out[r];
}

template <int R> __device__ __forceinline__ void dft_small(cplx (&x)[R]);
template <> __device__ __forceinline__ void dft_small<2>(cplx (&x)[2]) {
  const cplx a = x[0], b = x[1];
  x[0] = cadd(a, b); x[1] = csub(a, b);
}
template <> __device__ __forceinline__ void dft_small<3>(cplx (&x)[3]) { dft_prime<3>(x); }
template <> __device__ __forceinline__ void dft_small<4>(cplx (&x)[4]) {
  const cplx a = cadd(x[0], x[2]), b = csub(x[0], x[2]), c = cadd(x[1], x[3]), d = mul_mi(csub(x[1], x[3]));
  x[0] = cadd(a, c); x[1] = cadd(b, d); x[2] = csub(a, c); x[3] = csub(b, d);
}
template <> __device__ __forceinline__ void dft_small<5>(cplx (&x)[5]) { dft_prime<5>(x); }
template <> __device__ __forceinline__ void dft_small<7>(cplx (&x)[7]) { dft_prime<7>(x); }
template <> __device__ __forceinline__ void dft_small<8>(cplx (&x)[8]) {
  // two radix-4 on the even / odd inputs, then the radix-2 combination with w8^k
  cplx e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
  dft_small<4>(e); dft_small<4>(o);
  const double h = 0.70710678118654752440;
  const cplx o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * e^{-i pi/4}
  const cplx o2 = mul_mi(o[2]);
  const cplx o3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * e^{-3 i pi/4}
  x[0] = cadd(e[0], o[0]); x[4] = csub(e[0], o[0]);
  x[1] = cadd(e[1], o1);   x[5] = csub(e[1], o1);
  x[2] = cadd(e[2], o2);   x[6] = csub(e[2], o2);
  x[3] = cadd(e[3], o3);   x[7] = csub(e[3], o3);
}
template <> __device__ __forceinline__ void dft_small<11>(cplx (&x)[11]) { dft_prime<11>(x); }
template <> __device__ __forceinline__ void dft_small<13>(cplx (&x)[13]) { dft_prime<13>(x); }

// One Stockham stage of radix R on a tile held as [point][line]:
//   in[q + s (p + m j)] (j < R)  ->  out[q + s (R p + r)] = w_{R m}^{p r} DFT_R(in)_r,   p < m, q < s,  n = R m s.
// so / sl: output strides over the point and line index (sl != 1: transposed store for the next axis).
template <int R>
__device__ __forceinline__ void stockham_stage(const cplx* __restrict__ src, cplx* __restrict__ dst, int n, int m, int s,
                                               int pitch_in, long so, long sl, int lcnt, const cplx* __restrict__ W) {
  const int nbf = m * s;
  // two lines per thread (twiddle reuse) only when that still leaves every thread a butterfly
  const int lbw = (nbf * ((lcnt + 1) / 2) >= (int)blockDim.x) ? FFT_LB : 1;
  const int nlb = (lcnt + lbw - 1) / lbw;
  const int tot = nbf * nlb;
  for (int w = threadIdx.x; w < tot; w += blockDim.x) {
    const int lb = w % nlb, bf = w / nlb;
    const int p = bf / s, q = bf - p * s;
    cplx tw[R];
    {
      const int step = (int)(((long)p * s) % n);     // exponent of w_n per unit of r
      int e = 0;
#pragma unroll
      for (int r = 1; r < R; ++r) { e += step; if (e >= n) e -= n; tw[r] = W[e]; }
    }
    const cplx* xin = src + (long)(q + s * p) * pitch_in;
    const long xstep = (long)s * m * pitch_in;
#pragma unroll
    for (int l = 0; l < FFT_LB; ++l) {
      const int line = lb + l * nlb;
      if (l < lbw && line < lcnt) {
        cplx x[R];
#pragma unroll
        for (int j = 0; j < R; ++j) x[j] = xin[j * xstep + line];
        dft_small<R>(x);
        cplx* o = dst + (long)(q + s * R * p) * so + (long)line * sl;
        o[0] = x[0];
#pragma unroll
        for (int r = 1; r < R; ++r) o[(long)r * s * so] = cmul(x[r], tw[r]);
      }
    }
  }
}

// generic radix (primes 17 ... 61): direct R-point DFT per output
__device__ __noinline__ void stockham_stage_generic(int R, const cplx* __restrict__ src, cplx* __restrict__ dst, int n,
                                                       int m, int s, int pitch_in, long so, long sl, int lcnt,
                                                       const cplx* __restrict__ W) {
  const int nbf = m * s;
  const int tot = nbf * lcnt;
  const int nR = n / R;
  for (int w = threadIdx.x; w < tot; w += blockDim.x) {
    const int line = w % lcnt, bf = w / lcnt;
    const int p = bf / s, q = bf - p * s;
    cplx x[FFT_RMAX];
    const cplx* xin = src + (long)(q + s * p) * pitch_in + line;
    const long xstep = (long)s * m * pitch_in;
    for (int j = 0; j < R; ++j) x[j] = xin[j * xstep];
    const int step = (int)(((long)p * s) % n);
    int et = 0;
    for (int r = 0; r < R; ++r) {
      cplx acc = x[0];
      const int st = (int)(((long)nR * r) % n);
      int e = 0;
      for (int j = 1; j < R; ++j) { e += st; if (e >= n) e -= n; cfma(acc, x[j], W[e]); }
      dst[(long)(q + s * (R * p + r)) * so + (long)line * sl] = (r == 0) ? acc : cmul(acc, W[et]);
      et += step; if (et >= n) et -= n;
    }
  }
}

struct FftAxis {
  int n, nstages;
  int radix[FFT_MAXSTAGES];
  const cplx* tw;                 // n twiddles exp(-2 pi i j / n) (device)
};

// All stages of one axis on a tile.  a: input buffer (pitch pitch), b: the other buffer.  The LAST stage writes with
// strides (so_last, sl_last) into `last` (which may be a or b or a third layout).  Returns nothing; ends synchronised.
__device__ __forceinline__ void run_axis(const FftAxis& ax, cplx* a, cplx* b, int pitch, cplx* last, long so_last,
                                         long sl_last, int lcnt, const cplx* W) {
  const int n = ax.n;
  cplx* src = a;
  cplx* dst = b;
  int ncur = n, s = 1;
  for (int st = 0; st < ax.nstages; ++st) {
    const int R = ax.radix[st];
    const int m = ncur / R;
    const bool fin = (st == ax.nstages - 1);
    cplx* d = fin ? last : dst;
    const long so = fin ? so_last : pitch, sl = fin ? sl_last : 1;
    switch (R) {
      case 2: stockham_stage<2>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 3: stockham_stage<3>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 4: stockham_stage<4>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 5: stockham_stage<5>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 7: stockham_stage<7>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 8: stockham_stage<8>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 11: stockham_stage<11>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      case 13: stockham_stage<13>(src, d, n, m, s, pitch, so, sl, lcnt, W); break;
      default: stockham_stage_generic(R, src, d, n, m, s, pitch, so, sl, lcnt, W); break;
    }
    __syncthreads();
    cplx* t = src; src = dst; dst = t;
    ncur = m;
    s *= R;
  }
}

struct FftPlaneParams {
  cplx* data; long ldv;          // [nvec][ldv]
  int n1, n2, n3;
  FftAxis az, ay;                // z (length n3) and y (length n2)
  long nwork;                    // nvec * n1 planes
  const cplx* pre;               // [ng] or null
  const double* post;            // [ng] or null (only when n1 == 1)
};

// z + y of one x-plane.  Buffers: A as [z][y] (pitch pa >= n2, odd), B the same; the last z stage writes into the
// [y][kz] layout (pitch pb >= n3, odd) of the buffer the y stages start from.  Persistent CTAs of 256 threads.
__global__ void __launch_bounds__(512, 1) fft_plane_kernel(FftPlaneParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n2 = p.n2, n3 = p.n3;
  const int pa = n2 | 1, pb = n3 | 1;
  const int bufsz = max(n3 * pa, n2 * pb);
  const int nthr = blockDim.x;
  cplx* A = reinterpret_cast<cplx*>(smem_raw);
  cplx* B = A + bufsz;
  cplx* Wz = B + bufsz;
  cplx* Wy = Wz + n3;
  for (int i = threadIdx.x; i < n3; i += nthr) Wz[i] = p.az.tw[i];
  for (int i = threadIdx.x; i < n2; i += nthr) Wy[i] = p.ay.tw[i];
  const int plane_sz = n2 * n3;
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    const long poff = (long)plane * plane_sz;
    cplx* base = p.data + vec * p.ldv + poff;
    __syncthreads();
    // load [y][z] (z contiguous) -> A[z][y], phase fused
    for (int f = threadIdx.x; f < plane_sz; f += nthr) {
      const int y = f / n3, z = f - y * n3;
      cplx v = base[f];
      if (p.pre) v = cmul(v, p.pre[poff + f]);
      A[z * pa + y] = v;
    }
    __syncthreads();
    // z stages: lines = y (n2 of them).  Ping-pong parity: stage st reads (st even ? A : B); the last one writes the
    // [y][kz] layout (pitch pb) into the buffer it does not read.
    cplx* zlast = (p.az.nstages & 1) ? B : A;
    if (p.az.nstages == 0) {                         // n3 == 1: A[0][y] -> B[y][0]
      for (int f = threadIdx.x; f < plane_sz; f += nthr) B[f * pb] = A[f];
      __syncthreads();
      zlast = B;
    } else {
      run_axis(p.az, A, B, pa, zlast, /*so=*/1, /*sl=*/pb, n2, Wz);
    }
    cplx* ysrc = zlast;
    cplx* yoth = (zlast == A) ? B : A;
    cplx* ylast = ysrc;
    if (p.ay.nstages > 0) {
      ylast = (p.ay.nstages & 1) ? yoth : ysrc;
      run_axis(p.ay, ysrc, yoth, pb, ylast, pb, 1, n3, Wy);
    }
    // store [ky][kz]
    for (int f = threadIdx.x; f < plane_sz; f += nthr) {
      const int y = f / n3, z = f - y * n3;
      cplx v = ylast[y * pb + z];
      if (p.post) { const double wgt = p.post[poff + f]; v.x *= wgt; v.y *= wgt; }
      base[f] = v;
    }
  }
}

struct FftLinesParams {
  cplx* data; long vec_stride;
  FftAxis ax;
  long stride;        // element stride along the line
  long line_step;     // address step between consecutive lines of a run
  int lines_per_run;
  long run_stride;
  int nruns;
  int T;              // lines per tile
  long nwork;         // tiles in this launch (nvec * nruns * tiles_per_run)
  int contig;         // 1: lines are contiguous (stride == 1, line_step == n)
  const cplx* pre;
  const double* post;
};

__global__ void __launch_bounds__(FFT_THREADS, 2) fft_lines_kernel(FftLinesParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.ax.n, T = p.T;
  const int Tp = T | 1;
  cplx* X = reinterpret_cast<cplx*>(smem_raw);
  cplx* Y = X + (long)n * Tp;
  cplx* L = Y + (long)n * Tp;        // landing buffer of the next tile, same [point][line] layout
  cplx* W = L + (long)n * Tp;
  const int tiles_per_run = (p.lines_per_run + T - 1) / T;
  for (int i = threadIdx.x; i < n; i += FFT_THREADS) W[i] = p.ax.tw[i];
  auto locate = [&](long work, long& voff, int& lcnt, long& vecoff) {
    const int tile = (int)(work % tiles_per_run);
    long b = work / tiles_per_run;
    const int run = (int)(b % p.nruns);
    const long vec = b / p.nruns;
    const int l0 = tile * T;
    lcnt = min(T, p.lines_per_run - l0);
    voff = (long)run * p.run_stride + (long)l0 * p.line_step;
    vecoff = vec * p.vec_stride;
  };
  auto prefetch = [&](long work) {
    long voff, vecoff; int lcnt;
    locate(work, voff, lcnt, vecoff);
    const cplx* base = p.data + vecoff;
    const int tot = lcnt * n;
    if (p.contig) {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int l = f / n, i = f - l * n;
        cp_async16(L + i * Tp + l, base + voff + f, true);
      }
    } else {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int i = f / lcnt, l = f - i * lcnt;
        cp_async16(L + i * Tp + l, base + voff + (long)l * p.line_step + (long)i * p.stride, true);
      }
    }
    cp_async_commit();
  };
  if ((long)blockIdx.x < p.nwork) prefetch(blockIdx.x);
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    long voff, vecoff; int lcnt;
    locate(work, voff, lcnt, vecoff);
    cplx* base = p.data + vecoff;
    const int tot = lcnt * n;
    cp_async_wait<0>();
    __syncthreads();
    // landing buffer -> X (phase fused); same layout, conflict-free either way
    if (p.contig) {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int l = f / n, i = f - l * n;
        cplx v = L[i * Tp + l];
        if (p.pre) v = cmul(v, p.pre[voff + f]);
        X[i * Tp + l] = v;
      }
    } else {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int i = f / lcnt, l = f - i * lcnt;
        cplx v = L[i * Tp + l];
        if (p.pre) v = cmul(v, p.pre[voff + (long)l * p.line_step + (long)i * p.stride]);
        X[i * Tp + l] = v;
      }
    }
    __syncthreads();
    if (work + gridDim.x < p.nwork) prefetch(work + gridDim.x);
    cplx* last = (p.ax.nstages & 1) ? Y : X;
    if (p.ax.nstages > 0) run_axis(p.ax, X, Y, Tp, last, Tp, 1, lcnt, W);
    if (p.contig) {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int l = f / n, i = f - l * n;
        const long off = voff + f;
        cplx v = last[i * Tp + l];
        if (p.post) { const double wgt = p.post[off]; v.x *= wgt; v.y *= wgt; }
        base[off] = v;
      }
    } else {
      for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
        const int i = f / lcnt, l = f - i * lcnt;
        const long off = voff + (long)l * p.line_step + (long)i * p.stride;
        cplx v = last[i * Tp + l];
        if (p.post) { const double wgt = p.post[off]; v.x *= wgt; v.y *= wgt; }
        base[off] = v;
      }
    }
  }
}

// ---- host-side plan cache (twiddles live on the device, owned by the handle's plan table) ----
struct FftPlan {
  cplx* tw;
  int nstages;
  int radix[FFT_MAXSTAGES];
};

static std::map<std::pair<int, int>, FftPlan>& plan_table() {
  static std::map<std::pair<int, int>, FftPlan> t;
  return t;
}

// Radix schedule: 8s and 4s first (fewest stages for powers of two), then 2, then the odd primes in increasing order.
static bool choose_radices(int n, std::vector<int>& out) {
  out.clear();
  while (n % 8 == 0) { out.push_back(8); n /= 8; }
  while (n % 4 == 0) { out.push_back(4); n /= 4; }
  while (n % 2 == 0) { out.push_back(2); n /= 2; }
  for (int r = 3; n > 1; r += 2)
    while (n % r == 0) { out.push_back(r); n /= r; }
  for (int r : out)
    if (r > FFT_RMAX) return false;
  return (int)out.size() <= FFT_MAXSTAGES;
}

static int get_plan(Handle* h, int n, FftPlan** plan) {
  auto key = std::make_pair(h->device, n);
  auto it = plan_table().find(key);
  if (it == plan_table().end()) {
    FftPlan pl;
    std::vector<int> rad;
    if (!choose_radices(n, rad)) return ISDF_ESIZE;
    pl.nstages = (int)rad.size();
    for (int i = 0; i < FFT_MAXSTAGES; ++i) pl.radix[i] = (i < pl.nstages) ? rad[i] : 1;
    std::vector<cplx> tw(n);
    for (int j = 0; j < n; ++j) {
      const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
      tw[j] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    cudaError_t e = cudaMalloc(&pl.tw, sizeof(cplx) * n);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(pl.tw, tw.data(), sizeof(cplx) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    it = plan_table().insert(std::make_pair(key, pl)).first;
  }
  *plan = &it->second;
  return ISDF_OK;
}

static int make_axis(Handle* h, int n, FftAxis* ax) {
  FftPlan* pl;
  int rc = get_plan(h, n, &pl);
  if (rc != ISDF_OK) { snprintf(h->err, sizeof(h->err), "fft plan for n=%d failed (%d): a prime factor above %d", n, rc, FFT_RMAX); return rc; }
  ax->n = n; ax->nstages = pl->nstages; ax->tw = pl->tw;
  for (int i = 0; i < FFT_MAXSTAGES; ++i) ax->radix[i] = pl->radix[i];
  return ISDF_OK;
}

static int launch_lines(Handle* h, cplx* data, long nvec, long ldv, int n, long stride, long line_step,
                        int lines_per_run, long run_stride, int nruns, int contig, const cplx* pre, const double* post,
                        cudaStream_t st) {
  FftLinesParams p;
  int rc = make_axis(h, n, &p.ax);
  if (rc != ISDF_OK) return rc;
  p.data = data; p.vec_stride = ldv; p.stride = stride; p.line_step = line_step;
  p.lines_per_run = lines_per_run; p.run_stride = run_stride; p.nruns = nruns; p.contig = contig;
  p.pre = pre; p.post = post;
  int T = 32;
  auto bytes = [&](int t) { return ((size_t)3 * n * (t | 1) + n) * sizeof(cplx); };
  while (T > 1 && bytes(T) > (size_t)72 * 1024) T >>= 1;
  if (bytes(T) > (size_t)h->max_smem_optin) { snprintf(h->err, sizeof(h->err), "fft length %d too large", n); return ISDF_ESIZE; }
  if (T > lines_per_run) { T = 1; while (T * 2 <= lines_per_run) T *= 2; }
  p.T = T;
  const size_t smem = bytes(T);
  ISDF_CUDA(h, cudaFuncSetAttribute(fft_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long tiles_per_run = (lines_per_run + T - 1) / T;
  p.nwork = nvec * nruns * tiles_per_run;
  if (p.nwork <= 0) return ISDF_OK;
  int per_sm = (int)((size_t)h->max_smem_optin / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 2) per_sm = 2;                 // register budget of the kernel (__launch_bounds__(256, 2))
  long grid = (long)h->sm_count * per_sm;
  if (grid > p.nwork) grid = p.nwork;
  fft_lines_kernel<<<(unsigned)grid, FFT_THREADS, smem, st>>>(p);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

}  // namespace isdf

using namespace isdf;

// data: [nvec][ldv >= ng] c128 in place; forward (e^{-i}) unnormalised transform over mesh (C order, z fastest);
// out[v][G] = post[G] * sum_r data[v][r] * pre[r] * e^{-i G.r}.   group_vecs: vectors per L2-resident group.
extern "C" int isdf_fft3d_batched(void* hv, void* data, long nvec, long ldv, const int* mesh, const void* pre_dev,
                                  const double* post_dev, long group_vecs, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, data && mesh, "null pointer");
  const int n1 = mesh[0], n2 = mesh[1], n3 = mesh[2];
  ISDF_CHECK_ARG(h, n1 >= 1 && n2 >= 1 && n3 >= 1, "mesh");
  const long ng = (long)n1 * n2 * n3;
  ISDF_CHECK_ARG(h, ldv >= ng, "ldv < prod(mesh)");
  if (nvec <= 0) return ISDF_OK;
  if (group_vecs <= 0) {
    // measured on B200: one launch over the whole batch beats L2-sized groups (the kernels are latency-, not
    // HBM-bound: 0.95 vs 0.84 TB/s of minimum traffic at 64^3), so the default is no grouping
    group_vecs = nvec;
  }
  // z + y fused per plane when two padded plane buffers fit in shared memory
  const int pa = n2 | 1, pb = n3 | 1;
  const long bufsz = ((long)n3 * pa > (long)n2 * pb) ? (long)n3 * pa : (long)n2 * pb;
  const size_t plane_smem = (size_t)(2 * bufsz + n2 + n3) * sizeof(cplx);
  const bool fused = plane_smem <= (size_t)h->max_smem_optin - 1024;
  FftPlaneParams pp;
  if (fused) {
    int rc = make_axis(h, n3, &pp.az);
    if (rc) return rc;
    rc = make_axis(h, n2, &pp.ay);
    if (rc) return rc;
    ISDF_CUDA(h, cudaFuncSetAttribute(fft_plane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plane_smem));
  }
  for (long v0 = 0; v0 < nvec; v0 += group_vecs) {
    const long nv = (nvec - v0 < group_vecs) ? (nvec - v0) : group_vecs;
    cplx* d = (cplx*)data + v0 * ldv;
    int rc;
    if (fused) {
      pp.data = d; pp.ldv = ldv; pp.n1 = n1; pp.n2 = n2; pp.n3 = n3;
      pp.nwork = nv * n1; pp.pre = (const cplx*)pre_dev; pp.post = (n1 == 1) ? post_dev : nullptr;
      int per_sm = (int)((size_t)(h->max_smem_optin) / (plane_smem + 1024));
      if (per_sm < 1) per_sm = 1;
      if (per_sm > 2) per_sm = 2;               // register budget: 128 registers x 256 threads
      const int nthr = 256;                     // 512 threads for big planes measured slower (48^3: 4.1 vs 2.9 ms)
      long grid = (long)h->sm_count * per_sm;
      if (grid > pp.nwork) grid = pp.nwork;
      fft_plane_kernel<<<(unsigned)grid, nthr, plane_smem, st>>>(pp);
      ISDF_LAUNCH_CHECK(h);
    } else {
      // z: contiguous lines, n1*n2 of them per vector;  y: stride n3, runs over x
      rc = launch_lines(h, d, nv, ldv, n3, 1, n3, n1 * n2, 0, 1, 1, (const cplx*)pre_dev, (n1 == 1 && n2 == 1) ? post_dev : nullptr, st);
      if (rc) return rc;
      if (n2 > 1) {
        rc = launch_lines(h, d, nv, ldv, n2, n3, 1, n3, (long)n2 * n3, n1, 0, nullptr, (n1 == 1) ? post_dev : nullptr, st);
        if (rc) return rc;
      }
    }
    // x: stride n2*n3, one run of n2*n3 lines
    if (n1 > 1) {
      rc = launch_lines(h, d, nv, ldv, n1, (long)n2 * n3, 1, n2 * n3, 0, 1, 0, nullptr, post_dev, st);
      if (rc) return rc;
    }
  }
  return ISDF_OK;
}

extern "C" int isdf_fft_release_plans(void* hv) {
  Handle* h = (Handle*)hv;
  for (auto it = plan_table().begin(); it != plan_table().end();) {
    if (h == nullptr || it->first.first == h->device) {
      cudaFree(it->second.tw);
      it = plan_table().erase(it);
    } else {
      ++it;
    }
  }
  return ISDF_OK;
}
