// Batched 3-D complex128 FFT over the dense grid, fused with the e^{-iq.r} pre-phase and the
// sqrt(v(q+G) vol)/ng post-weight  (replaces pbctools.fft / get_coulG / ifft at
// /root/reference/fftisdf.py:113-119 -- the inverse FFT is removed analytically).
//
// One kernel does batched 1-D transforms along one mesh axis for a group of vectors:
//   * a tile of T lines (each of length n <= 128... any n) is staged in shared memory as [i][l]
//   * Stockham autosort stages (out-of-place ping-pong, natural-order output) with arbitrary
//     radices: each output point is a direct R-point DFT of its butterfly (generic for every R,
//     so awkward PySCF meshes -- 15, 31, 33, 37 -- need no special casing)
//   * global accesses are 16-byte and coalesced along the contiguous direction of each pass
// The host runs z, y, x passes back to back on L2-sized groups of vectors so that only the first
// read and the last write of every vector reach HBM.
#include <map>
#include <vector>
#include <math.h>
#include "common.cuh"

namespace isdf {

#ifndef ISDF_FFT_THREADS
#define ISDF_FFT_THREADS 128
#endif
#ifndef ISDF_FFT_OB
#define ISDF_FFT_OB 4
#endif
#ifndef ISDF_FFT_LB
#define ISDF_FFT_LB 4
#endif
#ifndef ISDF_FFT_T
#define ISDF_FFT_T 32
#endif
constexpr int FFT_THREADS = ISDF_FFT_THREADS;
constexpr int FFT_MAXSTAGES = 8;
constexpr int FFT_OB = ISDF_FFT_OB;   // outputs per thread (register blocking)
constexpr int FFT_LB = ISDF_FFT_LB;   // lines per thread

struct FftParams {
  cplx* data;        // in place
  long vec_stride;   // elements between vectors (= ng)
  int n;             // line length
  long stride;       // element stride along the line
  long line_step;    // address step between consecutive lines of a run
  int lines_per_run; // lines in a run
  long run_stride;   // address step between runs
  int nruns;         // runs per vector
  int T;             // lines per tile
  int contig;        // 1: lines are contiguous (stride == 1, line_step == n)
  int nstages;
  int radix[FFT_MAXSTAGES];
  const cplx* tw;    // n twiddles exp(-2 pi i j / n)
  const cplx* pre;   // [vec] complex pre-multiplier indexed by offset within the vector, or null
  const double* post;// [vec] real post-multiplier, or null
};

__global__ void __launch_bounds__(FFT_THREADS) fft_lines_kernel(FftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n = p.n, T = p.T;
  const int Tp = (T + FFT_LB) | 1;  // odd pitch (16-byte units), with room for the masked tail lines of a block
  cplx* X = reinterpret_cast<cplx*>(smem_raw);
  cplx* Y = X + (long)n * Tp;
  cplx* W = Y + (long)n * Tp;

  const int tiles_per_run = (p.lines_per_run + T - 1) / T;
  long bid = blockIdx.x;
  const int tile = (int)(bid % tiles_per_run); bid /= tiles_per_run;
  const int run = (int)(bid % p.nruns);
  const long vec = bid / p.nruns;
  const int l0 = tile * T;
  const int lcnt = min(T, p.lines_per_run - l0);
  const long voff = (long)run * p.run_stride + (long)l0 * p.line_step;  // offset within the vector
  cplx* base = p.data + vec * p.vec_stride;

  for (int i = threadIdx.x; i < n; i += FFT_THREADS) W[i] = p.tw[i];

  // ---- load (+ optional phase)
  if (p.contig) {
    const int tot = lcnt * n;
    for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
      const int l = f / n, i = f - l * n;
      const long off = voff + f;
      cplx v = base[off];
      if (p.pre) v = cmul(v, p.pre[off]);
      X[i * Tp + l] = v;
    }
  } else {
    const int tot = lcnt * n;
    for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
      const int i = f / lcnt, l = f - i * lcnt;
      const long off = voff + (long)l * p.line_step + (long)i * p.stride;
      cplx v = base[off];
      if (p.pre) v = cmul(v, p.pre[off]);
      X[i * Tp + l] = v;
    }
  }
  __syncthreads();

  // ---- Stockham stages.  ncur = current sub-transform length, s = stride (product of done radices)
  cplx* src = X;
  cplx* dst = Y;
  int ncur = n, s = 1;
  for (int st = 0; st < p.nstages; ++st) {
    const int R = p.radix[st];
    const int m = ncur / R;
    const int nR = n / R;
    // Register-blocked direct R-point DFT: a work item = (butterfly, chunk of FFT_OB outputs, FFT_LB lines).
    // Every x load is reused for FFT_OB outputs and every twiddle load for FFT_LB lines, which moves the
    // kernel from shared-memory-bandwidth bound (2 LDS per complex MAC) to the FP64 pipe (0.5 LDS/MAC).
    const int nbf = n / R;                       // butterflies per line
    const int nch = (R + FFT_OB - 1) / FFT_OB;   // output chunks per butterfly
    const int nlb = (lcnt + FFT_LB - 1) / FFT_LB;
    const int tot = nbf * nch * nlb;
    const long xstep = (long)s * m * Tp;
    for (int w = threadIdx.x; w < tot; w += FFT_THREADS) {
      const int lb = w % nlb;
      const int tmp = w / nlb;
      const int ch = tmp % nch;
      const int bf = tmp / nch;
      const int pp = bf / s, q = bf - pp * s;
      const int r10 = ch * FFT_OB;
      int step[FFT_OB], widx[FFT_OB];
#pragma unroll
      for (int j = 0; j < FFT_OB; ++j) {
        const int r1 = min(r10 + j, R - 1);
        step[j] = (int)(((long)nR * r1) % n);    // index step of w_R^{r1} per input r
        widx[j] = 0;
      }
      cplx acc[FFT_OB][FFT_LB];
#pragma unroll
      for (int j = 0; j < FFT_OB; ++j)
#pragma unroll
        for (int l = 0; l < FFT_LB; ++l) acc[j][l] = make_double2(0.0, 0.0);
      const cplx* xin = src + (long)(q + s * pp) * Tp + lb;
      for (int r = 0; r < R; ++r) {
        cplx xs[FFT_LB], ws[FFT_OB];
#pragma unroll
        for (int l = 0; l < FFT_LB; ++l) xs[l] = xin[r * xstep + l * nlb];   // lines lb + l*nlb (pad lines are zero-safe)
#pragma unroll
        for (int j = 0; j < FFT_OB; ++j) {
          ws[j] = W[widx[j]];
          widx[j] += step[j];
          if (widx[j] >= n) widx[j] -= n;
        }
#pragma unroll
        for (int j = 0; j < FFT_OB; ++j)
#pragma unroll
          for (int l = 0; l < FFT_LB; ++l) cfma(acc[j][l], xs[l], ws[j]);
      }
#pragma unroll
      for (int j = 0; j < FFT_OB; ++j) {
        const int r1 = r10 + j;
        if (r1 < R) {
          const int tidx = (int)(((long)pp * r1 % n) * s % n);
          const cplx tw = W[tidx];
          const int o = q + s * (R * pp + r1);
#pragma unroll
          for (int l = 0; l < FFT_LB; ++l) {
            const int line = lb + l * nlb;
            if (line < lcnt) dst[(long)o * Tp + line] = cmul(acc[j][l], tw);
          }
        }
      }
    }
    __syncthreads();
    cplx* tswap = src; src = dst; dst = tswap;
    ncur = m;
    s *= R;
  }

  // ---- store (+ optional weight)
  if (p.contig) {
    const int tot = lcnt * n;
    for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
      const int l = f / n, i = f - l * n;
      const long off = voff + f;
      cplx v = src[i * Tp + l];
      if (p.post) { const double wgt = p.post[off]; v.x *= wgt; v.y *= wgt; }
      base[off] = v;
    }
  } else {
    const int tot = lcnt * n;
    for (int f = threadIdx.x; f < tot; f += FFT_THREADS) {
      const int i = f / lcnt, l = f - i * lcnt;
      const long off = voff + (long)l * p.line_step + (long)i * p.stride;
      cplx v = src[i * Tp + l];
      if (p.post) { const double wgt = p.post[off]; v.x *= wgt; v.y *= wgt; }
      base[off] = v;
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

// factor n into radices minimising (sum of radices + per-stage overhead); primes stay whole.
static void choose_radices(int n, std::vector<int>& out) {
  out.clear();
  if (n == 1) return;
  // dynamic programme over divisors
  std::vector<int> best(n + 1, 1 << 30), choice(n + 1, 0);
  best[1] = 0;
  for (int v = 2; v <= n; ++v) {
    if (n % v) continue;
    for (int r = 2; r <= v; ++r) {
      if (v % r) continue;
      // cost r per point for this stage plus a per-stage overhead of 3 (sync + index math)
      const int c = best[v / r] + r + 3;
      if (best[v / r] < (1 << 30) && c < best[v]) { best[v] = c; choice[v] = r; }
    }
  }
  int v = n;
  while (v > 1) { out.push_back(choice[v]); v /= choice[v]; }
}

static int get_plan(Handle* h, int n, FftPlan** plan) {
  auto key = std::make_pair(h->device, n);
  auto it = plan_table().find(key);
  if (it == plan_table().end()) {
    FftPlan pl;
    std::vector<int> rad;
    choose_radices(n, rad);
    if ((int)rad.size() > FFT_MAXSTAGES) return ISDF_ESIZE;
    pl.nstages = (int)rad.size();
    for (int i = 0; i < pl.nstages; ++i) pl.radix[i] = rad[i];
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

static int launch_pass(Handle* h, cplx* data, long nvec, long ldv, int n, long stride, long line_step,
                       int lines_per_run, long run_stride, int nruns, int contig, const cplx* pre, const double* post,
                       cudaStream_t st) {
  FftPlan* pl;
  int rc = get_plan(h, n, &pl);
  if (rc != ISDF_OK) { snprintf(h->err, sizeof(h->err), "fft plan for n=%d failed (%d)", n, rc); return rc; }
  FftParams p;
  p.data = data; p.vec_stride = ldv; p.n = n; p.stride = stride; p.line_step = line_step;
  p.lines_per_run = lines_per_run; p.run_stride = run_stride; p.nruns = nruns; p.contig = contig;
  p.nstages = pl->nstages;
  for (int i = 0; i < FFT_MAXSTAGES; ++i) p.radix[i] = (i < pl->nstages) ? pl->radix[i] : 1;
  p.tw = pl->tw; p.pre = pre; p.post = post;
  int T = ISDF_FFT_T;
  auto bytes = [&](int t) { return ((size_t)2 * n * ((t + FFT_LB) | 1) + n) * sizeof(cplx); };
  while (T > 1 && bytes(T) > (size_t)96 * 1024) T >>= 1;
  if (bytes(T) > (size_t)h->max_smem_optin) { snprintf(h->err, sizeof(h->err), "fft length %d too large", n); return ISDF_ESIZE; }
  if (T > lines_per_run) { T = 1; while (T * 2 <= lines_per_run) T *= 2; }
  p.T = T;
  const size_t smem = bytes(T);
  static size_t configured = 0;
  if (smem > configured) {
    ISDF_CUDA(h, cudaFuncSetAttribute(fft_lines_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const long tiles_per_run = (lines_per_run + T - 1) / T;
  const long nblocks = nvec * nruns * tiles_per_run;
  if (nblocks <= 0) return ISDF_OK;
  if (nblocks >= (1L << 31)) { snprintf(h->err, sizeof(h->err), "fft grid too large"); return ISDF_ESIZE; }
  fft_lines_kernel<<<(unsigned)nblocks, FFT_THREADS, smem, st>>>(p);
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
    group_vecs = (long)(48.0 * 1024 * 1024 / ((double)ng * sizeof(cplx)));
    if (group_vecs < 1) group_vecs = 1;
  }
  for (long v0 = 0; v0 < nvec; v0 += group_vecs) {
    const long nv = (nvec - v0 < group_vecs) ? (nvec - v0) : group_vecs;
    cplx* d = (cplx*)data + v0 * ldv;
    int rc;
    // z: contiguous lines, n1*n2 of them per vector
    if (n3 > 1 || pre_dev) {
      rc = launch_pass(h, d, nv, ldv, n3, 1, n3, n1 * n2, 0, 1, 1, (const cplx*)pre_dev, (n1 == 1 && n2 == 1) ? post_dev : nullptr, st);
      if (rc) return rc;
    }
    // y: stride n3, runs over x
    if (n2 > 1) {
      rc = launch_pass(h, d, nv, ldv, n2, n3, 1, n3, (long)n2 * n3, n1, 0, nullptr, (n1 == 1) ? post_dev : nullptr, st);
      if (rc) return rc;
    }
    // x: stride n2*n3, one run of n2*n3 lines
    if (n1 > 1) {
      rc = launch_pass(h, d, nv, ldv, n1, (long)n2 * n3, 1, n2 * n3, 0, 1, 0, nullptr, post_dev, st);
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
