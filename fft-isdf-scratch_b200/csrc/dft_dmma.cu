// FP64 tensor-core DFT for awkward mesh lengths (large prime factors: 31, 37, 41, 43, 47, ...).
//
// PySCF's cutoff_to_mesh yields odd meshes whose prime factors are often large; a Stockham FFT then
// degenerates to an O(n^2) direct DFT on the FP64 FMA pipe.  Here every 1-D transform of length n <= 48 is
// a dense complex matrix product with the n x n DFT matrix, executed on the FP64 tensor pipe (DMMA.8x8x4)
// with the same conflict-free LDS.128 fragment scheme as the GEMM engine (gemm_c128.cuh):
//   * dft_zy_kernel: one CTA per (vector, x-plane); the n2 x n3 plane is staged once in shared memory,
//     transformed along z then along y (the second product reads the first one's output as a K-contiguous
//     operand, so no transpose is needed), and written back coalesced.  The e^{-iq.r} phase is fused into
//     the load.
//   * dft_x_kernel: lines along x for 64 consecutive (y,z) positions; the sqrt(v(q+G) vol)/ng weight is
//     fused into the store.
// Replaces pbctools.fft at /root/reference/fftisdf.py:113 (+ :99, :114-115) for such meshes.
#include <math.h>
#include <map>
#include <vector>
#include "common.cuh"

namespace isdf {

constexpr int DFT_MAXN = 48;
constexpr int DFT_THREADS = 256;
#ifndef ISDF_DFT_3M
#define ISDF_DFT_3M 1     // complex products in the 3M (Karatsuba) form: 3 DMMA per complex MAC (see gemm_c128.cuh)
#endif
constexpr bool DFT_3M = ISDF_DFT_3M != 0;

// One k-step of a complex 8x8 tile product.  4M: (re, im) accumulate directly.  3M: re = P1 = sum ar*br,
// im = P2 = sum ai*bi, p3 = sum (ar+ai)(br+bi); cmac_finish() turns them into the complex result.
__device__ __forceinline__ void cmac(double (&re)[2], double (&im)[2], double (&p3)[2], cplx a, cplx b) {
  if (DFT_3M) {
    dmma884(re[0], re[1], a.x, b.x);
    dmma884(im[0], im[1], a.y, b.y);
    dmma884(p3[0], p3[1], a.x + a.y, b.x + b.y);
  } else {
    dmma884(re[0], re[1], a.x, b.x);
    dmma884(re[0], re[1], a.y, -b.y);
    dmma884(im[0], im[1], a.x, b.y);
    dmma884(im[0], im[1], a.y, b.x);
  }
}
__device__ __forceinline__ cplx cmac_finish(double re, double im, double p3) {
  return DFT_3M ? make_double2(re - im, (p3 - re) - im) : make_double2(re, im);
}

struct DftParams {
  cplx* data; long ldv;        // [nvec][ldv]
  int n1, n2, n3;
  const cplx* w1; const cplx* w2; const cplx* w3;      // padded DFT matrices [np][np], W[i][k] = e^{-2 pi i ik/n}
  long nwork;                  // planes (zy kernel) or line tiles (x kernel) in this launch
  // fused exchange over NVLink peer memory (world > 1): the zy kernel GATHERS its planes from the ranks'
  // grid-column shards peer[r][(row0 + v) * ncol + (g - r*ncol)], the x kernel SCATTERS its output there.
  cplx* peer[8]; int world; long ncol; long row0;
  const cplx* pre;             // [ng] or null
  const double* post;          // [ng] or null
};

__device__ __forceinline__ int pad8(int n) { return (n + 7) & ~7; }

// copy the padded [np][np] DFT matrix from the plan into shared memory with row pitch ld
__device__ __forceinline__ void fill_dft_matrix(cplx* W, const cplx* wg, int np, int ld) {
  for (int w = threadIdx.x; w < np * np; w += DFT_THREADS) {
    const int i = w / np, k = w - i * np;
    W[i * ld + k] = wg[w];
  }
}

// acc (NT adjacent 8x8 complex tiles along n, as re[2], im[2] (, p3[2]) per lane) += A(8 x K) * B(K x 8 NT), a*b.
// NT = 2 shares the A fragment and doubles the independent accumulator chains per warp.
template <bool A_KSLOW, int NT>
__device__ __forceinline__ void tile_mma(double (&re)[NT][2], double (&im)[NT][2], double (&p3)[NT][2], const cplx* A,
                                         int lda, int m0, const cplx* B, int ldb, int n0, int K, int g, int t) {
  for (int k0 = 0; k0 < K; k0 += 4) {
    const cplx a = A_KSLOW ? A[(k0 + t) * lda + m0 + g] : A[(m0 + g) * lda + k0 + t];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const cplx b = B[(k0 + t) * ldb + n0 + 8 * j + g];
      cmac(re[j], im[j], p3[j], a, b);
    }
  }
}

// The 8x8 tiles of an (mt_n x nt_n)-tile product are dealt to the warps as jobs of two adjacent n-tiles (all the
// pairs first, then the odd last column as singles): job -> (mt, nt0, cnt).
__device__ __forceinline__ void dft_job(int job, int mt_n, int nt_n, int& mt, int& nt0, int& cnt) {
  const int npair = nt_n >> 1;
  if (job < mt_n * npair) { mt = job / npair; nt0 = 2 * (job - mt * npair); cnt = 2; }
  else { mt = job - mt_n * npair; nt0 = nt_n - 1; cnt = 1; }
}

// PIPE (n2 == n3): the DFT matrix is shared by both passes and the next plane is prefetched with cp.async into
// a landing buffer while the current plane is transformed.  With GATHER the prefetch pulls the plane straight
// from the owning ranks' shards over NVLink, so the ~2-3k-cycle peer latency is hidden behind the tensor pipe.
template <bool GATHER, bool PIPE>
__global__ void __launch_bounds__(DFT_THREADS, 2) dft_zy_kernel(DftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n2 = p.n2, n3 = p.n3;
  const int n2p = pad8(n2), n3p = pad8(n3);
  const int LDX = n2p + 2, LDY = n2p + 4, LDW3 = n3p + 2, LDW2 = n2p + 2;
  cplx* Xs = reinterpret_cast<cplx*>(smem_raw);   // [n3p][LDX]   Xs[z][y]
  cplx* Ys = Xs + n3p * LDX;                      // [n3p][LDY]   Ys[kz][y]
  cplx* W3 = Ys + n3p * LDY;                      // [n3p][LDW3]  W3[z][kz]
  cplx* W2 = PIPE ? W3 : W3 + n3p * LDW3;         // [n2p][LDW2]  W2[y][ky]  (same matrix when n2 == n3)
  cplx* Ls = W3 + n3p * LDW3;                     // PIPE only: [n2*n3] raw landing buffer of the next plane
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int plane_sz = n2 * n3;
  fill_dft_matrix(W3, p.w3, n3p, LDW3);
  if (!PIPE) fill_dft_matrix(W2, p.w2, n2p, LDW2);
  auto prefetch = [&](long work) {
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    if (GATHER) {
      const long g0 = (long)plane * plane_sz;
      const int owner0 = (int)(g0 / p.ncol);
      const long rem0 = g0 - owner0 * p.ncol;
      for (int w = tid; w < plane_sz; w += DFT_THREADS) {
        int owner = owner0;
        long rem = rem0 + w;
        while (rem >= p.ncol) { rem -= p.ncol; ++owner; }
        cp_async16(Ls + w, p.peer[owner] + (p.row0 + vec) * p.ncol + rem, true);
      }
    } else {
      const cplx* src = p.data + vec * p.ldv + (long)plane * plane_sz;
      for (int w = tid; w < plane_sz; w += DFT_THREADS) cp_async16(Ls + w, src + w, true);
    }
    cp_async_commit();
  };
  if (PIPE && (long)blockIdx.x < p.nwork) prefetch(blockIdx.x);
  // persistent over planes: the DFT matrices stay in shared memory
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    const long poff = (long)plane * plane_sz;
    cplx* base = p.data + vec * p.ldv + poff;
    if (PIPE) {
      cp_async_wait<0>();
      __syncthreads();          // landing buffer complete; every warp is done with the previous plane
      for (int w = tid; w < n3p * n2p; w += DFT_THREADS) {
        const int y = w / n3p, z = w - y * n3p;
        cplx v = make_double2(0.0, 0.0);
        if (y < n2 && z < n3) {
          v = Ls[y * n3 + z];
          if (p.pre) v = cmul(v, p.pre[poff + y * n3 + z]);
        }
        Xs[z * LDX + y] = v;
      }
      __syncthreads();
      if (work + gridDim.x < p.nwork) prefetch(work + gridDim.x);   // flies during both passes
    } else {
      // load plane transposed: Xs[z][y] = base[y*n3 + z] * pre ; zero padding
      for (int w = tid; w < n3p * n2p; w += DFT_THREADS) {
        const int y = w / n3p, z = w - y * n3p;   // z fastest: coalesced global reads
        cplx v = make_double2(0.0, 0.0);
        if (y < n2 && z < n3) {
          if (GATHER) {
            const long gidx = poff + y * n3 + z;
            const int owner = (int)(gidx / p.ncol);
            v = p.peer[owner][(p.row0 + vec) * p.ncol + (gidx - owner * p.ncol)];
          } else {
            v = base[y * n3 + z];
          }
          if (p.pre) v = cmul(v, p.pre[poff + y * n3 + z]);
        }
        Xs[z * LDX + y] = v;
      }
      __syncthreads();
    }
    // ---- pass 1: T[y][kz] = sum_z Xs[z][y] W3[z][kz]   -> Ys[kz][y]
    {
      const int mt_n = n2p >> 3, nt_n = n3p >> 3;
      const int njob = mt_n * ((nt_n >> 1) + (nt_n & 1));
      for (int w = warp; w < njob; w += DFT_THREADS / 32) {
        int mt, nt0, cnt;
        dft_job(w, mt_n, nt_n, mt, nt0, cnt);
        double re[2][2] = {}, im[2][2] = {}, p3[2][2] = {};
        if (cnt == 2) tile_mma<true, 2>(re, im, p3, Xs, LDX, mt * 8, W3, LDW3, nt0 * 8, n3p, g, t);
        else tile_mma<true, 1>(reinterpret_cast<double(&)[1][2]>(re), reinterpret_cast<double(&)[1][2]>(im),
                               reinterpret_cast<double(&)[1][2]>(p3), Xs, LDX, mt * 8, W3, LDW3, nt0 * 8, n3p, g, t);
        for (int j = 0; j < cnt; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e)
            Ys[((nt0 + j) * 8 + 2 * t + e) * LDY + mt * 8 + g] = cmac_finish(re[j][e], im[j][e], p3[j][e]);
      }
    }
    __syncthreads();
    // ---- pass 2: out[kz][ky] = sum_y Ys[kz][y] W2[y][ky]   -> global [ky][kz]
    {
      const int mt_n = n3p >> 3, nt_n = n2p >> 3;
      const int njob = mt_n * ((nt_n >> 1) + (nt_n & 1));
      for (int w = warp; w < njob; w += DFT_THREADS / 32) {
        int mt, nt0, cnt;
        dft_job(w, mt_n, nt_n, mt, nt0, cnt);
        double re[2][2] = {}, im[2][2] = {}, p3[2][2] = {};
        if (cnt == 2) tile_mma<false, 2>(re, im, p3, Ys, LDY, mt * 8, W2, LDW2, nt0 * 8, n2p, g, t);
        else tile_mma<false, 1>(reinterpret_cast<double(&)[1][2]>(re), reinterpret_cast<double(&)[1][2]>(im),
                                reinterpret_cast<double(&)[1][2]>(p3), Ys, LDY, mt * 8, W2, LDW2, nt0 * 8, n2p, g, t);
        const int kz = mt * 8 + g;
        for (int j = 0; j < cnt; ++j)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int ky = (nt0 + j) * 8 + 2 * t + e;
            if (kz < n3 && ky < n2) base[ky * n3 + kz] = cmac_finish(re[j][e], im[j][e], p3[j][e]);
          }
      }
    }
    // non-PIPE: Xs is rewritten by the next plane's load only after every warp finished pass 1 (barrier
    // above); Ys is rewritten in the next pass 1, which comes after the next load's barrier.
  }
}

constexpr int DFTX_LINES = 64;

// The line tiles are double-buffered: the next tile is fetched with cp.async (zero-filled padding) straight
// into the fragment layout while the current one is on the tensor pipe -- one barrier per tile.
template <bool SCATTER>
__global__ void __launch_bounds__(DFT_THREADS, 2) dft_x_kernel(DftParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n1 = p.n1;
  const long n23 = (long)p.n2 * p.n3;
  const int n1p = pad8(n1);
  const int LDX = DFTX_LINES + 2, LDW = n1p + 2;
  cplx* Xbuf = reinterpret_cast<cplx*>(smem_raw); // 2 x [n1p][LDX]  Xs[x][l]
  cplx* W1 = Xbuf + 2 * n1p * LDX;                // [n1p][LDW]  W1[x][kx]
  const int ntile = (int)((n23 + DFTX_LINES - 1) / DFTX_LINES);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  fill_dft_matrix(W1, p.w1, n1p, LDW);
  const int m0 = warp * 8;          // warp w owns lines 8w .. 8w+7 and all kx tiles (shared A fragment)
  const int nt_n = n1p >> 3;
  auto prefetch = [&](long work, cplx* dst) {
    const int tile = (int)(work % ntile);
    const long vec = work / ntile;
    const long l0 = (long)tile * DFTX_LINES;
    const int lcnt = (int)((n23 - l0 < DFTX_LINES) ? (n23 - l0) : DFTX_LINES);
    const cplx* src = p.data + vec * p.ldv + l0;
    for (int w = tid; w < n1p * DFTX_LINES; w += DFT_THREADS) {
      const int x = w / DFTX_LINES, l = w - x * DFTX_LINES;
      const bool v = x < n1 && l < lcnt;
      cp_async16(dst + x * LDX + l, v ? src + (long)x * n23 + l : src, v);
    }
    cp_async_commit();
  };
  if ((long)blockIdx.x < p.nwork) prefetch(blockIdx.x, Xbuf);
  int buf = 0;
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x, buf ^= 1) {
    const int tile = (int)(work % ntile);
    const long vec = work / ntile;
    const long l0 = (long)tile * DFTX_LINES;
    const int lcnt = (int)((n23 - l0 < DFTX_LINES) ? (n23 - l0) : DFTX_LINES);
    cplx* base = p.data + vec * p.ldv + l0;
    const cplx* Xs = Xbuf + buf * n1p * LDX;
    cp_async_wait<0>();
    __syncthreads();   // this tile has landed (also orders the W1 fill); every warp is done with the other buffer
    if (work + gridDim.x < p.nwork) prefetch(work + gridDim.x, Xbuf + (buf ^ 1) * n1p * LDX);
    double re[DFT_MAXN / 8][2], im[DFT_MAXN / 8][2], p3[DFT_3M ? DFT_MAXN / 8 : 1][2];
#pragma unroll
    for (int nt = 0; nt < DFT_MAXN / 8; ++nt) {
      re[nt][0] = re[nt][1] = im[nt][0] = im[nt][1] = 0.0;
      if (DFT_3M) p3[nt][0] = p3[nt][1] = 0.0;
    }
    for (int k0 = 0; k0 < n1p; k0 += 4) {
      const cplx a = Xs[(k0 + t) * LDX + m0 + g];
#pragma unroll
      for (int nt = 0; nt < DFT_MAXN / 8; ++nt) {
        if (nt < nt_n) {
          const cplx b = W1[(k0 + t) * LDW + nt * 8 + g];
          cmac(re[nt], im[nt], p3[DFT_3M ? nt : 0], a, b);
        }
      }
    }
    const int l = m0 + g;
#pragma unroll
    for (int nt = 0; nt < DFT_MAXN / 8; ++nt) {
      if (nt < nt_n) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int kx = nt * 8 + 2 * t + e;
          if (kx < n1 && l < lcnt) {
            cplx v = cmac_finish(re[nt][e], im[nt][e], p3[DFT_3M ? nt : 0][e]);
            const long off = (long)kx * n23 + l;
            if (p.post) { const double wgt = p.post[l0 + off]; v.x *= wgt; v.y *= wgt; }
            if (SCATTER) {
              const long gidx = l0 + off;
              const int owner = (int)(gidx / p.ncol);
              p.peer[owner][(p.row0 + vec) * p.ncol + (gidx - owner * p.ncol)] = v;
            } else {
              base[off] = v;
            }
          }
        }
      }
    }
  }
}

// padded DFT matrices W[i][k] = exp(-2 pi i ik/n) ([np][np], zero padding), cached per (device, n)
static std::map<std::pair<int, int>, cplx*>& dft_w_table() {
  static std::map<std::pair<int, int>, cplx*> t;
  return t;
}

static int get_w(Handle* h, int n, const cplx** out) {
  auto key = std::make_pair(h->device, n);
  auto it = dft_w_table().find(key);
  if (it == dft_w_table().end()) {
    const int np = (n + 7) & ~7;
    std::vector<cplx> w((size_t)np * np, make_double2(0.0, 0.0));
    for (int i = 0; i < n; ++i)
      for (int k = 0; k < n; ++k) {
        const long j = ((long)i * k) % n;
        const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
        w[(size_t)i * np + k] = make_double2((double)cosl(ang), (double)sinl(ang));
      }
    cplx* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(cplx) * w.size());
    if (e != cudaSuccess) return (int)e;
    e = cudaMemcpy(d, w.data(), sizeof(cplx) * w.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return (int)e;
    it = dft_w_table().insert(std::make_pair(key, d)).first;
  }
  *out = it->second;
  return ISDF_OK;
}

}  // namespace isdf

using namespace isdf;

extern "C" void dft_release_plans_internal(int device) {
  for (auto it = dft_w_table().begin(); it != dft_w_table().end();) {
    if (it->first.first == device) { cudaFree(it->second); it = dft_w_table().erase(it); } else { ++it; }
  }
}

static int dft3d_run(Handle* h, cplx* local, long nvec, long ldv, const int* mesh, const void* pre_dev,
                     const double* post_dev, cplx* const* peer, int world, long ncol, long row0, cudaStream_t st) {
  const int n1 = mesh[0], n2 = mesh[1], n3 = mesh[2];
  if (n1 < 2 || n2 < 2 || n3 < 2 || n1 > DFT_MAXN || n2 > DFT_MAXN || n3 > DFT_MAXN) return ISDF_ESIZE;
  const long ng = (long)n1 * n2 * n3;
  ISDF_CHECK_ARG(h, ldv >= ng, "ldv < prod(mesh)");
  if (nvec <= 0) return ISDF_OK;
  DftParams p;
  p.ldv = ldv; p.n1 = n1; p.n2 = n2; p.n3 = n3;
  p.pre = (const cplx*)pre_dev; p.post = post_dev;
  p.world = world; p.ncol = ncol; p.row0 = row0;
  for (int r = 0; r < 8; ++r) p.peer[r] = (peer && r < world) ? peer[r] : nullptr;
  int rc;
  if ((rc = get_w(h, n1, &p.w1)) || (rc = get_w(h, n2, &p.w2)) || (rc = get_w(h, n3, &p.w3))) {
    snprintf(h->err, sizeof(h->err), "dft matrix allocation failed (%d)", rc);
    return rc;
  }
  const int n1p = (n1 + 7) & ~7, n2p = (n2 + 7) & ~7, n3p = (n3 + 7) & ~7;
  const size_t sm_zy = (size_t)(n3p * (n2p + 2) + n3p * (n2p + 4) + n3p * (n3p + 2) + n2p * (n2p + 2)) * sizeof(cplx);
  const size_t sm_x = (size_t)(2 * n1p * (DFTX_LINES + 2) + n1p * (n1p + 2)) * sizeof(cplx);
  ISDF_CHECK_ARG(h, sm_zy <= (size_t)h->max_smem_optin && sm_x <= (size_t)h->max_smem_optin, "mesh too large");
  const bool p2p = peer != nullptr;
  // in-place hazard of the pipelined variant: the prefetch of a later plane must not race with pass-2 stores of
  // another CTA -- planes are disjoint, and a CTA only prefetches planes it will itself transform, so it is safe.
  const size_t sm_pipe = (size_t)(n3p * (n2p + 2) + n3p * (n2p + 4) + n3p * (n3p + 2) + n2 * n3) * sizeof(cplx);
  const bool pipe = n2 == n3 && 2 * sm_pipe <= (size_t)h->max_smem_optin;
  const size_t sm_zy_used = pipe ? sm_pipe : sm_zy;
  if (p2p) {
    if (pipe) ISDF_CUDA(h, (cudaFuncSetAttribute(dft_zy_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pipe)));
    else ISDF_CUDA(h, (cudaFuncSetAttribute(dft_zy_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_zy)));
    ISDF_CUDA(h, cudaFuncSetAttribute(dft_x_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_x));
  } else {
    if (pipe) ISDF_CUDA(h, (cudaFuncSetAttribute(dft_zy_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_pipe)));
    else ISDF_CUDA(h, (cudaFuncSetAttribute(dft_zy_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_zy)));
    ISDF_CUDA(h, cudaFuncSetAttribute(dft_x_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_x));
  }
  // Persistent CTAs: exactly the resident set (2 per SM), so that the DFT-matrix fill and the pipeline prologue
  // are paid once per CTA; the grid is trimmed so that every CTA runs the same number of rounds.  The whole
  // batch goes into one pair of launches (ncu: 64 MB groups did not keep the intermediate in L2 anyway, and
  // DRAM is at ~10 % of its peak in both kernels).
  const long group = 1L << 30;
  const long ntile = ((long)n2 * n3 + DFTX_LINES - 1) / DFTX_LINES;
  const long resident = (long)h->sm_count * 2;
  auto grid_for = [&](long nwork) {
    const long rounds = (nwork + resident - 1) / resident;
    return (unsigned)((nwork + rounds - 1) / rounds);
  };
  for (long v0 = 0; v0 < nvec; v0 += group) {
    const long nv = (nvec - v0 < group) ? (nvec - v0) : group;
    p.data = local + v0 * ldv;
    p.row0 = row0 + v0;
    p.nwork = nv * n1;
    const unsigned g1 = grid_for(p.nwork);
    if (p2p && pipe) dft_zy_kernel<true, true><<<g1, DFT_THREADS, sm_zy_used, st>>>(p);
    else if (p2p) dft_zy_kernel<true, false><<<g1, DFT_THREADS, sm_zy, st>>>(p);
    else if (pipe) dft_zy_kernel<false, true><<<g1, DFT_THREADS, sm_zy_used, st>>>(p);
    else dft_zy_kernel<false, false><<<g1, DFT_THREADS, sm_zy, st>>>(p);
    ISDF_LAUNCH_CHECK(h);
    p.nwork = nv * ntile;
    const unsigned g2 = grid_for(p.nwork);
    if (p2p) dft_x_kernel<true><<<g2, DFT_THREADS, sm_x, st>>>(p);
    else dft_x_kernel<false><<<g2, DFT_THREADS, sm_x, st>>>(p);
    ISDF_LAUNCH_CHECK(h);
  }
  return ISDF_OK;
}

// Same contract as isdf_fft3d_batched, every mesh axis in [2, 48].  Returns -2 (no launch) otherwise.
extern "C" int isdf_dft3d_dmma(void* hv, void* data, long nvec, long ldv, const int* mesh, const void* pre_dev,
                               const double* post_dev, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, data && mesh, "null pointer");
  return dft3d_run(h, (cplx*)data, nvec, ldv, mesh, pre_dev, post_dev, nullptr, 1, 0, 0, (cudaStream_t)stream);
}

// Multi-GPU variant with the all-to-all exchanges fused into the transform over NVLink peer memory.
// peers: HOST array of `world` (<= 8) device pointers, peers[r] = rank r's grid-column shard
// [rows][ncol] (peer-mapped, e.g. torch symmetric memory); this rank transforms the `nvec` vectors stored in
// rows row0 .. row0+nvec-1 of every shard: the z/y kernel gathers each plane from the owning ranks, the
// result stays in `work` ([nvec][ldv >= ng], local), the x kernel scatters its output (times `post`) back
// into the shards.  Callers must barrier across ranks before (shards complete) and after (scatter visible).
extern "C" int isdf_dft3d_dmma_p2p(void* hv, void* const* peers, int world, long ncol, long row0, void* work,
                                   long nvec, long ldv, const int* mesh, const void* pre_dev,
                                   const double* post_dev, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, peers && work && mesh, "null pointer");
  ISDF_CHECK_ARG(h, world >= 1 && world <= 8 && ncol >= 1 && row0 >= 0, "world/ncol/row0");
  ISDF_CHECK_ARG(h, (long)world * ncol >= (long)mesh[0] * mesh[1] * mesh[2], "shards do not cover the grid");
  return dft3d_run(h, (cplx*)work, nvec, ldv, mesh, pre_dev, post_dev, (cplx* const*)peers, world, ncol, row0,
                   (cudaStream_t)stream);
}
