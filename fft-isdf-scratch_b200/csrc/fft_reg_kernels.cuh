// Kernels and launchers of the register-resident FFT (see fft_reg.cu); included by the fft_reg_part*.cu translation
// units, each of which instantiates a slice of the size table (parallel compilation).
#pragma once
#include "fft_reg_core.cuh"

namespace isdf {
namespace fftreg {

// Multi-GPU (P2P = true): the grid columns of every vector are sharded over the ranks, peer[r] = rank r's shard
// [rows][ncol] in NVLink peer-mapped memory (rank r owns grid points [r ncol, (r+1) ncol)).  The plane kernel GATHERS
// its planes from the owning ranks (P2P loads) into the local work buffer `data`, the x pass SCATTERS the weighted
// result back into the shards (P2P stores): both all-to-all exchanges are fused into the transform.
struct PeerArgs {
  cplx* peer[8];
  long ncol, row0;
};

struct PlaneArgs {
  cplx* data; long ldv;
  int n1;                 // number of x planes per vector
  long nwork;             // nvec * n1
  const cplx* tw;         // [N] exp(-2 pi i m / N) (device)
  const cplx* pre;        // [ng] or null
  const double* post;     // [ng] or null (applied here only when n1 == 1)
  PeerArgs pr;
};

struct LinesArgs {
  cplx* data; long ldv;
  long stride;            // n2 * n3: element stride along x == number of lines
  int tiles;              // tiles per vector
  long nwork;             // nvec * tiles
  const cplx* tw;
  const double* post;
  PeerArgs pr;
};

// plane element idx of (vec, plane) -> address in the owning rank's shard
struct Gather {
  cplx* const* peer; long ncol, rowoff, rem0; int owner0;
  __device__ __forceinline__ Gather(cplx* const* peer_s, const PeerArgs& pr, long vec, long poff) : peer(peer_s), ncol(pr.ncol) {
    rowoff = (pr.row0 + vec) * pr.ncol;
    owner0 = (int)(poff / pr.ncol);
    rem0 = poff - owner0 * pr.ncol;
  }
  __device__ __forceinline__ cplx* at(long idx) const {
    long rem = rem0 + idx;
    int owner = owner0;
    while (rem >= ncol) { rem -= ncol; ++owner; }
    return peer[owner] + rowoff + rem;
  }
};

template <class AX, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_plane_kernel(PlaneArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  constexpr int N = AX::N;
  cplx* P = reinterpret_cast<cplx*>(smem_raw);
  cplx* TW = P + AX::SLOTS;
  for (int i = threadIdx.x; i < N; i += THREADS) TW[i] = p.tw[i];
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pr.peer[threadIdx.x];
  __syncthreads();
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    const long poff = (long)plane * N * N;
    cplx* base = p.data + vec * p.ldv + poff;
    const cplx* pre = p.pre ? p.pre + poff : nullptr;
    const double* post = p.post ? p.post + poff : nullptr;
    if (P2P) {
      const Gather ga(peer_s, p.pr, vec, poff);
      plane_z1<AX, THREADS>(threadIdx.x, [&](int idx) { cplx v = *ga.at(idx); if (pre) v = c_mul(v, pre[idx]); return v; }, P, TW);
    } else {
      plane_z1<AX, THREADS>(threadIdx.x, [&](int idx) { cplx v = base[idx]; if (pre) v = c_mul(v, pre[idx]); return v; }, P, TW);
    }
    __syncthreads();
    plane_z2<AX, THREADS>(threadIdx.x, P);
    __syncthreads();
    plane_y1<AX, THREADS>(threadIdx.x, P, TW);
    __syncthreads();
    plane_y2<AX, THREADS>(threadIdx.x, P, [&](int o, cplx v) {
      if (post) { const double w = post[o]; v.x *= w; v.y *= w; }
      base[o] = v;
    });
    __syncthreads();
  }
}

// x-pass store: grid index g of (vec) -> local vector or the owning rank's shard
template <bool P2P>
struct XStore {
  cplx* base; const double* post; long stride; cplx* const* peer; long ncol, rowoff; long l0;
  __device__ __forceinline__ void operator()(int k, int l, cplx v) const {
    const long g = (long)k * stride + l0 + l;
    if (post) { const double w = post[g]; v.x *= w; v.y *= w; }
    if (P2P) {
      const unsigned owner = (unsigned)g / (unsigned)ncol;
      peer[owner][rowoff + (g - (long)owner * ncol)] = v;
    } else {
      base[g] = v;
    }
  }
};

template <class AX, int T, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_lines_kernel(LinesArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  constexpr int N = AX::N;
  cplx* S = reinterpret_cast<cplx*>(smem_raw);
  cplx* TW = S + N * T;
  for (int i = threadIdx.x; i < N; i += THREADS) TW[i] = p.tw[i];
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pr.peer[threadIdx.x];
  __syncthreads();
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int tile = (int)(work % p.tiles);
    const long vec = work / p.tiles;
    const long l0 = (long)tile * T;
    const int lcnt = (int)((p.stride - l0 < T) ? (p.stride - l0) : T);
    cplx* base = p.data + vec * p.ldv;
    const cplx* src = base + l0;
    const long stride = p.stride;
    lines_s1<AX, T, THREADS>(threadIdx.x, [&](int x, int l) { return src[(long)x * stride + l]; }, lcnt, S, TW);
    __syncthreads();
    const XStore<P2P> st{base, p.post, stride, peer_s, p.pr.ncol, (p.pr.row0 + vec) * p.pr.ncol, l0};
    lines_s2<AX, T, THREADS>(threadIdx.x, S, lcnt, st);
    __syncthreads();
  }
}

// ---- prime lengths ---------------------------------------------------------------------------------------------
template <class AX, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_plane_direct_kernel(PlaneArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  constexpr int N = AX::N, G = AX::G, PITCH = AX::PITCH;
  constexpr int LP = (N + 31) / 32 * 32;            // lines padded to whole warps: the group index is warp-uniform
  cplx* A = reinterpret_cast<cplx*>(smem_raw);      // A[z][y]
  cplx* B = A + AX::SLOTS;                          // B[kz][y]
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pr.peer[threadIdx.x];
  __syncthreads();
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    const long poff = (long)plane * N * N;
    cplx* base = p.data + vec * p.ldv + poff;
    const cplx* pre = p.pre ? p.pre + poff : nullptr;
    const double* post = p.post ? p.post + poff : nullptr;
    const Gather ga(peer_s, p.pr, vec, P2P ? poff : 0);
#pragma unroll 4
    for (int f = threadIdx.x; f < N * N; f += THREADS) {
      const int y = f / N, z = f % N;
      cplx v = P2P ? *ga.at(f) : base[f];
      if (pre) v = c_mul(v, pre[f]);
      A[z * PITCH + y] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * LP; i += THREADS) {     // z axis: lines = y
      const int g = i / LP, y = i % LP;
      if (y < N) {
        direct_dispatch<N, G, 0>(g, [&](int j) { return A[j * PITCH + y]; },
                                 [&](int k, cplx v) { B[k * PITCH + y] = v; });
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < G * LP; i += THREADS) {     // y axis: lines = kz
      const int g = i / LP, kz = i % LP;
      if (kz < N) {
        const cplx* b = B + kz * PITCH;
        direct_dispatch<N, G, 0>(g, [&](int j) { return b[j]; },
                                 [&](int k, cplx v) {
                                   const int o = k * N + kz;
                                   if (post) { const double w = post[o]; v.x *= w; v.y *= w; }
                                   base[o] = v;
                                 });
      }
    }
    // the next round's loads overwrite A, last read before the previous barrier; B is rewritten after the next one
  }
}

template <class AX, int T, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_lines_direct_kernel(LinesArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  constexpr int N = AX::N, G = AX::G;
  static_assert(T % 32 == 0, "whole warps per group");
  cplx* S = reinterpret_cast<cplx*>(smem_raw);      // S[x][l]
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pr.peer[threadIdx.x];
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    const int tile = (int)(work % p.tiles);
    const long vec = work / p.tiles;
    const long l0 = (long)tile * T;
    const int lcnt = (int)((p.stride - l0 < T) ? (p.stride - l0) : T);
    cplx* base = p.data + vec * p.ldv;
    const cplx* src = base + l0;
    __syncthreads();
#pragma unroll 4
    for (int f = threadIdx.x; f < N * T; f += THREADS) {
      const int x = f / T, l = f % T;
      if (l < lcnt) S[f] = src[(long)x * p.stride + l];
    }
    __syncthreads();
    const XStore<P2P> st{base, p.post, p.stride, peer_s, p.pr.ncol, (p.pr.row0 + vec) * p.pr.ncol, l0};
    for (int i = threadIdx.x; i < G * T; i += THREADS) {
      const int g = i / T, l = i % T;
      if (l < lcnt) {
        direct_dispatch<N, G, 0>(g, [&](int j) { return S[j * T + l]; }, [&](int k, cplx v) { st(k, l, v); });
      }
    }
  }
}

// ---- host side: plan table ---------------------------------------------------------------------------------
struct RegPlan {
  int n;
  size_t plane_smem, lines_smem;
  int (*plane)(Handle*, const PlaneArgs&, bool p2p, cudaStream_t);
  int (*lines)(Handle*, const LinesArgs&, bool p2p, cudaStream_t);
  int T;
};

template <class K>
static inline int resident_grid(Handle* h, K kernel, int threads, size_t smem, long nwork, long* grid) {
  ISDF_CUDA(h, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  ISDF_CUDA(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
  if (per_sm < 1) { snprintf(h->err, sizeof(h->err), "fft kernel does not fit an SM (smem %zu)", smem); return ISDF_ESIZE; }
  long g = (long)h->sm_count * per_sm;
  *grid = g < nwork ? g : nwork;
  return ISDF_OK;
}

template <class K, class A>
static int launch_resident(Handle* h, K kernel, int threads, size_t smem, const A& a, cudaStream_t st) {
  long grid;
  int rc = resident_grid(h, kernel, threads, smem, a.nwork, &grid);
  if (rc) return rc;
  kernel<<<(unsigned)grid, threads, smem, st>>>(a);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
template <class AX, int THREADS, int MINB>
static int launch_plane(Handle* h, const PlaneArgs& a, bool p2p, cudaStream_t st) {
  const size_t smem = (size_t)(AX::SLOTS + AX::N) * sizeof(cplx);
  return p2p ? launch_resident(h, fftreg_plane_kernel<AX, THREADS, MINB, true>, THREADS, smem, a, st)
             : launch_resident(h, fftreg_plane_kernel<AX, THREADS, MINB, false>, THREADS, smem, a, st);
}
template <class AX, int T, int THREADS, int MINB>
static int launch_lines(Handle* h, const LinesArgs& a, bool p2p, cudaStream_t st) {
  const size_t smem = (size_t)(AX::N * T + AX::N) * sizeof(cplx);
  return p2p ? launch_resident(h, fftreg_lines_kernel<AX, T, THREADS, MINB, true>, THREADS, smem, a, st)
             : launch_resident(h, fftreg_lines_kernel<AX, T, THREADS, MINB, false>, THREADS, smem, a, st);
}
template <class AX, int THREADS, int MINB>
static int launch_plane_direct(Handle* h, const PlaneArgs& a, bool p2p, cudaStream_t st) {
  const size_t smem = (size_t)(2 * AX::SLOTS) * sizeof(cplx);
  return p2p ? launch_resident(h, fftreg_plane_direct_kernel<AX, THREADS, MINB, true>, THREADS, smem, a, st)
             : launch_resident(h, fftreg_plane_direct_kernel<AX, THREADS, MINB, false>, THREADS, smem, a, st);
}
template <class AX, int T, int THREADS, int MINB>
static int launch_lines_direct(Handle* h, const LinesArgs& a, bool p2p, cudaStream_t st) {
  const size_t smem = (size_t)(AX::N * T) * sizeof(cplx);
  return p2p ? launch_resident(h, fftreg_lines_direct_kernel<AX, T, THREADS, MINB, true>, THREADS, smem, a, st)
             : launch_resident(h, fftreg_lines_direct_kernel<AX, T, THREADS, MINB, false>, THREADS, smem, a, st);
}

// N = R1 x R2 | plane kernel: threads, min CTAs per SM | x pass: lines per tile, threads, min CTAs per SM
#define ISDF_FFT_TWO(N, R1, R2, PT, PB, T, LT, LB)                                                   \
  {N, 0, 0, launch_plane<TwoFactor<N, R1, R2>, PT, PB>, launch_lines<TwoFactor<N, R1, R2>, T, LT, LB>, T}
#define ISDF_FFT_DIRECT(N, G, PT, PB, T, LT, LB)                                                     \
  {N, 0, 0, launch_plane_direct<Direct<N, G>, PT, PB>, launch_lines_direct<Direct<N, G>, T, LT, LB>, T}


struct RegPlanSlice { const RegPlan* plans; int count; };
}  // namespace fftreg
}  // namespace isdf
