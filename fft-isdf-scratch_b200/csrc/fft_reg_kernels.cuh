// Kernels and launchers of the register-resident FFT (see fft_reg.cu); included by the fft_reg_part*.cu translation
// units, each of which instantiates a slice of the size table (parallel compilation).
#pragma once
#include "fft_reg_core.cuh"

namespace isdf {
namespace fftreg {

// Multi-GPU (P2P = true): the grid columns of every vector are sharded over the ranks, peer[r] = rank r's shard
// [rows][ncol] in NVLink peer-mapped memory (rank r owns grid points [r ncol, (r+1) ncol)).  The plane kernel GATHERS
// its planes from the owning ranks (P2P loads) into the local work buffer `data`, the x pass SCATTERS the weighted
// result back into the shards (P2P stores): both all-to-all exchanges are fused into the transform.  The gather uses
// plain loads into the butterfly registers; a cp.async pipeline that lands the NEXT plane in a second shared-memory
// buffer was measured slower (NiO 33^3 on 2 GPUs: 87 vs 52 ms per build) because it halves the CTAs per SM, and the
// NVLink request rate follows the number of resident CTAs.
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

// global-memory access flavours: 0 default, 1 streaming (evict-first: touched once), 2 L2 only (coherent with other SMs)
template <int KIND>
__device__ __forceinline__ cplx ldg(const cplx* p) {
  if (KIND == 1) return __ldcs(p);
  if (KIND == 2) return __ldcg(p);
  return *p;
}
template <int KIND>
__device__ __forceinline__ void stg(cplx* p, cplx v) {
  if (KIND == 1) __stcs(p, v);
  else *p = v;
}

// x-pass store: grid index g of (vec) -> local vector or the owning rank's shard
template <bool P2P, int STK>
struct XStore {
  cplx* base; const double* post; long stride; cplx* const* peer; long ncol, rowoff; long l0;
  __device__ __forceinline__ void operator()(int k, int l, cplx v) const {
    const long g = (long)k * stride + l0 + l;
    if (post) { const double w = post[g]; v.x *= w; v.y *= w; }
    if (P2P) {
      const unsigned owner = (unsigned)g / (unsigned)ncol;
      peer[owner][rowoff + (g - (long)owner * ncol)] = v;
    } else {
      stg<STK>(base + g, v);
    }
  }
};

// ---- one unit of work per CTA round: "ops" shared by the two-pass kernels and the fused single-launch kernel --------
// FUSED: input loads are streaming, the x pass reads the intermediate through L2 only and streams its result out, so
// that the intermediate written by the plane pass is what stays in L2.  Every op ends with its threads' last shared-
// memory reads still in flight: the caller barriers before the buffer is reused.
template <class AX_, int THREADS>
struct PlaneTwo {
  using AX = AX_;
  static constexpr int WORK = AX::SLOTS;          // complex elements of shared memory besides the twiddles
  template <bool P2P, bool FUSED>
  static __device__ __forceinline__ void run(const PlaneArgs& p, cplx* const* peer_s, long work, long next, cplx* P,
                                             const cplx* TW) {
    constexpr int N = AX::N;
    const int plane = (int)(work % p.n1);
    const long vec = work / p.n1;
    const long poff = (long)plane * N * N;
    cplx* base = p.data + vec * p.ldv + poff;
    (void)next;   // an L2 prefetch of the next plane measured slower here (64^2: 4.67 -> 4.22 TB/s): the pass is not latency-starved
    const cplx* pre = p.pre ? p.pre + poff : nullptr;
    const double* post = p.post ? p.post + poff : nullptr;
    if (P2P) {
      const Gather ga(peer_s, p.pr, vec, poff);
      plane_z1<AX, THREADS>(threadIdx.x, [&](int idx) { cplx v = *ga.at(idx); if (pre) v = c_mul(v, pre[idx]); return v; }, P, TW);
    } else {
      plane_z1<AX, THREADS>(threadIdx.x, [&](int idx) { cplx v = ldg<FUSED ? 1 : 0>(base + idx); if (pre) v = c_mul(v, pre[idx]); return v; }, P, TW);
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
  }
};

// Fire-and-forget L2 prefetch of the NEXT x-pass tile of a persistent CTA (N rows of T consecutive lines, T * 16 bytes
// each, up to 5 cache lines): its demand loads then find L2 instead of DRAM latency.  Pays only for the prime-length
// pass, whose long FP64 phase leaves the memory system idle; the two-factor passes are faster without it.
template <int N, int T, int THREADS>
__device__ __forceinline__ void prefetch_tile_l2(const LinesArgs& p, long next) {
  constexpr int LPR = (T * (int)sizeof(cplx) + 127) / 128 + 1;
  const long l0 = (next % p.tiles) * T;
  const long lcnt = (p.stride - l0 < T) ? (p.stride - l0) : T;
  const cplx* src = p.data + (next / p.tiles) * p.ldv + l0;
  for (int f = threadIdx.x; f < N * LPR; f += THREADS) {
    const int x = f / LPR, i = f % LPR;
    const char* row = (const char*)(src + (long)x * p.stride);
    const char* q = (const char*)((uintptr_t)row & ~(uintptr_t)127) + i * 128;
    if (q < row + lcnt * (long)sizeof(cplx)) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
  }
}

template <class AX_, int T_, int THREADS>
struct LinesTwo {
  using AX = AX_;
  static constexpr int T = T_;
  static constexpr int WORK = AX::N * T_;
  template <bool P2P, bool FUSED>
  static __device__ __forceinline__ void run(const LinesArgs& p, cplx* const* peer_s, long work, long next, cplx* S,
                                             const cplx* TW) {
    const int tile = (int)(work % p.tiles);
    const long vec = work / p.tiles;
    const long l0 = (long)tile * T;
    const int lcnt = (int)((p.stride - l0 < T) ? (p.stride - l0) : T);
    cplx* base = p.data + vec * p.ldv;
    const cplx* src = base + l0;
    const long stride = p.stride;
    (void)next;   // prefetching the next tile measured slower (64: 6.10 -> 5.00 TB/s, the pass already runs at 93 % of the copy rate)
    lines_s1<AX, T, THREADS>(threadIdx.x, [&](int x, int l) { return ldg<FUSED ? 2 : 0>(src + (long)x * stride + l); }, lcnt, S, TW);
    __syncthreads();
    const XStore<P2P, FUSED ? 1 : 0> st{base, p.post, stride, peer_s, p.pr.ncol, (p.pr.row0 + vec) * p.pr.ncol, l0};
    lines_s2<AX, T, THREADS>(threadIdx.x, S, lcnt, st);
  }
};

// prime lengths
template <class AX_, int THREADS>
struct PlaneDirect {
  using AX = AX_;
  static constexpr int WORK = 2 * AX::SLOTS;
  template <bool P2P, bool FUSED>
  static __device__ __forceinline__ void run(const PlaneArgs& p, cplx* const* peer_s, long work, long next, cplx* A,
                                             const cplx*) {
    constexpr int N = AX::N, G = AX::G, PITCH = AX::PITCH;
    (void)next;   // an L2 prefetch of the next plane measured slower here (64^2: 4.67 -> 4.22 TB/s): the pass is not latency-starved
    constexpr int LP = (N + 31) / 32 * 32;            // lines padded to whole warps: the group index is warp-uniform
    cplx* B = A + AX::SLOTS;                          // A[z][y], B[kz][y]
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
      cplx v = P2P ? *ga.at(f) : ldg<FUSED ? 1 : 0>(base + f);
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
  }
};

template <class AX_, int T_, int THREADS>
struct LinesDirect {
  using AX = AX_;
  static constexpr int T = T_;
  static constexpr int WORK = AX::N * T_;
  static_assert(T_ % 32 == 0, "whole warps per group");
  template <bool P2P, bool FUSED>
  static __device__ __forceinline__ void run(const LinesArgs& p, cplx* const* peer_s, long work, long next, cplx* S,
                                             const cplx*) {
    constexpr int N = AX::N, G = AX::G;
    if (next >= 0) prefetch_tile_l2<N, T, THREADS>(p, next);   // measured: 37-point x pass 3.24 -> 4.50 TB/s
    const int tile = (int)(work % p.tiles);
    const long vec = work / p.tiles;
    const long l0 = (long)tile * T;
    const int lcnt = (int)((p.stride - l0 < T) ? (p.stride - l0) : T);
    cplx* base = p.data + vec * p.ldv;
    const cplx* src = base + l0;
#pragma unroll 4
    for (int f = threadIdx.x; f < N * T; f += THREADS) {
      const int x = f / T, l = f % T;
      if (l < lcnt) S[f] = ldg<FUSED ? 2 : 0>(src + (long)x * p.stride + l);
    }
    __syncthreads();
    const XStore<P2P, FUSED ? 1 : 0> st{base, p.post, p.stride, peer_s, p.pr.ncol, (p.pr.row0 + vec) * p.pr.ncol, l0};
    for (int i = threadIdx.x; i < G * T; i += THREADS) {
      const int g = i / T, l = i % T;
      if (l < lcnt) {
        direct_dispatch<N, G, 0>(g, [&](int j) { return S[j * T + l]; }, [&](int k, cplx v) { st(k, l, v); });
      }
    }
  }
};

// ---- two-pass kernels: persistent CTAs, static round-robin over the units ------------------------------------
template <class OP, class ARGS, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_pass_kernel(ARGS p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  constexpr int N = OP::AX::N;
  cplx* TW = reinterpret_cast<cplx*>(smem_raw);
  cplx* W = TW + N;
  for (int i = threadIdx.x; i < N; i += THREADS) TW[i] = p.tw[i];
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pr.peer[threadIdx.x];
  __syncthreads();
  for (long work = blockIdx.x; work < p.nwork; work += gridDim.x) {
    OP::template run<P2P, false>(p, peer_s, work, (work + gridDim.x < p.nwork) ? work + gridDim.x : -1, W, TW);
    __syncthreads();
  }
}

// ---- fused single launch: plane pass and x pass of the whole batch in ONE persistent kernel ------------------------
// The batch is cut into groups of gv vectors (a few tens of MB).  Units are handed out by a global ticket counter in
// the order P(0), P(1), X(0), P(2), X(1), ...: an x unit of group g waits (acquire on done[g]) until all plane units of
// its group have been published (release), which by ticket order are already running or finished -- no CTA ever waits
// on work that has not been handed out, so the scheme cannot deadlock, and the intermediate of at most ~3 groups is
// live in L2 between its write and its only read: HBM sees each element once in and once out.
struct FusedArgs {
  PlaneArgs pl;
  LinesArgs ln;
  int* sync;              // [0] ticket, [1 + g] plane units of group g published
  int* err;               // host-mapped sticky flag: a dependency wait gave up
  long nvec;
  int gv, ngroups;
  int np, nx;             // plane / x units per (full) group
};

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <class POP, class LOP, int THREADS, int MINB, bool P2P>
__global__ void __launch_bounds__(THREADS, MINB) fftreg_fused_kernel(FusedArgs p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ cplx* peer_s[8];
  __shared__ int ticket_s;
  constexpr int NZ = POP::AX::N, NX = LOP::AX::N;
  cplx* TWz = reinterpret_cast<cplx*>(smem_raw);
  cplx* TWx = TWz + NZ;
  cplx* W = TWx + NX;
  for (int i = threadIdx.x; i < NZ; i += THREADS) TWz[i] = p.pl.tw[i];
  for (int i = threadIdx.x; i < NX; i += THREADS) TWx[i] = p.ln.tw[i];
  if (P2P && threadIdx.x < 8) peer_s[threadIdx.x] = p.pl.pr.peer[threadIdx.x];
  const long seg = (long)p.np + p.nx;
  const long total = (long)p.ngroups * seg;
  for (;;) {
    __syncthreads();                                   // previous unit's shared-memory reads are done; ticket_s is free
    if (threadIdx.x == 0) ticket_s = atomicAdd(p.sync, 1);
    __syncthreads();
    const long t = ticket_s;
    if (t >= total) break;
    // ticket order: P(0) | P(1) X(0) | P(2) X(1) | ... | X(G-1)
    int g; long r; bool is_plane;
    if (t < p.np) { g = 0; r = t; is_plane = true; }
    else {
      const long u = t - p.np;
      const int s = (int)(u / seg) + 1;
      r = u - (long)(s - 1) * seg;
      if (s < p.ngroups && r < p.np) { g = s; is_plane = true; }
      else { g = s - 1; if (s < p.ngroups) r -= p.np; is_plane = false; }
    }
    if (is_plane) {
      const long vec = (long)g * p.gv + r / p.pl.n1;
      if (vec < p.nvec) POP::template run<P2P, true>(p.pl, peer_s, vec * p.pl.n1 + r % p.pl.n1, -1, W, TWz);
      __threadfence();                                 // this thread's stores are visible device-wide ...
      __syncthreads();
      if (threadIdx.x == 0) atomicAdd(p.sync + 1 + g, 1);   // ... before the unit is published
    } else {
      if (threadIdx.x == 0) {
        int spins = 0;
        while (ld_acquire(p.sync + 1 + g) < p.np) {
          __nanosleep(128);
          if (++spins > (1 << 24)) { *((volatile int*)p.err) = 1; break; }
        }
      }
      __syncthreads();
      const long vec = (long)g * p.gv + r / p.ln.tiles;
      if (vec < p.nvec) LOP::template run<P2P, true>(p.ln, peer_s, vec * p.ln.tiles + r % p.ln.tiles, -1, W, TWx);
    }
  }
}

// ---- host side: plan table ---------------------------------------------------------------------------------
struct RegPlan {
  int n;
  int (*plane)(Handle*, const PlaneArgs&, bool p2p, cudaStream_t);
  int (*lines)(Handle*, const LinesArgs&, bool p2p, cudaStream_t);
  int (*fused)(Handle*, const FusedArgs&, bool p2p, cudaStream_t);     // cubic meshes: n1 == n2 == n3 == n
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
static int launch_resident(Handle* h, K kernel, int threads, size_t smem, const A& a, long nwork, cudaStream_t st) {
  long grid;
  int rc = resident_grid(h, kernel, threads, smem, nwork, &grid);
  if (rc) return rc;
  kernel<<<(unsigned)grid, threads, smem, st>>>(a);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
template <class OP, class ARGS, int THREADS, int MINB>
static int launch_pass(Handle* h, const ARGS& a, bool p2p, cudaStream_t st) {
  const size_t smem = (size_t)(OP::WORK + OP::AX::N) * sizeof(cplx);
  return p2p ? launch_resident(h, fftreg_pass_kernel<OP, ARGS, THREADS, MINB, true>, THREADS, smem, a, a.nwork, st)
             : launch_resident(h, fftreg_pass_kernel<OP, ARGS, THREADS, MINB, false>, THREADS, smem, a, a.nwork, st);
}
template <class POP, class LOP, int THREADS, int MINB>
static int launch_fused(Handle* h, const FusedArgs& a, bool p2p, cudaStream_t st) {
  constexpr int WORK = POP::WORK > LOP::WORK ? POP::WORK : LOP::WORK;
  const size_t smem = (size_t)(WORK + POP::AX::N + LOP::AX::N) * sizeof(cplx);
  const long total = (long)a.ngroups * ((long)a.np + a.nx);
  return p2p ? launch_resident(h, fftreg_fused_kernel<POP, LOP, THREADS, MINB, true>, THREADS, smem, a, total, st)
             : launch_resident(h, fftreg_fused_kernel<POP, LOP, THREADS, MINB, false>, THREADS, smem, a, total, st);
}

// N = R1 x R2 | plane kernel: threads, min CTAs per SM | x pass: lines per tile, threads, min CTAs per SM
#define ISDF_FFT_TWO(N, R1, R2, PT, PB, T, LT, LB)                                                        \
  {N, launch_pass<PlaneTwo<TwoFactor<N, R1, R2>, PT>, PlaneArgs, PT, PB>,                                 \
   launch_pass<LinesTwo<TwoFactor<N, R1, R2>, T, LT>, LinesArgs, LT, LB>,                                 \
   launch_fused<PlaneTwo<TwoFactor<N, R1, R2>, PT>, LinesTwo<TwoFactor<N, R1, R2>, T, PT>, PT, PB>, T}
#define ISDF_FFT_DIRECT(N, G, PT, PB, T, LT, LB)                                                          \
  {N, launch_pass<PlaneDirect<Direct<N, G>, PT>, PlaneArgs, PT, PB>,                                      \
   launch_pass<LinesDirect<Direct<N, G>, T, LT>, LinesArgs, LT, LB>,                                      \
   launch_fused<PlaneDirect<Direct<N, G>, PT>, LinesDirect<Direct<N, G>, T, PT>, PT, PB>, T}

struct RegPlanSlice { const RegPlan* plans; int count; };
}  // namespace fftreg
}  // namespace isdf
