// Batched diagonal-pivoted Cholesky (LAPACK xPSTRF semantics) on Hermitian PSD matrices, plus the
// helpers that turn its factor into the blocked triangular-sweep operators of the Theta fit.
//
// Serves two reference call sites:
//   * point selection: pyscf.lib.scipy_helper.pivoted_cholesky(x4) -> scipy dpstrf
//     (/root/reference/fftisdf.py:381-384).  x4 is real; it is carried as complex with zero
//     imaginary parts, which makes every operation below bit-identical to the real recurrence.
//   * the per-q least-squares fit scipy.linalg.lstsq(A_q, Y_q^T, "gelsy") (fftisdf.py:108): A_q is
//     Hermitian PSD, so a rank-revealing Cholesky A_q[piv,piv] = U^H U replaces QRCP.
//
// Pivot rule (dpstrf): at step j take the FIRST maximum, in current position order, of the
// residual diagonal a_ii - sum_{t in panel} |u_ti|^2; stop when it is <= n*eps*max_i a_ii (tol<0)
// or <= tol.  Positions are tracked instead of physically swapping rows/columns.
//
// Structure: right-looking with panels of nb steps.  The panel kernel (one CTA per matrix, all
// matrices of the batch concurrently) produces nb rows of U; the trailing update
// A -= U_panel^H U_panel runs on the DMMA GEMM engine.
#include <float.h>
#include <type_traits>
#include <cooperative_groups.h>
#include "gemm_c128.cuh"

namespace cg = cooperative_groups;

namespace isdf {

constexpr int PC_THREADS = 1024;
constexpr int PC_NCOL = 8;  // columns per thread -> n <= 8192
constexpr int PC_NB_MAX = 64;

struct PcholInfo {
  int rank;
  int done;
  double dstop;
  double next;  // value of the pivot that would come next (residual estimate, fftisdf.py:387)
};

__global__ void pchol_init_kernel(int* pos, PcholInfo* info, int* active, int n, int batch) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long)n * batch) pos[i] = (int)(i % n);
  if (i < batch) {
    info[i].rank = 0; info[i].done = 0; info[i].dstop = 0.0; info[i].next = 0.0;
    active[i] = 1;
  }
}

__global__ void __launch_bounds__(PC_THREADS, 1)
pchol_panel_kernel(const cplx* __restrict__ Aall, long lda, long strideA, int n, int j0, int nb, int max_steps,
                   double tol, cplx* Uall, long ldu, long strideU, int* posall, PcholInfo* infoall, int* active,
                   int nopivot) {
  const int b = blockIdx.x;
  PcholInfo* info = infoall + b;
  if (info->done) return;
  const cplx* A = Aall + (long)b * strideA;
  cplx* U = Uall + (long)b * strideU;
  int* pos = posall + (long)b * n;

  __shared__ cplx bp[PC_NB_MAX];
  __shared__ double red_v[32];
  __shared__ int red_pos[32];
  __shared__ int red_idx[32];
  __shared__ double w_v;
  __shared__ int w_pos, w_idx;

  extern __shared__ double s_aii[];  // [n] diagonal at panel start (dpstrf: work(n+i) = a_ii - work(i))
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double ssum[PC_NCOL];
  int mypos[PC_NCOL];
#pragma unroll
  for (int c = 0; c < PC_NCOL; ++c) {
    const int i = tid + c * PC_THREADS;
    ssum[c] = 0.0;
    if (i < n) {
      s_aii[i] = A[(long)i * lda + i].x;
      mypos[c] = pos[i];
    } else {
      mypos[c] = -1;
    }
  }
  double dstop = info->dstop;
  int steps_done = 0;
  bool stopped = false;

  for (int t = 0; t <= nb; ++t) {
    const int j = j0 + t;
    if (t == nb && j < max_steps) break;  // panel complete; more panels follow
    // ---- 1. first maximum of the residual diagonal over positions >= j
    double bv = -DBL_MAX;
    int bpos = 0x7fffffff, bidx = -1;
#pragma unroll
    for (int c = 0; c < PC_NCOL; ++c) {
      if (nopivot ? (mypos[c] == j) : (mypos[c] >= j)) {
        const double d = s_aii[tid + c * PC_THREADS] - ssum[c];
        if (d > bv || (d == bv && mypos[c] < bpos)) { bv = d; bpos = mypos[c]; bidx = tid + c * PC_THREADS; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_pos[warp] = bpos; red_idx[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
      bv = red_v[lane]; bpos = red_pos[lane]; bidx = red_idx[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
      }
      if (lane == 0) { w_v = bv; w_pos = bpos; w_idx = bidx; }
    }
    __syncthreads();
    const double dp = w_v;
    const int p = w_idx, ppos = w_pos;
    // ---- 2. stopping rule (dpstrf: ajj <= dstop or NaN; first pivot must be > 0)
    if (j == 0) dstop = (tol < 0.0) ? (double)n * DBL_EPSILON * dp : tol;
    if (p < 0 || j >= max_steps || !(dp > dstop) || !(dp > 0.0)) {
      if (tid == 0) {
        info->rank = j; info->done = 1; info->dstop = dstop;
        info->next = (p < 0) ? 0.0 : dp;
        active[b] = 0;
      }
      stopped = true;
      break;
    }
    // ---- 3. position bookkeeping (swap positions j <-> ppos)
#pragma unroll
    for (int c = 0; c < PC_NCOL; ++c) {
      const int i = tid + c * PC_THREADS;
      if (i == p) mypos[c] = j;
      else if (mypos[c] == j) mypos[c] = ppos;
    }
    // ---- 4. broadcast the pivot column's in-panel entries u_{t',p}
    if (tid < t) bp[tid] = __ldcg(&U[(long)(j0 + tid) * ldu + p]);
    __syncthreads();
    // ---- 5. new row of U, residual update.  The t-loop is outermost so that the PC_NCOL independent
    //         column updates of a thread overlap their U loads (latency-bound otherwise).
    const double rt = sqrt(dp);
    const double inv = 1.0 / rt;
    const cplx* Arow = A + (long)p * lda;
#pragma unroll
    for (int c0 = 0; c0 < PC_NCOL; c0 += 4) {
      if (tid + c0 * PC_THREADS >= n) break;
      cplx v[4];
      bool act[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = tid + (c0 + c) * PC_THREADS;
        act[c] = (i < n) && (i != p) && (mypos[c0 + c] > j);
        // row p of the Hermitian matrix from its LOWER triangle: A[p][i] (i < p) or conj(A[i][p]) (i > p)
        v[c] = make_double2(0.0, 0.0);
        if (act[c]) {
          if (i < p) v[c] = Arow[i];
          else { const cplx w = A[(long)i * lda + p]; v[c] = make_double2(w.x, -w.y); }
        }
      }
      for (int tt = 0; tt < t; ++tt) {
        const cplx bb = bp[tt];
        const cplx* Urow = U + (long)(j0 + tt) * ldu + tid + c0 * PC_THREADS;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (act[c]) {
            const cplx ui = Urow[c * PC_THREADS];
            // v -= conj(bb) * ui
            v[c].x -= bb.x * ui.x + bb.y * ui.y;
            v[c].y -= bb.x * ui.y - bb.y * ui.x;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = tid + (c0 + c) * PC_THREADS;
        if (i < n) {
          cplx u;
          if (i == p) {
            u = make_double2(rt, 0.0);
          } else if (act[c]) {
            u = make_double2(v[c].x * inv, v[c].y * inv);
            ssum[c0 + c] += u.x * u.x + u.y * u.y;
          } else {
            u = make_double2(0.0, 0.0);
          }
          U[(long)j * ldu + i] = u;
        }
      }
    }
    steps_done = t + 1;
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < PC_NCOL; ++c) {
    const int i = tid + c * PC_THREADS;
    if (i < n) pos[i] = mypos[c];
  }
  if (!stopped && tid == 0) { info->rank = j0 + steps_done; info->dstop = dstop; }
}

// ---------------------------------------------------------------------------------------------
// Cluster version of the panel kernel for large matrices: one thread-block cluster of PCC_CS CTAs per
// matrix, each CTA owning a contiguous slice of columns whose in-panel rows of U stay in ITS shared
// memory.  Per step there are two cluster barriers: (1) every CTA publishes its best candidate into
// every CTA's shared memory (DSMEM stores) and all reduce the PCC_CS candidates identically; (2) the
// CTA that owns the pivot column broadcasts that column's in-panel entries.  Same arithmetic and pivot
// rule as pchol_panel_kernel.
constexpr int PCC_CS = 8;        // portable cluster size
constexpr int PCC_THREADS = 512;
constexpr int PCC_NCOL = 2;      // columns per thread -> ncc <= 1024, n <= 8192

struct PcCand { double v; int pos; int idx; };

// REALP: the matrix is real (imaginary parts exactly zero, e.g. the selection matrix x4): the in-panel rows are
// kept as doubles, which doubles the panel width that fits in shared memory (half as many trailing updates).
template <bool REALP>
__global__ void __cluster_dims__(PCC_CS, 1, 1) __launch_bounds__(PCC_THREADS, 1)
pchol_panel_cluster_kernel(const cplx* __restrict__ Aall, long lda, long strideA, int n, int ncc, int j0, int nb,
                           int max_steps, double tol, cplx* Uall, long ldu, long strideU, int* posall,
                           PcholInfo* infoall, int* active, int nopivot) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int b = blockIdx.y;
  PcholInfo* info = infoall + b;
  if (info->done) return;   // uniform over the cluster (flag is only written after a cluster barrier)
  const cplx* A = Aall + (long)b * strideA;
  cplx* U = Uall + (long)b * strideU;
  int* pos = posall + (long)b * n;

  extern __shared__ __align__(16) unsigned char pcc_smem[];
  typedef typename std::conditional<REALP, double, cplx>::type pan_t;
  pan_t* up = reinterpret_cast<pan_t*>(pcc_smem);               // [nb][ncc] in-panel rows of U, own columns
  double* s_aii = reinterpret_cast<double*>(up + (long)nb * ncc);  // [ncc]
  // candidates and their in-panel columns, double-buffered by step parity: a CTA that runs one step ahead writes
  // the other buffer, and it cannot run two steps ahead because of the per-step cluster barrier
  __shared__ cplx bpc[2][PCC_CS][PC_NB_MAX];
  __shared__ PcCand cand[2][PCC_CS];
  __shared__ int s_best;                    // this CTA's candidate column (global index), -1 if none
  __shared__ double red_v[PCC_THREADS / 32];
  __shared__ int red_pos[PCC_THREADS / 32];
  __shared__ int red_idx[PCC_THREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c_lo = crank * ncc;
  double ssum[PCC_NCOL];
  int mypos[PCC_NCOL];
#pragma unroll
  for (int c = 0; c < PCC_NCOL; ++c) {
    const int lc = tid + c * PCC_THREADS;
    const int i = c_lo + lc;
    ssum[c] = 0.0;
    if (lc < ncc && i < n) {
      s_aii[lc] = A[(long)i * lda + i].x;
      mypos[c] = pos[i];
    } else {
      mypos[c] = -1;
    }
  }
  double dstop = info->dstop;
  int steps_done = 0;
  bool stopped = false;

  for (int t = 0; t <= nb; ++t) {
    const int j = j0 + t;
    if (t == nb && j < max_steps) break;
    // ---- 1. local candidate
    double bv = -DBL_MAX;
    int bpos = 0x7fffffff, bidx = -1;
#pragma unroll
    for (int c = 0; c < PCC_NCOL; ++c) {
      if (nopivot ? (mypos[c] == j) : (mypos[c] >= j)) {
        const int lc = tid + c * PCC_THREADS;
        const double d = s_aii[lc] - ssum[c];
        if (d > bv || (d == bv && mypos[c] < bpos)) { bv = d; bpos = mypos[c]; bidx = c_lo + lc; }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
    }
    if (lane == 0) { red_v[warp] = bv; red_pos[warp] = bpos; red_idx[warp] = bidx; }
    __syncthreads();
    if (warp == 0) {
      if (lane < PCC_THREADS / 32) { bv = red_v[lane]; bpos = red_pos[lane]; bidx = red_idx[lane]; }
      else { bv = -DBL_MAX; bpos = 0x7fffffff; bidx = -1; }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
      }
      if (lane < PCC_CS) {   // lane r publishes this CTA's candidate into CTA r's table
        PcCand (*remote)[PCC_CS] = cluster.map_shared_rank(cand, lane);
        remote[t & 1][crank].v = bv; remote[t & 1][crank].pos = bpos; remote[t & 1][crank].idx = bidx;
      }
      if (lane == 0) s_best = bidx;
    }
    __syncthreads();
    // publish the candidate's in-panel column next to it, so that ONE cluster barrier per step suffices: after it
    // every CTA knows the winner and already holds the winner's column u_{0..t-1, p}
    {
      const int bl = s_best - c_lo;
      if (s_best >= 0) {
        for (int w = tid; w < t * PCC_CS; w += PCC_THREADS) {
          const int tt = w / PCC_CS, r = w - tt * PCC_CS;
          cplx (*remote)[PCC_CS][PC_NB_MAX] = cluster.map_shared_rank(bpc, r);
          if constexpr (REALP) remote[t & 1][crank][tt] = make_double2(up[(long)tt * ncc + bl], 0.0);
          else remote[t & 1][crank][tt] = up[(long)tt * ncc + bl];
        }
      }
    }
    cluster.sync();
    // ---- 2. identical reduction of the PCC_CS candidates in every CTA
    double dp = -DBL_MAX;
    int ppos = 0x7fffffff, p = -1;
#pragma unroll
    for (int r = 0; r < PCC_CS; ++r) {
      const double ov = cand[t & 1][r].v;
      const int op = cand[t & 1][r].pos, oi = cand[t & 1][r].idx;
      if (ov > dp || (ov == dp && op < ppos)) { dp = ov; ppos = op; p = oi; }
    }
    if (j == 0) dstop = (tol < 0.0) ? (double)n * DBL_EPSILON * dp : tol;
    if (p < 0 || j >= max_steps || !(dp > dstop) || !(dp > 0.0)) {
      if (crank == 0 && tid == 0) {
        info->rank = j; info->done = 1; info->dstop = dstop;
        info->next = (p < 0) ? 0.0 : dp;
        active[b] = 0;
      }
      stopped = true;
      break;
    }
    // ---- 3. positions
#pragma unroll
    for (int c = 0; c < PCC_NCOL; ++c) {
      const int i = c_lo + tid + c * PCC_THREADS;
      if (mypos[c] >= 0) {
        if (i == p) mypos[c] = j;
        else if (mypos[c] == j) mypos[c] = ppos;
      }
    }
    // ---- 4a. issue the loads of row p (lower triangle) now: their latency overlaps the broadcast + barrier
    const double rt = sqrt(dp);
    const double inv = 1.0 / rt;
    const cplx* Arow = A + (long)p * lda;
    cplx v[PCC_NCOL];
    bool act[PCC_NCOL];
#pragma unroll
    for (int c = 0; c < PCC_NCOL; ++c) {
      const int lc = tid + c * PCC_THREADS;
      const int i = c_lo + lc;
      act[c] = (lc < ncc) && (i < n) && (i != p) && (mypos[c] > j);
      v[c] = make_double2(0.0, 0.0);
      if (act[c]) {
        if (i < p) v[c] = Arow[i];
        else { const cplx w = A[(long)i * lda + p]; v[c] = make_double2(w.x, -w.y); }
      }
    }
    // ---- 4b. the winner's column arrived with its candidate
    const cplx* bp = bpc[t & 1][p / ncc];
    // ---- 5. new row of U for the own columns
    for (int tt = 0; tt < t; ++tt) {
      const cplx bb = bp[tt];
#pragma unroll
      for (int c = 0; c < PCC_NCOL; ++c) {
        if (act[c]) {
          if constexpr (REALP) {
            v[c].x -= bb.x * up[(long)tt * ncc + tid + c * PCC_THREADS];
          } else {
            const cplx ui = up[(long)tt * ncc + tid + c * PCC_THREADS];
            v[c].x -= bb.x * ui.x + bb.y * ui.y;   // v -= conj(bb) * ui
            v[c].y -= bb.x * ui.y - bb.y * ui.x;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < PCC_NCOL; ++c) {
      const int lc = tid + c * PCC_THREADS;
      const int i = c_lo + lc;
      if (lc < ncc && i < n) {
        cplx u;
        if (i == p) {
          u = make_double2(rt, 0.0);
        } else if (act[c]) {
          u = make_double2(v[c].x * inv, v[c].y * inv);
          ssum[c] += u.x * u.x + u.y * u.y;
        } else {
          u = make_double2(0.0, 0.0);
        }
        if (t < nb) {
          if constexpr (REALP) up[(long)t * ncc + lc] = u.x;
          else up[(long)t * ncc + lc] = u;
        }
        U[(long)j * ldu + i] = u;
      }
    }
    steps_done = t + 1;
    // no barrier here: up[t][*] is ordered before the next step's publish by that step's block barrier, and the
    // cand/bpc buffers of this step are not overwritten before the next-but-one step (double buffering)
  }
#pragma unroll
  for (int c = 0; c < PCC_NCOL; ++c) {
    const int lc = tid + c * PCC_THREADS;
    const int i = c_lo + lc;
    if (lc < ncc && i < n) pos[i] = mypos[c];
  }
  if (!stopped && crank == 0 && tid == 0) { info->rank = j0 + steps_done; info->dstop = dstop; }
  // keep every CTA's shared memory alive until all remote stores into it have been consumed
  cluster.sync();
}

__global__ void pchol_finalize_kernel(const int* pos, const PcholInfo* info, int n, int batch, int* piv, int* rank,
                                      double* next) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long)n * batch) {
    const long b = i / n;
    piv[b * n + pos[i]] = (int)(i % n);
  }
  if (i < batch) {
    rank[i] = info[i].rank;
    if (next) next[i] = info[i].next;
  }
}

// ---------------------------------------------------------------------------------------------
// Blocked UNPIVOTED Cholesky A = U^H U (right-looking, 64-wide block columns) for the well-conditioned Gram matrices of
// the Cholesky-QR step: per block one CTA factorises the 64 x 64 diagonal block in shared memory (A_jj = L L^H,
// U_jj = L^H) and inverts L; the block row U_j,rest = L^-1 A_j,rest and the trailing update run on the GEMM engine.
// Stops at the first pivot <= tol (rows of an exactly singular trailing block: rank < n), like the pivoted kernels.
constexpr int CB = 64;
__global__ void __launch_bounds__(256) chol_diag_kernel(const cplx* __restrict__ Aall, long strideA, int n, int j0,
                                                        double tol, cplx* __restrict__ Uall, long strideU,
                                                        cplx* __restrict__ Winv, PcholInfo* infoall, int* active) {
  const int b = blockIdx.x;
  PcholInfo* info = infoall + b;
  cplx* W = Winv + (long)b * CB * CB;
  const int tid = threadIdx.x;
  const int nb = min(CB, n - j0);
  if (info->done) {
    for (int e = tid; e < CB * CB; e += 256) W[e] = make_double2(0.0, 0.0);
    return;
  }
  extern __shared__ __align__(16) unsigned char cd_smem[];
  cplx* S = reinterpret_cast<cplx*>(cd_smem);       // [CB][CB+1]  lower triangle: L
  cplx* X = S + CB * (CB + 1);                      // [CB][CB+1]  L^-1
  __shared__ int s_stop;
  const cplx* A = Aall + (long)b * strideA + (long)j0 * n + j0;
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e / CB, c = e - r * CB;
    S[r * (CB + 1) + c] = (r < nb && c <= r) ? A[(long)r * n + c] : make_double2(0.0, 0.0);
    X[r * (CB + 1) + c] = make_double2(0.0, 0.0);
  }
  if (tid == 0) s_stop = nb;
  __syncthreads();
  for (int k = 0; k < nb; ++k) {
    const double d = S[k * (CB + 1) + k].x;
    if (!(d > tol) || !(d > 0.0)) { if (tid == 0) s_stop = k; break; }       // uniform: d comes from shared memory
    const double lkk = sqrt(d), inv = 1.0 / lkk;
    __syncthreads();
    for (int i = k + tid; i < nb; i += 256) {
      cplx v = S[i * (CB + 1) + k];
      S[i * (CB + 1) + k] = (i == k) ? make_double2(lkk, 0.0) : make_double2(v.x * inv, v.y * inv);
    }
    __syncthreads();
    // trailing lower triangle: S[i][c] -= L[i][k] conj(L[c][k]),  i >= c > k
    const int m = nb - k - 1;
    for (int e = tid; e < m * m; e += 256) {
      const int i = k + 1 + e / m, c = k + 1 + e % m;
      if (c <= i) {
        const cplx li = S[i * (CB + 1) + k], lc = S[c * (CB + 1) + k];
        cplx v = S[i * (CB + 1) + c];
        v.x -= li.x * lc.x + li.y * lc.y;
        v.y -= li.y * lc.x - li.x * lc.y;
        S[i * (CB + 1) + c] = v;
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int kk = s_stop;                              // leading kk x kk block is factorised
  // X = L^-1 (lower), one column per thread by forward substitution
  if (tid < kk) {
    const int c = tid;
    for (int i = c; i < kk; ++i) {
      cplx acc = make_double2((i == c) ? 1.0 : 0.0, 0.0);
      for (int k = c; k < i; ++k) {
        const cplx l = S[i * (CB + 1) + k], x = X[k * (CB + 1) + c];
        acc.x -= l.x * x.x - l.y * x.y;
        acc.y -= l.x * x.y + l.y * x.x;
      }
      const double dinv = 1.0 / S[i * (CB + 1) + i].x;
      X[i * (CB + 1) + c] = make_double2(acc.x * dinv, acc.y * dinv);
    }
  }
  __syncthreads();
  cplx* U = Uall + (long)b * strideU + (long)j0 * n + j0;
  for (int e = tid; e < CB * CB; e += 256) {
    const int r = e / CB, c = e - r * CB;
    W[e] = (r < kk && c <= r) ? X[r * (CB + 1) + c] : make_double2(0.0, 0.0);
    if (r < nb && c < nb) {
      cplx u = make_double2(0.0, 0.0);
      if (r < kk && c >= r) { const cplx l = S[c * (CB + 1) + r]; u = make_double2(l.x, -l.y); }   // U_jj = L^H
      U[(long)r * n + c] = u;
    }
  }
  if (tid == 0) {
    if (kk < nb) { info->rank = j0 + kk; info->done = 1; active[b] = 0; }
    else info->rank = j0 + nb;
  }
}

// ---------------------------------------------------------------------------------------------
// Triangular-sweep operator construction.
// Up[a][b] = U[a][piv[b]] for a<=b<rank, identity on [rank, nP); Lp = Up^H.
__global__ void build_up_lp_kernel(const cplx* Uall, long ldu, long strideU, const int* pivall, const int* rank, int n,
                                   int nP, cplx* UpAll, cplx* LpAll) {
  const int b = blockIdx.z;
  const cplx* U = Uall + (long)b * strideU;
  const int* piv = pivall + (long)b * n;
  const int r = rank[b];
  cplx* Up = UpAll + (long)b * nP * nP;
  cplx* Lp = LpAll + (long)b * nP * nP;
  __shared__ cplx tile[32][33];
  const int a0 = blockIdx.y * 32, b0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int yy = ty; yy < 32; yy += 8) {
    const int a = a0 + yy, c = b0 + tx;
    cplx v = make_double2(0.0, 0.0);
    if (a < nP && c < nP) {
      if (a < r && c < r && c >= a) v = U[(long)a * ldu + piv[c]];
      else if (a == c && a >= r) v = make_double2(1.0, 0.0);
      Up[(long)a * nP + c] = v;
    }
    tile[yy][tx] = v;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {
    const int c = b0 + yy, a = a0 + tx;  // Lp[c][a] = conj(Up[a][c])
    if (a < nP && c < nP) {
      const cplx v = tile[tx][yy];
      Lp[(long)c * nP + a] = make_double2(v.x, -v.y);
    }
  }
}

// Invert the 64x64 upper-triangular diagonal blocks of Up; write Upinv into Ubwd's diagonal block
// and its conjugate transpose into Lfwd's diagonal block.
constexpr int TB = 64;
__global__ void __launch_bounds__(TB) tri_inv_blocks_kernel(const cplx* UpAll, int nP, cplx* UbwdAll, cplx* LfwdAll) {
  const int blk = blockIdx.x, b = blockIdx.y;
  const cplx* Up = UpAll + (long)b * nP * nP + (long)blk * TB * nP + blk * TB;
  cplx* Ub = UbwdAll + (long)b * nP * nP + (long)blk * TB * nP + blk * TB;
  cplx* Lf = LfwdAll + (long)b * nP * nP + (long)blk * TB * nP + blk * TB;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sU = reinterpret_cast<cplx*>(smem_raw);  // [TB][TB+1]
  cplx* sX = sU + TB * (TB + 1);                 // [TB][TB+1]  X[i][c]
  const int c = threadIdx.x;
  for (int i = 0; i < TB; ++i) sU[i * (TB + 1) + c] = Up[(long)i * nP + c];
  __syncthreads();
  // column c of the inverse: solve Up x = e_c by back substitution (x_i = 0 for i > c)
  for (int i = TB - 1; i >= 0; --i) {
    cplx acc = make_double2((i == c) ? 1.0 : 0.0, 0.0);
    if (i <= c) {
      for (int k = i + 1; k <= c; ++k) {
        const cplx u = sU[i * (TB + 1) + k];
        const cplx x = sX[k * (TB + 1) + c];
        acc.x -= u.x * x.x - u.y * x.y;
        acc.y -= u.x * x.y + u.y * x.x;
      }
      const cplx d = sU[i * (TB + 1) + i];
      const double den = d.x * d.x + d.y * d.y;
      const cplx q = make_double2((acc.x * d.x + acc.y * d.y) / den, (acc.y * d.x - acc.x * d.y) / den);
      sX[i * (TB + 1) + c] = q;
    } else {
      sX[i * (TB + 1) + c] = make_double2(0.0, 0.0);
    }
  }
  __syncthreads();
  for (int i = 0; i < TB; ++i) {
    const cplx x = sX[i * (TB + 1) + c];
    Ub[(long)i * nP + c] = x;                       // Upinv[i][c]
    const cplx y = sX[c * (TB + 1) + i];            // Upinv[c][i]
    Lf[(long)i * nP + c] = make_double2(y.x, -y.y); // Linv[i][c] = conj(Upinv[c][i])
  }
}

// Off-diagonal blocks: Ubwd[a][cb] = -Upinv_aa * Up[a][cb] (cb > a), Lfwd[a][cb] = -Linv_aa * Lp[a][cb] (cb < a).
__global__ void __launch_bounds__(256) apply_diag_inv_kernel(const cplx* UpAll, const cplx* LpAll, int nP, cplx* UbwdAll,
                                                             cplx* LfwdAll) {
  const int nblk = nP / TB;
  const int a = blockIdx.y, cb = blockIdx.x, b = blockIdx.z;
  if (cb == a) return;
  const bool upper = cb > a;
  const cplx* Src = (upper ? UpAll : LpAll) + (long)b * nP * nP + (long)a * TB * nP + cb * TB;
  cplx* Dst = (upper ? UbwdAll : LfwdAll) + (long)b * nP * nP + (long)a * TB * nP + cb * TB;
  const cplx* Inv = (upper ? UbwdAll : LfwdAll) + (long)b * nP * nP + (long)a * TB * nP + a * TB;
  (void)nblk;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sI = reinterpret_cast<cplx*>(smem_raw);  // [TB][TB+1]
  cplx* sS = sI + TB * (TB + 1);                 // [TB][TB+1]
  for (int w = threadIdx.x; w < TB * TB; w += 256) {
    const int i = w / TB, j = w % TB;
    sI[i * (TB + 1) + j] = Inv[(long)i * nP + j];
    sS[i * (TB + 1) + j] = Src[(long)i * nP + j];
  }
  __syncthreads();
  // 256 threads, each a 4x4 patch: rows ty*4.., cols tx*4..  (tx fastest)
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  cplx acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = make_double2(0.0, 0.0);
  for (int k = 0; k < TB; ++k) {
    cplx av[4], bv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) av[i] = sI[(ty * 4 + i) * (TB + 1) + k];
#pragma unroll
    for (int j = 0; j < 4; ++j) bv[j] = sS[k * (TB + 1) + tx + 16 * j];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) cfma(acc[i][j], av[i], bv[j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      Dst[(long)(ty * 4 + i) * nP + tx + 16 * j] = make_double2(-acc[i][j].x, -acc[i][j].y);
}

}  // namespace isdf

using namespace isdf;

extern "C" int isdf_pchol_workspace_bytes(int n, int batch, size_t* bytes) {
  if (!bytes || n <= 0 || batch <= 0) return ISDF_EARG;
  size_t b = (size_t)n * batch * sizeof(int);            // pos
  b = (b + 255) / 256 * 256;
  b += (size_t)batch * sizeof(PcholInfo);
  b = (b + 255) / 256 * 256;
  b += (size_t)batch * sizeof(int);                      // active
  *bytes = (b + 255) / 256 * 256;
  return ISDF_OK;
}

// a: [batch][n][n] c128 Hermitian PSD (only the LOWER triangle is read), lower triangle overwritten by the
// trailing Schur complements.
// u: [batch][ldu_rows][n] c128, rows j < rank hold row j of the factor A = U^H U in ORIGINAL column
//    order (column piv[j] carries the pivot); must be zero-initialised by the caller? -> zeroed here.
static int pchol_run(Handle* h, void* a, int n, int batch, int max_steps, double tol, int nb, void* u, int ldu_rows,
                     int* piv, int* rank, double* next_pivot, void* workspace, void* stream, bool is_real,
                     int nopivot = 0);

extern "C" int isdf_pchol(void* hv, void* a, int n, int batch, int max_steps, double tol, int nb, void* u,
                          int ldu_rows, int* piv, int* rank, double* next_pivot, void* workspace, void* stream) {
  return pchol_run((Handle*)hv, a, n, batch, max_steps, tol, nb, u, ldu_rows, piv, rank, next_pivot, workspace, stream,
                   false);
}

// Same as isdf_pchol for matrices whose imaginary parts are exactly zero (the selection matrix x4 of
// isdf_select_gram): identical arithmetic, wider panels.
extern "C" int isdf_pchol_real(void* hv, void* a, int n, int batch, int max_steps, double tol, int nb, void* u,
                               int ldu_rows, int* piv, int* rank, double* next_pivot, void* workspace, void* stream) {
  return pchol_run((Handle*)hv, a, n, batch, max_steps, tol, nb, u, ldu_rows, piv, rank, next_pivot, workspace, stream,
                   true);
}

// Unpivoted blocked Cholesky A = U^H U of [batch][n][n] Hermitian matrices (lower triangle read and destroyed): used by
// the Cholesky-QR that orthonormalises the rows of the scaled [R11 R12] factor of the gelsy fit (fftisdf.py:108).
// Stops at the first pivot <= tol (tol = 0: an exactly singular trailing block).  u [batch][ldu_rows >= n][n] upper
// triangular (rows >= rank zero), piv = identity, rank [batch].  workspace: isdf_pchol_workspace_bytes(n, batch) +
// batch * 64 * 64 * 16 bytes.
extern "C" int isdf_chol_nopivot(void* hv, void* a, int n, int batch, int max_steps, double tol, int nb, void* u,
                                 int ldu_rows, int* piv, int* rank, void* workspace, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  (void)nb;
  ISDF_CHECK_ARG(h, a && u && piv && rank && workspace, "null pointer");
  ISDF_CHECK_ARG(h, n >= 1 && batch >= 1 && batch <= 65535 && max_steps == n && ldu_rows >= n, "shape");
  char* w = (char*)workspace;
  int* pos = (int*)w;
  size_t off = ((size_t)n * batch * sizeof(int) + 255) / 256 * 256;
  PcholInfo* info = (PcholInfo*)(w + off);
  off += ((size_t)batch * sizeof(PcholInfo) + 255) / 256 * 256;
  int* active = (int*)(w + off);
  off += ((size_t)batch * sizeof(int) + 255) / 256 * 256;
  cplx* winv = (cplx*)(w + off);
  const long strideA = (long)n * n, strideU = (long)ldu_rows * n;
  ISDF_CUDA(h, cudaMemsetAsync(u, 0, (size_t)batch * ldu_rows * n * sizeof(cplx), st));
  {
    const long tot = (long)n * batch;
    pchol_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, info, active, n, batch);
    ISDF_LAUNCH_CHECK(h);
  }
  const size_t sm = 2 * CB * (CB + 1) * sizeof(cplx);
  ISDF_CUDA(h, cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  for (int j0 = 0; j0 < n; j0 += CB) {
    const int jb = (n - j0 < CB) ? (n - j0) : CB;
    chol_diag_kernel<<<batch, 256, sm, st>>>((const cplx*)a, strideA, n, j0, tol, (cplx*)u, strideU, winv, info, active);
    ISDF_LAUNCH_CHECK(h);
    const int rest = n - j0 - jb;
    if (rest <= 0) break;
    GemmParams p;
    // block row: U[j0 + r][c] = sum_k Linv[r][k] conj(A[c][j0 + k]),  c > j0 + jb   (A_j,c = conj(A_c,j): lower storage)
    p.A = winv; p.lda = CB; p.strideA = (long)CB * CB;
    p.B = (const cplx*)a + (long)(j0 + jb) * n + j0; p.ldb = n; p.strideB = strideA;
    p.C = (cplx*)u + (long)j0 * n + (j0 + jb); p.ldc = n; p.strideC = strideU;
    p.M = jb; p.N = rest; p.K = jb;
    p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
    p.perm = nullptr; p.stridePerm = 0; p.active = active; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
    ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJB, false, EPI_STORE>(p, batch, st)));
    // trailing update (lower triangle): A[c][c'] -= sum_r conj(U[r][c]) U[r][c'],  c, c' > j0 + jb
    GemmParams q;
    q.A = (const cplx*)u + (long)j0 * n + (j0 + jb); q.lda = n; q.strideA = strideU;
    q.B = q.A; q.ldb = n; q.strideB = strideU;
    q.C = (cplx*)a + (long)(j0 + jb) * n + (j0 + jb); q.ldc = n; q.strideC = strideA;
    q.M = rest; q.N = rest; q.K = jb;
    q.nseg = 1; q.segA = 0; q.segB = 0; q.alpha = 1.0;
    q.perm = nullptr; q.stridePerm = 0; q.active = active; q.ksplit = 1; q.kchunk = 0; q.strideSplit = 0;
    ISDF_CUDA(h, (launch_gemm<128, 64, true, true, MODE_CONJA, false, EPI_SUB_LOWER>(q, batch, st)));
  }
  {
    const long tot = (long)n * batch;
    pchol_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, info, n, batch, piv, rank, nullptr);
    ISDF_LAUNCH_CHECK(h);
  }
  return ISDF_OK;
}

static int pchol_run(Handle* h, void* a, int n, int batch, int max_steps, double tol, int nb, void* u, int ldu_rows,
                     int* piv, int* rank, double* next_pivot, void* workspace, void* stream, bool is_real,
                     int nopivot) {
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, a && u && piv && rank && workspace, "null pointer");
  ISDF_CHECK_ARG(h, n >= 1 && n <= PC_THREADS * PC_NCOL, "n must be in [1, 8192]");
  ISDF_CHECK_ARG(h, batch >= 1, "batch");
  ISDF_CHECK_ARG(h, max_steps >= 0 && max_steps <= n && max_steps <= ldu_rows, "max_steps");
  if (nb <= 0) nb = 32;
  ISDF_CHECK_ARG(h, nb <= PC_NB_MAX, "nb <= 64");
  char* w = (char*)workspace;
  int* pos = (int*)w;
  size_t off = ((size_t)n * batch * sizeof(int) + 255) / 256 * 256;
  PcholInfo* info = (PcholInfo*)(w + off);
  off += ((size_t)batch * sizeof(PcholInfo) + 255) / 256 * 256;
  int* active = (int*)(w + off);

  ISDF_CUDA(h, cudaMemsetAsync(u, 0, (size_t)batch * ldu_rows * n * sizeof(cplx), st));
  {
    const long tot = (long)n * batch;
    pchol_init_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, info, active, n, batch);
    ISDF_LAUNCH_CHECK(h);
  }
  const long strideA = (long)n * n, strideU = (long)ldu_rows * n;
  ISDF_CUDA(h, cudaFuncSetAttribute(pchol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    PC_THREADS * PC_NCOL * (int)sizeof(double)));
  // Large matrices: one 8-CTA cluster per matrix with the panel held in distributed shared memory.
  const bool use_cluster = (n >= 256);
  const int ncc = (n + PCC_CS - 1) / PCC_CS;
  size_t csmem = 0;
  if (use_cluster) {
    const size_t esz = is_real ? sizeof(double) : sizeof(cplx);
    const long avail = 200 * 1024 - (long)ncc * (long)sizeof(double);
    const int nb_fit = (int)(avail / ((long)ncc * (long)esz));
    ISDF_CHECK_ARG(h, nb_fit >= 4 && ncc <= PCC_THREADS * PCC_NCOL, "matrix too large for the cluster panel kernel");
    if (is_real && nb < PC_NB_MAX) nb = PC_NB_MAX;   // real panels: take the widest panel that fits
    if (nb > nb_fit) nb = nb_fit;
    csmem = (size_t)nb * ncc * esz + (size_t)ncc * sizeof(double);
    if (is_real) ISDF_CUDA(h, cudaFuncSetAttribute(pchol_panel_cluster_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
    else ISDF_CUDA(h, cudaFuncSetAttribute(pchol_panel_cluster_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem));
  }
  // The panel that reaches max_steps also evaluates the would-be next pivot and sets the stop flag.
  for (int j0 = 0; j0 == 0 || j0 < max_steps; j0 += nb) {
    if (use_cluster) {
      if (is_real)
        pchol_panel_cluster_kernel<true><<<dim3(PCC_CS, batch), PCC_THREADS, csmem, st>>>(
            (const cplx*)a, n, strideA, n, ncc, j0, nb, max_steps, tol, (cplx*)u, n, strideU, pos, info, active,
            nopivot);
      else
        pchol_panel_cluster_kernel<false><<<dim3(PCC_CS, batch), PCC_THREADS, csmem, st>>>(
            (const cplx*)a, n, strideA, n, ncc, j0, nb, max_steps, tol, (cplx*)u, n, strideU, pos, info, active,
            nopivot);
    } else {
      pchol_panel_kernel<<<batch, PC_THREADS, (size_t)n * sizeof(double), st>>>(
          (const cplx*)a, n, strideA, n, j0, nb, max_steps, tol, (cplx*)u, n, strideU, pos, info, active, nopivot);
    }
    ISDF_LAUNCH_CHECK(h);
    if (j0 + nb < max_steps) {
      // trailing update A -= U_p^H U_p (only needed while further pivot rows will be read)
      GemmParams p;
      p.A = (const cplx*)u + (long)j0 * n; p.lda = n; p.strideA = strideU;
      p.B = p.A; p.ldb = n; p.strideB = strideU;
      p.C = (cplx*)a; p.ldc = n; p.strideC = strideA;
      p.M = n; p.N = n; p.K = nb;
      p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
      p.perm = nullptr; p.stridePerm = 0; p.active = active; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
      ISDF_CUDA(h, (launch_gemm<128, 64, true, true, MODE_CONJA, false, EPI_SUB_LOWER>(p, batch, st)));
    }
  }
  {
    const long tot = (long)n * batch;
    pchol_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, info, n, batch, piv, rank, next_pivot);
    ISDF_LAUNCH_CHECK(h);
  }
  return ISDF_OK;
}

// From the pivoted factor build the block operators of the two triangular sweeps (block size 64):
//   forward  T_a <- sum_{b<=a} Lfwd[a][b] T_b,   backward T_a <- sum_{b>=a} Ubwd[a][b] T_b.
// work: 2 * batch * nP * nP c128 (Up, Lp).  lfwd, ubwd: batch * nP * nP c128 each.
extern "C" int isdf_trsm_prepare(void* hv, const void* u, int ldu_rows, const int* piv, const int* rank, int n, int nP,
                                 int batch, void* lfwd, void* ubwd, void* work, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, u && piv && rank && lfwd && ubwd && work, "null pointer");
  // nP may be smaller than n when every rank[b] <= nP (rows beyond the largest rank are never used)
  ISDF_CHECK_ARG(h, nP % TB == 0 && nP >= TB && n >= 1, "nP must be a positive multiple of 64");
  ISDF_CHECK_ARG(h, batch >= 1 && batch <= 65535, "batch");
  cplx* Up = (cplx*)work;
  cplx* Lp = Up + (long)batch * nP * nP;
  {
    dim3 grid((nP + 31) / 32, (nP + 31) / 32, batch), block(32, 8);
    build_up_lp_kernel<<<grid, block, 0, st>>>((const cplx*)u, n, (long)ldu_rows * n, piv, rank, n, nP, Up, Lp);
    ISDF_LAUNCH_CHECK(h);
  }
  const int nblk = nP / TB;
  const size_t sm = 2 * TB * (TB + 1) * sizeof(cplx);
  ISDF_CUDA(h, cudaFuncSetAttribute(tri_inv_blocks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  ISDF_CUDA(h, cudaFuncSetAttribute(apply_diag_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  {
    dim3 grid(nblk, batch);
    tri_inv_blocks_kernel<<<grid, TB, sm, st>>>(Up, nP, (cplx*)ubwd, (cplx*)lfwd);
    ISDF_LAUNCH_CHECK(h);
  }
  if (nblk > 1) {
    dim3 grid(nblk, nblk, batch);
    apply_diag_inv_kernel<<<grid, 256, sm, st>>>(Up, Lp, nP, (cplx*)ubwd, (cplx*)lfwd);
    ISDF_LAUNCH_CHECK(h);
  }
  return ISDF_OK;
}

// One block row of a sweep: 64 x 32 tiles for full 64-row blocks, the transposed 32 x 64 tile for a ragged tail
// of <= 32 live rows (every warp of the CTA then has live rows).
static cudaError_t sweep_launch(const GemmParams& p, int batch, cudaStream_t st) {
  // the sweeps update T in place (C rows alias B rows): race-free only while ONE M-tile covers the whole 64-row
  // block (32 rows for the ragged tail, which takes the transposed tile)
  static_assert(!gemm_is_3m(false) || (ISDF_GEMM_3M_BM >= TB && ISDF_GEMM_3M_BN >= 32), "sweep tile must cover a block row");
  static_assert(gemm_is_3m(false) || ISDF_GEMM_SMALL, "in-place sweeps need the 64-row tile");
  if (p.M <= 32) return launch_gemm<64, 128, false, true, MODE_AB, false, EPI_STORE>(p, batch, st);
  return launch_gemm<128, 64, false, true, MODE_AB, false, EPI_STORE>(p, batch, st);
}

// One direction only: backward = 0: T <- U^{-H} T with op = lfwd;  backward = 1: T <- U^{-1} T with op = ubwd.
extern "C" int isdf_trsm_sweep(void* hv, const void* op, void* t, int nP, int nact, long ng, long ldt, int batch,
                               int backward, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, op && t, "null pointer");
  ISDF_CHECK_ARG(h, nP % TB == 0 && ng >= 1 && ng < (1L << 31) && ldt >= ng, "shape");
  ISDF_CHECK_ARG(h, nact >= 0 && nact <= nP, "nact out of range");
  if (nact == 0) return ISDF_OK;
  const int nblk = (nact + TB - 1) / TB;
  GemmParams p;
  p.lda = nP; p.strideA = (long)nP * nP;
  p.ldb = ldt; p.strideB = (long)nP * ldt;
  p.ldc = ldt; p.strideC = (long)nP * ldt;
  p.N = (int)ng;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  if (!backward) {
    for (int a = 0; a < nblk; ++a) {
      p.A = (const cplx*)op + (long)a * TB * nP;
      p.B = (const cplx*)t;
      p.C = (cplx*)t + (long)a * TB * ldt;
      p.M = (nact - a * TB < TB) ? (nact - a * TB) : TB;
      p.K = ((a + 1) * TB < nact) ? (a + 1) * TB : nact;
      ISDF_CUDA(h, sweep_launch(p, batch, st));
    }
  } else {
    for (int a = nblk - 1; a >= 0; --a) {
      p.A = (const cplx*)op + (long)a * TB * nP + (long)a * TB;
      p.B = (const cplx*)t + (long)a * TB * ldt;
      p.C = (cplx*)t + (long)a * TB * ldt;
      p.M = (nact - a * TB < TB) ? (nact - a * TB) : TB;
      p.K = nact - a * TB;
      ISDF_CUDA(h, sweep_launch(p, batch, st));
    }
  }
  return ISDF_OK;
}

// In-place blocked substitution  T <- U^{-1} U^{-H} T  on T[batch][nP][ng] (row-major, ng contiguous).
// Only the first nact rows (nact >= every batch member's rank) are touched: rows at and beyond the rank are
// zero on input and stay zero, so neither their block rows nor their K range are executed.
extern "C" int isdf_trsm_sweeps(void* hv, const void* lfwd, const void* ubwd, void* t, int nP, int nact, long ng,
                                long ldt, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, lfwd && ubwd && t, "null pointer");
  ISDF_CHECK_ARG(h, nP % TB == 0 && ng >= 1 && ng < (1L << 31) && ldt >= ng, "shape");
  ISDF_CHECK_ARG(h, nact >= 0 && nact <= nP, "nact out of range");
  if (nact == 0) return ISDF_OK;
  const int nblk = (nact + TB - 1) / TB;
  GemmParams p;
  p.lda = nP; p.strideA = (long)nP * nP;
  p.ldb = ldt; p.strideB = (long)nP * ldt;
  p.ldc = ldt; p.strideC = (long)nP * ldt;
  p.N = (int)ng;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  for (int a = 0; a < nblk; ++a) {  // forward: rows 0..a
    p.A = (const cplx*)lfwd + (long)a * TB * nP;
    p.B = (const cplx*)t;
    p.C = (cplx*)t + (long)a * TB * ldt;
    p.M = (nact - a * TB < TB) ? (nact - a * TB) : TB;
    p.K = ((a + 1) * TB < nact) ? (a + 1) * TB : nact;
    ISDF_CUDA(h, sweep_launch(p, batch, st));
  }
  for (int a = nblk - 1; a >= 0; --a) {  // backward: rows a..end
    p.A = (const cplx*)ubwd + (long)a * TB * nP + (long)a * TB;
    p.B = (const cplx*)t + (long)a * TB * ldt;
    p.C = (cplx*)t + (long)a * TB * ldt;
    p.M = (nact - a * TB < TB) ? (nact - a * TB) : TB;
    p.K = nact - a * TB;
    ISDF_CUDA(h, sweep_launch(p, batch, st));
  }
  return ISDF_OK;
}
