// Rank-revealing QR of the per-q metric A_q with LAPACK ZGELSY semantics -- the factorisation inside
// scipy.linalg.lstsq(x4_q, y_q.T, lapack_driver="gelsy") at /root/reference/fftisdf.py:108:
//
//   zgeqp3 : Householder QR with column pivoting on downdated partial column norms (blocked zlaqps recurrence,
//            sqrt(eps) recomputation safeguard)                                     -> qrcp_cluster_kernel
//   zgelsy : rank = leading columns of R accepted by incremental condition estimation (zlaic1) with
//            rcond = eps (scipy's default cond)                                     -> gelsy_rank_kernel
//   ztzrzf / zunmqr / ztrsm / zunmrz: applied to the right-hand sides as three dense operators built once per q
//            (Q1 D^-1, the triangular factor of the row-scaled [R11 R12], E = P Z1^H); the small kernels at the end
//            of this file extract their ingredients, the products run on the DMMA GEMM engine (kernels.py).
//
// Layout: the working matrix is held COLUMN-major, w[c][i] = A[i][c], so that a column (pivot column, GEMV operand)
// is contiguous.  One thread-block cluster per matrix; CTA r owns the columns [r*ncc, (r+1)*ncc) and only ever
// touches those.  Columns are not swapped: pos[c] is the pivot position of column c (LAPACK's tie rule -- first
// maximum in current position order -- is kept by tracking positions exactly as the swaps would move them).
//
// Per pivot step k (panel-relative t):
//   S1  every CTA publishes its best candidate (value, position, index, recompute flag) into every CTA's shared
//       memory (DSMEM) -- cluster barrier -- all reduce the candidates identically;
//       the owner of the pivot column p applies the pending panel reflectors to it
//       (a_p -= V conj(F[p,:])), generates the reflector (zlarfg), writes v_k / tau_k / R_kk, and publishes
//       auxv = -tau V^H v_k and the pivot row of V;
//   S2  cluster barrier -- every CTA computes, for its live columns, F[c,t] = tau a_c^H v_k + F[c,:t] auxv
//       (warp per column, coalesced), the pivot-row update R[k,c] and the norm downdate.
// A panel ends after nb steps; the owner CTAs then apply the deferred rank-nb update to their columns.  (zlaqps also
// ends it when a column asks for its norm to be recomputed; here that column is brought up to date on the fly.)
#include <float.h>
#include <stdlib.h>
#include <cooperative_groups.h>
#include "gemm_c128.cuh"

namespace cg = cooperative_groups;

namespace isdf {

constexpr int QR_THREADS = 512;
constexpr int QR_NW = QR_THREADS / 32;
constexpr int QR_NB_MAX = 32;
constexpr int QR_CS_MAX = 16;
constexpr int QR_NSL = 16;             // logical row slices of the pivot column (independent of the cluster size)
constexpr int QR_NCOLT = 2;            // columns per thread in the per-column phases -> ncc <= 1024
constexpr int QR_RT_MIN = 64;          // rows per trailing-update tile (lower bound)

struct QrCand { double v; int pos; int idx; int flag; };
struct QrBcast { cplx aux[QR_NB_MAX]; cplx vrow[QR_NB_MAX]; cplx tau; };
struct QrPart { cplx dots[QR_NB_MAX]; cplx ak; double nrm2; double pad; };

__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// Fire-and-forget L2 prefetch of w[i0..n) by one warp (128-byte lines = 8 complex): turns the DRAM latency of the
// demand loads that follow into L2 latency without holding registers for the bytes in flight.
__device__ __forceinline__ void warp_prefetch_l2(const cplx* w, int i0, int n, int lane) {
  for (int i = (i0 & ~7) + lane * 8; i < n; i += 256)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(w + i));
}

// 2-norm of w[i0..n) by one warp
__device__ __forceinline__ double warp_col_norm(const cplx* __restrict__ w, int i0, int n, int lane) {
  double s0 = 0.0, s1 = 0.0;
  int i = i0 + lane;
  for (; i + 32 < n; i += 64) {
    const cplx a = w[i], b = w[i + 32];
    s0 += a.x * a.x + a.y * a.y;
    s1 += b.x * b.x + b.y * b.y;
  }
  if (i < n) { const cplx a = w[i]; s0 += a.x * a.x + a.y * a.y; }
  return sqrt(warp_sum(s0 + s1));
}

__global__ void __launch_bounds__(QR_THREADS, 1)
qrcp_cluster_kernel(cplx* __restrict__ Wall, long strideW, int n, int ncc, int nb, int nvb, cplx* __restrict__ Vall,
                    long strideV, cplx* __restrict__ tauall, int* __restrict__ posall) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int CS = (int)cluster.num_blocks();
  const int b = blockIdx.y;
  cplx* W = Wall + (long)b * strideW;
  cplx* V = Vall + (long)b * strideV;          // V[k][i] = v_k[i]  (k-major), zero for i < k
  cplx* tau_out = tauall + (long)b * n;

  extern __shared__ __align__(16) unsigned char qr_smem[];
  cplx* vbuf = reinterpret_cast<cplx*>(qr_smem);                 // [nvb] reflector / column / V tile
  cplx* Ft = vbuf + nvb;                                          // [nb][ncc]
  cplx* sdot = Ft + (long)nb * ncc;                               // [ncc]
  double* vn1 = reinterpret_cast<double*>(sdot + ncc);            // [ncc]
  double* vn2 = vn1 + ncc;                                        // [ncc]
  int* pos = reinterpret_cast<int*>(vn2 + ncc);                   // [ncc]
  int* mark = pos + ncc;                                          // [ncc]
  __shared__ QrCand cand[2][QR_CS_MAX];
  __shared__ QrBcast bc[1];
  __shared__ QrPart part[2][QR_NSL];
  __shared__ cplx fp[QR_NB_MAX];
  __shared__ cplx mydots[QR_NSL * QR_NB_MAX];
  __shared__ cplx tdots[QR_NB_MAX];
  __shared__ double s_nrm[QR_NSL];
  __shared__ double red_v[QR_NW];
  __shared__ int red_pos[QR_NW], red_idx[QR_NW], red_flag[QR_NW];
  __shared__ double s_scal[6];     // beta, tau.re, tau.im, scale.re, scale.im, spare

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c_lo = crank * ncc;
  const int nown = max(0, min(ncc, n - c_lo));                    // columns this CTA owns

  for (int lc = tid; lc < ncc; lc += QR_THREADS) { pos[lc] = (lc < nown) ? c_lo + lc : -1; mark[lc] = 0; }
  for (int lc = warp; lc < nown; lc += QR_NW) {
    const double nr = warp_col_norm(W + (long)(c_lo + lc) * n, 0, n, lane);
    if (lane == 0) { vn1[lc] = nr; vn2[lc] = nr; }
  }
  __syncthreads();

  const double tol3z = sqrt(DBL_EPSILON * 0.5);
  int it = 0;          // exchange counter (buffer parity)
  int k = 0;
  while (k < n) {
    const int j0 = k;
    int t = 0;
    double dp_start = 0.0;
    while (t < nb && k < n) {
      const int par = it & 1;
      ++it;
      // ---- S1: local candidate = first maximum of vn1 in position order, plus the recompute flag
      double bv = -1.0;
      int bpos = 0x7fffffff, bidx = -1, bflag = 0;
#pragma unroll
      for (int j = 0; j < QR_NCOLT; ++j) {
        const int lc = tid + j * QR_THREADS;
        if (lc < nown) {
          bflag |= mark[lc];
          const int ps = pos[lc];
          if (ps >= k) {
            const double d = vn1[lc];
            if (d > bv || (d == bv && ps < bpos)) { bv = d; bpos = ps; bidx = c_lo + lc; }
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        bflag |= __shfl_xor_sync(0xffffffffu, bflag, o);
        if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
      }
      if (lane == 0) { red_v[warp] = bv; red_pos[warp] = bpos; red_idx[warp] = bidx; red_flag[warp] = bflag; }
      __syncthreads();
      if (warp == 0) {
        if (lane < QR_NW) { bv = red_v[lane]; bpos = red_pos[lane]; bidx = red_idx[lane]; bflag = red_flag[lane]; }
        else { bv = -1.0; bpos = 0x7fffffff; bidx = -1; bflag = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
          const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
          const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
          bflag |= __shfl_xor_sync(0xffffffffu, bflag, o);
          if (ov > bv || (ov == bv && op < bpos)) { bv = ov; bpos = op; bidx = oi; }
        }
        if (lane < CS) {
          QrCand (*remote)[QR_CS_MAX] = cluster.map_shared_rank(cand, lane);
          remote[par][crank].v = bv; remote[par][crank].pos = bpos; remote[par][crank].idx = bidx;
          remote[par][crank].flag = bflag;
        }
      }
      cluster.sync();
      double dp = -1.0;
      int ppos = 0x7fffffff, p = -1, anyflag = 0;
      for (int r = 0; r < CS; ++r) {
        const double ov = cand[par][r].v;
        const int op = cand[par][r].pos, oi = cand[par][r].idx;
        anyflag |= cand[par][r].flag;
        if (oi >= 0 && (ov > dp || (ov == dp && op < ppos))) { dp = ov; ppos = op; p = oi; }
      }
      (void)anyflag;
      // Graded accuracy: the deferred update forms trailing entries as differences of panel-start magnitudes, so a panel
      // must not span a large decay of the pivot norms (LAPACK gets this for free: zlaqps closes the panel at every
      // norm recomputation, and small matrices take the unblocked zlaqp2).  Close it once the pivot norm has dropped
      // below 1/16 of the panel's first one: steep spectra (small metrics: 17 decades in 26 steps) degenerate to the
      // unblocked recurrence, flat ones (3120 x 3120: a factor 1.5 per 32 steps) keep full panels.
      if (t == 0) dp_start = dp;
      else if (dp < 0.0625 * dp_start) break;
      // ---- positions: column p takes position k, the column that sat at k moves to p's old position
#pragma unroll
      for (int j = 0; j < QR_NCOLT; ++j) {
        const int lc = tid + j * QR_THREADS;
        if (lc < nown) {
          if (c_lo + lc == p) pos[lc] = k;
          else if (pos[lc] == k) pos[lc] = ppos;
        }
      }
      const int owner = p / ncc;
      const int lp = p - owner * ncc;
      // ---- the pivot column, by ALL CTAs of the cluster: CTA r takes the r-th slice of the rows [k, n)
      //      (one row per thread).  fp = F[p][0..t) comes from the owner's shared memory (DSMEM read).
      if (tid < t) {
        const cplx* Fo = cluster.map_shared_rank(Ft, owner);
        fp[tid] = Fo[(long)tid * ncc + lp];
      }
      __syncthreads();
      // The rows [k, n) are cut into QR_NSL logical slices whatever the cluster size, every slice is reduced in a fixed
      // order, and the slice partials are summed in slice order: the factorisation is bit-identical for every
      // cluster size (the cluster size follows the number of matrices per GPU, i.e. the number of GPUs).
      const int spc = QR_NSL / CS;                       // logical slices per CTA
      const int sl = (n - k + QR_NSL - 1) / QR_NSL;      // rows per logical slice
      for (int s_loc = 0; s_loc < spc; ++s_loc) {
        const int i_lo = k + (crank * spc + s_loc) * sl;
        const int cnt = max(0, min(sl, n - i_lo));
        cplx* abuf = vbuf + (long)s_loc * sl;
        double ss = 0.0;
        for (int r = tid; r < cnt; r += QR_THREADS) {
          const int my_i = i_lo + r;
          cplx a = __ldcg(W + (long)p * n + my_i);
          // a_p -= sum_tt v_tt conj(F[p][tt]); loads in batches of 8 (the column of V is L2-resident, latency-bound)
          for (int t0 = 0; t0 < t; t0 += 8) {
            cplx vv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
              vv[u] = (t0 + u < t) ? __ldcg(V + (long)(j0 + t0 + u) * n + my_i) : make_double2(0.0, 0.0);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              if (t0 + u < t) {
                const cplx f = fp[t0 + u];
                a.x -= vv[u].x * f.x + vv[u].y * f.y;      // vv * conj(f)
                a.y -= vv[u].y * f.x - vv[u].x * f.y;
              }
            }
          }
          abuf[r] = a;                                    // slice buffer (the previous reflector is dead by now)
          if (my_i > k) ss += a.x * a.x + a.y * a.y;
        }
        ss = warp_sum(ss);
        if (lane == 0) red_v[warp] = ss;
        __syncthreads();
        if (tid == 0) {
          double tot = 0.0;
          for (int w = 0; w < QR_NW; ++w) tot += red_v[w];
          s_nrm[s_loc] = tot;
        }
        // partial dots[tt] = sum_{i in slice, i > k} conj(v_tt[i]) a[i]   (warp per tt)
        for (int tt = warp; tt < t; tt += QR_NW) {
          const cplx* Vt = V + (long)(j0 + tt) * n + i_lo;
          double sr = 0.0, si = 0.0;
          for (int r = lane; r < cnt; r += 32) {
            if (i_lo + r > k) {
              const cplx vv = __ldcg(Vt + r), av = abuf[r];
              sr += vv.x * av.x + vv.y * av.y;        // conj(vv) * av
              si += vv.x * av.y - vv.y * av.x;
            }
          }
          sr = warp_sum(sr); si = warp_sum(si);
          if (lane == 0) mydots[s_loc * QR_NB_MAX + tt] = make_double2(sr, si);
        }
        __syncthreads();
      }
      // publish {dots[0..t), a_k, |a|^2} of every local slice into every CTA's table
      {
        const int nval = t + 2;
        for (int e = tid; e < nval * spc * CS; e += QR_THREADS) {
          const int dst = e / (nval * spc);
          const int rem = e - dst * (nval * spc);
          const int s_loc = rem / nval, idx = rem - s_loc * nval;
          const int sid = crank * spc + s_loc;
          QrPart (*remote)[QR_NSL] = cluster.map_shared_rank(part, dst);
          if (idx < t) remote[par][sid].dots[idx] = mydots[s_loc * QR_NB_MAX + idx];
          else if (idx == t) remote[par][sid].ak = (sid == 0) ? vbuf[0] : make_double2(0.0, 0.0);
          else remote[par][sid].nrm2 = s_nrm[s_loc];
        }
      }
      cluster.sync();
      // ---- every CTA: totals, zlarfg, auxv; scales and stores its slice of v_k
      if (tid < t) {
        double sr = 0.0, si = 0.0;
        for (int r = 0; r < QR_NSL; ++r) { sr += part[par][r].dots[tid].x; si += part[par][r].dots[tid].y; }
        tdots[tid] = make_double2(sr, si);
      }
      if (tid == 32) {
        double tot = 0.0;
        for (int r = 0; r < QR_NSL; ++r) tot += part[par][r].nrm2;
        const cplx alpha = part[par][0].ak;
        double beta, tr, ti, sr, si;
        if (tot == 0.0 && alpha.y == 0.0) {
          beta = alpha.x; tr = 0.0; ti = 0.0; sr = 0.0; si = 0.0;
        } else {
          beta = -copysign(sqrt(alpha.x * alpha.x + alpha.y * alpha.y + tot), alpha.x);
          tr = (beta - alpha.x) / beta; ti = -alpha.y / beta;
          const double dr = alpha.x - beta, di = alpha.y, den = dr * dr + di * di;   // 1 / (alpha - beta)
          sr = dr / den; si = -di / den;
        }
        s_scal[0] = beta; s_scal[1] = tr; s_scal[2] = ti; s_scal[3] = sr; s_scal[4] = si;
      }
      __syncthreads();
      const double beta = s_scal[0];
      const cplx tau = make_double2(s_scal[1], s_scal[2]);
      const cplx scal = make_double2(s_scal[3], s_scal[4]);
      const bool ident = (tau.x == 0.0 && tau.y == 0.0);       // H = I: represented by v = 0
      if (tid < t) {
        // auxv[tt] = -tau (conj(v_tt[k]) + scale * dots[tt]);  vrow[tt] = v_tt[k]
        const cplx vr = __ldcg(V + (long)(j0 + tid) * n + k);
        const cplx sd = cmul(scal, tdots[tid]);
        const cplx in = ident ? make_double2(0.0, 0.0) : make_double2(vr.x + sd.x, -vr.y + sd.y);
        bc[0].aux[tid] = make_double2(-(tau.x * in.x - tau.y * in.y), -(tau.x * in.y + tau.y * in.x));
        bc[0].vrow[tid] = vr;
      }
      for (int s_loc = 0; s_loc < spc; ++s_loc) {
        const int i_lo = k + (crank * spc + s_loc) * sl;
        const int cnt = max(0, min(sl, n - i_lo));
        const cplx* abuf = vbuf + (long)s_loc * sl;
        for (int r = tid; r < cnt; r += QR_THREADS) {
          const int my_i = i_lo + r;
          cplx v;
          if (ident) v = make_double2(0.0, 0.0);
          else if (my_i == k) v = make_double2(1.0, 0.0);
          else v = cmul(abuf[r], scal);
          V[(long)k * n + my_i] = v;
        }
      }
      __syncthreads();    // the slice in vbuf is consumed before the full reflector overwrites it
      if (crank == owner && tid == 0) { W[(long)p * n + k] = make_double2(beta, 0.0); tau_out[k] = tau; }
      cluster.sync();
      // ---- every CTA holds tau, auxv, vrow; the full reflector comes back from global memory (L2)
      for (int i = k + tid; i < n; i += QR_THREADS) vbuf[i] = __ldcg(V + (long)k * n + i);
      __syncthreads();
      constexpr int par0 = 0;
      // GEMV, warp per live column (two columns at a time, four rows per lane in flight -> 8 independent 16-byte loads
      // per lane: the stage is latency-bound otherwise):  sdot[c] = sum_{i>=k} conj(a_c[i]) v_k[i]
      for (int lc = warp; lc < nown; lc += 2 * QR_NW) {
        const int lc2 = lc + QR_NW;
        const bool live0 = pos[lc] > k;
        const bool live1 = (lc2 < nown) && pos[lc2] > k;
        if (!live0 && !live1) continue;
        const cplx* W0 = W + (long)(c_lo + (live0 ? lc : lc2)) * n;
        const cplx* W1 = W + (long)(c_lo + (live1 ? lc2 : lc)) * n;
        double ar = 0.0, ai = 0.0, br = 0.0, bi = 0.0;
        int i = k + lane;
        for (; i + 224 < n; i += 256) {      // 16 independent 16-byte loads per lane in flight
          cplx a[8], c[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) { a[u] = W0[i + 32 * u]; c[u] = W1[i + 32 * u]; }
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const cplx v = vbuf[i + 32 * u];
            ar += a[u].x * v.x + a[u].y * v.y;  ai += a[u].x * v.y - a[u].y * v.x;
            br += c[u].x * v.x + c[u].y * v.y;  bi += c[u].x * v.y - c[u].y * v.x;
          }
        }
        for (; i + 96 < n; i += 128) {
          cplx a[4], c[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { a[u] = W0[i + 32 * u]; c[u] = W1[i + 32 * u]; }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const cplx v = vbuf[i + 32 * u];
            ar += a[u].x * v.x + a[u].y * v.y;  ai += a[u].x * v.y - a[u].y * v.x;
            br += c[u].x * v.x + c[u].y * v.y;  bi += c[u].x * v.y - c[u].y * v.x;
          }
        }
        for (; i < n; i += 32) {
          const cplx a = W0[i], c = W1[i], v = vbuf[i];
          ar += a.x * v.x + a.y * v.y;  ai += a.x * v.y - a.y * v.x;
          br += c.x * v.x + c.y * v.y;  bi += c.x * v.y - c.y * v.x;
        }
        ar = warp_sum(ar); ai = warp_sum(ai); br = warp_sum(br); bi = warp_sum(bi);
        if (lane == 0) {
          if (live0) sdot[lc] = make_double2(ar, ai);
          if (live1) sdot[lc2] = live0 ? make_double2(br, bi) : make_double2(ar, ai);
        }
      }
      __syncthreads();
      // per live column: F[c][t], pivot-row entry R[k][c], norm downdate
#pragma unroll
      for (int j = 0; j < QR_NCOLT; ++j) {
        const int lc = tid + j * QR_THREADS;
        if (lc < nown && pos[lc] > k) {
          const cplx s = sdot[lc];
          cplx f = cmul(tau, s);
          cplx* Wc = W + (long)(c_lo + lc) * n;
          cplx rk = Wc[k];
          for (int tt = 0; tt < t; ++tt) {
            const cplx ft = Ft[(long)tt * ncc + lc];
            cfma(f, ft, bc[par0].aux[tt]);
            const cplx vr = bc[par0].vrow[tt];
            rk.x -= vr.x * ft.x + vr.y * ft.y;     // vrow * conj(F)
            rk.y -= vr.y * ft.x - vr.x * ft.y;
          }
          Ft[(long)t * ncc + lc] = f;
          rk.x -= f.x; rk.y += f.y;                // vrow[t] = 1:  rk -= conj(f)
          Wc[k] = rk;
          const double v1 = vn1[lc];
          if (v1 != 0.0) {
            double temp = sqrt(rk.x * rk.x + rk.y * rk.y) / v1;
            temp = fmax(0.0, (1.0 + temp) * (1.0 - temp));
            const double q = v1 / vn2[lc];
            const double temp2 = temp * q * q;
            if (temp2 <= tol3z) mark[lc] = 1;
            else vn1[lc] = v1 * sqrt(temp);
          }
        }
      }
      __syncthreads();
      // Flagged columns get their partial norm recomputed exactly (dznrm2 of the rows below the pivot row).  zlaqps
      // closes the panel for that, because the column must be up to date first; on graded matrices (the metric spans
      // 15 decades) some column asks for it at almost every step, which would degrade the panel to width 1 and triple
      // the memory traffic.  Here the pending reflectors of the panel are applied to the flagged column on the fly
      // (same values, the panel stays open): u = a_c - sum_{tt<=t} v_tt conj(F[c][tt]), rows > k.
      // (the whole CTA works on one flagged column at a time: a single warp would need ~50 us per column, and every
      //  other CTA of the cluster waits for it at the next barrier)
      for (int lc = 0; lc < nown; ++lc) {
        if (!mark[lc]) continue;                      // block-uniform (shared memory, synchronised above)
        const cplx* Wc = W + (long)(c_lo + lc) * n;
        double ssq = 0.0;
        for (int i = k + 1 + tid; i < n; i += QR_THREADS) {
          cplx u = Wc[i];
          for (int t0 = 0; t0 <= t; t0 += 8) {
            cplx vv[8];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              vv[q] = (t0 + q <= t) ? __ldcg(V + (long)(j0 + t0 + q) * n + i) : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              if (t0 + q <= t) {
                const cplx f = Ft[(long)(t0 + q) * ncc + lc];
                u.x -= vv[q].x * f.x + vv[q].y * f.y;
                u.y -= vv[q].y * f.x - vv[q].x * f.y;
              }
            }
          }
          ssq += u.x * u.x + u.y * u.y;
        }
        ssq = warp_sum(ssq);
        if (lane == 0) red_v[warp] = ssq;
        __syncthreads();
        if (tid == 0) {
          double tot = 0.0;
          for (int w = 0; w < QR_NW; ++w) tot += red_v[w];
          const double nr = sqrt(tot);
          vn1[lc] = nr; vn2[lc] = nr; mark[lc] = 0;
        }
        __syncthreads();
      }
      __syncthreads();
      ++t; ++k;
    }
    // ---- close the panel: deferred update of the live columns, rows >= k:  a_c -= sum_tt v_tt conj(F[c][tt])
    if (t > 0 && k < n) {
      const int rt = max(32, min(256, (nvb / t) & ~31));
      for (int r0 = k; r0 < n; r0 += rt) {
        const int rows = min(rt, n - r0);
        __syncthreads();
        for (int e = tid; e < t * rows; e += QR_THREADS) {
          const int tt = e / rows, ii = e - tt * rows;
          vbuf[tt * rt + ii] = __ldcg(V + (long)(j0 + tt) * n + r0 + ii);
        }
        __syncthreads();
        // 2 columns x 2 rows per lane: each staged V element and each F element feeds two complex MACs
        for (int lc = warp; lc < nown; lc += 2 * QR_NW) {
          const int lc2 = lc + QR_NW;
          const bool live0 = pos[lc] >= k;
          const bool live1 = (lc2 < nown) && pos[lc2] >= k;
          if (!live0 && !live1) continue;
          const int la = live0 ? lc : lc2, lb = live1 ? lc2 : lc;
          cplx* Wa = W + (long)(c_lo + la) * n + r0;
          cplx* Wb = W + (long)(c_lo + lb) * n + r0;
          for (int ii = lane; ii < rows; ii += 64) {
            const int i2 = ii + 32;
            const bool two = i2 < rows;
            cplx a0 = make_double2(0.0, 0.0), a1 = a0, b0 = a0, b1 = a0;
            // issue the global loads first: their latency overlaps the t-loop
            cplx w0 = Wa[ii], u0 = Wb[ii];
            cplx w1 = two ? Wa[i2] : a0, u1 = two ? Wb[i2] : a0;
            for (int tt = 0; tt < t; ++tt) {
              const cplx v0 = vbuf[tt * rt + ii];
              const cplx v1 = two ? vbuf[tt * rt + i2] : make_double2(0.0, 0.0);
              const cplx fa = Ft[(long)tt * ncc + la];
              const cplx fb = Ft[(long)tt * ncc + lb];
              a0.x += v0.x * fa.x + v0.y * fa.y;  a0.y += v0.y * fa.x - v0.x * fa.y;   // v conj(f)
              a1.x += v1.x * fa.x + v1.y * fa.y;  a1.y += v1.y * fa.x - v1.x * fa.y;
              b0.x += v0.x * fb.x + v0.y * fb.y;  b0.y += v0.y * fb.x - v0.x * fb.y;
              b1.x += v1.x * fb.x + v1.y * fb.y;  b1.y += v1.y * fb.x - v1.x * fb.y;
            }
            w0.x -= a0.x; w0.y -= a0.y;
            Wa[ii] = w0;
            if (two) { w1.x -= a1.x; w1.y -= a1.y; Wa[i2] = w1; }
            if (live0 && live1) {
              u0.x -= b0.x; u0.y -= b0.y;
              Wb[ii] = u0;
              if (two) { u1.x -= b1.x; u1.y -= b1.y; Wb[i2] = u1; }
            }
          }
        }
      }
      __syncthreads();
    }
    // recompute the flagged norms (dznrm2 of the rows below the last pivot row)
    for (int lc = warp; lc < nown; lc += QR_NW) {
      if (mark[lc]) {    // warp-uniform
        double nr = 0.0;
        if (pos[lc] >= k && k < n) nr = warp_col_norm(W + (long)(c_lo + lc) * n, k, n, lane);
        __syncwarp();
        if (lane == 0) { vn1[lc] = nr; vn2[lc] = nr; mark[lc] = 0; }
      }
    }
    __syncthreads();
  }
  for (int lc = tid; lc < nown; lc += QR_THREADS) posall[(long)b * n + c_lo + lc] = pos[lc];
  cluster.sync();      // keep every CTA's shared memory alive until all remote stores into it have landed
}

__global__ void qrcp_finalize_kernel(const int* __restrict__ pos, int n, int batch, int* __restrict__ piv) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (long)n * batch) {
    const long b = i / n;
    piv[b * n + pos[i]] = (int)(i % n);
  }
}

// ---------------------------------------------------------------------------------------------
// zlaic1 (one step of incremental condition estimation), jobs 1 (largest) and 2 (smallest).
struct Laic1Out { double sestpr; cplx s, c; };

__device__ __forceinline__ cplx cdivr(cplx a, double d) { return make_double2(a.x / d, a.y / d); }
__device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }

__device__ Laic1Out laic1(int job, double sest, cplx alpha, cplx gamma) {
  const double eps = DBL_EPSILON * 0.5;
  const double absalp = sqrt(cabs2(alpha)), absgam = sqrt(cabs2(gamma)), absest = fabs(sest);
  Laic1Out o;
  const cplx one = make_double2(1.0, 0.0), zero = make_double2(0.0, 0.0);
  if (job == 1) {
    if (sest == 0.0) {
      const double s1 = fmax(absgam, absalp);
      if (s1 == 0.0) { o.s = zero; o.c = one; o.sestpr = 0.0; return o; }
      cplx s = cdivr(alpha, s1), c = cdivr(gamma, s1);
      const double tmp = sqrt(cabs2(s) + cabs2(c));
      o.s = cdivr(s, tmp); o.c = cdivr(c, tmp); o.sestpr = s1 * tmp; return o;
    }
    if (absgam <= eps * absest) {
      const double tmp = fmax(absest, absalp);
      const double s1 = absest / tmp, s2 = absalp / tmp;
      o.s = one; o.c = zero; o.sestpr = tmp * sqrt(s1 * s1 + s2 * s2); return o;
    }
    if (absalp <= eps * absest) {
      const double s1 = absgam, s2 = absest;
      if (s1 <= s2) { o.s = one; o.c = zero; o.sestpr = s2; }
      else { o.s = zero; o.c = one; o.sestpr = s1; }
      return o;
    }
    if (absest <= eps * absalp || absest <= eps * absgam) {
      const double s1 = absgam, s2 = absalp;
      if (s1 <= s2) {
        const double tmp = s1 / s2, scl = sqrt(1.0 + tmp * tmp);
        o.sestpr = s2 * scl; o.s = cdivr(cdivr(alpha, s2), scl); o.c = cdivr(cdivr(gamma, s2), scl);
      } else {
        const double tmp = s2 / s1, scl = sqrt(1.0 + tmp * tmp);
        o.sestpr = s1 * scl; o.s = cdivr(cdivr(alpha, s1), scl); o.c = cdivr(cdivr(gamma, s1), scl);
      }
      return o;
    }
    const double zeta1 = absalp / absest, zeta2 = absgam / absest;
    const double bb = (1.0 - zeta1 * zeta1 - zeta2 * zeta2) * 0.5;
    const double cc = zeta1 * zeta1;
    const double tt = (bb > 0.0) ? cc / (bb + sqrt(bb * bb + cc)) : sqrt(bb * bb + cc) - bb;
    cplx sine = cdivr(cdivr(alpha, absest), -tt);
    cplx cosine = cdivr(cdivr(gamma, absest), -(1.0 + tt));
    const double tmp = sqrt(cabs2(sine) + cabs2(cosine));
    o.s = cdivr(sine, tmp); o.c = cdivr(cosine, tmp); o.sestpr = sqrt(tt + 1.0) * absest;
    return o;
  }
  // job 2
  if (sest == 0.0) {
    o.sestpr = 0.0;
    cplx sine, cosine;
    if (fmax(absgam, absalp) == 0.0) { sine = one; cosine = zero; }
    else { sine = make_double2(-gamma.x, gamma.y); cosine = make_double2(alpha.x, -alpha.y); }
    const double s1 = fmax(sqrt(cabs2(sine)), sqrt(cabs2(cosine)));
    cplx s = cdivr(sine, s1), c = cdivr(cosine, s1);
    const double tmp = sqrt(cabs2(s) + cabs2(c));
    o.s = cdivr(s, tmp); o.c = cdivr(c, tmp); return o;
  }
  if (absgam <= eps * absest) { o.s = zero; o.c = one; o.sestpr = absgam; return o; }
  if (absalp <= eps * absest) {
    const double s1 = absgam, s2 = absest;
    if (s1 <= s2) { o.s = zero; o.c = one; o.sestpr = s1; }
    else { o.s = one; o.c = zero; o.sestpr = s2; }
    return o;
  }
  if (absest <= eps * absalp || absest <= eps * absgam) {
    const double s1 = absgam, s2 = absalp;
    const cplx ncg = make_double2(-gamma.x, gamma.y), ca = make_double2(alpha.x, -alpha.y);
    if (s1 <= s2) {
      const double tmp = s1 / s2, scl = sqrt(1.0 + tmp * tmp);
      o.sestpr = absest * (tmp / scl); o.s = cdivr(cdivr(ncg, s2), scl); o.c = cdivr(cdivr(ca, s2), scl);
    } else {
      const double tmp = s2 / s1, scl = sqrt(1.0 + tmp * tmp);
      o.sestpr = absest / scl; o.s = cdivr(cdivr(ncg, s1), scl); o.c = cdivr(cdivr(ca, s1), scl);
    }
    return o;
  }
  const double zeta1 = absalp / absest, zeta2 = absgam / absest;
  const double norma = fmax(1.0 + zeta1 * zeta1 + zeta1 * zeta2, zeta1 * zeta2 + zeta2 * zeta2);
  const double test = 1.0 + 2.0 * (zeta1 - zeta2) * (zeta1 + zeta2);
  cplx sine, cosine;
  if (test >= 0.0) {
    const double bb = (zeta1 * zeta1 + zeta2 * zeta2 + 1.0) * 0.5;
    const double cc = zeta2 * zeta2;
    const double tt = cc / (bb + sqrt(fabs(bb * bb - cc)));
    sine = cdivr(cdivr(alpha, absest), 1.0 - tt);
    cosine = cdivr(cdivr(gamma, absest), -tt);
    o.sestpr = sqrt(tt + 4.0 * eps * eps * norma) * absest;
  } else {
    const double bb = (zeta2 * zeta2 + zeta1 * zeta1 - 1.0) * 0.5;
    const double cc = zeta1 * zeta1;
    const double tt = (bb >= 0.0) ? -cc / (bb + sqrt(bb * bb + cc)) : bb - sqrt(bb * bb + cc);
    sine = cdivr(cdivr(alpha, absest), -tt);
    cosine = cdivr(cdivr(gamma, absest), -(1.0 + tt));
    o.sestpr = sqrt(1.0 + tt + 4.0 * eps * eps * norma) * absest;
  }
  const double tmp = sqrt(cabs2(sine) + cabs2(cosine));
  o.s = cdivr(sine, tmp); o.c = cdivr(cosine, tmp);
  return o;
}

// The rank loop of zgelsy on R (held as w[piv[i]][j] = R[j][i], j <= i).  One CTA per matrix.
constexpr int ICE_THREADS = 256;
__global__ void __launch_bounds__(ICE_THREADS)
gelsy_rank_kernel(const cplx* __restrict__ Wall, long strideW, const int* __restrict__ pivall, int n, double rcond,
                  cplx* __restrict__ xwork, int* __restrict__ rank) {
  const int b = blockIdx.x;
  const cplx* W = Wall + (long)b * strideW;
  const int* piv = pivall + (long)b * n;
  cplx* xmin = xwork + (long)b * 2 * n;
  cplx* xmax = xmin + n;
  __shared__ double red[ICE_THREADS / 32][4];
  __shared__ cplx s_sc[4];       // s1, c1, s2, c2
  __shared__ double s_est[2];    // smin, smax
  __shared__ int s_go;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    const cplx r00 = W[(long)piv[0] * n];
    const double a = sqrt(cabs2(r00));
    s_est[0] = a; s_est[1] = a;
    xmin[0] = make_double2(1.0, 0.0); xmax[0] = make_double2(1.0, 0.0);
    s_go = (a != 0.0);
  }
  __syncthreads();
  if (!s_go) { if (tid == 0) rank[b] = 0; return; }
  int r = 1;
  while (r < n) {
    const cplx* wc = W + (long)piv[r] * n;     // R[0..r][r]
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    for (int j = tid; j < r; j += ICE_THREADS) {
      const cplx w = wc[j], xm = xmin[j], xM = xmax[j];
      a0 += xm.x * w.x + xm.y * w.y;  a1 += xm.x * w.y - xm.y * w.x;     // conj(x) * w
      b0 += xM.x * w.x + xM.y * w.y;  b1 += xM.x * w.y - xM.y * w.x;
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); b0 = warp_sum(b0); b1 = warp_sum(b1);
    if (lane == 0) { red[warp][0] = a0; red[warp][1] = a1; red[warp][2] = b0; red[warp][3] = b1; }
    __syncthreads();
    if (tid == 0) {
      double s[4] = {0.0, 0.0, 0.0, 0.0};
      for (int w = 0; w < ICE_THREADS / 32; ++w) for (int e = 0; e < 4; ++e) s[e] += red[w][e];
      const cplx gamma = wc[r];
      const Laic1Out mn = laic1(2, s_est[0], make_double2(s[0], s[1]), gamma);
      const Laic1Out mx = laic1(1, s_est[1], make_double2(s[2], s[3]), gamma);
      if (mx.sestpr * rcond <= mn.sestpr) {
        s_sc[0] = mn.s; s_sc[1] = mn.c; s_sc[2] = mx.s; s_sc[3] = mx.c;
        s_est[0] = mn.sestpr; s_est[1] = mx.sestpr;
        s_go = 1;
      } else {
        s_go = 0;
      }
    }
    __syncthreads();
    if (!s_go) break;
    const cplx s1 = s_sc[0], s2 = s_sc[2];
    for (int j = tid; j < r; j += ICE_THREADS) { xmin[j] = cmul(s1, xmin[j]); xmax[j] = cmul(s2, xmax[j]); }
    if (tid == 0) { xmin[r] = s_sc[1]; xmax[r] = s_sc[3]; }
    ++r;
    __syncthreads();
  }
  if (tid == 0) rank[b] = r;
}

// ---------------------------------------------------------------------------------------------
// Ingredients of the three operators (all tiny, batched over q).
//   dinv[k]   = 1 / |R[k][k]|  (k < rank), 0 beyond
//   S[k][j]   = 1/tau_k (j == k < rank, tau_k != 0) | 1 (j == k otherwise) | G[k][j] (k < j < rank) | 0
//               -- the inverse of the compact-WY factor T of Q = H_0 ... H_{rank-1} = I - V T V^H
//   V1H[k][i] = conj(V[k][i]) for k <= i < rank (upper triangular), 0 elsewhere
__global__ void gelsy_extract_kernel(const cplx* __restrict__ Gall, const cplx* __restrict__ tauall,
                                     const int* __restrict__ rank, const cplx* __restrict__ Vall, long strideV,
                                     const cplx* __restrict__ Wall, long strideW, const int* __restrict__ pivall, int n,
                                     int rP, cplx* __restrict__ Sall, cplx* __restrict__ V1Hall,
                                     double* __restrict__ dinvall) {
  const int b = blockIdx.z;
  const int r = rank[b];
  const int k = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= rP || j >= rP) return;
  const cplx* G = Gall + (long)b * rP * rP;
  const cplx tau = (k < n) ? tauall[(long)b * n + k] : make_double2(0.0, 0.0);
  cplx s = make_double2(0.0, 0.0), v = make_double2(0.0, 0.0);
  if (j == k) {
    const double den = tau.x * tau.x + tau.y * tau.y;
    s = (k < r && den != 0.0) ? make_double2(tau.x / den, -tau.y / den) : make_double2(1.0, 0.0);
  } else if (k < j && j < r) {
    s = G[(long)j * rP + k];       // g = Vt Vt^H = (V^H V)^T: the strict upper triangle of V^H V sits in g's lower one
  }
  if (k <= j && j < r) { const cplx a = Vall[(long)b * strideV + (long)k * n + j]; v = make_double2(a.x, -a.y); }
  Sall[(long)b * rP * rP + (long)k * rP + j] = s;
  V1Hall[(long)b * rP * rP + (long)k * rP + j] = v;
  if (j == 0) {
    double d = 0.0;
    if (k < r) {
      const cplx rkk = Wall[(long)b * strideW + (long)pivall[(long)b * n + k] * n + k];
      d = 1.0 / sqrt(cabs2(rkk));
    }
    dinvall[(long)b * rP + k] = d;
  }
}

// rhat[k][c] = R[k][c] / |R[k][k]| = w[c][k] * dinv[k] for pos[c] >= k, k < rank; 0 elsewhere.  (rP x n, original
// column order, so that the orthonormalised rows are E^H directly.)
__global__ void gelsy_rhat_kernel(const cplx* __restrict__ Wall, long strideW, const int* __restrict__ posall,
                                  const double* __restrict__ dinvall, const int* __restrict__ rank, int n, int rP,
                                  cplx* __restrict__ Rh) {
  __shared__ cplx tile[32][33];
  const int b = blockIdx.z;
  const int r = rank[b];
  const int c0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;     // 32 x 8
  for (int yy = ty; yy < 32; yy += 8) {             // read w[c0+yy][k0+tx]
    const int c = c0 + yy, k = k0 + tx;
    cplx v = make_double2(0.0, 0.0);
    if (c < n && k < r && k < n && posall[(long)b * n + c] >= k) v = Wall[(long)b * strideW + (long)c * n + k];
    tile[yy][tx] = v;
  }
  __syncthreads();
  for (int yy = ty; yy < 32; yy += 8) {             // write rhat[k0+yy][c0+tx]
    const int k = k0 + yy, c = c0 + tx;
    if (k < rP && c < n) {
      const cplx v = tile[tx][yy];
      const double d = dinvall[(long)b * rP + k];
      Rh[(long)b * rP * n + (long)k * n + c] = make_double2(v.x * d, v.y * d);
    }
  }
}

// q1s[i][j] = (delta_ij - vm[i][j]) * dinv[j]  (j < rank), 0 beyond: the first `rank` columns of Q, scaled by 1/|R_jj|.
__global__ void gelsy_q1_finish_kernel(cplx* __restrict__ VM, const double* __restrict__ dinvall,
                                       const int* __restrict__ rank, int n, int rP) {
  const int b = blockIdx.z;
  const int r = rank[b];
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j >= rP) return;
  cplx* q = VM + (long)b * n * rP + (long)i * rP + j;
  if (j >= r) { *q = make_double2(0.0, 0.0); return; }
  const double d = dinvall[(long)b * rP + j];
  const cplx v = *q;
  *q = make_double2(((i == j ? 1.0 : 0.0) - v.x) * d, -v.y * d);
}

// Transposed form: on entry t1[j][i] = sum_k M[k][j] V[k][i] (= conj of (V M)^H); on return
// q1h[j][i] = conj(q1s[i][j]) = (delta_ij - conj(t1[j][i])) * dinv[j]  (rows j >= rank zero): D^-1 Q1^H, [rP x n].
__global__ void gelsy_q1h_finish_kernel(cplx* __restrict__ T1, const double* __restrict__ dinvall,
                                        const int* __restrict__ rank, int n, int rP) {
  const int b = blockIdx.z;
  const int r = rank[b];
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= rP || i >= n) return;
  cplx* q = T1 + (long)b * rP * n + (long)j * n + i;
  if (j >= r) { *q = make_double2(0.0, 0.0); return; }
  const double d = dinvall[(long)b * rP + j];
  const cplx v = *q;
  *q = make_double2(((i == j ? 1.0 : 0.0) - v.x) * d, v.y * d);
}

// w = (w + w^H) / 2 with an exactly real diagonal
__global__ void hermitize_kernel(cplx* __restrict__ Wm, int n) {
  const int b = blockIdx.z;
  const int i = blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j > i) return;
  cplx* M = Wm + (long)b * n * n;
  const cplx a = M[(long)i * n + j], c = M[(long)j * n + i];
  const double re = 0.5 * (a.x + c.x), im = (i == j) ? 0.0 : 0.5 * (a.y - c.y);
  M[(long)i * n + j] = make_double2(re, im);
  M[(long)j * n + i] = make_double2(re, -im);
}

}  // namespace isdf

using namespace isdf;

static size_t qr_smem_bytes(int n, int ncc, int nb, int* nvb_out) {
  const int nvb = (n > nb * QR_RT_MIN ? n : nb * QR_RT_MIN) + QR_NSL;   // + slack of the logical-slice layout
  if (nvb_out) *nvb_out = nvb;
  return (size_t)nvb * 16 + (size_t)nb * ncc * 16 + (size_t)ncc * (16 + 8 + 8 + 4 + 4);
}

/* a: [batch][n][n] c128, COLUMN-major working copy of the matrices to factor (a[c][i] = A[i][c]); on return
 * a[c][k] = R[k][c] for k <= pos(c).  vt [batch][n][n]: row k = reflector v_k (v_k[k] = 1, zeros before; an
 * identity reflector is stored as the zero vector).  tau [batch][n] c128.  piv [batch][n]: position -> column.
 * pos [batch][n]: column -> position. */
extern "C" int isdf_qrcp(void* hv, void* a, int n, int batch, void* vt, void* tau, int* piv, int* pos, void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, a && vt && tau && piv && pos, "null pointer");
  ISDF_CHECK_ARG(h, n >= 1 && batch >= 1 && batch <= 65535, "shape");
  // 16-CTA clusters (9 resident at a time) up to 18 matrices, 8-CTA clusters (18 at a time) beyond -- measured on B200
  // (tools/qrcp_cs_sweep.sh), 16 vs 8: n = 3120: 5 matrices 336 vs 552 ms, 18 matrices 1030 vs 1145 ms (two waves of 9
  // beat one wave of 18), 36 matrices 2029 vs 1782 ms; one matrix of 1000: 25.6 vs 40.7 ms, of 1622: 60.7 vs 98.9 ms;
  // 14 matrices of 520: 19.0 vs 12.6 ms.  The factorisation is bit-identical for every cluster size.
  int cs = (n >= 768 && batch <= 18) ? 16 : 8;
  if (n < 64) cs = 1; else if (n < 256) cs = 2;
  if (const char* e = getenv("ISDF_QR_CS")) { const int v = atoi(e); if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) cs = v; }   // tuning
  ISDF_CUDA(h, cudaFuncSetAttribute(qrcp_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  int ncc = 0, nb = 0, nvb = 0;
  size_t smem = 0;
  const long budget = (long)h->max_smem_optin - 34 * 1024;      // static shared memory of the kernel
  // The panel width fixes the order of the deferred updates, i.e. the rounding of R: it must not follow the cluster
  // size (which follows the number of matrices per GPU), or an N-GPU build would cut the eps-plateau of |R_kk| at other
  // ranks than the single-GPU build.  It is therefore the width that fits the 8-CTA geometry, for 16-CTA clusters too
  // (n = 3120: 16, although 32 would fit 16 CTAs).
  int nb_ref = QR_NB_MAX;
  while (nb_ref > 4 && (long)qr_smem_bytes(n, (n + 7) / 8, nb_ref, nullptr) > budget) nb_ref /= 2;
  for (;; cs /= 2) {
    ncc = (n + cs - 1) / cs;
    bool ok = ncc <= QR_THREADS * QR_NCOLT;
    if (ok) {
      nb = nb_ref;
      while (nb >= 4 && (long)qr_smem_bytes(n, ncc, nb, &nvb) > budget) nb /= 2;
      ok = nb >= 4;
    }
    if (ok) {
      smem = qr_smem_bytes(n, ncc, nb, &nvb);
      ISDF_CUDA(h, cudaFuncSetAttribute(qrcp_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(cs, batch); cfg.blockDim = dim3(QR_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int nclus = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&nclus, qrcp_cluster_kernel, &cfg);
      if (e == cudaSuccess && nclus >= 1) break;
      (void)cudaGetLastError();
    }
    ISDF_CHECK_ARG(h, cs > 1, "matrix too large for the QRCP cluster kernel");
  }
  ISDF_CUDA(h, cudaMemsetAsync(vt, 0, (size_t)batch * n * n * sizeof(cplx), st));
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs, batch); cfg.blockDim = dim3(QR_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    ISDF_CUDA(h, cudaLaunchKernelEx(&cfg, qrcp_cluster_kernel, (cplx*)a, (long)n * n, n, ncc, nb, nvb, (cplx*)vt,
                                    (long)n * n, (cplx*)tau, pos));
  }
  {
    const long tot = (long)n * batch;
    qrcp_finalize_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(pos, n, batch, piv);
    ISDF_LAUNCH_CHECK(h);
  }
  return ISDF_OK;
}

/* zgelsy's rank decision on the factor left in `a` by isdf_qrcp.  xwork: 2*batch*n c128.  rank [batch] (device). */
extern "C" int isdf_gelsy_rank(void* hv, const void* a, const int* piv, int n, int batch, double rcond, void* xwork,
                               int* rank, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && piv && xwork && rank && n >= 1 && batch >= 1, "args");
  gelsy_rank_kernel<<<batch, ICE_THREADS, 0, (cudaStream_t)stream>>>((const cplx*)a, (long)n * n, piv, n, rcond,
                                                                    (cplx*)xwork, rank);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

/* g [batch][rP][rP] = Vt Vt^H = (V^H V)^T restricted to the first rP reflectors (lower triangle read).  Outputs s, v1h [batch][rP][rP], dinv [batch][rP]. */
extern "C" int isdf_gelsy_extract(void* hv, const void* g, const void* tau, const int* rank, const void* vt,
                                  const void* a, const int* piv, int n, int rP, int batch, void* s, void* v1h,
                                  double* dinv, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, g && tau && rank && vt && a && piv && s && v1h && dinv, "null pointer");
  ISDF_CHECK_ARG(h, n >= 1 && rP >= 1 && batch >= 1 && batch <= 65535, "shape");
  dim3 block(32, 8), grid((rP + 31) / 32, (rP + 7) / 8, batch);
  gelsy_extract_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const cplx*)g, (const cplx*)tau, rank,
                                                                 (const cplx*)vt, (long)n * n, (const cplx*)a,
                                                                 (long)n * n, piv, n, rP, (cplx*)s, (cplx*)v1h, dinv);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_gelsy_rhat(void* hv, const void* a, const int* pos, const double* dinv, const int* rank, int n,
                               int rP, int batch, void* rhat, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && pos && dinv && rank && rhat, "null pointer");
  ISDF_CHECK_ARG(h, n >= 1 && rP >= 1 && batch >= 1 && batch <= 65535, "shape");
  dim3 block(32, 8), grid((n + 31) / 32, (rP + 31) / 32, batch);
  gelsy_rhat_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const cplx*)a, (long)n * n, pos, dinv, rank, n, rP,
                                                              (cplx*)rhat);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_gelsy_q1_finish(void* hv, void* vm, const double* dinv, const int* rank, int n, int rP, int batch,
                                    void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, vm && dinv && rank && n >= 1 && rP >= 1 && batch >= 1 && batch <= 65535, "args");
  dim3 block(32, 8), grid((rP + 31) / 32, (n + 7) / 8, batch);
  gelsy_q1_finish_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((cplx*)vm, dinv, rank, n, rP);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_gelsy_q1h_finish(void* hv, void* t1, const double* dinv, const int* rank, int n, int rP, int batch,
                                     void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, t1 && dinv && rank && n >= 1 && rP >= 1 && batch >= 1 && batch <= 65535, "args");
  dim3 block(32, 8), grid((n + 31) / 32, (rP + 7) / 8, batch);
  gelsy_q1h_finish_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((cplx*)t1, dinv, rank, n, rP);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_hermitize(void* hv, void* w, int n, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, w && n >= 1 && batch >= 1 && batch <= 65535, "args");
  dim3 block(32, 8), grid((n + 31) / 32, (n + 7) / 8, batch);
  hermitize_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((cplx*)w, n);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

// c[z][i][j] = sum_l a[z][l][i] * b[z][l][j]   (A^T B without conjugation; a [k][m], b [k][n] row-major)
extern "C" int isdf_gemm_tn(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                            void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && b && c, "null pointer");
  ISDF_CHECK_ARG(h, m >= 0 && n >= 0 && k >= 0 && batch >= 0 && batch <= 65535, "shape");
  GemmParams p;
  p.A = (const cplx*)a; p.lda = lda; p.strideA = strideA;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = (cplx*)c; p.ldc = ldc; p.strideC = strideC;
  p.M = m; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  ISDF_CUDA(h, (launch_gemm<128, 64, true, true, MODE_AB, false, EPI_STORE>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}
