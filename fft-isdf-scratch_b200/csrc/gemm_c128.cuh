// FP64 complex GEMM tile engine for sm_100a.
//
// Complex128 operands stay in numpy/torch interleaved (re,im) layout end to end.  Tiles are
// staged global -> shared with a cp.async (LDGSTS) ring; each lane pulls one complex
// element per fragment with a single conflict-free LDS.128 and feeds the re/im halves to the
// FP64 tensor pipe (mma.sync.m8n8k4.f64 == DMMA.8x8x4, the only FP64 MMA sm_100a has; tcgen05
// has no f64 kind).
//
// Complex products use the 3M (Karatsuba) form by default: with P1 = sum ar*br, P2 = sum ai*bi and
// P3 = sum (ar+ai)(br+bi), re = P1 - P2 and im = P3 - P1 - P2 -- three real MMAs per complex MAC instead of
// four, i.e. 25 % less work on the pipe that bounds every dense stage (the two extra additions per
// fragment element are noise).  The error is bounded normwise, ~2x the 4M constant, which is what the
// 1e-10 parity budget is stated in.  ISDF_GEMM_3M=0 builds the classical 4M form.  Real-only products
// (selection Gram) need two MMAs either way.
//
// Default configuration (measured on B200, profiles/r1_gemm_variants.txt): 256 threads, K step 16 complex,
// 2-stage cp.async ring, TWO CTAs per SM so that one CTA's barrier / prologue / epilogue bubbles are filled
// by the other.  4M: CTA tile 64x64, 8 warps of 32(m) x 16(n) (92 % of cuBLAS DGEMM).  3M keeps three
// accumulator sets, so the warp tile is 16 x 16 and the CTA tile 64x32 to stay within 128 registers.
// ISDF_GEMM_SMALL=0 selects the 512-thread 128x64 / 64x128 4M tiles (one CTA per SM, 86 %).
//
// Operand layouts: KCONTIG = [rows][K] row-major (K fastest), KSLOW = [K][rows] (rows fastest).
#pragma once
#include "common.cuh"

namespace isdf {

enum { MODE_CONJA = 0, MODE_CONJB = 1, MODE_AB = 2 };  // conj(a)*b, a*conj(b), a*b
enum { EPI_STORE = 0, EPI_HERK = 2, EPI_SQ_SYM = 3, EPI_SUB_LOWER = 4 };

struct GemmParams {
  const cplx* A; long lda; long strideA;
  const cplx* B; long ldb; long strideB;
  cplx* C;       long ldc; long strideC;
  int M, N, K;
  int nseg; long segA; long segB;  // accumulate over nseg K-segments (pointer += seg per segment)
  double alpha;
  const int* perm; long stridePerm;  // EPI_HERK: optional row/col scatter map per batch
  const int* active;                 // optional per-batch flag; batch skipped when 0
  int ksplit; int kchunk; long strideSplit;  // split-K: grid.z = batch*ksplit, partial results at C + ks*strideSplit
  // L2 rasterisation (plain EPI_STORE products only): CTAs are dealt in groups of group_m consecutive M-tiles per
  // N-tile, so that a B column panel fetched from DRAM is shared by group_m resident CTAs instead of being re-read for
  // every M-tile (ncu on the NiO fit GEMM: 1.23 TB of DRAM reads per launch for 32 GB of operands without it)
  int group_m = 0;
  // EPI_HERK across GPUs: per-batch destination (e.g. a slab in the OWNING rank's NVLink peer-mapped memory), written
  // with plain stores as the tiles finish, lower triangle only (the owner mirrors when it sums the slabs)
  cplx* const* Cbatch = nullptr;
  int lower_only = 0;
};

#ifndef ISDF_GEMM_BK
#define ISDF_GEMM_BK 16       // complex K elements per pipeline stage of the 4M / real-only kernels
#endif
#ifndef ISDF_GEMM_3M_BK
#define ISDF_GEMM_3M_BK 32    // ... of the 3M kernels (64x32 tiles: 2 x 54 KB per CTA, still two CTAs per SM)
#endif
#ifndef ISDF_GEMM_STAGES
#define ISDF_GEMM_STAGES 2
#endif
#ifndef ISDF_GEMM_SMALL
#define ISDF_GEMM_SMALL 1     // 1: 64x64 tiles, 256 threads, 2 CTAs per SM; 0: 128x64 / 64x128 tiles, 512 threads
#endif
#ifndef ISDF_GEMM_3M
#define ISDF_GEMM_3M 1        // 1: Karatsuba complex product (3 DMMA per complex MAC); 0: classical 4
#endif
#ifndef ISDF_GEMM_3M_BM
#define ISDF_GEMM_3M_BM 64    // CTA tile of the 3M kernels (warp tile 16x16)
#endif
#ifndef ISDF_GEMM_3M_BN
#define ISDF_GEMM_3M_BN 32
#endif
constexpr int GEMM_STAGES = ISDF_GEMM_STAGES;
__host__ __device__ constexpr bool gemm_is_3m(bool real_only) { return ISDF_GEMM_3M != 0 && !real_only; }
__host__ __device__ constexpr int gemm_bk(bool real_only) {   // multiple of 4
  return gemm_is_3m(real_only) ? ISDF_GEMM_3M_BK : ISDF_GEMM_BK;
}
#ifndef ISDF_GEMM_WM_3M
#define ISDF_GEMM_WM_3M 16    // warp tile rows of the 3M kernels (tuning)
#endif
#ifndef ISDF_GEMM_WM_4M
#define ISDF_GEMM_WM_4M 32
#endif
#ifndef ISDF_GEMM_4M_BN
#define ISDF_GEMM_4M_BN 64
#endif
#ifndef ISDF_GEMM_MINB
#define ISDF_GEMM_MINB 2      // resident CTAs per SM asked of the compiler for <= 256-thread tiles
#endif
__host__ __device__ constexpr int gemm_wm(bool real_only) {   // warp tile rows
  return gemm_is_3m(real_only) ? ISDF_GEMM_WM_3M : (real_only ? 32 : ISDF_GEMM_WM_4M);
}
__host__ __device__ constexpr int gemm_threads(int BM, int BN, bool real_only) {
  return (BM / gemm_wm(real_only)) * (BN / 16) * 32;
}
__host__ __device__ constexpr int gemm_min_blocks(int BM, int BN, bool real_only) {
  return gemm_threads(BM, BN, real_only) <= 256 ? ISDF_GEMM_MINB : 1;
}

template <int BM, int BN, bool A_KSLOW, bool B_KSLOW, int BK>
struct GemmSmem {
  static constexpr int LDA_S = A_KSLOW ? (BM + 2) : (BK + 4);
  static constexpr int LDB_S = B_KSLOW ? (BN + 2) : (BK + 4);
  static constexpr int A_TILE = A_KSLOW ? BK * LDA_S : BM * LDA_S;
  static constexpr int B_TILE = B_KSLOW ? BK * LDB_S : BN * LDB_S;
  static constexpr int BYTES = GEMM_STAGES * (A_TILE + B_TILE) * (int)sizeof(cplx);
};

template <int BM, int BN, bool A_KSLOW, bool B_KSLOW, int MODE, bool REAL_ONLY, int EPI>
__global__ void __launch_bounds__(gemm_threads(BM, BN, REAL_ONLY), gemm_min_blocks(BM, BN, REAL_ONLY))
    gemm_c128_kernel(GemmParams p) {
  constexpr bool K3M = gemm_is_3m(REAL_ONLY);
  constexpr int WM = gemm_wm(REAL_ONLY), MI = WM / 8;     // warp tile WM x 16 = MI x 2 DMMA tiles
  static_assert(BM % WM == 0 && BN % 16 == 0, "warp tile is WM x 16");
  constexpr int GEMM_THREADS = gemm_threads(BM, BN, REAL_ONLY);
  static_assert(GEMM_THREADS <= 1024, "CTA tile too large for this warp tile");
  constexpr int BK = gemm_bk(REAL_ONLY), STAGES = GEMM_STAGES;
  using S = GemmSmem<BM, BN, A_KSLOW, B_KSLOW, BK>;
  constexpr int LDA_S = S::LDA_S, LDB_S = S::LDB_S, A_TILE = S::A_TILE, B_TILE = S::B_TILE;
  constexpr int WARPS_N = BN / 16;
  constexpr bool SYMM = (EPI == EPI_HERK || EPI == EPI_SQ_SYM || EPI == EPI_SUB_LOWER);

  extern __shared__ __align__(16) unsigned char smem_raw[];
  cplx* sA = reinterpret_cast<cplx*>(smem_raw);
  cplx* sB = sA + STAGES * A_TILE;

  int bz = blockIdx.z;
  int ksp = 0;
  if (p.ksplit > 1) { ksp = bz % p.ksplit; bz /= p.ksplit; }
  if (p.active != nullptr && p.active[bz] == 0) return;
  int tile_x = blockIdx.x, tile_y = blockIdx.y;
  if (!SYMM && p.group_m > 1) {
    const long id = (long)blockIdx.y * gridDim.x + blockIdx.x;
    const long per = (long)p.group_m * gridDim.x;
    const int g0 = (int)(id / per) * p.group_m;
    const int rows = min(p.group_m, (int)gridDim.y - g0);
    const long r = id - (long)(id / per) * per;
    tile_y = g0 + (int)(r % rows);
    tile_x = (int)(r / rows);
  }
  const int m0 = tile_y * BM, n0 = tile_x * BN;
  if (SYMM && n0 >= m0 + BM) return;  // tile strictly above the diagonal: produced by mirroring

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / WARPS_N) * WM, wn0 = (warp % WARPS_N) * 16;
  const int M = p.M, N = p.N;
  int K = p.K;
  const cplx* Abase = p.A + (long)bz * p.strideA;
  const cplx* Bbase = p.B + (long)bz * p.strideB;
  if (p.ksplit > 1) {   // this CTA contracts over k in [k0, k0 + kchunk)
    const int k0 = ksp * p.kchunk;
    Abase += A_KSLOW ? (long)k0 * p.lda : (long)k0;
    Bbase += B_KSLOW ? (long)k0 * p.ldb : (long)k0;
    K = (K - k0 < p.kchunk) ? (K - k0) : p.kchunk;
    if (K < 0) K = 0;
  }
  const int ktiles = (K + BK - 1) / BK;
  const int nit = p.nseg * ktiles;
  // Ragged edges cost what they use: a warp whose block has no valid row/column (or, for the symmetric
  // epilogues, lies strictly above the diagonal) still helps stage the tiles but issues no DMMA.
  const bool warp_live = (m0 + wm0 < M) && (n0 + wn0 < N) && !(SYMM && n0 + wn0 > m0 + wm0 + WM - 1);

  auto load_tiles = [&](int it, int slot) {
    const int seg = it / ktiles;
    const int k0 = (it - seg * ktiles) * BK;
    const cplx* Ag = Abase + (long)seg * p.segA;
    const cplx* Bg = Bbase + (long)seg * p.segB;
    cplx* dA = sA + slot * A_TILE;
    cplx* dB = sB + slot * B_TILE;
    if (!A_KSLOW) {
#pragma unroll
      for (int c = tid; c < BM * BK; c += GEMM_THREADS) {
        const int row = c / BK, kk = c % BK;
        const bool v = (m0 + row < M) && (k0 + kk < K);
        const cplx* src = v ? Ag + (long)(m0 + row) * p.lda + (k0 + kk) : Ag;
        cp_async16(dA + row * LDA_S + kk, src, v);
      }
    } else {
#pragma unroll
      for (int c = tid; c < BK * BM; c += GEMM_THREADS) {
        const int kk = c / BM, m = c % BM;
        const bool v = (k0 + kk < K) && (m0 + m < M);
        const cplx* src = v ? Ag + (long)(k0 + kk) * p.lda + (m0 + m) : Ag;
        cp_async16(dA + kk * LDA_S + m, src, v);
      }
    }
    if (!B_KSLOW) {
#pragma unroll
      for (int c = tid; c < BN * BK; c += GEMM_THREADS) {
        const int row = c / BK, kk = c % BK;
        const bool v = (n0 + row < N) && (k0 + kk < K);
        const cplx* src = v ? Bg + (long)(n0 + row) * p.ldb + (k0 + kk) : Bg;
        cp_async16(dB + row * LDB_S + kk, src, v);
      }
    } else {
#pragma unroll
      for (int c = tid; c < BK * BN; c += GEMM_THREADS) {
        const int kk = c / BN, n = c % BN;
        const bool v = (k0 + kk < K) && (n0 + n < N);
        const cplx* src = v ? Bg + (long)(k0 + kk) * p.ldb + (n0 + n) : Bg;
        cp_async16(dB + kk * LDB_S + n, src, v);
      }
    }
  };

  // 4M: acc_re / acc_im.  3M: acc_re = P1 = sum ar*br, acc_im = P2 = sum ai*bi, acc_p3 = sum (ar +- ai)(br +- bi)
  double acc_re[MI][2][2];
  double acc_im[REAL_ONLY ? 1 : MI][2][2];
  double acc_p3[K3M ? MI : 1][2][2];
#pragma unroll
  for (int mi = 0; mi < MI; ++mi)
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
      acc_re[mi][ni][0] = 0.0; acc_re[mi][ni][1] = 0.0;
      if (!REAL_ONLY) { acc_im[mi][ni][0] = 0.0; acc_im[mi][ni][1] = 0.0; }
      if (K3M) { acc_p3[mi][ni][0] = 0.0; acc_p3[mi][ni][1] = 0.0; }
    }

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nit) load_tiles(s, s);
    cp_async_commit();
  }

  for (int it = 0; it < nit; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nx = it + STAGES - 1;
      if (nx < nit) load_tiles(nx, nx % STAGES);
      cp_async_commit();
    }
    if (!warp_live) continue;   // warp-uniform: this warp's block lies outside M x N (or above the diagonal)
    const cplx* tA = sA + (it % STAGES) * A_TILE;
    const cplx* tB = sB + (it % STAGES) * B_TILE;
    const int kleft = K - (it % ktiles) * BK;   // a short last K tile (e.g. K = nao = 26) skips its all-zero steps
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      if (ks * 4 >= kleft) break;
      cplx a[MI], b[2];
#pragma unroll
      for (int mi = 0; mi < MI; ++mi)
        a[mi] = A_KSLOW ? tA[(ks * 4 + t) * LDA_S + wm0 + mi * 8 + g] : tA[(wm0 + mi * 8 + g) * LDA_S + ks * 4 + t];
#pragma unroll
      for (int ni = 0; ni < 2; ++ni)
        b[ni] = B_KSLOW ? tB[(ks * 4 + t) * LDB_S + wn0 + ni * 8 + g] : tB[(wn0 + ni * 8 + g) * LDB_S + ks * 4 + t];
      if constexpr (K3M) {
        // conj(a) b: (ar - ai)(br + bi);  a conj(b): (ar + ai)(br - bi);  a b: (ar + ai)(br + bi)
        double as[MI], bs[2];
#pragma unroll
        for (int mi = 0; mi < MI; ++mi) as[mi] = (MODE == MODE_CONJA) ? a[mi].x - a[mi].y : a[mi].x + a[mi].y;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) bs[ni] = (MODE == MODE_CONJB) ? b[ni].x - b[ni].y : b[ni].x + b[ni].y;
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) {
            dmma884(acc_re[mi][ni][0], acc_re[mi][ni][1], a[mi].x, b[ni].x);
            dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], a[mi].y, b[ni].y);
            dmma884(acc_p3[mi][ni][0], acc_p3[mi][ni][1], as[mi], bs[ni]);
          }
      } else {
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) {
          const double br = b[ni].x, bi = b[ni].y;
          const double nbr = -br, nbi = -bi;
#pragma unroll
          for (int mi = 0; mi < MI; ++mi) {
            const double ar = a[mi].x, ai = a[mi].y;
            dmma884(acc_re[mi][ni][0], acc_re[mi][ni][1], ar, br);
            dmma884(acc_re[mi][ni][0], acc_re[mi][ni][1], ai, (MODE == MODE_AB) ? nbi : bi);
            if (!REAL_ONLY) {
              if (MODE == MODE_CONJA) {         // im = ar*bi - ai*br
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ar, bi);
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ai, nbr);
              } else if (MODE == MODE_CONJB) {  // im = ai*br - ar*bi
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ai, br);
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ar, nbi);
              } else {                          // im = ar*bi + ai*br
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ar, bi);
                dmma884(acc_im[mi][ni][0], acc_im[mi][ni][1], ai, br);
              }
            }
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  cplx* Cb = (EPI == EPI_HERK && p.Cbatch != nullptr) ? p.Cbatch[bz] : p.C + (long)bz * p.strideC + (long)ksp * p.strideSplit;
  const int* perm = (EPI == EPI_HERK && p.perm != nullptr) ? p.perm + (long)bz * p.stridePerm : nullptr;
  const double alpha = p.alpha;
#pragma unroll
  for (int mi = 0; mi < MI; ++mi) {
    const int r = m0 + wm0 + mi * 8 + g;
    if (r >= M) continue;
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = n0 + wn0 + ni * 8 + 2 * t + e;
        if (c >= N) continue;
        double vr = acc_re[mi][ni][e];
        double vi = REAL_ONLY ? 0.0 : acc_im[mi][ni][e];
        if (K3M) {
          const double p1 = vr, p2 = vi, p3 = acc_p3[mi][ni][e];
          if (MODE == MODE_AB) { vr = p1 - p2; vi = (p3 - p1) - p2; }
          else                 { vr = p1 + p2; vi = (p3 - p1) + p2; }
        }
        if (EPI == EPI_STORE) {
          Cb[(long)r * p.ldc + c] = make_double2(alpha * vr, alpha * vi);
        } else if (EPI == EPI_SUB_LOWER) {   // only the lower triangle is kept current (no mirrored traffic)
          if (r >= c) {
            cplx* q = Cb + (long)r * p.ldc + c;
            cplx o = *q;
            *q = make_double2(o.x - vr, o.y - vi);
          }
        } else if (EPI == EPI_HERK) {
          if (r >= c) {
            const int pr = perm ? perm[r] : r, pc = perm ? perm[c] : c;
            // the diagonal of B B^H is real: store an exact zero imaginary part
            Cb[(long)pr * p.ldc + pc] = make_double2(alpha * vr, (r == c) ? 0.0 : alpha * vi);
            if (r > c && !p.lower_only) Cb[(long)pc * p.ldc + pr] = make_double2(alpha * vr, -alpha * vi);
          }
        } else {  // EPI_SQ_SYM: out = alpha * re^2 (real, stored as complex with zero imaginary part)
          if (r >= c) {
            const double v = alpha * vr * vr;
            Cb[(long)r * p.ldc + c] = make_double2(v, 0.0);
            if (r > c) Cb[(long)c * p.ldc + r] = make_double2(v, 0.0);
          }
        }
      }
    }
  }
}

template <int BM_, int BN_, bool A_KSLOW, bool B_KSLOW, int MODE, bool REAL_ONLY, int EPI>
inline cudaError_t launch_gemm(const GemmParams& p, int batch, cudaStream_t st) {
  // 3M: the caller's aspect hint (BM_ < BN_: short and wide, e.g. the 64-row sweep blocks and their ragged tail)
  // picks the transposed tile, so that a block row of <= 32 live rows keeps every warp busy.
  constexpr bool K3M = gemm_is_3m(REAL_ONLY);
  constexpr bool WIDE = BM_ < BN_;
  constexpr int BM = K3M ? (WIDE ? ISDF_GEMM_3M_BN : ISDF_GEMM_3M_BM) : (ISDF_GEMM_SMALL ? 64 : BM_);
  constexpr int BN = K3M ? (WIDE ? ISDF_GEMM_3M_BM : ISDF_GEMM_3M_BN)
                         : (ISDF_GEMM_SMALL ? (REAL_ONLY ? 64 : ISDF_GEMM_4M_BN) : BN_);
  using S = GemmSmem<BM, BN, A_KSLOW, B_KSLOW, gemm_bk(REAL_ONLY)>;
  auto kern = gemm_c128_kernel<BM, BN, A_KSLOW, B_KSLOW, MODE, REAL_ONLY, EPI>;
  // the opt-in above 48 KB is a per-device attribute: remember it per device, not per process
  static bool configured[64] = {};
  int dev = 0;
  cudaError_t eg = cudaGetDevice(&dev);
  if (eg != cudaSuccess) return eg;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  if (p.M <= 0 || p.N <= 0 || batch <= 0) return cudaSuccess;
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, batch * (p.ksplit > 1 ? p.ksplit : 1));
  kern<<<grid, gemm_threads(BM, BN, REAL_ONLY), S::BYTES, st>>>(p);
  return cudaGetLastError();
}

}  // namespace isdf
