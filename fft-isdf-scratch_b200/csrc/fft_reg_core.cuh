// Register-resident FFT building blocks (/root/reference/fftisdf.py:113 pbctools.fft).
//
// Every axis length N handled here is either
//   * a two-factor length N = R1 * R2 (R1, R2 <= 16): one thread takes the R1 inputs of a butterfly column from global
//     memory straight into registers, runs a hard-coded R1-point DFT, multiplies by the stage twiddles, and the
//     ONLY shared-memory exchange of the axis hands the R2 inputs of the second butterfly to another thread; or
//   * a prime N <= 61: a direct symmetric DFT (inputs paired as x_j +- x_{N-j}: N-1 real-coefficient FMAs per
//     output pair instead of a complex product per term), output pairs dealt over G thread groups.
// All index arithmetic is compile-time (N, R1, R2, thread count are template parameters), so a stage is loads, FP64
// math and stores only.  The functions are __host__ __device__ so that tools/fft_reg_host_check.cu can run the very same
// phases thread by thread on the CPU against a naive DFT (index maps, twiddles, butterflies) without a GPU.
#pragma once
#include "common.cuh"
#include "fft_roots.cuh"

#define ISDF_HD __host__ __device__ __forceinline__

namespace isdf {
namespace fftreg {

ISDF_HD cplx c_add(cplx a, cplx b) { return make_double2(a.x + b.x, a.y + b.y); }
ISDF_HD cplx c_sub(cplx a, cplx b) { return make_double2(a.x - b.x, a.y - b.y); }
ISDF_HD cplx c_mul(cplx a, cplx b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
ISDF_HD cplx c_mul_mi(cplx a) { return make_double2(a.y, -a.x); }   // a * (-i)

// a * exp(-2 pi i m / N); m is a compile-time value after unrolling, so the branches fold
template <int N>
ISDF_HD cplx mul_root(cplx a, int m) {
  m %= N;
  if (m == 0) return a;
  if (4 * m == N) return make_double2(a.y, -a.x);
  if (2 * m == N) return make_double2(-a.x, -a.y);
  if (4 * m == 3 * N) return make_double2(-a.y, a.x);
  const double c = Roots<N>::c(m), s = Roots<N>::s(m);
  return make_double2(fma(a.y, s, a.x * c), fma(-a.x, s, a.y * c));
}

template <int R> ISDF_HD void rdft(cplx (&x)[R]);

// prime R: y_r = A_r - i B_r, y_{R-r} = A_r + i B_r with A_r = x0 + sum_j c(jr) (x_j + x_{R-j}), B_r = sum_j s(jr) (x_j - x_{R-j})
template <int R>
ISDF_HD void rdft_prime(cplx (&x)[R]) {
  constexpr int H = (R - 1) / 2;
  cplx s[H], d[H];
#pragma unroll
  for (int j = 0; j < H; ++j) { s[j] = c_add(x[j + 1], x[R - 1 - j]); d[j] = c_sub(x[j + 1], x[R - 1 - j]); }
  cplx y0 = x[0];
#pragma unroll
  for (int j = 0; j < H; ++j) y0 = c_add(y0, s[j]);
  cplx out[R];
  out[0] = y0;
#pragma unroll
  for (int r = 1; r <= H; ++r) {
    double ar = x[0].x, ai = x[0].y, br = 0.0, bi = 0.0;
#pragma unroll
    for (int j = 1; j <= H; ++j) {
      const double c = Roots<R>::c((j * r) % R), sn = Roots<R>::s((j * r) % R);
      ar = fma(c, s[j - 1].x, ar); ai = fma(c, s[j - 1].y, ai);
      if (j == 1) { br = sn * d[0].x; bi = sn * d[0].y; }
      else { br = fma(sn, d[j - 1].x, br); bi = fma(sn, d[j - 1].y, bi); }
    }
    out[r] = make_double2(ar + bi, ai - br);
    out[R - r] = make_double2(ar - bi, ai + br);
  }
#pragma unroll
  for (int r = 0; r < R; ++r) x[r] = out[r];
}

// Cooley-Tukey in registers, R = Ra * Rb:  j = j1 Rb + j2,  k = k1 + Ra k2
template <int Ra, int Rb>
ISDF_HD void rdft_ct(cplx (&x)[Ra * Rb]) {
  constexpr int R = Ra * Rb;
  cplx y[R];
#pragma unroll
  for (int j2 = 0; j2 < Rb; ++j2) {
    cplx t[Ra];
#pragma unroll
    for (int j1 = 0; j1 < Ra; ++j1) t[j1] = x[j1 * Rb + j2];
    rdft<Ra>(t);
#pragma unroll
    for (int k1 = 0; k1 < Ra; ++k1) y[k1 * Rb + j2] = mul_root<R>(t[k1], j2 * k1);
  }
#pragma unroll
  for (int k1 = 0; k1 < Ra; ++k1) {
    cplx u[Rb];
#pragma unroll
    for (int j2 = 0; j2 < Rb; ++j2) u[j2] = y[k1 * Rb + j2];
    rdft<Rb>(u);
#pragma unroll
    for (int k2 = 0; k2 < Rb; ++k2) x[k1 + Ra * k2] = u[k2];
  }
}

template <int R>
ISDF_HD void rdft(cplx (&x)[R]) {
  if constexpr (R == 1) {
  } else if constexpr (R == 2) {
    const cplx a = x[0], b = x[1];
    x[0] = c_add(a, b); x[1] = c_sub(a, b);
  } else if constexpr (R == 4) {
    const cplx a = c_add(x[0], x[2]), b = c_sub(x[0], x[2]), c = c_add(x[1], x[3]), d = c_mul_mi(c_sub(x[1], x[3]));
    x[0] = c_add(a, c); x[1] = c_add(b, d); x[2] = c_sub(a, c); x[3] = c_sub(b, d);
  } else if constexpr (R == 8) {
    cplx e[4] = {x[0], x[2], x[4], x[6]}, o[4] = {x[1], x[3], x[5], x[7]};
    rdft<4>(e); rdft<4>(o);
    const double h = 0.70710678118654752440;
    const cplx o1 = make_double2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));      // * e^{-i pi/4}
    const cplx o2 = c_mul_mi(o[2]);
    const cplx o3 = make_double2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));     // * e^{-3 i pi/4}
    x[0] = c_add(e[0], o[0]); x[4] = c_sub(e[0], o[0]);
    x[1] = c_add(e[1], o1);   x[5] = c_sub(e[1], o1);
    x[2] = c_add(e[2], o2);   x[6] = c_sub(e[2], o2);
    x[3] = c_add(e[3], o3);   x[7] = c_sub(e[3], o3);
  } else if constexpr (R == 3 || R == 5 || R == 7 || R == 11 || R == 13) {
    rdft_prime<R>(x);
  } else if constexpr (R == 6) {
    rdft_ct<2, 3>(x);
  } else if constexpr (R == 9) {
    rdft_ct<3, 3>(x);
  } else if constexpr (R == 10) {
    rdft_ct<2, 5>(x);
  } else if constexpr (R == 12) {
    rdft_ct<3, 4>(x);
  } else if constexpr (R == 15) {
    rdft_ct<3, 5>(x);
  } else if constexpr (R == 16) {
    rdft_ct<4, 4>(x);
  } else {
    static_assert(R == 1, "unsupported in-register radix");
  }
}

// ---- two-factor axis: geometry of the single in-place exchange buffer of a plane -------------------------------
//   N = R1 R2;  input index n = R2 j1 + j2;  output index k = k1 + R1 k2.
//   Plane buffer P[a][b], a = z slot, b = y slot, pitch PITCH (odd).  After the z axis, z slot a = k1 Q + k2 holds
//   kz = k1 + R1 k2 (Q = R2 rounded up to odd: the y-phase threads run in kz order, i.e. stride Q PITCH, odd).
template <int N_, int R1_, int R2_>
struct TwoFactor {
  static constexpr int N = N_, R1 = R1_, R2 = R2_;
  static_assert(R1_ * R2_ == N_, "N = R1 R2");
  static constexpr int Q = R2_ | 1;
  static constexpr int ROWS = R1_ * Q;
  static constexpr int PITCH = N_ | 1;
  static constexpr int SLOTS = ROWS * PITCH;           // complex elements of the plane buffer
  ISDF_HD static int slot_of_k(int k) { return (k % R1_) * Q + k / R1_; }   // z slot that holds output index k
  ISDF_HD static int slot_of_t(int t) { return (t / R2_) * Q + t % R2_; }   // t-th occupied z slot (slot order)
};

// Phase functions of the fused z+y plane transform (n2 == n3 == N).  `tid` in [0, THREADS).  ld(idx) / st(idx, v):
// element idx = y N + z of the plane in global memory (z contiguous; the pre-phase, the post-weight and -- in the
// multi-GPU kernels -- the NVLink peer addressing live in these functors).  TW[m] = exp(-2 pi i m / N), m < N
// (shared memory).
template <class AX, int THREADS, class LD>
ISDF_HD void plane_z1(int tid, LD&& ld, cplx* __restrict__ P, const cplx* __restrict__ TW) {
  constexpr int N = AX::N, R1 = AX::R1, R2 = AX::R2, TOT = N * R2, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 2
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int y = i / R2, j2 = i % R2;
      cplx v[R1];
#pragma unroll
      for (int j1 = 0; j1 < R1; ++j1) v[j1] = ld(y * N + R2 * j1 + j2);
      rdft<R1>(v);
#pragma unroll
      for (int k1 = 1; k1 < R1; ++k1) v[k1] = c_mul(v[k1], TW[j2 * k1]);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) P[(k1 * AX::Q + j2) * AX::PITCH + y] = v[k1];
    }
  }
}

template <class AX, int THREADS>
ISDF_HD void plane_z2(int tid, cplx* __restrict__ P) {
  constexpr int N = AX::N, R1 = AX::R1, R2 = AX::R2, TOT = N * R1, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int k1 = i / N, y = i % N;
      cplx* p = P + (k1 * AX::Q) * AX::PITCH + y;
      cplx v[R2];
#pragma unroll
      for (int j2 = 0; j2 < R2; ++j2) v[j2] = p[j2 * AX::PITCH];
      rdft<R2>(v);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) p[k2 * AX::PITCH] = v[k2];      // z slot k1 Q + k2 now holds kz = k1 + R1 k2
    }
  }
}

template <class AX, int THREADS>
ISDF_HD void plane_y1(int tid, cplx* __restrict__ P, const cplx* __restrict__ TW) {
  constexpr int N = AX::N, R1 = AX::R1, R2 = AX::R2, TOT = N * R2, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int j2 = i / N, t = i % N;
      cplx* p = P + AX::slot_of_t(t) * AX::PITCH + j2;
      cplx v[R1];
#pragma unroll
      for (int j1 = 0; j1 < R1; ++j1) v[j1] = p[R2 * j1];
      rdft<R1>(v);
#pragma unroll
      for (int k1 = 1; k1 < R1; ++k1) v[k1] = c_mul(v[k1], TW[j2 * k1]);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) p[R2 * k1] = v[k1];             // y slot R2 k1 + j2
    }
  }
}

template <class AX, int THREADS, class ST>
ISDF_HD void plane_y2(int tid, const cplx* __restrict__ P, ST&& st) {
  constexpr int N = AX::N, R1 = AX::R1, R2 = AX::R2, TOT = N * R1, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int k1 = i / N, kz = i % N;
      const cplx* p = P + AX::slot_of_k(kz) * AX::PITCH + R2 * k1;
      cplx v[R2];
#pragma unroll
      for (int j2 = 0; j2 < R2; ++j2) v[j2] = p[j2];
      rdft<R2>(v);
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) st((k1 + R1 * k2) * N + kz, v[k2]);
    }
  }
}

// ---- strided lines (the x pass): a tile of T consecutive lines, element stride `stride` along the line -----------
//   S[(k1 R2 + j2) T + l]; threads run over l fastest, so global accesses are coalesced and shared-memory accesses
//   unit-stride; the twiddle index is warp-uniform.
//   ld(x, l) / st(k, l, v): point x (k) of line l of the tile.
template <class AX, int T, int THREADS, class LD>
ISDF_HD void lines_s1(int tid, LD&& ld, int lcnt, cplx* __restrict__ S, const cplx* __restrict__ TW) {
  constexpr int R1 = AX::R1, R2 = AX::R2, TOT = R2 * T, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 2
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int j2 = i / T, l = i % T;
      if (l < lcnt) {
        cplx v[R1];
#pragma unroll
        for (int j1 = 0; j1 < R1; ++j1) v[j1] = ld(R2 * j1 + j2, l);
        rdft<R1>(v);
#pragma unroll
        for (int k1 = 1; k1 < R1; ++k1) v[k1] = c_mul(v[k1], TW[j2 * k1]);
#pragma unroll
        for (int k1 = 0; k1 < R1; ++k1) S[(k1 * R2 + j2) * T + l] = v[k1];
      }
    }
  }
}

template <class AX, int T, int THREADS, class ST>
ISDF_HD void lines_s2(int tid, const cplx* __restrict__ S, int lcnt, ST&& st) {
  constexpr int R1 = AX::R1, R2 = AX::R2, TOT = R1 * T, IT = (TOT + THREADS - 1) / THREADS;
#pragma unroll 1
  for (int it = 0; it < IT; ++it) {
    const int i = tid + it * THREADS;
    if (i < TOT) {
      const int k1 = i / T, l = i % T;
      if (l < lcnt) {
        cplx v[R2];
#pragma unroll
        for (int j2 = 0; j2 < R2; ++j2) v[j2] = S[(k1 * R2 + j2) * T + l];
        rdft<R2>(v);
#pragma unroll
        for (int k2 = 0; k2 < R2; ++k2) st(k1 + R1 * k2, l, v[k2]);
      }
    }
  }
}

// ---- prime axis: direct symmetric DFT, output pairs dealt over G groups ---------------------------------------
//   load(j) -> x_j,  store(k, value).  Group GI computes the pairs r = GI PG + 1 ... (and y_0 when GI == 0).
template <int N, int G, int GI, class LD, class ST>
ISDF_HD void direct_group(LD&& load, ST&& store) {
  constexpr int H = (N - 1) / 2, PG = (H + G - 1) / G, R0 = GI * PG + 1;
  constexpr int NP = (R0 + PG - 1 <= H) ? PG : (H - R0 + 1 > 0 ? H - R0 + 1 : 0);
  const cplx x0 = load(0);
  cplx y0 = x0;
  double ar[NP > 0 ? NP : 1], ai[NP > 0 ? NP : 1], br[NP > 0 ? NP : 1], bi[NP > 0 ? NP : 1];
#pragma unroll
  for (int p = 0; p < NP; ++p) { ar[p] = x0.x; ai[p] = x0.y; br[p] = 0.0; bi[p] = 0.0; }
#pragma unroll
  for (int j = 1; j <= H; ++j) {
    const cplx a = load(j), b = load(N - j);
    const cplx s = c_add(a, b), d = c_sub(a, b);
    if (GI == 0) y0 = c_add(y0, s);
#pragma unroll
    for (int p = 0; p < NP; ++p) {
      const int m = (j * (R0 + p)) % N;
      const double c = Roots<N>::c(m), sn = Roots<N>::s(m);
      ar[p] = fma(c, s.x, ar[p]); ai[p] = fma(c, s.y, ai[p]);
      br[p] = fma(sn, d.x, br[p]); bi[p] = fma(sn, d.y, bi[p]);
    }
  }
  if (GI == 0) store(0, y0);
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    store(R0 + p, make_double2(ar[p] + bi[p], ai[p] - br[p]));
    store(N - R0 - p, make_double2(ar[p] - bi[p], ai[p] + br[p]));
  }
}

template <int N, int G, int GI, class LD, class ST>
ISDF_HD void direct_dispatch(int g, LD&& load, ST&& store) {
  if constexpr (GI < G) {
    if (g == GI) direct_group<N, G, GI>(load, store);
    else direct_dispatch<N, G, GI + 1>(g, load, store);
  }
}

template <int N_, int G_>
struct Direct {
  static constexpr int N = N_, G = G_;
  static constexpr int PITCH = N_ | 1;
  static constexpr int SLOTS = N_ * PITCH;          // one plane buffer (the plane kernel uses two)
};

}  // namespace fftreg
}  // namespace isdf
