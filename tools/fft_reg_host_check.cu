// CPU check of the register-FFT phases (no GPU): compiles fft_reg_core.cuh as host code, runs every phase thread by
// thread and compares with a naive DFT.   nvcc -std=c++17 -O1 -I fft-isdf-scratch_b200/csrc tools/fft_reg_host_check.cu
#include <vector>
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "fft_reg_core.cuh"
using namespace isdf;
using namespace isdf::fftreg;
typedef std::complex<double> cd;

static double worst = 0;
static std::vector<cd> naive(const std::vector<cd>& x, int n, int stride, int count, int lstride) {   // lines of length n
  std::vector<cd> y(x.size());
  for (int l = 0; l < count; ++l)
    for (int k = 0; k < n; ++k) {
      cd acc = 0;
      for (int j = 0; j < n; ++j) acc += x[l * lstride + j * stride] * std::polar(1.0, -2 * M_PI * ((long)j * k % n) / n);
      y[l * lstride + k * stride] = acc;
    }
  return y;
}
static double relerr(const std::vector<cd>& a, const std::vector<cd>& b) {
  double num = 0, den = 0;
  for (size_t i = 0; i < a.size(); ++i) { num += std::norm(a[i] - b[i]); den += std::norm(b[i]); }
  return std::sqrt(num / den);
}
template <int R> static void check_rdft() {
  cplx x[R]; std::vector<cd> in(R);
  for (int i = 0; i < R; ++i) { in[i] = cd(drand48() - .5, drand48() - .5); x[i] = make_double2(in[i].real(), in[i].imag()); }
  rdft<R>(x);
  std::vector<cd> ref = naive(in, R, 1, 1, R), got(R);
  for (int i = 0; i < R; ++i) got[i] = cd(x[i].x, x[i].y);
  double e = relerr(got, ref); worst = std::max(worst, e);
  printf("rdft<%d>: %.2e\n", R, e);
}
template <class AX, int THREADS> static void check_plane() {
  constexpr int N = AX::N;
  std::vector<cd> in(N * N), pre(N * N); std::vector<double> post(N * N);
  for (auto& v : in) v = cd(drand48() - .5, drand48() - .5);
  for (auto& v : pre) v = cd(drand48() - .5, drand48() - .5);
  for (auto& v : post) v = drand48();
  std::vector<cd> t(N * N);
  for (int i = 0; i < N * N; ++i) t[i] = in[i] * pre[i];
  t = naive(t, N, 1, N, N);          // z
  // y: lines indexed by z, stride N
  std::vector<cd> ref(N * N);
  for (int z = 0; z < N; ++z)
    for (int k = 0; k < N; ++k) {
      cd acc = 0;
      for (int y = 0; y < N; ++y) acc += t[y * N + z] * std::polar(1.0, -2 * M_PI * ((long)y * k % N) / N);
      ref[k * N + z] = acc * post[k * N + z];
    }
  std::vector<cplx> P(AX::SLOTS, make_double2(1e300, 1e300)), TW(N), data(N * N), prec(N * N);
  for (int m = 0; m < N; ++m) TW[m] = make_double2(std::cos(2 * M_PI * m / N), -std::sin(2 * M_PI * m / N));
  for (int i = 0; i < N * N; ++i) { data[i] = make_double2(in[i].real(), in[i].imag()); prec[i] = make_double2(pre[i].real(), pre[i].imag()); }
  for (int tid = 0; tid < THREADS; ++tid) plane_z1<AX, THREADS>(tid, [&](int i) { return c_mul(data[i], prec[i]); }, P.data(), TW.data());
  for (int tid = 0; tid < THREADS; ++tid) plane_z2<AX, THREADS>(tid, P.data());
  for (int tid = 0; tid < THREADS; ++tid) plane_y1<AX, THREADS>(tid, P.data(), TW.data());
  for (int tid = 0; tid < THREADS; ++tid) plane_y2<AX, THREADS>(tid, P.data(), [&](int i, cplx v) { data[i] = make_double2(v.x * post[i], v.y * post[i]); });
  std::vector<cd> got(N * N);
  for (int i = 0; i < N * N; ++i) got[i] = cd(data[i].x, data[i].y);
  double e = relerr(got, ref); worst = std::max(worst, e);
  printf("plane N=%d (%d x %d), %d threads: %.2e\n", N, AX::R1, AX::R2, THREADS, e);
}
template <class AX, int T, int THREADS> static void check_lines(int lcnt) {
  constexpr int N = AX::N;
  const long stride = 1000;   // lines l0.. of a run of 1000
  std::vector<cd> in(N * stride); std::vector<double> post(N * stride);
  for (auto& v : in) v = cd(drand48() - .5, drand48() - .5);
  for (auto& v : post) v = drand48();
  std::vector<cplx> S(N * T), TW(N), data(N * stride);
  for (int m = 0; m < N; ++m) TW[m] = make_double2(std::cos(2 * M_PI * m / N), -std::sin(2 * M_PI * m / N));
  for (size_t i = 0; i < in.size(); ++i) data[i] = make_double2(in[i].real(), in[i].imag());
  const int l0 = 64;
  for (int tid = 0; tid < THREADS; ++tid) lines_s1<AX, T, THREADS>(tid, [&](int x, int l) { return data[x * stride + l0 + l]; }, lcnt, S.data(), TW.data());
  for (int tid = 0; tid < THREADS; ++tid) lines_s2<AX, T, THREADS>(tid, S.data(), lcnt, [&](int k, int l, cplx v) { const long o = k * stride + l0 + l; data[o] = make_double2(v.x * post[o], v.y * post[o]); });
  double num = 0, den = 0;
  for (int l = 0; l < T + 2; ++l)
    for (int k = 0; k < N; ++k) {
      cd ref;
      if (l < lcnt) {
        cd acc = 0;
        for (int j = 0; j < N; ++j) acc += in[j * stride + l0 + l] * std::polar(1.0, -2 * M_PI * ((long)j * k % N) / N);
        ref = acc * post[k * stride + l0 + l];
      } else ref = in[k * stride + l0 + l];       // untouched
      cd got(data[k * stride + l0 + l].x, data[k * stride + l0 + l].y);
      num += std::norm(got - ref); den += std::norm(ref);
    }
  double e = std::sqrt(num / den); worst = std::max(worst, e);
  printf("lines N=%d T=%d lcnt=%d: %.2e\n", N, T, lcnt, e);
}
template <int N, int G> static void check_direct() {
  std::vector<cd> in(N), got(N, cd(1e300, 0));
  for (auto& v : in) v = cd(drand48() - .5, drand48() - .5);
  for (int g = 0; g < G; ++g)
    direct_dispatch<N, G, 0>(g, [&](int j) { return make_double2(in[j].real(), in[j].imag()); },
                             [&](int k, cplx v) { got[k] = cd(v.x, v.y); });
  double e = relerr(got, naive(in, N, 1, 1, N)); worst = std::max(worst, e);
  printf("direct N=%d G=%d: %.2e\n", N, G, e);
}
#define ISDF_FFT_TWO(N, R1, R2, PT, PB, T, LT, LB) \
  (check_plane<TwoFactor<N, R1, R2>, PT>(), check_lines<TwoFactor<N, R1, R2>, T, LT>(T), check_lines<TwoFactor<N, R1, R2>, T, LT>(5), 0)
#define ISDF_FFT_DIRECT(N, G, PT, PB, T, LT, LB) (check_direct<N, G>(), 0)
int main() {
  check_rdft<2>(); check_rdft<3>(); check_rdft<4>(); check_rdft<5>(); check_rdft<6>(); check_rdft<7>(); check_rdft<8>();
  check_rdft<9>(); check_rdft<10>(); check_rdft<11>(); check_rdft<12>(); check_rdft<13>(); check_rdft<15>(); check_rdft<16>();
  int dummy[] = {0,
#include "fft_reg_sizes_p0.inc"
#include "fft_reg_sizes_p1.inc"
#include "fft_reg_sizes_p2.inc"
#include "fft_reg_sizes_p3.inc"
#include "fft_reg_sizes_p4.inc"
#include "fft_reg_sizes_p5.inc"
  };
  (void)dummy;
  printf("worst %.2e\n", worst);
  return worst < 1e-13 ? 0 : 1;
}
