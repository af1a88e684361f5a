// Kernel-tuning harness for the register-resident FFT (no Python): times the plane pass, the x pass and grouped
// runs of several thread / residency configurations per axis length.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fft-isdf-scratch_b200/csrc tools/fft_tune.cu -o tools/bin/fft_tune
#include <vector>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <cmath>
#include <functional>
#include "fft_reg_kernels.cuh"
using namespace isdf;
using namespace isdf::fftreg;
static float time_ms(const std::function<void()>& f, int reps);

static Handle H;
static cplx* g_data; static size_t g_bytes;
static cplx* g_tw;

static int* g_sync; static int* g_err_host; static int* g_err_dev;
template <class POP, class LOP, int PT, int PB>
static float time_fused(PlaneArgs pa, LinesArgs la, long nvec, long ng, int n1, long target_bytes) {
  long gv = target_bytes / (ng * 16); if (gv < 1) gv = 1; if (gv > nvec) gv = nvec;
  FusedArgs fa; fa.pl = pa; fa.ln = la; fa.sync = g_sync; fa.err = g_err_dev; fa.nvec = nvec; fa.gv = (int)gv;
  fa.ngroups = (int)((nvec + gv - 1) / gv); fa.np = (int)(gv * n1); fa.nx = (int)(gv * la.tiles);
  return time_ms([&] {
    cudaMemsetAsync(g_sync, 0, sizeof(int) * (fa.ngroups + 1), 0);
    launch_fused<POP, LOP, PT, PB>(&H, fa, false, 0);
  });
}
static float time_ms(const std::function<void()>& f, int reps = 3);
static float time_ms(const std::function<void()>& f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s (%s)\n", cudaGetErrorString(e), H.err);
  return ms / reps;
}

template <class AX, int PT, int PB, int T, int LT, int LB>
static void run(const char* tag, const char* only, long gvec_a = 0, long gvec_b = 0, long gvec_c = 0) {
  constexpr int N = AX::N;
  if (only && atoi(only) != N) return;
  const long ng = (long)N * N * N;
  const long nvec = (long)(g_bytes / (ng * sizeof(cplx)));
  std::vector<cplx> tw(N);
  for (int j = 0; j < N; ++j) tw[j] = make_double2(cos(-2 * M_PI * j / N), sin(-2 * M_PI * j / N));
  cudaMemcpy(g_tw, tw.data(), sizeof(cplx) * N, cudaMemcpyHostToDevice);
  PlaneArgs pa; memset(&pa, 0, sizeof(pa));
  pa.data = g_data; pa.ldv = ng; pa.n1 = N; pa.nwork = nvec * N; pa.tw = g_tw; pa.pr.ncol = 1;
  LinesArgs la; memset(&la, 0, sizeof(la));
  la.data = g_data; la.ldv = ng; la.stride = (long)N * N; la.tiles = (int)((la.stride + T - 1) / T); la.nwork = nvec * la.tiles;
  la.tw = g_tw; la.pr.ncol = 1;
  const double gb = 2.0 * 16 * nvec * ng / 1e9;
  float tp = time_ms([&] { launch_pass<PlaneTwo<AX, PT>, PlaneArgs, PT, PB>(&H, pa, false, 0); });
  float tl = time_ms([&] { launch_pass<LinesTwo<AX, T, LT>, LinesArgs, LT, LB>(&H, la, false, 0); });
  printf("N=%3d %-10s plane %dx%d: %7.3f ms %6.0f GB/s | x T=%d %dx%d: %7.3f ms %6.0f GB/s | both %6.0f GB/s", N, tag, PT, PB, tp,
         gb / tp * 1e3, T, LT, LB, tl, gb / tl * 1e3, gb / (tp + tl) * 1e3);
  for (long gv : {gvec_a, gvec_b, gvec_c}) {
    if (gv <= 0) continue;
    float tg = time_ms([&] {
      for (long v0 = 0; v0 < nvec; v0 += gv) {
        const long nv = (nvec - v0 < gv) ? nvec - v0 : gv;
        PlaneArgs p2 = pa; p2.data = g_data + v0 * ng; p2.nwork = nv * N;
        LinesArgs l2 = la; l2.data = p2.data; l2.nwork = nv * la.tiles;
        launch_pass<PlaneTwo<AX, PT>, PlaneArgs, PT, PB>(&H, p2, false, 0);
        launch_pass<LinesTwo<AX, T, LT>, LinesArgs, LT, LB>(&H, l2, false, 0);
      }
    });
    printf(" | g%ld %6.0f", gv, gb / tg * 1e3);
  }
  for (long mb : {8L, 16L, 32L}) {
    float tf = time_fused<PlaneTwo<AX, PT>, LinesTwo<AX, T, PT>, PT, PB>(pa, la, nvec, ng, N, mb << 20);
    printf(" | F%ldMB %6.0f", mb, gb / tf * 1e3);
  }
  printf(" err=%d\n", *g_err_host); fflush(stdout);
}

template <class AX, int PT, int PB, int T, int LT, int LB>
static void run_direct(const char* tag, const char* only) {
  constexpr int N = AX::N;
  if (only && atoi(only) != N) return;
  const long ng = (long)N * N * N;
  const long nvec = (long)(g_bytes / (ng * sizeof(cplx)));
  PlaneArgs pa; memset(&pa, 0, sizeof(pa));
  pa.data = g_data; pa.ldv = ng; pa.n1 = N; pa.nwork = nvec * N; pa.tw = g_tw; pa.pr.ncol = 1;
  LinesArgs la; memset(&la, 0, sizeof(la));
  la.data = g_data; la.ldv = ng; la.stride = (long)N * N; la.tiles = (int)((la.stride + T - 1) / T); la.nwork = nvec * la.tiles;
  la.tw = g_tw; la.pr.ncol = 1;
  const double gb = 2.0 * 16 * nvec * ng / 1e9;
  float tp = time_ms([&] { launch_pass<PlaneDirect<AX, PT>, PlaneArgs, PT, PB>(&H, pa, false, 0); });
  float tl = time_ms([&] { launch_pass<LinesDirect<AX, T, LT>, LinesArgs, LT, LB>(&H, la, false, 0); });
  printf("N=%3d %-10s plane %dx%d G%d: %7.3f ms %6.0f GB/s | x T=%d %dx%d: %7.3f ms %6.0f GB/s | both %6.0f GB/s", N, tag, PT, PB,
         AX::G, tp, gb / tp * 1e3, T, LT, LB, tl, gb / tl * 1e3, gb / (tp + tl) * 1e3);
  for (long mb : {8L, 16L, 32L}) {
    float tf = time_fused<PlaneDirect<AX, PT>, LinesDirect<AX, T, PT>, PT, PB>(pa, la, nvec, ng, N, mb << 20);
    printf(" | F%ldMB %6.0f", mb, gb / tf * 1e3);
  }
  printf(" err=%d\n", *g_err_host);
  fflush(stdout);
}

int main(int argc, char** argv) {
  const char* only = argc > 1 ? argv[1] : nullptr;
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  H.device = 0; H.sm_count = prop.multiProcessorCount; H.max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  g_bytes = (size_t)(getenv("FFT_TUNE_MB") ? atoi(getenv("FFT_TUNE_MB")) : 1280) << 20;
  cudaMalloc(&g_data, g_bytes); cudaMemset(g_data, 0, g_bytes);
  cudaMalloc(&g_tw, 4096);
  cudaMalloc(&g_sync, 4 << 20);
  cudaHostAlloc(&g_err_host, 4, cudaHostAllocMapped); *g_err_host = 0; cudaHostGetDevicePointer(&g_err_dev, g_err_host, 0);
#include "fft_tune_cases.inc"
  return 0;
}
