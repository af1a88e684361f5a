// C-ABI entry points built on the DMMA GEMM engine, plus the handle and small data-movement kernels.
#include "gemm_c128.cuh"

namespace isdf {

__global__ void conj_copy_kernel(const cplx* __restrict__ src, cplx* __restrict__ dst, long n) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const cplx v = src[i];
    dst[i] = make_double2(v.x, -v.y);
  }
}

// dst[z][i][:] = src[z][idx[z][i]][:]  (zeros when idx < 0)
__global__ void gather_rows_kernel(const cplx* __restrict__ src, long lds, long strideS, const int* __restrict__ idx,
                                   long strideI, int nrows, long ncols, cplx* __restrict__ dst, long ldd, long strideD) {
  const int z = blockIdx.z;
  const int i = blockIdx.y;
  const int r = idx[(long)z * strideI + i];
  const cplx* s = src + (long)z * strideS + (long)r * lds;
  cplx* d = dst + (long)z * strideD + (long)i * ldd;
  for (long c = (long)blockIdx.x * blockDim.x + threadIdx.x; c < ncols; c += (long)gridDim.x * blockDim.x)
    d[c] = (r >= 0) ? s[c] : make_double2(0.0, 0.0);
}

// W[z][perm[r]][perm[c]] = alpha * sum_ks part[ks][z][r][c] for r >= c, mirrored (deterministic split-K reduce)
__global__ void herk_splitk_reduce_kernel(const cplx* __restrict__ part, int ksplit, long strideSplit, int n,
                                          double alpha, const int* __restrict__ perm, long stridePerm,
                                          cplx* __restrict__ w, long ldw, long strideW) {
  const int z = blockIdx.z;
  const int r = blockIdx.y * 16 + threadIdx.y;
  const int c = blockIdx.x * 16 + threadIdx.x;
  if (r >= n || c > r) return;
  double sr = 0.0, si = 0.0;
  for (int ks = 0; ks < ksplit; ++ks) {
    const cplx v = part[(long)ks * strideSplit + ((long)z * n + r) * n + c];
    sr += v.x; si += v.y;
  }
  const int* pm = perm ? perm + (long)z * stridePerm : nullptr;
  const int pr = pm ? pm[r] : r, pc = pm ? pm[c] : c;
  cplx* W = w + (long)z * strideW;
  W[(long)pr * ldw + pc] = make_double2(alpha * sr, (r == c) ? 0.0 : alpha * si);
  if (r > c) W[(long)pc * ldw + pr] = make_double2(alpha * sr, -alpha * si);
}

// out[i] = scale * sum_z sum_n y[z][i][n] * conj(x[z][i][n])          (fftisdf.py:155-156, rho_I)
__global__ void rowdot_conj_sum_kernel(const cplx* __restrict__ y, const cplx* __restrict__ x, int nz, int nrows,
                                       int ncols, double scale, cplx* __restrict__ out) {
  const int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= nrows) return;
  double ar = 0.0, ai = 0.0;
  for (int z = 0; z < nz; ++z) {
    const cplx* yr = y + ((long)z * nrows + i) * ncols;
    const cplx* xr = x + ((long)z * nrows + i) * ncols;
    for (int n = lane; n < ncols; n += 32) {
      const cplx a = yr[n], b = xr[n];
      ar += a.x * b.x + a.y * b.y;
      ai += a.y * b.x - a.x * b.y;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ar += __shfl_xor_sync(0xffffffffu, ar, o);
    ai += __shfl_xor_sync(0xffffffffu, ai, o);
  }
  if (lane == 0) out[i] = make_double2(scale * ar, scale * ai);
}

// out[z][i][n] = v[i] * x[z][i][n]                                       (fftisdf.py:166, diag(v) X_k)
__global__ void scale_rows_kernel(const cplx* __restrict__ x, const cplx* __restrict__ v, int nrows, int ncols,
                                  long total, cplx* __restrict__ out) {
  const long stride = (long)gridDim.x * blockDim.x;
  for (long w = (long)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += stride) {
    const int i = (int)((w / ncols) % nrows);
    out[w] = cmul(v[i], x[w]);
  }
}

}  // namespace isdf

using namespace isdf;

extern "C" int isdf_abi_version(void) { return 1; }

extern "C" int isdf_create(int device, void** out) {
  if (!out) return ISDF_EARG;
  *out = nullptr;
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return (int)e;
  Handle* h = new Handle();
  h->device = device;
  h->err[0] = 0;
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { delete h; return (int)e; }
  if (prop.major != 10) {
    // built for sm_100a only: refuse anything else loudly (no fallback path exists)
    delete h;
    return ISDF_ESIZE;
  }
  h->scratch = nullptr; h->scratch_bytes = 0;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  *out = h;
  return ISDF_OK;
}

extern "C" int isdf_fft_release_plans(void* hv);
extern "C" void dft_release_plans_internal(int device);

extern "C" int isdf_destroy(void* hv) {
  if (!hv) return ISDF_OK;
  isdf_fft_release_plans(hv);
  dft_release_plans_internal(((Handle*)hv)->device);
  if (((Handle*)hv)->scratch) cudaFree(((Handle*)hv)->scratch);
  delete (Handle*)hv;
  return ISDF_OK;
}

extern "C" const char* isdf_last_error(void* hv) {
  return hv ? ((Handle*)hv)->err : "null handle";
}

// x4c[g][h] = ( (sum_k Re(conj(x0[k,g,:]) . x0[k,h,:]))^2 / nk , 0 )      fftisdf.py:376-379
extern "C" int isdf_select_gram(void* hv, const void* x0, int nk, int n0, int nao, void* x4c, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, x0 && x4c, "null pointer");
  ISDF_CHECK_ARG(h, nk >= 1 && n0 >= 1 && nao >= 1, "shape");
  GemmParams p;
  p.A = (const cplx*)x0; p.lda = nao; p.strideA = 0;
  p.B = (const cplx*)x0; p.ldb = nao; p.strideB = 0;
  p.C = (cplx*)x4c; p.ldc = n0; p.strideC = 0;
  p.M = n0; p.N = n0; p.K = nao;
  p.nseg = nk; p.segA = (long)n0 * nao; p.segB = (long)n0 * nao;
  p.alpha = 1.0 / (double)nk;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJA, true, EPI_SQ_SYM>(p, 1, (cudaStream_t)stream)));
  return ISDF_OK;
}

// c[z][i][j] = sum_l conj(a[z][i][l]) * b[z][j][l]       fftisdf.py:38 (x2_k), :76 (fx_k)
extern "C" int isdf_gram_conja(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                               void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && b && c, "null pointer");
  ISDF_CHECK_ARG(h, m >= 0 && n >= 0 && k >= 0 && batch >= 0 && batch <= 65535, "shape");
  ISDF_CHECK_ARG(h, (m + 127) / 128 <= 65535, "m too large for one launch");
  GemmParams p;
  p.A = (const cplx*)a; p.lda = lda; p.strideA = strideA;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = (cplx*)c; p.ldc = ldc; p.strideC = strideC;
  p.M = m; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJA, false, EPI_STORE>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

// c[z][i][j] = sum_l a[z][i][l] * conj(b[z][j][l])      (fx_k^T = X_k F_k^H: the transposed form of fftisdf.py:76)
extern "C" int isdf_gram_conjb(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                               void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && b && c, "null pointer");
  ISDF_CHECK_ARG(h, m >= 0 && n >= 0 && k >= 0 && batch >= 0 && batch <= 65535, "shape");
  ISDF_CHECK_ARG(h, (m + 127) / 128 <= 65535, "m too large for one launch");
  GemmParams p;
  p.A = (const cplx*)a; p.lda = lda; p.strideA = strideA;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = (cplx*)c; p.ldc = ldc; p.strideC = strideC;
  p.M = m; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJB, false, EPI_STORE>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

// w[z][perm[i]][perm[j]] = alpha * sum_g b[z][i][g] conj(b[z][j][g])   (Hermitian; lower tiles + mirror)
// fftisdf.py:121 in Parseval form: W_q = B B^H.
extern "C" int isdf_herk_scatter(void* hv, const void* b, long ldb, long strideB, int n, int k, double alpha,
                                 const int* perm, long stridePerm, void* w, long ldw, long strideW, int batch,
                                 void* stream) {
  Handle* h = (Handle*)hv;
  cudaStream_t st = (cudaStream_t)stream;
  ISDF_CHECK_ARG(h, b && w, "null pointer");
  ISDF_CHECK_ARG(h, n >= 0 && k >= 0 && batch >= 0 && batch <= 65535, "shape");
  if (n == 0 || batch == 0) return ISDF_OK;   // nothing to write (e.g. every rank is zero)
  GemmParams p;
  p.A = (const cplx*)b; p.lda = ldb; p.strideA = strideB;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = (cplx*)w; p.ldc = ldw; p.strideC = strideW;
  p.M = n; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = alpha;
  p.perm = perm; p.stridePerm = stridePerm; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  // Small n, long K: the lower-tile grid alone does not fill the GPU (2 CTAs per SM), so the contraction is split
  // along K into partial products that a second kernel sums in a fixed order (deterministic, exactly Hermitian).
  const long nt = (n + 63) / 64;
  const long tiles = nt * (nt + 1) / 2 * batch;
  const long slots = 2L * h->sm_count;
  int ksplit = 1;
  if (tiles < 4 * slots && k >= 4096) {
    ksplit = (int)((4 * slots + tiles - 1) / tiles);
    if (ksplit > 16) ksplit = 16;
    while (ksplit > 1 && k / ksplit < 1024) --ksplit;
  }
  if (ksplit <= 1) {
    ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJB, false, EPI_HERK>(p, batch, st)));
    return ISDF_OK;
  }
  const size_t need = (size_t)ksplit * batch * n * n * sizeof(cplx);
  if (h->scratch_bytes < need) {
    if (h->scratch) ISDF_CUDA(h, cudaFree(h->scratch));
    h->scratch = nullptr; h->scratch_bytes = 0;
    ISDF_CUDA(h, cudaMalloc(&h->scratch, need));
    h->scratch_bytes = need;
  }
  p.C = (cplx*)h->scratch; p.ldc = n; p.strideC = (long)n * n;
  p.perm = nullptr; p.stridePerm = 0; p.alpha = 1.0;
  p.ksplit = ksplit; p.kchunk = ((k + ksplit - 1) / ksplit + 31) / 32 * 32; p.strideSplit = (long)batch * n * n;
  ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJB, false, EPI_HERK>(p, batch, st)));
  dim3 grid((n + 15) / 16, (n + 15) / 16, batch), block(16, 16);
  herk_splitk_reduce_kernel<<<grid, block, 0, st>>>((const cplx*)h->scratch, ksplit, p.strideSplit, n, alpha, perm,
                                                    stridePerm, (cplx*)w, ldw, strideW);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

// W~ partial products across GPUs, reduce-scatter fused into the HERK: batch z is stored (lower triangle, plain
// stores as the tiles finish, i.e. overlapped with the tensor work) into dst[z], which the host points at this rank's
// slab inside the OWNING rank's NVLink peer-mapped buffer.  isdf_sum_slabs_herm then sums the slabs in rank order.
extern "C" int isdf_herk_to_peers(void* hv, const void* b, long ldb, long strideB, int n, int k, double alpha,
                                  void* const* dst_dev, long ldw, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, b && dst_dev, "null pointer");
  ISDF_CHECK_ARG(h, n >= 0 && k >= 0 && batch >= 0 && batch <= 65535 && ldw >= n, "shape");
  if (n == 0 || batch == 0) return ISDF_OK;
  GemmParams p;
  p.A = (const cplx*)b; p.lda = ldb; p.strideA = strideB;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = nullptr; p.ldc = ldw; p.strideC = 0;
  p.M = n; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = alpha;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  p.Cbatch = (cplx* const*)dst_dev; p.lower_only = 1;
  ISDF_CUDA(h, (launch_gemm<128, 64, false, false, MODE_CONJB, false, EPI_HERK>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

// out[z][r][c] = sum_{w < world} slabs[z][w][r][c] for r >= c (fixed order: deterministic), mirrored conjugate above,
// exact real diagonal.  slabs [batch][world][n][ld] (only the leading n x n lower triangles are read), out [batch][n][ldo].
__global__ void sum_slabs_herm_kernel(const cplx* __restrict__ slabs, int world, int n, long ld, long slab_stride,
                                      cplx* __restrict__ out, long ldo, long strideO) {
  const int z = blockIdx.z;
  const int r = blockIdx.y * blockDim.y + threadIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n || c > r) return;
  const cplx* s = slabs + (long)z * world * slab_stride + (long)r * ld + c;
  double re = 0.0, im = 0.0;
  for (int w = 0; w < world; ++w) { const cplx v = s[(long)w * slab_stride]; re += v.x; im += v.y; }
  cplx* o = out + (long)z * strideO;
  o[(long)r * ldo + c] = make_double2(re, (r == c) ? 0.0 : im);
  if (r > c) o[(long)c * ldo + r] = make_double2(re, -im);
}

extern "C" int isdf_sum_slabs_herm(void* hv, const void* slabs, int world, int n, long ld, long slab_stride, void* out,
                                   long ldo, long strideO, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, slabs && out, "null pointer");
  ISDF_CHECK_ARG(h, world >= 1 && n >= 0 && ld >= n && ldo >= n && batch >= 0 && batch <= 65535, "shape");
  if (n == 0 || batch == 0) return ISDF_OK;
  dim3 block(32, 8), grid((n + 31) / 32, (n + 7) / 8, batch);
  sum_slabs_herm_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const cplx*)slabs, world, n, ld, slab_stride,
                                                                 (cplx*)out, ldo, strideO);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

// c[z] = a[z] * b[z]  with a [m][k] row-major, b [k][n] row-major (plain complex GEMM, NN)
extern "C" int isdf_gemm_nn(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
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
  p.group_m = (m >= 512 && (long)n * k * 16 > (64L << 20)) ? 16 : 0;   // B panel larger than L2's share: rasterise
  ISDF_CUDA(h, (launch_gemm<64, 128, false, true, MODE_AB, false, EPI_STORE>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

// c[z][i][j] = sum_l conj(a[z][l][i]) * b[z][l][j]   (A^H B; a [k][m], b [k][n] row-major)   fftisdf.py:166, :225
extern "C" int isdf_gemm_hn(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
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
  ISDF_CUDA(h, (launch_gemm<128, 64, true, true, MODE_CONJA, false, EPI_STORE>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

// C = A^H B for a product KNOWN to be Hermitian (W = E^H (W~ E)): only the tiles on and below the diagonal are
// computed, the upper triangle is the mirrored conjugate and the diagonal gets an exact zero imaginary part
// (the HERK epilogue), i.e. gemm_hn + hermitize at ~half the tensor work.
extern "C" int isdf_gemm_hn_herm(void* hv, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                                 void* c, long ldc, long strideC, int n, int k, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, a && b && c, "null pointer");
  ISDF_CHECK_ARG(h, n >= 0 && k >= 0 && batch >= 0 && batch <= 65535, "shape");
  GemmParams p;
  p.A = (const cplx*)a; p.lda = lda; p.strideA = strideA;
  p.B = (const cplx*)b; p.ldb = ldb; p.strideB = strideB;
  p.C = (cplx*)c; p.ldc = ldc; p.strideC = strideC;
  p.M = n; p.N = n; p.K = k;
  p.nseg = 1; p.segA = 0; p.segB = 0; p.alpha = 1.0;
  p.perm = nullptr; p.stridePerm = 0; p.active = nullptr; p.ksplit = 1; p.kchunk = 0; p.strideSplit = 0;
  ISDF_CUDA(h, (launch_gemm<128, 64, true, true, MODE_CONJA, false, EPI_HERK>(p, batch, (cudaStream_t)stream)));
  return ISDF_OK;
}

extern "C" int isdf_rowdot_conj_sum(void* hv, const void* y, const void* x, int nz, int nrows, int ncols, double scale,
                                    void* out, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, y && x && out && nz >= 1 && nrows >= 1 && ncols >= 1, "args");
  rowdot_conj_sum_kernel<<<(nrows + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const cplx*)y, (const cplx*)x, nz, nrows,
                                                                         ncols, scale, (cplx*)out);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_scale_rows(void* hv, const void* x, const void* v, int nz, int nrows, int ncols, void* out,
                               void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, x && v && out && nz >= 1 && nrows >= 1 && ncols >= 1, "args");
  const long total = (long)nz * nrows * ncols;
  long blocks = (total + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  scale_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)x, (const cplx*)v, nrows, ncols,
                                                                        total, (cplx*)out);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_conj_copy(void* hv, const void* src, void* dst, long n, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, src && dst && n >= 0, "args");
  if (n == 0) return ISDF_OK;
  long blocks = (n + 255) / 256;
  if (blocks > 148L * 16) blocks = 148L * 16;
  conj_copy_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const cplx*)src, (cplx*)dst, n);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}

extern "C" int isdf_gather_rows(void* hv, const void* src, long lds, long strideS, const int* idx, long strideI,
                                int nrows, long ncols, void* dst, long ldd, long strideD, int batch, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, src && dst && idx, "null pointer");
  ISDF_CHECK_ARG(h, nrows >= 0 && nrows <= 65535 && batch >= 0 && batch <= 65535 && ncols >= 0, "shape");
  if (nrows == 0 || ncols == 0 || batch == 0) return ISDF_OK;
  long bx = (ncols + 255) / 256;
  if (bx > 64) bx = 64;
  dim3 grid((unsigned)bx, nrows, batch);
  gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const cplx*)src, lds, strideS, idx, strideI, nrows, ncols,
                                                             (cplx*)dst, ldd, strideD);
  ISDF_LAUNCH_CHECK(h);
  return ISDF_OK;
}
