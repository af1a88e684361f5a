// Batched 3-D complex128 FFT, register-resident kernels (replaces pbctools.fft / get_coulG / ifft at
// /root/reference/fftisdf.py:113-119, phase :99 and weights :114-115 fused; the inverse FFT is removed analytically,
// see DESIGN.md).  Two passes per batch, each reading and writing every element exactly once:
//
//   fftreg_plane_kernel : z + y of one x-plane per CTA round.  Two-factor lengths go global -> registers -> R1-point
//                         DFT -> twiddle -> shared (the one exchange of the axis) -> R2-point DFT; the plane buffer is
//                         transformed IN PLACE (a butterfly overwrites exactly the slots it read), so a 64 x 64 plane
//                         needs 75 KB and two CTAs share an SM; 3 exchanges x 32 B of shared-memory traffic per point.
//   fftreg_lines_kernel : the x pass on tiles of T consecutive (y,z) lines (coalesced 16-byte accesses, one exchange),
//                         Coulomb weight fused into the store.
//   Prime lengths (direct symmetric DFT) stage the plane / tile in shared memory first.
// Axis lengths are compile-time (fft_reg_core.cuh); isdf_fft3d_reg_supported() tells the host which meshes have an
// instantiation -- the others keep the generic Stockham / tensor-core DFT kernels.
#include "fft_reg_kernels.cuh"

namespace isdf {
namespace fftreg {

#define ISDF_FFT_PART(k) RegPlanSlice fft_reg_slice##k();
#include "fft_reg_parts.inc"
#undef ISDF_FFT_PART

static const RegPlan* find_plan(int n) {
  const RegPlanSlice sl[] = {
#define ISDF_FFT_PART(k) fft_reg_slice##k(),
#include "fft_reg_parts.inc"
#undef ISDF_FFT_PART
  };
  for (const RegPlanSlice& s : sl)
    for (int i = 0; i < s.count; ++i)
      if (s.plans[i].n == n) return &s.plans[i];
  return nullptr;
}

// ticket / publication counters of the fused launch (device) and its sticky time-out flag (host-mapped), per device
constexpr int FUSED_MAX_GROUPS = 1 << 16;
struct FusedSync { int* sync; int* err_host; int* err_dev; };
static FusedSync g_fsync[8];

static int fused_sync(Handle* h, FusedSync** out) {
  if (h->device < 0 || h->device >= 8) return ISDF_ESIZE;
  FusedSync& f = g_fsync[h->device];
  if (!f.sync) {
    ISDF_CUDA(h, cudaMalloc(&f.sync, sizeof(int) * (FUSED_MAX_GROUPS + 1)));
    ISDF_CUDA(h, cudaHostAlloc(&f.err_host, sizeof(int), cudaHostAllocMapped));
    *f.err_host = 0;
    ISDF_CUDA(h, cudaHostGetDevicePointer(&f.err_dev, f.err_host, 0));
  }
  *out = &f;
  return ISDF_OK;
}

static cplx* g_tw[8][512];     // per device, per length: exp(-2 pi i m / n) tables (a few KB in total, never freed)

static int twiddles(Handle* h, int n, const cplx** out) {
  if (h->device < 0 || h->device >= 8 || n >= 512) return ISDF_ESIZE;
  cplx*& t = g_tw[h->device][n];
  if (!t) {
    cplx* host = (cplx*)malloc(sizeof(cplx) * n);
    for (int j = 0; j < n; ++j) {
      const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
      host[j] = make_double2((double)cosl(ang), (double)sinl(ang));
    }
    cudaError_t e = cudaMalloc(&t, sizeof(cplx) * n);
    if (e == cudaSuccess) e = cudaMemcpy(t, host, sizeof(cplx) * n, cudaMemcpyHostToDevice);
    free(host);
    if (e != cudaSuccess) { t = nullptr; snprintf(h->err, sizeof(h->err), "fft twiddles: %s", cudaGetErrorString(e)); return (int)e; }
  }
  *out = t;
  return ISDF_OK;
}

}  // namespace fftreg
}  // namespace isdf

using namespace isdf;
using namespace isdf::fftreg;

// 1 when isdf_fft3d_reg has kernels for this mesh (n2 == n3 with an instantiated length; n1 == 1 or instantiated)
extern "C" int isdf_fft3d_reg_supported(const int* mesh) {
  if (!mesh) return 0;
  if (mesh[1] != mesh[2]) return 0;
  if (!find_plan(mesh[2])) return 0;
  if (mesh[0] != 1 && !find_plan(mesh[0])) return 0;
  return 1;
}

static int fft3d_reg_run(Handle* h, cplx* data, long nvec, long ldv, const int* mesh, const void* pre_dev,
                         const double* post_dev, long group_vecs, cplx* const* peers, int world, long ncol, long row0,
                         cudaStream_t st) {
  if (!isdf_fft3d_reg_supported(mesh)) {
    snprintf(h->err, sizeof(h->err), "isdf_fft3d_reg: no instantiation for mesh %d x %d x %d", mesh[0], mesh[1], mesh[2]);
    return ISDF_ESIZE;
  }
  const int n1 = mesh[0], n2 = mesh[1], n3 = mesh[2];
  const long ng = (long)n1 * n2 * n3;
  ISDF_CHECK_ARG(h, ldv >= ng, "ldv < prod(mesh)");
  ISDF_CHECK_ARG(h, ng < (1L << 31), "grid too large");
  if (nvec <= 0) return ISDF_OK;
  const bool p2p = peers != nullptr;
  ISDF_CHECK_ARG(h, !p2p || (world >= 1 && world <= 8 && ncol > 0 && ncol * world >= ng && n1 > 1), "peer layout");
  const RegPlan* pz = find_plan(n3);
  const RegPlan* px = (n1 > 1) ? find_plan(n1) : nullptr;
  const cplx *twz = nullptr, *twx = nullptr;
  int rc = twiddles(h, n3, &twz);
  if (rc) return rc;
  if (px) { rc = twiddles(h, n1, &twx); if (rc) return rc; }
  PeerArgs pr;
  for (int r = 0; r < 8; ++r) pr.peer[r] = (p2p && r < world) ? peers[r] : nullptr;
  pr.ncol = p2p ? ncol : 1; pr.row0 = row0;
  if (group_vecs == -2 && n1 == n2 && n1 == n3) {
    // cubic mesh, on request: both passes in one persistent launch, the intermediate stays in L2 (see
    // fftreg_fused_kernel).  Measured on B200 (tools/fft_tune.cu, profiles/README.md): no faster than two launches,
    // because the passes are bound on the SM side (same rate on L2-resident data), not by HBM -- kept selectable.
    FusedSync* fs;
    rc = fused_sync(h, &fs);
    if (rc) return rc;
    if (*fs->err_host) {
      snprintf(h->err, sizeof(h->err), "isdf_fft3d_reg: a dependency wait of an earlier fused launch timed out");
      return ISDF_ESIZE - 1;
    }
    const long target = 16L << 20;                       // bytes of intermediate per group
    long gv = target / (ng * (long)sizeof(cplx));
    if (gv < 1) gv = 1;
    if (gv > nvec) gv = nvec;
    long ngroups = (nvec + gv - 1) / gv;
    if (ngroups > FUSED_MAX_GROUPS) { gv = (nvec + FUSED_MAX_GROUPS - 1) / FUSED_MAX_GROUPS; ngroups = (nvec + gv - 1) / gv; }
    FusedArgs fa;
    fa.pl.data = data; fa.pl.ldv = ldv; fa.pl.n1 = n1; fa.pl.nwork = nvec * n1; fa.pl.tw = twz;
    fa.pl.pre = (const cplx*)pre_dev; fa.pl.post = nullptr; fa.pl.pr = pr;
    fa.ln.data = data; fa.ln.ldv = ldv; fa.ln.stride = (long)n2 * n3;
    fa.ln.tiles = (int)((fa.ln.stride + px->T - 1) / px->T); fa.ln.nwork = nvec * fa.ln.tiles; fa.ln.tw = twx;
    fa.ln.post = post_dev; fa.ln.pr = pr;
    fa.sync = fs->sync; fa.err = fs->err_dev; fa.nvec = nvec; fa.gv = (int)gv; fa.ngroups = (int)ngroups;
    fa.np = (int)(gv * n1); fa.nx = (int)(gv * fa.ln.tiles);
    ISDF_CUDA(h, cudaMemsetAsync(fs->sync, 0, sizeof(int) * (size_t)(ngroups + 1), st));
    return pz->fused(h, fa, p2p, st);
  }
  if (group_vecs <= 0) group_vecs = nvec;
  for (long v0 = 0; v0 < nvec; v0 += group_vecs) {
    const long nv = (nvec - v0 < group_vecs) ? (nvec - v0) : group_vecs;
    PlaneArgs pa;
    pa.data = data + v0 * ldv; pa.ldv = ldv; pa.n1 = n1; pa.nwork = nv * n1; pa.tw = twz;
    pa.pre = (const cplx*)pre_dev; pa.post = (n1 == 1) ? post_dev : nullptr;
    pa.pr = pr; pa.pr.row0 = row0 + v0;
    rc = pz->plane(h, pa, p2p, st);
    if (rc) return rc;
    if (px) {
      LinesArgs la;
      la.data = pa.data; la.ldv = ldv; la.stride = (long)n2 * n3;
      la.tiles = (int)((la.stride + px->T - 1) / px->T); la.nwork = nv * la.tiles; la.tw = twx; la.post = post_dev;
      la.pr = pa.pr;
      rc = px->lines(h, la, p2p, st);
      if (rc) return rc;
    }
  }
  return ISDF_OK;
}

// Same contract as isdf_fft3d_batched: data [nvec][ldv >= ng] c128 in place,
// out[v][G] = post[G] * sum_r data[v][r] * pre[r] * e^{-i G.r}.  group_vecs <= 0: one launch per pass over the whole
// batch (default); group_vecs > 0: two launches per group of that many vectors; group_vecs == -2 on a cubic mesh: both
// passes in one persistent launch with the intermediate held in L2 (dependency counters, fftreg_fused_kernel).
extern "C" int isdf_fft3d_reg(void* hv, void* data, long nvec, long ldv, const int* mesh, const void* pre_dev,
                              const double* post_dev, long group_vecs, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, data && mesh, "null pointer");
  return fft3d_reg_run(h, (cplx*)data, nvec, ldv, mesh, pre_dev, post_dev, group_vecs, nullptr, 1, 0, 0,
                       (cudaStream_t)stream);
}

// Multi-GPU form, same contract as isdf_dft3d_dmma_p2p: gather from / scatter to the ranks' grid-column shards
// peers[r] ([rows][ncol], NVLink peer-mapped) inside the transform; `work` [nvec][ldv >= ng] holds the intermediate.
extern "C" int isdf_fft3d_reg_p2p(void* hv, void* const* peers, int world, long ncol, long row0, void* work, long nvec,
                                  long ldv, const int* mesh, const void* pre_dev, const double* post_dev, void* stream) {
  Handle* h = (Handle*)hv;
  ISDF_CHECK_ARG(h, peers && work && mesh, "null pointer");
  return fft3d_reg_run(h, (cplx*)work, nvec, ldv, mesh, pre_dev, post_dev, -1, (cplx* const*)peers, world, ncol, row0,
                       (cudaStream_t)stream);
}
