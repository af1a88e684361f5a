/* libisdf_b200 -- C ABI of the B200-native FFT-ISDF build (sm_100a).
 *
 * This is the drop-in boundary for the hot path of yangjunjie0320/fft-isdf-scratch:
 * `build(df_obj)` (fftisdf.py:22-128) and `select_interpolation_points` (fftisdf.py:357-388).
 * The reference is pure Python; the native kernels it reaches through numpy/scipy/PySCF
 * (ZGEMM, dpstrf, zgelsy, FFT) are what the entry points below replace, one per call site.
 *
 * Conventions
 *   - every function returns int: 0 = ok, < 0 = argument error, > 0 = cudaError_t;
 *     isdf_last_error(handle) gives the message.  Nothing throws across the ABI.
 *   - all pointers are DEVICE pointers owned by the caller unless stated otherwise; complex128 is
 *     interleaved (re, im) doubles exactly as numpy / torch store it; matrices are row-major.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it.
 *   - one handle per (process, device); a handle is not thread-safe.  The handle owns only small
 *     plans (FFT twiddle tables, DFT matrices) and a grow-only scratch for split-K partial sums.
 *   - there is no CPU path: isdf_create fails on anything that is not compute capability 10.x.
 *
 * Stage map (SURVEY.md section 8b names the stages of the path; each is this sequence of entry points,
 * orchestrated by fft-isdf-scratch_b200/fftisdf.py:build, which also owns the multi-GPU exchanges):
 *   select points   fftisdf.py:357-388  isdf_select_gram -> isdf_pchol_real -> isdf_gather_rows
 *   build metric    fftisdf.py:38-48    isdf_gram_conja -> isdf_ktransform_square_rows (registers; k-mesh axes <= 4)
 *                                       or isdf_ktransform_square (shared memory; axes <= 8)
 *   build rhs       fftisdf.py:72-87    isdf_gram_conjb -> isdf_ktransform_square_rows, per grid block, rows written
 *                                       straight into pivot order
 *   fit theta       fftisdf.py:108      zgelsy on the device (default): isdf_qrcp -> isdf_gelsy_rank -> the dense operators
 *                                       G = U^-H D^-1 Q1^H and E^H per q (isdf_gelsy_extract / _rhat / _q1h_finish,
 *                                       isdf_chol_nopivot, isdf_herk_scatter, isdf_gemm_*) -> isdf_gemm_nn (Theta~ = G Y^T,
 *                                       per grid block);  fit = "cholesky": isdf_pchol -> isdf_trsm_prepare -> isdf_trsm_sweeps
 *   coulomb kernel  fftisdf.py:96-122   isdf_phase_table + isdf_coulomb_weights -> isdf_fft3d_reg (instantiated lengths; _p2p
 *                                       across GPUs), isdf_dft3d_dmma (other meshes with axes <= 48; _p2p) or
 *                                       isdf_fft3d_batched -> isdf_herk_scatter (isdf_herk_to_peers + isdf_sum_slabs_herm
 *                                       across GPUs) -> isdf_gemm_nn + isdf_gemm_hn_herm (W = E W~ E^H) -> isdf_conj_copy
 *   J / K           fftisdf.py:133-228  isdf_gemm_hn, isdf_rowdot_conj_sum, isdf_scale_rows, isdf_ktransform_rows_ex
 */
#ifndef ISDF_B200_H
#define ISDF_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

int isdf_abi_version(void);
int isdf_create(int device, void** handle_out);
int isdf_destroy(void* handle);
const char* isdf_last_error(void* handle);

/* fftisdf.py:376-379   x2 = sum_q Re(x0[q]^* x0[q]^T);  x4 = x2*x2/nkpt
 * x0 [nk][n0][nao] c128  ->  x4c [n0][n0] c128 with zero imaginary parts (exactly symmetric). */
int isdf_select_gram(void* handle, const void* x0, int nk, int n0, int nao, void* x4c, void* stream);

/* fftisdf.py:381-382  pyscf.lib.scipy_helper.pivoted_cholesky -> LAPACK dpstrf (upper), and the
 * rank-revealing factorisation used in place of the QRCP inside scipy lstsq(..., "gelsy") at :108.
 * a [batch][n][n] Hermitian PSD, lower triangle referenced (destroyed).  Runs at most max_steps pivots, stops earlier when the
 * pivot <= tol (tol < 0: n*eps*max diag, LAPACK's default).  nb = panel width (<= 64, 0 -> 32).
 * u [batch][ldu_rows][n]: row j = row j of U (A = U^H U) in ORIGINAL column order.
 * piv [batch][n] (0-based, position -> original index), rank [batch], next_pivot [batch] (value of the
 * pivot that would come next; sqrt of it is the reference's chol[nip,nip] at :387), may be NULL.
 * workspace: isdf_pchol_workspace_bytes(n, batch) bytes. */
int isdf_pchol_workspace_bytes(int n, int batch, size_t* bytes);
int isdf_pchol(void* handle, void* a, int n, int batch, int max_steps, double tol, int nb, void* u, int ldu_rows,
               int* piv, int* rank, double* next_pivot, void* workspace, void* stream);
/* Same for matrices whose imaginary parts are exactly zero (x4 of isdf_select_gram): identical arithmetic and
 * pivots, in-panel rows kept as doubles so that twice as wide a panel fits (half the trailing updates). */
int isdf_pchol_real(void* handle, void* a, int n, int batch, int max_steps, double tol, int nb, void* u, int ldu_rows,
                    int* piv, int* rank, double* next_pivot, void* workspace, void* stream);

/* fftisdf.py:38 (x2_k) and :76 (fx_k):  c[z][i][j] = sum_l conj(a[z][i][l]) * b[z][j][l]
 * a [m][k] (ld lda), b [n][k] (ld ldb), c [m][n] (ld ldc); batch strides in elements. */
int isdf_gram_conja(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                    void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream);

/* transposed form of :76:  c[z][i][j] = sum_l a[z][i][l] * conj(b[z][j][l])  (fx_k^T = X_k F_k^H). */
int isdf_gram_conjb(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                    void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream);

/* plain batched complex GEMM c[z] = a[z] b[z], a [m][k], b [k][n] (the kernel behind the sweeps). */
int isdf_gemm_nn(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                 void* c, long ldc, long strideC, int m, int n, int k, int batch, void* stream);

/* fftisdf.py:41-47 (metric, conj2=1) and :79-85 (right-hand side, conj2=0):
 *   s = phase @ v;  assert |Im s| small;  y = s*s;  out = phase^H @ y  or  phase^T @ y
 * in[k*in_sk + g*in_sg + i], g < ng, i < ni;  kmesh[3] (host) with every axis <= 8;
 * uaxes: device, 3 matrices [8][8] c128, U_a[m][j] = exp(2 pi i m j/N_a)/sqrt(N_a) (phase = U1 x U2 x U3);
 * out[slot*out_sq + (out_g0+g)*out_sg + row*out_si], slot = qslot[q] (NULL: q; <0: skip),
 * row = rowmap[slot*rowmap_sq + i] (NULL: i; <0: drop);  out_g_fast = 1 when out_sg == 1.
 * diag (device, 2 doubles, may be NULL): running max of |Im s| and |Re s| (atomic max). */
int isdf_ktransform_square(void* handle, const void* in, long in_sk, long in_sg, void* out, long out_sq, long out_sg,
                           long out_si, long out_g0, int ng, int ni, const int* kmesh, const void* uaxes, int conj2,
                           int out_g_fast, const int* qslot, const int* rowmap, long rowmap_sq, double* diag,
                           void* stream);

/* General form of isdf_ktransform_square (any k-mesh with axes <= 8), same modes as isdf_ktransform_rows_ex below:
 * mode 0 = square; mode 1 = multiply scale*Re(s) by the real R-space table[R*tab_sk + g*tab_sg + i] before the second
 * transform (fftisdf.py:215-223); mode 2 = write scale*Re(s) as a real table (doubles, strides in doubles) and stop
 * (fftisdf.py:205-207). */
int isdf_ktransform_ex(void* handle, const void* in, long in_sk, long in_sg, void* out, long out_sq, long out_sg,
                       long out_si, long out_g0, int ng, int ni, const int* kmesh, const void* uaxes, int conj2,
                       int out_g_fast, const int* qslot, const int* rowmap, long rowmap_sq, double* diag, int mode,
                       const double* table, long tab_sk, long tab_sg, double scale, void* stream);

/* Register-resident variant of isdf_ktransform_square for small k-meshes (every axis <= 4, nk <= 32):
 * in[k*in_sk + r*in_sr + c] (c contiguous) -> out[slot*out_sq + row*out_sr + out_c0 + c], row = rowmap[slot][r].
 * uaxes is a HOST pointer here (3 x [8][8] c128, copied to constant memory).  Returns -2 without launching
 * when the mesh has no instantiation (use isdf_ktransform_square then). */
int isdf_ktransform_square_rows(void* handle, const void* in, long in_sk, long in_sr, void* out, long out_sq,
                                long out_sr, long out_c0, int nrows, int ncols, const int* kmesh,
                                const void* uaxes_host, int conj2, const int* qslot, const int* rowmap,
                                long rowmap_sq, double* diag, void* stream);

/* General form of the register k-transform: mode 0 = square (as above); mode 1 = multiply scale*Re(s) by a real
 * R-space table[R*tab_sk + r*tab_sr + c] before the second transform (exchange: vs = ws * rhos^T, fftisdf.py:215-223);
 * mode 2 = write scale*Re(s) as a real table out[R*out_sq + r*out_sr + out_c0 + c] (doubles) and stop
 * (ws = Re(phase @ wq)*sqrt(nk), fftisdf.py:205-207). */
int isdf_ktransform_rows_ex(void* handle, const void* in, long in_sk, long in_sr, void* out, long out_sq, long out_sr,
                            long out_c0, int nrows, int ncols, const int* kmesh, const void* uaxes_host, int conj2,
                            const int* qslot, const int* rowmap, long rowmap_sq, double* diag, int mode,
                            const double* table, long tab_sk, long tab_sr, double scale, void* stream);

/* J/K consumers (fftisdf.py:133-228) building blocks:
 * isdf_gemm_hn: c[z] = a[z]^H b[z], a [k][m], b [k][n] row-major (X_k^H (...) at :166, :225);
 * isdf_rowdot_conj_sum: out[i] = scale * sum_z sum_n y[z][i][n] conj(x[z][i][n])  (rho_I, :155-156);
 * isdf_scale_rows: out[z][i][n] = v[i] * x[z][i][n]  (diag(v) X_k, :166). */
int isdf_gemm_hn(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB, void* c,
                 long ldc, long strideC, int m, int n, int k, int batch, void* stream);
int isdf_rowdot_conj_sum(void* handle, const void* y, const void* x, int nz, int nrows, int ncols, double scale,
                         void* out, void* stream);
int isdf_scale_rows(void* handle, const void* x, const void* v, int nz, int nrows, int ncols, void* out, void* stream);

/* fftisdf.py:108  scipy.linalg.lstsq(A_q, Y_q^T): block operators of the two triangular sweeps from the
 * pivoted factor (block size 64).  n = nip, nP = multiple of 64 >= max(rank) (usually n rounded up).  lfwd, ubwd [batch][nP][nP];
 * work 2*batch*nP*nP c128.  Rows/columns at positions >= rank are replaced by the identity. */
int isdf_trsm_prepare(void* handle, const void* u, int ldu_rows, const int* piv, const int* rank, int n, int nP,
                      int batch, void* lfwd, void* ubwd, void* work, void* stream);
/* In place T <- U^{-1} U^{-H} T on t [batch][nP][ldt] (ng columns used): Theta in pivot order. */
int isdf_trsm_sweeps(void* handle, const void* lfwd, const void* ubwd, void* t, int nP, int nact, long ng, long ldt, int batch,
                     void* stream);

/* One direction of the above: backward = 0: T <- U^{-H} T with op = lfwd; backward = 1: T <- U^{-1} T with op = ubwd. */
int isdf_trsm_sweep(void* handle, const void* op, void* t, int nP, int nact, long ng, long ldt, int batch,
                    int backward, void* stream);
/* Unpivoted blocked Cholesky A = U^H U (64-wide block columns: diagonal block in shared memory, block row and
 * trailing update on the GEMM engine); argument meaning as isdf_pchol with max_steps = n, piv comes back as the
 * identity.  Stops at the first pivot <= tol (rank < n).  workspace: isdf_pchol_workspace_bytes(n, batch) +
 * batch * 64 * 64 * 16 bytes. */
int isdf_chol_nopivot(void* handle, void* a, int n, int batch, int max_steps, double tol, int nb, void* u,
                      int ldu_rows, int* piv, int* rank, void* workspace, void* stream);

/* fftisdf.py:108  scipy.linalg.lstsq(x4_q, y_q.T, lapack_driver="gelsy") -> LAPACK ZGELSY, restated on the device.
 *
 * isdf_qrcp (zgeqp3): Householder QR with column pivoting, A P = Q R.  a [batch][n][n] is the COLUMN-major working
 * copy (a[c][i] = A[i][c]; for the Hermitian metric this is conj of the row-major matrix); on return
 * a[c][k] = R[k][c] for k <= pos[c].  vt [batch][n][n]: row k = Householder vector v_k (v_k[k] = 1, zeros before;
 * an identity reflector is the zero row).  tau [batch][n] c128 (H_k = I - tau_k v_k v_k^H, LAPACK's convention).
 * piv [batch][n] position -> column, pos [batch][n] column -> position (both 0-based, device).
 * isdf_gelsy_rank (zgelsy's loop over zlaic1): rank[b] = number of leading columns of R accepted by incremental
 * condition estimation, smax * rcond <= smin (the reference gets rcond = machine eps from scipy).  xwork: 2*batch*n
 * c128. */
int isdf_qrcp(void* handle, void* a, int n, int batch, void* vt, void* tau, int* piv, int* pos, void* stream);
int isdf_gelsy_rank(void* handle, const void* a, const int* piv, int n, int batch, double rcond, void* xwork,
                    int* rank, void* stream);
/* Ingredients of the three dense operators through which zgelsy's solution x = P Z^H [T11^-1 (Q^H b)(1:rank); 0]
 * (zunmqr, ztrsm, ztzrzf/zunmrz) is applied to all right-hand sides (sequenced by kernels.py:gelsy_factor):
 *   Q1 D^-1  = (I(:, :rank) - V S^-1 V(:rank, :)^H) D^-1,  S = diag(1/tau) + striu(V^H V)   (compact WY, D = |diag R|)
 *   E^H      = rows of D^-1 [R11 R12] P^T orthonormalised by Cholesky-QR (twice), U = their triangular factor
 * isdf_gelsy_extract: g [batch][rP][rP] = vt vt^H = (V^H V)^T (lower triangle read)  ->  s, v1h [batch][rP][rP] (S and V(:rank,:)^H), dinv [batch][rP].
 * isdf_gelsy_rhat: rhat [batch][rP][n] = D^-1 [R11 R12] P^T (original column order, zero rows beyond rank).
 * isdf_gelsy_q1_finish: vm [batch][n][rP] = V S^-1 V1^H on entry, Q1 D^-1 on return (zero columns beyond rank).
 * isdf_hermitize: w <- (w + w^H)/2.  isdf_gemm_tn: c = a^T b (no conjugation), a [k][m], b [k][n]. */
int isdf_gelsy_extract(void* handle, const void* g, const void* tau, const int* rank, const void* vt, const void* a,
                       const int* piv, int n, int rP, int batch, void* s, void* v1h, double* dinv, void* stream);
int isdf_gelsy_rhat(void* handle, const void* a, const int* pos, const double* dinv, const int* rank, int n, int rP,
                    int batch, void* rhat, void* stream);
int isdf_gelsy_q1_finish(void* handle, void* vm, const double* dinv, const int* rank, int n, int rP, int batch,
                         void* stream);
/* Transposed twin: t1 [batch][rP][n] = M^T V on entry, D^-1 Q1^H on return (zero rows beyond rank). */
int isdf_gelsy_q1h_finish(void* handle, void* t1, const double* dinv, const int* rank, int n, int rP, int batch,
                          void* stream);
int isdf_hermitize(void* handle, void* w, int n, int batch, void* stream);
/* Multi-GPU form of the W~ = B~ B~^H contraction (fftisdf.py:121) with the reduce-scatter over the ranks fused into
 * the product: batch z (a q-slot) is stored -- lower triangle, as its tiles finish -- into dst[z] (DEVICE array of
 * `batch` pointers), which the host points at this rank's slab [n][ldw] inside the slot OWNER's NVLink peer-mapped
 * buffer.  After a cross-rank barrier the owner calls isdf_sum_slabs_herm: out[z] = sum over the `world` slabs in
 * rank order (deterministic), mirrored, exact real diagonal.  slabs [batch][world][slab_stride], out [batch][n][ldo]. */
int isdf_herk_to_peers(void* handle, const void* b, long ldb, long strideB, int n, int k, double alpha,
                       void* const* dst_dev, long ldw, int batch, void* stream);
int isdf_sum_slabs_herm(void* handle, const void* slabs, int world, int n, long ld, long slab_stride, void* out,
                        long ldo, long strideO, int batch, void* stream);
/* c[z] [n,n] = a[z]^H b[z] for a product known to be Hermitian (W_q = E (W~ E^H), fftisdf.py:121 in the row space of
 * zgelsy): lower tiles only, mirrored conjugate above, exact real diagonal.  a, b: [k][n] per batch. */
int isdf_gemm_hn_herm(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB,
                      void* c, long ldc, long strideC, int n, int k, int batch, void* stream);
int isdf_gemm_tn(void* handle, const void* a, long lda, long strideA, const void* b, long ldb, long strideB, void* c,
                 long ldc, long strideC, int m, int n, int k, int batch, void* stream);

/* fftisdf.py:113-115  pbctools.fft(z_q * fq, mesh) * coulG * vol/ngrid  (the ifft at :118 is removed by
 * Parseval):  data [nvec][ldv >= ng] in place, out[v][G] = post[G] * sum_r data[v][r] pre[r] e^{-iG.r}.
 * mesh[3] host; pre [ng] c128 or NULL; post [ng] f64 or NULL; group_vecs <= 0 picks an L2-sized group. */
int isdf_fft3d_batched(void* handle, void* data, long nvec, long ldv, const int* mesh, const void* pre,
                       const double* post, long group_vecs, void* stream);
int isdf_fft_release_plans(void* handle);
/* Same contract, register-resident kernels with compile-time axis lengths (fft_reg.cu): two-factor lengths
 * N = R1 R2 (R1, R2 <= 16) and primes <= 61; a plane pass (z + y, in place in shared memory) and an x pass, each
 * touching every element once.  isdf_fft3d_reg_supported(mesh) -> 1 when mesh[1] == mesh[2] and the lengths are
 * instantiated; isdf_fft3d_reg returns -2 without launching otherwise (the host then keeps isdf_fft3d_batched /
 * isdf_dft3d_dmma).  group_vecs > 0 runs both passes per group of vectors, <= 0 per batch. */
int isdf_fft3d_reg_supported(const int* mesh);
int isdf_fft3d_reg(void* handle, void* data, long nvec, long ldv, const int* mesh, const void* pre,
                   const double* post, long group_vecs, void* stream);
/* Same contract as isdf_fft3d_batched for meshes with every axis in [2, 48]: each 1-D transform is a dense
 * product with the n x n DFT matrix on the FP64 tensor pipe (for lengths with large prime factors, e.g. 31,
 * 37, 41).  Returns -2 without launching when an axis is out of range. */
int isdf_dft3d_dmma(void* handle, void* data, long nvec, long ldv, const int* mesh, const void* pre,
                    const double* post, void* stream);
/* Multi-GPU form with both all-to-all exchanges fused into the transform over NVLink peer memory.
 * peers: HOST array of `world` (<= 8) peer-mapped device pointers, peers[r] = rank r's grid-column shard
 * [rows][ncol] (rank r owns grid points [r*ncol, (r+1)*ncol)).  This rank transforms the nvec vectors in rows
 * row0 .. row0+nvec-1 of every shard: the z/y kernel GATHERS each x-plane from the owning ranks (P2P loads), the
 * intermediate stays in `work` ([nvec][ldv >= ng], local), the x kernel SCATTERS post[G]*result into the shards
 * (P2P stores).  The caller barriers across ranks before (shards complete) and after (scatter visible). */
int isdf_dft3d_dmma_p2p(void* handle, void* const* peers, int world, long ncol, long row0, void* work, long nvec,
                        long ldv, const int* mesh, const void* pre, const double* post, void* stream);
/* The same fused exchange for the register-resident FFT kernels (meshes isdf_fft3d_reg_supported accepts, n1 > 1). */
int isdf_fft3d_reg_p2p(void* handle, void* const* peers, int world, long ncol, long row0, void* work, long nvec,
                       long ldv, const int* mesh, const void* pre, const double* post, void* stream);

/* Per-q tables of the Coulomb stage generated on the device:
 * fftisdf.py:114-115  get_coulG(cell, k=vq, mesh) (exxdiv=None, wrap_around=True) folded with vol/ngrid and
 * Parseval's 1/ngrid:  out[G] = sqrt(v(q+G) vol)/ng.  b [3][3] reciprocal vectors (rows) and kscaled [3]
 * (q in units of b) are HOST pointers; out is a device pointer.
 * fftisdf.py:99  fq = exp(-i coords . vq):  coords [ng][3] device, q [3] host (cartesian). */
int isdf_coulomb_weights(void* handle, const double* b_host, const double* kscaled_host, const int* mesh, double vol,
                         double* out, void* stream);
int isdf_phase_table(void* handle, const double* coords, const double* q_host, long ng, void* out, void* stream);

/* fftisdf.py:121  W_q = zeta_q @ z_q^H  in Parseval form:
 * w[z][perm[i]][perm[j]] = alpha * sum_g b[z][i][g] conj(b[z][j][g]);  exactly Hermitian output.
 * perm [batch][stridePerm...] position -> original index, or NULL. */
int isdf_herk_scatter(void* handle, const void* b, long ldb, long strideB, int n, int k, double alpha,
                      const int* perm, long stridePerm, void* w, long ldw, long strideW, int batch, void* stream);

/* Input producer for synthetic cells (fftisdf.py:350-352,367-370 pbc_eval_gto):
 * out[k][g][mu] = sum_T e^{ik.T} N_mu P_mu(r_g-c_mu-T) exp(-alpha_mu |r_g-c_mu-T|^2).
 * aos: device array of nao records {cx,cy,cz,alpha,norm, coef[3], pw[3][3] (int), nterm (int)} of
 * isdf_ao_desc_bytes() bytes each; images [nimg][3]; kphase [nk][nimg] c128 = exp(i k.T). */
int isdf_ao_desc_bytes(void);
int isdf_eval_ao(void* handle, const double* coords, long npts, const void* aos, int nao, const double* images,
                 int nimg, const void* kphase, int nk, void* out, void* stream);

/* data movement: dst = conj(src) (time-reversal partner W_{-q} = W_q^*), and row gather
 * dst[z][i][:] = src[z][idx[z][i]][:] (zeros where idx < 0)  (fftisdf.py:388 x0[:, mask, :]). */
int isdf_conj_copy(void* handle, const void* src, void* dst, long n, void* stream);
int isdf_gather_rows(void* handle, const void* src, long lds, long strideS, const int* idx, long strideI, int nrows,
                     long ncols, void* dst, long ldd, long strideD, int batch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ISDF_B200_H */
