"""Thin torch-tensor wrappers around the C-ABI entry points of libisdf_b200.so.

torch is used for device memory and streams only; every arithmetic operation on the hot path is
one of the hand-written sm_100a kernels behind `_cabi.Handle`.  All tensors are complex128 /
float64 / int32, contiguous, on the handle's device.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._cabi import Handle

KT_NMAX = 8
TB = 64  # triangular-sweep block size (csrc/pchol.cu)

c128 = torch.complex128


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk(t, dtype):
    assert t.is_cuda and t.dtype == dtype and t.is_contiguous(), (t.dtype, t.is_contiguous())


class _Timed:
    def __init__(self, ops, name):
        self.ops, self.name = ops, name

    def __enter__(self):
        if self.ops.kernel_events is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()

    def __exit__(self, *exc):
        if self.ops.kernel_events is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            self.ops.kernel_events.setdefault(self.name, []).append((self.e0, e1))
        return False


class IsdfOps:
    def __init__(self, device=0):
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.handle = Handle(device)
        self.lib = self.handle.lib
        self.h = self.handle.h
        self.launches = 0  # kernels launched through this object (bench.py's gpu_launches)
        self.kernel_events = None   # name -> [(start, stop)] CUDA event pairs when per-kernel timing is on

    def timed(self, name):
        """Context manager: CUDA-event bracket around one kernel (group) on the current stream, accumulated per name
        when `kernel_events` is a dict (bench.py's per-kernel roofline); free otherwise."""
        return _Timed(self, name)

    def kernel_ms(self):
        """Sum of the recorded brackets per name (synchronises)."""
        torch.cuda.synchronize(self.device)
        return {k: sum(a.elapsed_time(b) for a, b in v) for k, v in (self.kernel_events or {}).items()}

    # ---- K1: selection Gram  (fftisdf.py:376-379) --------------------------------------
    def select_gram(self, x0):
        _chk(x0, c128)
        nk, n0, nao = x0.shape
        x4c = torch.empty((n0, n0), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_select_gram(self.h, _ptr(x0), nk, n0, nao, _ptr(x4c), _stream()),
                          "isdf_select_gram")
        self.launches += 1
        return x4c

    # ---- K2/K5a: batched pivoted Cholesky ------------------------------------------------
    def pchol(self, a, max_steps, tol=-1.0, nb=32, real=False):
        """a: [batch, n, n] Hermitian PSD (destroyed).  Returns (u, piv, rank, next_pivot).
        real=True: imaginary parts are exactly zero (selection matrix) -> wider panels."""
        _chk(a, c128)
        batch, n, _ = a.shape
        ldu = max(1, max_steps)
        u = torch.empty((batch, ldu, n), dtype=c128, device=self.device)
        piv = torch.empty((batch, n), dtype=torch.int32, device=self.device)
        rank = torch.empty((batch,), dtype=torch.int32, device=self.device)
        nxt = torch.empty((batch,), dtype=torch.float64, device=self.device)
        nbytes = C.c_size_t()
        self.lib.isdf_pchol_workspace_bytes(n, batch, C.byref(nbytes))
        work = torch.empty((nbytes.value,), dtype=torch.uint8, device=self.device)
        fn = self.lib.isdf_pchol_real if real else self.lib.isdf_pchol
        self.handle.check(fn(self.h, _ptr(a), n, batch, int(max_steps), float(tol), int(nb), _ptr(u),
                             ldu, _ptr(piv), _ptr(rank), _ptr(nxt), _ptr(work), _stream()), "isdf_pchol")
        npan = max(1, -(-max_steps // nb))
        self.launches += 3 + 2 * npan
        return u, piv, rank, nxt

    # ---- batched conj(A) B^T  (fftisdf.py:38, :76) -----------------------------------------
    def gram_conja(self, a, b, out=None):
        """out[z] = conj(a[z]) @ b[z]^T; a, b may be strided views with a unit-stride last axis."""
        assert a.is_cuda and b.is_cuda and a.dtype == c128 and b.dtype == c128
        assert a.stride(2) == 1 and b.stride(2) == 1
        batch, m, k = a.shape
        _, n, _ = b.shape
        if out is None:
            out = torch.empty((batch, m, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gram_conja(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                   b.stride(0), _ptr(out), n, m * n, m, n, k, batch, _stream()),
                          "isdf_gram_conja")
        self.launches += 1
        return out

    def gram_conjb(self, a, b, out=None):
        """out[z] = a[z] @ conj(b[z])^T  (a [m,k], b [n,k] strided views with unit-stride last axis)."""
        assert a.is_cuda and b.is_cuda and a.dtype == c128 and b.dtype == c128
        assert a.stride(2) == 1 and b.stride(2) == 1
        batch, m, k = a.shape
        _, n, _ = b.shape
        if out is None:
            out = torch.empty((batch, m, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gram_conjb(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                   b.stride(0), _ptr(out), n, m * n, m, n, k, batch, _stream()),
                          "isdf_gram_conjb")
        self.launches += 1
        return out

    def gemm_nn(self, a, b):
        _chk(a, c128), _chk(b, c128)
        batch, m, k = a.shape
        _, _, n = b.shape
        out = torch.empty((batch, m, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gemm_nn(self.h, _ptr(a), k, m * k, _ptr(b), n, k * n, _ptr(out), n, m * n,
                                                m, n, k, batch, _stream()), "isdf_gemm_nn")
        self.launches += 1
        return out

    # ---- k<->R transform, square, k<->R transform ------------------------------------------
    def ktransform_rows(self, vin, in_sk, in_sr, out, out_sq, out_sr, out_c0, nrows, ncols, kmesh, uaxes_host, conj2,
                        qslot=None, rowmap=None, rowmap_sq=0, diag=None):
        """Register-resident k-transform (small k-meshes).  Returns False when the mesh is unsupported."""
        km = (C.c_int * 3)(*[int(x) for x in kmesh])
        assert uaxes_host.dtype == np.complex128 and uaxes_host.shape == (3, KT_NMAX, KT_NMAX)
        rc = self.lib.isdf_ktransform_square_rows(
            self.h, _ptr(vin), in_sk, in_sr, _ptr(out), out_sq, out_sr, out_c0, nrows, ncols, km,
            C.c_void_p(uaxes_host.ctypes.data), int(conj2), _ptr(qslot), _ptr(rowmap), rowmap_sq, _ptr(diag),
            _stream())
        if rc == -2:
            return False
        self.handle.check(rc, "isdf_ktransform_square_rows")
        self.launches += 1
        return True

    def ktransform_rows_ex(self, vin, in_sk, in_sr, out, out_sq, out_sr, out_c0, nrows, ncols, kmesh, uaxes_host, conj2,
                           mode, table=None, tab_sk=0, tab_sr=0, scale=1.0, diag=None):
        """mode 1: scale*Re(P v)*table then second transform; mode 2: write scale*Re(P v) as a real table."""
        km = (C.c_int * 3)(*[int(x) for x in kmesh])
        rc = self.lib.isdf_ktransform_rows_ex(
            self.h, _ptr(vin), in_sk, in_sr, _ptr(out), out_sq, out_sr, out_c0, nrows, ncols, km,
            C.c_void_p(uaxes_host.ctypes.data), int(conj2), None, None, 0, _ptr(diag), int(mode), _ptr(table),
            tab_sk, tab_sr, float(scale), _stream())
        if rc == -2:
            return False
        self.handle.check(rc, "isdf_ktransform_rows_ex")
        self.launches += 1
        return True

    def gemm_hn(self, a, b):
        """out[z] = a[z]^H @ b[z]; a [batch,k,m], b [batch,k,n]."""
        _chk(a, c128), _chk(b, c128)
        batch, k, m = a.shape
        n = b.shape[2]
        out = torch.empty((batch, m, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gemm_hn(self.h, _ptr(a), m, k * m, _ptr(b), n, k * n, _ptr(out), n, m * n,
                                                m, n, k, batch, _stream()), "isdf_gemm_hn")
        self.launches += 1
        return out

    def gemm_hn_herm(self, a, b):
        """out[z] = a[z]^H @ b[z] for a product known to be Hermitian: lower tiles + mirrored conjugate (exactly
        Hermitian output, half the tensor work of gemm_hn + hermitize); a, b [batch,k,n]."""
        _chk(a, c128), _chk(b, c128)
        batch, k, n = a.shape
        assert b.shape == a.shape
        out = torch.empty((batch, n, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gemm_hn_herm(self.h, _ptr(a), n, k * n, _ptr(b), n, k * n, _ptr(out), n, n * n,
                                                     n, k, batch, _stream()), "isdf_gemm_hn_herm")
        self.launches += 1
        return out

    def rowdot_conj_sum(self, y, x, scale):
        _chk(y, c128), _chk(x, c128)
        nz, nrows, ncols = x.shape
        out = torch.empty((nrows,), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_rowdot_conj_sum(self.h, _ptr(y), _ptr(x), nz, nrows, ncols, float(scale),
                                                        _ptr(out), _stream()), "isdf_rowdot_conj_sum")
        self.launches += 1
        return out

    def scale_rows(self, x, v):
        _chk(x, c128), _chk(v, c128)
        nz, nrows, ncols = x.shape
        out = torch.empty_like(x)
        self.handle.check(self.lib.isdf_scale_rows(self.h, _ptr(x), _ptr(v), nz, nrows, ncols, _ptr(out), _stream()),
                          "isdf_scale_rows")
        self.launches += 1
        return out

    def pack_uaxes_host(self, kmesh):
        from .pbc_tools import get_phase_axes
        u = np.zeros((3, KT_NMAX, KT_NMAX), dtype=np.complex128)
        for a, m in enumerate(get_phase_axes(kmesh)):
            assert m.shape[0] <= KT_NMAX, "k-mesh axis > 8 not supported by the k-transform kernels"
            u[a, : m.shape[0], : m.shape[1]] = m
        return u

    def pack_uaxes(self, kmesh):
        from .pbc_tools import get_phase_axes
        u = np.zeros((3, KT_NMAX, KT_NMAX), dtype=np.complex128)
        for a, m in enumerate(get_phase_axes(kmesh)):
            assert m.shape[0] <= KT_NMAX, "k-mesh axis > 8 not supported by the k-transform kernel"
            u[a, : m.shape[0], : m.shape[1]] = m
        return torch.from_numpy(u).to(self.device)

    def ktransform_square(self, vin, in_sk, in_sg, out, out_sq, out_sg, out_si, out_g0, ng, ni, kmesh, uaxes,
                          conj2, out_g_fast, qslot=None, rowmap=None, rowmap_sq=0, diag=None):
        km = (C.c_int * 3)(*[int(x) for x in kmesh])
        self.handle.check(self.lib.isdf_ktransform_square(
            self.h, _ptr(vin), in_sk, in_sg, _ptr(out), out_sq, out_sg, out_si, out_g0, ng, ni, km, _ptr(uaxes),
            int(conj2), int(out_g_fast), _ptr(qslot), _ptr(rowmap), rowmap_sq, _ptr(diag), _stream()),
            "isdf_ktransform_square")
        self.launches += 1

    def ktransform_general(self, vin, in_sk, in_sg, out, out_sq, out_sg, out_si, ng, ni, kmesh, uaxes, conj2, mode,
                           table=None, tab_sk=0, tab_sg=0, scale=1.0, diag=None):
        """Shared-memory k-transform (axes <= 8) with the exchange modes of ktransform_rows_ex."""
        km = (C.c_int * 3)(*[int(x) for x in kmesh])
        self.handle.check(self.lib.isdf_ktransform_ex(
            self.h, _ptr(vin), in_sk, in_sg, _ptr(out), out_sq, out_sg, out_si, 0, ng, ni, km, _ptr(uaxes),
            int(conj2), 0, None, None, 0, _ptr(diag), int(mode), _ptr(table), tab_sk, tab_sg, float(scale),
            _stream()), "isdf_ktransform_ex")
        self.launches += 1

    # ---- K5: triangular sweeps ----------------------------------------------------------------
    def trsm_prepare(self, u, piv, rank, nP):
        batch, ldu, n = u.shape
        lfwd = torch.zeros((batch, nP, nP), dtype=c128, device=self.device)
        ubwd = torch.zeros((batch, nP, nP), dtype=c128, device=self.device)
        work = torch.empty((2, batch, nP, nP), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_trsm_prepare(self.h, _ptr(u), ldu, _ptr(piv), _ptr(rank), n, nP, batch,
                                                     _ptr(lfwd), _ptr(ubwd), _ptr(work), _stream()),
                          "isdf_trsm_prepare")
        self.launches += 3
        return lfwd, ubwd

    def trsm_sweeps(self, lfwd, ubwd, t, nact=None):
        """nact: rows of t that can be non-zero (>= every rank in the batch); the rest is neither read nor written."""
        _chk(t, c128)
        batch, nP, ng = t.shape
        nact = nP if nact is None else int(nact)
        self.handle.check(self.lib.isdf_trsm_sweeps(self.h, _ptr(lfwd), _ptr(ubwd), _ptr(t), nP, nact, ng, ng, batch,
                                                    _stream()), "isdf_trsm_sweeps")
        self.launches += 2 * (-(-nact // TB))

    def trsm_sweep(self, op, t, nact=None, backward=False, ng=None):
        """One direction: t <- U^{-H} t (op = lfwd) or t <- U^{-1} t (op = ubwd, backward=True), in place."""
        _chk(t, c128)
        batch, nP, ldt = t.shape
        nact = nP if nact is None else int(nact)
        ng = ldt if ng is None else int(ng)
        self.handle.check(self.lib.isdf_trsm_sweep(self.h, _ptr(op), _ptr(t), nP, nact, ng, ldt, batch,
                                                   int(bool(backward)), _stream()), "isdf_trsm_sweep")
        self.launches += -(-nact // TB)

    def chol_nopivot(self, a, tol=0.0, nb=32):
        """Unpivoted Cholesky a = U^H U of [batch, n, n] Hermitian matrices (lower triangle read, destroyed);
        stops at the first pivot <= tol.  Returns (u [batch, n, n], piv (identity), rank)."""
        _chk(a, c128)
        batch, n, _ = a.shape
        u = torch.empty((batch, n, n), dtype=c128, device=self.device)
        piv = torch.empty((batch, n), dtype=torch.int32, device=self.device)
        rank = torch.empty((batch,), dtype=torch.int32, device=self.device)
        nbytes = C.c_size_t()
        self.lib.isdf_pchol_workspace_bytes(n, batch, C.byref(nbytes))
        work = torch.empty((nbytes.value + batch * 64 * 64 * 16,), dtype=torch.uint8, device=self.device)
        self.handle.check(self.lib.isdf_chol_nopivot(self.h, _ptr(a), n, batch, n, float(tol), int(nb), _ptr(u), n,
                                                     _ptr(piv), _ptr(rank), _ptr(work), _stream()),
                          "isdf_chol_nopivot")
        self.launches += 2 + 3 * (-(-n // 64))
        return u, piv, rank

    # ---- K5 (reference semantics): LAPACK zgelsy restated on the device (fftisdf.py:108) ---------------
    def qrcp(self, w):
        """Householder QR with column pivoting (zgeqp3) of [batch, n, n] matrices held COLUMN-major
        (w[b, c, i] = A[i, c]); w is overwritten (w[b, c, k] = R[k, c] for k <= pos[c]).
        Returns (vt [batch, n, n] reflectors by row, tau [batch, n], piv, pos)."""
        _chk(w, c128)
        batch, n, _ = w.shape
        vt = torch.empty((batch, n, n), dtype=c128, device=self.device)
        tau = torch.empty((batch, n), dtype=c128, device=self.device)
        piv = torch.empty((batch, n), dtype=torch.int32, device=self.device)
        pos = torch.empty((batch, n), dtype=torch.int32, device=self.device)
        self.handle.check(self.lib.isdf_qrcp(self.h, _ptr(w), n, batch, _ptr(vt), _ptr(tau), _ptr(piv), _ptr(pos),
                                             _stream()), "isdf_qrcp")
        self.launches += 2
        return vt, tau, piv, pos

    def gelsy_rank(self, w, piv, rcond):
        batch, n, _ = w.shape
        rank = torch.empty((batch,), dtype=torch.int32, device=self.device)
        xwork = torch.empty((batch, 2, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gelsy_rank(self.h, _ptr(w), _ptr(piv), n, batch, float(rcond), _ptr(xwork),
                                                   _ptr(rank), _stream()), "isdf_gelsy_rank")
        self.launches += 1
        return rank

    def gemm_tn(self, a, b):
        """out[z] = a[z]^T @ b[z] (no conjugation); a [batch,k,m] (row pitch a.stride(1)), b [batch,k,n]."""
        assert a.dtype == c128 and b.dtype == c128 and a.stride(2) == 1 and b.stride(2) == 1
        batch, k, m = a.shape
        n = b.shape[2]
        out = torch.empty((batch, m, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gemm_tn(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                b.stride(0), _ptr(out), n, m * n, m, n, k, batch, _stream()),
                          "isdf_gemm_tn")
        self.launches += 1
        return out

    def gemm_tn_strided(self, a, b, out):
        """out[z] = a[z]^T @ b[z] (no conjugation) into a strided view: a [batch,k,m], b [batch,k,n], out [batch,m,n]."""
        assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
        batch, k, m = a.shape
        n = b.shape[2]
        self.handle.check(self.lib.isdf_gemm_tn(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                b.stride(0), _ptr(out), out.stride(1), out.stride(0), m, n, k, batch,
                                                _stream()), "isdf_gemm_tn")
        self.launches += 1
        return out

    def gemm_hn_strided(self, a, b, out):
        """out[z] = a[z]^H @ b[z] into a strided view: a [batch,k,m], b [batch,k,n], out [batch,m,n] (unit last stride)."""
        assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
        batch, k, m = a.shape
        n = b.shape[2]
        self.handle.check(self.lib.isdf_gemm_hn(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                b.stride(0), _ptr(out), out.stride(1), out.stride(0), m, n, k, batch,
                                                _stream()), "isdf_gemm_hn")
        self.launches += 1
        return out

    def hermitize(self, w):
        _chk(w, c128)
        batch, n, _ = w.shape
        self.handle.check(self.lib.isdf_hermitize(self.h, _ptr(w), n, batch, _stream()), "isdf_hermitize")
        self.launches += 1
        return w

    def gelsy_qr(self, a_q, rcond):
        """Stage 1 of the gelsy fit for Hermitian a_q [batch, n, n] (row-major): QRCP + rank decision.
        Returns a state dict (device tensors); `rank` is still on the device."""
        _chk(a_q, c128)
        w = torch.empty_like(a_q)
        self.conj_copy(a_q, w)                    # column-major working copy: w[c][i] = A[i][c] = conj(A[c][i])
        with self.timed("qrcp"):
            vt, tau, piv, pos = self.qrcp(w)
        with self.timed("gelsy_rank"):
            rank = self.gelsy_rank(w, piv, rcond)
        return dict(w=w, vt=vt, tau=tau, piv=piv, pos=pos, rank=rank)

    def gelsy_operators(self, st, rP, debug=False):
        """Stage 2: the dense operators of x = P Z^H [T11^-1 (Q^H b)(:rank); 0] for every matrix of the batch:
             gt   [batch, rP, n]   G = U^-H D^-1 Q1^H  (zunmqr + ztrsm of zgelsy as ONE operator: U^-H is lower triangular
                                   and well conditioned, so row i of G carries the single scale 1/|R_ii| and the
                                   explicit product is row-wise as accurate as the two-step application)
             eh   [batch, rP, n]   E^H (orthonormal rows, zero beyond rank; E = P Z1^H of ztzrzf/zunmrz)
           so that  Theta~ = G Y^T  [rank x ng],  Theta = eh^H Theta~  and  W = eh^H W~ eh.
           D = |diag R|;  D^-1 [R11 R12] P^T = U^H E^H by Cholesky-QR (twice) with U upper triangular.
           debug=True also returns q1s [batch, n, rP] = Q1 D^-1 and lfwd (the block operators of U1^-H, U2^-H; U = U2 U1)."""
        w, vt, tau, piv, pos, rank = st["w"], st["vt"], st["tau"], st["piv"], st["pos"], st["rank"]
        batch, n, _ = w.shape
        assert rP % TB == 0
        kk = min(rP, n)
        ident = torch.arange(rP, dtype=torch.int32, device=self.device).repeat(batch, 1).contiguous()
        # --- Q1 through the compact-WY form of the first `rank` reflectors: Q1 = I(:, :r) - V S^-1 V(:r, :)^H
        vv = vt[:, :kk, :]
        with self.timed("ops_vhv"):
            # V^H V as conj(Vt Vt^H): HERK (half the flops); only the strict upper triangle of V^H V is used,
            # i.e. the lower triangle of Vt Vt^H transposed -- the extract kernel reads g[j][k] for k < j
            g = torch.zeros((batch, rP, rP), dtype=c128, device=self.device)
            self.herk_strided(vv, n, n * n, kk, n, 1.0, None, 0, g, rP, rP * rP, batch)
        s = torch.empty((batch, rP, rP), dtype=c128, device=self.device)
        m = torch.empty((batch, rP, rP), dtype=c128, device=self.device)
        dinv = torch.empty((batch, rP), dtype=torch.float64, device=self.device)
        with self.timed("ops_wy_solve"):
            self.handle.check(self.lib.isdf_gelsy_extract(self.h, _ptr(g), _ptr(tau), _ptr(rank), _ptr(vt), _ptr(w),
                                                          _ptr(piv), n, rP, batch, _ptr(s), _ptr(m), _ptr(dinv),
                                                          _stream()), "isdf_gelsy_extract")
            self.launches += 1
            _, ub = self.trsm_prepare(s, ident, rank, rP)
            self.trsm_sweep(ub, m, backward=True)                             # M = S^-1 V1^H
        del s, ub, g
        with self.timed("ops_q1h"):
            gt = torch.zeros((batch, rP, n), dtype=c128, device=self.device)
            self.gemm_tn_into(m[:, :kk, :], vv, gt)                           # M^T V  = conj((V M)^H)
            self.handle.check(self.lib.isdf_gelsy_q1h_finish(self.h, _ptr(gt), _ptr(dinv), _ptr(rank), n, rP, batch,
                                                             _stream()), "isdf_gelsy_q1h_finish")   # D^-1 Q1^H
            self.launches += 1
        out = {}
        if debug:
            q1s = torch.zeros((batch, n, rP), dtype=c128, device=self.device)
            self.gemm_tn_into(vv, m[:, :kk, :], q1s)
            self.handle.check(self.lib.isdf_gelsy_q1_finish(self.h, _ptr(q1s), _ptr(dinv), _ptr(rank), n, rP, batch,
                                                            _stream()), "isdf_gelsy_q1_finish")
            out["q1s"] = q1s
        del m
        # --- E^H and the triangular factor: Cholesky-QR (twice) of the row-scaled [R11 R12] P^T
        eh = torch.empty((batch, rP, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gelsy_rhat(self.h, _ptr(w), _ptr(pos), _ptr(dinv), _ptr(rank), n, rP, batch,
                                                   _ptr(eh), _stream()), "isdf_gelsy_rhat")
        self.launches += 1
        lfs = []
        for _ in range(2):
            with self.timed("ops_cholqr_herk"):
                gg = self.herk(eh)
            with self.timed("ops_cholqr_chol"):
                u, _, rk = self.chol_nopivot(gg)
            with self.timed("ops_cholqr_solve"):
                lf, _ = self.trsm_prepare(u, ident, rk, rP)
                self.trsm_sweep(lf, eh, backward=False)
            lfs.append(lf)
            del gg, u
        with self.timed("ops_g"):
            # G = U^-H (D^-1 Q1^H) with U = U2 U1:  U^-H = U2^-H U1^-H, applied as the two substitutions already set up
            self.trsm_sweep(lfs[0], gt, backward=False)
            self.trsm_sweep(lfs[1], gt, backward=False)
        out.update(gt=gt, eh=eh, chol_rank=rk)
        if debug:
            out["lfwd"] = lfs
        return out

    def gemm_nn_strided(self, a, b, out):
        """out[z] = a[z] @ b[z] into a strided view: a [batch,m,k], b [batch,k,n], out [batch,m,n] (unit last stride)."""
        assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
        batch, m, k = a.shape
        n = b.shape[2]
        self.handle.check(self.lib.isdf_gemm_nn(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                b.stride(0), _ptr(out), out.stride(1), out.stride(0), m, n, k, batch,
                                                _stream()), "isdf_gemm_nn")
        self.launches += 1
        return out

    def gram_conja_strided(self, a, b, out):
        assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
        batch, m, k = a.shape
        n = b.shape[1]
        self.handle.check(self.lib.isdf_gram_conja(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                   b.stride(0), _ptr(out), out.stride(1), out.stride(0), m, n, k,
                                                   batch, _stream()), "isdf_gram_conja")
        self.launches += 1
        return out

    def gemm_tn_into(self, a, b, out):
        assert a.stride(2) == 1 and b.stride(2) == 1 and out.stride(2) == 1
        batch, k, m = a.shape
        n = b.shape[2]
        self.handle.check(self.lib.isdf_gemm_tn(self.h, _ptr(a), a.stride(1), a.stride(0), _ptr(b), b.stride(1),
                                                b.stride(0), _ptr(out), out.stride(1), out.stride(0), m, n, k, batch,
                                                _stream()), "isdf_gemm_tn")
        self.launches += 1
        return out

    # ---- K6: batched 3-D FFT with fused phase / weight ----------------------------------------
    @staticmethod
    def _max_prime_factor(n):
        f, p = 1, 2
        while n > 1:
            if n % p == 0:
                f = p
                n //= p
            else:
                p += 1
        return f

    def fft3d_reg_supported(self, mesh):
        """True when the register-resident FFT kernels (fft_reg.cu) have an instantiation for this mesh."""
        m = (C.c_int * 3)(*[int(x) for x in mesh])
        return bool(self.lib.isdf_fft3d_reg_supported(m))

    def fft3d(self, data, mesh, pre=None, post=None, group_vecs=0, nvec=None, ldv=None, mode="auto"):
        """data: [..., ldv] rows of length prod(mesh) (row pitch ldv), transformed in place.
        mode: "reg" (register-resident two-factor / direct-prime kernels with compile-time axis lengths: one
        shared-memory exchange per axis, every element read and written once per pass), "stockham" (generic
        shared-memory Stockham FFT, any mesh), "dmma" (tensor-core dense DFT, axes <= 48) or "auto": "reg" when the
        mesh is instantiated, else the round-1 routing by measurement (Stockham when every axis factors into primes
        <= 13 and some axis is longer than 32, the tensor-core DFT for the other meshes with every axis in [2, 48],
        Stockham with its direct-DFT stage for whatever is left)."""
        assert data.is_cuda and data.dtype == c128 and data.stride(-1) == 1
        ng = int(np.prod(mesh))
        if ldv is None:
            ldv = ng
        if nvec is None:
            nvec = data.numel() // ldv
        m = (C.c_int * 3)(*[int(x) for x in mesh])
        fits = all(2 <= int(x) <= 48 for x in mesh)
        if mode == "auto":
            if self.lib.isdf_fft3d_reg_supported(m):
                mode = "reg"
            else:
                smooth = all(self._max_prime_factor(int(x)) <= 13 for x in mesh)
                mode = "stockham" if ((smooth and max(int(x) for x in mesh) > 32) or not fits) else "dmma"
        if mode in ("reg", "reg-fused"):
            # "reg-fused": cubic meshes, both passes in one persistent launch (measured no faster; kept selectable)
            if mode == "reg-fused":
                assert int(mesh[0]) == int(mesh[1]) == int(mesh[2])
                group_vecs = -2
            self.handle.check(self.lib.isdf_fft3d_reg(self.h, _ptr(data), nvec, ldv, m, _ptr(pre), _ptr(post),
                                                      int(group_vecs), _stream()), "isdf_fft3d_reg")
            ngroups = 1 if group_vecs <= 0 else -(-nvec // int(group_vecs))
            self.launches += (2 if int(mesh[0]) > 1 else 1) * ngroups if nvec > 0 else 0
            return
        if mode == "dmma":
            assert fits, "dmma DFT needs every mesh axis in [2, 48]"
            self.handle.check(self.lib.isdf_dft3d_dmma(self.h, _ptr(data), nvec, ldv, m, _ptr(pre), _ptr(post),
                                                       _stream()), "isdf_dft3d_dmma")
            self.launches += 2 if nvec > 0 else 0       # one zy launch + one x launch for the whole batch
            return
        self.handle.check(self.lib.isdf_fft3d_batched(self.h, _ptr(data), nvec, ldv, m, _ptr(pre), _ptr(post),
                                                      int(group_vecs), _stream()), "isdf_fft3d_batched")
        gv = group_vecs if group_vecs > 0 else max(1, int(48 * 1024 * 1024 / (ng * 16)))
        self.launches += 2 * (-(-nvec // gv))

    def dft3d_p2p(self, peer_ptrs, ncol, row0, work, nvec, mesh, pre=None, post=None, mode="auto"):
        """3-D transform with the all-to-all exchanges fused over NVLink peer memory (see the C header): the
        register-resident FFT kernels when the mesh is instantiated ("reg"), else the tensor-core DFT ("dmma")."""
        world = len(peer_ptrs)
        arr = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_ptrs])
        m = (C.c_int * 3)(*[int(x) for x in mesh])
        if mode == "auto":
            mode = "reg" if (self.lib.isdf_fft3d_reg_supported(m) and int(mesh[0]) > 1) else "dmma"
        fn, name = ((self.lib.isdf_fft3d_reg_p2p, "isdf_fft3d_reg_p2p") if mode == "reg"
                    else (self.lib.isdf_dft3d_dmma_p2p, "isdf_dft3d_dmma_p2p"))
        self.handle.check(fn(self.h, arr, world, int(ncol), int(row0), _ptr(work), int(nvec), work.shape[-1], m,
                             _ptr(pre), _ptr(post), _stream()), name)
        self.launches += 2 if nvec > 0 else 0

    # ---- K7: W = alpha * B B^H, scattered through perm -------------------------------------------
    def herk(self, b, alpha=1.0, perm=None, out=None):
        _chk(b, c128)
        batch, n, k = b.shape
        if out is None:
            out = torch.empty((batch, n, n), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_herk_scatter(self.h, _ptr(b), k, n * k, n, k, float(alpha), _ptr(perm),
                                                     (perm.shape[-1] if perm is not None else 0), _ptr(out), n, n * n,
                                                     batch, _stream()), "isdf_herk_scatter")
        self.launches += 1
        return out

    def herk_strided(self, b, ldb, strideB, n, k, alpha, perm, stride_perm, out, ldw, strideW, batch):
        self.handle.check(self.lib.isdf_herk_scatter(self.h, _ptr(b), ldb, strideB, n, k, float(alpha), _ptr(perm),
                                                     stride_perm, _ptr(out), ldw, strideW, batch, _stream()),
                          "isdf_herk_scatter")
        self.launches += 1

    def herk_to_peers(self, b, ldb, strideB, n, k, alpha, dst_ptrs, ldw, batch):
        """W~ partial products with the reduce-scatter fused in: batch z goes (lower triangle, plain NVLink stores as
        the tiles finish) to dst_ptrs[z], a DEVICE int64 tensor of peer-mapped addresses (see the C header)."""
        assert dst_ptrs.dtype == torch.int64 and dst_ptrs.is_cuda and dst_ptrs.numel() == batch
        self.handle.check(self.lib.isdf_herk_to_peers(self.h, _ptr(b), ldb, strideB, n, k, float(alpha), _ptr(dst_ptrs),
                                                      ldw, batch, _stream()), "isdf_herk_to_peers")
        self.launches += 1

    def sum_slabs_herm(self, slabs, world, n, out):
        """out[z] = sum of the `world` lower-triangular slabs slabs[z, w] in rank order, mirrored (exactly Hermitian)."""
        _chk(slabs, c128), _chk(out, c128)
        batch, w, ldn, ld = slabs.shape
        assert w == world and out.shape[0] == batch
        self.handle.check(self.lib.isdf_sum_slabs_herm(self.h, _ptr(slabs), world, n, ld, ldn * ld, _ptr(out),
                                                       out.shape[2], out.shape[1] * out.shape[2], batch, _stream()),
                          "isdf_sum_slabs_herm")
        self.launches += 1

    # ---- device-side per-q tables (fftisdf.py:99, :114-115) -----------------------------------
    def coulomb_weights(self, b, kscaled, mesh, vol, out):
        bb = (C.c_double * 9)(*[float(x) for x in np.asarray(b).reshape(9)])
        ks = (C.c_double * 3)(*[float(x) for x in kscaled])
        m = (C.c_int * 3)(*[int(x) for x in mesh])
        self.handle.check(self.lib.isdf_coulomb_weights(self.h, bb, ks, m, float(vol), _ptr(out), _stream()),
                          "isdf_coulomb_weights")
        self.launches += 1

    def phase_table(self, coords_dev, q, out):
        qq = (C.c_double * 3)(*[float(x) for x in q])
        self.handle.check(self.lib.isdf_phase_table(self.h, _ptr(coords_dev), qq, coords_dev.shape[0], _ptr(out),
                                                    _stream()), "isdf_phase_table")
        self.launches += 1

    # ---- periodic Gaussian AO evaluation (input producer for SyntheticCell) ---------------------
    def eval_ao(self, coords_dev, ao_desc_dev, nao, images_dev, kphase_dev, out=None):
        npts = coords_dev.shape[0]
        nk, nimg = kphase_dev.shape
        if out is None:
            out = torch.empty((nk, npts, nao), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_eval_ao(self.h, _ptr(coords_dev), npts, _ptr(ao_desc_dev), nao,
                                                _ptr(images_dev), nimg, _ptr(kphase_dev), nk, _ptr(out), _stream()),
                          "isdf_eval_ao")
        self.launches += 1
        return out

    def conj_copy(self, src, dst):
        self.handle.check(self.lib.isdf_conj_copy(self.h, _ptr(src), _ptr(dst), src.numel(), _stream()),
                          "isdf_conj_copy")
        self.launches += 1

    def gather_rows(self, src, idx, ncols=None, out=None):
        """src [batch, R, ncols], idx [batch, nrows] int32 -> out [batch, nrows, ncols] (zeros where idx<0)."""
        _chk(src, c128)
        batch, R, nc = src.shape
        nrows = idx.shape[-1]
        if out is None:
            out = torch.empty((batch, nrows, nc), dtype=c128, device=self.device)
        self.handle.check(self.lib.isdf_gather_rows(self.h, _ptr(src), nc, R * nc, _ptr(idx), nrows, nrows, nc,
                                                    _ptr(out), nc, nrows * nc, batch, _stream()), "isdf_gather_rows")
        self.launches += 1
        return out
