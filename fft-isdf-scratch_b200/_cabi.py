"""ctypes binding of libisdf_b200.so (the C-ABI drop-in boundary, include/isdf_b200.h).

No torch types cross the ABI: every call passes raw device pointers, explicit sizes/strides and a
cudaStream_t.  There is NO fallback: if the shared library is missing or the device is not a
compute-capability-10 GPU, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_LIB = None
_LIB_PATH = os.environ.get("ISDF_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                                                         "libisdf_b200.so")  # env override: kernel-tuning builds only

c_void_p, c_int, c_long, c_double, c_size_t = C.c_void_p, C.c_int, C.c_long, C.c_double, C.c_size_t
P_int = C.POINTER(C.c_int)

# name -> argtypes (all return int unless noted)
SIGNATURES = {
    "isdf_abi_version": [],
    "isdf_create": [c_int, C.POINTER(c_void_p)],
    "isdf_destroy": [c_void_p],
    "isdf_last_error": [c_void_p],
    "isdf_select_gram": [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "isdf_gram_conja": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                        c_int, c_int, c_int, c_int, c_void_p],
    "isdf_gram_conjb": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                        c_int, c_int, c_int, c_int, c_void_p],
    "isdf_ktransform_square_rows": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_long, c_int, c_int,
                                    P_int, c_void_p, c_int, c_void_p, c_void_p, c_long, c_void_p, c_void_p],
    "isdf_ktransform_rows_ex": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_long, c_int, c_int,
                                P_int, c_void_p, c_int, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_long,
                                c_long, c_double, c_void_p],
    "isdf_gemm_hn": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                     c_int, c_int, c_int, c_int, c_void_p],
    "isdf_rowdot_conj_sum": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_void_p, c_void_p],
    "isdf_scale_rows": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "isdf_herk_scatter": [c_void_p, c_void_p, c_long, c_long, c_int, c_int, c_double, c_void_p, c_long, c_void_p,
                          c_long, c_long, c_int, c_void_p],
    "isdf_gemm_nn": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                     c_int, c_int, c_int, c_int, c_void_p],
    "isdf_conj_copy": [c_void_p, c_void_p, c_void_p, c_long, c_void_p],
    "isdf_gather_rows": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_int, c_long, c_void_p, c_long,
                         c_long, c_int, c_void_p],
    "isdf_pchol_workspace_bytes": [c_int, c_int, C.POINTER(c_size_t)],
    "isdf_pchol": [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_int, c_void_p, c_void_p,
                   c_void_p, c_void_p, c_void_p],
    "isdf_pchol_real": [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_int, c_void_p, c_void_p,
                        c_void_p, c_void_p, c_void_p],
    "isdf_chol_nopivot": [c_void_p, c_void_p, c_int, c_int, c_int, c_double, c_int, c_void_p, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p],
    "isdf_trsm_sweep": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_long, c_int, c_int, c_void_p],
    "isdf_qrcp": [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p],
    "isdf_gelsy_rank": [c_void_p, c_void_p, c_void_p, c_int, c_int, c_double, c_void_p, c_void_p, c_void_p],
    "isdf_gelsy_extract": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                           c_void_p, c_void_p, c_void_p, c_void_p],
    "isdf_gelsy_rhat": [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p],
    "isdf_gelsy_q1_finish": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "isdf_gelsy_q1h_finish": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p],
    "isdf_hermitize": [c_void_p, c_void_p, c_int, c_int, c_void_p],
    "isdf_herk_to_peers": [c_void_p, c_void_p, c_long, c_long, c_int, c_int, c_double, c_void_p, c_long, c_int, c_void_p],
    "isdf_sum_slabs_herm": [c_void_p, c_void_p, c_int, c_int, c_long, c_long, c_void_p, c_long, c_long, c_int, c_void_p],
    "isdf_gemm_hn_herm": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                          c_int, c_int, c_int, c_void_p],
    "isdf_gemm_tn": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_void_p, c_long, c_long,
                     c_int, c_int, c_int, c_int, c_void_p],
    "isdf_trsm_prepare": [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                          c_void_p, c_void_p],
    "isdf_trsm_sweeps": [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_long, c_int, c_void_p],
    "isdf_ktransform_square": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_long, c_long, c_int,
                               c_int, P_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_long, c_void_p, c_void_p],
    "isdf_ktransform_ex": [c_void_p, c_void_p, c_long, c_long, c_void_p, c_long, c_long, c_long, c_long, c_int,
                           c_int, P_int, c_void_p, c_int, c_int, c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p,
                           c_long, c_long, c_double, c_void_p],
    "isdf_fft3d_batched": [c_void_p, c_void_p, c_long, c_long, P_int, c_void_p, c_void_p, c_long, c_void_p],
    "isdf_fft_release_plans": [c_void_p],
    "isdf_fft3d_reg_supported": [P_int],
    "isdf_fft3d_reg": [c_void_p, c_void_p, c_long, c_long, P_int, c_void_p, c_void_p, c_long, c_void_p],
    "isdf_fft3d_reg_p2p": [c_void_p, C.POINTER(c_void_p), c_int, c_long, c_long, c_void_p, c_long, c_long, P_int,
                           c_void_p, c_void_p, c_void_p],
    "isdf_dft3d_dmma": [c_void_p, c_void_p, c_long, c_long, P_int, c_void_p, c_void_p, c_void_p],
    "isdf_dft3d_dmma_p2p": [c_void_p, C.POINTER(c_void_p), c_int, c_long, c_long, c_void_p, c_long, c_long, P_int,
                            c_void_p, c_void_p, c_void_p],
    "isdf_ao_desc_bytes": [],
    "isdf_eval_ao": [c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p],
    "isdf_coulomb_weights": [c_void_p, C.POINTER(c_double), C.POINTER(c_double), P_int, c_double, c_void_p, c_void_p],
    "isdf_phase_table": [c_void_p, c_void_p, C.POINTER(c_double), c_long, c_void_p, c_void_p],
}


def lib_path():
    return _LIB_PATH


def load():
    """Load the shared library and declare every prototype.  Raises if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIB_PATH):
        raise RuntimeError(
            f"{_LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
            "the ISDF build has no CPU or PyTorch fallback")
    lib = C.CDLL(_LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.argtypes = argtypes
        fn.restype = C.c_char_p if name == "isdf_last_error" else C.c_int
    _LIB = lib
    return lib


class IsdfError(RuntimeError):
    pass


class Handle:
    """One handle per (process, device).  Not thread-safe per handle."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_void_p()
        rc = self.lib.isdf_create(int(device), C.byref(h))
        if rc != 0 or not h.value:
            raise IsdfError(f"isdf_create(device={device}) failed with status {rc}: needs an sm_100 (B200) GPU")
        self.h = h
        self.device = int(device)

    def check(self, rc: int, what: str):
        if rc != 0:
            msg = self.lib.isdf_last_error(self.h)
            raise IsdfError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.isdf_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
