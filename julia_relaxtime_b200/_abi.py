"""ctypes mirror of include/pnjl_b200.h (structs, record offsets, status bits)."""
import ctypes as C

import numpy as np

ABI_VERSION = 12
REC_DOUBLES = 32
REC_X, REC_MASS, REC_OMEGA, REC_PRESSURE, REC_RHO_NORM, REC_ENTROPY, REC_ENERGY = 0, 5, 8, 9, 10, 11, 12
REC_NQ, REC_NQBAR, REC_RESNORM, REC_ITER, REC_STATUS, REC_NEVAL, REC_RHO, REC_NTHERMO = 13, 16, 19, 20, 21, 22, 23, 26
REC_T, REC_MU, REC_XI, REC_NFUSED = 27, 28, 29, 30

ST_CONVERGED, ST_USED_TR, ST_TR_ATTEMPTED, ST_USED_MULTISEED = 1, 2, 4, 8
ST_SEED_SHIFT, ST_SEED_MASK, ST_PHASE_SWITCH, ST_NONFINITE, ST_ALL_SEEDS_FAILED = 4, 0x70, 128, 256, 512
ST_PROMOTED, ST_REFINED, ST_CAND_SHIFT, ST_CAND_MASK, ST_NO_RESULT = 1024, 2048, 12, 0x3000, 16384
ST_MASS_INVERSION = 32768

AUX_DOUBLES = 16
AUX_NAMES = ("A_u", "A_s", "G_u", "G_s", "K0_plus", "K0_minus", "K123_plus", "K123_minus", "K4567_plus", "K4567_minus",
             "K8_plus", "K8_minus", "K08_plus", "K08_minus", "det_K_plus", "det_K_minus")
AUX = {name: i for i, name in enumerate(AUX_NAMES)}

SEED_EXPLICIT, SEED_AUTO, SEED_MULTI = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class PnjlConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("hbarc", "Lambda", "m_ud0", "m_s0", "G", "K", "T0", "a0", "a1", "a2", "b3", "rho0")] + [
        ("Nc", C.c_int32), ("p_num", C.c_int32), ("t_num", C.c_int32),
        ("p_nodes", c_double_p), ("p_w", c_double_p), ("c_nodes", c_double_p), ("c_w", c_double_p),
        ("xtol", C.c_double), ("ftol", C.c_double), ("residual_norm_max", C.c_double), ("phi_tol", C.c_double),
        ("max_iter", C.c_int32), ("tr_fallback", C.c_int32), ("auto_multiseed_fallback", C.c_int32),
        ("omega_tie_rel", C.c_double), ("device", C.c_int32), ("lanes_per_solve", C.c_int32),
        ("predict_tol", C.c_double), ("isospin_symmetric", C.c_int32), ("schedule", C.c_int32),
        ("isotropic_collapse", C.c_int32)]


class PnjlBoundary(C.Structure):
    _fields_ = [("T_MeV", c_double_p), ("mu_c_MeV", c_double_p), ("n", C.c_int32), ("T_CEP_MeV", C.c_double)]


class PnjlStats(C.Structure):
    _fields_ = [("kernel_launches", C.c_int64), ("kernel_ms", C.c_double), ("lanes_per_solve", C.c_int32),
                ("blocks", C.c_int32), ("threads", C.c_int32), ("regs_per_thread", C.c_int32),
                ("smem_bytes", C.c_int32)]


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def iptr(a):
    return a.ctypes.data_as(c_int32_p)


def as_f64(a, n=None):
    a = np.ascontiguousarray(np.atleast_1d(a), dtype=np.float64)
    if n is not None and a.size != n:
        a = np.ascontiguousarray(np.broadcast_to(a, (n,)), dtype=np.float64)
    return a


def check_layout(L):
    """Assert that the ctypes mirrors above have the library's struct layout (pnjl_sizeof_* / pnjl_config_field_offset)."""
    L.pnjl_sizeof_config.restype = C.c_int64
    L.pnjl_sizeof_boundary.restype = C.c_int64
    L.pnjl_sizeof_stats.restype = C.c_int64
    L.pnjl_config_field_offset.restype = C.c_int64
    L.pnjl_config_field_offset.argtypes = [C.c_char_p]
    bad = []
    if L.pnjl_sizeof_config() != C.sizeof(PnjlConfig):
        bad.append("sizeof(pnjl_config) %d != %d" % (L.pnjl_sizeof_config(), C.sizeof(PnjlConfig)))
    if L.pnjl_sizeof_boundary() != C.sizeof(PnjlBoundary):
        bad.append("sizeof(pnjl_boundary)")
    if L.pnjl_sizeof_stats() != C.sizeof(PnjlStats):
        bad.append("sizeof(pnjl_stats)")
    for name, _ in PnjlConfig._fields_:
        off = L.pnjl_config_field_offset(name.encode())
        if off != getattr(PnjlConfig, name).offset:
            bad.append("offset of %s: library %d, mirror %d" % (name, off, getattr(PnjlConfig, name).offset))
    return bad


def check_records(out, n_doubles, what="records"):
    """A caller-supplied result buffer goes to the C ABI as a raw pointer: refuse anything the kernels / memcpy would overrun
    or mis-stride (wrong dtype, non-contiguous view, read-only, too small, not 16-byte aligned)."""
    if not isinstance(out, np.ndarray) or out.dtype != np.float64:
        raise ValueError("%s buffer must be a float64 numpy array" % what)
    if not out.flags["C_CONTIGUOUS"] or not out.flags["WRITEABLE"]:
        raise ValueError("%s buffer must be C-contiguous and writeable" % what)
    if out.size != n_doubles:
        raise ValueError("%s buffer holds %d doubles, the call writes %d" % (what, out.size, n_doubles))
    if out.ctypes.data % 16 != 0:
        raise ValueError("%s buffer must be 16-byte aligned" % what)
    return out
