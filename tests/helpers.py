"""Shared test plumbing: ctypes bindings for the compiled reference (oracle/_ref), for the CPU
restatement (oracle/libbla_oracle_*.so) and small utilities.  Test infrastructure only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def ctype_of(dtype):
    return C.c_double if np.dtype(dtype) == np.float64 else C.c_float


def make_matrix_struct(real):
    class Matrix(C.Structure):
        _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("data", C.POINTER(real))]

    return Matrix


MatrixF64 = make_matrix_struct(C.c_double)
MatrixF32 = make_matrix_struct(C.c_float)


class ConvDataF64(C.Structure):
    _fields_ = [(n, C.POINTER(MatrixF64)) for n in ("im2col", "kernel_matrix", "product", "output")]


class ConvDataF32(C.Structure):
    _fields_ = [(n, C.POINTER(MatrixF32)) for n in ("im2col", "kernel_matrix", "product", "output")]


def ensure_oracle_built():
    f64 = os.path.join(ORACLE_DIR, "libbla_oracle_f64.so")
    f32 = os.path.join(ORACLE_DIR, "libbla_oracle_f32.so")
    src = os.path.join(ORACLE_DIR, "bla_oracle.c")
    stale = any((not os.path.exists(p)) or os.path.getmtime(p) < os.path.getmtime(src) for p in (f64, f32))
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return f64, f32


_oracles = {}


def load_oracle(dtype):
    """The CPU restatement for np.float64 or np.float32."""
    dtype = np.dtype(dtype)
    if dtype not in _oracles:
        f64, f32 = ensure_oracle_built()
        lib = C.CDLL(f64 if dtype == np.float64 else f32)
        lib.orc_real_size.restype = C.c_int
        assert lib.orc_real_size() == dtype.itemsize
        real = ctype_of(dtype)
        lib.orc_frobenius.restype = real
        lib.orc_max.restype = real
        lib.orc_scale.argtypes = [C.c_size_t, C.c_void_p, real]
        lib.orc_dense_forward.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_double, C.c_void_p, C.c_void_p]
        lib.orc_mlp_step.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 8 + [C.c_double, C.c_int,
                                                                                C.c_void_p, C.c_void_p,
                                                                                C.c_void_p, C.c_int]
        lib.orc_hinge_iter.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_float, C.c_void_p]
        _oracles[dtype] = lib
    return _oracles[dtype]


_refs = {}


def ref_available(variant="f64"):
    return os.path.exists(os.path.join(REF_DIR, f"libref_{variant}.so"))


def load_ref(variant="f64"):
    """The reference's own C, compiled by oracle/build_ref.sh (variants: f64, f32, *_convfix)."""
    if variant not in _refs:
        lib = C.CDLL(os.path.join(REF_DIR, f"libref_{variant}.so"))
        f32 = variant.startswith("f32")
        real = C.c_float if f32 else C.c_double
        M = MatrixF32 if f32 else MatrixF64
        P = C.POINTER(M)
        lib.make_matrix.restype = P
        lib.clone_matrix.restype = P
        lib.clone_matrix.argtypes = [M]
        lib.matrix_multiply.restype = P
        lib.matrix_multiply.argtypes = [M, M]
        lib.matrix_scale.argtypes = [P, real]
        lib.matrix_row_sum.restype = P
        lib.matrix_row_sum.argtypes = [M]
        lib.matrix_col_sum.restype = P
        lib.matrix_col_sum.argtypes = [M]
        lib.frobenius_norm.restype = real
        lib.frobenius_norm.argtypes = [M]
        lib.max_value.restype = real
        lib.max_value.argtypes = [M]
        lib.free_matrix.argtypes = [P]
        lib.MatrixT = M
        lib.real = real
        lib.np_dtype = np.float32 if f32 else np.float64
        lib.ConvDataT = ConvDataF32 if f32 else ConvDataF64
        _refs[variant] = lib
    return _refs[variant]


def as_matrix(lib_or_struct, arr):
    """Wrap a C-contiguous 2-D numpy array as a by-value struct Matrix (no copy)."""
    M = getattr(lib_or_struct, "MatrixT", lib_or_struct)
    assert arr.flags["C_CONTIGUOUS"] and arr.ndim == 2
    real = M._fields_[2][1]._type_
    return M(arr.shape[0], arr.shape[1], arr.ctypes.data_as(C.POINTER(real)))


def matrix_to_numpy(mptr, dtype):
    m = mptr.contents if hasattr(mptr, "contents") else mptr
    n = m.rows * m.cols
    return np.ctypeslib.as_array(m.data, shape=(n,)).reshape(m.rows, m.cols).astype(dtype, copy=True)


def planes(lib, arr3):
    """[C][H][W] contiguous numpy -> ctypes array of C struct Matrix planes sharing its memory."""
    M = lib.MatrixT
    real = M._fields_[2][1]._type_
    Cn, H, W = arr3.shape
    out = (M * Cn)()
    for c in range(Cn):
        out[c] = M(H, W, arr3[c].ctypes.data_as(C.POINTER(real)))
    return out


def kernel_table(lib, k4):
    """[F][C][k][k] contiguous numpy -> Matrix** as the reference's conv() expects."""
    M = lib.MatrixT
    F = k4.shape[0]
    rows = [planes(lib, k4[f]) for f in range(F)]
    table = (C.POINTER(M) * F)(*[C.cast(r, C.POINTER(M)) for r in rows])
    table._keep = rows
    return table


def rel_err(got, want):
    """Norm-wise relative error ||got-want||_F / ||want||_F (SURVEY.md §7 hard part 6)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    denom = np.linalg.norm(want.ravel())
    if denom == 0:
        return float(np.linalg.norm(got.ravel()))
    return float(np.linalg.norm((got - want).ravel()) / denom)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)
