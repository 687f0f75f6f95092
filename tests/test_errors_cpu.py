"""Error behaviour of the C-ABI that needs no GPU: the reference's convention is one line on stdout and exit(1)
(lib/matrix.c:36-39, :96-99); the dimension checks come before any device work, so the same calls are made against
libbla.so and the compiled reference in child processes and their stdout / exit status compared.  Also: a compute call
without a usable GPU must fail loudly (no CPU fallback), and the Python mirror must refuse to load without the library."""
import os
import subprocess
import sys

import pytest

from helpers import REF_DIR, ROOT

LIB = os.path.join(ROOT, "big-linear-algebra_b200", "libbla.so")
REF = os.path.join(REF_DIR, "libref_f32.so")

CHILD = r"""
import ctypes as C, sys
lib = C.CDLL(sys.argv[1])
class Matrix(C.Structure):
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("data", C.POINTER(C.c_float))]
buf = (C.c_float * 64)()
a = Matrix(int(sys.argv[3]), int(sys.argv[4]), buf); b = Matrix(int(sys.argv[5]), int(sys.argv[6]), buf)
if sys.argv[2] == "multiply":
    lib.matrix_multiply.argtypes = [Matrix, Matrix]; lib.matrix_multiply.restype = C.c_void_p
    lib.matrix_multiply(a, b)
else:
    lib.matrix_multiply_elementwise.argtypes = [C.POINTER(Matrix), C.POINTER(Matrix)]
    lib.matrix_multiply_elementwise(C.byref(a), C.byref(b))
print("returned")
"""


def call(lib, op, *dims):
    p = subprocess.run([sys.executable, "-c", CHILD, lib, op, *map(str, dims)], capture_output=True, text=True, timeout=120)
    return p.returncode, p.stdout


@pytest.mark.parametrize("op,dims,text", [
    ("multiply", (2, 3, 4, 5), "Attempted to multiply 2x3 matrix by 4x5 matrix, exiting\n"),
    ("elementwise", (2, 3, 3, 2), "Attempted to multiply elements of 2x3 matrix by 3x2 matrix, exiting\n"),
])
def test_dimension_mismatch_is_one_line_and_exit_1(op, dims, text):
    got = call(LIB, op, *dims)
    assert got == (1, text)
    if os.path.exists(REF):
        assert call(REF, op, *dims) == got


def test_compute_without_a_gpu_fails_loudly():
    """no CPU fallback: the first call that needs the device says so on stdout and exits 1 (skipped where a GPU is present)"""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    code, out = call(LIB, "multiply", 2, 3, 3, 2)
    assert code == 1 and out.startswith("bla: no usable CUDA device") and "no CPU fallback" in out and "returned" not in out


def test_python_mirror_refuses_to_load_without_the_library(tmp_path):
    pkg = tmp_path / "big-linear-algebra_b200"
    pkg.mkdir()
    (pkg / "__init__.py").write_text(open(os.path.join(ROOT, "big-linear-algebra_b200", "__init__.py")).read())
    child = "import importlib, sys; sys.path.insert(0, sys.argv[1]); importlib.import_module('big-linear-algebra_b200')"
    p = subprocess.run([sys.executable, "-c", child, str(tmp_path)], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "ImportError" in p.stderr and "libbla.so" in p.stderr
