"""CPU parity of the CSV checkpoint codec (csrc/csv_codec.cu, include/lib/csv.h) with the reference's lib/csv.c: byte-identical
output of write_csv_contents, bit-identical values from read_csv_contents -- against the compiled reference (oracle/_ref) when it
is there, and always against a Python restatement of lib/csv.c:28-67 (Python's %f and float() are correctly rounded, like glibc)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from helpers import load_ref, ptr, ref_available

A_CSV = b"1,2.3,3,4,509,6,7,8,\n9,\n"        # data/a.csv of the reference (CSV reader known answer)


@pytest.fixture(scope="module")
def b():
    import bla_b200 as lib
    return lib


def ref_csv():
    lib = load_ref("f32")
    lib.read_csv_contents.restype = C.POINTER(C.c_float)
    lib.read_csv_contents.argtypes = [C.c_char_p]
    lib.write_csv_contents.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    return lib


def py_format(a):
    """lib/csv.c:56-67"""
    rows, cols = a.shape
    return "".join("".join("%f," % float(v) for v in a[r]) + "\n" for r in range(rows)).encode()


_NUM = re.compile(r"^[ \t\n\v\f\r]*[+-]?(?:inf(?:inity)?|nan|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)", re.I)


def py_parse(text):
    """lib/csv.c:28-54: a value ends at ',' or at a newline that follows a non-empty field; '\\r' is skipped; atof semantics"""
    out, field = [], []
    for ch in text.decode("latin-1"):
        if ch == "," or (ch == "\n" and field):
            m = _NUM.match("".join(field))
            out.append(np.float32(float(m.group(0))) if m else np.float32(0))
            field = []
        elif ch not in "\n\r":
            field.append(ch)
    return np.array(out, np.float32)


def special_floats():
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, 200000, dtype=np.uint64).astype(np.uint32)       # every exponent, both signs
    v = bits.view(np.float32)
    v = v[np.isfinite(v)]
    ties = np.array([k / 2.0 ** s for s in range(1, 30) for k in (1, 3, 5, 7, 1001, 4097)], np.float32)   # exact halves at 1e-6
    edge = np.array([0.0, -0.0, 1e-7, 5e-7, 4.9999997e-7, 5.0000003e-7, -5e-7, 0.9999995, 0.99999946, 1.0, 16777216.0, 1e10, 3.4e38,
                     -3.4e38, 9.2e18, 4.6e18, 1e-45, 123456.789, -0.02, 1 / 255.0], np.float32)
    return np.concatenate([v, ties, -ties, edge, rng.normal(0, 0.1, 50000).astype(np.float32)])


def test_format_is_byte_identical_to_printf(b):
    v = special_floats()
    v = v[: v.size // 7 * 7].reshape(-1, 7)
    want = py_format(v)
    n = b.bla_csv_format(ptr(v), 7, v.shape[0], None, 0)
    assert n == len(want)
    buf = C.create_string_buffer(n)
    assert b.bla_csv_format(ptr(v), 7, v.shape[0], buf, n) == n
    assert buf.raw == want


def test_format_inf_nan(b):
    v = np.array([[np.inf, -np.inf, np.nan]], np.float32)
    buf = C.create_string_buffer(64)
    n = b.bla_csv_format(ptr(v), 3, 1, buf, 64)
    assert buf.raw[:n] in (b"inf,-inf,nan,\n", b"inf,-inf,-nan,\n")


@pytest.mark.skipif(not ref_available("f32"), reason="oracle/_ref not built")
def test_write_and_read_match_the_compiled_reference(b, tmp_path):
    ref = ref_csv()
    v = special_floats()[:60000].reshape(-1, 100)
    ours, theirs = str(tmp_path / "ours.csv"), str(tmp_path / "ref.csv")
    b.write_csv_contents(ours.encode(), ptr(v), 100, v.shape[0])
    ref.write_csv_contents(theirs.encode(), ptr(v), 100, v.shape[0])
    assert open(ours, "rb").read() == open(theirs, "rb").read()
    got = np.ctypeslib.as_array(b.read_csv_contents(theirs.encode()), shape=(v.size,)).copy()
    want = np.ctypeslib.as_array(ref.read_csv_contents(theirs.encode()), shape=(v.size,)).copy()
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_reader_known_answer_a_csv(b, tmp_path):
    p = tmp_path / "a.csv"
    p.write_bytes(A_CSV)
    got = np.ctypeslib.as_array(b.read_csv_contents(str(p).encode()), shape=(9,))
    assert np.array_equal(got, np.array([1, 2.3, 3, 4, 509, 6, 7, 8, 9], np.float32))


def test_parse_tokenisation_quirks(b):
    text = (b"1.5,-2.25e-3,,+4,\r\n 7.0,8e2,0x10,\n\n9.75\n10,11\r\n12,abc,1.5abc,-.5,1e999,-1e-999,\n13,14")   # last row: 14 is dropped
    out = C.POINTER(C.c_float)()
    n = b.bla_csv_parse(text, len(text), C.byref(out))
    got = np.ctypeslib.as_array(out, shape=(n,)).copy()
    want = py_parse(text)
    want[6] = 16.0          # atof reads hexadecimal floats; the Python restatement's regex does not
    assert n == want.size
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (got, want)


def test_large_text_parallel_parse_and_round_trip(b):
    rng = np.random.default_rng(5)
    v = rng.normal(0, 3, (1500, 400)).astype(np.float32)                # ~6 MB of text: split across threads
    n = b.bla_csv_format(ptr(v), 400, 1500, None, 0)
    buf = C.create_string_buffer(n)
    b.bla_csv_format(ptr(v), 400, 1500, buf, n)
    assert buf.raw == py_format(v)
    out = C.POINTER(C.c_float)()
    cnt = b.bla_csv_parse(buf, n, C.byref(out))
    got = np.ctypeslib.as_array(out, shape=(cnt,)).copy()
    assert cnt == v.size
    want = np.array([np.float32(float("%f" % float(x))) for x in v.ravel()], np.float32)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.abs(got.astype(np.float64) - v.ravel()).max() <= 5e-7 + 2.0 ** -20          # six decimals + one float32 ulp at |x| < 16


def test_fixed_notation_fields_match_strtod(b):
    """The reader's one-walk path for [-]digits[.digits] (<= 15 digits: an exact integer over an exact power of ten, one IEEE
    division) and its hand-over to the general path (16+ digits, a lone sign or dot, signs the walk does not take) give the
    float atof-then-cast gives -- Python's float() is the same correctly rounded strtod."""
    rng = np.random.default_rng(15)
    toks = []
    for _ in range(40000):
        nd = int(rng.integers(1, 19))                                   # 1..18 digits: both sides of the 15-digit limit
        digits = "".join(rng.choice(list("0123456789"), nd))
        cut = int(rng.integers(0, nd + 1))
        t = digits[:cut] + ("." if cut < nd or rng.random() < 0.2 else "") + digits[cut:]
        toks.append(("-" if rng.random() < 0.4 else "") + t)
    toks += ["0", "-0", "-0.0", "0.000000", "-0.000000", "5.", "-.5", ".5", "999999999999999", "9999999999999999", "0.999999999999999",
             "123456789012345.", "1234567.12345678", "4.9999997e-7", "16777217", "0.1", "-0.3", "7.0000005", "33554434.000000"]
    text = (",".join(toks) + ",\n").encode()
    out = C.POINTER(C.c_float)()
    n = b.bla_csv_parse(text, len(text), C.byref(out))
    assert n == len(toks)
    got = np.ctypeslib.as_array(out, shape=(n,)).copy()
    want = np.array([np.float32(float(t)) for t in toks], np.float32)
    bad = np.nonzero(got.view(np.uint32) != want.view(np.uint32))[0]
    assert bad.size == 0, [(toks[i], got[i], want[i]) for i in bad[:5]]
