"""The host-side data objects of the drop-in (csrc/host_io.cu in libbla.so, host_io/mnist_csv.c in libbla_mnist_csv.so) against
the reference's own lib/mnist_csv2.c, lib/cifar10.c, lib/bmp.c and lib/mnist_csv.c compiled into oracle/_ref (build_ref.sh), call
by call on the same inputs and the same libc rand() stream: loader arrays, both draw sequences with the model's own epoch
reset, the character pictures on stdout, CIFAR records and file offset, BMP bytes.  No GPU: none of these functions needs one.
Where the compiled reference is absent the known-answer parts still run."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from helpers import REF_DIR, ROOT
from test_data_cpu import libc, reference_take_sequence

LIB = os.path.join(ROOT, "big-linear-algebra_b200", "libbla.so")
LIB_ROW = os.path.join(ROOT, "big-linear-algebra_b200", "libbla_mnist_csv.so")
REF_DATA = os.path.join(REF_DIR, "libref_data.so")
REF_ROW = os.path.join(REF_DIR, "libref_mnist_csv.so")

libc.fopen.restype = C.c_void_p
libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
libc.fclose.argtypes = [C.c_void_p]
libc.fflush.argtypes = [C.c_void_p]
libc.free.argtypes = [C.c_void_p]
libc.feof.argtypes = [C.c_void_p]


class MnistCSV2(C.Structure):          # lib/mnist_csv2.h:5-12
    _fields_ = [("file", C.c_void_p), ("X", C.POINTER(C.c_float)), ("y", C.POINTER(C.c_float)), ("num_examples", C.c_int),
                ("num_sampled", C.c_int), ("sampled", C.POINTER(C.c_char))]


class MnistExample(C.Structure):       # lib/mnist_csv2.h:14-18
    _fields_ = [("X", C.POINTER(C.c_float)), ("y", C.c_float), ("num_examples", C.c_int)]


class MnistCSV1(C.Structure):          # lib/mnist_csv.h:6-10
    _fields_ = [("file", C.c_void_p), ("buffer", C.POINTER(C.c_float)), ("num_lines", C.c_int)]


class BMPData(C.Structure):            # lib/bmp.h:6-12
    _fields_ = [("width", C.c_uint), ("height", C.c_uint), ("red", C.c_void_p), ("green", C.c_void_p), ("blue", C.c_void_p)]


def bind_data(path):
    lib = C.CDLL(path)
    lib.mnist_csv_init.argtypes = [C.POINTER(MnistCSV2)]
    for f in (lib.get_random_data_take, lib.get_random_data_replace):
        f.restype = MnistExample
        f.argtypes = [C.POINTER(MnistCSV2)]
    lib.visualize_digit_data.argtypes = [MnistExample]
    lib.fill_random_data.argtypes = [C.c_int, C.c_void_p]
    lib.write_bmp_data.argtypes = [C.c_char_p, C.POINTER(BMPData)]
    return lib


@pytest.fixture(scope="module")
def ours():
    if not (os.path.exists(LIB) and os.path.exists(LIB_ROW)):
        import __graft_entry__
        __graft_entry__.build()
    return bind_data(LIB)


@pytest.fixture(scope="module")
def ref():
    if not os.path.exists(REF_DATA):
        pytest.skip("compiled reference (oracle/_ref) not present")
    return bind_data(REF_DATA)


def captured_stdout(fn, tmp_path):
    """what C code prints on fd 1 while fn runs"""
    path = str(tmp_path / "stdout.txt")
    libc.fflush(None)
    saved = os.dup(1)
    fd = os.open(path, os.O_WRONLY | os.O_CREAT | os.O_TRUNC)
    os.dup2(fd, 1)
    try:
        fn()
        libc.fflush(None)
    finally:
        os.dup2(saved, 1)
        os.close(fd)
        os.close(saved)
    return open(path, "rb").read()


def write_mnist_csv(path, n, seed):
    rng = np.random.default_rng(seed)
    labels = rng.integers(0, 10, n)
    pixels = rng.integers(0, 256, (n, 784))
    with open(path, "w") as f:
        for i in range(n):
            f.write(str(labels[i]) + "," + ",".join(map(str, pixels[i])) + ",\n")
    return labels, pixels


def load(lib, path, tmp_path):
    csv = MnistCSV2()
    csv.file = libc.fopen(str(path).encode(), b"r")
    out = captured_stdout(lambda: lib.mnist_csv_init(C.byref(csv)), tmp_path)
    return csv, out


def index_of(csv, ex):
    return (C.addressof(ex.X.contents) - C.addressof(csv.X.contents)) // 4


@pytest.mark.parametrize("n", [3, 300, 5000])      # 5000: the threaded transpose
def test_mnist_loader_arrays(ours, ref, tmp_path, n):
    path = tmp_path / "mnist.csv"
    labels, pixels = write_mnist_csv(path, n, n)
    a, out_a = load(ours, path, tmp_path)
    r, out_r = load(ref, path, tmp_path)
    assert out_a == out_r == b"MNIST CSV file contents read!\n"
    assert a.num_examples == r.num_examples == n and a.num_sampled == 0
    Xa = np.ctypeslib.as_array(a.X, shape=(784, n)); Xr = np.ctypeslib.as_array(r.X, shape=(784, n))
    assert np.array_equal(Xa, Xr) and np.array_equal(Xa, pixels.T.astype(np.float32))           # feature-major, mnist_csv2.c:29
    assert np.array_equal(np.ctypeslib.as_array(a.y, shape=(n,)), labels.astype(np.float32))
    assert bytes(C.cast(a.sampled, C.POINTER(C.c_char * n)).contents) == b"\0" * n
    for p in (a.X, a.y, a.sampled, r.X, r.y, r.sampled):                                        # malloc'd: the caller frees them
        libc.free(C.cast(p, C.c_void_p))


@pytest.mark.parametrize("n,seed", [(1, 3), (2, 4), (50, 42), (1000, 7)])
def test_take_sequence_matches_the_reference_with_the_models_reset(ours, ref, tmp_path, n, seed):
    """get_random_data_take over two and a half passes, the flags cleared by the CALLER between epochs as mnist_nn.c:189-191
    does (the library has to notice), and once mid-epoch; plus the wrap the library does itself when every draw is used up."""
    path = tmp_path / "mnist.csv"
    write_mnist_csv(path, n, seed)
    seqs, flags = [], []
    for lib in (ours, ref):
        csv, _ = load(lib, path, tmp_path)
        libc.srand(seed)
        seq = []
        for epoch in range(2):
            C.memset(csv.sampled, 0, n); csv.num_sampled = 0                                     # the model's reset
            seq += [index_of(csv, lib.get_random_data_take(C.byref(csv))) for _ in range(n)]
        seq += [index_of(csv, lib.get_random_data_take(C.byref(csv))) for _ in range(n // 2)]    # the library's own wrap
        C.memset(csv.sampled, 0, n); csv.num_sampled = 0                                         # reset in the middle of a pass
        seq += [index_of(csv, lib.get_random_data_take(C.byref(csv))) for _ in range(n)]
        seqs.append(seq)
        flags.append((bytes(C.cast(csv.sampled, C.POINTER(C.c_char * n)).contents), csv.num_sampled))
    assert seqs[0] == seqs[1]
    assert flags[0] == flags[1]
    assert seqs[0][:n] == reference_take_sequence(n, n, seed)                                    # and the Python restatement


def test_replace_sequence_and_examples(ours, ref, tmp_path):
    path = tmp_path / "mnist.csv"
    labels, pixels = write_mnist_csv(path, 200, 9)
    got = []
    for lib in (ours, ref):
        csv, _ = load(lib, path, tmp_path)
        libc.srand(5)
        seq = []
        for _ in range(500):
            ex = lib.get_random_data_replace(C.byref(csv))
            i = index_of(csv, ex)
            assert ex.num_examples == 200 and ex.y == labels[i] and ex.X[3 * 200] == pixels[i, 3]   # pixel p at X[p * num_examples]
            seq.append(i)
        got.append(seq)
    assert got[0] == got[1] and len(set(got[0])) > 100


def test_digit_pictures(ours, ref, tmp_path):
    path = tmp_path / "mnist.csv"
    write_mnist_csv(path, 20, 11)
    outs = []
    for lib in (ours, ref):
        csv, _ = load(lib, path, tmp_path)
        libc.srand(2)
        ex = lib.get_random_data_take(C.byref(csv))
        outs.append(captured_stdout(lambda: lib.visualize_digit_data(ex), tmp_path))
    assert outs[0] == outs[1]
    lines = outs[0].split(b"\n")
    assert len(lines) == 32 and lines[0] == b"=" * 28 and lines[1].startswith(b"Data for digit ") and set(outs[0]) <= set(b"=Data for digit0123456789.: #\n")


def test_cifar_records(ours, ref, tmp_path):
    path = str(tmp_path / "data_batch.bin")
    rng = np.random.default_rng(3)
    blob = rng.integers(0, 256, 30730000, dtype=np.uint8)
    blob.tofile(path)
    for name, want in (("CIFAR10_NUM_EXAMPLES_PER_FILE", 10000), ("CIFAR10_LINE_LENGTH", 3073), ("CIFAR10_DATA_LENGTH", 3072),
                       ("CIFAR10_BATCH_FILE_SIZE", 30730000), ("CIFAR10_NUM_PIXELS", 1024), ("CIFAR10_EXAMPLE_DIM", 32)):
        assert C.c_uint.in_dll(ours, name).value == C.c_uint.in_dll(ref, name).value == want
    res = []
    for lib in (ours, ref):
        fd = os.open(path, os.O_RDONLY)
        libc.srand(77)
        arr = np.empty((6, 3072), np.uint8)
        offs = []
        for k in range(6):
            lib.fill_random_data(fd, arr[k].ctypes.data)
            offs.append(os.lseek(fd, 0, os.SEEK_CUR))
        os.close(fd)
        res.append((arr, offs))
    assert np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1]
    # known answer: the record the draw names, every plane with its rows bottom-up (cifar10.c:24-31)
    rec = (res[0][1][0] - 3073) // 3073
    want = blob[rec * 3073 + 1:(rec + 1) * 3073].reshape(3, 32, 32)[:, ::-1, :].reshape(-1)
    assert np.array_equal(res[0][0][0], want)


@pytest.mark.parametrize("w,h", [(32, 32), (5, 3), (1, 1), (64, 7)])
def test_bmp_bytes(ours, ref, tmp_path, w, h):
    rng = np.random.default_rng(w * 100 + h)
    planes = [rng.integers(0, 256, w * h, dtype=np.uint8) for _ in range(3)]
    d = BMPData(w, h, planes[0].ctypes.data, planes[1].ctypes.data, planes[2].ctypes.data)
    files = []
    for tag, lib in (("a", ours), ("r", ref)):
        p = tmp_path / f"{tag}.bmp"
        lib.write_bmp_data(str(p).encode(), C.byref(d))
        files.append(bytearray(p.read_bytes()))
    stride = (24 * w + 31) // 32 * 4
    assert len(files[0]) == len(files[1]) == 54 + stride * h
    files[1][47] = 0                     # bmp.c never assigns info-header byte 33: whatever was on its stack
    assert files[0] == files[1]
    head = bytes(files[0][:54])
    assert head[:2] == b"BM" and struct.unpack_from("<IHHI", head, 2) == (54 + stride * h, 0, 0, 54)
    assert struct.unpack_from("<IiiHHII", head, 14) == (40, w, h, 1, 24, 0, 0)
    row0 = bytes(files[0][54:54 + stride])
    assert row0[:3] == bytes([planes[2][0], planes[1][0], planes[0][0]]) and row0[3 * w:] == b"\0" * (stride - 3 * w)
    # no O_TRUNC (bmp.c:12): a longer file keeps its tail
    p = tmp_path / "long.bmp"
    p.write_bytes(b"\xAA" * (54 + stride * h + 10))
    ours.write_bmp_data(str(p).encode(), C.byref(d))
    assert p.read_bytes() == bytes(files[0]) + b"\xAA" * 10


def test_row_reader_of_the_hinge_model(tmp_path):
    """lib/mnist_csv.h: get_next_data row by row and the 0.32 / 0.6 picture, against lib/mnist_csv.c"""
    path = tmp_path / "rows.csv"
    labels, pixels = write_mnist_csv(path, 7, 21)
    libs = [C.CDLL(LIB_ROW)] + ([C.CDLL(REF_ROW)] if os.path.exists(REF_ROW) else [])
    rows, pics, rets = [], [], []
    for lib in libs:
        lib.get_next_data.argtypes = [C.POINTER(MnistCSV1)]
        lib.visualize_digit_data.argtypes = [C.POINTER(MnistCSV1)]
        buf = np.zeros(785, np.float32)
        csv = MnistCSV1(libc.fopen(str(path).encode(), b"r"), buf.ctypes.data_as(C.POINTER(C.c_float)), 7)
        got, ret = [], []
        for _ in range(7):
            ret.append(lib.get_next_data(C.byref(csv)))
            got.append(buf.copy())
        buf[1:] = np.linspace(0, 1, 784, dtype=np.float32)
        pics.append(captured_stdout(lambda: lib.visualize_digit_data(C.byref(csv)), tmp_path))
        libc.fclose(csv.file)
        rows.append(np.array(got)); rets.append(ret)
    want = np.concatenate([labels[:, None], pixels], axis=1).astype(np.float32)
    for r, ret in zip(rows, rets):
        assert np.array_equal(r, want) and ret == [0] * 7
    assert len(set(pics)) == 1 and pics[0].count(b"\n") == 31 and b"#" in pics[0] and b":" in pics[0]


def test_init_commands_of_the_programs_built_without_reference_objects(tmp_path):
    """`init` of mnist_nn and mnist_hinge only draws parameters with rand() and writes the checkpoint -- no GPU work -- so the
    programs built from the unchanged model sources with NO reference file under lib/ (oracle/_ref/bin/bla_only_*) run here
    and must leave byte-identical checkpoints to the reference's own builds."""
    import subprocess
    bins = os.path.join(REF_DIR, "bin")
    pairs = (("ref_mnist_nn_f64_b512", "bla_only_mnist_nn_b512", "mnist_nn"), ("ref_mnist_hinge_f32", "bla_only_mnist_hinge", "mnist_hinge"))
    if not all(os.path.exists(os.path.join(bins, b)) for p in pairs for b in p[:2]):
        pytest.skip("oracle/_ref programs not built")
    for ref_bin, our_bin, sub in pairs:
        dirs = []
        for tag, binary in (("r", ref_bin), ("o", our_bin)):
            d = tmp_path / f"{sub}_{tag}"
            (d / "data" / sub).mkdir(parents=True)
            subprocess.run([os.path.join(bins, binary), "init"], cwd=d, check=True, capture_output=True, timeout=120)
            dirs.append(d / "data" / sub)
        names = sorted(os.listdir(dirs[0]))
        assert names and names == sorted(os.listdir(dirs[1]))
        for f in names:
            assert (dirs[0] / f).read_bytes() == (dirs[1] / f).read_bytes(), (sub, f)
