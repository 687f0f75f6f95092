"""GPU tests of the checkpoint path: device-resident parameters <-> the reference's CSV files (lib/csv.c format,
model/mnist_nn.c:30-35 and model/cifar_unet.c:1484-1802 file layouts), staged through pinned memory by csrc/csv_codec.cu."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import load_ref, ptr, ref_available
from test_csv_cpu import py_format
import test_unet_gpu as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bla():
    import bla_b200 as b
    assert b.bla_device_count() >= 1
    return b


def six_decimals(a):
    return np.array([np.float32(float("%f" % float(v))) for v in a.ravel()], np.float32).reshape(a.shape)


def test_mlp_checkpoint_round_trip_in_the_reference_format(bla, tmp_path):
    b = bla
    dims = (C.c_int * 4)(784, 256, 128, 10)
    rng = np.random.default_rng(3)
    shapes = ((256, 784), (256,), (128, 256), (128,), (10, 128), (10,))
    p32 = [np.ascontiguousarray(rng.normal(0, 0.05, s), np.float32) for s in shapes]
    net = b.bla_mlp_create(dims, 64)
    b.bla_mlp_set_params(net, *[ptr(p) for p in p32])
    d0 = b.bla_d2h_bytes()
    b.bla_mlp_save_csv(net, str(tmp_path).encode())
    assert b.bla_d2h_bytes() - d0 == sum(p.nbytes for p in p32)           # device -> pinned -> text
    for l in range(3):
        w, bias = p32[2 * l], p32[2 * l + 1]
        assert (tmp_path / f"weights_{l + 1}.csv").read_bytes() == py_format(w)                    # one row per output unit
        assert (tmp_path / f"biases_{l + 1}.csv").read_bytes() == py_format(bias.reshape(-1, 1))   # mnist_nn.c:127: cols = 1
    if ref_available("f32"):                                                # the reference's own reader sees the same values
        ref = load_ref("f32")
        ref.read_csv_contents.restype = C.POINTER(C.c_float)
        ref.read_csv_contents.argtypes = [C.c_char_p]
        got = np.ctypeslib.as_array(ref.read_csv_contents(str(tmp_path / "weights_1.csv").encode()), shape=(256 * 784,))
        assert np.array_equal(got, six_decimals(p32[0]).ravel())
    net2 = b.bla_mlp_create(dims, 64)
    b.bla_mlp_load_csv(net2, str(tmp_path).encode())
    back = [np.empty_like(p) for p in p32]
    b.bla_mlp_get_params(net2, *[ptr(g) for g in back])
    for g, p in zip(back, p32):
        assert np.array_equal(g, six_decimals(p))
    b.bla_mlp_destroy(net); b.bla_mlp_destroy(net2)


def test_unet_checkpoint_uses_the_reference_directory_layout(bla, tmp_path):
    b = bla
    cfg = T.SMALL
    net, tensors = T.make_net(b, cfg, 1)
    b.bla_unet_init_params(net, 9)
    n = b.bla_unet_num_params(net)
    flat = np.empty(n, np.float32)
    b.bla_unet_get_params(net, ptr(flat))
    b.bla_unet_save_csv(net, str(tmp_path).encode())
    k2 = cfg["kernel_size"] ** 2
    by_name = {name: (off, cnt) for name, off, cnt in tensors}
    # cifar_unet.c:1493-1545: conv kernels as [F*C][k*k], residual conv = conv_3.csv, down / up convs = conv_0.csv
    off, cnt = by_name["down_1/resnet_1/conv_1"]
    assert (tmp_path / "down_1/resnet_1/conv_1.csv").read_bytes() == py_format(flat[off:off + cnt].reshape(-1, k2))
    off, cnt = by_name["down_1/resnet_1/residual_conv"]
    assert (tmp_path / "down_1/resnet_1/conv_3.csv").read_bytes() == py_format(flat[off:off + cnt].reshape(-1, 1))
    off, cnt = by_name["down_1/conv"]
    assert (tmp_path / "down_1/conv_0.csv").read_bytes() == py_format(flat[off:off + cnt].reshape(-1, k2))
    off, cnt = by_name["mid/self_attention/qkv"]
    qkv = flat[off:off + cnt].reshape(-1, 48)
    for i, nm in enumerate(("query", "key", "value")):
        assert (tmp_path / f"mid/{nm}.csv").read_bytes()   # the middle block's attention files lie in mid/ itself (:1601-1603) == py_format(np.ascontiguousarray(qkv[:, 16 * i:16 * i + 16]))
    off, cnt = by_name["up_4/resnet_2/time_weight"]
    assert (tmp_path / "up_4/resnet_2/time_weight.csv").read_bytes() == py_format(flat[off:off + cnt].reshape(cfg["time_dim"], -1))
    assert os.path.exists(tmp_path / "output_conv.csv") and os.path.exists(tmp_path / "up_3/self_attention_2/bias.csv")
    net2, _ = T.make_net(b, cfg, 1)
    b.bla_unet_load_csv(net2, str(tmp_path).encode())
    back = np.empty(n, np.float32)
    b.bla_unet_get_params(net2, ptr(back))
    for name, off, cnt in tensors:
        assert np.array_equal(back[off:off + cnt], six_decimals(flat[off:off + cnt])), name
    b.bla_unet_destroy(net); b.bla_unet_destroy(net2)


REF_BIN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "bin")
# blocks whose declared in_channels in save_parameters / load_parameters is smaller than what forward() uses (cifar_unet.c:1557,
# :1614-1653): conv_1.csv / conv_3.csv hold the first declared channels, the rest travels in *_rest.csv
TRUNCATED = {"down_1/resnet_2": (3, 128), "up_1/resnet_1": (256, 512), "up_2/resnet_1": (256, 512), "up_3/resnet_1": (256, 512),
             "up_4/resnet_1": (128, 256)}


def _files(root):
    out = {}
    for d, _, names in os.walk(root):
        for nm in names:
            out[os.path.relpath(os.path.join(d, nm), root)] = os.path.join(d, nm)
    return out


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "ref_cifar_unet_f64")), reason="compiled reference programs not present")
def test_unet_checkpoint_written_by_the_reference_loads_and_is_written_back_identically(bla, tmp_path):
    """`ref_cifar_unet_f64 init` (cifar_unet.c:1853-1858: init_parameters + save_parameters) writes data/cifar_unet; the library loads
    that directory into the reference-size model and writes it out again: every file the reference wrote comes back byte for byte
    (the unused residual kernels of in == out blocks excepted: same shape, zeros), and the only additional files are the *_rest.csv
    of the five blocks whose declared in_channels drop input channels."""
    import subprocess
    b = bla
    ref_dir = tmp_path / "ref"
    (ref_dir / "data").mkdir(parents=True)
    subprocess.run([os.path.join(REF_BIN, "ref_cifar_unet_f64"), "init"], cwd=ref_dir, check=True, timeout=300, stdout=subprocess.DEVNULL)
    theirs = _files(ref_dir / "data" / "cifar_unet")
    assert len(theirs) == 122
    net, tensors = T.make_net(b, T.FULL, 1)
    b.bla_unet_init_params(net, 5)
    n = b.bla_unet_num_params(net)
    before = np.empty(n, np.float32); b.bla_unet_get_params(net, ptr(before))
    b.bla_unet_load_csv(net, str(ref_dir / "data" / "cifar_unet").encode())
    after = np.empty(n, np.float32); b.bla_unet_get_params(net, ptr(after))
    by_name = {name: (off, cnt) for name, off, cnt in tensors}
    # a truncated block: the declared channels come from the file, the others keep what the model held
    off, cnt = by_name["up_4/resnet_1/conv_1"]
    w_before, w_after = before[off:off + cnt].reshape(128, 256, 9), after[off:off + cnt].reshape(128, 256, 9)
    text = np.array(open(theirs["up_4/resnet_1/conv_1.csv"]).read().replace("\n", "").split(",")[:-1], np.float64).astype(np.float32)
    assert np.array_equal(w_after[:, :128], text.reshape(128, 128, 9))
    assert np.array_equal(w_after[:, 128:], w_before[:, 128:])
    out_dir = tmp_path / "ours"
    b.bla_unet_save_csv(net, str(out_dir).encode())
    ours = _files(out_dir)
    extra = sorted(set(ours) - set(theirs))
    assert sorted(set(theirs) - set(ours)) == []
    assert extra == sorted(f"{blk}/{f}_rest.csv" for blk in TRUNCATED for f in ("conv_1", "conv_3") if not (blk == "down_1/resnet_2" and f == "conv_3"))
    unused = 0
    for rel, path in theirs.items():
        a, o = open(path, "rb").read(), open(ours[rel], "rb").read()
        blk = os.path.dirname(rel)
        never_read = (rel.endswith("conv_3.csv") and (blk + "/residual_conv") not in by_name) or \
                     (rel.endswith("conv_0.csv") and (blk + "/conv") not in by_name)
        if never_read:   # kernels forward() never reads (:1062-1066 residual of in == out blocks, :1131,:1141 skipped up convs)
            assert a.count(b"\n") == o.count(b"\n") and set(o) <= set(b"0.,\n"), rel
            unused += 1
        else:
            assert a == o, rel
    blocks = [nm[:-len('/conv_1')] for nm in by_name if nm.endswith('/conv_1')]
    assert unused == sum(1 for blk in blocks if blk + '/residual_conv' not in by_name) + 2   # + up_1 / up_2 conv_0
    b.bla_unet_destroy(net)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_BIN, "ref_unet_ckpt_f64")), reason="compiled reference programs not present")
def test_unet_checkpoint_written_here_goes_through_the_reference_loader(bla, tmp_path):
    """The other direction: a checkpoint of the reference-size model written by bla_unet_save_csv is read by the reference's
    load_parameters and written back by its save_parameters (oracle/build_ref.sh: ref_unet_ckpt_f64 calls exactly those two,
    cifar_unet.c:1545-1802); every file the reference knows is unchanged byte for byte, and loading the result here restores all
    parameters (the *_rest.csv files the reference left alone included)."""
    import subprocess
    b = bla
    net, tensors = T.make_net(b, T.FULL, 1)
    b.bla_unet_init_params(net, 6)
    n = b.bla_unet_num_params(net)
    flat = np.empty(n, np.float32); b.bla_unet_get_params(net, ptr(flat))
    work = tmp_path / "w"
    ck = work / "data" / "cifar_unet"
    ck.mkdir(parents=True)
    b.bla_unet_save_csv(net, str(ck).encode())
    written = {rel: open(path, "rb").read() for rel, path in _files(ck).items()}
    subprocess.run([os.path.join(REF_BIN, "ref_unet_ckpt_f64")], cwd=work, check=True, timeout=600, stdout=subprocess.DEVNULL)
    back = {rel: open(path, "rb").read() for rel, path in _files(ck).items()}
    assert set(back) == set(written)
    for rel in written:
        assert back[rel] == written[rel], rel
    net2, _ = T.make_net(b, T.FULL, 1)
    b.bla_unet_load_csv(net2, str(ck).encode())
    got = np.empty(n, np.float32); b.bla_unet_get_params(net2, ptr(got))
    for name, off, cnt in tensors:
        assert np.array_equal(got[off:off + cnt], six_decimals(flat[off:off + cnt])), name
    b.bla_unet_destroy(net); b.bla_unet_destroy(net2)


def test_codec_edge_cases(bla, tmp_path):
    """empty file, a single value without a terminator (dropped, as lib/csv.c:44-52 does), rows without the trailing comma (the
    reference overflows its buffer there, SURVEY D8: here every value is returned), zero-sized writes"""
    b = bla
    out = C.POINTER(C.c_float)()
    assert b.bla_csv_parse(b"", 0, C.byref(out)) == 0
    assert b.bla_csv_parse(b"42", 2, C.byref(out)) == 0
    n = b.bla_csv_parse(b"1,2\n3,4\n", 8, C.byref(out))
    assert n == 4 and [out[i] for i in range(4)] == [1.0, 2.0, 3.0, 4.0]
    p = tmp_path / "empty.csv"
    v = np.zeros((0, 4), np.float32)
    b.write_csv_contents(str(p).encode(), ptr(np.zeros(1, np.float32)), 4, 0)
    assert p.read_bytes() == b""
    d = b.bla_malloc_device(16)
    b.bla_csv_save(str(tmp_path / "one.csv").encode(), d, 1, 0)
    b.bla_free(d)
