"""-m gpu: the reference's unchanged model programs built with NO reference file under lib/ at all (oracle/_ref/bin/bla_only_*):
every header from include/lib, csv / mnist_csv2 / cifar10 / bmp from libbla.so (csrc/host_io.cu), mnist_hinge's row reader from
libbla_mnist_csv.so.  The builds that keep the reference's host objects are compared with the reference itself in
test_programs_gpu.py; the host objects alone are compared with the compiled reference in test_host_io_cpu.py."""
import os
import re

import numpy as np
import pytest

from test_programs_gpu import have, mnist_csv, read_csv, run, twin_dirs, write_csv

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not have("bla_only_main", "bla_only_my_first_model", "bla_only_mnist_hinge", "bla_only_mnist_nn_b512", "bla_main",
                         "bla_my_first_model", "bla_mnist_hinge", "bla_mnist_nn_b512"), reason="oracle/_ref programs not built")
def test_programs_linked_against_this_library_alone(tmp_path):
    """the same unchanged model sources built with NO reference file under lib/ -- every header from
    include/lib, csv / mnist_csv2 / cifar10 / bmp from libbla.so (csrc/host_io.cu), mnist_hinge's row reader from
    libbla_mnist_csv.so -- against the builds that keep the reference's host objects (themselves compared with the reference
    above).  Same library underneath, so the same text and the same checkpoints; the host objects on their own are compared
    with the compiled reference call by call in tests/test_host_io_cpu.py."""
    rng = np.random.default_rng(4)

    def fill(d):          # the fixtures of the four tests above, in one tree
        for sub in ("mnist_nn", "mnist", "mnist_hinge", "my_first_model"):
            (d / "data" / sub).mkdir(parents=True)
        mnist_csv(d / "data" / "mnist" / "mnist_train.csv", 1024, 3)
        mnist_csv(d / "data" / "mnist" / "mnist_test.csv", 60, 4)
        (d / "data" / "a.csv").write_text("1,2.3,3,\n4,509,6,\n7,8,9.0,")
        (d / "data" / "inputs.csv").write_text("3,\n7,\n9,")
        (d / "data" / "weights.csv").write_text("1,2,3,\n4,5,6,")
        (d / "data" / "biases.csv").write_text("0.1,\n0.2,")
        m = d / "data" / "my_first_model"
        write_csv(m / "input_nodes.csv", [[0.7], [-0.3]])
        write_csv(m / "hidden_weights.csv", np.round(rng.normal(0, 0.7, (3, 2)), 6))
        write_csv(m / "hidden_biases.csv", np.round(rng.normal(0, 0.5, (3, 1)), 6))
        write_csv(m / "output_weights.csv", np.round(rng.normal(0, 0.7, (2, 3)), 6))
        write_csv(m / "output_biases.csv", np.round(rng.normal(0, 0.5, (2, 1)), 6))
    ra, rb = twin_dirs(tmp_path, fill)
    number = r"-?\d+\.\d+(?:e[-+]?\d+)?"

    def same_text(want, got):
        assert re.sub(number, "#", got) == re.sub(number, "#", want)
        assert np.allclose([float(x) for x in re.findall(number, got)], [float(x) for x in re.findall(number, want)], rtol=2e-3, atol=1e-5)   # an accuracy may move by one sample

    def same_files(sub):
        names = sorted(os.listdir(os.path.join(ra, "data", sub)))
        assert names and names == sorted(os.listdir(os.path.join(rb, "data", sub)))
        for f in names:
            assert np.allclose(read_csv(os.path.join(rb, "data", sub, f)), read_csv(os.path.join(ra, "data", sub, f)), rtol=1e-5, atol=2e-6), f

    same_text(run("bla_main", ra), run("bla_only_main", rb))
    for args in (("run",), ("train", "400", "0.01")):
        same_text(run("bla_my_first_model", ra, *args), run("bla_only_my_first_model", rb, *args))
    same_files("my_first_model")
    for args in (("init",), ("train", "10", "0.001"), ("run", "60", "1000")):
        same_text(run("bla_mnist_hinge", ra, *args), run("bla_only_mnist_hinge", rb, *args))
    same_files("mnist_hinge")
    for args in (("init",), ("train", "2")):
        same_text(run("bla_mnist_nn_b512", ra, *args), run("bla_only_mnist_nn_b512", rb, *args))
    same_files("mnist_nn")
