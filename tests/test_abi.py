"""CPU-side checks of the drop-in boundary: libbla.so loads without a GPU, exports every symbol that
include/lib/*.h and include/bla.h declare, and exports the full symbol set the reference's model
programs import (probed there with nm).  No compute call is made."""
import os
import re
import subprocess

import pytest

from helpers import ROOT

INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(ROOT, "big-linear-algebra_b200", "libbla.so")
LIB_MNIST_CSV = os.path.join(ROOT, "big-linear-algebra_b200", "libbla_mnist_csv.so")   # lib/mnist_csv.h, mnist_hinge.c's reader

DECL = re.compile(r"^[A-Za-z_][\w\s\*]*?\b(\w+)\s*\(", re.M)


def declared_functions():
    names = set()
    for dirpath, _, files in os.walk(INCLUDE):
        for f in files:
            if not f.endswith(".h"):
                continue
            text = open(os.path.join(dirpath, f)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            text = re.sub(r"//.*", "", text)
            for m in DECL.finditer(text):
                if m.group(1) not in ("defined", "if", "sizeof"):
                    names.add(m.group(1))
    return names


def exported(lib=LIB):
    out = subprocess.check_output(["nm", "-D", "--defined-only", lib], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


@pytest.fixture(scope="module")
def syms():
    if not os.path.exists(LIB) or not os.path.exists(LIB_MNIST_CSV):
        import __graft_entry__
        __graft_entry__.build()
    return exported()


def test_every_declared_function_is_exported(syms):
    decl = declared_functions()
    assert len(decl) > 60
    # lib/mnist_csv.h's two functions live in libbla_mnist_csv.so: its visualize_digit_data is not lib/mnist_csv2.h's
    row_reader = exported(LIB_MNIST_CSV)
    assert {"get_next_data", "visualize_digit_data"} <= row_reader
    missing = sorted(decl - syms - {"get_next_data"})
    assert not missing, f"declared in include/ but not exported by libbla.so: {missing}"


def test_reference_import_sets_are_covered(syms):
    """What model/*.o need from the reference's lib/*.o: the compute objects, and the host I/O
    objects too (csv, mnist_csv2, cifar10, bmp from libbla.so; mnist_csv from libbla_mnist_csv.so), so that a model program
    links with no reference object at all."""
    need = """make_matrix clone_matrix free_matrix_data free_matrix matrix_multiply matrix_scale matrix_add print_matrix
    print_matrix_dim matrix_multiply_elementwise matrix_transpose matrix_row_sum matrix_col_sum frobenius_norm max_value
    matrix_z_score_normalize matrix_add_tile_columns matrix_add_tile_rows matrix_multiply_inplace feed_forward
    free_layer_data load_weights_from_csv load_biases_from_csv back_propagate_errors do_back_propagate_errors conv
    conv_ddx reshape_channels_matrix reshape_matrix_channels _im2col _col2im _reshape_kernels_matrix
    _reshape_matrix_kernels group_norm group_norm_ddx epsilon relu softmax softmax_row_wise load_matrix_from_csv
    random_gaussian PI read_csv_contents read_csv_contents_file write_csv_contents count_num_lines mnist_csv_init
    get_random_data_take get_random_data_replace visualize_digit_data fill_random_data write_bmp_data
    CIFAR10_NUM_EXAMPLES_PER_FILE CIFAR10_LINE_LENGTH CIFAR10_DATA_LENGTH CIFAR10_BATCH_FILE_SIZE CIFAR10_NUM_PIXELS
    CIFAR10_EXAMPLE_DIM""".split()
    assert not [n for n in need if n not in syms]


def test_python_binding_matches_library(syms):
    import bla_b200 as b
    assert b.MISSING == []
    assert set(b.PROTOTYPES) <= syms
    assert isinstance(b.bla_device_count(), int)        # answers 0 here, never exits
    assert b.bla_version().startswith(b"bla-b200")


def test_struct_layouts_match_the_reference_headers():
    import ctypes as C
    import bla_b200 as b
    assert C.sizeof(b.Matrix) == 16 and b.Matrix.data.offset == 8            # lib/matrix.h:7-11
    assert C.sizeof(b.ConvData) == 32                                        # lib/conv.h:6-11
    assert b.Layer.previous_layer.offset == 40 and b.Layer.activation.offset == 48
    assert b.Layer.has_previous_layer.offset == 64 and C.sizeof(b.Layer) == 72   # lib/layer.h:4-15
