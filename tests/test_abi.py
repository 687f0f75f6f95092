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


def exported():
    out = subprocess.check_output(["nm", "-D", "--defined-only", LIB], text=True)
    return {line.split()[-1] for line in out.splitlines() if line.strip()}


@pytest.fixture(scope="module")
def syms():
    if not os.path.exists(LIB):
        import __graft_entry__
        __graft_entry__.build()
    return exported()


def test_every_declared_function_is_exported(syms):
    decl = declared_functions()
    assert len(decl) > 60
    missing = sorted(decl - syms)
    assert not missing, f"declared in include/ but not exported by libbla.so: {missing}"


def test_reference_import_sets_are_covered(syms):
    """What model/*.o need from lib/{matrix,layer,conv,norm,util}.o;
    csv / mnist_csv / cifar10 / bmp symbols keep coming from the reference's own host objects."""
    need = """make_matrix clone_matrix free_matrix_data free_matrix matrix_multiply matrix_scale matrix_add print_matrix
    print_matrix_dim matrix_multiply_elementwise matrix_transpose matrix_row_sum matrix_col_sum frobenius_norm max_value
    matrix_z_score_normalize matrix_add_tile_columns matrix_add_tile_rows matrix_multiply_inplace feed_forward
    free_layer_data load_weights_from_csv load_biases_from_csv back_propagate_errors do_back_propagate_errors conv
    conv_ddx reshape_channels_matrix reshape_matrix_channels _im2col _col2im _reshape_kernels_matrix
    _reshape_matrix_kernels group_norm group_norm_ddx epsilon relu softmax softmax_row_wise load_matrix_from_csv
    random_gaussian PI""".split()
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
