"""big-linear-algebra_b200 -- Python-side mirror of the C-ABI of libbla.so (ctypes only).

The product is the shared library next to this file (built in-tree from csrc/*.cu for sm_100a by
`make -C big-linear-algebra_b200` / __graft_entry__.build()).  This module only declares the
prototypes of include/lib/*.h (the reference's own API) and include/bla.h (the additive
device-resident API) so tests and bench.py can call through the same boundary a relinked C
program uses.  There is no Python or CPU implementation of any operation here: if the library is
missing, import fails loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbla.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `make -C {_HERE}` (nvcc, sm_100a). "
        "There is no CPU or PyTorch fallback for the big-linear-algebra hot path.")

lib = C.CDLL(LIB_PATH)   # RTLD_LOCAL: the test-only reference build exports the same names

c_float_p = C.POINTER(C.c_float)


class Matrix(C.Structure):
    """struct Matrix of include/lib/matrix.h (lib/matrix.h:6-11 in the reference)."""
    _fields_ = [("rows", C.c_int), ("cols", C.c_int), ("data", c_float_p)]


MatrixP = C.POINTER(Matrix)


class ConvData(C.Structure):
    """struct ConvData of include/lib/conv.h (lib/conv.h:6-11)."""
    _fields_ = [("im2col", MatrixP), ("kernel_matrix", MatrixP), ("product", MatrixP), ("output", MatrixP)]


ACT_FN = C.CFUNCTYPE(None, c_float_p, C.c_int)


class Layer(C.Structure):
    """struct Layer of include/lib/layer.h (lib/layer.h:4-15)."""


Layer._fields_ = [("num_nodes", C.c_int), ("nodes", MatrixP), ("raw_nodes", MatrixP), ("weights", MatrixP),
                  ("biases", MatrixP), ("previous_layer", C.POINTER(Layer)), ("activation", ACT_FN),
                  ("activation_ddx", ACT_FN), ("has_previous_layer", C.c_char), ("has_nodes", C.c_char)]


class UnetConfig(C.Structure):
    """include/bla.h bla_unet_config (model/cifar_unet.c:26-37)"""
    _fields_ = [("image_side", C.c_int), ("dims", C.c_int * 4), ("time_dim", C.c_int), ("kernel_size", C.c_int),
                ("group_size", C.c_int), ("key_dim", C.c_int), ("dropout", C.c_float), ("max_imgs", C.c_int),
                ("seed", C.c_ulonglong)]


class Epilogue(C.Structure):
    """bla_epilogue of include/bla.h."""
    _fields_ = [("bias_rows", C.c_void_p), ("bias_cols", C.c_void_p), ("pre_activation", C.c_void_p),
                ("gate", C.c_void_p), ("activation", C.c_int), ("alpha", C.c_float)]


GEMM_FP32, GEMM_3XTF32, GEMM_AUTO = 0, 1, 2
ACT_IDENTITY, ACT_RELU, ACT_SCALE = 0, 1, 2
KIND_HOST, KIND_MANAGED, KIND_DEVICE, KIND_PINNED = 0, 1, 2, 3

# name -> (restype, argtypes); exactly the symbols include/lib/*.h and include/bla.h declare
PROTOTYPES = {
    # include/lib/matrix.h
    "make_matrix": (MatrixP, [C.c_int, C.c_int, C.c_void_p]),
    "clone_matrix": (MatrixP, [Matrix]),
    "free_matrix_data": (None, [MatrixP]),
    "free_matrix": (None, [MatrixP]),
    "matrix_multiply": (MatrixP, [Matrix, Matrix]),
    "matrix_scale": (None, [MatrixP, C.c_float]),
    "matrix_add": (None, [MatrixP, MatrixP]),
    "print_matrix": (None, [Matrix]),
    "print_matrix_dim": (None, [Matrix]),
    "matrix_multiply_elementwise": (None, [MatrixP, MatrixP]),
    "matrix_transpose": (None, [MatrixP]),
    "matrix_row_sum": (MatrixP, [Matrix]),
    "matrix_col_sum": (MatrixP, [Matrix]),
    "frobenius_norm": (C.c_float, [Matrix]),
    "max_value": (C.c_float, [Matrix]),
    "matrix_z_score_normalize": (None, [MatrixP]),
    "matrix_add_tile_columns": (None, [MatrixP, MatrixP]),
    "matrix_add_tile_rows": (None, [MatrixP, MatrixP]),
    "matrix_multiply_inplace": (None, [MatrixP, MatrixP, MatrixP]),
    # include/lib/util.h
    "relu": (None, [C.c_void_p, C.c_int]),
    "softmax": (None, [C.c_void_p, C.c_int, C.c_int]),
    "softmax_row_wise": (None, [C.c_void_p, C.c_int, C.c_int]),
    "load_matrix_from_csv": (None, [MatrixP, C.c_char_p, C.c_int, C.c_int]),
    "random_gaussian": (C.c_double, [C.c_void_p]),
    # include/lib/layer.h
    "feed_forward": (None, [C.POINTER(Layer)]),
    "free_layer_data": (None, [Layer]),
    "load_weights_from_csv": (None, [C.POINTER(Layer), C.c_char_p]),
    "load_biases_from_csv": (None, [C.POINTER(Layer), C.c_char_p]),
    "back_propagate_errors": (None, [C.POINTER(Layer), c_float_p, C.c_float]),
    "do_back_propagate_errors": (None, [C.POINTER(Layer), C.POINTER(Layer), MatrixP, C.c_float]),
    # include/lib/conv.h
    "conv": (None, [MatrixP, C.POINTER(MatrixP), C.POINTER(ConvData), C.c_int, C.c_int, C.c_int]),
    "reshape_channels_matrix": (None, [MatrixP, MatrixP]),
    "reshape_matrix_channels": (None, [MatrixP, MatrixP]),
    "conv_ddx": (None, [MatrixP, C.POINTER(ConvData), C.POINTER(ConvData), C.POINTER(MatrixP), MatrixP, C.c_int, C.c_int]),
    "_im2col": (None, [MatrixP, MatrixP, C.c_int, C.c_int, C.c_int]),
    "_col2im": (None, [MatrixP, MatrixP, C.c_int, C.c_int, C.c_int]),
    "_reshape_kernels_matrix": (None, [C.POINTER(MatrixP), MatrixP]),
    "_reshape_matrix_kernels": (None, [MatrixP, C.POINTER(MatrixP)]),
    # include/lib/norm.h
    "group_norm": (None, [MatrixP, MatrixP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "group_norm_ddx": (None, [MatrixP, MatrixP, MatrixP, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    # include/bla.h -- runtime
    "bla_device_count": (C.c_int, []),
    "bla_init": (None, [C.c_int]),
    "bla_sync": (None, []),
    "bla_stream": (C.c_void_p, []),
    "bla_set_stream": (None, [C.c_void_p]),
    "bla_version": (C.c_char_p, []),
    "bla_set_gemm_path": (None, [C.c_int]),
    "bla_get_gemm_path": (C.c_int, []),
    "bla_tc_available": (C.c_int, []),
    "bla_tc_launch_count": (C.c_ulonglong, []),
    "bla_tc_main_columns": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "bla_set_quirks": (None, [C.c_int]),
    "bla_get_quirks": (C.c_int, []),
    "bla_launch_count": (C.c_ulonglong, []),
    "bla_h2d_bytes": (C.c_ulonglong, []),
    "bla_d2h_bytes": (C.c_ulonglong, []),
    # include/bla.h -- memory
    "bla_malloc_device": (C.c_void_p, [C.c_size_t]),
    "bla_malloc_pinned": (C.c_void_p, [C.c_size_t]),
    "bla_malloc_managed": (C.c_void_p, [C.c_size_t]),
    "bla_free": (None, [C.c_void_p]),
    "bla_memory_kind": (C.c_int, [C.c_void_p]),
    "bla_matrix_device": (MatrixP, [C.c_int, C.c_int]),
    "bla_copy_h2d": (None, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bla_copy_d2h": (None, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bla_copy_d2d": (None, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "bla_memset_zero": (None, [C.c_void_p, C.c_size_t]),
    "bla_fill_uniform": (None, [C.c_void_p, C.c_size_t, C.c_ulonglong, C.c_float, C.c_float]),
    "bla_host_uniform": (None, [C.c_void_p, C.c_size_t, C.c_ulonglong, C.c_float, C.c_float]),
    "bla_u8_to_float": (None, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_float]),
    # include/bla.h -- GEMM and device twins of model-local loops
    "bla_gemm": (None, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "bla_gemm_ex": (None, [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(Epilogue)]),
    "bla_relu_ddx": (None, [C.c_void_p, C.c_int]),
    "bla_relu_backward": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "bla_softmax_xent": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "bla_conv2d_forward": (None, [C.c_void_p] * 3 + [C.c_int] * 7),
    "bla_conv2d_wgrad": (None, [C.c_void_p] * 3 + [C.c_int] * 7),
    "bla_conv2d_dgrad": (None, [C.c_void_p] * 3 + [C.c_int] * 7),
    "bla_group_norm": (None, [C.c_void_p] * 4 + [C.c_int] * 4),
    "bla_group_norm_ddx": (None, [C.c_void_p] * 5 + [C.c_int] * 4),
    # include/bla.h -- data pipeline
    "bla_mnist_from_csv": (C.c_void_p, [C.c_char_p]),
    "bla_mnist_from_arrays": (C.c_void_p, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "bla_mnist_destroy": (None, [C.c_void_p]),
    "bla_mnist_num_examples": (C.c_int, [C.c_void_p]),
    "bla_mnist_reset": (None, [C.c_void_p]),
    "bla_mnist_sample_take": (None, [C.c_void_p, C.c_int, C.c_void_p]),
    "bla_mnist_gather": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "bla_mlp_train_epoch": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "bla_hinge_create": (C.c_void_p, [C.c_int, C.c_int, C.c_int]),
    "bla_hinge_destroy": (None, [C.c_void_p]),
    "bla_hinge_set_weights": (None, [C.c_void_p, C.c_void_p]),
    "bla_hinge_get_weights": (None, [C.c_void_p, C.c_void_p]),
    "bla_hinge_iteration": (None, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]),
    "bla_cifar_open": (C.c_void_p, [C.c_char_p]),
    "bla_cifar_destroy": (None, [C.c_void_p]),
    "bla_cifar_num_examples": (C.c_int, [C.c_void_p]),
    "bla_cifar_sample": (None, [C.c_void_p, C.c_int, C.c_void_p]),
    "bla_cifar_gather": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    # include/lib/csv.h + include/bla.h -- CSV checkpoint codec
    "read_csv_contents": (C.POINTER(C.c_float), [C.c_char_p]),
    "write_csv_contents": (None, [C.c_char_p, C.c_void_p, C.c_int, C.c_int]),
    "bla_csv_parse": (C.c_size_t, [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(C.c_float))]),
    "bla_csv_format": (C.c_size_t, [C.c_void_p, C.c_int, C.c_size_t, C.c_char_p, C.c_size_t]),
    "bla_csv_save": (None, [C.c_char_p, C.c_void_p, C.c_int, C.c_size_t]),
    "bla_csv_load": (None, [C.c_char_p, C.c_void_p, C.c_size_t]),
    # include/bla.h -- fused self attention
    "bla_attention_forward": (None, [C.c_void_p] * 9 + [C.c_int] * 3),
    "bla_attention_backward": (None, [C.c_void_p] * 11 + [C.c_int] * 3),
    # include/bla.h -- CIFAR U-Net trainer
    "bla_unet_create": (C.c_void_p, [C.POINTER(UnetConfig)]),
    "bla_unet_destroy": (None, [C.c_void_p]),
    "bla_unet_num_params": (C.c_size_t, [C.c_void_p]),
    "bla_unet_num_tensors": (C.c_int, [C.c_void_p]),
    "bla_unet_tensor_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "bla_unet_tensor_offset": (C.c_size_t, [C.c_void_p, C.c_int]),
    "bla_unet_tensor_size": (C.c_size_t, [C.c_void_p, C.c_int]),
    "bla_unet_init_params": (None, [C.c_void_p, C.c_ulonglong]),
    "bla_unet_set_params": (None, [C.c_void_p, C.c_void_p]),
    "bla_unet_get_params": (None, [C.c_void_p, C.c_void_p]),
    "bla_unet_get_grads": (None, [C.c_void_p, C.c_void_p]),
    "bla_unet_save_csv": (None, [C.c_void_p, C.c_char_p]),
    "bla_unet_load_csv": (None, [C.c_void_p, C.c_char_p]),
    "bla_unet_forward": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bla_unet_train_step": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    # include/bla.h -- MNIST MLP trainer
    "bla_mlp_create": (C.c_void_p, [C.POINTER(C.c_int), C.c_int]),
    "bla_mlp_destroy": (None, [C.c_void_p]),
    "bla_mlp_dims": (None, [C.c_void_p, C.c_void_p]),
    "bla_mlp_set_host_packing": (None, [C.c_void_p, C.c_int]),
    "bla_pack_pixels": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bla_mlp_set_params": (None, [C.c_void_p] * 7),
    "bla_mlp_get_params": (None, [C.c_void_p] * 7),
    "bla_mlp_save_csv": (None, [C.c_void_p, C.c_char_p]),
    "bla_mlp_load_csv": (None, [C.c_void_p, C.c_char_p]),
    "bla_mlp_init_params": (None, [C.c_void_p, C.c_ulonglong]),
    "bla_mlp_train_step": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "bla_mlp_train_step_u8": (None, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "bla_mlp_set_host_chunking": (None, [C.c_void_p, C.c_int]),
    "bla_mlp_forward": (None, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "bla_mlp_read_stats": (None, [C.c_void_p, C.c_void_p]),
    # include/bla.h -- NCCL
    "bla_comm_unique_id": (None, [C.c_void_p]),
    "bla_comm_init": (None, [C.c_void_p, C.c_int, C.c_int]),
    "bla_comm_world": (C.c_int, []),
    "bla_comm_peer_windows": (C.c_int, []),
    "bla_comm_set_peer_windows": (None, [C.c_int]),
    "bla_comm_rank": (C.c_int, []),
    "bla_allreduce_sum_f32": (None, [C.c_void_p, C.c_size_t]),
    "bla_allreduce_sum_f64": (None, [C.c_void_p, C.c_size_t]),
    "bla_broadcast_f32": (None, [C.c_void_p, C.c_size_t, C.c_int]),
    "bla_comm_destroy": (None, []),
}

MISSING = []
for _name, (_res, _args) in PROTOTYPES.items():
    try:
        _fn = getattr(lib, _name)
    except AttributeError:
        MISSING.append(_name)
        continue
    _fn.restype = _res
    _fn.argtypes = _args
    globals()[_name] = _fn


# ---- small conveniences for tests / bench (no arithmetic) ---------------------------------------
def host_matrix(arr):
    """View a C-contiguous float32 2-D numpy array as a by-value struct Matrix (ordinary host memory)."""
    assert arr.dtype == np.float32 and arr.ndim == 2 and arr.flags["C_CONTIGUOUS"]
    return Matrix(arr.shape[0], arr.shape[1], arr.ctypes.data_as(c_float_p))


def device_matrix_from(arr):
    """Upload a float32 2-D numpy array into a new HBM-resident Matrix (free with free_matrix)."""
    arr = np.ascontiguousarray(arr, np.float32)
    m = lib.bla_matrix_device(arr.shape[0], arr.shape[1])
    lib.bla_copy_h2d(C.cast(m.contents.data, C.c_void_p), arr.ctypes.data_as(C.c_void_p), arr.nbytes)
    lib.bla_sync()
    return m


def to_numpy(m):
    """Copy any Matrix (host, managed or device data) into a fresh numpy array."""
    mm = m.contents if isinstance(m, MatrixP) else m
    out = np.empty((mm.rows, mm.cols), np.float32)
    addr = C.cast(mm.data, C.c_void_p)
    if out.nbytes == 0:
        return out
    kind = lib.bla_memory_kind(addr)
    if kind in (KIND_HOST, KIND_PINNED):
        C.memmove(out.ctypes.data, addr, out.nbytes)
    else:
        lib.bla_copy_d2h(out.ctypes.data_as(C.c_void_p), addr, out.nbytes)
        lib.bla_sync()
    return out


def planes(arr3):
    """[C][H][W] float32 numpy -> (Matrix * C) array of planes sharing its memory."""
    assert arr3.dtype == np.float32 and arr3.flags["C_CONTIGUOUS"]
    Cn, H, W = arr3.shape
    out = (Matrix * Cn)()
    for c in range(Cn):
        out[c] = Matrix(H, W, arr3[c].ctypes.data_as(c_float_p))
    return out


def kernel_table(k4):
    """[F][C][k][k] float32 numpy -> Matrix** as conv()/conv_ddx() expect."""
    rows = [planes(k4[f]) for f in range(k4.shape[0])]
    table = (MatrixP * len(rows))(*[C.cast(r, MatrixP) for r in rows])
    table._keep = rows
    return table
