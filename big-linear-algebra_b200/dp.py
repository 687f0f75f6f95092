"""Host-side data-parallel plumbing for the MLP step (no arithmetic on tensors here).

The reference lays a mini-batch out as COLUMNS (model/mnist_nn.c:199-201), so a data-parallel shard is a column
range; every gradient is a plain sum over samples and one all-reduce of the flat buffer restores the full-batch
gradient.  The only non-trivial piece is the reference's matrix_col_sum quirk (lib/matrix.c:138-148):
bias gradient i is the sum of the `global_batch` FLAT elements of the [rows x global_batch] matrix starting at
i*rows, i.e. at most two row segments; `quirk_window_segments` is the host twin of csrc/mlp.cu:bias_grad_kernel."""
import ctypes as C


def shard_columns(global_batch, world, rank):
    """Columns [offset, offset + count) of the global batch owned by `rank`; the last rank takes the remainder."""
    base = global_batch // world
    offset = rank * base
    count = global_batch - offset if rank == world - 1 else base
    return offset, count


def quirk_window_segments(i, rows, global_batch, col_offset, local_batch):
    """Local pieces of bias-gradient window i: a list of (row, local_col_begin, local_col_end)."""
    start = i * rows
    r0, s = divmod(start, global_batch)
    out = []
    lo, hi = max(s, col_offset), col_offset + local_batch
    if r0 < rows and lo < hi:
        out.append((r0, lo - col_offset, hi - col_offset))
    lo, hi = col_offset, min(s, col_offset + local_batch)
    if r0 + 1 < rows and lo < hi:
        out.append((r0 + 1, lo - col_offset, hi - col_offset))
    return out


def exchange_unique_id(bla, dist, rank, device=None):
    """Rank 0 creates the NCCL unique id through the C-ABI; it is broadcast with torch.distributed (any backend)
    and every rank returns the 128 raw bytes for bla_comm_init()."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        bla.bla_comm_unique_id(raw)
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, 0)
    return (C.c_ubyte * 128)(*buf.cpu().tolist())
