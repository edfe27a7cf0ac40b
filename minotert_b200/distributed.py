"""Multi-GPU exchange steps.  One process per GPU, torch.distributed for the plumbing.

The render itself never communicates: every rank holds the whole BVH, LUTs and blue noise and renders
either its own row slabs of the image (tile mode, mrt_set_partition) or the whole image for its own
frame counters (sample mode).  The only exchange is of the finished framebuffer / accumulator:

  tile mode   : ranks own interleaved slabs of `slab_rows` rows (slab j -> rank j % N), stored compactly.
                gather_tiles() collects the compact slabs on the root and scatters the rows back into
                image order using the same mapping the library uses (mrt_partition_rows_for).
  sample mode : reduce_samples() sums the fp32 RGBA accumulators (xyz = radiance sums, w = sample count)
                onto the root; dividing by w afterwards averages over all ranks' samples.

Works with the nccl backend on device tensors (NVLink/NVSwitch) and with gloo on CPU tensors (tests).
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

from . import capi


def partition_rows(rank, nranks, slab_rows, full_h):
    """Rows of the full image owned by `rank`, in its local storage order (host logic, no GPU needed)."""
    L = capi.load()
    n = C.c_uint32()
    if L.mrt_partition_rows_for(rank, nranks, slab_rows, full_h, None, C.byref(n)) != 0:
        raise ValueError(f"bad partition {rank}/{nranks} slab {slab_rows}")
    rows = np.zeros(n.value, np.uint32)
    if n.value:
        L.mrt_partition_rows_for(rank, nranks, slab_rows, full_h, rows.ctypes.data_as(C.c_void_p), C.byref(n))
    return rows


_ROW_CACHE = {}


def _rows_on_device(world, slab_rows, full_h, device):
    """(counts, one int64 row-index tensor per rank on `device`), cached: the mapping is static per partition."""
    key = (world, slab_rows, full_h, str(device))
    if key not in _ROW_CACHE:
        rows = [partition_rows(r, world, slab_rows, full_h) for r in range(world)]
        _ROW_CACHE[key] = ([len(x) for x in rows],
                           [torch.from_numpy(x.astype(np.int64)).to(device) for x in rows])
    return _ROW_CACHE[key]


def gather_tiles(local, full_h, slab_rows, dst=0, group=None):
    """local: [local_rows, W, C] tensor of this rank's slabs.  Returns the [full_h, W, C] image on dst, None elsewhere."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts, row_idx = _rows_on_device(world, slab_rows, full_h, local.device)
    assert local.shape[0] == counts[rank], f"rank {rank}: {local.shape[0]} local rows, expected {counts[rank]}"
    pad = max(counts)
    buf = local
    if local.shape[0] != pad:  # equal-size gather: pad the short ranks
        buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        buf[: local.shape[0]] = local
    buf = buf.contiguous()
    if rank == dst:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, parts, dst=dst, group=group)
        full = torch.empty((full_h,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        for r in range(world):
            full[row_idx[r]] = parts[r][: counts[r]]
        return full
    dist.gather(buf, None, dst=dst, group=group)
    return None


def reduce_samples(accum, dst=0, group=None):
    """In-place sum of fp32 RGBA accumulators onto dst (w channel carries the sample counts)."""
    dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return accum


def device_tensor(ptr, nbytes, dtype, device):
    """Zero-copy torch view of a borrowed device buffer returned by mrt_buffer."""
    item = torch.empty((), dtype=dtype).element_size()
    typestr = {torch.float32: "<f4", torch.uint8: "|u1", torch.uint32: "<u4", torch.float16: "<f2"}[dtype]

    class _Wrap:
        __cuda_array_interface__ = {"shape": (nbytes // item,), "typestr": typestr, "data": (ptr, False), "version": 2}

    return torch.as_tensor(_Wrap(), device=device)
