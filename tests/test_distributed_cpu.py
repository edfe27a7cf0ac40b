"""World-size-2 (and 3) gloo tests of the multi-GPU exchange logic on CPU tensors:
tile gather reassembles interleaved slabs into image order; sample reduce sums accumulators."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, full_h, w, slab, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from minotert_b200 import distributed as D
        rows = D.partition_rows(rank, world, slab, full_h)
        # tile mode: each local pixel encodes its global (row, column)
        local = torch.zeros((len(rows), w, 4), dtype=torch.uint8)
        for lr, y in enumerate(rows):
            local[lr, :, 0] = int(y) % 251
            local[lr, :, 1] = torch.arange(w) % 256
            local[lr, :, 2] = rank
        full = D.gather_tiles(local, full_h, slab, dst=0)
        ok = True
        if rank == 0:
            ys = torch.arange(full_h)
            ok &= bool(torch.equal(full[:, 0, 0].long(), ys % 251))
            ok &= bool(torch.equal(full[:, :, 1].long(), (torch.arange(w) % 256).expand(full_h, w)))
            ok &= bool(torch.equal(full[:, 0, 2].long(), (ys // slab) % world))
        else:
            ok &= full is None
        # sample mode: rank r contributes (r+1) everywhere with 2 samples
        acc = torch.full((full_h * w * 4,), float(rank + 1))
        acc.view(-1, 4)[:, 3] = 2.0
        D.reduce_samples(acc, dst=0)
        if rank == 0:
            a = acc.view(-1, 4)
            ok &= bool(torch.all(a[:, 0] == sum(range(1, world + 1)))) and bool(torch.all(a[:, 3] == 2.0 * world))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,full_h,slab", [(2, 64, 8), (2, 101, 8), (3, 50, 4)])
def test_gather_and_reduce_gloo(world, full_h, slab):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, full_h, 33, slab, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def test_partition_rows_cover_image():
    from minotert_b200 import distributed as D
    for world in (1, 2, 4, 8):
        rows = np.concatenate([D.partition_rows(r, world, 8, 2160) for r in range(world)])
        assert np.array_equal(np.sort(rows), np.arange(2160))
    with pytest.raises(ValueError):
        D.partition_rows(3, 2, 8, 64)


def _uid_worker(rank, world, port, q):
    """The plumbing bench.py --gpus N uses around mrt_group_create_rank: rank 0 makes the ncclUniqueId through the
    C ABI, the bytes travel by torch.distributed (gloo here, NCCL on the GPU box), every rank ends up with the same
    128 bytes and with complementary row tables.  (mrt_group_create_rank itself needs a GPU per rank.)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from minotert_b200 import capi
        from minotert_b200 import distributed as D
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(capi.Group.unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        rows = D.partition_rows(rank, world, 8, 2160)
        gathered = [None] * world
        dist.all_gather_object(gathered, (bytes(uid.numpy().tobytes()), rows.tolist()))
        ok = all(g[0] == gathered[0][0] for g in gathered) and any(b != 0 for b in gathered[0][0])
        allrows = np.sort(np.concatenate([np.array(g[1]) for g in gathered]))
        ok &= bool(np.array_equal(allrows, np.arange(2160)))
        # without a CUDA device the group cannot be made: it must fail loudly, not fall back
        try:
            capi.Group.rank_of(0, rank, world, gathered[0][0])
            ok &= torch.cuda.is_available()
        except capi.MinoteError as e:
            ok &= "no CPU fallback" in str(e) or "CUDA" in str(e)
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_group_unique_id_plumbing_gloo():
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by tests/test_gpu_group.py")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_uid_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results
