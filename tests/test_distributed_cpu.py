"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: one broadcast of the packed weight
blob at load, batch sharding with no data-path collective, max-over-ranks reduction of timings."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import vad_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from vad_b200.distributed import broadcast_weight_blob, shard_bounds
        from vad_b200.engine import pack_state
        st = O.make_state(0, 64, 3, 128)
        numel = sum(v.numel() for v in st.values())
        blob = pack_state(st, 3) if rank == 0 else None
        got = broadcast_weight_blob(blob, numel, torch.device("cpu"), src=0)
        ref = pack_state(st, 3)
        ok_blob = torch.equal(got, ref)
        # shard a batch of 7 clips: every clip owned by exactly one rank, results gathered by the test
        B, T = 7, 16
        x = O.make_input(3, B, T, 64)
        lo, hi = shard_bounds(B, world, rank)
        part = O.forward_prob(st, x[lo:hi]) if hi > lo else torch.zeros(0, T)
        # timing reduction the bench uses: max over ranks
        tms = torch.tensor([float(rank + 1)])
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        q.put((rank, ok_blob, lo, hi, part.numpy(), float(tms.item())))
    finally:
        dist.destroy_process_group()


def test_broadcast_and_shard_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(o[1] for o in out)                      # every rank holds rank 0's weights
    assert out[0][2] == 0 and out[-1][3] == 7 and out[0][3] == out[1][2]
    st = O.make_state(0, 64, 3, 128)
    full = O.forward_prob(st, O.make_input(3, 7, 16, 64)).numpy()
    import numpy as np
    np.testing.assert_array_equal(np.concatenate([o[4] for o in out]), full)   # shards tile the batch
    assert all(o[5] == 2.0 for o in out)               # max over ranks
