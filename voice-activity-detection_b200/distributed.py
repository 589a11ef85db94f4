"""Multi-GPU: one process per GPU, the clip batch sharded across ranks, ONE broadcast of the
packed weight blob at load (NCCL over NVLink/NVSwitch), no collective on the steady-state path
(SURVEY.md section 8e).  Replaces the reference's only multi-GPU mechanism, the training-time
``nn.DataParallel`` (vad/training/trainer.py:115-116), for inference.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of ``n_items`` clips: rank r owns [lo, hi); sizes differ by at most 1."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_assignment(lengths, world_size: int):
    """Mixed-length batches (config 5): greedy longest-first assignment of clips to ranks so that
    each rank gets a near-equal sum of T_b^2 (attention cost).  Returns a list of index lists."""
    order = sorted(range(len(lengths)), key=lambda i: -int(lengths[i]))
    load = [0] * world_size
    out = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda j: (load[j], j))
        out[r].append(i)
        load[r] += int(lengths[i]) ** 2
    return [sorted(ix) for ix in out]


def broadcast_weight_blob(blob: Optional[torch.Tensor], numel: int, device: torch.device,
                          src: int = 0) -> torch.Tensor:
    """Rank ``src`` passes the packed fp32 blob (any device); every rank returns the blob on
    ``device``.  A single collective, issued once at load time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert blob is not None
        return blob.to(device=device, dtype=torch.float32).contiguous()
    if dist.get_rank() == src:
        assert blob is not None and blob.numel() == numel
        buf = blob.to(device=device, dtype=torch.float32).contiguous()
    else:
        buf = torch.empty(numel, dtype=torch.float32, device=device)
    dist.broadcast(buf, src=src)
    return buf


def nccl_comm_ptr(device: torch.device) -> Optional[int]:
    """The raw ``ncclComm_t`` of the default process group for ``device`` (None when the backend is
    not NCCL or this torch build does not expose it)."""
    try:
        pg = dist.distributed_c10d._get_default_group()
        backend = pg._get_backend(device)
        ptr = backend._comm_ptr()
        return int(ptr) if ptr else None
    except Exception:
        return None


def load_engine_from_broadcast(engine, state_dict=None, src: int = 0, via: str = "auto"):
    """Rank ``src`` packs ``state_dict``; all ranks end up with the weights loaded -- ONE collective.

    ``via="c"``: rank ``src`` loads its blob and the library itself broadcasts it from handle to handle
    (``vadb_broadcast_weights`` on the process group's own ncclComm_t -- the path a maintainer binding only
    the C ABI would use); ``via="torch"``: ``dist.broadcast`` of the blob, then ``vadb_load_weights``
    (on_device=1) on every rank; ``"auto"``: the C entry when the NCCL communicator is reachable."""
    from .engine import pack_state
    multi = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
    is_src = (not multi) or dist.get_rank() == src
    blob = pack_state(state_dict, engine.num_layers) if is_src else None
    if multi and via in ("auto", "c") and dist.get_backend() == "nccl":
        # the communicator is created lazily by the first collective on this device
        dist.barrier(device_ids=[engine.device.index])
        comm = nccl_comm_ptr(engine.device)
        ok = torch.tensor([1 if comm else 0], device=engine.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)            # every rank must take the same path
        if int(ok.item()) == 1:
            if is_src:
                engine.load_blob(blob)
            engine.broadcast_weights(comm, src)
            engine.load_path = "vadb_broadcast_weights"
            return engine
        if via == "c":
            raise RuntimeError("the NCCL communicator of the default process group is not reachable")
    buf = broadcast_weight_blob(blob, engine.weight_count, engine.device, src)
    engine.load_blob(buf)
    engine.load_path = "dist.broadcast + vadb_load_weights"
    return engine
