"""Multi-GPU plumbing: independent slots are sharded over ranks (one process per GPU); the only
collective is a SUM / MAX all-reduce of a few counters.  The reference's only
multi-GPU hook is ``distribute="all"`` of Sionna's ``sim_ber`` (``scripts/evaluate.py:61,199``),
i.e. batch sharding — this module is its torch.distributed equivalent (NCCL on GPUs, gloo in the
CPU tests)."""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np


def slot_shard(num_slots: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of slot indices owned by ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(num_slots, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device_for_backend():
    import torch
    import torch.distributed as dist
    if dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def sum_counters(counters: Dict[str, int]) -> Dict[str, int]:
    """All-reduce (SUM) of integer counters such as {bit_errors, bits, block_errors, blocks}."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(counters)
    keys = sorted(counters)
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.tolist())}


def max_over_ranks(value: float) -> float:
    """Max over ranks of a device-measured duration (the time every multi-GPU number is quoted on)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=_device_for_backend())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def error_counters(llr: np.ndarray, bits: np.ndarray, active_tx: np.ndarray) -> Dict[str, int]:
    """Uncoded hard-decision counters of one shard (``llr > 0 <=> bit 1``, utils/neural_rx.py:864)."""
    n = llr.shape[-1]
    err = ((llr > 0).astype(np.uint8) != bits[..., :n]) & (active_tx[..., None] > 0)
    return {"bit_errors": int(err.sum()), "bits": int((active_tx > 0).sum()) * n,
            "slots": int(llr.shape[0])}
