"""Monte-Carlo loop of the BLER harness — what Sionna's ``sim_ber`` does for the reference in
``scripts/evaluate.py:193-202`` (batch 30, ``distribute="all"``, early stop on a block-error target), restated for
slot-sharded ranks: slot i of SNR point p is generated from seeds that depend on (p, i) only, the slots of a point
are split over the ranks (``distributed.slot_shard``), every rank pushes its share through ``llr_fn`` (the CUDA
receiver; the oracle in CPU tests) and the transport-block decoder, and the only collective is a SUM of
{bit_errors, bits, block_errors, blocks} — so the totals do not depend on the number of ranks as long as no early
stop triggers.  Host-side harness code; nothing on the receiver hot path imports it."""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy as np

from .distributed import slot_shard, sum_counters
from .synth import make_slots


def coded_slots(cfg, grid, enc, idx, point: int, ebno_db: float, active=None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Slots ``idx`` of SNR point ``point``: seeded TB bits [n, U, tb_size], received grids y, activity flags
    (``active`` [U] 0/1: the transmitters that are on in every slot, e.g. [1, 0] for the single-UE results of
    ``results/nrx_rt_results``; default all)."""
    U = grid.num_tx
    act1 = None if active is None else np.asarray(active, np.float32).reshape(1, U)
    tbs, ys, acts = [], [], []
    for i in idx:
        rng = np.random.default_rng([77, point, i])
        b = rng.integers(0, 2, (1, U, enc.tb_size), dtype=np.uint8)
        sb = make_slots(cfg, grid, batch=1, ebno_db=float(ebno_db), seed=1000 * point + i, coded_bits=enc(b), active=act1)
        tbs.append(b)
        ys.append(sb.y)
        acts.append(sb.active_tx)
    return np.concatenate(tbs), np.concatenate(ys), np.concatenate(acts)


def sim_point(llr_fn: Callable, cfg, grid, enc, dec, point: int, ebno_db: float, num_slots: int, batch: int = 30,
              rank: int = 0, world: int = 1, target_block_errors: int = 500, keep_first: int = 0, active=None
              ) -> Tuple[Dict[str, int], List[tuple]]:
    """One Eb/N0 point.  ``llr_fn(y, active_tx) -> llr [n, U, num_coded_bits]`` (NumPy or torch; the decoder runs where
    the LLRs live).  Returns the job-wide counters and, on rank 0, the first ``keep_first`` (y, active, crc_ok, b_hat)
    tuples for cross-checks."""
    s_lo, s_hi = slot_shard(num_slots, rank, world)
    c = {"bit_errors": 0, "bits": 0, "block_errors": 0, "blocks": 0}
    first: List[tuple] = []
    n_iter = -(-(-(-num_slots // world)) // batch)      # the same on every rank: the early-stop test is a collective
    for it in range(n_iter):
        idx = range(min(s_lo + it * batch, s_hi), min(s_lo + (it + 1) * batch, s_hi))
        if len(idx):
            tb_bits, y, act = coded_slots(cfg, grid, enc, idx, point, ebno_db, active)
            b_hat, ok = dec(llr_fn(y, act))
            if hasattr(b_hat, "cpu"):
                b_hat, ok = b_hat.cpu().numpy(), ok.cpu().numpy()
            on = act > 0
            c["bit_errors"] += int(((b_hat != tb_bits) & on[..., None]).sum())
            c["bits"] += int(on.sum()) * enc.tb_size
            c["block_errors"] += int((~ok & on).sum())
            c["blocks"] += int(on.sum())
            if rank == 0 and len(first) < keep_first:
                k = keep_first - len(first)
                first += list(zip(y[:k], act[:k], ok[:k], b_hat[:k]))
        if sum_counters(dict(c))["block_errors"] >= target_block_errors:
            break
    return sum_counters(c), first
