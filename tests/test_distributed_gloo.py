"""CPU test of the multi-GPU plumbing with world_size 2 over gloo: slot sharding covers every slot
exactly once and the counter all-reduce / max-over-ranks give every rank the global result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neural_rx_b200.distributed import error_counters, max_over_ranks, slot_shard, sum_counters


def test_slot_shard_partitions():
    for n in (0, 1, 7, 30, 31):
        for world in (1, 2, 4, 8):
            blocks = [slot_shard(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slot_shard(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)                       # same data on every rank, sharded by slot
    llr = rng.standard_normal((9, 2, 64)).astype(np.float32)
    bits = rng.integers(0, 2, (9, 2, 64)).astype(np.uint8)
    act = np.ones((9, 2), np.float32)
    act[4, 1] = 0
    lo, hi = slot_shard(9, rank, world)
    tot = sum_counters(error_counters(llr[lo:hi], bits[lo:hi], act[lo:hi]))
    q.put((rank, tot, max_over_ranks(1.0 + rank)))
    dist.destroy_process_group()


def test_counters_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    llr = rng.standard_normal((9, 2, 64)).astype(np.float32)
    bits = rng.integers(0, 2, (9, 2, 64)).astype(np.uint8)
    act = np.ones((9, 2), np.float32)
    act[4, 1] = 0
    want = error_counters(llr, bits, act)
    for _, tot, mx in res:
        assert tot == want and tot["slots"] == 9
        assert mx == 2.0


# ---- BLER harness (neural_rx_b200/bler.py): the counters of an Eb/N0 point do not depend on the number of ranks ----
def _bler_setup():
    from neural_rx_b200 import tb as TB
    from neural_rx_b200.config import get_config
    from neural_rx_b200.pusch import build_grid
    from oracle import nrx_oracle as O
    from tests.common import get_weights, oracle_arch, oracle_net
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=4)
    weights, _ = get_weights(cfg)
    enc = TB.pusch_tb_encoder(cfg, grid, base_graph="standin")
    dec = TB.TBDecoder(enc, num_bp_iter=8)
    net, arch = oracle_net(cfg, weights), oracle_arch(cfg)

    def llr_fn(y, act):                                    # the CPU oracle stands in for the CUDA receiver
        return np.asarray(O.receiver_forward(net, arch, y, grid.pilots, grid.pilot_mask, act)["llr"], np.float32)
    return cfg, grid, enc, dec, llr_fn


def _bler_worker(rank, world, port, q):
    from neural_rx_b200.bler import sim_point
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, grid, enc, dec, llr_fn = _bler_setup()
    tot, _ = sim_point(llr_fn, cfg, grid, enc, dec, point=1, ebno_db=3.0, num_slots=7, batch=2, rank=rank, world=world)
    q.put((rank, tot))
    dist.destroy_process_group()


def test_bler_point_counters_do_not_depend_on_world_size():
    from neural_rx_b200.bler import sim_point
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bler_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg, grid, enc, dec, llr_fn = _bler_setup()
    want, first = sim_point(llr_fn, cfg, grid, enc, dec, point=1, ebno_db=3.0, num_slots=7, batch=3, keep_first=2)
    assert want["blocks"] == 14 and want["bits"] == 14 * enc.tb_size and len(first) == 2
    assert 0 <= want["block_errors"] <= 14
    for _, tot in res:
        assert tot == want


def test_bler_point_single_active_user():
    """`active = [1, 0]` (the 1-UE evaluation of results/nrx_rt_results): only the active transmitter's blocks count."""
    from neural_rx_b200.bler import sim_point
    cfg, grid, enc, dec, llr_fn = _bler_setup()
    tot, first = sim_point(llr_fn, cfg, grid, enc, dec, point=0, ebno_db=6.0, num_slots=3, batch=2, keep_first=1, active=[1, 0])
    assert tot["blocks"] == 3 and tot["bits"] == 3 * enc.tb_size and 0 <= tot["block_errors"] <= 3
    assert np.array_equal(first[0][1], np.array([1, 0], np.float32))
