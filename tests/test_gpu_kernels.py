"""Per-kernel GPU parity on RANDOM tensors at edge-case widths: the aggregation kernel, the
read-out kernel and the fused sep-conv stack kernels (serial plan 1 and pipelined plan 5), each driven through
the C ABI test hooks (nrx_debug_*) and compared with the oracle's restatement of the reference block —
``AggregateUserStates`` (utils/neural_rx.py:135-207), ``ReadoutLLRs`` / ``ReadoutChEst`` (:309-404), ``StateInit``
(:61-132), ``UpdateState`` (:210-270) — which tests/test_oracle_pins.py and tests/test_ref_e2e_pins.py pin to
the reference's own code.  F = 5, 9, 10 are narrower than / equal to one 9-subcarrier tile (+1), 1584 and 3276
are the 132- and 273-PRB carriers: a tile-edge or chunk-edge bug cannot hide behind an end-to-end tolerance.

Tolerances: <= 4e-3 relative L2 against the oracle with the engine's rounding points emulated (fp16 operands,
fp32 accumulation), <= 1e-2 against exact fp32; plans 1 and 5 bit-identical."""
import dataclasses

import numpy as np
import pytest

from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import PuschGrid
from neural_rx_b200.weights import random_weights
from oracle import nrx_oracle as O
from tests.common import ENGINE_EMU, oracle_arch, oracle_net, rel_l2

pytestmark = pytest.mark.gpu

WIDTHS = [5, 9, 10, 1584, 3276]
T = 14


def _setup(F, U=2, seed=0):
    """nrx_rt architecture with seeded random weights on a synthetic F-subcarrier grid (no PUSCH structure is
    needed by these kernels: positional encoding random, pilot tables empty)."""
    from neural_rx_b200.engine import NrxEngine
    cfg = dataclasses.replace(get_config("nrx_rt"), max_num_tx=U, dmrs_port_sets=[[0], [2], [1], [3]][:U])
    cfg.validate()
    rng = np.random.default_rng(100 + F)
    mask = np.zeros((T, F), bool)
    mask[[2, 11]] = True
    data_index = np.full(T * F, -1, np.int32)
    data_index[~mask.reshape(-1)] = np.arange(int((~mask).sum()), dtype=np.int32)
    grid = PuschGrid(U, F, T, (2, 11), np.zeros((U, 2 * F), np.complex64), mask, np.zeros((U, T * F), np.int32),
                     rng.standard_normal((U, F, T, 2)).astype(np.float32), data_index, int((~mask).sum()), 1)
    weights = random_weights(cfg, seed=seed)
    return cfg, grid, weights, NrxEngine(cfg, weights, grid, device=0), rng


def _f16(a):
    return np.asarray(a, np.float32).astype(np.float16)


def _state_rows(s56, pe):
    """[B,U,F,T,56] fp16 values + pe [U,F,T,2] -> internal state rows [B,U,F,T,64]."""
    B = s56.shape[0]
    rows = np.zeros(s56.shape[:-1] + (64,), np.float16)
    rows[..., :56] = s56
    rows[..., 56:58] = _f16(np.broadcast_to(pe[None], (B,) + pe.shape))
    return rows


@pytest.mark.parametrize("F", WIDTHS)
@pytest.mark.parametrize("U,active", [(2, [[1, 1], [1, 0]]), (3, [[1, 1, 1], [0, 1, 1]])])
def test_aggregation_kernel(F, U, active):
    import torch
    cfg, grid, weights, eng, rng = _setup(F, U)
    B = 2
    s = _f16(1.5 * rng.standard_normal((B, U, F, T, 56)))
    act = np.asarray(active, np.float32)
    rows = _state_rows(s, grid.pos_enc)
    a = eng.debug_aggregate(0, torch.as_tensor(rows).cuda(), torch.as_tensor(act).cuda())
    torch.cuda.synchronize()
    a = a.cpu().numpy().astype(np.float32)
    net = oracle_net(cfg, weights)
    st, at = torch.as_tensor(s.astype(np.float32)), torch.as_tensor(act)
    with torch.no_grad():
        emu = O.aggregate_user_states(net["it"][0][0], st, at, ENGINE_EMU).numpy()
        ref = O.aggregate_user_states(net["it"][0][0], st, at).numpy()
    assert np.all(a[..., 56:] == 0)
    assert rel_l2(a[..., :56], emu) <= 4e-3
    assert rel_l2(a[..., :56], ref) <= 1e-2
    eng.close()


@pytest.mark.parametrize("F", WIDTHS)
@pytest.mark.parametrize("skip", [0, 1])
def test_pipelined_aggregation_equals_one_tile_kernel(F, skip):
    """nrx_agg_ws_kernel (two users, TMA ring + warp roles) against nrx_agg_kernel<2>: same arithmetic, same bits;
    slots with both / one / no active user, with and without inactive-user skipping, more tiles than CTAs at F = 3276."""
    import torch
    cfg, grid, weights, eng, rng = _setup(F, 2)
    B = 7 if F >= 1584 else 5
    act = np.asarray([[1, 1], [1, 0], [1, 1], [0, 1], [0, 0], [1, 1], [1, 1]][:B], np.float32)
    rows = torch.as_tensor(_state_rows(_f16(1.5 * rng.standard_normal((B, 2, F, T, 56))), grid.pos_enc)).cuda()
    act_d = torch.as_tensor(act).cuda()
    eng.set_skip_inactive(bool(skip))
    outs = []
    for pipelined in (1, 0):
        eng.debug_option(eng.OPT_AGG_PIPELINED, pipelined)
        for it in (0, 1):
            outs.append(eng.debug_aggregate(it, rows, act_d).cpu().numpy())
    torch.cuda.synchronize()
    for it in (0, 1):
        new, old = outs[it], outs[2 + it]
        keep = np.ones((B, 2), bool) if not skip else (act > 0)   # planes of skipped users are not written
        assert np.array_equal(new[keep].view(np.uint16), old[keep].view(np.uint16))
    assert not np.array_equal(outs[0], outs[1])
    eng.close()


@pytest.mark.parametrize("F", WIDTHS)
def test_readout_kernel(F):
    import torch
    cfg, grid, weights, eng, rng = _setup(F)
    B, U = 2, 2
    s = _f16(1.5 * rng.standard_normal((B, U, F, T, 56)))
    llr, h = eng.debug_readout(0, torch.as_tensor(_state_rows(s, grid.pos_enc)).cuda(), cfg.num_bits_per_symbol[0])
    torch.cuda.synchronize()
    net = oracle_net(cfg, weights)
    st = torch.as_tensor(s.astype(np.float32))
    with torch.no_grad():
        for got, head in ((llr, net["llr"][0]), (h, net["chest"])):
            assert rel_l2(got.cpu().numpy(), O.mlp(st, head, ENGINE_EMU).numpy()) <= 4e-3
            assert rel_l2(got.cpu().numpy(), O.mlp(st, head).numpy()) <= 1e-2
    eng.close()


def _run_stack(net_stack, z, emu):
    import torch
    with torch.no_grad():
        for l in net_stack[:-1]:
            z = O.sepconv(z, l, True, emu)
        return O.sepconv(z, net_stack[-1], False, emu)


@pytest.mark.parametrize("F", WIDTHS)
def test_update_stack_kernels(F):
    """UpdateState = concat [a, s, pe] -> three sep-convs -> + s (:249-270) on random a / s, both stack kernels."""
    import torch
    cfg, grid, weights, eng, rng = _setup(F)
    B, U = (1, 2) if F > 1000 else (2, 2)
    s = _f16(rng.standard_normal((B, U, F, T, 56)))
    a = _f16(rng.standard_normal((B, U, F, T, 56)))
    s_rows = torch.as_tensor(_state_rows(s, grid.pos_enc)).cuda()
    a_rows = np.zeros((B, U, F, T, 64), np.float16)
    a_rows[..., :56] = a
    a_rows = torch.as_tensor(a_rows).cuda()
    outs = []
    for plan in (1, 5):
        eng.set_fused(plan)
        o = eng.debug_stack(0, B, a=a_rows, s=s_rows)
        torch.cuda.synchronize()
        outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    net = oracle_net(cfg, weights)
    pe = np.broadcast_to(_f16(grid.pos_enc).astype(np.float32)[None], (B, U, F, T, 2))
    z = torch.as_tensor(np.concatenate([a.astype(np.float32), s.astype(np.float32), pe], -1).reshape(B * U, F, T, 114))
    s32 = torch.as_tensor(s.astype(np.float32).reshape(B * U, F, T, 56))
    got = outs[0].astype(np.float32).reshape(B * U, F, T, 64)
    assert np.array_equal(got[..., 56:58], pe.reshape(B * U, F, T, 2))         # the encoding rides along unchanged
    emu = (s32 + _run_stack(net["it"][0][1], z, ENGINE_EMU)).numpy()
    assert rel_l2(got[..., :56], emu) <= 4e-3
    if F <= 1584:
        ref = (s32 + _run_stack(net["it"][0][1], z, O.EXACT)).numpy()
        assert rel_l2(got[..., :56], ref) <= 1e-2
    eng.close()


@pytest.mark.parametrize("F", WIDTHS)
def test_init_stack_kernels(F):
    """StateInit = concat [y, pe, h_hat] -> three sep-convs (:106-132); the kernels append pe to the state rows."""
    import torch
    cfg, grid, weights, eng, rng = _setup(F)
    B, U = (1, 2) if F > 1000 else (2, 2)
    z18 = _f16(rng.standard_normal((B, U, F, T, 18)))
    z0 = np.zeros((B, U, F, T, 32), np.float16)
    z0[..., :18] = z18
    outs = []
    for plan in (1, 5):
        eng.set_fused(plan)
        o = eng.debug_stack(-1, B, z0=torch.as_tensor(z0).cuda())
        torch.cuda.synchronize()
        outs.append(o.cpu().numpy())
    assert np.array_equal(outs[0], outs[1])
    net = oracle_net(cfg, weights)
    z = torch.as_tensor(z18.astype(np.float32).reshape(B * U, F, T, 18))
    got = outs[0].astype(np.float32).reshape(B * U, F, T, 64)
    pe = np.broadcast_to(_f16(grid.pos_enc).astype(np.float32)[None], (B, U, F, T, 2)).reshape(B * U, F, T, 2)
    assert np.array_equal(got[..., 56:58], pe) and np.all(got[..., 58:] == 0)
    assert rel_l2(got[..., :56], _run_stack(net["init"][0], z, ENGINE_EMU).numpy()) <= 4e-3
    if F <= 1584:
        assert rel_l2(got[..., :56], _run_stack(net["init"][0], z, O.EXACT).numpy()) <= 1e-2
    eng.close()
