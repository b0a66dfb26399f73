"""BLER harness on the GPU: TB bits -> TBEncoder -> synthetic slot -> NeuralPUSCHReceiver on the
CUDA engine -> TBDecoder, through the reference's receiver call ``(b_hat, h_hat_refined, h_hat, tb_crc_status)``
(utils/neural_rx.py:1600-1603).  The LDPC code is the structural stand-in of neural_rx_b200/tb.py (the TS 38.212
shift tables are not in the image), so the checks are about the LLR source, not about 3GPP BLER values: blocks
decode and match the transmitted bits at high Eb/N0, fail at very low Eb/N0, and at the waterfall the CUDA LLRs
lead to the same CRC decisions as the oracle's LLRs for the same slots."""
import numpy as np
import pytest

from neural_rx_b200 import tb as TB
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from oracle import nrx_oracle as O
from tests.common import get_weights, oracle_arch, oracle_net

pytestmark = pytest.mark.gpu


def _slots(cfg, grid, enc, n, ebno, seed0):
    rng = np.random.default_rng([5, seed0])
    bits = rng.integers(0, 2, (n, grid.num_tx, enc.tb_size), dtype=np.uint8)
    coded = enc(bits)
    sbs = [make_slots(cfg, grid, batch=1, ebno_db=ebno, seed=seed0 + i, coded_bits=coded[i:i + 1]) for i in range(n)]
    return bits, np.concatenate([s.y for s in sbs]), np.concatenate([s.active_tx for s in sbs])


def test_receiver_call_returns_decoded_blocks():
    from neural_rx_b200.receiver import NeuralPUSCHReceiver
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=8)
    weights, kind = get_weights(cfg)
    real = kind == "shipped"
    rx = NeuralPUSCHReceiver(cfg, weights=weights, grid=grid, tb_decoding="standin")
    enc = rx.tb_encoders[0]
    assert enc.base_graph_source == "standin" and enc.num_coded_bits == grid.num_data_res * 4
    bits, y, act = _slots(cfg, grid, enc, 4, 12.0 if real else 40.0, 300)
    b_hat, h_ref, h_ls, status = rx((y, act), None)
    assert b_hat.shape == bits.shape and status.shape == act.shape and status.dtype == bool
    if real:                                    # the shipped weights demodulate; random ones need not
        assert status.all() and np.array_equal(b_hat, bits)
    bits, y, act = _slots(cfg, grid, enc, 2, -8.0, 400)
    b_hat, _, _, status = rx((y, act), None)
    assert not status.any()
    off = NeuralPUSCHReceiver(cfg, weights=weights, grid=grid, tb_decoding="off")
    llr, _, _, none = off((y, act), None)
    assert none is None and llr.shape == (2, grid.num_tx, enc.num_coded_bits)
    rx.engine.close()
    off.engine.close()


def test_crc_decisions_match_the_oracle_at_the_waterfall():
    import torch
    from neural_rx_b200.engine import NrxEngine
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=8)
    weights, kind = get_weights(cfg)
    real = kind == "shipped"
    if not real:
        pytest.skip("needs the shipped weight file")
    enc = TB.pusch_tb_encoder(cfg, grid, base_graph="standin")
    dec = TB.TBDecoder(enc)
    eng = NrxEngine(cfg, weights, grid, device=0)
    net, arch = oracle_net(cfg, weights), oracle_arch(cfg)
    n_ok = n_diff = n = 0
    for ebno, seed0 in ((1.0, 500), (3.0, 600)):
        bits, y, act = _slots(cfg, grid, enc, 4, ebno, seed0)
        llr = eng.forward(torch.as_tensor(y).cuda(), torch.as_tensor(act).cuda(), want=("llr",))["llr"]
        b_g, ok_g = dec(llr)                                    # decoder on the GPU
        b_g, ok_g = b_g.cpu().numpy(), ok_g.cpu().numpy()
        ref = O.receiver_forward(net, arch, y, grid.pilots, grid.pilot_mask, act)["llr"]
        b_o, ok_o = dec(np.asarray(ref, np.float32))            # the same decoder on the CPU
        assert np.array_equal(b_g[ok_g], bits[ok_g])            # a passed CRC means the right bits
        n_diff += int((ok_g != ok_o).sum())
        n_ok += int(ok_g.sum())
        n += ok_g.size
    assert 0 < n_ok < n                                         # these points straddle the waterfall
    assert n_diff <= 1
    eng.close()
