"""CPU tests: the oracle END TO END against the reference's own forward code.

``tests/golden/ref_e2e_fixtures.npz`` holds inputs and outputs of the reference's ``CGNN.forward``
(utils/neural_rx.py:544-595, running ``StateInit.forward`` :106-132, ``AggregateUserStates.forward``
:176-207, ``UpdateState.forward`` :249-270, ``CGNNIt.forward`` of utils/neural_rx copy_pytorch.py:311-321
and the read-outs :309-404), of ``NeuralReceiverONNX.forward`` (:1773-1812) and of
``post_process_llrs`` (utils/onnx_utils.py:472-516), executed in the build container by
``tests/golden/make_ref_e2e_fixtures.py`` with the shipped weights (fork defects routed around: see
that script's header).  Tolerance 1e-5 relative to the largest output magnitude (fp32 re-association
only; measured: 0 for the Sionna-shaped cases, 8e-7 for the Aerial-shaped one)."""
import os

import numpy as np
import pytest
import torch

from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.weights import load_weights
from oracle import nrx_oracle as O
from tests.common import oracle_arch, oracle_net, weight_path

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ref_e2e_fixtures.npz")
CASES = {"rt": "nrx_rt", "rt_it1": "nrx_rt", "varmcs": "nrx_rt_var_mcs", "masking": "nrx_large_var_mcs_64qam_masking",
         "large": "nrx_large", "large64": "nrx_large_64qam", "site": "nrx_site_specific_large"}
TOL = 1e-5


def _fx():
    return np.load(GOLDEN)


def _net(label):
    if weight_path(label) is None:
        pytest.skip(f"weights/{label}_weights not staged")
    cfg = get_config(label)
    w = load_weights(cfg, weight_path(label))
    return cfg, oracle_arch(cfg), oracle_net(cfg, w)


def _close(a, ref):
    return np.abs(np.asarray(a) - ref).max() <= TOL * np.abs(ref).max()


@pytest.mark.parametrize("key", sorted(CASES))
def test_cgnn_forward_matches_reference_forward(key):
    """oracle.cgnn_forward == reference CGNN.forward: normalisation broadcast, concat orders [y, pe, h_hat] /
    [a, s, pe], Var-IO blend, aggregation masks ([1,1], [1,0], [0,1]), residual, iteration loop and ``num_it``
    truncation, per-MCS heads and the masking-mode slice."""
    g = _fx()
    cfg, arch, net = _net(CASES[key])
    t = torch.as_tensor
    with torch.no_grad():
        llrs, h = O.cgnn_forward(net, arch, t(g[f"{key}_y8"]), t(g[f"{key}_pe"]), t(g[f"{key}_h_hat"]),
                                 t(g[f"{key}_active"]), t(g[f"{key}_mask"]), num_it=int(g[f"{key}_num_it"]))
    assert len(llrs) == len(cfg.num_bits_per_symbol)
    for m, l in enumerate(llrs):
        assert l.shape == g[f"{key}_llr{m}"].shape
        assert _close(l.numpy(), g[f"{key}_llr{m}"]), (key, m)
    assert _close(h.numpy(), g[f"{key}_h_ref"])


@pytest.mark.parametrize("key", sorted(CASES))
def test_receiver_forward_chain_matches_reference_forward(key):
    """The whole Sionna-shaped oracle chain from the COMPLEX grid (re|im re-layout, LS + FOCC + nearest-pilot
    estimate, positional encoding, CGNN, head selection) reproduces the reference-generated LLR grids."""
    g = _fx()
    cfg, arch, net = _net(CASES[key])
    grid = build_grid(cfg, n_size_bwp=int(g["n_prb"]))
    mask = g[f"{key}_mask"]
    for head in range(len(cfg.num_bits_per_symbol)):
        out = O.receiver_forward(net, arch, g[f"{key}_y"], grid.pilots, grid.pilot_mask, g[f"{key}_active"],
                                 mcs_arr_eval=[head], mcs_ue_mask_eval=mask, num_it=int(g[f"{key}_num_it"]))
        ref = g[f"{key}_llr{head}"]
        assert _close(out["llr_grid"][head], ref)
        assert _close(out["llr"], O.demap_llrs(ref, grid.pilot_mask))
    assert _close(out["h_hat_refined"], g[f"{key}_h_ref"])
    assert np.array_equal(out["h_hat"], g[f"{key}_h_hat"])


def test_aerial_forward_matches_reference_forward():
    """oracle.aerial_forward == reference NeuralReceiverONNX.forward: re|im concat, FOCC removal, per-PRB
    nearest-pilot gather, LLR layout [B,bits,U,F,T] and sign."""
    g = _fx()
    cfg, arch, net = _net("nrx_rt")
    ins = [g[f"aerial_in{i}"] for i in range(7)]
    out = O.aerial_forward(net, arch, *ins)
    assert out["llr"].shape == g["aerial_llr"].shape
    assert _close(out["llr"], g["aerial_llr"])
    assert _close(out["h_hat"], g["aerial_h"])


def test_demap_order_matches_reference_post_process_llrs():
    """oracle.demap_llrs == reference post_process_llrs (utils/onnx_utils.py:472-516) applied to the
    reference-generated Aerial LLR tensor."""
    g = _fx()
    grid = build_grid(get_config("nrx_rt"), n_size_bwp=int(g["n_prb"]))
    llr_sionna = -np.transpose(g["aerial_llr"], (0, 2, 3, 4, 1))            # [B,U,F,T,bits]
    assert np.array_equal(O.demap_llrs(llr_sionna, grid.pilot_mask), g["aerial_llr_demapped"])
