"""Pin the oracle against the REFERENCE'S OWN CODE, executed here.

    python tests/golden/make_ref_fixtures.py        (needs /root/reference; run in the build container)

The reference package cannot be imported (``utils/__init__.py`` pulls in TensorFlow and Sionna, which
are not installable offline) and its receiver does not run end to end.  Several of
its classes are nevertheless self-contained torch / NumPy code.  This script extracts exactly those
class definitions from the reference's source files with ``ast`` (nothing is copied into this repo),
executes them in a namespace that only provides torch / numpy, feeds them seeded inputs and stores
inputs + outputs as fixtures; ``tests/test_oracle_pins.py`` then checks the oracle against them:

* ``AggregateUserStates``            utils/neural_rx.py:135-207   (message MLP, masking, sum minus self, scaling)
* ``ReadoutLLRs`` / ``ReadoutChEst`` utils/neural_rx.py:309-404
* ``NearestNeighborInterpolator``    utils/neural_rx.py:919-1004  (nearest-pilot gather indices of the LS estimator)
* ``NRPreprocessing``                utils/neural_rx.py:1614-1700 (FOCC removal, per-PRB nearest-pilot template)
* positional encoding                utils/onnx_utils.py:203-247  (the NumPy part of the pre-computation inside
                                     ``NRXDataGenerator.__init__``: a code FRAGMENT, executed with the
                                     pilot-position list the preceding TF lines would have produced)
* ``SeparableConv2d``                utils/neural_rx copy_pytorch.py:34-51 (the torch twin of Keras
                                     SeparableConv2D the fork intended to use; that file is commented out,
                                     the class is un-commented on the fly)

Weights: the shipped ``nrx_rt_weights`` (Keras layouts: Dense kernel [in,out] -> ``nn.Linear.weight`` =
kernel^T; depthwise [3,3,C,1] -> ``Conv2d.weight[c,0,i,j]``; pointwise [1,1,Cin,Cout] ->
``Conv2d.weight[n,c,0,0]``).  Only inputs and outputs are stored, not the weights.
"""
import ast
import contextlib
import io
import os
import sys
import tempfile
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.weights import load_weights  # noqa: E402


def reference_classes(path, names, uncomment=False):
    """exec the named top-level class definitions of a reference source file."""
    text = open(path).read()
    if uncomment:
        text = "\n".join(l[2:] if l.startswith("# ") else (l[1:] if l.startswith("#") else l) for l in text.splitlines())
    tree = ast.parse(text)
    ns = {"torch": torch, "nn": nn, "F": F, "np": np, "os": os}
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):      # the fork's code prints debug lines
        return fn(*a, **k)


def main():
    rng = np.random.default_rng(20240901)
    live = reference_classes(os.path.join(REF, "utils", "neural_rx.py"),
                             ["AggregateUserStates", "ReadoutLLRs", "ReadoutChEst", "NearestNeighborInterpolator",
                              "NRPreprocessing"])
    twin = reference_classes(os.path.join(REF, "utils", "neural_rx copy_pytorch.py"), ["SeparableConv2d"], uncomment=True)
    cfg = get_config("nrx_rt")
    w = load_weights(cfg, os.path.join(REF, "weights", "nrx_rt_weights")).to_list()
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)
    out = {}

    # ---- AggregateUserStates of iteration 0: arrays 9..12 ----------------------
    agg = live["AggregateUserStates"](56, [64], 56)
    with torch.no_grad():
        agg._hidden_layers[0].weight.copy_(t(w[9].T)); agg._hidden_layers[0].bias.copy_(t(w[10]))
        agg._output_layer.weight.copy_(t(w[11].T)); agg._output_layer.bias.copy_(t(w[12]))
    s = (2.0 * rng.standard_normal((4, 3, 5, 14, 56))).astype(np.float32)
    act = np.array([[1, 1, 1], [1, 0, 1], [0, 1, 0], [0, 0, 0]], np.float32)
    with torch.no_grad():
        a = quiet(agg, [t(s), t(act)[:, :, None, None]])
    out.update(agg_s=s, agg_active=act, agg_a=a.numpy())

    # ---- read-outs: nrx_rt arrays 35..38 (LLR head) and 39..42 (channel estimate) -------------------
    ro = live["ReadoutLLRs"](4, [128], 56)
    ch = live["ReadoutChEst"](4, [128], 56)
    with torch.no_grad():
        ro._hidden_layers[0].weight.copy_(t(w[35].T)); ro._hidden_layers[0].bias.copy_(t(w[36]))
        ro._output_layer.weight.copy_(t(w[37].T)); ro._output_layer.bias.copy_(t(w[38]))
        ch._hidden_layers[0].weight.copy_(t(w[39].T)); ch._hidden_layers[0].bias.copy_(t(w[40]))
        ch._output_layer.weight.copy_(t(w[41].T)); ch._output_layer.bias.copy_(t(w[42]))
    s2 = (1.5 * rng.standard_normal((2, 2, 4, 14, 56))).astype(np.float32)
    with torch.no_grad():
        out.update(ro_s=s2, ro_llr=quiet(ro, t(s2)).numpy(), ro_h=quiet(ch, t(s2)).numpy())

    # ---- SeparableConv2d twin: first StateInit layer (arrays 0..2), channels-first, H = F, W = T -----
    sc = twin["SeparableConv2d"](18, 128, 3, bias=True)
    with torch.no_grad():
        sc.depthwise.weight.copy_(t(np.transpose(w[0], (2, 3, 0, 1))))        # [3,3,C,1] -> [C,1,3,3]
        sc.depthwise.bias.zero_()                                             # Keras has no depthwise bias
        sc.pointwise.weight.copy_(t(np.transpose(w[1], (3, 2, 0, 1))))        # [1,1,Cin,Cout] -> [Cout,Cin,1,1]
        sc.pointwise.bias.copy_(t(w[2]))
    x = rng.standard_normal((3, 7, 14, 18)).astype(np.float32)                # [N, F, T, C]
    with torch.no_grad():
        y = sc(t(x).permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    out.update(sep_x=x, sep_y=y.numpy())

    # ---- NearestNeighborInterpolator: gather indices for the 2-PRB PUSCH pilot pattern ---------------
    grid = build_grid(cfg, n_size_bwp=2)
    U, T, Fs = grid.num_tx, grid.num_ofdm_symbols, grid.num_subcarriers
    pp = types.SimpleNamespace(num_pilot_symbols=grid.pilots.shape[1],
                               mask=np.broadcast_to(grid.pilot_mask[None, None], (U, 1, T, Fs)).copy(),
                               pilots=grid.pilots[:, None, :])
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:                                # the class writes data/*.npy
        os.chdir(tmp)
        try:
            nn_i = quiet(live["NearestNeighborInterpolator"], pp)
        finally:
            os.chdir(cwd)
    out.update(nn_gather_ind=np.asarray(nn_i._gather_ind).reshape(U, T, Fs).astype(np.int32), nn_prb=np.int64(2))

    # ---- NRPreprocessing: FOCC removal and the per-PRB nearest-pilot template -------------------------
    pre = live["NRPreprocessing"](2)
    ofdm_pos = torch.tensor([[2, 11], [2, 11]])
    sc_pos = torch.tensor([[0, 2, 4, 6, 8, 10], [1, 3, 5, 7, 9, 11]])
    h = rng.standard_normal((2, 8, 2, 24)).astype(np.float32)                 # [B, 2N, U, n_pilots]
    with torch.no_grad():
        hf = quiet(pre._focc_removal, t(h))
        nn_idx, pe = quiet(pre._calculate_nn_indices, ofdm_pos, sc_pos, 14, 2)
    out.update(focc_in=h, focc_out=hf.numpy(), aer_ofdm_pos=ofdm_pos.numpy(), aer_sc_pos=sc_pos.numpy(),
               aer_nn_idx=nn_idx.numpy().astype(np.int32), aer_pe=pe.numpy().astype(np.float32))

    # ---- positional encoding: NumPy fragment of utils/onnx_utils.py (distance to the nearest own pilot) ----
    import textwrap
    lines = open(os.path.join(REF, "utils", "onnx_utils.py")).read().splitlines()
    i0 = next(i for i, l in enumerate(lines) if "# Distance to the nearest pilot in time" in l)
    i1 = next(i for i, l in enumerate(lines) if "nearest_pilot_dist = np.stack(" in l) + 3
    frag = textwrap.dedent("\n".join(lines[i0:i1]))
    grid4 = build_grid(cfg, n_size_bwp=4)
    Tn, Fn = grid4.num_ofdm_symbols, grid4.num_subcarriers
    ind = []                                   # what tf.where(|pilots_only| > 1e-3) yields, per tx: (t, f) row-major
    for u in range(grid4.num_tx):
        pu = np.zeros((Tn, Fn), np.complex64)
        pu[grid4.pilot_mask] = grid4.pilots[u]
        ind.append(np.argwhere(np.abs(pu) > 1e-3))
    loc = {"np": np, "max_num_tx": grid4.num_tx, "rg": types.SimpleNamespace(num_ofdm_symbols=Tn, fft_size=Fn),
           "pilot_ind_sorted": np.array(ind)}
    exec(frag, loc)
    out.update(pe_ref=np.transpose(loc["nearest_pilot_dist"], (0, 2, 1, 3)).astype(np.float32), pe_prb=np.int64(4))

    path = os.path.join(HERE, "ref_exec_fixtures.npz")
    np.savez_compressed(path, **out)
    print(path, {k: v.shape for k, v in out.items() if hasattr(v, "shape")})


if __name__ == "__main__":
    main()
