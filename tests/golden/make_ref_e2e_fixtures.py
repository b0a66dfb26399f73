"""END-TO-END pins: the reference's own forward code, executed here, produces the golden LLRs.

    python tests/golden/make_ref_e2e_fixtures.py     (needs /root/reference; run in the build container)

``make_ref_fixtures.py`` pins the oracle block by block.  This script pins the COMPOSITION: it runs the
reference's own ``forward`` methods — extracted from the reference sources with ``ast``, nothing is
copied into this repo — on seeded synthetic slots with the shipped weights and stores inputs + outputs
(``ref_e2e_fixtures.npz``):

* ``StateInit.forward``            utils/neural_rx.py:106-132   (tile y / pe, concat ``[y, pe, h_hat]``, stack)
* ``AggregateUserStates.forward``  utils/neural_rx.py:176-207
* ``UpdateState.forward``          utils/neural_rx.py:249-270   (concat ``[a, s, pe]``, stack, residual)
* ``CGNNIt.forward``               utils/neural_rx copy_pytorch.py:311-321 (aggregate, then update; the live
                                   class of utils/neural_rx.py:273-306 never calls UpdateState — SURVEY App. B)
* ``ReadoutLLRs`` / ``ReadoutChEst``  utils/neural_rx.py:309-404
* ``CGNN.forward``                 utils/neural_rx.py:544-595   (normalisation, Var-IO blend of the StateInit
                                   stacks, iteration loop, read-outs after the last iteration, masking slice)
* ``NeuralReceiverONNX.forward``   utils/neural_rx.py:1773-1812 with ``NRPreprocessing.forward`` :1698-1711
                                   (re|im concat, FOCC removal, interpolation, slicing, LLR permute + sign)
* ``post_process_llrs``            utils/onnx_utils.py:472-516  (Aerial LLR tensor -> coded-bit order of the TB
                                   decoder: sign, transposes, data-RE gather, (RE, bit) flattening) — TF code,
                                   executed against a NumPy stand-in of the four tensor utilities it calls

The constructors of these classes are defective in the fork (dense ``nn.Conv2d`` for
"sepconv", ``UpdateState`` builds layers without ``in_channels``, ``in_channels = 16``, ``CGNNIt`` builds no
``UpdateState``), so the objects are assembled with ``object.__new__`` + ``nn.Module.__init__`` and their
attributes are filled with the reference's own leaf classes; the FORWARD code is what runs unmodified.

Fork defects that had to be routed around (each one documented where it is handled):

1. sep-conv layers: the reference's ``SeparableConv2d`` twin (utils/neural_rx copy_pytorch.py:34-51, torch
   channels-first) is injected as ``_hidden_conv`` / ``_output_conv`` behind a channels-last <-> channels-first
   adapter, because ``StateInit.forward`` / ``UpdateState.forward`` feed ``[N, F, T, C]`` tensors as the TF
   original did.  Keras has no depthwise bias: the twin's depthwise bias is zero.
2. ``active_tx`` is handed over as ``[B, U, 1, 1]``: ``AggregateUserStates.forward`` does
   ``active_tx.unsqueeze(-1).expand_as(sp)`` (:192), which only broadcasts from that rank (the TF original
   used ``expand_to_rank``).  Values are unchanged.
3. ``mcs_ue_mask`` is handed over as ``[B, U, n_mcs, 1]``: ``CGNN.forward`` does
   ``mcs_ue_mask[:, :, idx:idx+1].unsqueeze(-1)`` (:563-568), which lines up with the 5-D state only from that
   rank (TF: ``expand_to_rank(gather(mcs_ue_mask, idx, axis=2), 5, axis=-1)``).  Values are unchanged.
4. ``NRPreprocessing._nn_interpolation`` (:1672-1696) cannot execute: its first ``view`` asks for a shape with
   more elements than the tensor has and ``gather(...expand_as...)`` mismatches ranks (a mechanical
   translation of TF's ``split_dim`` / ``tf.gather(batch_dims=...)``).  It is replaced by a six-line gather that
   uses the indices returned by the reference's own ``_calculate_nn_indices`` (executed) on the output of the
   reference's own ``_focc_removal`` (executed), and returns the layout the remaining reference lines
   (:1709-1710 slicing + permute) expect.  ``_calculate_nn_indices`` normalises the positional encoding with
   torch's UNBIASED ``.std()`` where the TF original used ``tf.math.reduce_std`` (population): during that call
   ``torch.Tensor.std`` is switched to the population estimator, and — second defect in the same method — its
   ``view(1, T, 12, 2)`` of values enumerated subcarrier-major is undone (see ``_nn_interp``).
5. debug ``print`` calls inside ``AggregateUserStates.forward`` are silenced.
6. ``StateInit.forward`` calls ``h_hat.view(...)`` (:120) on what ``NRPreprocessing.forward`` returns through a
   ``permute`` (:1710); torch refuses ``view`` on non-contiguous memory (TF's reshape does not care).  The
   stand-in of 4. therefore lays its result out in memory so that the permuted tensor is contiguous.  Values
   are unchanged.

Inputs are seeded synthetic slots (``neural_rx_b200.synth.make_slots``, 4 PRB); the CGNN inputs derived from
the complex grid (re|im re-layout, LS + FOCC + nearest-pilot estimate, positional encoding) are stored too, so
the fixtures do not depend on this repo's pre-processing staying unchanged.
"""
import contextlib
import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

from make_ref_fixtures import REF, quiet, reference_classes  # noqa: E402

from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.synth import aerial_inputs, make_slots  # noqa: E402
from neural_rx_b200.weights import load_weights  # noqa: E402
from oracle import nrx_oracle as O  # noqa: E402  (only for the pre-processing of the Sionna-shaped inputs)

N_PRB = 4


def t32(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32)


class ChannelsLast(nn.Module):
    """[N, F, T, C] in / out around a channels-first torch module (defect 1)."""

    def __init__(self, m):
        super().__init__()
        self.m = m

    def forward(self, x):
        return self.m(x.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)


def bare(cls):
    obj = object.__new__(cls)
    nn.Module.__init__(obj)
    return obj


class RefBuilder:
    """Assembles the reference's CGNN from the reference's own classes and a Keras weight list
    (creation order: SURVEY.md App. A.2)."""

    def __init__(self, live, twin, cfg, arrays):
        self.live, self.twin, self.cfg = live, twin, cfg
        self.a, self.i = list(arrays), 0

    def sep(self, cin, cout):
        dw, pw, b = self.a[self.i:self.i + 3]
        self.i += 3
        assert dw.shape == (3, 3, cin, 1) and pw.shape == (1, 1, cin, cout)
        m = self.twin["SeparableConv2d"](cin, cout, 3, bias=True)
        with torch.no_grad():
            m.depthwise.weight.copy_(t32(np.transpose(dw, (2, 3, 0, 1))))      # [3,3,C,1] -> [C,1,3,3]
            m.depthwise.bias.zero_()
            m.pointwise.weight.copy_(t32(np.transpose(pw, (3, 2, 0, 1))))      # [1,1,Cin,Cout] -> [Cout,Cin,1,1]
            m.pointwise.bias.copy_(t32(b))
        return ChannelsLast(m)

    def fill_dense(self, mod):
        layers = list(mod._hidden_layers) + [mod._output_layer]
        for l in layers:
            k, b = self.a[self.i:self.i + 2]
            self.i += 2
            assert k.shape == (l.in_features, l.out_features), (k.shape, l)
            with torch.no_grad():
                l.weight.copy_(t32(k.T))
                l.bias.copy_(t32(b))
        return mod

    def conv_stack(self, cls, cin, units, d_s):
        obj = bare(cls)
        hidden = []
        for n in units:
            hidden.append(self.sep(cin, n))
            cin = n
        obj._hidden_conv = nn.ModuleList(hidden)
        obj._output_conv = self.sep(cin, d_s)
        return obj

    def cgnn(self):
        cfg, live = self.cfg, self.live
        d_s, nrx = cfg.d_s, cfg.num_rx_antennas
        bits = list(cfg.num_bits_per_symbol)
        masking = bool(cfg.mcs_var_mcs_masking)
        g = bare(live["CGNN"])
        g._training, g._apply_multiloss, g._var_mcs_masking = False, False, masking
        n_io = 1 if masking else len(bits)
        inits = [self.conv_stack(live["StateInit"], 4 * nrx + 2, cfg.num_units_init, d_s) for _ in range(n_io)]
        g._s_init = inits if masking else nn.ModuleList(inits)      # a plain list in masking mode, as in :445-454
        its = []
        for i in range(cfg.num_nrx_iter):
            it = bare(self.twin["CGNNIt"])
            it._state_aggreg = self.fill_dense(live["AggregateUserStates"](d_s, list(cfg.num_units_agg[i]), d_s))
            it._state_update = self.conv_stack(live["UpdateState"], 2 * d_s + 2, cfg.num_units_state[i], d_s)
            its.append(it)
        g._iterations = nn.ModuleList(its)
        g._num_it = cfg.num_nrx_iter
        heads = [max(bits)] if masking else bits
        ro = [self.fill_dense(live["ReadoutLLRs"](b, list(cfg.num_units_readout), d_s)) for b in heads]
        g._readout_llrs = ro if masking else nn.ModuleList(ro)
        g._readout_chest = self.fill_dense(live["ReadoutChEst"](nrx, list(cfg.num_units_readout), d_s))
        g._num_mcss_supported = len(bits)
        g._num_bits_per_symbol = bits
        assert self.i == len(self.a), (self.i, len(self.a))
        return g


@contextlib.contextmanager
def population_std():
    """defect 4: tf.math.reduce_std is the population estimator."""
    orig = torch.Tensor.std
    torch.Tensor.std = lambda self, *a, **k: orig(self, *a, unbiased=False, **k)
    try:
        yield
    finally:
        torch.Tensor.std = orig


def _nn_interp(pre, h_hat, num_ofdm_symbols, dmrs_ofdm_pos, dmrs_subcarrier_pos):
    """Stand-in for the non-executable ``NRPreprocessing._nn_interpolation`` (defect 4).

    h_hat [B, 2N, U, n_pilots] (after the reference's ``_focc_removal``), pilots DMRS-symbol-major, subcarrier
    ascending.  The indices come from the reference's ``_calculate_nn_indices``: per RE of the 12 x T PRB
    template the winner among the candidates enumerated by ``meshgrid(dmrs_subcarrier_pos, dmrs_ofdm_pos)``,
    i.e. ``idx = k * n_sym + j`` (k-th pilot subcarrier of the PRB, j-th DMRS symbol).  That method enumerates
    REs subcarrier-major (``meshgrid(arange(12), arange(T))``, torch default 'ij') but then views the result as
    ``[.., T, 12]`` the way the TF original (``tf.meshgrid``, 'xy') laid it out; reading it back as [12, T]
    restores the per-RE values.  Output: [B, 1, 2N, U, 1, T, F] — the Sionna LS-estimator layout that the
    reference's remaining lines slice and permute."""
    B, C, U, n_p = h_hat.shape
    n_sc, n_sym = dmrs_subcarrier_pos.shape[1], dmrs_ofdm_pos.shape[1]
    n_prb = n_p // (n_sc * n_sym)
    with population_std():
        nn_idx, pe = pre._calculate_nn_indices(dmrs_ofdm_pos, dmrs_subcarrier_pos, num_ofdm_symbols, n_prb)
    idx = nn_idx.reshape(U, 12, num_ofdm_symbols)                          # per RE (sc, t) of the template
    # same un-view for the encoding: the fork's pe[u, b, a] holds the value of flat RE index a*12 + b
    pe = pe[:, :12].permute(0, 2, 1, 3).reshape(U, 12, num_ofdm_symbols, 2).repeat(1, n_prb, 1, 1)
    hp = h_hat.reshape(B, C, U, n_sym, n_prb, n_sc)
    # memory order [B, U, F, T, C] (defect 6): the reference's next lines slice + permute into that order and
    # StateInit.forward then calls .view on the result, which torch only allows on contiguous memory
    base = torch.zeros(B, U, 12 * n_prb, num_ofdm_symbols, C)
    for u in range(U):
        for sc in range(12):
            for t in range(num_ofdm_symbols):
                k, j = int(idx[u, sc, t]) // n_sym, int(idx[u, sc, t]) % n_sym
                base[:, u, sc::12, t, :] = hp[:, :, u, j, :, k].permute(0, 2, 1)
    return base.permute(0, 4, 1, 3, 2).unsqueeze(1).unsqueeze(4), pe


def run_cgnn(g, y8, pe, h_hat, active, mask):
    """reference CGNN.forward on [B,F,T,2N] / [U,F,T,2] / [B,U,F,T,2N]; defects 2, 3, 5."""
    B, U = active.shape
    n_mcs = mask.shape[-1]
    with torch.no_grad():
        llrs, h_hats = quiet(g, [t32(y8), t32(pe), t32(h_hat), t32(active).reshape(B, U, 1, 1),
                                 t32(mask).reshape(B, U, n_mcs, 1)])
    assert len(llrs) == 1 and len(h_hats) == 1                  # read-outs only after the last iteration (:582)
    return [l.numpy() for l in llrs[0]], h_hats[0].numpy()


class _TfShim:
    """The four TensorFlow / Sionna tensor utilities ``post_process_llrs`` uses, on NumPy arrays (generic library
    semantics: ``tf.transpose``, ``tf.gather`` with ``axis`` / ``batch_dims``, Sionna's ``flatten_dims`` /
    ``flatten_last_dims``).  TensorFlow is not installable here; the ORDER logic that runs is the reference's."""

    @staticmethod
    def transpose(x, perm):
        return np.transpose(x, perm)

    @staticmethod
    def gather(x, ind, axis=0, batch_dims=0):
        ind = np.asarray(ind)
        axis = axis % x.ndim
        if batch_dims == 0:
            return np.take(x, ind, axis=axis)
        assert batch_dims == 1 and axis == 1 and ind.ndim == 2 and ind.shape[0] == x.shape[0]
        return np.stack([np.take(x[i], ind[i], axis=0) for i in range(x.shape[0])], axis=0)


def flatten_dims(x, num_dims, axis):
    sh = list(x.shape)
    return x.reshape(sh[:axis] + [-1] + sh[axis + num_dims:])


def flatten_last_dims(x, num_dims=2):
    sh = list(x.shape)
    return x.reshape(sh[:-num_dims] + [-1])


def run_post_process_llrs(llr_aerial, grid):
    """Execute ``post_process_llrs`` (utils/onnx_utils.py:472-516, a method of the evaluation wrapper whose class
    needs TensorFlow to be defined) on an Aerial-layout LLR tensor [B,bits,U,F,T]: the function body is taken from
    the reference source with ``ast`` and run against the NumPy stand-ins above.  ``self`` carries what the
    constructor (:458-470) pre-computes: ``data_ind = argsort(flatten_last_dims(mask))[..., :num_data_symbols]``
    (TF's ascending argsort of a 0/1 mask; stable, so data REs keep their (t, f) order), ``eff_sub_ind`` = all
    subcarriers (PUSCH grid without guards / DC), ``num_tx``."""
    import ast
    import types
    path = os.path.join(REF, "utils", "onnx_utils.py")
    tree = ast.parse(open(path).read())
    fn = next(n for c in tree.body if isinstance(c, ast.ClassDef) for n in c.body
              if isinstance(n, ast.FunctionDef) and n.name == "post_process_llrs")
    ns = {"tf": _TfShim, "flatten_dims": flatten_dims, "flatten_last_dims": flatten_last_dims}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    U, T, Fs = grid.num_tx, grid.num_ofdm_symbols, grid.num_subcarriers
    mask = np.broadcast_to(grid.pilot_mask[None, None], (U, 1, T, Fs)).astype(np.int32)       # [num_tx, streams, T, F]
    data_ind = np.argsort(flatten_last_dims(mask), axis=-1, kind="stable")[..., :grid.num_data_res]
    me = types.SimpleNamespace(data_ind=data_ind, eff_sub_ind=np.arange(Fs), num_tx=U)
    return np.ascontiguousarray(ns["post_process_llrs"](me, llr_aerial), dtype=np.float32)


def main():
    live = reference_classes(os.path.join(REF, "utils", "neural_rx.py"),
                             ["StateInit", "AggregateUserStates", "UpdateState", "ReadoutLLRs", "ReadoutChEst", "CGNN",
                              "NRPreprocessing", "NeuralReceiverONNX"])
    twin = reference_classes(os.path.join(REF, "utils", "neural_rx copy_pytorch.py"), ["SeparableConv2d", "CGNNIt"],
                             uncomment=True)
    out = {"n_prb": np.int64(N_PRB)}

    # ---- Sionna-shaped path: complex grid -> (restated, block-pinned) pre-processing -> reference CGNN.forward ----
    cases = [
        # key, config, weights, active [B,U], mcs index per (slot, user) or None, num_it
        ("rt", "nrx_rt", "nrx_rt", [[1, 1], [1, 0], [0, 1]], None, None),
        ("rt_it1", "nrx_rt", "nrx_rt", [[1, 1]], None, 1),
        ("varmcs", "nrx_rt_var_mcs", "nrx_rt_var_mcs", [[1, 1], [1, 1], [1, 0], [1, 1]], [[0, 1], [1, 0], [1, 1], [0, 0]], None),
        ("masking", "nrx_large_var_mcs_64qam_masking", "nrx_large_var_mcs_64qam_masking", [[1, 1], [1, 0]],
         [[2, 0], [1, 1]], None),
        # the three 8-iteration BASELINE configs (configs[1], [3], [4])
        ("large", "nrx_large", "nrx_large", [[1, 1], [0, 1]], None, None),
        # (three slots: 64-QAM puts many LLRs next to zero — hard-decision agreement of the fp16 engine is 99.93 % on
        #  228 k-bit samples, bar 99.9 % — so one 4-PRB slot of 8 k bits is too small a sample for that statistic)
        ("large64", "nrx_large_64qam", "nrx_large_64qam", [[1, 1], [1, 1], [1, 1]], None, None),
        ("site", "nrx_site_specific_large", "nrx_site_specific_large", [[1, 1]], None, None),
    ]
    for key, label, wlabel, active, mcs, num_it in cases:
        cfg = get_config(label)
        grid = build_grid(cfg, n_size_bwp=N_PRB)
        arrays = load_weights(cfg, os.path.join(REF, "weights", f"{wlabel}_weights")).to_list()
        g = RefBuilder(live, twin, cfg, arrays).cgnn()
        if num_it is not None:
            g.num_it = num_it                                   # the reference's own setter (:537-542)
        active = np.asarray(active, np.float32)
        B, U = active.shape
        n_mcs = len(cfg.mcs_index)
        mcs = np.zeros((B, U), np.int64) if mcs is None else np.asarray(mcs, np.int64)
        ys = []
        for b in range(B):                                      # every slot with its own per-user MCS
            kw = dict(per_ue_power_norm=True, sparse_paths=24) if "site_specific" in label else {}
            sb = make_slots(cfg, grid, batch=1, ebno_db=8.0 + b, seed=9000 + 17 * len(out) + b,
                            mcs_per_ue=list(mcs[b]), active=active[b:b + 1], **kw)
            ys.append(sb.y)
        y = np.concatenate(ys, axis=0)
        mask = np.eye(n_mcs, dtype=np.float32)[mcs]             # [B,U,n_mcs] one-hot
        y8 = O.preprocess_y(y)
        pe = O.positional_encoding(grid.pilots, grid.pilot_mask)
        h_hat = O.ls_channel_estimate(y, grid.pilots, grid.pilot_mask)
        llrs, h_ref = run_cgnn(g, y8, pe, h_hat, active, mask)
        out.update({f"{key}_y": y.astype(np.complex64), f"{key}_y8": y8, f"{key}_pe": pe, f"{key}_h_hat": h_hat,
                    f"{key}_active": active, f"{key}_mask": mask, f"{key}_h_ref": h_ref,
                    f"{key}_num_it": np.int64(g.num_it)})
        for m, l in enumerate(llrs):
            out[f"{key}_llr{m}"] = l
        print(key, label, "heads", [l.shape for l in llrs], "h", h_ref.shape)

    # ---- Aerial / TensorRT-shaped path: reference NeuralReceiverONNX.forward ---------------------------------------
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=N_PRB)
    arrays = load_weights(cfg, os.path.join(REF, "weights", "nrx_rt_weights")).to_list()
    rx = bare(live["NeuralReceiverONNX"])
    rx._num_tx = grid.num_tx
    rx._cgnn = RefBuilder(live, twin, cfg, arrays).cgnn()
    rx._preprocessing = live["NRPreprocessing"](grid.num_tx)
    rx._preprocessing._nn_interpolation = lambda *a: _nn_interp(rx._preprocessing, *a)
    active = np.asarray([[1, 1], [1, 0], [0, 1]], np.float32)
    sb = make_slots(cfg, grid, batch=3, ebno_db=[7.0, 9.0, 11.0], seed=9901, active=active)
    ins = aerial_inputs(sb, grid)
    B, U = active.shape
    with torch.no_grad():
        llr, h = quiet(rx, [t32(ins[0]), t32(ins[1]), t32(ins[2]), t32(ins[3]), t32(ins[4]).reshape(B, U, 1, 1),
                            torch.as_tensor(ins[5]), torch.as_tensor(ins[6])])
    for i, a in enumerate(ins):
        out[f"aerial_in{i}"] = a
    out.update(aerial_llr=llr.numpy(), aerial_h=h.numpy(), aerial_y=sb.y.astype(np.complex64))
    print("aerial", llr.shape, h.shape)

    # ---- LLR -> coded-bit order: the reference's own post_process_llrs (utils/onnx_utils.py:472-516) -------------
    out["aerial_llr_demapped"] = run_post_process_llrs(llr.numpy(), grid)
    print("demapped", out["aerial_llr_demapped"].shape)

    path = os.path.join(HERE, "ref_e2e_fixtures.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
