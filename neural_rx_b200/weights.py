"""Weight files of the neural receiver.

``weights/<label>_weights`` is a ``pickle`` of the Keras ``model.get_weights()`` list — fp32 NumPy
arrays in layer-creation order (``utils/utils.py:34-70``).  This module walks that flat list with
the architecture read from the config and names every tensor:

    for each StateInit stack   : SepConv(18->h0) SepConv(h0->h1) ... SepConv(->d_s)
    for each iteration         : Dense(d_s->a0) ... Dense(->d_s)          (AggregateUserStates)
                                 SepConv(2*d_s+2->u0) ... SepConv(->d_s)   (UpdateState)
    for each LLR readout head  : Dense(d_s->r0) ... Dense(->bits)
    channel-estimate readout   : Dense(d_s->r0) ... Dense(->2*N_rx)

SeparableConv2D contributes ``depthwise_kernel [3,3,Cin,1]``, ``pointwise_kernel [1,1,Cin,Cout]``,
``bias [Cout]``; Dense contributes ``kernel [in,out]``, ``bias [out]``.
"""
from __future__ import annotations

import dataclasses
import pickle
from typing import List, Sequence

import numpy as np

from .config import NrxConfig


@dataclasses.dataclass
class SepConv:
    dw: np.ndarray   # [3, 3, Cin]   (Delta f, Delta t, channel) cross-correlation taps
    pw: np.ndarray   # [Cin, Cout]
    b: np.ndarray    # [Cout]

    @property
    def cin(self) -> int:
        return self.pw.shape[0]

    @property
    def cout(self) -> int:
        return self.pw.shape[1]


@dataclasses.dataclass
class Dense:
    k: np.ndarray    # [in, out]
    b: np.ndarray    # [out]


@dataclasses.dataclass
class IterationWeights:
    agg: List[Dense]        # hidden (ReLU) layers then linear output layer
    update: List[SepConv]   # hidden (ReLU) layers then linear output layer


@dataclasses.dataclass
class NrxWeights:
    state_init: List[List[SepConv]]     # [num_io_stacks][layers]
    iterations: List[IterationWeights]  # [num_nrx_iter]
    readout_llr: List[List[Dense]]      # [num_io_stacks][layers]
    readout_chest: List[Dense]

    def num_params(self) -> int:
        return sum(int(a.size) for a in self.to_list())

    def mac_per_pixel(self, head: int = 0, num_it: int | None = None) -> int:
        """Multiply-accumulates per user resource element (biases excluded)."""
        num_it = len(self.iterations) if num_it is None else num_it
        macs = sum(int(l.dw.size + l.pw.size) for l in self.state_init[head])
        for it in self.iterations[:num_it]:
            macs += sum(int(l.k.size) for l in it.agg)
            macs += sum(int(l.dw.size + l.pw.size) for l in it.update)
        macs += sum(int(l.k.size) for l in self.readout_llr[head])
        macs += sum(int(l.k.size) for l in self.readout_chest)
        return macs

    def to_list(self) -> List[np.ndarray]:
        """Back to the Keras ``get_weights()`` order (inverse of :func:`from_list`)."""
        out: List[np.ndarray] = []

        def sep(l: SepConv):
            out.extend([l.dw[..., None], l.pw[None, None], l.b])

        def den(l: Dense):
            out.extend([l.k, l.b])

        for stack in self.state_init:
            for l in stack:
                sep(l)
        for it in self.iterations:
            for l in it.agg:
                den(l)
            for l in it.update:
                sep(l)
        for head in self.readout_llr:
            for l in head:
                den(l)
        for l in self.readout_chest:
            den(l)
        return out


def _expected_layout(cfg: NrxConfig):
    """Yield ('sep'|'dense', cin, cout) in file order."""
    c_in0 = 2 * cfg.num_rx_antennas + 2 + (2 * cfg.num_rx_antennas if cfg.initial_chest else 0)
    for _ in range(cfg.num_io_stacks):
        cin = c_in0
        for n in list(cfg.num_units_init) + [cfg.d_s]:
            yield ("sep", cin, n)
            cin = n
    for i in range(cfg.num_nrx_iter):
        cin = cfg.d_s
        for n in list(cfg.num_units_agg[i]) + [cfg.d_s]:
            yield ("dense", cin, n)
            cin = n
        cin = 2 * cfg.d_s + 2
        for n in list(cfg.num_units_state[i]) + [cfg.d_s]:
            yield ("sep", cin, n)
            cin = n
    for bits in cfg.readout_bits:
        cin = cfg.d_s
        for n in list(cfg.num_units_readout) + [bits]:
            yield ("dense", cin, n)
            cin = n
    cin = cfg.d_s
    for n in list(cfg.num_units_readout) + [2 * cfg.num_rx_antennas]:
        yield ("dense", cin, n)
        cin = n


def from_list(cfg: NrxConfig, arrays: Sequence[np.ndarray]) -> NrxWeights:
    """Name the flat ``get_weights()`` list; raises ``ValueError`` on any shape mismatch."""
    arrays = [np.asarray(a, dtype=np.float32) for a in arrays]
    pos = 0
    layers = []
    for kind, cin, cout in _expected_layout(cfg):
        if kind == "sep":
            if pos + 3 > len(arrays):
                raise ValueError("weight list too short for the configured architecture")
            dw, pw, b = arrays[pos:pos + 3]
            pos += 3
            if dw.shape != (3, 3, cin, 1) or pw.shape != (1, 1, cin, cout) or b.shape != (cout,):
                raise ValueError(f"SeparableConv2D({cin}->{cout}) expected at index {pos - 3}, "
                                 f"got {dw.shape} {pw.shape} {b.shape}")
            layers.append(SepConv(dw[..., 0].copy(), pw[0, 0].copy(), b.copy()))
        else:
            if pos + 2 > len(arrays):
                raise ValueError("weight list too short for the configured architecture")
            k, b = arrays[pos:pos + 2]
            pos += 2
            if k.shape != (cin, cout) or b.shape != (cout,):
                raise ValueError(f"Dense({cin}->{cout}) expected at index {pos - 2}, got {k.shape} {b.shape}")
            layers.append(Dense(k.copy(), b.copy()))
    if pos != len(arrays):
        raise ValueError(f"weight list has {len(arrays)} arrays, architecture consumes {pos}")

    it = iter(layers)
    n_init = len(cfg.num_units_init) + 1
    state_init = [[next(it) for _ in range(n_init)] for _ in range(cfg.num_io_stacks)]
    iterations = []
    for i in range(cfg.num_nrx_iter):
        agg = [next(it) for _ in range(len(cfg.num_units_agg[i]) + 1)]
        upd = [next(it) for _ in range(len(cfg.num_units_state[i]) + 1)]
        iterations.append(IterationWeights(agg, upd))
    n_ro = len(cfg.num_units_readout) + 1
    readout_llr = [[next(it) for _ in range(n_ro)] for _ in range(len(cfg.readout_bits))]
    readout_chest = [next(it) for _ in range(n_ro)]
    return NrxWeights(state_init, iterations, readout_llr, readout_chest)


class _ArraysOnlyUnpickler(pickle.Unpickler):
    """The reference ``pickle.load``s the weight file (utils/utils.py:53-70), i.e. runs whatever the file asks for.
    A weight file is a list of NumPy arrays: this unpickler resolves only the globals NumPy's array pickles need
    and refuses everything else, so a tampered file cannot execute code."""

    _ALLOWED = {"_reconstruct", "ndarray", "dtype", "scalar", "_frombuffer"}

    def find_class(self, module, name):
        if module.split(".")[0] == "numpy" and name in self._ALLOWED:
            return super().find_class(module, name)
        raise pickle.UnpicklingError(f"weight file references {module}.{name}: only NumPy arrays are accepted")


def load_weights(cfg: NrxConfig, model_path: str) -> NrxWeights:
    """``utils.load_weights`` equivalent: unpickle the list and bind it to the architecture."""
    with open(model_path, "rb") as f:
        arrays = _ArraysOnlyUnpickler(f).load()
    if not isinstance(arrays, (list, tuple)):
        raise ValueError("weight file does not hold a list of arrays")
    return from_list(cfg, arrays)


def save_weights(weights: NrxWeights, model_path: str) -> None:
    """``utils.save_weights`` equivalent (same on-disk format)."""
    with open(model_path, "wb") as f:
        pickle.dump(weights.to_list(), f)


def random_weights(cfg: NrxConfig, seed: int = 0) -> NrxWeights:
    """Seeded random-init weights of the configured architecture (He-style scaling, non-zero
    biases) — used for synthetic benchmarks and for tests when no weight file is staged."""
    rng = np.random.default_rng(seed)
    arrays = []
    for kind, cin, cout in _expected_layout(cfg):
        if kind == "sep":
            arrays.append((rng.standard_normal((3, 3, cin, 1)) * 0.35).astype(np.float32))
            arrays.append((rng.standard_normal((1, 1, cin, cout)) * np.sqrt(1.6 / cin)).astype(np.float32))
        else:
            arrays.append((rng.standard_normal((cin, cout)) * np.sqrt(1.6 / cin)).astype(np.float32))
        arrays.append((rng.standard_normal((cout,)) * 0.1).astype(np.float32))
    return from_list(cfg, arrays)
