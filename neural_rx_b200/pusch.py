"""PUSCH resource-grid geometry the receiver needs: DMRS pilots, nearest-pilot gather tables,
positional encoding and the data-RE ordering of the demapper.

In the reference all of this comes from Sionna objects built in ``utils/parameters.py:139-261``
(``PUSCHConfig`` / ``PUSCHPilotPattern`` / ``ResourceGrid``); the receiver then derives
  * the nearest-neighbour interpolation indices of ``PUSCHLSChannelEstimator(interpolation_type="nn")``
    (call site ``utils/neural_rx copy_pytorch.py:848-854``; NumPy twin ``utils/neural_rx.py:950-1000``),
  * the positional encoding (``utils/onnx_utils.py:172-260``),
  * the data-RE gather of ``ResourceGridDemapper`` (``utils/neural_rx.py:849``, ordering spec
    ``utils/onnx_utils.py:462-516``).
Here they are computed in closed form / vectorised for DMRS configuration type 1, mapping type A,
single-symbol DMRS (the only family the reference configs use, ``config/*.cfg:31-40``).

Index conventions: grids are ``[t, f]`` flattened as ``t * F + f``; the per-UE pilot vector has
``n_dmrs * F`` slots ordered (DMRS symbol ascending, subcarrier ascending) with zeros off-comb.
"""
from __future__ import annotations

import dataclasses
from typing import Sequence

import numpy as np

from .config import NrxConfig


def gold_sequence(c_init: int, length: int) -> np.ndarray:
    """Pseudo-random sequence c(n) of TS 38.211 §5.2.1 (length-31 Gold sequence, Nc = 1600)."""
    nc = 1600
    n_total = nc + length
    x1 = np.zeros(n_total + 31, dtype=np.uint8)
    x2 = np.zeros(n_total + 31, dtype=np.uint8)
    x1[0] = 1
    for i in range(31):
        x2[i] = (c_init >> i) & 1
    for n in range(n_total):
        x1[n + 31] = x1[n + 3] ^ x1[n]
        x2[n + 31] = x2[n + 3] ^ x2[n + 2] ^ x2[n + 1] ^ x2[n]
    return (x1[nc:nc + length] ^ x2[nc:nc + length]).astype(np.uint8)


def dmrs_base_sequence(num_subcarriers: int, symbol: int, slot_number: int = 0, n_id: int = 1,
                       n_scid: int = 1, symbols_per_slot: int = 14) -> np.ndarray:
    """r(m), m = 0..F/2-1, of TS 38.211 §6.4.1.1.1.1 for one OFDM symbol (QPSK, unit modulus)."""
    c_init = ((1 << 17) * (symbols_per_slot * slot_number + symbol + 1) * (2 * n_id + 1)
              + 2 * n_id + n_scid) % (1 << 31)
    m = num_subcarriers // 2
    c = gold_sequence(c_init, 2 * m).astype(np.float64)
    return ((1 - 2 * c[0::2]) + 1j * (1 - 2 * c[1::2])) / np.sqrt(2.0)


@dataclasses.dataclass
class PuschGrid:
    """All constant tables of one (config, #PRB) pair."""

    num_tx: int
    num_subcarriers: int
    num_ofdm_symbols: int
    dmrs_symbols: Sequence[int]
    pilots: np.ndarray        # [U, n_dmrs*F] complex64, zeros off-comb (Sionna pilot_pattern.pilots[:,0])
    pilot_mask: np.ndarray    # [T, F] bool: RE carries a pilot slot (all subcarriers of DMRS symbols)
    nn_index: np.ndarray      # [U, T*F] int32: pilot slot whose (FOCC-averaged) LS estimate fills the RE
    pos_enc: np.ndarray       # [U, F, T, 2] float32 (time, freq) normalised distance to nearest pilot
    data_index: np.ndarray    # [T*F] int32: ordinal of the RE among data REs, -1 for pilot REs
    num_data_res: int
    focc_block: int           # 2 * num_cdm_groups_without_data consecutive pilot slots share one estimate

    @property
    def num_pilot_slots(self) -> int:
        return len(self.dmrs_symbols) * self.num_subcarriers


def _comb_offset(port_set: Sequence[int]) -> int:
    """Subcarrier offset Delta of the CDM group of a type-1 DMRS port (TS 38.211 Table 6.4.1.1.3-1)."""
    port = port_set[0]
    return 0 if port in (0, 1, 4, 5) else 1


def build_grid(cfg: NrxConfig, n_size_bwp: int | None = None, slot_number: int | None = None,
               n_id: int | None = None, n_scid: int | None = None, pilots: np.ndarray | None = None) -> PuschGrid:
    """Build the constant tables.  ``pilots`` ([U, n_dmrs*F] complex) overrides the 38.211 sequence
    (any values are fine — only the zero / non-zero pattern shapes the tables).  The DMRS scrambling
    parameters default to the cfg's (``slot_number``, ``n_scid``, ``dmrs_nid[u][n_scid]`` per transmitter:
    utils/parameters.py:140-192); the LS estimate divides by these pilot values, so they must be the
    transmitter's."""
    slot_number = cfg.slot_number if slot_number is None else slot_number
    n_scid = cfg.n_scid if n_scid is None else n_scid
    if cfg.n_start_grid != 0:
        raise NotImplementedError("n_start_grid != 0: the pilot sequence offset of a shifted carrier grid is not implemented")
    for ports in cfg.dmrs_port_sets[:cfg.max_num_tx]:
        if len(ports) != 1:
            raise NotImplementedError("one DMRS port (one layer) per transmitter only: multi-port sets are not implemented")
    F = 12 * (cfg.n_size_bwp if n_size_bwp is None else n_size_bwp)
    T = cfg.num_ofdm_symbols
    U = cfg.max_num_tx
    syms = list(cfg.dmrs_symbols)
    nd = len(syms)
    if cfg.num_cdm_groups_without_data != 2:
        raise NotImplementedError("num_cdm_groups_without_data = 2 (no data on DMRS symbols) only")

    if pilots is None:
        pilots = np.zeros((U, nd * F), dtype=np.complex64)
        for u in range(U):
            delta = _comb_offset(cfg.dmrs_port_sets[u])
            nid_u = n_id if n_id is not None else int(cfg.dmrs_nid[u % len(cfg.dmrs_nid)][n_scid])
            # a_k = beta * w_f(k') * r(2n + k'), k = 4n + 2k' + Delta; TS 38.211 Table 6.4.1.1.3-1 (config type 1):
            # w_f = [+1, +1] for the even ports 0, 2, 4, 6 and [+1, -1] for the odd ports 1, 3, 5, 7
            w_f = np.where(np.arange(F // 2) % 2 == 0, 1.0, -1.0) if cfg.dmrs_port_sets[u][0] % 2 else np.ones(F // 2)
            for j, l in enumerate(syms):
                r = dmrs_base_sequence(F, l, slot_number, nid_u, n_scid, T)
                pilots[u, j * F + delta + 2 * np.arange(F // 2)] = (np.sqrt(2.0) * w_f * r).astype(np.complex64)
    pilots = np.asarray(pilots, dtype=np.complex64)
    assert pilots.shape == (U, nd * F)

    pilot_mask = np.zeros((T, F), dtype=bool)
    pilot_mask[syms, :] = True

    # nearest non-zero pilot in Manhattan distance, ties -> lowest pilot slot (utils/neural_rx.py:973-992)
    nn_index = np.zeros((U, T * F), dtype=np.int32)
    pos_enc = np.zeros((U, F, T, 2), dtype=np.float32)
    f_all = np.arange(F)
    t_all = np.arange(T)
    for u in range(U):
        best_d = np.full((T, F), np.iinfo(np.int32).max, dtype=np.int64)
        best_i = np.zeros((T, F), dtype=np.int64)
        df_min = np.full(F, np.inf)
        dt_min = np.full(T, np.inf)
        for j, l in enumerate(syms):
            nz = np.flatnonzero(np.abs(pilots[u, j * F:(j + 1) * F]) > 0)
            if nz.size == 0:
                continue
            # nearest non-zero subcarrier on this symbol, ties -> lower subcarrier
            pos = np.searchsorted(nz, f_all)
            lo = nz[np.clip(pos - 1, 0, nz.size - 1)]
            hi = nz[np.clip(pos, 0, nz.size - 1)]
            d_lo = np.abs(f_all - lo)
            d_hi = np.abs(f_all - hi)
            pick = np.where(d_hi < d_lo, hi, lo)
            df = np.minimum(d_lo, d_hi)
            d = np.abs(t_all - l)[:, None] + df[None, :]
            upd = d < best_d            # strict: earlier symbol (lower slot) wins ties
            best_d = np.where(upd, d, best_d)
            best_i = np.where(upd, j * F + pick[None, :], best_i)
            # positional encoding uses |pilot| > 1e-3 (utils/onnx_utils.py:191); same set here
            df_min = np.minimum(df_min, df)
            dt_min = np.minimum(dt_min, np.abs(t_all - l))
        nn_index[u] = best_i.reshape(-1)
        dt = dt_min - dt_min.mean()
        if dt.std() > 0:
            dt = dt / dt.std()
        dfreq = df_min - df_min.mean()
        if dfreq.std() > 0:
            dfreq = dfreq / dfreq.std()
        pos_enc[u, :, :, 0] = dt[None, :]
        pos_enc[u, :, :, 1] = dfreq[:, None]

    flat_mask = pilot_mask.reshape(-1)
    data_index = np.full(T * F, -1, dtype=np.int32)
    data_index[~flat_mask] = np.arange(int((~flat_mask).sum()), dtype=np.int32)
    return PuschGrid(U, F, T, tuple(syms), pilots, pilot_mask, nn_index, pos_enc, data_index,
                     int((~flat_mask).sum()), 2 * cfg.num_cdm_groups_without_data)
