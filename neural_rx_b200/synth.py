"""Seeded synthetic PUSCH slots (stand-in for Sionna's ``PUSCHTransmitter`` + channel).

The reference evaluates on ``E2E_Model`` slots: random bits -> PUSCH transmitter -> DoubleTDL
channel -> AWGN (``utils/e2e_model.py:392-653``, ``utils/channel_models.py:39-161``); Sionna is not
available offline, so this module restates what reaches the receiver:

  * DMRS symbols carry the 38.211 type-1 comb pilots of :mod:`neural_rx_b200.pusch`;
  * data REs carry Gray-mapped QAM (``utils/siona_tf.py:748-905``) of iid uniform bits;
  * per (UE, rx antenna) an L-tap Rayleigh channel with exponential power-delay profile
    (rms delay spread 100 ns for UE0, 300 ns for UE1 at 30 kHz SCS) and per-tap Doppler
    (max 400 Hz / 100 Hz) — the delay/Doppler spreads of ``DoubleTDLChannel``
    (``utils/channel_models.py:112-136``); unit average gain (TPMI-2 precoder folded in);
  * AWGN with N0 from Eb/N0 via ``ebnodb2no`` (``utils/siona_tf.py:3125-3200``).

It is "TDL-like", good for LLR parity, uncoded-BER sanity and throughput — not for reproducing
the reference's BLER curves point by point.
"""
from __future__ import annotations

import dataclasses
from typing import Optional, Sequence

import numpy as np

from .config import NrxConfig, mcs_bits_per_symbol, mcs_code_rate
from .pusch import PuschGrid, build_grid


def pam_gray(b: np.ndarray) -> int:
    """Gray-labelled PAM point of TS 38.211 §5.1 (recursive form)."""
    if len(b) > 1:
        return (1 - 2 * int(b[0])) * (2 ** len(b[1:]) - pam_gray(b[1:]))
    return 1 - 2 * int(b[0])


def qam_constellation(num_bits_per_symbol: int) -> np.ndarray:
    """Unit-power QAM constellation; point n is labelled by the binary representation of n,
    even-position bits -> real PAM, odd-position bits -> imaginary PAM."""
    pts = np.zeros(2 ** num_bits_per_symbol, dtype=np.complex128)
    for i in range(pts.size):
        b = np.array(list(np.binary_repr(i, num_bits_per_symbol)), dtype=np.int16)
        pts[i] = pam_gray(b[0::2]) + 1j * pam_gray(b[1::2])
    n = num_bits_per_symbol // 2
    var = np.sum(np.linspace(1, 2 ** n - 1, 2 ** (n - 1)) ** 2) / (2 ** (n - 2))
    return (pts / np.sqrt(var)).astype(np.complex64)


def ebnodb2no(ebno_db: float, num_bits_per_symbol: int, coderate: float, grid: PuschGrid,
              cp_overhead: float = 288.0 / 4096.0) -> float:
    """N0 for a rate-adjusted Eb/N0 on a PUSCH grid (restates ``ebnodb2no`` with a resource grid:
    one stream per UE, pilot/CP overhead accounted for)."""
    ebno = 10.0 ** (ebno_db / 10.0)
    num_syms = grid.num_ofdm_symbols * (1.0 + cp_overhead) * grid.num_subcarriers
    energy_per_symbol = num_syms / grid.num_data_res
    return float(1.0 / (ebno * coderate * num_bits_per_symbol / energy_per_symbol))


@dataclasses.dataclass
class SlotBatch:
    y: np.ndarray            # [B, 1, N_rx, T, F] complex64 — the receiver input of the reference
    active_tx: np.ndarray    # [B, U] float32 0/1
    bits: np.ndarray         # [B, U, n_data_res * max_bits] uint8, (RE, bit) bit fastest; padded with 0
    bits_per_ue: np.ndarray  # [B, U] int32 bits per symbol actually used by each UE
    h: np.ndarray            # [B, U, F, T, N_rx] complex64 true effective channel
    no: np.ndarray           # [B] float32


def make_slots(cfg: NrxConfig, grid: Optional[PuschGrid] = None, batch: int = 1,
               ebno_db: float | Sequence[float] = 6.0, seed: int = 0,
               mcs_per_ue: Optional[Sequence[int]] = None,
               active: Optional[np.ndarray] = None,
               delay_spread_ns: Sequence[float] = (100.0, 300.0),
               doppler_hz: Sequence[float] = (400.0, 100.0),
               per_ue_power_norm: bool = False,
               sparse_paths: Optional[int] = None,
               coded_bits: Optional[np.ndarray] = None) -> SlotBatch:
    """Generate ``batch`` slots.  ``mcs_per_ue[u]`` indexes ``cfg.mcs_index`` (default: head 0).

    ``sparse_paths`` switches to a ray-traced-shape channel (that many discrete paths with
    log-normal powers and uniform delays) for the site-specific configuration;
    ``per_ue_power_norm`` mirrors ``channel_norm_eval = True``.  ``coded_bits`` [batch, U, n_data_res x bits] (the
    output of a transport-block encoder, neural_rx_b200/tb.py) replaces the random payload; channel and noise
    realisations of a seed do not depend on it.
    """
    grid = build_grid(cfg) if grid is None else grid
    U, F, T, N = grid.num_tx, grid.num_subcarriers, grid.num_ofdm_symbols, cfg.num_rx_antennas
    mcs_per_ue = [0] * U if mcs_per_ue is None else list(mcs_per_ue)
    bits_ue = [mcs_bits_per_symbol(cfg.mcs_index[m], cfg.mcs_table) for m in mcs_per_ue]
    max_bits = max(bits_ue)
    ebno = np.broadcast_to(np.asarray(ebno_db, dtype=np.float64), (batch,))
    act = np.ones((batch, U), np.float32) if active is None else np.asarray(active, np.float32).reshape(batch, U)

    scs = 30e3
    t_sym = (1.0 + 288.0 / 4096.0) / scs
    data_mask = ~grid.pilot_mask                      # [T, F]
    n_data = grid.num_data_res
    y = np.zeros((batch, 1, N, T, F), np.complex64)
    bits_out = np.zeros((batch, U, n_data * max_bits), np.uint8)
    h_out = np.zeros((batch, U, F, T, N), np.complex64)
    no_out = np.zeros(batch, np.float32)
    f_hz = (np.arange(F) - F / 2.0) * scs

    for b in range(batch):
        rng = np.random.default_rng([seed, b])
        no = ebnodb2no(ebno[b], bits_ue[0], mcs_code_rate(cfg.mcs_index[mcs_per_ue[0]], cfg.mcs_table), grid)
        no_out[b] = no
        rx = np.zeros((N, T, F), np.complex128)
        for u in range(U):
            # ---- transmit grid -------------------------------------------------------
            x = np.zeros((T, F), np.complex128)
            nb = bits_ue[u]
            bb = rng.integers(0, 2, size=(n_data, nb), dtype=np.uint8)
            if coded_bits is not None:
                bb = np.asarray(coded_bits[b, u], np.uint8)[:n_data * nb].reshape(n_data, nb)
            const = qam_constellation(nb)
            idx = np.zeros(n_data, np.int64)
            for k in range(nb):
                idx = (idx << 1) | bb[:, k]
            x[data_mask] = const[idx]                 # row-major (t, f) order == demapper order
            for j, l in enumerate(grid.dmrs_symbols):
                x[l, :] = grid.pilots[u, j * F:(j + 1) * F]
            bits_out[b, u, :n_data * nb] = bb.reshape(-1)
            # ---- channel -------------------------------------------------------------
            ds = delay_spread_ns[u % len(delay_spread_ns)] * 1e-9
            fd = doppler_hz[u % len(doppler_hz)]
            if sparse_paths is None:
                n_taps = 24
                tau = np.sort(rng.exponential(ds, size=n_taps))
                tau -= tau[0]
                pw = np.full(n_taps, 1.0 / n_taps)
            else:
                n_taps = int(sparse_paths)
                tau = np.sort(rng.uniform(0.0, 6.0 * ds, size=n_taps))
                tau -= tau[0]
                pw = 10.0 ** (-rng.uniform(0.0, 25.0, size=n_taps) / 10.0)
                pw /= pw.sum()
            hu = np.zeros((N, T, F), np.complex128)
            for a in range(N):
                g = (rng.standard_normal(n_taps) + 1j * rng.standard_normal(n_taps)) * np.sqrt(pw / 2.0)
                nu = fd * np.cos(rng.uniform(0, 2 * np.pi, size=n_taps))
                rot = np.exp(2j * np.pi * nu[None, :] * (np.arange(T)[:, None] * t_sym))      # [T, taps]
                ph = np.exp(-2j * np.pi * f_hz[:, None] * tau[None, :])                       # [F, taps]
                hu[a] = (rot * g[None, :]) @ ph.T
            if per_ue_power_norm:
                hu /= np.sqrt(np.mean(np.abs(hu) ** 2))
            h_out[b, u] = np.transpose(hu, (2, 1, 0)).astype(np.complex64)
            if act[b, u] > 0:
                rx += hu * x[None, :, :]
        noise = (rng.standard_normal((N, T, F)) + 1j * rng.standard_normal((N, T, F))) * np.sqrt(no / 2.0)
        y[b, 0] = (rx + noise).astype(np.complex64)

    return SlotBatch(y, act, bits_out, np.tile(np.asarray(bits_ue, np.int32), (batch, 1)), h_out, no_out)


def uncoded_ber(llr: np.ndarray, bits: np.ndarray, active_tx: np.ndarray, num_bits: int) -> float:
    """Hard-decision bit error rate over active UEs (``llr > 0  <=>  bit 1``, utils/neural_rx.py:864)."""
    n = llr.shape[-1]
    hard = (llr > 0).astype(np.uint8)
    err = (hard != bits[..., :n]).astype(np.float64)
    w = np.broadcast_to(active_tx[..., None], err.shape)
    return float((err * w).sum() / max(w.sum(), 1.0))


def aerial_inputs(sb, grid):
    """Re-shape a synthetic slot batch into the seven inputs of the Aerial / TensorRT-shaped
    receiver (what ``utils/onnx_utils.py:376-409`` does with the Sionna tensors): rx_slot
    [B,F,T,N_rx] real / imag, raw LS estimates ``y_p / p`` at every user's non-zero pilots
    [B,n_pilots,U,N_rx] (DMRS-symbol major, zero pilots dropped), the active-port mask and the DMRS
    symbol / in-PRB subcarrier positions."""
    y = np.transpose(sb.y[:, 0], (0, 3, 2, 1))                                # [B,F,T,N]
    U, F = grid.num_tx, grid.num_subcarriers
    ofdm_pos = np.tile(np.asarray(grid.dmrs_symbols, np.int32), (U, 1))
    hs, sc_pos = [], []
    for u in range(U):
        h_sym = []
        nz0 = None
        for j, l in enumerate(grid.dmrs_symbols):
            p = grid.pilots[u, j * F:(j + 1) * F]
            nz = np.flatnonzero(np.abs(p) > 0)
            nz0 = nz if nz0 is None else nz0
            h_sym.append(sb.y[:, 0, :, l, :][:, :, nz] / p[nz])               # [B,N,n_nz]
        hs.append(np.concatenate(h_sym, axis=-1))                             # [B,N,n_pilots]
        sc_pos.append(nz0[nz0 < 12])
    h = np.transpose(np.stack(hs, axis=1), (0, 3, 1, 2))                      # [B,n_pilots,U,N]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return [f32(y.real), f32(y.imag), f32(h.real), f32(h.imag), f32(sb.active_tx), ofdm_pos,
            np.asarray(sc_pos, np.int32)]
