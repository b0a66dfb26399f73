"""Shared helpers of the test-suite: oracle construction from a receiver config, staged weight
files, parity metrics.  (Test infrastructure — the only place besides bench/smoke that imports
``oracle``.)"""
from __future__ import annotations

import os

import numpy as np

from neural_rx_b200.config import NrxConfig, get_config
from neural_rx_b200.weights import NrxWeights, load_weights, random_weights
from oracle import nrx_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WEIGHT_DIRS = [os.path.join(ROOT, "weights")]      # staged (git-ignored) copies of the reference's weight files

#: emulation of where the CUDA engine rounds (fp16 operands everywhere, fp32 accumulate in GEMMs)
ENGINE_EMU = O.Emulation(act_fp16=True, weight_fp16=True, dw_weight_fp16=True, dw_acc_fp16=True, state_fp16=True)


def weight_path(label: str):
    for d in WEIGHT_DIRS:
        p = os.path.join(d, f"{label}_weights")
        if os.path.exists(p):
            return p
    return None


def get_weights(cfg: NrxConfig, prefer_real: bool = True, seed: int = 0) -> tuple[NrxWeights, str]:
    p = weight_path(cfg.label) if prefer_real else None
    if p is not None:
        return load_weights(cfg, p), "shipped"
    return random_weights(cfg, seed=seed), "random"


def oracle_arch(cfg: NrxConfig) -> O.OracleArch:
    return O.OracleArch(num_rx_ant=cfg.num_rx_antennas, d_s=cfg.d_s, num_it=cfg.num_nrx_iter,
                        num_units_init=tuple(cfg.num_units_init), num_units_agg=tuple(cfg.num_units_agg),
                        num_units_state=tuple(cfg.num_units_state), num_units_readout=tuple(cfg.num_units_readout),
                        num_bits_per_symbol=tuple(cfg.num_bits_per_symbol), var_mcs_masking=cfg.mcs_var_mcs_masking)


def oracle_net(cfg: NrxConfig, weights: NrxWeights, dtype=None):
    import torch
    return O.bind_weights(oracle_arch(cfg), weights.to_list(), dtype or torch.float32)


def rel_l2(a: np.ndarray, b: np.ndarray) -> float:
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def sign_agreement(a: np.ndarray, b: np.ndarray) -> float:
    return float(np.mean((np.asarray(a) > 0) == (np.asarray(b) > 0)))
