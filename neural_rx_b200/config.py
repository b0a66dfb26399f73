"""Receiver configuration: reader for the reference's ``config/*.cfg`` files plus built-in presets.

The reference parses its INI files with ``configparser.RawConfigParser`` and ``eval``s every value
into an attribute of ``Parameters`` (``utils/parameters.py:91-113``); in inference mode the
``*_eval`` keys overwrite their training twins (``utils/parameters.py:118-127``).  This module
mirrors exactly those two rules for the keys the CGNN hot path reads
(``utils/neural_rx.py:1445-1460, 638-662``) and nothing else (no Sionna objects are built).

Presets for the five BASELINE.json configs are embedded so that nothing under ``/root/reference``
is needed at run time (the GPU box does not have it).
"""
from __future__ import annotations

import configparser
import dataclasses
import os
from typing import List, Optional, Sequence

# 3GPP TS 38.214 Table 5.1.3.1-1 (mcs_table = 1): MCS index -> (bits per symbol, code rate x 1024)
_MCS_TABLE_1 = (
    [(2, r) for r in (120, 157, 193, 251, 308, 379, 449, 526, 602, 679)]
    + [(4, r) for r in (340, 378, 434, 490, 553, 616, 658)]
    + [(6, r) for r in (438, 466, 517, 567, 616, 666, 719, 772, 822, 873, 910, 948)]
)


def mcs_bits_per_symbol(mcs_index: int, mcs_table: int = 1) -> int:
    if mcs_table != 1:
        raise NotImplementedError("only mcs_table = 1 is used by the reference configs")
    return _MCS_TABLE_1[mcs_index][0]


def mcs_code_rate(mcs_index: int, mcs_table: int = 1) -> float:
    if mcs_table != 1:
        raise NotImplementedError("only mcs_table = 1 is used by the reference configs")
    return _MCS_TABLE_1[mcs_index][1] / 1024.0


class _DtypeNamespace:
    """Stands in for ``tf`` / ``torch`` when eval-ing ``nrx_dtype = tf.float32`` (cfg line 74/75)."""

    float32 = "float32"
    float16 = "float16"
    float64 = "float64"


@dataclasses.dataclass
class NrxConfig:
    """The subset of ``Parameters`` attributes that defines the neural receiver hot path."""

    label: str = "nrx_rt"
    # [system]
    n_size_bwp: int = 132            # PRBs (already the *_eval value in inference mode)
    num_rx_antennas: int = 4
    mcs_index: Sequence[int] = (14,)
    mcs_table: int = 1
    num_ofdm_symbols: int = 14       # symbol_allocation = [0, 14]
    dmrs_symbols: Sequence[int] = (2, 11)   # type-A pos 2, additional_position 1, length 1
    dmrs_port_sets: Sequence[Sequence[int]] = ((0,), (2,))
    num_cdm_groups_without_data: int = 2
    n_scid: int = 1                  # DMRS scrambling: the pilot VALUES follow from these four (config/nrx_rt.cfg:22-37)
    dmrs_nid: Sequence[Sequence[int]] = ((1, 1), (1, 1))    # per transmitter: N_ID^0, N_ID^1
    slot_number: int = 0
    n_start_grid: int = 0
    # [neural_receiver]
    num_nrx_iter: int = 2
    num_nrx_iter_eval: int = 2
    d_s: int = 56
    num_units_init: Sequence[int] = (128, 128)
    num_units_agg: Sequence[Sequence[int]] = ((64,), (64,))
    num_units_state: Sequence[Sequence[int]] = ((128, 128), (128, 128))
    num_units_readout: Sequence[int] = (128,)
    max_num_tx: int = 2
    initial_chest: Optional[str] = "ls"
    mask_pilots: bool = False
    mcs_var_mcs_masking: bool = False
    layer_type_dense: str = "dense"
    layer_type_conv: str = "sepconv"
    layer_type_readout: str = "dense"
    nrx_dtype: str = "float32"
    # [evaluation]
    batch_size_eval: int = 30
    snr_db_eval_min: float = -2.0
    snr_db_eval_max: float = 8.0
    snr_db_eval_stepsize: float = 1.0
    channel_norm: bool = False

    # ---- derived -------------------------------------------------------------------------
    @property
    def num_subcarriers(self) -> int:
        return 12 * self.n_size_bwp

    @property
    def num_bits_per_symbol(self) -> List[int]:
        return [mcs_bits_per_symbol(m, self.mcs_table) for m in self.mcs_index]

    @property
    def num_mcss_supported(self) -> int:
        return len(self.mcs_index)

    @property
    def num_io_stacks(self) -> int:
        """Number of StateInit / ReadoutLLRs copies (utils/neural_rx.py:445-454, 487-496)."""
        return 1 if self.mcs_var_mcs_masking else len(self.mcs_index)

    @property
    def readout_bits(self) -> List[int]:
        """Output width of each LLR readout head (max bits in masking mode)."""
        b = self.num_bits_per_symbol
        return [max(b)] if self.mcs_var_mcs_masking else list(b)

    def validate(self) -> None:
        if self.layer_type_conv != "sepconv" or self.layer_type_dense != "dense" \
                or self.layer_type_readout != "dense":
            # same error the reference raises (utils/neural_rx.py:93)
            raise NotImplementedError("Unknown layer_type selected.")
        if self.initial_chest not in ("ls",):
            raise NotImplementedError("the B200 engine implements initial_chest = 'ls' only")
        if self.mask_pilots:
            # utils/neural_rx.py:1476-1479
            raise ValueError("Cannot use initial channel estimator if pilots are masked.")
        if not (1 <= self.num_nrx_iter_eval <= self.num_nrx_iter):
            raise ValueError("Invalid number of iterations")
        if len(self.num_units_agg) != self.num_nrx_iter or len(self.num_units_state) != self.num_nrx_iter:
            raise ValueError("num_units_agg / num_units_state need one entry per iteration")


def _literal(expr: str):
    """Value of one cfg entry.  The reference ``eval``s every value (utils/parameters.py:104-110), i.e. executes
    whatever the file says; the cfg files only ever hold Python literals plus a dtype name (``tf.float32`` /
    ``torch.float32``), so this parser accepts exactly that and nothing executable."""
    import ast
    import re
    text = expr.strip()
    m = re.fullmatch(r"(?:tf|torch)\.([A-Za-z_][A-Za-z0-9_]*)", text)
    if m:
        return getattr(_DtypeNamespace, m.group(1))
    try:
        return ast.literal_eval(text)
    except (ValueError, SyntaxError) as exc:
        raise ValueError(f"cfg value {expr!r} is not a Python literal") from exc


def parse_cfg_text(text: str, training: bool = False) -> NrxConfig:
    """Parse a reference-format cfg (INI + Python expressions) into an :class:`NrxConfig`."""
    cp = configparser.RawConfigParser()
    cp.read_string(text)
    raw = {}
    for section in cp.sections():
        for option in cp.options(section):
            raw[option] = _literal(cp.get(section, option))
    if not training:   # utils/parameters.py:118-127
        for key in ("n_size_bwp", "channel_norm"):
            if f"{key}_eval" in raw:
                raw[key] = raw[f"{key}_eval"]
    sym = raw.get("symbol_allocation", [0, 14])
    if raw.get("dmrs_mapping_type", "A") != "A" or raw.get("dmrs_config_type", 1) != 1 \
            or raw.get("dmrs_length", 1) != 1:
        raise NotImplementedError("only DMRS mapping type A, config type 1, length 1 are supported")
    l0 = raw.get("dmrs_type_a_position", 2)
    add = raw.get("dmrs_additional_position", 1)
    # TS 38.211 Table 6.4.1.1.3-3, mapping type A, single-symbol DMRS, 14-symbol allocation
    dmrs_syms = {0: [l0], 1: [l0, 11], 2: [l0, 7, 11], 3: [l0, 5, 8, 11]}[add]
    cfg = NrxConfig(
        label=raw.get("label", "nrx"),
        n_size_bwp=int(raw["n_size_bwp"]),
        num_rx_antennas=int(raw["num_rx_antennas"]),
        mcs_index=tuple(raw["mcs_index"]),
        mcs_table=int(raw.get("mcs_table", 1)),
        num_ofdm_symbols=int(sym[1]),
        dmrs_symbols=tuple(dmrs_syms),
        dmrs_port_sets=tuple(tuple(p) for p in raw["dmrs_port_sets"]),
        num_cdm_groups_without_data=int(raw.get("num_cdm_groups_without_data", 2)),
        n_scid=int(raw.get("n_scid", 1)),
        dmrs_nid=tuple(tuple(int(v) for v in x) for x in raw.get("dmrs_nid", ((1, 1), (1, 1)))),
        slot_number=int(raw.get("slot_number", 0)),
        n_start_grid=int(raw.get("n_start_grid", 0)),
        num_nrx_iter=int(raw["num_nrx_iter"]),
        num_nrx_iter_eval=int(raw.get("num_nrx_iter_eval", raw["num_nrx_iter"])),
        d_s=int(raw["d_s"]),
        num_units_init=tuple(raw["num_units_init"]),
        num_units_agg=tuple(tuple(x) for x in raw["num_units_agg"]),
        num_units_state=tuple(tuple(x) for x in raw["num_units_state"]),
        num_units_readout=tuple(raw["num_units_readout"]),
        max_num_tx=int(raw["max_num_tx"]),
        initial_chest=None if raw.get("initial_chest") in (None, "None") else raw.get("initial_chest"),
        mask_pilots=bool(raw.get("mask_pilots", False)),
        mcs_var_mcs_masking=bool(raw.get("mcs_var_mcs_masking", False)),
        layer_type_dense=raw.get("layer_type_dense", "dense"),
        layer_type_conv=raw.get("layer_type_conv", "sepconv"),
        layer_type_readout=raw.get("layer_type_readout", "dense"),
        nrx_dtype=str(raw.get("nrx_dtype", "float32")),
        batch_size_eval=int(raw.get("batch_size_eval", 30)),
        snr_db_eval_min=float(raw.get("snr_db_eval_min", -2)),
        snr_db_eval_max=float(raw.get("snr_db_eval_max", 8)),
        snr_db_eval_stepsize=float(raw.get("snr_db_eval_stepsize", 1)),
        channel_norm=bool(raw.get("channel_norm", False)),
    )
    return cfg


def load_cfg(path: str, training: bool = False) -> NrxConfig:
    """``Parameters(config_name, training=False, system='dummy')`` equivalent for the hot path."""
    if not os.path.exists(path):
        raise FileNotFoundError("Unknown config file.")   # utils/parameters.py:99
    with open(path, "r", encoding="utf-8") as f:
        return parse_cfg_text(f.read(), training=training)


def _rt(**kw) -> NrxConfig:
    return dataclasses.replace(NrxConfig(), **kw)


def _large(**kw) -> NrxConfig:
    base = NrxConfig(
        label="nrx_large", num_nrx_iter=8, num_nrx_iter_eval=8,
        num_units_agg=tuple((64,) for _ in range(8)),
        num_units_state=tuple((128, 128) for _ in range(8)),
        snr_db_eval_max=7.0,
    )
    return dataclasses.replace(base, **kw)


#: presets equal to what ``load_cfg`` returns for the reference's cfg files (checked in tests
#: whenever /root/reference is present)
PRESETS = {
    "nrx_rt": _rt(label="nrx_rt"),
    "nrx_rt_64qam": _rt(label="nrx_rt_64qam", mcs_index=(19,), snr_db_eval_max=12.0),
    "nrx_rt_var_mcs": _rt(label="nrx_rt_var_mcs", mcs_index=(9, 14), snr_db_eval_min=-3.0),
    "nrx_large": _large(),
    "nrx_large_qpsk": _large(label="nrx_large_qpsk", mcs_index=(9,), snr_db_eval_min=-3.0, snr_db_eval_max=8.0),
    "nrx_large_64qam": _large(label="nrx_large_64qam", mcs_index=(19,), snr_db_eval_max=10.0),
    "nrx_large_var_mcs": _large(label="nrx_large_var_mcs", mcs_index=(9, 14), snr_db_eval_min=-3.0,
                                snr_db_eval_max=8.0),
    "nrx_large_var_mcs_64qam_masking": _large(
        label="nrx_large_var_mcs_64qam_masking", mcs_index=(9, 14, 19), mcs_var_mcs_masking=True,
        snr_db_eval_min=-3.0, snr_db_eval_max=10.0),
    "nrx_site_specific": _rt(label="nrx_site_specific", channel_norm=True, snr_db_eval_min=-3.0,
                             snr_db_eval_max=16.0),
    "nrx_site_specific_large": _large(label="nrx_site_specific_large", channel_norm=True,
                                      snr_db_eval_min=-3.0, snr_db_eval_max=16.0),
}


def get_config(name_or_path: str, training: bool = False) -> NrxConfig:
    """Resolve a preset name (``"nrx_rt"``, ``"nrx_rt.cfg"``) or a cfg file path."""
    if os.path.exists(name_or_path):
        return load_cfg(name_or_path, training=training)
    key = os.path.basename(name_or_path).replace(".cfg", "")
    if key in PRESETS:
        return dataclasses.replace(PRESETS[key])
    raise FileNotFoundError("Unknown config file.")
