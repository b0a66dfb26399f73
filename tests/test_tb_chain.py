"""Transport-block chain of the BLER harness (neural_rx_b200/tb.py) on the CPU: CRC known answers,
TBS rules, segmentation bookkeeping, LDPC encoding (H w = 0) on the structural stand-in base graphs, rate matching /
interleaving / scrambling round trips, the BP decoder, and the loader for base-graph files in Sionna's layout."""
import binascii
import os

import numpy as np
import pytest

from neural_rx_b200 import tb as TB
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid


def _bits_of(data: bytes) -> np.ndarray:
    return np.unpackbits(np.frombuffer(data, np.uint8))


def _crc_bitwise(bits, L, poly):
    reg = 0
    for b in list(bits) + [0] * L:
        top = (reg >> (L - 1)) & 1
        reg = ((reg << 1) & ((1 << L) - 1)) | int(b)
        if top:
            reg ^= poly
    return reg


@pytest.mark.parametrize("kind,check", [("24A", 0xCDE703), ("24B", 0x23EF52), ("16", 0x31C3)])
def test_crc_known_answers(kind, check):
    """CRC-24/LTE-A, CRC-24/LTE-B and CRC-16/XMODEM check values of "123456789" (zero initial state, no
    reflection — the TS 38.212 §5.1 definition); CRC16 also against binascii's CCITT implementation; every
    polynomial against a bit-serial shift register on random messages."""
    msg = _bits_of(b"123456789")
    L = TB._CRC_POLY[kind][0]
    val = int("".join(map(str, TB.crc_parity(msg, kind))), 2)
    assert val == check
    if kind == "16":
        assert val == binascii.crc_hqx(b"123456789", 0)
    rng = np.random.default_rng(1)
    for k in TB._CRC_POLY:
        Lk, poly = TB._CRC_POLY[k]
        m = rng.integers(0, 2, 257, dtype=np.uint8)
        assert int("".join(map(str, TB.crc_parity(m, k))), 2) == _crc_bitwise(m, Lk, poly)
        assert TB.crc_check(TB.crc_attach(m, k), k)
        bad = TB.crc_attach(m, k)
        bad[5] ^= 1
        assert not TB.crc_check(bad, k)
    assert L in (16, 24)


def test_tb_size_rules():
    """TS 38.214 §5.1.3.2: table branch for N_info <= 3824, formula branch above; byte aligned; with the CRCs the
    block splits into equal code blocks; grows with PRBs / rate / modulation."""
    assert TB.tb_size(1, 144, 2, 120 / 1024) in TB._TBS_TABLE
    assert TB.tb_size(1, 144, 2, 679 / 1024) == 184           # N_info = 190.97 -> n = 3, N' = 8 floor(190.97 / 8) = 184
    prev = 0
    for n_prb in (1, 4, 16, 52, 106, 132, 273):
        for qm, r in ((2, 679), (4, 553), (6, 517)):
            t = TB.tb_size(n_prb, 144, qm, r / 1024)
            assert t % 8 == 0 and t > 0
            enc = TB.TBEncoder(t, n_prb * 144 * qm, r / 1024, qm, base_graph="standin")
            assert enc.num_cbs * enc.k_prime == t + (24 if t > 3824 else 16) + (24 * enc.num_cbs if enc.num_cbs > 1 else 0)
            assert enc.k >= enc.k_prime and sum(enc.cb_e) == n_prb * 144 * qm
            assert abs(t / (n_prb * 144 * qm) - r / 1024) < 0.05 + 40 / (n_prb * 144 * qm)
        t16 = TB.tb_size(n_prb, 144, 4, 553 / 1024)
        assert t16 > prev
        prev = t16
    # nrx_large at 132 PRB, MCS 14: N_info = 132*144*4*553/1024 = 41060.25 -> n = 10, N' = 40960 -> C = 5 -> TBS = 40976
    assert TB.tb_size(132, 144, 4, 553 / 1024) == 40976


@pytest.mark.parametrize("bg,z", [(1, 2), (1, 44), (1, 384), (2, 6), (2, 52), (2, 384)])
def test_ldpc_encoder_satisfies_all_checks(bg, z):
    i_ls = next(i for i, zs in enumerate(TB.lifting_sets()) if z in zs)
    code = TB.LdpcCode(bg, z, i_ls, "standin")
    rng = np.random.default_rng(bg * 1000 + z)
    c = rng.integers(0, 2, (3, code.k), dtype=np.uint8)
    w = code.encode(c)
    assert w.shape == (3, code.cols * z) and np.array_equal(w[:, :code.k], c)
    assert code.syndrome_ok(w).all()
    w[0, 7] ^= 1
    assert not code.syndrome_ok(w)[0]


def test_lifting_size_selection():
    sets = TB.lifting_sets()
    assert sorted(z for zs in sets for z in zs)[:6] == [2, 3, 4, 5, 6, 7] and max(max(zs) for zs in sets) == 384
    assert sum(len(zs) for zs in sets) == 51
    assert TB.select_lifting(22, 8448) == (384, 1)
    assert TB.select_lifting(22, 8219) == (384, 1) and TB.select_lifting(22, 7744) == (352, 5)
    assert TB.select_lifting(10, 640)[0] == 64


@pytest.mark.parametrize("tbs,G,qm,rate", [(192, 576, 2, 0.66), (1544, 3168, 4, 0.54), (40976, 76032, 4, 0.54), (424, 6000, 6, 0.1)])
def test_tb_round_trip(tbs, G, qm, rate):
    """encode -> ideal LLRs -> decode returns the bits with CRC status True; a few dB of noise is corrected; random
    LLRs fail the CRC.  (424 bits in 6000 coded bits: repetition beyond the circular buffer.)"""
    enc = TB.TBEncoder(tbs, G, rate, qm, n_rnti=[1, 7], n_id=[1, 99], base_graph="standin")
    dec = TB.TBDecoder(enc, num_bp_iter=20, cn_type="boxplus")
    rng = np.random.default_rng(tbs)
    bits = rng.integers(0, 2, (2, 2, tbs), dtype=np.uint8)
    coded = enc(bits)
    assert coded.shape == (2, 2, G) and coded.dtype == np.uint8
    assert not np.array_equal(coded[:, 0], coded[:, 1])
    llr = 8.0 * (2.0 * coded.astype(np.float32) - 1.0)
    b_hat, ok = dec(llr)
    assert ok.all() and np.array_equal(b_hat, bits)
    noisy = 2.0 * (2.0 * coded - 1.0 + 0.55 * rng.standard_normal(coded.shape)) / 0.55 ** 2
    b_hat, ok = dec(noisy.astype(np.float32))
    assert ok.all() and np.array_equal(b_hat, bits)
    _, ok = dec(rng.standard_normal(coded.shape).astype(np.float32))
    assert not ok.any()
    import torch
    b_t, ok_t = dec(torch.as_tensor(noisy.astype(np.float32)))
    assert np.array_equal(b_t.numpy(), bits) and bool(ok_t.all())
    ms = TB.TBDecoder(enc, num_bp_iter=20, cn_type="minsum")
    b_hat, ok = ms(llr)
    assert ok.all() and np.array_equal(b_hat, bits)


def test_scrambler_and_rate_matching_details():
    enc = TB.TBEncoder(1544, 3168, 0.54, 4, base_graph="standin")
    plain = TB.TBEncoder(1544, 3168, 0.54, 4, use_scrambler=False, base_graph="standin")
    bits = np.random.default_rng(3).integers(0, 2, (1, 1544), dtype=np.uint8)
    from neural_rx_b200.pusch import gold_sequence
    assert np.array_equal(enc(bits) ^ plain(bits), gold_sequence((1 << 15) + 1, 3168)[None].astype(np.uint8))
    # no filler position is ever transmitted; the interleaver is a permutation
    Z = enc.z
    for sel, perm, E in zip(enc.cb_sel, enc.cb_perm, enc.cb_e):
        assert not np.any((sel >= enc.k_prime - 2 * Z) & (sel < enc.k - 2 * Z)) and sel.max() < enc.n_cb
        assert np.array_equal(np.sort(perm), np.arange(E))
    with pytest.raises(ValueError):
        enc(np.zeros((1, 100), np.uint8))


def test_base_graph_file_loader(tmp_path, monkeypatch):
    """load_base_graph reads files in the layout Sionna ships (sionna/fec/ldpc/codes/5G_bg1.csv): two header lines,
    then `row;col;s_0;...;s_7` with the row index left blank on continuation lines."""
    mats = [TB.standin_base_graph(2, i) for i in range(8)]
    lines = ["BG2;;;;;;;;;", "row;col;0;1;2;3;4;5;6;7"]
    for r in range(42):
        first = True
        for c in range(52):
            if mats[0][r, c] >= 0 or any(m[r, c] >= 0 for m in mats):
                vals = ";".join(str(int(m[r, c])) for m in mats)
                lines.append(f"{r if first else ''};{c};{vals}")
                first = False
    (tmp_path / "5G_bg2.csv").write_text("\n".join(lines) + "\n")
    monkeypatch.setenv("NRX_LDPC_BG_DIR", str(tmp_path))
    TB.load_base_graph.cache_clear()
    # the stand-in matrices of different lifting sets have different supports; compare on set 3's support
    got = TB.load_base_graph(2, 3)
    assert np.array_equal(got[mats[3] >= 0], mats[3][mats[3] >= 0])
    bm, src = TB.base_graph(2, 3, "auto")
    assert src == "3gpp"
    TB.load_base_graph.cache_clear()
    monkeypatch.delenv("NRX_LDPC_BG_DIR")
    if TB._bg_csv_path(1) is None:
        with pytest.raises(TB.BaseGraphUnavailable):
            TB.base_graph(1, 0, "3gpp")
        assert TB.base_graph(1, 0, "auto")[1] == "standin"


@pytest.mark.parametrize("label,n_prb", [("nrx_rt", 4), ("nrx_large", 132), ("nrx_large_64qam", 132), ("nrx_large_qpsk", 273)])
def test_pusch_tb_encoder_of_the_configs(label, n_prb):
    cfg = get_config(label)
    grid = build_grid(cfg, n_size_bwp=n_prb)
    enc = TB.pusch_tb_encoder(cfg, grid, base_graph="standin")
    assert enc.num_coded_bits == grid.num_data_res * cfg.num_bits_per_symbol[0]
    assert grid.num_data_res == n_prb * 144
    assert enc.bg == 1 or enc.tb_size <= 3824
    assert abs(enc.tb_size / enc.num_coded_bits - enc.coderate) < 0.02


@pytest.mark.parametrize("n_prb,qm,rate,bg,num_cbs", [
    (100, 2, 0.2, 2, 2),        # rate <= 0.25: BG2 although the block is long (K_cb = 3840), two code blocks
    (273, 2, 0.12, 2, 3),       # repetition: E_r = 26 208 > N_cb = 16 000
    (35, 4, 0.3, 1, 1),         # one BG1 block with fillers (K' = 6 040, K = 6 336)
    (66, 6, 0.55, 1, 4),
    (51, 6, 0.9, 1, 5),         # heavy puncturing, unequal E_r (8 808 / 8 814)
])
def test_segmentation_and_rate_matching_regimes(n_prb, qm, rate, bg, num_cbs):
    """Every branch of TS 38.212 §6.2.2-6.2.5 the PUSCH configs can reach, on sizes from the TBS rule: base-graph
    choice, multi-block segmentation with per-block CRC, fillers, repetition and puncturing, unequal E_r; the
    noiseless round trip returns the bits and a passed CRC."""
    tbs = TB.tb_size(n_prb, 144, qm, rate)
    G = n_prb * 144 * qm
    enc = TB.TBEncoder(tbs, G, rate, qm, base_graph="standin")
    assert (enc.bg, enc.num_cbs) == (bg, num_cbs)
    assert enc.k == (22 if bg == 1 else 10) * enc.z and enc.k_prime <= enc.k and enc.n_cb == (66 if bg == 1 else 50) * enc.z
    assert sum(enc.cb_e) == G and all(e % qm == 0 for e in enc.cb_e) and max(enc.cb_e) - min(enc.cb_e) in (0, qm)
    bits = np.random.default_rng(n_prb).integers(0, 2, (2, tbs), dtype=np.uint8)
    coded = enc(bits)
    b_hat, ok = TB.TBDecoder(enc)(6.0 * (2.0 * coded - 1.0).astype(np.float32))
    assert ok.all() and np.array_equal(b_hat, bits)
    flipped = coded.copy()
    flipped[0, : G // 3] ^= 1                                   # a third of block 0 wrong: the TB CRC must fail
    _, ok = TB.TBDecoder(enc, num_bp_iter=4)(6.0 * (2.0 * flipped - 1.0).astype(np.float32))
    assert not ok[0] and ok[1]


@pytest.mark.parametrize("tbs,G,qm,rate,kb_cols", [(24, 120, 2, 0.2, 6), (504, 1200, 2, 0.42, 8), (608, 1400, 4, 0.43, 9), (3824, 8000, 4, 0.48, 10)])
def test_short_blocks_on_bg2(tbs, G, qm, rate, kb_cols):
    """BG2 with K_b = 6 / 8 / 9 / 10 systematic columns in use (TS 38.212 §5.2.2): the unused columns are fillers."""
    enc = TB.TBEncoder(tbs, G, rate, qm, base_graph="standin")
    assert enc.bg == 2 and enc.tb_crc == "16" and enc.num_cbs == 1
    assert enc.z == TB.select_lifting(kb_cols, enc.k_prime)[0] and enc.k == 10 * enc.z
    bits = np.random.default_rng(tbs).integers(0, 2, (3, tbs), dtype=np.uint8)
    b_hat, ok = TB.TBDecoder(enc)(5.0 * (2.0 * enc(bits) - 1.0).astype(np.float32))
    assert ok.all() and np.array_equal(b_hat, bits)


def test_tb_size_matches_the_reference_notebook_printout():
    """The one transport-block size the reference tree shows: `model._transmitters[0].show()` in
    notebooks/jumpstart_tutorial.ipynb (nrx_rt training grid: 4 PRB, MCS 14, mcs_table 1) prints num_res_per_prb : 144,
    num_coded_bits : 2304, target_coderate : 0.5400390625, tb_size : 1256, n_rnti : 1, n_id : 1."""
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=4)
    enc = TB.pusch_tb_encoder(cfg, grid, base_graph="standin")
    assert grid.num_data_res // 4 == 144
    assert enc.num_coded_bits == 2304 and enc.coderate == 0.5400390625 and enc.tb_size == 1256
    assert TB.tb_size(4, 144, 4, 0.5400390625) == 1256
    assert np.array_equal(enc.scramble_seq[0], TB.gold_sequence((1 << 15) + 1, 2304))


def test_base_graph_files_of_an_installed_sionna_are_found_without_importing_it(tmp_path, monkeypatch):
    """A package named `sionna` on the path whose __init__ would raise if executed: its data file is still found."""
    pkg = tmp_path / "sionna"
    codes = pkg / "fec" / "ldpc" / "codes"
    codes.mkdir(parents=True)
    (pkg / "__init__.py").write_text("raise RuntimeError('sionna must not be imported by the TB chain')\n")
    (codes / "5G_bg2.csv").write_text("x\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    monkeypatch.delenv("NRX_LDPC_BG_DIR", raising=False)
    import importlib
    importlib.invalidate_caches()
    assert TB._bg_csv_path(2) == str(codes / "5G_bg2.csv")
    assert TB._bg_csv_path(1) is None
    import sys
    assert "sionna" not in sys.modules


def test_published_curves_of_the_reference_decode():
    """tests/golden/ref_published_curves.json = the reference's results/*_results decoded by tools/ref_results.py:
    the nrx_large 2-UE BLER curve falls from ~1 at -2 dB to 0 at 6 dB, more CGNN iterations never hurt at 2 dB
    (results/nrx_large_sweep_results), and the lookup used by tools/bler_sweep.py --published finds those points.  When
    /root/reference is present the committed JSON must equal a fresh decode."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    from tools import ref_results as R
    pub = R.load_published()
    c = pub["nrx_large_results"]
    assert c["ebno_db"] == [float(e) for e in range(-2, 7)]
    bler = c["bler"]["Neural Receiver|2|0"]
    assert bler[0] > 0.99 and bler[-1] == 0.0 and all(a >= b for a, b in zip(bler, bler[1:]))
    at2 = [R.published_at("nrx_large_sweep_results", f"Neural Receiver|2|0|{n}", 2.0) for n in range(1, 9)]
    assert all(a >= b for a, b in zip(at2, at2[1:])) and abs(at2[0] - 0.41205) < 1e-9 and abs(at2[-1] - 0.15865) < 1e-9
    assert R.published_at("nrx_large_results", "Neural Receiver|2|0", 2.5) is None
    assert R.published_at("no_such_file", "x", 0.0) is None
    ref_dir = "/root/reference/results"
    if os.path.isdir(ref_dir):
        assert {f: R.decode(os.path.join(ref_dir, f)) for f in sorted(os.listdir(ref_dir))} == pub
