"""5G-NR PUSCH transport-block chain for the BLER harness: what sits on either side of the
neural receiver in the reference's evaluation loop.

The reference takes all of it from Sionna 0.18 (``requirements.txt:1``): ``PUSCHTransmitter`` owns a
``TBEncoder`` (``utils/neural_rx.py:1402-1405``), ``NeuralPUSCHReceiver`` wraps it in a
``TBDecoder(encoder, num_bp_iter, cn_type)`` (``:1407-1413``) and applies it to the receiver's LLRs (``:1600``);
``scripts/evaluate.py:193-202`` counts block errors from the returned CRC status.  Sionna is not installed here,
so this module restates the published procedures the two classes implement:

* transport-block size — TS 38.214 §5.1.3.2 / §6.1.4.2 (``tb_size``);
* CRC attachment, code-block segmentation, LDPC base-graph / lifting-size selection — TS 38.212 §5.1, §5.2.2, §6.2.1-6.2.3;
* LDPC encoding from a base graph — TS 38.212 §5.3.2 (any base matrix with an invertible core parity block);
* rate matching (rv 0, no LBRM), bit interleaving, code-block concatenation — TS 38.212 §5.4.2, §5.5;
* scrambling — TS 38.211 §6.3.1.1 (Gold sequence, ``c_init = n_RNTI 2^15 + n_ID``);
* decoding: the inverse steps, flooding belief propagation with exact box-plus check nodes
  (``cn_type = 'boxplus'``, ``num_bp_iter = 20``: ``config/nrx_rt.cfg:46-47``), CRC checks.

**What is NOT here: the base-graph shift tables of TS 38.212 Tables 5.3.2-2 / 5.3.2-3** (316 + 197 entries x 8
lifting sets).  They exist neither in the reference tree nor in this image, and they are not reproduced from
memory.  ``load_base_graph`` reads them from Sionna's package data (``sionna/fec/ldpc/codes/5G_bg{1,2}.csv``) when
Sionna is installed, or from ``$NRX_LDPC_BG_DIR`` holding files in that same layout; without them
``standin_base_graph`` builds a seeded quasi-cyclic base matrix with the dimensions and the parity structure of
BG1 / BG2 (core 4 x (Kb + 4) with the double-diagonal parity block, degree-1 extension parities, the two punctured
high-degree columns).  Everything else in the chain — sizes, CRCs, segmentation, rate matching, interleaving,
scrambling, the decoder — is identical for both, so a stand-in BLER curve compares LLR sources (CUDA engine vs
oracle) through a real code of the right length and rate, but it is NOT comparable with ``results/*_results``
(different code, and the synthetic channel is TDL-like, not TDL-B/C).  ``TBEncoder.base_graph_source`` says which one
is in use.

Host-side harness code (NumPy for the encoder, torch tensor ops on CPU or GPU for the decoder); nothing on the
receiver hot path imports it.
"""
from __future__ import annotations

import functools
import math
import os
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np

from .config import NrxConfig, mcs_bits_per_symbol, mcs_code_rate
from .pusch import PuschGrid, gold_sequence

# ---------------------------------------------------------------------------------------------------------------
# CRC (TS 38.212 §5.1): generator polynomials without the leading term, MSB first
# ---------------------------------------------------------------------------------------------------------------
_CRC_POLY = {
    "24A": (24, 0x864CFB),   # x^24+x^23+x^18+x^17+x^14+x^11+x^10+x^7+x^6+x^5+x^4+x^3+x+1
    "24B": (24, 0x800063),   # x^24+x^23+x^6+x^5+x+1
    "24C": (24, 0xB2B117),   # x^24+x^23+x^21+x^20+x^17+x^15+x^13+x^12+x^8+x^4+x^2+x+1
    "16": (16, 0x1021),      # x^16+x^12+x^5+1
    "11": (11, 0x621),       # x^11+x^10+x^9+x^5+1
    "6": (6, 0x21),          # x^6+x^5+1
}


@functools.lru_cache(maxsize=64)
def _crc_matrix(kind: str, k: int) -> np.ndarray:
    """[k, L] uint8: row i = remainder of x^(k-1-i+L) modulo the generator, so that parity = bits @ M mod 2
    (the CRC with zero initial state is linear in the message)."""
    L, poly = _CRC_POLY[kind]
    rows = np.zeros((k, L), np.uint8)
    r = poly                                            # x^L mod g
    mask = (1 << L) - 1
    for i in range(k - 1, -1, -1):
        rows[i] = [(r >> (L - 1 - j)) & 1 for j in range(L)]
        r = ((r << 1) & mask) ^ (poly if (r >> (L - 1)) & 1 else 0)
    return rows


def crc_parity(bits: np.ndarray, kind: str) -> np.ndarray:
    """Parity bits [..., L] of the message bits [..., k] (first bit = highest power)."""
    bits = np.asarray(bits)
    m = _crc_matrix(kind, bits.shape[-1])
    return ((bits.astype(np.int64) @ m.astype(np.int64)) & 1).astype(np.uint8)


def crc_attach(bits: np.ndarray, kind: str) -> np.ndarray:
    return np.concatenate([np.asarray(bits, np.uint8), crc_parity(bits, kind)], axis=-1)


def crc_check(bits_with_crc: np.ndarray, kind: str) -> np.ndarray:
    """True where the trailing L bits are the CRC of the leading ones."""
    L = _CRC_POLY[kind][0]
    b = np.asarray(bits_with_crc, np.uint8)
    return np.all(crc_parity(b[..., :-L], kind) == b[..., -L:], axis=-1)


# ---------------------------------------------------------------------------------------------------------------
# Transport-block size (TS 38.214 §5.1.3.2, used for PUSCH through §6.1.4.2)
# ---------------------------------------------------------------------------------------------------------------
_TBS_TABLE = (   # Table 5.1.3.2-1 (N_info <= 3824)
    24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128, 136, 144, 152, 160, 168, 176, 184, 192, 208, 224, 240,
    256, 272, 288, 304, 320, 336, 352, 368, 384, 408, 432, 456, 480, 504, 528, 552, 576, 608, 640, 672, 704, 736, 768,
    808, 848, 888, 928, 984, 1032, 1064, 1128, 1160, 1192, 1224, 1256, 1288, 1320, 1352, 1416, 1480, 1544, 1608, 1672,
    1736, 1800, 1864, 1928, 2024, 2088, 2152, 2216, 2280, 2408, 2472, 2536, 2600, 2664, 2728, 2792, 2856, 2976, 3104,
    3240, 3368, 3496, 3624, 3752, 3824)
assert len(_TBS_TABLE) == 93


def tb_size(num_prb: int, num_res_per_prb: int, num_bits_per_symbol: int, target_coderate: float,
            num_layers: int = 1, tb_scaling: float = 1.0) -> int:
    """TBS for ``num_prb`` PRBs with ``num_res_per_prb`` = 12 N_symb - N_DMRS - N_oh data REs per PRB."""
    n_re = min(156, num_res_per_prb) * num_prb
    n_info = tb_scaling * n_re * target_coderate * num_bits_per_symbol * num_layers
    if n_info <= 3824:
        n = max(3, int(math.floor(math.log2(n_info))) - 6)
        n_q = max(24, (1 << n) * int(math.floor(n_info / (1 << n))))
        return next(t for t in _TBS_TABLE if t >= n_q)
    n = int(math.floor(math.log2(n_info - 24))) - 5
    n_q = max(3840, (1 << n) * int(round((n_info - 24) / (1 << n))))
    if target_coderate <= 0.25:
        c = -(-(n_q + 24) // 3816)
    elif n_q > 8424:
        c = -(-(n_q + 24) // 8424)
    else:
        c = 1
    return 8 * c * (-(-(n_q + 24) // (8 * c))) - 24


# ---------------------------------------------------------------------------------------------------------------
# LDPC base graphs
# ---------------------------------------------------------------------------------------------------------------
_LIFT_A = (2, 3, 5, 7, 9, 11, 13, 15)           # i_LS -> a;  Z = a 2^j <= 384  (TS 38.212 Table 5.3.2-1)
_BG_DIMS = {1: (46, 68, 22), 2: (42, 52, 10)}   # rows, columns, systematic columns


def lifting_sets() -> List[List[int]]:
    return [[a << j for j in range(8) if (a << j) <= 384] for a in _LIFT_A]


def select_lifting(kb: int, k_prime: int) -> Tuple[int, int]:
    """Smallest Z with kb Z >= K' and the index of its lifting set."""
    best = None
    for i_ls, zs in enumerate(lifting_sets()):
        for z in zs:
            if kb * z >= k_prime and (best is None or z < best[0]):
                best = (z, i_ls)
    if best is None:
        raise ValueError("code block too long for the largest lifting size")
    return best


class BaseGraphUnavailable(RuntimeError):
    pass


def _bg_csv_path(bg: int) -> Optional[str]:
    """Path of the base-graph file: $NRX_LDPC_BG_DIR first, then the package data of an installed Sionna (0.x:
    sionna/fec/ldpc/codes, 1.x: sionna/phy/fec/ldpc/codes) — located through the import machinery WITHOUT importing
    Sionna (that would pull TensorFlow into a process that owns a GPU)."""
    name = f"5G_bg{bg}.csv"
    d = os.environ.get("NRX_LDPC_BG_DIR")
    if d and os.path.exists(os.path.join(d, name)):
        return os.path.join(d, name)
    try:
        import importlib.util
        spec = importlib.util.find_spec("sionna")
    except (ImportError, ValueError):
        spec = None
    for root in (list(spec.submodule_search_locations or []) if spec is not None else []):
        for sub in (("fec", "ldpc", "codes"), ("phy", "fec", "ldpc", "codes")):
            p = os.path.join(root, *sub, name)
            if os.path.exists(p):
                return p
    return None


@functools.lru_cache(maxsize=16)
def load_base_graph(bg: int, i_ls: int) -> np.ndarray:
    """Base matrix [rows, cols] of shift values (-1 = no edge) of TS 38.212 Table 5.3.2-2 (bg = 1) / -3 (bg = 2) for
    lifting set ``i_ls``, read from a file in the layout Sionna ships (`;`-separated; two header lines; then one
    line per edge: row index (blank = same row as the line before); column index; the 8 shift values)."""
    p = _bg_csv_path(bg)
    if p is None:
        raise BaseGraphUnavailable(
            "the TS 38.212 base-graph tables are not in this image: install Sionna or point NRX_LDPC_BG_DIR at a "
            "directory with 5G_bg1.csv / 5G_bg2.csv (Sionna's layout); standin_base_graph() gives a structural twin")
    rows, cols, _ = _BG_DIMS[bg]
    bm = np.full((rows, cols), -1, np.int32)
    tab = np.genfromtxt(p, delimiter=";")
    r = 0
    for line in tab[2:]:
        if not np.isnan(line[0]):
            r = int(line[0])
        bm[r, int(line[1])] = int(line[2 + i_ls])
    return bm


@functools.lru_cache(maxsize=16)
def standin_base_graph(bg: int, i_ls: int, seed: int = 38212) -> np.ndarray:
    """Seeded quasi-cyclic base matrix with the shape and parity structure of BG1 / BG2 (NOT the 3GPP shift
    values): rows 0-3 connect most systematic columns and carry the double-diagonal parity block
    (column kb: rows 0, 1, 3 with shifts 1, 0, 1 — its circulants sum to the identity, which makes the block
    invertible — then an identity staircase), rows >= 4 are single-parity extension checks that always touch one of
    the two punctured columns 0 / 1, a few more systematic columns and sometimes a core parity column."""
    rows, cols, kb = _BG_DIMS[bg]
    rng = np.random.default_rng([seed, bg, i_ls])
    bm = np.full((rows, cols), -1, np.int32)
    core_w = 19 if bg == 1 else 7
    for r in range(4):
        keep = {0, 1} if r != 2 or bg == 2 else {0}
        keep |= set(rng.choice(np.arange(2, kb), size=core_w - len(keep), replace=False).tolist())
        for c in sorted(keep):
            bm[r, c] = int(rng.integers(0, 384))
    bm[0, kb], bm[1, kb], bm[3, kb] = 1, 0, 1
    for r in range(4):
        if r < 3:
            bm[r, kb + 1 + r] = 0
        if r > 0:
            bm[r, kb + r] = 0
    for r in range(4, rows):
        deg = int(rng.integers(3, 8)) if bg == 1 else int(rng.integers(2, 5))
        c_sys = {int(rng.integers(0, 2))} | set(rng.choice(np.arange(2, kb), size=deg - 1, replace=False).tolist())
        for c in sorted(c_sys):
            bm[r, c] = int(rng.integers(0, 384))
        if rng.random() < 0.5:
            bm[r, kb + int(rng.integers(0, 4))] = int(rng.integers(0, 384))
        bm[r, kb + r] = 0
    return bm


def base_graph(bg: int, i_ls: int, source: str = "auto") -> Tuple[np.ndarray, str]:
    """(base matrix, "3gpp" | "standin")."""
    if source not in ("auto", "3gpp", "standin"):
        raise ValueError("base_graph source must be 'auto', '3gpp' or 'standin'")
    if source != "standin":
        try:
            return load_base_graph(bg, i_ls), "3gpp"
        except BaseGraphUnavailable:
            if source == "3gpp":
                raise
    return standin_base_graph(bg, i_ls), "standin"


def _edges(bm: np.ndarray, z: int) -> Tuple[np.ndarray, np.ndarray]:
    """Check / variable index of every edge of the lifted graph: block (r, c) with shift s is the identity shifted
    right by s mod z, i.e. check r z + i  <->  variable c z + (i + s) mod z."""
    r, c = np.nonzero(bm >= 0)
    s = bm[r, c] % z
    i = np.arange(z)
    cn = (r[:, None] * z + i[None, :]).reshape(-1)
    vn = (c[:, None] * z + (i[None, :] + s[:, None]) % z).reshape(-1)
    return cn.astype(np.int64), vn.astype(np.int64)


def _gf2_inverse(a: np.ndarray) -> np.ndarray:
    """Inverse of a square 0/1 matrix over GF(2) (Gauss-Jordan on bit-packed rows)."""
    n = a.shape[0]
    m = np.packbits(np.concatenate([a.astype(np.uint8), np.eye(n, dtype=np.uint8)], axis=1), axis=1)
    for col in range(n):
        byte, bit = col >> 3, 7 - (col & 7)
        has = (m[:, byte] >> bit) & 1
        piv = col + int(np.argmax(has[col:]))
        if not has[piv]:
            raise ValueError("core parity block of the base graph is singular")
        if piv != col:
            m[[col, piv]] = m[[piv, col]]
            has[[col, piv]] = has[[piv, col]]
        has[col] = 0
        m[has.astype(bool)] ^= m[col]
    return np.unpackbits(m, axis=1)[:, n:2 * n]


class LdpcCode:
    """One lifted 5G-style LDPC code: systematic encoder and the edge lists the decoder walks."""

    def __init__(self, bg: int, z: int, i_ls: int, source: str = "auto"):
        self.bg, self.z, self.i_ls = bg, z, i_ls
        self.bm, self.source = base_graph(bg, i_ls, source)
        self.rows, self.cols, self.kb = _BG_DIMS[bg]
        self.k = self.kb * z                          # systematic bits incl. fillers
        self.n = (self.cols - 2) * z                  # transmitted-buffer length (first two columns punctured)
        self.cn, self.vn = _edges(self.bm, z)
        kb, Z = self.kb, z
        core = np.zeros((4 * Z, 4 * Z), np.uint8)
        for r in range(4):
            for c in range(4):
                s = self.bm[r, kb + c]
                if s >= 0:
                    i = np.arange(Z)
                    core[r * Z + i, c * Z + (i + s % Z) % Z] ^= 1
        self._core_inv = _gf2_inverse(core).astype(np.float32)

    def _block_mul(self, x: np.ndarray, rows: range, cols: range) -> np.ndarray:
        """sum_c H[r, c] x_c for the row blocks `rows` over the column blocks `cols`; x [B, cols, Z] -> [B, rows, Z]."""
        Z = self.z
        out = np.zeros((x.shape[0], len(rows), Z), np.uint8)
        for ri, r in enumerate(rows):
            for ci, c in enumerate(cols):
                s = self.bm[r, c]
                if s >= 0:
                    out[:, ri] ^= np.roll(x[:, ci], -(int(s) % Z), axis=-1)
        return out

    def encode(self, c: np.ndarray) -> np.ndarray:
        """c [B, k] systematic bits (fillers as zeros) -> full codeword [B, cols z] with H w = 0."""
        B, Z, kb = c.shape[0], self.z, self.kb
        s = np.asarray(c, np.uint8).reshape(B, kb, Z)
        lam = self._block_mul(s, range(4), range(kb)).reshape(B, 4 * Z)             # A s
        p1 = (np.rint(lam.astype(np.float32) @ self._core_inv.T).astype(np.int64) & 1).astype(np.uint8)
        p1b = p1.reshape(B, 4, Z)
        p2 = self._block_mul(s, range(4, self.rows), range(kb)) ^ \
            self._block_mul(p1b, range(4, self.rows), range(kb, kb + 4))
        return np.concatenate([s.reshape(B, -1), p1, p2.reshape(B, -1)], axis=1)

    def syndrome_ok(self, w: np.ndarray) -> np.ndarray:
        acc = np.zeros((w.shape[0], self.rows * self.z), np.int64)
        np.add.at(acc, (slice(None), self.cn), w[:, self.vn].astype(np.int64))
        return np.all((acc & 1) == 0, axis=1)


@functools.lru_cache(maxsize=8)
def _ldpc_code(bg: int, z: int, i_ls: int, source: str) -> LdpcCode:
    return LdpcCode(bg, z, i_ls, source)


# ---------------------------------------------------------------------------------------------------------------
# Transport-block encoder / decoder (the interface of Sionna's TBEncoder / TBDecoder)
# ---------------------------------------------------------------------------------------------------------------
class TBEncoder:
    """TB bits [..., tb_size] -> scrambled coded bits [..., num_coded_bits]  (TS 38.212 §6.2, one codeword, rv 0).

    Arguments follow Sionna's ``TBEncoder`` as the reference builds it inside ``PUSCHTransmitter``
    (``utils/parameters.py:186-246``): ``target_tb_size``, ``num_coded_bits``, ``target_coderate``,
    ``num_bits_per_symbol``, ``num_layers``, ``n_rnti`` / ``n_id`` (an int, or one value per transmitter for inputs
    shaped [..., num_tx, tb_size]), ``use_scrambler``.  ``base_graph``: "auto" (3GPP tables if available, else the
    stand-in), "3gpp", "standin"."""

    def __init__(self, target_tb_size: int, num_coded_bits: int, target_coderate: float, num_bits_per_symbol: int,
                 num_layers: int = 1, n_rnti: Union[int, Sequence[int]] = 1, n_id: Union[int, Sequence[int]] = 1,
                 use_scrambler: bool = True, base_graph: str = "auto"):
        A, G, Qm = int(target_tb_size), int(num_coded_bits), int(num_bits_per_symbol)
        if G % (Qm * num_layers):
            raise ValueError("num_coded_bits must be a multiple of num_bits_per_symbol x num_layers")
        if A >= G:
            raise ValueError("target_tb_size must be smaller than num_coded_bits")
        self.tb_size, self.num_coded_bits, self.coderate = A, G, float(target_coderate)
        self.num_bits_per_symbol, self.num_layers = Qm, int(num_layers)
        self.tb_crc = "24A" if A > 3824 else "16"
        B = A + _CRC_POLY[self.tb_crc][0]
        R = self.coderate
        self.bg = 2 if (A <= 292 or (A <= 3824 and R <= 0.67) or R <= 0.25) else 1
        k_cb = 8448 if self.bg == 1 else 3840
        if B <= k_cb:
            self.num_cbs, b_prime = 1, B
        else:
            self.num_cbs = -(-B // (k_cb - 24))
            b_prime = B + 24 * self.num_cbs
        if b_prime % self.num_cbs:
            raise ValueError("TB size does not split into equal code blocks (not a TS 38.214 TBS)")
        self.k_prime = b_prime // self.num_cbs                      # bits per code block incl. CRCs
        kb = 22 if self.bg == 1 else (10 if B > 640 else 9 if B > 560 else 8 if B > 192 else 6)
        self.z, self.i_ls = select_lifting(kb, self.k_prime)
        self.code = _ldpc_code(self.bg, self.z, self.i_ls, base_graph)
        self.base_graph_source = self.code.source
        self.k = self.code.k
        self.n_cb = self.code.n                                     # circular buffer (I_LBRM = 0)
        # rate matching: E_r per code block (§5.4.2.1)
        C, q = self.num_cbs, Qm * self.num_layers
        self.cb_e = [q * (G // (q * C)) if j <= C - (G // q) % C - 1 else q * -(-G // (q * C)) for j in range(C)]
        assert sum(self.cb_e) == G
        # positions of the circular buffer read by bit selection, fillers skipped (k0 = 0 for rv 0)
        Z = self.z
        fill_lo, fill_hi = self.k_prime - 2 * Z, self.k - 2 * Z
        valid = np.array([j for j in range(self.n_cb) if not (fill_lo <= j < fill_hi)], np.int64)
        self.cb_sel = [valid[np.arange(E) % valid.size] for E in self.cb_e]
        # bit interleaver (§5.4.2.2): f[i + j Qm] = e[i E/Qm + j]
        self.cb_perm = []
        for E in self.cb_e:
            j, i = np.meshgrid(np.arange(E // Qm), np.arange(Qm), indexing="ij")
            self.cb_perm.append((i * (E // Qm) + j).reshape(-1))
        self.use_scrambler = bool(use_scrambler)
        rn = np.atleast_1d(np.asarray(n_rnti, np.int64))
        ni = np.atleast_1d(np.asarray(n_id, np.int64))
        if rn.shape != ni.shape:
            raise ValueError("n_rnti and n_id need the same length")
        self.scramble_seq = np.stack([gold_sequence(int(r) * (1 << 15) + int(i), G) for r, i in zip(rn, ni)]).astype(np.uint8)

    def _seq(self, shape) -> np.ndarray:
        s = self.scramble_seq
        if s.shape[0] == 1:
            return s[0]
        if len(shape) < 2 or shape[-2] != s.shape[0]:
            raise ValueError("inputs must be shaped [..., num_tx, n] when n_rnti / n_id are given per transmitter")
        return s

    def code_blocks(self, bits: np.ndarray) -> np.ndarray:
        """[N, tb_size] -> [N, C, k]: TB CRC, segmentation, per-block CRC, fillers as zeros."""
        tb = crc_attach(bits, self.tb_crc)
        N, C = tb.shape[0], self.num_cbs
        if C == 1:
            cb = tb[:, None, :]
        else:
            cb = crc_attach(tb.reshape(N, C, -1), "24B")
        out = np.zeros((N, C, self.k), np.uint8)
        out[:, :, :self.k_prime] = cb
        return out

    def __call__(self, bits: np.ndarray) -> np.ndarray:
        bits = np.asarray(bits, np.uint8)
        if bits.shape[-1] != self.tb_size:
            raise ValueError(f"last dimension must be tb_size = {self.tb_size}")
        lead = bits.shape[:-1]
        cb = self.code_blocks(bits.reshape(-1, self.tb_size))
        N, C = cb.shape[:2]
        w = self.code.encode(cb.reshape(N * C, self.k)).reshape(N, C, -1)
        d = w[:, :, 2 * self.z:]
        out = np.concatenate([d[:, r, self.cb_sel[r]][:, self.cb_perm[r]] for r in range(C)], axis=1)
        out = out.reshape(lead + (self.num_coded_bits,))
        if self.use_scrambler:
            out = out ^ self._seq(out.shape)
        return out


class TBDecoder:
    """LLRs [..., num_coded_bits] (``llr > 0 <=> bit 1``, the receiver's convention, utils/neural_rx.py:864) ->
    (b_hat [..., tb_size] uint8, tb_crc_status [...] bool)  — Sionna's ``TBDecoder(encoder, num_bp_iter, cn_type)``
    as used at utils/neural_rx.py:1407-1413 / :1600.  ``cn_type``: "boxplus" (exact, the reference's setting) or
    "minsum".  Runs on the device of the input when it is a torch tensor, else on the CPU."""

    LLR_MAX = 20.0

    def __init__(self, encoder: TBEncoder, num_bp_iter: int = 20, cn_type: str = "boxplus"):
        if cn_type not in ("boxplus", "minsum"):
            raise ValueError("cn_type must be 'boxplus' or 'minsum'")
        self.encoder, self.num_bp_iter, self.cn_type = encoder, int(num_bp_iter), cn_type
        e = encoder
        Z = e.z
        # prune the extension checks whose degree-1 parity was never sent (their messages are exactly zero)
        last = max(int(s.max()) for s in e.cb_sel) + 2 * Z                 # highest codeword position with an LLR
        n_cols = max(e.code.kb + 4, -(-(last + 1) // Z))
        n_rows = n_cols - e.code.kb
        keep = e.code.cn < n_rows * Z
        self._cn = e.code.cn[keep]
        self._vn = e.code.vn[keep]
        self._n_cn, self._n_vn = n_rows * Z, n_cols * Z
        self._cache = {}

    def _idx(self, device):
        import torch
        if device not in self._cache:
            self._cache[device] = (torch.as_tensor(self._cn, device=device), torch.as_tensor(self._vn, device=device))
        return self._cache[device]

    def _bp(self, lch):
        """Flooding BP on channel LLRs lch [N, n_vn] in the log(p0/p1) convention; returns the a-posteriori LLRs."""
        import torch
        cn, vn = self._idx(lch.device)
        N = lch.shape[0]
        c2v = torch.zeros((N, cn.numel()), dtype=lch.dtype, device=lch.device)
        tot = lch
        for _ in range(self.num_bp_iter):
            v2c = (tot[:, vn] - c2v).clamp_(-self.LLR_MAX, self.LLR_MAX)
            neg = (v2c < 0)
            par = torch.zeros((N, self._n_cn), dtype=lch.dtype, device=lch.device).index_add_(1, cn, neg.to(lch.dtype))
            sign = 1.0 - 2.0 * torch.remainder(par[:, cn] + neg.to(lch.dtype), 2.0)       # parity of the OTHER edges
            mag = v2c.abs()
            if self.cn_type == "boxplus":
                phi = -torch.log(torch.tanh(mag.clamp(min=1e-7) * 0.5).clamp(min=1e-30))
                s = torch.zeros((N, self._n_cn), dtype=lch.dtype, device=lch.device).index_add_(1, cn, phi)
                ext = (s[:, cn] - phi).clamp_(min=1e-7)
                out = -torch.log(torch.tanh(ext * 0.5).clamp(min=1e-30))
            else:
                big = torch.full((N, self._n_cn), float("inf"), dtype=lch.dtype, device=lch.device)
                m1 = big.scatter_reduce(1, cn.expand(N, -1), mag, "amin")
                is1 = mag <= m1[:, cn]
                m2 = big.scatter_reduce(1, cn.expand(N, -1), torch.where(is1, torch.full_like(mag, float("inf")), mag), "amin")
                n1 = torch.zeros((N, self._n_cn), dtype=lch.dtype, device=lch.device).index_add_(1, cn, is1.to(lch.dtype))
                m2 = torch.where(n1 > 1, m1, m2)                                           # ties: the minimum stays
                out = torch.where(is1, m2[:, cn], m1[:, cn])
            c2v = sign * out.clamp_(max=self.LLR_MAX)
            tot = lch + torch.zeros_like(lch).index_add_(1, vn, c2v)
        return tot

    def __call__(self, llr):
        import torch
        e = self.encoder
        is_np = isinstance(llr, np.ndarray)
        x = torch.as_tensor(np.ascontiguousarray(llr, dtype=np.float32)) if is_np else llr.to(torch.float32)
        if x.shape[-1] != e.num_coded_bits:
            raise ValueError(f"last dimension must be num_coded_bits = {e.num_coded_bits}")
        lead = tuple(x.shape[:-1])
        dev = x.device
        if e.use_scrambler:
            seq = torch.as_tensor(e._seq(x.shape).astype(np.float32), device=dev)
            x = x * (1.0 - 2.0 * seq)                                   # descramble: flip where c = 1
        x = -x.reshape(-1, e.num_coded_bits)                            # -> log(p0/p1)
        N, C, Z = x.shape[0], e.num_cbs, e.z
        lch = torch.zeros((N, C, self._n_vn), dtype=torch.float32, device=dev)
        off = 0
        for r in range(C):
            E = e.cb_e[r]
            f = x[:, off:off + E]
            off += E
            inv = torch.as_tensor(np.argsort(e.cb_perm[r]), device=dev)
            sel = torch.as_tensor(e.cb_sel[r] + 2 * Z, device=dev)
            lch[:, r].index_add_(1, sel, f[:, inv])                     # de-interleave, combine repetitions
        lch[:, :, e.k_prime:e.k] = self.LLR_MAX                         # fillers are known zeros
        lch.clamp_(-self.LLR_MAX, self.LLR_MAX)
        post = self._bp(lch.reshape(N * C, self._n_vn))
        hard = (post[:, :e.k_prime] < 0).to(torch.uint8).cpu().numpy().reshape(N, C, e.k_prime)
        tb = hard[:, :, :-24].reshape(N, -1) if C > 1 else hard[:, 0]     # strip the code-block CRCs
        ok = crc_check(tb, e.tb_crc)                                        # status = the TB CRC, like Sionna's
        b_hat = tb[:, :e.tb_size].reshape(lead + (e.tb_size,))
        ok = ok.reshape(lead)
        if is_np:
            return b_hat, ok
        return torch.as_tensor(b_hat, device=dev), torch.as_tensor(ok, device=dev)


# ---------------------------------------------------------------------------------------------------------------
# PUSCH parameters of a receiver configuration
# ---------------------------------------------------------------------------------------------------------------
def pusch_tb_encoder(cfg: NrxConfig, grid: PuschGrid, mcs_list_idx: int = 0, n_rnti: Union[int, Sequence[int]] = 1,
                     n_id: Union[int, Sequence[int]] = 1, base_graph: str = "auto") -> TBEncoder:
    """The TBEncoder of ``sys_parameters.transmitters[mcs_list_idx]`` (utils/parameters.py:186-246): one layer per
    transmitter, TBS from the MCS and the slot's data REs, G = data REs x bits per symbol.  ``n_rntis`` / ``n_ids`` are
    [1, 1] in every shipped cfg (config/nrx_large.cfg:41-42)."""
    mcs = cfg.mcs_index[mcs_list_idx]
    qm = mcs_bits_per_symbol(mcs, cfg.mcs_table)
    rate = mcs_code_rate(mcs, cfg.mcs_table)
    n_prb = grid.num_subcarriers // 12
    if grid.num_data_res % n_prb:
        raise ValueError("data REs are not uniform over the PRBs")
    tbs = tb_size(n_prb, grid.num_data_res // n_prb, qm, rate)
    return TBEncoder(tbs, grid.num_data_res * qm, rate, qm, 1, n_rnti, n_id, True, base_graph)
