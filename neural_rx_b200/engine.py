"""ctypes binding of ``libnrx_b200.so`` (C ABI in ``include/nrx_b200.h``).

PyTorch only provides device buffers and the CUDA stream; every arithmetic step of the receiver
runs inside the library's sm_100a kernels.  A missing library is a hard error (no fallback).
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence

import numpy as np

from .config import NrxConfig
from .pusch import PuschGrid
from .weights import NrxWeights

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libnrx_b200.so")

NRX_MAX_IO = 4
NRX_MAX_DMRS = 4

#: every symbol declared in include/nrx_b200.h
EXPORTED_SYMBOLS = (
    "nrx_create", "nrx_destroy", "nrx_set_num_it", "nrx_get_num_it", "nrx_set_slots_per_pass", "nrx_set_fused", "nrx_set_host_chunk",
    "nrx_set_skip_inactive",
    "nrx_workspace_bytes", "nrx_forward", "nrx_forward_host", "nrx_forward_host_async", "nrx_wait", "nrx_set_aerial_dmrs", "nrx_forward_aerial", "nrx_launches_per_forward",
    "nrx_plan_stack_chunks", "nrx_plan_stack_range",
    "nrx_mac_per_pixel", "nrx_set_profiling", "nrx_get_profile", "nrx_last_error", "nrx_version",
    "nrx_debug_aggregate", "nrx_debug_stack", "nrx_debug_readout", "nrx_debug_option",
)

KERNEL_CLASSES = ("power", "prep", "sep_32x128", "sep_128x128", "sep_128x64_init_out",
                  "sep_128x64_update_out", "agg", "readout", "stack_init", "stack_update")


class NrxError(RuntimeError):
    """Non-zero status from the C ABI (message from ``nrx_last_error``)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"libnrx_b200 error {code}: {msg}")
        self.code = code


class ModelDesc(ctypes.Structure):
    """``nrx_model_desc`` of include/nrx_b200.h."""

    _fields_ = [
        ("num_rx_ant", ctypes.c_int32), ("max_num_tx", ctypes.c_int32),
        ("num_subcarriers", ctypes.c_int32), ("num_ofdm_symbols", ctypes.c_int32),
        ("d_s", ctypes.c_int32), ("num_it", ctypes.c_int32),
        ("units_init", ctypes.c_int32 * 2), ("units_agg", ctypes.c_int32),
        ("units_state", ctypes.c_int32 * 2), ("units_readout", ctypes.c_int32),
        ("n_io", ctypes.c_int32), ("io_bits", ctypes.c_int32 * NRX_MAX_IO),
        ("num_dmrs_symbols", ctypes.c_int32), ("dmrs_symbols", ctypes.c_int32 * NRX_MAX_DMRS),
        ("focc_block", ctypes.c_int32), ("num_data_res", ctypes.c_int32),
    ]


_lib = None


def load_library(path: Optional[str] = None) -> ctypes.CDLL:
    """Load the shared object and declare the prototypes.  Raises if it has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("NRX_B200_LIB") or _LIB_PATH      # env override: experiment builds only
    if not os.path.exists(p):
        raise RuntimeError(
            f"{p} not found: build the CUDA library first (python -m neural_rx_b200.build). "
            "The receiver has no CPU fallback.")
    lib = ctypes.CDLL(p)
    c_void_pp = ctypes.POINTER(ctypes.c_void_p)
    f32p, i32p = ctypes.c_void_p, ctypes.c_void_p
    lib.nrx_create.argtypes = [ctypes.POINTER(ModelDesc), ctypes.POINTER(ctypes.c_void_p),
                               ctypes.POINTER(ctypes.c_int64), ctypes.c_int32, f32p, i32p, f32p, i32p,
                               ctypes.c_int32, c_void_pp]
    lib.nrx_destroy.argtypes = [ctypes.c_void_p]
    lib.nrx_set_num_it.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_get_num_it.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32)]
    lib.nrx_set_slots_per_pass.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_set_fused.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_set_skip_inactive.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_set_host_chunk.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_workspace_bytes.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_size_t)]
    lib.nrx_forward.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, f32p,
                                i32p, i32p, ctypes.c_int32, ctypes.c_int32, f32p, f32p, f32p, f32p,
                                ctypes.c_void_p, ctypes.c_size_t]
    lib.nrx_forward_host.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, f32p, i32p, i32p,
                                     ctypes.c_int32, ctypes.c_int32, f32p, f32p, f32p, f32p]
    lib.nrx_forward_host_async.argtypes = lib.nrx_forward_host.argtypes + [ctypes.POINTER(ctypes.c_int64)]
    lib.nrx_wait.argtypes = [ctypes.c_void_p, ctypes.c_int64]
    lib.nrx_set_aerial_dmrs.argtypes = [ctypes.c_void_p, i32p, ctypes.c_int32, i32p, ctypes.c_int32]
    lib.nrx_forward_aerial.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32, f32p, f32p, f32p, f32p, f32p,
                                       f32p, f32p, ctypes.c_void_p, ctypes.c_size_t]
    lib.nrx_launches_per_forward.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int32)]
    i32o = ctypes.POINTER(ctypes.c_int32)
    lib.nrx_plan_stack_chunks.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, i32o]
    i64o = ctypes.POINTER(ctypes.c_int64)
    lib.nrx_plan_stack_range.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, i64o, i64o]
    lib.nrx_mac_per_pixel.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]
    lib.nrx_set_profiling.argtypes = [ctypes.c_void_p, ctypes.c_int32]
    lib.nrx_get_profile.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]
    vp = ctypes.c_void_p
    lib.nrx_debug_aggregate.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, vp, vp, vp]
    lib.nrx_debug_stack.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, vp, vp]
    lib.nrx_debug_readout.argtypes = [vp, vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, vp, vp, vp]
    lib.nrx_debug_option.argtypes = [vp, ctypes.c_int32, ctypes.c_int32]
    lib.nrx_last_error.restype = ctypes.c_char_p
    lib.nrx_version.restype = ctypes.c_char_p
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("nrx_last_error", "nrx_version"):
            fn.restype = ctypes.c_int
    if path is None:
        _lib = lib
    return lib


def make_desc(cfg: NrxConfig, grid: PuschGrid) -> ModelDesc:
    """Fill ``nrx_model_desc`` from the receiver configuration (what CGNN.__init__ reads,
    utils/neural_rx.py:407-530, 638-662) and the PUSCH grid."""
    cfg.validate()
    if len(cfg.num_units_init) != 2 or any(len(x) != 2 for x in cfg.num_units_state) \
            or any(len(x) != 1 for x in cfg.num_units_agg) or len(cfg.num_units_readout) != 1:
        raise NotImplementedError("the B200 engine implements 2 hidden sep-conv layers per stack and "
                                  "1 hidden layer in the aggregation / read-out MLPs")
    if len({tuple(x) for x in cfg.num_units_state}) != 1 or len({tuple(x) for x in cfg.num_units_agg}) != 1:
        raise NotImplementedError("all iterations must share the same layer widths")
    d = ModelDesc()
    d.num_rx_ant = cfg.num_rx_antennas
    d.max_num_tx = grid.num_tx
    d.num_subcarriers = grid.num_subcarriers
    d.num_ofdm_symbols = grid.num_ofdm_symbols
    d.d_s = cfg.d_s
    d.num_it = cfg.num_nrx_iter
    d.units_init[0], d.units_init[1] = cfg.num_units_init
    d.units_agg = cfg.num_units_agg[0][0]
    d.units_state[0], d.units_state[1] = cfg.num_units_state[0]
    d.units_readout = cfg.num_units_readout[0]
    bits = cfg.readout_bits
    d.n_io = len(bits)
    for i, b in enumerate(bits):
        d.io_bits[i] = b
    d.num_dmrs_symbols = len(grid.dmrs_symbols)
    for i, s in enumerate(grid.dmrs_symbols):
        d.dmrs_symbols[i] = s
    d.focc_block = grid.focc_block
    d.num_data_res = grid.num_data_res
    return d


class NrxEngine:
    """One engine = one (architecture, weight file, PUSCH grid) on one GPU."""

    def __init__(self, cfg: NrxConfig, weights: NrxWeights, grid: PuschGrid, device: int = 0):
        self._lib = load_library()
        self.cfg, self.grid, self.device = cfg, grid, int(device)
        self.desc = make_desc(cfg, grid)
        arrays = [np.ascontiguousarray(a, dtype=np.float32) for a in weights.to_list()]
        ptrs = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
        sizes = (ctypes.c_int64 * len(arrays))(*[a.size for a in arrays])
        pil = np.ascontiguousarray(grid.pilots.astype(np.complex64)).view(np.float32)
        nn = np.ascontiguousarray(grid.nn_index, dtype=np.int32)
        pe = np.ascontiguousarray(grid.pos_enc, dtype=np.float32)
        di = np.ascontiguousarray(grid.data_index, dtype=np.int32)
        handle = ctypes.c_void_p()
        self._check(self._lib.nrx_create(ctypes.byref(self.desc), ptrs, sizes, len(arrays), pil.ctypes.data,
                                         nn.ctypes.data, pe.ctypes.data, di.ctypes.data, self.device,
                                         ctypes.byref(handle)))
        self._h = handle
        self._ws = None

    # ---- plumbing ---------------------------------------------------------------------------
    def _check(self, rc: int) -> None:
        if rc != 0:
            raise NrxError(rc, self._lib.nrx_last_error().decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.nrx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_it(self) -> int:
        v = ctypes.c_int32()
        self._check(self._lib.nrx_get_num_it(self._h, ctypes.byref(v)))
        return v.value

    @num_it.setter
    def num_it(self, val: int) -> None:
        self._check(self._lib.nrx_set_num_it(self._h, int(val)))

    def set_slots_per_pass(self, slots: int) -> None:
        self._check(self._lib.nrx_set_slots_per_pass(self._h, int(slots)))
        self._ws = None

    def set_fused(self, fused) -> None:
        """6 (default): serial StateInit kernel + pipelined UpdateState kernels; 1/True: serial fused stack kernels +
        aggregation kernel; 2: fused stacks with the message MLP in their tail (two users); 5: warp-specialised
        pipelined kernels for both stacks; 0/False: one kernel per SeparableConv2D layer (3, 4: round-1 experiments,
        only in -DNRX_EXPERIMENTAL_PLANS builds)."""
        self._check(self._lib.nrx_set_fused(self._h, int(fused)))

    def set_skip_inactive(self, enable: bool) -> None:
        """Do not compute the planes of inactive users (their outputs become zeros; active users bit-identical)."""
        self._check(self._lib.nrx_set_skip_inactive(self._h, int(bool(enable))))

    def set_host_chunk(self, slots: int) -> None:
        """Slots per pipeline chunk of the host-buffer call (0 = default)."""
        self._check(self._lib.nrx_set_host_chunk(self._h, int(slots)))

    def workspace_bytes(self, batch: int) -> int:
        v = ctypes.c_size_t()
        self._check(self._lib.nrx_workspace_bytes(self._h, int(batch), ctypes.byref(v)))
        return v.value

    def launches_per_forward(self, batch: int) -> int:
        v = ctypes.c_int32()
        self._check(self._lib.nrx_launches_per_forward(self._h, int(batch), ctypes.byref(v)))
        return v.value

    def mac_per_pixel(self, llr_head: int = 0) -> int:
        v = ctypes.c_int64()
        self._check(self._lib.nrx_mac_per_pixel(self._h, int(llr_head), ctypes.byref(v)))
        return v.value

    def set_profiling(self, enable: bool) -> None:
        self._check(self._lib.nrx_set_profiling(self._h, int(bool(enable))))

    def get_profile(self) -> Dict[str, Dict[str, float]]:
        """Per-kernel-class CUDA-event time (ms) and launch count since the last call."""
        ms = (ctypes.c_double * len(KERNEL_CLASSES))()
        n = (ctypes.c_int64 * len(KERNEL_CLASSES))()
        self._check(self._lib.nrx_get_profile(self._h, ms, n))
        return {k: {"ms": ms[i], "launches": int(n[i])} for i, k in enumerate(KERNEL_CLASSES)}

    def flops_per_slot(self, llr_head: int = 0) -> float:
        """Algorithmic FLOPs of one slot: 2 * U * F * T * MAC_per_pixel."""
        g = self.grid
        return 2.0 * g.num_tx * g.num_subcarriers * g.num_ofdm_symbols * self.mac_per_pixel(llr_head)

    def _out_bits(self, llr_head: int, head_index, out_bits: Optional[int]) -> int:
        if out_bits is not None:
            return int(out_bits)
        bits = self.cfg.readout_bits
        return max(bits) if head_index is not None else bits[llr_head]

    # ---- device call (torch tensors) ------------------------------------------------------------
    def forward(self, y, active_tx, io_index=None, head_index=None, llr_head: int = 0,
                out_bits: Optional[int] = None, want: Sequence[str] = ("llr", "h_hat_refined", "h_hat"),
                out: Optional[Dict] = None, stream=None, workspace=None) -> Dict:
        """``y`` complex64 CUDA tensor [B,1,N_rx,T,F]; ``active_tx`` float32 CUDA [B,U];
        ``io_index`` / ``head_index`` int32 CUDA [B,U] or None.  Enqueues on the current stream
        and returns CUDA tensors (keys of ``want`` among llr, llr_grid, h_hat_refined, h_hat).

        Scratch memory: ``workspace`` (a uint8 CUDA tensor of at least ``workspace_bytes(B)`` bytes) or, by
        default, one engine-owned tensor that grows on demand and is shared by all calls — so calls of one engine
        must be issued on ONE stream at a time (two forwards in flight on different streams would race on it;
        pass distinct ``workspace`` tensors for that)."""
        import torch

        g, N = self.grid, self.cfg.num_rx_antennas
        B = int(y.shape[0])
        if tuple(y.shape) != (B, 1, N, g.num_ofdm_symbols, g.num_subcarriers) or y.dtype != torch.complex64:
            raise ValueError(f"y must be complex64 [B,1,{N},{g.num_ofdm_symbols},{g.num_subcarriers}], "
                             f"got {tuple(y.shape)} {y.dtype}")
        if not y.is_cuda or y.device.index != self.device:
            raise ValueError("y must live on the engine's CUDA device")
        U = g.num_tx
        y = y.contiguous()
        active_tx = active_tx.to(dtype=torch.float32).contiguous()
        if tuple(active_tx.shape) != (B, U):
            raise ValueError(f"active_tx must be [B,{U}]")
        dev = y.device
        bits = self._out_bits(llr_head, head_index, out_bits)
        res = {} if out is None else out
        per = g.num_subcarriers * g.num_ofdm_symbols
        shapes = {"llr": (B, U, g.num_data_res * bits), "llr_grid": (B, U, g.num_subcarriers, g.num_ofdm_symbols, bits),
                  "h_hat_refined": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N),
                  "h_hat": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N)}
        for k in want:
            if k not in res:
                res[k] = torch.empty(shapes[k], dtype=torch.float32, device=dev)
        need = self.workspace_bytes(B)
        if workspace is not None:
            if workspace.dtype != torch.uint8 or not workspace.is_cuda or workspace.numel() < need:
                raise ValueError(f"workspace must be a uint8 CUDA tensor of at least {need} bytes")
            ws = workspace
        else:
            if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
                self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
            ws = self._ws
        ptr = lambda k: res[k].data_ptr() if k in want else None
        iptr = lambda t: None if t is None else t.to(dtype=torch.int32).contiguous()
        io_t, head_t = iptr(io_index), iptr(head_index)
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        self._check(self._lib.nrx_forward(
            self._h, ctypes.c_void_p(st), B, y.data_ptr(), active_tx.data_ptr(),
            None if io_t is None else io_t.data_ptr(), None if head_t is None else head_t.data_ptr(),
            int(llr_head), bits, ptr("llr"), ptr("llr_grid"), ptr("h_hat_refined"), ptr("h_hat"),
            ws.data_ptr(), ws.numel()))
        res["_keepalive"] = (y, active_tx, io_t, head_t, ws)
        del per
        return res

    # ---- CUDA-graph replay of a fixed-shape forward (latency mode) ---------------------------------
    def capture(self, y, active_tx, io_index=None, head_index=None, llr_head: int = 0,
                out_bits: Optional[int] = None, want: Sequence[str] = ("llr", "h_hat_refined", "h_hat")):
        """Capture one forward on the given (static) input tensors into a CUDA graph.  nrx_forward
        only enqueues kernels (no allocation, no synchronisation), so it can be recorded as is.
        Returns ``(graph, outputs)``: overwrite ``y`` / ``active_tx`` in place, call ``graph.replay()``
        and read ``outputs`` (same tensors every time).

        The graph bakes device addresses in (inputs, outputs, scratch, tensor maps encoded from the scratch
        address), so it gets a workspace tensor of its own that lives as long as the graph object does
        (``graph._nrx_keepalive``) — later eager calls that grow or drop the engine's shared workspace, or
        ``set_slots_per_pass``, cannot pull memory from under a replay."""
        import torch

        outs: Dict = {}
        ws = torch.empty(self.workspace_bytes(int(y.shape[0])), dtype=torch.uint8, device=y.device)
        kw = dict(io_index=io_index, head_index=head_index, llr_head=llr_head, out_bits=out_bits, want=want, out=outs,
                  workspace=ws)
        side = torch.cuda.Stream(device=y.device)
        side.wait_stream(torch.cuda.current_stream(y.device))
        with torch.cuda.stream(side):
            self.forward(y, active_tx, **kw)              # warm-up: allocates outputs and the workspace
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                self.forward(y, active_tx, **kw)
        torch.cuda.current_stream(y.device).wait_stream(side)
        graph._nrx_keepalive = (ws, outs.get("_keepalive"), dict(outs))
        return graph, {k: v for k, v in outs.items() if not k.startswith("_")}

    # ---- Aerial / TensorRT-shaped call (NeuralReceiverONNX.forward, utils/neural_rx.py:1773-1812) ----
    def set_aerial_dmrs(self, dmrs_ofdm_pos, dmrs_subcarrier_pos) -> None:
        op = np.ascontiguousarray(dmrs_ofdm_pos, dtype=np.int32)
        sp = np.ascontiguousarray(dmrs_subcarrier_pos, dtype=np.int32)
        U = self.grid.num_tx
        if op.ndim != 2 or sp.ndim != 2 or op.shape[0] != U or sp.shape[0] != U:
            raise ValueError(f"dmrs_ofdm_pos / dmrs_subcarrier_pos must be [{U}, n]")
        self._check(self._lib.nrx_set_aerial_dmrs(self._h, op.ctypes.data, op.shape[1], sp.ctypes.data, sp.shape[1]))
        self._aerial_key = (op.tobytes(), sp.tobytes())
        self._aerial_pilots = op.shape[1] * (self.grid.num_subcarriers // 12) * sp.shape[1]

    def forward_aerial(self, rx_slot_real, rx_slot_imag, h_hat_real, h_hat_imag, active_dmrs_ports,
                       dmrs_ofdm_pos, dmrs_subcarrier_pos, stream=None):
        """float32 CUDA tensors rx_slot_* [B,F,T,N_rx], h_hat_* [B,n_pilots,U,N_rx],
        active_dmrs_ports [B,U]; integer host arrays dmrs_ofdm_pos [U,n_sym], dmrs_subcarrier_pos
        [U,n_sc].  Returns (llr [B,bits,U,F,T] = -LLR, h_hat [B,U,F,T,2N_rx]) CUDA tensors."""
        import torch

        op = np.ascontiguousarray(np.asarray(dmrs_ofdm_pos), dtype=np.int32)
        sp = np.ascontiguousarray(np.asarray(dmrs_subcarrier_pos), dtype=np.int32)
        if getattr(self, "_aerial_key", None) != (op.tobytes(), sp.tobytes()):
            self.set_aerial_dmrs(op, sp)
        g, N = self.grid, self.cfg.num_rx_antennas
        B, U, F, T = int(rx_slot_real.shape[0]), g.num_tx, g.num_subcarriers, g.num_ofdm_symbols
        ts = [rx_slot_real, rx_slot_imag, h_hat_real, h_hat_imag, active_dmrs_ports]
        shapes = [(B, F, T, N), (B, F, T, N), (B, self._aerial_pilots, U, N), (B, self._aerial_pilots, U, N), (B, U)]
        for i, (t, sh) in enumerate(zip(ts, shapes)):
            if tuple(t.shape) != sh:
                raise ValueError(f"input {i} must have shape {sh}, got {tuple(t.shape)}")
            if not t.is_cuda or t.device.index != self.device:
                raise ValueError("inputs must live on the engine's CUDA device")
            ts[i] = t.to(dtype=torch.float32).contiguous()
        dev = ts[0].device
        bits = self.cfg.readout_bits[0]
        llr = torch.empty((B, bits, U, F, T), dtype=torch.float32, device=dev)
        h = torch.empty((B, U, F, T, 2 * N), dtype=torch.float32, device=dev)
        need = self.workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need or self._ws.device != dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream(dev).cuda_stream if stream is None else stream
        self._check(self._lib.nrx_forward_aerial(self._h, ctypes.c_void_p(st), B, ts[0].data_ptr(), ts[1].data_ptr(),
                                                 ts[2].data_ptr(), ts[3].data_ptr(), ts[4].data_ptr(), llr.data_ptr(),
                                                 h.data_ptr(), self._ws.data_ptr(), self._ws.numel()))
        self._aerial_keepalive = ts
        return llr, h

    # ---- test hooks: one kernel on tensors in the internal activation layout ------------------------
    def _stream(self, t):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)

    OPT_AGG_PIPELINED = 1
    OPT_STACK_BALANCED = 2

    def debug_option(self, option: int, value: int):
        """Test switch (include/nrx_b200.h, nrx_debug_option)."""
        self._check(self._lib.nrx_debug_option(self._h, int(option), int(value)))

    def debug_aggregate(self, it: int, s, active_tx):
        """s: fp16 CUDA [B,U,F,T,64] (state rows) -> a fp16 [B,U,F,T,64]."""
        import torch
        s = s.contiguous()
        act = active_tx.to(dtype=torch.float32).contiguous()
        a = torch.empty_like(s)
        self._check(self._lib.nrx_debug_aggregate(self._h, self._stream(s), int(it), int(s.shape[0]), s.data_ptr(),
                                                  act.data_ptr(), a.data_ptr()))
        return a

    def debug_stack(self, it: int, batch: int, z0=None, a=None, s=None, stack: int = 0):
        """it >= 0: UpdateState_it on (a, s) fp16 [B,U,F,T,64]; it < 0: StateInit on z0 fp16 [B,U,F,T,32]."""
        import torch
        src = z0 if it < 0 else s
        out = torch.empty(tuple(src.shape[:-1]) + (64,), dtype=torch.float16, device=src.device)
        ptr = lambda t: None if t is None else t.contiguous().data_ptr()
        self._check(self._lib.nrx_debug_stack(self._h, self._stream(src), int(it), int(stack), int(batch), ptr(z0), ptr(a), ptr(s),
                                              out.data_ptr()))
        return out

    def debug_readout(self, head: int, s, out_bits: int):
        import torch
        s = s.contiguous()
        B, U, F, T = (int(x) for x in s.shape[:4])
        llr = torch.empty((B, U, F, T, out_bits), dtype=torch.float32, device=s.device)
        h = torch.empty((B, U, F, T, 2 * self.cfg.num_rx_antennas), dtype=torch.float32, device=s.device)
        self._check(self._lib.nrx_debug_readout(self._h, self._stream(s), int(head), B, int(out_bits), s.data_ptr(),
                                                llr.data_ptr(), h.data_ptr()))
        return llr, h

    # ---- host call (NumPy arrays, H2D + D2H inside) --------------------------------------------
    def forward_host(self, y: np.ndarray, active_tx: np.ndarray, io_index=None, head_index=None,
                     llr_head: int = 0, out_bits: Optional[int] = None,
                     want: Sequence[str] = ("llr", "h_hat_refined", "h_hat"),
                     out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """NumPy in, NumPy out.  ``out`` may hold preallocated C-contiguous float32 result arrays
        (e.g. page-locked ones from :func:`pinned_empty`, which are then DMA'd in place)."""
        g, N = self.grid, self.cfg.num_rx_antennas
        y = np.ascontiguousarray(y, dtype=np.complex64)
        B, U = y.shape[0], g.num_tx
        if y.shape != (B, 1, N, g.num_ofdm_symbols, g.num_subcarriers):
            raise ValueError(f"y must be [B,1,{N},{g.num_ofdm_symbols},{g.num_subcarriers}], got {y.shape}")
        act = np.ascontiguousarray(active_tx, dtype=np.float32)
        if act.shape != (B, U):
            raise ValueError(f"active_tx must be [B,{U}]")
        bits = self._out_bits(llr_head, head_index, out_bits)
        shapes = {"llr": (B, U, g.num_data_res * bits), "llr_grid": (B, U, g.num_subcarriers, g.num_ofdm_symbols, bits),
                  "h_hat_refined": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N),
                  "h_hat": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N)}
        res = {}
        for k in want:
            a = None if out is None else out.get(k)
            if a is not None and (a.shape != shapes[k] or a.dtype != np.float32 or not a.flags.c_contiguous):
                raise ValueError(f"out[{k!r}] must be a C-contiguous float32 array of shape {shapes[k]}")
            res[k] = np.empty(shapes[k], np.float32) if a is None else a
        io = None if io_index is None else np.ascontiguousarray(io_index, dtype=np.int32).reshape(B, U)
        hd = None if head_index is None else np.ascontiguousarray(head_index, dtype=np.int32).reshape(B, U)
        ptr = lambda k: res[k].ctypes.data if k in res else None
        self._check(self._lib.nrx_forward_host(
            self._h, B, y.ctypes.data, act.ctypes.data, None if io is None else io.ctypes.data,
            None if hd is None else hd.ctypes.data, int(llr_head), bits, ptr("llr"), ptr("llr_grid"),
            ptr("h_hat_refined"), ptr("h_hat")))
        return res


def _host_shapes(eng, B, bits):
    g, N, U = eng.grid, eng.cfg.num_rx_antennas, eng.grid.num_tx
    return {"llr": (B, U, g.num_data_res * bits), "llr_grid": (B, U, g.num_subcarriers, g.num_ofdm_symbols, bits),
            "h_hat_refined": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N),
            "h_hat": (B, U, g.num_subcarriers, g.num_ofdm_symbols, 2 * N)}


def forward_host_async(self, y: np.ndarray, active_tx: np.ndarray, out: Dict[str, np.ndarray], io_index=None,
                       head_index=None, llr_head: int = 0, out_bits: Optional[int] = None) -> int:
    """Asynchronous host-buffer call (``nrx_forward_host_async``): every array — ``y`` complex64, ``active_tx``
    float32, the optional int32 index arrays and the float32 result arrays in ``out`` (keys among llr, llr_grid,
    h_hat_refined, h_hat) — must be page-locked (:func:`pinned_empty`) and C-contiguous with the exact dtype; nothing
    is copied or converted on the host.  Returns a ticket; the arrays belong to the engine until ``wait(ticket)``."""
    g, N = self.grid, self.cfg.num_rx_antennas
    B, U = y.shape[0], g.num_tx
    if y.dtype != np.complex64 or not y.flags.c_contiguous or y.shape != (B, 1, N, g.num_ofdm_symbols, g.num_subcarriers):
        raise ValueError(f"y must be C-contiguous complex64 [B,1,{N},{g.num_ofdm_symbols},{g.num_subcarriers}]")
    if active_tx.dtype != np.float32 or not active_tx.flags.c_contiguous or active_tx.shape != (B, U):
        raise ValueError(f"active_tx must be C-contiguous float32 [B,{U}]")
    for name, a in (("io_index", io_index), ("head_index", head_index)):
        if a is not None and (a.dtype != np.int32 or not a.flags.c_contiguous or a.shape != (B, U)):
            raise ValueError(f"{name} must be C-contiguous int32 [B,{U}]")
    bits = self._out_bits(llr_head, head_index, out_bits)
    shapes = _host_shapes(self, B, bits)
    for k, a in out.items():
        if k not in shapes or a.shape != shapes[k] or a.dtype != np.float32 or not a.flags.c_contiguous:
            raise ValueError(f"out[{k!r}] must be a C-contiguous float32 array of shape {shapes.get(k)}")
    ptr = lambda k: out[k].ctypes.data if k in out else None
    ticket = ctypes.c_int64()
    self._check(self._lib.nrx_forward_host_async(
        self._h, B, y.ctypes.data, active_tx.ctypes.data, None if io_index is None else io_index.ctypes.data,
        None if head_index is None else head_index.ctypes.data, int(llr_head), bits, ptr("llr"), ptr("llr_grid"),
        ptr("h_hat_refined"), ptr("h_hat"), ctypes.byref(ticket)))
    if not hasattr(self, "_inflight"):
        self._inflight = {}
    self._inflight[ticket.value] = (y, active_tx, io_index, head_index, out)      # keep the buffers alive
    return ticket.value


def wait(self, ticket: int) -> Dict[str, np.ndarray]:
    """Block until the asynchronous call ``ticket`` has delivered its outputs; returns its ``out`` dict."""
    self._check(self._lib.nrx_wait(self._h, int(ticket)))
    held = getattr(self, "_inflight", {})
    for t in [t for t in held if t < ticket]:
        held.pop(t)                                                             # calls complete in order
    return held.pop(ticket, (None,) * 5)[4]


NrxEngine.forward_host_async = forward_host_async
NrxEngine.wait = wait


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """Page-locked host array (NumPy view of a pinned torch tensor): ``forward_host`` copies such
    buffers by DMA without staging."""
    import torch
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.complex64): torch.complex64,
           np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
    return torch.empty(tuple(shape), dtype=tdt, pin_memory=True).numpy()
