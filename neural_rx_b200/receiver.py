"""Receiver-facing API: the reference's object interface for the neural-receiver hot path, backed
by the CUDA engine.

Mirrors (names, argument meaning, error behaviour):

* ``NeuralPUSCHReceiver(sys_parameters, training=False)`` — ``utils/neural_rx.py:1384-1603``;
  inference call ``receiver((y, active_tx), no, mcs_arr_eval=[0], mcs_ue_mask_eval=None)``
  (``:1544-1603``; ``no`` never influences the LLRs, it only fed the discarded LS error variance).
  The reference then runs Sionna's ``TBDecoder`` on the LLRs (``:1600``).  The transport-block chain is
  third-party code outside the hot path: ``neural_rx_b200/tb.py`` restates it for the BLER
  harness, and it needs the TS 38.212 base-graph tables, which this image does not hold.  Default
  (``tb_decoding="auto"``): when the tables are found (Sionna installed, or ``$NRX_LDPC_BG_DIR``) the call returns
  ``(b_hat, h_hat_refined, h_hat, tb_crc_status)`` like the reference; otherwise the first element is the LLR
  tensor ``[B, U, n_coded_bits]`` the decoder would consume and the CRC status is ``None``.
  ``tb_decoding="standin"`` decodes with the structural stand-in code (harness / tests), ``"off"`` never decodes.
* ``CGNNOFDM`` LLR-level call — ``utils/neural_rx.py:813-881``:
  ``receiver.llrs((y, active_tx), mcs_arr_eval, mcs_ue_mask_eval)`` -> ``(llr, h_hat_refined)``.
* ``num_it`` property with the reference's assertion (``:532-542``).
* ``load_weights(receiver, path)`` — ``utils/utils.py:53-70``.

``sys_parameters`` is an :class:`~neural_rx_b200.config.NrxConfig` (the subset of the reference's
``Parameters`` object that the hot path reads), a preset name or a cfg file path.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple, Union

import numpy as np

from .config import NrxConfig, get_config
from .engine import NrxEngine
from .pusch import PuschGrid, build_grid
from .weights import NrxWeights, load_weights as _load_weight_file, random_weights


class NeuralPUSCHReceiver:
    def __init__(self, sys_parameters: Union[NrxConfig, str], training: bool = False,
                 weights: Optional[NrxWeights] = None, grid: Optional[PuschGrid] = None,
                 device: int = 0, tb_decoding: str = "auto", num_bp_iter: int = 20, cn_type: str = "boxplus",
                 n_rntis: Sequence[int] = (1, 1), n_ids: Sequence[int] = (1, 1), **kwargs):
        if training:
            raise NotImplementedError("the B200 engine implements the inference path only")
        cfg = get_config(sys_parameters) if isinstance(sys_parameters, str) else sys_parameters
        cfg.validate()
        self._sys_parameters = cfg
        self._grid = build_grid(cfg) if grid is None else grid
        self._device = device
        self._num_it = cfg.num_nrx_iter_eval
        self._engine: Optional[NrxEngine] = None
        # transport-block decoders, one per supported MCS (utils/neural_rx.py:1402-1413)
        if tb_decoding not in ("auto", "3gpp", "standin", "off"):
            raise ValueError("tb_decoding must be 'auto', '3gpp', 'standin' or 'off'")
        self._tb_encoders, self._tb_decoders = [], []
        if tb_decoding != "off":
            from . import tb as _tb
            U = self._grid.num_tx
            rn = (list(n_rntis) * U)[:U]
            ni = (list(n_ids) * U)[:U]
            try:
                for i in range(cfg.num_mcss_supported):
                    enc = _tb.pusch_tb_encoder(cfg, self._grid, i, rn, ni, "3gpp" if tb_decoding == "auto" else tb_decoding)
                    self._tb_encoders.append(enc)
                    self._tb_decoders.append(_tb.TBDecoder(enc, num_bp_iter=num_bp_iter, cn_type=cn_type))
            except (_tb.BaseGraphUnavailable, ValueError):
                # "auto": no tables in this installation (or a grid the TB chain has no size for): LLRs are returned
                if tb_decoding != "auto":
                    raise
                self._tb_encoders, self._tb_decoders = [], []
        # like Keras, the layers exist (randomly initialised) until load_weights() is called
        self.set_weights(random_weights(cfg) if weights is None else weights)

    # ---- weights ----------------------------------------------------------------------------
    def set_weights(self, weights: NrxWeights) -> None:
        if self._engine is not None:
            self._engine.close()
        self._weights = weights
        self._engine = NrxEngine(self._sys_parameters, weights, self._grid, self._device)
        self._engine.num_it = self._num_it

    def get_weights(self):
        return self._weights.to_list()

    # ---- properties -----------------------------------------------------------------------------
    @property
    def engine(self) -> NrxEngine:
        return self._engine

    @property
    def grid(self) -> PuschGrid:
        return self._grid

    @property
    def num_it(self) -> int:
        return self._num_it

    @num_it.setter
    def num_it(self, val: int) -> None:
        # utils/neural_rx.py:537-542
        assert 1 <= val <= self._sys_parameters.num_nrx_iter, "Invalid number of iterations"
        self._num_it = int(val)
        self._engine.num_it = self._num_it

    # ---- calls ------------------------------------------------------------------------------------
    def _indices(self, batch: int, mcs_arr_eval: Sequence[int], mcs_ue_mask_eval, per_user_heads: bool):
        cfg = self._sys_parameters
        U = self._grid.num_tx
        head = int(mcs_arr_eval[0])
        if not 0 <= head < cfg.num_mcss_supported:
            raise ValueError("mcs_arr_eval index out of range")
        io_index = head_index = None
        if mcs_ue_mask_eval is not None:
            m = np.asarray(mcs_ue_mask_eval, dtype=np.float32).reshape(-1, U, cfg.num_mcss_supported)
            m = np.broadcast_to(m, (batch, U, cfg.num_mcss_supported))
            if not np.all((m == 0) | (m == 1)) or not np.all(m.sum(-1) == 1):
                raise ValueError("mcs_ue_mask_eval must be one-hot over the supported MCSs")
            sel = np.argmax(m, axis=-1).astype(np.int32)
            if not cfg.mcs_var_mcs_masking:
                io_index = sel                       # one StateInit stack per MCS (:562-569)
            if per_user_heads:
                head_index = np.zeros_like(sel) if cfg.mcs_var_mcs_masking else sel
        llr_head = 0 if cfg.mcs_var_mcs_masking else head
        out_bits = None
        if not per_user_heads:
            out_bits = cfg.num_bits_per_symbol[head]   # masking mode: slice [..., :bits] (:586-588)
        return io_index, head_index, llr_head, out_bits

    def llrs(self, inputs: Tuple, mcs_arr_eval: Sequence[int] = (0,), mcs_ue_mask_eval=None,
             per_user_heads: bool = False, want: Sequence[str] = ("llr", "h_hat_refined", "h_hat")):
        """LLR-level call (``CGNNOFDM.forward`` inference branch).  ``inputs = (y, active_tx)``,
        NumPy arrays on the host or torch CUDA tensors.  Returns a dict with the requested
        outputs (``llr`` [B,U,n_coded_bits] of head ``mcs_arr_eval[0]`` for every user — the
        reference's behaviour — or, with ``per_user_heads=True``, each user's own head padded to
        the widest constellation)."""
        y, active_tx = inputs
        B = int(y.shape[0])
        io_index, head_index, llr_head, out_bits = self._indices(B, mcs_arr_eval, mcs_ue_mask_eval, per_user_heads)
        if isinstance(y, np.ndarray):
            return self._engine.forward_host(y, np.asarray(active_tx), io_index, head_index, llr_head, out_bits, want)
        import torch
        dev = y.device
        t = lambda a: None if a is None else torch.as_tensor(a, device=dev)
        out = self._engine.forward(y, torch.as_tensor(active_tx, device=dev), t(io_index), t(head_index),
                                   llr_head, out_bits, want)
        return {k: v for k, v in out.items() if not k.startswith("_")}

    def __call__(self, inputs: Tuple, no=None, mcs_arr_eval: Sequence[int] = (0,), mcs_ue_mask_eval=None):
        """``NeuralPUSCHReceiver.forward`` inference branch: ``(b_hat, h_hat_refined, h_hat, tb_crc_status)`` with
        transport-block decoding, ``(llr, h_hat_refined, h_hat, None)`` without — see the module doc-string."""
        out = self.llrs(inputs, mcs_arr_eval, mcs_ue_mask_eval)
        if not self._tb_decoders:
            return out["llr"], out["h_hat_refined"], out["h_hat"], None
        b_hat, tb_crc_status = self._tb_decoders[int(mcs_arr_eval[0])](out["llr"])     # :1600
        return b_hat, out["h_hat_refined"], out["h_hat"], tb_crc_status

    @property
    def tb_encoders(self):
        """One ``tb.TBEncoder`` per supported MCS (empty without the base-graph tables)."""
        return self._tb_encoders

    forward = __call__


class NeuralReceiverONNX:
    """Aerial / TensorRT-shaped wrapper — ``NeuralReceiverONNX`` of ``utils/neural_rx.py:1716-1812``
    (the model ``scripts/export_onnx.py`` exports; binding names ``:153-160``).

    ``forward([rx_slot_real, rx_slot_imag, h_hat_real, h_hat_imag, active_dmrs_ports,
    dmrs_ofdm_pos, dmrs_subcarrier_pos])`` -> ``(llr [B,bits,U,F,T], h_hat [B,U,F,T,2N_rx])`` with
    the Aerial LLR sign (minus the Sionna convention, ``:1809-1810``).  Inputs are NumPy arrays
    (copied to the device and back) or torch CUDA tensors (outputs stay on the device).  Built
    from the receiver configuration instead of the reference's positional layer-size arguments;
    single-MCS configurations only, as in the reference (``:1796``)."""

    def __init__(self, sys_parameters: Union[NrxConfig, str], weights: Optional[NrxWeights] = None,
                 n_size_bwp: Optional[int] = None, device: int = 0):
        cfg = get_config(sys_parameters) if isinstance(sys_parameters, str) else sys_parameters
        cfg.validate()
        if cfg.num_mcss_supported != 1:
            raise NotImplementedError("NeuralReceiverONNX has no support for mixed MCS")
        self._sys_parameters = cfg
        self._grid = build_grid(cfg, n_size_bwp=n_size_bwp)
        self._weights = random_weights(cfg) if weights is None else weights
        self._engine = NrxEngine(cfg, self._weights, self._grid, device)
        self._num_it = cfg.num_nrx_iter_eval
        self._engine.num_it = self._num_it

    @property
    def engine(self) -> NrxEngine:
        return self._engine

    @property
    def num_it(self) -> int:
        return self._num_it

    @num_it.setter
    def num_it(self, val: int) -> None:
        assert 1 <= val <= self._sys_parameters.num_nrx_iter, "Invalid number of iterations"   # :1766-1771
        self._num_it = int(val)
        self._engine.num_it = self._num_it

    def forward(self, inputs):
        import torch

        y_re, y_im, h_re, h_im, port_mask, ofdm_pos, sc_pos = inputs
        on_host = isinstance(y_re, np.ndarray)
        dev = torch.device("cuda", self._engine.device)
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(dev) if on_host else a
        to_np = lambda a: a.cpu().numpy() if hasattr(a, "cpu") else np.asarray(a)
        llr, h = self._engine.forward_aerial(t(y_re), t(y_im), t(h_re), t(h_im), t(port_mask), to_np(ofdm_pos),
                                             to_np(sc_pos))
        if on_host:
            torch.cuda.synchronize(dev)
            return llr.cpu().numpy(), h.cpu().numpy()
        return llr, h

    __call__ = forward


def load_weights(receiver: NeuralPUSCHReceiver, model_path: str) -> None:
    """``utils.load_weights(model, path)``: unpickle the Keras weight list and set it."""
    receiver.set_weights(_load_weight_file(receiver._sys_parameters, model_path))
