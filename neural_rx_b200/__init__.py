"""nrx-b200: B200-native (sm_100a) engine for the neural PUSCH receiver's full-slot hot path.

Host code is Python (PyTorch tensors as device buffers); all arithmetic runs in hand-written CUDA
kernels behind the C ABI of ``include/nrx_b200.h`` (``libnrx_b200.so``), bound with ctypes.
There is no CPU or library fallback: constructing an engine without the shared object raises.
"""
from .config import NrxConfig, get_config, load_cfg, PRESETS  # noqa: F401
from .weights import NrxWeights, load_weights, save_weights, random_weights  # noqa: F401
from .pusch import PuschGrid, build_grid  # noqa: F401

__all__ = ["NrxConfig", "get_config", "load_cfg", "PRESETS", "NrxWeights", "load_weights",
           "save_weights", "random_weights", "PuschGrid", "build_grid"]
