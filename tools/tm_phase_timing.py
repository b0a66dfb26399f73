"""Cycle accounting of the plan-4 stack kernel (debug build with -DNRX_PHASE_TIMING): per role of CTA 0,
the cycles spent waiting for the input, draining the accumulator, in the depthwise pass, and the
issue -> complete latency of each layer's GEMM.   python tools/tm_phase_timing.py [label] [batch]"""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from neural_rx_b200 import build as B
from neural_rx_b200 import engine as E

lib_path = os.path.join(ROOT, "neural_rx_b200", "libnrx_b200_timing.so")
if not os.path.exists(lib_path) or "--rebuild" in sys.argv:
    fine = ["-DNRX_TM_FINE"] if "--fine" in sys.argv else []
    subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DNRX_PHASE_TIMING"] + fine + ["-o", lib_path] + B.SOURCES, check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
label = args[0] if args else "nrx_large"
batch = int(args[1]) if len(args) > 1 else 30
E._LIB_PATH = lib_path
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from tests.common import get_weights

cfg = get_config(label); w, _ = get_weights(cfg); grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.repeat(sb.y, batch, axis=0)).cuda()
act = torch.ones((batch, 2), device="cuda")
eng = E.NrxEngine(cfg, w, grid); eng.set_fused(4)
lib = E.load_library()
buf = (ctypes.c_ulonglong * 32)()
for _ in range(2): eng.forward(y, act, want=("llr",))
torch.cuda.synchronize(); lib.nrx_debug_tm_cycles(buf)
n_fwd = 3
for _ in range(n_fwd): eng.forward(y, act, want=("llr",))
torch.cuda.synchronize(); lib.nrx_debug_tm_cycles(buf)
steps = max(buf[21], 1)
print(f"{label} batch {batch}: CTA 0, {steps} steps over {n_fwd} forwards, {buf[20] / steps:.0f} cycles per step (kernel total / steps)")
for r, name in enumerate(("layer 1 (role 0)", "layer 2 (role 1)", "layer 3 (role 2)")):
    print(f"  {name}: input waits {buf[3*r]/steps:7.0f}  drain {buf[3*r+1]/steps:7.0f}  pass {buf[3*r+2]/steps:7.0f}  epilogue {buf[15+r]/steps:6.0f}"
          f"  | GEMM issue->complete {buf[9+r]/steps:6.0f}  wait before issue {buf[12+r]/steps:6.0f}")
if buf[22] + buf[23] + buf[24] + buf[25]:
    print("  layer 2 pass, per step: accumulator load + wait %.0f  convert %.0f  shuffles + depthwise %.0f  A store %.0f"
          % tuple(buf[22 + i] / steps for i in range(4)))
