"""Per-phase cycle budget of the fused stack kernel (debug build with -DNRX_PHASE_TIMING).

    python tools/phase_timing.py [label] [batch]

Builds neural_rx_b200/libnrx_b200_timing.so, runs a few forwards and prints, per step of CTA 0,
the cycles between the phase boundaries of nrx_stack_kernel (update stack launches dominate)."""
import ctypes, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from neural_rx_b200 import build as B
from neural_rx_b200 import engine as E

lib_path = os.path.join(ROOT, "neural_rx_b200", "libnrx_b200_timing.so")
if not os.path.exists(lib_path) or "--rebuild" in sys.argv:
    subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-DNRX_PHASE_TIMING", "-o", lib_path] + B.SOURCES, check=True)
if "--build-only" in sys.argv:
    sys.exit(0)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
label = args[0] if args else "nrx_large"
batch = int(args[1]) if len(args) > 1 else 30
E._LIB_PATH = lib_path
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.weights import load_weights, random_weights

cfg = get_config(label)
p = os.path.join(ROOT, "weights", f"{label}_weights")
w = load_weights(cfg, p) if os.path.exists(p) else random_weights(cfg)
grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.repeat(sb.y, batch, axis=0)).cuda()
act = torch.ones((batch, 2), device="cuda")
eng = E.NrxEngine(cfg, w, grid)
eng.set_fused(int(os.environ.get("NRX_FUSED", "1")))
if os.environ.get("NRX_NUM_IT"):
    eng.num_it = int(os.environ["NRX_NUM_IT"])   # e.g. 1: StateInit weighs as much as UpdateState in the totals
lib = E.load_library()
buf = (ctypes.c_ulonglong * 32)()
for _ in range(2):
    eng.forward(y, act, want=("llr",))
torch.cuda.synchronize()
lib.nrx_debug_phase_cycles(buf)
n_fwd = 3
for _ in range(n_fwd):
    eng.forward(y, act, want=("llr",))
torch.cuda.synchronize()
lib.nrx_debug_phase_cycles(buf)
names = ["z wait+sync", "L1 dw", "L1 sync", "L1 issue+prefetch", "L1 mma wait", "L1 epilogue", "L1 sync",
         "L2 dw", "L2 sync", "L2 issue+carry", "L2 mma wait", "L2 epilogue", "L2 sync",
         "L3 dw", "L3 sync", "L3 issue+carry+prefetch", "L3 mma wait", "L3 epilogue(stage)", "L3 sync",
         "copy-out", "copy-out sync", "item/loop overhead", "wait for helper warps (ws kernel)"]
tot = sum(buf[i] for i in range(23))
print(f"{label} batch {batch}: total cycles CTA0 over {n_fwd} forwards = {tot}")
for i, n in enumerate(names):
    print(f"  {n:26s} {buf[i]:12d}  {100.0 * buf[i] / max(tot, 1):5.1f}%")
