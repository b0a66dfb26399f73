"""Per-role cycle budget of the pipelined aggregation kernel (debug build with -DNRX_PHASE_TIMING).

    python tools/agg_timing.py [label] [batch]

Builds neural_rx_b200/libnrx_b200_timing.so, runs a few forwards and prints where lane 0 of epilogue warp 0, the
tensor thread and the producer thread of CTA 0 spend their cycles (waits vs work)."""
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
eng.set_fused(1)
lib = E.load_library()
buf = (ctypes.c_ulonglong * 48)()
for _ in range(2):
    eng.forward(y, act, want=("llr",))
torch.cuda.synchronize()
lib.nrx_debug_ws_cycles(buf)
n_fwd = 3
for _ in range(n_fwd):
    eng.forward(y, act, want=("llr",))
torch.cuda.synchronize()
lib.nrx_debug_ws_cycles(buf)
names = {
    "E": ["loop", "E1: wait acc1", "E1: tmem ld", "E1: math + sts", "E1: fence + arrive", "E2: set-up", "E2: wait acc2",
          "E2: tmem ld + arrive", "E2: math + sts", "E2: copy-out"],
    "T": ["loop", "wait stage full", "issue gemm 1", "wait hidden", "wait acc2 free", "issue gemm 2"],
    "P": ["loop", "wait stage empty"],
}
launches = n_fwd * cfg.num_nrx_iter
for r, (role, nm) in enumerate(names.items()):
    vals = [buf[16 * r + i] for i in range(16)]
    tot = sum(vals)
    print(f"{role}: total {tot} cycles over {launches} launches = {tot / launches / 1e3:.1f} k cycles per launch")
    for i, n in enumerate(nm + ["-"] * (15 - len(nm)) + ["item switch / drain"]):
        if vals[i]:
            print(f"   {n:28s} {vals[i]:12d}  {100.0 * vals[i] / max(tot, 1):5.1f}%")
