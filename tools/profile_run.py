"""Short single-GPU workload for ncu: a few forwards of one config (default nrx_large, batch 4)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from neural_rx_b200.weights import random_weights, load_weights

label = sys.argv[1] if len(sys.argv) > 1 else "nrx_large"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = get_config(label)
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", f"{label}_weights")
w = load_weights(cfg, p) if os.path.exists(p) else random_weights(cfg)
grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.repeat(sb.y, batch, axis=0)).cuda()
act = torch.ones((batch, 2), device="cuda")
eng = NrxEngine(cfg, w, grid)
if "NRX_FUSED" in os.environ:
    eng.set_fused(int(os.environ["NRX_FUSED"]))        # default: the engine's default plan (6)
for _ in range(iters):
    out = eng.forward(y, act, want=("llr", "h_hat_refined"))
torch.cuda.synchronize()
print("ok", float(out["llr"].abs().mean()))
