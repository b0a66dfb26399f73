"""Per-kernel-class timing of one plan (nrx_large, batch 30): python tools/tm_time.py <plan>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from tests.common import get_weights
plan = int(sys.argv[1]); label = sys.argv[2] if len(sys.argv) > 2 else "nrx_large"; batch = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cfg = get_config(label); w, _ = get_weights(cfg); grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=2, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.tile(sb.y, (batch // 2 + 1, 1, 1, 1, 1))[:batch]).cuda()
act = torch.ones((batch, 2), device="cuda")
eng = NrxEngine(cfg, w, grid); eng.set_fused(plan)
for _ in range(3): eng.forward(y, act, want=("llr", "h_hat_refined"))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): eng.forward(y, act, want=("llr", "h_hat_refined"))
e1.record(); torch.cuda.synchronize()
print("lib", os.environ.get("NRX_B200_LIB", "default"), "plan", plan, label, batch, "ms/forward %.4f" % (e0.elapsed_time(e1) / 20))
