"""Per-launch CUDA-event times of every kernel class for nrx_large, 30 slots (with and without the LS estimate output)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.getcwd())
from neural_rx_b200 import engine as E
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from tests.common import get_weights
cfg = get_config("nrx_large"); w,_ = get_weights(cfg); grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.repeat(sb.y, 30, axis=0)).cuda(); act = torch.ones((30, 2), device="cuda")
eng = E.NrxEngine(cfg, w, grid)
for _ in range(3): eng.forward(y, act, want=("llr",))
torch.cuda.synchronize(); eng.set_profiling(True); eng.get_profile()
for plan in (1, 5):
    eng.set_fused(plan)
    for balanced in (1, 0):
        eng.debug_option(eng.OPT_STACK_BALANCED, balanced)
        for _ in range(3): eng.forward(y, act, want=("llr", "h_hat_refined"))
        torch.cuda.synchronize(); eng.get_profile()
        for _ in range(10): eng.forward(y, act, want=("llr", "h_hat_refined"))
        torch.cuda.synchronize(); pr = eng.get_profile()
        print("plan", plan, "balanced" if balanced else "uniform ", {k: round(v["ms"] / max(v["launches"], 1) * 1e3, 1) for k, v in pr.items() if v["launches"]})
