"""Batch-1 latency of one forward: CUDA-event total next to the sum of the per-kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from neural_rx_b200.weights import load_weights, random_weights

for label in sys.argv[1:] or ["nrx_rt", "nrx_large"]:
    cfg = get_config(label)
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "weights", f"{label}_weights")
    w = load_weights(cfg, p) if os.path.exists(p) else random_weights(cfg)
    grid = build_grid(cfg)
    sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
    y = torch.as_tensor(sb.y).cuda(); act = torch.ones((1, 2), device="cuda")
    eng = NrxEngine(cfg, w, grid)
    eng.set_fused(int(os.environ.get('NRX_FUSED', '1')))
    for _ in range(10):
        eng.forward(y, act, want=("llr", "h_hat_refined"))
    torch.cuda.synchronize()
    ts = []
    for _ in range(100):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); eng.forward(y, act, want=("llr", "h_hat_refined")); b.record(); b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    eng.set_profiling(True)
    for _ in range(20):
        eng.forward(y, act, want=("llr", "h_hat_refined"))
    prof = eng.get_profile()
    eng.set_profiling(False)
    tot = sum(v["ms"] for v in prof.values()) / 20 * 1e3
    print(f"{label}: p50 {np.median(ts):.1f} us, sum of kernel times {tot:.1f} us, launches {eng.launches_per_forward(1)}")
    for k, v in prof.items():
        if v["launches"]:
            print(f"   {k:24s} {v['ms'] / v['launches'] * 1e3:8.1f} us x {v['launches'] // 20}")
