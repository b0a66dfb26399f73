"""Batch-1 latency with the forward captured in a CUDA graph (torch.cuda.graph around the C-ABI call)."""
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
    eng.set_fused(2)
    outs = {}
    want = ("llr", "h_hat_refined")
    for _ in range(5):
        eng.forward(y, act, want=want, out=outs)
    torch.cuda.synchronize()
    ref = outs["llr"].clone()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        eng.forward(y, act, want=want, out=outs)
        with torch.cuda.graph(g, stream=s):
            eng.forward(y, act, want=want, out=outs)
    torch.cuda.synchronize()
    outs["llr"].zero_()
    g.replay(); torch.cuda.synchronize()
    same = bool(torch.equal(outs["llr"], ref))
    def p50(fn, n=200):
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
        ts = np.sort(ts); return ts[len(ts) // 2], ts[int(0.99 * len(ts))]
    e = p50(lambda: eng.forward(y, act, want=want, out=outs))
    r = p50(g.replay)
    print(f"{label}: eager p50 {e[0]:.1f} us p99 {e[1]:.1f} | graph p50 {r[0]:.1f} us p99 {r[1]:.1f} | graph result identical: {same}")
