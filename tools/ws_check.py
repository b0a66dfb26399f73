"""Plan 5 (warp-specialised pipelined stack kernels, nrx_stack_ws.cuh) against plan 1: every case in its own
process (a sticky CUDA error or a trapped barrier wait in one case cannot mask the others), bit-identity of all
outputs, and per-kernel-class event times.  usage: python tools/ws_check.py [label:n_prb:batch ...]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "%s")
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from tests.common import get_weights
label, n_prb, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = get_config(label); w,_ = get_weights(cfg); grid = build_grid(cfg, n_size_bwp=n_prb)
sb = make_slots(cfg, grid, batch=min(batch, 3), ebno_db=7.0, seed=200+n_prb)
idx = np.arange(batch) %% min(batch, 3)
y = torch.as_tensor(sb.y[idx]).cuda(); act = torch.as_tensor(sb.active_tx[idx]).cuda()
kw = {}
if cfg.num_io_stacks > 1:
    kw["io_index"] = torch.as_tensor(np.tile(np.array([[0, 1]], np.int32), (batch, 1))).cuda()
outs = {}
for mode in (1, 5):
    eng = NrxEngine(cfg, w, grid); eng.set_fused(mode)
    for rep in range(2):
        out = eng.forward(y, act, want=("llr","llr_grid","h_hat_refined","h_hat"), **kw)
        torch.cuda.synchronize()
    eng.set_profiling(True); eng.get_profile()
    for rep in range(5):
        eng.forward(y, act, want=("llr",), **kw)
    torch.cuda.synchronize()
    prof = eng.get_profile()
    outs[mode] = ({k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")},
                  {k: round(v["ms"] / max(v["launches"], 1) * 1e3, 1) for k, v in prof.items() if v["launches"]})
    eng.close()
same = all(np.array_equal(outs[1][0][k], outs[5][0][k]) for k in outs[1][0])
worst = max(float(np.abs(outs[1][0][k] - outs[5][0][k]).max()) for k in outs[1][0])
print("IDENTICAL" if same else "DIFFERENT max|d|=%%g" %% worst, "us/launch plan1", outs[1][1], "plan5", outs[5][1])
''' % ROOT
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
path = os.path.join(ROOT, "gpurun_out", "_ws_one.py")
open(path, "w").write(code)
cases = sys.argv[1:] or ["nrx_rt:4:3", "nrx_rt:1:2", "nrx_rt:11:5", "nrx_rt_var_mcs:7:3", "nrx_large:16:2", "nrx_rt:132:1", "nrx_large:132:30"]
for c in cases:
    label, prb, batch = c.split(":")
    try:
        r = subprocess.run([sys.executable, path, label, prb, batch], capture_output=True, text=True, timeout=240)
        print(c, (r.stdout.strip().splitlines() or ["-"])[-1], "|", (r.stderr.strip().splitlines() or ["-"])[-1][:200], flush=True)
    except subprocess.TimeoutExpired:
        print(c, "TIMEOUT", flush=True)
