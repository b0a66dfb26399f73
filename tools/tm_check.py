"""Plan 4 (TMEM-resident UpdateState stacks) against plan 1 on a few cases, each plan in its own
process (a sticky CUDA error cannot mask the other), then a timing of both plans.
usage: python tools/tm_check.py [quick]"""
import os, sys, subprocess
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time, numpy as np, torch
sys.path.insert(0, "%s")
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from tests.common import get_weights
label, n_prb, batch, mode, out, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6])
cfg = get_config(label); w,_ = get_weights(cfg); grid = build_grid(cfg, n_size_bwp=n_prb)
sb = make_slots(cfg, grid, batch=min(batch, 4), ebno_db=7.0, seed=200+n_prb)
y = np.tile(sb.y, ((batch + 3) // 4,) + (1,) * (sb.y.ndim - 1))[:batch]
a = np.tile(sb.active_tx, ((batch + 3) // 4, 1))[:batch]
eng = NrxEngine(cfg, w, grid); eng.set_fused(mode)
yt, at = torch.as_tensor(y).cuda(), torch.as_tensor(a).cuda()
for rep in range(2):
    o = eng.forward(yt, at, want=("llr", "h_hat_refined"))
    torch.cuda.synchronize()
np.savez(out, llr=o["llr"].cpu().numpy(), h=o["h_hat_refined"].cpu().numpy())
if reps:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): eng.forward(yt, at, want=("llr", "h_hat_refined"))
    e1.record(); torch.cuda.synchronize()
    print("MS_PER_FORWARD %%.4f" %% (e0.elapsed_time(e1) / reps))
print("OK", float(np.abs(o["llr"].cpu().numpy()).mean()))
''' % ROOT
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
one = os.path.join(ROOT, "gpurun_out", "_tm_one.py")
open(one, "w").write(code)
cases = [("nrx_rt", 4, 3, 0), ("nrx_rt", 1, 2, 0), ("nrx_large", 16, 2, 0), ("nrx_rt", 132, 1, 0), ("nrx_large", 132, 30, 20)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    cases = cases[:2]
for label, n_prb, batch, reps in cases:
    res = {}
    for mode in (1, 4):
        out = os.path.join(ROOT, "gpurun_out", f"_tm_{mode}.npz")
        if os.path.exists(out): os.remove(out)
        r = subprocess.run(["timeout", "120", sys.executable, one, label, str(n_prb), str(batch), str(mode), out, str(reps)],
                           capture_output=True, text=True)
        tail = (r.stdout.strip().splitlines() or ["-"])
        err = (r.stderr.strip().splitlines() or ["-"])[-1][:200]
        print(label, n_prb, batch, "plan", mode, "rc", r.returncode, " | ".join(tail[-2:]), "|", err, flush=True)
        if os.path.exists(out): res[mode] = np.load(out)
    if 1 in res and 4 in res:
        for k in ("llr", "h"):
            a, b = res[1][k].astype(np.float64), res[4][k].astype(np.float64)
            print("   ", k, "max|diff|", float(np.abs(a - b).max()), "rel_l2", float(np.linalg.norm(a - b) / max(np.linalg.norm(a), 1e-30)),
                  "identical", bool(np.array_equal(res[1][k], res[4][k])), "nan", int(np.isnan(b).sum()), flush=True)
