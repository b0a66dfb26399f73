"""Run every execution plan (nrx_set_fused 1, 2, 0, 3) in its own process on a few small cases and
report success / the CUDA error per plan — a sticky device error in one plan cannot mask the others
(compute-sanitizer is not available on the GPU pool).  usage: python tools/plan_check.py"""
import os, sys, subprocess
ROOT='/root/repo'
code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "%s")
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from neural_rx_b200.engine import NrxEngine
from tests.common import get_weights
label, n_prb, batch, mode = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
cfg = get_config(label); w,_ = get_weights(cfg); grid = build_grid(cfg, n_size_bwp=n_prb)
sb = make_slots(cfg, grid, batch=batch, ebno_db=7.0, seed=200+n_prb)
eng = NrxEngine(cfg, w, grid); eng.set_fused(mode)
for rep in range(3):
    out = eng.forward(torch.as_tensor(sb.y).cuda(), torch.as_tensor(sb.active_tx).cuda(), want=("llr","llr_grid","h_hat_refined","h_hat"))
    torch.cuda.synchronize()
print("OK", float(out["llr"].abs().mean()))
''' % ROOT
open('gpurun_out/_one.py','w').write(code)
for case in [("nrx_rt",4,3),("nrx_rt",1,2)]:
    for mode in (1, 2, 0, 3):
        r = subprocess.run([sys.executable,'gpurun_out/_one.py',case[0],str(case[1]),str(case[2]),str(mode)],capture_output=True,text=True,env=dict(os.environ,CUDA_LAUNCH_BLOCKING="1"))
        print(case, mode, (r.stdout.strip().splitlines() or ['-'])[-1], '|', (r.stderr.strip().splitlines() or ['-'])[-1][:160])
