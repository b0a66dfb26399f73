"""Timing experiments on the aggregation kernel: builds libnrx_b200_<tag>.so with extra -D flags and reports
the per-launch CUDA-event time of the aggregation kernel for nrx_large, batch 30 (results are NOT checked: some
experiments deliberately skip work).  usage: python tools/agg_exp.py tag=FLAG[,FLAG] ...   (--build-only to compile here)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_rx_b200 import build as B
variants = [a.split("=", 1) for a in sys.argv[1:] if "=" in a]
libs = {"base": os.path.join(ROOT, "neural_rx_b200", "libnrx_b200.so")}
for tag, flags in variants:
    lib = os.path.join(ROOT, "neural_rx_b200", f"libnrx_b200_{tag}.so")
    if "--build-only" in sys.argv or not os.path.exists(lib):
        subprocess.run([B._nvcc()] + B.NVCC_FLAGS + ["-D" + f for f in flags.split(",") if f] + ["-o", lib] + B.SOURCES, check=True)
    libs[tag] = lib
if "--build-only" in sys.argv:
    sys.exit(0)
code = r'''
import sys, os, numpy as np, torch
sys.path.insert(0, "%s")
from neural_rx_b200 import engine as E
E._LIB_PATH = sys.argv[1]
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots
from tests.common import get_weights
cfg = get_config("nrx_large"); w,_ = get_weights(cfg); grid = build_grid(cfg)
sb = make_slots(cfg, grid, batch=1, ebno_db=4.0, seed=1)
y = torch.as_tensor(np.repeat(sb.y, 30, axis=0)).cuda(); act = torch.ones((30, 2), device="cuda")
eng = E.NrxEngine(cfg, w, grid); eng.set_fused(int(sys.argv[2]))
for _ in range(3): eng.forward(y, act, want=("llr", "h_hat_refined"))
torch.cuda.synchronize(); eng.set_profiling(True); eng.get_profile()
for _ in range(10): eng.forward(y, act, want=("llr", "h_hat_refined"))
torch.cuda.synchronize(); pr = eng.get_profile()
print({k: round(v["ms"] / max(v["launches"], 1) * 1e3, 1) for k, v in pr.items() if v["launches"] and k.startswith("agg")})
''' % ROOT
path = os.path.join(ROOT, "gpurun_out", "_agg_exp.py")
os.makedirs(os.path.dirname(path), exist_ok=True)
open(path, "w").write(code)
for tag, lib in libs.items():
    for plan in (1,):
        r = subprocess.run([sys.executable, path, lib, str(plan)], capture_output=True, text=True, timeout=300)
        print(tag, "plan", plan, (r.stdout.strip().splitlines() or ["-"])[-1], "|", (r.stderr.strip().splitlines() or ["-"])[-1][:160], flush=True)
