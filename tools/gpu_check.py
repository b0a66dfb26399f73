"""Diagnostic: run the CUDA engine against the oracle on a few cases and print parity metrics.
usage: python tools/gpu_check.py [label:n_prb:batch ...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_rx_b200.config import get_config
from neural_rx_b200.pusch import build_grid
from neural_rx_b200.synth import make_slots, uncoded_ber
from neural_rx_b200.engine import NrxEngine
from oracle import nrx_oracle as O
from tests.common import get_weights, oracle_arch, oracle_net, rel_l2, sign_agreement, ENGINE_EMU

cases = sys.argv[1:] or ["nrx_rt:4:2", "nrx_rt:16:2", "nrx_rt:132:1"]
for c in cases:
    label, prb, batch = c.split(":")
    prb, batch = int(prb), int(batch)
    cfg = get_config(label)
    w, kind = get_weights(cfg)
    grid = build_grid(cfg, n_size_bwp=prb)
    sb = make_slots(cfg, grid, batch=batch, ebno_db=8.0, seed=11)
    t0 = time.time()
    eng = NrxEngine(cfg, w, grid)
    y = torch.as_tensor(sb.y).cuda(); act = torch.as_tensor(sb.active_tx).cuda()
    out = eng.forward(y, act, want=("llr", "llr_grid", "h_hat_refined", "h_hat"))
    torch.cuda.synchronize()
    got = {k: v.cpu().numpy() for k, v in out.items() if not k.startswith("_")}
    t1 = time.time()
    arch = oracle_arch(cfg); net = oracle_net(cfg, w)
    ref = O.receiver_forward(net, arch, sb.y, grid.pilots, grid.pilot_mask, sb.active_tx)
    emu = O.receiver_forward(net, arch, sb.y, grid.pilots, grid.pilot_mask, sb.active_tx, emu=ENGINE_EMU)
    print(f"== {c} weights={kind} engine {t1-t0:.2f}s  BER gpu {uncoded_ber(got['llr'], sb.bits, sb.active_tx, 4):.4f} oracle {uncoded_ber(ref['llr'], sb.bits, sb.active_tx, 4):.4f}")
    print(f"   h_hat(LS)   relL2 vs oracle {rel_l2(got['h_hat'], ref['h_hat']):.3e}")
    for name, r in (("exact", ref), ("emul ", emu)):
        print(f"   vs {name}: llr relL2 {rel_l2(got['llr'], r['llr']):.3e} agree {100*sign_agreement(got['llr'], r['llr']):.3f}%  "
              f"grid relL2 {rel_l2(got['llr_grid'], r['llr_grid'][0]):.3e}  h_ref relL2 {rel_l2(got['h_hat_refined'], r['h_hat_refined']):.3e}")
    hs = eng.forward_host(sb.y, sb.active_tx, want=("llr",))
    print(f"   host call == device call: {np.array_equal(hs['llr'], got['llr'])}")
    eng.close()
