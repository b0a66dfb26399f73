"""Coded BER / BLER Monte-Carlo sweep on 1..8 GPUs — the loop of ``scripts/evaluate.py:154-207`` (``sim_ber`` over
an Eb/N0 grid, batch 30, ``distribute="all"``) around the CUDA receiver, SURVEY.md §8f-1:

    TB bits -> neural_rx_b200.tb.TBEncoder -> QAM + DMRS on the resource grid -> synthetic channel + AWGN
            -> NeuralPUSCHReceiver (CUDA engine) -> LLRs -> tb.TBDecoder (20 box-plus BP iterations) -> b_hat, CRC

    python tools/bler_sweep.py [--config nrx_rt] [--n-prb 24] [--slots 60] [--ebno -2 8 1] [--check-oracle 4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \
        tools/bler_sweep.py ...

Slot i of point p is generated from seed 1000 p + i whatever the number of GPUs; slots are sharded over the ranks
and the only collective is one NCCL SUM of {bit_errors, bits, block_errors, blocks} per point, so the counters do not
depend on G.  Early stop per point like ``sim_ber``: ``--target-block-errors`` (500 in the reference).

The LDPC code is the 3GPP one when the TS 38.212 base-graph tables are available (Sionna installed, or
$NRX_LDPC_BG_DIR), else the structural stand-in of neural_rx_b200/tb.py — the header line says which.  With the stand-in
and the synthetic (TDL-like) channel the curve is NOT comparable with the reference's ``results/*_results``; what it
does show is whether the CUDA LLRs decode like the oracle's: ``--check-oracle K`` pushes the first K slots of every
point through the CPU oracle on rank 0 and prints block errors of both LLR sources for the same slots."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neural_rx_b200 import tb as TB  # noqa: E402
from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.bler import sim_point  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.weights import load_weights, random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="nrx_rt")
    ap.add_argument("--slots", type=int, default=60, help="slots per Eb/N0 point (whole job)")
    ap.add_argument("--batch", type=int, default=30)
    ap.add_argument("--ebno", type=float, nargs=3, default=None, metavar=("MIN", "MAX", "STEP"))
    ap.add_argument("--n-prb", type=int, default=None)
    ap.add_argument("--base-graph", default="auto", choices=("auto", "3gpp", "standin"))
    ap.add_argument("--target-block-errors", type=int, default=500)
    ap.add_argument("--check-oracle", type=int, default=0)
    ap.add_argument("--active", type=int, nargs="+", default=None,
                    help="transmitters that are on (e.g. 1 0: the single-UE evaluation of results/nrx_rt_results); default all")
    ap.add_argument("--published", default=None, metavar="RESULTS_FILE",
                    help="print the reference's published BLER (tests/golden/ref_published_curves.json, e.g. nrx_large_results "
                         "or nrx_large_sweep_results with --num-it) next to the measured one, for orientation")
    ap.add_argument("--num-it", type=int, nargs="+", default=None,
                    help="CGNN iterations to evaluate (num_it sweep of results/nrx_large_sweep_results); default: the cfg's")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from neural_rx_b200.receiver import NeuralPUSCHReceiver

    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = get_config(args.config)
    p = os.path.join(ROOT, "weights", f"{cfg.label}_weights")
    weights = load_weights(cfg, p) if os.path.exists(p) else random_weights(cfg, seed=0)
    grid = build_grid(cfg, n_size_bwp=args.n_prb)
    rx = NeuralPUSCHReceiver(cfg, weights=weights, grid=grid, device=local_rank, tb_decoding="off")
    enc = TB.pusch_tb_encoder(cfg, grid, 0, base_graph=args.base_graph)
    dec = TB.TBDecoder(enc, num_bp_iter=20, cn_type="boxplus")
    lo, hi, step = args.ebno or (cfg.snr_db_eval_min, cfg.snr_db_eval_max, cfg.snr_db_eval_stepsize)
    points = np.arange(lo, hi + 1e-9, step)
    U = grid.num_tx
    if rank == 0:
        print(f"# {cfg.label}: {grid.num_subcarriers // 12} PRB, {args.slots} slots/point over {world} GPU(s), "
              f"{'shipped' if os.path.exists(p) else 'random'} weights; TBS {enc.tb_size}, {enc.num_cbs} code block(s) of "
              f"{enc.k_prime} bits, BG{enc.bg} Z = {enc.z}, G = {enc.num_coded_bits}; base graph: {enc.base_graph_source}"
              + ("  (structural stand-in, NOT the TS 38.212 code: not comparable with results/*_results)"
                 if enc.base_graph_source != "3gpp" else ""))
        print("# ebno_db  bit_errors        bits        BER  block_errors  blocks       BLER   seconds" +
              ("   oracle_block_errors  engine_block_errors  (same slots)  decisions_differ" if args.check_oracle else ""))
    def llr_fn(y, act):
        return rx.llrs((torch.as_tensor(y).cuda(), torch.as_tensor(act).cuda()), want=("llr",))["llr"]

    def run_point(pi, ebno):
        t0 = time.perf_counter()
        tot, first = sim_point(llr_fn, cfg, grid, enc, dec, pi, float(ebno), args.slots, args.batch, rank, world,
                               args.target_block_errors, args.check_oracle, args.active)
        if rank != 0:
            return
        dt = time.perf_counter() - t0
        line = (f"{ebno:8.2f} {tot['bit_errors']:11d} {tot['bits']:11d} {tot['bit_errors'] / max(tot['bits'], 1):10.3e} "
                f"{tot['block_errors']:13d} {tot['blocks']:7d} {tot['block_errors'] / max(tot['blocks'], 1):10.3e} {dt:9.1f}")
        if first:                                       # test aid: the same slots through the CPU oracle (test infrastructure)
            from oracle import nrx_oracle as O
            from tests.common import oracle_arch, oracle_net
            net, arch = oracle_net(cfg, weights), oracle_arch(cfg)
            e_o = e_g = diff = 0
            for ys, a, ok_g, bh_g in first:
                ref = O.receiver_forward(net, arch, ys[None], grid.pilots, grid.pilot_mask, a[None], num_it=rx.num_it)["llr"]
                bh_o, ok_o = dec(np.asarray(ref, np.float32))
                e_o += int((~ok_o[0] & (a > 0)).sum())
                e_g += int((~ok_g & (a > 0)).sum())
                diff += int(((ok_o[0] != ok_g) & (a > 0)).sum())
            line += f"   {e_o:19d}  {e_g:19d}  {len(first) * U:12d}  {diff:16d}"
        if args.published:
            from tools.ref_results import published_at
            n_on = int(sum(args.active)) if args.active else U
            key = f"Neural Receiver|{n_on}|0" + (f"|{rx.num_it}" if "sweep" in args.published else "")
            ref = published_at(args.published, key, float(ebno))
            line += f"   published[{args.published}: {key}] " + ("n/a" if ref is None else f"{ref:.3e}")
        print(line, flush=True)

    for num_it in (args.num_it or [rx.num_it]):
        rx.num_it = num_it                              # utils/neural_rx.py:537-542: iterations can be dropped after training
        if rank == 0 and args.num_it:
            print(f"# num_it = {num_it}")
        for pi, ebno in enumerate(points):
            run_point(pi, ebno)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
