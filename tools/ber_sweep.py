"""Slot-sharded Eb/N0 sweep on 1..8 GPUs: uncoded BER and bit-wise mutual information of the
receiver's LLRs (what can be evaluated without the third-party TB/LDPC chain).

    python tools/ber_sweep.py [--config nrx_large_64qam] [--slots 60] [--ebno -2 10 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 \
        --master-port P tools/ber_sweep.py ...

Mirrors the reference's Monte-Carlo loop (``scripts/evaluate.py:154-207``: batch 30, sim_ber with
``distribute="all"``): slot i of an SNR point is generated from seed ``1000*point + i`` whatever the
number of GPUs, slots are sharded over the ranks (``neural_rx_b200.distributed.slot_shard``), each
rank runs its shard in batches through the engine, and the only collective is one NCCL SUM of
{bit_errors, bits, slots, bmi_sum} per point — so the counters are identical for every G.
``--check-oracle K`` also pushes the first K slots of every point through the CPU oracle on
rank 0 and prints its BER next to the engine's for the same slots."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.distributed import slot_shard, sum_counters  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.synth import make_slots  # noqa: E402
from neural_rx_b200.weights import load_weights, random_weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="nrx_large_64qam")
    ap.add_argument("--slots", type=int, default=60, help="slots per Eb/N0 point (whole job)")
    ap.add_argument("--batch", type=int, default=30)
    ap.add_argument("--ebno", type=float, nargs=3, default=None, metavar=("MIN", "MAX", "STEP"))
    ap.add_argument("--n-prb", type=int, default=None)
    ap.add_argument("--check-oracle", type=int, default=0)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    from neural_rx_b200.engine import NrxEngine

    rank, local_rank = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = get_config(args.config)
    p = os.path.join(ROOT, "weights", f"{cfg.label}_weights")
    weights = load_weights(cfg, p) if os.path.exists(p) else random_weights(cfg, seed=0)
    grid = build_grid(cfg, n_size_bwp=args.n_prb)
    eng = NrxEngine(cfg, weights, grid, device=local_rank)
    lo, hi, step = args.ebno or (cfg.snr_db_eval_min, cfg.snr_db_eval_max, cfg.snr_db_eval_stepsize)
    points = np.arange(lo, hi + 1e-9, step)
    bits_per_sym = cfg.num_bits_per_symbol[0]
    if rank == 0:
        print(f"# {cfg.label}: {grid.num_subcarriers // 12} PRB, {args.slots} slots/point over {world} GPU(s), "
              f"{'shipped' if os.path.exists(p) else 'random'} weights")
        print("# ebno_db  bit_errors        bits        BER        BMI   slots   seconds" +
              ("   oracle_BER  engine_BER(same slots)" if args.check_oracle else ""))
    for pi, ebno in enumerate(points):
        s_lo, s_hi = slot_shard(args.slots, rank, world)
        t0 = time.perf_counter()
        c = {"bit_errors": 0, "bits": 0, "slots": 0, "bmi_micro": 0}
        first = []
        for b0 in range(s_lo, s_hi, args.batch):
            idx = range(b0, min(b0 + args.batch, s_hi))
            sbs = [make_slots(cfg, grid, batch=1, ebno_db=float(ebno), seed=1000 * pi + i) for i in idx]
            y = np.concatenate([s.y for s in sbs])
            bits = np.concatenate([s.bits for s in sbs])
            act = np.concatenate([s.active_tx for s in sbs])
            out = eng.forward(torch.as_tensor(y).cuda(), torch.as_tensor(act).cuda(), want=("llr",))
            llr = out["llr"].cpu().numpy()
            n = llr.shape[-1]
            b = bits[..., :n].astype(np.float32)
            err = ((llr > 0) != (b > 0.5)) & (act[..., None] > 0)
            # bit-wise mutual information estimate: 1 - E[log2(1 + exp(-(2b-1) llr))]
            bmi = 1.0 - np.logaddexp(0.0, -(2.0 * b - 1.0) * llr) / np.log(2.0)
            c["bit_errors"] += int(err.sum())
            c["bits"] += int((act > 0).sum()) * n
            c["slots"] += len(sbs)
            c["bmi_micro"] += int(round(float((bmi * (act[..., None] > 0)).sum()) * 1e6))
            if rank == 0 and len(first) < args.check_oracle:
                first += [(s, l) for s, l in zip(sbs, llr)][:args.check_oracle - len(first)]
        tot = sum_counters(c)
        line = None
        if rank == 0:
            dt = time.perf_counter() - t0
            line = (f"{ebno:8.2f} {tot['bit_errors']:11d} {tot['bits']:11d} {tot['bit_errors'] / max(tot['bits'], 1):10.3e} "
                    f"{tot['bmi_micro'] * 1e-6 / max(tot['bits'], 1):10.4f} {tot['slots']:7d} {dt:9.1f}")
            if first:
                from oracle import nrx_oracle as O
                from tests.common import oracle_arch, oracle_net
                net, arch = oracle_net(cfg, weights), oracle_arch(cfg)
                e_o = e_g = nb = 0
                for s, l in first:
                    ref = O.receiver_forward(net, arch, s.y, grid.pilots, grid.pilot_mask, s.active_tx)["llr"][0]
                    b = s.bits[0][..., :ref.shape[-1]] > 0.5
                    e_o += int(((ref > 0) != b).sum())
                    e_g += int(((l > 0) != b).sum())
                    nb += b.size
                line += f"   {e_o / nb:10.3e}  {e_g / nb:10.3e}"
            print(line, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
