#!/usr/bin/env python
"""Benchmark of the neural-receiver full-slot hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

One *step* = one pass of the hot path over one batch of ``batch_size_eval`` = 30 synthetic slots of
BASELINE.json ``configs[1]`` (nrx_large, 132 PRB, 2 UEs, 4 rx antennas, 16-QAM) per GPU.  Prints ONE
JSON line: ``value`` = whole-job slots/s with inputs resident in HBM, ``e2e`` = the same through
the host-buffer C-ABI call on page-locked host buffers (every step's H2D and D2H copies inside the
timed region; two asynchronous calls in flight, the serving pattern), plus ``roofline`` (dominant
kernel, CUDA events around every launch, fraction of the burst AND of the sustained measured tensor
peak), ``sustained`` (the same loop run back to back for >= 5 s), ``cpu_baseline`` (the oracle port on
the host cores; rank 0 at N = 1 only, like the other single-GPU side measurements), ``latency_us`` (batch-1 p50 / p99 for nrx_rt and nrx_large: device-resident and through
host buffers) and ``clocks``.

``--impl reference`` times the reference's own CPU path on the same workload: whole 30-slot steps on all
host threads.  TensorFlow and Sionna are not installable here and the fork's torch port does not run
(SURVEY.md §0), so that arm is the oracle restatement (``oracle/nrx_oracle.py``, PyTorch-CPU fp32).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.synth import make_slots  # noqa: E402
from neural_rx_b200.weights import load_weights, random_weights  # noqa: E402

WORKLOAD = "nrx_large"
METRIC = "nrx_slots_per_s"
UNIT = "slots/s"


def workload_desc(cfg):
    """The SAME string in both arms (the driver compares them)."""
    return (f"{cfg.label}.cfg: {cfg.n_size_bwp} PRB x 14 symbols, "
            f"{cfg.max_num_tx} UEs, {cfg.num_rx_antennas} rx antennas, 16-QAM, {cfg.num_nrx_iter} CGNN iterations, "
            f"batch {cfg.batch_size_eval} slots per GPU per step")


def source_hash():
    """Hash of the kernel sources: keys the committed ncu traffic figures to the code they were measured on
    (the GPU box has no .git)."""
    import hashlib
    h = hashlib.sha256()
    d = os.path.join(ROOT, "neural_rx_b200", "csrc")
    for name in sorted(os.listdir(d)):
        with open(os.path.join(d, name), "rb") as f:
            h.update(name.encode() + b"\0" + f.read())
    return h.hexdigest()[:16]


def _weights(cfg):
    p = os.path.join(ROOT, "weights", f"{cfg.label}_weights")      # staged copy of the reference's weight file
    if os.path.exists(p):
        return load_weights(cfg, p), "shipped weight file"
    return random_weights(cfg, seed=0), "random-init weights"


def _measured_peaks():
    """(burst TFLOP/s, sustained TFLOP/s, HBM GB/s, source)."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1659.2), d.get("bf16_tflops_sustained", 1388.1), d.get("hbm_gbs", 6523.3), "MEASURED_PEAKS.json"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs.  Rank 0 samples its own GPU four
    times a second; the other ranks do not sample (every nvidia-smi query goes through driver-wide locks: eight
    ranks polling ten times a second were measured to cost the 8-GPU end-to-end run several percent)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown," \
        "clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.rows, self._stop, self._t, self.enabled = index, [], threading.Event(), None, enabled

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if out.returncode == 0 and out.stdout.strip():
                    self.rows.append([x.strip() for x in out.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.25)

    def __enter__(self):
        if self.enabled:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def oracle_forward_timer(cfg, weights, grid, threads: int):
    """Returns run(y, active, chunk) -> seconds for one pass of the CPU restatement of the path (oracle) over the
    slots of ``y`` in sub-batches of ``chunk`` slots (1 = slot by slot, as the reference's evaluation loop would call
    it with batch 1; larger = batched tensors)."""
    import torch
    from oracle import nrx_oracle as O
    from tests.common import oracle_arch
    torch.set_num_threads(threads)
    arch = oracle_arch(cfg)
    net = O.bind_weights(arch, weights.to_list())
    tables = dict(nn=grid.nn_index, pe=grid.pos_enc)

    def run(y, active, chunk):
        t0 = time.perf_counter()
        for j in range(0, y.shape[0], chunk):
            O.receiver_forward(net, arch, y[j:j + chunk], grid.pilots, grid.pilot_mask, active[j:j + chunk], tables=tables)
        return time.perf_counter() - t0
    return run


def pick_cpu_mode(run, sb):
    """Slot-by-slot or batched (5 slots per call: the activations of more slots no longer fit the host caches)?
    Timed on 5 slots each; the faster one is used."""
    run(sb.y[:1], sb.active_tx[:1], 1)                                          # warm-up (thread pool, allocator)
    t1 = run(sb.y[:5], sb.active_tx[:5], 1)
    t5 = run(sb.y[:5], sb.active_tx[:5], 5)
    return (5, t1, t5) if t5 < t1 else (1, t1, t5)


def run_reference(args, rank):
    """Reference arm: CPU implementation of the path on the host cores (rank 0 only): every step is one pass over
    the whole 30-slot batch of the workload."""
    if rank != 0:
        return
    cfg = get_config(WORKLOAD)
    weights, wsrc = _weights(cfg)
    grid = build_grid(cfg)
    B = cfg.batch_size_eval
    sb = make_slots(cfg, grid, batch=B, ebno_db=np.linspace(-2, 6, B), seed=1000)
    cores = os.cpu_count() or 1
    run = oracle_forward_timer(cfg, weights, grid, cores)
    chunk, t1, t5 = pick_cpu_mode(run, sb)
    for _ in range(max(args.warmup - 1, 0)):
        run(sb.y[:chunk], sb.active_tx[:chunk], chunk)
    t_steps = [run(sb.y, sb.active_tx, chunk) for _ in range(args.steps)]
    t_total = float(sum(t_steps))
    v = B * args.steps / t_total
    sample = (f"{args.steps} whole steps of {B} slots, {'slot by slot' if chunk == 1 else f'{chunk} slots per call'} "
              f"(5-slot probe: {t1:.2f} s slot by slot, {t5:.2f} s batched), {wsrc}")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": f"synthetic ({wsrc})",
        "config": {"workload": workload_desc(cfg),
                   "note": "TensorFlow/Sionna reference cannot run offline; its CPU restatement (oracle port) is timed"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def latency_percentiles(eng, y1, act1, n: int = 200):
    """Batch-1 latency of one forward (CUDA events around it): eager launches through the C ABI and
    the same forward replayed from a CUDA graph (NrxEngine.capture)."""
    import torch
    want = ("llr", "h_hat_refined")

    def measure(fn):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            b.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts = np.sort(np.asarray(ts))
        return float(ts[len(ts) // 2]), float(ts[min(len(ts) - 1, int(0.99 * len(ts)))])

    outs = {}
    e50, e99 = measure(lambda: eng.forward(y1, act1, want=want, out=outs))
    graph, _ = eng.capture(y1, act1, want=want)
    g50, g99 = measure(graph.replay)
    # the same slot through HOST buffers (page-locked): H2D of y, kernels, D2H of the LLRs and the refined channel
    # estimate, wall clock around the blocking C-ABI call — comparable with the reference's published end-to-end
    # latency (notebooks/real_time_nrx.ipynb:580-584: 1.409 ms incl. 0.050 H2D / 0.085 D2H, RTX 3090, TensorRT)
    from neural_rx_b200.engine import pinned_empty
    y_h = pinned_empty(tuple(y1.shape), np.complex64)
    y_h[...] = y1.cpu().numpy()
    a_h = pinned_empty(tuple(act1.shape), np.float32)
    a_h[...] = act1.cpu().numpy()
    g_ = eng.grid
    out_h = {"llr": pinned_empty((1, g_.num_tx, g_.num_data_res * eng.cfg.num_bits_per_symbol[0])),
             "h_hat_refined": pinned_empty((1, g_.num_tx, g_.num_subcarriers, g_.num_ofdm_symbols, 2 * eng.cfg.num_rx_antennas))}
    for _ in range(10):
        eng.forward_host(y_h, a_h, want=want, out=out_h)
    th = []
    for _ in range(n):
        t0 = time.perf_counter()
        eng.forward_host(y_h, a_h, want=want, out=out_h)
        th.append((time.perf_counter() - t0) * 1e6)
    th = np.sort(np.asarray(th))
    return {"p50": g50, "p99": g99, "eager_p50": e50, "eager_p99": e99, "host_p50": float(th[len(th) // 2]),
            "host_p99": float(th[min(len(th) - 1, int(0.99 * len(th)))]), "n": n,
            "mode": "p50/p99: cuda graph replay, device-resident; eager_*: plain launches; host_*: blocking host-buffer call "
                    "(pinned), H2D + kernels + D2H, wall clock"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--slots-per-pass", type=int, default=int(os.environ.get("NRX_SLOTS_PER_PASS", "0")))
    ap.add_argument("--fused", type=int, default=int(os.environ.get("NRX_FUSED", "6")),
                    help="execution plan of the sep-conv stacks (nrx_set_fused)")
    ap.add_argument("--host-chunk", type=int, default=int(os.environ.get("NRX_HOST_CHUNK", "0")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from neural_rx_b200.distributed import max_over_ranks, sum_counters
    from neural_rx_b200.engine import NrxEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the receiver has no CPU path)")
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # one disjoint block of host cores per rank: the ranks' copy / launch threads do not migrate over each other
        try:
            cpus = sorted(os.sched_getaffinity(0))
            per = max(len(cpus) // world, 1)
            os.sched_setaffinity(0, set(cpus[local_rank * per:(local_rank + 1) * per] or cpus))
        except OSError:
            pass
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = get_config(WORKLOAD)
    weights, wsrc = _weights(cfg)
    grid = build_grid(cfg)
    B = cfg.batch_size_eval
    eng = NrxEngine(cfg, weights, grid, device=local_rank)
    if args.slots_per_pass:
        eng.set_slots_per_pass(args.slots_per_pass)
    eng.set_fused(args.fused)

    # distinct synthetic batches per rank; rotating over NBUF x 21 MB of inputs (> 126 MB L2)
    NBUF = 8
    base = make_slots(cfg, grid, batch=B, ebno_db=np.linspace(-2, 6, B), seed=1000 + rank)
    rng = np.random.default_rng(77 + rank)
    ys_host, ys = [], []
    for i in range(NBUF):
        yi = base.y if i == 0 else (base.y * np.exp(1j * rng.uniform(0, 2 * np.pi)) +
                                    0.01 * (rng.standard_normal(base.y.shape) + 1j * rng.standard_normal(base.y.shape))
                                    ).astype(np.complex64)
        ys_host.append(np.ascontiguousarray(yi))
        ys.append(torch.as_tensor(yi).to(dev))
    act = torch.as_tensor(base.active_tx).to(dev)
    want = ("llr", "h_hat_refined")
    outs = {}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----------------------------------------------------
    for i in range(args.warmup):
        eng.forward(ys[i % NBUF], act, want=want, out=outs)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank, enabled=rank == 0)      # samples through the timed region and the profiled / e2e repeats of it
    clk.__enter__()
    e0.record()
    for i in range(args.steps):
        eng.forward(ys[i % NBUF], act, want=want, out=outs)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    ms = max_over_ranks(ms)
    value = world * B * args.steps / (ms * 1e-3)
    launches = eng.launches_per_forward(B) * args.steps

    # ---- per-kernel durations (second pass over the same steps, events around every launch) ------
    eng.set_profiling(True)
    for i in range(args.steps):
        eng.forward(ys[i % NBUF], act, want=want, out=outs)
    prof = eng.get_profile()
    eng.set_profiling(False)
    peak_burst, peak_sust, peak_hbm, peak_src = _measured_peaks()
    clocks_mid = clk.summary()
    # which peak applies: the burst figure while the SM clock sits at its maximum with no power capping (a short
    # run), the sustained one once the part is power-capped
    uncapped = (clocks_mid.get("sm_mhz") or 0) >= 0.95 * (clocks_mid.get("sm_max_mhz") or 1e9) and \
        "sw_power_cap" not in clocks_mid.get("reasons", [])
    peak_tf = peak_burst if uncapped else peak_sust
    P_step = B * grid.num_tx * grid.num_subcarriers * grid.num_ofdm_symbols     # user-REs per step
    dom = max(prof, key=lambda k: prof[k]["ms"])
    # algorithmic FLOPs of the dominant kernel per launch = 2 * MACs/pixel of that layer * pixels per launch
    cin0, ds, hid = 4 * cfg.num_rx_antennas + 2, cfg.d_s, cfg.num_units_state[0][0]
    mac_layer = {"sep_128x128": 9 * hid + hid * hid, "sep_32x128": 9 * cin0 + cin0 * hid,
                 "sep_128x64_init_out": 9 * hid + hid * ds, "sep_128x64_update_out": 9 * hid + hid * ds,
                 "stack_init": (9 + hid) * cin0 + (9 + hid) * hid + (9 + ds) * hid,
                 "stack_update": (9 + hid) * (2 * ds + 2) + (9 + hid) * hid + (9 + ds) * hid,
                 "agg": 2 * ds * cfg.num_units_agg[0][0],
                 "readout": 2 * ds * cfg.num_units_readout[0]
                 + cfg.num_units_readout[0] * (cfg.num_bits_per_symbol[0] + 2 * cfg.num_rx_antennas)}
    total_kernel_ms = sum(v["ms"] for v in prof.values())
    roof = None
    if dom in mac_layer and prof[dom]["launches"]:
        n_l = prof[dom]["launches"]
        launches_per_step = n_l / args.steps
        # every launch of a layer class covers all pixels of one pass; passes per step * layers per pass = launches
        passes = (B + (args.slots_per_pass or B) - 1) // (args.slots_per_pass or B)
        pixels_per_launch = P_step / passes
        flops_per_launch = 2.0 * mac_layer[dom] * pixels_per_launch
        dur_s = prof[dom]["ms"] * 1e-3 / n_l
        achieved = flops_per_launch / dur_s / 1e12
        # DRAM bytes per launch of that kernel from this round's committed `ncu --set full` capture — only if it was
        # taken on the kernel sources that are running now (tools/ncu_summary.py stores their hash)
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r02b_traffic.json")
        ncu_name = {"stack_update": "nrx_stack_ws_kernel<1>" if args.fused >= 5 else "nrx_stack_kernel<1, 0>",
                    "stack_init": "nrx_stack_ws_kernel<0>" if args.fused == 5 else "nrx_stack_kernel<0, 0>",
                    "agg": "nrx_agg_ws_kernel", "readout": "nrx_readout_kernel"}.get(dom)
        if os.path.exists(tpath) and ncu_name and passes == 1:
            with open(tpath) as f:
                tj = json.load(f)
            if tj.get("source_hash") == source_hash():
                traffic = tj.get("kernels", {}).get(ncu_name, {}).get("dram_bytes_per_launch")
                traffic_src = f"profiles/r02b_traffic.json (ncu --set full, source hash {tj.get('source_hash')})"
            else:
                traffic_src = "omitted: profiles/r02b_traffic.json was captured on other kernel sources"
        roof = {"bound": "tensor", "kernel": dom, "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf, "frac_of_burst_peak": achieved / peak_burst,
                "frac_of_sustained_peak": achieved / peak_sust,
                "peak_used": "burst" if uncapped else "sustained", "traffic": traffic, "traffic_source": traffic_src,
                "traffic_unit": "bytes per launch (ncu dram read+write)",
                # per user-RE: 11/9 x 256 B input window + 128 B residual re-read + 128 B new state
                "algorithmic_bytes_per_launch": 340.0 * pixels_per_launch if dom == "stack_update" else None,
                "peak_source": peak_src,
                "launch_us": dur_s * 1e6, "launches_per_step": launches_per_step,
                "share_of_kernel_time": prof[dom]["ms"] / max(total_kernel_ms, 1e-9)}
    whole = {"achieved_tflops": eng.flops_per_slot() * value / world / 1e12,
             "frac_of_peak": eng.flops_per_slot() * value / world / 1e12 / peak_tf,
             "frac_of_burst_peak": eng.flops_per_slot() * value / world / 1e12 / peak_burst}

    # ---- steady state: the same loop back to back for >= 5 s (a power-cap effect would show here) -------
    sust_clk = ClockSampler(local_rank, enabled=rank == 0)
    sust_clk.__enter__()
    n_sust = max(int(5.5 / max(ms / args.steps * 1e-3, 1e-6)), args.steps)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(n_sust):
        eng.forward(ys[i % NBUF], act, want=want, out=outs)
    s1.record()
    barrier()
    sust_ms = max_over_ranks(s0.elapsed_time(s1))
    sust_clk.__exit__(None, None, None)
    sustained = {"seconds": sust_ms * 1e-3, "steps": n_sust, "value": world * B * n_sust / (sust_ms * 1e-3), "unit": UNIT,
                 "clocks": sust_clk.summary()}

    # ---- end to end through the host-buffer C-ABI call ---------------------------------------------
    # inputs and outputs live in page-locked host memory (the contract's "pinned host memory"); every
    # step copies that step's y host->device and the LLRs + refined channel estimate device->host
    from neural_rx_b200.engine import pinned_empty
    ys_pin = []
    for yi in ys_host:
        a = pinned_empty(yi.shape, np.complex64)
        a[...] = yi
        ys_pin.append(a)
    g_ = grid
    out_pin = {"llr": pinned_empty((B, g_.num_tx, g_.num_data_res * cfg.num_bits_per_symbol[0])),
               "h_hat_refined": pinned_empty((B, g_.num_tx, g_.num_subcarriers, g_.num_ofdm_symbols,
                                              2 * cfg.num_rx_antennas))}
    if args.host_chunk:
        eng.set_host_chunk(args.host_chunk)
    # two result sets and two calls in flight: step i+1 is enqueued before step i is waited for, so its copy-in
    # overlaps step i's kernels and copy-outs (nrx_forward_host_async / nrx_wait)
    act_pin = pinned_empty(base.active_tx.shape, np.float32)
    act_pin[...] = base.active_tx
    out_pins = [out_pin, {k: pinned_empty(v.shape) for k, v in out_pin.items()}]
    for i in range(3):
        eng.wait(eng.forward_host_async(ys_pin[i % NBUF], act_pin, out_pins[i % 2]))
    barrier()
    t0 = time.perf_counter()
    prev = None
    for i in range(args.steps):
        t = eng.forward_host_async(ys_pin[i % NBUF], act_pin, out_pins[i % 2])
        if prev is not None:
            eng.wait(prev)
        prev = t
    res = eng.wait(prev)
    barrier()
    dt = time.perf_counter() - t0
    dt = max_over_ranks(dt)
    h2d = ys_pin[0].nbytes + act_pin.nbytes
    d2h = sum(v.nbytes for v in res.values())
    e2e = {"value": world * B * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "host_memory": "pinned", "calls_in_flight": 2,
           "host_chunk_slots": args.host_chunk or min(32, B)}
    # one blocking call per step (nothing overlaps between steps)
    t0 = time.perf_counter()
    for i in range(max(args.steps // 2, 1)):
        eng.forward_host(ys_pin[i % NBUF], base.active_tx, want=want, out=out_pin)
    barrier()
    e2e["blocking_call_value"] = world * B * max(args.steps // 2, 1) / max_over_ranks(time.perf_counter() - t0)
    # same call with ordinary (pageable) NumPy arrays, staged through the engine's pinned buffers
    t0 = time.perf_counter()
    for i in range(max(args.steps // 2, 1)):
        eng.forward_host(ys_host[i % NBUF], base.active_tx, want=want)
    barrier()
    e2e["pageable_value"] = world * B * max(args.steps // 2, 1) / max_over_ranks(time.perf_counter() - t0)

    clk.__exit__(None, None, None)
    # ---- counters over NCCL (the only collective: slots processed per rank) -------------------------
    slots_done = sum_counters({"slots": B * args.steps})["slots"]

    if rank == 0:
        extra = {}
        # single-GPU side measurements (batch-1 latency, nrx_rt, skipping, CPU baseline) run at N = 1 only: at N > 1 the
        # other ranks would sit idle behind rank 0 for minutes (torchrun also pins OMP to one thread per rank)
        single = world == 1
        if single and not args.no_latency:
            # batch-1 latency uses plan 2 (message MLP fused into the stack kernels: 12 instead of 20
            # launches for nrx_large); throughput above uses the default plan
            lat = {}
            y1, a1 = ys[0][:1].contiguous(), act[:1].contiguous()
            eng.set_fused(2)
            lat["nrx_large"] = latency_percentiles(eng, y1, a1)
            eng.set_fused(args.fused)
            cfg_rt = get_config("nrx_rt")
            w_rt, _ = _weights(cfg_rt)
            eng_rt = NrxEngine(cfg_rt, w_rt, grid, device=local_rank)
            eng_rt.set_fused(2)
            lat["nrx_rt"] = latency_percentiles(eng_rt, y1, a1)
            lat["plan"] = 2
            # the metric also names nrx_rt: its device-resident throughput on the same 30-slot batches
            eng_rt.set_fused(args.fused)
            outs_rt = {}
            for i in range(3):
                eng_rt.forward(ys[i % NBUF], act, want=want, out=outs_rt)
            torch.cuda.synchronize()
            r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n_rt = max(args.steps // 2, 5)
            r0.record()
            for i in range(n_rt):
                eng_rt.forward(ys[i % NBUF], act, want=want, out=outs_rt)
            r1.record()
            torch.cuda.synchronize()
            extra["nrx_rt_slots_per_s_one_gpu"] = B * n_rt / (r0.elapsed_time(r1) * 1e-3)
            eng_rt.close()
            extra["latency_us"] = lat
        # inactive-user skipping: the same 30-slot steps with one of the two users switched off
        # (the reference's 1-UE evaluation points, results/nrx_rt_results key (..., 1, 0)), everything computed as the
        # reference does vs. only the active planes
        act_half = act.clone()
        act_half[:, 1] = 0
        eng.set_fused(args.fused)

        def rate(n=max(args.steps // 2, 5)):
            for i in range(3):
                eng.forward(ys[i % NBUF], act_half, want=want, out=outs)
            torch.cuda.synchronize()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for i in range(n):
                eng.forward(ys[i % NBUF], act_half, want=want, out=outs)
            q1.record()
            torch.cuda.synchronize()
            return B * n / (q0.elapsed_time(q1) * 1e-3)
        if single:
            v_all = rate()
            eng.set_skip_inactive(True)
            v_skip = rate()
            eng.set_skip_inactive(False)
            extra["one_of_two_ues_active"] = {"all_planes_computed": v_all, "inactive_planes_skipped": v_skip, "unit": UNIT + " (one GPU)"}
        cpu = None
        if single and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            run = oracle_forward_timer(cfg, weights, grid, cores)
            chunk, _, _ = pick_cpu_mode(run, base)
            n_cpu = 3 * B                   # ~10 s of CPU work on the box's host cores
            dt_cpu = sum(run(base.y, base.active_tx, chunk) for _ in range(3))
            cpu = {"value": n_cpu / dt_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"3 whole steps ({n_cpu} slots) of the same batch through oracle/nrx_oracle.py (PyTorch-CPU fp32, "
                             f"{'slot by slot' if chunk == 1 else f'{chunk} slots per call'}), {dt_cpu:.1f} s"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16", "data": f"synthetic ({wsrc})",
            "config": {"workload": workload_desc(cfg),
                       "slots_per_pass": args.slots_per_pass or B, "plan": args.fused,
                       "l2": f"inputs rotate over {NBUF} distinct batches ({NBUF * ys_host[0].nbytes / 1e6:.0f} MB > 126 MB L2)",
                       "parallelism": f"slot-sharded x{world}, no data-path collective"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "whole_path": whole, "sustained": sustained,
            "kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
            "cpu_baseline": cpu, "clocks": clk.summary(), "slots_processed": slots_done,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
