"""Generate the committed golden fixtures: oracle (exact fp32) outputs on seeded synthetic slots.

    python tests/golden/make_golden.py

* nrx_rt_random_4prb.npz   — nrx_rt architecture, seeded random weights (reproducible anywhere)
* nrx_rt_shipped_4prb.npz  — same slots with the reference's weights/nrx_rt_weights (outputs only;
                             the weight file itself is not redistributed)
* nrx_large_shipped_2prb.npz — nrx_large (8 iterations) with weights/nrx_large_weights
The reference ships no LLR-level golden vectors: these pin the oracle against
accidental change and give the GPU tests a fixed target that does not need the oracle at run time.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.synth import make_slots  # noqa: E402
from neural_rx_b200.weights import load_weights, random_weights  # noqa: E402
from oracle import nrx_oracle as O  # noqa: E402
from tests.common import oracle_arch, oracle_net, weight_path  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(label, n_prb, weights, out_name, weight_seed=-1, batch=2, ebno=8.0, seed=2024):
    cfg = get_config(label)
    grid = build_grid(cfg, n_size_bwp=n_prb)
    sb = make_slots(cfg, grid, batch=batch, ebno_db=ebno, seed=seed)
    ref = O.receiver_forward(oracle_net(cfg, weights), oracle_arch(cfg), sb.y, grid.pilots, grid.pilot_mask,
                             sb.active_tx)
    np.savez_compressed(os.path.join(HERE, out_name), y=sb.y, active_tx=sb.active_tx, bits=sb.bits,
                        llr=ref["llr"].astype(np.float32), h_hat=ref["h_hat"].astype(np.float32),
                        h_hat_refined=ref["h_hat_refined"].astype(np.float32),
                        weight_seed=np.int64(weight_seed), n_prb=np.int64(n_prb))
    print(out_name, "written")


if __name__ == "__main__":
    cfg = get_config("nrx_rt")
    make("nrx_rt", 4, random_weights(cfg, seed=123), "nrx_rt_random_4prb.npz", weight_seed=123)
    p = weight_path("nrx_rt")
    if p:
        make("nrx_rt", 4, load_weights(cfg, p), "nrx_rt_shipped_4prb.npz")
    p = weight_path("nrx_large")
    if p:
        make("nrx_large", 2, load_weights(get_config("nrx_large"), p), "nrx_large_shipped_2prb.npz")
