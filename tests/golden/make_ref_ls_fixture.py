"""Pin the oracle's LS channel estimate (pilot gather order, safe division, nearest-pilot broadcast over the whole
grid — the Sionna-shaped entry of the receiver) against the REFERENCE'S OWN NumPy estimator, executed here.

    python tests/golden/make_ref_ls_fixture.py        (needs /root/reference; run in the build container)

``NeuralPUSCHReceiver.__init__`` builds ``MyLSChannelEstimatorNP(self.rg, interpolation_type="nn")``
(utils/neural_rx.py:1432) and ``estimate_channel`` calls it on the received grid.  The class
(utils/neural_rx.py:1129-1381) is plain NumPy: gather of the pilot REs (:1255-1264, :1354), safe division by the
pilots (:1289-1294), ``NearestNeighborInterpolator`` (:919-1004).  It is extracted with ``ast`` together with the
helpers it calls in the same file (``to_numpy`` :33, ``myexpand_to_rank`` :1084, ``RemoveNulledSubcarriers`` :884) —
nothing is copied into the repo — and run on a seeded 2-PRB slot with a stand-in for the resource-grid object
(pilot mask, pilots, effective subcarrier indices) and NumPy stand-ins for the two third-party calls on the path:
``sionna.utils.flatten_last_dims`` (reshape of the last two dimensions) and ``tf.gather(..., 2, batch_dims=2)``
(:1011; the gather INDICES are the reference's own and already pinned bit-exact by tests/golden/make_ref_fixtures.py).

What the fork's estimator does NOT do is the FOCC / CDM de-spreading of Sionna's ``PUSCHLSChannelEstimator``, which
the original receiver used and which the oracle applies between the division and the
interpolation; that step is pinned separately to the reference's ``NRPreprocessing._focc_removal``
(tests/golden/make_ref_fixtures.py).  The oracle is therefore called with ``focc=False`` for this fixture.
Inputs + outputs go to tests/golden/ref_ls_fixture.npz; tests/test_oracle_pins.py compares."""
import ast
import os
import sys
import tempfile
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

from make_ref_fixtures import quiet  # noqa: E402
from neural_rx_b200.config import get_config  # noqa: E402
from neural_rx_b200.pusch import build_grid  # noqa: E402
from neural_rx_b200.synth import make_slots  # noqa: E402


def reference_defs(path, names, extra_ns):
    """exec the named top-level classes / functions of a reference source file in a NumPy-only namespace."""
    tree = ast.parse(open(path).read())
    ns = {"np": np, "os": os}
    ns.update(extra_ns)
    for node in tree.body:
        if isinstance(node, (ast.ClassDef, ast.FunctionDef)) and node.name in names:
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    missing = [n for n in names if n not in ns]
    assert not missing, missing
    return ns


def flatten_last_dims(tensor, num_dims=2):
    """NumPy stand-in for sionna.utils.flatten_last_dims (documented semantics: merge the last `num_dims` axes)."""
    tensor = np.asarray(tensor)
    return tensor.reshape(tensor.shape[:-num_dims] + (-1,))


def _tf_gather(params, indices, axis, batch_dims):
    """NumPy stand-in for the one TensorFlow call on the path, ``tf.gather(inputs, gather_ind, 2, batch_dims=2)``
    (utils/neural_rx.py:1011), with TF's documented semantics: the first two axes of `params` and `indices` are batch
    axes, the gather runs along axis 2 of `params` with the remaining axes of `indices` taking its place."""
    assert axis == 2 and batch_dims == 2
    params, indices = np.asarray(params), np.asarray(indices)
    out = np.empty(params.shape[:2] + indices.shape[2:] + params.shape[3:], params.dtype)
    for a in range(params.shape[0]):
        for b in range(params.shape[1]):
            out[a, b] = np.take(params[a, b], indices[a, b], axis=0)
    return out


def main():
    ns = reference_defs(os.path.join(REF, "utils", "neural_rx.py"),
                        ["to_numpy", "myexpand_to_rank", "RemoveNulledSubcarriers", "NearestNeighborInterpolator",
                         "MyLSChannelEstimatorNP"], {"flatten_last_dims": flatten_last_dims, "tf": types.SimpleNamespace(gather=_tf_gather)})
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=2)
    U, T, Fs = grid.num_tx, grid.num_ofdm_symbols, grid.num_subcarriers
    pp = types.SimpleNamespace(num_pilot_symbols=grid.pilots.shape[1],
                               mask=np.broadcast_to(grid.pilot_mask[None, None], (U, 1, T, Fs)).copy(),
                               pilots=grid.pilots[:, None, :])
    rg = types.SimpleNamespace(pilot_pattern=pp, effective_subcarrier_ind=range(Fs))
    sb = make_slots(cfg, grid, batch=3, ebno_db=[3.0, 9.0, 15.0], seed=41, active=np.array([[1, 1], [1, 0], [1, 1]], np.float32))
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:                                # the interpolator writes data/*.npy
        os.chdir(tmp)
        try:
            est = quiet(ns["MyLSChannelEstimatorNP"], rg, interpolation_type="nn")
            h_hat, err_var = quiet(est, [sb.y, np.float32(0.1)])
        finally:
            os.chdir(cwd)
    h_hat = np.asarray(h_hat)
    assert h_hat.shape == (3, 1, cfg.num_rx_antennas, U, 1, T, Fs), h_hat.shape
    np.savez_compressed(os.path.join(HERE, "ref_ls_fixture.npz"), y=sb.y, pilots=grid.pilots, pilot_mask=grid.pilot_mask,
                        h_hat=h_hat.astype(np.complex64), n_prb=np.int64(2))
    print("ref_ls_fixture.npz:", h_hat.shape, float(np.abs(h_hat).mean()))


if __name__ == "__main__":
    main()
