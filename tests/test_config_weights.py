"""CPU tests of the host logic: cfg reader (mirrors utils/parameters.py:91-127), presets,
weight-list walker (utils/utils.py:34-70), receiver argument handling."""
import os

import numpy as np
import pytest

from neural_rx_b200.config import PRESETS, get_config, load_cfg, parse_cfg_text
from neural_rx_b200.weights import from_list, load_weights, random_weights, save_weights

REF_CFG = "/root/reference/config"


@pytest.mark.parametrize("label", sorted(PRESETS))
def test_presets_equal_reference_cfg(label):
    path = os.path.join(REF_CFG, label + ".cfg")
    if not os.path.exists(path):
        pytest.skip("/root/reference not present")
    assert load_cfg(path) == PRESETS[label]


def test_cfg_eval_rules():
    text = """
[global]
label = 'demo'
[system]
n_size_bwp = 4
num_rx_antennas = 4
mcs_index = [9, 14]
dmrs_port_sets = [[0], [2]]
dmrs_additional_position = 1
[neural_receiver]
num_nrx_iter = 2
d_s = 56
num_units_init = [128, 128]
num_units_agg = [[64], [64]]
num_units_state = [[128, 128], [128, 128]]
num_units_readout = [128]
max_num_tx = 2
nrx_dtype = tf.float32
[evaluation]
n_size_bwp_eval = 132
"""
    cfg = parse_cfg_text(text)                     # inference: *_eval overrides (parameters.py:118-127)
    assert cfg.n_size_bwp == 132 and cfg.num_subcarriers == 1584 and cfg.nrx_dtype == "float32"
    assert parse_cfg_text(text, training=True).n_size_bwp == 4
    assert cfg.num_bits_per_symbol == [2, 4] and cfg.dmrs_symbols == (2, 11)
    assert parse_cfg_text(text.replace("tf.float32", "torch.float32")).nrx_dtype == "float32"   # config/nrx_rt.cfg:75
    with pytest.raises(FileNotFoundError, match="Unknown config file"):
        get_config("does_not_exist.cfg")


def test_validate_errors_match_reference():
    import dataclasses
    cfg = get_config("nrx_rt")
    with pytest.raises(NotImplementedError, match="Unknown layer_type selected"):
        dataclasses.replace(cfg, layer_type_conv="conv").validate()
    with pytest.raises(ValueError, match="Cannot use initial channel estimator if pilots are masked"):
        dataclasses.replace(cfg, mask_pilots=True).validate()
    with pytest.raises(ValueError, match="Invalid number of iterations"):
        dataclasses.replace(cfg, num_nrx_iter_eval=3).validate()


def test_weight_list_roundtrip_and_shape_errors(tmp_path):
    cfg = get_config("nrx_rt_var_mcs")
    w = random_weights(cfg, seed=2)
    lst = w.to_list()
    assert len(lst) == 56 and lst[0].shape == (3, 3, 18, 1) and lst[1].shape == (1, 1, 18, 128)
    p = tmp_path / "w"
    save_weights(w, str(p))
    w2 = load_weights(cfg, str(p))
    assert all(np.array_equal(a, b) for a, b in zip(lst, w2.to_list()))
    with pytest.raises(ValueError):
        from_list(get_config("nrx_rt"), lst)              # 56 arrays into a 43-array architecture
    bad = [a.copy() for a in lst]
    bad[1] = bad[1][..., :64]
    with pytest.raises(ValueError, match="SeparableConv2D"):
        from_list(cfg, bad)


def test_receiver_requires_cuda_library_or_device():
    """No CPU fallback: without a GPU the engine fails loudly (NRX_ERR_CUDA), with one it builds."""
    import torch
    from neural_rx_b200.engine import NrxError
    from neural_rx_b200.pusch import build_grid
    from neural_rx_b200.receiver import NeuralPUSCHReceiver
    cfg = get_config("nrx_rt")
    grid = build_grid(cfg, n_size_bwp=1)
    if torch.cuda.is_available():
        rx = NeuralPUSCHReceiver(cfg, grid=grid)
        with pytest.raises(AssertionError, match="Invalid number of iterations"):
            rx.num_it = 5
    else:
        with pytest.raises(NrxError):
            NeuralPUSCHReceiver(cfg, grid=grid)
    with pytest.raises(NotImplementedError):
        NeuralPUSCHReceiver(cfg, training=True, grid=grid)


def test_weight_file_loader_accepts_arrays_only(tmp_path):
    """The reference unpickles weight files without restriction (utils/utils.py:53-70); the loader here resolves
    only the globals NumPy's array pickles need, so a file that asks for anything else is refused."""
    import pickle

    class Evil:
        def __reduce__(self):
            return (print, ("executed",))

    p = tmp_path / "evil_weights"
    p.write_bytes(pickle.dumps([Evil()]))
    with pytest.raises(pickle.UnpicklingError, match="only NumPy arrays"):
        load_weights(get_config("nrx_rt"), str(p))
    w = random_weights(get_config("nrx_rt"), seed=2)
    q = tmp_path / "ok_weights"
    q.write_bytes(pickle.dumps(w.to_list()))
    assert load_weights(get_config("nrx_rt"), str(q)).num_params() == w.num_params()


def test_cfg_values_are_literals_not_code():
    """The reference eval()s every cfg value (utils/parameters.py:104-110); the parser here accepts Python literals
    and the dtype names the cfg files use, nothing executable."""
    from neural_rx_b200.config import _literal
    assert _literal("[128, 128]") == [128, 128] and _literal("tf.float32") == "float32" and _literal("torch.float32") == "float32"
    with pytest.raises(ValueError, match="not a Python literal"):
        _literal("__import__('os').system('true')")
