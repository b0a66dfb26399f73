"""CPU test: libnrx_b200.so loads and exports every symbol include/nrx_b200.h declares; argument
validation that needs no device (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from neural_rx_b200.build import build_library
    from neural_rx_b200.engine import load_library
    build_library()
    return load_library()


def _declared():
    text = open(os.path.join(ROOT, "include", "nrx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nrx_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported(lib):
    from neural_rx_b200.engine import EXPORTED_SYMBOLS
    declared = _declared()
    assert declared == sorted(EXPORTED_SYMBOLS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert b"sm_100a" in lib.nrx_version()


def test_struct_layout_matches_header():
    from neural_rx_b200.engine import ModelDesc
    # 6 + 2 + 1 + 2 + 1 + 1 + 4 + 1 + 4 + 2 int32 fields
    assert ctypes.sizeof(ModelDesc) == 4 * 24


def test_argument_validation_without_device(lib):
    assert lib.nrx_create(None, None, None, 0, None, None, None, None, 0, None) == 1     # NRX_ERR_INVALID
    assert b"null argument" in lib.nrx_last_error()
    assert lib.nrx_set_num_it(None, 1) == 1
    assert lib.nrx_destroy(None) == 0


@pytest.mark.parametrize("planes,F", [(60, 1584), (2, 1584), (2, 48), (7, 12), (14, 132), (120, 3276)])
def test_stack_work_split_without_device(lib, planes, F):
    """Host-only planning helper: how the stack kernels cut the subcarrier axis over 148 CTAs: the chunks of a
    plane are a partition of [0, F) (c_j = j*F/n), every chunk has at least 5 subcarriers and the chosen count
    minimises waves x steps-per-item (9-subcarrier steps, 4-subcarrier run-in)."""
    sms = 148
    n = ctypes.c_int32()
    assert lib.nrx_plan_stack_chunks(planes, F, sms, ctypes.byref(n)) == 0
    n = n.value
    assert 1 <= n <= max(F // 5, 1)
    edges = [j * F // n for j in range(n + 1)]
    assert edges[0] == 0 and edges[-1] == F and all(b > a for a, b in zip(edges, edges[1:]))

    def cost(nn):
        steps = -(-(-(-F // nn) + 4) // 9)
        return -(-planes * nn // sms) * steps
    assert cost(n) == min(cost(nn) for nn in range(1, min(max(F // 5, 1), 512) + 1))
    assert lib.nrx_plan_stack_chunks(0, F, sms, None) == 1                 # NRX_ERR_INVALID
