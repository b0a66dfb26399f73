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


@pytest.mark.parametrize("planes,F,ctas", [(60, 1584, 148), (2, 1584, 148), (2, 48, 148), (7, 12, 148), (1, 3, 148), (120, 3276, 148), (59, 1584, 132)])
def test_stack_balanced_ranges_without_device(lib, planes, F, ctas):
    """Default work distribution of the fused stack kernels (nrx_plan_stack_range): the CTA ranges partition the
    planes x F line in order, empty ranges only at the end, no range shorter than 5 subcarriers unless the whole
    launch is, and the longest CTA walk (steps of 9 subcarriers, 4-subcarrier run-in per item) never exceeds what
    equal chunks per plane (nrx_plan_stack_chunks) cost."""
    W = planes * F
    a, b = ctypes.c_int64(), ctypes.c_int64()
    pos, worst, seen_empty = 0, 0, False
    for c in range(ctas):
        assert lib.nrx_plan_stack_range(planes, F, ctas, c, ctypes.byref(a), ctypes.byref(b)) == 0
        if a.value == b.value:
            seen_empty = True
            continue
        assert not seen_empty and a.value == pos and b.value > a.value
        assert b.value - a.value >= min(5, W)
        steps, g = 0, a.value
        while g < b.value:                                  # items: one per plane touched
            c0 = g % F
            c1 = min(F, c0 + b.value - g)
            steps += -(-(c1 - c0 + 4) // 9)
            g += c1 - c0
        worst = max(worst, steps)
        pos = b.value
    assert pos == W
    n = ctypes.c_int32()
    assert lib.nrx_plan_stack_chunks(planes, F, ctas, ctypes.byref(n)) == 0
    uniform = -(-planes * n.value // ctas) * -(-(-(-F // n.value) + 4) // 9)
    assert worst <= uniform
    if (planes, F, ctas) == (60, 1584, 148):
        assert uniform == 77 and worst <= 74
    assert lib.nrx_plan_stack_range(planes, F, ctas, ctas, ctypes.byref(a), ctypes.byref(b)) == 1     # NRX_ERR_INVALID
