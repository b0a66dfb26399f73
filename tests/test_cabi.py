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
    """Host-only planning helpers: how the stack kernels cut the subcarrier axis over 148 CTAs.
    Plans 1-3: chunks of a plane are a partition of [0, F) (c_j = j*F/n) and fill the machine;
    plan 4: jobs (eight per item, advanced in lock step) cover every subcarrier once, the step count
    is the longest job plus the six pipeline-fill steps, and the chosen split is never worse than
    one job per plane."""
    sms = 148
    n = ctypes.c_int32()
    assert lib.nrx_plan_stack_chunks(planes, F, sms, ctypes.byref(n)) == 0
    n = n.value
    assert 1 <= n <= max(F // 5, 1)
    edges = [j * F // n for j in range(n + 1)]
    assert edges[0] == 0 and edges[-1] == F and all(b > a for a, b in zip(edges, edges[1:]))

    j, items, steps = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    assert lib.nrx_plan_stack_jobs(planes, F, sms, ctypes.byref(j), ctypes.byref(items), ctypes.byref(steps)) == 0
    j, items, steps = j.value, items.value, steps.value
    assert 1 <= j <= F
    bounds = [i * F // j for i in range(j + 1)]
    lens = [b - a for a, b in zip(bounds, bounds[1:])]
    assert sum(lens) == F and min(lens) >= 1
    assert steps == max(lens) + 6 == -(-F // j) + 6
    assert items == -(-planes * j // 8)

    def cost(jj):
        it = -(-planes * jj // 8)
        return -(-it // sms) * (-(-F // jj) + 6)
    assert cost(j) <= cost(1) and cost(j) == min(cost(jj) for jj in range(1, min(F, 1024) + 1))
    assert lib.nrx_plan_stack_jobs(0, F, sms, None, None, None) == 1       # NRX_ERR_INVALID


def test_fragment_column_order(lib):
    """Plan 4 keeps output channel 16c + 4g + 2e + d in accumulator column 16c + 8e + 2g + d (the columns one
    thread of the 16x256b tcgen05 fragment owns are four consecutive channels): a permutation of every group of 16."""
    cols = [lib.nrx_fragment_column(n) for n in range(128)]
    assert sorted(cols) == list(range(128))
    for n in range(128):
        c, g, e, d = n // 16, (n % 16) // 4, (n % 4) // 2, n % 2
        assert cols[n] == 16 * c + 8 * e + 2 * g + d
    assert lib.nrx_fragment_column(-1) == -1
