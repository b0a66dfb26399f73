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
