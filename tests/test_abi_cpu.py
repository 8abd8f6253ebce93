"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/tb_knarpe.h declares;
argument validation returns the documented error codes without touching a GPU."""
import os
import re

import pytest

from trafficbotsv1_5_b200 import lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def handle():
    lib.build()
    return lib.load()


def test_exports_match_header(handle):
    hdr = open(os.path.join(ROOT, "include", "tb_knarpe.h")).read()
    declared = set(re.findall(r"^(?:int|const char\*)\s+(tb_\w+)\s*\(", hdr, flags=re.M))
    assert declared == set(lib.EXPORTS), declared ^ set(lib.EXPORTS)
    for name in declared:
        assert hasattr(handle, name), name
    assert handle.tb_version() >= 100


def test_error_codes_without_gpu(handle):
    # NULL pointers / bad shapes are rejected before any launch
    assert handle.tb_knn_select(None, None, None, None, 1, 1, 4, 1, 2, 1.0, None, None, None, 2, 0, None, None, 0, None) == -5
    one = 16  # any non-null, 16-byte aligned fake address; never dereferenced because validation fails first
    assert handle.tb_knn_select(one, one, one, one, 1, 1, 4, 1, 4, 1.0, one, one, one, 4, 0, None, None, 0, None) == -2  # K == T
    assert handle.tb_knn_select(one, one, one, one, 1, 1, 4096, 1, 4, 1.0, one, one, one, 4, 0, None, None, 0, None) == -3
    assert handle.tb_layernorm(one, 128, one, one, one, 128, 4, 96, 0, None) == -3
    assert handle.tb_layernorm(one, 128, one, one, one, 132, 4, 128, 2, None) == -4  # fp16 rows: ldy % 8
    assert handle.tb_linear(one, 4, one, None, 0, one, 4, 0, 4, 4, 0, None, None, 0, None, 0, None, 0, 0, None) == -1
    assert handle.tb_linear(one, 4, one, None, 0, one, 32, 8, 64, 4, 0, None, None, 0, None, 0, one, 64, 32, None) == -3  # fp16 out needs precision 1
    assert handle.tb_linear(one, 4, one, None, 0, one, 4, 8, 64, 4, 0, None, None, 0, None, 1, one, 64, 16, None) == -1  # col_h % 32
    assert handle.tb_strerror(-2).decode() == "need 0 < K < T"
