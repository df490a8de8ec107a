"""CPU: the tables the product generates itself equal the reference's (no GPU needed)."""
import ctypes
import importlib

import numpy as np
import pytest

import oracle

pkg = importlib.import_module("x265-yuuki-asuna_b200")
needs_ref = pytest.mark.skipif(not oracle.have_ref(8), reason="oracle/_ref not built")


@needs_ref
def test_dct_tables_match_reference():
    R = oracle.ref(8)
    for i, N in enumerate([4, 8, 16, 32]):
        ptr = R.ref_dct_table(i)
        ref = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_int16)), shape=(N, N))
        assert np.array_equal(pkg.dct_table(N), ref), N


@needs_ref
@pytest.mark.parametrize("depth", [8, 10])
def test_lambda_and_bitcost_tables_match_reference(depth):
    R = oracle.ref(depth)
    for qp in range(0, 70, 3):
        assert pkg.lambda_for_qp(qp, depth) == R.ref_lambda(qp), qp
    for qp in (0, 17, 30, 37, 51, 63):
        exp = np.empty(4 * 32768 + 1, dtype=np.uint16)
        R.ref_bitcost_table(qp, ctypes.c_void_p(exp.ctypes.data))
        got = pkg.bitcost_table(R.ref_lambda(qp))
        assert np.array_equal(got, exp), qp


def test_dct_table_golden_rows():
    """pinned without the reference: first rows of the HEVC matrices (public standard values)."""
    t32 = pkg.dct_table(32)
    assert list(t32[1][:16]) == [90, 90, 88, 85, 82, 78, 73, 67, 61, 54, 46, 38, 31, 22, 13, 4]
    assert list(t32[16][:4]) == [64, -64, -64, 64]
    assert list(pkg.dct_table(4).ravel()) == [64, 64, 64, 64, 83, 36, -36, -83, 64, -64, -64, 64, 36, -83, 83, -36]
