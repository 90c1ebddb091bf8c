"""Index algebra of the materialised BSE direct-term blocks (votca_b200/csrc/capi_bse.cu: dense_build, dense_apply)
on the CPU, for one, two and three ranks.

scratch/check_dense_bse_index.py emulates the generalised operand / C addressing of the GEMM (gemm_dmma.cuh) and the
pack kernel on flat NumPy buffers with the parameter values the C++ sets - m-cyclic Mmn shards, the gathered vv / cv
blocks at their padded pitch, the chunked build, the scatter of a rank's result rows into Y - and compares the sum of
the ranks' contributions with the defining sums of Hd and Hd2.  The device-side counterparts are
tests/test_gpu_kernels.py::test_bse_operator_materialised_blocks (one GPU) and tests/test_gpu_multi.py (two)."""
import importlib.util
import os

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    spec = importlib.util.spec_from_file_location("check_dense_bse_index",
                                                  os.path.join(os.path.dirname(HERE), "scratch", "check_dense_bse_index.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("case", [(7, 12, 14, 0, 4, 5, 3, 100), (6, 13, 13, 1, 3, 7, 2, 2), (5, 9, 11, 2, 3, 4, 4, 1),
                                  (4, 20, 21, 3, 5, 6, 1, 3), (3, 33, 34, 1, 16, 16, 2, 5)])
def test_block_build_and_product_index_algebra(world, case):
    naux, mtotal, ntotal, voff, vt, ct, k, lchunk = case
    assert _load().run(world, naux, mtotal, ntotal, voff, vt, ct, k, lchunk, seed=world)


@pytest.mark.parametrize("world", [1, 2, 3])
@pytest.mark.parametrize("case", [(5, 7, 19, 4, 11), (6, 8, 17, 3, 17), (4, 5, 16, 0, 9)])
def test_windowed_multiply_right_index_algebra(world, case):
    """gwbse_mmn_mul_right_window_dev / mmn_complete_rotation (capi_mmn.cu rotate_rows): the window rows first, in
    place through the second buffer, then the rest; device-side: test_mul_right_on_a_row_window_then_the_rest."""
    naux, mtotal, ntotal, n_lo, n_hi = case
    assert _load().run_window_rotation(world, naux, mtotal, ntotal, n_lo, n_hi, seed=world)
