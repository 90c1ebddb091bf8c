"""The cases of the reference's own tests of the boundary this library replaces - xtp/src/tests/test_cudapipeline.cc
(CudaPipeline::gemm with blocks / transposes / beta, diag_gemm, axpy; 1e-9 isApprox) and test_cudamatrix.cc (upload /
block / download) - through the C ABI entry points that stand in for them: gwbse_dgemm_dev (cudapipeline.cc:53-105),
gwbse_diag_scale_dev (cudapipeline.h:133-162), gwbse_axpy_dev (cudapipeline.cc:36-51), gwbse_h2d / gwbse_d2h
(cudamatrix.cc:60-95).  Blocks are expressed as the reference's CudaMatrixBlock does: pointer offset + leading dimension."""
import ctypes

import numpy as np
import pytest

from tests.helpers import rel_frob


pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.fixture(scope="module")
def ctx():
    from votca_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def off(p, doubles):
    return ctypes.c_void_p(p.value + 8 * doubles)


def rnd(rng, r, c):
    return np.asfortranarray(rng.uniform(-1.0, 1.0, (r, c)))  # Eigen::MatrixXd::Random


def gemm(ctx, ta, tb, m, n, k, A, lda, B, ldb, beta, C, ldc):
    ctx.call("gwbse_dgemm_dev", ta.encode(), tb.encode(), m, n, k, 1.0, A, lda, B, ldb, float(beta), C, ldc)


def test_matmul_and_blocks(ctx):
    rng = np.random.default_rng(0)
    A, B = rnd(rng, 6, 10), rnd(rng, 10, 6)
    dA, dB = ctx.upload(A), ctx.upload(B)
    # matmul
    dC = ctx.upload(np.zeros((6, 6)))
    gemm(ctx, "N", "N", 6, 6, 10, dA, 6, dB, 10, 0.0, dC, 6)
    assert rel_frob(A @ B, ctx.download(dC, (6, 6))) < TOL
    # matmul_Cb: result into the block (1, 2, 6, 6) of an 8 x 10 matrix
    C8 = np.zeros((8, 10), order="F")
    dC8 = ctx.upload(C8)
    gemm(ctx, "N", "N", 6, 6, 10, dA, 6, dB, 10, 0.0, off(dC8, 2 * 8 + 1), 8)
    got = ctx.download(dC8, (8, 10))
    assert rel_frob(A @ B, got[1:7, 2:8]) < TOL
    got[1:7, 2:8] = 0.0
    assert not got.any()  # nothing outside the block is touched
    # matmul_ABb: B operand is the block (2, 3, 10, 6) of a 15 x 10 matrix
    B15 = rnd(rng, 15, 10)
    dB15 = ctx.upload(B15)
    gemm(ctx, "N", "N", 6, 6, 10, dA, 6, off(dB15, 3 * 15 + 2), 15, 0.0, dC, 6)
    assert rel_frob(A @ B15[2:12, 3:9], ctx.download(dC, (6, 6))) < TOL
    # matmul_AbB: A operand is the block (2, 3, 6, 10) of a 10 x 15 matrix
    A15 = rnd(rng, 10, 15)
    dA15 = ctx.upload(A15)
    gemm(ctx, "N", "N", 6, 6, 10, off(dA15, 3 * 10 + 2), 10, dB, 10, 0.0, dC, 6)
    assert rel_frob(A15[2:8, 3:13] @ B, ctx.download(dC, (6, 6))) < TOL
    for p in (dA, dB, dC, dC8, dB15, dA15):
        ctx.free(p)


def test_matmul_add_and_transposes(ctx):
    rng = np.random.default_rng(1)
    A, B, C = rnd(rng, 6, 10), rnd(rng, 10, 6), rnd(rng, 6, 6)
    dA, dB, dC = ctx.upload(A), ctx.upload(B), ctx.upload(C)
    gemm(ctx, "N", "N", 6, 6, 10, dA, 6, dB, 10, 2.0, dC, 6)  # matmul_add: A B + 2 C
    assert rel_frob(A @ B + 2 * C, ctx.download(dC, (6, 6))) < TOL
    C9 = rnd(rng, 9, 9)  # matmul_AB_Cbadd: block (1, 1, 6, 6) += A B
    dC9 = ctx.upload(C9)
    gemm(ctx, "N", "N", 6, 6, 10, dA, 6, dB, 10, 1.0, off(dC9, 9 + 1), 9)
    ref = C9.copy()
    ref[1:7, 1:7] += A @ B
    assert rel_frob(ref, ctx.download(dC9, (9, 9))) < TOL
    Bt = rnd(rng, 6, 10)  # matmul_ABt
    dBt = ctx.upload(Bt)
    gemm(ctx, "N", "T", 6, 6, 10, dA, 6, dBt, 6, 0.0, dC, 6)
    assert rel_frob(A @ Bt.T, ctx.download(dC, (6, 6))) < TOL
    At = rnd(rng, 10, 6)  # matmul_AtB
    dAt = ctx.upload(At)
    gemm(ctx, "T", "N", 6, 6, 10, dAt, 10, dB, 10, 0.0, dC, 6)
    assert rel_frob(At.T @ B, ctx.download(dC, (6, 6))) < TOL
    gemm(ctx, "T", "T", 6, 6, 10, dAt, 10, dBt, 6, 0.0, dC, 6)  # matmul_AtBt
    assert rel_frob(At.T @ Bt.T, ctx.download(dC, (6, 6))) < TOL
    for p in (dA, dB, dC, dC9, dBt, dAt):
        ctx.free(p)


def test_diag_matrix_mul_and_axpy(ctx):
    rng = np.random.default_rng(2)
    A, b = rnd(rng, 6, 10), rng.uniform(-1.0, 1.0, 10)
    dA, db, dC = ctx.upload(A), ctx.upload(b), ctx.upload(np.zeros((6, 10)))
    ctx.call("gwbse_diag_scale_dev", b"R", 6, 10, dA, 6, db, dC, 6)  # diag_matrix_mul: A diag(b)
    assert rel_frob(A * b[None, :], ctx.download(dC, (6, 10))) < TOL
    A2 = rnd(rng, 10, 6)
    dA2, dC2 = ctx.upload(A2), ctx.upload(np.zeros((10, 6)))
    ctx.call("gwbse_diag_scale_dev", b"L", 10, 6, dA2, 10, db, dC2, 10)  # diag_matrix_mulT: diag(b) A
    assert rel_frob(b[:, None] * A2, ctx.download(dC2, (10, 6))) < TOL
    ctx.call("gwbse_diag_scale_dev", b"L", 10, 6, dA2, 10, db, dA2, 10)  # diag_matrix_mul_onemat: in place
    assert rel_frob(b[:, None] * A2, ctx.download(dA2, (10, 6))) < TOL
    X, Y = rnd(rng, 8, 10), rnd(rng, 8, 10)  # axpy: B + 3 A
    dX, dY = ctx.upload(X), ctx.upload(Y)
    ctx.call("gwbse_axpy_dev", 8, 10, 3.0, dX, 8, dY, 8)
    assert rel_frob(Y + 3.0 * X, ctx.download(dY, (8, 10))) < TOL
    for p in (dA, db, dC, dA2, dC2, dX, dY):
        ctx.free(p)


def test_cudamatrix_roundtrip_and_block(ctx):
    rng = np.random.default_rng(3)
    X = rnd(rng, 10, 8)  # create_cudamatrix: upload / download is the identity
    d = ctx.upload(X)
    assert np.array_equal(ctx.download(d, (10, 8)), X)
    # create_cudamatrixblock: block (2, 3, 4, 5) = pointer offset + leading dimension; copied out with a 1 x block GEMM-free path
    blk = ctx.upload(np.zeros((4, 5)))
    ctx.call("gwbse_axpy_dev", 4, 5, 1.0, off(d, 3 * 10 + 2), 10, blk, 4)
    assert np.array_equal(ctx.download(blk, (4, 5)), X[2:6, 3:8])
    ctx.free(d)
    ctx.free(blk)
