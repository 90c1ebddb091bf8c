"""Pins the oracle's DavidsonSolver and HamiltonianOperator on the reference's own unit tests
(xtp/src/tests/test_davidson.cc:59-393, test_bseoperator_btda.cc:54-96): lowest eigenvalues of the test
matrices against dense LAPACK at 1e-6, eigenvector weights at 1e-3, failure reporting after one iteration."""
import numpy as np
import pytest

from oracle.bse_operator import HamiltonianOperator
from oracle.davidson import DavidsonSolver
from tests.helpers import rel_frob


def init_matrix(n, eps):  # test_davidson.cc:42-55
    i, j = np.indices((n, n))
    with np.errstate(divide="ignore"):
        A = eps / (j - i).astype(float) ** 2
    A[np.diag_indices(n)] = np.sqrt(1.0 + np.arange(n))
    return A


def symm_matrix(n, eps, seed):  # test_davidson.cc:34-40 (Eigen::MatrixXd::Random is uniform on [-1, 1])
    R = eps * np.random.default_rng(seed).uniform(-1.0, 1.0, (n, n))
    return R + R.T


class Dense:
    def __init__(self, A, matrix_free=False):
        self.A, self.matrix_free = A, matrix_free

    def rows(self):
        return self.A.shape[0]

    def diagonal(self):
        return np.diag(self.A).copy()

    def matmul(self, X):
        if self.matrix_free:  # TestOperator::matmul builds the product row by row (test_davidson.cc:129-136)
            return np.stack([self.A[i] @ X for i in range(self.A.shape[0])])
        return self.A @ X

    matmul_factorised = matmul


@pytest.mark.parametrize("size,matrix_free", [(100, False), (400, False), (100, True), (400, True)])
def test_davidson_symmetric(size, matrix_free):
    A = init_matrix(size, 0.01)
    ds = DavidsonSolver()
    if matrix_free:
        ds.set_tolerance("normal")
        ds.set_size_update("safe")
    lam, vec = ds.solve(Dense(A, matrix_free), 10)
    ref = np.linalg.eigvalsh(A)[:10]
    assert ds.info == "Success" and rel_frob(ref, lam) < 1e-6
    assert np.abs(A @ vec - vec * lam).max() < 1e-3


def test_davidson_reports_failure_after_one_iteration():  # test_davidson.cc:107-125
    A = init_matrix(100, 0.01)
    ds = DavidsonSolver()
    ds.set_iter_max(1)
    lam, _ = ds.solve(Dense(A), 10)
    assert ds.info == "NoConvergence"
    assert not rel_frob(np.linalg.eigvalsh(A)[:10], lam + 1e-300) < 1e-6


def _positive_branch(H, neigen):  # index_eval + extract_eigenvectors, test_davidson.cc:232-262
    w, V = np.linalg.eig(H)
    w, V = w.real, V.real
    idx = [i for i in np.argsort(w, kind="stable") if w[i] > 0][:neigen]
    return w[idx], V[:, idx]


@pytest.mark.parametrize("size,update,space", [(60, "max", 0), (120, "safe", 50)])
def test_davidson_hamiltonian(size, update, space):
    R, C = Dense(init_matrix(size, 0.01)), Dense(symm_matrix(size, 0.01, 5))
    H = HamiltonianOperator(R, C)
    ds = DavidsonSolver()
    ds.set_tolerance("normal")
    ds.set_size_update(update)
    if space:
        ds.set_max_search_space(space)
    ds.set_matrix_type("HAM")
    lam, vec = ds.solve(H, 5)
    dense = H.matmul(np.eye(2 * size))
    ref_w, ref_v = _positive_branch(dense, 5)
    assert ds.info == "Success" and rel_frob(ref_w, np.sort(lam)) < 1e-6
    # eigenvector weights (cwiseAbs2) after a common normalisation
    a = ref_v / np.linalg.norm(ref_v, axis=0)
    b = vec / np.linalg.norm(vec, axis=0)
    assert rel_frob(a ** 2, b ** 2) < 1e-3


def test_hamiltonian_operator_matches_dense():  # test_bseoperator_btda.cc:54-96
    n = 60
    A, B = init_matrix(n, 0.01), symm_matrix(n, 0.01, 6)
    H = HamiltonianOperator(Dense(A), Dense(B))
    dense = np.block([[A, B], [-B, -A]])
    assert rel_frob(dense, H.matmul(np.eye(2 * n))) < 1e-9
    x = np.random.default_rng(7).standard_normal((2 * n, 3))
    assert rel_frob(dense @ x, H.matmul(x)) < 1e-12
    assert np.array_equal(H.diagonal(), np.concatenate([np.diag(A), -np.diag(A)]))
