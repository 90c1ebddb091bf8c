"""Oracle for DavidsonSolver (test infrastructure).

Follows xtp/include/votca/xtp/davidsonsolver.h:64-280 and
xtp/src/libxtp/davidsonsolver.cc:149-555 step by step (Ritz / harmonic Ritz,
DPR / Olsen correction, twice-applied Gram-Schmidt, restart).
Small dense eigenproblems go to LAPACK through NumPy/SciPy.
"""
import numpy as np
import scipy.linalg


class DavidsonSolver:
    def __init__(self):
        self.iter_max = 50
        self.tol = 1e-4
        self.max_search_space = 0
        self.correction = "DPR"
        self.update = "SAFE"
        self.matrix_type = "SYMM"
        self.info = "NoConvergence"
        self.i_iter = 0
        self.n_matmul_cols = 0

    def set_iter_max(self, n):
        self.iter_max = n

    def set_max_search_space(self, n):
        self.max_search_space = n

    def set_tolerance(self, tol):
        self.tol = {"loose": 1e-3, "normal": 1e-4, "strict": 1e-5, "lapack": 1e-9}[tol]

    def set_correction(self, m):
        if m not in ("DPR", "OLSEN"):
            raise RuntimeError(m + " is not a valid Davidson correction method")
        self.correction = m

    def set_size_update(self, u):
        if u not in ("min", "safe", "max"):
            raise RuntimeError(u + " is not a valid Davidson update")
        self.update = u.upper()

    def set_matrix_type(self, mt):
        if mt not in ("HAM", "SYMM"):
            raise RuntimeError(mt + " is not a valid Davidson matrix type")
        self.matrix_type = mt

    def _size_update(self, neigen):
        if self.update == "MIN":
            return neigen
        if self.update == "SAFE":
            return int(1.5 * neigen) if neigen < 20 else neigen + 10
        return 2 * neigen

    def _apply(self, A, X):
        self.n_matmul_cols += X.shape[1]
        return A.matmul(X)

    # davidsonsolver.h:64-203
    def solve(self, A, neigen, initial_guess=None):
        if self.max_search_space < neigen:
            self.max_search_space = neigen * 5
        op_size = A.rows()
        if self.max_search_space > op_size:
            self.max_search_space = op_size
        self.Adiag = np.asarray(A.diagonal())
        if initial_guess is None:
            size_initial_guess = 2 * neigen
            self.restart_size = size_initial_guess
            V = self._initial_vectors(size_initial_guess)
        else:
            if initial_guess.shape[0] != op_size:
                raise RuntimeError("DavidsonSolver::solve initial_guess has wrong number of rows.")
            if initial_guess.shape[1] < neigen:
                raise RuntimeError("DavidsonSolver::solve initial_guess has fewer columns than neigen.")
            self.restart_size = min(initial_guess.shape[1], self.max_search_space)
            V = np.array(initial_guess, dtype=np.float64)
            self._gramschmidt(V, 0)
        self.size_update = self._size_update(neigen)
        root_converged = np.zeros(self.size_update, dtype=bool)
        proj = {"V": V}
        self.history = []
        for self.i_iter in range(self.iter_max):
            self._update_projection(A, proj)
            rep = self._ritz(proj) if self.matrix_type == "SYMM" else self._harmonic_ritz(proj)
            res_norm = np.linalg.norm(rep["res"], axis=0)
            root_converged = res_norm[:self.size_update] < self.tol
            converged = bool(np.all(root_converged[:neigen]))
            self.history.append((self.i_iter, proj["V"].shape[1], res_norm[:neigen].max()))
            last = self.i_iter == self.iter_max - 1
            if converged:
                self._store(rep, neigen)
                self.info = "Success"
                break
            elif last:
                self._store(rep, neigen)
                for i in range(neigen):
                    if not root_converged[i]:
                        self.eigenvalues[i] = 0
                        self.eigenvectors[:, i] = 0
                self.info = "NoConvergence"
                break
            ext = self._extend(rep, proj, root_converged)
            if proj["V"].shape[1] > self.max_search_space:
                self._restart(rep, proj, ext)
        return self.eigenvalues, self.eigenvectors

    def _initial_vectors(self, n):
        guess = np.zeros((len(self.Adiag), n))
        idx = np.argsort(self.Adiag, kind="stable")
        if self.matrix_type == "SYMM":
            for j in range(n):
                guess[idx[j], j] = 1.0
        else:
            ind0 = len(self.Adiag) // 2
            for j in range(n):
                guess[idx[ind0 + j], j] = 1.0
        return guess

    # davidsonsolver.h:239-280
    def _update_projection(self, A, proj):
        V = proj["V"]
        if self.i_iter == 0:
            proj["AV"] = self._apply(A, V)
            proj["T"] = V.T @ proj["AV"]
            if self.matrix_type == "HAM":
                proj["AAV"] = self._apply(A, proj["AV"])
                proj["B"] = V.T @ proj["AAV"]
            return
        old = proj["AV"].shape[1]
        new = V.shape[1]
        nvec = new - old
        AVn = self._apply(A, V[:, old:])
        proj["AV"] = np.hstack([proj["AV"], AVn])
        T = np.zeros((new, new))
        T[:old, :old] = proj["T"]
        T[:, old:] = V.T @ AVn
        if self.matrix_type == "SYMM":
            T[old:, :old] = T[:old, old:].T
        else:
            T[old:, :old] = V[:, old:].T @ proj["AV"][:, :old]
            AAVn = self._apply(A, AVn)
            proj["AAV"] = np.hstack([proj["AAV"], AAVn])
            B = np.zeros((new, new))
            B[:old, :old] = proj["B"]
            B[:, old:] = V.T @ AAVn
            B[old:, :old] = V[:, old:].T @ proj["AAV"][:, :old]
            proj["B"] = B
        proj["T"] = T

    def _needed(self, proj):
        return min(proj["T"].shape[1], max(self.restart_size, self.size_update))

    # davidsonsolver.cc:221-239
    def _ritz(self, proj):
        ev, U = np.linalg.eigh(proj["T"], UPLO="L")
        n = self._needed(proj)
        lam, U = ev[:n], U[:, :n]
        q = proj["V"] @ U
        res = proj["AV"] @ U - q * lam[None, :]
        return {"lambda": lam, "U": U, "q": q, "res": res}

    # davidsonsolver.cc:241-332
    def _harmonic_ritz(self, proj):
        w, vr = scipy.linalg.eig(proj["T"], proj["B"])
        pairs = []  # [first, second]
        for i in range(len(w)):
            if w[i].imag != 0:
                found = False
                for pr in pairs:
                    if pr[1] > -1:
                        continue
                    if abs(w[pr[0]].real - w[i].real) < 1e-9 and abs(w[pr[0]].imag + w[i].imag) < 1e-9:
                        pr[1] = i
                        found = True
                if not found:
                    pairs.append([i, -1])
        for pr in pairs:
            if pr[1] < 0:
                raise RuntimeError("Eigenvalue:" + str(pr[0]) + " is complex but has no partner.")
        seconds = {pr[1] for pr in pairs}
        vals, vecs = [], []
        for i in range(len(w)):
            if i in seconds:
                continue
            vals.append(w[i].real)
            v = vr[:, i].real.copy()
            v /= np.linalg.norm(v)
            vecs.append(v)
        vals = np.array(vals)
        vecs = np.array(vecs).T
        n = self._needed(proj)
        idx = np.argsort(vals, kind="stable")[::-1][:n]
        U = vecs[:, idx]
        lam = np.diag(U.T @ proj["T"] @ U).copy()
        q = proj["V"] @ U
        res = proj["AV"] @ U - q * lam[None, :]
        return {"lambda": lam, "U": U, "q": q, "res": res}

    def _dpr(self, r, lam):
        with np.errstate(divide="ignore", invalid="ignore"):
            return -r / (self.Adiag - lam)

    def _correction(self, q, lam, r):
        if self.correction == "DPR":
            c = self._dpr(r, lam)
        else:
            delta = self._dpr(r, lam)
            num = -(q @ delta)
            den = -(q @ self._dpr(q, lam))
            c = delta + (num / den) * q
        return np.where(np.isfinite(c), c, 0.0)

    # davidsonsolver.cc:354-378
    def _extend(self, rep, proj, root_converged):
        nupdate = int(np.count_nonzero(~root_converged))
        V = proj["V"]
        old = V.shape[1]
        Vn = np.zeros((V.shape[0], old + nupdate))
        Vn[:, :old] = V
        k = 0
        for j in range(self.size_update):
            if root_converged[j]:
                continue
            w = self._correction(rep["q"][:, j], rep["lambda"][j], rep["res"][:, j])
            Vn[:, old + k] = w / np.linalg.norm(w)
            k += 1
        self._gramschmidt(Vn, old)
        proj["V"] = Vn
        return nupdate

    # davidsonsolver.cc:442-478
    @staticmethod
    def _gramschmidt(Q, nstart):
        nupdate = Q.shape[1] - nstart
        norms = np.linalg.norm(Q[:, nstart:], axis=0)
        for rep in range(2):
            if nstart > 0:
                Q[:, nstart:] -= Q[:, :nstart] @ (Q[:, :nstart].T @ Q[:, nstart:])
                Q[:, nstart:] /= np.linalg.norm(Q[:, nstart:], axis=0)[None, :]
            for j in range(nstart + 1, Q.shape[1]):
                rng = j - nstart
                Q[:, j] -= Q[:, nstart:j] @ (Q[:, nstart:j].T @ Q[:, j])
                if rep == 1 and np.linalg.norm(Q[:, j]) <= 1e-12 * norms[rng]:
                    raise RuntimeError("Linear dependencies in Gram-Schmidt.")
                Q[:, j] /= np.linalg.norm(Q[:, j])
        return nupdate

    # davidsonsolver.cc:490-513
    def _restart(self, rep, proj, newvectors):
        V = proj["V"]
        rs = self.restart_size
        newV = np.zeros((V.shape[0], newvectors + rs))
        newV[:, rs:] = V[:, V.shape[1] - newvectors:]
        if self.matrix_type == "SYMM":
            newV[:, :rs] = rep["q"][:, :rs]
            proj["AV"] = proj["AV"] @ rep["U"][:, :rs]
        else:
            Qo, _ = np.linalg.qr(rep["U"][:, :rs])
            newV[:, :rs] = V[:, :V.shape[1] - newvectors] @ Qo
            proj["AV"] = proj["AV"] @ Qo
            proj["AAV"] = proj["AAV"] @ Qo
            proj["B"] = newV[:, :rs].T @ proj["AAV"]
        proj["T"] = newV[:, :rs].T @ proj["AV"]
        proj["V"] = newV

    def _store(self, rep, neigen):
        self.eigenvalues = rep["lambda"][:neigen].copy()
        ev = rep["q"][:, :neigen].copy()
        self.eigenvectors = ev / np.linalg.norm(ev, axis=0)[None, :]
