"""Thin NumPy-facing wrapper over the C ABI (tests / bench plumbing only).

Every method maps 1:1 onto an entry point of include/gwbse_b200.h; host arrays are
converted to column-major float64 and results come back as NumPy arrays.  No math
happens here.
"""
import ctypes

import numpy as np

from ._capi import capi, fmat, ptr


class GwbseError(RuntimeError):
    pass


class Context:
    def __init__(self, device=0):
        self.api = capi()
        h = ctypes.c_void_p()
        if self.api.gwbse_ctx_create(int(device), ctypes.byref(h)) != 0:
            raise GwbseError(self.api.gwbse_create_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.api.gwbse_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GwbseError(self.api.gwbse_last_error(self.h).decode())

    def call(self, name, *args):
        self._ck(getattr(self.api, name)(self.h, *args))

    # ---- misc ----
    def sync(self):
        self.call("gwbse_sync")

    def set_option(self, key, value):
        self.call("gwbse_set_option", key.encode(), float(value))

    def launch_count(self):
        return int(self.api.gwbse_launch_count(self.h))

    def timer_start(self):
        self.call("gwbse_timer_start")

    def timer_stop_ms(self):
        ms = ctypes.c_float()
        self.call("gwbse_timer_stop_ms", ctypes.byref(ms))
        return float(ms.value)

    def comm_init(self, rank, world, uid):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(uid))
        self.call("gwbse_comm_init", int(rank), int(world), buf)

    def nccl_unique_id(self):
        buf = (ctypes.c_ubyte * 128)()
        if self.api.gwbse_nccl_unique_id(buf) != 0:
            raise GwbseError("cannot create NCCL unique id")
        return bytes(buf)

    # ---- device memory ----
    def malloc(self, n):
        p = ctypes.c_void_p()
        self.call("gwbse_dev_malloc", ctypes.c_size_t(int(n) * 8), ctypes.byref(p))
        return p

    def free(self, p):
        self.call("gwbse_dev_free", p)

    def upload(self, a):
        a = fmat(a)
        p = self.malloc(max(a.size, 1))
        self.call("gwbse_h2d", p, ptr(a), ctypes.c_size_t(a.size))
        return p

    def h2d(self, p, a):
        a = fmat(a)
        self.call("gwbse_h2d", p, ptr(a), ctypes.c_size_t(a.size))

    def download(self, p, shape):
        out = np.empty(shape, dtype=np.float64, order="F")
        self.call("gwbse_d2h", ptr(out), p, ctypes.c_size_t(out.size))
        return out

    # ---- dense primitives ----
    def dgemm(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, cfg=-1, splitk=0):
        self.call("gwbse_dgemm_dev_ex", ta.encode(), tb.encode(), m, n, k, float(alpha), A, lda, B, ldb, float(beta),
                  C, ldc, cfg, splitk)

    def gemm_host(self, ta, tb, alpha, A, B, beta=0.0, C=None, cfg=-1, splitk=0):
        """Convenience: op(A) op(B) on host arrays through the device GEMM."""
        A, B = fmat(A), fmat(B)
        m = A.shape[1] if ta == "T" else A.shape[0]
        k = A.shape[0] if ta == "T" else A.shape[1]
        n = B.shape[0] if tb == "T" else B.shape[1]
        Cm = np.zeros((m, n), order="F") if C is None else fmat(C).copy(order="F")
        dA, dB, dC = self.upload(A), self.upload(B), self.upload(Cm)
        self.dgemm(ta, tb, m, n, k, alpha, dA, max(A.shape[0], 1), dB, max(B.shape[0], 1), beta, dC, max(m, 1),
                   cfg, splitk)
        out = self.download(dC, (m, n))
        for p in (dA, dB, dC):
            self.free(p)
        return out

    def sym_eig(self, A):
        A = fmat(A).copy(order="F")
        n = A.shape[0]
        d = self.upload(A)
        w = np.empty(n)
        self.call("gwbse_sym_eig_dev", n, d, n, ptr(w))
        V = self.download(d, (n, n))
        self.free(d)
        return w, V

    def inverse(self, A):
        A = fmat(A).copy(order="F")
        n = A.shape[0]
        d = self.upload(A)
        self.call("gwbse_inverse_dev", n, d, n)
        out = self.download(d, (n, n))
        self.free(d)
        return out

    def lu_solve(self, A, B):
        A, B = fmat(A).copy(order="F"), fmat(B).copy(order="F")
        if B.ndim == 1:
            B = B[:, None].copy(order="F")
        n, nrhs = A.shape[0], B.shape[1]
        dA, dB = self.upload(A), self.upload(B)
        self.call("gwbse_lu_solve_dev", n, nrhs, dA, n, dB, n)
        out = self.download(dB, (n, nrhs))
        self.free(dA)
        self.free(dB)
        return out

    def gen_eig(self, T, B):
        T, B = fmat(T), fmat(B)
        n = T.shape[0]
        wr, wi, VR = np.empty(n), np.empty(n), np.empty((n, n), order="F")
        self.call("gwbse_gen_eig_host", n, ptr(T), ptr(B), ptr(wr), ptr(wi), ptr(VR))
        return wr, wi, VR

    # ---- Mmn ----
    def mmn_alloc(self, naux, mmin, mmax, nmin, nmax):
        self.call("gwbse_mmn_alloc", naux, mmin, mmax, nmin, nmax)
        self.naux, self.mtotal, self.ntotal = naux, mmax - mmin + 1, nmax - nmin + 1

    def mmn_set_mos(self, mos):
        mos = fmat(mos)
        self.call("gwbse_mmn_set_mos", ptr(mos), mos.shape[0], mos.shape[0], mos.shape[1])

    def mmn_fill_block(self, aux_offset, ao3c):
        """ao3c: (count, N, N) symmetric matrices."""
        a = np.ascontiguousarray(ao3c, dtype=np.float64)
        self.call("gwbse_mmn_fill_block", aux_offset, a.shape[0], ptr(a))

    def mmn_mul_right(self, R):
        R = fmat(R)
        self.call("gwbse_mmn_mul_right", ptr(R), R.shape[0])

    def mmn_mul_right_window(self, R, n_lo, n_hi):
        """M[m] <- M[m] R with only the rows n in [n_lo, n_hi) rotated at once (gwbse_mmn_mul_right_window_dev)."""
        R = fmat(R)
        d = self.upload(R)
        try:
            self.call("gwbse_mmn_mul_right_window_dev", d, R.shape[0], int(n_lo), int(n_hi))
            self.sync()
        finally:
            self.free(d)

    # ---- AO Coulomb integrals on the device ----
    def basis_create(self, l, nprim, centers, exps, coefs):
        """Flat shell arrays (see gwbse_basis_create) -> opaque device basis handle."""
        l = np.ascontiguousarray(l, dtype=np.int32)
        nprim = np.ascontiguousarray(nprim, dtype=np.int32)
        centers = np.ascontiguousarray(centers, dtype=np.float64)
        exps = np.ascontiguousarray(exps, dtype=np.float64)
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        if centers.shape != (len(l), 3) or len(nprim) != len(l) or len(exps) != nprim.sum() or len(coefs) != len(exps):
            raise ValueError("inconsistent basis arrays")
        h = ctypes.c_void_p()
        self.call("gwbse_basis_create", len(l), ptr(l), ptr(nprim), ptr(centers), ptr(exps), ptr(coefs),
                  ctypes.byref(h))
        return h

    def basis_destroy(self, basis):
        self.call("gwbse_basis_destroy", basis)

    def basis_size(self, basis):
        return int(self.api.gwbse_basis_size(basis))

    def ao3c_block(self, aux, dft, aux_offset, aux_count):
        """(aux_count, N, N): the output of ComputeAO3cBlock for aux functions [aux_offset, aux_offset+aux_count)."""
        n = self.basis_size(dft)
        out = np.empty((aux_count, n, n), dtype=np.float64)
        self.call("gwbse_ao3c_block", aux, dft, int(aux_offset), int(aux_count), ptr(out))
        return out

    def ao_coulomb2c(self, aux):
        n = self.basis_size(aux)
        out = np.empty((n, n), order="F")
        self.call("gwbse_ao_coulomb2c", aux, ptr(out), n)
        return out

    def ao_overlap(self, basis):
        n = self.basis_size(basis)
        out = np.empty((n, n), order="F")
        self.call("gwbse_ao_overlap", basis, ptr(out), n)
        return out

    def ao_dipole(self, basis):
        """(3, n, n): <mu | r_k | nu> about the origin."""
        n = self.basis_size(basis)
        out = np.empty((3, n, n))
        self.call("gwbse_ao_dipole", basis, ptr(out), n)
        return out

    def mmn_fill_from_basis(self, aux, dft, aux_block=64):
        self.call("gwbse_mmn_fill_from_basis", aux, dft, int(aux_block))

    def mmn_get_slice(self, m):
        out = np.empty((self.ntotal, self.naux), order="F")
        self.call("gwbse_mmn_get_slice", m, ptr(out), self.ntotal)
        return out

    def mmn_set_slice(self, m, a):
        a = fmat(a)
        self.call("gwbse_mmn_set_slice", m, ptr(a), a.shape[0])

    def mmn_get_all(self):
        return np.array([self.mmn_get_slice(m) for m in range(self.mtotal)])

    def mmn_set_all(self, M):
        for m in range(self.mtotal):
            self.mmn_set_slice(m, M[m])

    def pseudo_invsqrt(self, S, V, etol=5e-7):
        S, V = fmat(S), fmat(V)
        n = S.shape[0]
        L = np.empty((n, n), order="F")
        removed = ctypes.c_int()
        self.call("gwbse_pseudo_invsqrt", n, ptr(S), ptr(V), float(etol), ptr(L), ctypes.byref(removed))
        return L, int(removed.value)

    # ---- RPA ----
    def rpa_epsilon(self, kind, freq, eta, energies, homo, rpamin, rpamax, fetch=True):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        fre, fim = (freq.real, freq.imag) if isinstance(freq, complex) else (float(freq), 0.0)
        out = np.empty((self.naux, self.naux), order="F") if fetch else None
        self.call("gwbse_rpa_epsilon", kind, fre, fim, float(eta), ptr(e), homo, rpamin, rpamax, ptr(out),
                  self.naux)
        return out

    def rpa_h2p_apb(self, energies, homo, rpamin, rpamax):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        S = (homo + 1 - rpamin) * (rpamax - homo)
        d = self.malloc(S * S)
        self.call("gwbse_rpa_h2p_apb", ptr(e), homo, rpamin, rpamax, d, S)
        out = self.download(d, (S, S))
        self.free(d)
        return out

    # ---- Sigma ----
    def sigma_x(self, homo, rpamin, qpmin, qpmax):
        q = qpmax - qpmin + 1
        out = np.empty((q, q), order="F")
        self.call("gwbse_sigma_x", homo, rpamin, qpmin, qpmax, ptr(out), q)
        return out

    def sigma_ppm_set(self, weight, freq, energies, homo, rpamin, qpmin, eta):
        w, f, e = (np.ascontiguousarray(x, dtype=np.float64) for x in (weight, freq, energies))
        self.call("gwbse_sigma_ppm_set", ptr(w), ptr(f), ptr(e), homo, rpamin, qpmin, float(eta))

    def mmn_rotate(self, U, qpmin, qpmax):
        u = fmat(U)
        self.call("gwbse_mmn_rotate", ptr(u), u.shape[0], int(qpmin), int(qpmax))

    def sigma_update_energies(self, which, energies):
        e = np.ascontiguousarray(energies, dtype=np.float64)
        self.call("gwbse_sigma_update_energies", int(which), ptr(e))

    def _sigma_eval(self, fn, levels, freqs, deriv):
        lv = np.ascontiguousarray(levels, dtype=np.int32)
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        s = np.empty(len(lv))
        ds = np.empty(len(lv)) if deriv else None
        self.call(fn, len(lv), ptr(lv), ptr(fr), ptr(s), ptr(ds))
        return (s, ds) if deriv else s

    def sigma_ppm_eval(self, levels, freqs, deriv=False):
        return self._sigma_eval("gwbse_sigma_ppm_eval", levels, freqs, deriv)

    def sigma_ppm_offdiag(self, freqs):
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        q = len(fr)
        out = np.empty((q, q), order="F")
        self.call("gwbse_sigma_ppm_offdiag", q, ptr(fr), ptr(out), q)
        return out

    def sigma_exact_prepare(self, omegas, XpY, energies, homo, rpamin, rpamax, qpmin, qpmax, eta):
        om, e = (np.ascontiguousarray(x, dtype=np.float64) for x in (omegas, energies))
        d = self.upload(XpY)
        try:
            self.call("gwbse_sigma_exact_prepare", ptr(om), d, XpY.shape[0], ptr(e), homo, rpamin, rpamax, qpmin,
                      qpmax, float(eta))
        finally:
            self.free(d)

    def sigma_exact_eval(self, levels, freqs, deriv=False):
        return self._sigma_eval("gwbse_sigma_exact_eval", levels, freqs, deriv)

    def sigma_exact_offdiag(self, freqs):
        fr = np.ascontiguousarray(freqs, dtype=np.float64)
        q = len(fr)
        out = np.empty((q, q), order="F")
        self.call("gwbse_sigma_exact_offdiag", q, ptr(fr), ptr(out), q)
        return out

    # ---- BSE ----
    def bse_configure(self, homo, rpamin, vmin, cmax, eps_inv, Hqp):
        e = np.ascontiguousarray(eps_inv, dtype=np.float64)
        H = fmat(Hqp)
        self.call("gwbse_bse_configure", homo, rpamin, vmin, cmax, ptr(e), ptr(H), H.shape[0])
        self.bse_size = (homo - vmin + 1) * (cmax - homo)

    def bse_matmul(self, coeffs, X):
        X = fmat(X)
        if X.ndim == 1:
            X = X[:, None].copy(order="F")
        Y = np.empty_like(X, order="F")
        cqp, cx, cd, cd2 = coeffs
        self.call("gwbse_bse_matmul", cqp, cx, cd, cd2, X.shape[1], ptr(X), X.shape[0], ptr(Y), Y.shape[0])
        return Y

    def bse_diagonal(self, coeffs):
        out = np.empty(self.bse_size)
        cqp, cx, cd, cd2 = coeffs
        self.call("gwbse_bse_diagonal", cqp, cx, cd, cd2, ptr(out))
        return out

    # ---- Davidson helpers ----
    def gramschmidt(self, Q, nstart):
        Q = fmat(Q).copy(order="F")
        d = self.upload(Q)
        try:
            self.call("gwbse_gramschmidt_dev", Q.shape[0], Q.shape[1], nstart, d, Q.shape[0])
            return self.download(d, Q.shape)
        finally:
            self.free(d)

    def davidson_correction(self, diag, lam, R, Q, olsen=False):
        R, Q = fmat(R), fmat(Q)
        lam = np.ascontiguousarray(lam, dtype=np.float64)
        dd, dR, dQ = self.upload(diag), self.upload(R), self.upload(Q)
        dW = self.malloc(R.size)
        try:
            self.call("gwbse_davidson_correction_dev", R.shape[0], R.shape[1], int(olsen), dd, ptr(lam), dR,
                      R.shape[0], dQ, Q.shape[0], dW, R.shape[0])
            return self.download(dW, R.shape)
        finally:
            for p in (dd, dR, dQ, dW):
                self.free(p)


class Job:
    """GWBSE job facade of the C++ host layer (include/gwbse_host.h): options in, .orb-named arrays out."""

    def __init__(self, device=0):
        from ._capi import host_api
        self.api = host_api()
        h = ctypes.c_void_p()
        if self.api.gwbse_job_create(int(device), ctypes.byref(h)) != 0:
            raise GwbseError(self.api.gwbse_job_create_error().decode())
        self.h = h
        self._keep = []

    def close(self):
        if getattr(self, "h", None):
            self.api.gwbse_job_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise GwbseError(self.api.gwbse_job_error(self.h).decode())

    def comm_init(self, rank, world, uid):
        buf = (ctypes.c_ubyte * 128).from_buffer_copy(bytes(uid))
        self._ck(self.api.gwbse_job_comm_init(self.h, int(rank), int(world), buf))

    def set_options(self, **kw):
        """Keys as in gwbse.xml with '.' written as '__', e.g. gw__mode='G0W0'."""
        for k, v in kw.items():
            self.set_option(k.replace("__", "."), v)

    def set_option(self, key, value):
        if isinstance(value, bool):
            value = "true" if value else "false"
        self._ck(self.api.gwbse_job_set_option(self.h, key.encode(), str(value).encode()))

    def load_options_xml(self, path):
        self._ck(self.api.gwbse_job_load_options_xml(self.h, str(path).encode()))

    def set_scalar(self, name, value):
        self._ck(self.api.gwbse_job_set_scalar(self.h, name.encode(), float(value)))

    def set_array(self, name, a):
        a = fmat(a)
        if a.ndim == 1:
            a = a[:, None]
        self._ck(self.api.gwbse_job_set_array(self.h, name.encode(), ptr(a), a.shape[0], a.shape[1]))

    def set_ao3c(self, ao3c):
        """ao3c: (naux, N, N) C-contiguous; referenced, not copied."""
        a = np.ascontiguousarray(ao3c, dtype=np.float64)
        self._keep.append(a)
        naux, N = a.shape[0], a.shape[1]
        self._ck(self.api.gwbse_job_set_array(self.h, b"ao3c", ptr(a), N * N, naux))

    def set_basis(self, which, l, nprim, centers, exps, coefs):
        """which = 'dft' | 'aux': AO integrals are then produced on the device (no ao3c array needed)."""
        l = np.ascontiguousarray(l, dtype=np.int32)
        nprim = np.ascontiguousarray(nprim, dtype=np.int32)
        centers = np.ascontiguousarray(centers, dtype=np.float64)
        exps = np.ascontiguousarray(exps, dtype=np.float64)
        coefs = np.ascontiguousarray(coefs, dtype=np.float64)
        if centers.shape != (len(l), 3) or len(nprim) != len(l) or len(exps) != nprim.sum() or len(coefs) != len(exps):
            raise ValueError("inconsistent basis arrays")
        self._ck(self.api.gwbse_job_set_basis(self.h, which.encode(), len(l), ptr(l), ptr(nprim), ptr(centers),
                                              ptr(exps), ptr(coefs)))

    def set_orb_output(self, path):
        """Write the results as an .orb (HDF5) checkpoint at the end of run()."""
        self._ck(self.api.gwbse_job_set_orb_output(self.h, str(path).encode() if path else None))

    def set_summary_output(self, path):
        """Write <job>_summary.xml (GWBSE::addoutput) at the end of run()."""
        self._ck(self.api.gwbse_job_set_summary_output(self.h, str(path).encode() if path else None))

    def run(self):
        self._ck(self.api.gwbse_job_run(self.h))

    def run_coupling(self):
        """BSECoupling::CalculateCouplings on the job's dimer inputs and the "A.*" / "B.*" monomer inputs
        (include/gwbse_host.h: gwbse_job_run_coupling); options: set_option("bsecoupling.<key>", value)."""
        self._ck(self.api.gwbse_job_run_coupling(self.h))

    def run_uks(self):
        """GW_UKS (+ BSE_UKS for the task exciton_uks) on an unrestricted reference: beta channel inputs "mos_beta",
        "mo_energies_beta", "vxc_beta", scalar "homo_beta" (include/gwbse_host.h: gwbse_job_run_uks)."""
        self._ck(self.api.gwbse_job_run_uks(self.h))

    def coupling_xml(self):
        return self.api.gwbse_job_coupling_xml(self.h).decode()

    def get(self, name):
        r, c = ctypes.c_long(), ctypes.c_long()
        if self.api.gwbse_job_array_dims(self.h, name.encode(), ctypes.byref(r), ctypes.byref(c)) != 0:
            raise KeyError(name)
        out = np.empty((r.value, c.value), order="F")
        if out.size:
            self.api.gwbse_job_get_array(self.h, name.encode(), ptr(out))
        return out[:, 0] if c.value == 1 else out

    def scalar(self, name):
        v = ctypes.c_double()
        if self.api.gwbse_job_get_scalar(self.h, name.encode(), ctypes.byref(v)) != 0:
            raise KeyError(name)
        return float(v.value)

    def log(self):
        return self.api.gwbse_job_log(self.h).decode()

    def launch_count(self):
        return int(self.api.gwbse_job_launch_count(self.h))


def _job_extras():
    def set_ao3c_dev(self, nbasis, naux, dev_ptr):
        """AO tensor already on the device (int address, e.g. torch tensor.data_ptr())."""
        self._ck(self.api.gwbse_job_set_ao3c_dev(self.h, int(nbasis), int(naux), ctypes.c_void_p(int(dev_ptr))))

    def set_ao3c_host_ptr(self, nbasis, naux, host_ptr):
        """AO tensor in (pinned) host memory given by raw address; referenced, not copied."""
        self._ck(self.api.gwbse_job_set_array(self.h, b"ao3c", ctypes.c_void_p(int(host_ptr)), int(nbasis) * int(nbasis),
                                              int(naux)))

    def set_ao3c_partial(self, nbasis, naux, first_aux, count, ptr_, on_device):
        """This rank's share [first_aux, first_aux + count) of the AO tensor by raw host / device address."""
        self._ck(self.api.gwbse_job_set_ao3c_partial(self.h, int(nbasis), int(naux), int(first_aux), int(count),
                                                     ctypes.c_void_p(int(ptr_)), int(bool(on_device))))

    def set_ao3c_callback(self, nbasis, naux, producer):
        """AO integral producer called block by block, as the reference's libint loop would be
        (gwbse_job_set_ao3c_callback): producer(aux_offset, aux_count) -> array (aux_count, N, N)."""
        proto = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_long, ctypes.c_long, ctypes.POINTER(ctypes.c_double))
        n2 = int(nbasis) * int(nbasis)

        def trampoline(_user, aux_offset, aux_count, out):
            blk = np.ascontiguousarray(producer(int(aux_offset), int(aux_count)), dtype=np.float64)
            ctypes.memmove(out, blk.ctypes.data, 8 * n2 * int(aux_count))

        cb = proto(trampoline)
        self._keep.append(cb)
        self._ck(self.api.gwbse_job_set_ao3c_callback(self.h, int(nbasis), int(naux), cb, None))

    def kernel_ctx(self):
        """The underlying gwbse_b200 context as a Context-like wrapper (profiling, timers)."""
        c = Context.__new__(Context)
        c.api = capi()
        c.h = ctypes.c_void_p(self.api.gwbse_job_ctx(self.h))
        c.close = lambda: None  # owned by the job
        return c

    Job.set_ao3c_dev = set_ao3c_dev
    Job.set_ao3c_host_ptr = set_ao3c_host_ptr
    Job.set_ao3c_partial = set_ao3c_partial
    Job.set_ao3c_callback = set_ao3c_callback
    Job.kernel_ctx = kernel_ctx

    def gemm_profile(self, enable=True):
        self.call("gwbse_gemm_profile", int(bool(enable)))

    def gemm_stats(self):
        ms, fl, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_longlong()
        self.call("gwbse_gemm_stats", ctypes.byref(ms), ctypes.byref(fl), ctypes.byref(n))
        return {"ms": ms.value, "flops": fl.value, "launches": int(n.value)}

    def fp64_peak_probe(self):
        t = ctypes.c_double()
        self.call("gwbse_fp64_peak_probe", ctypes.byref(t))
        return t.value

    Context.gemm_profile = gemm_profile
    Context.gemm_stats = gemm_stats
    Context.fp64_peak_probe = fp64_peak_probe


_job_extras()


def _profile_report(self):
    buf = ctypes.create_string_buffer(16384)
    self.call("gwbse_profile_report", buf, ctypes.c_size_t(len(buf)))
    return buf.value.decode()


def _gemm_shape_report(self):
    buf = ctypes.create_string_buffer(65536)
    self.call("gwbse_gemm_shape_report", buf, ctypes.c_size_t(len(buf)))
    return buf.value.decode()


def _bse_stats(self, reset=False):
    fl, pr, co = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_longlong()
    self.call("gwbse_bse_stats", ctypes.byref(fl), ctypes.byref(pr), ctypes.byref(co), int(bool(reset)))
    return fl.value, pr.value, co.value


Context.bse_stats = _bse_stats


def _bse_dense_stats(self):
    """(blocks built, trial columns applied from a resident block, bytes resident now)"""
    b, c, r = ctypes.c_longlong(), ctypes.c_longlong(), ctypes.c_double()
    self.call("gwbse_bse_dense_stats", ctypes.byref(b), ctypes.byref(c), ctypes.byref(r))
    return int(b.value), int(c.value), float(r.value)


Context.bse_dense_stats = _bse_dense_stats
Context.profile_report = _profile_report
Context.gemm_shape_report = _gemm_shape_report
