import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")
    # tests/test_host_layer_on_mock_cpu.py re-runs the GPU tests of the C++ host layer in a child process against a
    # CPU stand-in for the kernel library (tests/host_harness/mock_b200_for_tests.cc, test infrastructure only):
    # the child redirects the bindings to the directory holding the mock build.  Never set outside that test.
    mock = os.environ.get("GWBSE_B200_TEST_MOCK_DIR")
    if mock:
        from votca_b200 import _capi
        _capi.LIBPATH = os.path.join(mock, "libgwbse_b200.so")
        _capi.HOST_LIBPATH = os.path.join(mock, "libgwbse_host.so")
        _capi._api = _capi.CApi(_capi.LIBPATH, _capi.HEADER)
        _capi._host_api = _capi.CApi(_capi.HOST_LIBPATH, _capi.HOST_HEADER)
        _install_mock_lapack_hook(_capi._api.lib)


_MOCK_HOOKS = []


def _install_mock_lapack_hook(lib):
    """T x = lambda B x for the mock's gwbse_gen_eig_host, with LAPACK dgeev's output convention (complex pairs in
    adjacent columns: real part, imaginary part), computed by scipy in the test process."""
    import ctypes

    import numpy as np
    import scipy.linalg
    dp = ctypes.POINTER(ctypes.c_double)
    proto = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, dp, dp, dp, dp, dp)

    def gen_eig(n, T, B, wr, wi, VR):
        try:
            Tm = np.ctypeslib.as_array(T, (n * n,)).reshape(n, n).T
            Bm = np.ctypeslib.as_array(B, (n * n,)).reshape(n, n).T
            w, V = scipy.linalg.eig(np.linalg.solve(Bm, Tm))
            out = np.zeros((n, n))
            j = 0
            while j < n:
                if abs(w[j].imag) > 0 and j + 1 < n:
                    out[:, j], out[:, j + 1] = V[:, j].real, V[:, j].imag
                    j += 2
                else:
                    out[:, j] = V[:, j].real
                    j += 1
            np.ctypeslib.as_array(wr, (n,))[:] = w.real
            np.ctypeslib.as_array(wi, (n,))[:] = w.imag
            np.ctypeslib.as_array(VR, (n * n,))[:] = out.T.ravel()
            return 0
        except Exception:
            return 1

    cb = proto(gen_eig)
    _MOCK_HOOKS.append(cb)
    lib.mock_set_gen_eig.argtypes = [proto]
    lib.mock_set_gen_eig.restype = None
    lib.mock_set_gen_eig(cb)


@pytest.fixture(scope="session")
def golden():
    from tests.helpers import load_golden
    return load_golden()


@pytest.fixture(scope="session")
def methane():
    """Methane / 3-21G AO integrals of the reference fixtures (oracle integral code)."""
    from tests.helpers import methane_integrals
    return methane_integrals()
