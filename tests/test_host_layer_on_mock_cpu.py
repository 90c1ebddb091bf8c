"""The C++ host layer (votca_b200/host: job facade, GW loop + QP search, BSE driver, Davidson, PPM, checkpoint writer,
basis-set path) on the CPU: the GPU tests of that layer are re-run, unchanged, in a child process in which the
kernel library is replaced by tests/host_harness/mock_b200_for_tests.cc - a naive single-rank CPU restatement of
the C ABI that exists for this purpose only (the product has no CPU fallback and nothing
under votca_b200/ references the mock).  This checks the host logic every round without a GPU; the kernels
themselves are only ever checked on the device.  The mock's own fidelity is checked by running the kernel tests
against it as well."""
import os
import re
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
MOCK = os.path.join(HERE, "host_harness", "build", "mock")


@pytest.fixture(scope="module")
def mock_dir():
    os.makedirs(MOCK, exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", os.path.join(MOCK, "libgwbse_b200.so"),
                    os.path.join(HERE, "host_harness", "mock_b200_for_tests.cc")], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-o",
                    os.path.join(MOCK, "libgwbse_host.so"), os.path.join(ROOT, "votca_b200", "host", "driver.cc"),
                    "-L" + MOCK, "-lgwbse_b200", "-Wl,-rpath,$ORIGIN"], check=True)
    return MOCK


def run_gpu_tests_on_mock(mock_dir, args):
    env = dict(os.environ, GWBSE_B200_TEST_MOCK_DIR=mock_dir)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "-rxXfE"] + args,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    tail = r.stdout[-4000:] + r.stderr[-2000:]
    m = re.search(r"(\d+) passed", r.stdout)
    return r.returncode, int(m.group(1)) if m else 0, tail


def test_mock_is_faithful_on_the_kernel_tests(mock_dir):
    """The stand-in passes the tests the CUDA kernels are checked with (tests/test_gpu_kernels.py: every C ABI entry
    point against the oracle and the reference's golden matrices; only the two large property tests are left out for
    time) - so what the host layer sees from it is what it sees from the device."""
    rc, passed, tail = run_gpu_tests_on_mock(mock_dir, ["tests/test_gpu_kernels.py", "-k",
                                                        "not medium_size and not full_size"])
    assert rc == 0 and passed >= 40, tail


def test_gw_and_bse_host_tests_pass_on_the_mock(mock_dir):
    """tests/test_gpu_host.py, all of it: golden G0W0 (test_gw.cc), canonical / Brent root search, evGW with the ppm,
    exact and cda integrators against the oracle, BSE TDA / full / triplets with dynamical screening (test_bse.cc),
    oscillator strengths, options XML, error convention; and tests/test_gpu_zz_reference_checkpoint.py: BSE on the
    reference's own dftgwbse checkpoints (water, d/f aux shells), the PPM known answer, the tier-R methane case."""
    rc, passed, tail = run_gpu_tests_on_mock(mock_dir, ["tests/test_gpu_host.py",
                                                        "tests/test_gpu_zz_reference_checkpoint.py"])
    assert rc == 0 and passed >= 17, tail


def test_basis_set_job_and_checkpoint_pass_on_the_mock(mock_dir):
    """Jobs fed with basis sets only (integrals from the shared ao3c_core source, interlevel dipoles formed by the
    host layer) against jobs fed with the oracle's arrays, the error texts of gwbse_basis_create, and the .orb
    checkpoint a job writes (tests/test_zzz_gpu_ao3c_device.py, tests/test_zz_gpu_orb_output.py)."""
    rc, passed, tail = run_gpu_tests_on_mock(mock_dir, [
        "tests/test_zzz_gpu_ao3c_device.py", "tests/test_zz_gpu_orb_output.py", "--runxfail", "-k",
        "not large_l"])
    assert rc == 0 and passed >= 4, tail


def test_late_gpu_test_files_are_sound_on_the_mock(mock_dir):
    """The GPU test files written after the round's GPU budget was spent (the reference's cudapipeline / cudamatrix
    cases, the gpu_benchmark tool) run clean against the mock: their own code - argument order, shapes, pointer
    offsets, expectations - is right, so a failure on the device would be the library's."""
    rc, passed, tail = run_gpu_tests_on_mock(mock_dir, [
        "tests/test_zz_gpu_cudapipeline_cases.py", "tests/test_zz_gpu_benchmark_tool.py", "--runxfail"])
    assert rc == 0 and passed >= 6, tail


def test_smoke_entry_point_logic_on_the_mock(mock_dir):
    """__graft_entry__.smoke() - the first thing the driver runs on the GPU box - executed against the mock: its own
    plumbing and its oracle comparison are sound (on the box it runs on the CUDA library, as everything else)."""
    code = (
        "import os, sys\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "from votca_b200 import _capi\n"
        f"_capi._api = _capi.CApi(os.path.join(r'{mock_dir}', 'libgwbse_b200.so'), _capi.HEADER)\n"
        f"_capi._host_api = _capi.CApi(os.path.join(r'{mock_dir}', 'libgwbse_host.so'), _capi.HOST_HEADER)\n"
        "import __graft_entry__ as g\n"
        "g.smoke()\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "smoke: max|dQP|" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
