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


MOCK_RUN_FILES = ["tests/test_gpu_kernels.py", "tests/test_gpu_host.py", "tests/test_gpu_zz_reference_checkpoint.py",
                  "tests/test_zzz_gpu_ao3c_device.py", "tests/test_zz_gpu_orb_output.py",
                  "tests/test_zz_gpu_cudapipeline_cases.py", "tests/test_zz_gpu_benchmark_tool.py",
                  "tests/test_zz_gpu_host_options.py", "tests/test_gpu_zz_qsgw.py", "tests/test_zz_gpu_bsecoupling.py", "tests/test_zz_gpu_uks.py", "tests/test_zz_gpu_dftgwbse_tool.py"]


@pytest.fixture(scope="module")
def mock_run(mock_dir):
    """One child process runs every GPU test file against the mock (shared oracle caches); {test id: outcome}.
    Left out for time: the two large property tests of the kernels and the (G G | I) integrals through the oracle."""
    env = dict(os.environ, GWBSE_B200_TEST_MOCK_DIR=mock_dir)
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--runxfail", "-rA",
                        "-k", "not medium_size and not full_size and not large_l"] + MOCK_RUN_FILES,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=2400)
    outcomes = dict((m.group(2), m.group(1)) for m in re.finditer(r"^(PASSED|FAILED|ERROR) (\S+)", r.stdout, re.M))
    return outcomes, r.stdout[-6000:] + r.stderr[-2000:]


def _check(mock_run, prefix, at_least):
    outcomes, tail = mock_run
    mine = {k: v for k, v in outcomes.items() if k.startswith(prefix)}
    bad = {k: v for k, v in mine.items() if v != "PASSED"}
    assert not bad and len(mine) >= at_least, (bad, len(mine), tail)


def test_mock_is_faithful_on_the_kernel_tests(mock_run):
    """The stand-in passes the tests the CUDA kernels are checked with (tests/test_gpu_kernels.py: every C ABI entry
    point against the oracle and the reference's golden matrices) - so what the host layer sees from it is what it
    sees from the device."""
    _check(mock_run, "tests/test_gpu_kernels.py", 40)


def test_gw_and_bse_host_tests_pass_on_the_mock(mock_run):
    """tests/test_gpu_host.py, all of it: golden G0W0 (test_gw.cc), canonical / Brent root search, evGW with the ppm,
    exact and cda integrators against the oracle, BSE TDA / full / triplets with dynamical screening (test_bse.cc),
    oscillator strengths, options XML, error convention; and tests/test_gpu_zz_reference_checkpoint.py: BSE on the
    reference's own dftgwbse checkpoints (water, d/f aux shells), the PPM known answer, the tier-R methane case."""
    _check(mock_run, "tests/test_gpu_host.py", 13)
    _check(mock_run, "tests/test_gpu_zz_reference_checkpoint.py", 4)


def test_basis_set_job_and_checkpoint_pass_on_the_mock(mock_run):
    """Jobs fed with basis sets only (integrals from the shared ao3c_core source, interlevel dipoles formed by the
    host layer) against jobs fed with the oracle's arrays, the error texts of gwbse_basis_create, and the .orb
    checkpoint a job writes (tests/test_zzz_gpu_ao3c_device.py, tests/test_zz_gpu_orb_output.py)."""
    _check(mock_run, "tests/test_zzz_gpu_ao3c_device.py", 4)
    _check(mock_run, "tests/test_zz_gpu_orb_output.py", 1)


def test_late_gpu_test_files_are_sound_on_the_mock(mock_run):
    """The GPU test files written after the round's GPU budget was spent (the reference's cudapipeline / cudamatrix
    cases, the gpu_benchmark tool, the less-travelled host options against the oracle) run clean against the mock:
    their own code - argument order, shapes, pointer offsets, expectations - is right, so a failure on the device
    would be the library's."""
    _check(mock_run, "tests/test_zz_gpu_cudapipeline_cases.py", 4)
    _check(mock_run, "tests/test_zz_gpu_benchmark_tool.py", 2)
    _check(mock_run, "tests/test_zz_gpu_host_options.py", 4)


def test_qsgw_passes_on_the_mock(mock_run):
    """QSGW loop of the host layer (tests/test_gpu_zz_qsgw.py): ppm and exact integrators against the oracle, virtual
    threshold trimming, BSE in the QP basis."""
    _check(mock_run, "tests/test_gpu_zz_qsgw.py", 3)


def test_bsecoupling_passes_on_the_mock(mock_run):
    """BSECoupling host class (projection algebra, perturbation / reduction methods, XML output) against the
    reference's known answers and the oracle (tests/test_zz_gpu_bsecoupling.py)."""
    _check(mock_run, "tests/test_zz_gpu_bsecoupling.py", 4)


def test_uks_twins_pass_on_the_mock(mock_run):
    """GW_UKS / BSE_UKS (votca_b200/host/uks.h, two contexts) against oracle/uks.py and, in the closed-shell limit,
    against the restricted path and the reference's gw/ref.mm (tests/test_zz_gpu_uks.py)."""
    _check(mock_run, "tests/test_zz_gpu_uks.py", 5)


def test_dftgwbse_tool_passes_on_the_mock(mock_run):
    _check(mock_run, "tests/test_zz_gpu_dftgwbse_tool.py", 2)


def test_every_host_layer_gpu_test_file_runs_on_the_mock():
    """Every GPU test file is either re-run against the mock above or listed here with the reason it is not (it
    needs the real kernels or several devices), so a new GPU test file cannot silently skip the CPU check."""
    import glob
    device_only = {"tests/test_gpu_multi.py": "needs two devices and NCCL",
                   "tests/test_gpu_zzz_large_pipeline.py": "forces the treecode / split-K / chunked paths of the CUDA library",
                   "tests/test_gpu_zzz_streaming_roofline.py": "times kernels"}
    gpu_files = [os.path.relpath(p, ROOT) for p in glob.glob(os.path.join(HERE, "test_*.py"))
                 if re.search(r"pytest\.mark\.gpu", open(p).read()) and os.path.basename(p) != os.path.basename(__file__)]
    missing = sorted(set(gpu_files) - set(MOCK_RUN_FILES) - set(device_only))
    assert gpu_files and not missing, missing


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


def test_fill_with_the_reference_argument_list_on_the_mock(mock_dir):
    """TCMatrix_gwbse::Fill(auxbasis, dftbasis, dft_orbitals) (threecenter.cc:72-90) of the C++ host layer, from RAW
    basis-set contractions (AOBasisData::NormalizeFromRawContractions): overlap, two- and three-centre integrals and
    V^-1/2 all behind the one call; result equals the oracle's TCMatrix.fill on methane def2-svp."""
    so = os.path.join(mock_dir, "libfill_overload.so")
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", so,
                    os.path.join(HERE, "host_harness", "fill_overload_harness.cc"), "-L" + mock_dir, "-lgwbse_b200",
                    "-Wl,-rpath,$ORIGIN"], check=True)
    code = (
        "import ctypes, json, sys\n"
        "import numpy as np\n"
        f"sys.path.insert(0, r'{ROOT}')\n"
        "from oracle import threecenter\n"
        "from tests import helpers\n"
        "from votca_b200 import realsys\n"
        f"lib = ctypes.CDLL(r'{so}')\n"
        "c = helpers.methane_svp_case()\n"
        "g = helpers.load_golden()\n"
        "el = [str(e) for e in g['molecule_methane_tutorial/elements']]\n"
        "pos = np.asarray(g['molecule_methane_tutorial/positions_bohr'])\n"
        "def raw(name):\n"
        "    bs = realsys.basis_set(name)\n"
        "    l, npr, cen, ex, co = [], [], [], [], []\n"
        "    for e, p in zip(el, pos):\n"
        "        for sl, prims in bs[e]:\n"
        "            l.append(sl); npr.append(len(prims)); cen.append(p)\n"
        "            ex += [q[0] for q in prims]; co += [q[1] for q in prims]\n"
        "    return (np.array(l, dtype=np.int32), np.array(npr, dtype=np.int32), np.ascontiguousarray(cen, dtype=np.float64),\n"
        "            np.array(ex), np.array(co))\n"
        "d, a = raw('def2-svp'), raw('aux-def2-svp')\n"
        "N, naux, q = c['dft'].size, c['aux'].size, c['q']\n"
        "C = np.asfortranarray(c['hf']['mos'])\n"
        "out = np.zeros((q, naux, N)); removed = ctypes.c_long(); err = ctypes.create_string_buffer(256)\n"
        "P = ctypes.c_void_p\n"
        "lib.fill_from_bases.argtypes = [ctypes.c_int] + [P] * 5 + [ctypes.c_int] + [P] * 5 + [P, ctypes.c_long, ctypes.c_long, P, P, ctypes.c_char_p, ctypes.c_int]\n"
        "rc = lib.fill_from_bases(len(d[0]), *[x.ctypes.data for x in d], len(a[0]), *[x.ctypes.data for x in a],\n"
        "                         C.ctypes.data, N, q - 1, out.ctypes.data, ctypes.addressof(removed), err, 256)\n"
        "assert rc == 0, err.value\n"
        "tc = threecenter.TCMatrix(naux, 0, q - 1, 0, N - 1)\n"
        "tc.fill_from_integrals(c['ao3c'], c['S'], c['V'], c['hf']['mos'])\n"
        "got = out.transpose(0, 2, 1)\n"
        "rel = np.linalg.norm(got - tc.M) / np.linalg.norm(tc.M)\n"
        "assert rel < 1e-10 and removed.value == tc.removed, (rel, removed.value, tc.removed)\n"
        "print('fill ok', rel)\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "fill ok" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_host_layer_is_clean_under_address_and_ub_sanitizers(tmp_path):
    """Host library and mock rebuilt with -fsanitize=address,undefined; a selection of the host-layer tests (GW golden,
    evGW, BSE TDA / full, options, callback producer, basis-set job, .orb output) re-run in a child process with the
    sanitizer runtimes preloaded: no out-of-bounds access, use-after-free or undefined behaviour in votca_b200/host."""
    d = str(tmp_path)
    flags = ["-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-fPIC", "-shared"]
    subprocess.run(["g++"] + flags + ["-o", os.path.join(d, "libgwbse_b200.so"),
                                      os.path.join(HERE, "host_harness", "mock_b200_for_tests.cc")], check=True)
    subprocess.run(["g++"] + flags + ["-pthread", "-o", os.path.join(d, "libgwbse_host.so"),
                                      os.path.join(ROOT, "votca_b200", "host", "driver.cc"), "-L" + d, "-lgwbse_b200",
                                      "-Wl,-rpath,$ORIGIN"], check=True)
    rt = [subprocess.run(["g++", "-print-file-name=" + n], capture_output=True, text=True).stdout.strip()
          for n in ("libasan.so", "libubsan.so")]
    env = dict(os.environ, GWBSE_B200_TEST_MOCK_DIR=d, LD_PRELOAD=":".join(rt),
               ASAN_OPTIONS="detect_leaks=0:halt_on_error=1", UBSAN_OPTIONS="print_stacktrace=1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", "--runxfail",
                        "tests/test_gpu_host.py", "tests/test_zz_gpu_host_options.py", "tests/test_zz_gpu_orb_output.py",
                        "tests/test_zzz_gpu_ao3c_device.py", "-k",
                        "not large_l and not cda and not exact and not water and not fill_from_basis"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=1500)
    out = r.stdout + r.stderr
    assert r.returncode == 0 and "AddressSanitizer" not in out and "runtime error:" not in out, out[-4000:]
    assert int(re.search(r"(\d+) passed", r.stdout).group(1)) >= 15
