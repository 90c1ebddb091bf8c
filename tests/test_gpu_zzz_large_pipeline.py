"""Full GW-BSE pipelines through the CUDA path against the oracle at sizes where the code paths of the benchmark are
active (the golden matrices of the reference's unit tests are 17-function systems):

* BASELINE.json config 2 - benzene def2-tzvp + aux-def2-tzvp, evGW with the EXACT self-energy (rpa.cc:204-326,
  sigma_exact.cc:29-148: 4221 x 4221 two-particle Hamiltonian, residues for 62 levels), full BSE (TDA off) 10 singlets,
  oscillator strengths; fed with host integrals and, separately, with integrals produced on the device;
* synthetic tier-S systems (N = 160 and N = 420) with evGW(ppm) + full BSE, with the treecode Sigma_c evaluator, a
  chunked BSE intermediate and both GEMM kernels (TMA-staged and cp.async) forced on - the paths the DCV5T / C60 runs
  take, which the small fixtures never reach.

Oracle numbers: tests/golden/*.npz, written by tests/golden/make_benzene_tzvp.py and make_synthetic_pipeline.py.
Tolerances are the north star's: QP and BSE energies 1e-6 Ha, oscillator strengths 1e-5."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _shell_sums(energies, values, tol=1e-4):
    """Sums of `values` over groups of (near-)degenerate energies: the members of a degenerate level are only
    defined up to a rotation, their sum is not."""
    out, i = [], 0
    while i < len(energies):
        j = i
        while j + 1 < len(energies) and abs(energies[j + 1] - energies[i]) < tol:
            j += 1
        out.append(float(np.sum(values[i:j + 1])))
        i = j + 1
    return np.array(out)


def _benzene_job(c, device_integrals):
    from oracle import bse as obse
    from votca_b200 import realsys
    from votca_b200.api import Job
    homo, q = c["homo"], c["q"]
    job = Job(0)
    job.set_scalar("homo", homo)
    job.set_array("mos", c["mos"])
    job.set_array("mo_energies", c["mo_energies"])
    job.set_array("vxc", c["vxc"])
    if device_integrals:
        s = realsys.system("benzene-tzvp")
        job.set_basis("dft", *s["dft"])
        job.set_basis("aux", *s["aux"])
    else:
        job.set_ao3c(c["ao3c"])
        job.set_array("aux_overlap", c["S"])
        job.set_array("aux_coulomb", c["V"])
        vt, ct = homo + 1, q - homo - 1
        for ax, d in zip("xyz", obse.free_transition_dipoles(c["dipole"], c["mos"], 0, vt, homo + 1, ct)):
            job.set_array("dipole_" + ax, d)
    job.set_options(tasks="gw,singlets", gw__mode="evGW", gw__sigma_integrator="exact", bse__exctotal=10,
                    bse__useTDA=False, bse__davidson__tolerance="lapack", bse__davidson__maxiter=200)
    return job


@pytest.mark.parametrize("device_integrals", [False, True], ids=["host-integrals", "device-integrals"])
def test_config2_benzene_tzvp_evgw_exact_full_bse(device_integrals):
    from tests.helpers import benzene_tzvp_case
    c = benzene_tzvp_case()
    job = _benzene_job(c, device_integrals)
    try:
        job.run()
        assert (job.scalar("qpmax"), job.scalar("bse_cmax"), job.scalar("rpamax")) == (c["q"] - 1, c["q"] - 1, 221)
        assert int(job.scalar("removed_functions")) == int(c["removed"])
        assert int(job.scalar("gw_iterations")) == int(c["gw_iterations"])
        assert np.abs(job.get("QPpert_energies") - c["QPpert_energies"]).max() < 1e-6
        assert np.abs(job.get("RPA_inputenergies") - c["RPA_inputenergies"]).max() < 1e-6
        es = job.get("BSE_singlet_eigenvalues")
        assert job.scalar("singlet_converged") == 1.0
        assert np.abs(es - c["BSE_singlet_eigenvalues"]).max() < 1e-6
        f_ref = _shell_sums(c["BSE_singlet_eigenvalues"], c["oscillator_strengths"])
        f_got = _shell_sums(es, job.get("oscillator_strengths"))
        assert f_ref.shape == f_got.shape and np.abs(f_ref - f_got).max() < 1e-5
        assert f_ref.max() > 1.0  # the bright E1u pair of benzene is in the window
    finally:
        job.close()


def _synthetic_job(name, exctotal=10):
    from votca_b200 import synthetic
    from votca_b200.api import Job
    N, naux, homo = synthetic.CONFIGS[name]
    s = synthetic.make_small(N, naux, homo)
    job = Job(0)
    job.set_scalar("homo", homo)
    for k in ("mos", "mo_energies", "vxc", "aux_overlap", "aux_coulomb"):
        job.set_array(k, s[k])
    job.set_ao3c(s["ao3c"])
    job.set_options(tasks="gw,singlets", gw__mode="evGW", gw__sigma_integrator="ppm", bse__exctotal=exctotal,
                    bse__useTDA=False)
    return job, s


def _check_synthetic(job, ref):
    job.run()
    assert int(job.scalar("gw_iterations")) == int(ref["gw_iterations"])
    assert np.abs(job.get("QPpert_energies") - ref["QPpert_energies"]).max() < 1e-6
    assert np.abs(np.diag(job.get("Sigma_x")) - ref["Sigma_x_diag"]).max() < 1e-8
    assert np.abs(np.diag(job.get("Sigma_c")) - ref["Sigma_c_diag"]).max() < 1e-6
    assert np.abs(job.get("Hqp") - ref["Hqp"]).max() < 1e-6
    assert job.scalar("singlet_converged") == 1.0
    assert np.abs(job.get("BSE_singlet_eigenvalues") - ref["BSE_singlet_eigenvalues"]).max() < 1e-6


@pytest.mark.parametrize("name", ["small", "medium"])
@pytest.mark.parametrize("tma", [1, 0], ids=["tma-gemm", "cpasync-gemm"])
@pytest.mark.parametrize("dense", [0, 1], ids=["factorised-direct", "materialised-direct"])
def test_synthetic_pipeline_with_the_benchmark_code_paths(name, tma, dense):
    """evGW(ppm) + full BSE against the oracle with: the treecode evaluator for every Sigma_c element (it normally
    starts at 32768 terms), the BSE intermediate cut into several chunks or the direct terms applied from their
    materialised blocks (the default policy, as in the benchmark), split-K plans (the planner picks them for the
    epsilon SYRK and the long-K BSE leg at these shapes) - once with the TMA-staged GEMM, once with cp.async."""
    path = os.path.join(GOLDEN, f"synthetic_{name}_evgw_ppm_bse.npz")
    if not os.path.exists(path):
        pytest.skip(f"{os.path.basename(path)} not generated (tests/golden/make_synthetic_pipeline.py {name})")
    with np.load(path) as z:
        ref = {k: z[k] for k in z.files}
    job, s = _synthetic_job(name)
    try:
        assert np.allclose(ref["input_checksum"], [s["ao3c"].sum(), s["mos"].sum(), s["aux_coulomb"].sum()], rtol=1e-9)
        k = job.kernel_ctx()
        k.set_option("tma", tma)
        k.set_option("sigma_tree_min_terms", 0)
        k.set_option("bse_chunk_bytes", 8 << 20)
        k.set_option("bse_dense", dense)
        builds0 = k.bse_dense_stats()[0]
        k.gemm_profile(True)
        _check_synthetic(job, ref)
        assert (k.bse_dense_stats()[0] - builds0 >= 2) == (dense == 1)  # A carries Hd, B carries Hd2
        shapes = k.gemm_shape_report()
        k.gemm_profile(False)
        import re
        assert bool(re.search(r" cfg1[0-4]", shapes)) == (tma == 1), shapes  # TMA tile shapes are cfg10..14
        assert any(f" sk{n}" in shapes for n in range(2, 65)), "no split-K plan was exercised"
    finally:
        job.kernel_ctx().set_option("tma", 1)
        job.kernel_ctx().set_option("bse_dense", 1)
        job.close()
