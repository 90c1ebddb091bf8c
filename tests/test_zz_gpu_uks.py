"""UKS twins through the CUDA path: GW_UKS and BSE_UKS of votca_b200/host/uks.h behind
gwbse_job_run_uks, against the oracle (oracle/uks.py) on an open-shell case and against the restricted CUDA path in the
closed-shell limit.  Two kernel-library contexts (one per spin channel) run on the same GPU."""
import numpy as np
import pytest

from oracle import bse as obse
from oracle import uks
from tests.helpers import load_golden, methane_integrals, methane_mmn, uks_case
from tests.test_oracle_golden import _gw_options

pytestmark = pytest.mark.gpu

GRID = dict(gw__qp_grid_steps=601, gw__qp_grid_spacing=0.005)


def make_job(c, mode="G0W0", tasks="gw", **opts):
    from votca_b200.api import Job
    m = methane_integrals()
    job = Job(0)
    job.set_ao3c(m["ao3c"])
    job.set_array("aux_overlap", m["S"])
    job.set_array("aux_coulomb", m["V"])
    job.set_array("mos", c["Ca"])
    job.set_array("mos_beta", c["Cb"])
    job.set_array("mo_energies", c["ea"])
    job.set_array("mo_energies_beta", c["eb"])
    job.set_array("vxc", c["vxc_a"])
    job.set_array("vxc_beta", c["vxc_b"])
    job.set_scalar("homo", c["homo_a"])
    job.set_scalar("homo_beta", c["homo_b"])
    job.set_options(tasks=tasks, ranges="full", gw__mode=mode, gw__sigma_integrator="ppm", gw__mixing_order=0,
                    gw__qp_sc_limit=1e-5, gw__qp_sc_max_iter=50, gw__sc_limit=1e-5, bse__exctotal=3, bse__useTDA=True,
                    bse__use_Hqp_offdiag=True, bse__davidson__tolerance="lapack", **GRID)
    for k, v in opts.items():
        job.set_option(k.replace("__", "."), v)
    return job


def test_closed_shell_limit_equals_the_restricted_path():
    g = load_golden()
    c = {"Ca": g["gw/mo_eigenvectors"], "Cb": g["gw/mo_eigenvectors"], "ea": g["inline/gw_mo_eigenvalues"],
         "eb": g["inline/gw_mo_eigenvalues"], "vxc_a": g["gw/vxc"], "vxc_b": g["gw/vxc"], "homo_a": 4, "homo_b": 4}
    for mode in ("G0W0", "evGW"):
        u = make_job(c, mode)
        u.run_uks()
        r = make_job(c, mode)
        r.run()
        for s in ("_alpha", "_beta"):
            assert np.abs(u.get("QPpert_energies" + s) - r.get("QPpert_energies")).max() < 1e-9
            assert np.abs(u.get("Hqp" + s) - r.get("Hqp")).max() < 1e-9
            assert np.abs(u.get("RPA_inputenergies" + s) - r.get("RPA_inputenergies")).max() < 1e-9
        # the reference's own fixture of this restricted case (test_gw.cc: gw/ref.mm, 1e-4 relative)
        if mode == "G0W0":
            ref = np.diag(g["gw/ref"])
            assert np.abs(u.get("QPpert_energies_alpha") - ref).max() / np.abs(ref).max() < 1e-4
        u.close()
        r.close()


@pytest.mark.parametrize("mode", ["G0W0", "evGW"])
def test_open_shell_gw_and_excitons_against_the_oracle(mode):
    c = uks_case()
    iters = 1 if mode == "G0W0" else 50
    og = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    og.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005, gw_sc_max_iterations=iters), c["homo_a"],
                 c["homo_b"])
    og.calculate_gw_perturbation()
    og.calculate_hqp()
    job = make_job(c, mode, tasks="gw,exciton_uks")
    job.run_uks()
    for s, tag in enumerate(("_alpha", "_beta")):
        assert np.abs(job.get("QPpert_energies" + tag) - og.get_gwa_results(s)).max() < 1e-6  # Hartree
        assert np.abs(job.get("Hqp" + tag) - og.get_hqp(s)).max() < 1e-6
        assert np.abs(job.get("RPA_inputenergies" + tag) - og.rpa.energies(s)).max() < 1e-6
    assert np.abs(job.get("QPpert_energies_alpha") - job.get("QPpert_energies_beta")).max() > 1e-3
    ob = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=3, useTDA=True, homo=4, qpmin=0, qpmax=16,
                        davidson_tolerance="lapack", davidson_maxiter=50, use_Hqp_offdiag=True)
    ob.configure(o, c["homo_a"], c["homo_b"], og.rpa.energies(0), og.rpa.energies(1), og.get_hqp(0), og.get_hqp(1))
    w = np.linalg.eigvalsh(ob.operator_tda().dense())
    got = job.get("BSE_uks_eigenvalues")
    assert job.scalar("uks_converged") == 1.0
    assert (job.scalar("bse_alpha_size"), job.scalar("bse_beta_size")) == (5 * 12, 4 * 13)
    assert np.abs(got - w[:3]).max() < 1e-6
    X = job.get("BSE_uks_eigenvectors")
    assert X.shape == (112, 3) and np.abs(X.T @ X - np.eye(3)).max() < 1e-8
    job.close()


def test_unrestricted_operator_against_the_oracle_matrix():
    """BSE_OPERATOR_UKS<1,1,1,0> (gwbse_bse_matmul_dev per channel + gwbse_bse_vc_project_dev / _vc_expand_dev across
    them) with Hqp reduced to its diagonal: the ten lowest excitons against the oracle's dense operator."""
    c = uks_case()
    job = make_job(c, "G0W0", tasks="gw,exciton_uks", bse__exctotal=10, bse__use_Hqp_offdiag=False)
    job.run_uks()
    og = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    og.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005), c["homo_a"], c["homo_b"])
    og.calculate_gw_perturbation()
    og.calculate_hqp()
    ob = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=10, useTDA=True, homo=4, qpmin=0, qpmax=16,
                        use_Hqp_offdiag=False)
    ob.configure(o, c["homo_a"], c["homo_b"], og.rpa.energies(0), og.rpa.energies(1), og.get_hqp(0), og.get_hqp(1))
    w = np.linalg.eigvalsh(ob.operator_tda().dense())
    assert np.abs(np.sort(job.get("BSE_uks_eigenvalues")) - w[:10]).max() < 1e-6
    job.close()


def test_full_unrestricted_bse_against_the_oracle():
    """bse.useTDA=false: [A B; -B -A] with A = <1,1,1,0>, B = <0,1,0,1> (cross-spin Hd2 block through
    gwbse_bse_hd2_cross_dev); 112 excitations -> the dense branch, as in the reference."""
    c = uks_case()
    og = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    og.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005), c["homo_a"], c["homo_b"])
    og.calculate_gw_perturbation()
    og.calculate_hqp()
    ob = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=5, useTDA=False, homo=4, qpmin=0, qpmax=16,
                        use_Hqp_offdiag=True)
    ob.configure(o, c["homo_a"], c["homo_b"], og.rpa.energies(0), og.rpa.energies(1), og.get_hqp(0), og.get_hqp(1))
    ref = ob.solve_excitons_uks_btda_dense()
    job = make_job(c, "G0W0", tasks="gw,exciton_uks", bse__useTDA=False, bse__exctotal=5)
    job.run_uks()
    assert np.abs(job.get("BSE_uks_eigenvalues") - ref["eigenvalues"]).max() < 1e-6
    X, Y = job.get("BSE_uks_eigenvectors"), job.get("BSE_uks_eigenvectors2")
    assert np.abs(np.abs(np.sum(X * X, axis=0) - np.sum(Y * Y, axis=0)) - 1.0).max() < 1e-8  # 1 / sqrt(abs(norm))
    for r in range(5):  # same state up to the degeneracy-free phase convention
        if r + 1 < 5 and abs(ref["eigenvalues"][r + 1] - ref["eigenvalues"][r]) < 1e-6:
            continue
        if r > 0 and abs(ref["eigenvalues"][r] - ref["eigenvalues"][r - 1]) < 1e-6:
            continue
        assert abs(abs(X[:, r] @ ref["eigenvectors"][:, r] - Y[:, r] @ ref["eigenvectors2"][:, r]) - 1.0) < 1e-5
    job.close()


@pytest.mark.parametrize("integrator", ["exact", "cda"])
def test_open_shell_gw_with_the_other_integrators_against_the_oracle(integrator):
    """sigma_integrator = exact for an unrestricted reference: RPA_UKS::Diagonalize_H2p over both channels' tensors
    (gwbse_rpa_h2p_block), the shared screening modes (gwbse_sigma_exact_project on either context) and the
    per-channel residues (gwbse_sigma_exact_prepare_modes); sigma_integrator = cda: the device-resident contour
    deformation of a channel with the dielectric matrix of both (gwbse_sigma_cda_set_partner).  Against
    oracle/uks.py; evGW so that the screening is rebuilt with updated energies of both channels."""
    c = uks_case()
    # the contour-deformation oracle assembles eps(z) per pole and evaluation in NumPy: a coarser scan and two
    # self-consistency steps keep it to seconds
    steps, spacing, iters = (601, 0.005, 3) if integrator == "exact" else (151, 0.02, 2)
    og = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    og.configure(_gw_options(qp_grid_steps=steps, qp_grid_spacing=spacing, gw_sc_max_iterations=iters,
                             sigma_integration=integrator, order=12, alpha=1e-3), c["homo_a"], c["homo_b"])
    og.calculate_gw_perturbation()
    og.calculate_hqp()
    job = make_job(c, "evGW", tasks="gw", gw__sigma_integrator=integrator, gw__sc_max_iter=iters,
                   gw__quadrature_order=12, gw__alpha=1e-3, gw__qp_grid_steps=steps, gw__qp_grid_spacing=spacing)
    job.run_uks()
    for s, tag in enumerate(("_alpha", "_beta")):
        assert np.abs(job.get("QPpert_energies" + tag) - og.get_gwa_results(s)).max() < 1e-6  # Hartree
        assert np.abs(job.get("Hqp" + tag) - og.get_hqp(s)).max() < 1e-6
        assert np.abs(job.get("RPA_inputenergies" + tag) - og.rpa.energies(s)).max() < 1e-6
    assert np.abs(job.get("QPpert_energies_alpha") - job.get("QPpert_energies_beta")).max() > 1e-3
    job.close()


def test_closed_shell_limit_of_the_exact_integrator_equals_the_restricted_path():
    g = load_golden()
    c = {"Ca": g["gw/mo_eigenvectors"], "Cb": g["gw/mo_eigenvectors"], "ea": g["inline/gw_mo_eigenvalues"],
         "eb": g["inline/gw_mo_eigenvalues"], "vxc_a": g["gw/vxc"], "vxc_b": g["gw/vxc"], "homo_a": 4, "homo_b": 4}
    u = make_job(c, "G0W0", gw__sigma_integrator="exact")
    u.run_uks()
    r = make_job(c, "G0W0", gw__sigma_integrator="exact")
    r.run()
    for s in ("_alpha", "_beta"):
        assert np.abs(u.get("QPpert_energies" + s) - r.get("QPpert_energies")).max() < 1e-8
        assert np.abs(u.get("Hqp" + s) - r.get("Hqp")).max() < 1e-8
    u.close()
    r.close()


def test_what_is_not_on_this_path_is_refused():
    c = uks_case()
    for kw, msg in ((dict(tasks="gw,singlets"), "not defined for open-shell"),
                    (dict(gw__mode="evGW", gw__do_qsgw=True), "restricted references only")):
        job = make_job(c, **kw)
        with pytest.raises(Exception, match=msg):
            job.run_uks()
        job.close()
    job = make_job(c, tasks="gw,exciton_uks")
    with pytest.raises(Exception, match="unrestricted reference"):
        job.run()  # the restricted driver refuses the unrestricted task
    job.close()


@pytest.mark.parametrize("tda", [True, False])
def test_unrestricted_dynamical_screening_against_the_oracle(tda):
    """bse.dyn_screen_max_iter > 0 for an unrestricted reference (BSE_UKS::Perturbative_DynamicalScreening,
    bse_uks.cc:640-702): the spin-summed eps(omega) re-diagonalised per excitation and iteration, expectation values
    of the unrestricted Hd operator <0,0,1,0> and, without the TDA, the cross term of Hd2 <0,0,0,1>."""
    c = uks_case()
    og = uks.GWUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]), c["vxc_a"], c["vxc_b"], c["ea"], c["eb"])
    og.configure(_gw_options(qp_grid_steps=601, qp_grid_spacing=0.005), c["homo_a"], c["homo_b"])
    og.calculate_gw_perturbation()
    og.calculate_hqp()
    job = make_job(c, "G0W0", tasks="gw,exciton_uks", bse__useTDA=tda, bse__exctotal=3, bse__dyn_screen_max_iter=4,
                   bse__dyn_screen_tol=1e-6)
    job.run_uks()
    es = {"eigenvalues": job.get("BSE_uks_eigenvalues").ravel(), "eigenvectors": job.get("BSE_uks_eigenvectors")}
    if not tda:
        es["eigenvectors2"] = job.get("BSE_uks_eigenvectors2")
    ob = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    o = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=3, useTDA=tda, homo=4, qpmin=0, qpmax=16,
                        use_Hqp_offdiag=True, max_dyn_iter=4, dyn_tolerance=1e-6)
    ob.configure(o, c["homo_a"], c["homo_b"], og.rpa.energies(0), og.rpa.energies(1), og.get_hqp(0), og.get_hqp(1))
    ref = ob.perturbative_dynamical_screening(es, og.rpa.energies(0), og.rpa.energies(1))
    got = job.get("BSE_uks_dynamic").ravel()
    assert got.shape == ref.shape == (3,)
    assert np.abs(got - ref).max() < 1e-6  # Hartree
    shift = got - es["eigenvalues"]
    assert np.all(np.abs(shift) > 1e-7) and np.all(np.abs(shift) < 0.05)  # a correction, and a small one
    job.close()


@pytest.mark.parametrize("tda", [True, False])
def test_unrestricted_transition_dipoles_and_oscillator_strengths(tda):
    """Orbitals::CalcCoupledTransition_Dipoles / Oscillatorstrengths for combined UKS excitons (orbitals.cc:798-877,
    645-674): interlevel dipoles of both spin channels formed on the device from the AO dipole matrices, no sqrt(2);
    oscillator strengths within 1e-5 of the oracle."""
    from oracle import integrals
    from tests.helpers import methane_integrals
    c = uks_case()
    dip = integrals.dipole(methane_integrals()["basis"])
    job = make_job(c, "G0W0", tasks="gw,exciton_uks", bse__useTDA=tda, bse__exctotal=4)
    for k, name in enumerate(("ao_dipole_x", "ao_dipole_y", "ao_dipole_z")):
        job.set_array(name, dip[k])
    job.run_uks()
    es = {"eigenvalues": job.get("BSE_uks_eigenvalues").ravel(), "eigenvectors": job.get("BSE_uks_eigenvectors")}
    if not tda:
        es["eigenvectors2"] = job.get("BSE_uks_eigenvectors2")
    ob = uks.BSEUKS(methane_mmn(c["Ca"]), methane_mmn(c["Cb"]))
    ob.opt = obse.BSEOptions(cmax=16, rpamax=16, rpamin=0, vmin=0, nmax=4, useTDA=tda, homo=4, qpmin=0, qpmax=16)
    ob.homo = (c["homo_a"], c["homo_b"])
    d, f = ob.transition_dipoles(es, dip, c["Ca"], c["Cb"])
    assert np.abs(job.get("uks_transition_dipoles").T - d).max() < 1e-8
    assert np.abs(job.get("uks_oscillator_strengths").ravel() - f).max() < 1e-8
    assert np.abs(f).max() > 1e-4  # (the toy doublet has excitations of negative energy: f carries their sign)
    log = job.log()
    assert "TrDipole length gauge[e*bohr]" in log and "XU1" in log
    assert "           alpha-sector: " in log and "beta-sector: " in log  # BSE_UKS::PrintWeightsUKS
    job.close()


def test_unrestricted_summary_xml(tmp_path):
    """<job>_summary.xml of an unrestricted run as GWBSE::addoutput writes it (gwbse.cc:591-633, 696-736): dft_alpha /
    dft_beta level tables and the exciton_uks levels with omega, f and Trdipole, in eV."""
    import xml.etree.ElementTree as ET
    from oracle import integrals
    from tests.helpers import methane_integrals
    c = uks_case()
    dip = integrals.dipole(methane_integrals()["basis"])
    job = make_job(c, "G0W0", tasks="gw,exciton_uks", bse__exctotal=3)
    for k, name in enumerate(("ao_dipole_x", "ao_dipole_y", "ao_dipole_z")):
        job.set_array(name, dip[k])
    out = tmp_path / "uks_summary.xml"
    job.set_summary_output(str(out))
    job.run_uks()
    g = ET.parse(out).getroot().find("GWBSE")
    assert g.get("units") == "eV"
    h2e = 27.21138602
    for tag, homo in (("alpha", c["homo_a"]), ("beta", c["homo_b"])):
        t = g.find("dft_" + tag)
        assert int(t.get("HOMO")) == homo and int(t.get("LUMO")) == homo + 1
        lv = t.findall("level")
        assert len(lv) == 17 and [int(x.get("number")) for x in lv] == list(range(17))
        gw = np.array([float(x.find("gw_energy").text) for x in lv])
        assert np.abs(gw - job.get("QPpert_energies_" + tag) * h2e).max() < 1e-6
        qp = np.array([float(x.find("qp_energy").text) for x in lv])
        assert np.abs(qp - job.get("QPdiag_eigenvalues_" + tag).ravel() * h2e).max() < 1e-6
    ex = g.find("exciton_uks").findall("level")
    assert [int(x.get("number")) for x in ex] == [1, 2, 3]
    om = np.array([float(x.find("omega").text) for x in ex])
    assert np.abs(om - job.get("BSE_uks_eigenvalues").ravel() * h2e).max() < 1e-6
    f = np.array([float(x.find("f").text) for x in ex])
    assert np.abs(f - job.get("uks_oscillator_strengths").ravel()).max() < 1e-6
    td = ex[0].find("Trdipole")
    assert td.get("unit") == "e*bohr" and td.get("gauge") == "length" and len(td.text.split()) == 3
    job.close()
