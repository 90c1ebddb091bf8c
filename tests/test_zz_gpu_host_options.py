"""Less-travelled options of the C++ host layer against the oracle on identical inputs (methane / 3-21G of the
reference's unit tests): fixed-point QP solver, scissor shift, linear mixing, Rebuild of the three-centre tensor every
iteration, explicit level ranges with a BSE window that sticks out of the QP window on both sides (AdjustHqpSize,
bse.cc:147-186), Olsen correction.  Runs on the GPU like tests/test_gpu_host.py, and in the CPU suite against the
test-only mock (tests/test_host_layer_on_mock_cpu.py)."""
import numpy as np
import pytest

from oracle import bse as obse
from oracle import gw as ogw
from tests.helpers import methane_mmn
from tests.test_gpu_host import GW_OPTS, make_job

pytestmark = pytest.mark.gpu


def _oracle_gw(golden, mmax=16, **kw):
    tc = methane_mmn(golden["gw/mo_eigenvectors"], mmax=mmax)
    opt = dict(homo=4, qpmin=0, qpmax=16, rpamin=0, rpamax=16, gw_sc_max_iterations=1, eta=1e-3,
               sigma_integration="ppm", gw_mixing_order=0, g_sc_limit=1e-5, g_sc_max_iterations=50)
    opt.update(kw)
    q0, q1 = opt["qpmin"], opt["qpmax"]
    g = ogw.GW(tc, golden["gw/vxc"][q0:q1 + 1, q0:q1 + 1], golden["inline/gw_mo_eigenvalues"])
    g.configure(ogw.GWOptions(**opt))
    g.calculate_gw_perturbation()
    return tc, g


def _job(golden, methane, **options):
    opts = dict(GW_OPTS)
    opts.update(options)
    job = make_job(methane, golden["gw/mo_eigenvectors"], golden["inline/gw_mo_eigenvalues"], **opts)
    job.set_array("vxc", golden["gw/vxc"])
    return job


def test_fixedpoint_qp_solver(golden, methane):
    job = _job(golden, methane, gw__qp_solver="fixedpoint")
    job.run()
    _, g = _oracle_gw(golden, qp_solver="fixedpoint")
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies")).max() < 1e-6
    job.close()


def test_scissor_shift_linear_mixing_and_rebuild_every_iteration(golden, methane):
    res = {}
    for rebuild in (1, 5):
        job = _job(golden, methane, gw__mode="evGW", gw__sc_max_iter=3, gw__scissor_shift=0.05, gw__mixing_order=1,
                   gw__mixing_alpha=0.6, gw__rebuild_3c_freq=rebuild)
        job.run()
        res[rebuild] = (job.get("QPpert_energies").copy(), job.get("RPA_inputenergies").copy(), job.scalar("gw_iterations"))
        job.close()
    _, g = _oracle_gw(golden, gw_sc_max_iterations=3, shift=0.05, gw_mixing_order=1, gw_mixing_alpha=0.6, reset_3c=5)
    assert np.abs(g.get_gwa_results() - res[5][0]).max() < 1e-6
    assert np.abs(g.rpa_input_energies() - res[5][1]).max() < 1e-6 and res[5][2] == g.iterations
    # rebuilding the tensor from the integrals (or the device snapshot) every iteration changes nothing
    assert np.abs(res[1][0] - res[5][0]).max() < 1e-9


def test_bse_window_beyond_the_qp_window(golden, methane):
    """ranges = explicit: QP window 1..12, BSE window 0..14 - Hqp is padded with RPA input energies on both sides."""
    job = _job(golden, methane, tasks="gw,singlets,triplets", bse__useTDA=True, bse__exctotal=3,
               bse__davidson__tolerance="strict", bse__davidson__correction="OLSEN", bse__use_Hqp_offdiag=True)
    job.set_options(ranges="explicit", rpamax=16, qpmin=1, qpmax=12, bsemin=0, bsemax=14)  # overrides make_job's "full"
    job.set_array("vxc", golden["gw/vxc"][1:13, 1:13])
    job.run()
    assert (job.scalar("qpmin"), job.scalar("qpmax"), job.scalar("bse_vmin"), job.scalar("bse_cmax")) == (1, 12, 0, 14)
    tc, g = _oracle_gw(golden, mmax=14, qpmin=1, qpmax=12)
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies")).max() < 1e-6
    g.calculate_hqp()
    b = obse.BSE(tc, factorised=True)
    b.configure(obse.BSEOptions(useTDA=True, homo=4, rpamin=0, rpamax=16, qpmin=1, qpmax=12, vmin=0, cmax=14, nmax=3,
                                davidson_tolerance="strict", davidson_correction="OLSEN", use_Hqp_offdiag=True),
                g.rpa_input_energies(), g.get_hqp())
    assert np.abs(b.solve_singlets()["eigenvalues"] - job.get("BSE_singlet_eigenvalues")).max() < 1e-6
    assert np.abs(b.solve_triplets()["eigenvalues"] - job.get("BSE_triplet_eigenvalues")).max() < 1e-6
    job.close()


def test_integral_producer_callback(golden, methane):
    """The path a maintainer's libint loop takes: TCMatrix_gwbse::Fill3cMO asks an AOIntegralSource
    for blocks of aux functions (here a callback, gwbse_job_set_ao3c_callback), stages them in page-locked memory and
    contracts them; same result as handing the whole tensor over."""
    from votca_b200.api import Job
    res, calls = [], []
    for use_callback in (False, True):
        job = Job(0)
        job.set_scalar("homo", 4)
        job.set_array("mos", golden["gw/mo_eigenvectors"])
        job.set_array("mo_energies", golden["inline/gw_mo_eigenvalues"])
        job.set_array("vxc", golden["gw/vxc"])
        job.set_array("aux_overlap", methane["S"])
        job.set_array("aux_coulomb", methane["V"])
        if use_callback:
            def producer(off, cnt):
                calls.append((off, cnt))
                return methane["ao3c"][off:off + cnt]
            job.set_ao3c_callback(17, 17, producer)
        else:
            job.set_ao3c(methane["ao3c"])
        job.set_options(ranges="full", **GW_OPTS)
        job.run()
        res.append(job.get("QPpert_energies").copy())
        job.close()
    assert sum(c for _, c in calls) == 17 and [o for o, _ in calls] == sorted(o for o, _ in calls)
    assert np.abs(res[0] - res[1]).max() < 1e-12


def test_switching_the_integral_producer_on_one_job(golden, methane):
    """One Job, three producers in a row - a partial host range, then the callback, then the whole array: a later
    producer must not be shadowed by state of an earlier one (a stale device pointer or 'held' range once made the
    callback silently unused)."""
    from votca_b200.api import Job
    job = Job(0)
    job.set_scalar("homo", 4)
    job.set_array("mos", golden["gw/mo_eigenvectors"])
    job.set_array("mo_energies", golden["inline/gw_mo_eigenvalues"])
    job.set_array("vxc", golden["gw/vxc"])
    job.set_array("aux_overlap", methane["S"])
    job.set_array("aux_coulomb", methane["V"])
    job.set_options(ranges="full", **GW_OPTS)
    whole = np.ascontiguousarray(methane["ao3c"])
    job.set_ao3c_partial(17, 17, 0, 17, whole.ctypes.data, False)
    job.run()
    first = job.get("QPpert_energies").copy()
    calls = []

    def producer(off, cnt):
        calls.append((off, cnt))
        return 2.0 * whole[off:off + cnt]  # deliberately different integrals: the callback has to be what is used

    job.set_ao3c_callback(17, 17, producer)
    job.run()
    second = job.get("QPpert_energies").copy()
    assert sum(c for _, c in calls) == 17
    assert np.abs(second - first).max() > 1e-3
    job.set_ao3c(whole)
    job.run()
    assert np.abs(job.get("QPpert_energies") - first).max() < 1e-12
    job.close()


def test_sigma_plot(golden, methane, tmp_path):
    """gw.sigma_plot (GW::PlotSigma, gw.cc:1132-1182): E_QP(omega) of the chosen states on a frequency grid, one
    grouped evaluation on the device, written in the reference's table format; states outside the qp window are
    dropped, ranges and duplicates are handled as IndexParser does."""
    out = tmp_path / "sigma.dat"
    job = _job(golden, methane, gw__sigma_plot__states="3:5 2,5 40", gw__sigma_plot__steps=21,
               gw__sigma_plot__spacing=0.02, gw__sigma_plot__filename=str(out))
    job.run()
    _, g = _oracle_gw(golden)
    levels, table = g.plot_sigma(21, 0.02, "3:5 2,5 40")
    assert levels == [2, 3, 4, 5]
    lines = out.read_text().splitlines()
    assert lines[0] == "#omega_2\tE_QP(omega)_2#\tomega_3\tE_QP(omega)_3#\tomega_4\tE_QP(omega)_4#\tomega_5\tE_QP(omega)_5"
    got = np.array([[float(x) for x in ln.split("\t")] for ln in lines[1:] if ln.strip()])
    assert got.shape == table.shape
    assert all(tok[0] in "+-" and len(tok.split(".")[1]) == 6 for tok in lines[1].split("\t"))  # "%+1.6f"
    assert np.abs(got - table).max() < 2e-6
    assert np.abs(g.get_gwa_results() - job.get("QPpert_energies")).max() < 1e-6
    # GW::PrintGWA_Energies (gw.cc:80-111)
    log = job.log()
    assert "  ====== Perturbative quasiparticle energies (Hartree) ====== " in log and "   DeltaHLGap = " in log
    qp = job.get("QPpert_energies")
    assert ("  HOMO  =    4 DFT = %+1.4f VXC = " % golden["inline/gw_mo_eigenvalues"][4]) in log
    assert ("GWA = %+1.4f" % qp[4]) in log and "  LUMO  =    5 DFT = " in log and "  Level =    0 DFT = " in log
    job.close()


@pytest.mark.parametrize("tda", [True, False])
def test_fragment_populations(golden, methane, tda, tmp_path):
    """bse.fragments: Lowdin populations of every exciton on groups of atoms (Lowdin::CalcChargeperFragment,
    populationanalysis.cc:47-84, with the densities of orbitals.cc:516-650) against oracle/population.py, which
    reproduces the known answers of the reference's test_populationanalysis.cc; TDA and full BSE (antiresonant part
    subtracted), singlets and triplets; fragments given as two <fragment> elements of an options file."""
    from oracle import population as opop
    from tests.helpers import methane_integrals
    m = methane_integrals()
    basis_atom = np.concatenate([[sh.atom] * (2 * sh.l + 1) for sh in m["basis"].shells]).astype(float)
    nuc = np.array([6.0, 1.0, 1.0, 1.0, 1.0])
    mos = golden["gw/mo_eigenvectors"]
    S_dft = np.linalg.inv(mos @ mos.T)  # the MOs of the fixture are orthonormal in the metric they were computed with
    xml = tmp_path / "opts.xml"
    xml.write_text("<options><gwbse><bse><fragments><fragment><indices>0</indices></fragment>"
                   "<fragment><indices>1:3 4</indices></fragment></fragments></bse></gwbse></options>")
    job = _job(golden, methane, tasks="gw,singlets,triplets", bse__exctotal=4, bse__useTDA=tda,
               bse__davidson__tolerance="lapack")
    job.load_options_xml(str(xml))
    job.set_array("ao_overlap", S_dft)
    job.set_array("basis_atom_index", basis_atom)
    job.set_array("nuclear_charges", nuc)
    job.run()
    for kind in ("singlet", "triplet"):
        X = job.get(f"BSE_{kind}_eigenvectors")
        Y = None if tda else job.get(f"BSE_{kind}_eigenvectors2")
        Gs, H, E = opop.fragment_populations(S_dft, mos, 4, 0, 16, basis_atom.astype(int), nuc, [[0], [1, 2, 3, 4]], X, Y)
        assert np.abs(job.get("fragment_gs").ravel() - Gs).max() < 1e-9
        assert np.abs(job.get(f"BSE_{kind}_fragment_hole") - H).max() < 1e-9
        assert np.abs(job.get(f"BSE_{kind}_fragment_electron") - E).max() < 1e-9
        # the exciton is neutral and normalised: hole and electron populations sum to +1 / -1 (TDA)
        if tda:
            assert np.abs(H.sum(axis=0) - 1.0).max() < 1e-8 and np.abs(E.sum(axis=0) + 1.0).max() < 1e-8
    assert abs(Gs.sum()) < 1e-8  # neutral molecule
    log = job.log()
    assert "Fragment    1 -- hole:" in log
    # the per-state report of BSE::Analyze_singlets / Analyze_triplets (bse.cc:394-490)
    assert "  ====== singlet energies (eV) ====== " in log and "  ====== triplet energies (eV) ====== " in log
    e1 = job.get("BSE_singlet_eigenvalues")[0] * 27.21138602
    assert ("   S =    1 Omega = %+1.12f eV  lamdba = %+3.2f nm <FT> = " % (e1, 1240.0 / e1)) in log
    assert "   T =    1 Omega = " in log and "<K_x>" in log and "HOMO-" in log and " -> LUMO+" in log
    job.close()


def test_ignore_corelevels_gw(golden, methane):
    """ignore_corelevels = GW: the carbon 1s level leaves the QP and BSE windows (one core level for CH4) - the same
    run as explicit ranges with qpmin = bsemin = 1."""
    res = {}
    for tag, kw in (("core", dict(ranges="full", ignore_corelevels="GW")),
                    ("explicit", dict(ranges="explicit", rpamax=16, qpmin=1, qpmax=16, bsemin=1, bsemax=16))):
        job = _job(golden, methane, tasks="gw,singlets", bse__exctotal=3, bse__useTDA=True)
        job.set_options(**kw)
        job.set_array("nuclear_charges", np.array([6.0, 1.0, 1.0, 1.0, 1.0]))
        job.run()
        res[tag] = (job.get("QPpert_energies").copy(), job.get("BSE_singlet_eigenvalues").copy(),
                    job.scalar("qpmin"), job.scalar("bse_vmin"), job.scalar("rpamin"))
        if tag == "core":
            assert "Ignoring 1 core levels for GW and beyond." in job.log()
        job.close()
    assert res["core"][2:] == (1.0, 1.0, 0.0) == res["explicit"][2:]
    assert np.abs(res["core"][0] - res["explicit"][0]).max() < 1e-12
    assert np.abs(res["core"][1] - res["explicit"][1]).max() < 1e-10
