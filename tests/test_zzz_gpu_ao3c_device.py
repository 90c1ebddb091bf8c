"""GPU parity of the device AO-integral producer through the C ABI: gwbse_ao3c_block,
gwbse_ao_coulomb2c and gwbse_mmn_fill_from_basis against the oracle's integrals / Mmn.

STATUS: written after this round's GPU budget was spent - the kernel's arithmetic is verified on the CPU
(tests/test_ao3c_core_cpu.py runs the same __host__ __device__ source, incl. ThreadSanitizer on the barriers), the
launch path itself has not executed on a device yet.  The file sorts last so that `pytest -x` reaches every
verified test first.

Replaces: ComputeAO3cBlock (libint2_calls.cc:544-593), AOCoulomb::Fill (libint2_calls.cc:224-271), and the host
producer side of TCMatrix_gwbse::Fill3cMO (libint2_calls.cc:595-651)."""
import numpy as np
import pytest

from oracle import threecenter
from tests import helpers
from tests.test_ao3c_core_cpu import _golden_basis, pack, relmax

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from votca_b200.api import Context
    c = Context(0)
    yield c
    c.close()


def _device_basis(ctx, ao):
    return ctx.basis_create(*pack(ao))


def test_ao3c_water_spdf_aux(ctx):
    w = helpers.water_integrals()
    aux, dft = _device_basis(ctx, w["aux"]), _device_basis(ctx, w["dft"])
    assert ctx.basis_size(aux) == w["aux"].size and ctx.basis_size(dft) == w["dft"].size
    before = ctx.launch_count()
    got = ctx.ao3c_block(aux, dft, 0, w["aux"].size)
    assert ctx.launch_count() > before
    assert relmax(w["ao3c"], got) < 1e-12
    # function ranges that cut through shells, and an empty one
    for lo, cnt in ((3, 7), (0, 1), (w["aux"].size - 2, 2), (5, 0)):
        part = ctx.ao3c_block(aux, dft, lo, cnt)
        assert part.shape[0] == cnt
        if cnt:
            assert relmax(w["ao3c"][lo:lo + cnt], part) < 1e-12
    assert relmax(w["V"], ctx.ao_coulomb2c(aux)) < 1e-12
    assert relmax(w["S"], ctx.ao_overlap(aux)) < 1e-13
    assert relmax(w["S_dft"], ctx.ao_overlap(dft)) < 1e-13
    assert relmax(w["dipole"], ctx.ao_dipole(dft)) < 1e-13
    # odd basis size (N = 13): fill_from_basis pads the device blocks to an even pitch
    N = w["dft"].size
    C = np.linalg.qr(np.random.default_rng(2).standard_normal((N, N)))[0]
    ctx.mmn_alloc(w["aux"].size, 0, N - 1, 0, N - 1)
    ctx.mmn_set_mos(C)
    ctx.mmn_fill_from_basis(aux, dft, aux_block=10)
    tc = threecenter.TCMatrix(w["aux"].size, 0, N - 1, 0, N - 1)
    tc.fill_3c_mo(w["ao3c"], C)
    assert helpers.rel_frob(tc.M, ctx.mmn_get_all()) < 1e-10
    ctx.basis_destroy(aux)
    ctx.basis_destroy(dft)


def test_ao3c_methane_def2svp_and_fill_from_basis(ctx):
    c = helpers.methane_svp_case()
    aux, dft = _device_basis(ctx, c["aux"]), _device_basis(ctx, c["dft"])
    got = ctx.ao3c_block(aux, dft, 0, c["aux"].size)
    assert relmax(c["ao3c"], got) < 1e-12
    assert relmax(c["V"], ctx.ao_coulomb2c(aux)) < 1e-12
    # Fill3cMO with the producer on the device == oracle Fill3cMO on the oracle's integrals (Mmn bar 1e-10)
    C = c["hf"]["mos"]
    N, q = c["dft"].size, c["q"]
    ctx.mmn_alloc(c["aux"].size, 0, q - 1, 0, N - 1)
    ctx.mmn_set_mos(C)
    ctx.mmn_fill_from_basis(aux, dft, aux_block=17)
    tc = threecenter.TCMatrix(c["aux"].size, 0, q - 1, 0, N - 1)
    tc.fill_3c_mo(c["ao3c"], C)
    assert helpers.rel_frob(tc.M, ctx.mmn_get_all()) < 1e-10
    ctx.basis_destroy(aux)
    ctx.basis_destroy(dft)


def test_ao3c_large_l_g_orbitals_i_aux(ctx):
    from oracle import integrals
    aux_ao, dft_ao = _golden_basis("I", "C2"), _golden_basis("G", "C2")
    aux, dft = _device_basis(ctx, aux_ao), _device_basis(ctx, dft_ao)
    assert relmax(integrals.coulomb3c(aux_ao, dft_ao), ctx.ao3c_block(aux, dft, 0, aux_ao.size)) < 1e-11
    ctx.basis_destroy(aux)
    ctx.basis_destroy(dft)


def test_job_from_basis_sets_equals_job_from_arrays():
    """The whole GW-BSE job with the integral producer on the device (Job.set_basis) against the same job fed with
    the oracle's integral arrays: QP / BSE energies within 1e-9 Ha."""
    from votca_b200.api import Job
    c = helpers.methane_svp_case()
    hf = c["hf"]
    N, q, homo = c["dft"].size, c["q"], c["homo"]
    vxc = hf["exchange_mo"][:q, :q]
    res = []
    for from_basis in (False, True):
        job = Job(0)
        job.set_scalar("homo", homo)
        job.set_array("mos", hf["mos"])
        job.set_array("mo_energies", hf["energies"])
        job.set_array("vxc", vxc)
        if from_basis:  # overlap, two- and three-centre Coulomb integrals all come from the device
            job.set_basis("dft", *pack(c["dft"]))
            job.set_basis("aux", *pack(c["aux"]))
        else:
            job.set_ao3c(c["ao3c"])
            job.set_array("aux_overlap", c["S"])
            job.set_array("aux_coulomb", c["V"])
            ct, vt = q - homo - 1, homo + 1
            C = hf["mos"]
            for ax, D in zip("xyz", c["dipole"]):
                job.set_array("dipole_" + ax, C[:, homo + 1:homo + 1 + ct].T @ D @ C[:, :vt])
        job.set_options(tasks="gw,singlets", gw__mode="G0W0", gw__sigma_integrator="ppm", bse__exctotal=5,
                        bse__useTDA=True)
        job.run()
        res.append((job.get("QPpert_energies").ravel().copy(), job.get("BSE_singlet_eigenvalues").ravel().copy(),
                    job.get("oscillator_strengths").ravel().copy()))
        job.close()
    assert np.abs(res[0][0] - res[1][0]).max() < 1e-9
    assert np.abs(res[0][1] - res[1][1]).max() < 1e-9
    # oscillator strengths: interlevel dipoles given (oracle AO dipoles) vs AO dipoles produced on the device; the
    # lowest singlets of methane are degenerate (T2), so compare the sum over the shell
    assert res[1][2].size == 5 and abs(res[0][2].sum() - res[1][2].sum()) < 1e-7


def test_basis_errors(ctx):
    from votca_b200.api import GwbseError
    with pytest.raises(GwbseError, match="angular momentum"):
        ctx.basis_create([7], [1], [[0.0, 0.0, 0.0]], [1.0], [1.0])
    with pytest.raises(GwbseError, match="exponent"):
        ctx.basis_create([0], [1], [[0.0, 0.0, 0.0]], [0.0], [1.0])
