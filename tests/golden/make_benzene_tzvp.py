#!/usr/bin/env python
"""Generates tests/golden/benzene_tzvp_evgw_exact.npz: BASELINE.json config 2 (benzene, def2-tzvp + aux-def2-tzvp,
evGW with the exact self-energy integrator, full BSE 10 singlets) computed by the CPU oracle on own integrals.

The reference's DFT cannot run here, so the orbitals are RI-RHF orbitals of oracle/scf.py ("tier R");
everything downstream - Mmn, RPA two-particle Hamiltonian, residues, QP search, evGW loop, BSE, oscillator
strengths - is the oracle's restatement of rpa.cc:204-326, sigma_exact.cc:29-148, gw.cc:218-776, bse.cc:266-360.
The fixture holds the inputs the GPU test needs (MOs, energies, exchange matrix) and the oracle's outputs.

The AO integrals come from tests/host_harness/ao3c_host.cc (the McMurchie-Davidson code of the oracle compiled with
g++; tests/test_ao3c_core_cpu.py holds it to 1e-12 of oracle/integrals.py on this very system): the NumPy version
needs about an hour for the 27 M three-centre integrals of this basis.

  python tests/golden/make_benzene_tzvp.py          (about ten minutes on 8 cores)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import basis as obasis  # noqa: E402
from oracle import bse as obse  # noqa: E402
from oracle import gw as ogw  # noqa: E402
from oracle import integrals, scf, threecenter  # noqa: E402
from votca_b200 import realsys  # noqa: E402


def oracle_basis(name, elements, positions):
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells]
          for el, shells in realsys.basis_set(name).items()}
    return obasis.AOBasis(bs, elements, positions)


def use_compiled_integrals():
    from tests import test_ao3c_core_cpu as hh
    lib = hh._bind(hh._build("libao3c_host.so", ["-O2"]))
    integrals.coulomb3c = lambda aux, dft: hh.ao3c(lib, aux, dft)
    integrals.coulomb2c = lambda ao: hh.coulomb2c(lib, ao)
    integrals.overlap = lambda ao: hh.overlap(lib, ao)
    integrals.dipole = lambda ao: hh.dipole(lib, ao)


def main():
    t0 = time.time()
    use_compiled_integrals()
    el, pos = realsys.benzene()
    dft = oracle_basis("def2-tzvp", el, pos)
    aux = oracle_basis("aux-def2-tzvp", el, pos)
    Z = [{"H": 1, "C": 6}[e] for e in el]
    print("N", dft.size, "Naux", aux.size, flush=True)
    hf = scf.rhf_ri(dft, aux, Z, pos, sum(Z))
    print("RHF", hf["total_energy"], hf["iterations"], "iterations", time.time() - t0, "s", flush=True)
    homo = sum(Z) // 2 - 1
    N = dft.size
    q = min(3 * homo + 1, N - 1) + 1
    e, C = hf["energies"], hf["mos"]
    vxc = hf["exchange_mo"][:q, :q]
    S, V = integrals.overlap(aux), integrals.coulomb2c(aux)
    ao3c = integrals.coulomb3c(aux, dft)
    dip = integrals.dipole(dft)
    print("integrals", time.time() - t0, "s", flush=True)
    tc = threecenter.TCMatrix(aux.size, 0, q - 1, 0, N - 1)
    tc.fill_from_integrals(ao3c, S, V, C)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "benzene_tzvp_evgw_exact.npz")
    if "--bse-only" in sys.argv:  # keep the (18 minute) evGW results of the existing fixture, redo the BSE part
        with np.load(out) as z:
            old = {k: z[k] for k in z.files}
        assert np.array_equal(old["mos"], C) and np.array_equal(old["mo_energies"], e)

        class G:
            iterations = int(old["gw_iterations"]) if "gw_iterations" in old else 8
            def get_gwa_results(self): return old["QPpert_energies"]
            def rpa_input_energies(self): return old["RPA_inputenergies"]
            def get_hqp(self): return old["Hqp"]
        g = G()
    else:
        g = ogw.GW(tc, vxc, e)
        g.configure(ogw.GWOptions(homo=homo, qpmin=0, qpmax=q - 1, rpamin=0, rpamax=N - 1, gw_sc_max_iterations=50,
                                  sigma_integration="exact", g_sc_max_iterations=100))
        g.calculate_gw_perturbation()
        g.calculate_hqp()
    print("evGW", getattr(g, "iterations", None), time.time() - t0, "s", flush=True)
    vt, ct = homo + 1, q - homo - 1
    b = obse.BSE(tc, factorised=True)
    # Davidson tolerance "lapack" (1e-9 on the residuals): benzene's E1u / E2g levels are degenerate pairs, at the
    # default tolerance the two partners differ by 1e-5 Ha, more than the 1e-6 Ha parity bar
    b.configure(obse.BSEOptions(useTDA=False, homo=homo, rpamin=0, rpamax=N - 1, qpmin=0, qpmax=q - 1, vmin=0,
                                cmax=q - 1, nmax=10, use_Hqp_offdiag=False, davidson_tolerance="lapack",
                                davidson_maxiter=200), g.rpa_input_energies(), g.get_hqp())
    es = b.solve_singlets()
    inter = obse.free_transition_dipoles(dip, C, 0, vt, homo + 1, ct)
    tdip = obse.coupled_transition_dipoles(es, inter, ct, vt, False)
    f = obse.oscillator_strengths(tdip, es["eigenvalues"])
    print("BSE", es["eigenvalues"], f, time.time() - t0, "s", flush=True)
    np.savez_compressed(out, gw_iterations=g.iterations, mos=C, mo_energies=e, vxc=vxc, homo=homo, q=q, removed=tc.removed,
                        rhf_energy=hf["total_energy"], QPpert_energies=g.get_gwa_results(),
                        RPA_inputenergies=g.rpa_input_energies(), Hqp=g.get_hqp(),
                        BSE_singlet_eigenvalues=es["eigenvalues"], oscillator_strengths=f,
                        transition_dipoles=np.array(tdip), ao3c_checksum=np.array([ao3c.sum(), np.abs(ao3c).sum()]),
                        ao3c_sample=ao3c[::97, ::13, ::7].copy())
    print("wrote", out, os.path.getsize(out), "bytes", flush=True)


if __name__ == "__main__":
    main()
