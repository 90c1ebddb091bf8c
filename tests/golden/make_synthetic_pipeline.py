#!/usr/bin/env python
"""Generates tests/golden/synthetic_<name>_evgw_ppm_bse.npz: the full GW-BSE pipeline (evGW with the plasmon-pole
model, full BSE 10 singlets) of the CPU oracle on a synthetic tier-S system of votca_b200/synthetic.py.  The GPU test
(tests/test_gpu_zzz_large_pipeline.py) regenerates the same inputs from the seed, runs the CUDA path with the
treecode Sigma_c evaluator, split-K plans and a chunked BSE intermediate forced on, and compares with these numbers.

  python tests/golden/make_synthetic_pipeline.py small      (minutes)
  python tests/golden/make_synthetic_pipeline.py medium     (much longer: 2 ms per Sigma_c evaluation in NumPy)
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bse as obse  # noqa: E402
from oracle import gw as ogw  # noqa: E402
from oracle import threecenter  # noqa: E402
from votca_b200 import synthetic  # noqa: E402


def main(name):
    t0 = time.time()
    N, naux, homo = synthetic.CONFIGS[name]
    s = synthetic.make_small(N, naux, homo)
    q = s["vxc"].shape[0]
    tc = threecenter.TCMatrix(naux, 0, q - 1, 0, N - 1)
    tc.fill_from_integrals(s["ao3c"], s["aux_overlap"], s["aux_coulomb"], s["mos"])
    g = ogw.GW(tc, s["vxc"], s["mo_energies"])
    g.configure(ogw.GWOptions(homo=homo, qpmin=0, qpmax=q - 1, rpamin=0, rpamax=N - 1, gw_sc_max_iterations=50,
                              sigma_integration="ppm", g_sc_max_iterations=100))
    g.calculate_gw_perturbation()
    g.calculate_hqp()
    print("evGW iterations", g.iterations, "sigma evaluations", g.sigma_evals, time.time() - t0, "s", flush=True)
    b = obse.BSE(tc, factorised=True)
    b.configure(obse.BSEOptions(useTDA=False, homo=homo, rpamin=0, rpamax=N - 1, qpmin=0, qpmax=q - 1, vmin=0,
                                cmax=q - 1, nmax=10, use_Hqp_offdiag=False), g.rpa_input_energies(), g.get_hqp())
    es = b.solve_singlets()
    print("BSE", es["eigenvalues"], time.time() - t0, "s", flush=True)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"synthetic_{name}_evgw_ppm_bse.npz")
    np.savez_compressed(out, config=np.array([N, naux, homo]), gw_iterations=g.iterations,
                        QPpert_energies=g.get_gwa_results(), RPA_inputenergies=g.rpa_input_energies(),
                        Sigma_x_diag=np.diag(g.Sigma_x).copy(), Sigma_c_diag=np.diag(g.Sigma_c).copy(),
                        Hqp=g.get_hqp(), BSE_singlet_eigenvalues=es["eigenvalues"],
                        input_checksum=np.array([s["ao3c"].sum(), s["mos"].sum(), s["aux_coulomb"].sum()]))
    print("wrote", out, os.path.getsize(out), "bytes", flush=True)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "small")
