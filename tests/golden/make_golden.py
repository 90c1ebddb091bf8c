#!/usr/bin/env python
"""Builds tests/golden/votca_fixtures.npz from the reference's own unit-test data.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Sources: xtp/src/tests/DataFiles/{threecenter_gwbse,rpa,sigma_exact,sigma_cda,
sigma_ppm,gw,bse,bse_operator,bsecoupling}/*.mm (MatrixMarket, 6 significant digits), the AO
integral references of aomatrix/, aomatrix3d/ and threecenter_dft/,
molecule.xyz and 3-21G.xml (identical in all of these directories), and the
inline vectors of test_sigma_*.cc, test_gw.cc, test_bse_operator.cc,
test_rpa_h2p.cc.  The npz is what the GPU box sees; /root/reference does not
exist there.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import basis as obasis  # noqa: E402
from oracle import mmio  # noqa: E402

REF = "/root/reference/xtp/src/tests/DataFiles"
DIRS = ["threecenter_gwbse", "rpa", "sigma_exact", "sigma_cda", "sigma_ppm", "gw", "bse", "bse_operator",
        "populationanalysis"]


def main():
    out = {}
    for d in DIRS:
        for fn in sorted(os.listdir(os.path.join(REF, d))):
            if fn.endswith(".mm"):
                out[f"{d}/{fn[:-3]}"] = mmio.read_matrix(os.path.join(REF, d, fn))
    elems, pos = obasis.read_xyz(os.path.join(REF, "rpa", "molecule.xyz"))
    out["molecule/elements"] = np.array(elems)
    out["molecule/positions_bohr"] = pos
    bs = obasis.load_basisset(os.path.join(REF, "rpa", "3-21G.xml"))
    out["basis/3-21G.json"] = np.array(json.dumps(bs))
    # inline vectors of the reference unit tests
    out["inline/sigma_exact_mo_energy"] = np.array(  # test_sigma_exact.cc:54-57, test_sigma_cda.cc:53-56
        [0.0468207, 0.0907801, 0.0907801, 0.104563, 0.592491, 0.663355, 0.663355, 0.768373, 1.69292,
         1.97724, 1.97724, 2.50877, 2.98732, 3.4418, 3.4418, 4.81084, 17.1838])
    out["inline/sigma_ppm_mo_energy"] = np.array(  # test_sigma_ppm.cc:57-60, test_bse_operator.cc:55-57
        [-0.612601, -0.341755, -0.341755, -0.341755, 0.137304, 0.16678, 0.16678, 0.16678, 0.671592,
         0.671592, 0.671592, 0.974255, 1.01205, 1.01205, 1.01205, 1.64823, 19.4429])
    out["inline/gw_mo_eigenvalues"] = np.array(  # test_gw.cc:46-49
        [-10.6784, -0.746424, -0.394948, -0.394948, -0.394948, 0.165212, 0.227713, 0.227713, 0.227713,
         0.763971, 0.763971, 0.763971, 1.05054, 1.13372, 1.13372, 1.13372, 1.72964])
    out["inline/bse_operator_epsilon_inv"] = np.array(  # test_bse_operator.cc:68-73
        [0.999807798016267, 0.994206065211371, 0.917916768047073, 0.902913813951883, 0.902913745974602,
         0.902913584797742, 0.853352878674581, 0.853352727016914, 0.853352541699637, 0.79703468058566,
         0.797034577207669, 0.797034400395582, 0.787701833916331, 0.518976361745313, 0.518975064844033,
         0.518973712898761, 0.459286057710524])
    out["inline/ppm_freq"] = np.array(  # test_ppm.cc:66-69
        [19.4503, 12.2429, 10.2167, 10.2167, 10.2167, 17.0832, 12.7117, 12.7117, 12.7117, 10.9455, 10.9455, 10.9455,
         11.1861, 9.60843, 9.60843, 9.60843, 9.61518])
    out["inline/ppm_weight"] = np.array(  # test_ppm.cc:70-74
        [3.59422e-05, 0.00121795, 0.00343632, 0.00343632, 0.00343632, 0.00394659, 0.012066, 0.012066, 0.012066,
         0.0241118, 0.0241118, 0.0241118, 0.0286786, 0.11036, 0.11036, 0.11036, 0.191732])
    out["inline/rpa_update_dft"] = np.array([-0.5, -0.4, -0.3, -0.2, -0.2, -0.1, 0, 0.1, 0.2, 0.3])  # test_rpa.cc:47-57
    out["inline/rpa_update_gw"] = np.array([-0.15, -0.05, 0.05, 0.15, 0.45, 0.55, 0.65])
    out["inline/rpa_update_ref"] = np.array([-0.85, -0.15, -0.05, 0.05, 0.15, 0.45, 0.55, 0.65, 0.75, 0.85])
    out["inline/rpa_h2p_erpa"] = np.array(-0.0587973)  # test_rpa_h2p.cc:86
    out["inline/rpa_h2p_omega"] = np.array(  # test_rpa_h2p.cc:62-71
        [0.104192, 0.104192, 0.187814, 0.559693, 0.559693, 0.572575, 0.577988, 0.577989, 0.579088, 0.618403,
         0.618403, 0.67005, 0.678538, 0.678538, 0.722771, 1.10797, 1.41413, 1.41413, 1.58866, 1.60381,
         1.60381, 1.64709, 1.87331, 1.87331, 1.88646, 1.8926, 1.89268, 1.89268, 1.933, 1.933, 2.01832,
         2.40974, 2.42192, 2.42192, 2.46371, 2.85829, 2.8853, 2.8853, 2.90367, 2.90367, 2.92541, 2.94702,
         3.3382, 3.3382, 3.35102, 3.3566, 3.3566, 3.35835, 3.39617, 3.39617, 4.22882, 4.71607, 4.72233,
         4.72233, 4.76567, 16.5917, 17.0793, 17.093, 17.093, 17.1377])
    # ---- AO integral fixtures (test_aomatrix.cc, test_aomatrix3d.cc, test_threecenter_dft.cc): the host-side
    # producer of the Fill inputs; these pin the oracle's integrals for d...i shells
    for d, files in (("aomatrix", ["overlap_ref", "coulomb_ref", "coulombinvsqrtgw_ref", "overlap_ref_contracted",
                                   "overlap_ref_gi", "coulomb_ref_gi", "kinetic_ref", "kinetic_ref_gi"]),
                     ("aopotential", ["esp_ref"]),
                     ("aomatrix3d", ["dip_ref_large_0", "dip_ref_large_1", "dip_ref_large_2"]),
                     ("threecenter_dft", ["Ref0", "Ref4", "RefList0", "RefList1", "RefList2", "RefList3"])):
        for fn in files:
            out[f"{d}/{fn}"] = mmio.read_matrix(os.path.join(REF, d, fn + ".mm"))
    for name, path in (("contracted", "aomatrix/contracted.xml"), ("G", "aomatrix/G.xml"), ("I", "aomatrix/I.xml")):
        out[f"basis/{name}.json"] = np.array(json.dumps(obasis.load_basisset(os.path.join(REF, path))))
    for name, path in (("C", "aomatrix/C.xyz"), ("C2", "aomatrix/C2.xyz")):
        elems, pos = obasis.read_xyz(os.path.join(REF, path))
        out[f"molecule_{name}/elements"] = np.array(elems)
        out[f"molecule_{name}/positions_bohr"] = pos
    # ---- integration-test checkpoints (xtp/src/tests/CMakeLists.txt:336-416): inputs and GW-BSE outputs of the
    # reference's own `xtp_tools -e dftgwbse` runs on water, 3-21G + aux-def2-svp (neutral: G0W0 exact sigma,
    # full BSE, 5 dynamical-screening iterations; neutral_tda: TDA, 5 states), read with oracle/orbfile.py
    from oracle.orbfile import OrbFile
    IT = os.path.join(REF, "xtp_tools_integration_tests")
    for tag in ("neutral", "neutral_tda"):
        f = OrbFile(os.path.join(IT, f"molecule_{tag}.orb"))
        at = f.attrs("/QMdata")
        for k in ("occupied_levels", "rpamin", "rpamax", "qpmin", "qpmax", "bse_vmin", "bse_cmax", "useTDA",
                  "use_Hqp_offdiag", "ScaHFX"):
            out[f"orb/{tag}/attr_{k}"] = np.array(at[k])
        for name, path in (("mo_energies", "mos/eigenvalues"), ("mos", "mos/eigenvectors"),
                           ("RPA_inputenergies", "RPA_inputenergies"), ("QPpert_energies", "QPpert_energies"),
                           ("QPdiag_eigenvalues", "QPdiag/eigenvalues"), ("QPdiag_eigenvectors", "QPdiag/eigenvectors"),
                           ("BSE_singlet_eigenvalues", "BSE_singlet/eigenvalues"),
                           ("BSE_singlet_eigenvectors", "BSE_singlet/eigenvectors"),
                           ("BSE_singlet_eigenvectors2", "BSE_singlet/eigenvectors2"),
                           ("BSE_singlet_dynamic", "BSE_singlet_dynamic")):
            out[f"orb/{tag}/{name}"] = f.read("/QMdata/" + path)
        n = out[f"orb/{tag}/BSE_singlet_eigenvalues"].shape[0]
        out[f"orb/{tag}/transition_dipoles"] = np.array(
            [f.read(f"/QMdata/transition_dipoles/ind{i}").ravel() for i in range(n)])
    elems, pos = obasis.read_xyz(os.path.join(IT, "molecule.xyz"))
    out["molecule_water/elements"] = np.array(elems)
    out["molecule_water/positions_bohr"] = pos
    out["basis/water_3-21G.json"] = np.array(json.dumps(obasis.load_basisset(os.path.join(IT, "3-21G.xml"))))
    auxbs = obasis.load_basisset("/root/reference/xtp/share/xtp/basis_sets/aux-def2-svp.xml")
    out["basis/aux-def2-svp_OH.json"] = np.array(json.dumps({el: auxbs[el] for el in ("O", "H")}))
    # ---- BASELINE config 0 (methane, def2-svp + aux-def2-svp, geometry of xtp-tutorials/tools/dftgwbse_CH4):
    # geometry and basis sets only; orbitals come from oracle/scf.py at test time ("tier R", own integrals)
    elems, pos = obasis.read_xyz("/root/reference/xtp-tutorials/tools/dftgwbse_CH4/methane.xyz")
    out["molecule_methane_tutorial/elements"] = np.array(elems)
    out["molecule_methane_tutorial/positions_bohr"] = pos
    for name in ("def2-svp", "aux-def2-svp"):
        bs = obasis.load_basisset(f"/root/reference/xtp/share/xtp/basis_sets/{name}.xml")
        out[f"basis/{name}_CH.json"] = np.array(json.dumps({el: bs[el] for el in ("C", "H")}))
    # ---- BSECoupling (test_bsecoupling.cc): methane monomer / dimer (B = A shifted by 4 bohr along x), 3-21G for
    # both the orbital and the auxiliary basis, monomer and dimer MOs, the dimer's QP eigenvectors, three monomer
    # singlets; inline eigenvalues of :84-88 (dimer MOs, unused by the coupling) and :113-119 (QPdiag)
    for fn in ("A_MOs", "AB_MOs", "Hqp", "spsi_ref"):
        out[f"bsecoupling/{fn}"] = mmio.read_matrix(os.path.join(REF, "bsecoupling", fn + ".mm"))
    elems, pos = obasis.read_xyz(os.path.join(REF, "bsecoupling", "molecule.xyz"))
    out["bsecoupling/elements"] = np.array(elems)
    out["bsecoupling/positions_bohr"] = pos
    out["bsecoupling/qpdiag_eigenvalues"] = np.array(
        [-10.504, -10.5038, -0.923616, -0.775673, -0.549084, -0.530193, -0.530193, -0.430293, -0.430293, -0.322766,
         0.267681, 0.307809, 0.326961, 0.326961, 0.36078, 0.381947, 0.414845, 0.414845, 0.906609, 0.906609, 0.993798,
         1.09114, 1.14639, 1.14639, 1.1966, 1.25629, 1.25629, 1.27991, 1.29122, 1.35945, 1.36705, 1.36705, 1.93286,
         2.11739])
    out["bsecoupling/known_answers_eV"] = np.array([23.662750, 9.529579])  # |j_diag|, |j_pert|, :138-139
    np.savez_compressed(os.path.join(HERE, "votca_fixtures.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
