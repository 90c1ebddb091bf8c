#!/usr/bin/env python
"""dftgwbse - the `gwbse` task of the reference's `xtp_tools -e dftgwbse` calculator (xtp/src/libxtp/tools/dftgwbse.cc:
63-130 driving GWBSEEngine::ExcitationEnergies, xtp/src/libxtp/gwbseengine.cc:91-238) on the B200 library.

The reference tool runs a DFT package first (tasks input, dft, parse) and hands the parsed Orbitals object to GWBSE.
DFT is outside this path, so the DFT results are an input file here; everything from `GWBSE::Initialize` on is the
reference's flow: the same options file, `<job_name>.orb` and `<job_name>_summary.xml` as outputs.

    python -m votca_b200.tools.dftgwbse -o dftgwbse.xml --dft dft_results.npz [--device 0]

options file: what `xtp_tools -e dftgwbse -o` takes - `<options><dftgwbse><job_name>..</job_name><tasks>..</tasks>
<gwbse>..</gwbse></dftgwbse></options>`; the `<gwbse>` subtree goes to the library unchanged (unknown keys are errors).
dft_results.npz (the members of the parsed Orbitals object GWBSE reads):
    mos (N x N, column-major MO coefficients), mo_energies (N), homo, vxc (q x q in the MO basis, qp window of the
    options), optional ScaHFX, dft_total_energy; and either
      elements + positions_bohr + basis + auxbasis   names of votca_b200/data basis sets: all AO integrals (three- and
                                                    two-centre Coulomb, overlap, dipoles) are produced on the device
    or ao3c (Naux x N x N) + aux_overlap + aux_coulomb [+ dipole_x/y/z interlevel dipoles]   host-computed integrals
    unrestricted references add mos_beta, mo_energies_beta, homo_beta, vxc_beta (tasks gw / exciton_uks).
    bse.fragments needs the nuclear charges: taken from `elements`, or nuclear_charges + basis_atom_index + ao_overlap.
"""
import argparse
import re
import sys

import numpy as np


def read_tool_options(path):
    """job_name and tasks of the <dftgwbse> element (dftgwbse.xml defaults: system; input,dft,parse,gwbse)"""
    text = open(path).read()
    text = re.sub(r"<!--.*?-->", "", text, flags=re.S)
    outer = re.sub(r"<gwbse\b.*?</gwbse>", "", text, flags=re.S)  # the engine's own keys, not those of the subpackage

    def leaf(tag, default):
        m = re.search(rf"<{tag}\b[^>]*>(.*?)</{tag}>", outer, flags=re.S)
        return m.group(1).strip() if m and m.group(1).strip() else default
    return leaf("job_name", "system"), [t.strip() for t in leaf("tasks", "input,dft,parse,gwbse").split(",")]


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-o", "--options", required=True, help="options XML of the dftgwbse tool")
    ap.add_argument("--dft", required=True, help="npz with the DFT results (see above)")
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    from votca_b200 import realsys
    from votca_b200.api import Job

    job_name, tasks = read_tool_options(a.options)
    if "gwbse" not in tasks:
        print("nothing to do: the only task of dftgwbse on this path is 'gwbse' (DFT results are an input)")
        return 0
    d = np.load(a.dft, allow_pickle=False)
    job = Job(a.device)
    job.load_options_xml(a.options)
    unrestricted = "mos_beta" in d.files
    for name in ("mos", "mo_energies", "vxc", "mos_beta", "mo_energies_beta", "vxc_beta", "aux_overlap", "aux_coulomb",
                 "dipole_x", "dipole_y", "dipole_z"):
        if name in d.files:
            job.set_array(name, d[name])
    for name in ("homo", "homo_beta", "ScaHFX", "dft_total_energy"):
        if name in d.files:
            job.set_scalar(name, float(d[name]))
    if "ao3c" in d.files:
        job.set_ao3c(d["ao3c"])
    elif "basis" in d.files:
        el = [str(e) for e in d["elements"]]
        job.set_basis("dft", *realsys.shell_arrays(str(d["basis"]), el, d["positions_bohr"]))
        job.set_basis("aux", *realsys.shell_arrays(str(d["auxbasis"]), el, d["positions_bohr"]))
    else:
        raise SystemExit("the DFT results hold neither AO integrals (ao3c) nor basis-set names (basis, auxbasis)")
    # bse.fragments (Lowdin populations of the excitons on atom groups): nuclear charges from the elements, or the
    # arrays nuclear_charges / basis_atom_index / ao_overlap of the DFT results when the integrals come as arrays
    if "elements" in d.files and "nuclear_charges" not in d.files:
        z = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8, "F": 9, "Ne": 10, "Na": 11, "Mg": 12,
             "Al": 13, "Si": 14, "P": 15, "S": 16, "Cl": 17, "Ar": 18}
        job.set_array("nuclear_charges", np.array([float(z[str(e)]) for e in d["elements"]]))
    for name in ("nuclear_charges", "basis_atom_index", "ao_overlap"):
        if name in d.files:
            job.set_array(name, np.asarray(d[name], dtype=np.float64))
    archive, summary = job_name + ".orb", job_name + "_summary.xml"
    if unrestricted:
        job.set_summary_output(summary)
        job.run_uks()
    else:
        job.set_orb_output(archive)
        job.set_summary_output(summary)
        job.run()
    sys.stdout.write(job.log())
    if unrestricted:
        np.savez(job_name + "_uks_results.npz", **{k: job.get(k) for k in (
            "QPpert_energies_alpha", "QPpert_energies_beta", "RPA_inputenergies_alpha", "RPA_inputenergies_beta",
            "Hqp_alpha", "Hqp_beta", "BSE_uks_eigenvalues", "BSE_uks_eigenvectors", "BSE_uks_eigenvectors2",
            "BSE_uks_dynamic", "uks_transition_dipoles", "uks_oscillator_strengths")})
        print(f"Saving data to {job_name}_uks_results.npz")
        print(f"Writing output to {summary}")
    else:
        print(f"Saving data to {archive}")
        print(f"Writing output to {summary}")
    job.close()
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
