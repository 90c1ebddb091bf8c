"""Oracle: Loewdin populations of excited states on fragments (TEST INFRASTRUCTURE - never imported by votca_b200/).

Restates
  Populationanalysis<true> (Lowdin)   xtp/src/libxtp/populationanalysis.cc:27-45 (charge per atom), :47-84 (fragments),
                                      :113-138 (S^1/2 D S^1/2, diagonal summed per atom)
  Orbitals::DensityMatrixGroundState  xtp/src/libxtp/orbitals.cc:193-223 (closed shell: 2 C_occ C_occ^T)
  Orbitals::DensityMatrixExcitedState xtp/src/libxtp/orbitals.cc:516-600 (_R from X, minus _AR from Y without the TDA)
  BSE::printFragInfo                  xtp/src/libxtp/gwbse/bse.cc:362-376
Pinned on the reference's known answers (xtp/src/tests/test_populationanalysis.cc:38-74, 95-180) in
tests/test_oracle_golden.py.
"""
import numpy as np


def sqrt_overlap(S):
    w, U = np.linalg.eigh(S)
    return (U * np.sqrt(w)) @ U.T


def lowdin_per_atom(dmat, S, basis_atom, natoms):
    """populationanalysis.cc:113-138: electrons per atom of a density matrix"""
    sq = sqrt_overlap(S)
    diag = np.diag(sq @ dmat @ sq)
    out = np.zeros(natoms)
    np.add.at(out, np.asarray(basis_atom, dtype=int), diag)
    return out


def ground_state_density(mos, nocc):
    return 2.0 * mos[:, :nocc] @ mos[:, :nocc].T


def excited_state_densities(mos, vmin, homo, cmax, X, Y=None):
    """(hole, electron) AO density matrices of one exciton with coefficients X (and Y without the TDA), index v * ct + c"""
    vt, ct = homo - vmin + 1, cmax - homo
    Cv, Cc = mos[:, vmin:homo + 1], mos[:, homo + 1:cmax + 1]

    def parts(coeffs):
        A = np.asarray(coeffs).reshape(vt, ct)          # A[v, c]
        return Cv @ (A @ A.T) @ Cv.T, Cc @ (A.T @ A) @ Cc.T   # CalcAuxMat_vv, CalcAuxMat_cc
    hole, elec = parts(X)
    if Y is not None:
        vv, cc = parts(Y)
        hole, elec = hole - cc, elec - vv               # orbitals.cc:523-530, :602-650
    return hole, elec


def fragment_populations(S, mos, homo, vmin, cmax, basis_atom, nuclear_charges, fragments, X, Y=None):
    """populationanalysis.cc:47-84.  fragments: lists of atom indices.  Returns Gs (nfrag), H, E (nfrag x nstates)."""
    nuc = np.asarray(nuclear_charges, dtype=float)
    nat = len(nuc)
    gs = nuc - lowdin_per_atom(ground_state_density(mos, homo + 1), S, basis_atom, nat)
    nstates = X.shape[1]
    H = np.zeros((len(fragments), nstates))
    E = np.zeros((len(fragments), nstates))
    for s in range(nstates):
        dh, de = excited_state_densities(mos, vmin, homo, cmax, X[:, s], None if Y is None else Y[:, s])
        ah = lowdin_per_atom(dh, S, basis_atom, nat)
        ae = -lowdin_per_atom(de, S, basis_atom, nat)
        for f, atoms in enumerate(fragments):
            H[f, s], E[f, s] = ah[list(atoms)].sum(), ae[list(atoms)].sum()
    Gs = np.array([gs[list(atoms)].sum() for atoms in fragments])
    return Gs, H, E
