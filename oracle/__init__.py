"""CPU oracle for the GW-BSE dense FP64 contraction path of VOTCA-XTP.

TEST INFRASTRUCTURE ONLY.  This package is a plain NumPy/SciPy restatement of
the reference algorithm (votca/votca, xtp/src/libxtp/gwbse/*, threecenter.cc,
davidsonsolver.cc ...).  It exists to check the CUDA product path
(`votca_b200`) and to serve as the timed CPU baseline in `bench.py`.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product never does.

Parity pinning: every module below is checked against the reference's own
MatrixMarket fixtures (xtp/src/tests/DataFiles/{threecenter_gwbse,rpa,
sigma_exact,sigma_cda,sigma_ppm,gw,bse,bse_operator}) in
`tests/test_oracle_golden.py` at the tolerances of the reference's unit tests.
AO integrals: overlap, 2-centre and 3-centre Coulomb and dipole integrals are pinned on
the reference's aomatrix/, aomatrix3d/ and threecenter_dft/ fixtures up to l = 6
(contracted S/P/D/F overlap, G-shell dipoles, I-shell overlap / Coulomb and
G x G | I three-centre integrals from the shipped "large_l" data, minus the one
function per I shell that data is self-inconsistent in).  No fixture holds a d/f/g
integral in the GW layout itself; that layout differs from the
pinned one only by Pseudo_InvSqrt_GWBSE, which is pinned separately.
"""
