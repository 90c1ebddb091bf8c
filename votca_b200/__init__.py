"""votca_b200: B200-native (sm_100a) GW-BSE dense contraction path behind VOTCA-XTP's dftgwbse.

`votca_b200.csrc`  hand-written CUDA kernels + the C ABI (include/gwbse_b200.h)
`votca_b200.host`  C++ host layer mirroring the reference classes (TCMatrix_gwbse, RPA, Sigma_*, GW,
                   BSE_OPERATOR, DavidsonSolver, BSE) on top of the C ABI
`votca_b200.api`   NumPy-facing ctypes plumbing used by tests and bench.py
There is no CPU fallback: importing works anywhere, creating a Context needs a B200.
"""
__version__ = "0.1.0"
