"""Is V^-1/2 on a second context perturbed by work running on the first one?  (C60 / aux-def2-tzvp, Naux = 4560)"""
import sys, threading, time
sys.path.insert(0, '.')
import numpy as np
from votca_b200 import realsys
from votca_b200.api import Context
s = realsys.system("c60-tzvp")
A, B = Context(0), Context(0)
dft, aux = A.basis_create(*s["dft"]), A.basis_create(*s["aux"])
S, V = A.ao_overlap(aux), A.ao_coulomb2c(aux)
w = np.linalg.eigvalsh(S)
print("S eig min", w[:4], "count < 5e-7:", int((w < 5e-7).sum()), flush=True)
L1, r1 = A.pseudo_invsqrt(S, V)
L2, r2 = B.pseudo_invsqrt(S, V)
print("alone: removed", r1, r2, "max|L1-L2|", np.abs(L1 - L2).max(), "max|L|", np.abs(L1).max(), flush=True)
n = s["nbasis"]
buf = A.malloc(64 * n * n)
stop = False
def busy():
    k = 0
    while not stop:
        A.call("gwbse_ao3c_block_dev", aux, dft, (64 * k) % 4480, 64, buf)
        k += 1
    A.sync()
for trial in range(3):
    stop = False
    t = threading.Thread(target=busy); t.start()
    time.sleep(0.2)
    t0 = time.perf_counter()
    L3, r3 = B.pseudo_invsqrt(S, V)
    dt = time.perf_counter() - t0
    stop = True; t.join()
    print(f"concurrent trial {trial}: removed {r3} max|L3-L1| {np.abs(L3 - L1).max():.3e} time {dt:.3f} s", flush=True)
t0 = time.perf_counter(); L4, r4 = B.pseudo_invsqrt(S, V); print("alone again", r4, np.abs(L4 - L1).max(), time.perf_counter() - t0)
