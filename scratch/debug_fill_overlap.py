"""C60 tier-R G0W0 (gw task only): Sigma_x depends on the filled Mmn only.  Run with and without GWBSE_NO_FILL_OVERLAP=1
and compare (scratch/gpu_call18.sh)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np
from votca_b200 import realsys, synthetic
from votca_b200.api import Job, Context
tag = sys.argv[1]
s = realsys.system("c60-tzvp")
N, naux, homo = s["nbasis"], s["naux"], 179
ctx = Context(0)
dft = ctx.basis_create(*s["dft"])
S = ctx.ao_overlap(dft)
ctx.basis_destroy(dft); ctx.close()
w, U = np.linalg.eigh(S)
rng = np.random.default_rng(5)
Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
C = (U / np.sqrt(w)) @ (U.T @ Q)
e = synthetic.spectrum(N, homo, rng)
q = 3 * homo + 2
job = Job(0)
job.set_scalar("homo", homo)
job.set_array("mos", C); job.set_array("mo_energies", e); job.set_array("vxc", np.diag(e[:q]) * 0.0)
job.set_basis("dft", *s["dft"]); job.set_basis("aux", *s["aux"])
job.set_options(tasks="gw", gw__mode="G0W0", gw__sigma_integrator="ppm")
out = {}
for rep in range(3):
    job.run()
    out[f"sx{rep}"] = job.get("Sigma_x").copy()
    out[f"qp{rep}"] = job.get("QPpert_energies").copy()
    print(tag, rep, "removed", job.scalar("removed_functions"), "Sigma_x[0,0]", out[f"sx{rep}"][0, 0], "trace", np.trace(out[f"sx{rep}"]), flush=True)
np.savez(f"gpurun_out/c19_{tag}.npz", **out)
