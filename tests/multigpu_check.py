"""Multi-GPU parity check (run under torchrun on >= 2 GPUs): the m-sharded GW-BSE job must reproduce the
single-GPU result of the same inputs.  Usage:
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multigpu_check.py [workload]
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from votca_b200 import synthetic  # noqa: E402
from votca_b200.api import Job  # noqa: E402


def make_job(s, device, **opts):
    job = Job(device)
    job.set_scalar("homo", s["homo"])
    for name in ("mos", "mo_energies", "vxc", "aux_overlap", "aux_coulomb"):
        job.set_array(name, s[name])
    job.set_ao3c(s["ao3c"])
    job.set_options(tasks="gw,singlets,triplets", gw__mode="evGW", gw__sc_max_iter=3, bse__exctotal=5, **opts)
    return job


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    name = sys.argv[1] if len(sys.argv) > 1 else "small"
    N, naux, homo = synthetic.CONFIGS[name]
    s = synthetic.make_small(N, naux, homo)
    worst = 0.0
    for tda in (True, False):
        ref = None
        if rank == 0:
            j1 = make_job(s, local, bse__useTDA=tda)
            j1.run()
            ref = {k: j1.get(k) for k in ("QPpert_energies", "Hqp", "BSE_singlet_eigenvalues",
                                          "BSE_triplet_eigenvalues", "RPA_inputenergies")}
            j1.close()
        job = make_job(s, local, bse__useTDA=tda)
        uid = [job.kernel_ctx().nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        job.comm_init(rank, world, uid[0])
        if tda:
            # hand this rank only its share of the AO integrals (aux-sharded fill + all-to-all);
            # the other pass keeps the full array on every rank
            lo, hi = rank * naux // world, (rank + 1) * naux // world
            share = np.ascontiguousarray(s["ao3c"][lo:hi])
            job.set_ao3c_partial(N, naux, lo, hi - lo, share.ctypes.data, False)
        job.run()
        if rank == 0:
            for k, v in ref.items():
                d = float(np.abs(v - job.get(k)).max())
                worst = max(worst, d)
                print(f"useTDA={tda} {k}: max|sharded - single| = {d:.3e}")
        job.close()
        dist.barrier()
    if rank == 0:
        print("MULTIGPU_CHECK", "PASS" if worst < 1e-8 else "FAIL", f"worst={worst:.3e} world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
