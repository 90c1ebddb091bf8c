#!/usr/bin/env python
"""bench.py - GW+BSE wall seconds per molecule on N B200s (BASELINE.json metric), FP64 TFLOP/s against the
measured DMMA peak, and the host-CPU reference beside it.

One "step" = the full GW-BSE stage of one molecule: Mmn fill (AO->MO three-centre transform + V^-1/2),
evGW with the plasmon-pole model (RPA epsilon, PPM rotation, batched QP root search, Hqp), then BSE
(static screening, 10 singlets by Davidson, TDA off = full BSE as the reference default).  The workload
is the DCV5T def2-tzvp size of SURVEY.md section 8 (N=1249, Naux=3177, homo=143 -> m=q=431, B=41328) on
synthetic tier-S inputs (votca_b200/synthetic.py).

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload NAME]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gwbse_wall_seconds_per_molecule"
UNIT = "s/molecule"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dcv5t-tzvp")
    ap.add_argument("--mode", default="evGW", choices=["evGW", "G0W0"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_flops(N, naux, homo, counts, k_bse=15):
    """mathematically necessary dense work of the contraction stages."""
    q = min(3 * homo + 1, N - 1) + 1
    m, n = q, N
    n_occ, n_unocc = homo + 1, N - homo - 1
    S = n_occ * n_unocc
    vt, ct = n_occ, q - n_occ
    B = vt * ct
    it = counts["gw_iterations"]
    f = {}
    f["fill"] = 2.0 * naux * N * m * (N + n)
    f["mul_right"] = 2.0 * m * n * naux * naux * (1 + it + 1)
    f["epsilon"] = S * naux * (naux + 1.0) * (2 * it + 1)
    f["sigma_x"] = q * (q + 1.0) * n_occ * naux
    f["sigma_c_offdiag"] = 2.0 * q * q * n * naux
    if "bse_algorithmic_flops" in counts:  # summed by the library per operator product (F_bse with its flags)
        f["bse_matvec"] = counts["bse_algorithmic_flops"]
    else:
        cols = counts.get("bse_operator_columns", 10 * k_bse)
        per_col_A = 4.0 * B * naux + 2.0 * naux * vt * ct * (vt + ct) + 2.0 * B * (vt + ct)
        f["bse_matvec"] = cols * per_col_A
    return f


def make_inputs_gpu(torch, N, naux, homo, device, seed, lo=0, hi=None):
    """Synthetic tier-S inputs generated on the device with torch (plumbing, not the product).  Every rank draws
    the same random stream and keeps the aux functions [lo, hi) of the AO tensor (its share of the fill)."""
    hi = naux if hi is None else hi
    from votca_b200 import synthetic
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    rng = np.random.default_rng(seed)
    e = synthetic.spectrum(N, homo, rng)
    Q, _ = torch.linalg.qr(torch.randn(N, N, dtype=torch.float64, device=device, generator=g))
    A = torch.randn(naux, naux, dtype=torch.float64, device=device, generator=g)
    V = (A @ A.T / naux + torch.eye(naux, dtype=torch.float64, device=device)).cpu().numpy()
    band = torch.tensor(synthetic.band_profile(N) * synthetic.ao3c_sigma(N, naux, homo), dtype=torch.float64,
                        device=device)
    ao = torch.empty((hi - lo, N, N), dtype=torch.float64, device=device)
    blk = max(1, (1 << 27) // (N * N))
    for a0 in range(0, naux, blk):
        a1 = min(naux, a0 + blk)
        G = torch.randn(a1 - a0, N, N, dtype=torch.float64, device=device, generator=g) * band
        b0, b1 = max(a0, lo), min(a1, hi)
        if b1 > b0:
            G = G[b0 - a0:b1 - a0]
            ao[b0 - lo:b1 - lo] = G + G.transpose(1, 2)
        del G
    return {"mos": Q.cpu().numpy(), "mo_energies": e, "aux_overlap": np.eye(naux), "aux_coulomb": V,
            "vxc": synthetic.make_vxc(e, homo, rng), "homo": homo, "ao3c_dev": ao}


def build_job(inp, N, naux, mode, device_index):
    from votca_b200.api import Job
    job = Job(device_index)
    job.set_scalar("homo", inp["homo"])
    for name in ("mos", "mo_energies", "vxc", "aux_overlap", "aux_coulomb"):
        job.set_array(name, inp[name])
    job.set_options(tasks="gw,singlets", gw__mode=mode, gw__sigma_integrator="ppm", bse__exctotal=10,
                    bse__useTDA=False)
    return job


def run_reference(args, N, naux, homo, rank):
    """--impl reference: the CPU formulation of the reference on the host cores (oracle port, bounded sample)."""
    from oracle import cpu_baseline
    if rank != 0:
        return
    counts = {"gw_iterations": 7 if args.mode == "evGW" else 1, "davidson_iterations": 8, "bse_analysis_matmuls": 4,
              "bse_operator_products": 4 * 9 + 8,
              "bse_operator_columns": 1047}  # trial columns through the operator, counted by the GPU run (119 TFLOP)
    q = min(3 * homo + 1, N - 1) + 1
    counts["sigma_evaluations"] = 756 * q * counts["gw_iterations"]  # evaluations per level and iteration of the
    # adaptive QP search on this workload (counted by the GPU run: 325873 per iteration for q = 431)
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_baseline.estimate(N, naux, homo, counts, sample_scale=1.0)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        est = cpu_baseline.estimate(N, naux, homo, counts, sample_scale=24.0)
        vals.append(est["total_seconds"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    cores = os.cpu_count()
    sample = ("per stage a few loop iterations of the reference CPU formulation (aux functions / m slices / "
              "occupied levels / sigma evaluations / BSE rows), scaled by the iteration counts of the workload; "
              f"{est['sampled_seconds']:.1f} s of CPU work per step")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, N, naux, homo), "mode": args.mode},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "stages": {k: round(s["sample_s"] * s["factor"], 3) for k, s in est["stages"].items()},
                             "factorised": factorised_summary(est)},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def factorised_summary(est):
    """CPU time of the same workload in the factorised formulation the GPU path uses (fill in the cheaper order, BSE
    operator without rebuilding H): reference-formulation value / this = algorithmic part of the speed-up."""
    return {"value": est["factorised_total_seconds"], "unit": UNIT,
            "stages": {k: round(s["sample_s"] * s["factor"], 3) for k, s in est["factorised_stages"].items()},
            "note": "replaces fill_3c, bse_hd_rows, bse_hx_blocks of the reference formulation; other stages unchanged"}


def workload_name(args, N, naux, homo):
    q = min(3 * homo + 1, N - 1) + 1
    return (f"{args.workload}: synthetic tier-S, N={N} basis, Naux={naux}, homo={homo}, m=q={q}, "
            f"{args.mode}(ppm) + full BSE 10 singlets (Davidson)")


def main():
    args = parse_args()
    from votca_b200 import synthetic
    N, naux, homo = synthetic.CONFIGS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, N, naux, homo, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the fill is sharded over aux functions: a rank holds (and uploads) only its share of the AO tensor
    aux_lo, aux_hi = rank * naux // world, (rank + 1) * naux // world
    inp = make_inputs_gpu(torch, N, naux, homo, device, 20261017, aux_lo, aux_hi)
    job = build_job(inp, N, naux, args.mode, local_rank)
    if world > 1:
        from votca_b200.api import Context
        uid = [None]
        if rank == 0:
            uid[0] = job.kernel_ctx().nccl_unique_id()
        dist.broadcast_object_list(uid, src=0)
        job.comm_init(rank, world, uid[0])
    kctx = job.kernel_ctx()
    peak = kctx.fp64_peak_probe()

    # ---------------- value: inputs resident in HBM ----------------
    job.set_ao3c_partial(N, naux, aux_lo, aux_hi - aux_lo, inp["ao3c_dev"].data_ptr(), True)
    for _ in range(args.warmup):
        job.run()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = job.launch_count()
    kctx.bse_stats(reset=True)
    kctx.gemm_profile(True)
    kctx.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        job.run()
    dev_ms = kctx.timer_stop_ms()
    barrier()
    wall = time.perf_counter() - t0
    gstats = kctx.gemm_stats()
    kctx.gemm_profile(False)
    launches = job.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    tms = torch.tensor([max(dev_ms, wall * 1e3)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_per_step = float(tms.item()) / args.steps
    counts = {"gw_iterations": int(job.scalar("gw_iterations")),
              "sigma_evaluations": job.scalar("sigma_evaluations"),
              "davidson_iterations": int(job.scalar("singlet_davidson_iterations")),
              "bse_analysis_matmuls": 4}
    bse_flops, bse_products, bse_columns = kctx.bse_stats()
    # operator products (A or B block applied to a block of trial vectors) per step: the reference rebuilds
    # every row of H for each of them whatever the block width
    counts["bse_operator_products"] = max(1, bse_products // args.steps)
    counts["bse_operator_columns"] = bse_columns // args.steps
    counts["bse_algorithmic_flops"] = bse_flops / args.steps
    stage_times = {k: job.scalar(k) for k in ("time_fill", "time_gw", "time_bse")}
    results = {"QP_homo": float(job.get("QPpert_energies")[homo]), "QP_lumo": float(job.get("QPpert_energies")[homo + 1]),
               "S1": float(job.get("BSE_singlet_eigenvalues")[0]), "singlet_converged": job.scalar("singlet_converged")}

    if os.environ.get("GWBSE_PROFILE"):
        # one extra, untimed step with the library's region profiler (per entry point device/host ms); every rank
        # runs it (the step contains collectives), rank 0 writes the report
        kctx.set_option("profile", 1)
        kctx.gemm_profile(True)
        job.run()
        rep = kctx.profile_report() + "\n" + kctx.gemm_shape_report()
        kctx.gemm_profile(False)
        kctx.set_option("profile", 0)
        if rank == 0:
            sys.stderr.write(rep + "\n" + "\n".join(l for l in job.log().splitlines()[-40:]) + "\n")
            with open(os.environ["GWBSE_PROFILE"], "w") as fh:
                fh.write(rep)
        barrier()

    # ---------------- e2e: AO integrals from pinned host memory, results back to the host ----------------
    e2e = None
    try:
        import psutil
        host_free = psutil.virtual_memory().available
    except Exception:
        host_free = 0
    ao_bytes = 8 * naux * N * N
    e2e_fits = host_free > 1.25 * ao_bytes  # the ranks together stage one copy of the AO tensor
    if not args.no_e2e and not e2e_fits and rank == 0:
        sys.stderr.write("e2e skipped: host memory cannot hold one pinned AO tensor per rank\n")
    if not args.no_e2e and e2e_fits:
        pinned = True
        try:
            host_ao = torch.empty((aux_hi - aux_lo, N, N), dtype=torch.float64, pin_memory=True)
        except Exception:
            pinned = False
            host_ao = torch.empty((aux_hi - aux_lo, N, N), dtype=torch.float64)
        host_ao.copy_(inp["ao3c_dev"])
        torch.cuda.synchronize()
        job.set_ao3c_partial(N, naux, aux_lo, aux_hi - aux_lo, host_ao.data_ptr(), False)
        job.run()  # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            job.run()
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
        te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        q = min(3 * homo + 1, N - 1) + 1
        h2d = 8 * (naux * N * N + world * (N * N + 2 * naux * naux))  # whole job: AO tensor once, small inputs per rank
        d2h = 8 * (2 * q * q + 2 * q + 2 * 10 * (homo + 1) * (q - homo - 1))
        e2e = {"value": float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "host_memory": "pinned" if pinned else "pageable",
               "stage_seconds": {k: job.scalar(k) for k in ("time_fill", "time_gw", "time_bse")}}
        if os.environ.get("GWBSE_PROFILE"):
            kctx.set_option("profile", 1)
            job.run()
            rep = kctx.profile_report()
            kctx.set_option("profile", 0)
            if rank == 0:
                with open(os.environ["GWBSE_PROFILE"] + ".e2e", "w") as fh:
                    fh.write(rep)
            barrier()
        del host_ao

    if rank != 0:
        return
    flops = algorithmic_flops(N, naux, homo, counts)
    total_flops = sum(flops.values())
    achieved = gstats["flops"] / (gstats["ms"] * 1e-3) / 1e12 if gstats["ms"] > 0 else 0.0
    traffic = None
    traffic_note = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            ncu = json.load(open(prof))
            traffic = ncu.get("gemm_dram_bytes_per_launch")
            traffic_note = ("dram bytes of one ncu --set full capture of the dominant launch shape (" +
                            str(ncu.get("report")) + f"); algorithmic bytes of that launch "
                            f"{ncu.get('algorithmic_bytes_per_launch')}")
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": ms_per_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args, N, naux, homo), "mode": args.mode,
                   "l2_policy": "inputs larger than L2 (Mmn 13.9 GB, AO tensor 39.6 GB)" if N > 1000 else
                   "small workload, L2-resident", "gw_iterations": counts["gw_iterations"],
                   "davidson_iterations": counts["davidson_iterations"], "results": results,
                   "stage_seconds": stage_times,
                   "ao_integrals": "supplied as a synthetic tensor (the reference's libint stage, row a2, is an input; "
                                   "the device producer gwbse_mmn_fill_from_basis is not on this path)"},
        "tflops": {"value": total_flops / (ms_per_step * 1e-3) / 1e12, "algorithmic_tflop_per_step": total_flops / 1e12,
                   "stages_tflop": {k: round(v / 1e12, 3) for k, v in flops.items()}},
        "roofline": {"bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA.8x8x4)", "achieved": achieved,
                     "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": traffic,
                     "traffic_note": traffic_note,
                     "peak_source": "live DMMA issue-rate probe in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                    "cuBLAS Dgemm 8192^3 measured 36.0 TF/s, profiles/r01_fp64_peak_probe.txt)",
                     "launches": gstats["launches"], "kernel_ms_per_step": gstats["ms"] / args.steps,
                     "kernel_share_of_step": gstats["ms"] / args.steps / ms_per_step},
        "gpu_launches": int(launches), "clocks": clocks, "e2e": e2e,
    }
    if not args.no_cpu and world == 1:  # reported at N=1 only
        from oracle import cpu_baseline
        est = cpu_baseline.estimate(N, naux, homo, counts, sample_scale=24.0)
        line["cpu_baseline"] = {
            "value": est["total_seconds"], "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
            "sample": ("reference CPU formulation (NumPy/OpenBLAS port, all host threads) timed per stage on a few "
                       f"loop iterations and scaled by the run's iteration counts; {est['sampled_seconds']:.1f} s "
                       "of CPU work"),
            "stages": {k: round(s["sample_s"] * s["factor"], 3) for k, s in est["stages"].items()},
            "factorised": factorised_summary(est)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
