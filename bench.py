#!/usr/bin/env python
"""bench.py - GW+BSE wall seconds per molecule on N B200s (BASELINE.json metric), FP64 TFLOP/s against the
measured DMMA peak, and the host-CPU reference beside it.

One "step" = the full GW-BSE stage of one molecule: Mmn fill (AO->MO three-centre transform + V^-1/2),
GW with the plasmon-pole model (RPA epsilon, PPM rotation, batched QP root search, Hqp), then BSE
(static screening, 10 singlets by Davidson, TDA off = full BSE as the reference default).

Workloads (SURVEY.md section 8 sizes):
  c60-tzvp    (headline) BASELINE config 4, the largest configuration that fits one GPU and the one BASELINE.json
              shards over 8: C60 def2-tzvp + aux-def2-tzvp, N=1860, Naux=4560, homo=179 -> m=q=539, B=64620; G0W0.
              Tier R: real geometry and basis sets, the AO Coulomb integrals are produced on the device
              (gwbse_mmn_fill_from_basis path: the 126 GB AO tensor never exists), synthetic orthonormal MOs.
  dcv5t-tzvp  (reported under "also") BASELINE config 3: N=1249, Naux=3177, homo=143 -> m=q=431, B=41328; evGW.
              Synthetic tier-S inputs (votca_b200/synthetic.py): the AO three-centre tensor is an input (39.6 GB,
              resident in HBM for `value`, in pinned host memory for `e2e`).
  benzene-tzvp, small, medium: short runs of the same two kinds.

  python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload NAME] [--mode evGW|G0W0]
"""
import os
import sys

if "--impl" in sys.argv and sys.argv[sys.argv.index("--impl") + 1:][:1] == ["reference"]:
    # the CPU arm uses every host core whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1);
    # must happen before NumPy / OpenBLAS / libgomp are loaded
    _n = str(os.cpu_count() or 1)
    for _k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_k] = _n

import argparse  # noqa: E402
import json  # noqa: E402
import subprocess  # noqa: E402
import threading  # noqa: E402
import time  # noqa: E402

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "gwbse_wall_seconds_per_molecule"
UNIT = "s/molecule"
TIER_R = ("c60-tzvp", "benzene-tzvp", "methane-svp")
DEFAULT_MODE = {"c60-tzvp": "G0W0"}
COUNTS_FILE = os.path.join(ROOT, "votca_b200", "data", "bench_counts.json")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c60-tzvp")
    ap.add_argument("--mode", default=None, choices=["evGW", "G0W0"])
    ap.add_argument("--also", default="auto",
                    help="second workload measured with 1 warm-up + 2 steps and reported under 'also' "
                         "(auto: dcv5t-tzvp beside the default c60-tzvp headline, none otherwise; '' switches it off)")
    ap.add_argument("--e2e-steps", type=int, default=3, help="upper bound on the timed steps of the end-to-end arm")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    a = ap.parse_args()
    if a.mode is None:
        a.mode = DEFAULT_MODE.get(a.workload, "evGW")
    if a.also == "auto":
        a.also = "dcv5t-tzvp" if a.workload == "c60-tzvp" else ""
    return a


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


def sizes(workload):
    from votca_b200 import synthetic
    N, naux, homo = synthetic.CONFIGS[workload]
    q = min(3 * homo + 1, N - 1) + 1
    return N, naux, homo, q


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


def workload_name(workload, mode, N, naux, homo):
    q = min(3 * homo + 1, N - 1) + 1
    kind = ("tier-R: real geometry + basis sets, AO integrals on the device, synthetic orthonormal MOs"
            if workload in TIER_R else "synthetic tier-S")
    return (f"{workload}: {kind}, N={N} basis, Naux={naux}, homo={homo}, m=q={q}, "
            f"{mode}(ppm) + full BSE 10 singlets (Davidson)")


def l2_policy(N, naux, q):
    mmn = 8.0 * q * ((N + 15) // 16 * 16) * naux
    ao = 8.0 * naux * N * N
    if mmn > 4 * 126e6:
        return f"inputs larger than L2 (Mmn {mmn / 1e9:.1f} GB, AO tensor {ao / 1e9:.1f} GB, L2 0.126 GB)"
    return f"small workload (Mmn {mmn / 1e6:.0f} MB): partly L2-resident, not a headline size"


# ------------------------------------------------------------------------------------------- reference arm
def reference_counts(workload, mode, q):
    """Iteration counts of the workload as counted by a GPU run (votca_b200/data/bench_counts.json); generic
    estimates for a (workload, mode) that has none on record."""
    try:
        with open(COUNTS_FILE) as fh:
            rec = json.load(fh).get(f"{workload}/{mode}")
    except OSError:
        rec = None
    if rec:
        return dict(rec), rec.get("source", "bench_counts.json")
    it = 7 if mode == "evGW" else 1
    return ({"gw_iterations": it, "sigma_evaluations": 756 * q * it, "davidson_iterations": 8,
             "bse_analysis_matmuls": 4, "bse_operator_products": 44, "bse_operator_columns": 1047},
            "generic estimate (no GPU run of this workload / mode on record)")


def thread_report():
    info = {"os_cpu_count": os.cpu_count(), "OMP_NUM_THREADS": os.environ.get("OMP_NUM_THREADS")}
    try:
        from threadpoolctl import threadpool_info
        info["blas"] = [{"api": t.get("internal_api"), "threads": t.get("num_threads")} for t in threadpool_info()]
    except Exception:
        pass
    return info


def run_reference(args, rank):
    """--impl reference: the CPU formulation of the reference on the host cores (oracle port, bounded sample)."""
    if rank != 0:
        return
    from oracle import cpu_baseline
    N, naux, homo, q = sizes(args.workload)
    counts, source = reference_counts(args.workload, args.mode, q)
    ao_system = args.workload if args.workload in TIER_R else None
    # Bounded: the whole run is to end within a few minutes whatever --steps is.  The first (untimed) call times the
    # single-call samples (naux x naux eigh / inverse, AO-integral shell triples) once and keeps them; a second one
    # measures what a unit of loop sampling costs on this box, and the sample size of the timed steps follows from
    # the budget (GWBSE_REF_BUDGET_S seconds, default 240).
    once = {}
    cpu_baseline.estimate(N, naux, homo, counts, sample_scale=1.0, ao_system=ao_system, once=once)
    t0 = time.perf_counter()
    cpu_baseline.estimate(N, naux, homo, counts, sample_scale=1.0, ao_system=ao_system, once=once)
    unit_s = max(time.perf_counter() - t0, 1e-3)
    budget = float(os.environ.get("GWBSE_REF_BUDGET_S", "240"))
    scale = float(min(24.0, max(1.0, np.floor(budget / max(args.steps, 1) / unit_s))))
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        est = cpu_baseline.estimate(N, naux, homo, counts, sample_scale=scale, ao_system=ao_system, once=once)
        vals.append(est["total_seconds"])
    wall = time.perf_counter() - t0
    v = float(np.mean(vals))
    threads = thread_report()
    sample = ("per stage a few loop iterations of the reference CPU formulation (aux functions / m slices / "
              "occupied levels / sigma evaluations / BSE rows), scaled by the iteration counts of the workload; "
              f"{est['sampled_seconds']:.1f} s of CPU work per step (sample scale {scale:g} of 24, chosen for a "
              f"{budget:.0f} s run; eigh / inverse / AO-integral samples timed once per run)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "extrapolated": True,
            "value_is": "sampled stage times x iteration counts of the workload (ms_per_step is the sampler's own wall time)",
            "config": {"workload": workload_name(args.workload, args.mode, N, naux, homo), "mode": args.mode,
                       "l2_policy": l2_policy(N, naux, q)},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
                             "threads": threads, "counts": counts, "counts_source": source,
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


# ------------------------------------------------------------------------------------------- inputs
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


def make_inputs_tier_r(torch, workload, device, local_rank, seed):
    """Real geometry and basis sets; MOs = S^-1/2 Q (orthonormal in the real AO metric, Q from a seeded QR), gapped
    synthetic spectrum, Vxc from the run's own exchange self-energy (set by calibrate_vxc)."""
    from votca_b200 import realsys, synthetic
    from votca_b200.api import Context
    s = realsys.system(workload)
    N, naux, homo, q = sizes(workload)
    assert (N, naux) == (s["nbasis"], s["naux"]), (N, naux, s["nbasis"], s["naux"])
    ctx = Context(local_rank)
    dft = ctx.basis_create(*s["dft"])
    S = torch.tensor(ctx.ao_overlap(dft), dtype=torch.float64, device=device)
    ctx.basis_destroy(dft)
    ctx.close()
    w, U = torch.linalg.eigh(S)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    Q, _ = torch.linalg.qr(torch.randn(N, N, dtype=torch.float64, device=device, generator=g))
    C = (U * w.rsqrt()) @ (U.T @ Q)
    rng = np.random.default_rng(seed)
    e = synthetic.spectrum(N, homo, rng)
    return {"mos": C.cpu().numpy(), "mo_energies": e, "vxc": synthetic.make_vxc(e, homo, rng), "homo": homo,
            "basis": s, "min_overlap_eigenvalue": float(w.min().item())}


def build_job(inp, mode, device_index, exctotal=10):
    from votca_b200.api import Job
    job = Job(device_index)
    job.set_scalar("homo", inp["homo"])
    for name in ("mos", "mo_energies", "vxc", "aux_overlap", "aux_coulomb"):
        if name in inp:
            job.set_array(name, inp[name])
    job.set_options(tasks="gw,singlets", gw__mode=mode, gw__sigma_integrator="ppm", bse__exctotal=exctotal,
                    bse__useTDA=False)
    return job


def set_basis(job, inp):
    job.set_basis("dft", *inp["basis"]["dft"])
    job.set_basis("aux", *inp["basis"]["aux"])


def calibrate_vxc(job, inp, rng_seed=11):
    """Tier R: Vxc := 0.9 diag(Sigma_x) + small symmetric noise, Sigma_x from one exchange-only pre-run, so that the
    quasiparticle corrections have the size they have on a real system (as tier S does by construction)."""
    job.set_option("tasks", "gw")
    job.set_option("gw.mode", "G0W0")
    job.run()
    sx = np.diag(job.get("Sigma_x")).copy()
    q = sx.size
    R = 0.005 * np.random.default_rng(rng_seed).standard_normal((q, q))
    inp["vxc"] = np.diag(0.9 * sx) + 0.5 * (R + R.T)
    job.set_array("vxc", inp["vxc"])
    job.set_option("tasks", "gw,singlets")
    return sx


# ------------------------------------------------------------------------------------------- one workload
class Bench:
    def __init__(self, args, torch, dist, rank, local_rank, world):
        self.args, self.torch, self.dist = args, torch, dist
        self.rank, self.local_rank, self.world = rank, local_rank, world
        self.device = torch.device("cuda", local_rank)
        self.uid = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxf(self, v):
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def connect(self, job):
        if self.world > 1:
            uid = [None]
            if self.rank == 0:
                uid[0] = job.kernel_ctx().nccl_unique_id()
            self.dist.broadcast_object_list(uid, src=0)
            job.comm_init(self.rank, self.world, uid[0])

    def run(self, workload, mode, steps, warmup, e2e_steps, with_e2e, profile_path=None):
        torch, args = self.torch, self.args
        N, naux, homo, q = sizes(workload)
        rank, world = self.rank, self.world
        tier_r = workload in TIER_R
        aux_lo, aux_hi = rank * naux // world, (rank + 1) * naux // world
        if tier_r:
            inp = make_inputs_tier_r(torch, workload, self.device, self.local_rank, 20261017)
        else:
            inp = make_inputs_gpu(torch, N, naux, homo, self.device, 20261017, aux_lo, aux_hi)
        job = build_job(inp, mode, self.local_rank)
        self.connect(job)
        kctx = job.kernel_ctx()
        peak = kctx.fp64_peak_probe()
        if tier_r:
            set_basis(job, inp)
            calibrate_vxc(job, inp)
            job.set_option("gw.mode", mode)
        else:
            job.set_ao3c_partial(N, naux, aux_lo, aux_hi - aux_lo, inp["ao3c_dev"].data_ptr(), True)

        # ---------------- value: inputs resident in HBM ----------------
        for _ in range(warmup):
            job.run()
        self.barrier()
        sampler = ClockSampler(self.local_rank)
        if rank == 0:
            sampler.start()
        launches0 = job.launch_count()
        kctx.bse_stats(reset=True)
        kctx.gemm_profile(True)
        kctx.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            job.run()
        dev_ms = kctx.timer_stop_ms()
        self.barrier()
        wall = time.perf_counter() - t0
        gstats = kctx.gemm_stats()
        kctx.gemm_profile(False)
        launches = job.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        ms_per_step = self.maxf(max(dev_ms, wall * 1e3)) / steps
        counts = {"gw_iterations": int(job.scalar("gw_iterations")),
                  "sigma_evaluations": job.scalar("sigma_evaluations"),
                  "davidson_iterations": int(job.scalar("singlet_davidson_iterations")),
                  "bse_analysis_matmuls": 4}
        bse_flops, bse_products, bse_columns = kctx.bse_stats()
        counts["bse_operator_products"] = max(1, bse_products // steps)
        counts["bse_operator_columns"] = bse_columns // steps
        counts["bse_algorithmic_flops"] = bse_flops / steps
        dn_builds, dn_columns, dn_bytes = kctx.bse_dense_stats()
        bse_dense = {"blocks_built": dn_builds, "columns_from_resident_blocks": dn_columns,
                     "resident_GB_this_rank": round(dn_bytes / 1e9, 2)}
        stage_times = {k: job.scalar(k) for k in ("time_fill", "time_gw", "time_bse")}
        qp = job.get("QPpert_energies")
        results = {"QP_homo": float(qp[homo]), "QP_lumo": float(qp[homo + 1]),
                   "S1": float(job.get("BSE_singlet_eigenvalues")[0]),
                   "singlet_converged": job.scalar("singlet_converged")}

        if profile_path:
            # one extra, untimed step with the library's region profiler (per entry point device/host ms); every rank
            # runs it (the step contains collectives), rank 0 writes the report
            kctx.set_option("profile", 1)
            kctx.gemm_profile(True)
            job.run()
            rep = kctx.profile_report() + "\n" + kctx.gemm_shape_report()
            kctx.gemm_profile(False)
            kctx.set_option("profile", 0)
            if rank == 0:
                sys.stderr.write(rep + "\n" + "\n".join(ln for ln in job.log().splitlines()[-40:]) + "\n")
                with open(profile_path, "w") as fh:
                    fh.write(rep)
            self.barrier()

        # ---------------- e2e: every input from host memory, results back to the host ----------------
        e2e = None
        if with_e2e:
            e2e = self.run_e2e(job, inp, workload, N, naux, homo, q, aux_lo, aux_hi, min(steps, e2e_steps), profile_path)
        job.close()
        del inp

        flops = algorithmic_flops(N, naux, homo, counts)
        total_flops = sum(flops.values())
        achieved = gstats["flops"] / (gstats["ms"] * 1e-3) / 1e12 if gstats["ms"] > 0 else 0.0
        rec = {"workload": workload, "mode": mode, "N": N, "naux": naux, "homo": homo, "q": q,
               "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup, "counts": counts,
               "stage_seconds": stage_times, "results": results, "launches": int(launches), "clocks": clocks,
               "flops": flops, "total_flops": total_flops, "gemm": gstats, "gemm_tflops": achieved, "peak": peak,
               "e2e": e2e, "bse_dense": bse_dense}
        return rec

    def run_e2e(self, job, inp, workload, N, naux, homo, q, aux_lo, aux_hi, steps, profile_path):
        torch, world = self.torch, self.world
        tier_r = workload in TIER_R
        host_ao = None
        if not tier_r:
            try:
                import psutil
                host_free = psutil.virtual_memory().available
            except Exception:
                host_free = 0
            ao_bytes = 8 * naux * N * N
            if host_free <= 1.25 * ao_bytes:  # the ranks together stage one copy of the AO tensor
                if self.rank == 0:
                    sys.stderr.write("e2e skipped: host memory cannot hold the pinned AO tensor\n")
                return None
            pinned = True
            try:
                host_ao = torch.empty((aux_hi - aux_lo, N, N), dtype=torch.float64, pin_memory=True)
            except Exception:
                pinned = False
                host_ao = torch.empty((aux_hi - aux_lo, N, N), dtype=torch.float64)
            host_ao.copy_(inp["ao3c_dev"])
            torch.cuda.synchronize()
            job.set_ao3c_partial(N, naux, aux_lo, aux_hi - aux_lo, host_ao.data_ptr(), False)
            h2d = 8 * (naux * N * N + world * (N * N + 2 * naux * naux))  # AO tensor once, small inputs per rank
        else:
            pinned = None
            nb = sum(a.nbytes for a in inp["basis"]["dft"]) + sum(a.nbytes for a in inp["basis"]["aux"])
            h2d = world * (8 * (N * N + N + q * q) + nb)  # MOs, energies, Vxc, basis tables per rank

        def one_step():
            if tier_r:  # the basis tables go host -> device again (pair records are rebuilt from them)
                set_basis(job, inp)
            for name in ("mos", "mo_energies", "vxc"):
                job.set_array(name, inp[name])
            job.run()
            return job.get("QPpert_energies"), job.get("BSE_singlet_eigenvalues"), job.get("BSE_singlet_eigenvectors")

        one_step()  # warm the staging buffers
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one_step()
        self.barrier()
        e2e_s = self.maxf((time.perf_counter() - t0) / steps)
        d2h = 8 * (2 * q * q + 2 * q + 2 * 10 * (homo + 1) * (q - homo - 1))
        e2e = {"value": e2e_s, "unit": UNIT, "steps": steps, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "host_memory": "pinned" if pinned else ("pageable" if pinned is False else "n/a"),
               "stage_seconds": {k: job.scalar(k) for k in ("time_fill", "time_gw", "time_bse")}}
        if profile_path:
            kctx = job.kernel_ctx()
            kctx.set_option("profile", 1)
            one_step()
            rep = kctx.profile_report()
            kctx.set_option("profile", 0)
            if self.rank == 0:
                with open(profile_path + ".e2e", "w") as fh:
                    fh.write(rep)
            self.barrier()
        del host_ao
        return e2e

    def sharded_vs_single(self):
        """The sharded path against a single-GPU recomputation on rank 0 (small synthetic workload, seconds): the
        largest deviation of the QP and BSE energies goes into the JSON line."""
        from votca_b200 import synthetic
        N, naux, homo = synthetic.CONFIGS["small"]
        s = synthetic.make_small(N, naux, homo)
        lo, hi = self.rank * naux // self.world, (self.rank + 1) * naux // self.world
        share = np.ascontiguousarray(s["ao3c"][lo:hi])
        job = build_job(s, "evGW", self.local_rank, exctotal=5)
        self.connect(job)
        job.set_ao3c_partial(N, naux, lo, hi - lo, share.ctypes.data, False)
        job.run()
        got = (job.get("QPpert_energies").copy(), job.get("BSE_singlet_eigenvalues").copy())
        job.close()
        dev = None
        if self.rank == 0:
            single = build_job(s, "evGW", self.local_rank, exctotal=5)
            single.set_ao3c(s["ao3c"])
            single.run()
            dev = {"workload": f"small: synthetic tier-S N={N}, Naux={naux}, evGW(ppm) + full BSE 5 singlets",
                   "ranks": self.world,
                   "max_abs_dev_QP_Ha": float(np.abs(got[0] - single.get("QPpert_energies")).max()),
                   "max_abs_dev_BSE_Ha": float(np.abs(got[1] - single.get("BSE_singlet_eigenvalues")).max())}
            single.close()
        self.barrier()
        return dev


def summarize(rec, with_cpu):
    """Nested record of a secondary workload."""
    cpu = None
    if with_cpu:
        from oracle import cpu_baseline
        est = cpu_baseline.estimate(rec["N"], rec["naux"], rec["homo"], rec["counts"], sample_scale=6.0,
                                    ao_system=rec["workload"] if rec["workload"] in TIER_R else None)
        cpu = {"value": est["total_seconds"], "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "extrapolated": True,
               "stages": {k: round(s["sample_s"] * s["factor"], 3) for k, s in est["stages"].items()},
               "factorised": est["factorised_total_seconds"]}
    out = {"value": rec["ms_per_step"] / 1e3, "unit": UNIT, "steps": rec["steps"], "warmup": rec["warmup"],
           "config": {"workload": workload_name(rec["workload"], rec["mode"], rec["N"], rec["naux"], rec["homo"]),
                      "mode": rec["mode"], "gw_iterations": rec["counts"]["gw_iterations"],
                      "davidson_iterations": rec["counts"]["davidson_iterations"], "results": rec["results"],
                      "stage_seconds": rec["stage_seconds"]},
           "tflops": rec["total_flops"] / (rec["ms_per_step"] * 1e-3) / 1e12,
           "gemm_tflops": rec["gemm_tflops"], "gemm_frac_of_peak": rec["gemm_tflops"] / rec["peak"] if rec["peak"] else None,
           "e2e": rec["e2e"], "gpu_launches": rec["launches"], "cpu_baseline": cpu}
    return out


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    b = Bench(args, torch, dist, rank, local_rank, world)
    try:
        parity = None
        if world > 1 and not args.no_parity:
            parity = b.sharded_vs_single()
        rec = b.run(args.workload, args.mode, args.steps, args.warmup, args.e2e_steps, not args.no_e2e,
                    os.environ.get("GWBSE_PROFILE"))
        also = None
        if args.also:
            torch.cuda.empty_cache()
            amode = DEFAULT_MODE.get(args.also, "evGW")
            prof = os.environ.get("GWBSE_PROFILE")
            also = b.run(args.also, amode, 2, 1, 1, not args.no_e2e, prof + ".also" if prof else None)
        if rank == 0:
            print(json.dumps(make_line(args, rec, also, parity, world)))
    finally:
        if world > 1:
            dist.destroy_process_group()


def make_line(args, rec, also, parity, world):
    N, naux, homo, q = rec["N"], rec["naux"], rec["homo"], rec["q"]
    counts, flops, gstats = rec["counts"], rec["flops"], rec["gemm"]
    ms_per_step, peak, achieved = rec["ms_per_step"], rec["peak"], rec["gemm_tflops"]
    tier_r = rec["workload"] in TIER_R
    traffic = traffic_note = None
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
    ao_note = ("produced on the device from the basis sets (gwbse_ao3c_block_dev inside the fill; the AO tensor never "
               "exists)" if tier_r else
               "supplied as a synthetic tensor (the reference's libint stage, row a2, is an input of this workload)")
    line = {
        "metric": METRIC, "value": ms_per_step / 1e3, "unit": UNIT, "n_gpus": world, "steps": rec["steps"],
        "warmup": rec["warmup"], "ms_per_step": ms_per_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(rec["workload"], rec["mode"], N, naux, homo), "mode": rec["mode"],
                   "l2_policy": l2_policy(N, naux, q)},  # the reference arm prints the same three keys
        "run": {"gw_iterations": counts["gw_iterations"], "davidson_iterations": counts["davidson_iterations"],
                "results": rec["results"], "stage_seconds": rec["stage_seconds"], "ao_integrals": ao_note,
                "bse_direct_terms": dict(rec["bse_dense"], note=(
                    "Hd / Hd2 applied from their B x B blocks, materialised once per (Mmn, screening) when that pays "
                    "back against the factorised products and the blocks fit in HBM (option bse_dense)"))},
        "tflops": {"value": rec["total_flops"] / (ms_per_step * 1e-3) / 1e12,
                   "algorithmic_tflop_per_step": rec["total_flops"] / 1e12,
                   "stages_tflop": {k: round(v / 1e12, 3) for k, v in flops.items()}},
        "roofline": {"bound": "tensor", "kernel": "gemm_tma_kernel / gemm_dmma_kernel (FP64 DMMA.8x8x4)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "traffic_note": traffic_note,
                     "peak_source": "live DMMA issue-rate probe in this run (MEASURED_PEAKS.json has no FP64 entry; "
                                    "cuBLAS Dgemm 8192^3 measured 36.0 TF/s, profiles/r01_fp64_peak_probe.txt)",
                     "launches": gstats["launches"], "kernel_ms_per_step": gstats["ms"] / rec["steps"],
                     "kernel_share_of_step": gstats["ms"] / rec["steps"] / ms_per_step},
        "gpu_launches": rec["launches"], "clocks": rec["clocks"], "e2e": rec["e2e"],
    }
    if parity is not None:
        line["sharded_vs_single"] = parity
    if also is not None:
        line["also"] = summarize(also, not args.no_cpu and world == 1)
    # iteration counts for the CPU arm (copied into votca_b200/data/bench_counts.json by the maintainer)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"bench_counts_{rec['workload']}_{rec['mode']}_{world}gpu.json"), "w") as fh:
            json.dump({f"{rec['workload']}/{rec['mode']}": {k: v for k, v in counts.items() if k != "bse_algorithmic_flops"}}, fh)
    except OSError:
        pass
    if not args.no_cpu and world == 1:  # reported at N=1 only
        from oracle import cpu_baseline
        est = cpu_baseline.estimate(N, naux, homo, counts, sample_scale=12.0,
                                    ao_system=rec["workload"] if tier_r else None)
        line["cpu_baseline"] = {
            "value": est["total_seconds"], "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "extrapolated": True,
            "sample": ("reference CPU formulation (NumPy/OpenBLAS port, all host threads) timed per stage on a few "
                       f"loop iterations and scaled by the run's iteration counts; {est['sampled_seconds']:.1f} s "
                       "of CPU work"),
            "stages": {k: round(s["sample_s"] * s["factor"], 3) for k, s in est["stages"].items()},
            "factorised": factorised_summary(est)}
    return line


if __name__ == "__main__":
    main()
