#!/usr/bin/env python
"""Round-2 tool (NOT yet run on a device): times the device AO-integral producer on a tier-R system and checks a few
aux blocks against the CPU harness built from the same source.

  python scratch/ao3c_bench.py --system c60-tzvp --aux-block 64 --check 2
  ncu --set full --clock-control none -k regex:ao3c_kernel -c 6 -o gpurun_out/ao3c python scratch/ao3c_bench.py --system benzene-tzvp
"""
import argparse
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from votca_b200 import realsys  # noqa: E402
from votca_b200.api import Context  # noqa: E402


def cpu_harness():
    src = os.path.join(ROOT, "tests", "host_harness", "ao3c_host.cc")
    out = os.path.join(ROOT, "gpurun_out", "libao3c_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-pthread", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    p, i = ctypes.c_void_p, ctypes.c_int
    lib.ao3c_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, p]
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--system", default="benzene-tzvp", choices=sorted(realsys.SYSTEMS))
    ap.add_argument("--aux-block", type=int, default=64)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", type=int, default=0, help="aux blocks to compare with the CPU harness (full CPU run!)")
    a = ap.parse_args()
    s = realsys.system(a.system)
    N, naux = s["nbasis"], s["naux"]
    ctx = Context(0)
    t0 = time.time()
    dft, aux = ctx.basis_create(*s["dft"]), ctx.basis_create(*s["aux"])
    print(f"{a.system}: N={N} Naux={naux} shells={len(s['dft'][0])}/{len(s['aux'][0])} basis upload {time.time() - t0:.2f} s")
    buf = ctx.malloc(a.aux_block * N * N)
    ctx.set_option("profile", 1)
    best = 1e30
    for rep in range(a.reps + 1):
        ctx.sync()
        ctx.timer_start()
        for a0 in range(0, naux, a.aux_block):
            cnt = min(a.aux_block, naux - a0)
            ctx.call("gwbse_ao3c_block_dev", aux, dft, a0, cnt, buf)
        ms = ctx.timer_stop_ms()
        if rep:
            best = min(best, ms)
        print(f"  pass {rep}: {ms:.1f} ms  ({naux * N * N / max(ms, 1e-9) / 1e6:.1f} G integrals/s incl. mirror)")
    print(f"best {best:.1f} ms; launches so far {ctx.launch_count()}")
    print(ctx.profile_report())
    if a.check:
        lib = cpu_harness()
        ref = np.empty((naux, N, N))
        pt = [x.ctypes.data for x in s["dft"]] + [len(s["aux"][0])] + [x.ctypes.data for x in s["aux"]]
        lib.ao3c_host(len(s["dft"][0]), *pt[:5], pt[5], *pt[6:], 1, ref.ctypes.data)
        rng = np.random.default_rng(0)
        for a0 in rng.integers(0, max(1, naux - a.aux_block), a.check):
            got = ctx.ao3c_block(aux, dft, int(a0), a.aux_block)
            err = np.abs(got - ref[a0:a0 + a.aux_block]).max() / np.abs(ref).max()
            print(f"  aux block {a0}: max rel deviation from the CPU run of the same source {err:.2e}")
    ctx.free(buf)
    ctx.basis_destroy(aux)
    ctx.basis_destroy(dft)
    ctx.close()


if __name__ == "__main__":
    main()
