#!/usr/bin/env python
"""gpu_benchmark - the reference's `xtp_tools -e gpu_benchmark` calculator (xtp/src/libxtp/tools/gpu_benchmark.cc:
101-203) on the B200 library: the same seven parts under the same names, each timed `repetitions` times (default 5,
share/xtp/xml/gpu_benchmark.xml) with the reference's statistics (mean and population standard deviation of the
wall time per repetition, gpu_benchmark.cc:45-50), written as the same <GPU_Benchmark> XML.

    Filling_ThreeCenter                    TCMatrix_gwbse::Fill
    Multiplication_of_tensor_with_matrix   TCMatrix_gwbse::MultiplyRightWithAuxMatrix (random aux x aux matrix)
    RPA_evaluation                         RPA::calculate_epsilon_i(0.5) + calculate_epsilon_r(-0.5)
    SingletOperator_TDA   <1,2,1,0>        BSE_OPERATOR::matmul on 100 random vectors (bse_operator.h:79-87)
    TripletOperator_TDA   <1,0,1,0>
    SingletOperator_BTDA_B <0,2,0,1>
    HxOperator            <0,1,0,0>

Inputs: a synthetic tier-S workload (--workload, AO integrals handed over from host memory exactly as the shim does),
or a tier-R system (--system, votca_b200/realsys.py) whose integrals are produced on the device, which is what the
reference's Fill times too (libint + transform).  Host arrays go in and come out of every timed call, as in the
reference (Eigen matrices); this is the per-call boundary, not the resident-tensor pipeline bench.py measures.

    python -m votca_b200.tools.gpu_benchmark --workload medium --repetitions 5 --outputfile gpu_benchmark.xml
"""
import argparse
import math
import time

import numpy as np


def calc_statistics(v):
    """mean and population standard deviation, gpu_benchmark.cc:45-50"""
    mean = sum(v) / len(v)
    return mean, math.sqrt(max(sum(x * x for x in v) / len(v) - mean * mean, 0.0))


def run_part(payload, name, repetitions, sync):
    timings = []
    print(name)
    for _ in range(repetitions):
        sync()
        t0 = time.perf_counter()
        payload()
        sync()
        timings.append(time.perf_counter() - t0)
    mean, std = calc_statistics(timings)
    print(f"avg:{mean} std:{std}")
    return name, mean, std, timings


def to_xml(header, parts):
    out = ["<GPU_Benchmark>"]
    for k, v in header.items():
        out.append(f"\t<{k}>{v}</{k}>")
    for name, mean, std, timings in parts:
        out.append(f"\t<{name}>")
        out.append(f"\t\t<avg>{mean:.6f}</avg>")
        out.append(f"\t\t<std>{std:.6f}</std>")
        out.append("\t\t<runs>")
        out += [f"\t\t\t<timing>{t:.6f}</timing>" for t in timings]
        out.append("\t\t</runs>")
        out.append(f"\t</{name}>")
    out.append("</GPU_Benchmark>")
    return "\n".join(out) + "\n"


OPERATORS = (("SingletOperator_TDA", (1, 2, 1, 0)), ("TripletOperator_TDA", (1, 0, 1, 0)),
             ("SingletOperator_BTDA_B", (0, 2, 0, 1)), ("HxOperator", (0, 1, 0, 0)))


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--workload", default="medium", help="synthetic tier-S size (votca_b200.synthetic.CONFIGS)")
    ap.add_argument("--system", default=None, help="tier-R system (votca_b200.realsys.SYSTEMS): integrals on the device")
    ap.add_argument("--repetitions", type=int, default=5)
    ap.add_argument("--outputfile", default="gpu_benchmark.xml")
    ap.add_argument("--spacesize", type=int, default=100)
    a = ap.parse_args(argv)
    import os

    from votca_b200 import realsys, synthetic
    from votca_b200.api import Context
    ctx = Context(0)
    rng = np.random.default_rng(1)
    if a.system:
        s = realsys.system(a.system)
        N, naux = s["nbasis"], s["naux"]
        homo = sum({"H": 1, "C": 6, "N": 7, "S": 16}[e] for e in s["elements"]) // 2 - 1
        mos = np.linalg.qr(rng.standard_normal((N, N)))[0]
        energies = synthetic.spectrum(N, homo, rng)
        dft, aux = ctx.basis_create(*s["dft"]), ctx.basis_create(*s["aux"])
        S_aux, V_aux = ctx.ao_overlap(aux), ctx.ao_coulomb2c(aux)
        basis_name, aux_name = realsys.SYSTEMS[a.system][1], realsys.SYSTEMS[a.system][2]
    else:
        N, naux, homo = synthetic.CONFIGS[a.workload]
        s = synthetic.make_small(N, naux, homo)
        mos, energies, S_aux, V_aux = s["mos"], s["mo_energies"], s["aux_overlap"], s["aux_coulomb"]
        basis_name = aux_name = "synthetic tier-S " + a.workload
    rpamin, rpamax = 0, N - 1
    qpmax = min(3 * homo + 1, N - 1)  # ranges = default (gwbse.cc:115-120): max_3c = max(bse_cmax, qpmax)
    print(f"Repetitions:{a.repetitions}\nNumber of CPUs:{os.cpu_count()} \nNumber gpus:1")
    print(f"BasisSet:{basis_name} size:{N}\nAuxBasisSet:{aux_name} size:{naux}\nrpamin:{rpamin} rpamax:{rpamax}")
    ctx.mmn_alloc(naux, rpamin, qpmax, rpamin, rpamax)

    def fill():
        ctx.mmn_set_mos(mos)
        if a.system:
            ctx.mmn_fill_from_basis(aux, dft, 64)
        else:
            for a0 in range(0, naux, 64):
                ctx.mmn_fill_block(a0, s["ao3c"][a0:a0 + 64])
        L, _ = ctx.pseudo_invsqrt(S_aux, V_aux)  # threecenter.cc:72-90
        ctx.mmn_mul_right(L)

    parts = [run_part(fill, "Filling_ThreeCenter", a.repetitions, ctx.sync)]
    op = rng.uniform(-1.0, 1.0, (naux, naux))
    parts.append(run_part(lambda: ctx.mmn_mul_right(op), "Multiplication_of_tensor_with_matrix", a.repetitions, ctx.sync))
    fill()  # Mmn.Rebuild()
    result = np.zeros((naux, naux))

    def rpa():
        result[...] += ctx.rpa_epsilon(0, 0.5, 1e-4, energies, homo, rpamin, rpamax)
        result[...] += ctx.rpa_epsilon(1, -0.5, 1e-4, energies, homo, rpamin, rpamax)

    parts.append(run_part(rpa, "RPA_evaluation", a.repetitions, ctx.sync))
    hq = qpmax + 1
    ctx.bse_configure(homo, rpamin, 0, qpmax, np.diag(result).copy(), rng.uniform(-1.0, 1.0, (hq, hq)))
    state = rng.uniform(-1.0, 1.0, (ctx.bse_size, a.spacesize))
    for name, coeffs in OPERATORS:
        parts.append(run_part(lambda c=coeffs: ctx.bse_matmul(c, state), name, a.repetitions, ctx.sync))
    header = {"Repetitions": a.repetitions, "CPUs": os.cpu_count(), "GPUs": 1, "MKL_overload": 0,
              "Basisset": basis_name, "Basissetsize": N, "AuxBasisset": aux_name, "AuxBasissetsize": naux,
              "rpamin": rpamin, "rpamax": rpamax}
    with open(a.outputfile, "w") as fh:
        fh.write(to_xml(header, parts))
    ctx.close()
    return parts


if __name__ == "__main__":
    main()
