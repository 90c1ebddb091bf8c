"""CPU tests of the device AO-integral code: votca_b200/csrc/ao3c_core.cuh is __host__ __device__
and parameterised on the warp barrier, so the very source the sm_100a kernel compiles runs here through a g++-built
harness (tests/host_harness/ao3c_host.cc) - serially, and with one thread per lane + std::barrier, the latter also
under ThreadSanitizer (checks that every shared-scratch dependency crosses a barrier).  Compared with the oracle's
McMurchie-Davidson integrals (oracle/integrals.py), which are pinned on the reference's fixtures up to l = 6.

Replaces on the device: ComputeAO3cBlock (xtp/src/libxtp/libint2_calls.cc:544-593) and AOCoulomb::Fill
(libint2_calls.cc:224-271)."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest
import mpmath

from oracle import basis as obasis
from oracle import integrals
from tests import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_harness", "ao3c_host.cc")
BUILD = os.path.join(HERE, "host_harness", "build")


def _build(name, extra):
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, name)
    subprocess.run(["g++", "-std=c++20", "-fPIC", "-shared", "-pthread", "-o", out, SRC] + extra, check=True)
    return out


def _bind(path):
    lib = ctypes.CDLL(path)
    p, i = ctypes.c_void_p, ctypes.c_int
    lib.ao3c_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, p]
    lib.coulomb2c_host.argtypes = [i, p, p, p, p, p, p]
    lib.ao3c_range_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, i, ctypes.c_long, p]
    lib.ao3c_grid_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, i, ctypes.c_long, p]
    lib.launch_config_host.argtypes = [i, i, i, ctypes.c_long, p]
    lib.dipole_host.argtypes = [i, p, p, p, p, p, p]
    lib.overlap_host.argtypes = [i, p, p, p, p, p, p]
    lib.surviving_pairs_host.argtypes = [i, p, p, p, p, p]
    lib.surviving_pairs_host.restype = ctypes.c_long
    lib.normalize_host.argtypes = [i, i, p, p, p]
    lib.boys_host.argtypes = [i, ctypes.c_double, p]
    lib.pure_matrix_host.argtypes = [i, p]
    return lib


@pytest.fixture(scope="module")
def lib():
    return _bind(_build("libao3c_host.so", ["-O2"]))


def pack(ao):
    """Flat shell arrays of an oracle AOBasis, the argument list of gwbse_basis_create."""
    l = np.array([s.l for s in ao.shells], dtype=np.int32)
    npr = np.array([len(s.exps) for s in ao.shells], dtype=np.int32)
    cen = np.ascontiguousarray(np.array([s.center for s in ao.shells], dtype=np.float64))
    ex = np.concatenate([s.exps for s in ao.shells]).astype(np.float64)
    co = np.concatenate([s.coefs for s in ao.shells]).astype(np.float64)
    return l, npr, cen, ex, co


def _ptrs(arrs):
    return [a.ctypes.data for a in arrs]


def ao3c(lib, aux, dft, nl=1):
    d, a = pack(dft), pack(aux)
    out = np.full((aux.size, dft.size, dft.size), np.nan)
    rc = lib.ao3c_host(len(d[0]), *_ptrs(d), len(a[0]), *_ptrs(a), nl, out.ctypes.data)
    assert rc == 0
    return out


def coulomb2c(lib, ao):
    a = pack(ao)
    out = np.full((ao.size, ao.size), np.nan)
    assert lib.coulomb2c_host(len(a[0]), *_ptrs(a), out.ctypes.data) == 0
    return out


def overlap(lib, ao):
    a = pack(ao)
    out = np.full((ao.size, ao.size), np.nan)
    assert lib.overlap_host(len(a[0]), *_ptrs(a), out.ctypes.data) == 0
    return out


def dipole(lib, ao):
    a = pack(ao)
    out = np.full((3, ao.size, ao.size), np.nan)
    assert lib.dipole_host(len(a[0]), *_ptrs(a), out.ctypes.data) == 0
    return out


def _golden_basis(name, mol):
    g = helpers.load_golden()
    bs = json.loads(str(g[f"basis/{name}.json"]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    return obasis.AOBasis(bs, [str(e) for e in g[f"molecule_{mol}/elements"]], g[f"molecule_{mol}/positions_bohr"])


def relmax(ref, val):
    return np.abs(ref - val).max() / np.abs(ref).max()


def test_boys_function(lib):
    out = ctypes.c_double()
    worst = 0.0
    rng = np.random.default_rng(3)
    xs = np.concatenate([[0.0, 1e-12, 0.05, 0.049999, 35.94, 35.95, 35.96, 36.0, 60.0, 250.0, 4000.0],
                         rng.uniform(0, 40, 200), rng.uniform(0, 1, 50)])
    mpmath.mp.dps = 40  # scipy's hyp1f1 (the oracle's Boys function) is itself only good to 5e-13 at large x
    for n in (0, 1, 2, 5, 8, 10, 14, 16):
        for x in xs:
            lib.boys_host(n, float(x), ctypes.byref(out))
            ref = mpmath.hyp1f1(n + 0.5, n + 1.5, -mpmath.mpf(float(x))) / (2 * n + 1)
            worst = max(worst, float(abs((mpmath.mpf(out.value) - ref) / ref)))
    assert worst < 1e-14, worst


def test_contraction_normalisation_matches_oracle_basis(lib):
    """normalize_contraction == oracle/basis.py (AOShell::LibintShell + normalizeContraction, aoshell.cc:65-89) for the
    contracted s/p/d/f shells of def2-svp / aux-def2-svp and the G / I shells of the large-l fixtures."""
    g = helpers.load_golden()
    n = 0
    for key, mol in (("def2-svp_CH", "methane_tutorial"), ("aux-def2-svp_CH", "methane_tutorial"), ("G", "C2"),
                     ("I", "C2"), ("contracted", "C")):
        raw = json.loads(str(g[f"basis/{key}.json"]))
        ao = _golden_basis(key, mol)
        shells = iter(ao.shells)
        for el in [str(e) for e in g[f"molecule_{mol}/elements"]]:
            for l, prims in raw[el]:
                sh = next(shells)
                ex = np.array([p[0] for p in prims], dtype=np.float64)
                co = np.array([p[1] for p in prims], dtype=np.float64)
                out = np.empty_like(ex)
                lib.normalize_host(int(l), len(ex), ex.ctypes.data, co.ctypes.data, out.ctypes.data)
                assert np.allclose(out, sh.coefs, rtol=1e-13, atol=0), (key, l)
                n += 1
    assert n > 40


def test_pure_matrices_match_oracle(lib):
    for l in range(7):
        T = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
        lib.pure_matrix_host(l, T.ctypes.data)
        assert np.abs(T - integrals.pure_transform(l)).max() < 1e-14


def test_methane_321g_matches_oracle(lib):
    m = helpers.methane_integrals()
    got = ao3c(lib, m["basis"], m["basis"])
    assert relmax(m["ao3c"], got) < 1e-12
    assert relmax(m["V"], coulomb2c(lib, m["basis"])) < 1e-12


def test_water_spdf_aux_matches_oracle(lib):
    w = helpers.water_integrals()
    got = ao3c(lib, w["aux"], w["dft"])
    assert relmax(w["ao3c"], got) < 1e-12
    assert np.abs(got - got.transpose(0, 2, 1)).max() == 0.0 or relmax(got, got.transpose(0, 2, 1)) < 1e-14
    assert relmax(w["V"], coulomb2c(lib, w["aux"])) < 1e-12
    assert relmax(w["S"], overlap(lib, w["aux"])) < 1e-13
    assert relmax(w["S_dft"], overlap(lib, w["dft"])) < 1e-13


def test_function_ranges_cutting_through_shells_and_pitched_blocks(lib):
    """gwbse_ao3c_block takes aux FUNCTION ranges: shells cut by the range are computed but only the functions inside
    are written (to the right slot), for every start / length, also with a padded column pitch."""
    w = helpers.water_integrals()
    d, a = pack(w["dft"]), pack(w["aux"])
    N, naux = w["dft"].size, w["aux"].size
    for f0, f1, pitch in [(0, naux, 0), (3, 10, 0), (0, 1, 0), (naux - 2, naux, 0), (5, 5, 0), (7, 31, N + 1),
                          (1, naux - 1, N + 3)]:
        P = pitch or N
        out = np.full((max(f1 - f0, 0), N, P), np.nan)
        rc = lib.ao3c_range_host(len(d[0]), *_ptrs(d), len(a[0]), *_ptrs(a), f0, f1, pitch, out.ctypes.data)
        assert rc == 0
        if f1 > f0:
            assert relmax(w["ao3c"][f0:f1], out[:, :, :N]) < 1e-12, (f0, f1, pitch)
            assert np.all(out[:, :, N:] == 0.0)


def test_emulated_launch_grid_equals_oracle(lib):
    """The launcher's whole sequence on the CPU - classes, launch geometry, CTA / warp / lane-group indexing of
    ao::cta_thread (the kernel body), group barriers - for the full tensor and for a range that cuts through shells,
    with a padded pitch.  What stays untested without a device are the CUDA runtime calls themselves."""
    w = helpers.water_integrals()
    d, a = pack(w["dft"]), pack(w["aux"])
    N, naux = w["dft"].size, w["aux"].size
    for f0, f1, pitch in [(0, naux, 0), (4, 37, N + 1)]:
        P = pitch or N
        out = np.full((f1 - f0, N, P), np.nan)
        assert lib.ao3c_grid_host(len(d[0]), *_ptrs(d), len(a[0]), *_ptrs(a), f0, f1, pitch, out.ctypes.data) == 0
        assert relmax(w["ao3c"][f0:f1], out[:, :, :N]) < 1e-12
        assert np.all(out[:, :, N:] == 0.0)


def test_launch_geometry_of_every_class(lib):
    """Every class up to (g g | i) gets a launch that fits the 227 KB opt-in shared memory of sm_100 and at most 256
    threads; a lane group has a quarter of the lanes its widest stage has entries (measured on C60 / def2-tzvp:
    2.65 s with one lane per entry, 2.39 s with a quarter), so narrow classes share a warp between 2 - 32 triples;
    small classes keep several CTAs per SM."""
    limit = 232448
    out = (ctypes.c_long * 6)()
    seen = set()
    for la in range(5):
        for lb in range(la + 1):
            for lc in range(7):
                lib.launch_config_host(la, lb, lc, limit, out)
                gl, gpw, wpc, wsd, smem, fits = list(out)
                assert fits == 1 and smem <= limit and smem == 8 * wsd * gpw * wpc, (la, lb, lc)
                assert gl in (1, 2, 4, 8, 16, 32) and gl * gpw == 32 and 1 <= wpc <= 8
                nc = lambda l: (l + 1) * (l + 2) // 2  # noqa: E731
                assert gl == 32 or nc(la) * nc(lb) * nc(lc) <= 4 * gl
                if smem * 4 <= limit:
                    assert wpc == 8 or 2 * smem * 4 > limit  # as many warps as the quarter-SM budget allows
                seen.add(gl)
    assert seen == {1, 4, 8, 16, 32}
    lib.launch_config_host(0, 0, 0, limit, out)
    assert list(out)[:3] == [1, 32, 8]
    lib.launch_config_host(4, 4, 6, limit, out)
    assert list(out)[:3] == [32, 1, 1] and out[4] > 100000


def test_dipole_integrals_match_oracle(lib):
    """<mu | r | nu> about the origin (AODipole, the input of Orbitals::CalcFreeTransition_Dipoles, orbitals.cc:742-760)
    against the oracle, whose dipoles are pinned on the reference's G-shell fixture (test_aomatrix3d.cc:91-123)."""
    w = helpers.water_integrals()
    assert relmax(w["dipole"], dipole(lib, w["dft"])) < 1e-13
    c = helpers.methane_svp_case()
    assert relmax(c["dipole"], dipole(lib, c["dft"])) < 1e-13
    g = _golden_basis("G", "C2")
    assert relmax(integrals.dipole(g), dipole(lib, g)) < 1e-12


def test_methane_def2svp_tier_r_matches_oracle(lib):
    c = helpers.methane_svp_case()  # d orbital shells, d/f aux shells
    got = ao3c(lib, c["aux"], c["dft"])
    assert relmax(c["ao3c"], got) < 1e-12
    assert relmax(c["V"], coulomb2c(lib, c["aux"])) < 1e-12


def test_lanes_with_barriers_equal_serial(lib):
    w = helpers.water_integrals()
    serial = ao3c(lib, w["aux"], w["dft"], nl=1)
    for nl in (2, 4, 5, 8, 16, 32):  # 4 / 8 / 16 / 32 are the lane-group sizes the launcher uses
        assert np.array_equal(serial, ao3c(lib, w["aux"], w["dft"], nl=nl)), nl


def test_g_orbitals_i_aux_large_l(lib):
    """(G G | I): la = lb = 4, lc = 6, L = 14 - the extreme class the reference ships data for
    (test_threecenter_dft.cc:76-115); oracle pinned there in test_oracle_golden.py::test_large_l_integrals."""
    aux, dft = _golden_basis("I", "C2"), _golden_basis("G", "C2")
    ref = integrals.coulomb3c(aux, dft)
    assert relmax(ref, ao3c(lib, aux, dft)) < 1e-11
    assert relmax(integrals.coulomb2c(aux), coulomb2c(lib, aux)) < 1e-11
    assert relmax(integrals.overlap(aux), overlap(lib, aux)) < 1e-12
    assert relmax(integrals.overlap(dft), overlap(lib, dft)) < 1e-12


def test_shell_pair_screening_far_apart_fragments(lib):
    """Two water molecules 60 bohr apart: shell pairs across the gap have no primitive pair above the kernel's
    threshold, are dropped from the launch lists and stay exact zeros; everything still matches the oracle."""
    g = helpers.load_golden()
    bs = json.loads(str(g["basis/water_3-21G.json"]))
    bs = {el: [(int(l), [tuple(p) for p in prims]) for l, prims in shells] for el, shells in bs.items()}
    el = [str(e) for e in g["molecule_water/elements"]]
    pos = np.asarray(g["molecule_water/positions_bohr"])
    dimer = obasis.AOBasis(bs, el + el, np.vstack([pos, pos + np.array([60.0, 0.0, 0.0])]))
    d = pack(dimer)
    ns = len(d[0])
    kept = lib.surviving_pairs_host(ns, *_ptrs(d))
    assert kept == ns * (ns + 1) // 2 - (ns // 2) ** 2  # exactly the cross pairs are gone
    ref = integrals.coulomb3c(dimer, dimer)
    got = ao3c(lib, dimer, dimer)
    assert relmax(ref, got) < 1e-12
    half = dimer.size // 2
    assert np.all(got[:, :half, half:] == 0.0) and np.abs(ref[:, :half, half:]).max() < 1e-30


def _rotated(ao, R, shift):
    import copy
    out = copy.deepcopy(ao)
    for sh in out.shells:
        sh.center = R @ sh.center + shift
    return out


@pytest.mark.parametrize("case", ["methane-def2svp", "G-I"])
def test_rotational_and_translational_invariance(lib, case):
    """Independent of any reference data (SURVEY.md 8c asks for it where libint parity is unpinned, i.e. beyond p
    shells): under a rigid rotation + translation the functions of every pure shell mix by an orthogonal matrix, so
    the Frobenius norm of each (aux shell | shell, shell) block, and the spectra of the overlap and two-centre
    Coulomb matrices, must not change.  Exercises the solid-harmonic transformation and the angular recursions of
    every l combination in the basis (up to f aux / d orbital, and (G G | I))."""
    if case == "methane-def2svp":
        c = helpers.methane_svp_case()
        aux, dft = c["aux"], c["dft"]
    else:
        aux, dft = _golden_basis("I", "C2"), _golden_basis("G", "C2")
    rng = np.random.default_rng(11)
    R = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    if np.linalg.det(R) < 0:
        R[:, 0] *= -1.0
    shift = np.array([0.7, -1.3, 2.1])
    aux2, dft2 = _rotated(aux, R, shift), _rotated(dft, R, shift)
    T0, T1 = ao3c(lib, aux, dft), ao3c(lib, aux2, dft2)
    worst, scale = 0.0, np.abs(T0).max()
    for sc in aux.shells:
        for sa in dft.shells:
            for sb in dft.shells:
                blk = (slice(sc.start, sc.start + sc.nfunc), slice(sa.start, sa.start + sa.nfunc),
                       slice(sb.start, sb.start + sb.nfunc))
                n0, n1 = np.linalg.norm(T0[blk]), np.linalg.norm(T1[blk])
                # blocks that vanish by symmetry are 1e-14 noise: measure against the block or 1e-3 of the largest
                worst = max(worst, abs(n0 - n1) / max(n0, 1e-3 * scale))
    assert worst < 1e-10, worst
    assert not np.allclose(T0, T1, atol=1e-3)  # the individual integrals do change
    for f in (overlap, coulomb2c):
        w0, w1 = np.linalg.eigvalsh(f(lib, aux)), np.linalg.eigvalsh(f(lib, aux2))
        assert np.abs(w0 - w1).max() < 1e-10 * np.abs(w0).max()


def test_invariance_benzene_def2_tzvp_every_class_up_to_ff_g(lib):
    """BASELINE config 1 basis (benzene, def2-tzvp + aux-def2-tzvp: s..f orbital shells, s..g aux shells, 27 M
    integrals): every angular-momentum class the benchmark molecules produce, checked by rigid-motion invariance of the
    per-shell-triple norms, permutational symmetry, and positive definiteness of the two-centre metric."""
    from votca_b200 import realsys
    el, pos = realsys.benzene()
    rng = np.random.default_rng(21)
    R = np.linalg.qr(rng.standard_normal((3, 3)))[0]
    if np.linalg.det(R) < 0:
        R[:, 0] *= -1.0
    pos2 = pos @ R.T + np.array([1.1, -0.4, 0.9])
    res = []
    for p in (pos, pos2):
        d, a = realsys.shell_arrays("def2-tzvp", el, p), realsys.shell_arrays("aux-def2-tzvp", el, p)
        N, naux = realsys.nfunc(d[0]), realsys.nfunc(a[0])
        out = np.empty((naux, N, N))
        assert lib.ao3c_host(len(d[0]), *_ptrs(d), len(a[0]), *_ptrs(a), 1, out.ctypes.data) == 0
        res.append((out, d, a))
    (T0, d, a), (T1, _, _) = res
    assert (T0.shape[1], T0.shape[0]) == (222, 546) and max(d[0]) == 3 and max(a[0]) == 4
    assert np.abs(T0 - T0.transpose(0, 2, 1)).max() < 1e-13 * np.abs(T0).max()
    fo = np.concatenate([[0], np.cumsum(2 * d[0] + 1)])
    fa = np.concatenate([[0], np.cumsum(2 * a[0] + 1)])
    n0 = np.add.reduceat(np.add.reduceat(np.add.reduceat(T0 ** 2, fa[:-1], 0), fo[:-1], 1), fo[:-1], 2)
    n1 = np.add.reduceat(np.add.reduceat(np.add.reduceat(T1 ** 2, fa[:-1], 0), fo[:-1], 1), fo[:-1], 2)
    scale = n0.max()
    assert np.abs(np.sqrt(n0) - np.sqrt(n1)).max() < 1e-10 * np.sqrt(scale)
    classes = {(int(lc), int(max(la, lb)), int(min(la, lb))) for lc in a[0] for la in d[0] for lb in d[0]}
    assert (4, 3, 3) in classes and len(classes) == 5 * 10
    V = np.full((naux, naux), np.nan)  # two-centre metric of the aux basis: symmetric positive definite
    assert lib.coulomb2c_host(len(a[0]), *_ptrs(a), V.ctypes.data) == 0
    assert np.abs(V - V.T).max() < 1e-12 * np.abs(V).max() and np.linalg.eigvalsh(V).min() > 0.0


def test_thread_sanitizer_finds_no_race():
    """Same source, 8 lanes on threads, under -fsanitize=thread: a missing barrier between two stages that share
    scratch would be reported as a data race (TSan exits non-zero)."""
    so = _build("libao3c_host_tsan.so", ["-O1", "-g", "-fsanitize=thread"])
    m = helpers.methane_integrals()
    d = pack(m["basis"])
    np.savez(os.path.join(BUILD, "tsan_in.npz"), l=d[0], np_=d[1], cen=d[2], ex=d[3], co=d[4])
    code = (
        "import ctypes, numpy as np, sys\n"
        f"z = np.load(r'{os.path.join(BUILD, 'tsan_in.npz')}')\n"
        f"lib = ctypes.CDLL(r'{so}')\n"
        "p, i = ctypes.c_void_p, ctypes.c_int\n"
        "lib.ao3c_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, p]\n"
        "a = [np.ascontiguousarray(z[k]) for k in ('l', 'np_', 'cen', 'ex', 'co')]\n"
        "ns = len(a[0]); n = int((2 * a[0] + 1).sum())\n"
        "out = np.zeros((n, n, n))\n"
        "pt = [x.ctypes.data for x in a]\n"
        "rc = lib.ao3c_host(ns, *pt, ns, *pt, 8, out.ctypes.data)\n"
        "lib.ao3c_grid_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, i, i, ctypes.c_long, p]\n"
        "rc |= lib.ao3c_grid_host(ns, *pt, ns, *pt, 0, n, 0, out.ctypes.data)  # lane groups of 4 / 16 / 32 side by side\n"
        "sys.exit(rc)\n")
    tsan_rt = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=tsan_rt, TSAN_OPTIONS="exitcode=66 halt_on_error=1")
    r = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-3000:]
    assert r.returncode == 0, (r.returncode, r.stderr[-3000:])


def test_scratch_bounds_under_address_sanitizer():
    """Each shell triple gets a heap scratch of exactly workspace_doubles(la, lb, lc) - the size the launcher reserves
    per lane group in shared memory - and the harness is built with -fsanitize=address: an access past that region in
    any stage (three-centre, two-centre, overlap, dipole modes; s..f orbital, s..f aux, and the (G G | I) class) would
    abort the child process."""
    so = _build("libao3c_host_asan.so", ["-O1", "-g", "-fsanitize=address", "-fno-omit-frame-pointer"])
    w, c = helpers.water_integrals(), helpers.methane_svp_case()
    cases = {"water": (pack(w["dft"]), pack(w["aux"])), "methane": (pack(c["dft"]), pack(c["aux"])),
             "gi": (pack(_golden_basis("G", "C2")), pack(_golden_basis("I", "C2")))}
    arrs = {f"{k}_{s}_{i}": a for k, (d, x) in cases.items() for s, t in (("d", d), ("a", x)) for i, a in enumerate(t)}
    np.savez(os.path.join(BUILD, "asan_in.npz"), **arrs)
    code = (
        "import ctypes, numpy as np, sys\n"
        f"z = np.load(r'{os.path.join(BUILD, 'asan_in.npz')}')\n"
        f"lib = ctypes.CDLL(r'{so}')\n"
        "p, i = ctypes.c_void_p, ctypes.c_int\n"
        "lib.ao3c_exact_scratch_host.argtypes = [i, p, p, p, p, p, i, p, p, p, p, p, p]\n"
        "for k in ('water', 'methane', 'gi'):\n"
        "    d = [np.ascontiguousarray(z[f'{k}_d_{j}']) for j in range(5)]\n"
        "    a = [np.ascontiguousarray(z[f'{k}_a_{j}']) for j in range(5)]\n"
        "    N, M = int((2 * d[0] + 1).sum()), int((2 * a[0] + 1).sum())\n"
        "    out = np.zeros((M, N, N))\n"
        "    rc = lib.ao3c_exact_scratch_host(len(d[0]), *[x.ctypes.data for x in d], len(a[0]), *[x.ctypes.data for x in a], out.ctypes.data)\n"
        "    assert rc == 0 and np.isfinite(out).all() and np.abs(out).max() > 0\n"
        "print('asan clean')\n")
    rt = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, LD_PRELOAD=rt, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1")
    r = subprocess.run([os.sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "asan clean" in r.stdout and "AddressSanitizer" not in r.stderr, r.stderr[-3000:]
