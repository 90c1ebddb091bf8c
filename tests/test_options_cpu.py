"""CPU test of the options layer of the C++ host (votca_b200/host/gwbse.h: Options), through a g++ harness: the
defaults equal the reference's share/xtp/xml/subpackages/gwbse.xml, and the option files of the reference's
dftgwbse integration tests load to the same key/value pairs (north star: "the same options XML").
Fixture: tests/golden/gwbse_xml_options.json <- tests/golden/make_options_fixture.py."""
import ctypes
import json
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
IT = "/root/reference/xtp/src/tests/DataFiles/xtp_tools_integration_tests"

# options of gwbse.xml without a default value to compare (fragment lists, the aux basis name)
NOT_CONSUMED = ("bse.fragments", "auxbasisset")


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_harness", "options_harness.cc")
    out = os.path.join(HERE, "host_harness", "build", "liboptions_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O0", "-fPIC", "-shared", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    p, s = ctypes.c_void_p, ctypes.c_char_p
    lib.opt_new.restype = p
    lib.opt_free.argtypes = [p]
    lib.opt_load_xml.argtypes = [p, s]
    lib.opt_set.argtypes = [p, s, s]
    lib.opt_get.argtypes = [p, s, ctypes.c_char_p, ctypes.c_int]
    L = ctypes.c_long
    lib.rpa_update_energies.argtypes = [L, L, L, p, L, p, L, L, p]
    lib.bse_ranked_guess.argtypes = [p, p, L, L, p, L]
    lib.bse_ranked_guess.restype = L
    lib.gwbse_write_results.argtypes = [s, L, L, L, L, L, ctypes.c_int]
    lib.gwbse_initialize_ranges.argtypes = [p, L, L, p, p, ctypes.c_char_p, ctypes.c_int]
    lib.gwbse_initialize_ranges_z.argtypes = [p, L, L, p, L, p, ctypes.c_char_p, ctypes.c_int]
    return lib


@pytest.fixture(scope="module")
def fixture():
    with open(os.path.join(HERE, "golden", "gwbse_xml_options.json")) as fh:
        return json.load(fh)


def get(lib, o, key):
    buf = ctypes.create_string_buffer(256)
    return buf.value.decode() if lib.opt_get(o, key.encode(), buf, 256) == 0 else None


def test_defaults_equal_gwbse_xml(lib, fixture):
    o = lib.opt_new()
    checked = 0
    for key, default in fixture["defaults"].items():
        if key.startswith(NOT_CONSUMED):
            continue
        if default in ("", "OPTIONAL", "REQUIRED"):
            assert get(lib, o, key) is None, key  # unset unless the user gives it
            continue
        assert get(lib, o, key) == default, key
        checked += 1
    assert checked >= 30
    lib.opt_free(o)


def _nested_xml(values):
    tree = {}
    for k, v in values.items():
        node = tree
        parts = k.split(".")
        for part in parts[:-1]:
            node = node.setdefault(part, {})
        node[parts[-1]] = v

    def emit(node):
        return "".join(f"<{k}>{emit(v) if isinstance(v, dict) else v}</{k}>" for k, v in node.items())
    return f"<options><dftgwbse><tasks>input,dft,parse,gwbse</tasks><gwbse>{emit(tree)}</gwbse></dftgwbse></options>"


def test_reference_option_files_load(lib, fixture, tmp_path):
    for fn, values in fixture["files"].items():
        path = os.path.join(IT, fn)  # the reference's own file where it exists, else the same content re-nested
        if not os.path.exists(path):
            path = str(tmp_path / fn)
            with open(path, "w") as fh:
                fh.write(_nested_xml(values))
        o = lib.opt_new()
        assert lib.opt_load_xml(o, path.encode()) == 0
        for key, val in values.items():
            assert get(lib, o, key) == val, (fn, key)
        # untouched keys keep their defaults; the dftgwbse-level <tasks> must not leak into gwbse.tasks
        assert get(lib, o, "gw.qp_solver") == "grid" and get(lib, o, "tasks") == values.get("tasks", "all")
        lib.opt_free(o)


def test_dotted_keys_with_reference_prefixes(lib):
    o = lib.opt_new()
    for key in ("options.dftgwbse.gwbse.gw.mode", "dftgwbse.gwbse.gw.mode", "gwbse.gw.mode", "gw.mode"):
        assert lib.opt_set(o, key.encode(), b"G0W0") == 0
        assert get(lib, o, "gw.mode") == "G0W0"
        lib.opt_set(o, b"gw.mode", b"evGW")
    assert lib.opt_load_xml(o, b"/nonexistent/options.xml") == 1
    lib.opt_free(o)


def test_unknown_and_unsupported_keys_are_rejected(lib):
    """A key that gwbse.xml does not have is an error (the reference validates options against that file), and a key
    it has but this path does not implement must not be swallowed: silently different physics is worse than a throw."""
    o = lib.opt_new()
    assert lib.opt_set(o, b"gw.moed", b"G0W0") == 1  # misspelled
    assert lib.opt_set(o, b"bse.davidson.tolerence", b"strict") == 1
    assert lib.opt_set(o, b"bse.fragments.fragment.indices", b"1:3;4 5") == 0  # GWBSE::FragmentPopulations
    assert lib.opt_set(o, b"bse.fragments.fragment.name", b"donor") == 1
    assert lib.opt_set(o, b"gw.sigma_plot.states", b"1 3 5") == 0  # GW::PlotSigma (tests/test_zz_gpu_host_options.py)
    assert lib.opt_set(o, b"gw.sigma_plot.steps", b"201") == 0
    assert lib.opt_set(o, b"gw.qsgw_max_iterations", b"7") == 0 and get(lib, o, "gw.qsgw_max_iterations") == "7"
    lib.opt_free(o)
    assert "exciton" in _ranges(lib, 4, 17, tasks="gw,excitons")
    assert "exciton" in _ranges(lib, 4, 17, tasks="exciton_uks")


def test_rpa_update_input_energies_known_answer(lib):
    """test_rpa.cc:41-67 on the C++ host class (RPA::UpdateRPAInputEnergies, rpa.cc:32-70): GW energies replace the
    window, levels outside it are shifted by the largest correction at the respective edge."""
    import numpy as np
    dft = np.array([-0.5, -0.4, -0.3, -0.2, -0.2, -0.1, 0, 0.1, 0.2, 0.3])
    gw = np.array([-0.15, -0.05, 0.05, 0.15, 0.45, 0.55, 0.65])
    out = np.zeros(10)
    n = lib.rpa_update_energies(4, 0, 9, dft.ctypes.data, 10, gw.ctypes.data, 7, 1, out.ctypes.data)
    assert n == 10
    ref = np.array([-0.85, -0.15, -0.05, 0.05, 0.15, 0.45, 0.55, 0.65, 0.75, 0.85])
    assert np.linalg.norm(out - ref) < 1e-4 * np.linalg.norm(ref)


def _ranges(lib, homo, nlevels, **opts):
    import numpy as np
    o = lib.opt_new()
    for k, v in opts.items():
        lib.opt_set(o, k.replace("__", ".").encode(), str(v).encode())
    r = np.zeros(6, dtype=np.int64)
    nmax = ctypes.c_long()
    err = ctypes.create_string_buffer(256)
    rc = lib.gwbse_initialize_ranges(o, homo, nlevels, r.ctypes.data, ctypes.byref(nmax), err, 256)
    lib.opt_free(o)
    return (tuple(int(x) for x in r), nmax.value) if rc == 0 else err.value.decode()


def test_level_ranges_as_gwbse_initialize(lib):
    """GWBSE::Initialize (gwbse.cc:60-233) on the C++ host: the four `ranges` modes, clamping, exctotal against the BSE
    size, and the reference's error texts.  Water of the reference's integration tests (13 levels, homo 4, default
    ranges) must give the ranges stored in its checkpoint: rpa 0..12, qp 0..12, bse 0..12."""
    assert _ranges(lib, 4, 13) == ((0, 12, 0, 12, 0, 12), 10)
    # default: qpmax = bse_cmax = 3 homo + 1 (SURVEY.md section 8 sizes: DCV5T homo 143 -> 431 levels in the window)
    assert _ranges(lib, 143, 1249)[0] == (0, 1248, 0, 430, 0, 430)
    assert _ranges(lib, 4, 17, ranges="full") == ((0, 16, 0, 16, 0, 16), 10)
    assert _ranges(lib, 4, 17, ranges="explicit", rpamax=14, qpmin=1, qpmax=9, bsemin=2, bsemax=40)[0] == (0, 14, 1, 9, 2, 16)
    assert _ranges(lib, 9, 40, ranges="factor", rpamax=0.5, qpmin=0.5, qpmax=0.5, bsemin=0.2, bsemax=1.0)[0] == \
        (0, 19, 4, 14, 7, 19)
    assert _ranges(lib, 4, 17, bse__exctotal=1000)[1] == 5 * 9  # capped at the number of transitions
    assert "unknown ranges" in _ranges(lib, 4, 17, ranges="nonsense")
    assert "Invalid GW level range" in _ranges(lib, 4, 17, ranges="explicit", rpamax=16, qpmin=9, qpmax=3, bsemin=0, bsemax=8)
    assert "must be given together" in _ranges(lib, 4, 17, gw__qp_grid_steps=201)
    assert "unknown gw.mode" in _ranges(lib, 4, 17, gw__mode="scGW")


def test_full_bse_ranked_initial_guess_equals_oracle(lib):
    """BuildFullBSEXRankedInitialGuess (bse_initialization.h:47-93): max(4 nroots, 8) unit vectors on the X block,
    ranked by sqrt(a^2 - b^2) with ties broken by a."""
    import numpy as np
    from oracle.bse import build_full_bse_x_ranked_initial_guess
    rng = np.random.default_rng(4)
    for n, nroots in ((45, 3), (12, 5), (6, 1), (30, 10)):
        a = np.round(rng.uniform(0.2, 1.5, n), 2)  # rounding makes ties
        b = rng.uniform(-0.2, 0.2, n)
        ref = build_full_bse_x_ranked_initial_guess(a, b, nroots)
        out = np.zeros(ref.size)
        k = lib.bse_ranked_guess(a.ctypes.data, b.ctypes.data, n, nroots, out.ctypes.data, out.size)
        assert k == ref.shape[1]
        assert np.array_equal(out.reshape(ref.shape[1], ref.shape[0]).T, ref)


def test_results_checkpoint_uses_the_names_and_types_of_the_reference(lib, tmp_path):
    """GWBSE::WriteToCpt (the GW-BSE part of Orbitals::WriteToCpt, orbitals.cc:990-1063): every attribute and dataset
    it writes exists under the same path, with the same HDF5 type and rank, in the .orb files the reference's own
    dftgwbse run produced (names / types recorded from molecule_neutral.orb and molecule_neutral_tda.orb)."""
    import numpy as np
    from oracle.orbfile import OrbFile
    path = tmp_path / "res.orb"
    assert lib.gwbse_write_results(str(path).encode(), 13, 4, 13, 40, 5, 0) == 0
    f = OrbFile(str(path))
    f.verify_checksums()
    # (name, numpy kind + size) of the scalar attributes of /QMdata in the reference's files
    ref_attrs = {"XTPVersion": "str", "version": "i4", "occupied_levels": "i8", "number_alpha_electrons": "i8",
                 "number_beta_electrons": "i8", "rpamin": "i8", "rpamax": "i8", "qpmin": "i8", "qpmax": "i8",
                 "bse_vmin": "i8", "bse_cmax": "i8", "ScaHFX": "f8", "useTDA": "i8", "use_Hqp_offdiag": "u1",
                 # members of an Orbitals object without unrestricted / embedding data (defaults, as in the reference's files)
                 "occupied_levels_beta": "i8", "charge": "i8", "spin": "i8", "active_electrons": "i8", "qm_energy": "f8",
                 "qm_package": "str", "XCFunctional": "str", "XC_grid_quality": "str", "ECP": "str", "CalcType": "str"}
    at = f.attrs("/QMdata")
    for name, val in at.items():
        if name == "is_qsgw":  # orbitals_version 9 (orbitals.h:842-844); the checked-in files are version 8
            assert np.asarray(val).dtype == np.uint8
            continue
        kind = ref_attrs[name]
        assert (isinstance(val, str) and kind == "str") or np.asarray(val).dtype == np.dtype(kind), name
    assert set(ref_attrs) <= set(at)
    ref_sets = {"RPA_inputenergies": (13, 1), "QPpert_energies": (13, 1), "BSE_singlet_dynamic": (5, 1),
                "BSE_triplet_dynamic": (0, 1), "mos/eigenvalues": (13, 1), "mos/eigenvectors": (13, 13),
                "mos/eigenvectors2": (0, 1), "QPdiag/eigenvalues": (13, 1), "QPdiag/eigenvectors": (13, 13),
                "QPdiag/eigenvectors2": (0, 1), "BSE_singlet/eigenvalues": (5, 1), "BSE_singlet/eigenvectors": (40, 5),
                "BSE_singlet/eigenvectors2": (40, 5), "BSE_triplet/eigenvalues": (0, 1),
                "BSE_triplet/eigenvectors": (0, 1), "BSE_triplet/eigenvectors2": (0, 1)}
    ref_sets.update({f"transition_dipoles/ind{i}": (3, 1) for i in range(5)})
    empty = (0, 1)  # how the reference's CheckpointWriter stores an empty Eigen object
    for grp in ("mos_beta", "mos_embedding", "QPdiag_alpha", "QPdiag_beta", "BSE_uks"):
        ref_sets.update({f"{grp}/eigenvalues": empty, f"{grp}/eigenvectors": empty, f"{grp}/eigenvectors2": empty})
    for name in ("occupations", "LMOs", "LMOs_energies", "inactivedensity", "TruncMOsFullBasis", "RPA_inputenergies_alpha",
                 "RPA_inputenergies_beta", "QPpert_energies_alpha", "QPpert_energies_beta", "BSE_uks_dynamic"):
        ref_sets[name] = empty
    got = {p[len("/QMdata/"):]: f.read(p).shape for p in f.walk("/QMdata")}
    assert got == ref_sets
    for grp in ("mos", "QPdiag", "BSE_singlet", "BSE_triplet", "mos_beta", "BSE_uks"):
        assert f.attrs("/QMdata/" + grp)["info"].dtype == np.int64
    assert f.read("/QMdata/BSE_singlet/eigenvectors2")[0, 0] == 7.5 and at["useTDA"] == 0 and at["occupied_levels"] == 5
    IT = "/root/reference/xtp/src/tests/DataFiles/xtp_tools_integration_tests/molecule_neutral.orb"
    if os.path.exists(IT):  # the expectations above are the reference file's own
        r = OrbFile(IT)
        rat = r.attrs("/QMdata")
        for name, kind in ref_attrs.items():
            assert (isinstance(rat[name], str) and kind == "str") or np.asarray(rat[name]).dtype == np.dtype(kind), name
        rsets = {p[len("/QMdata/"):] for p in r.walk("/QMdata")}
        assert {k for k in ref_sets if not k.startswith("transition_dipoles/")} <= rsets
        for k, shape in ref_sets.items():  # the empties are empty in the reference's file too
            if shape == empty and not k.startswith(("BSE_triplet", "QPdiag/", "mos/")):
                assert r.read("/QMdata/" + k).shape == empty, k
        # what the stand-alone writer leaves out are exactly the compound tables of the DFT side
        assert {k.split("/")[0] for k in rsets - set(ref_sets) if not k.startswith("transition_dipoles/")} <= \
            {"qmmolecule", "dft", "aux", "forces"}
        assert r.read("/QMdata/BSE_singlet/eigenvectors").shape[0] == r.read("/QMdata/BSE_singlet/eigenvectors2").shape[0]


def test_ignore_corelevels(lib):
    """gwbse.cc:46-58, 128-152: half the core electrons of corelevels.xml (C, N, O, F 2; Al, S 10; H 0) are cut from
    the RPA / GW / BSE windows; thiophene C4H4S has 4 + 5 = 9 core levels."""
    import numpy as np
    z = np.array([16.0] + [6.0] * 4 + [1.0] * 4)
    expect = {"none": (0, 0, 0), "BSE": (0, 0, 9), "GW": (0, 9, 9), "RPA": (9, 9, 9)}
    for mode, (rpamin, qpmin, vmin) in expect.items():
        o = lib.opt_new()
        assert lib.opt_set(o, b"ignore_corelevels", mode.encode()) == 0
        r = (ctypes.c_long * 6)()
        err = ctypes.create_string_buffer(512)
        rc = lib.gwbse_initialize_ranges_z(o, 21, 120, z.ctypes.data_as(ctypes.c_void_p), len(z), r, err, 512)
        assert rc == 0, err.value
        assert (r[0], r[2], r[4]) == (rpamin, qpmin, vmin), (mode, list(r))
        assert (r[1], r[3], r[5]) == (119, 3 * 21 + 1, 3 * 21 + 1)
        lib.opt_free(o)
    o = lib.opt_new()
    lib.opt_set(o, b"ignore_corelevels", b"GW")
    r = (ctypes.c_long * 6)()
    err = ctypes.create_string_buffer(512)
    assert lib.gwbse_initialize_ranges_z(o, 21, 120, None, 0, r, err, 512) == 1 and b"nuclear charges" in err.value
    zz = np.array([26.0])
    assert lib.gwbse_initialize_ranges_z(o, 21, 120, zz.ctypes.data_as(ctypes.c_void_p), 1, r, err, 512) == 1
    assert b"corelevels table" in err.value
    lib.opt_set(o, b"ignore_corelevels", b"sometimes")
    assert lib.gwbse_initialize_ranges_z(o, 21, 120, None, 0, r, err, 512) == 1
    lib.opt_free(o)
