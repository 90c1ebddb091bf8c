"""CPU tests of the .orb (HDF5) results writer, votca_b200/host/checkpoint.h, through a g++ harness.

No HDF5 library exists in this environment, so the evidence is: (1) the lookup3 checksum restated in
oracle/orbfile.py reproduces every superblock / object-header checksum libhdf5 wrote into the reference's own .orb
files, and the C++ checksum equals it; (2) files written by the C++ writer parse with the reader that is pinned on
those reference files, all their checksums verify, and (3) every numeric dataset and attribute of a reference
checkpoint replayed through the writer reads back identical.  Reference call surface:
xtp/include/votca/xtp/checkpointwriter.h:49-353, Orbitals::WriteToCpt (orbitals.cc:990-1063)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.orbfile import OrbFile, lookup3

HERE = os.path.dirname(os.path.abspath(__file__))
IT = "/root/reference/xtp/src/tests/DataFiles/xtp_tools_integration_tests"


@pytest.fixture(scope="module")
def lib():
    src = os.path.join(HERE, "host_harness", "cpt_harness.cc")
    out = os.path.join(HERE, "host_harness", "build", "libcpt_harness.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wall", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    p, l, d, s, i = ctypes.c_void_p, ctypes.c_long, ctypes.c_double, ctypes.c_char_p, ctypes.c_int
    lib.cpt_open.restype = p
    lib.cpt_open.argtypes = [s]
    lib.cpt_close.argtypes = [p]
    lib.cpt_attr.argtypes = [p, s, s, i, l, d, s]
    lib.cpt_dataset.argtypes = [p, s, s, l, l, p, i]
    lib.cpt_eigensystem.argtypes = [p, s, s, l, l, p, p, l, p, l]
    lib.cpt_vec3list.argtypes = [p, s, s, l, p]
    lib.cpt_group.argtypes = [p, s]
    lib.cpt_lookup3.argtypes = [p, l]
    lib.cpt_lookup3.restype = ctypes.c_uint
    return lib


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


def _groups(f, p):
    yield p
    for k in f.keys(p):
        c = p.rstrip("/") + "/" + k
        if f.is_group(c):
            yield from _groups(f, c)


def test_lookup3_cpp_equals_python_and_known_values(lib):
    rng = np.random.default_rng(5)
    for n in (0, 1, 3, 4, 11, 12, 13, 24, 25, 44, 257):
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert lib.cpt_lookup3(buf.ctypes.data, n) == lookup3(bytes(buf))
    # lookup3.c self-test vector: hashlittle("Four score and seven years ago", 30, 0) = 0x17770551
    assert lookup3(b"Four score and seven years ago") == 0x17770551


@pytest.mark.skipif(not os.path.isdir(IT), reason="reference tree not available")
def test_python_lookup3_reproduces_libhdf5_checksums():
    for fn in ("molecule_neutral.orb", "molecule_neutral_tda.orb", "molecule_ch4.orb", "molecule_cation.orb"):
        assert OrbFile(os.path.join(IT, fn)).verify_checksums() >= 60


def test_results_file_layout(lib, tmp_path):
    """What GWBSE results look like in the file: names and types of Orbitals::WriteToCpt (orbitals.cc:1031-1062)."""
    path = str(tmp_path / "results.orb").encode()
    rng = np.random.default_rng(1)
    h = lib.cpt_open(path)
    q, B, k = 7, 12, 3
    assert lib.cpt_attr(h, b"/QMdata", b"XTPVersion", 4, 0, 0.0, b"gwbse-b200") == 0
    assert lib.cpt_attr(h, b"/QMdata", b"version", 0, 8, 0.0, None) == 0
    for name, v in (("rpamin", 0), ("rpamax", 20), ("qpmin", 2), ("qpmax", 8), ("bse_vmin", 2), ("bse_cmax", 8)):
        assert lib.cpt_attr(h, b"/QMdata", name.encode(), 1, v, 0.0, None) == 0
    assert lib.cpt_attr(h, b"/QMdata", b"useTDA", 5, 0, 0.0, None) == 0
    assert lib.cpt_attr(h, b"/QMdata", b"use_Hqp_offdiag", 2, 1, 0.0, None) == 0
    assert lib.cpt_attr(h, b"/QMdata", b"ScaHFX", 3, 0, 0.25, None) == 0
    assert lib.cpt_attr(h, b"/QMdata", b"ECP", 4, 0, 0.0, b"") == 0
    qp = _f(rng.standard_normal(q))
    assert lib.cpt_dataset(h, b"/QMdata", b"QPpert_energies", q, 1, qp.ctypes.data, 1) == 0
    ev, V1, V2 = _f(np.sort(rng.random(k))), _f(rng.standard_normal((B, k))), _f(rng.standard_normal((B, k)))
    assert lib.cpt_eigensystem(h, b"/QMdata", b"BSE_singlet", B, k, ev.ctypes.data, V1.ctypes.data, k, V2.ctypes.data, 0) == 0
    assert lib.cpt_eigensystem(h, b"/QMdata", b"BSE_triplet", 0, 0, ev.ctypes.data, V1.ctypes.data, 0, V2.ctypes.data, 0) == 0
    td = np.ascontiguousarray(rng.standard_normal((k, 3)))
    assert lib.cpt_vec3list(h, b"/QMdata", b"transition_dipoles", k, td.ctypes.data) == 0
    M = _f(rng.standard_normal((5, 4)))
    assert lib.cpt_dataset(h, b"/QMdata/nested/deeper", b"M", 5, 4, M.ctypes.data, 0) == 0
    assert lib.cpt_attr(h, b"/QMdata", b"qpmax", 1, 9, 0.0, None) == 0  # reopened attribute is overwritten
    assert lib.cpt_close(h) == 0

    f = OrbFile(path.decode())
    assert f.verify_checksums() >= 12
    assert f.keys("/") == ["QMdata"]
    assert set(f.keys("/QMdata")) == {"QPpert_energies", "BSE_singlet", "BSE_triplet", "transition_dipoles", "nested"}
    at = f.attrs("/QMdata")
    assert at["XTPVersion"] == "gwbse-b200" and at["ECP"] == "" and at["version"] == 8 and at["version"].dtype == np.int32
    assert at["qpmax"] == 9 and at["qpmax"].dtype == np.int64 and at["useTDA"] == 0 and at["useTDA"].dtype == np.int64
    assert at["use_Hqp_offdiag"] == 1 and at["use_Hqp_offdiag"].dtype == np.uint8 and at["ScaHFX"] == 0.25
    got = f.read("/QMdata/QPpert_energies")
    assert got.shape == (q, 1) and np.array_equal(got.ravel(), qp)
    assert f.keys("/QMdata/BSE_singlet") == ["eigenvalues", "eigenvectors", "eigenvectors2"]
    assert f.attrs("/QMdata/BSE_singlet")["info"] == 0
    assert np.array_equal(f.read("/QMdata/BSE_singlet/eigenvectors"), V1)       # file (rows, cols) in C order
    assert np.array_equal(f.read("/QMdata/BSE_singlet/eigenvectors2"), V2)
    assert f.read("/QMdata/BSE_triplet/eigenvalues").shape == (0, 1)            # empty: dims {0, 1}, no storage
    assert f.read("/QMdata/BSE_triplet/eigenvectors").size == 0
    assert f.keys("/QMdata/transition_dipoles") == ["ind0", "ind1", "ind2"]
    assert np.array_equal(f.read("/QMdata/transition_dipoles/ind1").ravel(), td[1])
    assert np.array_equal(f.read("/QMdata/nested/deeper/M"), M)


@pytest.mark.skipif(not os.path.isdir(IT), reason="reference tree not available")
@pytest.mark.parametrize("fn", ["molecule_neutral.orb", "molecule_ch4.orb"])
def test_replay_of_a_reference_checkpoint_reads_back_identical(lib, tmp_path, fn):
    """Every float dataset and every attribute of the reference's own checkpoint, written again by the C++ writer
    (the compound tables of atoms / basis shells are DFT-side inputs and are skipped), reads back identical, and
    the object-header messages of a dataset are byte-identical to the ones libhdf5 produced."""
    src = OrbFile(os.path.join(IT, fn))
    path = str(tmp_path / "copy.orb")
    h = lib.cpt_open(path.encode())
    kinds = {np.dtype("<i4"): 0, np.dtype("<i8"): 1, np.dtype("u1"): 2, np.dtype("<f8"): 3}
    copied, skipped = [], []

    def groups(p):
        yield p
        for k in src.keys(p):
            c = p.rstrip("/") + "/" + k
            if src.is_group(c):
                yield from groups(c)

    for g in groups("/"):
        assert lib.cpt_group(h, g.encode()) == 0  # openChild: empty groups exist too
        for name, val in src.attrs(g).items():
            if isinstance(val, str):
                assert lib.cpt_attr(h, g.encode(), name.encode(), 4, 0, 0.0, val.encode()) == 0
            elif val is not None and np.ndim(val) == 0:
                kind = kinds[np.asarray(val).dtype]
                assert lib.cpt_attr(h, g.encode(), name.encode(), kind, int(val) if kind != 3 else 0,
                                    float(val) if kind == 3 else 0.0, None) == 0
    for ds in src.walk("/"):
        try:
            a = src.read(ds)
        except NotImplementedError:
            skipped.append(ds)
            continue
        if a.dtype != np.dtype("<f8") or a.ndim != 2:
            skipped.append(ds)
            continue
        g, name = ds.rsplit("/", 1)
        af = _f(a)
        assert lib.cpt_dataset(h, (g or "/").encode(), name.encode(), a.shape[0], a.shape[1], af.ctypes.data, 0) == 0
        copied.append(ds)
    assert lib.cpt_close(h) == 0
    assert len(copied) >= 25 and all(s.rsplit("/", 1)[1] in ("qmatoms", "Shells", "Contractions") for s in skipped), skipped

    dst = OrbFile(path)
    dst.verify_checksums()
    assert sorted(groups("/")) == sorted(g for g in _groups(dst, "/"))
    for ds in copied:
        assert np.array_equal(src.read(ds), dst.read(ds)), ds
    for g in groups("/"):
        a0, a1 = src.attrs(g), dst.attrs(g)
        for name, val in a0.items():
            if val is None or np.ndim(val):
                continue
            assert name in a1 and a1[name] == val and type(a1[name]) is type(val), (g, name)
    # message encodings of a dataset header: dataspace, datatype, fill value equal libhdf5's bytes
    ma = {t: (fl, body) for t, fl, body in src._messages(src._resolve("/QMdata/mos/eigenvectors"))}
    mb = {t: (fl, body) for t, fl, body in dst._messages(dst._resolve("/QMdata/mos/eigenvectors"))}
    for t in (0x01, 0x03, 0x05):
        assert ma[t] == mb[t], hex(t)
    assert ma[0x08][1][:2] == mb[0x08][1][:2] and ma[0x08][1][10:] == mb[0x08][1][10:]  # layout: version, class, size
