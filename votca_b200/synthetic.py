"""Synthetic GW-BSE inputs (tier S) for benchmarks and size-scaling tests.

Not part of the product path: it only manufactures inputs of the right shape and conditioning
(orthonormal MOs, gapped spectrum, band-decaying symmetric AO three-centre tensor, SPD aux metric).
"""
import numpy as np

CONFIGS = {
    # name: (nbasis, naux, homo)   sizes of SURVEY.md section 8 table (default ranges: qpmax = cmax = 3*homo+1)
    "methane-svp": (34, 104, 4),
    "benzene-tzvp": (222, 546, 20),
    "dcv5t-tzvp": (1249, 3177, 143),
    "c60-tzvp": (1860, 4560, 179),
    "tiny": (48, 96, 7),
    "small": (160, 384, 23),
    "medium": (420, 1024, 55),
}


def spectrum(nbasis, homo, rng):
    n_occ = homo + 1
    occ = np.sort(rng.uniform(-1.2, -0.25, n_occ))
    u = rng.uniform(0.0, 1.0, nbasis - n_occ)
    virt = np.sort(0.02 + 3.0 * u * u)
    return np.concatenate([occ, virt])


SIGMA_X_TARGET = 0.4  # Hartree: typical magnitude of the exchange self-energy of a valence level


def band_profile(nbasis):
    idx = np.arange(nbasis)
    return np.exp(-np.abs(idx[:, None] - idx[None, :]) / 16.0)


def ao3c_sigma(nbasis, naux, homo):
    """Std-dev s of the Gaussian entries G (T = (G + G^T) * band) such that the MO-basis tensor after the
    V^-1/2 metric has var(M) = SIGMA_X_TARGET / (n_occ * naux), i.e. Sigma_x ~ -0.4 Ha for every level."""
    var_m = SIGMA_X_TARGET / ((homo + 1) * naux)
    mean_band2 = float(np.mean(band_profile(nbasis) ** 2))
    metric = 0.58  # mean of 1/lambda over the spectrum of A A^T / naux + I (Marchenko-Pastur, ratio 1)
    return np.sqrt(var_m / (2.0 * mean_band2 * metric))


def make_vxc(e, homo, rng):
    q = min(3 * homo + 1, len(e) - 1) + 1
    R = 0.005 * rng.standard_normal((q, q))
    return np.diag(np.full(q, -0.9 * SIGMA_X_TARGET)) + 0.5 * (R + R.T)


def make_small(nbasis, naux, homo, seed=20261017):
    """Everything on the host with NumPy (sizes up to a few hundred basis functions)."""
    rng = np.random.default_rng(seed)
    e = spectrum(nbasis, homo, rng)
    Q, _ = np.linalg.qr(rng.standard_normal((nbasis, nbasis)))
    G = rng.standard_normal((naux, nbasis, nbasis)) * (band_profile(nbasis) * ao3c_sigma(nbasis, naux, homo))[None]
    ao3c = G + G.transpose(0, 2, 1)
    A = rng.standard_normal((naux, naux))
    V = A @ A.T / naux + np.eye(naux)
    S = np.eye(naux)
    return {"mos": Q, "mo_energies": e, "ao3c": ao3c, "aux_overlap": S, "aux_coulomb": V,
            "vxc": make_vxc(e, homo, rng), "homo": homo}
