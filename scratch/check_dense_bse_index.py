"""CPU check of the index algebra of the materialised BSE blocks (csrc/capi_bse.cu: dense_build / dense_apply).

The generalised GEMM addressing of gemm_dmma.cuh and the pack kernel are emulated on flat NumPy buffers with exactly
the parameter values the C++ sets, for 1, 2 and 3 ranks (m-cyclic Mmn shards, gathered vv / cv blocks), and the sum
of the ranks' contributions is compared with the defining sums
  Hd [(v1,c1),(v2,c2)] = sum_chi M[v1][v2,chi] e[chi] M[c1][c2,chi]
  Hd2[(v1,c1),(v2,c2)] = sum_chi M[c1][v2,chi] e[chi] M[v1][c2,chi].
Run: python scratch/check_dense_bse_index.py"""
import numpy as np

BIG = 1 << 30


class Op:
    def __init__(self, buf, off=0, s_ri=1, s_ro=0, s_ki=1, s_ko=0, Lr=BIG):
        self.buf, self.off, self.s_ri, self.s_ro, self.s_ki, self.s_ko, self.Lr = buf, off, s_ri, s_ro, s_ki, s_ko, Lr

    def rows(self, n, Ko, Ki):
        r = np.arange(n)
        ko, ki = np.meshgrid(np.arange(Ko), np.arange(Ki), indexing="ij")
        idx = (self.off + (r // self.Lr)[:, None] * self.s_ro + (r % self.Lr)[:, None] * self.s_ri
               + (ko.ravel() * self.s_ko + ki.ravel() * self.s_ki)[None, :])
        return self.buf[idx]


def gemm(A, B, M, N, Ki, C, coff, sC_mi=1, sC_mo=0, sC_ni=0, sC_no=0, Lm=BIG, Ln=BIG, alpha=1.0, beta=0.0, Ko=1):
    a, b = A.rows(M, Ko, Ki), B.rows(N, Ko, Ki)
    prod = alpha * (a @ b.T)
    r, c = np.arange(M), np.arange(N)
    idx = coff + ((r // Lm) * sC_mo + (r % Lm) * sC_mi)[:, None] + ((c // Ln) * sC_no + (c % Ln) * sC_ni)[None, :]
    assert len(np.unique(idx)) == idx.size, "C elements written twice"
    assert idx.min() >= 0 and idx.max() < C.size, "C out of bounds"
    C[idx] = prod + (beta * C[idx] if beta != 0.0 else 0.0)


def pack(src, soff, s_pole, s_outer, L1, L2, scale, plane, npoles):
    out = np.full(plane * npoles, np.nan)
    for p in range(npoles):
        f = 1.0 if scale is None else scale[p]
        for a in range(L1):
            out[p * plane + a * L2:p * plane + (a + 1) * L2] = f * src[soff + p * s_pole + a * s_outer:
                                                                      soff + p * s_pole + a * s_outer + L2]
    return out


def round_up(v, a):
    return (v + a - 1) // a * a


def run(world, naux, mtotal, ntotal, voff, vt, ct, k, lchunk_force, seed=0):
    rng = np.random.default_rng(seed)
    coff = voff + vt
    assert coff + ct <= mtotal and coff + ct <= ntotal
    Mfull = rng.standard_normal((mtotal, ntotal, naux))
    eps = rng.uniform(0.3, 1.0, naux)
    B = vt * ct
    Xin = rng.standard_normal((B, k))
    # reference
    Mvv = Mfull[voff:voff + vt, voff:voff + vt]
    Mcc = Mfull[coff:coff + ct, coff:coff + ct]
    Mcv = Mfull[coff:coff + ct, voff:voff + vt]
    Mvc = Mfull[voff:voff + vt, coff:coff + ct]
    Hd = np.einsum("abx,x,cdx->acbd", Mvv, eps, Mcc).reshape(B, B)     # (v1,c1),(v2,c2)
    Hd2 = np.einsum("cbx,x,adx->acbd", Mcv, eps, Mvc).reshape(B, B)    # M[c1][v2] e M[v1][c2]
    ref = {0: Hd @ Xin, 1: Hd2 @ Xin}
    npad = round_up(ntotal, 16)
    mlmax = (mtotal + world - 1) // world
    ldx = mlmax * npad
    vtp = round_up(vt, 2)
    # replicated blocks as gather_slices lays them out: out[chi * ldo + s * rpad + row]
    gvv = np.zeros(naux * vt * vtp)
    gcv = np.zeros(naux * ct * vtp)
    for x in range(naux):
        for s in range(vt):
            gvv[x * vt * vtp + s * vtp:x * vt * vtp + s * vtp + vt] = Mfull[voff + s, voff:voff + vt, x]
        for s in range(ct):
            gcv[x * ct * vtp + s * vtp:x * ct * vtp + s * vtp + vt] = Mfull[coff + s, voff:voff + vt, x]
    for kind in (0, 1):
        Y = np.zeros(B * k)
        ldy = ldin = B
        for rank in range(world):
            # local shard X[chi * ldx + ml * npad + n], ml = m // world for m % world == rank
            X = np.zeros(ldx * naux)
            for m in range(rank, mtotal, world):
                for x in range(naux):
                    X[x * ldx + (m // world) * npad:x * ldx + (m // world) * npad + ntotal] = Mfull[m, :, x]

            def first_owned(s0):
                return s0 + ((rank - s0 % world) % world + world) % world

            def owned_count(s0, ns):
                f = first_owned(s0)
                return 0 if f >= s0 + ns else (s0 + ns - f + world - 1) // world
            nvloc, v_rel0 = owned_count(voff, vt), first_owned(voff) - voff
            lvfirst = (voff + v_rel0) // world if nvloc else 0
            ncloc, c_rel0 = owned_count(coff, ct), first_owned(coff) - coff
            lcfirst = (coff + c_rel0) // world if ncloc else 0
            if world == 1:
                vv = (X, voff * npad + voff, npad, ldx)
                cv = (X, coff * npad + voff, npad, ldx)
            else:
                vv = (gvv, 0, vtp, vt * vtp)
                cv = (gcv, 0, vtp, ct * vtp)
            # ---------------- dense_build
            ld = round_up(B, 2)
            if ld % 256 == 0:
                ld += 2
            ncols = vt * ncloc if kind == 0 else nvloc * ct
            if ncols == 0:
                continue
            H = np.full(ld * ncols, np.nan)
            nloc = ncloc if kind == 0 else nvloc
            small_rows = vt * vt if kind == 0 else ct * vt
            small_plane = round_up(small_rows, 2)
            lchunk = min(lchunk_force, nloc)
            big_plane = round_up(lchunk * ct, 2)
            if kind == 0:
                packA = pack(vv[0], vv[1], vv[3], vv[2], vt, vt, eps, small_plane, naux)
            else:
                packA = pack(cv[0], cv[1], cv[3], cv[2], ct, vt, eps, small_plane, naux)
            for a in range(0, nloc, lchunk):
                n1 = min(lchunk, nloc - a)
                lfirst = (lcfirst if kind == 0 else lvfirst) + a
                packB = pack(X, lfirst * npad + coff, ldx, npad, n1, ct, None, big_plane, naux)
                if kind == 0:
                    gemm(Op(packB, s_ri=1, s_ki=big_plane), Op(packA, s_ri=1, s_ki=small_plane), n1 * ct, vt * vt, naux,
                         H, a * ld, Lm=ct, sC_mo=ld, sC_mi=1, Ln=vt, sC_no=ncloc * ld, sC_ni=ct)
                else:
                    gemm(Op(packB, s_ri=1, s_ki=big_plane), Op(packA, s_ri=1, s_ki=small_plane), n1 * ct, ct * vt, naux,
                         H, a * ct * ld, Lm=ct, sC_mo=ct * ld, sC_mi=1, Ln=vt, sC_no=ld, sC_ni=ct)
            # every element of the B x ncols block written
            Hm = H.reshape(ncols, ld)[:, :B]
            assert not np.isnan(Hm).any(), "block not completely written"
            # ---------------- dense_apply (alpha = -1 as for cd = 1)
            Xbuf = Xin.ravel(order="F").copy()
            if kind == 0:
                gemm(Op(H, s_ri=ld, s_ki=1), Op(Xbuf, s_ri=ldin, s_ki=1), ncols, k, B, Y, c_rel0, Lm=ncloc, sC_mo=ct,
                     sC_mi=world, sC_ni=ldy, alpha=-1.0, beta=1.0)
            else:
                gemm(Op(H, s_ri=ld, s_ki=1), Op(Xbuf, s_ri=ldin, s_ki=1), ncols, k, B, Y, v_rel0 * ct, Lm=ct,
                     sC_mo=world * ct, sC_mi=1, sC_ni=ldy, alpha=-1.0, beta=1.0)
        got = Y.reshape(k, B).T
        err = np.abs(got + ref[kind]).max() / np.abs(ref[kind]).max()
        assert err < 1e-12, (world, kind, err)
    return True


if __name__ == "__main__":
    for world in (1, 2, 3):
        for (naux, mt, nt, voff, vt, ct, k, lch) in [(7, 12, 14, 0, 4, 5, 3, 100), (6, 13, 13, 1, 3, 7, 2, 2),
                                                     (5, 9, 11, 2, 3, 4, 4, 1), (4, 20, 21, 3, 5, 6, 1, 3),
                                                     (3, 33, 34, 1, 16, 16, 2, 5)]:
            run(world, naux, mt, nt, voff, vt, ct, k, lch, seed=world)
    print("dense BSE block index algebra: ok (worlds 1, 2, 3; Hd and Hd2; chunked builds)")


def run_window_rotation(world, naux, mtotal, ntotal, n_lo, n_hi, seed=0):
    """capi_mmn.cu rotate_rows / gwbse_mmn_mul_right_window_dev / mmn_complete_rotation: rows n in [n_lo, n_hi) of every
    local slice through the second buffer and back in place, then the rest; (chi, ml) is one row index of pitch npad."""
    rng = np.random.default_rng(seed)
    Mfull = rng.standard_normal((mtotal, ntotal, naux))
    R = rng.standard_normal((naux, naux))
    ref = Mfull @ R
    npad = round_up(ntotal, 16)
    mlmax = (mtotal + world - 1) // world
    ldx = mlmax * npad
    ldp = round_up(naux, 2)
    Rp = np.zeros(ldp * naux)
    for j in range(naux):
        Rp[j * ldp:j * ldp + naux] = R[:, j]          # column-major R at pitch ldp: R[k, j] at j * ldp + k
    for rank in range(world):
        X = np.zeros(ldx * naux)
        for m in range(rank, mtotal, world):
            for x in range(naux):
                X[x * ldx + (m // world) * npad:x * ldx + (m // world) * npad + ntotal] = Mfull[m, :, x]
        X2 = np.full(ldx * (naux + world), np.nan)

        def rotate_rows(n0, n1):
            nw = n1 - n0
            if nw <= 0:
                return
            # C[(ml, n), j] = sum_k X[k * ldx + ml * npad + n0 + n] * Rp[j * ldp + k]
            gemm(Op(X, off=n0, s_ri=1, s_ro=npad, s_ki=ldx, Lr=nw), Op(Rp, s_ri=ldp, s_ki=1), mlmax * nw, naux, naux,
                 X2, n0, Lm=nw, sC_mi=1, sC_mo=npad, sC_ni=ldx)
            # cudaMemcpy2D: width nw, height mlmax * naux, pitch npad on both sides
            for r in range(mlmax * naux):
                X[n0 + r * npad:n0 + r * npad + nw] = X2[n0 + r * npad:n0 + r * npad + nw]

        lo = n_lo & ~1
        rotate_rows(lo, n_hi)
        # window rows carry the rotation, the others are untouched
        for m in range(rank, mtotal, world):
            got = np.array([X[x * ldx + (m // world) * npad:x * ldx + (m // world) * npad + ntotal] for x in range(naux)]).T
            assert np.abs(got[lo:n_hi] - ref[m, lo:n_hi]).max() < 1e-12
            assert np.array_equal(got[:lo], Mfull[m, :lo]) and np.array_equal(got[n_hi:], Mfull[m, n_hi:])
        rotate_rows(0, lo)
        rotate_rows(n_hi, ntotal)
        for m in range(rank, mtotal, world):
            got = np.array([X[x * ldx + (m // world) * npad:x * ldx + (m // world) * npad + ntotal] for x in range(naux)]).T
            assert np.abs(got - ref[m]).max() < 1e-12
    return True


if __name__ == "__main__":
    for world in (1, 2, 3):
        for (naux, mt, nt, lo, hi) in [(5, 7, 19, 4, 11), (6, 8, 17, 3, 17), (4, 5, 16, 0, 9)]:
            run_window_rotation(world, naux, mt, nt, lo, hi, seed=world)
    print("windowed MultiplyRight index algebra: ok")
