// Three-centre Coulomb integrals (P | mu nu) over contracted real-solid-harmonic Gaussian shells: the work of
// ComputeAO3cBlock (xtp/src/libxtp/libint2_calls.cc:544-593, libint2 Operator::coulomb, BraKet::xs_xx) and, with
// a unit partner, of AOCoulomb::Fill (libint2_calls.cc:224-271, BraKet::xs_xs), restated for one warp per
// (orbital shell pair, auxiliary shell).
//
// Scheme: McMurchie-Davidson.  Per primitive pair the 1-D Hermite coefficients E^{ab}_t of the three directions
// (built once per basis on the host, ao3c_tables.h: only primitive pairs above the screening threshold have a
// record, shell pairs without any are never launched), per primitive triple the Hermite Coulomb tensor R_{tuv} (Boys function from a grid + 8-term Taylor step, then
// the level recursion R^n -> R^{n-1}), the auxiliary side folded in first (G), then the pair side; cartesian ->
// pure transformation of the three indices at the end.  The 32 lanes split the entries of every stage; stages are
// separated by a warp barrier.  The code is __host__ __device__ and parameterised on the barrier so the CPU
// harness (tests/host_harness/ao3c_host.cc) runs the same source with std::barrier + one thread per lane
// (ThreadSanitizer then checks the barrier placement).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AO_HD __host__ __device__ __forceinline__
#else
#define AO_HD inline
#endif

namespace gwbse {
namespace ao {

struct BasisView {
  int nshell, nfunc;
  const int* l;
  const int* np;
  const int* prim0;
  const int* func0;
  const double* center;  // 3 per shell
  const double* exps;
  const double* coefs;
  const double* herm1;  // HERM1 per primitive
};

struct TableView {
  const double* boys;    // [points][orders]
  const uint32_t* tuv;   // [nherm]: t | u << 8 | v << 16
  const double* pure;    // concatenated (2l+1) x ncart(l) matrices
  int pure_off[8];
  int boys_orders, boys_taylor, herm1_stride, herm1_dim;
  double boys_dx, boys_xmax;
};

// where a block of integrals goes: element (aux function k, mu, nu) at
//   base[(k - func_begin) * stride_k + mu * stride_mu + nu * stride_nu]   for func_begin <= k < func_end
struct OutSpec {
  double* base;
  long long stride_k, stride_mu, stride_nu;
  int func_begin, func_end;
  int mirror;  // also write (k, nu, mu)
};

// One shell pair of a launch list.  b < 0: unit partner (exponent 0, coefficient 1, s type, on the centre of a).
// Its npp surviving primitive pairs have records of rec_doubles doubles each at pool[off ...]:
//   [0] p = a + b   [1..3] P = (a A + b B) / p   [4] c_a c_b   [5 ...] E[d][i][j][t], d = x, y, z
// (the Gaussian product factor exp(-mu |AB|^2) is inside E[d][0][0][0], direction by direction)
struct PairEntry {
  int a, b;
  int npp, rec_doubles;
  long long off;
};

AO_HD int nc_of(int l) { return (l + 1) * (l + 2) / 2; }
AO_HD int nh_of(int L) { return (L + 1) * (L + 2) * (L + 3) / 6; }
AO_HD void unpack_tuv(uint32_t w, int& t, int& u, int& v) {
  t = (int)(w & 0xffu);
  u = (int)((w >> 8) & 0xffu);
  v = (int)((w >> 16) & 0xffu);
}
AO_HD int hidx(int t, int u, int v) {
  const int N = t + u + v, w = u + v;
  return N * (N + 1) * (N + 2) / 6 + w * (w + 1) / 2 + v;
}

// doubles of scratch one warp needs for the class (la, lb, lc)
AO_HD int workspace_doubles(int la, int lb, int lc) {
  const int Lab = la + lb, L = Lab + lc;
  const int nca = nc_of(la), ncb = nc_of(lb), ncc = nc_of(lc);
  // E, R (two levels) and G; after the primitive loops the same region holds the block with its aux index transformed
  int erg = 3 * (la + 1) * (lb + 1) * (Lab + 1) + 2 * nh_of(L) + nh_of(Lab) * ncc;
  const int pc = (2 * lc + 1) * nca * ncb;
  if (pc > erg) erg = pc;
  return erg + nca * ncb * ncc + (L + 1);
}

// F_n(x) for one order: Taylor step off the tabulated grid, or erf + upward recursion beyond it
AO_HD double boys_one(const TableView& tb, int n, double x) {
  if (x < tb.boys_xmax) {
    const int j = (int)(x / tb.boys_dx + 0.5);
    const double d = (double)j * tb.boys_dx - x;  // -(x - x_j)
    const double* f = tb.boys + (long long)j * tb.boys_orders + n;
    double term = 1.0, s = f[0];
    for (int k = 1; k < tb.boys_taylor; ++k) {
      term *= d / (double)k;
      s += f[k] * term;
    }
    return s;
  }
  const double ex = exp(-x);
  double f = 0.5 * sqrt(3.14159265358979323846 / x) * erf(sqrt(x));
  const double inv2x = 0.5 / x;
  for (int k = 0; k < n; ++k) f = ((2 * k + 1) * f - ex) * inv2x;
  return f;
}

// pe.b < 0: unit partner -> two-centre integrals (sc | a)
// sc = -1:  no Coulomb operator at all -> overlap <a | b> (AOOverlap::Fill, libint2_calls.cc:163-165), written as
//           "aux function" 0
// sc = -2, -3, -4:  dipole <a | x | b>, <a | y | b>, <a | z | b> about the origin (AODipole::Fill, libint2
//           Operator::emultipole1), likewise
template <class Sync>
AO_HD void triple_block(const BasisView& dft, const BasisView& aux, const TableView& tb, const PairEntry& pe,
                        const double* pool, int sc, double* ws, int lane, int nl, Sync& sync, const OutSpec& out) {
  const int sa = pe.a, sb = pe.b;
  const bool unit_b = sb < 0, overlap = sc < 0;
  const int dipole_dir = -2 - sc;  // 0, 1, 2 for the dipole components, -1 for the plain overlap
  const int la = dft.l[sa], lb = unit_b ? 0 : dft.l[sb], lc = overlap ? 0 : aux.l[sc];
  const int Lab = la + lb, L = Lab + lc;
  const int nca = nc_of(la), ncb = nc_of(lb), ncc = nc_of(lc);
  const int npa = 2 * la + 1, npb = 2 * lb + 1, npc = 2 * lc + 1;
  const int nhab = nh_of(Lab), nhL = nh_of(L);
  const int T1 = Lab + 1, ej = (lb + 1) * T1;  // E[d][(i*(lb+1)+j)*T1 + t]
  const int esz = (la + 1) * ej;

  double* E = ws;
  double* R = E + 3 * esz;
  double* G = R + 2 * nhL;
  int erg = 3 * esz + 2 * nhL + nhab * ncc;
  if (npc * nca * ncb > erg) erg = npc * nca * ncb;
  double* acc = ws + erg;
  double* seed = acc + nca * ncb * ncc;

  const double Cx = overlap ? 0.0 : aux.center[3 * sc], Cy = overlap ? 0.0 : aux.center[3 * sc + 1],
               Cz = overlap ? 0.0 : aux.center[3 * sc + 2];
  const int pc0 = overlap ? 0 : aux.prim0[sc], nprc = overlap ? 0 : aux.np[sc];
  const uint32_t* cart_a = tb.tuv + nh_of(la - 1);  // the degree-l entries are the cartesian components
  const uint32_t* cart_b = tb.tuv + nh_of(lb - 1);
  const uint32_t* cart_c = tb.tuv + nh_of(lc - 1);
  // lane-strided walks over (h, c) and (ja, jb, c) item spaces without per-item divisions: start decomposition
  // and step decomposition are fixed per lane and class
  const int step_q = nl / ncc, step_r = nl % ncc;            // it += nl  ->  (it / ncc, it % ncc) += (q, r) + carry
  const int step_q2 = step_q / ncb, step_r2 = step_q % ncb;  // ab += q   ->  (ab / ncb, ab % ncb) += (q2, r2)
  const int c0 = lane % ncc, ab0 = lane / ncc, ja0 = ab0 / ncb, jb0 = ab0 % ncb;
  const double sgn_c = (lc & 1) ? -1.0 : 1.0;  // (-1)^(tau+nu+phi): the aux Hermite indices have the parity of lc
  const int H1 = tb.herm1_dim;

  for (int i = lane; i < nca * ncb * ncc; i += nl) acc[i] = 0.0;
  sync();

  for (int ipp = 0; ipp < pe.npp; ++ipp) {
    {
      const double* rec = pool + pe.off + (long long)ipp * pe.rec_doubles;
      const double p = rec[0], Px = rec[1], Py = rec[2], Pz = rec[3], cab = rec[4];
      // ---- E^{ab} of this primitive pair: global -> scratch -------------------------------------------------
      for (int i = lane; i < 3 * esz; i += nl) E[i] = rec[5 + i];
      sync();
      if (overlap) {
        // <a|b> += c_a c_b (pi/p)^(3/2) E^x_0 E^y_0 E^z_0;  dipole along d: E^d_0 -> E^d_1 + P_d E^d_0
        const double pref = cab * 5.568327996831708 / (p * sqrt(p));  // pi^(3/2)
        const double Pd = dipole_dir == 0 ? Px : (dipole_dir == 1 ? Py : Pz);
        for (int it = lane; it < nca * ncb; it += nl) {
          const int ja = it / ncb, jb = it % ncb;
          int ax, ay, az, bx, by, bz;
          unpack_tuv(cart_a[ja], ax, ay, az);
          unpack_tuv(cart_b[jb], bx, by, bz);
          const double* Ex = E + ax * ej + bx * T1;
          const double* Ey = E + esz + ay * ej + by * T1;
          const double* Ez = E + 2 * esz + az * ej + bz * T1;
          double fx = Ex[0], fy = Ey[0], fz = Ez[0];
          if (dipole_dir == 0) fx = (ax + bx >= 1 ? Ex[1] : 0.0) + Pd * Ex[0];
          if (dipole_dir == 1) fy = (ay + by >= 1 ? Ey[1] : 0.0) + Pd * Ey[0];
          if (dipole_dir == 2) fz = (az + bz >= 1 ? Ez[1] : 0.0) + Pd * Ez[0];
          acc[it] += pref * fx * fy * fz;
        }
        sync();
      }
      for (int ic = 0; ic < nprc; ++ic) {
        const double g = aux.exps[pc0 + ic];
        const double alpha = p * g / (p + g);
        const double X = Px - Cx, Y = Py - Cy, Z = Pz - Cz;
        const double pref = cab * aux.coefs[pc0 + ic] * 34.986836655249725 / (p * g * sqrt(p + g));  // 2 pi^(5/2)
        // ---- seeds (-2 alpha)^n F_n(alpha |PC|^2) ------------------------------------------------------------
        const double xarg = alpha * (X * X + Y * Y + Z * Z);
        for (int n = lane; n <= L; n += nl) {
          double s = boys_one(tb, n, xarg);
          for (int k = 0; k < n; ++k) s *= -2.0 * alpha;
          seed[n] = s;
        }
        sync();
        // ---- R^n_{tuv}, n = L .. 0, level n in buffer n & 1 -------------------------------------------------
        for (int n = L; n >= 0; --n) {
          double* cur = R + (n & 1) * nhL;
          const double* prev = R + ((n + 1) & 1) * nhL;
          const int cnt = nh_of(L - n);
          for (int h = lane; h < cnt; h += nl) {
            int t, u, v;
            unpack_tuv(tb.tuv[h], t, u, v);
            double val;
            if (h == 0) {
              val = seed[n];
            } else if (t > 0) {
              val = X * prev[hidx(t - 1, u, v)];
              if (t > 1) val += (t - 1) * prev[hidx(t - 2, u, v)];
            } else if (u > 0) {
              val = Y * prev[hidx(t, u - 1, v)];
              if (u > 1) val += (u - 1) * prev[hidx(t, u - 2, v)];
            } else {
              val = Z * prev[hidx(t, u, v - 1)];
              if (v > 1) val += (v - 1) * prev[hidx(t, u, v - 2)];
            }
            cur[h] = val;
          }
          sync();
        }
        // ---- G[h][c] = pref * sum_{tau nu phi} (-1)^(..) e_c R_{t+tau,u+nu,v+phi} ---------------------------
        const double* e1 = aux.herm1 + (long long)(pc0 + ic) * tb.herm1_stride;
        for (int it = lane, h = ab0, c = c0; it < nhab * ncc; it += nl) {
          int t, u, v, cx, cy, cz;
          unpack_tuv(tb.tuv[h], t, u, v);
          unpack_tuv(cart_c[c], cx, cy, cz);
          double s = 0.0;
          for (int tau = cx & 1; tau <= cx; tau += 2) {
            const double ex = e1[cx * H1 + tau];
            for (int nu = cy & 1; nu <= cy; nu += 2) {
              const double exy = ex * e1[cy * H1 + nu];
              for (int phi = cz & 1; phi <= cz; phi += 2)
                s += exy * e1[cz * H1 + phi] * R[hidx(t + tau, u + nu, v + phi)];
            }
          }
          G[it] = pref * sgn_c * s;
          h += step_q;
          c += step_r;
          if (c >= ncc) c -= ncc, ++h;
        }
        sync();
        // ---- acc[a][b][c] += sum_{tuv} Ex Ey Ez G[tuv][c] ---------------------------------------------------
        for (int it = lane, c = c0, ja = ja0, jb = jb0; it < nca * ncb * ncc; it += nl) {
          int ax, ay, az, bx, by, bz;
          unpack_tuv(cart_a[ja], ax, ay, az);
          unpack_tuv(cart_b[jb], bx, by, bz);
          const double* Ex = E + ax * ej + bx * T1;
          const double* Ey = E + esz + ay * ej + by * T1;
          const double* Ez = E + 2 * esz + az * ej + bz * T1;
          double s = 0.0;
          for (int t = 0; t <= ax + bx; ++t)
            for (int u = 0; u <= ay + by; ++u) {
              const double exy = Ex[t] * Ey[u];
              // hidx(t, u, v + 1) - hidx(t, u, v) = (N+1)(N+2)/2 + (u+v) + 2 with N = t+u+v: walk it, no division
              int N = t + u, w = u, h = hidx(t, u, 0);
              for (int v = 0; v <= az + bz; ++v) {
                s += exy * Ez[v] * G[h * ncc + c];
                h += (N + 1) * (N + 2) / 2 + w + 2;
                ++N;
                ++w;
              }
            }
          acc[it] += s;
          c += step_r;
          jb += step_r2;
          ja += step_q2;
          if (c >= ncc) c -= ncc, ++jb;
          if (jb >= ncb) jb -= ncb, ++ja;
        }
        sync();
      }
    }
  }

  // ---- cartesian -> pure, one index after the other.  E, R and G are free now: the aux index goes
  //      acc[(ja, jb), c] -> pc[(m * nca + ja) * ncb + jb] in that region, shell a pc -> half[(m * npa + ma) * ncb + jb]
  //      back in the accumulator region (npc * npa * ncb <= nca * ncb * ncc), shell b from half straight to memory
  const double* Tc = tb.pure + tb.pure_off[lc];
  const double* Ta = tb.pure + tb.pure_off[la];
  const double* Tb = tb.pure + tb.pure_off[lb];
  double* pc = ws;
  double* half = acc;
  for (int it = lane; it < npc * nca * ncb; it += nl) {
    const int ab = it % (nca * ncb), m = it / (nca * ncb);
    double s = 0.0;
    for (int c = 0; c < ncc; ++c) s += Tc[m * ncc + c] * acc[ab * ncc + c];
    pc[it] = s;
  }
  sync();
  const int fa = dft.func0[sa], fb = unit_b ? 0 : dft.func0[sb], fc = overlap ? 0 : aux.func0[sc];
  for (int it = lane; it < npc * npa * ncb; it += nl) {
    const int jb = it % ncb, mm = it / ncb, ma = mm % npa, m = mm / npa;
    double s = 0.0;
    for (int ja = 0; ja < nca; ++ja) s += Ta[ma * nca + ja] * pc[(m * nca + ja) * ncb + jb];
    half[it] = s;
  }
  sync();
  // shell b, and out (only the aux functions inside the requested range); ma fastest: consecutive lanes write
  // consecutive mu
  for (int it = lane; it < npc * npb * npa; it += nl) {
    const int ma = it % npa, mm = it / npa, mb = mm % npb, m = mm / npb;
    const int k = fc + m;
    if (k < out.func_begin || k >= out.func_end) continue;
    double s = 0.0;
    const double* h = half + (m * npa + ma) * ncb;
    for (int jb = 0; jb < ncb; ++jb) s += Tb[mb * ncb + jb] * h[jb];
    double* dst = out.base + (long long)(k - out.func_begin) * out.stride_k;
    const long long mu = fa + ma, nu = fb + mb;
    dst[mu * out.stride_mu + nu * out.stride_nu] = s;
    if (out.mirror && sa != sb) dst[nu * out.stride_mu + mu * out.stride_nu] = s;
  }
  sync();  // the scratch may be reused by the caller's next triple
}

// What one thread of the launch does: thread tid of CTA `block` (nwarps warps) belongs to lane group `sub` of its warp;
// group g = (block * nwarps + warp) * groups_per_warp + sub works on shell pair g / naux_shells and aux shell
// aux_shells[g % naux_shells] (consecutive groups share the pair: same records in L1/L2, neighbouring output rows)
// with its own slice of the CTA's scratch.  sync_of(sub, group_lanes) hands out the group's barrier.
template <class SyncFactory>
AO_HD void cta_thread(long long block, int nwarps, int tid, const BasisView& dft, const BasisView& aux, const TableView& tb,
                      const PairEntry* pairs, long long npairs, const double* pool, const int* aux_shells,
                      int naux_shells, const OutSpec& out, int ws_doubles, int group_lanes, double* scratch,
                      SyncFactory& sync_of) {
  const int warp = tid >> 5, lane = tid & 31;
  const int groups_per_warp = 32 / group_lanes, sub = lane / group_lanes, glane = lane % group_lanes;
  const long long g = (block * nwarps + warp) * groups_per_warp + sub;
  if (g >= npairs * naux_shells) return;  // whole groups leave; the barriers are group-local
  const PairEntry pe = pairs[g / naux_shells];
  auto sync = sync_of(sub, group_lanes);
  triple_block(dft, aux, tb, pe, pool, aux_shells[(int)(g % naux_shells)],
               scratch + ((long long)warp * groups_per_warp + sub) * ws_doubles, glane, group_lanes, sync, out);
}

}  // namespace ao
}  // namespace gwbse
