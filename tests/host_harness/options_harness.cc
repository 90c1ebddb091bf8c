// C entry points over the Options class of votca_b200/host/gwbse.h (device-free part) for tests/test_options_cpu.py.
#include <cstring>
#include <string>

#include "../../votca_b200/host/gwbse.h"

using namespace votca::xtp;

extern "C" {
// RPA::UpdateRPAInputEnergies (rpa.cc:32-70) is pure host arithmetic: run it on an RPA object whose device handle
// is never touched (the Device behind the reference is deliberately not constructed - there is no GPU here).
int rpa_update_energies(long homo, long rpamin, long rpamax, const double* dft, long ndft, const double* gw, long ngw,
                        long qpmin, double* out) {
  try {
    alignas(Device) static unsigned char no_device[sizeof(Device)];
    const Device& dev = *reinterpret_cast<const Device*>(no_device);
    Logger log;
    TCMatrix_gwbse Mmn(dev);
    RPA rpa(log, Mmn);
    rpa.configure(homo, rpamin, rpamax);
    rpa.UpdateRPAInputEnergies(VectorXd(dft, ndft), VectorXd(gw, ngw), qpmin);
    const VectorXd& e = rpa.getRPAInputEnergies();
    for (long i = 0; i < e.size(); ++i) out[i] = e(i);
    return (int)e.size();
  } catch (...) {
    return -1;
  }
}

// BuildFullBSEXRankedInitialGuess (bse_initialization.h:47-93): out is 2n x nguess column-major; returns nguess
long bse_ranked_guess(const double* adiag, const double* bdiag, long n, long nroots, double* out, long cap) {
  try {
    const MatrixXd g = BuildFullBSEXRankedInitialGuess(VectorXd(adiag, n), VectorXd(bdiag, n), nroots);
    if (g.rows() * g.cols() > cap) return -2;
    for (long j = 0; j < g.cols(); ++j)
      for (long i = 0; i < g.rows(); ++i) out[j * g.rows() + i] = g(i, j);
    return g.cols();
  } catch (...) {
    return -1;
  }
}

// GWBSE::Initialize (gwbse.cc:60-233): level ranges and derived sizes from the options; ranges[6] = rpamin, rpamax,
// qpmin, qpmax, bse_vmin, bse_cmax; returns 0, or 1 with the exception text in err
int gwbse_initialize_ranges(void* options, long homo, long nlevels, long* ranges, long* bse_nmax, char* err, int cap) {
  try {
    alignas(Device) static unsigned char no_device[sizeof(Device)];
    const Device& dev = *reinterpret_cast<const Device*>(no_device);
    Logger log;
    GWBSE g(dev, log);
    MatrixXd mos(nlevels, nlevels);
    VectorXd e(nlevels);
    GWBSE::Inputs in;
    in.homo = homo;
    in.mos = &mos;
    in.mo_energies = &e;
    g.Initialize(*static_cast<Options*>(options), in);
    const GW::options& gw = g.gw_options();
    const BSE::options& bse = g.bse_options();
    const long r[6] = {gw.rpamin, gw.rpamax, gw.qpmin, gw.qpmax, bse.vmin, bse.cmax};
    for (int i = 0; i < 6; ++i) ranges[i] = r[i];
    *bse_nmax = bse.nmax;
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), cap - 1);
    err[cap - 1] = 0;
    return 1;
  }
}

// the same with the nuclear charges of the atoms given (ignore_corelevels, gwbse.cc:46-58, 128-152)
int gwbse_initialize_ranges_z(void* options, long homo, long nlevels, const double* z, long nz, long* ranges, char* err,
                              int cap) {
  try {
    alignas(Device) static unsigned char no_device[sizeof(Device)];
    const Device& dev = *reinterpret_cast<const Device*>(no_device);
    Logger log;
    GWBSE g(dev, log);
    MatrixXd mos(nlevels, nlevels);
    VectorXd e(nlevels), zz(nz);
    for (long i = 0; i < nz; ++i) zz(i) = z[i];
    GWBSE::Inputs in;
    in.homo = homo;
    in.mos = &mos;
    in.mo_energies = &e;
    if (nz > 0) in.nuclear_charges = &zz;
    g.Initialize(*static_cast<Options*>(options), in);
    const GW::options& gw = g.gw_options();
    const BSE::options& bse = g.bse_options();
    const long r[6] = {gw.rpamin, gw.rpamax, gw.qpmin, gw.qpmax, bse.vmin, bse.cmax};
    for (int i = 0; i < 6; ++i) ranges[i] = r[i];
    return 0;
  } catch (const std::exception& e) {
    std::strncpy(err, e.what(), cap - 1);
    err[cap - 1] = 0;
    return 1;
  }
}

// GWBSE::WriteToCpt on made-up results of the given sizes (values i + 0.5): checks names / types / shapes of the file
int gwbse_write_results(const char* path, long nlevels, long homo, long q, long bse_size, long nstates, int tda) {
  try {
    alignas(Device) static unsigned char no_device[sizeof(Device)];
    const Device& dev = *reinterpret_cast<const Device*>(no_device);
    Logger log;
    GWBSE g(dev, log);
    MatrixXd mos(nlevels, nlevels, 0.25);
    VectorXd e(nlevels, -0.5);
    GWBSE::Inputs in;
    in.homo = homo;
    in.mos = &mos;
    in.mo_energies = &e;
    in.ScaHFX = 0.25;
    Options opt;
    opt.set("bse.useTDA", tda ? "true" : "false");
    g.Initialize(opt, in);
    GWBSE::Results r;
    r.rpamin = 0, r.rpamax = nlevels - 1, r.qpmin = 0, r.qpmax = q - 1, r.bse_vmin = 0, r.bse_cmax = q - 1;
    r.RPA_inputenergies = VectorXd(nlevels, 1.5);
    r.QPpert_energies = VectorXd(q, 2.5);
    r.QPdiag_eigenvalues = VectorXd(q, 3.5);
    r.QPdiag_eigenvectors = MatrixXd(q, q, 4.5);
    r.BSE_singlet.eigenvalues = VectorXd(nstates, 5.5);
    r.BSE_singlet.eigenvectors = MatrixXd(bse_size, nstates, 6.5);
    if (!tda) r.BSE_singlet.eigenvectors2 = MatrixXd(bse_size, nstates, 7.5);
    r.BSE_singlet.success = true;
    for (long s2 = 0; s2 < nstates; ++s2) r.transition_dipoles.emplace_back(3, 8.5 + (double)s2);
    r.BSE_singlet_dynamic = VectorXd(nstates, 9.5);
    g.WriteToCpt(r, path);
    return 0;
  } catch (...) {
    return 1;
  }
}

void* opt_new() { return new Options(); }
void opt_free(void* o) { delete static_cast<Options*>(o); }
int opt_load_xml(void* o, const char* path) {
  try {
    static_cast<Options*>(o)->LoadFromXML(path);
    return 0;
  } catch (...) {
    return 1;
  }
}
int opt_set(void* o, const char* key, const char* value) {
  try {
    static_cast<Options*>(o)->set(key, value);
    return 0;
  } catch (...) {
    return 1;
  }
}
// 0: ok, 1: key unknown / empty
int opt_get(void* o, const char* key, char* out, int cap) {
  try {
    Options* opt = static_cast<Options*>(o);
    if (!opt->exists(key)) return 1;
    const std::string v = opt->str(key);
    std::strncpy(out, v.c_str(), cap - 1);
    out[cap - 1] = 0;
    return 0;
  } catch (...) {
    return 1;
  }
}
}
