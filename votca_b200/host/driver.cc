// C facade of the host layer (include/gwbse_host.h).
#include "../../include/gwbse_host.h"

#include <cstring>
#include <map>
#include <memory>

#include "aobasis.h"
#include "bsecoupling.h"
#include "gwbse.h"

using namespace votca;
using namespace votca::xtp;

namespace {

struct ArraySource : AOIntegralSource {
  Index N = 0, naux = 0;
  const double* ao3c = nullptr;
  const double* ao3c_dev = nullptr;
  gwbse_ao3c_fn fn = nullptr;
  void* user = nullptr;
  Index first_aux = 0, held = -1;  // the arrays hold aux functions [first_aux, first_aux + held); -1: all
  const MatrixXd* S = nullptr;
  const MatrixXd* V = nullptr;
  size_t local(Index aux_offset, Index aux_count) const {
    const Index n = held < 0 ? naux : held;
    if (aux_offset < first_aux || aux_offset + aux_count > first_aux + n)
      throw std::runtime_error("AO three-centre integrals for aux functions " + std::to_string(aux_offset) + ".." +
                               std::to_string(aux_offset + aux_count - 1) + " were not supplied to this rank");
    return static_cast<size_t>(aux_offset - first_aux) * N * N;
  }
  Index AuxSize() const override { return naux; }
  Index BasisSize() const override { return N; }
  void ComputeAO3cBlock(Index aux_offset, Index aux_count, double* out) const override {
    if (fn) {
      fn(user, aux_offset, aux_count, out);
    } else {
      if (!ao3c) throw std::runtime_error("no AO three-centre integrals supplied (ao3c)");
      std::memcpy(out, ao3c + local(aux_offset, aux_count), sizeof(double) * aux_count * N * N);
    }
  }
  const double* HostBlock(Index aux_offset, Index aux_count) const override {
    return (!fn && ao3c) ? ao3c + local(aux_offset, aux_count) : nullptr;
  }
  const double* DeviceBlock(Index aux_offset, Index aux_count) const override {
    return ao3c_dev ? ao3c_dev + local(aux_offset, aux_count) : nullptr;
  }
  const MatrixXd& AuxOverlap() const override { return *S; }
  const MatrixXd& AuxCoulomb() const override { return *V; }
};

}  // namespace

struct gwbse_job {
  std::unique_ptr<Device> dev;
  Logger log;
  Options options;
  std::map<std::string, MatrixXd> in;
  std::map<std::string, double> scalars;
  std::map<std::string, MatrixXd> out;
  std::map<std::string, double> out_scalars;
  ArraySource ints;
  // AO integrals produced on the device from the basis sets (gwbse_job_set_basis) instead of "ao3c"
  std::unique_ptr<AOBasisData> basis_data[2];  // 0 = dft, 1 = aux
  // device copies of the basis tables (shell data + primitive-pair records), kept across runs until the basis is
  // set again: they are inputs resident in HBM, not results
  std::unique_ptr<DeviceAOBasis> dev_basis[2];
  std::string orb_path;  // gwbse_job_set_orb_output: results are written there at the end of gwbse_job_run
  std::string summary_path;  // gwbse_job_set_summary_output: <job>_summary.xml of the dftgwbse tool
  std::map<std::string, std::string> coupling_options;  // "bsecoupling.*" keys (bsecoupling.xml)
  std::string coupling_xml;                             // BSECoupling::Addoutput of the last gwbse_job_run_coupling
  std::string err;
  mutable std::string logcache;
};

static thread_local std::string g_job_create_error;

#define JOB_BEGIN(job) \
  if (!(job)) return 1; \
  try {
#define JOB_END(job)                   \
  return 0;                            \
  }                                    \
  catch (const std::exception& e) {    \
    (job)->err = e.what();             \
    return 1;                          \
  }

static MatrixXd vec2mat(const VectorXd& v) { return MatrixXd(v.data(), v.size(), 1, std::max<Index>(v.size(), 1)); }
static VectorXd mat2vec(const MatrixXd& m) { return VectorXd(m.data(), m.size()); }

extern "C" {

int gwbse_job_create(int device, gwbse_job** out) {
  if (!out) return 1;
  *out = nullptr;
  try {
    auto job = std::make_unique<gwbse_job>();
    job->dev = std::make_unique<Device>(device);
    *out = job.release();
    return 0;
  } catch (const std::exception& e) {
    g_job_create_error = e.what();
    return 1;
  }
}

void gwbse_job_destroy(gwbse_job* job) { delete job; }
const char* gwbse_job_error(const gwbse_job* job) { return job ? job->err.c_str() : "null job"; }
const char* gwbse_job_create_error(void) { return g_job_create_error.c_str(); }
const char* gwbse_job_log(const gwbse_job* job) {
  if (!job) return "";
  job->logcache = job->log.str();
  return job->logcache.c_str();
}

int gwbse_job_comm_init(gwbse_job* job, int rank, int world, const unsigned char* id128) {
  JOB_BEGIN(job)
  job->dev->check(gwbse_comm_init(job->dev->ctx(), rank, world, id128));
  JOB_END(job)
}

int gwbse_job_set_option(gwbse_job* job, const char* key, const char* value) {
  JOB_BEGIN(job)
  const std::string k(key ? key : "");
  const size_t at = k.find("bsecoupling.");
  if (at != std::string::npos) {
    static const char* known[] = {"spin", "use_perturbation", "output_tb", "moleculeA.states", "moleculeA.occLevels",
                                  "moleculeA.unoccLevels", "moleculeB.states", "moleculeB.occLevels",
                                  "moleculeB.unoccLevels"};
    const std::string sub = k.substr(at + 12);
    bool ok = false;
    for (const char* n : known) ok = ok || sub == n;
    if (!ok) throw std::runtime_error("unknown option '" + k + "' (not a key of bsecoupling.xml)");
    job->coupling_options[sub] = value ? value : "";
  } else {
    job->options.set(key, value);
  }
  JOB_END(job)
}

int gwbse_job_load_options_xml(gwbse_job* job, const char* path) {
  JOB_BEGIN(job)
  job->options.LoadFromXML(path);
  JOB_END(job)
}

int gwbse_job_set_scalar(gwbse_job* job, const char* name, double value) {
  JOB_BEGIN(job)
  job->scalars[name] = value;
  JOB_END(job)
}

int gwbse_job_set_array(gwbse_job* job, const char* name, const double* data, long rows, long cols) {
  JOB_BEGIN(job)
  const std::string n(name);
  if (n == "ao3c") {
    job->ints.ao3c = data;
    job->ints.ao3c_dev = nullptr;
    job->ints.fn = nullptr;
    job->ints.first_aux = 0;
    job->ints.held = -1;
    job->ints.naux = cols;
    job->ints.N = static_cast<Index>(std::llround(std::sqrt(static_cast<double>(rows))));
    if (job->ints.N * job->ints.N != rows) throw std::runtime_error("ao3c: rows must be N*N");
  } else {
    job->in[n] = MatrixXd(data, rows, cols, std::max<long>(rows, 1));
  }
  JOB_END(job)
}

int gwbse_job_set_ao3c_callback(gwbse_job* job, long nbasis, long naux, gwbse_ao3c_fn fn, void* user) {
  JOB_BEGIN(job)
  job->ints.fn = fn;
  job->ints.user = user;
  job->ints.ao3c = nullptr;
  job->ints.ao3c_dev = nullptr;  // an earlier set_ao3c_dev / set_ao3c_partial must not shadow the callback
  job->ints.first_aux = 0;
  job->ints.held = -1;
  job->ints.N = nbasis;
  job->ints.naux = naux;
  JOB_END(job)
}

int gwbse_job_set_ao3c_dev(gwbse_job* job, long nbasis, long naux, const double* ao3c_dev) {
  JOB_BEGIN(job)
  job->ints.ao3c_dev = ao3c_dev;
  job->ints.ao3c = nullptr;
  job->ints.fn = nullptr;
  job->ints.first_aux = 0;
  job->ints.held = -1;
  job->ints.N = nbasis;
  job->ints.naux = naux;
  JOB_END(job)
}

int gwbse_job_set_ao3c_partial(gwbse_job* job, long nbasis, long naux, long first_aux, long count, const double* data,
                               int on_device) {
  JOB_BEGIN(job)
  if (first_aux < 0 || count < 0 || first_aux + count > naux) throw std::runtime_error("ao3c: invalid aux range");
  job->ints.ao3c = on_device ? nullptr : data;
  job->ints.ao3c_dev = on_device ? data : nullptr;
  job->ints.fn = nullptr;
  job->ints.first_aux = first_aux;
  job->ints.held = count;
  job->ints.N = nbasis;
  job->ints.naux = naux;
  JOB_END(job)
}

int gwbse_job_set_basis(gwbse_job* job, const char* which, int nshell, const int* l, const int* nprim,
                        const double* centers, const double* exps, const double* coefs) {
  JOB_BEGIN(job)
  const std::string w(which ? which : "");
  if (w != "dft" && w != "aux") throw std::runtime_error("basis must be 'dft' or 'aux'");
  if (nshell < 1 || !l || !nprim || !centers || !exps || !coefs) throw std::runtime_error("invalid basis description");
  auto d = std::make_unique<AOBasisData>();
  d->l.assign(l, l + nshell);
  d->nprim.assign(nprim, nprim + nshell);
  d->centers.assign(centers, centers + 3 * static_cast<size_t>(nshell));
  size_t np = 0;
  for (int s = 0; s < nshell; ++s) np += static_cast<size_t>(std::max(nprim[s], 0));
  d->exps.assign(exps, exps + np);
  d->coefs.assign(coefs, coefs + np);
  job->basis_data[w == "aux" ? 1 : 0] = std::move(d);
  job->dev_basis[w == "aux" ? 1 : 0].reset();
  JOB_END(job)
}

int gwbse_job_set_orb_output(gwbse_job* job, const char* path) {
  JOB_BEGIN(job)
  job->orb_path = path ? path : "";
  JOB_END(job)
}

int gwbse_job_set_summary_output(gwbse_job* job, const char* path) {
  JOB_BEGIN(job)
  job->summary_path = path ? path : "";
  JOB_END(job)
}

void* gwbse_job_ctx(gwbse_job* job) { return job ? job->dev->ctx() : nullptr; }

int gwbse_job_run(gwbse_job* job) {
  JOB_BEGIN(job)
  auto need = [&](const char* n) -> const MatrixXd& {
    auto it = job->in.find(n);
    if (it == job->in.end()) throw std::runtime_error(std::string("input array '") + n + "' not set");
    return it->second;
  };
  GWBSE::Inputs in;
  if (!job->scalars.count("homo")) throw std::runtime_error("input scalar 'homo' not set");
  in.homo = static_cast<Index>(job->scalars["homo"]);
  in.ScaHFX = job->scalars.count("ScaHFX") ? job->scalars["ScaHFX"] : 0.0;
  const MatrixXd& mos = need("mos");
  const VectorXd mo_e = mat2vec(need("mo_energies"));
  in.mos = &mos;
  in.mo_energies = &mo_e;
  if (job->in.count("vxc")) in.vxc = &job->in["vxc"];
  // integral producer: the device (both basis sets given, no ao3c array / callback), else the supplied arrays
  DeviceAOBasis *dft_basis = nullptr, *aux_basis = nullptr;
  std::unique_ptr<DeviceAOIntegrals> device_ints;
  const bool have_arrays = job->ints.ao3c || job->ints.ao3c_dev || job->ints.fn;
  if (!have_arrays && job->basis_data[0] && job->basis_data[1]) {
    for (int b = 0; b < 2; ++b)
      if (!job->dev_basis[b]) job->dev_basis[b] = std::make_unique<DeviceAOBasis>(*job->dev, *job->basis_data[b]);
    dft_basis = job->dev_basis[0].get();
    aux_basis = job->dev_basis[1].get();
    const MatrixXd* S = job->in.count("aux_overlap") ? &job->in["aux_overlap"] : nullptr;
    if (S && aux_basis->AOBasisSize() != S->rows()) throw std::runtime_error("aux overlap does not match the aux basis");
    if (dft_basis->AOBasisSize() != mos.rows()) throw std::runtime_error("MO coefficients do not match the dft basis");
    device_ints = std::make_unique<DeviceAOIntegrals>(*job->dev, *aux_basis, *dft_basis, S,
                                                      job->in.count("aux_coulomb") ? &job->in["aux_coulomb"] : nullptr);
    in.integrals = device_ints.get();
  } else {
    job->ints.S = &need("aux_overlap");
    job->ints.V = &need("aux_coulomb");
    if (job->ints.naux != job->ints.S->rows()) throw std::runtime_error("aux matrices do not match ao3c");
    in.integrals = &job->ints;
  }
  std::vector<MatrixXd> dip, ao_dip;
  if (job->in.count("ao_dipole_x") && job->in.count("ao_dipole_y") && job->in.count("ao_dipole_z")) {
    ao_dip = {job->in["ao_dipole_x"], job->in["ao_dipole_y"], job->in["ao_dipole_z"]};
    in.ao_dipoles = &ao_dip;
  } else if (dft_basis && !job->in.count("dipole_x")) {  // AO dipoles from the device, interlevel dipoles formed by GWBSE
    ao_dip = dft_basis->Dipoles();
    in.ao_dipoles = &ao_dip;
  }
  if (job->in.count("dipole_x") && job->in.count("dipole_y") && job->in.count("dipole_z")) {
    dip = {job->in["dipole_x"], job->in["dipole_y"], job->in["dipole_z"]};
    in.interlevel_dipoles = &dip;
  }
  // bse.fragments: overlap / atom map from the arrays given, else from the dft basis
  MatrixXd ao_overlap;
  std::vector<Index> basis_atom;
  VectorXd nuclear_charges;
  if (job->in.count("nuclear_charges")) {
    nuclear_charges = mat2vec(job->in["nuclear_charges"]);
    in.nuclear_charges = &nuclear_charges;
    if (job->in.count("ao_overlap")) {
      in.ao_overlap = &job->in["ao_overlap"];
    } else if (job->basis_data[0]) {
      if (!job->dev_basis[0]) job->dev_basis[0] = std::make_unique<DeviceAOBasis>(*job->dev, *job->basis_data[0]);
      ao_overlap = job->dev_basis[0]->Overlap();
      in.ao_overlap = &ao_overlap;
    }
    if (job->in.count("basis_atom_index")) {
      const VectorXd v = mat2vec(job->in["basis_atom_index"]);
      for (Index i = 0; i < v.size(); ++i) basis_atom.push_back(static_cast<Index>(std::llround(v(i))));
      in.basis_atom = &basis_atom;
    } else if (job->basis_data[0]) {
      basis_atom = job->basis_data[0]->AtomOfFunction();
      in.basis_atom = &basis_atom;
    }
  }
  VectorXd rpa_in;
  if (job->in.count("Hqp") && job->in.count("RPA_inputenergies")) {
    in.Hqp = &job->in["Hqp"];
    rpa_in = mat2vec(job->in["RPA_inputenergies"]);
    in.rpa_input_energies = &rpa_in;
  }
  GWBSE gwbse(*job->dev, job->log);
  gwbse.Initialize(job->options, in);
  GWBSE::Results r = gwbse.Evaluate();
  if (!job->orb_path.empty() && job->dev->rank() == 0) gwbse.WriteToCpt(r, job->orb_path);
  if (!job->summary_path.empty() && job->dev->rank() == 0)
    gwbse.WriteSummaryXML(r, job->summary_path,
                          job->scalars.count("dft_total_energy") ? job->scalars["dft_total_energy"] : 0.0);
  auto& o = job->out;
  o.clear();
  o["RPA_inputenergies"] = vec2mat(r.RPA_inputenergies);
  o["QPpert_energies"] = vec2mat(r.QPpert_energies);
  o["QPdiag_eigenvalues"] = vec2mat(r.QPdiag_eigenvalues);
  o["QPdiag_eigenvectors"] = r.QPdiag_eigenvectors;
  o["Hqp"] = r.Hqp;
  o["Sigma_x"] = r.Sigma_x;
  o["Sigma_c"] = r.Sigma_c;
  o["BSE_singlet_eigenvalues"] = vec2mat(r.BSE_singlet.eigenvalues);
  o["BSE_singlet_eigenvectors"] = r.BSE_singlet.eigenvectors;
  o["BSE_singlet_eigenvectors2"] = r.BSE_singlet.eigenvectors2;
  o["BSE_triplet_eigenvalues"] = vec2mat(r.BSE_triplet.eigenvalues);
  o["BSE_triplet_eigenvectors"] = r.BSE_triplet.eigenvectors;
  o["BSE_triplet_eigenvectors2"] = r.BSE_triplet.eigenvectors2;
  o["BSE_singlet_dynamic"] = vec2mat(r.BSE_singlet_dynamic);
  o["BSE_triplet_dynamic"] = vec2mat(r.BSE_triplet_dynamic);
  o["oscillator_strengths"] = vec2mat(r.oscillator_strengths);
  MatrixXd td(3, static_cast<Index>(r.transition_dipoles.size()));
  for (size_t s = 0; s < r.transition_dipoles.size(); ++s)
    for (Index i = 0; i < 3; ++i) td(i, s) = r.transition_dipoles[s](i);
  o["transition_dipoles"] = td;
  o["singlet_qp_contrib"] = vec2mat(r.singlet_analysis.qp_contrib);
  o["singlet_direct_contrib"] = vec2mat(r.singlet_analysis.direct_contrib);
  o["singlet_exchange_contrib"] = vec2mat(r.singlet_analysis.exchange_contrib);
  o["triplet_qp_contrib"] = vec2mat(r.triplet_analysis.qp_contrib);
  o["triplet_direct_contrib"] = vec2mat(r.triplet_analysis.direct_contrib);
  o["fragment_gs"] = vec2mat(r.fragment_gs);
  o["BSE_singlet_fragment_hole"] = r.singlet_fragment_hole;
  o["BSE_singlet_fragment_electron"] = r.singlet_fragment_electron;
  o["BSE_triplet_fragment_hole"] = r.triplet_fragment_hole;
  o["BSE_triplet_fragment_electron"] = r.triplet_fragment_electron;
  auto& s = job->out_scalars;
  s["rpamin"] = r.rpamin;
  s["rpamax"] = r.rpamax;
  s["qpmin"] = r.qpmin;
  s["qpmax"] = r.qpmax;
  s["bse_vmin"] = r.bse_vmin;
  s["bse_cmax"] = r.bse_cmax;
  s["removed_functions"] = r.removed_functions;
  s["gw_iterations"] = r.gw_iterations;
  s["qsgw_iterations"] = r.qsgw_iterations;
  s["is_qsgw"] = r.is_qsgw ? 1.0 : 0.0;
  s["singlet_davidson_iterations"] = r.singlet_davidson_iterations;
  s["triplet_davidson_iterations"] = r.triplet_davidson_iterations;
  s["singlet_converged"] = r.BSE_singlet.success;
  s["triplet_converged"] = r.BSE_triplet.success;
  s["sigma_batches"] = static_cast<double>(r.sigma_batches);
  s["sigma_evaluations"] = static_cast<double>(r.sigma_evaluations);
  s["time_fill"] = r.time_fill;
  s["time_gw"] = r.time_gw;
  s["time_bse"] = r.time_bse;
  JOB_END(job)
}

// BSECoupling through the job facade: the dimer uses the job's own inputs (mos, Hqp, RPA_inputenergies, integrals or
// basis sets, scalars homo / rpamin / rpamax / qpmin / qpmax / bse_vmin / bse_cmax / use_Hqp_offdiag), the monomers
// the arrays "A.mos", "A.BSE_singlet_eigenvectors", ... and the scalars "A.bse_vmin" ... "A.bse_cmax" (same for B).
int gwbse_job_run_coupling(gwbse_job* job) {
  JOB_BEGIN(job)
  auto need = [&](const std::string& n) -> const MatrixXd& {
    auto it = job->in.find(n);
    if (it == job->in.end()) throw std::runtime_error("input array '" + n + "' not set");
    return it->second;
  };
  auto opt = [&](const std::string& n) -> const MatrixXd* {
    auto it = job->in.find(n);
    return it == job->in.end() ? nullptr : &it->second;
  };
  auto scalar = [&](const std::string& n) -> Index {
    auto it = job->scalars.find(n);
    if (it == job->scalars.end()) throw std::runtime_error("input scalar '" + n + "' not set");
    return static_cast<Index>(std::llround(it->second));
  };
  BSECoupling::options co;
  auto copt = [&](const char* k, const std::string& dflt) {
    auto it = job->coupling_options.find(k);
    return it == job->coupling_options.end() ? dflt : it->second;
  };
  auto truth = [](const std::string& v) { return v == "true" || v == "1"; };
  co.spin = copt("spin", co.spin);
  co.use_perturbation = truth(copt("use_perturbation", "true"));
  co.output_tb = truth(copt("output_tb", "false"));
  co.statesA = std::stol(copt("moleculeA.states", "5"));
  co.occLevelsA = std::stol(copt("moleculeA.occLevels", "5"));
  co.unoccLevelsA = std::stol(copt("moleculeA.unoccLevels", "5"));
  co.statesB = std::stol(copt("moleculeB.states", "5"));
  co.occLevelsB = std::stol(copt("moleculeB.occLevels", "5"));
  co.unoccLevelsB = std::stol(copt("moleculeB.unoccLevels", "5"));

  std::vector<VectorXd> keep;  // energies as vectors, alive until the end of the run
  keep.reserve(8);
  auto fragment = [&](const std::string& tag) {
    CouplingOrbitals f;
    f.mos = &need(tag + ".mos");
    f.bse_vmin = scalar(tag + ".bse_vmin");
    f.bse_vmax = scalar(tag + ".bse_vmax");
    f.bse_cmin = scalar(tag + ".bse_cmin");
    f.bse_cmax = scalar(tag + ".bse_cmax");
    f.singlets = opt(tag + ".BSE_singlet_eigenvectors");
    f.triplets = opt(tag + ".BSE_triplet_eigenvectors");
    if (const MatrixXd* e = opt(tag + ".BSE_singlet_eigenvalues")) {
      keep.push_back(mat2vec(*e));
      f.singlet_energies = &keep.back();
    }
    if (const MatrixXd* e = opt(tag + ".BSE_triplet_eigenvalues")) {
      keep.push_back(mat2vec(*e));
      f.triplet_energies = &keep.back();
    }
    return f;
  };
  const CouplingOrbitals A = fragment("A"), B = fragment("B");
  CouplingOrbitals AB;
  AB.mos = &need("mos");
  AB.homo = scalar("homo");
  AB.rpamin = scalar("rpamin");
  AB.rpamax = scalar("rpamax");
  AB.qpmin = scalar("qpmin");
  AB.qpmax = scalar("qpmax");
  AB.bse_vmin = scalar("bse_vmin");
  AB.bse_cmax = scalar("bse_cmax");
  AB.bse_vmax = AB.homo;
  AB.bse_cmin = AB.homo + 1;
  AB.use_Hqp_offdiag = !job->scalars.count("use_Hqp_offdiag") || job->scalars["use_Hqp_offdiag"] != 0.0;
  AB.Hqp = &need("Hqp");
  keep.push_back(mat2vec(need("RPA_inputenergies")));
  AB.rpa_input_energies = &keep.back();

  // integrals of the dimer: produced on the device from the basis sets, else the supplied arrays (as in gwbse_job_run)
  std::unique_ptr<DeviceAOIntegrals> device_ints;
  const AOIntegralSource* ints = nullptr;
  MatrixXd overlap;
  const bool have_arrays = job->ints.ao3c || job->ints.ao3c_dev || job->ints.fn;
  if (!have_arrays && job->basis_data[0] && job->basis_data[1]) {
    for (int b = 0; b < 2; ++b)
      if (!job->dev_basis[b]) job->dev_basis[b] = std::make_unique<DeviceAOBasis>(*job->dev, *job->basis_data[b]);
    device_ints = std::make_unique<DeviceAOIntegrals>(*job->dev, *job->dev_basis[1], *job->dev_basis[0], opt("aux_overlap"),
                                                      opt("aux_coulomb"));
    ints = device_ints.get();
    if (!opt("dft_overlap")) {  // CouplingBase::CalculateOverlapMatrix on the device
      const Index n = job->dev_basis[0]->AOBasisSize();
      overlap = MatrixXd(n, n);
      job->dev->check(gwbse_ao_overlap(job->dev->ctx(), job->dev_basis[0]->handle(), overlap.data(), (int)n));
    }
  } else {
    job->ints.S = &need("aux_overlap");
    job->ints.V = &need("aux_coulomb");
    ints = &job->ints;
  }
  AB.overlap = opt("dft_overlap") ? opt("dft_overlap") : &overlap;
  if (AB.overlap->rows() != AB.mos->rows()) throw std::runtime_error("input array 'dft_overlap' not set");

  BSECoupling coupling(*job->dev, job->log);
  coupling.Initialize(co);
  coupling.CalculateCouplings(A, B, AB, *ints);
  job->coupling_xml = coupling.Addoutput();
  auto& o = job->out;
  o.clear();
  job->out_scalars.clear();
  for (int spin = 0; spin < 2; ++spin) {
    if (!(spin == 0 ? coupling.doSinglets() : coupling.doTriplets())) continue;
    const BSECoupling::Channel& ch = spin == 0 ? coupling.singlet() : coupling.triplet();
    const std::string s = spin == 0 ? "singlet" : "triplet";
    o["JAB_" + s + "_pert"] = ch.JAB[0];
    o["JAB_" + s + "_diag"] = ch.JAB[1];
    o["J_dimer_" + s] = ch.J_dimer;
    o["S_dimer_" + s] = ch.S_dimer;
    job->out_scalars["xi_" + s] = ch.diag.xi;
    job->out_scalars["pt_rm_discrepancy_" + s] = ch.diag.pt_rm_discrepancy;
    job->out_scalars["downfolding_safe_" + s] = ch.diag.downfolding_safe ? 1.0 : 0.0;
  }
  job->out_scalars["levA"] = (double)coupling.levA();
  job->out_scalars["levB"] = (double)coupling.levB();
  JOB_END(job)
}

// Unrestricted reference: GW_UKS + BSE_UKS (tasks "gw" / "exciton_uks").  Alpha channel = the restricted inputs
// ("mos", "mo_energies", "vxc", scalar "homo"), beta channel = "mos_beta", "mo_energies_beta", "vxc_beta", "homo_beta".
int gwbse_job_run_uks(gwbse_job* job) {
  JOB_BEGIN(job)
  auto need = [&](const char* n) -> const MatrixXd& {
    auto it = job->in.find(n);
    if (it == job->in.end()) throw std::runtime_error(std::string("input array '") + n + "' not set");
    return it->second;
  };
  GWBSE::Inputs in;
  in.unrestricted = true;
  if (!job->scalars.count("homo") || !job->scalars.count("homo_beta"))
    throw std::runtime_error("input scalars 'homo' and 'homo_beta' not set");
  in.homo = static_cast<Index>(job->scalars["homo"]);
  in.homo_beta = static_cast<Index>(job->scalars["homo_beta"]);
  in.ScaHFX = job->scalars.count("ScaHFX") ? job->scalars["ScaHFX"] : 0.0;
  const MatrixXd& mos = need("mos");
  const MatrixXd& mos_b = need("mos_beta");
  const VectorXd mo_e = mat2vec(need("mo_energies")), mo_e_b = mat2vec(need("mo_energies_beta"));
  in.mos = &mos;
  in.mos_beta = &mos_b;
  in.mo_energies = &mo_e;
  in.mo_energies_beta = &mo_e_b;
  in.vxc = &need("vxc");
  in.vxc_beta = &need("vxc_beta");
  std::unique_ptr<DeviceAOIntegrals> device_ints;
  const bool have_arrays = job->ints.ao3c || job->ints.ao3c_dev || job->ints.fn;
  if (!have_arrays && job->basis_data[0] && job->basis_data[1]) {
    for (int b = 0; b < 2; ++b)
      if (!job->dev_basis[b]) job->dev_basis[b] = std::make_unique<DeviceAOBasis>(*job->dev, *job->basis_data[b]);
    device_ints = std::make_unique<DeviceAOIntegrals>(
        *job->dev, *job->dev_basis[1], *job->dev_basis[0], job->in.count("aux_overlap") ? &job->in["aux_overlap"] : nullptr,
        job->in.count("aux_coulomb") ? &job->in["aux_coulomb"] : nullptr);
    in.integrals = device_ints.get();
  } else {
    job->ints.S = &need("aux_overlap");
    job->ints.V = &need("aux_coulomb");
    in.integrals = &job->ints;
  }
  // AO dipole matrices <mu|r_k|nu>: from the dft basis on the device, or the arrays ao_dipole_x / _y / _z
  std::vector<MatrixXd> ao_dip;
  if (job->in.count("ao_dipole_x") && job->in.count("ao_dipole_y") && job->in.count("ao_dipole_z")) {
    ao_dip = {job->in["ao_dipole_x"], job->in["ao_dipole_y"], job->in["ao_dipole_z"]};
    in.ao_dipoles = &ao_dip;
  } else if (job->basis_data[0]) {
    if (!job->dev_basis[0]) job->dev_basis[0] = std::make_unique<DeviceAOBasis>(*job->dev, *job->basis_data[0]);
    ao_dip = job->dev_basis[0]->Dipoles();
    in.ao_dipoles = &ao_dip;
  }
  GWBSE gwbse(*job->dev, job->log);
  gwbse.Initialize(job->options, in);
  GWBSE::ResultsUKS r = gwbse.EvaluateUKS();
  if (!job->summary_path.empty() && job->dev->rank() == 0)
    gwbse.WriteSummaryXML(r, job->summary_path,
                          job->scalars.count("dft_total_energy") ? job->scalars["dft_total_energy"] : 0.0);
  auto& o = job->out;
  o.clear();
  for (int s = 0; s < 2; ++s) {
    const std::string sp = s == 0 ? "_alpha" : "_beta";
    o["RPA_inputenergies" + sp] = vec2mat(r.RPA_inputenergies[s]);
    o["QPpert_energies" + sp] = vec2mat(r.QPpert_energies[s]);
    o["QPdiag_eigenvalues" + sp] = vec2mat(r.QPdiag_eigenvalues[s]);
    o["QPdiag_eigenvectors" + sp] = r.QPdiag_eigenvectors[s];
    o["Hqp" + sp] = r.Hqp[s];
    o["Sigma_x" + sp] = r.Sigma_x[s];
    o["Sigma_c" + sp] = r.Sigma_c[s];
  }
  o["BSE_uks_eigenvalues"] = vec2mat(r.BSE_uks.eigenvalues);
  o["BSE_uks_eigenvectors"] = r.BSE_uks.eigenvectors;
  o["BSE_uks_eigenvectors2"] = r.BSE_uks.eigenvectors2;
  o["BSE_uks_dynamic"] = vec2mat(r.BSE_uks_dynamic);
  o["uks_oscillator_strengths"] = vec2mat(r.oscillator_strengths);
  MatrixXd utd(3, static_cast<Index>(r.transition_dipoles.size()));
  for (size_t st = 0; st < r.transition_dipoles.size(); ++st)
    for (Index i = 0; i < 3; ++i) utd(i, st) = r.transition_dipoles[st](i);
  o["uks_transition_dipoles"] = utd;
  auto& sc = job->out_scalars;
  sc.clear();
  sc["rpamin"] = r.rpamin;
  sc["rpamax"] = r.rpamax;
  sc["qpmin"] = r.qpmin;
  sc["qpmax"] = r.qpmax;
  sc["bse_vmin"] = r.bse_vmin;
  sc["bse_cmax"] = r.bse_cmax;
  sc["bse_alpha_size"] = r.alpha_size;
  sc["bse_beta_size"] = r.beta_size;
  sc["gw_iterations"] = r.gw_iterations;
  sc["uks_davidson_iterations"] = r.davidson_iterations;
  sc["uks_converged"] = r.BSE_uks.success;
  sc["removed_functions"] = r.removed_functions;
  sc["time_fill"] = r.time_fill;
  sc["time_gw"] = r.time_gw;
  sc["time_bse"] = r.time_bse;
  JOB_END(job)
}

const char* gwbse_job_coupling_xml(const gwbse_job* job) { return job ? job->coupling_xml.c_str() : ""; }

int gwbse_job_array_dims(const gwbse_job* job, const char* name, long* rows, long* cols) {
  if (!job) return 1;
  auto it = job->out.find(name);
  if (it == job->out.end()) return 1;
  *rows = it->second.rows();
  *cols = it->second.cols();
  return 0;
}

int gwbse_job_get_array(const gwbse_job* job, const char* name, double* out) {
  if (!job) return 1;
  auto it = job->out.find(name);
  if (it == job->out.end()) return 1;
  std::memcpy(out, it->second.data(), sizeof(double) * it->second.size());
  return 0;
}

int gwbse_job_get_scalar(const gwbse_job* job, const char* name, double* out) {
  if (!job) return 1;
  auto it = job->out_scalars.find(name);
  if (it == job->out_scalars.end()) return 1;
  *out = it->second;
  return 0;
}

long long gwbse_job_launch_count(const gwbse_job* job) { return job ? gwbse_launch_count(job->dev->ctx()) : 0; }

}  // extern "C"
