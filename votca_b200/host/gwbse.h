// GWBSE - host mirror of the orchestrator xtp/src/libxtp/gwbse/gwbse.cc:60-578 (Initialize: options ->
// level ranges and solver settings, same option keys as share/xtp/xml/subpackages/gwbse.xml) and
// gwbse.cc:822-1250 (Evaluate: Mmn -> GW -> BSE).  Inputs that the reference takes from the Orbitals
// object (MO coefficients/energies, homo, Vxc from libxc, AO integrals from libint, AO dipoles) are
// supplied by the caller; outputs use the dataset names of the .orb file (orbitals.cc:990-1063).
#pragma once
#include <chrono>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>

#include "bse.h"
#include "checkpoint.h"
#include "gw.h"
#include "uks.h"

namespace votca {
namespace xtp {

// Minimal stand-in for tools::Property restricted to what GWBSE::Initialize reads: a flat map of
// dotted keys ("gw.mode", "bse.davidson.tolerance", ...) with the defaults of gwbse.xml.
class Options {
 public:
  Options() {
    // share/xtp/xml/subpackages/gwbse.xml
    const char* defaults[][2] = {
        {"tasks", "all"}, {"ranges", "default"}, {"ignore_corelevels", "none"},
        {"gw.mode", "evGW"}, {"gw.scissor_shift", "0.0"}, {"gw.sigma_integrator", "ppm"}, {"gw.eta", "1e-3"},
        {"gw.alpha", "1e-3"}, {"gw.quadrature_scheme", "legendre"}, {"gw.quadrature_order", "12"},
        {"gw.qp_solver", "grid"}, {"gw.qp_grid_search_mode", "adaptive_with_dense_fallback"},
        {"gw.qp_root_finder", "bisection"}, {"gw.qp_restrict_search", "true"}, {"gw.qp_zero_margin", "1e-6"},
        {"gw.qp_virtual_min_energy", "-0.1"}, {"gw.qp_sc_max_iter", "100"}, {"gw.qp_sc_limit", "1e-5"},
        {"gw.sc_max_iter", "50"}, {"gw.mixing_order", "20"}, {"gw.sc_limit", "1e-5"}, {"gw.mixing_alpha", "0.7"},
        {"gw.rebuild_3c_freq", "5"}, {"gw.do_qsgw", "false"}, {"gw.qsgw_max_iterations", "20"},
        {"gw.qsgw_sc_limit", "1e-5"}, {"gw.qsgw_max_virt_correction", "0.5"}, {"bse.exctotal", "10"}, {"bse.useTDA", "false"},
        {"bse.dyn_screen_max_iter", "0"}, {"bse.dyn_screen_tol", "1e-5"}, {"bse.davidson.correction", "DPR"},
        {"bse.davidson.tolerance", "normal"}, {"bse.davidson.update", "safe"}, {"bse.davidson.maxiter", "50"},
        {"bse.use_Hqp_offdiag", "false"}, {"bse.print_weight", "0.5"}, {"gw.sigma_plot.steps", "201"},
        {"gw.sigma_plot.spacing", "1e-2"}, {"gw.sigma_plot.filename", "QPenergies_sigma.dat"}};
    for (auto& d : defaults) kv_[d[0]] = d[1];
  }
  void set(const std::string& key, const std::string& value) {
    std::string k = key;
    for (const char* pre : {"options.dftgwbse.gwbse.", "dftgwbse.gwbse.", "gwbse.", "."})
      if (k.rfind(pre, 0) == 0) {
        k = k.substr(std::string(pre).size());
        break;
      }
    check_key(k, value);
    kv_[k] = value;
  }
  // Keys of share/xtp/xml/subpackages/gwbse.xml.  An unknown key is an error (the reference's OptionsHandler rejects
  // it against the same file), and a known key that asks for something this path does not implement is an error
  // too instead of silently different physics.
  static void check_key(const std::string& k, const std::string& value) {
    static const char* known[] = {
        "tasks", "ranges", "rpamax", "qpmin", "qpmax", "bsemin", "bsemax", "ignore_corelevels", "auxbasisset",
        "gw.mode", "gw.scissor_shift", "gw.sigma_integrator", "gw.eta", "gw.alpha", "gw.quadrature_scheme",
        "gw.quadrature_order", "gw.qp_solver", "gw.qp_grid_search_mode", "gw.qp_root_finder",
        "gw.qp_full_window_half_width", "gw.qp_dense_spacing", "gw.qp_adaptive_shell_width",
        "gw.qp_adaptive_shell_count", "gw.qp_grid_steps", "gw.qp_grid_spacing", "gw.qp_restrict_search",
        "gw.qp_zero_margin", "gw.qp_virtual_min_energy", "gw.qp_sc_max_iter", "gw.qp_sc_limit", "gw.sc_max_iter",
        "gw.mixing_order", "gw.sc_limit", "gw.mixing_alpha", "gw.do_qsgw", "gw.qsgw_max_iterations",
        "gw.qsgw_sc_limit", "gw.qsgw_max_virt_correction", "gw.rebuild_3c_freq", "gw.sigma_plot.states",
        "gw.sigma_plot.steps", "gw.sigma_plot.spacing", "gw.sigma_plot.filename", "bse.exctotal", "bse.useTDA",
        "bse.dyn_screen_max_iter", "bse.dyn_screen_tol", "bse.davidson.correction", "bse.davidson.tolerance",
        "bse.davidson.update", "bse.davidson.maxiter", "bse.use_Hqp_offdiag", "bse.print_weight"};
    bool ok = k.rfind("bse.fragments", 0) == 0;
    for (const char* n : known) ok = ok || k == n;
    if (!ok) throw std::runtime_error("unknown option '" + k + "' (not a key of gwbse.xml)");
    if (value.empty()) return;
    if (k.rfind("bse.fragments", 0) == 0 && k != "bse.fragments.fragment.indices")
      throw std::runtime_error("unknown option '" + k + "' (bse.fragments holds fragment.indices elements)");
  }
  bool exists(const std::string& key) const {
    auto it = kv_.find(key);
    return it != kv_.end() && !it->second.empty();
  }
  std::string str(const std::string& key) const {
    auto it = kv_.find(key);
    if (it == kv_.end()) throw std::runtime_error("option '" + key + "' not found");
    return it->second;
  }
  double dbl(const std::string& key) const { return std::stod(str(key)); }
  Index idx(const std::string& key) const { return static_cast<Index>(std::stol(str(key))); }
  bool flag(const std::string& key) const {
    const std::string v = str(key);
    if (v == "true" || v == "1") return true;
    if (v == "false" || v == "0") return false;
    throw std::runtime_error("option '" + key + "' is not a bool: " + v);
  }

  // Reads an options XML as the reference accepts with -o (elements nest, text is the value); every leaf
  // below <gwbse> becomes a dotted key.  Attributes and comments are ignored.
  void LoadFromXML(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw std::runtime_error("cannot open options file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    std::vector<std::string> stack;
    size_t i = 0;
    std::string text;
    while (i < s.size()) {
      if (s[i] == '<') {
        if (s.compare(i, 4, "<!--") == 0) {
          i = s.find("-->", i);
          i = (i == std::string::npos) ? s.size() : i + 3;
          continue;
        }
        size_t j = s.find('>', i);
        if (j == std::string::npos) throw std::runtime_error("malformed XML in " + path);
        std::string tag = s.substr(i + 1, j - i - 1);
        i = j + 1;
        if (tag.empty() || tag[0] == '?' || tag[0] == '!') continue;
        if (tag[0] == '/') {
          std::string t;
          for (char c : text)
            if (!std::isspace(static_cast<unsigned char>(c)) || !t.empty()) t += c;
          while (!t.empty() && std::isspace(static_cast<unsigned char>(t.back()))) t.pop_back();
          if (!t.empty() && !stack.empty()) {
            std::string key;
            bool in_gwbse = false;
            for (const auto& e : stack) {
              if (in_gwbse) key += (key.empty() ? "" : ".") + e;
              if (e == "gwbse") in_gwbse = true;
            }
            if (in_gwbse && !key.empty()) {
              check_key(key, t);
              // one <fragment><indices> element per fragment (gwbse.cc:392-400): kept in order, ';' separated
              if (key == "bse.fragments.fragment.indices" && kv_.count(key) && !kv_[key].empty())
                kv_[key] += ";" + t;
              else
                kv_[key] = t;
            }
          }
          text.clear();
          if (!stack.empty()) stack.pop_back();
        } else {
          const bool selfclose = tag.back() == '/';
          std::string name = tag.substr(0, tag.find_first_of(" \t\r\n/"));
          text.clear();
          if (!selfclose) stack.push_back(name);
        }
      } else {
        text += s[i++];
      }
    }
  }

 private:
  std::map<std::string, std::string> kv_;
};

class GWBSE {
 public:
  struct Inputs {
    Index homo = -1;                       // Orbitals::getHomo()
    const MatrixXd* mos = nullptr;         // MOs().eigenvectors()
    const VectorXd* mo_energies = nullptr;  // MOs().eigenvalues()
    const MatrixXd* vxc = nullptr;         // qptotal x qptotal, CalculateVXC (gwbse.cc:750-788) is host/libxc work
    double ScaHFX = 0.0;                   // Orbitals::getScaHFX()
    const AOIntegralSource* integrals = nullptr;
    const std::vector<MatrixXd>* interlevel_dipoles = nullptr;  // optional: CalcFreeTransition_Dipoles
    // optional alternative: AO dipole matrices <mu|r_k|nu> (AODipole::Fill, e.g. DeviceAOBasis::Dipoles()); the
    // interlevel dipoles are then formed here as Orbitals::CalcFreeTransition_Dipoles does (orbitals.cc:742-760)
    const std::vector<MatrixXd>* ao_dipoles = nullptr;
    // bse.fragments (Lowdin populations of the excitons on groups of atoms, populationanalysis.cc:47-84): the AO
    // overlap of the dft basis, the atom every basis function sits on, the nuclear charges (QMAtom::getNuccharge)
    const MatrixXd* ao_overlap = nullptr;
    const std::vector<Index>* basis_atom = nullptr;
    const VectorXd* nuclear_charges = nullptr;
    // BSE-only runs (do_gw == false): orbitals_.QPdiag / RPAInputEnergies from a previous run
    const MatrixXd* Hqp = nullptr;
    const VectorXd* rpa_input_energies = nullptr;
    // unrestricted reference (Orbitals::hasUnrestrictedOrbitals): homo / mos / mo_energies / vxc above are the alpha
    // channel's, these the beta channel's; evaluated by EvaluateUKS
    bool unrestricted = false;
    Index homo_beta = -1;
    const MatrixXd* mos_beta = nullptr;
    const VectorXd* mo_energies_beta = nullptr;
    const MatrixXd* vxc_beta = nullptr;
  };
  // what GWBSE::Evaluate stores in an unrestricted Orbitals object (gwbse.cc:976-997, 1186-1190)
  struct ResultsUKS {
    Index rpamin = 0, rpamax = 0, qpmin = 0, qpmax = 0, bse_vmin = 0, bse_cmax = 0;
    VectorXd RPA_inputenergies[2], QPpert_energies[2], QPdiag_eigenvalues[2];
    MatrixXd QPdiag_eigenvectors[2], Hqp[2], Sigma_x[2], Sigma_c[2];
    EigenSystem BSE_uks;
    VectorXd BSE_uks_dynamic;
    std::vector<VectorXd> transition_dipoles;  // Orbitals::CalcCoupledTransition_Dipoles(ExcitonUKS)
    VectorXd oscillator_strengths;            // Orbitals::Oscillatorstrengths(ExcitonUKS)
    Index alpha_size = 0, beta_size = 0, gw_iterations = 0, davidson_iterations = 0, removed_functions = 0;
    double time_fill = 0, time_gw = 0, time_bse = 0;
  };
  struct Results {
    Index rpamin = 0, rpamax = 0, qpmin = 0, qpmax = 0, bse_vmin = 0, bse_cmax = 0;
    VectorXd RPA_inputenergies, QPpert_energies;
    VectorXd QPdiag_eigenvalues;
    MatrixXd QPdiag_eigenvectors, Hqp, Sigma_x, Sigma_c;
    EigenSystem BSE_singlet, BSE_triplet;
    VectorXd BSE_singlet_dynamic, BSE_triplet_dynamic;
    std::vector<VectorXd> transition_dipoles;
    VectorXd oscillator_strengths;
    BSE::Interaction singlet_analysis, triplet_analysis;
    // QMFragment<BSE_Population> per fragment (qmfragment.h, bse.cc:362-376): Gs, and H / E per state (nfrag x nstates)
    VectorXd fragment_gs;
    MatrixXd singlet_fragment_hole, singlet_fragment_electron, triplet_fragment_hole, triplet_fragment_electron;
    Index removed_functions = 0, gw_iterations = 0, qsgw_iterations = 0;
    bool is_qsgw = false;
    Index singlet_davidson_iterations = 0, triplet_davidson_iterations = 0;
    std::size_t sigma_batches = 0, sigma_evaluations = 0;
    double time_fill = 0, time_gw = 0, time_bse = 0;
  };

  GWBSE(const Device& dev, Logger& log) : dev_(dev), log_(log) {}

  // gwbse.cc:60-578
  void Initialize(const Options& options, const Inputs& in) {
    in_ = in;
    if (!in.mos || !in.mo_energies || in.homo < 0) throw std::runtime_error("GWBSE: MOs, energies and homo required");
    Index rpamax = 0, rpamin = 0, qpmin = 0, qpmax = 0, bse_vmin = 0, bse_cmax = 0;
    if (in.unrestricted && (!in.mos_beta || !in.mo_energies_beta || in.homo_beta < 0))
      throw std::runtime_error("GWBSE: an unrestricted reference needs beta MOs, energies and homo");
    // gwbse.cc:70-79: the level ranges of an unrestricted reference follow the larger of the two occupations
    const Index homo = in.unrestricted ? std::max(in.homo, in.homo_beta) : in.homo;
    const Index num_of_levels = in.mos->cols();
    const Index num_of_occlevels = homo + 1;
    const std::string ranges = options.str("ranges");
    if (ranges == "factor") {
      rpamax = Index(options.dbl("rpamax") * double(num_of_levels)) - 1;
      qpmin = num_of_occlevels - Index(options.dbl("qpmin") * double(num_of_occlevels)) - 1;
      qpmax = num_of_occlevels + Index(options.dbl("qpmax") * double(num_of_occlevels)) - 1;
      bse_vmin = num_of_occlevels - Index(options.dbl("bsemin") * double(num_of_occlevels)) - 1;
      bse_cmax = num_of_occlevels + Index(options.dbl("bsemax") * double(num_of_occlevels)) - 1;
    } else if (ranges == "explicit") {
      rpamax = options.idx("rpamax");
      qpmin = options.idx("qpmin");
      qpmax = options.idx("qpmax");
      bse_vmin = options.idx("bsemin");
      bse_cmax = options.idx("bsemax");
    } else if (ranges == "default") {
      rpamax = num_of_levels - 1;
      qpmin = 0;
      qpmax = 3 * homo + 1;
      bse_vmin = 0;
      bse_cmax = 3 * homo + 1;
    } else if (ranges == "full") {
      rpamax = num_of_levels - 1;
      qpmin = 0;
      qpmax = num_of_levels - 1;
      bse_vmin = 0;
      bse_cmax = num_of_levels - 1;
    } else {
      throw std::runtime_error("unknown ranges option '" + ranges + "'");
    }
    // gwbse.cc:128-152 with GWBSE::CountCoreLevels (:46-58): core electrons per element as share/xtp/ecps/corelevels.xml
    // lists them, elements identified by their nuclear charge
    const std::string ignore_corelevels = options.str("ignore_corelevels");
    if (ignore_corelevels == "RPA" || ignore_corelevels == "GW" || ignore_corelevels == "BSE") {
      if (!in.nuclear_charges)
        throw std::runtime_error("ignore_corelevels needs the nuclear charges of the atoms (input 'nuclear_charges')");
      const Index ignored_corelevels = CountCoreLevels(*in.nuclear_charges);
      if (ignore_corelevels == "RPA") rpamin = ignored_corelevels;
      if (ignore_corelevels == "GW" || ignore_corelevels == "RPA")
        if (qpmin < ignored_corelevels) qpmin = ignored_corelevels;
      if (bse_vmin < ignored_corelevels) bse_vmin = ignored_corelevels;
      log_(" Ignoring " + std::to_string(ignored_corelevels) + " core levels for " + ignore_corelevels + " and beyond.");
    } else if (ignore_corelevels != "none") {
      throw std::runtime_error("ignore_corelevels is none, RPA, GW or BSE");
    }
    if (rpamax >= num_of_levels) rpamax = num_of_levels - 1;
    if (qpmax >= num_of_levels) qpmax = num_of_levels - 1;
    if (bse_cmax >= num_of_levels) bse_cmax = num_of_levels - 1;
    if (bse_vmin < 0) bse_vmin = 0;
    if (qpmin < 0) qpmin = 0;
    const Index bse_vmax = homo, bse_cmin = homo + 1;
    if (rpamin < 0 || rpamax < 0 || rpamin > rpamax)
      throw std::runtime_error("Invalid RPA level range after setup/clamping.");
    if (qpmin < 0 || qpmax < 0 || qpmin > qpmax) throw std::runtime_error("Invalid GW level range after setup/clamping.");
    if (bse_vmin < 0 || bse_vmax < 0 || bse_vmin > bse_vmax)
      throw std::runtime_error("Invalid BSE occupied level range after setup/clamping.");
    if (bse_cmin < 0 || bse_cmax < 0 || bse_cmin > bse_cmax)
      throw std::runtime_error("Invalid BSE virtual level range after setup/clamping.");
    if (bse_vmax >= num_of_levels || bse_cmax >= num_of_levels)
      throw std::runtime_error("BSE level range exceeds available orbital indices.");
    gwopt_.homo = homo;
    gwopt_.qpmin = qpmin;
    gwopt_.qpmax = qpmax;
    gwopt_.rpamin = rpamin;
    gwopt_.rpamax = rpamax;
    bseopt_.vmin = bse_vmin;
    bseopt_.cmax = bse_cmax;
    bseopt_.homo = homo;
    bseopt_.qpmin = qpmin;
    bseopt_.qpmax = qpmax;
    bseopt_.rpamin = rpamin;
    bseopt_.rpamax = rpamax;
    const Index bse_size = (bse_vmax - bse_vmin + 1) * (bse_cmax - bse_cmin + 1);
    log_(" RPA level range [" + std::to_string(rpamin) + ":" + std::to_string(rpamax) + "]");
    log_(" GW  level range [" + std::to_string(qpmin) + ":" + std::to_string(qpmax) + "]");
    log_(" BSE level range occ[" + std::to_string(bse_vmin) + ":" + std::to_string(bse_vmax) + "]  virt[" +
         std::to_string(bse_cmin) + ":" + std::to_string(bse_cmax) + "]");
    gwopt_.reset_3c = options.idx("gw.rebuild_3c_freq");
    bseopt_.nmax = options.idx("bse.exctotal");
    if (bseopt_.nmax > bse_size || bseopt_.nmax < 0) bseopt_.nmax = bse_size;
    bseopt_.davidson_correction = options.str("bse.davidson.correction");
    bseopt_.davidson_tolerance = options.str("bse.davidson.tolerance");
    bseopt_.davidson_update = options.str("bse.davidson.update");
    bseopt_.davidson_maxiter = options.idx("bse.davidson.maxiter");
    bseopt_.useTDA = options.flag("bse.useTDA");
    log_(bseopt_.useTDA ? " BSE type: TDA" : " BSE type: full");
    const Index full_bse_size = bseopt_.useTDA ? bse_size : 2 * bse_size;
    log_(" BSE Hamiltonian has size " + std::to_string(full_bse_size) + "x" + std::to_string(full_bse_size));
    bseopt_.use_Hqp_offdiag = options.flag("bse.use_Hqp_offdiag");
    bseopt_.max_dyn_iter = options.idx("bse.dyn_screen_max_iter");
    bseopt_.dyn_tolerance = options.dbl("bse.dyn_screen_tol");
    do_dynamical_screening_bse_ = bseopt_.max_dyn_iter > 0;
    const std::string mode = options.str("gw.mode");
    if (mode == "G0W0") {
      gwopt_.gw_sc_max_iterations = 1;
    } else if (mode == "evGW") {
      gwopt_.gw_sc_max_iterations = options.idx("gw.sc_max_iter");
    } else {
      throw std::runtime_error("unknown gw.mode '" + mode + "'");
    }
    log_(" Running GW as: " + mode);
    gwopt_.ScaHFX = in.ScaHFX;
    gwopt_.shift = options.dbl("gw.scissor_shift");
    gwopt_.g_sc_limit = options.dbl("gw.qp_sc_limit");
    gwopt_.g_sc_max_iterations = options.idx("gw.qp_sc_max_iter");
    gwopt_.gw_sc_limit = options.dbl("gw.sc_limit");
    bseopt_.min_print_weight = options.dbl("bse.print_weight");
    std::string tasks = options.str("tasks");
    for (char& c : tasks) c = static_cast<char>(std::tolower(static_cast<unsigned char>(c)));
    // gwbse.cc:345-358: the unrestricted task is driven by GWBSE_UKS (gwbse_job_run_uks); here it is an error as it
    // is in the reference for a restricted Orbitals object
    do_bse_exciton_uks_ = tasks.find("exciton") != std::string::npos;
    if (do_bse_exciton_uks_ && !in.unrestricted)
      throw std::runtime_error("tasks 'excitons' / 'exciton_uks' need an unrestricted reference (gwbse_job_run_uks)");
    if (in.unrestricted && (tasks.find("singlets") != std::string::npos || tasks.find("triplets") != std::string::npos))
      throw std::runtime_error(
          "Invalid gwbse task for UKS reference: 'singlets' and 'triplets' are not defined for open-shell systems.\n"
          "Use 'exciton_uks' (or 'excitons') instead.");
    // gwbse.cc:531-547
    gwopt_.do_qsgw = options.flag("gw.do_qsgw");
    gwopt_.qsgw_max_iterations = options.idx("gw.qsgw_max_iterations");
    gwopt_.qsgw_sc_limit = options.dbl("gw.qsgw_sc_limit");
    gwopt_.qsgw_max_virt_correction = options.dbl("gw.qsgw_max_virt_correction");
    if (gwopt_.do_qsgw)
      log_(" QSGW enabled: max_iter=" + std::to_string(gwopt_.qsgw_max_iterations) +
           " sc_limit=" + std::to_string(gwopt_.qsgw_sc_limit) + " Ha");
    do_gw_ = tasks.find("gw") != std::string::npos;
    if (do_bse_exciton_uks_) do_gw_ = true;
    if (tasks.find("all") != std::string::npos && !in.unrestricted) do_gw_ = do_bse_singlets_ = do_bse_triplets_ = true;
    if (tasks.find("singlets") != std::string::npos) do_bse_singlets_ = true;
    if (tasks.find("triplets") != std::string::npos) do_bse_triplets_ = true;
    gwopt_.sigma_integration = options.str("gw.sigma_integrator");
    log_(" Sigma integration: " + gwopt_.sigma_integration);
    gwopt_.eta = options.dbl("gw.eta");
    if (gwopt_.sigma_integration == "exact")
      log_(" RPA Hamiltonian size: " + std::to_string((homo + 1 - rpamin) * (rpamax - homo)));
    if (gwopt_.sigma_integration == "cda") {
      gwopt_.order = options.idx("gw.quadrature_order");
      gwopt_.quadrature_scheme = options.str("gw.quadrature_scheme");
      gwopt_.alpha = options.dbl("gw.alpha");
    }
    if (options.exists("gw.sigma_plot.states")) {  // gwbse.cc:562-577
      sigma_plot_states_ = options.str("gw.sigma_plot.states");
      sigma_plot_steps_ = options.idx("gw.sigma_plot.steps");
      sigma_plot_spacing_ = options.dbl("gw.sigma_plot.spacing");
      sigma_plot_filename_ = options.str("gw.sigma_plot.filename");
      log_(" Sigma plot states: " + sigma_plot_states_);
      log_(" Sigma plot steps: " + std::to_string(sigma_plot_steps_));
      log_(" Sigma plot spacing: " + std::to_string(sigma_plot_spacing_));
      log_(" Sigma plot filename: " + sigma_plot_filename_);
    }
    fragments_.clear();
    if (options.exists("bse.fragments.fragment.indices")) {  // gwbse.cc:392-400, :878-885
      std::stringstream list(options.str("bse.fragments.fragment.indices"));
      std::string item;
      while (std::getline(list, item, ';')) {
        fragments_.push_back(GW::ParseIndexList(item));
        log_(" Fragment " + std::to_string(fragments_.size() - 1) + " size:" + std::to_string(fragments_.back().size()));
      }
    }
    gwopt_.qp_solver = options.str("gw.qp_solver");
    if (gwopt_.qp_solver == "grid") {
      if (options.exists("gw.qp_full_window_half_width"))
        gwopt_.qp_full_window_half_width = options.dbl("gw.qp_full_window_half_width");
      if (options.exists("gw.qp_dense_spacing")) gwopt_.qp_dense_spacing = options.dbl("gw.qp_dense_spacing");
      if (options.exists("gw.qp_adaptive_shell_width"))
        gwopt_.qp_adaptive_shell_width = options.dbl("gw.qp_adaptive_shell_width");
      if (options.exists("gw.qp_adaptive_shell_count"))
        gwopt_.qp_adaptive_shell_count = options.idx("gw.qp_adaptive_shell_count");
      const bool has_steps = options.exists("gw.qp_grid_steps"), has_spacing = options.exists("gw.qp_grid_spacing");
      if (has_steps != has_spacing)
        throw std::runtime_error(
            "Deprecated gw.qp_grid_steps and gw.qp_grid_spacing must be given together if either is used.");
      if (has_steps) {
        gwopt_.qp_grid_steps = options.idx("gw.qp_grid_steps");
        gwopt_.qp_grid_spacing = options.dbl("gw.qp_grid_spacing");
      }
      qp_solver::NormalizeGridSearchOptions(gwopt_);
    }
    gwopt_.qp_root_finder = options.str("gw.qp_root_finder");
    gwopt_.gw_mixing_order = options.idx("gw.mixing_order");
    gwopt_.gw_mixing_alpha = options.dbl("gw.mixing_alpha");
    gwopt_.qp_grid_search_mode = options.str("gw.qp_grid_search_mode");
    gwopt_.qp_restrict_search = options.flag("gw.qp_restrict_search");
    gwopt_.qp_zero_margin = options.dbl("gw.qp_zero_margin");
    gwopt_.qp_virtual_min_energy = options.dbl("gw.qp_virtual_min_energy");
  }

  const GW::options& gw_options() const { return gwopt_; }
  const BSE::options& bse_options() const { return bseopt_; }

  // gwbse.cc:896-997 + 1150-1190, the is_uks branches: two Mmn tensors (one kernel-library context per spin channel
  // on the same GPU), GW_UKS, then the combined exciton problem
  ResultsUKS EvaluateUKS() {
    using clock = std::chrono::system_clock;
    if (!in_.unrestricted) throw std::runtime_error("GWBSE::EvaluateUKS needs an unrestricted reference");
    if (dev_.world() > 1) throw std::runtime_error("the unrestricted path is single-GPU in this build");
    if (!in_.integrals) throw std::runtime_error("GWBSE: AO integral source required");
    if (!in_.vxc || !in_.vxc_beta) throw std::runtime_error("GWBSE: vxc (alpha and beta) is an input of this path");
    ResultsUKS res;
    res.rpamin = gwopt_.rpamin;
    res.rpamax = gwopt_.rpamax;
    res.qpmin = gwopt_.qpmin;
    res.qpmax = gwopt_.qpmax;
    res.bse_vmin = bseopt_.vmin;
    res.bse_cmax = bseopt_.cmax;
    auto t0 = clock::now();
    Device dev_beta(dev_.index());
    TCMatrix_gwbse_spin Mmn(dev_, dev_beta);
    const Index max_3c = std::max(bseopt_.cmax, gwopt_.qpmax);
    for (int s = 0; s < 2; ++s) {
      TCMatrix_gwbse& M = s == 0 ? Mmn.alpha : Mmn.beta;
      M.Initialize(in_.integrals->AuxSize(), gwopt_.rpamin, max_3c, gwopt_.rpamin, gwopt_.rpamax);
      log_(s == 0 ? " Calculating alpha spin Mmn " : " Calculating beta spin Mmn ");
      M.Fill(*in_.integrals, s == 0 ? *in_.mos : *in_.mos_beta);
      M.device().sync();
    }
    res.removed_functions = Mmn.alpha.Removedfunctions();
    res.time_fill = std::chrono::duration<double>(clock::now() - t0).count();
    t0 = clock::now();
    GW_UKS::options o;
    static_cast<GW::options&>(o) = gwopt_;
    o.homo_alpha = in_.homo;
    o.homo_beta = in_.homo_beta;
    GW_UKS gw(log_, Mmn, *in_.vxc, *in_.vxc_beta, *in_.mo_energies, *in_.mo_energies_beta);
    gw.configure(o);
    gw.CalculateGWPerturbation();
    res.QPpert_energies[0] = gw.getGWAResultsAlpha();
    res.QPpert_energies[1] = gw.getGWAResultsBeta();
    res.RPA_inputenergies[0] = gw.RPAInputEnergiesAlpha();
    res.RPA_inputenergies[1] = gw.RPAInputEnergiesBeta();
    gw.CalculateHQP();
    res.Hqp[0] = gw.getHQPAlpha();
    res.Hqp[1] = gw.getHQPBeta();
    auto ea = gw.DiagonalizeQPHamiltonianAlpha(), eb = gw.DiagonalizeQPHamiltonianBeta();
    res.QPdiag_eigenvalues[0] = ea.first;
    res.QPdiag_eigenvectors[0] = ea.second;
    res.QPdiag_eigenvalues[1] = eb.first;
    res.QPdiag_eigenvectors[1] = eb.second;
    for (int s = 0; s < 2; ++s) {
      res.Sigma_x[s] = gw.Sigma_x(s == 0 ? Spin::Alpha : Spin::Beta);
      res.Sigma_c[s] = gw.Sigma_c(s == 0 ? Spin::Alpha : Spin::Beta);
    }
    res.gw_iterations = gw.iterations();
    res.time_gw = std::chrono::duration<double>(clock::now() - t0).count();
    log_(" UKS GW calculation took " + std::to_string(res.time_gw) + " seconds.");
    if (do_bse_exciton_uks_) {
      t0 = clock::now();
      // Mmn still carries the orthogonal plasmon-pole rotation of the last GW iteration; eps(0), its spectrum and
      // the operator are invariant under it (as in the restricted path, where BSE follows GW on the same tensor)
      BSE_UKS bse(log_, Mmn);
      bse.configure(bseopt_, in_.homo, in_.homo_beta, res.RPA_inputenergies[0], res.RPA_inputenergies[1], res.Hqp[0],
                    res.Hqp[1]);
      res.BSE_uks = bse.Solve_excitons_uks();
      log_(" Solved combined UKS BSE exciton problem ");
      ExcitonUKSOperator_TDA H = bse.getExcitonOperator_TDA();
      res.alpha_size = H.alpha_size();
      res.beta_size = H.beta_size();
      res.davidson_iterations = bse.last_davidson_iterations();
      if (in_.ao_dipoles) {
        // Orbitals::CalcCoupledTransition_Dipoles(ExcitonUKS) (orbitals.cc:798-877): d = -(sum (X+Y)_alpha o D_alpha
        // + sum (X+Y)_beta o D_beta), no sqrt(2): both spin sectors are explicit components of the eigenvector;
        // f = 2/3 Omega |d|^2 (orbitals.cc:645-674)
        const std::vector<MatrixXd> Da = CalcFreeTransition_Dipoles(*in_.ao_dipoles, *in_.mos, in_.homo),
                                    Db = CalcFreeTransition_Dipoles(*in_.ao_dipoles, *in_.mos_beta, in_.homo_beta);
        const Index nst = res.BSE_uks.eigenvalues.size(), na = res.alpha_size;
        res.oscillator_strengths = VectorXd(nst);
        log_(bseopt_.useTDA ? "  ====== combined UKS TDA exciton energies (eV) ====== "
                            : "  ====== combined UKS full-BSE exciton energies (eV) ====== ");
        for (Index s = 0; s < nst; ++s) {
          VectorXd d(3, 0.0);
          for (int k = 0; k < 3; ++k) {
            double acc = 0.0;
            for (int sp = 0; sp < 2; ++sp) {
              const MatrixXd& D = (sp == 0 ? Da : Db)[static_cast<size_t>(k)];  // ct x vt, index c + ct * v
              const Index off = sp == 0 ? 0 : na;
              for (Index j = 0; j < D.cols(); ++j)
                for (Index i = 0; i < D.rows(); ++i) {
                  double c = res.BSE_uks.eigenvectors(off + j * D.rows() + i, s);
                  if (!bseopt_.useTDA) c += res.BSE_uks.eigenvectors2(off + j * D.rows() + i, s);
                  acc += c * D(i, j);
                }
            }
            d(k) = -acc;
          }
          res.transition_dipoles.push_back(d);
          const double d2 = d(0) * d(0) + d(1) * d(1) + d(2) * d(2);
          res.oscillator_strengths(s) = d2 * 2.0 / 3.0 * res.BSE_uks.eigenvalues(s);
          char buf[200];
          std::snprintf(buf, sizeof(buf), "  XU%-4ld %+1.6f", (long)(s + 1), res.BSE_uks.eigenvalues(s) * 27.21138602);
          log_(buf);
          std::snprintf(buf, sizeof(buf),
                        "           TrDipole length gauge[e*bohr]  dx = %+1.4f dy = %+1.4f dz = %+1.4f |d|^2 = %+1.4f f = %+1.4f",
                        d(0), d(1), d(2), d2, res.oscillator_strengths(s));
          log_(buf);
          PrintWeightsUKS(res, s);
        }
      }
      if (do_dynamical_screening_bse_)  // gwbse.cc:1207-1210
        res.BSE_uks_dynamic =
            bse.Perturbative_DynamicalScreening(res.BSE_uks, res.RPA_inputenergies[0], res.RPA_inputenergies[1]);
      res.time_bse = std::chrono::duration<double>(clock::now() - t0).count();
      log_(" BSE calculation took " + std::to_string(res.time_bse) + " seconds.");
    }
    log_(" GWBSE calculation finished ");
    return res;
  }

  // gwbse.cc:822-1250
  Results Evaluate() {
    using clock = std::chrono::system_clock;
    Results res;
    res.rpamin = gwopt_.rpamin;
    res.rpamax = gwopt_.rpamax;
    res.qpmin = gwopt_.qpmin;
    res.qpmax = gwopt_.qpmax;
    res.bse_vmin = bseopt_.vmin;
    res.bse_cmax = bseopt_.cmax;
    if (!in_.integrals) throw std::runtime_error("GWBSE: AO integral source required");
    if (!do_gw_ && !(in_.Hqp && in_.rpa_input_energies))
      throw std::runtime_error("You want no GW calculation but the orb file has no matching QP coefficients.");
    auto t0 = clock::now();
    TCMatrix_gwbse Mmn(dev_);
    const Index max_3c = std::max(bseopt_.cmax, gwopt_.qpmax);
    Mmn.Initialize(in_.integrals->AuxSize(), gwopt_.rpamin, max_3c, gwopt_.rpamin, gwopt_.rpamax);
    log_(" Calculating Mmn_beta (3-center-repulsion x orbitals)  ");
    Mmn.Fill(*in_.integrals, *in_.mos);
    res.removed_functions = Mmn.Removedfunctions();
    log_(" Removed " + std::to_string(Mmn.Removedfunctions()) + " functions from Aux Coulomb matrix to avoid near linear dependencies");
    res.time_fill = std::chrono::duration<double>(clock::now() - t0).count();
    MatrixXd Hqp;
    VectorXd rpa_e;
    if (do_gw_) {
      auto start = clock::now();
      if (!in_.vxc) throw std::runtime_error("GWBSE: Vxc matrix required for GW");
      GW gw(log_, Mmn, *in_.vxc, *in_.mo_energies);
      gw.configure(gwopt_);
      gw.CalculateGWPerturbation();
      if (!sigma_plot_states_.empty())  // gwbse.cc:1009-1012
        gw.PlotSigma(sigma_plot_filename_, sigma_plot_steps_, sigma_plot_spacing_, sigma_plot_states_);
      res.QPpert_energies = gw.getGWAResults();
      res.RPA_inputenergies = gw.RPAInputEnergies();
      if (gwopt_.do_qsgw) {
        // gwbse.cc:1019-1075: QSGW loop, then Mmn rebuilt in the QP wavefunction basis for the BSE, Hqp diagonal
        gw.CalculateQSGW();
        const VectorXd qsgw_energies = gw.getGWAResults();
        const MatrixXd& U = gw.getQSGWRotation();
        const Index qptotal = gwopt_.qpmax - gwopt_.qpmin + 1;
        qsgw_mos_ = *in_.mos;
        {
          // C_qp[:, qp window] = C[:, qp window] U (small host product: N x q x q)
          const MatrixXd Cw = in_.mos->block(0, gwopt_.qpmin, in_.mos->rows(), qptotal) * U;
          qsgw_mos_.setBlock(0, gwopt_.qpmin, Cw);
        }
        log_(" Rebuilding Mmn in QSGW QP wavefunction basis");
        Mmn.Fill(*in_.integrals, qsgw_mos_);
        Hqp = asDiagonal(qsgw_energies);
        res.QPdiag_eigenvalues = qsgw_energies;
        res.QPdiag_eigenvectors = U;
        res.is_qsgw = true;
        res.qsgw_iterations = gw.qsgw_iterations();
        res.RPA_inputenergies = gw.RPAInputEnergies();
      } else {
        gw.CalculateHQP();
        Hqp = gw.getHQP();
        auto es = gw.DiagonalizeQPHamiltonian();
        res.QPdiag_eigenvalues = es.first;
        res.QPdiag_eigenvectors = es.second;
      }
      res.Sigma_x = gw.Sigma_x();
      res.Sigma_c = gw.Sigma_c();
      res.gw_iterations = gw.iterations();
      res.sigma_batches = gw.sigma_batches();
      res.sigma_evaluations = gw.sigma_evaluations();
      res.time_gw = std::chrono::duration<double>(clock::now() - start).count();
      char buf[96];
      std::snprintf(buf, sizeof(buf), " GW calculation took %f seconds.", res.time_gw);
      log_(buf);
      rpa_e = res.RPA_inputenergies;
    } else {
      Hqp = *in_.Hqp;
      rpa_e = *in_.rpa_input_energies;
      res.RPA_inputenergies = rpa_e;
    }
    res.Hqp = Hqp;
    if (do_bse_singlets_ || do_bse_triplets_) {
      auto start = clock::now();
      BSE bse(log_, Mmn);
      bse.configure(bseopt_, rpa_e, Hqp);
      if (do_bse_triplets_) {
        res.BSE_triplet = bse.Solve_triplets();
        res.triplet_davidson_iterations = bse.last_davidson_iterations();
        res.triplet_analysis = bse.Analyze_eh_interaction(false, res.BSE_triplet);
        if (!fragments_.empty())
          FragmentPopulations(res.BSE_triplet, res.fragment_gs, res.triplet_fragment_hole, res.triplet_fragment_electron);
        ReportStates(false, res.BSE_triplet, res.triplet_analysis, res, res.triplet_fragment_hole,
                     res.triplet_fragment_electron);
      }
      if (do_bse_singlets_) {
        res.BSE_singlet = bse.Solve_singlets();
        res.singlet_davidson_iterations = bse.last_davidson_iterations();
        std::vector<MatrixXd> free_dipoles;
        const std::vector<MatrixXd>* interlevel = in_.interlevel_dipoles;
        if (!interlevel && in_.ao_dipoles) {
          free_dipoles = CalcFreeTransition_Dipoles(*in_.ao_dipoles);
          interlevel = &free_dipoles;
        }
        if (interlevel) {
          res.transition_dipoles = bse.CalcCoupledTransition_Dipoles(res.BSE_singlet, *interlevel);
          res.oscillator_strengths = BSE::Oscillatorstrengths(res.transition_dipoles, res.BSE_singlet.eigenvalues);
        }
        res.singlet_analysis = bse.Analyze_eh_interaction(true, res.BSE_singlet);
        if (!fragments_.empty())
          FragmentPopulations(res.BSE_singlet, res.fragment_gs, res.singlet_fragment_hole, res.singlet_fragment_electron);
        ReportStates(true, res.BSE_singlet, res.singlet_analysis, res, res.singlet_fragment_hole,
                     res.singlet_fragment_electron);
      }
      if (do_dynamical_screening_bse_) {
        if (do_bse_triplets_) res.BSE_triplet_dynamic = bse.Perturbative_DynamicalScreening(res.BSE_triplet, rpa_e);
        if (do_bse_singlets_) res.BSE_singlet_dynamic = bse.Perturbative_DynamicalScreening(res.BSE_singlet, rpa_e);
      }
      res.time_bse = std::chrono::duration<double>(clock::now() - start).count();
      char buf[96];
      std::snprintf(buf, sizeof(buf), " BSE calculation took %f seconds.", res.time_bse);
      log_(buf);
    }
    log_(" GWBSE calculation finished ");
    return res;
  }

  // BSE_UKS::PrintWeightsUKS (bse_uks.cc:520-595): weight of the two spin sectors, then the (at most eight) largest
  // transitions above bse.print_weight, by magnitude
  void PrintWeightsUKS(const ResultsUKS& res, Index state) const {
    struct Contribution {
      double weight;
      bool is_alpha;
      Index v, c;
    };
    const Index na = res.alpha_size, nb = res.beta_size;
    const Index ct[2] = {bseopt_.cmax - in_.homo, bseopt_.cmax - in_.homo_beta};
    const bool tda = res.BSE_uks.eigenvectors2.size() == 0;
    auto weight = [&](Index k) {
      double w = res.BSE_uks.eigenvectors(k, state) * res.BSE_uks.eigenvectors(k, state);
      if (!tda) w -= res.BSE_uks.eigenvectors2(k, state) * res.BSE_uks.eigenvectors2(k, state);
      return w;
    };
    double sector[2] = {0.0, 0.0};
    std::vector<Contribution> contributions;
    for (Index k = 0; k < na + nb; ++k) {
      const bool a = k < na;
      const Index i = a ? k : k - na;
      const double w = weight(k);
      sector[a ? 0 : 1] += w;
      if (std::abs(w) > bseopt_.min_print_weight) contributions.push_back({w, a, i / ct[a ? 0 : 1], i % ct[a ? 0 : 1]});
    }
    char buf[160];
    std::snprintf(buf, sizeof(buf), "           alpha-sector: %+6.2f%%   beta-sector: %+6.2f%%", 100.0 * sector[0],
                  100.0 * sector[1]);
    log_(buf);
    std::stable_sort(contributions.begin(), contributions.end(),
                     [](const Contribution& x, const Contribution& y) { return std::abs(x.weight) > std::abs(y.weight); });
    const size_t nprint = std::min<size_t>(8, contributions.size());
    for (size_t i = 0; i < nprint; ++i) {
      const Contribution& c = contributions[i];
      const Index homo = c.is_alpha ? in_.homo : in_.homo_beta;
      std::snprintf(buf, sizeof(buf), "           [%s] HOMO-%-3ld -> LUMO+%-3ld  : %+6.2f%%", c.is_alpha ? "alpha" : "beta ",
                    (long)(homo - (bseopt_.vmin + c.v)), (long)c.c, 100.0 * c.weight);
      log_(buf);
    }
    if (contributions.size() > nprint) {
      std::snprintf(buf, sizeof(buf), "           ... %ld more contributions above threshold",
                    (long)(contributions.size() - nprint));
      log_(buf);
    }
  }

  // The per-state report of BSE::Analyze_singlets / Analyze_triplets (bse.cc:394-490) with PrintWeights (:378-392) and
  // printFragInfo (:362-376), line for line
  void ReportStates(bool singlet, const EigenSystem& es, const BSE::Interaction& act, const Results& res,
                    const MatrixXd& frag_h, const MatrixXd& frag_e) const {
    const double hrt2ev = 27.21138602;
    log_(singlet ? "  ====== singlet energies (eV) ====== " : "  ====== triplet energies (eV) ====== ");
    const Index vt = bseopt_.homo - bseopt_.vmin + 1, ct = bseopt_.cmax - bseopt_.homo;
    const bool tda = es.eigenvectors2.size() == 0;
    char buf[256];
    for (Index i = 0; i < std::min<Index>(bseopt_.nmax, es.eigenvalues.size()); ++i) {
      const double e = hrt2ev * es.eigenvalues(i);
      if (singlet)
        std::snprintf(buf, sizeof(buf),
                      "  %2s = %4ld Omega = %+1.12f eV  lamdba = %+3.2f nm <FT> = %+1.4f <K_x> = %+1.4f <K_d> = %+1.4f", "S",
                      (long)(i + 1), e, 1240.0 / e, hrt2ev * act.qp_contrib(i), hrt2ev * act.exchange_contrib(i),
                      hrt2ev * act.direct_contrib(i));
      else
        std::snprintf(buf, sizeof(buf), "  %2s = %4ld Omega = %+1.12f eV  lamdba = %+3.2f nm <FT> = %+1.4f <K_d> = %+1.4f",
                      "T", (long)(i + 1), e, 1240.0 / e, hrt2ev * act.qp_contrib(i), hrt2ev * act.direct_contrib(i));
      log_(buf);
      if (singlet && static_cast<size_t>(i) < res.transition_dipoles.size() && i < res.oscillator_strengths.size()) {
        const VectorXd& d = res.transition_dipoles[static_cast<size_t>(i)];
        std::snprintf(buf, sizeof(buf),
                      "           TrDipole length gauge[e*bohr]  dx = %+1.4f dy = %+1.4f dz = %+1.4f |d|^2 = %+1.4f f = %+1.4f",
                      d(0), d(1), d(2), d(0) * d(0) + d(1) * d(1) + d(2) * d(2), res.oscillator_strengths(i));
        log_(buf);
      }
      for (Index v = 0; v < vt; ++v)  // PrintWeights: index ct * v + c (vc2index)
        for (Index c = 0; c < ct; ++c) {
          const Index k = v * ct + c;
          double w = es.eigenvectors(k, i) * es.eigenvectors(k, i);
          if (!tda) w -= es.eigenvectors2(k, i) * es.eigenvectors2(k, i);
          if (w > bseopt_.min_print_weight) {
            std::snprintf(buf, sizeof(buf), "           HOMO-%-3ld -> LUMO+%-3ld  : %3.1f%%",
                          (long)(bseopt_.homo - (bseopt_.vmin + v)), (long)c, 100.0 * w);
            log_(buf);
          }
        }
      for (Index f = 0; f < frag_h.rows() && i < frag_h.cols(); ++f) {
        const double dq = frag_h(f, i) + frag_e(f, i), qeff = dq + res.fragment_gs(f);
        std::snprintf(buf, sizeof(buf),
                      "           Fragment %4d -- hole: %5.1f%%  electron: %5.1f%%  dQ: %+5.2f  Qeff: %+5.2f", (int)f,
                      100.0 * frag_h(f, i), -100.0 * frag_e(f, i), dq, qeff);
        log_(buf);
      }
      log_(singlet ? "" : "   ");
    }
  }

  // GWBSE::CountCoreLevels (gwbse.cc:46-58): half the core electrons of share/xtp/ecps/corelevels.xml
  static Index CountCoreLevels(const VectorXd& nuclear_charges) {
    Index core_electrons = 0;
    for (Index a = 0; a < nuclear_charges.size(); ++a) {
      const long z = std::lround(nuclear_charges(a));
      switch (z) {
        case 1: break;                                  // H
        case 6: case 7: case 8: case 9: core_electrons += 2; break;   // C N O F
        case 13: case 16: core_electrons += 10; break;  // Al S
        case 47: core_electrons += 28; break;           // Ag
        case 80: core_electrons += 60; break;           // Hg
        default:
          throw std::runtime_error("ignore_corelevels: element with nuclear charge " + std::to_string(z) +
                                   " is not in the corelevels table");
      }
    }
    return core_electrons / 2;
  }

  // Lowdin::CalcChargeperFragment (populationanalysis.cc:47-84) for the excitons of one spin type, with
  // Orbitals::DensityMatrixGroundState / DensityMatrixExcitedState (orbitals.cc:193-223, 516-650) folded in: with
  // T = S^1/2 C, the electrons an AO density C_a K C_b^T puts on basis function mu are sum_kl T[mu,k] K[k,l] T[mu,l],
  // and K = A A^T (CalcAuxMat_vv) or A^T A (CalcAuxMat_cc) makes that a row-wise sum of squares of T_v A or T_c A^T.
  // S^1/2 C (eigensolver + two GEMMs) and the per-state products run on the device; the N x N densities are never
  // formed.
  void FragmentPopulations(const EigenSystem& es, VectorXd& gs, MatrixXd& H, MatrixXd& E) {
    if (!in_.ao_overlap || !in_.basis_atom || !in_.nuclear_charges)
      throw std::runtime_error(
          "bse.fragments needs the AO overlap of the dft basis (or the basis itself), the atom of every basis "
          "function and the nuclear charges");
    const MatrixXd& C = *in_.mos;
    const MatrixXd& S = *in_.ao_overlap;
    const Index N = C.rows(), homo = bseopt_.homo, vmin = bseopt_.vmin, cmax = bseopt_.cmax;
    const Index vt = homo - vmin + 1, ct = cmax - homo, ncol = cmax + 1, nat = in_.nuclear_charges->size();
    if (S.rows() != N || S.cols() != N || static_cast<Index>(in_.basis_atom->size()) != N)
      throw std::runtime_error("bse.fragments: AO overlap / atom map do not match the dft basis");
    for (Index a : *in_.basis_atom)
      if (a < 0 || a >= nat) throw std::runtime_error("bse.fragments: atom index of a basis function out of range");
    for (const auto& f : fragments_)
      for (Index a : f)
        if (a < 0 || a >= nat) throw std::runtime_error("bse.fragments: atom index " + std::to_string(a) + " out of range");
    // T = S^1/2 C = U sqrt(w) U^T C for the levels 0 .. cmax
    Device::Buffer U = dev_.upload(S), Cd = dev_.upload(C), T1 = dev_.alloc(static_cast<size_t>(N * ncol)),
                   Td = dev_.alloc(static_cast<size_t>(N * ncol));
    VectorXd w(N);
    dev_.check(gwbse_sym_eig_dev(dev_.ctx(), (int)N, U.get(), (int)N, w.data()));
    VectorXd sq(N);
    for (Index i = 0; i < N; ++i) {
      if (w(i) <= 0.0) throw std::runtime_error("bse.fragments: AO overlap is not positive definite");
      sq(i) = std::sqrt(w(i));
    }
    Device::Buffer dsq = dev_.upload(sq);
    dev_.gemm('T', 'N', N, ncol, N, 1.0, U.get(), N, Cd.get(), N, 0.0, T1.get(), N);
    dev_.check(gwbse_diag_scale_dev(dev_.ctx(), 'L', (int)N, (int)ncol, T1.get(), (int)N, dsq.get(), T1.get(), (int)N));
    dev_.gemm('N', 'N', N, ncol, N, 1.0, U.get(), N, T1.get(), N, 0.0, Td.get(), N);
    auto per_atom = [&](const MatrixXd& G) {  // electrons per atom: rows of G squared, summed over the atom's functions
      VectorXd out(nat, 0.0);
      for (Index j = 0; j < G.cols(); ++j)
        for (Index mu = 0; mu < N; ++mu) out((*in_.basis_atom)[static_cast<size_t>(mu)]) += G(mu, j) * G(mu, j);
      return out;
    };
    auto per_fragment = [&](const VectorXd& atoms, Index f) {
      double q = 0.0;
      for (Index a : fragments_[static_cast<size_t>(f)]) q += atoms(a);
      return q;
    };
    const Index nfrag = static_cast<Index>(fragments_.size());
    {  // ground state, closed shell: 2 sum_occ T[mu,i]^2
      const MatrixXd Tocc = dev_.download(Td.get(), N, homo + 1);
      VectorXd atoms = per_atom(Tocc);
      gs = VectorXd(nfrag);
      VectorXd q(nat);
      for (Index a = 0; a < nat; ++a) q(a) = (*in_.nuclear_charges)(a) - 2.0 * atoms(a);
      for (Index f = 0; f < nfrag; ++f) gs(f) = per_fragment(q, f);
    }
    const Index nstates = es.eigenvalues.size();
    H = MatrixXd(nfrag, nstates);
    E = MatrixXd(nfrag, nstates);
    const bool tda = es.eigenvectors2.size() == 0;
    Device::Buffer Ad = dev_.alloc(static_cast<size_t>(vt * ct)), Gv = dev_.alloc(static_cast<size_t>(N * ct)),
                   Gc = dev_.alloc(static_cast<size_t>(N * vt));
    const double* Tv = Td.get() + vmin * N;
    const double* Tc = Td.get() + (homo + 1) * N;
    // coefficient vector as the ct x vt matrix A(c, v) (orbitals.cc:578-590):
    // vv[mu] = sum_c (sum_v T_v[mu,v] A(c,v))^2, cc[mu] = sum_v (sum_c T_c[mu,c] A(c,v))^2
    auto densities = [&](const MatrixXd& coeffs, Index state, VectorXd& vv, VectorXd& cc) {
      dev_.check(gwbse_h2d(dev_.ctx(), Ad.get(), coeffs.data() + state * coeffs.rows(), static_cast<size_t>(vt * ct)));
      dev_.gemm('N', 'T', N, ct, vt, 1.0, Tv, N, Ad.get(), ct, 0.0, Gv.get(), N);
      dev_.gemm('N', 'N', N, vt, ct, 1.0, Tc, N, Ad.get(), ct, 0.0, Gc.get(), N);
      vv = per_atom(dev_.download(Gv.get(), N, ct));
      cc = per_atom(dev_.download(Gc.get(), N, vt));
    };
    for (Index s = 0; s < nstates; ++s) {
      VectorXd xvv, xcc;
      densities(es.eigenvectors, s, xvv, xcc);
      VectorXd atom_h = xvv, atom_e(nat);
      for (Index a = 0; a < nat; ++a) atom_e(a) = -xcc(a);
      if (!tda) {  // antiresonant part (orbitals.cc:523-530): hole -= C_c Y_cc C_c^T, electron -= C_v Y_vv C_v^T
        VectorXd yvv, ycc;
        densities(es.eigenvectors2, s, yvv, ycc);
        for (Index a = 0; a < nat; ++a) {
          atom_h(a) -= ycc(a);
          atom_e(a) += yvv(a);
        }
      }
      for (Index f = 0; f < nfrag; ++f) {
        H(f, s) = per_fragment(atom_h, f);
        E(f, s) = per_fragment(atom_e, f);
      }
    }
  }

  // Orbitals::CalcFreeTransition_Dipoles (orbitals.cc:742-760): interlevel[k] = empty^T * D_k * occ with
  // empty = MOs(:, bse_cmin..bse_cmax), occ = MOs(:, bse_vmin..homo); two small GEMMs per direction on the device
  std::vector<MatrixXd> CalcFreeTransition_Dipoles(const std::vector<MatrixXd>& ao_dipoles) const {
    return CalcFreeTransition_Dipoles(ao_dipoles, *in_.mos, bseopt_.homo);
  }
  // the same for one spin channel of an unrestricted reference (orbitals.cc:826-846): the channel's MOs and homo
  std::vector<MatrixXd> CalcFreeTransition_Dipoles(const std::vector<MatrixXd>& ao_dipoles, const MatrixXd& C,
                                                   Index homo) const {
    const Index N = C.rows(), vt = homo - bseopt_.vmin + 1, ct = bseopt_.cmax - homo;
    if (ao_dipoles.size() != 3) throw std::runtime_error("three AO dipole matrices expected");
    Device::Buffer Cd = dev_.upload(C), T = dev_.alloc(static_cast<size_t>(N * vt)),
                   I = dev_.alloc(static_cast<size_t>(std::max<Index>(ct * vt, 1)));
    std::vector<MatrixXd> out;
    for (const MatrixXd& D : ao_dipoles) {
      if (D.rows() != N || D.cols() != N) throw std::runtime_error("AO dipole matrix does not match the basis size");
      Device::Buffer Dd = dev_.upload(D);
      dev_.gemm('N', 'N', N, vt, N, 1.0, Dd.get(), N, Cd.get() + bseopt_.vmin * N, N, 0.0, T.get(), N);
      dev_.gemm('T', 'N', ct, vt, N, 1.0, Cd.get() + (homo + 1) * N, N, T.get(), N, 0.0, I.get(), ct);
      out.push_back(dev_.download(I.get(), ct, vt));
    }
    return out;
  }

  // GWBSE::addoutput (gwbse.cc:580-738) as the dftgwbse tool writes it to <job>_summary.xml
  // (tools/dftgwbse.cc:120-128: the "output" property, tab-indented, attributes in alphabetical order; energies in
  // eV with boost::format("%+1.6f ")).  dft_total_energy: Orbitals::getDFTTotalEnergy() in Hartree.
  void WriteSummaryXML(const Results& r, const std::string& filename, double dft_total_energy) const {
    const double hrt2ev = 27.21138602;  // tools::conv::hrt2ev, tools/include/votca/tools/constants.h:53
    auto ev = [&](double x) {
      char b[48];
      std::snprintf(b, sizeof(b), "%+1.6f ", x * hrt2ev);
      return std::string(b);
    };
    std::ofstream f(filename);
    if (!f) throw std::runtime_error("cannot write summary file " + filename);
    f << "<output>\n";
    if (do_gw_) {
      f << "\t<GWBSE DFTEnergy=\"" << ev(dft_total_energy) << "\" units=\"eV\">\n";
      f << "\t\t<dft HOMO=\"" << gwopt_.homo << "\" LUMO=\"" << gwopt_.homo + 1 << "\">\n";
      for (Index state = 0; state < gwopt_.qpmax + 1 - gwopt_.qpmin; ++state) {
        f << "\t\t\t<level number=\"" << state + gwopt_.qpmin << "\">\n";
        f << "\t\t\t\t<dft_energy>" << ev((*in_.mo_energies)(state + gwopt_.qpmin)) << "</dft_energy>\n";
        f << "\t\t\t\t<gw_energy>" << ev(r.QPpert_energies(state)) << "</gw_energy>\n";
        f << "\t\t\t\t<qp_energy>" << ev(r.QPdiag_eigenvalues(state)) << "</qp_energy>\n";
        f << "\t\t\t</level>\n";
      }
      f << "\t\t</dft>\n";
    } else {
      f << "\t<GWBSE>\n";
    }
    if (do_bse_singlets_) {
      f << "\t\t<singlets>\n";
      for (Index state = 0; state < std::min<Index>(bseopt_.nmax, r.BSE_singlet.eigenvalues.size()); ++state) {
        f << "\t\t\t<level number=\"" << state + 1 << "\">\n";
        f << "\t\t\t\t<omega>" << ev(r.BSE_singlet.eigenvalues(state)) << "</omega>\n";
        if (static_cast<size_t>(state) < r.transition_dipoles.size()) {
          const VectorXd& d = r.transition_dipoles[static_cast<size_t>(state)];
          const double fosc = 2 * (d(0) * d(0) + d(1) * d(1) + d(2) * d(2)) * r.BSE_singlet.eigenvalues(state) / 3.0;
          char b[96];
          std::snprintf(b, sizeof(b), "%+1.6f ", fosc);
          f << "\t\t\t\t<f>" << b << "</f>\n";
          std::snprintf(b, sizeof(b), "%+1.4f %+1.4f %+1.4f", d(0), d(1), d(2));
          f << "\t\t\t\t<Trdipole gauge=\"length\" unit=\"e*bohr\">" << b << "</Trdipole>\n";
        }
        f << "\t\t\t</level>\n";
      }
      f << "\t\t</singlets>\n";
    }
    if (do_bse_triplets_) {
      f << "\t\t<triplets>\n";
      for (Index state = 0; state < std::min<Index>(bseopt_.nmax, r.BSE_triplet.eigenvalues.size()); ++state) {
        f << "\t\t\t<level number=\"" << state + 1 << "\">\n";
        f << "\t\t\t\t<omega>" << ev(r.BSE_triplet.eigenvalues(state)) << "</omega>\n";
        f << "\t\t\t</level>\n";
      }
      f << "\t\t</triplets>\n";
    }
    f << "\t</GWBSE>\n</output>\n";
  }

  // the same file for an unrestricted reference (gwbse.cc:591-633 dft_alpha / dft_beta, :696-736 exciton_uks)
  void WriteSummaryXML(const ResultsUKS& r, const std::string& filename, double dft_total_energy) const {
    const double hrt2ev = 27.21138602;
    auto ev = [&](double x) {
      char b[48];
      std::snprintf(b, sizeof(b), "%+1.6f ", x * hrt2ev);
      return std::string(b);
    };
    std::ofstream f(filename);
    if (!f) throw std::runtime_error("cannot write summary file " + filename);
    f << "<output>\n";
    if (do_gw_) {
      f << "\t<GWBSE DFTEnergy=\"" << ev(dft_total_energy) << "\" units=\"eV\">\n";
      for (int s = 0; s < 2; ++s) {
        const Index homo = s == 0 ? in_.homo : in_.homo_beta;
        const VectorXd& e = s == 0 ? *in_.mo_energies : *in_.mo_energies_beta;
        const char* tag = s == 0 ? "dft_alpha" : "dft_beta";
        f << "\t\t<" << tag << " HOMO=\"" << homo << "\" LUMO=\"" << homo + 1 << "\">\n";
        for (Index state = 0; state < gwopt_.qpmax + 1 - gwopt_.qpmin; ++state) {
          f << "\t\t\t<level number=\"" << state + gwopt_.qpmin << "\">\n";
          f << "\t\t\t\t<dft_energy>" << ev(e(state + gwopt_.qpmin)) << "</dft_energy>\n";
          f << "\t\t\t\t<gw_energy>" << ev(r.QPpert_energies[s](state)) << "</gw_energy>\n";
          f << "\t\t\t\t<qp_energy>" << ev(r.QPdiag_eigenvalues[s](state)) << "</qp_energy>\n";
          f << "\t\t\t</level>\n";
        }
        f << "\t\t</" << tag << ">\n";
      }
    } else {
      f << "\t<GWBSE>\n";
    }
    if (do_bse_exciton_uks_) {
      f << "\t\t<exciton_uks>\n";
      for (Index state = 0; state < std::min<Index>(bseopt_.nmax, r.BSE_uks.eigenvalues.size()); ++state) {
        f << "\t\t\t<level number=\"" << state + 1 << "\">\n";
        f << "\t\t\t\t<omega>" << ev(r.BSE_uks.eigenvalues(state)) << "</omega>\n";
        if (static_cast<size_t>(state) < r.transition_dipoles.size() && state < r.oscillator_strengths.size()) {
          const VectorXd& d = r.transition_dipoles[static_cast<size_t>(state)];
          char b[96];
          std::snprintf(b, sizeof(b), "%+1.6f ", r.oscillator_strengths(state));
          f << "\t\t\t\t<f>" << b << "</f>\n";
          std::snprintf(b, sizeof(b), "%+1.4f %+1.4f %+1.4f", d(0), d(1), d(2));
          f << "\t\t\t\t<Trdipole gauge=\"length\" unit=\"e*bohr\">" << b << "</Trdipole>\n";
        }
        f << "\t\t\t</level>\n";
      }
      f << "\t\t</exciton_uks>\n";
    }
    f << "\t</GWBSE>\n</output>\n";
  }

  // The GW-BSE part of Orbitals::WriteToCpt (orbitals.cc:990-1063): same group (/QMdata), names and HDF5 types for
  // everything this stage reads or produces, and the scalar / empty members of an Orbitals object that has no
  // unrestricted, embedding or localised-orbital data (defaults of orbitals.h).  NOT written: the compound tables of
  // the DFT side - qmmolecule (atoms), dft / aux (basis shells) - which belong to the caller's Orbitals object.
  // Orbitals::ReadFromCpt opens "qmmolecule" unconditionally (orbitals.cc:1093-1094), so this file alone is a results
  // dump (read it with any HDF5 tool, or merge it into the .orb of the DFT run); in an integrated build the
  // reference's own CheckpointWriter serialises the Orbitals object the shim fills.
  void WriteToCpt(const Results& r, const std::string& filename) const {
    CheckpointFile cpf(filename);
    CheckpointWriter w = cpf.getWriter("/QMdata");
    w(std::string("gwbse-b200"), "XTPVersion");
    w(int(9), "version");  // Orbitals::orbitals_version(), orbitals.h:844
    w(long(in_.homo + 1), "occupied_levels");
    w(long(in_.homo + 1), "occupied_levels_beta");
    w(long(in_.homo + 1), "number_alpha_electrons");
    w(long(in_.homo + 1), "number_beta_electrons");
    w(long(0), "charge");
    w(long(1), "spin");
    const MatrixXd none;
    const VectorXd nov;
    w.WriteEigenSystem(*in_.mo_energies, *in_.mos, none, 0, "mos");
    w.WriteEigenSystem(nov, none, none, 0, "mos_beta");
    w(nov, "occupations");
    w(long(0), "active_electrons");
    w.WriteEigenSystem(nov, none, none, 0, "mos_embedding");
    w(none, "LMOs");
    w(nov, "LMOs_energies");
    w(none, "inactivedensity");
    w(none, "TruncMOsFullBasis");
    w(0.0, "qm_energy");
    w(std::string("gwbse-b200"), "qm_package");
    w(std::string(""), "XCFunctional");
    w(std::string("medium"), "XC_grid_quality");
    w(std::string(""), "ECP");
    w(std::string("NoEmbedding"), "CalcType");
    w(long(r.rpamin), "rpamin");
    w(long(r.rpamax), "rpamax");
    w(long(r.qpmin), "qpmin");
    w(long(r.qpmax), "qpmax");
    w(long(r.bse_vmin), "bse_vmin");
    w(long(r.bse_cmax), "bse_cmax");
    w(in_.ScaHFX, "ScaHFX");
    w(bseopt_.useTDA, "useTDA");
    w(r.RPA_inputenergies, "RPA_inputenergies");
    w(r.QPpert_energies, "QPpert_energies");
    w.WriteEigenSystem(r.QPdiag_eigenvalues, r.QPdiag_eigenvectors, none, 0, "QPdiag");
    w.WriteEigenSystem(r.BSE_singlet.eigenvalues, r.BSE_singlet.eigenvectors, r.BSE_singlet.eigenvectors2,
                       r.BSE_singlet.eigenvalues.size() && !r.BSE_singlet.success ? 2 : 0, "BSE_singlet");
    w(r.transition_dipoles, "transition_dipoles");
    w.WriteEigenSystem(r.BSE_triplet.eigenvalues, r.BSE_triplet.eigenvectors, r.BSE_triplet.eigenvectors2,
                       r.BSE_triplet.eigenvalues.size() && !r.BSE_triplet.success ? 2 : 0, "BSE_triplet");
    w(std::uint8_t(bseopt_.use_Hqp_offdiag ? 1u : 0u), "use_Hqp_offdiag");
    w(std::uint8_t(r.is_qsgw ? 1u : 0u), "is_qsgw");
    w(r.BSE_singlet_dynamic, "BSE_singlet_dynamic");
    w(r.BSE_triplet_dynamic, "BSE_triplet_dynamic");
    // spin-resolved GW / BSE members (gw_uks, bse_uks): empty on the restricted path
    w(nov, "RPA_inputenergies_alpha");
    w(nov, "RPA_inputenergies_beta");
    w(nov, "QPpert_energies_alpha");
    w(nov, "QPpert_energies_beta");
    w.WriteEigenSystem(nov, none, none, 0, "QPdiag_alpha");
    w.WriteEigenSystem(nov, none, none, 0, "QPdiag_beta");
    w.WriteEigenSystem(nov, none, none, 0, "BSE_uks");
    w(nov, "BSE_uks_dynamic");
    cpf.Close();
  }

 private:
  const Device& dev_;
  Logger& log_;
  Inputs in_;
  GW::options gwopt_;
  std::vector<std::vector<Index>> fragments_;  // atom indices per fragment (bse.fragments.fragment.indices)
  std::string sigma_plot_states_, sigma_plot_filename_;
  Index sigma_plot_steps_ = 201;
  double sigma_plot_spacing_ = 1e-2;
  BSE::options bseopt_;
  MatrixXd qsgw_mos_;  // MO coefficients with the QP-window columns rotated to the QSGW wavefunctions
  bool do_gw_ = false, do_bse_singlets_ = false, do_bse_triplets_ = false, do_dynamical_screening_bse_ = false;
  bool do_bse_exciton_uks_ = false;
};

}  // namespace xtp
}  // namespace votca
