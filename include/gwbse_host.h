/* gwbse_host.h - C entry points of libgwbse_host.so, the C++ host layer above gwbse_b200.h.
 *
 * The host layer mirrors the reference's classes (headers under votca_b200/host: TCMatrix_gwbse, RPA, Sigma_*, GW,
 * BSE_OPERATOR, DavidsonSolver, BSE, GWBSE).  This "job" facade exposes GWBSE::Initialize/Evaluate
 * (xtp/src/libxtp/gwbse/gwbse.cc:60-578, 822-1250) to non-C++ callers (tests, bench.py):
 *   - options use the keys of share/xtp/xml/subpackages/gwbse.xml ("gw.mode", "bse.exctotal", ...) and can
 *     be loaded from the same options XML the reference takes with `xtp_tools -e dftgwbse -o opts.xml`;
 *   - inputs are what the reference reads from the Orbitals object (orbitals.cc:990-1063);
 *   - outputs carry the .orb dataset names (QPpert_energies, QPdiag_eigenvalues, BSE_singlet_eigenvalues ...).
 * All functions return 0 on success; gwbse_job_error() gives the message of the last failure.
 */
#ifndef GWBSE_HOST_H
#define GWBSE_HOST_H
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct gwbse_job gwbse_job;

int gwbse_job_create(int device, gwbse_job** out);
void gwbse_job_destroy(gwbse_job* job);
const char* gwbse_job_error(const gwbse_job* job);
const char* gwbse_job_create_error(void);
const char* gwbse_job_log(const gwbse_job* job);

/* multi-GPU: call before gwbse_job_run on every rank (id from gwbse_nccl_unique_id on rank 0) */
int gwbse_job_comm_init(gwbse_job* job, int rank, int world, const unsigned char* id128);

int gwbse_job_set_option(gwbse_job* job, const char* key, const char* value);
int gwbse_job_load_options_xml(gwbse_job* job, const char* path);
/* scalars: "homo", "ScaHFX" */
int gwbse_job_set_scalar(gwbse_job* job, const char* name, double value);
/* arrays (column-major, copied): "mos" (nbasis x nmo), "mo_energies" (nmo x 1), "vxc" (q x q),
 * "aux_overlap", "aux_coulomb" (naux x naux), "dipole_x|y|z" (ctotal x vtotal interlevel dipoles),
 * "Hqp", "RPA_inputenergies" (BSE-only runs); "ao_dipole_x|y|z" (N x N AO dipole matrices, AODipole::Fill: the
 * interlevel dipoles are then formed on the device, for both spin channels of an unrestricted reference);
 * "nuclear_charges" (natoms; ignore_corelevels and bse.fragments), "ao_overlap" (N x N) and "basis_atom_index" (N, the
 * atom of every basis function) for bse.fragments - the last two come from the dft basis when gwbse_job_set_basis
 * gave one.  Results of the optional analyses: "fragment_gs", "BSE_singlet|triplet_fragment_hole|electron"
 * (nfragments x nstates); unrestricted runs: "BSE_uks_dynamic", "uks_transition_dipoles", "uks_oscillator_strengths".
 * "ao3c": naux matrices N x N, rows = N*N, cols = naux; NOT copied, must stay alive until run returns. */
int gwbse_job_set_array(gwbse_job* job, const char* name, const double* data, long rows, long cols);
/* AO integral producer callback instead of "ao3c": fill(user, aux_offset, aux_count, out[aux_count*N*N]) */
typedef void (*gwbse_ao3c_fn)(void* user, long aux_offset, long aux_count, double* out);
int gwbse_job_set_ao3c_callback(gwbse_job* job, long nbasis, long naux, gwbse_ao3c_fn fn, void* user);

/* AO integrals already resident on the device (naux contiguous N x N matrices) */
int gwbse_job_set_ao3c_dev(gwbse_job* job, long nbasis, long naux, const double* ao3c_dev);
/* Multi-GPU jobs contract the aux functions rank by rank (gwbse_mmn_fill_begin/end): a rank only needs the
 * integrals of its share [gwbse_shard_aux_begin(naux, rank, world), gwbse_shard_aux_begin(naux, rank + 1, world)).
 * data holds aux functions [first_aux, first_aux + count) (host memory, or device memory with on_device = 1). */
int gwbse_job_set_ao3c_partial(gwbse_job* job, long nbasis, long naux, long first_aux, long count, const double* data,
                               int on_device);
/* AO integrals produced on the GPU instead of "ao3c" (gwbse_b200.h: gwbse_basis_create, gwbse_ao3c_block_dev,
 * gwbse_ao_coulomb2c - the device stand-ins for ComputeAO3cBlock / AOCoulomb::Fill, libint2_calls.cc:544-593,
 * 224-271): which = "dft" or "aux", arguments as gwbse_basis_create.  Used when both are set and no ao3c array or
 * callback is; "aux_overlap" and "aux_coulomb" then become optional (computed on the device: gwbse_ao_overlap,
 * gwbse_ao_coulomb2c), and so do "dipole_x|y|z" (AO dipoles from gwbse_ao_dipole, interlevel dipoles formed as
 * Orbitals::CalcFreeTransition_Dipoles does). */
int gwbse_job_set_basis(gwbse_job* job, const char* which, int nshell, const int* l, const int* nprim,
                        const double* centers, const double* exps, const double* coefs);
/* Results as an .orb checkpoint: gwbse_job_run (rank 0) writes /QMdata with the names and HDF5 types of
 * Orbitals::WriteToCpt (orbitals.cc:990-1063) for everything this stage reads or produces - mos, the level ranges,
 * RPA_inputenergies, QPpert_energies, QPdiag, BSE_singlet / BSE_triplet (eigenvalues, eigenvectors, eigenvectors2,
 * info), transition_dipoles (ind0..), BSE_*_dynamic, useTDA, use_Hqp_offdiag, ScaHFX.  Written without an HDF5
 * library (votca_b200/host/checkpoint.h).  NULL or "" switches it off. */
int gwbse_job_set_orb_output(gwbse_job* job, const char* path);
/* GWBSE::addoutput (gwbse.cc:580-738) written as the dftgwbse tool does (tools/dftgwbse.cc:120-128,
 * <job>_summary.xml): DFT / GW / QP level energies, singlet and triplet excitation energies, oscillator strengths
 * and transition dipoles, eV; gwbse_job_run_uks writes the dft_alpha / dft_beta tables and the exciton_uks levels.
 * The input scalar "dft_total_energy" (Hartree) fills the DFTEnergy attribute.                                   */
int gwbse_job_set_summary_output(gwbse_job* job, const char* path);
/* the kernel-library context of this job (gwbse_b200.h), e.g. for gwbse_gemm_stats / timers */
void* gwbse_job_ctx(gwbse_job* job);

int gwbse_job_run(gwbse_job* job);

/* BSECoupling::CalculateCouplings + Addoutput (xtp/src/libxtp/bsecoupling.cc:356-612, 127-254; the calculator behind
 * `xtp_tools -e bsecoupling` / the iexcitoncl job type): exciton couplings between monomers A and B from the GW-BSE
 * results of the dimer.  Dimer inputs are the job's own ("mos", "Hqp" = QPdiag eigenvectors * eigenvalues *
 * eigenvectors^T, "RPA_inputenergies", AO integrals as arrays or basis sets, "dft_overlap" unless a dft basis is set,
 * scalars homo, rpamin, rpamax, qpmin, qpmax, bse_vmin, bse_cmax, use_Hqp_offdiag); monomer X in {A, B}: arrays
 * "X.mos", "X.BSE_singlet_eigenvectors" / "X.BSE_triplet_eigenvectors" (+ "_eigenvalues"), scalars "X.bse_vmin",
 * "X.bse_vmax", "X.bse_cmin", "X.bse_cmax".  Options: gwbse_job_set_option with the keys of bsecoupling.xml prefixed
 * "bsecoupling." (spin, use_perturbation, output_tb, moleculeA.states, moleculeA.occLevels, ...).
 * Outputs: arrays JAB_{singlet,triplet}_{pert,diag} (Hartree, (levA+levB)^2), J_dimer_*, S_dimer_*; scalars xi_*,
 * pt_rm_discrepancy_*, downfolding_safe_*, levA, levB; gwbse_job_coupling_xml = the <bsecoupling> output subtree. */
int gwbse_job_run_coupling(gwbse_job* job);

/* GWBSE::Evaluate for an unrestricted reference (the is_uks branches of gwbse.cc:896-997, 1150-1190): GW_UKS
 * (xtp/src/libxtp/gwbse/gw_uks.cc: spin-summed RPA screening rpa_uks.cc, one shared plasmon-pole model, per-spin
 * Sigma_x / Sigma_c / QP search) and, for the task "exciton_uks" ("excitons"), the combined exciton problem of BSE_UKS
 * (bse_uks.cc, bse_operator_uks.cc).  Alpha channel: the restricted inputs; beta channel: arrays "mos_beta",
 * "mo_energies_beta", "vxc_beta", scalar "homo_beta".  Outputs: the restricted names with "_alpha" / "_beta"
 * appended (RPA_inputenergies, QPpert_energies, QPdiag_eigenvalues / _eigenvectors, Hqp, Sigma_x, Sigma_c),
 * "BSE_uks_eigenvalues" / "BSE_uks_eigenvectors" (rows: alpha (v c) then beta (v c)), scalars bse_alpha_size,
 * bse_beta_size, gw_iterations, uks_converged; "BSE_uks_eigenvectors2" (Y) for bse.useTDA=false.
 * Scope: sigma_integrator=ppm, one GPU.                                                                            */
int gwbse_job_run_uks(gwbse_job* job);
const char* gwbse_job_coupling_xml(const gwbse_job* job);

int gwbse_job_array_dims(const gwbse_job* job, const char* name, long* rows, long* cols);
int gwbse_job_get_array(const gwbse_job* job, const char* name, double* out);
int gwbse_job_get_scalar(const gwbse_job* job, const char* name, double* out);
long long gwbse_job_launch_count(const gwbse_job* job);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GWBSE_HOST_H */
