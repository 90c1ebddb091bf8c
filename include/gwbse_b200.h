/* gwbse_b200.h - C ABI of libgwbse_b200.so
 *
 * B200-native (sm_100a) implementation of the dense FP64 contraction path
 * behind VOTCA-XTP's `xtp_tools -e dftgwbse`.  This is the boundary a
 * maintainer binds to from the reference's C++ classes; every entry point
 * names the reference interface it replaces (paths relative to the votca
 * repository root).  INTEGRATION.md shows the reference-side shim.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure;
 *    gwbse_last_error(ctx) returns the message (the C++ shim rethrows it as
 *    std::runtime_error, the reference's only error convention,
 *    xtp/src/libxtp/cudamatrix.cc:25-37).
 *  - pointers are HOST pointers unless the parameter name ends in `_dev`.
 *  - host matrices are column-major with explicit leading dimension (Eigen
 *    MatrixXd layout).
 *  - one context per process/GPU; calls on one context are serialised by the
 *    caller (the reference's OpenMP_CUDA has the same rule,
 *    xtp/include/votca/xtp/openmp_cuda.h:49-65).
 *  - there is no CPU fallback: without a CUDA device gwbse_ctx_create fails.
 *
 * Device layout of Mmn ("aux-major"):  element (m, n, chi) of the reference's
 * matrix_[m](n, chi) (xtp/include/votca/xtp/threecenter.h:125-134) lives at
 *     X[chi * ldx + mloc * npad + n],  ldx = mlocal * npad,
 * i.e. one (mlocal*npad) x naux column-major matrix.  mloc is the index of m
 * among the levels owned by this rank (m-cyclic sharding: owner = m % world).
 */
#ifndef GWBSE_B200_H
#define GWBSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct gwbse_ctx gwbse_ctx;

/* ---- context ----------------------------------------------------------- */
/* replaces OpenMP_CUDA::OpenMP_CUDA / CudaPipeline ctor (openmp_cuda.cc:48-71,
 * cudapipeline.h:89-104): binds one GPU, creates stream + solver handles.   */
int gwbse_ctx_create(int device, gwbse_ctx** out);
void gwbse_ctx_destroy(gwbse_ctx* ctx);
const char* gwbse_last_error(const gwbse_ctx* ctx);
const char* gwbse_create_error(void); /* message of a failed gwbse_ctx_create */
int gwbse_sync(gwbse_ctx* ctx);
/* number of kernels of this library launched so far on ctx (bench.py gpu_launches) */
long long gwbse_launch_count(const gwbse_ctx* ctx);
int gwbse_profile_report(gwbse_ctx* ctx, char* buf, size_t buflen); /* also resets the counters */
int gwbse_device_count(void); /* OpenMP_CUDA::AvailableGPUs, openmp_cuda.cc:30-46 */
/* tuning knobs: "bse_chunk_bytes" (size of the Hd/Hd2 intermediate held at once), "profile" (0/1: time every
 * entry point with CUDA events on the context's stream; read the table with gwbse_profile_report),
 * "sigma_tree_min_terms" (Sigma_c diagonal elements use the treecode over sorted pole positions when
 * n * npoles reaches this, the term-by-term kernel below it; default 32768), "sigma_tree_bytes" (budget of
 * the per-level moment store of the treecode; default 8 GiB) */
int gwbse_set_option(gwbse_ctx* ctx, const char* key, double value);
/* Per-kernel accounting of the DMMA GEMM (bench.py roofline): when enabled every GEMM launch is bracketed
 * by CUDA events on the context's stream; stats = summed kernel milliseconds, algorithmic flops, launches. */
int gwbse_gemm_profile(gwbse_ctx* ctx, int enable);
int gwbse_gemm_stats(gwbse_ctx* ctx, double* ms, double* flops, long long* launches);
/* the same accounting broken down by GEMM shape / tile configuration, as a text table (most expensive first) */
int gwbse_gemm_shape_report(gwbse_ctx* ctx, char* buf, size_t buflen);
/* Live FP64 tensor (DMMA) issue-rate probe: register-resident mma.sync loop on every SM -> TFLOP/s.
 * This is the roofline denominator for the contraction kernels (MEASURED_PEAKS.json has no FP64 entry). */
int gwbse_fp64_peak_probe(gwbse_ctx* ctx, double* tflops);
/* CUDA-event timers on the context's stream (bench.py times kernels with these) */
int gwbse_timer_start(gwbse_ctx* ctx);
int gwbse_timer_stop_ms(gwbse_ctx* ctx, float* ms);

/* ---- multi-GPU (one process per GPU, NCCL over NVLink) ------------------ */
/* m-cyclic sharding of Mmn; replaces the thread-per-GPU loop + host reduction
 * of OpenMP_CUDA::getReductionVar (openmp_cuda.cc:480-493).                 */
int gwbse_nccl_unique_id(unsigned char* id128);
int gwbse_comm_init(gwbse_ctx* ctx, int rank, int world, const unsigned char* id128);
int gwbse_comm_rank(const gwbse_ctx* ctx);
int gwbse_comm_world(const gwbse_ctx* ctx);
int gwbse_comm_allreduce_host(gwbse_ctx* ctx, double* buf, size_t n); /* sum, in place */
/* the m-cyclic sharding plan (pure functions, usable without a GPU) */
int gwbse_shard_owner(int m, int world);
int gwbse_shard_local_index(int m, int world);
int gwbse_shard_local_count(int total, int rank, int world);

/* ---- raw device memory (CudaMatrix, cudamatrix.h:95-190) ---------------- */
int gwbse_dev_malloc(gwbse_ctx* ctx, size_t bytes, double** out_dev);
int gwbse_dev_free(gwbse_ctx* ctx, double* p_dev);
int gwbse_h2d(gwbse_ctx* ctx, double* dst_dev, const double* src, size_t n);
int gwbse_d2h(gwbse_ctx* ctx, double* dst, const double* src_dev, size_t n);
int gwbse_d2d(gwbse_ctx* ctx, double* dst_dev, const double* src_dev, size_t n);
int gwbse_dev_memset_zero(gwbse_ctx* ctx, double* dst_dev, size_t n);
int gwbse_dev_mem_info(gwbse_ctx* ctx, size_t* free_bytes, size_t* total_bytes);

/* ---- dense primitives on device matrices -------------------------------- */
/* CudaPipeline::gemm (cudapipeline.h:106-131): C = alpha op(A) op(B) + beta C,
 * column-major, transa/transb in {'N','T'}; hand-written DMMA kernel.       */
int gwbse_dgemm_dev(gwbse_ctx* ctx, char transa, char transb, int m, int n, int k, double alpha,
                    const double* A_dev, int lda, const double* B_dev, int ldb, double beta, double* C_dev,
                    int ldc);
/* same with tile-shape / split-K override (tests, tuning): cfg 0=128x128, 1=128x32, 2=64x64, -1 auto */
int gwbse_dgemm_dev_ex(gwbse_ctx* ctx, char transa, char transb, int m, int n, int k, double alpha,
                       const double* A_dev, int lda, const double* B_dev, int ldb, double beta, double* C_dev,
                       int ldc, int cfg, int splitk);
/* CudaPipeline::diag_gemm (cudapipeline.h:133-162): C = A diag(d) (side 'R') or diag(d) A (side 'L') */
int gwbse_diag_scale_dev(gwbse_ctx* ctx, char side, int m, int n, const double* A_dev, int lda,
                         const double* d_dev, double* C_dev, int ldc);
/* CudaPipeline::axpy (cudapipeline.cc:36-51): Y += alpha X over an m x n block */
int gwbse_axpy_dev(gwbse_ctx* ctx, int m, int n, double alpha, const double* X_dev, int ldx, double* Y_dev,
                   int ldy);
/* column 2-norms of an m x n device matrix -> host (DavidsonSolver residual norms) */
int gwbse_colnorms_dev(gwbse_ctx* ctx, int m, int n, const double* A_dev, int lda, double* norms);
/* A(:,j) *= s[j]  (s on host) */
int gwbse_scale_cols_dev(gwbse_ctx* ctx, int m, int n, double* A_dev, int lda, const double* s);
/* column-wise dot products d[j] = X(:,j) . Y(:,j) -> host (BSE::ExpectationValue, bse.cc:489-498) */
int gwbse_coldots_dev(gwbse_ctx* ctx, int m, int n, const double* X_dev, int ldx, const double* Y_dev, int ldy,
                      double* dots);

/* symmetric eigen-decomposition, A (n x n, device, lower used) -> eigenvectors in place, w on host.
 * Eigen::SelfAdjointEigenSolver call sites: ppm.cc:37, bse.cc:194, aomatrix.cc:56,73, rpa.cc:328-345.
 * Non-contraction auxiliary: cuSOLVER Dsyevd.                                */
int gwbse_sym_eig_dev(gwbse_ctx* ctx, int n, double* A_dev, int lda, double* w);
/* general inverse (MatrixXd::inverse(): ppm.cc:44, sigma_cda.cc:43, ImaginaryAxisIntegration.cc:97):
 * LU with partial pivoting, result overwrites A.  cuSOLVER getrf/getrs.     */
int gwbse_inverse_dev(gwbse_ctx* ctx, int n, double* A_dev, int lda);
/* solve A x = b for nrhs right-hand sides (partialPivLu().solve, sigma_cda.cc:57-60); A destroyed */
int gwbse_lu_solve_dev(gwbse_ctx* ctx, int n, int nrhs, double* A_dev, int lda, double* B_dev, int ldb);
/* real non-symmetric generalized problem T x = lambda B x (Eigen::GeneralizedEigenSolver,
 * davidsonsolver.cc:252): small, host in/out.  wr/wi eigenvalues, VR right eigenvectors as LAPACK dgeev. */
int gwbse_gen_eig_host(gwbse_ctx* ctx, int n, const double* T, const double* B, double* wr, double* wi,
                       double* VR);

/* ---- Mmn: TCMatrix_gwbse (threecenter.h:41-142) ------------------------- */
/* TCMatrix_gwbse::Initialize (threecenter.cc:29-48) */
int gwbse_mmn_alloc(gwbse_ctx* ctx, int naux, int mmin, int mmax, int nmin, int nmax);
int gwbse_mmn_free(gwbse_ctx* ctx);
/* geometry queries (auxsize/msize/nsize + device layout) */
int gwbse_mmn_dims(const gwbse_ctx* ctx, int* naux, int* mtotal, int* ntotal, int* mlocal, int* npad);
/* TCMatrix_gwbse::Fill3cMO for a block of auxiliary functions
 * (libint2_calls.cc:595-651 + OpenMP_CUDA::MultiplyLeftRight, openmp_cuda.cc:172-192).
 * ao3c: aux_count symmetric N x N matrices (column-major, contiguous) = the
 * output of ComputeAO3cBlock (libint2_calls.cc:544-593); mos: N x nmo
 * coefficient matrix (MOs().eigenvectors()).  Contracts
 * M[m](n, aux_offset+k) = sum_{mu,nu} C(mu,nmin+n) ao3c[k](mu,nu) C(nu,mmin+m). */
int gwbse_mmn_set_mos(gwbse_ctx* ctx, const double* mos, int ldmos, int nbasis, int nmo);
/* Host variant: the block is streamed through two alternating device staging buffers on a copy stream, so
 * the transfer of one sub-block overlaps the contraction of the previous one; it returns when the last copy has
 * left the host buffer (the GEMMs keep running on the context's stream).  Full PCIe rate needs page-locked
 * memory: gwbse_host_malloc / gwbse_host_free hand out such buffers for the integral producer to write into
 * (the role of CudaMatrix's staging copies, cudamatrix.cc:60-95). */
int gwbse_mmn_fill_block(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c);
int gwbse_host_malloc(size_t bytes, void** out);
int gwbse_host_free(void* p);
int gwbse_mmn_fill_block_dev(gwbse_ctx* ctx, int aux_offset, int aux_count, const double* ao3c_dev);
/* Multi-GPU fill sharded over aux functions (the reference's own parallel loop, libint2_calls.cc:621-622):
 * between fill_begin(ctx, 1) and fill_end every rank passes only the aux functions of its share
 * [begin, end) = gwbse_shard_aux_range(ctx, rank) to gwbse_mmn_fill_block(_dev); they are contracted for all m
 * and fill_end (collective) moves them into the m-sharded tensor with one all-to-all over NVLink.  Each rank
 * then needs 1/world of the AO integrals on its host and over its PCIe link.  fill_begin(ctx, 0), or no call
 * at all, keeps the replicated mode: every rank passes every aux block and keeps its own m slices. */
int gwbse_shard_aux_range(const gwbse_ctx* ctx, int rank, int* begin, int* end);
int gwbse_shard_aux_begin(int naux, int rank, int world); /* share of rank r: [begin(r), begin(r+1)) */
int gwbse_mmn_fill_begin(gwbse_ctx* ctx, int aux_sharded);
int gwbse_mmn_fill_end(gwbse_ctx* ctx);
/* ---- AO Coulomb integrals on the device ---------------- */
/* A basis as AOBasis::Fill lays it out (xtp/src/libxtp/aobasis.cc:85-105): shells in atom order, functions
 * appended shell by shell, pure (real solid harmonic) functions m = -l..l (aoshell.cc:65-79).  l[s] in 0..6,
 * nprim[s] primitives per shell, centers: 3 doubles per shell (bohr); exps / coefs run over all primitives in
 * shell order.  coefs must already contain libint's primitive normalisation and VOTCA's shell norm
 * (AOShell::normalizeContraction, aoshell.cc:81-89) - i.e. the numbers the libint2::Shell holds. */
typedef struct gwbse_basis gwbse_basis;
/* host helper (no device work): raw contraction factors of a basis-set file -> the coefs gwbse_basis_create
 * expects, i.e. AOShell::LibintShell + AOShell::normalizeContraction (aoshell.cc:65-89) */
int gwbse_basis_normalize(int nshell, const int* l, const int* nprim, const double* exps, const double* contractions,
                          double* coefs_out);
int gwbse_basis_create(gwbse_ctx* ctx, int nshell, const int* l, const int* nprim, const double* centers,
                       const double* exps, const double* coefs, gwbse_basis** out);
int gwbse_basis_destroy(gwbse_ctx* ctx, gwbse_basis* basis);
int gwbse_basis_size(const gwbse_basis* basis); /* number of functions */
/* ComputeAO3cBlock (libint2_calls.cc:544-593; libint2 Operator::coulomb, BraKet::xs_xx) for the aux FUNCTIONS
 * [aux_offset, aux_offset + aux_count): aux_count symmetric N x N matrices (P | mu nu), column-major,
 * contiguous - exactly what gwbse_mmn_fill_block(_dev) consumes.  McMurchie-Davidson on the GPU, one warp per
 * (orbital shell pair, aux shell).  The range need not respect shell boundaries. */
int gwbse_ao3c_block_dev(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_offset,
                         int aux_count, double* out_dev);
int gwbse_ao3c_block(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_offset, int aux_count,
                     double* out);
/* AOCoulomb::Fill (libint2_calls.cc:224-271; BraKet::xs_xs): V(P, Q) = (P | Q), naux x naux to the host */
int gwbse_ao_coulomb2c(gwbse_ctx* ctx, const gwbse_basis* aux, double* V, int ld);
/* AOOverlap::Fill (libint2_calls.cc:163-165): S(mu, nu) = <mu | nu>, n x n to the host */
int gwbse_ao_overlap(gwbse_ctx* ctx, const gwbse_basis* basis, double* S, int ld);
/* AODipole::Fill (libint2 Operator::emultipole1; input of Orbitals::CalcFreeTransition_Dipoles,
 * orbitals.cc:742-760): D[k](mu, nu) = <mu | r_k | nu> about the origin, k = x, y, z: three n x n matrices (ld x n
 * doubles each, one after the other) to the host */
int gwbse_ao_dipole(gwbse_ctx* ctx, const gwbse_basis* basis, double* D, int ld);
/* TCMatrix_gwbse::Fill3cMO (libint2_calls.cc:595-651) with the integral producer on the GPU: blocks of aux_block
 * aux functions are computed into a device buffer and contracted from there (gwbse_mmn_fill_block_dev); no AO
 * integral crosses PCIe.  Needs gwbse_mmn_alloc + gwbse_mmn_set_mos.  Multi-GPU: every rank produces and
 * contracts its own share of the aux functions, then the all-to-all of gwbse_mmn_fill_end (collective). */
int gwbse_mmn_fill_from_basis(gwbse_ctx* ctx, const gwbse_basis* aux, const gwbse_basis* dft, int aux_block);
/* TCMatrix_gwbse::MultiplyRightWithAuxMatrix (threecenter.cc:54-65,
 * OpenMP_CUDA::MultiplyRight openmp_cuda.cc:131-150): M[m] <- M[m] * R      */
int gwbse_mmn_mul_right(gwbse_ctx* ctx, const double* R, int ldr);
int gwbse_mmn_mul_right_dev(gwbse_ctx* ctx, const double* R_dev, int ldr);
/* The same product when the caller will next read only the rows n in [n_lo, n_hi) of every slice - the BSE operator
 * after BSE::SetupDirectInteractionOperator (bse.cc:188-204) reads n in the (v, c) window only, a third of the rows
 * of a full-range tensor.  Those rows are rotated now; the others keep the rotation pending and get it, transparently,
 * when any entry point that may read them runs (every one except gwbse_bse_* on a covered window).  Results are those
 * of gwbse_mmn_mul_right_dev.  R_dev is copied.                                                                     */
int gwbse_mmn_mul_right_window_dev(gwbse_ctx* ctx, const double* R_dev, int ldr, int n_lo, int n_hi);
/* TCMatrix_gwbse::Rotate (threecenter.cc:108-131, QSGW): for the m slices of the QP window,
 * M[m].middleRows(qpmin - nmin, q) <- U^T * M[m].middleRows(qpmin - nmin, q); U is q x q, q = qpmax - qpmin + 1 */
int gwbse_mmn_rotate(gwbse_ctx* ctx, const double* U, int ldu, int qpmin, int qpmax);
/* AOCoulomb::Pseudo_InvSqrt_GWBSE (aomatrix.cc:53-86) on the device; S, V: naux x naux host.  L_out may be
 * NULL: the result then only stays on the device (gwbse_pseudo_invsqrt_result_dev, valid until the next call)
 * for gwbse_mmn_mul_right_dev, which is how TCMatrix_gwbse::Fill uses it (threecenter.cc:72-90). */
int gwbse_pseudo_invsqrt(gwbse_ctx* ctx, int naux, const double* S, const double* V, double etol, double* L_out,
                         int* removed);
const double* gwbse_pseudo_invsqrt_result_dev(gwbse_ctx* ctx);
/* operator[] (threecenter.h:125-134): copy slice M[m] (ntotal x naux, col-major, ld) to/from host.
 * m is the global storage index (m_abs - mmin); must be owned by this rank. */
int gwbse_mmn_get_slice(gwbse_ctx* ctx, int m, double* out, int ld);
int gwbse_mmn_set_slice(gwbse_ctx* ctx, int m, const double* in, int ld);
/* TCMatrix_gwbse::Rebuild support (gw.cc:242-246): keep / restore a pristine copy on the device */
int gwbse_mmn_snapshot(gwbse_ctx* ctx);
int gwbse_mmn_restore(gwbse_ctx* ctx);

/* ---- RPA (rpa.h:35-115) -------------------------------------------------- */
/* RPA::calculate_epsilon_i / _r(double) / _r(complex) (rpa.cc:75-202 +
 * OpenMP_CUDA::A_TDA openmp_cuda.cc:218-256): kind 0 = imaginary axis,
 * 1 = real axis, 2 = complex.  energies: RPA input energies (rpatotal).
 * Result (naux x naux) stays on the device (gwbse_rpa_epsilon_ptr) and is
 * copied to eps_out if non-NULL.  Multi-GPU: partials are all-reduced.     */
int gwbse_rpa_epsilon(gwbse_ctx* ctx, int kind, double freq_re, double freq_im, double eta,
                      const double* energies, int homo, int rpamin, int rpamax, double* eps_out, int ld);
double* gwbse_rpa_epsilon_ptr(gwbse_ctx* ctx);
/* RPA::setQSGWRotation (rpa.h:59-66) and Sigma_base::setQSGWRotation (sigma_base.h:51-58): while a rotation U
 * (qptotal x qptotal, DFT-MOs -> QP wavefunctions) is registered, every RPA sum over hole slices - epsilon
 * (rpa.cc:95-118, 162-181), A+B (rpa.cc:288-310), the exact-sigma residues (sigma_exact.cc:119-145) - uses
 * sum_vp U(vp, v) Mmn[vp + qpmin - rpamin] for the occupied slices v inside the QP window.  U = NULL clears it. */
int gwbse_rpa_set_qsgw_rotation(gwbse_ctx* ctx, const double* U, int ldu, int qptotal, int qpmin, int homo);
/* RPA::Calculate_H2p_ApB (rpa.cc:281-326): (A+B) two-particle matrix, S x S, lower triangle */
int gwbse_rpa_h2p_apb(gwbse_ctx* ctx, const double* energies, int homo, int rpamin, int rpamax,
                      double* apb_out_dev, int ld);
/* One block of the unrestricted A+B matrix, RPA_UKS::Calculate_H2p_ApB (rpa_uks.cc:475-540):
 *   block[(v1,c1), (v2,c2)] = alpha sum_chi M_ctx[v1][c1,chi] M_other[v2][c2,chi]
 * rows: particle-hole pairs of `ctx` (homo), columns: those of `other` (homo_other), which holds the other spin
 * channel's Mmn on the same GPU, idle during the call (it may be `ctx` itself).  diag(AmB) is added by the caller. */
int gwbse_rpa_h2p_block(gwbse_ctx* ctx, gwbse_ctx* other, int homo, int homo_other, int rpamin, int rpamax,
                        double alpha, double* block_dev, int ld);

/* ---- Sigma (sigma_base.h:33-106) ---------------------------------------- */
/* Sigma_base::CalcExchangeMatrix (sigma_base.cc:36-52): q x q, host out */
int gwbse_sigma_x(gwbse_ctx* ctx, int homo, int rpamin, int qpmin, int qpmax, double* out, int ld);
/* Sigma_PPM::CalcCorrelationDiagElement / ...Derivative (sigma_ppm.cc:37-91), batched over
 * nreq (level, frequency) requests.  weights/freqs: PPM parameters (naux), energies: RPA input.
 * dsigma may be NULL.  lumo_abs = homo + 1 exactly as the reference uses it. */
int gwbse_sigma_ppm_set(gwbse_ctx* ctx, const double* ppm_weight, const double* ppm_freq, const double* energies,
                        int homo, int rpamin, int qpmin, double eta);
/* The evaluators read rpa_.getRPAInputEnergies() at evaluation time (sigma_ppm.cc:53, sigma_exact.cc:51);
 * evGW changes them between PrepareScreening and the final CalcCorrelationDiag (gw.cc:258-308), so the host
 * refreshes the device copy before evaluating.  which: 0 = ppm, 1 = exact.                                */
int gwbse_sigma_update_energies(gwbse_ctx* ctx, int which, const double* energies);
int gwbse_sigma_ppm_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma,
                         double* dsigma);
/* Grouped form of the two evaluators above (which: 0 = ppm, 1 = exact): group g evaluates level levels[g]
 * at the frequencies freqs[group_ptr[g] .. group_ptr[g+1]); the level's data is streamed once per group.
 * This is what the QP root search uses: its scan / sub-step / bisection frequencies are known ahead
 * (qp_solver_utils.h:484-628) and are evaluated together.                                              */
int gwbse_sigma_eval_groups(gwbse_ctx* ctx, int which, int ngroups, const int* levels, const int* group_ptr,
                            const double* freqs, double* sigma, double* dsigma);
/* Sigma_PPM::CalcCorrelationOffDiagElement for all pairs (sigma_ppm.cc:93-126 via
 * Sigma_base::CalcCorrelationOffDiag sigma_base.cc:65-78): q x q symmetric, zero diagonal */
int gwbse_sigma_ppm_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld);
/* Sigma_Exact (sigma_exact.cc:29-148): residues from XpY (S x S device), then batched evaluation */
int gwbse_sigma_exact_prepare(gwbse_ctx* ctx, const double* rpa_omegas, const double* XpY_dev, int ldxpy,
                              const double* energies, int homo, int rpamin, int rpamax, int qpmin, int qpmax,
                              double eta);
/* The two halves of the call above on their own, for the unrestricted evaluator (Sigma_Exact_UKS::PrepareScreening,
 * sigma_exact_uks.cc:37-61, with the screening modes of RPA_UKS::BuildCachedScreeningModes, rpa_uks.cc:91-161):
 * project: Z[chi, s] (+)= sum_{v,c} M[v][c,chi] XpY[(v,c), s] for the channel held by ctx (XpY_dev points at the
 *          channel's first row of the combined eigenvectors; accumulate != 0 adds to Z, which may live in the other
 *          channel's context on the same GPU);
 * prepare_modes: residues M_i Z of the qp window for nmodes screening modes (columns of Z) with frequencies omegas;
 *          diag_pref / offdiag_pref multiply the pole sums of the diagonal / off-diagonal elements (closed shell:
 *          2 and 1, sigma_exact.cc:57,106; per spin channel: 1 and 0.5, sigma_exact_uks.cc:82,140).               */
int gwbse_sigma_exact_project(gwbse_ctx* ctx, const double* XpY_dev, int ldxpy, int ncols, int homo, int rpamin,
                              int rpamax, int accumulate, double* Z_dev, int ldz);
int gwbse_sigma_exact_prepare_modes(gwbse_ctx* ctx, const double* omegas, int nmodes, const double* Z_dev, int ldz,
                                    const double* energies, int homo, int rpamin, int rpamax, int qpmin, int qpmax,
                                    double eta, double diag_pref, double offdiag_pref);
int gwbse_sigma_exact_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, double* sigma,
                           double* dsigma);
int gwbse_sigma_exact_offdiag(gwbse_ctx* ctx, int q, const double* freqs, double* out, int ld);

/* ---- Sigma_CDA: contour deformation (device resident) ---------------------
 * Sigma_CDA::PrepareScreening (self_energy_evaluators/sigma_cda.cc:30-45) +
 * ImaginaryAxisIntegration::CalcDielInvVector (ImaginaryAxisIntegration.cc:90-102):
 * kappa_0 = eps(0)^-1 - 1, kappa_j = -(eps(i w_j)^-1 - 1) + kappa_0 exp(-(alpha w_j)^2)
 * for the `order` mapped quadrature nodes (points / weights, host), and the row
 * forms Q_j[level][n] = (I kappa_j)[n,:] . I[n,:], I = Mmn[level], for every local
 * level of the qp window (one DMMA GEMM per node over the whole Mmn).  eta is
 * RPA::getEta() (rpa.h:90).  `energies`: RPA input energies (rpamax-rpamin+1).  */
int gwbse_sigma_cda_prepare(gwbse_ctx* ctx, int order, const double* points, const double* weights, int symmetry,
                            double alpha, const double* energies, int homo, int rpamin, int rpamax, int qpmin,
                            int qpmax, double eta);
/* Sigma_CDA::CalcCorrelationDiagElement (sigma_cda.cc:118-124) for nreq (level,
 * frequency) requests: SigmaGQDiag (ImaginaryAxisIntegration.cc:145-176) and the
 * Gaussian tail from the Q tables, CalcResidueContribution (sigma_cda.cc:81-116)
 * with one eps(|e_n - w| + i eta) assembly + LU solve per pole inside the contour.
 * levels are relative to qpmin and must be owned by this rank.                    */
int gwbse_sigma_cda_eval(gwbse_ctx* ctx, int nreq, const int* levels, const double* freqs, const double* energies,
                         double* sigma);
/* Sigma_CDA_UKS (self_energy_evaluators/sigma_cda_uks.cc:44-159): the restricted formulas with the dielectric matrix
 * of RPA_UKS (rpa_uks.cc:203-367, both spin channels) and the channel's own energies / Mmn.  While a partner is
 * registered, every eps(z) the two calls above assemble is the mean of this context's and the partner's matrices
 * (their weights carry the closed-shell factor 2).  energies_other: the partner's RPA input energies (ntotal),
 * to be refreshed whenever they change; other = NULL returns to the restricted evaluator.  The partner context lives
 * on the same GPU and is idle during the calls.                                                                    */
int gwbse_sigma_cda_set_partner(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, const double* energies_other);

/* ---- BSE operator (bse_operator.h:32-87) --------------------------------- */
/* BSE_OPERATOR::configure + ctor data (bse_operator.cc:29-38): eps_inv (naux), Hqp ((vt+ct)^2, ld) */
int gwbse_bse_configure(gwbse_ctx* ctx, int homo, int rpamin, int vmin, int cmax, const double* eps_inv,
                        const double* Hqp, int ldh);
/* BSE_OPERATOR<cqp,cx,cd,cd2>::matmul (bse_operator.cc:40-119): Y = H X, X,Y: bse_size x k */
int gwbse_bse_matmul(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X, int ldx, double* Y,
                     int ldy);
int gwbse_bse_matmul_dev(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, int k, const double* X_dev, int ldx,
                         double* Y_dev, int ldy);
/* The two halves of the exchange term as separate calls - what BSE_OPERATOR_UKS needs for the coupling between the
 * spin channels (add_direct_cross_tda_block, bse_operator_uks.cc:174-211: transition densities of the input channel
 * contracted with the trial vectors, screened, expanded in the transition densities of the output channel), each
 * half on the context that holds that channel's Mmn:
 *   project: W[chi, kv]     = sum_{v,c} M[v][c, chi] X[(v,c), kv]                       W_dev: naux x k, ld naux
 *   expand : Y[(v,c), kv] += alpha sum_chi M[v][c, chi] (screened ? eps_inv[chi] : 1) W[chi, kv]
 * with the level ranges / eps_inv of the last gwbse_bse_configure of that context.                                */
int gwbse_bse_vc_project_dev(gwbse_ctx* ctx, int k, const double* X_dev, int ldx, double* W_dev);
int gwbse_bse_vc_expand_dev(gwbse_ctx* ctx, double alpha, int screened, int k, const double* W_dev, double* Y_dev,
                            int ldy);
/* Cross-spin block of the unrestricted full-BSE B operator (BSE_OPERATOR_UKS::add_direct2_block with different
 * output and input channels, bse_operator_uks.cc:136-172, 252-255):
 *   Y[(v1,c1), kv] += alpha sum_{v2,c2,chi} M_ctx[c1][v2,chi] eps_inv[chi] M_other[v1][c2,chi] X[(v2,c2), kv]
 * (v1, c1) in the ranges of `ctx` (its last gwbse_bse_configure), (v2, c2) in those of the other channel
 * (homo_other with the same vmin / cmax).  `other` is the context holding the other channel's Mmn on the same GPU;
 * it must be idle (synchronised) during the call.                                                               */
int gwbse_bse_hd2_cross_dev(gwbse_ctx* ctx, gwbse_ctx* other, int homo_other, double alpha, int k,
                            const double* X_dev, int ldx, double* Y_dev, int ldy);
/* accounting for bench.py: algorithmic flops (F_bse), operator products and trial columns applied
 * through gwbse_bse_matmul(_dev) since the last reset */
int gwbse_bse_stats(gwbse_ctx* ctx, double* algo_flops, long long* products, long long* columns, int reset);
/* The direct terms Hd / Hd2 of the operator (bse_operator.cc:61-116 rebuilds their rows for every product) are
 * applied in factorised form until materialising the B x B block pays back, then from the resident block (option
 * "bse_dense": 0 never, 1 when it pays back - default, 2 always; "bse_dense_payback": fraction of a build the
 * factorised work under one (Mmn, screening, window) has to reach).  builds: blocks formed so far; columns: trial
 * columns applied from a resident block; resident_bytes: bytes of the blocks valid right now on this rank.          */
int gwbse_bse_dense_stats(gwbse_ctx* ctx, long long* builds, long long* columns, double* resident_bytes);
/* BSE_OPERATOR::diagonal (bse_operator.cc:134-175) */
int gwbse_bse_diagonal(gwbse_ctx* ctx, int cqp, int cx, int cd, int cd2, double* diag);

/* ---- Davidson helpers (davidsonsolver.cc) -------------------------------- */
/* DavidsonSolver::gramschmidt (davidsonsolver.cc:442-478) on a device matrix Q (rows x ncols):
 * orthonormalises columns [nstart, ncols) against the earlier ones, twice.  */
int gwbse_gramschmidt_dev(gwbse_ctx* ctx, int rows, int ncols, int nstart, double* Q_dev, int ldq);
/* DavidsonSolver::computeCorrectionVector, DPR / Olsen (davidsonsolver.cc:392-440):
 * W(:,j) = normalised correction for residual R(:,j), Ritz vector Q(:,j), value lambda[j].
 * olsen = 0 -> DPR.  diag_dev: operator diagonal (rows).                    */
int gwbse_davidson_correction_dev(gwbse_ctx* ctx, int rows, int ncols, int olsen, const double* diag_dev,
                                  const double* lambda, const double* R_dev, int ldr, const double* Q_dev,
                                  int ldq, double* W_dev, int ldw);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* GWBSE_B200_H */
