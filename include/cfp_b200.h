/*
 * cfp_b200.h -- C ABI of the B200-native cosmicfishpie hot path (libcfp_b200.so).
 *
 * Conventions
 *   - every entry point is extern "C", returns 0 on success or a negative cfp_status;
 *     the message of the last failure on the calling thread is cfp_last_error().
 *   - no exceptions, no torch/STL types cross the boundary; plain pointers and sizes.
 *   - pointers suffixed _h are HOST pointers, _d are DEVICE pointers; the caller owns every
 *     buffer it passes; objects returned through `**out` are owned by the library and are
 *     released with the matching *_destroy/_free call.
 *   - all floating point data is IEEE FP64, arrays are C-contiguous (row-major).
 *   - `stream` is a cudaStream_t passed as void* (NULL = the context's own stream).
 *   - one cfp_ctx per GPU; plans and banks belong to the context that created them.
 *     A context is not thread-safe; use one per host thread.
 *
 * Each entry point cites the reference (santiagocasas/cosmicfishpie v1.3.0) interface it
 * replaces as `file:line`, relative to the reference repository root.
 */
#ifndef CFP_B200_H
#define CFP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFP_ABI_VERSION 1

typedef enum {
  CFP_OK = 0,
  CFP_ERR_INVALID = -1, /* bad argument (NULL, size, range) */
  CFP_ERR_CUDA = -2,    /* CUDA runtime error, see cfp_last_error() */
  CFP_ERR_NOMEM = -3,
  CFP_ERR_STATE = -4 /* object used before it was prepared / run */
} cfp_status;

typedef struct cfp_ctx cfp_ctx;
typedef struct cfp_bank cfp_bank;
typedef struct cfp_spectro_plan cfp_spectro_plan;
typedef struct cfp_photo_plan cfp_photo_plan;

/* ---------------------------------------------------------------- context ------------- */
int cfp_abi_version(void);
const char* cfp_last_error(void);
int cfp_ctx_create(int device, cfp_ctx** out);
int cfp_ctx_destroy(cfp_ctx* ctx);
int cfp_ctx_sync(cfp_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
/* A second stream owned by the context, for the `stream` argument of the *_run / *_chi2 entry points: a caller can
 * run one batch there (from a worker thread) while it creates the next plan -- whose uploads use the context's own
 * stream -- and so overlap PCIe with compute.  The library itself never launches on it.                          */
void* cfp_ctx_aux_stream(cfp_ctx* ctx);
/* One more: consecutive likelihood chunks launched on alternating streams (cfp_photo_plan_chi2_launch) overlap each
 * other's kernels -- the C_ell contraction of chunk i+1 runs on the tensor path while the factorisations of chunk i
 * wait on instruction issue. */
void* cfp_ctx_aux_stream2(cfp_ctx* ctx);
int64_t cfp_ctx_launch_count(const cfp_ctx* ctx);
/* device time [ms] of the last *_run call, measured with CUDA events on the launch stream */
double cfp_ctx_last_run_ms(const cfp_ctx* ctx);

/* Measured FP64 throughput of this GPU in TFLOP/s (the roofline denominators for the FP64 kernels;
 * MEASURED_PEAKS.json records HBM and bf16 only): kind 0 = DFMA on the vector pipe, kind 1 = DMMA
 * (mma.sync m8n8k4 f64, the FP64 tensor path).  Runs a ~20 ms register-resident microbenchmark. */
int cfp_peak_fp64(cfp_ctx* ctx, int kind, double* tflops_out);

/* ---------------------------------------------------------------- table bank ----------
 * Replaces the spline facade `cosmo_functions.matpow / f_growthrate / growth`
 * (cosmicfishpie/cosmology/cosmology.py:1682-1738, 1859-1877, 1911-1949) built by
 * `external_input.calculate_interpol_results` (cosmology.py:1423-1540): for every cosmology
 * of the stencil the FITPACK bicubic (tx, ty, c) of each tabulated function of (z, k).
 *
 * blob_h  : all knots and coefficients, concatenated
 * desc_h  : int64 [n_cosmo][n_tables][CFP_BANK_DESC] = {nx, ny, off_tx, off_ty, off_c}
 *           nx, ny = number of knots in z and k (degree 3 both ways; the coefficient
 *           matrix is (nx-4) x (ny-4), z-major); offsets are in doubles into blob_h.
 * Table slots are the caller's convention; this package uses CFP_TAB_*.
 */
#define CFP_BANK_DESC 5
#define CFP_TAB_PLIN 0
#define CFP_TAB_PNL 1
#define CFP_TAB_F 2
#define CFP_TAB_D 3
#define CFP_TAB_PLIN_CB 4 /* cdm+baryon ("clustering" tracer) spectra and growth rate, cosmology.py:1508-1531 */
#define CFP_TAB_PNL_CB 5
#define CFP_TAB_F_CB 6
#define CFP_NTAB 7

int cfp_bank_upload(cfp_ctx* ctx, const double* blob_h, size_t n_doubles, const int64_t* desc_h,
                    int n_cosmo, int n_tables, cfp_bank** out);
int cfp_bank_free(cfp_ctx* ctx, cfp_bank* bank);

/* The same bank assembled from one host block per cosmology (rows_h[c], row_len[c] doubles), so that a caller
 * that caches the fitted tables of each cosmology (cosmology.py:1249-1540 caches them per `cosmo_functions`
 * instance) re-uploads them without re-packing.  desc_h offsets are in doubles from the start of the
 * cosmology's OWN row.  Rows allocated with cfp_host_alloc are page-locked and copy at full PCIe rate.   */
int cfp_bank_upload_rows(cfp_ctx* ctx, const double* const* rows_h, const int64_t* row_len,
                         const int64_t* desc_h, int n_cosmo, int n_tables, cfp_bank** out);

/* Page-locked host memory for buffers that are uploaded repeatedly (table rows, likelihood batches). */
int cfp_host_alloc(cfp_ctx* ctx, size_t bytes, void** out);
int cfp_host_free(cfp_ctx* ctx, void* p);
/* Point-wise bicubic evaluation, semantics of FITPACK bispeu / RectBivariateSpline.__call__
 * (grid=False), including clamping of the arguments to the knot range
 * (cosmology.py:1736-1738, 1876, 1944).  x = z, y = k, n points, out_h[n]. */
int cfp_bank_eval(cfp_ctx* ctx, const cfp_bank* bank, int cosmo, int table, const double* z_h,
                  const double* k_h, size_t n, double* out_h);

/* ---------------------------------------------------------------- spectro (GCsp) ------
 * One plan = one Fisher-matrix job: the observed galaxy power spectrum on the
 * (stencil point s, z-bin b, mu, k) lattice, the finite-difference / analytic derivatives and
 * the Fisher reduction.  Replaces, fused into two kernels:
 *   ComputeGalSpectro.observed_Pgg / lnpobs_gg   cosmicfishpie/LSSsurvey/spectro_obs.py:845-933
 *   SpectroCov.veff                              cosmicfishpie/LSSsurvey/spectro_cov.py:207-228
 *   SpectroDerivs.get_obs / exact_derivs         cosmicfishpie/LSSsurvey/spectro_cov.py:481-532
 *   derivatives.derivative_3pt (any linear stencil) cosmicfishpie/fishermatrix/derivatives.py:93-207
 *   FisherMatrix.fish_integrand / fisher_integral / pk_LSS_Fisher
 *                                                cosmicfishpie/fishermatrix/cosmicfish.py:407-542
 *
 * Per (stencil point, z-bin) scalar block, CFP_SP_NSCAL doubles (host-evaluated 1-D splines
 * and nuisance values; see cosmicfishpie_b200/spectro.py for how each is obtained):
 */
enum {
  CFP_SP_Z = 0,       /* bin-centre redshift */
  CFP_SP_COSMO,       /* index of the cosmology in the bank (as a double) */
  CFP_SP_QPAR,        /* H_fid(z)/H(z)                 spectro_obs.py:342-358 */
  CFP_SP_QPER,        /* D_A(z)/D_A,fid(z)             spectro_obs.py:360-376 */
  CFP_SP_HUBBLE,      /* H(z) of the sample cosmology  spectro_obs.py:505 */
  CFP_SP_DZ,          /* sigma_z, or sigma_z (1+z)     spectro_obs.py:500-504 */
  CFP_SP_BIAS,        /* b_g of the bin                nuisance.py:249-301 */
  CFP_SP_A1,          /* Q-bias scale dependence       nuisance.py:306-325 */
  CFP_SP_A2,
  CFP_SP_SIGMAP,      /* sigma_p (after rescale)       spectro_obs.py:678-697 */
  CFP_SP_I0,          /* P_theta-theta moments         spectro_obs.py:699-755 */
  CFP_SP_I1,
  CFP_SP_I2,
  CFP_SP_SVRESCALE,   /* sigma_v rescale factor        nuisance.py:346-350 */
  CFP_SP_FS8,         /* multiplies f(z,k): sigma8(z) if bfs8terms else 1   spectro_obs.py:624-627 */
  CFP_SP_INVS8SQ,     /* 1/sigma8(z)^2 if bfs8terms else 1                  spectro_obs.py:773-777 */
  CFP_SP_KHFAC,       /* h/h_fid if kh_rescaling_bug else 1                 spectro_obs.py:419-425 */
  CFP_SP_PSHOT,       /* extra shot noise added to P_obs (0 in the reference, nuisance.py:364-366) */
  CFP_SP_NSCAL
};
/* flags (bit mask) */
#define CFP_SPF_AP 1u              /* Alcock-Paczynski remap on      (settings AP_effect) */
#define CFP_SPF_FOG 2u             /* Lorentzian FoG on              (settings FoG_switch) */
#define CFP_SPF_LINEAR 4u          /* GCsp_linear: no FoG, no dewiggling */
#define CFP_SPF_KH_BEFORE_ERR 8u   /* kh_rescaling_beforespecerr_bug */

/* Create a plan and upload the job description.
 *   S, Z, Nmu, Nk, npar : stencil points, z-bins, lattice, free parameters
 *   k_h[Nk], mu_h[Nmu]  : FisherMatrix.set_pk_settings grids (cosmicfish.py:366-381), 1/Mpc
 *   wk_h[Nk], wmu_h[Nmu]: quadrature weights reproducing scipy.integrate.simpson on those grids
 *   scal_h[S][Z][CFP_SP_NSCAL]
 *   alias_h[S][Z]       : s if the point must be evaluated in this bin, 0 if its (cosmology,
 *                         scalars) are identical to the fiducial point 0 there (e.g. a bias step
 *                         of another bin): the fiducial lattice is reused, not recomputed
 *   nnw, pnw_k_h[n_cosmo][nnw], pnw_p_h[n_cosmo][Z][nnw]: no-wiggle spectrum samples, the knots
 *                         and values of the degree-1 spline of cosmology.py:1804-1812
 *   stw_h[npar][S], stden_h[npar] : derivative stencil, d_p = (sum_s stw[p][s] lnP_s) / stden[p]
 *                         (3PT: weights +1/-1 on theta+-h, denominator 2h  derivatives.py:93-111;
 *                         any linear stencil -- 5PT, 4PT_FWD derivatives.py:209-225 -- fits).
 *                         A parameter whose stencil points all alias to the fiducial in a bin has
 *                         an exactly zero derivative there, as in the reference.
 *   ps_bin_h[npar]      : -1 for finite-difference parameters; for an analytic shot-noise
 *                         parameter Ps_i the z-bin index i-1: d_p = delta(b, i-1) / P_obs(ps_src)
 *   ps_src              : stencil point whose P_obs enters the analytic derivative (the reference
 *                         uses the LAST evaluated point, spectro_cov.py:465,526)
 *   nbar_h[Z], vol_h[Z] : fiducial number density and survey volume per bin (spectro_cov.py:169-205)
 *   tab_plin, tab_f     : bank table slots of P_lin(z,k) and f(z,k)
 */
int cfp_spectro_plan_create(cfp_ctx* ctx, const cfp_bank* bank, int S, int Z, int Nmu, int Nk,
                            int npar, const double* k_h, const double* mu_h, const double* wk_h,
                            const double* wmu_h, const double* scal_h, const int32_t* alias_h,
                            int nnw, const double* pnw_k_h, const double* pnw_p_h,
                            const double* stw_h, const double* stden_h, const int32_t* ps_bin_h,
                            int ps_src, const double* nbar_h, const double* vol_h, unsigned flags,
                            int tab_plin, int tab_f, cfp_spectro_plan** out);
int cfp_spectro_plan_destroy(cfp_spectro_plan* plan);
/* Restrict the plan to a slice of the flattened (mu,k) lattice [cell_begin, cell_end) of every
 * bin: the multi-GPU partition (each rank owns a slice; partial Fishers are summed). */
int cfp_spectro_plan_set_slice(cfp_spectro_plan* plan, int64_t cell_begin, int64_t cell_end);
/* Launch: table preparation (z-contraction + piecewise-polynomial form) and the fused
 * lattice / derivative / Fisher kernel.  materialize: bit 0 = keep lnP[S][Z][Nmu][Nk],
 * bit 1 = keep derivs[npar][Z][Nmu][Nk] and veff[Z][Nmu][Nk] (the API-preserving mode). */
#define CFP_MAT_LNP 1
#define CFP_MAT_DERIVS 2
int cfp_spectro_plan_run(cfp_spectro_plan* plan, int materialize, void* stream);
/* Device time [ms] of the fused lattice/derivative/Fisher kernel alone in the last run (CUDA events on
 * the launch stream around that one launch); synchronises. */
double cfp_spectro_plan_kernel_ms(cfp_spectro_plan* plan);
/* Parts of the lattice pass of the last run: 0 = the whole pass (as above), 1 = the weights pre-pass, 2 = the main
 * kernel (spectro_fisher_pairs_kernel); -1 when the last run did not take the pair kernels. */
double cfp_spectro_plan_kernel_ms_part(cfp_spectro_plan* plan, int which);
/* The factors of P_obs of stencil point s as separate lattices, terms_h[5][Z][Nmu][Nk] (host): 0 Kaiser term, 1
 * Fingers-of-God factor, 2 de-wiggled spectrum, 3 redshift-error factor, 4 P_obs.  Replaces ComputeGalSpectro.kaiserTerm /
 * FingersOfGod / dewiggled_pdd / spec_err_z / observed_Pgg (cosmicfishpie/LSSsurvey/spectro_obs.py:476-506, 584-676,
 * 811-911) evaluated on a lattice. */
int cfp_spectro_plan_terms(cfp_spectro_plan* plan, int s, double* terms_h);
/* Test hook: y[i] = f(x[i]) on the device for the branch-free FP64 functions of the lattice loops, so that tests can
 * sweep their stated domains against libm.  which: 0 exp(x), x <= 0 (table version); 1 log(x), x > 0 normal;
 * 2 1/x; 3 1/sqrt(x); 4 1/x by division; 5 exp(x), x <= 0 (pair-kernel version).  Host buffers. */
int cfp_math_selftest(cfp_ctx* ctx, int which, const double* x_h, int n, double* y_h);
/* Device pointer of the result F[Z][npar][npar] (valid until the plan is destroyed). */
const double* cfp_spectro_plan_fisher_d(const cfp_spectro_plan* plan);
/* Copy results to the host (synchronises). Any pointer may be NULL. */
int cfp_spectro_plan_fetch(cfp_spectro_plan* plan, double* fisher_h, double* lnp_h,
                           double* derivs_h, double* veff_h);

/* ---------------------------------------------------------------- photo (WL/GCph/3x2pt)
 * One plan = one photometric Fisher job.  Replaces:
 *   ComputeCls.sqrtP_limber                      cosmicfishpie/LSSsurvey/photo_obs.py:457-512
 *   much_faster_integral_efficiency              cosmicfishpie/LSSsurvey/photo_obs.py:113-160
 *   ComputeCls.lensing_kernel / galaxy_kernel / genwindow   photo_obs.py:514-722
 *   ComputeCls.clsintegral / computecls_vectorized          photo_obs.py:798-928
 *   PhotoCov.getclsnoise / get_covmat            cosmicfishpie/LSSsurvey/photo_cov.py:164-245
 *   derivatives.derivative_3pt                   cosmicfishpie/fishermatrix/derivatives.py:93-207
 *   FisherMatrix.photo_LSS_fishermatrix_einsum   cosmicfishpie/fishermatrix/cosmicfish.py:643-744
 *
 * Tomographic rows are ordered as the reference orders its covariance columns: all bins of
 * observable 0, then all bins of observable 1 (cosmicfish.py:686-691).  kind[a] = 0 for a
 * lensing (WL) row, 1 for a clustering (GCph) row.
 *
 *   S, NC, Nl, Nz, Nb, npar
 *   cosmo_of_s_h[S]      : cosmology index of each stencil point
 *   ell_h[Nl], z_h[Nz]   : photo_obs.py:309-316 grids;  dz = mean spacing (photo_obs.py:316)
 *   chi_h[NC][Nz], hub_h[NC][Nz] : comoving distance and H(z) on the z grid (host 1-D splines)
 *   wlpref_h[NC]         : 1.5 H0^2 Omega_m                        photo_obs.py:584
 *   kind_h[Nb]           : 0 = WL row, 1 = GCph row
 *   ngal_h[Nb][Nz]       : normalised n_i(z) of each row's bin       photo_window.py:220-242
 *   bias_h[S][Nb][Nz]    : galaxy bias b(z) per stencil point (rows of kind 0: ignored)
 *   ia_h[S][Nz]          : intrinsic-alignment window IA(z) per stencil point  nuisance.py:155-191
 *   lmin_h[Nb], lmax_h[Nb] : multipole range of each row's observable  photo_obs.py:899-900
 *   noise_h[Nb]          : shape / shot noise added on the diagonal    photo_cov.py:111-114
 *   sqrtfsky_h[Nb]       : sqrt(f_sky) of each row's observable        photo_cov.py:228
 *   stw_h[npar][S], stden_h[npar] : derivative stencil (as for the spectro plan)
 *   kmax                 : Limber k cut (strict <)                     photo_obs.py:482
 *   tab_p                : bank slot of the matter power spectrum used (P_nl or P_lin)
 */
int cfp_photo_plan_create(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz,
                          int Nb, int npar, const int32_t* cosmo_of_s_h, const double* ell_h,
                          const double* z_h, double dz, const double* chi_h, const double* hub_h,
                          const double* wlpref_h, const int32_t* kind_h, const double* ngal_h,
                          const double* bias_h, const double* ia_h, const double* lmin_h,
                          const double* lmax_h, const double* noise_h, const double* sqrtfsky_h,
                          const double* stw_h, const double* stden_h, double kmax, int tab_p,
                          cfp_photo_plan** out);

/* The same with the galaxy bias given on Kb <= Nz samples per (point, row) and a map zmap_h[Nz] -> sample: the
 * reference's `binned` model (nuisance.py:98-230) is constant inside each tomographic bin, so a batch of B
 * likelihood points uploads B*Nb*n_bins doubles instead of B*Nb*Nz.  bias_h[S][Nb][Kb].                         */
int cfp_photo_plan_create_zmap(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb, int npar,
                               const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h, double dz,
                               const double* chi_h, const double* hub_h, const double* wlpref_h,
                               const int32_t* kind_h, const double* ngal_h, const double* bias_h, int Kb,
                               const int32_t* zmap_h, const double* ia_h, const double* lmin_h, const double* lmax_h,
                               const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                               const double* stden_h, double kmax, int tab_p, cfp_photo_plan** out);

/* As cfp_photo_plan_create_zmap with ONE bias vector per point shared by every clustering row, bias_h[S][Kb]: the
 * `binned` model assigns the same per-bin values to all rows (nuisance.py:118-128), so a likelihood batch uploads
 * S*n_bins doubles. */
int cfp_photo_plan_create_zmap_shared(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb, int npar,
                                      const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h, double dz,
                                      const double* chi_h, const double* hub_h, const double* wlpref_h,
                                      const int32_t* kind_h, const double* ngal_h, const double* bias_h, int Kb,
                                      const int32_t* zmap_h, const double* ia_h, const double* lmin_h,
                                      const double* lmax_h, const double* noise_h, const double* sqrtfsky_h,
                                      const double* stw_h, const double* stden_h, double kmax, int tab_p,
                                      cfp_photo_plan** out);

/* The general form: NW sets of tomographic n_i(z) windows (ngal_h[NW][Nb][Nz]) with one index per stencil point
 * (win_of_s_h[S], NULL when NW == 1) -- stencil points that vary a photo-z parameter carry their own windows
 * (photo_cov.py:116-160 builds a new GalaxyPhotoDist per parameter point); zmap_h may be NULL (bias_h[S][Nb][Nz]). */
int cfp_photo_plan_create_ex(cfp_ctx* ctx, const cfp_bank* bank, int S, int NC, int Nl, int Nz, int Nb, int npar,
                             const int32_t* cosmo_of_s_h, const double* ell_h, const double* z_h, double dz,
                             const double* chi_h, const double* hub_h, const double* wlpref_h, const int32_t* kind_h,
                             int NW, const int32_t* win_of_s_h, const double* ngal_h, const double* bias_h, int Kb,
                             const int32_t* zmap_h, const double* ia_h, const double* lmin_h, const double* lmax_h,
                             const double* noise_h, const double* sqrtfsky_h, const double* stw_h,
                             const double* stden_h, double kmax, int tab_p, cfp_photo_plan** out);

/* Intrinsic-alignment windows of all points computed on the device (the eNLA model, cosmicfishpie/LSSsurvey/
 * nuisance.py:190-230): point s has W_IA(z_j) = fac_h[s] * zdep_h[j]^eta_h[s] * (lum_h[j]^beta_h[s] / growth_h[c_s][j])
 * on the nuisance grid of nia redshifts, zdep = (1 + z) / (1 + pivot_z_IA), fac = -A_IA C_IA Omega_m(0); sample iz of
 * the plan's z grid takes (1 - t_h[iz]) W_IA(z_j) + t_h[iz] W_IA(z_{j+1}) with j = j_h[iz] (the reference's degree-1
 * spline).  Replaces the `ia_h` array of plan creation (which may then be NULL): a likelihood batch uploads three
 * numbers per point instead of Nz.  Call before the plan runs. */
int cfp_photo_plan_set_ia_enla(cfp_photo_plan* plan, int nia, const double* zdep_h, const double* lum_h,
                               const double* growth_h, const double* fac_h, const double* eta_h, const double* beta_h,
                               const int32_t* j_h, const double* t_h);

/* Table slot of the spectrum the CLUSTERING rows use when it differs from tab_p (GCph_Tracer = "clustering":
 * photo_obs.py:483-507 evaluates sqrt(P) of GCph on the cdm+baryon spectrum, lensing rows stay on total matter).
 * Call between create and run.                                                                                    */
int cfp_photo_plan_set_gc_table(cfp_photo_plan* plan, int tab_p_gc);
int cfp_photo_plan_destroy(cfp_photo_plan* plan);
/* multi-GPU partition: this rank owns multipoles [l_begin, l_end) */
int cfp_photo_plan_set_slice(cfp_photo_plan* plan, int l_begin, int l_end);
/* Launch K3 (Limber kernels), K4 (FP64 tensor-core C_ell contraction) and K5 (Cholesky + trace
 * Fisher).  C_ell of all stencil points is always kept: cls[S][Nl][Nb][Nb]. */
int cfp_photo_plan_run(cfp_photo_plan* plan, void* stream);
/* Device time [ms] of one kernel of the last run: which = 0 Limber table, 1 C_ell contraction (DMMA),
 * 2 Cholesky/trace Fisher; synchronises. */
/* cfp_photo_plan_chi2 in two halves: _launch enqueues everything on `stream` (NULL: the context stream) and returns
 * without waiting, _collect waits for the result and copies chi2_h[S] out.  A sampler prepares and uploads its next
 * batch between the two, so the device never idles (PhotometricLikelihood.chi2_matrix does exactly that). */
int cfp_photo_plan_chi2_launch(cfp_photo_plan* plan, const double* data_h, int nblk, const int32_t* blk_off_h,
                               const int32_t* blk_n_h, const double* blk_w_h, void* stream);
int cfp_photo_plan_chi2_collect(cfp_photo_plan* plan, double* chi2_h);
double cfp_photo_plan_kernel_ms(cfp_photo_plan* plan, int which);
/* Radial kernels of the last run on the plan's z grid, [S][Nb][Nz] each (host): U = W_a / sqrt(H) (lensing efficiency
 * term or galaxy kernel x bias), V = W^IA_a / sqrt(H).  Replaces ComputeCls.lensing_kernel / galaxy_kernel
 * (cosmicfishpie/LSSsurvey/photo_obs.py:514-615) at the grid redshifts. */
int cfp_photo_plan_fetch_windows(cfp_photo_plan* plan, double* U_h, double* V_h);
/* Fisher matrix from caller-supplied arrays -- the reference's seam photo_LSS_fishermatrix_einsum(noisy_cls, covmat,
 * derivs), cosmicfishpie/fishermatrix/cosmicfish.py:643-744: cov_h[Nl][Nb][Nb] (the covariance matrices), dC_h[npar][Nl][Nb][Nb]
 * (derivatives of the spectra), ell_h[Nl] (bin edges) -> fisher_h[npar][npar] = sum over the Nl - 1 bins of
 * (l_centre + 1/2) dl Tr(cov^-1 dC_p cov^-1 dC_q), left-edge arrays.  Host buffers, Nb <= 64. */
int cfp_photo_fisher(cfp_ctx* ctx, int Nl, int Nb, int npar, const double* cov_h, const double* dC_h, const double* ell_h,
                     double* fisher_h);
const double* cfp_photo_plan_fisher_d(const cfp_photo_plan* plan);
/* fisher_h[npar][npar], cls_h[S][Nl][Nb][Nb] (signal only, no noise), sqrtp_h[NC][Nl][Nz] */
int cfp_photo_plan_fetch(cfp_photo_plan* plan, double* fisher_h, double* cls_h, double* sqrtp_h);

/* ---------------------------------------------------------------- batched likelihoods ---
 * The stencil points of a plan double as a BATCH of parameter points (S = batch size): the same kernels
 * evaluate the observables of every point, then a chi^2 kernel replaces the Fisher reduction.
 *
 * Photometric (replaces _cells_from_cls, _chi2_per_obs, PhotometricLikelihood.compute_chi2,
 * cosmicfishpie/likelihood/photo_like.py:28-97,180-256): for every point s and multipole l and each block b
 * (a contiguous range of tomographic rows [off, off+n)):
 *     q = tr(A^-1 B) + ln det A - ln det B - n,   A = theory block (C_l + N), B = data block
 * (identical to the reference's determinant-ratio form by Cramer's rule), chi2[s] = sum_b sum_l w[b][l] q.
 * The caller folds (2l+1), the multipole quadrature of photo_like.py:189-204,90-96 and the f_sky factors into w.
 *   data_h[Nl][Nb][Nb] : data spectra INCLUDING noise, or NULL = stencil point 0 plus the plan's noise
 *   nblk, blk_off_h[nblk], blk_n_h[nblk] (n <= 32), blk_w_h[nblk][Nl]
 *   chi2_h[S]          : output
 */
int cfp_photo_plan_chi2(cfp_photo_plan* plan, const double* data_h, int nblk, const int32_t* blk_off_h,
                        const int32_t* blk_n_h, const double* blk_w_h, double* chi2_h, void* stream);

/* Spectroscopic wedges (replaces observable_Pgg, compute_wedge_chi2, compute_theory_spectro,
 * cosmicfishpie/likelihood/spectro_like.py:25-41,143-233): for every point s
 *     chi2[s] = sum_z sum_cells wmu wk 2 k^2 V_z / (8 pi^2) (P_d - P_t)^2 / P_d^2,
 *     P_t = P_obs(s) + 1/nbar_z + shot_h[s][z],   P_d = data
 * with the plan's Simpson weights, volumes and number densities.
 *   data_h[Z][Nmu][Nk] : data spectra including noise, or NULL = stencil point 0 + 1/nbar_z + shot_data_h[z]
 *   shot_h[S][Z], shot_data_h[Z] (may be NULL = 0)
 * Points of a bin whose scalar blocks differ only in the galaxy bias are evaluated together: groups of >= 4 through
 * the ten lattice moments of the group (one pass over the lattice per group and bin, chi2 = v^T M v per member),
 * smaller groups in one pass of the direct kernel.  Environment CFP_NO_MOMENTS=1 (read per call) disables the moment
 * path; the same applies to cfp_spectro_plan_chi2_legendre.
 */
int cfp_spectro_plan_chi2(cfp_spectro_plan* plan, const double* data_h, const double* shot_h,
                          const double* shot_data_h, double* chi2_h, void* stream);

/* The same with the data lattice already resident on the device (cfp_dev_upload): a sampler evaluates
 * thousands of batches against ONE data vector, which therefore crosses PCIe once, not once per batch. */
int cfp_spectro_plan_chi2_dev(cfp_spectro_plan* plan, const double* data_d, const double* shot_h, double* chi2_h,
                              void* stream);

/* Legendre multipoles (replaces legendre_Pgg + compute_chi2_legendre over a batch, spectro_like.py:48-56,125-140,
 * 192-233), fused: chi2[s] = sum_{k,z} d^T C^-1 d,  d_l = P_l,data(k,z) - sum_mu lw_l(mu) P_t(s; k, mu, z), l = 0,2,4.
 *   pl_data_h[Nk][Z][3], icov_h[Nk][Z][3][3] (cfp_legendre_covariance), lw_h[3][Nmu] = (2l+1) * Simpson weight * L_l(mu),
 *   shot_h[S][Z].  Nmu <= 256.                                                                                    */
int cfp_spectro_plan_chi2_legendre(cfp_spectro_plan* plan, const double* pl_data_h, const double* icov_h,
                                   const double* lw_h, const double* shot_h, double* chi2_h, void* stream);
int cfp_dev_upload(cfp_ctx* ctx, const double* src_h, size_t n_doubles, double** out_d);
int cfp_dev_release(cfp_ctx* ctx, double* p_d);

/* Stand-alone forms on spectra the caller already holds (host arrays): the operator seams
 * `_chi2_per_obs(cell_fid, cell_th, ells, dells)` (photo_like.py:78-97) and
 * `compute_wedge_chi2(P_obs_data, P_obs_theory, cosmoFM)` (spectro_like.py:143-189).
 *   theory_h[S][Nl][Nb][Nb], data_h[Nl][Nb][Nb] : spectra including noise; blocks/weights as above
 *   theory_h[S][Z][Nmu][Nk], data_h[Z][Nmu][Nk], k_h[Nk], wk_h[Nk], wmu_h[Nmu], vol_h[Z]
 */
int cfp_cells_chi2(cfp_ctx* ctx, int S, int Nl, int Nb, const double* theory_h, const double* data_h, int nblk,
                   const int32_t* blk_off_h, const int32_t* blk_n_h, const double* blk_w_h, double* chi2_h);
int cfp_wedge_chi2(cfp_ctx* ctx, int S, int Z, int Nmu, int Nk, const double* theory_h, const double* data_h,
                   const double* k_h, const double* wk_h, const double* wmu_h, const double* vol_h, double* chi2_h);

/* Legendre-multipole form of the spectroscopic likelihood (ell = 0, 2, 4), host arrays in and out:
 *   cfp_legendre_project      legendre_Pgg                  cosmicfishpie/likelihood/spectro_like.py:48-56
 *       lat_h[S][Z][Nmu][Nk], lw_h[3][Nmu] = (2 ell + 1) * Simpson weight * L_ell(mu)  ->  pl_h[S][Nk][Z][3]
 *   cfp_legendre_covariance   compute_covariance_legendre   spectro_like.py:69-122 (k-bin integrated Gaussian
 *       covariance with the Wigner-3j sum matrices of utilities/legendre_tools.py:56-123, then its inverse)
 *       pl_h[Nk][Z][3] (data), k_h[Nk], vol_h[Z], m_h[6][3][3] = m00,m22,m44,m02,m24,m04
 *       ->  cov_h[Nk][Z][3][3], icov_h[Nk][Z][3][3]
 *   cfp_legendre_chi2         compute_chi2_legendre         spectro_like.py:125-140
 *       pl_data_h[Nk][Z][3], pl_theory_h[S][Nk][Z][3], icov_h  ->  chi2_h[S]
 */
int cfp_legendre_project(cfp_ctx* ctx, int S, int Z, int Nmu, int Nk, const double* lat_h, const double* lw_h,
                         double* pl_h);
int cfp_legendre_covariance(cfp_ctx* ctx, int Nk, int Z, const double* pl_h, const double* k_h, const double* vol_h,
                            const double* m_h, double* cov_h, double* icov_h);
int cfp_legendre_chi2(cfp_ctx* ctx, int S, int Nk, int Z, const double* pl_data_h, const double* pl_theory_h,
                      const double* icov_h, double* chi2_h);

/* Cosmology emulation on the device: a new bank = the NC rows of `base` followed by M rows whose bicubic
 * coefficient arrays are the weighted sums sum_c W_h[m][c] * row_c (knots copied).  Interpolating-spline fitting is
 * linear in the table values, so with the weights of a Taylor / interpolation scheme over tabulated cosmologies
 * (cosmicfishpie_b200/bank.py) row NC + m is the table set of the emulated cosmology m, built without any host fit or
 * H2D of tables -- where the reference snaps an off-grid cosmology to the nearest folder
 * (cosmicfishpie/cosmology/cosmology.py:1375-1421).  All rows of `base` must share table shapes and knots. */
int cfp_bank_combine(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out);
/* The same, but the new bank holds the M combined rows ONLY, in the order of W_h's rows (a row of W_h that is one-hot
 * reproduces a base row): the bank of a likelihood batch whose every point has its own emulated cosmology, built
 * without the tables leaving the device. */
int cfp_bank_combine_rows(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out);

/* ---- device-side preparation of the cosmology input ---------------------------------------------
 * What `external_input.calculate_interpol_results` and the no-wiggle / velocity-moment helpers redo on the host for
 * every cosmology of a forecast, batched over all cosmologies of a job (csrc/cfp_prep.cu).
 *
 * cfp_bank_fit: the bank of cfp_bank_upload, fitted on the device from the RAW tables -- for every cosmology c and
 * table slot t with tab_h[c * n_tables + t] != NULL (a [Nz][Nk] array; the same slots for every cosmology) the
 * interpolating bicubic spline RectBivariateSpline(z, k_c, scale * table, kx=3, ky=3) on FITPACK's knots
 * (cosmicfishpie/cosmology/cosmology.py:1463-1532; scale_h[c][t] is the unit factor of :1438-1444).
 * z_h[Nz] is shared, k_h[n_cosmo][Nk] per cosmology (the k-units factor depends on h). */
int cfp_bank_fit(cfp_ctx* ctx, int n_cosmo, int n_tables, int Nz, int Nk, const double* z_h, const double* k_h,
                 const double* const* tab_h, const double* scale_h, cfp_bank** out);

/* Background functions of every cosmology at the redshifts zq_h[nq]: out_h[n_cosmo][3 + n_extra][nq] =
 * { H(z), comoving distance, angular-diameter distance, extra_0(z), ... } where H and the extras (sigma8(z) ...,
 * hub_h[n_cosmo][Nz], extra_h[n_cosmo][n_extra][Nz]) are InterpolatedUnivariateSpline fits of the tables, the
 * comoving distance to each tabulated redshift is the ntrap-point trapezoid of 1/H on linspace(0, z, ntrap) and the
 * two distances are splines through those values (cosmology.py:36-55, 1447-1475, 1500-1506, 1618-1680).
 * coef_h (may be NULL): [n_cosmo][3 + n_extra][Nz] B-spline coefficients of the same functions on the not-a-knot knot
 * vector of z_h (z_0 x4, z_2 .. z_{Nz-3}, z_{Nz-1} x4), i.e. the `c` of each spline's FITPACK (t, c, 3); nq may be 0. */
int cfp_background_fit(cfp_ctx* ctx, int n_cosmo, int Nz, const double* z_h, const double* hub_h, int n_extra,
                       const double* extra_h, int ntrap, int nq, const double* zq_h, double* out_h, double* coef_h);

/* cfp_bank_eval for every cosmology of the bank at once: out_h[n_cosmo][n] (growth(z, k), f_growthrate(z, k),
 * cosmology.py:1859-1949). */
int cfp_bank_eval_many(cfp_ctx* ctx, const cfp_bank* bank, int table, size_t n, const double* z_h, const double* k_h,
                       double* out_h);

/* No-wiggle spectrum of every cosmology at the redshifts zq_h[nzq] (cosmology.py:1760-1813): ln P(z, k) on the samples
 * k_h[n_cosmo][nsamp], smoothed by a Savitzky-Golay filter given as a linear operator, exponentiated:
 * out_h[n_cosmo][nzq][nsamp] (the values of the degree-1 spline whose knots are k_h).  Filter f (filt_of_c_h[c] selects one
 * per cosmology: the window length depends on h) is fdesc_h[f] = {win, off, ne, wc} and, at fblob_h + foff_h[f],
 * taps[win] | EL[ne][wc] | ER[ne][wc]: row s of the output is sum_u taps[u] x[s + off + u] in the interior, row s < ne
 * is EL[s] . x[0 .. wc), row nsamp - ne + s is ER[s] . x[nsamp - wc .. nsamp) (scipy's mode="interp" edge fit). */
int cfp_bank_nowiggle(cfp_ctx* ctx, const cfp_bank* bank, int table, int nzq, const double* zq_h, int nsamp,
                      const double* k_h, int n_filt, const int32_t* filt_of_c_h, const int32_t* fdesc_h,
                      const int64_t* foff_h, const double* fblob_h, size_t fblob_len, double* out_h);

/* Velocity moments (1 / 6 pi^2) int P(z, k) f(z, k)^m dk, m = 0, 1, 2, by the trapezoid rule on k_h[nk], of every
 * cosmology at zq_h[nzq]: out_h[n_cosmo][nzq][3] (LSSsurvey/spectro_obs.py:724-755). */
int cfp_bank_theta_moments(cfp_ctx* ctx, const cfp_bank* bank, int tab_p, int tab_f, int nzq, const double* zq_h, int nk,
                           const double* k_h, double* out_h);

/* ---- batched Fisher-matrix post-processing ------------------------------------------------------
 * B symmetric positive-definite matrices of n <= 64 parameters in one launch.
 * cfp_fisher_batch_post: sigma_h[B][n] = coef * sqrt(diag(F^-1))  (fisher_matrix.get_confidence_bounds,
 *   cosmicfishpie/analysis/fisher_matrix.py:705-733, coef = sqrt(2) erfinv(confidence level)); with m > 0 also the
 *   Fisher matrix marginalised onto the parameters keep_h[m]: marg_h[B][m][m] = ((F^-1)[keep][keep])^-1
 *   (analysis/fisher_operations.py:177-237) and fom_h[B] = sqrt(det marg) (for m = 2, e.g. (w0, wa): the dark-energy
 *   figure of merit).  A matrix that is not positive definite gives NaN.
 * cfp_fisher_batch_add: out_h[B][n_out][n_out] = a_h[b] scattered through map_a_h[na] + b_h[b] scattered through
 *   map_b_h[nb] -- the sum over the union of parameters of fisher_matrix.__add__ (analysis/fisher_matrix.py:462-548),
 *   the name -> index maps being the caller's.  Host buffers. */
int cfp_fisher_batch_post(cfp_ctx* ctx, int B, int n, const double* fisher_h, int m, const int32_t* keep_h, double coef,
                          double* sigma_h, double* marg_h, double* fom_h);
int cfp_fisher_batch_add(cfp_ctx* ctx, int B, int na, int nb, int n_out, const int32_t* map_a_h, const int32_t* map_b_h,
                         const double* a_h, const double* b_h, double* out_h);

#ifdef __cplusplus
}
#endif
#endif /* CFP_B200_H */
