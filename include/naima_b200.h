/* naima_b200.h -- C ABI of the B200-native naima likelihood hot path.
 *
 * The reference (zblz/naima, pure Python) has no FFI; the functions below are
 * the entry points a ctypes/cffi binding in the reference would bind to replace
 * its NumPy hot path.  Each entry cites the reference code it replaces (paths
 * relative to src/naima/ of zblz/naima @ ba20a64).
 *
 * Conventions
 *  - plain C, no C++/torch types; every array pointer is a DEVICE pointer to
 *    contiguous float64 unless the name ends in `_host`;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *    calls only enqueue work and return immediately;
 *  - return value: 0 ok, <0 argument error (NB_E*), >0 a cudaError_t;
 *  - no global state; the library never allocates device memory for the
 *    caller except inside nb_plan objects (nb_plan.h section below).
 *
 * Units (fixed, stripped on the host): energies eV, B in G, T in K, lengths cm,
 * densities cm^-3, angles rad; differential spectra 1/(s eV).
 *
 * Particle distributions (models.py): `kind` is NB_PD_*, `pd_params` is
 * [W][NB_PD_MAXPAR] with the reference eval() argument order after `e`:
 *   NB_PD_PL      amplitude[1/eV], e_0, alpha                      models.py:87-92
 *   NB_PD_ECPL    amplitude, e_0, alpha, e_cutoff, beta            models.py:156-161
 *   NB_PD_BPL     amplitude, e_0, e_break, alpha_1, alpha_2        models.py:233-238
 *   NB_PD_ECBPL   amplitude, e_0, e_break, alpha_1, alpha_2, e_cutoff, beta  :329-335
 *   NB_PD_LOGPAR  amplitude, e_0, alpha, beta                      models.py:401-407
 *
 * Integration layout ("rows"): every process reduces to log-log trapezoids
 * (utils.py:285-355) of n[w,j] * K[r,j] over the particle grid x[j], where the
 * emissivity table K is walker independent.  Row r = c * N_E + e for
 * component c (seed photon field, ee/ep, ...) and photon energy e.  Tables are
 * [R][pitch] with pitch >= N, pitch even.
 */
#ifndef NAIMA_B200_H
#define NAIMA_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define NB_PD_PL 0
#define NB_PD_ECPL 1
#define NB_PD_BPL 2
#define NB_PD_ECBPL 3
#define NB_PD_LOGPAR 4
#define NB_PD_MAXPAR 8

#define NB_PP_GEANT4 0
#define NB_PP_PYTHIA8 1
#define NB_PP_SIBYLL 2
#define NB_PP_QGSJET 3

#define NB_EINVAL (-1)    /* bad argument (null pointer, negative size, bad kind) */
#define NB_ETOOLARGE (-2) /* grid does not fit the kernel's shared-memory tiling */
#define NB_EALIGN (-3)    /* pointer / pitch alignment requirement violated */

#define NB_MAX_TERMS 16

/* library version (major*10000 + minor*100 + patch) */
int nb_version(void);
/* human readable text for a return code of this library */
const char* nb_strerror(int code);
/* shared-memory bytes nb_contract needs for a grid of N nodes (0 if too large) */
int nb_contract_smem_bytes(int N, int rows_per_tile);

/* --- log-log trapezoid as a stand-alone op (utils.py:285-355) ------------------
 * out[r] = trapz_loglog(y[r][0..N), x) in the reference's operation order;
 * x is shared (x_ld == 0) or per row (x[r][0..N), row stride x_ld).  `intervals`
 * (may be NULL) receives the N-1 interval values per row; `out` may be NULL. */
int nb_trapz_loglog(const double* y, int R, int N, int ld, const double* x, int x_ld,
                    double* out, double* intervals, void* stream);

/* --- particle distribution ------------------------------------------------
 * out[w][i] = PD.eval(e[i], params[w])                     models.py:87-335 */
int nb_pdist_eval(int kind, const double* pd_params, int W, const double* e_eV, int N,
                  double* out, void* stream);
/* same with an output row pitch out_ld >= N (rows inside a wider per-walker record) */
int nb_pdist_eval_ld(int kind, const double* pd_params, int W, const double* e_eV, int N,
                     double* out, int out_ld, void* stream);

/* Per-walker integration operands on a particle grid x[N]
 * (BaseElectron._gam/_nelec radiative.py:147-160, BaseProton._Ep/_J :1002-1015):
 *   e_j   = (x[j] * e_mul1) * e_mul2        [eV]
 *   n_j   = PD.eval(e_j) * n_scale
 *   xn    [w][j] = x[j] * n_j                              j < N
 *   ds1   [w][j] = ln(n_{j+1}/n_j) * invdlx[j] + 1         j < N-1
 * invdlx[j] = 1/ln(x[j+1]/x[j]).  wpitch >= N. */
int nb_pd_prep(int kind, const double* pd_params, int W, const double* x, int N,
               double e_mul1, double e_mul2, double n_scale, const double* invdlx,
               double* xn, double* ds1, int wpitch, void* stream);

/* nb_pd_prep that also stores n[w][j] itself in nraw (may be NULL); the
 * exact contraction (nb_contract exact != 0) consumes nraw. */
int nb_pd_prep_ex(int kind, const double* pd_params, int W, const double* x, int N,
                  double e_mul1, double e_mul2, double n_scale, const double* invdlx,
                  double* xn, double* ds1, double* nraw, int wpitch, void* stream);

/* Total particle energy  W = trapz_loglog(x*n, x*x_to_energy)  in the unit of
 * x_to_energy (We radiative.py:162-195, Wp :1017-1055), reference operation
 * order.  out[W]. */
int nb_particle_energy(int kind, const double* pd_params, int W, const double* x, int N,
                       double e_mul1, double e_mul2, double n_scale, double x_to_energy,
                       double* out, void* stream);

/* --- emissivity tables (walker independent) --------------------------------
 * IC on grey-body seeds: Khangulyan+14 Eq.14 (theta[s] NaN) or Eq.11
 * (radiative.py:547-607).  Eph = photon energy / mec2.  Writes rows
 * [row0 + s*N_E + e]. */
int nb_ic_planck_table(const double* gam, int N, const double* Eph, int N_E,
                       const double* seed_T, const double* seed_theta, int S, double* K,
                       int pitch, int row0, void* stream);

/* IC on a monochromatic (Ns == 1, phn = energy density in mec2/cm3) or tabulated
 * (Ns > 1, phn = dn/dE in 1/(mec2 cm3)) isotropic seed: Aharonian & Atoyan 81
 * Eq.22 incl. the inner log-log trapezoid over seed energy
 * (radiative.py:609-655).  eps0 = seed energies / mec2.  Writes rows
 * [row0 + e]. */
int nb_ic_seed_table(const double* gam, int N, const double* Eph, int N_E,
                     const double* eps0, const double* phn, int Ns, double* K, int pitch,
                     int row0, void* stream);

/* Bremsstrahlung, Baring+99 (radiative.py:838-938): rows [row0 + e] get
 * sigma_ee / mec2_eV (cm2/eV), rows [row0 + N_E + e] get sigma_1 (cm2/mec2). */
int nb_brems_table(const double* gam, int N, const double* eps, int N_E, double* K,
                   int pitch, int row0, void* stream);

/* Pion decay dsigma/dEgamma [cm2/GeV], Kafexhiu+14 analytic (radiative.py:1215-1482). */
int nb_pp_analytic_table(int hiEmodel, int nuclear_enhancement, const double* Ep_GeV, int N,
                         const double* Eg_GeV, int N_E, double* K, int pitch, int row0,
                         void* stream);

/* Same from the bicubic B-spline of the packaged lookup table
 * (LookupTable radiative.py:1770-1797 = FITPACK bispev with clamping):
 * tx[nx], ty[ny] knots, c[(nx-4)*(ny-4)] coefficients. */
int nb_pp_lut_table(const double* tx, int nx, const double* ty, int ny, const double* c,
                    const double* Ep_GeV, int N, const double* Eg_GeV, int N_E, double* K,
                    int pitch, int row0, void* stream);

/* lrs[r][j] = ln(K[r][j+1]/K[r][j]) * invdlx[j]   (j < N-1; NaN for sign changes,
 * which selects trapz_loglog's log branch, utils.py:341-345; 2^600 when either end
 * point is zero: such an interval is zero whatever the slope, utils.py:347, and the
 * finite sentinel lets the lean cell of nb_contract_ex do without a zero test). */
int nb_table_finalize(const double* K, int R, int N, int pitch, const double* invdlx,
                      double* lrs, void* stream);

/* --- the hot contraction ---------------------------------------------------
 * out[w][r] = coef[r] * trapz_loglog(n[w,:] * K[r,:], x)     utils.py:285-355
 * for all W walkers and R rows in one launch.  dlx[j] = ln(x[j+1]/x[j]).
 * coef may be NULL (=1).  `exact` != 0 evaluates every interval in the
 * reference's own operation order (log10/pow per interval): then `xn` must
 * hold n[w][j] itself (nraw of nb_pd_prep_ex), `xgrid` the grid x[N], and
 * lrs/ds1/dlx are ignored (may be NULL).
 * K_wstride: 0 for a shared table, else the per-walker table stride in
 * doubles (self-Compton). */
int nb_contract(const double* K, const double* lrs, int R, int N, int pitch,
                long long K_wstride, const double* xn, const double* ds1, int wpitch, int W,
                const double* dlx, const double* xgrid, const double* coef, double* out,
                int exact, void* stream);

/* First non-zero node of every row (row_j0[R], int32; N for an all-zero row) and
 * *flags |= 1 (int32, zero it first) when the table holds a negative or non-finite entry. */
int nb_table_scan(const double* K, int R, int N, int pitch, int* row_j0, int* flags,
                  void* stream);

/* nb_contract with the table's row_j0 (may be NULL): a row tile starts integrating at its
 * first live column (leading zeros -- kinematic limits, thresholds -- are skipped exactly),
 * and mode 0: careful cell (every trapz_loglog edge case per interval), 1: reference
 * operation order (= nb_contract exact), 2: lean cell -- 8 fp64 instructions per interval,
 * irregular slopes (|b+1| <= 1e-10, NaN, inf) detected per (walker, row tile) and
 * re-integrated with the careful cell; requires lrs from nb_table_finalize (zero end
 * points carry its finite sentinel slope) and a table without negative entries. */
int nb_contract_ex(const double* K, const double* lrs, int R, int N, int pitch,
                   const int* row_j0, const double* xn, const double* ds1, int wpitch, int W,
                   const double* dlx, const double* xgrid, const double* coef, double* out,
                   int mode, void* stream);

/* --- synchrotron, fused (Synchrotron._spectrum radiative.py:282-342) -------
 * out[w][e] = spectrum in 1/(s eV) for B[w] (Gauss) on grid gam[N] with
 * operands xn/ds1 from nb_pd_prep; E_erg[N_E] photon energies in erg.
 * gm2[j] = gam[j]^-2 and g23[j] = cbrt(gam[j]^-2) are the grid's walker-independent
 * node tables (both NULL: computed per node in the kernel); cbrtE[e] = cbrt(E_erg[e]) (NULL:
 * computed per lane); out_ld >= N_E is the row pitch of out (0 = N_E). */
int nb_synchrotron(const double* gam, int N, const double* gm2, const double* g23,
                   const double* xn, const double* ds1, int wpitch, const double* invdlx,
                   const double* dlx, const double* B, int W, const double* E_erg,
                   const double* cbrtE, int N_E, double* out, int out_ld, void* stream);

/* --- combine + likelihood (BaseRadiative.flux radiative.py:102-111,
 * lnprobmodel/lnprob core.py:64-121) ------------------------------------------
 * model[w][e] = unit_fac[e] * sum_groups ( (sum_{terms in group} wscale[w] * src[w][off+e]) / div_g )
 * Terms and groups are summed in the order given.  */
typedef struct nb_term {
  const double* src;    /* [W][ld] rows produced by nb_contract / nb_synchrotron */
  const double* wscale; /* NULL, or a per-walker factor [W] (fitted nh, n0, ...) */
  int ld;               /* row length of src */
  int off;              /* first row of this component */
  int group_end;        /* != 0: close the group after this term, dividing by div */
  double div;           /* 4 pi d^2, or 1 */
} nb_term;

/* flux_model[W][flux_ld] (may be NULL; flux_ld >= N_E is the row pitch, 0 = N_E) receives
 * the model in data units in its first N_E columns;
 * lnp[W] (may be NULL) receives lnprob:
 *   sum_{!ul} -(m-f)^2 / (2 s^2), s = err_hi if m > f else err_lo,
 *   + n_viol * ln(1 - cl[n_viol]) when the table has upper limits,
 *   + prior[w] (prior == NULL: 0); if prior[w] is +-inf the result is prior[w].
 * ul is int32[N_E]. */
int nb_combine_lnprob(const nb_term* terms_host, int n_terms, int W, int N_E,
                      const double* unit_fac, const double* data_flux, const double* err_lo,
                      const double* err_hi, const int* ul, const double* cl,
                      const double* prior, double* flux_model, int flux_ld, double* lnp,
                      void* stream);

/* --- parameter map + priors ------------------------------------------------
 * What the user's model(pars, data) / lnprior(pars) callbacks do on the host in
 * the reference (examples/RXJ1713_SynIC.py:19-62, core.py:34-58), declaratively:
 *   out[dst_off_k + w * dst_stride_k] = scale_k * f_k(pars[w][src_k])
 *                                       (src_k < 0: the constant scale_k)
 *   f: NB_FN_ID x, NB_FN_POW10 10**x, NB_FN_EXP e**x
 * so one launch fills every [W][NB_PD_MAXPAR] particle-distribution block
 * (stride NB_PD_MAXPAR) and every per-walker scalar column (stride 1) of `out`.
 *   prior[w] = sum over entries, in order, of
 *     NB_PRIOR_UNIFORM      a <= v <= b ? 0 : -inf                   core.py:34-39
 *     NB_PRIOR_NORMAL       -0.5*(2 pi b) - (v-a)^2/(2 b)  (literal) core.py:42-44
 *     NB_PRIOR_LOGUNIFORM   v > 0 && v >= a && v <= b ? 1/v : -inf   core.py:47-58
 * map_host / priors_host are HOST arrays (copied into the launch).  prior_out may
 * be NULL.  Evaluated by nb_walker_prep[_move] and the self-contained kernels. */
#define NB_FN_ID 0
#define NB_FN_POW10 1
#define NB_FN_EXP 2
#define NB_PRIOR_UNIFORM 0
#define NB_PRIOR_NORMAL 1
#define NB_PRIOR_LOGUNIFORM 2
#define NB_MAX_MAP 32
#define NB_MAX_PRIORS 16
typedef struct nb_parmap {
  int src;
  int fn;
  double scale;
  long long dst_off;
  int dst_stride;
} nb_parmap;
typedef struct nb_prior {
  int par;
  int kind;
  double a, b;
} nb_prior;

/* --- per-walker set-up, fused ----------------------------------------------------
 * nb_param_map, then for every job the nb_pd_prep operands of one particle
 * distribution on one grid and/or its total particle energy (nb_particle_energy), in
 * ONE launch (one CTA per walker): what BaseElectron._nelec / _gam / compute_We
 * (radiative.py:147-195) and the user's model() parameter arithmetic do per lnprob call.
 * `pm` is the nb_param_map output buffer; job.pd_off is the offset (in doubles) of the
 * job's [W][NB_PD_MAXPAR] parameter block inside it.  jobs_host is a HOST array. */
#define NB_MAX_PREP_JOBS 8
typedef struct nb_prep_job {
  int kind;             /* NB_PD_* */
  int N;                /* grid nodes */
  long long pd_off;
  const double* x;      /* grid [N] */
  const double* invdlx; /* [N-1]; may be NULL when xn == NULL */
  double e_mul1, e_mul2, n_scale;
  double* xn;           /* [W][wpitch] or NULL (energy only) */
  double* ds1;
  double* nraw;         /* may be NULL */
  int wpitch;
  int pad_;
  double x_to_energy;
  double* energy_out;   /* element w at energy_out[w * energy_stride], or NULL */
  long long energy_stride; /* 0 = 1 */
} nb_prep_job;
int nb_walker_prep(const double* pars, int W, int P, const nb_parmap* map_host, int n_map,
                   double* pm, const nb_prior* priors_host, int n_priors, double* prior_out,
                   const nb_prep_job* jobs_host, int n_jobs, void* stream);

/* --- IC on a tabulated seed, fused (synchrotron self-Compton) ----------------
 * radiative.py:609-655 + 684 in one launch, reference operation order:
 *   out[w][out_off + e] = Eph[e] * trapz_loglog(nraw[w,:] * K_w[e,:], gam)
 *   K_w[e,j] = 3/4 sigma_T c / gam_j^2 *
 *              trapz_loglog_s(f_AA81(gam_j, eps0_s, Eph_e) * phn[w][s] / eps0_s, eps0)
 * phn_wstride: Ns for a per-walker seed density phn[W][Ns], 0 for a shared one. */
int nb_ic_seed_spectrum(const double* gam, int N, const double* nraw, int wpitch,
                        const double* Eph, int N_E, const double* eps0, const double* phn,
                        int Ns, int phn_wstride, int W, double* out, int out_ld, int out_off,
                        void* stream);

/* --- synchrotron self-Compton: IC on a per-walker tabulated seed, hoisted ----------
 * The same integral as nb_ic_seed_spectrum (radiative.py:609-655 + 684) for seed
 * densities that differ per walker (examples/CrabNebula_SynSSC.py:24-28), with the
 * walker-independent factor f_AA81(gam_g, eps0_s, Eph_e) tabulated once:
 *   nb_ssc_table   KL[s][r] = (F[s+1][r], log-slope of F over interval s; sentinel for zero
 *                  end points) as interleaved pairs of doubles, s < Ns-1, and F0[r] = F[0][r];
 *                  r = e*N + g, row pitch Rp (multiple of 128, >= N_E*N); coef[r] =
 *                  3/4 sigma_T c / gam_g^2.  invdlx_s[s] = 1/ln(eps0[s+1]/eps0[s]).
 *   nb_ssc_seed    sxn[w][s] = sum_k fac_k * src_k[w][off_k + s]  (dn/dE in 1/(mec2 cm3)
 *                  from luminosities in 1/(s eV): Lsy / (4 pi R^2 c) * 2.24 * mec2[eV]),
 *                  sds[w][s] = its slope term ln(sxn[s+1]/sxn[s]) * invdlx_s[s].
 *   nb_ssc_inner   inner[w][r] = coef[r] * trapz_loglog_s(F[:, r] * sxn[w, :] / eps0, eps0)
 *                  -- lean cell, rows with an irregular slope redone with the careful cell.
 *   nb_ssc_outer   out[w][out_off + e] = coef_e[e] * trapz_loglog_g(n_e[w, :] * inner[w, e, :],
 *                  gam) with the electron operands xn / ds1 of nb_pd_prep; coef_e = Eph/E_eV. */
#define NB_SSC_MAX_SRC 4
typedef struct nb_ssc_src {
  const double* src; /* [W][ld] */
  int ld, off;
  double fac;
} nb_ssc_src;
int nb_ssc_table(const double* gam, int N, const double* Eph, int N_E, const double* eps0,
                 const double* invdlx_s, int Ns, double* KL, double* F0, double* coef,
                 long long Rp, void* stream);
int nb_ssc_seed(const nb_ssc_src* src_host, int n_src, int W, int Ns, const double* invdlx_s,
                double* sxn, double* sds, int spitch, void* stream);
int nb_ssc_inner(const double* KL, const double* F0, const double* coef, long long Rp, int Ns,
                 const double* sxn, const double* sds, int spitch, int W, const double* dlx_s,
                 double* inner, void* stream);
int nb_ssc_outer(const double* inner, long long Rp, int N, int N_E, int W, const double* xn,
                 const double* ds1, int wpitch, const double* dlx, const double* invdlx,
                 const double* coef_e, double* out, int out_ld, int out_off, void* stream);

/* --- ensemble stretch move (emcee StretchMove, driven from core.py:127-160) --
 * Device-resident: the draws of n_steps ensemble steps are resident as s_idx/c_idx
 * [n_steps][2][Ns] (int32), zz/lnu [n_steps][2][Ns]; the current step t is read from
 * *step (device int32).  The whole red-blue half-step is three launches --
 *   nb_walker_prep_move  (proposal q[i][:] = c[i][:] - (c[i][:] - s[i][:]) * zz[i] of
 *                         the active half + nb_walker_prep: the set-up CTAs compute the
 *                         proposals themselves and publish them to `pars`)
 *   nb_contract / nb_synchrotron_fused ... (independent, may run concurrently)
 *   nb_combine_lnprob_update (nb_combine_lnprob + accept step: proposal i is accepted when
 *                         (P-1) ln zz + new_lp - lp[s_idx] > ln u, overwriting the
 *                         walker's coords / lp / blob record and counting in n_accepted;
 *                         the walker's row of step t is appended to chain / chain_lp /
 *                         chain_blobs by the warp that decides its proposal; a NaN new_lp
 *                         is never accepted but IS written to chain_lp)
 * `step` is read at kernel start by all kernels of a step and incremented by the last
 * CTA of the split == 1 combine kernel to finish (ticket in `sync`, an int32 scratch
 * word that must be zero before the first launch). */
#define NB_TIMELINE_CAP 8192 /* half-steps kept by nb_stretch.timeline (a ring) */
#define NB_TIMELINE_COLS 16  /* stamps per half-step */
typedef struct nb_stretch {
  double* coords;      /* [W][P] */
  double* lp;          /* [W] */
  double* blobs;       /* [W][nb] or NULL */
  int nb, W, P, Ns, split;
  int i0;              /* first proposal of the half handled by this launch (walker sharding:
                          a rank evaluates proposals [i0, i0 + W_launch) of the Ns) */
  int pars_ld;         /* row pitch of the published proposals `pars` (0 = P) */
  int pad_;
  int* step;           /* device int32: current step index t */
  int* sync;           /* device int32 scratch (ticket counter) */
  const int* s_idx;    /* [n_steps][2][Ns] */
  const int* c_idx;
  const double* zz;
  const double* lnu;
  int* n_accepted;     /* [W] */
  double* chain;       /* [n_steps][W][P] or NULL */
  double* chain_lp;    /* [n_steps][W] or NULL */
  double* chain_blobs; /* [n_steps][W][nb] or NULL */
  /* replicated state updated by peers (nb_combine_lnprob_update_push): kernels that read
   * `coords` first wait until wait_flags[0 .. wait_world) have all reached *wait_gen */
  const unsigned long long* wait_flags; /* NULL: no waiting */
  const unsigned long long* wait_gen;
  int wait_world;
  int pad2_;
  /* optional diagnostic (NULL in production): [NB_TIMELINE_CAP][NB_TIMELINE_COLS] %globaltimer
   * stamps (ns), row (2 * *step + split) mod NB_TIMELINE_CAP of the half-step:
   * [0] first kernel entered (CTA 0 of nb_walker_prep_move), [1] its wait for the peers' flags is
   * over, [2] accept kernel entered (CTA 0), [3] last CTA before it releases this rank's flag
   * (sharded runs), [4] last CTA of the accept kernel done, [5] / [6] synchrotron kernel: CTA 0
   * entered / last CTA done, [7] last CTA of the set-up launch done, [8] / [9] total-energy blob
   * launch, [10] / [11] contraction kernel(s): first CTA entered / last CTA done */
  unsigned long long* timeline;
} nb_stretch;
int nb_walker_prep_move(const nb_stretch* mv_host, double* pars, int W, int P,
                        const nb_parmap* map_host, int n_map, double* pm,
                        const nb_prior* priors_host, int n_priors, double* prior_out,
                        const nb_prep_job* jobs_host, int n_jobs, void* stream);
/* lnp is required; when mv.nb > 0 so is flux_model, whose rows [Ns][flux_ld] are the
 * per-walker blob records (model flux in the first N_E columns, further blobs written
 * there by earlier kernels): the first mv.nb columns (N_E <= mv.nb <= flux_ld) of an
 * accepted proposal's row replace the walker's blob record.  `pars` holds the
 * proposals published by nb_walker_prep_move. */
int nb_combine_lnprob_update(const nb_stretch* mv_host, const double* pars,
                             const nb_term* terms_host, int n_terms, int W, int N_E,
                             const double* unit_fac, const double* data_flux,
                             const double* err_lo, const double* err_hi, const int* ul,
                             const double* cl, const double* prior, double* flux_model,
                             int flux_ld, double* lnp, void* stream);
/* nb_combine_lnprob with a stride between consecutive walkers' lnp (lnp[w * lnp_ld]), for
 * writing straight into packed per-proposal records */
int nb_combine_lnprob_ld(const nb_term* terms_host, int n_terms, int W, int N_E,
                         const double* unit_fac, const double* data_flux, const double* err_lo,
                         const double* err_hi, const int* ul, const double* cl,
                         const double* prior, double* flux_model, int flux_ld, double* lnp,
                         int lnp_ld, void* stream);

/* Walker sharding: every rank evaluates a slice of the half-ensemble's proposals into
 * packed records  pack[i][0..nb) = blob record, pack[i][nb] = lnprob,
 * pack[i][nb+1 .. nb+1+P) = proposal  (row pitch ld >= nb + 1 + P), the slices are
 * all-gathered, and this kernel then runs the accept step + chain append for ALL Ns
 * proposals of the half identically on every rank (mv.i0 is ignored).  Step counter
 * protocol as for nb_combine_lnprob_update. */
int nb_stretch_update_packed(const nb_stretch* mv_host, const double* pack, int ld,
                             void* stream);

/* --- walker sharding without a collective launch: replicated state over NVLink --------
 * The ensemble state and the chain live in buffers that every rank has mapped from every
 * peer (symmetric memory).  See nb_combine_lnprob_update_push below. */
#define NB_MAX_PEERS 16
typedef struct nb_peers {
  int world, rank;
  int i0;                               /* unused (reserved) */
  int ld;                               /* unused (reserved) */
  double* pack[NB_MAX_PEERS];           /* unused (reserved) */
  unsigned long long* flags[NB_MAX_PEERS]; /* peer r's flag array [world] */
  unsigned long long* gen;              /* this rank's half-step generation counter */
  int* ticket;                          /* int32 scratch, zero before the first launch */
  double* mc_pack;                      /* unused (reserved) */
  /* replicated-state mode (nb_combine_lnprob_update_push): up to two symmetric arenas that
   * hold the ensemble state and the chain on every rank; a local pointer inside arena k
   * maps to arena_peer[k][r] + offset on rank r and to arena_mc[k] + offset for a
   * multicast store (arena_mc[k] == NULL: one store per peer) */
  char* arena_local[2];
  char* arena_mc[2];
  char* arena_peer[2][NB_MAX_PEERS];
  unsigned long long arena_bytes[2];
  unsigned long long* mc_flags;         /* multicast address of the flag arrays, or NULL */
} nb_peers;
/* Replicated-state sharding: nb_combine_lnprob_update for this rank's proposals
 * [mv.i0, mv.i0 + W) whose accept step writes the walkers' new state (coords, lp, blob
 * record) and their chain rows into EVERY rank's copy (n_accepted stays local to the
 * deciding rank: sum it over ranks on the host) (multimem.st
 * through the NVSwitch, or one store per peer), then raises flags[rank] = *gen + 1 on every
 * rank and advances *gen.  The next half-step's kernels wait on the flags through
 * nb_stretch.wait_*: no collective and no separate accept kernel -- the sharded step has
 * exactly the launches of the single-GPU step. */
/* One-CTA kernel that returns once every rank's pushes of all completed half-steps have
 * landed in this rank's copy (mv.wait_*): enqueue it before reading the replicated state
 * or the chain from the host or with a copy. */
int nb_peer_wait(const nb_stretch* mv_host, void* stream);
int nb_combine_lnprob_update_push(const nb_stretch* mv_host, const nb_peers* peers_host,
                                  const double* pars, const nb_term* terms_host, int n_terms,
                                  int W, int N_E, const double* unit_fac,
                                  const double* data_flux, const double* err_lo,
                                  const double* err_hi, const int* ul, const double* cl,
                                  const double* prior, double* flux_model, int flux_ld,
                                  double* lnp, void* stream);

/* --- self-contained component kernels -------------------------------------------
 * nb_synchrotron with the walker's operands derived INSIDE the kernel from
 * the raw parameters (each warp / CTA repeats the few-hundred-instruction parameter map
 * and evaluates the particle distribution at the nodes it integrates), so that a
 * likelihood evaluation needs no set-up launch in front of its components.
 *   src: where the parameters come from -- pars[W][P] (dense), or, with mv != NULL, the
 *        stretch-move proposals of the active half (as nb_walker_prep_move);
 *        map_host/n_map as nb_param_map (dst_off / dst_stride identify the entries:
 *        parameter k of the distribution is the entry with dst_off == pd_off + k and
 *        dst_stride == NB_PD_MAXPAR).  P <= 32 and n_map <= 32.
 *   pd:  the particle distribution and the grid's walker-independent tables
 *        (lnx[j] = ln x[j]). */
typedef struct nb_walker_src {
  const double* pars;
  int P;
  int n_map;
  const nb_parmap* map_host;
  const nb_stretch* mv_host; /* NULL: parameters from `pars` */
} nb_walker_src;
typedef struct nb_pd_desc {
  int kind;          /* NB_PD_* */
  int pad_;
  long long pd_off;
  double e_mul1, e_mul2, n_scale;
  const double* lnx;    /* [N] */
  const double* invdlx; /* [N-1] */
} nb_pd_desc;
/* b_entry: index of the map entry holding B [G]; gm2 / g23 / cbrtE / out_ld as nb_synchrotron */
int nb_synchrotron_fused(const nb_walker_src* src, const nb_pd_desc* pd, int b_entry,
                         const double* gam, int N, const double* gm2, const double* g23,
                         const double* dlx, int W, const double* E_erg, const double* cbrtE,
                         int N_E, double* out, int out_ld, void* stream);

/* --- pion decay, Kelner+06 (PionDecayKelner06 radiative.py:1543-1767) ----------------
 * The reference integrates KAB06 Eq. 71 (photon energies >= Etrans) and the delta-functional
 * approximation Eq. 78 (below) with adaptive QUADPACK at epsrel = 1e-3, one Python callback
 * per sample point.  Here both are log-log trapezoids over a per-row proton-energy grid
 * (100 nodes per decade from the row's own lower limit):
 *   nb_kelner_table  for row r (photon energy Eg_TeV[r]; hi[r] != 0: full calculation, else
 *                    delta-functional): Ep[r][j] = E0_r 10^(decades j/(N-1)) and the
 *                    walker-independent integrand kernel Kk[r][j]
 *                    (c sigma_inel F_gamma(Eg/Ep, Ep)/Ep, or 2 c sigma_inel/sqrt(Epi^2-m_pi^2)).
 *   nb_kelner_rows   out[w][r] = trapz_loglog(J_w(Ep[r,:]) Kk[r,:], Ep[r,:]) in 1/(s TeV)
 *                    for n_H = 1 and nhat = 1, J = PD.eval per TeV. */
int nb_kelner_table(const double* Eg_TeV, const int* hi, int R, int N, double decades,
                    double* Ep, double* Kk, void* stream);
int nb_kelner_rows(int kind, const double* pd_params, int W, const double* Ep, const double* Kk,
                   int R, int N, double* out, void* stream);

/* --- measurement aid: how often the lean cell fell back to the careful cell ----------
 * out_host[0]: (walker, row tile) pairs of nb_contract_ex mode 2 that were re-integrated,
 * out_host[1]: rows of nb_ssc_inner that were (each for all walkers of its thread), since
 * the library was loaded or the last reset.  Synchronous (cudaMemcpyFromSymbol). */
int nb_fallback_counts(unsigned long long* out_host, int reset);
/* Diagnostic timelines (nb_stretch.timeline): call before freeing a timeline buffer -- the
 * contraction kernel reaches the current row through a device-global pointer that the set-up
 * kernel publishes. */
int nb_timeline_reset(void);
/* Preferred shared-memory carve-out (percent of the SM's L1 / shared-memory array; -1: the
 * driver's choice, the default) that the CALLING THREAD's subsequent launches of the set-up,
 * contraction, synchrotron and combine kernels carry as a launch attribute (recorded in the
 * kernel node when a stream is capturing).  Kernels that run on parallel branches of one
 * evaluation only share an SM if they ask for the same split (an SM reconfigures only when
 * idle): a plan with a synchrotron and a tabulated component sets 50 around its launches. */
int nb_launch_carveout(int percent);

/* --- measurement aid: fp64 FMA throughput probe -------------------------------
 * Runs blocks x threads threads doing iters x 16 dependent-chain-free DFMAs each;
 * the caller times it with CUDA events to obtain the fp64 roofline denominator. */
int nb_fp64_peak_probe(double* out, int blocks, int threads, int iters, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NAIMA_B200_H */
