// emu.cpp -- TEST-ONLY host build of naima_b200/csrc/nb_math.cuh.
//
// The per-cell arithmetic of the CUDA kernels lives in `__host__ __device__`
// functions; this file compiles them with g++ and emulates the kernels' thread
// decomposition (lane ranges, shuffle-tree order) in plain loops so that the
// formulas can be checked against the oracle on a box without a GPU
// (tests/test_host_emu.py).  It is never loaded by the product: naima_b200 has
// no CPU execution path.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "../../naima_b200/csrc/nb_math.cuh"

using namespace nb;

static double tree32(const double* v) {
  // __shfl_down_sync tree of warp_sum(): offsets 16, 8, 4, 2, 1; lane 0 result
  double t[32];
  memcpy(t, v, sizeof(t));
  for (int o = 16; o > 0; o >>= 1)
    for (int l = 0; l < 32; ++l) t[l] = t[l] + ((l + o < 32) ? t[l + o] : t[l]);
  return t[0];
}

extern "C" {

void emu_pd_eval(int kind, const double* p, const double* e, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = pd_eval(kind, p, e[i]);
}

void emu_interval_exact(const double* x, const double* y, int n, double* out) {
  for (int i = 0; i < n - 1; ++i) out[i] = interval_exact(x[i], x[i + 1], y[i], y[i + 1]);
}

void emu_gtilde(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = gtilde(x[i]);
}

void emu_ic_planck(const double* gam, int N, const double* Eph, int N_E, double T,
                   double theta, double* out) {
  for (int e = 0; e < N_E; ++e)
    for (int j = 0; j < N; ++j)
      out[e * N + j] = (theta != theta) ? ic_iso_planck(gam[j], T, Eph[e])
                                        : ic_ani_planck(gam[j], T, Eph[e], theta);
}

void emu_ic_seed(const double* gam, int N, const double* Eph, int N_E, const double* eps0,
                 const double* phn, int Ns, double* out) {
  for (int e = 0; e < N_E; ++e)
    for (int j = 0; j < N; ++j) {
      double g = gam[j], ep = Eph[e], v;
      if (Ns == 1) {
        v = ic_mono_f(g, eps0[0], ep);
        v *= phn[0] / (eps0[0] * eps0[0]);
      } else {
        double x1 = eps0[0];
        double y1 = ic_mono_f(g, x1, ep) * phn[0] / x1;
        double acc = 0.0;
        for (int s = 1; s < Ns; ++s) {
          double x2 = eps0[s];
          double y2 = ic_mono_f(g, x2, ep) * phn[s] / x2;
          acc += interval_exact(x1, x2, y1, y2);
          x1 = x2;
          y1 = y2;
        }
        v = acc;
      }
      v *= (3.0 / 4.0) * SIGT * 29979245800.0 / (g * g);
      out[e * N + j] = v;
    }
}

void emu_brems(const double* gam, int N, const double* eps, int N_E, double* see, double* s1) {
  for (int e = 0; e < N_E; ++e)
    for (int j = 0; j < N; ++j) {
      see[e * N + j] = brems_sigma_ee(gam[j], eps[e]) / MEC2_EV;
      s1[e * N + j] = brems_sigma_1(gam[j], eps[e]);
    }
}

void emu_pp_diffsigma(const double* Ep, int N, double Eg, int model, int nuc, double* out) {
  for (int j = 0; j < N; ++j) out[j] = pp_diffsigma(Ep[j], Eg, model, nuc);
}

void emu_bspl(const double* tx, int nx, const double* ty, int ny, const double* c,
              const double* x, int n, double y, double* out) {
  for (int i = 0; i < n; ++i) out[i] = bspl_eval2d(tx, nx, ty, ny, c, x[i], y);
}

void emu_exp_neg(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = exp_neg(x[i]);
}

// pd_prep_chunk, reference-order branch (nraw != NULL): feeds the exact contraction
void emu_pd_prep(int kind, const double* p, const double* x, int N, double m1, double m2,
                 double ns, const double* invdlx, double* xn, double* ds1, double* nraw) {
  for (int j = 0; j < N; ++j) {
    nraw[j] = pd_eval(kind, p, (x[j] * m1) * m2) * ns;
    xn[j] = x[j] * nraw[j];
  }
  for (int j = 0; j < N - 1; ++j) ds1[j] = log(nraw[j + 1] / nraw[j]) * invdlx[j] + 1.0;
  ds1[N - 1] = 0.0;
}

// pd_prep_chunk, log-space branch: feeds the hoisted contraction and the synchrotron kernel
void emu_pd_prep_log(int kind, const double* p, const double* x, int N, double m1, double m2,
                     double ns, const double* invdlx, double* xn, double* ds1, double* n) {
  PdLog S = pd_log_setup(kind, p, ns);
  PdNode* nd = (PdNode*)malloc(sizeof(PdNode) * N);
  for (int j = 0; j < N; ++j) {
    double e = (x[j] * m1) * m2;
    nd[j] = pd_log_node(S, e);
    n[j] = pd_log_value(S, nd[j]);
    xn[j] = x[j] * n[j];
  }
  for (int j = 0; j < N - 1; ++j) ds1[j] = pd_log_ds1(S, nd[j], nd[j + 1], invdlx[j]);
  ds1[N - 1] = 0.0;
  free(nd);
}

// table_finalize_kernel
void emu_finalize(const double* K, int R, int N, int pitch, const double* invdlx, double* lrs) {
  for (int r = 0; r < R; ++r)
    for (int j = 0; j < pitch; ++j)
      lrs[r * pitch + j] = (j < N - 1) ? slope_or_sentinel(K[r * pitch + j], K[r * pitch + j + 1],
                                                           invdlx[j])
                                       : 0.0;
}

// contract_kernel<RT=1>, one walker; lane decomposition + shuffle tree.  mode 0: careful
// cell, 1: reference operation order, 2: lean cell with the per-row fall-back; skip != 0:
// start at the row's first live column like the kernel does with row_j0
void emu_contract(const double* K, const double* lrs, int R, int N, int pitch, const double* xn,
                  const double* ds1, const double* dlx, const double* xgrid, int mode, int skip,
                  double* out, int* n_fallback) {
  int nint = N - 1;
  if (n_fallback) *n_fallback = 0;
  for (int r = 0; r < R; ++r) {
    const double* Kr = K + (size_t)r * pitch;
    const double* Lr = lrs + (size_t)r * pitch;
    int jt = 0;
    if (skip) {
      int j0 = N;
      for (int j = 0; j < N; ++j)
        if (Kr[j] != 0.0) { j0 = j; break; }
      jt = (j0 - 1 > 0 ? j0 - 1 : 0) & ~1;
    }
    if (jt >= nint) { out[r] = 0.0; continue; }
    int m = odd_chunk(nint - jt);
    double part[32];
    unsigned worst[32];
    for (int pass = 0; pass < 2; ++pass) {
      bool redo = false;
      for (int lane = 0; lane < 32; ++lane) {
        int i0 = jt + lane * m, i1 = i0 + m < nint ? i0 + m : nint;
        double acc = 0.0;
        worst[lane] = 0u;
        if (i0 < nint) {
          if (mode == 1)
            contract_lane_exact<1>(xn, xgrid, Kr, pitch, i0, i1, &acc);
          else if (mode == 2 && pass == 0)
            worst[lane] = (m >= 16 ? contract_lane_lean<1, 4>(xn, ds1, Kr, Lr, pitch, i0, i1, &acc)
                                   : contract_lane_lean<1, 1>(xn, ds1, Kr, Lr, pitch, i0, i1, &acc));
          else
            contract_lane_fast<1, 1>(xn, ds1, dlx, Kr, Lr, pitch, i0, i1, &acc);
        }
        part[lane] = acc;
        redo = redo || worst[lane] >= NB_REG_RANGE;
      }
      if (!redo) break;
      if (n_fallback && pass == 0) ++*n_fallback;
    }
    out[r] = tree32(part);
  }
}

// synchrotron_kernel for one walker
void emu_synchrotron(const double* gam, int N, const double* xn, const double* ds1,
                     const double* invdlx, const double* dlx, double B, const double* E_erg,
                     int N_E, double* out) {
  double* iec = (double*)malloc(sizeof(double) * N);
  double* cb = (double*)malloc(sizeof(double) * N);
  double ikB, cbk;
  syn_walker(B, &ikB, &cbk);
  for (int j = 0; j < N; ++j) {  // the kernels' node tables gm2 = g^-2, g23 = cbrt(g^-2)
    double gm2 = 1.0 / (gam[j] * gam[j]);
    iec[j] = ikB * gm2;
    cb[j] = cbk * cbrt(gm2);
  }
  int nint = N - 1;
  for (int e = 0; e < N_E; ++e) {
    double tot = 0.0;
    int js = syn_first_node(gam, N, B, E_erg[e]);
    int len = nint - js;
    if (len > 0) {
      int m = odd_chunk2(len);
      double halves[2];
      for (int half = 0; half < 2; ++half) {
        double part[32];
        for (int lane = 0; lane < 32; ++lane) {
          int i0 = js + (half * 32 + lane) * m, i1 = i0 + m < nint ? i0 + m : nint;
          part[lane] = (i0 < nint) ? syn_lane(E_erg[e], cbrt(E_erg[e]), iec, cb, xn, ds1, invdlx,
                                              dlx, i0, i1)
                                   : 0.0;
        }
        halves[half] = tree32(part);
      }
      tot = halves[0] + halves[1];
    }
    out[e] = syn_finish(B, E_erg[e], tot);
  }
  free(iec);
  free(cb);
}

// synchrotron_fused_kernel for one walker: node set-up from the ln x table, then the
// same pair-of-warps integration as emu_synchrotron
void emu_synchrotron_fused(int kind, const double* p, double m1, double m2, double ns,
                           const double* gam, const double* lnx, int N, const double* invdlx,
                           const double* dlx, double B, const double* E_erg, int N_E,
                           double* out) {
  PdLog S = pd_log_setup(kind, p, ns);
  pd_log_setup_grid(S, m1, m2);
  double* xn = (double*)malloc(sizeof(double) * N);
  double* ds1 = (double*)malloc(sizeof(double) * N);
  for (int j = 0; j < N; ++j) {
    PdNode nd = pd_log_node_tab(S, gam[j], lnx[j]);
    xn[j] = gam[j] * pd_log_value_fast(S, nd);
    ds1[j] = 0.0;
    if (j < N - 1) ds1[j] = pd_log_ds1(S, nd, pd_log_node_tab(S, gam[j + 1], lnx[j + 1]), invdlx[j]);
  }
  emu_synchrotron(gam, N, xn, ds1, invdlx, dlx, B, E_erg, N_E, out);
  free(xn);
  free(ds1);
}

// ssc_table_kernel + ssc_seed_kernel + ssc_inner_kernel + ssc_outer_kernel for one walker:
// phn[Ns] seed density in 1/(mec2 cm3), xn/ds1 the electron operands on gam[N];
// out[e] = Eph/E * integral (E_eV given for the last factor); n_fallback counts rows redone
// with the careful cell
void emu_ssc(const double* gam, int N, const double* Eph, const double* E_eV, int N_E,
             const double* eps0, const double* phn, int Ns, const double* xn, const double* ds1,
             const double* dlx, const double* invdlx, double* out, int* n_fallback) {
  double* dls = (double*)malloc(sizeof(double) * Ns);
  double* inv = (double*)malloc(sizeof(double) * Ns);
  double* sds = (double*)malloc(sizeof(double) * Ns);
  double* F = (double*)malloc(sizeof(double) * Ns);
  double* L = (double*)malloc(sizeof(double) * Ns);
  double* inner = (double*)malloc(sizeof(double) * N);
  for (int s = 0; s < Ns - 1; ++s) {
    dls[s] = log(eps0[s + 1] / eps0[s]);
    inv[s] = 1.0 / dls[s];
    sds[s] = slope_or_sentinel(phn[s], phn[s + 1], inv[s]);
  }
  if (n_fallback) *n_fallback = 0;
  for (int e = 0; e < N_E; ++e) {
    for (int j = 0; j < N; ++j) {
      for (int s = 0; s < Ns; ++s) F[s] = ic_mono_f(gam[j], eps0[s], Eph[e]);
      for (int s = 0; s < Ns - 1; ++s) L[s] = slope_or_sentinel(F[s], F[s + 1], inv[s]);
      double acc = 0.0, prev = phn[0] * F[0];
      unsigned worst = 0u;
      for (int s = 0; s < Ns - 1; ++s) {
        double xy2 = phn[s + 1] * F[s + 1];
        cell_lean(prev, xy2, sds[s] + L[s], acc, worst);
        prev = xy2;
      }
      if (worst >= NB_REG_RANGE) {
        if (n_fallback) ++*n_fallback;
        acc = 0.0;
        double xy1 = phn[0] * F[0];
        for (int s = 0; s < Ns - 1; ++s) {
          double xy2 = phn[s + 1] * F[s + 1];
          acc += interval_fast(xy1, xy2, sds[s] + L[s], dls[s]);
          xy1 = xy2;
        }
      }
      inner[j] = acc * ((3.0 / 4.0) * SIGT * 29979245800.0 / (gam[j] * gam[j]));
    }
    double part[32];
    for (int lane = 0; lane < 32; ++lane) {
      double acc = 0.0;
      for (int j = lane; j < N - 1; j += 32) {
        double bp1 = ds1[j] + log(inner[j + 1] / inner[j]) * invdlx[j];
        acc += interval_fast(xn[j] * inner[j], xn[j + 1] * inner[j + 1], bp1, dlx[j]);
      }
      part[lane] = acc;
    }
    out[e] = Eph[e] / E_eV[e] * tree32(part);
  }
  free(dls); free(inv); free(sds); free(F); free(L); free(inner);
}

// kelner_table_kernel + kelner_rows_kernel for one walker (serial sum instead of the CTA's
// tree: the comparison is against adaptive quadrature at 1e-3)
void emu_kelner(int kind, const double* p, const double* Eg, const int* hi, int R, int N,
                double decades, double* out) {
  double* ep = (double*)malloc(sizeof(double) * N);
  double* y = (double*)malloc(sizeof(double) * N);
  for (int r = 0; r < R; ++r) {
    double e0 = Eg[r];
    if (!hi[r]) {
      double Epimin = Eg[r] + KEL_MPI_TEV * KEL_MPI_TEV / (4 * Eg[r]);
      e0 = KEL_MP_TEV + Epimin / KEL_KPI;
    }
    for (int j = 0; j < N; ++j) {
      ep[j] = e0 * pow(10.0, decades * j / (N - 1));
      double kk = hi[r] ? kel_kernel_hi(ep[j], Eg[r]) : kel_kernel_lo(ep[j]);
      y[j] = pd_eval(kind, p, ep[j] * 1e12) * 1e12 * kk;
    }
    double acc = 0.0;
    for (int j = 0; j < N - 1; ++j) acc += interval_exact(ep[j], ep[j + 1], y[j], y[j + 1]);
    out[r] = acc;
  }
  free(ep);
  free(y);
}

// combine_lnprob_kernel
void emu_combine_lnprob(const nb_term* terms, int n_terms, int W, int N_E,
                        const double* unit_fac, const double* data_flux, const double* err_lo,
                        const double* err_hi, const int* ul, const double* cl,
                        const double* prior, double* flux_model, double* lnp) {
  CombineArgs a;
  for (int t = 0; t < n_terms; ++t) a.terms[t] = terms[t];
  a.n_terms = n_terms; a.W = W; a.N_E = N_E; a.unit_fac = unit_fac;
  a.data_flux = data_flux; a.err_lo = err_lo; a.err_hi = err_hi; a.ul = ul; a.cl = cl;
  a.prior = prior; a.flux_model = flux_model; a.flux_ld = N_E; a.lnp = lnp; a.lnp_ld = 1;
  for (int w = 0; w < W; ++w) combine_lnprob_walker(a, w);
}

}  // extern "C"
