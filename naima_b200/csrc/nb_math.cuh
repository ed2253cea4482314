// nb_math.cuh -- per-cell fp64 arithmetic of the naima likelihood hot path.
//
// Everything here is `__host__ __device__` scalar code with no CUDA-only
// constructs, so the formulas can be unit-tested on the CPU box (tests compile
// this header with g++ into a throw-away checker library, tests/host_emu/) and
// are then used unchanged by the sm_100a kernels in nb_kernels.cu.  The host
// build is a TEST aid only; the product has no CPU path.
//
// Reference formulas (zblz/naima @ ba20a64, paths relative to src/naima/):
//   pd_eval            models.py:87-92,156-161,233-238,329-335,401-407
//   interval_*         utils.py:336-348   (trapz_loglog)
//   gtilde             radiative.py:300-311
//   G12 / G34          radiative.py:345-367
//   ic_iso/ani_planck  radiative.py:547-607
//   ic_mono_f          radiative.py:626-636
//   brems_*            radiative.py:838-928
//   pp_*               radiative.py:1215-1482
//   bspl_*             scipy FITPACK fpbspl/fpbisp, called at radiative.py:1793-1797
#pragma once
#include <math.h>
#include <stddef.h>
#include <string.h>

#include "../../include/naima_b200.h"

#if defined(__CUDACC__)
#define NB_HD __host__ __device__ __forceinline__
#else
#define NB_HD inline
#endif

namespace nb {

// CODATA 2018 (astropy >= 6.1), cgs
constexpr double C_CGS = 29979245800.0;
constexpr double E_ESU = 4.803204712570263e-10;
constexpr double HBAR_CGS = 1.0545718176461565e-27;
constexpr double M_E_G = 9.1093837015e-28;
constexpr double ALPHA_FS = 0.0072973525693;
constexpr double EV_ERG = 1.602176634e-12;
constexpr double MEC2_ERG = M_E_G * C_CGS * C_CGS;
constexpr double MEC2_EV = MEC2_ERG / EV_ERG;
constexpr double R0_CM = E_ESU * E_ESU / MEC2_ERG;
constexpr double MPC2_GEV = 0.9382720881604903;
constexpr double M_PI0 = 0.1349766;
constexpr double T_TH = 0.27966184;
constexpr double NB_PI = 3.141592653589793;
constexpr double PI26 = NB_PI * NB_PI / 6.0;

enum PdKind { PD_PL = 0, PD_ECPL = 1, PD_BPL = 2, PD_ECBPL = 3, PD_LOGPAR = 4 };
constexpr int PD_MAXPAR = 8;

// ---------------------------------------------------------------------------
// particle distributions: p[] in the order of the reference eval() signature
// after `e` (energies eV, amplitude 1/eV)
// ---------------------------------------------------------------------------
NB_HD double pd_eval(int kind, const double* p, double e) {
  switch (kind) {
    case PD_PL: {
      double xx = e / p[1];
      return p[0] * pow(xx, -p[2]);
    }
    case PD_ECPL: {
      double xx = e / p[1];
      return p[0] * pow(xx, -p[2]) * exp(-pow(e / p[3], p[4]));
    }
    case PD_BPL: {
      bool lo = e < p[2];
      double K = lo ? 1.0 : pow(p[2] / p[1], p[4] - p[3]);
      double a = lo ? p[3] : p[4];
      return p[0] * K * pow(e / p[1], -a);
    }
    case PD_ECBPL: {
      bool lo = e < p[2];
      double K = lo ? 1.0 : pow(p[2] / p[1], p[4] - p[3]);
      double a = lo ? p[3] : p[4];
      double ee2 = e / p[5];
      return p[0] * K * pow(e / p[1], -a) * exp(-pow(ee2, p[6]));
    }
    case PD_LOGPAR: {
      double ee = e / p[1];
      double ex = -p[2] - p[3] * log(ee);
      return p[0] * pow(ee, ex);
    }
  }
  return NAN;
}

// ---------------------------------------------------------------------------
// constants of the per-cell math.  On the device they live in constant memory so that
// DFMA/DMUL take them straight from the constant bank; as literals every use costs two
// extra moves per 64-bit constant, which was a third of the synchrotron kernel's
// instructions.
// ---------------------------------------------------------------------------
#define NB_CONST_TABLE                                                                      \
  { /* 0..12: 1/k!, k = 0..12 (exp) */                                                      \
    1.0, 1.0, 0.5, 0.16666666666666666, 0.041666666666666664,                  \
    0.008333333333333333, 0.001388888888888889, 0.0001984126984126984, 2.48015873015873e-05,             \
    2.7557319223985893e-06, 2.755731922398589e-07, 2.505210838544172e-08, 2.08767569878681e-09,                                                                                     \
    /* 13..16: log2(e), 2^52 + 2^51, ln2 hi, ln2 lo */                                      \
    1.4426950408889634, 6755399441055744.0, 6.93147180369123816490e-01,                     \
    1.90821492927058770002e-10,                                                             \
    /* 17..22: AKP10 Gtilde: 1.808, 3.4, 2.210, 0.347, 1.353, 0.217 */                      \
    1.808, 3.4, 2.210, 0.347, 1.353, 0.217,                                                 \
    /* 23..27: atanh series 1/3, 1/5, 1/7, 1/9, 1/11 */                                     \
    1.0 / 3.0, 1.0 / 5.0, 1.0 / 7.0, 1.0 / 9.0, 1.0 / 11.0                                  \
  }
#if defined(__CUDACC__)
__constant__ double NB_C_DEV[28] = NB_CONST_TABLE;
#endif
static const double NB_C_HOST[28] = NB_CONST_TABLE;
#if defined(__CUDA_ARCH__)
#define NB_K(i) NB_C_DEV[i]
#else
#define NB_K(i) NB_C_HOST[i]
#endif

// exp(-x) for x >= 0 (NaN propagates): k = round(-x log2 e), r = -x - k ln2 in two
// pieces, degree-12 Taylor polynomial on |r| <= ln2/2 (max error 3.7e-16 incl.
// rounding), scaling by 2^k in two factors so that the result underflows gradually.
NB_HD int nb_loword(double v) {
#if defined(__CUDA_ARCH__)
  return __double2loint(v);
#else
  long long b;
  memcpy(&b, &v, 8);
  return (int)(b & 0xffffffffLL);
#endif
}

NB_HD double nb_pow2(int k) {  // 2^k, -1022 <= k <= 1023
#if defined(__CUDA_ARCH__)
  return __hiloint2double((k + 1023) << 20, 0);
#else
  long long b = (long long)(k + 1023) << 52;
  double v;
  memcpy(&v, &b, 8);
  return v;
#endif
}

NB_HD double exp_neg(double x) {
  const bool ovf = x < -709.0;  // exp(-x) overflows (only reachable with B < 0)
  x = (x > 1100.0) ? 1100.0 : x;
  x = ovf ? -709.0 : x;
  double t = fma(-x, NB_K(13), NB_K(14));
  int k = nb_loword(t);
  double kd = t - NB_K(14);
  double r = fma(kd, -NB_K(15), -x);
  r = fma(kd, -NB_K(16), r);
  double p = fma(NB_K(12), r, NB_K(11));
  p = fma(p, r, NB_K(10));
  p = fma(p, r, NB_K(9));
  p = fma(p, r, NB_K(8));
  p = fma(p, r, NB_K(7));
  p = fma(p, r, NB_K(6));
  p = fma(p, r, NB_K(5));
  p = fma(p, r, NB_K(4));
  p = fma(p, r, NB_K(3));
  p = fma(p, r, NB_K(2));
  p = fma(p, r, 1.0);
  p = fma(p, r, 1.0);
  int k1 = k >> 1, k2 = k - k1;
  double v = (p * nb_pow2(k1)) * nb_pow2(k2);
  return ovf ? INFINITY : v;
}

// ---------------------------------------------------------------------------
// the same distributions in log space, for the integration operands.  With
// L = ln(e/e_0):  ln n = P - c,  P = ln|A| - alpha L (+ ln K beyond the break),
// c = (e/e_c)^beta = exp(beta ln(e/e_c)) (e/e_c itself for beta == 1).  One or two
// log and one or two exp per node instead of two pow and one exp, and the logarithmic
// slope between neighbouring nodes comes out without cancellation:
// d ln n / d ln x = -alpha - (c2 - c1)/ln(x2/x1).
// Agrees with pd_eval to (|alpha L| + c) ulps, i.e. a few 1e-15 wherever n matters.
// ---------------------------------------------------------------------------
struct PdLog {
  int kind;
  double sgn;    // sign of amplitude * n_scale (0 if zero)
  double lnA;    // ln|amplitude * n_scale|
  double e0;     // e_0
  double a1, a2; // index below / above the break (a1 only: no break)
  double eb;     // e_break (BPL, ECBPL)
  double lnK;    // (a2 - a1) ln(e_b / e_0)
  double ec;     // e_cutoff
  double beta;   // cutoff exponent, or LogParabola curvature
  bool cutoff, broken;
  // grid-relative constants for pd_log_node_tab (e = x * m with m = e_mul1 * e_mul2)
  double m, lnme0, mec, lnmec;  // m, ln(m / e_0), m / e_c, ln(m / e_c)
};

struct PdNode {
  double P, c, L;
  int side;
};

NB_HD PdLog pd_log_setup(int kind, const double* p, double n_scale) {
  PdLog s;
  s.kind = kind;
  double A = p[0] * n_scale;
  s.sgn = (A > 0.0) ? 1.0 : ((A < 0.0) ? -1.0 : A);  // nan stays nan
  s.lnA = log(fabs(A));
  s.e0 = p[1];
  s.cutoff = (kind == PD_ECPL || kind == PD_ECBPL);
  s.broken = (kind == PD_BPL || kind == PD_ECBPL);
  s.a1 = s.a2 = 0.0; s.eb = 0.0; s.lnK = 0.0; s.ec = 1.0; s.beta = 0.0;
  if (kind == PD_PL) {
    s.a1 = s.a2 = p[2];
  } else if (kind == PD_ECPL) {
    s.a1 = s.a2 = p[2];
    s.ec = p[3];
    s.beta = p[4];
  } else if (kind == PD_BPL || kind == PD_ECBPL) {
    s.eb = p[2];
    s.a1 = p[3];
    s.a2 = p[4];
    s.lnK = (p[4] - p[3]) * log(p[2] / p[1]);
    if (kind == PD_ECBPL) {
      s.ec = p[5];
      s.beta = p[6];
    }
  } else {  // PD_LOGPAR
    s.a1 = s.a2 = p[2];
    s.beta = p[3];
  }
  return s;
}

// grid-relative constants: with them a node costs no log (ln x comes from the grid's
// walker-independent table) and, for beta == 1, a single exp
NB_HD void pd_log_setup_grid(PdLog& s, double e_mul1, double e_mul2) {
  s.m = e_mul1 * e_mul2;
  s.lnme0 = log(s.m / s.e0);
  s.mec = s.m / s.ec;
  s.lnmec = log(s.mec);
}

// node x of a grid whose ln(x) is tabulated
NB_HD PdNode pd_log_node_tab(const PdLog& s, double x, double lnx) {
  PdNode nd;
  nd.L = lnx + s.lnme0;
  nd.side = (s.broken && !(x * s.m < s.eb)) ? 1 : 0;
  nd.c = 0.0;
  if (s.cutoff) nd.c = (s.beta == 1.0) ? x * s.mec : exp(s.beta * (lnx + s.lnmec));
  if (s.kind == PD_LOGPAR)
    nd.P = s.lnA - (s.a1 + s.beta * nd.L) * nd.L;
  else if (nd.side)
    nd.P = (s.lnA + s.lnK) - s.a2 * nd.L;
  else
    nd.P = s.lnA - s.a1 * nd.L;
  return nd;
}

// e: node energy [eV]
NB_HD PdNode pd_log_node(const PdLog& s, double e) {
  PdNode nd;
  nd.L = log(e / s.e0);
  nd.side = (s.broken && !(e < s.eb)) ? 1 : 0;
  nd.c = 0.0;
  if (s.cutoff) {
    double r = e / s.ec;
    nd.c = (s.beta == 1.0) ? r : exp(s.beta * log(r));
  }
  if (s.kind == PD_LOGPAR)
    nd.P = s.lnA - (s.a1 + s.beta * nd.L) * nd.L;
  else if (nd.side)
    nd.P = (s.lnA + s.lnK) - s.a2 * nd.L;
  else
    nd.P = s.lnA - s.a1 * nd.L;
  return nd;
}

NB_HD double pd_log_value(const PdLog& s, const PdNode& nd) { return s.sgn * exp(nd.P - nd.c); }

// same through exp_neg (no libm call in the hot loops)
NB_HD double pd_log_value_fast(const PdLog& s, const PdNode& nd) {
  return s.sgn * exp_neg(nd.c - nd.P);
}

// d ln n / d ln x + 1 over the interval (a, b); invdlx = 1/ln(x_b/x_a)
NB_HD double pd_log_ds1(const PdLog& s, const PdNode& a, const PdNode& b, double invdlx) {
  double slope;
  if (s.kind == PD_LOGPAR)
    slope = -(s.a1 + s.beta * (a.L + b.L));
  else if (a.side == b.side)
    slope = a.side ? -s.a2 : -s.a1;
  else
    slope = (b.P - a.P) * invdlx;
  return (slope - (b.c - a.c) * invdlx) + 1.0;
}

// ---------------------------------------------------------------------------
// log-log trapezoid, one interval
// ---------------------------------------------------------------------------
// reference operation order (utils.py:336-348)
NB_HD double interval_exact(double x1, double x2, double y1, double y2) {
  double b = log10(y2 / y1) / log10(x2 / x1);
  double v;
  if (fabs(b + 1.0) > 1e-10)
    v = (y1 * (x2 * pow(x2 / x1, b) - x1)) / (b + 1);
  else
    v = x1 * y1 * log(x2 / x1);
  if (y1 == 0.0 || y2 == 0.0 || x1 == x2) v = 0.0;
  return v;
}

// 1/b to ~1 ulp: hardware seed (MUFU.RCP64H, rel. error <= 2^-23) + one cubically
// convergent Newton step (3 DFMA).  Not correctly rounded -- the IEEE division it
// replaces costs ~25 instructions plus a slow path for zero/NaN operands, which the
// zero-padded emissivity tables hit all the time.  Zero, inf, NaN and denormal b give
// NaN/inf here; every caller masks those cases with its own selects.
NB_HD double fast_rcp(double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  double e = fma(-b, r, 1.0);
  e = fma(e, e, e);
  return fma(r, e, r);
#else
  return 1.0 / b;
#endif
}

// 1/sqrt(d) to ~1 ulp for normal positive d: hardware seed (MUFU.RSQ64H, <= 2^-23) and
// one cubically convergent step, y (1 + e/2 + 3 e^2/8) with e = 1 - d y^2.
NB_HD double fast_rsqrt(double d) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-(d * y), y, 1.0);
  double t = fma(e, 0.375, 0.5) * e;
  return fma(y, t, y);
#else
  return 1.0 / sqrt(d);
#endif
}

// hoisted form: xy = x*y at both nodes, bp1 = b + 1 with
// b = ln(y2/y1)/ln(x2/x1), dlx = ln(x2/x1).  (x2/x1)^b == y2/y1, so
// y1 (x2 (x2/x1)^b - x1) == x2 y2 - x1 y1.
NB_HD double interval_fast(double xy1, double xy2, double bp1, double dlx) {
  double v = (xy2 - xy1) * fast_rcp(bp1);
#if defined(__CUDA_ARCH__)
  // the classification runs on the integer pipe (the fp64 pipe is the bottleneck):
  // regular slope <=> 1e-10 < |bp1| < inf, tested on the high word (cut at 1.0000002e-10;
  // both branches agree to 1e-12 there); NaN/inf slopes take the log branch like the
  // reference's `abs(b + 1) > 1e-10` being False for NaN
  unsigned hb = (unsigned)__double2hiint(bp1) & 0x7fffffffu;
  bool regular = (hb - 0x3DDB7CE0u) < (0x7FF00000u - 0x3DDB7CE0u);
  if (!regular) v = xy1 * dlx;
  unsigned z1 = ((unsigned)__double2hiint(xy1) << 1) | (unsigned)__double2loint(xy1);
  unsigned z2 = ((unsigned)__double2hiint(xy2) << 1) | (unsigned)__double2loint(xy2);
  if (z1 == 0u || z2 == 0u) v = 0.0;
#else
  if (!(fabs(bp1) > 1e-10)) v = xy1 * dlx;  // also the NaN-slope branch
  if (xy1 == 0.0 || xy2 == 0.0) v = 0.0;
#endif
  return v;
}

// Sentinel slope for intervals with a zero end point.  Tables (nb_table_finalize) and the
// self-Compton seed operands store it instead of the +-inf / NaN of ln(0/K).  Its
// reciprocal is subnormal, which the hardware seed (rcp.approx.ftz.f64) flushes to zero:
// (x2 y2 - x1 y1) * 0 -- the interval contributes exactly nothing, as utils.py:347 demands,
// without a zero test in the lean cell.  Table and walker sentinels may add up (1.5 * 2^1023
// is still finite).  The careful cell (interval_fast) zeroes such intervals by its own test.
constexpr double NB_BIG_SLOPE = 6.741349255733685e307;  // 1.5 * 2^1022

NB_HD unsigned nb_hiword(double v) {
#if defined(__CUDA_ARCH__)
  return (unsigned)__double2hiint(v);
#else
  unsigned long long b;
  memcpy(&b, &v, 8);
  return (unsigned)(b >> 32);
#endif
}

// interval_fast classifies a slope as regular when 1e-10 < |b + 1| < inf, tested on the high
// word: (hi & 0x7fffffff) - NB_REG_LO < NB_REG_RANGE (unsigned)
constexpr unsigned NB_REG_LO = 0x3DDB7CE0u;
constexpr unsigned NB_REG_RANGE = 0x7FF00000u - 0x3DDB7CE0u;

// The lean cell: acc += (xy2 - xy1) / bp1 with the reciprocal from the hardware seed and one
// cubic step folded into the accumulation -- 8 fp64 instructions, one MUFU and three integer
// instructions, no select.  It does NOT handle irregular slopes (|b + 1| <= 1e-10, NaN, inf):
// `worst` keeps the running maximum of the classification word, and the caller re-integrates
// with interval_fast when worst >= NB_REG_RANGE at the end (rare: tables with sign changes,
// NaN parameters).  Zero end points must carry the NB_BIG_SLOPE sentinel.
NB_HD void cell_lean(double xy1, double xy2, double bp1, double& acc, unsigned& worst) {
  const unsigned t = (nb_hiword(bp1) & 0x7fffffffu) - NB_REG_LO;
  worst = (t > worst) ? t : worst;
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(bp1));
  const double e = fma(-bp1, r, 1.0);
  const double p = fma(e, e, e);
  const double dr = (xy2 - xy1) * r;
  acc += dr;
  acc = fma(dr, p, acc);
#else
  double r = 1.0 / bp1;
  if (fabs(r) < 2.2250738585072014e-308) r = 0.0;  // the seed's flush-to-zero
  acc += (xy2 - xy1) * r;
#endif
}

// ---------------------------------------------------------------------------
// synchrotron AKP10 Eq. D7
// ---------------------------------------------------------------------------
NB_HD double gtilde(double x) {
  double cb = cbrt(x);
  double cb2 = cb * cb;
  double cb4 = cb2 * cb2;
  double gt1 = 1.808 * cb / sqrt(1 + 3.4 * cb2);
  double gt2 = 1 + 2.210 * cb2 + 0.347 * cb4;
  double gt3 = 1 + 1.353 * cb2 + 0.217 * cb4;
  return gt1 * (gt2 / gt3) * exp(-x);
}

// rational part of gtilde as a function of cb = cbrt(x) (no exp)
NB_HD double gtilde_rational(double cb) {
  double cb2 = cb * cb;
  double cb4 = cb2 * cb2;
  double gt1 = 1.808 * cb / sqrt(1 + 3.4 * cb2);
  double gt2 = 1 + 2.210 * cb2 + 0.347 * cb4;
  double gt3 = 1 + 1.353 * cb2 + 0.217 * cb4;
  return gt1 * (gt2 / gt3);
}

// ---------------------------------------------------------------------------
// Khangulyan+14 IC on Planck seeds
// ---------------------------------------------------------------------------
NB_HD double G12(double x, double al, double a, double be, double b) {
  double G = (PI26 + x) * exp(-x);
  double tmp = 1 + b * pow(x, be);
  double g = 1.0 / (a * pow(x, al) / tmp + 1.0);
  return G * g;
}

NB_HD double G34(double x, double al, double a, double be, double b, double c) {
  double tmp = (1 + c * x) / (1 + PI26 * c * x);
  double G = PI26 * tmp * exp(-x);
  tmp = 1 + b * pow(x, be);
  double g = 1.0 / (a * pow(x, al) / tmp + 1.0);
  return G * g;
}

constexpr double KTOMEC2 = 1.6863699549e-10;
constexpr double IC_PLANCK_NORM = 2.6318735743809104e16;

// gam: electron Lorentz factor, T_K: seed temperature, Eph: photon energy / mec2
NB_HD double ic_iso_planck(double gam, double T_K, double Eph) {
  double Tm = T_K * KTOMEC2;
  double z = Eph / gam;
  double x = z / (1 - z) / (4.0 * gam * Tm);
  double cs = z * z / (2 * (1 - z)) * G34(x, 0.606, 0.443, 1.481, 0.540, 0.319) +
              G34(x, 0.461, 0.726, 1.457, 0.382, 6.620);
  double r = Tm / gam;
  double tmp = r * r;
  tmp *= IC_PLANCK_NORM;
  cs = tmp * cs;
  return ((Eph < gam) && (gam > 1)) ? cs : 0.0;
}

NB_HD double ic_ani_planck(double gam, double T_K, double Eph, double theta) {
  double Tm = T_K * KTOMEC2;
  double z = Eph / gam;
  double ttheta = 2.0 * gam * Tm * (1.0 - cos(theta));
  double x = z / (1 - z) / ttheta;
  double cs = z * z / (2 * (1 - z)) * G12(x, 0.857, 0.153, 1.840, 0.254) +
              G12(x, 0.691, 1.330, 1.668, 0.534);
  double r = Tm / gam;
  double tmp = r * r;
  tmp *= IC_PLANCK_NORM;
  cs = tmp * cs;
  return ((Eph < gam) && (gam > 1)) ? cs : 0.0;
}

NB_HD double heaviside(double x) {
  double s = (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : x);  // np.sign (nan -> nan, 0 -> 0)
  return (s + 1) / 2.0;
}

constexpr double SIGT = 6.652458734983284e-25;

// Aharonian & Atoyan 81 Eq. 22 integrand incl. the two step functions and the
// NaN -> 0 clean-up (radiative.py:626-636); photE0 in mec2 units.
NB_HD double ic_mono_f(double gam, double photE0, double Eph) {
  double b = 4 * photE0 * gam;
  double w = Eph / gam;
  double q = w / (b * (1 - w));
  double bq = b * q;
  double fic = 2 * q * log(q) + (1 + 2 * q) * (1 - q) +
               (1.0 / 2.0) * (bq * bq) * (1 - q) / (1 + b * q);
  double g = fic * heaviside(1 - q) * heaviside(q - 1.0 / (4 * (gam * gam)));
  return (g != g) ? 0.0 : g;
}

// slope term ln(y2/y1)/ln(x2/x1) of a tabulated factor between neighbouring nodes, with the
// finite sentinel for zero end points (see NB_BIG_SLOPE); NaN (sign change) stays NaN and
// selects the careful cell
NB_HD double slope_or_sentinel(double y1, double y2, double invdlx) {
  return (y1 == 0.0 || y2 == 0.0) ? NB_BIG_SLOPE : log(y2 / y1) * invdlx;
}

// ---------------------------------------------------------------------------
// bremsstrahlung, Baring+99; cross sections in cm2 per mec2 of photon energy
// ---------------------------------------------------------------------------
NB_HD double brems_sigma_1(double gam, double eps) {
  double s1 = 4 * (R0_CM * R0_CM) * ALPHA_FS / eps;
  double s2 = 1 + (1.0 / 3.0 - eps / gam) * (1 - eps / gam);
  double s3 = log(2 * gam * (gam - eps) / eps) - 1.0 / 2.0;
  if (gam < eps) s3 = 0.0;
  return s1 * s2 * s3;
}

NB_HD double brems_sigma_2(double gam, double eps) {
  double s0 = (R0_CM * R0_CM) * ALPHA_FS / (3 * eps);
  double e2 = eps * eps;
  double r;
  if (eps <= 0.5) {
    double s1_1 = 16 * (1 - eps + e2) * log(gam / eps);
    double s1_2 = -1 / e2 + 3 / eps - 4 - 4 * eps - 8 * e2;
    double s1_3 = -2 * (1 - 2 * eps) * log(1 - 2 * eps);
    double s1_4 = 1 / (4 * (e2 * eps)) - 1 / (2 * e2) + 3 / eps - 2 + 4 * eps;
    r = s1_1 + s1_2 + s1_3 * s1_4;
  } else {
    double s2_1 = 2 / eps;
    double s2_2 = (4 - 1 / eps + 1 / (4 * e2)) * log(2 * gam);
    double s2_3 = -2 + 2 / eps - 5 / (8 * e2);
    r = s2_1 * (s2_2 + s2_3);
  }
  return s0 * r * heaviside(gam - eps);
}

NB_HD double brems_sigma_ee_rel(double gam, double eps) {
  double A = 1 - 8.0 / 3.0 * pow(gam - 1, 0.2) / (gam + 1) * pow(eps / gam, 1.0 / 3.0);
  return (brems_sigma_1(gam, eps) + brems_sigma_2(gam, eps)) * A;
}

NB_HD double brems_F(double x, double gam) {
  double g2 = gam * gam;
  double beta = sqrt(1 - 1.0 / g2);
  double B = 1 + 0.5 * (g2 - 1);
  double C = 10 * x * gam * beta * (2 + gam * beta);
  C /= 1 + (x * x) * (g2 - 1);
  double tmx = 2 - x;
  double x2 = x * x;
  double F_1 = (17 - 3 * x2 / (tmx * tmx) - C) * sqrt(1 - x);
  double F_2 = 12 * tmx - 7 * x2 / tmx - 3 * (x2 * x2) / (tmx * tmx * tmx);
  double F_3 = log((1 + sqrt(1 - x)) / sqrt(x));
  return B * F_1 + F_2 * F_3;
}

NB_HD double brems_sigma_ee_nonrel(double gam, double eps) {
  double s0 = 4 * (R0_CM * R0_CM) * ALPHA_FS / (15 * eps);
  double g2m1 = gam * gam - 1;
  double x = 4 * eps / g2m1;
  double s = s0 * brems_F(x, gam);
  if (eps >= 0.25 * (gam * gam - 1.0)) s = 0.0;
  if (gam < 1.0) s = 0.0;
  return s;
}

NB_HD double brems_sigma_ee(double gam, double eps) {
  const double gam_trans = 2e6 * EV_ERG / MEC2_ERG;
  return (gam <= gam_trans) ? brems_sigma_ee_nonrel(gam, eps) : brems_sigma_ee_rel(gam, eps);
}

// ---------------------------------------------------------------------------
// pion decay, Kafexhiu+14 analytic parametrisation
// ---------------------------------------------------------------------------
enum PpModel { PP_GEANT4 = 0, PP_PYTHIA8 = 1, PP_SIBYLL = 2, PP_QGSJET = 3 };

NB_HD double pp_sigma_inel(double Tp) {
  double L = log(Tp / T_TH);
  double sigma = 30.7 - 0.96 * L + 0.18 * (L * L);
  double t = 1 - pow(T_TH / Tp, 1.9);
  sigma *= t * t * t;
  return sigma * 1e-27;
}

NB_HD double pp_sigma_pi_loE(double Tp) {
  const double m_p = MPC2_GEV, m_pi = M_PI0;
  const double Mres = 1.1883, Gres = 0.2264;
  double s = 2 * m_p * (Tp + 2 * m_p);
  double gamma = sqrt(Mres * Mres * (Mres * Mres + Gres * Gres));
  double K = sqrt(8.0) * Mres * Gres * gamma;
  K /= NB_PI * sqrt(Mres * Mres + gamma);
  double fBW = m_p * K;
  double t = (sqrt(s) - m_p);
  double u = t * t - Mres * Mres;
  fBW /= u * u + Mres * Mres * Gres * Gres;
  double v = s - m_pi * m_pi - 4 * (m_p * m_p);
  double mu = sqrt(v * v - 16 * (m_pi * m_pi) * (m_p * m_p));
  mu /= 2 * m_pi * sqrt(s);
  const double sigma0 = 7.66e-3;
  double mu2 = mu * mu;
  double mu5 = mu2 * mu2 * mu;
  double sigma1pi = sigma0 * pow(mu, 1.95) * (1 + mu + mu5) * pow(fBW, 1.86);
  double sigma2pi = 5.7 / (1 + exp(-9.3 * (Tp - 1.4)));
  if (Tp < 0.56) sigma2pi = 0.0;
  return (sigma1pi + sigma2pi) * 1e-27;
}

NB_HD double pp_sigma_pi_midE(double Tp) {
  double Qp = (Tp - T_TH) / MPC2_GEV;
  double multip = -6e-3 + 0.237 * Qp - 0.023 * (Qp * Qp);
  return pp_sigma_inel(Tp) * multip;
}

NB_HD double pp_sigma_pi_hiE(double Tp, const double* a) {
  double csip = (Tp - 3.0) / MPC2_GEV;
  double m1 = a[0] * pow(csip, a[3]) * (1 + exp(-a[1] * pow(csip, a[4])));
  double m2 = 1 - exp(-a[2] * pow(csip, 0.25));
  return pp_sigma_inel(Tp) * (m1 * m2);
}

NB_HD void pp_a_table(int model, double* a) {
  const double A[4][5] = {{0.728, 0.596, 0.491, 0.2503, 0.117},
                          {0.652, 0.0016, 0.488, 0.1928, 0.483},
                          {5.436, 0.254, 0.072, 0.075, 0.166},
                          {0.908, 0.0009, 6.089, 0.176, 0.448}};
  for (int i = 0; i < 5; ++i) a[i] = A[model][i];
}

NB_HD double pp_etrans(int model) { return model == PP_PYTHIA8 ? 50.0 : 100.0; }

NB_HD double pp_sigma_pi(double Tp, int model) {
  double a[5];
  if (Tp < 2.0) return pp_sigma_pi_loE(Tp);
  if (Tp < 5.0) return pp_sigma_pi_midE(Tp);
  if (Tp < pp_etrans(model)) {
    pp_a_table(PP_GEANT4, a);
    return pp_sigma_pi_hiE(Tp, a);
  }
  if (Tp >= pp_etrans(model)) {
    pp_a_table(model, a);
    return pp_sigma_pi_hiE(Tp, a);
  }
  return 0.0;  // NaN Tp: np.zeros_like default
}

NB_HD double pp_EpimaxLAB(double Tp) {
  const double m_p = MPC2_GEV, m_pi = M_PI0;
  double s = 2 * m_p * (Tp + 2 * m_p);
  double EpiCM = (s - 4 * (m_p * m_p) + m_pi * m_pi) / (2 * sqrt(s));
  double PpiCM = sqrt(EpiCM * EpiCM - m_pi * m_pi);
  double gCM = (Tp + 2 * m_p) / sqrt(s);
  double betaCM = sqrt(1 - 1.0 / (gCM * gCM));
  return gCM * (EpiCM + PpiCM * betaCM);
}

NB_HD double pp_Egmax(double Tp) {
  const double m_pi = M_PI0;
  double gpiLAB = pp_EpimaxLAB(Tp) / m_pi;
  double betapiLAB = sqrt(1 - 1.0 / (gpiLAB * gpiLAB));
  return (m_pi / 2) * gpiLAB * (1 + betapiLAB);
}

NB_HD double pp_Amax(double Tp, int model) {
  const double m_p = MPC2_GEV;
  if (Tp < 1.0) {
    return 5.9 * pp_sigma_pi(Tp, model) / pp_EpimaxLAB(Tp);
  } else if (Tp >= 1.0) {
    double b1, b2, b3;
    if (Tp < 5.0) {
      b1 = 9.53; b2 = 0.52; b3 = 0.054;
    } else {
      b1 = 9.13; b2 = 0.35; b3 = 9.7e-3;
    }
    if (Tp >= pp_etrans(model)) {
      const double Bm[4][3] = {{9.13, 0.35, 9.7e-3},
                               {9.06, 0.3795, 0.01105},
                               {10.77, 0.412, 0.01264},
                               {13.16, 0.4419, 0.01439}};
      b1 = Bm[model][0]; b2 = Bm[model][1]; b3 = Bm[model][2];
    }
    double thetap = Tp / m_p;
    double lt = log(thetap);
    return b1 * pow(thetap, -b2) * exp(b3 * (lt * lt)) * pp_sigma_pi(Tp, model) / m_p;
  }
  return 0.0;
}

NB_HD double pp_F_func(double Tp, double Egamma, double lamb, double alpha, double beta,
                       double gamma) {
  const double m_pi = M_PI0;
  double Egmax = pp_Egmax(Tp);
  double Yg = Egamma + m_pi * m_pi / (4 * Egamma);
  double Ygmax = Egmax + m_pi * m_pi / (4 * Egmax);
  double Xg = (Yg - m_pi) / (Ygmax - m_pi);
  if (Xg > 1) Xg = 1.0;
  double C = lamb * m_pi / Ygmax;
  double F = pow(1 - pow(Xg, alpha), beta);
  F /= pow(1 + Xg / C, gamma);
  return F;
}

NB_HD double pp_mu(double Tp) {
  double q = (Tp - 1.0) / MPC2_GEV;
  const double x = 5.0 / 4.0;
  return x * pow(q, x) * exp(-x * q);
}

NB_HD double pp_F(double Tp, double Egamma, int model) {
  double F = 0.0;
  if (Tp >= T_TH && Tp <= 1.0) {
    double thetap = Tp / MPC2_GEV;
    double kappa = 3.29 - pow(thetap, -1.5) / 5.0;
    F = pp_F_func(Tp, Egamma, 1.0, 1.0, kappa, 0.0);
  }
  if (Tp > 1.0 && Tp <= 4.0) {
    double mu = pp_mu(Tp);
    F = pp_F_func(Tp, Egamma, 3.0, 1.0, mu + 2.45, mu + 1.45);
  }
  if (Tp > 4.0 && Tp <= 20.0) {
    double mu = pp_mu(Tp);
    F = pp_F_func(Tp, Egamma, 3.0, 1.0, 1.5 * mu + 4.95, mu + 1.50);
  }
  if (Tp > 20.0 && Tp <= 100.0) F = pp_F_func(Tp, Egamma, 3.0, 0.5, 4.2, 1.0);
  if (Tp > pp_etrans(model)) {
    const double Fm[4][4] = {{3.0, 0.5, 4.9, 1.0},
                             {3.5, 0.5, 4.0, 1.0},
                             {3.55, 0.5, 3.6, 1.0},
                             {3.55, 0.5, 4.5, 1.0}};
    F = pp_F_func(Tp, Egamma, Fm[model][0], Fm[model][1], Fm[model][2], Fm[model][3]);
  }
  return F;
}

NB_HD double pp_nuclear_factor(double Tp) {
  const double sigmaRpp = 10 * NB_PI * 1e-27;
  double sigmainel = pp_sigma_inel(Tp);
  double sigmainel0 = pp_sigma_inel(1e3);
  double f = sigmainel / sigmainel0;
  double f2 = (f > 1) ? f : 1.0;
  double G = 1.0 + log(f2);
  double eps = (Tp > T_TH) ? (1.37 + (0.29 + 0.1) * sigmaRpp * G / sigmainel) : 0.0;
  if (Tp > T_TH && Tp < 1.0) eps = 1.9141;
  return eps;
}

// dsigma/dEgamma [cm2/GeV]; Ep, Egamma in GeV
NB_HD double pp_diffsigma(double Ep, double Egamma, int model, int nuclear_enhancement) {
  double Tp = Ep - MPC2_GEV;
  double ds = pp_Amax(Tp, model) * pp_F(Tp, Egamma, model);
  if (nuclear_enhancement) ds *= pp_nuclear_factor(Tp);
  return ds;
}

// ---------------------------------------------------------------------------
// pion decay, Kelner+06 (radiative.py:1592-1646); energies in TeV
// ---------------------------------------------------------------------------
NB_HD double kel_sigma_inel(double Ep) {  // KAB06 Eq. 73, 79 [cm2]
  double L = log(Ep);
  double sigma = 34.3 + 1.88 * L + 0.25 * (L * L);
  if (Ep <= 0.1) {
    const double Eth = 1.22e-3;
    double r = Eth / Ep;
    double r2 = r * r;
    double t = 1 - r2 * r2;
    sigma *= (t * t) * heaviside(Ep - Eth);
  }
  return sigma * 1e-27;
}

NB_HD double kel_Fgamma(double x, double Ep) {  // KAB06 Eq. 58-61
  double L = log(Ep);
  double B = 1.30 + 0.14 * L + 0.011 * (L * L);
  double beta = 1.0 / (1.79 + 0.11 * L + 0.008 * (L * L));
  double k = 1.0 / (0.801 + 0.049 * L + 0.014 * (L * L));
  double xb = pow(x, beta);
  double lx = log(x);
  double q = (1 - xb) / (1 + k * xb * (1 - xb));
  double q2 = q * q;
  double F1 = B * (lx / x) * (q2 * q2);
  double F2 = 1.0 / lx - (4 * beta * xb) / (1 - xb) -
              (4 * k * beta * xb * (1 - 2 * xb)) / (1 + k * xb * (1 - xb));
  return F1 * F2;
}

constexpr double KEL_KPI = 0.17;
constexpr double KEL_MPI_TEV = 1.349766e-4;
constexpr double KEL_MP_TEV = MPC2_GEV * 1e-3;

// integrand kernels per unit proton energy, to be multiplied with J(Ep) [1/TeV]:
//   full calculation (Eq. 71 with x = Eg/Ep, dx/x = -dEp/Ep):  c sigma F(Eg/Ep, Ep) / Ep
//   delta-functional approximation (Eq. 78 with Epi = Kpi (Ep - mp)):
//                                                 2 c sigma / sqrt(Epi^2 - m_pi^2)
NB_HD double kel_kernel_hi(double Ep, double Eg) {
  double x = Eg / Ep;
  if (!(x < 1.0)) return 0.0;  // F -> 0 as x -> 1 (the limit of 0 * inf)
  double v = C_CGS * kel_sigma_inel(Ep) * kel_Fgamma(x, Ep) / Ep;
  return (v == v) ? v : 0.0;
}

NB_HD double kel_kernel_lo(double Ep) {
  double Epi = KEL_KPI * (Ep - KEL_MP_TEV);
  double d = Epi * Epi - KEL_MPI_TEV * KEL_MPI_TEV;
  if (!(d > 0.0)) return 0.0;
  return 2.0 * C_CGS * kel_sigma_inel(Ep) / sqrt(d);
}

// ---------------------------------------------------------------------------
// FITPACK cubic B-spline pieces (fpbspl / fpbisp), 0-based
// ---------------------------------------------------------------------------
// interval search of fpbisp: returns l with t[l] <= x < t[l+1], l in [3, n-5]
NB_HD int bspl_interval(const double* t, int n, double x) {
  int l = 3;
  while (l != n - 5 && x >= t[l + 1]) ++l;
  return l;
}

// the 4 non-zero cubic B-splines at x for knot interval l (fpbspl, k = 3)
NB_HD void bspl_basis(const double* t, double x, int l, double* h) {
  double hh[3];
  h[0] = 1.0;
  for (int j = 1; j <= 3; ++j) {
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
    for (int i = 0; i < j; ++i) {
      int li = l + i + 1;
      int lj = li - j;
      double f = hh[i] / (t[li] - t[lj]);
      h[i] = h[i] + f * (t[li] - x);
      h[i + 1] = f * (x - t[lj]);
    }
  }
}

// ---------------------------------------------------------------------------
// likelihood (core.py:64-94): one data point's Gaussian term
// ---------------------------------------------------------------------------
NB_HD double lnprob_term(double model, double flux, double err_lo, double err_hi) {
  double d = model - flux;
  double s = (d > 0) ? err_hi : err_lo;
  return -(d * d) / (2.0 * (s * s));
}

// priors (core.py:34-58), literal formulas
NB_HD double prior_eval(int kind, double v, double a, double b) {
  if (kind == NB_PRIOR_UNIFORM) return (a <= v && v <= b) ? 0.0 : -INFINITY;
  if (kind == NB_PRIOR_NORMAL) return -0.5 * (2 * NB_PI * b) - ((v - a) * (v - a)) / (2.0 * b);
  // log-uniform: returns 1/value (sic)
  if (v > 0 && v >= a) return (v <= b) ? 1 / v : -INFINITY;
  return -INFINITY;
}

// ---------------------------------------------------------------------------
// per-lane bodies of the hot kernels (lane = contiguous interval range [i0,i1))
// ---------------------------------------------------------------------------
NB_HD int odd_chunk(int nint) {
  int m = (nint + 31) / 32;
  if (m < 1) m = 1;
  if ((m & 1) == 0) ++m;
  return m;
}

// odd chunk length for 64 lanes (a pair of warps per row)
NB_HD int odd_chunk2(int nint) {
  int m = (nint + 63) / 64;
  if (m < 1) m = 1;
  if ((m & 1) == 0) ++m;
  return m;
}

// Walker operands come from global memory (L2).  Two loop shapes: PF == 1 fetches the
// operands of interval i + 1 before the RT cells of interval i (short lane ranges: the
// 13-interval ranges of the 370-node IC grid); PF == 4 runs four intervals behind its loads
// -- a rotating window of prefetched (x*n, slope[, dlx]) triples, indexed statically through
// the unrolled body -- so that their latency hides behind 4 * RT cells even when few warps
// are resident (long ranges: 869-node grids).
template <int RT, int PF>
NB_HD void contract_lane_fast(const double* xnw, const double* dsw, const double* dlx,
                              const double* sK, const double* sL, int pitch, int i0, int i1,
                              double* acc) {
  double prev[RT];
  const double n1 = xnw[i0];
#pragma unroll
  for (int r = 0; r < RT; ++r) prev[r] = n1 * sK[r * pitch + i0];
  if (PF == 1) {
    double n2 = xnw[i0 + 1], d = dsw[i0], dl = dlx[i0];
    for (int i = i0; i < i1; ++i) {
      const int in = (i + 1 < i1) ? i + 1 : i;  // the last prefetch repeats (unused)
      const double n2n = xnw[in + 1], dn = dsw[in], dln = dlx[in];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const double xy2 = n2 * sK[r * pitch + i + 1];
        const double bp1 = d + sL[r * pitch + i];
        acc[r] += interval_fast(prev[r], xy2, bp1, dl);
        prev[r] = xy2;
      }
      n2 = n2n;
      d = dn;
      dl = dln;
    }
    return;
  }
  double n2v[PF], dv[PF], dlv[PF];
#pragma unroll
  for (int k = 0; k < PF; ++k) {
    const int ik = (i0 + k < i1) ? i0 + k : i1 - 1;  // clamped: loads past the range repeat
    n2v[k] = xnw[ik + 1];
    dv[k] = dsw[ik];
    dlv[k] = dlx[ik];
  }
  for (int i = i0; i < i1; i += PF) {
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      if (i + k < i1) {
        const double n2 = n2v[k], d = dv[k], dl = dlv[k];
        const int ip = (i + k + PF < i1) ? i + k + PF : i1 - 1;
        n2v[k] = xnw[ip + 1];
        dv[k] = dsw[ip];
        dlv[k] = dlx[ip];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const double xy2 = n2 * sK[r * pitch + i + k + 1];
          const double bp1 = d + sL[r * pitch + i + k];
          acc[r] += interval_fast(prev[r], xy2, bp1, dl);
          prev[r] = xy2;
        }
      }
    }
  }
}

// lean contraction (cell_lean): same operands as contract_lane_fast; returns the running
// maximum of the slope classification word (>= NB_REG_RANGE: some interval was irregular and
// the caller must redo the range with contract_lane_fast)
template <int RT, int PF>
NB_HD unsigned contract_lane_lean(const double* xnw, const double* dsw, const double* sK,
                                  const double* sL, int pitch, int i0, int i1, double* acc) {
  double prev[RT];
  unsigned worst = 0u;
  const double n1 = xnw[i0];
#pragma unroll
  for (int r = 0; r < RT; ++r) prev[r] = n1 * sK[r * pitch + i0];
  if (PF == 1) {
    double n2 = xnw[i0 + 1], d = dsw[i0];
#pragma unroll 2
    for (int i = i0; i < i1; ++i) {
      const int in = (i + 1 < i1) ? i + 1 : i;  // the last prefetch repeats (unused)
      const double n2n = xnw[in + 1], dn = dsw[in];
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const double xy2 = n2 * sK[r * pitch + i + 1];
        cell_lean(prev[r], xy2, d + sL[r * pitch + i], acc[r], worst);
        prev[r] = xy2;
      }
      n2 = n2n;
      d = dn;
    }
    return worst;
  }
  double n2v[PF], dv[PF];
#pragma unroll
  for (int k = 0; k < PF; ++k) {
    const int ik = (i0 + k < i1) ? i0 + k : i1 - 1;
    n2v[k] = xnw[ik + 1];
    dv[k] = dsw[ik];
  }
  for (int i = i0; i < i1; i += PF) {
#pragma unroll
    for (int k = 0; k < PF; ++k) {
      if (i + k < i1) {
        const double n2 = n2v[k], d = dv[k];
        const int ip = (i + k + PF < i1) ? i + k + PF : i1 - 1;
        n2v[k] = xnw[ip + 1];
        dv[k] = dsw[ip];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const double xy2 = n2 * sK[r * pitch + i + k + 1];
          cell_lean(prev[r], xy2, d + sL[r * pitch + i + k], acc[r], worst);
          prev[r] = xy2;
        }
      }
    }
  }
  return worst;
}

// exact contraction: reference operation order per interval (nw = n itself)
template <int RT>
NB_HD void contract_lane_exact(const double* nw, const double* xg, const double* sK, int pitch,
                               int i0, int i1, double* acc) {
  double prev[RT];
  double n1 = nw[i0];
#pragma unroll
  for (int r = 0; r < RT; ++r) prev[r] = n1 * sK[r * pitch + i0];
  for (int i = i0; i < i1; ++i) {
    double n2 = nw[i + 1];
    double x1 = xg[i], x2 = xg[i + 1];
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      double y2 = n2 * sK[r * pitch + i + 1];
      acc[r] += interval_exact(x1, x2, prev[r], y2);
      prev[r] = y2;
    }
  }
}

// synchrotron per-node setup: 1/Ec and cbrt(1/Ec), Ec = 3 e hbar B g^2 / (2 m_e c)
// (radiative.py:331-332)
NB_HD void syn_node(double g, double B, double* iec, double* cb) {
  double Ec = 3 * E_ESU * HBAR_CGS * B * (g * g);
  Ec /= 2 * (M_E_G * C_CGS);
  double i = 1.0 / Ec;
  *iec = i;
  *cb = cbrt(i);
}

// the same split into a per-walker and a per-node factor: Ec = kB g^2, so
// 1/Ec = (1/kB) gm2[j] and cbrt(1/Ec) = cbrt(1/kB) g23[j] with the grid's walker-independent
// tables gm2 = g^-2, g23 = cbrt(g^-2) -- two multiplications per node instead of a division
// and a cbrt (agrees with syn_node to 2 ulp)
NB_HD void syn_walker(double B, double* ikB, double* cbk) {
  double kB = 3 * E_ESU * HBAR_CGS * B;
  kB /= 2 * (M_E_G * C_CGS);
  *ikB = 1.0 / kB;
  *cbk = cbrt(*ikB);
}

// rational part of Gtilde with one reciprocal square root:
//   R = 1.808 cb gt2 / (gt3 sqrt(1 + 3.4 cb^2)) = 1.808 cb gt2 rsqrt(gt3^2 (1 + 3.4 cb^2))
NB_HD double gtilde_rational_fast(double cb) {
  double cb2 = cb * cb;
  double cb4 = cb2 * cb2;
  double gt2 = fma(NB_K(20), cb4, fma(NB_K(19), cb2, 1.0));
  double gt3 = fma(NB_K(22), cb4, fma(NB_K(21), cb2, 1.0));
  double d = (gt3 * gt3) * fma(NB_K(18), cb2, 1.0);
  return (NB_K(17) * cb) * gt2 * fast_rsqrt(d);
}

// ln(R2/R1) for neighbouring nodes: 2 atanh((R2-R1)/(R2+R1)) by its series while
// |s| <= 0.05 (truncation < 2e-17 relative; default grids have |s| < 0.008), log otherwise
NB_HD double log_ratio(double R1, double R2) {
  double s = (R2 - R1) * fast_rcp(R2 + R1);
  double s2 = s * s;
  double p = fma(s2, NB_K(27), NB_K(26));
  p = fma(p, s2, NB_K(25));
  p = fma(p, s2, NB_K(24));
  p = fma(p, s2, NB_K(23));
  p = fma(p, s2, 1.0);
  double v = (2.0 * s) * p;
  if (!(fabs(s) <= 0.05)) v = log(R2 / R1);
  return v;
}

// exp(-x) contributes exactly 0 once x > 745.14; nodes are ordered by decreasing x
// (x = E / (kB g^2)), so every interval left of the first node with x <= 746 is zero.
// Returns that node's index (0 when B is not a positive finite number).
NB_HD int syn_first_node(const double* gam, int N, double B, double E_erg) {
  double kB = 3 * E_ESU * HBAR_CGS * B;
  kB /= 2 * (M_E_G * C_CGS);
  double g2 = E_erg / (746.0 * kB);  // x <= 746  <=>  g^2 >= g2
  if (!(g2 > 0.0) || !(g2 < 1e300)) return 0;
  // first j with gam[j]^2 >= g2: guess from a log-uniform grid, then walk (0-2 steps on
  // the reference's np.logspace grids; correct, if slow, on any increasing grid)
  double l0 = log(gam[0]), l1 = log(gam[N - 1]);
  double t = (0.5 * log(g2) - l0) / (l1 - l0) * (N - 1);
  int j = (t > 0.0) ? ((t < (double)N) ? (int)t : N) : 0;
  while (j > 0 && gam[j - 1] * gam[j - 1] >= g2) --j;
  while (j < N && gam[j] * gam[j] < g2) ++j;
  return j;
}

// synchrotron lane: integral of x*n*Gtilde(E/Ec) over intervals [i0,i1) with
// ln(y2/y1) = ln(n2/n1) + ln(R2/R1) - (x2 - x1).  Two intervals per iteration: the two new
// nodes' rational / exponential chains are independent of each other and of the carried node,
// so they overlap in the pipeline (the kernel is latency bound); the accumulation order is
// that of the one-interval loop.
NB_HD double syn_lane(double E, double cbE, const double* s_iec, const double* s_cb,
                      const double* s_xn, const double* s_ds, const double* s_idl,
                      const double* s_dl, int i0, int i1) {
  double acc = 0.0;
  double x1 = E * s_iec[i0];
  double R1 = gtilde_rational_fast(cbE * s_cb[i0]);
  double xy1 = s_xn[i0] * (R1 * exp_neg(x1));
  int i = i0;
  for (; i + 1 < i1; i += 2) {
    const double x2 = E * s_iec[i + 1], x3 = E * s_iec[i + 2];
    const double R2 = gtilde_rational_fast(cbE * s_cb[i + 1]);
    const double R3 = gtilde_rational_fast(cbE * s_cb[i + 2]);
    const double xy2 = s_xn[i + 1] * (R2 * exp_neg(x2));
    const double xy3 = s_xn[i + 2] * (R3 * exp_neg(x3));
    const double bpa = s_ds[i] + (log_ratio(R1, R2) - (x2 - x1)) * s_idl[i];
    const double bpb = s_ds[i + 1] + (log_ratio(R2, R3) - (x3 - x2)) * s_idl[i + 1];
    acc += interval_fast(xy1, xy2, bpa, s_dl[i]);
    acc += interval_fast(xy2, xy3, bpb, s_dl[i + 1]);
    x1 = x3;
    R1 = R3;
    xy1 = xy3;
  }
  if (i < i1) {
    const double x2 = E * s_iec[i + 1];
    const double R2 = gtilde_rational_fast(cbE * s_cb[i + 1]);
    const double xy2 = s_xn[i + 1] * (R2 * exp_neg(x2));
    const double bp1 = s_ds[i] + (log_ratio(R1, R2) - (x2 - x1)) * s_idl[i];
    acc += interval_fast(xy1, xy2, bp1, s_dl[i]);
  }
  return acc;
}

// CS1 * integral * (erg -> eV):  CS1 = sqrt(3) e^3 B / (2 pi m_e c^2 hbar E)
// (radiative.py:319-328, 340)
NB_HD double syn_finish(double B, double E_erg, double integral) {
  double CS1_0 = sqrt(3.0) * (E_ESU * E_ESU * E_ESU) * B;
  double CS1_1 = 2 * NB_PI * M_E_G * (C_CGS * C_CGS) * HBAR_CGS * E_erg;
  return (CS1_0 / CS1_1) * integral * EV_ERG;
}

// ---------------------------------------------------------------------------
// combine + likelihood for one walker (flux radiative.py:102-111; core.py:64-121)
// ---------------------------------------------------------------------------
struct CombineArgs {
  nb_term terms[NB_MAX_TERMS];
  int n_terms;
  int W, N_E;
  const double* unit_fac;
  const double* data_flux;
  const double* err_lo;
  const double* err_hi;
  const int* ul;
  const double* cl;
  const double* prior;
  double* flux_model;
  int flux_ld;  // row pitch of flux_model (>= N_E)
  double* lnp;
  int lnp_ld;  // stride between consecutive walkers' lnp (1 = dense)
};

NB_HD double combine_model(const CombineArgs& a, int w, int e) {
  // all loads first (independent: they overlap instead of queueing behind the
  // data-dependent group logic), then the sums in the order given
  double v[NB_MAX_TERMS], ws[NB_MAX_TERMS];
#pragma unroll
  for (int t = 0; t < NB_MAX_TERMS; ++t) {
    v[t] = 0.0;
    ws[t] = 1.0;
    if (t < a.n_terms) {
      const nb_term& T = a.terms[t];
      v[t] = T.src[(size_t)w * T.ld + T.off + e];
      if (T.wscale) ws[t] = T.wscale[w];
    }
  }
  const double uf = a.unit_fac[e];
  double total = 0.0, g = 0.0;
  bool first_in_group = true, first_group = true;
#pragma unroll
  for (int t = 0; t < NB_MAX_TERMS; ++t) {
    if (t < a.n_terms) {
      const nb_term& T = a.terms[t];
      double x = T.wscale ? v[t] * ws[t] : v[t];
      g = first_in_group ? x : g + x;
      first_in_group = false;
      if (T.group_end) {
        if (T.div != 1.0) g = g / T.div;
        total = first_group ? g : total + g;
        first_group = false;
        first_in_group = true;
      }
    }
  }
  return total * uf;
}

// Sum of the n non-upper-limit terms get(e), e ascending, in numpy's pairwise
// summation order for n < 8 and 8 <= n <= 128 (np.sum of core.py:87); blocks of 8
// accumulators beyond that.
template <class GetT>
NB_HD double numpy_order_sum(int N_E, const int* ul, int n, GetT get) {
  double r[8];
  double seq = 0.0;
  int k = 0;  // index among the non-UL points
  const int nblk = n - (n % 8);
  for (int e = 0; e < N_E; ++e) {
    if (ul[e]) continue;
    double t = get(e);
    if (n < 8) {
      seq = (k == 0) ? t : seq + t;
    } else if (k < 8) {
      r[k] = t;
    } else if (k < nblk) {
      r[k & 7] += t;
    } else {
      if (k == nblk) seq = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
      seq += t;
    }
    ++k;
  }
  if (n >= 8 && n == nblk) seq = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
  if (n == 0) seq = 0.0;
  return seq;
}

// lnprob from the Gaussian sum, the upper-limit violations and the prior (core.py:89-121)
NB_HD double lnprob_finish(const CombineArgs& a, int w, double gauss_sum, int nul, int nviol) {
  double total = gauss_sum;
  if (nul > 0) {
    double clv = (nviol < a.N_E) ? a.cl[nviol] : NAN;
    total += nviol * log(1.0 - clv);
  }
  double pr = a.prior ? a.prior[w] : 0.0;
  return isinf(pr) ? pr : total + pr;
}

// serial form (one walker): host emulation and the reference for the warp kernel
NB_HD void combine_lnprob_walker(const CombineArgs& a, int w) {
  if (!a.lnp) {
    for (int e = 0; e < a.N_E; ++e) a.flux_model[(size_t)w * a.flux_ld + e] = combine_model(a, w, e);
    return;
  }
  int n = 0, nviol = 0, nul = 0;
  for (int e = 0; e < a.N_E; ++e) {
    double m = combine_model(a, w, e);
    if (a.flux_model) a.flux_model[(size_t)w * a.flux_ld + e] = m;
    if (a.ul[e]) {
      ++nul;
      if (m > a.data_flux[e]) ++nviol;
    } else {
      ++n;
    }
  }
  double seq = numpy_order_sum(a.N_E, a.ul, n, [&](int e) {
    return lnprob_term(combine_model(a, w, e), a.data_flux[e], a.err_lo[e], a.err_hi[e]);
  });
  a.lnp[(size_t)w * a.lnp_ld] = lnprob_finish(a, w, seq, nul, nviol);
}

// FITPACK bispev at one point with clamping to the knot range
NB_HD double bspl_eval2d(const double* tx, int nx, const double* ty, int ny, const double* c,
                         double xx, double yy) {
  double tb = tx[3], te = tx[nx - 4];
  if (xx < tb) xx = tb;
  if (xx > te) xx = te;
  tb = ty[3];
  te = ty[ny - 4];
  if (yy < tb) yy = tb;
  if (yy > te) yy = te;
  int lx = bspl_interval(tx, nx, xx);
  int ly = bspl_interval(ty, ny, yy);
  double hx[4], hy[4];
  bspl_basis(tx, xx, lx, hx);
  bspl_basis(ty, yy, ly, hy);
  int nky1 = ny - 4;
  double sp = 0.0;
  for (int i1 = 0; i1 < 4; ++i1) {
    const double* crow = c + (size_t)(lx - 3 + i1) * nky1 + (ly - 3);
    for (int j1 = 0; j1 < 4; ++j1) sp = sp + crow[j1] * hx[i1] * hy[j1];
  }
  return sp;
}

}  // namespace nb
