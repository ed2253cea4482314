// nb_kernels.cu -- sm_100a kernels and C ABI of the naima likelihood hot path.
//
// Design: every radiative process is  spec[w,e] = sum_c coef *
// trapz_loglog(n[w,:] * K_c[e,:], x).  The emissivity tables K are walker
// independent, so they are built once per (grid, photon energies, seeds) by the
// *_table kernels in the reference's operation order.  One likelihood evaluation of W
// walkers is then
//   walker_prep_kernel        parameter map + priors (+ stretch-move proposals), the
//                             particle distribution on the grids in log space
//   contract_kernel           log-log trapezoid contraction; table tile staged in shared
//                             memory by TMA bulk copies, warp-shuffle reduction along
//                             the integration axis                (|| on a forked branch:)
//   synchrotron_fused_kernel  self-contained: parameters -> operands -> AKP10 integral
//   combine_lnprob_kernel     component sums, units, Gaussian/upper-limit likelihood
//                             (+ emcee accept step, chain append, and -- walker sharding
//                             -- the replicated state written to every rank over NVLink)
// captured with the device-resident ensemble state in one CUDA graph per ensemble step.
// All arithmetic is IEEE fp64; no tensor cores (there is no dense contraction).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <cuda/ptx>

#include "../../include/naima_b200.h"
#include "nb_math.cuh"

namespace ptx = cuda::ptx;
using namespace nb;

#define NB_CHECK_LAUNCH()                         \
  do {                                            \
    cudaError_t e__ = cudaGetLastError();         \
    if (e__ != cudaSuccess) return (int)e__;      \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return (cudaStream_t)s; }

// Launch with the caller's preferred shared-memory carve-out (nb_launch_carveout): an SM changes
// its L1 / shared-memory split only when it is idle, so kernels of one evaluation that run on
// parallel graph branches can only share an SM if they ask for the same split.  Measured on the
// C3 half-step: without it the contraction's first CTA starts 21 us after the set-up kernel
// ended (when synchrotron CTAs retire), with a common 50 % split 1.5 us after it.
static thread_local int g_launch_carveout = -1;
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t st, Args&&... args) {
  if (g_launch_carveout < 0) {
    kernel<<<grid, block, smem, st>>>(args...);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributePreferredSharedMemoryCarveout;
  at[0].val.sharedMemCarveout = (unsigned)g_launch_carveout;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#define NB_LAUNCH(...)                                   \
  do {                                                   \
    cudaError_t e__ = launch_k(__VA_ARGS__);             \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

// how often the lean cell had to be redone with the careful cell: [0] contraction (walker,
// row tile) pairs, [1] self-Compton rows (x walkers per thread).  Read with nb_fallback_counts.
__device__ unsigned long long g_fallbacks[2];
// diagnostic timelines (nb_stretch.timeline): the row of the half-step in flight, published by
// the set-up kernel for the kernels that do not see the stretch descriptor (the contraction)
__device__ unsigned long long* g_timeline_row;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  return v;
}

// RT sums at once: on return lane (32 / RT) * r holds the sum over the 32 lanes of acc[r], with
// the pairing tree of warp_sum (offsets 16, 8, 4, 2, 1), hence bit-identical to it.  While more
// than one row is left, the two halves of every lane pair split the rows between them (each
// keeps half and hands the other half over), so the tree costs RT - 1 + log2(32 / RT) shuffles
// instead of 5 RT: 9 instead of 40 for eight rows.
template <int RT>
__device__ __forceinline__ double warp_sum_rows(double (&acc)[RT], const int lane) {
  int off = 16;
#pragma unroll
  for (int n = RT; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int r = 0; r < n / 2; ++r) {
      const double send = upper ? acc[r] : acc[r + n / 2];
      const double keep = upper ? acc[r + n / 2] : acc[r];
      acc[r] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  double v = acc[0];
  for (; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Spin loops on flags that another GPU (or a bulk copy) raises carry a watchdog: after
// NB_WATCHDOG_NS of waiting the kernel traps -- a dead peer then surfaces as a CUDA error on
// this rank instead of hanging every GPU of the job.
constexpr unsigned long long NB_WATCHDOG_NS = 20ull * 1000ull * 1000ull * 1000ull;  // 20 s

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Replicated ensemble state (walker sharding with peer pushes): spin until every rank's
// flag has reached this rank's generation count, i.e. all pushes of the previous
// half-step have landed in our copy.  Called by whole CTAs before they read coords.
__device__ __forceinline__ unsigned long long* timeline_row(const nb_stretch& mv) {
  return mv.timeline +
         NB_TIMELINE_COLS * ((size_t)(2 * *mv.step + mv.split) & (size_t)(NB_TIMELINE_CAP - 1));
}

__device__ __forceinline__ void wait_for_peers(const nb_stretch& mv, bool first_kernel = false) {
  const bool stamp = first_kernel && mv.timeline && threadIdx.x == 0 && blockIdx.x == 0 &&
                     blockIdx.y == 0;
  if (stamp) timeline_row(mv)[0] = global_timer_ns();
  if (mv.wait_flags != nullptr) {
    if ((int)threadIdx.x < mv.wait_world) {
      const unsigned long long need = *mv.wait_gen;
      unsigned long long v, t0 = 0;
      unsigned spins = 0;
      for (;;) {
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];"
                     : "=l"(v)
                     : "l"(mv.wait_flags + threadIdx.x)
                     : "memory");
        if (v >= need) break;
        if ((++spins & 1023u) == 0u) {  // look at the clock every 1024 polls
          const unsigned long long now = global_timer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > NB_WATCHDOG_NS) __trap();  // a peer never arrived
        }
      }
    }
    __syncthreads();
  }
  if (stamp) timeline_row(mv)[1] = global_timer_ns();
}

// ---------------------------------------------------------------------------
// particle distribution kernels
// ---------------------------------------------------------------------------
__global__ void pdist_eval_kernel(int kind, const double* __restrict__ params, int W,
                                  const double* __restrict__ e, int N, double* __restrict__ out,
                                  int out_ld) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int w = blockIdx.y;
  if (i >= N || w >= W) return;
  double p[PD_MAXPAR];
#pragma unroll
  for (int k = 0; k < PD_MAXPAR; ++k) p[k] = params[w * PD_MAXPAR + k];
  out[(size_t)w * out_ld + i] = pd_eval(kind, p, e[i]);
}

// One chunk of PREP_CHUNK = 512 nodes of one walker (two per thread: the transcendental
// chains of the two nodes interleave): n on the grid, x*n and the logarithmic slope term
// ds1 per interval.  nraw != NULL selects the reference-order evaluation (pd_eval with
// pow, slope from log(n2/n1)) that the exact contraction consumes; otherwise the
// log-space form (pd_log_*).  s_n / s_nd: PREP_CHUNK + 1 entries of scratch shared memory.
constexpr int PREP_CHUNK = 512;

__device__ __forceinline__ void pd_prep_chunk(int kind, const double* pp, const PdLog& S,
                                              const double* x, int N,
                                              double e_mul1, double e_mul2, double n_scale,
                                              const double* invdlx, double* xn, double* ds1,
                                              double* nraw, size_t row, int j0, double* s_n,
                                              PdNode* s_nd) {
  const int tid = threadIdx.x;
  const bool next = (tid == 255 && j0 + PREP_CHUNK < N);  // first node of the next chunk
  double xj[2] = {0.0, 0.0}, nj[2] = {0.0, 0.0};
  PdNode nd[2];
  if (nraw) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int l = tid + 256 * u, j = j0 + l;
      if (j < N) {
        xj[u] = x[j];
        nj[u] = pd_eval(kind, pp, (xj[u] * e_mul1) * e_mul2) * n_scale;
        s_n[l] = nj[u];
      }
    }
    if (next)
      s_n[PREP_CHUNK] = pd_eval(kind, pp, (x[j0 + PREP_CHUNK] * e_mul1) * e_mul2) * n_scale;
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int l = tid + 256 * u, j = j0 + l;
      if (j < N) {
        xn[row + j] = xj[u] * nj[u];
        nraw[row + j] = nj[u];
        ds1[row + j] = (j < N - 1) ? log(s_n[l + 1] / nj[u]) * invdlx[j] + 1.0 : 0.0;
      }
    }
  } else {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int l = tid + 256 * u, j = j0 + l;
      nd[u].P = nd[u].c = nd[u].L = 0.0;
      nd[u].side = 0;
      if (j < N) {
        xj[u] = x[j];
        nd[u] = pd_log_node(S, (xj[u] * e_mul1) * e_mul2);
        nj[u] = pd_log_value(S, nd[u]);
        s_nd[l] = nd[u];
      }
    }
    if (next) s_nd[PREP_CHUNK] = pd_log_node(S, (x[j0 + PREP_CHUNK] * e_mul1) * e_mul2);
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int l = tid + 256 * u, j = j0 + l;
      if (j < N) {
        xn[row + j] = xj[u] * nj[u];
        ds1[row + j] = (j < N - 1) ? pd_log_ds1(S, nd[u], s_nd[l + 1], invdlx[j]) : 0.0;
      }
    }
  }
}

__global__ void __launch_bounds__(256) pd_prep_kernel(
    int kind, const double* __restrict__ params, int W, const double* __restrict__ x, int N,
    double e_mul1, double e_mul2, double n_scale, const double* __restrict__ invdlx,
    double* __restrict__ xn, double* __restrict__ ds1, double* __restrict__ nraw, int wpitch) {
  __shared__ double s_n[PREP_CHUNK + 1];
  __shared__ PdNode s_nd[PREP_CHUNK + 1];
  __shared__ PdLog s_S;
  int w = blockIdx.y;
  double p[PD_MAXPAR];
#pragma unroll
  for (int k = 0; k < PD_MAXPAR; ++k) p[k] = params[w * PD_MAXPAR + k];
  if (threadIdx.x == 0) s_S = pd_log_setup(kind, p, n_scale);
  __syncthreads();
  pd_prep_chunk(kind, p, s_S, x, N, e_mul1, e_mul2, n_scale, invdlx, xn, ds1, nraw,
                (size_t)w * wpitch, blockIdx.x * PREP_CHUNK, s_n, s_nd);
}

// W = trapz_loglog(x*n, x*x_to_energy) in reference operation order; one CTA
// (128 threads) per walker, fixed-order tree reduction.
__global__ void particle_energy_kernel(int kind, const double* __restrict__ params,
                                       const double* __restrict__ x, int N, double e_mul1,
                                       double e_mul2, double n_scale, double x_to_energy,
                                       double* __restrict__ out) {
  __shared__ double s_part[128];
  int w = blockIdx.x;
  double p[PD_MAXPAR];
#pragma unroll
  for (int k = 0; k < PD_MAXPAR; ++k) p[k] = params[w * PD_MAXPAR + k];
  double acc = 0.0;
  for (int i = threadIdx.x; i < N - 1; i += 128) {
    double x1 = x[i], x2 = x[i + 1];
    double n1 = pd_eval(kind, p, (x1 * e_mul1) * e_mul2) * n_scale;
    double n2 = pd_eval(kind, p, (x2 * e_mul1) * e_mul2) * n_scale;
    acc += interval_exact(x1 * x_to_energy, x2 * x_to_energy, x1 * n1, x2 * n2);
  }
  s_part[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 64; s > 0; s >>= 1) {
    if (threadIdx.x < s) s_part[threadIdx.x] += s_part[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[w] = s_part[0];
}

// ---------------------------------------------------------------------------
// emissivity table builders (walker independent, reference operation order)
// ---------------------------------------------------------------------------
__global__ void ic_planck_table_kernel(const double* __restrict__ gam, int N,
                                       const double* __restrict__ Eph, int N_E,
                                       const double* __restrict__ seed_T,
                                       const double* __restrict__ seed_theta, int S,
                                       double* __restrict__ K, int pitch, int row0) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;  // s*N_E + e
  if (j >= pitch) return;
  int s = r / N_E, e = r - s * N_E;
  double v = 0.0;
  if (j < N) {
    double th = seed_theta[s];
    if (th != th) v = ic_iso_planck(gam[j], seed_T[s], Eph[e]);
    else v = ic_ani_planck(gam[j], seed_T[s], Eph[e], th);
  }
  K[(size_t)(row0 + r) * pitch + j] = v;
}

// phn_wstride == 0: shared seed density; else per-walker density / table
__global__ void ic_seed_table_kernel(const double* __restrict__ gam, int N,
                                     const double* __restrict__ Eph, int N_E,
                                     const double* __restrict__ eps0,
                                     const double* __restrict__ phn, int Ns, int phn_wstride,
                                     double* __restrict__ K, int pitch, int row0,
                                     long long K_wstride) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int e = blockIdx.y;
  int w = blockIdx.z;
  if (j >= pitch) return;
  const double* ph = phn + (size_t)w * phn_wstride;
  double v = 0.0;
  if (j < N) {
    double g = gam[j], ep = Eph[e];
    if (Ns == 1) {
      double e0 = eps0[0];
      v = ic_mono_f(g, e0, ep);
      v *= ph[0] / (e0 * e0);
    } else {
      double x1 = eps0[0];
      double y1 = ic_mono_f(g, x1, ep) * ph[0] / x1;
      double acc = 0.0;
      for (int s = 1; s < Ns; ++s) {
        double x2 = eps0[s];
        double y2 = ic_mono_f(g, x2, ep) * ph[s] / x2;
        acc += interval_exact(x1, x2, y1, y2);
        x1 = x2;
        y1 = y2;
      }
      v = acc;
    }
    v *= (3.0 / 4.0) * SIGT * 29979245800.0 / (g * g);
  }
  K[(size_t)w * K_wstride + (size_t)(row0 + e) * pitch + j] = v;
}

__global__ void brems_table_kernel(const double* __restrict__ gam, int N,
                                   const double* __restrict__ eps, int N_E,
                                   double* __restrict__ K, int pitch, int row0) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int e = blockIdx.y;
  if (j >= pitch) return;
  double vee = 0.0, vep = 0.0;
  if (j < N) {
    vee = brems_sigma_ee(gam[j], eps[e]) / MEC2_EV;
    vep = brems_sigma_1(gam[j], eps[e]);
  }
  K[(size_t)(row0 + e) * pitch + j] = vee;
  K[(size_t)(row0 + N_E + e) * pitch + j] = vep;
}

__global__ void pp_analytic_table_kernel(int model, int nuc, const double* __restrict__ Ep,
                                         int N, const double* __restrict__ Eg, int N_E,
                                         double* __restrict__ K, int pitch, int row0) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int e = blockIdx.y;
  if (j >= pitch) return;
  double v = 0.0;
  if (j < N) v = pp_diffsigma(Ep[j], Eg[e], model, nuc);
  K[(size_t)(row0 + e) * pitch + j] = v;
}

// FITPACK bispev (fpbisp) with clamping to the knot range
__global__ void pp_lut_table_kernel(const double* __restrict__ tx, int nx,
                                    const double* __restrict__ ty, int ny,
                                    const double* __restrict__ c, const double* __restrict__ Ep,
                                    int N, const double* __restrict__ Eg, int N_E,
                                    double* __restrict__ K, int pitch, int row0) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int e = blockIdx.y;
  if (j >= pitch) return;
  double v = 0.0;
  if (j < N) v = bspl_eval2d(tx, nx, ty, ny, c, log10(Ep[j]), log10(Eg[e]));
  K[(size_t)(row0 + e) * pitch + j] = v;
}

__global__ void table_finalize_kernel(const double* __restrict__ K, int N, int pitch,
                                      const double* __restrict__ invdlx,
                                      double* __restrict__ lrs) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  size_t r = blockIdx.y;
  if (j >= pitch) return;
  double v = 0.0;
  if (j < N - 1) {
    const double k1 = K[r * pitch + j], k2 = K[r * pitch + j + 1];
    // zero end point: the interval is zero whatever the slope (utils.py:347); store the
    // finite sentinel instead of ln(0) so that the lean cell needs no zero test
    v = (k1 == 0.0 || k2 == 0.0) ? NB_BIG_SLOPE : log(k2 / k1) * invdlx[j];
  }
  lrs[r * pitch + j] = v;
}

// per row: index of the first non-zero node (N when the row is all zero) -- the contraction
// skips the leading zeros of a row tile (the kinematic limit of IC / bremsstrahlung, the
// pion-production threshold); flags |= 1 when any entry is negative or not finite (such
// tables keep the careful cell: their NaN slopes select trapz_loglog's log branch)
__global__ void table_scan_kernel(const double* __restrict__ K, int R, int N, int pitch,
                                  int* __restrict__ row_j0, int* __restrict__ flags) {
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= R) return;
  int first = N;
  bool bad = false;
  for (int j = lane; j < N; j += 32) {
    const double k = K[(size_t)r * pitch + j];
    if (k != 0.0 && j < first) first = j;
    bad = bad || !(k >= 0.0) || !(k < INFINITY);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
  const bool anybad = __any_sync(0xffffffffu, bad);
  if (lane == 0) {
    row_j0[r] = first;
    if (anybad) atomicOr(flags, 1);
  }
}

// ---------------------------------------------------------------------------
// the hot contraction
// ---------------------------------------------------------------------------
// CTA = 8 warps.  blockIdx.x selects a tile of RT table rows, staged in shared
// memory with TMA bulk copies (K and lrs, one copy per row from the tile's first live
// column on); blockIdx.y selects a group of walkers, one walker per warp at a time.  Lane l
// integrates the contiguous interval range [jt + l*m, jt + (l+1)*m) (m odd => conflict-free
// 64-bit smem reads), carrying x*y of the previous node in a register; the 32 partial sums
// are combined with a shuffle tree.
// MODE 0: careful cell (interval_fast), 1: reference operation order (interval_exact),
// 2: lean cell (cell_lean) with a per-(walker, tile) fall-back to the careful cell.
struct ContractArgs {
  const double* K;
  const double* lrs;
  int R, N, pitch;
  const double* xn;   // fast: x*n ; exact: n
  const double* ds1;
  int wpitch, W;
  const double* dlx;
  const double* xgrid;
  const double* coef;
  double* out;
  const int* row_j0;  // first non-zero node per row, or NULL
  int w_per_cta;
};


template <int RT, int MODE>
__global__ void __launch_bounds__(256, MODE == 1 ? 1 : 3) contract_kernel(ContractArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned long long* const tl_row = threadIdx.x == 0 ? g_timeline_row : nullptr;  // diagnostic
  if (tl_row && blockIdx.x == 0 && blockIdx.y == 0) tl_row[10] = global_timer_ns();
  constexpr bool EXACT = MODE == 1;
  double* sK = reinterpret_cast<double*>(smem_raw);
  double* sL = sK + (size_t)RT * a.pitch;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sL + (EXACT ? 0 : (size_t)RT * a.pitch));

  const int row0 = blockIdx.x * RT;
  const int nrows = min(RT, a.R - row0);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nint = a.N - 1;

  // first live column of the tile (even: 16-byte aligned bulk copies)
  int jt = 0;
  if (a.row_j0) {
    jt = a.N;
    for (int r = 0; r < nrows; ++r) jt = min(jt, a.row_j0[row0 + r]);
    jt = max(jt - 1, 0) & ~1;  // the interval left of the first non-zero node is zero anyway
  }
  const int wbeg = blockIdx.y * a.w_per_cta;
  const int wend = min(wbeg + a.w_per_cta, a.W);
  if (jt >= nint) {  // all-zero tile
    for (int k = threadIdx.x; k < (wend - wbeg) * nrows; k += blockDim.x)
      a.out[(size_t)(wbeg + k / nrows) * a.R + row0 + k % nrows] = 0.0;
    return;
  }

  if (threadIdx.x == 0) {
    ptx::mbarrier_init(bar, 1);
    ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(a.pitch - jt) * 8u;
    ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, bar,
                                   (EXACT ? 1u : 2u) * bytes * (uint32_t)nrows);
    for (int r = 0; r < nrows; ++r) {
      const size_t off = (size_t)r * a.pitch + jt;
      ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, sK + off,
                         a.K + (size_t)row0 * a.pitch + off, bytes, bar);
      if (!EXACT)
        ptx::cp_async_bulk(ptx::space_cluster, ptx::space_global, sL + off,
                           a.lrs + (size_t)row0 * a.pitch + off, bytes, bar);
    }
  }
  if (nrows < RT) {  // rows beyond the table: zeros (their results are not stored)
    for (int k = threadIdx.x; k < (RT - nrows) * a.pitch; k += blockDim.x) {
      sK[(size_t)nrows * a.pitch + k] = 0.0;
      if (!EXACT) sL[(size_t)nrows * a.pitch + k] = NB_BIG_SLOPE;
    }
    __syncthreads();
  }
  {  // a bulk copy that never completes (bad pointer, lost transaction): fail loudly
    unsigned long long t0 = 0;
    for (unsigned spins = 0; !ptx::mbarrier_try_wait_parity(bar, 0);) {
      if ((++spins & 1023u) == 0u) {
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        else if (now - t0 > NB_WATCHDOG_NS) __trap();
      }
    }
  }

  const int m = odd_chunk(nint - jt);
  const bool deep = m >= 16;  // long lane ranges: run four intervals behind the operand loads
  const int i0 = jt + lane * m;
  const int i1 = min(i0 + m, nint);

  for (int w = wbeg + warp; w < wend; w += 8) {
    double acc[RT];
#pragma unroll
    for (int r = 0; r < RT; ++r) acc[r] = 0.0;
    const double* xnw = a.xn + (size_t)w * a.wpitch;
    if (MODE == 2) {
      const double* dsw = a.ds1 + (size_t)w * a.wpitch;
      unsigned worst = 0u;
      if (i0 < nint)
        worst = deep ? contract_lane_lean<RT, 4>(xnw, dsw, sK, sL, a.pitch, i0, i1, acc)
                     : contract_lane_lean<RT, 1>(xnw, dsw, sK, sL, a.pitch, i0, i1, acc);
      if (__any_sync(0xffffffffu, worst >= NB_REG_RANGE)) {  // irregular slope somewhere: redo
        if (lane == 0) atomicAdd(&g_fallbacks[0], 1ull);
#pragma unroll
        for (int r = 0; r < RT; ++r) acc[r] = 0.0;
        if (i0 < nint) contract_lane_fast<RT, 1>(xnw, dsw, a.dlx, sK, sL, a.pitch, i0, i1, acc);
      }
    } else if (i0 < nint) {
      if (EXACT)
        contract_lane_exact<RT>(xnw, a.xgrid, sK, a.pitch, i0, i1, acc);
      else if (deep)
        contract_lane_fast<RT, 4>(xnw, a.ds1 + (size_t)w * a.wpitch, a.dlx, sK, sL, a.pitch, i0,
                                  i1, acc);
      else
        contract_lane_fast<RT, 1>(xnw, a.ds1 + (size_t)w * a.wpitch, a.dlx, sK, sL, a.pitch, i0,
                                  i1, acc);
    }
    {  // lane (32 / RT) * r ends up with row r's sum
      double v = warp_sum_rows<RT>(acc, lane);
      const int r = lane / (32 / RT);
      if ((lane & (32 / RT - 1)) == 0 && r < nrows) {
        const int row = row0 + r;
        if (a.coef) v *= a.coef[row];
        a.out[(size_t)w * a.R + row] = v;
      }
    }
  }
  if (tl_row) atomicMax(&tl_row[11], global_timer_ns());
}

// ---------------------------------------------------------------------------
// combine + likelihood: one warp per walker.  Lanes evaluate the model flux and the
// Gaussian terms of the photon energies e = lane, lane + 32, ... into shared memory;
// lane 0 then adds the terms in numpy's summation order (a few hundred cycles).
// ---------------------------------------------------------------------------
constexpr int COMBINE_WARPS = 4;

struct CombineKernelArgs {
  CombineArgs c;
  int has_mv;  // != 0: accept/reject the proposals and append the chain (nb_stretch)
  nb_stretch mv;
  const double* pars;  // [Ns][P] proposals
  int has_peers;       // != 0 (with has_mv): replicated state, the accept step writes every
                       // rank's copy (nb_peers)
  nb_peers peers;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Replicated state: a store that lands in every rank's copy of the symmetric arena that
// holds `local` (one multimem.st through the NVSwitch when a multicast mapping exists,
// else one store per peer, own rank included).
__device__ __forceinline__ int arena_of(const nb_peers& pr, const void* local) {
  const char* p = reinterpret_cast<const char*>(local);
  return (p >= pr.arena_local[1] && p < pr.arena_local[1] + pr.arena_bytes[1]) ? 1 : 0;
}
__device__ __forceinline__ void bcast_store(const nb_peers& pr, double* local, double v) {
  const int k = arena_of(pr, local);
  const size_t off = reinterpret_cast<char*>(local) - pr.arena_local[k];
  if (pr.arena_mc[k]) {
    asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(pr.arena_mc[k] + off), "d"(v)
                 : "memory");
  } else {
    for (int p = 0; p < pr.world; ++p) *reinterpret_cast<double*>(pr.arena_peer[k][p] + off) = v;
  }
}
__device__ __forceinline__ void bcast_store(const nb_peers& pr, int* local, int v) {
  const int k = arena_of(pr, local);
  const size_t off = reinterpret_cast<char*>(local) - pr.arena_local[k];
  if (pr.arena_mc[k]) {
    asm volatile("multimem.st.weak.global.u32 [%0], %1;" ::"l"(pr.arena_mc[k] + off), "r"(v)
                 : "memory");
  } else {
    for (int p = 0; p < pr.world; ++p) *reinterpret_cast<int*>(pr.arena_peer[k][p] + off) = v;
  }
}

__global__ void __launch_bounds__(COMBINE_WARPS * 32) combine_lnprob_kernel(
    const __grid_constant__ CombineKernelArgs ka) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const CombineArgs& a = ka.c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int w = blockIdx.x * COMBINE_WARPS + warp;
  const int t_step = ka.has_mv ? *ka.mv.step : 0;  // before anybody can increment it
  // diagnostic time stamps (see nb_stretch.timeline); the row is fixed before the last CTA
  // can advance the step counter
  unsigned long long* tl_row = (ka.has_mv && ka.mv.timeline && threadIdx.x == 0)
                                   ? ka.mv.timeline + NB_TIMELINE_COLS * ((size_t)(2 * t_step + ka.mv.split) &
                                                           (size_t)(NB_TIMELINE_CAP - 1))
                                   : nullptr;
  if (tl_row && blockIdx.x == 0) tl_row[2] = global_timer_ns();
  // the accept step's operands do not depend on the model: fetch them up front so that
  // their (cold) latency overlaps the component loads
  int sidx = 0;
  double mv_zz = 1.0, mv_lnu = 0.0, lp_old = 0.0;
  size_t mv_base = 0;
  if (ka.has_mv && w < a.W) {
    mv_base = ((size_t)t_step * 2 + ka.mv.split) * ka.mv.Ns + ka.mv.i0 + w;
    sidx = ka.mv.s_idx[mv_base];
    mv_zz = ka.mv.zz[mv_base];
    mv_lnu = ka.mv.lnu[mv_base];
    lp_old = ka.mv.lp[sidx];
  }
  if (w < a.W) {
    // s_t[k]: Gaussian term of the k-th non-upper-limit point (k ascending with e)
    double* s_t = reinterpret_cast<double*>(smem_raw) + (size_t)warp * a.N_E;
    int n = 0, nviol = 0, nul = 0;
    for (int e0 = 0; e0 < a.N_E; e0 += 32) {
      const int e = e0 + lane;
      bool is_pt = false;
      double t = 0.0;
      if (e < a.N_E) {
        // issue the data loads before the model (independent; the store below would
        // otherwise fence them behind the component loads)
        int ule = 0;
        double df = 0.0, elo = 0.0, ehi = 0.0;
        if (a.lnp) {
          ule = a.ul[e];
          df = a.data_flux[e];
          elo = a.err_lo[e];
          ehi = a.err_hi[e];
        }
        double m = combine_model(a, w, e);
        if (a.flux_model) a.flux_model[(size_t)w * a.flux_ld + e] = m;
        if (a.lnp) {
          if (ule) {
            ++nul;
            if (m > df) ++nviol;
          } else {
            is_pt = true;
            t = lnprob_term(m, df, elo, ehi);
          }
        }
      }
      unsigned mask = __ballot_sync(0xffffffffu, is_pt);
      if (is_pt) s_t[n + __popc(mask & ((1u << lane) - 1u))] = t;
      n += __popc(mask);
    }
    if (a.lnp) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        nviol += __shfl_xor_sync(0xffffffffu, nviol, o);
        nul += __shfl_xor_sync(0xffffffffu, nul, o);
      }
      __syncwarp();
      // numpy's summation order (np.sum of core.py:87; see numpy_order_sum): n < 8
      // sequential; else 8 strided accumulators r[j] = sum_i t[8 i + j] over the full
      // blocks of 8, combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) -- exactly the
      // xor-1, xor-2, xor-4 shuffle tree over lanes 0..7 -- then the tail sequentially
      double seq = 0.0;
      const int nblk = n - (n % 8);
      if (n >= 8) {
        double r = 0.0;
        if (lane < 8) {
          r = s_t[lane];
          for (int i = 8 + lane; i < nblk; i += 8) r += s_t[i];
        }
        r += __shfl_xor_sync(0xffffffffu, r, 1);
        r += __shfl_xor_sync(0xffffffffu, r, 2);
        r += __shfl_xor_sync(0xffffffffu, r, 4);
        seq = r;
      }
      double lv = 0.0;
      if (lane == 0) {
        for (int i = (n >= 8 ? nblk : 0); i < n; ++i) seq = (i == 0) ? s_t[0] : seq + s_t[i];
        lv = lnprob_finish(a, w, seq, nul, nviol);
        a.lnp[(size_t)w * a.lnp_ld] = lv;
      }
      if (ka.has_mv) {
        // emcee's accept step for proposal w of the active half, then this walker's
        // row of the chain (its state is final for step t once its half is decided)
        const nb_stretch& mv = ka.mv;
        int acc = 0;
        if (lane == 0) {
          double lnpdiff = (mv.P - 1) * log(mv_zz) + lv - lp_old;
          acc = lnpdiff > mv_lnu;
        }
        acc = __shfl_sync(0xffffffffu, acc, 0);
        __syncwarp();  // this warp's flux_model row is visible to all its lanes
        const size_t W_ = (size_t)mv.W;
        const bool bc = ka.has_peers != 0;
        const nb_peers& pr = ka.peers;
        for (int d = lane; d < mv.P; d += 32) {
          double v = acc ? ka.pars[(size_t)w * mv.pars_ld + d] : mv.coords[(size_t)sidx * mv.P + d];
          if (acc) {
            if (bc) bcast_store(pr, &mv.coords[(size_t)sidx * mv.P + d], v);
            else mv.coords[(size_t)sidx * mv.P + d] = v;
          }
          if (mv.chain) {
            double* dst = &mv.chain[((size_t)t_step * W_ + sidx) * mv.P + d];
            if (bc) bcast_store(pr, dst, v);
            else *dst = v;
          }
        }
        if (mv.nb > 0) {
          for (int d = lane; d < mv.nb; d += 32) {
            double v = acc ? a.flux_model[(size_t)w * a.flux_ld + d]
                           : mv.blobs[(size_t)sidx * mv.nb + d];
            if (acc) {
              if (bc) bcast_store(pr, &mv.blobs[(size_t)sidx * mv.nb + d], v);
              else mv.blobs[(size_t)sidx * mv.nb + d] = v;
            }
            if (mv.chain_blobs) {
              double* dst = &mv.chain_blobs[((size_t)t_step * W_ + sidx) * mv.nb + d];
              if (bc) bcast_store(pr, dst, v);
              else *dst = v;
            }
          }
        }
        if (lane == 0) {
          if (acc) {
            if (bc) bcast_store(pr, &mv.lp[sidx], lv);
            else mv.lp[sidx] = lv;
            // acceptance counts stay local to the deciding rank (replicated mode: the
            // host sums them over ranks; a read-modify-write of replicated data would
            // depend on the arrival order of different ranks' stores)
            mv.n_accepted[sidx] += 1;
          }
          if (mv.chain_lp) {
            // a NaN log-probability is never accepted (the comparison above is false) but
            // must not stay hidden: it goes into the chain row, where the host finds it and
            // raises emcee's "Probability function returned NaN"
            const double rec = (acc || lv != lv) ? lv : lp_old;
            double* dst = &mv.chain_lp[(size_t)t_step * W_ + sidx];
            if (bc) bcast_store(pr, dst, rec);
            else *dst = rec;
          }
        }
      }
    }
  }
  if (ka.has_peers) {
    // replicated state: the accept step above already wrote every rank's copy; the last CTA
    // to finish raises this rank's flag on every peer.  Ordering (PTX memory model, release
    // pattern with cumulativity): each CTA's stores -> bar.sync -> fence.gpu + ticket (thread
    // 0) -> the last CTA observes all tickets -> fence.gpu -> ONE release.sys store of the
    // flag, which makes every store that happens-before it visible to the acquiring peer.
    // The system-scope round trip is paid once per half-step, not once per thread and again
    // by the last CTA.
    const nb_peers& pr = ka.peers;
    const unsigned long long epoch = *pr.gen + 1ull;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      int ticket = atomicAdd(pr.ticket, 1);
      if (ticket == (int)gridDim.x - 1) {
        *pr.ticket = 0;
        __threadfence();
        if (tl_row) tl_row[3] = global_timer_ns();
        if (pr.mc_flags) {
          asm volatile("multimem.st.release.sys.global.u64 [%0], %1;" ::"l"(pr.mc_flags + pr.rank),
                       "l"(epoch)
                       : "memory");
        } else {
          __threadfence_system();
          for (int p = 0; p < pr.world; ++p)
            asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(pr.flags[p] + pr.rank),
                         "l"(epoch)
                         : "memory");
        }
        *pr.gen = epoch;
      }
    }
  }
  if (ka.has_mv && ka.mv.split == 1) {
    // the last CTA to finish advances the step counter (no CTA reads it any more)
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      int ticket = atomicAdd(ka.mv.sync, 1);
      if (ticket == (int)gridDim.x - 1) {
        *ka.mv.sync = 0;
        *ka.mv.step = t_step + 1;
      }
    }
  }
  if (tl_row) atomicMax(&tl_row[4], global_timer_ns());
}

// ---------------------------------------------------------------------------
// parameter map + priors: one thread per walker
// ---------------------------------------------------------------------------
struct ParamMapArgs {
  nb_parmap map[NB_MAX_MAP];
  nb_prior pri[NB_MAX_PRIORS];
  int n_out, n_pri, W, P;
  const double* pars;
  double* out;
  double* prior_out;
};

// ---------------------------------------------------------------------------
// fused per-walker set-up: parameter map + priors, then every particle-distribution
// job (integration operands and/or total particle energy).  One CTA per walker.
// ---------------------------------------------------------------------------
// blockIdx.x = walker, blockIdx.y = work item: (job, first node) for the integration
// operands (256 nodes per CTA) or (job, -1) for a total-energy reduction.  Every CTA
// re-derives its walker's mapped parameters (a handful of pow/exp) so that no CTA waits
// for another; CTA y == 0 also publishes them (and the prior) for the later kernels.
struct PrepItem {
  short job;
  short energy;  // != 0: total-energy item
  int j0;
};
constexpr int NB_MAX_PREP_ITEMS = 48;
constexpr int NB_MAX_MOVE_PAR = 32;
struct WalkerPrepArgs {
  ParamMapArgs pm;
  nb_prep_job jobs[NB_MAX_PREP_JOBS];
  PrepItem items[NB_MAX_PREP_ITEMS];
  int n_jobs, n_items;
  int has_mv;        // != 0: parameters are stretch-move proposals computed here
  nb_stretch mv;
  double* pars_out;  // where CTA y == 0 publishes the proposals
};

__global__ void __launch_bounds__(256) walker_prep_kernel(const __grid_constant__ WalkerPrepArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_node = reinterpret_cast<double*>(smem_raw);  // energy items: n at every node
  __shared__ double s_pm[NB_MAX_MAP];
  __shared__ double s_n[PREP_CHUNK + 1];
  __shared__ PdNode s_nd[PREP_CHUNK + 1];
  __shared__ double s_red[256];
  const int w = blockIdx.x;
  const int tid = threadIdx.x;
  const bool publish = blockIdx.y == 0;
  const double* p = a.pm.pars + (size_t)w * a.pm.P;
  __shared__ double s_q[NB_MAX_MOVE_PAR];
  const bool setup_launch = a.pm.prior_out != nullptr;  // else: the total-energy blob launch
  if (tid == 0 && w == 0 && blockIdx.y == 0) {
    unsigned long long* row = (a.has_mv && a.mv.timeline) ? timeline_row(a.mv) : nullptr;
    if (setup_launch) g_timeline_row = row;
    else if (row) row[8] = global_timer_ns();
  }
  if (a.has_mv) {
    wait_for_peers(a.mv, setup_launch);
    // emcee stretch move: q = c - (c - s) zz, numpy's rounding (no FMA contraction)
    if (tid < a.pm.P) {
      const size_t base = ((size_t)(*a.mv.step) * 2 + a.mv.split) * a.mv.Ns + a.mv.i0 + w;
      double c = a.mv.coords[(size_t)a.mv.c_idx[base] * a.pm.P + tid];
      double sv = a.mv.coords[(size_t)a.mv.s_idx[base] * a.pm.P + tid];
      double q = __dsub_rn(c, __dmul_rn(__dsub_rn(c, sv), a.mv.zz[base]));
      s_q[tid] = q;
      if (publish) a.pars_out[(size_t)w * a.mv.pars_ld + tid] = q;
    }
    __syncthreads();
    p = s_q;
  }
  if (tid < a.pm.n_out) {
    const nb_parmap& m = a.pm.map[tid];
    double v = m.scale;
    if (m.src >= 0) {
      double x = p[m.src];
      if (m.fn == NB_FN_POW10) x = exp10(x);
      else if (m.fn == NB_FN_EXP) x = exp(x);
      v = x * m.scale;
    }
    s_pm[tid] = v;
    if (publish) a.pm.out[m.dst_off + (long long)w * m.dst_stride] = v;
  }
  if (publish && tid == 32 && a.pm.prior_out) {
    double lp = 0.0;
    for (int k = 0; k < a.pm.n_pri; ++k)
      lp += prior_eval(a.pm.pri[k].kind, p[a.pm.pri[k].par], a.pm.pri[k].a, a.pm.pri[k].b);
    a.pm.prior_out[w] = lp;
  }
  if ((int)blockIdx.y >= a.n_items) return;
  __syncthreads();
  const PrepItem it = a.items[blockIdx.y];
  const nb_prep_job& J = a.jobs[it.job];
  // this item's distribution parameters: lane q < 8 of warp 0 finds parameter q among the
  // mapped entries, thread 0 derives the log-space constants once for the whole CTA
  __shared__ double s_pp[PD_MAXPAR];
  __shared__ PdLog s_S;
  if (tid < PD_MAXPAR) {
    double v = 0.0;
    for (int k = 0; k < a.pm.n_out; ++k)
      if (a.pm.map[k].dst_stride == PD_MAXPAR && a.pm.map[k].dst_off - J.pd_off == tid)
        v = s_pm[k];
    s_pp[tid] = v;
  }
  __syncthreads();
  if (tid == 0) s_S = pd_log_setup(J.kind, s_pp, J.n_scale);
  __syncthreads();
  const double* pp = s_pp;
  const PdLog& S = s_S;
  if (!it.energy) {
    pd_prep_chunk(J.kind, pp, S, J.x, J.N, J.e_mul1, J.e_mul2, J.n_scale, J.invdlx, J.xn, J.ds1,
                  J.nraw, (size_t)w * J.wpitch, it.j0, s_n, s_nd);
  } else {
    for (int i = tid; i < J.N; i += 256) {
      double e = (J.x[i] * J.e_mul1) * J.e_mul2;
      s_node[i] = pd_log_value(S, pd_log_node(S, e));
    }
    __syncthreads();
    double acc = 0.0;
    for (int i = tid; i < J.N - 1; i += 256) {
      double x1 = J.x[i], x2 = J.x[i + 1];
      acc += interval_exact(x1 * J.x_to_energy, x2 * J.x_to_energy, x1 * s_node[i],
                            x2 * s_node[i + 1]);
    }
    s_red[tid] = acc;
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
      if (tid < s2) s_red[tid] += s_red[tid + s2];
      __syncthreads();
    }
    if (tid == 0) J.energy_out[(size_t)w * (J.energy_stride > 0 ? J.energy_stride : 1)] = s_red[0];
  }
  if (a.has_mv && a.mv.timeline && tid == 0)
    atomicMax(&timeline_row(a.mv)[setup_launch ? 7 : 9], global_timer_ns());
}

// ---------------------------------------------------------------------------
// self-contained component kernels: walker parameters derived per warp
// ---------------------------------------------------------------------------
struct WalkerSrc {
  ParamMapArgs pm;  // pars / P / map / n_out (out, priors unused)
  int has_mv;
  nb_stretch mv;
};

struct PdDesc {
  int kind;
  long long pd_off;
  double e_mul1, e_mul2, n_scale;
  const double* lnx;
  const double* invdlx;
};

// All 32 lanes of a warp cooperate and all receive: the NB_PD_MAXPAR parameters of the
// distribution block at pd_off and the mapped value of entry scalar_entry, for walker w.
__device__ __forceinline__ void warp_walker_params(const WalkerSrc& s, int w, long long pd_off,
                                                   int scalar_entry, double* pp,
                                                   double* scalar) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int P = s.pm.P;
  double qv = 0.0;
  if (lane < P) {
    if (s.has_mv) {
      const size_t base = ((size_t)(*s.mv.step) * 2 + s.mv.split) * s.mv.Ns + s.mv.i0 + w;
      double c = s.mv.coords[(size_t)s.mv.c_idx[base] * P + lane];
      double sv = s.mv.coords[(size_t)s.mv.s_idx[base] * P + lane];
      qv = __dsub_rn(c, __dmul_rn(__dsub_rn(c, sv), s.mv.zz[base]));
    } else {
      qv = s.pm.pars[(size_t)w * P + lane];
    }
  }
  const bool active = lane < s.pm.n_out;
  int src = 0, fn = 0, stride = 0;
  long long off = -1;
  double scale = 0.0;
  if (active) {
    const nb_parmap& m = s.pm.map[lane];
    src = m.src;
    fn = m.fn;
    scale = m.scale;
    off = m.dst_off;
    stride = m.dst_stride;
  }
  double x = __shfl_sync(FULL, qv, src >= 0 ? src : 0);
  double v = scale;
  if (active && src >= 0) {
    if (fn == NB_FN_POW10) x = exp10(x);
    else if (fn == NB_FN_EXP) x = exp(x);
    v = x * scale;
  }
#pragma unroll
  for (int k = 0; k < PD_MAXPAR; ++k) {
    const bool mine = active && stride == PD_MAXPAR && off == pd_off + k;
    const unsigned b = __ballot_sync(FULL, mine);
    const double t = __shfl_sync(FULL, v, b ? (__ffs(b) - 1) : 0);
    pp[k] = b ? t : 0.0;
  }
  const double sc = __shfl_sync(FULL, v, scalar_entry >= 0 ? scalar_entry : 0);
  *scalar = scalar_entry >= 0 ? sc : 0.0;
}

// ---------------------------------------------------------------------------
// synchrotron: CTA = (walker, strided slice of photon energies)
// ---------------------------------------------------------------------------
struct SynArgs {
  const double* gam;
  int N;
  const double* gm2;  // g^-2 per node, or NULL (computed here)
  const double* g23;  // cbrt(g^-2) per node, or NULL
  const double* xn;   // operand arrays [W][wpitch] (plain kernel; NULL in the fused kernel)
  const double* ds1;
  int wpitch;
  const double* invdlx;
  const double* dlx;
  const double* B;    // [W] (plain kernel)
  int W;
  const double* E_erg;
  const double* cbrtE;  // cbrt(E_erg) per photon energy, or NULL (computed per lane)
  int N_E;
  double* out;
  int out_ld;         // row pitch of out (>= N_E)
  int e_per_cta;
};

struct SynFusedArgs {
  SynArgs a;
  WalkerSrc src;
  PdDesc pd;
  int b_entry;
};

// FUSED: warp 0 derives the walker's parameters (proposal -> parameter map -> log-space
// constants) and the CTA evaluates the particle distribution at the nodes itself; else the
// operands come from nb_pd_prep's arrays.
template <bool FUSED>
__device__ __forceinline__ void synchrotron_cta(const SynArgs& a, const SynFusedArgs* fa) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // per node: 1/Ec, cbrt(1/Ec), x*n, ds1, invdlx, dlx; per photon energy: first live node
  double* s_iec = reinterpret_cast<double*>(smem_raw);
  double* s_cb = s_iec + a.N;
  double* s_xn = s_cb + a.N;
  double* s_ds = s_xn + a.N;
  double* s_idl = s_ds + a.N;
  double* s_dl = s_idl + a.N;
  int* s_js = reinterpret_cast<int*>(s_dl + a.N);  // [e_per_cta]
  __shared__ int s_jmin;
  __shared__ double s_B;
  __shared__ PdLog s_S;

  const int w = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nint = a.N - 1;
  // this CTA's photon energies: e = blockIdx.y + k * gridDim.y, k < ne (strided, so that
  // every slice gets the same mix of cheap and expensive energies)
  const int nsl = gridDim.y;
  const int ne = (a.N_E - (int)blockIdx.y + nsl - 1) / nsl;
  unsigned long long* tl_row = nullptr;  // diagnostic stamps [5] entry of CTA (0,0), [6] last end
  if (FUSED && fa->src.has_mv && fa->src.mv.timeline && threadIdx.x == 0) {
    tl_row = timeline_row(fa->src.mv);
    if (blockIdx.x == 0 && blockIdx.y == 0) tl_row[5] = global_timer_ns();
  }
  if (FUSED) {
    if (fa->src.has_mv) wait_for_peers(fa->src.mv);
    if (warp == 0) {
      double pp[PD_MAXPAR], Bv;
      warp_walker_params(fa->src, w, fa->pd.pd_off, fa->b_entry, pp, &Bv);
      if (lane == 0) {
        PdLog S = pd_log_setup(fa->pd.kind, pp, fa->pd.n_scale);
        pd_log_setup_grid(S, fa->pd.e_mul1, fa->pd.e_mul2);
        s_S = S;
        s_B = Bv;
        s_jmin = a.N;
      }
    }
  } else if (threadIdx.x == 0) {
    s_B = a.B[w];
    s_jmin = a.N;
  }
  __syncthreads();
  const double Bw = s_B;
  // nodes whose exp(-E/Ec) underflows to zero contribute nothing: find, per photon
  // energy, the first node that can be non-zero, and only set up nodes from the
  // smallest of them on
  for (int k = threadIdx.x; k < ne; k += blockDim.x) {
    int js = syn_first_node(a.gam, a.N, Bw, a.E_erg[blockIdx.y + k * nsl]);
    s_js[k] = js;
    atomicMin(&s_jmin, js);
  }
  __syncthreads();
  const int jmin = s_jmin;
  double ikB, cbk;
  syn_walker(Bw, &ikB, &cbk);
  for (int j = jmin + threadIdx.x; j < a.N; j += blockDim.x) {
    const double g = a.gam[j];
    if (a.gm2) {
      s_iec[j] = ikB * a.gm2[j];
      s_cb[j] = cbk * a.g23[j];
    } else {
      syn_node(g, Bw, &s_iec[j], &s_cb[j]);
    }
    if (FUSED) {
      const PdNode nd = pd_log_node_tab(s_S, g, fa->pd.lnx[j]);
      s_xn[j] = g * pd_log_value_fast(s_S, nd);
      if (j < nint) {
        const double idl = a.invdlx[j];
        const PdNode nd2 = pd_log_node_tab(s_S, a.gam[j + 1], fa->pd.lnx[j + 1]);
        s_ds[j] = pd_log_ds1(s_S, nd, nd2, idl);
        s_idl[j] = idl;
        s_dl[j] = a.dlx[j];
      }
    } else {
      s_xn[j] = a.xn[(size_t)w * a.wpitch + j];
      s_ds[j] = a.ds1[(size_t)w * a.wpitch + j];
      if (j < nint) {
        s_idl[j] = a.invdlx[j];
        s_dl[j] = a.dlx[j];
      }
    }
  }
  __syncthreads();

  // each photon energy is integrated by a pair of warps (64 lanes, ~5 intervals per lane
  // on the default grids): the per-lane chains are serial, so short chains and many warps
  // are what keeps the fp64 pipe busy.  The two partial sums meet in shared memory.
  double* s_part = reinterpret_cast<double*>(s_js + a.e_per_cta + (a.e_per_cta & 1));  // [epc][2]
  const int pair = warp >> 1, half = warp & 1;
  for (int k = pair; k < ne; k += 4) {
    const double E = a.E_erg[blockIdx.y + k * nsl];
    const int js = s_js[k];
    const int len = nint - js;
    double acc = 0.0;
    if (len > 0) {
      const double cbE = a.cbrtE ? a.cbrtE[blockIdx.y + k * nsl] : cbrt(E);
      const int m = odd_chunk2(len);
      const int i0 = js + (half * 32 + lane) * m;
      const int i1 = min(i0 + m, nint);
      if (i0 < nint) acc = syn_lane(E, cbE, s_iec, s_cb, s_xn, s_ds, s_idl, s_dl, i0, i1);
      acc = warp_sum(acc);
    }
    if (lane == 0) s_part[2 * k + half] = acc;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < ne; k += blockDim.x) {
    const int e = blockIdx.y + k * nsl;
    const double acc = s_part[2 * k] + s_part[2 * k + 1];
    a.out[(size_t)w * a.out_ld + e] = syn_finish(Bw, a.E_erg[e], acc);
  }
  if (tl_row) atomicMax(&tl_row[6], global_timer_ns());
}

__global__ void __launch_bounds__(256) synchrotron_kernel(const __grid_constant__ SynArgs a) {
  synchrotron_cta<false>(a, nullptr);
}

__global__ void __launch_bounds__(256) synchrotron_fused_kernel(
    const __grid_constant__ SynFusedArgs fa) {
  synchrotron_cta<true>(fa.a, &fa);
}

// ---------------------------------------------------------------------------
// IC on a tabulated seed, fused: CTA = (photon energy e, walker w)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ic_seed_spectrum_kernel(
    const double* __restrict__ gam, int N, const double* __restrict__ nraw, int wpitch,
    const double* __restrict__ Eph, const double* __restrict__ eps0,
    const double* __restrict__ phn, int Ns, int phn_wstride, double* __restrict__ out,
    int out_ld, int out_off) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_e0 = reinterpret_cast<double*>(smem_raw);  // [Ns]
  double* s_ph = s_e0 + Ns;                            // [Ns] phn/eps0
  double* s_y = s_ph + Ns;                             // [N]  n_e * K
  double* s_red = s_y + N;                             // [8]  per-warp partial sums
  const int e = blockIdx.x, w = blockIdx.y;
  const double* ph = phn + (size_t)w * phn_wstride;
  for (int s = threadIdx.x; s < Ns; s += blockDim.x) {
    double x = eps0[s];
    s_e0[s] = x;
    s_ph[s] = ph[s] / x;
  }
  __syncthreads();
  const double ep = Eph[e];
  for (int j = threadIdx.x; j < N; j += blockDim.x) {
    double g = gam[j];
    double v;
    if (Ns == 1) {
      v = ic_mono_f(g, s_e0[0], ep) * (s_ph[0] / s_e0[0]);
    } else {
      double x1 = s_e0[0];
      double y1 = ic_mono_f(g, x1, ep) * s_ph[0];
      double acc = 0.0;
      for (int s = 1; s < Ns; ++s) {
        double x2 = s_e0[s];
        double y2 = ic_mono_f(g, x2, ep) * s_ph[s];
        acc += interval_exact(x1, x2, y1, y2);
        x1 = x2;
        y1 = y2;
      }
      v = acc;
    }
    v *= (3.0 / 4.0) * SIGT * 29979245800.0 / (g * g);
    s_y[j] = nraw[(size_t)w * wpitch + j] * v;
  }
  __syncthreads();
  double acc = 0.0;
  for (int j = threadIdx.x; j < N - 1; j += blockDim.x)
    acc += interval_exact(gam[j], gam[j + 1], s_y[j], s_y[j + 1]);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += s_red[k];
    out[(size_t)w * out_ld + out_off + e] = ep * t;
  }
}

// ---------------------------------------------------------------------------
// synchrotron self-Compton: IC on a seed photon field that differs per walker
// (radiative.py:609-655 with the seed density of examples/CrabNebula_SynSSC.py:24-28)
//
//   spec[w][e] = Eph_e/E_e * trapz_g( n_e[w][g] * 3/4 sigma_T c / g^2 *
//                       trapz_s( F[e][g][s] * phn[w][s] / eps0_s , eps0 ), gam )
//
// F = f_AA81(gam_g, eps0_s, Eph_e) (incl. its two step functions) does not depend on the
// walker: it is tabulated once, s-major (Ft[s][r], r = e*N + g, so that consecutive threads
// = consecutive rows read consecutive addresses), together with its log-slopes along s.
// Per half-step:  ssc_seed_kernel (seed density and its slopes from the synchrotron
// luminosities) -> ssc_inner_kernel (the N_s-long inner trapezoid of every (e, g) row for
// all walkers: 8.7 M rows x intervals per walker at C4, the heavy part) -> ssc_outer_kernel
// (the outer trapezoid over gam).
// ---------------------------------------------------------------------------
__global__ void ssc_table_kernel(const double* __restrict__ gam, int N,
                                 const double* __restrict__ Eph, int N_E,
                                 const double* __restrict__ eps0,
                                 const double* __restrict__ invdlx_s, int Ns,
                                 double2* __restrict__ KL, double* __restrict__ F0,
                                 double* __restrict__ coef, long long Rp) {
  // KL[s][r] = (F[s+1][r], log-slope of F over interval s): what cell s of row r needs, in
  // one 128-bit load; F0[r] = F[0][r]
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= Rp) return;
  const long long R = (long long)N_E * N;
  if (r >= R) {  // padding rows: zero integrand, sentinel slope
    for (int s = 0; s < Ns - 1; ++s) KL[s * Rp + r] = make_double2(0.0, NB_BIG_SLOPE);
    F0[r] = 0.0;
    coef[r] = 0.0;
    return;
  }
  const int e = (int)(r / N), j = (int)(r - (long long)e * N);
  const double g = gam[j], ep = Eph[e];
  double f1 = ic_mono_f(g, eps0[0], ep);
  F0[r] = f1;
  for (int s = 1; s < Ns; ++s) {
    const double f2 = ic_mono_f(g, eps0[s], ep);
    KL[(s - 1) * Rp + r] = make_double2(f2, slope_or_sentinel(f1, f2, invdlx_s[s - 1]));
    f1 = f2;
  }
  coef[r] = (3.0 / 4.0) * SIGT * 29979245800.0 / (g * g);
}

struct SscSeedArgs {
  const double* src[NB_SSC_MAX_SRC];  // luminosities [W][ld] in 1/(s eV)
  int ld[NB_SSC_MAX_SRC], off[NB_SSC_MAX_SRC];
  double fac[NB_SSC_MAX_SRC];         // -> dn/dE in 1/(mec2 cm3)
  int n_src, W, Ns, spitch;
  const double* invdlx_s;
  double* sxn;  // [W][spitch] seed density (the contraction's x*y operand: eps0 * phn/eps0)
  double* sds;  // [W][spitch] its slope term d ln(phn/eps0)/d ln eps0 + 1
};

__global__ void ssc_seed_kernel(const __grid_constant__ SscSeedArgs a) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
  if (s >= a.spitch) return;
  double p1 = 0.0, p2 = 0.0;
  if (s < a.Ns)
    for (int k = 0; k < a.n_src; ++k) {
      const double* row = a.src[k] + (size_t)w * a.ld[k] + a.off[k];
      p1 += a.fac[k] * row[s];
      if (s + 1 < a.Ns) p2 += a.fac[k] * row[s + 1];
    }
  a.sxn[(size_t)w * a.spitch + s] = p1;
  a.sds[(size_t)w * a.spitch + s] = (s + 1 < a.Ns) ? slope_or_sentinel(p1, p2, a.invdlx_s[s]) : 0.0;
}

struct SscInnerArgs {
  const double2* KL;   // [Ns-1][Rp]
  const double* F0;    // [Rp]
  const double* coef;  // [Rp]
  long long Rp;        // row pitch of the s-major table (multiple of 128)
  int Ns;
  const double* sxn;
  const double* sds;
  int spitch, W;
  const double* dlx_s;  // [Ns-1] ln(eps0[s+1]/eps0[s]) (careful cell only)
  double* inner;        // [W][Rp]
};

// thread = one (e, g) row, WT walkers in registers; the walkers' seed operands sit in shared
// memory as (x*y at s+1, slope at s) pairs: one 128-bit broadcast load per cell.  The row's
// table entries stream from L2 two intervals ahead of their use.
constexpr int SSC_STAGES = 8;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int WT>
__global__ void __launch_bounds__(128, 6) ssc_inner_kernel(
    const __grid_constant__ SscInnerArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double2* s_op = reinterpret_cast<double2*>(smem_raw);           // [Ns][WT]
  double2* s_ring = s_op + (size_t)WT * a.Ns;                     // [SSC_STAGES][128]
  double* s_x0 = reinterpret_cast<double*>(s_ring + SSC_STAGES * 128);  // [WT]
  const int w0 = blockIdx.x * WT;
  const int Ns = a.Ns;
  for (int k = threadIdx.x; k < WT * Ns; k += blockDim.x) {
    const int s = k / WT, wl = k - s * WT, w = w0 + wl;
    double2 v = make_double2(0.0, NB_BIG_SLOPE);
    if (w < a.W && s + 1 < Ns)
      v = make_double2(a.sxn[(size_t)w * a.spitch + s + 1], a.sds[(size_t)w * a.spitch + s]);
    s_op[k] = v;
  }
  if (threadIdx.x < WT)
    s_x0[threadIdx.x] = (w0 + (int)threadIdx.x < a.W) ? a.sxn[(size_t)(w0 + threadIdx.x) * a.spitch]
                                                      : 0.0;
  __syncthreads();
  const long long r = (long long)blockIdx.y * blockDim.x + threadIdx.x;  // < Rp by construction
  const double2* KLc = a.KL + r;
  double acc[WT], prev[WT];
  unsigned worst = 0u;
  const double k1 = a.F0[r];
#pragma unroll
  for (int w = 0; w < WT; ++w) {
    acc[w] = 0.0;
    prev[w] = s_x0[w] * k1;
  }
  const int nint = Ns - 1;
  // The row's table entries reach the thread through a ring of SSC_STAGES 16-byte slots of its
  // own in shared memory, filled by cp.async SSC_STAGES - 1 intervals ahead: no registers are
  // tied up by loads in flight and the L2 latency of the stream is off the dependency chain
  // (held in registers two intervals ahead, the loads were the kernel's first stall reason).
  // Every thread reads only the slots it wrote itself: cp.async.wait_group is all the
  // synchronisation needed.
  double2* ring = s_ring + threadIdx.x;  // slot of stage g: ring[g * 128]
#pragma unroll
  for (int g = 0; g < SSC_STAGES - 1; ++g) {
    if (g < nint) cp_async16(ring + g * 128, KLc + (long long)g * a.Rp);
    cp_async_commit();
  }
  const double2* op_s = s_op;
  const double2* src = KLc + (long long)(SSC_STAGES - 1) * a.Rp;  // next entry to fetch
  // one interval: wait for its slot, refill the slot of the previous interval, WT cells
  auto interval = [&](const int s, const int slot) {
    cp_async_wait<SSC_STAGES - 2>();  // the group of interval s has landed
    const double2 kl = ring[slot * 128];
    if (s + SSC_STAGES - 1 < nint)
      cp_async16(ring + ((slot + SSC_STAGES - 1) % SSC_STAGES) * 128, src);
    cp_async_commit();
    src += a.Rp;
    const double k2 = kl.x, l = kl.y;
#pragma unroll
    for (int w = 0; w < WT; ++w) {
      const double2 op = op_s[w];
      const double xy2 = op.x * k2;
      cell_lean(prev[w], xy2, op.y + l, acc[w], worst);
      prev[w] = xy2;
    }
    op_s += WT;
  };
  int s = 0;
  for (; s + SSC_STAGES <= nint; s += SSC_STAGES) {  // slot numbers are compile-time constants
#pragma unroll
    for (int g = 0; g < SSC_STAGES; ++g) interval(s + g, g);
  }
  for (int g = 0; s < nint; ++s, ++g) interval(s, g);  // fewer than SSC_STAGES left
  const double cf = a.coef[r];
  if (worst >= NB_REG_RANGE) {
    // an irregular slope somewhere on this row (sign change in the table, |b+1| <= 1e-10,
    // NaN operands): redo the row with the careful cell
    atomicAdd(&g_fallbacks[1], 1ull);
#pragma unroll 1
    for (int w = 0; w < WT; ++w) {
      double xy1 = s_x0[w] * k1, t = 0.0;
      for (int s = 0; s < nint; ++s) {
        const double2 op = s_op[s * WT + w];
        const double2 kl = KLc[(long long)s * a.Rp];
        const double xy2 = op.x * kl.x;
        t += interval_fast(xy1, xy2, op.y + kl.y, a.dlx_s[s]);
        xy1 = xy2;
      }
      if (w0 + w < a.W) a.inner[(size_t)(w0 + w) * a.Rp + r] = t * cf;
    }
    return;
  }
#pragma unroll
  for (int w = 0; w < WT; ++w)
    if (w0 + w < a.W) a.inner[(size_t)(w0 + w) * a.Rp + r] = acc[w] * cf;
}

struct SscOuterArgs {
  const double* inner;  // [W][Rp], row e*N + g
  long long Rp;
  int N, N_E, W;
  const double* xn;     // electron operands on the gam grid [W][wpitch]
  const double* ds1;
  int wpitch;
  const double* dlx;
  const double* invdlx;
  const double* coef_e;  // [N_E] Eph/E_eV
  double* out;
  int out_ld, out_off;
};

// one warp per (walker, photon energy): outer trapezoid over gam, careful cell (the inner
// integral's own slope needs one log per node)
__global__ void __launch_bounds__(256) ssc_outer_kernel(const __grid_constant__ SscOuterArgs a) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (gw >= a.W * a.N_E) return;
  const int w = gw / a.N_E, e = gw - w * a.N_E;
  const double* in = a.inner + (size_t)w * a.Rp + (size_t)e * a.N;
  const double* xn = a.xn + (size_t)w * a.wpitch;
  const double* ds = a.ds1 + (size_t)w * a.wpitch;
  double acc = 0.0;
  for (int j = lane; j < a.N - 1; j += 32) {
    const double y1 = in[j], y2 = in[j + 1];
    // ln(y2 / y1) of neighbouring nodes by the atanh series (log when they differ by more
    // than 10 %, or at zero / NaN end points, where interval_fast decides anyway)
    const double bp1 = ds[j] + log_ratio(y1, y2) * a.invdlx[j];
    acc += interval_fast(xn[j] * y1, xn[j + 1] * y2, bp1, a.dlx[j]);
  }
  acc = warp_sum(acc);
  if (lane == 0) a.out[(size_t)w * a.out_ld + a.out_off + e] = a.coef_e[e] * acc;
}

// ---------------------------------------------------------------------------
// pion decay, Kelner+06: every photon energy integrates over its OWN proton-energy grid
// (the lower limit depends on the photon energy), so the particle distribution is evaluated
// per (walker, row, node).  CTA = (row, walker); rows carry their grid Ep[r][N] (TeV) and the
// walker-independent integrand kernel Kk[r][N]; reference-order trapezoid, fixed-order sum.
// ---------------------------------------------------------------------------
__global__ void kelner_table_kernel(const double* __restrict__ Eg, const int* __restrict__ hi,
                                    int R, int N, double decades, double* __restrict__ Ep,
                                    double* __restrict__ Kk) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (j >= N || r >= R) return;
  const double eg = Eg[r];
  double e0;
  if (hi[r]) {
    e0 = eg;
  } else {
    const double Epimin = eg + KEL_MPI_TEV * KEL_MPI_TEV / (4 * eg);
    e0 = KEL_MP_TEV + Epimin / KEL_KPI;
  }
  const double ep = e0 * exp10(decades * j / (N - 1));
  Ep[(size_t)r * N + j] = ep;
  Kk[(size_t)r * N + j] = hi[r] ? kel_kernel_hi(ep, eg) : kel_kernel_lo(ep);
}

__global__ void __launch_bounds__(256) kelner_rows_kernel(
    int kind, const double* __restrict__ params, const double* __restrict__ Ep,
    const double* __restrict__ Kk, int R, int N, double* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_y = reinterpret_cast<double*>(smem_raw);  // [N]
  __shared__ double s_red[256];
  const int r = blockIdx.x, w = blockIdx.y;
  double p[PD_MAXPAR];
#pragma unroll
  for (int k = 0; k < PD_MAXPAR; ++k) p[k] = params[w * PD_MAXPAR + k];
  const double* ep = Ep + (size_t)r * N;
  const double* kk = Kk + (size_t)r * N;
  // J per TeV: PD.eval(E [eV]) [1/eV] * 1e12  (radiative.py:1589-1590)
  for (int j = threadIdx.x; j < N; j += blockDim.x)
    s_y[j] = pd_eval(kind, p, ep[j] * 1e12) * 1e12 * kk[j];
  __syncthreads();
  double acc = 0.0;
  for (int j = threadIdx.x; j < N - 1; j += blockDim.x)
    acc += interval_exact(ep[j], ep[j + 1], s_y[j], s_y[j + 1]);
  s_red[threadIdx.x] = acc;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2) s_red[threadIdx.x] += s_red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(size_t)w * R + r] = s_red[0];
}

// ---------------------------------------------------------------------------
// accept step + chain append from all-gathered packed records (walker sharding): one
// warp per proposal of the active half, identical on every rank
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) stretch_update_packed_kernel(
    const __grid_constant__ nb_stretch mv, const double* __restrict__ pack, int ld) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = blockIdx.x * 4 + warp;
  const int t_step = *mv.step;
  if (i < mv.Ns) {
    const size_t base = ((size_t)t_step * 2 + mv.split) * mv.Ns + i;
    const int sidx = mv.s_idx[base];
    const double* rec = pack + (size_t)i * ld;
    const double lv = rec[mv.nb];
    const double lp_old = mv.lp[sidx];
    const double lnpdiff = (mv.P - 1) * log(mv.zz[base]) + lv - lp_old;
    const int acc = lnpdiff > mv.lnu[base];
    const size_t W_ = (size_t)mv.W;
    for (int d = lane; d < mv.P; d += 32) {
      double v = acc ? rec[mv.nb + 1 + d] : mv.coords[(size_t)sidx * mv.P + d];
      if (acc) mv.coords[(size_t)sidx * mv.P + d] = v;
      if (mv.chain) mv.chain[((size_t)t_step * W_ + sidx) * mv.P + d] = v;
    }
    for (int d = lane; d < mv.nb; d += 32) {
      double v = acc ? rec[d] : mv.blobs[(size_t)sidx * mv.nb + d];
      if (acc) mv.blobs[(size_t)sidx * mv.nb + d] = v;
      if (mv.chain_blobs) mv.chain_blobs[((size_t)t_step * W_ + sidx) * mv.nb + d] = v;
    }
    __syncwarp();
    if (lane == 0) {
      if (acc) {
        mv.lp[sidx] = lv;
        mv.n_accepted[sidx] += 1;
      }
      // a NaN log-probability is recorded in the chain (the host raises on it)
      if (mv.chain_lp) mv.chain_lp[(size_t)t_step * W_ + sidx] = (acc || lv != lv) ? lv : lp_old;
    }
  }
  if (mv.split == 1) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      int ticket = atomicAdd(mv.sync, 1);
      if (ticket == (int)gridDim.x - 1) {
        *mv.sync = 0;
        *mv.step = t_step + 1;
      }
    }
  }
}

__global__ void peer_wait_kernel(const __grid_constant__ nb_stretch mv) { wait_for_peers(mv); }

// ---------------------------------------------------------------------------
// fp64 FMA throughput probe (roofline denominator, measured by the caller)
// ---------------------------------------------------------------------------
__global__ void fp64_probe_kernel(double* out, int iters) {
  double a[16];
  const double m = 1.0000001, c = 1e-9;
#pragma unroll
  for (int k = 0; k < 16; ++k) a[k] = threadIdx.x * 1e-3 + k;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
  }
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < 16; ++k) t += a[k];
  if (t == -1.0) out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}


// ---------------------------------------------------------------------------
// a1 as a stand-alone op: trapz_loglog of R rows y[r][0..N) over x (shared x[N]
// or per-row x[r][0..N) when x_ld > 0), reference operation order; warp per row
// ---------------------------------------------------------------------------
__global__ void trapz_loglog_kernel(const double* __restrict__ y, int R, int N, int ld,
                                    const double* __restrict__ x, int x_ld,
                                    double* __restrict__ out, double* __restrict__ iv) {
  int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (r >= R) return;
  const double* yr = y + (size_t)r * ld;
  const double* xr = x + (size_t)r * x_ld;
  double acc = 0.0;
  for (int i = lane; i < N - 1; i += 32) {
    double v = interval_exact(xr[i], xr[i + 1], yr[i], yr[i + 1]);
    if (iv) iv[(size_t)r * (N - 1) + i] = v;
    acc += v;
  }
  acc = warp_sum(acc);
  if (lane == 0 && out) out[r] = acc;
}

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int nb_version(void) { return 100; }

const char* nb_strerror(int code) {
  if (code == 0) return "ok";
  if (code == NB_EINVAL) return "invalid argument";
  if (code == NB_ETOOLARGE) return "grid too large for the shared-memory tiling";
  if (code == NB_EALIGN) return "alignment requirement violated";
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "unknown error";
}

int nb_contract_smem_bytes(int N, int rows_per_tile) {
  int pitch = (N + 1) & ~1;
  long long b = 2LL * rows_per_tile * pitch * 8 + 16;
  return b > 227 * 1024 ? 0 : (int)b;
}

int nb_pdist_eval_ld(int kind, const double* pd_params, int W, const double* e_eV, int N,
                     double* out, int out_ld, void* stream) {
  if (!pd_params || !e_eV || !out || W < 0 || N < 0 || kind < 0 || kind > NB_PD_LOGPAR ||
      out_ld < N)
    return NB_EINVAL;
  if (W == 0 || N == 0) return 0;
  dim3 grid((N + 255) / 256, W);
  pdist_eval_kernel<<<grid, 256, 0, as_stream(stream)>>>(kind, pd_params, W, e_eV, N, out,
                                                         out_ld);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_pdist_eval(int kind, const double* pd_params, int W, const double* e_eV, int N,
                  double* out, void* stream) {
  return nb_pdist_eval_ld(kind, pd_params, W, e_eV, N, out, N, stream);
}

// nraw is an extension used by the exact contraction: n[w][j] itself
int nb_pd_prep_ex(int kind, const double* pd_params, int W, const double* x, int N,
                  double e_mul1, double e_mul2, double n_scale, const double* invdlx,
                  double* xn, double* ds1, double* nraw, int wpitch, void* stream) {
  if (!pd_params || !x || !invdlx || !xn || !ds1 || W < 0 || N < 2 || wpitch < N ||
      kind < 0 || kind > NB_PD_LOGPAR)
    return NB_EINVAL;
  if (W == 0) return 0;
  dim3 grid((N + PREP_CHUNK - 1) / PREP_CHUNK, W);
  pd_prep_kernel<<<grid, 256, 0, as_stream(stream)>>>(kind, pd_params, W, x, N, e_mul1, e_mul2,
                                                      n_scale, invdlx, xn, ds1, nraw, wpitch);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_pd_prep(int kind, const double* pd_params, int W, const double* x, int N, double e_mul1,
               double e_mul2, double n_scale, const double* invdlx, double* xn, double* ds1,
               int wpitch, void* stream) {
  return nb_pd_prep_ex(kind, pd_params, W, x, N, e_mul1, e_mul2, n_scale, invdlx, xn, ds1,
                       nullptr, wpitch, stream);
}

int nb_particle_energy(int kind, const double* pd_params, int W, const double* x, int N,
                       double e_mul1, double e_mul2, double n_scale, double x_to_energy,
                       double* out, void* stream) {
  if (!pd_params || !x || !out || W < 0 || N < 2 || kind < 0 || kind > NB_PD_LOGPAR)
    return NB_EINVAL;
  if (W == 0) return 0;
  particle_energy_kernel<<<W, 128, 0, as_stream(stream)>>>(kind, pd_params, x, N, e_mul1,
                                                           e_mul2, n_scale, x_to_energy, out);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_ic_planck_table(const double* gam, int N, const double* Eph, int N_E,
                       const double* seed_T, const double* seed_theta, int S, double* K,
                       int pitch, int row0, void* stream) {
  if (!gam || !Eph || !seed_T || !seed_theta || !K || N < 2 || N_E < 1 || S < 1 ||
      pitch < N || row0 < 0)
    return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, S * N_E);
  ic_planck_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(gam, N, Eph, N_E, seed_T,
                                                              seed_theta, S, K, pitch, row0);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_ic_seed_table(const double* gam, int N, const double* Eph, int N_E, const double* eps0,
                     const double* phn, int Ns, double* K, int pitch, int row0, void* stream) {
  if (!gam || !Eph || !eps0 || !phn || !K || N < 2 || N_E < 1 || Ns < 1 || pitch < N ||
      row0 < 0)
    return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, N_E, 1);
  ic_seed_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(gam, N, Eph, N_E, eps0, phn, Ns, 0,
                                                            K, pitch, row0, 0);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_brems_table(const double* gam, int N, const double* eps, int N_E, double* K, int pitch,
                   int row0, void* stream) {
  if (!gam || !eps || !K || N < 2 || N_E < 1 || pitch < N || row0 < 0) return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, N_E);
  brems_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(gam, N, eps, N_E, K, pitch, row0);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_pp_analytic_table(int hiEmodel, int nuclear_enhancement, const double* Ep_GeV, int N,
                         const double* Eg_GeV, int N_E, double* K, int pitch, int row0,
                         void* stream) {
  if (!Ep_GeV || !Eg_GeV || !K || N < 2 || N_E < 1 || pitch < N || row0 < 0 || hiEmodel < 0 ||
      hiEmodel > NB_PP_QGSJET)
    return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, N_E);
  pp_analytic_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(
      hiEmodel, nuclear_enhancement, Ep_GeV, N, Eg_GeV, N_E, K, pitch, row0);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_pp_lut_table(const double* tx, int nx, const double* ty, int ny, const double* c,
                    const double* Ep_GeV, int N, const double* Eg_GeV, int N_E, double* K,
                    int pitch, int row0, void* stream) {
  if (!tx || !ty || !c || !Ep_GeV || !Eg_GeV || !K || nx < 8 || ny < 8 || N < 2 || N_E < 1 ||
      pitch < N || row0 < 0)
    return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, N_E);
  pp_lut_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(tx, nx, ty, ny, c, Ep_GeV, N, Eg_GeV,
                                                           N_E, K, pitch, row0);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_table_finalize(const double* K, int R, int N, int pitch, const double* invdlx,
                      double* lrs, void* stream) {
  if (!K || !invdlx || !lrs || R < 1 || N < 2 || pitch < N) return NB_EINVAL;
  dim3 grid((pitch + 127) / 128, R);
  table_finalize_kernel<<<grid, 128, 0, as_stream(stream)>>>(K, N, pitch, invdlx, lrs);
  NB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"

template <int RT, int MODE>
static int launch_contract(const ContractArgs& a, int smem, cudaStream_t st) {
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(contract_kernel<RT, MODE>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid((a.R + RT - 1) / RT, (a.W + a.w_per_cta - 1) / a.w_per_cta);
  NB_LAUNCH(contract_kernel<RT, MODE>, grid, dim3(256), (size_t)smem, st, a);
  return 0;
}

template <int MODE>
static int dispatch_contract(int RT, const ContractArgs& a, int smem, cudaStream_t st) {
  if (RT == 8) return launch_contract<8, MODE>(a, smem, st);
  if (RT == 4) return launch_contract<4, MODE>(a, smem, st);
  return launch_contract<2, MODE>(a, smem, st);
}

extern "C" {

int nb_table_scan(const double* K, int R, int N, int pitch, int* row_j0, int* flags,
                  void* stream) {
  if (!K || !row_j0 || !flags || R < 1 || N < 2 || pitch < N) return NB_EINVAL;
  table_scan_kernel<<<(R * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(K, R, N, pitch, row_j0,
                                                                         flags);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_contract_ex(const double* K, const double* lrs, int R, int N, int pitch,
                   const int* row_j0, const double* xn, const double* ds1, int wpitch, int W,
                   const double* dlx, const double* xgrid, const double* coef, double* out,
                   int mode, void* stream) {
  if (!K || !xn || !out || R < 1 || N < 2 || pitch < N || wpitch < N || W < 0 || mode < 0 ||
      mode > 2)
    return NB_EINVAL;
  const bool exact = mode == 1;
  if (!exact && (!lrs || !ds1 || !dlx)) return NB_EINVAL;
  if (exact && !xgrid) return NB_EINVAL;
  if ((pitch & 1) || ((uintptr_t)K & 15) || (lrs && ((uintptr_t)lrs & 15))) return NB_EALIGN;
  if (W == 0) return 0;
  ContractArgs a;
  a.K = K; a.lrs = lrs; a.R = R; a.N = N; a.pitch = pitch;
  a.xn = xn; a.ds1 = ds1; a.wpitch = wpitch; a.W = W;
  a.dlx = dlx; a.xgrid = xgrid; a.coef = coef; a.out = out; a.row_j0 = row_j0;
  // rows per tile: largest of 8/4/2 whose K+lrs tile stays <= ~96 KB (2 CTAs/SM)
  int narr = exact ? 1 : 2;
  long long row_bytes = (long long)narr * pitch * 8;
  int RT = 8;
  while (RT > 2 && RT * row_bytes > 96 * 1024) RT >>= 1;
  if (RT * row_bytes + 16 > 227 * 1024) return NB_ETOOLARGE;
  int smem = (int)(RT * row_bytes + 16);
  // walkers per CTA (multiple of the 8 warps): as few as keeps the whole grid resident in
  // one wave (148 SMs x CTAs that fit by shared memory, at most 3 by registers) -- a second
  // partial wave costs more than longer CTAs
  int resident = (int)((227LL * 1024) / (smem + 1024));
  if (resident > (exact ? 2 : 3)) resident = exact ? 2 : 3;  // registers (launch bounds)
  if (resident < 1) resident = 1;
  int row_tiles = (R + RT - 1) / RT;
  int wpc = 8;
  while (wpc < 64 && (long long)row_tiles * ((W + wpc - 1) / wpc) > 148LL * resident) wpc <<= 1;
  a.w_per_cta = wpc;
  cudaStream_t st = as_stream(stream);
  if (mode == 1) return dispatch_contract<1>(RT, a, smem, st);
  if (mode == 2) return dispatch_contract<2>(RT, a, smem, st);
  return dispatch_contract<0>(RT, a, smem, st);
}

int nb_contract(const double* K, const double* lrs, int R, int N, int pitch,
                long long K_wstride, const double* xn, const double* ds1, int wpitch, int W,
                const double* dlx, const double* xgrid, const double* coef, double* out,
                int exact, void* stream) {
  if (K_wstride != 0) return NB_EINVAL;  // per-walker seed fields: nb_ssc_*
  return nb_contract_ex(K, lrs, R, N, pitch, nullptr, xn, ds1, wpitch, W, dlx, xgrid, coef, out,
                        exact ? 1 : 0, stream);
}

static int syn_geometry(SynArgs& a, int N, int W, int N_E, long long* smem) {
  // photon energies per CTA: every CTA repeats the per-walker node set-up, so take as
  // many as still leaves >= 2 CTAs per SM (148 SMs), but at least one per warp
  int epc = (N_E + 7) & ~7;
  // about 1.5 CTAs per SM: the kernel runs beside the contraction, whose CTAs need registers
  // left over on the SMs (four synchrotron CTAs would take the whole register file: measured,
  // the contraction then starts only when they retire)
  while (epc > 8 && (long long)W * ((N_E + epc - 1) / epc) < 200) epc = ((epc / 2) + 7) & ~7;
  a.e_per_cta = epc;
  *smem = 6LL * N * 8 + 4LL * (epc + 1) + 16LL * epc + 8;
  return (*smem > 224 * 1024) ? NB_ETOOLARGE : 0;
}

int nb_synchrotron(const double* gam, int N, const double* gm2, const double* g23,
                   const double* xn, const double* ds1, int wpitch, const double* invdlx,
                   const double* dlx, const double* B, int W, const double* E_erg,
                   const double* cbrtE, int N_E, double* out, int out_ld, void* stream) {
  if (!gam || !xn || !ds1 || !invdlx || !dlx || !B || !E_erg || !out || N < 2 || wpitch < N ||
      W < 0 || N_E < 1 || (gm2 == nullptr) != (g23 == nullptr))
    return NB_EINVAL;
  if (out_ld == 0) out_ld = N_E;
  if (out_ld < N_E) return NB_EINVAL;
  if (W == 0) return 0;
  SynArgs a;
  a.gam = gam; a.N = N; a.gm2 = gm2; a.g23 = g23; a.xn = xn; a.ds1 = ds1; a.wpitch = wpitch;
  a.invdlx = invdlx; a.dlx = dlx; a.B = B; a.W = W; a.E_erg = E_erg; a.cbrtE = cbrtE;
  a.N_E = N_E; a.out = out; a.out_ld = out_ld;
  long long smem;
  int rc = syn_geometry(a, N, W, N_E, &smem);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(synchrotron_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(W, (N_E + a.e_per_cta - 1) / a.e_per_cta);
  NB_LAUNCH(synchrotron_kernel, grid, dim3(256), (size_t)smem, as_stream(stream), a);
  return 0;
}

static int fill_walker_src(WalkerSrc& s, const nb_walker_src* src, int W) {
  if (!src || src->P < 1 || src->P > 32 || src->n_map < 0 || src->n_map > 32 ||
      (src->n_map > 0 && !src->map_host) || (!src->mv_host && !src->pars))
    return NB_EINVAL;
  for (int k = 0; k < src->n_map; ++k) {
    s.pm.map[k] = src->map_host[k];
    if (s.pm.map[k].src >= src->P || s.pm.map[k].fn < 0 || s.pm.map[k].fn > NB_FN_EXP)
      return NB_EINVAL;
  }
  s.pm.n_out = src->n_map;
  s.pm.n_pri = 0;
  s.pm.W = W;
  s.pm.P = src->P;
  s.pm.pars = src->pars;
  s.pm.out = nullptr;
  s.pm.prior_out = nullptr;
  s.has_mv = src->mv_host ? 1 : 0;
  if (src->mv_host) {
    const nb_stretch* mv = src->mv_host;
    if (!mv->coords || !mv->step || !mv->s_idx || !mv->c_idx || !mv->zz || mv->P != src->P ||
        mv->i0 < 0 || mv->i0 + W > mv->Ns || mv->split < 0 || mv->split > 1)
      return NB_EINVAL;
    s.mv = *mv;
  }
  return 0;
}

static int fill_pd_desc(PdDesc& d, const nb_pd_desc* pd) {
  if (!pd || pd->kind < 0 || pd->kind > NB_PD_LOGPAR || pd->pd_off < 0 || !pd->lnx || !pd->invdlx)
    return NB_EINVAL;
  d.kind = pd->kind;
  d.pd_off = pd->pd_off;
  d.e_mul1 = pd->e_mul1;
  d.e_mul2 = pd->e_mul2;
  d.n_scale = pd->n_scale;
  d.lnx = pd->lnx;
  d.invdlx = pd->invdlx;
  return 0;
}

int nb_synchrotron_fused(const nb_walker_src* src, const nb_pd_desc* pd, int b_entry,
                         const double* gam, int N, const double* gm2, const double* g23,
                         const double* dlx, int W, const double* E_erg, const double* cbrtE,
                         int N_E, double* out, int out_ld, void* stream) {
  if (!gam || !dlx || !E_erg || !out || N < 2 || W < 0 || N_E < 1 || b_entry < 0 ||
      (gm2 == nullptr) != (g23 == nullptr))
    return NB_EINVAL;
  if (out_ld == 0) out_ld = N_E;
  if (out_ld < N_E) return NB_EINVAL;
  SynFusedArgs fa;
  int rc = fill_walker_src(fa.src, src, W);
  if (rc) return rc;
  if (b_entry >= fa.src.pm.n_out) return NB_EINVAL;
  rc = fill_pd_desc(fa.pd, pd);
  if (rc) return rc;
  if (W == 0) return 0;
  fa.b_entry = b_entry;
  SynArgs& a = fa.a;
  a.gam = gam; a.N = N; a.gm2 = gm2; a.g23 = g23; a.xn = nullptr; a.ds1 = nullptr; a.wpitch = 0;
  a.invdlx = pd->invdlx; a.dlx = dlx; a.B = nullptr; a.W = W; a.E_erg = E_erg;
  a.cbrtE = cbrtE; a.N_E = N_E; a.out = out; a.out_ld = out_ld;
  long long smem;
  rc = syn_geometry(a, N, W, N_E, &smem);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(synchrotron_fused_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(W, (N_E + a.e_per_cta - 1) / a.e_per_cta);
  NB_LAUNCH(synchrotron_fused_kernel, grid, dim3(256), (size_t)smem, as_stream(stream), fa);
  return 0;
}

static int launch_combine(const nb_peers* peers, const nb_stretch* mv, const double* pars,
                          const nb_term* terms_host,
                          int n_terms, int W, int N_E, const double* unit_fac,
                          const double* data_flux, const double* err_lo, const double* err_hi,
                          const int* ul, const double* cl, const double* prior,
                          double* flux_model, int flux_ld, double* lnp, int lnp_ld,
                          void* stream) {
  if (!terms_host || n_terms < 1 || n_terms > NB_MAX_TERMS || W < 0 || N_E < 1 || !unit_fac)
    return NB_EINVAL;
  if (flux_ld == 0) flux_ld = N_E;
  if (flux_ld < N_E) return NB_EINVAL;
  if (lnp && (!data_flux || !err_lo || !err_hi || !ul || !cl)) return NB_EINVAL;
  if (!lnp && !flux_model) return NB_EINVAL;
  CombineKernelArgs ka;
  CombineArgs& a = ka.c;
  for (int t = 0; t < n_terms; ++t) a.terms[t] = terms_host[t];
  if (!a.terms[n_terms - 1].group_end) return NB_EINVAL;
  a.n_terms = n_terms; a.W = W; a.N_E = N_E; a.unit_fac = unit_fac;
  a.data_flux = data_flux; a.err_lo = err_lo; a.err_hi = err_hi; a.ul = ul; a.cl = cl;
  a.prior = prior; a.flux_model = flux_model; a.flux_ld = flux_ld; a.lnp = lnp;
  a.lnp_ld = lnp_ld > 0 ? lnp_ld : 1;
  ka.has_mv = mv ? 1 : 0;
  ka.pars = pars;
  ka.has_peers = peers ? 1 : 0;
  if (peers) {
    if (!mv || peers->world < 1 || peers->world > NB_MAX_PEERS || peers->rank < 0 ||
        peers->rank >= peers->world || !peers->gen || !peers->ticket || !flux_model || !lnp)
      return NB_EINVAL;
    if (!peers->arena_local[0] || peers->arena_bytes[0] == 0) return NB_EINVAL;
    for (int k = 0; k < 2; ++k)
      if (peers->arena_local[k] && !peers->arena_mc[k])
        for (int p = 0; p < peers->world; ++p)
          if (!peers->arena_peer[k][p]) return NB_EINVAL;
    if (!peers->mc_flags)
      for (int p = 0; p < peers->world; ++p)
        if (!peers->flags[p]) return NB_EINVAL;
    ka.peers = *peers;
  }
  if (mv) {
    const bool sharded = peers != nullptr;  // replicated state: this rank's slice only
    if (!lnp || !pars || !mv->coords || !mv->lp || !mv->step || !mv->sync || !mv->s_idx ||
        !mv->zz || !mv->lnu || !mv->n_accepted || (!sharded && (mv->Ns != W || mv->i0 != 0)) ||
        (sharded && (mv->i0 < 0 || mv->i0 + W > mv->Ns)) || mv->P < 1 ||
        mv->W < W || (mv->pars_ld != 0 && mv->pars_ld < mv->P) ||
        mv->split < 0 || mv->split > 1 || mv->nb < 0 ||
        (mv->nb > 0 && (!mv->blobs || !flux_model || mv->nb < N_E || mv->nb > flux_ld)))
      return NB_EINVAL;
    ka.mv = *mv;
    if (ka.mv.pars_ld == 0) ka.mv.pars_ld = mv->P;
  }
  if (W == 0) return 0;
  size_t smem = (size_t)COMBINE_WARPS * N_E * sizeof(double);
  if (smem > 48 * 1024) return NB_ETOOLARGE;
  NB_LAUNCH(combine_lnprob_kernel, dim3((W + COMBINE_WARPS - 1) / COMBINE_WARPS),
            dim3(COMBINE_WARPS * 32), smem, as_stream(stream), ka);
  return 0;
}

int nb_combine_lnprob(const nb_term* terms_host, int n_terms, int W, int N_E,
                      const double* unit_fac, const double* data_flux, const double* err_lo,
                      const double* err_hi, const int* ul, const double* cl, const double* prior,
                      double* flux_model, int flux_ld, double* lnp, void* stream) {
  return launch_combine(nullptr, nullptr, nullptr, terms_host, n_terms, W, N_E, unit_fac,
                        data_flux, err_lo, err_hi, ul, cl, prior, flux_model, flux_ld, lnp, 1,
                        stream);
}

int nb_combine_lnprob_ld(const nb_term* terms_host, int n_terms, int W, int N_E,
                         const double* unit_fac, const double* data_flux, const double* err_lo,
                         const double* err_hi, const int* ul, const double* cl,
                         const double* prior, double* flux_model, int flux_ld, double* lnp,
                         int lnp_ld, void* stream) {
  if (lnp_ld < 1) return NB_EINVAL;
  return launch_combine(nullptr, nullptr, nullptr, terms_host, n_terms, W, N_E, unit_fac,
                        data_flux, err_lo, err_hi, ul, cl, prior, flux_model, flux_ld, lnp,
                        lnp_ld, stream);
}

int nb_combine_lnprob_update(const nb_stretch* mv_host, const double* pars,
                             const nb_term* terms_host, int n_terms, int W, int N_E,
                             const double* unit_fac, const double* data_flux,
                             const double* err_lo, const double* err_hi, const int* ul,
                             const double* cl, const double* prior, double* flux_model,
                             int flux_ld, double* lnp, void* stream) {
  if (!mv_host) return NB_EINVAL;
  return launch_combine(nullptr, mv_host, pars, terms_host, n_terms, W, N_E, unit_fac, data_flux,
                        err_lo, err_hi, ul, cl, prior, flux_model, flux_ld, lnp, 1, stream);
}

int nb_combine_lnprob_update_push(const nb_stretch* mv_host, const nb_peers* peers_host,
                                  const double* pars, const nb_term* terms_host, int n_terms,
                                  int W, int N_E, const double* unit_fac,
                                  const double* data_flux, const double* err_lo,
                                  const double* err_hi, const int* ul, const double* cl,
                                  const double* prior, double* flux_model, int flux_ld,
                                  double* lnp, void* stream) {
  if (!mv_host || !peers_host) return NB_EINVAL;
  return launch_combine(peers_host, mv_host, pars, terms_host, n_terms, W, N_E, unit_fac,
                        data_flux, err_lo, err_hi, ul, cl, prior, flux_model, flux_ld, lnp, 1,
                        stream);
}

static int fill_param_map(ParamMapArgs& a, const double* pars, int W, int P,
                          const nb_parmap* map_host, int n_out, double* out,
                          const nb_prior* priors_host, int n_priors, double* prior_out) {
  if (!pars || W < 0 || P < 1 || n_out < 0 || n_out > NB_MAX_MAP || n_priors < 0 ||
      n_priors > NB_MAX_PRIORS || (n_out > 0 && (!map_host || !out)) ||
      (n_priors > 0 && !priors_host))
    return NB_EINVAL;
  for (int k = 0; k < n_out; ++k) {
    a.map[k] = map_host[k];
    if (a.map[k].src >= P || a.map[k].fn < 0 || a.map[k].fn > NB_FN_EXP ||
        a.map[k].dst_off < 0 || a.map[k].dst_stride < 0)
      return NB_EINVAL;
  }
  for (int k = 0; k < n_priors; ++k) {
    a.pri[k] = priors_host[k];
    if (a.pri[k].par < 0 || a.pri[k].par >= P || a.pri[k].kind < 0 ||
        a.pri[k].kind > NB_PRIOR_LOGUNIFORM)
      return NB_EINVAL;
  }
  a.n_out = n_out; a.n_pri = n_priors; a.W = W; a.P = P;
  a.pars = pars; a.out = out; a.prior_out = prior_out;
  return 0;
}

static int launch_walker_prep(const nb_stretch* mv, double* pars_out, const double* pars, int W,
                              int P, const nb_parmap* map_host, int n_map, double* pm,
                              const nb_prior* priors_host, int n_priors, double* prior_out,
                              const nb_prep_job* jobs_host, int n_jobs, void* stream) {
  WalkerPrepArgs a;
  int rc = fill_param_map(a.pm, pars, W, P, map_host, n_map, pm, priors_host, n_priors,
                          prior_out);
  if (rc) return rc;
  a.has_mv = mv ? 1 : 0;
  a.pars_out = pars_out;
  if (mv) {
    if (!mv->coords || !mv->step || !mv->s_idx || !mv->c_idx || !mv->zz || mv->P != P ||
        mv->i0 < 0 || mv->i0 + W > mv->Ns || mv->split < 0 || mv->split > 1 || !pars_out ||
        (mv->pars_ld != 0 && mv->pars_ld < P))
      return NB_EINVAL;
    if (P > NB_MAX_MOVE_PAR) return NB_ETOOLARGE;
    a.mv = *mv;
    if (a.mv.pars_ld == 0) a.mv.pars_ld = P;
  }
  if (n_jobs < 0 || n_jobs > NB_MAX_PREP_JOBS || (n_jobs > 0 && (!jobs_host || !pm)))
    return NB_EINVAL;
  for (int k = 0; k < n_jobs; ++k) {
    const nb_prep_job& J = jobs_host[k];
    if (J.kind < 0 || J.kind > NB_PD_LOGPAR || J.N < 2 || !J.x || J.pd_off < 0 ||
        (!J.xn && !J.energy_out) || (J.xn && (!J.ds1 || !J.invdlx || J.wpitch < J.N)))
      return NB_EINVAL;
    a.jobs[k] = J;
  }
  a.n_jobs = n_jobs;
  a.n_items = 0;
  int max_energy_N = 0;
  for (int k = 0; k < n_jobs; ++k) {
    const nb_prep_job& J = a.jobs[k];
    if (J.xn)
      for (int j0 = 0; j0 < J.N; j0 += PREP_CHUNK) {
        if (a.n_items == NB_MAX_PREP_ITEMS) return NB_ETOOLARGE;
        a.items[a.n_items++] = PrepItem{(short)k, 0, j0};
      }
    if (J.energy_out) {
      if (a.n_items == NB_MAX_PREP_ITEMS) return NB_ETOOLARGE;
      a.items[a.n_items++] = PrepItem{(short)k, 1, 0};
      if (J.N > max_energy_N) max_energy_N = J.N;
    }
  }
  if (W == 0) return 0;
  size_t smem = (size_t)max_energy_N * sizeof(double);
  if (smem > 200 * 1024) return NB_ETOOLARGE;
  if (smem > 32 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(walker_prep_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  dim3 grid(W, a.n_items > 0 ? a.n_items : 1);
  NB_LAUNCH(walker_prep_kernel, grid, dim3(256), smem, as_stream(stream), a);
  return 0;
}

int nb_walker_prep(const double* pars, int W, int P, const nb_parmap* map_host, int n_map,
                   double* pm, const nb_prior* priors_host, int n_priors, double* prior_out,
                   const nb_prep_job* jobs_host, int n_jobs, void* stream) {
  return launch_walker_prep(nullptr, nullptr, pars, W, P, map_host, n_map, pm, priors_host,
                            n_priors, prior_out, jobs_host, n_jobs, stream);
}

int nb_walker_prep_move(const nb_stretch* mv_host, double* pars, int W, int P,
                        const nb_parmap* map_host, int n_map, double* pm,
                        const nb_prior* priors_host, int n_priors, double* prior_out,
                        const nb_prep_job* jobs_host, int n_jobs, void* stream) {
  if (!mv_host) return NB_EINVAL;
  return launch_walker_prep(mv_host, pars, pars, W, P, map_host, n_map, pm, priors_host,
                            n_priors, prior_out, jobs_host, n_jobs, stream);
}

int nb_ic_seed_spectrum(const double* gam, int N, const double* nraw, int wpitch,
                        const double* Eph, int N_E, const double* eps0, const double* phn,
                        int Ns, int phn_wstride, int W, double* out, int out_ld, int out_off,
                        void* stream) {
  if (!gam || !nraw || !Eph || !eps0 || !phn || !out || N < 2 || wpitch < N || N_E < 1 ||
      Ns < 1 || W < 0 || out_off < 0 || out_ld < out_off + N_E ||
      (phn_wstride != 0 && phn_wstride < Ns))
    return NB_EINVAL;
  if (W == 0) return 0;
  if (W > 65535) return NB_ETOOLARGE;
  long long smem = (2LL * Ns + N + 8) * 8;
  if (smem > 227 * 1024) return NB_ETOOLARGE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ic_seed_spectrum_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  dim3 grid(N_E, W);
  ic_seed_spectrum_kernel<<<grid, 256, (int)smem, as_stream(stream)>>>(
      gam, N, nraw, wpitch, Eph, eps0, phn, Ns, phn_wstride, out, out_ld, out_off);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_ssc_table(const double* gam, int N, const double* Eph, int N_E, const double* eps0,
                 const double* invdlx_s, int Ns, double* KL, double* F0, double* coef,
                 long long Rp, void* stream) {
  if (!gam || !Eph || !eps0 || !invdlx_s || !KL || !F0 || !coef || N < 2 || N_E < 1 || Ns < 2 ||
      Rp < (long long)N * N_E || (Rp & 127) || ((uintptr_t)KL & 15))
    return NB_EINVAL;
  ssc_table_kernel<<<(unsigned)((Rp + 127) / 128), 128, 0, as_stream(stream)>>>(
      gam, N, Eph, N_E, eps0, invdlx_s, Ns, reinterpret_cast<double2*>(KL), F0, coef, Rp);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_ssc_seed(const nb_ssc_src* src_host, int n_src, int W, int Ns, const double* invdlx_s,
                double* sxn, double* sds, int spitch, void* stream) {
  if (!src_host || n_src < 1 || n_src > NB_SSC_MAX_SRC || W < 0 || Ns < 2 || !invdlx_s || !sxn ||
      !sds || spitch < Ns)
    return NB_EINVAL;
  SscSeedArgs a;
  for (int k = 0; k < n_src; ++k) {
    if (!src_host[k].src || src_host[k].off < 0 || src_host[k].ld < src_host[k].off + Ns)
      return NB_EINVAL;
    a.src[k] = src_host[k].src;
    a.ld[k] = src_host[k].ld;
    a.off[k] = src_host[k].off;
    a.fac[k] = src_host[k].fac;
  }
  a.n_src = n_src; a.W = W; a.Ns = Ns; a.spitch = spitch; a.invdlx_s = invdlx_s;
  a.sxn = sxn; a.sds = sds;
  if (W == 0) return 0;
  if (W > 65535) return NB_ETOOLARGE;
  dim3 grid((spitch + 127) / 128, W);
  ssc_seed_kernel<<<grid, 128, 0, as_stream(stream)>>>(a);
  NB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"

template <int WT>
static int launch_ssc_inner(const SscInnerArgs& a, cudaStream_t st) {
  size_t smem = (size_t)WT * a.Ns * sizeof(double2) + SSC_STAGES * 128 * sizeof(double2) +
                WT * sizeof(double);
  if (smem > 200 * 1024) return NB_ETOOLARGE;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(ssc_inner_kernel<WT>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  // walker groups on the fast grid axis: the CTAs that share a row tile run together, so
  // the table streams from HBM once and is re-read from L2
  dim3 grid((a.W + WT - 1) / WT, (unsigned)(a.Rp / 128));
  ssc_inner_kernel<WT><<<grid, 128, smem, st>>>(a);
  NB_CHECK_LAUNCH();
  return 0;
}

extern "C" {

int nb_ssc_inner(const double* KL, const double* F0, const double* coef, long long Rp, int Ns,
                 const double* sxn, const double* sds, int spitch, int W, const double* dlx_s,
                 double* inner, void* stream) {
  if (!KL || !F0 || !coef || !sxn || !sds || !dlx_s || !inner || Rp < 128 || (Rp & 127) ||
      Ns < 2 || spitch < Ns || W < 0 || ((uintptr_t)KL & 15))
    return NB_EINVAL;
  if (W == 0) return 0;
  if (Rp / 128 > 65535) return NB_ETOOLARGE;
  SscInnerArgs a;
  a.KL = reinterpret_cast<const double2*>(KL); a.F0 = F0; a.coef = coef; a.Rp = Rp; a.Ns = Ns;
  a.sxn = sxn; a.sds = sds; a.spitch = spitch; a.W = W; a.dlx_s = dlx_s; a.inner = inner;
  // eight walkers per thread, 6 CTAs = 24 warps per SM.  Sixteen walkers per thread (4 CTAs per
  // SM, half the table traffic) measured the same: 820.9 against 809.7 us per launch at C4,
  // fp64 pipe 57 % against 58 % (profiles/r02_ncu_c4_ssc_inner_wt16.md) -- the kernel is bound
  // by the fp64 dependency chains of the lean cell, not by the table stream.
  return launch_ssc_inner<8>(a, as_stream(stream));
}

int nb_ssc_outer(const double* inner, long long Rp, int N, int N_E, int W, const double* xn,
                 const double* ds1, int wpitch, const double* dlx, const double* invdlx,
                 const double* coef_e, double* out, int out_ld, int out_off, void* stream) {
  if (!inner || !xn || !ds1 || !dlx || !invdlx || !coef_e || !out || N < 2 || N_E < 1 || W < 0 ||
      Rp < (long long)N * N_E || wpitch < N || out_off < 0 || out_ld < out_off + N_E)
    return NB_EINVAL;
  if (W == 0) return 0;
  SscOuterArgs a;
  a.inner = inner; a.Rp = Rp; a.N = N; a.N_E = N_E; a.W = W; a.xn = xn; a.ds1 = ds1;
  a.wpitch = wpitch; a.dlx = dlx; a.invdlx = invdlx; a.coef_e = coef_e; a.out = out;
  a.out_ld = out_ld; a.out_off = out_off;
  long long warps = (long long)W * N_E;
  ssc_outer_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, as_stream(stream)>>>(a);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_stretch_update_packed(const nb_stretch* mv, const double* pack, int ld, void* stream) {
  if (!mv || !pack || !mv->coords || !mv->lp || !mv->step || !mv->sync || !mv->s_idx ||
      !mv->zz || !mv->lnu || !mv->n_accepted || mv->P < 1 || mv->Ns < 0 || mv->W < mv->Ns ||
      mv->split < 0 || mv->split > 1 || mv->nb < 0 || (mv->nb > 0 && !mv->blobs) ||
      ld < mv->nb + 1 + mv->P)
    return NB_EINVAL;
  if (mv->Ns == 0) return 0;
  stretch_update_packed_kernel<<<(mv->Ns + 3) / 4, 128, 0, as_stream(stream)>>>(*mv, pack, ld);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_peer_wait(const nb_stretch* mv, void* stream) {
  if (!mv || !mv->wait_flags || !mv->wait_gen || mv->wait_world < 1 ||
      mv->wait_world > NB_MAX_PEERS)
    return NB_EINVAL;
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(*mv);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_kelner_table(const double* Eg_TeV, const int* hi, int R, int N, double decades,
                    double* Ep, double* Kk, void* stream) {
  if (!Eg_TeV || !hi || !Ep || !Kk || R < 1 || N < 2 || !(decades > 0.0)) return NB_EINVAL;
  dim3 grid((N + 127) / 128, R);
  kelner_table_kernel<<<grid, 128, 0, as_stream(stream)>>>(Eg_TeV, hi, R, N, decades, Ep, Kk);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_kelner_rows(int kind, const double* pd_params, int W, const double* Ep, const double* Kk,
                   int R, int N, double* out, void* stream) {
  if (!pd_params || !Ep || !Kk || !out || W < 0 || R < 1 || N < 2 || kind < 0 ||
      kind > NB_PD_LOGPAR)
    return NB_EINVAL;
  if (W == 0) return 0;
  if (W > 65535) return NB_ETOOLARGE;
  size_t smem = (size_t)N * sizeof(double);
  if (smem > 48 * 1024) return NB_ETOOLARGE;
  dim3 grid(R, W);
  kelner_rows_kernel<<<grid, 256, smem, as_stream(stream)>>>(kind, pd_params, Ep, Kk, R, N, out);
  NB_CHECK_LAUNCH();
  return 0;
}

int nb_timeline_reset(void) {
  // the contraction kernel stamps through g_timeline_row: forget a row of a timeline buffer
  // that is about to be freed
  unsigned long long* null_row = nullptr;
  cudaError_t e = cudaMemcpyToSymbol(g_timeline_row, &null_row, sizeof(null_row));
  return e == cudaSuccess ? 0 : (int)e;
}

int nb_launch_carveout(int percent) {
  if (percent < -1 || percent > 100) return NB_EINVAL;
  g_launch_carveout = percent;
  return 0;
}

int nb_fallback_counts(unsigned long long* out_host, int reset) {
  if (!out_host) return NB_EINVAL;
  cudaError_t e = cudaMemcpyFromSymbol(out_host, g_fallbacks, sizeof(g_fallbacks));
  if (e != cudaSuccess) return (int)e;
  if (reset) {
    const unsigned long long z[2] = {0ull, 0ull};
    e = cudaMemcpyToSymbol(g_fallbacks, z, sizeof(z));
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

int nb_fp64_peak_probe(double* out, int blocks, int threads, int iters, void* stream) {
  if (!out || blocks < 1 || threads < 1 || threads > 1024 || iters < 1) return NB_EINVAL;
  fp64_probe_kernel<<<blocks, threads, 0, as_stream(stream)>>>(out, iters);
  NB_CHECK_LAUNCH();
  return 0;
}


int nb_trapz_loglog(const double* y, int R, int N, int ld, const double* x, int x_ld,
                    double* out, double* intervals, void* stream) {
  if (!y || !x || (!out && !intervals) || R < 0 || N < 1 || ld < N || (x_ld != 0 && x_ld < N))
    return NB_EINVAL;
  if (R == 0) return 0;
  trapz_loglog_kernel<<<(R * 32 + 255) / 256, 256, 0, as_stream(stream)>>>(y, R, N, ld, x, x_ld,
                                                                           out, intervals);
  NB_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"
