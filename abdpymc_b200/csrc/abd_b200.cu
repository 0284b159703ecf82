// libabd_b200.so -- kernels (sm_100a) and the C ABI declared in include/abd_b200.h.
//
// HBM layout (all built once in abd_create):
//   pcr[N], vac[N]              one bit mask over gaps per individual (uint32 if G <= 31)
//   per antigen a in {N, S}:    CSR by individual -- rp_a[N+1] (int32), and row arrays sorted by
//                               (individual, gap): od_a[R_a] f64, x_a[R_a] f64,
//                               meta_a[R_a] u32 = individual << 6 | gap
//   per chain (resident state): i_raw[C][G][N] int8, waner[C][N] int8  (the reference layout)
//
// Kernels
//   k_sums      one CTA per (tile of individuals, chain): int8 columns -> bit masks ->
//               constrained infections in shared memory; one thread per OD row evaluates the
//               titer in closed form from the masks (power tables in shared memory), the
//               logistic curve and the 13 gradient accumulators; warp-shuffle + block
//               reduction; the last CTA of each chain (atomic ticket) reduces the tiles in a
//               fixed order and applies priors / transforms (single launch, deterministic).
//   k_finalize  the same finalisation as a separate launch (after an all-reduce when
//               individuals are sharded across GPUs).
//   k_gibbs     one warp per (individual, chain), lanes over that individual's OD rows; G+1
//               sequential single-bit updates, both states evaluated, Philox4x32-10.
//   k_determ    one thread per (individual, chain): i, ab_n_mu, ab_s_mu for a recorded draw.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/abd_b200.h"
#include "abd_device.cuh"

using namespace abd;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess)                                                                    \
      return fail(ABD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
  } while (0)

// ------------------------------------------------------------------------------------------
// device-side view of the cohort
// ------------------------------------------------------------------------------------------
struct DevCohort {
  int G, N;
  unsigned ind_offset;       // global index of this shard's first individual (RNG streams)
  const void* pcr;
  const void* vac;
  const int* rp[2];          // [0] = N antigen, [1] = S antigen
  const double* od[2];
  const double* x[2];
  // factored mode (the cohort has <= 32 distinct log_dilution values, the usual case):
  const uint32_t* rcx[2];      // per row: cell index << 5 | index of its dilution in xlev
  const double* xlev;          // [32] the distinct dilutions, ascending (padded with 0)
  int n_xlev;                  // 0: not available (k_sums stages the dilutions as doubles)
  const uint32_t* meta[2];     // per row: individual << 6 | gap
  const uint32_t* rowcell[2];  // per row: index of its (individual, gap) cell
  const uint32_t* cmeta[2];    // per cell: individual << 6 | gap
  Chunks ch;
};

constexpr int kSumsBlock = 256;
constexpr int kSumsWarps = kSumsBlock / 32;
constexpr int kAuxDoubles = 128;  // 17 x PriorPre (7) + LikPre (6), padded
constexpr int kTileMaxInds = 128;
constexpr int kGibbsWarps = 8;

// per-individual state staged in shared memory: constrained infections, vaccinations, waner
// (packed in the top bit of the vaccination mask; usable gaps <= width - 1)
template <typename M>
struct IndState {
  M inf, vacw;
};
template <typename M>
__device__ __forceinline__ M top_bit() { return (M)1 << (sizeof(M) * 8 - 1); }

__device__ __forceinline__ double load_param(const double* theta, int theta_is_q, int c, int k13) {
  if (theta_is_q) {
    const int j = kQOfTheta[k13];
    return backward(theta[(size_t)c * 17 + j], kQTransform[j]);
  }
  return theta[(size_t)c * 13 + k13];
}

// One warp: pw[k] = rho^k and (optionally) dpw[k] = k rho^(k-1) for k < G by a shuffle scan.
__device__ __forceinline__ void fill_pow_warp(double rho, int G, int lane, double* pw, double* dpw) {
  double v = rho;  // inclusive prefix product: rho^(lane+1)
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, v, off);
    if (lane >= off) v *= u;
  }
  const double r32 = __shfl_sync(0xffffffffu, v, 31);  // rho^32
  double prev = __shfl_up_sync(0xffffffffu, v, 1);      // rho^lane
  if (lane == 0) prev = 1.0;
  if (lane == 0) {
    pw[0] = 1.0;
    if (dpw) dpw[0] = 0.0;
  }
  for (int blk = 0; blk * 32 < G; ++blk) {
    const int k = blk * 32 + lane + 1;
    const double scale = blk ? r32 : 1.0;  // G <= 63: at most two blocks
    if (k < G) {
      pw[k] = v * scale;
      if (dpw) dpw[k] = (double)k * prev * scale;
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_sums
// ------------------------------------------------------------------------------------------
struct FinalizeCfg {
  int mode;        // 0: write raw sums only; 1: loglik + grad13; 2: joint logp + dlogp17
  Totals tot;
  double* out_val;   // [C]
  double* out_grad;  // [C][13] or [C][17] (may be null)
};

struct SumsCfg {
  int ntiles;
  int cap_n, cap_s;      // staged doubles per antigen (even)
  int capr_n, capr_s;    // staged row->cell words per antigen (multiple of 4)
  int capk_n, capk_s;    // staged cell-meta words per antigen (multiple of 4)
  int chains_per_cta;
  int C;
};

// Fused all-reduce over NVLink peer memory (individuals sharded over `world` GPUs of one node):
// the CTA that finishes a chain last on each rank stores its 16 raw sums straight into every
// peer's exchange buffer, raises a per-(rank, chain) flag there, waits for the flags the peers
// raise in ITS buffer, adds the `world` contributions in rank order (bitwise the same result on
// every rank) and finalises -- compute, exchange and finalisation in one launch, no NCCL call.
// Buffers are double-buffered on the parity of a per-chain sequence number kept on the device.
constexpr int kMaxPeers = 8;
constexpr unsigned long long kXchTimeoutNs = 10ull * 1000 * 1000 * 1000;  // a peer that is 10 s late is declared lost
struct XchCfg {
  int world, rank, cmax;                 // world == 0: not sharded
  double* data[kMaxPeers];               // peer r's buffer: [2][world][cmax][16] doubles ...
  unsigned long long* flag[kMaxPeers];   // ... followed by [2][world][cmax] flags
  unsigned* seq;                         // [cmax] local sequence numbers
  unsigned* err;
};

// Trajectory mode (n_steps > 0): the kernel stays resident for n_steps leapfrog steps of
// Hamiltonian dynamics over q17 with the binary state fixed.  Per step every CTA evaluates its
// tile at the current position; the CTA that finishes a chain last finalises logp / gradient,
// advances (q, p) and publishes them with a per-chain generation counter the other CTAs of that
// chain wait on -- no kernel launch, no re-staging of the cohort between evaluations.
struct TrajCfg {
  int n_steps;             // 0 = plain evaluation
  double* q;               // [C][17] in: start position, out: end position
  double* p;               // [C][17] in/out momentum
  double* grad;            // [C][17] in: gradient at q, out: gradient at the end position
  double* logp;            // [C]     out: logp at the end position
  const double* eps;       // [C] step sizes
  const double* inv_mass;  // [17][17] (symmetric) inverse mass matrix, shared by all chains
  double* state;           // [C][34] scratch: position and half-step momentum of the current step
  unsigned* gen;           // [C] generation counters (zeroed before the launch)
  unsigned* err;           // set to 1 if a wait timed out
};

// tile descriptor (48 bytes, three 16-byte loads): individuals [i0, i1); per antigen the OD rows
// [r0, r1) and the (individual, gap) cells [c0, c1) of those individuals
struct __align__(16) TileDesc {
  int i0, i1, rn0, rn1;
  int rs0, rs1, cn0, cn1;
  int cs0, cs1, pad0, pad1;
};

// trajectory of one cell: titer m, decaying part T (or U) and its rho-derivative; in factored mode
// also Em = exp(b m) (capped so that exp(-b x) * Em <= e^700)
template <bool FX>
struct CellValT {
  double m, T, dT;
};
template <>
struct __align__(16) CellValT<true> {
  double m, T, dT, Em;
};
constexpr int kMaxXLevels = 32;

#ifdef ABD_PHASE_TIMING
__device__ unsigned long long g_phase[4096][16];
__device__ unsigned long long g_span[256][2];  // per launch: first CTA start, last CTA end
__device__ unsigned g_span_idx, g_span_done;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t_;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
  return t_;
}
#define SPAN_BEGIN()                                                                      \
  unsigned span_slot_ = 0;                                                                \
  if (threadIdx.x == 0) {                                                                 \
    span_slot_ = *(volatile unsigned*)&g_span_idx & 255u;                                 \
    atomicMin(&g_span[span_slot_][0], gtime());                                           \
  }
#define SPAN_END()                                                                        \
  if (threadIdx.x == 0) {                                                                 \
    atomicMax(&g_span[span_slot_][1], gtime());                                           \
    if (atomicAdd(&g_span_done, 1u) == gridDim.x * gridDim.y - 1) {                       \
      g_span_done = 0;                                                                    \
      __threadfence();                                                                    \
      atomicAdd(&g_span_idx, 1u);                                                         \
    }                                                                                     \
  }
#define PHASE(i)                                                                         \
  do {                                                                                   \
    if (tid == 0 && blockIdx.y * gridDim.x + blockIdx.x < 4096) {                        \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      g_phase[blockIdx.y * gridDim.x + blockIdx.x][i] = t_;                              \
    }                                                                                    \
  } while (0)
#define PHASEW(i, cond)                                                                  \
  do {                                                                                   \
    if ((cond) && blockIdx.y * gridDim.x + blockIdx.x < 4096) {                          \
      unsigned long long t_;                                                             \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                             \
      g_phase[blockIdx.y * gridDim.x + blockIdx.x][i] = t_;                              \
    }                                                                                    \
  } while (0)
#else
#define PHASE(i)
#define PHASEW(i, cond)
#define SPAN_BEGIN()
#define SPAN_END()
#endif

#ifndef ABD_SUMS_MINB
#define ABD_SUMS_MINB 3
#endif
// XT: how the dilutions reach the row loop.  uint8_t = factored mode: a row carries the index of
// its dilution (packed with its cell index), exp(-b (x - m)) = exp(-b x) * exp(b m) costs one exp
// per CELL plus a per-chain table over the distinct dilutions; double = general fallback, the
// dilutions themselves are staged and every row evaluates its own exp.
template <typename M, typename XT, bool TRAJ>
__global__ void __launch_bounds__(kSumsBlock, ABD_SUMS_MINB)
k_sums(const DevCohort dc, const TileDesc* __restrict__ tiles, const SumsCfg cfg,
       const double* __restrict__ theta, const int theta_is_q,
       const int8_t* __restrict__ i_raw, const int8_t* __restrict__ waner,
       double* __restrict__ partial, unsigned* __restrict__ ticket, double* __restrict__ sums,
       const FinalizeCfg fin, const Priors* __restrict__ priors, double* __restrict__ aux, const TrajCfg traj,
       const XchCfg xch) {
  const int tile = blockIdx.x, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int G = dc.G, N = dc.N, ntiles = cfg.ntiles;

  // dynamic shared memory: staged rows (od, x, row->cell), staged cell meta, per-cell values
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  constexpr bool kFX = sizeof(XT) == 1;
  using CellVal = CellValT<kFX>;
  double* s_od_n = reinterpret_cast<double*>(dyn_smem);
  double* s_od_s = s_od_n + cfg.cap_n;
  CellVal* s_cv_n = reinterpret_cast<CellVal*>(s_od_s + cfg.cap_s);
  CellVal* s_cv_s = s_cv_n + cfg.capk_n;
  double* s_x_n = reinterpret_cast<double*>(s_cv_s + cfg.capk_s);   // cap_n doubles (not in factored mode)
  double* s_x_s = s_x_n + (kFX ? 0 : cfg.cap_n);
  uint32_t* s_rc_n = reinterpret_cast<uint32_t*>(s_x_s + (kFX ? 0 : cfg.cap_s));
  uint32_t* s_rc_s = s_rc_n + cfg.capr_n;
  uint32_t* s_cm_n = s_rc_s + cfg.capr_s;
  uint32_t* s_cm_s = s_cm_n + cfg.capk_n;

  __shared__ double s_th[16];
  __shared__ double s_pw[4][kMaxGaps];  // rho_n^k, d/drho; rho_s^k, d/drho
  __shared__ double s_tab[kExpTab];
  __shared__ IndState<M> s_ind[kTileMaxInds];
  __shared__ double s_red[kSumsWarps][kNSums];
  __shared__ double s_fin[kSumsBlock / 16][kNSums];
  __shared__ int s_last;
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ PriorPre s_pre[17];
  __shared__ LikPre s_lik;
  __shared__ double s_q[17], s_ph[17], s_out[18];  // trajectory mode: position, half-step momentum, logp + gradient
  __shared__ double s_im[TRAJ ? 17 * 17 : 1];       // trajectory mode: inverse mass matrix
  __shared__ double2 s_xe[kFX ? 2 : 1][kMaxXLevels];  // factored mode: {x_j, exp(-b x_j)} per antigen
  __shared__ double s_zmax[2];                      // cap on b m so that exp(-b x) exp(b m) <= e^700
  __shared__ int s_direct;                          // |b| x too large to factor: rows take their own exp

  PHASE(0);
  SPAN_BEGIN();
  griddep_launch_dependents();
  const int4 d0 = reinterpret_cast<const int4*>(tiles + tile)[0];
  const int4 d1 = reinterpret_cast<const int4*>(tiles + tile)[1];
  const int2 d2 = reinterpret_cast<const int2*>(tiles + tile)[4];
  const int i0 = d0.x, ni = d0.y - d0.x;
  const int rn0 = d0.z, rn1 = d0.w, rs0 = d1.x, rs1 = d1.y;
  const int cn0 = d1.z, cn1 = d1.w, cs0 = d2.x, cs1 = d2.y;

  // ---- stage this tile's rows and cell table in shared memory: one thread, bulk async copies ----
  const int an0 = rn0 & ~1, as0 = rs0 & ~1;      // 16-byte aligned starts (doubles)
  const int qn0 = rn0 & ~3, qs0 = rs0 & ~3;      // (32-bit words)
  const int kn0 = cn0 & ~3, ks0 = cs0 & ~3;
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    const uint32_t bn = (uint32_t)(((rn1 + 1) & ~1) - an0) * 8u, bs = (uint32_t)(((rs1 + 1) & ~1) - as0) * 8u;
    const uint32_t bqn = (uint32_t)(((rn1 + 3) & ~3) - qn0) * 4u, bqs = (uint32_t)(((rs1 + 3) & ~3) - qs0) * 4u;
    const uint32_t bkn = (uint32_t)(((cn1 + 3) & ~3) - kn0) * 4u, bks = (uint32_t)(((cs1 + 3) & ~3) - ks0) * 4u;
    mbar_expect_tx(&s_bar, (kFX ? bn + bs : 2 * bn + 2 * bs) + bqn + bqs + bkn + bks);
    if (bkn) bulk_g2s(s_cm_n, dc.cmeta[0] + kn0, bkn, &s_bar);
    if (bks) bulk_g2s(s_cm_s, dc.cmeta[1] + ks0, bks, &s_bar);
    if (bn) {
      bulk_g2s(s_od_n, dc.od[0] + an0, bn, &s_bar);
      if (!kFX) bulk_g2s(s_x_n, dc.x[0] + an0, bn, &s_bar);
    }
    if (bs) {
      bulk_g2s(s_od_s, dc.od[1] + as0, bs, &s_bar);
      if (!kFX) bulk_g2s(s_x_s, dc.x[1] + as0, bs, &s_bar);
    }
    if (bqn) bulk_g2s(s_rc_n, (kFX ? dc.rcx[0] : dc.rowcell[0]) + qn0, bqn, &s_bar);
    if (bqs) bulk_g2s(s_rc_s, (kFX ? dc.rcx[1] : dc.rowcell[1]) + qs0, bqs, &s_bar);
  }
  fill_exp_table(s_tab, tid, kSumsBlock);
  // the immutable PCR+ / vaccination masks of this thread's individual (read before the wait)
  M my_pcr = 0, my_vac = 0;
  if (tid < ni) {
    my_pcr = reinterpret_cast<const M*>(dc.pcr)[i0 + tid];
    my_vac = reinterpret_cast<const M*>(dc.vac)[i0 + tid];
  }
  double x_lev = 0.0;  // factored mode, last warp: lane j holds the j-th distinct dilution
  if (kFX && warp == kSumsWarps - 1) x_lev = dc.xlev[lane];
  __syncthreads();  // the exp table is usable from here on (still before the dependency wait)
  // everything above reads only the immutable cohort; parameters, chain state and the reduction
  // scratch may be written by the previous kernel in the stream
  griddep_wait();
  if (TRAJ) {
    for (int k = tid; k < 17 * 17; k += kSumsBlock) s_im[k] = traj.inv_mass[k];
    __syncthreads();
  }
  PHASE(1);

  for (int cc = 0; cc < cfg.chains_per_cta; ++cc) {
    const int c = blockIdx.y * cfg.chains_per_cta + cc;
    if (c >= cfg.C) break;

    const int nsteps = TRAJ ? traj.n_steps : 1;
    double cnt_i = 0.0, cnt_w = 0.0;  // this thread's individual: sum(i_raw), waner
    for (int step = 0; step < nsteps; ++step) {
    double acc[kNSums];
#pragma unroll
    for (int k = 0; k < kNSums; ++k) acc[k] = 0.0;
    acc[S_KI] = cnt_i;
    acc[S_KW] = cnt_w;

    // trajectory mode: this step's position.  Step 0: every warp that needs it advances the
    // start point itself (p_half = p + eps/2 g, q' = q + eps Sigma p_half: 17 FMAs per lane);
    // later steps: wait for the chain's previous finaliser to publish (q', p_half).
    double q_lane = 0.0, ph_lane = 0.0;  // lane k < 17: component k
    if (TRAJ) {
      if (step > 0) {
        if (tid == 0) {
          unsigned seen, polls = 0;
          do {
            asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(seen) : "l"(traj.gen + c) : "memory");
          } while (seen < (unsigned)step && ++polls < (1u << 26));
          if (seen < (unsigned)step) *traj.err = 1u;  // watchdog: never hang the GPU
        }
        __syncthreads();
        if (lane < 17) {
          q_lane = __ldcg(traj.state + (size_t)c * 34 + lane);
          ph_lane = __ldcg(traj.state + (size_t)c * 34 + 17 + lane);
        }
      } else if (warp >= 4) {
        const double e = traj.eps[c];
        double mine = 0.0;
        if (lane < 17) mine = fma(0.5 * e, traj.grad[(size_t)c * 17 + lane], traj.p[(size_t)c * 17 + lane]);
        double dot = 0.0;
        for (int j = 0; j < 17; ++j) {
          const double pj = __shfl_sync(0xffffffffu, mine, j);
          if (lane < 17) dot = fma(s_im[lane * 17 + j], pj, dot);
        }
        ph_lane = mine;
        if (lane < 17) q_lane = fma(e, dot, traj.q[(size_t)c * 17 + lane]);
      }
    }
    // parameter k13 of this chain for the calling warp (uniform over the warp)
    auto param13 = [&](int k13) -> double {
      if (TRAJ) {
        const int j = kQOfTheta[k13];
        return backward(__shfl_sync(0xffffffffu, q_lane, j), kQTransform[j]);
      }
      return load_param(theta, theta_is_q, c, k13);
    };

    // ---- phase 0: warps 0-3: one thread per individual reads its int8 column (every load of
    //      the warp is one coalesced 32-byte segment) and applies the infection constraints;
    //      warps 4-7: parameters, power tables, dilution table.  (Fetching the block as aligned
    //      32-bit words + shared-memory atomics / ballots was measured: slower, the transposition
    //      costs more than the byte loads save.) ----
    if (tid < kTileMaxInds) {
      if (tid < ni && step == 0) {
        const int8_t* col = i_raw + (size_t)c * G * N + i0 + tid;
        int8_t bytes[sizeof(M) * 8];
#pragma unroll
        for (int t = 0; t < (int)sizeof(M) * 8; ++t) {
          bytes[t] = (t < G) ? __ldg(col) : (int8_t)0;
          col += N;
        }
        M raw = 0;
#pragma unroll
        for (int t = 0; t < (int)sizeof(M) * 8; ++t) raw |= (M)(bytes[t] != 0) << t;
        const int w = waner[(size_t)c * N + i0 + tid] != 0;
        IndState<M> st;
        st.inf = constrain<M>(raw, my_pcr, dc.ch);
        st.vacw = my_vac | (w ? top_bit<M>() : (M)0);
        s_ind[tid] = st;
        acc[S_KI] = cnt_i = (double)popc(raw);
        acc[S_KW] = cnt_w = (double)w;
      }
      PHASEW(12, tid == 0);
    } else if (warp == 4) {
      fill_pow_warp(param13(N_RHO), G, lane, s_pw[0], s_pw[1]);
      PHASEW(13, lane == 0);
    } else if (warp == 5) {
      fill_pow_warp(param13(S_RHO), G, lane, s_pw[2], s_pw[3]);
    } else if (warp == 6) {
      if (TRAJ) {
        const int j = kQOfTheta[lane < 13 ? lane : 0];
        const double v = backward(__shfl_sync(0xffffffffu, q_lane, j), kQTransform[j]);
        if (lane < 13) s_th[lane] = v;
      } else if (lane < 13) {
        s_th[lane] = load_param(theta, theta_is_q, c, lane);
      }
      PHASEW(14, lane == 0);
    } else if (warp == kSumsWarps - 1) {
      if (kFX) {
        // per-chain table {x_j, exp(-b x_j)} over the distinct dilutions, both antigens, and the
        // cap on b m that keeps the product <= e^700 (b is not transformed: read it directly)
        const double zn = -param13(N_B) * x_lev, zs = -param13(S_B) * x_lev;
        double mn = (lane < dc.n_xlev) ? fabs(zn) : 0.0, ms = (lane < dc.n_xlev) ? fabs(zs) : 0.0;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
          mn = fmax(mn, __shfl_xor_sync(0xffffffffu, mn, off));
          ms = fmax(ms, __shfl_xor_sync(0xffffffffu, ms, off));
        }
        // NaN / huge |b| x: every row evaluates its own exponential (same arithmetic as the fallback)
        const bool direct = !(mn < 300.0 && ms < 300.0);
        s_xe[0][lane] = make_double2(x_lev, direct ? 0.0 : fast_exp(zn, s_tab));
        s_xe[kFX ? 1 : 0][lane] = make_double2(x_lev, direct ? 0.0 : fast_exp(zs, s_tab));
        if (lane == 0) {
          s_zmax[0] = 700.0 - mn;
          s_zmax[1] = 700.0 - ms;
          s_direct = direct;
        }
      }
      if (TRAJ && lane < 17) {  // keeps (q, p_half) for the finaliser
        s_q[lane] = q_lane;
        s_ph[lane] = ph_lane;
      }
      PHASEW(15, lane == 0);
    }
    __syncthreads();
    PHASE(2);
    // The finaliser's parameter-only part (priors, transforms, logs: ~2 us of cold libm code) is
    // taken off the critical path: warp 7 of the chain's first tile computes it now, instead of
    // processing cells / rows, and parks it in global memory for whichever CTA finishes last.
    // (the row -> thread assignment must not depend on fin.mode: abd_sums_dev + abd_finalize_* has
    // to reproduce the fused launch bit for bit)
    const bool aux_cta = (tile == 0);
    const int nwork = aux_cta ? kSumsBlock - 32 : kSumsBlock;
    if (aux_cta && fin.mode != 0 && warp == kSumsWarps - 1) {
      double* a = aux + (size_t)c * kAuxDoubles;
      if (fin.mode == 2 && lane < 17) {
        const PriorPre pp = prior_pre(lane, TRAJ ? s_q[lane] : theta[(size_t)c * 17 + lane], priors->v[lane]);
        double* o = a + lane * 7;
        o[0] = pp.lpA, o[1] = pp.dA, o[2] = pp.f, o[3] = pp.lx, o[4] = pp.l1mx, o[5] = pp.x, o[6] = pp.omx;
      }
      if (lane == 17) {
        const LikPre lk = lik_pre(s_th[N_SIGMA], s_th[S_SIGMA]);
        double* o = a + 17 * 7;
        o[0] = lk.lsn, o[1] = lk.lss, o[2] = lk.ivn, o[3] = lk.ivs, o[4] = lk.isn, o[5] = lk.iss;
      }
    }
    if (cc == 0 && step == 0) mbar_wait(&s_bar, 0);
    PHASE(3);

    // ---- phase 1: one thread per (individual, gap) cell: titer in closed form from the masks ----
    {
      const double init = s_th[N_INIT], perm = s_th[N_PERM], temp = s_th[N_TEMP];
      const double bn_fx = s_th[N_B], zmax_n = s_zmax[0];
      for (int k = cn0 + tid; k < cn1 && tid < nwork; k += nwork) {
        const uint32_t mt = s_cm_n[k - kn0];
        const int t = mt & 63, li = (int)(mt >> 6) - i0;
        double P, T, dT;
        traj_n<M>(s_ind[li].inf, t, s_pw[0], s_pw[1], P, T, dT);
        CellVal cv;
        cv.m = init + perm * P + temp * T;
        cv.T = T;
        cv.dT = (P != 0.0) ? dT : -0.0;  // sign bit of dT carries "never exposed" (P = 0)
        if constexpr (kFX) {
          const double z = bn_fx * cv.m;
          cv.Em = fast_exp((z > zmax_n) ? zmax_n : z, s_tab);  // NaN stays NaN
        }
        s_cv_n[k - cn0] = cv;
      }
    }
    {
      const double init = s_th[S_INIT], perm = s_th[S_PERM];
      const double bs_fx = s_th[S_B], zmax_s = s_zmax[1];
      for (int k = cs0 + tid; k < cs1 && tid < nwork; k += nwork) {
        const uint32_t mt = s_cm_s[k - ks0];
        const int t = mt & 63, li = (int)(mt >> 6) - i0;
        const IndState<M> st = s_ind[li];
        double P, U, dU;
        traj_s<M>(st.inf, st.vacw & ~top_bit<M>(), (st.vacw & top_bit<M>()) != 0, t, s_pw[2], s_pw[3], P, U, dU);
        CellVal cv;
        cv.m = init + perm * P + U;
        cv.T = U;
        cv.dT = (P != 0.0) ? dU : -0.0;
        if constexpr (kFX) {
          const double z = bs_fx * cv.m;
          cv.Em = fast_exp((z > zmax_s) ? zmax_s : z, s_tab);
        }
        s_cv_s[k - cs0] = cv;
      }
    }
    __syncthreads();
    PHASE(4);

    // ---- phase 2: one thread per OD row; everything comes from shared memory, no divergence.
    //      P (ever exposed) travels in the sign bit of dT ----
    //      The loops exist twice: `direct` (a chain whose |b| x is too large to factor, or the
    //      fallback kernel) evaluates one exponential per row, the common case only a product ----
    auto rows = [&](auto direct_tag) {
      constexpr bool kDirect = decltype(direct_tag)::value;
      {
        const double b = s_th[N_B], d = s_th[N_D];
        const uint32_t* rcp = s_rc_n - qn0;
        const double* odp = s_od_n - an0;
        const CellVal* cvp = s_cv_n - cn0;
#pragma unroll 2
        for (int r = rn0 + tid; r < rn1 && tid < nwork; r += nwork) {
          double sg, res, q, xm;
          const uint32_t rc = rcp[r];
          const CellVal cv = cvp[kFX ? (rc >> 5) : rc];
          if constexpr (kFX) {
            const double2 xe = s_xe[0][rc & 31];
            xm = xe.x - cv.m;
            if constexpr (kDirect) row_eval(xe.x, odp[r], cv.m, b, d, s_tab, sg, res, q, xm);
            else row_eval_E(xe.y * cv.Em, odp[r], d, sg, res, q);
          } else {
            row_eval((s_x_n - an0)[r], odp[r], cv.m, b, d, s_tab, sg, res, q, xm);
          }
          acc[SN_0] = fma(res, res, acc[SN_0]);
          acc[SN_1] = fma(res, sg, acc[SN_1]);
          acc[SN_2] = fma(q, xm, acc[SN_2]);
          acc[SN_QINIT] += q;
          acc[SN_QPERM] += (__double2hiint(cv.dT) < 0) ? 0.0 : q;
          acc[SN_QTEMP] = fma(q, cv.T, acc[SN_QTEMP]);
          acc[SN_QRHO] = fma(q, cv.dT, acc[SN_QRHO]);
        }
      }
      {
        const double b = s_th[S_B], d = s_th[S_D];
        const uint32_t* rcp = s_rc_s - qs0;
        const double* odp = s_od_s - as0;
        const CellVal* cvp = s_cv_s - cs0;
#pragma unroll 2
        for (int r = rs0 + tid; r < rs1 && tid < nwork; r += nwork) {
          double sg, res, q, xm;
          const uint32_t rc = rcp[r];
          const CellVal cv = cvp[kFX ? (rc >> 5) : rc];
          if constexpr (kFX) {
            const double2 xe = s_xe[1][rc & 31];
            xm = xe.x - cv.m;
            if constexpr (kDirect) row_eval(xe.x, odp[r], cv.m, b, d, s_tab, sg, res, q, xm);
            else row_eval_E(xe.y * cv.Em, odp[r], d, sg, res, q);
          } else {
            row_eval((s_x_s - as0)[r], odp[r], cv.m, b, d, s_tab, sg, res, q, xm);
          }
          acc[SS_0] = fma(res, res, acc[SS_0]);
          acc[SS_1] = fma(res, sg, acc[SS_1]);
          acc[SS_2] = fma(q, xm, acc[SS_2]);
          acc[SS_QINIT] += q;
          acc[SS_QPERM] += (__double2hiint(cv.dT) < 0) ? 0.0 : q;
          acc[SS_QRHO] = fma(q, cv.dT, acc[SS_QRHO]);
        }
      }
    };
    if (!kFX || s_direct) rows(std::true_type{});
    else rows(std::false_type{});

    PHASE(5);
    // ---- block reduction: butterfly inside a warp, shared memory across warps ----
    {
      const double tot = warp_reduce16(acc, lane);
      if ((lane & 1) == 0) s_red[warp][warp_reduce16_index(lane)] = tot;
    }
    __syncthreads();
    if (tid < kNSums) {
      double v = 0.0;
#pragma unroll
      for (int wv = 0; wv < kSumsWarps; ++wv) v += s_red[wv][tid];
      partial[((size_t)c * ntiles + tile) * kNSums + tid] = v;
    }

    PHASE(6);
    // ---- last CTA of this chain: ordered reduction over tiles, then finalise.  One thread
    //      releases the CTA's partials (barrier, then fence + ticket) and acquires the others'
    //      (then barrier): the grid-sync idiom with ONE acq_rel atomic per CTA instead of a pair of
    //      sequentially consistent fences around a relaxed one ----
    __syncthreads();
    if (tid == 0) {
      unsigned prev;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(ticket + c) : "memory");
      s_last = (prev == (unsigned)(ntiles - 1));
    }
    __syncthreads();
    PHASE(7);
    if (s_last) {
      // the parameter-only part of the finaliser was parked in `aux` by the chain's first tile
      if (fin.mode && tid < kAuxDoubles) {
        const double v = __ldcg(aux + (size_t)c * kAuxDoubles + tid);
        if (tid < 17 * 7) reinterpret_cast<double*>(s_pre)[tid] = v;
        else if (tid < 17 * 7 + 6) reinterpret_cast<double*>(&s_lik)[tid - 17 * 7] = v;
      }
      const int k = tid & 15, g = tid >> 4;  // 16 groups of 16 values
      double v = 0.0;
      const double* src = partial + (size_t)c * ntiles * kNSums + k;
      for (int tl = g; tl < ntiles; tl += 8 * (kSumsBlock / 16)) {  // 8 loads in flight, fixed order
        double ld[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int tt = tl + u * (kSumsBlock / 16);
          ld[u] = (tt < ntiles) ? __ldcg(src + (size_t)tt * kNSums) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) v += ld[u];
      }
      s_fin[g][k] = v;
      PHASE(9);
      __syncthreads();
      if (tid < kNSums) {
        double tot = 0.0;
#pragma unroll
        for (int gg = 0; gg < kSumsBlock / 16; ++gg) tot += s_fin[gg][tid];
        s_red[0][tid] = tot;
        if (sums) sums[(size_t)c * kNSums + tid] = tot;
      }
      __syncthreads();
      if (!TRAJ && xch.world > 1) {
        __shared__ unsigned s_seq;
        if (tid == 0) s_seq = xch.seq[c] + 1u;
        __syncthreads();
        const unsigned seq = s_seq, par = seq & 1u;
        const size_t slot = ((size_t)par * xch.world + xch.rank) * xch.cmax + c;  // my slot in a peer's buffer
        if (tid < kNSums * xch.world) {
          const int r = tid >> 4, k = tid & 15;
          xch.data[r][slot * kNSums + k] = s_red[0][k];
        }
        __threadfence_system();
        __syncthreads();
        if (tid < xch.world) {
          unsigned long long* f = xch.flag[tid] + slot;
          asm volatile("st.release.sys.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)seq) : "memory");
          // wait for peer `tid`'s contribution to arrive in MY buffer
          const unsigned long long* mine = xch.flag[xch.rank] + ((size_t)par * xch.world + tid) * xch.cmax + c;
          unsigned long long seen, t_start, t_now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
          do {
            asm volatile("ld.acquire.sys.u64 %0, [%1];" : "=l"(seen) : "l"(mine) : "memory");
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_now));
          } while (seen != (unsigned long long)seq && t_now - t_start < kXchTimeoutNs);
          if (seen != (unsigned long long)seq) *xch.err = 1u;  // watchdog: never hang the GPU
        }
        __syncthreads();
        if (tid < kNSums) {
          double tot = 0.0;
          for (int r = 0; r < xch.world; ++r)
            tot += __ldcg(xch.data[xch.rank] + (((size_t)par * xch.world + r) * xch.cmax + c) * kNSums + tid);
          s_red[0][tid] = tot;
          if (sums) sums[(size_t)c * kNSums + tid] = tot;
        }
        if (tid == 0) xch.seq[c] = seq;
        __syncthreads();
      }
      PHASE(10);
      if (tid == 0) ticket[c] = 0;  // re-arm for the next launch
      if (warp == 0) {
        if (fin.mode == 1) {
          if (lane == 0)
            finalize_loglik_post(s_th, s_lik, s_red[0], fin.tot, &fin.out_val[c],
                                 fin.out_grad ? &fin.out_grad[(size_t)c * 13] : nullptr);
        } else if (fin.mode == 2 && !TRAJ) {
          finalize_logp_post(lane, s_pre, s_th, s_lik, s_red[0], fin.tot, &fin.out_val[c],
                             fin.out_grad ? &fin.out_grad[(size_t)c * 17] : nullptr);
        } else if (TRAJ && fin.mode == 2) {
          // trajectory mode: finish this leapfrog step and either publish the next position or
          // write the end point
          finalize_logp_post(lane, s_pre, s_th, s_lik, s_red[0], fin.tot, &s_out[0], &s_out[1]);
          __syncwarp();
          const double e = traj.eps[c];
          const double g = lane < 17 ? s_out[1 + lane] : 0.0;
          const double ph = lane < 17 ? s_ph[lane] : 0.0;
          if (step == nsteps - 1) {
            if (lane < 17) {
              traj.q[(size_t)c * 17 + lane] = s_q[lane];
              traj.p[(size_t)c * 17 + lane] = fma(0.5 * e, g, ph);
              traj.grad[(size_t)c * 17 + lane] = g;
            }
            if (lane == 0) traj.logp[c] = s_out[0];
          } else {
            const double ph2 = fma(e, g, ph);  // two half steps: end of this step + start of the next
            double dot = 0.0;
            for (int j = 0; j < 17; ++j) {
              const double pj = __shfl_sync(0xffffffffu, ph2, j);
              if (lane < 17) dot = fma(s_im[lane * 17 + j], pj, dot);
            }
            if (lane < 17) {
              traj.state[(size_t)c * 34 + lane] = fma(e, dot, s_q[lane]);
              traj.state[(size_t)c * 34 + 17 + lane] = ph2;
            }
            __threadfence();
            __syncwarp();
            if (lane == 0) {
              const unsigned nxt = (unsigned)step + 1u;
              asm volatile("st.release.gpu.u32 [%0], %1;" ::"l"(traj.gen + c), "r"(nxt) : "memory");
            }
          }
        }
      }
    }
    if (s_last) PHASE(11);
    PHASE(8);
    __syncthreads();  // shared memory is reused by the next step / chain
    }  // step
  }
  SPAN_END();
}

// one warp per chain
__global__ void k_finalize(const int C, const double* __restrict__ theta,
                           const double* __restrict__ sums, const FinalizeCfg fin,
                           const Priors* __restrict__ priors) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  if (fin.mode == 1) {
    if (lane == 0)
      finalize_loglik(&theta[(size_t)c * 13], &sums[(size_t)c * kNSums], fin.tot, &fin.out_val[c],
                      fin.out_grad ? &fin.out_grad[(size_t)c * 13] : nullptr);
  } else {
    __shared__ double s_th[4][16];  // 128 threads = 4 chains per CTA
    __shared__ PriorPre s_pre[4][17];
    const int wv = threadIdx.x >> 5;
    if (lane < 13) s_th[wv][lane] = load_param(theta, 1, c, lane);
    if (lane < 17) s_pre[wv][lane] = prior_pre(lane, theta[(size_t)c * 17 + lane], priors->v[lane]);
    __syncwarp();
    finalize_logp_post(lane, s_pre[wv], s_th[wv], lik_pre(s_th[wv][N_SIGMA], s_th[wv][S_SIGMA]),
                       &sums[(size_t)c * kNSums], fin.tot, &fin.out_val[c],
                       fin.out_grad ? &fin.out_grad[(size_t)c * 17] : nullptr);
  }
}

// ------------------------------------------------------------------------------------------
// k_gibbs
// ------------------------------------------------------------------------------------------
struct GibbsCfg {
  uint64_t seed, sweep;
  int mode;          // ABD_GIBBS_*; -1 = conditional log-odds only (no update, no RNG)
  double transit_p;
  double* out_i;     // [C][G][N] (mode -1)
  double* out_w;     // [C][N]
  unsigned long long* stats;  // [C][2] proposals, accepted flips (may be null)
};

// One OD row of one individual under a candidate state (value only, times -1/(2 sigma^2)).
// N and S rows share one code path: lanes differ only in their per-lane parameters.
struct RowPar {
  double init, perm, tfac, b, d, nh;
};
template <typename M>
__device__ __forceinline__ double gibbs_row(double x, double od, int t, bool is_s, M inf, M vac, int w,
                                            const RowPar& rp, const double (*s_pw)[kMaxGaps],
                                            const double* __restrict__ s_tab) {
  const M lm = low_mask<M>(t);
  M e = inf & lm, v = is_s ? (vac & lm) : (M)0;
  const double P = (e | v) ? rp.perm : 0.0;
  const double* pw = is_s ? (w ? s_pw[1] : s_pw[2]) : s_pw[0];  // s_pw[2] = all ones (rho_ind = 1)
  double T = 0.0;
  while (e) {
    T += pw[t - ctz(e)];
    e &= e - 1;
  }
  while (v) {
    T += pw[t - ctz(v)];
    v &= v - 1;
  }
  return rp.nh * row_resid2(x, od, fma(rp.tfac, T, rp.init + P), rp.b, rp.d, s_tab);
}

// Persistent warps: every warp repeatedly claims the next (chain, individual) from a global
// counter.  Items are ordered chain-major and, inside a chain, by decreasing OD-row count
// (longest job first), so the tail at the end of the launch is one short job.
template <typename M>
__global__ void __launch_bounds__(kGibbsWarps * 32, 3)
k_gibbs(const DevCohort dc, const int* __restrict__ order, const int C,
        const double* __restrict__ theta, const int theta_is_q,
        const double* __restrict__ p_arr, const double* __restrict__ pw_arr,
        int8_t* __restrict__ i_raw, int8_t* __restrict__ waner, unsigned* __restrict__ queue,
        const GibbsCfg cfg) {
  constexpr int NSLOT = sizeof(M) / 4;  // proposals owned per lane
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = dc.G, N = dc.N;

  __shared__ double s_tab[kExpTab];
  __shared__ double s_th_all[kGibbsWarps][20];            // theta13, -1/(2 sigma^2) x 2, logit p / p_w, p / p_w
  __shared__ double s_pw_all[kGibbsWarps][3][kMaxGaps];   // rho_n^k, rho_s^k, ones
  __shared__ unsigned char s_ord[kGibbsWarps][kMaxGaps];
  double* s_th = s_th_all[warp];
  const double (*s_pw)[kMaxGaps] = s_pw_all[warp];

  fill_exp_table(s_tab, tid, kGibbsWarps * 32);
  for (int k = lane; k < kMaxGaps; k += 32) s_pw_all[warp][2][k] = 1.0;
  __syncthreads();

  const int nprop = G + 1;  // proposal j < G flips i_raw[j, n]; j == G flips waner[n]
  const unsigned n_items = (unsigned)C * (unsigned)N;
  unsigned n_prop = 0, n_acc = 0;
  int cur_c = -1;

  while (true) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(queue, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int c = (int)(item / (unsigned)N);
    const int n = order[item - (unsigned)c * (unsigned)N];

    if (c != cur_c) {  // (re)load this chain's parameters into the warp's shared-memory slot
      if (cur_c >= 0 && cfg.stats && lane == 0) {
        atomicAdd(&cfg.stats[(size_t)cur_c * 2], (unsigned long long)n_prop);
        atomicAdd(&cfg.stats[(size_t)cur_c * 2 + 1], (unsigned long long)n_acc);
      }
      n_prop = n_acc = 0;
      cur_c = c;
      __syncwarp();
      fill_pow_warp(load_param(theta, theta_is_q, c, N_RHO), G, lane, s_pw_all[warp][0], nullptr);
      fill_pow_warp(load_param(theta, theta_is_q, c, S_RHO), G, lane, s_pw_all[warp][1], nullptr);
      if (lane < 13) s_th[lane] = load_param(theta, theta_is_q, c, lane);
      if (lane == 13 || lane == 14) {
        const double sg = load_param(theta, theta_is_q, c, lane == 13 ? N_SIGMA : S_SIGMA);
        s_th[lane] = -0.5 / (sg * sg);
      }
      if (lane == 15 || lane == 16) {
        const int which = lane - 15;
        double lo;
        if (theta_is_q) {  // logit of a logodds-transformed value is the value itself
          lo = theta[(size_t)c * 17 + (which ? kQ_PW : kQ_P)];
        } else {
          const double p = which ? pw_arr[c] : p_arr[c];
          lo = log(p) - log1p(-p);
        }
        s_th[lane] = lo;
        s_th[lane + 2] = 1.0 / (1.0 + exp(-lo));
      }
      __syncwarp();
    }

    // ---- this individual's column of i_raw (lane t reads gap t), waner, masks, rows ----
    int8_t* col = i_raw + (size_t)c * G * N + n;
    M raw = 0;
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) {
      const int t = lane + 32 * sl;
      const int8_t b = (t < G) ? col[(size_t)t * N] : (int8_t)0;
      raw |= (M)__ballot_sync(0xffffffffu, b != 0) << (32 * sl);
    }
    const M raw_in = raw;
    int w = waner[(size_t)c * N + n] != 0;
    const int w_in = w;
    const M pcr = reinterpret_cast<const M*>(dc.pcr)[n];
    const M vac = reinterpret_cast<const M*>(dc.vac)[n];
    const int rn0 = dc.rp[0][n], cnt_n = dc.rp[0][n + 1] - rn0;
    const int rs0 = dc.rp[1][n], cnt_s = dc.rp[1][n + 1] - rs0;
    const int nrows = cnt_n + cnt_s;

    // rows of this individual, one per lane (N rows first, then S rows); kept in registers
    auto load_row = [&](int l, double& x, double& od, int& t, bool& is_s) {
      is_s = l >= cnt_n;
      const int r = is_s ? rs0 + (l - cnt_n) : rn0 + l;
      if (l < nrows) {
        x = (is_s ? dc.x[1] : dc.x[0])[r];
        od = (is_s ? dc.od[1] : dc.od[0])[r];
        t = (int)((is_s ? dc.meta[1] : dc.meta[0])[r] & 63u);
      } else {
        x = 0.0;
        od = 0.0;
        t = -1;
      }
    };
    auto row_par = [&](bool is_s) {
      RowPar rp;
      rp.init = s_th[is_s ? S_INIT : N_INIT];
      rp.perm = s_th[is_s ? S_PERM : N_PERM];
      rp.tfac = is_s ? 1.0 : s_th[N_TEMP];
      rp.b = s_th[is_s ? S_B : N_B];
      rp.d = s_th[is_s ? S_D : N_D];
      rp.nh = s_th[is_s ? 14 : 13];
      return rp;
    };
    double x0, od0;
    int t0;
    bool s0;
    load_row(lane, x0, od0, t0, s0);
    const RowPar rp0 = row_par(s0);
    // latest sampled gap of either antigen / of the S antigen (rows are sorted by gap)
    int t_last = t0, t_last_s = s0 ? t0 : -1;
    for (int l = lane + 32; l < nrows; l += 32) {
      const bool is_s = l >= cnt_n;
      const int r = is_s ? rs0 + (l - cnt_n) : rn0 + l;
      const int t = (int)((is_s ? dc.meta[1] : dc.meta[0])[r] & 63u);
      t_last = max(t_last, t);
      if (is_s) t_last_s = max(t_last_s, t);
    }
    t_last = __reduce_max_sync(0xffffffffu, t_last);
    t_last_s = __reduce_max_sync(0xffffffffu, t_last_s);

    auto indiv_ll = [&](M inf_, int w_) {
      double a = (t0 >= 0) ? gibbs_row<M>(x0, od0, t0, s0, inf_, vac, w_, rp0, s_pw, s_tab) : 0.0;
      for (int l = lane + 32; l < nrows; l += 32) {  // individuals with more than 32 rows
        double x, od;
        int t;
        bool is_s;
        load_row(l, x, od, t, is_s);
        a += gibbs_row<M>(x, od, t, is_s, inf_, vac, w_, row_par(is_s), s_pw, s_tab);
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
      return a;
    };

    M inf = constrain<M>(raw, pcr, dc.ch);
    double ll = indiv_ll(inf, w);

    // ---- lane-owned proposals: Philox key (visiting order), transit / accept uniforms ----
    M act = 0;              // steps that are not skipped
    int jp_step[NSLOT];     // lane k (+32 sl): proposal visited at step k
    double acc_u[NSLOT];    // log(u) (Metropolis) or u (heat bath) of the proposals this lane owns
#pragma unroll
    for (int sl = 0; sl < NSLOT; ++sl) acc_u[sl] = 0.0;
    if (cfg.mode >= 0) {
      uint32_t key[NSLOT];
      int rank[NSLOT];
      bool skip[NSLOT];
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const uint4 r = philox4x32_10(
            make_uint4((uint32_t)(lane + 32 * sl), (uint32_t)n + dc.ind_offset, (uint32_t)c, (uint32_t)cfg.sweep),
            make_uint2((uint32_t)cfg.seed, (uint32_t)(cfg.seed >> 32) ^ (uint32_t)(cfg.sweep >> 32)));
        key[sl] = r.x;
        const double u_t = u01(r.y), u_a = u01(r.z);
        skip[sl] = (cfg.mode == ABD_GIBBS_METROPOLIS) && !(u_t <= cfg.transit_p);
        acc_u[sl] = (cfg.mode == ABD_GIBBS_METROPOLIS) ? log(u_a) : u_a;
        rank[sl] = 0;
      }
      for (int jj = 0; jj < nprop; ++jj) {
        const uint32_t kj = __shfl_sync(0xffffffffu, key[jj >> 5], jj & 31);
#pragma unroll
        for (int sl = 0; sl < NSLOT; ++sl) {
          const int me = lane + 32 * sl;
          rank[sl] += (kj < key[sl]) || (kj == key[sl] && jj < me);
        }
      }
      __syncwarp();
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int me = lane + 32 * sl;
        if (me < nprop) s_ord[warp][rank[sl]] = (unsigned char)me;
      }
      __syncwarp();
      // the proposal visited at the step this lane stands for; skipped steps drop out of `act`
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int step = lane + 32 * sl;
        const int jp = (step < nprop) ? (int)s_ord[warp][step] : 0;
        jp_step[sl] = jp;
        bool sk = __shfl_sync(0xffffffffu, (int)skip[0], jp & 31);
        if (NSLOT > 1) {
          const bool sk1 = __shfl_sync(0xffffffffu, (int)skip[NSLOT - 1], jp & 31);
          if (jp >> 5) sk = sk1;
        }
        act |= (M)__ballot_sync(0xffffffffu, step < nprop && !sk) << (32 * sl);
      }
    } else {
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) jp_step[sl] = lane + 32 * sl;
      act = low_mask<M>(G);  // every proposal 0..G, in order
    }

    // Everything about a proposal that depends only on the current state lives in the lane that
    // owns it and is refreshed (lane-parallel) after an accepted flip of an infection bit:
    //   inf2  constrained infections if the flip were accepted
    //   code  2: the data term changes, the individual's likelihood must be evaluated;
    //         1 / 0: it cannot change (the constrained infections differ only after the last
    //         sampled gap, or -- waner bit -- there is no S sample after the first exposure), so
    //         logp(prop) - logp(cur) = +-logit(p) and the flip decision (1 / 0) is already known
    M inf2_own[NSLOT];
    int code_own[NSLOT];
    auto refresh = [&]() {
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int me = lane + 32 * sl;
        const bool is_w = (me == G);
        const M i2 = (me < G) ? constrain<M>(raw ^ ((M)1 << me), pcr, dc.ch) : inf;
        inf2_own[sl] = i2;
        const int cur_bit = is_w ? w : (int)((raw >> (me < G ? me : 0)) & 1);
        bool affected;
        if (is_w) {
          const M ex = inf | vac;
          affected = ex != 0 && ctz(ex | top_bit<M>()) < t_last_s;
        } else {
          const M diff = inf ^ i2;
          affected = diff != 0 && ctz(diff | top_bit<M>()) <= t_last;
        }
        bool flip0;
        if (cfg.mode == ABD_GIBBS_METROPOLIS) {
          const double lo = s_th[is_w ? 16 : 15];
          const double delta = cur_bit ? -lo : lo;
          flip0 = isfinite(delta) && (acc_u[sl] < delta);
        } else {
          flip0 = ((acc_u[sl] <= s_th[is_w ? 18 : 17]) ? 1 : 0) != cur_bit;  // s_th[17 / 18] = sigmoid(logit)
        }
        code_own[sl] = affected ? 2 : (flip0 ? 1 : 0);
      }
    };
    refresh();

    while (act) {
      const int step = ctz(act);
      act &= act - 1;
      int jp = jp_step[0];
      if (NSLOT > 1 && (step >> 5)) jp = jp_step[NSLOT - 1];
      jp = __shfl_sync(0xffffffffu, jp, step & 31);
      const int owner = jp & 31;
      const bool hi = NSLOT > 1 && (jp >> 5);
      const int code = __shfl_sync(0xffffffffu, hi ? code_own[NSLOT - 1] : code_own[0], owner);
      const bool is_w = (jp == G);
      if (cfg.mode >= 0) {
        ++n_prop;
        if (code == 0) continue;
      }
      const M inf2 = __shfl_sync(0xffffffffu, hi ? inf2_own[NSLOT - 1] : inf2_own[0], owner);
      const int w2 = is_w ? (w ^ 1) : w;
      double ll2 = ll;
      bool flip = true;
      if (code == 2 || cfg.mode < 0) {
        const int cur_bit = is_w ? w : (int)((raw >> jp) & 1);
        if (code == 2) ll2 = indiv_ll(inf2, w2);
        const double lo = s_th[is_w ? 16 : 15];
        const double d10 = cur_bit ? (ll - ll2 + lo) : (ll2 - ll + lo);  // log-odds of 1 versus 0
        if (cfg.mode < 0) {
          if (lane == 0) {
            if (is_w) cfg.out_w[(size_t)c * N + n] = d10;
            else cfg.out_i[((size_t)c * G + jp) * N + n] = d10;
          }
          continue;
        }
        const double au = __shfl_sync(0xffffffffu, hi ? acc_u[NSLOT - 1] : acc_u[0], owner);
        if (cfg.mode == ABD_GIBBS_METROPOLIS) {
          const double delta = cur_bit ? -d10 : d10;  // logp(proposed) - logp(current)
          flip = isfinite(delta) && (au < delta);
        } else {
          const double p1 = 1.0 / (1.0 + exp(-d10));
          flip = ((au <= p1) ? 1 : 0) != cur_bit;
        }
      }
      if (flip) {
        ++n_acc;
        ll = ll2;
        if (is_w) {
          w = w2;
        } else {
          raw ^= (M)1 << jp;
          inf = inf2;
          refresh();
        }
      }
    }

    if (cfg.mode >= 0) {  // write back only what changed
      const M changed = raw ^ raw_in;
#pragma unroll
      for (int sl = 0; sl < NSLOT; ++sl) {
        const int t = lane + 32 * sl;
        if (t < G && ((changed >> t) & 1)) col[(size_t)t * N] = (int8_t)((raw >> t) & 1);
      }
      if (lane == 0 && w != w_in) waner[(size_t)c * N + n] = (int8_t)w;
    }
  }
  if (cfg.mode >= 0 && cfg.stats && lane == 0 && cur_c >= 0) {
    atomicAdd(&cfg.stats[(size_t)cur_c * 2], (unsigned long long)n_prop);
    atomicAdd(&cfg.stats[(size_t)cur_c * 2 + 1], (unsigned long long)n_acc);
  }
}

// ------------------------------------------------------------------------------------------
// k_hmc_begin / k_hmc_end -- the two ends of an HMC transition, one warp per chain
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_hmc_begin(const int C, const double* __restrict__ q, const double* __restrict__ grad, const double* __restrict__ logp,
            const double* __restrict__ linv_t, const uint64_t seed, const uint64_t iter, double* __restrict__ qw,
            double* __restrict__ pw, double* __restrict__ gw, double* __restrict__ h0) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  double z = 0.0;
  if (lane < 17) {  // Box-Muller on two of the four Philox words
    const uint4 r = philox4x32_10(make_uint4((uint32_t)lane, (uint32_t)c, (uint32_t)iter, 0x484d4331u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iter >> 32)));
    z = sqrt(-2.0 * log(u01(r.x))) * cospi(2.0 * u01(r.y));
  }
  double p = 0.0, kin = z * z;
  for (int j = 0; j < 17; ++j) {
    const double zj = __shfl_sync(0xffffffffu, z, j);
    if (lane < 17) p = fma(linv_t[lane * 17 + j], zj, p);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) kin += __shfl_xor_sync(0xffffffffu, kin, off);
  if (lane < 17) {
    qw[(size_t)c * 17 + lane] = q[(size_t)c * 17 + lane];
    gw[(size_t)c * 17 + lane] = grad[(size_t)c * 17 + lane];
    pw[(size_t)c * 17 + lane] = p;
  }
  if (lane == 0) h0[c] = -logp[c] + 0.5 * kin;
}

__global__ void __launch_bounds__(128)
k_hmc_end(const int C, double* __restrict__ q, double* __restrict__ grad, double* __restrict__ logp,
          const double* __restrict__ qw, const double* __restrict__ pw, const double* __restrict__ gw,
          const double* __restrict__ lpw, const double* __restrict__ inv_mass, const double* __restrict__ h0,
          const uint64_t seed, const uint64_t iter, double* __restrict__ accept_out, double* __restrict__ da,
          double* __restrict__ eps, const int adapt, const double target) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  const double p = lane < 17 ? pw[(size_t)c * 17 + lane] : 0.0;
  double v = 0.0;
  for (int j = 0; j < 17; ++j) {
    const double pj = __shfl_sync(0xffffffffu, p, j);
    if (lane < 17) v = fma(inv_mass[lane * 17 + j], pj, v);
  }
  double kin = p * v;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) kin += __shfl_xor_sync(0xffffffffu, kin, off);
  const double h1 = -lpw[c] + 0.5 * kin;
  double dh = h0[c] - h1;
  if (!isfinite(dh)) dh = -INFINITY;
  const double acc = exp(fmin(dh, 0.0));
  const uint4 r = philox4x32_10(make_uint4(0u, (uint32_t)c, (uint32_t)iter, 0x41434331u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iter >> 32)));
  const bool take = (u01(r.x) - 2.3283064365386963e-10) < acc;  // u in [0, 1)
  if (take && lane < 17) {
    q[(size_t)c * 17 + lane] = qw[(size_t)c * 17 + lane];
    grad[(size_t)c * 17 + lane] = gw[(size_t)c * 17 + lane];
  }
  if (lane == 0) {
    if (take) logp[c] = lpw[c];
    accept_out[c] = acc;
    if (adapt) {  // Nesterov dual averaging of log(step size): gamma 0.05, t0 10, kappa 0.75
      double* s = da + (size_t)c * 4;
      const double t = s[3] + 1.0, eta = 1.0 / (t + 10.0);
      const double hbar = (1.0 - eta) * s[1] + eta * (target - acc);
      const double log_eps = s[0] - sqrt(t) / 0.05 * hbar;
      const double w = pow(t, -0.75);
      s[1] = hbar;
      s[2] = w * log_eps + (1.0 - w) * s[2];
      s[3] = t;
      eps[c] = exp(log_eps);
    }
  }
}

// ------------------------------------------------------------------------------------------
// k_determ
// ------------------------------------------------------------------------------------------
template <typename M>
__global__ void __launch_bounds__(128)
k_determ(const DevCohort dc, const double* __restrict__ theta13, const int8_t* __restrict__ i_raw,
         const int8_t* __restrict__ waner, int8_t* __restrict__ out_i, double* __restrict__ out_mu_n,
         double* __restrict__ out_mu_s) {
  const int c = blockIdx.y, n = blockIdx.x * blockDim.x + threadIdx.x;
  const int G = dc.G, N = dc.N;
  __shared__ double s_th[16];
  if (threadIdx.x < 13) s_th[threadIdx.x] = theta13[(size_t)c * 13 + threadIdx.x];
  __syncthreads();
  if (n >= N) return;
  M raw = 0;
  const int8_t* col = i_raw + (size_t)c * G * N + n;
  for (int t = 0; t < G; ++t) raw |= (M)(col[(size_t)t * N] != 0) << t;
  const int w = waner[(size_t)c * N + n] != 0;
  const M inf = constrain<M>(raw, reinterpret_cast<const M*>(dc.pcr)[n], dc.ch);
  const M vac = reinterpret_cast<const M*>(dc.vac)[n];
  const double rho_n = s_th[N_RHO], rho_s = w ? s_th[S_RHO] : 1.0;
  double T = 0.0, U = 0.0, Pn = 0.0, Ps = 0.0;
  for (int t = 0; t < G; ++t) {  // the recurrence of abd.py:277-293
    const int it = (int)((inf >> t) & 1), vt = (int)((vac >> t) & 1);
    T = T * rho_n + it;
    U = U * rho_s + (it + vt);
    if (it) Pn = 1.0;
    if (it | vt) Ps = 1.0;
    const size_t o = ((size_t)c * G + t) * N + n;
    if (out_i) out_i[o] = (int8_t)it;
    if (out_mu_n) out_mu_n[o] = s_th[N_PERM] * Pn + s_th[N_TEMP] * T + s_th[N_INIT];
    if (out_mu_s) out_mu_s[o] = s_th[S_PERM] * Ps + U + s_th[S_INIT];
  }
}

// pinned host memory -> device memory by the SMs (see copy_state_h2d); n16 16-byte words + rem bytes
__global__ void __launch_bounds__(256)
k_pull(const uint4* __restrict__ src, uint4* __restrict__ dst, const size_t n16, const size_t rem) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent loads in flight per thread
    const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a, dst[i + stride] = b, dst[i + 2 * stride] = c, dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = src[i];
  if (blockIdx.x == 0 && threadIdx.x < rem)
    reinterpret_cast<uint8_t*>(dst + n16)[threadIdx.x] = reinterpret_cast<const uint8_t*>(src + n16)[threadIdx.x];
}

__global__ void k_debug_fast_math(const long long n, const double* __restrict__ z, double* __restrict__ out_exp,
                                  double* __restrict__ out_rcp) {
  __shared__ double s_tab[kExpTab];
  fill_exp_table(s_tab, threadIdx.x, blockDim.x);
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out_exp[i] = fast_exp(z[i], s_tab);
    out_rcp[i] = fast_rcp(1.0 + fabs(z[i]));
  }
}

}  // namespace

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
struct abd_handle {
  int device = 0;
  int G = 0, N = 0;
  bool wide = false;  // uint64 masks
  int64_t R[2] = {0, 0};
  Totals tot{};
  DevCohort dc{};
  std::vector<void*> owned;   // device allocations freed in abd_destroy
  std::vector<int> h_rp[2];   // host copy of the row CSR pointers (for tilings)
  std::vector<int> h_cp[2];   // host copy of the cell CSR pointers
  cudaStream_t stream = nullptr;
  Priors* d_priors = nullptr;
  int64_t launches = 0;

  struct Tiling {
    int ntiles = 0;
    void* d_tiles = nullptr;   // TileDesc[ntiles]
    int cap_n = 0, cap_s = 0, capr_n = 0, capr_s = 0, capk_n = 0, capk_s = 0;  // staging capacities (elements)
    size_t smem = 0;                                     // dynamic shared memory per CTA
    int occ = 0;                                         // resident CTAs per SM (0 = not queried yet)
  };
  std::map<int, Tiling> tilings;  // keyed by number of tiles requested
  int tile_rows_override = 0;     // 0 = automatic
  int chains_per_cta_override = 0;
  int n_sms = 148;
  size_t smem_optin = 0;
  int* d_order = nullptr;       // individuals by decreasing OD-row count (Gibbs work queue)
  unsigned* d_queue = nullptr;  // Gibbs work-queue counter
  int gibbs_ctas = 0;
  // peer exchange (individual sharding over NVLink)
  XchCfg xch{};
  void* xch_local = nullptr;
  bool xch_active = false;  // set for the duration of a fused sharded launch
  void* xch_peer[kMaxPeers] = {};
  std::vector<double> x_levels; // distinct log_dilution values (ascending) when there are <= 32 of them
  bool fx = false;              // factored mode: rows carry (cell, dilution index), see k_sums
  bool use_pdl = true;          // programmatic dependent launch (ABD_B200_NO_PDL=1 disables)
  bool use_pull = true;         // SM-driven upload of pinned chain state (ABD_B200_NO_PULL=1 disables)

  // per-chain scratch
  int cap_chains = 0;
  int8_t* d_iraw = nullptr;
  int8_t* d_waner = nullptr;
  double* d_theta = nullptr;   // [C][17]
  double* d_p = nullptr;       // [C][2]
  double* d_sums = nullptr;    // [C][16]
  double* d_out = nullptr;     // [C][18]
  unsigned* d_ticket = nullptr;
  double* d_aux = nullptr;     // [C][kAuxDoubles] finaliser inputs that depend on parameters only
  double* d_traj = nullptr;    // [C][34] trajectory-mode scratch (position, half-step momentum)
  unsigned* d_gen = nullptr;   // [C + 1] generation counters + error flag
  unsigned long long* d_stats = nullptr;
  double* d_partial = nullptr;
  size_t cap_partial = 0;
  double* h_pin = nullptr;     // pinned staging [C][40]
  double* h_pin_dev = nullptr; // the same buffer as the device sees it
};

namespace {

template <typename T>
int dev_alloc(abd_handle* h, T** p, size_t n, bool owned = true) {
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T));
  if (e != cudaSuccess) return fail(ABD_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  *p = (T*)q;
  if (owned) h->owned.push_back(q);
  return ABD_OK;
}

template <typename T>
int upload(abd_handle* h, T** dptr, const std::vector<T>& v) {
  int rc = dev_alloc(h, dptr, v.size());
  if (rc) return rc;
  if (!v.empty()) CU(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return ABD_OK;
}

Priors default_priors(int G) {
  // abd.py:424 p ~ Beta(1, G-1); :329-340 N response; :367-388 S response; :464-467 sigmoids
  auto normal = [](double mu, double sd) { return PriorSpec{0, mu, sd, -kHalfLog2Pi - std::log(sd)}; };
  auto gamma = [](double mu, double sd) {
    const double a = mu * mu / (sd * sd), b = mu / (sd * sd);
    return PriorSpec{1, a, b, -std::lgamma(a) + a * std::log(b)};
  };
  auto beta = [](double a, double b) {
    return PriorSpec{2, a, b, -(std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b))};
  };
  auto expo = [](double lam) { return PriorSpec{3, lam, 0.0, std::log(lam)}; };
  Priors p;
  p.v[0] = beta(1.0, (double)(G - 1));
  p.v[1] = gamma(2.0, 0.5);
  p.v[2] = gamma(1.0, 0.5);
  p.v[3] = beta(10.0, 1.0);
  p.v[4] = normal(-2.0, 1.0);
  p.v[5] = gamma(2.0, 0.5);
  p.v[6] = beta(10.0, 1.0);
  p.v[7] = beta(1.0, 1.0);
  p.v[8] = gamma(1.0, 0.5);
  p.v[9] = gamma(1.0, 0.5);
  p.v[10] = normal(-2.0, 1.0);
  p.v[11] = normal(-1.0, 0.5);
  p.v[12] = normal(2.0, 0.5);
  p.v[13] = expo(1.0);
  p.v[14] = normal(-1.0, 0.5);
  p.v[15] = normal(2.0, 0.5);
  p.v[16] = expo(1.0);
  return p;
}

// CSR by individual for one antigen (counting sort by individual, then by gap inside)
int build_rows(abd_handle* h, int a, int64_t R, const double* x, const double* od, const int32_t* gap,
               const int32_t* ind) {
  const int N = h->N, G = h->G;
  if (R > 0 && (!x || !od || !gap || !ind)) return fail(ABD_ERR_INVALID, "NULL row array");
  if (R >= (int64_t)1 << 31) return fail(ABD_ERR_INVALID, "too many OD rows for int32 CSR");
  std::vector<int>& rp = h->h_rp[a];
  rp.assign((size_t)N + 1, 0);
  for (int64_t r = 0; r < R; ++r) {
    if (ind[r] < 0 || ind[r] >= N) return fail(ABD_ERR_INVALID, "individual index out of range");
    if (gap[r] < 0 || gap[r] >= G) return fail(ABD_ERR_INVALID, "gap index out of range");
    rp[(size_t)ind[r] + 1]++;
  }
  for (int n = 0; n < N; ++n) rp[n + 1] += rp[n];
  std::vector<int64_t> order((size_t)R);
  {
    std::vector<int> cur(rp.begin(), rp.end() - 1);
    for (int64_t r = 0; r < R; ++r) order[(size_t)cur[ind[r]]++] = r;
  }
  for (int n = 0; n < N; ++n)
    std::stable_sort(order.begin() + rp[n], order.begin() + rp[n + 1],
                     [&](int64_t p, int64_t q) { return gap[p] < gap[q]; });
  // padded so that the 16-byte-aligned bulk copies of k_sums may read past the last row
  std::vector<double> xs((size_t)R + 2, 0.0), ods((size_t)R + 2, 0.0);
  std::vector<uint32_t> meta((size_t)R + 4, 0u);
  // cells: the distinct (individual, gap) pairs, in row order; every row points at its cell
  std::vector<uint32_t> rowcell((size_t)R + 4, 0u), cmeta;
  std::vector<int>& cp = h->h_cp[a];
  cp.assign((size_t)N + 1, 0);
  for (int64_t k = 0; k < R; ++k) {
    const int64_t r = order[(size_t)k];
    xs[(size_t)k] = x[r];
    ods[(size_t)k] = od[r];
    const uint32_t m = ((uint32_t)ind[r] << 6) | (uint32_t)gap[r];
    meta[(size_t)k] = m;
    if (cmeta.empty() || cmeta.back() != m) {
      cmeta.push_back(m);
      cp[(size_t)ind[r] + 1]++;
    }
    rowcell[(size_t)k] = (uint32_t)cmeta.size() - 1;
  }
  for (int n = 0; n < N; ++n) cp[n + 1] += cp[n];
  cmeta.resize(cmeta.size() + 4, 0u);
  int rc;
  int* d_rp;
  double *d_x, *d_od;
  uint32_t *d_meta, *d_rowcell, *d_cmeta;
  if ((rc = upload(h, &d_rp, rp))) return rc;
  if ((rc = upload(h, &d_x, xs))) return rc;
  if ((rc = upload(h, &d_od, ods))) return rc;
  if ((rc = upload(h, &d_meta, meta))) return rc;
  if ((rc = upload(h, &d_rowcell, rowcell))) return rc;
  if ((rc = upload(h, &d_cmeta, cmeta))) return rc;
  h->dc.rcx[a] = nullptr;
  if (h->fx) {
    std::vector<uint32_t> rcx((size_t)R + 4, 0u);
    for (int64_t k = 0; k < R; ++k) {
      const int xi = (int)(std::lower_bound(h->x_levels.begin(), h->x_levels.end(), xs[(size_t)k]) - h->x_levels.begin());
      rcx[(size_t)k] = (rowcell[(size_t)k] << 5) | (uint32_t)xi;
    }
    uint32_t* d_rcx;
    if ((rc = upload(h, &d_rcx, rcx))) return rc;
    h->dc.rcx[a] = d_rcx;
  }
  h->dc.rp[a] = d_rp;
  h->dc.x[a] = d_x;
  h->dc.od[a] = d_od;
  h->dc.meta[a] = d_meta;
  h->dc.rowcell[a] = d_rowcell;
  h->dc.cmeta[a] = d_cmeta;
  h->R[a] = R;
  return ABD_OK;
}

// Split the individuals into `want` contiguous tiles of (nearly) equal OD-row count, at most
// kTileMaxInds individuals each, and size the shared-memory staging buffers for the largest.
int get_tiling(abd_handle* h, int want, abd_handle::Tiling** out) {
  auto it = h->tilings.find(want);
  if (it == h->tilings.end()) {
    const int N = h->N;
    const std::vector<int>&rn = h->h_rp[0], &rs = h->h_rp[1];
    const double total = (double)rn[N] + (double)rs[N];
    std::vector<int> ti{0};
    int inds = 0;
    for (int n = 0; n < N; ++n) {
      // boundary before individual n if the running row count passed the next multiple of total/want
      const double before = (double)rn[n] + (double)rs[n];
      const double target = total * (double)ti.size() / (double)want;
      if (inds > 0 && (inds == kTileMaxInds || (before >= target && (int)ti.size() < want))) {
        ti.push_back(n);
        inds = 0;
      }
      ++inds;
    }
    ti.push_back(N);
    abd_handle::Tiling t;
    t.ntiles = (int)ti.size() - 1;
    const std::vector<int>&cn = h->h_cp[0], &cs = h->h_cp[1];
    auto span = [](int lo, int hi, int al) { return ((hi + al - 1) & ~(al - 1)) - (lo & ~(al - 1)); };
    for (int k = 0; k < t.ntiles; ++k) {
      const int a = ti[k], b = ti[k + 1];
      t.cap_n = std::max(t.cap_n, span(rn[a], rn[b], 2));
      t.cap_s = std::max(t.cap_s, span(rs[a], rs[b], 2));
      t.capr_n = std::max(t.capr_n, span(rn[a], rn[b], 4));
      t.capr_s = std::max(t.capr_s, span(rs[a], rs[b], 4));
      t.capk_n = std::max(t.capk_n, span(cn[a], cn[b], 4));
      t.capk_s = std::max(t.capk_s, span(cs[a], cs[b], 4));
    }
    t.cap_n = std::max(t.cap_n, 2);
    t.cap_s = std::max(t.cap_s, 2);
    t.capr_n = std::max(t.capr_n, 4);
    t.capr_s = std::max(t.capr_s, 4);
    t.capk_n = std::max(t.capk_n, 4);
    t.capk_s = std::max(t.capk_s, 4);
    t.smem = (size_t)(t.cap_n + t.cap_s) * 8 + (size_t)(t.capr_n + t.capr_s) * 4 +
             (h->fx ? (size_t)(t.capk_n + t.capk_s) * (sizeof(CellValT<true>) + 4)
                    : (size_t)(t.capk_n + t.capk_s) * (sizeof(CellValT<false>) + 4) + (size_t)(t.cap_n + t.cap_s) * 8);
    if (t.smem + 12 * 1024 > h->smem_optin)
      return fail(ABD_ERR_INVALID, "an individual tile does not fit in shared memory (too many OD rows per 128 individuals)");
    std::vector<TileDesc> desc((size_t)t.ntiles);
    for (int k = 0; k < t.ntiles; ++k) {
      const int a = ti[k], b = ti[k + 1];
      desc[(size_t)k] = TileDesc{a, b, rn[a], rn[b], rs[a], rs[b], cn[a], cn[b], cs[a], cs[b], 0, 0};
    }
    TileDesc* d_desc = nullptr;
    int rc = upload(h, &d_desc, desc);
    if (rc) return rc;
    t.d_tiles = d_desc;
    it = h->tilings.emplace(want, t).first;
  }
  *out = &it->second;
  return ABD_OK;
}

int ensure_chains(abd_handle* h, int C) {
  if (C <= 0) return fail(ABD_ERR_INVALID, "n_chains must be positive");
  if (C > 65535) return fail(ABD_ERR_INVALID, "n_chains must be <= 65535");
  if (C <= h->cap_chains) return ABD_OK;
  CU(cudaStreamSynchronize(h->stream));
  int8_t *old_i = h->d_iraw, *old_w = h->d_waner;
  const int oldC = h->cap_chains;
  const size_t gn = (size_t)h->G * h->N;
  int rc;
  int8_t *ni, *nw;
  if ((rc = dev_alloc(h, &ni, (size_t)C * gn, false))) return rc;
  if ((rc = dev_alloc(h, &nw, (size_t)C * h->N, false))) return rc;
  CU(cudaMemset(ni, 0, (size_t)C * gn));
  CU(cudaMemset(nw, 0, (size_t)C * h->N));
  if (oldC) {
    CU(cudaMemcpy(ni, old_i, (size_t)oldC * gn, cudaMemcpyDeviceToDevice));
    CU(cudaMemcpy(nw, old_w, (size_t)oldC * h->N, cudaMemcpyDeviceToDevice));
  }
  for (void* p : {(void*)old_i, (void*)old_w, (void*)h->d_theta, (void*)h->d_p, (void*)h->d_sums,
                  (void*)h->d_out, (void*)h->d_ticket, (void*)h->d_stats, (void*)h->d_aux, (void*)h->d_traj, (void*)h->d_gen})
    if (p) cudaFree(p);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  h->d_iraw = ni;
  h->d_waner = nw;
  if ((rc = dev_alloc(h, &h->d_theta, (size_t)C * 17, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_p, (size_t)C * 2, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_sums, (size_t)C * kNSums, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_out, (size_t)C * 18, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_ticket, (size_t)C, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_aux, (size_t)C * kAuxDoubles, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_traj, (size_t)C * 34, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_gen, (size_t)C + 1, false))) return rc;
  if ((rc = dev_alloc(h, &h->d_stats, (size_t)C * 2, false))) return rc;
  CU(cudaMemset(h->d_ticket, 0, (size_t)C * sizeof(unsigned)));
  CU(cudaMallocHost((void**)&h->h_pin, (size_t)C * 40 * sizeof(double)));
  h->h_pin_dev = nullptr;
  {
    void* dp = nullptr;
    if (cudaHostGetDevicePointer(&dp, h->h_pin, 0) == cudaSuccess) h->h_pin_dev = (double*)dp;
    else cudaGetLastError();
  }
  h->cap_chains = C;
  return ABD_OK;
}

// Grid plan: CTAs = tiles x chain groups.  Up to a few waves' worth of work the grid is sized
// to exactly `waves` full waves of resident CTAs (n_sms x CTAs-per-SM), so every SM gets the
// same number of equally sized tiles and there is no partial last wave; beyond that the tail is
// negligible and tiles simply hold ~`target` OD rows (a handful per thread).
void plan_grid(const abd_handle* h, int C, int ctas_per_sm, int* want_tiles, int* chains_per_cta) {
  const double rows = (double)(h->R[0] + h->R[1]);
  const int target = h->tile_rows_override > 0 ? h->tile_rows_override : 2200;
  int cpc = h->chains_per_cta_override > 0 ? h->chains_per_cta_override : (C >= 32 ? 4 : 1);
  cpc = std::min(cpc, C);
  const int groups = (C + cpc - 1) / cpc;
  const int min_tiles = (h->N + kTileMaxInds - 1) / kTileMaxInds;
  const int resident = h->n_sms * ctas_per_sm;
  int tiles = std::max(min_tiles, (int)std::ceil(rows / target));
  if ((long)tiles * groups <= 6L * resident) {
    for (int waves = 1; waves <= 6; ++waves) {
      const int t = (resident * waves) / groups;
      if (t >= min_tiles && t >= 1 && rows / t <= target * 1.02) {
        tiles = t;
        break;
      }
    }
  }
  *want_tiles = std::max(1, std::min(tiles, h->N));
  *chains_per_cta = cpc;
}

// The dynamic shared-memory limit of a kernel is a per-device function attribute shared by every
// handle of the process: only ever raise it (a small cohort must not lower it under a large one).
template <typename M, typename XT, bool TRAJ>
cudaError_t sums_smem_attr(int device, size_t smem) {
  static size_t configured[64] = {};
  const int d = (device >= 0 && device < 64) ? device : 0;
  const size_t want = std::max<size_t>(smem, 48 * 1024);
  if (want <= configured[d]) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(k_sums<M, XT, TRAJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_sums<M, XT, TRAJ>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return e;
  configured[d] = want;
  return cudaSuccess;
}

template <typename M, typename XT>
cudaError_t sums_occupancy(int device, size_t smem, int* occ) {
  cudaError_t e = sums_smem_attr<M, XT, false>(device, smem);
  if (e != cudaSuccess) return e;
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_sums<M, XT, false>, kSumsBlock, smem);
}

template <typename M, typename XT>
int launch_sums_t(abd_handle* h, const abd_handle::Tiling& tl, const SumsCfg& cfg, dim3 grid, const double* theta,
                  int theta_is_q, const int8_t* i_raw, const int8_t* waner, double* sums, const FinalizeCfg& fin,
                  const TrajCfg& traj, cudaStream_t st) {
  if (std::getenv("ABD_B200_VERBOSE")) {
    const int occ = tl.occ;
    std::fprintf(stderr, "[abd_b200] k_sums grid (%u, %u) dyn smem %zu B, occupancy %d CTAs/SM, caps rows %d/%d cells %d/%d\n",
                 grid.x, grid.y, tl.smem, occ, tl.cap_n, tl.cap_s, tl.capk_n, tl.capk_s);
  }
  const int tj = traj.n_steps > 0 ? 1 : 0;
  if (tj) CU((sums_smem_attr<M, XT, true>(h->device, tl.smem)));
  else CU((sums_smem_attr<M, XT, false>(h->device, tl.smem)));
  cudaLaunchConfig_t lc{};
  lc.gridDim = grid;
  lc.blockDim = dim3(kSumsBlock);
  lc.dynamicSmemBytes = tl.smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  const TileDesc* tiles = reinterpret_cast<const TileDesc*>(tl.d_tiles);
  const Priors* pri = h->d_priors;
  if (tj) {  // CTAs wait on one another: they must all be resident
    int occ = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_sums<M, XT, true>, kSumsBlock, tl.smem));
    if ((long)grid.x * grid.y > (long)occ * h->n_sms)
      return fail(ABD_ERR_INVALID, "abd_leapfrog_dev: the grid does not fit on the device at once");
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    lc.attrs = attr;
    lc.numAttrs = 1;
    CU(cudaLaunchKernelEx(&lc, k_sums<M, XT, true>, h->dc, tiles, cfg, theta, theta_is_q, i_raw, waner, h->d_partial,
                          h->d_ticket, sums, fin, pri, h->d_aux, traj, XchCfg{}));
  } else {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr;
    lc.numAttrs = h->use_pdl ? 1 : 0;
    CU(cudaLaunchKernelEx(&lc, k_sums<M, XT, false>, h->dc, tiles, cfg, theta, theta_is_q, i_raw, waner, h->d_partial,
                          h->d_ticket, sums, fin, pri, h->d_aux, traj, h->xch_active ? h->xch : XchCfg{}));
  }
  return ABD_OK;
}

int launch_sums(abd_handle* h, int C, const double* theta, int theta_is_q, const int8_t* i_raw,
                const int8_t* waner, double* sums, const FinalizeCfg& fin, cudaStream_t st,
                const TrajCfg* traj_in = nullptr) {
  TrajCfg traj{};
  if (traj_in) traj = *traj_in;
  // plan for the kernel's register-limited occupancy first; if the tiles that plan needs do not
  // fit that many CTAs per SM (shared memory), plan again for what actually fits
  int want, cpc, rc;
  abd_handle::Tiling* tl = nullptr;
  for (int occ = ABD_SUMS_MINB; occ >= 1; --occ) {
    plan_grid(h, C, occ, &want, &cpc);
    if (traj.n_steps > 0 && cpc != 1) {  // re-plan with one chain per CTA, exactly one wave
      cpc = 1;
      want = std::max((h->N + kTileMaxInds - 1) / kTileMaxInds, (h->n_sms * occ) / C);
      want = std::max(1, std::min(want, h->N));
    }
    if ((rc = get_tiling(h, want, &tl))) return rc;
    if (!tl->occ) {
      int o = 0;
      cudaError_t e;
      if (h->wide)
        e = h->fx ? sums_occupancy<uint64_t, uint8_t>(h->device, tl->smem, &o) : sums_occupancy<uint64_t, double>(h->device, tl->smem, &o);
      else
        e = h->fx ? sums_occupancy<uint32_t, uint8_t>(h->device, tl->smem, &o) : sums_occupancy<uint32_t, double>(h->device, tl->smem, &o);
      CU(e);
      tl->occ = std::max(o, 1);
    }
    if (tl->occ >= occ) break;
  }
  if (traj.n_steps > 0) {
    // persistent mode: one chain per CTA and every CTA resident at once
    if (cpc != 1 || (long)tl->ntiles * C > (long)h->n_sms * tl->occ)
      return fail(ABD_ERR_INVALID, "abd_leapfrog_dev: too many chains for one resident grid on this cohort; "
                                   "use abd_logp_dlogp_dev per step");
  }
  const size_t need = (size_t)C * tl->ntiles * kNSums;
  if (need > h->cap_partial) {
    CU(cudaStreamSynchronize(st));
    if (h->d_partial) cudaFree(h->d_partial);
    h->d_partial = nullptr;
    h->cap_partial = 0;
    if ((rc = dev_alloc(h, &h->d_partial, need, false))) return rc;
    h->cap_partial = need;
  }
  SumsCfg cfg{tl->ntiles, tl->cap_n, tl->cap_s, tl->capr_n, tl->capr_s, tl->capk_n, tl->capk_s, cpc, C};
  dim3 grid(tl->ntiles, (C + cpc - 1) / cpc);
  if (h->wide)
    rc = h->fx ? launch_sums_t<uint64_t, uint8_t>(h, *tl, cfg, grid, theta, theta_is_q, i_raw, waner, sums, fin, traj, st)
               : launch_sums_t<uint64_t, double>(h, *tl, cfg, grid, theta, theta_is_q, i_raw, waner, sums, fin, traj, st);
  else
    rc = h->fx ? launch_sums_t<uint32_t, uint8_t>(h, *tl, cfg, grid, theta, theta_is_q, i_raw, waner, sums, fin, traj, st)
               : launch_sums_t<uint32_t, double>(h, *tl, cfg, grid, theta, theta_is_q, i_raw, waner, sums, fin, traj, st);
  if (rc) return rc;
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int launch_gibbs(abd_handle* h, int C, const double* theta, int theta_is_q, const double* p,
                 const double* pw, int8_t* i_raw, int8_t* waner, const GibbsCfg& cfg, cudaStream_t st) {
  if (!h->gibbs_ctas) {
    int occ = 0;
    if (h->wide)
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gibbs<uint64_t>, kGibbsWarps * 32, 0));
    else
      CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gibbs<uint32_t>, kGibbsWarps * 32, 0));
    h->gibbs_ctas = h->n_sms * std::max(occ, 1);
  }
  const long items = (long)C * h->N;
  const int grid = (int)std::min<long>(h->gibbs_ctas, (items + kGibbsWarps - 1) / kGibbsWarps);
  CU(cudaMemsetAsync(h->d_queue, 0, sizeof(unsigned), st));
  if (h->wide)
    k_gibbs<uint64_t><<<grid, kGibbsWarps * 32, 0, st>>>(h->dc, h->d_order, C, theta, theta_is_q, p, pw, i_raw, waner,
                                                          h->d_queue, cfg);
  else
    k_gibbs<uint32_t><<<grid, kGibbsWarps * 32, 0, st>>>(h->dc, h->d_order, C, theta, theta_is_q, p, pw, i_raw, waner,
                                                          h->d_queue, cfg);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int set_device(const abd_handle* h) {
  CU(cudaSetDevice(h->device));
  return ABD_OK;
}

// copy optional host state into the resident buffers
// Host -> device copy of a chain-state array.  A copy-engine transfer of ~1 MB from pinned memory
// takes ~60 us on a B200 host (PCIe gen5, measured, whatever the stream count); when the source is
// pinned (page-locked, hence mapped into the device's address space under UVA) and 16-byte aligned,
// the SMs pull it themselves with 16-byte loads -- every byte of the transfer is in flight at once
// and the copy runs at the link's bandwidth.  Pageable sources take cudaMemcpyAsync.
int copy_state_h2d(abd_handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes >= (4u << 10) && h->use_pull && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, src) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
      const size_t n16 = bytes / 16, rem = bytes - n16 * 16;
      const int grid = (int)std::min<size_t>((size_t)h->n_sms * 4, (n16 + 255) / 256);
      k_pull<<<grid, 256, 0, h->stream>>>(reinterpret_cast<const uint4*>(at.devicePointer), reinterpret_cast<uint4*>(dst), n16,
                                          rem);
      CU(cudaGetLastError());
      h->launches++;
      return ABD_OK;
    }
    cudaGetLastError();  // not a CUDA-known pointer: clear the sticky-free error and copy the ordinary way
  }
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
  return ABD_OK;
}

// Device -> host copy of a chain-state array: pinned, 16-byte aligned destinations are written by
// the SMs (posted PCIe writes), anything else by the copy engine.
int copy_state_d2h(abd_handle* h, void* dst, const void* src, size_t bytes) {
  if (bytes >= (4u << 10) && h->use_pull && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, dst) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
      const size_t n16 = bytes / 16, rem = bytes - n16 * 16;
      const int grid = (int)std::min<size_t>((size_t)h->n_sms * 4, (n16 + 255) / 256);
      k_pull<<<grid, 256, 0, h->stream>>>(reinterpret_cast<const uint4*>(src), reinterpret_cast<uint4*>(at.devicePointer), n16,
                                          rem);
      CU(cudaGetLastError());
      h->launches++;
      return ABD_OK;
    }
    cudaGetLastError();
  }
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  return ABD_OK;
}

int stage_state(abd_handle* h, int C, const int8_t* i_raw, const int8_t* waner) {
  const size_t gn = (size_t)h->G * h->N;
  int rc;
  if (i_raw && (rc = copy_state_h2d(h, h->d_iraw, i_raw, (size_t)C * gn))) return rc;
  if (waner && (rc = copy_state_h2d(h, h->d_waner, waner, (size_t)C * h->N))) return rc;
  return ABD_OK;
}

#define PROLOGUE(h, C)                                         \
  if (!(h)) return fail(ABD_ERR_INVALID, "NULL handle");       \
  {                                                            \
    int rc_ = set_device(h);                                   \
    if (rc_) return rc_;                                       \
    rc_ = ensure_chains(h, C);                                 \
    if (rc_) return rc_;                                       \
  }

}  // namespace

extern "C" {

const char* abd_last_error(void) { return g_err.c_str(); }
int abd_version(void) { return 100; }

int abd_create(abd_handle** out, const abd_cohort* co, int device) {
  if (!out || !co) return fail(ABD_ERR_INVALID, "NULL argument");
  *out = nullptr;
  const int G = co->n_gaps, N = co->n_inds;
  if (G < 1 || G > ABD_MAX_GAPS) return fail(ABD_ERR_INVALID, "n_gaps must be in [1, 63]");
  if (N < 1 || N >= (1 << 26)) return fail(ABD_ERR_INVALID, "n_inds must be in [1, 2^26)");
  if (!co->vacs) return fail(ABD_ERR_INVALID, "vacs is NULL");
  // check_splits, abd.py:604-622
  if (co->n_splits < 0 || co->n_splits > 2)
    return fail(ABD_ERR_INVALID, "only implemented 1-3 time chunks (0-2 splits)");  // abd.py:882
  for (int k = 0; k < co->n_splits; ++k) {
    if (co->splits[k] < 0) return fail(ABD_ERR_INVALID, "split indexes must be positive");
    if (co->splits[k] > G) return fail(ABD_ERR_INVALID, "largest split must be less than n_gaps - 1");
  }
  if (co->n_splits == 2 && co->splits[0] == co->splits[1]) return fail(ABD_ERR_INVALID, "splits not unique");
  if (co->n_splits == 2 && co->splits[0] > co->splits[1])
    return fail(ABD_ERR_INVALID, "splits must be in ascending order");

  int ndev = 0;
  CU(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(ABD_ERR_CUDA, "no such CUDA device");
  CU(cudaSetDevice(device));

  abd_handle* h = new abd_handle();
  h->device = device;
  if (const char* e = std::getenv("ABD_B200_NO_PDL")) h->use_pdl = !(e[0] == '1');
  if (const char* e = std::getenv("ABD_B200_NO_PULL")) h->use_pull = !(e[0] == '1');
  {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) h->n_sms = v;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess) h->smem_optin = (size_t)v;
  }
  h->G = G;
  h->N = N;
  h->wide = G > 31;
  auto bail = [&](int rc) {
    abd_destroy(h);
    return rc;
  };
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(ABD_ERR_CUDA, "cudaStreamCreate failed"));

  // bit masks per individual
  std::vector<uint64_t> pcr((size_t)N, 0), vac((size_t)N, 0);
  for (int t = 0; t < G; ++t)
    for (int n = 0; n < N; ++n) {
      if (co->pcrpos && co->pcrpos[(size_t)t * N + n]) pcr[n] |= 1ull << t;
      if (co->vacs[(size_t)t * N + n]) vac[n] |= 1ull << t;
    }
  int rc;
  if (h->wide) {
    uint64_t *dp, *dv;
    if ((rc = upload(h, &dp, pcr)) || (rc = upload(h, &dv, vac))) return bail(rc);
    h->dc.pcr = dp;
    h->dc.vac = dv;
  } else {
    std::vector<uint32_t> p32(pcr.begin(), pcr.end()), v32(vac.begin(), vac.end());
    uint32_t *dp, *dv;
    if ((rc = upload(h, &dp, p32)) || (rc = upload(h, &dv, v32))) return bail(rc);
    h->dc.pcr = dp;
    h->dc.vac = dv;
  }
  h->dc.G = G;
  h->dc.N = N;
  h->dc.ind_offset = (unsigned)co->ind_offset;
  // time chunks, abd.py:865-882
  h->dc.ch.n = co->n_splits + 1;
  {
    int edges[4] = {0, 0, 0, 0};
    for (int k = 0; k < co->n_splits; ++k) edges[k + 1] = co->splits[k];
    edges[co->n_splits + 1] = G;
    for (int k = 0; k < 3; ++k) h->dc.ch.mask[k] = 0;
    for (int k = 0; k <= co->n_splits; ++k)
      for (int t = edges[k]; t < edges[k + 1]; ++t) h->dc.ch.mask[k] |= 1ull << t;
  }
  // distinct dilutions over both antigens: <= 32 of them (and no NaN) enables the factored mode
  {
    std::vector<double>& lv = h->x_levels;
    bool ok = std::getenv("ABD_B200_NO_FACTORED") == nullptr;
    for (int a = 0; a < 2 && ok; ++a) {
      const double* xa = a ? co->x_s : co->x_n;
      const int64_t Ra = a ? co->n_rows_s : co->n_rows_n;
      if (Ra > 0 && !xa) break;  // build_rows reports the NULL
      double last = std::nan("");
      for (int64_t r = 0; r < Ra && ok; ++r) {
        const double v = xa[r];
        if (v == last) continue;
        last = v;
        auto it = std::lower_bound(lv.begin(), lv.end(), v);
        if (it != lv.end() && *it == v) continue;
        if (!(v == v) || (int)lv.size() == kMaxXLevels) ok = false;
        else lv.insert(it, v);
      }
    }
    // a row packs its cell index in 27 bits
    h->fx = ok && (co->n_rows_n < ((int64_t)1 << 27)) && (co->n_rows_s < ((int64_t)1 << 27));
    if (!h->fx) lv.clear();
    std::vector<double> padded(lv);
    padded.resize(kMaxXLevels, 0.0);
    double* d_lv;
    if ((rc = upload(h, &d_lv, padded))) return bail(rc);
    h->dc.xlev = d_lv;
    h->dc.n_xlev = h->fx ? (int)lv.size() : 0;
  }
  if ((rc = build_rows(h, 0, co->n_rows_n, co->x_n, co->od_n, co->gap_n, co->ind_n))) return bail(rc);
  if ((rc = build_rows(h, 1, co->n_rows_s, co->x_s, co->od_s, co->gap_s, co->ind_s))) return bail(rc);

  {
    std::vector<int> order((size_t)N);
    for (int n = 0; n < N; ++n) order[(size_t)n] = n;
    auto nrows = [&](int n) {
      return (h->h_rp[0][n + 1] - h->h_rp[0][n]) + (h->h_rp[1][n + 1] - h->h_rp[1][n]);
    };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return nrows(a) > nrows(b); });
    if ((rc = upload(h, &h->d_order, order))) return bail(rc);
    if ((rc = dev_alloc(h, &h->d_queue, 1))) return bail(rc);
  }
  const double tn = co->total_inds > 0 ? (double)co->total_inds : (double)N;
  h->tot.rows_n = co->total_rows_n > 0 ? (double)co->total_rows_n : (double)co->n_rows_n;
  h->tot.rows_s = co->total_rows_s > 0 ? (double)co->total_rows_s : (double)co->n_rows_s;
  h->tot.bits_i = tn * G;
  h->tot.bits_w = tn;

  Priors pr = default_priors(G);
  if ((rc = dev_alloc(h, &h->d_priors, 1))) return bail(rc);
  if (cudaMemcpy(h->d_priors, &pr, sizeof(pr), cudaMemcpyHostToDevice) != cudaSuccess)
    return bail(fail(ABD_ERR_CUDA, "cudaMemcpy(priors) failed"));
  *out = h;
  return ABD_OK;
}

int abd_destroy(abd_handle* h) {
  if (!h) return ABD_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (void* p : h->xch_peer)
    if (p) cudaIpcCloseMemHandle(p);
  if (h->xch_local) cudaFree(h->xch_local);
  for (void* p : h->owned) cudaFree(p);
  for (void* p : {(void*)h->d_iraw, (void*)h->d_waner, (void*)h->d_theta, (void*)h->d_p, (void*)h->d_sums,
                  (void*)h->d_out, (void*)h->d_ticket, (void*)h->d_stats, (void*)h->d_partial, (void*)h->d_aux, (void*)h->d_traj,
                  (void*)h->d_gen})
    if (p) cudaFree(p);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return ABD_OK;
}

int abd_sizes(const abd_handle* h, int32_t* g, int32_t* n, int64_t* rs, int64_t* rn) {
  if (!h) return fail(ABD_ERR_INVALID, "NULL handle");
  if (g) *g = h->G;
  if (n) *n = h->N;
  if (rs) *rs = h->R[1];
  if (rn) *rn = h->R[0];
  return ABD_OK;
}

// B_const = 2 G N + 20 R + 8 (N + 1); B_chain = G N + N + 13*8 + 14*8
int64_t abd_algorithmic_bytes_logp(const abd_handle* h, int C) {
  if (!h) return 0;
  const int64_t G = h->G, N = h->N, R = h->R[0] + h->R[1];
  return (int64_t)C * (G * N + N + 216) + 2 * G * N + 20 * R + 8 * (N + 1);
}
int64_t abd_algorithmic_bytes_gibbs(const abd_handle* h, int C) {
  if (!h) return 0;
  const int64_t G = h->G, N = h->N, R = h->R[0] + h->R[1];
  return (int64_t)C * (2 * (G * N + N) + 120) + 2 * G * N + 20 * R + 8 * (N + 1);
}
int64_t abd_launch_count(const abd_handle* h) { return h ? h->launches : 0; }

int abd_set_tuning(abd_handle* h, int rows_per_tile, int chains_per_cta) {
  if (!h) return fail(ABD_ERR_INVALID, "NULL handle");
  if (rows_per_tile < 0 || chains_per_cta < 0) return fail(ABD_ERR_INVALID, "tuning values must be >= 0");
  h->tile_rows_override = rows_per_tile;
  h->chains_per_cta_override = chains_per_cta;
  return ABD_OK;
}

int abd_upload_state(abd_handle* h, int C, const int8_t* i_raw, const int8_t* waner) {
  PROLOGUE(h, C);
  if (!i_raw || !waner) return fail(ABD_ERR_INVALID, "NULL state");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return ABD_OK;
}

int abd_download_state(abd_handle* h, int C, int8_t* i_raw, int8_t* waner) {
  PROLOGUE(h, C);
  const size_t gn = (size_t)h->G * h->N;
  int rc;
  if (i_raw && (rc = copy_state_d2h(h, i_raw, h->d_iraw, (size_t)C * gn))) return rc;
  if (waner && (rc = copy_state_d2h(h, waner, h->d_waner, (size_t)C * h->N))) return rc;
  CU(cudaStreamSynchronize(h->stream));
  return ABD_OK;
}

int abd_state_dev(abd_handle* h, int C, int8_t** i_raw, int8_t** waner) {
  PROLOGUE(h, C);
  if (i_raw) *i_raw = h->d_iraw;
  if (waner) *waner = h->d_waner;
  return ABD_OK;
}

int abd_loglik_grad(abd_handle* h, int C, const double* theta13, const int8_t* i_raw, const int8_t* waner,
                    double* out_loglik, double* out_grad, int64_t* out_counts) {
  PROLOGUE(h, C);
  if (!theta13 || !out_loglik) return fail(ABD_ERR_INVALID, "NULL argument");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  std::memcpy(h->h_pin, theta13, (size_t)C * 13 * sizeof(double));
  CU(cudaMemcpyAsync(h->d_theta, h->h_pin, (size_t)C * 13 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  FinalizeCfg fin{1, h->tot, h->d_out, h->d_out + C};
  if ((rc = launch_sums(h, C, h->d_theta, 0, h->d_iraw, h->d_waner, h->d_sums, fin, h->stream))) return rc;
  double* hp = h->h_pin;
  CU(cudaMemcpyAsync(hp, h->d_out, (size_t)C * 14 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (out_counts)
    CU(cudaMemcpyAsync(hp + (size_t)C * 14, h->d_sums, (size_t)C * kNSums * sizeof(double),
                       cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  std::memcpy(out_loglik, hp, (size_t)C * sizeof(double));
  if (out_grad) std::memcpy(out_grad, hp + C, (size_t)C * 13 * sizeof(double));
  if (out_counts)
    for (int c = 0; c < C; ++c) {
      out_counts[2 * c] = (int64_t)hp[(size_t)C * 14 + (size_t)c * kNSums + S_KI];
      out_counts[2 * c + 1] = (int64_t)hp[(size_t)C * 14 + (size_t)c * kNSums + S_KW];
    }
  return ABD_OK;
}

int abd_logp_dlogp(abd_handle* h, int C, const double* q17, const int8_t* i_raw, const int8_t* waner,
                   double* out_logp, double* out_dlogp) {
  PROLOGUE(h, C);
  if (!q17 || !out_logp) return fail(ABD_ERR_INVALID, "NULL argument");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  // (Letting the kernel read the 17 scalars and write its results in place in the pinned staging
  // buffer over PCIe was measured: 81 us per call instead of 41 -- every CTA pays sysmem latency for
  // its parameter loads.  Small copy-engine transfers on both sides of the launch it is.)
  // The RESULTS, however, are written by the finishing warps straight into the pinned staging
  // buffer (posted PCIe writes cost the kernel nothing): no device -> host copy after the launch.
  double* ho = h->h_pin + (size_t)C * 17;
  std::memcpy(h->h_pin, q17, (size_t)C * 17 * sizeof(double));
  CU(cudaMemcpyAsync(h->d_theta, h->h_pin, (size_t)C * 17 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (h->h_pin_dev && h->use_pull) {
    double* dout = h->h_pin_dev + (size_t)C * 17;
    FinalizeCfg fin{2, h->tot, dout, dout + C};
    if ((rc = launch_sums(h, C, h->d_theta, 1, h->d_iraw, h->d_waner, h->d_sums, fin, h->stream))) return rc;
  } else {
    FinalizeCfg fin{2, h->tot, h->d_out, h->d_out + C};
    if ((rc = launch_sums(h, C, h->d_theta, 1, h->d_iraw, h->d_waner, h->d_sums, fin, h->stream))) return rc;
    CU(cudaMemcpyAsync(ho, h->d_out, (size_t)C * 18 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  CU(cudaStreamSynchronize(h->stream));
  std::memcpy(out_logp, ho, (size_t)C * sizeof(double));
  if (out_dlogp) std::memcpy(out_dlogp, ho + C, (size_t)C * 17 * sizeof(double));
  return ABD_OK;
}

static int stage_gibbs_params(abd_handle* h, int C, const double* theta13, const double* p, const double* pw) {
  double* hp = h->h_pin;
  std::memcpy(hp, theta13, (size_t)C * 13 * sizeof(double));
  for (int c = 0; c < C; ++c) {
    hp[(size_t)C * 13 + c] = p[c];
    hp[(size_t)C * 14 + c] = pw[c];
  }
  CU(cudaMemcpyAsync(h->d_theta, hp, (size_t)C * 13 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CU(cudaMemcpyAsync(h->d_p, hp + (size_t)C * 13, (size_t)C * 2 * sizeof(double), cudaMemcpyHostToDevice,
                     h->stream));
  return ABD_OK;
}

int abd_cond_logodds(abd_handle* h, int C, const double* theta13, const double* p, const double* p_w,
                     const int8_t* i_raw, const int8_t* waner, double* out_i, double* out_w) {
  PROLOGUE(h, C);
  if (!theta13 || !p || !p_w || !out_i || !out_w) return fail(ABD_ERR_INVALID, "NULL argument");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  if ((rc = stage_gibbs_params(h, C, theta13, p, p_w))) return rc;
  const size_t gn = (size_t)h->G * h->N;
  double *d_oi = nullptr, *d_ow = nullptr;
  if ((rc = dev_alloc(h, &d_oi, (size_t)C * gn, false))) return rc;
  if ((rc = dev_alloc(h, &d_ow, (size_t)C * h->N, false))) {
    cudaFree(d_oi);
    return rc;
  }
  GibbsCfg cfg{0, 0, -1, 1.0, d_oi, d_ow, nullptr};
  rc = launch_gibbs(h, C, h->d_theta, 0, h->d_p, h->d_p + C, h->d_iraw, h->d_waner, cfg, h->stream);
  cudaError_t e1 = cudaMemcpyAsync(out_i, d_oi, (size_t)C * gn * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaMemcpyAsync(out_w, d_ow, (size_t)C * h->N * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e3 = cudaStreamSynchronize(h->stream);
  cudaFree(d_oi);
  cudaFree(d_ow);
  if (rc) return rc;
  CU(e1);
  CU(e2);
  CU(e3);
  return ABD_OK;
}

int abd_gibbs_sweep(abd_handle* h, int C, const double* theta13, const double* p, const double* p_w,
                    int8_t* i_raw, int8_t* waner, uint64_t seed, uint64_t sweep_idx, int mode,
                    double transit_p, int64_t* out_stats) {
  PROLOGUE(h, C);
  if (!theta13 || !p || !p_w) return fail(ABD_ERR_INVALID, "NULL argument");
  if (mode != ABD_GIBBS_METROPOLIS && mode != ABD_GIBBS_HEATBATH) return fail(ABD_ERR_INVALID, "bad mode");
  if (!(transit_p >= 0.0 && transit_p <= 1.0)) return fail(ABD_ERR_INVALID, "transit_p must be in [0, 1]");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  if ((rc = stage_gibbs_params(h, C, theta13, p, p_w))) return rc;
  CU(cudaMemsetAsync(h->d_stats, 0, (size_t)C * 2 * sizeof(unsigned long long), h->stream));
  GibbsCfg cfg{seed, sweep_idx, mode, transit_p, nullptr, nullptr, h->d_stats};
  if ((rc = launch_gibbs(h, C, h->d_theta, 0, h->d_p, h->d_p + C, h->d_iraw, h->d_waner, cfg, h->stream)))
    return rc;
  const size_t gn = (size_t)h->G * h->N;
  if (i_raw && (rc = copy_state_d2h(h, i_raw, h->d_iraw, (size_t)C * gn))) return rc;
  if (waner && (rc = copy_state_d2h(h, waner, h->d_waner, (size_t)C * h->N))) return rc;
  if (out_stats)
    CU(cudaMemcpyAsync(h->h_pin, h->d_stats, (size_t)C * 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                       h->stream));
  CU(cudaStreamSynchronize(h->stream));
  if (out_stats) std::memcpy(out_stats, h->h_pin, (size_t)C * 2 * sizeof(int64_t));
  return ABD_OK;
}

int abd_deterministics(abd_handle* h, int C, const double* theta13, const int8_t* i_raw, const int8_t* waner,
                       int8_t* out_i, double* out_mu_n, double* out_mu_s) {
  PROLOGUE(h, C);
  if (!theta13) return fail(ABD_ERR_INVALID, "NULL argument");
  int rc = stage_state(h, C, i_raw, waner);
  if (rc) return rc;
  std::memcpy(h->h_pin, theta13, (size_t)C * 13 * sizeof(double));
  CU(cudaMemcpyAsync(h->d_theta, h->h_pin, (size_t)C * 13 * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  const size_t gn = (size_t)h->G * h->N * C;
  int8_t* d_i = nullptr;
  double *d_n = nullptr, *d_s = nullptr;
  auto cleanup = [&]() {
    if (d_i) cudaFree(d_i);
    if (d_n) cudaFree(d_n);
    if (d_s) cudaFree(d_s);
  };
  if (out_i && (rc = dev_alloc(h, &d_i, gn, false))) return rc;
  if (out_mu_n && (rc = dev_alloc(h, &d_n, gn, false))) {
    cleanup();
    return rc;
  }
  if (out_mu_s && (rc = dev_alloc(h, &d_s, gn, false))) {
    cleanup();
    return rc;
  }
  rc = abd_deterministics_dev(h, C, h->d_theta, h->d_iraw, h->d_waner, d_i, d_n, d_s, h->stream);
  cudaError_t e = cudaSuccess;
  if (!rc && out_i) e = cudaMemcpyAsync(out_i, d_i, gn, cudaMemcpyDeviceToHost, h->stream);
  if (!rc && e == cudaSuccess && out_mu_n)
    e = cudaMemcpyAsync(out_mu_n, d_n, gn * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (!rc && e == cudaSuccess && out_mu_s)
    e = cudaMemcpyAsync(out_mu_s, d_s, gn * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  cudaError_t e2 = cudaStreamSynchronize(h->stream);
  cleanup();
  if (rc) return rc;
  CU(e);
  CU(e2);
  return ABD_OK;
}

// ---- device-pointer variants ----------------------------------------------------------------
int abd_sums_dev(abd_handle* h, int C, const double* theta, int theta_is_q17, const int8_t* i_raw,
                 const int8_t* waner, double* sums, void* stream) {
  PROLOGUE(h, C);
  if (!theta || !i_raw || !waner || !sums) return fail(ABD_ERR_INVALID, "NULL argument");
  FinalizeCfg fin{0, h->tot, nullptr, nullptr};
  return launch_sums(h, C, theta, theta_is_q17, i_raw, waner, sums, fin, (cudaStream_t)stream);
}

int abd_finalize_loglik_dev(abd_handle* h, int C, const double* theta13, const double* sums, double* out_loglik,
                            double* out_grad, void* stream) {
  PROLOGUE(h, C);
  if (!theta13 || !sums || !out_loglik) return fail(ABD_ERR_INVALID, "NULL argument");
  FinalizeCfg fin{1, h->tot, out_loglik, out_grad};
  k_finalize<<<(C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(C, theta13, sums, fin, h->d_priors);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int abd_finalize_logp_dev(abd_handle* h, int C, const double* q17, const double* sums, double* out_logp,
                          double* out_dlogp, void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !sums || !out_logp) return fail(ABD_ERR_INVALID, "NULL argument");
  FinalizeCfg fin{2, h->tot, out_logp, out_dlogp};
  k_finalize<<<(C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(C, q17, sums, fin, h->d_priors);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int abd_loglik_grad_dev(abd_handle* h, int C, const double* theta13, const int8_t* i_raw, const int8_t* waner,
                        double* out_loglik, double* out_grad, void* stream) {
  PROLOGUE(h, C);
  if (!theta13 || !i_raw || !waner || !out_loglik) return fail(ABD_ERR_INVALID, "NULL argument");
  FinalizeCfg fin{1, h->tot, out_loglik, out_grad};
  return launch_sums(h, C, theta13, 0, i_raw, waner, h->d_sums, fin, (cudaStream_t)stream);
}

int abd_logp_dlogp_dev(abd_handle* h, int C, const double* q17, const int8_t* i_raw, const int8_t* waner,
                       double* out_logp, double* out_dlogp, void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !i_raw || !waner || !out_logp) return fail(ABD_ERR_INVALID, "NULL argument");
  FinalizeCfg fin{2, h->tot, out_logp, out_dlogp};
  return launch_sums(h, C, q17, 1, i_raw, waner, h->d_sums, fin, (cudaStream_t)stream);
}

int abd_gibbs_sweep_dev(abd_handle* h, int C, const double* theta, int theta_is_q17, const double* p,
                        const double* p_w, int8_t* i_raw, int8_t* waner, uint64_t seed, uint64_t sweep_idx,
                        int mode, double transit_p, unsigned long long* stats, void* stream) {
  PROLOGUE(h, C);
  if (!theta || !i_raw || !waner) return fail(ABD_ERR_INVALID, "NULL argument");
  if (!theta_is_q17 && (!p || !p_w)) return fail(ABD_ERR_INVALID, "p / p_w required with theta13");
  if (mode != ABD_GIBBS_METROPOLIS && mode != ABD_GIBBS_HEATBATH) return fail(ABD_ERR_INVALID, "bad mode");
  GibbsCfg cfg{seed, sweep_idx, mode, transit_p, nullptr, nullptr, stats};
  return launch_gibbs(h, C, theta, theta_is_q17, p, p_w, i_raw, waner, cfg, (cudaStream_t)stream);
}

int abd_deterministics_dev(abd_handle* h, int C, const double* theta13, const int8_t* i_raw, const int8_t* waner,
                           int8_t* out_i, double* out_mu_n, double* out_mu_s, void* stream) {
  PROLOGUE(h, C);
  if (!theta13 || !i_raw || !waner) return fail(ABD_ERR_INVALID, "NULL argument");
  dim3 grid((h->N + 127) / 128, C);
  if (h->wide)
    k_determ<uint64_t><<<grid, 128, 0, (cudaStream_t)stream>>>(h->dc, theta13, i_raw, waner, out_i, out_mu_n, out_mu_s);
  else
    k_determ<uint32_t><<<grid, 128, 0, (cudaStream_t)stream>>>(h->dc, theta13, i_raw, waner, out_i, out_mu_n, out_mu_s);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int abd_leapfrog_dev(abd_handle* h, int C, int n_steps, double* q17, double* p17, double* grad17, double* logp,
                     const double* eps, const double* inv_mass, const int8_t* i_raw, const int8_t* waner, void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !p17 || !grad17 || !logp || !eps || !inv_mass || !i_raw || !waner) return fail(ABD_ERR_INVALID, "NULL argument");
  if (n_steps < 1 || n_steps > 4096) return fail(ABD_ERR_INVALID, "n_steps must be in [1, 4096]");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemsetAsync(h->d_gen, 0, ((size_t)C + 1) * sizeof(unsigned), st));
  TrajCfg traj{n_steps, q17, p17, grad17, logp, eps, inv_mass, h->d_traj, h->d_gen, h->d_gen + C};
  FinalizeCfg fin{2, h->tot, nullptr, nullptr};
  return launch_sums(h, C, q17, 1, i_raw, waner, nullptr, fin, st, &traj);
}

int abd_leapfrog_status(abd_handle* h, int C) {
  PROLOGUE(h, C);
  unsigned err = 0;
  CU(cudaMemcpy(&err, h->d_gen + C, sizeof(unsigned), cudaMemcpyDeviceToHost));
  return err ? fail(ABD_ERR_CUDA, "abd_leapfrog_dev: a CTA timed out waiting for its chain (grid not co-resident?)") : ABD_OK;
}

int abd_hmc_begin_dev(abd_handle* h, int C, const double* q17, const double* grad17, const double* logp,
                      const double* linv_t, uint64_t seed, uint64_t iter, double* qw, double* pw, double* gw, double* h0,
                      void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !grad17 || !logp || !linv_t || !qw || !pw || !gw || !h0) return fail(ABD_ERR_INVALID, "NULL argument");
  k_hmc_begin<<<(C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(C, q17, grad17, logp, linv_t, seed, iter, qw, pw, gw, h0);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int abd_hmc_end_dev(abd_handle* h, int C, double* q17, double* grad17, double* logp, const double* qw, const double* pw,
                    const double* gw, const double* lpw, const double* inv_mass, const double* h0, uint64_t seed,
                    uint64_t iter, double* accept_out, double* da, double* eps, int adapt, double target_accept,
                    void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !grad17 || !logp || !qw || !pw || !gw || !lpw || !inv_mass || !h0 || !accept_out)
    return fail(ABD_ERR_INVALID, "NULL argument");
  if (adapt && (!da || !eps)) return fail(ABD_ERR_INVALID, "adapt needs the dual-averaging state and eps");
  k_hmc_end<<<(C + 3) / 4, 128, 0, (cudaStream_t)stream>>>(C, q17, grad17, logp, qw, pw, gw, lpw, inv_mass, h0, seed, iter,
                                                            accept_out, da, eps, adapt, target_accept);
  CU(cudaGetLastError());
  h->launches++;
  return ABD_OK;
}

int abd_xch_alloc(abd_handle* h, int world, int rank, int max_chains, void* out_ipc_handle) {
  if (!h) return fail(ABD_ERR_INVALID, "NULL handle");
  if (world < 2 || world > kMaxPeers || rank < 0 || rank >= world || max_chains < 1 || !out_ipc_handle)
    return fail(ABD_ERR_INVALID, "abd_xch_alloc: need 2 <= world <= 8, 0 <= rank < world, max_chains >= 1");
  if (world * kNSums > kSumsBlock) return fail(ABD_ERR_INVALID, "world too large");
  int rc = set_device(h);
  if (rc) return rc;
  if (h->xch_local) return fail(ABD_ERR_INVALID, "abd_xch_alloc: already allocated");
  const size_t nd = (size_t)2 * world * max_chains * kNSums, nf = (size_t)2 * world * max_chains;
  const size_t bytes = nd * sizeof(double) + nf * sizeof(unsigned long long);
  CU(cudaMalloc(&h->xch_local, bytes));
  CU(cudaMemset(h->xch_local, 0, bytes));
  unsigned* seq;
  if ((rc = dev_alloc(h, &seq, (size_t)max_chains + 1))) return rc;
  CU(cudaMemset(seq, 0, ((size_t)max_chains + 1) * sizeof(unsigned)));
  h->xch = XchCfg{};
  h->xch.world = world;
  h->xch.rank = rank;
  h->xch.cmax = max_chains;
  h->xch.seq = seq;
  h->xch.err = seq + max_chains;
  cudaIpcMemHandle_t ipc;
  CU(cudaIpcGetMemHandle(&ipc, h->xch_local));
  static_assert(sizeof(ipc) == ABD_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(out_ipc_handle, &ipc, sizeof(ipc));
  return ABD_OK;
}

int abd_xch_connect(abd_handle* h, const void* all_ipc_handles) {
  if (!h || !all_ipc_handles) return fail(ABD_ERR_INVALID, "NULL argument");
  if (!h->xch_local) return fail(ABD_ERR_INVALID, "abd_xch_connect: call abd_xch_alloc first");
  int rc = set_device(h);
  if (rc) return rc;
  const int world = h->xch.world;
  const size_t nd = (size_t)2 * world * h->xch.cmax * kNSums;
  for (int r = 0; r < world; ++r) {
    void* base = h->xch_local;
    if (r != h->xch.rank) {
      cudaIpcMemHandle_t ipc;
      std::memcpy(&ipc, (const char*)all_ipc_handles + (size_t)r * sizeof(ipc), sizeof(ipc));
      CU(cudaIpcOpenMemHandle(&base, ipc, cudaIpcMemLazyEnablePeerAccess));
      h->xch_peer[r] = base;
    }
    h->xch.data[r] = (double*)base;
    h->xch.flag[r] = (unsigned long long*)((double*)base + nd);
  }
  return ABD_OK;
}

int abd_logp_dlogp_sharded_dev(abd_handle* h, int C, const double* q17, const int8_t* i_raw, const int8_t* waner,
                               double* out_logp, double* out_dlogp, void* stream) {
  PROLOGUE(h, C);
  if (!q17 || !i_raw || !waner || !out_logp) return fail(ABD_ERR_INVALID, "NULL argument");
  if (!h->xch_local || !h->xch.data[h->xch.world - 1] || !h->xch.data[0])
    return fail(ABD_ERR_INVALID, "abd_logp_dlogp_sharded_dev: call abd_xch_alloc and abd_xch_connect first");
  if (C > h->xch.cmax) return fail(ABD_ERR_INVALID, "more chains than abd_xch_alloc reserved");
  FinalizeCfg fin{2, h->tot, out_logp, out_dlogp};
  h->xch_active = true;
  const int rc = launch_sums(h, C, q17, 1, i_raw, waner, h->d_sums, fin, (cudaStream_t)stream);
  h->xch_active = false;
  return rc;
}

int abd_xch_status(abd_handle* h) {
  if (!h || !h->xch_local) return fail(ABD_ERR_INVALID, "no exchange buffer");
  int rc = set_device(h);
  if (rc) return rc;
  unsigned err = 0;
  CU(cudaMemcpy(&err, h->xch.err, sizeof(unsigned), cudaMemcpyDeviceToHost));
  return err ? fail(ABD_ERR_CUDA, "a peer's contribution did not arrive (timeout)") : ABD_OK;
}

int abd_debug_fast_math(int device, int64_t n, const double* z, double* out_exp, double* out_rcp) {
  if (n < 0 || !z || !out_exp || !out_rcp) return fail(ABD_ERR_INVALID, "bad argument");
  CU(cudaSetDevice(device));
  double *dz = nullptr, *de = nullptr, *dr = nullptr;
  const size_t bytes = (size_t)std::max<int64_t>(n, 1) * sizeof(double);
  CU(cudaMalloc((void**)&dz, bytes));
  CU(cudaMalloc((void**)&de, bytes));
  CU(cudaMalloc((void**)&dr, bytes));
  cudaError_t e = cudaMemcpy(dz, z, (size_t)n * sizeof(double), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_debug_fast_math<<<296, 256>>>(n, dz, de, dr);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_exp, de, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(out_rcp, dr, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(dz);
  cudaFree(de);
  cudaFree(dr);
  CU(e);
  return ABD_OK;
}

#ifdef ABD_PHASE_TIMING
int abd_debug_spans(unsigned long long* out, int reset) {
  if (reset) {
    unsigned long long init[256][2];
    for (auto& r : init) { r[0] = ~0ull; r[1] = 0; }
    unsigned z = 0;
    CU(cudaMemcpyToSymbol(g_span, init, sizeof(init)));
    CU(cudaMemcpyToSymbol(g_span_idx, &z, sizeof(z)));
    CU(cudaMemcpyToSymbol(g_span_done, &z, sizeof(z)));
    return ABD_OK;
  }
  CU(cudaMemcpyFromSymbol(out, g_span, sizeof(unsigned long long) * 512));
  return ABD_OK;
}
int abd_debug_phase_times(unsigned long long* out, int n_ctas) {
  CU(cudaMemcpyFromSymbol(out, g_phase, sizeof(unsigned long long) * 16 * (size_t)n_ctas));
  return ABD_OK;
}
#endif

}  // extern "C"
