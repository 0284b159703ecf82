// Types and helpers shared by the kernels of libabd_b200.so and the host side that launches them
// (device view of the cohort, launch configurations, warp-level power tables).
#pragma once
#include "abd_device.cuh"

namespace {
using namespace abd;

// ------------------------------------------------------------------------------------------
// device-side view of the cohort
// ------------------------------------------------------------------------------------------
struct DevCohort {
  int G, N;
  unsigned ind_offset;       // global index of this shard's first individual (RNG streams)
  unsigned chain_offset;     // global index of this handle's first chain (RNG streams; abd_set_chain_offset)
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
  const int* rowperm[2];       // per (sorted) row: its index among the antigen's rows as the caller passed them
  Chunks ch;
};

constexpr int kSumsBlock = 256;
constexpr int kSumsWarps = kSumsBlock / 32;
constexpr int kAuxDoubles = 128;  // 17 x PriorPre (7) + LikPre (6), padded
constexpr int kTileMaxInds = 128;
#ifndef ABD_GIBBS_WARPS
#define ABD_GIBBS_WARPS 8
#endif
#ifndef ABD_GIBBS_MINB
#define ABD_GIBBS_MINB 4
#endif
constexpr int kGibbsWarps = ABD_GIBBS_WARPS;  // warps per CTA of the Gibbs kernels; ABD_GIBBS_MINB CTAs per SM

// per-individual state staged in shared memory: constrained infections, vaccinations, waner
// (packed in the top bit of the vaccination mask; usable gaps <= width - 1)
template <typename M>
struct IndState {
  M inf, vacw;
};
template <typename M>
__device__ __forceinline__ M top_bit() { return (M)1 << (sizeof(M) * 8 - 1); }

// Packed resident chain state, one entry per (chain, individual), kept beside the int8 boundary
// arrays by every library call that writes the resident state (abd_upload_state, the Gibbs sweeps):
//   rw  = the individual's i_raw column as a bit mask over gaps | ab_s_waner in the top bit
//   inf = constrain(raw, pcr, chunks), the constrained infections (abd.py:640-667)
// so that an evaluation reads 8 (16) bytes per individual and chain instead of G + 1 strided bytes
// and does not repeat the integer prologue: the binary state only changes once per Gibbs sweep.
template <typename M>
struct __align__(2 * sizeof(M)) PackedState {
  M rw, inf;
};

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

// The same, interleaved: tab[k] = {rho^k, k rho^(k-1)}.
__device__ __forceinline__ void fill_pow2_warp(double rho, int G, int lane, double2* tab) {
  double v = rho;  // inclusive prefix product: rho^(lane+1)
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const double u = __shfl_up_sync(0xffffffffu, v, off);
    if (lane >= off) v *= u;
  }
  const double r32 = __shfl_sync(0xffffffffu, v, 31);  // rho^32
  double prev = __shfl_up_sync(0xffffffffu, v, 1);      // rho^lane
  if (lane == 0) {
    prev = 1.0;
    tab[0] = make_double2(1.0, 0.0);
  }
  for (int blk = 0; blk * 32 < G; ++blk) {
    const int k = blk * 32 + lane + 1;
    const double scale = blk ? r32 : 1.0;  // G <= 63: at most two blocks
    if (k < G) tab[k] = make_double2(v * scale, (double)k * prev * scale);
  }
}

}  // namespace
