// k_nuts_begin / k_nuts_leaf / k_nuts_end -- the No-U-Turn sampler's tree bookkeeping on the device,
// one warp per chain (lane k < 17 owns component k of every vector).
//
// What pm.sample runs for the 17 continuous value variables (abd.py:922) is NUTS: multinomial sampling
// inside sub-trees, biased progressive sampling between them, the generalised U-turn criterion on every
// balanced sub-tree (Betancourt 2017; as Stan / PyMC implement it).  The chains of a batch double in
// LOCKSTEP: the host enqueues, for tree depth j = 0, 1, ..., the 2^j leaves of the new sub-tree -- each leaf
// one single-step leapfrog launch (k_sums in trajectory mode, all chains at once) followed by one k_nuts_leaf
// launch that folds the new state into every chain's tree -- and reads ONE word per depth (does any chain
// still want to double?).  A chain whose tree has terminated, or whose current sub-tree turned / diverged,
// is masked: its work vectors keep being integrated, nothing is read from them.
//
// Per-chain state (doubles): 12 + 2 D vectors of 17 (D = maximum depth) and 16 scalars, see NutsLayout.
#pragma once
#include "abd_kernels_common.cuh"

namespace {
using namespace abd;

constexpr int kNutsMaxDepth = 10;

struct NutsLayout {
  int D;  // checkpoint slots
  __host__ __device__ int vec(int k) const { return k * 17; }
  // vectors
  enum { QL = 0, PL, GL, QR, PR, GR, QPROP, GPROP, RHO, SQ, SG, SRHO, NVEC };
  __host__ __device__ int pck(int i) const { return (NVEC + i) * 17; }
  __host__ __device__ int rck(int i) const { return (NVEC + D + i) * 17; }
  __host__ __device__ int scal() const { return (NVEC + 2 * D) * 17; }
  // scalars (offsets from scal())
  enum { LPPROP = 0, LOGW, SLP, SLOGW, H0, SUMACC, NACC, ACTIVE, SSTOP, DIVERGED, DEPTH, FWD, NSCAL = 16 };
  __host__ __device__ int size() const { return scal() + NSCAL; }
};

// dot over the 17 components held by lanes 0..16 (every lane gets the result)
__device__ __forceinline__ double dot17(double a, double b, int lane) {
  double v = lane < 17 ? a * b : 0.0;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// (Sigma x)[lane] for a symmetric 17 x 17 Sigma in shared / global memory
__device__ __forceinline__ double matvec17(const double* __restrict__ S, double x, int lane) {
  double v = 0.0;
  for (int j = 0; j < 17; ++j) {
    const double xj = __shfl_sync(0xffffffffu, x, j);
    if (lane < 17) v = fma(S[lane * 17 + j], xj, v);
  }
  return v;
}
// the generalised U-turn criterion for a span with end momenta p_a, p_b and momentum sum rho
__device__ __forceinline__ bool nuts_turning(const double* __restrict__ S, double pa, double pb, double rho, int lane) {
  const double adj = rho - 0.5 * (pa + pb);
  const double sa = matvec17(S, adj, lane);
  return dot17(pa, sa, lane) <= 0.0 || dot17(pb, sa, lane) <= 0.0;
}
__device__ __forceinline__ double nuts_logaddexp(double a, double b) {
  const double m = fmax(a, b);
  if (!(m > -INFINITY)) return -INFINITY;
  return m + log(exp(a - m) + exp(b - m));
}
__device__ __forceinline__ double nuts_uniform(uint64_t seed, uint64_t iter, unsigned chain, unsigned tag) {
  const uint4 r = philox4x32_10(make_uint4(tag, chain, (uint32_t)iter, 0x4e555453u),
                                make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iter >> 32)));
  return u01(r.x) - 2.3283064365386963e-10;  // [0, 1)
}

// start the sub-tree of depth j for this chain: direction, empty sub-tree, work vectors = the end to extend
__device__ __forceinline__ void nuts_open_subtree(double* st, const NutsLayout& L, int j, int c, int lane, uint64_t seed,
                                                  uint64_t iter, unsigned chain_global, const double* __restrict__ eps,
                                                  double* qw, double* pw, double* gw, double* eps_signed) {
  double* sc = st + L.scal();
  const bool fwd = nuts_uniform(seed, iter, chain_global, 0x10000u + (unsigned)j) < 0.5;
  if (lane < 17) {
    const int e = fwd ? NutsLayout::QR : NutsLayout::QL;  // QL/PL/GL and QR/PR/GR are consecutive
    qw[(size_t)c * 17 + lane] = st[L.vec(e) + lane];
    pw[(size_t)c * 17 + lane] = st[L.vec(e + 1) + lane];
    gw[(size_t)c * 17 + lane] = st[L.vec(e + 2) + lane];
    st[L.vec(NutsLayout::SRHO) + lane] = 0.0;
  }
  if (lane == 0) {
    sc[NutsLayout::FWD] = fwd ? 1.0 : 0.0;
    sc[NutsLayout::SLOGW] = -INFINITY;
    sc[NutsLayout::SLP] = sc[NutsLayout::LPPROP];
    sc[NutsLayout::SSTOP] = 0.0;
    eps_signed[c] = fwd ? eps[c] : -eps[c];
  }
}

__global__ void __launch_bounds__(128)
k_nuts_begin(const int C, const NutsLayout L, const double* __restrict__ q, const double* __restrict__ grad,
             const double* __restrict__ logp, const double* __restrict__ linv_t, const double* __restrict__ eps,
             const uint64_t seed, const uint64_t iter, const unsigned chain_offset, double* __restrict__ state,
             double* __restrict__ qw, double* __restrict__ pw, double* __restrict__ gw, double* __restrict__ eps_signed,
             int* __restrict__ any_active) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c == 0 && lane <= L.D) any_active[lane] = (lane == 0);  // max_depth + 1 words; [j]: does any chain want depth j?
  if (c >= C) return;
  double* st = state + (size_t)c * L.size();
  double* sc = st + L.scal();
  double z = 0.0;
  if (lane < 17) {  // Box-Muller on two of the four Philox words (the stream of k_hmc_begin)
    const uint4 r = philox4x32_10(make_uint4((uint32_t)lane, (uint32_t)c + chain_offset, (uint32_t)iter, 0x484d4331u),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(iter >> 32)));
    z = sqrt(-2.0 * log(u01(r.x))) * cospi(2.0 * u01(r.y));
  }
  const double p0 = matvec17(linv_t, z, lane);  // p = (L^T)^-1 z ~ N(0, M)
  const double kin = dot17(z, z, lane);
  if (lane < 17) {
    const double qv = q[(size_t)c * 17 + lane], gv = grad[(size_t)c * 17 + lane];
    st[L.vec(NutsLayout::QL) + lane] = qv, st[L.vec(NutsLayout::QR) + lane] = qv, st[L.vec(NutsLayout::QPROP) + lane] = qv;
    st[L.vec(NutsLayout::PL) + lane] = p0, st[L.vec(NutsLayout::PR) + lane] = p0, st[L.vec(NutsLayout::RHO) + lane] = p0;
    st[L.vec(NutsLayout::GL) + lane] = gv, st[L.vec(NutsLayout::GR) + lane] = gv, st[L.vec(NutsLayout::GPROP) + lane] = gv;
  }
  if (lane == 0) {
    sc[NutsLayout::LPPROP] = logp[c];
    sc[NutsLayout::LOGW] = 0.0;
    sc[NutsLayout::H0] = -logp[c] + 0.5 * kin;
    sc[NutsLayout::SUMACC] = 0.0, sc[NutsLayout::NACC] = 0.0;
    sc[NutsLayout::ACTIVE] = 1.0, sc[NutsLayout::DIVERGED] = 0.0, sc[NutsLayout::DEPTH] = 0.0;
  }
  __syncwarp();
  nuts_open_subtree(st, L, 0, c, lane, seed, iter, (unsigned)c + chain_offset, eps, qw, pw, gw, eps_signed);
}

// What one warp does for its chain once leaf n of the depth-j sub-tree has been integrated into (qw, pw, gw, lpw).
// Called by k_nuts_leaf (its own launch) and by the finishing warp of a single-step leapfrog launch of k_sums (the
// leaf and its bookkeeping are then ONE launch).  s_im: the 17 x 17 metric in shared memory.
struct NutsLeafArgs {
  double* state;            // nullptr: no tree bookkeeping
  NutsLayout L;
  int j, n, max_depth;
  const double* eps;        // unsigned per-chain step sizes
  uint64_t seed, iter;
  unsigned chain_offset;
  double* eps_signed;
  int* any_active;
};
__device__ __noinline__ void nuts_leaf_chain(const NutsLeafArgs a, const int c, const int lane, double* qw, double* pw,
                                             double* gw, const double* lpw, const double* s_im) {
  const NutsLayout L = a.L;
  const int j = a.j, n = a.n, max_depth = a.max_depth;
  const uint64_t seed = a.seed, iter = a.iter;
  const double* eps = a.eps;
  double* eps_signed = a.eps_signed;
  int* any_active = a.any_active;
  double* st = a.state + (size_t)c * L.size();
  double* sc = st + L.scal();
  const unsigned cg = (unsigned)c + a.chain_offset;
  const bool active = sc[NutsLayout::ACTIVE] != 0.0;
  bool s_stop = sc[NutsLayout::SSTOP] != 0.0;
  const bool run = active && !s_stop;
  const size_t o = (size_t)c * 17 + lane;
  if (run) {
    const double pn = lane < 17 ? pw[o] : 0.0, qn = lane < 17 ? qw[o] : 0.0, gn = lane < 17 ? gw[o] : 0.0;
    const double lpn = lpw[c];
    const double kin = dot17(pn, matvec17(s_im, pn, lane), lane);
    double dh = sc[NutsLayout::H0] - (-lpn + 0.5 * kin);
    if (!isfinite(dh)) dh = -INFINITY;
    const bool div = fabs(dh) > 1000.0;  // PyMC's divergence threshold on the energy error
    const bool ok = !div;
    double s_log_w = sc[NutsLayout::SLOGW];
    const double new_w = nuts_logaddexp(s_log_w, dh);
    const bool take = ok && nuts_uniform(seed, iter, cg, 0x20000u + ((unsigned)j << 11) + (unsigned)n) < exp(dh - new_w);
    double s_rho = lane < 17 ? st[L.vec(NutsLayout::SRHO) + lane] : 0.0;
    if (ok) s_rho += pn;
    if (lane < 17) {
      if (take) st[L.vec(NutsLayout::SQ) + lane] = qn, st[L.vec(NutsLayout::SG) + lane] = gn;
      st[L.vec(NutsLayout::SRHO) + lane] = s_rho;
    }
    if (ok) s_log_w = new_w;
    if (div) s_stop = true;
    // sub-tree U-turn checks with O(depth) checkpoints: leaf n (odd) closes the balanced sub-trees that start at the
    // leaves whose checkpoints sit at idx_min .. idx_max; an even leaf opens one at idx_max
    if (j > 0 && ok) {
      const int idx_max = __popc((unsigned)n >> 1);
      if ((n & 1) == 0) {
        if (lane < 17) st[L.pck(idx_max) + lane] = pn, st[L.rck(idx_max) + lane] = s_rho;
      } else {
        const int trailing = __ffs(~n) - 1;  // number of trailing one bits of n
        const int idx_min = idx_max - trailing + 1;
        for (int i = idx_max; i >= idx_min; --i) {
          const double pc = lane < 17 ? st[L.pck(i) + lane] : 0.0, rc = lane < 17 ? st[L.rck(i) + lane] : 0.0;
          if (nuts_turning(s_im, pc, pn, s_rho - rc + pc, lane)) s_stop = true;
        }
      }
    }
    if (lane == 0) {
      sc[NutsLayout::SUMACC] += exp(fmin(dh, 0.0));
      sc[NutsLayout::NACC] += 1.0;
      if (take) sc[NutsLayout::SLP] = lpn;
      sc[NutsLayout::SLOGW] = s_log_w;
      sc[NutsLayout::SSTOP] = s_stop ? 1.0 : 0.0;
      if (div) sc[NutsLayout::DIVERGED] = 1.0;
    }
    __syncwarp();
  }
  if (n != (1 << j) - 1) return;
  // ---- the sub-tree is complete: merge it into the tree (chains that neither turned inside it nor diverged) ----
  bool still = false;
  if (active) {
    const bool good = !s_stop;
    const bool fwd = sc[NutsLayout::FWD] != 0.0;
    const double log_w = sc[NutsLayout::LOGW], s_log_w = sc[NutsLayout::SLOGW];
    double rho = lane < 17 ? st[L.vec(NutsLayout::RHO) + lane] : 0.0;
    if (good) {
      const bool take = nuts_uniform(seed, iter, cg, 0x30000u + (unsigned)j) < exp(fmin(s_log_w - log_w, 0.0));
      rho += lane < 17 ? st[L.vec(NutsLayout::SRHO) + lane] : 0.0;
      if (lane < 17) {
        if (take) {
          st[L.vec(NutsLayout::QPROP) + lane] = st[L.vec(NutsLayout::SQ) + lane];
          st[L.vec(NutsLayout::GPROP) + lane] = st[L.vec(NutsLayout::SG) + lane];
        }
        st[L.vec(NutsLayout::RHO) + lane] = rho;
        const int e = fwd ? NutsLayout::QR : NutsLayout::QL;  // the extended end is the last leaf
        st[L.vec(e) + lane] = qw[o], st[L.vec(e + 1) + lane] = pw[o], st[L.vec(e + 2) + lane] = gw[o];
      }
      __syncwarp();
      if (lane == 0) {
        if (take) sc[NutsLayout::LPPROP] = sc[NutsLayout::SLP];
        sc[NutsLayout::LOGW] = nuts_logaddexp(log_w, s_log_w);
      }
      const double pl = lane < 17 ? st[L.vec(NutsLayout::PL) + lane] : 0.0, pr = lane < 17 ? st[L.vec(NutsLayout::PR) + lane] : 0.0;
      still = !nuts_turning(s_im, pl, pr, rho, lane) && (j + 1 < max_depth);
    }
    if (lane == 0) {
      sc[NutsLayout::DEPTH] += 1.0;
      sc[NutsLayout::ACTIVE] = still ? 1.0 : 0.0;
    }
    __syncwarp();
  }
  if (still) {
    nuts_open_subtree(st, L, j + 1, c, lane, seed, iter, cg, eps, qw, pw, gw, eps_signed);
    if (lane == 0) atomicOr(&any_active[j + 1], 1);
  }
}

// Leaf n of the depth-j sub-tree has just been integrated into (qw, pw, gw, lpw): the bookkeeping as its own launch.
__global__ void __launch_bounds__(128)
k_nuts_leaf(const int C, const NutsLeafArgs a, double* qw, double* pw, double* gw, const double* lpw,
            const double* __restrict__ inv_mass) {
  __shared__ double s_im[17 * 17];
  // launched with programmatic stream serialisation: the metric (written by the host only) is staged while the leapfrog
  // launch this leaf belongs to is still finishing; everything else waits for it; the next leapfrog launch may start its
  // own prologue (immutable cohort data) right away
  griddep_launch_dependents();
  for (int k = threadIdx.x; k < 17 * 17; k += blockDim.x) s_im[k] = inv_mass[k];
  __syncthreads();
  griddep_wait();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  nuts_leaf_chain(a, c, lane, qw, pw, gw, lpw, s_im);
}

// The transition's end: the chain moves to the tree's proposal; sample stats; dual averaging of the step size.
__global__ void __launch_bounds__(128)
k_nuts_end(const int C, const NutsLayout L, double* __restrict__ q, double* __restrict__ grad, double* __restrict__ logp,
           const double* __restrict__ state, double* __restrict__ accept_out, double* __restrict__ depth_out,
           double* __restrict__ diverged_out, double* __restrict__ da, double* __restrict__ eps, const int adapt,
           const double target) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (c >= C) return;
  const double* st = state + (size_t)c * L.size();
  const double* sc = st + L.scal();
  if (lane < 17) {
    q[(size_t)c * 17 + lane] = st[L.vec(NutsLayout::QPROP) + lane];
    grad[(size_t)c * 17 + lane] = st[L.vec(NutsLayout::GPROP) + lane];
  }
  if (lane == 0) {
    logp[c] = sc[NutsLayout::LPPROP];
    const double acc = sc[NutsLayout::SUMACC] / fmax(sc[NutsLayout::NACC], 1.0);
    accept_out[c] = acc;
    if (depth_out) depth_out[c] = sc[NutsLayout::DEPTH];
    if (diverged_out) diverged_out[c] = sc[NutsLayout::DIVERGED];
    if (adapt) {  // Nesterov dual averaging of log(step size), as k_hmc_end
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

}  // namespace
