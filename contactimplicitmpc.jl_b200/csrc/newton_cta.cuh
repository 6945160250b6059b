// Batched `newton_solve!` (src/controller/newton.jl:169-288), :configuration mode, TrackingObjective — the hot variant of
// the Monte-Carlo benchmark — with ONE CTA OF FOUR WARPS PER ROLLOUT (round 2; replaces the warp-per-rollout kernel of
// newton_kernel.cuh, whose state machine, sweep protocol and linear algebra it keeps; that kernel remains the path for
// horizons whose δz does not fit in shared memory).
//
// Why.  The warp-per-rollout kernel is bound by exposed memory latency: every rollout walks its ten δz blocks three times
// (residual, assembly, Δx), one 2.6 KB block in flight at a time (ncu: long_scoreboard 2.9–5.8 stalls per issue, 29 % issue
// utilisation, more resident warps measured neutral).  Here
//   * ALL of a rollout's δz (H × nd × ncol doubles, 26 KB for the quadruped) is brought into shared memory by H bulk-async
//     copies (cp.async.bulk + mbarrier, the 1-D TMA path) issued by one thread at kernel entry, while the other threads
//     load the candidate point: one memory latency per rollout and sweep instead of thirty, δz read once instead of 3×;
//   * for a KKT solve the columns of δz are scaled in place by √Q⁻¹ (Y = C Q⁻¹ Cᵀ + ρI = Z̃ Z̃ᵀ + …), so the assembly's inner
//     loops are two shared-memory loads per FMA and nothing else;
//   * the Z̃ Z̃ᵀ part of block column t+1 (which does not depend on the factorisation) is computed by warps 1–3 while
//     warp 0 factorises the diagonal block of column t in registers; the pending Cholesky updates are subtracted by all
//     128 threads; `trsm` of the two sub-diagonal blocks (warps 1, 2) runs beside the forward substitution (warp 0);
//   * the backward pass prefetches the factor blocks of stage t−1 while stage t is substituted.
// The serial core — the 11 Cholesky steps and the two 11-step substitutions per stage — is unchanged.
#pragma once
#include <cstdint>

#include "dims.cuh"
#include "ip_kernel.cuh"  // mbarrier / bulk-copy wrappers
#include "newton_kernel.cuh"

namespace cimpc {

constexpr int NEWTON_CTA_WARPS = 4;
constexpr int NEWTON_CTA_THREADS = 32 * NEWTON_CTA_WARPS;
constexpr int NEWTON_CTA_MIN_CTAS = 5;  // ≤ 102 registers per thread; shared memory allows 5 quadruped rollouts per SM
constexpr size_t NEWTON_CTA_MAX_SMEM = 227 * 1024;

template <class D>
struct NewtonCtaSmem {
  static constexpr int NQ = D::NQ, NU = D::NU, ND = D::ND, NCOL = D::NCOL;
  static constexpr int BS = ND * ND, NR = NU + NQ, ZB = ND * NCOL;
  // a δz block must be a whole number of 16-byte units for the bulk copy
  static constexpr bool SUPPORTED = D::MODE == 0 && ND <= 32 && ZB % 2 == 0;
  // δz of every stage; candidate q, u, ν; rhs/solution; d; r_x; √Q⁻¹ r_x; eight ND×ND blocks (working column A0–A2,
  // factor blocks P1, P2, Q2, the pre-computed Z̃Z̃ᵀ parts ZZ0, ZZ1); potrf column ring, diagonal reciprocals, reduction
  // scratch, mbarrier
  __host__ __device__ static constexpr int doubles(int H) {
    return H * ZB + (H + 2) * NQ + H * NU + 3 * H * ND + 2 * H * NR + 8 * BS + 3 * ND + NEWTON_CTA_WARPS + 2;
  }
};

template <class D>
__global__ void __launch_bounds__(NEWTON_CTA_THREADS, NEWTON_CTA_MIN_CTAS) newton_step_cta_kernel(const NewtonParams p, double* __restrict__ lscratch) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NCOL = D::NCOL, NZ = D::NZ, NTH = D::NTH;
  constexpr int NR = NU + NQ, BS = ND * ND, T = NEWTON_CTA_THREADS, ZB = ND * NCOL;
  constexpr int NTRI = ND * (ND + 1) / 2;  // lower-triangular entries of a diagonal block
  static_assert(NewtonCtaSmem<D>::SUPPORTED, "one lane per block row; 16-byte δz blocks");
  constexpr unsigned FULLM = 0xffffffffu;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, H = p.H, R = p.R;
  const int cur = *p.par;
  const int slot = blockIdx.x;
  if (slot >= p.act_count[cur]) return;  // (CTA-uniform)
  const int r = p.act_list[(size_t)cur * R + slot];
  const int phase = p.phase[r];
  if (phase == NP_DONE) return;

  extern __shared__ __align__(16) double sm[];
  double* zall = sm;                    // H×(ND×NCOL)  δz of every stage, column-major per stage
  double* cq = zall + H * ZB;           // (H+2)×NQ
  double* cu = cq + (H + 2) * NQ;       // H×NU
  double* cnu = cu + H * NU;            // H×ND
  double* gv = cnu + H * ND;            // H×ND   rhs → y → Δν
  double* dv = gv + H * ND;             // H×ND   d_t
  double* rx = dv + H * ND;             // H×NR   [r_u; r_q] per stage
  double* rxs = rx + H * NR;            // H×NR   √Q⁻¹ r_x
  double* blk = rxs + H * NR;           // 8 blocks
  double* colb = blk + 8 * BS;          // 2×ND  column of the Cholesky step in flight
  double* rdg = colb + 2 * ND;          // ND    reciprocals of the diagonal of L_tt
  double* red = rdg + ND;               // NEWTON_CTA_WARPS partial sums
  uint64_t* bar = reinterpret_cast<uint64_t*>(red + NEWTON_CTA_WARPS);

  // ---- δz of all stages: H bulk copies in flight at once ----
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(bar, (uint32_t)(H * ZB * sizeof(double)));
    for (int s_ = 0; s_ < H; ++s_) bulk_g2s(zall + s_ * ZB, p.dz + ((size_t)s_ * R + r) * ZB, ZB * sizeof(double), bar);
  }
  const double* cand_q = p.cand_q + (size_t)r * (H + 2) * NQ;
  const double* cand_u = p.cand_u + (size_t)r * H * NU;
  const double* cand_nu = p.cand_nu + (size_t)r * H * ND;
  for (int e = tid; e < (H + 2) * NQ; e += T) cq[e] = cand_q[e];
  for (int e = tid; e < H * NU; e += T) cu[e] = cand_u[e];
  for (int e = tid; e < H * ND; e += T) cnu[e] = cand_nu[e];
  auto QIu = [&](int t, int k) -> double { return __ldg(p.obj_ui + t * NU + k); };
  auto QIq = [&](int t, int k) -> double { return __ldg(p.obj_qi + t * NQ + k); };
  auto SQu = [&](int t, int k) -> double { return __ldg(p.obj_usi + t * NU + k); };  // √(1 / obj_u)
  auto SQq = [&](int t, int k) -> double { return __ldg(p.obj_qsi + t * NQ + k); };
  // δz of stage s: element (row a, column c) — the δq0 | δq1 | δu1 views (implicit_dynamics.jl:82-86)
  auto Z = [&](int s_, int c, int a) -> double { return zall[s_ * ZB + c * ND + a]; };
  __syncthreads();
  // d_t = z*_t[1:nq] − q_{t+2}   (implicit_dynamics.jl:180-182)
  for (int e = tid; e < H * ND; e += T) {
    const int t = e / ND, i = e % ND;
    dv[e] = p.z[((size_t)t * R + r) * NZ + i] - cq[(t + 2) * NQ + i];
  }
  while (!mbar_try_wait(bar, 0)) {}
  // residual!  (newton_residual.jl:113-138), primal rows — same order of operations as newton_step_kernel
  for (int e = tid; e < H * NR; e += T) {
    const int t = e / NR, c = e % NR;
    double acc;
    if (c < NU) {  // u_t:  obj.u (u − u_ref) + δu1_tᵀ ν_t
      acc = p.obj_u[t * NU + c] * (cu[t * NU + c] - p.ref_u[t * NU + c]);
      for (int i = 0; i < ND; ++i) acc = fma(Z(t, 2 * NQ + c, i), cnu[t * ND + i], acc);
    } else {  // q_{t+2}: obj.q (q − q_ref) − ν_t + δq1_{t+1}ᵀ ν_{t+1} + δq0_{t+2}ᵀ ν_{t+2}
      const int k = c - NU;
      acc = p.obj_q[t * NQ + k] * (cq[(t + 2) * NQ + k] - p.ref_q[(t + 2) * NQ + k]) - cnu[t * ND + k];
      if (t + 1 < H)
        for (int i = 0; i < ND; ++i) acc = fma(Z(t + 1, NQ + k, i), cnu[(t + 1) * ND + i], acc);
      if (t + 2 < H)
        for (int i = 0; i < ND; ++i) acc = fma(Z(t + 2, k, i), cnu[(t + 2) * ND + i], acc);
    }
    rx[e] = acc;
  }
  __syncthreads();
  // r_cand = ‖res‖₁  (newton.jl:198, 241) — fixed-order reduction: deterministic
  double r_cand = 0.0;
  for (int e = tid; e < H * NR; e += T) r_cand += fabs(rx[e]);
  for (int e = tid; e < H * ND; e += T) r_cand += fabs(dv[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r_cand += __shfl_xor_sync(FULLM, r_cand, o);
  if (lane == 0) red[wid] = r_cand;
  __syncthreads();
  r_cand = 0.0;
#pragma unroll
  for (int w = 0; w < NEWTON_CTA_WARPS; ++w) r_cand += red[w];

  // ---- state machine (every thread evaluates the same scalars) ----
  const int len = H * (NR + ND);
  double r_norm = p.r_norm[r], alpha = p.alpha[r], beta = p.beta[r];
  int ls = p.ls_it[r], l = p.newton_it[r];
  int action;  // 0: re-evaluate at a smaller α, 1: solve for a new direction, 2: finished
  bool accept = false;
  if (phase == NP_INIT) {
    r_norm = r_cand;
    action = 1;
  } else {
    if (r_cand * r_cand >= (1.0 - 0.001 * alpha) * r_norm * r_norm) {  // newton.jl:245
      alpha *= 0.5;
      ls += 1;
      if (ls > 6) accept = true;  // the halved α is applied without being evaluated (newton.jl:249-251, 273)
      action = accept ? 1 : 0;
    } else {
      accept = true;
      action = 1;
    }
    if (accept) {
      r_norm = r_cand;
      beta = (ls > 6) ? fmin(beta * 1.3, 1.0e2) : fmax(1.0e1, beta / 1.3);  // newton.jl:280
      l += 1;
    }
  }
  if (action == 1 && (r_norm / (double)len < p.r_tol || l >= p.max_iter)) action = 2;  // newton.jl:202-206
  __syncthreads();  // everyone has read the per-rollout scalars
  if (tid == 0) {
    p.sweeps[r] += 1;
    p.r_norm[r] = r_norm;
    p.alpha[r] = alpha;
    p.beta[r] = beta;
    p.ls_it[r] = ls;
    p.newton_it[r] = l;
  }
  const double alpha_acc = alpha;
  double* traj_q = p.traj_q + (size_t)r * (H + 2) * NQ;
  double* traj_u = p.traj_u + (size_t)r * H * NU;
  double* nu = p.nu + (size_t)r * H * ND;
  double* delta = p.delta + (size_t)r * H * (NR + ND);
  double* cq_g = p.cand_q + (size_t)r * (H + 2) * NQ;
  double* cu_g = p.cand_u + (size_t)r * H * NU;
  double* cnu_g = p.cand_nu + (size_t)r * H * ND;

  if (accept) {  // update_traj!(traj, traj, ν, ν, Δ, α)  (newton.jl:273)
    for (int e = tid; e < H * NQ; e += T) {
      const int t = e / NQ, k = e % NQ;
      traj_q[(t + 2) * NQ + k] -= alpha_acc * delta[t * (NR + ND) + NU + k];
    }
    for (int e = tid; e < H * NU; e += T) traj_u[e] -= alpha_acc * delta[(e / NU) * (NR + ND) + e % NU];
    for (int e = tid; e < H * ND; e += T) nu[e] -= alpha_acc * delta[(e / ND) * (NR + ND) + NR + e % ND];
    __syncthreads();
  }

  if (action == 2) {  // finished: switch the rollout's subproblems off
    for (int t = tid; t < H; t += T) p.knot[(size_t)t * R + r] = -1;
    if (tid == 0) {
      p.phase[r] = NP_DONE;
      atomicSub(p.n_active, 1);
    }
    return;
  }

  double alpha_next;
  if (action == 1) {
    // ---------------- jacobian! + linear_solve! through the dual Schur complement ----------------
    const double rho = (double)H * beta * p.kappa;
    double* Ls = lscratch + (size_t)r * ((size_t)H * 3 * BS);
    // Z̃ = δz · √Q⁻¹ column by column (δq0_s acts on q_s = stage s−2's variable, δq1_s on stage s−1's, δu1_s on u_s); the
    // columns that act on the fixed q_0, q_1 become zero, so every product below runs over all ncol columns
    for (int e = tid; e < H * ZB; e += T) {
      const int s_ = e / ZB, c = (e % ZB) / ND;
      const double w = c < NQ ? (s_ >= 2 ? SQq(s_ - 2, c) : 0.0) : c < 2 * NQ ? (s_ >= 1 ? SQq(s_ - 1, c - NQ) : 0.0) : SQu(s_, c - 2 * NQ);
      zall[e] *= w;
    }
    for (int e = tid; e < H * NR; e += T) {
      const int t = e / NR, c = e % NR;
      rxs[e] = rx[e] * (c < NU ? SQu(t, c) : SQq(t, c - NU));
    }
    __syncthreads();
    // g = C Q⁻¹ r_x − r_ν, every stage at once
    for (int e = tid; e < H * ND; e += T) {
      const int t = e / ND, a = e % ND;
      double acc = -rx[t * NR + NU + a] * QIq(t, a) - dv[e];
      for (int k = 0; k < NU; ++k) acc = fma(Z(t, 2 * NQ + k, a), rxs[t * NR + k], acc);
      if (t >= 1)
        for (int k = 0; k < NQ; ++k) acc = fma(Z(t, NQ + k, a), rxs[(t - 1) * NR + NU + k], acc);
      if (t >= 2)
        for (int k = 0; k < NQ; ++k) acc = fma(Z(t, k, a), rxs[(t - 2) * NR + NU + k], acc);
      gv[e] = acc;
    }
    // sliding window of blocks (column-major ND×ND): working column A0, A1, A2, the factor blocks P1 = L_{t,t−1},
    // P2 = L_{t+1,t−1}, Q2 = L_{t,t−2} of the two previous block columns, and the factor-independent parts ZZ0, ZZ1
    double *A0 = blk, *A1 = blk + BS, *A2 = blk + 2 * BS, *P1 = blk + 3 * BS, *P2 = blk + 4 * BS, *Q2 = blk + 5 * BS;
    double *ZZ0 = blk + 6 * BS, *ZZ1 = blk + 7 * BS;
    // work items of a block column: NTRI lower-triangular entries of the diagonal block, then the BS entries of the first
    // sub-diagonal block
    auto item_rc = [&](int it) -> int {  // (a, b) packed as a | b << 8
      int a, b;
      if (it < NTRI) {
        b = 0;
        int rem = it;
        while (rem >= ND - b) {
          rem -= ND - b;
          ++b;
        }
        a = b + rem;
      } else {
        a = (it - NTRI) % ND;
        b = (it - NTRI) / ND;
      }
      return a | (b << 8);
    };
    // the items of this thread in the two mappings (all 128 threads / warps 1–3), decoded once
    constexpr int NIT = (NTRI + BS + T - 1) / T, NIT3 = (NTRI + BS + T - 33) / (T - 32);
    int ab_all[NIT], ab_w13[NIT3];
#pragma unroll
    for (int k = 0; k < NIT; ++k) ab_all[k] = item_rc(tid + k * T);
#pragma unroll
    for (int k = 0; k < NIT3; ++k) ab_w13[k] = item_rc((tid >= 32 ? tid - 32 : tid) + k * (T - 32));
    // ZZ0 = Z̃_t Z̃_tᵀ + Qq_t⁻¹ + ρI (lower triangle);  ZZ1 = −δq1_{t+1} Qq_t⁻¹ + Z̃q0_{t+1} Z̃q1_tᵀ
    auto zz_item = [&](int t, int it, int ab) {
      const int a = ab & 255, b = ab >> 8;
      if (it < NTRI) {
        double acc = (a == b) ? QIq(t, a) + rho : 0.0;
        const double* za = zall + t * ZB + a;
        const double* zb = zall + t * ZB + b;
#pragma unroll 6
        for (int c = 0; c < NCOL; ++c) acc = fma(za[c * ND], zb[c * ND], acc);
        ZZ0[a + b * ND] = acc;
      } else if (it < NTRI + BS && t + 1 < H) {
        double c1 = -Z(t + 1, NQ + b, a) * SQq(t, b);
        const double* za = zall + (t + 1) * ZB + a;
        const double* zb = zall + t * ZB + NQ * ND + b;
#pragma unroll
        for (int k = 0; k < NQ; ++k) c1 = fma(za[k * ND], zb[k * ND], c1);
        ZZ1[a + b * ND] = c1;
      }
    };
#pragma unroll
    for (int k = 0; k < NIT; ++k) zz_item(0, tid + k * T, ab_all[k]);
    __syncthreads();
    for (int t = 0; t < H; ++t) {
      // ---- phase A (all threads): block column t = factor-independent part − pending Cholesky updates; the two
      //      sub-diagonal blocks of column t−1 (now P1, P2) go to the scratch for the backward pass ----
#pragma unroll
      for (int kk = 0; kk < NIT; ++kk) {
        const int it = tid + kk * T, a = ab_all[kk] & 255, b = ab_all[kk] >> 8;
        if (it < NTRI) {
          double acc = ZZ0[a + b * ND];
          if (t >= 1)
            for (int k = 0; k < ND; ++k) acc = fma(-P1[a + k * ND], P1[b + k * ND], acc);
          if (t >= 2)
            for (int k = 0; k < ND; ++k) acc = fma(-Q2[a + k * ND], Q2[b + k * ND], acc);
          A0[a + b * ND] = acc;
        } else if (it < NTRI + BS && t + 1 < H) {
          double c1 = ZZ1[a + b * ND];
          if (t >= 1)
            for (int k = 0; k < ND; ++k) c1 = fma(-P2[a + k * ND], P1[b + k * ND], c1);
          A1[a + b * ND] = c1;
        }
      }
      if (t >= 1)
        for (int e = tid; e < BS; e += T) {
          Ls[(size_t)(3 * t - 2) * BS + e] = P1[e];
          if (t + 1 < H) Ls[(size_t)(3 * t - 1) * BS + e] = P2[e];
        }
      __syncthreads();
      // ---- phase B: warp 0 factorises A0 = L Lᵀ (row in registers, one published column per step); warps 1–3 compute
      //      the factor-independent part of block column t+1 ----
      if (wid == 0) {
        double rw[ND];
#pragma unroll
        for (int c = 0; c < ND; ++c) rw[c] = (lane < ND && c <= lane) ? A0[lane + c * ND] : 0.0;
#pragma unroll
        for (int j = 0; j < ND; ++j) {
          const double ajj = __shfl_sync(FULLM, rw[j], j);
          const double inv = rsqrt(ajj);  // 1 / l_jj
          const double lij = (lane == j) ? ajj * inv : rw[j] * inv;
          rw[j] = lij;
          double* cb = colb + (j & 1) * ND;
          if (lane >= j && lane < ND) cb[lane] = lij;
          if (lane == j) rdg[j] = inv;
          __syncwarp();
#pragma unroll
          for (int c = j + 1; c < ND; ++c)
            if (c <= lane) rw[c] = fma(-lij, cb[c], rw[c]);
        }
#pragma unroll
        for (int c = 0; c < ND; ++c)
          if (lane < ND && c <= lane) A0[lane + c * ND] = rw[c];
      } else if (t + 1 < H) {
#pragma unroll
        for (int k = 0; k < NIT3; ++k) zz_item(t + 1, tid - 32 + k * (T - 32), ab_w13[k]);
      }
      __syncthreads();
      // ---- phase C: warps 1, 2: trsm  [A1; A2] ← [A1; A2] L⁻ᵀ (one row per lane, the row in registers; the (t+2,t)
      //      block −δq0_{t+2} Qq_t⁻¹ is read straight from Z̃);  warp 0: forward substitution
      //      y_t = L_tt⁻¹ (g_t − L_{t,t−1} y_{t−1} − L_{t,t−2} y_{t−2}) by shuffle;  warp 3: L_tt to the scratch ----
      if (wid == 1 || wid == 2) {
        const bool second = wid == 2;
        if (lane < ND && (second ? t + 2 < H : t + 1 < H)) {
          double* X = (second ? A2 : A1) + lane;
          double x[ND];
#pragma unroll
          for (int c = 0; c < ND; ++c) x[c] = second ? -Z(t + 2, c, lane) * SQq(t, c) : X[c * ND];
#pragma unroll
          for (int c = 0; c < ND; ++c) {
            double sacc = x[c];
#pragma unroll
            for (int k = 0; k < c; ++k) sacc = fma(-x[k], A0[c + k * ND], sacc);
            x[c] = sacc * rdg[c];
          }
#pragma unroll
          for (int c = 0; c < ND; ++c) X[c * ND] = x[c];
        }
      } else if (wid == 0) {
        double sreg = 0.0, lrow[ND];
        if (lane < ND) {
          sreg = gv[t * ND + lane];
          if (t >= 1)
            for (int k = 0; k < ND; ++k) sreg = fma(-P1[lane + k * ND], gv[(t - 1) * ND + k], sreg);
          if (t >= 2)
            for (int k = 0; k < ND; ++k) sreg = fma(-Q2[lane + k * ND], gv[(t - 2) * ND + k], sreg);
        }
        const double rdl = (lane < ND) ? rdg[lane] : 0.0;
#pragma unroll
        for (int c = 0; c < ND; ++c) lrow[c] = (lane < ND && c < lane) ? A0[lane + c * ND] : 0.0;
#pragma unroll
        for (int c = 0; c < ND; ++c) {
          const double yc = __shfl_sync(FULLM, sreg * rdl, c);
          sreg = (lane == c) ? yc : fma(-lrow[c], yc, sreg);  // lrow[c] = 0 for the rows above c
        }
        if (lane < ND) gv[t * ND + lane] = sreg;
      } else {
        for (int e = lane; e < BS; e += 32) Ls[(size_t)(3 * t) * BS + e] = A0[e];
      }
      __syncthreads();
      // ---- slide the window ----
      double* oldQ2 = Q2;
      Q2 = P2;     // L_{t+1,t−1} is L_{t',t'−2} of the next column t' = t+1
      P2 = A2;     // L_{t+2,t}   →  L_{t'+1,t'−1}
      double* oldP1 = P1;
      P1 = A1;     // L_{t+1,t}   →  L_{t',t'−1}
      A1 = oldP1;
      A2 = oldQ2;
    }
    // ---- backward: Δν_t = L_tt⁻ᵀ (y_t − L_{t+1,t}ᵀ Δν_{t+1} − L_{t+2,t}ᵀ Δν_{t+2}); the three factor blocks of stage
    //      t − 1 are fetched (warps 1–3) into the other half of the block area while warp 0 substitutes stage t ----
    auto load_factor = [&](int t_, int first, int stride) {
      const double* Lg = Ls + (size_t)(3 * t_) * BS;
      double* dst = blk + (t_ & 1) * 3 * BS;
      const int nblk = 1 + (t_ + 1 < H ? 1 : 0) + (t_ + 2 < H ? 1 : 0);
      for (int e = first; e < nblk * BS; e += stride) dst[e] = Lg[e];
    };
    load_factor(H - 1, tid, T);
    __syncthreads();
    for (int t = H - 1; t >= 0; --t) {
      if (wid == 0) {
        const double *L0 = blk + (t & 1) * 3 * BS, *L1 = L0 + BS, *L2 = L0 + 2 * BS;
        double sreg = 0.0, rdl = 0.0, lcol[ND];
        if (lane < ND) {
          sreg = gv[t * ND + lane];
          if (t + 1 < H)
            for (int k = 0; k < ND; ++k) sreg = fma(-L1[k + lane * ND], gv[(t + 1) * ND + k], sreg);
          if (t + 2 < H)
            for (int k = 0; k < ND; ++k) sreg = fma(-L2[k + lane * ND], gv[(t + 2) * ND + k], sreg);
          rdl = 1.0 / L0[lane + lane * ND];
        }
#pragma unroll
        for (int c = 0; c < ND; ++c) lcol[c] = (lane < ND && c > lane) ? L0[c + lane * ND] : 0.0;
#pragma unroll
        for (int cc = 0; cc < ND; ++cc) {
          const int c = ND - 1 - cc;
          const double xc = __shfl_sync(FULLM, sreg * rdl, c);
          sreg = (lane == c) ? xc : fma(-lcol[c], xc, sreg);  // lcol[c] = 0 for the rows below c
        }
        if (lane < ND) gv[t * ND + lane] = sreg;
      } else if (t > 0) {
        load_factor(t - 1, tid - 32, T - 32);
      }
      __syncthreads();
    }
    // Δx = Q⁻¹ (r_x − Cᵀ Δν) = Q⁻¹ r_x − √Q⁻¹ Z̃ᵀ Δν;  Δ = [Δu, Δq, Δν] per stage
    for (int e = tid; e < H * NR; e += T) {
      const int t = e / NR, c = e % NR;
      double dot = 0.0;
      if (c < NU) {
        for (int i = 0; i < ND; ++i) dot = fma(Z(t, 2 * NQ + c, i), gv[t * ND + i], dot);
        delta[t * (NR + ND) + c] = fma(QIu(t, c), rx[e], -SQu(t, c) * dot);
      } else {
        const int k = c - NU;
        if (t + 1 < H)
          for (int i = 0; i < ND; ++i) dot = fma(Z(t + 1, NQ + k, i), gv[(t + 1) * ND + i], dot);
        if (t + 2 < H)
          for (int i = 0; i < ND; ++i) dot = fma(Z(t + 2, k, i), gv[(t + 2) * ND + i], dot);
        delta[t * (NR + ND) + c] = fma(QIq(t, k), rx[e] + gv[t * ND + k], -SQq(t, k) * dot);
      }
    }
    for (int e = tid; e < H * ND; e += T) delta[(e / ND) * (NR + ND) + NR + e % ND] = gv[e];
    __syncthreads();
    alpha_next = 1.0;
    if (tid == 0) {
      p.alpha[r] = 1.0;
      p.ls_it[r] = 0;
      p.phase[r] = NP_LS;
    }
  } else {
    alpha_next = alpha_acc;  // back-track: same Δ, smaller α
  }

  // candidate = traj − α Δ  (update_traj!, newton_residual.jl:160-176), then the next sweep's inputs
  for (int e = tid; e < 2 * NQ; e += T) cq[e] = traj_q[e];
  for (int e = tid; e < H * NQ; e += T) {
    const int t = e / NQ, k = e % NQ;
    cq[(t + 2) * NQ + k] = traj_q[(t + 2) * NQ + k] - alpha_next * delta[t * (NR + ND) + NU + k];
  }
  for (int e = tid; e < H * NU; e += T) cu[e] = traj_u[e] - alpha_next * delta[(e / NU) * (NR + ND) + e % NU];
  for (int e = tid; e < H * ND; e += T) cnu_g[e] = nu[e] - alpha_next * delta[(e / ND) * (NR + ND) + NR + e % ND];
  __syncthreads();
  for (int e = tid; e < (H + 2) * NQ; e += T) cq_g[e] = cq[e];
  for (int e = tid; e < H * NU; e += T) cu_g[e] = cu[e];
  // θ_t = [q_t; q_{t+1}; u_t; w_t; μ; h]  (update_θ!, trajectory.jl:67-82), cold start q2 = q_{t+2}
  for (int e = tid; e < H * NTH; e += T) {
    const int t = e / NTH, c = e % NTH;
    double v;
    if (c < NQ) v = cq[t * NQ + c];
    else if (c < 2 * NQ) v = cq[(t + 1) * NQ + c - NQ];
    else if (c < 2 * NQ + NU) v = cu[t * NU + c - 2 * NQ];
    else if (c < 2 * NQ + NU + NW) v = p.w[t * NW + c - 2 * NQ - NU];
    else if (c == 2 * NQ + NU + NW) v = p.call->mu;
    else v = p.call->h;
    p.theta[((size_t)t * R + r) * NTH + c] = v;
  }
  for (int e = tid; e < H * NQ; e += T) {
    const int t = e / NQ, k = e % NQ;
    p.q2[((size_t)t * R + r) * NQ + k] = cq[(t + 2) * NQ + k];
  }
  if (tid == 0) p.act_list[(size_t)(cur ^ 1) * R + atomicAdd(&p.act_count[cur ^ 1], 1)] = r;  // takes part in the next sweep
}

}  // namespace cimpc
