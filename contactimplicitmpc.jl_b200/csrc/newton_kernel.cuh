// Batched `newton_solve!` (src/controller/newton.jl:169-288), :configuration mode, TrackingObjective.
//
// One CTA per Monte-Carlo rollout.  A call of newton_step_kernel consumes the result of ONE
// `implicit_dynamics!` sweep (done by ip_solve_kernel over all (stage, rollout) subproblems) and advances
// that rollout's state machine:
//     INIT  — first sweep at the reset trajectory: residual, convergence check, KKT solve, first candidate
//     LS    — sweep at a candidate: Armijo-type test of newton.jl:245; back-track (α/2) or accept, then
//             β update, convergence / iteration-cap check, next KKT solve and candidate
//     DONE  — its subproblems are switched off (knot = −1)
// and writes the θ / cold-start inputs of the next sweep.  The host alternates the two kernels until
// no rollout is active (one 4-byte read per round).
//
// KKT solve.  The reference assembles the sparse symmetric matrix
//        R = [ Q   Cᵀ ]      Q  = diag(obj.u, obj.q)                     (hessian!, newton_jacobian.jl:209-215)
//            [ C  −ρI ]      C  = [δu1_t on u_t, δq1_t on q_{t+1}, δq0_t on q_t, −I on q_{t+2}]   (:121-126, :159-181)
//                            ρ  = H·β·κ  (`reg_du .-= βκ` once per stage, :185)
// (300×300 for the quadruped) and calls a dense / sparse LU (src/solver/lu.jl:4-12).  Q is diagonal and
// positive, so here Δν solves the dual Schur complement (C Q⁻¹ Cᵀ + ρI) Δν = C Q⁻¹ r_x − r_ν — a
// block-pentadiagonal SPD system of H blocks of nd (110×110, half-bandwidth 32) factorised by banded
// Cholesky in shared memory — and Δx = Q⁻¹ (r_x − Cᵀ Δν).  Same Δ as R \ r up to round-off
// (tests/test_gpu_newton.py), ≈ 150× fewer flops than the dense LU.
#pragma once
#include <cstdint>

#include "dims.cuh"

namespace cimpc {

enum NewtonPhase : int { NP_INIT = 0, NP_LS = 1, NP_DONE = 2 };

// Per-call arguments of one batched newton_solve! (device copy read by the reset / finish kernels of the solve graph,
// so that the instantiated graph never changes).
struct NewtonCall {
  const double* q0;       // nq × R
  const double* q1;       // nq × R
  const uint8_t* active;  // R or null
  const double* alt;      // nc × R per-rollout altitude offsets (policy.jl:111-115, update_altitude!), or null
  double* u_out;          // nu × R
  double* q_out;          // nq × (H+2) × R or null
  double* y_out;          // nyd × H × R or null
  int32_t* info;          // 4 × R or null
  int32_t warm_start;
  int32_t pad;
  double mu, h;           // θ entries μ and h of this solve
};

struct NewtonParams {
  int R, H;
  // shared by every rollout (they track the same gait in lock-step, policy.jl:100-107, 131)
  const double* ref_q;    // (H+2) × nq   `ref_traj.q[1:H+2]`
  const double* ref_u;    // H × nu
  const double* w;        // H × nw       disturbances in θ (zeros in the reference's MPC)
  const int32_t* window;  // H            knot of every stage
  const double* obj_q;    // H × nq       diagonal of obj.q[t]
  const double* obj_u;    // H × nu
  const NewtonCall* call; // per-call arguments (device)
  double kappa;
  double r_tol, beta_init;
  int max_iter;
  // per rollout state
  double *traj_q, *traj_u, *nu;        // R×(H+2)×nq, R×H×nu, R×H×nd   accepted point
  double *cand_q, *cand_u, *cand_nu;   // trial point (the one the last sweep evaluated)
  double* delta;                       // R×H×(nu+nq+nd)  Newton direction [Δu, Δq, Δν] per stage
  double *r_norm, *alpha, *beta;       // R
  int *ls_it, *newton_it, *sweeps, *phase;  // R
  // implicit-dynamics buffers, stage-major: subproblem (t, r) at index t*R + r
  int32_t* knot;     // H×R   window[t] while the rollout is active, −1 afterwards
  double* theta;     // H×R×nθ
  double* q2;        // H×R×nq
  const double* z;   // H×R×nz
  const double* dz;  // H×R×(nd×ncol), column-major per subproblem
  int* n_active;     // [1] rollouts still iterating (decremented when a rollout finishes)
  // compaction of the sweeps: two lists of rollouts (unordered) and their lengths.  List *par is the one the CURRENT
  // sweep runs over (ip_solve_kernel and newton_step read it); rollouts that ask for another sweep append themselves to
  // the other one; newton_advance_kernel flips `par` between sweeps.
  int32_t* act_list;  // 2 × R
  int* act_count;     // [2]
  int* par;           // [1]
  int* sweep_ctr;     // [1] sweeps done in this solve
  double* alt;        // R × nc  altitude offsets of every rollout for this solve (zeros when the caller passes none)
  // general variant only (newton_general.cuh): contact forces y = [γ; b] as Newton variables (:configurationforce)
  // and the velocity weights of a TrackingVelocityObjective
  double *traj_y = nullptr, *cand_y = nullptr;  // R×H×nyd
  const double* ref_y = nullptr;                // H×nyd   `ref_traj.γ`, `ref_traj.b`
  const double* obj_y = nullptr;                // H×nyd   diagonals of obj.γ, obj.b
  const double* obj_v = nullptr;                // H×nq    diagonals of obj.v (null: TrackingObjective)
  // dense-weight variant only (newton_dense.cuh)
  const double* obj_qi = nullptr;               // H×nq    1 / obj_q   (specialised kernel)
  const double* obj_ui = nullptr;               // H×nu    1 / obj_u
  const double* obj_qsi = nullptr;              // H×nq    √(1 / obj_q)  (newton_cta.cuh)
  const double* obj_usi = nullptr;              // H×nu    √(1 / obj_u)
  const double* obj_qd = nullptr;               // H×nq×nq column-major  obj.q[t]
  const double* obj_e = nullptr;                // H×nq×nq column-major  E_t = L_t⁻ᵀ, obj.q[t] = L_t L_tᵀ
  const double* q_tgt = nullptr;                // H×nq    obj.q_target (null: zeros)
  const double* v_tgt = nullptr;                // H×nq    obj.v_target (null: zeros)
};

template <class D>
struct NewtonSmem {
  static constexpr int NQ = D::NQ, NU = D::NU, ND = D::ND, NCOL = D::NCOL;
  static_assert(D::MODE == 0, "device Newton: :configuration mode");
  static_assert(ND <= 32, "one lane per block row");
  static constexpr int BS = ND * ND;  // one ND×ND block, column-major
  // per-warp shared memory (doubles): candidate q, u, ν; rhs/solution; d; r_x; Q⁻¹; six sliding-window
  // factor blocks; a TWO-stage window of δz blocks (stages t, t+1; the single use of stage t+2 — the (t+2, t) block —
  // reads it from global memory) and Q⁻¹ from the per-objective reciprocal table in global memory (uniform, L1-resident
  // loads): 17.3 instead of 21.4 KB per rollout, 12 instead of 10 resident warps per SM
  __host__ __device__ static constexpr int per_warp(int H) {
    return (H + 2) * NQ + H * NU + 3 * H * ND + H * (NU + NQ) + 6 * BS + 2 * ND * NCOL + 3 * ND + 8;
  }
  // global scratch per rollout: the three block columns L_tt, L_{t+1,t}, L_{t+2,t} of every stage
  __host__ __device__ static constexpr size_t l_doubles(int H) { return (size_t)H * 3 * BS; }
};

constexpr int NEWTON_WARPS = 2;  // rollouts per CTA (one warp each; 21 KB of shared memory per rollout)

// One warp per rollout; everything is warp-synchronous (no __syncthreads).
//
// KKT solve.  Y = C Q⁻¹ Cᵀ + ρI is block-pentadiagonal (H×H blocks of nd).  It is assembled, factorised
// (block Cholesky) and forward-substituted one block column at a time inside a six-block sliding window
// in shared memory; the factor's block columns go to a global scratch (L2-resident) for the backward pass.
template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_step_kernel(const NewtonParams p, double* __restrict__ lscratch) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NCOL = D::NCOL, NZ = D::NZ, NTH = D::NTH;
  constexpr int NR = NU + NQ;  // primal block of one stage: [u_t; q_{t+2}]
  using SM = NewtonSmem<D>;
  constexpr int BS = SM::BS;
  constexpr unsigned FULLM = 0xffffffffu;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, H = p.H, R = p.R;
  const int cur = *p.par;
  const int slot = blockIdx.x * (THREADS / 32) + wid;
  if (slot >= p.act_count[cur]) return;
  const int r = p.act_list[(size_t)cur * R + slot];
  const int phase = p.phase[r];
  if (phase == NP_DONE) return;

  extern __shared__ __align__(16) double sm[];
  double* base = sm + (size_t)wid * SM::per_warp(H);
  double* cq = base;                    // (H+2)×NQ
  double* cu = cq + (H + 2) * NQ;       // H×NU
  double* cnu = cu + H * NU;            // H×ND
  double* gv = cnu + H * ND;            // H×ND   rhs → y → Δν
  double* dv = gv + H * ND;             // H×ND   d_t
  double* rx = dv + H * ND;             // H×NR   [r_u; r_q] per stage
  double* blk = rx + H * NR;            // 6 factor blocks
  double* zwin = blk + 6 * BS;          // δz of stages t, t+1 (ring of 2)
  double* colb = zwin + 2 * ND * NCOL;  // 2×ND  column of the Cholesky step in flight (double-buffered by parity)
  double* rdg = colb + 2 * ND;          // ND    reciprocals of the diagonal of L_tt

  const double* cand_q = p.cand_q + (size_t)r * (H + 2) * NQ;
  const double* cand_u = p.cand_u + (size_t)r * H * NU;
  const double* cand_nu = p.cand_nu + (size_t)r * H * ND;
  for (int e = lane; e < (H + 2) * NQ; e += 32) cq[e] = cand_q[e];
  for (int e = lane; e < H * NU; e += 32) cu[e] = cand_u[e];
  for (int e = lane; e < H * ND; e += 32) cnu[e] = cand_nu[e];
  __syncwarp();
  // δz of stage t: element (row a, column c) — the δq0 | δq1 | δu1 views (implicit_dynamics.jl:82-86)
  auto DZ = [&](int t, int c, int a) -> double { return p.dz[(((size_t)t * R + r) * NCOL + c) * ND + a]; };
  __syncwarp();
  auto QIu = [&](int t, int k) -> double { return __ldg(p.obj_ui + t * NU + k); };
  auto QIq = [&](int t, int k) -> double { return __ldg(p.obj_qi + t * NQ + k); };

  // d_t = z*_t[1:nq] − q_{t+2}   (implicit_dynamics.jl:180-182)
  for (int e = lane; e < H * ND; e += 32) {
    const int t = e / ND, i = e % ND;
    dv[e] = p.z[((size_t)t * R + r) * NZ + i] - cq[(t + 2) * NQ + i];
  }
  // residual!  (newton_residual.jl:113-138), primal rows
  for (int e = lane; e < H * NR; e += 32) {
    const int t = e / NR, c = e % NR;
    double acc;
    if (c < NU) {  // u_t:  obj.u (u − u_ref) + δu1_tᵀ ν_t
      acc = p.obj_u[t * NU + c] * (cu[t * NU + c] - p.ref_u[t * NU + c]);
      for (int i = 0; i < ND; ++i) acc = fma(DZ(t, 2 * NQ + c, i), cnu[t * ND + i], acc);
    } else {  // q_{t+2}: obj.q (q − q_ref) − ν_t + δq1_{t+1}ᵀ ν_{t+1} + δq0_{t+2}ᵀ ν_{t+2}
      const int k = c - NU;
      acc = p.obj_q[t * NQ + k] * (cq[(t + 2) * NQ + k] - p.ref_q[(t + 2) * NQ + k]) - cnu[t * ND + k];
      if (t + 1 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 1, NQ + k, i), cnu[(t + 1) * ND + i], acc);
      if (t + 2 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 2, k, i), cnu[(t + 2) * ND + i], acc);
    }
    rx[e] = acc;
  }
  __syncwarp();
  // r_cand = ‖res‖₁  (newton.jl:198, 241) — fixed-order reduction: deterministic
  double r_cand = 0.0;
  for (int e = lane; e < H * NR; e += 32) r_cand += fabs(rx[e]);
  for (int e = lane; e < H * ND; e += 32) r_cand += fabs(dv[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r_cand += __shfl_xor_sync(FULLM, r_cand, o);

  // ---- state machine (uniform across the warp: every lane evaluates the same scalars) ----
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
  if (lane == 0) {
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
    for (int e = lane; e < H * NQ; e += 32) {
      const int t = e / NQ, k = e % NQ;
      traj_q[(t + 2) * NQ + k] -= alpha_acc * delta[t * (NR + ND) + NU + k];
    }
    for (int e = lane; e < H * NU; e += 32) traj_u[e] -= alpha_acc * delta[(e / NU) * (NR + ND) + e % NU];
    for (int e = lane; e < H * ND; e += 32) nu[e] -= alpha_acc * delta[(e / ND) * (NR + ND) + NR + e % ND];
    __syncwarp();
  }

  if (action == 2) {  // finished: switch the rollout's subproblems off
    for (int t = lane; t < H; t += 32) p.knot[(size_t)t * R + r] = -1;
    if (lane == 0) {
      p.phase[r] = NP_DONE;
      atomicSub(p.n_active, 1);
    }
    return;
  }

  double alpha_next;
  if (action == 1) {
    // ---------------- jacobian! + linear_solve! through the dual Schur complement ----------------
    const double rho = (double)H * beta * p.kappa;
    double* Ls = lscratch + (size_t)r * SM::l_doubles(H);
    // g = C Q⁻¹ r_x − r_ν
    for (int e = lane; e < H * ND; e += 32) {
      const int t = e / ND, a = e % ND;
      double acc = -rx[t * NR + NU + a] * QIq(t, a) - dv[e];
      for (int k = 0; k < NU; ++k) acc = fma(DZ(t, 2 * NQ + k, a), rx[t * NR + k] * QIu(t, k), acc);
      if (t >= 1)
        for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, NQ + k, a), rx[(t - 1) * NR + NU + k] * QIq(t - 1, k), acc);
      if (t >= 2)
        for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, k, a), rx[(t - 2) * NR + NU + k] * QIq(t - 2, k), acc);
      gv[e] = acc;
    }
    __syncwarp();
    // sliding window of blocks (column-major ND×ND): working column A0, A1, A2 and the factor blocks
    // P1 = L_{t,t−1}, P2 = L_{t+1,t−1}, Q2 = L_{t,t−2} of the two previous block columns
    double *A0 = blk, *A1 = blk + BS, *A2 = blk + 2 * BS, *P1 = blk + 3 * BS, *P2 = blk + 4 * BS, *Q2 = blk + 5 * BS;
    // δz window: stage s lives in slot s % 2 (coalesced 330-double copies, L2 → shared)
    auto load_stage = [&](int s_) {
      const double* src = p.dz + ((size_t)s_ * R + r) * (ND * NCOL);
      double* dst = zwin + (s_ % 2) * (ND * NCOL);
      for (int e = lane; e < ND * NCOL; e += 32) dst[e] = src[e];
    };
    auto ZW = [&](int s_, int c, int a) -> double { return zwin[(s_ % 2) * (ND * NCOL) + c * ND + a]; };
    load_stage(0);
    for (int t = 0; t < H; ++t) {
      if (t + 1 < H) load_stage(t + 1);  // overwrites stage t − 1, whose last reader was block column t − 1
      __syncwarp();
      // ---- assemble block column t of Y minus the pending Cholesky updates ----
      for (int e = lane; e < BS; e += 32) {
        const int a = e % ND, b = e / ND;
        // (t,t): δu1 Qu⁻¹ δu1ᵀ + Qq_t⁻¹ + δq1 Qq_{t−1}⁻¹ δq1ᵀ + δq0 Qq_{t−2}⁻¹ δq0ᵀ + ρI   (lower triangle)
        double acc = 0.0;
        if (a >= b) {
          for (int k = 0; k < NU; ++k) acc = fma(ZW(t, 2 * NQ + k, a) * QIu(t, k), ZW(t, 2 * NQ + k, b), acc);
          if (a == b) acc += QIq(t, a) + rho;
          if (t >= 1) {
            for (int k = 0; k < NQ; ++k) acc = fma(ZW(t, NQ + k, a) * QIq(t - 1, k), ZW(t, NQ + k, b), acc);
            for (int k = 0; k < ND; ++k) acc = fma(-P1[a + k * ND], P1[b + k * ND], acc);
          }
          if (t >= 2) {
            for (int k = 0; k < NQ; ++k) acc = fma(ZW(t, k, a) * QIq(t - 2, k), ZW(t, k, b), acc);
            for (int k = 0; k < ND; ++k) acc = fma(-Q2[a + k * ND], Q2[b + k * ND], acc);
          }
        }
        A0[e] = acc;
        // (t+1,t): −δq1_{t+1} Qq_t⁻¹ + δq0_{t+1} Qq_{t−1}⁻¹ δq1_tᵀ  − L_{t+1,t−1} L_{t,t−1}ᵀ
        if (t + 1 < H) {
          double c1 = -ZW(t + 1, NQ + b, a) * QIq(t, b);
          if (t >= 1) {
            for (int k = 0; k < NQ; ++k) c1 = fma(ZW(t + 1, k, a) * QIq(t - 1, k), ZW(t, NQ + k, b), c1);
            for (int k = 0; k < ND; ++k) c1 = fma(-P2[a + k * ND], P1[b + k * ND], c1);
          }
          A1[e] = c1;
        }
        // (t+2,t): −δq0_{t+2} Qq_t⁻¹
        if (t + 2 < H) A2[e] = -DZ(t + 2, b, a) * QIq(t, b);
      }
      __syncwarp();
      // ---- potrf: A0 = L Lᵀ, lane = row, the row lives in REGISTERS; one published column + ≤ ND−1 FMAs per step
      //      (v1 did a read-modify-write of shared memory per element: 4× the dependent latency) ----
      {
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
      }
      __syncwarp();
      // ---- trsm: A1 ← A1 L⁻ᵀ, A2 ← A2 L⁻ᵀ (one row of [A1; A2] per lane pass, the row in registers);
      //      forward substitution for y_t ----
      {
        const int nrows = (t + 1 < H ? ND : 0) + (t + 2 < H ? ND : 0);
        for (int row = lane; row < nrows; row += 32) {
          double* X = (row < ND) ? (A1 + row) : (A2 + row - ND);
          double x[ND];
#pragma unroll
          for (int c = 0; c < ND; ++c) x[c] = X[c * ND];
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
        __syncwarp();
        // y_t = L_tt⁻¹ (g_t − L_{t,t−1} y_{t−1} − L_{t,t−2} y_{t−2}): the right-hand side stays in a register of its
        // row's lane, y_c travels by shuffle (no shared-memory round trip and no __syncwarp per step)
        {
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
        }
        __syncwarp();
      }
      // ---- keep the block column for the backward pass, slide the window ----
      for (int e = lane; e < BS; e += 32) {
        Ls[(size_t)(3 * t) * BS + e] = A0[e];
        if (t + 1 < H) Ls[(size_t)(3 * t + 1) * BS + e] = A1[e];
        if (t + 2 < H) Ls[(size_t)(3 * t + 2) * BS + e] = A2[e];
      }
      __syncwarp();
      double* oldQ2 = Q2;
      Q2 = P2;     // L_{t+1,t−1} is L_{t',t'−2} of the next column t' = t+1
      P2 = A2;     // L_{t+2,t}   →  L_{t'+1,t'−1}
      double* oldP1 = P1;
      P1 = A1;     // L_{t+1,t}   →  L_{t',t'−1}
      A1 = oldP1;
      A2 = oldQ2;
    }
    // ---- backward: Δν_t = L_tt⁻ᵀ (y_t − L_{t+1,t}ᵀ Δν_{t+1} − L_{t+2,t}ᵀ Δν_{t+2}) ----
    // The three factor blocks of a stage are one contiguous run of the scratch: one coalesced copy into the (now idle)
    // block area, then the same register / shuffle substitution as in the forward pass.
    for (int t = H - 1; t >= 0; --t) {
      const double* Lg = Ls + (size_t)(3 * t) * BS;
      const int nblk = 1 + (t + 1 < H ? 1 : 0) + (t + 2 < H ? 1 : 0);
      for (int e = lane; e < nblk * BS; e += 32) blk[e] = Lg[e];
      __syncwarp();
      const double *L0 = blk, *L1 = blk + BS, *L2 = blk + 2 * BS;
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
      __syncwarp();
    }
    // Δx = Q⁻¹ (r_x − Cᵀ Δν);  Δ = [Δu, Δq, Δν] per stage
    for (int e = lane; e < H * NR; e += 32) {
      const int t = e / NR, c = e % NR;
      double acc = rx[e];
      if (c < NU) {
        for (int i = 0; i < ND; ++i) acc = fma(-DZ(t, 2 * NQ + c, i), gv[t * ND + i], acc);
        delta[t * (NR + ND) + c] = QIu(t, c) * acc;
      } else {
        const int k = c - NU;
        acc += gv[t * ND + k];
        if (t + 1 < H)
          for (int i = 0; i < ND; ++i) acc = fma(-DZ(t + 1, NQ + k, i), gv[(t + 1) * ND + i], acc);
        if (t + 2 < H)
          for (int i = 0; i < ND; ++i) acc = fma(-DZ(t + 2, k, i), gv[(t + 2) * ND + i], acc);
        delta[t * (NR + ND) + c] = QIq(t, k) * acc;
      }
    }
    for (int e = lane; e < H * ND; e += 32) delta[(e / ND) * (NR + ND) + NR + e % ND] = gv[e];
    __syncwarp();
    alpha_next = 1.0;
    if (lane == 0) {
      p.alpha[r] = 1.0;
      p.ls_it[r] = 0;
      p.phase[r] = NP_LS;
    }
  } else {
    alpha_next = alpha_acc;  // back-track: same Δ, smaller α
  }

  // candidate = traj − α Δ  (update_traj!, newton_residual.jl:160-176), then the next sweep's inputs
  for (int e = lane; e < 2 * NQ; e += 32) cq[e] = traj_q[e];
  for (int e = lane; e < H * NQ; e += 32) {
    const int t = e / NQ, k = e % NQ;
    cq[(t + 2) * NQ + k] = traj_q[(t + 2) * NQ + k] - alpha_next * delta[t * (NR + ND) + NU + k];
  }
  for (int e = lane; e < H * NU; e += 32) cu[e] = traj_u[e] - alpha_next * delta[(e / NU) * (NR + ND) + e % NU];
  for (int e = lane; e < H * ND; e += 32) cnu_g[e] = nu[e] - alpha_next * delta[(e / ND) * (NR + ND) + NR + e % ND];
  __syncwarp();
  for (int e = lane; e < (H + 2) * NQ; e += 32) cq_g[e] = cq[e];
  for (int e = lane; e < H * NU; e += 32) cu_g[e] = cu[e];
  // θ_t = [q_t; q_{t+1}; u_t; w_t; μ; h]  (update_θ!, trajectory.jl:67-82), cold start q2 = q_{t+2}
  for (int e = lane; e < H * NTH; e += 32) {
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
  for (int e = lane; e < H * NQ; e += 32) {
    const int t = e / NQ, k = e % NQ;
    p.q2[((size_t)t * R + r) * NQ + k] = cq[(t + 2) * NQ + k];
  }
  if (lane == 0) p.act_list[(size_t)(cur ^ 1) * R + atomicAdd(&p.act_count[cur ^ 1], 1)] = r;  // takes part in the next sweep
}

// reset!  (newton.jl:130-167): traj ← ref (cold) ; q[1], q[2] ← q0, q1 ; candidate ← traj ; first sweep inputs.
template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_reset_kernel(const NewtonParams p, const NewtonCall* __restrict__ call) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NTH = D::NTH, NYD = D::NYD;
  const int r = blockIdx.x, tid = threadIdx.x, H = p.H, R = p.R;
  if (r >= R) return;
  const double* __restrict__ q0 = call->q0;
  const double* __restrict__ q1 = call->q1;
  const uint8_t* __restrict__ active = call->active;
  const int warm_start = call->warm_start;
  const double* __restrict__ alt = call->alt;
  if constexpr (NYD > 0) {  // γ, b follow the reference on a cold start (newton.jl:139-151); candidate ← traj
    double* ty = p.traj_y + (size_t)r * H * NYD;
    double* cy = p.cand_y + (size_t)r * H * NYD;
    for (int e = tid; e < H * NYD; e += THREADS) {
      const double v = warm_start ? ty[e] : p.ref_y[e];
      ty[e] = v;
      cy[e] = v;
    }
  }
  double* traj_q = p.traj_q + (size_t)r * (H + 2) * NQ;
  double* traj_u = p.traj_u + (size_t)r * H * NU;
  double* nu = p.nu + (size_t)r * H * ND;
  if (!warm_start) {
    for (int e = tid; e < (H + 2) * NQ; e += THREADS) traj_q[e] = p.ref_q[e];
    for (int e = tid; e < H * NU; e += THREADS) traj_u[e] = p.ref_u[e];
    for (int e = tid; e < H * ND; e += THREADS) nu[e] = 0.0;
  }
  __syncthreads();
  for (int e = tid; e < NQ; e += THREADS) {
    traj_q[e] = q0[(size_t)r * NQ + e];
    traj_q[NQ + e] = q1[(size_t)r * NQ + e];
  }
  __syncthreads();
  double* cq = p.cand_q + (size_t)r * (H + 2) * NQ;
  double* cu = p.cand_u + (size_t)r * H * NU;
  double* cnu = p.cand_nu + (size_t)r * H * ND;
  for (int e = tid; e < (H + 2) * NQ; e += THREADS) cq[e] = traj_q[e];
  for (int e = tid; e < H * NU; e += THREADS) cu[e] = traj_u[e];
  for (int e = tid; e < H * ND; e += THREADS) cnu[e] = nu[e];
  for (int e = tid; e < H * NTH; e += THREADS) {
    const int t = e / NTH, c = e % NTH;
    double v;
    if (c < NQ) v = traj_q[t * NQ + c];
    else if (c < 2 * NQ) v = traj_q[(t + 1) * NQ + c - NQ];
    else if (c < 2 * NQ + NU) v = traj_u[t * NU + c - 2 * NQ];
    else if (c < 2 * NQ + NU + NW) v = p.w[t * NW + c - 2 * NQ - NU];
    else if (c == 2 * NQ + NU + NW) v = p.call->mu;
    else v = p.call->h;
    p.theta[((size_t)t * R + r) * NTH + c] = v;
  }
  for (int e = tid; e < H * NQ; e += THREADS) {
    const int t = e / NQ, k = e % NQ;
    p.q2[((size_t)t * R + r) * NQ + k] = traj_q[(t + 2) * NQ + k];
  }
  for (int e = tid; e < D::NC; e += THREADS) p.alt[(size_t)r * D::NC + e] = alt ? alt[(size_t)r * D::NC + e] : 0.0;
  const bool on = (active == nullptr) || active[r] != 0;  // inactive rollouts (ended simulations) are not solved
  for (int t = tid; t < H; t += THREADS) p.knot[(size_t)t * R + r] = on ? p.window[t] : -1;
  if (tid == 0) {
    p.phase[r] = on ? NP_INIT : NP_DONE;
    if (!on) atomicSub(p.n_active, 1);
    else p.act_list[atomicAdd(&p.act_count[0], 1)] = r;  // list 0 is the first sweep's (par = 0 after the reset)
    p.alpha[r] = 1.0;
    p.beta[r] = p.beta_init;
    p.r_norm[r] = 0.0;
    p.ls_it[r] = 0;
    p.newton_it[r] = 0;
    p.sweeps[r] = 0;
  }
}

}  // namespace cimpc
