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

struct NewtonParams {
  int R, H;
  // shared by every rollout (they track the same gait in lock-step, policy.jl:100-107, 131)
  const double* ref_q;    // (H+2) × nq   `ref_traj.q[1:H+2]`
  const double* ref_u;    // H × nu
  const double* w;        // H × nw       disturbances in θ (zeros in the reference's MPC)
  const int32_t* window;  // H            knot of every stage
  const double* obj_q;    // H × nq       diagonal of obj.q[t]
  const double* obj_u;    // H × nu
  double mu, h, kappa;
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
};

template <class D>
struct NewtonSmem {
  static constexpr int NQ = D::NQ, NU = D::NU, ND = D::ND, NCOL = D::NCOL;
  static_assert(D::MODE == 0, "device Newton: :configuration mode");
  static constexpr int KD = 3 * ND - 1;   // half bandwidth of the block-pentadiagonal Schur complement
  static constexpr int LDB = KD + 1;
  __host__ __device__ static constexpr int doubles(int H) {
    return H * ND * NCOL            // dz of the H stages
           + LDB * H * ND           // banded Y
           + (H + 2) * NQ + H * NU  // candidate q, u
           + 3 * H * ND             // ν_cand, g / Δν, d
           + H * (NU + NQ)          // r_x
           + 2 * H * (NU + NQ)      // Q⁻¹ (as vectors), v = Q⁻¹ r_x
           + 64;
  }
};

template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_step_kernel(const NewtonParams p) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NCOL = D::NCOL, NZ = D::NZ, NTH = D::NTH;
  constexpr int NR = NU + NQ;  // primal block of one stage: [u_t; q_{t+2}]
  using SM = NewtonSmem<D>;
  constexpr int KD = SM::KD, LDB = SM::LDB;
  const int r = blockIdx.x, tid = threadIdx.x, H = p.H, R = p.R;
  if (r >= R) return;
  int phase = p.phase[r];
  if (phase == NP_DONE) return;
  const int N = H * ND;  // dual dimension

  extern __shared__ __align__(16) double sm[];
  double* DZ = sm;                       // [t][col][row]
  double* Yb = DZ + H * ND * NCOL;       // banded lower: Yb[(i-j) + j*LDB]
  double* cq = Yb + LDB * N;             // (H+2)×NQ
  double* cu = cq + (H + 2) * NQ;        // H×NU
  double* cnu = cu + H * NU;             // H×ND
  double* gv = cnu + H * ND;             // H×ND   rhs, then Δν
  double* dv = gv + H * ND;              // H×ND   d_t
  double* rx = dv + H * ND;              // H×NR   [r_u; r_q] per stage
  double* qi = rx + H * NR;              // H×NR   Q⁻¹
  double* vx = qi + H * NR;              // H×NR   Q⁻¹ r_x
  double* red = vx + H * NR;             // reduction scratch (64)
  __shared__ int s_action;               // 0: re-evaluate at a smaller α, 1: solve for a new direction, 2: finished
  __shared__ double s_alpha;

  const double* cand_q = p.cand_q + (size_t)r * (H + 2) * NQ;
  const double* cand_u = p.cand_u + (size_t)r * H * NU;
  const double* cand_nu = p.cand_nu + (size_t)r * H * ND;
  for (int e = tid; e < (H + 2) * NQ; e += THREADS) cq[e] = cand_q[e];
  for (int e = tid; e < H * NU; e += THREADS) cu[e] = cand_u[e];
  for (int e = tid; e < H * ND; e += THREADS) cnu[e] = cand_nu[e];
  for (int t = 0; t < H; ++t) {
    const double* src = p.dz + ((size_t)t * R + r) * (ND * NCOL);
    for (int e = tid; e < ND * NCOL; e += THREADS) DZ[t * ND * NCOL + e] = src[e];
  }
  for (int e = tid; e < H * NR; e += THREADS) {
    const int t = e / NR, c = e % NR;
    qi[e] = 1.0 / (c < NU ? p.obj_u[t * NU + c] : p.obj_q[t * NQ + c - NU]);
  }
  __syncthreads();
  // d_t = z*_t[1:nq] − q_{t+2}   (implicit_dynamics.jl:180-182)
  for (int e = tid; e < H * ND; e += THREADS) {
    const int t = e / ND, i = e % ND;
    dv[e] = p.z[((size_t)t * R + r) * NZ + i] - cq[(t + 2) * NQ + i];
  }
  // residual!  (newton_residual.jl:113-138), primal rows
  for (int e = tid; e < H * NR; e += THREADS) {
    const int t = e / NR, c = e % NR;
    double acc;
    if (c < NU) {  // u_t:  obj.u (u − u_ref) + δu1_tᵀ ν_t
      acc = p.obj_u[t * NU + c] * (cu[t * NU + c] - p.ref_u[t * NU + c]);
      const double* col = DZ + t * ND * NCOL + (2 * NQ + c) * ND;
      for (int i = 0; i < ND; ++i) acc = fma(col[i], cnu[t * ND + i], acc);
    } else {  // q_{t+2}: obj.q (q − q_ref) − ν_t + δq1_{t+1}ᵀ ν_{t+1} + δq0_{t+2}ᵀ ν_{t+2}
      const int k = c - NU;
      acc = p.obj_q[t * NQ + k] * (cq[(t + 2) * NQ + k] - p.ref_q[(t + 2) * NQ + k]) - cnu[t * ND + k];
      if (t + 1 < H) {
        const double* col = DZ + (t + 1) * ND * NCOL + (NQ + k) * ND;
        for (int i = 0; i < ND; ++i) acc = fma(col[i], cnu[(t + 1) * ND + i], acc);
      }
      if (t + 2 < H) {
        const double* col = DZ + (t + 2) * ND * NCOL + (k)*ND;
        for (int i = 0; i < ND; ++i) acc = fma(col[i], cnu[(t + 2) * ND + i], acc);
      }
    }
    rx[e] = acc;
  }
  __syncthreads();
  // r_cand = ‖res‖₁  (newton.jl:198, 241) — fixed-order reduction: deterministic
  {
    double part = 0.0;
    for (int e = tid; e < H * NR; e += THREADS) part += fabs(rx[e]);
    for (int e = tid; e < H * ND; e += THREADS) part += fabs(dv[e]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
  }
  __syncthreads();
  if (tid == 0) {
    double r_cand = 0.0;
    for (int wgt = 0; wgt < THREADS / 32; ++wgt) r_cand += red[wgt];
    const int len = H * (NR + ND);
    double r_norm = p.r_norm[r], alpha = p.alpha[r], beta = p.beta[r];
    int ls = p.ls_it[r], l = p.newton_it[r];
    p.sweeps[r] += 1;
    int action;
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
    s_action = action | (accept ? 4 : 0);
    s_alpha = alpha;
    p.r_norm[r] = r_norm;
    p.alpha[r] = alpha;
    p.beta[r] = beta;
    p.ls_it[r] = ls;
    p.newton_it[r] = l;
  }
  __syncthreads();
  const int action = s_action & 3;
  const bool accept = (s_action & 4) != 0;
  const double alpha_acc = s_alpha;
  double* traj_q = p.traj_q + (size_t)r * (H + 2) * NQ;
  double* traj_u = p.traj_u + (size_t)r * H * NU;
  double* nu = p.nu + (size_t)r * H * ND;
  double* delta = p.delta + (size_t)r * H * (NR + ND);
  double* cq_g = p.cand_q + (size_t)r * (H + 2) * NQ;
  double* cu_g = p.cand_u + (size_t)r * H * NU;
  double* cnu_g = p.cand_nu + (size_t)r * H * ND;

  if (accept) {  // update_traj!(traj, traj, ν, ν, Δ, α)  (newton.jl:273)
    for (int e = tid; e < H * NQ; e += THREADS) {
      const int t = e / NQ, k = e % NQ;
      traj_q[(t + 2) * NQ + k] -= alpha_acc * delta[t * (NR + ND) + NU + k];
    }
    for (int e = tid; e < H * NU; e += THREADS) {
      const int t = e / NU, k = e % NU;
      traj_u[e] -= alpha_acc * delta[t * (NR + ND) + k];
    }
    for (int e = tid; e < H * ND; e += THREADS) {
      const int t = e / ND, k = e % ND;
      nu[e] -= alpha_acc * delta[t * (NR + ND) + NR + k];
    }
    __syncthreads();
  }

  if (action == 2) {  // finished: switch the rollout's subproblems off
    for (int t = tid; t < H; t += THREADS) p.knot[(size_t)t * R + r] = -1;
    if (tid == 0) {
      p.phase[r] = NP_DONE;
      atomicSub(p.n_active, 1);
    }
    return;
  }

  double alpha_next;
  if (action == 1) {
    // ---------------- jacobian! + linear_solve! through the dual Schur complement ----------------
    const double rho = (double)H * p.beta[r] * p.kappa;
    for (int e = tid; e < LDB * N; e += THREADS) Yb[e] = 0.0;
    for (int e = tid; e < H * NR; e += THREADS) vx[e] = qi[e] * rx[e];
    __syncthreads();
    // Y blocks (t,t), (t,t−1), (t,t−2); element (a,b) of block (t,s) lives at row t·nd+a, column s·nd+b
    const int per = 3 * ND * ND;
    for (int e = tid; e < H * per; e += THREADS) {
      const int t = e / per, rem = e % per, which = rem / (ND * ND), a = (rem % (ND * ND)) % ND, b = (rem % (ND * ND)) / ND;
      const int s = t - which;
      if (s < 0) continue;
      const int i = t * ND + a, j = s * ND + b;
      if (i < j) continue;  // lower triangle only
      const double* Zt = DZ + t * ND * NCOL;
      double acc = 0.0;
      if (which == 0) {
        for (int k = 0; k < NU; ++k) acc = fma(Zt[(2 * NQ + k) * ND + a] * qi[t * NR + k], Zt[(2 * NQ + k) * ND + b], acc);
        if (a == b) acc += qi[t * NR + NU + a] + rho;
        if (t >= 1)
          for (int k = 0; k < NQ; ++k)
            acc = fma(Zt[(NQ + k) * ND + a] * qi[(t - 1) * NR + NU + k], Zt[(NQ + k) * ND + b], acc);
        if (t >= 2)
          for (int k = 0; k < NQ; ++k) acc = fma(Zt[k * ND + a] * qi[(t - 2) * NR + NU + k], Zt[k * ND + b], acc);
      } else if (which == 1) {
        // shared variables of rows t and t−1: q_{t+1} (δq1_t vs −I) and q_t (δq0_t vs δq1_{t−1})
        acc = -Zt[(NQ + b) * ND + a] * qi[(t - 1) * NR + NU + b];
        if (t >= 2) {
          const double* Zs = DZ + (t - 1) * ND * NCOL;
          for (int k = 0; k < NQ; ++k) acc = fma(Zt[k * ND + a] * qi[(t - 2) * NR + NU + k], Zs[(NQ + k) * ND + b], acc);
        }
      } else {
        acc = -Zt[b * ND + a] * qi[(t - 2) * NR + NU + b];  // q_t: δq0_t vs −I of row t−2
      }
      Yb[(i - j) + j * LDB] = acc;
    }
    // g = C Q⁻¹ r_x − r_ν
    for (int e = tid; e < H * ND; e += THREADS) {
      const int t = e / ND, a = e % ND;
      const double* Zt = DZ + t * ND * NCOL;
      double acc = -vx[t * NR + NU + a] - dv[e];
      for (int k = 0; k < NU; ++k) acc = fma(Zt[(2 * NQ + k) * ND + a], vx[t * NR + k], acc);
      if (t >= 1)
        for (int k = 0; k < NQ; ++k) acc = fma(Zt[(NQ + k) * ND + a], vx[(t - 1) * NR + NU + k], acc);
      if (t >= 2)
        for (int k = 0; k < NQ; ++k) acc = fma(Zt[k * ND + a], vx[(t - 2) * NR + NU + k], acc);
      gv[e] = acc;
    }
    __syncthreads();
    // banded Cholesky Y = L Lᵀ (right-looking, in place)
    for (int j = 0; j < N; ++j) {
      const int m = min(KD, N - 1 - j);
      const double dj = sqrt(Yb[j * LDB]);
      __syncthreads();
      if (tid == 0) Yb[j * LDB] = dj;
      const double inv = 1.0 / dj;
      for (int i = 1 + tid; i <= m; i += THREADS) Yb[i + j * LDB] *= inv;
      __syncthreads();
      for (int e = tid; e < m * m; e += THREADS) {
        const int i1 = 1 + e % m, i2 = 1 + e / m;
        if (i2 >= i1) Yb[(i2 - i1) + (j + i1) * LDB] = fma(-Yb[i2 + j * LDB], Yb[i1 + j * LDB], Yb[(i2 - i1) + (j + i1) * LDB]);
      }
      __syncthreads();
    }
    // L y = g, Lᵀ Δν = y  (one warp; the band fits 32 lanes + 1)
    if (tid < 32) {
      for (int j = 0; j < N; ++j) {
        const int m = min(KD, N - 1 - j);
        const double yj = gv[j] / Yb[j * LDB];
        __syncwarp();
        if (tid == 0) gv[j] = yj;
        for (int i = 1 + tid; i <= m; i += 32) gv[j + i] = fma(-Yb[i + j * LDB], yj, gv[j + i]);
        __syncwarp();
      }
      for (int j = N - 1; j >= 0; --j) {
        const int m = min(KD, N - 1 - j);
        double part = 0.0;
        for (int i = 1 + tid; i <= m; i += 32) part = fma(Yb[i + j * LDB], gv[j + i], part);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        __syncwarp();
        if (tid == 0) gv[j] = (gv[j] - part) / Yb[j * LDB];
        __syncwarp();
      }
    }
    __syncthreads();
    // Δx = Q⁻¹ (r_x − Cᵀ Δν);  Δ = [Δu, Δq, Δν] per stage
    for (int e = tid; e < H * NR; e += THREADS) {
      const int t = e / NR, c = e % NR;
      double acc = rx[e];
      if (c < NU) {
        const double* col = DZ + t * ND * NCOL + (2 * NQ + c) * ND;
        for (int i = 0; i < ND; ++i) acc = fma(-col[i], gv[t * ND + i], acc);
      } else {
        const int k = c - NU;
        acc += gv[t * ND + k];
        if (t + 1 < H) {
          const double* col = DZ + (t + 1) * ND * NCOL + (NQ + k) * ND;
          for (int i = 0; i < ND; ++i) acc = fma(-col[i], gv[(t + 1) * ND + i], acc);
        }
        if (t + 2 < H) {
          const double* col = DZ + (t + 2) * ND * NCOL + k * ND;
          for (int i = 0; i < ND; ++i) acc = fma(-col[i], gv[(t + 2) * ND + i], acc);
        }
      }
      delta[t * (NR + ND) + c] = qi[e] * acc;
    }
    for (int e = tid; e < H * ND; e += THREADS) delta[(e / ND) * (NR + ND) + NR + e % ND] = gv[e];
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
  for (int e = tid; e < 2 * NQ; e += THREADS) cq[e] = traj_q[e];
  for (int e = tid; e < H * NQ; e += THREADS) {
    const int t = e / NQ, k = e % NQ;
    cq[(t + 2) * NQ + k] = traj_q[(t + 2) * NQ + k] - alpha_next * delta[t * (NR + ND) + NU + k];
  }
  for (int e = tid; e < H * NU; e += THREADS) cu[e] = traj_u[e] - alpha_next * delta[(e / NU) * (NR + ND) + e % NU];
  for (int e = tid; e < H * ND; e += THREADS) cnu_g[e] = nu[e] - alpha_next * delta[(e / ND) * (NR + ND) + NR + e % ND];
  __syncthreads();
  for (int e = tid; e < (H + 2) * NQ; e += THREADS) cq_g[e] = cq[e];
  for (int e = tid; e < H * NU; e += THREADS) cu_g[e] = cu[e];
  // θ_t = [q_t; q_{t+1}; u_t; w_t; μ; h]  (update_θ!, trajectory.jl:67-82), cold start q2 = q_{t+2}
  for (int e = tid; e < H * NTH; e += THREADS) {
    const int t = e / NTH, c = e % NTH;
    double v;
    if (c < NQ) v = cq[t * NQ + c];
    else if (c < 2 * NQ) v = cq[(t + 1) * NQ + c - NQ];
    else if (c < 2 * NQ + NU) v = cu[t * NU + c - 2 * NQ];
    else if (c < 2 * NQ + NU + NW) v = p.w[t * NW + c - 2 * NQ - NU];
    else if (c == 2 * NQ + NU + NW) v = p.mu;
    else v = p.h;
    p.theta[((size_t)t * R + r) * NTH + c] = v;
  }
  for (int e = tid; e < H * NQ; e += THREADS) {
    const int t = e / NQ, k = e % NQ;
    p.q2[((size_t)t * R + r) * NQ + k] = cq[(t + 2) * NQ + k];
  }
}

// reset!  (newton.jl:130-167): traj ← ref (cold) ; q[1], q[2] ← q0, q1 ; candidate ← traj ; first sweep inputs.
template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_reset_kernel(const NewtonParams p, const double* __restrict__ q0,
                                                               const double* __restrict__ q1, int warm_start) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NTH = D::NTH;
  const int r = blockIdx.x, tid = threadIdx.x, H = p.H, R = p.R;
  if (r >= R) return;
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
    else if (c == 2 * NQ + NU + NW) v = p.mu;
    else v = p.h;
    p.theta[((size_t)t * R + r) * NTH + c] = v;
  }
  for (int e = tid; e < H * NQ; e += THREADS) {
    const int t = e / NQ, k = e % NQ;
    p.q2[((size_t)t * R + r) * NQ + k] = traj_q[(t + 2) * NQ + k];
  }
  for (int t = tid; t < H; t += THREADS) p.knot[(size_t)t * R + r] = p.window[t];
  if (tid == 0) {
    p.phase[r] = NP_INIT;
    p.alpha[r] = 1.0;
    p.beta[r] = p.beta_init;
    p.r_norm[r] = 0.0;
    p.ls_it[r] = 0;
    p.newton_it[r] = 0;
    p.sweeps[r] = 0;
  }
}

}  // namespace cimpc
