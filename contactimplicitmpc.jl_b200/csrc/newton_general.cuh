// Batched `newton_solve!` (src/controller/newton.jl:169-288) — GENERAL variant: both modes of `ImplicitTrajectory`
// (:configuration, :configurationforce; implicit_dynamics.jl:37-46) and both objectives (`TrackingObjective`,
// `TrackingVelocityObjective`; src/controller/objective.jl:3-47).  The (:configuration, TrackingObjective) case of the
// Monte-Carlo benchmark keeps its specialised kernel (newton_kernel.cuh); this one is what e.g. the flamingo policy of
// examples/flamingo/flat.jl and test/controller/mpc_flamingo.jl needs.  Same state machine, same host protocol.
//
// KKT solve.  The reference factorises  R = [Q Cᵀ; C −ρI]  (newton_jacobian.jl:148-198) with
//     Q = blockdiag(Qu_t, Qγb_t, Qq_t) + the velocity couplings ±v_t between consecutive q (hessian!, :221-248),
//     C row block t:  δu1_t on u_t,  δq1_t on q_{t+1},  δq0_t on q_t,  −I on (q_{t+2}, γ_t, b_t).
// Here the same linear system is solved through an augmented dual Schur complement:
//   * force rows.  Every example of the reference weights γ, b with 1e-100.  With a numerically zero
//     weight the γb rows of the system give  Δν^y_t = −r_y,t  exactly; its coupling into the (u, q) rows moves to the
//     right-hand side and  Δy_t = δu1_t[y] Δu_t + δq1_t[y] Δx_{t−1} + δq0_t[y] Δx_{t−2} − ρ Δν^y_t − d^y_t  is
//     backed out afterwards.  (The host entry point refuses weights that are not negligible against Qq, Qu.)
//   * velocity cost.  ½ Σ v_t ‖x_t − x_{t−1}‖² is written with slack s_t = x_t − x_{t−1} and multiplier λ_t, which
//     keeps the primal Hessian diagonal:  P = diag(Qu, Qq, V).  The dual unknowns of stage t are μ_t = [Δν^q_t; Δλ_t]
//     (block size NB = 2 nq, or nq without the velocity cost) and
//         Y = C̃ P⁻¹ C̃ᵀ + diag(ρ I, V⁻¹)      is block-pentadiagonal SPD,
//     factorised by the same sliding-window block Cholesky as the specialised kernel.  Eliminating s, λ gives back
//     exactly the reference's system, so Δ equals R \ r to round-off (prototype vs dense solve: 1e-11; tests/test_gpu_newton.py).
#pragma once
#include <cstdint>

#include "dims.cuh"
#include "newton_kernel.cuh"

namespace cimpc {

template <class D, bool VEL>
struct NewtonGSmem {
  static constexpr int NQ = D::NQ, NU = D::NU, ND = D::ND, NYD = D::NYD, NCOL = D::NCOL;
  static constexpr int NB = VEL ? 2 * NQ : NQ;  // dual block of one stage: [ν^q; λ]
  static constexpr int BS = NB * NB;
  static constexpr int NRX = NU + NQ;           // primal block of the reduced system: [u_t; x_t = q_{t+2}]
  static constexpr int DL = NU + NQ + ND + NYD; // Δ of one stage: [Δu; Δx; Δν (nd); Δy]
  __host__ __device__ static constexpr int per_warp(int H) {
    return (H + 2) * NQ + H * NU + H * ND + H * NYD   // candidate q, u, ν, y
           + H * NB + H * ND + 2 * H * NRX + H * NYD  // g → μ ; d ; r_x ; Q⁻¹ ; r_y → Δν^y
           + (VEL ? H * NQ : 0)                       // V⁻¹
           + 6 * BS + 3 * ND * NCOL + 3 * NB + 8;     // factor blocks, δz window, Cholesky column (×2) + reciprocal diagonal
  }
  __host__ __device__ static constexpr size_t l_doubles(int H) { return (size_t)H * 3 * BS; }
};

template <class D, bool VEL, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_step_general_kernel(const NewtonParams p, double* __restrict__ lscratch) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NYD = D::NYD, NCOL = D::NCOL, NZ = D::NZ, NTH = D::NTH;
  using SM = NewtonGSmem<D, VEL>;
  constexpr int NB = SM::NB, BS = SM::BS, NRX = SM::NRX, DL = SM::DL;
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
  double* cy = cnu + H * ND;            // H×NYD   [γ; b]
  double* gv = cy + H * NYD;            // H×NB    rhs → y → μ = [Δν^q; Δλ]
  double* dv = gv + H * NB;             // H×ND    d_t
  double* rx = dv + H * ND;             // H×NRX   [r_u; r_x] per stage
  double* qi = rx + H * NRX;            // H×NRX   Q⁻¹ (diagonal)
  double* ry = qi + H * NRX;            // H×NYD   r_y, then Δν^y
  double* vi = ry + H * NYD;            // H×NQ    V⁻¹ (VEL)
  double* blk = vi + (VEL ? H * NQ : 0);  // 6 factor blocks
  double* zwin = blk + 6 * BS;          // δz of stages t, t+1, t+2 (ring of 3)
  double* colb = zwin + 3 * ND * NCOL;  // 2×NB  column of the Cholesky step in flight (double-buffered by parity)
  double* rdg = colb + 2 * NB;          // NB    reciprocals of the diagonal of L_tt

  const double* cand_q = p.cand_q + (size_t)r * (H + 2) * NQ;
  const double* cand_u = p.cand_u + (size_t)r * H * NU;
  const double* cand_nu = p.cand_nu + (size_t)r * H * ND;
  for (int e = lane; e < (H + 2) * NQ; e += 32) cq[e] = cand_q[e];
  for (int e = lane; e < H * NU; e += 32) cu[e] = cand_u[e];
  for (int e = lane; e < H * ND; e += 32) cnu[e] = cand_nu[e];
  if constexpr (NYD > 0) {
    const double* cand_y = p.cand_y + (size_t)r * H * NYD;
    for (int e = lane; e < H * NYD; e += 32) cy[e] = cand_y[e];
  }
  // δz of stage t: element (row a of nd, column c) — the δq0 | δq1 | δu1 views (implicit_dynamics.jl:82-86)
  auto DZ = [&](int t, int c, int a) -> double { return p.dz[(((size_t)t * R + r) * NCOL + c) * ND + a]; };
  for (int e = lane; e < H * NRX; e += 32) {
    const int t = e / NRX, c = e % NRX;
    qi[e] = 1.0 / (c < NU ? p.obj_u[t * NU + c] : p.obj_q[t * NQ + c - NU]);
  }
  if constexpr (VEL)
    for (int e = lane; e < H * NQ; e += 32) vi[e] = 1.0 / p.obj_v[e];
  __syncwarp();
  auto QIu = [&](int t, int k) -> double { return qi[t * NRX + k]; };
  auto QIq = [&](int t, int k) -> double { return qi[t * NRX + NU + k]; };

  // d_t = z*_t[1:nd] − [q_{t+2}; γ_t; b_t]   (implicit_dynamics.jl:180-190)
  for (int e = lane; e < H * ND; e += 32) {
    const int t = e / ND, i = e % ND;
    const double cur = (i < NQ) ? cq[(t + 2) * NQ + i] : cy[t * NYD + (i - NQ)];
    dv[e] = p.z[((size_t)t * R + r) * NZ + i] - cur;
  }
  // residual! + gradient!  (newton_residual.jl:113-138, 178-281), primal rows
  for (int e = lane; e < H * NRX; e += 32) {
    const int t = e / NRX, c = e % NRX;
    double acc;
    if (c < NU) {  // u_t:  obj.u (u − u_ref) + δu1_tᵀ ν_t
      acc = p.obj_u[t * NU + c] * (cu[t * NU + c] - p.ref_u[t * NU + c]);
      for (int i = 0; i < ND; ++i) acc = fma(DZ(t, 2 * NQ + c, i), cnu[t * ND + i], acc);
    } else {  // q_{t+2}: obj.q (q − q_ref) − ν^q_t + δq1_{t+1}ᵀ ν_{t+1} + δq0_{t+2}ᵀ ν_{t+2}  (+ velocity terms)
      const int k = c - NU;
      acc = p.obj_q[t * NQ + k] * (cq[(t + 2) * NQ + k] - p.ref_q[(t + 2) * NQ + k]) - cnu[t * ND + k];
      if (t + 1 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 1, NQ + k, i), cnu[(t + 1) * ND + i], acc);
      if (t + 2 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 2, k, i), cnu[(t + 2) * ND + i], acc);
      if constexpr (VEL) {  // v_t (q_{t+2} − q_{t+1}) on row t, −v_{t+1} (q_{t+3} − q_{t+2}) from stage t+1
        acc += p.obj_v[t * NQ + k] * (cq[(t + 2) * NQ + k] - cq[(t + 1) * NQ + k]);
        if (t + 1 < H) acc -= p.obj_v[(t + 1) * NQ + k] * (cq[(t + 3) * NQ + k] - cq[(t + 2) * NQ + k]);
      }
    }
    rx[e] = acc;
  }
  if constexpr (NYD > 0)  // (γ, b) rows: obj.γb (y − y_ref) − ν^y_t
    for (int e = lane; e < H * NYD; e += 32) {
      const int t = e / NYD, j = e % NYD;
      ry[e] = p.obj_y[e] * (cy[e] - p.ref_y[e]) - cnu[t * ND + NQ + j];
    }
  __syncwarp();
  // r_cand = ‖res‖₁  (newton.jl:198, 241) — fixed-order reduction: deterministic
  double r_cand = 0.0;
  for (int e = lane; e < H * NRX; e += 32) r_cand += fabs(rx[e]);
  for (int e = lane; e < H * NYD; e += 32) r_cand += fabs(ry[e]);
  for (int e = lane; e < H * ND; e += 32) r_cand += fabs(dv[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r_cand += __shfl_xor_sync(FULLM, r_cand, o);

  // ---- state machine (identical to newton_step_kernel) ----
  const int len = H * (NRX + NYD + ND);
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
      if (ls > 6) accept = true;
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
  if (action == 1 && (r_norm / (double)len < p.r_tol || l >= p.max_iter)) action = 2;
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
  double* traj_y = (NYD > 0) ? p.traj_y + (size_t)r * H * NYD : nullptr;
  double* nu = p.nu + (size_t)r * H * ND;
  double* delta = p.delta + (size_t)r * H * DL;
  double* cq_g = p.cand_q + (size_t)r * (H + 2) * NQ;
  double* cu_g = p.cand_u + (size_t)r * H * NU;
  double* cnu_g = p.cand_nu + (size_t)r * H * ND;

  if (accept) {  // update_traj!(traj, traj, ν, ν, Δ, α)  (newton.jl:273)
    for (int e = lane; e < H * NQ; e += 32) traj_q[(e / NQ + 2) * NQ + e % NQ] -= alpha_acc * delta[(e / NQ) * DL + NU + e % NQ];
    for (int e = lane; e < H * NU; e += 32) traj_u[e] -= alpha_acc * delta[(e / NU) * DL + e % NU];
    for (int e = lane; e < H * ND; e += 32) nu[e] -= alpha_acc * delta[(e / ND) * DL + NRX + e % ND];
    if constexpr (NYD > 0)
      for (int e = lane; e < H * NYD; e += 32) traj_y[e] -= alpha_acc * delta[(e / NYD) * DL + NRX + ND + e % NYD];
    __syncwarp();
  }

  if (action == 2) {
    for (int t = lane; t < H; t += 32) p.knot[(size_t)t * R + r] = -1;
    if (lane == 0) {
      p.phase[r] = NP_DONE;
      atomicSub(p.n_active, 1);
    }
    return;
  }

  double alpha_next;
  if (action == 1) {
    const double rho = (double)H * beta * p.kappa;
    double* Ls = lscratch + (size_t)r * SM::l_doubles(H);
    if constexpr (NYD > 0) {
      // zero-weight force rows: Δν^y = −r_y; move its coupling into the (u, x) right-hand sides
      for (int e = lane; e < H * NYD; e += 32) ry[e] = -ry[e];
      __syncwarp();
      for (int e = lane; e < H * NRX; e += 32) {
        const int t = e / NRX, c = e % NRX;
        double acc = 0.0;
        if (c < NU) {
          for (int j = 0; j < NYD; ++j) acc = fma(DZ(t, 2 * NQ + c, NQ + j), ry[t * NYD + j], acc);
        } else {
          const int k = c - NU;
          if (t + 1 < H)
            for (int j = 0; j < NYD; ++j) acc = fma(DZ(t + 1, NQ + k, NQ + j), ry[(t + 1) * NYD + j], acc);
          if (t + 2 < H)
            for (int j = 0; j < NYD; ++j) acc = fma(DZ(t + 2, k, NQ + j), ry[(t + 2) * NYD + j], acc);
        }
        rx[e] -= acc;
      }
      __syncwarp();
    }
    // g = C̃ P⁻¹ r_primal − r_dual
    for (int e = lane; e < H * NB; e += 32) {
      const int t = e / NB, a = e % NB;
      double acc;
      if (a < NQ) {
        acc = -rx[t * NRX + NU + a] * QIq(t, a) - dv[t * ND + a];
        for (int k = 0; k < NU; ++k) acc = fma(DZ(t, 2 * NQ + k, a), rx[t * NRX + k] * QIu(t, k), acc);
        if (t >= 1)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, NQ + k, a), rx[(t - 1) * NRX + NU + k] * QIq(t - 1, k), acc);
        if (t >= 2)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, k, a), rx[(t - 2) * NRX + NU + k] * QIq(t - 2, k), acc);
      } else {
        const int k = a - NQ;
        acc = rx[t * NRX + NU + k] * QIq(t, k);
        if (t >= 1) acc -= rx[(t - 1) * NRX + NU + k] * QIq(t - 1, k);
      }
      gv[e] = acc;
    }
    __syncwarp();
    double *A0 = blk, *A1 = blk + BS, *A2 = blk + 2 * BS, *P1 = blk + 3 * BS, *P2 = blk + 4 * BS, *Q2 = blk + 5 * BS;
    auto load_stage = [&](int s_) {
      const double* src = p.dz + ((size_t)s_ * R + r) * (ND * NCOL);
      double* dst = zwin + (s_ % 3) * (ND * NCOL);
      for (int e = lane; e < ND * NCOL; e += 32) dst[e] = src[e];
    };
    auto ZW = [&](int s_, int c, int a) -> double { return zwin[(s_ % 3) * (ND * NCOL) + c * ND + a]; };
    auto Aq = [&](int s_, int a, int k) -> double { return ZW(s_, NQ + k, a); };      // δq1_s[a][k]
    auto Bq = [&](int s_, int a, int k) -> double { return ZW(s_, k, a); };           // δq0_s[a][k]
    auto Uq = [&](int s_, int a, int k) -> double { return ZW(s_, 2 * NQ + k, a); };  // δu1_s[a][k]
    load_stage(0);
    if (H > 1) load_stage(1);
    for (int t = 0; t < H; ++t) {
      if (t + 2 < H) load_stage(t + 2);
      __syncwarp();
      // ---- block column t of Y (entries derived in the header comment) minus the pending Cholesky updates ----
      for (int e = lane; e < BS; e += 32) {
        const int a = e % NB, b = e / NB;
        const bool aN = a < NQ, bN = b < NQ;
        const int a_ = aN ? a : a - NQ, b_ = bN ? b : b - NQ;
        double acc = 0.0;
        if (a >= b) {  // (t,t), lower triangle
          if (aN) {    // then b is an ν row too
            for (int k = 0; k < NU; ++k) acc = fma(Uq(t, a, k) * QIu(t, k), Uq(t, b, k), acc);
            if (a == b) acc += QIq(t, a) + rho;
            if (t >= 1)
              for (int k = 0; k < NQ; ++k) acc = fma(Aq(t, a, k) * QIq(t - 1, k), Aq(t, b, k), acc);
            if (t >= 2)
              for (int k = 0; k < NQ; ++k) acc = fma(Bq(t, a, k) * QIq(t - 2, k), Bq(t, b, k), acc);
          } else if (bN) {  // λ row × ν row
            if (a_ == b) acc -= QIq(t, b);
            if (t >= 1) acc -= QIq(t - 1, a_) * Aq(t, b, a_);
          } else if (a == b) {  // λ × λ
            acc = QIq(t, a_) + (t >= 1 ? QIq(t - 1, a_) : 0.0) + vi[t * NQ + a_];
          }
          if (t >= 1)
            for (int k = 0; k < NB; ++k) acc = fma(-P1[a + k * NB], P1[b + k * NB], acc);
          if (t >= 2)
            for (int k = 0; k < NB; ++k) acc = fma(-Q2[a + k * NB], Q2[b + k * NB], acc);
        }
        A0[e] = acc;
        if (t + 1 < H) {  // (t+1, t): row a of stage t+1, column b of stage t
          double c1 = 0.0;
          if (aN && bN) {
            c1 = -Aq(t + 1, a, b) * QIq(t, b);
            if (t >= 1)
              for (int k = 0; k < NQ; ++k) c1 = fma(Bq(t + 1, a, k) * QIq(t - 1, k), Aq(t, b, k), c1);
          } else if (aN) {  // ν_{t+1} × λ_t
            c1 = Aq(t + 1, a, b_) * QIq(t, b_);
            if (t >= 1) c1 -= Bq(t + 1, a, b_) * QIq(t - 1, b_);
          } else if (bN) {  // λ_{t+1} × ν_t
            if (a_ == b) c1 = QIq(t, b);
          } else {          // λ_{t+1} × λ_t
            if (a_ == b_) c1 = -QIq(t, a_);
          }
          if (t >= 1)
            for (int k = 0; k < NB; ++k) c1 = fma(-P2[a + k * NB], P1[b + k * NB], c1);
          A1[e] = c1;
        }
        if (t + 2 < H) {  // (t+2, t)
          double c2 = 0.0;
          if (aN) c2 = (bN ? -1.0 : 1.0) * Bq(t + 2, a, b_) * QIq(t, b_);
          A2[e] = c2;
        }
      }
      __syncwarp();
      // ---- potrf: A0 = L Lᵀ with the rows in registers (row i of lane i, and row 32 + i when NB > 32), one published
      //      column per step; trsm rows and the forward substitution in registers / by shuffle (as newton_kernel.cuh) ----
      {
        constexpr int RPL = (NB + 31) / 32;  // rows per lane
        double rw[RPL][NB];
#pragma unroll
        for (int q_ = 0; q_ < RPL; ++q_)
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            const int i = lane + 32 * q_;
            rw[q_][c] = (i < NB && c <= i) ? A0[i + c * NB] : 0.0;
          }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
          const double ajj = __shfl_sync(FULLM, rw[j / 32][j], j % 32);
          const double inv = rsqrt(ajj);
          double* cb = colb + (j & 1) * NB;
#pragma unroll
          for (int q_ = 0; q_ < RPL; ++q_) {
            const int i = lane + 32 * q_;
            const double lij = (i == j) ? ajj * inv : rw[q_][j] * inv;
            rw[q_][j] = lij;
            if (i >= j && i < NB) cb[i] = lij;
          }
          if (lane == 0) rdg[j] = inv;
          __syncwarp();
#pragma unroll
          for (int q_ = 0; q_ < RPL; ++q_) {
            const int i = lane + 32 * q_;
            const double lij = rw[q_][j];
#pragma unroll
            for (int c = j + 1; c < NB; ++c)
              if (c <= i && i < NB) rw[q_][c] = fma(-lij, cb[c], rw[q_][c]);
          }
        }
#pragma unroll
        for (int q_ = 0; q_ < RPL; ++q_)
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            const int i = lane + 32 * q_;
            if (i < NB && c <= i) A0[i + c * NB] = rw[q_][c];
          }
      }
      __syncwarp();
      {
        const int nrows = (t + 1 < H ? NB : 0) + (t + 2 < H ? NB : 0);
        for (int row = lane; row < nrows; row += 32) {
          double* X = (row < NB) ? (A1 + row) : (A2 + row - NB);
          double x[NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) x[c] = X[c * NB];
#pragma unroll
          for (int c = 0; c < NB; ++c) {
            double sacc = x[c];
#pragma unroll
            for (int k = 0; k < c; ++k) sacc = fma(-x[k], A0[c + k * NB], sacc);
            x[c] = sacc * rdg[c];
          }
#pragma unroll
          for (int c = 0; c < NB; ++c) X[c * NB] = x[c];
        }
        __syncwarp();
        for (int i = lane; i < NB; i += 32) {
          double s = gv[t * NB + i];
          if (t >= 1)
            for (int k = 0; k < NB; ++k) s = fma(-P1[i + k * NB], gv[(t - 1) * NB + k], s);
          if (t >= 2)
            for (int k = 0; k < NB; ++k) s = fma(-Q2[i + k * NB], gv[(t - 2) * NB + k], s);
          gv[t * NB + i] = s;
        }
        __syncwarp();
        for (int c = 0; c < NB; ++c) {
          const double yc = gv[t * NB + c] * rdg[c];
          __syncwarp();
          if (lane == 0) gv[t * NB + c] = yc;
          for (int i = c + 1 + lane; i < NB; i += 32) gv[t * NB + i] = fma(-A0[i + c * NB], yc, gv[t * NB + i]);
          __syncwarp();
        }
      }
      for (int e = lane; e < BS; e += 32) {
        Ls[(size_t)(3 * t) * BS + e] = A0[e];
        if (t + 1 < H) Ls[(size_t)(3 * t + 1) * BS + e] = A1[e];
        if (t + 2 < H) Ls[(size_t)(3 * t + 2) * BS + e] = A2[e];
      }
      __syncwarp();
      double* oldQ2 = Q2;
      Q2 = P2;
      P2 = A2;
      double* oldP1 = P1;
      P1 = A1;
      A1 = oldP1;
      A2 = oldQ2;
    }
    // ---- backward: μ_t = L_tt⁻ᵀ (y_t − L_{t+1,t}ᵀ μ_{t+1} − L_{t+2,t}ᵀ μ_{t+2}) ----
    for (int t = H - 1; t >= 0; --t) {
      const double* L0 = Ls + (size_t)(3 * t) * BS;
      const double* L1 = L0 + BS;
      const double* L2 = L0 + 2 * BS;
      for (int e = lane; e < BS; e += 32) A0[e] = L0[e];
      for (int i = lane; i < NB; i += 32) rdg[i] = 1.0 / L0[i + i * NB];
      for (int i = lane; i < NB; i += 32) {
        double s = gv[t * NB + i];
        if (t + 1 < H)
          for (int k = 0; k < NB; ++k) s = fma(-L1[k + i * NB], gv[(t + 1) * NB + k], s);
        if (t + 2 < H)
          for (int k = 0; k < NB; ++k) s = fma(-L2[k + i * NB], gv[(t + 2) * NB + k], s);
        gv[t * NB + i] = s;
      }
      __syncwarp();
      for (int c = NB - 1; c >= 0; --c) {
        const double xc = gv[t * NB + c] * rdg[c];
        __syncwarp();
        if (lane == 0) gv[t * NB + c] = xc;
        for (int i = lane; i < c; i += 32) gv[t * NB + i] = fma(-A0[c + i * NB], xc, gv[t * NB + i]);
        __syncwarp();
      }
    }
    // Δ[u, x] = P⁻¹ (r − C̃ᵀ μ)
    for (int e = lane; e < H * NRX; e += 32) {
      const int t = e / NRX, c = e % NRX;
      double acc = rx[e];
      if (c < NU) {
        for (int i = 0; i < NQ; ++i) acc = fma(-DZ(t, 2 * NQ + c, i), gv[t * NB + i], acc);
        delta[t * DL + c] = QIu(t, c) * acc;
      } else {
        const int k = c - NU;
        acc += gv[t * NB + k];
        if (t + 1 < H)
          for (int i = 0; i < NQ; ++i) acc = fma(-DZ(t + 1, NQ + k, i), gv[(t + 1) * NB + i], acc);
        if (t + 2 < H)
          for (int i = 0; i < NQ; ++i) acc = fma(-DZ(t + 2, k, i), gv[(t + 2) * NB + i], acc);
        if constexpr (VEL) {
          acc -= gv[t * NB + NQ + k];
          if (t + 1 < H) acc += gv[(t + 1) * NB + NQ + k];
        }
        delta[t * DL + c] = QIq(t, k) * acc;
      }
    }
    for (int e = lane; e < H * NQ; e += 32) delta[(e / NQ) * DL + NRX + e % NQ] = gv[(e / NQ) * NB + e % NQ];
    if constexpr (NYD > 0) {
      __syncwarp();
      for (int e = lane; e < H * NYD; e += 32) {
        const int t = e / NYD, j = e % NYD;
        double acc = -rho * ry[e] - dv[t * ND + NQ + j];
        for (int c = 0; c < NU; ++c) acc = fma(DZ(t, 2 * NQ + c, NQ + j), delta[t * DL + c], acc);
        if (t >= 1)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, NQ + k, NQ + j), delta[(t - 1) * DL + NU + k], acc);
        if (t >= 2)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, k, NQ + j), delta[(t - 2) * DL + NU + k], acc);
        delta[t * DL + NRX + ND + j] = acc;        // Δy
        delta[t * DL + NRX + NQ + j] = ry[e];      // Δν^y
      }
    }
    __syncwarp();
    alpha_next = 1.0;
    if (lane == 0) {
      p.alpha[r] = 1.0;
      p.ls_it[r] = 0;
      p.phase[r] = NP_LS;
    }
  } else {
    alpha_next = alpha_acc;
  }

  // candidate = traj − α Δ  (update_traj!, newton_residual.jl:160-176), then the next sweep's inputs
  for (int e = lane; e < 2 * NQ; e += 32) cq[e] = traj_q[e];
  for (int e = lane; e < H * NQ; e += 32) {
    const int t = e / NQ, k = e % NQ;
    cq[(t + 2) * NQ + k] = traj_q[(t + 2) * NQ + k] - alpha_next * delta[t * DL + NU + k];
  }
  for (int e = lane; e < H * NU; e += 32) cu[e] = traj_u[e] - alpha_next * delta[(e / NU) * DL + e % NU];
  for (int e = lane; e < H * ND; e += 32) cnu_g[e] = nu[e] - alpha_next * delta[(e / ND) * DL + NRX + e % ND];
  if constexpr (NYD > 0) {
    double* cy_g = p.cand_y + (size_t)r * H * NYD;
    for (int e = lane; e < H * NYD; e += 32) cy_g[e] = traj_y[e] - alpha_next * delta[(e / NYD) * DL + NRX + ND + e % NYD];
  }
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
  if (lane == 0) p.act_list[(size_t)(cur ^ 1) * R + atomicAdd(&p.act_count[cur ^ 1], 1)] = r;
}

}  // namespace cimpc
