// Batched `newton_solve!` (src/controller/newton.jl:169-288) — DENSE-WEIGHT variant: per-stage full matrices `obj.q[t]`
// (e.g. `relative_state_cost`, src/dynamics/centroidal_quadruped/model.jl:168-183, the objective of
// examples/centroidal_quadruped/flat_trot.jl:37-42) and a `TrackingVelocityObjective` with non-zero `v_target`
// (src/controller/objective.jl:18-47; gradient newton_residual.jl:254-281).  :configuration mode.  Same state machine,
// same sweep protocol and the same block-pentadiagonal Cholesky as newton_general.cuh; what changes is the assembly.
//
// KKT solve.  R = [P Cᵀ; C −ρI] with P = blockdiag(Qu_t, Q_t [, V_t]) — Q_t dense SPD.  With Q_t = L_t L_tᵀ and
// E_t = L_t⁻ᵀ (upper triangular, formed on the host once per objective) the configuration block is whitened,
// x_t = E_t x̃_t, so that P̃ = I on x̃ and the dual Schur complement is a sum of plain outer products of the (whitened)
// coefficient rows
//     ν_s (dynamics rows of stage s):   on x̃_s: −E_s      on x̃_{s−1}: Ã_s = δq1_s E_{s−1}      on x̃_{s−2}: B̃_s = δq0_s E_{s−2}
//     λ_s (velocity slack rows, VEL):   on x̃_s: +E_s      on x̃_{s−1}: −E_{s−1}
//     Y[(i,a),(j,b)] = Σ_p Σ_k K_i,p[a,k] K_j,p[b,k]  (+ δu1 Qu⁻¹ δu1ᵀ + ρI on ν×ν,  + V⁻¹ on λ×λ)
// — E_t E_tᵀ = Q_t⁻¹, so for a diagonal Q_t this is the matrix of newton_general.cuh entry by entry.  Δ equals R \ r to
// round-off (tests/test_newton_kkt_math.py states the elimination in numpy for dense Q_t; tests/test_gpu_newton.py holds
// the kernel to the oracle's dense KKT solve).
#pragma once
#include <cstdint>

#include "dims.cuh"
#include "newton_kernel.cuh"

namespace cimpc {

template <class D, bool VEL>
struct NewtonDSmem {
  static constexpr int NQ = D::NQ, NU = D::NU, ND = D::ND, NCOL = D::NCOL;
  static_assert(D::MODE == 0, "dense-weight Newton: :configuration mode");
  static constexpr int NB = VEL ? 2 * NQ : NQ;  // dual block of one stage: [ν; λ]
  static constexpr int BS = NB * NB;
  static constexpr int NRX = NU + NQ;
  static constexpr int DL = NU + NQ + ND;       // Δ of one stage: [Δu; Δx; Δν]
  __host__ __device__ static constexpr int per_warp(int H) {
    return (H + 2) * NQ + H * NU + H * ND           // candidate q, u, ν
           + H * NB + H * ND + H * NRX + H * NU     // g → μ ; d ; r_x ; Qu⁻¹
           + 2 * H * NQ                             // whitened residual / work vector ; Q_t⁻¹ r_x
           + (VEL ? H * NQ : 0)                     // V⁻¹
           + 6 * BS + 3 * ND * NCOL + 3 * NB + 8;   // factor blocks, δz window (whitened in place), Cholesky column, diag
  }
  __host__ __device__ static constexpr size_t l_doubles(int H) { return (size_t)H * 3 * BS; }
};

// One coefficient row K_{s,p}[a, :] of the whitened constraint Jacobian: value(k) = sign · ptr[k · stride] (null: zero).
struct CoefRow {
  const double* ptr;
  int stride;
  double sign;
};

template <class D, bool VEL, int THREADS>
__global__ void __launch_bounds__(THREADS) newton_step_dense_kernel(const NewtonParams p, double* __restrict__ lscratch) {
  constexpr int NQ = D::NQ, NU = D::NU, NW = D::NW, ND = D::ND, NCOL = D::NCOL, NZ = D::NZ, NTH = D::NTH;
  using SM = NewtonDSmem<D, VEL>;
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
  double* gv = cnu + H * ND;            // H×NB    rhs → y → μ = [Δν; Δλ]
  double* dv = gv + H * NB;             // H×ND    d_t
  double* rx = dv + H * ND;             // H×NRX   [r_u; r_x] per stage
  double* qiu = rx + H * NRX;           // H×NU    Qu⁻¹ (diagonal)
  double* wv = qiu + H * NU;            // H×NQ    work: E_tᵀ r_x, later the back-substitution vector
  double* yv = wv + H * NQ;             // H×NQ    Q_t⁻¹ r_x
  double* vi = yv + H * NQ;             // H×NQ    V⁻¹ (VEL)
  double* blk = vi + (VEL ? H * NQ : 0);  // 6 factor blocks
  double* zwin = blk + 6 * BS;          // δz of stages t, t+1, t+2 (ring of 3), δq0 / δq1 whitened in place
  double* colb = zwin + 3 * ND * NCOL;  // 2×NB
  double* rdg = colb + 2 * NB;          // NB

  const double* cand_q = p.cand_q + (size_t)r * (H + 2) * NQ;
  const double* cand_u = p.cand_u + (size_t)r * H * NU;
  const double* cand_nu = p.cand_nu + (size_t)r * H * ND;
  for (int e = lane; e < (H + 2) * NQ; e += 32) cq[e] = cand_q[e];
  for (int e = lane; e < H * NU; e += 32) cu[e] = cand_u[e];
  for (int e = lane; e < H * ND; e += 32) cnu[e] = cand_nu[e];
  auto DZ = [&](int t, int c, int a) -> double { return p.dz[(((size_t)t * R + r) * NCOL + c) * ND + a]; };
  auto QD = [&](int t, int i, int j) -> double { return __ldg(p.obj_qd + ((size_t)t * NQ + j) * NQ + i); };  // Q_t[i][j]
  auto EM = [&](int t, int i, int j) -> double { return __ldg(p.obj_e + ((size_t)t * NQ + j) * NQ + i); };   // E_t[i][j]
  for (int e = lane; e < H * NU; e += 32) qiu[e] = 1.0 / p.obj_u[e];
  if constexpr (VEL)
    for (int e = lane; e < H * NQ; e += 32) vi[e] = 1.0 / p.obj_v[e];
  __syncwarp();
  auto QT = [&](int t, int k) -> double { return p.q_tgt ? p.q_tgt[t * NQ + k] : 0.0; };
  auto VT = [&](int t, int k) -> double { return p.v_tgt ? p.v_tgt[t * NQ + k] : 0.0; };

  // d_t = z*_t[1:nq] − q_{t+2}   (implicit_dynamics.jl:180-182)
  for (int e = lane; e < H * ND; e += 32) {
    const int t = e / ND, i = e % ND;
    dv[e] = p.z[((size_t)t * R + r) * NZ + i] - cq[(t + 2) * NQ + i];
  }
  // residual! + gradient!  (newton_residual.jl:113-138, 254-281), primal rows
  for (int e = lane; e < H * NRX; e += 32) {
    const int t = e / NRX, c = e % NRX;
    double acc;
    if (c < NU) {
      acc = p.obj_u[t * NU + c] * (cu[t * NU + c] - p.ref_u[t * NU + c]);
      for (int i = 0; i < ND; ++i) acc = fma(DZ(t, 2 * NQ + c, i), cnu[t * ND + i], acc);
    } else {
      const int k = c - NU;
      acc = -cnu[t * ND + k];
      for (int m = 0; m < NQ; ++m)  // obj.q[t] (q − (q_ref + q_target))
        acc = fma(QD(t, k, m), cq[(t + 2) * NQ + m] - (p.ref_q[(t + 2) * NQ + m] + QT(t, m)), acc);
      if (t + 1 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 1, NQ + k, i), cnu[(t + 1) * ND + i], acc);
      if (t + 2 < H)
        for (int i = 0; i < ND; ++i) acc = fma(DZ(t + 2, k, i), cnu[(t + 2) * ND + i], acc);
      if constexpr (VEL) {  // v_t (q_{t+2} − q_{t+1} − v_target_t) on row t, its negative from stage t+1 on row t
        acc += p.obj_v[t * NQ + k] * (cq[(t + 2) * NQ + k] - cq[(t + 1) * NQ + k] - VT(t, k));
        if (t + 1 < H) acc -= p.obj_v[(t + 1) * NQ + k] * (cq[(t + 3) * NQ + k] - cq[(t + 2) * NQ + k] - VT(t + 1, k));
      }
    }
    rx[e] = acc;
  }
  __syncwarp();
  double r_cand = 0.0;
  for (int e = lane; e < H * NRX; e += 32) r_cand += fabs(rx[e]);
  for (int e = lane; e < H * ND; e += 32) r_cand += fabs(dv[e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r_cand += __shfl_xor_sync(FULLM, r_cand, o);

  // ---- state machine (identical to newton_step_kernel) ----
  const int len = H * (NRX + ND);
  double r_norm = p.r_norm[r], alpha = p.alpha[r], beta = p.beta[r];
  int ls = p.ls_it[r], l = p.newton_it[r];
  int action;
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
  double* nu = p.nu + (size_t)r * H * ND;
  double* delta = p.delta + (size_t)r * H * DL;
  double* cq_g = p.cand_q + (size_t)r * (H + 2) * NQ;
  double* cu_g = p.cand_u + (size_t)r * H * NU;
  double* cnu_g = p.cand_nu + (size_t)r * H * ND;

  if (accept) {  // update_traj!(traj, traj, ν, ν, Δ, α)  (newton.jl:273)
    for (int e = lane; e < H * NQ; e += 32) traj_q[(e / NQ + 2) * NQ + e % NQ] -= alpha_acc * delta[(e / NQ) * DL + NU + e % NQ];
    for (int e = lane; e < H * NU; e += 32) traj_u[e] -= alpha_acc * delta[(e / NU) * DL + e % NU];
    for (int e = lane; e < H * ND; e += 32) nu[e] -= alpha_acc * delta[(e / ND) * DL + NRX + e % ND];
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
    // wv_t = E_tᵀ r_x,t ;  yv_t = E_t wv_t = Q_t⁻¹ r_x,t      (E upper triangular)
    for (int e = lane; e < H * NQ; e += 32) {
      const int t = e / NQ, k = e % NQ;
      double acc = 0.0;
      for (int m = 0; m <= k; ++m) acc = fma(EM(t, m, k), rx[t * NRX + NU + m], acc);
      wv[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < H * NQ; e += 32) {
      const int t = e / NQ, k = e % NQ;
      double acc = 0.0;
      for (int m = k; m < NQ; ++m) acc = fma(EM(t, k, m), wv[t * NQ + m], acc);
      yv[e] = acc;
    }
    __syncwarp();
    // g = C̃ P⁻¹ r_primal − r_dual
    for (int e = lane; e < H * NB; e += 32) {
      const int t = e / NB, a = e % NB;
      double acc;
      if (a < NQ) {
        acc = -yv[t * NQ + a] - dv[t * ND + a];
        for (int k = 0; k < NU; ++k) acc = fma(DZ(t, 2 * NQ + k, a), rx[t * NRX + k] * qiu[t * NU + k], acc);
        if (t >= 1)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, NQ + k, a), yv[(t - 1) * NQ + k], acc);
        if (t >= 2)
          for (int k = 0; k < NQ; ++k) acc = fma(DZ(t, k, a), yv[(t - 2) * NQ + k], acc);
      } else {
        const int k = a - NQ;
        acc = yv[t * NQ + k];
        if (t >= 1) acc -= yv[(t - 1) * NQ + k];
      }
      gv[e] = acc;
    }
    __syncwarp();
    double *A0 = blk, *A1 = blk + BS, *A2 = blk + 2 * BS, *P1 = blk + 3 * BS, *P2 = blk + 4 * BS, *Q2 = blk + 5 * BS;
    // δz window: stage s in slot s % 3; δq1_s ← δq1_s E_{s−1}, δq0_s ← δq0_s E_{s−2} in place (row a by lane a; E upper
    // triangular, so column k of the product only needs columns ≤ k of the row: walk k downwards)
    auto load_stage = [&](int s_) {
      const double* src = p.dz + ((size_t)s_ * R + r) * (ND * NCOL);
      double* dst = zwin + (s_ % 3) * (ND * NCOL);
      for (int e = lane; e < ND * NCOL; e += 32) dst[e] = src[e];
      __syncwarp();
      if (lane < ND) {
        const int a = lane;
        if (s_ >= 1)
          for (int k = NQ - 1; k >= 0; --k) {
            double acc = 0.0;
            for (int m = 0; m <= k; ++m) acc = fma(dst[(NQ + m) * ND + a], EM(s_ - 1, m, k), acc);
            dst[(NQ + k) * ND + a] = acc;
          }
        if (s_ >= 2)
          for (int k = NQ - 1; k >= 0; --k) {
            double acc = 0.0;
            for (int m = 0; m <= k; ++m) acc = fma(dst[m * ND + a], EM(s_ - 2, m, k), acc);
            dst[k * ND + a] = acc;
          }
      }
      __syncwarp();
    };
    auto ZWp = [&](int s_) -> const double* { return zwin + (s_ % 3) * (ND * NCOL); };
    // coefficient row of dual row `a` of stage s_ on the whitened configuration of stage p_
    auto coef = [&](int s_, int a, int p_) -> CoefRow {
      if (p_ < 0) return CoefRow{nullptr, 0, 0.0};
      if (a < NQ) {
        if (p_ == s_) return CoefRow{p.obj_e + (size_t)s_ * NQ * NQ + a, NQ, -1.0};
        if (p_ == s_ - 1) return CoefRow{ZWp(s_) + NQ * ND + a, ND, 1.0};
        return CoefRow{ZWp(s_) + a, ND, 1.0};
      }
      const int a_ = a - NQ;
      if (p_ == s_) return CoefRow{p.obj_e + (size_t)s_ * NQ * NQ + a_, NQ, 1.0};
      if (p_ == s_ - 1) return CoefRow{p.obj_e + (size_t)(s_ - 1) * NQ * NQ + a_, NQ, -1.0};
      return CoefRow{nullptr, 0, 0.0};
    };
    auto dotk = [&](const CoefRow& x, const CoefRow& y, double acc) -> double {
      if (x.ptr == nullptr || y.ptr == nullptr) return acc;
      const double sg = x.sign * y.sign;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll 2
      for (int k = 0; k + 1 < NQ; k += 2) {
        s0 = fma(x.ptr[k * x.stride], y.ptr[k * y.stride], s0);
        s1 = fma(x.ptr[(k + 1) * x.stride], y.ptr[(k + 1) * y.stride], s1);
      }
      if (NQ & 1) s0 = fma(x.ptr[(NQ - 1) * x.stride], y.ptr[(NQ - 1) * y.stride], s0);
      return fma(sg, s0 + s1, acc);
    };
    load_stage(0);
    if (H > 1) load_stage(1);
    for (int t = 0; t < H; ++t) {
      if (t + 2 < H) load_stage(t + 2);
      __syncwarp();
      const double* Zt = ZWp(t);
      for (int e = lane; e < BS; e += 32) {
        const int a = e % NB, b = e / NB;
        const bool aN = a < NQ, bN = b < NQ;
        double acc = 0.0;
        if (a >= b) {  // (t,t), lower triangle
          for (int dp = 0; dp <= 2; ++dp) acc = dotk(coef(t, a, t - dp), coef(t, b, t - dp), acc);
          if (aN && bN) {
            for (int k = 0; k < NU; ++k) acc = fma(Zt[(2 * NQ + k) * ND + a] * qiu[t * NU + k], Zt[(2 * NQ + k) * ND + b], acc);
            if (a == b) acc += rho;
          } else if (!aN && a == b) {
            if constexpr (VEL) acc += vi[t * NQ + a - NQ];
          }
          if (t >= 1)
            for (int k = 0; k < NB; ++k) acc = fma(-P1[a + k * NB], P1[b + k * NB], acc);
          if (t >= 2)
            for (int k = 0; k < NB; ++k) acc = fma(-Q2[a + k * NB], Q2[b + k * NB], acc);
        }
        A0[e] = acc;
        if (t + 1 < H) {  // (t+1, t): shared configurations x̃_t, x̃_{t−1}
          double c1 = 0.0;
          c1 = dotk(coef(t + 1, a, t), coef(t, b, t), c1);
          c1 = dotk(coef(t + 1, a, t - 1), coef(t, b, t - 1), c1);
          if (t >= 1)
            for (int k = 0; k < NB; ++k) c1 = fma(-P2[a + k * NB], P1[b + k * NB], c1);
          A1[e] = c1;
        }
        if (t + 2 < H) A2[e] = dotk(coef(t + 2, a, t), coef(t, b, t), 0.0);  // (t+2, t): shared x̃_t
      }
      __syncwarp();
      // ---- potrf / trsm / forward substitution: as newton_general.cuh ----
      {
        constexpr int RPL = (NB + 31) / 32;
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
    // Δu = Qu⁻¹ (r_u − δu1ᵀ Δν);  Δx_t = E_t E_tᵀ (r_x + Δν_t − δq1_{t+1}ᵀ Δν_{t+1} − δq0_{t+2}ᵀ Δν_{t+2} − Δλ_t + Δλ_{t+1})
    for (int e = lane; e < H * NRX; e += 32) {
      const int t = e / NRX, c = e % NRX;
      double acc = rx[e];
      if (c < NU) {
        for (int i = 0; i < NQ; ++i) acc = fma(-DZ(t, 2 * NQ + c, i), gv[t * NB + i], acc);
        delta[t * DL + c] = qiu[t * NU + c] * acc;
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
        yv[t * NQ + k] = acc;
      }
    }
    __syncwarp();
    for (int e = lane; e < H * NQ; e += 32) {
      const int t = e / NQ, k = e % NQ;
      double acc = 0.0;
      for (int m = 0; m <= k; ++m) acc = fma(EM(t, m, k), yv[t * NQ + m], acc);
      wv[e] = acc;
    }
    __syncwarp();
    for (int e = lane; e < H * NQ; e += 32) {
      const int t = e / NQ, k = e % NQ;
      double acc = 0.0;
      for (int m = k; m < NQ; ++m) acc = fma(EM(t, k, m), wv[t * NQ + m], acc);
      delta[t * DL + NU + k] = acc;
    }
    for (int e = lane; e < H * NQ; e += 32) delta[(e / NQ) * DL + NRX + e % NQ] = gv[(e / NQ) * NB + e % NQ];
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
  __syncwarp();
  for (int e = lane; e < (H + 2) * NQ; e += 32) cq_g[e] = cq[e];
  for (int e = lane; e < H * NU; e += 32) cu_g[e] = cu[e];
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
