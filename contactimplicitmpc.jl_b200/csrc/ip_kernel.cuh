// Batched Mehrotra predictor-corrector interior-point solve of the linearized contact subproblem.
//
// One group of G lanes (G = 4, 16 or 32; 32/G subproblems per warp) owns one subproblem: lane l
// holds row l of every vector (x_l, y1_l, y2_l, residual rows) and row l of the ny×ny Schur
// complement in registers.  Cross-row traffic is warp shuffles of width G.  Per-knot constants
// (Dims::LinStore) are shared by every subproblem on the same reference knot and are read through
// L1/L2.  Persistent grid-stride loop over the batch.
//
// Replaces, per subproblem (reference file:line):
//   rlin!                         src/controller/linearized_solver.jl:364-373     -> residual()
//   rzlin! + schur_factorize!     linearized_solver.jl:378-399, src/solver/schur.jl:80-88 -> factor()
//   linear_solve!(Δ, rz, r)       linearized_solver.jl:424-444, schur.jl:93-110   -> lu_solve() + products
//   linear_solve!(δz, rz, rθ)     linearized_solver.jl:451-479                    -> sensitivities
//   residual_/bilinear_violation  linearized_solver.jl:401-409                    -> group max-reductions
//   general_correction_term!      linearized_solver.jl:411-418
//   interior_point_solve!         RoboDojo 0.1.3 (external; iteration defined in oracle/ip.py)
// The ny×ny system is factorised by LU with partial pivoting (rows stay in their lanes; the pivot
// order is tracked) instead of the reference's modified Gram-Schmidt QR (src/solver/qr.jl:113-158):
// same solution to fp64 round-off, one third of the flops, no column dot-products across lanes.
#pragma once
#include <cfloat>
#include <cstdint>
#include <utility>

#include "dims.cuh"
#include "../../include/cimpc_b200.h"

namespace cimpc {

struct IpParams {
  int64_t n;
  const int32_t* __restrict__ knot;
  const double* __restrict__ theta;
  const double* __restrict__ q2_init;
  const double* __restrict__ alt;  // may be null
  const double* __restrict__ lin;  // H_ref × LIN_STRIDE
  int32_t h_ref;
  double* __restrict__ z_out;
  double* __restrict__ dz_out;
  uint8_t* __restrict__ status;
  int32_t* __restrict__ iters;
  cimpc_ip_opts o;
};

constexpr unsigned FULL = 0xffffffffu;

// Guaranteed compile-time unrolling (register arrays must never be indexed dynamically).
template <int B, int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, B + Is>{}), ...);
}
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (E > B) static_for_impl<B>(std::make_integer_sequence<int, E - B>{}, static_cast<F&&>(f));
}
constexpr int UNPIV = 1 << 20;

template <int G>
__device__ __forceinline__ double bc(double v, int src) {
  return __shfl_sync(FULL, v, src, G);
}
template <int G>
__device__ __forceinline__ double gmax(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o, G));
  return v;
}
template <int G>
__device__ __forceinline__ double gmin(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o, G));
  return v;
}
template <int G>
__device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
  return v;
}

// Per-lane LU state of the NY×NY Schur complement S = S0 − diag(Ry2 ŷ2 / ŷ1):
// a[j] = row `l` of the factors (multipliers below the pivot order, U on and above it).
template <int NY>
struct LU {
  double a[NY];
  int piv[NY];  // lane (row) chosen as pivot of elimination step k — identical in all lanes
  int mystep;   // elimination step at which this lane's row was the pivot row
  int mypl;     // piv[l]: lane that holds unknown number l after back-substitution
  double invp;  // 1 / (pivot element of this lane's row)
};

template <class D>
__device__ __forceinline__ void factor(LU<D::NY>& f, int l, int gshift) {
  constexpr int NY = D::NY, G = D::G;
  f.mystep = UNPIV;
  f.mypl = 0;
  f.invp = 1.0;
  static_for<0, NY>([&](auto K) {
    constexpr int k = decltype(K)::value;
    const bool unp = (f.mystep == UNPIV) && (l < NY);
    const double cand = unp ? fabs(f.a[k]) : -1.0;
    const double m = gmax<G>(cand);
    unsigned bal = __ballot_sync(FULL, cand == m);
    bal = (G == 32) ? bal : ((bal >> gshift) & ((1u << (G & 31)) - 1u));
    const int p = __ffs(bal) - 1;
    f.piv[k] = p;
    if (l == k) f.mypl = p;
    const double inv_own = 1.0 / f.a[k];
    const double pinv = bc<G>(inv_own, p);
    const bool is_p = (l == p);
    if (is_p) {
      f.mystep = k;
      f.invp = inv_own;
    }
    const bool elim = unp && !is_p;
    const double mult = f.a[k] * pinv;
    if (elim) f.a[k] = mult;
    static_for<k + 1, NY>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double pj = bc<G>(f.a[j], p);
      if (elim) f.a[j] = fma(-mult, pj, f.a[j]);
    });
  });
}

// Solve S t = w for NR right-hand sides; on entry w[r] = row l of RHS r, on exit w[r] = t_l.
template <class D, int NR>
__device__ __forceinline__ void lu_solve(const LU<D::NY>& f, double (&w)[NR]) {
  constexpr int NY = D::NY, G = D::G;
  static_for<0, NY>([&](auto K) {  // forward: apply the row eliminations to the RHS
    constexpr int k = decltype(K)::value;
    const bool upd = f.mystep > k;
    static_for<0, NR>([&](auto R) {
      constexpr int r = decltype(R)::value;
      const double wp = bc<G>(w[r], f.piv[k]);
      if (upd) w[r] = fma(-f.a[k], wp, w[r]);
    });
  });
  static_for<0, NY>([&](auto K) {  // backward: U t = c, unknown k lives in lane piv[k]
    constexpr int k = NY - 1 - decltype(K)::value;
    const bool upd = f.mystep < k;
    const bool mine = f.mystep == k;
    static_for<0, NR>([&](auto R) {
      constexpr int r = decltype(R)::value;
      const double tk = bc<G>(w[r] * f.invp, f.piv[k]);
      if (upd) w[r] = fma(-f.a[k], tk, w[r]);
      if (mine) w[r] = tk;
    });
  });
  static_for<0, NR>([&](auto R) {
    constexpr int r = decltype(R)::value;
    w[r] = bc<G>(w[r], f.mypl);  // unknown l → lane l
  });
}

// rlin!: rows of [rdyn; rrst; rbil] at (x, y1, y2) with central-path parameter kappa.
template <class D>
__device__ __forceinline__ void residual(const double* __restrict__ L, int lx, int ly, bool hx, bool hy,
                                         double cdyn, double crst, double ry2, double x, double y1,
                                         double y2, double kappa, double& rdyn, double& rrst,
                                         double& rbil) {
  constexpr int NX = D::NX, NY = D::NY, G = D::G;
  double ad = cdyn, ar = fma(ry2, y2, crst);
#pragma unroll
  for (int j = 0; j < NX; ++j) {
    const double xb = bc<G>(x, j);
    ad = fma(__ldg(L + D::O_DX + lx + j * NX), xb, ad);
    ar = fma(__ldg(L + D::O_RX + ly + j * NY), xb, ar);
  }
#pragma unroll
  for (int j = 0; j < NY; ++j) {
    const double yb = bc<G>(y1, j);
    ad = fma(__ldg(L + D::O_DY1 + lx + j * NX), yb, ad);
    ar = fma(__ldg(L + D::O_RY1 + ly + j * NY), yb, ar);
  }
  rdyn = hx ? ad : 0.0;
  rrst = hy ? ar : 0.0;
  rbil = hy ? fma(y1, y2, -kappa) : 0.0;
}

template <class D>
__device__ __forceinline__ double step_length(bool hy, double y1, double y2, double d1, double d2,
                                              double tau) {
  // largest α ≤ 1 with y − αΔ ≥ (1−τ) y on the orthant (fraction to the boundary)
  double a = 1.0;
  if (hy && d1 > 0.0) a = fmin(a, tau * y1 / d1);
  if (hy && d2 > 0.0) a = fmin(a, tau * y2 / d2);
  return gmin<D::G>(a);
}

template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS) ip_solve_kernel(const IpParams p) {
  constexpr int NX = D::NX, NY = D::NY, NTH = D::NTH, NCOL = D::NCOL, NZ = D::NZ, NC = D::NC;
  constexpr int G = D::G, PPW = 32 / G;
  constexpr int NTHR = (NTH + G - 1) / G;

  const int lane = threadIdx.x & 31;
  const int l = lane % G;
  const int gi = lane / G;
  const int gshift = gi * G;
  const bool hx = l < NX, hy = l < NY;
  const int lx = hx ? l : 0, ly = hy ? l : 0;
  const int64_t warp0 = (int64_t)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (THREADS / 32);
  const cimpc_ip_opts o = p.o;

  for (int64_t wb = warp0 * PPW; wb < p.n; wb += nwarps * PPW) {
    const int64_t prob = wb + gi;
    const bool valid = prob < p.n;
    const int64_t pi = valid ? prob : wb;  // idle groups shadow the warp's first problem
    int kn = p.knot[pi];
    kn = (kn < 0 || kn >= p.h_ref) ? 0 : kn;  // range is validated by the host entry point
    const double* __restrict__ L = p.lin + (int64_t)kn * D::LIN_STRIDE;

    // ---- prologue: θ-dependent constants  c = c0 + Rθ θ (+ alt on the impact rows) ----
    double th[NTHR];
#pragma unroll
    for (int r = 0; r < NTHR; ++r) {
      const int j = l + r * G;
      th[r] = (j < NTH) ? p.theta[pi * NTH + j] : 0.0;
    }
    double cdyn = __ldg(L + D::O_CD + lx), crst = __ldg(L + D::O_CR + ly);
#pragma unroll
    for (int j = 0; j < NTH; ++j) {
      const double tb = bc<G>(th[j / G], j % G);
      cdyn = fma(__ldg(L + D::O_RTD + lx + j * NX), tb, cdyn);
      crst = fma(__ldg(L + D::O_RTR + ly + j * NY), tb, crst);
    }
    if (p.alt != nullptr && l < NC) crst += p.alt[pi * NC + l];
    const double ry2 = hy ? __ldg(L + D::O_RY2 + ly) : 0.0;

    // ---- cold start: z = 1, z[q2] = q2_init  (z_initialize!) ----
    double x = hx ? p.q2_init[pi * NX + lx] : 0.0;
    double y1 = 1.0, y2 = 1.0;
    double rdyn, rrst, rbil;
    residual<D>(L, lx, ly, hx, hy, cdyn, crst, ry2, x, y1, y2, 0.0, rdyn, rrst, rbil);
    double r_vio = gmax<G>(fmax(fabs(rdyn), fabs(rrst)));
    double k_vio = gmax<G>(fabs(rbil));

    bool done = !valid;
    int iters = 0;
    double reg = 0.0;
    LU<NY> f;

    for (int it = 0; it < o.max_iter; ++it) {
      if (r_vio < o.r_tol && k_vio < o.kappa_tol) done = true;
      if (__all_sync(FULL, done)) break;

      const double reg_it = (k_vio < o.kappa_reg) ? k_vio * o.gamma_reg : 0.0;
      const double y1r = fmax(y1, reg_it), y2r = fmax(y2, reg_it);
      // rzlin!: S = (Ry1 − Rx Dx⁻¹ Dy1) − diag(Ry2 ŷ2 / ŷ1), then factorise
      const double dd = ry2 * y2r / y1r;
#pragma unroll
      for (int j = 0; j < NY; ++j) {
        const double s0 = hy ? __ldg(L + D::O_S0 + ly + j * NY) : 0.0;
        f.a[j] = (j == l) ? s0 - dd : s0;
      }
      factor<D>(f, l, gshift);

      // constant products with u = rdyn (shared by predictor and corrector)
      double cu = 0.0, au = 0.0;
#pragma unroll
      for (int j = 0; j < NX; ++j) {
        const double ub = bc<G>(rdyn, j);
        cu = fma(__ldg(L + D::O_CAI + ly + j * NY), ub, cu);
        au = fma(__ldg(L + D::O_AI + lx + j * NX), ub, au);
      }

      // ---- predictor (affine) direction: only Δy1, Δy2 are needed ----
      double w[1];
      w[0] = hy ? cu - (rrst - ry2 * rbil / y1r) : 0.0;
      lu_solve<D, 1>(f, w);
      const double dy1a = -w[0];
      const double dy2a = hy ? (rbil - y2r * dy1a) / y1r : 0.0;
      const double a_aff = step_length<D>(hy, y1, y2, dy1a, dy2a, 1.0);
      const double mu = gsum<G>(hy ? y1 * y2 : 0.0) / (double)NY;
      const double mu_aff =
          gsum<G>(hy ? (y1 - a_aff * dy1a) * (y2 - a_aff * dy2a) : 0.0) / (double)NY;
      double sg = fmin(fmax(mu_aff / mu, 0.0), 1.0);
      sg = sg * sg * sg;
      const double kap = fmax(sg * mu, o.kappa_tol / o.undercut);

      // ---- corrector: rbil = y1∘y2 − κ + Δy1aff∘Δy2aff ----
      const double rbc = hy ? fma(y1, y2, -kap) + dy1a * dy2a : 0.0;
      w[0] = hy ? cu - (rrst - ry2 * rbc / y1r) : 0.0;
      lu_solve<D, 1>(f, w);
      const double t = w[0];
      const double dy1 = -t;
      const double dy2 = hy ? (rbc - y2r * dy1) / y1r : 0.0;
      double dx = au;
#pragma unroll
      for (int j = 0; j < NY; ++j) dx = fma(__ldg(L + D::O_AIB + lx + j * NX), bc<G>(t, j), dx);
      if (!hx) dx = 0.0;

      const double vmax = fmax(r_vio, k_vio);
      const double tau = fmax(1.0 - o.eps_min, 1.0 - vmax * vmax);
      double alpha = step_length<D>(hy, y1, y2, dy1, dy2, tau);

      // ---- candidate + back-tracking on the violations (trial max_ls is accepted unconditionally) ----
      bool acc = done;
      double xc = x, y1c = y1, y2c = y2, rdc = rdyn, rrc = rrst, rbc2 = rbil, rvc = r_vio, kvc = k_vio;
      for (int ls = 0; ls <= o.max_ls; ++ls) {
        const double xt = x - alpha * dx, y1t = y1 - alpha * dy1, y2t = y2 - alpha * dy2;
        double rd, rr, rb;
        residual<D>(L, lx, ly, hx, hy, cdyn, crst, ry2, xt, y1t, y2t, 0.0, rd, rr, rb);
        const double rv = gmax<G>(fmax(fabs(rd), fabs(rr)));
        const double kv = gmax<G>(fabs(rb));
        if (!acc) {
          xc = xt; y1c = y1t; y2c = y2t; rdc = rd; rrc = rr; rbc2 = rb; rvc = rv; kvc = kv;
          if (rv <= r_vio || kv <= k_vio || ls == o.max_ls) acc = true;
          else alpha *= o.ls_scale;
        }
        if (__all_sync(FULL, acc)) break;
      }
      if (!done) {
        x = xc; y1 = y1c; y2 = y2c; rdyn = rdc; rrst = rrc; rbil = rbc2; r_vio = rvc; k_vio = kvc;
        reg = reg_it;
        ++iters;
      }
    }
    const bool conv = (r_vio < o.r_tol) && (k_vio < o.kappa_tol);

    // ---- outputs: z*, status, iteration count ----
    if (valid) {
      double* zo = p.z_out + prob * NZ;
      if (hx) zo[l] = x;
      if (hy) {
        zo[NX + l] = y1;
        zo[NX + NY + l] = y2;
      }
      if (l == 0) {
        p.status[prob] = conv ? 1 : 0;
        p.iters[prob] = iters;
      }
    }

    // ---- differentiate_solution!: δz = −rz⁻¹ rθ on the consumed rows / columns ----
    if (o.diff_sol) {
      const double reg_d = fmax(reg, o.kappa_tol * o.gamma_reg);
      const double y1r = fmax(y1, reg_d), y2r = fmax(y2, reg_d);
      const double dd = ry2 * y2r / y1r;
#pragma unroll
      for (int j = 0; j < NY; ++j) {
        const double s0 = hy ? __ldg(L + D::O_S0 + ly + j * NY) : 0.0;
        f.a[j] = (j == l) ? s0 - dd : s0;
      }
      factor<D>(f, l, gshift);
      constexpr int CH = (NCOL % 6 == 0) ? 6 : ((NCOL % 5 == 0) ? 5 : 1);
      double* dzo = p.dz_out + prob * (int64_t)(D::ND * NCOL);
      for (int c0 = 0; c0 < NCOL; c0 += CH) {
        double w[CH];
#pragma unroll
        for (int r = 0; r < CH; ++r) w[r] = hy ? __ldg(L + D::O_W + ly + (c0 + r) * NY) : 0.0;
        lu_solve<D, CH>(f, w);  // w = S⁻¹ (CAi Rθdyn − Rθrst)[:, c] = −δy1
        double dxo[CH];
#pragma unroll
        for (int r = 0; r < CH; ++r) dxo[r] = __ldg(L + D::O_AR + lx + (c0 + r) * NX);
#pragma unroll
        for (int j = 0; j < NY; ++j) {
          const double m = __ldg(L + D::O_AIB + lx + j * NX);
#pragma unroll
          for (int r = 0; r < CH; ++r) dxo[r] = fma(m, bc<G>(w[r], j), dxo[r]);
        }
        if (valid) {
#pragma unroll
          for (int r = 0; r < CH; ++r) {
            if (hx) dzo[(c0 + r) * D::ND + l] = -dxo[r];
            if (D::NYD > 0 && l < D::NYD) dzo[(c0 + r) * D::ND + NX + l] = w[r];
          }
        }
      }
    }
  }
}

}  // namespace cimpc
