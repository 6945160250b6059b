// Batched interior-point solve of the linearized contact subproblem, kernel v3: TWO ROWS PER LANE.
// MEASURED AND REJECTED (profiles/r02_ip_two_rows_per_lane.md): 0.80-0.90x the throughput of ip_kernel.cuh's v2 on every
// BASELINE config.  Compiled only with -DCIMPC_WITH_IP_V3 so that the measurement stays reproducible
// (scripts/gpu_ip_ab.py, scripts/gpu_ip_ab_ncu.sh); the shipped library does not contain it.
//
// Same algorithm and the same LinStore / scratch layout as v2, per-row arithmetic written operation by operation like
// v2's (same pivots, same summation trees; results agree to round-off amplified by the conditioning of S - identical
// status, iteration counts equal wherever the line search is off).  What changes is the mapping: v2 gives one subproblem
// G = pow2ceil(ny) lanes, one row per lane; here a subproblem owns GL = G/2 lanes and lane l holds rows l and l + GL:
//   * a warp holds 64/G subproblems of the same knot and their lanes read the SAME constant words, so a warp-wide load
//     of the per-knot constants serves twice as many subproblems (LDS.128 instructions per subproblem -32 %);
//   * per-lane control (loop counters, addresses, pivot search, predicates) is shared by two rows, and every dependent
//     FMA chain has an independent sibling (ILP 2 at half the lanes).
// What the measurement showed: a broadcast load costs one shared-memory wavefront per DISTINCT address, i.e. per
// subproblem, whatever the number of lanes or rows it feeds - the pivot-row / iterate broadcasts, the larger half of the
// shared traffic, do not shrink (wavefronts per subproblem -12 % only); warp instructions per subproblem fall by 8 %, not
// 25 % (DFMA, selects and register moves are per row); and the 168 registers of two rows per lane leave 12 instead of 16
// resident warps per SM on a kernel whose stalls are dependency latencies (`wait`, `short_scoreboard`), with a code
// footprint 1.5x larger (`no_instruction` 0.37 -> 0.95 stalls per issue).
// The psi lanes' passenger rows (ip_kernel.cuh, load_schur) are rows NR..NY-1 of slot 1, still free riders of the
// warp-wide update instructions.
//
// Reference functions replaced: the same list as ip_kernel.cuh (rlin!, rzlin!, schur_factorize!, linear_solve!,
// residual_/bilinear_violation, general_correction_term!: src/controller/linearized_solver.jl:364-479,
// src/solver/schur.jl:80-110; interior_point_solve!: RoboDojo 0.1.3, iteration defined in oracle/ip.py).
#pragma once
#include "ip_kernel.cuh"

namespace cimpc {

template <class D>
struct GroupScratch2 : GroupScratch<D> {
  using Base = GroupScratch<D>;
  static constexpr int O_RB = Base::DOUBLES;  // per-row exchange: w_l (load_schur) / rhs_l (schur_solve) of the reduced rows
  static constexpr int DOUBLES = round_up(O_RB + D::NRP, 2);
};

template <class D>
struct Ctx2 {  // per-lane registers: rows l (slot 0) and l + GL (slot 1) of one subproblem
  double x[2], y1[2], y2[2];
  double cdyn[2], crst[2], ry2[2];
  double rdyn[2], rrst[2], rbil[2];
  double bv[2];
  int cid[2], b0[2];
  double wl[2], a1[2], a2[2], a3[2];
  bool dpiv[2];
  bool any_b;
  double M[2][D::NRP];
  double msc[2];
  int mystep[2];
};

#define CIMPC_ROWS(r, row)                            \
  static_for<0, 2>([&](auto R_) {                     \
    constexpr int r = decltype(R_)::value;            \
    const int row = l + r * GL;                       \
    (void)row;
#define CIMPC_ROWS_END \
  });

template <int GL>
__device__ __forceinline__ double gmax2(double a, double b) {
  double v = fmax(a, b);
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o, GL));
  return v;
}
template <int GL>
__device__ __forceinline__ double gmin2(double a, double b) {
  double v = fmin(a, b);
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o, GL));
  return v;
}
// a + b is the first level (offset G/2) of v2's width-G butterfly: same summation tree, same bits
template <int GL>
__device__ __forceinline__ double gsum2(double a, double b) {
  double v = a + b;
#pragma unroll
  for (int o = GL / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, GL);
  return v;
}

template <int GL>
__device__ __forceinline__ unsigned group_bits(unsigned bal, int gshift) {
  return (bal >> gshift) & ((1u << GL) - 1u);
}

// rlin! for both rows of the lane.
template <class D>
__device__ __forceinline__ void residual2(const double* __restrict__ Ls, double* __restrict__ sc, int l,
                                          const Ctx2<D>& c, const double (&x)[2], const double (&y1)[2],
                                          const double (&y2)[2], double kappa, double (&rdyn)[2], double (&rrst)[2],
                                          double (&rbil)[2]) {
  constexpr int NX = D::NX, NY = D::NY, G = D::G, GL = G / 2;
  using S = GroupScratch2<D>;
  CIMPC_ROWS(r, row)
    if (row < NX) sc[S::O_XY + row] = x[r];
    if (row < NY) sc[S::O_XY + NX + row] = y1[r];
  CIMPC_ROWS_END
  __syncwarp();
  double ad[2], ar[2], ad2[2] = {0.0, 0.0}, ar2[2] = {0.0, 0.0};
  CIMPC_ROWS(r, row)
    ad[r] = c.cdyn[r];
    ar[r] = fma(c.ry2[r], y2[r], c.crst[r]);
  CIMPC_ROWS_END
  const double* R = Ls + D::O_RES + 2 * l;
  constexpr int NJ = NX + NY;
#pragma unroll 2
  for (int j = 0; j + 1 < NJ; j += 2) {
    const double2 v = lds2(sc + S::O_XY + j);
    CIMPC_ROWS(r, row)
      const double2 c0 = lds2(R + 2 * r * GL + (j)*G * 2), c1 = lds2(R + 2 * r * GL + (j + 1) * G * 2);
      ad[r] = fma(c0.x, v.x, ad[r]);
      ar[r] = fma(c0.y, v.x, ar[r]);
      ad2[r] = fma(c1.x, v.y, ad2[r]);
      ar2[r] = fma(c1.y, v.y, ar2[r]);
    CIMPC_ROWS_END
  }
  if constexpr (NJ & 1) {
    const double v = sc[S::O_XY + NJ - 1];
    CIMPC_ROWS(r, row)
      const double2 c0 = lds2(R + 2 * r * GL + (NJ - 1) * G * 2);
      ad[r] = fma(c0.x, v, ad[r]);
      ar[r] = fma(c0.y, v, ar[r]);
    CIMPC_ROWS_END
  }
  __syncwarp();
  CIMPC_ROWS(r, row)
    const double sd = ad[r] + ad2[r], sr = ar[r] + ar2[r];
    rdyn[r] = row < NX ? sd : 0.0;
    rrst[r] = row < NY ? sr : 0.0;
    rbil[r] = row < NY ? fma(y1[r], y2[r], -kappa) : 0.0;
  CIMPC_ROWS_END
}

// Gauss-Jordan inversion of the reduced block, two rows per lane (see invert() in ip_kernel.cuh for the scheme:
// deferred pivot scaling, reciprocal formed during the pivot search, pivot row double-buffered by step parity).
// The pivot of a step is the unpivoted row with the largest leading 32 bits of |a|, lowest ROW index on ties —
// the choice v2 makes with one row per lane.
template <class D, bool RECORD_PL>
__device__ __forceinline__ void invert2(Ctx2<D>& c, double* __restrict__ sc, int l, int gi, int gshift) {
  constexpr int NY = D::NRP, G = D::G, GL = G / 2, NGRP = 32 / GL;
  using S = GroupScratch2<D>;
  c.mystep[0] = c.mystep[1] = UNPIV;
  c.msc[0] = c.msc[1] = 0.0;
#pragma unroll 1
  for (int kb = 0; kb < NY; kb += 4) {
    static_for<0, 4>([&](auto U) {
      constexpr int u = decltype(U)::value;
      double* const prow = sc + S::O_PROW + (u & 1) * NY;
      double myinv[2];
      unsigned cand[2];
      CIMPC_ROWS(r, row)
        const bool unp = (c.mystep[r] == UNPIV) && row < NY;
        myinv[r] = __drcp_rn(c.M[r][u]);
        cand[r] = unp ? ((unsigned)__double2hiint(c.M[r][u]) & 0x7fffffffu) + 1u : 0u;
      CIMPC_ROWS_END
      const unsigned mc = max(cand[0], cand[1]);
      unsigned m = 0u;
      static_for<0, NGRP>([&](auto Gi) {
        constexpr int g = decltype(Gi)::value;
        const unsigned mg = __reduce_max_sync(FULL, gi == g ? mc : 0u);
        if (gi == g) m = mg;
      });
      const unsigned bal0 = group_bits<GL>(__ballot_sync(FULL, cand[0] == m), gshift);
      const unsigned bal1 = group_bits<GL>(__ballot_sync(FULL, cand[1] == m), gshift);
      const int prow_idx = bal0 ? __ffs(bal0) - 1 : __ffs(bal1) - 1 + GL;
      bool is_p[2];
      CIMPC_ROWS(r, row)
        is_p[r] = (row == prow_idx);
        if (is_p[r]) {  // publish the pivot row
          c.mystep[r] = kb + u;
          if (RECORD_PL) reinterpret_cast<int*>(sc + S::O_PL)[kb + u] = row;
          static_for<0, NY / 2>([&](auto J) {
            constexpr int j = decltype(J)::value;
            *reinterpret_cast<double2*>(prow + 2 * j) =
                make_double2(2 * j == u ? myinv[r] : c.M[r][2 * j], 2 * j + 1 == u ? myinv[r] : c.M[r][2 * j + 1]);
          });
          c.msc[r] = myinv[r];
        }
      CIMPC_ROWS_END
      __syncwarp();
      const double pinv = prow[u];
      double g[2];
      CIMPC_ROWS(r, row)
        g[r] = is_p[r] ? 0.0 : -c.M[r][u] * pinv;
      CIMPC_ROWS_END
      static_for<0, NY / 2>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const double2 pr = lds2(prow + 2 * j);
        CIMPC_ROWS(r, row)
          if constexpr (2 * j != u) c.M[r][2 * j] = fma(g[r], pr.x, c.M[r][2 * j]);
          if constexpr (2 * j + 1 != u) c.M[r][2 * j + 1] = fma(g[r], pr.y, c.M[r][2 * j + 1]);
        CIMPC_ROWS_END
      });
      CIMPC_ROWS(r, row)
        c.M[r][u] = is_p[r] ? 1.0 : g[r];
      CIMPC_ROWS_END
    });
    // rotate both rows left by four column slots
    CIMPC_ROWS(r, row)
      const double t0 = c.M[r][0], t1 = c.M[r][1], t2 = c.M[r][2], t3 = c.M[r][3];
      static_for<0, NY - 4>([&](auto J) {
        constexpr int j = decltype(J)::value;
        c.M[r][j] = c.M[r][j + 4];
      });
      c.M[r][NY - 4] = t0;
      c.M[r][NY - 3] = t1;
      c.M[r][NY - 2] = t2;
      c.M[r][NY - 1] = t3;
    CIMPC_ROWS_END
  }
}

// t = S_red⁻¹ w for both rows of the lane; raw = un-scaled dot products (passenger rows: −A[b0,:] t_r).
template <class D>
__device__ __forceinline__ void apply_inverse2(const Ctx2<D>& c, double* __restrict__ sc, int l, const double (&w)[2],
                                               double (&t)[2], double (&raw)[2]) {
  constexpr int NY = D::NRP, GL = D::G / 2;
  using S = GroupScratch2<D>;
  CIMPC_ROWS(r, row)
    if (row < NY) sc[S::O_WV + c.mystep[r]] = w[r];
  CIMPC_ROWS_END
  __syncwarp();
  double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0}, a2[2] = {0.0, 0.0}, a3[2] = {0.0, 0.0};
  static_for<0, NY / 4>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const double2 v = lds2(sc + S::O_WV + 4 * j), ww = lds2(sc + S::O_WV + 4 * j + 2);
    CIMPC_ROWS(r, row)
      a0[r] = fma(c.M[r][4 * j], v.x, a0[r]);
      a1[r] = fma(c.M[r][4 * j + 1], v.y, a1[r]);
      a2[r] = fma(c.M[r][4 * j + 2], ww.x, a2[r]);
      a3[r] = fma(c.M[r][4 * j + 3], ww.y, a3[r]);
    CIMPC_ROWS_END
  });
  CIMPC_ROWS(r, row)
    raw[r] = (a0[r] + a1[r]) + (a2[r] + a3[r]);
    const double tt = raw[r] * c.msc[r];
    if (row < NY) sc[S::O_TV + c.mystep[r]] = tt;
  CIMPC_ROWS_END
  __syncwarp();
  CIMPC_ROWS(r, row)
    t[r] = row < NY ? sc[S::O_TV + row] : 0.0;
  CIMPC_ROWS_END
}

template <class D>
__device__ __forceinline__ double step_length2(const double (&y1)[2], const double (&y2)[2], const double (&d1)[2],
                                               const double (&d2)[2], double tau, int l) {
  constexpr int GL = D::G / 2;
  double a[2] = {1.0, 1.0};
  CIMPC_ROWS(r, row)
    const bool hy = row < D::NY;
    if (hy && d1[r] > 0.0) a[r] = fmin(a[r], tau * y1[r] / d1[r]);
    if (hy && d2[r] > 0.0) a[r] = fmin(a[r], tau * y2[r] / d2[r]);
  CIMPC_ROWS_END
  return gmin2<GL>(a[0], a[1]);
}

// rzlin!: rows l and l + GL of S_red (or of S_redᵀ); see load_schur() in ip_kernel.cuh for the row combinations.
template <class D, bool TRANSPOSED>
__device__ __forceinline__ void load_schur2(Ctx2<D>& c, const double* __restrict__ Ls, double* __restrict__ sc, int l,
                                            int gshift, double reg) {
  constexpr int NY = D::NY, NR = D::NR, NRP = D::NRP, G = D::G, GL = G / 2;
  static_assert(NRP == NR, "two-rows-per-lane kernel: reduced block without identity padding");
  using S = GroupScratch2<D>;
  CIMPC_ROWS(r, row)
    const bool hy = row < NY, hr = row < NR, hp = hy && row >= NR;
    const double y1r = fmax(c.y1[r], reg), y2r = fmax(c.y2[r], reg);
    c.wl[r] = hy ? c.ry2[r] * y2r / y1r : 1.0;
    if (hp) *reinterpret_cast<double2*>(sc + S::O_V + 2 * (row - NR)) = make_double2(1.0 / c.wl[r], c.wl[r]);
    if (hr) sc[S::O_RB + row] = c.wl[r];
  CIMPC_ROWS_END
  __syncwarp();
  double wb0[2];
  CIMPC_ROWS(r, row)
    const bool hy = row < NY, hr = row < NR, hp = hy && row >= NR;
    c.a1[r] = 1.0; c.a2[r] = 0.0; c.a3[r] = 0.0;
    c.dpiv[r] = true;
    if (hr && c.cid[r] >= 0) {
      const double2 vw = lds2(sc + S::O_V + 2 * c.cid[r]);  // {1/w_ψ, w_ψ}
      c.dpiv[r] = vw.y * fabs(Ls[D::O_IBV0 + row]) >= PSI_PIVOT_RATIO;
      if (c.dpiv[r]) {
        c.a2[r] = vw.x;
      } else if (row == c.b0[r]) {
        c.a2[r] = 1.0 / c.bv[r];
        c.a1[r] = vw.y * c.a2[r];
      } else {
        c.a3[r] = -Ls[D::O_RHO + row];
      }
    }
    if (hp) {
      c.dpiv[r] = c.wl[r] * fabs(Ls[D::O_IBV0 + row]) >= PSI_PIVOT_RATIO;
      c.a2[r] = 1.0 / c.wl[r];
    }
    // w of the first row of this row's contact (own value on rows without one)
    wb0[r] = c.b0[r] >= 0 ? sc[S::O_RB + c.b0[r]] : c.wl[r];
  CIMPC_ROWS_END
  c.any_b = group_bits<GL>(__ballot_sync(FULL, !c.dpiv[0] || !c.dpiv[1]), gshift) != 0u;
  if constexpr (!TRANSPOSED) {
    CIMPC_ROWS(r, row)
      const bool hy = row < NY, hr = row < NR, hp = hy && row >= NR;
      const bool use_ab = (hr && c.a3[r] != 0.0) || hp;  // rows l ≠ b0 of a B-pivot contact; passengers
      const double a1 = hr ? c.a1[r] : 0.0;
      const double ax = hr ? (c.a3[r] != 0.0 ? c.a3[r] : c.a2[r]) : (hp ? 1.0 : 0.0);
      const double* A0 = Ls + D::O_S0 + row;
      const double* AX = Ls + (use_ab ? D::O_AB : D::O_BC) + row;
      const int jb = use_ab ? c.b0[r] : -1;
      const double dl = hr ? a1 * c.wl[r] : 0.0, db = ax * wb0[r];
      static_for<0, NRP>([&](auto J) {
        constexpr int j = decltype(J)::value;
        double m = fma(a1, A0[j * G], ax * AX[j * G]);
        if (j == row) m -= dl;
        if (j == jb) m -= db;
        c.M[r][j] = m;
      });
    CIMPC_ROWS_END
  } else {
    CIMPC_ROWS(r, row)
      const bool hr = row < NR;
      if (hr) {
        *reinterpret_cast<double2*>(sc + S::O_VR + 4 * row) = make_double2(c.a1[r], c.a2[r]);
        sc[S::O_VR + 4 * row + 2] = c.a3[r];
      }
    CIMPC_ROWS_END
    __syncwarp();
    bool is_b0_bpiv[2];
    CIMPC_ROWS(r, row)
      is_b0_bpiv[r] = row < NR && c.cid[r] >= 0 && !c.dpiv[r] && row == c.b0[r];
    CIMPC_ROWS_END
    static_for<0, NRP>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 a12 = lds2(sc + S::O_VR + 4 * j);
      const double a3 = sc[S::O_VR + 4 * j + 2];
      const int cidj = (int)Ls[D::O_CIDU + j];
      const double rhoj = Ls[D::O_RHOU + j];
      CIMPC_ROWS(r, row)
        const bool hr = row < NR;
        double m = a12.x * Ls[D::O_S0T + row + j * G];
        m = fma(a12.y, Ls[D::O_BCT + row + j * G], m);
        m = fma(a3, Ls[D::O_ABT + row + j * G], m);
        if (j == row && hr) m = fma(-c.a1[r], c.wl[r], m);
        if (is_b0_bpiv[r] && j != row && cidj == c.cid[r]) m = fma(rhoj, c.wl[r], m);
        c.M[r][j] = hr ? m : 0.0;
      CIMPC_ROWS_END
    });
  }
}

// t = S⁻¹ rhs for the full ny-vector through the reduced inverse (schur_solve() in ip_kernel.cuh), two rows per lane.
template <class D>
__device__ __forceinline__ void schur_solve2(const Ctx2<D>& c, const double* __restrict__ Ls, double* __restrict__ sc,
                                             int l, const double (&rhs)[2], double (&t)[2]) {
  constexpr int NY = D::NY, NR = D::NR, G = D::G, GL = G / 2;
  using S = GroupScratch2<D>;
  CIMPC_ROWS(r, row)
    const bool hy = row < NY, hr = row < NR, hp = hy && row >= NR;
    if (hp) sc[S::O_G + row - NR] = rhs[r];
    if (hr) sc[S::O_RB + row] = rhs[r];
  CIMPC_ROWS_END
  __syncwarp();
  double rb0[2], rr[2];
  CIMPC_ROWS(r, row)
    const bool hr = row < NR;
    rb0[r] = c.b0[r] >= 0 ? sc[S::O_RB + c.b0[r]] : rhs[r];  // rhs of the contact's first row
    rr[r] = 0.0;
    if (hr) {
      rr[r] = c.a1[r] * rhs[r];
      if (c.cid[r] >= 0)
        rr[r] = (c.a3[r] != 0.0) ? fma(c.a3[r], rb0[r], rr[r]) : fma(c.a2[r] * c.bv[r], sc[S::O_G + c.cid[r]], rr[r]);
    }
  CIMPC_ROWS_END
  double tr[2], raw[2];
  apply_inverse2<D>(c, sc, l, rr, tr, raw);
  CIMPC_ROWS(r, row)
    const bool hy = row < NY, hr = row < NR, hp = hy && row >= NR;
    if constexpr ((r + 1) * GL > NR) {  // this slot holds ψ rows
      const double tpsi_b = (rb0[r] + raw[r]) * Ls[D::O_IBV0 + row];  // raw = −A[b0,:] t_r (passenger rows)
      double a = -rhs[r];
#pragma unroll
      for (int k = 0; k < D::NFR; ++k) {
        const double cv = Ls[D::O_CRW + (2 * k) * G + row];
        const int cj = (int)Ls[D::O_CRW + (2 * k + 1) * G + row];
        a = fma(cv, sc[S::O_TV + cj], a);
      }
      const double tpsi = c.dpiv[r] ? c.a2[r] * a : tpsi_b;
      t[r] = hr ? tr[r] : (hp ? tpsi : 0.0);
    } else {
      t[r] = hr ? tr[r] : 0.0;
    }
  CIMPC_ROWS_END
}

template <class D>
__device__ __forceinline__ void fold_row_combination2(const double* __restrict__ Ls, const double* __restrict__ sc,
                                                      double* __restrict__ rowp) {
  using S = GroupScratch2<D>;
#pragma unroll 1
  for (int k = 0; k < D::NRP; ++k) {
    const double a1 = sc[S::O_VR + 4 * k];
    if (a1 != 1.0) rowp[k] *= a1;
  }
#pragma unroll 1
  for (int k = 0; k < D::NRP; ++k) {
    const double a3 = sc[S::O_VR + 4 * k + 2];
    if (a3 != 0.0) {
      const int b0 = (int)Ls[D::O_B0U + k];
      rowp[b0] = fma(a3, rowp[k], rowp[b0]);
    }
  }
}

// differentiate_solution! restricted to the consumed rows / columns (sensitivities() in ip_kernel.cuh).
template <class D>
__device__ __forceinline__ void sensitivities2(Ctx2<D>& c, const double* __restrict__ Ls, double* __restrict__ sc,
                                               int l, int gi, int gshift, double reg, double* __restrict__ dzo,
                                               bool valid) {
  constexpr int NX = D::NX, NY = D::NRP, G = D::G, GL = G / 2, NCOL = D::NCOL, ND = D::ND, NYD = D::NYD;
  static_assert(NYD == 0 || NYD == D::NR, "force rows = reduced rows");
  using S = GroupScratch2<D>;
  load_schur2<D, true>(c, Ls, sc, l, gshift, reg);
  invert2<D, (NYD > 0)>(c, sc, l, gi, gshift);  // c.M[r][j] = S_red⁻¹[q_j][k],  k = c.mystep[r]
  CIMPC_ROWS(r, row)
    if (row < NY) {
#pragma unroll 1
      for (int i = 0; i < NX; ++i) sc[S::O_AIBP + i * NY + c.mystep[r]] = Ls[D::O_AIBR + i * G + row];
    }
  CIMPC_ROWS_END
  __syncwarp();
#pragma unroll 1
  for (int i = 0; i < NX; ++i) {
    double a0[2] = {0.0, 0.0}, a1[2] = {0.0, 0.0};
    static_for<0, NY / 2>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 v = lds2(sc + S::O_AIBP + i * NY + 2 * j);
      CIMPC_ROWS(r, row)
        a0[r] = fma(c.M[r][2 * j], v.x, a0[r]);
        a1[r] = fma(c.M[r][2 * j + 1], v.y, a1[r]);
      CIMPC_ROWS_END
    });
    CIMPC_ROWS(r, row)
      if (row < NY) sc[S::O_P1 + i * S::LDP + c.mystep[r]] = (a0[r] + a1[r]) * c.msc[r];
    CIMPC_ROWS_END
  }
  __syncwarp();
  double Ar[2][NYD > 0 ? NY : 1];
  if constexpr (NYD > 0) {
    const int* pl = reinterpret_cast<const int*>(sc + S::O_PL);
    CIMPC_ROWS(r, row)
      if (row < NY) {
        static_for<0, NY>([&](auto J) {
          constexpr int j = decltype(J)::value;
          sc[S::O_AIBP + pl[j] * S::LDP + c.mystep[r]] = c.M[r][j] * c.msc[r];
        });
      }
    CIMPC_ROWS_END
    __syncwarp();
    CIMPC_ROWS(r, row)
      if (c.any_b && row < NYD) fold_row_combination2<D>(Ls, sc, sc + S::O_AIBP + row * S::LDP);
      static_for<0, NY>([&](auto J) {
        constexpr int j = decltype(J)::value;
        Ar[r][j] = (row < NYD) ? sc[S::O_AIBP + row * S::LDP + j] : 0.0;
      });
    CIMPC_ROWS_END
  }
  CIMPC_ROWS(r, row)
    if (c.any_b && row < NX) fold_row_combination2<D>(Ls, sc, sc + S::O_P1 + row * S::LDP);
    static_for<0, NY>([&](auto J) {
      constexpr int j = decltype(J)::value;
      c.M[r][j] = row < NX ? sc[S::O_P1 + row * S::LDP + j] : 0.0;
    });
  CIMPC_ROWS_END
  __syncwarp();
#pragma unroll 1
  for (int col = 0; col < NCOL; ++col) {
    const double* Wc = Ls + D::O_W + col * NY;
    double a0[2], a1[2] = {0.0, 0.0}, b0[2] = {0.0, 0.0}, b1[2] = {0.0, 0.0};
    CIMPC_ROWS(r, row)
      a0[r] = Ls[D::O_AR + col * G + row];
    CIMPC_ROWS_END
    static_for<0, NY / 2>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 w = lds2(Wc + 2 * j);
      CIMPC_ROWS(r, row)
        a0[r] = fma(c.M[r][2 * j], w.x, a0[r]);
        a1[r] = fma(c.M[r][2 * j + 1], w.y, a1[r]);
        if constexpr (NYD > 0) {
          b0[r] = fma(Ar[r][2 * j], w.x, b0[r]);
          b1[r] = fma(Ar[r][2 * j + 1], w.y, b1[r]);
        }
      CIMPC_ROWS_END
    });
    if (valid) {
      CIMPC_ROWS(r, row)
        if (row < NX) dzo[col * ND + row] = -(a0[r] + a1[r]);
        if constexpr (NYD > 0) {
          if (row < NYD) dzo[col * ND + NX + row] = b0[r] + b1[r];
        }
      CIMPC_ROWS_END
    }
  }
}

template <class D, int THREADS>
struct KernelSmem2 {
  static constexpr int GL = D::G / 2;
  static constexpr int GROUPS = THREADS / GL;
  static constexpr int GS = round_up(GroupScratch2<D>::DOUBLES, 2);
  static constexpr size_t BYTES = (size_t)(D::SMEM_DOUBLES + GROUPS * GS) * 8 + 64;
};

template <class D, int THREADS, int MIN_CTAS>
__global__ void __launch_bounds__(THREADS, MIN_CTAS) ip_solve2_kernel(const IpParams p) {
  constexpr int NX = D::NX, NY = D::NY, NTH = D::NTH, NZ = D::NZ, NC = D::NC, NR = D::NR;
  constexpr int G = D::G, GL = G / 2, PPW = 32 / GL;
  static_assert(NC <= GL, "impact rows live in slot 0");
  using S = GroupScratch2<D>;
  using KS = KernelSmem2<D, THREADS>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Ls = reinterpret_cast<double*>(smem_raw);
  double* scratch = Ls + D::SMEM_DOUBLES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + KS::GROUPS * KS::GS);
  int* ctl = reinterpret_cast<int*>(bar + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  const int l = lane % GL, gi = lane / GL, gshift = gi * GL;
  double* sc = scratch + (size_t)(tid / GL) * KS::GS;
  const cimpc_ip_opts o = p.o;

  // slot → subproblem (32-bit arithmetic on purpose, see ip_kernel.cuh)
  const int32_t* act = p.act;
  int n_act = p.n_act;
  int64_t n_slots = p.n;
  if (p.par != nullptr) {
    const int cur = *p.par;
    act = p.act2 + (size_t)cur * p.R;
    n_act = p.cnt2[cur];
    n_slots = (int64_t)p.stages * n_act;
  }
  const int nroll = p.R;
  auto sub = [=](int64_t slot) -> int64_t {
    if (act == nullptr) return slot;
    const unsigned s = (unsigned)slot, t = s / (unsigned)n_act;
    return (int64_t)t * nroll + act[s - t * (unsigned)n_act];
  };
  const int64_t c0 = n_slots * (int64_t)blockIdx.x / gridDim.x, c1 = n_slots * (int64_t)(blockIdx.x + 1) / gridDim.x;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  int staged = -1;
  uint32_t phase = 0;

  for (int64_t cur = c0; cur < c1;) {
    if (tid == 0) {
      ctl[3] = (int)(c1 - c0);
      ctl[1] = (int)(c1 - c0);
      ctl[0] = 0;
    }
    __syncthreads();
    for (int64_t i = cur + tid; i < c1; i += THREADS)
      if (p.knot[sub(i)] >= 0) {
        atomicMin(&ctl[3], (int)(i - c0));
        break;
      }
    __syncthreads();
    const int64_t first = c0 + ctl[3];
    if (first >= c1) break;
    const int kraw = p.knot[sub(first)];
    const bool bad_knot = kraw >= p.h_ref;
    const int kn = bad_knot ? 0 : kraw;
    for (int64_t i = first + 1 + tid; i < c1; i += THREADS) {
      const int ki = p.knot[sub(i)];
      if (ki >= 0 && ki != kraw) {
        atomicMin(&ctl[1], (int)(i - c0));
        break;
      }
    }
    __syncthreads();
    const int64_t seg_end = c0 + ctl[1];
    cur = first;
    if (bad_knot) {
      for (int64_t i = first + tid; i < seg_end; i += THREADS) {
        const int64_t pr = sub(i);
        if (p.knot[pr] >= 0) {
          p.status[pr] = 0;
          p.iters[pr] = 0;
        }
      }
      __syncthreads();
      cur = seg_end;
      continue;
    }
    if (kn != staged) {
      if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(D::SMEM_DOUBLES * 8));
        bulk_g2s(Ls, p.lin + (int64_t)kn * D::LIN_STRIDE, (uint32_t)(D::SMEM_DOUBLES * 8), bar);
      }
      while (!mbar_try_wait(bar, phase)) {
      }
      phase ^= 1u;
      staged = kn;
    }
    const double* __restrict__ Lg = p.lin + (int64_t)kn * D::LIN_STRIDE;

    for (;;) {
      int base = 0;
      if (lane == 0) base = atomicAdd(&ctl[0], PPW);
      base = __shfl_sync(FULL, base, 0);
      const int64_t wb = cur + base;
      if (wb >= seg_end) break;
      const int64_t slot = wb + gi;
      const int64_t prob = slot < seg_end ? sub(slot) : 0;
      const bool valid = slot < seg_end && p.knot[prob] >= 0;
      if (!__any_sync(FULL, valid)) continue;
      const int64_t pi = valid ? prob : sub(first);

      Ctx2<D> c;
#pragma unroll 1
      for (int j = l; j < NTH; j += GL) sc[S::O_XY + j] = p.theta[pi * NTH + j];
      __syncwarp();
      {
        double cd[2], cr[2];
        CIMPC_ROWS(r, row)
          const double2 c00 = __ldg(reinterpret_cast<const double2*>(Lg + D::O_C0) + row);
          cd[r] = c00.x;
          cr[r] = c00.y;
        CIMPC_ROWS_END
        const double2* R = reinterpret_cast<const double2*>(Lg + D::O_RTH) + l;
#pragma unroll 2
        for (int j = 0; j < NTH; ++j) {
          const double t = sc[S::O_XY + j];
          CIMPC_ROWS(r, row)
            const double2 rv = __ldg(R + r * GL + j * G);
            cd[r] = fma(rv.x, t, cd[r]);
            cr[r] = fma(rv.y, t, cr[r]);
          CIMPC_ROWS_END
        }
        if (p.alt != nullptr && l < NC) {
          const int64_t arow = p.alt_by_rollout ? (int64_t)((unsigned)pi % (unsigned)nroll) : pi;
          cr[0] += p.alt[arow * NC + l];
        }
        CIMPC_ROWS(r, row)
          c.cdyn[r] = cd[r];
          c.crst[r] = cr[r];
        CIMPC_ROWS_END
      }
      __syncwarp();
      CIMPC_ROWS(r, row)
        c.ry2[r] = Ls[D::O_RY2 + row];
        c.bv[r] = Ls[D::O_BV + row];
        c.cid[r] = (int)Ls[D::O_CID + row];
        c.b0[r] = (row < NR) ? (int)Ls[D::O_B0 + row] : (int)Ls[D::O_PB0 + row];
        c.x[r] = row < NX ? p.q2_init[pi * NX + row] : 0.0;
        c.y1[r] = 1.0;
        c.y2[r] = 1.0;
      CIMPC_ROWS_END
      residual2<D>(Ls, sc, l, c, c.x, c.y1, c.y2, 0.0, c.rdyn, c.rrst, c.rbil);
      double r_vio = gmax2<GL>(fmax(fabs(c.rdyn[0]), fabs(c.rrst[0])), fmax(fabs(c.rdyn[1]), fabs(c.rrst[1])));
      double k_vio = gmax2<GL>(fabs(c.rbil[0]), fabs(c.rbil[1]));

      bool done = !valid;
      int iters = 0;
      double reg = 0.0;

#pragma unroll 1
      for (int it = 0; it < o.max_iter; ++it) {
        if (r_vio < o.r_tol && k_vio < o.kappa_tol) done = true;
        if (!(r_vio == r_vio) || !(k_vio == k_vio)) done = true;
        if (__all_sync(FULL, done)) break;

        const double reg_it = (k_vio < o.kappa_reg) ? k_vio * o.gamma_reg : 0.0;
        load_schur2<D, false>(c, Ls, sc, l, gshift, reg_it);
        invert2<D, false>(c, sc, l, gi, gshift);

        // cu = CAi rdyn, au = Ai rdyn
        double cu[2] = {0.0, 0.0}, au[2] = {0.0, 0.0};
        CIMPC_ROWS(r, row)
          if (row < NX) sc[S::O_XY + row] = c.rdyn[r];
        CIMPC_ROWS_END
        __syncwarp();
        {
          const double* CA = Ls + D::O_CA2 + 2 * l;
#pragma unroll 1
          for (int j = 0; j < NX; ++j) {
            const double ub = sc[S::O_XY + j];
            CIMPC_ROWS(r, row)
              const double2 k2 = lds2(CA + 2 * r * GL + j * G * 2);
              cu[r] = fma(k2.x, ub, cu[r]);
              au[r] = fma(k2.y, ub, au[r]);
            CIMPC_ROWS_END
          }
        }
        __syncwarp();

        // ---- predictor ----
        double y1r[2], y2r[2], iy1[2], rhs[2], t[2], dy1a[2], dy2a[2];
        CIMPC_ROWS(r, row)
          const bool hy = row < NY;
          y1r[r] = fmax(c.y1[r], reg_it);
          y2r[r] = fmax(c.y2[r], reg_it);
          iy1[r] = __drcp_rn(y1r[r]);
          rhs[r] = hy ? cu[r] - (c.rrst[r] - c.ry2[r] * c.rbil[r] * iy1[r]) : 0.0;
        CIMPC_ROWS_END
        schur_solve2<D>(c, Ls, sc, l, rhs, t);
        double pm[2], pa[2];
        CIMPC_ROWS(r, row)
          const bool hy = row < NY;
          dy1a[r] = -t[r];
          dy2a[r] = hy ? (c.rbil[r] - y2r[r] * dy1a[r]) * iy1[r] : 0.0;
        CIMPC_ROWS_END
        const double a_aff = step_length2<D>(c.y1, c.y2, dy1a, dy2a, 1.0, l);
        CIMPC_ROWS(r, row)
          const bool hy = row < NY;
          pm[r] = hy ? c.y1[r] * c.y2[r] : 0.0;
          pa[r] = hy ? (c.y1[r] - a_aff * dy1a[r]) * (c.y2[r] - a_aff * dy2a[r]) : 0.0;
        CIMPC_ROWS_END
        const double mu = gsum2<GL>(pm[0], pm[1]) / (double)NY;
        const double mu_aff = gsum2<GL>(pa[0], pa[1]) / (double)NY;
        double sg = fmin(fmax(mu_aff / mu, 0.0), 1.0);
        sg = sg * sg * sg;
        const double kap = fmax(sg * mu, o.kappa_tol / o.undercut);

        // ---- corrector ----
        double rbc[2], dy1[2], dy2[2], dx[2];
        CIMPC_ROWS(r, row)
          const bool hy = row < NY;
          rbc[r] = hy ? fma(c.y1[r], c.y2[r], -kap) + dy1a[r] * dy2a[r] : 0.0;
          rhs[r] = hy ? cu[r] - (c.rrst[r] - c.ry2[r] * rbc[r] * iy1[r]) : 0.0;
        CIMPC_ROWS_END
        schur_solve2<D>(c, Ls, sc, l, rhs, t);
        CIMPC_ROWS(r, row)
          const bool hy = row < NY;
          dy1[r] = -t[r];
          dy2[r] = hy ? (rbc[r] - y2r[r] * dy1[r]) * iy1[r] : 0.0;
          dx[r] = au[r];
        CIMPC_ROWS_END
        {
          const double* AB = Ls + D::O_AIBC + l;
#pragma unroll 2
          for (int j = 0; j < D::NRP; j += 2) {
            const double2 tv = lds2(sc + S::O_TV + j);
            CIMPC_ROWS(r, row)
              dx[r] = fma(AB[r * GL + j * G], tv.x, dx[r]);
              dx[r] = fma(AB[r * GL + (j + 1) * G], tv.y, dx[r]);
            CIMPC_ROWS_END
          }
        }
        CIMPC_ROWS(r, row)
          if (!(row < NX)) dx[r] = 0.0;
        CIMPC_ROWS_END

        const double vmax = fmax(r_vio, k_vio);
        const double tau = fmax(1.0 - o.eps_min, 1.0 - vmax * vmax);
        double alpha = step_length2<D>(c.y1, c.y2, dy1, dy2, tau, l);

        bool acc = done;
        double xc[2], y1c[2], y2c[2], rdc[2], rrc[2], rbc2[2];
        double rvc = r_vio, kvc = k_vio;
        CIMPC_ROWS(r, row)
          xc[r] = c.x[r]; y1c[r] = c.y1[r]; y2c[r] = c.y2[r];
          rdc[r] = c.rdyn[r]; rrc[r] = c.rrst[r]; rbc2[r] = c.rbil[r];
        CIMPC_ROWS_END
#pragma unroll 1
        for (int ls = 0; ls <= o.max_ls; ++ls) {
          double xt[2], y1t[2], y2t[2], rd[2], rr[2], rb[2];
          CIMPC_ROWS(r, row)
            xt[r] = c.x[r] - alpha * dx[r];
            y1t[r] = c.y1[r] - alpha * dy1[r];
            y2t[r] = c.y2[r] - alpha * dy2[r];
          CIMPC_ROWS_END
          residual2<D>(Ls, sc, l, c, xt, y1t, y2t, 0.0, rd, rr, rb);
          const double rv = gmax2<GL>(fmax(fabs(rd[0]), fabs(rr[0])), fmax(fabs(rd[1]), fabs(rr[1])));
          const double kv = gmax2<GL>(fabs(rb[0]), fabs(rb[1]));
          if (!acc) {
            CIMPC_ROWS(r, row)
              xc[r] = xt[r]; y1c[r] = y1t[r]; y2c[r] = y2t[r];
              rdc[r] = rd[r]; rrc[r] = rr[r]; rbc2[r] = rb[r];
            CIMPC_ROWS_END
            rvc = rv;
            kvc = kv;
            if (rv <= r_vio || kv <= k_vio || ls == o.max_ls) acc = true;
            else alpha *= o.ls_scale;
          }
          if (__all_sync(FULL, acc)) break;
        }
        if (!done) {
          CIMPC_ROWS(r, row)
            c.x[r] = xc[r]; c.y1[r] = y1c[r]; c.y2[r] = y2c[r];
            c.rdyn[r] = rdc[r]; c.rrst[r] = rrc[r]; c.rbil[r] = rbc2[r];
          CIMPC_ROWS_END
          r_vio = rvc;
          k_vio = kvc;
          reg = reg_it;
          ++iters;
        }
      }
      const bool conv = (r_vio < o.r_tol) && (k_vio < o.kappa_tol);

      if (valid) {
        double* zo = p.z_out + prob * NZ;
        CIMPC_ROWS(r, row)
          if (row < NX) zo[row] = c.x[r];
          if (row < NY) {
            zo[NX + row] = c.y1[r];
            zo[NX + NY + row] = c.y2[r];
          }
        CIMPC_ROWS_END
        if (l == 0) {
          p.status[prob] = conv ? 1 : 0;
          p.iters[prob] = iters;
        }
      }
      if (o.diff_sol) {
        const double reg_d = fmax(reg, o.kappa_tol * o.gamma_reg);
        sensitivities2<D>(c, Ls, sc, l, gi, gshift, reg_d, p.dz_out + prob * (int64_t)(D::ND * D::NCOL), valid);
      }
      __syncwarp();
    }
    __syncthreads();
    cur = seg_end;
  }
}

#undef CIMPC_ROWS
#undef CIMPC_ROWS_END

}  // namespace cimpc
