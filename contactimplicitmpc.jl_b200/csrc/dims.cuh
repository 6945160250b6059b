// Compile-time sizes of one robot's linearized contact subproblem, and the layout of the
// per-knot constant block ("LinStore") in HBM.
//
// Sizes follow src/simulation/index.jl:371-390 (num_var, num_data) and
// src/controller/linearized_solver.jl:78-80 (nx = nq, ny = 2nc + nb) of the reference.
#pragma once
#include <cstdint>

namespace cimpc {

__host__ __device__ constexpr int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}
__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }
__host__ __device__ constexpr int round_up(int v, int m) { return (v + m - 1) / m * m; }

template <int NQ_, int NU_, int NW_, int NC_, int NB_, int MODE_>
struct Dims {
  static constexpr int NQ = NQ_, NU = NU_, NW = NW_, NC = NC_, NB = NB_, MODE = MODE_;
  static constexpr int NX = NQ;
  static constexpr int NY = 2 * NC + NB;
  static constexpr int NZ = NQ + 4 * NC + 2 * NB;
  static constexpr int NTH = 2 * NQ + NU + NW + 2;
  static constexpr int NCOL = 2 * NQ + NU;                      // q0, q1, u1 columns of δz
  static constexpr int NYD = MODE ? (NC + NB) : 0;              // y1 rows handed to Newton (γ1, b1)
  static constexpr int ND = NQ + NYD;                           // implicit_dynamics.jl:37-43
  // lanes cooperating on one subproblem: one lane per row of x / y1 / y2
  static constexpr int G = imax(4, pow2_ceil(imax(NX, NY)));
  static_assert(G <= 32, "subproblem rows must fit one warp");

  // ---- LinStore: per-knot constants, all column-major fp64, offsets in doubles ----
  static constexpr int O_DX = 0;                        // Dx      NX×NX   rz0[dyn, x]
  static constexpr int O_DY1 = O_DX + NX * NX;          // Dy1     NX×NY   rz0[dyn, y1]
  static constexpr int O_RX = O_DY1 + NX * NY;          // Rx      NY×NX   rz0[rst, x]
  static constexpr int O_RY1 = O_RX + NY * NX;          // Ry1     NY×NY   rz0[rst, y1]
  static constexpr int O_RY2 = O_RY1 + NY * NY;         // Ry2     NY      diag rz0[rst, y2]
  static constexpr int O_RTD = O_RY2 + NY;              // Rθdyn   NX×NTH  rθ0[dyn, :]
  static constexpr int O_RTR = O_RTD + NX * NTH;        // Rθrst   NY×NTH  rθ0[rst, :]
  static constexpr int O_CD = O_RTR + NY * NTH;         // cdyn    NX      rdyn0 − Dx x0 − Dy1 y10 − Rθdyn θ0
  static constexpr int O_CR = O_CD + NX;                // crst    NY      rrst0 − Rx x0 − Ry1 y10 − Ry2∘y20 − Rθrst θ0
  static constexpr int O_AI = O_CR + NY;                // Ai      NX×NX   Dx⁻¹                 (schur.jl:39)
  static constexpr int O_CAI = O_AI + NX * NX;          // CAi     NY×NX   Rx Dx⁻¹              (schur.jl:40)
  static constexpr int O_AIB = O_CAI + NY * NX;         // AiB     NX×NY   Dx⁻¹ Dy1
  static constexpr int O_S0 = O_AIB + NX * NY;          // S0      NY×NY   Ry1 − Rx Dx⁻¹ Dy1    (schur.jl:41,85)
  static constexpr int O_W = O_S0 + NY * NY;            // W       NY×NCOL CAi Rθdyn − Rθrst  (first NCOL columns)
  static constexpr int O_AR = O_W + NY * NCOL;          // AR      NX×NCOL Ai Rθdyn           (first NCOL columns)
  static constexpr int LIN_DOUBLES = O_AR + NX * NCOL;
  static constexpr int LIN_STRIDE = round_up(LIN_DOUBLES, 16);  // 128 B multiples (bulk-copy friendly)
};

// The robots of BASELINE.json's configs (SURVEY.md §2 dimension table).
//                         nq  nu nw nc nb
#define CIMPC_FOR_EACH_MODEL(X) \
  X(hopper2d, 4, 2, 2, 1, 2)    \
  X(quadruped, 11, 8, 2, 4, 8)  \
  X(flamingo, 9, 6, 2, 4, 8)    \
  X(centroidal, 18, 12, 3, 4, 16)

}  // namespace cimpc
