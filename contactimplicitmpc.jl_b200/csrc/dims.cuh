// Compile-time sizes of one robot's linearized contact subproblem, and the layout of the
// per-knot constant block ("LinStore") in HBM / shared memory.
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
  static_assert(NX <= NY, "row padding below assumes nx <= ny");
  static_assert(NY % 4 == 0, "Gauss-Jordan block size");

  // ---- LinStore: per-knot constants, fp64, offsets in doubles --------------------------------
  // Every lane-indexed matrix is padded to G rows and stored "lane fastest" so that the G lanes of a
  // group read consecutive words (conflict-free LDS); pairs that are always consumed together are
  // interleaved so that one 16-byte load feeds two FMAs.  Zero padding makes idle lanes harmless.
  //
  // Shared-memory part (copied per CTA by one cp.async.bulk when the knot changes):
  static constexpr int O_RES = 0;                         // (NX+NY)×G×2  {[Dx Dy1][l][j], [Rx Ry1][l][j]}
  static constexpr int O_CA2 = O_RES + (NX + NY) * G * 2;  // NX×G×2       {CAi[l][j], Ai[l][j]}
  static constexpr int O_AIBC = O_CA2 + NX * G * 2;        // NY×G         AiB[l][j] at j*G + l
  static constexpr int O_AIBR = O_AIBC + NY * G;           // NX×G         AiB[i][l] at i*G + l
  static constexpr int O_S0 = O_AIBR + NX * G;             // NY×G         S0[l][j]  at j*G + l
  static constexpr int O_S0T = O_S0 + NY * G;              // NY×G         S0[j][l]  at j*G + l
  static constexpr int O_RY2 = O_S0T + NY * G;             // G            Ry2[l]
  static constexpr int O_W = O_RY2 + G;                    // NCOL×NY      W[k][c]   at c*NY + k   (uniform reads)
  static constexpr int O_AR = O_W + NCOL * NY;             // NCOL×G       AR[l][c]  at c*G + l
  static constexpr int SMEM_DOUBLES = round_up(O_AR + NCOL * G, 2);
  // Global-only part (touched once per subproblem, in the prologue; served by L1/L2):
  static constexpr int O_C0 = SMEM_DOUBLES;                // G×2          {cdyn0[l], crst0[l]}
  static constexpr int O_RTH = O_C0 + G * 2;               // NTH×G×2      {Rθdyn[l][j], Rθrst[l][j]}
  static constexpr int LIN_DOUBLES = O_RTH + NTH * G * 2;
  static constexpr int LIN_STRIDE = round_up(LIN_DOUBLES, 16);  // 128 B multiples
  static_assert((SMEM_DOUBLES * 8) % 16 == 0, "bulk copy size must be a multiple of 16 B");
};

// The robots of BASELINE.json's configs (SURVEY.md §2 dimension table).
//                         nq  nu nw nc nb
#define CIMPC_FOR_EACH_MODEL(X) \
  X(hopper2d, 4, 2, 2, 1, 2)    \
  X(quadruped, 11, 8, 2, 4, 8)  \
  X(flamingo, 9, 6, 2, 4, 8)    \
  X(centroidal, 18, 12, 3, 4, 16)

}  // namespace cimpc
