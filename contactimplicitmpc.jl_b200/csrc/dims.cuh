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

  // ---- reduced Schur complement -----------------------------------------------------------------
  // In S = Ry1 − diag(Ry2 ŷ2/ŷ1) − Rx Dx⁻¹ Dy1 the friction-cone rows `fri` (s2 − (μγ1 − E b1)) do not depend on q2
  // and the ψ1 columns do not enter the dynamics (src/simulation/simulation.jl:133-158: Rx[fri,:] = 0, Dy1[:,ψ] = 0,
  // Ry1[fri,ψ] = 0), so with y1 = [γ1; b1 | ψ1], rst = [imp; mdp | fri]
  //     S = [ A  B ]     A (NR×NR) iterate-dependent only on its diagonal,  B = Ry1[imp∪mdp, ψ] (= [0; −Eᵀ]) constant,
  //         [ C  D ]     C = Ry1[fri, γb] (= [−μI  E]) constant,            D = −diag(w_ψ),  w = Ry2 ŷ2/ŷ1,
  // and t_ψ is eliminated in closed form before the dense inversion (NR = nc + nb rows: 16 → 12 for the quadruped,
  // 24 → 20 centroidal).  Column ψ_c of S holds −w_ψ,c (row fri_c) and the entries B[l, c] of the rows l of contact c
  // (one non-zero per row); partial pivoting on that column picks the larger of the two, PER CONTACT AND ITERATE:
  //   w_ψ,c ≥ |B| ("D pivot", sticking contact):  row l ← A[l,:] + (B[l,c]/w_ψ,c) C[c,:],   t_ψ = (C t_r − r_ψ)/w_ψ
  //   w_ψ,c < |B| ("B pivot", sliding / open contact; b0 = first row of contact c, ρ_l = B[l,c]/B[b0,c]):
  //        row b0 ← C[c,:] + (w_ψ,c/B[b0,c]) A[b0,:],   row l ≠ b0 ← A[l,:] − ρ_l A[b0,:],   t_ψ = (r_b0 − A[b0,:] t_r)/B[b0,c]
  // so every multiplier is ≤ 1 — eliminating with 1/w_ψ alone loses cond ≈ 1/w_ψ digits on sliding contacts (measured:
  // 1e-5 instead of 1e-13 on one of 40 960 benchmark problems).  Every row is a combination of three CONSTANT rows
  // (A0[l,:], BC[l,:] = B[l,c]·C[c,:], AB[l,:] = A0[b0,:]) with three scalars, plus its diagonal terms.
  static constexpr int NPSI = NC;                                 // eliminated pairs (ψ1, s2)
  static constexpr int NR = NY - NPSI;                            // rows of the dense block that is inverted
  static constexpr int NRP = round_up(NR, 4);                     // padded with identity rows (Gauss-Jordan block of 4)
  static constexpr int NFR = 1 + NB / NC;                         // non-zeros of a row of C: γ_c and the b's of contact c
  static_assert(NRP <= G, "reduced rows must fit the lane group");

  // ---- LinStore: per-knot constants, fp64, offsets in doubles --------------------------------
  // Every lane-indexed matrix is padded to G rows and stored "lane fastest" so that the G lanes of a
  // group read consecutive words (conflict-free LDS); pairs that are always consumed together are
  // interleaved so that one 16-byte load feeds two FMAs.  Zero padding makes idle lanes harmless.
  //
  // Shared-memory part (copied per CTA by one cp.async.bulk when the knot changes):
  static constexpr int O_RES = 0;                         // (NX+NY)×G×2  {[Dx Dy1][l][j], [Rx Ry1][l][j]}
  static constexpr int O_CA2 = O_RES + (NX + NY) * G * 2;  // NX×G×2       {CAi[l][j], Ai[l][j]}
  static constexpr int O_AIBC = O_CA2 + NX * G * 2;        // NRP×G        AiB[l][j] at j*G + l   (ψ columns are zero)
  static constexpr int O_AIBR = O_AIBC + NRP * G;          // NX×G         AiB[i][l] at i*G + l
  static constexpr int O_S0 = O_AIBR + NX * G;             // NRP×G        A0[l][j]  at j*G + l   (identity padding)
  static constexpr int O_S0T = O_S0 + NRP * G;             // NRP×G        A0[j][l]  at j*G + l
  static constexpr int O_BC = O_S0T + NRP * G;             // NRP×G        BC[l][j]  at j*G + l
  static constexpr int O_BCT = O_BC + NRP * G;             // NRP×G        BC[j][l]  at j*G + l
  static constexpr int O_AB = O_BCT + NRP * G;             // NRP×G        AB[l][j] = A0[b0(l)][j] at j*G + l
  static constexpr int O_ABT = O_AB + NRP * G;             // NRP×G        AB[j][l]  at j*G + l
  static constexpr int O_CRW = O_ABT + NRP * G;            // NFR×G×2      k-th non-zero of row l−NR of C: {value, column} (ψ lanes)
  // per-lane scalars (G each): Ry2[l]; B[l,c(l)]; c(l) (−1: none); lane of b0(c(l)) (−1: none); ρ_l;
  // 1/B[b0,c] of the lane's contact; ψ lanes: lane of b0(c)
  static constexpr int O_RY2 = O_CRW + NFR * G * 2;
  static constexpr int O_BV = O_RY2 + G;
  static constexpr int O_CID = O_BV + G;
  static constexpr int O_B0 = O_CID + G;
  static constexpr int O_RHO = O_B0 + G;
  static constexpr int O_PB0 = O_RHO + G;
  static constexpr int O_IBV0 = O_PB0 + G;
  // per-row scalars read at uniform addresses (transposed build, transformed sensitivity right-hand sides)
  static constexpr int O_CIDU = O_IBV0 + G;                // NRP
  static constexpr int O_RHOU = O_CIDU + NRP;              // NRP
  static constexpr int O_B0U = O_RHOU + NRP;               // NRP          b0(c(j)), j itself for rows without ψ coupling
  static constexpr int O_W = O_B0U + NRP;                   // NCOL×NRP     W[k][c]   at c*NRP + k   (uniform reads)
  static constexpr int O_AR = O_W + NCOL * NRP;            // NCOL×G       AR[l][c]  at c*G + l
  static constexpr int ITER_DOUBLES = round_up(O_AR + NCOL * G, 2);  // what the iteration itself reads
  // Prologue part (touched once per subproblem: c = c0 + Rθ θ):
  static constexpr int O_C0 = ITER_DOUBLES;                // G×2          {cdyn0[l], crst0[l]}
  static constexpr int O_RTH = O_C0 + G * 2;               // NTH×G×2      {Rθdyn[l][j], Rθrst[l][j]}
  static constexpr int LIN_DOUBLES = O_RTH + NTH * G * 2;
  static constexpr int LIN_STRIDE = round_up(LIN_DOUBLES, 16);  // 128 B multiples
  // The instances that share a warp between subproblems (G <= 16: 9 KB of Rθ for the quadruped) stage the prologue part
  // too: every subproblem reads all of it, and from L1 / L2 that was 5.7 GB of loads per 655 360-subproblem sweep behind
  // `long_scoreboard` stalls.  The one-subproblem-per-warp instances (centroidal: 27 KB of Rθ on top of 96 KB) keep it in L2.
#ifndef CIMPC_STAGE_RTH
#define CIMPC_STAGE_RTH 1
#endif
  static constexpr bool RTH_IN_SMEM = CIMPC_STAGE_RTH && G <= 16;
  static constexpr int SMEM_DOUBLES = RTH_IN_SMEM ? round_up(LIN_DOUBLES, 2) : ITER_DOUBLES;
  static_assert((SMEM_DOUBLES * 8) % 16 == 0, "bulk copy size must be a multiple of 16 B");
};

// The robots of BASELINE.json's configs (SURVEY.md §2 dimension table).
//                         nq  nu nw nc nb
// The *_payload entries share the sizes (and therefore the solver kernels) of their nominal robot; what differs is the
// generated residual used by the simulator step and the device linearization (cimpc_create_named selects them).
// The *_piecewise entries are the planar robots on the terrain `piecewise1_2D_lc` (get_simulation(robot,
// "piecewise1_2D_lc", "piecewise", approx = true)): the generated residual evaluates the surface under every contact.
#define CIMPC_FOR_EACH_MODEL(X)        \
  X(hopper2d, 4, 2, 2, 1, 2)           \
  X(quadruped, 11, 8, 2, 4, 8)         \
  X(flamingo, 9, 6, 2, 4, 8)           \
  X(centroidal, 18, 12, 3, 4, 16)      \
  X(quadruped_payload, 11, 8, 2, 4, 8) \
  X(centroidal_payload, 18, 12, 3, 4, 16) \
  X(hopper2d_piecewise, 4, 2, 2, 1, 2) \
  X(flamingo_piecewise, 9, 6, 2, 4, 8) \
  X(quadruped_piecewise, 11, 8, 2, 4, 8)

}  // namespace cimpc
