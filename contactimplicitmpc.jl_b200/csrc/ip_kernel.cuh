// Batched Mehrotra predictor-corrector interior-point solve of the linearized contact subproblem
// (kernel v2: compact rolled loops, shared-memory exchange, TMA-staged per-knot constants).
//
// Mapping.  One group of G lanes (G = 4, 16 or 32; 32/G subproblems per warp) owns one subproblem;
// lane l holds row l of every vector (x_l, y1_l, y2_l, residual rows) and row l of the ny×ny inverse of
// the Schur complement in registers.  Vectors are exchanged between the rows of a group through a
// small per-group shared-memory scratch (one STS per lane, 16-byte broadcast LDS on the way back),
// scalars are reduced with width-G shuffles.  The per-knot constants (Dims::LinStore, ≈ 27 KB for the
// quadruped) are shared by every subproblem on the same reference knot: each CTA stages them once
// into shared memory with a single bulk-async copy (cp.async.bulk + mbarrier, the 1-D TMA path) and
// re-stages only when its contiguous slice of the batch crosses into the next knot.  Warps pull
// subproblem pairs from a CTA-local counter, so a slow subproblem only delays its own warp.
//
// Replaces, per subproblem (reference file:line):
//   rlin!                         src/controller/linearized_solver.jl:364-373     -> residual()
//   rzlin! + schur_factorize!     linearized_solver.jl:378-399, src/solver/schur.jl:80-88 -> load_schur() + invert()
//   linear_solve!(Δ, rz, r)       linearized_solver.jl:424-444, schur.jl:93-110   -> apply_inverse() + products
//   linear_solve!(δz, rz, rθ)     linearized_solver.jl:451-479                    -> sensitivities()
//   residual_/bilinear_violation  linearized_solver.jl:401-409                    -> group max-reductions
//   general_correction_term!      linearized_solver.jl:411-418
//   interior_point_solve!         RoboDojo 0.1.3 (external; iteration defined in oracle/ip.py)
// Linear algebra.  The reference re-factorises S = D − C A⁻¹ B (ny×ny) by modified Gram-Schmidt QR at
// every iteration and back-substitutes (src/solver/qr.jl:113-158).  Here S is INVERTED in place by
// Gauss-Jordan elimination with partial (row) pivoting — rows never leave their lanes, the pivot order
// is tracked — so both solves of an iteration and all sensitivity right-hand sides become
// matrix-vector / matrix-matrix products with no serial substitution chain.  Same solution to fp64
// round-off (tests/test_gpu_parity.py), see DESIGN.md for the flop / latency argument.
#pragma once
#include <cfloat>
#include <cstdint>
#include <utility>

#include "dims.cuh"
#include "fastdiv.cuh"
#include "../../include/cimpc_b200.h"

// Unroll factors of the product loops (summation order unchanged: only the loads move ahead of the FMA chains).
#ifndef CIMPC_UNROLL_CA
#define CIMPC_UNROLL_CA 11            // cu = CAi rdyn, au = Ai rdyn: was 1 (one exposed shared-memory latency per term)
#endif
#ifndef CIMPC_UNROLL_DX
#define CIMPC_UNROLL_DX 6
#endif
#ifndef CIMPC_UNROLL_SENS
#define CIMPC_UNROLL_SENS 2
#endif
#ifndef CIMPC_UNROLL_RES
#define CIMPC_UNROLL_RES 4            // residual: trips of two terms
#endif
#ifndef CIMPC_UNROLL_PRO
#define CIMPC_UNROLL_PRO 4            // prologue c = c0 + Rθ θ
#endif
#ifndef CIMPC_UNROLL_GJ
#define CIMPC_UNROLL_GJ 1             // Gauss-Jordan: trips of four columns (1 = rolled, the row rotated by four per trip)
#endif
#ifndef CIMPC_UNROLL_SENS_SCATTER
#define CIMPC_UNROLL_SENS_SCATTER 4
#endif

namespace cimpc {

// (the G = 32 instance keeps the rolled loops: at its 128-register cap the unrolled ones measured −0.8 %)
template <class D> struct Unroll {
  static constexpr bool R = D::G == 32;
  static constexpr int CA = R ? 1 : CIMPC_UNROLL_CA, DX = R ? 2 : CIMPC_UNROLL_DX, SENS = R ? 1 : CIMPC_UNROLL_SENS,
                       SENS_SCATTER = R ? 1 : CIMPC_UNROLL_SENS_SCATTER, RES = R ? 2 : CIMPC_UNROLL_RES,
                       PRO = R ? 2 : CIMPC_UNROLL_PRO, GJ = R ? 1 : CIMPC_UNROLL_GJ;
};

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
  // optional compaction (batched newton_solve!): slot s of the launch is subproblem (s / n_act) · R + act[s % n_act],
  // i.e. stage-major over the rollouts that still iterate; null = slot s is subproblem s
  const int32_t* __restrict__ act = nullptr;
  int32_t n_act = 0;
  int32_t R = 0;
  // device-driven compaction (the newton_solve! graph, api.cu): when `par` is set, the list of this sweep is
  // act2 + (*par)·R with cnt2[*par] entries and the launch covers stages·cnt2[*par] slots — nothing comes from the host
  const int32_t* __restrict__ act2 = nullptr;
  const int* __restrict__ cnt2 = nullptr;
  const int* __restrict__ par = nullptr;
  int32_t stages = 0;
  // alt holds one row per ROLLOUT (subproblem index mod R) instead of one per subproblem: `set_altitude!` gives every
  // stage of a policy the same offsets (src/controller/implicit_dynamics.jl:141-154)
  int32_t alt_by_rollout = 0;
};

constexpr unsigned FULL = 0xffffffffu;
constexpr int UNPIV = 1 << 20;
// Elimination of ψ1 (dims.cuh): the D pivot (−w_ψ) is kept while w_ψ ≥ PSI_PIVOT_RATIO·|B|; 1 = partial pivoting on the
// ψ column (every multiplier ≤ 1).  Measured with 1e-3 (multipliers up to 1e3): agreement with the LU oracle unchanged at
// κ_tol ≥ 1e-5, but iteration counts at κ_tol = 1e-8 start to depend on 1e-13 perturbations of the linearization.
constexpr double PSI_PIVOT_RATIO = 1.0;

// Guaranteed compile-time unrolling (register arrays must never be indexed dynamically).
template <int B, int... Is, class F>
__device__ __forceinline__ void static_for_impl(std::integer_sequence<int, Is...>, F&& f) {
  (f(std::integral_constant<int, B + Is>{}), ...);
}
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (E > B) static_for_impl<B>(std::make_integer_sequence<int, E - B>{}, static_cast<F&&>(f));
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
// Group max / min of NON-NEGATIVE doubles (violation norms, step lengths) by integer REDUX: the bit pattern of a
// non-negative double is monotone in its value, so the maximum is the lexicographic maximum of (high word, low word) —
// two masked REDUX per word and warp instead of log2(G) rounds of two 32-bit shuffles, a compare and two selects
// (≈ 14 instead of ≈ 30 instructions per reduction, six reductions per iteration, and a shorter dependent chain).
// Same value, same bits.  A NaN (high word 0x7ff8…) wins the maximum, i.e. it reaches r_vio / κ_vio and ends the
// iteration through the NaN exit, where a shuffle / fmax chain would have dropped it.
#ifndef CIMPC_IP_REDUX_MINMAX
#define CIMPC_IP_REDUX_MINMAX 1
#endif
template <int G, bool MAX>
__device__ __forceinline__ unsigned gredux(unsigned v, int gshift) {
  // inactive value: 0 for max, ~0 for min
  constexpr unsigned NEUTRAL = MAX ? 0u : 0xffffffffu;
  if constexpr (G == 32) {
    return MAX ? __reduce_max_sync(FULL, v) : __reduce_min_sync(FULL, v);
  } else {
    static_assert(G == 16, "two groups per warp");
    const unsigned a = gshift == 0 ? v : NEUTRAL, b = gshift == 0 ? NEUTRAL : v;
    const unsigned ra = MAX ? __reduce_max_sync(FULL, a) : __reduce_min_sync(FULL, a);
    const unsigned rb = MAX ? __reduce_max_sync(FULL, b) : __reduce_min_sync(FULL, b);
    return gshift == 0 ? ra : rb;
  }
}
template <int G, bool MAX>
__device__ __forceinline__ double gminmax_nonneg(double v, int gshift) {
  const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
  const unsigned mh = gredux<G, MAX>(hi, gshift);
  const unsigned ml = gredux<G, MAX>(hi == mh ? lo : (MAX ? 0u : 0xffffffffu), gshift);
  return __hiloint2double((int)mh, (int)ml);
}
template <int G>
__device__ __forceinline__ double gmax_nn(double v, int gshift) {
#if CIMPC_IP_REDUX_MINMAX
  if constexpr (G >= 16) return gminmax_nonneg<G, true>(v, gshift);
#endif
  return gmax<G>(v);
}
template <int G>
__device__ __forceinline__ double gmin_nn(double v, int gshift) {
#if CIMPC_IP_REDUX_MINMAX
  if constexpr (G >= 16) return gminmax_nonneg<G, false>(v, gshift);
#endif
  return gmin<G>(v);
}
template <int G>
__device__ __forceinline__ double gsum(double v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
  return v;
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---- bulk-async (TMA 1-D) staging of one knot's constants ------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // make the init visible to the async proxy
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Per-group shared-memory scratch (doubles).
template <class D>
struct GroupScratch {
  static constexpr int NX = D::NX, NY = D::NY, NRP = D::NRP;
  static constexpr int XYN = round_up(imax(NX + NY, D::NTH), 2);
  static constexpr int O_XY = 0;                 // [x; y1] exchange (θ in the prologue, rdyn for the products)
  static constexpr int O_WV = O_XY + XYN;        // RHS of a solve, permuted to pivot order
  static constexpr int O_TV = O_WV + NRP;        // solution of a solve, natural order
  static constexpr int O_PROW = O_TV + NRP;      // pivot row of a Gauss-Jordan step, double-buffered by step parity
  static constexpr int O_V = O_PROW + 2 * NRP;   // {1/w_ψ,c, w_ψ,c} of the eliminated pairs
  static constexpr int O_G = O_V + 2 * D::NPSI;  // rhs_ψ,c of the solve in flight
  static constexpr int O_VR = O_G + round_up(D::NPSI, 2);  // {a1_j, a2_j, a3_j, ·} per ROW j (transposed load)
  static constexpr int O_SENS = O_VR + 4 * NRP;  // sensitivity workspace
  static constexpr int LDP = NRP + 1;            // padded leading dimension of lane-major scratch matrices
  static constexpr int O_AIBP = O_SENS;          // NX×NRP   AiB with columns permuted to pivot order
  static constexpr int A_SZ = imax(NX * NRP, D::MODE ? NRP * LDP : 0);  // AiBp, later aliased by S_red⁻¹ rows
  static constexpr int O_P1 = O_SENS + round_up(A_SZ, 2);    // NX×LDP  P1 = AiB S_red⁻¹
  static constexpr int O_PL = O_P1 + round_up(NX * LDP, 2);  // NRP ints: pivot lane of each step
  static constexpr int DOUBLES = round_up(O_PL + (NRP + 1) / 2, 2);
};

template <class D>
struct Ctx {  // per-lane registers of one subproblem
  double x, y1, y2;
  double cdyn, crst, ry2;
  double rdyn, rrst, rbil;
  // elimination of ψ1 (dims.cuh): constants of this lane's row, and the pivot choice of the current iterate
  double bv;        // B[l, c(l)]: coupling of reduced row l to its eliminated ψ (0: none)
  int cid;          // c(l), −1: none
  int b0;           // lane of the first row of contact c(l) (reduced lanes) / of contact l − NR (ψ lanes); −1: none
  double wl;        // w_l = Ry2 ŷ2 / ŷ1 of this lane's pair at the current iterate
  double a1, a2, a3;  // row l of S_red = a1·A[l,:] + a2·BC[l,:] + a3·A[b0,:]   (a2 doubles as v_ψ = 1/w_ψ on ψ lanes)
  bool dpiv;        // this lane's contact is eliminated with the D pivot
  bool any_b;       // group-uniform: some contact of this subproblem uses the B pivot
  double M[D::NRP]; // row of P·S_red⁻¹ in pivot-order column slots, WITHOUT its pivot scaling (see invert)
  double msc;       // 1 / pivot of this lane's row: the row of the inverse is msc · M[]
  int mystep;       // Gauss-Jordan step at which this lane's row was the pivot row
};

// rlin!: rows of [rdyn; rrst; rbil] at (x, y1, y2) with central-path parameter kappa.
template <class D>
__device__ __forceinline__ void residual(const double* __restrict__ Ls, double* __restrict__ sc, int l, bool hx,
                                         bool hy, double cdyn, double crst, double ry2, double x, double y1,
                                         double y2, double kappa, double& rdyn, double& rrst, double& rbil) {
  constexpr int NX = D::NX, NY = D::NY, G = D::G;
  using S = GroupScratch<D>;
  if (hx) sc[S::O_XY + l] = x;
  if (hy) sc[S::O_XY + NX + l] = y1;
  __syncwarp();
  // two accumulator pairs: the fp64 FMA chains are the latency that matters here
  double ad = cdyn, ar = fma(ry2, y2, crst), ad2 = 0.0, ar2 = 0.0;
  const double* R = Ls + D::O_RES + 2 * l;
  constexpr int NJ = NX + NY;
#pragma unroll Unroll<D>::RES
  for (int j = 0; j + 1 < NJ; j += 2) {
    const double2 v = lds2(sc + S::O_XY + j);
    const double2 c0 = lds2(R + (j)*G * 2), c1 = lds2(R + (j + 1) * G * 2);
    ad = fma(c0.x, v.x, ad);
    ar = fma(c0.y, v.x, ar);
    ad2 = fma(c1.x, v.y, ad2);
    ar2 = fma(c1.y, v.y, ar2);
  }
  if constexpr (NJ & 1) {
    const double v = sc[S::O_XY + NJ - 1];
    const double2 c0 = lds2(R + (NJ - 1) * G * 2);
    ad = fma(c0.x, v, ad);
    ar = fma(c0.y, v, ar);
  }
  ad += ad2;
  ar += ar2;
  __syncwarp();
  rdyn = hx ? ad : 0.0;
  rrst = hy ? ar : 0.0;
  rbil = hy ? fma(y1, y2, -kappa) : 0.0;
}

// In-place Gauss-Jordan inversion with partial pivoting of the matrix whose row l is c.M[] (lane l).
// Columns are processed four at a time with static register indices, then the row is rotated by four
// so that the loop body is the same code for every block (compact, I-cache resident).
// On exit lane l = p_k (the pivot lane of step k = c.mystep) holds row k of (QA)⁻¹ = A⁻¹Qᵀ:
//   c.msc · c.M[j] = A⁻¹[mystep][p_j].
// The pivot row is never normalised in place: scaling a row commutes with every later row operation on it,
// so the factor 1/pivot is kept aside (c.msc) and applied to the dot products that use the row.  The pivot
// lane then takes part in the uniform update with factor 0 — no per-step clearing of its 2·NY registers
// (that clearing was 38 % of the instructions of an inversion; +4 % throughput, the kernel is bound by the
// shared-memory broadcast, not by issue).  Eight columns per rolled trip instead of four was measured
// slower (58.2 vs 59.9 M subproblems/s).
template <class D, bool RECORD_PL>
__device__ __forceinline__ void invert(Ctx<D>& c, double* __restrict__ sc, int l, int gshift, bool hy) {
  constexpr int NY = D::NRP, G = D::G;  // the reduced block (dims.cuh); `hy` = lane owns one of its rows
  using S = GroupScratch<D>;
  c.mystep = UNPIV;
  c.msc = 0.0;
#pragma unroll Unroll<D>::GJ
  for (int kb = 0; kb < NY; kb += 4) {
    static_for<0, 4>([&](auto U) {
      constexpr int u = decltype(U)::value;
      double* const prow = sc + S::O_PROW + (u & 1) * NY;
      // partial pivoting on the leading 32 bits of |a| (monotone for non-negative doubles): the chosen
      // pivot is within 2^-20 of the column maximum, which is all that stability needs
      const bool unp = (c.mystep == UNPIV) && hy;
      // every lane forms the reciprocal of ITS candidate while the pivot search is in flight; the winner
      // publishes it in the pivot slot of the row (that slot is read for nothing else), which takes the
      // MUFU + Newton-step latency of 1/pivot off the critical path between publish and update
      const double myinv = rcp_fast(c.M[u]);
      const unsigned cand = unp ? ((unsigned)__double2hiint(c.M[u]) & 0x7fffffffu) + 1u : 0u;
      // group maximum: one REDUX per group of the warp (independent, so one REDUX latency) instead of a
      // log2(G)-deep shuffle chain on the critical path of every elimination step
      unsigned m;
      if constexpr (G == 32) {
        m = __reduce_max_sync(FULL, cand);
      } else if constexpr (G == 16) {
        const unsigned m0 = __reduce_max_sync(FULL, gshift == 0 ? cand : 0u);
        const unsigned m1 = __reduce_max_sync(FULL, gshift == 0 ? 0u : cand);
        m = gshift == 0 ? m0 : m1;
      } else {
        m = cand;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o, G));
      }
      unsigned bal = __ballot_sync(FULL, cand == m);
      bal = (G == 32) ? bal : ((bal >> gshift) & ((1u << (G & 31)) - 1u));
      const int p = __ffs(bal) - 1;
      const bool is_p = (l == p);
      if (is_p) {  // publish the pivot row
        c.mystep = kb + u;
        if (RECORD_PL) reinterpret_cast<int*>(sc + S::O_PL)[kb + u] = l;
        static_for<0, NY / 2>([&](auto J) {
          constexpr int j = decltype(J)::value;
          *reinterpret_cast<double2*>(prow + 2 * j) =
              make_double2(2 * j == u ? myinv : c.M[2 * j], 2 * j + 1 == u ? myinv : c.M[2 * j + 1]);
        });
        c.msc = myinv;
      }
      __syncwarp();
      const double pinv = prow[u];
      // update factor: other rows −a_iu / piv; the pivot row itself stays (its scaling is deferred)
      const double g = is_p ? 0.0 : -c.M[u] * pinv;
      static_for<0, NY / 2>([&](auto J) {
        constexpr int j = decltype(J)::value;
        const double2 pr = lds2(prow + 2 * j);
        if constexpr (2 * j != u) c.M[2 * j] = fma(g, pr.x, c.M[2 * j]);
        if constexpr (2 * j + 1 != u) c.M[2 * j + 1] = fma(g, pr.y, c.M[2 * j + 1]);
      });
      c.M[u] = is_p ? 1.0 : g;  // new pivot-column entry (unscaled 1/piv on the pivot row)
      // no second __syncwarp: the next step publishes into the OTHER buffer, and a lane reaches the publish of
      // step u + 2 only through the __syncwarp of step u + 1, i.e. after every lane has read this buffer
    });
    // rotate the row left by four column slots
    const double t0 = c.M[0], t1 = c.M[1], t2 = c.M[2], t3 = c.M[3];
    static_for<0, NY - 4>([&](auto J) {
      constexpr int j = decltype(J)::value;
      c.M[j] = c.M[j + 4];
    });
    c.M[NY - 4] = t0;
    c.M[NY - 3] = t1;
    c.M[NY - 2] = t2;
    c.M[NY - 1] = t3;
  }
}

// t = S⁻¹ w with the inverse held as c.M: w is scattered to pivot order, every lane forms its dot
// product, the result is scattered back to natural order.  Returns t_l; sc[O_TV + j] = t_j for all j.
template <class D>
__device__ __forceinline__ double apply_inverse(const Ctx<D>& c, double* __restrict__ sc, int l, bool hy, double w,
                                                double& raw) {
  constexpr int NY = D::NRP;
  using S = GroupScratch<D>;
  if (hy) sc[S::O_WV + c.mystep] = w;
  __syncwarp();
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  static_for<0, NY / 4>([&](auto J) {
    constexpr int j = decltype(J)::value;
    const double2 v = lds2(sc + S::O_WV + 4 * j), w = lds2(sc + S::O_WV + 4 * j + 2);
    a0 = fma(c.M[4 * j], v.x, a0);
    a1 = fma(c.M[4 * j + 1], v.y, a1);
    a2 = fma(c.M[4 * j + 2], w.x, a2);
    a3 = fma(c.M[4 * j + 3], w.y, a3);
  });
  raw = (a0 + a1) + (a2 + a3);  // lanes outside the block: product of their passenger row (load_schur) with the RHS
  a0 = raw * c.msc;
  if (hy) sc[S::O_TV + c.mystep] = a0;
  __syncwarp();
  return hy ? sc[S::O_TV + l] : 0.0;
}

template <class D>
__device__ __forceinline__ double step_length(bool hy, double y1, double y2, double d1, double d2, double tau, int gshift) {
  // largest α ≤ 1 with y − αΔ ≥ (1−τ) y on the orthant (fraction to the boundary)
  double a = 1.0;
  if (hy && d1 > 0.0) a = fmin(a, div_fast(tau * y1, d1));
  if (hy && d2 > 0.0) a = fmin(a, div_fast(tau * y2, d2));
  return gmin_nn<D::G>(a, gshift);  // a in (0, 1]
}

// rzlin!: row l of S_red (or of S_redᵀ), the Schur complement with ψ1 eliminated (dims.cuh):
//   row l = a1·(A0[l,:] − w_l e_l) + a2·BC[l,:] + a3·(A0[b0,:] − w_b0 e_b0)
// with (a1, a2, a3) = (1, 0, 0) rows without ψ coupling; (1, 1/w_ψ, 0) D pivot; (w_ψ/B_l, 1/B_l, 0) B pivot, l = b0;
// (1, 0, −ρ_l) B pivot, l ≠ b0.  Only one of a2, a3 is non-zero, so the row is built from TWO constant rows in one
// pass for every case: a1·A0[l,:] + ax·aux[l,:], aux = BC or AB selected per lane.
// The ψ lanes (which own no row of the block) load A[b0,:] of their contact as PASSENGER rows: they are never pivots, the
// Gauss-Jordan updates turn them into −A[b0,:] S_red⁻¹, and apply_inverse then hands them −A[b0,:] t_r — the t_ψ of the
// B pivot — at no extra instruction.
template <class D, bool TRANSPOSED>
__device__ __forceinline__ void load_schur(Ctx<D>& c, const double* __restrict__ Ls, double* __restrict__ sc, int l,
                                           int gshift, bool hy, double reg) {
  constexpr int NR = D::NR, NRP = D::NRP, G = D::G;
  using S = GroupScratch<D>;
  const bool hr = l < NR, hm = l < NRP, hp = hy && l >= NR;
  const double y1r = fmax(c.y1, reg), y2r = fmax(c.y2, reg);
  c.wl = hy ? div_fast(c.ry2 * y2r, y1r) : 1.0;
  const double iwl = rcp_fast(c.wl);
  if (hp) *reinterpret_cast<double2*>(sc + S::O_V + 2 * (l - NR)) = make_double2(iwl, c.wl);
  __syncwarp();
  c.a1 = 1.0; c.a2 = 0.0; c.a3 = 0.0;
  c.dpiv = true;
  if (hr && c.cid >= 0) {
    const double2 vw = lds2(sc + S::O_V + 2 * c.cid);  // {1/w_ψ, w_ψ}
    c.dpiv = vw.y * fabs(Ls[D::O_IBV0 + l]) >= PSI_PIVOT_RATIO;  // same expression on the ψ lane
    if (c.dpiv) {
      c.a2 = vw.x;
    } else if (l == c.b0) {
      c.a2 = Ls[D::O_IBV0 + l];  // 1 / B[b0, c]: this lane is b0 (same bits as 1.0 / c.bv, prep_kernel divides the same pair)
      c.a1 = vw.y * c.a2;
    } else {
      c.a3 = -Ls[D::O_RHO + l];
    }
  }
  if (hp) {  // ψ lanes: pivot choice of their own contact, a2 = 1/w_ψ
    c.dpiv = c.wl * fabs(Ls[D::O_IBV0 + l]) >= PSI_PIVOT_RATIO;
    c.a2 = iwl;
  }
  {  // does THIS subproblem have a contact on its B pivot (group-uniform)
    const unsigned bal = __ballot_sync(FULL, !c.dpiv);
    c.any_b = ((G == 32) ? bal : ((bal >> gshift) & ((1u << (G & 31)) - 1u))) != 0u;
  }
  // w of the first row of this lane's contact (own value on lanes without one)
  const double wb0 = __shfl_sync(FULL, c.wl, gshift + (c.b0 >= 0 ? c.b0 : l));
  if constexpr (!TRANSPOSED) {
    const bool use_ab = (hr && c.a3 != 0.0) || (hp && NRP == NR);     // rows l ≠ b0 of a B-pivot contact; passengers
    const double a1 = hr ? c.a1 : ((hm && NRP != NR) ? 1.0 : 0.0);      // padding rows keep their identity row of A0
    const double ax = hr ? (c.a3 != 0.0 ? c.a3 : c.a2) : ((hp && NRP == NR) ? 1.0 : 0.0);
    const double* A0 = Ls + D::O_S0 + l;
    const double* AX = Ls + (use_ab ? D::O_AB : D::O_BC) + l;
    const int jb = use_ab ? c.b0 : -1;
    const double dl = hr ? a1 * c.wl : 0.0, db = ax * wb0;
    static_for<0, NRP>([&](auto J) {
      constexpr int j = decltype(J)::value;
      double m = fma(a1, A0[j * G], ax * AX[j * G]);
      if (j == l) m -= dl;
      if (j == jb) m -= db;
      c.M[j] = m;
    });
  } else {
    if (hm) {
      *reinterpret_cast<double2*>(sc + S::O_VR + 4 * l) = make_double2(hr ? c.a1 : 1.0, hr ? c.a2 : 0.0);
      sc[S::O_VR + 4 * l + 2] = hr ? c.a3 : 0.0;
    }
    __syncwarp();
    const double* A0 = Ls + D::O_S0T + l;
    const double* BC = Ls + D::O_BCT + l;
    const double* AB = Ls + D::O_ABT + l;
    // column l of S_red: element j comes from row j
    const bool is_b0_bpiv = hr && c.cid >= 0 && !c.dpiv && l == c.b0;
    static_for<0, NRP>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 a12 = lds2(sc + S::O_VR + 4 * j);
      const double a3 = sc[S::O_VR + 4 * j + 2];
      double m = a12.x * A0[j * G];
      m = fma(a12.y, BC[j * G], m);
      m = fma(a3, AB[j * G], m);
      if (j == l && hr) m = fma(-c.a1, c.wl, m);  // own diagonal: row l = column l
      // rows j ≠ b0 of a B-pivot contact carry −a3_j·w_b0 in column b0: this lane IS b0, so w_b0 = w_l
      if (is_b0_bpiv && j != l && (int)Ls[D::O_CIDU + j] == c.cid) m = fma(Ls[D::O_RHOU + j], c.wl, m);
      c.M[j] = hm ? m : 0.0;
    });
  }
}

// t = S⁻¹ rhs for the full ny-vector rhs (lane l holds rhs_l) through the reduced inverse held in c.M:
//   rhs_red,l = a1·rhs_l + a2·B_l·rhs_ψ + a3·rhs_b0,   t_r = S_red⁻¹ rhs_red,
//   t_ψ = (C t_r − rhs_ψ)/w_ψ (D pivot)   or   (rhs_b0 − A[b0,:] t_r)/B[b0,c] (B pivot).
// Returns t_l; sc[O_TV + j] = t_j for the reduced rows j.
template <class D>
__device__ __forceinline__ double schur_solve(const Ctx<D>& c, const double* __restrict__ Ls, double* __restrict__ sc,
                                              int l, int gshift, bool hy, double rhs) {
  constexpr int NR = D::NR, NRP = D::NRP, G = D::G;
  using S = GroupScratch<D>;
  const bool hr = l < NR, hm = l < NRP, hp = hy && l >= NR;
  if (hp) sc[S::O_G + l - NR] = rhs;
  __syncwarp();
  const double rb0 = __shfl_sync(FULL, rhs, gshift + (c.b0 >= 0 ? c.b0 : l));  // rhs of the contact's first row
  double rr = 0.0;
  if (hr) {
    rr = c.a1 * rhs;
    if (c.cid >= 0) rr = (c.a3 != 0.0) ? fma(c.a3, rb0, rr) : fma(c.a2 * c.bv, sc[S::O_G + c.cid], rr);
  }
  double raw;
  const double tr = apply_inverse<D>(c, sc, l, hm, rr, raw);
  double tpsi_b;  // B pivot: t_ψ = (rhs_b0 − A[b0,:] t_r) / B[b0,c]
  if constexpr (NRP == NR) {
    tpsi_b = (rb0 + raw) * Ls[D::O_IBV0 + l];  // raw = −A[b0,:] t_r on the ψ lanes (passenger rows)
  } else {
    // (hopper sizes: the ψ lane doubles as the identity padding row of the block, so the product is formed explicitly)
    double e0 = hr ? fma(c.wl, tr, rhs) : 0.0, e1 = 0.0;
    const double* A0 = Ls + D::O_S0 + l;
#pragma unroll
    for (int j = 0; j < NRP; j += 2) {
      const double2 tv = lds2(sc + S::O_TV + j);
      e0 = fma(-A0[j * G], tv.x, e0);
      e1 = fma(-A0[(j + 1) * G], tv.y, e1);
    }
    tpsi_b = __shfl_sync(FULL, e0 + e1, gshift + (c.b0 >= 0 ? c.b0 : l)) * Ls[D::O_IBV0 + l];
  }
  double a = -rhs;
#pragma unroll
  for (int k = 0; k < D::NFR; ++k) {
    const double cv = Ls[D::O_CRW + (2 * k) * G + l];
    const int cj = (int)Ls[D::O_CRW + (2 * k + 1) * G + l];
    a = fma(cv, sc[S::O_TV + cj], a);
  }
  const double tpsi = c.dpiv ? c.a2 * a : tpsi_b;
  return hr ? tr : (hp ? tpsi : 0.0);
}

// A contact on its B pivot has its rows of the reduced system COMBINED (load_schur): S_red t_r = T·rhs with
// T = diag(a1) + Σ_k a3_k e_k e_b0(k)ᵀ.  For the sensitivities the right-hand sides are the constant W, so T is folded into
// the row vectors that multiply W instead:  row ← row·T  (own row of this lane in shared memory, in place: slot b0 is
// the only target of additions and the rows k with a3_k ≠ 0 have a1_k = 1).  Identity for D-pivot contacts.
template <class D>
__device__ __forceinline__ void fold_row_combination(const double* __restrict__ Ls, const double* __restrict__ sc,
                                                     double* __restrict__ row) {
  using S = GroupScratch<D>;
#pragma unroll 1
  for (int k = 0; k < D::NRP; ++k) {
    const double a1 = sc[S::O_VR + 4 * k];
    if (a1 != 1.0) row[k] *= a1;
  }
#pragma unroll 1
  for (int k = 0; k < D::NRP; ++k) {
    const double a3 = sc[S::O_VR + 4 * k + 2];
    if (a3 != 0.0) {
      const int b0 = (int)Ls[D::O_B0U + k];
      row[b0] = fma(a3, row[k], row[b0]);
    }
  }
}

// differentiate_solution! restricted to the rows / columns Newton consumes:
//   δx  = −(AR + AiB S⁻¹ W)          rows 1:nx      (δq0, δq1, δu1 blocks of q2)
//   δy1 =  S⁻¹ W                      rows nx+1:nd   (γ1, b1; :configurationforce only)
// with W = CAi Rθdyn − Rθrst and AR = Ai Rθdyn precomputed per knot.  Sᵀ is inverted so that lane q_k
// holds COLUMN k of S⁻¹; P1 = AiB S⁻¹ then needs no cross-lane sums, and the final products run with
// lane = output row and broadcast (uniform-address) loads of W.
template <class D>
__device__ __forceinline__ void sensitivities(Ctx<D>& c, const double* __restrict__ Ls, double* __restrict__ sc, int l,
                                              int gshift, bool hx, bool hy, double reg, double* __restrict__ dzo,
                                              bool valid) {
  // Only the reduced rows enter: AiB[:, ψ] = 0 and W[ψ, (q0, q1, u1)] = 0 (checked by prep_kernel), so
  // δx = −(AR + AiB_r S_red⁻¹ W_r) and δ[γ1; b1] = S_red⁻¹ W_r need neither t_ψ nor the ψ rows of W.
  constexpr int NX = D::NX, NY = D::NRP, G = D::G, NCOL = D::NCOL, ND = D::ND, NYD = D::NYD;
  static_assert(NYD == 0 || NYD == D::NR, "force rows = reduced rows");
  using S = GroupScratch<D>;
  const bool hm = l < NY;
  load_schur<D, true>(c, Ls, sc, l, gshift, hy, reg);
  invert<D, (NYD > 0)>(c, sc, l, gshift, hm);  // c.M[j] = S_red⁻¹[q_j][k],  k = c.mystep, lane = q_k
  // AiBp[i][m] = AiB[i][q_m]: lane l = q_m owns column l of AiB
  if (hm) {
#pragma unroll Unroll<D>::SENS_SCATTER
    for (int i = 0; i < NX; ++i) sc[S::O_AIBP + i * NY + c.mystep] = Ls[D::O_AIBR + i * G + l];
  }
  __syncwarp();
  // P1[i][k] = Σ_m AiBp[i][m] · S_red⁻¹[q_m][k]
#pragma unroll Unroll<D>::SENS
  for (int i = 0; i < NX; ++i) {
    double a0 = 0.0, a1 = 0.0;
    static_for<0, NY / 2>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 v = lds2(sc + S::O_AIBP + i * NY + 2 * j);
      a0 = fma(c.M[2 * j], v.x, a0);
      a1 = fma(c.M[2 * j + 1], v.y, a1);
    });
    if (hm) sc[S::O_P1 + i * S::LDP + c.mystep] = (a0 + a1) * c.msc;
  }
  __syncwarp();
  double Ar[NYD > 0 ? NY : 1];
  if constexpr (NYD > 0) {
    // rows of S_red⁻¹ for the force rows: Ainv[r][k] = S_red⁻¹[r][k], scattered from the column-held inverse
    const int* pl = reinterpret_cast<const int*>(sc + S::O_PL);
    if (hm) {
      static_for<0, NY>([&](auto J) {
        constexpr int j = decltype(J)::value;
        sc[S::O_AIBP + pl[j] * S::LDP + c.mystep] = c.M[j] * c.msc;
      });
    }
    __syncwarp();
    if (c.any_b && l < NYD) fold_row_combination<D>(Ls, sc, sc + S::O_AIBP + l * S::LDP);
    static_for<0, NY>([&](auto J) {
      constexpr int j = decltype(J)::value;
      Ar[j] = (l < NYD) ? sc[S::O_AIBP + l * S::LDP + j] : 0.0;
    });
  }
  // row l of P1 (c.M is dead: reuse its registers)
  if (c.any_b && hx) fold_row_combination<D>(Ls, sc, sc + S::O_P1 + l * S::LDP);
  static_for<0, NY>([&](auto J) {
    constexpr int j = decltype(J)::value;
    c.M[j] = hx ? sc[S::O_P1 + l * S::LDP + j] : 0.0;
  });
  __syncwarp();
#pragma unroll Unroll<D>::SENS
  for (int col = 0; col < NCOL; ++col) {
    const double* Wc = Ls + D::O_W + col * NY;
    double a0 = Ls[D::O_AR + col * G + l], a1 = 0.0, b0 = 0.0, b1 = 0.0;
    static_for<0, NY / 2>([&](auto J) {
      constexpr int j = decltype(J)::value;
      const double2 w = lds2(Wc + 2 * j);
      a0 = fma(c.M[2 * j], w.x, a0);
      a1 = fma(c.M[2 * j + 1], w.y, a1);
      if constexpr (NYD > 0) {
        b0 = fma(Ar[2 * j], w.x, b0);
        b1 = fma(Ar[2 * j + 1], w.y, b1);
      }
    });
    if (valid) {
      if (hx) dzo[col * ND + l] = -(a0 + a1);
      if constexpr (NYD > 0) {
        if (l < NYD) dzo[col * ND + NX + l] = b0 + b1;
      }
    }
  }
}

template <class D, int THREADS>
struct KernelSmem {
  static constexpr int GROUPS = THREADS / D::G;
  static constexpr int GS = round_up(GroupScratch<D>::DOUBLES, 2);
  static constexpr size_t BYTES = (size_t)(D::SMEM_DOUBLES + GROUPS * GS) * 8 + 64;
};

// Two CTAs of 256 threads per SM (128 registers) for the instances that share a warp between subproblems; measured
// alternatives on the quadruped batch: 2 × 320 threads at 96 registers 61.9 M/s, 2 × 288 61.6, 3 × 192 59.6, against
// 63.5 M/s here (the shared-memory pipe, not occupancy, is the limiter) and 44 M/s with one CTA of 148 registers.
template <class D, int THREADS>
__global__ void __launch_bounds__(THREADS, (D::G == 32 ? 1 : 2)) ip_solve_kernel(const IpParams p) {
  constexpr int NX = D::NX, NY = D::NY, NTH = D::NTH, NZ = D::NZ, NC = D::NC;
  constexpr int G = D::G, PPW = 32 / G;
  using S = GroupScratch<D>;
  using KS = KernelSmem<D, THREADS>;

  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Ls = reinterpret_cast<double*>(smem_raw);                      // staged knot constants
  double* scratch = Ls + D::SMEM_DOUBLES;                                // per-group scratch
  uint64_t* bar = reinterpret_cast<uint64_t*>(scratch + KS::GROUPS * KS::GS);
  int* ctl = reinterpret_cast<int*>(bar + 1);  // [0] next subproblem of the segment, [1] segment end, [3] first active

  const int tid = threadIdx.x, lane = tid & 31;
  const int l = lane % G, gi = lane / G, gshift = gi * G;
  const bool hx = l < NX, hy = l < NY;
  double* sc = scratch + (size_t)(tid / G) * KS::GS;
  const cimpc_ip_opts o = p.o;
  const double kappa_floor = o.kappa_tol / o.undercut;  // κ never goes below this (loop-invariant)

  // slot → subproblem (identity unless the launch is compacted).  32-bit arithmetic on purpose: a 64-bit division is
  // an out-of-line call in SASS, and with that call present this kernel computed garbage (bisected on B200, nvcc 12.9).
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
    const unsigned s = (unsigned)slot, t = s / (unsigned)n_act;  // a compacted launch has < 2^31 slots
    return (int64_t)t * nroll + act[s - t * (unsigned)n_act];
  };
  // this CTA's contiguous slice of the batch
  const int64_t c0 = n_slots * (int64_t)blockIdx.x / gridDim.x, c1 = n_slots * (int64_t)(blockIdx.x + 1) / gridDim.x;
  if (tid == 0) mbar_init(bar, 1);
  __syncthreads();
  int staged = -1;
  uint32_t phase = 0;

  for (int64_t cur = c0; cur < c1;) {
    // ---- segment = maximal run of subproblems on the same reference knot; knot −1 marks an inactive
    //      subproblem (finished rollout): it belongs to any segment and is skipped, outputs untouched ----
    if (tid == 0) {
      ctl[3] = (int)(c1 - c0);  // first active subproblem at or after cur
      ctl[1] = (int)(c1 - c0);  // segment end
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
    if (first >= c1) break;  // nothing left to do in this CTA's slice (CTA-uniform)
    const int kraw = p.knot[sub(first)];
    // a knot outside the uploaded reference is a caller bug: its subproblems are not solved, status 0, iters 0
    // (the host entry point rejects it up front; the device entry point cannot look at device memory)
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
    if (bad_knot) {  // CTA-uniform: flag the whole run and move on
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
    if (kn != staged) {  // CTA-uniform
      if (tid == 0) {
        mbar_expect_tx(bar, (uint32_t)(D::SMEM_DOUBLES * 8));
        bulk_g2s(Ls, p.lin + (int64_t)kn * D::LIN_STRIDE, (uint32_t)(D::SMEM_DOUBLES * 8), bar);
      }
      while (!mbar_try_wait(bar, phase)) {
      }
      phase ^= 1u;
      staged = kn;
    }
    // prologue constants: staged with the rest where they fit (dims.cuh), else read through L1 / L2
    const double* __restrict__ Lg = D::RTH_IN_SMEM ? Ls : p.lin + (int64_t)kn * D::LIN_STRIDE;

    // ---- warps pull PPW subproblems at a time from the segment ----
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
      const int64_t pi = valid ? prob : sub(first);  // idle groups shadow an active subproblem

      Ctx<D> c;
      // prologue: θ-dependent constants  c = c0 + Rθ θ (+ alt on the impact rows)
#pragma unroll 1
      for (int j = l; j < NTH; j += G) sc[S::O_XY + j] = p.theta[pi * NTH + j];
      __syncwarp();
      {
        const double2 c00 = reinterpret_cast<const double2*>(Lg + D::O_C0)[l];
        double cd = c00.x, cr = c00.y;
        const double2* R = reinterpret_cast<const double2*>(Lg + D::O_RTH) + l;
#pragma unroll Unroll<D>::PRO
        for (int j = 0; j < NTH; ++j) {
          const double2 r = R[j * G];
          const double t = sc[S::O_XY + j];
          cd = fma(r.x, t, cd);
          cr = fma(r.y, t, cr);
        }
        if (p.alt != nullptr && l < NC) {
          const int64_t arow = p.alt_by_rollout ? (int64_t)((unsigned)pi % (unsigned)nroll) : pi;  // 32-bit on purpose
          cr += p.alt[arow * NC + l];
        }
        c.cdyn = cd;
        c.crst = cr;
      }
      __syncwarp();
      c.ry2 = Ls[D::O_RY2 + l];
      c.bv = Ls[D::O_BV + l];
      c.cid = (int)Ls[D::O_CID + l];
      c.b0 = (l < D::NR) ? (int)Ls[D::O_B0 + l] : (int)Ls[D::O_PB0 + l];
      // cold start: z = 1, z[q2] = q2_init  (z_initialize!)
      c.x = hx ? p.q2_init[pi * NX + l] : 0.0;
      c.y1 = 1.0;
      c.y2 = 1.0;
      residual<D>(Ls, sc, l, hx, hy, c.cdyn, c.crst, c.ry2, c.x, c.y1, c.y2, 0.0, c.rdyn, c.rrst, c.rbil);
      double r_vio = gmax_nn<G>(fmax(fabs(c.rdyn), fabs(c.rrst)), gshift);
      double k_vio = gmax_nn<G>(fabs(c.rbil), gshift);

      bool done = !valid;
      int iters = 0;
      double reg = 0.0;

#pragma unroll 1
      for (int it = 0; it < o.max_iter; ++it) {
        if (r_vio < o.r_tol && k_vio < o.kappa_tol) done = true;
        if (!(r_vio == r_vio) || !(k_vio == k_vio)) done = true;  // NaN (singular pivot): give up, status 0
        if (__all_sync(FULL, done)) break;

        const double reg_it = (k_vio < o.kappa_reg) ? k_vio * o.gamma_reg : 0.0;
        const double y1r = fmax(c.y1, reg_it), y2r = fmax(c.y2, reg_it);
        // rzlin! + schur_factorize!  →  S⁻¹
        load_schur<D, false>(c, Ls, sc, l, gshift, hy, reg_it);
        invert<D, false>(c, sc, l, gshift, l < D::NRP);

        // constant products with u = rdyn (shared by predictor and corrector): cu = CAi u, au = Ai u
        double cu = 0.0, au = 0.0;
        if (hx) sc[S::O_XY + l] = c.rdyn;
        __syncwarp();
        {
          const double* CA = Ls + D::O_CA2 + 2 * l;
#pragma unroll Unroll<D>::CA
          for (int j = 0; j < NX; ++j) {
            const double2 k2 = lds2(CA + j * G * 2);
            const double ub = sc[S::O_XY + j];
            cu = fma(k2.x, ub, cu);
            au = fma(k2.y, ub, au);
          }
        }
        __syncwarp();

        // ---- predictor (affine) direction: only Δy1, Δy2 are needed ----
        // (the five divisions by ŷ1 of one iteration share one correctly-rounded reciprocal)
        const double iy1 = rcp_fast(y1r);
        double t = schur_solve<D>(c, Ls, sc, l, gshift, hy, hy ? cu - (c.rrst - c.ry2 * c.rbil * iy1) : 0.0);
        const double dy1a = -t;
        const double dy2a = hy ? (c.rbil - y2r * dy1a) * iy1 : 0.0;
        const double a_aff = step_length<D>(hy, c.y1, c.y2, dy1a, dy2a, 1.0, gshift);
        const double mu = over_n<NY>(gsum<G>(hy ? c.y1 * c.y2 : 0.0));
        const double mu_aff = over_n<NY>(gsum<G>(hy ? (c.y1 - a_aff * dy1a) * (c.y2 - a_aff * dy2a) : 0.0));
        double sg = fmin(fmax(div_fast(mu_aff, mu), 0.0), 1.0);
        sg = sg * sg * sg;
        const double kap = fmax(sg * mu, kappa_floor);

        // ---- corrector: rbil = y1∘y2 − κ + Δy1aff∘Δy2aff ----
        const double rbc = hy ? fma(c.y1, c.y2, -kap) + dy1a * dy2a : 0.0;
        t = schur_solve<D>(c, Ls, sc, l, gshift, hy, hy ? cu - (c.rrst - c.ry2 * rbc * iy1) : 0.0);
        const double dy1 = -t;
        const double dy2 = hy ? (rbc - y2r * dy1) * iy1 : 0.0;
        double dx = au;
        {
          const double* AB = Ls + D::O_AIBC + l;
#pragma unroll Unroll<D>::DX
          for (int j = 0; j < D::NRP; j += 2) {  // AiB[:, ψ] = 0: the reduced part of t is all Δx needs
            const double2 tv = lds2(sc + S::O_TV + j);
            dx = fma(AB[j * G], tv.x, dx);
            dx = fma(AB[(j + 1) * G], tv.y, dx);
          }
        }
        if (!hx) dx = 0.0;

        const double vmax = fmax(r_vio, k_vio);
        const double tau = fmax(1.0 - o.eps_min, 1.0 - vmax * vmax);
        double alpha = step_length<D>(hy, c.y1, c.y2, dy1, dy2, tau, gshift);

        // ---- candidate + back-tracking on the violations (trial max_ls is accepted unconditionally) ----
        double xc, y1c, y2c, rdc, rrc, rbc2, rvc, kvc;
        if constexpr (D::G < 32) {
          // Trial 0 is formed in the candidate registers themselves; the copy-in of a later trial is needed on ≈ 1 % of
          // the iterations only (same trials, same tests, same order as the loop over ls = 0 .. max_ls below; +1.2 %).
          xc = c.x - alpha * dx; y1c = c.y1 - alpha * dy1; y2c = c.y2 - alpha * dy2;
          residual<D>(Ls, sc, l, hx, hy, c.cdyn, c.crst, c.ry2, xc, y1c, y2c, 0.0, rdc, rrc, rbc2);
          rvc = gmax_nn<G>(fmax(fabs(rdc), fabs(rrc)), gshift);
          kvc = gmax_nn<G>(fabs(rbc2), gshift);
          bool acc = done || rvc <= r_vio || kvc <= k_vio || o.max_ls == 0;
          if (!acc) alpha *= o.ls_scale;
          if (!__all_sync(FULL, acc)) {
#pragma unroll 1
            for (int ls = 1; ls <= o.max_ls; ++ls) {
              const double xt = c.x - alpha * dx, y1t = c.y1 - alpha * dy1, y2t = c.y2 - alpha * dy2;
              double rd, rr, rb;
              residual<D>(Ls, sc, l, hx, hy, c.cdyn, c.crst, c.ry2, xt, y1t, y2t, 0.0, rd, rr, rb);
              const double rv = gmax_nn<G>(fmax(fabs(rd), fabs(rr)), gshift);
              const double kv = gmax_nn<G>(fabs(rb), gshift);
              if (!acc) {
                xc = xt; y1c = y1t; y2c = y2t; rdc = rd; rrc = rr; rbc2 = rb; rvc = rv; kvc = kv;
                if (rv <= r_vio || kv <= k_vio || ls == o.max_ls) acc = true;
                else alpha *= o.ls_scale;
              }
              if (__all_sync(FULL, acc)) break;
            }
          }
        } else {
          // (one subproblem per warp, 20-wide rows: a second inlined copy of the residual costs more in spills than the
          // copies it saves — measured −1.4 % — so this instance keeps the single call site)
          bool acc = done;
          xc = c.x; y1c = c.y1; y2c = c.y2; rdc = c.rdyn; rrc = c.rrst; rbc2 = c.rbil; rvc = r_vio; kvc = k_vio;
#pragma unroll 1
          for (int ls = 0; ls <= o.max_ls; ++ls) {
            const double xt = c.x - alpha * dx, y1t = c.y1 - alpha * dy1, y2t = c.y2 - alpha * dy2;
            double rd, rr, rb;
            residual<D>(Ls, sc, l, hx, hy, c.cdyn, c.crst, c.ry2, xt, y1t, y2t, 0.0, rd, rr, rb);
            const double rv = gmax_nn<G>(fmax(fabs(rd), fabs(rr)), gshift);
            const double kv = gmax_nn<G>(fabs(rb), gshift);
            if (!acc) {
              xc = xt; y1c = y1t; y2c = y2t; rdc = rd; rrc = rr; rbc2 = rb; rvc = rv; kvc = kv;
              if (rv <= r_vio || kv <= k_vio || ls == o.max_ls) acc = true;
              else alpha *= o.ls_scale;
            }
            if (__all_sync(FULL, acc)) break;
          }
        }
        if (!done) {
          c.x = xc; c.y1 = y1c; c.y2 = y2c; c.rdyn = rdc; c.rrst = rrc; c.rbil = rbc2; r_vio = rvc; k_vio = kvc;
          reg = reg_it;
          ++iters;
        }
      }
      const bool conv = (r_vio < o.r_tol) && (k_vio < o.kappa_tol);

      // ---- outputs: z*, status, iteration count ----
      if (valid) {
        double* zo = p.z_out + prob * NZ;
        if (hx) zo[l] = c.x;
        if (hy) {
          zo[NX + l] = c.y1;
          zo[NX + NY + l] = c.y2;
        }
        if (l == 0) {
          p.status[prob] = conv ? 1 : 0;
          p.iters[prob] = iters;
        }
      }
      // ---- differentiate_solution! ----
      if (o.diff_sol) {
        const double reg_d = fmax(reg, o.kappa_tol * o.gamma_reg);
        sensitivities<D>(c, Ls, sc, l, gshift, hx, hy, reg_d, p.dz_out + prob * (int64_t)(D::ND * D::NCOL), valid);
      }
      __syncwarp();
    }
    __syncthreads();  // every warp is done with this segment (and with Ls) before the next one is staged
    cur = seg_end;
  }
}

}  // namespace cimpc
