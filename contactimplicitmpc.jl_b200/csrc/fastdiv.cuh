// Reciprocal / division helpers of the solver kernels.
#pragma once

namespace cimpc {

// IEEE divisions and __drcp_rn cost this path more than their flops: nvcc expands a / b into MUFU.RCP64H + ~8 DFMA + a range
// check with an out-of-line slow path behind a BSSY / BRA pair (~25 instructions, and the branch serialises the chain the
// quotient sits on: every step length of an iteration), __drcp_rn into ~11.  Here:
//   rcp_fast(a)    = hardware seed (MUFU.RCP64H, <= 1e-6 relative) + two Newton steps: 5 instructions, no branch;
//   div_fast(a, b) = q = a·r, then Markstein's correction q + r·(a - b·q) with r = rcp_fast(b): 8 instructions, no branch
//                    (correctly rounded whenever r is the correctly rounded reciprocal).
// Measured on B200 with profiles/tools/rcp_check.cu on 2^24 random operands over 600 binades: rcp_fast equals __drcp_rn and
// div_fast equals the IEEE quotient BIT FOR BIT on every sample (profiles/r02_fastdiv.md) - the kernels keep their results.
// Special operands differ, and only where the solve has already failed: 0, +-inf and denormal divisors (flushed) give NaN
// instead of +-inf / 0, so a singular pivot or a collapsed slack still ends in the NaN exit of the iteration (status 0).
// -DCIMPC_IP_FASTDIV=0 restores the compiler's sequences.
#ifndef CIMPC_IP_FASTDIV
#define CIMPC_IP_FASTDIV 1
#endif
__device__ __forceinline__ double rcp_fast(double a) {
#if CIMPC_IP_FASTDIV
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
  double e = fma(-a, r, 1.0);
  r = fma(r, e, r);
  e = fma(-a, r, 1.0);
  return fma(r, e, r);
#else
  return __drcp_rn(a);
#endif
}
__device__ __forceinline__ double div_fast(double a, double b) {
#if CIMPC_IP_FASTDIV
  const double r = rcp_fast(b);
  const double q = a * r;
  return fma(fma(-q, b, a), r, q);
#else
  return a / b;
#endif
}
// v / N for a compile-time N (a multiplication when N is a power of two, else the corrected quotient)
template <int N>
__device__ __forceinline__ double over_n(double v) {
#if CIMPC_IP_FASTDIV
  constexpr double c = 1.0 / (double)N;
  if constexpr ((N & (N - 1)) == 0) return v * c;
  const double q = v * c;
  return fma(fma(-q, (double)N, v), c, q);
#else
  return v / (double)N;
#endif
}

}  // namespace cimpc
