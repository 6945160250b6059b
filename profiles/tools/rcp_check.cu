// Accuracy of the branch-free reciprocal / division of csrc/fastdiv.cuh against the correctly rounded operations:
// mismatching results and max error in ulp over 2^24 random operands spread over 600 binades, the ratios the IP iteration
// actually forms (y/Δy with y in 1e-12..1e3), v/24, and the special values.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../contactimplicitmpc.jl_b200/csrc -o rcp_check rcp_check.cu
#include <cstdio>
#include <cstdint>
#include <cmath>
#include <cuda_runtime.h>
#define CIMPC_IP_FASTDIV 1
#include "fastdiv.cuh"
using namespace cimpc;
__device__ unsigned long long mix(unsigned long long i) {
  unsigned long long h = i * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32; return h;
}
__device__ double rnd(unsigned long long h, int emin, int espan) {
  const double m = 1.0 + (double)(h >> 12) * (1.0 / 4503599627370496.0);
  double a = ldexp(m, emin + (int)((h >> 3) % (unsigned)espan));
  return (h & 1) ? -a : a;
}
__device__ double ulps(double r, double ref) { return fabs(r - ref) / fabs(ldexp(1.0, ilogb(ref) - 52)); }
// out: [0] max ulp rcp, [1] seed rel err, [2] max ulp div (wide), [3] max ulp div (IP range), [4] max ulp /24; cnt: mismatches of the same four
__global__ void k(unsigned long long n, double* out, unsigned long long* cnt, double* specials) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  double m0 = 0, m1 = 0, m2 = 0, m3 = 0, m4 = 0;
  unsigned long long c0 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0;
  double m5 = 0;
  for (; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long h = mix(i), g = mix(i + 0x5555555555ull);
    const double a = rnd(h, -300, 600), b = rnd(g, -300, 600);
    const double r = rcp_fast(a), ref = __drcp_rn(a);
    m0 = fmax(m0, ulps(r, ref)); c0 += (r != ref);
    double s; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(a));
    m1 = fmax(m1, fabs(s * a - 1.0));
    const double aw = rnd(h, -150, 300), bw = rnd(g, -150, 300);  // quotient stays finite and normal
    const double q = div_fast(aw, bw), qr = __ddiv_rn(aw, bw);
    m2 = fmax(m2, ulps(q, qr)); c2 += (q != qr);
    const double y = fabs(rnd(h, -40, 50)), d = fabs(rnd(g, -40, 60));
    const double q2 = div_fast(y, d), q2r = __ddiv_rn(y, d);
    m3 = fmax(m3, ulps(q2, q2r)); c3 += (q2 != q2r);
    const double q3 = over_n<24>(y), q3r = __ddiv_rn(y, 24.0);
    m4 = fmax(m4, ulps(q3, q3r)); c4 += (q3 != q3r);
    // the clamp of differentiate_solution!: divisor = κ_tol·γ_reg for the usual tolerances (one fixed divisor, many numerators)
    const double clampv[4] = {1e-8 * 0.1, 1e-4 * 0.1, 2e-4 * 0.1, 1e-5 * 0.1};
    const double bc = clampv[i & 3];
    const double q4 = div_fast(y, bc), q4r = __ddiv_rn(y, bc);
    c5 += (q4 != q4r); m5 = fmax(m5, ulps(q4, q4r));
    const double r4 = rcp_fast(q4r), r4r = __drcp_rn(q4r);   // 1 / w of such a quotient
    c6 += (r4 != r4r);
  }
  atomicMax((unsigned long long*)&out[0], (unsigned long long)__double_as_longlong(m0));
  atomicMax((unsigned long long*)&out[1], (unsigned long long)__double_as_longlong(m1));
  atomicMax((unsigned long long*)&out[2], (unsigned long long)__double_as_longlong(m2));
  atomicMax((unsigned long long*)&out[3], (unsigned long long)__double_as_longlong(m3));
  atomicMax((unsigned long long*)&out[4], (unsigned long long)__double_as_longlong(m4));
  atomicAdd(&cnt[0], c0); atomicAdd(&cnt[1], c2); atomicAdd(&cnt[2], c3); atomicAdd(&cnt[3], c4); atomicAdd(&cnt[4], c5); atomicAdd(&cnt[5], c6);
  atomicMax((unsigned long long*)&out[5], (unsigned long long)__double_as_longlong(m5));
  if (blockIdx.x == 0 && threadIdx.x == 0) for (int t = 0; t < 4; ++t) { const double cv[4] = {1e-8 * 0.1, 1e-4 * 0.1, 2e-4 * 0.1, 1e-5 * 0.1}; specials[6 + t] = (rcp_fast(cv[t]) == __drcp_rn(cv[t])) ? 1.0 : 0.0; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    specials[0] = rcp_fast(0.0); specials[1] = rcp_fast(INFINITY); specials[2] = rcp_fast(1e-310); specials[3] = rcp_fast(1e300);
    specials[4] = div_fast(1.0, 0.0); specials[5] = div_fast(0.0, 3.0);
  }
}
int main() {
  double* d; unsigned long long* c;
  cudaMalloc(&d, 24 * sizeof(double)); cudaMemset(d, 0, 24 * sizeof(double));
  cudaMalloc(&c, 8 * sizeof(unsigned long long)); cudaMemset(c, 0, 8 * sizeof(unsigned long long));
  k<<<1184, 256>>>(1ull << 24, d, c, d + 8);
  double h[24]; unsigned long long hc[8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hc, c, sizeof(hc), cudaMemcpyDeviceToHost);
  printf("{\"samples\": %llu, \"rcp\": {\"mismatches_vs_drcp_rn\": %llu, \"max_ulp\": %.3f, \"seed_max_rel_error\": %.3e}, "
         "\"div_wide\": {\"mismatches_vs_ddiv_rn\": %llu, \"max_ulp\": %.3f}, \"div_ip_range\": {\"mismatches_vs_ddiv_rn\": %llu, \"max_ulp\": %.3f}, "
         "\"over_24\": {\"mismatches_vs_ddiv_rn\": %llu, \"max_ulp\": %.3f}, \"div_by_clamp\": {\"mismatches_vs_ddiv_rn\": %llu, \"max_ulp\": %.3f, \"rcp_of_quotient_mismatches\": %llu, \"rcp_of_clamp_exact\": [%g, %g, %g, %g]}, "
         "\"specials\": {\"rcp(0)\": \"%g\", \"rcp(inf)\": \"%g\", \"rcp(1e-310)\": \"%g\", \"rcp(1e300)\": \"%g\", \"1/0\": \"%g\", \"0/3\": \"%g\"}}\n",
         1ull << 24, hc[0], h[0], h[1], hc[1], h[2], hc[2], h[3], hc[3], h[4], hc[4], h[5], hc[5], h[14], h[15], h[16], h[17], h[8], h[9], h[10], h[11], h[12], h[13]);
  return 0;
}
