// Measures the fp64 FMA issue peak (DFMA) and a 64-bit shuffle / shared-broadcast rate on the GPU it
// runs on — the companion denominators for the fp64-bound ip_solve_kernel (MEASURED_PEAKS.json only
// holds HBM and bf16).  Prints one JSON line.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void shfl_kernel(double* out, int iters) {
  double x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __shfl_sync(0xffffffffu, x[i], (it + i) & 15, 16);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  if (s == 123.456) out[0] = s;
}

__global__ void lds_bcast_kernel(double* out, int iters) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  double s = 0;
  const double2* v = reinterpret_cast<const double2*>(sm);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      double2 t = v[(it * 8 + i) & 511];  // all lanes read the same 16 B: broadcast
      s += t.x * t.y;
    }
  }
  if (s == 123.456) out[0] = s;
}

// fp64 tensor-core rate: DMMA m8n8k4 (mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64; SASS `DMMA.884`), four independent
// accumulator fragments per warp.  One instruction = 8·8·4 FMAs = 512 flops per warp.
__global__ void dmma_kernel(double* out, int iters) {
  double c[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* d;
  cudaMalloc(&d, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 4096;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dfma_kernel<8><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double fl = 2.0 * blocks * threads * (double)iters * 8;
  const double dfma_tf = fl / (best * 1e-3) / 1e12;
  best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    shfl_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  // 64-bit shuffles (= 2 SHFL.32) per second, chip-wide, in warp-instructions
  const double shfl64_per_s = (double)blocks * (threads / 32) * (double)iters * 8 / (best * 1e-3);
  best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    lds_bcast_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double lds128_per_s = (double)blocks * (threads / 32) * (double)iters * 8 / (best * 1e-3);
  best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    dmma_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  const double dmma_tf = 2.0 * 256.0 * (double)blocks * (threads / 32) * (double)iters * 4 / (best * 1e-3) / 1e12;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_m8n8k4_tflops\": %.2f, "
         "\"shfl64_warp_instr_per_s\": %.4g, "
         "\"lds128_bcast_warp_instr_per_s\": %.4g, \"clock_khz_max\": %d, "
         "\"how\": \"8 independent DFMA chains/thread (4 independent DMMA.884 accumulator fragments/warp), "
         "8 CTAs/SM x 256 thr, best of 4 after warm-up\"}\n",
         p.name, p.multiProcessorCount, dfma_tf, dmma_tf, shfl64_per_s, lds128_per_s, clk);
  return 0;
}
