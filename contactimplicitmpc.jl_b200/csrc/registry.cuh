// Runtime view of one compiled (robot, mode) kernel instance.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "ip_kernel.cuh"
#ifdef CIMPC_WITH_IP_V3
#include "ip_kernel2.cuh"  // measured-and-rejected two-rows-per-lane variant (profiles/r02_ip_two_rows_per_lane.md)
#endif
#include "lin_kernel.cuh"
#include "newton_kernel.cuh"
#include "newton_cta.cuh"
#include "newton_general.cuh"
#include "newton_dense.cuh"
#include "sim_kernel.cuh"

namespace cimpc {

struct LinLayout {  // runtime copy of Dims<...> (offsets in doubles)
  int nx, ny, nz, nth, ncol, nd, group, nr, nrp, nfr;
  int o_res, o_ca2, o_aibc, o_aibr, o_s0, o_s0t, o_bc, o_bct, o_ab, o_abt, o_crw, o_ry2, o_bv, o_cid, o_b0, o_rho, o_pb0, o_ibv0,
      o_cidu, o_rhou, o_b0u, o_w, o_ar, o_c0, o_rth;
  int smem_doubles, stride;
};

struct ModelEntry {
  const char* name;
  cimpc_model_desc desc;
  LinLayout lay;
  cudaError_t (*launch)(const IpParams& p, int sm_count, cudaStream_t s);
  cudaError_t (*occupancy)(int* blocks_per_sm);
  // device Newton (:configuration instances only, else null)
  cudaError_t (*newton_reset)(const NewtonParams& p, const NewtonCall* call, cudaStream_t s);
  cudaError_t (*newton_step)(const NewtonParams& p, double* lscratch, cudaStream_t s);
  size_t (*newton_scratch)(int H);  // doubles of global scratch per rollout
  // general device Newton (both modes; [0] TrackingObjective, [1] TrackingVelocityObjective)
  cudaError_t (*newton_step_g[2])(const NewtonParams& p, double* lscratch, cudaStream_t s);
  size_t (*newton_scratch_g[2])(int H);
  // dense-weight device Newton (:configuration instances only; [0] without / [1] with velocity cost)
  cudaError_t (*newton_step_d[2])(const NewtonParams& p, double* lscratch, cudaStream_t s);
  size_t (*newton_scratch_d[2])(int H);
  // simulator step (generated residual of this robot)
  cudaError_t (*sim_step)(const SimParams& p, cudaStream_t s);
  size_t (*sim_scratch)(int R);  // doubles
  // device linearization (generated r, rz, rθ of this robot at the reference knots)
  cudaError_t (*linearize)(const LinEvalParams& p, cudaStream_t s);
};

constexpr int NEWTON_THREADS = 32 * NEWTON_WARPS;

// Opt-in to more than 48 KB of dynamic shared memory.  The attribute is PER DEVICE and a process may hold contexts on
// several GPUs (and drive them from several host threads: rollout.py::GroupedRollouts), so the largest size configured
// so far is kept per device of the calling thread, under a lock.
constexpr int MAX_DEVICES = 64;
struct SmemOptIn {
  std::mutex mu;
  size_t configured[MAX_DEVICES] = {};
  template <class K>
  cudaError_t ensure(K kernel, size_t bytes) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= MAX_DEVICES) return cudaErrorInvalidDevice;
    std::lock_guard<std::mutex> lk(mu);
    if (bytes <= configured[dev] || bytes <= 48 * 1024) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) configured[dev] = bytes;
    return e;
  }
};

template <class D>
cudaError_t launch_newton_reset(const NewtonParams& p, const NewtonCall* call, cudaStream_t s) {
  newton_reset_kernel<D, NEWTON_THREADS><<<p.R, NEWTON_THREADS, 0, s>>>(p, call);
  return cudaGetLastError();
}

// CIMPC_NEWTON_KERNEL=warp selects the round-1 warp-per-rollout kernel (A/B measurements; same arithmetic up to the
// summation order of the residual's δzᵀν terms).
inline bool newton_use_cta_kernel() {
  static const bool cta = [] {
    const char* v = getenv("CIMPC_NEWTON_KERNEL");
    return !(v && v[0] == 'w');
  }();
  return cta;
}

template <class D>
cudaError_t launch_newton_step(const NewtonParams& p, double* lscratch, cudaStream_t s) {
  if constexpr (NewtonCtaSmem<D>::SUPPORTED) {
    // one CTA per rollout with all of its δz in shared memory; longer horizons take the warp-per-rollout kernel below
    const size_t cbytes = (size_t)NewtonCtaSmem<D>::doubles(p.H) * sizeof(double);
    if (newton_use_cta_kernel() && cbytes <= NEWTON_CTA_MAX_SMEM) {
      static SmemOptIn copt;
      if (cudaError_t e = copt.ensure(newton_step_cta_kernel<D>, cbytes); e != cudaSuccess) return e;
      newton_step_cta_kernel<D><<<p.R, NEWTON_CTA_THREADS, cbytes, s>>>(p, lscratch);
      return cudaGetLastError();
    }
  }
  const size_t bytes = (size_t)NewtonSmem<D>::per_warp(p.H) * NEWTON_WARPS * sizeof(double);
  static SmemOptIn opt;
  if (cudaError_t e = opt.ensure(newton_step_kernel<D, NEWTON_THREADS>, bytes); e != cudaSuccess) return e;
  const int grid = (p.R + NEWTON_WARPS - 1) / NEWTON_WARPS;
  newton_step_kernel<D, NEWTON_THREADS><<<grid, NEWTON_THREADS, bytes, s>>>(p, lscratch);
  return cudaGetLastError();
}

template <class D, bool VEL>
cudaError_t launch_newton_step_general(const NewtonParams& p, double* lscratch, cudaStream_t s) {
  const size_t bytes = (size_t)NewtonGSmem<D, VEL>::per_warp(p.H) * NEWTON_WARPS * sizeof(double);
  static SmemOptIn opt;
  if (cudaError_t e = opt.ensure(newton_step_general_kernel<D, VEL, NEWTON_THREADS>, bytes); e != cudaSuccess) return e;
  const int grid = (p.R + NEWTON_WARPS - 1) / NEWTON_WARPS;
  newton_step_general_kernel<D, VEL, NEWTON_THREADS><<<grid, NEWTON_THREADS, bytes, s>>>(p, lscratch);
  return cudaGetLastError();
}

template <class D, bool VEL>
cudaError_t launch_newton_step_dense(const NewtonParams& p, double* lscratch, cudaStream_t s) {
  const size_t bytes = (size_t)NewtonDSmem<D, VEL>::per_warp(p.H) * NEWTON_WARPS * sizeof(double);
  static SmemOptIn opt;
  if (cudaError_t e = opt.ensure(newton_step_dense_kernel<D, VEL, NEWTON_THREADS>, bytes); e != cudaSuccess) return e;
  const int grid = (p.R + NEWTON_WARPS - 1) / NEWTON_WARPS;
  newton_step_dense_kernel<D, VEL, NEWTON_THREADS><<<grid, NEWTON_THREADS, bytes, s>>>(p, lscratch);
  return cudaGetLastError();
}

template <class D, bool VEL>
size_t newton_scratch_dense_doubles(int H) {
  return NewtonDSmem<D, VEL>::l_doubles(H);
}

template <class D, bool VEL>
size_t newton_scratch_general_doubles(int H) {
  return NewtonGSmem<D, VEL>::l_doubles(H);
}

template <class D>
size_t newton_scratch_doubles(int H) {
  return NewtonSmem<D>::l_doubles(H);
}

// One CTA of GEN::NS warps per 32-rollout tile; SIM_MIN_CTAS caps the registers (65536 / (256 · 2) → 128; the register LU holds a 28-double row).
constexpr int SIM_MIN_CTAS = 2;

template <class GEN>
cudaError_t launch_sim_step(const SimParams& p, cudaStream_t s) {
  const int grid = (p.R + 31) / 32;
  const size_t bytes = SimLayout<GEN>::SMEM_DOUBLES * sizeof(double);
  static SmemOptIn opt;
  if (cudaError_t e = opt.ensure(sim_step_kernel<GEN, SIM_MIN_CTAS>, bytes); e != cudaSuccess) return e;
  sim_step_kernel<GEN, SIM_MIN_CTAS><<<grid, GEN::NS * 32, bytes, s>>>(p);
  return cudaGetLastError();
}

template <class GEN>
size_t sim_scratch_doubles(int R) {
  return (size_t)((R + 31) / 32) * SimLayout<GEN>::TILE * 32;
}

template <class GEN>
cudaError_t launch_linearize(const LinEvalParams& p, cudaStream_t s) {
  const size_t bytes = (size_t)2 * GEN::NTRIG * 32 * sizeof(double);
  static_assert((size_t)2 * GEN::NTRIG * 32 * sizeof(double) <= 48 * 1024, "trig table must fit default shared memory");
  linearize_kernel<GEN><<<(p.H + 31) / 32, GEN::NS * 32, bytes, s>>>(p);
  return cudaGetLastError();
}

// CTA size of ip_solve_kernel.  One-subproblem-per-warp instances (G = 32: centroidal, 74 KB of staged constants) use
// 512 threads: one CTA per SM either way, but 16 instead of 8 resident warps (the register cap of 128 costs some spills).
#ifndef CIMPC_IP_THREADS_SHARED
#define CIMPC_IP_THREADS_SHARED 256  // CTA size of the instances that share a warp between subproblems (G < 32)
#endif
#ifndef CIMPC_IP_THREADS_G32
#define CIMPC_IP_THREADS_G32 512  // CTA size of the one-subproblem-per-warp instances (G = 32)
#endif
template <class D>
constexpr int ip_threads() { return D::G == 32 ? CIMPC_IP_THREADS_G32 : CIMPC_IP_THREADS_SHARED; }

template <class D>
LinLayout layout_of() {
  LinLayout l;
  l.nx = D::NX; l.ny = D::NY; l.nz = D::NZ; l.nth = D::NTH; l.ncol = D::NCOL; l.nd = D::ND; l.group = D::G;
  l.nr = D::NR; l.nrp = D::NRP; l.nfr = D::NFR;
  l.o_res = D::O_RES; l.o_ca2 = D::O_CA2; l.o_aibc = D::O_AIBC; l.o_aibr = D::O_AIBR; l.o_s0 = D::O_S0;
  l.o_s0t = D::O_S0T; l.o_bc = D::O_BC; l.o_bct = D::O_BCT; l.o_crw = D::O_CRW; l.o_ry2 = D::O_RY2; l.o_bv = D::O_BV;
  l.o_cid = D::O_CID; l.o_ab = D::O_AB; l.o_abt = D::O_ABT; l.o_b0 = D::O_B0; l.o_rho = D::O_RHO; l.o_pb0 = D::O_PB0;
  l.o_ibv0 = D::O_IBV0; l.o_cidu = D::O_CIDU; l.o_rhou = D::O_RHOU; l.o_b0u = D::O_B0U; l.o_w = D::O_W; l.o_ar = D::O_AR; l.o_c0 = D::O_C0; l.o_rth = D::O_RTH;
  l.smem_doubles = D::SMEM_DOUBLES; l.stride = D::LIN_STRIDE;
  return l;
}

// ---- kernel v3 (ip_kernel2.cuh): two rows per lane, for the instances with G >= 16 ------------------------------
// Measured slower than v2 (0.80-0.90x, profiles/r02_ip_two_rows_per_lane.md), so it is compiled only with
// -DCIMPC_WITH_IP_V3 (CIMPC_EXTRA_NVCC_FLAGS); then CIMPC_IP_KERNEL=v2 / v3 selects the kernel per launch
// (scripts/gpu_ip_ab.py).
#ifndef CIMPC_IP2_THREADS_16
#define CIMPC_IP2_THREADS_16 384  // G = 16 (quadruped, flamingo): 8 lanes per subproblem, 48 subproblems per CTA
#endif
#ifndef CIMPC_IP2_MINCTAS_16
#define CIMPC_IP2_MINCTAS_16 1
#endif
#ifndef CIMPC_IP2_THREADS_32
#define CIMPC_IP2_THREADS_32 256  // G = 32 (centroidal): 16 lanes per subproblem, 16 subproblems per CTA
#endif
#ifndef CIMPC_IP_DEFAULT_VERSION
#define CIMPC_IP_DEFAULT_VERSION 2
#endif
template <class D>
constexpr bool ip2_supported() {
#ifdef CIMPC_WITH_IP_V3
  return D::G >= 16 && D::NRP == D::NR;
#else
  return false;
#endif
}
template <class D>
constexpr int ip2_threads() { return D::G == 32 ? CIMPC_IP2_THREADS_32 : CIMPC_IP2_THREADS_16; }
template <class D>
constexpr int ip2_minctas() { return D::G == 32 ? 1 : CIMPC_IP2_MINCTAS_16; }
inline int ip_kernel_version() {  // read per launch (a launch costs more than a getenv): tests switch it inside one process
  const char* e = getenv("CIMPC_IP_KERNEL");
  if (e && e[0] == 'v' && (e[1] == '2' || e[1] == '3')) return e[1] - '0';
  return CIMPC_IP_DEFAULT_VERSION;
}

template <class D>
cudaError_t prepare_ip() {  // opt in to > 48 KB of dynamic shared memory (once per instance and device)
  static SmemOptIn opt;
  cudaError_t e = opt.ensure(ip_solve_kernel<D, ip_threads<D>()>, KernelSmem<D, ip_threads<D>()>::BYTES);
#ifdef CIMPC_WITH_IP_V3
  if constexpr (ip2_supported<D>()) {
    static SmemOptIn opt2;
    if (e == cudaSuccess)
      e = opt2.ensure(ip_solve2_kernel<D, ip2_threads<D>(), ip2_minctas<D>()>, KernelSmem2<D, ip2_threads<D>()>::BYTES);
  }
#endif
  return e;
}

template <class D>
cudaError_t occupancy_ip(int* blocks_per_sm) {
  cudaError_t e = prepare_ip<D>();
  if (e != cudaSuccess) return e;
#ifdef CIMPC_WITH_IP_V3
  if constexpr (ip2_supported<D>()) {
    if (ip_kernel_version() == 3)
      return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm,
                                                           ip_solve2_kernel<D, ip2_threads<D>(), ip2_minctas<D>()>,
                                                           ip2_threads<D>(), KernelSmem2<D, ip2_threads<D>()>::BYTES);
  }
#endif
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ip_solve_kernel<D, ip_threads<D>()>, ip_threads<D>(),
                                                       KernelSmem<D, ip_threads<D>()>::BYTES);
}

template <class D>
cudaError_t launch_ip(const IpParams& p, int sm_count, cudaStream_t s) {
  int occ = 1;
  cudaError_t e = occupancy_ip<D>(&occ);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  // persistent grid: one wave of resident CTAs, each owning a contiguous slice of the batch
  // (a slice of >= 4 passes keeps the knot constants staged in shared memory amortised)
#ifdef CIMPC_WITH_IP_V3
  if constexpr (ip2_supported<D>()) {
    if (ip_kernel_version() == 3) {
      constexpr int PPB = ip2_threads<D>() / (D::G / 2);  // subproblems in flight per CTA
      const int64_t need = (p.n + 4 * PPB - 1) / (4 * PPB), cap = (int64_t)sm_count * occ;
      int grid = (int)(need < cap ? need : cap);
      if (grid < 1) grid = 1;
      ip_solve2_kernel<D, ip2_threads<D>(), ip2_minctas<D>()>
          <<<grid, ip2_threads<D>(), KernelSmem2<D, ip2_threads<D>()>::BYTES, s>>>(p);
      return cudaGetLastError();
    }
  }
#endif
  constexpr int PPW = 32 / D::G;
  constexpr int PPB = PPW * (ip_threads<D>() / 32);  // subproblems in flight per CTA
  int64_t need = (p.n + 4 * PPB - 1) / (4 * PPB);
  int64_t cap = (int64_t)sm_count * occ;
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  ip_solve_kernel<D, ip_threads<D>()><<<grid, ip_threads<D>(), KernelSmem<D, ip_threads<D>()>::BYTES, s>>>(p);
  return cudaGetLastError();
}

#define CIMPC_DECLARE_ENTRY(name, nq, nu, nw, nc, nb) \
  const ModelEntry* entries_##name(int* count);
CIMPC_FOR_EACH_MODEL(CIMPC_DECLARE_ENTRY)
#undef CIMPC_DECLARE_ENTRY

#define CIMPC_DEFINE_ENTRIES(name_, nq, nu, nw, nc, nb)                                           \
  const ModelEntry* entries_##name_(int* count) {                                                 \
    using D0 = Dims<nq, nu, nw, nc, nb, 0>;                                                       \
    using D1 = Dims<nq, nu, nw, nc, nb, 1>;                                                       \
    using GEN = gen_##name_::Gen;                                                                 \
    static_assert(GEN::NQ == nq && GEN::NU == nu && GEN::NW == nw && GEN::NC == nc && GEN::NB == nb, "generated model"); \
    static const ModelEntry e[2] = {                                                              \
        {#name_, {nq, nu, nw, nc, nb, 0}, layout_of<D0>(), &launch_ip<D0>, &occupancy_ip<D0>,     \
         &launch_newton_reset<D0>, &launch_newton_step<D0>, &newton_scratch_doubles<D0>,          \
         {&launch_newton_step_general<D0, false>, &launch_newton_step_general<D0, true>},         \
         {&newton_scratch_general_doubles<D0, false>, &newton_scratch_general_doubles<D0, true>}, \
         {&launch_newton_step_dense<D0, false>, &launch_newton_step_dense<D0, true>},             \
         {&newton_scratch_dense_doubles<D0, false>, &newton_scratch_dense_doubles<D0, true>},     \
         &launch_sim_step<GEN>, &sim_scratch_doubles<GEN>, &launch_linearize<GEN>},               \
        {#name_, {nq, nu, nw, nc, nb, 1}, layout_of<D1>(), &launch_ip<D1>, &occupancy_ip<D1>,     \
         &launch_newton_reset<D1>, nullptr, nullptr,                                              \
         {&launch_newton_step_general<D1, false>, &launch_newton_step_general<D1, true>},         \
         {&newton_scratch_general_doubles<D1, false>, &newton_scratch_general_doubles<D1, true>}, \
         {nullptr, nullptr}, {nullptr, nullptr},                                                  \
         &launch_sim_step<GEN>, &sim_scratch_doubles<GEN>, &launch_linearize<GEN>}};              \
    *count = 2;                                                                                   \
    return e;                                                                                     \
  }

// A variant of a robot (payload, terrain): the solver kernels of the base robot (same sizes: the SAME instantiations, not
// compiled a second time) with the variant's own generated residual behind the simulator step and the linearization.
#define CIMPC_DEFINE_VARIANT_ENTRIES(name_, base_, nq, nu, nw, nc, nb)                            \
  const ModelEntry* entries_##name_(int* count) {                                                 \
    using GEN = gen_##name_::Gen;                                                                 \
    static_assert(GEN::NQ == nq && GEN::NU == nu && GEN::NW == nw && GEN::NC == nc && GEN::NB == nb, "generated model"); \
    static ModelEntry e[2];                                                                       \
    static std::once_flag once;                                                                   \
    std::call_once(once, [] {                                                                     \
      int c = 0;                                                                                  \
      const ModelEntry* b = entries_##base_(&c);                                                  \
      for (int i = 0; i < 2; ++i) {                                                               \
        e[i] = b[i];                                                                              \
        e[i].name = #name_;                                                                       \
        e[i].sim_step = &launch_sim_step<GEN>;                                                    \
        e[i].sim_scratch = &sim_scratch_doubles<GEN>;                                             \
        e[i].linearize = &launch_linearize<GEN>;                                                  \
      }                                                                                           \
    });                                                                                           \
    *count = 2;                                                                                   \
    return e;                                                                                     \
  }

}  // namespace cimpc
