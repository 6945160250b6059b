// Runtime view of one compiled (robot, mode) kernel instance.
#pragma once
#include <cuda_runtime.h>

#include "ip_kernel.cuh"

namespace cimpc {

struct LinLayout {  // runtime copy of Dims<...>::O_* (doubles)
  int nx, ny, nz, nth, ncol, nd, group;
  int o_dx, o_dy1, o_rx, o_ry1, o_ry2, o_rtd, o_rtr, o_cd, o_cr, o_ai, o_cai, o_aib, o_s0, o_w, o_ar;
  int stride;
};

struct ModelEntry {
  const char* name;
  cimpc_model_desc desc;
  LinLayout lay;
  cudaError_t (*launch)(const IpParams& p, int sm_count, cudaStream_t s);
  cudaError_t (*occupancy)(int* blocks_per_sm);
};

constexpr int IP_THREADS = 128;

template <class D>
LinLayout layout_of() {
  LinLayout l;
  l.nx = D::NX; l.ny = D::NY; l.nz = D::NZ; l.nth = D::NTH; l.ncol = D::NCOL; l.nd = D::ND; l.group = D::G;
  l.o_dx = D::O_DX; l.o_dy1 = D::O_DY1; l.o_rx = D::O_RX; l.o_ry1 = D::O_RY1; l.o_ry2 = D::O_RY2;
  l.o_rtd = D::O_RTD; l.o_rtr = D::O_RTR; l.o_cd = D::O_CD; l.o_cr = D::O_CR; l.o_ai = D::O_AI;
  l.o_cai = D::O_CAI; l.o_aib = D::O_AIB; l.o_s0 = D::O_S0; l.o_w = D::O_W; l.o_ar = D::O_AR;
  l.stride = D::LIN_STRIDE;
  return l;
}

template <class D>
cudaError_t occupancy_ip(int* blocks_per_sm) {
  return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, ip_solve_kernel<D, IP_THREADS>,
                                                       IP_THREADS, 0);
}

template <class D>
cudaError_t launch_ip(const IpParams& p, int sm_count, cudaStream_t s) {
  constexpr int PPW = 32 / D::G;
  constexpr int PPB = PPW * (IP_THREADS / 32);  // subproblems per CTA pass
  int occ = 1;
  cudaError_t e = occupancy_ip<D>(&occ);
  if (e != cudaSuccess) return e;
  if (occ < 1) occ = 1;
  int64_t need = (p.n + PPB - 1) / PPB;
  int64_t cap = (int64_t)sm_count * occ;  // persistent: one wave of resident CTAs
  int grid = (int)(need < cap ? need : cap);
  if (grid < 1) grid = 1;
  ip_solve_kernel<D, IP_THREADS><<<grid, IP_THREADS, 0, s>>>(p);
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
    static const ModelEntry e[2] = {                                                              \
        {#name_, {nq, nu, nw, nc, nb, 0}, layout_of<D0>(), &launch_ip<D0>, &occupancy_ip<D0>},    \
        {#name_, {nq, nu, nw, nc, nb, 1}, layout_of<D1>(), &launch_ip<D1>, &occupancy_ip<D1>}};   \
    *count = 2;                                                                                   \
    return e;                                                                                     \
  }

}  // namespace cimpc
