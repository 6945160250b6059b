// Device-side linearization of the contact dynamics at the reference knots.
//
// Replaces `LinearizedStep(s, z, θ, κ)` / `update!(lin, s, z, θ)` (src/controller/linearized_step.jl:10-29, 48-55):
// the code-generated r!, rz!, rθ! of the robot (`gen/residual_<robot>.h`, the product's counterpart of
// src/simulation/code_gen_simulation.jl:114-197) are evaluated at (z0[t], θ0[t], κ) for every knot t and written
// as the dense column-major r0 (nz), rz0 (nz × nz), rθ0 (nz × nθ) that `prep_kernel` slices into the RLin / RZLin /
// RθLin / Schur constants — the same three arrays `cimpc_upload_linearization` takes from the host.
//
// Mapping = the simulator kernel's generated-code phase: one CTA of GEN::NS warps owns a tile of 32 knots, lane j of
// every warp works on knot j, warp s evaluates slice s of the outputs; the sin / cos of the trig atoms are computed
// once per knot by ONE sincos call site and shared through shared memory.  Set-up work (60-70 knots per gait,
// re-run when the caller re-linearizes), not a throughput kernel.
#pragma once
#include <cstdint>

namespace cimpc {

struct LinEvalParams {
  int H;
  const double* z0;   // nz × H
  const double* th0;  // nθ × H
  double kappa;
  double* r0;    // nz × H
  double* rz0;   // nz × nz × H   (zero-filled by the caller; only structural non-zeros are written)
  double* rth0;  // nz × nθ × H   (idem)
};

template <class GEN>
__global__ void __launch_bounds__(GEN::NS * 32) linearize_kernel(const LinEvalParams p) {
  constexpr int NZ = GEN::NZ, NTH = GEN::NTH, NTRIG = GEN::NTRIG, WARPS = GEN::NS;
  extern __shared__ double trig[];  // [2·NTRIG][32]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + lane;
  const bool valid = t < p.H;
  const int tc = valid ? t : p.H - 1;
  const double* zt = p.z0 + (size_t)tc * NZ;
  const double* tt = p.th0 + (size_t)tc * NTH;
  auto z = [&](int i) { return zt[i]; };
  auto th = [&](int i) { return tt[i]; };
  auto tr = [&](int i) { return trig[i * 32 + lane]; };
  for (int k = wid; k < NTRIG; k += WARPS) {
    if (GEN::is_terr(k)) continue;
    double sn, cs;
    sincos(GEN::trig_arg(k, z, th), &sn, &cs);
    trig[(2 * k) * 32 + lane] = sn;
    trig[(2 * k + 1) * 32 + lane] = cs;
  }
  __syncthreads();
  if constexpr (GEN::NTERR > 0) {  // terrain atoms: surface height, slope and rotation under every contact point
    for (int i = wid; i < GEN::NTERR; i += WARPS) {
      const double x = GEN::terr_x(i, z, th, tr);
      double f, g, c, sn;
      GEN::terr_atoms(x, f, g, c, sn);
      trig[(2 * (GEN::TERR0 + 2 * i)) * 32 + lane] = f;
      trig[(2 * (GEN::TERR0 + 2 * i) + 1) * 32 + lane] = g;
      trig[(2 * (GEN::TERR0 + 2 * i + 1)) * 32 + lane] = c;
      trig[(2 * (GEN::TERR0 + 2 * i + 1) + 1) * 32 + lane] = sn;
    }
    __syncthreads();
  }
  if (!valid) return;
  double* r0 = p.r0 + (size_t)t * NZ;
  double* rz = p.rz0 + (size_t)t * NZ * NZ;
  double* rt = p.rth0 + (size_t)t * NZ * NTH;
  GEN::r_slice(wid, z, th, tr, p.kappa, [&](int i, double v) { r0[i] = v; });
  GEN::rz_slice(wid, z, th, tr, [&](int k, double v) { rz[GEN::row(k) + (size_t)GEN::col(k) * NZ] = v; });
  GEN::rth_slice(wid, z, th, tr, [&](int k, double v) { rt[GEN::trow(k) + (size_t)GEN::tcol(k) * NZ] = v; });
}

}  // namespace cimpc
