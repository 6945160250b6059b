// Kernel instances for a planar robot on the piecewise terrain (src/simulation/environments/piecewise.jl): same sizes as
// the robot on flat ground (the solver kernels are the base robot's, taken from its entries), its own
// generated residual with the terrain atoms.
#include "gen/residual_quadruped_piecewise.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_VARIANT_ENTRIES(quadruped_piecewise, quadruped, 11, 8, 2, 4, 8)
}
