// Kernel instances (both Newton modes) for the payload variant of a BASELINE robot: same sizes as the nominal robot
// (the solver kernels are the base robot's, taken from its entries), its own generated residual.
#include "gen/residual_quadruped_payload.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_VARIANT_ENTRIES(quadruped_payload, quadruped, 11, 8, 2, 4, 8)
}
