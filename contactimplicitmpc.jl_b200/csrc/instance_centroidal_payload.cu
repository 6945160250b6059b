// Kernel instances (both Newton modes) for the payload variant of a BASELINE robot: same sizes as the nominal robot
// (the solver kernels are the base robot's, taken from its entries), its own generated residual.
#include "gen/residual_centroidal_payload.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_VARIANT_ENTRIES(centroidal_payload, centroidal, 18, 12, 3, 4, 16)
}
