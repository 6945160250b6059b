// Kernel instances (both Newton modes) for one robot of BASELINE.json's configs.
#include "gen/residual_hopper2d.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_ENTRIES(hopper2d, 4, 2, 2, 1, 2)
}
