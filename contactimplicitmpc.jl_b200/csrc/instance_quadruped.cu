// Kernel instances (both Newton modes) for one robot of BASELINE.json's configs.
#include "gen/residual_quadruped.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_ENTRIES(quadruped, 11, 8, 2, 4, 8)
}
