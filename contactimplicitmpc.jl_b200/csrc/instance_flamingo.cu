// Kernel instances (both Newton modes) for one robot of BASELINE.json's configs.
#include "gen/residual_flamingo.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_ENTRIES(flamingo, 9, 6, 2, 4, 8)
}
