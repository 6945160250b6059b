// Kernel instances (both Newton modes) for one robot of BASELINE.json's configs.
#include "gen/residual_centroidal.h"
#include "registry.cuh"
namespace cimpc {
CIMPC_DEFINE_ENTRIES(centroidal, 18, 12, 3, 4, 16)
}
