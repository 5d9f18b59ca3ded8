// placeholder until the collocation kernel lands (keeps the dispatch table complete)
#pragma once
#include "common.cuh"
#define RB_COLLOC_CPB 4
#define RB_COLLOC_NCR(nz, nu) (8 * ((nz) + (nu) + 2) + (nz) + (nu) + 2)
template <class PF>
__global__ void colloc_cells_kernel(const RbDev d, const RbBatch b) {}
