// pfem2_common.cuh -- what every kernel header of the particle step shares (sm_100a).
//
// Data layout in HBM: particles are one array of 64-byte records (four 16-byte fields, accessed
// through strided field views); cell c owns the segment [cell_start[c], cell_start[c+1]) of the SORTED ORDER, which is
// either the physical order of the array or -- the default, lazy re-sort -- a permutation src[] over a dense array.
// Mesh data the path reads is repacked once into one 64-byte CellGeom record per cell.  All particle counts live in
// device memory (Counters); kernels are grid-stride and read the live count themselves, so a step issues no
// device->host copy.
#pragma once

#include "pfem2_device.cuh"
#include "pfem2_sort.cuh"
#include "pfem2_tma.cuh"

#include <cuda.h> // CUtensorMap (type only; the encoder is resolved at run time through cudaGetDriverEntryPoint)

namespace pfem2 {

constexpr int kThreads = 256;

// Counters.overflow bits (1: particle capacity)
constexpr int kOverflowMigration = 4; // Counters.overflow bit: migration buffer too small / non-adjacent destination
constexpr int kOverflowP2PTimeout = 8; // Counters.overflow bit: a neighbour's delivery did not arrive in time

} // namespace pfem2
