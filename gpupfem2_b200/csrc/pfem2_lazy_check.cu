// Compile check of the DRAFT lazy re-sort kernels (pfem2_lazy.cuh): `make lazy-check`.  Not linked into the library.
#include "pfem2_lazy.cuh"

namespace pfem2 {
template __global__ void k_project_cells_lazy<2>(int, int, ParticleSoA, const unsigned *, const int *, double *);
template __global__ void k_project_cells_lazy<4>(int, int, ParticleSoA, const unsigned *, const int *, double *);
template __global__ void k_project_cells_lazy<8>(int, int, ParticleSoA, const unsigned *, const int *, double *);
template __global__ void k_project_cells_lazy<16>(int, int, ParticleSoA, const unsigned *, const int *, double *);
void lazy_check_refs(void **out)
{
    out[0] = (void *)k_iota;
    out[1] = (void *)k_rank;
    out[2] = (void *)k_reseed_lazy;
    out[3] = (void *)k_materialize;
    out[4] = (void *)k_advect_locate_lazy<0, true, false, 3>; // the default instantiation of the move pass (level <= 4, S = 3)
    out[5] = (void *)k_advect_locate_lazy<0, true, true, 0>;
}
} // namespace pfem2
