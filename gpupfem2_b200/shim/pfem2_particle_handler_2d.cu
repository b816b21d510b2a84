// pfem2_particle_handler_2d.cu -- member functions of the drop-in ParticleHandler2D: thin forwards to the C ABI.
// Replaces the host methods of reference src/particles/particle_handler_2d.cu:238-423 (no kernels here).
#include "pfem2_particle_handler_2d.cuh"

#include "pfem2_b200.h"

#include <cstdio>
#include <cstdlib>

void ParticleHandler2D::check(int rc, const char *what) const
{
    if (rc == PFEM2_OK) return;
    // the reference prints and exits on any CUDA failure (src/common/cuda_helper.cuh:57-70)
    fprintf(stderr, "pfem2_b200: %s failed (code %d): %s\n", what, rc, pfem2_last_error(handle));
    exit(EXIT_FAILURE);
}

ParticleHandler2D::ParticleHandler2D(const Mesh2D *mesh_, int cellDivisionLevel) : mesh(mesh_), handle(nullptr)
{
    // the mesh kernels (initMesh) run on the null stream like everything else in the reference; the library also
    // uses the null stream by default, so plain stream order makes the inverse Jacobians visible here
    pfem2_mesh_view view;
    view.n_nodes = mesh->getVertices().size;
    view.n_cells = mesh->getCells().size;
    view.d_vertices = reinterpret_cast<const double *>(mesh->getVertices().data);
    view.d_cells = reinterpret_cast<const unsigned *>(mesh->getCells().data);
    view.d_inv_jacobi = reinterpret_cast<const double *>(mesh->getInvJacobi().data);
    view.d_nbr_offsets = mesh->getCellNeighborsOffsets().data;
    view.d_nbr_indices = mesh->getCellNeighborIndices().data;
    pfem2_options opt;
    pfem2_default_options(&opt);
    opt.capacity_factor = 1.1; // CONSTANTS::MEMORY_REALLOCATION_COEFFICIENT; the library grows on demand like resize()
    if (const char *e = getenv("PFEM2_LAZY_SORT")) opt.lazy_sort = atoi(e) != 0; // in-situ A/B runs: 0 = physical re-sort in every advect
    if (const char *e = getenv("PFEM2_GRAPH_ADVECT")) opt.graph_advect = atoi(e); // in-situ A/B runs: -1 = plain launches
    const int rc = pfem2_create(&handle, &view, cellDivisionLevel, &opt);
    if (rc != PFEM2_OK) {
        fprintf(stderr, "pfem2_b200: pfem2_create failed (code %d): %s\n", rc, pfem2_last_error(nullptr));
        exit(EXIT_FAILURE);
    }
}

ParticleHandler2D::~ParticleHandler2D()
{
    flushCountLine();
    pfem2_destroy(handle);
}

void ParticleHandler2D::flushCountLine() const
{
    if (!countLinePending) return;
    countLinePending = false;
    int n = 0;
    check(pfem2_particle_count(handle, &n), "getParticleCount");
    printf("Particle handler contains %d particles\n", n); // particle_handler_2d.cu:341
}

void ParticleHandler2D::seedParticles()
{
    flushCountLine();
    check(pfem2_seed(handle), "seedParticles");
    printf("Created %d particles\n", getParticleCount());
}

void ParticleHandler2D::initParticleVelocity(const deviceVector<double*> &velocitySolution)
{
    flushCountLine();
    check(pfem2_init_velocity_ptrs(handle, velocitySolution.data), "initParticleVelocity");
}

void ParticleHandler2D::advectParticles(const deviceVector<double*> &velocitySolution, double timeStep, int particleSubsteps)
{
    flushCountLine();
    check(pfem2_advect_ptrs(handle, velocitySolution.data, timeStep, particleSubsteps), "advectParticles");
    countLinePending = true; // "Particle handler contains %d particles": printed by the next call into the handler
}

void ParticleHandler2D::correctParticleVelocity(const deviceVector<double*> &velocitySolution,
                                                const deviceVector<double*> &velocitySolutionOld)
{
    flushCountLine();
    check(pfem2_correct_ptrs(handle, velocitySolution.data, velocitySolutionOld.data), "correctParticleVelocity");
}

void ParticleHandler2D::projectVelocityOntoGrid(deviceVector<double*> &velocity)
{
    check(pfem2_project_ptrs(handle, velocity.data), "projectVelocityOntoGrid");
    flushCountLine(); // (after the enqueue: the projection is already queued behind the advect when the host waits for the count)
}

void ParticleHandler2D::projectVelocityOntoGrid(deviceVector<double*> &velocity, deviceVector<double*> &velocityCopy)
{
    check(pfem2_project_dual_ptrs(handle, velocity.data, velocityCopy.data), "projectVelocityOntoGrid");
    flushCountLine();
}

const Particle2D *ParticleHandler2D::getParticles() const
{
    flushCountLine();
    const void *aos = nullptr;
    check(pfem2_export_aos(handle, &aos, nullptr), "getParticles");
    return static_cast<const Particle2D *>(aos);
}

int ParticleHandler2D::getParticleCount() const
{
    flushCountLine();
    int n = 0;
    check(pfem2_particle_count(handle, &n), "getParticleCount");
    return n;
}
