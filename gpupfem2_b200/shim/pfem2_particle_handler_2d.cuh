// pfem2_particle_handler_2d.cuh -- drop-in `ParticleHandler2D` for gpuPfem2, backed by libpfem2_b200.so.
//
// This header stands in for the reference's src/particles/particle_handler_2d.cuh: it declares a class with
// the same name and the same public members (reference particle_handler_2d.cuh:12-30), so that
// cases/Cylinder2D/main.cu, cases/PoiseuilleFlow2D/main.cu and src/data_export.cu compile UNCHANGED against
// it.  Everything behind the public interface is different: the object only owns a handle of the C ABI in
// include/pfem2_b200.h; particle storage, kernels and scratch live inside the library.
//
// Integration: either overwrite src/particles/particle_handler_2d.{cuh,cu} in a checkout with
// these two files, or -- without touching the tree -- pre-include this header in every translation unit
// (`nvcc -include pfem2_particle_handler_2d.cuh`, what the shim Makefile does), compile
// pfem2_particle_handler_2d.cu instead of the reference's particle_handler_2d.cu, and link -lpfem2_b200.
//
// It is compiled against the reference's own headers (Mesh2D, deviceVector, Particle2D), which it uses as
// boundary types only.
// Same include guard as the reference header on purpose: once this file has been seen (the shim build pre-includes
// it with `nvcc -include`), a later `#include "particles/particle_handler_2d.cuh"` from inside the reference tree --
// which the quoted-include rule would resolve to the reference's own file -- becomes a no-op.
#if defined(__CUDACC__) && !defined(PARTICLE_HANDLER_2D_CUH) // plain C++ units of the reference never see the class
#define PARTICLE_HANDLER_2D_CUH

#include "particles/particle_2d.cuh"     // reference: the 96-byte AoS record DataExport reads (data_export.cu:97-104)
#include "common/device_vector.cuh"      // reference: deviceVector<T> { T *data; int size; int capacity; }
#include "mesh_2d.cuh"                   // reference: Mesh2D getters = the path's read-only inputs

struct pfem2_handle;

class ParticleHandler2D
{
public:
    // mesh_ is borrowed for the lifetime of the handler (as in the reference); cellDivisionLevel is clamped to [1, 4].
    ParticleHandler2D(const Mesh2D *mesh_, int cellDivisionLevel);
    ~ParticleHandler2D();
    ParticleHandler2D(const ParticleHandler2D &) = delete;
    ParticleHandler2D &operator=(const ParticleHandler2D &) = delete;

    void seedParticles();                                                           // prints "Created %d particles"
    void initParticleVelocity(const deviceVector<double*> &velocitySolution);

    // S x (advect + locate + delete), distribution check + re-seed; prints "Particle handler contains %d particles"
    void advectParticles(const deviceVector<double*> &velocitySolution, double timeStep, int particleSubsteps);

    void correctParticleVelocity(const deviceVector<double*> &velocitySolution, const deviceVector<double*> &velocitySolutionOld);

    // writes the projected nodal velocity through the two pointers held in `velocity`
    void projectVelocityOntoGrid(deviceVector<double*> &velocity);
    // extension: also writes the projected field into `velocityCopy` (replaces the copy_d2d pair that follows the call in the cases)
    void projectVelocityOntoGrid(deviceVector<double*> &velocity, deviceVector<double*> &velocityCopy);

    // device pointer to 96-byte Particle2D records, materialised on demand from the library's SoA storage;
    // valid until the next mutating call
    const Particle2D *getParticles() const;
    int getParticleCount() const;

private:
    void check(int rc, const char *what) const; // reference error behaviour: message on stderr + exit(EXIT_FAILURE)

    // advectParticles' stdout line needs the new count, i.e. a wait for the whole advect.  The line is printed by the NEXT call
    // into the handler instead (in the cases: projectVelocityOntoGrid, two statements later, with nothing printed in between:
    // cases/Cylinder2D/main.cu:797-803), when the count has long arrived through the asynchronous read-back -- same text, same
    // position in the output, no host stall inside advectParticles.
    void flushCountLine() const;
    mutable bool countLinePending = false;

    const Mesh2D *mesh;
    pfem2_handle *handle;
};

#endif // PARTICLE_HANDLER_2D_CUH
