/*
 * ref_harness.cu -- drives the UNMODIFIED reference ParticleHandler2D (gpuPfem2) on a B200.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY.  Compiled here (where /root/reference exists) against the
 * reference's own headers and linked with oracle/_ref/libgpuPfem2Lib.so, which is built from the
 * reference sources where they lie (oracle/Makefile).  The binary lands in oracle/_ref/ (git-ignored,
 * travels to the GPU box).  Nothing of the reference is copied into this repository.
 *
 * It does two things:
 *   ref_harness dump  <case.bin> <out_prefix>     run a case file and dump particle + nodal state
 *   ref_harness timecase <case.bin> <steps> <warmup>   time the particle step on a case file (shipped meshes)
 *   ref_harness time  <nx> <ny> <lx> <ly> <level> <substeps> <dt> <umax> <steps> <warmup> [colmajor]
 *                                                  time the particle step on a synthetic channel
 *   ref_harness digest <nx> <ny> <lx> <ly> <level> <substeps> <dt> <umax> <steps> <out_prefix>
 *                                                  full-size parity: per step one JSON line with the particle count and
 *                                                  order-independent checksums of the reference's particle state (wrapping
 *                                                  int64 sums of the bit patterns of x, y, L0, L1, L2 and of the cell ids;
 *                                                  plain double sums of the velocities); the projected nodal field of the
 *                                                  last step goes to <out_prefix>_w.bin
 *
 * Meshes are injected into Mesh2D through its const getters (the members are deviceVectors with
 * public allocate()), bypassing loadMeshFromFile and its O(C^2) neighbour fill (mesh_2d.cu:107-139);
 * initMesh() (mesh_2d.cu:98-105) then computes areas and inverse Jacobians with the reference kernels.
 *
 * Step protocol ("isolated mode", the FEM stage replaced by a frozen nodal field F):
 *   advectParticles(F, dt, S) ; projectVelocityOntoGrid(W) ; correctParticleVelocity(F, W)
 */
#include "particles/particle_handler_2d.cuh"
#include "mesh_2d.cuh"
#include "common/cuda_memory.cuh"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <unistd.h>

template <class T> static deviceVector<T> &mut(const deviceVector<T> &v) { return const_cast<deviceVector<T> &>(v); }
template <class T> static std::vector<T> &mut(const std::vector<T> &v) { return const_cast<std::vector<T> &>(v); }

static void inject_mesh(Mesh2D &mesh, const std::vector<Point2> &verts, const std::vector<uint3> &cells,
                        const std::vector<int> &off, const std::vector<int> &idx)
{
    mut(mesh.getVertices()).allocate((int)verts.size());
    mut(mesh.getCells()).allocate((int)cells.size());
    mut(mesh.getEdgeBoundaryIDs()).allocate((int)cells.size());
    copy_h2d(verts.data(), mesh.getVertices().data, verts.size());
    copy_h2d(cells.data(), mesh.getCells().data, cells.size());
    set_value_device(mesh.getEdgeBoundaryIDs().data, -1, cells.size());
    mut(mesh.getCellNeighborsOffsets()).allocate((int)off.size());
    mut(mesh.getCellNeighborIndices()).allocate((int)idx.size());
    copy_h2d(off.data(), mesh.getCellNeighborsOffsets().data, off.size());
    copy_h2d(idx.data(), mesh.getCellNeighborIndices().data, idx.size());
    mut(mesh.getHostVertices()) = verts;
    mut(mesh.getHostCells()) = cells;
    mesh.initMesh();
    checkCudaErrors(cudaDeviceSynchronize());
}

// O(C) vertex-sharing one-ring, ascending (same definition as mesh_2d.cu:107-139)
static void one_ring(int n_nodes, const std::vector<uint3> &cells, std::vector<int> &off, std::vector<int> &idx)
{
    const int C = (int)cells.size();
    std::vector<int> voff(n_nodes + 1, 0);
    auto node = [&](int c, int k) { return k == 0 ? cells[c].x : (k == 1 ? cells[c].y : cells[c].z); };
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < 3; ++k) ++voff[node(c, k) + 1];
    for (int n = 0; n < n_nodes; ++n) voff[n + 1] += voff[n];
    std::vector<int> vcell(voff[n_nodes]), cur(voff.begin(), voff.end() - 1);
    for (int c = 0; c < C; ++c)
        for (int k = 0; k < 3; ++k) vcell[cur[node(c, k)]++] = c;
    off.assign(C + 1, 0);
    idx.clear();
    idx.reserve((size_t)C * 13);
    std::vector<int> buf;
    for (int c = 0; c < C; ++c) {
        buf.clear();
        for (int k = 0; k < 3; ++k)
            for (int p = voff[node(c, k)]; p < voff[node(c, k) + 1]; ++p)
                if (vcell[p] != c) buf.push_back(vcell[p]);
        std::sort(buf.begin(), buf.end());
        buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
        idx.insert(idx.end(), buf.begin(), buf.end());
        off[c + 1] = (int)idx.size();
    }
}

struct NodalField {
    deviceVector<double> comp[2];
    deviceVector<double *> ptrs;
    void init(int n, const double *hx, const double *hy)
    {
        comp[0].allocate(n);
        comp[1].allocate(n);
        if (hx) copy_h2d(hx, comp[0].data, n); else comp[0].clearValues();
        if (hy) copy_h2d(hy, comp[1].data, n); else comp[1].clearValues();
        ptrs.allocate(2);
        double *h[2] = {comp[0].data, comp[1].data};
        copy_h2d(h, ptrs.data, 2);
        checkCudaErrors(cudaDeviceSynchronize());
    }
};

template <class T> static void rd(FILE *f, T *p, size_t n)
{
    if (fread(p, sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(2); }
}
template <class T> static void wr(FILE *f, const T *p, size_t n)
{
    if (fwrite(p, sizeof(T), n, f) != n) { fprintf(stderr, "short write\n"); exit(2); }
}

static void dump_state(const std::string &path, const ParticleHandler2D &ph, const NodalField &W, int n_nodes)
{
    const int n = ph.getParticleCount();
    std::vector<Particle2D> hp(n);
    copy_d2h(ph.getParticles(), hp.data(), n);
    std::vector<double> wx(n_nodes), wy(n_nodes);
    copy_d2h(W.comp[0].data, wx.data(), n_nodes);
    copy_d2h(W.comp[1].data, wy.data(), n_nodes);
    checkCudaErrors(cudaDeviceSynchronize());
    std::vector<double> col(n);
    std::vector<unsigned> ucol(n);
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
    int64_t hdr[2] = {n, n_nodes};
    wr(f, hdr, 2);
    auto put = [&](auto get) { for (int i = 0; i < n; ++i) col[i] = get(hp[i]); wr(f, col.data(), n); };
    put([](const Particle2D &p) { return p.getPosition().x; });
    put([](const Particle2D &p) { return p.getPosition().y; });
    put([](const Particle2D &p) { return p.getLocalPosition().x; });
    put([](const Particle2D &p) { return p.getLocalPosition().y; });
    put([](const Particle2D &p) { return p.getLocalPosition().z; });
    put([](const Particle2D &p) { return p.getVelocity().x; });
    put([](const Particle2D &p) { return p.getVelocity().y; });
    for (int i = 0; i < n; ++i) ucol[i] = hp[i].getCellID();
    wr(f, ucol.data(), n);
    for (int i = 0; i < n; ++i) ucol[i] = hp[i].getID();
    wr(f, ucol.data(), n);
    wr(f, wx.data(), n_nodes);
    wr(f, wy.data(), n_nodes);
    fclose(f);
}

/* case file (little endian), written by gpupfem2_b200/casefile.py:
 *   int64 magic(0x50464d32), N, C, nnz, level, substeps, nsteps, ndump ; double dt ; int64 dump_steps[ndump]
 *   double vertices[2N] ; uint32 cells[3C] ; int32 nbr_off[C+1] ; int32 nbr_idx[nnz] ; double Fx[N] ; double Fy[N]
 * dump step 0 = state after seedParticles + initParticleVelocity(F), step k = after k full steps. */
static int run_dump(const char *case_path, const char *out_prefix)
{
    FILE *f = fopen(case_path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", case_path); return 2; }
    int64_t hdr[8];
    rd(f, hdr, 8);
    if (hdr[0] != 0x50464d32) { fprintf(stderr, "bad magic\n"); return 2; }
    const int N = (int)hdr[1], C = (int)hdr[2], nnz = (int)hdr[3], level = (int)hdr[4], S = (int)hdr[5];
    const int nsteps = (int)hdr[6], ndump = (int)hdr[7];
    double dt;
    rd(f, &dt, 1);
    std::vector<int64_t> dumps(ndump);
    rd(f, dumps.data(), ndump);
    std::vector<Point2> verts(N);
    std::vector<uint3> cells(C);
    std::vector<int> off(C + 1), idx(nnz);
    std::vector<double> fx(N), fy(N);
    rd(f, (double *)verts.data(), 2 * (size_t)N);
    rd(f, (unsigned *)cells.data(), 3 * (size_t)C);
    rd(f, off.data(), C + 1);
    rd(f, idx.data(), nnz);
    rd(f, fx.data(), N);
    rd(f, fy.data(), N);
    fclose(f);

    Mesh2D mesh;
    inject_mesh(mesh, verts, cells, off, idx);
    {   // the inverse Jacobians the reference computed, so every implementation can share their bits
        std::vector<double> invj(4 * (size_t)C);
        copy_d2h((const double *)mesh.getInvJacobi().data, invj.data(), invj.size());
        checkCudaErrors(cudaDeviceSynchronize());
        FILE *g = fopen((std::string(out_prefix) + "_invj.bin").c_str(), "wb");
        wr(g, invj.data(), invj.size());
        fclose(g);
    }
    NodalField F, W;
    F.init(N, fx.data(), fy.data());
    W.init(N, nullptr, nullptr);

    ParticleHandler2D ph(&mesh, level);
    ph.seedParticles();
    ph.initParticleVelocity(F.ptrs);
    checkCudaErrors(cudaDeviceSynchronize());

    auto wants = [&](int s) { return std::find(dumps.begin(), dumps.end(), (int64_t)s) != dumps.end(); };
    char name[64];
    if (wants(0)) { snprintf(name, sizeof name, "_step%05d.bin", 0); dump_state(out_prefix + std::string(name), ph, W, N); }
    FILE *cf = fopen((std::string(out_prefix) + "_counts.txt").c_str(), "w");
    for (int s = 1; s <= nsteps; ++s) {
        ph.advectParticles(F.ptrs, dt, S);
        ph.projectVelocityOntoGrid(W.ptrs);
        ph.correctParticleVelocity(F.ptrs, W.ptrs);
        checkCudaErrors(cudaDeviceSynchronize());
        fprintf(cf, "%d %d\n", s, ph.getParticleCount());
        if (wants(s)) { snprintf(name, sizeof name, "_step%05d.bin", s); dump_state(out_prefix + std::string(name), ph, W, N); }
    }
    fclose(cf);
    return 0;
}

/* time a case file (small shipped meshes): same protocol as run_dump, no dumps, CUDA-event + wall timing */
static int run_timecase(const char *case_path, int steps, int warmup)
{
    FILE *f = fopen(case_path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", case_path); return 2; }
    int64_t hdr[8];
    rd(f, hdr, 8);
    if (hdr[0] != 0x50464d32) { fprintf(stderr, "bad magic\n"); return 2; }
    const int N = (int)hdr[1], C = (int)hdr[2], nnz = (int)hdr[3], level = (int)hdr[4], S = (int)hdr[5], ndump = (int)hdr[7];
    double dt;
    rd(f, &dt, 1);
    std::vector<int64_t> dumps(ndump);
    rd(f, dumps.data(), ndump);
    std::vector<Point2> verts(N);
    std::vector<uint3> cells(C);
    std::vector<int> off(C + 1), idx(nnz);
    std::vector<double> fx(N), fy(N);
    rd(f, (double *)verts.data(), 2 * (size_t)N);
    rd(f, (unsigned *)cells.data(), 3 * (size_t)C);
    rd(f, off.data(), C + 1);
    rd(f, idx.data(), nnz);
    rd(f, fx.data(), N);
    rd(f, fy.data(), N);
    fclose(f);
    Mesh2D mesh;
    inject_mesh(mesh, verts, cells, off, idx);
    NodalField F, W;
    F.init(N, fx.data(), fy.data());
    W.init(N, nullptr, nullptr);
    fflush(stdout);
    FILE *real_out = fdopen(dup(fileno(stdout)), "w");
    if (!freopen("/dev/null", "w", stdout)) return 2;
    ParticleHandler2D ph(&mesh, level);
    ph.seedParticles();
    ph.initParticleVelocity(F.ptrs);
    for (int s = 0; s < warmup; ++s) {
        ph.advectParticles(F.ptrs, dt, S);
        ph.projectVelocityOntoGrid(W.ptrs);
        ph.correctParticleVelocity(F.ptrs, W.ptrs);
    }
    checkCudaErrors(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    long long psteps = 0;
    const auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0);
    for (int s = 0; s < steps; ++s) {
        ph.advectParticles(F.ptrs, dt, S);
        psteps += ph.getParticleCount();
        ph.projectVelocityOntoGrid(W.ptrs);
        ph.correctParticleVelocity(F.ptrs, W.ptrs);
    }
    cudaEventRecord(e1);
    checkCudaErrors(cudaDeviceSynchronize());
    const auto t1 = std::chrono::steady_clock::now();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    fprintf(real_out,
            "{\"impl\": \"reference-cuda\", \"cells\": %d, \"nodes\": %d, \"particles\": %d, \"steps\": %d, "
            "\"ms_per_step\": %.6f, \"wall_ms_per_step\": %.6f, \"particle_steps_per_s\": %.6e}\n",
            C, N, ph.getParticleCount(), steps, ms / steps, std::chrono::duration<double, std::milli>(t1 - t0).count() / steps,
            psteps / (ms * 1e-3));
    fflush(real_out);
    return 0;
}

/* synthetic structured channel (same generator as gpupfem2_b200/mesh.py: structured_channel) */
static void channel(int nx, int ny, double lx, double ly, bool colmajor, std::vector<Point2> &verts, std::vector<uint3> &cells)
{
    const double hx = lx / nx, hy = ly / ny;
    verts.resize((size_t)(nx + 1) * (ny + 1));
    cells.resize((size_t)2 * nx * ny);
    auto nid = [&](int i, int j) { return colmajor ? (unsigned)(i * (ny + 1) + j) : (unsigned)(j * (nx + 1) + i); };
    for (int j = 0; j <= ny; ++j)
        for (int i = 0; i <= nx; ++i) verts[nid(i, j)] = {i * hx, j * hy};
    for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i) {
            const size_t q = colmajor ? (size_t)i * ny + j : (size_t)j * nx + i;
            const unsigned a = nid(i, j), b = nid(i + 1, j), c = nid(i + 1, j + 1), d = nid(i, j + 1);
            cells[2 * q] = {a, b, c};
            cells[2 * q + 1] = {a, c, d};
        }
}

static int run_time(int argc, char **argv)
{
    if (argc < 12) { fprintf(stderr, "usage: time nx ny lx ly level substeps dt umax steps warmup [colmajor]\n"); return 2; }
    const int nx = atoi(argv[2]), ny = atoi(argv[3]);
    const double lx = atof(argv[4]), ly = atof(argv[5]);
    const int level = atoi(argv[6]), S = atoi(argv[7]);
    const double dt = atof(argv[8]), umax = atof(argv[9]);
    const int steps = atoi(argv[10]), warmup = atoi(argv[11]);
    const bool colmajor = argc > 12 ? atoi(argv[12]) != 0 : true;

    std::vector<Point2> verts;
    std::vector<uint3> cells;
    channel(nx, ny, lx, ly, colmajor, verts, cells);
    std::vector<int> off, idx;
    one_ring((int)verts.size(), cells, off, idx);
    const int N = (int)verts.size();
    std::vector<double> fx(N), fy(N, 0.0);
    for (int i = 0; i < N; ++i) fx[i] = 4.0 * umax * verts[i].y * (ly - verts[i].y) / (ly * ly);

    Mesh2D mesh;
    inject_mesh(mesh, verts, cells, off, idx);
    NodalField F, W;
    F.init(N, fx.data(), fy.data());
    W.init(N, nullptr, nullptr);
    // silence the reference's per-step printf during timing (stdout -> /dev/null), keep stderr
    fflush(stdout);
    FILE *real_out = fdopen(dup(fileno(stdout)), "w");
    if (!freopen("/dev/null", "w", stdout)) return 2;

    ParticleHandler2D ph(&mesh, level);
    ph.seedParticles();
    ph.initParticleVelocity(F.ptrs);
    checkCudaErrors(cudaDeviceSynchronize());

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    long long psteps = 0;
    for (int s = 0; s < warmup; ++s) {
        ph.advectParticles(F.ptrs, dt, S);
        ph.projectVelocityOntoGrid(W.ptrs);
        ph.correctParticleVelocity(F.ptrs, W.ptrs);
    }
    checkCudaErrors(cudaDeviceSynchronize());
    const auto t0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0);
    for (int s = 0; s < steps; ++s) {
        ph.advectParticles(F.ptrs, dt, S);
        psteps += ph.getParticleCount();
        ph.projectVelocityOntoGrid(W.ptrs);
        ph.correctParticleVelocity(F.ptrs, W.ptrs);
    }
    cudaEventRecord(e1);
    checkCudaErrors(cudaDeviceSynchronize());
    const auto t1 = std::chrono::steady_clock::now();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double wall_ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    fprintf(real_out,
            "{\"impl\": \"reference-cuda\", \"cells\": %d, \"nodes\": %d, \"particles\": %d, \"steps\": %d, "
            "\"ms_per_step\": %.6f, \"wall_ms_per_step\": %.6f, \"particle_steps_per_s\": %.6e}\n",
            (int)cells.size(), N, ph.getParticleCount(), steps, ms / steps, wall_ms / steps, psteps / (ms * 1e-3));
    fflush(real_out);
    return 0;
}

/* full-size parity against the reference itself without a multi-GB dump: order-independent checksums per step */
static int run_digest(int argc, char **argv)
{
    if (argc < 12) { fprintf(stderr, "usage: digest nx ny lx ly level substeps dt umax steps out_prefix\n"); return 2; }
    const int nx = atoi(argv[2]), ny = atoi(argv[3]);
    const double lx = atof(argv[4]), ly = atof(argv[5]);
    const int level = atoi(argv[6]), S = atoi(argv[7]);
    const double dt = atof(argv[8]), umax = atof(argv[9]);
    const int steps = atoi(argv[10]);
    const std::string prefix = argv[11];
    std::vector<Point2> verts;
    std::vector<uint3> cells;
    channel(nx, ny, lx, ly, true, verts, cells);
    std::vector<int> off, idx;
    one_ring((int)verts.size(), cells, off, idx);
    const int N = (int)verts.size();
    std::vector<double> fx(N), fy(N, 0.0);
    for (int i = 0; i < N; ++i) fx[i] = 4.0 * umax * verts[i].y * (ly - verts[i].y) / (ly * ly);
    Mesh2D mesh;
    inject_mesh(mesh, verts, cells, off, idx);
    NodalField F, W;
    F.init(N, fx.data(), fy.data());
    W.init(N, nullptr, nullptr);
    fflush(stdout);
    FILE *real_out = fdopen(dup(fileno(stdout)), "w");
    if (!freopen("/dev/null", "w", stdout)) return 2;
    ParticleHandler2D ph(&mesh, level);
    ph.seedParticles();
    ph.initParticleVelocity(F.ptrs);
    checkCudaErrors(cudaDeviceSynchronize());
    std::vector<Particle2D> hp;
    auto bits = [](double v) { int64_t b; memcpy(&b, &v, 8); return (uint64_t)b; };
    for (int s = 0; s <= steps; ++s) {
        if (s > 0) {
            ph.advectParticles(F.ptrs, dt, S);
            ph.projectVelocityOntoGrid(W.ptrs);
            ph.correctParticleVelocity(F.ptrs, W.ptrs);
            checkCudaErrors(cudaDeviceSynchronize());
        }
        const int n = ph.getParticleCount();
        hp.resize(n);
        copy_d2h(ph.getParticles(), hp.data(), n);
        checkCudaErrors(cudaDeviceSynchronize());
        uint64_t cs[6] = {0, 0, 0, 0, 0, 0};
        double vx = 0.0, vy = 0.0;
        for (int i = 0; i < n; ++i) {
            const Particle2D &p = hp[i];
            cs[0] += bits(p.getPosition().x);
            cs[1] += bits(p.getPosition().y);
            cs[2] += bits(p.getLocalPosition().x);
            cs[3] += bits(p.getLocalPosition().y);
            cs[4] += bits(p.getLocalPosition().z);
            cs[5] += (uint64_t)p.getCellID();
            vx += p.getVelocity().x;
            vy += p.getVelocity().y;
        }
        fprintf(real_out, "{\"step\": %d, \"count\": %d, \"checksum\": [%lld, %lld, %lld, %lld, %lld, %lld], \"vsum\": [%.17g, %.17g]}\n", s, n,
                (long long)cs[0], (long long)cs[1], (long long)cs[2], (long long)cs[3], (long long)cs[4], (long long)cs[5], vx, vy);
        fflush(real_out);
    }
    std::vector<double> w(2 * (size_t)N);
    copy_d2h(W.comp[0].data, w.data(), N);
    copy_d2h(W.comp[1].data, w.data() + N, N);
    checkCudaErrors(cudaDeviceSynchronize());
    FILE *g = fopen((prefix + "_w.bin").c_str(), "wb");
    if (!g) return 2;
    wr(g, w.data(), w.size());
    fclose(g);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc >= 12 && !strcmp(argv[1], "digest")) return run_digest(argc, argv);
    if (argc >= 4 && !strcmp(argv[1], "dump")) return run_dump(argv[2], argv[3]);
    if (argc >= 2 && !strcmp(argv[1], "time")) return run_time(argc, argv);
    if (argc >= 5 && !strcmp(argv[1], "timecase")) return run_timecase(argv[2], atoi(argv[3]), atoi(argv[4]));
    fprintf(stderr, "usage: ref_harness dump <case.bin> <out_prefix> | time ...\n");
    return 2;
}
