/*
 * pfem2_oracle.h -- C ABI of the CPU oracle for the PFEM-2 particle step.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a C++/OpenMP restatement of the reference's CUDA
 * particle path (gpuPfem2 src/particles/particle_handler_2d.cu, particle_2d.cu, geometry.cuh,
 * mesh_2d.cu:21-34,107-139).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it; the product library (gpupfem2_b200/csrc) never does.
 *
 * Parity pin: checked against particle/nodal dumps produced by the reference's own CUDA code
 * (oracle/ref_harness.cu linked against oracle/_ref/libgpuPfem2Lib.so, run on a B200 through
 * gpurun); the dumps are committed under tests/golden/ with the script that made them.
 */
#ifndef PFEM2_ORACLE_H
#define PFEM2_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_handle orc_handle;

/* ---- mesh helpers (reference: src/mesh_2d.cu) ---- */

/* inverse Jacobian, 4 doubles per cell, same operation order as kCalculateInvJacobi (mesh_2d.cu:21-34) */
void orc_inv_jacobi(int n_cells, const double *vertices, const unsigned *cells, double *inv_jacobi);

/* vertex-sharing one-ring CSR in ascending order (mesh_2d.cu:107-139), O(C) instead of O(C^2).
 * offsets has n_cells+1 entries.  Call with indices == NULL to get offsets only (offsets[n_cells] = nnz). */
void orc_one_ring(int n_nodes, int n_cells, const unsigned *cells, int *offsets, int *indices);

/* ---- particle handler (reference: class ParticleHandler2D, particle_handler_2d.cuh:9-54) ---- */

/* subcell_mode: 0 = reference-exact flat index with spill, 1 = clamped.
 * max_level: reference value is 4 (constants.h:15); larger values are an extension. */
orc_handle *orc_create(int n_nodes, int n_cells, const double *vertices, const unsigned *cells,
                       const double *inv_jacobi, const int *nbr_offsets, const int *nbr_indices,
                       int cell_division_level, int max_level, int subcell_mode);
void orc_destroy(orc_handle *h);

int orc_particles_per_cell(const orc_handle *h);
/* sub-cell centres, 3 doubles each (particle_handler_2d.cu:248-274) */
void orc_subcell_centers(const orc_handle *h, double *out);

int orc_seed(orc_handle *h);                                                  /* :304-320 -> particle count */
void orc_init_velocity(orc_handle *h, const double *vx, const double *vy);    /* :322-326 */
int orc_advect(orc_handle *h, const double *vx, const double *vy, double dt, int substeps); /* :328-342 */
/* the two halves of orc_advect, and the projection without its division, for multi-rank (strip partition) tests */
int orc_move(orc_handle *h, const double *vx, const double *vy, double dt, int substeps);
int orc_check_distribution(orc_handle *h, const double *vx, const double *vy, int own_lo, int own_hi);
void orc_project_accumulate(orc_handle *h, double *acc3);
void orc_project(orc_handle *h, double *vx, double *vy);                      /* :350-361, writes in place */
void orc_correct(orc_handle *h, const double *vx, const double *vy,
                 const double *vx_old, const double *vy_old);                 /* :344-348 */

int orc_count(const orc_handle *h);
/* counters of the last orc_advect call: lost[substep] for up to 16 substeps, and added */
void orc_last_stats(const orc_handle *h, int *lost_per_substep, int *added);

/* copy particle state out / in (any pointer may be NULL) */
void orc_download(const orc_handle *h, double *x, double *y, double *l0, double *l1, double *l2,
                  double *vx, double *vy, unsigned *cell, unsigned *id);
void orc_upload(orc_handle *h, int n, const double *x, const double *y, const double *l0, const double *l1,
                const double *l2, const double *vx, const double *vy, const unsigned *cell, const unsigned *id);

/* element-wise device-function restatements, for unit tests */
void orc_to_local(const double *inv_jacobi4, const double *v3, double px, double py, double *L3);
int orc_inside(const double *L3);
int orc_subcell(const double *L3, int level);

int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
