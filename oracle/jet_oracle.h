/*
 * jet_oracle.h — TEST INFRASTRUCTURE.  Common C interface of the two CPU oracles:
 *
 *   oracle/libjetport.so       jet_port.c  — a plain-C restatement of Jet's hybrid-liquid hot path
 *                              (jeto_kind() == "port"), every function citing the reference file:line.
 *   oracle/_ref/libjetref.so   ref_shim.cpp — the UNMODIFIED reference (compiled from /root/reference
 *                              by oracle/Makefile) behind the same functions (jeto_kind() == "reference").
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load these.
 * The product (include/jetcu.h, fluid-engine-dev_b200/) never links, loads or calls them.
 *
 * Signatures deliberately mirror include/jetcu.h (minus the ctx), so that a parity test is
 * "same arguments into jetcu_* and jeto_*, compare outputs".  Struct layouts are identical.
 */
#ifndef JET_ORACLE_H_
#define JET_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct jeto_row { double center, right, up, front; } jeto_row;

typedef struct jeto_mgpcg_params {
    uint32_t max_iterations, max_levels, n_restriction, n_correction, n_coarsest, n_final;
    double tolerance, sor_factor;
    int32_t use_red_black, linear_solver;   /* linear_solver: 0 MGPCG, 1 ICCG, | 2 = useCompressedLinearSystem */
} jeto_mgpcg_params;

typedef struct jeto_mgpcg_result {
    uint32_t iterations;
    int32_t returned;
    double last_residual, true_residual_l2, b_l2, solve_ms;
} jeto_mgpcg_result;

typedef struct jeto_grid {
    uint64_t nx, ny, nz;
    double hx, hy, hz, ox, oy, oz;
} jeto_grid;

typedef struct jeto_hybrid_params {
    jeto_grid grid;
    int32_t mode, closed_domain_boundary_flag;
    double pic_blending_factor;
    double gravity[3];
    double max_cfl;
    jeto_mgpcg_params mgpcg;
} jeto_hybrid_params;

typedef struct jeto_step_stats {
    uint32_t pcg_iterations;
    double pcg_last_residual, pcg_true_residual_l2, pcg_b_l2;
    double ms_total, ms_bin, ms_p2g, ms_sdf, ms_extrapolate, ms_bc, ms_pressure, ms_g2p, ms_move;
    double ms_mgpcg_solve;   /* layout of jetcu_step_stats; the oracles leave it 0 */
} jeto_step_stats;

const char* jeto_kind(void);          /* "port" or "reference" */
void jeto_set_num_threads(int n);     /* OpenMP threads used by the oracle (0 = all) */
int jeto_num_threads(void);

int jeto_fdm_mvm(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* v, double* result);
int jeto_fdm_residual(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* x, const double* b, double* result);
int jeto_fdm_relax_red_black(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b,
                             double sor_factor, uint32_t iterations, double* x);
int jeto_fdm_relax(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b,
                   double sor_factor, uint32_t iterations, double* x);   /* FdmGaussSeidelSolver3::relax (lexicographic) */
int jeto_fdm_restrict(uint64_t nx, uint64_t ny, uint64_t nz, const double* finer, double* coarser);
int jeto_fdm_correct(uint64_t nx, uint64_t ny, uint64_t nz, const double* coarser, double* finer);
int jeto_fdm_dot(uint64_t n, const double* a, const double* b, double* result);
int jeto_fdm_axpy(uint64_t n, double a, const double* x, const double* y, double* result);
int jeto_fdm_vcycle(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                    const double* b, double* x, const jeto_mgpcg_params* params);
int jeto_mgpcg_solve(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, const jeto_row* const* A,
                     const double* b, double* x, const jeto_mgpcg_params* params, jeto_mgpcg_result* result);
/* FdmIccgSolver3::solve(FdmLinearSystem3*) and its Preconditioner::build / solve (src/jet/fdm_iccg_solver3.cpp:14-70,152-188) */
int jeto_iccg_solve(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* x,
                    uint32_t max_iterations, double tolerance, jeto_mgpcg_result* result);
int jeto_iccg_precondition(uint64_t nx, uint64_t ny, uint64_t nz, const jeto_row* A, const double* b, double* d, double* y,
                           double* x);
uint32_t jeto_mg_num_levels(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t max_levels);

int jeto_pressure_solve(const jeto_grid* grid, const double* u_in, const double* v_in, const double* w_in,
                        const double* collider_sdf, const double* bnd_u, const double* bnd_v, const double* bnd_w,
                        const double* fluid_sdf, const jeto_mgpcg_params* params,
                        double* u_out, double* v_out, double* w_out, double* pressure_out, jeto_mgpcg_result* result);
int jeto_pressure_get_system(uint32_t level, jeto_row* A, double* b, float* fluid_sdf,
                             float* u_weights, float* v_weights, float* w_weights);

int jeto_bin_points(uint64_t n, const double* positions, uint64_t res_x, uint64_t res_y, uint64_t res_z,
                    double grid_spacing, uint64_t* keys_unsorted, uint64_t* sorted_indices,
                    uint64_t* keys_sorted, uint64_t* start_table, uint64_t* end_table);
/* PointHashGridSearcher3 (serial) bucket membership: bucket_of_point[i] = key, rank_in_bucket[i] = position of
 * point i inside buckets()[key]; reference only (the port returns the stable order by construction). */
int jeto_bin_points_serial(uint64_t n, const double* positions, uint64_t res_x, uint64_t res_y, uint64_t res_z,
                           double grid_spacing, uint64_t* bucket_of_point, uint64_t* rank_in_bucket);

typedef struct jeto_hybrid jeto_hybrid;
int jeto_hybrid_create(const jeto_hybrid_params* params, jeto_hybrid** out);
int jeto_hybrid_destroy(jeto_hybrid*);
int jeto_hybrid_set_particles(jeto_hybrid*, uint64_t n, const double* positions, const double* velocities);
int jeto_hybrid_append_particles(jeto_hybrid*, uint64_t n, const double* positions, const double* velocities);
int jeto_hybrid_emit_box(jeto_hybrid*, const double lower[3], const double upper[3], double spacing,
                         const double initial_velocity[3], uint64_t* n_emitted);
/* Analytic rigid collider (RigidBodyCollider3, identity transforms) for the hybrid solver.  One explicit surface, or an
 * ImplicitSurfaceSet3 of explicit surfaces (as_set != 0; needed for n > 1):
 *   kind 1 Box3 (a = lower corner, b = upper corner), 2 Sphere3 (a = centre, b[0] = radius),
 *        3 Cylinder3 (a = centre, b[0] = radius, b[1] = height), 4 Plane3 (a = point, b = unit normal);
 *   flipped = Surface3::isNormalFlipped.  n == 0 removes the collider.  Same layout as jetcu_collider_shape. */
typedef struct jeto_collider_shape {
    int32_t kind, flipped;
    double a[3], b[3];
} jeto_collider_shape;
/* VolumeParticleEmitter3 (one shot, jitter 0, GridPointGenerator3, maxRegion = domain) over one explicit surface or an
 * ImplicitSurfaceSet3 of them (shapes as for jeto_hybrid_set_collider_shapes). */
int jeto_hybrid_emit_shapes(jeto_hybrid*, const jeto_collider_shape* shapes, uint32_t n, int as_set,
                            int lattice /* 0 GridPointGenerator3, 1 BccLatticePointGenerator (the emitter's default) */, double spacing,
                            const double initial_velocity[3], uint64_t* n_emitted);
uint64_t jeto_hybrid_num_particles(const jeto_hybrid*);
int jeto_hybrid_get_particles(jeto_hybrid*, double* positions, double* velocities);
int jeto_hybrid_get_affine(jeto_hybrid*, double* cx, double* cy, double* cz);
int jeto_hybrid_set_collider_shapes(jeto_hybrid*, const jeto_collider_shape* shapes, uint32_t n, int as_set,
                                    const double linear_velocity[3], const double angular_velocity[3], double friction);
/* shorthand for one un-flipped box (kind 1) or sphere (kind 2); kind 0 removes the collider */
int jeto_hybrid_set_collider_shape(jeto_hybrid*, int kind, const double a[3], const double b[3],
                                   const double linear_velocity[3], const double angular_velocity[3], double friction);
/* GridFractionalBoundaryConditionSolver3::_colliderSdf after the next constrainVelocity / as rasterised (cell-centred). */
int jeto_hybrid_get_collider_sdf(jeto_hybrid*, double* sdf);
int jeto_hybrid_set_velocity(jeto_hybrid*, const double* u, const double* v, const double* w);
int jeto_hybrid_get_velocity(jeto_hybrid*, double* u, double* v, double* w);
int jeto_hybrid_get_sdf(jeto_hybrid*, double* sdf);
int jeto_hybrid_get_pressure(jeto_hybrid*, double* pressure);
int jeto_hybrid_get_markers(jeto_hybrid*, char* um, char* vm, char* wm);
int jeto_hybrid_cfl(jeto_hybrid*, double dt, double* cfl);
int jeto_hybrid_num_sub_steps(jeto_hybrid*, double dt, uint32_t* n);
int jeto_hybrid_step(jeto_hybrid*, double dt, jeto_step_stats* stats);
int jeto_hybrid_advance_frame(jeto_hybrid*, double frame_dt, uint32_t* n_sub_steps, jeto_step_stats* last_stats);
int jeto_hybrid_apply_boundary_condition(jeto_hybrid*);
int jeto_hybrid_transfer_particles_to_grids(jeto_hybrid*);
int jeto_hybrid_build_signed_distance_field(jeto_hybrid*);
int jeto_hybrid_extrapolate_velocity_to_air(jeto_hybrid*);
int jeto_hybrid_begin_advance_time_step(jeto_hybrid*, double dt);
int jeto_hybrid_compute_external_forces(jeto_hybrid*, double dt);
int jeto_hybrid_compute_pressure(jeto_hybrid*, double dt, jeto_mgpcg_result* result);
int jeto_hybrid_transfer_grids_to_particles(jeto_hybrid*);
int jeto_hybrid_move_particles(jeto_hybrid*, double dt);
int jeto_hybrid_compute_advection(jeto_hybrid*, double dt);
/* GridFluidSolver3::setViscosityCoefficient / computeViscosity (src/jet/grid_fluid_solver3.cpp:231-241): when > kEpsilonD the
 * step solves the three face grids with the solver's GridBackwardEulerDiffusionSolver3 (default boundary type, ICCG(100,
 * kEpsilonD)) between the external forces and the pressure projection.  results: the u, v, w ICCG solves (port) / the w solve
 * in all three slots (reference shim: the class exposes its last solve only). */
int jeto_hybrid_set_viscosity(jeto_hybrid*, double coefficient);
int jeto_hybrid_compute_viscosity(jeto_hybrid*, double dt, jeto_mgpcg_result results[3]);

/* ---------------------------------------------------------------- GridBackwardEulerDiffusionSolver3
 * solve(FaceCenteredGrid3) / solve(ScalarGrid3 = CellCenteredScalarGrid3) of src/jet/grid_backward_euler_diffusion_solver3.cpp
 * with FdmIccgSolver3(max_iterations, tolerance) (the class default is (100, kEpsilonD), :20-23).  boundary_type 0 = Dirichlet,
 * 1 = Neumann (header default).  collider_sdf / fluid_sdf: cell-centred SDF arrays (CellCenteredScalarGrid3, sampled linearly at
 * the data positions) or NULL for the constant fields the fluid solvers pass when there is none (+kMaxD / -kMaxD). */
int jeto_diffusion_solve_faces(const jeto_grid* grid, const double* u, const double* v, const double* w, double coefficient, double dt,
                               int boundary_type, const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations,
                               double tolerance, double* u_out, double* v_out, double* w_out, jeto_mgpcg_result results[3]);
int jeto_diffusion_solve_cells(const jeto_grid* grid, const double* source, double coefficient, double dt, int boundary_type,
                               const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations, double tolerance,
                               double* dest, char* markers /* optional: kFluid 0 / kAir 1 / kBoundary 2 (port only) */,
                               jeto_mgpcg_result* result);

/* ---------------------------------------------------------------- GridSmokeSolver3 (BASELINE config 5)
 * GridSmokeSolver3 (src/jet/grid_smoke_solver3.cpp) on GridFluidSolver3::onAdvanceTimeStep
 * (src/jet/grid_fluid_solver3.cpp:179-210): buoyancy, fractional pressure projection with MGPCG on an all-liquid domain
 * (fluidSdf = -kMaxD, no collider), cubic semi-Lagrangian advection (CubicSemiLagrangian3, the solver's default) of
 * density / temperature / velocity, decay.  No emitter, no collider, diffusion and viscosity coefficients 0 (defaults). */
typedef struct jeto_smoke_params {
    jeto_grid grid;
    int32_t closed_domain_boundary_flag;
    int32_t advection;   /* 0: CubicSemiLagrangian3 (the solver's default), 1: SemiLagrangian3 (linear sampler) */
    double gravity[3];
    double max_cfl;
    double buoyancy_density_factor, buoyancy_temperature_factor;   /* defaults -0.000625, 5.0 */
    double density_decay, temperature_decay;                       /* defaults 0.001, 0.001 */
    jeto_mgpcg_params mgpcg;
} jeto_smoke_params;
typedef struct jeto_smoke jeto_smoke;
int jeto_smoke_create(const jeto_smoke_params* params, jeto_smoke** out);
int jeto_smoke_destroy(jeto_smoke*);
/* any pointer may be NULL (field left as it is / not returned); cell-centred density & temperature, face-centred u, v, w */
int jeto_smoke_set_fields(jeto_smoke*, const double* density, const double* temperature, const double* u, const double* v,
                          const double* w);
int jeto_smoke_get_fields(jeto_smoke*, double* density, double* temperature, double* u, double* v, double* w, double* pressure);
/* VolumeGridEmitter3 with step-function targets on smoke density [dmin, dmax] and temperature [tmin, tmax], fired once (shapes
 * as for jeto_hybrid_set_collider_shapes): the smoke source of src/examples/smoke_sim. */
int jeto_smoke_emit_step(jeto_smoke*, const jeto_collider_shape* shapes, uint32_t n, int as_set, double dmin, double dmax,
                         double tmin, double tmax);
/* GridFluidSolver3::setCollider(RigidBodyCollider3 over analytic surfaces) on the smoke solver (src/examples/smoke_sim uses a
 * sphere): constrainVelocity's collider branch, the collider SDF in the pressure system, SemiLagrangian3's boundary handling
 * (cells inside the collider are not advected, back-traces stop at the surface) and extrapolateIntoCollider of the scalars. */
int jeto_smoke_set_collider_shapes(jeto_smoke*, const jeto_collider_shape* shapes, uint32_t n, int as_set,
                                   const double linear_velocity[3], const double angular_velocity[3], double friction);
int jeto_smoke_cfl(jeto_smoke*, double dt, double* cfl);                               /* GridFluidSolver3::cfl */
int jeto_smoke_num_sub_steps(jeto_smoke*, double dt, uint32_t* n);                   /* ::numberOfSubTimeSteps */
int jeto_smoke_step(jeto_smoke*, double dt, jeto_mgpcg_result* result);            /* onAdvanceTimeStep */
int jeto_smoke_compute_external_forces(jeto_smoke*, double dt);                    /* computeBuoyancyForce + BC */
int jeto_smoke_compute_pressure(jeto_smoke*, double dt, jeto_mgpcg_result* result);
int jeto_smoke_compute_advection(jeto_smoke*, double dt);
int jeto_smoke_end_advance_time_step(jeto_smoke*, double dt);                      /* computeDiffusion: diffusion + decay */
/* GridFluidSolver3::setViscosityCoefficient, GridSmokeSolver3::setSmokeDiffusionCoefficient / setTemperatureDiffusionCoefficient:
 * coefficients > kEpsilonD switch computeViscosity (step) / the scalar diffusion solves (end_advance_time_step) on. */
int jeto_smoke_set_diffusion(jeto_smoke*, double viscosity, double smoke_diffusion, double temperature_diffusion);
int jeto_smoke_compute_viscosity(jeto_smoke*, double dt);

#ifdef __cplusplus
}
#endif
#endif /* JET_ORACLE_H_ */
