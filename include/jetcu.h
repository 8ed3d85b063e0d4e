/*
 * jetcu.h — C ABI of the B200-native hybrid-liquid hot path (drop-in for Jet's CPU path).
 *
 * Everything here is plain C: opaque handles, plain pointers and sizes, `int` status returns
 * (0 = JETCU_OK; never throws, never aborts).  All array arguments are HOST pointers owned by
 * the caller unless the name ends in `_dev`; the context owns all device scratch.  Arrays use
 * Jet's Array3 layout: x-fastest, contiguous (reference: include/jet/detail/array3-inl.h:144).
 *
 * Each entry point cites the reference interface it replaces (paths relative to the Jet tree,
 * doyubkim/fluid-engine-dev).  INTEGRATION.md shows the C++ adapters that bind them.
 *
 * One ctx per GPU; calls on one ctx must be serialised by the caller (Jet's solvers are not
 * re-entrant either: SURVEY.md §8b).  There is NO CPU fallback: every call fails with
 * JETCU_ERR_CUDA if no sm_100-class device is usable.
 */
#ifndef JETCU_H_
#define JETCU_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ status */
enum {
    JETCU_OK = 0,
    JETCU_ERR_INVALID_ARG = 1, /* mirrors std::invalid_argument from JET_THROW_INVALID_ARG_IF (include/jet/macros.h:37-41) */
    JETCU_ERR_CUDA = 2,        /* CUDA runtime error; text in jetcu_last_error() */
    JETCU_ERR_UNSUPPORTED = 3, /* e.g. ICCG or lexicographic Gauss-Seidel under a multi-GPU communicator */
    JETCU_ERR_OOM = 4,
    JETCU_ERR_NCCL = 5
};

typedef struct jetcu_ctx jetcu_ctx;

/* Creates a context bound to CUDA device `device`. */
int jetcu_ctx_create(int device, jetcu_ctx** out);
int jetcu_ctx_destroy(jetcu_ctx* ctx);
/* Last error text for ctx (or for the calling thread's last failed create when ctx == NULL). */
const char* jetcu_last_error(const jetcu_ctx* ctx);
/* Library / build identification ("jetcu <ver> sm_100a"). */
const char* jetcu_version(void);
/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
uint64_t jetcu_launch_count(const jetcu_ctx* ctx);
/* cudaStream_t the ctx launches on (as void*), so callers can record CUDA events on it. */
void* jetcu_stream(const jetcu_ctx* ctx);
int jetcu_synchronize(jetcu_ctx* ctx);

/* Optional live timing of kernel families (bench.py's roofline): when enabled, groups of launches are bracketed
 * by CUDA events on the ctx stream.  `units` counts the family's work items (cells x sweeps for RELAX_L0, cells
 * for MVM_L0 ...), so achieved bytes/s = units * bytes_per_unit / total_ms.  Not part of the reference API. */
enum {
    JETCU_PROF_RELAX_L0 = 0,   /* red-black SOR sweeps on the finest MG level (unit: cell-sweep, 56 B) */
    JETCU_PROF_RELAX_COARSE,   /* the same on all coarser levels */
    JETCU_PROF_MVM_L0,         /* PCG matvec (unit: cell, 48 B) */
    JETCU_PROF_TRANSFER,       /* residual+restrict / correct on all levels (unit: fine cell) */
    JETCU_PROF_BLAS1,          /* dot / axpy / fused PCG updates (unit: element) */
    JETCU_PROF_P2G,            /* particle-to-grid gather, all three components (unit: particle) */
    JETCU_PROF_KINDS = 8
};
int jetcu_profile_enable(jetcu_ctx* ctx, int on);   /* on != 0 also clears the accumulators */
int jetcu_profile_get(jetcu_ctx* ctx, int kind, double* total_ms, uint64_t* launches, uint64_t* units);

/* ------------------------------------------------------------------ multi-GPU (z-slabs over NCCL / NVLink)
 * The reference has no domain decomposition (one process, one allocation: SURVEY.md §8e); this is the new part.
 * One process per GPU, one ctx per process.  After jetcu_comm_init the SAME calls (jetcu_mgpcg_solve,
 * jetcu_pressure_solve, jetcu_hybrid_*) run SPMD: every rank passes the same full-size grid arguments and receives
 * the same full-size results, the MGPCG solve inside works on this rank's z-slab with one halo-plane exchange per
 * smoother colour pass / matvec / transfer, slab-invariant dot products (every rank sees the same PCG scalars) and
 * the coarse MG levels replicated; particles live on the rank that owns their cell plane and migrate each step.
 * Results are bit-identical to the single-GPU run.  nz must be a multiple of 2*world with >= 4 planes per rank.
 * The 128-byte NCCL id is created on rank 0 and handed to the other ranks out of band (torch.distributed, MPI ...). */
#define JETCU_UNIQUE_ID_BYTES 128
int jetcu_comm_unique_id(uint8_t id[JETCU_UNIQUE_ID_BYTES]);
int jetcu_comm_init(jetcu_ctx* ctx, int world, int rank, const uint8_t id[JETCU_UNIQUE_ID_BYTES]);
int jetcu_comm_destroy(jetcu_ctx* ctx);
int jetcu_comm_world(const jetcu_ctx* ctx);
int jetcu_comm_rank(const jetcu_ctx* ctx);
/* messages / payload bytes this rank has sent since jetcu_comm_init (bench.py's halo accounting) */
int jetcu_comm_stats(const jetcu_ctx* ctx, uint64_t* messages, uint64_t* bytes_sent);
/* 1 when halo planes travel through peer memory (JETCU_P2P_HALO=1 at jetcu_comm_init: a CUDA-IPC mailbox per rank, written by
 * the neighbours' kernels over NVLink, two kernels per exchange and no host involvement) instead of NCCL send/recv; the switch
 * is taken by all ranks together after a self-checking exchange, and never silently: it falls back with a line on stderr.
 * Results are bit-identical either way. */
int jetcu_comm_p2p_enabled(const jetcu_ctx*);
/* Diagnostic: average device time (ms) of one halo exchange of `depth` planes of `plane_bytes` with both neighbours, over
 * `iters` back-to-back exchanges on a scratch slab — the latency that bounds the slab-distributed V-cycle. */
int jetcu_comm_halo_benchmark(jetcu_ctx*, uint64_t plane_bytes, int depth, int iters, double* ms_per_exchange);
/* cell planes [z0, z1) of an nz-plane grid owned by `rank` (uniform slabs; nz % world must be 0) */
int jetcu_slab_range(uint64_t nz, int world, int rank, uint64_t* z0, uint64_t* z1);
/* Host-only views of the decomposition logic (no GPU needed; what the world_size-2 CPU tests drive):
 * - the MG plan: levels [0, *dist_levels) are slab-distributed with this rank's planes [z0[l], z1[l]), coarser levels
 *   are replicated (z0 = 0, z1 = nz >> l); fails if level 0 cannot be split (nz % (2*world) != 0 or < 4 planes/rank);
 * - the particle routing rule: bit q set = a particle in `cell_plane` held by rank `me` is copied to rank q (q owns the
 *   plane or keeps it as one of its JETCU_SLAB_GHOST_PLANES ghost planes). */
#define JETCU_SLAB_GHOST_PLANES 2
int jetcu_slab_plan(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels, int world, int rank,
                    uint32_t* dist_levels, uint64_t* z0, uint64_t* z1);
uint64_t jetcu_slab_route_mask(int64_t cell_plane, uint64_t nz, int world, int me);

/* ------------------------------------------------------------------ types */

/* == jet::FdmMatrixRow3 (include/jet/fdm_linear_system3.h:18-30): 32-byte AoS row. */
typedef struct jetcu_matrix_row {
    double center, right, up, front;
} jetcu_matrix_row;

/* == ctor arguments of jet::FdmMgpcgSolver3 (include/jet/fdm_mgpcg_solver3.h:35-41). */
typedef struct jetcu_mgpcg_params {
    uint32_t max_iterations;    /* numberOfCgIter */
    uint32_t max_levels;        /* maxNumberOfLevels */
    uint32_t n_restriction;     /* numberOfRestrictionIter (pre-smoothing sweeps) */
    uint32_t n_correction;      /* numberOfCorrectionIter (post-smoothing sweeps, levels > 0) */
    uint32_t n_coarsest;        /* numberOfCoarsestIter */
    uint32_t n_final;           /* numberOfFinalIter (post-smoothing sweeps on level 0) */
    double tolerance;           /* maxTolerance */
    double sor_factor;          /* sorFactor */
    int32_t use_red_black;      /* useRedBlackOrdering; 0 = the natural (lexicographic) Gauss-Seidel ordering, the reference
                                   constructors' default: tile-hyperplane wavefront on the device, single GPU only
                                   (JETCU_ERR_UNSUPPORTED under a multi-rank communicator) */
    int32_t linear_solver;      /* which FdmLinearSystemSolver3 the pressure solve runs: JETCU_SOLVER_MGPCG (0, all fields
                                   above apply) or JETCU_SOLVER_ICCG (FdmIccgSolver3: max_iterations and tolerance apply,
                                   one level), optionally | JETCU_SOLVER_COMPRESSED (see jetcu_pressure_solve) */
} jetcu_mgpcg_params;
enum { JETCU_SOLVER_MGPCG = 0, JETCU_SOLVER_ICCG = 1, JETCU_SOLVER_COMPRESSED = 2 };

typedef struct jetcu_mgpcg_result {
    uint32_t iterations;        /* FdmMgpcgSolver3::lastNumberOfIterations() */
    int32_t returned;           /* what FdmMgpcgSolver3::solve() returns (src/jet/fdm_mgpcg_solver3.cpp:78-79) */
    double last_residual;       /* FdmMgpcgSolver3::lastResidual() = sqrt(|r.M^-1 r|) */
    double true_residual_l2;    /* ||b - A x||_2 recomputed at exit (not in the reference) */
    double b_l2;                /* ||b||_2 */
    double solve_ms;            /* device time of the solve (CUDA events) */
} jetcu_mgpcg_result;

/* Uniform grid descriptor == jet::Grid3 (resolution, gridSpacing, origin; src/jet/grid3.cpp:30-75). */
typedef struct jetcu_grid {
    uint64_t nx, ny, nz;
    double hx, hy, hz;
    double ox, oy, oz;
} jetcu_grid;

/* ------------------------------------------------ seam 1: linear-system level
 * FdmBlas3 / FdmGaussSeidelSolver3 / FdmMgUtils3 building blocks, exposed so that every kernel
 * can be checked bit-for-bit against the reference on the same inputs.  Sizes are the size of
 * the (finer) array; vectors are nx*ny*nz doubles.                                            */

/* FdmBlas3::mvm (src/jet/fdm_linear_system3.cpp:79-96). */
int jetcu_fdm_mvm(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                  const jetcu_matrix_row* A, const double* v, double* result);
/* FdmBlas3::residual (src/jet/fdm_linear_system3.cpp:98-116): result = b - A x. */
int jetcu_fdm_residual(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                       const jetcu_matrix_row* A, const double* x, const double* b, double* result);
/* `iterations` x FdmGaussSeidelSolver3::relaxRedBlack (src/jet/fdm_gauss_seidel_solver3.cpp:150-221), in place on x. */
int jetcu_fdm_relax_red_black(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                              const jetcu_matrix_row* A, const double* b, double sor_factor,
                              uint32_t iterations, double* x);
/* `iterations` x FdmGaussSeidelSolver3::relax (src/jet/fdm_gauss_seidel_solver3.cpp:103-120), the natural (lexicographic)
 * ordering — what FdmMgSolver3 / FdmMgpcgSolver3 smooth with unless useRedBlackOrdering is set — in place on x.  The
 * reference's loop is serial; a cell depends on its three lower neighbours' new and its three upper neighbours' old values,
 * so the device sweeps 8^3 tiles in tile-hyperplane order, consecutive sweeps pipelined two hyperplanes apart: the same
 * operands per cell, bit-identical to the serial loop. */
int jetcu_fdm_relax(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                    const jetcu_matrix_row* A, const double* b, double sor_factor,
                    uint32_t iterations, double* x);
/* FdmMgUtils3::restrict (src/jet/fdm_mg_linear_system3.cpp:43-98); (nx,ny,nz) = finer size, all even. */
int jetcu_fdm_restrict(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                       const double* finer, double* coarser);
/* FdmMgUtils3::correct (src/jet/fdm_mg_linear_system3.cpp:100-179): finer += P coarser.
 * The reference's k-clamp tests n.y instead of n.z (:159); we implement the intended n.z test, which is identical whenever
 * ny == nz — the shapes on which the reference is well defined (for ny > nz it reads a coarse plane beyond the array) and on
 * which every parity claim of this library is made.  ny != nz is accepted as an EXTENSION with the n.z clamp: a multi-level
 * V-cycle on the 1 x 2 x 1 / 3 x 2 x 1.5 domains of Jet's examples, where Jet itself can only run ICCG. */
int jetcu_fdm_correct(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz,
                      const double* coarser, double* finer);
/* FdmBlas3::dot (src/jet/fdm_linear_system3.cpp:49-65). Deterministic two-stage tree (not the serial order). */
int jetcu_fdm_dot(jetcu_ctx*, uint64_t n, const double* a, const double* b, double* result);
/* FdmBlas3::axpy (src/jet/fdm_linear_system3.cpp:67-77): result = a*x + y. */
int jetcu_fdm_axpy(jetcu_ctx*, uint64_t n, double a, const double* x, const double* y, double* result);

/* One MG V-cycle == FdmMgSolver3::solve(FdmMgLinearSystem3*) (src/jet/fdm_mg_solver3.cpp:69-74,
 * include/jet/detail/mg-inl.h:16-68) on x (in/out, finest level) with rhs b.  A[l] is level l's matrix
 * (level 0 = finest = nx*ny*nz, each next level halves every axis). */
int jetcu_fdm_vcycle(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels,
                     const jetcu_matrix_row* const* A, const double* b, double* x,
                     const jetcu_mgpcg_params* params);

/* FdmMgpcgSolver3::solve(FdmMgLinearSystem3*) (src/jet/fdm_mgpcg_solver3.cpp:54-80 +
 * include/jet/detail/cg-inl.h:18-94).  x is output only (the reference zeroes it, :61). */
int jetcu_mgpcg_solve(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz, uint32_t levels,
                      const jetcu_matrix_row* const* A, const double* b, double* x,
                      const jetcu_mgpcg_params* params, jetcu_mgpcg_result* result);

/* == jet::FdmIccgSolver3::solve(FdmLinearSystem3*) (src/jet/fdm_iccg_solver3.cpp:152-188): incomplete-Cholesky preconditioned
 * CG, x = 0 initial guess, exit on sigma <= tolerance^2 or max_iterations; the preconditioner's lexicographic recurrences
 * (:14-70) run as tile-hyperplane sweeps with the reference's per-cell arithmetic.  The default linear solver of Jet's
 * pressure solvers (100 iterations, 1e-6).  Single GPU. */
int jetcu_iccg_solve(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* b, double* x,
                     uint32_t max_iterations, double tolerance, jetcu_mgpcg_result* result);
/* Preconditioner::build + Preconditioner::solve on their own (parity tests): d (optional), y and x receive the incomplete-
 * Cholesky diagonal, the forward and the backward substitution of b. */
int jetcu_iccg_precondition(jetcu_ctx*, uint64_t nx, uint64_t ny, uint64_t nz, const jetcu_matrix_row* A, const double* b,
                            double* d, double* y, double* x);

/* == jet::GridBackwardEulerDiffusionSolver3::solve (src/jet/grid_backward_euler_diffusion_solver3.cpp), SURVEY §8f-4: the
 * FaceCenteredGrid3 overload (:108-160: u, v, w one after the other, each with its own markers) and the ScalarGrid3 overload
 * (:22-50) on a CellCenteredScalarGrid3.  Host buffers, x-fastest.  boundary_type 0 = Dirichlet, 1 = Neumann (the class
 * default, header :38).  collider_sdf / fluid_sdf: cell-centred SDF arrays sampled linearly at the data positions, or NULL
 * for the constant fields the fluid solvers pass when there is none (collider grid filled with kMaxD; ConstantScalarField3(
 * -kMaxD)).  Linear solver: FdmIccgSolver3(max_iterations, tolerance) — the class default is (100, kEpsilonD), :20-23.
 * markers (optional): kFluid 0 / kAir 1 / kBoundary 2 per sample (:16-18).  Single GPU. */
int jetcu_diffusion_solve_faces(jetcu_ctx*, const jetcu_grid* grid, const double* u, const double* v, const double* w,
                                double coefficient, double dt, int boundary_type, const double* collider_sdf,
                                const double* fluid_sdf, uint32_t max_iterations, double tolerance, double* u_out, double* v_out,
                                double* w_out, jetcu_mgpcg_result results[3]);
int jetcu_diffusion_solve_cells(jetcu_ctx*, const jetcu_grid* grid, const double* source, double coefficient, double dt,
                                int boundary_type, const double* collider_sdf, const double* fluid_sdf, uint32_t max_iterations,
                                double tolerance, double* dest, char* markers, jetcu_mgpcg_result* result);

/* Number of MG levels Jet builds for a finest resolution (FdmMgUtils3::resizeArrayWithFinest,
 * include/jet/detail/fdm_mg_linear_system3-inl.h:34-50). */
uint32_t jetcu_mg_num_levels(uint64_t nx, uint64_t ny, uint64_t nz, uint32_t max_levels);

/* ------------------------------------------------ seam 2: pressure projection
 * GridFractionalSinglePhasePressureSolver3::solve (src/jet/grid_fractional_single_phase_pressure_solver3.cpp:456-484)
 * with an FdmMgpcgSolver3 installed through setLinearSystemSolver (:497-510), i.e. buildWeights (:520-636),
 * buildSystem (:651-713, uncompressed MG path), MGPCG, applyPressureGradient (:715-765).
 *
 *  u_in/v_in/w_in    face-centred input velocity, sizes (nx+1,ny,nz) / (nx,ny+1,nz) / (nx,ny,nz+1)
 *  collider_sdf      cell-centred boundary SDF (a CellCenteredScalarGrid3's data()); NULL => +kMaxD (no collider)
 *  bnd_u/bnd_v/bnd_w boundaryVelocity sampled at the u/v/w face positions (x/y/z component resp.); NULL => 0
 *  fluid_sdf         cell-centred fluid SDF; NULL => -kMaxD (all liquid, GridFluidSolver3::fluidSdf, src/jet/grid_fluid_solver3.cpp:313-315)
 *  u_out/v_out/w_out in/out: faces the reference does not touch keep their incoming value (may alias the inputs)
 *  pressure_out      nullable; == GridFractionalSinglePhasePressureSolver3::pressure()
 */
int jetcu_pressure_solve(jetcu_ctx*, const jetcu_grid* grid,
                         const double* u_in, const double* v_in, const double* w_in,
                         const double* collider_sdf,
                         const double* bnd_u, const double* bnd_v, const double* bnd_w,
                         const double* fluid_sdf,
                         const jetcu_mgpcg_params* params,
                         double* u_out, double* v_out, double* w_out,
                         double* pressure_out, jetcu_mgpcg_result* result);

/* Test/inspection hook: copies out what buildWeights/buildSystem produced in the last jetcu_pressure_solve
 * on this ctx.  Any pointer may be NULL.  level < number of levels built. */
int jetcu_pressure_get_system(jetcu_ctx*, uint32_t level, jetcu_matrix_row* A, double* b,
                              float* fluid_sdf, float* u_weights, float* v_weights, float* w_weights);

/* Test/inspection hook: one MG V-cycle (as jetcu_fdm_vcycle) on the device-resident system the last
 * jetcu_pressure_solve on this ctx built, i.e. through the packed-matrix kernels the solver itself uses.
 * b, x: finest-level host vectors; x is in/out. */
int jetcu_pressure_vcycle(jetcu_ctx*, const double* b, double* x, const jetcu_mgpcg_params* params);

/* ------------------------------------------------ particle binning
 * PointParallelHashGridSearcher3::build (src/jet/point_parallel_hash_grid_searcher3.cpp:52-146) with the
 * key arithmetic of getHashKeyFromPosition (:232-269).  positions are AoS xyz doubles (Vector3D).
 * The sort is STABLE, so inside a bucket points are in ascending original index, i.e. exactly the order of
 * PointHashGridSearcher3::buckets() (src/jet/point_hash_grid_searcher3.cpp:50-69).  Empty buckets hold
 * JETCU_EMPTY_BUCKET (== jet::kMaxSize) in both tables.  Any output may be NULL. */
#define JETCU_EMPTY_BUCKET UINT64_MAX
int jetcu_bin_points(jetcu_ctx*, uint64_t n, const double* positions,
                     uint64_t res_x, uint64_t res_y, uint64_t res_z, double grid_spacing,
                     uint64_t* keys_unsorted,   /* n: getHashKeyFromPosition(points[i]) */
                     uint64_t* sorted_indices,  /* n: PointParallelHashGridSearcher3::sortedIndices() */
                     uint64_t* keys_sorted,     /* n: ::keys() */
                     uint64_t* start_table,     /* res_x*res_y*res_z: ::startIndexTable() */
                     uint64_t* end_table);      /* ::endIndexTable() */

/* ------------------------------------------------ seam 3: device-resident PIC / FLIP / APIC solver
 * Mirrors PicSolver3 / FlipSolver3 / ApicSolver3 (src/jet/{pic,flip,apic}_solver3.cpp) driven by
 * GridFluidSolver3::onAdvanceTimeStep (src/jet/grid_fluid_solver3.cpp:179-210).  Particles and grids stay
 * in HBM between calls; host copies happen only in the set_* / get_* calls. */
typedef struct jetcu_hybrid jetcu_hybrid;

enum { JETCU_PIC = 0, JETCU_FLIP = 1, JETCU_APIC = 2 };

/* == jet::kDirection* (include/jet/constants.h). */
enum {
    JETCU_DIR_LEFT = 1, JETCU_DIR_RIGHT = 2, JETCU_DIR_DOWN = 4, JETCU_DIR_UP = 8,
    JETCU_DIR_BACK = 16, JETCU_DIR_FRONT = 32, JETCU_DIR_ALL = 63
};

typedef struct jetcu_hybrid_params {
    jetcu_grid grid;
    int32_t mode;                        /* JETCU_PIC / FLIP / APIC */
    int32_t closed_domain_boundary_flag; /* GridFluidSolver3::closedDomainBoundaryFlag (default JETCU_DIR_ALL) */
    double pic_blending_factor;          /* FlipSolver3::picBlendingFactor (default 0) */
    double gravity[3];                   /* default (0,-9.8,0) */
    double max_cfl;                      /* default 5 */
    jetcu_mgpcg_params mgpcg;            /* pressure solver = fractional + MGPCG */
} jetcu_hybrid_params;

typedef struct jetcu_step_stats {
    uint32_t pcg_iterations;
    double pcg_last_residual;
    double pcg_true_residual_l2;
    double pcg_b_l2;
    double ms_total, ms_bin, ms_p2g, ms_sdf, ms_extrapolate, ms_bc, ms_pressure, ms_g2p, ms_move;
    double ms_mgpcg_solve;      /* device time of the MGPCG solve alone (ms_pressure also holds the system build and the gradient) */
} jetcu_step_stats;

void jetcu_hybrid_default_params(jetcu_hybrid_params* p);
int jetcu_hybrid_create(jetcu_ctx*, const jetcu_hybrid_params* params, jetcu_hybrid** out);
int jetcu_hybrid_destroy(jetcu_hybrid*);
/* GridFluidSolver3's setters between steps (setGravity, setMaxCfl, setClosedDomainBoundaryFlag,
 * FlipSolver3::setPicBlendingFactor, a re-configured linear solver): everything in `params` except grid and mode, which are
 * fixed at creation (JETCU_ERR_INVALID_ARG if they differ).  The device-resident state is untouched. */
int jetcu_hybrid_update_params(jetcu_hybrid*, const jetcu_hybrid_params* params);

/* ParticleSystemData3::positions()/velocities() (AoS xyz doubles). set replaces all particles. */
int jetcu_hybrid_set_particles(jetcu_hybrid*, uint64_t n, const double* positions, const double* velocities);
/* ParticleSystemData3::addParticles (src/jet/particle_system_data3.cpp:158-197) as a host emitter uses it between steps:
 * appends n_new particles (original indices num_particles() ...) to the device-resident set, which keeps its state —
 * in particular ApicSolver3's affine vectors, which `_cX.resize(n)` (src/jet/apic_solver3.cpp:36-38) preserves for the old
 * particles and zero-initialises for the new ones.  velocities may be NULL (zero). */
int jetcu_hybrid_append_particles(jetcu_hybrid*, uint64_t n_new, const double* positions, const double* velocities);
/* Emits the lattice VolumeParticleEmitter3 + GridPointGenerator3 produce for an axis-aligned box
 * (src/jet/volume_particle_emitter3.cpp:65-140, src/jet/grid_point_generator3.cpp:12-37; jitter 0, one shot),
 * directly on the device, in the reference's k,j,i order.  Appends; returns the count via n_emitted. */
int jetcu_hybrid_emit_box(jetcu_hybrid*, const double lower[3], const double upper[3], double spacing,
                          const double initial_velocity[3], uint64_t* n_emitted);
uint64_t jetcu_hybrid_num_particles(const jetcu_hybrid*);        /* all ranks together in slab mode */
uint64_t jetcu_hybrid_num_local_particles(const jetcu_hybrid*);  /* owned by this rank (== the above on one GPU) */
/* Returned in the caller's ORIGINAL particle order (the device keeps particles cell-sorted). */
int jetcu_hybrid_get_particles(jetcu_hybrid*, double* positions, double* velocities);
/* Slab mode (jetcu_comm_init'ed ctx): jetcu_hybrid_set_particles takes the WHOLE particle set on every rank and keeps
 * the particles whose cell plane the rank owns; jetcu_hybrid_get_particles fills the entries of the particles the rank
 * owns and zeroes the others (summing the ranks' arrays gives the whole set).  The two calls below move only a rank's
 * own particles: `ids` are the original indices (< n_global), positions / velocities AoS in the same order.  The
 * particles handed to set_local may lie anywhere: the next step migrates them to the owner of their cell plane. */
int jetcu_hybrid_get_local_particles(jetcu_hybrid*, uint64_t* ids, double* positions, double* velocities);
int jetcu_hybrid_set_local_particles(jetcu_hybrid*, uint64_t n_local, const uint64_t* ids, const double* positions,
                                     const double* velocities, uint64_t n_global);
/* APIC affine vectors _cX,_cY,_cZ (src/jet/apic_solver3.h), original order; NULL entries skipped. */
int jetcu_hybrid_get_affine(jetcu_hybrid*, double* cx, double* cy, double* cz);

/* Collider as the grid solver sees it: cell-centred SDF (GridFractionalBoundaryConditionSolver3::onColliderUpdated,
 * src/jet/grid_fractional_boundary_condition_solver3.cpp:244-280) + rigid-body velocity
 * (RigidBodyCollider3::velocityAt, src/jet/rigid_body_collider3.cpp:25-28) + friction.  sdf == NULL => no collider. */
int jetcu_hybrid_set_collider(jetcu_hybrid*, const double* collider_sdf, const double linear_velocity[3],
                              const double angular_velocity[3], const double rotation_origin[3],
                              double friction_coefficient);

/* Analytic rigid collider == RigidBodyCollider3 (src/jet/rigid_body_collider3.cpp) installed with
 * GridFluidSolver3::setCollider, identity transforms.  Its surface is ONE explicit surface (as_set == 0, n == 1) or an
 * ImplicitSurfaceSet3 of up to JETCU_COLLIDER_MAX_SHAPES explicit surfaces (as_set != 0; src/jet/implicit_surface_set3.cpp,
 * members wrapped in SurfaceToImplicit3 as addExplicitSurface does, queried through the set's Bvh3 in the reference's
 * traversal order):
 *   BOX      Box3      (src/jet/box3.cpp)       a = lower corner, b = upper corner
 *   SPHERE   Sphere3   (src/jet/sphere3.cpp)    a = centre, b[0] = radius
 *   CYLINDER Cylinder3 (src/jet/cylinder3.cpp)  a = centre, b[0] = radius, b[1] = height (axis y)
 *   PLANE    Plane3    (src/jet/plane3.cpp)     a = point, b = unit normal
 * flipped = Surface3::isNormalFlipped (the spherical tank of src/examples/hybrid_liquid_sim/main.cpp:454-463); n == 0
 * removes the collider.  Drives the collider branch of GridFractionalBoundaryConditionSolver3::constrainVelocity
 * (src/jet/grid_fractional_boundary_condition_solver3.cpp:26-280), extrapolateIntoCollider
 * (src/jet/grid_fluid_solver3.cpp:355-369), the moving-boundary terms of the pressure system and
 * Collider3::resolveCollision (src/jet/collider3.cpp:19-74) after moveParticles.  [Written against the pinned oracle after
 * round 1's GPU minutes were spent: bit-identical to it on the CPU (tests/test_hostcheck_collider.py);
 * tests/firstrun/zz_collider.py is the first GPU run.] */
enum { JETCU_COLLIDER_NONE = 0, JETCU_COLLIDER_BOX = 1, JETCU_COLLIDER_SPHERE = 2, JETCU_COLLIDER_CYLINDER = 3,
       JETCU_COLLIDER_PLANE = 4 };
#define JETCU_COLLIDER_MAX_SHAPES 8
typedef struct jetcu_collider_shape {
    int32_t kind;     /* JETCU_COLLIDER_BOX .. JETCU_COLLIDER_PLANE */
    int32_t flipped;  /* Surface3::isNormalFlipped */
    double a[3], b[3];
} jetcu_collider_shape;
int jetcu_hybrid_set_collider_shapes(jetcu_hybrid*, const jetcu_collider_shape* shapes, uint32_t n, int as_set,
                                     const double linear_velocity[3], const double angular_velocity[3], double friction);
/* shorthand for one un-flipped surface of `kind`; JETCU_COLLIDER_NONE removes the collider */
int jetcu_hybrid_set_collider_shape(jetcu_hybrid*, int kind, const double a[3], const double b[3],
                                    const double linear_velocity[3], const double angular_velocity[3], double friction);
/* The particle frame files of the reference's examples (saveParticleAsPos / saveParticleAsXyz,
 * src/examples/hybrid_liquid_sim/main.cpp:29-63), byte-identical to what they write for the same positions:
 *   POS  jet::serialize(const uint8_t*, size_t, ...) (src/jet/serialization.cpp:25-38): a FlatBuffers `FlatData` table — a
 *        fixed 24-byte frame (root offset, vtable, table, u32 length), the raw Vector3D array, zero padding to 4 bytes;
 *   XYZ  one "x y z\n" line per particle, doubles as operator<< prints them by default ("%g").
 * jetcu_hybrid_save_positions takes the positions from the device in the original particle order.  The remaining entry points
 * are host-only and usable without a GPU. */
enum { JETCU_FRAME_POS = 0, JETCU_FRAME_XYZ = 1 };
int jetcu_hybrid_save_positions(jetcu_hybrid*, const char* path, int format);
uint64_t jetcu_pos_frame_size(uint64_t payload_bytes);
int jetcu_pos_frame_header(uint64_t payload_bytes, uint8_t header[24]);
int jetcu_format_xyz(const double* positions, uint64_t n, char* out, uint64_t cap, uint64_t* out_size);
int jetcu_write_pos_file(const char* path, const double* positions, uint64_t n);
int jetcu_write_xyz_file(const char* path, const double* positions, uint64_t n);

/* VolumeParticleEmitter3 (src/jet/volume_particle_emitter3.cpp:65-142; one shot, jitter 0, GridPointGenerator3, maxRegion =
 * the domain) over an analytic implicit surface, on the device: one explicit surface (Builder::withSurface wraps it in
 * SurfaceToImplicit3) or an ImplicitSurfaceSet3 of explicit surfaces, on either of Jet's point generators — the emitters of
 * src/examples/hybrid_liquid_sim
 * (half-space + sphere, main.cpp:110-141; two boxes, :220-246).  Region = domain ∩ surface bounding box when bounded; every
 * lattice point with signedDistance <= 0 is appended in the generator's k, j, i order with velocity `initial_velocity`.
 * Shapes as for jetcu_hybrid_set_collider_shapes.  jetcu_hybrid_emit_box is the single-Box3 form. */
enum { JETCU_LATTICE_GRID = 0,   /* GridPointGenerator3 (src/jet/grid_point_generator3.cpp:15-38): what hybrid_liquid_sim installs */
       JETCU_LATTICE_BCC = 1 }; /* BccLatticePointGenerator (src/jet/bcc_lattice_point_generator.cpp:12-45): the emitter's default */
int jetcu_hybrid_emit_shapes(jetcu_hybrid*, const jetcu_collider_shape* shapes, uint32_t n, int as_set, int lattice, double spacing,
                             const double initial_velocity[3], uint64_t* n_emitted);
/* GridFractionalBoundaryConditionSolver3::colliderSdf() (cell-centred); kMaxD everywhere without a collider. */
int jetcu_hybrid_get_collider_sdf(jetcu_hybrid*, double* sdf);
int jetcu_hybrid_set_velocity(jetcu_hybrid*, const double* u, const double* v, const double* w);
int jetcu_hybrid_get_velocity(jetcu_hybrid*, double* u, double* v, double* w);
int jetcu_hybrid_get_sdf(jetcu_hybrid*, double* sdf);              /* PicSolver3::signedDistanceField() */
int jetcu_hybrid_get_pressure(jetcu_hybrid*, double* pressure);
int jetcu_hybrid_get_markers(jetcu_hybrid*, char* um, char* vm, char* wm); /* PicSolver3::_uMarkers.. */

/* GridFluidSolver3::cfl (src/jet/grid_fluid_solver3.cpp:57-72; note: no abs) and numberOfSubTimeSteps (:212-217). */
int jetcu_hybrid_cfl(jetcu_hybrid*, double dt, double* cfl);
int jetcu_hybrid_num_sub_steps(jetcu_hybrid*, double dt, uint32_t* n);

/* One GridFluidSolver3::onAdvanceTimeStep(dt) (src/jet/grid_fluid_solver3.cpp:179-210). stats may be NULL. */
int jetcu_hybrid_step(jetcu_hybrid*, double dt, jetcu_step_stats* stats);
/* PhysicsAnimation::advanceTimeStep adaptive sub-stepping over one frame (src/jet/physics_animation.cpp:75-123). */
int jetcu_hybrid_advance_frame(jetcu_hybrid*, double frame_dt, uint32_t* n_sub_steps, jetcu_step_stats* last_stats);

/* Stage-level entry points == the virtuals an adapter overrides (include/jet/pic_solver3.h:75-81,
 * include/jet/grid_fluid_solver3.h:208-240); also what the parity tests drive one by one. */
int jetcu_hybrid_apply_boundary_condition(jetcu_hybrid*);            /* GridFluidSolver3::applyBoundaryCondition */
int jetcu_hybrid_transfer_particles_to_grids(jetcu_hybrid*);         /* transferFromParticlesToGrids (PIC/FLIP/APIC) */
int jetcu_hybrid_build_signed_distance_field(jetcu_hybrid*);         /* PicSolver3::buildSignedDistanceField */
int jetcu_hybrid_extrapolate_velocity_to_air(jetcu_hybrid*);         /* PicSolver3::extrapolateVelocityToAir */
int jetcu_hybrid_begin_advance_time_step(jetcu_hybrid*, double dt);  /* beginAdvanceTimeStep incl. PicSolver3::onBeginAdvanceTimeStep */
int jetcu_hybrid_compute_external_forces(jetcu_hybrid*, double dt);  /* computeGravity + BC */
int jetcu_hybrid_compute_pressure(jetcu_hybrid*, double dt, jetcu_mgpcg_result* result); /* computePressure + BC */
/* == GridFluidSolver3::setViscosityCoefficient / computeViscosity (src/jet/grid_fluid_solver3.cpp:231-241).  A coefficient
 * > kEpsilonD makes jetcu_hybrid_step run computeViscosity between the external forces and the pressure projection: the three
 * face grids through GridBackwardEulerDiffusionSolver3 (its default boundary type, Neumann; ICCG(100, kEpsilonD)) with
 * boundarySdf = the collider SDF and fluidSdf = the particle SDF, then applyBoundaryCondition.  results (optional): the u, v, w
 * ICCG solves.  Single GPU (ICCG is not slab-distributed: JETCU_ERR_UNSUPPORTED under a communicator). */
int jetcu_hybrid_set_viscosity(jetcu_hybrid*, double coefficient);
int jetcu_hybrid_compute_viscosity(jetcu_hybrid*, double dt, jetcu_mgpcg_result results[3]);
int jetcu_hybrid_transfer_grids_to_particles(jetcu_hybrid*);         /* transferFromGridsToParticles */
int jetcu_hybrid_move_particles(jetcu_hybrid*, double dt);           /* PicSolver3::moveParticles */
int jetcu_hybrid_compute_advection(jetcu_hybrid*, double dt);        /* PicSolver3::computeAdvection */

/* ------------------------------------------------ device-resident smoke solver (BASELINE config 5)
 * Mirrors GridSmokeSolver3 (src/jet/grid_smoke_solver3.cpp) on GridFluidSolver3::onAdvanceTimeStep
 * (src/jet/grid_fluid_solver3.cpp:179-210): buoyancy, fractional pressure projection with MGPCG on the all-liquid domain,
 * cubic semi-Lagrangian advection (CubicSemiLagrangian3, the solver's default) of density / temperature / velocity,
 * decay.  No emitter, no collider, zero diffusion and viscosity coefficients (the defaults).  Fields stay in HBM between
 * calls.  [First CUDA implementation written against the pinned oracle after round 1's GPU minutes had run out; its
 * kernel bodies are checked on the CPU by tests/hostcheck, tests/firstrun/zz_smoke.py is the first GPU run.] */
typedef struct jetcu_smoke jetcu_smoke;
enum { JETCU_ADVECT_CUBIC = 0, JETCU_ADVECT_LINEAR = 1 };
typedef struct jetcu_smoke_params {
    jetcu_grid grid;
    int32_t closed_domain_boundary_flag;   /* GridFluidSolver3::closedDomainBoundaryFlag (default JETCU_DIR_ALL) */
    int32_t advection;                     /* GridFluidSolver3::advectionSolver(): JETCU_ADVECT_CUBIC = CubicSemiLagrangian3 (Jet's
                                              default), JETCU_ADVECT_LINEAR = SemiLagrangian3 (src/examples/smoke_sim example 2) */
    double gravity[3];                     /* default (0,-9.8,0): only its direction matters (buoyancy "up") */
    double max_cfl;                        /* default 5 */
    double buoyancy_density_factor;        /* GridSmokeSolver3::buoyancySmokeDensityFactor (default -0.000625) */
    double buoyancy_temperature_factor;    /* ::buoyancyTemperatureFactor (default 5.0) */
    double density_decay;                  /* ::smokeDecayFactor (default 0.001) */
    double temperature_decay;              /* ::temperatureDecayFactor (default 0.001) */
    jetcu_mgpcg_params mgpcg;
} jetcu_smoke_params;
void jetcu_smoke_default_params(jetcu_smoke_params* p);
int jetcu_smoke_create(jetcu_ctx*, const jetcu_smoke_params* params, jetcu_smoke** out);
int jetcu_smoke_destroy(jetcu_smoke*);
/* smokeDensity() / temperature() (cell-centred) and velocity() (face-centred); NULL = leave / skip */
int jetcu_smoke_set_fields(jetcu_smoke*, const double* density, const double* temperature, const double* u, const double* v,
                           const double* w);
int jetcu_smoke_get_fields(jetcu_smoke*, double* density, double* temperature, double* u, double* v, double* w, double* pressure);
/* VolumeGridEmitter3 with addStepFunctionTarget(smokeDensity, density_min, density_max) and
 * addStepFunctionTarget(temperature, temperature_min, temperature_max) (src/jet/volume_grid_emitter3.cpp:33-42, 83-99), fired
 * once on the device: value = max(old, (max - min) * (1 - smearedHeavisideSdf(sdf / min h)) + min) at every cell centre — the
 * smoke source of src/examples/smoke_sim/main.cpp:203-216.  Source region: shapes as for jetcu_hybrid_set_collider_shapes
 * (Builder::withSourceRegion wraps one explicit surface in SurfaceToImplicit3, a set stays an ImplicitSurfaceSet3).  A
 * continuous emitter (isOneShot = false) is this call before every step. */
int jetcu_smoke_emit_step(jetcu_smoke*, const jetcu_collider_shape* shapes, uint32_t n, int as_set, double density_min,
                          double density_max, double temperature_min, double temperature_max);
/* GridFluidSolver3::setCollider(RigidBodyCollider3 over analytic surfaces) on the smoke solver (src/examples/smoke_sim puts a
 * sphere in the plume): the collider branch of constrainVelocity, the collider SDF and velocity in the pressure system,
 * SemiLagrangian3's boundary handling (cells inside the collider are not advected, back-traces stop at the surface —
 * src/jet/semi_lagrangian3.cpp:36-40, 146-153) and extrapolateIntoCollider of density and temperature
 * (src/jet/grid_fluid_solver3.cpp:355-369).  Shapes as for jetcu_hybrid_set_collider_shapes; n == 0 removes the collider. */
int jetcu_smoke_set_collider_shapes(jetcu_smoke*, const jetcu_collider_shape* shapes, uint32_t n, int as_set,
                                    const double linear_velocity[3], const double angular_velocity[3], double friction);
/* GridFluidSolver3::cfl / numberOfSubTimeSteps (src/jet/grid_fluid_solver3.cpp:57-72, 212-217) evaluated on the device-resident
 * velocity, so that adaptive sub-stepping does not need the fields on the host. */
/* Device time (ms, CUDA events) of the stages of the last jetcu_smoke_step: buoyancy + boundary condition, pressure projection,
 * advection of all five fields, decay.  Any pointer may be NULL. */
int jetcu_smoke_last_step_ms(jetcu_smoke*, double* forces, double* pressure, double* advection, double* decay);
int jetcu_smoke_cfl(jetcu_smoke*, double dt, double* cfl);
int jetcu_smoke_num_sub_steps(jetcu_smoke*, double dt, uint32_t* n);
int jetcu_smoke_step(jetcu_smoke*, double dt, jetcu_mgpcg_result* result);                 /* onAdvanceTimeStep */
int jetcu_smoke_compute_external_forces(jetcu_smoke*, double dt);                          /* computeBuoyancyForce + BC */
int jetcu_smoke_compute_pressure(jetcu_smoke*, double dt, jetcu_mgpcg_result* result);     /* computePressure + BC */
int jetcu_smoke_compute_advection(jetcu_smoke*, double dt);                                /* computeAdvection */
int jetcu_smoke_end_advance_time_step(jetcu_smoke*, double dt);                            /* computeDiffusion (diffusion + decay) */
/* == GridFluidSolver3::setViscosityCoefficient, GridSmokeSolver3::setSmokeDiffusionCoefficient / setTemperatureDiffusionCoefficient.
 * Coefficients > kEpsilonD switch on computeViscosity in jetcu_smoke_step (src/jet/grid_fluid_solver3.cpp:231-241) and the scalar
 * solves of computeDiffusion in jetcu_smoke_end_advance_time_step (src/jet/grid_smoke_solver3.cpp:85-117), each through
 * GridBackwardEulerDiffusionSolver3 (Neumann, ICCG(100, kEpsilonD)) with the collider SDF as boundary.  Single GPU. */
int jetcu_smoke_set_diffusion(jetcu_smoke*, double viscosity, double smoke_diffusion, double temperature_diffusion);
int jetcu_smoke_compute_viscosity(jetcu_smoke*, double dt);

#ifdef __cplusplus
}
#endif
#endif /* JETCU_H_ */
