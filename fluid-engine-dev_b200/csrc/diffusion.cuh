// diffusion.cuh — GridBackwardEulerDiffusionSolver3 on the device (diffusion.cu).
#pragma once
#include "fdm.cuh"
#include "collider_ops.cuh"

// kind 0 = cell-centred scalar grid, 1/2/3 = u/v/w array of a face-centred grid; all pointers are device memory;
// csdf / fsdf = cell-centred collider / fluid SDF or nullptr (no collider / everything fluid)
int diffusion_solve_array(jetcu_ctx* ctx, const GridD& g, int kind, const double* f, double coefficient, double dt,
                          int boundary_type, const double* csdf, const double* fsdf, uint32_t max_iterations, double tolerance,
                          double* dest, char* markers_out, jetcu_mgpcg_result* result);
int diffusion_solve_faces(jetcu_ctx* ctx, const GridD& g, const double* u, const double* v, const double* w, double coefficient,
                          double dt, int boundary_type, const double* csdf, const double* fsdf, uint32_t max_iterations,
                          double tolerance, double* ou, double* ov, double* ow, jetcu_mgpcg_result results[3]);
