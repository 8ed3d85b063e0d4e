// pressure.cuh — device-side GridFractionalSinglePhasePressureSolver3 (weights, 7-point system, gradient).
#pragma once
#include "fdm.cuh"

struct PressureSystem {
    jetcu_grid grid{};
    int levels = 0;
    Dim3i dims[JC_MAX_LEVELS];
    float* sdf[JC_MAX_LEVELS] = {};  // _fluidSdf[l]
    float* uw[JC_MAX_LEVELS] = {};   // _uWeights[l]  (nx+1, ny, nz)
    float* vw[JC_MAX_LEVELS] = {};
    float* ww[JC_MAX_LEVELS] = {};
    double* center[JC_MAX_LEVELS] = {};  // packed matrix: diagonal ...
    float* codes[JC_MAX_LEVELS] = {};    // ... and the three +face weights (3 planes of n floats)
    uint8_t* code8[JC_MAX_LEVELS] = {};  // ... and their 2-bit classes, one byte per cell
    MgLevels mg;                     // A[l] (views of center/codes), x/b/buf per level
    double* b0 = nullptr;            // finest-level right-hand side
    double* x0 = nullptr;            // finest-level solution == pressure()
    uint32_t built_max_levels = 0;
};

void pressure_free(jetcu_ctx* ctx, PressureSystem* ps);
// (Re)allocates ps for `grid` and params->max_levels if its shape changed.
int pressure_prepare(jetcu_ctx* ctx, PressureSystem* ps, const jetcu_grid* grid, uint32_t max_levels);
// All pointers are DEVICE pointers; collider_sdf / bnd_* / fluid_sdf may be null (defaults as in jetcu.h).
// u_out may alias u_in (etc).
int pressure_solve_device(jetcu_ctx* ctx, PressureSystem* ps, const double* u_in, const double* v_in,
                          const double* w_in, const double* collider_sdf, const double* bnd_u, const double* bnd_v,
                          const double* bnd_w, const double* fluid_sdf, const jetcu_mgpcg_params* params, double* u_out,
                          double* v_out, double* w_out, jetcu_mgpcg_result* result);
