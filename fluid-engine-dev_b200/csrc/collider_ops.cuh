// collider_ops.cuh — per-cell / per-face bodies of the grid kernels around the collider and the boundary conditions,
// as __host__ __device__ functions: hybrid.cu's kernels are thin wrappers around them, and tests/hostcheck compiles the
// very same source for the CPU and checks it bit-for-bit against the oracle (no GPU needed).
#pragma once
#include "../../include/jetcu.h"
#include "collider.cuh"

struct GridD {
    int nx, ny, nz;
    double hx, hy, hz, ox, oy, oz;
    double ihx, ihy, ihz;  // exact 1/h for power-of-two spacings, else 0 (jc_norm)
};

// Component grid (AXIS = 0/1/2 -> u/v/w): size and data origin (src/jet/face_centered_grid3.cpp:430-444).
template <int AXIS>
__host__ __device__ __forceinline__ void comp_grid(const GridD& g, int* sx, int* sy, int* sz, double* ox, double* oy, double* oz) {
    *sx = g.nx + (AXIS == 0); *sy = g.ny + (AXIS == 1); *sz = g.nz + (AXIS == 2);
    *ox = g.ox + 0.5 * (AXIS == 0 ? 0.0 : g.hx);
    *oy = g.oy + 0.5 * (AXIS == 1 ? 0.0 : g.hy);
    *oz = g.oz + 0.5 * (AXIS == 2 ? 0.0 : g.hz);
}

__host__ __device__ __forceinline__ GridView comp_view(const GridD& g, int axis, const double* data) {
    GridView v;
    v.data = data;
    v.nx = g.nx + (axis == 0); v.ny = g.ny + (axis == 1); v.nz = g.nz + (axis == 2);
    v.hx = g.hx; v.hy = g.hy; v.hz = g.hz;
    v.ox = g.ox + 0.5 * (axis == 0 ? 0.0 : g.hx);
    v.oy = g.oy + 0.5 * (axis == 1 ? 0.0 : g.hy);
    v.oz = g.oz + 0.5 * (axis == 2 ? 0.0 : g.hz);
    v.ihx = g.ihx; v.ihy = g.ihy; v.ihz = g.ihz;
    return v;
}


// One cell of one sweep of extrapolateToRegion, include/jet/detail/array_utils-inl.h:191-239: reads only cells valid at
// sweep start, writes only invalid ones (so it is safe in place); neighbour order +i,-i,+j,-j,+k,-k; divide by count.
__host__ __device__ __forceinline__ void extrapolate_cell(double* data, const char* v0, char* v1, int nx, int ny, int nz, int i,
                                                          int j, int k) {
    const size_t sy = (size_t)nx, sz = (size_t)nx * ny;
    size_t c = (size_t)k * sz + (size_t)j * sy + i;
    if (v0[c]) { v1[c] = 1; return; }
    double sum = 0.0;
    unsigned count = 0;
    if (i + 1 < nx && v0[c + 1]) { sum += data[c + 1]; ++count; }
    if (i > 0 && v0[c - 1]) { sum += data[c - 1]; ++count; }
    if (j + 1 < ny && v0[c + sy]) { sum += data[c + sy]; ++count; }
    if (j > 0 && v0[c - sy]) { sum += data[c - sy]; ++count; }
    if (k + 1 < nz && v0[c + sz]) { sum += data[c + sz]; ++count; }
    if (k > 0 && v0[c - sz]) { sum += data[c - sz]; ++count; }
    if (count > 0) { data[c] = sum / (double)count; v1[c] = 1; }
    else v1[c] = 0;
}

// Closed-wall part of GridFractionalBoundaryConditionSolver3::constrainVelocity
// (src/jet/grid_fractional_boundary_condition_solver3.cpp:190-232), element t of the largest wall.
__host__ __device__ __forceinline__ void zero_walls_elem(const GridD& g, int flag, double* u, double* v, double* w, size_t t) {
    size_t nyz = (size_t)g.ny * g.nz, nxz = (size_t)g.nx * g.nz, nxy = (size_t)g.nx * g.ny;
    if (t < nyz) {
        int j = (int)(t % g.ny), k = (int)(t / g.ny);
        size_t rowu = ((size_t)k * g.ny + j) * (g.nx + 1);
        if (flag & JETCU_DIR_LEFT) u[rowu] = 0;
        if (flag & JETCU_DIR_RIGHT) u[rowu + g.nx] = 0;
    }
    if (t < nxz) {
        int i = (int)(t % g.nx), k = (int)(t / g.nx);
        size_t base = (size_t)k * (g.ny + 1) * g.nx + i;
        if (flag & JETCU_DIR_DOWN) v[base] = 0;
        if (flag & JETCU_DIR_UP) v[base + (size_t)g.ny * g.nx] = 0;
    }
    if (t < nxy) {
        if (flag & JETCU_DIR_BACK) w[t] = 0;
        if (flag & JETCU_DIR_FRONT) w[(size_t)g.nz * nxy + t] = 0;
    }
}

// GridFractionalBoundaryConditionSolver3::onColliderUpdated — src/jet/grid_fractional_boundary_condition_solver3.cpp:244-280:
// _colliderSdf(i,j,k) = SurfaceToImplicit3(surface).signedDistance(dataPosition(i,j,k)).
__host__ __device__ __forceinline__ double collider_sdf_cell(const GridD& g, const ColliderD& c, int i, int j, int k) {
    const double pt[3] = {(g.ox + 0.5 * g.hx) + g.hx * (double)i, (g.oy + 0.5 * g.hy) + g.hy * (double)j,
                          (g.oz + 0.5 * g.hz) + g.hz * (double)k};
    return cd_signed_distance(c, pt);
}

__host__ __device__ __forceinline__ GridView collider_view(const GridD& g, const double* sdf) {
    GridView v;
    v.data = sdf; v.nx = g.nx; v.ny = g.ny; v.nz = g.nz; v.hx = g.hx; v.hy = g.hy; v.hz = g.hz;
    v.ox = g.ox + 0.5 * g.hx; v.oy = g.oy + 0.5 * g.hy; v.oz = g.oz + 0.5 * g.hz;
    v.ihx = g.ihx; v.ihy = g.ihy; v.ihz = g.ihz;
    return v;
}

template <int AXIS>
__host__ __device__ __forceinline__ void face_position(const GridD& g, int i, int j, int k, double* pt) {
    const double fox = g.ox + 0.5 * (AXIS == 0 ? 0.0 : g.hx), foy = g.oy + 0.5 * (AXIS == 1 ? 0.0 : g.hy),
                 foz = g.oz + 0.5 * (AXIS == 2 ? 0.0 : g.hz);
    pt[0] = fox + g.hx * (double)i; pt[1] = foy + g.hy * (double)j; pt[2] = foz + g.hz * (double)k;
}

// constrainVelocity, first pass (…boundary_condition_solver3.cpp:55-104): faces whose open fraction is 0 take the
// collider's velocity and are marked invalid (0) for the extrapolation that follows.
template <int AXIS>
__host__ __device__ __forceinline__ void constrain_marker_face(const GridD& g, const ColliderD& c, const double* csdf,
                                                               double* data, char* marker, int i, int j, int k) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1);
    const GridView cs = collider_view(g, csdf);
    double pt[3];
    face_position<AXIS>(g, i, j, k, pt);
    const double hx = AXIS == 0 ? 0.5 * g.hx : 0.0, hy = AXIS == 1 ? 0.5 * g.hy : 0.0, hz = AXIS == 2 ? 0.5 * g.hz : 0.0;
    const double phi0 = jc_sample(cs, pt[0] - hx, pt[1] - hy, pt[2] - hz);
    const double phi1 = jc_sample(cs, pt[0] + hx, pt[1] + hy, pt[2] + hz);
    double frac = jc_frac_inside_sdf(phi0, phi1);
    frac = 1.0 - jc_clamp(frac, 0.0, 1.0);
    const size_t f = ((size_t)k * sy + j) * sx + i;
    if (frac > 0.0) {
        marker[f] = 1;
    } else {
        double cv[3];
        cd_velocity_at(c, pt, cv);
        data[f] = cv[AXIS];
        marker[f] = 0;
    }
}

// boundaryVelocity.sample(facePosition) component as the pressure solver reads it (…pressure_solver3.cpp:242-255):
// the collider's velocity field is collider->velocityAt (…boundary_condition_solver3.cpp:265-270).
template <int AXIS>
__host__ __device__ __forceinline__ void boundary_velocity_face(const GridD& g, const ColliderD& c, double* bnd, int i, int j, int k) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1);
    double pt[3], cv[3];
    face_position<AXIS>(g, i, j, k, pt);
    cd_velocity_at(c, pt, cv);
    bnd[((size_t)k * sy + j) * sx + i] = cv[AXIS];
}

// constrainVelocity, no-flux pass (…:115-187): inside the collider the (extrapolated) fluid velocity keeps only its
// tangential part relative to the collider, with friction; reads the whole field, writes the temporary.
template <int AXIS>
__host__ __device__ __forceinline__ void constrain_project_face(const GridD& g, const ColliderD& c, const double* csdf,
                                                                const double* u, const double* v, const double* w, double* temp,
                                                                int i, int j, int k) {
    const int sx = g.nx + (AXIS == 0), sy = g.ny + (AXIS == 1);
    const GridView cs = collider_view(g, csdf);
    double pt[3];
    face_position<AXIS>(g, i, j, k, pt);
    const size_t f = ((size_t)k * sy + j) * sx + i;
    const double* self = AXIS == 0 ? u : (AXIS == 1 ? v : w);
    if (jc_sample(cs, pt[0], pt[1], pt[2]) < 0) {  // isInsideSdf
        double cv[3], gr[3], vel[3];
        cd_velocity_at(c, pt, cv);
        vel[0] = jc_sample(comp_view(g, 0, u), pt[0], pt[1], pt[2]);
        vel[1] = jc_sample(comp_view(g, 1, v), pt[0], pt[1], pt[2]);
        vel[2] = jc_sample(comp_view(g, 2, w), pt[0], pt[1], pt[2]);
        cd_grid_gradient(cs, pt[0], pt[1], pt[2], gr);
        if (cd_len2(gr) > 0.0) {
            const double l = cd_len(gr);
            const double n[3] = {gr[0] / l, gr[1] / l, gr[2] / l};
            const double velr[3] = {vel[0] - cv[0], vel[1] - cv[1], vel[2] - cv[2]};
            double velt[3];
            cd_project_friction(velr, n, c.friction, velt);
            temp[f] = velt[AXIS] + cv[AXIS];
        } else {
            temp[f] = cv[AXIS];
        }
    } else {
        temp[f] = self[f];
    }
}

// GridFluidSolver3::extrapolateIntoCollider(ScalarGrid3*) marker — src/jet/grid_fluid_solver3.cpp:355-369
__host__ __device__ __forceinline__ char outside_collider_cell(const GridD& g, const double* csdf, int i, int j, int k) {
    const GridView cs = collider_view(g, csdf);
    const double px = (g.ox + 0.5 * g.hx) + g.hx * (double)i, py = (g.oy + 0.5 * g.hy) + g.hy * (double)j,
                 pz = (g.oz + 0.5 * g.hz) + g.hz * (double)k;
    return (jc_sample(cs, px, py, pz) < 0) ? 0 : 1;
}

// GridPointGenerator3::forEachPoint, point t of the (cx, cy, .) lattice in the generator's k, j, i order
// (src/jet/grid_point_generator3.cpp:24-33): position = index * spacing + lowerCorner
__host__ __device__ __forceinline__ void lattice_point(size_t t, unsigned cx, unsigned cy, double spacing, double lx, double ly,
                                                       double lz, double p[3]) {
    int i = (int)(t % cx), j = (int)((t / cx) % cy), k = (int)(t / ((size_t)cx * cy));
    p[0] = i * spacing + lx; p[1] = j * spacing + ly; p[2] = k * spacing + lz;
}

// The emitters' point sequences.  bcc = 0: GridPointGenerator3 (above); bcc = 1: BccLatticePointGenerator, VolumeParticleEmitter3's
// default (src/jet/bcc_lattice_point_generator.cpp:12-45): layers half a spacing apart in z, every other layer shifted by half a
// spacing in x and y, so even layers hold cx0 * cy0 points and odd layers cx1 * cy1.
struct LatticeD {
    int bcc;
    unsigned cx0, cy0, cx1, cy1, cz;
    double spacing, half, lx, ly, lz;
};
// counts with the generators' own loop tests (`i * spacing + offset <= width`), total number of points
inline size_t cd_lattice_make(const double lo[3], const double hi[3], double spacing, int bcc, LatticeD* L) {
    L->bcc = bcc; L->spacing = spacing; L->half = spacing / 2.0; L->lx = lo[0]; L->ly = lo[1]; L->lz = lo[2];
    const double w = hi[0] - lo[0], hgt = hi[1] - lo[1], d = hi[2] - lo[2];
    auto count = [](double step, double offset, double width) { unsigned c = 0; for (int i = 0; i * step + offset <= width; ++i) ++c; return c; };
    if (!bcc) {
        L->cx0 = L->cx1 = 0; L->cy0 = L->cy1 = 0; L->cz = 0;
        for (int i = 0; i * spacing <= w; ++i) ++L->cx0;      // literally GridPointGenerator3's tests (no "+ 0.0")
        for (int j = 0; j * spacing <= hgt; ++j) ++L->cy0;
        for (int k = 0; k * spacing <= d; ++k) ++L->cz;
        L->cx1 = L->cx0; L->cy1 = L->cy0;
        return (size_t)L->cx0 * L->cy0 * L->cz;
    }
    L->cx0 = count(spacing, 0.0, w); L->cx1 = count(spacing, L->half, w);
    L->cy0 = count(spacing, 0.0, hgt); L->cy1 = count(spacing, L->half, hgt);
    L->cz = 0;
    for (int k = 0; k * L->half <= d; ++k) ++L->cz;
    const size_t n0 = (size_t)L->cx0 * L->cy0, n1 = (size_t)L->cx1 * L->cy1;
    return (size_t)(L->cz / 2) * (n0 + n1) + (size_t)(L->cz & 1u) * n0;
}
__host__ __device__ __forceinline__ void lattice_point_l(const LatticeD& L, size_t t, double p[3]) {
    if (!L.bcc) { lattice_point(t, L.cx0, L.cy0, L.spacing, L.lx, L.ly, L.lz, p); return; }
    const size_t n0 = (size_t)L.cx0 * L.cy0, n1 = (size_t)L.cx1 * L.cy1, pair = n0 + n1;
    const size_t q = t / pair, r = t % pair;
    const bool odd = r >= n0;
    const size_t idx = odd ? r - n0 : r;
    const unsigned cx = odd ? L.cx1 : L.cx0;
    const double offset = odd ? L.half : 0.0;
    const int i = (int)(idx % cx), j = (int)(idx / cx), k = (int)(2 * q + (odd ? 1 : 0));
    p[0] = i * L.spacing + offset + L.lx; p[1] = j * L.spacing + offset + L.ly; p[2] = k * L.half + L.lz;
}

