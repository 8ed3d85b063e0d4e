// collider.cuh — analytic rigid collider of the hybrid solver on the device (SURVEY.md §8a-a9 with a collider, §8f-1):
// RigidBodyCollider3 (src/jet/rigid_body_collider3.cpp) over a Box3 (src/jet/box3.cpp) or Sphere3 (src/jet/sphere3.cpp)
// with identity transforms.  Closed forms replace the reference's per-point virtual Surface3 calls; every expression
// keeps the reference's operand order (Vector3: include/jet/detail/vector3-inl.h) and the TU is built with -fmad=false,
// so the collider SDF, the constrained velocities and the resolved particles are meant to be BIT-IDENTICAL to the CPU
// path (checked against oracle/jet_port_collider.inc, which is pinned to the unmodified reference).
#pragma once
#include "sampling.cuh"

struct ColliderD {
    int kind = 0;        // 0 none, 1 Box3 (a = lowerCorner, b = upperCorner), 2 Sphere3 (a = center, b[0] = radius)
    double a[3] = {}, b[3] = {};
    double lin[3] = {}, ang[3] = {};   // RigidBodyCollider3::linearVelocity / angularVelocity
    double friction = 0.0;             // Collider3::frictionCoefficient()
};

__host__ __device__ __forceinline__ double cd_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__host__ __device__ __forceinline__ double cd_len2(const double* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
__host__ __device__ __forceinline__ double cd_len(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
__host__ __device__ __forceinline__ double cd_max(double a, double b) { return (a < b) ? b : a; }  // std::max

// Transform3 with identity orientation and zero translation, spelt out (include/jet/detail/transform3-inl.h:45-47,
// 72-79 over Matrix3x3::mul, matrix3x3-inl.h:208-213): exact no-ops except that -0.0 becomes +0.0.
__host__ __device__ __forceinline__ void cd_to_local(const double* p, double* out) {
    const double t0 = p[0] - 0.0, t1 = p[1] - 0.0, t2 = p[2] - 0.0;
    out[0] = 1.0 * t0 + 0.0 * t1 + 0.0 * t2;
    out[1] = 0.0 * t0 + 1.0 * t1 + 0.0 * t2;
    out[2] = 0.0 * t0 + 0.0 * t1 + 1.0 * t2;
}
__host__ __device__ __forceinline__ void cd_to_world_dir(const double* d, double* out) {
    out[0] = 1.0 * d[0] + 0.0 * d[1] + 0.0 * d[2];
    out[1] = 0.0 * d[0] + 1.0 * d[1] + 0.0 * d[2];
    out[2] = 0.0 * d[0] + 0.0 * d[1] + 1.0 * d[2];
}
__host__ __device__ __forceinline__ void cd_to_world(const double* p, double* out) {
    double t[3];
    cd_to_world_dir(p, t);
    out[0] = t[0] + 0.0; out[1] = t[1] + 0.0; out[2] = t[2] + 0.0;
}

// normal of box plane q: +x, +y, +z (through upperCorner), -x, -y, -z (through lowerCorner) — src/jet/box3.cpp:30-35
__host__ __device__ __forceinline__ void cd_box_normal(int q, double* n) {
    n[0] = n[1] = n[2] = 0.0;
    n[q % 3] = q < 3 ? 1.0 : -1.0;
}

// Plane3::closestPointLocal — src/jet/plane3.cpp:35-38: r = p - point; r - normal.dot(r) * normal + point
__host__ __device__ __forceinline__ void cd_plane_closest(const double* n, const double* pt, const double* p, double* out) {
    const double r[3] = {p[0] - pt[0], p[1] - pt[1], p[2] - pt[2]};
    const double d = cd_dot(n, r);
#pragma unroll
    for (int a = 0; a < 3; ++a) out[a] = (r[a] - n[a] * d) + pt[a];
}

__host__ __device__ __forceinline__ bool cd_box_contains(const ColliderD& c, const double* p) {  // bounding_box3-inl.h:77-91
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (c.b[a] < p[a] || c.a[a] > p[a]) return false;
    return true;
}

// Box3::closestPointLocal (src/jet/box3.cpp:28-55) and closestNormalLocal (:57-100) in one pass: both scan the six
// planes with the same strict `<` on the squared distance, so they pick the same plane.
__host__ __device__ __forceinline__ void cd_box_closest_local(const ColliderD& c, const double* p, double* point, double* normal) {
    if (cd_box_contains(c, p)) {
        double n[3], best[3];
        cd_box_normal(0, n);
        cd_plane_closest(n, c.b, p, best);
        double d[3] = {best[0] - p[0], best[1] - p[1], best[2] - p[2]};
        double best2 = cd_len2(d);
        int bq = 0;
#pragma unroll
        for (int q = 1; q < 6; ++q) {
            double cand[3];
            cd_box_normal(q, n);
            cd_plane_closest(n, q < 3 ? c.b : c.a, p, cand);
            const double e[3] = {cand[0] - p[0], cand[1] - p[1], cand[2] - p[2]};
            const double e2 = cd_len2(e);
            if (e2 < best2) { best[0] = cand[0]; best[1] = cand[1]; best[2] = cand[2]; best2 = e2; bq = q; }
        }
        point[0] = best[0]; point[1] = best[1]; point[2] = best[2];
        cd_box_normal(bq, normal);
    } else {
        double r[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) { point[a] = jc_clamp(p[a], c.a[a], c.b[a]); r[a] = p[a] - point[a]; }
        double n[3];
        cd_box_normal(0, n);
        double bestcos = cd_dot(n, r);
        int bq = 0;
#pragma unroll
        for (int q = 1; q < 6; ++q) {
            cd_box_normal(q, n);
            const double cs = cd_dot(n, r);
            if (cs > bestcos) { bq = q; bestcos = cs; }
        }
        cd_box_normal(bq, normal);
    }
}

// Sphere3::closestNormalLocal / closestPointLocal — src/jet/sphere3.cpp:25-39
__host__ __device__ __forceinline__ void cd_sphere_closest_local(const ColliderD& c, const double* p, double* point, double* normal) {
    const double eps = 2.220446049250313e-16;  // Vector3::isSimilar's default epsilon
    if (fabs(c.a[0] - p[0]) < eps && fabs(c.a[1] - p[1]) < eps && fabs(c.a[2] - p[2]) < eps) {
        normal[0] = 1.0; normal[1] = 0.0; normal[2] = 0.0;
    } else {
        const double r[3] = {p[0] - c.a[0], p[1] - c.a[1], p[2] - c.a[2]};
        const double l = cd_len(r);
        normal[0] = r[0] / l; normal[1] = r[1] / l; normal[2] = r[2] / l;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) point[a] = normal[a] * c.b[0] + c.a[a];  // radius * normal + center
}

__host__ __device__ __forceinline__ void cd_shape_closest_local(const ColliderD& c, const double* p, double* point, double* normal) {
    if (c.kind == 1) cd_box_closest_local(c, p, point, normal); else cd_sphere_closest_local(c, p, point, normal);
}

// One query of the surface at world point p — Surface3::closestPoint / closestNormal / closestDistance / isInside
// (src/jet/surface3.cpp:22-79, isNormalFlipped = false).
struct ColliderQuery {
    double point[3], normal[3], distance;
    bool inside;
};

__host__ __device__ __forceinline__ ColliderQuery cd_query(const ColliderD& c, const double* p) {
    ColliderQuery q;
    double l[3], cp[3], n[3];
    cd_to_local(p, l);
    cd_shape_closest_local(c, l, cp, n);
    cd_to_world(cp, q.point);
    cd_to_world_dir(n, q.normal);
    q.normal[0] *= 1.0; q.normal[1] *= 1.0; q.normal[2] *= 1.0;  // result *= isNormalFlipped ? -1 : 1
    const double r[3] = {l[0] - cp[0], l[1] - cp[1], l[2] - cp[2]};
    q.inside = cd_dot(r, n) < 0.0;  // Surface3::isInsideLocal
    if (c.kind == 2) {              // Sphere3::closestDistanceLocal: |center.distanceTo(p) - radius|
        const double s[3] = {c.a[0] - l[0], c.a[1] - l[1], c.a[2] - l[2]};
        q.distance = fabs(cd_len(s) - c.b[0]);
    } else {                        // Surface3::closestDistanceLocal: p.distanceTo(closestPointLocal(p))
        q.distance = cd_len(r);
    }
    return q;
}

// SurfaceToImplicit3(surface)->signedDistance(pt) — implicit_surface3.cpp:26-29, surface_to_implicit3.cpp:67-72
__host__ __device__ __forceinline__ double cd_signed_distance(const ColliderD& c, const double* pt) {
    double l[3];
    cd_to_local(pt, l);  // the wrapper's own (identity) transform
    const ColliderQuery q = cd_query(c, l);
    const double r[3] = {q.point[0] - l[0], q.point[1] - l[1], q.point[2] - l[2]};
    const double d = cd_len(r);  // x.distanceTo(otherPoint)
    return q.inside ? -d : d;
}

// RigidBodyCollider3::velocityAt — src/jet/rigid_body_collider3.cpp:25-28: linear + angular x (p - translation)
__host__ __device__ __forceinline__ void cd_velocity_at(const ColliderD& c, const double* p, double* out) {
    const double r[3] = {p[0] - 0.0, p[1] - 0.0, p[2] - 0.0};
    const double cr[3] = {c.ang[1] * r[2] - r[1] * c.ang[2], c.ang[2] * r[0] - r[2] * c.ang[0], c.ang[0] * r[1] - r[0] * c.ang[1]};
    out[0] = c.lin[0] + cr[0]; out[1] = c.lin[1] + cr[1]; out[2] = c.lin[2] + cr[2];
}

// projectAndApplyFriction — src/jet/physics_helpers.h:33-45
__host__ __device__ __forceinline__ void cd_project_friction(const double* vel, const double* n, double friction, double* out) {
    const double d = cd_dot(vel, n);
    double velt[3] = {vel[0] - n[0] * d, vel[1] - n[1] * d, vel[2] - n[2] * d};  // vel.projected(normal)
    if (cd_len2(velt) > 0) {
        const double veln = cd_max(-cd_dot(vel, n), 0.0);
        const double s = cd_max(1.0 - friction * veln / cd_len(velt), 0.0);
        velt[0] *= s; velt[1] *= s; velt[2] *= s;
    }
    out[0] = velt[0]; out[1] = velt[1]; out[2] = velt[2];
}

// ScalarGrid3::gradient(x) — src/jet/scalar_grid3.cpp:89-102: sum over the 8 corners of getCoordinatesAndWeights of
// weight * gradient3 (src/jet/fdm_utils.cpp:49-67: 0.5 * (right-left, up-down, front-back) / gridSpacing, clamped).
__host__ __device__ __forceinline__ void cd_grid_gradient(const GridView& g, double px, double py, double pz, double* out) {
    const BaseFrac bf = jc_base_mul(g.nx, g.ny, g.nz, 1.0 / g.hx, 1.0 / g.hy, 1.0 / g.hz, g.ox, g.oy, g.oz, px, py, pz);
    const size_t sy = (size_t)g.nx, sz = (size_t)g.nx * g.ny;
    out[0] = out[1] = out[2] = 0.0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int i = (q & 1) ? bf.ip1 : bf.i, j = (q & 2) ? bf.jp1 : bf.j, k = (q & 4) ? bf.kp1 : bf.k;
        const double w = jc_corner_weight(bf, q);
        const double* d = g.data;
        const double left = d[k * sz + j * sy + (i > 0 ? i - 1 : i)], right = d[k * sz + j * sy + (i + 1 < g.nx ? i + 1 : i)];
        const double down = d[k * sz + (j > 0 ? j - 1 : j) * sy + i], up = d[k * sz + (j + 1 < g.ny ? j + 1 : j) * sy + i];
        const double back = d[(k > 0 ? k - 1 : k) * sz + j * sy + i], front = d[(k + 1 < g.nz ? k + 1 : k) * sz + j * sy + i];
        const double gx = ((right - left) * 0.5) / g.hx, gy = ((up - down) * 0.5) / g.hy, gz = ((front - back) * 0.5) / g.hz;
        out[0] += gx * w; out[1] += gy * w; out[2] += gz * w;
    }
}

// Collider3::resolveCollision(radius = 0, restitution = 0, &pos, &vel) — src/jet/collider3.cpp:19-74
__host__ __device__ __forceinline__ void cd_resolve_collision(const ColliderD& c, double* pos, double* vel) {
    const double radius = 0.0, restitution = 0.0;
    const ColliderQuery q = cd_query(c, pos);
    double cvel[3];
    cd_velocity_at(c, pos, cvel);
    if (q.inside || q.distance < radius) {
        const double target[3] = {q.point[0] + q.normal[0] * radius, q.point[1] + q.normal[1] * radius, q.point[2] + q.normal[2] * radius};
        const double rel[3] = {vel[0] - cvel[0], vel[1] - cvel[1], vel[2] - cvel[2]};
        const double ndot = cd_dot(q.normal, rel);
        double reln[3] = {q.normal[0] * ndot, q.normal[1] * ndot, q.normal[2] * ndot};
        double relt[3] = {rel[0] - reln[0], rel[1] - reln[1], rel[2] - reln[2]};
        if (ndot < 0.0) {
            const double f = -restitution - 1.0;
            const double dn[3] = {reln[0] * f, reln[1] * f, reln[2] * f};
            reln[0] *= -restitution; reln[1] *= -restitution; reln[2] *= -restitution;
            if (cd_len2(relt) > 0.0) {
                const double scale = cd_max(1.0 - c.friction * cd_len(dn) / cd_len(relt), 0.0);
                relt[0] *= scale; relt[1] *= scale; relt[2] *= scale;
            }
            vel[0] = (reln[0] + relt[0]) + cvel[0]; vel[1] = (reln[1] + relt[1]) + cvel[1]; vel[2] = (reln[2] + relt[2]) + cvel[2];
        }
        pos[0] = target[0]; pos[1] = target[1]; pos[2] = target[2];
    }
}
