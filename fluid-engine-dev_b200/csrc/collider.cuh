// collider.cuh — analytic rigid collider of the hybrid solver on the device (SURVEY.md §8a-a9 with a collider, §8f-1):
// RigidBodyCollider3 (src/jet/rigid_body_collider3.cpp) over ONE explicit surface — Box3 (src/jet/box3.cpp), Sphere3
// (sphere3.cpp), Cylinder3 (cylinder3.cpp), Plane3 (plane3.cpp), optionally normal-flipped — or over an
// ImplicitSurfaceSet3 (implicit_surface_set3.cpp) of up to 8 of them, with identity transforms.  Closed forms replace the
// reference's per-point virtual Surface3 calls; the set's Bvh3 (include/jet/detail/bvh3-inl.h) is built on the host with
// the reference's split rule and walked on the device in the reference's order, so that ties between members resolve the
// same way.  Every expression keeps the reference's operand order (Vector3: include/jet/detail/vector3-inl.h) and the
// TU is built with -fmad=false, so the collider SDF, the constrained velocities and the resolved particles are
// BIT-IDENTICAL to the CPU path for boxes, spheres and planes; Cylinder3 calls atan2 / cos / sin, where the device's libm
// may differ from glibc in the last bit (checked against oracle/jet_port_collider.inc, which is pinned to the
// unmodified reference; on the CPU — tests/hostcheck — every kind is bit-identical).
#pragma once
#include "sampling.cuh"

#define JC_COLLIDER_MAX_SHAPES 8
#define JC_SHAPE_BOX 1       // a = lowerCorner, b = upperCorner
#define JC_SHAPE_SPHERE 2    // a = center, b[0] = radius
#define JC_SHAPE_CYLINDER 3  // a = center, b[0] = radius, b[1] = height (axis y)
#define JC_SHAPE_PLANE 4     // a = point, b = normal

struct ShapeD {
    int kind = 0, flipped = 0;  // flipped: Surface3::isNormalFlipped
    double a[3] = {}, b[3] = {};
};

struct BvhNodeD {  // Bvh3<T>::Node — bvh3-inl.h:17-45
    double lo[3], hi[3];
    int leaf, item, child, pad;
};

struct ColliderD {
    int kind = 0;            // 0: no collider
    int n = 0, is_set = 0;   // is_set: the surface is ImplicitSurfaceSet3(s[0..n)); else it is s[0] itself
    int nnodes = 0, nitems = 0, nunb = 0;
    ShapeD s[JC_COLLIDER_MAX_SHAPES];
    BvhNodeD nodes[2 * JC_COLLIDER_MAX_SHAPES];
    int items[JC_COLLIDER_MAX_SHAPES] = {};  // Bvh3::_items: the bounded members in insertion order
    int unb[JC_COLLIDER_MAX_SHAPES] = {};    // ImplicitSurfaceSet3::_unboundedSurfaces
    double lin[3] = {}, ang[3] = {};   // RigidBodyCollider3::linearVelocity / angularVelocity
    double friction = 0.0;             // Collider3::frictionCoefficient()
};

__host__ __device__ __forceinline__ double cd_dot(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__host__ __device__ __forceinline__ double cd_len2(const double* a) { return a[0] * a[0] + a[1] * a[1] + a[2] * a[2]; }
__host__ __device__ __forceinline__ double cd_len(const double* a) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
__host__ __device__ __forceinline__ double cd_max(double a, double b) { return (a < b) ? b : a; }  // std::max
__host__ __device__ __forceinline__ double cd_min(double a, double b) { return (b < a) ? b : a; }  // std::min
__host__ __device__ __forceinline__ double cd_dist(const double* a, const double* b) {            // a.distanceTo(b)
    const double r[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    return cd_len(r);
}

// Transform3 with identity orientation and zero translation, spelt out (include/jet/detail/transform3-inl.h:45-47,
// 72-79 over Matrix3x3::mul, matrix3x3-inl.h:208-213): exact no-ops except that -0.0 may become +0.0.
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
// Transform2 with orientation 0 (cos 1, sin 0) and zero translation — include/jet/detail/transform2-inl.h:48-54, 85-99
__host__ __device__ __forceinline__ void cd2_to_local(const double* p, double* out) {
    const double cs = 1.0, sn = 0.0;
    const double x = p[0] - 0.0, y = p[1] - 0.0;
    out[0] = cs * x + sn * y;
    out[1] = -sn * x + cs * y;
}
__host__ __device__ __forceinline__ void cd2_to_world(const double* p, double* out) {
    const double cs = 1.0, sn = 0.0;
    out[0] = cs * p[0] - sn * p[1] + 0.0;
    out[1] = sn * p[0] + cs * p[1] + 0.0;
}
__host__ __device__ __forceinline__ void cd2_to_world_dir(const double* d, double* out) {
    const double cs = 1.0, sn = 0.0;
    out[0] = cs * d[0] - sn * d[1];
    out[1] = sn * d[0] + cs * d[1];
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

__host__ __device__ __forceinline__ bool cd_box_contains(const ShapeD& c, const double* p) {  // bounding_box3-inl.h:77-91
#pragma unroll
    for (int a = 0; a < 3; ++a)
        if (c.b[a] < p[a] || c.a[a] > p[a]) return false;
    return true;
}

// Box3::closestPointLocal (src/jet/box3.cpp:28-55) and closestNormalLocal (:57-100) in one pass: both scan the six
// planes with the same strict `<` on the squared distance, so they pick the same plane.
__host__ __device__ __forceinline__ void cd_box_closest_local(const ShapeD& c, const double* p, double* point, double* normal) {
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
__host__ __device__ __forceinline__ void cd_sphere_closest_local(const ShapeD& c, const double* p, double* point, double* normal) {
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

// ---- Box2 (src/jet/box2.cpp:28-96) over Plane2 (plane2.cpp:28-36) behind Surface2's public functions
// (surface2.cpp:23-79): the 2-D profile Cylinder3 queries.  lo / hi = bound.lowerCorner / upperCorner. -----------------
__host__ __device__ __forceinline__ void cd_box2_normal(int q, double* n) {  // (1,0) (0,1) through hi, (-1,0) (0,-1) through lo
    n[0] = n[1] = 0.0;
    n[q & 1] = q < 2 ? 1.0 : -1.0;
}
__host__ __device__ __forceinline__ void cd_plane2_closest(const double* n, const double* pt, const double* p, double* out) {
    double l[2], q[2];
    cd2_to_local(p, l);
    const double r[2] = {l[0] - pt[0], l[1] - pt[1]};
    const double d = n[0] * r[0] + n[1] * r[1];
    q[0] = (r[0] - n[0] * d) + pt[0]; q[1] = (r[1] - n[1] * d) + pt[1];
    cd2_to_world(q, out);
}
__host__ __device__ __forceinline__ bool cd_box2_contains(const double* lo, const double* hi, const double* p) {
    if (hi[0] < p[0] || lo[0] > p[0]) return false;
    if (hi[1] < p[1] || lo[1] > p[1]) return false;
    return true;
}
__host__ __device__ __forceinline__ double cd2_dist2(const double* a, const double* b) {
    const double x = a[0] - b[0], y = a[1] - b[1];
    return x * x + y * y;
}
// Box2::closestPointLocal and closestNormalLocal in one pass (same plane scan, same strict comparisons)
__host__ __device__ __forceinline__ void cd_box2_closest_local(const double* lo, const double* hi, const double* p, double* point,
                                                               double* normal) {
    if (cd_box2_contains(lo, hi, p)) {
        double n[2], best[2];
        cd_box2_normal(0, n);
        cd_plane2_closest(n, hi, p, best);
        double best2 = cd2_dist2(best, p);
        int bq = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            double cand[2];
            cd_box2_normal(q, n);
            cd_plane2_closest(n, q < 2 ? hi : lo, p, cand);
            const double e2 = cd2_dist2(cand, p);
            if (e2 < best2) { best[0] = cand[0]; best[1] = cand[1]; best2 = e2; bq = q; }
        }
        point[0] = best[0]; point[1] = best[1];
        cd_box2_normal(bq, normal);
    } else {
        point[0] = jc_clamp(p[0], lo[0], hi[0]); point[1] = jc_clamp(p[1], lo[1], hi[1]);
        const double r[2] = {p[0] - point[0], p[1] - point[1]};
        double n[2];
        cd_box2_normal(0, n);
        double bestcos = n[0] * r[0] + n[1] * r[1];
        int bq = 0;
#pragma unroll
        for (int q = 1; q < 4; ++q) {
            cd_box2_normal(q, n);
            const double cs = n[0] * r[0] + n[1] * r[1];
            if (cs > bestcos) { bq = q; bestcos = cs; }
        }
        cd_box2_normal(bq, normal);
    }
}

// ---- Cylinder3 — src/jet/cylinder3.cpp:32-68 ------------------------------------------------------------------------
struct CylProfile {
    double r[3], rr[2], lo[2], hi[2];
};
__host__ __device__ __forceinline__ CylProfile cd_cyl_profile(const ShapeD& c, const double* p) {
    CylProfile f;
    f.r[0] = p[0] - c.a[0]; f.r[1] = p[1] - c.a[1]; f.r[2] = p[2] - c.a[2];
    f.rr[0] = sqrt(f.r[0] * f.r[0] + f.r[2] * f.r[2]); f.rr[1] = f.r[1];
    // Box2(Vector2D(-radius, -0.5 * height), Vector2D(radius, 0.5 * height)) -> BoundingBox2D(point1, point2): min / max
    const double p1[2] = {-c.b[0], -0.5 * c.b[1]}, p2[2] = {c.b[0], 0.5 * c.b[1]};
    f.lo[0] = cd_min(p1[0], p2[0]); f.lo[1] = cd_min(p1[1], p2[1]);
    f.hi[0] = cd_max(p1[0], p2[0]); f.hi[1] = cd_max(p1[1], p2[1]);
    return f;
}
// closestPointLocal (:32-42) and closestNormalLocal (:52-68): box.closestPoint(rr) / box.closestNormal(rr) are Surface2's
// public functions (identity Transform2 around the *Local ones)
__host__ __device__ __forceinline__ void cd_cyl_closest_local(const ShapeD& c, const double* p, double* point, double* normal) {
    const CylProfile f = cd_cyl_profile(c, p);
    double l[2], q[2], n2[2], cp[2], cn[2];
    cd2_to_local(f.rr, l);
    cd_box2_closest_local(f.lo, f.hi, l, q, n2);
    cd2_to_world(q, cp);
    cd2_to_world_dir(n2, cn);
    cn[0] *= 1.0; cn[1] *= 1.0;
    const double angle = atan2(f.r[2], f.r[0]);
    point[0] = cp[0] * cos(angle) + c.a[0]; point[1] = cp[1] + c.a[1]; point[2] = cp[0] * sin(angle) + c.a[2];
    if (cn[1] > 0) { normal[0] = 0.0; normal[1] = 1.0; normal[2] = 0.0; }
    else if (cn[1] < 0) { normal[0] = 0.0; normal[1] = -1.0; normal[2] = 0.0; }
    else {  // Vector3D(r.x, 0, r.z).normalized()
        const double v[3] = {f.r[0], 0.0, f.r[2]};
        const double len = cd_len(v);
        normal[0] = v[0] / len; normal[1] = v[1] / len; normal[2] = v[2] / len;
    }
}
// closestDistanceLocal (:44-50): box.closestDistance(rr) = Surface2::closestDistanceLocal on toLocal(rr)
__host__ __device__ __forceinline__ double cd_cyl_closest_distance_local(const ShapeD& c, const double* p) {
    const CylProfile f = cd_cyl_profile(c, p);
    double l[2], q[2], n2[2];
    cd2_to_local(f.rr, l);
    cd_box2_closest_local(f.lo, f.hi, l, q, n2);
    return sqrt(cd2_dist2(l, q));
}

// ---- one explicit surface: closestPointLocal + closestNormalLocal by kind, then Surface3's public functions
// (src/jet/surface3.cpp:22-79) in one query ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ void cd_shape_closest_local(const ShapeD& c, const double* p, double* point, double* normal) {
    if (c.kind == JC_SHAPE_BOX) cd_box_closest_local(c, p, point, normal);
    else if (c.kind == JC_SHAPE_SPHERE) cd_sphere_closest_local(c, p, point, normal);
    else if (c.kind == JC_SHAPE_CYLINDER) cd_cyl_closest_local(c, p, point, normal);
    else {  // Plane3: plane3.cpp:35-43
        cd_plane_closest(c.b, c.a, p, point);
        normal[0] = c.b[0]; normal[1] = c.b[1]; normal[2] = c.b[2];
    }
}

struct ColliderQuery {
    double point[3], normal[3], distance;
    bool inside;
};

// closestPoint / closestNormal / closestDistance / isInside of shape `c` at world point p
__host__ __device__ __forceinline__ ColliderQuery cd_surf_query(const ShapeD& c, const double* p) {
    ColliderQuery q;
    double l[3], cp[3], n[3];
    cd_to_local(p, l);
    cd_shape_closest_local(c, l, cp, n);
    cd_to_world(cp, q.point);
    cd_to_world_dir(n, q.normal);
    const double f = c.flipped ? -1.0 : 1.0;  // result *= isNormalFlipped ? -1 : 1
    q.normal[0] *= f; q.normal[1] *= f; q.normal[2] *= f;
    const double r[3] = {l[0] - cp[0], l[1] - cp[1], l[2] - cp[2]};
    const bool in = cd_dot(r, n) < 0.0;       // Surface3::isInsideLocal
    q.inside = (c.flipped != 0) == !in;       // isNormalFlipped == !isInsideLocal(...)
    if (c.kind == JC_SHAPE_SPHERE) q.distance = fabs(cd_dist(c.a, l) - c.b[0]);   // Sphere3::closestDistanceLocal
    else if (c.kind == JC_SHAPE_CYLINDER) q.distance = cd_cyl_closest_distance_local(c, l);
    else q.distance = cd_len(r);              // Surface3::closestDistanceLocal: p.distanceTo(closestPointLocal(p))
    return q;
}
__host__ __device__ __forceinline__ double cd_surf_closest_distance(const ShapeD& c, const double* p) {
    double l[3];
    cd_to_local(p, l);
    if (c.kind == JC_SHAPE_SPHERE) return fabs(cd_dist(c.a, l) - c.b[0]);
    if (c.kind == JC_SHAPE_CYLINDER) return cd_cyl_closest_distance_local(c, l);
    double cp[3], n[3];
    cd_shape_closest_local(c, l, cp, n);
    return cd_dist(l, cp);
}

// ---- SurfaceToImplicit3(shape) with its own identity transform (surface_to_implicit3.cpp:41-76) ----------------------
__host__ __device__ __forceinline__ ColliderQuery cd_wrap_query(const ShapeD& c, const double* p) {
    double l[3];
    cd_to_local(p, l);
    const ColliderQuery s = cd_surf_query(c, l);
    ColliderQuery q;
    cd_to_world(s.point, q.point);
    cd_to_world_dir(s.normal, q.normal);
    q.normal[0] *= 1.0; q.normal[1] *= 1.0; q.normal[2] *= 1.0;
    q.distance = s.distance;
    q.inside = s.inside;
    return q;
}
__host__ __device__ __forceinline__ double cd_wrap_closest_distance(const ShapeD& c, const double* p) {
    double l[3];
    cd_to_local(p, l);
    return cd_surf_closest_distance(c, l);
}
// ImplicitSurface3::signedDistance (implicit_surface3.cpp:26-29) over SurfaceToImplicit3::signedDistanceLocal (:67-72)
__host__ __device__ __forceinline__ double cd_wrap_signed_distance(const ShapeD& c, const double* pt) {
    double l[3];
    cd_to_local(pt, l);  // the wrapper's own (identity) transform
    const ColliderQuery q = cd_surf_query(c, l);
    const double d = cd_dist(q.point, l);  // x.distanceTo(otherPoint)
    return q.inside ? -d : d;
}

// ---- Bvh3::nearest (bvh3-inl.h:80-165) with distanceFunc = member->closestDistance(pt) ----------------------------------
struct BvhNearest {
    int item;  // index into ColliderD::s, -1 = nullptr
    double distance;
};
__host__ __device__ __forceinline__ BvhNearest cd_bvh_nearest(const ColliderD& c, const double* pt) {
    BvhNearest best;
    best.item = -1;
    best.distance = DBL_MAX;
    if (c.nnodes == 0) return best;
    int todo[JC_COLLIDER_MAX_SHAPES], todoPos = 0, node = 0;
    for (;;) {
        const BvhNodeD& nd = c.nodes[node];
        if (nd.leaf) {
            const int m = c.items[nd.item];
            const double dist = cd_wrap_closest_distance(c.s[m], pt);
            if (dist < best.distance) { best.distance = dist; best.item = m; }
            if (todoPos > 0) node = todo[--todoPos]; else break;
        } else {
            const double bestDistSqr = best.distance * best.distance;
            const int left = node + 1, right = nd.child;
            double dl[3], dr[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                dl[a] = jc_clamp(pt[a], c.nodes[left].lo[a], c.nodes[left].hi[a]) - pt[a];
                dr[a] = jc_clamp(pt[a], c.nodes[right].lo[a], c.nodes[right].hi[a]) - pt[a];
            }
            const double dlsq = cd_len2(dl), drsq = cd_len2(dr);
            const bool visitLeft = dlsq < bestDistSqr, visitRight = drsq < bestDistSqr;
            if (visitLeft && visitRight) {
                int first, second;
                if (dlsq < drsq) { first = left; second = right; } else { first = right; second = left; }
                todo[todoPos++] = second;
                node = first;
            } else if (visitLeft) node = left;
            else if (visitRight) node = right;
            else if (todoPos > 0) node = todo[--todoPos];
            else break;
        }
    }
    return best;
}

// ---- ImplicitSurfaceSet3::closestPointLocal / closestNormalLocal / closestDistanceLocal / isInsideLocal
// (implicit_surface_set3.cpp:94-161, 206-214) in one query at the set-local point p ---------------------------------------
__host__ __device__ __forceinline__ ColliderQuery cd_set_query_local(const ColliderD& c, const double* p) {
    ColliderQuery q;
    q.point[0] = q.point[1] = q.point[2] = DBL_MAX;
    q.normal[0] = 1.0; q.normal[1] = 0.0; q.normal[2] = 0.0;
    const BvhNearest nn = cd_bvh_nearest(c, p);
    if (nn.item >= 0) {
        const ColliderQuery w = cd_wrap_query(c.s[nn.item], p);
        q.point[0] = w.point[0]; q.point[1] = w.point[1]; q.point[2] = w.point[2];
        q.normal[0] = w.normal[0]; q.normal[1] = w.normal[1]; q.normal[2] = w.normal[2];
    }
    double minDist = nn.distance;
    for (int u = 0; u < c.nunb; ++u) {
        const ColliderQuery w = cd_wrap_query(c.s[c.unb[u]], p);
        const double dist = cd_dist(w.point, p);  // pt.distanceTo(otherPoint)
        if (dist < minDist) {
            minDist = dist;
            q.point[0] = w.point[0]; q.point[1] = w.point[1]; q.point[2] = w.point[2];
            q.normal[0] = w.normal[0]; q.normal[1] = w.normal[1]; q.normal[2] = w.normal[2];
        }
    }
    q.distance = minDist;
    q.inside = false;
    for (int i = 0; i < c.n; ++i) {
        double l[3];
        cd_to_local(p, l);  // the member wrapper's transform, then the shape's own inside cd_surf_query
        if (cd_surf_query(c.s[i], l).inside) { q.inside = true; break; }
    }
    return q;
}

// One query of the collider's surface (Collider3::surface()) at world point p through Surface3's public functions
__host__ __device__ __forceinline__ ColliderQuery cd_query(const ColliderD& c, const double* p) {
    if (!c.is_set) return cd_surf_query(c.s[0], p);
    double l[3];
    cd_to_local(p, l);
    const ColliderQuery s = cd_set_query_local(c, l);
    ColliderQuery q;
    cd_to_world(s.point, q.point);
    cd_to_world_dir(s.normal, q.normal);
    q.normal[0] *= 1.0; q.normal[1] *= 1.0; q.normal[2] *= 1.0;
    q.distance = s.distance;
    q.inside = s.inside;
    return q;
}

// onColliderUpdated's implicitSurface->signedDistance(pt) (…boundary_condition_solver3.cpp:255-264): the set itself (min
// over the members, implicit_surface_set3.cpp:216-224) or SurfaceToImplicit3(surface) around an explicit surface
__host__ __device__ __forceinline__ double cd_signed_distance(const ColliderD& c, const double* pt) {
    if (!c.is_set) return cd_wrap_signed_distance(c.s[0], pt);
    double l[3];
    cd_to_local(pt, l);
    double sdf = DBL_MAX;
    for (int i = 0; i < c.n; ++i) sdf = cd_min(sdf, cd_wrap_signed_distance(c.s[i], l));
    return sdf;
}

// ---- host side: Surface3::boundingBox() of the members and Bvh3::build (bvh3-inl.h:49-69, 522-581) ----------------------
// Transform3::toWorld(BoundingBox3D) with the identity transform (transform3-inl.h:87-98): the box of the eight transformed
// corners, where corner(i) = (width, height, depth) * (-+1/2) + midPoint (bounding_box3-inl.h:150-152, 201-208) is NOT
// exactly the stored corner — the round trip can move a bound by an ulp, and the BVH is built from the moved bounds.
inline void cd_bbox_to_world(const double* l, const double* u, double* lo, double* hi) {
    const double size[3] = {u[0] - l[0], u[1] - l[1], u[2] - l[2]};
    const double mid[3] = {(u[0] + l[0]) / 2.0, (u[1] + l[1]) / 2.0, (u[2] + l[2]) / 2.0};
    for (int a = 0; a < 3; ++a) { lo[a] = DBL_MAX; hi[a] = -DBL_MAX; }
    for (int i = 0; i < 8; ++i) {
        const double off[3] = {(i & 1) ? 0.5 : -0.5, (i & 2) ? 0.5 : -0.5, (i & 4) ? 0.5 : -0.5};
        const double corner[3] = {size[0] * off[0] + mid[0], size[1] * off[1] + mid[1], size[2] * off[2] + mid[2]};
        double w[3];
        cd_to_world(corner, w);
        for (int a = 0; a < 3; ++a) { lo[a] = cd_min(lo[a], w[a]); hi[a] = cd_max(hi[a], w[a]); }
    }
}
// ImplicitSurface3::boundingBox() of SurfaceToImplicit3(shape): the shape's boundingBoxLocal through the shape's and the
// wrapper's (identity) transforms
inline void cd_shape_bounding_box(const ShapeD& c, double* lo, double* hi) {
    double p1[3], p2[3], l0[3], h0[3], l1[3], h1[3];
    for (int a = 0; a < 3; ++a) {
        if (c.kind == JC_SHAPE_BOX) { p1[a] = c.a[a]; p2[a] = c.b[a]; }
        else if (c.kind == JC_SHAPE_SPHERE) { p1[a] = c.a[a] - c.b[0]; p2[a] = c.a[a] + c.b[0]; }            // sphere3.cpp:96-99
        else { const double e = (a == 1) ? 0.5 * c.b[1] : c.b[0]; p1[a] = c.a[a] - e; p2[a] = c.a[a] + e; }  // cylinder3.cpp:223-226
        l0[a] = cd_min(p1[a], p2[a]); h0[a] = cd_max(p1[a], p2[a]);                                          // BoundingBox3D(p1, p2)
    }
    cd_bbox_to_world(l0, h0, l1, h1);
    cd_bbox_to_world(l1, h1, lo, hi);
}
inline int cd_bvh_qsplit(int* idx, int n, double pivot, int axis, const double (*blo)[3], const double (*bhi)[3]) {
    int ret = 0;
    for (int i = 0; i < n; ++i) {
        const double centroid = 0.5f * (blo[idx[i]][axis] + bhi[idx[i]][axis]);
        if (centroid < pivot) {
            const int t = idx[i]; idx[i] = idx[ret]; idx[ret] = t;
            ret++;
        }
    }
    if (ret == 0 || ret == n) ret = n >> 1;
    return ret;
}
inline void cd_bvh_build(ColliderD& c, int node, int* idx, int n, const double (*blo)[3], const double (*bhi)[3]) {
    c.nnodes++;  // _nodes.push_back(Node()): `node` is its index
    BvhNodeD& nd = c.nodes[node];
    nd.pad = 0;
    if (n == 1) {
        nd.leaf = 1; nd.item = idx[0]; nd.child = -1;
        for (int a = 0; a < 3; ++a) { nd.lo[a] = blo[idx[0]][a]; nd.hi[a] = bhi[idx[0]][a]; }
        return;
    }
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int i = 0; i < n; ++i)
        for (int a = 0; a < 3; ++a) { lo[a] = cd_min(lo[a], blo[idx[i]][a]); hi[a] = cd_max(hi[a], bhi[idx[i]][a]); }
    const double d[3] = {hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]};
    int axis;
    if (d[0] > d[1] && d[0] > d[2]) axis = 0; else axis = (d[1] > d[2]) ? 1 : 2;
    const double pivot = 0.5 * (hi[axis] + lo[axis]);
    const int mid = cd_bvh_qsplit(idx, n, pivot, axis, blo, bhi);
    cd_bvh_build(c, node + 1, idx, mid, blo, bhi);
    nd.leaf = 0; nd.item = -1; nd.child = c.nnodes;
    for (int a = 0; a < 3; ++a) { nd.lo[a] = lo[a]; nd.hi[a] = hi[a]; }
    cd_bvh_build(c, nd.child, idx + mid, n - mid, blo, bhi);
}
// ImplicitSurfaceSet3::buildBvh (implicit_surface_set3.cpp:233-246) + the _unboundedSurfaces list (:86-91)
inline void cd_build_set(ColliderD& c) {
    double blo[JC_COLLIDER_MAX_SHAPES][3], bhi[JC_COLLIDER_MAX_SHAPES][3];
    c.nnodes = c.nitems = c.nunb = 0;
    for (int i = 0; i < c.n; ++i) {
        if (c.s[i].kind == JC_SHAPE_PLANE) { c.unb[c.nunb++] = i; continue; }  // Plane3::isBounded() == false
        cd_shape_bounding_box(c.s[i], blo[c.nitems], bhi[c.nitems]);
        c.items[c.nitems++] = i;
    }
    if (c.nitems == 0) return;
    int idx[JC_COLLIDER_MAX_SHAPES];
    for (int i = 0; i < c.nitems; ++i) idx[i] = i;
    cd_bvh_build(c, 0, idx, c.nitems, blo, bhi);
}

// VolumeParticleEmitter3::emit's region (src/jet/volume_particle_emitter3.cpp:74-79): maxRegion [lo, hi] ∩ the implicit
// surface's boundingBox() when the surface is bounded (no Plane3 member).  The surface is SurfaceToImplicit3(shape) for one
// explicit surface, the set itself otherwise: Surface3::boundingBox() = toWorld(_bvh.boundingBox()), the merge of the members.
inline void cd_emit_region(const ColliderD& e, double* lo, double* hi) {
    for (int i = 0; i < e.n; ++i)
        if (e.s[i].kind == JC_SHAPE_PLANE) return;
    double sl[3], sh[3];
    if (e.is_set) {
        double ml[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mh[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int i = 0; i < e.n; ++i) {
            double bl[3], bh[3];
            cd_shape_bounding_box(e.s[i], bl, bh);
            for (int a = 0; a < 3; ++a) { ml[a] = cd_min(ml[a], bl[a]); mh[a] = cd_max(mh[a], bh[a]); }
        }
        cd_bbox_to_world(ml, mh, sl, sh);
    } else {
        cd_shape_bounding_box(e.s[0], sl, sh);
    }
    for (int a = 0; a < 3; ++a) { lo[a] = cd_max(lo[a], sl[a]); hi[a] = cd_min(hi[a], sh[a]); }
}
// ... and its acceptance test at lattice point p (jitter 0): _implicitSurface->signedDistance(p) <= 0 (:93)
__host__ __device__ __forceinline__ bool cd_emit_accept(const ColliderD& e, const double* p) { return cd_signed_distance(e, p) <= 0.0; }

// The ColliderD of jetcu_hybrid_set_collider_shapes' arguments (already validated): members, rigid-body velocity, friction
// clamped at 0 as Collider3::setFrictionCoefficient does (collider3.cpp:78-80), and — for a set — the Bvh3, built once:
// Collider3::update -> updateQueryEngine -> buildBvh gives the same tree at every step for a static set.
inline ColliderD cd_make_collider(const jetcu_collider_shape* shapes, uint32_t n, int as_set, const double* lin, const double* ang,
                                  double friction) {
    ColliderD c;
    if (n == 0) return c;
    c.kind = 1;
    c.n = (int)n;
    c.is_set = as_set ? 1 : 0;
    for (uint32_t i = 0; i < n; ++i) {
        c.s[i].kind = shapes[i].kind;
        c.s[i].flipped = shapes[i].flipped ? 1 : 0;
        for (int q = 0; q < 3; ++q) { c.s[i].a[q] = shapes[i].a[q]; c.s[i].b[q] = shapes[i].b[q]; }
    }
    for (int q = 0; q < 3; ++q) { c.lin[q] = lin ? lin[q] : 0.0; c.ang[q] = ang ? ang[q] : 0.0; }
    c.friction = friction > 0.0 ? friction : 0.0;
    if (c.is_set) cd_build_set(c);
    return c;
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
