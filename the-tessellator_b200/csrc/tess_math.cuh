// tess_math.cuh — device twin of the reference's float.rs / vector3.rs (f64).
//
// The GPU computes in the reference's own precision and operation order: every sum is
// left-associated exactly as the Rust expression is written and nothing is contracted into an
// FMA (Rust never contracts).  The translation unit is compiled with -fmad=false; the explicit
// __dmul_rn/__dadd_rn intrinsics below make the intent independent of that flag.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tess {

struct Vec3 {  // vector3.rs:24-28
    double x, y, z;
};

__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double addd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double subd(double a, double b) { return __dsub_rn(a, b); }

/// vector3.rs:38-40  a.x*b.x + a.y*b.y + a.z*b.z  (left-associated)
__device__ __forceinline__ double dot(const Vec3& a, const Vec3& b) {
    return addd(addd(mul(a.x, b.x), mul(a.y, b.y)), mul(a.z, b.z));
}
__device__ __forceinline__ double dot3(double ax, double ay, double az, double bx, double by, double bz) {
    return addd(addd(mul(ax, bx), mul(ay, by)), mul(az, bz));
}
/// vector3.rs:44-50
__device__ __forceinline__ Vec3 cross(const Vec3& a, const Vec3& b) {
    return {subd(mul(a.y, b.z), mul(a.z, b.y)), subd(mul(a.z, b.x), mul(a.x, b.z)), subd(mul(a.x, b.y), mul(a.y, b.x))};
}
/// vector3.rs:53-59
__device__ __forceinline__ Vec3 scale(const Vec3& a, double s) { return {mul(a.x, s), mul(a.y, s), mul(a.z, s)}; }
/// vector3.rs:63-65
__device__ __forceinline__ double mag_sq(const Vec3& a) { return dot(a, a); }
/// vector3.rs:94-104 / 106-116
__device__ __forceinline__ Vec3 add(const Vec3& a, const Vec3& b) { return {addd(a.x, b.x), addd(a.y, b.y), addd(a.z, b.z)}; }
__device__ __forceinline__ Vec3 sub(const Vec3& a, const Vec3& b) { return {subd(a.x, b.x), subd(a.y, b.y), subd(a.z, b.z)}; }

struct Plane {  // vector3.rs:158-164
    double nx, ny, nz, off;
};

/// Plane::halfway_from_origin_to (vector3.rs:223-225):
///   unit() = scale(1.0 / mag())  (:75-77), point.scale(0.5), offset = dot(unit, half) (:229-239)
__device__ __forceinline__ Plane halfway_from_origin_to(const Vec3& p) {
    const double m = __dsqrt_rn(mag_sq(p));
    const double inv = __ddiv_rn(1.0, m);
    const Vec3 n = scale(p, inv);
    const Vec3 h = scale(p, 0.5);
    return {n.x, n.y, n.z, dot(n, h)};
}

/// Plane::signed_distance (vector3.rs:191-193) = offset_inverse(v) - plane_offset
__device__ __forceinline__ double signed_distance(const Plane& pl, double x, double y, double z) {
    return subd(dot3(pl.nx, pl.ny, pl.nz, x, y, z), pl.off);
}

/// Plane::intersection (vector3.rs:213-219): a + (b - a) * ((off - n.a) / (n.b - n.a))
__device__ __forceinline__ Vec3 intersection(const Plane& pl, const Vec3& a, const Vec3& b) {
    const double ao = dot3(pl.nx, pl.ny, pl.nz, a.x, a.y, a.z);
    const double bo = dot3(pl.nx, pl.ny, pl.nz, b.x, b.y, b.z);
    const double t = __ddiv_rn(subd(pl.off, ao), subd(bo, ao));
    return add(a, scale(sub(b, a), t));
}

/// Polyhedron::tolerance (polyhedron.rs:221-223)
#define TESS_TOL 1e-12

/// `as usize` (float.rs:138-142) saturates: NaN/negative -> 0.  Grid indices fit u32.
__device__ __forceinline__ uint32_t sat_u32(double v) { return __double2uint_rz(v); }

/// Celery::get_{x,y,z}_cell_index (celery.rs:269-314)
__device__ __forceinline__ uint32_t axis_index(double v, double vmin, double vmax, double inv, uint32_t cpd) {
    if (v >= vmax) return cpd - 1;
    const uint32_t i = sat_u32(mul(subd(v, vmin), inv));
    return i < cpd - 1 ? i : cpd - 1;
}

}  // namespace tess
