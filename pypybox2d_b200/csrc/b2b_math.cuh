// b2b_math.cuh -- scalar-type-generic 2D math and tuning constants for the device step.
//
// Conventions are the reference's (pypybox2d/common.py, summarised in SURVEY A.2):
//   cross(v, s) = (s*v.y, -s*v.x)        common.py:204-208
//   cross(s, v) = (-s*v.y, s*v.x)        common.py:195-202
//   R(theta): col1 = (c, s), col2 = (-s, c); R*v, R^T*v, xf*v = p + R*v, xf^T*v = R^T*(v - p)
//   normalize returns its input below epsilon   common.py:216-222
// Operand order inside each expression follows the reference so that the fp64 instantiation
// (compiled with -fmad=false) reproduces the oracle to the last bit wherever libm is not involved.
//
// Everything here is B2B_HD so that tests/hostcheck can instantiate the same source with g++;
// the product library only ever contains the device instantiations.
#pragma once

#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define B2B_HD __host__ __device__ __forceinline__
#define B2B_D __device__ __forceinline__
#define B2B_HD_COLD __host__ __device__ __noinline__
#define B2B_D_NOINLINE __device__ __noinline__
#else
#define B2B_HD inline
#define B2B_D inline
#define B2B_HD_COLD inline
#define B2B_D_NOINLINE inline
#endif

namespace b2b {

// ---- flag bits (pypybox2d_b200/schema.py) ----
enum : uint32_t {
    BODY_TYPE_MASK = 0x3u, BODY_AWAKE = 0x4u, BODY_ACTIVE = 0x8u, BODY_ALLOW_SLEEP = 0x10u,
    BODY_FIXED_ROT = 0x20u, BODY_BULLET = 0x40u, BODY_ISLAND = 0x80u,
    BODY_WOKEN = 0x100u,   // transient, shared-memory copy only: woken by the island walk this step
    CONTACT_TOUCHING = 0x1u, CONTACT_ENABLED = 0x2u, CONTACT_FILTER = 0x4u, CONTACT_ISLAND = 0x8u,
    CONTACT_SENSOR = 0x10u,      // one of the two fixtures is a sensor (set when the contact is created): no manifold, touching = GJK overlap
    WORLD_ALLOW_SLEEP = 0x1u, WORLD_WARM_STARTING = 0x2u, WORLD_NEW_FIXTURE = 0x4u, WORLD_AUTO_CLEAR = 0x8u,
    WORLD_GRAPH_VALID = 0x10u,   // the island lists / level schedule of the last step still describe this world
    WORLD_CONTINUOUS = 0x20u,    // continuous_physics (world.py:78): World.solve_toi runs after the discrete solve
    WORLD_TOI_DIRTY = 0x40u,     // device bookkeeping: some body's sweep.alpha0 is non-zero (a TOI event was solved last step)
    STATUS_CONTACT_OVERFLOW = 0x1u, STATUS_PAIR_OVERFLOW = 0x2u, STATUS_NAN = 0x4u, STATUS_ISLAND_OVERFLOW = 0x8u,
    STATUS_TOI_STUCK = 0x10u,    // a time-of-impact push-back loop was cut off (b2b_toi.cuh TOI_MAX_PUSH_BACK)
    STATUS_JOINT_OVERFLOW = 0x20u // more joints than the batch's joint capacity
};
enum { BODY_STATIC = 0, BODY_DYNAMIC = 1, BODY_KINEMATIC = 2 };
enum { MANIFOLD_CIRCLES = 0, MANIFOLD_FACE_A = 1, MANIFOLD_FACE_B = 2 };
enum { SHAPE_CIRCLE = 0, SHAPE_POLYGON = 1, SHAPE_LOOP = 2, SHAPE_EDGE = 3 };
enum { MP_VERTEX = 0, MP_FACE = 1 };
enum { SHAPE_DATA = 32, MAX_POLY = 8 };

// ---- tuning constants: pypybox2d/settings.py:32-134 ----
template <class real> struct K {
    static constexpr real pi = real(3.141592653589793);
    static constexpr real epsilon = real(2.220446049250313e-16);         // fp64 epsilon, settings.py:32
    static constexpr real epsilon_sqr = real(4.930380657631324e-32);
    static constexpr real max_float = sizeof(real) == 8 ? real(1.7976931348623157e308) : real(3.4028234e38);
    static constexpr real aabb_extension = real(0.1);
    static constexpr real aabb_multiplier = real(2.0);
    static constexpr real linear_slop = real(0.005);
    static constexpr real angular_slop = real(2.0 / 180.0 * 3.141592653589793);
    static constexpr real polygon_radius = real(2.0 * 0.005);
    static constexpr real velocity_threshold = real(1.0);
    static constexpr real max_linear_correction = real(0.2);
    static constexpr real max_angular_correction = real(8.0 / 180.0 * 3.141592653589793);   // settings.py:92
    static constexpr real max_translation = real(2.0);
    static constexpr real max_translation_sqr = real(4.0);
    static constexpr real max_rotation = real(0.5 * 3.141592653589793);
    static constexpr real max_rotation_sqr = real(2.4674011002723395);           // (0.5 * pi) ** 2 as CPython evaluates it (settings.py:104)
    static constexpr real baumgarte = real(0.2);
    static constexpr real time_to_sleep = real(0.5);
    static constexpr real linear_sleep_tol_sqr = real(0.0001);   // 0.01 ** 2 (settings.py:119)
    static constexpr real angular_sleep_tol_sqr =
        real((2.0 / 180.0 * 3.141592653589793) * (2.0 / 180.0 * 3.141592653589793));
    static constexpr real max_condition_number = real(1000.0);           // contact.py:488
    static constexpr real relative_tol = real(0.98);                     // collision.py:39
    static constexpr real absolute_tol = real(0.001);                    // collision.py:40
};

template <class real> struct vec2 {
    real x, y;
};

template <class real> B2B_HD vec2<real> mk(real x, real y) { vec2<real> r; r.x = x; r.y = y; return r; }
template <class real> B2B_HD vec2<real> operator+(vec2<real> a, vec2<real> b) { return mk<real>(a.x + b.x, a.y + b.y); }
template <class real> B2B_HD vec2<real> operator-(vec2<real> a, vec2<real> b) { return mk<real>(a.x - b.x, a.y - b.y); }
template <class real> B2B_HD vec2<real> operator-(vec2<real> a) { return mk<real>(-a.x, -a.y); }
template <class real> B2B_HD vec2<real> operator*(real s, vec2<real> a) { return mk<real>(s * a.x, s * a.y); }
template <class real> B2B_HD real dot(vec2<real> a, vec2<real> b) { return a.x * b.x + a.y * b.y; }
template <class real> B2B_HD real cross(vec2<real> a, vec2<real> b) { return a.x * b.y - a.y * b.x; }
template <class real> B2B_HD vec2<real> cross_vs(vec2<real> a, real s) { return mk<real>(s * a.y, -s * a.x); }
template <class real> B2B_HD vec2<real> cross_sv(real s, vec2<real> a) { return mk<real>(-s * a.y, s * a.x); }
// tangent = normal.cross(1.0)
template <class real> B2B_HD vec2<real> right_perp(vec2<real> n) { return mk<real>(n.y, -n.x); }
// The reference writes many squares as `x ** 2`, which CPython evaluates with libm's pow: Vec2.length /
// length_squared / normalized (common.py:186-190, 216-222), the effective-mass terms (contact.py:444-457, 489, 841-842),
// the step clamps and sleep tests (island.py:124-128, 169-170), the radius tests of the colliders (collision.py:56-336).
// glibc's pow(x, 2.0) is NOT always the correctly rounded product: it differs from x * x by one ulp in ~0.09 % of
// evaluations.  sq_r is that square: the host build of the fp64 step (tests/hostcheck) calls pow, like the oracle, and is
// bit-identical to it on every scene; the device computes the correctly rounded product (CUDA's pow is another
// algorithm again, 2 ulp) -- together with sincos (<= 2 ulp from glibc's) these are the ONLY operations in which the fp64
// kernels differ from the reference, one ulp at a time (tests/test_gpu_parity.py test_p_exact_fp64_one_step).
B2B_HD float sq_r(float x) { return x * x; }
B2B_HD double sq_r(double x)
{
#if defined(__CUDA_ARCH__) || defined(B2B_SQ_PRODUCT)      // (B2B_SQ_PRODUCT: tools/sincos_ulp_probe.py measures what the product costs)
    return x * x;
#else
    return pow(x, 2.0);
#endif
}
template <class real> B2B_HD real length_sqr(vec2<real> a) { return sq_r(a.x) + sq_r(a.y); }

B2B_HD float sqrt_r(float x) { return sqrtf(x); }
B2B_HD double sqrt_r(double x) { return sqrt(x); }
B2B_HD float abs_r(float x) { return fabsf(x); }
B2B_HD double abs_r(double x) { return fabs(x); }
B2B_HD void sincos_r(float a, float* s, float* c)
{
#if defined(__CUDA_ARCH__)
    sincosf(a, s, c);
#else
    *s = sinf(a); *c = cosf(a);
#endif
}
B2B_HD void sincos_r(double a, double* s, double* c)
{
#if defined(__CUDA_ARCH__)
    sincos(a, s, c);
#else
    *s = sin(a); *c = cos(a);
#if defined(B2B_SINCOS_ULP_PROBE)
    // tests/hostcheck experiment (tools/sincos_ulp_probe.py): what a libm that differs from glibc's in the last place does
    // to a trajectory.  About half of the results are moved by one ulp, chosen by the bits of the angle.
    {
        uint64_t u; memcpy(&u, &a, 8);
        u ^= u >> 29; u *= 0xbf58476d1ce4e5b9ull; u ^= u >> 32;
        if (u & 1) *s = nextafter(*s, (u & 2) ? 2.0 : -2.0);
        if (u & 4) *c = nextafter(*c, (u & 8) ? 2.0 : -2.0);
    }
#endif
#endif
}

// Python's min/max: the first argument wins ties (matters only for signed zeros)
template <class real> B2B_HD real pymin(real a, real b) { return (b < a) ? b : a; }
template <class real> B2B_HD real pymax(real a, real b) { return (b > a) ? b : a; }
template <class real> B2B_HD real clamp_r(real v, real lo, real hi) { return pymax(lo, pymin(v, hi)); }

template <class real> B2B_HD vec2<real> normalize(vec2<real> a)
{
    real len = sqrt_r(sq_r(a.x) + sq_r(a.y));
    if (len < K<real>::epsilon) return a;
    real inv = real(1) / len;
    return mk<real>(a.x * inv, a.y * inv);
}

// Rigid transform with a general 2x2 rotation block (EPCollider composes two rotations).
template <class real> struct xform {
    vec2<real> p, c1, c2;
};
template <class real> B2B_HD xform<real> xf_make(vec2<real> p, real c, real s)
{
    xform<real> t; t.p = p; t.c1 = mk<real>(c, s); t.c2 = mk<real>(-s, c); return t;
}
template <class real> B2B_HD xform<real> xf_load(const real* r) { return xf_make<real>(mk<real>(r[0], r[1]), r[2], r[3]); }
template <class real> B2B_HD vec2<real> rot_mul(const xform<real>& t, vec2<real> v)
{
    return mk<real>(t.c1.x * v.x + t.c2.x * v.y, t.c1.y * v.x + t.c2.y * v.y);
}
template <class real> B2B_HD vec2<real> rot_mul_t(const xform<real>& t, vec2<real> v)
{
    return mk<real>(dot(t.c1, v), dot(t.c2, v));
}
template <class real> B2B_HD vec2<real> xf_mul(const xform<real>& t, vec2<real> v)
{
    return mk<real>(t.p.x + t.c1.x * v.x + t.c2.x * v.y, t.p.y + t.c1.y * v.x + t.c2.y * v.y);
}
template <class real> B2B_HD vec2<real> xf_mul_t(const xform<real>& t, vec2<real> v) { return rot_mul_t(t, v - t.p); }
// Transform(angle=a) with position = c - R * local_center   (body.py:906-908, contact.py:433-436)
template <class real> B2B_HD xform<real> xf_from_sweep(vec2<real> c, real a, vec2<real> lc)
{
    real s, co;
    sincos_r(a, &s, &co);
    xform<real> t = xf_make<real>(mk<real>(real(0), real(0)), co, s);
    t.p = c - rot_mul(t, lc);
    return t;
}

// AABB.test_overlap: touching boxes overlap (common.py:830-839).  Boxes are (lx, ly, ux, uy).
template <class real> B2B_HD bool aabb_overlap(const real* a, const real* b)
{
    real d1x = b[0] - a[2], d1y = b[1] - a[3];
    real d2x = a[0] - b[2], d2y = a[1] - b[3];
    if (d1x > real(0) || d1y > real(0)) return false;
    if (d2x > real(0) || d2y > real(0)) return false;
    return true;
}

B2B_HD uint32_t make_id(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}

}  // namespace b2b
