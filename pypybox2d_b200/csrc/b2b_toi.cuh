// b2b_toi.cuh -- continuous collision: GJK distance, separation function, time of impact.
//
// Device restatement (one thread per contact) of
//   Sweep.get_transform / advance / normalize     pypybox2d/common.py:893-923
//   DistanceProxy.get_support                     distance.py:393-396
//   Simplex (read_cache, solve2, solve3, ...)     distance.py:76-352
//   shape_distance                                distance.py:419-561
//   SeparationFunction                            toi.py:46-221
//   time_of_impact                                toi.py:223-391
// used by World.solve_toi (world.py:594-821), which world_step::phase_toi (b2b_step.cuh) drives.
// Operand order follows the reference expression by expression (see b2b_math.cuh); the quirks that matter are kept:
//   * SeparationFunction.initialize leaves the POINTS axis un-normalised (toi.py:73-75 builds a normalised copy and drops it);
//   * solve2 / solve3 REBIND a simplex vertex to a copy (self._v1 = copy(self._v2)) while shape_distance keeps writing new
//     support points through the list it took before the loop (distance.py:458): after a rebind the list entry and the
//     attribute are different objects, so a later support point written "into" that slot never reaches the simplex.
#pragma once
#include "b2b_collide.cuh"

namespace b2b {

enum { TOI_UNKNOWN = 0, TOI_FAILED = 1, TOI_OVERLAPPED = 2, TOI_TOUCHING = 3, TOI_SEPARATED = 4 };
enum { TOI_MAX_ITERS = 20, DISTANCE_MAX_ITERS = 20, TOI_MAX_ROOT_ITERS = 50, MAX_SUB_STEPS = 8, MAX_TOI_CONTACTS = 32,
       // the reference's push-back loop has no bound (toi.py:374-376, commented out); in exact arithmetic it ends after at most
       // one pass per vertex pair.  A bound far above that keeps a rounding cycle from hanging a kernel (the world's status
       // word gets STATUS_TOI_STUCK).
       TOI_MAX_PUSH_BACK = 256 };

B2B_HD float floor_r(float x) { return floorf(x); }
B2B_HD double floor_r(double x) { return floor(x); }

template <class real> struct sweep_t {
    vec2<real> lc, c0, c;
    real a0, a, alpha0;
};
// (out of line everywhere: every copy carries a sincos with its large-argument path, and the TOI kernel is bound by instruction fetch)
template <class real> B2B_HD_COLD xform<real> sweep_xf(const sweep_t<real>& s, real beta)                   // common.py:893-904
{
    const vec2<real> p = (real(1) - beta) * s.c0 + beta * s.c;
    const real angle = (real(1) - beta) * s.a0 + beta * s.a;
    return xf_from_sweep<real>(p, angle, s.lc);
}
template <class real> B2B_HD void sweep_advance(sweep_t<real>& s, real alpha)                          // common.py:906-915
{
    const real beta = (alpha - s.alpha0) / (real(1) - s.alpha0);
    s.c0 = (real(1) - beta) * s.c0 + beta * s.c;
    s.a0 = (real(1) - beta) * s.a0 + beta * s.a;
    s.alpha0 = alpha;
}
template <class real> B2B_HD void sweep_normalize(sweep_t<real>& s)                                    // common.py:917-923
{
    const real two_pi = real(2) * K<real>::pi;
    const real d = two_pi * floor_r(s.a0 / two_pi);
    s.a0 -= d;
    s.a -= d;
}

// DistanceProxy (distance.py:354-409): a circle is its centre, a polygon its vertices, an edge its two end points
template <class real> B2B_HD int proxy_count(const shape_view<real>& s)
{
    return (s.type == SHAPE_CIRCLE) ? 1 : (s.type == SHAPE_POLYGON) ? s.count : 2;
}
// get_support: dots.index(max(dots)) -- the first index that holds the maximum
template <class real> B2B_HD int proxy_support(const shape_view<real>& s, vec2<real> d)
{
    const int n = proxy_count(s);
    int best = 0;
    real bv = dot(s.vert(0), d);
    for (int i = 1; i < n; ++i) {
        const real v = dot(s.vert(i), d);
        if (v > bv) { bv = v; best = i; }
    }
    return best;
}

template <class real> struct simplex_cache {
    real metric;
    int count;
    int ia[3], ib[3];
};
template <class real> struct simplex_vertex {
    vec2<real> wa, wb, w;
    real a;
    int ia, ib;
};
template <class real> B2B_HD real dist_r(vec2<real> a, vec2<real> b)                                   // common.py:947-949
{
    const vec2<real> c = a - b;
    return sqrt_r(sq_r(c.x) + sq_r(c.y));
}

// The simplex: v[k] is the CURRENT value of the attribute _v(k+1); linked[k] says whether that attribute is still the object
// the list entry k refers to (true until solve2 / solve3 rebind it to a copy).
template <class real> struct simplex_t {
    simplex_vertex<real> v[3];
    bool linked[3];
    int count;
    B2B_HD void rebind(int dst, int src) { v[dst] = v[src]; linked[dst] = false; }
    B2B_HD real metric() const                                                                          // distance.py:189-203
    {
        if (count == 1) return real(0);
        if (count == 2) return dist_r(v[0].w, v[1].w);
        return cross(v[1].w - v[0].w, v[2].w - v[0].w);
    }
    B2B_HD void solve2()                                                                                // distance.py:205-253
    {
        const vec2<real> w1 = v[0].w, w2 = v[1].w, e12 = w2 - w1;
        const real d12_2 = -dot(w1, e12);
        if (d12_2 <= real(0)) { v[0].a = real(1); count = 1; return; }
        const real d12_1 = dot(w2, e12);
        if (d12_1 <= real(0)) { v[1].a = real(1); count = 1; rebind(0, 1); return; }
        const real inv_d12 = real(1) / (d12_1 + d12_2);
        v[0].a = d12_1 * inv_d12; v[1].a = d12_2 * inv_d12; count = 2;
    }
    B2B_HD void solve3()                                                                                // distance.py:255-352
    {
        const vec2<real> w1 = v[0].w, w2 = v[1].w, w3 = v[2].w;
        const vec2<real> e12 = w2 - w1;
        const real w1e12 = dot(w1, e12), w2e12 = dot(w2, e12), d12_1 = w2e12, d12_2 = -w1e12;
        const vec2<real> e13 = w3 - w1;
        const real w1e13 = dot(w1, e13), w3e13 = dot(w3, e13), d13_1 = w3e13, d13_2 = -w1e13;
        const vec2<real> e23 = w3 - w2;
        const real w2e23 = dot(w2, e23), w3e23 = dot(w3, e23), d23_1 = w3e23, d23_2 = -w2e23;
        const real n123 = cross(e12, e13);
        const real d123_1 = n123 * cross(w2, w3), d123_2 = n123 * cross(w3, w1), d123_3 = n123 * cross(w1, w2);
        if (d12_2 <= real(0) && d13_2 <= real(0)) { v[0].a = real(1); count = 1; return; }
        if (d12_1 > real(0) && d12_2 > real(0) && d123_3 <= real(0)) {
            const real inv = real(1) / (d12_1 + d12_2);
            v[0].a = d12_1 * inv; v[1].a = d12_2 * inv; count = 2; return;
        }
        if (d13_1 > real(0) && d13_2 > real(0) && d123_2 <= real(0)) {
            const real inv = real(1) / (d13_1 + d13_2);
            v[0].a = d13_1 * inv; v[2].a = d13_2 * inv; count = 2; rebind(1, 2); return;
        }
        if (d12_1 <= real(0) && d23_2 <= real(0)) { v[1].a = real(1); count = 1; rebind(0, 1); return; }
        if (d13_1 <= real(0) && d23_1 <= real(0)) { v[2].a = real(1); count = 1; rebind(0, 2); return; }
        if (d23_1 > real(0) && d23_2 > real(0) && d123_1 <= real(0)) {
            const real inv = real(1) / (d23_1 + d23_2);
            v[1].a = d23_1 * inv; v[2].a = d23_2 * inv; count = 2; rebind(0, 2); return;
        }
        const real inv = real(1) / (d123_1 + d123_2 + d123_3);
        v[0].a = d123_1 * inv; v[1].a = d123_2 * inv; v[2].a = d123_3 * inv; count = 3;
    }
};

// shape_distance: distance.py:419-561.  The cache is read and written.  use_radii: World.solve_toi passes False; test_overlap
// (sensor contacts) passes True -- the distance then is the gap between the rounded shapes, 0 when they overlap.
template <class real>
B2B_HD_COLD real shape_distance(const shape_view<real>& pa, const shape_view<real>& pb, const xform<real>& xa,
                                const xform<real>& xb, simplex_cache<real>& cache, bool use_radii = false)
{
    simplex_t<real> S;
    for (int k = 0; k < 3; ++k) {
        S.linked[k] = true;
        S.v[k].wa = S.v[k].wb = S.v[k].w = mk<real>(real(0), real(0));
        S.v[k].a = real(0); S.v[k].ia = 0; S.v[k].ib = 0;
    }
    // read_cache: distance.py:97-135
    S.count = cache.count;
    for (int i = 0; i < S.count; ++i) {
        simplex_vertex<real>& v = S.v[i];
        v.ia = cache.ia[i]; v.ib = cache.ib[i];
        v.wa = xf_mul(xa, pa.vert(v.ia));
        v.wb = xf_mul(xb, pb.vert(v.ib));
        v.w = v.wb - v.wa;
        v.a = real(0);
    }
    if (S.count > 1) {
        const real metric1 = cache.metric, metric2 = S.metric();
        if (metric2 < (real(0.5) * metric1) || (real(2) * metric1) < metric2 || metric2 < K<real>::epsilon) S.count = 0;
    }
    if (S.count == 0) {
        simplex_vertex<real>& v = S.v[0];
        v.ia = 0; v.ib = 0;
        v.wa = xf_mul(xa, pa.vert(0));
        v.wb = xf_mul(xb, pb.vert(0));
        v.w = v.wb - v.wa;
        S.count = 1;
    }
    int iter = 0;
    while (iter < (int)DISTANCE_MAX_ITERS) {
        int save_a[3], save_b[3];
        const int nsave = S.count;
        for (int i = 0; i < 3; ++i) { save_a[i] = S.v[i].ia; save_b[i] = S.v[i].ib; }
        if (S.count == 2) S.solve2();
        else if (S.count == 3) { S.solve3(); break; }
        // search_direction: distance.py:152-169
        vec2<real> d;
        if (S.count == 1) d = -S.v[0].w;
        else {
            const vec2<real> e12 = S.v[1].w - S.v[0].w;
            const real sgn = cross(e12, -S.v[0].w);
            d = (sgn > real(0)) ? cross_sv(real(1), e12) : cross_vs(e12, real(1));
        }
        if (length_sqr(d) < K<real>::epsilon_sqr) break;
        // the new support point goes to the LIST entry S.count: it reaches the simplex only while that entry is linked
        const int n = S.count;
        const int ia = proxy_support(pa, rot_mul_t(xa, -d));
        const vec2<real> wa = xf_mul(xa, pa.vert(ia));
        const int ib = proxy_support(pb, rot_mul_t(xb, d));
        const vec2<real> wb = xf_mul(xb, pb.vert(ib));
        if (S.linked[n]) {
            simplex_vertex<real>& v = S.v[n];
            v.ia = ia; v.wa = wa; v.ib = ib; v.wb = wb; v.w = wb - wa;
        }
        ++iter;
        bool dup = false;
        for (int i = 0; i < 3; ++i) if (i < nsave && save_a[i] == ia && save_b[i] == ib) dup = true;
        if (dup) break;
        S.count += 1;
    }
    // witness_points: distance.py:171-187
    vec2<real> point_a, point_b;
    if (S.count == 1) { point_a = S.v[0].wa; point_b = S.v[0].wb; }
    else if (S.count == 2) {
        point_a = S.v[0].a * S.v[0].wa + S.v[1].a * S.v[1].wa;
        point_b = S.v[0].a * S.v[0].wb + S.v[1].a * S.v[1].wb;
    } else {
        point_a = S.v[0].a * S.v[0].wa + S.v[1].a * S.v[1].wa + S.v[2].a * S.v[2].wa;
        point_b = point_a;
    }
    const real dist = dist_r(point_a, point_b);
    // write_cache: distance.py:137-142
    cache.metric = S.metric();
    cache.count = S.count;
    for (int i = 0; i < 3; ++i) if (i < S.count) { cache.ia[i] = S.v[i].ia; cache.ib[i] = S.v[i].ib; }
    if (use_radii) {                                                                                    // distance.py:541-559
        const real ra = pa.radius, rb = pb.radius;
        if (dist > (ra + rb) && dist > K<real>::epsilon) return dist - (ra + rb);
        return real(0);
    }
    return dist;
}

// test_overlap: distance.py:411-417
template <class real>
B2B_HD bool shapes_overlap(const shape_view<real>& pa, const shape_view<real>& pb, const xform<real>& xa, const xform<real>& xb)
{
    simplex_cache<real> cache;
    cache.metric = real(0); cache.count = 0;
    for (int k = 0; k < 3; ++k) { cache.ia[k] = 0; cache.ib[k] = 0; }
    return shape_distance(pa, pb, xa, xb, cache, true) < (real(10) * K<real>::epsilon);
}

// SeparationFunction: toi.py:46-221
enum { SEP_POINTS = 0, SEP_FACE_A = 1, SEP_FACE_B = 2 };
template <class real> struct sep_function {
    shape_view<real> pa, pb;
    sweep_t<real> sa, sb;
    int type;
    vec2<real> axis, local_point;

    B2B_HD_COLD void initialize(const simplex_cache<real>& cache, real t1)
    {
        const xform<real> xa = sweep_xf(sa, t1), xb = sweep_xf(sb, t1);
        const int ia0 = cache.ia[0], ib0 = cache.ib[0];
        local_point = mk<real>(real(0), real(0));
        if (cache.count == 1) {
            type = SEP_POINTS;
            const vec2<real> point_a = xf_mul(xa, pa.vert(ia0)), point_b = xf_mul(xb, pb.vert(ib0));
            axis = point_b - point_a;                    // NOT normalised (toi.py:73-75)
            return;
        }
        const int ia1 = cache.ia[1], ib1 = cache.ib[1];
        if (ia0 == ia1) {
            type = SEP_FACE_B;
            const vec2<real> b1 = pb.vert(ib0), b2 = pb.vert(ib1);
            axis = normalize(cross_vs(b2 - b1, real(1)));
            const vec2<real> normal = rot_mul(xb, axis);
            local_point = real(0.5) * (b1 + b2);
            const vec2<real> point_b = xf_mul(xb, local_point);
            const vec2<real> point_a = xf_mul(xa, pa.vert(ia0));
            const real s = dot(point_a - point_b, normal);
            if (s < real(0)) axis = -axis;
        } else {
            type = SEP_FACE_A;
            const vec2<real> a1 = pa.vert(ia0), a2 = pa.vert(ia1);
            axis = normalize(cross_vs(a2 - a1, real(1)));
            const vec2<real> normal = rot_mul(xa, axis);
            local_point = real(0.5) * (a1 + a2);
            const vec2<real> point_a = xf_mul(xa, local_point);
            const vec2<real> point_b = xf_mul(xb, pb.vert(ib0));
            const real s = dot(point_b - point_a, normal);
            if (s < real(0)) axis = -axis;
        }
    }
    B2B_HD_COLD real find_min_separation(real t, int& index_a, int& index_b) const                          // toi.py:126-183
    {
        const xform<real> xa = sweep_xf(sa, t), xb = sweep_xf(sb, t);
        if (type == SEP_POINTS) {
            const vec2<real> axis_a = rot_mul_t(xa, axis), axis_b = rot_mul_t(xb, -axis);
            index_a = proxy_support(pa, axis_a);
            index_b = proxy_support(pb, axis_b);
            const vec2<real> point_a = xf_mul(xa, pa.vert(index_a)), point_b = xf_mul(xb, pb.vert(index_b));
            return dot(point_b - point_a, axis);
        } else if (type == SEP_FACE_A) {
            const vec2<real> normal = rot_mul(xa, axis);
            const vec2<real> point_a = xf_mul(xa, local_point);
            const vec2<real> axis_b = rot_mul_t(xb, -normal);
            index_a = -1;
            index_b = proxy_support(pb, axis_b);
            const vec2<real> point_b = xf_mul(xb, pb.vert(index_b));
            return dot(point_b - point_a, normal);
        } else {
            const vec2<real> normal = rot_mul(xb, axis);
            const vec2<real> point_b = xf_mul(xb, local_point);
            const vec2<real> axis_a = rot_mul_t(xa, -normal);
            index_b = -1;
            index_a = proxy_support(pa, axis_a);
            const vec2<real> point_a = xf_mul(xa, pa.vert(index_a));
            return dot(point_a - point_b, normal);
        }
    }
    B2B_HD_COLD real evaluate(int index_a, int index_b, real t) const                                        // toi.py:185-221
    {
        const xform<real> xa = sweep_xf(sa, t), xb = sweep_xf(sb, t);
        if (type == SEP_POINTS) {
            const vec2<real> point_a = xf_mul(xa, pa.vert(index_a)), point_b = xf_mul(xb, pb.vert(index_b));
            return dot(point_b - point_a, axis);
        } else if (type == SEP_FACE_A) {
            const vec2<real> normal = rot_mul(xa, axis);
            const vec2<real> point_a = xf_mul(xa, local_point);
            const vec2<real> point_b = xf_mul(xb, pb.vert(index_b));
            return dot(point_b - point_a, normal);
        } else {
            const vec2<real> normal = rot_mul(xb, axis);
            const vec2<real> point_b = xf_mul(xb, local_point);
            const vec2<real> point_a = xf_mul(xa, pa.vert(index_a));
            return dot(point_a - point_b, normal);
        }
    }
};

// time_of_impact: toi.py:223-391.  `stuck` is raised when the (unbounded in the reference) push-back loop hit TOI_MAX_PUSH_BACK.
template <class real>
B2B_HD_COLD int time_of_impact(const shape_view<real>& pa, const shape_view<real>& pb, sweep_t<real> sweep_a,
                               sweep_t<real> sweep_b, real t_max, real& out_t, bool& stuck)
{
    int state = TOI_UNKNOWN;
    real output_t = t_max;
    sweep_normalize(sweep_a);
    sweep_normalize(sweep_b);
    const real total_radius = pa.radius + pb.radius;
    const real target = pymax(K<real>::linear_slop, total_radius - real(3) * K<real>::linear_slop);
    const real tolerance = real(0.25) * K<real>::linear_slop;
    real t1 = real(0);
    int iter = 0;
    simplex_cache<real> cache;
    cache.metric = real(0); cache.count = 0;
    for (int k = 0; k < 3; ++k) { cache.ia[k] = 0; cache.ib[k] = 0; }
    sep_function<real> fcn;
    fcn.pa = pa; fcn.pb = pb; fcn.sa = sweep_a; fcn.sb = sweep_b;
    for (;;) {
        const xform<real> xa = sweep_xf(sweep_a, t1), xb = sweep_xf(sweep_b, t1);
        const real out_distance = shape_distance(pa, pb, xa, xb, cache);
        if (out_distance <= real(0)) { state = TOI_OVERLAPPED; output_t = real(0); break; }
        if (out_distance < target + tolerance) { state = TOI_TOUCHING; output_t = t1; break; }
        fcn.initialize(cache, t1);
        bool done = false;
        real t2 = t_max;
        int push_back_iter = 0;
        for (;;) {
            int index_a, index_b;
            real s2 = fcn.find_min_separation(t2, index_a, index_b);
            if (s2 > (target + tolerance)) { state = TOI_SEPARATED; output_t = t_max; done = true; break; }
            if (s2 > (target - tolerance)) { t1 = t2; break; }
            real s1 = fcn.evaluate(index_a, index_b, t1);
            if (s1 < (target - tolerance)) { state = TOI_FAILED; output_t = t1; done = true; break; }
            if (s1 <= (target + tolerance)) { state = TOI_TOUCHING; output_t = t1; done = true; break; }
            int root_iter_count = 0;
            real a1 = t1, a2 = t2;
            for (;;) {
                real t;
                if (root_iter_count & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = real(0.5) * (a1 + a2);
                const real s = fcn.evaluate(index_a, index_b, t);
                if (abs_r(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                ++root_iter_count;
                if (root_iter_count == (int)TOI_MAX_ROOT_ITERS) break;
            }
            if (++push_back_iter == (int)TOI_MAX_PUSH_BACK) { stuck = true; state = TOI_FAILED; output_t = t1; done = true; break; }
        }
        ++iter;
        if (done) break;
        if (iter == (int)TOI_MAX_ITERS) { state = TOI_FAILED; output_t = t1; break; }
    }
    out_t = output_t;
    return state;
}

}  // namespace b2b
