// b2b_query.cuh -- the callers either side of the step: World.query_aabb, World.ray_cast, Fixture.test_point.
//
// Reference: World.query_aabb / ray_cast pypybox2d/world.py:371-440; DynamicTree.query / ray_cast dynamictree.py:190-262
// (the tree only prunes: a proxy is reported iff ITS fat box passes the tests a leaf has to pass, so the answer is a scan
// over the world's proxies -- a few hundred 16-byte boxes, one coalesced pass of a warp -- instead of a walk of a tree the
// device does not keep); Shape.test_point / ray_cast shapes.py:141-176 (circle), 256-314 (polygon), 655-700 (edge).
// The reference reports hits in tree order; here they come out in ascending fixture order (the SET is the reference's; Box2D
// makes no promise about callback order).  One warp per query, ballot compaction, results into caller-owned device arrays.
#pragma once
#include "b2b_collide.cuh"

namespace b2b {

// Shape.ray_cast (max_fraction = 1: World.ray_cast always passes 1.0, world.py:425)
template <class real>
B2B_HD bool shape_ray_cast(const shape_view<real>& s, const xform<real>& xf, vec2<real> p1, vec2<real> p2, vec2<real>& normal_out,
                           real& fraction_out)
{
    const real max_fraction = real(1);
    if (s.type == SHAPE_CIRCLE) {
        const vec2<real> position = xf.p + rot_mul(xf, s.vert(0));
        const vec2<real> sv = p1 - position;
        const real b = dot(sv, sv) - sq_r(s.radius);
        const vec2<real> r = p2 - p1;
        const real c = dot(sv, r), rr = dot(r, r);
        const real sigma = c * c - rr * b;
        if (sigma < real(0) || rr < K<real>::epsilon) return false;
        real a = -(c + sqrt_r(sigma));
        if (real(0) <= a && a <= max_fraction * rr) {
            a /= rr;
            fraction_out = a;
            normal_out = normalize(sv + a * r);
            return true;
        }
        return false;
    }
    const vec2<real> q1 = rot_mul_t(xf, p1 - xf.p), q2 = rot_mul_t(xf, p2 - xf.p);
    const vec2<real> d = q2 - q1;
    if (s.type == SHAPE_POLYGON) {
        real lower = real(0), upper = max_fraction;
        int tagged = -1;
        for (int i = 0; i < s.count; ++i) {
            const vec2<real> normal = s.norm(i), vertex = s.vert(i);
            const real numerator = dot(normal, vertex - q1);
            const real denominator = dot(normal, d);
            if (denominator == real(0)) {
                if (numerator < real(0)) return false;
            } else {
                if (denominator < real(0) && numerator < lower * denominator) { lower = numerator / denominator; tagged = i; }
                else if (denominator > real(0) && numerator < upper * denominator) upper = numerator / denominator;
            }
            if (upper < lower) return false;
        }
        if (tagged < 0) return false;
        normal_out = rot_mul(xf, s.norm(tagged));
        fraction_out = lower;
        return true;
    }
    // edge (shapes.py:655-700; the port returns the normal in the edge's LOCAL frame)
    const vec2<real> v1 = s.vert(0), v2 = s.vert(1);
    const vec2<real> e = v2 - v1;
    vec2<real> normal = normalize(mk<real>(e.y, -e.x));
    const real numerator = dot(normal, v1 - q1);
    const real denominator = dot(normal, d);
    if (denominator == real(0)) return false;
    const real t = numerator / denominator;
    if (t < real(0) || real(1) < t) return false;
    const vec2<real> q = q1 + t * d;
    const vec2<real> r = v2 - v1;
    const real rr = dot(r, r);
    if (rr == real(0)) return false;
    const real sp = dot(q - v1, r) / rr;
    if (sp < real(0) || real(1) < sp) return false;
    fraction_out = t;
    if (numerator > real(0)) normal = -normal;
    normal_out = normal;
    return true;
}

// what DynamicTree.ray_cast asks of a leaf box before it reports it (dynamictree.py:226-246), un-clipped segment
template <class real> B2B_HD bool ray_leaf_test(const real* box, vec2<real> p1, vec2<real> p2)
{
    const vec2<real> r = normalize(p2 - p1);
    const vec2<real> v = cross_sv(real(1), r);
    const vec2<real> abs_v = mk<real>(abs_r(v.x), abs_r(v.y));
    const vec2<real> t = p1 + real(1) * (p2 - p1);
    const real seg[4] = {pymin(p1.x, t.x), pymin(p1.y, t.y), pymax(p1.x, t.x), pymax(p1.y, t.y)};
    if (!aabb_overlap<real>(box, seg)) return false;
    const vec2<real> c = real(0.5) * (mk<real>(box[0], box[1]) + mk<real>(box[2], box[3]));
    const vec2<real> h = real(0.5) * (mk<real>(box[2], box[3]) - mk<real>(box[0], box[1]));
    const real separation = abs_r(dot(v, p1 - c)) - dot(abs_v, h);
    return !(separation > real(0));
}

// Fixture.test_point (an edge has no interior)
template <class real> B2B_HD bool shape_test_point(const shape_view<real>& s, const xform<real>& xf, vec2<real> p)
{
    if (s.type == SHAPE_CIRCLE) {
        const vec2<real> d = p - (xf.p + rot_mul(xf, s.vert(0)));
        return dot(d, d) <= sq_r(s.radius);
    }
    if (s.type == SHAPE_POLYGON) {
        const vec2<real> pl = rot_mul_t(xf, p - xf.p);
        for (int i = 0; i < s.count; ++i)
            if (dot(s.norm(i), pl - s.vert(i)) > real(0)) return false;
        return true;
    }
    return false;
}

}  // namespace b2b
