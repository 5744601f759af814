// b2b_collide.cuh -- narrowphase manifold generation (kernel K1's per-pair work).
//
// Device restatement of the reference's five colliders; results must be identical to
//   collide_circles            pypybox2d/collision.py:46-66
//   collide_polygon_circle     collision.py:68-141
//   collide_polygons           collision.py:144-241 + collision_util.py:31-164
//   collide_edge_circle        collision.py:246-353
//   collide_edge_polygon       collision.py:355-690 (EPCollider)
// including the quirks listed in SURVEY A.9 (#1 polygon centroid is (0,0), #4 un-rotated offset of
// Transform.mul_t(Transform), #5 no early return in the polygon-axis loop, #6 id permutation).
//
// Shapes are rows of the shared shape table: meta = type | count<<8 | has_v0<<16 | has_v3<<17,
// data = polygon {8 vertices, 8 normals}, circle {center}, edge {v1, v2, v0, v3}.
#pragma once
#include "b2b_math.cuh"

namespace b2b {

template <class real> struct shape_view {
    int type, count, has0, has3;
    real radius;
    real aabb_radius;        // the margin of Shape.compute_aabb: the radius, or 0 for a child edge of a Loop (shapes.py:822-837)
    const real* d;
    B2B_HD vec2<real> vert(int i) const { return mk<real>(d[2 * i], d[2 * i + 1]); }
    B2B_HD vec2<real> norm(int i) const { return mk<real>(d[16 + 2 * i], d[16 + 2 * i + 1]); }
};

template <class real>
B2B_HD shape_view<real> shape_load(const int32_t* s_meta, const real* s_radius, const real* s_data, int row)
{
    shape_view<real> s;
    int m = s_meta[row];
    s.type = m & 0xff; s.count = (m >> 8) & 0xff; s.has0 = (m >> 16) & 1; s.has3 = (m >> 17) & 1;
    s.radius = s_radius[row];
    s.aabb_radius = ((m >> 18) & 1) ? real(0) : s.radius;
    s.d = s_data + (size_t)SHAPE_DATA * row;
    return s;
}

template <class real> struct manifold {
    int type, count;
    vec2<real> local_normal, local_point;
    vec2<real> pt[2];
    uint32_t id[2];
};

template <class real> struct clip_vertex {
    vec2<real> v;
    int ia, ib, ta, tb;
};

// ---------------------------------------------------------------- circles
template <class real>
B2B_HD_COLD void collide_circles(manifold<real>& m, const shape_view<real>& a, const xform<real>& xa,
                            const shape_view<real>& b, const xform<real>& xb)
{
    m.count = 0;
    vec2<real> ca = a.vert(0), cb = b.vert(0);
    vec2<real> d = xf_mul(xb, cb) - xf_mul(xa, ca);
    real rsum = a.radius + b.radius;
    if (length_sqr(d) > sq_r(rsum)) return;
    m.type = MANIFOLD_CIRCLES;
    m.local_point = ca;
    m.local_normal = mk<real>(real(0), real(0));
    m.count = 1;
    m.pt[0] = cb;
    m.id[0] = 0u;
}

// ---------------------------------------------------------------- polygon vs circle
template <class real>
B2B_HD_COLD void collide_polygon_circle(manifold<real>& m, const shape_view<real>& pa, const xform<real>& xa,
                                   const shape_view<real>& cb, const xform<real>& xb)
{
    m.count = 0;
    const vec2<real> center_b = cb.vert(0);
    const vec2<real> c_local = xf_mul_t(xa, xf_mul(xb, center_b));
    const real radius = pa.radius + cb.radius;
    const int n = pa.count;
    int best = 0;
    real separation = -K<real>::max_float;
    for (int i = 0; i < n; ++i) {
        real s = dot(pa.norm(i), c_local - pa.vert(i));
        if (s > radius) return;
        if (s > separation) { separation = s; best = i; }
    }
    const int i1 = best, i2 = (best + 1 < n) ? best + 1 : 0;
    const vec2<real> v1 = pa.vert(i1), v2 = pa.vert(i2);
    if (separation < K<real>::epsilon) {                       // center inside the polygon
        m.count = 1; m.type = MANIFOLD_FACE_A;
        m.local_normal = pa.norm(best);
        m.local_point = real(0.5) * (v1 + v2);
        m.pt[0] = center_b; m.id[0] = 0u;
        return;
    }
    const real u1 = dot(c_local - v1, v2 - v1);
    const real u2 = dot(c_local - v2, v1 - v2);
    if (u1 < real(0)) {
        vec2<real> d = c_local - v1;
        if (dot(d, d) > sq_r(radius)) return;
        m.local_normal = normalize(d);
        m.local_point = v1;
    } else if (u2 < real(0)) {
        vec2<real> d = c_local - v2;
        if (dot(d, d) > sq_r(radius)) return;
        m.local_normal = normalize(d);
        m.local_point = v2;
    } else {
        vec2<real> face_center = real(0.5) * (v1 + v2);
        if (dot(c_local - face_center, pa.norm(i1)) > radius) return;
        m.local_normal = pa.norm(i1);
        m.local_point = face_center;
    }
    m.count = 1; m.type = MANIFOLD_FACE_A;
    m.pt[0] = center_b; m.id[0] = 0u;
}

// ---------------------------------------------------------------- polygon vs polygon (SAT + clip)
// separation of poly2 from edge `e1` of poly1 (collision_util.py:31-52)
// Inlining attributes of the two SAT helpers (experiment switches, tools/build_variant.sh).  The edge separation is OUT of line:
// the hill climb calls it from four places and the narrowphase kernel is bound by instruction fetch; one copy instead of
// eight (4,056 instead of 5,344 SASS instructions) measured 7.36 -> 7.02 ms per 131,072 pyramid worlds.  The hill climb itself
// out of line loses (7.99 ms), both out of line is neutral (7.37 ms).
#ifndef B2B_SAT_EDGE_ATTR
#define B2B_SAT_EDGE_ATTR B2B_HD_COLD
#endif
#ifndef B2B_SAT_MAX_ATTR
#define B2B_SAT_MAX_ATTR B2B_HD
#endif
template <class real>
B2B_SAT_EDGE_ATTR real sat_edge_separation(const shape_view<real>& p1, const xform<real>& x1, int e1,
                                const shape_view<real>& p2, const xform<real>& x2)
{
    const vec2<real> n_world = rot_mul(x1, p1.norm(e1));
    const vec2<real> n_in2 = rot_mul_t(x2, n_world);
    int support = 0;
    real best = dot(p2.vert(0), n_in2);
    for (int i = 1; i < p2.count; ++i) {
        real d = dot(p2.vert(i), n_in2);
        if (d < best) { best = d; support = i; }
    }
    return dot(xf_mul(x2, p2.vert(support)) - xf_mul(x1, p1.vert(e1)), n_world);
}

// hill-climb over poly1's edges from the edge facing poly2 (collision_util.py:54-109)
template <class real>
B2B_SAT_MAX_ATTR real sat_max_separation(int& edge_out, const shape_view<real>& p1, const xform<real>& x1,
                               const shape_view<real>& p2, const xform<real>& x2)
{
    const int n1 = p1.count;
    const vec2<real> zero = mk<real>(real(0), real(0));                   // Polygon._centroid, A.9 #1
    const vec2<real> d_local = rot_mul_t(x1, xf_mul(x2, zero) - xf_mul(x1, zero));
    int edge = 0;
    real best_dot = dot(p1.norm(0), d_local);
    for (int i = 1; i < n1; ++i) {
        real d = dot(p1.norm(i), d_local);
        if (d > best_dot) { best_dot = d; edge = i; }
    }
    real s = sat_edge_separation(p1, x1, edge, p2, x2);
    const int prev = (edge == 0) ? n1 - 1 : edge - 1;
    const real s_prev = sat_edge_separation(p1, x1, prev, p2, x2);
    const int next = (edge + 1 == n1) ? 0 : edge + 1;
    const real s_next = sat_edge_separation(p1, x1, next, p2, x2);
    int best_edge, step;
    real best_sep;
    if (s_prev > s && s_prev > s_next) { step = -1; best_edge = prev; best_sep = s_prev; }
    else if (s_next > s) { step = 1; best_edge = next; best_sep = s_next; }
    else { edge_out = edge; return s; }
    for (;;) {
        edge = (step < 0) ? ((best_edge == 0) ? n1 - 1 : best_edge - 1) : ((best_edge + 1 == n1) ? 0 : best_edge + 1);
        s = sat_edge_separation(p1, x1, edge, p2, x2);
        if (s > best_sep) { best_edge = edge; best_sep = s; }
        else break;
    }
    edge_out = best_edge;
    return best_sep;
}

// Sutherland-Hodgman step (collision_util.py:135-164); returns the number of output vertices (<= 2 here
// because two points on the inside never straddle the plane).
template <class real>
B2B_HD int clip_segment(clip_vertex<real> out[2], const clip_vertex<real> in[2], vec2<real> normal, real offset,
                        int vertex_index_a)
{
    int n = 0;
    const real d0 = dot(normal, in[0].v) - offset;
    const real d1 = dot(normal, in[1].v) - offset;
    if (d0 <= real(0)) out[n++] = in[0];
    if (d1 <= real(0)) out[n++] = in[1];
    if (d0 * d1 < real(0)) {
        const real t = d0 / (d0 - d1);
        clip_vertex<real> cv;
        cv.v = in[0].v + t * (in[1].v - in[0].v);
        cv.ia = vertex_index_a; cv.ib = in[0].ib; cv.ta = MP_VERTEX; cv.tb = MP_FACE;
        if (n < 2) out[n] = cv;
        ++n;
    }
    return n;
}

template <class real>
B2B_HD void collide_polygons(manifold<real>& m, const shape_view<real>& pa, const xform<real>& xa,
                             const shape_view<real>& pb, const xform<real>& xb)
{
    m.count = 0;
    const real total_radius = pa.radius + pb.radius;
    int edge_a, edge_b;
    const real sep_a = sat_max_separation(edge_a, pa, xa, pb, xb);
    if (sep_a > total_radius) return;
    const real sep_b = sat_max_separation(edge_b, pb, xb, pa, xa);
    if (sep_b > total_radius) return;

    const bool flip = sep_b > (K<real>::relative_tol * sep_a + K<real>::absolute_tol);
    const shape_view<real>& ref = flip ? pb : pa;       // reference polygon (poly1)
    const shape_view<real>& inc = flip ? pa : pb;       // incident polygon (poly2)
    const xform<real>& xr = flip ? xb : xa;
    const xform<real>& xi = flip ? xa : xb;
    const int edge1 = flip ? edge_b : edge_a;
    m.type = flip ? MANIFOLD_FACE_B : MANIFOLD_FACE_A;

    // incident edge: the edge of `inc` most anti-parallel to the reference normal (collision_util.py:111-133)
    clip_vertex<real> incident[2];
    {
        const vec2<real> n_in_inc = rot_mul_t(xi, rot_mul(xr, ref.norm(edge1)));
        int idx = 0;
        real best = dot(n_in_inc, inc.norm(0));
        for (int i = 1; i < inc.count; ++i) {
            real d = dot(n_in_inc, inc.norm(i));
            if (d < best) { best = d; idx = i; }
        }
        const int i2 = (idx + 1 == inc.count) ? 0 : idx + 1;
        incident[0].v = xf_mul(xi, inc.vert(idx)); incident[0].ia = edge1; incident[0].ib = idx;
        incident[0].ta = MP_FACE; incident[0].tb = MP_VERTEX;
        incident[1].v = xf_mul(xi, inc.vert(i2)); incident[1].ia = edge1; incident[1].ib = i2;
        incident[1].ta = MP_FACE; incident[1].tb = MP_VERTEX;
    }
    const int iv1 = edge1, iv2 = (edge1 + 1 == ref.count) ? 0 : edge1 + 1;
    const vec2<real> v11_l = ref.vert(iv1), v12_l = ref.vert(iv2);
    const vec2<real> local_tangent = normalize(v12_l - v11_l);
    const vec2<real> tangent = rot_mul(xr, local_tangent);
    const vec2<real> normal = right_perp(tangent);
    const vec2<real> v11 = xf_mul(xr, v11_l), v12 = xf_mul(xr, v12_l);
    const real front_offset = dot(normal, v11);
    const real side1 = -dot(tangent, v11) + total_radius;
    const real side2 = dot(tangent, v12) + total_radius;
    clip_vertex<real> c1[2], c2[2];
    if (clip_segment(c1, incident, -tangent, side1, iv1) < 2) return;
    if (clip_segment(c2, c1, tangent, side2, iv2) < 2) return;
    m.local_normal = right_perp(local_tangent);
    m.local_point = real(0.5) * (v11_l + v12_l);
    int count = 0;
    for (int i = 0; i < 2; ++i) {
        const real sep = dot(normal, c2[i].v) - front_offset;
        if (sep <= total_radius) {
            m.pt[count] = xf_mul_t(xi, c2[i].v);
            m.id[count] = flip ? make_id(c2[i].ib, c2[i].ia, c2[i].tb, c2[i].ta)
                               : make_id(c2[i].ia, c2[i].ib, c2[i].ta, c2[i].tb);
            ++count;
        }
    }
    m.count = count;
}

// ---------------------------------------------------------------- edge vs circle
template <class real>
B2B_HD_COLD void collide_edge_circle(manifold<real>& m, const shape_view<real>& ea, const xform<real>& xa,
                                const shape_view<real>& cb, const xform<real>& xb)
{
    m.count = 0;
    const vec2<real> center_b = cb.vert(0);
    const vec2<real> Q = xf_mul_t(xa, xf_mul(xb, center_b));
    const vec2<real> A = ea.vert(0), B = ea.vert(1);
    const vec2<real> e = B - A;
    const real u = dot(e, B - Q);
    const real v = dot(e, Q - A);
    const real radius = ea.radius + cb.radius;
    if (v <= real(0)) {                                         // vertex region A
        vec2<real> d = Q - A;
        if (dot(d, d) > sq_r(radius)) return;
        if (ea.has0) {
            vec2<real> A1 = ea.vert(2);
            if (dot(A - A1, A - Q) > real(0)) return;           // belongs to the previous edge
        }
        m.count = 1; m.type = MANIFOLD_CIRCLES; m.local_normal = mk<real>(real(0), real(0)); m.local_point = A;
        m.id[0] = make_id(0, 0, MP_VERTEX, MP_VERTEX); m.pt[0] = center_b;
        return;
    }
    if (u <= real(0)) {                                         // vertex region B
        vec2<real> d = Q - B;
        if (dot(d, d) > sq_r(radius)) return;
        if (ea.has3) {
            vec2<real> B2 = ea.vert(3);
            if (dot(B2 - B, Q - B) > real(0)) return;           // belongs to the next edge
        }
        m.count = 1; m.type = MANIFOLD_CIRCLES; m.local_normal = mk<real>(real(0), real(0)); m.local_point = B;
        m.id[0] = make_id(1, 0, MP_VERTEX, MP_VERTEX); m.pt[0] = center_b;
        return;
    }
    const real den = dot(e, e);                                 // face region
    const vec2<real> P = (real(1) / den) * (u * A + v * B);
    const vec2<real> d = Q - P;
    if (dot(d, d) > sq_r(radius)) return;
    vec2<real> n = mk<real>(-e.y, e.x);
    if (dot(n, Q - A) < real(0)) n = mk<real>(-n.x, -n.y);
    m.count = 1; m.type = MANIFOLD_FACE_A; m.local_normal = normalize(n); m.local_point = A;
    m.id[0] = make_id(0, 0, MP_FACE, MP_VERTEX); m.pt[0] = center_b;
}

// ---------------------------------------------------------------- edge vs polygon
// Polygon vertices/normals are expressed in the edge's frame on demand (no local arrays -> no local memory).
template <class real> struct ep_ctx {
    xform<real> xf;                 // "edge frame <- polygon frame" as the reference composes it (A.9 #4)
    vec2<real> v1, v2, enormal;
    vec2<real> l11, l12, l21, l22;  // admissible normal ranges from adjacency
    real radius;
    const shape_view<real>* poly;
    B2B_HD vec2<real> bvert(int i) const { return xf_mul(xf, poly->vert(i)); }
    B2B_HD vec2<real> bnorm(int i) const { return rot_mul(xf, poly->norm(i)); }
    B2B_HD vec2<real> avert(int i) const { return i == 0 ? v1 : v2; }
    B2B_HD vec2<real> anorm(int i) const { return i == 0 ? enormal : -enormal; }
    B2B_HD bool admissible(vec2<real> n) const
    {
        const real slop = -K<real>::angular_slop;
        return (cross(n, l11) >= slop) && (cross(l12, n) >= slop) && (cross(n, l21) >= slop) && (cross(l22, n) >= slop);
    }
};

template <class real>
B2B_HD_COLD void collide_edge_polygon(manifold<real>& m, const shape_view<real>& ea, const xform<real>& xa,
                                 const shape_view<real>& pb, const xform<real>& xb)
{
    ep_ctx<real> c;
    c.xf.p = xb.p - xa.p;
    c.xf.c1 = mk<real>(dot(xa.c1, xb.c1), dot(xa.c2, xb.c1));
    c.xf.c2 = mk<real>(dot(xa.c1, xb.c2), dot(xa.c2, xb.c2));
    c.v1 = ea.vert(0); c.v2 = ea.vert(1);
    {
        vec2<real> e = c.v2 - c.v1;
        c.enormal = normalize(mk<real>(e.y, -e.x));
    }
    c.poly = &pb;
    c.radius = real(2) * K<real>::polygon_radius;
    const vec2<real> zero = mk<real>(real(0), real(0));
    c.l11 = c.l12 = c.l21 = c.l22 = zero;
    m.count = 0;

    // adjacency limits (collision.py:537-597)
    const vec2<real> center_b = xf_mul(c.xf, zero);
    if (ea.has0) {
        const vec2<real> v0 = ea.vert(2);
        const vec2<real> e0 = c.v1 - v0, e1 = c.v2 - c.v1;
        const vec2<real> n0 = normalize(mk<real>(e0.y, -e0.x)), n1 = normalize(mk<real>(e1.y, -e1.x));
        const bool convex = cross(n0, n1) >= real(0);
        const bool front0 = dot(n0, center_b - v0) >= real(0);
        const bool front1 = dot(n1, center_b - c.v1) >= real(0);
        if (convex) {
            if (front0 || front1) { c.l11 = n1; c.l12 = n0; } else { c.l11 = -n1; c.l12 = -n0; }
        } else {
            if (front0 && front1) { c.l11 = n0; c.l12 = n1; } else { c.l11 = -n0; c.l12 = -n1; }
        }
    }
    if (ea.has3) {
        const vec2<real> v3 = ea.vert(3);
        const vec2<real> e1 = c.v2 - c.v1, e2 = v3 - c.v2;
        const vec2<real> n1 = normalize(mk<real>(e1.y, -e1.x)), n2 = normalize(mk<real>(e2.y, -e2.x));
        const bool convex = cross(n1, n2) >= real(0);
        const bool front1 = dot(n1, center_b - c.v1) >= real(0);
        const bool front2 = dot(n2, center_b - c.v2) >= real(0);
        if (convex) {
            if (front1 || front2) { c.l21 = n2; c.l22 = n1; } else { c.l21 = -n2; c.l22 = -n1; }
        } else {
            if (front1 && front2) { c.l21 = n1; c.l22 = n2; } else { c.l21 = -n1; c.l22 = -n2; }
        }
    }

    // edge axis (collision.py:599-629): first admissible normal that separates wins immediately
    int e_type = 0, e_index = -1;
    real e_sep = -K<real>::max_float;
    for (int i = 0; i < 2; ++i) {
        const vec2<real> n = c.anorm(i);
        if (!c.admissible(n)) continue;
        real s = dot(n, c.bvert(0) - c.v1);
        for (int k = 1; k < pb.count; ++k) {
            real d = dot(n, c.bvert(k) - c.v1);
            if (d < s) s = d;
        }
        if (s > c.radius) { e_type = 1; e_index = i; e_sep = s; break; }
        if (s > e_sep) { e_type = 1; e_index = i; e_sep = s; }
    }
    if (e_type == 0) return;
    if (e_sep > c.radius) return;

    // polygon axis (collision.py:631-664): keeps scanning after a separating axis (A.9 #5)
    int p_type = 0, p_index = -1;
    real p_sep = -K<real>::max_float;
    for (int i = 0; i < pb.count; ++i) {
        const vec2<real> n = -c.bnorm(i);
        if (!c.admissible(n)) continue;
        const vec2<real> bv = c.bvert(i);
        const real s = pymin(dot(n, bv - c.v1), dot(n, bv - c.v2));
        if (s > c.radius) { p_type = 2; p_index = i; p_sep = s; }
        if (s > p_sep) { p_type = 2; p_index = i; p_sep = s; }
    }
    if (p_type != 0 && p_sep > c.radius) return;

    const bool use_poly = (p_type != 0) && (p_sep > K<real>::relative_tol * e_sep + K<real>::absolute_tol);
    const int edge1 = use_poly ? p_index : e_index;
    m.type = use_poly ? MANIFOLD_FACE_B : MANIFOLD_FACE_A;
    const int n1 = use_poly ? pb.count : 2;
    const int n2 = use_poly ? 2 : pb.count;

    // incident edge on the other proxy (collision.py:666-690)
    clip_vertex<real> incident[2];
    {
        const vec2<real> nref = use_poly ? c.bnorm(edge1) : c.anorm(edge1);
        int idx = 0;
        real best = dot(nref, use_poly ? c.anorm(0) : c.bnorm(0));
        for (int i = 1; i < n2; ++i) {
            real d = dot(nref, use_poly ? c.anorm(i) : c.bnorm(i));
            if (d < best) { best = d; idx = i; }
        }
        const int i2 = (idx + 1 == n2) ? 0 : idx + 1;
        incident[0].v = use_poly ? c.avert(idx) : c.bvert(idx);
        incident[0].ia = edge1; incident[0].ib = idx; incident[0].ta = MP_FACE; incident[0].tb = MP_VERTEX;
        incident[1].v = use_poly ? c.avert(i2) : c.bvert(i2);
        incident[1].ia = edge1; incident[1].ib = i2; incident[1].ta = MP_FACE; incident[1].tb = MP_VERTEX;
    }
    const int iv1 = edge1, iv2 = (edge1 + 1 == n1) ? 0 : edge1 + 1;
    const vec2<real> v11 = use_poly ? c.bvert(iv1) : c.avert(iv1);
    const vec2<real> v12 = use_poly ? c.bvert(iv2) : c.avert(iv2);
    const vec2<real> tangent = normalize(v12 - v11);
    const vec2<real> normal = right_perp(tangent);
    const vec2<real> plane_point = real(0.5) * (v11 + v12);
    const real front_offset = dot(normal, v11);
    const real side1 = -dot(tangent, v11) + c.radius;
    const real side2 = dot(tangent, v12) + c.radius;
    clip_vertex<real> c1[2], c2[2];
    if (clip_segment(c1, incident, -tangent, side1, iv1) < 2) return;
    if (clip_segment(c2, c1, tangent, side2, iv2) < 2) return;
    if (!use_poly) {
        m.local_normal = normal;
        m.local_point = plane_point;
    } else {
        m.local_normal = rot_mul_t(c.xf, normal);
        m.local_point = xf_mul_t(c.xf, plane_point);
    }
    int count = 0;
    for (int i = 0; i < 2; ++i) {
        const real sep = dot(normal, c2[i].v) - front_offset;
        if (sep <= c.radius) {
            if (!use_poly) {
                m.pt[count] = xf_mul_t(c.xf, c2[i].v);
                m.id[count] = make_id(c2[i].ia, c2[i].ib, c2[i].ta, c2[i].tb);
            } else {
                m.pt[count] = c2[i].v;
                m.id[count] = make_id(c2[i].tb, c2[i].ta, c2[i].ib, c2[i].ia);   // A.9 #6
            }
            ++count;
        }
    }
    m.count = count;
}

// dispatch in registry order (contact.py:294-311): A/B already ordered by contact creation
template <class real>
B2B_HD void collide_shapes(manifold<real>& m, const shape_view<real>& a, const xform<real>& xa,
                           const shape_view<real>& b, const xform<real>& xb)
{
    if (a.type == SHAPE_POLYGON && b.type == SHAPE_POLYGON) collide_polygons(m, a, xa, b, xb);
    else if (a.type == SHAPE_EDGE && b.type == SHAPE_POLYGON) collide_edge_polygon(m, a, xa, b, xb);
    else if (a.type == SHAPE_CIRCLE && b.type == SHAPE_CIRCLE) collide_circles(m, a, xa, b, xb);
    else if (a.type == SHAPE_POLYGON && b.type == SHAPE_CIRCLE) collide_polygon_circle(m, a, xa, b, xb);
    else if (a.type == SHAPE_EDGE && b.type == SHAPE_CIRCLE) collide_edge_circle(m, a, xa, b, xb);
    else m.count = 0;
}

// Shape.compute_aabb (shapes.py:178-189 circle, 316-334 polygon, 702-718 edge, 822-837 a Loop's child edge) -> (lx, ly, ux, uy)
template <class real> B2B_HD void shape_aabb(const shape_view<real>& s, const xform<real>& t, real out[4])
{
    vec2<real> lo, hi;
    if (s.type == SHAPE_CIRCLE) {
        lo = hi = t.p + rot_mul(t, s.vert(0));
    } else {
        const int n = (s.type == SHAPE_POLYGON) ? s.count : 2;
        lo = hi = xf_mul(t, s.vert(0));
        for (int i = 1; i < n; ++i) {
            vec2<real> v = xf_mul(t, s.vert(i));
            lo = mk<real>(pymin(lo.x, v.x), pymin(lo.y, v.y));
            hi = mk<real>(pymax(hi.x, v.x), pymax(hi.y, v.y));
        }
    }
    if (s.type == SHAPE_EDGE && s.aabb_radius == real(0)) { out[0] = lo.x; out[1] = lo.y; out[2] = hi.x; out[3] = hi.y; return; }   // Loop.compute_aabb
    out[0] = lo.x - s.radius; out[1] = lo.y - s.radius; out[2] = hi.x + s.radius; out[3] = hi.y + s.radius;
}

}  // namespace b2b
