/*
 * b2_oracle.c -- TEST INFRASTRUCTURE ONLY (see b2_oracle.h).
 *
 * Sequential fp64 restatement of pypybox2d's World.step with continuous_physics=False and
 * pair-emission order pinned (sorted by proxy index).  Every function cites the reference
 * lines it follows; expressions keep the reference's operand order so that, built with
 * -ffp-contract=off against the same libm, results are bit-identical to the pure-Python
 * reference (x ** 2 is written pow(x, 2.0) exactly where the reference writes it).
 *
 * Not a product path: nothing in pypybox2d_b200/ may load this library.
 */
#include "b2_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: pypybox2d/settings.py:32-134 ---- */
#define EPS            DBL_EPSILON
#define EPS_SQR        pow(DBL_EPSILON, 2.0)        /* settings.py:40 */
#define MAX_FLOAT      DBL_MAX
#define PI_            3.141592653589793
#define AABB_EXT       0.1
#define AABB_MULT      2.0
#define LINEAR_SLOP    0.005
#define ANGULAR_SLOP   (2.0 / 180.0 * PI_)
#define POLY_RADIUS    (2.0 * LINEAR_SLOP)
#define VEL_THRESHOLD  1.0
#define MAX_LIN_CORR   0.2
#define MAX_TRANS      2.0
#define MAX_TRANS_SQR  pow(MAX_TRANS, 2.0)          /* settings.py:99: MAX_TRANSLATION**2 */
#define MAX_ROT        (0.5 * PI_)
#define MAX_ROT_SQR    pow(MAX_ROT, 2.0)            /* settings.py:104 */
#define BAUMGARTE      0.2
#define TIME_TO_SLEEP  0.5
#define LIN_SLEEP_TOL_SQR pow(0.01, 2.0)             /* settings.py:119 */
#define ANG_SLEEP_TOL_SQR pow(2.0 / 180.0 * PI_, 2.0) /* settings.py:123 */
#define REL_TOL        0.98
#define ABS_TOL        0.001

/* flag bits: pypybox2d_b200/schema.py */
#define BODY_TYPE_MASK 0x3u
#define BODY_AWAKE     0x4u
#define BODY_ACTIVE    0x8u
#define BODY_ALLOW_SLEEP 0x10u
#define BODY_ISLAND    0x80u
#define T_STATIC 0
#define T_DYNAMIC 1
#define C_TOUCHING 0x1u
#define C_ENABLED  0x2u
#define C_FILTER   0x4u
#define C_ISLAND   0x8u
#define C_SENSOR   0x10u      /* one of the contact's fixtures is a sensor (fixture.py:163-168); set when the contact is created */
#define W_ALLOW_SLEEP 0x1u
#define W_WARM_START  0x2u
#define W_NEW_FIXTURE 0x4u
#define W_AUTO_CLEAR  0x8u
#define ST_CONTACT_OVERFLOW 0x1u
#define ST_PAIR_OVERFLOW    0x2u
#define ST_ISLAND_OVERFLOW  0x8u
#define M_CIRCLES 0
#define M_FACE_A  1
#define M_FACE_B  2
#define S_CIRCLE 0
#define S_POLYGON 1
#define S_EDGE 3
#define MP_VERTEX 0
#define MP_FACE 1
#define MAXV 8

/* Python's min/max return the first argument unless the second compares strictly better */
#define PYMIN(a, b) (((b) < (a)) ? (b) : (a))
#define PYMAX(a, b) (((b) > (a)) ? (b) : (a))

typedef struct { double x, y; } v2;
typedef struct { v2 p, c1, c2; } xform;   /* position, rotation columns (common.py:360-503,616-685) */

static inline v2 V(double x, double y) { v2 r = {x, y}; return r; }
static inline v2 vadd(v2 a, v2 b) { return V(a.x + b.x, a.y + b.y); }
static inline v2 vsub(v2 a, v2 b) { return V(a.x - b.x, a.y - b.y); }
static inline v2 vneg(v2 a) { return V(-a.x, -a.y); }
static inline v2 vscale(double s, v2 a) { return V(s * a.x, s * a.y); }          /* common.py:113-115 */
static inline double vdot(v2 a, v2 b) { return a.x * b.x + a.y * b.y; }          /* common.py:213-215 */
static inline double vcross(v2 a, v2 b) { return a.x * b.y - a.y * b.x; }        /* common.py:210-212 */
static inline v2 vcross_s(v2 a, double s) { return V(s * a.y, -s * a.x); }       /* common.py:204-208 */
static inline v2 scross(double s, v2 a) { return V(-s * a.y, s * a.x); }         /* common.py:195-202 */
static inline double sq(double x) { return pow(x, 2.0); }                        /* Python's x ** 2 */
static inline double vlen2(v2 a) { return sq(a.x) + sq(a.y); }                   /* common.py:189-190 */
static inline v2 vnorm(v2 a)                                                      /* common.py:216-222 */
{
    double len = sqrt(sq(a.x) + sq(a.y));
    if (len < EPS) return a;
    double inv = 1.0 / len;
    return V(a.x * inv, a.y * inv);
}
static inline double dist2(v2 a, v2 b) { v2 c = vsub(a, b); return vdot(c, c); } /* common.py:950-952 */
static inline double pyclamp(double v, double lo, double hi)                      /* common.py:978-979 */
{
    double m = PYMIN(v, hi);
    return PYMAX(lo, m);
}

static inline xform xf_make(v2 p, double c, double s)
{
    xform t; t.p = p; t.c1 = V(c, s); t.c2 = V(-s, c); return t;
}
static inline xform xf_row(const double* r) { return xf_make(V(r[0], r[1]), r[2], r[3]); }
/* Transform(angle=a); position = c - R * local_center  (body.py:906-908, contact.py:433-436) */
static inline xform xf_from_sweep(v2 c, double a, v2 lc)
{
    xform t = xf_make(V(0, 0), cos(a), sin(a));
    v2 rl = V(t.c1.x * lc.x + t.c2.x * lc.y, t.c1.y * lc.x + t.c2.y * lc.y);
    t.p = vsub(c, rl);
    return t;
}
static inline v2 xf_mul(const xform* t, v2 v)                                     /* common.py:677-685 */
{
    return V(t->p.x + t->c1.x * v.x + t->c2.x * v.y, t->p.y + t->c1.y * v.x + t->c2.y * v.y);
}
static inline v2 rot_mul(const xform* t, v2 v)                                    /* common.py:404-411 */
{
    return V(t->c1.x * v.x + t->c2.x * v.y, t->c1.y * v.x + t->c2.y * v.y);
}
static inline v2 rot_mul_t(const xform* t, v2 v)                                  /* common.py:467-472 */
{
    return V(vdot(t->c1, v), vdot(t->c2, v));
}
static inline v2 xf_mul_t(const xform* t, v2 v) { return rot_mul_t(t, vsub(v, t->p)); } /* common.py:669-670 */

/* ---- shape table access ---- */
typedef struct { int type, count, has0, has3, loop_child; double radius; const double* d; } shape_t;
static inline shape_t shape_get(const orc_world* w, int row)
{
    shape_t s; int m = w->s_meta[row];
    s.type = m & 0xff; s.count = (m >> 8) & 0xff; s.has0 = (m >> 16) & 1; s.has3 = (m >> 17) & 1;
    s.loop_child = (m >> 18) & 1;           /* a child edge of a Loop shape (contact.py:363-373): Loop.compute_aabb has no radius margin */
    s.radius = w->s_radius[row]; s.d = w->s_data + 32 * (size_t)row;
    return s;
}
static inline v2 pvert(const shape_t* s, int i) { return V(s->d[2 * i], s->d[2 * i + 1]); }
static inline v2 pnorm(const shape_t* s, int i) { return V(s->d[16 + 2 * i], s->d[16 + 2 * i + 1]); }

/* ---- manifold working copy (contact_util.py:50-166) ---- */
typedef struct {
    int type, count;
    v2 local_normal, local_point;
    v2 pt[2];
    uint32_t id[2];
} manifold_t;
static inline uint32_t mkid(int a, int b, int c, int d)
{
    return (uint32_t)(a & 0xff) | ((uint32_t)(b & 0xff) << 8) | ((uint32_t)(c & 0xff) << 16) | ((uint32_t)(d & 0xff) << 24);
}
typedef struct { v2 v; int ia, ib, ta, tb; } clipv;

/* ---- collide_circles: collision.py:46-66 ---- */
static void collide_circles(manifold_t* m, const shape_t* a, const xform* xa, const shape_t* b, const xform* xb)
{
    m->count = 0;
    v2 posa = V(a->d[0], a->d[1]), posb = V(b->d[0], b->d[1]);
    v2 pa = xf_mul(xa, posa), pb = xf_mul(xb, posb);
    v2 d = vsub(pb, pa);
    double dist_sqr = vlen2(d);
    double radius = a->radius + b->radius;
    if (dist_sqr > sq(radius)) return;
    m->type = M_CIRCLES;
    m->local_point = posa;
    m->local_normal = V(0, 0);
    m->count = 1;
    m->pt[0] = posb;
    m->id[0] = mkid(0, 0, 0, 0);
}

/* ---- collide_polygon_circle: collision.py:68-141 ---- */
static void collide_polygon_circle(manifold_t* m, const shape_t* pa, const xform* xa, const shape_t* cb, const xform* xb)
{
    m->count = 0;
    v2 posb = V(cb->d[0], cb->d[1]);
    v2 c = xf_mul(xb, posb);
    v2 c_local = xf_mul_t(xa, c);
    int normal_index = 0;
    double separation = -MAX_FLOAT;
    double radius = pa->radius + cb->radius;
    int n = pa->count;
    for (int i = 0; i < n; ++i) {
        double s = vdot(pnorm(pa, i), vsub(c_local, pvert(pa, i)));
        if (s > radius) return;
        if (s > separation) { separation = s; normal_index = i; }
    }
    int i1 = normal_index, i2 = (i1 + 1) % n;
    v2 v1 = pvert(pa, i1), v2_ = pvert(pa, i2);
    if (separation < EPS) {
        m->count = 1; m->type = M_FACE_A;
        m->local_normal = pnorm(pa, normal_index);
        m->local_point = vscale(0.5, vadd(v1, v2_));
        m->pt[0] = posb; m->id[0] = mkid(0, 0, 0, 0);
        return;
    }
    double u1 = vdot(vsub(c_local, v1), vsub(v2_, v1));
    double u2 = vdot(vsub(c_local, v2_), vsub(v1, v2_));
    if (u1 < 0.0) {
        if (dist2(c_local, v1) > sq(radius)) return;
        m->local_normal = vnorm(vsub(c_local, v1));
        m->local_point = v1;
    } else if (u2 < 0.0) {
        if (dist2(c_local, v2_) > sq(radius)) return;
        m->local_normal = vnorm(vsub(c_local, v2_));
        m->local_point = v2_;
    } else {
        v2 face_center = vscale(0.5, vadd(v1, v2_));
        separation = vdot(vsub(c_local, face_center), pnorm(pa, i1));
        if (separation > radius) return;
        m->local_normal = pnorm(pa, i1);
        m->local_point = face_center;
    }
    m->count = 1; m->type = M_FACE_A;
    m->pt[0] = posb; m->id[0] = mkid(0, 0, 0, 0);
}

/* ---- polygon SAT helpers: collision_util.py:31-164 ---- */
static double edge_separation(const shape_t* p1, const xform* x1, int edge1, const shape_t* p2, const xform* x2)
{
    v2 normal1_world = rot_mul(x1, pnorm(p1, edge1));
    v2 normal1 = rot_mul_t(x2, normal1_world);
    int index = 0; double min_dot = 0;
    for (int i = 0; i < p2->count; ++i) {
        double d = vdot(pvert(p2, i), normal1);
        if (i == 0 || d < min_dot) { min_dot = d; index = i; }
    }
    v2 v1 = xf_mul(x1, pvert(p1, edge1));
    v2 v2_ = xf_mul(x2, pvert(p2, index));
    return vdot(vsub(v2_, v1), normal1_world);
}

static double max_separation(int* edge_out, const shape_t* p1, const xform* x1, const shape_t* p2, const xform* x2)
{
    int count1 = p1->count;
    /* Polygon._centroid is always (0,0) in the reference */
    v2 zero = V(0, 0);
    v2 d = vsub(xf_mul(x2, zero), xf_mul(x1, zero));
    v2 d_local1 = rot_mul_t(x1, d);
    int edge = 0; double max_dot = 0;
    for (int i = 0; i < count1; ++i) {
        double dt = vdot(pnorm(p1, i), d_local1);
        if (i == 0 || dt > max_dot) { max_dot = dt; edge = i; }
    }
    double s = edge_separation(p1, x1, edge, p2, x2);
    int prev_edge = (edge - 1 + count1) % count1;
    double s_prev = edge_separation(p1, x1, prev_edge, p2, x2);
    int next_edge = (edge + 1) % count1;
    double s_next = edge_separation(p1, x1, next_edge, p2, x2);
    int best_edge, increment; double best_separation;
    if (s_prev > s && s_prev > s_next) { increment = -1; best_edge = prev_edge; best_separation = s_prev; }
    else if (s_next > s) { increment = 1; best_edge = next_edge; best_separation = s_next; }
    else { *edge_out = edge; return s; }
    for (;;) {
        if (increment == -1) edge = (best_edge - 1 + count1) % count1;
        else edge = (best_edge + 1) % count1;
        s = edge_separation(p1, x1, edge, p2, x2);
        if (s > best_separation) { best_edge = edge; best_separation = s; }
        else break;
    }
    *edge_out = best_edge;
    return best_separation;
}

static void find_incident_edge(clipv c[2], const shape_t* p1, const xform* x1, int edge1, const shape_t* p2, const xform* x2)
{
    v2 normal1 = rot_mul_t(x2, rot_mul(x1, pnorm(p1, edge1)));
    int index = 0; double min_dot = 0;
    for (int i = 0; i < p2->count; ++i) {
        double d = vdot(normal1, pnorm(p2, i));
        if (i == 0 || d < min_dot) { min_dot = d; index = i; }
    }
    int i1 = index, i2 = (i1 + 1) % p2->count;
    c[0].v = xf_mul(x2, pvert(p2, i1)); c[0].ia = edge1; c[0].ib = i1; c[0].ta = MP_FACE; c[0].tb = MP_VERTEX;
    c[1].v = xf_mul(x2, pvert(p2, i2)); c[1].ia = edge1; c[1].ib = i2; c[1].ta = MP_FACE; c[1].tb = MP_VERTEX;
}

static int clip_segment_to_line(clipv out[3], const clipv in[2], v2 normal, double offset, int vertex_index_a)
{
    int n = 0;
    double distance0 = vdot(normal, in[0].v) - offset;
    double distance1 = vdot(normal, in[1].v) - offset;
    if (distance0 <= 0.0) out[n++] = in[0];
    if (distance1 <= 0.0) out[n++] = in[1];
    if (distance0 * distance1 < 0.0) {
        double interp = distance0 / (distance0 - distance1);
        clipv cv;
        cv.v = vadd(in[0].v, vscale(interp, vsub(in[1].v, in[0].v)));
        cv.ia = vertex_index_a; cv.ib = in[0].ib; cv.ta = MP_VERTEX; cv.tb = MP_FACE;
        out[n++] = cv;
    }
    return n;
}

/* ---- collide_polygons: collision.py:144-241 ---- */
static void collide_polygons(manifold_t* m, const shape_t* pa, const xform* xa, const shape_t* pb, const xform* xb)
{
    m->count = 0;
    double total_radius = pa->radius + pb->radius;
    int edge_a, edge_b;
    double separation_a = max_separation(&edge_a, pa, xa, pb, xb);
    if (separation_a > total_radius) return;
    double separation_b = max_separation(&edge_b, pb, xb, pa, xa);
    if (separation_b > total_radius) return;

    const shape_t *poly1, *poly2; const xform *xf1, *xf2; int edge1, flip;
    if (separation_b > (REL_TOL * separation_a + ABS_TOL)) {
        poly1 = pb; poly2 = pa; xf1 = xb; xf2 = xa; edge1 = edge_b; m->type = M_FACE_B; flip = 1;
    } else {
        poly1 = pa; poly2 = pb; xf1 = xa; xf2 = xb; edge1 = edge_a; m->type = M_FACE_A; flip = 0;
    }
    clipv incident[2];
    find_incident_edge(incident, poly1, xf1, edge1, poly2, xf2);
    int iv1 = edge1, iv2 = (edge1 + 1) % poly1->count;
    v2 v11 = pvert(poly1, iv1), v12 = pvert(poly1, iv2);
    v2 localtangent = vnorm(vsub(v12, v11));
    v2 local_normal = vcross_s(localtangent, 1.0);
    v2 planepoint = vscale(0.5, vadd(v11, v12));
    v2 tangent = rot_mul(xf1, localtangent);
    v2 normal = vcross_s(tangent, 1.0);
    v11 = xf_mul(xf1, v11);
    v12 = xf_mul(xf1, v12);
    double front_offset = vdot(normal, v11);
    double side_offset1 = -vdot(tangent, v11) + total_radius;
    double side_offset2 = vdot(tangent, v12) + total_radius;
    clipv cp1[3], cp2[3];
    if (clip_segment_to_line(cp1, incident, vneg(tangent), side_offset1, iv1) < 2) return;
    if (clip_segment_to_line(cp2, cp1, tangent, side_offset2, iv2) < 2) return;
    m->local_normal = local_normal;
    m->local_point = planepoint;
    int point_count = 0;
    for (int i = 0; i < 2; ++i) {
        double separation = vdot(normal, cp2[i].v) - front_offset;
        if (separation <= total_radius) {
            m->pt[point_count] = xf_mul_t(xf2, cp2[i].v);
            if (flip) m->id[point_count] = mkid(cp2[i].ib, cp2[i].ia, cp2[i].tb, cp2[i].ta);
            else m->id[point_count] = mkid(cp2[i].ia, cp2[i].ib, cp2[i].ta, cp2[i].tb);
            ++point_count;
        }
    }
    m->count = point_count;
}

/* ---- collide_edge_circle: collision.py:246-353 ---- */
static void collide_edge_circle(manifold_t* m, const shape_t* ea, const xform* xa, const shape_t* cb, const xform* xb)
{
    m->count = 0;
    v2 posb = V(cb->d[0], cb->d[1]);
    v2 Q = xf_mul_t(xa, xf_mul(xb, posb));
    v2 A = V(ea->d[0], ea->d[1]), B = V(ea->d[2], ea->d[3]);
    v2 e = vsub(B, A);
    double u = vdot(e, vsub(B, Q));
    double v = vdot(e, vsub(Q, A));
    double radius = ea->radius + cb->radius;
    if (v <= 0.0) {
        v2 P = A, d = vsub(Q, P);
        double dd = vdot(d, d);
        if (dd > sq(radius)) return;
        if (ea->has0) {
            v2 A1 = V(ea->d[4], ea->d[5]), B1 = A;
            v2 e1 = vsub(B1, A1);
            double u1 = vdot(e1, vsub(B1, Q));
            if (u1 > 0.0) return;
        }
        m->count = 1; m->type = M_CIRCLES; m->local_normal = V(0, 0); m->local_point = P;
        m->id[0] = mkid(0, 0, MP_VERTEX, MP_VERTEX); m->pt[0] = posb;
        return;
    }
    if (u <= 0.0) {
        v2 P = B, d = vsub(Q, P);
        double dd = vdot(d, d);
        if (dd > sq(radius)) return;
        if (ea->has3) {
            v2 B2 = V(ea->d[6], ea->d[7]), A2 = B;
            v2 e2 = vsub(B2, A2);
            double v2s = vdot(e2, vsub(Q, A2));
            if (v2s > 0.0) return;
        }
        m->count = 1; m->type = M_CIRCLES; m->local_normal = V(0, 0); m->local_point = P;
        m->id[0] = mkid(1, 0, MP_VERTEX, MP_VERTEX); m->pt[0] = posb;
        return;
    }
    double den = vdot(e, e);
    v2 P = vscale(1.0 / den, vadd(vscale(u, A), vscale(v, B)));
    v2 d = vsub(Q, P);
    double dd = vdot(d, d);
    if (dd > sq(radius)) return;
    v2 n = V(-e.y, e.x);
    if (vdot(n, vsub(Q, A)) < 0.0) n = V(-n.x, -n.y);
    n = vnorm(n);
    m->count = 1; m->type = M_FACE_A; m->local_normal = n; m->local_point = A;
    m->id[0] = mkid(0, 0, MP_FACE, MP_VERTEX); m->pt[0] = posb;
}

/* ---- EPCollider: collision.py:355-690 ---- */
enum { AX_UNKNOWN = 0, AX_EDGE_A = 1, AX_EDGE_B = 2 };
typedef struct { int type, index; double separation; } epaxis;
typedef struct {
    xform xf;
    v2 v0, v1, v2_, v3; int has0, has3;
    v2 enormal;
    v2 avert[2], anorm[2];
    v2 bcentroid, bvert[MAXV], bnorm[MAXV]; int bcount;
    double radius;
    v2 l11, l12, l21, l22;
} epc_t;

static int ep_valid(const epc_t* c, v2 n)
{
    int valid1 = (vcross(n, c->l11) >= -ANGULAR_SLOP) && (vcross(c->l12, n) >= -ANGULAR_SLOP);
    int valid2 = (vcross(n, c->l21) >= -ANGULAR_SLOP) && (vcross(c->l22, n) >= -ANGULAR_SLOP);
    return valid1 && valid2;
}

static void ep_incident(clipv out[2], const v2* normals1, int edge1, const v2* verts2, const v2* normals2, int n2)
{
    v2 normal1 = normals1[edge1];
    int index = 0; double min_dot = 0;
    for (int i = 0; i < n2; ++i) {
        double d = vdot(normal1, normals2[i]);
        if (i == 0 || d < min_dot) { min_dot = d; index = i; }
    }
    int i1 = index, i2 = (i1 + 1) % n2;
    out[0].v = verts2[i1]; out[0].ia = edge1; out[0].ib = i1; out[0].ta = MP_FACE; out[0].tb = MP_VERTEX;
    out[1].v = verts2[i2]; out[1].ia = edge1; out[1].ib = i2; out[1].ta = MP_FACE; out[1].tb = MP_VERTEX;
}

static void collide_edge_polygon(manifold_t* m, const shape_t* ea, const xform* xa, const shape_t* pb, const xform* xb)
{
    epc_t c;
    /* Transform.mul_t(Transform): offset is NOT rotated (common.py:671-673) */
    c.xf.p = vsub(xb->p, xa->p);
    c.xf.c1 = V(vdot(xa->c1, xb->c1), vdot(xa->c2, xb->c1));
    c.xf.c2 = V(vdot(xa->c1, xb->c2), vdot(xa->c2, xb->c2));
    c.v1 = V(ea->d[0], ea->d[1]); c.v2_ = V(ea->d[2], ea->d[3]);
    c.v0 = V(ea->d[4], ea->d[5]); c.v3 = V(ea->d[6], ea->d[7]);
    c.has0 = ea->has0; c.has3 = ea->has3;
    v2 e = vsub(c.v2_, c.v1);
    c.enormal = vnorm(V(e.y, -e.x));
    c.avert[0] = c.v1; c.avert[1] = c.v2_;
    c.anorm[0] = c.enormal; c.anorm[1] = vneg(c.enormal);
    c.bcentroid = xf_mul(&c.xf, V(0, 0));
    c.bcount = pb->count;
    for (int i = 0; i < pb->count; ++i) {
        c.bvert[i] = xf_mul(&c.xf, pvert(pb, i));
        c.bnorm[i] = rot_mul(&c.xf, pnorm(pb, i));
    }
    c.radius = 2.0 * POLY_RADIUS;
    c.l11 = c.l12 = c.l21 = c.l22 = V(0, 0);

    m->count = 0;
    /* compute_adjacency: collision.py:537-597 */
    {
        v2 center_b = c.bcentroid;
        if (c.has0) {
            v2 e0 = vsub(c.v1, c.v0), e1 = vsub(c.v2_, c.v1);
            v2 n0 = vnorm(V(e0.y, -e0.x)), n1 = vnorm(V(e1.y, -e1.x));
            int convex = vcross(n0, n1) >= 0.0;
            int front0 = vdot(n0, vsub(center_b, c.v0)) >= 0.0;
            int front1 = vdot(n1, vsub(center_b, c.v1)) >= 0.0;
            if (convex) {
                if (front0 || front1) { c.l11 = n1; c.l12 = n0; } else { c.l11 = vneg(n1); c.l12 = vneg(n0); }
            } else {
                if (front0 && front1) { c.l11 = n0; c.l12 = n1; } else { c.l11 = vneg(n0); c.l12 = vneg(n1); }
            }
        } else { c.l11 = V(0, 0); c.l12 = V(0, 0); }
        if (c.has3) {
            v2 e1 = vsub(c.v2_, c.v1), e2 = vsub(c.v3, c.v2_);
            v2 n1 = vnorm(V(e1.y, -e1.x)), n2 = vnorm(V(e2.y, -e2.x));
            int convex = vcross(n1, n2) >= 0.0;
            int front1 = vdot(n1, vsub(center_b, c.v1)) >= 0.0;
            int front2 = vdot(n2, vsub(center_b, c.v2_)) >= 0.0;
            if (convex) {
                if (front1 || front2) { c.l21 = n2; c.l22 = n1; } else { c.l21 = vneg(n2); c.l22 = vneg(n1); }
            } else {
                if (front1 && front2) { c.l21 = n1; c.l22 = n2; } else { c.l21 = vneg(n1); c.l22 = vneg(n2); }
            }
        } else { c.l21 = V(0, 0); c.l22 = V(0, 0); }
    }
    /* compute_edge_separation: collision.py:599-629 */
    epaxis edge_axis; edge_axis.type = AX_UNKNOWN; edge_axis.index = -1; edge_axis.separation = -MAX_FLOAT;
    for (int i = 0; i < 2; ++i) {
        v2 n = c.anorm[i];
        if (!ep_valid(&c, n)) continue;
        epaxis axis; axis.type = AX_EDGE_A; axis.index = i;
        double mn = 0;
        for (int k = 0; k < c.bcount; ++k) {
            double d = vdot(n, vsub(c.bvert[k], c.v1));
            if (k == 0 || d < mn) mn = d;
        }
        axis.separation = mn;
        if (axis.separation > c.radius) { edge_axis = axis; goto edge_done; }
        if (axis.separation > edge_axis.separation) edge_axis = axis;
    }
edge_done:
    if (edge_axis.type == AX_UNKNOWN) return;
    if (edge_axis.separation > c.radius) return;
    /* compute_polygon_separation: collision.py:631-664 (no early return) */
    epaxis polygon_axis; polygon_axis.type = AX_UNKNOWN; polygon_axis.index = -1; polygon_axis.separation = -MAX_FLOAT;
    for (int i = 0; i < c.bcount; ++i) {
        v2 n = vneg(c.bnorm[i]);
        if (!ep_valid(&c, n)) continue;
        double s1 = vdot(n, vsub(c.bvert[i], c.v1));
        double s2 = vdot(n, vsub(c.bvert[i], c.v2_));
        double s = PYMIN(s1, s2);
        if (s > c.radius) { polygon_axis.type = AX_EDGE_B; polygon_axis.index = i; polygon_axis.separation = s; }
        if (s > polygon_axis.separation) { polygon_axis.type = AX_EDGE_B; polygon_axis.index = i; polygon_axis.separation = s; }
    }
    if (polygon_axis.type != AX_UNKNOWN && polygon_axis.separation > c.radius) return;
    epaxis primary;
    if (polygon_axis.type == AX_UNKNOWN) primary = edge_axis;
    else if (polygon_axis.separation > REL_TOL * edge_axis.separation + ABS_TOL) primary = polygon_axis;
    else primary = edge_axis;

    const v2 *verts1, *norms1, *verts2, *norms2; int n1, n2;
    if (primary.type == AX_EDGE_A) {
        verts1 = c.avert; norms1 = c.anorm; n1 = 2; verts2 = c.bvert; norms2 = c.bnorm; n2 = c.bcount;
        m->type = M_FACE_A;
    } else {
        verts1 = c.bvert; norms1 = c.bnorm; n1 = c.bcount; verts2 = c.avert; norms2 = c.anorm; n2 = 2;
        m->type = M_FACE_B;
    }
    int edge1 = primary.index;
    clipv incident[2];
    ep_incident(incident, norms1, edge1, verts2, norms2, n2);
    int iv1 = edge1, iv2 = (edge1 + 1) % n1;
    v2 v11 = verts1[iv1], v12 = verts1[iv2];
    v2 tangent = vnorm(vsub(v12, v11));
    v2 normal = vcross_s(tangent, 1.0);
    v2 plane_point = vscale(0.5, vadd(v11, v12));
    double front_offset = vdot(normal, v11);
    double side_offset1 = -vdot(tangent, v11) + c.radius;
    double side_offset2 = vdot(tangent, v12) + c.radius;
    clipv cp1[3], cp2[3];
    if (clip_segment_to_line(cp1, incident, vneg(tangent), side_offset1, iv1) < 2) return;
    int ncp2 = clip_segment_to_line(cp2, cp1, tangent, side_offset2, iv2);
    if (ncp2 < 2) return;
    if (primary.type == AX_EDGE_A) {
        m->local_normal = normal;
        m->local_point = plane_point;
    } else {
        m->local_normal = rot_mul_t(&c.xf, normal);
        m->local_point = xf_mul_t(&c.xf, plane_point);
    }
    int point_count = 0;
    for (int i = 0; i < ncp2; ++i) {
        double separation = vdot(normal, cp2[i].v) - front_offset;
        if (separation <= c.radius) {
            if (primary.type == AX_EDGE_A) {
                m->pt[point_count] = xf_mul_t(&c.xf, cp2[i].v);
                m->id[point_count] = mkid(cp2[i].ia, cp2[i].ib, cp2[i].ta, cp2[i].tb);
            } else {
                m->pt[point_count] = cp2[i].v;
                /* id stored as (type_b, type_a, index_b, index_a): SURVEY A.9 #6, collision.py:532 */
                m->id[point_count] = mkid(cp2[i].tb, cp2[i].ta, cp2[i].ib, cp2[i].ia);
            }
            ++point_count;
        }
    }
    m->count = point_count;
}

/* dispatch: contact.py:294-311,343-361 (shape A/B already in registry order) */
static void evaluate(manifold_t* m, const shape_t* a, const xform* xa, const shape_t* b, const xform* xb)
{
    if (a->type == S_CIRCLE && b->type == S_CIRCLE) collide_circles(m, a, xa, b, xb);
    else if (a->type == S_POLYGON && b->type == S_POLYGON) collide_polygons(m, a, xa, b, xb);
    else if (a->type == S_POLYGON && b->type == S_CIRCLE) collide_polygon_circle(m, a, xa, b, xb);
    else if (a->type == S_EDGE && b->type == S_CIRCLE) collide_edge_circle(m, a, xa, b, xb);
    else if (a->type == S_EDGE && b->type == S_POLYGON) collide_edge_polygon(m, a, xa, b, xb);
    else m->count = 0;
}

/* ---- manifold <-> contact rows ---- */
static void manifold_load(const orc_world* w, int ci, manifold_t* m)
{
    int mt = w->c_mtype[ci];
    m->type = mt & 0xff; m->count = (mt >> 8) & 0xff;
    const double* mf = w->c_mf + 4 * (size_t)ci; const double* pts = w->c_pts + 4 * (size_t)ci;
    m->local_normal = V(mf[0], mf[1]); m->local_point = V(mf[2], mf[3]);
    m->pt[0] = V(pts[0], pts[1]); m->pt[1] = V(pts[2], pts[3]);
    m->id[0] = w->c_id[2 * ci]; m->id[1] = w->c_id[2 * ci + 1];
}
static void manifold_store(orc_world* w, int ci, const manifold_t* m)
{
    w->c_mtype[ci] = m->type | (m->count << 8);
    double* mf = w->c_mf + 4 * (size_t)ci; double* pts = w->c_pts + 4 * (size_t)ci;
    mf[0] = m->local_normal.x; mf[1] = m->local_normal.y; mf[2] = m->local_point.x; mf[3] = m->local_point.y;
    pts[0] = m->pt[0].x; pts[1] = m->pt[0].y; pts[2] = m->pt[1].x; pts[3] = m->pt[1].y;
    w->c_id[2 * ci] = m->id[0]; w->c_id[2 * ci + 1] = m->id[1];
}

static inline int body_type(const orc_world* w, int b) { return (int)(w->b_flags[b] & BODY_TYPE_MASK); }
static inline int body_awake(const orc_world* w, int b) { return (w->b_flags[b] & BODY_AWAKE) != 0; }

/* Body.awake = True (body.py:800-804,812) */
static void body_wake(orc_world* w, int b)
{
    if (!(w->b_flags[b] & BODY_AWAKE)) w->b_pos[4 * b + 3] = 0.0;
    w->b_flags[b] |= BODY_AWAKE;
}
/* Body.awake = False (body.py:805-812) */
static void body_sleep(orc_world* w, int b)
{
    w->b_pos[4 * b + 3] = 0.0;
    w->b_vel[4 * b + 0] = 0.0; w->b_vel[4 * b + 1] = 0.0; w->b_vel[4 * b + 2] = 0.0; w->b_vel[4 * b + 3] = 0.0;
    w->b_force[2 * b + 0] = 0.0; w->b_force[2 * b + 1] = 0.0;
    w->b_flags[b] &= ~BODY_AWAKE;
}

static int sensor_test_overlap(const orc_world* w, int fa, int fb, const xform* xa, const xform* xb);

/* ---- Contact.update: contact.py:210-268 (no listeners) ---- */
static void contact_update(orc_world* w, int ci)
{
    if (w->c_flags[ci] & C_SENSOR) {
        /* contact.py:232-239: a sensor contact has no manifold points; touching = the shapes overlap (GJK with radii);
         * nobody is woken by a change (contact.py:257-259 sits in the other branch) */
        int fa = w->c_fix[2 * ci], fb = w->c_fix[2 * ci + 1];
        xform xa = xf_row(w->b_xf + 4 * (size_t)w->f_body[fa]), xb = xf_row(w->b_xf + 4 * (size_t)w->f_body[fb]);
        int touching = sensor_test_overlap(w, fa, fb, &xa, &xb);
        w->c_flags[ci] |= C_ENABLED;
        w->c_mtype[ci] = w->c_mtype[ci] & 0xff;                                  /* point_count = 0 */
        if (touching) w->c_flags[ci] |= C_TOUCHING; else w->c_flags[ci] &= ~C_TOUCHING;
        return;
    }
    manifold_t m; manifold_load(w, ci, &m);
    manifold_t old = m;
    double* imp = w->c_imp + 4 * (size_t)ci;
    double old_ni[2] = {imp[0], imp[2]}, old_ti[2] = {imp[1], imp[3]};
    w->c_flags[ci] |= C_ENABLED;
    int was_touching = (w->c_flags[ci] & C_TOUCHING) != 0;
    int fa = w->c_fix[2 * ci], fb = w->c_fix[2 * ci + 1];
    int ba = w->f_body[fa], bb = w->f_body[fb];
    xform xa = xf_row(w->b_xf + 4 * (size_t)ba), xb = xf_row(w->b_xf + 4 * (size_t)bb);
    shape_t sa = shape_get(w, w->f_shape[fa]), sb = shape_get(w, w->f_shape[fb]);
    evaluate(&m, &sa, &xa, &sb, &xb);
    int touching = m.count > 0;
    /* match ids: dict of old used points, last duplicate wins (contact.py:247-255) */
    for (int i = 0; i < m.count; ++i) {
        double ni = 0.0, ti = 0.0;
        for (int j = 0; j < old.count; ++j)
            if (old.id[j] == m.id[i]) { ni = old_ni[j]; ti = old_ti[j]; }
        imp[2 * i] = ni; imp[2 * i + 1] = ti;
    }
    manifold_store(w, ci, &m);
    if (touching != was_touching) { body_wake(w, ba); body_wake(w, bb); }
    if (touching) w->c_flags[ci] |= C_TOUCHING; else w->c_flags[ci] &= ~C_TOUCHING;
}

static inline int aabb_overlap(const double* a, const double* b)               /* common.py:830-839 */
{
    double d1x = b[0] - a[2], d1y = b[1] - a[3];
    double d2x = a[0] - b[2], d2y = a[1] - b[3];
    if (d1x > 0.0 || d1y > 0.0) return 0;
    if (d2x > 0.0 || d2y > 0.0) return 0;
    return 1;
}

static void contact_copy(orc_world* w, int dst, int src)
{
    w->c_fix[2 * dst] = w->c_fix[2 * src]; w->c_fix[2 * dst + 1] = w->c_fix[2 * src + 1];
    w->c_flags[dst] = w->c_flags[src];
    w->c_mat[2 * dst] = w->c_mat[2 * src]; w->c_mat[2 * dst + 1] = w->c_mat[2 * src + 1];
    w->c_mtype[dst] = w->c_mtype[src];
    memcpy(w->c_mf + 4 * (size_t)dst, w->c_mf + 4 * (size_t)src, 4 * sizeof(double));
    memcpy(w->c_pts + 4 * (size_t)dst, w->c_pts + 4 * (size_t)src, 4 * sizeof(double));
    memcpy(w->c_imp + 4 * (size_t)dst, w->c_imp + 4 * (size_t)src, 4 * sizeof(double));
    w->c_id[2 * dst] = w->c_id[2 * src]; w->c_id[2 * dst + 1] = w->c_id[2 * src + 1];
}

/* ---- ContactManager.collide: contact_manager.py:174-223 ---- */
static void cm_collide(orc_world* w)
{
    int nc = w->w_count[2];
    uint8_t* remove = (uint8_t*)calloc((size_t)nc + 1, 1);
    for (int ci = 0; ci < nc; ++ci) {
        int fa = w->c_fix[2 * ci], fb = w->c_fix[2 * ci + 1];
        int ba = w->f_body[fa], bb = w->f_body[fb];
        int active_a = body_awake(w, ba) && body_type(w, ba) != T_STATIC;
        int active_b = body_awake(w, bb) && body_type(w, bb) != T_STATIC;
        if (!active_a && !active_b) continue;
        /* filter flag is never raised on this path (no joints, no refilter) */
        if (!aabb_overlap(w->f_fat + 4 * (size_t)fa, w->f_fat + 4 * (size_t)fb)) { remove[ci] = 1; continue; }
        contact_update(w, ci);
    }
    /* destroy (contact_manager.py:160-172): order-preserving removal; bodies are NOT woken (A.9 #12) */
    int dst = 0;
    for (int ci = 0; ci < nc; ++ci) {
        if (remove[ci]) continue;
        if (dst != ci) contact_copy(w, dst, ci);
        ++dst;
    }
    w->w_count[2] = dst;
    free(remove);
}

/* ---- BroadPhase.update_pairs (order pinned) + ContactManager.find_new_contacts:
 *      broadphase.py:83-109, dynamictree.py:190-207, contact_manager.py:102-158 ---- */
static void cm_find_new_contacts(orc_world* w)
{
    int nf = w->w_count[1];
    int npairs = 0;
    for (int a = 0; a < nf; ++a) {
        if (!w->f_moved[a]) continue;
        const double* fa_box = w->f_fat + 4 * (size_t)a;
        for (int b = 0; b < nf; ++b) {
            if (b == a) continue;
            if (!aabb_overlap(w->f_fat + 4 * (size_t)b, fa_box)) continue;     /* node.aabb.test_overlap(query) */
            if (b < a && w->f_moved[b]) continue;                              /* (b, a) already in the pair buffer */
            if (npairs < w->npair_cap) { w->p_pairs[2 * npairs] = a; w->p_pairs[2 * npairs + 1] = b; }
            else w->w_status[0] |= ST_PAIR_OVERFLOW;
            ++npairs;

            int fixture_a = a, fixture_b = b;
            int body_a = w->f_body[fixture_a], body_b = w->f_body[fixture_b];
            if (body_a == body_b) continue;
            int exists = 0;
            int nc = w->w_count[2];
            for (int ci = 0; ci < nc; ++ci) {
                int ca = w->c_fix[2 * ci], cb = w->c_fix[2 * ci + 1];
                if ((ca == fixture_a && cb == fixture_b) || (ca == fixture_b && cb == fixture_a)) { exists = 1; break; }
            }
            if (exists) continue;
            /* Body._should_collide: body.py:910-928 */
            if (body_type(w, body_b) != T_DYNAMIC && body_type(w, body_a) != T_DYNAMIC) continue;
            {
                int blocked = 0;
                for (int k = w->b_jadj[2 * body_b], ke = k + w->b_jadj[2 * body_b + 1]; k < ke; ++k) {
                    const int32_t* def = w->j_def + 4 * (size_t)w->j_adj[k];
                    if (((def[1] == body_b && def[2] == body_a) || (def[2] == body_b && def[1] == body_a)) && !(def[3] & 0x1)) { blocked = 1; break; }
                }
                if (blocked) continue;
            }
            /* World.default_contact_filter: world.py:958-970 */
            {
                uint32_t fa0 = w->f_filter[2 * fixture_a], fa1 = w->f_filter[2 * fixture_a + 1];
                uint32_t fb0 = w->f_filter[2 * fixture_b], fb1 = w->f_filter[2 * fixture_b + 1];
                int ga = (int16_t)(fa1 & 0xffff), gb = (int16_t)(fb1 & 0xffff);
                uint32_t cat_a = fa0 & 0xffff, mask_a = fa0 >> 16, cat_b = fb0 & 0xffff, mask_b = fb0 >> 16;
                int ok;
                if (ga == gb && ga != 0) ok = ga > 0;
                else ok = ((mask_a & cat_b) != 0) && ((cat_a & mask_b) != 0);
                if (!ok) continue;
            }
            /* Contact._create: contact.py:294-326 */
            int t1 = w->s_meta[w->f_shape[fixture_a]] & 0xff, t2 = w->s_meta[w->f_shape[fixture_b]] & 0xff;
            int primary;
            if ((t1 == S_CIRCLE && t2 == S_CIRCLE) || (t1 == S_POLYGON && t2 == S_POLYGON) ||
                (t1 == S_POLYGON && t2 == S_CIRCLE) || (t1 == S_EDGE && t2 == S_CIRCLE) ||
                (t1 == S_EDGE && t2 == S_POLYGON)) primary = 1;
            else if ((t1 == S_CIRCLE && t2 == S_POLYGON) || (t1 == S_CIRCLE && t2 == S_EDGE) ||
                     (t1 == S_POLYGON && t2 == S_EDGE)) primary = 0;
            else continue;                                                     /* unregistered pair (edge-edge) */
            if (!primary) { int t = fixture_a; fixture_a = fixture_b; fixture_b = t; }
            if (nc >= w->nc_cap) { w->w_status[0] |= ST_CONTACT_OVERFLOW; continue; }
            int ci = nc;
            w->c_fix[2 * ci] = fixture_a; w->c_fix[2 * ci + 1] = fixture_b;
            w->c_flags[ci] = (((w->f_filter[2 * fixture_a + 1] | w->f_filter[2 * fixture_b + 1]) >> 16) & 1u) ? C_SENSOR : 0;
            /* mix_friction / mix_restitution: contact_util.py:36-48 */
            w->c_mat[2 * ci] = sqrt(w->f_mat[2 * fixture_a] * w->f_mat[2 * fixture_b]);
            {
                double r1 = w->f_mat[2 * fixture_a + 1], r2 = w->f_mat[2 * fixture_b + 1];
                w->c_mat[2 * ci + 1] = PYMAX(r1, r2);
            }
            w->c_mtype[ci] = 0;
            memset(w->c_mf + 4 * (size_t)ci, 0, 4 * sizeof(double));
            memset(w->c_pts + 4 * (size_t)ci, 0, 4 * sizeof(double));
            memset(w->c_imp + 4 * (size_t)ci, 0, 4 * sizeof(double));
            w->c_id[2 * ci] = 0; w->c_id[2 * ci + 1] = 0;
            w->w_count[2] = nc + 1;
        }
    }
    for (int a = 0; a < nf; ++a) w->f_moved[a] = 0;
    w->w_count[3] = npairs;
}

/* ---- Shape.compute_aabb: shapes.py:178-189 (circle), 316-334 (polygon), 702-718 (edge) ---- */
static void shape_aabb(const shape_t* s, const xform* t, double out[4])
{
    if (s->type == S_CIRCLE) {
        v2 pos = V(s->d[0], s->d[1]);
        v2 p = vadd(t->p, rot_mul(t, pos));
        out[0] = p.x - s->radius; out[1] = p.y - s->radius; out[2] = p.x + s->radius; out[3] = p.y + s->radius;
    } else if (s->type == S_POLYGON) {
        v2 lower = xf_mul(t, pvert(s, 0)), upper = lower;
        for (int i = 0; i < s->count; ++i) {
            v2 v = xf_mul(t, pvert(s, i));
            lower = V(PYMIN(lower.x, v.x), PYMIN(lower.y, v.y));
            upper = V(PYMAX(upper.x, v.x), PYMAX(upper.y, v.y));
        }
        out[0] = lower.x - s->radius; out[1] = lower.y - s->radius; out[2] = upper.x + s->radius; out[3] = upper.y + s->radius;
    } else {
        v2 a = xf_mul(t, V(s->d[0], s->d[1])), b = xf_mul(t, V(s->d[2], s->d[3]));
        v2 lower = V(PYMIN(a.x, b.x), PYMIN(a.y, b.y)), upper = V(PYMAX(a.x, b.x), PYMAX(a.y, b.y));
        if (s->loop_child) { out[0] = lower.x; out[1] = lower.y; out[2] = upper.x; out[3] = upper.y; return; }   /* shapes.py:822-837 */
        out[0] = lower.x - s->radius; out[1] = lower.y - s->radius; out[2] = upper.x + s->radius; out[3] = upper.y + s->radius;
    }
}

/* ---- Fixture._synchronize + BroadPhase.move_proxy + DynamicTree.move_proxy:
 *      fixture.py:287-298, broadphase.py:61-68, dynamictree.py:152-188 ---- */
static void fixture_synchronize(orc_world* w, int f, const xform* xf1, const xform* xf2)
{
    shape_t s = shape_get(w, w->f_shape[f]);
    double a1[4], a2[4];
    shape_aabb(&s, xf1, a1);
    shape_aabb(&s, xf2, a2);
    double* tight = w->f_tight + 4 * (size_t)f;
    tight[0] = PYMIN(a1[0], a2[0]); tight[1] = PYMIN(a1[1], a2[1]);
    tight[2] = PYMAX(a1[2], a2[2]); tight[3] = PYMAX(a1[3], a2[3]);
    v2 displacement = vsub(xf2->p, xf1->p);
    double* fat = w->f_fat + 4 * (size_t)f;
    /* AABB.contains: common.py:863-868 */
    if (fat[0] <= tight[0] && fat[1] <= tight[1] && tight[2] <= fat[2] && tight[3] <= fat[3]) return;
    double lx = tight[0] - AABB_EXT, ly = tight[1] - AABB_EXT, ux = tight[2] + AABB_EXT, uy = tight[3] + AABB_EXT;
    v2 d = vscale(AABB_MULT, displacement);
    if (d.x < 0.0) { lx = lx + d.x; ly = ly + 0; } else { ux = ux + d.x; uy = uy + 0; }
    if (d.y < 0.0) { lx = lx + 0; ly = ly + d.y; } else { ux = ux + 0; uy = uy + d.y; }
    fat[0] = lx; fat[1] = ly; fat[2] = ux; fat[3] = uy;
    w->f_moved[f] = 1;
}

/* ---- ContactSolver + Island.solve: contact.py:375-864, contact_util.py:195-360, island.py:65-179 ---- */
typedef struct {
    v2 ra, rb;
    double normal_impulse, tangent_impulse, normal_mass, tangent_mass, velocity_bias;
} vcp_t;
typedef struct {
    int contact;
    int ia, ib;                 /* island indices */
    double ima, imb, iia, iib;  /* inv mass / inv inertia */
    double friction, restitution;
    v2 normal;
    int npoints;                /* len(vc.points): may shrink to 1 */
    vcp_t p[2];
    double k11, k12, k22;       /* vc.K */
    double nm_c1x, nm_c1y, nm_c2x, nm_c2y;   /* vc.normal_mass = K.inverse (columns) */
    /* position constraint */
    v2 lca, lcb;                /* local centers */
    double ra_, rb_;            /* radii */
    v2 local_normal, local_point;
    int type, n_local_points;
    v2 local_points[2];
} cc_t;

typedef struct { v2 c; double a; } pos_t;
typedef struct { v2 v; double w; } vel_t;

/* WorldManifold: contact_util.py:195-245 */
static int world_manifold(const cc_t* cc, int mcount, const xform* xa, const xform* xb, v2* normal_out, v2 pts[2])
{
    v2 normal = V(0, 0);
    int n = 0;
    if (mcount == 0) { *normal_out = normal; return 0; }
    if (cc->type == M_CIRCLES) {
        normal = V(1.0, 0.0);
        v2 point_a = xf_mul(xa, cc->local_point);
        v2 point_b = xf_mul(xb, cc->local_points[0]);
        if (dist2(point_a, point_b) > EPS_SQR) normal = vnorm(vsub(point_b, point_a));
        v2 c_a = vadd(point_a, vscale(cc->ra_, normal));
        v2 c_b = vsub(point_b, vscale(cc->rb_, normal));
        pts[0] = vscale(0.5, vadd(c_a, c_b));
        n = 1;
    } else if (cc->type == M_FACE_A) {
        normal = rot_mul(xa, cc->local_normal);
        v2 plane_point = xf_mul(xa, cc->local_point);
        for (int i = 0; i < mcount; ++i) {
            v2 clip_point = xf_mul(xb, cc->local_points[i]);
            v2 c_a = vadd(clip_point, vscale(cc->ra_ - vdot(vsub(clip_point, plane_point), normal), normal));
            v2 c_b = vsub(clip_point, vscale(cc->rb_, normal));
            pts[i] = vscale(0.5, vadd(c_a, c_b));
        }
        n = mcount;
    } else {
        normal = rot_mul(xb, cc->local_normal);
        v2 plane_point = xf_mul(xb, cc->local_point);
        for (int i = 0; i < mcount; ++i) {
            v2 clip_point = xf_mul(xa, cc->local_points[i]);
            v2 c_b = vadd(clip_point, vscale(cc->rb_ - vdot(vsub(clip_point, plane_point), normal), normal));
            v2 c_a = vsub(clip_point, vscale(cc->ra_, normal));
            pts[i] = vscale(0.5, vadd(c_a, c_b));
        }
        n = mcount;
        normal = vneg(normal);
    }
    *normal_out = normal;
    return n;
}

/* ContactPositionConstraint.solver_manifold: contact_util.py:323-360 */
static void solver_manifold(const cc_t* cc, const xform* xa, const xform* xb, int index, v2* normal, v2* point, double* separation)
{
    if (cc->type == M_CIRCLES) {
        v2 point_a = xf_mul(xa, cc->local_point);
        v2 point_b = xf_mul(xb, cc->local_points[0]);
        *normal = vnorm(vsub(point_b, point_a));
        *point = vscale(0.5, vadd(point_a, point_b));
        *separation = vdot(vsub(point_b, point_a), *normal) - cc->ra_ - cc->rb_;
    } else if (cc->type == M_FACE_A) {
        *normal = rot_mul(xa, cc->local_normal);
        v2 plane_point = xf_mul(xa, cc->local_point);
        v2 clip_point = xf_mul(xb, cc->local_points[index]);
        *separation = vdot(vsub(clip_point, plane_point), *normal) - cc->ra_ - cc->rb_;
        *point = clip_point;
    } else {
        *normal = rot_mul(xb, cc->local_normal);
        v2 plane_point = xf_mul(xb, cc->local_point);
        v2 clip_point = xf_mul(xa, cc->local_points[index]);
        *separation = vdot(vsub(clip_point, plane_point), *normal) - cc->ra_ - cc->rb_;
        *point = clip_point;
        *normal = vneg(*normal);
    }
}

static inline v2 rel_vel(v2 vb, double wb, v2 rb, v2 va, double wa, v2 ra)
{
    /* vb + scalar_cross(wb, rb) - va - scalar_cross(wa, ra) */
    return vsub(vsub(vadd(vb, scross(wb, rb)), va), scross(wa, ra));
}

/* ContactSolver.__init__: contact.py:376-412 (warm: step.warm_starting, contact_util.py:260-265) */
static void solver_ctor(const orc_world* w, cc_t* C, const int* contacts, int ncontacts, int warm, double dt_ratio)
{
    for (int k = 0; k < ncontacts; ++k) {
        int ci = contacts[k];
        cc_t* cc = &C[k];
        memset(cc, 0, sizeof(*cc));
        cc->contact = ci;
        int fa = w->c_fix[2 * ci], fb = w->c_fix[2 * ci + 1];
        int ba = w->f_body[fa], bb = w->f_body[fb];
        cc->ia = w->b_island[ba]; cc->ib = w->b_island[bb];
        cc->ima = w->b_mass[4 * ba]; cc->imb = w->b_mass[4 * bb];
        cc->iia = w->b_mass[4 * ba + 1]; cc->iib = w->b_mass[4 * bb + 1];
        cc->lca = V(w->b_mass[4 * ba + 2], w->b_mass[4 * ba + 3]);
        cc->lcb = V(w->b_mass[4 * bb + 2], w->b_mass[4 * bb + 3]);
        cc->friction = w->c_mat[2 * ci]; cc->restitution = w->c_mat[2 * ci + 1];
        cc->ra_ = w->s_radius[w->f_shape[fa]]; cc->rb_ = w->s_radius[w->f_shape[fb]];
        manifold_t m; manifold_load(w, ci, &m);
        cc->type = m.type; cc->local_normal = m.local_normal; cc->local_point = m.local_point;
        cc->n_local_points = m.count; cc->npoints = m.count;
        const double* imp = w->c_imp + 4 * (size_t)ci;
        for (int j = 0; j < m.count; ++j) {
            cc->local_points[j] = m.pt[j];
            /* VelocityConstraintPoint: contact_util.py:260-265 */
            if (warm) { cc->p[j].normal_impulse = dt_ratio * imp[2 * j]; cc->p[j].tangent_impulse = dt_ratio * imp[2 * j + 1]; }
            else { cc->p[j].normal_impulse = imp[2 * j]; cc->p[j].tangent_impulse = imp[2 * j + 1]; }
        }
    }
}

/* initialize_velocity_constraints: contact.py:414-496 */
static void solver_init_velocity(cc_t* C, int ncontacts, const pos_t* P, const vel_t* Vv)
{
    for (int k = 0; k < ncontacts; ++k) {
        cc_t* cc = &C[k];
        double ma = cc->ima, mb = cc->imb, ia = cc->iia, ib = cc->iib;
        v2 ca = P[cc->ia].c, cb = P[cc->ib].c; double aa = P[cc->ia].a, ab = P[cc->ib].a;
        v2 va = Vv[cc->ia].v, vb = Vv[cc->ib].v; double wa = Vv[cc->ia].w, wb = Vv[cc->ib].w;
        xform xa = xf_from_sweep(ca, aa, cc->lca), xb = xf_from_sweep(cb, ab, cc->lcb);
        v2 wpts[2];
        int nw = world_manifold(cc, cc->n_local_points, &xa, &xb, &cc->normal, wpts);
        int np = cc->npoints < nw ? cc->npoints : nw;
        for (int j = 0; j < np; ++j) {
            vcp_t* p = &cc->p[j];
            v2 ra = vsub(wpts[j], ca), rb = vsub(wpts[j], cb);
            p->ra = ra; p->rb = rb;
            double rn_a = sq(vcross(ra, cc->normal)), rn_b = sq(vcross(rb, cc->normal));
            double k_normal = ma + mb + ia * rn_a + ib * rn_b;
            p->normal_mass = (k_normal > 0.0) ? 1.0 / k_normal : 0.0;
            v2 tangent = vcross_s(cc->normal, 1.0);
            double rt_a = sq(vcross(ra, tangent)), rt_b = sq(vcross(rb, tangent));
            double k_tangent = ma + mb + ia * rt_a + ib * rt_b;
            p->tangent_mass = (k_tangent > 0.0) ? 1.0 / k_tangent : 0.0;
            p->velocity_bias = 0.0;
            double v_rel = vdot(cc->normal, rel_vel(vb, wb, rb, va, wa, ra));
            if (v_rel < -VEL_THRESHOLD) p->velocity_bias = -cc->restitution * v_rel;
        }
        if (cc->npoints == 2) {
            v2 ra1 = cc->p[0].ra, rb1 = cc->p[0].rb, ra2 = cc->p[1].ra, rb2 = cc->p[1].rb;
            double rn1a = vcross(ra1, cc->normal), rn1b = vcross(rb1, cc->normal);
            double rn2a = vcross(ra2, cc->normal), rn2b = vcross(rb2, cc->normal);
            double k11 = ma + mb + ia * rn1a * rn1a + ib * rn1b * rn1b;
            double k22 = ma + mb + ia * rn2a * rn2a + ib * rn2b * rn2b;
            double k12 = ma + mb + ia * rn1a * rn2a + ib * rn1b * rn2b;
            if (sq(k11) < (1000.0 * (k11 * k22 - sq(k12)))) {
                cc->k11 = k11; cc->k12 = k12; cc->k22 = k22;
                /* Mat22.inverse: common.py:505-519 */
                double a = k11, b = k12, c = k12, d = k22;
                double det = a * d - b * c;
                if (det != 0.0) det = 1.0 / det;
                cc->nm_c1x = det * d; cc->nm_c1y = -det * c; cc->nm_c2x = -det * b; cc->nm_c2y = det * a;
            } else {
                cc->npoints = 1;
            }
        }
    }
}

/* solve_velocity_constraints: contact.py:527-799 (one sweep) */
static void solver_velocity_sweep(cc_t* C, int ncontacts, vel_t* Vv)
{
    for (int k = 0; k < ncontacts; ++k) {
        cc_t* cc = &C[k];
        double ima = cc->ima, imb = cc->imb, iia = cc->iia, iib = cc->iib;
        v2 normal = cc->normal, tangent = vcross_s(normal, 1.0);
        v2 va = Vv[cc->ia].v, vb = Vv[cc->ib].v; double wa = Vv[cc->ia].w, wb = Vv[cc->ib].w;
        double friction = cc->friction;
        for (int j = 0; j < cc->npoints; ++j) {
            vcp_t* p = &cc->p[j];
            v2 dv = rel_vel(vb, wb, p->rb, va, wa, p->ra);
            double vt = vdot(dv, tangent);
            double lambda = p->tangent_mass * (-vt);
            double max_friction = friction * p->normal_impulse;
            double new_impulse = pyclamp(p->tangent_impulse + lambda, -max_friction, max_friction);
            lambda = new_impulse - p->tangent_impulse;
            v2 Pv = vscale(lambda, tangent);
            va = vsub(va, vscale(ima, Pv));
            wa = wa - iia * vcross(p->ra, Pv);
            vb = vadd(vb, vscale(imb, Pv));
            wb = wb + iib * vcross(p->rb, Pv);
            p->tangent_impulse = new_impulse;
        }
        if (cc->npoints == 1) {
            vcp_t* p = &cc->p[0];
            v2 dv = rel_vel(vb, wb, p->rb, va, wa, p->ra);
            double vn = vdot(dv, normal);
            double lambda = -p->normal_mass * (vn - p->velocity_bias);
            double t = p->normal_impulse + lambda;
            double new_impulse = PYMAX(t, 0.0);
            lambda = new_impulse - p->normal_impulse;
            v2 Pv = vscale(lambda, normal);
            va = vsub(va, vscale(ima, Pv));
            wa = wa - iia * vcross(p->ra, Pv);
            vb = vadd(vb, vscale(imb, Pv));
            wb = wb + iib * vcross(p->rb, Pv);
            p->normal_impulse = new_impulse;
        } else if (cc->npoints == 2) {
            vcp_t *cp1 = &cc->p[0], *cp2 = &cc->p[1];
            v2 a = V(cp1->normal_impulse, cp2->normal_impulse);
            v2 dv1 = rel_vel(vb, wb, cp1->rb, va, wa, cp1->ra);
            v2 dv2 = rel_vel(vb, wb, cp2->rb, va, wa, cp2->ra);
            double vn1 = vdot(dv1, normal), vn2 = vdot(dv2, normal);
            v2 b = V(vn1 - cp1->velocity_bias, vn2 - cp2->velocity_bias);
            /* b -= K * a  with K = Mat22((k11,k12),(k12,k22)) */
            v2 Ka = V(cc->k11 * a.x + cc->k12 * a.y, cc->k12 * a.x + cc->k22 * a.y);
            b = vsub(b, Ka);
            v2 x; int solved = 0;
            /* case 1 */
            {
                v2 nb = V(cc->nm_c1x * b.x + cc->nm_c2x * b.y, cc->nm_c1y * b.x + cc->nm_c2y * b.y);
                x = vneg(nb);
                if (x.x >= 0.0 && x.y >= 0.0) solved = 1;
            }
            if (!solved) {                                   /* case 2 */
                x = V(-cp1->normal_mass * b.x, 0.0);
                double vn2c = cc->k12 * x.x + b.y;          /* K.col1.y */
                if (x.x >= 0.0 && vn2c >= 0.0) solved = 1;
            }
            if (!solved) {                                   /* case 3 */
                x = V(0.0, -cp2->normal_mass * b.y);
                double vn1c = cc->k12 * x.y + b.x;          /* K.col2.x */
                if (x.y >= 0.0 && vn1c >= 0.0) solved = 1;
            }
            if (!solved) {                                   /* case 4 */
                x = V(0.0, 0.0);
                if (b.x >= 0.0 && b.y >= 0.0) solved = 1;
            }
            if (solved) {
                v2 d = vsub(x, a);
                v2 p1 = vscale(d.x, normal), p2 = vscale(d.y, normal);
                v2 p12 = vadd(p1, p2);
                va = vsub(va, vscale(ima, p12));
                wa = wa - iia * (vcross(cp1->ra, p1) + vcross(cp2->ra, p2));
                vb = vadd(vb, vscale(imb, p12));
                wb = wb + iib * (vcross(cp1->rb, p1) + vcross(cp2->rb, p2));
                cp1->normal_impulse = x.x;
                cp2->normal_impulse = x.y;
            }
        }
        Vv[cc->ia].v = va; Vv[cc->ia].w = wa;
        Vv[cc->ib].v = vb; Vv[cc->ib].w = wb;
    }
}


/* =====================================================================================================
 * Joints in the island solver: joints/revolute.py, joints/distance.py, joints/weld.py on the
 * pure-Python Vec3 / Mat33 of common.py:227-358, 533-614.  One joint = row k of j_def (type, body_a, body_b,
 * flags | _limit_state << 8), j_par (parameters, layout in pypybox2d_b200/joints.py), j_imp (_impulse x, y, z,
 * _motor_impulse) and j_tmp (the attributes _init_velocity_constraints leaves for the sweeps).
 * ===================================================================================================== */
#define J_TMP 28                                    /* reals per j_tmp row (pypybox2d_b200/schema.py JOINT_TMP) */
#define J_REVOLUTE 1
#define J_DISTANCE 2
#define J_WELD     3
#define J_PRISMATIC 4
#define J_ROPE     5
#define J_FRICTION 6
#define J_WHEEL    7
#define J_PULLEY   8
#define J_MOUSE    9
#define JF_COLLIDE 0x1
#define JF_LIMIT   0x2
#define JF_MOTOR   0x4
#define JF_ISLAND  0x8
#define JF_LS_SHIFT 8
#define LS_INACTIVE 0
#define LS_LOWER 1
#define LS_UPPER 2
#define LS_EQUAL 3
#define MAX_ANG_CORR (8.0 / 180.0 * PI_)            /* settings.py:92 */
#define ST_JOINT_OVERFLOW 0x20u

typedef struct { double x, y, z; } v3;
static inline v3 V3(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline double v3dot(v3 a, v3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }               /* common.py:342-344 */
static inline v3 v3cross(v3 s, v3 v) { return V3(s.y * v.z - s.z * v.y, s.z * v.x - s.x * v.z, s.x * v.y - s.y * v.x); } /* :353-358 */
static inline v3 v3neg(v3 a) { return V3(-a.x, -a.y, -a.z); }
static inline v3 v3add(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
typedef struct { v3 c1, c2, c3; } m33;
static v3 m33_solve3(const m33* m, v3 b)                                                                /* common.py:587-599 */
{
    double det = v3dot(m->c1, v3cross(m->c2, m->c3));
    if (det != 0.0) det = 1.0 / det;
    return V3(det * v3dot(b, v3cross(m->c2, m->c3)), det * v3dot(m->c1, v3cross(b, m->c3)), det * v3dot(m->c1, v3cross(m->c2, b)));
}
static v2 m33_solve2(const m33* m, v2 vec)                                                              /* common.py:601-614 */
{
    double a11 = m->c1.x, a12 = m->c2.x, a21 = m->c1.y, a22 = m->c2.y;
    double det = a11 * a22 - a12 * a21;
    if (det != 0.0) det = 1.0 / det;
    return V(det * (a22 * vec.x - a12 * vec.y), det * (a11 * vec.y - a21 * vec.x));
}
static inline xform rot_make(double a) { return xf_make(V(0, 0), cos(a), sin(a)); }                     /* Mat22(angle=a) */
static inline m33 m33_load(const double* t) { m33 m; m.c1 = V3(t[0], t[1], t[2]); m.c2 = V3(t[3], t[4], t[5]); m.c3 = V3(t[6], t[7], t[8]); return m; }
static inline void m33_store(double* t, const m33* m)
{
    t[0] = m->c1.x; t[1] = m->c1.y; t[2] = m->c1.z; t[3] = m->c2.x; t[4] = m->c2.y; t[5] = m->c2.z; t[6] = m->c3.x; t[7] = m->c3.y; t[8] = m->c3.z;
}


/* MouseJoint (mouse.py:138-229): one body (j_def body_a = -1); par: [2:4] local_anchor_b, [4:6] target, [6] max_force, [7] frequency,
 * [8] damping_ratio, [9] Body.mass; tmp: [2:4] rb, [4:8] the inverse of K by columns, [8:10] c, [10] gamma */
static void mouse_init(orc_world* w, int k, double dt, double dt_ratio, int warm, pos_t* P, vel_t* Vv)
{
    const int32_t* def = w->j_def + 4 * (size_t)k;
    const double* par = w->j_par + 12 * (size_t)k;
    double* imp = w->j_imp + 4 * (size_t)k;
    double* tmp = w->j_tmp + J_TMP * (size_t)k;
    int bb = def[2], ib_ = w->b_island[bb];
    v2 cb = P[ib_].c; double ab = P[ib_].a;
    v2 vb = Vv[ib_].v; double wb = Vv[ib_].w;
    double mb = w->b_mass[4 * bb], ib = w->b_mass[4 * bb + 1];
    v2 lcb = V(w->b_mass[4 * bb + 2], w->b_mass[4 * bb + 3]);
    xform qb = rot_make(ab);
    double mass = par[9];
    double omega = 2.0 * PI_ * par[7];
    double d = 2.0 * mass * par[8] * omega;
    double kk = mass * sq(omega);
    double gamma = dt * (d + dt * kk);
    if (gamma != 0.0) gamma = 1.0 / gamma;
    double beta = dt * kk * gamma;
    v2 rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
    double k1x = mb + ib * sq(rb.y) + gamma, k1y = -ib * rb.x * rb.y;
    double k2x = k1y, k2y = mb + ib * sq(rb.x) + gamma;
    double det = k1x * k2y - k2x * k1y;
    if (det != 0.0) det = 1.0 / det;
    tmp[2] = rb.x; tmp[3] = rb.y;
    tmp[4] = det * k2y; tmp[5] = -det * k1y; tmp[6] = -det * k2x; tmp[7] = det * k1x;
    v2 c = vscale(beta, vsub(vadd(cb, rb), V(par[4], par[5])));
    tmp[8] = c.x; tmp[9] = c.y; tmp[10] = gamma;
    wb *= 0.98;
    if (warm) {
        imp[0] = dt_ratio * imp[0]; imp[1] = dt_ratio * imp[1];
        v2 I = V(imp[0], imp[1]);
        vb = vadd(vb, vscale(mb, I));
        wb += ib * vcross(rb, I);
    } else { imp[0] = 0.0; imp[1] = 0.0; }
    Vv[ib_].v = vb; Vv[ib_].w = wb;
}
static void mouse_solve_velocity(orc_world* w, int k, double dt, vel_t* Vv)
{
    const int32_t* def = w->j_def + 4 * (size_t)k;
    const double* par = w->j_par + 12 * (size_t)k;
    double* imp = w->j_imp + 4 * (size_t)k;
    const double* tmp = w->j_tmp + J_TMP * (size_t)k;
    int bb = def[2], ib_ = w->b_island[bb];
    v2 vb = Vv[ib_].v; double wb = Vv[ib_].w;
    double mb = w->b_mass[4 * bb], ib = w->b_mass[4 * bb + 1];
    v2 rb = V(tmp[2], tmp[3]);
    v2 Cdot = vadd(vb, scross(wb, rb));
    v2 rhs = vneg(vadd(vadd(Cdot, V(tmp[8], tmp[9])), vscale(tmp[10], V(imp[0], imp[1]))));
    v2 impulse = V(tmp[4] * rhs.x + tmp[6] * rhs.y, tmp[5] * rhs.x + tmp[7] * rhs.y);
    v2 old_impulse = V(imp[0], imp[1]);
    v2 cur = vadd(old_impulse, impulse);
    double max_impulse = dt * par[6];
    if (vlen2(cur) > sq(max_impulse)) { double f = max_impulse / sqrt(sq(cur.x) + sq(cur.y)); cur = V(f * cur.x, f * cur.y); }
    imp[0] = cur.x; imp[1] = cur.y;
    impulse = vsub(cur, old_impulse);
    vb = vadd(vb, vscale(mb, impulse));
    wb += ib * vcross(rb, impulse);
    Vv[ib_].v = vb; Vv[ib_].w = wb;
}

/* _init_velocity_constraints: revolute.py:208-300, distance.py:155-250, weld.py:106-169 */
static void joint_init(orc_world* w, int k, double dt, double dt_ratio, int warm, pos_t* P, vel_t* Vv)
{
    if (w->j_def[4 * (size_t)k] == J_MOUSE) { mouse_init(w, k, dt, dt_ratio, warm, P, Vv); return; }
    int32_t* def = w->j_def + 4 * (size_t)k;
    const double* par = w->j_par + 12 * (size_t)k;
    double* imp = w->j_imp + 4 * (size_t)k;
    double* tmp = w->j_tmp + J_TMP * (size_t)k;
    int type = def[0], ba = def[1], bb = def[2], flags = def[3];
    int ia_ = w->b_island[ba], ib_ = w->b_island[bb];
    v2 lca = V(w->b_mass[4 * ba + 2], w->b_mass[4 * ba + 3]), lcb = V(w->b_mass[4 * bb + 2], w->b_mass[4 * bb + 3]);
    double ma = w->b_mass[4 * ba], mb = w->b_mass[4 * bb], ia = w->b_mass[4 * ba + 1], ib = w->b_mass[4 * bb + 1];
    v2 ca = P[ia_].c, cb = P[ib_].c; double aa = P[ia_].a, ab = P[ib_].a;
    v2 va = Vv[ia_].v, vb = Vv[ib_].v; double wa = Vv[ia_].w, wb = Vv[ib_].w;
    xform qa = rot_make(aa), qb = rot_make(ab);
    v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
    tmp[0] = ra.x; tmp[1] = ra.y; tmp[2] = rb.x; tmp[3] = rb.y;
    if (type == J_REVOLUTE) {
        m33 m;
        m.c3 = V3(-ra.y * ia - rb.y * ib, ra.x * ia + rb.x * ib, ia + ib);
        m.c2 = V3(-ra.y * ra.x * ia - rb.y * rb.x * ib, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, m.c3.y);
        m.c1 = V3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, m.c2.x, m.c3.x);
        m33_store(tmp + 4, &m);
        double motor_mass = ia + ib;
        int fixed_rotation = (motor_mass == 0.0);
        if (motor_mass > 0.0) motor_mass = 1.0 / motor_mass;
        tmp[13] = motor_mass;
        int ls = (flags >> JF_LS_SHIFT) & 3;
        if (!(flags & JF_MOTOR) || fixed_rotation) imp[3] = 0.0;
        if ((flags & JF_LIMIT) && !fixed_rotation) {
            double joint_angle = ab - aa - par[4];
            if (fabs(par[6] - par[5]) < 2.0 * ANGULAR_SLOP) ls = LS_EQUAL;
            else if (joint_angle <= par[5]) {
                if (ls != LS_LOWER) imp[2] = imp[2] - imp[2];
                ls = LS_LOWER;
            } else if (joint_angle >= par[6]) {
                if (ls != LS_UPPER) imp[2] = imp[2] - imp[2];
                ls = LS_UPPER;
            } else {
                ls = LS_INACTIVE;
                imp[2] = imp[2] - imp[2];
            }
        } else ls = LS_INACTIVE;
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (warm) {
            imp[0] = dt_ratio * imp[0]; imp[1] = dt_ratio * imp[1]; imp[2] = dt_ratio * imp[2];
            imp[3] *= dt_ratio;
            v2 Pv = V(imp[0], imp[1]);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * (vcross(ra, Pv) + imp[3] + imp[2]);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * (vcross(rb, Pv) + imp[3] + imp[2]);
        } else { imp[0] = imp[1] = imp[2] = 0.0; imp[3] = 0.0; }
    } else if (type == J_DISTANCE) {
        v2 u = vsub(vsub(vadd(cb, rb), ca), ra);
        double length = sqrt(sq(u.x) + sq(u.y));
        v2 un = (length > LINEAR_SLOP) ? vscale(1.0 / length, u) : V(0.0, 0.0);     /* self._u; the local `u` stays as it was */
        tmp[4] = un.x; tmp[5] = un.y;
        double cr_au = sq(vcross(ra, u)), cr_bu = sq(vcross(rb, u));
        double inv_mass = ma + ia * cr_au + mb + ib * cr_bu;
        double mass = (inv_mass != 0.0) ? 1.0 / inv_mass : 0.0;
        double gamma = tmp[7], bias = tmp[8];                                       /* attributes persist when frequency == 0 */
        if (!(par[5] > 0.0)) { gamma = 0.0; bias = 0.0; }                           /* (never assigned: stay at the constructor's 0) */
        if (par[5] > 0.0) {
            double Cc = length - par[4];
            double omega = 2.0 * PI_ * par[5];
            double d = 2.0 * mass * par[6] * omega;
            double kk = mass * sq(omega);
            gamma = dt * (d + dt * kk);
            gamma = (gamma != 0.0) ? 1.0 / gamma : 0.0;
            bias = Cc * dt * kk * gamma;
            double m2 = inv_mass + gamma;
            mass = (m2 != 0.0) ? 1.0 / m2 : 0.0;
        }
        tmp[6] = mass; tmp[7] = gamma; tmp[8] = bias;
        if (warm) {
            imp[0] *= dt_ratio;
            v2 Pv = vscale(imp[0], un);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * vcross(ra, Pv);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * vcross(rb, Pv);
        } else imp[0] = 0.0;
    } else if (type == J_WELD) {
        m33 m;
        double c2x = -ra.y * ra.x * ia - rb.y * rb.x * ib;
        double c3x = -ra.y * ia - rb.y * ib;
        m.c1 = V3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, c2x, c3x);
        m.c2 = V3(c2x, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, c3x);          /* weld.py:148-150: c3x, as written */
        m.c3 = V3(c3x, ra.x * ia + rb.x * ib, ia + ib);
        m33_store(tmp + 4, &m);
        if (warm) {
            imp[0] = dt_ratio * imp[0]; imp[1] = dt_ratio * imp[1]; imp[2] = dt_ratio * imp[2];
            v2 Pv = V(imp[0], imp[1]);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * (vcross(ra, Pv) + imp[2]);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * (vcross(rb, Pv) + imp[2]);
        } else imp[0] = imp[1] = imp[2] = 0.0;
    }
    else if (type == J_PRISMATIC) {                                                  /* prismatic.py:232-335 */
        v2 d = vsub(vsub(vadd(cb, rb), ca), ra);
        v2 lx = V(par[9], par[10]), ly = scross(1.0, lx);
        v2 axis = rot_mul(&qa, lx);
        double a1 = vcross(vadd(d, ra), axis), a2 = vcross(rb, axis);
        double motor_mass = ma + mb + ia * sq(a1) + ib * sq(a2);
        if (motor_mass > 0.0) motor_mass = 1.0 / motor_mass;
        v2 perp = rot_mul(&qa, ly);
        double s1 = vcross(vadd(d, ra), perp), s2 = vcross(rb, perp);
        double k11 = ma + mb + ia * sq(s1) + ib * sq(s2);
        double k12 = ia * s1 + ib * s2;
        double k13 = ia * s1 * a1 + ib * s2 * a2;
        double k22 = ia + ib;
        if (k22 == 0.0) k22 = 1.0;
        double k23 = ia * a1 + ib * a2;
        double k33 = ma + mb + ia * sq(a1) + ib * sq(a2);
        tmp[4] = axis.x; tmp[5] = axis.y; tmp[6] = perp.x; tmp[7] = perp.y; tmp[8] = s1; tmp[9] = s2; tmp[10] = a1; tmp[11] = a2;
        tmp[12] = motor_mass; tmp[13] = k11; tmp[14] = k12; tmp[15] = k13; tmp[16] = k22; tmp[17] = k23; tmp[18] = k33;
        int ls = (flags >> JF_LS_SHIFT) & 3;
        if (flags & JF_LIMIT) {
            double joint_translation = vdot(axis, d);
            if (fabs(par[6] - par[5]) < (2.0 * LINEAR_SLOP)) ls = LS_EQUAL;
            else if (joint_translation <= par[5]) {
                if (ls != LS_LOWER) { ls = LS_LOWER; imp[2] = imp[2] - imp[2]; }
            } else if (joint_translation >= par[6]) {
                if (ls != LS_UPPER) { ls = LS_UPPER; imp[2] = imp[2] - imp[2]; }
            } else { ls = LS_INACTIVE; imp[2] = imp[2] - imp[2]; }
        } else { ls = LS_INACTIVE; imp[2] = imp[2] - imp[2]; }
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (!(flags & JF_MOTOR)) imp[3] = 0.0;
        if (warm) {
            imp[0] = dt_ratio * imp[0]; imp[1] = dt_ratio * imp[1]; imp[2] = dt_ratio * imp[2];
            imp[3] *= dt_ratio;
            double ix = imp[0], iy = imp[1], iz = imp[2], mi = imp[3];
            v2 Pv = vadd(vscale(ix, perp), vscale(mi + iz, axis));
            double L1 = ix * s1 + iy + (mi + iz) * a1;
            double L2 = ix * s2 + iy + (mi + iz) * a2;
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * L1;
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * L2;
        } else { imp[0] = imp[1] = imp[2] = 0.0; imp[3] = 0.0; }
    }
    else if (type == J_ROPE) {                                                       /* rope.py:106-176 */
        v2 u0 = vsub(vsub(vadd(cb, rb), ca), ra);                                    /* self._u: never normalised (`u /= ...` rebinds the local) */
        double length = sqrt(sq(u0.x) + sq(u0.y));
        tmp[4] = u0.x; tmp[5] = u0.y; tmp[7] = length;
        double Cc = length - par[4];
        int ls = (Cc > 0.0) ? LS_UPPER : LS_INACTIVE;
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (!(length > LINEAR_SLOP)) { tmp[6] = 0.0; imp[0] = 0.0; return; }        /* rope.py:147-151: returns before the velocities are stored */
        v2 u = V(u0.x / length, u0.y / length);
        double cra = vcross(ra, u), crb = vcross(rb, u);
        double inv_mass = ma + ia * cra * cra + mb + ib * crb * crb;
        tmp[6] = (inv_mass != 0.0) ? 1.0 / inv_mass : 0.0;
        if (warm) {
            imp[0] *= dt_ratio;
            v2 Pv = vscale(imp[0], u0);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * vcross(ra, Pv);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * vcross(rb, Pv);
        } else imp[0] = 0.0;
    }
    else if (type == J_FRICTION) {                                                   /* friction.py:105-176 */
        double k1x = ma + mb + ia * ra.y * ra.y + ib * rb.y * rb.y;
        double k1y = -ia * ra.x * ra.y - ib * rb.x * rb.y;
        double k2x = k1y;
        double k2y = ma + mb + ia * ra.x * ra.x + ib * rb.x * rb.x;
        double det = k1x * k2y - k2x * k1y;                                          /* Mat22.inverse: common.py:506-519 */
        if (det != 0.0) det = 1.0 / det;
        tmp[4] = det * k2y; tmp[5] = -det * k1y; tmp[6] = -det * k2x; tmp[7] = det * k1x;
        double angular_mass = ia + ib;
        if (angular_mass > 0.0) angular_mass = 1.0 / angular_mass;
        tmp[8] = angular_mass;
        if (warm) {
            imp[0] = dt_ratio * imp[0]; imp[1] = dt_ratio * imp[1];
            imp[2] *= dt_ratio;
            v2 Pv = V(imp[0], imp[1]);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * (vcross(ra, Pv) + imp[2]);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * (vcross(rb, Pv) + imp[2]);
        } else { imp[0] = imp[1] = 0.0; imp[2] = 0.0; }
    }
    else if (type == J_WHEEL) {                                                      /* wheel.py:202-318 */
        v2 d = vsub(vsub(vadd(cb, rb), ca), ra);
        v2 lx = V(par[4], par[5]), ly = scross(1.0, lx);
        v2 ay = rot_mul(&qa, ly);
        double s_ay = vcross(vadd(d, ra), ay), s_by = vcross(rb, ay);
        double mass = ma + mb + ia * sq(s_ay) + ib * sq(s_by);
        if (mass > 0.0) mass = 1.0 / mass;
        double spring_mass = 0.0, bias = 0.0, gamma = 0.0;
        v2 ax = V(tmp[9], tmp[10]); double s_ax = tmp[11], s_bx = tmp[12];          /* attributes keep their last values when frequency <= 0 */
        if (par[6] > 0.0) {
            ax = rot_mul(&qa, lx);
            s_ax = vcross(vadd(d, ra), ax); s_bx = vcross(rb, ax);
            double inv_mass = ma + mb + ia * sq(s_ax) + ib * sq(s_bx);
            if (inv_mass > 0.0) {
                spring_mass = 1.0 / inv_mass;
                double Cc = vdot(d, ax);
                double omega = 2.0 * PI_ * par[6];
                double dd = 2.0 * spring_mass * par[7] * omega;
                double kk = spring_mass * sq(omega);
                gamma = dt * (dd + dt * kk);
                if (gamma > 0.0) gamma = 1.0 / gamma;
                bias = Cc * dt * kk * gamma;
                spring_mass = inv_mass + gamma;
                if (spring_mass > 0.0) spring_mass = 1.0 / spring_mass;
            }
        } else imp[1] = 0.0;
        double motor_mass;
        if (flags & JF_MOTOR) {
            motor_mass = ia + ib;
            if (motor_mass > 0.0) motor_mass = 1.0 / motor_mass;
        } else { motor_mass = 0.0; imp[2] = 0.0; }
        tmp[4] = ay.x; tmp[5] = ay.y; tmp[6] = s_ay; tmp[7] = s_by; tmp[8] = mass; tmp[9] = ax.x; tmp[10] = ax.y; tmp[11] = s_ax; tmp[12] = s_bx;
        tmp[13] = spring_mass; tmp[14] = bias; tmp[15] = gamma; tmp[16] = motor_mass;
        if (warm) {
            imp[0] *= dt_ratio; imp[1] *= dt_ratio; imp[2] *= dt_ratio;
            v2 Pv = vadd(vscale(imp[0], ay), vscale(imp[1], ax));
            double La = imp[0] * s_ay + imp[1] * s_ax + imp[2];
            double Lb = imp[0] * s_by + imp[1] * s_bx + imp[2];
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * La;
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * Lb;
        } else { imp[0] = 0.0; imp[1] = 0.0; imp[2] = 0.0; }
    }
    else if (type == J_PULLEY) {                                                     /* pulley.py:156-228 */
        v2 ua = vsub(vadd(ca, ra), V(par[4], par[5])), ub = vsub(vadd(cb, rb), V(par[6], par[7]));
        double length_a = sqrt(sq(ua.x) + sq(ua.y)), length_b = sqrt(sq(ub.x) + sq(ub.y));
        ua = (length_a > 10.0 * LINEAR_SLOP) ? V(ua.x / length_a, ua.y / length_a) : V(0.0, 0.0);
        ub = (length_b > 10.0 * LINEAR_SLOP) ? V(ub.x / length_b, ub.y / length_b) : V(0.0, 0.0);
        tmp[4] = ua.x; tmp[5] = ua.y; tmp[6] = ub.x; tmp[7] = ub.y;
        double rua = vcross(ra, ua), rub = vcross(rb, ub);
        double ema = ma + ia * sq(rua), emb = mb + ib * sq(rub);                     /* the locals `ma`, `mb` are rebound: the warm start below uses them */
        double mass = ema + par[8] * par[8] * emb;
        if (mass > 0.0) mass = 1.0 / mass;
        tmp[8] = mass;
        if (warm) {
            imp[0] *= dt_ratio;
            v2 Pa = vscale(-imp[0], ua), Pb = vscale(-par[8] * imp[0], ub);
            va = vadd(va, vscale(ema, Pa));
            wa += ia * vcross(ra, Pa);
            vb = vadd(vb, vscale(emb, Pb));
            wb += ib * vcross(rb, Pb);
        } else imp[0] = 0.0;
    }
    Vv[ia_].v = va; Vv[ia_].w = wa;
    Vv[ib_].v = vb; Vv[ib_].w = wb;
}

static inline m33 m33_sym(const double* k)     /* Mat33((k11, k12, k13), (k12, k22, k23), (k13, k23, k33)) */
{
    m33 m; m.c1 = V3(k[0], k[1], k[2]); m.c2 = V3(k[1], k[3], k[4]); m.c3 = V3(k[2], k[4], k[5]); return m;
}

/* _solve_velocity_constraints: revolute.py:302-378, distance.py:252-275, weld.py:171-199 */
static void joint_solve_velocity(orc_world* w, int k, double dt, double inv_dt, vel_t* Vv)
{
    if (w->j_def[4 * (size_t)k] == J_MOUSE) { mouse_solve_velocity(w, k, dt, Vv); return; }
    const int32_t* def = w->j_def + 4 * (size_t)k;
    const double* par = w->j_par + 12 * (size_t)k;
    double* imp = w->j_imp + 4 * (size_t)k;
    const double* tmp = w->j_tmp + J_TMP * (size_t)k;
    int type = def[0], ba = def[1], bb = def[2], flags = def[3];
    int ia_ = w->b_island[ba], ib_ = w->b_island[bb];
    double ma = w->b_mass[4 * ba], mb = w->b_mass[4 * bb], ia = w->b_mass[4 * ba + 1], ib = w->b_mass[4 * bb + 1];
    v2 va = Vv[ia_].v, vb = Vv[ib_].v; double wa = Vv[ia_].w, wb = Vv[ib_].w;
    v2 ra = V(tmp[0], tmp[1]), rb = V(tmp[2], tmp[3]);
    if (type == J_REVOLUTE) {
        m33 m = m33_load(tmp + 4);
        int ls = (flags >> JF_LS_SHIFT) & 3;
        int fixed_rotation = ((ia + ib) == 0.0);
        if ((flags & JF_MOTOR) && ls != LS_EQUAL && !fixed_rotation) {
            double Cdot = wb - wa - par[8];
            double impulse = tmp[13] * (-Cdot);
            double old_impulse = imp[3];
            double max_impulse = dt * par[7];
            imp[3] = pyclamp(imp[3] + impulse, -max_impulse, max_impulse);
            impulse = imp[3] - old_impulse;
            wa -= ia * impulse;
            wb += ib * impulse;
        }
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE && !fixed_rotation) {
            v2 Cdot1 = vsub(vsub(vadd(vb, scross(wb, rb)), va), scross(wa, ra));
            double Cdot2 = wb - wa;
            v3 Cdot = V3(Cdot1.x, Cdot1.y, Cdot2);
            v3 impulse = v3neg(m33_solve3(&m, Cdot));
            if (ls == LS_EQUAL) {
                imp[0] += impulse.x; imp[1] += impulse.y; imp[2] += impulse.z;
            } else if (ls == LS_LOWER || ls == LS_UPPER) {
                double new_impulse = imp[2] + impulse.z;
                if ((ls == LS_LOWER) ? (new_impulse < 0.0) : (new_impulse > 0.0)) {
                    v2 rhs = vadd(vneg(Cdot1), vscale(imp[2], V(m.c3.x, m.c3.y)));
                    v2 reduced = m33_solve2(&m, rhs);
                    impulse = V3(reduced.x, reduced.y, -imp[2]);
                    imp[0] += reduced.x; imp[1] += reduced.y; imp[2] += -imp[2];
                } else {
                    imp[0] += impulse.x; imp[1] += impulse.y; imp[2] += impulse.z;
                }
            }
            v2 Pv = V(impulse.x, impulse.y);
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * (vcross(ra, Pv) + impulse.z);
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * (vcross(rb, Pv) + impulse.z);
        } else {
            v2 Cdot = vsub(vsub(vadd(vb, scross(wb, rb)), va), scross(wa, ra));
            v2 impulse = m33_solve2(&m, vneg(Cdot));
            imp[0] += impulse.x; imp[1] += impulse.y; imp[2] += 0.0;
            va = vsub(va, vscale(ma, impulse));
            wa -= ia * vcross(ra, impulse);
            vb = vadd(vb, vscale(mb, impulse));
            wb += ib * vcross(rb, impulse);
        }
    } else if (type == J_DISTANCE) {
        v2 u = V(tmp[4], tmp[5]);
        v2 vpa = vadd(va, scross(wa, ra)), vpb = vadd(vb, scross(wb, rb));
        double Cdot = vdot(u, vsub(vpb, vpa));
        double impulse = -tmp[6] * (Cdot + tmp[8] + tmp[7] * imp[0]);
        imp[0] += impulse;
        v2 Pv = vscale(impulse, u);
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * vcross(ra, Pv);
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * vcross(rb, Pv);
    } else if (type == J_WELD) {
        m33 m = m33_load(tmp + 4);
        v2 Cdot1 = vsub(vsub(vadd(vb, scross(wb, rb)), va), scross(wa, ra));
        double Cdot2 = wb - wa;
        v3 impulse = m33_solve3(&m, v3neg(V3(Cdot1.x, Cdot1.y, Cdot2)));
        imp[0] += impulse.x; imp[1] += impulse.y; imp[2] += impulse.z;
        v2 Pv = V(impulse.x, impulse.y);
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * (vcross(ra, Pv) + impulse.z);
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * (vcross(rb, Pv) + impulse.z);
    }
    else if (type == J_PRISMATIC) {                                                  /* prismatic.py:337-419 */
        v2 axis = V(tmp[4], tmp[5]), perp = V(tmp[6], tmp[7]);
        double s1 = tmp[8], s2 = tmp[9], a1 = tmp[10], a2 = tmp[11];
        m33 m = m33_sym(tmp + 13);
        int ls = (flags >> JF_LS_SHIFT) & 3;
        if ((flags & JF_MOTOR) && ls != LS_EQUAL) {
            double Cdot = vdot(axis, vsub(vb, va)) + a2 * wb - a1 * wa;
            double impulse = tmp[12] * (par[8] - Cdot);
            double old_impulse = imp[3];
            double max_impulse = dt * par[7];
            imp[3] = pyclamp(imp[3] + impulse, -max_impulse, max_impulse);
            impulse = imp[3] - old_impulse;
            v2 Pv = vscale(impulse, axis);
            double L1 = impulse * a1, L2 = impulse * a2;
            va = vsub(va, vscale(ma, Pv));
            wa -= ia * L1;
            vb = vadd(vb, vscale(mb, Pv));
            wb += ib * L2;
        }
        v2 Cdot1 = V(vdot(perp, vsub(vb, va)) + s2 * wb - s1 * wa, wb - wa);
        v2 Pv; double L1, L2;
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE) {
            double Cdot2 = vdot(axis, vsub(vb, va)) + a2 * wb - a1 * wa;
            v3 old_impulse = V3(imp[0], imp[1], imp[2]);
            v3 df = m33_solve3(&m, v3neg(V3(Cdot1.x, Cdot1.y, Cdot2)));
            v3 cur = v3add(old_impulse, df);
            double iz;
            if (ls == LS_LOWER) iz = PYMAX(cur.z, 0.0);
            else if (ls == LS_UPPER) iz = PYMIN(cur.z, 0.0);
            else iz = cur.z;
            v2 b = vsub(vneg(Cdot1), vscale(cur.z - old_impulse.z, V(m.c3.x, m.c3.y)));
            v2 f2r = vadd(m33_solve2(&m, b), V(old_impulse.x, old_impulse.y));
            imp[0] = f2r.x; imp[1] = f2r.y; imp[2] = iz;
            df = V3(f2r.x - old_impulse.x, f2r.y - old_impulse.y, iz - old_impulse.z);
            Pv = vadd(vscale(df.x, perp), vscale(df.z, axis));
            L1 = df.x * s1 + df.y + df.z * a1;
            L2 = df.x * s2 + df.y + df.z * a2;
        } else {
            v2 df = m33_solve2(&m, vneg(Cdot1));
            imp[0] += df.x; imp[1] += df.y; imp[2] += 0.0;
            Pv = vscale(df.x, perp);
            L1 = df.x * s1 + df.y;
            L2 = df.x * s2 + df.y;
        }
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * L1;
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * L2;
    }
    else if (type == J_ROPE) {                                                       /* rope.py:178-207 */
        v2 u0 = V(tmp[4], tmp[5]);
        v2 vpa = vadd(va, scross(wa, ra)), vpb = vadd(vb, scross(wb, rb));
        double Cc = tmp[7] - par[4];
        double Cdot = vdot(u0, vsub(vpb, vpa));
        if (Cc < 0.0) Cdot += inv_dt * Cc;
        double impulse = -tmp[6] * Cdot;
        double old_impulse = imp[0];
        imp[0] = PYMIN(0.0, imp[0] + impulse);
        impulse = imp[0] - old_impulse;
        v2 Pv = vscale(impulse, u0);
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * vcross(ra, Pv);
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * vcross(rb, Pv);
    }
    else if (type == J_FRICTION) {                                                   /* friction.py:178-230: the result goes to `positions`, */
        double Cdot = wb - wa;                                                       /* i.e. nowhere; only the accumulated impulses change */
        double impulse = -tmp[8] * Cdot;
        double old_impulse = imp[2];
        double max_impulse = dt * par[5];
        imp[2] = pyclamp(imp[2] + impulse, -max_impulse, max_impulse);
        impulse = imp[2] - old_impulse;
        wa -= ia * impulse;
        wb += ib * impulse;
        v2 Cd = vsub(vsub(vadd(vb, scross(wb, rb)), va), scross(wa, ra));
        v2 li = vneg(V(tmp[4] * Cd.x + tmp[6] * Cd.y, tmp[5] * Cd.x + tmp[7] * Cd.y));
        v2 lin = vadd(V(imp[0], imp[1]), li);
        max_impulse = dt * par[4];
        if (vlen2(lin) > sq(max_impulse)) { v2 n = vnorm(lin); lin = V(max_impulse * n.x, max_impulse * n.y); }
        imp[0] = lin.x; imp[1] = lin.y;
        return;
    }
    else if (type == J_WHEEL) {                                                      /* wheel.py:320-373 */
        v2 ay = V(tmp[4], tmp[5]), ax = V(tmp[9], tmp[10]);
        double s_ay = tmp[6], s_by = tmp[7], s_ax = tmp[11], s_bx = tmp[12];
        double Cdot = vdot(ax, vsub(vb, va)) + s_bx * wb - s_ax * wa;
        double impulse = -tmp[13] * (Cdot + tmp[14] + tmp[15] * imp[1]);
        imp[1] += impulse;
        v2 Pv = vscale(impulse, ax);
        double La = impulse * s_ax, Lb = impulse * s_bx;
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * La;
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * Lb;
        Cdot = wb - wa - par[9];
        impulse = -tmp[16] * Cdot;
        double old_impulse = imp[2];
        double max_impulse = dt * par[8];
        imp[2] = pyclamp(imp[2] + impulse, -max_impulse, max_impulse);
        impulse = imp[2] - old_impulse;
        wa -= ia * impulse;
        wb += ib * impulse;
        Cdot = vdot(ay, vsub(vb, va)) + s_by * wb - s_ay * wa;
        impulse = -tmp[8] * Cdot;
        imp[0] += impulse;
        Pv = vscale(impulse, ay);
        La = impulse * s_ay; Lb = impulse * s_by;
        va = vsub(va, vscale(ma, Pv));
        wa -= ia * La;
        vb = vadd(vb, vscale(mb, Pv));
        wb += ib * Lb;
    }
    else if (type == J_PULLEY) {                                                     /* pulley.py:230-256 */
        v2 ua = V(tmp[4], tmp[5]), ub = V(tmp[6], tmp[7]);
        v2 vpa = vadd(va, scross(wa, ra)), vpb = vadd(vb, scross(wb, rb));
        double Cdot = -vdot(ua, vpa) - par[8] * vdot(ub, vpb);
        double impulse = -tmp[8] * Cdot;
        imp[0] += impulse;
        v2 Pa = vscale(-impulse, ua), Pb = vscale(-par[8] * impulse, ub);
        va = vadd(va, vscale(ma, Pa));
        wa += ia * vcross(ra, Pa);
        vb = vadd(vb, vscale(mb, Pb));
        wb += ib * vcross(rb, Pb);
    }
    Vv[ia_].v = va; Vv[ia_].w = wa;
    Vv[ib_].v = vb; Vv[ib_].w = wb;
}

/* _solve_position_constraints: revolute.py:380-453, distance.py:277-311, weld.py:201-257; returns "within tolerance" */
static int joint_solve_position(orc_world* w, int k, pos_t* P)
{
    if (w->j_def[4 * (size_t)k] == J_MOUSE) return 1;                                  /* mouse.py:227-229 */
    const int32_t* def = w->j_def + 4 * (size_t)k;
    const double* par = w->j_par + 12 * (size_t)k;
    const double* tmp = w->j_tmp + J_TMP * (size_t)k;
    int type = def[0], ba = def[1], bb = def[2], flags = def[3];
    int ia_ = w->b_island[ba], ib_ = w->b_island[bb];
    v2 lca = V(w->b_mass[4 * ba + 2], w->b_mass[4 * ba + 3]), lcb = V(w->b_mass[4 * bb + 2], w->b_mass[4 * bb + 3]);
    double ma = w->b_mass[4 * ba], mb = w->b_mass[4 * bb], ia = w->b_mass[4 * ba + 1], ib = w->b_mass[4 * bb + 1];
    v2 ca = P[ia_].c, cb = P[ib_].c; double aa = P[ia_].a, ab = P[ib_].a;
    int ok = 1;
    if (type == J_REVOLUTE) {
        int ls = (flags >> JF_LS_SHIFT) & 3;
        int fixed_rotation = ((ia + ib) == 0.0);
        double angular_error = 0.0, position_error = 0.0;
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE && !fixed_rotation) {
            double angle = ab - aa - par[4];
            double limit_impulse = 0.0;
            if (ls == LS_EQUAL) {
                double Cc = pyclamp(angle - par[5], -MAX_ANG_CORR, MAX_ANG_CORR);
                limit_impulse = -tmp[13] * Cc;
                angular_error = fabs(Cc);
            } else if (ls == LS_LOWER) {
                double Cc = angle - par[5];
                angular_error = -Cc;
                Cc = pyclamp(Cc + ANGULAR_SLOP, -MAX_ANG_CORR, 0.0);
                limit_impulse = -tmp[13] * Cc;
            } else if (ls == LS_UPPER) {
                double Cc = angle - par[6];
                angular_error = Cc;
                Cc = pyclamp(Cc - ANGULAR_SLOP, 0.0, MAX_ANG_CORR);
                limit_impulse = -tmp[13] * Cc;
            }
            aa -= ia * limit_impulse;
            ab += ib * limit_impulse;
        }
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 Cc = vsub(vsub(vadd(cb, rb), ca), ra);
        position_error = sqrt(sq(Cc.x) + sq(Cc.y));
        double k1x = ma + mb + ia * ra.y * ra.y + ib * rb.y * rb.y;
        double k1y = -ia * ra.x * ra.y - ib * rb.x * rb.y;
        double k2x = k1y;
        double k2y = ma + mb + ia * ra.x * ra.x + ib * rb.x * rb.x;
        double det = k1x * k2y - k2x * k1y;                                          /* Mat22.solve: common.py:521-531 */
        if (det != 0.0) det = 1.0 / det;
        v2 impulse = vneg(V(det * (k2y * Cc.x - k2x * Cc.y), det * (k1x * Cc.y - k1y * Cc.x)));
        ca = vsub(ca, vscale(ma, impulse));
        aa -= ia * vcross(ra, impulse);
        cb = vadd(cb, vscale(mb, impulse));
        ab += ib * vcross(rb, impulse);
        ok = (position_error <= LINEAR_SLOP) && (angular_error <= ANGULAR_SLOP);
    } else if (type == J_DISTANCE) {
        if (par[5] > 0.0) return 1;
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 u = vsub(vsub(vadd(cb, rb), ca), ra);
        double length = sqrt(sq(u.x) + sq(u.y));
        u = vnorm(u);
        double Cc = length - par[4];
        double impulse = -tmp[6] * pyclamp(Cc, -MAX_LIN_CORR, MAX_LIN_CORR);
        v2 Pv = vscale(impulse, u);
        ca = vsub(ca, vscale(ma, Pv));
        aa -= ia * vcross(ra, Pv);
        cb = vadd(cb, vscale(mb, Pv));
        ab += ib * vcross(rb, Pv);
        ok = fabs(Cc) < LINEAR_SLOP;
    } else if (type == J_WELD) {
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 C1 = vsub(vsub(vadd(cb, rb), ca), ra);
        double C2 = ab - aa - par[4];
        double position_error = sqrt(sq(C1.x) + sq(C1.y)), angular_error = fabs(C2);
        m33 m;
        double c2x = -ra.y * ra.x * ia - rb.y * rb.x * ib;
        double c3x = -ra.y * ia - rb.y * ib;
        double c3y = ra.x * ia + rb.x * ib;
        m.c1 = V3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, c2x, c3x);
        m.c2 = V3(c2x, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, c3y);
        m.c3 = V3(-ra.y * ia - rb.y * ib, ra.x * ia + rb.x * ib, ia + ib);
        v3 impulse = v3neg(m33_solve3(&m, V3(C1.x, C1.y, C2)));
        v2 Pv = V(impulse.x, impulse.y);
        ca = vsub(ca, vscale(ma, Pv));
        aa -= ia * (vcross(ra, Pv) + impulse.z);
        cb = vadd(cb, vscale(mb, Pv));
        ab += ib * (vcross(rb, Pv) + impulse.z);
        ok = (position_error <= LINEAR_SLOP) && (angular_error <= ANGULAR_SLOP);
    }
    else if (type == J_PRISMATIC) {                                                  /* prismatic.py:421-513 */
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 d = vsub(vsub(vadd(cb, rb), ca), ra);
        v2 lx = V(par[9], par[10]), ly = scross(1.0, lx);
        v2 axis = rot_mul(&qa, lx);
        double a1 = vcross(vadd(d, ra), axis), a2 = vcross(rb, axis);
        v2 perp = rot_mul(&qa, ly);
        double s1 = vcross(vadd(d, ra), perp), s2 = vcross(rb, perp);
        v2 C1 = V(vdot(perp, d), ab - aa - par[4]);
        double linear_error = fabs(C1.x), angular_error = fabs(C1.y);
        int active = 0;
        double C2 = 0.0;
        if (flags & JF_LIMIT) {
            double translation = vdot(axis, d);
            if (fabs(par[6] - par[5]) < 2.0 * LINEAR_SLOP) {
                C2 = pyclamp(translation, -MAX_LIN_CORR, MAX_LIN_CORR);
                linear_error = PYMAX(linear_error, translation);
                active = 1;
            } else if (translation <= par[5]) {
                C2 = pyclamp(translation - par[5] + LINEAR_SLOP, -MAX_LIN_CORR, 0.0);
                linear_error = PYMAX(linear_error, par[5] - translation);
                active = 1;
            } else if (translation >= par[6]) {
                C2 = pyclamp(translation - par[6] - LINEAR_SLOP, 0.0, MAX_LIN_CORR);
                linear_error = PYMAX(linear_error, translation - par[6]);
                active = 1;
            }
        }
        v3 impulse;
        if (active) {
            double k[6];
            k[0] = ma + mb + ia * sq(s1) + ib * sq(s2);
            k[1] = ia * s1 + ib * s2;
            k[2] = ia * s1 * a1 + ib * s2 * a2;
            k[3] = ia + ib;
            if (k[3] == 0.0) k[3] = 1.0;
            k[4] = ia * a1 + ib * a2;
            k[5] = ma + mb + ia * sq(a1) + ib * sq(a2);
            m33 m = m33_sym(k);
            impulse = m33_solve3(&m, v3neg(V3(C1.x, C1.y, C2)));
        } else {
            double k11 = ma + mb + ia * s1 * s1 + ib * s2 * s2;
            double k12 = ia * s1 + ib * s2;
            double k22 = ia + ib;
            if (k22 == 0.0) k22 = 1.0;
            m33 m; m.c1 = V3(k11, k12, 0.0); m.c2 = V3(k12, k22, 0.0); m.c3 = V3(0.0, 0.0, 0.0);
            v2 impulse1 = m33_solve2(&m, vneg(C1));
            impulse = V3(impulse1.x, impulse1.y, 0.0);
        }
        v2 Pv = vadd(vscale(impulse.x, perp), vscale(impulse.z, axis));
        double L1 = impulse.x * s1 + impulse.y + impulse.z * a1;
        double L2 = impulse.x * s2 + impulse.y + impulse.z * a2;
        ca = vsub(ca, vscale(ma, Pv));
        cb = vadd(cb, vscale(mb, Pv));
        aa -= ia * L1;
        ab += ib * L2;
        ok = (linear_error <= LINEAR_SLOP) && (angular_error <= ANGULAR_SLOP);
    }
    else if (type == J_ROPE) {                                                       /* rope.py:209-248 */
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 u = vsub(vsub(vadd(cb, rb), ca), ra);
        double length = sqrt(sq(u.x) + sq(u.y));
        u = vnorm(u);
        double Cc = pyclamp(length - par[4], 0.0, MAX_LIN_CORR);
        double impulse = -tmp[6] * Cc;
        v2 Pv = vscale(impulse, u);
        ca = vsub(ca, vscale(ma, Pv));
        aa -= ia * vcross(ra, Pv);
        cb = vadd(cb, vscale(mb, Pv));
        ab += ib * vcross(rb, Pv);
        ok = (length - par[4]) < LINEAR_SLOP;
    }
    else if (type == J_FRICTION) return 1;                                           /* friction.py:232-233 */
    else if (type == J_WHEEL) {                                                      /* wheel.py:375-416 */
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 d = vsub(vadd(vsub(cb, ca), rb), ra);
        double kk = ma + mb + ia * sq(tmp[6]) + ib * sq(tmp[7]);
        v2 ly = scross(1.0, V(par[4], par[5]));
        v2 ay = rot_mul(&qa, ly);
        double s_ay = vcross(vadd(d, ra), ay), s_by = vcross(rb, ay);
        double Cc = vdot(d, ay);
        double impulse = (kk != 0.0) ? -Cc / kk : 0.0;
        v2 Pv = vscale(impulse, ay);
        double La = impulse * s_ay, Lb = impulse * s_by;
        ca = vsub(ca, vscale(ma, Pv));
        aa -= ia * La;
        cb = vadd(cb, vscale(mb, Pv));
        ab += ib * Lb;
        ok = fabs(Cc) <= LINEAR_SLOP;
    }
    else if (type == J_PULLEY) {                                                     /* pulley.py:258-318 */
        xform qa = rot_make(aa), qb = rot_make(ab);
        v2 ra = rot_mul(&qa, vsub(V(par[0], par[1]), lca)), rb = rot_mul(&qb, vsub(V(par[2], par[3]), lcb));
        v2 ua = vsub(vadd(ca, ra), V(par[4], par[5])), ub = vsub(vadd(cb, rb), V(par[6], par[7]));
        double length_a = sqrt(sq(ua.x) + sq(ua.y)), length_b = sqrt(sq(ub.x) + sq(ub.y));
        ua = (length_a > 10.0 * LINEAR_SLOP) ? V(ua.x / length_a, ua.y / length_a) : V(0.0, 0.0);
        ub = (length_b > 10.0 * LINEAR_SLOP) ? V(ub.x / length_b, ub.y / length_b) : V(0.0, 0.0);
        double rua = vcross(ra, ua), rub = vcross(rb, ub);
        double ema = ma + ia * sq(rua), emb = mb + ib * sq(rub);
        double mass = ema + sq(par[8]) * emb;
        if (mass > 0.0) mass = 1.0 / mass;
        double Cc = par[9] - length_a - par[8] * length_b;
        double linear_error = fabs(Cc);
        double impulse = -mass * Cc;
        v2 Pa = vscale(-impulse, ua), Pb = vscale(-par[8] * impulse, ub);
        ca = vadd(ca, vscale(ema, Pa));
        aa += ia * vcross(ra, Pa);
        cb = vadd(cb, vscale(emb, Pb));
        ab += ib * vcross(rb, Pb);
        ok = linear_error < LINEAR_SLOP;
    }
    P[ia_].c = ca; P[ia_].a = aa;
    P[ib_].c = cb; P[ib_].a = ab;
    return ok;
}

static void island_solve(orc_world* w, const int* bodies, int nbodies, const int* contacts, int ncontacts, const int* joints, int njoints,
                         double dt, double dt_ratio, int vel_iters, int pos_iters, int* pos_iters_used)
{
    pos_t* P = (pos_t*)malloc(sizeof(pos_t) * (size_t)(nbodies + 1));
    vel_t* Vv = (vel_t*)malloc(sizeof(vel_t) * (size_t)(nbodies + 1));
    cc_t* C = (cc_t*)malloc(sizeof(cc_t) * (size_t)(ncontacts + 1));
    v2 gravity = V(w->w_gravity[0], w->w_gravity[1]);
    int warm = (w->w_flags[0] & W_WARM_START) != 0;

    /* island.py:71-95 */
    for (int i = 0; i < nbodies; ++i) {
        int b = bodies[i];
        double* pos = w->b_pos + 4 * (size_t)b; double* vel = w->b_vel + 4 * (size_t)b;
        v2 c = V(pos[0], pos[1]); double a = pos[2];
        v2 v = V(vel[0], vel[1]); double wv = vel[2];
        w->b_pos0[4 * b + 0] = c.x; w->b_pos0[4 * b + 1] = c.y; w->b_pos0[4 * b + 2] = a;
        if (body_type(w, b) == T_DYNAMIC) {
            const double* mass = w->b_mass + 4 * (size_t)b; const double* damp = w->b_damp + 4 * (size_t)b;
            v2 force = V(w->b_force[2 * b], w->b_force[2 * b + 1]);
            v2 acc = vadd(vscale(damp[2], gravity), vscale(mass[0], force));
            v = vadd(v, vscale(dt, acc));
            wv = wv + dt * mass[1] * vel[3];
            v = vscale(pyclamp(1.0 - dt * damp[0], 0.0, 1.0), v);
            wv = wv * pyclamp(1.0 - dt * damp[1], 0.0, 1.0);
        }
        P[i].c = c; P[i].a = a; Vv[i].v = v; Vv[i].w = wv;
    }

    solver_ctor(w, C, contacts, ncontacts, warm, dt_ratio);
    solver_init_velocity(C, ncontacts, P, Vv);

    /* warm_start: contact.py:498-525 */
    if (warm) {
        for (int k = 0; k < ncontacts; ++k) {
            cc_t* cc = &C[k];
            v2 normal = cc->normal, tangent = vcross_s(normal, 1.0);
            v2 va = Vv[cc->ia].v, vb = Vv[cc->ib].v; double wa = Vv[cc->ia].w, wb = Vv[cc->ib].w;
            for (int j = 0; j < cc->npoints; ++j) {
                vcp_t* p = &cc->p[j];
                v2 pp = vadd(vscale(p->normal_impulse, normal), vscale(p->tangent_impulse, tangent));
                wa = wa - cc->iia * vcross(p->ra, pp);
                va = vsub(va, vscale(cc->ima, pp));
                wb = wb + cc->iib * vcross(p->rb, pp);
                vb = vadd(vb, vscale(cc->imb, pp));
            }
            Vv[cc->ia].v = va; Vv[cc->ia].w = wa;
            Vv[cc->ib].v = vb; Vv[cc->ib].w = wb;
        }
    }

    /* island.py:107-113 */
    for (int k = 0; k < njoints; ++k) joint_init(w, joints[k], dt, dt_ratio, warm, P, Vv);
    for (int it = 0; it < vel_iters; ++it) {
        for (int k = 0; k < njoints; ++k) joint_solve_velocity(w, joints[k], dt, 1.0 / dt, Vv);
        solver_velocity_sweep(C, ncontacts, Vv);
    }

    /* store_impulses: contact.py:801-806 (zip over vc.points, which may be shorter) */
    for (int k = 0; k < ncontacts; ++k) {
        cc_t* cc = &C[k];
        double* imp = w->c_imp + 4 * (size_t)cc->contact;
        int n = cc->npoints < cc->n_local_points ? cc->npoints : cc->n_local_points;
        for (int j = 0; j < n; ++j) { imp[2 * j] = cc->p[j].normal_impulse; imp[2 * j + 1] = cc->p[j].tangent_impulse; }
    }

    /* integrate positions: island.py:119-138 */
    for (int i = 0; i < nbodies; ++i) {
        v2 c = P[i].c; double a = P[i].a; v2 v = Vv[i].v; double wv = Vv[i].w;
        v2 translation = vscale(dt, v);
        if (vdot(translation, translation) > MAX_TRANS_SQR) {
            double ratio = MAX_TRANS / sqrt(sq(translation.x) + sq(translation.y));
            v = vscale(ratio, v);
        }
        double rotation = dt * wv;
        if (sq(rotation) > MAX_ROT_SQR) {
            double ratio = MAX_ROT / fabs(rotation);
            wv = wv * ratio;
        }
        c = vadd(c, vscale(dt, v));
        a = a + dt * wv;
        P[i].c = c; P[i].a = a; Vv[i].v = v; Vv[i].w = wv;
    }

    /* position iterations: island.py:141-153, contact.py:809-864 */
    int position_solved = 0, used = 0;
    for (int it = 0; it < pos_iters; ++it) {
        double min_separation = 0.0;
        ++used;
        for (int k = 0; k < ncontacts; ++k) {
            cc_t* cc = &C[k];
            double ima = cc->ima, imb = cc->imb, iia = cc->iia, iib = cc->iib;
            v2 ca = P[cc->ia].c, cb = P[cc->ib].c; double aa = P[cc->ia].a, ab = P[cc->ib].a;
            for (int j = 0; j < cc->n_local_points; ++j) {
                xform xa = xf_from_sweep(ca, aa, cc->lca), xb = xf_from_sweep(cb, ab, cc->lcb);
                v2 normal, point; double separation;
                solver_manifold(cc, &xa, &xb, j, &normal, &point, &separation);
                v2 ra = vsub(point, ca), rb = vsub(point, cb);
                min_separation = PYMIN(min_separation, separation);
                double Cc = pyclamp(BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LIN_CORR, 0.0);
                double rn_a = sq(vcross(ra, normal)), rn_b = sq(vcross(rb, normal));
                double K = ima + imb + (iia * rn_a) + (iib * rn_b);
                double impulse = (K > 0.0) ? -Cc / K : 0.0;
                v2 Pv = vscale(impulse, normal);
                ca = vsub(ca, vscale(ima, Pv));
                aa = aa - iia * vcross(ra, Pv);
                cb = vadd(cb, vscale(imb, Pv));
                ab = ab + iib * vcross(rb, Pv);
            }
            P[cc->ia].c = ca; P[cc->ia].a = aa;
            P[cc->ib].c = cb; P[cc->ib].a = ab;
        }
        /* island.py:145-153 */
        int joints_okay = 1;
        for (int k = 0; k < njoints; ++k) { int okay = joint_solve_position(w, joints[k], P); joints_okay = joints_okay && okay; }
        if (min_separation >= -3.0 * LINEAR_SLOP && joints_okay) { position_solved = 1; break; }
    }
    *pos_iters_used = used;

    /* write back + Body._synchronize_transform: island.py:155-159, body.py:906-908 */
    for (int i = 0; i < nbodies; ++i) {
        int b = bodies[i];
        w->b_pos[4 * b + 0] = P[i].c.x; w->b_pos[4 * b + 1] = P[i].c.y; w->b_pos[4 * b + 2] = P[i].a;
        w->b_vel[4 * b + 0] = Vv[i].v.x; w->b_vel[4 * b + 1] = Vv[i].v.y; w->b_vel[4 * b + 2] = Vv[i].w;
        v2 lc = V(w->b_mass[4 * b + 2], w->b_mass[4 * b + 3]);
        xform t = xf_from_sweep(P[i].c, P[i].a, lc);
        w->b_xf[4 * b + 0] = t.p.x; w->b_xf[4 * b + 1] = t.p.y; w->b_xf[4 * b + 2] = t.c1.x; w->b_xf[4 * b + 3] = t.c1.y;
    }

    /* sleep: island.py:163-179 -- the reference's rule as written:
     * `body` is the LAST island body; only the last non-static body `b` is put to sleep. */
    if (w->w_flags[0] & W_ALLOW_SLEEP) {
        double min_sleep_time = MAX_FLOAT;
        int last_body = bodies[nbodies - 1];
        int last_allow = (w->b_flags[last_body] & BODY_ALLOW_SLEEP) != 0;
        int b_last_nonstatic = -1;
        for (int i = 0; i < nbodies; ++i) {
            int b = bodies[i];
            if (body_type(w, b) == T_STATIC) continue;
            b_last_nonstatic = b;
            double wv = w->b_vel[4 * b + 2];
            v2 v = V(w->b_vel[4 * b + 0], w->b_vel[4 * b + 1]);
            if (!last_allow || sq(wv) > ANG_SLEEP_TOL_SQR || vlen2(v) > LIN_SLEEP_TOL_SQR) {
                w->b_pos[4 * b + 3] = 0.0;
                min_sleep_time = 0.0;
            } else {
                w->b_pos[4 * b + 3] = w->b_pos[4 * b + 3] + dt;
                min_sleep_time = PYMIN(min_sleep_time, w->b_pos[4 * b + 3]);
            }
        }
        if (min_sleep_time >= TIME_TO_SLEEP && position_solved && b_last_nonstatic >= 0)
            body_sleep(w, b_last_nonstatic);
    }
    free(P); free(Vv); free(C);
}

/* ---- World._solve: world.py:475-592 ---- */
static void world_solve(orc_world* w, double dt, double dt_ratio, int vel_iters, int pos_iters)
{
    int nb = w->w_count[0], nf = w->w_count[1], nc = w->w_count[2];
    for (int b = 0; b < nb; ++b) { w->b_flags[b] &= ~BODY_ISLAND; w->b_island[b] = -1; }
    for (int c = 0; c < nc; ++c) w->c_flags[c] &= ~C_ISLAND;
    int nj = w->w_joint[0], lj = 0;
    if (nj > w->nj_cap) { w->w_status[0] |= ST_JOINT_OVERFLOW; nj = w->nj_cap; }
    for (int j = 0; j < nj; ++j) w->j_def[4 * j + 3] &= ~JF_ISLAND;                /* world.py:491-492 */

    /* Body._contacts == world contact list filtered by body, same order (contact_manager.py:157-158,170-172) */
    int* adj_start = (int*)calloc((size_t)nb + 2, sizeof(int));
    int* adj = (int*)malloc(sizeof(int) * (size_t)(2 * nc + 1));
    for (int c = 0; c < nc; ++c) {
        adj_start[w->f_body[w->c_fix[2 * c]] + 1]++;
        adj_start[w->f_body[w->c_fix[2 * c + 1]] + 1]++;
    }
    for (int b = 0; b < nb; ++b) adj_start[b + 1] += adj_start[b];
    int* fill = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    memcpy(fill, adj_start, sizeof(int) * (size_t)nb);
    for (int c = 0; c < nc; ++c) {
        int ba = w->f_body[w->c_fix[2 * c]], bb = w->f_body[w->c_fix[2 * c + 1]];
        adj[fill[ba]++] = c;
        adj[fill[bb]++] = c;
    }

    int* stack = (int*)malloc(sizeof(int) * (size_t)(nb + 1));
    int n_islands = 0, lb = 0, lc = 0, pos_max = 0;
    for (int seed = 0; seed < nb; ++seed) {
        uint32_t fl = w->b_flags[seed];
        if (fl & BODY_ISLAND) continue;
        if (!(fl & BODY_AWAKE) || !(fl & BODY_ACTIVE)) continue;
        if ((fl & BODY_TYPE_MASK) == T_STATIC) continue;
        int body_start = lb, contact_start = lc, joint_start = lj;
        int sp = 0;
        stack[sp++] = seed;
        w->b_flags[seed] |= BODY_ISLAND;
        int overflow = 0;
        while (sp > 0) {
            int b = stack[--sp];
            if (lb >= w->nl_cap) { overflow = 1; break; }
            w->b_island[b] = lb - body_start;
            w->l_body[lb++] = b;
            body_wake(w, b);
            if (body_type(w, b) == T_STATIC) continue;
            /* the comprehension filters first, then the loop appends (world.py:528-546) */
            int first = lc;
            for (int k = adj_start[b]; k < adj_start[b + 1]; ++k) {
                int c = adj[k];
                uint32_t cf = w->c_flags[c];
                if ((cf & C_ISLAND) || !(cf & C_ENABLED) || !(cf & C_TOUCHING) || (cf & C_SENSOR)) continue;   /* world.py:528-533 */
                w->l_contact[lc++] = c;
            }
            for (int k = first; k < lc; ++k) {
                int c = w->l_contact[k];
                w->c_flags[c] |= C_ISLAND;
                int ba = w->f_body[w->c_fix[2 * c]], bb = w->f_body[w->c_fix[2 * c + 1]];
                int other = (b == ba) ? bb : ba;
                if (w->b_flags[other] & BODY_ISLAND) continue;
                stack[sp++] = other;
                w->b_flags[other] |= BODY_ISLAND;
            }
            /* world.py:548-572: Body.joints in append order */
            for (int k = w->b_jadj[2 * b], ke = k + w->b_jadj[2 * b + 1]; k < ke; ++k) {
                int j = w->j_adj[k];
                int32_t* def = w->j_def + 4 * (size_t)j;
                if (def[3] & JF_ISLAND) continue;
                int other = (def[1] == b) ? def[2] : def[1];
                if (other < 0) { w->l_joint[lj++] = j; def[3] |= JF_ISLAND; continue; }   /* a joint on one body (MouseJoint): world.py:556-560 */
                if (!(w->b_flags[other] & BODY_ACTIVE)) continue;
                w->l_joint[lj++] = j;
                def[3] |= JF_ISLAND;
                if (w->b_flags[other] & BODY_ISLAND) continue;
                stack[sp++] = other;
                w->b_flags[other] |= BODY_ISLAND;
            }
        }
        if (overflow) { w->w_status[0] |= ST_ISLAND_OVERFLOW; break; }
        w->l_start[2 * n_islands] = body_start; w->l_start[2 * n_islands + 1] = contact_start;
        w->l_jstart[n_islands] = joint_start;
        int used = 0;
        island_solve(w, w->l_body + body_start, lb - body_start, w->l_contact + contact_start, lc - contact_start,
                     w->l_joint + joint_start, lj - joint_start, dt, dt_ratio, vel_iters, pos_iters, &used);
        if (used > pos_max) pos_max = used;
        for (int k = body_start; k < lb; ++k)
            if (body_type(w, w->l_body[k]) == T_STATIC) w->b_flags[w->l_body[k]] &= ~BODY_ISLAND;
        ++n_islands;
    }
    w->w_island[0] = n_islands; w->w_island[1] = lb; w->w_island[2] = lc; w->w_island[3] = pos_max;
    w->w_joint[1] = lj;

    /* world.py:582-589 + body.py:898-904 */
    for (int f = 0; f < nf; ++f) {
        int b = w->f_body[f];
        if (!(w->b_flags[b] & BODY_ISLAND) || body_type(w, b) == T_STATIC) continue;
        v2 c0 = V(w->b_pos0[4 * b], w->b_pos0[4 * b + 1]); double a0 = w->b_pos0[4 * b + 2];
        v2 lcen = V(w->b_mass[4 * b + 2], w->b_mass[4 * b + 3]);
        xform xf1 = xf_from_sweep(c0, a0, lcen);
        xform xf2 = xf_row(w->b_xf + 4 * (size_t)b);
        fixture_synchronize(w, f, &xf1, &xf2);
    }
    free(adj_start); free(adj); free(fill); free(stack);
    cm_find_new_contacts(w);
}


/* =====================================================================================================
 * Continuous collision (World(continuous_physics=True), the reference's default): world.py:594-821,
 * island.py:181-243, contact.py:866-933, toi.py:46-391, distance.py:39-561, common.py:871-923,
 * body.py:930-937.  No sensors, no sub_stepping.
 * ===================================================================================================== */
#define W_CONTINUOUS       0x20u
#define BODY_BULLET        0x40u
#define MAX_SUB_STEPS      8         /* settings.py:73 */
#define MAX_TOI_CONTACTS   32        /* settings.py:79 */
#define TOI_MAX_ITERS      20        /* settings.py:82 */
#define TOI_BAUMGARTE      0.75      /* settings.py:110 */
#define DISTANCE_MAX_ITERS 20        /* settings.py:129 */

/* Sweep: common.py:871-923.  alpha0 lives in the spare column of b_pos0. */
typedef struct { v2 lc, c0, c; double a0, a, alpha0; } sweep_t;
static sweep_t sweep_load(const orc_world* w, int b)
{
    sweep_t s;
    s.lc = V(w->b_mass[4 * b + 2], w->b_mass[4 * b + 3]);
    s.c0 = V(w->b_pos0[4 * b], w->b_pos0[4 * b + 1]); s.a0 = w->b_pos0[4 * b + 2]; s.alpha0 = w->b_pos0[4 * b + 3];
    s.c = V(w->b_pos[4 * b], w->b_pos[4 * b + 1]); s.a = w->b_pos[4 * b + 2];
    return s;
}
static void sweep_store(orc_world* w, int b, const sweep_t* s)
{
    w->b_pos0[4 * b] = s->c0.x; w->b_pos0[4 * b + 1] = s->c0.y; w->b_pos0[4 * b + 2] = s->a0; w->b_pos0[4 * b + 3] = s->alpha0;
    w->b_pos[4 * b] = s->c.x; w->b_pos[4 * b + 1] = s->c.y; w->b_pos[4 * b + 2] = s->a;
}
/* Sweep.get_transform: common.py:893-904 */
static xform sweep_xf(const sweep_t* s, double beta)
{
    v2 p = vadd(vscale(1.0 - beta, s->c0), vscale(beta, s->c));
    double angle = (1.0 - beta) * s->a0 + beta * s->a;
    return xf_from_sweep(p, angle, s->lc);
}
/* Sweep.advance: common.py:906-915 */
static void sweep_advance(sweep_t* s, double alpha)
{
    double beta = (alpha - s->alpha0) / (1.0 - s->alpha0);
    s->c0 = vadd(vscale(1.0 - beta, s->c0), vscale(beta, s->c));
    s->a0 = (1.0 - beta) * s->a0 + beta * s->a;
    s->alpha0 = alpha;
}
/* Sweep.normalize: common.py:917-923 */
static void sweep_normalize(sweep_t* s)
{
    double d = 2.0 * PI_ * floor(s->a0 / (2.0 * PI_));
    s->a0 -= d;
    s->a -= d;
}
/* Body._synchronize_transform: body.py:906-908 */
static void body_sync_transform(orc_world* w, int b)
{
    v2 lc = V(w->b_mass[4 * b + 2], w->b_mass[4 * b + 3]);
    xform t = xf_from_sweep(V(w->b_pos[4 * b], w->b_pos[4 * b + 1]), w->b_pos[4 * b + 2], lc);
    w->b_xf[4 * b + 0] = t.p.x; w->b_xf[4 * b + 1] = t.p.y; w->b_xf[4 * b + 2] = t.c1.x; w->b_xf[4 * b + 3] = t.c1.y;
}
/* Body._advance: body.py:930-937 (the rotation is rebuilt from the angle, the position from the centre: the same
 * arithmetic as _synchronize_transform) */
static void body_advance(orc_world* w, int b, double alpha)
{
    sweep_t s = sweep_load(w, b);
    sweep_advance(&s, alpha);
    s.c = s.c0; s.a = s.a0;
    sweep_store(w, b, &s);
    body_sync_transform(w, b);
}

/* DistanceProxy: distance.py:354-409 */
typedef struct { int count; double radius; const double* v; } dproxy;
static dproxy dproxy_get(const orc_world* w, int shape_row)
{
    shape_t s = shape_get(w, shape_row);
    dproxy p; p.v = s.d; p.radius = s.radius;
    p.count = (s.type == S_CIRCLE) ? 1 : (s.type == S_POLYGON) ? s.count : 2;
    return p;
}
static inline v2 dvert(const dproxy* p, int i) { return V(p->v[2 * i], p->v[2 * i + 1]); }
/* get_support: distance.py:393-396 -- dots.index(max(dots)): the first index holding the maximum */
static int dproxy_support(const dproxy* p, v2 d)
{
    int best = 0; double bv = vdot(dvert(p, 0), d);
    for (int i = 1; i < p->count; ++i) { double v = vdot(dvert(p, i), d); if (v > bv) { bv = v; best = i; } }
    return best;
}

typedef struct { double metric; int count; int ia[3], ib[3]; } simplex_cache;
typedef struct { v2 wa, wb, w; double a; int ia, ib; } svert;
static inline double vdist(v2 a, v2 b) { v2 c = vsub(a, b); return sqrt(sq(c.x) + sq(c.y)); }   /* common.py:947-949 */

/* Simplex: distance.py:76-352.  The reference holds its three vertices as attributes that solve2 / solve3 REBIND to copies
 * (self._v1 = copy(self._v2)), while shape_distance keeps writing new support points through the list it took BEFORE the
 * loop (vertices = simplex.vertices, distance.py:458): after a rebind the list entry and the attribute are different
 * objects.  Restated with explicit object identities: `pool` are the objects, cur[k] the attribute _v(k+1), lst[k] the
 * entry of the list taken before the loop. */
typedef struct { svert pool[3 + 2 * DISTANCE_MAX_ITERS + 4]; int npool; int cur[3], lst[3]; int count; } simplex_t;
static int simplex_copy(simplex_t* s, int src) { s->pool[s->npool] = s->pool[src]; return s->npool++; }
static double simplex_metric(const simplex_t* s)                                   /* distance.py:189-203 */
{
    const svert *v1 = &s->pool[s->cur[0]], *v2_ = &s->pool[s->cur[1]], *v3 = &s->pool[s->cur[2]];
    if (s->count == 1) return 0.0;
    if (s->count == 2) return vdist(v1->w, v2_->w);
    return vcross(vsub(v2_->w, v1->w), vsub(v3->w, v1->w));
}
static void simplex_solve2(simplex_t* s)                                           /* distance.py:205-253 */
{
    svert *v1 = &s->pool[s->cur[0]], *v2_ = &s->pool[s->cur[1]];
    v2 w1 = v1->w, w2 = v2_->w, e12 = vsub(w2, w1);
    double d12_2 = -vdot(w1, e12);
    if (d12_2 <= 0.0) { v1->a = 1.0; s->count = 1; return; }
    double d12_1 = vdot(w2, e12);
    if (d12_1 <= 0.0) { v2_->a = 1.0; s->count = 1; s->cur[0] = simplex_copy(s, s->cur[1]); return; }
    double inv_d12 = 1.0 / (d12_1 + d12_2);
    v1->a = d12_1 * inv_d12; v2_->a = d12_2 * inv_d12; s->count = 2;
}
static void simplex_solve3(simplex_t* s)                                           /* distance.py:255-352 */
{
    svert *v1 = &s->pool[s->cur[0]], *v2_ = &s->pool[s->cur[1]], *v3 = &s->pool[s->cur[2]];
    v2 w1 = v1->w, w2 = v2_->w, w3 = v3->w;
    v2 e12 = vsub(w2, w1);
    double w1e12 = vdot(w1, e12), w2e12 = vdot(w2, e12), d12_1 = w2e12, d12_2 = -w1e12;
    v2 e13 = vsub(w3, w1);
    double w1e13 = vdot(w1, e13), w3e13 = vdot(w3, e13), d13_1 = w3e13, d13_2 = -w1e13;
    v2 e23 = vsub(w3, w2);
    double w2e23 = vdot(w2, e23), w3e23 = vdot(w3, e23), d23_1 = w3e23, d23_2 = -w2e23;
    double n123 = vcross(e12, e13);
    double d123_1 = n123 * vcross(w2, w3), d123_2 = n123 * vcross(w3, w1), d123_3 = n123 * vcross(w1, w2);
    if (d12_2 <= 0.0 && d13_2 <= 0.0) { v1->a = 1.0; s->count = 1; return; }
    if (d12_1 > 0.0 && d12_2 > 0.0 && d123_3 <= 0.0) {
        double inv = 1.0 / (d12_1 + d12_2);
        v1->a = d12_1 * inv; v2_->a = d12_2 * inv; s->count = 2; return;
    }
    if (d13_1 > 0.0 && d13_2 > 0.0 && d123_2 <= 0.0) {
        double inv = 1.0 / (d13_1 + d13_2);
        v1->a = d13_1 * inv; v3->a = d13_2 * inv; s->count = 2; s->cur[1] = simplex_copy(s, s->cur[2]); return;
    }
    if (d12_1 <= 0.0 && d23_2 <= 0.0) { v2_->a = 1.0; s->count = 1; s->cur[0] = simplex_copy(s, s->cur[1]); return; }
    if (d13_1 <= 0.0 && d23_1 <= 0.0) { v3->a = 1.0; s->count = 1; s->cur[0] = simplex_copy(s, s->cur[2]); return; }
    if (d23_1 > 0.0 && d23_2 > 0.0 && d123_1 <= 0.0) {
        double inv = 1.0 / (d23_1 + d23_2);
        v2_->a = d23_1 * inv; v3->a = d23_2 * inv; s->count = 2; s->cur[0] = simplex_copy(s, s->cur[2]); return;
    }
    double inv = 1.0 / (d123_1 + d123_2 + d123_3);
    v1->a = d123_1 * inv; v2_->a = d123_2 * inv; v3->a = d123_3 * inv; s->count = 3;
}

/* shape_distance: distance.py:419-561.  Returns the distance; the cache is read and written. */
static double shape_distance(const dproxy* pa, const dproxy* pb, const xform* xa, const xform* xb, int use_radii,
                             simplex_cache* cache)
{
    simplex_t S; memset(&S, 0, sizeof(S));
    S.npool = 3;
    for (int k = 0; k < 3; ++k) { S.cur[k] = k; S.lst[k] = k; }
    /* read_cache: distance.py:97-135 */
    S.count = cache->count;
    for (int i = 0; i < S.count; ++i) {
        svert* v = &S.pool[S.cur[i]];
        v->ia = cache->ia[i]; v->ib = cache->ib[i];
        v->wa = xf_mul(xa, dvert(pa, v->ia));
        v->wb = xf_mul(xb, dvert(pb, v->ib));
        v->w = vsub(v->wb, v->wa);
        v->a = 0.0;
    }
    if (S.count > 1) {
        double metric1 = cache->metric, metric2 = simplex_metric(&S);
        if (metric2 < (0.5 * metric1) || (2.0 * metric1) < metric2 || metric2 < EPS) S.count = 0;
    }
    if (S.count == 0) {
        svert* v = &S.pool[S.cur[0]];
        v->ia = 0; v->ib = 0;
        v->wa = xf_mul(xa, dvert(pa, 0));
        v->wb = xf_mul(xb, dvert(pb, 0));
        v->w = vsub(v->wb, v->wa);
        S.count = 1;
    }
    int iter = 0;
    while (iter < DISTANCE_MAX_ITERS) {
        int save_a[3], save_b[3], nsave = S.count;
        for (int i = 0; i < nsave; ++i) { save_a[i] = S.pool[S.cur[i]].ia; save_b[i] = S.pool[S.cur[i]].ib; }
        if (S.count == 2) simplex_solve2(&S);
        else if (S.count == 3) { simplex_solve3(&S); break; }
        /* search_direction: distance.py:152-169 */
        v2 d;
        if (S.count == 1) d = vneg(S.pool[S.cur[0]].w);
        else {
            v2 e12 = vsub(S.pool[S.cur[1]].w, S.pool[S.cur[0]].w);
            double sgn = vcross(e12, vneg(S.pool[S.cur[0]].w));
            d = (sgn > 0.0) ? scross(1.0, e12) : vcross_s(e12, 1.0);
        }
        if (vlen2(d) < EPS_SQR) break;
        svert* vertex = &S.pool[S.lst[S.count]];
        vertex->ia = dproxy_support(pa, rot_mul_t(xa, vneg(d)));
        vertex->wa = xf_mul(xa, dvert(pa, vertex->ia));
        vertex->ib = dproxy_support(pb, rot_mul_t(xb, d));
        vertex->wb = xf_mul(xb, dvert(pb, vertex->ib));
        vertex->w = vsub(vertex->wb, vertex->wa);
        ++iter;
        int dup = 0;
        for (int i = 0; i < nsave; ++i) if (save_a[i] == vertex->ia && save_b[i] == vertex->ib) dup = 1;
        if (dup) break;
        S.count += 1;
    }
    /* witness_points: distance.py:171-187 */
    const svert *v1 = &S.pool[S.cur[0]], *v2_ = &S.pool[S.cur[1]], *v3 = &S.pool[S.cur[2]];
    v2 point_a, point_b;
    if (S.count == 1) { point_a = v1->wa; point_b = v1->wb; }
    else if (S.count == 2) {
        point_a = vadd(vscale(v1->a, v1->wa), vscale(v2_->a, v2_->wa));
        point_b = vadd(vscale(v1->a, v1->wb), vscale(v2_->a, v2_->wb));
    } else {
        point_a = vadd(vadd(vscale(v1->a, v1->wa), vscale(v2_->a, v2_->wa)), vscale(v3->a, v3->wa));
        point_b = point_a;
    }
    double dist = vdist(point_a, point_b);
    /* write_cache: distance.py:137-142 */
    cache->metric = simplex_metric(&S);
    cache->count = S.count;
    for (int i = 0; i < S.count; ++i) { cache->ia[i] = S.pool[S.cur[i]].ia; cache->ib[i] = S.pool[S.cur[i]].ib; }
    if (use_radii) {
        double ra = pa->radius, rb = pb->radius;
        if (dist > (ra + rb) && dist > EPS) dist -= ra + rb;
        else dist = 0.0;
    }
    return dist;
}

/* test_overlap: distance.py:411-417 */
static int sensor_test_overlap(const orc_world* w, int fa, int fb, const xform* xa, const xform* xb)
{
    dproxy pa = dproxy_get(w, w->f_shape[fa]), pb = dproxy_get(w, w->f_shape[fb]);
    simplex_cache cache; memset(&cache, 0, sizeof(cache));
    return shape_distance(&pa, &pb, xa, xb, 1, &cache) < (10.0 * EPS);
}

/* SeparationFunction: toi.py:46-221 */
enum { SEP_POINTS, SEP_FACE_A, SEP_FACE_B };
typedef struct { const dproxy *pa, *pb; const sweep_t *sa, *sb; int type; v2 axis, local_point; } sepfn;
static void sep_initialize(sepfn* f, const simplex_cache* cache, double t1)
{
    xform xa = sweep_xf(f->sa, t1), xb = sweep_xf(f->sb, t1);
    int ia0 = cache->ia[0], ib0 = cache->ib[0];
    if (cache->count == 1) {
        /* toi.py:64-75: `s = self.axis.normalized` builds a new vector; the axis itself stays UN-normalised */
        f->type = SEP_POINTS;
        v2 point_a = xf_mul(&xa, dvert(f->pa, ia0)), point_b = xf_mul(&xb, dvert(f->pb, ib0));
        f->axis = vsub(point_b, point_a);
        return;
    }
    int ia1 = cache->ia[1], ib1 = cache->ib[1];
    if (ia0 == ia1) {
        f->type = SEP_FACE_B;
        v2 b1 = dvert(f->pb, ib0), b2 = dvert(f->pb, ib1);
        f->axis = vnorm(vcross_s(vsub(b2, b1), 1.0));
        v2 normal = rot_mul(&xb, f->axis);
        f->local_point = vscale(0.5, vadd(b1, b2));
        v2 point_b = xf_mul(&xb, f->local_point);
        v2 point_a = xf_mul(&xa, dvert(f->pa, ia0));
        double s = vdot(vsub(point_a, point_b), normal);
        if (s < 0.0) f->axis = vneg(f->axis);
    } else {
        f->type = SEP_FACE_A;
        v2 a1 = dvert(f->pa, ia0), a2 = dvert(f->pa, ia1);
        f->axis = vnorm(vcross_s(vsub(a2, a1), 1.0));
        v2 normal = rot_mul(&xa, f->axis);
        f->local_point = vscale(0.5, vadd(a1, a2));
        v2 point_a = xf_mul(&xa, f->local_point);
        v2 point_b = xf_mul(&xb, dvert(f->pb, ib0));
        double s = vdot(vsub(point_b, point_a), normal);
        if (s < 0.0) f->axis = vneg(f->axis);
    }
}
static double sep_find_min(const sepfn* f, double t, int* index_a, int* index_b)     /* toi.py:126-183 */
{
    xform xa = sweep_xf(f->sa, t), xb = sweep_xf(f->sb, t);
    if (f->type == SEP_POINTS) {
        v2 axis_a = rot_mul_t(&xa, f->axis), axis_b = rot_mul_t(&xb, vneg(f->axis));
        *index_a = dproxy_support(f->pa, axis_a);
        *index_b = dproxy_support(f->pb, axis_b);
        v2 point_a = xf_mul(&xa, dvert(f->pa, *index_a)), point_b = xf_mul(&xb, dvert(f->pb, *index_b));
        return vdot(vsub(point_b, point_a), f->axis);
    } else if (f->type == SEP_FACE_A) {
        v2 normal = rot_mul(&xa, f->axis);
        v2 point_a = xf_mul(&xa, f->local_point);
        v2 axis_b = rot_mul_t(&xb, vneg(normal));
        *index_a = -1;
        *index_b = dproxy_support(f->pb, axis_b);
        v2 point_b = xf_mul(&xb, dvert(f->pb, *index_b));
        return vdot(vsub(point_b, point_a), normal);
    } else {
        v2 normal = rot_mul(&xb, f->axis);
        v2 point_b = xf_mul(&xb, f->local_point);
        v2 axis_a = rot_mul_t(&xa, vneg(normal));
        *index_b = -1;
        *index_a = dproxy_support(f->pa, axis_a);
        v2 point_a = xf_mul(&xa, dvert(f->pa, *index_a));
        return vdot(vsub(point_a, point_b), normal);
    }
}
static double sep_evaluate(const sepfn* f, int index_a, int index_b, double t)       /* toi.py:185-221 */
{
    xform xa = sweep_xf(f->sa, t), xb = sweep_xf(f->sb, t);
    if (f->type == SEP_POINTS) {
        v2 point_a = xf_mul(&xa, dvert(f->pa, index_a)), point_b = xf_mul(&xb, dvert(f->pb, index_b));
        return vdot(vsub(point_b, point_a), f->axis);
    } else if (f->type == SEP_FACE_A) {
        v2 normal = rot_mul(&xa, f->axis);
        v2 point_a = xf_mul(&xa, f->local_point);
        v2 point_b = xf_mul(&xb, dvert(f->pb, index_b));
        return vdot(vsub(point_b, point_a), normal);
    } else {
        v2 normal = rot_mul(&xb, f->axis);
        v2 point_b = xf_mul(&xb, f->local_point);
        v2 point_a = xf_mul(&xa, dvert(f->pa, index_a));
        return vdot(vsub(point_a, point_b), normal);
    }
}

/* time_of_impact: toi.py:223-391 */
enum { TOI_UNKNOWN, TOI_FAILED, TOI_OVERLAPPED, TOI_TOUCHING, TOI_SEPARATED };
static int time_of_impact(const dproxy* pa, const dproxy* pb, sweep_t sweep_a, sweep_t sweep_b, double t_max, double* out_t)
{
    int state = TOI_UNKNOWN;
    double output_t = t_max;
    sweep_normalize(&sweep_a);
    sweep_normalize(&sweep_b);
    double total_radius = pa->radius + pb->radius;
    double target = PYMAX(LINEAR_SLOP, total_radius - 3.0 * LINEAR_SLOP);
    double tolerance = 0.25 * LINEAR_SLOP;
    double t1 = 0.0;
    int iter = 0;
    simplex_cache cache; memset(&cache, 0, sizeof(cache));
    for (;;) {
        xform xa = sweep_xf(&sweep_a, t1), xb = sweep_xf(&sweep_b, t1);
        double out_distance = shape_distance(pa, pb, &xa, &xb, 0, &cache);
        if (out_distance <= 0.0) { state = TOI_OVERLAPPED; output_t = 0.0; break; }
        if (out_distance < target + tolerance) { state = TOI_TOUCHING; output_t = t1; break; }
        sepfn fcn; memset(&fcn, 0, sizeof(fcn));
        fcn.pa = pa; fcn.pb = pb; fcn.sa = &sweep_a; fcn.sb = &sweep_b;
        sep_initialize(&fcn, &cache, t1);
        int done = 0;
        double t2 = t_max;
        for (;;) {
            int index_a, index_b;
            double s2 = sep_find_min(&fcn, t2, &index_a, &index_b);
            if (s2 > (target + tolerance)) { state = TOI_SEPARATED; output_t = t_max; done = 1; break; }
            if (s2 > (target - tolerance)) { t1 = t2; break; }
            double s1 = sep_evaluate(&fcn, index_a, index_b, t1);
            if (s1 < (target - tolerance)) { state = TOI_FAILED; output_t = t1; done = 1; break; }
            if (s1 <= (target + tolerance)) { state = TOI_TOUCHING; output_t = t1; done = 1; break; }
            int root_iter_count = 0;
            double a1 = t1, a2 = t2;
            for (;;) {
                double t;
                if (root_iter_count & 1) t = a1 + (target - s1) * (a2 - a1) / (s2 - s1);
                else t = 0.5 * (a1 + a2);
                double s = sep_evaluate(&fcn, index_a, index_b, t);
                if (fabs(s - target) < tolerance) { t2 = t; break; }
                if (s > target) { a1 = t; s1 = s; } else { a2 = t; s2 = s; }
                ++root_iter_count;
                if (root_iter_count == 50) break;
            }
        }
        ++iter;
        if (done) break;
        if (iter == TOI_MAX_ITERS) { state = TOI_FAILED; output_t = t1; break; }
    }
    *out_t = output_t;
    return state;
}

/* solve_toi_position_constraints: contact.py:866-933 */
static int solver_toi_position_sweep(const cc_t* C, int ncontacts, pos_t* P, int toi_a, int toi_b)
{
    double min_separation = 0.0;
    for (int k = 0; k < ncontacts; ++k) {
        const cc_t* cc = &C[k];
        v2 ca = P[cc->ia].c, cb = P[cc->ib].c; double aa = P[cc->ia].a, ab = P[cc->ib].a;
        double ima = 0.0, iia = 0.0, imb = 0.0, iib = 0.0;
        if (cc->ia == toi_a || cc->ia == toi_b) { ima = cc->ima; iia = cc->iia; }
        if (cc->ib == toi_a || cc->ib == toi_b) { imb = cc->imb; iib = cc->iib; }
        for (int j = 0; j < cc->n_local_points; ++j) {
            xform xa = xf_from_sweep(ca, aa, cc->lca), xb = xf_from_sweep(cb, ab, cc->lcb);
            v2 normal, point; double separation;
            solver_manifold(cc, &xa, &xb, j, &normal, &point, &separation);
            v2 ra = vsub(point, ca), rb = vsub(point, cb);
            min_separation = PYMIN(min_separation, separation);
            double Cc = pyclamp(TOI_BAUMGARTE * (separation + LINEAR_SLOP), -MAX_LIN_CORR, 0.0);
            double rn_a = sq(vcross(ra, normal)), rn_b = sq(vcross(rb, normal));
            double K = ima + imb + (iia * rn_a) + (iib * rn_b);
            double impulse = (K > 0.0) ? -Cc / K : 0.0;
            v2 Pv = vscale(impulse, normal);
            ca = vsub(ca, vscale(ima, Pv));
            aa = aa - iia * vcross(ra, Pv);
            cb = vadd(cb, vscale(imb, Pv));
            ab = ab + iib * vcross(rb, Pv);
        }
        P[cc->ia].c = ca; P[cc->ia].a = aa;
        P[cc->ib].c = cb; P[cc->ib].a = ab;
    }
    return min_separation >= -1.5 * LINEAR_SLOP;
}

/* Island.solve_toi: island.py:181-243 */
static void island_solve_toi(orc_world* w, const int* bodies, int nbodies, const int* contacts, int ncontacts,
                             double dt, int vel_iters, int toi_a, int toi_b)
{
    pos_t P[2 * MAX_TOI_CONTACTS + 1]; vel_t Vv[2 * MAX_TOI_CONTACTS + 1]; cc_t C[MAX_TOI_CONTACTS + 1];
    for (int i = 0; i < nbodies; ++i) {
        int b = bodies[i];
        P[i].c = V(w->b_pos[4 * b], w->b_pos[4 * b + 1]); P[i].a = w->b_pos[4 * b + 2];
        Vv[i].v = V(w->b_vel[4 * b], w->b_vel[4 * b + 1]); Vv[i].w = w->b_vel[4 * b + 2];
    }
    solver_ctor(w, C, contacts, ncontacts, 0, 1.0);
    for (int it = 0; it < 20; ++it)
        if (solver_toi_position_sweep(C, ncontacts, P, toi_a, toi_b)) break;
    /* leap of faith to the new safe state */
    {
        int ba = bodies[toi_a], bb = bodies[toi_b];
        w->b_pos0[4 * ba] = P[toi_a].c.x; w->b_pos0[4 * ba + 1] = P[toi_a].c.y; w->b_pos0[4 * ba + 2] = P[toi_a].a;
        w->b_pos0[4 * bb] = P[toi_b].c.x; w->b_pos0[4 * bb + 1] = P[toi_b].c.y; w->b_pos0[4 * bb + 2] = P[toi_b].a;
    }
    solver_init_velocity(C, ncontacts, P, Vv);
    for (int it = 0; it < vel_iters; ++it) solver_velocity_sweep(C, ncontacts, Vv);
    /* (impulses are not stored) integrate: island.py:213-241 */
    for (int i = 0; i < nbodies; ++i) {
        int b = bodies[i];
        v2 c = P[i].c; double a = P[i].a; v2 v = Vv[i].v; double wv = Vv[i].w;
        v2 translation = vscale(dt, v);
        if (vlen2(translation) > MAX_TRANS_SQR) {
            double ratio = MAX_TRANS / sqrt(sq(translation.x) + sq(translation.y));
            v = vscale(ratio, v);                       /* v *= ratio */
        }
        double rotation = dt * wv;
        if (sq(rotation) > MAX_ROT_SQR) {
            double ratio = MAX_ROT / fabs(rotation);
            wv = wv * ratio;
        }
        c = vadd(c, vscale(dt, v));
        a = a + dt * wv;
        w->b_pos[4 * b] = c.x; w->b_pos[4 * b + 1] = c.y; w->b_pos[4 * b + 2] = a;
        w->b_vel[4 * b] = v.x; w->b_vel[4 * b + 1] = v.y; w->b_vel[4 * b + 2] = wv;
        body_sync_transform(w, b);
    }
}

/* World.solve_toi: world.py:594-821 */
static void world_solve_toi(orc_world* w, double dt, int vel_iters)
{
    int nb = w->w_count[0];
    int cap = w->nc_cap;
    double* toi = (double*)malloc(sizeof(double) * (size_t)(cap + 1));
    uint8_t* toi_flag = (uint8_t*)calloc((size_t)cap + 1, 1);
    int* toi_count = (int*)calloc((size_t)cap + 1, sizeof(int));
    /* world.py:599-609 (step_complete is always true here: no sub_stepping) */
    for (int b = 0; b < nb; ++b) { w->b_flags[b] &= ~BODY_ISLAND; w->b_pos0[4 * b + 3] = 0.0; }
    for (int c = 0; c < w->w_count[2]; ++c) { w->c_flags[c] &= ~C_ISLAND; toi[c] = 1.0; }

    for (;;) {
        int min_contact = -1;
        double min_alpha = 1.0;
        int nc = w->w_count[2];
        for (int c = 0; c < nc; ++c) {
            if (!(w->c_flags[c] & C_ENABLED) || toi_count[c] > MAX_SUB_STEPS) continue;
            double alpha = 1.0;
            if (toi_flag[c]) alpha = toi[c];
            else {
                int fa = w->c_fix[2 * c], fb = w->c_fix[2 * c + 1];
                if (w->c_flags[c] & C_SENSOR) continue;                           /* world.py:631-633 */
                int ba = w->f_body[fa], bb = w->f_body[fb];
                int type_a = body_type(w, ba), type_b = body_type(w, bb);
                int active_a = body_awake(w, ba) && type_a != T_STATIC;
                int active_b = body_awake(w, bb) && type_b != T_STATIC;
                if (!active_a && !active_b) continue;
                int collide_a = (w->b_flags[ba] & BODY_BULLET) || type_a != T_DYNAMIC;
                int collide_b = (w->b_flags[bb] & BODY_BULLET) || type_b != T_DYNAMIC;
                if (!collide_a && !collide_b) continue;
                sweep_t sa = sweep_load(w, ba), sb = sweep_load(w, bb);
                double alpha0 = sa.alpha0;
                if (sa.alpha0 < sb.alpha0) { alpha0 = sb.alpha0; sweep_advance(&sa, alpha0); sweep_store(w, ba, &sa); }
                else if (sb.alpha0 < sa.alpha0) { sweep_advance(&sb, alpha0); sweep_store(w, bb, &sb); }
                dproxy pa = dproxy_get(w, w->f_shape[fa]), pb = dproxy_get(w, w->f_shape[fb]);
                double beta;
                int state = time_of_impact(&pa, &pb, sa, sb, 1.0, &beta);
                if (state == TOI_TOUCHING) { double t = alpha0 + (1.0 - alpha0) * beta; alpha = PYMIN(t, 1.0); }
                else alpha = 1.0;
                toi[c] = alpha;
                toi_flag[c] = 1;
            }
            if (alpha < min_alpha) { min_contact = c; min_alpha = alpha; }
        }
        if (min_contact < 0 || (1.0 - (10.0 * EPS)) < min_alpha) break;

        int mc = min_contact;
        int fa = w->c_fix[2 * mc], fb = w->c_fix[2 * mc + 1];
        int body_a = w->f_body[fa], body_b = w->f_body[fb];
        sweep_t backup_a = sweep_load(w, body_a), backup_b = sweep_load(w, body_b);
        body_advance(w, body_a, min_alpha);
        body_advance(w, body_b, min_alpha);
        contact_update(w, mc);
        toi_flag[mc] = 0;
        toi_count[mc] += 1;
        if (!(w->c_flags[mc] & C_ENABLED) || !(w->c_flags[mc] & C_TOUCHING)) {
            w->c_flags[mc] &= ~C_ENABLED;
            sweep_store(w, body_a, &backup_a);
            sweep_store(w, body_b, &backup_b);
            body_sync_transform(w, body_a);
            body_sync_transform(w, body_b);
            continue;
        }
        body_wake(w, body_a);
        body_wake(w, body_b);

        int ibodies[2 * MAX_TOI_CONTACTS + 2], icontacts[MAX_TOI_CONTACTS + 2];
        int nib = 0, nic = 0;
        w->b_island[body_a] = nib; ibodies[nib++] = body_a;
        w->b_island[body_b] = nib; ibodies[nib++] = body_b;
        icontacts[nic++] = mc;
        w->b_flags[body_a] |= BODY_ISLAND;
        w->b_flags[body_b] |= BODY_ISLAND;
        w->c_flags[mc] |= C_ISLAND;
        int seeds[2] = {body_a, body_b};
        for (int si = 0; si < 2; ++si) {
            int body = seeds[si];
            if (body_type(w, body) != T_DYNAMIC) continue;
            for (int c = 0; c < nc; ++c) {                      /* body._contacts: the world's list filtered by body */
                int cfa = w->c_fix[2 * c], cfb = w->c_fix[2 * c + 1];
                int cba = w->f_body[cfa], cbb = w->f_body[cfb];
                if (cba != body && cbb != body) continue;
                if (nib == 2 * MAX_TOI_CONTACTS) break;
                if (nic == MAX_TOI_CONTACTS) break;
                if (w->c_flags[c] & C_ISLAND) continue;
                int other = (cba == body) ? cbb : cba;
                if (body_type(w, other) == T_DYNAMIC && !(w->b_flags[body] & BODY_BULLET) && !(w->b_flags[other] & BODY_BULLET))
                    continue;
                if (w->c_flags[c] & C_SENSOR) continue;                           /* world.py:757-759 */
                sweep_t backup = sweep_load(w, other);
                if (!(w->b_flags[other] & BODY_ISLAND)) body_advance(w, other, min_alpha);
                contact_update(w, c);
                if (!(w->c_flags[c] & C_ENABLED) || !(w->c_flags[c] & C_TOUCHING)) {
                    sweep_store(w, other, &backup);
                    body_sync_transform(w, other);
                    continue;
                }
                w->c_flags[c] |= C_ISLAND;
                icontacts[nic++] = c;
                if (w->b_flags[other] & BODY_ISLAND) continue;
                w->b_flags[other] |= BODY_ISLAND;
                if (body_type(w, other) != T_STATIC) body_wake(w, other);
                w->b_island[other] = nib; ibodies[nib++] = other;
            }
        }
        double sub_dt = (1.0 - min_alpha) * dt;
        island_solve_toi(w, ibodies, nib, icontacts, nic, sub_dt, vel_iters, w->b_island[body_a], w->b_island[body_b]);

        for (int i = 0; i < nib; ++i) {
            int body = ibodies[i];
            w->b_flags[body] &= ~BODY_ISLAND;
            if (body_type(w, body) != T_DYNAMIC) continue;
            /* Body._synchronize_fixtures: body.py:898-904 */
            v2 c0 = V(w->b_pos0[4 * body], w->b_pos0[4 * body + 1]); double a0 = w->b_pos0[4 * body + 2];
            v2 lcen = V(w->b_mass[4 * body + 2], w->b_mass[4 * body + 3]);
            xform xf1 = xf_from_sweep(c0, a0, lcen);
            xform xf2 = xf_row(w->b_xf + 4 * (size_t)body);
            for (int f = 0; f < w->w_count[1]; ++f) if (w->f_body[f] == body) fixture_synchronize(w, f, &xf1, &xf2);
            for (int c = 0; c < nc; ++c) {
                int cba = w->f_body[w->c_fix[2 * c]], cbb = w->f_body[w->c_fix[2 * c + 1]];
                if (cba != body && cbb != body) continue;
                toi_flag[c] = 0;
                w->c_flags[c] &= ~C_ISLAND;
            }
        }
        {
            int before = w->w_count[2];
            cm_find_new_contacts(w);
            for (int c = before; c < w->w_count[2]; ++c) { toi[c] = 1.0; toi_flag[c] = 0; toi_count[c] = 0; }
        }
    }
    free(toi); free(toi_flag); free(toi_count);
}


/* =====================================================================================================
 * Queries: World.query_aabb / World.ray_cast world.py:371-440, DynamicTree.query / ray_cast
 * dynamictree.py:190-262, Shape.test_point / ray_cast shapes.py:141-176, 256-314, 655-700.  The tree only
 * prunes: a leaf is reported iff its fat box passes the tests below, so the answers are restated as a scan
 * over the proxies in fixture order (the reference reports them in tree order; the SET is the same).
 * ===================================================================================================== */
/* fixtures whose fat AABB overlaps `box` (lx, ly, ux, uy), ascending; returns the total count */
int orc_query_aabb(const orc_world* w, const double* box, int32_t* out, int max_out)
{
    int n = 0, nf = w->w_count[1];
    for (int f = 0; f < nf; ++f) {
        if (!aabb_overlap(w->f_fat + 4 * (size_t)f, box)) continue;            /* node.aabb.test_overlap(aabb) */
        if (n < max_out) out[n] = f;
        ++n;
    }
    return n;
}

/* Shape.ray_cast: 1 on a hit */
static int shape_ray_cast(const shape_t* s, const xform* xf, v2 p1, v2 p2, double max_fraction, v2* normal_out, double* fraction_out)
{
    if (s->type == S_CIRCLE) {                                                  /* shapes.py:147-176 */
        v2 position = vadd(xf->p, rot_mul(xf, V(s->d[0], s->d[1])));
        v2 sv = vsub(p1, position);
        double b = vdot(sv, sv) - sq(s->radius);
        v2 r = vsub(p2, p1);
        double c = vdot(sv, r), rr = vdot(r, r);
        double sigma = c * c - rr * b;
        if (sigma < 0.0 || rr < EPS) return 0;
        double a = -(c + sqrt(sigma));
        if (0.0 <= a && a <= max_fraction * rr) {
            a /= rr;
            *fraction_out = a;
            *normal_out = vnorm(vadd(sv, vscale(a, r)));
            return 1;
        }
        return 0;
    }
    if (s->type == S_POLYGON) {                                                 /* shapes.py:266-314 */
        v2 q1 = rot_mul_t(xf, vsub(p1, xf->p)), q2 = rot_mul_t(xf, vsub(p2, xf->p));
        v2 d = vsub(q2, q1);
        double lower = 0.0, upper = max_fraction;
        int tagged = -1;
        for (int i = 0; i < s->count; ++i) {
            v2 normal = pnorm(s, i), vertex = pvert(s, i);
            double numerator = vdot(normal, vsub(vertex, q1));
            double denominator = vdot(normal, d);
            if (denominator == 0.0) {
                if (numerator < 0.0) return 0;
            } else {
                if (denominator < 0.0 && numerator < lower * denominator) { lower = numerator / denominator; tagged = i; }
                else if (denominator > 0.0 && numerator < upper * denominator) upper = numerator / denominator;
            }
            if (upper < lower) return 0;
        }
        if (tagged >= 0) {
            *normal_out = rot_mul(xf, pnorm(s, tagged));
            *fraction_out = lower;
            return 1;
        }
        return 0;
    }
    {                                                                           /* edge: shapes.py:655-700 */
        v2 q1 = rot_mul_t(xf, vsub(p1, xf->p)), q2 = rot_mul_t(xf, vsub(p2, xf->p));
        v2 d = vsub(q2, q1);
        v2 v1 = V(s->d[0], s->d[1]), v2_ = V(s->d[2], s->d[3]);
        v2 e = vsub(v2_, v1);
        v2 normal = vnorm(V(e.y, -e.x));
        double numerator = vdot(normal, vsub(v1, q1));
        double denominator = vdot(normal, d);
        if (denominator == 0.0) return 0;
        double t = numerator / denominator;
        if (t < 0.0 || 1.0 < t) return 0;
        v2 q = vadd(q1, vscale(t, d));
        v2 r = vsub(v2_, v1);
        double rr = vdot(r, r);
        if (rr == 0.0) return 0;
        double sp = vdot(vsub(q, v1), r) / rr;
        if (sp < 0.0 || 1.0 < sp) return 0;
        *fraction_out = t;
        if (numerator > 0.0) normal = vneg(normal);
        *normal_out = normal;                           /* (the port returns the normal in the edge's LOCAL frame) */
        return 1;
    }
}

/* DynamicTree.ray_cast's tests on a leaf box (dynamictree.py:226-246), un-clipped segment: max_fraction = 1 */
static int ray_leaf_test(const double* box, v2 p1, v2 p2)
{
    v2 r = vnorm(vsub(p2, p1));
    v2 v = scross(1.0, r);
    v2 abs_v = V(fabs(v.x), fabs(v.y));
    v2 t = vadd(p1, vscale(1.0, vsub(p2, p1)));
    double seg[4] = {PYMIN(p1.x, t.x), PYMIN(p1.y, t.y), PYMAX(p1.x, t.x), PYMAX(p1.y, t.y)};
    if (!aabb_overlap(box, seg)) return 0;
    v2 c = vscale(0.5, vadd(V(box[0], box[1]), V(box[2], box[3])));
    v2 h = vscale(0.5, vsub(V(box[2], box[3]), V(box[0], box[1])));
    double separation = fabs(vdot(v, vsub(p1, c))) - vdot(abs_v, h);
    return !(separation > 0.0);
}

/* World.ray_cast with a callback that never clips (returns 1): every fixture the ray hits, ascending fixture index.
 * out_hit[k] = {fraction, normal.x, normal.y, point.x, point.y}; returns the total count */
int orc_ray_cast(const orc_world* w, const double* p1p2, int32_t* out_fix, double* out_hit, int max_out)
{
    v2 p1 = V(p1p2[0], p1p2[1]), p2 = V(p1p2[2], p1p2[3]);
    int n = 0, nf = w->w_count[1];
    for (int f = 0; f < nf; ++f) {
        if (!ray_leaf_test(w->f_fat + 4 * (size_t)f, p1, p2)) continue;
        shape_t s = shape_get(w, w->f_shape[f]);
        xform xf = xf_row(w->b_xf + 4 * (size_t)w->f_body[f]);
        v2 normal; double fraction;
        if (!shape_ray_cast(&s, &xf, p1, p2, 1.0, &normal, &fraction)) continue;
        if (n < max_out) {
            v2 point = vadd(vscale(1.0 - fraction, p1), vscale(fraction, p2));   /* world.py:427 */
            out_fix[n] = f;
            out_hit[5 * n] = fraction; out_hit[5 * n + 1] = normal.x; out_hit[5 * n + 2] = normal.y;
            out_hit[5 * n + 3] = point.x; out_hit[5 * n + 4] = point.y;
        }
        ++n;
    }
    return n;
}

/* Fixture.test_point (fixture.py:228-233; circles and polygons: shapes.py:141-145, 256-263; an edge contains no point) */
int orc_test_point(const orc_world* w, int f, const double* p)
{
    shape_t s = shape_get(w, w->f_shape[f]);
    xform xf = xf_row(w->b_xf + 4 * (size_t)w->f_body[f]);
    v2 pt = V(p[0], p[1]);
    if (s.type == S_CIRCLE) {
        v2 center = vadd(xf.p, rot_mul(&xf, V(s.d[0], s.d[1])));
        v2 d = vsub(pt, center);
        return vdot(d, d) <= sq(s.radius);
    }
    if (s.type == S_POLYGON) {
        v2 pl = rot_mul_t(&xf, vsub(pt, xf.p));
        for (int i = 0; i < s.count; ++i)
            if (vdot(pnorm(&s, i), vsub(pl, pvert(&s, i))) > 0.0) return 0;
        return 1;
    }
    return 0;
}

/* ---- World.step: world.py:307-354 ---- */
int orc_step(orc_world* w, double dt, double inv_dt0, int vel_iters, int pos_iters)
{
    if (w->w_flags[0] & W_NEW_FIXTURE) {
        cm_find_new_contacts(w);
        w->w_flags[0] &= ~W_NEW_FIXTURE;
    }
    double dt_ratio = inv_dt0 * dt;
    cm_collide(w);
    if (dt > 0.0) world_solve(w, dt, dt_ratio, vel_iters, pos_iters);
    if ((w->w_flags[0] & W_CONTINUOUS) && dt > 0.0) world_solve_toi(w, dt, vel_iters);
    if (w->w_flags[0] & W_AUTO_CLEAR) {
        int nb = w->w_count[0];
        for (int b = 0; b < nb; ++b) { w->b_force[2 * b] = 0.0; w->b_force[2 * b + 1] = 0.0; w->b_vel[4 * b + 3] = 0.0; }
    }
    return (w->w_status[0] != 0) ? -1 : 0;
}

int orc_run(orc_world* w, double dt, double inv_dt0, int vel_iters, int pos_iters, int n_steps)
{
    int rc = 0;
    for (int s = 0; s < n_steps; ++s) {
        rc |= orc_step(w, dt, inv_dt0, vel_iters, pos_iters);
        if (dt > 0.0) inv_dt0 = 1.0 / dt;
    }
    return rc;
}

int orc_collide_pair(const orc_world* w, int32_t sa, int32_t sb, const double* xfa, const double* xfb,
                     int32_t* mtype, double* mf, double* pts, uint32_t* id)
{
    manifold_t m; memset(&m, 0, sizeof(m));
    shape_t a = shape_get(w, sa), b = shape_get(w, sb);
    xform xa = xf_row(xfa), xb = xf_row(xfb);
    evaluate(&m, &a, &xa, &b, &xb);
    *mtype = m.type | (m.count << 8);
    mf[0] = m.local_normal.x; mf[1] = m.local_normal.y; mf[2] = m.local_point.x; mf[3] = m.local_point.y;
    pts[0] = m.pt[0].x; pts[1] = m.pt[0].y; pts[2] = m.pt[1].x; pts[3] = m.pt[1].y;
    id[0] = m.id[0]; id[1] = m.id[1];
    return 0;
}

const char* orc_field_names(void)
{
    return "w_count,w_flags,w_status,w_gravity,w_island,"
           "b_xf,b_pos,b_vel,b_pos0,b_force,b_mass,b_damp,b_flags,b_island,"
           "f_body,f_shape,f_mat,f_filter,f_fat,f_tight,f_moved,"
           "c_fix,c_flags,c_mat,c_mtype,c_mf,c_pts,c_imp,c_id,"
           "p_pairs,l_body,l_contact,l_start,s_meta,s_radius,s_data,"
           "w_joint,j_def,j_par,j_imp,j_tmp,j_adj,b_jadj,l_joint,l_jstart";
}
