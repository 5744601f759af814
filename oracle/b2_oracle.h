/*
 * b2_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, fp64, single world, strictly sequential) of the reference's
 * World.step hot path (pypybox2d r331, /root/reference/pypybox2d/world.py:307-354 and callees).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or as the reported CPU baseline.  The product
 * (pypybox2d_b200/) never links, imports or calls anything under oracle/.
 *
 * Parity pin: the restatement is checked step by step against the live Python reference
 * (pair-emission order pinned, SURVEY §0 finding 1) by oracle/gen_golden.py, and against the
 * committed fixtures it wrote under tests/golden/ by tests/test_oracle_golden.py.
 *
 * The world is a struct of pointers to ONE world's rows of the arrays described in
 * pypybox2d_b200/schema.py (FIELDS, same names, same order), all reals as double.
 */
#ifndef B2_ORACLE_H
#define B2_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_world {
    /* capacities */
    int32_t nb_cap, nf_cap, nc_cap, npair_cap, nl_cap, n_shapes, nj_cap;
    /* per world */
    int32_t*  w_count;    /* [4] n_bodies, n_fixtures, n_contacts, n_pairs */
    uint32_t* w_flags;    /* [1] */
    uint32_t* w_status;   /* [1] */
    double*   w_gravity;  /* [2] */
    int32_t*  w_island;   /* [4] n_islands, n_island_bodies, n_island_contacts, max pos iters used */
    /* per body */
    double*   b_xf;       /* [nb][4] */
    double*   b_pos;      /* [nb][4] */
    double*   b_vel;      /* [nb][4] */
    double*   b_pos0;     /* [nb][4] */
    double*   b_force;    /* [nb][2] */
    double*   b_mass;     /* [nb][4] */
    double*   b_damp;     /* [nb][4] */
    uint32_t* b_flags;    /* [nb] */
    int32_t*  b_island;   /* [nb] */
    /* per fixture / proxy */
    int32_t*  f_body;
    int32_t*  f_shape;
    double*   f_mat;      /* [nf][2] */
    uint32_t* f_filter;   /* [nf][2] */
    double*   f_fat;      /* [nf][4] */
    double*   f_tight;    /* [nf][4] */
    uint8_t*  f_moved;
    /* per contact */
    int32_t*  c_fix;      /* [nc][2] */
    uint32_t* c_flags;
    double*   c_mat;      /* [nc][2] */
    int32_t*  c_mtype;
    double*   c_mf;       /* [nc][4] */
    double*   c_pts;      /* [nc][4] */
    double*   c_imp;      /* [nc][4] */
    uint32_t* c_id;       /* [nc][2] */
    /* observable scratch */
    int32_t*  p_pairs;    /* [npair][2] */
    int32_t*  l_body;     /* [nl] */
    int32_t*  l_contact;  /* [nc] */
    int32_t*  l_start;    /* [nb][2] */
    /* shared shape table */
    int32_t*  s_meta;
    double*   s_radius;
    double*   s_data;     /* [n_shapes][32] */
    /* joints (layout: pypybox2d_b200/joints.py) */
    int32_t*  w_joint;    /* [4] n_joints, n_island_joints */
    int32_t*  j_def;      /* [nj][4] type, body_a, body_b, flags | limit_state << 8 */
    double*   j_par;      /* [nj][12] */
    double*   j_imp;      /* [nj][4] */
    double*   j_tmp;      /* [nj][28] */
    int32_t*  j_adj;      /* [2 nj] Body.joints rows */
    int32_t*  b_jadj;     /* [nb][2] start, count in j_adj */
    int32_t*  l_joint;    /* [nj] island joint lists */
    int32_t*  l_jstart;   /* [nb] per island: start in l_joint */
} orc_world;

/* One World.step (world.py:307-354) with continuous_physics=False.  inv_dt0 is World._inv_dt0
 * before the call.  Returns 0, or <0 on capacity overflow (also flagged in w_status). */
int orc_step(orc_world* w, double dt, double inv_dt0, int vel_iters, int pos_iters);

/* Queries (world.py:371-440): fixtures whose fat AABB overlaps `box`; every fixture hit by the segment p1p2 = (x1, y1, x2, y2)
 * with out_hit[k] = {fraction, normal.x, normal.y, point.x, point.y}; Fixture.test_point.  Lists are ascending in the
 * fixture index; the returned count is the total (it may exceed max_out). */
int orc_query_aabb(const orc_world* w, const double* box, int32_t* out, int max_out);
int orc_ray_cast(const orc_world* w, const double* p1p2, int32_t* out_fix, double* out_hit, int max_out);
int orc_test_point(const orc_world* w, int f, const double* p);

/* n_steps calls of orc_step with the reference's inv_dt0 bookkeeping (world.py:348-349). */
int orc_run(orc_world* w, double dt, double inv_dt0, int vel_iters, int pos_iters, int n_steps);

/* Single narrowphase evaluation (test hook): runs the collider for shape rows sa/sb with
 * transforms xfa/xfb = (px, py, cos, sin) on a zero manifold; out: mtype, mf[4], pts[4], id[2]. */
int orc_collide_pair(const orc_world* w, int32_t sa, int32_t sb, const double* xfa, const double* xfb,
                     int32_t* mtype, double* mf, double* pts, uint32_t* id);

const char* orc_field_names(void);

#ifdef __cplusplus
}
#endif
#endif
