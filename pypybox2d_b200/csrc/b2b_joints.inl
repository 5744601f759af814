// b2b_joints.inl -- joints of the island solver, included INSIDE world_step<real> (b2b_step.cuh).
//
// RevoluteJoint (pypybox2d/joints/revolute.py:208-453), DistanceJoint (joints/distance.py:155-311), WeldJoint
// (joints/weld.py:106-257), PrismaticJoint (joints/prismatic.py:232-513), RopeJoint (joints/rope.py:106-248), FrictionJoint
// (joints/friction.py:105-233), WheelJoint (joints/wheel.py:202-416), PulleyJoint (joints/pulley.py:156-318) and MouseJoint (joints/mouse.py:138-229) on the reference's pure-Python Vec3 / Mat33 (common.py:227-358, 533-614), operand by operand: the
// fp64 host build of this file is bit-identical to the oracle, which is pinned bit-identical to the live reference.
//
// Island.solve visits an island's joints one after the other, before the contacts in every velocity iteration and after them in
// every position iteration (island.py:107-113, 145-148); a joint reads and writes the velocities / positions of its two bodies,
// which live in the team's shared memory (state_get / state_set).  Joint lists are short next to contact lists (a bridge has one
// joint per plank, a pile has hundreds of contacts), and consecutive joints of a chain share a body, so the sweep over an island's
// joints is a dependent chain like the contact sweep: one lane (the team's first) walks it while the others wait at the barrier.
// Only batches created with a joint capacity (b2b_caps::nj > 0) run the kernels instantiated with these hooks.
//
// Rows (layout: pypybox2d_b200/joints.py): j_def = (type, body_a, body_b, flags | _limit_state << 8), j_par[12], j_imp = (_impulse
// x, y, z, _motor_impulse), j_tmp[20] = what _init_velocity_constraints leaves for the sweeps (ra, rb, then per type: the 3x3 mass
// matrix by columns and the motor mass; or u, mass, gamma, bias; prismatic: axis, perp, s1, s2, a1, a2, motor mass and the six
// distinct entries of its symmetric K); tmp[20:24] = the two bodies' inverse mass and inertia, tmp[24:28] = the local anchors minus
// the local centres of mass, both copied by the set-up pass so that the velocity and position passes need no row of the bodies
// (their addresses would only be known after the joint's definition has arrived: a third dependent round trip per fetch).

enum { JOINT_REVOLUTE = 1, JOINT_DISTANCE = 2, JOINT_WELD = 3, JOINT_PRISMATIC = 4, JOINT_ROPE = 5, JOINT_FRICTION = 6, JOINT_WHEEL = 7, JOINT_PULLEY = 8, JOINT_MOUSE = 9 };
enum { JF_COLLIDE = 0x1, JF_LIMIT = 0x2, JF_MOTOR = 0x4, JF_ISLAND = 0x8, JF_LS_SHIFT = 8 };
enum { LS_INACTIVE = 0, LS_LOWER = 1, LS_UPPER = 2, LS_EQUAL = 3 };               // joints/joint.py:36
enum { JOINT_PAR = 12, JOINT_TMP = 28 };        // (schema.py JOINT_PAR / JOINT_TMP)

struct V3 { real x, y, z; };
B2B_D static V3 mk3(real x, real y, real z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
B2B_D static real dot3(V3 a, V3 b) { return (a.x * b.x) + (a.y * b.y) + (a.z * b.z); }                                  // common.py:342-344
B2B_D static V3 cross3(V3 s, V3 v) { return mk3(s.y * v.z - s.z * v.y, s.z * v.x - s.x * v.z, s.x * v.y - s.y * v.x); } // common.py:353-358
B2B_D static V3 neg3(V3 a) { return mk3(-a.x, -a.y, -a.z); }
struct M33 { V3 c1, c2, c3; };
B2B_D static V3 solve33(const M33& m, V3 b)                                                                             // common.py:587-599
{
    real det = dot3(m.c1, cross3(m.c2, m.c3));
    if (det != real(0)) det = real(1) / det;
    return mk3(det * dot3(b, cross3(m.c2, m.c3)), det * dot3(m.c1, cross3(b, m.c3)), det * dot3(m.c1, cross3(m.c2, b)));
}
B2B_D static V2 solve22(const M33& m, V2 vec)                                                                           // common.py:601-614
{
    const real a11 = m.c1.x, a12 = m.c2.x, a21 = m.c1.y, a22 = m.c2.y;
    real det = a11 * a22 - a12 * a21;
    if (det != real(0)) det = real(1) / det;
    return mk<real>(det * (a22 * vec.x - a12 * vec.y), det * (a11 * vec.y - a21 * vec.x));
}
B2B_D static XF rot_of(real a)                                                                                          // Mat22(angle=a)
{
    real s, c;
    sincos_r(a, &s, &c);
    return xf_make<real>(mk<real>(real(0), real(0)), c, s);
}
B2B_D static M33 m33_load(const real* t) { M33 m; m.c1 = mk3(t[0], t[1], t[2]); m.c2 = mk3(t[3], t[4], t[5]); m.c3 = mk3(t[6], t[7], t[8]); return m; }
B2B_D static M33 m33_sym(const real* k)      // Mat33((k11, k12, k13), (k12, k22, k23), (k13, k23, k33))
{
    M33 m; m.c1 = mk3(k[0], k[1], k[2]); m.c2 = mk3(k[1], k[3], k[4]); m.c3 = mk3(k[2], k[4], k[5]); return m;
}
B2B_D static void m33_store(real* t, const M33& m)
{
    t[0] = m.c1.x; t[1] = m.c1.y; t[2] = m.c1.z; t[3] = m.c2.x; t[4] = m.c2.y; t[5] = m.c2.z; t[6] = m.c3.x; t[7] = m.c3.y; t[8] = m.c3.z;
}

// One joint's rows in registers.  The lanes of a team fetch the rows of t_n consecutive island joints at once (one memory round
// trip for t_n joints) and then take turns applying them to the shared body state in list order (joint_sweep below): the
// dependent part of a joint visit is then shared-memory reads, arithmetic and shared-memory writes only.
struct alignas(16) int4v { int32_t x, y, z, w; };
struct jrow {
    int k;
    int32_t def[4];
    real par[JOINT_PAR], imp[4], tmp[JOINT_TMP];
    quad mqa, mqb, pa, pb;
};
enum { JFETCH_INIT = 0, JFETCH_VELOCITY = 1, JFETCH_POSITION = 2 };
B2B_D void joint_fetch(int k, jrow& r, int what) const
{
    r.k = k;
    const int4v d = *reinterpret_cast<const int4v*>(j_def + 4 * (size_t)k);
    r.def[0] = d.x; r.def[1] = d.y; r.def[2] = d.z; r.def[3] = d.w;
    const quad* par = reinterpret_cast<const quad*>(j_par + JOINT_PAR * (size_t)k);
    const quad p0 = par[0], p1 = par[1], p2 = par[2];
    r.par[0] = p0.x; r.par[1] = p0.y; r.par[2] = p0.z; r.par[3] = p0.w; r.par[4] = p1.x; r.par[5] = p1.y; r.par[6] = p1.z; r.par[7] = p1.w;
    r.par[8] = p2.x; r.par[9] = p2.y; r.par[10] = p2.z; r.par[11] = p2.w;
    if (what != JFETCH_POSITION) {
        const quad q = ldq(j_imp + 4 * (size_t)k);
        r.imp[0] = q.x; r.imp[1] = q.y; r.imp[2] = q.z; r.imp[3] = q.w;
    }
    if (what != JFETCH_INIT) {
        const quad* t = reinterpret_cast<const quad*>(j_tmp + JOINT_TMP * (size_t)k);
#pragma unroll
        for (int i = 0; i < JOINT_TMP / 4; ++i) { const quad q = t[i]; r.tmp[4 * i] = q.x; r.tmp[4 * i + 1] = q.y; r.tmp[4 * i + 2] = q.z; r.tmp[4 * i + 3] = q.w; }
    }
    if (what == JFETCH_INIT) {
        const int ba = (r.def[1] >= 0) ? r.def[1] : r.def[2];      // a joint on one body (MouseJoint) has body_a = -1
        r.mqa = ldq(b_mass + 4 * (size_t)ba); r.mqb = ldq(b_mass + 4 * (size_t)r.def[2]);
        r.pa = ldq(b_pos + 4 * (size_t)ba); r.pb = ldq(b_pos + 4 * (size_t)r.def[2]);
    } else {
        r.mqa = mkq(r.tmp[20], r.tmp[21], real(0), real(0)); r.mqb = mkq(r.tmp[22], r.tmp[23], real(0), real(0));
    }
}
B2B_D void joint_store(const jrow& r, int what)
{
    if (what == JFETCH_POSITION) return;
    stq(j_imp + 4 * (size_t)r.k, mkq(r.imp[0], r.imp[1], r.imp[2], r.imp[3]));
    if (what == JFETCH_INIT) {
        j_def[4 * (size_t)r.k + 3] = r.def[3];
        quad* t = reinterpret_cast<quad*>(j_tmp + JOINT_TMP * (size_t)r.k);
#pragma unroll
        for (int i = 0; i < JOINT_TMP / 4; ++i) t[i] = mkq(r.tmp[4 * i], r.tmp[4 * i + 1], r.tmp[4 * i + 2], r.tmp[4 * i + 3]);
    }
}
// One pass over the island's joints [j0, j1) of l_joint, in list order (island.py:107-113, 145-148); `apply` is called by the lane
// whose turn it is with the joint's rows; returns on every lane the AND of the values apply returned (the position pass's
// "joints_okay").  B2B_SYNC orders the lanes' accesses to the shared body state.
template <class Apply> B2B_D int joint_sweep(int j0, int j1, int what, Apply apply)
{
    int all_ok = 1;
    for (int k0 = j0; k0 < j1; k0 += t_n) {
        jrow r;
        const int mine = k0 + t_id;
        if (mine < j1) joint_fetch(l_joint[mine], r, what);
#if defined(__CUDA_ARCH__)
        // the rows of this lane's NEXT joint are asked for now (L2 prefetch): they arrive while the team applies this round's joints
        if (mine + t_n < j1) {
            const int nk = l_joint[mine + t_n];
            B2B_PREFETCH(j_def + 4 * (size_t)nk);
            B2B_PREFETCH(j_par + JOINT_PAR * (size_t)nk); B2B_PREFETCH(j_par + JOINT_PAR * (size_t)nk + JOINT_PAR - 1);
            if (what != JFETCH_POSITION) B2B_PREFETCH(j_imp + 4 * (size_t)nk);
            if (what != JFETCH_INIT) { B2B_PREFETCH(j_tmp + JOINT_TMP * (size_t)nk); B2B_PREFETCH(j_tmp + JOINT_TMP * (size_t)nk + JOINT_TMP - 1); }
        }
#endif
        const int n = (j1 - k0 < t_n) ? j1 - k0 : t_n;
        for (int i = 0; i < n; ++i) {
            if (t_id == i) {
                if (!apply(r)) all_ok = 0;
                joint_store(r, what);
            }
            B2B_SYNC();
        }
    }
#if defined(__CUDA_ARCH__)
    for (int o = t_n >> 1; o > 0; o >>= 1) all_ok &= __shfl_xor_sync(t_mask, all_ok, o);
#endif
    return all_ok;
}

// MouseJoint (mouse.py:138-229): one body (j_def body_a = -1); par: [2:4] local_anchor_b, [4:6] target, [6] max_force, [7] frequency,
// [8] damping_ratio, [9] Body.mass; tmp: [2:4] rb, [4:8] the inverse of K by columns, [8:10] c, [10] gamma
B2B_D void mouse_init(jrow& r, real dt, real dt_ratio, bool warm)
{
    const real* par = r.par;
    real* imp = r.imp;
    real* tmp = r.tmp;
    const int bb = r.def[2];
    const real mb = r.mqb.x, ib = r.mqb.y;
    const V2 lcb = mk<real>(r.mqb.z, r.mqb.w);
    const V2 cb = mk<real>(r.pb.x, r.pb.y);
    const real ab = r.pb.z;
    const quad sb = state_get(bb);
    V2 vb = mk<real>(sb.x, sb.y);
    real wb = sb.z;
    const XF qb = rot_of(ab);
    const real mass = par[9];
    const real omega = real(2) * K<real>::pi * par[7];
    const real d = real(2) * mass * par[8] * omega;
    const real kk = mass * sq_r(omega);
    real gamma = dt * (d + dt * kk);
    if (gamma != real(0)) gamma = real(1) / gamma;
    const real beta = dt * kk * gamma;
    const V2 rb = rot_mul(qb, mk<real>(par[2], par[3]) - lcb);
    const real k1x = mb + ib * sq_r(rb.y) + gamma, k1y = -ib * rb.x * rb.y;
    const real k2x = k1y, k2y = mb + ib * sq_r(rb.x) + gamma;
    real det = k1x * k2y - k2x * k1y;
    if (det != real(0)) det = real(1) / det;
    tmp[2] = rb.x; tmp[3] = rb.y;
    tmp[4] = det * k2y; tmp[5] = -det * k1y; tmp[6] = -det * k2x; tmp[7] = det * k1x;
    const V2 c = beta * (cb + rb - mk<real>(par[4], par[5]));
    tmp[8] = c.x; tmp[9] = c.y; tmp[10] = gamma;
    wb = wb * real(0.98);
    real i0 = imp[0], i1 = imp[1];
    if (warm) {
        i0 = dt_ratio * i0; i1 = dt_ratio * i1;
        const V2 I = mk<real>(i0, i1);
        vb = vb + mb * I;
        wb = wb + ib * cross(rb, I);
    } else { i0 = real(0); i1 = real(0); }
    imp[0] = i0; imp[1] = i1;
    state_set(bb, vb.x, vb.y, wb);
}
B2B_D void mouse_solve_velocity(jrow& r, real dt)
{
    const real* par = r.par;
    real* imp = r.imp;
    const real* tmp = r.tmp;
    const int bb = r.def[2];
    const real mb = r.mqb.x, ib = r.mqb.y;
    const quad sb = state_get(bb);
    V2 vb = mk<real>(sb.x, sb.y);
    real wb = sb.z;
    const V2 rb = mk<real>(tmp[2], tmp[3]);
    const V2 Cdot = vb + cross_sv(wb, rb);
    const V2 rhs = -(Cdot + mk<real>(tmp[8], tmp[9]) + tmp[10] * mk<real>(imp[0], imp[1]));
    V2 impulse = mk<real>(tmp[4] * rhs.x + tmp[6] * rhs.y, tmp[5] * rhs.x + tmp[7] * rhs.y);
    const V2 old_impulse = mk<real>(imp[0], imp[1]);
    V2 cur = old_impulse + impulse;
    const real max_impulse = dt * par[6];
    if (sq_r(cur.x) + sq_r(cur.y) > sq_r(max_impulse)) { const real f = max_impulse / sqrt_r(sq_r(cur.x) + sq_r(cur.y)); cur = mk<real>(f * cur.x, f * cur.y); }
    imp[0] = cur.x; imp[1] = cur.y;
    impulse = cur - old_impulse;
    vb = vb + mb * impulse;
    wb = wb + ib * cross(rb, impulse);
    state_set(bb, vb.x, vb.y, wb);
}

// _init_velocity_constraints (revolute.py:208-300, distance.py:155-250, weld.py:106-169).  Velocities: the team's shared state;
// positions: b_pos (the sweeps before this step's integration -- the velocity kernel runs before integrate_position).
// (HOT: only the revolute joint's code -- the instantiation for batches whose joints are all revolute joints (bridges, chains,
// ragdolls: b2b_caps::joint_types), where the type dispatch compiles away.  Measured on bridge x 131,072 worlds, velocity kernel:
// 3.16 ms when four joint types existed, 3.71 ms with all nine in line; a hot / cold split INSIDE one kernel -- the other eight
// types behind one out-of-line call -- is slower than either (4.33 ms taking the caller's rows, 4.62 ms taking a copy:
// the rows then live in local memory).  B2B_JOINT_COLD keeps those experiments.)
template <bool HOT> B2B_D void joint_init_impl(jrow& r, real dt, real dt_ratio, bool warm)
{
    int32_t* def = r.def;
    const real* par = r.par;
    real* imp = r.imp;
    real* tmp = r.tmp;
#pragma unroll
    for (int i = 0; i < JOINT_TMP; ++i) tmp[i] = real(0);
    tmp[20] = r.mqa.x; tmp[21] = r.mqa.y; tmp[22] = r.mqb.x; tmp[23] = r.mqb.y;
    {
        const V2 la = mk<real>(par[0], par[1]) - mk<real>(r.mqa.z, r.mqa.w), lb = mk<real>(par[2], par[3]) - mk<real>(r.mqb.z, r.mqb.w);
        tmp[24] = la.x; tmp[25] = la.y; tmp[26] = lb.x; tmp[27] = lb.y;
    }
    if (!HOT && def[0] == JOINT_MOUSE) { mouse_init(r, dt, dt_ratio, warm); return; }
    const int type = HOT ? (int)JOINT_REVOLUTE : def[0];
    const int ba = def[1], bb = def[2], flags = def[3];
    const quad mqa = r.mqa, mqb = r.mqb;
    const real ma = mqa.x, ia = mqa.y, mb = mqb.x, ib = mqb.y;
    const V2 lca = mk<real>(mqa.z, mqa.w), lcb = mk<real>(mqb.z, mqb.w);
    const quad pa = r.pa, pb = r.pb;
    const V2 ca = mk<real>(pa.x, pa.y), cb = mk<real>(pb.x, pb.y);
    const real aa = pa.z, ab = pb.z;
    const quad sa = state_get(ba), sb = state_get(bb);
    V2 va = mk<real>(sa.x, sa.y), vb = mk<real>(sb.x, sb.y);
    real wa = sa.z, wb = sb.z;
    const XF qa = rot_of(aa), qb = rot_of(ab);
    const V2 ra = rot_mul(qa, mk<real>(par[0], par[1]) - lca), rb = rot_mul(qb, mk<real>(par[2], par[3]) - lcb);
    tmp[0] = ra.x; tmp[1] = ra.y; tmp[2] = rb.x; tmp[3] = rb.y;
    if (type == JOINT_REVOLUTE) {
        M33 m;
        m.c3 = mk3(-ra.y * ia - rb.y * ib, ra.x * ia + rb.x * ib, ia + ib);
        m.c2 = mk3(-ra.y * ra.x * ia - rb.y * rb.x * ib, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, m.c3.y);
        m.c1 = mk3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, m.c2.x, m.c3.x);
        m33_store(tmp + 4, m);
        real motor_mass = ia + ib;
        const bool fixed_rotation = (motor_mass == real(0));
        if (motor_mass > real(0)) motor_mass = real(1) / motor_mass;
        tmp[13] = motor_mass;
        int ls = (flags >> JF_LS_SHIFT) & 3;
        real i0 = imp[0], i1 = imp[1], i2 = imp[2], im = imp[3];
        if (!(flags & JF_MOTOR) || fixed_rotation) im = real(0);
        if ((flags & JF_LIMIT) && !fixed_rotation) {
            const real joint_angle = ab - aa - par[4];
            if (abs_r(par[6] - par[5]) < real(2) * K<real>::angular_slop) ls = LS_EQUAL;
            else if (joint_angle <= par[5]) {
                if (ls != LS_LOWER) i2 = i2 - i2;
                ls = LS_LOWER;
            } else if (joint_angle >= par[6]) {
                if (ls != LS_UPPER) i2 = i2 - i2;
                ls = LS_UPPER;
            } else {
                ls = LS_INACTIVE;
                i2 = i2 - i2;
            }
        } else ls = LS_INACTIVE;
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (warm) {
            i0 = dt_ratio * i0; i1 = dt_ratio * i1; i2 = dt_ratio * i2;
            im = im * dt_ratio;
            const V2 P = mk<real>(i0, i1);
            va = va - ma * P;
            wa = wa - ia * (cross(ra, P) + im + i2);
            vb = vb + mb * P;
            wb = wb + ib * (cross(rb, P) + im + i2);
        } else { i0 = i1 = i2 = real(0); im = real(0); }
        imp[0] = i0; imp[1] = i1; imp[2] = i2; imp[3] = im;
    } else if (type == JOINT_DISTANCE) {
        const V2 u = cb + rb - ca - ra;
        const real length = sqrt_r(sq_r(u.x) + sq_r(u.y));
        const V2 un = (length > K<real>::linear_slop) ? (real(1) / length) * u : mk<real>(real(0), real(0));   // self._u; the local `u` is not normalised
        tmp[4] = un.x; tmp[5] = un.y;
        const real cr_au = sq_r(cross(ra, u)), cr_bu = sq_r(cross(rb, u));
        const real inv_mass = ma + ia * cr_au + mb + ib * cr_bu;
        real mass = (inv_mass != real(0)) ? real(1) / inv_mass : real(0);
        real gamma = real(0), bias = real(0);
        if (par[5] > real(0)) {
            const real C = length - par[4];
            const real omega = real(2) * K<real>::pi * par[5];
            const real d = real(2) * mass * par[6] * omega;
            const real kk = mass * sq_r(omega);
            gamma = dt * (d + dt * kk);
            gamma = (gamma != real(0)) ? real(1) / gamma : real(0);
            bias = C * dt * kk * gamma;
            const real m2 = inv_mass + gamma;
            mass = (m2 != real(0)) ? real(1) / m2 : real(0);
        }
        tmp[6] = mass; tmp[7] = gamma; tmp[8] = bias;
        real i0 = imp[0];
        if (warm) {
            i0 = i0 * dt_ratio;
            const V2 P = i0 * un;
            va = va - ma * P;
            wa = wa - ia * cross(ra, P);
            vb = vb + mb * P;
            wb = wb + ib * cross(rb, P);
        } else i0 = real(0);
        imp[0] = i0;
    } else if (type == JOINT_WELD) {
        M33 m;
        const real c2x = -ra.y * ra.x * ia - rb.y * rb.x * ib;
        const real c3x = -ra.y * ia - rb.y * ib;
        m.c1 = mk3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, c2x, c3x);
        m.c2 = mk3(c2x, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, c3x);          // weld.py:148-150: c3x, as written
        m.c3 = mk3(c3x, ra.x * ia + rb.x * ib, ia + ib);
        m33_store(tmp + 4, m);
        real i0 = imp[0], i1 = imp[1], i2 = imp[2];
        if (warm) {
            i0 = dt_ratio * i0; i1 = dt_ratio * i1; i2 = dt_ratio * i2;
            const V2 P = mk<real>(i0, i1);
            va = va - ma * P;
            wa = wa - ia * (cross(ra, P) + i2);
            vb = vb + mb * P;
            wb = wb + ib * (cross(rb, P) + i2);
        } else i0 = i1 = i2 = real(0);
        imp[0] = i0; imp[1] = i1; imp[2] = i2;
    }
    else if (type == JOINT_PRISMATIC) {                                              // prismatic.py:232-335
        const V2 d = cb + rb - ca - ra;
        const V2 lx = mk<real>(par[9], par[10]), ly = cross_sv(real(1), lx);
        const V2 axis = rot_mul(qa, lx);
        const real a1 = cross(d + ra, axis), a2 = cross(rb, axis);
        real motor_mass = ma + mb + ia * sq_r(a1) + ib * sq_r(a2);
        if (motor_mass > real(0)) motor_mass = real(1) / motor_mass;
        const V2 perp = rot_mul(qa, ly);
        const real s1 = cross(d + ra, perp), s2 = cross(rb, perp);
        const real k11 = ma + mb + ia * sq_r(s1) + ib * sq_r(s2);
        const real k12 = ia * s1 + ib * s2;
        const real k13 = ia * s1 * a1 + ib * s2 * a2;
        real k22 = ia + ib;
        if (k22 == real(0)) k22 = real(1);
        const real k23 = ia * a1 + ib * a2;
        const real k33 = ma + mb + ia * sq_r(a1) + ib * sq_r(a2);
        tmp[4] = axis.x; tmp[5] = axis.y; tmp[6] = perp.x; tmp[7] = perp.y; tmp[8] = s1; tmp[9] = s2; tmp[10] = a1; tmp[11] = a2;
        tmp[12] = motor_mass; tmp[13] = k11; tmp[14] = k12; tmp[15] = k13; tmp[16] = k22; tmp[17] = k23; tmp[18] = k33;
        int ls = (flags >> JF_LS_SHIFT) & 3;
        real i0 = imp[0], i1 = imp[1], i2 = imp[2], im = imp[3];
        if (flags & JF_LIMIT) {
            const real joint_translation = dot(axis, d);
            if (abs_r(par[6] - par[5]) < (real(2) * K<real>::linear_slop)) ls = LS_EQUAL;
            else if (joint_translation <= par[5]) {
                if (ls != LS_LOWER) { ls = LS_LOWER; i2 = i2 - i2; }
            } else if (joint_translation >= par[6]) {
                if (ls != LS_UPPER) { ls = LS_UPPER; i2 = i2 - i2; }
            } else { ls = LS_INACTIVE; i2 = i2 - i2; }
        } else { ls = LS_INACTIVE; i2 = i2 - i2; }
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (!(flags & JF_MOTOR)) im = real(0);
        if (warm) {
            i0 = dt_ratio * i0; i1 = dt_ratio * i1; i2 = dt_ratio * i2;
            im = im * dt_ratio;
            const V2 P = i0 * perp + (im + i2) * axis;
            const real L1 = i0 * s1 + i1 + (im + i2) * a1;
            const real L2 = i0 * s2 + i1 + (im + i2) * a2;
            va = va - ma * P;
            wa = wa - ia * L1;
            vb = vb + mb * P;
            wb = wb + ib * L2;
        } else { i0 = i1 = i2 = real(0); im = real(0); }
        imp[0] = i0; imp[1] = i1; imp[2] = i2; imp[3] = im;
    }
    else if (type == JOINT_ROPE) {                                                   // rope.py:106-176
        const V2 u0 = cb + rb - ca - ra;                                             // self._u: never normalised (`u /= ...` rebinds the local)
        const real length = sqrt_r(sq_r(u0.x) + sq_r(u0.y));
        tmp[4] = u0.x; tmp[5] = u0.y; tmp[7] = length;
        const real C = length - par[4];
        const int ls = (C > real(0)) ? LS_UPPER : LS_INACTIVE;
        def[3] = (flags & ~(3 << JF_LS_SHIFT)) | (ls << JF_LS_SHIFT);
        if (!(length > K<real>::linear_slop)) { tmp[6] = real(0); imp[0] = real(0); return; }   // rope.py:147-151: returns before the velocities are stored
        const V2 u = mk<real>(u0.x / length, u0.y / length);
        const real cra = cross(ra, u), crb = cross(rb, u);
        const real inv_mass = ma + ia * cra * cra + mb + ib * crb * crb;
        tmp[6] = (inv_mass != real(0)) ? real(1) / inv_mass : real(0);
        real i0 = imp[0];
        if (warm) {
            i0 = i0 * dt_ratio;
            const V2 P = i0 * u0;
            va = va - ma * P;
            wa = wa - ia * cross(ra, P);
            vb = vb + mb * P;
            wb = wb + ib * cross(rb, P);
        } else i0 = real(0);
        imp[0] = i0;
    }
    else if (type == JOINT_FRICTION) {                                               // friction.py:105-176
        const real k1x = ma + mb + ia * ra.y * ra.y + ib * rb.y * rb.y;
        const real k1y = -ia * ra.x * ra.y - ib * rb.x * rb.y;
        const real k2x = k1y;
        const real k2y = ma + mb + ia * ra.x * ra.x + ib * rb.x * rb.x;
        real det = k1x * k2y - k2x * k1y;                                            // Mat22.inverse, common.py:506-519
        if (det != real(0)) det = real(1) / det;
        tmp[4] = det * k2y; tmp[5] = -det * k1y; tmp[6] = -det * k2x; tmp[7] = det * k1x;
        real angular_mass = ia + ib;
        if (angular_mass > real(0)) angular_mass = real(1) / angular_mass;
        tmp[8] = angular_mass;
        real i0 = imp[0], i1 = imp[1], i2 = imp[2];
        if (warm) {
            i0 = dt_ratio * i0; i1 = dt_ratio * i1;
            i2 = i2 * dt_ratio;
            const V2 P = mk<real>(i0, i1);
            va = va - ma * P;
            wa = wa - ia * (cross(ra, P) + i2);
            vb = vb + mb * P;
            wb = wb + ib * (cross(rb, P) + i2);
        } else { i0 = i1 = real(0); i2 = real(0); }
        imp[0] = i0; imp[1] = i1; imp[2] = i2;
    }
    else if (type == JOINT_WHEEL) {                                                  // wheel.py:202-318
        const V2 d = cb + rb - ca - ra;
        const V2 lx = mk<real>(par[4], par[5]), ly = cross_sv(real(1), lx);
        const V2 ay = rot_mul(qa, ly);
        const real s_ay = cross(d + ra, ay), s_by = cross(rb, ay);
        real mass = ma + mb + ia * sq_r(s_ay) + ib * sq_r(s_by);
        if (mass > real(0)) mass = real(1) / mass;
        real spring_mass = real(0), bias = real(0), gamma = real(0);
        V2 ax = mk<real>(real(0), real(0));                                          // (the attributes of a wheel without a spring stay at the constructor's zeros)
        real s_ax = real(0), s_bx = real(0);
        real i0 = imp[0], i1 = imp[1], i2 = imp[2];
        if (par[6] > real(0)) {
            ax = rot_mul(qa, lx);
            s_ax = cross(d + ra, ax); s_bx = cross(rb, ax);
            const real inv_mass = ma + mb + ia * sq_r(s_ax) + ib * sq_r(s_bx);
            if (inv_mass > real(0)) {
                spring_mass = real(1) / inv_mass;
                const real C = dot(d, ax);
                const real omega = real(2) * K<real>::pi * par[6];
                const real dd = real(2) * spring_mass * par[7] * omega;
                const real kk = spring_mass * sq_r(omega);
                gamma = dt * (dd + dt * kk);
                if (gamma > real(0)) gamma = real(1) / gamma;
                bias = C * dt * kk * gamma;
                spring_mass = inv_mass + gamma;
                if (spring_mass > real(0)) spring_mass = real(1) / spring_mass;
            }
        } else i1 = real(0);
        real motor_mass;
        if (flags & JF_MOTOR) {
            motor_mass = ia + ib;
            if (motor_mass > real(0)) motor_mass = real(1) / motor_mass;
        } else { motor_mass = real(0); i2 = real(0); }
        tmp[4] = ay.x; tmp[5] = ay.y; tmp[6] = s_ay; tmp[7] = s_by; tmp[8] = mass; tmp[9] = ax.x; tmp[10] = ax.y; tmp[11] = s_ax; tmp[12] = s_bx;
        tmp[13] = spring_mass; tmp[14] = bias; tmp[15] = gamma; tmp[16] = motor_mass;
        if (warm) {
            i0 = i0 * dt_ratio; i1 = i1 * dt_ratio; i2 = i2 * dt_ratio;
            const V2 P = i0 * ay + i1 * ax;
            const real La = i0 * s_ay + i1 * s_ax + i2;
            const real Lb = i0 * s_by + i1 * s_bx + i2;
            va = va - ma * P;
            wa = wa - ia * La;
            vb = vb + mb * P;
            wb = wb + ib * Lb;
        } else { i0 = real(0); i1 = real(0); i2 = real(0); }
        imp[0] = i0; imp[1] = i1; imp[2] = i2;
    }
    else if (type == JOINT_PULLEY) {                                                 // pulley.py:156-228
        V2 ua = ca + ra - mk<real>(par[4], par[5]), ub = cb + rb - mk<real>(par[6], par[7]);
        const real length_a = sqrt_r(sq_r(ua.x) + sq_r(ua.y)), length_b = sqrt_r(sq_r(ub.x) + sq_r(ub.y));
        ua = (length_a > real(10) * K<real>::linear_slop) ? mk<real>(ua.x / length_a, ua.y / length_a) : mk<real>(real(0), real(0));
        ub = (length_b > real(10) * K<real>::linear_slop) ? mk<real>(ub.x / length_b, ub.y / length_b) : mk<real>(real(0), real(0));
        tmp[4] = ua.x; tmp[5] = ua.y; tmp[6] = ub.x; tmp[7] = ub.y;
        const real rua = cross(ra, ua), rub = cross(rb, ub);
        const real ema = ma + ia * sq_r(rua), emb = mb + ib * sq_r(rub);             // the port rebinds `ma`, `mb`: its warm start uses these
        real mass = ema + par[8] * par[8] * emb;
        if (mass > real(0)) mass = real(1) / mass;
        tmp[8] = mass;
        real i0 = imp[0];
        if (warm) {
            i0 = i0 * dt_ratio;
            const V2 Pa = (-i0) * ua, Pb = (-par[8] * i0) * ub;
            va = va + ema * Pa;
            wa = wa + ia * cross(ra, Pa);
            vb = vb + emb * Pb;
            wb = wb + ib * cross(rb, Pb);
        } else i0 = real(0);
        imp[0] = i0;
    }
    state_set(ba, va.x, va.y, wa);
    state_set(bb, vb.x, vb.y, wb);
}

// _solve_velocity_constraints (revolute.py:302-378, distance.py:252-275, weld.py:171-199)
template <bool HOT> B2B_D void joint_solve_velocity_impl(jrow& r, real dt, real inv_dt)
{
    const int32_t* def = r.def;
    const real* par = r.par;
    real* imp = r.imp;
    const real* tmp = r.tmp;
    if (!HOT && def[0] == JOINT_MOUSE) { mouse_solve_velocity(r, dt); return; }
    const int type = HOT ? (int)JOINT_REVOLUTE : def[0];
    const int ba = def[1], bb = def[2], flags = def[3];
    const real ma = r.mqa.x, ia = r.mqa.y, mb = r.mqb.x, ib = r.mqb.y;
    const quad sa = state_get(ba), sb = state_get(bb);
    V2 va = mk<real>(sa.x, sa.y), vb = mk<real>(sb.x, sb.y);
    real wa = sa.z, wb = sb.z;
    const V2 ra = mk<real>(tmp[0], tmp[1]), rb = mk<real>(tmp[2], tmp[3]);
    if (type == JOINT_REVOLUTE) {
        const M33 m = m33_load(tmp + 4);
        const int ls = (flags >> JF_LS_SHIFT) & 3;
        const bool fixed_rotation = ((ia + ib) == real(0));
        real i0 = imp[0], i1 = imp[1], i2 = imp[2];
        if ((flags & JF_MOTOR) && ls != LS_EQUAL && !fixed_rotation) {
            const real Cdot = wb - wa - par[8];
            real impulse = tmp[13] * (-Cdot);
            const real old_impulse = imp[3];
            const real max_impulse = dt * par[7];
            const real im = clamp_r(old_impulse + impulse, -max_impulse, max_impulse);
            imp[3] = im;
            impulse = im - old_impulse;
            wa = wa - ia * impulse;
            wb = wb + ib * impulse;
        }
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE && !fixed_rotation) {
            const V2 Cdot1 = vb + cross_sv(wb, rb) - va - cross_sv(wa, ra);
            const real Cdot2 = wb - wa;
            V3 impulse = neg3(solve33(m, mk3(Cdot1.x, Cdot1.y, Cdot2)));
            if (ls == LS_EQUAL) {
                i0 = i0 + impulse.x; i1 = i1 + impulse.y; i2 = i2 + impulse.z;
            } else {
                const real new_impulse = i2 + impulse.z;
                if ((ls == LS_LOWER) ? (new_impulse < real(0)) : (new_impulse > real(0))) {
                    const V2 rhs = (-Cdot1) + i2 * mk<real>(m.c3.x, m.c3.y);
                    const V2 reduced = solve22(m, rhs);
                    impulse = mk3(reduced.x, reduced.y, -i2);
                    i0 = i0 + reduced.x; i1 = i1 + reduced.y; i2 = i2 + (-i2);
                } else {
                    i0 = i0 + impulse.x; i1 = i1 + impulse.y; i2 = i2 + impulse.z;
                }
            }
            const V2 P = mk<real>(impulse.x, impulse.y);
            va = va - ma * P;
            wa = wa - ia * (cross(ra, P) + impulse.z);
            vb = vb + mb * P;
            wb = wb + ib * (cross(rb, P) + impulse.z);
        } else {
            const V2 Cdot = vb + cross_sv(wb, rb) - va - cross_sv(wa, ra);
            const V2 impulse = solve22(m, -Cdot);
            i0 = i0 + impulse.x; i1 = i1 + impulse.y; i2 = i2 + real(0);
            va = va - ma * impulse;
            wa = wa - ia * cross(ra, impulse);
            vb = vb + mb * impulse;
            wb = wb + ib * cross(rb, impulse);
        }
        imp[0] = i0; imp[1] = i1; imp[2] = i2;
    } else if (type == JOINT_DISTANCE) {
        const V2 u = mk<real>(tmp[4], tmp[5]);
        const V2 vpa = va + cross_sv(wa, ra), vpb = vb + cross_sv(wb, rb);
        const real Cdot = dot(u, vpb - vpa);
        const real i0 = imp[0];
        const real impulse = -tmp[6] * (Cdot + tmp[8] + tmp[7] * i0);
        imp[0] = i0 + impulse;
        const V2 P = impulse * u;
        va = va - ma * P;
        wa = wa - ia * cross(ra, P);
        vb = vb + mb * P;
        wb = wb + ib * cross(rb, P);
    } else if (type == JOINT_WELD) {
        const M33 m = m33_load(tmp + 4);
        const V2 Cdot1 = vb + cross_sv(wb, rb) - va - cross_sv(wa, ra);
        const real Cdot2 = wb - wa;
        const V3 impulse = solve33(m, neg3(mk3(Cdot1.x, Cdot1.y, Cdot2)));
        imp[0] = imp[0] + impulse.x; imp[1] = imp[1] + impulse.y; imp[2] = imp[2] + impulse.z;
        const V2 P = mk<real>(impulse.x, impulse.y);
        va = va - ma * P;
        wa = wa - ia * (cross(ra, P) + impulse.z);
        vb = vb + mb * P;
        wb = wb + ib * (cross(rb, P) + impulse.z);
    }
    else if (type == JOINT_PRISMATIC) {                                              // prismatic.py:337-419
        const V2 axis = mk<real>(tmp[4], tmp[5]), perp = mk<real>(tmp[6], tmp[7]);
        const real s1 = tmp[8], s2 = tmp[9], a1 = tmp[10], a2 = tmp[11];
        const M33 m = m33_sym(tmp + 13);
        const int ls = (flags >> JF_LS_SHIFT) & 3;
        if ((flags & JF_MOTOR) && ls != LS_EQUAL) {
            const real Cdot = dot(axis, vb - va) + a2 * wb - a1 * wa;
            real impulse = tmp[12] * (par[8] - Cdot);
            const real old_impulse = imp[3];
            const real max_impulse = dt * par[7];
            const real im = clamp_r(old_impulse + impulse, -max_impulse, max_impulse);
            imp[3] = im;
            impulse = im - old_impulse;
            const V2 P = impulse * axis;
            const real L1 = impulse * a1, L2 = impulse * a2;
            va = va - ma * P;
            wa = wa - ia * L1;
            vb = vb + mb * P;
            wb = wb + ib * L2;
        }
        const V2 Cdot1 = mk<real>(dot(perp, vb - va) + s2 * wb - s1 * wa, wb - wa);
        V2 P;
        real L1, L2;
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE) {
            const real Cdot2 = dot(axis, vb - va) + a2 * wb - a1 * wa;
            const V3 old_impulse = mk3(imp[0], imp[1], imp[2]);
            V3 df = solve33(m, neg3(mk3(Cdot1.x, Cdot1.y, Cdot2)));
            const real cz = old_impulse.z + df.z;
            real iz;
            if (ls == LS_LOWER) iz = pymax(cz, real(0));
            else if (ls == LS_UPPER) iz = pymin(cz, real(0));
            else iz = cz;
            const V2 b = (-Cdot1) - (cz - old_impulse.z) * mk<real>(m.c3.x, m.c3.y);
            const V2 f2r = solve22(m, b) + mk<real>(old_impulse.x, old_impulse.y);
            imp[0] = f2r.x; imp[1] = f2r.y; imp[2] = iz;
            df = mk3(f2r.x - old_impulse.x, f2r.y - old_impulse.y, iz - old_impulse.z);
            P = df.x * perp + df.z * axis;
            L1 = df.x * s1 + df.y + df.z * a1;
            L2 = df.x * s2 + df.y + df.z * a2;
        } else {
            const V2 df = solve22(m, -Cdot1);
            imp[0] = imp[0] + df.x; imp[1] = imp[1] + df.y; imp[2] = imp[2] + real(0);
            P = df.x * perp;
            L1 = df.x * s1 + df.y;
            L2 = df.x * s2 + df.y;
        }
        va = va - ma * P;
        wa = wa - ia * L1;
        vb = vb + mb * P;
        wb = wb + ib * L2;
    }
    else if (type == JOINT_ROPE) {                                                   // rope.py:178-207
        const V2 u0 = mk<real>(tmp[4], tmp[5]);
        const V2 vpa = va + cross_sv(wa, ra), vpb = vb + cross_sv(wb, rb);
        const real C = tmp[7] - par[4];
        real Cdot = dot(u0, vpb - vpa);
        if (C < real(0)) Cdot = Cdot + inv_dt * C;
        real impulse = -tmp[6] * Cdot;
        const real old_impulse = imp[0];
        const real ni = pymin(real(0), old_impulse + impulse);
        imp[0] = ni;
        impulse = ni - old_impulse;
        const V2 P = impulse * u0;
        va = va - ma * P;
        wa = wa - ia * cross(ra, P);
        vb = vb + mb * P;
        wb = wb + ib * cross(rb, P);
    }
    else if (type == JOINT_FRICTION) {                                               // friction.py:178-230: the port stores the result into
        const real Cdot = wb - wa;                                                   // `positions`, i.e. nowhere: only the accumulated impulses change
        real impulse = -tmp[8] * Cdot;
        const real old_impulse = imp[2];
        real max_impulse = dt * par[5];
        const real na = clamp_r(old_impulse + impulse, -max_impulse, max_impulse);
        imp[2] = na;
        impulse = na - old_impulse;
        wa = wa - ia * impulse;
        wb = wb + ib * impulse;
        const V2 Cd = vb + cross_sv(wb, rb) - va - cross_sv(wa, ra);
        const V2 li = -mk<real>(tmp[4] * Cd.x + tmp[6] * Cd.y, tmp[5] * Cd.x + tmp[7] * Cd.y);
        V2 lin = mk<real>(imp[0], imp[1]) + li;
        max_impulse = dt * par[4];
        if (sq_r(lin.x) + sq_r(lin.y) > sq_r(max_impulse)) { const V2 n = normalize(lin); lin = mk<real>(max_impulse * n.x, max_impulse * n.y); }
        imp[0] = lin.x; imp[1] = lin.y;
        return;
    }
    else if (type == JOINT_WHEEL) {                                                  // wheel.py:320-373
        const V2 ay = mk<real>(tmp[4], tmp[5]), ax = mk<real>(tmp[9], tmp[10]);
        const real s_ay = tmp[6], s_by = tmp[7], s_ax = tmp[11], s_bx = tmp[12];
        real Cdot = dot(ax, vb - va) + s_bx * wb - s_ax * wa;
        real impulse = -tmp[13] * (Cdot + tmp[14] + tmp[15] * imp[1]);
        imp[1] = imp[1] + impulse;
        V2 P = impulse * ax;
        real La = impulse * s_ax, Lb = impulse * s_bx;
        va = va - ma * P;
        wa = wa - ia * La;
        vb = vb + mb * P;
        wb = wb + ib * Lb;
        Cdot = wb - wa - par[9];
        impulse = -tmp[16] * Cdot;
        const real old_impulse = imp[2];
        const real max_impulse = dt * par[8];
        const real nm = clamp_r(old_impulse + impulse, -max_impulse, max_impulse);
        imp[2] = nm;
        impulse = nm - old_impulse;
        wa = wa - ia * impulse;
        wb = wb + ib * impulse;
        Cdot = dot(ay, vb - va) + s_by * wb - s_ay * wa;
        impulse = -tmp[8] * Cdot;
        imp[0] = imp[0] + impulse;
        P = impulse * ay;
        La = impulse * s_ay; Lb = impulse * s_by;
        va = va - ma * P;
        wa = wa - ia * La;
        vb = vb + mb * P;
        wb = wb + ib * Lb;
    }
    else if (type == JOINT_PULLEY) {                                                 // pulley.py:230-256
        const V2 ua = mk<real>(tmp[4], tmp[5]), ub = mk<real>(tmp[6], tmp[7]);
        const V2 vpa = va + cross_sv(wa, ra), vpb = vb + cross_sv(wb, rb);
        const real Cdot = -dot(ua, vpa) - par[8] * dot(ub, vpb);
        const real impulse = -tmp[8] * Cdot;
        imp[0] = imp[0] + impulse;
        const V2 Pa = (-impulse) * ua, Pb = (-par[8] * impulse) * ub;
        va = va + ma * Pa;
        wa = wa + ia * cross(ra, Pa);
        vb = vb + mb * Pb;
        wb = wb + ib * cross(rb, Pb);
    }
    state_set(ba, va.x, va.y, wa);
    state_set(bb, vb.x, vb.y, wb);
}

// _solve_position_constraints (revolute.py:380-453, distance.py:277-311, weld.py:201-257) on the sweeps in the team's shared
// state; true when the joint's errors are within tolerance
template <bool HOT> B2B_D bool joint_solve_position_impl(const jrow& r)
{
    const int32_t* def = r.def;
    const real* par = r.par;
    const real* tmp = r.tmp;
    if (!HOT && def[0] == JOINT_MOUSE) return true;                                  // mouse.py:227-229
    const int type = HOT ? (int)JOINT_REVOLUTE : def[0];
    const int ba = def[1], bb = def[2], flags = def[3];
    const quad mqa = r.mqa, mqb = r.mqb;
    const real ma = mqa.x, ia = mqa.y, mb = mqb.x, ib = mqb.y;
    const quad sa = state_get(ba), sb = state_get(bb);
    V2 ca = mk<real>(sa.x, sa.y), cb = mk<real>(sb.x, sb.y);
    real aa = sa.z, ab = sb.z;
    bool ok = true;
    if (type == JOINT_REVOLUTE) {
        const int ls = (flags >> JF_LS_SHIFT) & 3;
        const bool fixed_rotation = ((ia + ib) == real(0));
        real angular_error = real(0), position_error = real(0);
        if ((flags & JF_LIMIT) && ls != LS_INACTIVE && !fixed_rotation) {
            const real angle = ab - aa - par[4];
            real limit_impulse = real(0);
            if (ls == LS_EQUAL) {
                const real C = clamp_r(angle - par[5], -K<real>::max_angular_correction, K<real>::max_angular_correction);
                limit_impulse = -tmp[13] * C;
                angular_error = abs_r(C);
            } else if (ls == LS_LOWER) {
                real C = angle - par[5];
                angular_error = -C;
                C = clamp_r(C + K<real>::angular_slop, -K<real>::max_angular_correction, real(0));
                limit_impulse = -tmp[13] * C;
            } else if (ls == LS_UPPER) {
                real C = angle - par[6];
                angular_error = C;
                C = clamp_r(C - K<real>::angular_slop, real(0), K<real>::max_angular_correction);
                limit_impulse = -tmp[13] * C;
            }
            aa = aa - ia * limit_impulse;
            ab = ab + ib * limit_impulse;
        }
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        const V2 C = cb + rb - ca - ra;
        position_error = sqrt_r(sq_r(C.x) + sq_r(C.y));
        const real k1x = ma + mb + ia * ra.y * ra.y + ib * rb.y * rb.y;
        const real k1y = -ia * ra.x * ra.y - ib * rb.x * rb.y;
        const real k2x = k1y;
        const real k2y = ma + mb + ia * ra.x * ra.x + ib * rb.x * rb.x;
        real det = k1x * k2y - k2x * k1y;                                            // Mat22.solve, common.py:521-531
        if (det != real(0)) det = real(1) / det;
        const V2 impulse = -mk<real>(det * (k2y * C.x - k2x * C.y), det * (k1x * C.y - k1y * C.x));
        ca = ca - ma * impulse;
        aa = aa - ia * cross(ra, impulse);
        cb = cb + mb * impulse;
        ab = ab + ib * cross(rb, impulse);
        ok = (position_error <= K<real>::linear_slop) && (angular_error <= K<real>::angular_slop);
    } else if (type == JOINT_DISTANCE) {
        if (par[5] > real(0)) return true;                                           // no position correction for soft constraints
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        V2 u = cb + rb - ca - ra;
        const real length = sqrt_r(sq_r(u.x) + sq_r(u.y));
        u = normalize(u);
        const real C = length - par[4];
        const real impulse = -tmp[6] * clamp_r(C, -K<real>::max_linear_correction, K<real>::max_linear_correction);
        const V2 P = impulse * u;
        ca = ca - ma * P;
        aa = aa - ia * cross(ra, P);
        cb = cb + mb * P;
        ab = ab + ib * cross(rb, P);
        ok = abs_r(C) < K<real>::linear_slop;
    } else if (type == JOINT_WELD) {
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        const V2 C1 = cb + rb - ca - ra;
        const real C2 = ab - aa - par[4];
        const real position_error = sqrt_r(sq_r(C1.x) + sq_r(C1.y)), angular_error = abs_r(C2);
        M33 m;
        const real c2x = -ra.y * ra.x * ia - rb.y * rb.x * ib;
        const real c3x = -ra.y * ia - rb.y * ib;
        const real c3y = ra.x * ia + rb.x * ib;
        m.c1 = mk3(ma + mb + ra.y * ra.y * ia + rb.y * rb.y * ib, c2x, c3x);
        m.c2 = mk3(c2x, ma + mb + ra.x * ra.x * ia + rb.x * rb.x * ib, c3y);
        m.c3 = mk3(-ra.y * ia - rb.y * ib, ra.x * ia + rb.x * ib, ia + ib);
        const V3 impulse = neg3(solve33(m, mk3(C1.x, C1.y, C2)));
        const V2 P = mk<real>(impulse.x, impulse.y);
        ca = ca - ma * P;
        aa = aa - ia * (cross(ra, P) + impulse.z);
        cb = cb + mb * P;
        ab = ab + ib * (cross(rb, P) + impulse.z);
        ok = (position_error <= K<real>::linear_slop) && (angular_error <= K<real>::angular_slop);
    }
    else if (type == JOINT_PRISMATIC) {                                              // prismatic.py:421-513
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        const V2 d = cb + rb - ca - ra;
        const V2 lx = mk<real>(par[9], par[10]), ly = cross_sv(real(1), lx);
        const V2 axis = rot_mul(qa, lx);
        const real a1 = cross(d + ra, axis), a2 = cross(rb, axis);
        const V2 perp = rot_mul(qa, ly);
        const real s1 = cross(d + ra, perp), s2 = cross(rb, perp);
        const V2 C1 = mk<real>(dot(perp, d), ab - aa - par[4]);
        real linear_error = abs_r(C1.x);
        const real angular_error = abs_r(C1.y);
        bool active = false;
        real C2 = real(0);
        if (flags & JF_LIMIT) {
            const real translation = dot(axis, d);
            if (abs_r(par[6] - par[5]) < real(2) * K<real>::linear_slop) {
                C2 = clamp_r(translation, -K<real>::max_linear_correction, K<real>::max_linear_correction);
                linear_error = pymax(linear_error, translation);
                active = true;
            } else if (translation <= par[5]) {
                C2 = clamp_r(translation - par[5] + K<real>::linear_slop, -K<real>::max_linear_correction, real(0));
                linear_error = pymax(linear_error, par[5] - translation);
                active = true;
            } else if (translation >= par[6]) {
                C2 = clamp_r(translation - par[6] - K<real>::linear_slop, real(0), K<real>::max_linear_correction);
                linear_error = pymax(linear_error, translation - par[6]);
                active = true;
            }
        }
        V3 impulse;
        if (active) {
            real k[6];
            k[0] = ma + mb + ia * sq_r(s1) + ib * sq_r(s2);
            k[1] = ia * s1 + ib * s2;
            k[2] = ia * s1 * a1 + ib * s2 * a2;
            k[3] = ia + ib;
            if (k[3] == real(0)) k[3] = real(1);
            k[4] = ia * a1 + ib * a2;
            k[5] = ma + mb + ia * sq_r(a1) + ib * sq_r(a2);
            impulse = solve33(m33_sym(k), neg3(mk3(C1.x, C1.y, C2)));
        } else {
            const real k11 = ma + mb + ia * s1 * s1 + ib * s2 * s2;
            const real k12 = ia * s1 + ib * s2;
            real k22 = ia + ib;
            if (k22 == real(0)) k22 = real(1);
            M33 m; m.c1 = mk3(k11, k12, real(0)); m.c2 = mk3(k12, k22, real(0)); m.c3 = mk3(real(0), real(0), real(0));
            const V2 impulse1 = solve22(m, -C1);
            impulse = mk3(impulse1.x, impulse1.y, real(0));
        }
        const V2 P = impulse.x * perp + impulse.z * axis;
        const real L1 = impulse.x * s1 + impulse.y + impulse.z * a1;
        const real L2 = impulse.x * s2 + impulse.y + impulse.z * a2;
        ca = ca - ma * P;
        cb = cb + mb * P;
        aa = aa - ia * L1;
        ab = ab + ib * L2;
        ok = (linear_error <= K<real>::linear_slop) && (angular_error <= K<real>::angular_slop);
    }
    else if (type == JOINT_ROPE) {                                                   // rope.py:209-248
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        V2 u = cb + rb - ca - ra;
        const real length = sqrt_r(sq_r(u.x) + sq_r(u.y));
        u = normalize(u);
        const real C = clamp_r(length - par[4], real(0), K<real>::max_linear_correction);
        const real impulse = -tmp[6] * C;
        const V2 P = impulse * u;
        ca = ca - ma * P;
        aa = aa - ia * cross(ra, P);
        cb = cb + mb * P;
        ab = ab + ib * cross(rb, P);
        ok = (length - par[4]) < K<real>::linear_slop;
    }
    else if (type == JOINT_FRICTION) return true;                                    // friction.py:232-233
    else if (type == JOINT_WHEEL) {                                                  // wheel.py:375-416
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        const V2 d = (cb - ca) + rb - ra;
        const real kk = ma + mb + ia * sq_r(tmp[6]) + ib * sq_r(tmp[7]);
        const V2 ly = cross_sv(real(1), mk<real>(par[4], par[5]));
        const V2 ay = rot_mul(qa, ly);
        const real s_ay = cross(d + ra, ay), s_by = cross(rb, ay);
        const real C = dot(d, ay);
        const real impulse = (kk != real(0)) ? -C / kk : real(0);
        const V2 P = impulse * ay;
        const real La = impulse * s_ay, Lb = impulse * s_by;
        ca = ca - ma * P;
        aa = aa - ia * La;
        cb = cb + mb * P;
        ab = ab + ib * Lb;
        ok = abs_r(C) <= K<real>::linear_slop;
    }
    else if (type == JOINT_PULLEY) {                                                 // pulley.py:258-318
        const XF qa = rot_of(aa), qb = rot_of(ab);
        const V2 ra = rot_mul(qa, mk<real>(tmp[24], tmp[25])), rb = rot_mul(qb, mk<real>(tmp[26], tmp[27]));
        V2 ua = ca + ra - mk<real>(par[4], par[5]), ub = cb + rb - mk<real>(par[6], par[7]);
        const real length_a = sqrt_r(sq_r(ua.x) + sq_r(ua.y)), length_b = sqrt_r(sq_r(ub.x) + sq_r(ub.y));
        ua = (length_a > real(10) * K<real>::linear_slop) ? mk<real>(ua.x / length_a, ua.y / length_a) : mk<real>(real(0), real(0));
        ub = (length_b > real(10) * K<real>::linear_slop) ? mk<real>(ub.x / length_b, ub.y / length_b) : mk<real>(real(0), real(0));
        const real rua = cross(ra, ua), rub = cross(rb, ub);
        const real ema = ma + ia * sq_r(rua), emb = mb + ib * sq_r(rub);
        real mass = ema + sq_r(par[8]) * emb;
        if (mass > real(0)) mass = real(1) / mass;
        const real C = par[9] - length_a - par[8] * length_b;
        const real linear_error = abs_r(C);
        const real impulse = -mass * C;
        const V2 Pa = (-impulse) * ua, Pb = (-par[8] * impulse) * ub;
        ca = ca + ema * Pa;
        aa = aa + ia * cross(ra, Pa);
        cb = cb + emb * Pb;
        ab = ab + ib * cross(rb, Pb);
        ok = linear_error < K<real>::linear_slop;
    }
    state_set(ba, ca.x, ca.y, aa);
    state_set(bb, cb.x, cb.y, ab);
    return ok;
}

B2B_D_NOINLINE void joint_init_cold(jrow& r, real dt, real dt_ratio, bool warm) { joint_init_impl<false>(r, dt, dt_ratio, warm); }
B2B_D_NOINLINE void joint_solve_velocity_cold(jrow& r, real dt, real inv_dt) { joint_solve_velocity_impl<false>(r, dt, inv_dt); }
B2B_D_NOINLINE bool joint_solve_position_cold(const jrow& r) { return joint_solve_position_impl<false>(r); }
#ifndef B2B_JOINT_COLD
#define B2B_JOINT_COLD 0         // experiments: 0 all nine types in line, 1 the out-of-line copy takes the caller's rows, 2 it takes a copy of them
#endif
#if B2B_JOINT_COLD == 0
B2B_D void joint_init(jrow& r, real dt, real dt_ratio, bool warm) { joint_init_impl<false>(r, dt, dt_ratio, warm); }
B2B_D void joint_solve_velocity(jrow& r, real dt, real inv_dt) { joint_solve_velocity_impl<false>(r, dt, inv_dt); }
B2B_D bool joint_solve_position(const jrow& r) { return joint_solve_position_impl<false>(r); }
#elif B2B_JOINT_COLD == 1
B2B_D void joint_init(jrow& r, real dt, real dt_ratio, bool warm)
{
    if (r.def[0] == JOINT_REVOLUTE) joint_init_impl<true>(r, dt, dt_ratio, warm); else joint_init_cold(r, dt, dt_ratio, warm);
}
B2B_D void joint_solve_velocity(jrow& r, real dt, real inv_dt)
{
    if (r.def[0] == JOINT_REVOLUTE) joint_solve_velocity_impl<true>(r, dt, inv_dt); else joint_solve_velocity_cold(r, dt, inv_dt);
}
B2B_D bool joint_solve_position(const jrow& r)
{
    return (r.def[0] == JOINT_REVOLUTE) ? joint_solve_position_impl<true>(r) : joint_solve_position_cold(r);
}
#else
B2B_D void joint_init(jrow& r, real dt, real dt_ratio, bool warm)
{
    // (the out-of-line copy gets its own jrow: handing out `r` itself would pin the caller's rows to local memory on the hot path too)
    if (r.def[0] == JOINT_REVOLUTE) joint_init_impl<true>(r, dt, dt_ratio, warm); else { jrow c = r; joint_init_cold(c, dt, dt_ratio, warm); r = c; }
}
B2B_D void joint_solve_velocity(jrow& r, real dt, real inv_dt)
{
    if (r.def[0] == JOINT_REVOLUTE) joint_solve_velocity_impl<true>(r, dt, inv_dt); else { jrow c = r; joint_solve_velocity_cold(c, dt, inv_dt); r = c; }
}
B2B_D bool joint_solve_position(const jrow& r)
{
    if (r.def[0] == JOINT_REVOLUTE) return joint_solve_position_impl<true>(r);
    jrow c = r;
    return joint_solve_position_cold(c);
}
#endif

// one island's joint range in l_joint
B2B_D void island_joint_range(int isl, int n_islands, int& j0, int& j1) const
{
    j0 = l_jstart[isl];
    j1 = (isl + 1 < n_islands) ? l_jstart[isl + 1] : w_joint[1];
}
