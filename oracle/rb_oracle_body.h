/* rb_oracle_body.h -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
 *
 * Plain-C restatement of the reference's impulse/friction hot path, included twice by
 * rb_oracle.c with REAL = double (suffix _f64) and REAL = float (suffix _f32).
 * Every function cites the reference lines it follows (paths relative to the reference root).
 * The operation ORDER of the reference's NumPy expressions is kept literally and the file is
 * compiled with -ffp-contract=off, so that in double precision the results agree with the
 * reference run under the fake MuJoCo backend to a few ulp (pinned by tests/golden/).
 *
 * MuJoCo pieces (third-party `mujoco` wheel, unpinned, absent here) follow SURVEY.md Appendix A.
 */

#define V3DOT(a, b) (((a)[0] * (b)[0] + (a)[1] * (b)[1]) + (a)[2] * (b)[2])

static void SUF(cross)(const REAL *a, const REAL *b, REAL *o) {
    /* numpy.cross: each component is (a_i*b_j) - (a_j*b_i), two roundings then a subtraction */
    REAL o0 = a[1] * b[2] - a[2] * b[1];
    REAL o1 = a[2] * b[0] - a[0] * b[2];
    REAL o2 = a[0] * b[1] - a[1] * b[0];
    o[0] = o0; o[1] = o1; o[2] = o2;
}

static void SUF(matvec3)(const REAL *A, const REAL *x, REAL *y) {
    REAL y0 = (A[0] * x[0] + A[1] * x[1]) + A[2] * x[2];
    REAL y1 = (A[3] * x[0] + A[4] * x[1]) + A[5] * x[2];
    REAL y2 = (A[6] * x[0] + A[7] * x[1]) + A[8] * x[2];
    y[0] = y0; y[1] = y1; y[2] = y2;
}

/* numpy.linalg.inv == LAPACK gesv(A, I): LU with partial pivoting, then forward/back solves. */
static void SUF(inv3)(const REAL *Ain, REAL *X) {
    REAL A[9];
    int piv[3] = {0, 1, 2};
    for (int i = 0; i < 9; ++i) A[i] = Ain[i];
    for (int k = 0; k < 3; ++k) {
        int p = k;
        REAL best = SUF(absr)(A[3 * k + k]);
        for (int i = k + 1; i < 3; ++i) {
            REAL v = SUF(absr)(A[3 * i + k]);
            if (v > best) { best = v; p = i; }
        }
        if (p != k) {
            for (int j = 0; j < 3; ++j) { REAL t = A[3 * k + j]; A[3 * k + j] = A[3 * p + j]; A[3 * p + j] = t; }
            int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
        }
        for (int i = k + 1; i < 3; ++i) {
            A[3 * i + k] = A[3 * i + k] / A[3 * k + k];
            for (int j = k + 1; j < 3; ++j) A[3 * i + j] = A[3 * i + j] - A[3 * i + k] * A[3 * k + j];
        }
    }
    for (int c = 0; c < 3; ++c) {
        REAL y[3];
        for (int i = 0; i < 3; ++i) {            /* L y = P e_c */
            REAL s = (piv[i] == c) ? (REAL)1 : (REAL)0;
            for (int j = 0; j < i; ++j) s = s - A[3 * i + j] * y[j];
            y[i] = s;
        }
        for (int i = 2; i >= 0; --i) {           /* U x = y */
            REAL s = y[i];
            for (int j = i + 1; j < 3; ++j) s = s - A[3 * i + j] * X[3 * j + c];
            X[3 * i + c] = s / A[3 * i + i];
        }
    }
}

/* A4  compute_inertia_tensor_world  src/physics/collision.py:51-53 (== time_integeration.py:8-10,
 * multi_sphere_bounce.py:35-37).  SciPy Rotation.from_quat normalises; as_matrix formula from
 * scipy/spatial/transform (x2 - y2 - z2 + w2, 2*(xy - zw), ...).  q is wxyz. */
static void SUF(rot_scipy)(const REAL *q, REAL *R) {
    REAL w = q[0], x = q[1], y = q[2], z = q[3];
    REAL nrm = SUF(sqrtr)(((x * x + y * y) + z * z) + w * w);
    x = x / nrm; y = y / nrm; z = z / nrm; w = w / nrm;
    REAL x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    REAL xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = ((x2 - y2) - z2) + w2;  R[1] = 2 * (xy - zw);          R[2] = 2 * (xz + yw);
    R[3] = 2 * (xy + zw);          R[4] = ((-x2 + y2) - z2) + w2; R[5] = 2 * (yz - xw);
    R[6] = 2 * (xz - yw);          R[7] = 2 * (yz + xw);          R[8] = ((-x2 - y2) + z2) + w2;
}

static void SUF(inertia_world_one)(const REAL *idiag, const REAL *q, REAL *Iw) {
    REAL R[9], M[9];
    SUF(rot_scipy)(q, R);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) M[3 * i + j] = R[3 * i + j] * idiag[j];       /* R @ diag(I) */
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)                                               /* ... @ R.T   */
            Iw[3 * i + j] = (M[3 * i] * R[3 * j] + M[3 * i + 1] * R[3 * j + 1]) + M[3 * i + 2] * R[3 * j + 2];
}

void SUF(rbo_inertia_world)(long n, const REAL *idiag3, const REAL *quat4, REAL *out9) {
    for (long e = 0; e < n; ++e) SUF(inertia_world_one)(idiag3 + 3 * e, quat4 + 4 * e, out9 + 9 * e);
}

/* A1  compute_collision_impulse_friction  src/physics/collision.py:7-48.
 * Returns 1 when an impulse was computed (u_n < 0), 0 on the early return (:32-33). */
static int SUF(impulse_friction_one)(REAL mass, const REAL *vel, const REAL *omega, const REAL *r,
                                     const REAL *n, REAL e, REAL mu, REAL *jn, REAL *jt) {
    REAL wxr[3], u[3], ut[3];
    SUF(cross)(omega, r, wxr);                                   /* :26 */
    for (int i = 0; i < 3; ++i) u[i] = vel[i] + wxr[i];
    REAL un = V3DOT(u, n);                                       /* :28 */
    for (int i = 0; i < 3; ++i) ut[i] = u[i] - un * n[i];        /* :29 */
    jt[0] = jt[1] = jt[2] = 0;
    if (un >= 0) { *jn = 0; return 0; }                          /* :32-33 */
    REAL k = ((REAL)1.0 / mass) + (REAL)(1.0 / 18);              /* :36 */
    *jn = (-(1 + e)) * un / k;                                   /* :39 */
    REAL tn = SUF(sqrtr)(V3DOT(ut, ut));                         /* :43 */
    if (tn > (REAL)1e-6) {
        REAL maxf = mu * SUF(absr)(*jn);                         /* :44 */
        REAL s = -(maxf < tn ? maxf : tn);                       /* :45 */
        for (int i = 0; i < 3; ++i) jt[i] = s * (ut[i] / tn);    /* :45-46 */
    }
    return 1;
}

void SUF(rbo_impulse_friction)(long n, const REAL *mass, const REAL *vel3, const REAL *omega3, const REAL *r3,
                               const REAL *normal3, const REAL *e, const REAL *mu, REAL *jn, REAL *jt3) {
    for (long i = 0; i < n; ++i)
        SUF(impulse_friction_one)(mass[i], vel3 + 3 * i, omega3 + 3 * i, r3 + 3 * i, normal3 + 3 * i, e[i], mu[i],
                                  jn + i, jt3 + 3 * i);
}

/* A2  apply_impulse_friction  src/physics/physics_utils.py:25-49 */
static void SUF(apply_impulse_friction_one)(REAL *vel, REAL *omega, REAL mass, const REAL *Iw, const REAL *r,
                                            const REAL *n, REAL jn, const REAL *jt) {
    REAL J[3], dv[3], rxJ[3], Iinv[9], dw[3];
    for (int i = 0; i < 3; ++i) J[i] = jn * n[i] + jt[i];        /* :42-43,45 */
    for (int i = 0; i < 3; ++i) dv[i] = J[i] / mass;             /* :45 */
    SUF(cross)(r, J, rxJ);
    SUF(inv3)(Iw, Iinv);                                         /* :46 */
    SUF(matvec3)(Iinv, rxJ, dw);
    for (int i = 0; i < 3; ++i) { vel[i] = vel[i] + dv[i]; omega[i] = omega[i] + dw[i]; }   /* :49 */
}

void SUF(rbo_apply_impulse_friction)(long n, const REAL *vel3, const REAL *omega3, const REAL *mass, const REAL *Iw9,
                                     const REAL *r3, const REAL *normal3, const REAL *jn, const REAL *jt3,
                                     REAL *vel_out, REAL *omega_out) {
    for (long i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) { vel_out[3 * i + c] = vel3[3 * i + c]; omega_out[3 * i + c] = omega3[3 * i + c]; }
        SUF(apply_impulse_friction_one)(vel_out + 3 * i, omega_out + 3 * i, mass[i], Iw9 + 9 * i, r3 + 3 * i,
                                        normal3 + 3 * i, jn[i], jt3 + 3 * i);
    }
}

/* A3  apply_impulse  src/physics/physics_utils.py:4-22 */
void SUF(rbo_apply_impulse)(long n, const REAL *vel3, const REAL *omega3, const REAL *mass, const REAL *Iw9,
                            const REAL *r3, const REAL *normal3, const REAL *impulse, REAL *vel_out, REAL *omega_out) {
    for (long i = 0; i < n; ++i) {
        REAL Jn[3], rxJ[3], Iinv[9], dw[3];
        REAL s = impulse[i] / mass[i];                                               /* :19 */
        for (int c = 0; c < 3; ++c) Jn[c] = impulse[i] * normal3[3 * i + c];         /* :21 */
        SUF(cross)(r3 + 3 * i, Jn, rxJ);
        SUF(inv3)(Iw9 + 9 * i, Iinv);
        SUF(matvec3)(Iinv, rxJ, dw);
        for (int c = 0; c < 3; ++c) {
            vel_out[3 * i + c] = vel3[3 * i + c] + s * normal3[3 * i + c];
            omega_out[3 * i + c] = omega3[3 * i + c] + dw[c];
        }
    }
}

/* mju_mulQuat: Hamilton product, wxyz */
static void SUF(mulquat)(const REAL *a, const REAL *b, REAL *res) {
    res[0] = ((a[0] * b[0] - a[1] * b[1]) - a[2] * b[2]) - a[3] * b[3];
    res[1] = ((a[0] * b[1] + a[1] * b[0]) + a[2] * b[3]) - a[3] * b[2];
    res[2] = ((a[0] * b[2] - a[1] * b[3]) + a[2] * b[0]) + a[3] * b[1];
    res[3] = ((a[0] * b[3] + a[1] * b[2]) - a[2] * b[1]) + a[3] * b[0];
}

/* MuJoCo mju_quat2Mat on the NORMALISED joint quaternion (mj_kinematics), Appendix A.1 */
static void SUF(rot_mujoco)(const REAL *qin, REAL *R) {
    REAL nrm = SUF(sqrtr)(((qin[0] * qin[0] + qin[1] * qin[1]) + qin[2] * qin[2]) + qin[3] * qin[3]);
    REAL w = qin[0] / nrm, x = qin[1] / nrm, y = qin[2] / nrm, z = qin[3] / nrm;
    R[0] = ((w * w + x * x) - y * y) - z * z; R[1] = 2 * (x * y - w * z);               R[2] = 2 * (x * z + w * y);
    R[3] = 2 * (x * y + w * z);               R[4] = ((w * w - x * x) + y * y) - z * z; R[5] = 2 * (y * z - w * x);
    R[6] = 2 * (x * z - w * y);               R[7] = 2 * (y * z + w * x);               R[8] = ((w * w - x * x) - y * y) + z * z;
}

typedef struct { REAL dist; REAL pos[3]; REAL n[3]; } SUF(contact_t);

/* Appendix A.2 plane(g1)-sphere(g2) */
static int SUF(plane_sphere)(const REAL *pp, const REAL *n, const REAL *c, REAL rad, SUF(contact_t) *out) {
    REAL d[3] = {c[0] - pp[0], c[1] - pp[1], c[2] - pp[2]};
    REAL dist = V3DOT(d, n) - rad;
    if (dist > 0) return 0;
    REAL s = rad + (REAL)0.5 * dist;
    for (int i = 0; i < 3; ++i) { out->pos[i] = c[i] - n[i] * s; out->n[i] = n[i]; }
    out->dist = dist;
    return 1;
}

/* Appendix A.2 plane(g1)-box(g2): vertices in index order, bit0->x bit1->y bit2->z, stop after 4 */
static int SUF(plane_box)(const REAL *pp, const REAL *n, const REAL *c, const REAL *q, const REAL *half,
                          SUF(contact_t) *out) {
    REAL R[9];
    SUF(rot_mujoco)(q, R);
    REAL d[3] = {c[0] - pp[0], c[1] - pp[1], c[2] - pp[2]};
    REAL d0 = V3DOT(d, n);
    int cnt = 0;
    for (int i = 0; i < 8 && cnt < 4; ++i) {
        REAL v[3] = {(i & 1) ? half[0] : -half[0], (i & 2) ? half[1] : -half[1], (i & 4) ? half[2] : -half[2]};
        REAL corner[3];
        SUF(matvec3)(R, v, corner);
        REAL ld = V3DOT(n, corner);
        if (d0 + ld > 0 || ld > 0) continue;
        REAL dist = d0 + ld;
        REAL hs = (REAL)0.5 * dist;
        for (int k = 0; k < 3; ++k) { out[cnt].pos[k] = (c[k] + corner[k]) - n[k] * hs; out[cnt].n[k] = n[k]; }
        out[cnt].dist = dist;
        ++cnt;
    }
    return cnt;
}

/* Appendix A.2 sphere(g1)-sphere(g2), g1 the lower id */
static int SUF(sphere_sphere)(const REAL *c1, REAL r1, const REAL *c2, REAL r2, SUF(contact_t) *out) {
    REAL d[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
    REAL L = SUF(sqrtr)((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
    REAL dist = (L - r1) - r2;
    if (dist > 0) return 0;
    REAL n[3] = {1, 0, 0};
    if (L >= (REAL)1e-15) { n[0] = d[0] / L; n[1] = d[1] / L; n[2] = d[2] / L; }
    REAL s = r1 + (REAL)0.5 * dist;
    for (int i = 0; i < 3; ++i) { out->pos[i] = c1[i] + n[i] * s; out->n[i] = n[i]; }
    out->dist = dist;
    return 1;
}

/* velocity update shared by A5/A6/A7/A9: collision.py:66-70 */
static void SUF(external_forces)(REAL *vel, REAL *omega, REAL mass, const REAL *Iw, const REAL *g, const REAL *xfrc,
                                 REAL dt) {
    REAL Iinv[9], tdt[3], dw[3];
    for (int i = 0; i < 3; ++i) {
        REAL force = (xfrc ? xfrc[i] : (REAL)0) + mass * g[i];   /* :66 */
        vel[i] = vel[i] + (force / mass) * dt;                   /* :69 */
        tdt[i] = (xfrc ? xfrc[3 + i] : (REAL)0) * dt;            /* :67,70 */
    }
    SUF(inv3)(Iw, Iinv);
    SUF(matvec3)(Iinv, tdt, dw);
    for (int i = 0; i < 3; ++i) omega[i] = omega[i] + dw[i];     /* :70 */
}

/* quaternion integrate collision.py:91-95 */
static void SUF(integrate_quat)(REAL *q, const REAL *omega, REAL dt) {
    REAL oq[4] = {0, omega[0], omega[1], omega[2]}, res[4], qn[4];
    SUF(mulquat)(oq, q, res);                                             /* :93 */
    for (int i = 0; i < 4; ++i) qn[i] = q[i] + ((REAL)0.5 * res[i]) * dt; /* :94 */
    REAL nrm = SUF(sqrtr)(((qn[0] * qn[0] + qn[1] * qn[1]) + qn[2] * qn[2]) + qn[3] * qn[3]);
    for (int i = 0; i < 4; ++i) q[i] = qn[i] / nrm;                       /* :95 */
}

/* A5 / A6 (scheme 0)  custom_step_with_impulse_collision_friction  collision.py:56-102 ==
 *                     timestep_integration  time_integeration.py:13-72
 * A7      (scheme 1)  general  time_integeration.py:75-141
 * geom: 0 = sphere (size[0] = radius), 1 = box (size = half extents).
 * Layout is the reference's: qpos[E][7] (xyz, wxyz), qvel[E][6].  calls/impulses may be NULL. */
void SUF(rbo_step_body_plane)(long E, int steps, int scheme, int geom, REAL *qpos7, REAL *qvel6, const REAL *mass,
                              const REAL *inertia3, const REAL *size3, const REAL *plane_pos3,
                              const REAL *plane_normal3, const REAL *gravity3, const REAL *xfrc6, REAL dt,
                              const REAL *rest, const REAL *fric, REAL thr, unsigned *calls, unsigned *impulses) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long e = 0; e < E; ++e) {
        REAL *qp = qpos7 + 7 * e, *qv = qvel6 + 6 * e;
        unsigned nc = 0, ni = 0;
        for (int s = 0; s < steps; ++s) {
            SUF(contact_t) con[4];
            int ncon = geom == 0 ? SUF(plane_sphere)(plane_pos3, plane_normal3, qp, size3[3 * e], con)
                                 : SUF(plane_box)(plane_pos3, plane_normal3, qp, qp + 3, size3 + 3 * e, con);
            REAL Iw[9];
            SUF(inertia_world_one)(inertia3 + 3 * e, qp + 3, Iw);             /* :62 */
            REAL vel[3] = {qv[0], qv[1], qv[2]}, om[3] = {qv[3], qv[4], qv[5]};
            REAL ppred[3] = {qp[0] + vel[0] * dt, qp[1] + vel[1] * dt, qp[2] + vel[2] * dt}; /* general :106 */
            SUF(external_forces)(vel, om, mass[e], Iw, gravity3, xfrc6 ? xfrc6 + 6 * e : 0, dt);
            for (int i = 0; i < ncon; ++i) {                                  /* :72 */
                if (!(con[i].dist == con[i].dist) || !(con[i].dist < 0)) continue;     /* :74 */
                REAL r[3] = {con[i].pos[0] - qp[0], con[i].pos[1] - qp[1], con[i].pos[2] - qp[2]};  /* :75 */
                if (SUF(absr)(con[i].dist) < thr) continue;                   /* :79-80 */
                REAL jn, jt[3];
                ++nc;
                ni += SUF(impulse_friction_one)(mass[e], vel, om, r, con[i].n, rest[e], fric[e], &jn, jt);
                SUF(apply_impulse_friction_one)(vel, om, mass[e], Iw, r, con[i].n, jn, jt);  /* :86 */
            }
            if (scheme == 0) {
                for (int i = 0; i < 3; ++i) qp[i] = qp[i] + vel[i] * dt;      /* :90 */
                SUF(integrate_quat)(qp + 3, om, dt);                          /* :91-95 */
            } else {
                for (int i = 0; i < 3; ++i) qp[i] = ppred[i];                 /* general :134-137 */
            }
            for (int i = 0; i < 3; ++i) { qv[i] = vel[i]; qv[3 + i] = om[i]; }
        }
        if (calls) calls[e] += nc;
        if (impulses) impulses[e] += ni;
    }
}

/* A9  custom_step_multi_sphere  src/simulation/multi_sphere_bounce.py:42-92 with the repairs of
 * ball b <-> slices 7b / 6b; a contact belongs to ball b iff one of its
 * geoms is ball b's; contacts visited in MuJoCo order (ground, then partners by ascending index);
 * the normal is used as generated (geom1 -> geom2), never flipped; no isnan test, no threshold.
 * Layout: qpos[E][B][7], qvel[E][B][6]; mass/radius/inertia per body [E][B](,3). */
void SUF(rbo_step_multi_sphere)(long E, int B, int steps, REAL *qpos, REAL *qvel, const REAL *mass,
                                const REAL *inertia3, const REAL *radius, const REAL *plane_pos3,
                                const REAL *plane_normal3, const REAL *gravity3, REAL dt, REAL rest, REAL fric,
                                unsigned *calls, unsigned *impulses) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long e = 0; e < E; ++e) {
        REAL *QP = qpos + (long)7 * B * e, *QV = qvel + (long)6 * B * e;
        REAL *p0 = (REAL *)malloc(sizeof(REAL) * 3 * B);
        for (int s = 0; s < steps; ++s) {
            for (int b = 0; b < B; ++b)                      /* mj_forward once per step (:43) */
                for (int i = 0; i < 3; ++i) p0[3 * b + i] = QP[7 * b + i];
            for (int b = 0; b < B; ++b) {                    /* :46 */
                long eb = e * B + b;
                REAL *qp = QP + 7 * b, *qv = QV + 6 * b;
                REAL Iw[9];
                SUF(inertia_world_one)(inertia3 + 3 * eb, qp + 3, Iw);         /* :55 */
                REAL vel[3] = {qv[0], qv[1], qv[2]}, om[3] = {qv[3], qv[4], qv[5]};
                SUF(external_forces)(vel, om, mass[eb], Iw, gravity3, 0, dt);  /* :58-61 */
                for (int j = -1; j < B; ++j) {               /* :64, MuJoCo contact order for ball b */
                    SUF(contact_t) c;
                    int hit;
                    if (j == b) continue;
                    if (j < 0) hit = SUF(plane_sphere)(plane_pos3, plane_normal3, p0 + 3 * b, radius[eb], &c);
                    else if (j < b) hit = SUF(sphere_sphere)(p0 + 3 * j, radius[e * B + j], p0 + 3 * b, radius[eb], &c);
                    else hit = SUF(sphere_sphere)(p0 + 3 * b, radius[eb], p0 + 3 * j, radius[e * B + j], &c);
                    if (!hit || !(c.dist < 0)) continue;     /* :66 */
                    REAL r[3] = {c.pos[0] - qp[0], c.pos[1] - qp[1], c.pos[2] - qp[2]};    /* :67 */
                    REAL jn, jt[3];
                    int imp = SUF(impulse_friction_one)(mass[eb], vel, om, r, c.n, rest, fric, &jn, jt);  /* :69 */
                    if (calls) calls[eb] += 1;
                    if (impulses) impulses[eb] += (unsigned)imp;
                    SUF(apply_impulse_friction_one)(vel, om, mass[eb], Iw, r, c.n, jn, jt);               /* :72 */
                }
                for (int i = 0; i < 3; ++i) qp[i] = qp[i] + vel[i] * dt;       /* :77 */
                SUF(integrate_quat)(qp + 3, om, dt);                           /* :78-82 */
                for (int i = 0; i < 3; ++i) { qv[i] = vel[i]; qv[3 + i] = om[i]; }  /* :85-88 */
            }
        }
        free(p0);
    }
}

/* ------------------------------------------------------------------------------------------------
 * N4: multi-body scenes with spheres AND boxes.  NOT in the reference (its scripts only ever
 * meet plane-sphere, plane-box and sphere-sphere): the STEP is the repaired A9 loop above, body by body, with the
 * reference's A1 / A2 / A4 per contact; the CONTACT SET is widened by the two pair functions below, which are this
 * project's own specification (DESIGN.md "N4"), written in the conventions of Appendix A.2 -- contact = {dist, pos midway
 * between the two surfaces, normal from geom1 to geom2}.  Parity of these two against MuJoCo's mjc_SphereBox /
 * mjc_BoxBox is NOT claimed (box-box here: vertices inside the other box, then edges passing through it; no SAT, no
 * clipped face polygons).
 * Geoms may sit at an offset in their body's frame (gpos, gquat: N1 remainder); the impulse arm is still taken from the
 * body origin qpos[:3], as the reference does (collision.py:75).
 * ------------------------------------------------------------------------------------------------ */
static void SUF(plane_box_rot)(const REAL *pp, const REAL *n, const REAL *c, const REAL *R, const REAL *half,
                               SUF(contact_t) *out, int *count) {
    REAL d[3] = {c[0] - pp[0], c[1] - pp[1], c[2] - pp[2]};
    REAL d0 = V3DOT(d, n);
    int cnt = 0;
    for (int i = 0; i < 8 && cnt < 4; ++i) {
        REAL v[3] = {(i & 1) ? half[0] : -half[0], (i & 2) ? half[1] : -half[1], (i & 4) ? half[2] : -half[2]};
        REAL corner[3];
        SUF(matvec3)(R, v, corner);
        REAL ld = V3DOT(n, corner);
        if (d0 + ld > 0 || ld > 0) continue;
        REAL dist = d0 + ld;
        REAL hs = (REAL)0.5 * dist;
        for (int k = 0; k < 3; ++k) { out[cnt].pos[k] = (c[k] + corner[k]) - n[k] * hs; out[cnt].n[k] = n[k]; }
        out[cnt].dist = dist;
        ++cnt;
    }
    *count = cnt;
}

/* x in the frame of a box at c with rotation R (columns = box axes): R^T (x - c) */
static void SUF(to_box_frame)(const REAL *c, const REAL *R, const REAL *x, REAL *o) {
    REAL d[3] = {x[0] - c[0], x[1] - c[1], x[2] - c[2]};
    for (int k = 0; k < 3; ++k) o[k] = (R[k] * d[0] + R[3 + k] * d[1]) + R[6 + k] * d[2];
}

/* sphere - box.  The closest point of the box to the sphere centre is the centre clamped to the box in the box frame; a
 * centre inside the box leaves through the nearest face.  geom1 is the body with the LOWER INDEX, as for sphere pairs
 * (MuJoCo itself would put the lower geom TYPE first; with the A9 loop's never-flipped normal that would make every
 * sphere blind to every box, so the index rule is kept for all pair types): sign = +1 when the sphere is geom1 (normal
 * sphere -> box), -1 when the box is. */
static int SUF(sphere_box)(const REAL *cs, REAL rad, const REAL *cb, const REAL *Rb, const REAL *h, REAL sign,
                           SUF(contact_t) *out) {
    REAL c[3], cl[3], e[3], nl[3], pl[3], dist;
    SUF(to_box_frame)(cb, Rb, cs, c);
    for (int k = 0; k < 3; ++k) { cl[k] = c[k] < -h[k] ? -h[k] : (c[k] > h[k] ? h[k] : c[k]); e[k] = cl[k] - c[k]; }
    REAL L = SUF(sqrtr)((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]);
    if (L >= (REAL)1e-15) {
        dist = L - rad;
        if (dist > 0) return 0;
        REAL s = rad + (REAL)0.5 * dist;
        for (int k = 0; k < 3; ++k) { nl[k] = e[k] / L; pl[k] = c[k] + nl[k] * s; }
    } else {
        int ax = 0;
        REAL depth = h[0] - SUF(absr)(c[0]);
        for (int k = 1; k < 3; ++k) { REAL dk = h[k] - SUF(absr)(c[k]); if (dk < depth) { depth = dk; ax = k; } }
        dist = -(rad + depth);
        REAL s = (REAL)0.5 * (rad - depth);
        for (int k = 0; k < 3; ++k) { nl[k] = 0; pl[k] = c[k]; }
        nl[ax] = c[ax] >= 0 ? (REAL)-1 : (REAL)1;
        pl[ax] = c[ax] + nl[ax] * s;
    }
    REAL nw[3], pw[3];
    SUF(matvec3)(Rb, nl, nw);
    SUF(matvec3)(Rb, pl, pw);
    for (int k = 0; k < 3; ++k) { out->n[k] = sign * nw[k]; out->pos[k] = cb[k] + pw[k]; }
    out->dist = dist;
    return 1;
}

/* Features of box V (centre cv, rotation Rv, half extents hv) inside box F: first its vertices (index order, bit0->x
 * bit1->y bit2->z) that lie inside F, then its edges that pass through F without either end point inside it and without
 * both end points beyond one face of F -- the edge is clipped against F's three slabs in F's frame and the middle of the
 * clipped piece is taken (covers edge-edge crossings and an edge lying across a face: crossed planks).  Every such point
 * of V leaves F through F's nearest face.  sign = +1 when F is geom1 (the face normal already points geom1 -> geom2),
 * -1 when F is geom2.  Edge e = 4*a + c: along axis a of V, c = the signs of the other two axes (lower axis in bit 0).
 * Appends to out[], at most `room` contacts. */
static int SUF(box_features_in_box)(const REAL *cv, const REAL *Rv, const REAL *hv, const REAL *cf, const REAL *Rf,
                                    const REAL *hf, REAL sign, SUF(contact_t) *out, int room, int phase, REAL (*l)[3],
                                    int *inside_io) {
    /* phase 0: the vertex pass (fills l[8][3] = V's vertices in F's frame and the mask of those inside F);
     * phase 1: the edge pass (reads both) */
    int cnt = 0, inside = phase ? *inside_io : 0;
    for (int i = 0; i < 8 && phase == 0; ++i) {
        REAL v[3] = {(i & 1) ? hv[0] : -hv[0], (i & 2) ? hv[1] : -hv[1], (i & 4) ? hv[2] : -hv[2]};
        REAL corner[3], x[3];
        SUF(matvec3)(Rv, v, corner);
        for (int k = 0; k < 3; ++k) x[k] = cv[k] + corner[k];
        SUF(to_box_frame)(cf, Rf, x, l[i]);
        int ax = 0;
        REAL depth = hf[0] - SUF(absr)(l[i][0]);
        for (int k = 1; k < 3; ++k) { REAL dk = hf[k] - SUF(absr)(l[i][k]); if (dk < depth) { depth = dk; ax = k; } }
        if (!(depth > 0)) continue;                       /* outside F (or on its surface) */
        inside |= 1 << i;
        if (cnt >= room) continue;
        REAL sg = l[i][ax] >= 0 ? (REAL)1 : (REAL)-1;     /* outward normal of the nearest face: sg * axis ax of F */
        REAL m[3] = {sg * Rf[ax], sg * Rf[3 + ax], sg * Rf[6 + ax]};
        REAL hd = (REAL)0.5 * depth;
        for (int k = 0; k < 3; ++k) { out[cnt].pos[k] = x[k] + m[k] * hd; out[cnt].n[k] = sign * m[k]; }
        out[cnt].dist = -depth;
        ++cnt;
    }
    if (phase == 0) { *inside_io = inside; return cnt; }
    for (int e = 0; e < 12 && cnt < room; ++e) {
        int a = e >> 2, b1 = a == 0 ? 1 : 0, b2 = a == 2 ? 1 : 2;
        int i0 = ((e & 1) ? 1 << b1 : 0) | ((e & 2) ? 1 << b2 : 0), i1 = i0 | (1 << a);
        if (inside & ((1 << i0) | (1 << i1))) continue;   /* an end point inside F: a vertex contact already */
        const REAL *p0 = l[i0], *p1 = l[i1];
        int beyond = 0;
        for (int k = 0; k < 3; ++k)
            if ((p0[k] >= hf[k] && p1[k] >= hf[k]) || (p0[k] <= -hf[k] && p1[k] <= -hf[k])) beyond = 1;
        if (beyond) continue;                             /* both ends beyond one face: the edge cannot enter F */
        REAL t0 = 0, t1 = 1, d[3];
        int empty = 0;
        for (int k = 0; k < 3; ++k) {
            d[k] = p1[k] - p0[k];
            if (d[k] == 0) { if (!(hf[k] - SUF(absr)(p0[k]) > 0)) empty = 1; continue; }
            REAL ta = (-hf[k] - p0[k]) / d[k], tb = (hf[k] - p0[k]) / d[k];
            REAL lo = ta < tb ? ta : tb, hi = ta < tb ? tb : ta;
            if (lo > t0) t0 = lo;
            if (hi < t1) t1 = hi;
        }
        if (empty || !(t0 < t1)) continue;
        REAL tm = (REAL)0.5 * (t0 + t1), lm[3];
        for (int k = 0; k < 3; ++k) lm[k] = p0[k] + d[k] * tm;
        int ax = 0;
        REAL depth = hf[0] - SUF(absr)(lm[0]);
        for (int k = 1; k < 3; ++k) { REAL dk = hf[k] - SUF(absr)(lm[k]); if (dk < depth) { depth = dk; ax = k; } }
        if (!(depth > 0)) continue;
        REAL v0[3] = {(i0 & 1) ? hv[0] : -hv[0], (i0 & 2) ? hv[1] : -hv[1], (i0 & 4) ? hv[2] : -hv[2]};
        REAL v1[3] = {(i1 & 1) ? hv[0] : -hv[0], (i1 & 2) ? hv[1] : -hv[1], (i1 & 4) ? hv[2] : -hv[2]};
        REAL w0[3], w1[3], x[3];
        SUF(matvec3)(Rv, v0, w0);
        SUF(matvec3)(Rv, v1, w1);
        for (int k = 0; k < 3; ++k) { REAL x0 = cv[k] + w0[k], x1 = cv[k] + w1[k]; x[k] = x0 + (x1 - x0) * tm; }
        REAL sg = lm[ax] >= 0 ? (REAL)1 : (REAL)-1;
        REAL m[3] = {sg * Rf[ax], sg * Rf[3 + ax], sg * Rf[6 + ax]};
        REAL hd = (REAL)0.5 * depth;
        for (int k = 0; k < 3; ++k) { out[cnt].pos[k] = x[k] + m[k] * hd; out[cnt].n[k] = sign * m[k]; }
        out[cnt].dist = -depth;
        ++cnt;
    }
    return cnt;
}

/* box (geom1 = lower id) - box (geom2): nothing when a face normal of either box separates them; else the vertices of
 * geom2 inside geom1, the vertices of geom1 inside geom2 (<= 8), and, while the pair has fewer than four contacts, the
 * edges of geom2 through geom1 and then the edges of geom1 through geom2 */
static int SUF(box_box)(const REAL *c1, const REAL *R1, const REAL *h1, const REAL *c2, const REAL *R2, const REAL *h2,
                        SUF(contact_t) *out) {
    /* separating-axis test on the six face normals first: boxes apart along one of them have no contact */
    REAL t[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]}, C[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C[3 * i + j] = SUF(absr)((R1[i] * R2[j] + R1[3 + i] * R2[3 + j]) + R1[6 + i] * R2[6 + j]);
    for (int i = 0; i < 3; ++i) {
        REAL d1 = SUF(absr)((t[0] * R1[i] + t[1] * R1[3 + i]) + t[2] * R1[6 + i]);
        REAL d2 = SUF(absr)((t[0] * R2[i] + t[1] * R2[3 + i]) + t[2] * R2[6 + i]);
        if (d1 > h1[i] + ((h2[0] * C[3 * i] + h2[1] * C[3 * i + 1]) + h2[2] * C[3 * i + 2])) return 0;
        if (d2 > h2[i] + ((h1[0] * C[i] + h1[1] * C[3 + i]) + h1[2] * C[6 + i])) return 0;
    }
    REAL l21[8][3], l12[8][3];
    int in21 = 0, in12 = 0;
    int cnt = SUF(box_features_in_box)(c2, R2, h2, c1, R1, h1, (REAL)1, out, 8, 0, l21, &in21);
    cnt += SUF(box_features_in_box)(c1, R1, h1, c2, R2, h2, (REAL)-1, out + cnt, 8 - cnt, 0, l12, &in12);
    /* edge contacts only fill a pair up to four contacts: a pair that already rests on three or four vertices has its
     * support, and most touching pairs of a settled pile are of that kind */
    int room = cnt < 4 ? 4 - cnt : 0;
    int ne = SUF(box_features_in_box)(c2, R2, h2, c1, R1, h1, (REAL)1, out + cnt, room, 1, l21, &in21);
    ne += SUF(box_features_in_box)(c1, R1, h1, c2, R2, h2, (REAL)-1, out + cnt + ne, room - ne, 1, l12, &in12);
    return cnt + ne;
}

/* world pose of a body's geom: centre and rotation matrix (mju_quat2Mat of the normalised quaternion) */
static void SUF(geom_pose)(const REAL *p, const REAL *q, int has_off, const REAL *gpos, const REAL *gquat, REAL *c, REAL *R) {
    if (!has_off) {
        for (int k = 0; k < 3; ++k) c[k] = p[k];
        SUF(rot_mujoco)(q, R);
        return;
    }
    REAL Rb[9], off[3], qq[4];
    SUF(rot_mujoco)(q, Rb);
    SUF(matvec3)(Rb, gpos, off);
    for (int k = 0; k < 3; ++k) c[k] = p[k] + off[k];
    SUF(mulquat)(q, gquat, qq);
    SUF(rot_mujoco)(qq, R);
}

/* gtype[B]: 0 sphere (size[0] = radius), 1 box (size = half extents); mass[B], inertia3[B][3], size3[B][3] and the geom
 * offsets gpos3[B][3] / gquat4[B][4] (NULL = none) describe ONE environment's bodies and are shared by all E.
 * Layout: qpos[E][B][7], qvel[E][B][6]; counters per body [E][B]. */
void SUF(rbo_step_multi_body)(long E, int B, int steps, REAL *qpos, REAL *qvel, const int *gtype, const REAL *mass,
                              const REAL *inertia3, const REAL *size3, const REAL *gpos3, const REAL *gquat4,
                              const REAL *plane_pos3, const REAL *plane_normal3, const REAL *gravity3, REAL dt,
                              REAL rest, REAL fric, unsigned *calls, unsigned *impulses) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long e = 0; e < E; ++e) {
        REAL *QP = qpos + (long)7 * B * e, *QV = qvel + (long)6 * B * e;
        REAL *gc = (REAL *)malloc(sizeof(REAL) * 12 * B);            /* start-of-step geom centre [3] + rotation [9] */
        for (int s = 0; s < steps; ++s) {
            for (int b = 0; b < B; ++b)                              /* mj_forward once per step (:43) */
                SUF(geom_pose)(QP + 7 * b, QP + 7 * b + 3, gpos3 != 0, gpos3 ? gpos3 + 3 * b : 0, gquat4 ? gquat4 + 4 * b : 0,
                               gc + 12 * b, gc + 12 * b + 3);
            for (int b = 0; b < B; ++b) {                            /* :46 */
                long eb = e * B + b;
                REAL *qp = QP + 7 * b, *qv = QV + 6 * b;
                REAL Iw[9];
                SUF(inertia_world_one)(inertia3 + 3 * b, qp + 3, Iw);              /* :55 */
                REAL vel[3] = {qv[0], qv[1], qv[2]}, om[3] = {qv[3], qv[4], qv[5]};
                SUF(external_forces)(vel, om, mass[b], Iw, gravity3, 0, dt);       /* :58-61 */
                for (int j = -1; j < B; ++j) {                       /* :64, MuJoCo contact order for body b */
                    SUF(contact_t) con[8];
                    int ncon = 0;
                    if (j == b) continue;
                    const REAL *cb_ = gc + 12 * b, *Rb_ = cb_ + 3, *hb_ = size3 + 3 * b;
                    if (j < 0) {
                        if (gtype[b] == 0) ncon = SUF(plane_sphere)(plane_pos3, plane_normal3, cb_, hb_[0], con);
                        else SUF(plane_box_rot)(plane_pos3, plane_normal3, cb_, Rb_, hb_, con, &ncon);
                    } else {
                        int lo = j < b ? j : b, hi = j < b ? b : j;
                        const REAL *cl = gc + 12 * lo, *ch = gc + 12 * hi;
                        const REAL *hl = size3 + 3 * lo, *hh = size3 + 3 * hi;
                        if (gtype[lo] == 0 && gtype[hi] == 0) ncon = SUF(sphere_sphere)(cl, hl[0], ch, hh[0], con);
                        else if (gtype[lo] == 1 && gtype[hi] == 1) ncon = SUF(box_box)(cl, cl + 3, hl, ch, ch + 3, hh, con);
                        else if (gtype[lo] == 0) ncon = SUF(sphere_box)(cl, hl[0], ch, ch + 3, hh, (REAL)1, con);
                        else ncon = SUF(sphere_box)(ch, hh[0], cl, cl + 3, hl, (REAL)-1, con);
                    }
                    for (int i = 0; i < ncon; ++i) {
                        if (!(con[i].dist < 0)) continue;            /* :66 */
                        REAL r[3] = {con[i].pos[0] - qp[0], con[i].pos[1] - qp[1], con[i].pos[2] - qp[2]};   /* :67 */
                        REAL jn, jt[3];
                        int imp = SUF(impulse_friction_one)(mass[b], vel, om, r, con[i].n, rest, fric, &jn, jt);   /* :69 */
                        if (calls) calls[eb] += 1;
                        if (impulses) impulses[eb] += (unsigned)imp;
                        SUF(apply_impulse_friction_one)(vel, om, mass[b], Iw, r, con[i].n, jn, jt);                /* :72 */
                    }
                }
                for (int i = 0; i < 3; ++i) qp[i] = qp[i] + vel[i] * dt;           /* :77 */
                SUF(integrate_quat)(qp + 3, om, dt);                               /* :78-82 */
                for (int i = 0; i < 3; ++i) { qv[i] = vel[i]; qv[3 + i] = om[i]; } /* :85-88 */
            }
        }
        free(gc);
    }
}

/* A10  compute_collision_impulse  src/simulation/ball_collision.py:53-68 (I_inv = iinv * Id, :39-41) */
static void SUF(two_ball_impulse)(REAL mass, REAL iinv, const REAL *v, const REAL *w, const REAL *r, const REAL *n,
                                  REAL e, REAL mu, REAL *J) {
    REAL wxr[3], vc[3], vt[3], t1[3], t2[3], tdir[3];
    SUF(cross)(w, r, wxr);
    for (int i = 0; i < 3; ++i) vc[i] = v[i] + wxr[i];                         /* :54 */
    REAL vn = V3DOT(vc, n);                                                    /* :55 */
    for (int i = 0; i < 3; ++i) vt[i] = vc[i] - vn * n[i];                     /* :56 */
    REAL tn = SUF(sqrtr)(V3DOT(vt, vt));                                       /* :57 */
    SUF(cross)(r, n, t1);
    for (int i = 0; i < 3; ++i) t1[i] = iinv * t1[i];
    SUF(cross)(t1, r, t2);
    REAL denom_n = ((REAL)1.0 / mass) + V3DOT(n, t2);                          /* :59 */
    REAL jn = (-(1 + e)) * vn / denom_n;                                       /* :60 */
    for (int i = 0; i < 3; ++i) tdir[i] = (tn > (REAL)1e-8) ? vt[i] / tn : (REAL)0;   /* :62 */
    SUF(cross)(r, tdir, t1);
    for (int i = 0; i < 3; ++i) t1[i] = iinv * t1[i];
    SUF(cross)(t1, r, t2);
    REAL denom_t = ((REAL)1.0 / mass) + V3DOT(tdir, t2);                       /* :63-64 */
    REAL jt = (-tn) / denom_t;                                                 /* :65 */
    REAL lim = mu * SUF(absr)(jn);
    if (jt < -lim) jt = -lim;                                                  /* :66 np.clip */
    if (jt > lim) jt = lim;
    for (int i = 0; i < 3; ++i) J[i] = jn * n[i] + jt * tdir[i];               /* :68 */
}

void SUF(rbo_two_ball_impulse)(long n, const REAL *mass, const REAL *iinv, const REAL *v3, const REAL *w3,
                               const REAL *r3, const REAL *n3, const REAL *e, const REAL *mu, REAL *J3) {
    for (long i = 0; i < n; ++i)
        SUF(two_ball_impulse)(mass[i], iinv[i], v3 + 3 * i, w3 + 3 * i, r3 + 3 * i, n3 + 3 * i, e[i], mu[i], J3 + 3 * i);
}

/* A11  step_with_custom_collisions  src/simulation/ball_collision.py:73-125.
 * qpos[E][14], qvel[E][12]; mass[E][2]; radius[E] (the reference hard-codes 0.1 for both, :23). */
void SUF(rbo_step_two_ball)(long E, int steps, REAL *qpos14, REAL *qvel12, const REAL *mass2, const REAL *radius,
                            const REAL *gravity3, REAL dt, REAL rest, REAL fric, unsigned *ground_hits,
                            unsigned *pair_hits) {
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (long e = 0; e < E; ++e) {
        REAL *qp = qpos14 + 14 * e, *qv = qvel12 + 12 * e;
        REAL rad = radius[e];
        REAL m[2] = {mass2[2 * e], mass2[2 * e + 1]};
        REAL iinv[2];
        for (int b = 0; b < 2; ++b) iinv[b] = (REAL)1.0 / ((((REAL)2.0 / (REAL)5.0) * m[b]) * (rad * rad));  /* :39-41 */
        for (int s = 0; s < steps; ++s) {
            for (int b = 0; b < 2; ++b)
                for (int i = 0; i < 3; ++i) qv[6 * b + i] = qv[6 * b + i] + gravity3[i] * dt;      /* :77-78 */
            for (int b = 0; b < 2; ++b) {                                                      /* :81-97 */
                REAL *p = qp + 7 * b, *v = qv + 6 * b, *w = qv + 6 * b + 3;
                const REAL nz[3] = {0, 0, 1};
                if (p[2] < rad) {                                                              /* :90 */
                    REAL cp[3] = {p[0] - rad * nz[0], p[1] - rad * nz[1], p[2] - rad * nz[2]}; /* :91 */
                    REAL r[3] = {cp[0] - p[0], cp[1] - p[1], cp[2] - p[2]};                    /* :92 */
                    REAL J[3], rxJ[3];
                    SUF(two_ball_impulse)(m[b], iinv[b], v, w, r, nz, rest, fric, J);          /* :93 */
                    SUF(cross)(r, J, rxJ);
                    for (int i = 0; i < 3; ++i) { v[i] = v[i] + J[i] / m[b]; w[i] = w[i] + iinv[b] * rxJ[i]; }  /* :95-96 */
                    p[2] = rad;                                                                /* :97 */
                    if (ground_hits) ground_hits[e] += 1;
                }
            }
            REAL *p1 = qp, *p2 = qp + 7;
            REAL diff[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};                      /* :100 */
            REAL dist = SUF(sqrtr)(V3DOT(diff, diff));                                         /* :101 */
            REAL tol = (REAL)0.01;
            if (dist < 2 * rad + tol) {                                                        /* :103 */
                REAL n[3], cp[3], r1[3], r2[3], J[3], x1[3], x2[3];
                for (int i = 0; i < 3; ++i) {
                    n[i] = diff[i] / (dist + (REAL)1e-8);                                      /* :104 */
                    cp[i] = (p1[i] + p2[i]) / (REAL)2.0;                                       /* :105 */
                    r1[i] = cp[i] - p1[i];
                    r2[i] = cp[i] - p2[i];
                }
                SUF(two_ball_impulse)(m[0], iinv[0], qv, qv + 3, r1, n, rest, fric, J);        /* :109-110 */
                SUF(cross)(r1, J, x1);
                SUF(cross)(r2, J, x2);
                for (int i = 0; i < 3; ++i) {
                    qv[i] = qv[i] + J[i] / m[0];                                               /* :111 */
                    qv[3 + i] = qv[3 + i] + iinv[0] * x1[i];                                   /* :112 */
                    qv[6 + i] = qv[6 + i] - J[i] / m[1];                                       /* :113 */
                    qv[9 + i] = qv[9 + i] - iinv[1] * x2[i];                                   /* :114 */
                }
                REAL corr = ((2 * rad + tol) - dist) / (REAL)2.0;                              /* :116 */
                for (int i = 0; i < 3; ++i) { p1[i] = p1[i] - corr * n[i]; p2[i] = p2[i] + corr * n[i]; }  /* :117-118 */
                if (pair_hits) pair_hits[e] += 1;
            }
            for (int b = 0; b < 2; ++b)
                for (int i = 0; i < 3; ++i) qp[7 * b + i] = qp[7 * b + i] + qv[6 * b + i] * dt;    /* :121-122 */
        }
    }
}

#undef V3DOT
