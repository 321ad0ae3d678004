/* rbsim_b200.h -- C ABI of the B200-native batched rigid-body stepper (librbsim_b200.so).
 *
 * The reference (pratyay2510/RigidBody-Simulation) has no FFI layer: its boundary for this path is
 * a set of Python functions.  Each entry point below replaces one of them for E independent
 * environments at once, and is what a ctypes / cffi binding on the reference side would call
 * (see INTEGRATION.md for the stubs).  Reference paths are relative to the reference root.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the name ends in _host;
 *   - `dtype` selects the arithmetic type of all `void*` arrays: RBS_F32 or RBS_F64
 *     (the reference computes in float64);
 *   - scalars that may be uniform or per-environment come as a pair {pointer, value}: a non-NULL
 *     pointer (array of n_env) wins, otherwise `value` is used for every environment;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every device entry
 *     point only enqueues work on that stream and returns without synchronising;
 *   - return value: 0 on success, a negative RBS_E* code otherwise; rbs_last_error() returns a
 *     thread-local message for the last failure.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *     RBS_ECUDA.
 *
 * State layout (device, SoA): `state` holds 13 component rows per body,
 *     row:  0 px 1 py 2 pz | 3 qw 4 qx 5 qy 6 qz | 7 vx 8 vy 9 vz | 10 wx 11 wy 12 wz
 *   single-body / two-ball steppers ("env-major"):  element (row c, body b, env e) at
 *       state[(c * n_body + b) * stride + e]            (stride >= n_env)
 *   multi-sphere stepper ("body-fastest"):          element (row c, env e, body b) at
 *       state[c * stride + e * n_body + b]              (stride >= n_env * n_body)
 *   which is the reference's qpos[7] (xyz + wxyz quaternion) and qvel[6] (linear, angular) per body
 *   (SURVEY.md Appendix A.1), transposed so that consecutive threads read consecutive addresses.
 */
#ifndef RBSIM_B200_H
#define RBSIM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define RBS_ABI_VERSION 1

#if defined(__GNUC__)
#define RBS_API __attribute__((visibility("default")))
#else
#define RBS_API
#endif

enum { RBS_F32 = 0, RBS_F64 = 1 };
enum { RBS_GEOM_SPHERE = 0, RBS_GEOM_BOX = 1 };
enum { RBS_SCHEME_A = 0, RBS_SCHEME_GENERAL = 1 };
enum { RBS_INERTIA_GENERAL = 0, RBS_INERTIA_ISOTROPIC = 1 };
enum { RBS_ARITH_STRICT = 0, RBS_ARITH_FAST = 1 };
enum { RBS_OK = 0, RBS_EINVAL = -1, RBS_ECUDA = -2, RBS_ENOMEM = -3 };

RBS_API int rbs_version(void);
RBS_API const char *rbs_last_error(void);
/* number of kernels this library has launched in the calling process (for launch accounting) */
RBS_API unsigned long long rbs_launch_count(void);
RBS_API int rbs_device_count(void);

/* Tuning knobs of the launch dispatch (new; the reference has nothing to tune).  None changes what is computed, only
 * which kernel instantiation / launch shape computes it.  Every knob starts from the environment variable of the
 * same name in upper case with an RBS_ prefix (RBS_PF_MIN_SUBSTEPS ...).  Names: "minb", "pf_min_substeps",
 * "pf_packed", "strict_minb", "strict_compact", "strict_tb_minb", "strict_ms_regs", "box_minb", "box_compact", "tb_minb", "ms_skin_percent", "ms_kernel", "ms_walk_cost", "ms_tight_span", "ms_regs", "probe_mode", "host_chunks".
 * "strict_compact" picks the strict single-body stepper: 0 = tuned (spheres: state resident in shared memory, contacts
 * queued across the CTA; uniform boxes in large batches: the same with a per-CTA contact-density vote), -1 = one thread per
 * environment throughout; the other codes are the A/B variants listed in DESIGN.md section 3 -- all give the same bits.
 * rbs_set_option returns RBS_EINVAL for an unknown name; rbs_get_option returns LONG_MIN for one. */
RBS_API int rbs_set_option(const char *name, long value);
RBS_API long rbs_get_option(const char *name);

/* ---------------------------------------------------------------------------------------------
 * Free functions, one work item per thread.  Vector arguments are [n][3] (row-major), matrices
 * [n][3][3]; outputs must not alias inputs.
 * --------------------------------------------------------------------------------------------- */

/* compute_collision_impulse_friction  -- src/physics/collision.py:7-48
 * out_jn[n], out_jt[n][3]; out_flag[n] (optional, may be NULL) = 1 where an impulse was computed
 * (u_n < 0), 0 where the function took its early return.  inertia_world is accepted by the
 * reference and unused (its k = 1/m + 1/18 is inertia-free), so it is not part of this ABI. */
RBS_API int rbs_impulse_friction(int dtype, long n, const void *mass, double mass_u, const void *vel, const void *omega,
                         const void *contact_point, const void *normal, const void *restitution,
                         double restitution_u, const void *friction, double friction_u, void *out_jn, void *out_jt,
                         unsigned char *out_flag, void *stream);

/* apply_impulse_friction  -- src/physics/physics_utils.py:25-49 */
RBS_API int rbs_apply_impulse_friction(int dtype, long n, const void *vel, const void *omega, const void *mass, double mass_u,
                               const void *inertia_world, const void *contact_point, const void *normal,
                               const void *jn, const void *jt, void *out_vel, void *out_omega, void *stream);

/* apply_impulse  -- src/physics/physics_utils.py:4-22 */
RBS_API int rbs_apply_impulse(int dtype, long n, const void *vel, const void *omega, const void *mass, double mass_u,
                      const void *inertia_world, const void *contact_point, const void *normal, const void *impulse,
                      double impulse_u, void *out_vel, void *out_omega, void *stream);

/* compute_inertia_tensor_world  -- src/physics/collision.py:51-53 (== time_integeration.py:8-10,
 * src/simulation/multi_sphere_bounce.py:35-37).  inertia_diag[n][3], quat[n][4] wxyz -> out[n][3][3] */
RBS_API int rbs_inertia_world(int dtype, long n, const void *inertia_diag, const void *quat, void *out, void *stream);

/* compute_collision_impulse  -- src/simulation/ball_collision.py:53-68 with I_inv = inv_inertia * Id
 * (compute_inverse_inertia, :39-41).  Returns the impulse VECTOR out_J[n][3]. */
RBS_API int rbs_two_ball_impulse(int dtype, long n, const void *mass, double mass_u, const void *inv_inertia,
                         double inv_inertia_u, const void *v_lin, const void *v_ang, const void *r, const void *normal,
                         const void *restitution, double restitution_u, const void *friction, double friction_u,
                         void *out_J, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Fused steppers: `substeps` integration steps per launch, state held in registers in between.
 * --------------------------------------------------------------------------------------------- */

/* One free body (sphere or box) above one plane per environment.
 *   scheme RBS_SCHEME_A       : custom_step_with_impulse_collision_friction, src/physics/collision.py:56-102
 *                               == timestep_integration, src/physics/time_integeration.py:13-72
 *   scheme RBS_SCHEME_GENERAL : general, src/physics/time_integeration.py:75-141
 * including the plane-sphere / plane-box narrow phase that the reference obtains from
 * mujoco.mj_forward. */
typedef struct rbs_body_plane_args {
    int dtype;                 /* RBS_F32 | RBS_F64 */
    int geom;                  /* RBS_GEOM_* */
    int scheme;                /* RBS_SCHEME_* */
    int inertia_mode;          /* RBS_INERTIA_ISOTROPIC is valid only when the three principal moments are equal */
    long n_env;
    long stride;               /* elements between state rows, >= n_env */
    long param_stride;         /* row stride of the multi-row per-env arrays (inertia, size, xfrc); 0 = n_env.
                                  Lets a call cover a window of a larger batch: offset every pointer by the
                                  window start, pass the window length as n_env and the batch size here. */
    int substeps;              /* >= 1 */
    int arith;                 /* RBS_ARITH_STRICT: the reference's rounding sequence (bit-faithful);
                                  RBS_ARITH_FAST: FMA / reciprocal-multiply re-association, <= 1e-12 per step in
                                  fp64; implemented for scheme A + isotropic inertia (sphere or box) */
    void *state;               /* [13][stride], env-major, n_body = 1 */
    const void *mass;          double mass_u;
    const void *inertia;       double inertia_u[3];   /* [3][n_env] body-frame principal moments */
    const void *size;          double size_u[3];      /* [3][n_env]: sphere radius in row 0; box half extents */
    const void *restitution;   double restitution_u;
    const void *friction;      double friction_u;
    const void *xfrc;          /* [6][n_env] applied force(3)+torque(3) (data.xfrc_applied), or NULL */
    double plane_point[3];
    double plane_normal[3];    /* unit normal = z axis of the plane geom */
    double gravity[3];
    double dt;
    double contact_threshold;  /* contacts with |dist| < threshold are skipped (collision.py:79-80) */
    unsigned *n_contacts;      /* [n_env] += contacts handed to the impulse routine, or NULL */
    unsigned *n_impulses;      /* [n_env] += contacts that produced an impulse (u_n < 0), or NULL */
    void *trajectory;          /* [substeps][trajectory_envs][3] (dtype, device) or NULL: position of the first
                                  trajectory_envs environments after EVERY substep of the launch -- what the reference's
                                  logger.record sees per frame (src/viewer/mujoco_viewer.py:116-119).  Device-resident
                                  stepping only (rbs_run_body_plane_host refuses it). */
    long trajectory_envs;      /* clamped to n_env; ignored when trajectory is NULL */
    void *stream;
} rbs_body_plane_args;
RBS_API int rbs_step_body_plane(const rbs_body_plane_args *a);

/* Two balls + ground: step_with_custom_collisions, src/simulation/ball_collision.py:73-125
 * (quaternions are never read or written by this step; rows 3..6 of `state` are untouched). */
typedef struct rbs_two_ball_args {
    int dtype;
    int substeps;
    int arith;                 /* RBS_ARITH_STRICT | RBS_ARITH_FAST */
    int reserved;
    long n_env;
    long stride;
    void *state;               /* [13][2][stride], env-major */
    const void *mass;          double mass_u[2];      /* [2][n_env] */
    const void *radius;        double radius_u;       /* [n_env]; the reference hard-codes 0.1 (:23) */
    double gravity[3];
    double dt;
    double restitution;
    double friction;
    unsigned *n_ground_hits;   /* [n_env] or NULL */
    unsigned *n_pair_hits;     /* [n_env] or NULL */
    void *stream;
} rbs_two_ball_args;
RBS_API int rbs_step_two_ball(const rbs_two_ball_args *a);

/* n_body spheres + ground plane per environment, all-pairs contacts:
 * custom_step_multi_sphere, src/simulation/multi_sphere_bounce.py:42-92, with the index / ownership
 * repairs documented in DESIGN.md (the shipped file raises IndexError on its first step). */
typedef struct rbs_multi_sphere_args {
    int dtype;
    int substeps;
    int n_body;                /* 1 .. 1024 */
    int inertia_mode;
    int arith;                 /* RBS_ARITH_FAST needs RBS_INERTIA_ISOTROPIC (spheres: always valid) */
    int list_skin_percent;     /* partner-list skin in % of the radius: 0 = adaptive per CTA (starts at 50, or at
                                  RBS_MS_SKIN_PERCENT; launches of < 4 substeps scan instead); > 0 = pinned; < 0 = no skin (all partners are scanned every
                                  substep).  A tuning knob: results never depend on it. */
    long n_env;
    long stride;               /* >= n_env * n_body */
    void *state;               /* [13][stride], body-fastest */
    const void *mass;          double mass_u;         /* [n_env * n_body] */
    const void *inertia;       double inertia_u[3];   /* [3][n_env * n_body] */
    const void *radius;        double radius_u;       /* [n_env * n_body] */
    double plane_point[3];
    double plane_normal[3];
    double gravity[3];
    double dt;
    double restitution;
    double friction;
    unsigned *n_contacts;      /* [n_env * n_body] or NULL */
    unsigned *n_impulses;      /* [n_env * n_body] or NULL */
    void *stream;
} rbs_multi_sphere_args;
RBS_API int rbs_step_multi_sphere(const rbs_multi_sphere_args *a);

/* ---------------------------------------------------------------------------------------------
 * SURVEY.md section 8(f) row N4 -- multi-body scenes with spheres AND boxes (new behaviour: the reference's scripts
 * only ever meet plane-sphere, plane-box and sphere-sphere contacts).  The step is the loop rbs_step_multi_sphere
 * replaces (src/simulation/multi_sphere_bounce.py:42-92, repaired), body by body with the reference's impulse
 * (src/physics/collision.py:7-48), its application (src/physics/physics_utils.py:25-49) and the literal world inertia
 * (collision.py:51-53) per contact, strict arithmetic; the contact set adds sphere-box and box-box (vertices inside
 * the other box, edges passing through it) pairs, geom1 = the lower body index, normal geom1 -> geom2, never flipped (DESIGN.md "N4").
 * body_table: DEVICE array [n_body][RBS_BODY_TABLE_WIDTH] of dtype, shared by every environment; per body
 *   [0] geom type (0 sphere, 1 box)   [1..3] size (radius,-,- | half extents)   [4] mass   [5..7] principal inertia
 *   [8..10] geom position and [11..14] geom quaternion (wxyz) in the body frame (read only when has_offset != 0)
 *   [15] radius of the geom's bounding sphere times (1 + 1e-6) (broad phase).
 * State layout: body-fastest, as rbs_step_multi_sphere.
 * --------------------------------------------------------------------------------------------- */
#define RBS_BODY_TABLE_WIDTH 16
typedef struct rbs_multi_body_args {
    int dtype;
    int substeps;
    int n_body;                /* 1 .. 256 */
    int has_offset;            /* 0: every geom sits at its body's origin with the body's orientation */
    long n_env;
    long stride;               /* >= n_env * n_body */
    void *state;               /* [13][stride], body-fastest */
    const void *body_table;    /* [n_body][RBS_BODY_TABLE_WIDTH] */
    double plane_point[3];
    double plane_normal[3];
    double gravity[3];
    double dt;
    double restitution;
    double friction;
    unsigned *n_contacts;      /* [n_env * n_body] or NULL */
    unsigned *n_impulses;      /* [n_env * n_body] or NULL */
    void *stream;
} rbs_multi_body_args;
RBS_API int rbs_step_multi_body(const rbs_multi_body_args *a);

/* ---------------------------------------------------------------------------------------------
 * Layout conversion between the reference's per-env qpos[7*B] / qvel[6*B] and the SoA state.
 * body_fastest = 0 -> env-major layout, 1 -> body-fastest layout (see top of file).
 * --------------------------------------------------------------------------------------------- */
RBS_API int rbs_pack_state(int dtype, long n_env, int n_body, int body_fastest, const void *qpos, const void *qvel,
                   void *state, long stride, void *stream);
RBS_API int rbs_unpack_state(int dtype, long n_env, int n_body, int body_fastest, const void *state, long stride, void *qpos,
                     void *qvel, void *stream);

/* mj_resetData for a subset of the batch (the viewer's BACKSPACE handler, src/viewer/mujoco_viewer.py:62-65):
 * qpos <- qpos0[7 * n_body] (device, dtype), qvel <- 0 and the event counters <- 0 for every environment whose
 * env_mask byte (device, [n_env]) is non-zero; env_mask == NULL resets all.  Counter arrays may be NULL; they are
 * indexed like the state ([n_body][n_env] env-major, [n_env][n_body] body-fastest). */
RBS_API int rbs_reset_envs(int dtype, long n_env, int n_body, int body_fastest, void *state, long stride, const void *qpos0,
                   const unsigned char *env_mask, unsigned *n_contacts, unsigned *n_impulses, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Host-buffer drivers: the call a reference-side `for _ in range(steps): step_function(model, data, dt)`
 * loop is replaced by.  qpos_host / qvel_host are HOST arrays in the reference layout
 * ([n_env][7*B], [n_env][6*B]); they are copied to the device, advanced `total_steps` steps in
 * launches of `a->substeps`, copied back, and the stream is synchronised before returning.
 * `a->state` and `a->stride` are ignored (an internal workspace is used); per-env parameter
 * pointers in `a` stay DEVICE pointers.  Pinned host memory gives full PCIe bandwidth.
 * --------------------------------------------------------------------------------------------- */
RBS_API int rbs_run_body_plane_host(const rbs_body_plane_args *a, void *qpos_host, void *qvel_host, long total_steps);
RBS_API int rbs_run_two_ball_host(const rbs_two_ball_args *a, void *qpos_host, void *qvel_host, long total_steps);
RBS_API int rbs_run_multi_sphere_host(const rbs_multi_sphere_args *a, void *qpos_host, void *qvel_host, long total_steps);
RBS_API int rbs_run_multi_body_host(const rbs_multi_body_args *a, void *qpos_host, void *qvel_host, long total_steps);
/* frees the cached device workspace of the host-buffer drivers */
RBS_API int rbs_release_workspace(void);

/* ---------------------------------------------------------------------------------------------
 * Run statistics in one pass (new behaviour; the reference only appends positions to Python lists,
 * src/visualization/logger_base.py:22-32).  `state` holds n_bodies columns with row stride `stride` (either
 * layout flattened: n_bodies = n_env * n_body).  out[5] (device, double) must be initialised by the caller with
 * {0, 0, -1e300, 0, 0}; the call accumulates
 *   out[0] += kinetic energy (linear + rotational), out[1] += potential energy -m g.p,
 *   out[2]  = max height along -g (or +z when g = 0), out[3] += n_contacts, out[4] += n_impulses.
 * mass [n_bodies] / inertia [3][inertia_stride] may be NULL (uniform values are used).
 * --------------------------------------------------------------------------------------------- */
RBS_API int rbs_stats(int dtype, long n_bodies, const void *state, long stride, const void *mass, double mass_u,
                      const void *inertia, long inertia_stride, const double inertia_u[3], const double gravity[3],
                      const unsigned *n_contacts, const unsigned *n_impulses, double *out, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Measurement helper (not part of the reference's surface): a dependent-chain-free FMA loop used by
 * bench.py to measure the FP32 / FP64 CUDA-core peak of the device it runs on.
 * Writes nothing useful to `sink` (n_threads elements) but keeps the compiler honest.
 * flops performed = 2 * n_threads * iters * 8.
 * --------------------------------------------------------------------------------------------- */
RBS_API int rbs_fma_probe(int dtype, long n_threads, int iters, void *sink, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RBSIM_B200_H */
