// rbs_capi.cu -- extern "C" boundary of librbsim_b200.so (declared in include/rbsim_b200.h).
//
// Argument validation, dtype / template dispatch, launch configuration, error capture and the
// host-buffer drivers.  Compiled with -fmad=false so that the kernels in rbs_kernels.cuh reproduce
// the reference's rounding sequence ("strict" arithmetic policy).  There is no CPU fallback: without
// a usable CUDA device every compute entry point returns RBS_ECUDA.
#include "rbs_kernels.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <climits>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <mutex>

#include "../../include/rbsim_b200.h"

namespace {

thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int check_launch(const char *what) {
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) return fail(RBS_ECUDA, "%s: %s", what, cudaGetErrorString(err));
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return RBS_OK;
}

// Tuning knobs (rbs_set_option / rbs_get_option).  Each starts from its environment variable, so a deployment can
// still pin it from outside; tests and benchmarks flip them in-process.  None of them changes what is computed,
// only which instantiation / launch shape computes it.
struct Option {
    const char *name, *env;
    std::atomic<long> value;
    long fallback;
};
Option g_options[] = {
    {"minb", "RBS_MINB", {0}, 0},                            // resident CTAs per SM of the sphere steppers (0 = tuned default)
    {"pf_min_substeps", "RBS_PF_MIN_SUBSTEPS", {0}, 4},      // shortest launch that takes the plane-frame kernels
    {"pf_packed", "RBS_PF_PACKED", {0}, 1},                  // float sphere stepper: packed fp32x2 kernel (1) or scalar (0)
    {"strict_minb", "RBS_STRICT_MINB", {0}, 0},              // resident CTAs per SM of the strict literal-inertia stepper (2, 4, 5, 6; 0 = tuned)
    {"strict_compact", "RBS_STRICT_COMPACT", {0}, 0},        // strict single-body stepper: -1 = thread per environment, 4 / 5 = CTA-compacted contact path, 0 = tuned
    {"strict_tb_minb", "RBS_STRICT_TB_MINB", {0}, 0},        // resident CTAs per SM of the strict two-ball stepper (3, 4, 5; 0 = tuned)
    {"strict_ms_regs", "RBS_STRICT_MS_REGS", {0}, 0},        // register cap of the strict literal-inertia multi-sphere stepper (168, 128, 96; 0 = tuned)
    {"box_minb", "RBS_BOX_MINB", {0}, 6},                    // resident CTAs per SM of the plane-frame box kernel
    {"box_compact", "RBS_BOX_COMPACT", {0}, 0},              // plane-frame box kernel: CTA-level compaction of contacts (measured slower: off)
    {"tb_minb", "RBS_TB_MINB", {0}, 0},                      // resident CTAs per SM of the two-ball fast kernel (5, 6 or 8; 0 = 6 in double, 8 in float)
    {"tb_uniform", "RBS_TB_UNIFORM", {0}, 1},                // two-ball fast stepper, uniform radius: radius / reach as launch parameters (1) or registers (0)
    {"tb_packed", "RBS_TB_PACKED", {0}, 0},                  // float two-ball fast stepper: packed fp32x2 kernel, two envs per thread (1) or scalar (0: measured equal, fewer ragged waves)
    {"mb_minb", "RBS_MB_MINB", {0}, 0},                      // resident CTAs per SM of the multi-body stepper (1, 2, 3; 0 = tuned)
    {"ms_skin_percent", "RBS_MS_SKIN_PERCENT", {0}, 50},     // starting skin of the adaptive partner lists
    {"ms_kernel", "RBS_MS_KERNEL", {0}, 2},                  // multi-sphere fast policy: 2 = plane-frame kernel, 1 = first generation
    {"ms_walk_cost", "RBS_MS_WALK_COST", {0}, 40},           // plane-frame multi-sphere kernel: cost of a list entry per substep (skin controller)
    {"ms_tight_span", "RBS_MS_TIGHT_SPAN", {0}, 32},         // plane-frame multi-sphere kernel: substeps per TIGHT (no-skin) span, 0 = never
    {"ms_regs", "RBS_MS_REGS", {0}, 96},                     // register cap of the frictionless plane-frame multi-sphere kernel (96 or 128)
    {"probe_mode", "RBS_PROBE_MODE", {0}, 1},                // rbs_fma_probe operand mode
    {"host_chunks", "RBS_HOST_CHUNKS", {0}, 16},             // pipeline depth of rbs_run_body_plane_host
    {"host_streams", "RBS_HOST_STREAMS", {0}, 3},            // compute streams of the host-buffer pipeline = chunks stepped concurrently (1..4; profiles/r2_ab_host_pipeline.jsonl)
    {"host_wave_ctas", "RBS_HOST_WAVE_CTAS", {0}, 0},        // CTAs per SM that make one chunk quantum of rbs_run_body_plane_host (0 = 4)
};
std::once_flag g_options_once;
Option *find_option(const char *name) {
    std::call_once(g_options_once, [] {
        for (Option &o : g_options) {
            const char *e = getenv(o.env);
            o.value.store(e ? atol(e) : o.fallback, std::memory_order_relaxed);
        }
    });
    if (!name) return nullptr;
    for (Option &o : g_options)
        if (strcmp(o.name, name) == 0) return &o;
    return nullptr;
}
inline long option(const char *name) { return find_option(name)->value.load(std::memory_order_relaxed); }

inline unsigned blocks_for(long n, int block) { return (unsigned)((n + block - 1) / block); }
inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }
inline bool bad_dtype(int dtype) { return dtype != RBS_F32 && dtype != RBS_F64; }
inline size_t elem_size(int dtype) { return dtype == RBS_F64 ? sizeof(double) : sizeof(float); }

// A launch may cover a window [off, off+cnt) of the environments described by `a` (the host-buffer driver
// pipelines chunks); per-env parameter arrays keep their full-size row stride.
struct Window {
    long off, cnt;             // environments [off, off + cnt) of the batch described by the argument struct
    void *state;               // SoA state of the window
    long stride;
    cudaStream_t stream;
};

inline Window whole(const rbs_body_plane_args *a) { return {0, a->n_env, a->state, a->stride, as_stream(a->stream)}; }

// Plane frame of the fused fast kernels (double on the host): rows t1, t2, n of the world->plane rotation, the
// quaternion of that rotation (wxyz) and gravity*dt expressed in the frame.  x' is taken along n x g, so gravity has no
// x' component and the kernels add two components per substep.
void plane_frame(const double *n, const double *g, double dt, double *R, double *q, double *gdt_pf) {
    double t1[3] = {0, 0, 0}, t2[3];
    const double c[3] = {n[1] * g[2] - n[2] * g[1], n[2] * g[0] - n[0] * g[2], n[0] * g[1] - n[1] * g[0]};
    const double clen = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    const double glen = sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]);
    if (clen > 1e-9 * glen) {
        for (int i = 0; i < 3; ++i) t1[i] = c[i] / clen;
    } else {                            // g parallel to n (flat ground) or g = 0: any tangent will do
        const int ax = fabs(n[0]) < 0.9 ? 0 : 1;
        t1[ax] = 1.0;
        const double d = n[ax];
        double len = 0.0;
        for (int i = 0; i < 3; ++i) { t1[i] -= d * n[i]; len += t1[i] * t1[i]; }
        len = sqrt(len);
        for (int i = 0; i < 3; ++i) t1[i] /= len;
    }
    t2[0] = n[1] * t1[2] - n[2] * t1[1]; t2[1] = n[2] * t1[0] - n[0] * t1[2]; t2[2] = n[0] * t1[1] - n[1] * t1[0];
    const double M[9] = {t1[0], t1[1], t1[2], t2[0], t2[1], t2[2], n[0], n[1], n[2]};
    for (int i = 0; i < 9; ++i) R[i] = M[i];
    const double tr = R[0] + R[4] + R[8];
    if (tr > 0) {
        const double s4 = 2.0 * sqrt(tr + 1.0);
        q[0] = 0.25 * s4; q[1] = (R[7] - R[5]) / s4; q[2] = (R[2] - R[6]) / s4; q[3] = (R[3] - R[1]) / s4;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        const double s4 = 2.0 * sqrt(1.0 + R[0] - R[4] - R[8]);
        q[0] = (R[7] - R[5]) / s4; q[1] = 0.25 * s4; q[2] = (R[1] + R[3]) / s4; q[3] = (R[2] + R[6]) / s4;
    } else if (R[4] > R[8]) {
        const double s4 = 2.0 * sqrt(1.0 + R[4] - R[0] - R[8]);
        q[0] = (R[2] - R[6]) / s4; q[1] = (R[1] + R[3]) / s4; q[2] = 0.25 * s4; q[3] = (R[5] + R[7]) / s4;
    } else {
        const double s4 = 2.0 * sqrt(1.0 + R[8] - R[0] - R[4]);
        q[0] = (R[3] - R[1]) / s4; q[1] = (R[2] + R[6]) / s4; q[2] = (R[5] + R[7]) / s4; q[3] = 0.25 * s4;
    }
    for (int i = 0; i < 3; ++i) gdt_pf[i] = (R[3 * i] * g[0] + R[3 * i + 1] * g[1] + R[3 * i + 2] * g[2]) * dt;
}

template <typename T> rbs::BodyPlaneParams<T> make_params(const rbs_body_plane_args *a, const Window &w) {
    rbs::BodyPlaneParams<T> p;
    auto at = [&](const void *base) { return base ? static_cast<const T *>(base) + w.off : nullptr; };
    p.n_env = w.cnt;
    p.stride = w.stride;
    p.pstride = a->param_stride > 0 ? a->param_stride : a->n_env;
    p.substeps = a->substeps;
    p.state = static_cast<T *>(w.state);
    p.mass = at(a->mass);
    p.inertia = at(a->inertia);
    p.size = at(a->size);
    p.rest = at(a->restitution);
    p.fric = at(a->friction);
    p.xfrc = at(a->xfrc);
    p.mass_u = (T)a->mass_u;
    p.rest_u = (T)a->restitution_u;
    p.fric_u = (T)a->friction_u;
    for (int i = 0; i < 3; ++i) {
        p.inertia_u[i] = (T)a->inertia_u[i];
        p.size_u[i] = (T)a->size_u[i];
        p.pp[i] = (T)a->plane_point[i];
        p.pn[i] = (T)a->plane_normal[i];
        p.g[i] = (T)a->gravity[i];
    }
    p.dt = (T)a->dt;
    p.thr = (T)a->contact_threshold;
    for (int i = 0; i < 3; ++i) p.gdt[i] = p.g[i] * p.dt;
    p.hdt = (T)0.5 * p.dt;
    p.inv_hdt = (T)1 / p.hdt;
    p.inv_dt = (T)0.5 * p.inv_hdt;
    {
        double R[9], q[4], gpf[3];
        plane_frame(a->plane_normal, a->gravity, a->dt, R, q, gpf);
        for (int i = 0; i < 9; ++i) p.frame[i] = (T)R[i];
        for (int i = 0; i < 4; ++i) p.frame_q[i] = (T)q[i];
        for (int i = 0; i < 3; ++i) p.gdt_pf[i] = (T)gpf[i];
    }
    p.n_contacts = a->n_contacts ? a->n_contacts + w.off : nullptr;
    p.n_impulses = a->n_impulses ? a->n_impulses + w.off : nullptr;
    p.traj = static_cast<T *>(a->trajectory);
    p.traj_envs = a->trajectory ? (a->trajectory_envs < w.cnt ? a->trajectory_envs : w.cnt) : 0;
    return p;
}

// Resident CTAs per SM the headline kernels are compiled for (register cap 65536 / (128 * MINB)).
// Measured on B200, 1M envs fp64 (profiles/r1_summary.md): strict policy -- fused launches fastest at 6 (80 regs),
// the one-substep streaming launch at 8 (64 regs, more loads in flight); fast policy -- 6 (79 regs, no spills) for both.
// RBS_MINB overrides for experiments.
int tuning_minb(int substeps, int arith) {
    const int forced = (int)option("minb");
    if (forced) return forced;
    if (arith == RBS_ARITH_FAST) return 6;
    return substeps <= 2 ? 8 : 6;
}

template <typename T, int GEOM, int SCHEME> void launch_body_plane_iso(const rbs_body_plane_args *a, const Window &w) {
    const rbs::BodyPlaneParams<T> p = make_params<T>(a, w);
    const unsigned grid = blocks_for(w.cnt, rbs::kBlock);
    cudaStream_t st = w.stream;
    if (a->trajectory) {                     // per-substep position log of the sampled environments
        if (a->inertia_mode == RBS_INERTIA_ISOTROPIC) rbs::step_body_plane_kernel<T, GEOM, SCHEME, 1, 4, true><<<grid, rbs::kBlock, 0, st>>>(p);
        else rbs::step_body_plane_kernel<T, GEOM, SCHEME, 0, 2, true><<<grid, rbs::kBlock, 0, st>>>(p);
        return;
    }
    if (a->inertia_mode == RBS_INERTIA_ISOTROPIC) {
        if (GEOM == 0 && SCHEME == 0) {      // occupancy variants of the headline kernel
            switch (tuning_minb(a->substeps, RBS_ARITH_STRICT)) {
                case 6: rbs::step_body_plane_kernel<T, GEOM, SCHEME, 1, 6><<<grid, rbs::kBlock, 0, st>>>(p); return;
                case 8: rbs::step_body_plane_kernel<T, GEOM, SCHEME, 1, 8><<<grid, rbs::kBlock, 0, st>>>(p); return;
                default: break;
            }
        }
        rbs::step_body_plane_kernel<T, GEOM, SCHEME, 1, 4><<<grid, rbs::kBlock, 0, st>>>(p);
    } else {
        if constexpr (SCHEME == 0) {
            // the contact path compacted across the CTA (bit-identical results; strict_compact = 0 keeps one environment's
            // whole substep in its own thread).  Not with an applied wrench: then every lane builds the inverse anyway.
            // Option strict_compact: -1 = never, 4 / 5 = always (at 4 / 5 resident CTAs per SM), 2x-4x = K environments per thread
            // in registers (measured slower), 5x-8x = K environments per thread RESIDENT IN SHARED MEMORY (uniform mass / size /
            // inertia), 0 = where it measured faster (profiles/r2_ab_strict.jsonl, r2_ab_strict_resident.jsonl: sphere 2.28e10 ->
            // 2.56e10 compacted at 5 CTAs -> 3.60e10 resident at K = 3, 4 CTAs; cube 1.02e10 -> 7.7e9 / 1.08e10 / 6.6e9 (incline),
            // so boxes stayed on the thread-per-environment kernel until the rolled hybrid below).
            long compact = option("strict_compact");
            // boxes: the rolled K = 4 resident kernel whose dense CTAs take the thread-per-environment loop (374: cube bounce
            // 1.32e10 -> 1.64e10, incline 9.6e9 -> 8.7e9 at 1M environments -- 512 environments per CTA are 3.5 waves there --
            // profiles/r2_ab_strict_hybrid.jsonl); short launches and batches that would not fill the GPU with 512-environment CTAs
            // keep the thread-per-environment kernel
            if (compact == 0) compact = GEOM == 0 ? ((!a->mass && !a->size && !a->inertia) ? 64 : 5)
                                                  : ((!a->mass && !a->size && !a->inertia && a->substeps >= 16 && w.cnt >= (1L << 18)) ? 374 : -1);
            if (!a->xfrc && compact >= 50 && !a->mass && !a->size && !a->inertia) {
                // state resident in shared memory, K environments per thread: compact = 50 + 10*(K - 2) + resident CTAs, K = 2..4
#define RBS_RES(KK, MB, WP, ...)                                                                                              \
    do {                                                                                                                 \
        const size_t smem__ = (size_t)(KK) * rbs::kBlock * (13 * sizeof(T) + 4);                                         \
        cudaFuncSetAttribute(rbs::step_body_plane_resident_kernel<T, GEOM, KK, MB, WP, ##__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem__); \
        cudaFuncSetAttribute(rbs::step_body_plane_resident_kernel<T, GEOM, KK, MB, WP, ##__VA_ARGS__>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        rbs::step_body_plane_resident_kernel<T, GEOM, KK, MB, WP, ##__VA_ARGS__><<<blocks_for(w.cnt, (KK) * rbs::kBlock), rbs::kBlock, smem__, st>>>(p); \
    } while (0)
                // (the A/B of profiles/r2_ab_strict_resident.jsonl also had K = 2..5 at 3..6 resident CTAs; the instantiations kept
                // are the best of each K).  1KM: the same with one queue per WARP (K environments per lane, M resident CTAs): bit-identical,
                // 3.0 / 3.3 / 3.4e10 at 124 / 134 / 135 against 3.6e10 (profiles/r2_ab_strict_warp_queues.jsonl; K = 4, 6 slower still).
                switch (compact) {
                    case 54: RBS_RES(2, 4, false); return;
                    case 65: RBS_RES(3, 5, false); return;
                    case 74: RBS_RES(4, 4, false); return;
                    case 264: RBS_RES(3, 4, false, true); return;     // 2KM: phases A and C as rolled loops over the K columns (cube bounce:
                                                                      // 1.56 / 1.63e10 at 264 / 274 against 1.32e10 thread-per-env; incline and sphere slower)
                    case 274: RBS_RES(4, 4, false, true); return;
                    case 364: RBS_RES(3, 4, false, true, true); return;     // 3KM: rolled, and dense CTAs take the thread-per-environment loop
                    case 374: RBS_RES(4, 4, false, true, true); return;
                    case 124: RBS_RES(2, 4, true); return;
                    case 134: RBS_RES(3, 4, true); return;
                    case 135: RBS_RES(3, 5, true); return;
                    default: RBS_RES(3, 4, false); return;
                }
#undef RBS_RES
            }
            if (!a->xfrc && compact >= 20 && compact < 50) {
                // K environments per thread (K * 128 per CTA), all warps work through the contact phase: compact = 10*K + resident CTAs
#define RBS_CM(KK, MB)                                                                                                   \
    do {                                                                                                                 \
        const size_t smem__ = (size_t)(KK) * rbs::kBlock * (13 * sizeof(T) + 3 * sizeof(unsigned));                      \
        cudaFuncSetAttribute(rbs::step_body_plane_compact_multi_kernel<T, GEOM, KK, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem__); \
        cudaFuncSetAttribute(rbs::step_body_plane_compact_multi_kernel<T, GEOM, KK, MB>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
        rbs::step_body_plane_compact_multi_kernel<T, GEOM, KK, MB><<<blocks_for(w.cnt, (KK) * rbs::kBlock), rbs::kBlock, smem__, st>>>(p); \
    } while (0)
                // (profiles/r2_ab_strict_multi_env.jsonl measured K = 2..4 at 3..6 resident CTAs, all slower; two are kept as the record)
                switch (compact) {
                    case 25: RBS_CM(2, 5); return;
                    default: RBS_CM(4, 4); return;
                }
#undef RBS_CM
            }
            if (!a->xfrc && compact > 0) {
                if (compact >= 8) {
                    cudaFuncSetAttribute(rbs::step_body_plane_compact_kernel<T, GEOM, 8>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    rbs::step_body_plane_compact_kernel<T, GEOM, 8><<<grid, rbs::kBlock, 0, st>>>(p);
                } else if (compact >= 6) {
                    cudaFuncSetAttribute(rbs::step_body_plane_compact_kernel<T, GEOM, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    rbs::step_body_plane_compact_kernel<T, GEOM, 6><<<grid, rbs::kBlock, 0, st>>>(p);
                } else if (compact >= 5) {
                    cudaFuncSetAttribute(rbs::step_body_plane_compact_kernel<T, GEOM, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    rbs::step_body_plane_compact_kernel<T, GEOM, 5><<<grid, rbs::kBlock, 0, st>>>(p);
                } else {
                    cudaFuncSetAttribute(rbs::step_body_plane_compact_kernel<T, GEOM, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                    rbs::step_body_plane_compact_kernel<T, GEOM, 4><<<grid, rbs::kBlock, 0, st>>>(p);
                }
                return;
            }
        }
        // literal inv(R diag(I) R^T): resident CTAs per SM (register cap 255 / 128 / 96 / 80) -- option strict_minb
        if constexpr (SCHEME == 0) {
            // 0 = measured best on B200 (profiles/r2_ab_strict.jsonl: +14 % sphere, +25-28 % cube over the uncapped build)
            int minb = (int)option("strict_minb");
            if (minb == 0) minb = GEOM == 0 ? 5 : 4;
            switch (minb) {
                case 4: rbs::step_body_plane_kernel<T, GEOM, SCHEME, 0, 4><<<grid, rbs::kBlock, 0, st>>>(p); return;
                case 5: rbs::step_body_plane_kernel<T, GEOM, SCHEME, 0, 5><<<grid, rbs::kBlock, 0, st>>>(p); return;
                case 6: rbs::step_body_plane_kernel<T, GEOM, SCHEME, 0, 6><<<grid, rbs::kBlock, 0, st>>>(p); return;
                default: break;
            }
        }
        rbs::step_body_plane_kernel<T, GEOM, SCHEME, 0, 2><<<grid, rbs::kBlock, 0, st>>>(p);
    }
}

template <typename T> void launch_sphere_plane_fast(const rbs_body_plane_args *a, const Window &w) {
    const rbs::BodyPlaneParams<T> p = make_params<T>(a, w);
    const unsigned grid = blocks_for(w.cnt, rbs::kBlock);
    cudaStream_t st = w.stream;
    if (a->xfrc) {
        rbs::step_sphere_plane_fast_kernel<T, 4, true><<<grid, rbs::kBlock, 0, st>>>(p);
        return;
    }
    if (a->substeps >= option("pf_min_substeps") && !a->trajectory) {   // fused launches: work in the plane frame (two rotations per launch pay off)
        const bool count = a->n_contacts || a->n_impulses, thr = a->contact_threshold > 0;
        if constexpr (sizeof(T) == 4) {
            // float: two environments per thread, packed fp32x2 arithmetic (bit-identical to the scalar kernel; the
            // scalar one stays reachable with RBS_PF_PACKED=0 so that the tests can compare the two in one process)
            if (option("pf_packed") != 0) {
                const unsigned grid2 = blocks_for(w.cnt, 2 * rbs::kBlock);
#define RBS_PF2(COUNT, THR) rbs::step_sphere_plane_pf2_kernel<6, COUNT, THR><<<grid2, rbs::kBlock, 0, st>>>(p)
                if (count) { if (thr) RBS_PF2(true, true); else RBS_PF2(true, false); }
                else { if (thr) RBS_PF2(false, true); else RBS_PF2(false, false); }
#undef RBS_PF2
                return;
            }
        }
        const int pf_minb = tuning_minb(a->substeps, RBS_ARITH_FAST);
#define RBS_PF(MINB, COUNT, THR) rbs::step_sphere_plane_pf_kernel<T, MINB, COUNT, THR><<<grid, rbs::kBlock, 0, st>>>(p)
#define RBS_PF_MINB(COUNT, THR) do { if (pf_minb == 8) RBS_PF(8, COUNT, THR); else if (pf_minb == 7) RBS_PF(7, COUNT, THR); else if (pf_minb == 5) RBS_PF(5, COUNT, THR); else RBS_PF(6, COUNT, THR); } while (0)
        if (count) { if (thr) RBS_PF_MINB(true, true); else RBS_PF_MINB(true, false); }
        else { if (thr) RBS_PF_MINB(false, true); else RBS_PF_MINB(false, false); }
#undef RBS_PF_MINB
#undef RBS_PF
        return;
    }
    switch (tuning_minb(a->substeps, RBS_ARITH_FAST)) {
        case 4: rbs::step_sphere_plane_fast_kernel<T, 4, false><<<grid, rbs::kBlock, 0, st>>>(p); break;
        case 5: rbs::step_sphere_plane_fast_kernel<T, 5, false><<<grid, rbs::kBlock, 0, st>>>(p); break;
        case 6: rbs::step_sphere_plane_fast_kernel<T, 6, false><<<grid, rbs::kBlock, 0, st>>>(p); break;
        default: rbs::step_sphere_plane_fast_kernel<T, 8, false><<<grid, rbs::kBlock, 0, st>>>(p); break;
    }
}

template <typename T> void launch_box_plane_fast(const rbs_body_plane_args *a, const Window &w) {
    const rbs::BodyPlaneParams<T> p = make_params<T>(a, w);
    if (!a->xfrc && !a->trajectory && a->substeps >= option("pf_min_substeps")) {   // fused launches: plane frame (two rotations per launch pay off)
        // resident CTAs per SM (register cap 128 / 96 / 80); measured on B200, 1M cubes fp64, 128 fused substeps:
        // 6.14e10 / 6.49e10 / 6.55e10 env-substeps/s bouncing, 4.39e10 / 4.63e10 / 4.69e10 sliding on the incline
        const int minb = (int)option("box_minb");
        const unsigned grid = blocks_for(w.cnt, rbs::kBlock);
        if (option("box_compact") != 0) {
            // CTA-level compaction of the contact path (bit-identical results; box_compact = 0 keeps one environment's
            // whole substep in its own thread).  ~29 KB of static shared memory per CTA: ask for the large carve-out.
            if (minb >= 6) {
                cudaFuncSetAttribute(rbs::step_box_plane_pfc_kernel<T, 6>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                rbs::step_box_plane_pfc_kernel<T, 6><<<grid, rbs::kBlock, 0, w.stream>>>(p);
            } else {
                cudaFuncSetAttribute(rbs::step_box_plane_pfc_kernel<T, 5>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
                rbs::step_box_plane_pfc_kernel<T, 5><<<grid, rbs::kBlock, 0, w.stream>>>(p);
            }
            return;
        }
        switch (minb) {
            case 4: rbs::step_box_plane_pf_kernel<T, 4><<<grid, rbs::kBlock, 0, w.stream>>>(p); break;
            case 5: rbs::step_box_plane_pf_kernel<T, 5><<<grid, rbs::kBlock, 0, w.stream>>>(p); break;
            default: rbs::step_box_plane_pf_kernel<T, 6><<<grid, rbs::kBlock, 0, w.stream>>>(p); break;
        }
        return;
    }
    rbs::step_box_plane_fast_kernel<T, 4><<<blocks_for(w.cnt, rbs::kBlock), rbs::kBlock, 0, w.stream>>>(p);
}

template <typename T> void launch_body_plane(const rbs_body_plane_args *a, const Window &w) {
    if (a->arith == RBS_ARITH_FAST)
        return a->geom == RBS_GEOM_SPHERE ? launch_sphere_plane_fast<T>(a, w) : launch_box_plane_fast<T>(a, w);
    if (a->geom == RBS_GEOM_SPHERE) {
        if (a->scheme == RBS_SCHEME_A) launch_body_plane_iso<T, 0, 0>(a, w);
        else launch_body_plane_iso<T, 0, 1>(a, w);
    } else {
        if (a->scheme == RBS_SCHEME_A) launch_body_plane_iso<T, 1, 0>(a, w);
        else launch_body_plane_iso<T, 1, 1>(a, w);
    }
}

int launch_body_plane_any(const rbs_body_plane_args *a, const Window &w) {
    if (a->dtype == RBS_F64) launch_body_plane<double>(a, w);
    else launch_body_plane<float>(a, w);
    return check_launch("rbs_step_body_plane");
}

int validate_body_plane(const rbs_body_plane_args *a, bool need_state) {
    if (!a) return fail(RBS_EINVAL, "rbs_step_body_plane: null args");
    if (bad_dtype(a->dtype)) return fail(RBS_EINVAL, "rbs_step_body_plane: dtype %d is not RBS_F32/RBS_F64", a->dtype);
    if (a->geom != RBS_GEOM_SPHERE && a->geom != RBS_GEOM_BOX) return fail(RBS_EINVAL, "rbs_step_body_plane: bad geom %d", a->geom);
    if (a->scheme != RBS_SCHEME_A && a->scheme != RBS_SCHEME_GENERAL) return fail(RBS_EINVAL, "rbs_step_body_plane: bad scheme %d", a->scheme);
    if (a->inertia_mode != RBS_INERTIA_GENERAL && a->inertia_mode != RBS_INERTIA_ISOTROPIC)
        return fail(RBS_EINVAL, "rbs_step_body_plane: bad inertia_mode %d", a->inertia_mode);
    if (a->arith != RBS_ARITH_STRICT && a->arith != RBS_ARITH_FAST) return fail(RBS_EINVAL, "rbs_step_body_plane: bad arith %d", a->arith);
    if (a->arith == RBS_ARITH_FAST && !(a->scheme == RBS_SCHEME_A && a->inertia_mode == RBS_INERTIA_ISOTROPIC))
        return fail(RBS_EINVAL, "rbs_step_body_plane: RBS_ARITH_FAST is implemented for scheme A + isotropic inertia only");
    if (a->n_env < 0) return fail(RBS_EINVAL, "rbs_step_body_plane: n_env %ld < 0", a->n_env);
    if (a->substeps < 1) return fail(RBS_EINVAL, "rbs_step_body_plane: substeps %d < 1", a->substeps);
    if (a->param_stride != 0 && a->param_stride < a->n_env)
        return fail(RBS_EINVAL, "rbs_step_body_plane: param_stride %ld < n_env %ld", a->param_stride, a->n_env);
    if (a->trajectory && a->trajectory_envs < 0) return fail(RBS_EINVAL, "rbs_step_body_plane: trajectory_envs %ld < 0", a->trajectory_envs);
    if (need_state) {
        if (a->n_env > 0 && !a->state) return fail(RBS_EINVAL, "rbs_step_body_plane: null state");
        if (a->stride < a->n_env) return fail(RBS_EINVAL, "rbs_step_body_plane: stride %ld < n_env %ld", a->stride, a->n_env);
    }
    if (a->inertia_mode == RBS_INERTIA_ISOTROPIC && !a->inertia &&
        !(a->inertia_u[0] == a->inertia_u[1] && a->inertia_u[1] == a->inertia_u[2]))
        return fail(RBS_EINVAL, "rbs_step_body_plane: RBS_INERTIA_ISOTROPIC needs three equal principal moments");
    return RBS_OK;
}

inline Window whole(const rbs_two_ball_args *a) { return {0, a->n_env, a->state, a->stride, as_stream(a->stream)}; }
inline Window whole(const rbs_multi_sphere_args *a) { return {0, a->n_env, a->state, a->stride, as_stream(a->stream)}; }

template <typename T> rbs::TwoBallParams<T> make_params(const rbs_two_ball_args *a, const Window &w) {
    rbs::TwoBallParams<T> p;
    p.n_env = w.cnt;
    p.stride = w.stride;
    p.pstride = a->n_env;                        // mass is [2][n_env] of the whole batch
    p.substeps = a->substeps;
    p.state = static_cast<T *>(w.state);
    p.mass = a->mass ? static_cast<const T *>(a->mass) + w.off : nullptr;
    p.radius = a->radius ? static_cast<const T *>(a->radius) + w.off : nullptr;
    p.mass_u[0] = (T)a->mass_u[0];
    p.mass_u[1] = (T)a->mass_u[1];
    p.radius_u = (T)a->radius_u;
    for (int i = 0; i < 3; ++i) p.g[i] = (T)a->gravity[i];
    p.dt = (T)a->dt;
    p.rest = (T)a->restitution;
    p.fric = (T)a->friction;
    for (int i = 0; i < 3; ++i) p.gdt[i] = p.g[i] * p.dt;
    p.neg1pe = -((T)1 + p.rest);
    p.reach_u = std::fma((T)2, p.radius_u, (T)0.01);     // the device expressions of step_two_ball_fast_kernel, in T
    p.reach2_u = (p.reach_u * p.reach_u) * (T)1.0001;
    p.n_ground = a->n_ground_hits ? a->n_ground_hits + w.off : nullptr;
    p.n_pair = a->n_pair_hits ? a->n_pair_hits + w.off : nullptr;
    return p;
}

int validate_two_ball(const rbs_two_ball_args *a, bool need_state) {
    if (!a) return fail(RBS_EINVAL, "rbs_step_two_ball: null args");
    if (bad_dtype(a->dtype)) return fail(RBS_EINVAL, "rbs_step_two_ball: dtype %d is not RBS_F32/RBS_F64", a->dtype);
    if (a->n_env < 0) return fail(RBS_EINVAL, "rbs_step_two_ball: n_env %ld < 0", a->n_env);
    if (a->substeps < 1) return fail(RBS_EINVAL, "rbs_step_two_ball: substeps %d < 1", a->substeps);
    if (a->arith != RBS_ARITH_STRICT && a->arith != RBS_ARITH_FAST) return fail(RBS_EINVAL, "rbs_step_two_ball: bad arith %d", a->arith);
    if (need_state) {
        if (a->n_env > 0 && !a->state) return fail(RBS_EINVAL, "rbs_step_two_ball: null state");
        if (a->stride < a->n_env) return fail(RBS_EINVAL, "rbs_step_two_ball: stride %ld < n_env %ld", a->stride, a->n_env);
    }
    return RBS_OK;
}

template <typename T> rbs::MultiSphereParams<T> make_params(const rbs_multi_sphere_args *a, const Window &w, int env_per_block) {
    rbs::MultiSphereParams<T> p;
    const long boff = w.off * a->n_body;         // per-body arrays are indexed env * n_body + body
    p.n_env = w.cnt;
    p.stride = w.stride;
    p.pstride = a->n_env * a->n_body;            // inertia is [3][n_env * n_body] of the whole batch
    p.substeps = a->substeps;
    p.n_body = a->n_body;
    p.env_per_block = env_per_block;
    p.state = static_cast<T *>(w.state);
    p.mass = a->mass ? static_cast<const T *>(a->mass) + boff : nullptr;
    p.inertia = a->inertia ? static_cast<const T *>(a->inertia) + boff : nullptr;
    p.radius = a->radius ? static_cast<const T *>(a->radius) + boff : nullptr;
    p.mass_u = (T)a->mass_u;
    p.radius_u = (T)a->radius_u;
    for (int i = 0; i < 3; ++i) {
        p.inertia_u[i] = (T)a->inertia_u[i];
        p.pp[i] = (T)a->plane_point[i];
        p.pn[i] = (T)a->plane_normal[i];
        p.g[i] = (T)a->gravity[i];
    }
    p.dt = (T)a->dt;
    p.rest = (T)a->restitution;
    p.fric = (T)a->friction;
    for (int i = 0; i < 3; ++i) p.gdt[i] = p.g[i] * p.dt;
    p.hdt = (T)0.5 * p.dt;
    // 0 = adaptive per CTA, starting at 50 % (or at RBS_MS_SKIN_PERCENT); > 0 = pinned; < 0 = no lists.  A launch of
    // fewer than 4 substeps (the reference's per-frame call) cannot amortise a list and scans every substep.
    const int default_skin = option("ms_skin_percent") > 0 ? (int)option("ms_skin_percent") : 50;
    int skin_pct = a->list_skin_percent != 0 ? a->list_skin_percent : default_skin;
    if (a->list_skin_percent == 0 && a->substeps < 4) skin_pct = -1;
    p.skin = skin_pct > 0 ? (T)(skin_pct * 0.01) : T(0);
    p.skin_adapt = a->list_skin_percent == 0 && skin_pct > 0;
    p.walk_cost = (int)option("ms_walk_cost");
    p.tight_span = (int)option("ms_tight_span");
    {
        double R[9], q[4], gpf[3];
        plane_frame(a->plane_normal, a->gravity, a->dt, R, q, gpf);
        for (int i = 0; i < 9; ++i) p.frame[i] = (T)R[i];
        for (int i = 0; i < 4; ++i) p.frame_q[i] = (T)q[i];
        for (int i = 0; i < 3; ++i) p.gdt_pf[i] = (T)gpf[i];
    }
    p.n_contacts = a->n_contacts ? a->n_contacts + boff : nullptr;
    p.n_impulses = a->n_impulses ? a->n_impulses + boff : nullptr;
    return p;
}

int validate_multi_sphere(const rbs_multi_sphere_args *a, bool need_state) {
    if (!a) return fail(RBS_EINVAL, "rbs_step_multi_sphere: null args");
    if (bad_dtype(a->dtype)) return fail(RBS_EINVAL, "rbs_step_multi_sphere: dtype %d is not RBS_F32/RBS_F64", a->dtype);
    if (a->n_body < 1 || a->n_body > 1024) return fail(RBS_EINVAL, "rbs_step_multi_sphere: n_body %d outside 1..1024", a->n_body);
    if (a->n_env < 0) return fail(RBS_EINVAL, "rbs_step_multi_sphere: n_env %ld < 0", a->n_env);
    if (a->substeps < 1) return fail(RBS_EINVAL, "rbs_step_multi_sphere: substeps %d < 1", a->substeps);
    if (a->inertia_mode != RBS_INERTIA_GENERAL && a->inertia_mode != RBS_INERTIA_ISOTROPIC)
        return fail(RBS_EINVAL, "rbs_step_multi_sphere: bad inertia_mode %d", a->inertia_mode);
    if (a->arith != RBS_ARITH_STRICT && a->arith != RBS_ARITH_FAST) return fail(RBS_EINVAL, "rbs_step_multi_sphere: bad arith %d", a->arith);
    if (a->arith == RBS_ARITH_FAST && a->inertia_mode != RBS_INERTIA_ISOTROPIC)
        return fail(RBS_EINVAL, "rbs_step_multi_sphere: RBS_ARITH_FAST needs RBS_INERTIA_ISOTROPIC");
    if (need_state) {
        if (a->n_env > 0 && !a->state) return fail(RBS_EINVAL, "rbs_step_multi_sphere: null state");
        if (a->stride < a->n_env * a->n_body)
            return fail(RBS_EINVAL, "rbs_step_multi_sphere: stride %ld < n_env*n_body %ld", a->stride, a->n_env * a->n_body);
    }
    return RBS_OK;
}

inline void multi_sphere_shape(int B, int *threads, int *epb) {
    int t = ((B + 31) / 32) * 32;
    if (t < 128) t = 128;
    *threads = t;
    *epb = t / B;
}

template <typename T> int launch_multi_sphere(const rbs_multi_sphere_args *a, const Window &w) {
    const int B = a->n_body;
    int threads, epb;
    multi_sphere_shape(B, &threads, &epb);
    const rbs::MultiSphereParams<T> p = make_params<T>(a, w, epb);
    const unsigned grid = (unsigned)((w.cnt + epb - 1) / epb);
    // two buffers of centres + fp32 relative copies + partner lists (ceil(B/64) words per thread, word-major)
    const size_t smem = 2 * (size_t)epb * B * 4 * sizeof(T) + (size_t)epb * B * sizeof(float4) +
                        (size_t)((B + 63) / 64) * threads * sizeof(unsigned long long);
    cudaStream_t st = w.stream;
    const bool iso = a->inertia_mode == RBS_INERTIA_ISOTROPIC;
    // above 48 KB (B > ~450) the dynamic shared memory needs the opt-in attribute.  It is a property of the function on
    // the CURRENT device, so it is set on every such launch (cheap) rather than remembered per process.
#define RBS_MS_LAUNCH(KERNEL)                                                                                   \
    do {                                                                                                        \
        if (smem > 48 * 1024) {                                                                                 \
            const cudaError_t e__ = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e__ != cudaSuccess) {                                                                           \
                cudaGetLastError();                                                                             \
                return fail(RBS_ECUDA, "rbs_step_multi_sphere: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e__)); \
            }                                                                                                   \
        }                                                                                                       \
        KERNEL<<<grid, threads, smem, st>>>(p);                                                                 \
    } while (0)
    if (a->arith == RBS_ARITH_FAST && option("ms_kernel") >= 2) {
        // plane-frame kernel (SoA centres in both precisions, two-phase list walk); MU0: frictionless spheres
        const size_t smem_pf = rbs::PairListsSoA<T>::smem_bytes(epb, B, threads);
        const size_t smem_old = smem;
        (void)smem_old;
#define RBS_MS_LAUNCH_PF(KERNEL)                                                                                \
    do {                                                                                                        \
        if (smem_pf > 48 * 1024) {                                                                              \
            const cudaError_t e__ = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pf); \
            if (e__ != cudaSuccess) {                                                                           \
                cudaGetLastError();                                                                             \
                return fail(RBS_ECUDA, "rbs_step_multi_sphere: %zu bytes of shared memory: %s", smem_pf, cudaGetErrorString(e__)); \
            }                                                                                                   \
        }                                                                                                       \
        KERNEL<<<grid, threads, smem_pf, st>>>(p);                                                              \
    } while (0)
        const bool mu0 = a->friction == 0.0;
        if (threads <= 256) {
            if (mu0 && option("ms_regs") >= 128) RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 256, true, 128>));
            else if (mu0) RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 256, true>));
            else RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 256, false>));
        }
        else if (threads <= 512) { if (mu0) RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 512, true>)); else RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 512, false>)); }
        else { if (mu0) RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 1024, true>)); else RBS_MS_LAUNCH_PF((rbs::step_multi_sphere_pf_kernel<T, 1024, false>)); }
#undef RBS_MS_LAUNCH_PF
        return RBS_OK;
    }
    if (a->arith == RBS_ARITH_FAST) {
        if (threads <= 256) RBS_MS_LAUNCH((rbs::step_multi_sphere_fast_kernel<T, 256>));
        else if (threads <= 512) RBS_MS_LAUNCH((rbs::step_multi_sphere_fast_kernel<T, 512>));
        else RBS_MS_LAUNCH((rbs::step_multi_sphere_fast_kernel<T, 1024>));
        return RBS_OK;
    }
    if (threads <= 256) {
        // register cap of the 256-thread class (option strict_ms_regs; 0 = measured best, profiles/r2_ab_strict.jsonl)
        int regs = (int)option("strict_ms_regs");
        if (regs == 0) regs = 128;
        if (iso) RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 1, 256>));
        else if (regs <= 96) RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 0, 256, 96>));
        else if (regs <= 128) RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 0, 256, 128>));
        else RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 0, 256>));
    } else if (threads <= 512) {
        if (iso) RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 1, 512>));
        else RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 0, 512>));
    } else {
        if (iso) RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 1, 1024>));
        else RBS_MS_LAUNCH((rbs::step_multi_sphere_kernel<T, 0, 1024>));
    }
#undef RBS_MS_LAUNCH
    return RBS_OK;
}

inline Window whole(const rbs_multi_body_args *a) { return {0, a->n_env, a->state, a->stride, as_stream(a->stream)}; }

template <typename T> int launch_multi_body(const rbs_multi_body_args *a, const Window &w) {
    rbs::MultiBodyParams<T> p;
    const int B = a->n_body;
    p.n_env = w.cnt;
    p.stride = w.stride;
    p.substeps = a->substeps;
    p.n_body = B;
    p.env_per_block = 256 / B;
    p.has_offset = a->has_offset;
    p.state = static_cast<T *>(w.state);
    p.table = static_cast<const T *>(a->body_table);
    for (int i = 0; i < 3; ++i) {
        p.pp[i] = (T)a->plane_point[i];
        p.pn[i] = (T)a->plane_normal[i];
        p.g[i] = (T)a->gravity[i];
    }
    p.dt = (T)a->dt;
    p.rest = (T)a->restitution;
    p.fric = (T)a->friction;
    p.n_contacts = a->n_contacts ? a->n_contacts + w.off * B : nullptr;    // counters are indexed env * n_body + body
    p.n_impulses = a->n_impulses ? a->n_impulses + w.off * B : nullptr;
    const int threads = ((p.env_per_block * B + 31) / 32) * 32;
    const size_t smem = ((size_t)B * rbs::kBodyTable + 2 * (size_t)p.env_per_block * B * 12) * sizeof(T);
    // resident CTAs per SM (option mb_minb: 1 = uncapped registers, 2 = 128, 3 = 80; 0 = tuned default)
    int minb = (int)option("mb_minb");
    if (minb == 0) minb = 2;                                 // profiles/r2_multi_body.jsonl
#define RBS_MB_LAUNCH(KERNEL)                                                                                              \
    do {                                                                                                                   \
        if (smem > 48 * 1024) {                                                                                            \
            cudaError_t e__ = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);        \
            if (e__ != cudaSuccess)                                                                                        \
                return fail(RBS_ECUDA, "rbs_step_multi_body: %zu bytes of shared memory: %s", smem, cudaGetErrorString(e__)); \
        }                                                                                                                  \
        KERNEL<<<blocks_for(w.cnt, p.env_per_block), threads, smem, w.stream>>>(p);                                        \
    } while (0)
    if (minb >= 3) RBS_MB_LAUNCH((rbs::step_multi_body_kernel<T, 256, 3>));
    else if (minb == 2) RBS_MB_LAUNCH((rbs::step_multi_body_kernel<T, 256, 2>));
    else RBS_MB_LAUNCH((rbs::step_multi_body_kernel<T, 256, 1>));
#undef RBS_MB_LAUNCH
    return check_launch("rbs_step_multi_body");
}

int validate_multi_body(const rbs_multi_body_args *a, bool need_state) {
    if (!a) return fail(RBS_EINVAL, "rbs_step_multi_body: null args");
    if (bad_dtype(a->dtype)) return fail(RBS_EINVAL, "rbs_step_multi_body: dtype %d is not RBS_F32/RBS_F64", a->dtype);
    if (a->n_body < 1 || a->n_body > 256) return fail(RBS_EINVAL, "rbs_step_multi_body: n_body %d outside 1..256", a->n_body);
    if (a->n_env < 0) return fail(RBS_EINVAL, "rbs_step_multi_body: n_env %ld < 0", a->n_env);
    if (a->substeps < 1) return fail(RBS_EINVAL, "rbs_step_multi_body: substeps %d < 1", a->substeps);
    if (!a->body_table) return fail(RBS_EINVAL, "rbs_step_multi_body: null body_table");
    if (need_state) {
        if (a->n_env > 0 && !a->state) return fail(RBS_EINVAL, "rbs_step_multi_body: null state");
        if (a->stride < a->n_env * a->n_body)
            return fail(RBS_EINVAL, "rbs_step_multi_body: stride %ld < n_env*n_body %ld", a->stride, a->n_env * a->n_body);
    }
    return RBS_OK;
}

int launch_multi_body_any(const rbs_multi_body_args *a, const Window &w) {
    return a->dtype == RBS_F64 ? launch_multi_body<double>(a, w) : launch_multi_body<float>(a, w);
}

int launch_two_ball_any(const rbs_two_ball_args *a, const Window &w) {
    const unsigned grid = blocks_for(w.cnt, rbs::kBlock);
    cudaStream_t st = w.stream;
    if (a->arith == RBS_ARITH_FAST) {
        // gravity along z only (every shipped model): the additions of +0.0 to the horizontal velocities are not issued
        const bool gz = a->gravity[0] == 0.0 && a->gravity[1] == 0.0;
        // resident CTAs per SM (register cap 64 / 80 / 96); spins and per-ball constants live in shared memory
        int minb = (int)option("tb_minb");
        if (minb == 0) minb = (a->dtype == RBS_F64 && (a->radius != nullptr || option("tb_uniform") == 0)) ? 6 : 8;   // measured on B200, 1M envs: profiles/r2_ab_two_ball.jsonl, r2_ab_two_ball_uniform_radius.jsonl
#define RBS_TB_UR(T, GZ, MINB, UR)                                                                                       \
    do {                                                                                                                 \
        cudaFuncSetAttribute(rbs::step_two_ball_fast_kernel<T, GZ, MINB, UR>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                             cudaSharedmemCarveoutMaxShared);                                                            \
        rbs::step_two_ball_fast_kernel<T, GZ, MINB, UR><<<grid, rbs::kBlock, 0, st>>>(make_params<T>(a, w));                \
    } while (0)
        // uniform radius (no per-environment array): radius / reach / reach^2 as launch parameters (option tb_uniform = 0 keeps registers)
        const bool ur = a->radius == nullptr && option("tb_uniform") != 0;
#define RBS_TB(T, GZ, MINB) do { if (ur) RBS_TB_UR(T, GZ, MINB, true); else RBS_TB_UR(T, GZ, MINB, false); } while (0)
#define RBS_TB_MINB(T, GZ) do { if (minb >= 8) RBS_TB(T, GZ, 8); else if (minb >= 6) RBS_TB(T, GZ, 6); else RBS_TB(T, GZ, 5); } while (0)
#define RBS_TB2(GZ, MINB)                                                                                                \
    do {                                                                                                                 \
        cudaFuncSetAttribute(rbs::step_two_ball_fast2_kernel<GZ, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout,    \
                             cudaSharedmemCarveoutMaxShared);                                                            \
        rbs::step_two_ball_fast2_kernel<GZ, MINB><<<blocks_for(w.cnt, 2 * rbs::kBlock), rbs::kBlock, 0, st>>>(make_params<float>(a, w)); \
    } while (0)
#define RBS_TB2_MINB(GZ) do { if (minb >= 8) RBS_TB2(GZ, 8); else if (minb >= 6) RBS_TB2(GZ, 6); else RBS_TB2(GZ, 4); } while (0)
        if (a->dtype == RBS_F64) { if (gz) RBS_TB_MINB(double, true); else RBS_TB_MINB(double, false); }
        else if (option("tb_packed") != 0) {   // two environments per thread on packed fp32x2 instructions (same bits as the scalar kernel)
            if (option("tb_minb") == 0) minb = 6;
            if (gz) RBS_TB2_MINB(true); else RBS_TB2_MINB(false);
        }
        else { if (gz) RBS_TB_MINB(float, true); else RBS_TB_MINB(float, false); }
#undef RBS_TB2_MINB
#undef RBS_TB2
#undef RBS_TB_MINB
#undef RBS_TB
#undef RBS_TB_UR
    } else {
        // strict policy: resident CTAs per SM (option strict_tb_minb: 3 = uncapped, 4 = 128 registers, 5 = 96; 0 = measured best)
        int minb = (int)option("strict_tb_minb");
        if (minb == 0) minb = 5;               // profiles/r2_ab_strict.jsonl: 9.1e10 (uncapped) -> 1.03e11 (128) -> 1.08e11 (96 registers)
        if (a->dtype == RBS_F64) {
            if (minb >= 5) rbs::step_two_ball_kernel<double, 5><<<grid, rbs::kBlock, 0, st>>>(make_params<double>(a, w));
            else if (minb == 4) rbs::step_two_ball_kernel<double, 4><<<grid, rbs::kBlock, 0, st>>>(make_params<double>(a, w));
            else rbs::step_two_ball_kernel<double, 3><<<grid, rbs::kBlock, 0, st>>>(make_params<double>(a, w));
        } else {
            if (minb >= 5) rbs::step_two_ball_kernel<float, 5><<<grid, rbs::kBlock, 0, st>>>(make_params<float>(a, w));
            else if (minb == 4) rbs::step_two_ball_kernel<float, 4><<<grid, rbs::kBlock, 0, st>>>(make_params<float>(a, w));
            else rbs::step_two_ball_kernel<float, 3><<<grid, rbs::kBlock, 0, st>>>(make_params<float>(a, w));
        }
    }
    return check_launch("rbs_step_two_ball");
}

int launch_multi_sphere_any(const rbs_multi_sphere_args *a, const Window &w) {
    const int rc = a->dtype == RBS_F64 ? launch_multi_sphere<double>(a, w) : launch_multi_sphere<float>(a, w);
    if (rc) return rc;
    return check_launch("rbs_step_multi_sphere");
}

// Host-buffer driver resources, ONE SET PER DEVICE: a cached device workspace and the streams / events of the chunk
// pipeline, created at the first host-buffer call made with that device current.  Calls on different devices share
// nothing and run concurrently; calls on the same device are serialised by that device's mutex (they would contend for
// the same copy engines anyway).  Nothing else in the library keeps state between calls.
constexpr int kMaxChunks = 32;
constexpr int kMaxComputeStreams = 4;
constexpr int kMaxDevices = 64;
struct Pipe {
    std::mutex mutex;
    bool ready = false;
    int sm_count = 148;
    void *ws = nullptr;
    size_t ws_bytes = 0;
    cudaStream_t in = nullptr, out = nullptr, compute[kMaxComputeStreams] = {};
    cudaEvent_t start = nullptr, finished = nullptr, arrived[kMaxChunks] = {}, stepped[kMaxChunks] = {};
};
Pipe g_pipes[kMaxDevices];

int workspace(Pipe &pp, size_t bytes, void **out) {
    if (bytes > pp.ws_bytes) {
        if (pp.ws) cudaFree(pp.ws);
        pp.ws = nullptr;
        pp.ws_bytes = 0;
        cudaError_t err = cudaMalloc(&pp.ws, bytes);
        if (err != cudaSuccess) {
            cudaGetLastError();
            return fail(err == cudaErrorMemoryAllocation ? RBS_ENOMEM : RBS_ECUDA, "workspace of %zu bytes: %s", bytes,
                        cudaGetErrorString(err));
        }
        pp.ws_bytes = bytes;
    }
    *out = pp.ws;
    return RBS_OK;
}

// the current device's pipeline (nullptr + error when there is no usable device)
Pipe *current_pipe() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        fail(RBS_ECUDA, "no CUDA device");
        return nullptr;
    }
    if (dev < 0 || dev >= kMaxDevices) {
        fail(RBS_EINVAL, "device ordinal %d outside 0..%d", dev, kMaxDevices - 1);
        return nullptr;
    }
    return &g_pipes[dev];
}

// call with pp.mutex held
int pipe_init(Pipe &pp) {
    if (pp.ready) return RBS_OK;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaError_t err = cudaSuccess;
    auto ok = [&](cudaError_t e) { if (err == cudaSuccess) err = e; };
    ok(cudaDeviceGetAttribute(&pp.sm_count, cudaDevAttrMultiProcessorCount, dev));
    ok(cudaStreamCreateWithFlags(&pp.in, cudaStreamNonBlocking));
    ok(cudaStreamCreateWithFlags(&pp.out, cudaStreamNonBlocking));
    for (int i = 0; i < kMaxComputeStreams; ++i) ok(cudaStreamCreateWithFlags(&pp.compute[i], cudaStreamNonBlocking));
    ok(cudaEventCreateWithFlags(&pp.start, cudaEventDisableTiming));
    ok(cudaEventCreateWithFlags(&pp.finished, cudaEventDisableTiming));
    for (int i = 0; i < kMaxChunks; ++i) {
        ok(cudaEventCreateWithFlags(&pp.arrived[i], cudaEventDisableTiming));
        ok(cudaEventCreateWithFlags(&pp.stepped[i], cudaEventDisableTiming));
    }
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(RBS_ECUDA, "pipeline resources: %s", cudaGetErrorString(err));
    }
    pp.ready = true;
    return RBS_OK;
}

// argument checks shared by the host-buffer drivers; 1 = nothing to do
int host_driver_prologue(long n_env, const void *qpos_host, const void *qvel_host, long total_steps) {
    if (total_steps < 0) return fail(RBS_EINVAL, "total_steps %ld < 0", total_steps);
    if (n_env == 0) return 1;
    if (!qpos_host || !qvel_host) return fail(RBS_EINVAL, "null host buffer");
    return RBS_OK;
}

#define RBS_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t err__ = (call);                                                      \
        if (err__ != cudaSuccess) return fail(RBS_ECUDA, #call ": %s", cudaGetErrorString(err__)); \
    } while (0)

// The host-buffer drivers, pipelined over chunks of environments: while chunk c is being stepped, chunk c+1 is on its
// way in over PCIe and chunk c-1 on its way out (three internal streams; the caller's stream is joined at the end).
// Chunks are whole numbers of `quantum` environments (one wave of resident CTAs of the stepping kernel), so that no
// chunk ends in a ragged partial wave.  Only the first copy-in and the last copy-out are exposed, so the first and the
// last chunk are one quantum; the ones in between share the rest (at most `host_chunks` chunks, option of that name).
// `launch(local_args, window)` enqueues one launch of local_args->substeps substeps on the window's stream.
template <typename Args, typename LaunchFn>
int run_host_pipelined(const Args *a, int n_body, int body_fastest, long envs_per_sm_wave, void *qpos_host, void *qvel_host,
                       long total_steps, LaunchFn launch) {
    int rc = host_driver_prologue(a->n_env, qpos_host, qvel_host, total_steps);
    if (rc) return rc < 0 ? rc : RBS_OK;
    Pipe *pipe = current_pipe();
    if (!pipe) return RBS_ECUDA;
    Pipe &g_pipe = *pipe;
    std::lock_guard<std::mutex> lock(g_pipe.mutex);
    rc = pipe_init(g_pipe);
    if (rc) return rc;
    long quantum = (long)g_pipe.sm_count * envs_per_sm_wave;
    const size_t es = elem_size(a->dtype);
    const long E = a->n_env;
    const size_t per_env = (size_t)n_body * es;              // bytes of one scalar row entry group per environment
    void *base = nullptr;
    rc = workspace(g_pipe, (size_t)E * 26 * per_env, &base);
    if (rc) return rc;
    char *qpos_d = static_cast<char *>(base), *qvel_d = qpos_d + (size_t)E * 7 * per_env, *state_d = qvel_d + (size_t)E * 6 * per_env;
    long max_chunks = option("host_chunks");
    if (max_chunks < 3) max_chunks = 3;
    if (max_chunks > kMaxChunks) max_chunks = kMaxChunks;
    if (quantum < 1) quantum = 1;
    // the two exposed chunks (first copy-in, last copy-out) are half a wave: the pipeline fills and drains sooner
    long edge = (long)g_pipe.sm_count * (envs_per_sm_wave >= 2 ? envs_per_sm_wave / 2 : 1);
    long offs[kMaxChunks + 1];
    int n_chunks = 0;
    offs[0] = 0;
    if (E <= 3 * quantum) {
        offs[++n_chunks] = E;
    } else {
        offs[++n_chunks] = edge;
        const long middle = E - 2 * edge;
        const long n_mid = max_chunks - 2;
        long per = (middle + n_mid - 1) / n_mid;
        per = ((per + quantum - 1) / quantum) * quantum;
        for (long done = 0; done < middle; done += per) offs[n_chunks + 1] = offs[n_chunks] + (middle - done < per ? middle - done : per), ++n_chunks;
        offs[n_chunks + 1] = E, ++n_chunks;
    }
    long n_streams = option("host_streams");                 // chunks in flight on the GPU at once
    if (n_streams < 1) n_streams = 1;
    if (n_streams > kMaxComputeStreams) n_streams = kMaxComputeStreams;
    cudaStream_t user = as_stream(a->stream);
    RBS_CUDA(cudaEventRecord(g_pipe.start, user));
    RBS_CUDA(cudaStreamWaitEvent(g_pipe.in, g_pipe.start, 0));
    for (int c = 0; c < n_chunks; ++c) {
        const size_t off = (size_t)offs[c], cnt = (size_t)(offs[c + 1] - offs[c]);
        RBS_CUDA(cudaMemcpyAsync(qpos_d + off * 7 * per_env, (char *)qpos_host + off * 7 * per_env, cnt * 7 * per_env, cudaMemcpyHostToDevice, g_pipe.in));
        RBS_CUDA(cudaMemcpyAsync(qvel_d + off * 6 * per_env, (char *)qvel_host + off * 6 * per_env, cnt * 6 * per_env, cudaMemcpyHostToDevice, g_pipe.in));
        RBS_CUDA(cudaEventRecord(g_pipe.arrived[c], g_pipe.in));
    }
    for (int c = 0; c < n_chunks; ++c) {
        const long off = offs[c], cnt = offs[c + 1] - offs[c];
        cudaStream_t cs = g_pipe.compute[c % n_streams];
        char *state_c = state_d + (size_t)off * 13 * per_env;
        char *qp_c = qpos_d + (size_t)off * 7 * per_env, *qv_c = qvel_d + (size_t)off * 6 * per_env;
        const long stride_c = body_fastest ? cnt * n_body : cnt;
        RBS_CUDA(cudaStreamWaitEvent(cs, g_pipe.arrived[c], 0));
        rc = rbs_pack_state(a->dtype, cnt, n_body, body_fastest, qp_c, qv_c, state_c, stride_c, cs);
        if (rc) return rc;
        Args local = *a;
        const Window w{off, cnt, state_c, stride_c, cs};
        for (long done = 0; done < total_steps;) {
            const long k = total_steps - done < a->substeps ? total_steps - done : a->substeps;
            local.substeps = (int)k;
            rc = launch(&local, w);
            if (rc) return rc;
            done += k;
        }
        rc = rbs_unpack_state(a->dtype, cnt, n_body, body_fastest, state_c, stride_c, qp_c, qv_c, cs);
        if (rc) return rc;
        RBS_CUDA(cudaEventRecord(g_pipe.stepped[c], cs));
        RBS_CUDA(cudaStreamWaitEvent(g_pipe.out, g_pipe.stepped[c], 0));
        RBS_CUDA(cudaMemcpyAsync((char *)qpos_host + (size_t)off * 7 * per_env, qp_c, (size_t)cnt * 7 * per_env, cudaMemcpyDeviceToHost, g_pipe.out));
        RBS_CUDA(cudaMemcpyAsync((char *)qvel_host + (size_t)off * 6 * per_env, qv_c, (size_t)cnt * 6 * per_env, cudaMemcpyDeviceToHost, g_pipe.out));
    }
    RBS_CUDA(cudaEventRecord(g_pipe.finished, g_pipe.out));
    RBS_CUDA(cudaStreamWaitEvent(user, g_pipe.finished, 0));
    RBS_CUDA(cudaStreamSynchronize(user));
    return RBS_OK;
}

}  // namespace

extern "C" {

int rbs_version(void) { return RBS_ABI_VERSION; }
const char *rbs_last_error(void) { return g_err; }
unsigned long long rbs_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int rbs_set_option(const char *name, long value) {
    Option *o = find_option(name);
    if (!o) return fail(RBS_EINVAL, "rbs_set_option: unknown option '%s'", name ? name : "(null)");
    o->value.store(value, std::memory_order_relaxed);
    return RBS_OK;
}

long rbs_get_option(const char *name) {
    Option *o = find_option(name);
    if (!o) {
        fail(RBS_EINVAL, "rbs_get_option: unknown option '%s'", name ? name : "(null)");
        return LONG_MIN;
    }
    return o->value.load(std::memory_order_relaxed);
}

int rbs_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int rbs_impulse_friction(int dtype, long n, const void *mass, double mass_u, const void *vel, const void *omega,
                         const void *contact_point, const void *normal, const void *restitution,
                         double restitution_u, const void *friction, double friction_u, void *out_jn, void *out_jt,
                         unsigned char *out_flag, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_impulse_friction: bad dtype %d", dtype);
    if (n < 0) return fail(RBS_EINVAL, "rbs_impulse_friction: n %ld < 0", n);
    if (n == 0) return RBS_OK;
    if (!vel || !omega || !contact_point || !normal || !out_jn || !out_jt) return fail(RBS_EINVAL, "rbs_impulse_friction: null array");
    const unsigned grid = blocks_for(n, 256);
    if (dtype == RBS_F64)
        rbs::impulse_friction_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const double *)mass, mass_u, (const double *)vel, (const double *)omega, (const double *)contact_point,
            (const double *)normal, (const double *)restitution, restitution_u, (const double *)friction, friction_u,
            (double *)out_jn, (double *)out_jt, out_flag);
    else
        rbs::impulse_friction_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const float *)mass, (float)mass_u, (const float *)vel, (const float *)omega, (const float *)contact_point,
            (const float *)normal, (const float *)restitution, (float)restitution_u, (const float *)friction,
            (float)friction_u, (float *)out_jn, (float *)out_jt, out_flag);
    return check_launch("rbs_impulse_friction");
}

int rbs_apply_impulse_friction(int dtype, long n, const void *vel, const void *omega, const void *mass, double mass_u,
                               const void *inertia_world, const void *contact_point, const void *normal,
                               const void *jn, const void *jt, void *out_vel, void *out_omega, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_apply_impulse_friction: bad dtype %d", dtype);
    if (n < 0) return fail(RBS_EINVAL, "rbs_apply_impulse_friction: n %ld < 0", n);
    if (n == 0) return RBS_OK;
    if (!vel || !omega || !inertia_world || !contact_point || !normal || !jn || !jt || !out_vel || !out_omega)
        return fail(RBS_EINVAL, "rbs_apply_impulse_friction: null array");
    const unsigned grid = blocks_for(n, 128);
    if (dtype == RBS_F64)
        rbs::apply_impulse_kernel<double, 1><<<grid, 128, 0, as_stream(stream)>>>(
            n, (const double *)vel, (const double *)omega, (const double *)mass, mass_u, (const double *)inertia_world,
            (const double *)contact_point, (const double *)normal, (const double *)jn, 0.0, (const double *)jt,
            (double *)out_vel, (double *)out_omega);
    else
        rbs::apply_impulse_kernel<float, 1><<<grid, 128, 0, as_stream(stream)>>>(
            n, (const float *)vel, (const float *)omega, (const float *)mass, (float)mass_u, (const float *)inertia_world,
            (const float *)contact_point, (const float *)normal, (const float *)jn, 0.0f, (const float *)jt,
            (float *)out_vel, (float *)out_omega);
    return check_launch("rbs_apply_impulse_friction");
}

int rbs_apply_impulse(int dtype, long n, const void *vel, const void *omega, const void *mass, double mass_u,
                      const void *inertia_world, const void *contact_point, const void *normal, const void *impulse,
                      double impulse_u, void *out_vel, void *out_omega, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_apply_impulse: bad dtype %d", dtype);
    if (n < 0) return fail(RBS_EINVAL, "rbs_apply_impulse: n %ld < 0", n);
    if (n == 0) return RBS_OK;
    if (!vel || !omega || !inertia_world || !contact_point || !normal || !out_vel || !out_omega)
        return fail(RBS_EINVAL, "rbs_apply_impulse: null array");
    const unsigned grid = blocks_for(n, 128);
    if (dtype == RBS_F64)
        rbs::apply_impulse_kernel<double, 0><<<grid, 128, 0, as_stream(stream)>>>(
            n, (const double *)vel, (const double *)omega, (const double *)mass, mass_u, (const double *)inertia_world,
            (const double *)contact_point, (const double *)normal, (const double *)impulse, impulse_u, nullptr,
            (double *)out_vel, (double *)out_omega);
    else
        rbs::apply_impulse_kernel<float, 0><<<grid, 128, 0, as_stream(stream)>>>(
            n, (const float *)vel, (const float *)omega, (const float *)mass, (float)mass_u, (const float *)inertia_world,
            (const float *)contact_point, (const float *)normal, (const float *)impulse, (float)impulse_u, nullptr,
            (float *)out_vel, (float *)out_omega);
    return check_launch("rbs_apply_impulse");
}

int rbs_inertia_world(int dtype, long n, const void *inertia_diag, const void *quat, void *out, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_inertia_world: bad dtype %d", dtype);
    if (n < 0) return fail(RBS_EINVAL, "rbs_inertia_world: n %ld < 0", n);
    if (n == 0) return RBS_OK;
    if (!inertia_diag || !quat || !out) return fail(RBS_EINVAL, "rbs_inertia_world: null array");
    const unsigned grid = blocks_for(n, 256);
    if (dtype == RBS_F64)
        rbs::inertia_world_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n, (const double *)inertia_diag, (const double *)quat, (double *)out);
    else
        rbs::inertia_world_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n, (const float *)inertia_diag, (const float *)quat, (float *)out);
    return check_launch("rbs_inertia_world");
}

int rbs_two_ball_impulse(int dtype, long n, const void *mass, double mass_u, const void *inv_inertia,
                         double inv_inertia_u, const void *v_lin, const void *v_ang, const void *r, const void *normal,
                         const void *restitution, double restitution_u, const void *friction, double friction_u,
                         void *out_J, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_two_ball_impulse: bad dtype %d", dtype);
    if (n < 0) return fail(RBS_EINVAL, "rbs_two_ball_impulse: n %ld < 0", n);
    if (n == 0) return RBS_OK;
    if (!v_lin || !v_ang || !r || !normal || !out_J) return fail(RBS_EINVAL, "rbs_two_ball_impulse: null array");
    const unsigned grid = blocks_for(n, 256);
    if (dtype == RBS_F64)
        rbs::two_ball_impulse_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const double *)mass, mass_u, (const double *)inv_inertia, inv_inertia_u, (const double *)v_lin,
            (const double *)v_ang, (const double *)r, (const double *)normal, (const double *)restitution, restitution_u,
            (const double *)friction, friction_u, (double *)out_J);
    else
        rbs::two_ball_impulse_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(
            n, (const float *)mass, (float)mass_u, (const float *)inv_inertia, (float)inv_inertia_u, (const float *)v_lin,
            (const float *)v_ang, (const float *)r, (const float *)normal, (const float *)restitution, (float)restitution_u,
            (const float *)friction, (float)friction_u, (float *)out_J);
    return check_launch("rbs_two_ball_impulse");
}

int rbs_step_body_plane(const rbs_body_plane_args *a) {
    int rc = validate_body_plane(a, true);
    if (rc) return rc;
    if (a->n_env == 0) return RBS_OK;
    return launch_body_plane_any(a, whole(a));
}

int rbs_step_two_ball(const rbs_two_ball_args *a) {
    int rc = validate_two_ball(a, true);
    if (rc) return rc;
    if (a->n_env == 0) return RBS_OK;
    return launch_two_ball_any(a, whole(a));
}

int rbs_step_multi_sphere(const rbs_multi_sphere_args *a) {
    int rc = validate_multi_sphere(a, true);
    if (rc) return rc;
    if (a->n_env == 0) return RBS_OK;
    return launch_multi_sphere_any(a, whole(a));
}

int rbs_step_multi_body(const rbs_multi_body_args *a) {
    int rc = validate_multi_body(a, true);
    if (rc) return rc;
    if (a->n_env == 0) return RBS_OK;
    return launch_multi_body_any(a, whole(a));
}

int rbs_pack_state(int dtype, long n_env, int n_body, int body_fastest, const void *qpos, const void *qvel,
                   void *state, long stride, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_pack_state: bad dtype %d", dtype);
    if (n_env < 0 || n_body < 1) return fail(RBS_EINVAL, "rbs_pack_state: bad sizes");
    if (n_env == 0) return RBS_OK;
    if (!qpos || !qvel || !state) return fail(RBS_EINVAL, "rbs_pack_state: null array");
    if (stride < (body_fastest ? n_env * n_body : n_env)) return fail(RBS_EINVAL, "rbs_pack_state: stride too small");
    const unsigned grid = blocks_for(n_env * n_body, 256);
    if (dtype == RBS_F64)
        rbs::convert_state_kernel<double, 1><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (double *)qpos, (double *)qvel, (double *)state, stride);
    else
        rbs::convert_state_kernel<float, 1><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (float *)qpos, (float *)qvel, (float *)state, stride);
    return check_launch("rbs_pack_state");
}

int rbs_unpack_state(int dtype, long n_env, int n_body, int body_fastest, const void *state, long stride, void *qpos,
                     void *qvel, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_unpack_state: bad dtype %d", dtype);
    if (n_env < 0 || n_body < 1) return fail(RBS_EINVAL, "rbs_unpack_state: bad sizes");
    if (n_env == 0) return RBS_OK;
    if (!qpos || !qvel || !state) return fail(RBS_EINVAL, "rbs_unpack_state: null array");
    if (stride < (body_fastest ? n_env * n_body : n_env)) return fail(RBS_EINVAL, "rbs_unpack_state: stride too small");
    const unsigned grid = blocks_for(n_env * n_body, 256);
    if (dtype == RBS_F64)
        rbs::convert_state_kernel<double, 0><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (double *)qpos, (double *)qvel, (double *)state, stride);
    else
        rbs::convert_state_kernel<float, 0><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (float *)qpos, (float *)qvel, (float *)state, stride);
    return check_launch("rbs_unpack_state");
}

int rbs_reset_envs(int dtype, long n_env, int n_body, int body_fastest, void *state, long stride, const void *qpos0,
                   const unsigned char *env_mask, unsigned *n_contacts, unsigned *n_impulses, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_reset_envs: bad dtype %d", dtype);
    if (n_env < 0 || n_body < 1) return fail(RBS_EINVAL, "rbs_reset_envs: bad sizes");
    if (n_env == 0) return RBS_OK;
    if (!state || !qpos0) return fail(RBS_EINVAL, "rbs_reset_envs: null array");
    if (stride < (body_fastest ? n_env * n_body : n_env)) return fail(RBS_EINVAL, "rbs_reset_envs: stride too small");
    const unsigned grid = blocks_for(n_env * n_body, 256);
    if (dtype == RBS_F64)
        rbs::reset_envs_kernel<double><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (double *)state, stride,
                                                                             (const double *)qpos0, env_mask, n_contacts, n_impulses);
    else
        rbs::reset_envs_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(n_env, n_body, body_fastest, (float *)state, stride,
                                                                            (const float *)qpos0, env_mask, n_contacts, n_impulses);
    return check_launch("rbs_reset_envs");
}

int rbs_run_body_plane_host(const rbs_body_plane_args *a, void *qpos_host, void *qvel_host, long total_steps) {
    int rc = validate_body_plane(a, false);
    if (rc) return rc;
    if (a->trajectory) return fail(RBS_EINVAL, "rbs_run_body_plane_host: trajectory sampling needs device-resident stepping");
    const long wave_ctas = option("host_wave_ctas") > 0 ? option("host_wave_ctas") : 4;
    return run_host_pipelined(a, 1, 0, wave_ctas * rbs::kBlock, qpos_host, qvel_host, total_steps, launch_body_plane_any);
}

int rbs_run_two_ball_host(const rbs_two_ball_args *a, void *qpos_host, void *qvel_host, long total_steps) {
    int rc = validate_two_ball(a, false);
    if (rc) return rc;
    return run_host_pipelined(a, 2, 0, 8L * rbs::kBlock, qpos_host, qvel_host, total_steps, launch_two_ball_any);
}

int rbs_run_multi_sphere_host(const rbs_multi_sphere_args *a, void *qpos_host, void *qvel_host, long total_steps) {
    int rc = validate_multi_sphere(a, false);
    if (rc) return rc;
    int threads, epb;
    multi_sphere_shape(a->n_body, &threads, &epb);
    // one wave = the CTAs resident at once (about five 128-thread CTAs per SM), in environments
    const long resident = threads <= 256 ? 5 : (threads <= 512 ? 2 : 1);
    return run_host_pipelined(a, a->n_body, 1, resident * epb, qpos_host, qvel_host, total_steps, launch_multi_sphere_any);
}

int rbs_run_multi_body_host(const rbs_multi_body_args *a, void *qpos_host, void *qvel_host, long total_steps) {
    int rc = validate_multi_body(a, false);
    if (rc) return rc;
    // one wave = two resident 256-thread CTAs per SM, in environments
    return run_host_pipelined(a, a->n_body, 1, 2L * (256 / a->n_body), qpos_host, qvel_host, total_steps, launch_multi_body_any);
}

int rbs_release_workspace(void) {
    Pipe *pipe = current_pipe();          // the workspace of the CURRENT device
    if (!pipe) return RBS_OK;
    std::lock_guard<std::mutex> lock(pipe->mutex);
    if (pipe->ws) cudaFree(pipe->ws);
    pipe->ws = nullptr;
    pipe->ws_bytes = 0;
    return RBS_OK;
}

int rbs_stats(int dtype, long n_bodies, const void *state, long stride, const void *mass, double mass_u,
              const void *inertia, long inertia_stride, const double inertia_u[3], const double gravity[3],
              const unsigned *n_contacts, const unsigned *n_impulses, double *out, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_stats: bad dtype %d", dtype);
    if (n_bodies < 0) return fail(RBS_EINVAL, "rbs_stats: n_bodies %ld < 0", n_bodies);
    if (n_bodies == 0) return RBS_OK;
    if (!state || !out || !inertia_u || !gravity) return fail(RBS_EINVAL, "rbs_stats: null argument");
    if (stride < n_bodies) return fail(RBS_EINVAL, "rbs_stats: stride %ld < n_bodies %ld", stride, n_bodies);
    const double gn = sqrt(gravity[0] * gravity[0] + gravity[1] * gravity[1] + gravity[2] * gravity[2]);
    const double ux = gn > 0 ? -gravity[0] / gn : 0.0, uy = gn > 0 ? -gravity[1] / gn : 0.0, uz = gn > 0 ? -gravity[2] / gn : 1.0;
    long blocks = (n_bodies + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == RBS_F64)
        rbs::stats_kernel<double><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
            n_bodies, (const double *)state, stride, (const double *)mass, mass_u, (const double *)inertia, inertia_stride,
            inertia_u[0], inertia_u[1], inertia_u[2], gravity[0], gravity[1], gravity[2], ux, uy, uz, n_contacts, n_impulses, out);
    else
        rbs::stats_kernel<float><<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(
            n_bodies, (const float *)state, stride, (const float *)mass, (float)mass_u, (const float *)inertia, inertia_stride,
            inertia_u[0], inertia_u[1], inertia_u[2], gravity[0], gravity[1], gravity[2], ux, uy, uz, n_contacts, n_impulses, out);
    return check_launch("rbs_stats");
}

int rbs_fma_probe(int dtype, long n_threads, int iters, void *sink, void *stream) {
    if (bad_dtype(dtype)) return fail(RBS_EINVAL, "rbs_fma_probe: bad dtype %d", dtype);
    if (n_threads <= 0 || n_threads % 256 || iters < 1 || !sink) return fail(RBS_EINVAL, "rbs_fma_probe: n_threads must be a positive multiple of 256");
    const unsigned grid = (unsigned)(n_threads / 256);
    const int mode = (int)option("probe_mode");
    cudaStream_t st = as_stream(stream);
    if (dtype == RBS_F64) {
        if (mode == 0) rbs::fma_probe_kernel<double, 0><<<grid, 256, 0, st>>>(iters, (double *)sink, 0.999999, 1e-7);
        else if (mode == 2) rbs::fma_probe_kernel<double, 2><<<grid, 256, 0, st>>>(iters, (double *)sink, 0.999999, 1e-7);
        else rbs::fma_probe_kernel<double, 1><<<grid, 256, 0, st>>>(iters, (double *)sink, 0.999999, 1e-7);
    } else {
        if (mode == 0) rbs::fma_probe_kernel<float, 0><<<grid, 256, 0, st>>>(iters, (float *)sink, 0.999999f, 1e-7f);
        else if (mode == 2) rbs::fma_probe_kernel<float, 2><<<grid, 256, 0, st>>>(iters, (float *)sink, 0.999999f, 1e-7f);
        else rbs::fma_probe_kernel<float, 1><<<grid, 256, 0, st>>>(iters, (float *)sink, 0.999999f, 1e-7f);
    }
    return check_launch("rbs_fma_probe");
}

}  // extern "C"
