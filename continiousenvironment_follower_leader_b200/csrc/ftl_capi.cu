// ftl_capi.cu -- the C-ABI of include/ftl.h (libftl.so), the ray kernel, the bookkeeping / finishing kernels and the
// state-exchange kernels.
//
// Launch structure of one ftl_step() -- a chain of programmatic dependent launches on the caller's stream; each kernel
// waits per group of 32 envs on a flag the kernel it depends on publishes, so they overlap wherever the data allows:
//   k_kin<NB>   (ftl_step_nb.cu) one thread per env: the F sub-frames of robot kinematics fused in registers, per-frame
//               records, tracker scans, history snapshot, observation outputs            -> kin_flag
//   k_book      one thread per env: the bookkeeping of the F recorded frames (exact green-zone flags, trail, timers,
//               early stopping, reward, done), episode statistics.  Needs kin_flag; runs BESIDE k_rays -> book_flag
//   k_rays      one warp per env: history ray casting with static/corridor de-duplication.  Needs kin_flag only
//   k_finish    one thread per env, almost always idle: the (edge, ray) pairs whose float32 predicates were inconclusive,
//               then the in-place auto-reset of finished envs as a copy of the scenario's reset image (state, first
//               observation and its rays).  Needs all of k_rays and book_flag
// All are HBM/ALU streaming kernels without tensor-core work (ray casting is not a contraction).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false (see build.py); fused
// multiply-adds are written explicitly where wanted so float results match the reference's rounding.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "ftl_launch.h"
#include "ftl_rays.cuh"
#include "ftl_step.cuh"
#include "ftl_state_io.cuh"
#include "ftl_reset_image.cuh"
#include "ftl_book.cuh"

using namespace ftl;

// =================================================================================================
// kernels
// =================================================================================================
#ifndef FTL_NO_PDL
#define FTL_NO_PDL 0   // 1: plain stream order between the kernels of a step
#endif
#ifndef FTL_FUSED_BOOK
#define FTL_FUSED_BOOK 1   // 1: k_kin also does the bookkeeping of its env (no k_book launch); 0: separate k_book beside k_rays.
                           // Same-box A/B (profiles/r02_ab_log.txt): fused 0.370 ms per step, split 0.408 -- the bookkeeping is
                           // latency-bound either way and a separate kernel holds registers the ray kernel wants
#endif
#ifndef FTL_RAYS_WARPS
#define FTL_RAYS_WARPS 2   // envs per block.  A block lives as long as its slowest env: 2-warp blocks measured 2.3 % faster
#endif                     // than 4-warp ones (and they fit sooner into what a finished block of the step kernels frees)
#ifndef FTL_RAYS_MINB
#if FTL_RAYS_LANES == 32
#define FTL_RAYS_MINB (28 / FTL_RAYS_WARPS)   // 28 warps per SM: 72 registers, no spills; shared memory (7.2 KB per warp) allows no more
#else
#define FTL_RAYS_MINB (20 / FTL_RAYS_WARPS)   // two envs per warp: 11 KB of shared memory per warp
#endif
#endif
constexpr int kRaysEnvsPerWarp = 32 / FTL_RAYS_LANES;

__global__ void __launch_bounds__(32 * FTL_RAYS_WARPS, FTL_RAYS_MINB)
k_rays(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, const double2* __restrict__ rot,
       float* __restrict__ rays_out, int smem_per_warp, int first_env, int end_env, int wait_seq) {
    asm volatile("griddepcontrol.launch_dependents;");   // k_finish may be scheduled behind the last wave of this grid
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int sub = (int)(threadIdx.x & 31) / FTL_RAYS_LANES;          // which env of the warp this lane serves
    const int i0 = first_env + (blockIdx.x * (blockDim.x >> 5) + warp) * kRaysEnvsPerWarp;   // the warp's envs share a group of 32
    if (i0 >= end_env) return;
    if (wait_seq) wait_group_flag(s.kin_flag, i0 >> 5, wait_seq);   // launched early: wait for these envs' kinematics
    const int i = i0 + sub;
    if (i >= end_env) return;
    RayShared& sh = *reinterpret_cast<RayShared*>(smem + (size_t)(warp * kRaysEnvsPerWarp + sub) * smem_per_warp);
    rays_warp(cfg, s, pool, rot, i, sh, rays_out);
}

#ifndef FTL_BOOK_THREADS
#define FTL_BOOK_THREADS 64
#endif
__global__ void __launch_bounds__(FTL_BOOK_THREADS)
k_book(const __grid_constant__ DevCfg cfg, const __grid_constant__ DevState s, const __grid_constant__ DevOutputs img_out,
       const __grid_constant__ DevOutputs out, double* __restrict__ stats, int n_scenarios, int wait_seq, int seq) {
    asm volatile("griddepcontrol.launch_dependents;");
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // whole warps: n is padded to a multiple of 32
    if (i >= s.n) return;
    if (wait_seq) wait_group_flag(s.kin_flag, i >> 5, wait_seq);
    book_env(cfg, s, img_out, out, stats, n_scenarios, i);
    publish_group_flag(s.book_flag, i >> 5, seq);
}

// One thread per env, almost always idle: the (edge, ray) pairs of k_rays whose float32 predicates were inconclusive are
// redone in float64, and the rays of envs that finished in this step are replaced by those of their next episode's first
// observation (ftl_reset_image.cuh).  Serial behind k_rays, so it does as little as possible.
__global__ void __launch_bounds__(128, 12)   // 40 registers: its blocks are resident (waiting) beside the last ray blocks
k_finish(const __grid_constant__ DevCfg cfg, const __grid_constant__ DevState s, const __grid_constant__ DevPool pool,
         const __grid_constant__ DevOutputs img_out, float* __restrict__ rays_out, int first_env, int end_env, int wait_seq,
         int renew, int pdl) {
    asm volatile("griddepcontrol.launch_dependents;");            // a programmatic dependent (the rollout's policy kernel) may be scheduled behind this grid
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");   // launched early: wait for all of k_rays
    const int i = first_env + blockIdx.x * blockDim.x + threadIdx.x;   // first_env is a multiple of 32: a warp = one group
    const int lane = threadIdx.x & 31;
    if (i - lane >= end_env) return;                              // whole warps leave together
    if (i < end_env) rays_exact_env(cfg, s, pool, i, rays_out);
    if (!renew) return;
    if (wait_seq) wait_group_flag(s.book_flag, (i - lane) >> 5, wait_seq);   // `done` is the bookkeeping kernel's
    const int scen = i < end_env ? pending_reset_scenario(cfg, s, pool.n_scenarios, i) : -1;
    unsigned todo = __ballot_sync(0xffffffffu, scen >= 0);
    const int rpe = cfg.rays_per_env;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int env = __shfl_sync(0xffffffffu, i, src), sc = __shfl_sync(0xffffffffu, scen, src);
        for (int k = lane; k < rpe; k += 32) rays_out[(size_t)env * rpe + k] = img_out.rays[(size_t)sc * rpe + k];
    }
}

// the state half of the auto-reset outside a step (ftl_get_state / ftl_set_state / ftl_reset after a step)
__global__ void __launch_bounds__(128)
k_apply_resets(const __grid_constant__ DevCfg cfg, const __grid_constant__ DevState s, const __grid_constant__ DevPool pool,
               const __grid_constant__ DevState img) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    apply_pending_resets(cfg, s, pool, img, i);
}

// ---- state exchange: SoA <-> FtlEnvState (AoS) ----------------------------------------------------
__global__ void k_pack_state(const DevState s, int first, int count, FtlEnvState* __restrict__ dst) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    pack_env(s, first + j, dst[j]);
}

// FollowerInfo / LeaderTrackDetector_vector outputs from the stored state (launched only when requested)
__global__ void k_optional_sensors(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, const DevOutputs out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < out.n) write_optional_sensors(cfg.c, s, pool, out, i);
}

__global__ void k_unpack_state(const __grid_constant__ DevCfg cfg, const DevState s, int first, int count,
                               const FtlEnvState* __restrict__ src) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    unpack_env(cfg, s, first + j, src[j]);
}

// ---- FP32 FMA microbenchmark (ftl_measure_fp32_peak): 8 independent chains per thread, all in registers ---------
__global__ void __launch_bounds__(256) k_fma_peak(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
#pragma unroll 4
    for (int i = 0; i < iters; i++) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// =================================================================================================
// host side: handle, C-ABI
// =================================================================================================
static thread_local std::string g_err;
void ftl_set_error_message(const char* msg) { g_err = msg; }   // for ftl_scenario_gen.cpp
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t err__ = (expr);                                                                      \
        if (err__ != cudaSuccess)                                                                        \
            return fail(FTL_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(err__));            \
    } while (0)

struct FtlHandle_ {
    DevCfg cfg;
    int n = 0, n_pad = 0, device = 0;
    int n_robots = 2;
    DevState st{};
    DevState image{};   // env s = scenario s right after reset (built by ftl_upload_scenarios); the in-step auto-reset copies from it
    DevOutputs image_out{};   // ... and its first observation (numerical_features, leader_target, rays)
    DevPool pool{};
    bool have_pool = false, was_reset = false;
    bool maybe_pending_resets = false;   // a step ran since the last k_apply_resets: finished envs may await their state copy
    std::vector<void*> allocs, pool_allocs;
    double* d_stats = nullptr;
    // device-side staging for the host-buffer entry points
    void* d_actions = nullptr;
    int* d_in_frames = nullptr;       // FtlStepInputs staging of the host path (allocated on first use)
    double* d_in_draws = nullptr;
    uint8_t* d_mask = nullptr;
    int* d_scen_ids = nullptr;
    DevOutputs d_out{};
    FtlEnvState* d_state_stage = nullptr;
    int state_stage_cap = 0;
    int64_t launches = 0;
    int step_seq = 0;          // sequence number of the last ftl_step (kin_flag / book_flag protocol)
    int step_phase = 0;        // FTL_OPT_STEP_PHASE: 0 = a whole step per call, 1 = k_kin only, 2 = the ray kernels of the step begun before
    int pending_wait_seq = 0;  // ... the sequence number the second half waits for
    bool no_overlap = false;   // FTL_OPT_NO_OVERLAP: plain stream order between the kernels of a step
    bool kin_pdl = false;      // FTL_OPT_KIN_PDL: k_kin launched as a programmatic dependent of the kernel in front of it
    int rays_total = 0;
    bool rays_smem_opted = false;
    double2* d_rot = nullptr;   // (cos, sin)(k * 360/R) per flat ray
    cudaStream_t copy_stream = nullptr;   // host path: D2H copies overlap the ray kernel chunk by chunk
    cudaStream_t own_stream = nullptr;    // ftl_host_stream
    cudaStream_t pending_stream = nullptr;   // stream of an ftl_step_host_begin that has not been waited for
    bool pending = false;
    cudaEvent_t chunk_ev[8] = {};
    // optional per-kernel timing (ftl_profile): kProfEvents events per step on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
};

constexpr int kProfEvents = 4;
static cudaEvent_t prof_event(FtlHandle_* h, cudaStream_t st) {
    if (h->prof_used == h->prof_events.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->prof_events.push_back(e);
    }
    cudaEvent_t e = h->prof_events[h->prof_used++];
    cudaEventRecord(e, st);
    return e;
}

static std::vector<double2> ray_rotation_table(const FtlConfig& c) {
    std::vector<double2> rot;
    for (int s = 0; s < c.n_ray_sensors; s++)
        for (int k = 0; k < c.ray[s].lasers_count; k++) {
            double th = ray_angle(c.ray[s], k) * kDeg2Rad;
            rot.push_back(make_double2(std::cos(th), std::sin(th)));
        }
    return rot;
}

static size_t laser_floats(const FtlConfig& c) {   // floats per env of FtlOutputs.laser
    return c.laser_points > 0 ? (size_t)c.laser_beams * (c.laser_only_distances ? 1 : 2) : 0;
}
static int laser_beam_count(double available_angle, double angle_step) {   // SEN:86-98
    const int border = (int)((available_angle < 360 ? available_angle : 360) / 2);
    double diff = 0;
    int n = 1;
    while (diff < border) { diff += angle_step; n += 2; }
    return n;
}

static float sq_threshold(double limit) {
    // largest float x with sqrtf(x) <= (float)limit: the reference compares float32 square roots
    float lim = (float)limit;
    if (!(lim >= 0.f)) return -1.f;
    float x = lim * lim;
    while (sqrtf(x) > lim) x = nextafterf(x, 0.f);
    for (;;) {
        float y = nextafterf(x, INFINITY);
        if (sqrtf(y) <= lim) x = y; else break;
    }
    return x;
}

template <typename T>
static cudaError_t dalloc(FtlHandle_* h, T** p, size_t count, std::vector<void*>* list = nullptr) {
    void* q = nullptr;
    size_t bytes = (count ? count : 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) return e;
    e = cudaMemset(q, 0, bytes);
    if (e != cudaSuccess) return e;
    (list ? *list : h->allocs).push_back(q);
    *p = (T*)q;
    return cudaSuccess;
}

// the per-env arrays of a DevState whose n / n_real / n_bears are set (the env batch, and the per-scenario reset image)
static cudaError_t alloc_state(FtlHandle_* h, DevState& s, const FtlConfig& c, std::vector<void*>* list) {
    const size_t n = s.n;
    const int nb = s.n_bears, nr = 2 + nb;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(dalloc(h, &s.gd, GD_COUNT * n, list));
    ok(dalloc(h, &s.rd, (size_t)nr * RD_COUNT * n, list));
    ok(dalloc(h, &s.bear_tgt, (size_t)nb * 2 * n, list));
    ok(dalloc(h, &s.gi, GI_COUNT * n, list));
    ok(dalloc(h, &s.ri, (size_t)nr * n, list));
    ok(dalloc(h, &s.bear_idx, (size_t)nb * n, list));
    ok(dalloc(h, &s.gf, GF_COUNT * n, list));
    ok(dalloc(h, &s.pos, (size_t)nr * n, list));
    ok(dalloc(h, &s.rect, (size_t)nr * n, list));
    ok(dalloc(h, &s.trail, n * c.trail_cap, list));
    ok(dalloc(h, &s.trail_d, n * c.trail_cap, list));
    ok(dalloc(h, &s.trail_s, n * c.trail_cap, list));
    ok(dalloc(h, &s.hist, n * c.corridor_cap, list));
    ok(dalloc(h, &s.corridor, n * c.corridor_cap, list));
    ok(dalloc(h, &s.seg_d, n * c.corridor_cap, list));
    ok(dalloc(h, &s.seg_f, n * c.corridor_cap, list));
    ok(dalloc(h, &s.snap_range, (size_t)FTL_MAX_HIST * n, list));
    ok(dalloc(h, &s.snap_rect, (size_t)FTL_MAX_HIST * (1 + nb) * n, list));
    ok(dalloc(h, &s.unc_rec, n * kUncPerEnv, list));
    ok(dalloc(h, &s.unc_count, n, list));
    ok(dalloc(h, &s.rec_f, n * (size_t)c.frames_per_step, list));
    ok(dalloc(h, &s.rec_l, n * (size_t)c.frames_per_step, list));
    ok(dalloc(h, &s.rec_bits, n * (size_t)c.frames_per_step, list));
    ok(dalloc(h, &s.rec_lbits, n * (size_t)c.frames_per_step, list));
    ok(dalloc(h, &s.kin_flag, n / 32 + 1, list));
    ok(dalloc(h, &s.book_flag, n / 32 + 1, list));
    return e;
}

static int validate(const FtlConfig* c, int n_envs) {
    if (!c) return fail(FTL_ERR_INVALID, "config is NULL");
    if (c->abi_version != FTL_ABI_VERSION) return fail(FTL_ERR_INVALID, "FtlConfig.abi_version mismatch");
    if (n_envs <= 0) return fail(FTL_ERR_INVALID, "n_envs must be positive");
    if (c->frames_per_step < 1) return fail(FTL_ERR_INVALID, "frames_per_step must be >= 1");
    if (c->n_bears < 0 || c->n_bears > FTL_MAX_BEARS) return fail(FTL_ERR_INVALID, "n_bears out of range");
    if (c->n_ray_sensors < 0 || c->n_ray_sensors > FTL_MAX_RAY_SENSORS)
        return fail(FTL_ERR_INVALID, "n_ray_sensors out of range");
    if (c->n_ray_sensors > 0 && !c->tracker_enabled)
        return fail(FTL_ERR_INVALID, "ray sensors need LeaderPositionsTracker_v2 (CLS:263-280)");
    if (c->corridor_cap < 2 || (c->corridor_cap & (c->corridor_cap - 1)) || c->corridor_cap > 512)
        return fail(FTL_ERR_INVALID, "corridor_cap must be a power of two in [2, 512]");
    if (c->static_cap < 1 || c->static_cap > 64) return fail(FTL_ERR_INVALID, "static_cap must be in [1, 64]");
    if (c->route_cap < 2) return fail(FTL_ERR_INVALID, "route_cap must be >= 2");
    if (c->trail_cap < 8) return fail(FTL_ERR_INVALID, "trail_cap too small");
    if (c->tracker_enabled && (c->saving_period < 1 || c->tracker_scans_per_step < 1))
        return fail(FTL_ERR_INVALID, "tracker saving_period / scans_per_step must be positive");
    if (c->n_speed_regime > FTL_MAX_REGIME || c->n_accel_regime > FTL_MAX_REGIME)
        return fail(FTL_ERR_INVALID, "too many regime keys");
    for (int s = 0; s < c->n_ray_sensors; s++) {
        const FtlRaySensorConfig& r = c->ray[s];
        if (r.lasers_count < 1) return fail(FTL_ERR_INVALID, "lasers_count must be positive (SEN:761)");
        if (r.max_prev_obs < 1 || r.max_prev_obs > FTL_MAX_HIST)
            return fail(FTL_ERR_INVALID, "max_prev_obs must be in [1, FTL_MAX_HIST] (SEN:876)");
        if (c->fused_sensor_prev && r.max_prev_obs != c->ray[0].max_prev_obs)
            return fail(FTL_ERR_INVALID, "fused_sensor_prev needs the same max_prev_obs on every sensor (WRP:209-210)");
        if (r.n_custom_angles != 0 && (r.n_custom_angles != r.lasers_count || r.n_custom_angles > FTL_MAX_CUSTOM_ANGLES))
            return fail(FTL_ERR_INVALID, "n_custom_angles must be 0 or lasers_count (<= FTL_MAX_CUSTOM_ANGLES)");
        if (r.react_to_obstacles < 0 || r.react_to_obstacles > 3)
            return fail(FTL_ERR_INVALID, "react_to_obstacles must be True/'all'/'static'/'dynamic'/False (SEN:650-661)");
    }
    {
        int total = 0;
        for (int s = 0; s < c->n_ray_sensors; s++) total += c->ray[s].lasers_count;
        if (total >= kMaxTotalRays) return fail(FTL_ERR_INVALID, "more than 4095 rays over all sensors");
    }
    if (c->track_vector_len < 0 || (c->track_vector_len > 0 && !c->tracker_enabled))
        return fail(FTL_ERR_INVALID, "LeaderTrackDetector_vector needs LeaderPositionsTracker_v2 (CLS:240-244)");
    if (c->radar_sectors < 0 || (c->radar_sectors > 0 && (!c->tracker_enabled || c->radar_len < 1 || c->radar_mode < 0 || c->radar_mode > 2)))
        return fail(FTL_ERR_INVALID, "LeaderTrackDetector_radar needs LeaderPositionsTracker_v2, a positive length and mode 0..2 (CLS:240-244, SEN:402-421)");
    if (c->laser_points < 0 || (c->laser_points > 0 && (!(c->laser_angle_step > 0) || c->static_cap > 64 ||
                                                         c->laser_beams != laser_beam_count(c->laser_available_angle, c->laser_angle_step))))
        return fail(FTL_ERR_INVALID, "LaserSensor: points_number, angle_step must be positive and laser_beams = ftl_laser_beam_count(...)");
    return FTL_OK;
}

static DevOutputs to_dev_outputs(const FtlOutputs* o, int n_real) {
    DevOutputs d{};
    d.n = n_real;
    if (o) {
        d.numerical_features = o->numerical_features; d.leader_target = o->leader_target; d.rays = o->rays;
        d.reward = o->reward; d.done = o->done; d.status = o->status;
        d.follower_info = o->follower_info; d.track_vectors = o->track_vectors;
        d.radar = o->radar;
        d.laser = o->laser;
    }
    return d;
}

static void launch_reset(FtlHandle_* h, const DevState& s, const uint8_t* mask, const int* ids, const DevOutputs& o,
                         int reset_filler, cudaStream_t st) {
    switch (h->cfg.c.n_bears) {
        case 0: ftl_launch_reset_nb0(h->cfg, s, h->pool, mask, ids, o, reset_filler, st); break;
        case 1: ftl_launch_reset_nb1(h->cfg, s, h->pool, mask, ids, o, reset_filler, st); break;
        case 2: ftl_launch_reset_nb2(h->cfg, s, h->pool, mask, ids, o, reset_filler, st); break;
        case 3: ftl_launch_reset_nb3(h->cfg, s, h->pool, mask, ids, o, reset_filler, st); break;
        default: ftl_launch_reset_nb4(h->cfg, s, h->pool, mask, ids, o, reset_filler, st); break;
    }
}

static int flush_pending_resets(ftl_handle h, cudaStream_t st);
static int launch_optional_sensors(ftl_handle h, const DevOutputs& o, cudaStream_t st) {
    if (!o.follower_info && !(o.track_vectors && h->cfg.c.track_vector_len > 0) &&
        !(o.radar && h->cfg.c.radar_sectors > 0) && !(o.laser && h->cfg.c.laser_points > 0))
        return FTL_OK;
    // FollowerInfo / LeaderTrackDetector_* read the stored state, which for an env that just finished must already be
    // the first state of its next episode, like the rest of the observation
    int rc = flush_pending_resets(h, st);
    if (rc) return rc;
    k_optional_sensors<<<(h->n + 127) / 128, 128, 0, st>>>(h->cfg, h->st, h->pool, o);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FTL_OK;
}

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                              Args&&... args) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at; lc.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kernel, std::forward<Args>(args)...);
}

// k_rays + k_finish for the envs [first_env, end_env) of state `s` (first_env a multiple of 32).  wait_seq != 0: the
// kernels are programmatic dependents of what is in front of them in the stream and wait on the step's flags instead.
// renew: inside ftl_step -- k_finish also delivers the first rays of the next episode of envs that finished.
static int launch_rays_and_finish(ftl_handle h, const DevState& s, const DevOutputs& o, cudaStream_t st, int first_env,
                                  int end_env, int wait_seq, bool renew) {
    if (end_env <= first_env || !o.rays || h->rays_total == 0) return FTL_OK;
    const bool pdl = wait_seq != 0;
    const int warps = FTL_RAYS_WARPS, threads = warps * 32;
    const int per_warp = (int)((ray_shared_bytes(h->rays_total, h->cfg.ray_hmax) + 15) & ~(size_t)15);
    const int smem = per_warp * warps * kRaysEnvsPerWarp;   // per_warp: one env's block
    if (smem > 48 * 1024 && !h->rays_smem_opted) {
        CUDA_TRY(cudaFuncSetAttribute(k_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        h->rays_smem_opted = true;
    }
    const int blocks = (end_env - first_env + warps * kRaysEnvsPerWarp - 1) / (warps * kRaysEnvsPerWarp);
    const double2* rot = h->d_rot;
    CUDA_TRY(launch_pdl(k_rays, dim3(blocks), dim3(threads), smem, st, pdl, h->cfg, s, h->pool, rot, o.rays, per_warp,
                        first_env, end_env, wait_seq));
    CUDA_TRY(launch_pdl(k_finish, dim3((end_env - first_env + 127) / 128), dim3(128), 0, st, pdl, h->cfg, s, h->pool,
                        h->image_out, o.rays, first_env, end_env, renew ? wait_seq : 0, renew ? 1 : 0, pdl ? 1 : 0));
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
    return FTL_OK;
}

// the state half of pending auto-resets, for callers that look at or replace state between steps
static int flush_pending_resets(ftl_handle h, cudaStream_t st) {
    if (!h->maybe_pending_resets || !h->have_pool) return FTL_OK;
    k_apply_resets<<<(h->st.n + 127) / 128, 128, 0, st>>>(h->cfg, h->st, h->pool, h->image);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->maybe_pending_resets = false;
    return FTL_OK;
}

static int copy_outputs_to_host(ftl_handle h, const FtlOutputs* o, cudaStream_t st) {
    const size_t n = h->n;
    const DevOutputs& d = h->d_out;
    if (o->numerical_features) CUDA_TRY(cudaMemcpyAsync(o->numerical_features, d.numerical_features, 40 * n, cudaMemcpyDeviceToHost, st));
    if (o->leader_target) CUDA_TRY(cudaMemcpyAsync(o->leader_target, d.leader_target, 8 * n, cudaMemcpyDeviceToHost, st));
    if (o->rays && h->cfg.rays_per_env) CUDA_TRY(cudaMemcpyAsync(o->rays, d.rays, sizeof(float) * h->cfg.rays_per_env * n, cudaMemcpyDeviceToHost, st));
    if (o->reward) CUDA_TRY(cudaMemcpyAsync(o->reward, d.reward, 4 * n, cudaMemcpyDeviceToHost, st));
    if (o->done) CUDA_TRY(cudaMemcpyAsync(o->done, d.done, n, cudaMemcpyDeviceToHost, st));
    if (o->status) CUDA_TRY(cudaMemcpyAsync(o->status, d.status, 4 * n, cudaMemcpyDeviceToHost, st));
    if (o->follower_info) CUDA_TRY(cudaMemcpyAsync(o->follower_info, d.follower_info, 8 * n, cudaMemcpyDeviceToHost, st));
    if (o->track_vectors && h->cfg.c.track_vector_len)
        CUDA_TRY(cudaMemcpyAsync(o->track_vectors, d.track_vectors, 8 * n * h->cfg.c.track_vector_len, cudaMemcpyDeviceToHost, st));
    if (o->radar && h->cfg.c.radar_sectors)
        CUDA_TRY(cudaMemcpyAsync(o->radar, d.radar, 4 * n * h->cfg.c.radar_sectors, cudaMemcpyDeviceToHost, st));
    if (o->laser && h->cfg.c.laser_points)
        CUDA_TRY(cudaMemcpyAsync(o->laser, d.laser, laser_floats(h->cfg.c) * 4 * n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return FTL_OK;
}

static FtlOutputs staged_outputs(ftl_handle h, const FtlOutputs* want) {
    // only compute what the caller asked for
    FtlOutputs o{};
    const DevOutputs& d = h->d_out;
    if (want) {
        if (want->numerical_features) o.numerical_features = d.numerical_features;
        if (want->leader_target) o.leader_target = d.leader_target;
        if (want->rays) o.rays = d.rays;
        if (want->reward) o.reward = d.reward;
        if (want->done) o.done = d.done;
        if (want->status) o.status = d.status;
        if (want->follower_info) o.follower_info = d.follower_info;
        if (want->track_vectors && h->cfg.c.track_vector_len) o.track_vectors = d.track_vectors;
        if (want->radar && h->cfg.c.radar_sectors) o.radar = d.radar;
        if (want->laser && h->cfg.c.laser_points) o.laser = d.laser;
    }
    return o;
}

static int ensure_stage(ftl_handle h, int count) {
    if (h->state_stage_cap >= count) return FTL_OK;
    if (h->d_state_stage) cudaFree(h->d_state_stage);
    h->d_state_stage = nullptr;
    h->state_stage_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&h->d_state_stage, sizeof(FtlEnvState) * (size_t)count));
    h->state_stage_cap = count;
    return FTL_OK;
}

extern "C" {

int ftl_abi_version(void) { return FTL_ABI_VERSION; }
const char* ftl_last_error(void) { return g_err.c_str(); }
#define FTL_STR2(x) #x
#define FTL_STR(x) FTL_STR2(x)
const char* ftl_build_info(void) {
    return "edge_cap=" FTL_STR(FTL_EDGE_CAP) " pair_cap=" FTL_STR(FTL_PAIR_CAP) " unc_per_env=" FTL_STR(FTL_UNC_PER_ENV)
           " rays_lanes=" FTL_STR(FTL_RAYS_LANES) " scan_wide=" FTL_STR(FTL_SCAN_WIDE) " walk_wide=" FTL_STR(FTL_WALK_WIDE);
}

int ftl_create(const FtlConfig* cfg, int32_t n_envs, int32_t device, int64_t env_id_base, ftl_handle* out) {
    if (!out) return fail(FTL_ERR_INVALID, "out handle is NULL");
    *out = nullptr;
    int rc = validate(cfg, n_envs);
    if (rc) return rc;
    CUDA_TRY(cudaSetDevice(device));
    FtlHandle_* h = new FtlHandle_();
    h->n = n_envs;
    h->device = device;
    const FtlConfig& c = *cfg;
    DevCfg& d = h->cfg;
    memset(&d, 0, sizeof d);
    d.c = c;
    d.env_id_base = env_id_base;
    d.rays_per_env = 0;
    for (int s = 0; s < c.n_ray_sensors; s++) d.rays_per_env += sensor_width(c.ray[s]);
    ray_out_layout(d);
    d.ray_hmax = ray_hmax(c);
    ray_static_tables(d);
    h->rays_total = total_rays(c);
    d.rays_total = h->rays_total;
    if (h->rays_total > 0) {   // the ray kernel keeps its tables in dynamic shared memory: refuse what cannot fit, here
        int optin = 0;
        CUDA_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
        const size_t need = ((ray_shared_bytes(h->rays_total, d.ray_hmax) + 15) & ~(size_t)15) * FTL_RAYS_WARPS * kRaysEnvsPerWarp;
        if (need > (size_t)optin) {
            const int rt = h->rays_total;
            delete h;
            return fail(FTL_ERR_INVALID, "ray sensors need " + std::to_string(need) + " bytes of shared memory per block (" +
                                             std::to_string(rt) + " rays, " + std::to_string(ray_hmax(c)) +
                                             " history rows); the device allows " + std::to_string(optin) +
                                             ": reduce lasers_count or max_prev_obs");
        }
    }
    d.eps_f32 = (float)c.leader_pos_epsilon;
    d.dev_f32 = (float)c.max_dev;
    d.eps2_f32 = sq_threshold(c.leader_pos_epsilon);
    d.dev2_f32 = sq_threshold(c.max_dev);
    d.min_dist2_f32 = sq_threshold(c.min_distance);
    d.max_distance_f32 = (float)c.max_distance;
    d.es_far_f32 = (float)(c.max_distance * c.es_max_distance_coef);
    d.trail_seed_denom_f32 = (float)(c.trajectory_saving_period * c.leader.max_speed);
    d.corridor_length_f32 = (float)c.corridor_length;
    d.corridor_width_f32 = (float)c.corridor_width;
    auto inflate = [&](const FtlRobotConfig& r) {
        double half_diag = 0.5 * std::sqrt((double)r.width * r.width + (double)r.height * r.height);
        return (float)(half_diag + c.frames_per_step * std::fabs(r.max_speed) + 4.0);
    };
    d.static_inflate[0] = inflate(c.follower);
    d.static_inflate[1] = inflate(c.leader);

    h->n_pad = (n_envs + 31) & ~31;   // whole warps: the step kernel uses warp-wide collectives
    const size_t n = h->n_pad;
    const int nb = c.n_bears, nr = 2 + nb;
    h->n_robots = nr;
    DevState& s = h->st;
    s.n = h->n_pad;
    s.n_real = n_envs;
    s.n_bears = nb;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
    ok(alloc_state(h, s, c, nullptr));
    ok(dalloc(h, &h->d_stats, (size_t)FTL_STAT_COUNT));
    ok(dalloc(h, &h->d_rot, (size_t)h->rays_total));
    if (e == cudaSuccess && h->rays_total > 0) {
        std::vector<double2> rot = ray_rotation_table(c);
        ok(cudaMemcpy(h->d_rot, rot.data(), sizeof(double2) * rot.size(), cudaMemcpyHostToDevice));
    }
    size_t action_bytes = c.action_mode == FTL_ACTION_CONTINUOUS ? 8 : 4;
    ok(dalloc(h, (char**)&h->d_actions, action_bytes * n));
    ok(dalloc(h, &h->d_mask, n));
    ok(dalloc(h, &h->d_scen_ids, n));
    h->d_out.n = n_envs;
    ok(dalloc(h, &h->d_out.numerical_features, 10 * n));
    ok(dalloc(h, &h->d_out.leader_target, 2 * n));
    ok(dalloc(h, &h->d_out.rays, (size_t)(d.rays_per_env ? d.rays_per_env : 1) * n));
    ok(dalloc(h, &h->d_out.reward, n));
    ok(dalloc(h, &h->d_out.done, n));
    ok(dalloc(h, &h->d_out.status, 4 * n));
    ok(dalloc(h, &h->d_out.follower_info, 2 * n));
    ok(dalloc(h, &h->d_out.track_vectors, (size_t)(c.track_vector_len > 0 ? c.track_vector_len : 1) * 2 * n));
    ok(dalloc(h, &h->d_out.radar, (size_t)(c.radar_sectors > 0 ? c.radar_sectors : 1) * n));
    ok(dalloc(h, &h->d_out.laser, (laser_floats(c) ? laser_floats(c) : 1) * n));
    if (e != cudaSuccess) {
        for (void* p : h->allocs) cudaFree(p);
        delete h;
        return fail(e == cudaErrorMemoryAllocation ? FTL_ERR_NOMEM : FTL_ERR_CUDA,
                    std::string("device allocation failed: ") + cudaGetErrorString(e));
    }
    *out = h;
    return FTL_OK;
}

int ftl_destroy(ftl_handle h) {
    if (!h) return FTL_OK;
    cudaSetDevice(h->device);
    for (void* p : h->allocs) cudaFree(p);
    for (void* p : h->pool_allocs) cudaFree(p);
    if (h->d_state_stage) cudaFree(h->d_state_stage);
    for (cudaEvent_t e : h->prof_events) cudaEventDestroy(e);
    for (cudaEvent_t e : h->chunk_ev) if (e) cudaEventDestroy(e);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
    return FTL_OK;
}

int ftl_rays_per_env(ftl_handle h) { return h ? h->cfg.rays_per_env : 0; }
int ftl_laser_beam_count(double available_angle, double angle_step) { return angle_step > 0 ? laser_beam_count(available_angle, angle_step) : 0; }
int ftl_num_envs(ftl_handle h) { return h ? h->n : 0; }
int64_t ftl_launch_count(ftl_handle h) { return h ? h->launches : 0; }

int ftl_upload_scenarios(ftl_handle h, const FtlScenarioPool* p) {
    if (!h || !p) return fail(FTL_ERR_INVALID, "NULL argument");
    const FtlConfig& c = h->cfg.c;
    if (p->n_scenarios <= 0) return fail(FTL_ERR_INVALID, "scenario pool is empty");
    if (p->static_cap != c.static_cap || p->route_cap != c.route_cap)
        return fail(FTL_ERR_INVALID, "pool static_cap/route_cap differ from the configuration");
    const int32_t* ns = p->n_static;
    const int32_t* nr = p->n_route;
    for (int s = 0; s < p->n_scenarios; s++) {
        if (ns[s] < 0 || ns[s] > c.static_cap) return fail(FTL_ERR_INVALID, "n_static out of range");
        if (nr[s] < 2 || nr[s] > c.route_cap) return fail(FTL_ERR_INVALID, "a route needs 2..route_cap waypoints");
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    // the old pool dies here: whatever fails below, no step may run on freed memory, and envs keep scenario ids of
    // the old pool, so a full ftl_reset is required after every upload
    h->have_pool = false;
    h->was_reset = false;
    h->pool = DevPool{};
    h->image = DevState{};
    h->image_out = DevOutputs{};
    for (void* q : h->pool_allocs) cudaFree(q);
    h->pool_allocs.clear();
    const size_t S = p->n_scenarios;
    DevPool& d = h->pool;
    d.n_scenarios = p->n_scenarios;
    int4* sr; int* dns; int2* rt; int* dnr; float2* lp; double* ld; float2* fp; double* fd;
    CUDA_TRY(dalloc(h, &sr, S * c.static_cap, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &dns, S, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &rt, S * c.route_cap, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &dnr, S, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &lp, S, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &ld, S, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &fp, S, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &fd, S, &h->pool_allocs));
    CUDA_TRY(cudaMemcpy(sr, p->static_rects, S * c.static_cap * sizeof(int4), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dns, p->n_static, S * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(rt, p->route, S * c.route_cap * sizeof(int2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(dnr, p->n_route, S * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(lp, p->leader_pos, S * sizeof(float2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(ld, p->leader_dir, S * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(fp, p->follower_pos, S * sizeof(float2), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(fd, p->follower_dir, S * sizeof(double), cudaMemcpyHostToDevice));
    d.static_rects = sr; d.n_static = dns; d.route = rt; d.n_route = dnr;
    d.leader_pos = lp; d.leader_dir = ld; d.follower_pos = fp; d.follower_dir = fd;
    {   // cell grid of nearby static rectangles (near_static_masks_grid): each rectangle marks the cells its grown box touches
        const int gw = (c.game_width >> kNearGridShift) + 1, gh = (c.game_height >> kNearGridShift) + 1;
        const float inf = std::fmax(h->cfg.static_inflate[0], h->cfg.static_inflate[1]) + 1.f;
        const float cell = (float)(1 << kNearGridShift);
        std::vector<uint64_t> grid(S * gw * gh, 0);
        for (size_t sc = 0; sc < S; sc++) {
            uint64_t* g = grid.data() + sc * gw * gh;
            const int32_t* rects = p->static_rects + sc * c.static_cap * 4;   // x, y, w, h
            for (int k = 0; k < ns[sc]; k++) {
                const int32_t* q = rects + 4 * k;
                const float lo_x = (float)q[0] - inf, hi_x = (float)(q[0] + q[2]) + inf;
                const float lo_y = (float)q[1] - inf, hi_y = (float)(q[1] + q[3]) + inf;
                int cx0 = (int)std::floor(lo_x / cell) - 1, cx1 = (int)std::floor(hi_x / cell) + 1;
                int cy0 = (int)std::floor(lo_y / cell) - 1, cy1 = (int)std::floor(hi_y / cell) + 1;
                cx0 = cx0 < 0 ? 0 : cx0; cy0 = cy0 < 0 ? 0 : cy0;
                cx1 = cx1 >= gw ? gw - 1 : cx1; cy1 = cy1 >= gh ? gh - 1 : cy1;
                for (int cy = cy0; cy <= cy1; cy++)
                    for (int cx = cx0; cx <= cx1; cx++) {
                        // closed cell [x0, x0 + cell] x [y0, y0 + cell] against the grown box
                        const float x0 = cx * cell, y0 = cy * cell;
                        if (x0 + cell >= lo_x && x0 <= hi_x && y0 + cell >= lo_y && y0 <= hi_y)
                            g[cy * gw + cx] |= (uint64_t)1 << k;
                    }
            }
        }
        uint64_t* dg;
        CUDA_TRY(dalloc(h, &dg, grid.size(), &h->pool_allocs));
        CUDA_TRY(cudaMemcpy(dg, grid.data(), grid.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
        d.near_grid = dg; d.near_grid_w = gw; d.near_grid_h = gh;
    }
    // the reset image: one k_reset over a DevState with one env per scenario
    DevState& im = h->image;
    im = DevState{};
    im.n = (p->n_scenarios + 31) & ~31;
    im.n_real = p->n_scenarios;
    im.n_bears = c.n_bears;
    {
        cudaError_t e = alloc_state(h, im, c, &h->pool_allocs);
        if (e != cudaSuccess)
            return fail(e == cudaErrorMemoryAllocation ? FTL_ERR_NOMEM : FTL_ERR_CUDA,
                        std::string("reset image allocation failed: ") + cudaGetErrorString(e));
    }
    int* ids;
    CUDA_TRY(dalloc(h, &ids, (size_t)im.n, &h->pool_allocs));
    std::vector<int> iota(im.n);
    for (int k = 0; k < im.n; k++) iota[k] = k < p->n_scenarios ? k : 0;
    CUDA_TRY(cudaMemcpy(ids, iota.data(), sizeof(int) * iota.size(), cudaMemcpyHostToDevice));
    // the first observation of every scenario: k_reset writes numerical_features / leader_target, the ray kernels the rays
    DevOutputs io{};
    io.n = p->n_scenarios;
    CUDA_TRY(dalloc(h, &io.numerical_features, (size_t)im.n * 10, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &io.leader_target, (size_t)im.n * 2, &h->pool_allocs));
    CUDA_TRY(dalloc(h, &io.rays, (size_t)im.n * (h->cfg.rays_per_env ? h->cfg.rays_per_env : 1), &h->pool_allocs));
    launch_reset(h, im, nullptr, ids, io, 1, nullptr);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->image_out = io;
    int rc = launch_rays_and_finish(h, im, io, nullptr, 0, p->n_scenarios, 0, false);
    if (rc) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    h->have_pool = true;
    return FTL_OK;
}

int ftl_reset(ftl_handle h, const uint8_t* mask_dev, const int32_t* scenario_ids_dev, const FtlOutputs* out_dev,
              void* cuda_stream) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    if (!h->have_pool) return fail(FTL_ERR_STATE, "ftl_upload_scenarios must be called before ftl_reset");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevOutputs o = to_dev_outputs(out_dev, h->n);
    if (h->was_reset && mask_dev) {   // envs outside the mask that finished in the last step keep their auto-reset
        int rc0 = flush_pending_resets(h, st);
        if (rc0) return rc0;
    }
    h->maybe_pending_resets = false;
    const int reset_filler = (!mask_dev || !h->was_reset) ? 1 : 0;
    launch_reset(h, h->st, mask_dev, scenario_ids_dev, o, reset_filler, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->was_reset = true;
    int rc = launch_rays_and_finish(h, h->st, o, st, 0, h->n, 0, false);
    if (rc) return rc;
    return launch_optional_sensors(h, o, st);
}

static void launch_kin(FtlHandle_* h, const void* actions, const DevOutputs& o, int seq, int pdl, cudaStream_t st) {
    switch (h->cfg.c.n_bears) {
        case 0: ftl_launch_kin_nb0(h->cfg, h->st, h->pool, h->image, h->image_out, actions, o, h->d_stats, seq, FTL_FUSED_BOOK, pdl, st); break;
        case 1: ftl_launch_kin_nb1(h->cfg, h->st, h->pool, h->image, h->image_out, actions, o, h->d_stats, seq, FTL_FUSED_BOOK, pdl, st); break;
        case 2: ftl_launch_kin_nb2(h->cfg, h->st, h->pool, h->image, h->image_out, actions, o, h->d_stats, seq, FTL_FUSED_BOOK, pdl, st); break;
        case 3: ftl_launch_kin_nb3(h->cfg, h->st, h->pool, h->image, h->image_out, actions, o, h->d_stats, seq, FTL_FUSED_BOOK, pdl, st); break;
        default: ftl_launch_kin_nb4(h->cfg, h->st, h->pool, h->image, h->image_out, actions, o, h->d_stats, seq, FTL_FUSED_BOOK, pdl, st); break;
    }
    h->launches++;
}

// k_kin and k_book of one step.  *wait_seq = the sequence number the dependent kernels wait for, or 0 when the kernels
// run in plain stream order (per-kernel timing, a stream that is being captured, FTL_NO_PDL).
static int launch_step_front(ftl_handle h, const void* actions_dev, const DevOutputs& o, cudaStream_t st, int* wait_seq) {
    // While the stream is being captured the sequence number would be frozen into the graph: a replay would find the
    // previous replay's flags already equal to it and skip the wait.  Captured steps therefore use plain stream order.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &cap));
    const bool overlap = !h->profiling && !FTL_NO_PDL && !h->no_overlap && cap == cudaStreamCaptureStatusNone;
    const int seq = h->step_seq = (h->step_seq % 0x3fffffff) + 1;   // never 0
    if (h->profiling) prof_event(h, st);
    launch_kin(h, actions_dev, o, seq, (overlap && h->kin_pdl) ? 1 : 0, st);
    CUDA_TRY(cudaGetLastError());
    if (h->profiling) prof_event(h, st);
    if (!FTL_FUSED_BOOK) {
        CUDA_TRY(launch_pdl(k_book, dim3((h->st.n + FTL_BOOK_THREADS - 1) / FTL_BOOK_THREADS), dim3(FTL_BOOK_THREADS), 0, st,
                            overlap, h->cfg, h->st, h->image_out, o, h->d_stats, h->pool.n_scenarios, overlap ? seq : 0, seq));
        h->launches++;
    }
    if (h->profiling) prof_event(h, st);
    *wait_seq = overlap ? seq : 0;
    h->maybe_pending_resets = true;
    return FTL_OK;
}

// the per-step inputs travel inside the DevState argument of the launches of this step only
struct StepInputsScope {
    ftl_handle h;
    StepInputsScope(ftl_handle h_, const FtlStepInputs* in) : h(h_) {
        h->st.in_frames = in ? in->frames_per_step : nullptr;
        h->st.in_draws = in ? in->regime_draws : nullptr;
    }
    ~StepInputsScope() { h->st.in_frames = nullptr; h->st.in_draws = nullptr; }
};

int ftl_step_ex(ftl_handle h, const void* actions_dev, const FtlStepInputs* in_dev, const FtlOutputs* out_dev,
                void* cuda_stream) {
    if (!h || !actions_dev) return fail(FTL_ERR_INVALID, "NULL argument");
    if (!h->was_reset) return fail(FTL_ERR_STATE, "ftl_reset must be called before ftl_step");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevOutputs o = to_dev_outputs(out_dev, h->n);
    StepInputsScope scope(h, in_dev);
    int wait_seq = 0;
    int rc = FTL_OK;
    if (h->step_phase != 2) {
        rc = launch_step_front(h, actions_dev, o, st, &wait_seq);
        if (rc) return rc;
        h->pending_wait_seq = wait_seq;
        if (h->step_phase == 1) return FTL_OK;
    } else {
        wait_seq = h->pending_wait_seq;
    }
    rc = launch_rays_and_finish(h, h->st, o, st, 0, h->n, wait_seq, true);
    if (h->profiling) prof_event(h, st);
    if (rc) return rc;
    return launch_optional_sensors(h, o, st);
}

int ftl_step(ftl_handle h, const void* actions_dev, const FtlOutputs* out_dev, void* cuda_stream) {
    return ftl_step_ex(h, actions_dev, nullptr, out_dev, cuda_stream);
}

void* ftl_host_stream(ftl_handle h) {
    if (!h) return nullptr;
    if (!h->own_stream) {
        cudaSetDevice(h->device);
        if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) h->own_stream = nullptr;
    }
    return (void*)h->own_stream;
}

int ftl_step_host_wait(ftl_handle h) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    if (!h->pending) return FTL_OK;
    h->pending = false;
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->copy_stream) CUDA_TRY(cudaStreamSynchronize(h->copy_stream));
    CUDA_TRY(cudaStreamSynchronize(h->pending_stream));
    return FTL_OK;
}

static int step_host_begin(ftl_handle h, const void* actions_host, const FtlStepInputs* in_host, const FtlOutputs* out_host,
                           void* cuda_stream);

int ftl_step_host(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream) {
    int rc = step_host_begin(h, actions_host, nullptr, out_host, cuda_stream);
    if (rc) return rc;
    return ftl_step_host_wait(h);
}

int ftl_step_host_ex(ftl_handle h, const void* actions_host, const FtlStepInputs* in_host, const FtlOutputs* out_host,
                     void* cuda_stream) {
    int rc = step_host_begin(h, actions_host, in_host, out_host, cuda_stream);
    if (rc) return rc;
    return ftl_step_host_wait(h);
}

int ftl_step_host_begin(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream) {
    return step_host_begin(h, actions_host, nullptr, out_host, cuda_stream);
}

static int step_host_begin(ftl_handle h, const void* actions_host, const FtlStepInputs* in_host, const FtlOutputs* out_host,
                           void* cuda_stream) {
    if (!h || !actions_host) return fail(FTL_ERR_INVALID, "NULL argument");
    if (!h->was_reset) return fail(FTL_ERR_STATE, "ftl_reset must be called before ftl_step");
    if (h->pending) return fail(FTL_ERR_STATE, "ftl_step_host_begin: the previous step has not been waited for");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t n = h->n;
    if (!h->copy_stream) {
        CUDA_TRY(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        for (auto& e : h->chunk_ev) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t cs = h->copy_stream;
    size_t action_bytes = (h->cfg.c.action_mode == FTL_ACTION_CONTINUOUS ? 8 : 4) * n;
    CUDA_TRY(cudaMemcpyAsync(h->d_actions, actions_host, action_bytes, cudaMemcpyHostToDevice, st));
    FtlStepInputs in_dev{};
    if (in_host && in_host->frames_per_step) {
        if (!h->d_in_frames) CUDA_TRY(dalloc(h, &h->d_in_frames, (size_t)h->n_pad));
        CUDA_TRY(cudaMemcpyAsync(h->d_in_frames, in_host->frames_per_step, sizeof(int) * n, cudaMemcpyHostToDevice, st));
        in_dev.frames_per_step = h->d_in_frames;
    }
    if (in_host && in_host->regime_draws) {
        const size_t cnt = (size_t)h->cfg.c.frames_per_step;
        if (!h->d_in_draws) CUDA_TRY(dalloc(h, &h->d_in_draws, (size_t)h->n_pad * cnt));
        CUDA_TRY(cudaMemcpyAsync(h->d_in_draws, in_host->regime_draws, sizeof(double) * n * cnt, cudaMemcpyHostToDevice, st));
        in_dev.regime_draws = h->d_in_draws;
    }
    StepInputsScope scope(h, &in_dev);
    FtlOutputs so = staged_outputs(h, out_host);
    if (!out_host) so.rays = h->d_out.rays;   // nobody reads them, but the step stays the same work
    h->pending_stream = st;
    const bool want_rays = out_host && out_host->rays && h->cfg.rays_per_env;
    if (!want_rays) {   // the whole step in one go, then the copies
        int rc = ftl_step_ex(h, h->d_actions, &in_dev, &so, cuda_stream);
        if (rc) return rc;
    } else {
        // kinematics + bookkeeping for everybody, then the ray and finishing kernels in chunks of envs: the D2H copy of
        // chunk c runs on the copy stream while chunk c+1 is being cast
        DevOutputs o = to_dev_outputs(&so, h->n);
        int wait_seq = 0;
        int rc = launch_step_front(h, h->d_actions, o, st, &wait_seq);
        if (rc) return rc;
        const int chunks = h->n >= 8192 ? 6 : 1;
        const size_t row = sizeof(float) * (size_t)h->cfg.rays_per_env;
        for (int c = 0; c < chunks; c++) {
            int first = (int)(((long long)h->n * c / chunks) & ~31LL), end = (int)(((long long)h->n * (c + 1) / chunks) & ~31LL);
            if (c == chunks - 1) end = h->n;
            rc = launch_rays_and_finish(h, h->st, o, st, first, end, wait_seq, true);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(h->chunk_ev[1 + c], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, h->chunk_ev[1 + c], 0));
            CUDA_TRY(cudaMemcpyAsync((char*)out_host->rays + row * first, (const char*)h->d_out.rays + row * first,
                                     row * (size_t)(end - first), cudaMemcpyDeviceToHost, cs));
        }
        if (h->profiling) prof_event(h, st);
        rc = launch_optional_sensors(h, o, st);
        if (rc) return rc;
    }
    if (out_host) {   // the small outputs: after the last finishing kernel (it rewrites the observation of reset envs)
        const DevOutputs& d = h->d_out;
        CUDA_TRY(cudaEventRecord(h->chunk_ev[0], st));
        CUDA_TRY(cudaStreamWaitEvent(cs, h->chunk_ev[0], 0));
        if (out_host->numerical_features) CUDA_TRY(cudaMemcpyAsync(out_host->numerical_features, d.numerical_features, 40 * n, cudaMemcpyDeviceToHost, cs));
        if (out_host->leader_target) CUDA_TRY(cudaMemcpyAsync(out_host->leader_target, d.leader_target, 8 * n, cudaMemcpyDeviceToHost, cs));
        if (out_host->reward) CUDA_TRY(cudaMemcpyAsync(out_host->reward, d.reward, 4 * n, cudaMemcpyDeviceToHost, cs));
        if (out_host->done) CUDA_TRY(cudaMemcpyAsync(out_host->done, d.done, n, cudaMemcpyDeviceToHost, cs));
        if (out_host->status) CUDA_TRY(cudaMemcpyAsync(out_host->status, d.status, 4 * n, cudaMemcpyDeviceToHost, cs));
        if (out_host->follower_info) CUDA_TRY(cudaMemcpyAsync(out_host->follower_info, d.follower_info, 8 * n, cudaMemcpyDeviceToHost, cs));
        if (out_host->track_vectors && h->cfg.c.track_vector_len)
            CUDA_TRY(cudaMemcpyAsync(out_host->track_vectors, d.track_vectors, 8 * n * h->cfg.c.track_vector_len, cudaMemcpyDeviceToHost, cs));
        if (out_host->radar && h->cfg.c.radar_sectors)
            CUDA_TRY(cudaMemcpyAsync(out_host->radar, d.radar, 4 * n * h->cfg.c.radar_sectors, cudaMemcpyDeviceToHost, cs));
        if (out_host->laser && h->cfg.c.laser_points)
            CUDA_TRY(cudaMemcpyAsync(out_host->laser, d.laser, laser_floats(h->cfg.c) * 4 * n, cudaMemcpyDeviceToHost, cs));
    }
    h->pending = true;
    return FTL_OK;
}

int ftl_reset_host(ftl_handle h, const uint8_t* mask_host, const int32_t* scenario_ids_host, const FtlOutputs* out_host,
                   void* cuda_stream) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (mask_host) CUDA_TRY(cudaMemcpyAsync(h->d_mask, mask_host, h->n, cudaMemcpyHostToDevice, st));
    if (scenario_ids_host) CUDA_TRY(cudaMemcpyAsync(h->d_scen_ids, scenario_ids_host, 4 * (size_t)h->n, cudaMemcpyHostToDevice, st));
    FtlOutputs o = staged_outputs(h, out_host);
    int rc = ftl_reset(h, mask_host ? h->d_mask : nullptr, scenario_ids_host ? h->d_scen_ids : nullptr, &o, cuda_stream);
    if (rc) return rc;
    if (!out_host) { CUDA_TRY(cudaStreamSynchronize(st)); return FTL_OK; }
    return copy_outputs_to_host(h, out_host, st);
}

int ftl_get_state(ftl_handle h, int32_t first, int32_t count, const FtlStateBuffers* b) {
    if (!h || !b) return fail(FTL_ERR_INVALID, "NULL argument");
    if (first < 0 || count < 0 || first + count > h->n) return fail(FTL_ERR_INVALID, "env range out of bounds");
    if (count == 0) return FTL_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    {
        int rc0 = flush_pending_resets(h, nullptr);
        if (rc0) return rc0;
    }
    const FtlConfig& c = h->cfg.c;
    if (b->env) {
        int rc = ensure_stage(h, count);
        if (rc) return rc;
        k_pack_state<<<(count + 127) / 128, 128>>>(h->st, first, count, h->d_state_stage);
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaMemcpy(b->env, h->d_state_stage, sizeof(FtlEnvState) * (size_t)count, cudaMemcpyDeviceToHost));
    }
    if (b->trail) CUDA_TRY(cudaMemcpy(b->trail, h->st.trail + (size_t)first * c.trail_cap, sizeof(float2) * (size_t)c.trail_cap * count, cudaMemcpyDeviceToHost));
    if (b->hist) CUDA_TRY(cudaMemcpy(b->hist, h->st.hist + (size_t)first * c.corridor_cap, sizeof(double2) * (size_t)c.corridor_cap * count, cudaMemcpyDeviceToHost));
    if (b->corridor) CUDA_TRY(cudaMemcpy(b->corridor, h->st.corridor + (size_t)first * c.corridor_cap, sizeof(float4) * (size_t)c.corridor_cap * count, cudaMemcpyDeviceToHost));
    return FTL_OK;
}

int ftl_set_state(ftl_handle h, int32_t first, int32_t count, const FtlStateBuffers* b) {
    if (!h || !b) return fail(FTL_ERR_INVALID, "NULL argument");
    if (first < 0 || count < 0 || first + count > h->n) return fail(FTL_ERR_INVALID, "env range out of bounds");
    if (!b->env || !b->trail) return fail(FTL_ERR_INVALID, "set_state needs at least env and trail");
    if (!h->have_pool) return fail(FTL_ERR_STATE, "ftl_upload_scenarios must be called before ftl_set_state");
    if (count == 0) return FTL_OK;
    const FtlConfig& c = h->cfg.c;
    // every index the kernels will follow is checked here, on the host, before anything reaches the device
    for (int k = 0; k < count; k++) {
        const FtlEnvState& e = b->env[k];
        const char* bad = nullptr;
        const int64_t live = (int64_t)e.ring_head - (int64_t)e.ring_tail;
        if (e.trail_len < 0 || e.trail_len > c.trail_cap) bad = "trail_len outside [0, trail_cap]";
        else if (e.ring_tail < 0 || live < 0 || live > c.corridor_cap) bad = "ring_head - ring_tail outside [0, corridor_cap]";
        else if (e.scenario_id < 0 || e.scenario_id >= h->pool.n_scenarios) bad = "scenario_id outside the uploaded pool";
        else if (e.snap_pushes < 0) bad = "snap_pushes is negative";
        else if (e.cur_target_id < 0) bad = "cur_target_id is negative";
        else if (e.mission_status < 0 || e.mission_status > 3 || e.agent_status < 0 || e.agent_status > 7 ||
                 e.leader_status < 0 || e.leader_status > 3) bad = "status code out of range";
        for (int q = 0; !bad && q < c.n_bears; q++)
            if (e.bear_index[q] < 0 || e.bear_index[q] > 3) bad = "bear_index outside 0..3";
        for (int q = 0; !bad && q < FTL_MAX_HIST; q++) {
            const FtlSnapshot& sn = e.snap[q];
            if (!sn.valid) continue;
            const int64_t span = (int64_t)sn.corr_head - (int64_t)sn.corr_tail;
            if (sn.corr_tail < 0 || span < 0 || span > c.corridor_cap) bad = "history snapshot range outside the corridor ring";
        }
        if (bad) return fail(FTL_ERR_INVALID, "ftl_set_state: env " + std::to_string(first + k) + ": " + bad);
    }
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    {
        int rc0 = flush_pending_resets(h, nullptr);
        if (rc0) return rc0;
        CUDA_TRY(cudaDeviceSynchronize());
    }
    CUDA_TRY(cudaMemcpy(h->st.trail + (size_t)first * c.trail_cap, b->trail, sizeof(float2) * (size_t)c.trail_cap * count, cudaMemcpyHostToDevice));
    if (b->hist) CUDA_TRY(cudaMemcpy(h->st.hist + (size_t)first * c.corridor_cap, b->hist, sizeof(double2) * (size_t)c.corridor_cap * count, cudaMemcpyHostToDevice));
    if (b->corridor) CUDA_TRY(cudaMemcpy(h->st.corridor + (size_t)first * c.corridor_cap, b->corridor, sizeof(float4) * (size_t)c.corridor_cap * count, cudaMemcpyHostToDevice));
    int rc = ensure_stage(h, count);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpy(h->d_state_stage, b->env, sizeof(FtlEnvState) * (size_t)count, cudaMemcpyHostToDevice));
    k_unpack_state<<<(count + 127) / 128, 128>>>(h->cfg, h->st, first, count, h->d_state_stage);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaDeviceSynchronize());
    h->was_reset = true;
    return FTL_OK;
}

// ---- rgb_array rasteriser (Game.render / _show_tick, ENV:1196-1302) -----------------------
// One thread per output pixel, one block per 16 x 16 tile of one env's image; the primitives are read straight from the
// SoA state (broadcast loads: every thread of a block walks the same lists).  Layer order as _show_tick: background, the
// leader's route + finish point, the green zone, the min-distance ring, the objects (hit boxes instead of sprites: the
// image files are not part of the simulator), the tracker's history and corridor (LeaderPositionsTracker_v2.show,
// SEN:329-339), the current target ring.  Text is not drawn.
namespace {
struct Rgb { unsigned char r, g, b; };
__device__ __forceinline__ float seg_dist2(float px, float py, float ax, float ay, float bx, float by) {
    const float vx = bx - ax, vy = by - ay, wx = px - ax, wy = py - ay;
    const float vv = vx * vx + vy * vy;
    float t = vv > 0.f ? (wx * vx + wy * vy) / vv : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    const float dx = wx - t * vx, dy = wy - t * vy;
    return dx * dx + dy * dy;
}
__device__ __forceinline__ bool in_rect(float px, float py, int4 q) {   // pygame.Rect: [x, x + w) x [y, y + h)
    return px >= (float)q.x && px < (float)(q.x + q.z) && py >= (float)q.y && py < (float)(q.y + q.w);
}

__global__ void __launch_bounds__(256)
k_render(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, int first, int scale, int W, int H,
         unsigned char* __restrict__ rgb) {
    const FtlConfig& c = cfg.c;
    const int env = first + blockIdx.z;
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    if (x >= W || y >= H) return;
    const float px = ((float)x + 0.5f) * (float)scale, py = ((float)y + 0.5f) * (float)scale;   // world coordinates
    const float thin = 0.5f * (float)scale;   // half width of a one-pixel line, in world units
    const size_t n = s.n;
    const int scen = s.gi[(size_t)GI_SCENARIO * n + env], trail_len = s.gi[(size_t)GI_TRAIL_LEN * n + env];
    const int flags = s.gi[(size_t)GI_FLAGS * n + env];
    Rgb col = {255, 255, 255};
    const Rgb red = {255, 0, 0}, green = {0, 255, 0}, gray = {30, 30, 30}, blue = {0, 0, 255};
    // (1) the leader's route (show_leader_path_flag) and its finish point, ENV:1233-1238
    const int n_route = pool.n_route[scen];
    const int2* route = pool.route + (size_t)scen * c.route_cap;
    if (n_route > 2) {
        for (int k = 0; k + 1 < n_route; k++) {
            const int2 a = route[k], b = route[k + 1];
            if (seg_dist2(px, py, (float)a.x, (float)a.y, (float)b.x, (float)b.y) <= thin * thin) { col = red; break; }
        }
    }
    if (n_route > 0) {
        const int2 f = route[n_route - 1];
        const float dx = px - (float)f.x, dy = py - (float)f.y, r = fmaxf(5.f, thin);
        if (dx * dx + dy * dy <= r * r) col = red;
    }
    // (2) the green zone: a disc of radius max_dev around every green trail point, ENV:1246-1250 (the points whose
    //     membership the step kernel left undecided -- at most a handful at the window's far end -- are drawn too)
    {
        const float2* trail = s.trail + (size_t)env * c.trail_cap;
        const int g_lo = s.gi[(size_t)GI_G_LO * n + env] - s.gi[(size_t)GI_G_UNC * n + env], hi = trail_len - 2;
        if (hi - g_lo + 1 > 5) {
            const float r2 = cfg.dev_f32 * cfg.dev_f32;
            for (int k = g_lo < 0 ? 0 : g_lo; k <= hi; k++) {
                const float2 p = trail[k];
                const float dx = px - p.x, dy = py - p.y;
                if (dx * dx + dy * dy <= r2) { col = green; break; }
            }
        }
    }
    // (3) the min-distance ring around the leader, two pixels wide while the follower is too close, ENV:1251-1259
    const float2 lp = s.pos[(size_t)1 * n + env];
    {
        const float d = sqrtf((px - lp.x) * (px - lp.x) + (py - lp.y) * (py - lp.y));
        const float half = ((flags & FL_TOO_CLOSE) ? 2.f : 1.f) * thin;
        if (fabsf(d - (float)c.min_distance) <= half) col = red;
    }
    // (4) objects, ENV:1266-1272: static obstacles, then leader, follower, bears (their integer hit boxes)
    {
        const int4* statics = pool.static_rects + (size_t)scen * c.static_cap;
        const int n_static = pool.n_static[scen];
        for (int k = 0; k < n_static; k++)
            if (in_rect(px, py, statics[k])) { col = gray; break; }
        if (in_rect(px, py, s.rect[(size_t)1 * n + env])) col = blue;
        if (in_rect(px, py, s.rect[(size_t)0 * n + env])) col = Rgb{255, 140, 0};
        for (int b = 0; b < s.n_bears; b++)
            if (in_rect(px, py, s.rect[(size_t)(2 + b) * n + env])) col = Rgb{139, 69, 19};
    }
    // (5) the tracker: history points and the corridor's two walls + end caps, SEN:329-339
    if (c.tracker_enabled) {
        const int tail = s.gi[(size_t)GI_RING_TAIL * n + env], head = s.gi[(size_t)GI_RING_HEAD * n + env];
        const int mask = c.corridor_cap - 1;
        const double2* hist = s.hist + (size_t)env * c.corridor_cap;
        const float4* corr = s.corridor + (size_t)env * c.corridor_cap;
        const float rp = fmaxf(3.f, thin), hw = fmaxf(1.5f, thin);
        bool wall = false, point = false;
        for (int k = tail; k < head && !point; k++) {
            const double2 h = hist[k & mask];
            const float dx = px - (float)h.x, dy = py - (float)h.y;
            point = dx * dx + dy * dy <= rp * rp;
        }
        if (head - tail > 1) {
            for (int k = tail; k + 1 < head && !wall; k++) {
                const float4 a = corr[k & mask], b = corr[(k + 1) & mask];
                wall = seg_dist2(px, py, a.x, a.y, b.x, b.y) <= hw * hw || seg_dist2(px, py, a.z, a.w, b.z, b.w) <= hw * hw;
            }
            const float4 a = corr[tail & mask], b = corr[(head - 1) & mask];
            wall = wall || seg_dist2(px, py, a.x, a.y, a.z, a.w) <= hw * hw || seg_dist2(px, py, b.x, b.y, b.z, b.w) <= hw * hw;
        }
        if (point) col = Rgb{80, 10, 10};
        if (wall) col = Rgb{150, 120, 50};
    }
    // (6) the leader's current target: a ring of radius 10, two pixels wide, ENV:1279
    {
        int tid = s.gi[(size_t)GI_TARGET_ID * n + env];
        tid = tid < n_route ? tid : n_route - 1;
        if (tid >= 0) {
            const int2 t = route[tid];
            const float d = sqrtf((px - (float)t.x) * (px - (float)t.x) + (py - (float)t.y) * (py - (float)t.y));
            if (fabsf(d - 10.f) <= fmaxf(1.f, thin)) col = red;
        }
    }
    unsigned char* o = rgb + (((size_t)blockIdx.z * H + y) * W + x) * 3;
    o[0] = col.r; o[1] = col.g; o[2] = col.b;
}
}  // namespace

int ftl_set_option(ftl_handle h, int32_t option, int32_t value) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    switch (option) {
        case FTL_OPT_KIN_PDL: h->kin_pdl = value != 0; return FTL_OK;
        case FTL_OPT_STEP_PHASE: if (value < 0 || value > 2) return fail(FTL_ERR_INVALID, "FTL_OPT_STEP_PHASE: 0, 1 or 2"); h->step_phase = value; return FTL_OK;
        case FTL_OPT_NO_OVERLAP: h->no_overlap = value != 0; return FTL_OK;
        default: return fail(FTL_ERR_INVALID, "unknown option");
    }
}

int ftl_render(ftl_handle h, int32_t first, int32_t count, int32_t scale, uint8_t* rgb_dev, void* cuda_stream) {
    if (!h || !rgb_dev) return fail(FTL_ERR_INVALID, "NULL argument");
    if (first < 0 || count < 0 || first + count > h->n) return fail(FTL_ERR_INVALID, "env range out of bounds");
    if (scale < 1) return fail(FTL_ERR_INVALID, "scale must be >= 1");
    if (!h->have_pool || !h->was_reset) return fail(FTL_ERR_STATE, "ftl_render needs an uploaded scenario pool and a reset");
    if (count == 0) return FTL_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    int rc = flush_pending_resets(h, st);   // a finished env is drawn as the first frame of its next episode
    if (rc) return rc;
    const int W = (h->cfg.c.game_width + scale - 1) / scale, H = (h->cfg.c.game_height + scale - 1) / scale;
    k_render<<<dim3((W + 15) / 16, (H + 15) / 16, count), 256, 0, st>>>(h->cfg, h->st, h->pool, first, scale, W, H, rgb_dev);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FTL_OK;
}

int ftl_render_host(ftl_handle h, int32_t first, int32_t count, int32_t scale, uint8_t* rgb_host) {
    if (!h || !rgb_host) return fail(FTL_ERR_INVALID, "NULL argument");
    if (scale < 1) return fail(FTL_ERR_INVALID, "scale must be >= 1");
    if (count <= 0) return count == 0 ? FTL_OK : fail(FTL_ERR_INVALID, "env range out of bounds");
    CUDA_TRY(cudaSetDevice(h->device));
    const size_t W = (h->cfg.c.game_width + scale - 1) / scale, H = (h->cfg.c.game_height + scale - 1) / scale;
    const size_t bytes = W * H * 3 * (size_t)count;
    uint8_t* d = nullptr;
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMalloc((void**)&d, bytes));
    int rc = ftl_render(h, first, count, scale, d, nullptr);
    cudaError_t e = rc == FTL_OK ? cudaMemcpy(rgb_host, d, bytes, cudaMemcpyDeviceToHost) : cudaSuccess;
    cudaFree(d);
    if (rc) return rc;
    CUDA_TRY(e);
    return FTL_OK;
}

int ftl_stats(ftl_handle h, double* stats_dev, int32_t reset_after, void* cuda_stream) {
    if (!h || !stats_dev) return fail(FTL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CUDA_TRY(cudaMemcpyAsync(stats_dev, h->d_stats, sizeof(double) * FTL_STAT_COUNT, cudaMemcpyDeviceToDevice, st));
    if (reset_after) CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, sizeof(double) * FTL_STAT_COUNT, st));
    return FTL_OK;
}

int ftl_measure_fp32_peak(int32_t device, double* tflops_out) {
    if (!tflops_out) return fail(FTL_ERR_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(device));
    int sms = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int blocks = sms * 8, threads = 256, iters = 1 << 15;
    float* buf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&buf, sizeof(float) * (size_t)blocks * threads));
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {   // the first launches warm the clocks up; the best of the rest is the peak
        cudaEventRecord(e0);
        k_fma_peak<<<blocks, threads>>>(buf, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaError_t err = cudaEventSynchronize(e1);
        if (err != cudaSuccess) { cudaFree(buf); return fail(FTL_ERR_CUDA, cudaGetErrorString(err)); }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
        if (rep >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(buf);
    *tflops_out = best;
    return FTL_OK;
}

int ftl_profile(ftl_handle h, int32_t enable) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    h->profiling = enable != 0;
    h->prof_used = 0;
    return FTL_OK;
}

static int profile_collect(ftl_handle h, double ms_out[3], int64_t* steps) {
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    ms_out[0] = ms_out[1] = ms_out[2] = 0.0;
    const size_t n = h->prof_used / kProfEvents;   // per step: before k_kin, after it, after k_book, after k_rays + k_finish
    for (size_t k = 0; k < n; k++)
        for (int j = 0; j < 3; j++) {
            float ms = 0;
            CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[kProfEvents * k + j], h->prof_events[kProfEvents * k + j + 1]));
            ms_out[j] += ms;
        }
    if (steps) *steps = (int64_t)n;
    h->prof_used = 0;
    return FTL_OK;
}

int ftl_profile_read(ftl_handle h, double* step_kernel_ms, double* ray_kernel_ms, int64_t* steps) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    double ms[3];
    int rc = profile_collect(h, ms, steps);
    if (rc) return rc;
    if (step_kernel_ms) *step_kernel_ms = ms[0] + ms[1];
    if (ray_kernel_ms) *ray_kernel_ms = ms[2];
    return FTL_OK;
}

int ftl_profile_read_kernels(ftl_handle h, double* kin_ms, double* book_ms, double* rays_ms, int64_t* steps) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    double ms[3];
    int rc = profile_collect(h, ms, steps);
    if (rc) return rc;
    if (kin_ms) *kin_ms = ms[0];
    if (book_ms) *book_ms = ms[1];
    if (rays_ms) *rays_ms = ms[2];
    return FTL_OK;
}

}  // extern "C"
