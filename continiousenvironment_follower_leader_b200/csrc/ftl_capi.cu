// ftl_capi.cu -- the C-ABI of include/ftl.h (libftl.so), the ray kernel and the state-exchange kernels.
//
// Launch structure of one ftl_step():
//   k_step<NB>    (ftl_step_nb.cu) one thread per env: the F sub-frames fused in registers (a kinematics pass and a
//                 bookkeeping pass), tracker scans, history snapshot, non-ray outputs, episode statistics, optional
//                 in-place auto-reset (a copy of the scenario's reset image)
//   k_rays        one warp per env: history ray casting with static/corridor de-duplication; launched as a programmatic
//                 dependent of k_step, it waits per group of 32 envs on the flag the owning warp of k_step publishes
//   k_rays_exact  one thread per env, almost always idle: the pairs whose float32 predicates were inconclusive
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

using namespace ftl;

// =================================================================================================
// kernels
// =================================================================================================
#ifndef FTL_NO_PDL
#define FTL_NO_PDL 0   // 1: plain stream order between k_step and k_rays
#endif
#ifndef FTL_EXACT_PDL
#define FTL_EXACT_PDL 1   // k_rays_exact as a programmatic dependent launch of k_rays (hides its launch latency)
#endif
#ifndef FTL_RAYS_WARPS
#define FTL_RAYS_WARPS 2   // envs per block.  A block lives as long as its slowest env: 2-warp blocks measured 2.3 % faster
#endif                     // than 4-warp ones (and they fit sooner into what a finished k_step block frees, section 4.5)
#ifndef FTL_RAYS_MINB
#define FTL_RAYS_MINB (28 / FTL_RAYS_WARPS)   // 28 warps per SM: 72 registers, no spills; shared memory (7.2 KB per warp) allows no more
#endif
__global__ void __launch_bounds__(32 * FTL_RAYS_WARPS, FTL_RAYS_MINB)
k_rays(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, const double2* __restrict__ rot,
       float* __restrict__ rays_out, int smem_per_warp, int first_env, int end_env, int wait_seq) {
#if FTL_EXACT_PDL
    asm volatile("griddepcontrol.launch_dependents;");   // k_rays_exact may be scheduled behind the last wave of this grid
#endif
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5;
    const int i = first_env + blockIdx.x * (blockDim.x >> 5) + warp;  // one warp per env
    if (i >= end_env) return;
    if (wait_seq) {
        // launched as a programmatic dependent of k_step (see there): wait until the warp of k_step that owns this
        // env's group of 32 has published step `wait_seq`.  Bounded: a flag that never arrives is a bug, not a hang.
        if ((threadIdx.x & 31) == 0) {
            const int* flag = s.step_flag + (i >> 5);
            int seen, spins = 0;
            for (;;) {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
                if (seen == wait_seq) break;
                __nanosleep(200);
                if (++spins > (1 << 24)) __trap();
            }
        }
        __syncwarp();
    }
    RayShared& sh = *reinterpret_cast<RayShared*>(smem + (size_t)warp * smem_per_warp);
    rays_warp(cfg, s, pool, rot, i, sh, rays_out);
}

__global__ void __launch_bounds__(128)
k_rays_exact(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, float* __restrict__ rays_out,
             int first_env, int end_env) {
#if FTL_EXACT_PDL
    asm volatile("griddepcontrol.wait;" ::: "memory");   // launched early (programmatic dependent of k_rays): wait for all of it
#endif
    const int i = first_env + blockIdx.x * blockDim.x + threadIdx.x;   // one thread per env, idle unless a pair was inconclusive
    if (i >= end_env) return;
    rays_exact_env(cfg, s, pool, i, rays_out);
}

// ---- state exchange: SoA <-> FtlEnvState (AoS) ----------------------------------------------------
__global__ void k_pack_state(const DevState s, int first, int count, FtlEnvState* __restrict__ dst) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    pack_env(s, first + j, dst[j]);
}

// FollowerInfo / LeaderTrackDetector_vector outputs from the stored state (launched only when requested)
__global__ void k_optional_sensors(const __grid_constant__ DevCfg cfg, const DevState s, const DevOutputs out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < out.n) write_optional_sensors(cfg.c, s, out, i);
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
    DevPool pool{};
    bool have_pool = false, was_reset = false;
    std::vector<void*> allocs, pool_allocs;
    double* d_stats = nullptr;
    // device-side staging for the host-buffer entry points
    void* d_actions = nullptr;
    uint8_t* d_mask = nullptr;
    int* d_scen_ids = nullptr;
    DevOutputs d_out{};
    FtlEnvState* d_state_stage = nullptr;
    int state_stage_cap = 0;
    int64_t launches = 0;
    int step_seq = 0;          // sequence number of the last k_step launch (step_flag protocol)
    int rays_total = 0;
    bool rays_smem_opted = false;
    double2* d_rot = nullptr;   // (cos, sin)(k * 360/R) per flat ray
    cudaStream_t copy_stream = nullptr;   // host path: D2H copies overlap the ray kernel chunk by chunk
    cudaStream_t own_stream = nullptr;    // ftl_host_stream
    cudaStream_t pending_stream = nullptr;   // stream of an ftl_step_host_begin that has not been waited for
    bool pending = false;
    cudaEvent_t chunk_ev[8] = {};
    // optional per-kernel timing (ftl_profile): three events per step on the launching stream
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;
    size_t prof_used = 0;
};

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
    ok(dalloc(h, &s.step_flag, n / 32 + 1, list));
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

static int launch_optional_sensors(ftl_handle h, const DevOutputs& o, cudaStream_t st) {
    if (!o.follower_info && !(o.track_vectors && h->cfg.c.track_vector_len > 0) &&
        !(o.radar && h->cfg.c.radar_sectors > 0))
        return FTL_OK;
    k_optional_sensors<<<(h->n + 127) / 128, 128, 0, st>>>(h->cfg, h->st, o);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return FTL_OK;
}

static int launch_rays(ftl_handle h, float* rays, cudaStream_t st, int first_env = 0, int end_env = -1, int wait_seq = 0) {
    if (!rays || h->rays_total == 0) return FTL_OK;
    if (end_env < 0) end_env = h->n;
    if (end_env <= first_env) return FTL_OK;
    const int warps = FTL_RAYS_WARPS, threads = warps * 32;
    const int per_warp = (int)((ray_shared_bytes(h->rays_total, h->cfg.ray_hmax) + 15) & ~(size_t)15);
    const int smem = per_warp * warps;
    if (smem > 48 * 1024 && !h->rays_smem_opted) {
        CUDA_TRY(cudaFuncSetAttribute(k_rays, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        h->rays_smem_opted = true;
    }
    int blocks = (end_env - first_env + warps - 1) / warps;
    if (wait_seq) {
        // programmatic dependent launch behind k_step: the blocks start while k_step's last warps are still running
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(blocks); lc.blockDim = dim3(threads); lc.dynamicSmemBytes = smem; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        const double2* rot = h->d_rot;
        CUDA_TRY(cudaLaunchKernelEx(&lc, k_rays, h->cfg, h->st, h->pool, rot, rays, per_warp, first_env, end_env, wait_seq));
    } else {
        k_rays<<<blocks, threads, smem, st>>>(h->cfg, h->st, h->pool, h->d_rot, rays, per_warp, first_env, end_env, 0);
    }
#if FTL_EXACT_PDL
    {   // its blocks are scheduled while the last ray blocks run; griddepcontrol.wait at its top keeps the order
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3((end_env - first_env + 127) / 128); lc.blockDim = dim3(128); lc.dynamicSmemBytes = 0; lc.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        lc.attrs = at; lc.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&lc, k_rays_exact, h->cfg, h->st, h->pool, rays, first_env, end_env));
    }
#else
    k_rays_exact<<<(end_env - first_env + 127) / 128, 128, 0, st>>>(h->cfg, h->st, h->pool, rays, first_env, end_env);
#endif
    h->launches += 2;
    CUDA_TRY(cudaGetLastError());
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
        const size_t need = ((ray_shared_bytes(h->rays_total, d.ray_hmax) + 15) & ~(size_t)15) * FTL_RAYS_WARPS;
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
    DevOutputs none{};
    launch_reset(h, im, nullptr, ids, none, 1, nullptr);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
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
    const int reset_filler = (!mask_dev || !h->was_reset) ? 1 : 0;
    launch_reset(h, h->st, mask_dev, scenario_ids_dev, o, reset_filler, st);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->was_reset = true;
    int rc = launch_optional_sensors(h, o, st);
    if (rc) return rc;
    return launch_rays(h, o.rays, st);
}

int ftl_step(ftl_handle h, const void* actions_dev, const FtlOutputs* out_dev, void* cuda_stream) {
    if (!h || !actions_dev) return fail(FTL_ERR_INVALID, "NULL argument");
    if (!h->was_reset) return fail(FTL_ERR_STATE, "ftl_reset must be called before ftl_step");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    DevOutputs o = to_dev_outputs(out_dev, h->n);
    if (h->profiling) prof_event(h, st);
    const int seq = h->step_seq = (h->step_seq % 0x3fffffff) + 1;   // never 0
    switch (h->cfg.c.n_bears) {
        case 0: ftl_launch_step_nb0(h->cfg, h->st, h->pool, h->image, actions_dev, o, h->d_stats, seq, st); break;
        case 1: ftl_launch_step_nb1(h->cfg, h->st, h->pool, h->image, actions_dev, o, h->d_stats, seq, st); break;
        case 2: ftl_launch_step_nb2(h->cfg, h->st, h->pool, h->image, actions_dev, o, h->d_stats, seq, st); break;
        case 3: ftl_launch_step_nb3(h->cfg, h->st, h->pool, h->image, actions_dev, o, h->d_stats, seq, st); break;
        default: ftl_launch_step_nb4(h->cfg, h->st, h->pool, h->image, actions_dev, o, h->d_stats, seq, st); break;
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    int rc = launch_optional_sensors(h, o, st);
    if (rc) return rc;
    // The ray kernel overlaps the tail of k_step (programmatic dependent launch + per-group flags) unless something
    // was launched in between or per-kernel timing is on.
    // Not while the stream is being captured either: the sequence number is a kernel argument, so a replayed graph
    // would find last replay's flags already equal to it and skip the wait -- captured steps use plain stream order.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_TRY(cudaStreamIsCapturing(st, &cap));
    const bool overlap = !h->profiling && !FTL_NO_PDL && cap == cudaStreamCaptureStatusNone &&
                         !(o.follower_info || (o.track_vectors && h->cfg.c.track_vector_len > 0) ||
                           (o.radar && h->cfg.c.radar_sectors > 0));
    if (h->profiling) prof_event(h, st);
    rc = launch_rays(h, o.rays, st, 0, -1, overlap ? seq : 0);
    if (h->profiling) prof_event(h, st);
    return rc;
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

int ftl_step_host(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream) {
    int rc = ftl_step_host_begin(h, actions_host, out_host, cuda_stream);
    if (rc) return rc;
    return ftl_step_host_wait(h);
}

int ftl_step_host_begin(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream) {
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
    FtlOutputs o = staged_outputs(h, out_host);
    // fused step kernel, everything except the rays
    FtlOutputs no_rays = o;
    no_rays.rays = nullptr;
    int rc = ftl_step(h, h->d_actions, &no_rays, cuda_stream);
    if (rc) return rc;
    h->pending_stream = st;
    if (!out_host) {
        rc = launch_rays(h, h->d_out.rays, st);
        if (rc) return rc;
        h->pending = true;
        return FTL_OK;
    }
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
    // ray kernel in chunks of envs: the D2H copy of chunk c runs while chunk c+1 is being cast
    if (out_host->rays && h->cfg.rays_per_env) {
        const int chunks = h->n >= 8192 ? 6 : 1;
        const size_t row = sizeof(float) * (size_t)h->cfg.rays_per_env;
        for (int c = 0; c < chunks; c++) {
            int first = (int)((long long)h->n * c / chunks), end = (int)((long long)h->n * (c + 1) / chunks);
            rc = launch_rays(h, h->d_out.rays, st, first, end);
            if (rc) return rc;
            CUDA_TRY(cudaEventRecord(h->chunk_ev[1 + c], st));
            CUDA_TRY(cudaStreamWaitEvent(cs, h->chunk_ev[1 + c], 0));
            CUDA_TRY(cudaMemcpyAsync((char*)out_host->rays + row * first, (const char*)d.rays + row * first,
                                     row * (size_t)(end - first), cudaMemcpyDeviceToHost, cs));
        }
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

int ftl_profile_read(ftl_handle h, double* step_kernel_ms, double* ray_kernel_ms, int64_t* steps) {
    if (!h) return fail(FTL_ERR_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaDeviceSynchronize());
    double a = 0, b = 0;
    size_t n = h->prof_used / 3;
    for (size_t k = 0; k < n; k++) {
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[3 * k], h->prof_events[3 * k + 1]));
        a += ms;
        CUDA_TRY(cudaEventElapsedTime(&ms, h->prof_events[3 * k + 1], h->prof_events[3 * k + 2]));
        b += ms;
    }
    if (step_kernel_ms) *step_kernel_ms = a;
    if (ray_kernel_ms) *ray_kernel_ms = b;
    if (steps) *steps = (int64_t)n;
    h->prof_used = 0;
    return FTL_OK;
}

}  // extern "C"
