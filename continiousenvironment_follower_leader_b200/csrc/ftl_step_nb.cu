// ftl_step_nb.cu -- the fused step / reset kernels for ONE value of NB (number of dynamic obstacles),
// compiled once per NB in 0..FTL_MAX_BEARS with -DFTL_NB=<n> (build.py) so the robots live in registers.
//
//   k_step<NB>   one thread per env: the F sub-frames fused in registers (robots, collisions, exact
//                green-zone flags from cached bounds, reward/done bookkeeping), the two tracker scans,
//                the history snapshot, non-ray outputs, episode statistics, optional in-place auto-reset (a copy
//                of the scenario's reset image)
//   k_reset<NB>  masked re-initialisation from the scenario pool
#include <cuda_runtime.h>

#include "ftl_launch.h"

#ifndef FTL_SYNC_EVERY
#define FTL_SYNC_EVERY 0
#endif
#if FTL_SYNC_EVERY > 0
__host__ __device__ static __forceinline__ void ftl_frame_sync(int f) {
#if defined(__CUDA_ARCH__)
    if (f % FTL_SYNC_EVERY == 0) __syncthreads();
#endif
}
#define FTL_FRAME_SYNC(f) ftl_frame_sync(f)
#endif
#include "ftl_step.cuh"

#ifndef FTL_NB
#error "compile with -DFTL_NB=<number of bears>"
#endif

using namespace ftl;

__device__ __forceinline__ void add_stat(double* stats, int which, double v) { atomicAdd(stats + which, v); }

// In-step auto-reset.  What Game.reset() leaves behind depends on the scenario alone, so ftl_upload_scenarios runs
// k_reset once per scenario into an "image" DevState (env s of the image = scenario s after reset) and the step kernel
// only copies: the scalars by the lane that owns the env (same load/store helpers as the step itself), the rows
// (trail seed, tracker ring, first history snapshot) by the whole warp, coalesced.  The serial reset in one lane used
// to be the tail of the kernel: with random actions ~1 env in 130 finishes per step, i.e. every fifth warp.
template <int NB>
__device__ __forceinline__ void reset_scalars_from_image(const DevCfg& cfg, const DevState& s, const DevPool& pool,
                                                         const DevState& img, const DevOutputs& out, int i, int scen,
                                                         World<NB>& w, Episode& e, int* ring_head) {
    const int accel_consumed = e.accel_consumed, episode = e.episode;
    world_load<NB>(img, scen, w);
    episode_load(img, scen, e);
    e.accel_consumed = accel_consumed;   // never restored by the reference, ENV:1170
    e.episode = episode + 1;
    GreenCache gc;
    Tracker t;
    int snap_pushes;
    cache_load(img, scen, gc, t, &snap_pushes);
    cache_store(s, i, gc, t, snap_pushes);
    if (snap_pushes > 0) {   // the reset's own sensor pass pushed history entry 0
        s.snap_range[i] = img.snap_range[scen];
#pragma unroll
        for (int k = 0; k < 1 + NB; k++) s.snap_rect[(size_t)k * s.n + i] = img.snap_rect[(size_t)k * img.n + scen];
    }
    *ring_head = t.ring_head;
    write_outputs<NB>(cfg, pool, out, i, w, e, true);
}

__device__ __forceinline__ void reset_rows_from_image(const FtlConfig& c, const DevState& s, const DevState& img, int env,
                                                      int scen, int trail_len, int ring_head, int lane) {
    const size_t to = (size_t)env * c.trail_cap, from = (size_t)scen * c.trail_cap;
    for (int k = lane; k < trail_len; k += 32) {
        s.trail[to + k] = img.trail[from + k];
        s.trail_d[to + k] = img.trail_d[from + k];
        s.trail_s[to + k] = img.trail_s[from + k];
    }
    const size_t rto = (size_t)env * c.corridor_cap, rfrom = (size_t)scen * c.corridor_cap;
    const int nring = ring_head < c.corridor_cap ? ring_head : c.corridor_cap;
    for (int k = lane; k < nring; k += 32) {
        s.hist[rto + k] = img.hist[rfrom + k];
        s.corridor[rto + k] = img.corridor[rfrom + k];
        s.seg_d[rto + k] = img.seg_d[rfrom + k];
        s.seg_f[rto + k] = img.seg_f[rfrom + k];
    }
}

#ifndef FTL_STEP_THREADS
#define FTL_STEP_THREADS 64   // 65536 envs are only ~14 warps per SM: small blocks spread them evenly over the 148 SMs
#endif
#ifndef FTL_STEP_MINBLOCKS
#define FTL_STEP_MINBLOCKS 8    // caps the kernel at 128 registers: all ~14 warps/SM of a 65536-env batch resident in one wave
#endif                            // (measured: 128 regs 0.276 ms, 168 regs 0.31 ms, 210 regs 0.325 ms, 96 regs 0.29 ms per step)
template <int NB>
__global__ void __launch_bounds__(FTL_STEP_THREADS, FTL_STEP_MINBLOCKS)
k_step(const __grid_constant__ DevCfg cfg, const __grid_constant__ DevState s, const __grid_constant__ DevPool pool,
       const __grid_constant__ DevState img, const void* __restrict__ actions, const __grid_constant__ DevOutputs out,
       double* __restrict__ stats, int seq) {
    // The ray kernel is launched as a programmatic dependent of this one: its blocks may start as soon as every block
    // of this grid is running, take the SM resources that finished blocks free, and wait per env group on step_flag.
    // This kernel is one wave whose end is set by its slowest warps (exact green-zone scans, resets); the ray kernel
    // fills that tail.
    asm volatile("griddepcontrol.launch_dependents;");
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    World<NB> w;
    Episode e;
    world_load<NB>(s, i, w);
    episode_load(s, i, e);
    const bool was_done = (e.flags & FL_DONE) != 0;
    double a0, a1;
    decode_action(cfg.c, actions, i, s.n_real, &a0, &a1);
    env_step<NB>(cfg, s, pool, i, a0, a1, w, e);
    write_outputs<NB>(cfg, pool, out, i, w, e, false);
    const bool done = (e.flags & FL_DONE) != 0;
    if (done && !was_done && i < s.n_real) {  // episode statistics (reduced across ranks with NCCL by the host)
        add_stat(stats, FTL_STAT_EPISODES, 1.0);
        add_stat(stats, FTL_STAT_RETURN_SUM, e.overall);
        add_stat(stats, FTL_STAT_LENGTH_SUM, (double)e.step_count);
        int mission = (e.flags >> FL_MISSION_SHIFT) & 3, leader_st = (e.flags >> FL_LEADER_SHIFT) & 3;
        if (e.flags & FL_CRASH) add_stat(stats, FTL_STAT_CRASH, 1.0);
        if (mission == FTL_MISSION_SUCCESS) add_stat(stats, FTL_STAT_SUCCESS, 1.0);
        if (mission == FTL_MISSION_FINISHED_BY_TIME) add_stat(stats, FTL_STAT_TIMEOUT, 1.0);
        if (leader_st == FTL_LEADER_CRASH) add_stat(stats, FTL_STAT_LEADER_CRASH, 1.0);
        if (e.overflow) add_stat(stats, FTL_STAT_OVERFLOW, 1.0);
    }
    // reward/done/status of the finished episode stay in `out`; the observation becomes the first one of the next
    // episode (vector-env convention)
    const bool renew = (cfg.c.auto_reset || i >= s.n_real) && done;
    int scen = 0, ring_head = 0;
    if (renew) {
        scen = next_scenario(cfg, pool.n_scenarios, i, e.episode);
        reset_scalars_from_image<NB>(cfg, s, pool, img, out, i, scen, w, e, &ring_head);
    }
    world_store<NB>(s, i, w);
    episode_store(s, i, e);
    unsigned todo = __ballot_sync(0xffffffffu, renew);   // whole warps reach this point (n is padded to 32)
    const int lane = threadIdx.x & 31;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        reset_rows_from_image(cfg.c, s, img, __shfl_sync(0xffffffffu, i, src), __shfl_sync(0xffffffffu, scen, src),
                              __shfl_sync(0xffffffffu, e.trail_len, src), __shfl_sync(0xffffffffu, ring_head, src), lane);
    }
    // publish: every lane's stores, then the warp's flag (release); k_rays acquires it before reading this group's state
    __threadfence();
    __syncwarp();
    if (lane == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(s.step_flag + (i >> 5)), "r"(seq) : "memory");
}

template <int NB>
__global__ void __launch_bounds__(128)
k_reset(const __grid_constant__ DevCfg cfg, const DevState s, const DevPool pool, const uint8_t* __restrict__ mask,
        const int* __restrict__ scenario_ids, const DevOutputs out, int reset_filler) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= s.n) return;
    if (i < s.n_real ? (mask && !mask[i]) : !reset_filler) return;
    World<NB> w;
    Episode e;
    int episodes = s.gi[(size_t)GI_EPISODE * s.n + i];
    int scen = (scenario_ids && i < s.n_real) ? scenario_ids[i] : next_scenario(cfg, pool.n_scenarios, i, episodes);
    scen = scen < 0 ? 0 : scen >= pool.n_scenarios ? pool.n_scenarios - 1 : scen;
    env_reset<NB>(cfg, s, pool, i, scen, w, e);
    write_outputs<NB>(cfg, pool, out, i, w, e, false);
    world_store<NB>(s, i, w);
    episode_store(s, i, e);
}


#define FTL_CAT2(a, b) a##b
#define FTL_CAT(a, b) FTL_CAT2(a, b)

void FTL_CAT(ftl_launch_step_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const DevState& img,
                                          const void* actions, const DevOutputs& out, double* stats, int seq,
                                          cudaStream_t st) {
#ifdef FTL_STEP_LAUNCH_THREADS
    int threads = FTL_STEP_LAUNCH_THREADS;
#else
    int threads = FTL_STEP_THREADS;
#endif
    int blocks = (s.n + threads - 1) / threads;
    k_step<FTL_NB><<<blocks, threads, 0, st>>>(cfg, s, pool, img, actions, out, stats, seq);
}
void FTL_CAT(ftl_launch_reset_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const uint8_t* mask,
                                           const int* ids, const DevOutputs& out, int reset_filler, cudaStream_t st) {
    int threads = 128, blocks = (s.n + threads - 1) / threads;
    k_reset<FTL_NB><<<blocks, threads, 0, st>>>(cfg, s, pool, mask, ids, out, reset_filler);
}
