// ftl_step_nb.cu -- the fused step / reset kernels for ONE value of NB (number of dynamic obstacles),
// compiled once per NB in 0..FTL_MAX_BEARS with -DFTL_NB=<n> (build.py) so the robots live in registers.
//
//   k_step<NB>   one thread per env: the F sub-frames fused in registers (robots, collisions, exact
//                green-zone flags from cached bounds, reward/done bookkeeping), the two tracker scans,
//                the history snapshot, non-ray outputs, episode statistics, optional in-place auto-reset
//   k_reset<NB>  masked re-initialisation from the scenario pool
#include <cuda_runtime.h>

#include "ftl_launch.h"
#include "ftl_step.cuh"

#ifndef FTL_NB
#error "compile with -DFTL_NB=<number of bears>"
#endif

using namespace ftl;

__device__ __forceinline__ void add_stat(double* stats, int which, double v) { atomicAdd(stats + which, v); }

// The in-step auto-reset is a cold path: kept out of line with its own registers so that it neither grows the
// hot loop nor forces the step's state through local memory.
template <int NB>
__device__ __noinline__ void reset_in_place(const DevCfg& cfg, const DevState& s, const DevPool& pool,
                                            const DevOutputs& out, int i, int episode) {
    World<NB> w;
    Episode e;
    int scen = next_scenario(cfg, pool.n_scenarios, i, episode);
    env_reset<NB>(cfg, s, pool, i, scen, w, e);
    write_outputs<NB>(cfg, pool, out, i, w, e, true);
    world_store<NB>(s, i, w);
    episode_store(s, i, e);
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
       const void* __restrict__ actions, const __grid_constant__ DevOutputs out, double* __restrict__ stats) {
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
    if ((cfg.c.auto_reset || i >= s.n_real) && done) {
        // reward/done/status of the finished episode stay in `out`; the observation becomes the first
        // one of the next episode (vector-env convention).  accel_consumed / episode are read back from the
        // arrays by env_reset, so store them first.
        s.gi[(size_t)GI_ACCEL_CONSUMED * s.n + i] = e.accel_consumed;
        s.gi[(size_t)GI_EPISODE * s.n + i] = e.episode;
        reset_in_place<NB>(cfg, s, pool, out, i, e.episode);
    } else {
        world_store<NB>(s, i, w);
        episode_store(s, i, e);
    }
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

void FTL_CAT(ftl_launch_step_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const void* actions,
                                          const DevOutputs& out, double* stats, cudaStream_t st) {
    int threads = FTL_STEP_THREADS, blocks = (s.n + threads - 1) / threads;
    k_step<FTL_NB><<<blocks, threads, 0, st>>>(cfg, s, pool, actions, out, stats);
}
void FTL_CAT(ftl_launch_reset_nb, FTL_NB)(const DevCfg& cfg, const DevState& s, const DevPool& pool, const uint8_t* mask,
                                           const int* ids, const DevOutputs& out, int reset_filler, cudaStream_t st) {
    int threads = 128, blocks = (s.n + threads - 1) / threads;
    k_reset<FTL_NB><<<blocks, threads, 0, st>>>(cfg, s, pool, mask, ids, out, reset_filler);
}
