// ftl_book.cuh -- the bookkeeping half of a step for one env (device only), and the flag protocol between the kernels.
#pragma once

#include "ftl_reset_image.cuh"

namespace ftl {

// Wait until the warp that owns env group `group` has published sequence number `seq` on `flags` (one lane spins, the
// warp follows).  Bounded: a flag that never arrives is a bug, not a hang.
__device__ __forceinline__ void wait_group_flag(const int* flags, int group, int seq) {
    if ((threadIdx.x & 31) == 0) {
        const int* flag = flags + group;
        int seen, spins = 0;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
            if (seen == seq) break;
            __nanosleep(200);
            if (++spins > (1 << 24)) __trap();
        }
    }
    __syncwarp();
}
__device__ __forceinline__ void publish_group_flag(int* flags, int group, int seq) {
    __threadfence();
    __syncwarp();
    if ((threadIdx.x & 31) == 0)
        asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flags + group), "r"(seq) : "memory");
}


__device__ __forceinline__ void add_stat(double* stats, int which, double v) { atomicAdd(stats + which, v); }

// ENV:960-1139 for the frames the kinematics recorded, the step's reward / done / info outputs, the episode statistics
// and, for an env that finished and is renewed, the first observation of its next episode.  All 32 lanes of a warp call
// this together (the exact green-zone scans are warp collectives).
__device__ __forceinline__ void book_env(const DevCfg& cfg, const DevState& s, const DevOutputs& img_out, const DevOutputs& out,
                                         double* __restrict__ stats, int n_scenarios, int i,
                                         const FrameRec* rec_from = nullptr) {
    const FtlConfig& c = cfg.c;
    Episode e;
    episode_load_book(s, i, e);
    const bool was_done = (e.flags & FL_DONE) != 0;
    float2* trail = s.trail + (size_t)i * c.trail_cap;
    float* trail_d = s.trail_d + (size_t)i * c.trail_cap;
    double* trail_s = s.trail_s + (size_t)i * c.trail_cap;
    GreenCache gc;
    green_load(s, i, gc);
    green_cache_hydrate(trail, trail_s, e.trail_len, gc);
#ifdef FTL_PF_WINDOW
    // the window start moves when the trail grows (green_window_update reads trail_s[g_lo - g_unc ...]): ask for that line now
    if (gc.g_lo - gc.g_unc >= 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(trail_s + (gc.g_lo - gc.g_unc)));
#endif
    // the kinematics' records: where the caller put them (k_kin with the bookkeeping fused: shared memory), else HBM
    const FrameRec rec = rec_from ? *rec_from : FrameRec{s.rec_f + i, s.rec_l + i, s.rec_bits + i, s.rec_lbits + i, (size_t)s.n};
    const int fps = env_frames(cfg, s, i);
    // the warp's trip count: every lane joins the collectives of the exact scans (fps only differs between envs when the
    // caller passes FtlStepInputs.frames_per_step)
    const int nf_loop = s.in_frames ? __reduce_max_sync(0xffffffffu, fps) : fps;
    book_frames(cfg, trail, trail_d, trail_s, e, gc, rec, fps, fps, nf_loop);
    green_store(s, i, gc);
    episode_store_book(s, i, e);
    write_outputs_episode(out, i, e);
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
    // a finished env that is renewed: reward / done / info codes of the finished episode stay in `out`, the observation
    // becomes the first one of its next episode (vector-env convention).  The kinematics already wrote this step's
    // observation; the rays follow in k_finish, the state in the next step's k_kin.
    if (done && (c.auto_reset || i >= s.n_real) && i < out.n) {
        const int scen = next_scenario(cfg, n_scenarios, i, s.gi[(size_t)GI_EPISODE * s.n + i]);
        if (out.numerical_features)
            for (int k = 0; k < 10; k++) out.numerical_features[(size_t)i * 10 + k] = img_out.numerical_features[(size_t)scen * 10 + k];
        if (out.leader_target) {
            out.leader_target[2 * (size_t)i] = img_out.leader_target[2 * (size_t)scen];
            out.leader_target[2 * (size_t)i + 1] = img_out.leader_target[2 * (size_t)scen + 1];
        }
    }
}

}  // namespace ftl
