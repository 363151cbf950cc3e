// ftl_reset_image.cuh -- in-step auto-reset as a copy (device only).
//
// What Game.reset() leaves behind depends on the scenario alone, so ftl_upload_scenarios runs k_reset and the ray
// kernels once per scenario into an "image" DevState (env s of the image = scenario s after reset) plus its first
// observation (numerical_features, leader_target, rays).  An env that finishes (done && auto_reset) is handled in three
// places, none of them on the critical path of the step:
//   k_book    (the kernel that finds out) replaces the env's observation outputs by the image's numerical_features /
//             leader_target -- reward, done and info codes of the finished episode stay (vector-env convention);
//   k_finish  copies the image's rays over the rows the ray kernel cast for the finished episode's last pose;
//   k_kin     of the NEXT step (or k_apply_resets, when the host reads or writes state in between) copies the state:
//             scalars and robots one item per lane, the rows (trail seed, tracker ring, first history snapshot)
//             coalesced, all by the whole warp.  With random actions about one env in 130 finishes per step, i.e. every
//             fifth warp has one.
#pragma once

#include "ftl_step.cuh"

namespace ftl {

// the scenario an env whose flags say "finished" moves to, or -1 when it stays as it is
__device__ __forceinline__ int pending_reset_scenario(const DevCfg& cfg, const DevState& s, int n_scenarios, int i) {
    const int flags = s.gi[(size_t)GI_FLAGS * s.n + i];
    if (!(flags & FL_DONE) || !(cfg.c.auto_reset || i >= s.n_real)) return -1;
    return next_scenario(cfg, n_scenarios, i, s.gi[(size_t)GI_EPISODE * s.n + i]);
}

__device__ __forceinline__ void reset_state_from_image(const FtlConfig& c, const DevState& s, const DevState& img, int env,
                                                       int scen, int lane) {
    const size_t n = s.n, m = img.n;
    for (int k = lane; k < GI_COUNT; k += 32) {
        int v = img.gi[(size_t)k * m + scen];
        if (k == GI_ACCEL_CONSUMED) continue;                        // never restored by the reference, ENV:1170
        if (k == GI_EPISODE) v = s.gi[(size_t)k * n + env] + 1;
        s.gi[(size_t)k * n + env] = v;
    }
    for (int k = lane; k < GD_COUNT; k += 32) s.gd[(size_t)k * n + env] = img.gd[(size_t)k * m + scen];
    for (int k = lane; k < GF_COUNT; k += 32) s.gf[(size_t)k * n + env] = img.gf[(size_t)k * m + scen];
    const int nb = s.n_bears, nr = 2 + nb;
    for (int k = lane; k < nr; k += 32) {
        s.pos[(size_t)k * n + env] = img.pos[(size_t)k * m + scen];
        s.rect[(size_t)k * n + env] = img.rect[(size_t)k * m + scen];
        s.ri[(size_t)k * n + env] = img.ri[(size_t)k * m + scen];
    }
    for (int k = lane; k < nr * RD_COUNT; k += 32) s.rd[(size_t)k * n + env] = img.rd[(size_t)k * m + scen];
    for (int k = lane; k < nb * 2; k += 32) s.bear_tgt[(size_t)k * n + env] = img.bear_tgt[(size_t)k * m + scen];
    for (int k = lane; k < nb; k += 32) s.bear_idx[(size_t)k * n + env] = img.bear_idx[(size_t)k * m + scen];
    const int snap_pushes = img.gi[(size_t)GI_SNAP_PUSHES * m + scen];
    if (snap_pushes > 0) {   // the reset's own sensor pass pushed history entry 0
        if (lane == 0) s.snap_range[env] = img.snap_range[scen];
        for (int k = lane; k < 1 + nb; k += 32) s.snap_rect[(size_t)k * n + env] = img.snap_rect[(size_t)k * m + scen];
    }
    const int trail_len = img.gi[(size_t)GI_TRAIL_LEN * m + scen], ring_head = img.gi[(size_t)GI_RING_HEAD * m + scen];
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

// all 32 lanes of a warp call this with their env; afterwards the stored state of every finished env of the warp is the
// first state of its next episode
__device__ __forceinline__ void apply_pending_resets(const DevCfg& cfg, const DevState& s, const DevPool& pool,
                                                     const DevState& img, int i) {
    const int scen = pending_reset_scenario(cfg, s, pool.n_scenarios, i);
    unsigned todo = __ballot_sync(0xffffffffu, scen >= 0);
    if (!todo) return;
    const int lane = threadIdx.x & 31;
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        reset_state_from_image(cfg.c, s, img, __shfl_sync(0xffffffffu, i, src), __shfl_sync(0xffffffffu, scen, src), lane);
    }
    __syncwarp();   // the copies of the other lanes are this lane's state
}

}  // namespace ftl
