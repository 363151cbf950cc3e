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

#ifndef FTL_RESET_COPY_SERIAL
// Loads first, stores afterwards: the compiler cannot prove that the image and the state do not alias, so a copy written
// item by item is a chain of a dozen dependent L2 round trips (measured: 11 us in the prologue of one warp in five --
// the warps that then finish last).  Every group has at most 32 items (one per lane); the rows take one or two rounds.
static_assert(GI_COUNT <= 32 && GD_COUNT <= 32 && GF_COUNT <= 32 && kMaxRobots * RD_COUNT <= 32, "one item per lane");
__device__ __forceinline__ void reset_state_from_image(const FtlConfig& c, const DevState& s, const DevState& img, int env,
                                                       int scen, int lane) {
    const size_t n = s.n, m = img.n;
    const int nb = s.n_bears, nr = 2 + nb;
    // ---- round 1: every scalar group, one item per lane ------------------------------------------------------------
    int v_gi = 0, v_ri = 0, v_bidx = 0, v_episode = 0;
    double v_gd = 0, v_rd = 0, v_btgt = 0;
    float v_gf = 0;
    float2 v_pos = make_float2(0.f, 0.f);
    int4 v_rect = make_int4(0, 0, 0, 0), v_srect = make_int4(0, 0, 0, 0);
    int2 v_srange = make_int2(0, 0);
    if (lane < GI_COUNT) v_gi = img.gi[(size_t)lane * m + scen];
    if (lane == GI_EPISODE) v_episode = s.gi[(size_t)lane * n + env];
    if (lane < GD_COUNT) v_gd = img.gd[(size_t)lane * m + scen];
    if (lane < GF_COUNT) v_gf = img.gf[(size_t)lane * m + scen];
    if (lane < nr) {
        v_pos = img.pos[(size_t)lane * m + scen];
        v_rect = img.rect[(size_t)lane * m + scen];
        v_ri = img.ri[(size_t)lane * m + scen];
    }
    if (lane < nr * RD_COUNT) v_rd = img.rd[(size_t)lane * m + scen];
    if (lane < nb * 2) v_btgt = img.bear_tgt[(size_t)lane * m + scen];
    if (lane < nb) v_bidx = img.bear_idx[(size_t)lane * m + scen];
    if (lane == 0) v_srange = img.snap_range[scen];
    if (lane < 1 + nb) v_srect = img.snap_rect[(size_t)lane * m + scen];
    const int snap_pushes = __shfl_sync(0xffffffffu, v_gi, GI_SNAP_PUSHES);
    const int trail_len = __shfl_sync(0xffffffffu, v_gi, GI_TRAIL_LEN), ring_head = __shfl_sync(0xffffffffu, v_gi, GI_RING_HEAD);
    // ---- round 2: the first 32 items of the rows (a reset leaves a short trail seed and one or two ring entries) ------
    const size_t to = (size_t)env * c.trail_cap, from = (size_t)scen * c.trail_cap;
    const size_t rto = (size_t)env * c.corridor_cap, rfrom = (size_t)scen * c.corridor_cap;
    const int nring = ring_head < c.corridor_cap ? ring_head : c.corridor_cap;
    float2 t_p = make_float2(0.f, 0.f);
    float t_d = 0.f, r_f = 0.f;
    double t_s = 0, r_d = 0;
    double2 r_h = make_double2(0, 0);
    float4 r_c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < trail_len) { t_p = img.trail[from + lane]; t_d = img.trail_d[from + lane]; t_s = img.trail_s[from + lane]; }
    if (lane < nring) { r_h = img.hist[rfrom + lane]; r_c = img.corridor[rfrom + lane]; r_d = img.seg_d[rfrom + lane]; r_f = img.seg_f[rfrom + lane]; }
    // ---- stores ----------------------------------------------------------------------------------------------------
    if (lane < GI_COUNT && lane != GI_ACCEL_CONSUMED)                // never restored by the reference, ENV:1170
        s.gi[(size_t)lane * n + env] = lane == GI_EPISODE ? v_episode + 1 : v_gi;
    if (lane < GD_COUNT) s.gd[(size_t)lane * n + env] = v_gd;
    if (lane < GF_COUNT) s.gf[(size_t)lane * n + env] = v_gf;
    if (lane < nr) {
        s.pos[(size_t)lane * n + env] = v_pos;
        s.rect[(size_t)lane * n + env] = v_rect;
        s.ri[(size_t)lane * n + env] = v_ri;
    }
    if (lane < nr * RD_COUNT) s.rd[(size_t)lane * n + env] = v_rd;
    if (lane < nb * 2) s.bear_tgt[(size_t)lane * n + env] = v_btgt;
    if (lane < nb) s.bear_idx[(size_t)lane * n + env] = v_bidx;
    if (snap_pushes > 0) {   // the reset's own sensor pass pushed history entry 0
        if (lane == 0) s.snap_range[env] = v_srange;
        if (lane < 1 + nb) s.snap_rect[(size_t)lane * n + env] = v_srect;
    }
    if (lane < trail_len) { s.trail[to + lane] = t_p; s.trail_d[to + lane] = t_d; s.trail_s[to + lane] = t_s; }
    if (lane < nring) { s.hist[rto + lane] = r_h; s.corridor[rto + lane] = r_c; s.seg_d[rto + lane] = r_d; s.seg_f[rto + lane] = r_f; }
    // ---- the rest of longer rows ---------------------------------------------------------------------------------------
    for (int k = lane + 32; k < trail_len; k += 32) {
        const float2 p = img.trail[from + k];
        const float d = img.trail_d[from + k];
        const double sk = img.trail_s[from + k];
        s.trail[to + k] = p; s.trail_d[to + k] = d; s.trail_s[to + k] = sk;
    }
    for (int k = lane + 32; k < nring; k += 32) {
        const double2 h = img.hist[rfrom + k];
        const float4 cc = img.corridor[rfrom + k];
        const double d = img.seg_d[rfrom + k];
        const float f = img.seg_f[rfrom + k];
        s.hist[rto + k] = h; s.corridor[rto + k] = cc; s.seg_d[rto + k] = d; s.seg_f[rto + k] = f;
    }
}
#else
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

#endif

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
