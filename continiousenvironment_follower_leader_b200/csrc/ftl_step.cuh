// ftl_step.cuh -- one environment step (F fused sub-frames + tracker + bookkeeping), reset and outputs.
// Host+device like ftl_device.cuh; NB (number of dynamic obstacles) is a template parameter so that
// the robots live in registers.
#pragma once

#include "ftl_device.cuh"

// Hook at the top of every sub-frame: the step kernel may re-align the warps of a block there (instruction-cache
// locality: the frame loop is larger than the SM's instruction cache); nothing elsewhere.
#ifndef FTL_FRAME_SYNC
#define FTL_FRAME_SYNC(f) ((void)0)
#endif

namespace ftl {

FTL_HD void episode_load(const DevState& s, int i, Episode& e) {
    const int* g = s.gi + i;
    size_t n = s.n;
    e.step_count = g[GI_STEP_COUNT * n];
    e.cur_target_id = g[GI_TARGET_ID * n];
    e.finish_timer = g[GI_FINISH_TIMER * n];
    e.flags = g[GI_FLAGS * n];
    e.trail_len = g[GI_TRAIL_LEN * n];
    e.scenario = g[GI_SCENARIO * n];
    e.episode = g[GI_EPISODE * n];
    e.overflow = g[GI_OVERFLOW * n];
    e.accel_consumed = g[GI_ACCEL_CONSUMED * n];
    const double* d = s.gd + i;
    e.acc_penalty = d[GD_ACC_PENALTY * n];
    e.overall = d[GD_OVERALL * n];
    e.last_reward = d[GD_LAST_REWARD * n];
    e.speed_mult = d[GD_SPEED_MULT * n];
    e.lead_acc = d[GD_LEAD_ACC * n];
    e.lead_cum = d[GD_LEAD_CUM * n];
}
FTL_HD void episode_store(const DevState& s, int i, const Episode& e) {
    int* g = s.gi + i;
    size_t n = s.n;
    g[GI_STEP_COUNT * n] = e.step_count;
    g[GI_TARGET_ID * n] = e.cur_target_id;
    g[GI_FINISH_TIMER * n] = e.finish_timer;
    g[GI_FLAGS * n] = e.flags;
    g[GI_TRAIL_LEN * n] = e.trail_len;
    g[GI_SCENARIO * n] = e.scenario;
    g[GI_EPISODE * n] = e.episode;
    g[GI_OVERFLOW * n] = e.overflow;
    g[GI_ACCEL_CONSUMED * n] = e.accel_consumed;
    double* d = s.gd + i;
    d[GD_ACC_PENALTY * n] = e.acc_penalty;
    d[GD_OVERALL * n] = e.overall;
    d[GD_LAST_REWARD * n] = e.last_reward;
    d[GD_SPEED_MULT * n] = e.speed_mult;
    d[GD_LEAD_ACC * n] = e.lead_acc;
    d[GD_LEAD_CUM * n] = e.lead_cum;
}
FTL_HD void green_load(const DevState& s, int i, GreenCache& gc) {
    const int* g = s.gi + i;
    size_t n = s.n;
    gc.g_lo = g[GI_G_LO * n];
    gc.g_unc = g[GI_G_UNC * n];
    gc.a_star = g[GI_A_STAR * n];
    gc.b_star = g[GI_B_STAR * n];
    gc.nice_from = g[GI_NICE_FROM * n];
    gc.lb_g = s.gf[GF_LB_GREEN * n + i];
    gc.lb_all = s.gf[GF_LB_ALL * n + i];
    gc.sgx = s.gf[GF_SCAN_GX * n + i];
    gc.sgy = s.gf[GF_SCAN_GY * n + i];
    gc.sax = s.gf[GF_SCAN_AX * n + i];
    gc.say = s.gf[GF_SCAN_AY * n + i];
}
FTL_HD void green_store(const DevState& s, int i, const GreenCache& gc) {
    int* g = s.gi + i;
    size_t n = s.n;
    g[GI_G_LO * n] = gc.g_lo;
    g[GI_G_UNC * n] = gc.g_unc;
    g[GI_A_STAR * n] = gc.a_star;
    g[GI_B_STAR * n] = gc.b_star;
    g[GI_NICE_FROM * n] = gc.nice_from;
    s.gf[GF_LB_GREEN * n + i] = gc.lb_g;
    s.gf[GF_LB_ALL * n + i] = gc.lb_all;
    s.gf[GF_SCAN_GX * n + i] = gc.sgx;
    s.gf[GF_SCAN_GY * n + i] = gc.sgy;
    s.gf[GF_SCAN_AX * n + i] = gc.sax;
    s.gf[GF_SCAN_AY * n + i] = gc.say;
}
FTL_HD void tracker_load(const DevState& s, int i, Tracker& t, int* snap_pushes) {
    const int* g = s.gi + i;
    size_t n = s.n;
    t.saving_counter = g[GI_SAVING_COUNTER * n];
    t.ring_tail = g[GI_RING_TAIL * n];
    t.ring_head = g[GI_RING_HEAD * n];
    t.hist_f64_end = g[GI_HIST_F64_END * n];
    *snap_pushes = g[GI_SNAP_PUSHES * n];
}
FTL_HD void tracker_store(const DevState& s, int i, const Tracker& t, int snap_pushes) {
    int* g = s.gi + i;
    size_t n = s.n;
    g[GI_SAVING_COUNTER * n] = t.saving_counter;
    g[GI_RING_TAIL * n] = t.ring_tail;
    g[GI_RING_HEAD * n] = t.ring_head;
    g[GI_HIST_F64_END * n] = t.hist_f64_end;
    g[GI_SNAP_PUSHES * n] = snap_pushes;
}
FTL_HD void cache_load(const DevState& s, int i, GreenCache& gc, Tracker& t, int* snap_pushes) {
    green_load(s, i, gc);
    tracker_load(s, i, t, snap_pushes);
}
FTL_HD void cache_store(const DevState& s, int i, const GreenCache& gc, const Tracker& t, int snap_pushes) {
    green_store(s, i, gc);
    tracker_store(s, i, t, snap_pushes);
}

// The two kernels of a step own disjoint halves of the episode record: the kinematics kernel the leader's route cursor
// and regime state (it only READS the frame counter and the flags), the bookkeeping kernel everything else.
FTL_HD void episode_load_kin(const DevState& s, int i, Episode& e) {
    const int* g = s.gi + i;
    size_t n = s.n;
    e.step_count = g[GI_STEP_COUNT * n];
    e.cur_target_id = g[GI_TARGET_ID * n];
    e.flags = g[GI_FLAGS * n];
    e.scenario = g[GI_SCENARIO * n];
    e.episode = g[GI_EPISODE * n];
    e.accel_consumed = g[GI_ACCEL_CONSUMED * n];
    e.overflow = 0;   // bits raised by this step's tracker scans; merged into the stored word by episode_store_kin
    const double* d = s.gd + i;
    e.speed_mult = d[GD_SPEED_MULT * n];
    e.lead_acc = d[GD_LEAD_ACC * n];
    e.lead_cum = d[GD_LEAD_CUM * n];
    e.finish_timer = -1; e.trail_len = 0;
    e.acc_penalty = e.overall = e.last_reward = 0.0;
}
FTL_HD void episode_store_kin(const DevState& s, int i, const Episode& e) {
    int* g = s.gi + i;
    size_t n = s.n;
    g[GI_TARGET_ID * n] = e.cur_target_id;
    g[GI_ACCEL_CONSUMED * n] = e.accel_consumed;
    if (e.overflow) g[GI_OVERFLOW * n] |= e.overflow;
    double* d = s.gd + i;
    d[GD_SPEED_MULT * n] = e.speed_mult;
    d[GD_LEAD_ACC * n] = e.lead_acc;
    d[GD_LEAD_CUM * n] = e.lead_cum;
}
FTL_HD void episode_load_book(const DevState& s, int i, Episode& e) {
    const int* g = s.gi + i;
    size_t n = s.n;
    e.step_count = g[GI_STEP_COUNT * n];
    e.finish_timer = g[GI_FINISH_TIMER * n];
    e.flags = g[GI_FLAGS * n];
    e.trail_len = g[GI_TRAIL_LEN * n];
    e.overflow = g[GI_OVERFLOW * n];
    const double* d = s.gd + i;
    e.acc_penalty = d[GD_ACC_PENALTY * n];
    e.overall = d[GD_OVERALL * n];
    e.last_reward = d[GD_LAST_REWARD * n];
    e.cur_target_id = 0; e.scenario = 0; e.episode = 0; e.accel_consumed = 0;
    e.speed_mult = e.lead_acc = e.lead_cum = 0.0;
}
FTL_HD void episode_store_book(const DevState& s, int i, const Episode& e) {
    int* g = s.gi + i;
    size_t n = s.n;
    g[GI_STEP_COUNT * n] = e.step_count;
    g[GI_FINISH_TIMER * n] = e.finish_timer;
    g[GI_FLAGS * n] = e.flags;
    g[GI_TRAIL_LEN * n] = e.trail_len;
    g[GI_OVERFLOW * n] = e.overflow;
    double* d = s.gd + i;
    d[GD_ACC_PENALTY * n] = e.acc_penalty;
    d[GD_OVERALL * n] = e.overall;
    d[GD_LAST_REWARD * n] = e.last_reward;
}

FTL_HD int2 route_point(const DevPool& pool, const FtlConfig& c, int scenario, int k) {
    return pool.route[(size_t)scenario * c.route_cap + k];
}

constexpr int kFrameChunk = 10;   // frames per pass (frames_per_step = 10 in one go)
enum { REC_FOLLOWER_HIT = 1, REC_TOO_CLOSE = 2, REC_LEADER_HIT = 4, REC_LEADER_FINISHED = 8 };

template <int NB>
struct World {  // registers of one env during a step
    Robot follower, leader;
    Robot bear[NB > 0 ? NB : 1];
    double btx[NB > 0 ? NB : 1], bty[NB > 0 ? NB : 1];
    int bidx[NB > 0 ? NB : 1];
};

template <int NB>
FTL_HD void world_load(const DevState& s, int i, World<NB>& w) {
    robot_load(s, 0, i, w.follower);
    robot_load(s, 1, i, w.leader);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        robot_load(s, 2 + b, i, w.bear[b]);
        w.btx[b] = s.bear_tgt[((size_t)b * 2 + 0) * s.n + i];
        w.bty[b] = s.bear_tgt[((size_t)b * 2 + 1) * s.n + i];
        w.bidx[b] = s.bear_idx[(size_t)b * s.n + i];
    }
}
template <int NB>
FTL_HD void world_store(const DevState& s, int i, const World<NB>& w) {
    robot_store(s, 0, i, w.follower);
    robot_store(s, 1, i, w.leader);
#pragma unroll
    for (int b = 0; b < NB; b++) {
        robot_store(s, 2 + b, i, w.bear[b]);
        s.bear_tgt[((size_t)b * 2 + 0) * s.n + i] = w.btx[b];
        s.bear_tgt[((size_t)b * 2 + 1) * s.n + i] = w.bty[b];
        s.bear_idx[(size_t)b * s.n + i] = w.bidx[b];
    }
}

// ---- sensors that are cheap and serial: tracker scans + history snapshot (CLS:255-288, SEN:894-895) ----------
FTL_HD void sense_serial_in(const DevCfg& cfg, const DevState& s, int i, const TrackerInput& in, int4 leader_rect,
                            const int4* bear_rects, int nb, Tracker& t, int* snap_pushes, int* overflow) {
    const FtlConfig& c = cfg.c;
    if (!c.tracker_enabled) return;
    double2* hist = s.hist + (size_t)i * c.corridor_cap;
    float4* corr = s.corridor + (size_t)i * c.corridor_cap;
    double* seg_d = s.seg_d + (size_t)i * c.corridor_cap;
    float* seg_f = s.seg_f + (size_t)i * c.corridor_cap;
    for (int k = 0; k < c.tracker_scans_per_step; k++) tracker_scan(cfg, t, hist, corr, seg_d, seg_f, in, overflow);
    if (c.n_ray_sensors > 0 && t.ring_head - t.ring_tail > 1) {
        int slot = *snap_pushes % FTL_MAX_HIST;
        s.snap_range[(size_t)slot * s.n + i] = make_int2(t.ring_tail, t.ring_head);
        int4* sr = s.snap_rect + ((size_t)slot * (1 + nb)) * s.n + i;
        sr[0] = leader_rect;
        for (int b = 0; b < nb; b++) sr[(size_t)(1 + b) * s.n] = bear_rects[b];
        *snap_pushes += 1;
    }
}
template <int NB>
FTL_HD void sense_serial(const DevCfg& cfg, const DevState& s, int i, const World<NB>& w, Tracker& t, int* snap_pushes,
                         int* overflow) {
    TrackerInput in = {w.follower.px, w.follower.py, w.leader.px, w.leader.py, w.follower.dir};
    int4 bears[NB > 0 ? NB : 1];
#pragma unroll
    for (int b = 0; b < NB; b++) bears[b] = make_int4(w.bear[b].rx, w.bear[b].ry, w.bear[b].rw, w.bear[b].rh);
    sense_serial_in(cfg, s, i, in, make_int4(w.leader.rx, w.leader.ry, w.leader.rw, w.leader.rh), bears, NB, t,
                    snap_pushes, overflow);
}

// FollowerInfo.scan (SEN:834-842: float64 quotients stored as float32) and LeaderTrackDetector_vector.scan
// (SEN:365-380) on the tracker ring as the step left it: np.array(slice) - position with a float32 position, cast
// into the float32 buffer.  Reads the stored state after the step / reset kernel (its own small kernel, launched
// only when one of these outputs is requested: the step kernel of the BASELINE.json configurations is untouched).
// LeaderTrackDetector_radar.scan (SEN:425-461; rotateVector / calculateAngle MSC:47-62): for every chosen history
// point the angle to the follower's right-hand vector, negated behind the follower, selects one of R sectors of the
// front half plane; a sector reports the distance to its nearest point, 0 when empty.  The chosen points are a float64
// array as soon as one of them is a float64 point the tracker seeded, float32 otherwise (numpy's np.array(list)), and
// distances / vectors keep that type; dot products are float64 (fma form of the gemv numpy dispatches to).
FTL_HD void radar_scan(const FtlConfig& c, const double2* hist, int mask, int tail, int head, int f64_end, float2 fp,
                       double dir, float* radar) {
    const int R = c.radar_sectors;
    for (int k = 0; k < R; k++) radar[k] = 0.f;
    const int n = head - tail;
    if (n <= 0) return;
    int cnt = n < c.radar_len ? n : c.radar_len, first = tail;
    if (c.radar_mode == 0) first = head - cnt;      // "new"
    else if (c.radar_mode == 2) cnt = n;            // "near": all points (the reference's sort does not change minima)
    const bool is64 = first < f64_end;
    double rdir = dir + 90;
    if (rdir >= 360) rdir -= 360;
    double wdx, wdy, wrx, wry;
    sincos_deg(dir, &wdy, &wdx);
    sincos_deg(rdir, &wry, &wrx);
    const double nwd = sqrt(fma(wdy, wdy, wdx * wdx)), nwr = sqrt(fma(wry, wry, wrx * wrx));
    const double pi = 3.141592653589793, sa = pi / R;
    for (int k = 0; k < cnt; k++) {
        const double2 h = hist[(first + k) & mask];
        double vx, vy, dist;
        if (is64) {
            vx = h.x - (double)fp.x; vy = h.y - (double)fp.y;
            const double sx = vx * vx, sy = vy * vy;
            dist = sqrt(sx + sy);
        } else {
            const float fx = (float)h.x - fp.x, fy = (float)h.y - fp.y;
            dist = (double)sqrtf(d2_f32((float)h.x, (float)h.y, fp.x, fp.y));
            vx = fx; vy = fy;
        }
        const double ad = acos(fma(vx, wdx, vy * wdy) / (dist * nwd));
        double ar = acos(fma(vx, wrx, vy * wry) / (dist * nwr));
        if (ad > pi / 2) ar = -ar;
        if (!(ar >= 0.0)) continue;                  // behind the follower, or NaN (the reference drops those too)
        int q = (int)(ar / sa);                      // candidate sector; the reference's own comparisons decide
        q = q < 1 ? 0 : q - 1;
        for (int t = 0; t < 3 && q < R; t++, q++)
            if (ar >= sa * q && ar < sa * (q + 1)) {
                const float d32 = (float)dist;
                if (radar[q] == 0.f || d32 < radar[q]) radar[q] = d32;
                break;
            }
    }
}

// LaserSensor.scan (SEN:63-136; distance_to_rect MSC:29-44).  The reference's operand types under numpy >= 2 (the
// follower position is a float32 array, everything else python floats, which are "weak"): beam angles and their
// cos / sin in float64, the beam end and the sample points in float32.  Hit boxes are tested with pygame's collidepoint
// on the truncated coordinates.
FTL_HD bool laser_in_range(int4 q, float px, float py, double limit) {   // nearest of the 4 corners + 4 edge midpoints
    const int xs[3] = {q.x, q.x + (q.z >> 1), q.x + q.z}, ys[3] = {q.y, q.y + (q.w >> 1), q.y + q.w};
    double best = 1e300;
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            if (a == 1 && b == 1) continue;
            const double dx = (double)px - xs[a], dy = (double)py - ys[b];
            const double sx = dx * dx, sy = dy * dy;
            const double d = sqrt(sx + sy);
            if (d < best) best = d;
        }
    return best <= limit;
}
FTL_HD bool laser_collidepoint(int4 q, float px, float py) {
    const int x = (int)px, y = (int)py;
    return x >= q.x && x < q.x + q.z && y >= q.y && y < q.y + q.w;
}
FTL_HD_NOINLINE void laser_scan(const FtlConfig& c, const DevState& s, const DevPool& pool, int i, float* out) {
    const float2 fp = s.pos[i];
    const double dir = s.rd[(size_t)RD_DIR * s.n + i];
    const int scen = s.gi[(size_t)GI_SCENARIO * s.n + i];
    const int4* statics = pool.static_rects + (size_t)scen * c.static_cap;
    const int n_static = pool.n_static[scen], nb = s.n_bears;
    // objects_in_range: the leader, walls and rocks, bears (SEN:77-84); at most 1 + 64 + 4 objects in two masks
    const double limit = c.laser_range + c.laser_reach_extra;
    uint64_t smask = 0;
    unsigned dmask = 0;
    for (int k = 0; k < n_static; k++)
        if (laser_in_range(statics[k], fp.x, fp.y, limit)) smask |= (uint64_t)1 << k;
    for (int k = 0; k < 1 + nb; k++)
        if (laser_in_range(s.rect[(size_t)(1 + k) * s.n + i], fp.x, fp.y, limit)) dmask |= 1u << k;
    const int border = (int)(c.laser_available_angle / 2), P = c.laser_points;
    double diff = 0;
    for (int beam = 0; beam < c.laser_beams; beam++) {
        double angle;   // SEN:93-98: -direction, then alternately +diff and -diff
        if (beam == 0) angle = -dir;
        else {
            if (beam & 1) diff += c.laser_angle_step;
            angle = angle_correction((beam & 1) ? -dir + diff : -dir - diff);
        }
        (void)border;
        double sn, cs;
        sincos_deg(angle, &sn, &cs);
        const float x2 = fp.x + (float)(c.laser_range * cs), y2 = fp.y - (float)(c.laser_range * sn);
        float hx = x2, hy = y2;
        for (int k = 0; k < P; k++) {
            const double u = (double)k / P;
            const float uf = (float)u, vf = (float)(1 - u);
            const float ax = x2 * uf, bx = fp.x * vf, ay = y2 * uf, by = fp.y * vf;
            const float cx = ax + bx, cy = ay + by;
            bool hit = false;
            for (uint64_t m = smask; m && !hit; m &= m - 1) hit = laser_collidepoint(statics[first_bit64(m)], cx, cy);
            for (int q = 0; q < 1 + nb && !hit; q++)
                if (dmask & (1u << q)) hit = laser_collidepoint(s.rect[(size_t)(1 + q) * s.n + i], cx, cy);
            if (hit) { hx = cx; hy = cy; break; }
        }
        const float dx = hx - fp.x, dy = hy - fp.y;
        if (c.laser_only_distances) {
            const float sx = dx * dx, sy = dy * dy;
            out[beam] = sqrtf(sx + sy);           // np.linalg.norm(points - position, axis=1) on float32
        } else {
            out[2 * beam] = dx;
            out[2 * beam + 1] = dy;
        }
    }
}

FTL_HD void write_optional_sensors(const FtlConfig& c, const DevState& s, const DevPool& pool, const DevOutputs& out, int i) {
    const float2 fp = s.pos[i];   // robot 0 = follower
    if (out.laser && c.laser_points > 0)
        laser_scan(c, s, pool, i, out.laser + (size_t)i * c.laser_beams * (c.laser_only_distances ? 1 : 2));
    if (out.follower_info) {
        const double fspeed = s.rd[(size_t)RD_SPEED * s.n + i], fdir = s.rd[(size_t)RD_DIR * s.n + i];
        out.follower_info[2 * (size_t)i] = (float)(fspeed / c.follower.max_speed);
        out.follower_info[2 * (size_t)i + 1] = (float)(fdir / 360);
    }
    const int mask = c.corridor_cap - 1;
    const int tail = s.gi[(size_t)GI_RING_TAIL * s.n + i], head = s.gi[(size_t)GI_RING_HEAD * s.n + i];
    const double2* hist = s.hist + (size_t)i * c.corridor_cap;
    if (out.radar && c.radar_sectors > 0)
        radar_scan(c, hist, mask, tail, head, s.gi[(size_t)GI_HIST_F64_END * s.n + i], fp,
                   s.rd[(size_t)RD_DIR * s.n + i], out.radar + (size_t)i * c.radar_sectors);
    if (!out.track_vectors || c.track_vector_len <= 0) return;
    const int P = c.track_vector_len;
    float* v = out.track_vectors + (size_t)i * P * 2;
    const int len = head - tail, cnt = len < P ? len : P;
    const int first = c.track_vector_mode == 0 ? head - cnt : tail;   // "new": the last P points, "old": the first P
    for (int k = 0; k < P; k++) {
        float2 o = make_float2(0.f, 0.f);
        if (k < cnt) {
            const double2 p = hist[(first + k) & mask];
            o = make_float2((float)(p.x - (double)fp.x), (float)(p.y - (double)fp.y));
        }
        v[2 * k] = o.x;
        v[2 * k + 1] = o.y;
    }
}

// observation part of the outputs, ENV:1789-1810 (what the kinematics kernel and the reset know)
template <int NB>
FTL_HD void write_outputs_obs(const DevCfg& cfg, const DevPool& pool, const DevOutputs& out, int i, const World<NB>& w,
                              int scenario, int cur_target_id) {
    const FtlConfig& c = cfg.c;
    if (i >= out.n) return;  // filler env of the last warp
    if (out.numerical_features) {  // ENV:1793-1802
        float* nf = out.numerical_features + (size_t)i * 10;
        nf[0] = w.leader.px; nf[1] = w.leader.py;
        nf[2] = (float)w.leader.speed; nf[3] = (float)w.leader.dir; nf[4] = (float)w.leader.rot;
        nf[5] = w.follower.px; nf[6] = w.follower.py;
        nf[7] = (float)w.follower.speed; nf[8] = (float)w.follower.dir; nf[9] = (float)w.follower.rot;
    }
    if (out.leader_target) {  // ENV:1803-1806
        int n_route = pool.n_route[scenario];
        int tid = cur_target_id < n_route ? cur_target_id : n_route - 1;
        int2 p = route_point(pool, c, scenario, tid), last = route_point(pool, c, scenario, n_route - 1);
        if (n_route > 1 && p.x == last.x && p.y == last.y) p = route_point(pool, c, scenario, n_route - 2);
        out.leader_target[2 * (size_t)i] = p.x;
        out.leader_target[2 * (size_t)i + 1] = p.y;
    }
}
// reward / done / info codes of the step (what the bookkeeping kernel knows)
FTL_HD void write_outputs_episode(const DevOutputs& out, int i, const Episode& e) {
    if (i >= out.n) return;
    if (out.reward) out.reward[i] = (float)e.last_reward;
    if (out.done) out.done[i] = (uint8_t)((e.flags & FL_DONE) != 0);
    if (out.status) {
        out.status[4 * (size_t)i + 0] = (uint8_t)((e.flags >> FL_MISSION_SHIFT) & 3);
        out.status[4 * (size_t)i + 1] = (uint8_t)((e.flags >> FL_AGENT_SHIFT) & 7);
        out.status[4 * (size_t)i + 2] = (uint8_t)((e.flags >> FL_LEADER_SHIFT) & 3);
        out.status[4 * (size_t)i + 3] = (uint8_t)((e.flags & FL_CRASH) != 0);
    }
}
template <int NB>
FTL_HD void write_outputs(const DevCfg& cfg, const DevPool& pool, const DevOutputs& out, int i, const World<NB>& w,
                          const Episode& e, bool obs_only) {
    write_outputs_obs<NB>(cfg, pool, out, i, w, e.scenario, e.cur_target_id);
    if (!obs_only) write_outputs_episode(out, i, e);
}

// ---- reset: the tail of Game.reset once the scenario exists (ENV:495-543) ---------------------------------------
template <int NB>
FTL_HD void env_reset(const DevCfg& cfg, const DevState& s, const DevPool& pool, int i, int scenario, World<NB>& w,
                      Episode& e) {
    const FtlConfig& c = cfg.c;
    float2 lp = pool.leader_pos[scenario], fp = pool.follower_pos[scenario];
    robot_init(w.leader, c.leader, lp.x, lp.y, pool.leader_dir[scenario]);
    robot_init(w.follower, c.follower, fp.x, fp.y, pool.follower_dir[scenario]);
#pragma unroll
    for (int b = 0; b < NB; b++) {  // ENV:687-718, 761-770
        float bx = (b % 2 == 0) ? lp.x + 150.f : lp.x - 150.f;
        float by = (b % 2 == 0) ? lp.y - 150.f : lp.y + 150.f;
        robot_init(w.bear[b], c.bear, bx, by, 0.0);
        w.btx[b] = (double)(lp.x - 150.f);
        w.bty[b] = (double)(lp.y - 150.f);
        w.bidx[b] = 0;
    }
    int accel_consumed = s.gi[(size_t)GI_ACCEL_CONSUMED * s.n + i];  // never restored by the reference, ENV:1170
    int episodes = s.gi[(size_t)GI_EPISODE * s.n + i];
    e.step_count = 0;
    e.cur_target_id = 1;
    e.finish_timer = -1;
    e.flags = 0;
    e.scenario = scenario;
    e.episode = episodes + 1;
    e.overflow = 0;
    e.accel_consumed = accel_consumed;
    e.acc_penalty = e.overall = e.last_reward = 0.0;
    e.speed_mult = 1.0;
    e.lead_acc = e.lead_cum = 0.0;
    // seed the trail, ENV:533-539 (float32 linspace follower -> leader)
    float2* trail = s.trail + (size_t)i * c.trail_cap;
    float* trail_d = s.trail_d + (size_t)i * c.trail_cap;
    double* trail_s = s.trail_s + (size_t)i * c.trail_cap;
    float q = sqrtf(d2_f32(fp.x, fp.y, lp.x, lp.y)) / cfg.trail_seed_denom_f32;
    int m = (int)q;
    if (m > c.trail_cap) { m = c.trail_cap; e.overflow |= 1; }
    if (m > 0) {
        float dx = lp.x - fp.x, dy = lp.y - fp.y;
        float stepx = m > 1 ? dx / (float)(m - 1) : 0.f, stepy = m > 1 ? dy / (float)(m - 1) : 0.f;
        for (int k = 0; k < m; k++) {
            float tx = (float)k * stepx, ty = (float)k * stepy;
            if (k == m - 1 && m > 1) trail_push(trail, trail_d, trail_s, k, lp.x, lp.y);
            else trail_push(trail, trail_d, trail_s, k, tx + fp.x, ty + fp.y);
        }
    }
    e.trail_len = m;
    GreenCache gc;
    green_cache_invalidate(cfg, trail_d, m, gc);
    Tracker t = {0, 0, 0, 0};
    // ring indices restart at 0; old snapshots are dropped by zeroing the push counter
    int snap_pushes = 0;
    sense_serial<NB>(cfg, s, i, w, t, &snap_pushes, &e.overflow);
    cache_store(s, i, gc, t, snap_pushes);
}

// ---- one step: Game.step without the ray sensors (ENV:908-945, 947-1141) ------------------------------------------
// The frames of a step are run as two passes that only communicate through per-frame records (FrameRec):
//   kinematics   (kin_begin + kin_frames)  robots, waypoints, bears, the integer collision tests.  Nothing in it depends
//                on the green-zone flags, the reward or the counters -- the reference's simulation goes on after a crash
//                -- so it records, per frame, the two positions and four bits the bookkeeping needs;
//   bookkeeping  (book_frames)  replays ENV:960-1139 in order: crash flags, exact green-zone flags, trail append, finish
//                timer, early stopping, reward, counters.
// On the GPU they are two kernels (k_kin, k_book): the ray kernel only needs the kinematics and the tracker, so the
// bookkeeping -- whose exact green-zone scans are the long tail of a step -- runs beside the ray kernel instead of in
// front of it; the records live in HBM, frame-major.  The host build (tests) interleaves the passes chunk by chunk.
struct FrameRec {   // rec.f[j * stride] = follower position after frame j of the step, ...
    float2* f;
    float2* l;
    unsigned char* bits;
    unsigned char* lbits;   // optional second byte of bits (the role-split kernel's leader warp writes its own); may be NULL
    size_t stride;
};

// frames this env runs in this step (ENV:939-940 via FtlStepInputs.frames_per_step) and its row of regime draws
FTL_HD int env_frames(const DevCfg& cfg, const DevState& s, int i) {
    const int cap = cfg.c.frames_per_step;
    if (!s.in_frames) return cap;
    const int f = s.in_frames[i];
    return f < 1 ? 1 : f > cap ? cap : f;
}
FTL_HD const double* env_draws(const DevCfg& cfg, const DevState& s, int i) {
    return s.in_draws ? s.in_draws + (size_t)i * cfg.c.frames_per_step : nullptr;
}

struct KinCtx {   // per-step constants of the kinematics pass
    const int4* statics;
    int n_route;
    int2 target;            // the leader's current waypoint; re-read only when the index advances
    uint64_t fmask, lmask;  // static rectangles within reach of the follower / leader during this step
    bool lfin;
};

template <int NB>
FTL_HD void kin_begin(const DevCfg& cfg, const DevPool& pool, double a0, double a1, World<NB>& w, const Episode& e,
                      KinCtx& k) {
    const FtlConfig& c = cfg.c;
    k.statics = pool.static_rects + (size_t)e.scenario * c.static_cap;
    const int n_static = pool.n_static[e.scenario];
    k.n_route = pool.n_route[e.scenario];
    command_forward(w.follower, c.follower, a0);  // ENV:927-933
    if (a1 < 0) command_turn(w.follower, c.follower, fabs(a1), -1);
    else if (a1 > 0) command_turn(w.follower, c.follower, a1, 1);
    else command_turn(w.follower, c.follower, 0, 0);
    k.target = route_point(pool, c, e.scenario, e.cur_target_id < k.n_route ? e.cur_target_id : k.n_route - 1);
    near_static_masks_grid(pool, e.scenario, k.statics, n_static, make_float2(w.follower.px, w.follower.py),
                           cfg.static_inflate[0], make_float2(w.leader.px, w.leader.py), cfg.static_inflate[1], &k.fmask,
                           &k.lmask);
    k.lfin = (e.flags & FL_LEADER_FINISHED) != 0;
}

// frames [frame0, frame0 + nf) of the step; record j of `rec` receives frame frame0 + j.  step_base = the frame counter
// at the start of the step (the regimes are keyed by the frame counter, ENV:1143-1174).
template <int NB>
FTL_HD void kin_frames(const DevCfg& cfg, const DevPool& pool, int i, World<NB>& w, Episode& e, KinCtx& k,
                       const FrameRec& rec, int step_base, int frame0, int nf, int fps, const double* draws) {
    const FtlConfig& c = cfg.c;
    const int4* statics = k.statics;
    for (int j = 0; j < nf; j++) {
        FTL_FRAME_SYNC(frame0 + j);
        int bits = 0;
        // (1) follower, ENV:957-964
#ifndef FTL_OUTLINE_ROBOTS
        robot_move(w.follower, c.follower);
#else
        w.follower = robot_move_nv(w.follower, &c.follower);
#endif
        if (!c.ignore_follower_collisions) {
            bool hit = robots_collide(w.follower, w.leader) || collide_static_masked(w.follower, statics, k.fmask) ||
                       out_of_bounds(c, w.follower);
#pragma unroll
            for (int b = 0; b < NB; b++) hit = hit || robots_collide(w.follower, w.bear[b]);
            if (hit) bits |= REC_FOLLOWER_HIT;
        }
        // too_close uses the leader before its move, ENV:973
        if (d2_f32(w.leader.px, w.leader.py, w.follower.px, w.follower.py) <= cfg.min_dist2_f32) bits |= REC_TOO_CLOSE;
        // (3) waypoint advance, ENV:978-983
        if (dist_f64_nv((double)w.leader.px, (double)w.leader.py, (double)k.target.x, (double)k.target.y) <
            c.leader_pos_epsilon) {
            e.cur_target_id += 1;
            if (e.cur_target_id >= k.n_route) k.lfin = true;   // cur_target_point keeps its last value
            else k.target = route_point(pool, c, e.scenario, e.cur_target_id);
        }
        // (4) bears, ENV:987-995
#pragma unroll
        for (int b = 0; b < NB; b++) {
            bear_target(c, b, w.bear[b], w.leader, &w.btx[b], &w.bty[b], &w.bidx[b]);
#ifndef FTL_OUTLINE_ROBOTS
            move_to_the_point(w.bear[b], c.bear, w.btx[b], w.bty[b], false, 0.0);
#else
            w.bear[b] = move_to_the_point_nv(w.bear[b], &c.bear, w.btx[b], w.bty[b], 0, 0.0);
#endif
        }
        // (5) leader, ENV:1048-1072
        if (!k.lfin) {
            e.step_count = step_base + frame0 + j;   // the regimes are keyed by the frame counter
            double speed = c.n_speed_regime > 0 ? leader_speed(cfg, e, i, draws ? draws + frame0 + j : nullptr)
                                                : c.leader.max_speed;
            double accel = c.n_accel_regime > 0 ? leader_accel(cfg, e) / fps : 0.0;
#ifndef FTL_OUTLINE_ROBOTS
            move_to_the_point(w.leader, c.leader, (double)k.target.x, (double)k.target.y, true, speed + accel);
#else
            w.leader = move_to_the_point_nv(w.leader, &c.leader, (double)k.target.x, (double)k.target.y, 1, speed + accel);
#endif
        } else {
            command_forward(w.leader, c.leader, 0);
            command_turn(w.leader, c.leader, 0, 0);
            bits |= REC_LEADER_FINISHED;
        }
        if (robots_collide(w.leader, w.follower) || collide_static_masked(w.leader, statics, k.lmask) ||
            out_of_bounds(c, w.leader))
            bits |= REC_LEADER_HIT;
        rec.f[(size_t)j * rec.stride] = make_float2(w.follower.px, w.follower.py);
        rec.l[(size_t)j * rec.stride] = make_float2(w.leader.px, w.leader.py);
        rec.bits[(size_t)j * rec.stride] = (unsigned char)bits;
        if (rec.lbits) rec.lbits[(size_t)j * rec.stride] = 0;
    }
    e.step_count = step_base;
}

// flags, trail, timers, reward of the env's nf recorded frames (ENV:960-1139); e.step_count is the running frame
// counter, fps the frames_per_step of this step (finish timer).  nf_loop >= nf: the trip count of the warp (every lane
// takes part in the collectives of the exact scans even when it has no frame left).
FTL_HD void book_frames(const DevCfg& cfg, float2* trail, float* trail_d, double* trail_s, Episode& e, GreenCache& gc,
                        const FrameRec& rec, int nf, int fps, int nf_loop) {
    const FtlConfig& c = cfg.c;
    for (int j = 0; j < nf_loop; j++) {
        if (j >= nf) {
            bool ib, ot;
            green_flags(cfg, trail, trail_d, e.trail_len, 0.f, 0.f, gc, &ib, &ot, false);
            continue;
        }
        const int bits = rec.bits[(size_t)j * rec.stride] | (rec.lbits ? rec.lbits[(size_t)j * rec.stride] : 0);
        const float2 fp = rec.f[(size_t)j * rec.stride], lp = rec.l[(size_t)j * rec.stride];
        int mission = FTL_MISSION_IN_PROGRESS, agent = FTL_AGENT_MOVING, leader_st = FTL_LEADER_MOVING;
        if (bits & REC_FOLLOWER_HIT) {   // ENV:960-964
            e.flags |= FL_CRASH | FL_DONE;
            mission = FTL_MISSION_FAIL;
            agent = FTL_AGENT_CRASH;
        }
        // (2) green zone + flags, ENV:966-973
        bool in_box, on_trace;
        green_flags(cfg, trail, trail_d, e.trail_len, fp.x, fp.y, gc, &in_box, &on_trace);
        const bool too_close = (bits & REC_TOO_CLOSE) != 0;
        if (bits & REC_LEADER_FINISHED) {
            e.flags |= FL_LEADER_FINISHED;
            leader_st = FTL_LEADER_FINISHED;
        }
        if (bits & REC_LEADER_HIT) {     // ENV:1068-1072
            e.flags |= FL_DONE;
            mission = FTL_MISSION_FAIL;
            leader_st = FTL_LEADER_CRASH;
        }
        // (6) trail append on the virtual clock, ENV:1074-1075
        if (e.step_count % c.trajectory_saving_period == 0) {
            if (e.trail_len < c.trail_cap) {
                // trail_push with the previous tail in registers
                const int k = e.trail_len;
                const float2 p = lp, q = gc.last_pt;
                float dk = 0.f;
                trail[k] = p;
                if (k > 0) {
                    dk = sqrtf(d2_f32(p.x, p.y, q.x, q.y));   // euclidean(newer, older) in float32, ENV:1838
                    gc.last_s = gc.last_s + (double)dk;
                } else {
                    gc.last_s = 0.0;
                }
                trail_d[k] = dk;
                trail_s[k] = gc.last_s;
                gc.last_pt = p;
                e.trail_len++;
                green_cache_appended(cfg, trail_s, e.trail_len, p, q, dk, gc);
            } else {
                e.overflow |= 1;
            }
        }
        // (7) finish timer, ENV:1077-1087
        if ((e.flags & FL_LEADER_FINISHED) && in_box) {
            if (e.finish_timer < 0) {
                e.finish_timer = 0;
            } else {
                e.finish_timer += 1;
                if (e.finish_timer > fps * 20) {
                    mission = FTL_MISSION_SUCCESS;
                    leader_st = FTL_LEADER_FINISHED;
                    agent = FTL_AGENT_FINISHED;
                    e.flags |= FL_DONE;
                }
            }
        }
        // (8) early stopping, ENV:1088-1107
        if (e.step_count > c.warm_start) {
            if (c.es_has_low_reward && e.acc_penalty < c.es_low_reward) {
                mission = FTL_MISSION_FAIL; leader_st = FTL_LEADER_MOVING; agent = FTL_AGENT_LOW_REWARD;
                e.flags |= FL_CRASH | FL_DONE;
            }
            if (c.es_has_max_distance_coef) {
                float d = sqrtf(d2_f32(fp.x, fp.y, lp.x, lp.y));
                if (d > cfg.es_far_f32) {
                    mission = FTL_MISSION_FAIL; leader_st = FTL_LEADER_MOVING; agent = FTL_AGENT_TOO_FAR;
                    e.flags |= FL_CRASH | FL_DONE;
                }
            }
        }
        // (9) reward + counters, ENV:1109-1139
        double r = reward_of(c, e, too_close, in_box, on_trace);
        if (r < 0) e.acc_penalty += r; else e.acc_penalty = 0;
        e.overall += r;
        e.step_count += 1;
        if (e.step_count > c.max_steps) {
            mission = FTL_MISSION_FINISHED_BY_TIME; leader_st = FTL_LEADER_MOVING; agent = FTL_AGENT_MOVING;
            e.flags |= FL_DONE;
        }
        e.last_reward = c.aggregate_reward ? e.overall : r;
        e.flags = (e.flags & (FL_DONE | FL_CRASH | FL_LEADER_FINISHED)) | (in_box ? FL_IN_BOX : 0) |
                  (on_trace ? FL_ON_TRACE : 0) | (too_close ? FL_TOO_CLOSE : 0) | (mission << FL_MISSION_SHIFT) |
                  (agent << FL_AGENT_SHIFT) | (leader_st << FL_LEADER_SHIFT);
    }
}

// The whole step for one env with a complete Episode record: the host build of the tests (and the definition of what the
// two kernels compute together).  The passes alternate over chunks of kFrameChunk frames.
template <int NB>
FTL_HD void env_step(const DevCfg& cfg, const DevState& s, const DevPool& pool, int i, double a0, double a1,
                     World<NB>& w, Episode& e) {
    const FtlConfig& c = cfg.c;
    GreenCache gc;
    Tracker t;
    int snap_pushes;
    float2* trail = s.trail + (size_t)i * c.trail_cap;
    float* trail_d = s.trail_d + (size_t)i * c.trail_cap;
    double* trail_s = s.trail_s + (size_t)i * c.trail_cap;
    cache_load(s, i, gc, t, &snap_pushes);
    green_cache_hydrate(trail, trail_s, e.trail_len, gc);
    KinCtx k;
    kin_begin<NB>(cfg, pool, a0, a1, w, e, k);
    float2 rec_f[kFrameChunk], rec_l[kFrameChunk];
    unsigned char rec_bits[kFrameChunk];
    const FrameRec rec = {rec_f, rec_l, rec_bits, nullptr, 1};
    const int step_base = e.step_count, fps = env_frames(cfg, s, i);
    const double* draws = env_draws(cfg, s, i);
    for (int f0 = 0; f0 < fps; f0 += kFrameChunk) {
        const int nf = fps - f0 < kFrameChunk ? fps - f0 : kFrameChunk;
        kin_frames<NB>(cfg, pool, i, w, e, k, rec, step_base, f0, nf, fps, draws);
        e.step_count = step_base + f0;
        book_frames(cfg, trail, trail_d, trail_s, e, gc, rec, nf, fps, nf);
    }
    sense_serial<NB>(cfg, s, i, w, t, &snap_pushes, &e.overflow);
    cache_store(s, i, gc, t, snap_pushes);
}

// ---- action decode, ENV:918-925 -------------------------------------------------------------------------------
FTL_HD void decode_action(const FtlConfig& c, const void* actions, int i, int n_real, double* a0, double* a1) {
    if (i >= n_real) { *a0 = 0.0; *a1 = 0.0; return; }  // filler envs idle
    if (c.action_mode == FTL_ACTION_DISCRETE) {
        int a = ((const int32_t*)actions)[i];
        a = a < 0 ? 0 : a > 4 ? 4 : a;
        *a0 = c.follower.max_speed;
        *a1 = c.discrete_rotation_table[a];
    } else if (c.action_mode == FTL_ACTION_CONST_SPEED) {
        *a0 = c.const_speed_action;
        *a1 = (double)((const float*)actions)[i];
    } else {
        *a0 = (double)((const float*)actions)[2 * (size_t)i];
        *a1 = (double)((const float*)actions)[2 * (size_t)i + 1];
    }
}

// the scenario an env moves to when nobody names one: same rule as the oracle's next_scenario()
FTL_HD int next_scenario(const DevCfg& cfg, int n_scenarios, int i, int episode_count) {
    int64_t g = cfg.env_id_base + i;
    return (int)((g + (int64_t)episode_count * 7919) % n_scenarios);
}

}  // namespace ftl
