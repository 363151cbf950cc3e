// ftl_state_io.cuh -- conversion between the device structure-of-arrays and the canonical
// FtlEnvState record of include/ftl.h (ftl_get_state / ftl_set_state: teacher forcing, snapshots).
#pragma once

#include <string.h>

#include "ftl_step.cuh"

namespace ftl {

FTL_HD void pack_robot(const DevState& s, int k, int i, FtlRobotState& o) {
    Robot r;
    robot_load(s, k, i, r);
    o.pos[0] = r.px; o.pos[1] = r.py;
    o.rect[0] = r.rx; o.rect[1] = r.ry; o.rect[2] = r.rw; o.rect[3] = r.rh;
    o.dir = r.dir; o.speed = r.speed; o.rot_speed = r.rot; o.des_speed = r.des_speed; o.des_rot_speed = r.des_rot;
    o.rot_dir = r.rot_dir; o.des_rot_dir = r.des_rot_dir;
}
FTL_HD void unpack_robot(const DevState& s, int k, int i, const FtlRobotState& o) {
    Robot r;
    r.px = o.pos[0]; r.py = o.pos[1];
    r.rx = o.rect[0]; r.ry = o.rect[1]; r.rw = o.rect[2]; r.rh = o.rect[3];
    r.dir = o.dir; r.speed = o.speed; r.rot = o.rot_speed; r.des_speed = o.des_speed; r.des_rot = o.des_rot_speed;
    r.rot_dir = o.rot_dir; r.des_rot_dir = o.des_rot_dir;
    robot_store(s, k, i, r);
}

FTL_HD void pack_env(const DevState& s, int i, FtlEnvState& o) {
    memset(&o, 0, sizeof o);
    pack_robot(s, 0, i, o.follower);
    pack_robot(s, 1, i, o.leader);
    for (int b = 0; b < s.n_bears; b++) {
        pack_robot(s, 2 + b, i, o.bear[b]);
        o.bear_target[b][0] = s.bear_tgt[((size_t)b * 2 + 0) * s.n + i];
        o.bear_target[b][1] = s.bear_tgt[((size_t)b * 2 + 1) * s.n + i];
        o.bear_index[b] = s.bear_idx[(size_t)b * s.n + i];
    }
    Episode e;
    episode_load(s, i, e);
    GreenCache gc;
    Tracker t;
    int pushes;
    cache_load(s, i, gc, t, &pushes);
    o.accumulated_penalty = e.acc_penalty; o.overall_reward = e.overall; o.last_reward = e.last_reward;
    o.cur_speed_multiplier = e.speed_mult; o.cur_leader_acceleration = e.lead_acc;
    o.cur_leader_cumulative_speed = e.lead_cum;
    o.accel_consumed = e.accel_consumed; o.scenario_id = e.scenario;
    o.cur_target_id = e.cur_target_id; o.leader_finished = (e.flags & FL_LEADER_FINISHED) != 0;
    o.step_count = e.step_count; o.finish_timer = e.finish_timer;
    o.done = (e.flags & FL_DONE) != 0; o.crash = (e.flags & FL_CRASH) != 0;
    o.is_in_box = (e.flags & FL_IN_BOX) != 0; o.is_on_trace = (e.flags & FL_ON_TRACE) != 0;
    o.too_close = (e.flags & FL_TOO_CLOSE) != 0;
    o.mission_status = (e.flags >> FL_MISSION_SHIFT) & 3; o.agent_status = (e.flags >> FL_AGENT_SHIFT) & 7;
    o.leader_status = (e.flags >> FL_LEADER_SHIFT) & 3;
    o.trail_len = e.trail_len; o.saving_counter = t.saving_counter;
    o.ring_tail = t.ring_tail; o.ring_head = t.ring_head; o.hist_f64_end = t.hist_f64_end;
    o.snap_pushes = pushes; o.episode_count = e.episode; o.overflow = e.overflow;
    for (int j2 = 0; j2 < FTL_MAX_HIST; j2++) {  // snap[0] oldest ... snap[MAX-1] newest
        int age = FTL_MAX_HIST - 1 - j2;
        if (age < pushes) {
            int slot = (pushes - 1 - age) % FTL_MAX_HIST;
            int2 rg = s.snap_range[(size_t)slot * s.n + i];
            o.snap[j2].valid = 1; o.snap[j2].corr_tail = rg.x; o.snap[j2].corr_head = rg.y;
            for (int k = 0; k < 1 + s.n_bears; k++) {
                int4 q = s.snap_rect[((size_t)slot * (1 + s.n_bears) + k) * s.n + i];
                o.snap[j2].dyn_rect[k][0] = q.x; o.snap[j2].dyn_rect[k][1] = q.y;
                o.snap[j2].dyn_rect[k][2] = q.z; o.snap[j2].dyn_rect[k][3] = q.w;
            }
        }
    }
}

FTL_HD void unpack_env(const DevCfg& cfg, const DevState& s, int i, const FtlEnvState& o) {
    unpack_robot(s, 0, i, o.follower);
    unpack_robot(s, 1, i, o.leader);
    for (int b = 0; b < s.n_bears; b++) {
        unpack_robot(s, 2 + b, i, o.bear[b]);
        s.bear_tgt[((size_t)b * 2 + 0) * s.n + i] = o.bear_target[b][0];
        s.bear_tgt[((size_t)b * 2 + 1) * s.n + i] = o.bear_target[b][1];
        s.bear_idx[(size_t)b * s.n + i] = o.bear_index[b];
    }
    Episode e;
    e.acc_penalty = o.accumulated_penalty; e.overall = o.overall_reward; e.last_reward = o.last_reward;
    e.speed_mult = o.cur_speed_multiplier; e.lead_acc = o.cur_leader_acceleration; e.lead_cum = o.cur_leader_cumulative_speed;
    e.accel_consumed = o.accel_consumed; e.scenario = o.scenario_id; e.cur_target_id = o.cur_target_id;
    e.step_count = o.step_count; e.finish_timer = o.finish_timer; e.trail_len = o.trail_len;
    e.episode = o.episode_count; e.overflow = o.overflow;
    e.flags = (o.done ? FL_DONE : 0) | (o.crash ? FL_CRASH : 0) | (o.leader_finished ? FL_LEADER_FINISHED : 0) |
              (o.is_in_box ? FL_IN_BOX : 0) | (o.is_on_trace ? FL_ON_TRACE : 0) | (o.too_close ? FL_TOO_CLOSE : 0) |
              (o.mission_status << FL_MISSION_SHIFT) | (o.agent_status << FL_AGENT_SHIFT) |
              (o.leader_status << FL_LEADER_SHIFT);
    episode_store(s, i, e);
    Tracker t = {o.saving_counter, o.ring_tail, o.ring_head, o.hist_f64_end};
    GreenCache gc;  // derived data: rebuild from the (already uploaded) trail
    {
        float2* trail = s.trail + (size_t)i * cfg.c.trail_cap;
        float* trail_d = s.trail_d + (size_t)i * cfg.c.trail_cap;
        double* trail_s = s.trail_s + (size_t)i * cfg.c.trail_cap;
        for (int k = 0; k < o.trail_len; k++) trail_push(trail, trail_d, trail_s, k, trail[k].x, trail[k].y);
        green_cache_invalidate(cfg, trail_d, o.trail_len, gc);
    }
    {   // derived as well: the tracker's segment lengths, from the (already uploaded) history ring
        const int cap = cfg.c.corridor_cap;
        const double2* hist = s.hist + (size_t)i * cap;
        for (int k = t.ring_tail; k + 1 < t.ring_head; k++)
            tracker_seg_store(hist, s.seg_d + (size_t)i * cap, s.seg_f + (size_t)i * cap, cap - 1, k);
    }
    int pushes = o.snap_pushes;
    cache_store(s, i, gc, t, pushes);
    for (int j2 = 0; j2 < FTL_MAX_HIST; j2++) {
        int age = FTL_MAX_HIST - 1 - j2;
        if (age < pushes && o.snap[j2].valid) {
            int slot = (pushes - 1 - age) % FTL_MAX_HIST;
            s.snap_range[(size_t)slot * s.n + i] = make_int2(o.snap[j2].corr_tail, o.snap[j2].corr_head);
            for (int k = 0; k < 1 + s.n_bears; k++)
                s.snap_rect[((size_t)slot * (1 + s.n_bears) + k) * s.n + i] =
                    make_int4(o.snap[j2].dyn_rect[k][0], o.snap[j2].dyn_rect[k][1], o.snap[j2].dyn_rect[k][2],
                              o.snap[j2].dyn_rect[k][3]);
        }
    }
}


}  // namespace ftl
