/*
 * ftl_oracle.c -- CPU restatement of the reference "follow the leader" step.  TEST INFRASTRUCTURE.
 *
 * This file is the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it; the product package never does.
 *
 * It restates, literally and scalar, one Game.step of the reference (paths relative to the
 * reference root; ENV = src/continuous_grid_arctic/follow_the_leader_continuous_env.py,
 * CLS = .../utils/classes.py, SEN = .../utils/sensors.py, MSC = .../utils/misc.py):
 *
 *   oracle function          follows
 *   -----------------------  ---------------------------------------------------------------
 *   robot_controller()       AbstractRobot._turn_processing/_speed_processing   CLS:134-163
 *   robot_move()             AbstractRobot.move                                  CLS:165-182
 *   rotated_size()           pygame 2.1.2 transform.rotate (third party, requirements.txt:5;
 *                            restated from its published algorithm, see oracle/shims/pygame)
 *   move_to_the_point()      AbstractRobot.move_to_the_point + angle_to_point    CLS:184-215, MSC:16-26
 *   collision_*()            Game._collision_check + pygame.Rect.colliderect     ENV:1176-1194
 *   green_zone_and_flags()   _trajectory_in_box + _check_agent_position          ENV:1828-1843, 1906-1937
 *   bear_target()            _choose_points_for_bear_stat / _move_bear_v4        ENV:819-837, 722-758
 *   leader_speed()/accel()   _process_leader_speed/acceleration_regime           ENV:1143-1174
 *   frame_step()             Game.frame_step                                     ENV:947-1141
 *   reward()                 Game._reward_computation                            ENV:1869-1904
 *   tracker_scan()           LeaderPositionsTracker_v2.scan                      SEN:243-327
 *   ray_sensor_scan()        LeaderCorridor_Prev_lasers_v2.scan + helpers        SEN:608-673, 883-962
 *   env_step()               Game.step + RobotWithSensors.use_sensors            ENV:908-945, CLS:255-288
 *   env_reset()              the tail of Game.reset after scenario creation      ENV:495-543
 *
 * Arithmetic types follow what the reference computes under numpy 2.3 / scipy 1.18 with python-float
 * actions (the stack in the build container): python floats are double; robot positions, the leader
 * trail and the green-zone distances are float32; the tracker history is float64 until the seeded
 * points are popped, float32 afterwards; ray predicates mix float32 and float64 exactly as numpy's
 * promotion rules make them ("precision map").  numpy's pairwise summation order and
 * the FMA forms of the BLAS kernels numpy dispatched to on the build machine are restated too.
 *
 * PINNING: the reference has no tests and no golden vectors.  This restatement is
 * pinned against traces of the UNMODIFIED reference run head-less in the build container
 * (oracle/gen_golden.py -> tests/golden/ *.npz, checked by tests/test_oracle_golden.py).  The
 * pygame.Rect / transform.rotate arithmetic inside those traces comes from oracle/shims/pygame, a
 * restatement of pygame 2.1.2 that could not be checked against real pygame offline; for that
 * boundary parity is unpinned.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/ftl.h"

#ifdef _OPENMP
#include <omp.h>
#endif

#define DEG2RAD (M_PI / 180.0)
#define RAD2DEG (180.0 / M_PI)

typedef struct FtlOracle {
    FtlConfig cfg;
    int n_envs;
    int64_t env_id_base;
    int n_threads;
    /* scenario pool (owned copies) */
    int n_scen;
    int32_t *s_static, *s_nstatic, *s_route, *s_nroute;
    float *s_lpos, *s_fpos;
    double *s_ldir, *s_fdir;
    /* state */
    FtlEnvState* env;
    float* trail;     /* [N][trail_cap][2] */
    double* hist;     /* [N][corridor_cap][2] */
    float* corridor;  /* [N][corridor_cap][4] */
    double* corridor64; /* [N][cap][4] the same points before the float32 cast of SEN:672 (LeaderCorridor_lasers_compas reads these) */
    int rays_per_env;
    double stats[FTL_STAT_COUNT];
    char err[256];
    /* optional per-step inputs (FtlStepInputs of include/ftl.h), valid during one ftl_oracle_step_ex call */
    const int32_t* in_frames;   /* [N]     frames of this step per env (ENV:939-940), NULL = cfg.frames_per_step */
    const double* in_draws;     /* [N][F]  the random() behind random.uniform of list-valued speed regimes (ENV:1155-1156) */
} FtlOracle;

/* ------------------------------------------------------------------------------------------------ */
/* small helpers                                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static double angle_correction(double a) { /* MSC:6-13 */
    if (a >= 360) return a - 360;
    if (a < 0) return 360 + a;
    return a;
}

static double angle_to_point(double cx, double cy, double tx, double ty) { /* MSC:16-26, float64 operands */
    double rx = tx - cx, ry = ty - cy, res;
    if (rx > 0)
        res = atan(ry / rx) * RAD2DEG;
    else if (rx < 0)
        res = atan(ry / rx) * RAD2DEG + 180;
    else
        res = 0;
    return angle_correction(res);
}

/* numpy add.reduce order for a contiguous float array (pairwise_sum in loops_utils.h.src) */
static float np_sum_f32(const float* a, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        float r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_f32(a, n2) + np_sum_f32(a + n2, n - n2);
    }
}
static double np_sum_f64(const double* a, int n) {
    if (n < 8) {
        double r = 0.;
        for (int i = 0; i < n; i++) r += a[i];
        return r;
    } else if (n <= 128) {
        double r[8];
        int i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return np_sum_f64(a, n2) + np_sum_f64(a + n2, n - n2);
    }
}

/* float32 euclidean distance the way scipy.spatial.distance.euclidean -> np.linalg.norm(axis=-1)
 * evaluates it for two float32 vectors: squares and sum rounded to float32, float32 sqrt. */
static float dist_f32(float ax, float ay, float bx, float by) {
    float dx = ax - bx, dy = ay - by;
    float sx = dx * dx, sy = dy * dy;
    float s = sx + sy;
    return sqrtf(s);
}
static double dist_f64(double ax, double ay, double bx, double by) {
    double dx = ax - bx, dy = ay - by;
    double sx = dx * dx, sy = dy * dy;
    return sqrt(sx + sy);
}

/* Philox4x32-10, used only for list-valued leader_speed_regime entries (the reference draws them
 * from python's global Mersenne Twister, ENV:1155-1156, which no batched implementation can share;
 * both this oracle and the CUDA path key the draw by (global env id, episode, frame)). */
static void philox4x32(uint32_t c[4], const uint32_t k_in[2]) {
    uint32_t k0 = k_in[0], k1 = k_in[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static double regime_uniform(int64_t env_global, int episode, int frame) {
    uint32_t c[4] = {(uint32_t)frame, (uint32_t)episode, 0x46544c31u, 0u};
    uint32_t k[2] = {(uint32_t)env_global, (uint32_t)((uint64_t)env_global >> 32)};
    philox4x32(c, k);
    return (double)c[0] * (1.0 / 4294967296.0);
}

/* ------------------------------------------------------------------------------------------------ */
/* robots                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
static void rect_place(FtlRobotState* r, int w, int h) { /* image.get_rect(center=position, ...), CLS:50,56 */
    r->rect[2] = w;
    r->rect[3] = h;
    r->rect[0] = (int)r->pos[0] - (w >> 1);
    r->rect[1] = (int)r->pos[1] - (h >> 1);
}

static void rotated_size(int w, int h, double angle_py, int* ow, int* oh) {
    float angle = (float)angle_py; /* PyArg "f" */
    if (fmod((double)angle, 90.0) == 0.0) {
        int q = (int)angle / 90;
        int turns = q % 4;
        if (turns < 0) turns += 4;
        if (turns & 1) { *ow = h; *oh = w; } else { *ow = w; *oh = h; }
        return;
    }
    double rad = angle * .01745329251994329;
    double s = sin(rad), c = cos(rad);
    double cx = c * w, cy = c * h, sx = s * w, sy = s * h;
    double mx = fmax(fmax(fmax(fabs(cx + sy), fabs(cx - sy)), fabs(-cx + sy)), fabs(-cx - sy));
    double my = fmax(fmax(fmax(fabs(sx + cy), fabs(sx - cy)), fabs(-sx + cy)), fabs(-sx - cy));
    *ow = (int)mx;
    *oh = (int)my;
}

static void command_turn(FtlRobotState* r, const FtlRobotConfig* c, double des, int dir) { /* CLS:109-117 */
    r->des_rot_speed = (c->max_rotation_speed < des) ? c->max_rotation_speed : des;
    r->des_rot_dir = dir;
}
static void command_forward(FtlRobotState* r, const FtlRobotConfig* c, double des) { /* CLS:119-127 */
    if (des > c->max_speed) des = c->max_speed;
    if (des < c->min_speed) des = c->min_speed;
    r->des_speed = des;
}

static void robot_controller(FtlRobotState* r, const FtlRobotConfig* c) { /* CLS:129-163 */
    double change;
    if (r->rot_dir == 0) r->rot_dir = r->des_rot_dir;
    if (r->rot_dir == r->des_rot_dir) {
        double needed = fabs(r->rot_speed - r->des_rot_speed);
        change = (c->max_rotation_speed_change < needed) ? c->max_rotation_speed_change : needed;
        if (r->des_rot_speed < r->rot_speed) change = -1 * change;
    } else {
        double needed = fabs(r->des_rot_speed + r->rot_speed);
        change = -((c->max_rotation_speed_change < needed) ? c->max_rotation_speed_change : needed);
    }
    double nrs = r->rot_speed + change;
    if (nrs < 0) r->rot_dir = -1 * r->rot_dir;
    r->rot_speed = fabs(nrs);

    double needed = fabs(r->speed - r->des_speed);
    double sc = (needed < c->max_speed_change) ? needed : c->max_speed_change;
    if (r->speed > r->des_speed) sc = -1 * sc;
    r->speed = r->speed + sc;
}

static void robot_move(FtlRobotState* r, const FtlRobotConfig* c) { /* CLS:165-182 */
    robot_controller(r, c);
    if (r->rot_speed != 0) {
        r->dir = angle_correction(r->dir + r->rot_dir * r->rot_speed);
        int cx = r->rect[0] + (r->rect[2] >> 1), cy = r->rect[1] + (r->rect[3] >> 1);
        int nw, nh;
        rotated_size(c->width, c->height, -r->dir, &nw, &nh);
        r->rect[2] = nw;
        r->rect[3] = nh;
        r->rect[0] = cx - (nw >> 1);
        r->rect[1] = cy - (nh >> 1);
    }
    double th = r->dir * DEG2RAD;
    float mx = (float)(cos(th) * r->speed), my = (float)(sin(th) * r->speed);
    float px = r->pos[0] + mx, py = r->pos[1] + my;
    r->pos[0] = px;
    r->pos[1] = py;
    double dx = (double)r->pos[0] - (double)(r->rect[0] + (r->rect[2] >> 1));
    double dy = (double)r->pos[1] - (double)(r->rect[1] + (r->rect[3] >> 1));
    if (dx != 0 || dy != 0) {
        r->rect[0] += (int)dx;
        r->rect[1] += (int)dy;
    }
}

/* speed < 0 means "speed is None" (bears): desired speed = distance to the point, CLS:187-190 */
static void move_to_the_point(FtlRobotState* r, const FtlRobotConfig* c, double tx, double ty, int has_speed,
                              double speed) {
    double new_speed = has_speed ? speed : dist_f64((double)r->pos[0], (double)r->pos[1], tx, ty);
    int desirable_angle = (int)angle_to_point((double)r->pos[0], (double)r->pos[1], tx, ty);
    int cur_direction = (int)r->dir;
    int delta_turn, nrd;
    if (desirable_angle - cur_direction > 0) {
        if (desirable_angle - cur_direction > 180) {
            delta_turn = cur_direction + (360 - desirable_angle);
            nrd = -1;
        } else {
            delta_turn = desirable_angle - cur_direction;
            nrd = 1;
        }
    } else {
        if (cur_direction - desirable_angle > 180) {
            nrd = 1;
            delta_turn = (360 - cur_direction) + desirable_angle;
        } else {
            nrd = -1;
            delta_turn = cur_direction - desirable_angle;
        }
    }
    command_turn(r, c, (double)delta_turn, nrd);
    command_forward(r, c, new_speed);
    robot_move(r, c);
}

/* ------------------------------------------------------------------------------------------------ */
/* collisions (integer, pygame.Rect.colliderect)                                                     */
/* ------------------------------------------------------------------------------------------------ */
static int rects_collide(const int32_t* a, const int32_t* b) {
    if (a[2] == 0 || a[3] == 0 || b[2] == 0 || b[3] == 0) return 0;
    return a[0] < b[0] + b[2] && a[1] < b[1] + b[3] && a[0] + a[2] > b[0] && a[1] + a[3] > b[1];
}
static int out_of_bounds(const FtlConfig* cfg, const FtlRobotState* r) { /* ENV:1182-1183 */
    return r->pos[0] > (float)cfg->game_width || r->pos[1] > (float)cfg->game_height || r->pos[0] < 0 || r->pos[1] < 0;
}

static int collision_follower(const FtlOracle* o, const FtlEnvState* e) { /* ENV:1187-1194 */
    const FtlConfig* cfg = &o->cfg;
    const int32_t* sr = o->s_static + (size_t)e->scenario_id * cfg->static_cap * 4;
    int ns = o->s_nstatic[e->scenario_id];
    if (rects_collide(e->follower.rect, e->leader.rect)) return 1;
    for (int i = 0; i < ns; i++)
        if (rects_collide(e->follower.rect, sr + 4 * i)) return 1;
    for (int b = 0; b < cfg->n_bears; b++)
        if (rects_collide(e->follower.rect, e->bear[b].rect)) return 1;
    return out_of_bounds(cfg, &e->follower);
}
static int collision_leader(const FtlOracle* o, const FtlEnvState* e) { /* ENV:1180-1186 */
    const FtlConfig* cfg = &o->cfg;
    const int32_t* sr = o->s_static + (size_t)e->scenario_id * cfg->static_cap * 4;
    int ns = o->s_nstatic[e->scenario_id];
    if (rects_collide(e->leader.rect, e->follower.rect)) return 1;
    for (int i = 0; i < ns; i++)
        if (rects_collide(e->leader.rect, sr + 4 * i)) return 1;
    return out_of_bounds(cfg, &e->leader);
}

/* ------------------------------------------------------------------------------------------------ */
/* green zone and position flags                                                                     */
/* ------------------------------------------------------------------------------------------------ */
static void green_zone_and_flags(const FtlOracle* o, FtlEnvState* e, const float* trail) {
    const FtlConfig* cfg = &o->cfg;
    int n = e->trail_len;
    /* _trajectory_in_box, ENV:1828-1843: walk back from the newest point; float32 accumulation */
    int n_green = 0;
    int first_green = n - 2; /* green points are trail[first_green], trail[first_green-1], ... */
    float acc = 0.f;
    for (int i = n - 2; i >= 0; i--) {
        float t = acc + dist_f32(trail[2 * (i + 1)], trail[2 * (i + 1) + 1], trail[2 * i], trail[2 * i + 1]);
        acc = t;
        if (acc <= (float)cfg->max_distance)
            n_green++;
        else
            break;
    }
    float fx = e->follower.pos[0], fy = e->follower.pos[1];
    e->is_in_box = 0;
    e->is_on_trace = 0;
    /* _check_agent_position, ENV:1906-1937 */
    if (n_green > 2) {
        float best = 0.f;
        for (int k = 0; k < n_green; k++) { /* list order: newest first; argmin keeps the first minimum */
            int i = first_green - k;
            float dx = trail[2 * i] - fx, dy = trail[2 * i + 1] - fy;
            float sx = dx * dx, sy = dy * dy;
            float d2 = sx + sy;
            if (k == 0 || d2 < best) best = d2;
        }
        float g = sqrtf(best);
        if (g <= (float)cfg->leader_pos_epsilon) {
            e->is_on_trace = 1;
            e->is_in_box = 1;
        } else if (g <= (float)cfg->max_dev) {
            e->is_in_box = 1;
        } else {
            float bestall = 0.f;
            for (int i = 0; i < n; i++) {
                float dx = trail[2 * i] - fx, dy = trail[2 * i + 1] - fy;
                float sx = dx * dx, sy = dy * dy;
                float d2 = sx + sy;
                if (i == 0 || d2 < bestall) bestall = d2;
            }
            if (sqrtf(bestall) <= (float)cfg->leader_pos_epsilon) e->is_on_trace = 1;
        }
    }
    float dl = dist_f32(e->leader.pos[0], e->leader.pos[1], fx, fy);
    e->too_close = dl <= (float)cfg->min_distance;
}

/* ------------------------------------------------------------------------------------------------ */
/* bears                                                                                             */
/* ------------------------------------------------------------------------------------------------ */
static void rot_point(const FtlRobotState* leader, double radius, double angle_deg, double* ox, double* oy) {
    /* leader.position + rotateVector([radius, 0], angle), MSC:47-53 */
    double th = angle_deg * DEG2RAD;
    *ox = (double)leader->pos[0] + cos(th) * radius;
    *oy = (double)leader->pos[1] + sin(th) * radius;
}

static void bear_target(const FtlOracle* o, FtlEnvState* e, int idx) {
    const FtlConfig* cfg = &o->cfg;
    FtlRobotState* b = &e->bear[idx];
    double d = dist_f64((double)b->pos[0], (double)b->pos[1], e->bear_target[idx][0], e->bear_target[idx][1]);
    if (cfg->move_bear_v4 && (idx % 2)) { /* _move_bear_v4, ENV:722-758 */
        if (d < cfg->leader_pos_epsilon) e->bear_index[idx] += 1;
        if (e->bear_index[idx] > 3) e->bear_index[idx] = 0;
        double p[4][2]; /* p1..p4 */
        rot_point(&e->leader, 150, e->leader.dir + 140, &p[0][0], &p[0][1]);
        rot_point(&e->leader, 150, e->leader.dir - 140, &p[1][0], &p[1][1]);
        rot_point(&e->leader, 250, e->leader.dir - 160, &p[2][0], &p[2][1]);
        rot_point(&e->leader, 250, e->leader.dir + 160, &p[3][0], &p[3][1]);
        static const int order[4][4] = {{0, 1, 3, 2}, {3, 2, 0, 1}, {1, 3, 2, 0}, {2, 0, 1, 3}};
        int k = order[idx & 3][e->bear_index[idx]];
        e->bear_target[idx][0] = p[k][0];
        e->bear_target[idx][1] = p[k][1];
    } else { /* _choose_points_for_bear_stat, ENV:819-837 */
        if (d < cfg->leader_pos_epsilon) {
            e->bear_index[idx] += 1;
            if (e->bear_index[idx] > 1) e->bear_index[idx] = 0;
        }
        double radius = 100 * (idx + 1);
        double ang = e->bear_index[idx] == 0 ? e->leader.dir - 130 : e->leader.dir + 130;
        rot_point(&e->leader, radius, ang, &e->bear_target[idx][0], &e->bear_target[idx][1]);
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* leader speed / acceleration regimes                                                               */
/* ------------------------------------------------------------------------------------------------ */
static double leader_speed(const FtlOracle* o, FtlEnvState* e, int env_index, const double* draw) { /* ENV:1143-1157 */
    const FtlConfig* cfg = &o->cfg;
    int sel = -1;
    for (int k = 0; k < cfg->n_speed_regime; k++)
        if (cfg->speed_regime_key[k] <= e->step_count) sel = k;
    if (sel >= 0) {
        if (cfg->speed_regime_is_range[sel]) {
            /* random.uniform(a, b) = a + (b - a) * random(): the caller's recorded draw, or Philox */
            double u = draw ? *draw : regime_uniform(o->env_id_base + env_index, e->episode_count, e->step_count);
            e->cur_speed_multiplier = cfg->speed_regime_lo[sel] + (cfg->speed_regime_hi[sel] - cfg->speed_regime_lo[sel]) * u;
        } else {
            e->cur_speed_multiplier = cfg->speed_regime_lo[sel];
        }
    }
    return cfg->leader.max_speed * e->cur_speed_multiplier;
}
static double leader_accel(const FtlOracle* o, FtlEnvState* e) { /* ENV:1159-1174 */
    const FtlConfig* cfg = &o->cfg;
    for (int k = 0; k < cfg->n_accel_regime; k++) {
        if (e->accel_consumed & (1 << k)) continue;
        if (cfg->accel_regime_key[k] <= e->step_count) {
            e->cur_leader_acceleration = cfg->accel_regime_val[k];
            e->cur_leader_cumulative_speed = e->cur_leader_acceleration;
            e->accel_consumed |= (1 << k);
        }
    }
    e->cur_leader_cumulative_speed += e->cur_leader_acceleration;
    return e->cur_leader_cumulative_speed * cfg->leader.max_speed;
}

/* ------------------------------------------------------------------------------------------------ */
/* reward                                                                                            */
/* ------------------------------------------------------------------------------------------------ */
static double reward(const FtlConfig* cfg, const FtlEnvState* e) { /* ENV:1869-1904; stop_signal is never set */
    double r = 0;
    r += cfg->leader_movement_reward;
    if (e->too_close) {
        r += cfg->too_close_penalty;
    } else {
        if (e->is_in_box && e->is_on_trace)
            r += cfg->reward_in_box;
        else if (e->is_in_box)
            r += cfg->reward_in_dev;
        else if (e->is_on_trace)
            r += cfg->reward_on_track;
        else if (e->step_count > cfg->warm_start)
            r += cfg->not_on_track_penalty;
    }
    if (e->crash) r += cfg->crash_penalty;
    return r;
}

/* ------------------------------------------------------------------------------------------------ */
/* one frame                                                                                         */
/* ------------------------------------------------------------------------------------------------ */
static void frame_step(FtlOracle* o, int env_index, int fps, const double* draw) { /* ENV:947-1141; fps = self.frames_per_step */
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[env_index];
    float* trail = o->trail + (size_t)env_index * cfg->trail_cap * 2;
    const int32_t* route = o->s_route + (size_t)e->scenario_id * cfg->route_cap * 2;
    int n_route = o->s_nroute[e->scenario_id];

    e->is_in_box = 0;
    e->is_on_trace = 0;
    int mission = FTL_MISSION_IN_PROGRESS, agent = FTL_AGENT_MOVING, leader_st = FTL_LEADER_MOVING;

    robot_move(&e->follower, &cfg->follower);
    if (!cfg->ignore_follower_collisions && collision_follower(o, e)) {
        e->crash = 1;
        e->done = 1;
        mission = FTL_MISSION_FAIL;
        agent = FTL_AGENT_CRASH;
    }
    green_zone_and_flags(o, e, trail);

    /* waypoint advance, ENV:978-983 */
    {
        int tid = e->cur_target_id < n_route ? e->cur_target_id : n_route - 1;
        /* cur_target_point keeps its last value once the route is exhausted */
        double tx = route[2 * tid], ty = route[2 * tid + 1];
        if (dist_f64((double)e->leader.pos[0], (double)e->leader.pos[1], tx, ty) < cfg->leader_pos_epsilon) {
            e->cur_target_id += 1;
            if (e->cur_target_id >= n_route) e->leader_finished = 1;
        }
    }
    /* bears, ENV:987-995 */
    for (int b = 0; b < cfg->n_bears; b++) {
        bear_target(o, e, b);
        move_to_the_point(&e->bear[b], &cfg->bear, e->bear_target[b][0], e->bear_target[b][1], 0, 0.0);
    }
    /* leader, ENV:1048-1065 */
    if (!e->leader_finished) {
        double speed = cfg->n_speed_regime > 0 ? leader_speed(o, e, env_index, draw) : cfg->leader.max_speed;
        double accel = cfg->n_accel_regime > 0 ? leader_accel(o, e) / fps : 0;
        int tid = e->cur_target_id;
        move_to_the_point(&e->leader, &cfg->leader, (double)route[2 * tid], (double)route[2 * tid + 1], 1, speed + accel);
    } else {
        command_forward(&e->leader, &cfg->leader, 0);
        command_turn(&e->leader, &cfg->leader, 0, 0);
        leader_st = FTL_LEADER_FINISHED;
    }
    if (collision_leader(o, e)) { /* ENV:1068-1072 */
        e->done = 1;
        mission = FTL_MISSION_FAIL;
        leader_st = FTL_LEADER_CRASH;
    }
    /* trail append on the virtual clock, ENV:1074-1075 (ticks == frames since reset) */
    if (e->step_count % cfg->trajectory_saving_period == 0) {
        if (e->trail_len < cfg->trail_cap) {
            trail[2 * e->trail_len] = e->leader.pos[0];
            trail[2 * e->trail_len + 1] = e->leader.pos[1];
            e->trail_len++;
        } else {
            e->overflow |= 1;
        }
    }
    /* finish timer, ENV:1077-1087 */
    if (e->leader_finished && e->is_in_box) {
        if (e->finish_timer < 0) {
            e->finish_timer = 0;
        } else {
            e->finish_timer += 1;
            if (e->finish_timer > fps * 20) {
                mission = FTL_MISSION_SUCCESS;
                leader_st = FTL_LEADER_FINISHED;
                agent = FTL_AGENT_FINISHED;
                e->done = 1;
            }
        }
    }
    /* early stopping, ENV:1088-1107 */
    if (e->step_count > cfg->warm_start) {
        if (cfg->es_has_low_reward && e->accumulated_penalty < cfg->es_low_reward) {
            mission = FTL_MISSION_FAIL;
            leader_st = FTL_LEADER_MOVING;
            agent = FTL_AGENT_LOW_REWARD;
            e->crash = 1;
            e->done = 1;
        }
        if (cfg->es_has_max_distance_coef) {
            float d = dist_f32(e->follower.pos[0], e->follower.pos[1], e->leader.pos[0], e->leader.pos[1]);
            /* np.linalg.norm(float32) > python float: compared in float32 */
            if (d > (float)(cfg->max_distance * cfg->es_max_distance_coef)) {
                mission = FTL_MISSION_FAIL;
                leader_st = FTL_LEADER_MOVING;
                agent = FTL_AGENT_TOO_FAR;
                e->crash = 1;
                e->done = 1;
            }
        }
    }
    double r = reward(cfg, e);
    if (r < 0)
        e->accumulated_penalty += r;
    else
        e->accumulated_penalty = 0;
    e->overall_reward += r;

    e->step_count += 1;
    if (e->step_count > cfg->max_steps) { /* ENV:1129-1134 */
        mission = FTL_MISSION_FINISHED_BY_TIME;
        leader_st = FTL_LEADER_MOVING;
        agent = FTL_AGENT_MOVING;
        e->done = 1;
    }
    e->last_reward = cfg->aggregate_reward ? e->overall_reward : r;
    e->mission_status = mission;
    e->agent_status = agent;
    e->leader_status = leader_st;
}

/* ------------------------------------------------------------------------------------------------ */
/* LeaderPositionsTracker_v2.scan                                                                    */
/* ------------------------------------------------------------------------------------------------ */
#define RING(i) ((i) & (cap - 1))

static double path_length(const FtlEnvState* e, const double* hist, int cap) {
    /* np.sum(np.linalg.norm(h[:-1] - h[1:], axis=1)) in float64 while any seeded (float64) point is
     * alive, float32 afterwards (np.array() of an all-float32 deque is float32), SEN:283-285 */
    int n = e->ring_head - e->ring_tail;
    if (n < 2) return 0.0;
    if (e->ring_tail < e->hist_f64_end) {
        double d[512];
        for (int k = 0; k < n - 1 && k < 512; k++) {
            const double* a = hist + 2 * RING(e->ring_tail + k);
            const double* b = hist + 2 * RING(e->ring_tail + k + 1);
            d[k] = dist_f64(a[0], a[1], b[0], b[1]);
        }
        return np_sum_f64(d, n - 1);
    } else {
        float d[512];
        for (int k = 0; k < n - 1 && k < 512; k++) {
            const double* a = hist + 2 * RING(e->ring_tail + k);
            const double* b = hist + 2 * RING(e->ring_tail + k + 1);
            d[k] = dist_f32((float)a[0], (float)a[1], (float)b[0], (float)b[1]);
        }
        return (double)np_sum_f32(d, n - 1);
    }
}

/* border points for the segment hist[ia] -> hist[ib], anchored at hist[ianchor]; SEN:302-317 */
static void corridor_entry(const FtlOracle* o, const FtlEnvState* e, const double* hist, int ia, int ib, int ianchor,
                           float out[4], double* out64) {
    const FtlConfig* cfg = &o->cfg;
    int cap = cfg->corridor_cap;
    const double* pa = hist + 2 * RING(ia);
    const double* pb = hist + 2 * RING(ib);
    const double* pc = hist + 2 * RING(ianchor);
    double vx, vy;
    int f64 = (ia < e->hist_f64_end) || (ib < e->hist_f64_end);
    if (f64) {
        vx = pb[0] - pa[0];
        vy = pb[1] - pa[1];
        double nrm = sqrt(fma(vy, vy, vx * vx)); /* ddot(n=2) as dispatched on the build machine */
        double s = cfg->corridor_width / nrm;
        vx *= s;
        vy *= s;
    } else {
        float fx = (float)pb[0] - (float)pa[0], fy = (float)pb[1] - (float)pa[1];
        float sx = fx * fx, sy = fy * fy;
        float ss = sx + sy;
        float nrm = sqrtf(ss);
        float s = (float)cfg->corridor_width / nrm;
        float gx = fx * s, gy = fy * s;
        vx = gx;
        vy = gy;
    }
    /* rotateVector(v, +-90): np.dot(rot, v) = (fma(R00, v0, R01*v1), fma(R10, v0, R11*v1)) */
    const double c90 = 6.123233995736766e-17;
    double rx = fma(c90, vx, -1.0 * vy), ry = fma(1.0, vx, c90 * vy);
    double lx = fma(c90, vx, 1.0 * vy), ly = fma(-1.0, vx, c90 * vy);
    rx += pc[0]; ry += pc[1];
    lx += pc[0]; ly += pc[1];
    out[0] = (float)rx; out[1] = (float)ry; out[2] = (float)lx; out[3] = (float)ly;
    if (out64) { out64[0] = rx; out64[1] = ry; out64[2] = lx; out64[3] = ly; }
}

static void tracker_scan(FtlOracle* o, int env_index) { /* SEN:243-327 */
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[env_index];
    int cap = cfg->corridor_cap;
    double* hist = o->hist + (size_t)env_index * cap * 2;
    float* corr = o->corridor + (size_t)env_index * cap * 4;
    double* corr64 = o->corridor64 + (size_t)env_index * cap * 4;

    if (e->saving_counter % cfg->saving_period == 0) {
        int n = e->ring_head - e->ring_tail;
        if (n > 0) {
            const double* last = hist + 2 * RING(e->ring_head - 1);
            if (last[0] == (double)e->leader.pos[0] && last[1] == (double)e->leader.pos[1]) return; /* SEN:247-251 */
        }
        if (n == 0 && e->saving_counter == 0) {
            double sx, sy;
            if (cfg->start_corridor_behind_follower) { /* SEN:257-272 */
                double th = angle_correction(e->follower.dir + 180) * DEG2RAD;
                sx = 50 * cos(th) + (double)e->follower.pos[0];
                sy = 50 * sin(th) + (double)e->follower.pos[1];
            } else {
                sx = e->follower.pos[0];
                sy = e->follower.pos[1];
            }
            double lx = e->leader.pos[0], ly = e->leader.pos[1];
            int m;
            if (cfg->start_corridor_behind_follower) {
                m = (int)(dist_f64(sx, sy, lx, ly) / (cfg->saving_period * 5 * cfg->leader.max_speed));
                if (m > cap) { m = cap; e->overflow |= 2; }
                /* np.linspace in float64: i*step + start, last = stop */
                double stepx = m > 1 ? (lx - sx) / (m - 1) : 0, stepy = m > 1 ? (ly - sy) / (m - 1) : 0;
                for (int i = 0; i < m; i++) {
                    double tx = (double)i * stepx, ty = (double)i * stepy;
                    hist[2 * RING(e->ring_head + i)] = tx + sx;
                    hist[2 * RING(e->ring_head + i) + 1] = ty + sy;
                }
            } else { /* SEN:273-281: both ends float32 -> float32 distance, float32 linspace */
                float fsx = e->follower.pos[0], fsy = e->follower.pos[1], flx = e->leader.pos[0], fly = e->leader.pos[1];
                float q = dist_f32(fsx, fsy, flx, fly) / (float)(cfg->saving_period * 5 * cfg->leader.max_speed);
                m = (int)q;
                if (m > cap) { m = cap; e->overflow |= 2; }
                float ddx = flx - fsx, ddy = fly - fsy;
                float stepx = m > 1 ? ddx / (float)(m - 1) : 0.f, stepy = m > 1 ? ddy / (float)(m - 1) : 0.f;
                for (int i = 0; i < m; i++) {
                    float tx = (float)i * stepx, ty = (float)i * stepy;
                    float vx = tx + fsx, vy = ty + fsy;
                    hist[2 * RING(e->ring_head + i)] = vx;
                    hist[2 * RING(e->ring_head + i) + 1] = vy;
                }
            }
            if (m > 1) {
                hist[2 * RING(e->ring_head + m - 1)] = lx;
                hist[2 * RING(e->ring_head + m - 1) + 1] = ly;
            }
            e->ring_head += m;
            /* float64 typed unless start_corridor_behind_follower is off (then float32 linspace) */
            e->hist_f64_end = cfg->start_corridor_behind_follower ? e->ring_head : e->ring_tail;
        } else {
            if (e->ring_head - e->ring_tail >= cap) { /* ring full: drop the oldest (overflow) */
                e->ring_tail++;
                e->overflow |= 2;
            }
            hist[2 * RING(e->ring_head)] = e->leader.pos[0];
            hist[2 * RING(e->ring_head) + 1] = e->leader.pos[1];
            e->ring_head++;
        }
        int corr_len = e->saving_counter == 0 ? 0 : (e->ring_head - 1 - e->ring_tail);
        /* trim, SEN:286-292 (hist and corridor are popped together) */
        while (path_length(e, hist, cap) > ((e->ring_tail < e->hist_f64_end) ? cfg->corridor_length
                                                                               : (double)(float)cfg->corridor_length)) {
            e->ring_tail++;
            corr_len--;
        }
        n = e->ring_head - e->ring_tail;
        if (n > 1) {
            if (e->saving_counter == 0) { /* SEN:300-308: m-1 entries, anchors hist[0..m-2] */
                for (int i = n - 1; i > 0; i--) {
                    int anchor = e->ring_tail + (n - i - 1);
                    corridor_entry(o, e, hist, e->ring_tail + i - 1, e->ring_tail + i, anchor, corr + 4 * RING(anchor),
                                   corr64 + 4 * RING(anchor));
                }
                /* plus the duplicate anchored at hist[-2] from the last segment, stored at slot head-1 */
                corridor_entry(o, e, hist, e->ring_head - 2, e->ring_head - 1, e->ring_head - 2,
                               corr + 4 * RING(e->ring_head - 1), corr64 + 4 * RING(e->ring_head - 1));
            } else {
                corridor_entry(o, e, hist, e->ring_head - 2, e->ring_head - 1, e->ring_head - 2,
                               corr + 4 * RING(e->ring_head - 1), corr64 + 4 * RING(e->ring_head - 1));
            }
        }
        (void)corr_len;
    }
    e->saving_counter += 1;
}

/* ------------------------------------------------------------------------------------------------ */
/* LeaderCorridor_Prev_lasers_v2.scan                                                                */
/* ------------------------------------------------------------------------------------------------ */
#define MAX_EDGES 2048

static void push_rect_edges(float* ed, int* ne, const int32_t* r) { /* SEN:668-671 */
    float l = (float)r[0], t = (float)r[1], rr = (float)(r[0] + r[2]), b = (float)(r[1] + r[3]);
    float q[4][4] = {{l, b, rr, b}, {rr, t, rr, b}, {rr, t, l, t}, {l, b, l, t}};
    for (int k = 0; k < 4 && *ne < MAX_EDGES; k++, (*ne)++) memcpy(ed + 4 * (*ne), q[k], sizeof(float) * 4);
}

/* edge list of one history entry, in the order collect_obstacle_edges builds it, SEN:642-673 */
static int collect_edges(const FtlOracle* o, const FtlEnvState* e, const FtlRaySensorConfig* sc, const FtlSnapshot* sn,
                         const float* corr, float* ed) {
    const FtlConfig* cfg = &o->cfg;
    int cap = cfg->corridor_cap;
    int ne = 0;
    if (sc->react_to_safe_corridor) {
        for (int i = sn->corr_tail; i < sn->corr_head - 1 && ne + 2 <= MAX_EDGES; i++) {
            const float* a = corr + 4 * RING(i);
            const float* b = corr + 4 * RING(i + 1);
            float e0[4] = {a[0], a[1], b[0], b[1]}, e1[4] = {a[2], a[3], b[2], b[3]};
            memcpy(ed + 4 * ne++, e0, sizeof e0);
            memcpy(ed + 4 * ne++, e1, sizeof e1);
        }
    }
    if (sc->react_to_green_zone) {
        const float* a = corr + 4 * RING(sn->corr_tail);
        const float* b = corr + 4 * RING(sn->corr_head - 1);
        float e0[4] = {a[0], a[1], a[2], a[3]}, e1[4] = {b[0], b[1], b[2], b[3]};
        memcpy(ed + 4 * ne++, e0, sizeof e0);
        memcpy(ed + 4 * ne++, e1, sizeof e1);
    }
    int mode = sc->react_to_obstacles;
    if (mode == FTL_REACT_ALL || mode == FTL_REACT_STATIC) { /* game_object_list minus the follower */
        push_rect_edges(ed, &ne, sn->dyn_rect[0]);
        const int32_t* sr = o->s_static + (size_t)e->scenario_id * cfg->static_cap * 4;
        for (int i = 0; i < o->s_nstatic[e->scenario_id]; i++) push_rect_edges(ed, &ne, sr + 4 * i);
    }
    if (mode == FTL_REACT_ALL || mode == FTL_REACT_DYNAMIC)
        for (int b = 0; b < cfg->n_bears; b++) push_rect_edges(ed, &ne, sn->dyn_rect[1 + b]);
    return ne;
}

/* ccw with numpy's operand types: edge points and the follower position are float32, the ray end is
 * float64 (SEN:608-609 on the arrays SEN:903-906 builds) */
static int ccw_ffd(float ax, float ay, float cx, float cy, double dx, double dy) { /* ccw(A, C, D) */
    float ux = cx - ax, uy = cy - ay;
    double lhs = (dy - (double)ay) * (double)ux, rhs = (double)uy * (dx - (double)ax);
    return lhs > rhs;
}
static int ccw_fff(float ax, float ay, float bx, float by, float cx, float cy) { /* ccw(A, B, C) all float32 */
    float a = cy - ay, b = bx - ax, c = by - ay, d = cx - ax;
    float lhs = a * b, rhs = c * d;
    return lhs > rhs;
}
static int ccw_ffD(float ax, float ay, float bx, float by, double dx, double dy) { /* ccw(A, B, D) */
    float ux = bx - ax, uy = by - ay;
    double lhs = (dy - (double)ay) * (double)ux, rhs = (double)uy * (dx - (double)ax);
    return lhs > rhs;
}

/* LeaderCorridor_lasers_compas.scan, SEN:1138-1240: the corridor in float64 (corridor_lines is built from the tracker's
 * float64 points without the float32 cast of SEN:672), lines ordered front wall, back wall, left walls, right walls;
 * np.argmin takes the first of equal minima. */
static void compas_test(double ax, double ay, double bx, double by, double cx, double cy, double ex, double ey, int orient,
                        double* best_sel, double* best_x, double* best_y, int* best_orient) {
    /* ccw(A, B, C) = (C.y - A.y) * (B.x - A.x) > (B.y - A.y) * (C.x - A.x), SEN:608-609 */
    int acd = (ey - ay) * (cx - ax) > (cy - ay) * (ex - ax);
    int bcd = (ey - by) * (cx - bx) > (cy - by) * (ex - bx);
    if (acd == bcd) return;
    int abc = (cy - ay) * (bx - ax) > (by - ay) * (cx - ax);
    int abd = (ey - ay) * (bx - ax) > (by - ay) * (ex - ax);
    if (abc == abd) return;
    double dax = bx - ax, day = by - ay, dbx = ex - cx, dby = ey - cy, dpx = ax - cx, dpy = ay - cy;
    double dapx = -day, dapy = dax;
    double denom = fma(dapx, dbx, dapy * dby); /* np.dot((k,2),(2,1)) */
    double m0 = dapx * dpx, m1 = dapy * dpy;
    double t = (m0 + m1) / denom;
    double xx = t * dbx + cx, xy = t * dby + cy;
    double ddx = xx - cx, ddy = xy - cy;
    double sx = ddx * ddx, sy = ddy * ddy;
    double sel = sqrt(sx + sy); /* np.linalg.norm(x - pos, axis=1) */
    if (*best_orient < 0 || sel < *best_sel) { *best_sel = sel; *best_x = xx; *best_y = xy; *best_orient = orient; }
}

static void compas_sensor_scan(const FtlOracle* o, int env_index, int sensor, float* out) {
    const FtlConfig* cfg = &o->cfg;
    const FtlEnvState* e = &o->env[env_index];
    const FtlRaySensorConfig* sc = &cfg->ray[sensor];
    int cap = cfg->corridor_cap;
    const double* corr = o->corridor64 + (size_t)env_index * cap * 4;
    int R = sc->lasers_count, H = sc->max_prev_obs;
    double cx = e->follower.pos[0], cy = e->follower.pos[1];
    double L = sc->laser_length, period = 360.0 / R;
    for (int j = 0; j < H; j++) {
        const FtlSnapshot* sn = &e->snap[FTL_MAX_HIST - H + j];
        float* row = out + (size_t)j * 5 * R;
        for (int k = 0; k < 5 * R; k++) row[k] = 0.f;
        for (int i = 0; i < R; i++) {
            double ang = (e->follower.dir + sc->first_laser_angle_offset) + i * period;
            double th = ang * DEG2RAD;
            double ex = cx + cos(th) * L, ey = cy + sin(th) * L;
            double bsel = 0, bx_ = ex, by_ = ey;
            int borient = -1;
            if (sn->valid && sn->corr_head - sn->corr_tail > 1) {
                const double* f = corr + 4 * RING(sn->corr_head - 1);
                const double* b = corr + 4 * RING(sn->corr_tail);
                compas_test(f[0], f[1], f[2], f[3], cx, cy, ex, ey, 0, &bsel, &bx_, &by_, &borient);
                compas_test(b[0], b[1], b[2], b[3], cx, cy, ex, ey, 1, &bsel, &bx_, &by_, &borient);
                for (int q = sn->corr_tail; q < sn->corr_head - 1; q++) {
                    const double* a = corr + 4 * RING(q);
                    const double* n = corr + 4 * RING(q + 1);
                    compas_test(a[2], a[3], n[2], n[3], cx, cy, ex, ey, 2, &bsel, &bx_, &by_, &borient);
                }
                for (int q = sn->corr_tail; q < sn->corr_head - 1; q++) {
                    const double* a = corr + 4 * RING(q);
                    const double* n = corr + 4 * RING(q + 1);
                    compas_test(a[0], a[1], n[0], n[1], cx, cy, ex, ey, 3, &bsel, &bx_, &by_, &borient);
                }
            }
            double ddx = bx_ - cx, ddy = by_ - cy;
            float v = (float)sqrt(fma(ddy, ddy, ddx * ddx)); /* np.linalg.norm(collide - pos), 1-D: the FMA form */
            row[borient < 0 ? i : i + R * (1 + borient)] = v;
        }
    }
}

static void ray_sensor_scan(const FtlOracle* o, int env_index, int sensor, float* out) {
    const FtlConfig* cfg = &o->cfg;
    const FtlEnvState* e = &o->env[env_index];
    const FtlRaySensorConfig* sc = &cfg->ray[sensor];
    if (sc->compas) { compas_sensor_scan(o, env_index, sensor, out); return; }
    const float* corr = o->corridor + (size_t)env_index * cfg->corridor_cap * 4;
    int R = sc->lasers_count, H = sc->max_prev_obs;
    int width = sc->pad_sectors ? 4 * R : R;
    float px = e->follower.pos[0], py = e->follower.pos[1];
    double L = sc->laser_length, period = 360.0 / R;
    static _Thread_local float ed[MAX_EDGES * 4];

    for (int j = 0; j < H; j++) {
        const FtlSnapshot* sn = &e->snap[FTL_MAX_HIST - H + j];
        int ne = sn->valid ? collect_edges(o, e, sc, sn, corr, ed) : 0;
        float* row = out + (size_t)j * width;
        for (int k = 0; k < width; k++) row[k] = 0.f;
        for (int i = 0; i < R; i++) {
            /* SEN:888-891; the fan of LeaderCorridor_lasers (SEN:678-700) comes in as custom angles */
            double ang = (e->follower.dir + sc->first_laser_angle_offset) +
                         (sc->n_custom_angles ? sc->custom_angle[i] : i * period);
            double th = ang * DEG2RAD;
            double ex = (double)px + cos(th) * L, ey = (double)py + sin(th) * L; /* laser end, float64 */
            double best = -1, bx_ = 0, by_ = 0;
            for (int k = 0; k < ne; k++) {
                float ax = ed[4 * k], ay = ed[4 * k + 1], bx = ed[4 * k + 2], by = ed[4 * k + 3];
                int hit = (ccw_ffd(ax, ay, px, py, ex, ey) != ccw_ffd(bx, by, px, py, ex, ey)) &&
                          (ccw_fff(ax, ay, bx, by, px, py) != ccw_ffD(ax, ay, bx, by, ex, ey));
                if (!hit) continue;
                /* seg_intersect(a1=A, a2=B, b1=pos, b2=end), SEN:628-640 */
                float dax = bx - ax, day = by - ay;        /* da, float32 */
                double dbx = ex - (double)px, dby = ey - (double)py; /* db, float64 */
                float dpx = ax - px, dpy = ay - py;        /* dp, float32 */
                float dapx = -day, dapy = dax;             /* perp(da) */
                double denom = fma((double)dapx, dbx, (double)dapy * dby); /* np.dot((k,2),(2,1)) */
                float m0 = dapx * dpx, m1 = dapy * dpy;
                float num = m0 + m1;                       /* float32 */
                double t = (double)num / denom;
                double xx = t * dbx + (double)px, xy = t * dby + (double)py;
                double dist = dist_f64(xx, xy, (double)px, (double)py);
                if (best < 0 || dist < best) { best = dist; bx_ = xx; by_ = xy; }
            }
            if (best < 0) { bx_ = ex; by_ = ey; } /* no hit: the laser end point, SEN:926 */
            /* obs_item[i] = np.linalg.norm(collide - pos): axis=None -> x.dot(x) (ddot, FMA form) */
            double ddx = bx_ - (double)px, ddy = by_ - (double)py;
            double val = sqrt(fma(ddy, ddy, ddx * ddx));
            float fv = (float)val;
            if (sc->pad_sectors) { /* SEN:932-953 */
                double in_sector = R / 4.0;
                int sector = (i < in_sector) ? 0 : (i < 2 * in_sector) ? 1 : (i < 3 * in_sector) ? 2 : 3;
                row[sector * R + i] = fv;
            } else {
                row[i] = fv;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* sensors + observation                                                                             */
/* ------------------------------------------------------------------------------------------------ */
static int sensor_width(const FtlRaySensorConfig* sc) {
    return sc->max_prev_obs * (sc->compas ? 5 * sc->lasers_count : sc->pad_sectors ? 4 * sc->lasers_count : sc->lasers_count);
}

static void write_outputs(FtlOracle* o, int i, const FtlOutputs* out) {
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[i];
    if (out->numerical_features) { /* ENV:1793-1802 */
        float* nf = out->numerical_features + (size_t)i * 10;
        nf[0] = e->leader.pos[0]; nf[1] = e->leader.pos[1];
        nf[2] = (float)e->leader.speed; nf[3] = (float)e->leader.dir; nf[4] = (float)e->leader.rot_speed;
        nf[5] = e->follower.pos[0]; nf[6] = e->follower.pos[1];
        nf[7] = (float)e->follower.speed; nf[8] = (float)e->follower.dir; nf[9] = (float)e->follower.rot_speed;
    }
    if (out->leader_target) { /* ENV:1803-1806 */
        const int32_t* route = o->s_route + (size_t)e->scenario_id * cfg->route_cap * 2;
        int n_route = o->s_nroute[e->scenario_id];
        int tid = e->cur_target_id < n_route ? e->cur_target_id : n_route - 1;
        if (n_route > 1 && route[2 * tid] == route[2 * (n_route - 1)] && route[2 * tid + 1] == route[2 * (n_route - 1) + 1])
            tid = n_route - 2;
        out->leader_target[2 * i] = route[2 * tid];
        out->leader_target[2 * i + 1] = route[2 * tid + 1];
    }
    if (out->reward) out->reward[i] = (float)e->last_reward;
    if (out->done) out->done[i] = (uint8_t)e->done;
    if (out->status) {
        out->status[4 * i] = (uint8_t)e->mission_status;
        out->status[4 * i + 1] = (uint8_t)e->agent_status;
        out->status[4 * i + 2] = (uint8_t)e->leader_status;
        out->status[4 * i + 3] = (uint8_t)e->crash;
    }
}

/* LeaderTrackDetector_radar.scan, SEN:425-461, with rotateVector / calculateAngle of MSC:47-62.
 * chosen_dots = np.array(slice of the history): float64 as soon as one chosen point is one of the float64 points
 * the tracker seeded (SEN:257-272), float32 otherwise; everything derived from it (the vectors to the follower,
 * np.linalg.norm(axis=1) = sqrt(x*x + y*y) without fusion) keeps that type.  v.dot(w) is a float64 gemv
 * (fma(v0, w0, v1*w1), like the other (k,2).(2,) products of this file), np.linalg.norm(w) = sqrt(ddot) (fma form). */
static void radar_scan(FtlOracle* o, int i, float* radar) {
    const FtlConfig* cfg = &o->cfg;
    const FtlEnvState* e = &o->env[i];
    const int R = cfg->radar_sectors, cap = cfg->corridor_cap;
    const double* hist = o->hist + (size_t)i * cap * 2;
    for (int k = 0; k < R; k++) radar[k] = 0.f;
    const int n = e->ring_head - e->ring_tail;
    if (n <= 0) return;
    int cnt = n < cfg->radar_len ? n : cfg->radar_len, first = e->ring_tail;
    if (cfg->radar_mode == 0) first = e->ring_head - cnt;          /* "new": islice(len - P, len) */
    else if (cfg->radar_mode == 2) cnt = n;                        /* "near": every point (the sort does not matter) */
    const int is64 = first < e->hist_f64_end;
    const double dir = e->follower.dir;
    double rdir = dir + 90;
    if (rdir >= 360) rdir -= 360;
    const double th = dir * (M_PI / 180.0), thr = rdir * (M_PI / 180.0); /* np.radians */
    const double wdx = cos(th), wdy = sin(th), wrx = cos(thr), wry = sin(thr); /* np.dot(rot, [1, 0]) */
    const double nwd = sqrt(fma(wdy, wdy, wdx * wdx)), nwr = sqrt(fma(wry, wry, wrx * wrx));
    const double sa = M_PI / R;
    const float px = e->follower.pos[0], py = e->follower.pos[1];
    for (int k = 0; k < cnt; k++) {
        const double hx = hist[2 * RING(first + k)], hy = hist[2 * RING(first + k) + 1];
        double vx, vy, dist;
        if (is64) {
            vx = hx - (double)px; vy = hy - (double)py;
            const double sx = vx * vx, sy = vy * vy;
            dist = sqrt(sx + sy);
        } else {
            const float fx = (float)hx - px, fy = (float)hy - py;
            const float sx = fx * fx, sy = fy * fy;
            const float s = sx + sy;
            dist = (double)sqrtf(s);
            vx = fx; vy = fy;
        }
        const double ad = acos(fma(vx, wdx, vy * wdy) / (dist * nwd));
        double ar = acos(fma(vx, wrx, vy * wry) / (dist * nwr));
        if (ad > M_PI / 2) ar = -ar;
        for (int q = 0; q < R; q++) { /* SEN:454-459, literally */
            if (ar >= sa * q && ar < sa * (q + 1)) {
                const float d32 = (float)dist;
                if (radar[q] == 0.f || d32 < radar[q]) radar[q] = d32;
            }
        }
    }
}

/* LaserSensor.scan, SEN:63-136, with distance_to_rect of MSC:29-44.  Operand types as numpy >= 2 sees them: the follower
 * position is a float32 array element, every other operand a python float (weak): beam end and sample points are float32. */
static int laser_rect_in_range(const int32_t* r, float px, float py, double limit) {
    /* topleft, bottomleft, topright, bottomright, midtop, midleft, midbottom, midright (pygame: mid = x + w // 2) */
    int cx = r[0] + (r[2] >> 1), cy = r[1] + (r[3] >> 1);
    int pts[8][2] = {{r[0], r[1]}, {r[0], r[1] + r[3]}, {r[0] + r[2], r[1]}, {r[0] + r[2], r[1] + r[3]},
                     {cx, r[1]}, {r[0], cy}, {cx, r[1] + r[3]}, {r[0] + r[2], cy}};
    double best = INFINITY;
    for (int k = 0; k < 8; k++) {
        double d = dist_f64((double)px, (double)py, (double)pts[k][0], (double)pts[k][1]); /* scipy euclidean */
        if (d < best) best = d;
    }
    return best <= limit;
}
static int laser_collidepoint(const int32_t* r, float px, float py) { /* pygame.Rect.collidepoint on truncated ints */
    int x = (int)px, y = (int)py;
    return x >= r[0] && x < r[0] + r[2] && y >= r[1] && y < r[1] + r[3];
}
static void laser_scan(FtlOracle* o, int i, float* out) {
    const FtlConfig* cfg = &o->cfg;
    const FtlEnvState* e = &o->env[i];
    float x1 = e->follower.pos[0], y1 = e->follower.pos[1];
    /* objects_in_range: game_object_list minus the follower (leader, walls, rocks) + game_dynamic_list (bears) */
    const int32_t* objs[1 + 64 + FTL_MAX_BEARS];
    int no = 0;
    double limit = cfg->laser_range + cfg->laser_reach_extra;
    if (laser_rect_in_range(e->leader.rect, x1, y1, limit)) objs[no++] = e->leader.rect;
    const int32_t* sr = o->s_static + (size_t)e->scenario_id * cfg->static_cap * 4;
    for (int k = 0; k < o->s_nstatic[e->scenario_id]; k++)
        if (laser_rect_in_range(sr + 4 * k, x1, y1, limit)) objs[no++] = sr + 4 * k;
    for (int b = 0; b < cfg->n_bears; b++)
        if (laser_rect_in_range(e->bear[b].rect, x1, y1, limit)) objs[no++] = e->bear[b].rect;
    int border = (int)(cfg->laser_available_angle / 2);
    int beam = 0;
    double diff = 0;
    while (beam < cfg->laser_beams) {
        double angles[2];
        int na = 1;
        if (beam == 0) {
            angles[0] = -e->follower.dir;
        } else {
            if (!(diff < border)) break;
            diff += cfg->laser_angle_step;
            angles[0] = angle_correction(-e->follower.dir + diff);
            angles[1] = angle_correction(-e->follower.dir - diff);
            na = 2;
        }
        for (int a = 0; a < na; a++, beam++) {
            double th = angles[a] * DEG2RAD; /* math.radians */
            float x2 = x1 + (float)(cfg->laser_range * cos(th)), y2 = y1 - (float)(cfg->laser_range * sin(th));
            float hx = x2, hy = y2;
            for (int k = 0; k < cfg->laser_points; k++) {
                double u = (double)k / cfg->laser_points;
                float uf = (float)u, vf = (float)(1 - u);
                float ax = x2 * uf, bx = x1 * vf, ay = y2 * uf, by = y1 * vf;
                float cx = ax + bx, cy = ay + by;
                int hit = 0;
                for (int q = 0; q < no && !hit; q++) hit = laser_collidepoint(objs[q], cx, cy);
                if (hit) { hx = cx; hy = cy; break; }
            }
            float dx = hx - x1, dy = hy - y1;
            if (cfg->laser_only_distances) {
                float sx = dx * dx, sy = dy * dy;
                out[beam] = sqrtf(sx + sy);
            } else {
                out[2 * beam] = dx;
                out[2 * beam + 1] = dy;
            }
        }
    }
}

static void use_sensors(FtlOracle* o, int i, const FtlOutputs* out) { /* CLS:255-288 */
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[i];
    if (cfg->tracker_enabled)
        for (int k = 0; k < cfg->tracker_scans_per_step; k++) tracker_scan(o, i);
    if (out && out->follower_info) { /* FollowerInfo.scan, SEN:834-842: float64 quotients stored as float32 */
        out->follower_info[2 * i] = (float)(e->follower.speed / cfg->follower.max_speed);
        out->follower_info[2 * i + 1] = (float)(e->follower.dir / 360);
    }
    if (out && out->track_vectors && cfg->track_vector_len > 0) { /* LeaderTrackDetector_vector.scan, SEN:365-380 */
        const int P = cfg->track_vector_len, cap = cfg->corridor_cap;
        const double* hist = o->hist + (size_t)i * cap * 2;
        float* v = out->track_vectors + (size_t)i * P * 2;
        int len = e->ring_head - e->ring_tail;
        int cnt = len < P ? len : P;
        int first = cfg->track_vector_mode == 0 ? e->ring_head - cnt : e->ring_tail; /* "new": the last P, "old": the first P */
        for (int k = 0; k < 2 * P; k++) v[k] = 0.f;
        for (int k = 0; k < cnt; k++) { /* np.array(slice) - position (float32), cast into the float32 buffer */
            v[2 * k] = (float)(hist[2 * RING(first + k)] - (double)e->follower.pos[0]);
            v[2 * k + 1] = (float)(hist[2 * RING(first + k) + 1] - (double)e->follower.pos[1]);
        }
    }
    if (out && out->radar && cfg->radar_sectors > 0) radar_scan(o, i, out->radar + (size_t)i * cfg->radar_sectors);
    if (out && out->laser && cfg->laser_points > 0)
        laser_scan(o, i, out->laser + (size_t)i * cfg->laser_beams * (cfg->laser_only_distances ? 1 : 2));
    if (cfg->n_ray_sensors > 0) {
        /* history_obstacles_list.pop(0); append(current), SEN:894-895 (one shared ring: every sensor
         * snapshots the same world at the same instants) */
        int live = e->ring_head - e->ring_tail;
        if (live > 1) {
            memmove(&e->snap[0], &e->snap[1], sizeof(FtlSnapshot) * (FTL_MAX_HIST - 1));
            FtlSnapshot* sn = &e->snap[FTL_MAX_HIST - 1];
            sn->valid = 1;
            sn->corr_tail = e->ring_tail;
            sn->corr_head = e->ring_head;
            sn->pad_ = 0;
            memcpy(sn->dyn_rect[0], e->leader.rect, sizeof(int32_t) * 4);
            for (int b = 0; b < FTL_MAX_BEARS; b++)
                if (b < cfg->n_bears)
                    memcpy(sn->dyn_rect[1 + b], e->bear[b].rect, sizeof(int32_t) * 4);
                else
                    memset(sn->dyn_rect[1 + b], 0, sizeof(int32_t) * 4);
            if (e->snap_pushes < (1 << 30)) e->snap_pushes++;
        }
        if (out && out->rays) {
            float* dst = out->rays + (size_t)i * o->rays_per_env;
            for (int s = 0; s < cfg->n_ray_sensors; s++) {
                if (live > 1) {
                    ray_sensor_scan(o, i, s, dst);
                } else { /* the reference raises here (SEN:892, unbound all_obs_arr); report "no hit" */
                    const FtlRaySensorConfig* sc = &cfg->ray[s];
                    for (int k = 0; k < sensor_width(sc); k++) dst[k] = (float)sc->laser_length;
                }
                dst += sensor_width(&cfg->ray[s]);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------ */
/* reset / step of one env                                                                           */
/* ------------------------------------------------------------------------------------------------ */
static void robot_init(FtlRobotState* r, const FtlRobotConfig* c, float x, float y, double dir) {
    memset(r, 0, sizeof *r);
    r->pos[0] = x;
    r->pos[1] = y;
    r->dir = dir;
    rect_place(r, c->width, c->height);
}

static void env_reset(FtlOracle* o, int i, int scenario, const FtlOutputs* out) { /* ENV:434-543 with the scenario as data */
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[i];
    int accel_consumed = e->accel_consumed; /* the reference never restores the consumed keys, ENV:1170 */
    int episodes = e->episode_count;
    memset(e, 0, sizeof *e);
    e->accel_consumed = accel_consumed;
    e->episode_count = episodes + 1;
    e->scenario_id = scenario;
    robot_init(&e->leader, &cfg->leader, o->s_lpos[2 * scenario], o->s_lpos[2 * scenario + 1], o->s_ldir[scenario]);
    robot_init(&e->follower, &cfg->follower, o->s_fpos[2 * scenario], o->s_fpos[2 * scenario + 1], o->s_fdir[scenario]);
    for (int b = 0; b < cfg->n_bears; b++) { /* _create_dyn_obs + _reset_pose_bear, ENV:687-718, 761-770 */
        float bx = (b % 2 == 0) ? e->leader.pos[0] + 150.f : e->leader.pos[0] - 150.f;
        float by = (b % 2 == 0) ? e->leader.pos[1] - 150.f : e->leader.pos[1] + 150.f;
        robot_init(&e->bear[b], &cfg->bear, bx, by, 0.0);
        float tx = e->leader.pos[0] - 150.f, ty = e->leader.pos[1] - 150.f;
        e->bear_target[b][0] = tx;
        e->bear_target[b][1] = ty;
        e->bear_index[b] = 0;
    }
    e->cur_target_id = 1;
    e->finish_timer = -1;
    e->cur_speed_multiplier = 1;
    /* seed the trail, ENV:533-539: float32 linspace follower -> leader */
    float* trail = o->trail + (size_t)i * cfg->trail_cap * 2;
    float fx = e->follower.pos[0], fy = e->follower.pos[1], lx = e->leader.pos[0], ly = e->leader.pos[1];
    float denom = (float)(cfg->trajectory_saving_period * cfg->leader.max_speed);
    float q = dist_f32(fx, fy, lx, ly) / denom;
    int m = (int)q;
    if (m > cfg->trail_cap) { m = cfg->trail_cap; e->overflow |= 1; }
    if (m > 0) {
        float dx = lx - fx, dy = ly - fy;
        float stepx = m > 1 ? dx / (float)(m - 1) : 0.f, stepy = m > 1 ? dy / (float)(m - 1) : 0.f;
        for (int k = 0; k < m; k++) {
            float tx = (float)k * stepx, ty = (float)k * stepy;
            float vx = tx + fx, vy = ty + fy;
            trail[2 * k] = vx;
            trail[2 * k + 1] = vy;
        }
        if (m > 1) {
            trail[2 * (m - 1)] = lx;
            trail[2 * (m - 1) + 1] = ly;
        }
    }
    e->trail_len = m;
    e->mission_status = FTL_MISSION_IN_PROGRESS;
    use_sensors(o, i, out);
    if (out) write_outputs(o, i, out);
}

static void env_step(FtlOracle* o, int i, const void* actions, const FtlOutputs* out) { /* ENV:908-945 */
    const FtlConfig* cfg = &o->cfg;
    FtlEnvState* e = &o->env[i];
    double a0, a1;
    if (cfg->action_mode == FTL_ACTION_DISCRETE) {
        int a = ((const int32_t*)actions)[i];
        a0 = cfg->follower.max_speed;
        a1 = cfg->discrete_rotation_table[a];
    } else if (cfg->action_mode == FTL_ACTION_CONST_SPEED) {
        a0 = cfg->const_speed_action;
        a1 = (double)((const float*)actions)[i];
    } else {
        a0 = (double)((const float*)actions)[2 * i];
        a1 = (double)((const float*)actions)[2 * i + 1];
    }
    command_forward(&e->follower, &cfg->follower, a0);
    if (a1 < 0)
        command_turn(&e->follower, &cfg->follower, fabs(a1), -1);
    else if (a1 > 0)
        command_turn(&e->follower, &cfg->follower, a1, 1);
    else
        command_turn(&e->follower, &cfg->follower, 0, 0);
    int fps = cfg->frames_per_step;
    if (o->in_frames) {   /* random_frames_per_step: the caller drew this step's count, ENV:939-940 */
        fps = o->in_frames[i];
        if (fps < 1) fps = 1;
        if (fps > cfg->frames_per_step) fps = cfg->frames_per_step;
    }
    const double* draws = o->in_draws ? o->in_draws + (size_t)i * cfg->frames_per_step : NULL;
    for (int f = 0; f < fps; f++) frame_step(o, i, fps, draws ? draws + f : NULL);
    use_sensors(o, i, out);
    if (out) write_outputs(o, i, out);
}

/* ------------------------------------------------------------------------------------------------ */
/* public API (ctypes)                                                                               */
/* ------------------------------------------------------------------------------------------------ */
static int rays_per_env(const FtlConfig* cfg) {
    int n = 0;
    for (int s = 0; s < cfg->n_ray_sensors; s++) n += sensor_width(&cfg->ray[s]);
    return n;
}

int ftl_oracle_sizeof_env_state(void) { return (int)sizeof(FtlEnvState); }
int ftl_oracle_sizeof_config(void) { return (int)sizeof(FtlConfig); }

FtlOracle* ftl_oracle_create(const FtlConfig* cfg, int n_envs, int64_t env_id_base, int n_threads) {
    if (!cfg || cfg->abi_version != FTL_ABI_VERSION || n_envs <= 0) return NULL;
    if (cfg->corridor_cap & (cfg->corridor_cap - 1)) return NULL;
    if (cfg->n_bears > FTL_MAX_BEARS || cfg->n_ray_sensors > FTL_MAX_RAY_SENSORS) return NULL;
    FtlOracle* o = (FtlOracle*)calloc(1, sizeof *o);
    o->cfg = *cfg;
    o->n_envs = n_envs;
    o->env_id_base = env_id_base;
    o->n_threads = n_threads > 0 ? n_threads : 1;
    o->env = (FtlEnvState*)calloc(n_envs, sizeof(FtlEnvState));
    o->trail = (float*)calloc((size_t)n_envs * cfg->trail_cap * 2, sizeof(float));
    o->hist = (double*)calloc((size_t)n_envs * cfg->corridor_cap * 2, sizeof(double));
    o->corridor = (float*)calloc((size_t)n_envs * cfg->corridor_cap * 4, sizeof(float));
    o->corridor64 = (double*)calloc((size_t)n_envs * cfg->corridor_cap * 4, sizeof(double));
    o->rays_per_env = rays_per_env(cfg);
    return o;
}

void ftl_oracle_destroy(FtlOracle* o) {
    if (!o) return;
    free(o->s_static); free(o->s_nstatic); free(o->s_route); free(o->s_nroute);
    free(o->s_lpos); free(o->s_fpos); free(o->s_ldir); free(o->s_fdir);
    free(o->env); free(o->trail); free(o->hist); free(o->corridor); free(o->corridor64);
    free(o);
}

int ftl_oracle_rays_per_env(const FtlOracle* o) { return o->rays_per_env; }

static void* dup_mem(const void* p, size_t bytes) {
    void* q = malloc(bytes ? bytes : 1);
    memcpy(q, p, bytes);
    return q;
}

int ftl_oracle_upload_scenarios(FtlOracle* o, const FtlScenarioPool* p) {
    if (p->static_cap != o->cfg.static_cap || p->route_cap != o->cfg.route_cap || p->n_scenarios <= 0) return FTL_ERR_INVALID;
    size_t S = p->n_scenarios;
    free(o->s_static); free(o->s_nstatic); free(o->s_route); free(o->s_nroute);
    free(o->s_lpos); free(o->s_fpos); free(o->s_ldir); free(o->s_fdir);
    o->n_scen = (int)S;
    o->s_static = (int32_t*)dup_mem(p->static_rects, S * p->static_cap * 4 * sizeof(int32_t));
    o->s_nstatic = (int32_t*)dup_mem(p->n_static, S * sizeof(int32_t));
    o->s_route = (int32_t*)dup_mem(p->route, S * p->route_cap * 2 * sizeof(int32_t));
    o->s_nroute = (int32_t*)dup_mem(p->n_route, S * sizeof(int32_t));
    o->s_lpos = (float*)dup_mem(p->leader_pos, S * 2 * sizeof(float));
    o->s_fpos = (float*)dup_mem(p->follower_pos, S * 2 * sizeof(float));
    o->s_ldir = (double*)dup_mem(p->leader_dir, S * sizeof(double));
    o->s_fdir = (double*)dup_mem(p->follower_dir, S * sizeof(double));
    return FTL_OK;
}

/* next scenario of env i when the caller does not name one: round-robin over the pool in strides of
 * the global env count, the same rule the CUDA path uses */
static int next_scenario(const FtlOracle* o, int i) {
    const FtlEnvState* e = &o->env[i];
    int64_t g = o->env_id_base + i;
    return (int)((g + (int64_t)e->episode_count * 7919) % o->n_scen);
}

int ftl_oracle_reset(FtlOracle* o, const uint8_t* mask, const int32_t* scenario_ids, const FtlOutputs* out) {
    if (!o->n_scen) return FTL_ERR_STATE;
#pragma omp parallel for num_threads(o->n_threads) schedule(static)
    for (int i = 0; i < o->n_envs; i++) {
        if (mask && !mask[i]) continue;
        int s = scenario_ids ? scenario_ids[i] : next_scenario(o, i);
        env_reset(o, i, s, out);
    }
    return FTL_OK;
}

int ftl_oracle_step(FtlOracle* o, const void* actions, const FtlOutputs* out) {
    if (!o->n_scen) return FTL_ERR_STATE;
    int auto_reset = o->cfg.auto_reset;
#pragma omp parallel for num_threads(o->n_threads) schedule(dynamic, 64)
    for (int i = 0; i < o->n_envs; i++) {
        env_step(o, i, actions, out);
        if (auto_reset && o->env[i].done) {
            /* reward/done/status of the finished episode stay in `out`; the observation is replaced
             * by the first observation of the next episode (vector-env convention) */
            FtlOutputs o2 = *out;
            o2.reward = NULL; o2.done = NULL; o2.status = NULL;
            env_reset(o, i, next_scenario(o, i), &o2);
        }
    }
    return FTL_OK;
}

int ftl_oracle_step_ex(FtlOracle* o, const void* actions, const int32_t* frames, const double* draws, const FtlOutputs* out) {
    o->in_frames = frames;
    o->in_draws = draws;
    int rc = ftl_oracle_step(o, actions, out);
    o->in_frames = NULL;
    o->in_draws = NULL;
    return rc;
}

int ftl_oracle_get_state(FtlOracle* o, int first, int n, const FtlStateBuffers* b) {
    const FtlConfig* cfg = &o->cfg;
    if (first < 0 || n < 0 || first + n > o->n_envs) return FTL_ERR_INVALID;
    if (b->env) memcpy(b->env, o->env + first, sizeof(FtlEnvState) * n);
    if (b->trail) memcpy(b->trail, o->trail + (size_t)first * cfg->trail_cap * 2, sizeof(float) * 2 * cfg->trail_cap * n);
    if (b->hist) memcpy(b->hist, o->hist + (size_t)first * cfg->corridor_cap * 2, sizeof(double) * 2 * cfg->corridor_cap * n);
    if (b->corridor) memcpy(b->corridor, o->corridor + (size_t)first * cfg->corridor_cap * 4, sizeof(float) * 4 * cfg->corridor_cap * n);
    return FTL_OK;
}

int ftl_oracle_set_state(FtlOracle* o, int first, int n, const FtlStateBuffers* b) {
    const FtlConfig* cfg = &o->cfg;
    if (first < 0 || n < 0 || first + n > o->n_envs) return FTL_ERR_INVALID;
    if (b->env) memcpy(o->env + first, b->env, sizeof(FtlEnvState) * n);
    if (b->trail) memcpy(o->trail + (size_t)first * cfg->trail_cap * 2, b->trail, sizeof(float) * 2 * cfg->trail_cap * n);
    if (b->hist) memcpy(o->hist + (size_t)first * cfg->corridor_cap * 2, b->hist, sizeof(double) * 2 * cfg->corridor_cap * n);
    if (b->corridor) {
        memcpy(o->corridor + (size_t)first * cfg->corridor_cap * 4, b->corridor, sizeof(float) * 4 * cfg->corridor_cap * n);
        /* an injected state only carries the float32 points */
        for (size_t k = 0; k < (size_t)4 * cfg->corridor_cap * n; k++)
            o->corridor64[(size_t)first * cfg->corridor_cap * 4 + k] = (double)b->corridor[k];
    }
    return FTL_OK;
}
