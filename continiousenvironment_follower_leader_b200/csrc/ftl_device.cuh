// ftl_device.cuh -- per-environment logic of the batched "follow the leader" simulator.
//
// Everything here is FTL_HD (host + device) inline code: the kernels in ftl_kernels.cu call it with
// one thread per environment, and tests/hostsim compiles the very same functions for the CPU so the
// kernel logic is unit-tested in a container without a GPU.  (The host build is a test target; the
// product path is libftl.so only.)
//
// What each block restates (paths relative to the reference root, ENV/CLS/SEN/MSC as in
// include/ftl.h): the arithmetic TYPES are chosen so that results match the reference bit for bit
// wherever the reference's own result is type-stable: python floats -> double, np.float32 arrays ->
// float with separately rounded operations (the translation unit is compiled with -fmad=false;
// fused operations are written explicitly).
//
// The structure is NOT the reference's: state is structure-of-arrays in HBM, the F sub-frames of a
// step run in registers, the O(trail) green-zone searches are replaced by cached exact bounds
// (see GreenCache), and static-obstacle collision tests are pre-filtered once per step.
#pragma once

#include <math.h>
#include <stdint.h>

#include "../../include/ftl.h"

#if defined(__CUDACC__)
#define FTL_HD __host__ __device__ __forceinline__
#define FTL_HD_NOINLINE static __host__ __device__ __noinline__
#else
#define FTL_HD inline
#define FTL_HD_NOINLINE static
#include <cstring>
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
#endif

#if defined(FTL_COUNT_HOOK) && !defined(__CUDA_ARCH__)
void ftl_count_hook(int k, long long v);   // diagnostic host build only (tools/scan_stats.py): one event per call
#define FTL_COUNT(k, v) ftl_count_hook(k, v)
#elif defined(FTL_COUNT_RESCANS) && !defined(__CUDA_ARCH__)
extern long long g_ftl_counters[8];   // test-only instrumentation (tests/hostsim)
#define FTL_COUNT(k, v) (g_ftl_counters[k] += (v))
#else
#define FTL_COUNT(k, v) ((void)0)
#endif

namespace ftl {

constexpr int kMaxRobots = 2 + FTL_MAX_BEARS;  // follower, leader, bears
constexpr double kDeg2Rad = 3.14159265358979323846 / 180.0;
constexpr double kRad2Deg = 180.0 / 3.14159265358979323846;

// ---- state layout (structure of arrays, env index fastest) ---------------------------------------
enum GlobalF64 { GD_ACC_PENALTY, GD_OVERALL, GD_LAST_REWARD, GD_SPEED_MULT, GD_LEAD_ACC, GD_LEAD_CUM, GD_COUNT };
enum RobotF64 { RD_DIR, RD_SPEED, RD_ROT, RD_DES_SPEED, RD_DES_ROT, RD_COUNT };
enum GlobalI32 {
    GI_STEP_COUNT, GI_TARGET_ID, GI_FINISH_TIMER, GI_FLAGS, GI_TRAIL_LEN, GI_SAVING_COUNTER, GI_RING_TAIL,
    GI_RING_HEAD, GI_HIST_F64_END, GI_SNAP_PUSHES, GI_SCENARIO, GI_EPISODE, GI_OVERFLOW, GI_ACCEL_CONSUMED,
    // derived caches (not part of FtlEnvState; invalidated by set_state/reset)
    GI_G_LO, GI_G_UNC, GI_A_STAR, GI_B_STAR, GI_NICE_FROM, GI_COUNT
};
enum GlobalF32 { GF_LB_GREEN, GF_LB_ALL, GF_SCAN_GX, GF_SCAN_GY, GF_SCAN_AX, GF_SCAN_AY, GF_COUNT };

// flag bits of GI_FLAGS
enum {
    FL_DONE = 1, FL_CRASH = 2, FL_LEADER_FINISHED = 4, FL_IN_BOX = 8, FL_ON_TRACE = 16, FL_TOO_CLOSE = 32,
    FL_MISSION_SHIFT = 8, FL_AGENT_SHIFT = 10, FL_LEADER_SHIFT = 13
};

struct UncRec { float ax, ay, bx, by; int f; int rows; };  // an (edge, ray) pair left to the exact ray pass

struct DevState {
    int n;             // padded to a multiple of 32 (array stride and grid size); envs >= n_real are filler
    int n_real;        // environments the caller sees
    int n_bears;
    double* gd;        // [GD_COUNT][n]
    double* rd;        // [2+n_bears][RD_COUNT][n]
    double* bear_tgt;  // [n_bears][2][n]
    int* gi;           // [GI_COUNT][n]
    int* ri;           // [2+n_bears][n]   (rot_dir+1) | (des_rot_dir+1)<<2
    int* bear_idx;     // [n_bears][n]
    float* gf;         // [GF_COUNT][n]
    float2* pos;       // [2+n_bears][n]
    int4* rect;        // [2+n_bears][n]
    float2* trail;     // [n][trail_cap]
    float* trail_d;    // [n][trail_cap]  float32 distance to the previous trail point (the reference's term)
    double* trail_s;   // [n][trail_cap]  cumulative arc length in float64 (derived; decides the green window)
    double2* hist;     // [n][corridor_cap]
    float4* corridor;  // [n][corridor_cap]   (right.x, right.y, left.x, left.y)
    double* seg_d;     // [n][corridor_cap]   |hist[k+1] - hist[k]| in float64 and ...
    float* seg_f;      // [n][corridor_cap]   ... in float32 (derived: the tracker's length test sums them, SEN:283-287)
    int2* snap_range;  // [FTL_MAX_HIST][n]   ring slot = push index % FTL_MAX_HIST
    int4* snap_rect;   // [FTL_MAX_HIST][1+n_bears][n]
    UncRec* unc_rec;   // [n][kUncPerEnv]  scratch of the ray pass
    int* unc_count;    // [n]
    // per-frame records the kinematics kernel leaves for the bookkeeping kernel (frame-major, env fastest)
    float2* rec_f;     // [frames_per_step][n]  follower position after frame j
    float2* rec_l;     // [frames_per_step][n]  leader position after frame j
    unsigned char* rec_bits;   // [frames_per_step][n]  REC_* bits of frame j
    unsigned char* rec_lbits;  // [frames_per_step][n]  ... the leader's bits when a separate warp computes them
    // optional per-step inputs (FtlStepInputs; NULL = configuration value / Philox), set by the host before each launch
    const int* in_frames;      // [n]  frames of this step per env
    const double* in_draws;    // [n][frames_per_step]  uniform draws of list-valued speed regimes, per frame
    int* kin_flag;     // [n / 32]  sequence number of the last step whose kinematics the owning warp of k_kin has published
    int* book_flag;    // [n / 32]  ... whose bookkeeping the owning warp of k_book has published
};

struct DevPool {  // scenario pool in device (or host, for hostsim) memory
    int n_scenarios;
    const int4* static_rects;  // [S][static_cap]
    const int* n_static;       // [S]
    const int2* route;         // [S][route_cap]
    const int* n_route;        // [S]
    const float2* leader_pos;
    const double* leader_dir;
    const float2* follower_pos;
    const double* follower_dir;
    // optional (NULL: every rectangle is examined): per scenario a grid of 64-px cells over the field; bit k of a cell
    // is set when static rectangle k, grown by near_grid_inflate, touches the cell -- a superset of the rectangles
    // near_rect() can accept for any position inside the cell
    const uint64_t* near_grid;   // [S][near_grid_h][near_grid_w]
    int near_grid_w, near_grid_h;
};
constexpr int kNearGridShift = 6;

struct DevOutputs {
    int n;  // number of real envs: rows >= n are never written
    float* numerical_features;
    int* leader_target;
    float* rays;
    float* reward;
    uint8_t* done;
    uint8_t* status;
    float* follower_info;   // [N][2] or NULL
    float* track_vectors;   // [N][track_vector_len][2] or NULL
    float* radar;           // [N][radar_sectors] or NULL
    float* laser;           // [N][laser_beams][2] (or [N][laser_beams]) or NULL
};

// per-sensor constants of the ray pass that do not depend on the env (filled by ray_static_tables, ftl_rays.cuh)
struct RaySensorStatic {
    int base, R, H, cls_mask;      // first flat ray, rays, history rows, EdgeClass bits the sensor reacts to
    float L, inv_period, eps, inv_R;
};

// Configuration as the kernels want it: the user's FtlConfig plus values derived once on the host.
struct DevCfg {
    FtlConfig c;
    int64_t env_id_base;
    int rays_per_env;
    int rays_total;              // number of rays over all sensors
    int ray_hmax;                // largest max_prev_obs over the ray sensors (rows of minima the ray pass keeps)
    RaySensorStatic ray_static[FTL_MAX_RAY_SENSORS];
    float ray_reach[8];          // per EdgeClass: largest laser_length among the sensors reacting to it (or < 0)
    // float32 thresholds on SQUARED distances, exactly equivalent to the reference's comparisons of
    // float32 square roots (largest x with sqrtf(x) <= (float)limit)
    float eps2_f32, dev2_f32, min_dist2_f32;
    float eps_f32, dev_f32;      // (float)leader_pos_epsilon, (float)max_dev
    float max_distance_f32;      // (float)max_distance for the green-zone walk
    float es_far_f32;            // (float)(max_distance * max_distance_coef)
    float trail_seed_denom_f32;  // (float)(trajectory_saving_period * leader.max_speed)
    float corridor_length_f32;
    float corridor_width_f32;
    float static_inflate[2];     // pre-filter margin for follower / leader static collisions
    // where sensor s writes inside an env's block of `rays`: cell (row j, column q) = base + j * stride + q
    int ray_out_base[FTL_MAX_RAY_SENSORS], ray_out_stride[FTL_MAX_RAY_SENSORS];
    int ray_out_vec4;            // every lasers_count is a multiple of 4 and the layout is the raw one: rows leave as float4
    int ray_out_fused_vec4;      // the same for the fused sensorPrev layout
    int ray_compas_mask;         // bit s: ray sensor s is a LeaderCorridor_lasers_compas (cast by the per-env exact pass)
};

// ---- tiny helpers -----------------------------------------------------------------------------------
FTL_HD double angle_correction(double a) {  // MSC:6-13
    if (a >= 360) return a - 360;
    if (a < 0) return 360 + a;
    return a;
}

FTL_HD double angle_to_point(double cx, double cy, double tx, double ty) {  // MSC:16-26
    double rx = tx - cx, ry = ty - cy, res;
    if (rx > 0)
        res = atan(ry / rx) * kRad2Deg;
    else if (rx < 0)
        res = atan(ry / rx) * kRad2Deg + 180;
    else
        res = 0;
    return angle_correction(res);
}

// int(angle_to_point(...)) as move_to_the_point uses it (CLS:192): only the integer part of the bearing reaches
// the state, so it is taken from a float32 evaluation whenever that is at least 2e-3 degrees away from an integer
// (float32 atan + scaling is good to ~1e-4 degrees), and from the float64 formula otherwise.  Same result as the
// float64 formula in every case, ~5x fewer instructions on the common path.
FTL_HD int bearing_int(double cx, double cy, double tx, double ty) {
    const double rxd = tx - cx, ryd = ty - cy;
    const float rx = (float)rxd, ry = (float)ryd;
    if (fabsf(rx) > 1e-3f) {
        float a = atanf(ry / rx) * 57.29577951308232f;
        if (rx < 0.f) a += 180.f;
        if (a < 0.f) a += 360.f;          // angle_correction; a in (-90, 270) here
        float fl = floorf(a);
        float frac = a - fl;
        if (frac > 2e-3f && frac < 1.f - 2e-3f && a > 2e-3f && a < 360.f - 2e-3f) return (int)fl;
    }
    return (int)angle_to_point(cx, cy, tx, ty);
}

FTL_HD float d2_f32(float ax, float ay, float bx, float by) {  // float32 (a-b)^2 summed, numpy order
    float dx = ax - bx, dy = ay - by;
    float sx = dx * dx, sy = dy * dy;
    return sx + sy;
}
FTL_HD double dist_f64(double ax, double ay, double bx, double by) {
    double dx = ax - bx, dy = ay - by;
    double sx = dx * dx, sy = dy * dy;
    return sqrt(sx + sy);
}

// One shared copy of the float64 sincos / distance code for the hot loop of the step kernel: its frame loop is
// bound by instruction fetch (ncu: no_instruction stalls), so every inlined copy of these ~200-instruction
// sequences costs more than the call.
struct SinCos { double s, c; };
FTL_HD void sincos_deg(double deg, double* s, double* c);
FTL_HD_NOINLINE SinCos sincos_deg_nv(double deg) {
    SinCos r;
    sincos_deg(deg, &r.s, &r.c);
    return r;
}
FTL_HD double dist_f64(double ax, double ay, double bx, double by);
FTL_HD_NOINLINE double dist_f64_nv(double ax, double ay, double bx, double by) { return dist_f64(ax, ay, bx, by); }

FTL_HD void sincos_deg(double deg, double* s, double* c) {
    double th = deg * kDeg2Rad;
#if defined(__CUDA_ARCH__)
    sincos(th, s, c);
#else
    *s = sin(th);
    *c = cos(th);
#endif
}

// ---- robots -------------------------------------------------------------------------------------------
struct Robot {
    float px, py;
    int rx, ry, rw, rh;
    double dir, speed, rot, des_speed, des_rot;
    int rot_dir, des_rot_dir;
};

FTL_HD void robot_load(const DevState& s, int k, int i, Robot& r) {
    float2 p = s.pos[(size_t)k * s.n + i];
    int4 q = s.rect[(size_t)k * s.n + i];
    r.px = p.x; r.py = p.y;
    r.rx = q.x; r.ry = q.y; r.rw = q.z; r.rh = q.w;
    const double* d = s.rd + (size_t)k * RD_COUNT * s.n + i;
    r.dir = d[(size_t)RD_DIR * s.n];
    r.speed = d[(size_t)RD_SPEED * s.n];
    r.rot = d[(size_t)RD_ROT * s.n];
    r.des_speed = d[(size_t)RD_DES_SPEED * s.n];
    r.des_rot = d[(size_t)RD_DES_ROT * s.n];
    int f = s.ri[(size_t)k * s.n + i];
    r.rot_dir = (f & 3) - 1;
    r.des_rot_dir = ((f >> 2) & 3) - 1;
}
FTL_HD void robot_store(const DevState& s, int k, int i, const Robot& r) {
    s.pos[(size_t)k * s.n + i] = make_float2(r.px, r.py);
    s.rect[(size_t)k * s.n + i] = make_int4(r.rx, r.ry, r.rw, r.rh);
    double* d = s.rd + (size_t)k * RD_COUNT * s.n + i;
    d[(size_t)RD_DIR * s.n] = r.dir;
    d[(size_t)RD_SPEED * s.n] = r.speed;
    d[(size_t)RD_ROT * s.n] = r.rot;
    d[(size_t)RD_DES_SPEED * s.n] = r.des_speed;
    d[(size_t)RD_DES_ROT * s.n] = r.des_rot;
    s.ri[(size_t)k * s.n + i] = (r.rot_dir + 1) | ((r.des_rot_dir + 1) << 2);
}

FTL_HD void robot_init(Robot& r, const FtlRobotConfig& c, float x, float y, double dir) {
    r.px = x; r.py = y;
    r.rw = c.width; r.rh = c.height;
    r.rx = (int)x - (c.width >> 1);   // image.get_rect(center=position), CLS:50/56
    r.ry = (int)y - (c.height >> 1);
    r.dir = dir;
    r.speed = r.rot = r.des_speed = r.des_rot = 0.0;
    r.rot_dir = r.des_rot_dir = 0;
}

// pygame 2.1.2 transform.rotate bounding box (third party; see oracle/shims/pygame/transform.py)
FTL_HD void rotated_size(int w, int h, double angle_py, int* ow, int* oh) {
    float angle = (float)angle_py;
    // fmod(angle, 90) == 0.  q only has to be the right candidate when angle IS a multiple of 90 (|q| <= 4: the
    // product with the rounded reciprocal is within 1e-15 of the integer), so no float64 division is needed
    double q = rint((double)angle * (1.0 / 90.0));
    if ((double)angle == 90.0 * q) {
        int turns = ((int)angle / 90) % 4;
        if (turns < 0) turns += 4;
        if (turns & 1) { *ow = h; *oh = w; } else { *ow = w; *oh = h; }
        return;
    }
    {   // float32 evaluation decides unless a bound is within 1e-3 of an integer (sizes are < 100 px: float32 error ~1e-5)
        float snf, csf;
#if defined(__CUDA_ARCH__)
        sincosf(angle * 0.017453292519943295f, &snf, &csf);
#else
        snf = sinf(angle * 0.017453292519943295f); csf = cosf(angle * 0.017453292519943295f);
#endif
        float cxf = fabsf(csf) * (float)w, cyf = fabsf(csf) * (float)h, sxf = fabsf(snf) * (float)w, syf = fabsf(snf) * (float)h;
        float mxf = cxf + syf, myf = sxf + cyf;   // max |cx +- sy| = |cx| + |sy|, same for the other axis
        float fx = mxf - floorf(mxf), fy = myf - floorf(myf);
        if (fx > 1e-3f && fx < 1.f - 1e-3f && fy > 1e-3f && fy < 1.f - 1e-3f) {
            *ow = (int)mxf;
            *oh = (int)myf;
            return;
        }
    }
    double rad = angle * .01745329251994329, sn, cs;
#if defined(__CUDA_ARCH__)
    sincos(rad, &sn, &cs);
#else
    sn = sin(rad); cs = cos(rad);
#endif
    double cx = cs * w, cy = cs * h, sx = sn * w, sy = sn * h;
    double mx = fmax(fmax(fmax(fabs(cx + sy), fabs(cx - sy)), fabs(-cx + sy)), fabs(-cx - sy));
    double my = fmax(fmax(fmax(fabs(sx + cy), fabs(sx - cy)), fabs(-sx + cy)), fabs(-sx - cy));
    *ow = (int)mx;
    *oh = (int)my;
}

FTL_HD void command_turn(Robot& r, const FtlRobotConfig& c, double des, int dir) {  // CLS:109-117
    r.des_rot = (c.max_rotation_speed < des) ? c.max_rotation_speed : des;
    r.des_rot_dir = dir;
}
FTL_HD void command_forward(Robot& r, const FtlRobotConfig& c, double des) {  // CLS:119-127
    if (des > c.max_speed) des = c.max_speed;
    if (des < c.min_speed) des = c.min_speed;
    r.des_speed = des;
}

FTL_HD void robot_move(Robot& r, const FtlRobotConfig& c) {  // CLS:129-182
    double change;
    if (r.rot_dir == 0) r.rot_dir = r.des_rot_dir;
    if (r.rot_dir == r.des_rot_dir) {
        double needed = fabs(r.rot - r.des_rot);
        change = (c.max_rotation_speed_change < needed) ? c.max_rotation_speed_change : needed;
        if (r.des_rot < r.rot) change = -change;
    } else {
        double needed = fabs(r.des_rot + r.rot);
        change = -((c.max_rotation_speed_change < needed) ? c.max_rotation_speed_change : needed);
    }
    double nrs = r.rot + change;
    if (nrs < 0) r.rot_dir = -r.rot_dir;
    r.rot = fabs(nrs);
    double needed = fabs(r.speed - r.des_speed);
    double sc = (needed < c.max_speed_change) ? needed : c.max_speed_change;
    if (r.speed > r.des_speed) sc = -sc;
    r.speed = r.speed + sc;

    if (r.rot != 0) {
        r.dir = angle_correction(r.dir + r.rot_dir * r.rot);
        int cx = r.rx + (r.rw >> 1), cy = r.ry + (r.rh >> 1);
        int nw, nh;
        rotated_size(c.width, c.height, -r.dir, &nw, &nh);
        r.rw = nw; r.rh = nh;
        r.rx = cx - (nw >> 1);
        r.ry = cy - (nh >> 1);
    }
#ifdef FTL_INLINE_SINCOS
    double sn, cs;
    sincos_deg(r.dir, &sn, &cs);
#else
    double sn, cs;
    sincos_deg(r.dir, &sn, &cs);   // inline: the kinematics pass fits the instruction cache (measured -0.5 % per step)
#endif
    float mx = (float)(cs * r.speed), my = (float)(sn * r.speed);
    r.px = r.px + mx;
    r.py = r.py + my;
    double dx = (double)r.px - (double)(r.rx + (r.rw >> 1));
    double dy = (double)r.py - (double)(r.ry + (r.rh >> 1));
    r.rx += (int)dx;  // move_ip truncates toward zero; (0, 0) is a no-op
    r.ry += (int)dy;
}

FTL_HD void move_to_the_point(Robot& r, const FtlRobotConfig& c, double tx, double ty, bool has_speed,
                              double speed) {  // CLS:184-215
    double new_speed = has_speed ? speed : dist_f64_nv((double)r.px, (double)r.py, tx, ty);
    int desirable_angle = bearing_int((double)r.px, (double)r.py, tx, ty);
    int cur = (int)r.dir;
    int delta, nrd;
    if (desirable_angle - cur > 0) {
        if (desirable_angle - cur > 180) { delta = cur + (360 - desirable_angle); nrd = -1; }
        else { delta = desirable_angle - cur; nrd = 1; }
    } else {
        if (cur - desirable_angle > 180) { nrd = 1; delta = (360 - cur) + desirable_angle; }
        else { nrd = -1; delta = cur - desirable_angle; }
    }
    command_turn(r, c, (double)delta, nrd);
    command_forward(r, c, new_speed);
    robot_move(r, c);
}

// Out-of-line versions for the step kernel: the robot code (two float64 sincos, one atan, the rotated bounding box)
// is ~1.5k instructions; inlined at every call site the frame loop overflows the instruction caches.  Robots
// travel by value (registers), so nothing is forced into local memory.
FTL_HD_NOINLINE Robot robot_move_nv(Robot r, const FtlRobotConfig* c) {
    robot_move(r, *c);
    return r;
}
FTL_HD_NOINLINE Robot move_to_the_point_nv(Robot r, const FtlRobotConfig* c, double tx, double ty, int has_speed,
                                           double speed) {
    double new_speed = has_speed ? speed : dist_f64_nv((double)r.px, (double)r.py, tx, ty);
    int desirable_angle = bearing_int((double)r.px, (double)r.py, tx, ty);
    int cur = (int)r.dir;
    int delta, nrd;
    if (desirable_angle - cur > 0) {
        if (desirable_angle - cur > 180) { delta = cur + (360 - desirable_angle); nrd = -1; }
        else { delta = desirable_angle - cur; nrd = 1; }
    } else {
        if (cur - desirable_angle > 180) { nrd = 1; delta = (360 - cur) + desirable_angle; }
        else { nrd = -1; delta = cur - desirable_angle; }
    }
    command_turn(r, *c, (double)delta, nrd);
    command_forward(r, *c, new_speed);
    return robot_move_nv(r, c);
}

// ---- collisions -----------------------------------------------------------------------------------------
FTL_HD bool rects_collide(int ax, int ay, int aw, int ah, int bx, int by, int bw, int bh) {  // pygame colliderect
    if (aw == 0 || ah == 0 || bw == 0 || bh == 0) return false;
    return ax < bx + bw && ay < by + bh && ax + aw > bx && ay + ah > by;
}
FTL_HD bool robots_collide(const Robot& a, const Robot& b) {
    return rects_collide(a.rx, a.ry, a.rw, a.rh, b.rx, b.ry, b.rw, b.rh);
}
FTL_HD bool out_of_bounds(const FtlConfig& c, const Robot& r) {  // ENV:1182-1183
    return r.px > (float)c.game_width || r.py > (float)c.game_height || r.px < 0 || r.py < 0;
}

// Static rectangles that the robot could possibly touch during one step: those whose box, grown by
// `inflate` (half the largest hit-box diagonal + the distance covered in F frames + slack), contains
// the robot's position at the start of the step.  Exact: the per-frame test still uses the integer
// rectangles, the mask only skips rectangles that are provably out of reach.
FTL_HD bool near_rect(int4 q, float px, float py, float inflate) {
    return px >= (float)q.x - inflate && px <= (float)(q.x + q.z) + inflate && py >= (float)q.y - inflate &&
           py <= (float)(q.y + q.w) + inflate;
}
// one pass over the scenario's rectangles for both robots (the table is read once per step)
FTL_HD void near_static_masks(const int4* rects, int n_static, float2 p0, float inflate0, float2 p1, float inflate1,
                              uint64_t* m0, uint64_t* m1) {
    uint64_t a = 0, b = 0;
#pragma unroll 4
    for (int k = 0; k < n_static; k++) {
        int4 q = rects[k];
        if (near_rect(q, p0.x, p0.y, inflate0)) a |= (uint64_t)1 << k;
        if (near_rect(q, p1.x, p1.y, inflate1)) b |= (uint64_t)1 << k;
    }
    *m0 = a;
    *m1 = b;
}
FTL_HD int first_bit64(uint64_t m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}
// the same masks from the scenario's cell grid: only the few rectangles of the robot's cell are examined.  Positions
// outside the field (possible after a crash without auto-reset) take the full pass.
FTL_HD bool near_grid_cell(const DevPool& pool, float2 p, int* cell) {
    if (!(p.x >= 0.f && p.y >= 0.f)) return false;
    const int cx = (int)p.x >> kNearGridShift, cy = (int)p.y >> kNearGridShift;
    if (cx >= pool.near_grid_w || cy >= pool.near_grid_h) return false;
    *cell = cy * pool.near_grid_w + cx;
    return true;
}
FTL_HD uint64_t near_refine(const int4* rects, uint64_t cand, float2 p, float inflate) {
    uint64_t m = 0;
    while (cand) {
        const int k = first_bit64(cand);
        cand &= cand - 1;
        if (near_rect(rects[k], p.x, p.y, inflate)) m |= (uint64_t)1 << k;
    }
    return m;
}
FTL_HD void near_static_masks_grid(const DevPool& pool, int scenario, const int4* rects, int n_static, float2 p0,
                                   float inflate0, float2 p1, float inflate1, uint64_t* m0, uint64_t* m1) {
    int c0, c1;
#ifdef FTL_NO_NEAR_GRID
    if (false)
#else
    if (pool.near_grid && near_grid_cell(pool, p0, &c0) && near_grid_cell(pool, p1, &c1))
#endif
    {
        const uint64_t* g = pool.near_grid + (size_t)scenario * pool.near_grid_w * pool.near_grid_h;
        const uint64_t g0 = g[c0], g1 = g[c1];
        *m0 = near_refine(rects, g0, p0, inflate0);
        *m1 = near_refine(rects, g1, p1, inflate1);
        return;
    }
    near_static_masks(rects, n_static, p0, inflate0, p1, inflate1, m0, m1);
}
// one robot's mask (the role-split kinematics kernel computes the follower's and the leader's in different warps)
FTL_HD uint64_t near_static_mask_one(const DevPool& pool, int scenario, const int4* rects, int n_static, float2 p,
                                     float inflate) {
    int cell;
    if (pool.near_grid && near_grid_cell(pool, p, &cell)) {
        const uint64_t* g = pool.near_grid + (size_t)scenario * pool.near_grid_w * pool.near_grid_h;
        return near_refine(rects, g[cell], p, inflate);
    }
    uint64_t a = 0;
    for (int k = 0; k < n_static; k++)
        if (near_rect(rects[k], p.x, p.y, inflate)) a |= (uint64_t)1 << k;
    return a;
}
FTL_HD bool collide_static_masked(const Robot& r, const int4* rects, uint64_t mask) {
    while (mask) {
#if defined(__CUDA_ARCH__)
        int k = __ffsll((long long)mask) - 1;
#else
        int k = __builtin_ctzll(mask);
#endif
        mask &= mask - 1;
        int4 q = rects[k];
        if (rects_collide(r.rx, r.ry, r.rw, r.rh, q.x, q.y, q.z, q.w)) return true;
    }
    return false;
}

// ---- green zone: exact flags from cached bounds ------------------------------------------------------------
// The reference recomputes, every frame, (1) the "green" suffix of the leader trail whose arc length
// from the newest point stays <= max_distance and (2) the nearest green point to the follower, and,
// when that is farther than max_dev, (3) the nearest point of the WHOLE trail (ENV:1828-1843,
// 1906-1937; half of the reference's run time).  The flags only depend on whether those minima are
// below two thresholds, so this implementation keeps
//   * g_lo: first green index, recomputed (same float32 accumulation order) only when the trail grows;
//   * for each of the two searches a witness index (a point known to be in the set: its exact distance
//     is an upper bound of the minimum) and a lower bound of the minimum that is decreased every frame
//     by the follower's displacement (triangle inequality) and folded with each newly added point.
// A full scan (bit-identical arithmetic to the reference) is only done when the bounds cannot decide a
// comparison, so the flags are always exactly those of the reference.
struct GreenCache {
    int g_lo;      // first trail index that is CERTAINLY green; green = [g_lo - (0..g_unc), trail_len-2]
    int g_unc;     // how many points just below g_lo could not be decided without the exact float32 walk
    int a_star;    // witness for the green minimum (-1 = none)
    int b_star;    // witness for the whole-trail minimum (-1 = none)
    float lb_g;    // lower bound of the distance from (sgx, sgy) to every green point (negative = unknown)
    float lb_all;  // lower bound of the distance from (sax, say) to every trail point
    float sgx, sgy, sax, say;  // follower positions at the last full scans
    int nice_from;  // all trail_d[k], k >= nice_from, are multiples of 2^-15 (exact float32 sums)
    // register copies, not persisted: coordinates of the two witnesses and the newest trail point / length
    float2 a_pt, b_pt, last_pt;
    double last_s;
};

constexpr float kBoundSlack = 2e-3f;  // covers float32 rounding of positions/distances up to ~4000 px

// _trajectory_in_box, ENV:1828-1843, literally: float32 accumulation of the stored segment lengths from
// the newest point backwards.  trail_d[k] = float32 |trail[k] - trail[k-1]|.
FTL_HD int green_lo_exact(const float* trail_d, int n, float max_distance_f32) {
    int lo = n - 1;  // empty
    float acc = 0.f;
    FTL_COUNT(2, 1);
#if defined(__CUDA_ARCH__) && !defined(FTL_NO_WALK_PREFETCH)
    // the walk below is a chain of ~20 dependent batches of loads over ~640 bytes: ask for all of its cache lines first,
    // so that only the first batch pays the DRAM round trip
    for (int k = n - 1; k > 0 && k > n - 1 - 256; k -= 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(trail_d + k));
#endif
    // same additions in the same order; the loads are issued eight at a time so that the walk is not one
    // memory round trip per point
    for (int i = n - 2; i >= 0; i -= 8) {
        float d[8];
#pragma unroll
        for (int j = 0; j < 8; j++) d[j] = trail_d[i + 1 - j > 1 ? i + 1 - j : 1];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (j > i) return lo;
            acc = acc + d[j];
            if (acc <= max_distance_f32)
                lo = i - j;
            else
                return lo;
        }
    }
    return lo;
}

// The same index without the walk: the float32 running sum differs from the float64 cumulative length by at
// most (#terms) * 0.5 ulp, so membership is decided from trail_s unless a point is that close to the threshold.
// Such points (the leader covers 1.25 px per saved point and max_distance is 160 of those, so near-ties are the
// rule on straight stretches) are kept as "uncertain" and only resolved with the exact walk if the flags ever
// depend on them.  The window start only moves forward (float addition is monotone).
FTL_HD void green_window_update(const double* trail_s, int n, double total, float max_distance_f32, int nice_from,
                                int* g_lo, int* g_unc) {   // total = trail_s[n-1]
    if (n < 2) { *g_lo = n - 1; *g_unc = 0; return; }
    const double maxd = (double)max_distance_f32;
    int lo = *g_lo - *g_unc;
    if (lo < 0) lo = 0;
    int first_unc = -1;
    for (; lo <= n - 2; lo++) {
        double len = total - trail_s[lo];
        double err = (double)(n - 1 - lo) * maxd * 1.2e-7 + 1e-9;
        if (lo + 1 >= nice_from && maxd < 512.0) {
            // every term of this window is a multiple of 2^-15 and the sum stays below 2^9: the float32
            // running sum is exact, so the float64 length (snapped to the grid) decides, ties included
            len = rint(len * 32768.0) * (1.0 / 32768.0);
            err = 0.0;
        }
        if (len > maxd + err) continue;                                   // certainly outside
        if (len > maxd - err) { if (first_unc < 0) first_unc = lo; continue; }  // too close to call
        break;                                                            // certainly inside
    }
    *g_lo = lo;   // n-1 when nothing is certainly inside
    *g_unc = first_unc < 0 ? 0 : lo - first_unc;
}

FTL_HD float scan_min_d2(const float2* trail, int lo, int hi, float fx, float fy, int* arg) {
    // min over [lo, hi] of the reference's float32 squared distance
    float best = 3.0e38f;
    int bi = -1;
    FTL_COUNT(0, 1);
    FTL_COUNT(1, hi - lo + 1);
    for (int i = hi; i >= lo; i--) {
        float2 p = trail[i];
        float d2 = d2_f32(p.x, p.y, fx, fy);
        if (d2 < best) { best = d2; bi = i; }
    }
    *arg = bi;
    return best;
}

// 1: min <= thr, 0: min > thr, -1: the bounds cannot tell.  ub2: squared distance to a witness inside the
// set; lb: lower bound of the distances from the scan position; disp2: squared displacement since that scan.
FTL_HD int bound_decide(float ub2, float lb, float disp2, float thr, float thr2) {
    if (ub2 <= thr2) return 1;
    float m = lb - thr - kBoundSlack;
    if (m > 0.f && m * m > disp2 * 1.0001f) return 0;
    return -1;
}

// Full scans are requested, not performed, by the per-thread logic: on the GPU the 32 lanes of a warp serve
// each request together (coalesced loads, ~60 warp instructions instead of a ~200-iteration serial loop in one
// lane while 31 wait); the host build runs the same arithmetic serially.
struct ScanMin { float best; int arg; };
#ifndef FTL_SCAN_UNROLL
#define FTL_SCAN_UNROLL 4
#endif
#if !defined(FTL_SCAN_WIDE) && !defined(FTL_SCAN_NARROW)
#define FTL_SCAN_WIDE 8   // measured (profiles/r02_ab_log.txt): k_kin 0.164 -> 0.158 ms
#endif
#ifndef FTL_OUTLINE_SCAN   // measured: out of line is 2% slower (k_step 0.2045 -> 0.2091 ms)
FTL_HD ScanMin warp_scan_min_impl(bool need, const float2* trail, int lo, int hi, float fx, float fy) {
#else
FTL_HD_NOINLINE ScanMin warp_scan_min_impl(bool need, const float2* trail, int lo, int hi, float fx, float fy) {
#endif
    ScanMin r;
    int* arg = &r.arg;
#if defined(__CUDA_ARCH__)
    const unsigned full = 0xffffffffu;
    const int lane = (int)(threadIdx.x & 31);
    unsigned pending = __ballot_sync(full, need);
    float my_best = 3.0e38f;
    int my_arg = -1;
    while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1;
        const float2* t = (const float2*)(uintptr_t)__shfl_sync(full, (unsigned long long)(uintptr_t)trail, src);
        const int blo = __shfl_sync(full, lo, src), bhi = __shfl_sync(full, hi, src);
        const float bfx = __shfl_sync(full, fx, src), bfy = __shfl_sync(full, fy, src);
        float best = 3.0e38f;
        int bi = -1;
        int k = bhi - lane;
#if defined(FTL_SCAN_WIDE)
        // FTL_SCAN_WIDE loads in flight per lane, predicated, so that a scan of up to 32 * FTL_SCAN_WIDE points is ONE
        // memory round trip (the unrolled-then-serial loop below pays one per remainder iteration); same visiting order
        for (; k >= blo; k -= 32 * FTL_SCAN_WIDE) {
            float2 p[FTL_SCAN_WIDE];
#pragma unroll
            for (int u = 0; u < FTL_SCAN_WIDE; u++) {
                const int kk = k - 32 * u;
                p[u] = t[kk >= blo ? kk : k];
            }
#pragma unroll
            for (int u = 0; u < FTL_SCAN_WIDE; u++) {
                const int kk = k - 32 * u;
                float d2 = d2_f32(p[u].x, p[u].y, bfx, bfy);
                if (kk >= blo && d2 < best) { best = d2; bi = kk; }
            }
        }
#else
#if FTL_SCAN_UNROLL > 1
        // several loads in flight per lane (the scan is a chain of DRAM round trips otherwise); same visiting order
        for (; k - 32 * (FTL_SCAN_UNROLL - 1) >= blo; k -= 32 * FTL_SCAN_UNROLL) {
            float2 p[FTL_SCAN_UNROLL];
#pragma unroll
            for (int u = 0; u < FTL_SCAN_UNROLL; u++) p[u] = t[k - 32 * u];
#pragma unroll
            for (int u = 0; u < FTL_SCAN_UNROLL; u++) {
                float d2 = d2_f32(p[u].x, p[u].y, bfx, bfy);
                if (d2 < best) { best = d2; bi = k - 32 * u; }
            }
        }
#endif
        for (; k >= blo; k -= 32) {
            float2 p = t[k];
            float d2 = d2_f32(p.x, p.y, bfx, bfy);
            if (d2 < best) { best = d2; bi = k; }
        }
#endif
#ifndef FTL_SCAN_SHUFFLE   // arg-min in two hardware reductions (squared distances are non-negative floats: they order like
                           // their bits; among equal minima the largest index wins, as in the shuffle tree below); measured
                           // (r02_ab_log.txt (21)): k_kin 0.1342 -> 0.1328 ms
        {
            const unsigned mbits = __reduce_min_sync(full, __float_as_uint(best));
            const int widx = __reduce_max_sync(full, __float_as_uint(best) == mbits ? bi : -1);
            best = __uint_as_float(mbits); bi = widx;
        }
#else
#pragma unroll
        for (int off = 16; off; off >>= 1) {
            float ob = __shfl_xor_sync(full, best, off);
            int oi = __shfl_xor_sync(full, bi, off);
            if (ob < best || (ob == best && oi > bi)) { best = ob; bi = oi; }
        }
#endif
        if (lane == src) { my_best = best; my_arg = bi; }
    }
    *arg = my_arg;
    r.best = my_best;
    return r;
#else
    if (!need) { r.arg = -1; r.best = 3.0e38f; return r; }
    r.best = scan_min_d2(trail, lo, hi, fx, fy, arg);
    return r;
#endif
}
// the common case -- no lane of the warp needs a scan -- is decided by one ballot
FTL_HD float warp_scan_min(bool need, const float2* trail, int lo, int hi, float fx, float fy, int* arg) {
#if defined(__CUDA_ARCH__)
    if (__ballot_sync(0xffffffffu, need) == 0u) { *arg = -1; return 3.0e38f; }
#endif
    const ScanMin r = warp_scan_min_impl(need, trail, lo, hi, fx, fy);
    *arg = r.arg;
    return r.best;
}

#ifndef FTL_OUTLINE_SCAN   // measured: out of line is 2% slower (k_step 0.2045 -> 0.2091 ms)
FTL_HD int green_lo_exact_nv(const float* trail_d, int n, float max_distance_f32) {
#else
FTL_HD_NOINLINE int green_lo_exact_nv(const float* trail_d, int n, float max_distance_f32) {
#endif
    return green_lo_exact(trail_d, n, max_distance_f32);
}
FTL_HD void green_resolve(const DevCfg& cfg, const float* trail_d, int n, GreenCache& gc) {
    gc.g_lo = green_lo_exact_nv(trail_d, n, cfg.max_distance_f32);
    gc.g_unc = 0;
}
// The same walk as a warp collective (every lane of the warp calls it; `need` = this lane wants its window resolved).
// In one lane the walk is a chain of ~20 memory round trips (eight loads each); here the warp loads the newest
// 32 * kWalkWide segment lengths of the requesting env at once, coalesced, and every lane replays the reference's
// additions in the reference's order from shuffled values -- same float32 sums, one round trip per 192 points.
#ifndef FTL_WALK_WIDE
#define FTL_WALK_WIDE 6
#endif
FTL_HD void warp_green_resolve(bool need, const DevCfg& cfg, const float* trail_d, int n, GreenCache& gc) {
#if defined(__CUDA_ARCH__) && !defined(FTL_SERIAL_WALK)   // measured (r02_ab_log.txt (16)): k_kin 0.1449 -> 0.1349 ms
    const unsigned full = 0xffffffffu;
    unsigned pending = __ballot_sync(full, need);
    if (pending == 0u) return;
    const int lane = (int)(threadIdx.x & 31);
    const float maxd = cfg.max_distance_f32;
    while (pending) {
        const int src = __ffs((int)pending) - 1;
        pending &= pending - 1;
        const float* td = (const float*)(uintptr_t)__shfl_sync(full, (unsigned long long)(uintptr_t)trail_d, src);
        const int bn = __shfl_sync(full, n, src);
        int lo = bn - 1;     // green_lo_exact: term t (t = 0, 1, ...) is trail_d[bn - 1 - t], accepted index bn - 2 - t
        float acc = 0.f;
        bool done = bn < 2;
        for (int t0 = 0; !done && t0 <= bn - 2; t0 += 32 * FTL_WALK_WIDE) {
            float d[FTL_WALK_WIDE];
#pragma unroll
            for (int u = 0; u < FTL_WALK_WIDE; u++) {
                const int idx = bn - 1 - (t0 + 32 * u + lane);
                d[u] = td[idx > 1 ? idx : 1];
            }
#pragma unroll
            for (int u = 0; u < FTL_WALK_WIDE; u++) {
                if (done) break;
                for (int l = 0; l < 32; l++) {
                    const float v = __shfl_sync(full, d[u], l);
                    const int t = t0 + 32 * u + l;
                    if (t > bn - 2) { done = true; break; }
                    acc = acc + v;
                    if (acc <= maxd) lo = bn - 2 - t;
                    else { done = true; break; }
                }
            }
        }
        if (lane == src) { gc.g_lo = lo; gc.g_unc = 0; }
    }
#else
    if (need) green_resolve(cfg, trail_d, n, gc);
#endif
}

// `active` = false: the lane has no frame to evaluate (fewer frames this step than its warp's longest env) and only
// takes part in the warp's collectives.
FTL_HD void green_flags(const DevCfg& cfg, const float2* trail, const float* trail_d, int n, float fx, float fy,
                        GreenCache& gc, bool* in_box, bool* on_trace, bool active = true) {  // ENV:1906-1931; all lanes of a warp call this
    FTL_COUNT(3, 1);
    const int hi = n - 2;
    // len(green) > 2 ?  (only ambiguous in the first frames of an episode)
    warp_green_resolve(active && gc.g_unc > 0 && hi - gc.g_lo + 1 <= 2 && hi - gc.g_lo + 1 + gc.g_unc > 2, cfg, trail_d, n, gc);
    const bool have_green = active && hi - gc.g_lo + 1 > 2;
    int le_eps = 0, le_dev = 0;
    bool need = false;
    if (have_green) {
        if (gc.a_star < gc.g_lo) {  // the witness left the window: its oldest certain point is the next guess
            gc.a_star = gc.g_lo;
            gc.a_pt = trail[gc.a_star];
        }
        float ub2 = 3.0e38f;
        if (gc.a_star <= hi) ub2 = d2_f32(gc.a_pt.x, gc.a_pt.y, fx, fy);
        float disp2 = d2_f32(fx, fy, gc.sgx, gc.sgy);
        le_eps = bound_decide(ub2, gc.lb_g, disp2, cfg.eps_f32, cfg.eps2_f32);
        le_dev = le_eps == 1 ? 1 : bound_decide(ub2, gc.lb_g, disp2, cfg.dev_f32, cfg.dev2_f32);
        need = le_eps < 0 || (le_eps == 0 && le_dev < 0);
#ifdef FTL_TIMING_NOSCAN   // timing experiment only (results differ): what the exact scans cost
        if (need) { need = false; le_eps = le_eps < 0 ? 0 : le_eps; le_dev = le_dev < 0 ? 0 : le_dev; }
#endif
    }
    {
        int arg;
        float m = warp_scan_min(need, trail, gc.g_lo, hi, fx, fy, &arg);
        float m_low = m;   // lower bound material: certain and uncertain points together
        if (need) {
            for (int u = gc.g_lo - gc.g_unc; u < gc.g_lo; u++) {
                float2 p = trail[u];
                m_low = fminf(m_low, d2_f32(p.x, p.y, fx, fy));
            }
        }
        const int old_lo = gc.g_lo;
        const bool resolve = need && m_low < m;   // an undecided point would be the nearest one: now its membership matters
        warp_green_resolve(resolve, cfg, trail_d, n, gc);
        if (need) {
            if (resolve) {
                for (int u = gc.g_lo; u < old_lo; u++) {
                    float2 p = trail[u];
                    float d2 = d2_f32(p.x, p.y, fx, fy);
                    if (d2 < m) { m = d2; arg = u; }
                }
                m_low = m;
            }
            gc.a_star = arg;
            gc.a_pt = trail[arg];
            gc.lb_g = sqrtf(m_low) - kBoundSlack;
            gc.sgx = fx; gc.sgy = fy;
            le_eps = m <= cfg.eps2_f32;
            le_dev = m <= cfg.dev2_f32;
        }
    }
    // whole-trail fallback, ENV:1924-1931
    const bool fallback = have_green && le_eps != 1 && le_dev != 1;
    int le = 0;
    need = false;
    if (fallback) {
        float ub2a = 3.0e38f;
        if (gc.b_star >= 0 && gc.b_star < n) ub2a = d2_f32(gc.b_pt.x, gc.b_pt.y, fx, fy);
        float disp2a = d2_f32(fx, fy, gc.sax, gc.say);
        le = bound_decide(ub2a, gc.lb_all, disp2a, cfg.eps_f32, cfg.eps2_f32);
        FTL_COUNT(4, 1);
        need = le < 0;
#ifdef FTL_TIMING_NOSCAN
        if (need) { need = false; le = 0; }
#endif
    }
    {
        int arg;
        float m = warp_scan_min(need, trail, 0, n - 1, fx, fy, &arg);
        if (need) {
            FTL_COUNT(5, 1);
            FTL_COUNT(6, n);
            gc.b_star = arg;
            gc.b_pt = trail[arg];
            gc.lb_all = sqrtf(m) - kBoundSlack;
            gc.sax = fx; gc.say = fy;
            le = m <= cfg.eps2_f32;
        }
    }
    *in_box = have_green && (le_eps == 1 || le_dev == 1);
    *on_trace = have_green && (le_eps == 1 || (fallback && le == 1));
}

// the trail grew: trail[n-1] is new (n = new length); trail[n-2] enters the green set.  New points are folded
// into the bounds as distances from the respective scan positions.
// the trail grew: `p` was appended at index n-1 (n = new length), the previous newest point `q` (index n-2)
// enters the green set.  New points are folded into the bounds as distances from the respective scan positions.
FTL_HD void green_cache_appended(const DevCfg& cfg, const double* trail_s, int n, float2 p, float2 q, float d_new,
                                 GreenCache& gc) {
    if (d_new * 32768.f != rintf(d_new * 32768.f)) gc.nice_from = n;  // term n-1 is not on the 2^-15 grid
    green_window_update(trail_s, n, gc.last_s, cfg.max_distance_f32, gc.nice_from, &gc.g_lo, &gc.g_unc);
    float d = sqrtf(d2_f32(p.x, p.y, gc.sax, gc.say)) - kBoundSlack;
    if (d < gc.lb_all) gc.lb_all = d;
    if (n >= 2) {
        float dq = sqrtf(d2_f32(q.x, q.y, gc.sgx, gc.sgy)) - kBoundSlack;
        if (dq < gc.lb_g) gc.lb_g = dq;
    }
}
FTL_HD void green_cache_invalidate(const DevCfg& cfg, const float* trail_d, int n, GreenCache& gc) {
    gc.g_lo = green_lo_exact(trail_d, n, cfg.max_distance_f32);
    gc.g_unc = 0;
    gc.nice_from = n;  // conservative: nothing known about the existing terms
    gc.a_star = -1;
    gc.b_star = -1;
    gc.lb_g = -1.f;
    gc.lb_all = -1.f;
    gc.sgx = gc.sgy = gc.sax = gc.say = 0.f;
    gc.a_pt = gc.b_pt = make_float2(0.f, 0.f);
}

// register copies of the witnesses and of the trail tail (after a cache_load)
FTL_HD void green_cache_hydrate(const float2* trail, const double* trail_s, int n, GreenCache& gc) {
    gc.a_pt = (gc.a_star >= 0 && gc.a_star < n) ? trail[gc.a_star] : make_float2(0.f, 0.f);
    gc.b_pt = (gc.b_star >= 0 && gc.b_star < n) ? trail[gc.b_star] : make_float2(0.f, 0.f);
    gc.last_pt = n > 0 ? trail[n - 1] : make_float2(0.f, 0.f);
    gc.last_s = n > 0 ? trail_s[n - 1] : 0.0;
}

// append one point to the trail and its derived arrays
FTL_HD void trail_push(float2* trail, float* trail_d, double* trail_s, int k, float x, float y) {
    trail[k] = make_float2(x, y);
    if (k == 0) {
        trail_d[0] = 0.f;
        trail_s[0] = 0.0;
    } else {
        float2 prev = trail[k - 1];
        float d = sqrtf(d2_f32(x, y, prev.x, prev.y));  // euclidean(newer, older) in float32, ENV:1838
        trail_d[k] = d;
        trail_s[k] = trail_s[k - 1] + (double)d;
    }
}

// ---- Philox4x32-10 for list-valued speed regimes (same keying as the oracle) --------------------------------
FTL_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
FTL_HD double regime_uniform(int64_t env_global, int episode, int frame) {
    uint32_t c0 = (uint32_t)frame, c1 = (uint32_t)episode, c2 = 0x46544c31u, c3 = 0u;
    uint32_t k0 = (uint32_t)env_global, k1 = (uint32_t)((uint64_t)env_global >> 32);
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return (double)c0 * (1.0 / 4294967296.0);
}

// ---- episode scalars kept in registers during a step --------------------------------------------------------
struct Episode {
    int step_count, cur_target_id, finish_timer, flags, trail_len, scenario, episode, overflow, accel_consumed;
    double acc_penalty, overall, last_reward, speed_mult, lead_acc, lead_cum;
};

FTL_HD double reward_of(const FtlConfig& c, const Episode& e, bool too_close, bool in_box, bool on_trace) {  // ENV:1869-1904
    // (written as an initialisation, not as 0 + x: ptxas 12.9 -O3 dropped the zero initialiser of the
    // accumulator in the NB=0 instantiation; tests/test_gpu_parity.py's no-bear golden trace guards this)
    double r = c.leader_movement_reward;
    if (too_close) {
        r += c.too_close_penalty;
    } else {
        if (in_box && on_trace) r += c.reward_in_box;
        else if (in_box) r += c.reward_in_dev;
        else if (on_trace) r += c.reward_on_track;
        else if (e.step_count > c.warm_start) r += c.not_on_track_penalty;
    }
    if (e.flags & FL_CRASH) r += c.crash_penalty;
    return r;
}

FTL_HD double leader_speed(const DevCfg& cfg, Episode& e, int env_index, const double* draw) {  // ENV:1143-1157
    const FtlConfig& c = cfg.c;
    int sel = -1;
    for (int k = 0; k < c.n_speed_regime; k++)
        if (c.speed_regime_key[k] <= e.step_count) sel = k;
    if (sel >= 0) {
        if (c.speed_regime_is_range[sel]) {
            // random.uniform(a, b) = a + (b - a) * random(): u is the caller's recorded draw, or Philox
            double u = draw ? *draw : regime_uniform(cfg.env_id_base + env_index, e.episode, e.step_count);
            e.speed_mult = c.speed_regime_lo[sel] + (c.speed_regime_hi[sel] - c.speed_regime_lo[sel]) * u;
        } else {
            e.speed_mult = c.speed_regime_lo[sel];
        }
    }
    return c.leader.max_speed * e.speed_mult;
}
FTL_HD double leader_accel(const DevCfg& cfg, Episode& e) {  // ENV:1159-1174
    const FtlConfig& c = cfg.c;
    for (int k = 0; k < c.n_accel_regime; k++) {
        if (e.accel_consumed & (1 << k)) continue;
        if (c.accel_regime_key[k] <= e.step_count) {
            e.lead_acc = c.accel_regime_val[k];
            e.lead_cum = e.lead_acc;
            e.accel_consumed |= (1 << k);
        }
    }
    e.lead_cum += e.lead_acc;
    return e.lead_cum * c.leader.max_speed;
}

// bear target, ENV:722-758 and 819-837
FTL_HD void bear_target(const FtlConfig& c, int idx, const Robot& bear, const Robot& leader, double* tx, double* ty,
                        int* index) {
    double d = dist_f64_nv((double)bear.px, (double)bear.py, *tx, *ty);
    double radius, ang;
    if (c.move_bear_v4 && (idx & 1)) {
        if (d < c.leader_pos_epsilon) *index += 1;
        if (*index > 3) *index = 0;
        // order[idx][k] -> which of p1..p4: idx1 [p4,p3,p1,p2], idx3 [p3,p1,p2,p4]
        int which;
        if (idx == 1) which = (*index == 0) ? 3 : (*index == 1) ? 2 : (*index == 2) ? 0 : 1;
        else which = (*index == 0) ? 2 : (*index == 1) ? 0 : (*index == 2) ? 1 : 3;
        radius = which < 2 ? 150.0 : 250.0;
        ang = which == 0 ? leader.dir + 140 : which == 1 ? leader.dir - 140 : which == 2 ? leader.dir - 160 : leader.dir + 160;
    } else {
        if (d < c.leader_pos_epsilon) {
            *index += 1;
            if (*index > 1) *index = 0;
        }
        radius = 100.0 * (idx + 1);
        ang = *index == 0 ? leader.dir - 130 : leader.dir + 130;
    }
    const SinCos sc_ = sincos_deg_nv(ang);
    *tx = (double)leader.px + sc_.c * radius;
    *ty = (double)leader.py + sc_.s * radius;
}

// ---- tracker (LeaderPositionsTracker_v2.scan, SEN:243-327) ---------------------------------------------------
struct Tracker {
    int saving_counter, ring_tail, ring_head, hist_f64_end;
};

// numpy add.reduce (pairwise_sum) order, generated on the fly over k = 0..n-1 of f(k)
template <typename T, typename F>
FTL_HD T np_pairwise_sum(int n, int base, F f) {
    if (n < 8) {
        T r = 0;
        for (int i = 0; i < n; i++) r += f(base + i);
        return r;
    }
    T r0 = f(base + 0), r1 = f(base + 1), r2 = f(base + 2), r3 = f(base + 3), r4 = f(base + 4), r5 = f(base + 5),
      r6 = f(base + 6), r7 = f(base + 7);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += f(base + i); r1 += f(base + i + 1); r2 += f(base + i + 2); r3 += f(base + i + 3);
        r4 += f(base + i + 4); r5 += f(base + i + 5); r6 += f(base + i + 6); r7 += f(base + i + 7);
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
    for (; i < n; i++) res += f(base + i);
    return res;
}
template <typename T, typename F>
FTL_HD T np_sum(int n, F f) {  // n <= 512 here (corridor_cap): at most two levels of blocking
    if (n <= 128) return np_pairwise_sum<T>(n, 0, f);
    int n2 = n / 2;
    n2 -= n2 % 8;
    T a, b;
    if (n2 <= 128) a = np_pairwise_sum<T>(n2, 0, f);
    else { int m = n2 / 2; m -= m % 8; a = np_pairwise_sum<T>(m, 0, f) + np_pairwise_sum<T>(n2 - m, m, f); }
    int nr = n - n2;
    if (nr <= 128) b = np_pairwise_sum<T>(nr, n2, f);
    else { int m = nr / 2; m -= m % 8; b = np_pairwise_sum<T>(m, n2, f) + np_pairwise_sum<T>(nr - m, n2 + m, f); }
    return a + b;
}

// The length test of SEN:283-287 re-measures every segment of the history on every save (and once more per dropped
// point); a segment's length depends on its two end points only, so both forms the reference can ask for (float64
// while a seeded float64 point is still in the list, float32 afterwards) are stored when the segment is created and
// the test only sums them, in numpy's order.  Segment k joins ring entries k and k+1 and lives in slot k.
FTL_HD void tracker_seg_store(const double2* hist, double* seg_d, float* seg_f, int mask, int k) {
    const double2 a = hist[k & mask], b = hist[(k + 1) & mask];
    seg_d[k & mask] = dist_f64(a.x, a.y, b.x, b.y);
    seg_f[k & mask] = sqrtf(d2_f32((float)a.x, (float)a.y, (float)b.x, (float)b.y));
}

#ifndef FTL_INLINE_NPSUM   // measured (profiles/r02_ab_log.txt): k_kin 0.1585 -> 0.1538 ms
// Compact form of the same sums: inlined in the tracker's trim loop, np_sum's six blocked cases (each unrolled by the
// compiler), twice, are 145 KB of the step kernel's 300 KB of code -- for a function that runs two or three times per
// step.  One rolled, out-of-line block sum per type instead; np_sum's splitting is kept literally.
template <typename T>
FTL_HD_NOINLINE T ring_block_sum(const T* a, int start, int mask, int n) {   // numpy pairwise_sum for n <= 128, ring-indexed
    if (n < 8) {
        T r = 0;
#pragma unroll 1
        for (int i = 0; i < n; i++) r += a[(start + i) & mask];
        return r;
    }
    T r0 = a[(start + 0) & mask], r1 = a[(start + 1) & mask], r2 = a[(start + 2) & mask], r3 = a[(start + 3) & mask],
      r4 = a[(start + 4) & mask], r5 = a[(start + 5) & mask], r6 = a[(start + 6) & mask], r7 = a[(start + 7) & mask];
    int i;
#pragma unroll 1
    for (i = 8; i < n - (n % 8); i += 8) {
        r0 += a[(start + i) & mask]; r1 += a[(start + i + 1) & mask]; r2 += a[(start + i + 2) & mask]; r3 += a[(start + i + 3) & mask];
        r4 += a[(start + i + 4) & mask]; r5 += a[(start + i + 5) & mask]; r6 += a[(start + i + 6) & mask]; r7 += a[(start + i + 7) & mask];
    }
    T res = ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
#pragma unroll 1
    for (; i < n; i++) res += a[(start + i) & mask];
    return res;
}
template <typename T>
FTL_HD T ring_np_sum(const T* a, int start, int mask, int n) {   // np_sum over a[(start + k) & mask], k < n <= 512
    if (n <= 128) return ring_block_sum<T>(a, start, mask, n);
    int n2 = n / 2;
    n2 -= n2 % 8;
    T x, y;
    if (n2 <= 128) x = ring_block_sum<T>(a, start, mask, n2);
    else { int m = n2 / 2; m -= m % 8; x = ring_block_sum<T>(a, start, mask, m) + ring_block_sum<T>(a, start + m, mask, n2 - m); }
    int nr = n - n2;
    if (nr <= 128) y = ring_block_sum<T>(a, start + n2, mask, nr);
    else { int m = nr / 2; m -= m % 8; y = ring_block_sum<T>(a, start + n2, mask, m) + ring_block_sum<T>(a, start + n2 + m, mask, nr - m); }
    return x + y;
}
FTL_HD bool tracker_len_exceeds(const double* seg_d, const float* seg_f, int tail, int n, int mask, bool f64,
                                double limit_d, float limit_f) {
#if defined(__CUDA_ARCH__) && !defined(FTL_NO_RING_PREFETCH)   // measured (r02_ab_log.txt (18)): k_kin 0.1352 -> 0.1342 ms
    // the block sums below read the live ring in dependent batches of eight loads: ask for all of its lines first (the ring
    // of one env is (mask + 1) * 8 or * 4 contiguous bytes), so that only the first batch pays the DRAM round trip
    {
        const char* ring = f64 ? reinterpret_cast<const char*>(seg_d) : reinterpret_cast<const char*>(seg_f);
        const int bytes = (mask + 1) * (f64 ? 8 : 4);
        for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L1 [%0];" ::"l"(ring + o));
    }
#endif
    if (f64) return ring_np_sum<double>(seg_d, tail, mask, n - 1) > limit_d;
    return ring_np_sum<float>(seg_f, tail, mask, n - 1) > limit_f;
}
FTL_HD bool tracker_too_long(const DevCfg& cfg, const Tracker& t, const double* seg_d, const float* seg_f, int cap) {
    int n = t.ring_head - t.ring_tail;
    if (n < 2) return false;
    return tracker_len_exceeds(seg_d, seg_f, t.ring_tail, n, cap - 1, t.ring_tail < t.hist_f64_end, cfg.c.corridor_length,
                               cfg.corridor_length_f32);
}
#else
FTL_HD bool tracker_too_long(const DevCfg& cfg, const Tracker& t, const double* seg_d, const float* seg_f, int cap) {
    int n = t.ring_head - t.ring_tail;
    if (n < 2) return false;
    int mask = cap - 1, tail = t.ring_tail;
    if (tail < t.hist_f64_end) {
        double len = np_sum<double>(n - 1, [&](int k) { return seg_d[(tail + k) & mask]; });
        return len > cfg.c.corridor_length;
    } else {
        float len = np_sum<float>(n - 1, [&](int k) { return seg_f[(tail + k) & mask]; });
        return len > cfg.corridor_length_f32;
    }
}

#endif

FTL_HD float4 corridor_entry(const DevCfg& cfg, const Tracker& t, const double2* hist, int cap, int ia, int ib,
                             int ianchor) {  // SEN:302-317
    int mask = cap - 1;
    double2 pa = hist[ia & mask], pb = hist[ib & mask], pc = hist[ianchor & mask];
    double vx, vy;
    if (ia < t.hist_f64_end || ib < t.hist_f64_end) {
        vx = pb.x - pa.x;
        vy = pb.y - pa.y;
        double nrm = sqrt(fma(vy, vy, vx * vx));
        double s = cfg.c.corridor_width / nrm;
        vx *= s;
        vy *= s;
    } else {
        float fx = (float)pb.x - (float)pa.x, fy = (float)pb.y - (float)pa.y;
        float sx = fx * fx, sy = fy * fy;
        float ss = sx + sy;
        float s = cfg.corridor_width_f32 / sqrtf(ss);
        float gx = fx * s, gy = fy * s;
        vx = gx;
        vy = gy;
    }
    const double c90 = 6.123233995736766e-17;
    double rx = fma(c90, vx, -1.0 * vy), ry = fma(1.0, vx, c90 * vy);
    double lx = fma(c90, vx, 1.0 * vy), ly = fma(-1.0, vx, c90 * vy);
    rx += pc.x; ry += pc.y;
    lx += pc.x; ly += pc.y;
    return make_float4((float)rx, (float)ry, (float)lx, (float)ly);
}

struct TrackerInput {  // what the tracker reads from the robots, by value
    float fpx, fpy, lpx, lpy;
    double fdir;
};

FTL_HD_NOINLINE void tracker_scan(const DevCfg& cfg, Tracker& t, double2* hist, float4* corr, double* seg_d, float* seg_f,
                                  TrackerInput in, int* overflow) {
    struct { float px, py; double dir; } follower = {in.fpx, in.fpy, in.fdir};
    struct { float px, py; } leader = {in.lpx, in.lpy};
    const FtlConfig& c = cfg.c;
    int cap = c.corridor_cap, mask = cap - 1;
    if (t.saving_counter % c.saving_period == 0) {
        int n = t.ring_head - t.ring_tail;
        if (n > 0) {
            double2 last = hist[(t.ring_head - 1) & mask];
            if (last.x == (double)leader.px && last.y == (double)leader.py) return;  // SEN:247-251
        }
        if (n == 0 && t.saving_counter == 0) {
            double lx = leader.px, ly = leader.py;
            int m;
            if (c.start_corridor_behind_follower) {  // SEN:257-272
                double sn, cs;
                sincos_deg(angle_correction(follower.dir + 180), &sn, &cs);
                double sx = 50 * cs + (double)follower.px, sy = 50 * sn + (double)follower.py;
                m = (int)(dist_f64(sx, sy, lx, ly) / (c.saving_period * 5 * c.leader.max_speed));
                if (m > cap) { m = cap; *overflow |= 2; }
                double stepx = m > 1 ? (lx - sx) / (m - 1) : 0, stepy = m > 1 ? (ly - sy) / (m - 1) : 0;
                for (int i = 0; i < m; i++) {
                    double tx = (double)i * stepx, ty = (double)i * stepy;
                    hist[(t.ring_head + i) & mask] = make_double2(tx + sx, ty + sy);
                }
                t.hist_f64_end = t.ring_head + m;
            } else {  // SEN:273-281, float32 linspace
                float fsx = follower.px, fsy = follower.py, flx = leader.px, fly = leader.py;
                float q = sqrtf(d2_f32(fsx, fsy, flx, fly)) / (float)(c.saving_period * 5 * c.leader.max_speed);
                m = (int)q;
                if (m > cap) { m = cap; *overflow |= 2; }
                float ddx = flx - fsx, ddy = fly - fsy;
                float stepx = m > 1 ? ddx / (float)(m - 1) : 0.f, stepy = m > 1 ? ddy / (float)(m - 1) : 0.f;
                for (int i = 0; i < m; i++) {
                    float tx = (float)i * stepx, ty = (float)i * stepy;
                    float vx = tx + fsx, vy = ty + fsy;
                    hist[(t.ring_head + i) & mask] = make_double2((double)vx, (double)vy);
                }
                t.hist_f64_end = t.ring_tail;
            }
            if (m > 1) hist[(t.ring_head + m - 1) & mask] = make_double2(lx, ly);
            for (int i = 0; i + 1 < m; i++) tracker_seg_store(hist, seg_d, seg_f, mask, t.ring_head + i);
            t.ring_head += m;
        } else {
            if (t.ring_head - t.ring_tail >= cap) { t.ring_tail++; *overflow |= 2; }
            hist[t.ring_head & mask] = make_double2((double)leader.px, (double)leader.py);
            t.ring_head++;
            if (t.ring_head - t.ring_tail >= 2) tracker_seg_store(hist, seg_d, seg_f, mask, t.ring_head - 2);
        }
        while (tracker_too_long(cfg, t, seg_d, seg_f, cap)) t.ring_tail++;  // SEN:286-292
        n = t.ring_head - t.ring_tail;
        if (n > 1) {
            if (t.saving_counter == 0) {  // SEN:300-308
                for (int i = n - 1; i > 0; i--) {
                    int anchor = t.ring_tail + (n - i - 1);
                    corr[anchor & mask] = corridor_entry(cfg, t, hist, cap, t.ring_tail + i - 1, t.ring_tail + i, anchor);
                }
            }
            corr[(t.ring_head - 1) & mask] =
                corridor_entry(cfg, t, hist, cap, t.ring_head - 2, t.ring_head - 1, t.ring_head - 2);
        }
    }
    t.saving_counter += 1;
}

}  // namespace ftl
