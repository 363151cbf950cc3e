// ftl_rays.cuh -- history ray sensors (LeaderCorridor_Prev_lasers_v2.scan, SEN:883-962, with the
// segment test of SEN:608-640), restructured for the GPU:
//
//   * the reference re-casts every ray against H stored copies of the whole edge list (walls, rocks,
//     leader, bears, corridor sides, end caps).  Static rectangles are identical in every copy, so they
//     are cast ONCE per ray and merged into each valid history row; the corridor copies are sub-ranges
//     [tail_j, head_j) of one ring, so every ring segment is cast once and merged into the rows whose
//     range contains it.  Only the dynamic rectangles (leader + bears) and the end caps are cast per row.
//   * static rectangles are culled with the ray's bounding box before their four edges are tested.
//   * a history entry is stored by reference (ring range + dynamic rectangles), 8 + 16*(1+B) bytes instead
//     of ~2.5 KB of edges.
//
// Arithmetic: float32, with the reference's own float32 operations reproduced where it uses float32
// (edge vectors, the numerator of the intersection parameter, ccw(A,B,C)); the reference evaluates the
// predicates that involve the ray end point in float64 -- here they are float32 cross products with
// fused multiply-adds (documented tolerance: 1e-4 relative on the distances).
#pragma once

#include "ftl_device.cuh"

namespace ftl {

constexpr float kNoHit = 3.0e38f;

// difference of products a*b - c*d with one rounding error (Kahan): exact products via FMA residuals
FTL_HD float diff_of_products(float a, float b, float c, float d) {
    float w = c * d;
    float e = fmaf(-c, d, w);   // w - c*d exactly
    float f = fmaf(a, b, -w);   // a*b - w rounded once
    return f + e;
}

// distance along the ray (origin P, vector D of length L) to segment AB, or kNoHit.
// Mirrors intersect()/seg_intersect(): strict ccw tests, t = num/denom, value = t * L.
FTL_HD float seg_hit(float px, float py, float dx, float dy, float L, float ax, float ay, float bx, float by) {
    float uax = px - ax, uay = py - ay;  // C - A, float32 like the reference
    float ubx = px - bx, uby = py - by;  // C - B
    // ccw(A,C,D) != ccw(B,C,D): A and B on different sides of the ray's line
    float ca = diff_of_products(uax, dy, uay, dx);
    float cb = diff_of_products(ubx, dy, uby, dx);
    if ((ca > 0.f) == (cb > 0.f)) return kNoHit;
    float ex = bx - ax, ey = by - ay;
    float p3 = uay * ex, q3 = ey * uax;  // ccw(A,B,C): float32 products compared, exactly the reference's
    bool c3 = p3 > q3;
    float s1 = diff_of_products(uay, ex, ey, uax);
    float s2 = diff_of_products(ex, dy, ey, dx);   // = dap . db, the reference's denominator
    bool c4 = (s1 + s2) > 0.f;                       // ccw(A,B,D)
    if (c3 == c4) return kNoHit;
    float m0 = (-ey) * (-uax), m1 = ex * (-uay);    // np.multiply(dap, dp) in float32
    float num = m0 + m1;
    float t = num / s2;
    return fabsf(t) * L;
}

FTL_HD float rect_hit(float px, float py, float dx, float dy, float L, int4 q, float lox, float hix, float loy,
                      float hiy) {
    float l = (float)q.x, t = (float)q.y, r = (float)(q.x + q.z), b = (float)(q.y + q.w);
    if (r < lox || l > hix || b < loy || t > hiy) return kNoHit;  // outside the ray's bounding box
    float m = seg_hit(px, py, dx, dy, L, l, b, r, b);             // SEN:668-671 edge order
    m = fminf(m, seg_hit(px, py, dx, dy, L, r, t, r, b));
    m = fminf(m, seg_hit(px, py, dx, dy, L, r, t, l, t));
    m = fminf(m, seg_hit(px, py, dx, dy, L, l, b, l, t));
    return m;
}

struct RayEnv {  // per-env inputs of the ray pass, loaded once
    float px, py;
    double dir;
    int scenario, snap_pushes;
};

FTL_HD void ray_env_load(const DevState& s, int i, RayEnv& r) {
    float2 p = s.pos[i];  // robot 0 = follower
    r.px = p.x; r.py = p.y;
    r.dir = s.rd[(size_t)RD_DIR * s.n + i];
    r.scenario = s.gi[(size_t)GI_SCENARIO * s.n + i];
    r.snap_pushes = s.gi[(size_t)GI_SNAP_PUSHES * s.n + i];
}

// One ray of one sensor of one env: writes H values (history rows, oldest first) into rows[].
FTL_HD void cast_ray(const DevCfg& cfg, const DevState& s, const DevPool& pool, int i, const RayEnv& re,
                     const FtlRaySensorConfig& sc, int k, float* rows) {
    const FtlConfig& c = cfg.c;
    const int H = sc.max_prev_obs;
    const int NBr = s.n_bears;
    const float L = (float)sc.laser_length;
    double ang = (re.dir + sc.first_laser_angle_offset) + k * (360.0 / sc.lasers_count);
    double sn, cs;
    sincos_deg(ang, &sn, &cs);
    const float dx = (float)(cs * sc.laser_length), dy = (float)(sn * sc.laser_length);
    const float px = re.px, py = re.py;
    const float lox = fminf(px, px + dx) - 1.f, hix = fmaxf(px, px + dx) + 1.f;
    const float loy = fminf(py, py + dy) - 1.f, hiy = fmaxf(py, py + dy) + 1.f;

    // history rows: row j has age H-1-j; valid once enough scans happened (SEN:964-968 seeds zeros)
    int tail[FTL_MAX_HIST], head[FTL_MAX_HIST];
    bool valid[FTL_MAX_HIST];
    int min_tail = 0x7fffffff, max_head = -0x7fffffff;
#pragma unroll
    for (int j = 0; j < FTL_MAX_HIST; j++) {
        rows[j] = kNoHit;
        valid[j] = false;
        tail[j] = head[j] = 0;
        if (j < H) {
            int age = H - 1 - j;
            if (age < re.snap_pushes) {
                int slot = (re.snap_pushes - 1 - age) % FTL_MAX_HIST;
                int2 rg = s.snap_range[(size_t)slot * s.n + i];
                tail[j] = rg.x; head[j] = rg.y;
                valid[j] = true;
                min_tail = rg.x < min_tail ? rg.x : min_tail;
                max_head = rg.y > max_head ? rg.y : max_head;
            }
        }
    }
    const int mode = sc.react_to_obstacles;
    // ---- static rectangles: once for all rows ------------------------------------------------------------
    if (mode == FTL_REACT_ALL || mode == FTL_REACT_STATIC) {
        const int4* statics = pool.static_rects + (size_t)re.scenario * c.static_cap;
        const int n_static = pool.n_static[re.scenario];
        float m = kNoHit;
        for (int q = 0; q < n_static; q++) m = fminf(m, rect_hit(px, py, dx, dy, L, statics[q], lox, hix, loy, hiy));
#pragma unroll
        for (int j = 0; j < FTL_MAX_HIST; j++)
            if (valid[j]) rows[j] = fminf(rows[j], m);
    }
    // ---- dynamic rectangles: per row (leader belongs to game_object_list: ALL and STATIC) -----------------
#pragma unroll
    for (int j = 0; j < FTL_MAX_HIST; j++) {
        if (!valid[j]) continue;
        int age = H - 1 - j;
        int slot = (re.snap_pushes - 1 - age) % FTL_MAX_HIST;
        const int4* sr = s.snap_rect + ((size_t)slot * (1 + NBr)) * s.n + i;
        if (mode == FTL_REACT_ALL || mode == FTL_REACT_STATIC)
            rows[j] = fminf(rows[j], rect_hit(px, py, dx, dy, L, sr[0], lox, hix, loy, hiy));
        if (mode == FTL_REACT_ALL || mode == FTL_REACT_DYNAMIC)
            for (int b = 0; b < NBr; b++)
                rows[j] = fminf(rows[j], rect_hit(px, py, dx, dy, L, sr[(size_t)(1 + b) * s.n], lox, hix, loy, hiy));
    }
    // ---- corridor ring: every side segment once, merged into the rows whose range holds it ------------------
    const float4* corr = s.corridor + (size_t)i * c.corridor_cap;
    const int cmask = c.corridor_cap - 1;
    if (sc.react_to_safe_corridor && max_head > min_tail) {
        float4 a = corr[min_tail & cmask];
        for (int q = min_tail; q < max_head - 1; q++) {
            float4 b = corr[(q + 1) & cmask];
            float m = kNoHit;
            bool rbox = !(fmaxf(a.x, b.x) < lox || fminf(a.x, b.x) > hix || fmaxf(a.y, b.y) < loy || fminf(a.y, b.y) > hiy);
            bool lbox = !(fmaxf(a.z, b.z) < lox || fminf(a.z, b.z) > hix || fmaxf(a.w, b.w) < loy || fminf(a.w, b.w) > hiy);
            if (rbox) m = seg_hit(px, py, dx, dy, L, a.x, a.y, b.x, b.y);
            if (lbox) m = fminf(m, seg_hit(px, py, dx, dy, L, a.z, a.w, b.z, b.w));
            if (m < kNoHit) {
#pragma unroll
                for (int j = 0; j < FTL_MAX_HIST; j++)
                    if (valid[j] && q >= tail[j] && q < head[j] - 1) rows[j] = fminf(rows[j], m);
            }
            a = b;
        }
    }
    if (sc.react_to_green_zone) {  // end caps of each stored corridor, SEN:648-650
#pragma unroll
        for (int j = 0; j < FTL_MAX_HIST; j++) {
            if (!valid[j]) continue;
            float4 a = corr[tail[j] & cmask], b = corr[(head[j] - 1) & cmask];
            rows[j] = fminf(rows[j], seg_hit(px, py, dx, dy, L, a.x, a.y, a.z, a.w));
            rows[j] = fminf(rows[j], seg_hit(px, py, dx, dy, L, b.x, b.y, b.z, b.w));
        }
    }
#pragma unroll
    for (int j = 0; j < FTL_MAX_HIST; j++)
        if (rows[j] >= kNoHit) rows[j] = L;  // no hit: the laser end point, SEN:926-930
}

// where ray k of a sensor lands in the env's output vector; pad_sectors layout of SEN:932-953
FTL_HD void store_ray_rows(const FtlRaySensorConfig& sc, float* dst, int k, const float* rows) {
    const int R = sc.lasers_count, H = sc.max_prev_obs;
    if (!sc.pad_sectors) {
        for (int j = 0; j < H; j++) dst[(size_t)j * R + k] = rows[j];
    } else {
        double in_sector = R / 4.0;
        int sector = (k < in_sector) ? 0 : (k < 2 * in_sector) ? 1 : (k < 3 * in_sector) ? 2 : 3;
        for (int j = 0; j < H; j++)
            for (int q = 0; q < 4; q++) dst[(size_t)j * 4 * R + (size_t)q * R + k] = (q == sector) ? rows[j] : 0.f;
    }
}

FTL_HD int sensor_width(const FtlRaySensorConfig& sc) {
    return sc.max_prev_obs * (sc.pad_sectors ? 4 * sc.lasers_count : sc.lasers_count);
}

// flat ray index -> (sensor, ray, output offset)
FTL_HD bool locate_ray(const FtlConfig& c, int flat, int* sensor, int* k, int* offset) {
    int off = 0;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) {
        int R = c.ray[sidx].lasers_count;
        if (flat < R) { *sensor = sidx; *k = flat; *offset = off; return true; }
        flat -= R;
        off += sensor_width(c.ray[sidx]);
    }
    return false;
}

FTL_HD int total_rays(const FtlConfig& c) {
    int n = 0;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) n += c.ray[sidx].lasers_count;
    return n;
}

}  // namespace ftl
