// ftl_rays.cuh -- history ray sensors (LeaderCorridor_Prev_lasers_v2.scan, SEN:883-962, with the
// segment test of SEN:608-640), restructured for the GPU:
//
//   * the reference re-casts every ray against H stored copies of the whole edge list (walls, rocks,
//     leader, bears, corridor sides, end caps).  Static rectangles are identical in every copy, so they
//     are cast ONCE per ray and merged into each valid history row; the corridor copies are sub-ranges
//     [tail_j, head_j) of one ring, so every ring segment is cast once and merged into the rows whose
//     range contains it.  Only the dynamic rectangles (leader + bears) and the end caps are cast per row.
//   * a history entry is stored by reference (ring range + dynamic rectangles), 8 + 16*(1+B) bytes instead
//     of ~2.5 KB of edges.
//   * edges are culled against the sensors' reach, rectangles reduced to their front-facing edges, and each
//     surviving edge is only tested against the rays inside the angle it subtends from the follower.
//
// Arithmetic: float32, with the reference's own float32 operations reproduced where it uses float32
// (edge vectors, the numerator of the intersection parameter, ccw(A,B,C)); the reference evaluates the
// predicates that involve the ray end point in float64 -- here they are float32 cross products with
// error-free products (documented tolerance: 1e-4 relative on the distances).
#pragma once

#include "ftl_device.cuh"

namespace ftl {
// what lands in `rays`: the distance, or the wrapper's clip(v / laser_length, 0, 1) (float32 division, WRP:211)
FTL_HD float ray_out_value(const FtlConfig& c, float L, float v) {
    if (!c.fused_sensor_prev) return v;
    v = v / L;
    return v < 0.f ? 0.f : v > 1.f ? 1.f : v;
}


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
// *uncertain is set when one of the three predicates that involve the ray end point is closer to zero than the
// float32 evaluation can resolve against the reference's float64 one (the ray grazes a vertex or ends on the
// edge's line within ~1e-4 px); the caller then repeats the test with the reference's float64 arithmetic.
FTL_HD float seg_hit(float px, float py, float dx, float dy, float L, float ax, float ay, float bx, float by,
                     bool* uncertain) {
    float uax = px - ax, uay = py - ay;  // C - A, float32 like the reference
    float ubx = px - bx, uby = py - by;  // C - B
    // ccw(A,C,D) != ccw(B,C,D): A and B on different sides of the ray's line
    float ca = diff_of_products(uax, dy, uay, dx);
    float cb = diff_of_products(ubx, dy, uby, dx);
    // one bound for all three predicates.  Edges are culled to |C-A|_1, |C-B|_1 <= 2(L+1), so the rounding of the
    // float32 evaluation against the reference's float64 one stays below M(L+M)*2.5e-7 <= 1.5e-6 L^2: a ray that
    // passes a vertex (or ends at the edge's line) closer than ~3e-4 px is re-done in float64.
    const float bound = L * L * 1.5e-6f;
    bool unc = fminf(fabsf(ca), fabsf(cb)) <= bound;
    if (!unc && (ca > 0.f) == (cb > 0.f)) { *uncertain = false; return kNoHit; }
    float ex = bx - ax, ey = by - ay;
    float p3 = uay * ex, q3 = ey * uax;  // ccw(A,B,C): float32 products compared, exactly the reference's
    bool c3 = p3 > q3;
    float s1 = diff_of_products(uay, ex, ey, uax);
    float s2 = diff_of_products(ex, dy, ey, dx);   // = dap . db, the reference's denominator
    float s12 = s1 + s2;
    unc = unc || fabsf(s12) <= 2.f * bound;
    *uncertain = unc;
    if (unc) return kNoHit;
    bool c4 = s12 > 0.f;                             // ccw(A,B,D)
    if (c3 == c4) return kNoHit;
    float m0 = (-ey) * (-uax), m1 = ex * (-uay);    // np.multiply(dap, dp) in float32
    float num = m0 + m1;
    float t = num / s2;
    return fabsf(t) * L;
}

// The same test with the reference's arithmetic (SEN:608-640 on the arrays SEN:903-906 builds): float32 edge and
// follower position, float64 ray end point.  Returns the float32 value the reference stores, or kNoHit.
FTL_HD_NOINLINE float seg_hit_exact(float px, float py, double ex, double ey, float ax, float ay, float bx, float by) {
    float uax = px - ax, uay = py - ay, ubx = px - bx, uby = py - by;
    bool cA = (ey - (double)ay) * (double)uax > (double)uay * (ex - (double)ax);   // ccw(A, C, D)
    bool cB = (ey - (double)by) * (double)ubx > (double)uby * (ex - (double)bx);   // ccw(B, C, D)
    if (cA == cB) return kNoHit;
    float edx = bx - ax, edy = by - ay;
    float p3 = uay * edx, q3 = edy * uax;
    bool c3 = p3 > q3;                                                             // ccw(A, B, C), float32
    bool c4 = (ey - (double)ay) * (double)edx > (double)edy * (ex - (double)ax);   // ccw(A, B, D)
    if (c3 == c4) return kNoHit;
    double dbx = ex - (double)px, dby = ey - (double)py;
    float dpx = ax - px, dpy = ay - py, dapx = -edy, dapy = edx;
    double denom = fma((double)dapx, dbx, (double)dapy * dby);
    float m0 = dapx * dpx, m1 = dapy * dpy;
    float num = m0 + m1;
    double t = (double)num / denom;
    double xx = t * dbx + (double)px, xy = t * dby + (double)py;
    double ddx = xx - (double)px, ddy = xy - (double)py;
    return (float)sqrt(fma(ddy, ddy, ddx * ddx));
}

FTL_HD int sensor_row_width(const FtlRaySensorConfig& sc) {
    return sc.compas ? 5 * sc.lasers_count : sc.pad_sectors ? 4 * sc.lasers_count : sc.lasers_count;
}
FTL_HD int sensor_width(const FtlRaySensorConfig& sc) { return sc.max_prev_obs * sensor_row_width(sc); }

// direction of ray k relative to (heading + first_laser_angle_offset), degrees: SEN:888-891, or the fixed fan of
// LeaderCorridor_lasers (SEN:678-700)
FTL_HD double ray_angle(const FtlRaySensorConfig& sc, int k) {
    return sc.n_custom_angles ? sc.custom_angle[k] : k * (360.0 / sc.lasers_count);
}

// raw: sensor blocks one after the other, each [H][W]; fused (WRP:203-221): one [H][sum W] matrix
inline void ray_out_layout(DevCfg& d) {
    const FtlConfig& c = d.c;
    int sum_w = 0, off = 0, col = 0;
    for (int s = 0; s < c.n_ray_sensors; s++) sum_w += sensor_row_width(c.ray[s]);
    for (int s = 0; s < c.n_ray_sensors; s++) {
        const int w = sensor_row_width(c.ray[s]);
        d.ray_out_base[s] = c.fused_sensor_prev ? col : off;
        d.ray_out_stride[s] = c.fused_sensor_prev ? sum_w : w;
        col += w;
        off += w * c.ray[s].max_prev_obs;
    }
    d.ray_out_vec4 = c.n_ray_sensors > 0 && !c.fused_sensor_prev;
    d.ray_out_fused_vec4 = c.n_ray_sensors > 0 && c.fused_sensor_prev && d.rays_per_env % 4 == 0;
    for (int s = 0; s < c.n_ray_sensors; s++)
        if (c.ray[s].pad_sectors || c.ray[s].compas || c.ray[s].lasers_count % 4 != 0) d.ray_out_vec4 = d.ray_out_fused_vec4 = 0;
}

FTL_HD int total_rays(const FtlConfig& c) {
    int n = 0;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) n += c.ray[sidx].lasers_count;
    return n;
}

FTL_HD int f2i_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i; memcpy(&i, &f, 4); return i;
#endif
}
FTL_HD float i2f_bits(int i) {
#if defined(__CUDA_ARCH__)
    return __int_as_float(i);
#else
    float f; memcpy(&f, &i, 4); return f;
#endif
}

// =====================================================================================================
// Warp-cooperative ray pass (one warp per env).
//
// Written as "lane phases": every FTL_LANES(lane) block is executed by the 32 lanes of a warp on the
// GPU and by a plain loop over lane = 0..31 in the host test build; phases only communicate through
// the RayShared block (shared memory on the GPU), so the same source runs in both places.
//
//   setup   per-sensor table, ray direction table (one float64 sincos per sensor, rotated by a host-made
//           table of cos/sin(k * 360/R)), result arrays
//   A1      raw items (static rects, dynamic rect snapshots, corridor segments, end caps), one per lane:
//           culled against the sensors' reach box; rectangles keep only their front-facing edges; survivors
//           go to a compact edge list in shared memory
//   A2      one edge per lane: the angular interval the edge subtends from the follower (polynomial atan2,
//           widened) selects the few rays that can reach it -> compact (edge, ray) pair list
//   B       one pair per lane: the exact segment test; results merged with integer atomicMin (non-negative
//           floats order like ints) into res[age][ray] / res[static][ray]
//   out     rows assembled (static minimum merged into every valid history row), written coalesced
// =====================================================================================================
// FTL_RAYS_LANES: lanes that serve one env -- 32 (one env per warp) or 16 (two envs per warp: each half-warp runs the
// pass of its own env with its own shared block and only ever synchronises with its own 16 lanes; the halves share the
// warp's instruction stream wherever their control flow agrees).  The item lists of an env are short (37 rectangles,
// ~10 stored rectangles, ~50 edges, 48 rays): with 32 lanes 18 are active on average.
#ifndef FTL_RAYS_LANES
#define FTL_RAYS_LANES 32
#endif
constexpr int kLanes = FTL_RAYS_LANES;
static_assert(kLanes == 32 || kLanes == 16, "FTL_RAYS_LANES: 32 or 16");
#if defined(__CUDA_ARCH__)
#define FTL_LANES(lane) for (int lane = (int)(threadIdx.x & (kLanes - 1)), ftl_once_ = 1; ftl_once_; ftl_once_ = 0)
#define FTL_LANE_MASK() (kLanes == 32 ? 0xffffffffu : (0xffffu << (threadIdx.x & 16)))
#define FTL_WARP_SYNC() __syncwarp(FTL_LANE_MASK())
// A decision every lane of the env must take alike, from a shared word that the lanes go on to modify: __syncwarp() orders
// memory but does not make the lanes run in lockstep afterwards, so a lane that reads the word late may already see the
// appends of a faster one (found with a 24-edge list, where "flush first?" sits on its boundary all the time: some lanes
// flushed and some did not, profiles/r02_ab_log.txt (24)).  The first lane's reading is broadcast; nobody gets past the
// shuffle before everyone has arrived.
#define FTL_UNIFORM_INT(val_) __shfl_sync(FTL_LANE_MASK(), (val_), (int)(threadIdx.x & (32 - kLanes)))
FTL_HD int smem_atomic_add(int* p, int v) { return atomicAdd(p, v); }
FTL_HD void smem_atomic_min(int* p, int v) { atomicMin(p, v); }
#else
#ifdef FTL_DBG_LANE_ORDER   // host test builds: another lane order (the GPU serves the shared lists in no particular order)
#define FTL_LANES(lane) for (int ftl_k_ = 0, lane = FTL_DBG_LANE_ORDER % kLanes; ftl_k_ < kLanes; ++ftl_k_, lane = (lane * 5 + FTL_DBG_LANE_ORDER) % kLanes)
#else
#define FTL_LANES(lane) for (int lane = 0; lane < kLanes; ++lane)
#endif
#define FTL_WARP_SYNC() ((void)0)
#define FTL_UNIFORM_INT(val_) (val_)
FTL_HD int smem_atomic_add(int* p, int v) { int o = *p; *p = o + v; return o; }
FTL_HD void smem_atomic_min(int* p, int v) { if (v < *p) *p = v; }
#endif

#ifndef FTL_EDGE_CAP
#if FTL_RAYS_LANES == 32
#define FTL_EDGE_CAP 176   // measured: 160 -> 176 saves the mid-env flush of about half the envs (-2.8 % k_rays); 224 costs occupancy
#else
#define FTL_EDGE_CAP 112   // a round of 16 lanes adds at most 64 edges: the same 48 edges of head room before a flush is forced
#endif
#endif
#ifndef FTL_PAIR_CAP
#if FTL_RAYS_LANES == 32
#define FTL_PAIR_CAP 320
#else
#define FTL_PAIR_CAP 256
#endif
#endif
constexpr int kEdgeCap = FTL_EDGE_CAP;  // compact edge list per flush (overridable: tests build with tiny lists)
constexpr int kPairCap = FTL_PAIR_CAP;  // (edge, ray) pairs per flush
constexpr int kCorridorChunk = 2 * kLanes;  // corridor ring entries per batch (2 edges each)
constexpr int kStaticBit = 8;       // row-mask bit of the static minimum (merged into all valid rows)
constexpr int kNoHitBits = 0x7f7fffff;
// a pair word: flat ray (12 bits), edge slot (8 bits), rows of the ray's sensor that the edge belongs to (9 bits)
constexpr int kPairEdgeShift = 12, kPairRowsShift = 20;
static_assert(kEdgeCap <= (1 << (kPairRowsShift - kPairEdgeShift)), "edge slot does not fit the pair word");
#if defined(__CUDACC__) && !defined(FTL_ALLOW_EDGE_OVERFLOW)
// The list is flushed whenever the next round of appends (at most 4 edges per lane, or one corridor batch) might not fit,
// so with this much room the overflow path of edge_append (edge_inline) is never taken in a product build: which edges
// would be tested in place depends on the order the lanes get their slots in, and with it which inconclusive pairs are
// recorded -- results would stay within tolerance but no longer be bit-reproducible.  The test build with tiny lists
// (build.build_small_lists, libftl_hostsim_smallcaps.so) defines FTL_ALLOW_EDGE_OVERFLOW.
static_assert(kEdgeCap >= 4 * kLanes + 2 * FTL_MAX_HIST, "FTL_EDGE_CAP too small for a GPU build: the edge list must never overflow");
#endif
constexpr int kMaxTotalRays = 1 << kPairEdgeShift;   // checked at ftl_create
enum EdgeClass { EC_STATIC = 0, EC_LEADER = 1, EC_BEAR = 2, EC_CORRIDOR = 3, EC_CAP = 4, EC_COUNT = 5 };

struct RayEdge { float ax, ay, bx, by; int mask; };  // mask: bits 0..8 rows (ages / static), bits 16.. class

#ifndef FTL_UNC_PER_ENV
#define FTL_UNC_PER_ENV 16
#endif
constexpr int kUncPerEnv = FTL_UNC_PER_ENV;  // records per env for the exact pass; more -> the env is recast exactly

struct RaySensorTab {
    int base, R, H, cls_mask;      // cls_mask: which EdgeClass this sensor reacts to
    float L, theta0, inv_period, eps, inv_R, pad_;
    double cs0, sn0;               // cos/sin of the direction of ray 0
};

struct alignas(16) RayShared {   // 16-byte multiple: the arrays behind it are read as int4 for the float4 row output
    float px, py;
    double dir;                                    // follower heading (float64, for the exact fallback)
    int scenario, snap_pushes, n_valid, ne, np, rt, ns, hmax;
    float reach[EC_COUNT];                         // largest laser_length among sensors reacting to the class
    int tail[FTL_MAX_HIST], head[FTL_MAX_HIST];    // by age (0 = newest)
    RaySensorTab sen[FTL_MAX_RAY_SENSORS];
    RayEdge e[kEdgeCap];
    int pair[kPairCap];                            // rows << 20 | edge << 12 | flat ray
#ifdef FTL_DBG_EDGES
    int dbg[16];                                   // diagnostic counters (appended, in place, listed, flushes, pairs, ...)
#endif
    int nu;                                        // (edge, ray) pairs of this env whose float32 predicates were inconclusive
    UncRec* unc;                                   // this env's slice of DevState.unc_rec
    // arrays of length rays_total behind the struct: dx, dy, len (float), res[hmax + 1] (int): one row of minima per
    // history age in use plus the row of the static minimum
};

struct RayArrays { float *dx, *dy, *len; int* res; int rt, hmax; };   // hmax: largest max_prev_obs = index of the static row

FTL_HD int ray_hmax(const FtlConfig& c) {
    int h = 1;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) h = c.ray[sidx].max_prev_obs > h ? c.ray[sidx].max_prev_obs : h;
    return h;
}
FTL_HD size_t ray_shared_bytes(int rays_total, int hmax) {
    return sizeof(RayShared) + (size_t)rays_total * (3 * 4 + (hmax + 1) * 4);
}
FTL_HD RayArrays ray_arrays(RayShared* sh, int rt, int hmax) {
    RayArrays a;
    float* base = (float*)(sh + 1);
    a.dx = base; a.dy = base + rt; a.len = base + 2 * rt;
    a.res = (int*)(base + 3 * rt);
    a.rt = rt;
    a.hmax = hmax;
    return a;
}

FTL_HD int sensor_class_mask(const FtlRaySensorConfig& sc) {
    if (sc.compas) return 0;   // cast by the per-env exact pass (compas_exact_env): no edge class reaches its rays here
    int m = 0, mode = sc.react_to_obstacles;
    if (mode == FTL_REACT_ALL || mode == FTL_REACT_STATIC) m |= (1 << EC_STATIC) | (1 << EC_LEADER);  // game_object_list
    if (mode == FTL_REACT_ALL || mode == FTL_REACT_DYNAMIC) m |= 1 << EC_BEAR;                        // game_dynamic_list
    if (sc.react_to_safe_corridor) m |= 1 << EC_CORRIDOR;
    if (sc.react_to_green_zone) m |= 1 << EC_CAP;
    return m;
}

// atan2 in degrees, |error| < 1e-3 degrees (odd minimax polynomial on [0,1] + octant folding)
// env-independent part of the per-sensor tables, computed once at ftl_create
inline void ray_static_tables(DevCfg& d) {
    const FtlConfig& c = d.c;
    for (int k = 0; k < 8; k++) d.ray_reach[k] = -1e30f;
    d.ray_compas_mask = 0;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++)
        if (c.ray[sidx].compas) d.ray_compas_mask |= 1 << sidx;
    int base = 0;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) {
        const FtlRaySensorConfig& sc = c.ray[sidx];
        RaySensorStatic& st = d.ray_static[sidx];
        st.base = base; st.R = sc.lasers_count; st.H = sc.max_prev_obs;
        st.cls_mask = sensor_class_mask(sc);
        st.L = (float)sc.laser_length;
        st.inv_period = (float)sc.lasers_count / 360.f;
        // rays that are not evenly spaced (custom angles) cannot be binned by angle: a huge margin makes every
        // edge a candidate for every ray of the sensor (cnt >= R in ray_flush)
        st.eps = sc.n_custom_angles ? 1e6f : 0.02f + 5e-5f * (float)sc.lasers_count;
        st.inv_R = 1.f / (float)sc.lasers_count;
        for (int k = 0; k < EC_COUNT; k++)
            if (st.cls_mask & (1 << k)) d.ray_reach[k] = fmaxf(d.ray_reach[k], st.L);
        base += sc.lasers_count;
    }
}

FTL_HD float atan2_deg_approx(float y, float x) {
    float ax = fabsf(x), ay = fabsf(y);
    float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#if defined(__CUDA_ARCH__)
    float t = __fdividef(mn, fmaxf(mx, 1e-30f));   // candidates only need to be a superset: 2 ulp are inside the margin
#else
    float t = mn / fmaxf(mx, 1e-30f);
#endif
    float t2 = t * t;
    float p = fmaf(t2, -0.01172120f, 0.05265332f);
    p = fmaf(t2, p, -0.11643287f);
    p = fmaf(t2, p, 0.19354346f);
    p = fmaf(t2, p, -0.33262347f);
    p = fmaf(t2, p, 0.99997726f);
    float r = p * t * 57.29577951308232f;
    if (ay > ax) r = 90.f - r;
    if (x < 0.f) r = 180.f - r;
    return y < 0.f ? -r : r;
}

// hot paths only RECORD an inconclusive pair; k_rays_exact (one thread per env, almost always idle) redoes it in
// float64.  No float64 code and no calls in the ray kernel's loops.
FTL_HD void unc_push(RayShared& sh, float ax, float ay, float bx, float by, int f, int rows) {
    int slot = smem_atomic_add(&sh.nu, 1);
    if (slot < kUncPerEnv) {
        UncRec r = {ax, ay, bx, by, f, rows};
        sh.unc[slot] = r;
    }
}

FTL_HD void hit_merge(const RayArrays& ra, int f, int rows, float d) {
    if (d >= kNoHit) return;
    int bits = f2i_bits(d);
#ifndef FTL_MERGE_PREDICATED   // measured (r02_ab_log.txt (27)): k_rays 0.2245 -> 0.2205 ms
    if (rows & (1 << kStaticBit)) smem_atomic_min(&ra.res[ra.hmax * ra.rt + f], bits);
    rows &= (1 << kStaticBit) - 1;
    while (rows) {   // one atomic per row the edge belongs to (bits at or above hmax are never set)
#if defined(__CUDA_ARCH__)
        const int a = __ffs(rows) - 1;
#else
        const int a = __builtin_ctz((unsigned)rows);
#endif
        rows &= rows - 1;
        smem_atomic_min(&ra.res[a * ra.rt + f], bits);
    }
#else
    // constant trip count (unrolled, predicated): bits at or above hmax are never set, so no row beyond hmax is touched
    for (int a = 0; a < FTL_MAX_HIST; a++)
        if (rows & (1 << a)) smem_atomic_min(&ra.res[a * ra.rt + f], bits);
    if (rows & (1 << kStaticBit)) smem_atomic_min(&ra.res[ra.hmax * ra.rt + f], bits);
#endif
}

FTL_HD void edge_ray_test(RayShared& sh, const RayArrays& ra, int f, int rows, float ax, float ay, float bx, float by) {
    bool uncertain;
    float d = seg_hit(sh.px, sh.py, ra.dx[f], ra.dy[f], ra.len[f], ax, ay, bx, by, &uncertain);
    if (uncertain) unc_push(sh, ax, ay, bx, by, f, rows);
    else hit_merge(ra, f, rows, d);
}

// all candidate rays of one edge, tested in place (only used when the shared lists are full)
#ifdef FTL_DBG_INLINE_EDGE
FTL_HD void edge_inline(
#else
FTL_HD_NOINLINE void edge_inline(
#endif
RayShared& sh, int rt, const RayEdge ed, int n_sensors) {
    const RayArrays ra = ray_arrays(&sh, rt, sh.hmax);
    const int cls_bit = ed.mask >> 16;
    for (int sidx = 0; sidx < n_sensors; sidx++) {
        const RaySensorTab& st = sh.sen[sidx];
        if (!(st.cls_mask & cls_bit)) continue;
        int rows = ed.mask & 0x1ff;
        if (!(rows & (1 << kStaticBit))) rows &= (1 << st.H) - 1;
        if (!rows) continue;
        for (int k = 0; k < st.R; k++) edge_ray_test(sh, ra, st.base + k, rows, ed.ax, ed.ay, ed.bx, ed.by);
    }
}

FTL_HD void edge_append(RayShared& sh, float ax, float ay, float bx, float by, int mask) {
    int slot = smem_atomic_add(&sh.ne, 1);
    RayEdge ed = {ax, ay, bx, by, mask};
#ifdef FTL_DBG_EDGES
    smem_atomic_add(&sh.dbg[0], 1);
    for (int k = 0; k < EC_COUNT; k++) if (mask & (1 << (16 + k))) smem_atomic_add(&sh.dbg[8 + k], 1);
    if (slot >= kEdgeCap) { smem_atomic_add(&sh.dbg[1], 1); smem_atomic_add(&sh.dbg[7], (int)(ax + ay + bx + by)); }
    smem_atomic_add(&sh.dbg[6], (int)(ax + ay + bx + by));   // order-independent checksum of what was appended
#endif
    if (slot < kEdgeCap)
        sh.e[slot] = ed;
    else
        edge_inline(sh, sh.rt, ed, sh.ns);
}

// A1 for one rectangle: reach cull + front-facing edges (edge order/orientation of SEN:668-671).  A ray from
// outside enters through a front-facing edge; the exit through a back-facing edge is farther and never the
// minimum the reference reports.  (inside only happens after a crash: then every edge is kept.)
FTL_HD void edge_store(RayShared& sh, int slot, float ax, float ay, float bx, float by, int mask) {
    RayEdge ed = {ax, ay, bx, by, mask};
    if (slot < kEdgeCap)
        sh.e[slot] = ed;
    else
        edge_inline(sh, sh.rt, ed, sh.ns);
}
FTL_HD void rect_append(RayShared& sh, int4 q, int cls, int rows) {
    const float l = (float)q.x, t = (float)q.y, r = (float)(q.x + q.z), b = (float)(q.y + q.w);
    const float px = sh.px, py = sh.py, reach = sh.reach[cls] + 1.f;
    if (r < px - reach || l > px + reach || b < py - reach || t > py + reach) return;
    const int mask = rows | (1 << (16 + cls));
    const bool inside = px >= l && px <= r && py >= t && py <= b;
#ifdef FTL_APPEND_PER_EDGE   // measured (r02_ab_log.txt (28)): one atomic per edge
    if (inside || py > b) edge_append(sh, l, b, r, b, mask);
    if (inside || px > r) edge_append(sh, r, t, r, b, mask);
    if (inside || py < t) edge_append(sh, r, t, l, t, mask);
    if (inside || px < l) edge_append(sh, l, b, l, t, mask);
#else
    // one reservation for the rectangle's front-facing edges (two from outside), then the stores
    const bool e0 = inside || py > b, e1 = inside || px > r, e2 = inside || py < t, e3 = inside || px < l;
    int slot = smem_atomic_add(&sh.ne, (int)e0 + (int)e1 + (int)e2 + (int)e3);
    if (e0) edge_store(sh, slot++, l, b, r, b, mask);
    if (e1) edge_store(sh, slot++, r, t, r, b, mask);
    if (e2) edge_store(sh, slot++, r, t, l, t, mask);
    if (e3) edge_store(sh, slot++, l, b, l, t, mask);
#endif
}
FTL_HD bool seg_in_reach(const RayShared& sh, float ax, float ay, float bx, float by, int cls) {
    const float px = sh.px, py = sh.py, reach = sh.reach[cls] + 1.f;
    return !(fmaxf(ax, bx) < px - reach || fminf(ax, bx) > px + reach || fmaxf(ay, by) < py - reach ||
             fminf(ay, by) > py + reach);
}
FTL_HD void seg_append(RayShared& sh, float ax, float ay, float bx, float by, int cls, int rows) {
    if (seg_in_reach(sh, ax, ay, bx, by, cls)) edge_append(sh, ax, ay, bx, by, rows | (1 << (16 + cls)));
}

// A2 + B over the current edge list, then empty it.  One out-of-line copy for the five places that may flush: the
// kernel is bound by instruction fetch as much as by issue slots, and warps in different flushes share these lines.
#ifdef FTL_INLINE_FLUSH
FTL_HD void ray_flush(RayShared& sh, int n_sensors) {
#else
FTL_HD_NOINLINE void ray_flush(RayShared& sh, int n_sensors) {
#endif
    const RayArrays ra = ray_arrays(&sh, sh.rt, sh.hmax);
    FTL_WARP_SYNC();
    const int ne = sh.ne < kEdgeCap ? sh.ne : kEdgeCap;
#ifdef FTL_DBG_EDGES
    FTL_LANES(lane) {
        if (lane == 0) { sh.dbg[2] += ne; sh.dbg[3] += 1; }
        for (int ei = lane; ei < ne; ei += kLanes) smem_atomic_add(&sh.dbg[7], (int)(sh.e[ei].ax + sh.e[ei].ay + sh.e[ei].bx + sh.e[ei].by));
    }
#endif
    // ---- A2: edges -> (edge, ray) pairs ----------------------------------------------------------------------
    FTL_LANES(lane) {
        for (int ei = lane; ei < ne; ei += kLanes) {
            const RayEdge ed = sh.e[ei];
            const int cls_bit = ed.mask >> 16;
            const float ba = atan2_deg_approx(ed.ay - sh.py, ed.ax - sh.px);
            const float bb = atan2_deg_approx(ed.by - sh.py, ed.bx - sh.px);
            // the follower on an end point of the edge: its bearing is meaningless -> every ray is a candidate
            const bool on_end = fabsf(ed.ax - sh.px) + fabsf(ed.ay - sh.py) < 1e-3f ||
                                fabsf(ed.bx - sh.px) + fabsf(ed.by - sh.py) < 1e-3f;
            const float exlo = fminf(ed.ax, ed.bx), exhi = fmaxf(ed.ax, ed.bx);
            const float eylo = fminf(ed.ay, ed.by), eyhi = fmaxf(ed.ay, ed.by);
            for (int sidx = 0; sidx < n_sensors; sidx++) {
                const RaySensorTab& st = sh.sen[sidx];
                if (!(st.cls_mask & cls_bit)) continue;
                int rows = ed.mask & 0x1ff;
                if (!(rows & (1 << kStaticBit))) rows &= (1 << st.H) - 1;   // rows this sensor keeps
                if (!rows) continue;
                // reach of THIS sensor (the list was culled with the class maximum)
                const float reach = st.L + 1.f;
                if (exhi < sh.px - reach || exlo > sh.px + reach || eyhi < sh.py - reach || eylo > sh.py + reach) continue;
                const float Rf = (float)st.R;
                float ka = (ba - st.theta0) * st.inv_period;
                float d = (bb - ba) * st.inv_period;
                d = d - Rf * rintf(d * st.inv_R);              // short way round, about (-R/2, R/2]
                int klo = (int)ceilf(fminf(ka, ka + d) - st.eps), khi = (int)floorf(fmaxf(ka, ka + d) + st.eps);
                int cnt = khi - klo + 1;
                // the follower (almost) on the edge's line between its end points: the edge subtends ~180 degrees and
                // "the short way round" is ambiguous -> every ray is a candidate
                const bool degenerate = on_end || fabsf(d) > 0.49f * Rf;
                if (cnt <= 0 && !degenerate) continue;
                if (cnt >= st.R || degenerate) { klo = 0; cnt = st.R; }
                int slot = smem_atomic_add(&sh.np, cnt);
                int kk = klo;                                  // klo mod R; |klo| < 3R (bearings and theta0 are bounded)
                while (kk < 0) kk += st.R;
                while (kk >= st.R) kk -= st.R;
                const int packed = (rows << kPairRowsShift) | (ei << kPairEdgeShift);
                for (int k = 0; k < cnt; k++, kk = (kk + 1 == st.R) ? 0 : kk + 1) {
                    if (slot + k < kPairCap)
                        sh.pair[slot + k] = packed | (st.base + kk);
                    else
                        edge_ray_test(sh, ra, st.base + kk, rows, ed.ax, ed.ay, ed.bx, ed.by);   // list full: in place
                }
            }
        }
    }
    FTL_WARP_SYNC();
    // ---- B: uniform pair tests ----------------------------------------------------------------------------------
    const int np = sh.np < kPairCap ? sh.np : kPairCap;
#ifdef FTL_DBG_EDGES
    FTL_LANES(lane) { if (lane == 0) { sh.dbg[4] += np; sh.dbg[5] += sh.np - np; } }
#endif
    FTL_LANES(lane) {
        for (int t = lane; t < np; t += kLanes) {
            const int pr = sh.pair[t], ei = (pr >> kPairEdgeShift) & ((1 << (kPairRowsShift - kPairEdgeShift)) - 1);
            const int f = pr & ((1 << kPairEdgeShift) - 1), rows = pr >> kPairRowsShift;
            const RayEdge ed = sh.e[ei];
            edge_ray_test(sh, ra, f, rows, ed.ax, ed.ay, ed.bx, ed.by);
        }
    }
    FTL_WARP_SYNC();
    FTL_LANES(lane) { if (lane == 0) { sh.ne = 0; sh.np = 0; } }
    FTL_WARP_SYNC();
}

// the whole ray pass of env i.  rot: [rays_total] (cos, sin)(k * 360/R) per flat ray, made on the host.
// The same rows written as ContinuousObserveModifier_sensorPrev's matrix (WRP:203-221): [H][sum of widths], values
// clip(v / laser_length, 0, 1).  Out of line so that the default path's instruction footprint does not grow.
FTL_HD_NOINLINE void ray_rows_write_fused(const DevCfg& cfg, const RayShared& sh, const RayArrays& ra, int i, int lane,
                                          float* rays_out) {
    const FtlConfig& c = cfg.c;
    const int rt = sh.rt, n_valid = sh.n_valid;
    for (int sidx = 0; sidx < sh.ns; sidx++) {
        const FtlRaySensorConfig& sc = c.ray[sidx];
        if (sc.compas) continue;   // written by compas_exact_env
        const int R = sc.lasers_count, H = sc.max_prev_obs, base = sh.sen[sidx].base;
        const float L = (float)sc.laser_length;
        float* dst = rays_out + (size_t)i * cfg.rays_per_env + cfg.ray_out_base[sidx];
        const int stride = cfg.ray_out_stride[sidx];
        const double in_sector = R / 4.0;
        const int* srow = ra.res + ra.hmax * rt + base;
        if (cfg.ray_out_fused_vec4 && (((size_t)rays_out) & 15) == 0) {
            // four consecutive rays of one row per lane, as in the raw float4 path: every sensor's column block starts
            // at a multiple of four floats in a row that is a multiple of four floats long
            const int Q = R >> 2;
            for (int e = lane; e < H * Q; e += kLanes) {
                const int j = e / Q, k = (e - j * Q) << 2, age = H - 1 - j;
                float4 v = make_float4(L, L, L, L);
                if (age < n_valid) {
                    const int4 a = *reinterpret_cast<const int4*>(ra.res + age * rt + base + k);
                    const int4 sb = *reinterpret_cast<const int4*>(srow + k);
                    const int b0 = sb.x < a.x ? sb.x : a.x, b1 = sb.y < a.y ? sb.y : a.y;
                    const int b2 = sb.z < a.z ? sb.z : a.z, b3 = sb.w < a.w ? sb.w : a.w;
                    if (b0 != kNoHitBits) v.x = i2f_bits(b0);
                    if (b1 != kNoHitBits) v.y = i2f_bits(b1);
                    if (b2 != kNoHitBits) v.z = i2f_bits(b2);
                    if (b3 != kNoHitBits) v.w = i2f_bits(b3);
                }
                v.x = ray_out_value(c, L, v.x); v.y = ray_out_value(c, L, v.y);
                v.z = ray_out_value(c, L, v.z); v.w = ray_out_value(c, L, v.w);
                *reinterpret_cast<float4*>(dst + j * stride + k) = v;
            }
            continue;
        }
        for (int e = lane; e < H * R; e += kLanes) {
            const int j = e / R, k = e - j * R, age = H - 1 - j;
            float v = L;
            if (age < n_valid) {
                int bits = ra.res[age * rt + base + k], sb = srow[k];
                bits = sb < bits ? sb : bits;
                if (bits != kNoHitBits) v = i2f_bits(bits);
            }
            v = ray_out_value(c, L, v);
            if (!sc.pad_sectors) {
                dst[j * stride + k] = v;
            } else {   // SEN:932-953
                int ksec = (k < in_sector) ? 0 : (k < 2 * in_sector) ? 1 : (k < 3 * in_sector) ? 2 : 3;
                for (int sec = 0; sec < 4; sec++) dst[j * stride + sec * R + k] = sec == ksec ? v : 0.f;
            }
        }
    }
}

FTL_HD void rays_warp(const DevCfg& cfg, const DevState& s, const DevPool& pool, const double2* rot, int i,
                      RayShared& sh, float* rays_out) {
    const FtlConfig& c = cfg.c;
    const int rt = cfg.rays_total;
    const RayArrays ra = ray_arrays(&sh, rt, cfg.ray_hmax);
    const int NBr = s.n_bears, ns = c.n_ray_sensors;
#if defined(__CUDA_ARCH__) && !defined(FTL_NO_RAYS_PREFETCH)
    // The env's lists are reached through a chain of dependent loads (push counter -> stored ranges -> corridor entries,
    // snapshot rectangles); their ADDRESSES only depend on the env index, so the corridor ring's lines (corridor_cap entries of 16
    // bytes) are requested now, while the setup below runs (k_rays 0.2451 -> 0.2431 ms).
    {
        const int pl = (int)(threadIdx.x & (kLanes - 1));
        const char* ring = reinterpret_cast<const char*>(s.corridor + (size_t)i * c.corridor_cap);
        for (int k = pl; k * 128 < c.corridor_cap * 16; k += kLanes) asm volatile("prefetch.global.L1 [%0];" ::"l"(ring + k * 128));
#ifdef FTL_RECT_PREFETCH   // measured: the rectangle history too makes the kernel slower (0.2431 -> 0.2454 ms)
        for (int k = pl; k < FTL_MAX_HIST * (1 + NBr); k += kLanes)
            asm volatile("prefetch.global.L1 [%0];" ::"l"(s.snap_rect + (size_t)k * s.n + i));
#endif
    }
#endif
    const double dir = s.rd[(size_t)RD_DIR * s.n + i];
    // ---- setup: lane 0 the scalars, lanes < ns the sensor tables (static part from DevCfg), lanes < 5 the class
    //      reaches, the last eight lanes the stored corridor ranges -------------------------------------------------------
    FTL_LANES(lane) {
        const int pushes = s.gi[(size_t)GI_SNAP_PUSHES * s.n + i];
        if (lane == 0) {
            float2 p = s.pos[i];
            sh.px = p.x; sh.py = p.y;
            sh.dir = dir;
            sh.scenario = s.gi[(size_t)GI_SCENARIO * s.n + i];
            sh.snap_pushes = pushes;
#ifdef FTL_RAYS_ALL_AGES   // measured (r02_ab_log.txt (25)): the snapshot ring keeps FTL_MAX_HIST entries, no sensor looks further back than ray_hmax
            sh.n_valid = pushes < FTL_MAX_HIST ? pushes : FTL_MAX_HIST;
#else
            sh.n_valid = pushes < cfg.ray_hmax ? pushes : cfg.ray_hmax;   // ages no sensor keeps are not cast at all
#endif
            sh.ne = 0; sh.np = 0; sh.nu = 0; sh.rt = rt; sh.ns = ns; sh.hmax = cfg.ray_hmax;
#ifdef FTL_DBG_EDGES
            for (int k = 0; k < 16; k++) sh.dbg[k] = 0;
#endif
            sh.unc = s.unc_rec + (size_t)i * kUncPerEnv;
        }
        if (lane < ns) {
            const RaySensorStatic& ss = cfg.ray_static[lane];
            RaySensorTab st;
            st.base = ss.base; st.R = ss.R; st.H = ss.H; st.cls_mask = ss.cls_mask;
            st.L = ss.L; st.inv_period = ss.inv_period; st.eps = ss.eps; st.inv_R = ss.inv_R; st.pad_ = 0.f;
            const double a0 = dir + c.ray[lane].first_laser_angle_offset;
            st.theta0 = (float)a0;
            sincos_deg(a0, &st.sn0, &st.cs0);
            sh.sen[lane] = st;
        }
        if (lane < EC_COUNT) sh.reach[lane] = cfg.ray_reach[lane];
#if defined(__CUDA_ARCH__) && !defined(FTL_NO_RAYS_PREFETCH)
        {   // every slot of the range ring is loaded at once (no wait for the push counter), then handed to its age
            constexpr int kR0 = kLanes - FTL_MAX_HIST;   // the last FTL_MAX_HIST lanes of the env's lane group
            int2 rg_slot = make_int2(0, 0);
            if (lane >= kR0) rg_slot = s.snap_range[(size_t)(lane - kR0) * s.n + i];
            const int age = lane - kR0;
            const int src = kR0 + (((pushes - 1 - age) % FTL_MAX_HIST) + FTL_MAX_HIST) % FTL_MAX_HIST;
            const int src_abs = (src & (kLanes - 1)) + (int)(threadIdx.x & (32 - kLanes));
            const int rx = __shfl_sync(FTL_LANE_MASK(), rg_slot.x, src_abs), ry = __shfl_sync(FTL_LANE_MASK(), rg_slot.y, src_abs);
            if (lane >= kR0) {
                const bool live = age < pushes;
                sh.tail[age] = live ? rx : 0; sh.head[age] = live ? ry : 0;
            }
        }
#else
        if (lane >= kLanes - FTL_MAX_HIST) {
            const int age = lane - (kLanes - FTL_MAX_HIST);
            int2 rg = make_int2(0, 0);
            if (age < pushes) rg = s.snap_range[(size_t)((pushes - 1 - age) % FTL_MAX_HIST) * s.n + i];
            sh.tail[age] = rg.x; sh.head[age] = rg.y;
        }
#endif
    }
    FTL_WARP_SYNC();
    FTL_LANES(lane) {
        int sidx = 0;
        for (int f = lane; f < rt; f += kLanes) {
            while (sidx + 1 < ns && f >= sh.sen[sidx + 1].base) sidx++;
            const FtlRaySensorConfig& sc = c.ray[sidx];
            const double cs0 = sh.sen[sidx].cs0, sn0 = sh.sen[sidx].sn0;
            double2 r = rot[f];   // (cos, sin) of k * period
            ra.dx[f] = (float)((cs0 * r.x - sn0 * r.y) * sc.laser_length);
            ra.dy[f] = (float)((sn0 * r.x + cs0 * r.y) * sc.laser_length);
            ra.len[f] = (float)sc.laser_length;
#ifdef FTL_RES_INIT_PER_RAY
            for (int a = 0; a <= ra.hmax; a++) ra.res[a * rt + f] = kNoHitBits;
#endif
        }
#ifndef FTL_RES_INIT_PER_RAY   // measured (r02_ab_log.txt (30)): the rows of minima as 16-byte stores (288 words: three rounds, not twelve stores per lane)
        const int n_res = (ra.hmax + 1) * rt;
        if ((rt & 3) == 0) {   // res starts 12 * rt bytes behind a 16-byte boundary
            int4* r4 = reinterpret_cast<int4*>(ra.res);
            for (int k = lane; k < (n_res >> 2); k += kLanes) r4[k] = make_int4(kNoHitBits, kNoHitBits, kNoHitBits, kNoHitBits);
        } else {
            for (int k = lane; k < n_res; k += kLanes) ra.res[k] = kNoHitBits;
        }
#endif
    }
    FTL_WARP_SYNC();
    const int n_valid = sh.n_valid;
    if (n_valid > 0) {
        // ---- A1: rectangles -----------------------------------------------------------------------------
        const int4* statics = pool.static_rects + (size_t)sh.scenario * c.static_cap;
        const int n_static = pool.n_static[sh.scenario];
        const int n_dyn = n_valid * (1 + NBr);
#ifdef FTL_RAYS_SPLIT_RECT_ROUNDS   // measured (r02_ab_log.txt (26)): static and stored rectangles in rounds of their own
        if (sh.reach[EC_STATIC] > 0.f) {
            for (int q0 = 0; q0 < n_static; q0 += kLanes) {
                FTL_WARP_SYNC();
                if (FTL_UNIFORM_INT(sh.ne) + 4 * kLanes > kEdgeCap) ray_flush(sh, ns);   // a round adds at most 4 edges per lane
                FTL_LANES(lane) {
                    if (q0 + lane < n_static) rect_append(sh, statics[q0 + lane], EC_STATIC, 1 << kStaticBit);
                }
            }
        }
        for (int q0 = 0; q0 < n_dyn; q0 += kLanes) {
            FTL_WARP_SYNC();
            if (FTL_UNIFORM_INT(sh.ne) + 4 * kLanes > kEdgeCap) ray_flush(sh, ns);
            FTL_LANES(lane) {
                int q = q0 + lane;
                if (q < n_dyn) {
                    int age = q / (1 + NBr), k = q - age * (1 + NBr);
                    int cls = k == 0 ? EC_LEADER : EC_BEAR;
                    if (sh.reach[cls] > 0.f) {
                        int slot = (sh.snap_pushes - 1 - age) % FTL_MAX_HIST;
                        rect_append(sh, s.snap_rect[((size_t)slot * (1 + NBr) + k) * s.n + i], cls, 1 << age);
                    }
                }
            }
        }
#else
        {   // static rectangles and the stored leader / bear rectangles as ONE item list: 37 + 10 items are two rounds, not three
            const int ns_on = sh.reach[EC_STATIC] > 0.f ? n_static : 0;
            const int n_rect = ns_on + n_dyn;
            for (int q0 = 0; q0 < n_rect; q0 += kLanes) {
                FTL_WARP_SYNC();
                if (FTL_UNIFORM_INT(sh.ne) + 4 * kLanes > kEdgeCap) ray_flush(sh, ns);   // a round adds at most 4 edges per lane
                FTL_LANES(lane) {
                    const int q = q0 + lane;
                    if (q < n_rect) {
                        const int4* src = statics + q;
                        int cls = EC_STATIC, rows = 1 << kStaticBit;
                        if (q >= ns_on) {
                            const int d = q - ns_on, age = d / (1 + NBr), k = d - age * (1 + NBr);
                            const int slot = (sh.snap_pushes - 1 - age) % FTL_MAX_HIST;
                            cls = k == 0 ? EC_LEADER : EC_BEAR;
                            rows = 1 << age;
                            src = s.snap_rect + ((size_t)slot * (1 + NBr) + k) * s.n + i;
                        }
                        if (sh.reach[cls] > 0.f) rect_append(sh, *src, cls, rows);
                    }
                }
            }
        }
#endif
        // ---- A1: corridor sides (union of the stored ranges) and end caps; the list is only flushed when the next
        //      batch might not fit ----------------------------------------------------------------------------------
        const float4* corr = s.corridor + (size_t)i * c.corridor_cap;
        const int cmask = c.corridor_cap - 1;
        if (sh.reach[EC_CORRIDOR] > 0.f) {
            int min_tail = sh.tail[0], max_head = sh.head[0];
            for (int a = 1; a < n_valid; a++) {
                min_tail = sh.tail[a] < min_tail ? sh.tail[a] : min_tail;
                max_head = sh.head[a] > max_head ? sh.head[a] : max_head;
            }
            for (int q0 = min_tail; q0 < max_head - 1; q0 += kCorridorChunk) {
                FTL_WARP_SYNC();
                if (FTL_UNIFORM_INT(sh.ne) + 2 * kCorridorChunk > kEdgeCap) ray_flush(sh, ns);
                FTL_LANES(lane) {
                    int q1 = q0 + kCorridorChunk < max_head - 1 ? q0 + kCorridorChunk : max_head - 1;
                    for (int q = q0 + lane; q < q1; q += kLanes) {
                        int rows = 0;
                        for (int a = 0; a < n_valid; a++)
                            if (q >= sh.tail[a] && q < sh.head[a] - 1) rows |= 1 << a;
                        if (rows) {
                            float4 a4 = corr[q & cmask], b4 = corr[(q + 1) & cmask];
                            seg_append(sh, a4.x, a4.y, b4.x, b4.y, EC_CORRIDOR, rows);   // (one reservation for both sides: slower)
                            seg_append(sh, a4.z, a4.w, b4.z, b4.w, EC_CORRIDOR, rows);
                        }
                    }
                }
            }
        }
        if (sh.reach[EC_CAP] > 0.f) {
            FTL_WARP_SYNC();
            if (FTL_UNIFORM_INT(sh.ne) + 2 * FTL_MAX_HIST > kEdgeCap) ray_flush(sh, ns);
            FTL_LANES(lane) {
                if (lane < 2 * n_valid) {   // SEN:648-650
                    int age = lane >> 1;
                    int idx = (lane & 1) ? sh.head[age] - 1 : sh.tail[age];
                    float4 a4 = corr[idx & cmask];
                    seg_append(sh, a4.x, a4.y, a4.z, a4.w, EC_CAP, 1 << age);
                }
            }
        }
        ray_flush(sh, ns);
    }
    // ---- out: assemble rows and write (row-major per sensor; consecutive lanes write consecutive floats) ------
    FTL_LANES(lane) {
#ifndef FTL_NO_FUSED
        if (c.fused_sensor_prev) {
            ray_rows_write_fused(cfg, sh, ra, i, lane, rays_out);
        } else
#endif
#ifndef FTL_NO_VEC4_OUT
        if (cfg.ray_out_vec4 && (((size_t)rays_out) & 15) == 0) {
            // four consecutive rays of one row per lane: two 16-byte shared loads, one 16-byte store
#ifndef FTL_RAYS_OUT_PER_SENSOR
            // one item list over all sensors (15 + 45 float4 items are two rounds, not three): a sensor block is [H][R] and the
            // blocks follow each other, so item t of the env simply lands at float4 index t
            int T = 0;
            for (int sidx = 0; sidx < ns; sidx++) T += sh.sen[sidx].H * (sh.sen[sidx].R >> 2);
            float* dst = rays_out + (size_t)i * cfg.rays_per_env;
            for (int t = lane; t < T; t += kLanes) {
                int sidx = 0, s0 = 0;
                for (; sidx + 1 < ns; sidx++) {
                    const int cnt = sh.sen[sidx].H * (sh.sen[sidx].R >> 2);
                    if (t < s0 + cnt) break;
                    s0 += cnt;
                }
                const RaySensorTab& st = sh.sen[sidx];
                const int Q = st.R >> 2, e = t - s0;
                const int j = (int)(((float)e + 0.5f) * (4.f * st.inv_R));   // e / Q without an integer division (H*Q < 2^20)
                const int k = (e - j * Q) << 2;
                const int age = st.H - 1 - j;
                const float L = st.L;
                float4 v = make_float4(L, L, L, L);
                if (age < n_valid) {
                    const int4 a = *reinterpret_cast<const int4*>(ra.res + age * rt + st.base + k);
                    const int4 sb = *reinterpret_cast<const int4*>(ra.res + ra.hmax * rt + st.base + k);
                    const int b0 = sb.x < a.x ? sb.x : a.x, b1 = sb.y < a.y ? sb.y : a.y;
                    const int b2 = sb.z < a.z ? sb.z : a.z, b3 = sb.w < a.w ? sb.w : a.w;
                    if (b0 != kNoHitBits) v.x = i2f_bits(b0);
                    if (b1 != kNoHitBits) v.y = i2f_bits(b1);
                    if (b2 != kNoHitBits) v.z = i2f_bits(b2);
                    if (b3 != kNoHitBits) v.w = i2f_bits(b3);
                }
                *reinterpret_cast<float4*>(dst + 4 * t) = v;
            }
#else
            int off = 0;
            for (int sidx = 0; sidx < ns; sidx++) {
                const FtlRaySensorConfig& sc = c.ray[sidx];
                const int R = sc.lasers_count, H = sc.max_prev_obs, base = sh.sen[sidx].base, Q = R >> 2;
                const float L = (float)sc.laser_length;
                float* dst = rays_out + (size_t)i * cfg.rays_per_env + off;
                const float inv_Q = 1.0f / (float)Q;
                const int* srow = ra.res + ra.hmax * rt + base;
                for (int e = lane; e < H * Q; e += kLanes) {
                    const int j = (int)(((float)e + 0.5f) * inv_Q);   // e / Q without an integer division
                    const int k = (e - j * Q) << 2;
                    const int age = H - 1 - j;
                    float4 v = make_float4(L, L, L, L);
                    if (age < n_valid) {
                        const int4 a = *reinterpret_cast<const int4*>(ra.res + age * rt + base + k);
                        const int4 sb = *reinterpret_cast<const int4*>(srow + k);
                        const int b0 = sb.x < a.x ? sb.x : a.x, b1 = sb.y < a.y ? sb.y : a.y;
                        const int b2 = sb.z < a.z ? sb.z : a.z, b3 = sb.w < a.w ? sb.w : a.w;
                        if (b0 != kNoHitBits) v.x = i2f_bits(b0);
                        if (b1 != kNoHitBits) v.y = i2f_bits(b1);
                        if (b2 != kNoHitBits) v.z = i2f_bits(b2);
                        if (b3 != kNoHitBits) v.w = i2f_bits(b3);
                    }
                    *reinterpret_cast<float4*>(dst + j * R + k) = v;
                }
                off += H * R;
            }
#endif
        } else
#endif
        {
            int off = 0;
            for (int sidx = 0; sidx < ns; sidx++) {
                const FtlRaySensorConfig& sc = c.ray[sidx];
                if (sc.compas) { off += sensor_width(sc); continue; }   // written by compas_exact_env
                const int R = sc.lasers_count, H = sc.max_prev_obs, base = sh.sen[sidx].base;
                const float L = (float)sc.laser_length;
                float* dst = rays_out + (size_t)i * cfg.rays_per_env + off;
                const int nsec = sc.pad_sectors ? 4 : 1;
                const double in_sector = R / 4.0;
                const float inv_R = 1.0f / (float)R;
                const int* srow = ra.res + ra.hmax * rt + base;
                for (int e = lane; e < H * R; e += kLanes) {
                    int j = (int)(((float)e + 0.5f) * inv_R);      // e / R without an integer division (H*R < 2^20)
                    int k = e - j * R;
                    const int age = H - 1 - j;
                    float v = L;
                    if (age < n_valid) {
                        int bits = ra.res[age * rt + base + k], sb = srow[k];
                        bits = sb < bits ? sb : bits;
                        if (bits != kNoHitBits) v = i2f_bits(bits);
                    }
                    if (nsec == 1) {
                        dst[e] = v;
                    } else {   // SEN:932-953: four sector-masked copies side by side
                        int ksec = (k < in_sector) ? 0 : (k < 2 * in_sector) ? 1 : (k < 3 * in_sector) ? 2 : 3;
                        for (int sec = 0; sec < 4; sec++) dst[(j * 4 + sec) * R + k] = sec == ksec ? v : 0.f;
                    }
                }
                off += sensor_width(sc);
            }
        }
        if (lane == 0) s.unc_count[i] = sh.nu;   // 0 almost always; > kUncPerEnv: recast the env exactly
#ifdef FTL_DBG_EDGES
        if (lane == 0 && i == FTL_DBG_EDGES)
            printf("DBG env %d: appended %d in_place %d listed %d flushes %d pairs %d pairs_in_place %d sum_app %d sum_listed %d nu %d by class %d %d %d %d %d\n", i,
                   sh.dbg[0], sh.dbg[1], sh.dbg[2], sh.dbg[3], sh.dbg[4], sh.dbg[5], sh.dbg[6], sh.dbg[7], sh.nu,
                   sh.dbg[8], sh.dbg[9], sh.dbg[10], sh.dbg[11], sh.dbg[12]);
#endif
    }
    FTL_WARP_SYNC();
}

// =====================================================================================================
// Exact pass (one thread per env): pairs whose float32 predicates were inconclusive are redone with the
// reference's float64 arithmetic and merged into the already written rows; an env with more such pairs than
// fit its record slice is recast from scratch, edge list by edge list, like the reference does.
// =====================================================================================================
struct ExactEnv {
    float px, py;
    double dir;
    int n_valid, pushes, scenario;
};

FTL_HD void exact_ray_end(const FtlRaySensorConfig& sc, const ExactEnv& ee, int k, double* ex, double* ey) {
    double ang = (ee.dir + sc.first_laser_angle_offset) + ray_angle(sc, k);   // SEN:888-891
    double sn, cs;
    sincos_deg(ang, &sn, &cs);
    *ex = (double)ee.px + cs * sc.laser_length;
    *ey = (double)ee.py + sn * sc.laser_length;
}

// index of (row j, ray k) relative to the sensor's output base; pad_sectors layout of SEN:932-953
FTL_HD int ray_out_index(const FtlRaySensorConfig& sc, int stride, int j, int k) {
    const int R = sc.lasers_count;
    if (!sc.pad_sectors) return j * stride + k;
    double in_sector = R / 4.0;
    int ksec = (k < in_sector) ? 0 : (k < 2 * in_sector) ? 1 : (k < 3 * in_sector) ? 2 : 3;
    return j * stride + ksec * R + k;
}

FTL_HD float exact_rect_min(const ExactEnv& ee, double ex, double ey, int4 q) {
    float l = (float)q.x, t = (float)q.y, r = (float)(q.x + q.z), b = (float)(q.y + q.w);
    float m = seg_hit_exact(ee.px, ee.py, ex, ey, l, b, r, b);
    m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, r, t, r, b));
    m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, r, t, l, t));
    return fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, l, b, l, t));
}

FTL_HD_NOINLINE void rays_exact_recast(const DevCfg& cfg, const DevState& s, const DevPool& pool, int i,
                                       const ExactEnv& ee, float* rays_out) {
    const FtlConfig& c = cfg.c;
    const int NBr = s.n_bears, cmask = c.corridor_cap - 1;
    const float4* corr = s.corridor + (size_t)i * c.corridor_cap;
    const int4* statics = pool.static_rects + (size_t)ee.scenario * c.static_cap;
    const int n_static = pool.n_static[ee.scenario];
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) {
        const FtlRaySensorConfig& sc = c.ray[sidx];
        if (sc.compas) continue;
        const int cls = sensor_class_mask(sc);
        const int stride = cfg.ray_out_stride[sidx];
        float* dst = rays_out + (size_t)i * cfg.rays_per_env + cfg.ray_out_base[sidx];
        for (int k = 0; k < sc.lasers_count; k++) {
            double ex, ey;
            exact_ray_end(sc, ee, k, &ex, &ey);
            float m_static = kNoHit;
            if (cls & (1 << EC_STATIC))
                for (int q = 0; q < n_static; q++) m_static = fminf(m_static, exact_rect_min(ee, ex, ey, statics[q]));
            for (int j = 0; j < sc.max_prev_obs; j++) {
                const int age = sc.max_prev_obs - 1 - j;
                float m = kNoHit;
                if (age < ee.n_valid) {
                    m = m_static;
                    const int slot = (ee.pushes - 1 - age) % FTL_MAX_HIST;
                    const int2 rg = s.snap_range[(size_t)slot * s.n + i];
                    for (int r = 0; r < 1 + NBr; r++)
                        if (cls & (1 << (r == 0 ? EC_LEADER : EC_BEAR)))
                            m = fminf(m, exact_rect_min(ee, ex, ey, s.snap_rect[((size_t)slot * (1 + NBr) + r) * s.n + i]));
                    if (cls & (1 << EC_CORRIDOR))
                        for (int q = rg.x; q < rg.y - 1; q++) {
                            float4 a4 = corr[q & cmask], b4 = corr[(q + 1) & cmask];
                            m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, a4.x, a4.y, b4.x, b4.y));
                            m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, a4.z, a4.w, b4.z, b4.w));
                        }
                    if (cls & (1 << EC_CAP)) {
                        float4 a4 = corr[rg.x & cmask], b4 = corr[(rg.y - 1) & cmask];
                        m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, a4.x, a4.y, a4.z, a4.w));
                        m = fminf(m, seg_hit_exact(ee.px, ee.py, ex, ey, b4.x, b4.y, b4.z, b4.w));
                    }
                }
                dst[ray_out_index(sc, stride, j, k)] =
                    ray_out_value(c, (float)sc.laser_length, m >= kNoHit ? (float)sc.laser_length : m);
            }
        }
    }
}

// LeaderCorridor_lasers_compas.scan (SEN:1138-1240) for one env, with the reference's own float64 sequence: every ray
// of every stored corridor (front cap, back cap, left walls, right walls -- the order of corridor_lines, which decides
// ties through np.argmin) is tested with all-float64 ccw predicates, the nearest hit picks the column block.  One
// thread per env: this sensor is outside the shipped configurations and is not worth a place in the warp kernel.
// (The reference keeps its corridor points in float64; the ring here holds them rounded to float32 -- the distances
// agree to ~1e-6 relative, a ray that grazes a corridor vertex within ~3e-5 px may decide differently.)
struct CompasHit { double sel, x, y; int orient; };
FTL_HD void compas_test(const ExactEnv& ee, double ex, double ey, float fax, float fay, float fbx, float fby, int orient,
                        CompasHit* best) {
    const double ax = fax, ay = fay, bx = fbx, by = fby, cx = ee.px, cy = ee.py;
    // ccw(A, B, C) = (C.y - A.y) * (B.x - A.x) > (B.y - A.y) * (C.x - A.x), SEN:608-609, everything float64
    const bool acd = (ey - ay) * (cx - ax) > (cy - ay) * (ex - ax);
    const bool bcd = (ey - by) * (cx - bx) > (cy - by) * (ex - bx);
    if (acd == bcd) return;
    const bool abc = (cy - ay) * (bx - ax) > (by - ay) * (cx - ax);
    const bool abd = (ey - ay) * (bx - ax) > (by - ay) * (ex - ax);
    if (abc == abd) return;
    // seg_intersect(a1 = A, a2 = B, b1 = pos, b2 = end), SEN:628-640
    const double dax = bx - ax, day = by - ay, dbx = ex - cx, dby = ey - cy, dpx = ax - cx, dpy = ay - cy;
    const double dapx = -day, dapy = dax;
    const double denom = fma(dapx, dbx, dapy * dby);   // np.dot((k,2), (2,1))
    const double m0 = dapx * dpx, m1 = dapy * dpy;
    const double t = (m0 + m1) / denom;
    const double xx = t * dbx + cx, xy = t * dby + cy;
    const double ddx = xx - cx, ddy = xy - cy;
    const double sx = ddx * ddx, sy = ddy * ddy;
    const double sel = sqrt(sx + sy);                  // np.linalg.norm(x - pos, axis=1)
    if (best->orient < 0 || sel < best->sel) { best->sel = sel; best->x = xx; best->y = xy; best->orient = orient; }
}

FTL_HD_NOINLINE void compas_exact_env(const DevCfg& cfg, const DevState& s, int i, float* rays_out) {
    const FtlConfig& c = cfg.c;
    ExactEnv ee;
    float2 p = s.pos[i];
    ee.px = p.x; ee.py = p.y;
    ee.dir = s.rd[(size_t)RD_DIR * s.n + i];
    ee.pushes = s.gi[(size_t)GI_SNAP_PUSHES * s.n + i];
    ee.n_valid = ee.pushes < FTL_MAX_HIST ? ee.pushes : FTL_MAX_HIST;
    ee.scenario = 0;
    const int cmask = c.corridor_cap - 1;
    const float4* corr = s.corridor + (size_t)i * c.corridor_cap;
    for (int sidx = 0; sidx < c.n_ray_sensors; sidx++) {
        const FtlRaySensorConfig& sc = c.ray[sidx];
        if (!sc.compas) continue;
        const int R = sc.lasers_count, H = sc.max_prev_obs, stride = cfg.ray_out_stride[sidx];
        const float L = (float)sc.laser_length;
        float* dst = rays_out + (size_t)i * cfg.rays_per_env + cfg.ray_out_base[sidx];
        for (int j = 0; j < H; j++) {
            const int age = H - 1 - j;
            int2 rg = make_int2(0, 0);
            if (age < ee.n_valid) rg = s.snap_range[(size_t)((ee.pushes - 1 - age) % FTL_MAX_HIST) * s.n + i];
            float* row = dst + (size_t)j * stride;
            for (int k = 0; k < 5 * R; k++) row[k] = 0.f;
            for (int k = 0; k < R; k++) {
                double ex, ey;
                exact_ray_end(sc, ee, k, &ex, &ey);
                CompasHit best = {0.0, ex, ey, -1};
                if (age < ee.n_valid && rg.y - rg.x > 1) {
                    const float4 f4 = corr[(rg.y - 1) & cmask], b4 = corr[rg.x & cmask];
                    compas_test(ee, ex, ey, f4.x, f4.y, f4.z, f4.w, 0, &best);   // front wall = corridor[-1]
                    compas_test(ee, ex, ey, b4.x, b4.y, b4.z, b4.w, 1, &best);   // back wall = corridor[0]
                    for (int q = rg.x; q < rg.y - 1; q++) {                      // left walls: corridor[q][1] -> corridor[q + 1][1]
                        const float4 a4 = corr[q & cmask], n4 = corr[(q + 1) & cmask];
                        compas_test(ee, ex, ey, a4.z, a4.w, n4.z, n4.w, 2, &best);
                    }
                    for (int q = rg.x; q < rg.y - 1; q++) {                      // right walls: corridor[q][0] -> corridor[q + 1][0]
                        const float4 a4 = corr[q & cmask], n4 = corr[(q + 1) & cmask];
                        compas_test(ee, ex, ey, a4.x, a4.y, n4.x, n4.y, 3, &best);
                    }
                }
                // obs_item[...] = np.linalg.norm(collide - pos): 1-D -> sqrt(dot(x, x)), the FMA form
                const double ddx = best.x - (double)ee.px, ddy = best.y - (double)ee.py;
                const float v = ray_out_value(c, L, (float)sqrt(fma(ddy, ddy, ddx * ddx)));
                row[best.orient < 0 ? k : k + R * (1 + best.orient)] = v;
            }
        }
    }
}

FTL_HD void rays_exact_env(const DevCfg& cfg, const DevState& s, const DevPool& pool, int i, float* rays_out) {
    if (cfg.ray_compas_mask) compas_exact_env(cfg, s, i, rays_out);
    const int count = s.unc_count[i];
    if (count == 0) return;
    const FtlConfig& c = cfg.c;
    ExactEnv ee;
    float2 p = s.pos[i];
    ee.px = p.x; ee.py = p.y;
    ee.dir = s.rd[(size_t)RD_DIR * s.n + i];
    ee.pushes = s.gi[(size_t)GI_SNAP_PUSHES * s.n + i];
    ee.n_valid = ee.pushes < FTL_MAX_HIST ? ee.pushes : FTL_MAX_HIST;
    ee.scenario = s.gi[(size_t)GI_SCENARIO * s.n + i];
    if (count > kUncPerEnv) {
        rays_exact_recast(cfg, s, pool, i, ee, rays_out);
        return;
    }
    const UncRec* rec = s.unc_rec + (size_t)i * kUncPerEnv;
    for (int t = 0; t < count; t++) {
        const UncRec r = rec[t];
        int sidx = 0, base = 0;
        while (sidx + 1 < c.n_ray_sensors && r.f >= base + c.ray[sidx].lasers_count) {
            base += c.ray[sidx].lasers_count;
            sidx++;
        }
        const FtlRaySensorConfig& sc = c.ray[sidx];
        double ex, ey;
        exact_ray_end(sc, ee, r.f - base, &ex, &ey);
        float d = seg_hit_exact(ee.px, ee.py, ex, ey, r.ax, r.ay, r.bx, r.by);
        if (d >= kNoHit) continue;
        d = ray_out_value(c, (float)sc.laser_length, d);   // monotone, so the minimum can be taken on the written scale
        float* dst = rays_out + (size_t)i * cfg.rays_per_env + cfg.ray_out_base[sidx];
        for (int j = 0; j < sc.max_prev_obs; j++) {
            const int age = sc.max_prev_obs - 1 - j;
            if (age >= ee.n_valid) continue;
            if (!((r.rows >> age) & 1) && !((r.rows >> kStaticBit) & 1)) continue;
            float* cell = dst + ray_out_index(sc, cfg.ray_out_stride[sidx], j, r.f - base);
            if (d < *cell) *cell = d;
        }
    }
}

}  // namespace ftl
