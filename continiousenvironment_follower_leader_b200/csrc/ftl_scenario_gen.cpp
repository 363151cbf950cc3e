// ftl_scenario_gen.cpp -- host-side scenario generation (no CUDA): what Game.reset() does before the first sensor
// scan, for many seeds at once on all host cores, so that scenario pools for 1M-env batches do not go through Python.
//
// Follows follow_the_leader_continuous_env.py: _create_robots (ENV:545-596), _create_obstacles (ENV:613-677),
// generate_finish_point (ENV:1614-1630, distance_to_rect of utils/misc.py:29-44), _pos_follower_behind_leader
// (ENV:598-611), and the grids of generate_trajectory_dstar (ENV:1493-1507) / generate_trajectory_astar
// (ENV:1632-1712).  The reference draws from python's global `random` after env.seed(s) (ENV:429-432); the draws are
// reproduced here with CPython's generator (MT19937 seeded by init_by_array, randrange -> _randbelow_with_getrandbits),
// one private generator per scenario, in the reference's call order -- the layout of seed s (leader start, rocks,
// finish point, follower distance) is the one the reference builds.  Routes are shortest 8-connected paths on the
// reference's grid with its obstacle inflation; the reference's D* breaks ties by iterating a python set of objects
// (utils/dstar.py:88,130), so its routes are matched as valid shortest routes, not waypoint for waypoint.
// The python twin of this file is scenario_gen.py; tests/test_scenario_gen_native.py demands identical pools.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ftl.h"

namespace {

// ---- CPython's random.Random for integer seeds (Modules/_randommodule.c, Lib/random.py) -------------------------
struct PyRandom {
    uint32_t mt[624];
    int idx;
    void init_genrand(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    void init_by_array(const uint32_t* key, int len) {
        init_genrand(19650218u);
        int i = 1, j = 0;
        for (int k = (624 > len ? 624 : len); k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
            i++; j++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
            if (j >= len) j = 0;
        }
        for (int k = 623; k; k--) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            i++;
            if (i >= 624) { mt[0] = mt[623]; i = 1; }
        }
        mt[0] = 0x80000000u;
    }
    void seed(int64_t a) {  // random.seed(int): abs value as little-endian 32-bit words, at least one
        uint64_t v = a < 0 ? (uint64_t)(-(a + 1)) + 1u : (uint64_t)a;
        uint32_t key[2] = {(uint32_t)(v & 0xffffffffu), (uint32_t)(v >> 32)};
        init_by_array(key, key[1] ? 2 : 1);
    }
    uint32_t next_u32() {
        if (idx >= 624) {
            int kk;
            for (kk = 0; kk < 624 - 397; kk++) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            for (; kk < 623; kk++) {
                uint32_t y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            uint32_t y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    uint32_t randbelow(uint32_t n) {  // _randbelow_with_getrandbits, n in [1, 2^31]
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) k++;
        uint32_t r;
        do r = next_u32() >> (32 - k); while (r >= n);
        return r;
    }
    // random.randrange(start, stop, step) for positive steps; false = the ValueError("empty range")
    bool randrange(long start, long stop, long step, long* out) {
        long width = stop - start;
        long n = step == 1 ? width : (width + step - 1) / step;
        if (n <= 0) return false;
        *out = start + step * (long)randbelow((uint32_t)n);
        return true;
    }
};

const double kPi = 3.141592653589793;
double py_radians(double x) { return x * (kPi / 180.0); }
double py_degrees(double x) { return x * (180.0 / kPi); }
double angle_correction(double a) { return a >= 360 ? a - 360 : a < 0 ? 360 + a : a; }  // utils/misc.py:6-13
double angle_to_point(double cx, double cy, double tx, double ty) {                     // utils/misc.py:16-26
    double rx = tx - cx, ry = ty - cy, res;
    if (rx > 0) res = py_degrees(std::atan(ry / rx));
    else if (rx < 0) res = py_degrees(std::atan(ry / rx)) + 180;
    else res = 0;
    return angle_correction(res);
}

struct Rect { int x, y, w, h; };
// image.get_rect(center=(cx, cy)) after transform.scale(image, (w, h)), CLS:42-50: truncating centre
Rect rect_centered(double cx, double cy, int w, int h) { return {(int)cx - (w >> 1), (int)cy - (h >> 1), w, h}; }
bool collidepoint(const Rect& r, long px, long py) { return r.x <= px && px < r.x + r.w && r.y <= py && py < r.y + r.h; }
double distance_to_rect(long px, long py, const Rect& r) {  // utils/misc.py:29-44: 4 corners + 4 edge midpoints
    const int qx[8] = {r.x, r.x, r.x + r.w, r.x + r.w, r.x + (r.w >> 1), r.x, r.x + (r.w >> 1), r.x + r.w};
    const int qy[8] = {r.y, r.y + r.h, r.y, r.y + r.h, r.y, r.y + (r.h >> 1), r.y + r.h, r.y + (r.h >> 1)};
    double best = 1e300;
    for (int k = 0; k < 8; k++) best = std::min(best, std::hypot((double)(px - qx[k]), (double)(py - qy[k])));
    return best;
}

// 8-connected Dijkstra with euclidean step costs; the queue is ordered by (distance, x, y) like python's heap of
// (d, (x, y)) tuples, so both implementations expand cells in the same order and pick the same parents
struct Cell { int x, y; };
bool shortest_path(const std::vector<uint8_t>& blocked, int nx, int ny, Cell start, Cell goal, std::vector<Cell>* path) {
    auto inside = [&](int x, int y) { return 0 <= x && x < nx && 0 <= y && y < ny; };
    if (!inside(start.x, start.y) || !inside(goal.x, goal.y)) return false;
    struct Item { double d; int x, y; };
    auto later = [](const Item& a, const Item& b) {
        if (a.d != b.d) return a.d > b.d;
        if (a.x != b.x) return a.x > b.x;
        return a.y > b.y;
    };
    std::priority_queue<Item, std::vector<Item>, decltype(later)> heap(later);
    std::vector<double> dist((size_t)nx * ny, 1e30);
    std::vector<int> parent((size_t)nx * ny, -1);
    dist[(size_t)start.x * ny + start.y] = 0.0;
    heap.push({0.0, start.x, start.y});
    const int dxs[8] = {-1, -1, -1, 0, 0, 1, 1, 1}, dys[8] = {-1, 0, 1, -1, 1, -1, 0, 1};
    bool reached = false;
    while (!heap.empty()) {
        Item u = heap.top();
        heap.pop();
        if (u.x == goal.x && u.y == goal.y) { reached = true; break; }
        if (u.d > dist[(size_t)u.x * ny + u.y]) continue;
        for (int k = 0; k < 8; k++) {
            int vx = u.x + dxs[k], vy = u.y + dys[k];
            if (!inside(vx, vy) || blocked[(size_t)vx * ny + vy]) continue;
            double nd = u.d + ((dxs[k] && dys[k]) ? 1.4142135623730951 : 1.0);
            if (nd < dist[(size_t)vx * ny + vy]) {
                dist[(size_t)vx * ny + vy] = nd;
                parent[(size_t)vx * ny + vy] = u.x * ny + u.y;
                heap.push({nd, vx, vy});
            }
        }
    }
    if (!reached && dist[(size_t)goal.x * ny + goal.y] >= 1e30) return false;
    path->clear();
    int cur = goal.x * ny + goal.y, s = start.x * ny + start.y;
    path->push_back({cur / ny, cur % ny});
    while (cur != s) {
        cur = parent[cur];
        path->push_back({cur / ny, cur % ny});
    }
    std::reverse(path->begin(), path->end());
    return true;
}

struct Route { std::vector<Cell> pts; bool found; };

// the map generate_trajectory_dstar builds (ENV:1493-1507), searched for a shortest path to the first goal; with
// multiple_end_points two more legs, goal to goal on the same map, are appended (ENV:1552-1587)
struct Px { long x, y; };
Route plan_dstar_grid(const FtlScenarioGenConfig& g, long sx_px, long sy_px, const std::vector<Px>& goals_px,
                      const std::vector<Rect>& statics) {
    const int sg = g.step_grid, nx = g.game_width / sg, ny = g.game_height / sg;
    const int margin = (int)std::floor(g.leader_margin * std::max(g.leader_width_f, g.leader_height_f) / sg);
    std::vector<uint8_t> blocked((size_t)nx * ny, 0);
    for (const Rect& r : statics) {
        const int cx = r.x + (r.w >> 1), cy = r.y + (r.h >> 1);
        const int pmx = cx / sg, pmy = cy / sg;   // coordinates are non-negative: floor == truncation
        const int hh = (int)std::floor((r.h / 2.0) / sg) + margin, hw = (int)std::floor((r.w / 2.0) / sg) + margin;
        for (int i = pmx - hw; i < pmx + hw; i++)
            for (int j = pmy - hh; j < pmy + hh; j++)
                if (0 <= i && i < nx && 0 <= j && j < ny) blocked[(size_t)i * ny + j] = 1;
    }
    auto cell_of = [&](long px, long py) { return Cell{(int)(px / (double)sg), (int)(py / (double)sg)}; };
    auto is_blocked = [&](Cell c) { return 0 <= c.x && c.x < nx && 0 <= c.y && c.y < ny && blocked[(size_t)c.x * ny + c.y]; };
    Cell start = cell_of(sx_px, sy_px), goal = cell_of(goals_px[0].x, goals_px[0].y);
    if (0 <= start.x && start.x < nx && 0 <= start.y && start.y < ny) blocked[(size_t)start.x * ny + start.y] = 0;
    Route out;
    std::vector<Cell> path;
    if (is_blocked(goal) || !shortest_path(blocked, nx, ny, start, goal, &path)) {
        // unreachable target: the reference reports found_target_point=False (dstar.py:183-188) and SkipBadSeeds
        // re-resets; a short stub route keeps the episode well defined
        out.pts = {{start.x * sg, start.y * sg}, {std::max(start.x - 1, 0) * sg, start.y * sg}};
        out.found = false;
        return out;
    }
    // D* lists every cell from the start up to (not including) the goal cell (dstar.py:176-195)
    size_t count = path.size() > 2 ? path.size() - 1 : path.size();
    for (size_t k = 0; k < count; k++) out.pts.push_back({path[k].x * sg, path[k].y * sg});
    out.found = true;
    for (size_t leg = 1; leg < goals_px.size(); leg++) {
        // the next leg starts in the previous goal cell; a leg that cannot be planned adds nothing and clears the flag
        const Cell nxt = cell_of(goals_px[leg].x, goals_px[leg].y);
        if (is_blocked(nxt) || is_blocked(goal) || !shortest_path(blocked, nx, ny, goal, nxt, &path)) {
            out.found = false;
        } else {
            for (size_t k = 0; k + 1 < path.size(); k++) out.pts.push_back({path[k].x * sg, path[k].y * sg});
        }
        goal = nxt;
    }
    return out;
}

// generate_trajectory_astar (ENV:1632-1712): 20-px grid, obstacles grown by 2*max(leader size)
Route plan_astar_grid(const FtlScenarioGenConfig& g, long sx_px, long sy_px, long gx_px, long gy_px,
                      const std::vector<Rect>& statics) {
    const int sg = 20, nx = (int)(g.game_width / (double)sg), ny = (int)(g.game_height / (double)sg);
    const int grow = (int)(std::max(g.leader_width_f, g.leader_height_f) * 2);
    std::vector<uint8_t> blocked((size_t)nx * ny, 0);
    for (const Rect& r : statics) {
        const int sx = std::max((int)((r.x - grow) / (double)sg), 0), ex = std::min((int)((r.x + r.w + grow) / (double)sg), nx - 1);
        const int sy = std::max((int)((r.y - grow) / (double)sg), 0), ey = std::min((int)((r.y + r.h + grow) / (double)sg), ny - 1);
        for (int i = sx; i < ex; i++)
            for (int j = sy; j < ey; j++) blocked[(size_t)i * ny + j] = 1;
    }
    Cell start = {(int)(sx_px / (double)sg), (int)(sy_px / (double)sg)}, end = {(int)(gx_px / (double)sg), (int)(gy_px / (double)sg)};
    Route out;
    out.found = false;   // the reference only sets found_target_point inside the D* branch (ENV:1543)
    std::vector<Cell> p;
    if (g.add_obstacles && !statics.empty()) {
        const Rect &w1 = statics[0], &w2 = statics[1];
        const int bridge_y = (int)(((w1.y + (w1.h >> 1)) + (w2.y + (w2.h >> 1))) / 2.0 / sg);
        for (int i = (int)(w1.x / (double)sg - grow / (double)sg); i < (int)((w1.x + w1.w) / (double)sg + grow / (double)sg); i++)
            if (0 <= i && i < nx && 0 <= bridge_y && bridge_y < ny) blocked[(size_t)i * ny + bridge_y] = 0;
        const double eps = g.leader_pos_epsilon;
        Cell first = {(int)((w1.x + w1.w + eps) / sg), bridge_y}, second = {(int)((w1.x - eps) / sg), bridge_y};
        if (!shortest_path(blocked, nx, ny, start, first, &p)) return out;
        for (const Cell& c : p) out.pts.push_back({c.x * sg, c.y * sg});
        out.pts.push_back({(int)(w1.x + w1.w + eps), sg * bridge_y});   // self.first_bridge_point, ENV:1680-1694
        if (shortest_path(blocked, nx, ny, second, end, &p))
            for (const Cell& c : p) out.pts.push_back({c.x * sg, c.y * sg});
        return out;
    }
    if (shortest_path(blocked, nx, ny, start, end, &p))
        for (const Cell& c : p) out.pts.push_back({c.x * sg, c.y * sg});
    return out;
}

// FtlScenarioPool is declared with const pointers (it is an input everywhere else); here the caller hands in
// writable host arrays
struct PoolOut {
    int static_cap, route_cap;
    int32_t *static_rects, *n_static, *route, *n_route;
    float *leader_pos, *follower_pos;
    double *leader_dir, *follower_dir;
    uint8_t* found_target_point;
    explicit PoolOut(const FtlScenarioPool& p)
        : static_cap(p.static_cap), route_cap(p.route_cap), static_rects(const_cast<int32_t*>(p.static_rects)),
          n_static(const_cast<int32_t*>(p.n_static)), route(const_cast<int32_t*>(p.route)),
          n_route(const_cast<int32_t*>(p.n_route)), leader_pos(const_cast<float*>(p.leader_pos)),
          follower_pos(const_cast<float*>(p.follower_pos)), leader_dir(const_cast<double*>(p.leader_dir)),
          follower_dir(const_cast<double*>(p.follower_dir)),
          found_target_point(const_cast<uint8_t*>(p.found_target_point)) {}
};

// one scenario; returns 0 or a negative error code with *why filled in
int generate_one(const FtlScenarioGenConfig& g, int64_t seed, const PoolOut& out, int slot, std::string* why) {
    PyRandom rnd;
    rnd.seed(seed);
    const int W = g.game_width, H = g.game_height;
    const double max_d = g.max_distance, min_d = g.min_distance;
    auto draw = [&](double a, double b, long step, long* v) {   // _randrange: integral floats are accepted (py3.7)
        if (a != std::floor(a) || b != std::floor(b)) { *why = "non-integer arg for randrange()"; return false; }
        if (!rnd.randrange((long)a, (long)b, step, v)) { *why = "empty range for randrange()"; return false; }
        return true;
    };
    // ---- _create_robots, ENV:545-596
    long lx, ly, dist0;
    if (!draw(W / 2.0 + max_d, W - max_d, 10, &lx) || !draw(max_d, H - max_d, 10, &ly)) return FTL_ERR_INVALID;
    const double leader_start_direction = angle_to_point((double)lx, (double)ly, (double)(int)(W / 2.0), (double)(int)(H / 2.0));
    const Rect leader_rect = rect_centered((double)lx, (double)ly, g.leader_width, g.leader_height);
    if (!draw((double)(int)(min_d * 1.1), (double)(int)(max_d * 0.9), 1, &dist0)) return FTL_ERR_INVALID;
    const double theta0 = py_radians(angle_correction(leader_start_direction + 180));
    const float f0x = (float)((double)dist0 * std::cos(theta0) + (double)lx);   // GameObject keeps float32 positions, CLS:47
    const float f0y = (float)((double)dist0 * std::sin(theta0) + (double)ly);
    const Rect follower_rect0 = rect_centered((double)f0x, (double)f0y, g.follower_width, g.follower_height);
    std::vector<Rect> objects = {leader_rect, follower_rect0}, statics;
    // ---- _create_obstacles, ENV:613-677
    if (g.add_obstacles) {
        const int b0 = g.bridge_size[0], b1 = g.bridge_size[1];
        const int bh = (H - b0) / 2;   // floor division of non-negative ints
        const Rect wall1 = rect_centered(W / 2.0, (double)(bh / 2), b1, bh);
        const Rect wall2 = rect_centered(W / 2.0, (double)((H / 2) + (bh / 2) + (b0 / 2)), b1, bh);
        const int wall_start_x = wall1.x, wall_end_x = wall1.x + wall1.w, obstacle_size = 50;
        const Rect bridge_rect = {(int)(wall_start_x - g.leader_width_f * 4),
                                  (int)(wall1.y + wall1.h - g.leader_height_f * g.leader_margin),
                                  (int)(wall1.w + 8 * g.leader_width_f),
                                  (int)(wall2.y - (wall1.y + wall1.h) + 3 * g.leader_height_f)};
        statics.push_back(wall1);
        statics.push_back(wall2);
        for (int k = 0; k < g.obstacle_number; k++) {
            long px, py;
            for (;;) {
                if (!draw(130, W - 120, g.step_grid, &px) || !draw(20, H - 20, g.step_grid, &py)) return FTL_ERR_INVALID;
                if (collidepoint(leader_rect, px, py) || collidepoint(follower_rect0, px, py) ||
                    (wall_start_x <= px && px <= wall_end_x) || collidepoint(bridge_rect, px, py) ||
                    std::hypot((double)(lx - px), (double)(ly - py)) <= max_d + obstacle_size / 2.0)
                    continue;
                break;
            }
            statics.push_back(rect_centered((double)px, (double)py, obstacle_size, obstacle_size));
        }
        objects.insert(objects.end(), statics.begin(), statics.end());
    }
    if ((int)statics.size() > out.static_cap) { *why = "more static rectangles than static_cap"; return FTL_ERR_INVALID; }
    // ---- generate_finish_point, ENV:471-482, 1614-1630
    auto finish_point = [&](double l0, double l1, double r0, double r1, Px* fp) {
        for (;;) {
            if (!draw(l0, r0, 10, &fp->x) || !draw(l1, r1, 10, &fp->y)) return false;
            bool ok = true;
            for (const Rect& r : objects)
                if (collidepoint(r, fp->x, fp->y) || distance_to_rect(fp->x, fp->y, r) < g.leader_pos_epsilon) ok = false;
            if (ok) return true;
        }
    };
    std::vector<Px> goals(1);
    if (!finish_point(20, 20, (double)(int)(W / 2.0), H - 20, &goals[0])) return FTL_ERR_INVALID;
    if (g.multiple_end_points) {
        if (g.path_finding != 0) { *why = "Only dstar pathfinding function supports multiple end points (ENV:239-243)"; return FTL_ERR_INVALID; }
        for (int k = 0; k < 2; k++) {   // each in the half of the field (upper / lower) the previous one is not in
            Px fp;
            const bool lower = goals.back().y >= H / 2.0;
            if (!(lower ? finish_point(20, 20, W - 20, (double)(int)(H / 2.0), &fp)
                        : finish_point(20, (double)(int)(H / 2.0), W - 20, H - 20, &fp))) return FTL_ERR_INVALID;
            goals.push_back(fp);
        }
    }
    const long fx = goals[0].x, fy = goals[0].y;
    // ---- route
    Route route;
    if (g.path_finding == 0) {
        if (!g.add_obstacles) { *why = "'Game' object has no attribute 'obstacles1' (ENV:1501: dstar needs add_obstacles)"; return FTL_ERR_INVALID; }
        route = plan_dstar_grid(g, lx, ly, goals, statics);
    } else {
        route = plan_astar_grid(g, lx, ly, fx, fy, statics);
    }
    if (route.pts.size() < 2) { *why = "route planning produced fewer than two waypoints"; return FTL_ERR_STATE; }
    if ((int)route.pts.size() > out.route_cap) { *why = "route longer than route_cap"; return FTL_ERR_INVALID; }
    // ---- leader heading and follower placement, ENV:525-526, 598-611
    const double leader_dir = angle_to_point((double)lx, (double)ly, (double)route.pts[1].x, (double)route.pts[1].y);
    long dist;
    if (!draw((double)(int)(min_d * 1.1), (double)(int)(max_d * 0.9), 1, &dist)) return FTL_ERR_INVALID;
    const double theta = angle_correction(leader_dir + 180);
    const double pfx = (double)dist * std::cos(py_radians(theta)) + (double)lx;
    const double pfy = (double)dist * std::sin(py_radians(theta)) + (double)ly;
    const double follower_dir = angle_to_point(pfx, pfy, (double)lx, (double)ly);
    // ---- store
    int32_t* sr = out.static_rects + (size_t)slot * out.static_cap * 4;
    memset(sr, 0, sizeof(int32_t) * 4 * (size_t)out.static_cap);
    for (size_t k = 0; k < statics.size(); k++) {
        sr[4 * k] = statics[k].x; sr[4 * k + 1] = statics[k].y; sr[4 * k + 2] = statics[k].w; sr[4 * k + 3] = statics[k].h;
    }
    out.n_static[slot] = (int32_t)statics.size();
    int32_t* rt = out.route + (size_t)slot * out.route_cap * 2;
    memset(rt, 0, sizeof(int32_t) * 2 * (size_t)out.route_cap);
    for (size_t k = 0; k < route.pts.size(); k++) { rt[2 * k] = route.pts[k].x; rt[2 * k + 1] = route.pts[k].y; }
    out.n_route[slot] = (int32_t)route.pts.size();
    out.leader_pos[2 * slot] = (float)lx; out.leader_pos[2 * slot + 1] = (float)ly;
    out.leader_dir[slot] = leader_dir;
    out.follower_pos[2 * slot] = (float)pfx; out.follower_pos[2 * slot + 1] = (float)pfy;
    out.follower_dir[slot] = follower_dir;
    if (out.found_target_point) out.found_target_point[slot] = route.found ? 1 : 0;
    return FTL_OK;
}

}  // namespace

void ftl_set_error_message(const char* msg);   // ftl_capi.cu / hostsim.cpp: what ftl_last_error() returns

extern "C" int ftl_generate_scenarios(const FtlScenarioGenConfig* cfg, const int64_t* seeds, int32_t n,
                                      const FtlScenarioPool* out, int32_t n_threads) {
    if (!cfg || !seeds || !out || n < 0 || out->n_scenarios < n) { ftl_set_error_message("ftl_generate_scenarios: bad arguments"); return FTL_ERR_INVALID; }
    if (!out->static_rects || !out->n_static || !out->route || !out->n_route || !out->leader_pos || !out->leader_dir ||
        !out->follower_pos || !out->follower_dir) { ftl_set_error_message("ftl_generate_scenarios: NULL pool array"); return FTL_ERR_INVALID; }
    if (cfg->step_grid < 1 || cfg->game_width < 1 || cfg->game_height < 1) { ftl_set_error_message("ftl_generate_scenarios: bad geometry"); return FTL_ERR_INVALID; }
    int threads = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    threads = std::max(1, std::min(threads, n > 0 ? n : 1));
    std::vector<int> rc((size_t)threads, FTL_OK);
    std::vector<std::string> why((size_t)threads);
    std::vector<int64_t> bad_seed((size_t)threads, 0);
    const PoolOut po(*out);
    auto work = [&](int t) {
        for (int i = t; i < n; i += threads) {
            std::string w;
            int r = generate_one(*cfg, seeds[i], po, i, &w);
            if (r != FTL_OK && rc[t] == FTL_OK) { rc[t] = r; why[t] = w; bad_seed[t] = seeds[i]; }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < threads; t++)
        if (rc[t] != FTL_OK) {
            ftl_set_error_message(("ftl_generate_scenarios: seed " + std::to_string((long long)bad_seed[t]) + ": " + why[t]).c_str());
            return rc[t];
        }
    return FTL_OK;
}
