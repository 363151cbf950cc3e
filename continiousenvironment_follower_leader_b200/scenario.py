"""Scenario pools: what ``Game.reset()`` builds before the first sensor scan, as arrays.

The reference creates one scenario per ``reset()`` with python's global RNG and a D*/A* planner
(follow_the_leader_continuous_env.py:434-677, 1493-1712; one to two seconds each).  The kernels
consume scenarios as data (include/ftl.h, FtlScenarioPool); this module holds the container and a
synthetic generator that follows the reference's placement rules with numpy.
"""
import ctypes as C
import math

import numpy as np

from . import abi


class ScenarioPool:
    """Host arrays of S scenarios.  Field meanings: include/ftl.h, FtlScenarioPool."""

    def __init__(self, n, static_cap, route_cap):
        self.n, self.static_cap, self.route_cap = int(n), int(static_cap), int(route_cap)
        self.static_rects = np.zeros((n, static_cap, 4), np.int32)
        self.n_static = np.zeros(n, np.int32)
        self.route = np.zeros((n, route_cap, 2), np.int32)
        self.n_route = np.zeros(n, np.int32)
        self.leader_pos = np.zeros((n, 2), np.float32)
        self.leader_dir = np.zeros(n, np.float64)
        self.follower_pos = np.zeros((n, 2), np.float32)
        self.follower_dir = np.zeros(n, np.float64)
        self.found_target_point = np.ones(n, np.uint8)

    def c_struct(self):
        p = abi.FtlScenarioPool()
        p.n_scenarios, p.static_cap, p.route_cap = self.n, self.static_cap, self.route_cap
        for k in ("static_rects", "n_static", "route", "n_route", "leader_pos", "leader_dir", "follower_pos",
                  "follower_dir", "found_target_point"):
            a = getattr(self, k)
            assert a.flags["C_CONTIGUOUS"]
            setattr(p, k, a.ctypes.data)
        return p

    def set(self, i, static_rects, route, leader_pos, leader_dir, follower_pos, follower_dir, found=True):
        static_rects = np.asarray(static_rects, np.int32).reshape(-1, 4)
        route = np.asarray(route, np.int32).reshape(-1, 2)
        if len(static_rects) > self.static_cap:
            raise ValueError("scenario has %d static rects, static_cap is %d" % (len(static_rects), self.static_cap))
        if len(route) > self.route_cap:
            raise ValueError("route has %d waypoints, route_cap is %d" % (len(route), self.route_cap))
        if len(route) < 2:
            raise ValueError("a route needs at least two waypoints (ENV:506-514)")
        self.static_rects[i, :len(static_rects)] = static_rects
        self.n_static[i] = len(static_rects)
        self.route[i, :len(route)] = route
        self.n_route[i] = len(route)
        self.leader_pos[i] = leader_pos
        self.leader_dir[i] = leader_dir
        self.follower_pos[i] = follower_pos
        self.follower_dir[i] = follower_dir
        self.found_target_point[i] = bool(found)

    def save(self, path):
        np.savez_compressed(path, **{k: getattr(self, k) for k in (
            "static_rects", "n_static", "route", "n_route", "leader_pos", "leader_dir", "follower_pos",
            "follower_dir", "found_target_point")})

    @classmethod
    def from_arrays(cls, d):
        n, static_cap = d["static_rects"].shape[:2]
        pool = cls(n, static_cap, d["route"].shape[1])
        for k in ("static_rects", "n_static", "route", "n_route", "leader_pos", "leader_dir", "follower_pos",
                  "follower_dir", "found_target_point"):
            getattr(pool, k)[...] = d[k]
        return pool


def _angle_correction(a):
    if a >= 360:
        return a - 360
    if a < 0:
        return 360 + a
    return a


def angle_to_point(cur, target):
    """utils/misc.py:16-26 on float64 operands."""
    rx, ry = float(target[0]) - float(cur[0]), float(target[1]) - float(cur[1])
    if rx > 0:
        res = math.degrees(math.atan(ry / rx))
    elif rx < 0:
        res = math.degrees(math.atan(ry / rx)) + 180
    else:
        res = 0
    return _angle_correction(res)


def place_follower(leader_pos, leader_dir, distance):
    """_pos_follower_behind_leader, ENV:598-611: returns (float32 position, float64 direction)."""
    theta = _angle_correction(leader_dir + 180)
    fx = distance * math.cos(math.radians(theta)) + float(leader_pos[0])
    fy = distance * math.sin(math.radians(theta)) + float(leader_pos[1])
    fdir = angle_to_point((fx, fy), (float(leader_pos[0]), float(leader_pos[1])))
    return np.array((fx, fy), dtype=np.float32), fdir


def synthetic_pool(game_config, n_scenarios, seed=0, min_waypoints=40, max_waypoints=110):
    """Seeded synthetic scenarios following the reference's placement rules:

    * leader start on the 10-px grid, x in [W/2 + max_distance, W - max_distance), y in
      [max_distance, H - max_distance) (ENV:548-550);
    * two bridge walls (ENV:617-633) and ``obstacle_number`` 50x50 rocks on the grid obeying the
      rejection rules of ENV:654-662 when ``add_obstacles``;
    * an 8-connected grid polyline route that keeps clear of the rocks by the leader margin and goes
      through the bridge gap (a greedy walker, NOT D*: routes are valid, not D*-identical);
    * follower 1.1*min_distance .. 0.9*max_distance pixels behind the leader (ENV:599-600).
    """
    c = game_config.c
    g = game_config.kwargs
    rng = np.random.RandomState(seed)
    W, H, grid = c.game_width, c.game_height, int(g["step_grid"])
    add_obstacles = bool(g["add_obstacles"])
    n_rocks = int(g["obstacle_number"]) if add_obstacles else 0
    pool = ScenarioPool(n_scenarios, c.static_cap, c.route_cap)
    max_waypoints = min(max_waypoints, c.route_cap)
    bridge_h = (H - g["bridge_size"][0]) // 2
    lw, lh = c.leader.width, c.leader.height
    margin = int(g["leader_margin"] * max(lw, lh)) + 12
    for s in range(n_scenarios):
        for _attempt in range(1000):
            lx = int(rng.choice(np.arange(W // 2 + int(c.max_distance), W - int(c.max_distance), 10)))
            ly = int(rng.choice(np.arange(int(c.max_distance), H - int(c.max_distance), 10)))
            rects = []
            if add_obstacles:
                ww = int(g["bridge_size"][1])
                p1 = (W / 2, bridge_h // 2)
                p2 = (W / 2, (H // 2) + (bridge_h // 2) + (g["bridge_size"][0] // 2))
                for p in (p1, p2):
                    rects.append((int(p[0]) - (ww >> 1), int(p[1]) - (bridge_h >> 1), ww, bridge_h))
                wall_x0, wall_x1 = rects[0][0], rects[0][0] + ww
                gap_top, gap_bot = rects[0][1] + bridge_h, rects[1][1]
                while len(rects) < 2 + n_rocks:
                    x = int(rng.choice(np.arange(130, W - 120, grid)))
                    y = int(rng.choice(np.arange(20, H - 20, grid)))
                    if wall_x0 - 4 * lw <= x <= wall_x1 + 4 * lw and gap_top - 2 * lh <= y <= gap_bot + 2 * lh:
                        continue
                    if wall_x0 <= x <= wall_x1:
                        continue
                    if math.hypot(x - lx, y - ly) <= c.max_distance + 25:
                        continue
                    rects.append((x - 25, y - 25, 50, 50))
            route = _walk_route(rng, (lx, ly), rects, W, H, grid, margin, min_waypoints, max_waypoints,
                                (W // 2, H // 2) if add_obstacles else None)
            if route is not None:
                break
        else:
            raise RuntimeError("could not build a synthetic scenario")
        ldir = angle_to_point((lx, ly), route[1])
        dist = int(rng.randint(int(c.min_distance * 1.1), int(c.max_distance * 0.9)))
        fpos, fdir = place_follower((lx, ly), ldir, dist)
        pool.set(s, rects, route, (lx, ly), ldir, fpos, fdir)
    return pool


def _blocked(x, y, rects, margin, W, H):
    if x < 30 or y < 30 or x > W - 30 or y > H - 30:
        return True
    for (rx, ry, rw, rh) in rects:
        if rx - margin <= x <= rx + rw + margin and ry - margin <= y <= ry + rh + margin:
            return True
    return False


def _walk_route(rng, start, rects, W, H, grid, margin, min_wp, max_wp, bridge):
    """Greedy 8-connected walk on the grid towards random goals (left half of the map when there is a
    bridge, through the gap), avoiding inflated rectangles."""
    n_target = int(rng.randint(min_wp, max_wp + 1))
    goals = []
    if bridge is not None:
        goals.append((bridge[0] + 60, bridge[1]))
        goals.append((bridge[0] - 60, bridge[1]))
    for _ in range(8):
        gx = int(rng.randint(4, (W // 2 if bridge is not None else W) // grid - 4)) * grid
        gy = int(rng.randint(4, H // grid - 4)) * grid
        goals.append((gx, gy))
    route = [tuple(start)]
    x, y = start
    gi = 0
    stuck = 0
    prev = None
    while len(route) < n_target and gi < len(goals):
        gx, gy = goals[gi]
        if abs(gx - x) < grid and abs(gy - y) < grid:
            gi += 1
            continue
        sx = grid * int(np.sign(gx - x)) if abs(gx - x) >= grid else 0
        sy = grid * int(np.sign(gy - y)) if abs(gy - y) >= grid else 0
        cands = [(sx, sy), (sx, 0), (0, sy), (sx, -sy if sy else grid), (-sx if sx else grid, sy),
                 (0, grid), (0, -grid), (grid, 0), (-grid, 0)]
        moved = False
        # walls of the bridge may only be crossed inside the gap
        in_gap = bridge is not None and gi <= 1
        use_rects = rects[2:] if in_gap and abs(y - bridge[1]) <= 20 else rects
        for (dx, dy) in cands:
            if dx == 0 and dy == 0:
                continue
            nx, ny = x + dx, y + dy
            if (nx, ny) == prev:
                continue
            if _blocked(nx, ny, use_rects, margin, W, H):
                continue
            prev = (x, y)
            x, y = nx, ny
            route.append((x, y))
            moved = True
            break
        if not moved:
            stuck += 1
            gi += 1
            if stuck > 6:
                return None
    if len(route) < min_wp:
        return None
    return route
