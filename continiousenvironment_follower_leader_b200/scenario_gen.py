"""Host-side scenario generation: what ``Game.reset()`` does before the first sensor scan.

Restates ``_create_robots`` (ENV:545-596), ``_create_obstacles`` (ENV:613-677), ``generate_finish_point``
(ENV:1614-1630, with ``distance_to_rect`` of utils/misc.py:29-44) and ``_pos_follower_behind_leader``
(ENV:598-611) with the same calls to python's global ``random`` in the same order, so that after
``env.seed(s)`` the layout (leader start, rocks, finish point, follower distance) is the one the reference
draws for that seed.  pygame.Rect arithmetic is restated inline (integer x, y, w, h; truncating centre).

The leader route is planned with a plain 8-connected shortest-path search on the reference's grid and
obstacle inflation (ENV:1493-1507 for the D* map, ENV:1632-1712 for the A* variant).  The reference's D*
breaks ties by iterating a python ``set`` of objects (utils/dstar.py:88,130), which depends on memory
addresses; its routes are therefore reproduced as *valid shortest routes on the same grid*, not waypoint
for waypoint.  An explicit ``trajectory=`` is used verbatim, as in the reference.
"""
import heapq
import math
import random

import numpy as np

from .scenario import angle_to_point, place_follower, _angle_correction


def _rect(cx, cy, w, h):
    """image.get_rect(center=(cx, cy), width=w, height=h) after transform.scale(image, (w, h)), CLS:42-50."""
    w, h = int(w), int(h)
    return (int(cx) - (w >> 1), int(cy) - (h >> 1), w, h)


def _collidepoint(r, p):
    px, py = int(p[0]), int(p[1])
    return r[0] <= px < r[0] + r[2] and r[1] <= py < r[1] + r[3]


def _distance_to_rect(p, r):  # utils/misc.py:29-44
    x, y, w, h = r
    pts = [(x, y), (x, y + h), (x + w, y), (x + w, y + h), (x + (w >> 1), y), (x, y + (h >> 1)),
           (x + (w >> 1), y + h), (x + w, y + (h >> 1))]
    return min(math.hypot(p[0] - q[0], p[1] - q[1]) for q in pts)


def _randrange(a, b, step=1):
    """random.randrange with the python 3.7 tolerance for integral floats (the reference passes floats at ENV:549)."""
    def as_int(v):
        iv = int(v)
        if iv != v:
            raise ValueError("non-integer arg for randrange()")
        return iv
    return random.randrange(as_int(a), as_int(b), as_int(step))


class Scenario:
    """One reset's worth of data (see include/ftl.h, FtlScenarioPool)."""

    def __init__(self):
        self.static_rects = []
        self.route = []
        self.leader_pos = None
        self.leader_dir = 0.0
        self.follower_pos = None
        self.follower_dir = 0.0
        self.finish_point = None
        self.finish_points = None      # multiple_end_points: the three finish points, ENV:471-482
        self.found_target_point = False


def generate(gc, trajectory=None):
    """Draw one scenario with the global ``random`` state, like Game.reset (ENV:434-543)."""
    g, c = gc.kwargs, gc.c
    W, H = c.game_width, c.game_height
    max_distance, min_distance = c.max_distance, c.min_distance
    sc = Scenario()
    # ---- _create_robots, ENV:545-596 ----------------------------------------------------------------
    lx = _randrange(W / 2 + max_distance, W - max_distance, 10)
    ly = _randrange(max_distance, H - max_distance, 10)
    leader_start_direction = angle_to_point((lx, ly), (int(W / 2), int(H / 2)))
    leader_w, leader_h = c.leader.width, c.leader.height          # integer sprite
    leader_wf = g["leader_size"][0] * g["pixels_to_meter"]        # the float attributes the reference keeps
    leader_hf = g["leader_size"][1] * g["pixels_to_meter"]
    leader_rect = _rect(lx, ly, leader_w, leader_h)
    dist0 = _randrange(int(min_distance * 1.1), int(max_distance * 0.9), 1)
    theta0 = math.radians(_angle_correction(leader_start_direction + 180))
    f0 = (dist0 * math.cos(theta0) + lx, dist0 * math.sin(theta0) + ly)
    f0 = (float(np.float32(f0[0])), float(np.float32(f0[1])))     # GameObject keeps float32 positions, CLS:47
    follower_rect0 = _rect(f0[0], f0[1], c.follower.width, c.follower.height)
    objects = [leader_rect, follower_rect0]
    # ---- _create_obstacles, ENV:613-677 --------------------------------------------------------------
    if g["add_obstacles"]:
        bridge = g["bridge_size"]
        bh = (H - bridge[0]) // 2
        p1 = (W / 2, bh // 2)
        p2 = (W / 2, (H // 2) + (bh // 2) + (bridge[0] // 2))
        wall1, wall2 = _rect(p1[0], p1[1], bridge[1], bh), _rect(p2[0], p2[1], bridge[1], bh)
        wall_start_x, wall_end_x = wall1[0], wall1[0] + wall1[2]
        obstacle_size = 50
        # pygame.Rect(...) of floats truncates each argument
        bridge_rect = (int(wall_start_x - leader_wf * 4), int(wall1[1] + wall1[3] - leader_hf * g["leader_margin"]),
                       int(wall1[2] + 8 * leader_wf), int(wall2[1] - (wall1[1] + wall1[3]) + 3 * leader_hf))
        rocks = []
        for _ in range(int(g["obstacle_number"])):
            while True:
                p = (_randrange(130, W - 120, g["step_grid"]), _randrange(20, H - 20, g["step_grid"]))
                if _collidepoint(leader_rect, p) or _collidepoint(follower_rect0, p) or \
                        (wall_start_x <= p[0] <= wall_end_x) or _collidepoint(bridge_rect, p) or \
                        math.hypot(lx - p[0], ly - p[1]) <= max_distance + obstacle_size / 2:
                    continue
                break
            rocks.append((p, _rect(p[0], p[1], obstacle_size, obstacle_size)))
        sc.static_rects = [wall1, wall2] + [r for _, r in rocks]
        objects += sc.static_rects
    # ---- route ------------------------------------------------------------------------------------------
    if trajectory is not None:
        sc.route = [tuple(int(v) for v in p) for p in trajectory]
        sc.found_target_point = False       # the reference only sets it inside the D* branch (ENV:1543)
        sc.finish_point = sc.route[-1]
    else:
        def finish_point(left_top, right_bottom):       # generate_finish_point, ENV:1614-1630
            while True:
                fp = (_randrange(left_top[0], right_bottom[0], 10), _randrange(left_top[1], right_bottom[1], 10))
                ok = True
                for r in objects:
                    if _collidepoint(r, fp) or _distance_to_rect(fp, r) < c.leader_pos_epsilon:
                        ok = False
                if ok:
                    return fp
        fp = finish_point([20, 20], [int(W / 2), H - 20])                                       # ENV:471
        sc.finish_point = fp
        goals = [fp]
        if g["multiple_end_points"]:
            # two more finish points, each in the half of the field (upper / lower) the previous one is not in, ENV:472-482
            for _ in range(2):
                if goals[-1][1] >= H / 2:
                    goals.append(finish_point([20, 20], [W - 20, int(H / 2)]))
                else:
                    goals.append(finish_point([20, int(H / 2)], [W - 20, H - 20]))
            sc.finish_points = list(goals)
        if g["path_finding_algorythm"] == "dstar":
            if not g["add_obstacles"]:
                # the reference dereferences self.obstacles1 here (ENV:1501)
                raise AttributeError("'Game' object has no attribute 'obstacles1'")
            sc.route, sc.found_target_point = _plan_dstar_grid(gc, (lx, ly), goals, sc.static_rects, leader_wf, leader_hf)
        else:
            sc.route = _plan_astar_grid(gc, (lx, ly), fp, sc.static_rects, leader_wf, leader_hf)
            sc.found_target_point = False
    if len(sc.route) < 2:
        raise RuntimeError("route planning produced fewer than two waypoints")
    # ---- leader heading and follower placement, ENV:525-526, 598-611 -----------------------------------------
    sc.leader_pos = np.array((lx, ly), np.float32)
    sc.leader_dir = angle_to_point((lx, ly), sc.route[1])
    dist = _randrange(int(min_distance * 1.1), int(max_distance * 0.9), 1)
    sc.follower_pos, sc.follower_dir = place_follower((lx, ly), sc.leader_dir, dist)
    return sc


def gen_config(gc):
    """GameConfig -> the FtlScenarioGenConfig of include/ftl.h (what ftl_generate_scenarios needs)."""
    from . import abi
    g, c = gc.kwargs, gc.c
    if g["trajectory"] is not None:
        raise ValueError("an explicit trajectory= needs no generated route: use generate()")
    s = abi.FtlScenarioGenConfig()
    s.game_width, s.game_height = c.game_width, c.game_height
    s.min_distance, s.max_distance, s.leader_pos_epsilon = c.min_distance, c.max_distance, c.leader_pos_epsilon
    s.leader_width, s.leader_height = c.leader.width, c.leader.height
    s.follower_width, s.follower_height = c.follower.width, c.follower.height
    s.leader_width_f = g["leader_size"][0] * g["pixels_to_meter"]
    s.leader_height_f = g["leader_size"][1] * g["pixels_to_meter"]
    s.add_obstacles, s.obstacle_number, s.step_grid = int(bool(g["add_obstacles"])), int(g["obstacle_number"]), int(g["step_grid"])
    s.bridge_size[0], s.bridge_size[1] = int(g["bridge_size"][0]), int(g["bridge_size"][1])
    s.leader_margin = float(g["leader_margin"])
    s.path_finding = 0 if g["path_finding_algorythm"] == "dstar" else 1
    s.multiple_end_points = int(bool(g["multiple_end_points"]))
    return s


def generate_pool_native(gc, seeds, n_threads=0, lib_path=None):
    """Scenario pool for ``seeds`` through the C++ generator (ftl_generate_scenarios in libftl.so, all host cores):
    the same layouts and routes as ``random.seed(s); generate(gc)`` per seed, without the Python loop."""
    import ctypes as C
    from . import capi
    from .scenario import ScenarioPool
    L = capi.load(lib_path)
    seeds = np.ascontiguousarray(seeds, dtype=np.int64)
    pool = ScenarioPool(len(seeds), gc.c.static_cap, gc.c.route_cap)
    st, cfg = pool.c_struct(), gen_config(gc)
    capi.check(L, L.ftl_generate_scenarios(C.byref(cfg), seeds.ctypes.data, len(seeds), C.byref(st), int(n_threads)),
               "ftl_generate_scenarios")
    return pool


def _shortest_path(blocked, nx, ny, start, goal, max_iter=None):
    """8-connected Dijkstra with euclidean step costs (the metric of dstar.State.cost / astar)."""
    if not (0 <= start[0] < nx and 0 <= start[1] < ny and 0 <= goal[0] < nx and 0 <= goal[1] < ny):
        return None
    dist = {start: 0.0}
    parent = {}
    heap = [(0.0, start)]
    nbrs = [(-1, -1), (-1, 0), (-1, 1), (0, -1), (0, 1), (1, -1), (1, 0), (1, 1)]
    it = 0
    while heap:
        d, u = heapq.heappop(heap)
        if u == goal:
            break
        if d > dist.get(u, 1e30):
            continue
        it += 1
        if max_iter is not None and it > max_iter * 50:
            return None
        for dx, dy in nbrs:
            v = (u[0] + dx, u[1] + dy)
            if not (0 <= v[0] < nx and 0 <= v[1] < ny) or v in blocked:
                continue
            nd = d + (1.4142135623730951 if dx and dy else 1.0)
            if nd < dist.get(v, 1e30):
                dist[v] = nd
                parent[v] = u
                heapq.heappush(heap, (nd, v))
    if goal not in dist:
        return None
    path = [goal]
    while path[-1] != start:
        path.append(parent[path[-1]])
    return path[::-1]


def _plan_dstar_grid(gc, start_px, goals_px, static_rects, leader_wf, leader_hf):
    """The map generate_trajectory_dstar builds (ENV:1493-1507), searched for a shortest path to goals_px[0]; with
    multiple_end_points two more legs on fresh copies of the map, goal to goal, appended (ENV:1552-1587)."""
    g, c = gc.kwargs, gc.c
    sg = g["step_grid"]
    nx, ny = c.game_width // sg, c.game_height // sg
    margin = int(g["leader_margin"] * max(leader_wf, leader_hf) // sg)
    blocked = set()
    # the reference iterates rocks first, then the two walls, using start_position and the float sizes
    for (x, y, w, h) in static_rects:
        cx, cy = x + (w >> 1), y + (h >> 1)
        pm = (int(cx // sg), int(cy // sg))
        hh = int((h / 2) // sg) + margin
        hw = int((w / 2) // sg) + margin
        for i in range(pm[0] - hw, pm[0] + hw):
            for j in range(pm[1] - hh, pm[1] + hh):
                if 0 <= i < nx and 0 <= j < ny:
                    blocked.add((i, j))
    start = (int(start_px[0] / sg), int(start_px[1] / sg))
    cells_of = lambda p: (int(p[0] / sg), int(p[1] / sg))  # noqa: E731
    goal = cells_of(goals_px[0])
    blocked.discard(start)
    path = None if goal in blocked else _shortest_path(blocked, nx, ny, start, goal)
    if path is None:
        # unreachable target: the reference walks parents until path_finding_iterations runs out and
        # reports found_target_point=False (dstar.py:183-188); SkipBadSeeds then re-resets.  Return a
        # short stub route so the episode is well defined.
        stub = [start, (max(start[0] - 1, 0), start[1])]
        return [(p[0] * sg, p[1] * sg) for p in stub], False
    # D* lists every cell from the start up to (not including) the goal cell (dstar.py:176-195)
    cells = path[:-1] if len(path) > 2 else path
    found = True
    for nxt in goals_px[1:]:
        # the next leg starts in the previous goal cell; a leg that cannot be planned adds nothing and clears the flag
        # (upstream its point list is whatever the aborted parent walk left behind, ENV:1611)
        nxt = cells_of(nxt)
        leg = None if (nxt in blocked or goal in blocked) else _shortest_path(blocked, nx, ny, goal, nxt)
        if leg is None:
            found = False
        else:
            cells = cells + leg[:-1]
        goal = nxt
    return [(p[0] * sg, p[1] * sg) for p in cells], found


def _plan_astar_grid(gc, start_px, goal_px, static_rects, leader_wf, leader_hf):
    """generate_trajectory_astar (ENV:1632-1712): 20-px grid, obstacles grown by 2*max(leader size)."""
    g, c = gc.kwargs, gc.c
    sg = 20
    nx, ny = int(c.game_width / sg), int(c.game_height / sg)
    grow = int(max(leader_wf, leader_hf) * 2)
    blocked = set()
    for (x, y, w, h) in static_rects:
        sx, ex = max(int((x - grow) / sg), 0), min(int((x + w + grow) / sg), nx - 1)
        sy, ey = max(int((y - grow) / sg), 0), min(int((y + h + grow) / sg), ny - 1)
        for i in range(sx, ex):
            for j in range(sy, ey):
                blocked.add((i, j))
    start = (int(start_px[0] / sg), int(start_px[1] / sg))
    end = (int(goal_px[0] / sg), int(goal_px[1] / sg))
    scale = lambda path: [(p[0] * sg, p[1] * sg) for p in path]  # noqa: E731  (astar.return_path rescales by 20)
    if g["add_obstacles"] and static_rects:
        wall1 = static_rects[0]
        wall2 = static_rects[1]
        bridge_y = int(((wall1[1] + (wall1[3] >> 1)) + (wall2[1] + (wall2[3] >> 1))) / 2 / sg)
        for i in range(int(wall1[0] / sg - grow / sg), int((wall1[0] + wall1[2]) / sg + grow / sg)):
            blocked.discard((i, bridge_y))
        eps = c.leader_pos_epsilon
        first = (int((wall1[0] + wall1[2] + eps) / sg), bridge_y)
        second = (int((wall1[0] - eps) / sg), bridge_y)
        p1 = _shortest_path(blocked, nx, ny, start, first)
        if p1 is None:
            return []
        path = scale(p1)
        path.append((int(wall1[0] + wall1[2] + eps), sg * bridge_y))   # self.first_bridge_point, ENV:1680-1694
        p2 = _shortest_path(blocked, nx, ny, second, end)
        return path + (scale(p2) if p2 else [])
    p = _shortest_path(blocked, nx, ny, start, end)
    return scale(p) if p else []
