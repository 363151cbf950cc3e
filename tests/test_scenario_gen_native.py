"""The C++ scenario generator (ftl_generate_scenarios, csrc/ftl_scenario_gen.cpp) against its python twin
(scenario_gen.generate, itself pinned to the reference's layouts in tests/test_gym_surface.py) and against the
layouts stored in the reference traces.  Host-only code: runs without a GPU, through libftl.so itself."""
import random
import time

import numpy as np
import pytest

import parity
from continiousenvironment_follower_leader_b200 import scenario_gen
from continiousenvironment_follower_leader_b200.config import GameConfig, cfg3_sensors, TEST_GAME_MANUAL_GAZEBO_KWARGS
from continiousenvironment_follower_leader_b200.scenario import ScenarioPool

CASES = {
    "default": dict(),
    "cfg3": dict(bear_number=1, follower_sensors=cfg3_sensors()),
    "astar_no_obstacles": dict(add_obstacles=False, add_bear=False, path_finding_algorythm="astar"),
    "astar_obstacles": dict(path_finding_algorythm="astar", obstacle_number=20),
    "gazebo": dict(TEST_GAME_MANUAL_GAZEBO_KWARGS, route_cap=256),
    "multiple_end_points": dict(multiple_end_points=True),
}


def _python_pool(gc, seeds):
    pool = ScenarioPool(len(seeds), gc.c.static_cap, gc.c.route_cap)
    for i, s in enumerate(seeds):
        random.seed(int(s))
        sc = scenario_gen.generate(gc)
        pool.set(i, sc.static_rects, sc.route, sc.leader_pos, sc.leader_dir, sc.follower_pos, sc.follower_dir,
                 sc.found_target_point)
    return pool


@pytest.mark.parametrize("name", sorted(CASES))
def test_native_generator_equals_the_python_one(name):
    gc = GameConfig(**CASES[name])
    seeds = [0, 1, 2, 3, 5, 7, 11, 23, 1234567, 2 ** 33 + 17] + list(range(100, 106))
    want = _python_pool(gc, seeds)
    got = scenario_gen.generate_pool_native(gc, seeds, n_threads=3)
    for k in ("n_static", "static_rects", "n_route", "route", "leader_pos", "follower_pos", "found_target_point"):
        assert np.array_equal(getattr(got, k), getattr(want, k)), k
    # headings come out of atan / cos / sin of the same libm: bit-equal
    assert np.array_equal(got.leader_dir, want.leader_dir) and np.array_equal(got.follower_dir, want.follower_dir)


@pytest.mark.parametrize("trace", ["cfg3_seed5_follow", "cfg3_seed11_random", "cfg1_auto_seed0_random", "flat_sensors_seed9"])
def test_native_generator_draws_the_reference_layout(trace):
    d, meta = parity.load_trace(parity.GOLDEN_DIR + "/" + trace + ".npz")
    gc = GameConfig(**meta["kwargs"])
    pool = scenario_gen.generate_pool_native(gc, [meta["seed"]])
    ns = int(pool.n_static[0])
    assert np.array_equal(pool.static_rects[0, :ns], d["scen_static_rects"])    # walls + rocks, same MT19937 draws
    assert np.array_equal(pool.leader_pos[0], d["scen_leader_pos"])
    assert tuple(pool.route[0, 0]) == tuple(d["scen_route"][0])


def _multi_end_cases():
    import json
    with open(parity.GOLDEN_DIR + "/multi_end_points.json") as f:
        return json.load(f)


@pytest.mark.parametrize("native", [False, True])
def test_multiple_end_points_against_the_reference(native):
    """Game(multiple_end_points=True).reset() of the unmodified reference (oracle/gen_multi_end_golden.py): three finish
    points from the same MT19937 draws, three D* legs appended (ENV:471-482, 1552-1611).  Waypoints differ by D*'s
    tie-breaks; the finish points, the number of waypoints (same grid, same metric), the first waypoint and the joints
    of the legs do not."""
    gold = _multi_end_cases()
    gc = GameConfig(**gold["kwargs"])
    assert gc.c.route_cap >= 512
    seeds = [c["seed"] for c in gold["cases"]]
    pool = scenario_gen.generate_pool_native(gc, seeds) if native else None
    for i, case in enumerate(gold["cases"]):
        if native:
            route = [tuple(int(v) for v in p) for p in pool.route[i, :pool.n_route[i]]]
            found = bool(pool.found_target_point[i])
        else:
            random.seed(case["seed"])
            sc = scenario_gen.generate(gc)
            assert [list(p) for p in sc.finish_points] == case["finish_points"]
            route, found = [tuple(p) for p in sc.route], sc.found_target_point
        assert found == case["found_target_point"]
        assert len(route) == case["n_route"]
        assert list(route[0]) == case["route_first"]
        sg = gc.kwargs["step_grid"]
        cell = lambda p: (int(p[0] / sg) * sg, int(p[1] / sg) * sg)
        ref_route = [tuple(p) for p in case["route"]]
        # a leg lists its cells from its start up to, not including, its goal: the first two finish cells are on the
        # route (as the first cell of the next leg), at the same index as upstream; the last one is one step past the end
        for fp in case["finish_points"][:2]:
            assert cell(fp) in route and route.index(cell(fp)) == ref_route.index(cell(fp))
        last = cell(case["finish_points"][2])
        assert max(abs(route[-1][0] - last[0]), abs(route[-1][1] - last[1])) == sg
        # consecutive waypoints are neighbouring cells, and the route (with its last step, which neither lists) has the
        # reference's length in pixels
        length = lambda r: sum(np.hypot(a[0] - b[0], a[1] - b[1]) for a, b in zip(r[:-1], r[1:]))
        assert all(max(abs(a[0] - b[0]), abs(a[1] - b[1])) == sg for a, b in zip(route[:-1], route[1:]))
        assert abs(length(route + [last]) - length(ref_route + [last])) < 1e-6 * length(ref_route)


def test_native_generator_rejects_what_the_reference_rejects():
    gc = GameConfig(add_obstacles=False, add_bear=False)       # dstar without the bridge objects: ENV:1501
    with pytest.raises(ValueError):
        scenario_gen.generate_pool_native(gc, [1])


def test_native_generator_is_fast_enough_for_large_pools():
    gc = GameConfig(bear_number=1, follower_sensors=cfg3_sensors(), route_cap=256)
    t0 = time.time()
    pool = scenario_gen.generate_pool_native(gc, np.arange(1000, 1256))
    dt = time.time() - t0
    assert pool.n_route.min() >= 2 and dt < 60.0
    print("256 scenarios in %.2f s" % dt)
