"""An independent numpy restatement of the rgb_array rasteriser's specification (csrc/ftl_capi.cu: k_render, which follows
Game._show_tick, ENV:1229-1302) -- test infrastructure: the GPU test compares ftl_render with it pixel by pixel.  The
reference draws sprites and text through pygame; neither is available here, so this is a check of the kernel against its
own written specification (layers, colours, geometry), not against pygame's pixels."""
import numpy as np

F = np.float32


def _seg_d2(px, py, ax, ay, bx, by):
    ax, ay, bx, by = F(ax), F(ay), F(bx), F(by)
    vx, vy = F(bx - ax), F(by - ay)
    wx, wy = (px - ax).astype(F), (py - ay).astype(F)
    vv = F(F(vx * vx) + F(vy * vy))
    if vv > 0:
        t = ((wx * vx).astype(F) + (wy * vy).astype(F)).astype(F) / vv
    else:
        t = np.zeros_like(px)
    t = np.clip(t.astype(F), F(0), F(1))
    dx, dy = (wx - (t * vx).astype(F)).astype(F), (wy - (t * vy).astype(F)).astype(F)
    return ((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)


def _in_rect(px, py, q):
    x, y, w, h = [int(v) for v in q]
    return (px >= F(x)) & (px < F(x + w)) & (py >= F(y)) & (py < F(y + h))


def green_lo(trail, n, max_distance):
    """_trajectory_in_box (ENV:1828-1843): float32 running sum of the segment lengths from the newest point backwards."""
    lo, acc = n - 1, F(0)
    for i in range(n - 2, -1, -1):
        d = trail[i + 1] - trail[i]
        acc = F(acc + F(np.sqrt(F(F(d[0] * d[0]) + F(d[1] * d[1])))))
        if acc <= F(max_distance):
            lo = i
        else:
            break
    return lo


def render_env(cfg, pool, env, trail, hist, corridor, scale):
    """cfg: FtlConfig; pool: ScenarioPool; env: one FtlEnvState record; the env's trail / hist / corridor rows."""
    W, H = -(-cfg.game_width // scale), -(-cfg.game_height // scale)
    xs = ((np.arange(W, dtype=F) + F(0.5)) * F(scale)).astype(F)
    ys = ((np.arange(H, dtype=F) + F(0.5)) * F(scale)).astype(F)
    px, py = np.meshgrid(xs, ys)
    thin = F(0.5 * scale)
    img = np.full((H, W, 3), 255, np.uint8)

    def paint(mask, col):
        img[mask] = col

    scen = int(env["scenario_id"])
    n_route = int(pool.n_route[scen])
    route = pool.route[scen]
    if n_route > 2:
        m = np.zeros((H, W), bool)
        for k in range(n_route - 1):
            m |= _seg_d2(px, py, route[k][0], route[k][1], route[k + 1][0], route[k + 1][1]) <= F(thin * thin)
        paint(m, (255, 0, 0))
    if n_route > 0:
        f = route[n_route - 1]
        r = max(F(5), thin)
        dx, dy = (px - F(f[0])).astype(F), (py - F(f[1])).astype(F)
        paint(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F) <= F(r * r), (255, 0, 0))
    n = int(env["trail_len"])
    lo = green_lo(trail, n, cfg.max_distance)
    hi = n - 2
    green_exact = np.zeros((H, W), bool)
    green_wide = np.zeros((H, W), bool)       # with up to 3 more points at the far end: what the kernel may also draw
    dev = F(cfg.max_dev)
    if hi - lo + 1 > 5:
        for k in range(max(lo - 3, 0), hi + 1):
            dx, dy = (px - trail[k][0]).astype(F), (py - trail[k][1]).astype(F)
            m = ((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F) <= F(dev * dev)
            green_wide |= m
            if k >= lo:
                green_exact |= m
    paint(green_exact, (0, 255, 0))
    lp = np.array(env["leader"]["pos"], F)
    dx, dy = (px - lp[0]).astype(F), (py - lp[1]).astype(F)
    d = np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)
    half = F((2.0 if env["too_close"] else 1.0) * thin)
    paint(np.abs((d - F(cfg.min_distance)).astype(F)) <= half, (255, 0, 0))
    m = np.zeros((H, W), bool)
    for k in range(int(pool.n_static[scen])):
        m |= _in_rect(px, py, pool.static_rects[scen][k])
    paint(m, (30, 30, 30))
    paint(_in_rect(px, py, env["leader"]["rect"]), (0, 0, 255))
    paint(_in_rect(px, py, env["follower"]["rect"]), (255, 140, 0))
    for b in range(cfg.n_bears):
        paint(_in_rect(px, py, env["bear"][b]["rect"]), (139, 69, 19))
    if cfg.tracker_enabled:
        tail, head, mask = int(env["ring_tail"]), int(env["ring_head"]), cfg.corridor_cap - 1
        rp, hw = max(F(3), thin), max(F(1.5), thin)
        pm = np.zeros((H, W), bool)
        for k in range(tail, head):
            h = hist[k & mask]
            dx, dy = (px - F(h[0])).astype(F), (py - F(h[1])).astype(F)
            pm |= ((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F) <= F(rp * rp)
        wm = np.zeros((H, W), bool)
        if head - tail > 1:
            for k in range(tail, head - 1):
                a, b = corridor[k & mask], corridor[(k + 1) & mask]
                wm |= _seg_d2(px, py, a[0], a[1], b[0], b[1]) <= F(hw * hw)
                wm |= _seg_d2(px, py, a[2], a[3], b[2], b[3]) <= F(hw * hw)
            a, b = corridor[tail & mask], corridor[(head - 1) & mask]
            wm |= _seg_d2(px, py, a[0], a[1], a[2], a[3]) <= F(hw * hw)
            wm |= _seg_d2(px, py, b[0], b[1], b[2], b[3]) <= F(hw * hw)
        paint(pm, (80, 10, 10))
        paint(wm, (150, 120, 50))
    tid = min(int(env["cur_target_id"]), n_route - 1)
    if tid >= 0:
        t = route[tid]
        dx, dy = (px - F(t[0])).astype(F), (py - F(t[1])).astype(F)
        d = np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)
        paint(np.abs((d - F(10)).astype(F)) <= max(F(1), thin), (255, 0, 0))
    return img, green_wide & ~green_exact
