// hostsim.cpp -- TEST TARGET: compiles the device functions of csrc/*.cuh for the CPU.
//
// There is no GPU in the build container, so the logic the CUDA kernels run (ftl_device.cuh,
// ftl_step.cuh, ftl_rays.cuh, ftl_state_io.cuh -- all host+device inline code) is compiled here with
// g++ and driven through the same host-buffer C-ABI entry points as libftl.so (ftl_create,
// ftl_upload_scenarios, ftl_reset_host, ftl_step_host, ftl_get_state, ftl_set_state).  `pytest -m "not
// gpu"` checks it against the oracle and the golden traces.  The product package never loads this
// library: libftl.so (CUDA) is the only product path.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../continiousenvironment_follower_leader_b200/csrc/ftl_rays.cuh"
#include "../../continiousenvironment_follower_leader_b200/csrc/ftl_state_io.cuh"
#include "../../continiousenvironment_follower_leader_b200/csrc/ftl_step.cuh"

using namespace ftl;

static std::string g_err;
void ftl_set_error_message(const char* msg) { g_err = msg; }   // for ftl_scenario_gen.cpp

struct FtlHandle_ {
    DevCfg cfg;
    int n = 0;
    DevState st{};
    DevPool pool{};
    std::vector<void*> allocs, pool_allocs;
    bool have_pool = false;
    int rays_total = 0;
    std::vector<double2> rot;
};

template <typename T>
static T* zalloc(std::vector<void*>& list, size_t count) {
    void* p = calloc(count ? count : 1, sizeof(T));
    list.push_back(p);
    return (T*)p;
}

static float sq_threshold(double limit) {
    float lim = (float)limit;
    if (!(lim >= 0.f)) return -1.f;
    float x = lim * lim;
    while (sqrtf(x) > lim) x = nextafterf(x, 0.f);
    for (;;) {
        float y = nextafterf(x, INFINITY);
        if (sqrtf(y) <= lim) x = y; else break;
    }
    return x;
}

template <int NB>
static void step_all(FtlHandle_* h, const void* actions, const DevOutputs& out) {
    for (int i = 0; i < h->n; i++) {
        World<NB> w;
        Episode e;
        world_load<NB>(h->st, i, w);
        episode_load(h->st, i, e);
        double a0, a1;
        decode_action(h->cfg.c, actions, i, h->n, &a0, &a1);
        env_step<NB>(h->cfg, h->st, h->pool, i, a0, a1, w, e);
        write_outputs<NB>(h->cfg, h->pool, out, i, w, e, false);
        if (h->cfg.c.auto_reset && (e.flags & FL_DONE)) {
            int scen = next_scenario(h->cfg, h->pool.n_scenarios, i, e.episode);
            env_reset<NB>(h->cfg, h->st, h->pool, i, scen, w, e);
            write_outputs<NB>(h->cfg, h->pool, out, i, w, e, true);
        }
        world_store<NB>(h->st, i, w);
        episode_store(h->st, i, e);
    }
}
template <int NB>
static void reset_all(FtlHandle_* h, const uint8_t* mask, const int* ids, const DevOutputs& out) {
    for (int i = 0; i < h->n; i++) {
        if (mask && !mask[i]) continue;
        World<NB> w;
        Episode e;
        int episodes = h->st.gi[(size_t)GI_EPISODE * h->n + i];
        int scen = ids ? ids[i] : next_scenario(h->cfg, h->pool.n_scenarios, i, episodes);
        env_reset<NB>(h->cfg, h->st, h->pool, i, scen, w, e);
        write_outputs<NB>(h->cfg, h->pool, out, i, w, e, false);
        world_store<NB>(h->st, i, w);
        episode_store(h->st, i, e);
    }
}
static void optional_all(FtlHandle_* h, const DevOutputs& out) {
    if (!out.follower_info && !out.track_vectors && !out.radar && !out.laser) return;
    for (int i = 0; i < h->n; i++) write_optional_sensors(h->cfg.c, h->st, h->pool, out, i);
}
static void rays_all(FtlHandle_* h, float* rays) {
    if (!rays || !h->rays_total) return;
    std::vector<unsigned char> buf(ray_shared_bytes(h->rays_total, h->cfg.ray_hmax) + 16);
    RayShared& sh = *reinterpret_cast<RayShared*>(buf.data());
    for (int i = 0; i < h->n; i++) rays_warp(h->cfg, h->st, h->pool, h->rot.data(), i, sh, rays);
    for (int i = 0; i < h->n; i++) rays_exact_env(h->cfg, h->st, h->pool, i, rays);
}

static DevOutputs dev_out(const FtlOutputs* o) {
    DevOutputs d{};
    d.n = 1 << 30;
    if (o) {
        d.numerical_features = o->numerical_features; d.leader_target = o->leader_target; d.rays = o->rays;
        d.reward = o->reward; d.done = o->done; d.status = o->status;
        d.follower_info = o->follower_info; d.track_vectors = o->track_vectors; d.radar = o->radar; d.laser = o->laser;
    }
    return d;
}

extern "C" {
const char* ftl_last_error(void) { return g_err.c_str(); }
int ftl_abi_version(void) { return FTL_ABI_VERSION; }

int ftl_create(const FtlConfig* cfg, int32_t n_envs, int32_t device, int64_t env_id_base, ftl_handle* out) {
    const FtlConfig& c = *cfg;
    if (c.static_cap > 64 || (c.corridor_cap & (c.corridor_cap - 1))) { g_err = "bad caps"; return FTL_ERR_INVALID; }
    FtlHandle_* h = new FtlHandle_();
    h->n = n_envs;
    DevCfg& d = h->cfg;
    memset(&d, 0, sizeof d);
    d.c = c;
    d.env_id_base = env_id_base;
    for (int s = 0; s < c.n_ray_sensors; s++) d.rays_per_env += sensor_width(c.ray[s]);
    ray_out_layout(d);
    d.ray_hmax = ray_hmax(c);
    ray_static_tables(d);
    h->rays_total = total_rays(c);
    d.rays_total = h->rays_total;
    for (int s = 0; s < c.n_ray_sensors; s++)
        for (int k = 0; k < c.ray[s].lasers_count; k++) {
            double th = ray_angle(c.ray[s], k) * kDeg2Rad;
            h->rot.push_back(make_double2(std::cos(th), std::sin(th)));
        }
    d.eps_f32 = (float)c.leader_pos_epsilon;
    d.dev_f32 = (float)c.max_dev;
    d.eps2_f32 = sq_threshold(c.leader_pos_epsilon);
    d.dev2_f32 = sq_threshold(c.max_dev);
    d.min_dist2_f32 = sq_threshold(c.min_distance);
    d.max_distance_f32 = (float)c.max_distance;
    d.es_far_f32 = (float)(c.max_distance * c.es_max_distance_coef);
    d.trail_seed_denom_f32 = (float)(c.trajectory_saving_period * c.leader.max_speed);
    d.corridor_length_f32 = (float)c.corridor_length;
    d.corridor_width_f32 = (float)c.corridor_width;
    auto inflate = [&](const FtlRobotConfig& r) {
        double half_diag = 0.5 * std::sqrt((double)r.width * r.width + (double)r.height * r.height);
        return (float)(half_diag + c.frames_per_step * std::fabs(r.max_speed) + 4.0);
    };
    d.static_inflate[0] = inflate(c.follower);
    d.static_inflate[1] = inflate(c.leader);
    size_t n = n_envs;
    int nb = c.n_bears, nr = 2 + nb;
    DevState& s = h->st;
    s.n = n_envs; s.n_real = n_envs; s.n_bears = nb;
    s.gd = zalloc<double>(h->allocs, GD_COUNT * n);
    s.rd = zalloc<double>(h->allocs, (size_t)nr * RD_COUNT * n);
    s.bear_tgt = zalloc<double>(h->allocs, (size_t)nb * 2 * n);
    s.gi = zalloc<int>(h->allocs, GI_COUNT * n);
    s.ri = zalloc<int>(h->allocs, (size_t)nr * n);
    s.bear_idx = zalloc<int>(h->allocs, (size_t)nb * n);
    s.gf = zalloc<float>(h->allocs, GF_COUNT * n);
    s.pos = zalloc<float2>(h->allocs, (size_t)nr * n);
    s.rect = zalloc<int4>(h->allocs, (size_t)nr * n);
    s.trail = zalloc<float2>(h->allocs, n * c.trail_cap);
    s.trail_d = zalloc<float>(h->allocs, n * c.trail_cap);
    s.trail_s = zalloc<double>(h->allocs, n * c.trail_cap);
    s.hist = zalloc<double2>(h->allocs, n * c.corridor_cap);
    s.corridor = zalloc<float4>(h->allocs, n * c.corridor_cap);
    s.seg_d = zalloc<double>(h->allocs, n * c.corridor_cap);
    s.seg_f = zalloc<float>(h->allocs, n * c.corridor_cap);
    s.snap_range = zalloc<int2>(h->allocs, (size_t)FTL_MAX_HIST * n);
    s.snap_rect = zalloc<int4>(h->allocs, (size_t)FTL_MAX_HIST * (1 + nb) * n);
    s.unc_rec = zalloc<UncRec>(h->allocs, n * kUncPerEnv);
    s.unc_count = zalloc<int>(h->allocs, n);
    *out = h;
    return FTL_OK;
}
int ftl_destroy(ftl_handle h) {
    if (!h) return 0;
    for (void* p : h->allocs) free(p);
    for (void* p : h->pool_allocs) free(p);
    delete h;
    return 0;
}
int ftl_rays_per_env(ftl_handle h) { return h->cfg.rays_per_env; }
int ftl_num_envs(ftl_handle h) { return h->n; }

int ftl_upload_scenarios(ftl_handle h, const FtlScenarioPool* p) {
    const FtlConfig& c = h->cfg.c;
    if (p->static_cap != c.static_cap || p->route_cap != c.route_cap) { g_err = "caps differ"; return FTL_ERR_INVALID; }
    for (void* q : h->pool_allocs) free(q);
    h->pool_allocs.clear();
    size_t S = p->n_scenarios;
    auto dup = [&](const void* src, size_t bytes) { void* q = malloc(bytes ? bytes : 1); memcpy(q, src, bytes); h->pool_allocs.push_back(q); return q; };
    DevPool& d = h->pool;
    d.n_scenarios = p->n_scenarios;
    d.static_rects = (const int4*)dup(p->static_rects, S * c.static_cap * 16);
    d.n_static = (const int*)dup(p->n_static, S * 4);
    d.route = (const int2*)dup(p->route, S * c.route_cap * 8);
    d.n_route = (const int*)dup(p->n_route, S * 4);
    d.leader_pos = (const float2*)dup(p->leader_pos, S * 8);
    d.leader_dir = (const double*)dup(p->leader_dir, S * 8);
    d.follower_pos = (const float2*)dup(p->follower_pos, S * 8);
    d.follower_dir = (const double*)dup(p->follower_dir, S * 8);
    h->have_pool = true;
    return FTL_OK;
}

int ftl_reset_host(ftl_handle h, const uint8_t* mask, const int32_t* ids, const FtlOutputs* out, void*) {
    if (!h->have_pool) return FTL_ERR_STATE;
    DevOutputs o = dev_out(out);
    switch (h->cfg.c.n_bears) {
        case 0: reset_all<0>(h, mask, ids, o); break;
        case 1: reset_all<1>(h, mask, ids, o); break;
        case 2: reset_all<2>(h, mask, ids, o); break;
        case 3: reset_all<3>(h, mask, ids, o); break;
        default: reset_all<4>(h, mask, ids, o); break;
    }
    optional_all(h, o);
    rays_all(h, o.rays);
    return FTL_OK;
}
int ftl_step_host_ex(ftl_handle h, const void* actions, const FtlStepInputs* in, const FtlOutputs* out, void*);
int ftl_step_host(ftl_handle h, const void* actions, const FtlOutputs* out, void*) {
    return ftl_step_host_ex(h, actions, nullptr, out, nullptr);
}
int ftl_step_host_ex(ftl_handle h, const void* actions, const FtlStepInputs* in, const FtlOutputs* out, void*) {
    DevOutputs o = dev_out(out);
    struct Scope {   // the per-step inputs are visible to the device functions through the DevState, for this step only
        DevState& st;
        Scope(DevState& s, const FtlStepInputs* in) : st(s) {
            st.in_frames = in ? in->frames_per_step : nullptr;
            st.in_draws = in ? in->regime_draws : nullptr;
        }
        ~Scope() { st.in_frames = nullptr; st.in_draws = nullptr; }
    } scope(h->st, in);
    switch (h->cfg.c.n_bears) {
        case 0: step_all<0>(h, actions, o); break;
        case 1: step_all<1>(h, actions, o); break;
        case 2: step_all<2>(h, actions, o); break;
        case 3: step_all<3>(h, actions, o); break;
        default: step_all<4>(h, actions, o); break;
    }
    optional_all(h, o);
    rays_all(h, o.rays);
    return FTL_OK;
}

int ftl_get_state(ftl_handle h, int32_t first, int32_t count, const FtlStateBuffers* b) {
    const FtlConfig& c = h->cfg.c;
    if (b->env) for (int j = 0; j < count; j++) pack_env(h->st, first + j, b->env[j]);
    if (b->trail) memcpy(b->trail, h->st.trail + (size_t)first * c.trail_cap, sizeof(float2) * (size_t)c.trail_cap * count);
    if (b->hist) memcpy(b->hist, h->st.hist + (size_t)first * c.corridor_cap, sizeof(double2) * (size_t)c.corridor_cap * count);
    if (b->corridor) memcpy(b->corridor, h->st.corridor + (size_t)first * c.corridor_cap, sizeof(float4) * (size_t)c.corridor_cap * count);
    return FTL_OK;
}
int ftl_set_state(ftl_handle h, int32_t first, int32_t count, const FtlStateBuffers* b) {
    const FtlConfig& c = h->cfg.c;
    memcpy(h->st.trail + (size_t)first * c.trail_cap, b->trail, sizeof(float2) * (size_t)c.trail_cap * count);
    if (b->hist) memcpy(h->st.hist + (size_t)first * c.corridor_cap, b->hist, sizeof(double2) * (size_t)c.corridor_cap * count);
    if (b->corridor) memcpy(h->st.corridor + (size_t)first * c.corridor_cap, b->corridor, sizeof(float4) * (size_t)c.corridor_cap * count);
    for (int j = 0; j < count; j++) unpack_env(h->cfg, h->st, first + j, b->env[j]);
    return FTL_OK;
}
}  // extern "C"
