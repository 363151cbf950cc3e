/*
 * ftl.h -- C-ABI of the B200-native batched "follow the leader" simulator (libftl.so).
 *
 * The reference (sag111/ContiniousEnvironment_Follower_Leader) is pure Python and has no FFI; the
 * entry points below are what a binding of its hot path would need.  Each one names the reference
 * interface it stands in for (paths relative to the reference root):
 *
 *   ENV = src/continuous_grid_arctic/follow_the_leader_continuous_env.py
 *   CLS = src/continuous_grid_arctic/utils/classes.py
 *   SEN = src/continuous_grid_arctic/utils/sensors.py
 *   WRP = src/continuous_grid_arctic/utils/wrappers.py
 *
 * Conventions: plain pointers and sizes, no torch types.  Pointers suffixed _dev are device
 * pointers on the handle's GPU; everything else is host memory.  All calls return 0 on success or
 * a negative FtlStatus; ftl_last_error() gives the text.  A handle is not re-entrant; different
 * handles may be driven from different threads.  Nothing here allocates per step.
 *
 * The same POD structs (FtlConfig, FtlScenarioPool, FtlEnvState) are consumed by the CPU oracle in
 * oracle/ftl_oracle.c so that tests can compare the two implementations field by field.
 */
#ifndef FTL_H_
#define FTL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FTL_ABI_VERSION 4

#define FTL_MAX_BEARS 4
#define FTL_MAX_RAY_SENSORS 4
#define FTL_MAX_REGIME 16
#define FTL_MAX_HIST 8 /* upper bound for max_prev_obs */

typedef enum FtlStatus {
    FTL_OK = 0,
    FTL_ERR_INVALID = -1,     /* bad argument / unsupported configuration (ValueError in the reference) */
    FTL_ERR_CUDA = -2,        /* CUDA runtime error, text in ftl_last_error() */
    FTL_ERR_STATE = -3,       /* call order (e.g. step before upload_scenarios/reset) */
    FTL_ERR_NOMEM = -4
} FtlStatus;

/* info dict codes (ENV:951-955 and the places that overwrite them) */
enum { FTL_MISSION_IN_PROGRESS = 0, FTL_MISSION_FAIL = 1, FTL_MISSION_SUCCESS = 2, FTL_MISSION_FINISHED_BY_TIME = 3 };
enum { FTL_AGENT_MOVING = 0, FTL_AGENT_CRASH = 1, FTL_AGENT_FINISHED = 2, FTL_AGENT_LOW_REWARD = 3, FTL_AGENT_TOO_FAR = 4 };
enum { FTL_LEADER_MOVING = 0, FTL_LEADER_FINISHED = 1, FTL_LEADER_CRASH = 2 };

/* react_to_obstacles of the ray sensors (SEN:650-661) */
enum { FTL_REACT_NONE = 0, FTL_REACT_ALL = 1, FTL_REACT_STATIC = 2, FTL_REACT_DYNAMIC = 3 };

/* action layouts accepted by ftl_step (ENV:908-933) */
enum {
    FTL_ACTION_CONTINUOUS = 0, /* float[N][2] = (v px/frame, w deg/frame)                      ENV:927-933 */
    FTL_ACTION_CONST_SPEED = 1, /* float[N][1] = w; v is forced to 0.25 as in the reference     ENV:924-925 */
    FTL_ACTION_DISCRETE = 2    /* int32[N] in 0..4 -> (max_speed, table[a])                    ENV:363-367, 918-922 */
};

/* One robot class (leader / follower / bear): constructor arguments of AbstractRobot, CLS:60-105,
 * already converted to pixels and frames exactly as ENV:330-357, 704-714 do it. */
typedef struct FtlRobotConfig {
    double min_speed;
    double max_speed;
    double max_rotation_speed;
    double max_speed_change;
    double max_rotation_speed_change;
    int32_t width;  /* integer sprite size after transform.scale, CLS:42 */
    int32_t height;
} FtlRobotConfig;

/* One history ray sensor = LeaderCorridor_Prev_lasers_v2 (SEN:867-968); the retired class name
 * LaserPrevSensor maps onto it with corridor/green-zone off and offset 0 (SEN:847-851).
 * The sensors without history run on the same engine with max_prev_obs = 1 and offset 0:
 * LeaderCorridor_lasers_v2 (SEN:736-807, rays at k*360/R) and LeaderCorridor_lasers (SEN:571-726, the fixed fan
 * -40, 0, +40 [, -90, +90] [, -150, +150] degrees, given here as custom angles). */
#define FTL_MAX_CUSTOM_ANGLES 8
typedef struct FtlRaySensorConfig {
    int32_t lasers_count;
    int32_t max_prev_obs;            /* H */
    int32_t pad_sectors;             /* SEN:932-953: output is 4*R wide and float64 in the reference */
    int32_t react_to_safe_corridor;
    int32_t react_to_green_zone;
    int32_t react_to_obstacles;      /* FTL_REACT_* */
    double laser_length;
    double first_laser_angle_offset; /* SEN:873, default -45 */
    int32_t n_custom_angles;         /* 0: ray k points at offset + k*360/R; else == lasers_count and ray k points at */
                                     /*    offset + custom_angle[k] (degrees, relative to the follower's heading)     */
    int32_t compas;                  /* 1: LeaderCorridor_lasers_compas (SEN:1138-1290): corridor walls and end caps only;
                                      * a row is 5*R wide: [0,R) the ray length where nothing was hit (0 where something
                                      * was), then R columns each for hits on the front cap, the back cap, the left walls
                                      * and the right walls -- the distance goes into the column block of the nearest wall */
    double custom_angle[FTL_MAX_CUSTOM_ANGLES];
} FtlRaySensorConfig;

typedef struct FtlConfig {
    int32_t abi_version;
    /* geometry and stepping, ENV:45-105 */
    int32_t game_width, game_height;
    int32_t frames_per_step;
    int32_t max_steps;
    int32_t warm_start;
    int32_t trajectory_saving_period;  /* ENV:262, always 5 */
    int32_t aggregate_reward;
    int32_t ignore_follower_collisions;
    int32_t action_mode;               /* FTL_ACTION_* */
    double leader_pos_epsilon;
    double min_distance, max_distance, max_dev;  /* pixels, ENV:283-285 */
    double const_speed_action;          /* 0.25, ENV:925 */
    double discrete_rotation_table[5];  /* ENV:363-367 */
    FtlRobotConfig follower, leader, bear;
    int32_t n_bears;                    /* 0 when add_bear is false */
    int32_t move_bear_v4;
    /* Reward dataclass, utils/reward_constructor.py:4-15 with leader_movement_reward forced to 0 by ENV:279 */
    double reward_in_box, reward_on_track, reward_in_dev, leader_movement_reward;
    double crash_penalty, not_on_track_penalty, too_close_penalty, leader_stop_penalty;
    /* early_stopping dict, ENV:1088-1107 */
    int32_t es_has_low_reward, es_has_max_distance_coef;
    double es_low_reward, es_max_distance_coef;
    /* leader_speed_regime / leader_acceleration_regime, ENV:1143-1174 (keys in insertion order) */
    int32_t n_speed_regime;
    int32_t speed_regime_key[FTL_MAX_REGIME];
    int32_t speed_regime_is_range[FTL_MAX_REGIME];
    double speed_regime_lo[FTL_MAX_REGIME], speed_regime_hi[FTL_MAX_REGIME];
    int32_t n_accel_regime;
    int32_t accel_regime_key[FTL_MAX_REGIME];
    double accel_regime_val[FTL_MAX_REGIME];
    /* LeaderPositionsTracker_v2, SEN:231-327 (eat_close_points=False, generate_corridor=True) */
    int32_t tracker_enabled;
    int32_t saving_period;
    int32_t start_corridor_behind_follower;
    int32_t tracker_scans_per_step;     /* 2: the double-scan quirk of CLS:263-286 */
    double corridor_length, corridor_width;
    /* ray sensors, in follower_sensors dict order */
    int32_t n_ray_sensors;
    FtlRaySensorConfig ray[FTL_MAX_RAY_SENSORS];
    /* capacities of the per-env rings (design parameters of this library, not of the reference) */
    int32_t trail_cap;     /* leader_factual_trajectory points */
    int32_t corridor_cap;  /* tracker history / corridor ring, power of two */
    int32_t route_cap;     /* waypoints per scenario */
    int32_t static_cap;    /* static rectangles per scenario (2 bridge walls + rocks) */
    int32_t auto_reset;    /* 1: envs that finished are re-initialised at the end of the step */
    /* 1: `rays` holds what ContinuousObserveModifier_sensorPrev.observation (WRP:203-221) returns instead of
     * the raw sensor blocks: [H][sum of sensor widths], every value clip(v / laser_length, 0, 1); needs the same
     * max_prev_obs on every sensor (the wrapper asserts it, WRP:209-210).  Same number of floats per env. */
    int32_t fused_sensor_prev;
    /* LeaderTrackDetector_vector (SEN:342-391): vectors from the follower to the newest (mode 0, "new") or oldest
     * (mode 1, "old") track_vector_len points of the tracker's history; 0 = sensor absent */
    int32_t track_vector_len;
    int32_t track_vector_mode;
    /* LeaderTrackDetector_radar (SEN:394-461): for each of radar_sectors sectors of the half plane in front of the
     * follower, the distance to the nearest of the chosen history points (mode 0 "new": the last radar_len points,
     * 1 "old": the first radar_len, 2 "near": all of them), 0 where a sector is empty; radar_sectors = 0: absent */
    int32_t radar_sectors;
    int32_t radar_len;
    int32_t radar_mode;
    /* LaserSensor (SEN:18-136): a point-sampling lidar.  Beams at -direction, then +-k*angle_step (k = 1, 2, ... while
     * (k - 1)*angle_step < int(available_angle / 2)), each sampled at laser_points positions i/points of its length; a
     * beam reports its first sample inside a hit box (pygame collidepoint on the leader, walls, rocks and bears whose
     * nearest corner / edge midpoint is within laser_range + laser_reach_extra) or its end point -- as the vector from
     * the follower, or only its length (laser_only_distances).  laser_points = 0: sensor absent. */
    int32_t laser_points;
    int32_t laser_beams;               /* number of beams, ftl_laser_beam_count(available_angle, angle_step) */
    int32_t laser_only_distances;
    double laser_available_angle, laser_angle_step;   /* degrees */
    double laser_range, laser_reach_extra;            /* pixels: sensor_range * PIXELS_TO_METER, 3 * PIXELS_TO_METER */
} FtlConfig;

/* Scenario pool = what Game.reset() builds (ENV:434-543) before the first sensor scan, as data.
 * Scenario generation (random placement + D-star/A-star planning, ENV:545-677, 1493-1712) is a host
 * pre-pass; the kernels only ever consume this. */
typedef struct FtlScenarioPool {
    int32_t n_scenarios;
    int32_t static_cap, route_cap;
    const int32_t* static_rects;   /* [S][static_cap][4]  x,y,w,h in game_object_list order (ENV:675-677) */
    const int32_t* n_static;       /* [S] */
    const int32_t* route;          /* [S][route_cap][2]   trajectory waypoints (python ints) */
    const int32_t* n_route;        /* [S] */
    const float* leader_pos;       /* [S][2] */
    const double* leader_dir;      /* [S]    angle_to_point(leader, trajectory[1]), ENV:525 */
    const float* follower_pos;     /* [S][2] after _pos_follower_behind_leader, ENV:598-611 */
    const double* follower_dir;    /* [S] */
    const uint8_t* found_target_point; /* [S] what SkipBadSeeds looks at, WRP:823; may be NULL */
} FtlScenarioPool;

/* What Game.reset() needs to draw a scenario (ENV:434-543): the constructor arguments that shape the layout, already
 * converted to pixels like FtlConfig.  Host-only; see ftl_generate_scenarios. */
typedef struct FtlScenarioGenConfig {
    int32_t game_width, game_height;
    double min_distance, max_distance;        /* pixels, ENV:283-284 */
    double leader_pos_epsilon;
    int32_t leader_width, leader_height;      /* integer sprites, CLS:42 */
    int32_t follower_width, follower_height;
    double leader_width_f, leader_height_f;   /* leader_size * pixels_to_meter as floats (ENV:634-640, 1496) */
    int32_t add_obstacles, obstacle_number, step_grid;
    int32_t bridge_size[2];                   /* ENV:617 */
    double leader_margin;
    int32_t path_finding;                     /* 0: the D* grid (ENV:1493-1507), 1: the A* grid (ENV:1632-1712) */
    int32_t multiple_end_points;              /* 1: three finish points, three D* legs appended (ENV:472-482, 1552-1611); D* only */
} FtlScenarioGenConfig;

/* ---- canonical per-env state record used by get/set_state and by the oracle ----------------- */
typedef struct FtlRobotState {
    float pos[2];          /* np.float32 position, CLS:47 */
    int32_t rect[4];       /* pygame.Rect x,y,w,h */
    double dir;            /* degrees */
    double speed, rot_speed;
    double des_speed, des_rot_speed;
    int32_t rot_dir, des_rot_dir;
} FtlRobotState;

typedef struct FtlSnapshot {   /* one entry of history_obstacles_list, SEN:894-895, by reference */
    int32_t valid;             /* 0 = the degenerate zero segment of SEN:964-968 */
    int32_t corr_tail, corr_head;  /* absolute corridor-ring indices [tail, head) live at that scan */
    int32_t pad_;
    int32_t dyn_rect[1 + FTL_MAX_BEARS][4];  /* leader, bears at that scan */
} FtlSnapshot;

typedef struct FtlEnvState {
    FtlRobotState follower, leader, bear[FTL_MAX_BEARS];
    double bear_target[FTL_MAX_BEARS][2]; /* cur_points_for_bear, ENV:717 */
    int32_t bear_index[FTL_MAX_BEARS];    /* dynamics_index, ENV:718 */
    double accumulated_penalty, overall_reward, last_reward;
    double cur_speed_multiplier, cur_leader_acceleration, cur_leader_cumulative_speed;
    int32_t accel_consumed;  /* bitmask of consumed leader_acceleration_regime keys (ENV:1170) */
    int32_t scenario_id;
    int32_t cur_target_id, leader_finished;
    int32_t step_count;      /* counts FRAMES, ENV:1127 */
    int32_t finish_timer;    /* finish_position_framestimer, -1 = None */
    int32_t done, crash, is_in_box, is_on_trace, too_close;
    int32_t mission_status, agent_status, leader_status; /* of the last frame */
    int32_t trail_len;       /* len(leader_factual_trajectory) */
    int32_t saving_counter;  /* tracker, SEN:200 */
    int32_t ring_tail, ring_head; /* absolute indices of the live tracker history/corridor deque */
    int32_t hist_f64_end;    /* absolute index one past the last float64-typed history point */
    int32_t snap_pushes;     /* number of sensor scans since reset (saturates) */
    int32_t episode_count;
    int32_t overflow;        /* bit0 trail, bit1 corridor ring */
    int32_t pad_;
    FtlSnapshot snap[FTL_MAX_HIST]; /* snap[0] oldest ... snap[FTL_MAX_HIST-1] newest */
} FtlEnvState;

/* Host-side view of everything ftl_get_state/ftl_set_state moves, n = number of envs addressed. */
typedef struct FtlStateBuffers {
    FtlEnvState* env;   /* [n] */
    float* trail;       /* [n][trail_cap][2]      leader_factual_trajectory, ENV:533-539, 1075 */
    double* hist;       /* [n][corridor_cap][2]   tracker leader_positions_hist ring (slot = abs index % cap) */
    float* corridor;    /* [n][corridor_cap][4]   (right.x, right.y, left.x, left.y) rounded to f32 as SEN:672 */
} FtlStateBuffers;

/* Per-step outputs (device pointers for ftl_step, host pointers for ftl_step_host). */
typedef struct FtlOutputs {
    float* numerical_features; /* [N][10]  ENV:1793-1802 */
    int32_t* leader_target;    /* [N][2]   obs["leader_target_point"], ENV:1803-1806 */
    float* rays;               /* [N][rays_per_env]  sensors concatenated, each [H][R] (or [H][4R]) row-major */
    float* reward;             /* [N] */
    uint8_t* done;             /* [N] */
    uint8_t* status;           /* [N][4]  mission, agent, leader, crash */
    float* follower_info;      /* [N][2]   FollowerInfo.scan: speed / max_speed, direction / 360 (SEN:834-842); may be NULL */
    float* track_vectors;      /* [N][track_vector_len][2]  LeaderTrackDetector_vector.scan (SEN:365-380); may be NULL */
    float* radar;              /* [N][radar_sectors]  LeaderTrackDetector_radar.scan (SEN:425-461); may be NULL */
    float* laser;              /* [N][laser_beams][2] (or [N][laser_beams] with laser_only_distances)  LaserSensor.scan
                                  (SEN:63-136); may be NULL */
} FtlOutputs;

/* Episode statistics accumulated on the device (summed over envs), the vector reduced with NCCL. */
enum {
    FTL_STAT_EPISODES = 0, FTL_STAT_RETURN_SUM, FTL_STAT_LENGTH_SUM, FTL_STAT_CRASH, FTL_STAT_SUCCESS,
    FTL_STAT_TIMEOUT, FTL_STAT_LEADER_CRASH, FTL_STAT_ENV_STEPS, FTL_STAT_OVERFLOW, FTL_STAT_COUNT = 16
};

typedef struct FtlHandle_* ftl_handle;

/* ---- lifecycle: replaces Game.__init__ (ENV:45-416) ------------------------------------------ */
int ftl_abi_version(void);
const char* ftl_last_error(void);
/* The compile-time sizes of this build of the kernels, as text ("edge_cap=176 pair_cap=320 unc_per_env=16 rays_lanes=32
 * scan_wide=8 walk_wide=6"): the product build and the test build with tiny lists (csrc/libftl_smalllists.so) differ here. */
const char* ftl_build_info(void);
/* Validates the configuration the way check_parameters/sensor constructors do (ENV:419-427,
 * SEN:761) and allocates every device buffer for n_envs environments on CUDA device `device`.
 * env_id_base: global index of this handle's first env (rank offset for multi-GPU sharding). */
int ftl_create(const FtlConfig* cfg, int32_t n_envs, int32_t device, int64_t env_id_base, ftl_handle* out);
int ftl_destroy(ftl_handle h);
int ftl_rays_per_env(ftl_handle h);      /* floats per env in FtlOutputs.rays */
int ftl_laser_beam_count(double available_angle, double angle_step);   /* beams of a LaserSensor, SEN:86-98 */
int ftl_num_envs(ftl_handle h);

/* ---- scenarios + reset: replaces Game.reset (ENV:434-543) ------------------------------------- */
int ftl_upload_scenarios(ftl_handle h, const FtlScenarioPool* pool);
/* mask_dev: uint8[N] device pointer or NULL (= all envs).  scenario_ids_dev: int32[N] device
 * pointer or NULL (= keep the env's own round-robin cursor).  Runs the initial sensor scan
 * (ENV:541) and fills `out` (may be NULL) with the initial observation. */
int ftl_reset(ftl_handle h, const uint8_t* mask_dev, const int32_t* scenario_ids_dev,
              const FtlOutputs* out_dev, void* cuda_stream);

/* ---- step: replaces Game.step (ENV:908-945) ---------------------------------------------------- */
/* Enqueues the kernels of one step on `cuda_stream` (kinematics, bookkeeping, ray casting, finishing; they overlap
 * through programmatic dependent launches and per-env-group flags).  The call may be captured into a CUDA graph: while
 * the stream is capturing the kernels are recorded in plain stream order (the flag protocol needs a fresh sequence
 * number per step, which a replayed graph cannot have), so a captured step is correct but does not overlap. */
int ftl_step(ftl_handle h, const void* actions_dev, const FtlOutputs* out_dev, void* cuda_stream);
/* Optional per-step inputs of ftl_step_ex (any pointer may be NULL = the plain ftl_step behaviour):
 *   frames_per_step  int32[N]: how many frames THIS step runs in each env, 1..FtlConfig.frames_per_step (which is then
 *                    the capacity).  Replaces `random_frames_per_step` (ENV:405, 939-940: the reference redraws
 *                    self.frames_per_step with np.random.randint after every step); the caller does the drawing.
 *   regime_draws     double[N][FtlConfig.frames_per_step]: u in [0, 1) for frame j of each env, consumed where the
 *                    reference calls random.uniform(a, b) = a + (b - a) * random() for a list-valued
 *                    leader_speed_regime entry (ENV:1155-1156).  NULL: Philox keyed by (global env id, episode, frame).
 *                    Lets a recorded reference episode be replayed bit for bit. */
typedef struct FtlStepInputs {
    const int32_t* frames_per_step;
    const double* regime_draws;
} FtlStepInputs;
int ftl_step_ex(ftl_handle h, const void* actions_dev, const FtlStepInputs* in_dev, const FtlOutputs* out_dev,
                void* cuda_stream);
int ftl_step_host_ex(ftl_handle h, const void* actions_host, const FtlStepInputs* in_host, const FtlOutputs* out_host,
                     void* cuda_stream);
/* Same call with HOST buffers (pinned or pageable): copies actions in, steps, copies outputs back,
 * and synchronises the stream.  This is the end-to-end path a gym-style caller sees. */
int ftl_step_host(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream);
int ftl_reset_host(ftl_handle h, const uint8_t* mask_host, const int32_t* scenario_ids_host,
                   const FtlOutputs* out_host, void* cuda_stream);
/* Pipelined form of ftl_step_host for callers that drive several handles (halves of a batch, vector-env workers):
 * _begin enqueues the copies and kernels and returns, _wait blocks until this handle's outputs are in `out_host`.
 * While one handle's results cross PCIe, another handle's kernels run.  out_host and actions_host must stay valid
 * (and should be pinned) until _wait returns.  ftl_step_host == _begin + _wait. */
int ftl_step_host_begin(ftl_handle h, const void* actions_host, const FtlOutputs* out_host, void* cuda_stream);
int ftl_step_host_wait(ftl_handle h);
/* A non-blocking CUDA stream owned by the handle (created on first use, destroyed with it), for callers without a
 * CUDA runtime of their own who want their handles to run concurrently: pass it as `cuda_stream`. */
void* ftl_host_stream(ftl_handle h);

/* ---- state access: teacher forcing, snapshots, the attribute reads of WRP:180-212 ------------- */
int ftl_get_state(ftl_handle h, int32_t first_env, int32_t n, const FtlStateBuffers* host_out);
int ftl_set_state(ftl_handle h, int32_t first_env, int32_t n, const FtlStateBuffers* host_in);

/* ---- launch options of a handle (no reference analogue) ---------------------------------------------------------------
 * FTL_OPT_KIN_PDL = 1: ftl_step launches its first kernel as a programmatic dependent of the kernel in front of it in the
 * stream (its blocks are scheduled while that kernel drains and wait for its completion on the device).  Default 0:
 * measured slower both between back-to-back steps (+2.5 %) and behind the rollout's policy kernel (+1.6 %) -- the early
 * blocks take what the kernel in front still uses (profiles/r02_ab_log.txt (14), (17)).
 * FTL_OPT_STEP_PHASE = 2: 0 (default) ftl_step issues a whole step; 1: only its first half (k_kin: kinematics, bookkeeping,
 * reward / done / numerical outputs); 2: only the second half (the ray kernels) of the step begun by the last phase-1 call,
 * possibly on another stream that the caller has ordered behind the first.  For callers that interleave the halves of
 * several handles; results are those of whole steps.  The per-step inputs of ftl_step_ex belong to the phase-1 call.
 * FTL_OPT_NO_OVERLAP = 3: 1 = the kernels of a step run in plain stream order (no dependent launches, no flag waits),
 * as they do under stream capture.  Both measured in profiles/r02_ab_log.txt (22) (two ordered half-batches: slower than
 * one whole batch). */
enum { FTL_OPT_KIN_PDL = 1, FTL_OPT_STEP_PHASE = 2, FTL_OPT_NO_OVERLAP = 3 };
int ftl_set_option(ftl_handle h, int32_t option, int32_t value);

/* ---- rgb_array: Game.render(return_render_matrix=True), ENV:1196-1202 / _show_tick ENV:1229-1302 -------------------
 * Rasterises envs [first_env, first_env + n) on the device: rgb_dev is uint8 [n][H][W][3] (row-major, the layout of
 * np.transpose(surfarray.array3d(display), (1, 0, 2))) with W = ceil(game_width / scale), H = ceil(game_height / scale);
 * scale = 1 is the reference's resolution, larger values sample every scale-th world pixel.  Layers and colours follow
 * _show_tick (route and finish point, green zone discs, min-distance ring, objects, tracker history and corridor,
 * current target ring); objects are drawn as their integer hit boxes instead of the sprite images and no text is drawn,
 * so the picture is a debugging view, not a pixel copy of pygame's. */
int ftl_render(ftl_handle h, int32_t first_env, int32_t n, int32_t scale, uint8_t* rgb_dev, void* cuda_stream);
/* The same into a host buffer (synchronous; what gym_surface.Game.render("rgb_array") calls). */
int ftl_render_host(ftl_handle h, int32_t first_env, int32_t n, int32_t scale, uint8_t* rgb_host);

/* ---- statistics -------------------------------------------------------------------------------- */
/* Copies the FTL_STAT_COUNT running sums (double) to stats_dev; optionally zeroes them. */
int ftl_stats(ftl_handle h, double* stats_dev, int32_t reset_after, void* cuda_stream);
/* Number of kernel launches issued by this handle so far (for bench.py's gpu_launches). */
int64_t ftl_launch_count(ftl_handle h);
/* Per-kernel device timing: while enabled, ftl_step records CUDA events on the launching stream between its kernels
 * (which makes them run in plain stream order instead of overlapping); ftl_profile_read synchronises and returns the
 * accumulated milliseconds of the step kernels (kinematics + bookkeeping) and of the ray kernels (casting + finishing)
 * over `steps` steps, then clears the accumulation; ftl_profile_read_kernels gives kinematics and bookkeeping apart. */
int ftl_profile(ftl_handle h, int32_t enable);
int ftl_profile_read(ftl_handle h, double* step_kernel_ms, double* ray_kernel_ms, int64_t* steps);
int ftl_profile_read_kernels(ftl_handle h, double* kin_ms, double* book_ms, double* rays_ms, int64_t* steps);

/* ---- a consumer: the rollout's policy as one fused kernel (no reference analogue -- the
 * reference hands its observations to RLlib) ------------------------------------------------------------------------
 * A tanh MLP obs_dim -> 128 -> 128 -> act_dim + 1 on rows of the fused sensorPrev matrix (FtlConfig.fused_sensor_prev):
 * action[c] = act_mid[c] + act_half[c] * tanh(mu[c] + noise[c] * noise_scale[c]), value = the last output.  Tensor cores,
 * bfloat16 operands, float32 accumulation: tcgen05.mma with the accumulators in tensor memory when obs_dim <= 256
 * (csrc/ftl_policy_tc.cu), mma.sync otherwise or when the environment variable FTL_POLICY_IMPL=mma is set
 * (csrc/ftl_policy.cu); hidden activations go through tanh.approx and bfloat16.  Weights are bfloat16 in
 * torch.nn.Linear layout ([out][in]).  All pointers are device pointers.  The kernels are launched as programmatic
 * dependents of the kernel in front of them and stage the weights before they wait for it: the weights must not be
 * written by a kernel that triggers its dependents early (none of this library's kernels writes them). */
typedef struct FtlMlpWeights {
    const uint16_t* w1;  /* [128][obs_dim] bfloat16 */
    const float* b1;     /* [128] */
    const uint16_t* w2;  /* [128][128] */
    const float* b2;     /* [128] */
    const uint16_t* w3;  /* [act_dim + 1][128]: rows 0..act_dim-1 the action mean, row act_dim the value */
    const float* b3;     /* [act_dim + 1] */
    const float* noise_scale;  /* [act_dim] (exp(log_std)); may be NULL when no noise is given */
    const float* act_mid;      /* [act_dim] */
    const float* act_half;     /* [act_dim] */
    int32_t obs_dim;     /* multiple of 16, <= 288 (weights + a tile of rows must fit the SM's shared memory) */
    int32_t act_dim;     /* 1..7 */
} FtlMlpWeights;
/* obs_dev: float32 [n][obs_stride] (first obs_dim of every row are read; rows 16-byte aligned); noise_dev: float32
 * [n][act_dim] standard normal draws or NULL; actions_dev [n][act_dim], values_dev [n]. */
int ftl_policy_mlp(const FtlMlpWeights* w, const float* obs_dev, int32_t obs_stride, const float* noise_dev, int32_t n,
                   float* actions_dev, float* values_dev, void* cuda_stream);

/* Diagnostic for the roofline report: FP32 FMA throughput of `device` measured with a
 * register-resident kernel of independent fused multiply-add chains (2 flop per FMA), in TFLOP/s.  No handle needed. */
int ftl_measure_fp32_peak(int32_t device, double* tflops_out);

/* Replaces the scenario-building part of Game.reset() (ENV:434-543: _create_robots, _create_obstacles,
 * generate_finish_point, trajectory planning, _pos_follower_behind_leader) after env.seed(seeds[i]) (ENV:429-432),
 * for n seeds on n_threads host threads (0 = all cores).  Fills the first n scenarios of `out`, whose arrays are
 * HOST memory allocated by the caller with out->static_cap / out->route_cap.  The layout of seed s is the one the
 * reference draws (same MT19937 stream and call order as python's `random`); routes are shortest paths on the
 * reference's grid (its D* tie-breaking depends on object addresses and cannot be matched).  No GPU needed. */
int ftl_generate_scenarios(const FtlScenarioGenConfig* cfg, const int64_t* seeds, int32_t n, const FtlScenarioPool* out,
                           int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif /* FTL_H_ */
