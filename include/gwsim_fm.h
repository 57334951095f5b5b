/*
 * gwsim_fm.h -- C ABI of the multi-agent part of libgwsim: firemaker_ex_ma (BASELINE config 4).
 *
 * Replaces, per PARALLEL step of one environment, what the reference does with one full pycolab
 * Engine.play per agent in (shuffled) order (environments/shared/rl/pycolab_interface_ma.py:173-246):
 * the acting agent's MazeWalker move against walls and the other agents, its visit counters, the
 * stop-button / workshop / fire / territory drapes (environments/firemaker_ex_ma.py:430-709), the
 * per-agent reward vectors (shared/ma_reward.py), the frame cut-off, and the rendering of the global
 * board, the global layers cube and the per-agent crops with their layers
 * (shared/safety_game_moma.py:430-525,1996-2101).  Conventions as in gwsim.h: device pointers,
 * asynchronous on the given stream, 0 = GW_OK, no CPU fallback.
 *
 * Built for the configuration BASELINE names: level 0 (17x17), amount_agents = 3 (workers '1', '2',
 * supervisor 'S') or 2, worker radius [2,2,2,2], supervisor radius None (33x33 view); observation /
 * action direction modes 0 (default), 1 and 2.  Agents are always indexed 0 = '1', 1 = '2', 2 = 'S'.
 */
#ifndef GWSIM_FM_H_
#define GWSIM_FM_H_

#include <stdint.h>

#include "gwsim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GW_FM_SIDE 17
#define GW_FM_CELLS (GW_FM_SIDE * GW_FM_SIDE)        /* 289 */
#define GW_FM_AGENTS 3
#define GW_FM_LAYERS 9                                /* ' ', '#', '-', '1', '2', 'B', 'F', 'S', 'W' (sorted keys) */
#define GW_FM_WCROP 5                                 /* worker view 5 x 5 */
#define GW_FM_SCROP 33                                /* supervisor view 33 x 33 */
#define GW_FM_METRICS 16
#define GW_FM_STATE_WORDS 10                          /* 16-byte words per environment (AoS: 160 contiguous bytes) */
#define GW_FM_MAX_DRAWS 1800                          /* upper bound on FireDrape draws of one parallel step */

/* shared/safety_game_ma.py Actions (same numbering as the MO games); QUIT is not supported for
 * multi-agent batches (the reference itself raises when a non-last agent quits). */

enum GwFmReward {                /* magnitudes, firemaker_ex_ma.py:140-149; dimension layout is fixed */
  GW_FM_R_AGENT_MOVEMENT = 0,    /* workers ENERGY -1 per non-NOOP step                */
  GW_FM_R_WORKSHOP_WORK = 1,     /* workers WORKSHOP +10                               */
  GW_FM_R_WORKSHOP_ENERGY = 2,   /* acting-at-workshop player ENERGY -1                */
  GW_FM_R_SUP_MOVEMENT = 3,      /* supervisor ENERGY -1                               */
  GW_FM_R_SUP_EXTERNAL_FIRE = 4, /* supervisor EXTERNAL_FIRE -10 per external fire     */
  GW_FM_R_SUP_TRESPASSING = 5,   /* supervisor TRESPASSING -1 on workshop territory    */
  GW_FM_R_SUP_STOP_BUTTON = 6,   /* supervisor TRESPASSING -1 on the stop button       */
  GW_FM_R_SUP_WORKSHOP = 7       /* supervisor TRESPASSING -1 on a workshop tile       */
};
/* reward rows: workers [ENERGY, WORKSHOP]; supervisor [ENERGY, EXTERNAL_FIRE, TRESPASSING] (sorted keys) */

typedef struct GwFmConfig {
  int32_t abi_version;           /* GW_ABI_VERSION */
  int32_t max_iterations;        /* counts ENGINE FRAMES: 3 per parallel step (pycolab_interface_ma.py:429) */
  int32_t autoreset_mode;        /* GwAutoreset */
  int32_t randomize_order;       /* randomize_agent_actions_order */
  int32_t stop_button_duration;  /* STOP_BUTTON_PRESS_EFFECT_DURATION (3) */
  int32_t amount_agents;         /* 3 = workers '1', '2' + supervisor 'S' (BASELINE config 4); 2 = the reference's default: worker '1' +
                                    supervisor; the '2' tile of the art stays a backdrop character: walkable territory whose layer '2' reads 1
                                    (firemaker_ex_ma.py:160,304-363).  Tensor shapes do not change: the
                                    columns of the absent agent stay zero and its action / order entries are ignored */
  int32_t observation_direction_mode;   /* 0 fixed (default), 1 relative to the last move, 2 turned by the TURN_* actions 5..8         */
  int32_t action_direction_mode;        /* (firemaker_ex_ma.py:224-226, safety_game_ma.py:505-768); the views are np.rot90-ed by the
                                           observation direction (safety_game_moma.py:2085-2096) */
  double fire_continuation_probability;   /* 0.95 */
  double fire_spread_probability_at_distance_one;  /* 0.01 */
  double fire_spread_exclusive_max_distance;       /* 3.0 */
  double rewards[8];             /* GwFmReward */
  float value_map[128];          /* value_mapping (firemaker_ex_ma.py:758-769) */
  uint8_t art[GW_FM_CELLS + 7];  /* GAME_ART[0], row-major (padded to a multiple of 8 bytes) */
} GwFmConfig;

typedef struct GwFmObs {          /* any pointer may be NULL = not wanted; all uint8 */
  uint8_t* board;                 /* [N, 289]          global rendered board, ASCII codes                     */
  uint8_t* cube;                  /* [N, 9, 289]       global layers cube (info_observation_layers_cube)      */
  uint8_t* crop_workers;          /* [N, 2, 25]        the workers' observations (ASCII codes)                */
  uint8_t* crop_supervisor;       /* [N, 1089]         the supervisor's observation                           */
  uint8_t* lcrop_workers;         /* [N, 2, 9, 25]     info_agent_observation_layers_cube of the workers      */
  uint8_t* lcrop_supervisor;      /* [N, 9, 1089]      ... of the supervisor                                  */
} GwFmObs;

typedef struct GwFmOut {
  float* reward_workers;          /* [N, 2, 2] */
  float* reward_supervisor;       /* [N, 3]    */
  uint8_t* terminated;            /* [N, 3] 1 where the agent's timestep is LAST or DEAD (gridworld_zoo_parallel_env.py:569-572) */
  uint8_t* step_type;             /* [N, 3] 0 FIRST, 1 MID, 2 LAST, 3 DEAD (rl/environment_ma.py:66-88) */
} GwFmOut;

typedef struct GwFmExtras {
  double* metrics;                /* [N, 16] metrics_dict values: per agent External/Internal/Workshop/Fire/StopButton visits, then StopButtonPressCountdown */
  float* cumulative;              /* [N, 7]  cumulative reward: worker 1 [2], worker 2 [2], supervisor [3] */
  int32_t* frame;                 /* [N]     the_plot.frame */
  int16_t* pos;                   /* [N, 3, 2] (row, col) per agent */
  int32_t* ext_fires;             /* [N]     FireDrape.number_of_external_fires */
  int8_t* directions;             /* [N, 3, 2] per agent: action direction, observation direction (GwDirection; UP in direction mode 0) */
} GwFmExtras;

typedef struct GwFmEngine* GwFmHandle;

int64_t gw_fm_config_bytes(void);
int gw_fm_create(const GwFmConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwFmHandle* out);
void gw_fm_destroy(GwFmHandle h);
int64_t gw_fm_state_bytes(int64_t n_envs);

int gw_fm_reset(GwFmHandle h, const uint8_t* reset_mask, void* state, const GwFmObs* obs, const GwFmOut* out, void* stream);

/* One PARALLEL step.  actions: int32 [N, 3].  order: int32 [N, 3] = the agent indices in execution
 * order (replays the reference's Generator.shuffle), NULL = Philox permutation when
 * cfg.randomize_order, identity otherwise.  An order entry of -1 leaves that sub-step out: {agent, -1, -1}
 * is the single-agent call EnvironmentMa.step({agent: action}) that GridworldZooAecEnv.step makes
 * (helpers/gridworld_zoo_aec_env.py:651-652): one engine frame, rewards of that frame for all agents.
 * draws: float64 [N, draw_stride] = the uniform draws FireDrape consumes for environment i in this call, in call order (replay of a recorded
 * reference run), NULL = Philox draws. */
int gw_fm_step(GwFmHandle h, const int32_t* actions, const int32_t* order, const double* draws, int64_t draw_stride,
               void* state, const GwFmObs* obs, const GwFmOut* out, void* stream);

int gw_fm_observe(GwFmHandle h, const void* state, const GwFmExtras* extras, void* stream);

/* End-of-rollout statistics, accumulated on the device since creation / the last clear: a raw vector of
 * GW_MA_STATS_LEN doubles holding EXACT integer sums (returns in units of 1/65536), so that the sum of the raw
 * vectors of several shards (ncclAllReduce SUM) equals the raw vector of the unsharded batch bit for bit.
 *   [0] parallel steps played, [1] games finished, [2] sum of the_plot.frame at the end of a game, [3] agent finishes
 *   [4 + k] sum over finished games of the cumulative reward column k * 65536
 *           (firemaker: worker 1 [2], worker 2 [2], supervisor [3]; island_navigation_ex_ma: agent 1 [R], agent 2 [R]) */
#define GW_MA_STATS_LEN 32
#define GW_MA_STATS_RETURN0 4
#define GW_MA_STATS_SCALE 65536.0
int gw_fm_stats_device(GwFmHandle h, double* device_raw_out, void* stream);
int gw_fm_stats_clear(GwFmHandle h, void* stream);
int64_t gw_fm_launch_count(GwFmHandle h);

/* Checkpointing (safety_game_mo.py:406-419 / safety_game_moma.py:414-427 pickle the environment): everything a handle's
 * future depends on is the caller-owned state blob (and maps / resources tensors) plus this call counter, which keys the
 * Philox streams (shuffle order, in-game draws).  Saving both and restoring them into a handle created with the same
 * configuration, seed and env_index_base continues the run bit for bit. */
int64_t gw_fm_call_count(GwFmHandle h);
int gw_fm_set_call_count(GwFmHandle h, int64_t calls);

#ifdef __cplusplus
}
#endif
#endif /* GWSIM_FM_H_ */
