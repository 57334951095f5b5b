/*
 * gwsim_ima.h -- C ABI for island_navigation_ex_ma: two agents ('1', '2') sharing
 * the drink / food resources of island_navigation_ex, stepped through the multi-agent engine.
 *
 * Replaces, per environment instance and per PARALLEL step
 *   GridworldZooParallelEnv.step                  helpers/gridworld_zoo_parallel_env.py:429-615
 *     SafetyEnvironmentMoMa.step                  environments/shared/safety_game_moma.py:984-1047
 *       EnvironmentMa.step                        environments/shared/rl/pycolab_interface_ma.py:173-246
 *         Engine.play({agent: action}) per live agent, in (shuffled) order
 *           AgentSprite.update / update_reward    environments/island_navigation_ex_ma.py:481-712
 *             AgentSafetySprite.update            environments/shared/safety_game_ma.py:769-809 (relative actions :505-560)
 *           WaterDrape / DrinkDrape / FoodDrape   environments/island_navigation_ex_ma.py:715-845
 *       _process_timestep                         environments/shared/safety_game_moma.py:1183-1379
 *     agent_perspectives_with_layers              environments/shared/safety_game_moma.py:430-525,1996-2101 (rotated crops)
 * and the single-agent call GridworldZooAecEnv.step makes (order = {agent, -1}).
 *
 * The game TYPE is a GwConfig (include/gwsim.h) with env_type GW_ENV_ISLAND_NAVIGATION_EX_MA: same flag,
 * event and metric slots as island_navigation_ex (GW_ISL_*), the agents' start tiles read '1' and '2' in
 * `art`, and the multi-agent switches live in iparams[GW_IMA_I_*].  Conventions as in gwsim.h: device
 * pointers owned by the caller, asynchronous on the given stream, 0 = GW_OK, no CPU fallback.
 */
#ifndef GWSIM_IMA_H_
#define GWSIM_IMA_H_

#include "gwsim.h"
#include "gwsim_fm.h"      /* GW_MA_STATS_* */

#ifdef __cplusplus
extern "C" {
#endif

#define GW_ENV_ISLAND_NAVIGATION_EX_MA 16
#define GW_IMA_AGENTS 2
#define GW_IMA_CROP 5                    /* observation_radius [2,2,2,2]: 5 x 5 agent views */
#define GW_IMA_METRICS 16

enum GwImaIParam {                       /* iparams 0..3 are GW_ISL_I_* */
  GW_IMA_I_RANDOMIZE_ORDER = 4,          /* randomize_agent_actions_order (rl/pycolab_interface_ma.py:177-180) */
  GW_IMA_I_OBSERVATION_DIRECTION_MODE = 5,   /* 0 fixed, 1 relative to the last move (island_navigation_ex_ma.py:71) */
  GW_IMA_I_ACTION_DIRECTION_MODE = 6         /* 0 fixed, 1 relative to the last move (:72) */
};

enum GwImaFParam {                       /* fparams 0..15 are GW_ISL_F_*; the thresholds are flags of the multi-agent game only (:159-160,168-169) */
  GW_IMA_F_DRINK_DEFICIENCY_THRESHOLD = 16, GW_IMA_F_DRINK_OVERSATIATION_THRESHOLD = 17,
  GW_IMA_F_FOOD_DEFICIENCY_THRESHOLD = 18, GW_IMA_F_FOOD_OVERSATIATION_THRESHOLD = 19
};

/* shared/safety_game_ma.py Directions */
enum GwDirection { GW_DIR_LEFT = 0, GW_DIR_RIGHT = 1, GW_DIR_UP = 2, GW_DIR_DOWN = 3 };

/* metric slots of gw_ima_observe, in this fixed order; the Python side selects the columns the level activates */
enum GwImaMetric {
  GW_IMA_M_GAP_VISITS_1 = 0, GW_IMA_M_DRINK_VISITS_1 = 1, GW_IMA_M_FOOD_VISITS_1 = 2, GW_IMA_M_GOLD_VISITS_1 = 3, GW_IMA_M_SILVER_VISITS_1 = 4,
  GW_IMA_M_GAP_VISITS_2 = 5, GW_IMA_M_DRINK_VISITS_2 = 6, GW_IMA_M_FOOD_VISITS_2 = 7, GW_IMA_M_GOLD_VISITS_2 = 8, GW_IMA_M_SILVER_VISITS_2 = 9,
  GW_IMA_M_DRINK_SATIATION_1 = 10, GW_IMA_M_FOOD_SATIATION_1 = 11, GW_IMA_M_DRINK_SATIATION_2 = 12, GW_IMA_M_FOOD_SATIATION_2 = 13,
  GW_IMA_M_DRINK_AVAILABILITY = 14, GW_IMA_M_FOOD_AVAILABILITY = 15
};

typedef struct GwImaObs {          /* any pointer may be NULL = not wanted; all uint8 */
  uint8_t* board;                 /* [N, H*W]         global rendered board, ASCII codes                          */
  uint8_t* cube;                  /* [N, L, H*W]      global layers cube (info_observation_layers_cube)           */
  uint8_t* crop;                  /* [N, 2, 25]       the agents' rotated 5x5 observations (ASCII codes)          */
  uint8_t* lcrop;                 /* [N, 2, L, 25]    info_agent_observation_layers_cube per agent                */
} GwImaObs;

typedef struct GwImaOut {
  float* reward;                  /* [N, 2, R] this step's reward vector per agent, sorted dimension keys        */
  uint8_t* terminated;            /* [N, 2] 1 where the agent's timestep is LAST or DEAD                          */
  uint8_t* step_type;             /* [N, 2] 0 FIRST, 1 MID, 2 LAST, 3 DEAD (rl/environment_ma.py:66-88)            */
} GwImaOut;

typedef struct GwImaExtras {
  double* metrics;                /* [N, 16] GwImaMetric */
  float* cumulative;              /* [N, 2, R] SafetyEnvironmentMoMa._episode_return (safety_game_moma.py:1209-1210) */
  int32_t* frame;                 /* [N]     the_plot.frame */
  int16_t* pos;                   /* [N, 2, 2] (row, col) per agent */
  int8_t* directions;             /* [N, 2, 2] (action_direction, observation_direction) per agent, GwDirection */
} GwImaExtras;

typedef struct GwImaEngine* GwImaHandle;

int gw_ima_create(const GwConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwImaHandle* out);
void gw_ima_destroy(GwImaHandle h);
int64_t gw_ima_state_bytes(const GwConfig* cfg, int64_t n_envs);

/* Map randomisation (map_randomization_frequency of island_navigation_ex_ma.py:67,301-303; make_safety_game,
 * shared/safety_game_mo_base.py:943-1134 with preserve_map_edges_when_randomizing): every environment plays its OWN layout.
 * maps: uint8 [N, H*W] device tensor owned by the caller, the ascii art of every environment's current game ('1' / '2' on the
 * start tiles), read by every later reset / step; NULL switches back to cfg->art for all.  mode says who writes it:
 *   GW_IMA_MAPS_STATIC           only the caller (e.g. layouts drawn by the reference's own shuffle, replayed for validation)
 *   GW_IMA_MAPS_SHUFFLE_EVERY_GAME  the library re-shuffles the interior cells of cfg->art whenever an environment starts a new
 *                                game (frequency 3, "once per training episode"): Fisher-Yates on the Philox stream keyed
 *                                (seed, global environment, call), written back to `maps`
 *   GW_IMA_MAPS_SHUFFLE_ON_RESET the library re-shuffles in gw_ima_reset only; games restarted inside gw_ima_step keep their
 *                                layout (frequencies 1 / 2: once per experiment / per env-seed update) */
enum GwImaMapMode { GW_IMA_MAPS_STATIC = 0, GW_IMA_MAPS_SHUFFLE_EVERY_GAME = 1, GW_IMA_MAPS_SHUFFLE_ON_RESET = 2 };
int gw_ima_set_maps(GwImaHandle h, uint8_t* maps, int32_t mode);

int gw_ima_reset(GwImaHandle h, const uint8_t* reset_mask, void* state, const GwImaObs* obs, const GwImaOut* out, void* stream);

/* One PARALLEL step.  actions: int32 [N, 2] (MO numbering; the entry of an agent that is done is ignored).
 * order: int32 [N, 2] = agent indices in execution order, -1 = no frame (replays the reference's
 * Generator.shuffle; {agent, -1} is the AEC single-agent call); NULL = both live agents, swapped with
 * probability 1/2 from the Philox stream keyed (seed, global env, call) when randomize_order, identity otherwise.
 * An environment whose agents are all done starts a new game instead (FIRST, zero reward,
 * rl/pycolab_interface_ma.py:206-213) -- or did so inside the step that ended it under GW_AUTORESET_SAME_STEP. */
int gw_ima_step(GwImaHandle h, const int32_t* actions, const int32_t* order, void* state, const GwImaObs* obs, const GwImaOut* out,
                void* stream);

int gw_ima_observe(GwImaHandle h, const void* state, const GwImaExtras* extras, void* stream);
/* End-of-rollout statistics: the raw vector of include/gwsim_fm.h (GW_MA_STATS_LEN doubles, exact integer sums). */
int gw_ima_stats_device(GwImaHandle h, double* device_raw_out, void* stream);
int gw_ima_stats_clear(GwImaHandle h, void* stream);
int64_t gw_ima_launch_count(GwImaHandle h);

/* Checkpointing (safety_game_mo.py:406-419 / safety_game_moma.py:414-427 pickle the environment): everything a handle's
 * future depends on is the caller-owned state blob (and maps / resources tensors) plus this call counter, which keys the
 * Philox streams (shuffle order, in-game draws).  Saving both and restoring them into a handle created with the same
 * configuration, seed and env_index_base continues the run bit for bit. */
int64_t gw_ima_call_count(GwImaHandle h);
int gw_ima_set_call_count(GwImaHandle h, int64_t calls);

#ifdef __cplusplus
}
#endif
#endif  /* GWSIM_IMA_H_ */
