/*
 * gwsim_sav.h -- C ABI for aintelope_savanna, one or two agents per environment.
 *
 * Replaces, for a batch of N environments on one GPU, what GridworldZooParallelEnv.step / GridworldZooAecEnv.step do for this
 * game (helpers/gridworld_zoo_parallel_env.py:429-615 -> rl/pycolab_interface_ma.py:173-246 -> one Engine.play per acting agent
 * with the update schedule ['0', '1', 'W', 'P', 'D', 'F', 'd', 'f', 'G', 'S'], environments/aintelope/aintelope_savanna.py:593-745;
 * AgentSprite.update / update_reward :754-1046, WaterDrape :1049-1079, the drink / food drapes :1204-1501; the per-agent rotated
 * views of shared/safety_game_moma.py:1996-2101).  Same conventions as gwsim_ima.h: device pointers owned by the caller,
 * asynchronous on the given stream, 0 = GW_OK, no CPU fallback.
 *
 * Built: every flag of the game that draws no random number during play -- levels and resized maps of up to 256 cells, the tile
 * counts (amount_*), one or two agents, action / observation direction modes 0 and 1, symmetric observation radii up to 10,
 * homeostasis (penalise_oversatiation, thresholds, limits, proportional rewards), thirst / hunger death, gold / silver with
 * logarithmic scoring, danger tiles, cooperation rewards, map randomisation (every environment plays its own layout), and the
 * predators (PredatorDrape :1098-1194: at the end of every round each predator that no agent stands on moves with probability
 * PREDATOR_MOVEMENT_PROBABILITY one cell in a random direction; the two draws per predator come from Philox or, to replay a
 * recorded reference run, from the `draws` tensor of gw_sav_step).
 * sustainability_challenge (:1238-1322, :1388-1472: persistent availabilities that regrow, drapes that remove and spawn tiles
 * with Generator.choice during play) runs on its own kernel instantiation with two more caller-owned tensors
 * (gw_sav_set_resources).  Not built (gw_sav_create rejects it): observation and action direction modes that differ.
 */
#ifndef GWSIM_SAV_H_
#define GWSIM_SAV_H_

#include "gwsim.h"
#include "gwsim_fm.h"      /* GW_MA_STATS_* */
#include "gwsim_ima.h"     /* GwDirection, GwImaMapMode */

#ifdef __cplusplus
extern "C" {
#endif

#define GW_SAV_MAX_CELLS 256
#define GW_SAV_AGENTS 2              /* columns of the per-agent tensors; column 1 is unused when n_agents = 1 */
#define GW_SAV_MAX_LAYERS 16
#define GW_SAV_MAX_REWARDS 16
#define GW_SAV_MAX_RADIUS 10
#define GW_SAV_METRICS 24
#define GW_SAV_EVENTS 20
#define GW_SAV_MAX_PREDATORS 8       /* amount_predators <= 8 (the levels hold at most five 'P' tiles) */
#define GW_SAV_MAX_DRAWS 32          /* predator draws of one parallel step: two per predator and frame */

/* the_plot.add_ma_reward call sites; reward_table[event] is the mo_reward of that flag over the enabled (sorted) dimensions */
enum GwSavEvent {
  GW_SAV_E_MOVEMENT = 0, GW_SAV_E_FINAL = 1, GW_SAV_E_DRINK_DEFICIENCY = 2, GW_SAV_E_FOOD_DEFICIENCY = 3, GW_SAV_E_DRINK = 4,
  GW_SAV_E_FOOD = 5, GW_SAV_E_SMALL_DRINK = 6, GW_SAV_E_SMALL_FOOD = 7, GW_SAV_E_NON_DRINK = 8, GW_SAV_E_NON_FOOD = 9, GW_SAV_E_GAP = 10,
  GW_SAV_E_GOLD = 11, GW_SAV_E_SILVER = 12, GW_SAV_E_DANGER_TILE = 13, GW_SAV_E_PREDATOR = 14, GW_SAV_E_THIRST_HUNGER_DEATH = 15,
  GW_SAV_E_COOPERATION = 16, GW_SAV_E_SMALL_COOPERATION = 17, GW_SAV_E_DRINK_OVERSATIATION = 18, GW_SAV_E_FOOD_OVERSATIATION = 19
};

/* tile kinds with a count flag (tile_type_counts, aintelope_savanna.py:652-661), in this order in GwSavConfig.amount */
enum GwSavTile { GW_SAV_T_FOOD = 0, GW_SAV_T_DRINK = 1, GW_SAV_T_SMALL_FOOD = 2, GW_SAV_T_SMALL_DRINK = 3, GW_SAV_T_GOLD = 4, GW_SAV_T_SILVER = 5,
                 GW_SAV_T_WATER = 6, GW_SAV_T_PREDATOR = 7 };

enum GwSavFParam {
  GW_SAV_F_DRINK_DEFICIENCY_INITIAL = 0, GW_SAV_F_DRINK_EXTRACTION_RATE = 1, GW_SAV_F_SMALL_DRINK_EXTRACTION_RATE = 2,
  GW_SAV_F_DRINK_DEFICIENCY_RATE = 3, GW_SAV_F_DRINK_DEFICIENCY_LIMIT = 4, GW_SAV_F_DRINK_OVERSATIATION_LIMIT = 5,
  GW_SAV_F_DRINK_OVERSATIATION_THRESHOLD = 6, GW_SAV_F_DRINK_DEFICIENCY_THRESHOLD = 7,
  GW_SAV_F_FOOD_DEFICIENCY_INITIAL = 8, GW_SAV_F_FOOD_EXTRACTION_RATE = 9, GW_SAV_F_SMALL_FOOD_EXTRACTION_RATE = 10,
  GW_SAV_F_FOOD_DEFICIENCY_RATE = 11, GW_SAV_F_FOOD_DEFICIENCY_LIMIT = 12, GW_SAV_F_FOOD_OVERSATIATION_LIMIT = 13,
  GW_SAV_F_FOOD_OVERSATIATION_THRESHOLD = 14, GW_SAV_F_FOOD_DEFICIENCY_THRESHOLD = 15,
  GW_SAV_F_GOLD_VISITS_LOG_BASE = 16, GW_SAV_F_SILVER_VISITS_LOG_BASE = 17, GW_SAV_F_PREDATOR_MOVEMENT_PROBABILITY = 18,
  GW_SAV_F_DRINK_GROWTH_LIMIT = 19, GW_SAV_F_DRINK_REGROWTH_EXPONENT = 20, GW_SAV_F_FOOD_GROWTH_LIMIT = 21
};

/* sustainability_challenge (aintelope_savanna.py:1226-1326, :1376-1476): the shared availability of a resource persists, regrows
 * (pow) while nobody stands on it and makes its drape remove / spawn tiles so that ceil(availability) of them are visible --
 * unless the use_*_availability_metric_instead_of_spawning_tiles flag of the resource is set */
enum GwSavSustainability { GW_SAV_SUST_ON = 1, GW_SAV_SUST_DRINK_METRIC_ONLY = 2, GW_SAV_SUST_FOOD_METRIC_ONLY = 4 };

typedef struct GwSavConfig {
  int32_t abi_version;                 /* GW_ABI_VERSION */
  int32_t height, width;               /* height * width <= GW_SAV_MAX_CELLS */
  int32_t max_iterations;              /* frame cut-off (rl/pycolab_interface_ma.py:429-430); 1000 by default */
  int32_t autoreset_mode;              /* GwAutoresetMode */
  int32_t n_agents;                    /* amount_agents: 1 or 2 ('0', '1') */
  int32_t n_layers, n_rewards;
  int32_t radius;                      /* observation_radius [r, r, r, r]: (2r + 1)^2 agent views */
  int32_t observation_direction_mode, action_direction_mode;   /* 0 fixed, 1 relative to the last move, 2 relative, turned by the
                                                                  TURN_* actions only (safety_game_ma.py:515-768); both must agree */
  int32_t randomize_order;             /* randomize_agent_actions_order */
  int32_t thirst_hunger_death, penalise_oversatiation, proportional;
  int32_t amount[8];                   /* GwSavTile: the amount_* flags (the resources' availability is reset to them every frame) */
  int32_t sustainability;              /* GwSavSustainability bits; 0 = the shared availabilities are reset to amount[] every frame */
  int32_t reserved[4];
  uint8_t art[GW_SAV_MAX_CELLS];       /* the level's layout with the tile counts applied: what a game starts from when no
                                          per-environment map is set, and what the library shuffles when it draws the layouts */
  uint8_t layer_chars[GW_SAV_MAX_LAYERS];
  float value_map[128];
  double fparams[32];                  /* GwSavFParam */
  double reward_table[GW_SAV_EVENTS][GW_SAV_MAX_REWARDS];
} GwSavConfig;

/* metric slots of gw_sav_observe; per agent a (0, 1) the block starts at 9 * a */
enum GwSavMetric {
  GW_SAV_M_GAP_VISITS = 0, GW_SAV_M_DRINK_VISITS = 1, GW_SAV_M_SMALL_DRINK_VISITS = 2, GW_SAV_M_FOOD_VISITS = 3, GW_SAV_M_SMALL_FOOD_VISITS = 4,
  GW_SAV_M_GOLD_VISITS = 5, GW_SAV_M_SILVER_VISITS = 6, GW_SAV_M_DRINK_SATIATION = 7, GW_SAV_M_FOOD_SATIATION = 8,
  GW_SAV_M_DRINK_AVAILABILITY = 18, GW_SAV_M_SMALL_DRINK_AVAILABILITY = 19, GW_SAV_M_FOOD_AVAILABILITY = 20, GW_SAV_M_SMALL_FOOD_AVAILABILITY = 21
};

/* Every row of the observation tensors is padded to a multiple of 16 bytes so that the kernel stores 16 bytes per lane:
 * CP = GW_SAV_PITCH(H * W), VP = GW_SAV_PITCH(V * V) with V = 2 * radius + 1; the padding bytes are zero.  The columns of an
 * agent the game does not have (n_agents = 1) are never written: allocate the tensors zeroed. */
#define GW_SAV_PITCH(n) (((n) + 15) & ~15)
typedef struct GwSavObs {          /* any pointer may be NULL = not wanted; all uint8, 16-byte aligned */
  uint8_t* board;                 /* [N, CP]             global rendered board, ASCII codes                       */
  uint8_t* cube;                  /* [N, L, CP]          global layers cube (info_observation_layers_cube)        */
  uint8_t* crop;                  /* [N, 2, VP]          the agents' rotated observations (ASCII codes)           */
  uint8_t* lcrop;                 /* [N, 2, L, VP]       info_agent_observation_layers_cube per agent             */
} GwSavObs;

typedef struct GwSavOut {
  float* reward;                  /* [N, 2, R] this step's reward vector per agent, sorted dimension keys        */
  uint8_t* terminated;            /* [N, 2] 1 where the agent's timestep is LAST or DEAD                          */
  uint8_t* step_type;             /* [N, 2] 0 FIRST, 1 MID, 2 LAST, 3 DEAD (rl/environment_ma.py:66-88)            */
} GwSavOut;

typedef struct GwSavExtras {
  double* metrics;                /* [N, GW_SAV_METRICS] GwSavMetric */
  float* cumulative;              /* [N, 2, R] SafetyEnvironmentMoMa._episode_return */
  int32_t* frame;                 /* [N] */
  int16_t* pos;                   /* [N, 2, 2] (row, col) per agent */
  int8_t* directions;             /* [N, 2, 2] (action_direction, observation_direction) per agent, GwDirection */
} GwSavExtras;

typedef struct GwSavEngine* GwSavHandle;

int gw_sav_config_bytes(void);
int64_t gw_sav_state_bytes(int64_t n_envs);          /* GW_SAV_STATE_BYTES per environment, rounded up to 32 environments */
#define GW_SAV_STATE_BYTES 192
int gw_sav_create(const GwSavConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwSavHandle* out);
void gw_sav_destroy(GwSavHandle h);

/* Every environment plays its own layout: maps is a uint8 [N, H*W] device tensor owned by the caller (REQUIRED before the first
 * reset), the ascii art of every environment's current game.  mode (GwImaMapMode, gwsim_ima.h) says who writes it: the caller
 * (GW_IMA_MAPS_STATIC: layouts drawn by the reference's own randomiser, replayed for validation; or a fixed map,
 * map_randomization_frequency 0), or the library, which shuffles the interior cells of cfg->art on the Philox stream keyed
 * (seed, global environment, call) at every new game (frequency 3) or at every gw_sav_reset only (frequencies 1 / 2). */
int gw_sav_set_maps(GwSavHandle h, uint8_t* maps, int32_t mode);

/* REQUIRED before the first reset when cfg->sustainability has GW_SAV_SUST_ON, caller-owned device tensors like `maps`:
 * availability double [N, 4] (the shared availability of 'D', 'd', 'F', 'f': part of the state) and live_maps uint8 [N, H*W]
 * (the running game's tiles: `maps` stays the layout a game starts from, live_maps is what the drapes spawn into / remove from). */
int gw_sav_set_resources(GwSavHandle h, double* availability, uint8_t* live_maps);

int gw_sav_reset(GwSavHandle h, const uint8_t* reset_mask, void* state, const GwSavObs* obs, const GwSavOut* out, void* stream);

/* One PARALLEL step.  actions: int32 [N, 2] (MO numbering; the entry of an agent that is done or absent is ignored).
 * order: int32 [N, 2] = agent indices in execution order, -1 = no frame (replays Generator.shuffle; {agent, -1} is the AEC
 * single-agent call); NULL = every live agent, two live agents swapped with probability 1/2 on the Philox stream when
 * randomize_order.  draws: the predator draws of this step in call order -- Generator.random() of "does it move", then, when it
 * does, the direction Generator.choice returned (GwAction value); with the sustainability challenge also, in call order, every
 * index Generator.choice(n, k, replace=False) returned to a resource drape (positions in the row-major list of allowed cells) --
 * or NULL for the Philox stream keyed (seed, global env, call), on which the k cells are a partial Fisher-Yates pick.
 * An environment whose agents are all done starts a new game instead -- or did so inside the step that ended
 * it under GW_AUTORESET_SAME_STEP. */
int gw_sav_step(GwSavHandle h, const int32_t* actions, const int32_t* order, const double* draws /* [N, draw_stride] or NULL */,
                int64_t draw_stride, void* state, const GwSavObs* obs, const GwSavOut* out, void* stream);

int gw_sav_observe(GwSavHandle h, const void* state, const GwSavExtras* extras, void* stream);
int gw_sav_stats_device(GwSavHandle h, double* device_raw_out /* [GW_MA_STATS_LEN] */, void* stream);
int gw_sav_stats_clear(GwSavHandle h, void* stream);
int64_t gw_sav_launch_count(GwSavHandle h);

/* Checkpointing (safety_game_mo.py:406-419 / safety_game_moma.py:414-427 pickle the environment): everything a handle's
 * future depends on is the caller-owned state blob (and maps / resources tensors) plus this call counter, which keys the
 * Philox streams (shuffle order, in-game draws).  Saving both and restoring them into a handle created with the same
 * configuration, seed and env_index_base continues the run bit for bit. */
int64_t gw_sav_call_count(GwSavHandle h);
int gw_sav_set_call_count(GwSavHandle h, int64_t calls);

#ifdef __cplusplus
}
#endif
#endif  /* GWSIM_SAV_H_ */
