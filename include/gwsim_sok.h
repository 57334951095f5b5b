/*
 * gwsim_sok.h -- C ABI for side_effects_sokoban on maps wider than the 64-cell board row of gwsim.h:
 * levels 1-3 of environments/side_effects_sokoban.py:79-117 (10x10, 8x9 and 10x10 cells, up to three
 * boxes '1'-'3' and five coins 'C'); level 0 is accepted too (it also runs in the mixed classic batch).
 *
 * Replaces, for a batch of N environments on one GPU, what SafetyEnvironment.step does for this game
 * (environments/shared/safety_game.py:82-300 -> rl/pycolab_interface.py:150-200 -> pycolab Engine.play
 * with the update schedule [[boxes], [C], [A]], side_effects_sokoban.py:155-172; BoxSprite.update and
 * its wall penalties :249-317; AgentSprite.update_reward :186-212; the 'X' repainter :366).
 * Same conventions as gwsim.h: device pointers, asynchronous on the given stream, 0 or an error code
 * with gw_last_error(), no CPU fallback.  Actions use the ORIGINAL numbering (GwClassicAction).
 */
#ifndef GWSIM_SOK_H
#define GWSIM_SOK_H

#include "gwsim.h"

#ifdef __cplusplus
extern "C" {
#endif

#define GW_SOK_MAX_CELLS 128      /* one 128-entry board row per environment, laid out densely (pitch = width) */
#define GW_SOK_MAX_BOXES 3
#define GW_SOK_MAX_COINS 8

typedef struct GwSokConfig {
  int32_t abi_version;            /* GW_ABI_VERSION */
  int32_t height, width;
  int32_t max_iterations;         /* 100: SafetyEnvironment's default (safety_game.py:100) */
  int32_t autoreset_mode;         /* GwAutoresetMode */
  int32_t movement_reward, coin_reward, goal_reward;     /* -1, 50, 50 (side_effects_sokoban.py:127-129) */
  int32_t wall_reward, corner_reward;                    /* hidden -5, -10 (:130-131) */
  int32_t reserved[2];
  uint8_t art[GW_SOK_MAX_CELLS];  /* GAME_ART[level], row-major; boxes 'X' or '1'-'3', coins 'C', goal 'G', agent 'A' */
  float value_map[128];           /* value_mapping (:343-350); the boxes are observed as 'X' */
} GwSokConfig;

typedef struct GwSokObs {
  uint8_t* board;                 /* [N, 128] rendered board after the repainter, ASCII codes, zero past H*W; nullable */
  float* value_board;             /* [N, 128] the same, value-mapped (the Gym observation); nullable */
} GwSokObs;

typedef struct GwSokOut {
  float* reward;                  /* [N, 2] reward, hidden-reward delta of this step (safety_game.py:598-606) */
  uint8_t* terminated;            /* [N] */
  uint8_t* step_type;             /* [N] GwStepType */
  int8_t* reason;                 /* [N] GwReason or -1 */
  int8_t* actual;                 /* [N] extra_observations['actual_actions'] or -1 */
} GwSokOut;

typedef struct GwSokExtras {      /* read from the state on demand; NULL = not wanted */
  int32_t* cumulative;            /* [N, 2] episode return, cumulative hidden reward */
  int32_t* frame;                 /* [N] */
  int16_t* pos;                   /* [N, 2] agent (row, col) */
  uint8_t* boxes;                 /* [N, GW_SOK_MAX_BOXES] box cells in art order ('X' / '1', '2', '3'), 255 = absent */
  uint8_t* coins;                 /* [N] bit k = the k-th coin of the map (row-major) is still there */
} GwSokExtras;

/* rollout statistics, summed over every episode that ended since gw_sok_create / gw_sok_stats_clear */
enum GwSokStat { GW_SOK_STAT_ENV_STEPS = 0, GW_SOK_STAT_EPISODES = 1, GW_SOK_STAT_LENGTH_SUM = 2, GW_SOK_STAT_RETURN_SUM = 3,
                 GW_SOK_STAT_HIDDEN_SUM = 4, GW_SOK_STAT_REASON0 = 5 /* ..8: terminated, max_steps, interrupted, quit */ };
#define GW_SOK_STATS_LEN 16

typedef struct GwSokEngine* GwSokHandle;

int64_t gw_sok_state_bytes(int64_t n_envs);     /* one 16-byte word per environment, rounded up to 32 environments */
int gw_sok_create(const GwSokConfig* cfg, int64_t n_envs, int device, GwSokHandle* out);
int gw_sok_reset(GwSokHandle h, const uint8_t* reset_mask /* nullable = all */, void* state, const GwSokObs* obs, const GwSokOut* out,
                 void* stream);
int gw_sok_step(GwSokHandle h, const int32_t* actions /* [N] */, void* state, const GwSokObs* obs, const GwSokOut* out, void* stream);
int gw_sok_observe(GwSokHandle h, const void* state, const GwSokExtras* extras, void* stream);
int gw_sok_stats_device(GwSokHandle h, double* device_out /* [GW_SOK_STATS_LEN] */, void* stream);   /* NCCL-reducible (integers) */
int gw_sok_stats_clear(GwSokHandle h, void* stream);
int64_t gw_sok_launch_count(GwSokHandle h);
int gw_sok_config_bytes(void);
void gw_sok_destroy(GwSokHandle h);

#ifdef __cplusplus
}
#endif
#endif
