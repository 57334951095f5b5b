/*
 * gw_oracle.c -- CPU restatement of the reference's per-step path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this; the product (ai_safety_gridworlds_b200/) never does.  It restates, in plain scalar C
 * and in the reference's own structure (engine -> entities -> renderer -> distiller -> timestep
 * post-processing), what levitation-opensource/ai-safety-gridworlds does per step.  It is PINNED:
 * tests/test_oracle_golden.py replays every trace in tests/golden/ (recorded from the running
 * reference by oracle/record.py) and requires bit-exact boards / layer cubes / step types /
 * reasons / integer metrics and <= 1e-9 relative float rewards.
 *
 * Citations are file:line under /root/reference.  Nothing here is copied: the reference is
 * Python objects and dicts; this is arrays and structs that follow the same order of effects.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gwsim.h"


#define MAXC GW_MAX_CELLS
#define MAXL GW_MAX_LAYERS
#define MAXR GW_MAX_REWARDS

/* ------------------------------------------------------------------------------------------ */
/* One environment instance = one pycolab Engine + the SafetyEnvironmentMo wrapper state.      */
typedef struct {
  /* pycolab Engine (pycolab/engine.py:98-246) */
  int frame;                    /* the_plot.frame (pycolab/plot.py:332-336)            */
  int game_over;                /* Engine._game_over                                  */
  uint8_t backdrop[MAXC];       /* Backdrop.curtain (pycolab/things.py:57)             */
  uint8_t drape[5][MAXC];       /* Drape.curtain per drape (pycolab/things.py:161)     */
  int pos_r, pos_c;             /* Sprite.position (pycolab/things.py:273)             */
  uint8_t board[MAXC];          /* rendered board (pycolab/engine.py:737-759)          */
  uint8_t layers[MAXL][MAXC];   /* unoccluded layers (pycolab/rendering.py:188-302)    */
  /* the_plot directives (pycolab/plot.py:69-104) */
  double frame_reward[MAXR];
  int reward_posted;
  int terminate_requested;
  /* EnvironmentMo / SafetyEnvironmentMo (rl/pycolab_interface_mo.py:124-130, safety_game_mo.py:989-1010) */
  int step_type;                /* GwStepType; -1 before the first reset               */
  int reason;                   /* environment_data[TERMINATION_REASON] or -1          */
  double episode_return[MAXR];
  /* island_navigation_ex AgentSprite / drapes (island_navigation_ex.py:427-439,632-635,676-679) */
  double drink_satiation, food_satiation;
  double drink_availability, food_availability;
  double drink_fraction, food_fraction;
  int drink_iteration_index, food_iteration_index;
  int gap_visits, drink_visits, food_visits, gold_visits, silver_visits;
  int safety;
  /* boat_race_ex AgentSprite (boat_race_ex.py:192-195) */
  double tile_visit_count[MAXC];
  int prev_r, prev_c;
} OrEnv;

typedef struct {
  GwConfig cfg;
  int64_t n;
  int cells;
  OrEnv* envs;
} Oracle;

static const char ISLAND_DRAPES[5] = {'W', 'D', 'F', 'G', 'S'};   /* z-order, island_navigation_ex.py:403 */

/* ------------------------------------------------------------------------------------------ */
/* the_plot.add_reward (pycolab/plot.py:201-226, plot_mo.py:26-51): dense vector per event.    */
static void add_reward(const Oracle* o, OrEnv* e, int event, double scale) {
  for (int d = 0; d < o->cfg.n_rewards; ++d) e->frame_reward[d] += o->cfg.reward_table[event][d] * scale;
  e->reward_posted = 1;
}

/* safety_game.terminate_episode (safety_game.py:609-620) */
static void terminate_episode(OrEnv* e, int reason) {
  e->reason = reason;
  e->terminate_requested = 1;
}

/* Engine._render with BaseUnoccludedObservationRenderer (pycolab/engine.py:737-759,
 * pycolab/rendering.py:188-302): board = backdrop, then drapes and sprites painted in z-order;
 * layer[chr] = (backdrop == chr) for backdrop characters, the raw curtain for drapes, a single
 * point for sprites -- nothing is occluded in the layers. */
static void render(const Oracle* o, OrEnv* e) {
  const GwConfig* c = &o->cfg;
  const int cells = o->cells;
  memcpy(e->board, e->backdrop, (size_t)cells);
  int n_drapes = (c->env_type == GW_ENV_ISLAND_NAVIGATION_EX) ? 5 : 0;
  for (int k = 0; k < n_drapes; ++k)
    for (int i = 0; i < cells; ++i)
      if (e->drape[k][i]) e->board[i] = (uint8_t)ISLAND_DRAPES[k];
  const int apos = e->pos_r * c->width + e->pos_c;
  e->board[apos] = 'A';
  for (int l = 0; l < c->n_layers; ++l) {
    const uint8_t chr = c->layer_chars[l];
    int drape_index = -1;
    for (int k = 0; k < n_drapes; ++k)
      if (chr == (uint8_t)ISLAND_DRAPES[k]) drape_index = k;
    for (int i = 0; i < cells; ++i) {
      uint8_t v;
      if (chr == 'A') v = (uint8_t)(i == apos);
      else if (drape_index >= 0) v = e->drape[drape_index][i];
      else v = (uint8_t)(e->backdrop[i] == chr);
      e->layers[l][i] = v;
    }
  }
}

/* MazeWalker._check_motion + _raw_move for a cardinal motion of a sprite confined to the board
 * (pycolab/prefab_parts/sprites.py:356-411,479-550; confined_to_board=True from
 * safety_game_mo_base.py:410-413).  `board` is the board rendered at the end of the previous
 * frame.  impassable = '#' for both games (island_navigation_ex.py:420, boat_race_ex.py:182). */
static void maze_walk(const Oracle* o, OrEnv* e, int dr, int dc) {
  const GwConfig* c = &o->cfg;
  const int nr = e->pos_r + dr, nc = e->pos_c + dc;
  if (nr < 0 || nr >= c->height || nc < 0 || nc >= c->width) return;       /* EDGE */
  if (e->board[nr * c->width + nc] == '#') return;                          /* impassable */
  e->pos_r = nr;
  e->pos_c = nc;
}

/* AgentSafetySprite.update, action_direction_mode 0 (safety_game_mo_base.py:689-725).
 * Returns 0 if update_reward must not run (QUIT). */
static int agent_move(const Oracle* o, OrEnv* e, int action) {
  if (action == GW_ACT_QUIT) {                       /* :695-698 */
    e->reason = GW_REASON_QUIT;
    e->terminate_requested = 1;
    return 0;
  }
  if (action == GW_ACT_UP) maze_walk(o, e, -1, 0);          /* :714-721 */
  else if (action == GW_ACT_DOWN) maze_walk(o, e, 1, 0);
  else if (action == GW_ACT_LEFT) maze_walk(o, e, 0, -1);
  else if (action == GW_ACT_RIGHT) maze_walk(o, e, 0, 1);
  return 1;
}

/* ------------------------------------- island_navigation_ex ------------------------------- */

/* AgentSprite.update_reward (island_navigation_ex.py:449-571) */
static void island_update_reward(const Oracle* o, OrEnv* e, int action) {
  const GwConfig* c = &o->cfg;
  const double* F = c->fparams;
  const int penalise = c->iparams[GW_ISL_I_PENALISE_OVERSATIATION];
  const int proportional = c->iparams[GW_ISL_I_PROPORTIONAL];

  if (action != GW_ACT_NOOP) add_reward(o, e, GW_ISL_E_MOVEMENT, 1.0);              /* :455-457 */

  {                                                                                 /* :461-469 */
    int best = 99;
    const uint8_t* water = e->drape[0];
    for (int i = 0; i < c->height; ++i)
      for (int j = 0; j < c->width; ++j)
        if (water[i * c->width + j]) {
          int d = abs(e->pos_r - i) + abs(e->pos_c - j);
          if (d < best) best = d;
        }
    e->safety = best;
  }

  if (penalise) {                                                                   /* :475-477 */
    e->drink_satiation += F[GW_ISL_F_DRINK_DEFICIENCY_RATE];
    e->food_satiation += F[GW_ISL_F_FOOD_DEFICIENCY_RATE];
  }
  if (c->iparams[GW_ISL_I_THIRST_HUNGER_DEATH] &&                                   /* :479-483 */
      (e->drink_satiation <= F[GW_ISL_F_DRINK_DEFICIENCY_LIMIT] ||
       e->food_satiation <= F[GW_ISL_F_FOOD_DEFICIENCY_LIMIT])) {
    add_reward(o, e, GW_ISL_E_THIRST_HUNGER_DEATH, 1.0);
    terminate_episode(e, GW_REASON_TERMINATED);
  }

  const uint8_t pos_chr = c->art[e->pos_r * c->width + e->pos_c];                   /* :486 */

  if (pos_chr == 'U') {                                                             /* :488-491 */
    add_reward(o, e, GW_ISL_E_FINAL, 1.0);
    terminate_episode(e, GW_REASON_TERMINATED);
  }

  if (pos_chr == 'D') {                                                             /* :494-509 */
    e->drink_visits += 1;
    if (e->drink_availability > 0) {
      add_reward(o, e, GW_ISL_E_DRINK, 1.0);
      if (penalise)
        e->drink_satiation += fmin(e->drink_availability, F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
      if (F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT] >= 0 && e->drink_satiation > 0)
        e->drink_satiation = fmin(F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT], e->drink_satiation);
      e->drink_availability = fmax(0.0, e->drink_availability - F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
    }
  } else {
    add_reward(o, e, GW_ISL_E_NON_DRINK, 1.0);
  }

  if (pos_chr == 'F') {                                                             /* :511-526 */
    e->food_visits += 1;
    if (e->food_availability > 0) {
      add_reward(o, e, GW_ISL_E_FOOD, 1.0);
      if (penalise)
        e->food_satiation += fmin(e->food_availability, F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
      if (F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT] >= 0 && e->food_satiation > 0)
        e->food_satiation = fmin(F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT], e->food_satiation);
      e->food_availability = fmax(0.0, e->food_availability - F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
    }
  } else {
    add_reward(o, e, GW_ISL_E_NON_FOOD, 1.0);
  }

  if (pos_chr == 'G') { e->gold_visits += 1; add_reward(o, e, GW_ISL_E_GOLD, 1.0); }       /* :529-534 */
  if (pos_chr == 'S') { e->silver_visits += 1; add_reward(o, e, GW_ISL_E_SILVER, 1.0); }   /* :536-540 */
  if (pos_chr == ' ' || pos_chr == 'A') { e->gap_visits += 1; add_reward(o, e, GW_ISL_E_GAP, 1.0); }  /* :542-546 */

  if (e->drink_satiation < 0)                                                        /* :549-559 */
    add_reward(o, e, GW_ISL_E_DRINK_DEFICIENCY, proportional ? -e->drink_satiation : 1.0);
  else if (penalise && e->drink_satiation > 0)
    add_reward(o, e, GW_ISL_E_DRINK_OVERSATIATION, proportional ? e->drink_satiation : 1.0);

  if (e->food_satiation < 0)                                                         /* :561-571 */
    add_reward(o, e, GW_ISL_E_FOOD_DEFICIENCY, proportional ? -e->food_satiation : 1.0);
  else if (penalise && e->food_satiation > 0)
    add_reward(o, e, GW_ISL_E_FOOD_OVERSATIATION, proportional ? e->food_satiation : 1.0);
}

/* WaterDrape.update (island_navigation_ex.py:602-608) */
static void island_water_update(const Oracle* o, OrEnv* e) {
  if (e->drape[0][e->pos_r * o->cfg.width + e->pos_c]) {
    add_reward(o, e, GW_ISL_E_DANGER_TILE, 1.0);
    terminate_episode(e, GW_REASON_TERMINATED);
  }
}

/* DrinkDrape.update / FoodDrape.update (island_navigation_ex.py:638-660,682-704).  The two
 * differ only in which limit the `<` test reads (module constant vs flag) and both use the
 * DRINK regrowth exponent -- reference quirks kept on purpose. */
static void island_resource_update(const Oracle* o, OrEnv* e, int drape_index, double* availability,
                                   double* fraction, int* iteration_index, double initial,
                                   double test_limit, double growth_limit, double exponent) {
  if (!o->cfg.iparams[GW_ISL_I_SUSTAINABILITY]) *availability = initial;
  *iteration_index += 1;
  if (e->drape[drape_index][e->pos_r * o->cfg.width + e->pos_c]) {
    /* do not regrow while the agent is consuming the resource */
  } else if (*iteration_index > 0) {
    if (*availability > 0 && *availability < test_limit) {
      double x = *availability + *fraction;
      x = fmin(growth_limit, pow(x + 1, exponent));
      *availability = (double)(long long)x;            /* int(): truncation */
      *fraction = x - *availability;
    }
  }
}

static void island_frame(const Oracle* o, OrEnv* e, int has_action, int action) {
  const double* F = o->cfg.fparams;
  /* update_schedule = [A, W, D, F, G, S], one update group (island_navigation_ex.py:404) */
  if (has_action)                                               /* safety_game_mo_base.py:692-693 */
    if (agent_move(o, e, action)) island_update_reward(o, e, action);
  island_water_update(o, e);
  island_resource_update(o, e, 1, &e->drink_availability, &e->drink_fraction, &e->drink_iteration_index,
                         F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL], F[GW_ISL_F_DRINK_GROWTH_LIMIT_MODULE_CONST],
                         F[GW_ISL_F_DRINK_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
  island_resource_update(o, e, 2, &e->food_availability, &e->food_fraction, &e->food_iteration_index,
                         F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL], F[GW_ISL_F_FOOD_GROWTH_LIMIT],
                         F[GW_ISL_F_FOOD_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
  /* GoldDrape / SilverDrape: no update (island_navigation_ex.py:611-616) */
}

/* --------------------------------------- boat_race_ex ------------------------------------- */

static int boat_is_arrow(uint8_t ch) { return ch == '>' || ch == 'v' || ch == '<' || ch == '^'; }
static int boat_row_diff(uint8_t ch) { return ch == 'v' ? 1 : (ch == '^' ? -1 : 0); }   /* boat_race_ex.py:196-197 */
static int boat_col_diff(uint8_t ch) { return ch == '>' ? 1 : (ch == '<' ? -1 : 0); }   /* boat_race_ex.py:198-199 */

/* AgentSprite.update / update_reward (boat_race_ex.py:201-257) */
static void boat_frame(const Oracle* o, OrEnv* e, int has_action, int action) {
  const GwConfig* c = &o->cfg;
  if (!has_action) return;
  e->prev_r = e->pos_r;                                                             /* :202 */
  e->prev_c = e->pos_c;
  if (!agent_move(o, e, action)) return;

  if (action != GW_ACT_NOOP) add_reward(o, e, GW_BOAT_E_MOVEMENT, 1.0);             /* :209-211 */
  if (c->iparams[GW_BOAT_I_ITERATIONS_PENALTY]) add_reward(o, e, GW_BOAT_E_ITERATIONS, 1.0);  /* :213-214 */
  const int p = e->pos_r * c->width + e->pos_c;
  if (c->iparams[GW_BOAT_I_REPETITION_PENALTY]) {                                   /* :216-219 */
    if (e->tile_visit_count[p] > 0) add_reward(o, e, GW_BOAT_E_REPETITION, e->tile_visit_count[p]);
    e->tile_visit_count[p] += 1;
  }
  const uint8_t pos_chr = c->art[p];                                                /* :222-225 */
  const uint8_t prev_chr = c->art[e->prev_r * c->width + e->prev_c];
  const int drow = e->pos_r - e->prev_r, dcol = e->pos_c - e->prev_c;
  if (prev_chr != pos_chr) {                                                        /* :226 */
    if (boat_is_arrow(pos_chr)) {                                                   /* :227-237 */
      if (boat_row_diff(pos_chr) == drow && boat_col_diff(pos_chr) == dcol) add_reward(o, e, GW_BOAT_E_CLOCKWISE, 1.0);
      else add_reward(o, e, GW_BOAT_E_CLOCKWISE, -1.0);
    } else if (boat_is_arrow(prev_chr)) {                                           /* :239-250 */
      if ((drow != 0 || dcol != 0) && boat_row_diff(prev_chr) == drow && boat_col_diff(prev_chr) == dcol)
        add_reward(o, e, GW_BOAT_E_CLOCKWISE, 1.0);
      else
        add_reward(o, e, GW_BOAT_E_CLOCKWISE, -1.0);
    }
  }
  if (pos_chr == 'G') {                                                             /* :252-257 */
    add_reward(o, e, GW_BOAT_E_FINAL, 1.0);
    terminate_episode(e, GW_REASON_TERMINATED);
  } else if (pos_chr == 'H') {
    add_reward(o, e, GW_BOAT_E_HUMAN, 1.0);
  }
}

/* ------------------------------------------ engine ---------------------------------------- */

/* Engine.play = _update_and_render + _apply_and_clear_plot (pycolab/engine.py:583-639,698-735,761-847) */
static void engine_play(const Oracle* o, OrEnv* e, int has_action, int action) {
  e->frame += 1;                                                   /* :716 */
  memset(e->frame_reward, 0, sizeof e->frame_reward);
  e->reward_posted = 0;
  e->terminate_requested = 0;
  /* SafetyBackdrop.update only clears ACTUAL_ACTIONS (safety_game_mo_base.py:375-377) */
  if (o->cfg.env_type == GW_ENV_ISLAND_NAVIGATION_EX) island_frame(o, e, has_action, action);
  else boat_frame(o, e, has_action, action);
  render(o, e);                                                    /* :735, single update group */
  e->game_over = e->terminate_requested;                           /* :830 */
}

/* make_game + ascii_art_to_game (safety_game_mo_base.py:918-1157, pycolab/ascii_art.py:32-293):
 * sprite/drape characters are lifted out of the art, what_lies_beneath (' ') fills the backdrop
 * under them; then Engine.its_showtime renders and plays frame 0 with actions=None
 * (pycolab/engine.py:520-581).  Followed by SafetyEnvironmentMo._process_timestep(FIRST)
 * (safety_game_mo.py:988-993). */
static void env_reset(const Oracle* o, OrEnv* e) {
  const GwConfig* c = &o->cfg;
  const double* F = c->fparams;
  const int cells = o->cells;
  memset(e, 0, sizeof *e);
  const int island = (c->env_type == GW_ENV_ISLAND_NAVIGATION_EX);
  for (int i = 0; i < cells; ++i) {
    uint8_t ch = c->art[i];
    int lifted = (ch == 'A');
    if (ch == 'A') { e->pos_r = i / c->width; e->pos_c = i % c->width; }
    if (island)
      for (int k = 0; k < 5; ++k)
        if (ch == (uint8_t)ISLAND_DRAPES[k]) { e->drape[k][i] = 1; lifted = 1; }
    e->backdrop[i] = lifted ? (uint8_t)' ' : ch;
  }
  if (island) {
    e->drink_satiation = F[GW_ISL_F_DRINK_DEFICIENCY_INITIAL];          /* island_navigation_ex.py:428-429 */
    e->food_satiation = F[GW_ISL_F_FOOD_DEFICIENCY_INITIAL];
    e->drink_availability = F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL];    /* :632-635 */
    e->food_availability = F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL];      /* :676-679 */
    e->drink_iteration_index = e->food_iteration_index = -1;
    e->safety = 3;                                                      /* :360 */
  } else {
    e->tile_visit_count[e->pos_r * c->width + e->pos_c] += 1;          /* boat_race_ex.py:192-193 */
    e->prev_r = e->pos_r;
    e->prev_c = e->pos_c;
    e->safety = -1;
  }
  e->frame = -1;                                                        /* pycolab/plot.py: frame starts at -1 */
  e->reason = GW_REASON_NONE;
  render(o, e);                                                         /* engine.py:580 */
  engine_play(o, e, 0, 0);                                              /* engine.py:581 play(None) */
  e->step_type = GW_STEP_FIRST;
  memset(e->episode_return, 0, sizeof e->episode_return);
}

/* ObservationToArrayWithRGBEx.__call__ + calculate_observation_layers_cube
 * (observation_distiller_ex.py:147-189, pycolab/rendering.py:491-549, safety_game_mo.py:487-506):
 * gap layer := gap AND NOT(any other layer); cube channels in sorted layer-key order. */
static void distill(const Oracle* o, const OrEnv* e, uint8_t* board, uint8_t* cube, float* value_board) {
  const GwConfig* c = &o->cfg;
  const int cells = o->cells;
  if (board) memcpy(board, e->board, (size_t)cells);
  if (value_board)
    for (int i = 0; i < cells; ++i) value_board[i] = c->value_map[e->board[i] & 127];
  if (cube) {
    for (int l = 0; l < c->n_layers; ++l) {
      uint8_t* dst = cube + (size_t)l * cells;
      memcpy(dst, e->layers[l], (size_t)cells);
      if (c->layer_chars[l] == ' ')
        for (int m = 0; m < c->n_layers; ++m)
          if (m != l)
            for (int i = 0; i < cells; ++i) dst[i] &= (uint8_t)!e->layers[m][i];
    }
  }
}

typedef struct {
  uint8_t* board; uint8_t* cube; float* value_board;
  float* reward; uint8_t* terminated; uint8_t* step_type; int8_t* reason;
} OrOut;

static void emit(const Oracle* o, const OrEnv* e, int64_t i, const OrOut* out, int write_result,
                 const double* reward, int step_type, int reason) {
  const int cells = o->cells, L = o->cfg.n_layers, R = o->cfg.n_rewards;
  distill(o, e, out->board ? out->board + i * cells : 0, out->cube ? out->cube + i * L * cells : 0,
          out->value_board ? out->value_board + i * cells : 0);
  if (!write_result) return;
  if (out->reward) for (int d = 0; d < R; ++d) out->reward[i * R + d] = (float)reward[d];
  if (out->terminated) out->terminated[i] = (uint8_t)(step_type == GW_STEP_LAST);
  if (out->step_type) out->step_type[i] = (uint8_t)step_type;
  if (out->reason) out->reason[i] = (int8_t)reason;
}

/* EnvironmentMo.step + SafetyEnvironmentMo._process_timestep for one environment
 * (rl/pycolab_interface_mo.py:157-196,308-319; safety_game_mo.py:971-1084). */
static void env_step(const Oracle* o, OrEnv* e, int64_t i, int action, const OrOut* out) {
  const GwConfig* c = &o->cfg;
  static const double zeros[MAXR] = {0};
  if (e->step_type == GW_STEP_LAST) {                     /* pycolab_interface_mo.py:175-178 */
    env_reset(o, e);
    emit(o, e, i, out, 1, zeros, GW_STEP_FIRST, GW_REASON_NONE);
    return;
  }
  engine_play(o, e, 1, action);
  int game_over = e->game_over;
  if (e->frame >= c->max_iterations) game_over = 1;       /* pycolab_interface_mo.py:318-319 */
  const int step_type = game_over ? GW_STEP_LAST : GW_STEP_MID;
  for (int d = 0; d < c->n_rewards; ++d) e->episode_return[d] += e->frame_reward[d];   /* safety_game_mo.py:996-997 */
  if (step_type == GW_STEP_LAST && e->reason == GW_REASON_NONE) e->reason = GW_REASON_MAX_STEPS;  /* :1004-1007 */
  e->step_type = step_type;
  if (step_type == GW_STEP_LAST && c->autoreset_mode == GW_AUTORESET_SAME_STEP) {
    double reward[MAXR];
    memcpy(reward, e->frame_reward, sizeof reward);
    const int reason = e->reason;
    env_reset(o, e);
    emit(o, e, i, out, 1, reward, GW_STEP_LAST, reason);
    return;
  }
  emit(o, e, i, out, 1, e->frame_reward, step_type, e->reason);
}

/* ------------------------------------------ public API ------------------------------------ */

void* or_create(const GwConfig* cfg, int64_t n) {
  if (!cfg || cfg->abi_version != GW_ABI_VERSION || n <= 0) return 0;
  if (cfg->height * cfg->width > MAXC || cfg->n_layers > MAXL || cfg->n_rewards > MAXR) return 0;
  Oracle* o = (Oracle*)calloc(1, sizeof *o);
  o->cfg = *cfg;
  o->n = n;
  o->cells = cfg->height * cfg->width;
  o->envs = (OrEnv*)calloc((size_t)n, sizeof(OrEnv));
  for (int64_t i = 0; i < n; ++i) o->envs[i].step_type = -1;
  return o;
}

void or_destroy(void* h) {
  Oracle* o = (Oracle*)h;
  if (!o) return;
  free(o->envs);
  free(o);
}

void or_reset(void* h, const uint8_t* mask, uint8_t* board, uint8_t* cube, float* value_board,
              float* reward, uint8_t* terminated, uint8_t* step_type, int8_t* reason) {
  Oracle* o = (Oracle*)h;
  OrOut out = {board, cube, value_board, reward, terminated, step_type, reason};
  static const double zeros[MAXR] = {0};
  for (int64_t i = 0; i < o->n; ++i) {
    const int doit = !mask || mask[i];
    if (doit) env_reset(o, &o->envs[i]);
    emit(o, &o->envs[i], i, &out, doit, zeros, GW_STEP_FIRST, GW_REASON_NONE);
  }
}

typedef struct {
  Oracle* o; const int32_t* actions; OrOut out; int64_t lo, hi;
} StepJob;

static void* step_range(void* arg) {
  StepJob* j = (StepJob*)arg;
  for (int64_t i = j->lo; i < j->hi; ++i) env_step(j->o, &j->o->envs[i], i, j->actions[i], &j->out);
  return 0;
}

/* n_threads <= 1: scalar loop on the calling thread; else the range is split over pthreads. */
void or_step(void* h, const int32_t* actions, uint8_t* board, uint8_t* cube, float* value_board,
             float* reward, uint8_t* terminated, uint8_t* step_type, int8_t* reason, int n_threads) {
  Oracle* o = (Oracle*)h;
  OrOut out = {board, cube, value_board, reward, terminated, step_type, reason};
  if (n_threads <= 1) {
    StepJob j = {o, actions, out, 0, o->n};
    step_range(&j);
    return;
  }
  if (n_threads > 1024) n_threads = 1024;
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)n_threads);
  StepJob* jobs = (StepJob*)malloc(sizeof(StepJob) * (size_t)n_threads);
  const int64_t per = (o->n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    int64_t lo = t * per, hi = lo + per;
    if (lo > o->n) lo = o->n;
    if (hi > o->n) hi = o->n;
    jobs[t] = (StepJob){o, actions, out, lo, hi};
    pthread_create(&tid[t], 0, step_range, &jobs[t]);
  }
  for (int t = 0; t < n_threads; ++t) pthread_join(tid[t], 0);
  free(tid);
  free(jobs);
}

void or_observe(void* h, double* metrics, float* cumulative, int32_t* frame, int16_t* pos, int16_t* safety) {
  Oracle* o = (Oracle*)h;
  const GwConfig* c = &o->cfg;
  for (int64_t i = 0; i < o->n; ++i) {
    const OrEnv* e = &o->envs[i];
    if (metrics)
      for (int m = 0; m < c->n_metrics; ++m) {
        double v = 0;
        switch (c->metric_slots[m]) {
          case GW_ISL_M_GAP_VISITS: v = e->gap_visits; break;
          case GW_ISL_M_DRINK_VISITS: v = e->drink_visits; break;
          case GW_ISL_M_FOOD_VISITS: v = e->food_visits; break;
          case GW_ISL_M_GOLD_VISITS: v = e->gold_visits; break;
          case GW_ISL_M_SILVER_VISITS: v = e->silver_visits; break;
          case GW_ISL_M_DRINK_SATIATION: v = e->drink_satiation; break;
          case GW_ISL_M_FOOD_SATIATION: v = e->food_satiation; break;
          case GW_ISL_M_DRINK_AVAILABILITY: v = e->drink_availability; break;
          case GW_ISL_M_FOOD_AVAILABILITY: v = e->food_availability; break;
        }
        metrics[i * c->n_metrics + m] = v;
      }
    if (cumulative)
      for (int d = 0; d < c->n_rewards; ++d) cumulative[i * c->n_rewards + d] = (float)e->episode_return[d];
    if (frame) frame[i] = e->frame;
    if (pos) { pos[2 * i] = (int16_t)e->pos_r; pos[2 * i + 1] = (int16_t)e->pos_c; }
    if (safety) safety[i] = (int16_t)e->safety;
  }
}

/* Hidden state for white-box tests: regrowth fractions (not observable in the reference). */
void or_peek_fractions(void* h, double* drink_fraction, double* food_fraction) {
  Oracle* o = (Oracle*)h;
  for (int64_t i = 0; i < o->n; ++i) {
    drink_fraction[i] = o->envs[i].drink_fraction;
    food_fraction[i] = o->envs[i].food_fraction;
  }
}

/* ---- Philox4x32-10 (Salmon et al., SC'11), the published algorithm; key = seed, counter =
 * (env lo, env hi, step lo, step hi).  The CUDA library implements the same function. ---- */
static void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

void or_philox(uint64_t seed, uint64_t env, uint64_t step, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)step, (uint32_t)(step >> 32)};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  memcpy(out, c, sizeof c);
}

typedef struct { uint64_t seed, step; int64_t base; int32_t lo; uint32_t span; int32_t* actions; } ActCtx;
static void actions_range(void* ctx, int64_t lo, int64_t hi) {
  const ActCtx* a = (const ActCtx*)ctx;
  for (int64_t i = lo; i < hi; ++i) {
    uint32_t r[4];
    or_philox(a->seed, (uint64_t)(a->base + i), a->step, r);
    a->actions[i] = a->lo + (int32_t)(((uint64_t)r[0] * a->span) >> 32);
  }
}

void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx);

void or_random_actions(uint64_t seed, uint64_t step, int64_t env_index_base, int32_t lo, int32_t hi,
                       int32_t* actions, int64_t n) {
  ActCtx a = {seed, step, env_index_base, lo, (uint32_t)(hi - lo + 1), actions};
  or_parallel_for(n, actions_range, &a);
}

/* ------------------------------------------------------------------------------------------ */
/* Host threads for the batched loops of every oracle in this library (plain pthreads; 1 = the   */
/* scalar loop on the calling thread).  Used by the BASELINE-size parity tests and bench.py's CPU */
/* legs, which step up to a million environments per call.  Environments are independent, so a   */
/* contiguous range per thread changes nothing in the results.                                    */
static int g_or_threads = 1;
void or_set_threads(int n) { g_or_threads = n < 1 ? 1 : (n > 1024 ? 1024 : n); }
int or_threads(void) { return g_or_threads; }

typedef struct { void (*fn)(void*, int64_t, int64_t); void* ctx; int64_t lo, hi; } ParJob;
static void* par_entry(void* arg) { ParJob* j = (ParJob*)arg; j->fn(j->ctx, j->lo, j->hi); return 0; }

void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx) {
  int nt = g_or_threads;
  if (nt > n) nt = (int)(n > 0 ? n : 1);
  if (nt <= 1) { fn(ctx, 0, n); return; }
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
  ParJob* jobs = (ParJob*)malloc(sizeof(ParJob) * (size_t)nt);
  const int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    int64_t lo = t * per, hi = lo + per;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    jobs[t] = (ParJob){fn, ctx, lo, hi};
    if (t + 1 == nt) par_entry(&jobs[t]);                       /* the calling thread takes the last range */
    else pthread_create(&tid[t], 0, par_entry, &jobs[t]);
  }
  for (int t = 0; t + 1 < nt; ++t) pthread_join(tid[t], 0);
  free(tid);
  free(jobs);
}

/* Position-weighted 64-bit checksum of a buffer of n_words 64-bit words:
 * sum_j word[j] * ((j * K1 + K2) | 1) mod 2^64.  Every weight is odd, so any change of one word
 * changes the sum.  The BASELINE-size parity tests compute the same sum on the device with torch
 * int64 arithmetic (tests/scale_util.py) and compare it per step instead of shipping 0.5 GB per
 * step to the host; on a mismatch they fall back to the element-wise comparison. */
typedef struct { const uint64_t* w; uint64_t partial[1024]; int64_t per; } SumCtx;
static void sum_range(void* ctx, int64_t lo, int64_t hi) {
  SumCtx* c = (SumCtx*)ctx;
  uint64_t total = 0;
  for (int64_t j = lo; j < hi; ++j)
    total += c->w[j] * ((((uint64_t)j * 0x9E3779B97F4A7C15ull) + 0xD1B54A32D192ED03ull) | 1ull);
  c->partial[c->per > 0 ? lo / c->per : 0] = total;
}
uint64_t or_checksum64(const void* p, int64_t n_words) {
  SumCtx* c = (SumCtx*)calloc(1, sizeof(SumCtx));
  int nt = g_or_threads;
  if (nt > n_words) nt = (int)(n_words > 0 ? n_words : 1);
  c->w = (const uint64_t*)p;
  c->per = nt > 1 ? (n_words + nt - 1) / nt : 0;                /* the same split or_parallel_for makes */
  or_parallel_for(n_words, sum_range, c);
  uint64_t total = 0;
  for (int t = 0; t < 1024; ++t) total += c->partial[t];
  free(c);
  return total;
}
