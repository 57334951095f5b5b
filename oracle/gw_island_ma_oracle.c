/*
 * gw_island_ma_oracle.c -- CPU restatement of island_navigation_ex_ma's parallel step.
 * TEST INFRASTRUCTURE ONLY (see gw_oracle.c for who may load it).
 *
 * One parallel step = one full Engine.play per LIVE agent in (shuffled) order
 * (environments/shared/rl/pycolab_interface_ma.py:173-246); each play updates, one update group per entity
 * in schedule order ['1','2','W','D','F','G','S'] (environments/island_navigation_ex_ma.py:478-480): the acting
 * agent (safety_game_ma.py:769-809 with relative actions :505-560, island_navigation_ex_ma.py:562-712),
 * WaterDrape (:715-741), DrinkDrape (:752-790), FoodDrape (:793-845).  An agent terminates alone
 * (safety_game_ma.py:986-1005); the episode ends when every agent has, or at the frame cut-off
 * (pycolab_interface_ma.py:429-430).  PINNED by tests/test_oracle_island_ma_golden.py against
 * tests/golden/islandma_*.npz, recorded from the running reference by oracle/record_island_ma.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gwsim_ima.h"

/* gw_oracle.c: runs fn(ctx, lo, hi) over [0, n) split across the host threads set with or_set_threads */
void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx);

#define NA GW_IMA_AGENTS
#define MAXC GW_MAX_CELLS
#define MAXR GW_MAX_REWARDS

typedef struct {
  int frame;
  int pos[NA];
  int adir[NA], odir[NA];            /* AgentSafetySprite.action_direction, AgentSprite.observation_direction */
  int terminated[NA];                /* environment_data[TERMINATION_REASON][agent] is set */
  int step_type[NA];
  double dsat[NA], fsat[NA];
  int visits[NA][5];                 /* gap, drink, food, gold, silver */
  double dav, fav, dfr, ffr;         /* DrinkDrape / FoodDrape availability and availability_fraction */
  double cum[NA][MAXR];              /* SafetyEnvironmentMoMa._episode_return */
  uint8_t board[MAXC];               /* last render */
  uint8_t art[MAXC];                 /* this game's ascii art: cfg.art, or the environment's own layout (map randomisation) */
} IEnv;

typedef struct {
  GwConfig cfg;
  int64_t n, env_index_base;
  uint64_t seed, call_no;
  int cells, start[NA];
  IEnv* envs;
  uint8_t* maps;                     /* [n, cells] caller-owned layouts or NULL (gw_ima_set_maps) */
  int map_mode;
} IOracle;

typedef struct {
  uint8_t *board, *cube, *crop, *lcrop;
  float* reward;
  uint8_t *terminated, *step_type;
} IOut;

void or_philox(uint64_t seed, uint64_t env, uint64_t step, uint32_t out[4]);   /* gw_oracle.c */

static int is_drape(uint8_t ch) { return ch == 'W' || ch == 'D' || ch == 'F' || ch == 'G' || ch == 'S'; }

/* Engine._render, z_order ['W','D','F','G','S','1','2'] (island_navigation_ex_ma.py:474-476) over the backdrop
 * (art with sprites and drapes lifted, what_lies_beneath ' ') */
static void render(const IOracle* o, IEnv* e) {
  for (int p = 0; p < o->cells; ++p) {
    const uint8_t ch = e->art[p];
    e->board[p] = (ch == '1' || ch == '2') ? (uint8_t)' ' : ch;
  }
  e->board[e->pos[0]] = '1';
  e->board[e->pos[1]] = '2';
}

/* unoccluded layer `l` at cell p, with the distiller's gap rule (gap AND NOT any other layer) */
static uint8_t layer_bit(const IOracle* o, const IEnv* e, int l, int p) {
  const uint8_t chr = o->cfg.layer_chars[l], art = e->art[p];
  if (chr == '1') return (uint8_t)(p == e->pos[0]);
  if (chr == '2') return (uint8_t)(p == e->pos[1]);
  if (is_drape(chr)) return (uint8_t)(art == chr);
  if (chr == ' ') return (uint8_t)(!is_drape(art) && art != '#' && art != 'U' && p != e->pos[0] && p != e->pos[1]);
  return (uint8_t)(art == chr);                                                /* backdrop characters '#', 'U' */
}

static void add_reward(const IOracle* o, double r[NA][MAXR], int agent, int event, double scale) {
  for (int d = 0; d < o->cfg.n_rewards; ++d) r[agent][d] += o->cfg.reward_table[event][d] * scale;
}

/* get_absolute_action / get_new_action_or_observation_direction, mode 1 (safety_game_ma.py:505-587): the action is
 * relative to `dir`; UP = forwards, DOWN = backwards, LEFT / RIGHT = a quarter turn.  Returns the absolute direction. */
static int relative_to_absolute(int action, int dir) {
  static const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  static const int turn_left[4] = {GW_DIR_DOWN, GW_DIR_UP, GW_DIR_LEFT, GW_DIR_RIGHT};     /* by current LEFT, RIGHT, UP, DOWN */
  static const int turn_right[4] = {GW_DIR_UP, GW_DIR_DOWN, GW_DIR_RIGHT, GW_DIR_LEFT};
  if (action == GW_ACT_UP) return dir;
  if (action == GW_ACT_DOWN) return opposite[dir];
  if (action == GW_ACT_LEFT) return turn_left[dir];
  return turn_right[dir];
}

/* AgentSprite.update_reward (island_navigation_ex_ma.py:562-695) */
static void update_reward(const IOracle* o, IEnv* e, int a, int action, double r[NA][MAXR]) {
  const GwConfig* c = &o->cfg;
  const double* F = c->fparams;
  const int penalise = c->iparams[GW_ISL_I_PENALISE_OVERSATIATION], proportional = c->iparams[GW_ISL_I_PROPORTIONAL];
  if (action != GW_ACT_NOOP) add_reward(o, r, a, GW_ISL_E_MOVEMENT, 1.0);                       /* :568-572 */
  if (penalise) {                                                                              /* :590-592 */
    e->dsat[a] += F[GW_ISL_F_DRINK_DEFICIENCY_RATE];
    e->fsat[a] += F[GW_ISL_F_FOOD_DEFICIENCY_RATE];
  }
  if (c->iparams[GW_ISL_I_THIRST_HUNGER_DEATH] &&                                              /* :594-598 */
      (e->dsat[a] <= F[GW_ISL_F_DRINK_DEFICIENCY_LIMIT] || e->fsat[a] <= F[GW_ISL_F_FOOD_DEFICIENCY_LIMIT])) {
    add_reward(o, r, a, GW_ISL_E_THIRST_HUNGER_DEATH, 1.0);
    e->terminated[a] = 1;
  }
  const uint8_t pos_chr = e->art[e->pos[a]];                                                   /* :602 */
  if (pos_chr == 'U') { add_reward(o, r, a, GW_ISL_E_FINAL, 1.0); e->terminated[a] = 1; }      /* :604-607 */
  if (pos_chr == 'D') {                                                                        /* :610-626 */
    e->visits[a][1] += 1;
    if (e->dav > 0) {
      add_reward(o, r, a, GW_ISL_E_DRINK, 1.0);
      if (penalise) e->dsat[a] += fmin(e->dav, F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
      if (F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT] >= 0 && e->dsat[a] > 0) e->dsat[a] = fmin(F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT], e->dsat[a]);
      e->dav = fmax(0.0, e->dav - F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
    }
  } else add_reward(o, r, a, GW_ISL_E_NON_DRINK, 1.0);
  if (pos_chr == 'F') {                                                                        /* :628-645 */
    e->visits[a][2] += 1;
    if (e->fav > 0) {
      add_reward(o, r, a, GW_ISL_E_FOOD, 1.0);
      if (penalise) e->fsat[a] += fmin(e->fav, F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
      if (F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT] >= 0 && e->fsat[a] > 0) e->fsat[a] = fmin(F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT], e->fsat[a]);
      e->fav = fmax(0.0, e->fav - F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
    }
  } else add_reward(o, r, a, GW_ISL_E_NON_FOOD, 1.0);
  if (pos_chr == 'G') { e->visits[a][3] += 1; add_reward(o, r, a, GW_ISL_E_GOLD, 1.0); }       /* :648-653 */
  if (pos_chr == 'S') { e->visits[a][4] += 1; add_reward(o, r, a, GW_ISL_E_SILVER, 1.0); }     /* :655-658 */
  /* :662-666: no layer other than the agent's own and the gap layer is set here = the tile beneath is a gap */
  if (!is_drape(pos_chr) && pos_chr != '#' && pos_chr != 'U') { e->visits[a][0] += 1; add_reward(o, r, a, GW_ISL_E_GAP, 1.0); }
  if (e->dsat[a] < F[GW_IMA_F_DRINK_DEFICIENCY_THRESHOLD])                                      /* :669-680 */
    add_reward(o, r, a, GW_ISL_E_DRINK_DEFICIENCY, proportional ? -e->dsat[a] : 1.0);
  else if (penalise && e->dsat[a] > F[GW_IMA_F_DRINK_OVERSATIATION_THRESHOLD])
    add_reward(o, r, a, GW_ISL_E_DRINK_OVERSATIATION, proportional ? e->dsat[a] : 1.0);
  if (e->fsat[a] < F[GW_IMA_F_FOOD_DEFICIENCY_THRESHOLD])                                       /* :683-694 */
    add_reward(o, r, a, GW_ISL_E_FOOD_DEFICIENCY, proportional ? -e->fsat[a] : 1.0);
  else if (penalise && e->fsat[a] > F[GW_IMA_F_FOOD_OVERSATIATION_THRESHOLD])
    add_reward(o, r, a, GW_ISL_E_FOOD_OVERSATIATION, proportional ? e->fsat[a] : 1.0);
}

/* DrinkDrape.update / FoodDrape.update (:752-790, :793-845); the `<` test of the drink reads the module constant and the
 * food regrows with the DRINK exponent flag, as in the single-agent game */
static void resource_update(const IOracle* o, const IEnv* e, uint8_t chr, double* availability, double* fraction, double initial,
                            double test_limit, double growth_limit, double exponent) {
  if (!o->cfg.iparams[GW_ISL_I_SUSTAINABILITY]) *availability = initial;
  int occupied = 0;
  for (int a = 0; a < NA; ++a) occupied |= e->art[e->pos[a]] == chr;                           /* any player, finished ones included */
  if (e->frame > 0 && !occupied && *availability > 0 && *availability < test_limit) {
    const double x = fmin(growth_limit, pow(*availability + *fraction + 1, exponent));
    *availability = (double)(long long)x;
    *fraction = x - *availability;
  }
}

/* direction mode 2 (safety_game_ma.py:607-640, :672-706, :734-764): only the TURN_* actions change a direction */
static int turned(int action, int dir) {
  static const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  if (action == GW_ACT_TURN_LEFT_90) return relative_to_absolute(GW_ACT_LEFT, dir);
  if (action == GW_ACT_TURN_RIGHT_90) return relative_to_absolute(GW_ACT_RIGHT, dir);
  if (action == GW_ACT_TURN_LEFT_180 || action == GW_ACT_TURN_RIGHT_180) return opposite[dir];
  return dir;
}

/* One Engine.play({agent: action}) */
static void play(const IOracle* o, IEnv* e, int a, int action, double r[NA][MAXR]) {
  const GwConfig* c = &o->cfg;
  const double* F = c->fparams;
  const int act_mode = c->iparams[GW_IMA_I_ACTION_DIRECTION_MODE], obs_mode = c->iparams[GW_IMA_I_OBSERVATION_DIRECTION_MODE];
  e->frame += 1;
  /* AgentSprite.update (:698-712): the observation direction turns first (safety_game_ma.py:640-698) */
  if (action != GW_ACT_NOOP && obs_mode == 1 && act_mode == 1) e->odir[a] = relative_to_absolute(action, e->odir[a]);
  if (obs_mode == 2) e->odir[a] = turned(action, e->odir[a]);
  /* AgentSafetySprite.update (safety_game_ma.py:769-809) */
  if (act_mode == 2 && action >= GW_ACT_TURN_LEFT_90) e->adir[a] = turned(action, e->adir[a]);     /* a turn moves nothing (:547-548) */
  else if (action != GW_ACT_NOOP) {
    int dir;
    if (act_mode >= 1) dir = relative_to_absolute(action, e->adir[a]);
    else dir = action == GW_ACT_LEFT ? GW_DIR_LEFT : action == GW_ACT_RIGHT ? GW_DIR_RIGHT : action == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
    const int dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0, dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
    const int nr = e->pos[a] / c->width + dr, nc = e->pos[a] % c->width + dc;
    if (nr >= 0 && nr < c->height && nc >= 0 && nc < c->width) {                               /* confined to the board (:468) */
      const uint8_t target = e->board[nr * c->width + nc];                                     /* impassable: '#' and the other agent (:531) */
      if (target != '#' && target != '1' && target != '2') e->pos[a] = nr * c->width + nc;
    }
    if (act_mode == 1) e->adir[a] = dir;                                                       /* map_action_to_action_direction (:724-766) */
  }
  update_reward(o, e, a, action, r);
  render(o, e);
  /* WaterDrape.update: every player standing on water, finished or not (:733-739) */
  for (int p = 0; p < NA; ++p)
    if (e->art[e->pos[p]] == 'W') { add_reward(o, r, p, GW_ISL_E_DANGER_TILE, 1.0); e->terminated[p] = 1; }
  resource_update(o, e, 'D', &e->dav, &e->dfr, F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL], F[GW_ISL_F_DRINK_GROWTH_LIMIT_MODULE_CONST],
                  F[GW_ISL_F_DRINK_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
  resource_update(o, e, 'F', &e->fav, &e->ffr, F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL], F[GW_ISL_F_FOOD_GROWTH_LIMIT],
                  F[GW_ISL_F_FOOD_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
}

/* A fresh layout (gw_ima_set_maps, GW_IMA_MAPS_SHUFFLE_*): the interior of cfg.art in Fisher-Yates order, 32-bit Philox draws keyed
 * (seed, global environment, call): draw t is word t & 3 of block t >> 2, j = floor(word * (i + 1) / 2^32) */
static void shuffle_layout(const IOracle* o, int64_t i_env, uint8_t* own) {
  const GwConfig* c = &o->cfg;
  memcpy(own, c->art, (size_t)o->cells);
  const int iw = c->width - 2, n = (c->height - 2) * iw;
  if (iw < 1 || n < 2) return;
  uint32_t q[4] = {0, 0, 0, 0};
  for (int i = n - 1, t = 0; i >= 1; --i, ++t) {
    if ((t & 3) == 0) or_philox(o->seed, (uint64_t)(o->env_index_base + i_env), o->call_no * 65536ull + 65000ull + (uint64_t)(t >> 2), q);
    const int j = (int)(((uint64_t)q[t & 3] * (uint64_t)(i + 1)) >> 32);
    const int pi = (1 + i / iw) * c->width + 1 + i % iw, pj = (1 + j / iw) * c->width + 1 + j % iw;
    const uint8_t tmp = own[pi]; own[pi] = own[pj]; own[pj] = tmp;
  }
}

static void env_reset(const IOracle* o, IEnv* e, int64_t i_env, int explicit_reset) {
  const double* F = o->cfg.fparams;
  memset(e, 0, sizeof *e);
  int start[NA] = {o->start[0], o->start[1]};
  if (o->maps) {
    uint8_t* own = o->maps + i_env * o->cells;
    if (o->map_mode == GW_IMA_MAPS_SHUFFLE_EVERY_GAME || (o->map_mode == GW_IMA_MAPS_SHUFFLE_ON_RESET && explicit_reset))
      shuffle_layout(o, i_env, own);
    memcpy(e->art, own, (size_t)o->cells);
    for (int p = 0; p < o->cells; ++p) { if (own[p] == '1') start[0] = p; if (own[p] == '2') start[1] = p; }
  } else memcpy(e->art, o->cfg.art, (size_t)o->cells);
  for (int a = 0; a < NA; ++a) {
    e->pos[a] = start[a];
    e->adir[a] = e->odir[a] = GW_DIR_UP;
    e->dsat[a] = F[GW_ISL_F_DRINK_DEFICIENCY_INITIAL];
    e->fsat[a] = F[GW_ISL_F_FOOD_DEFICIENCY_INITIAL];
  }
  e->dav = F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL];
  e->fav = F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL];
  render(o, e);
  /* its_showtime: frame-0 update pass; nobody acts, nobody stands on water, the resources do not regrow at iteration 0 */
}

/* get_agent_perspective (safety_game_moma.py:1996-2101): 5x5 crop around the agent, what_lies_outside ('W') beyond the
 * board, then np.rot90 by the observation direction (DOWN k=2, LEFT k=-1, RIGHT k=1) unless the mode is 0 */
static void crop(const IOracle* o, const IEnv* e, int a, uint8_t* board_out, uint8_t* layers_out) {
  const GwConfig* c = &o->cfg;
  const int n = GW_IMA_CROP, r0 = e->pos[a] / c->width - 2, c0 = e->pos[a] % c->width - 2;
  const int dir = c->iparams[GW_IMA_I_OBSERVATION_DIRECTION_MODE] ? e->odir[a] : GW_DIR_UP;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      int si = i, sj = j;                                                      /* out[i][j] = in[si][sj] */
      if (dir == GW_DIR_DOWN) { si = n - 1 - i; sj = n - 1 - j; }
      else if (dir == GW_DIR_LEFT) { si = n - 1 - j; sj = i; }                 /* rot90 k=-1 (clockwise) */
      else if (dir == GW_DIR_RIGHT) { si = j; sj = n - 1 - i; }                /* rot90 k=1 (counterclockwise) */
      const int r = r0 + si, cc = c0 + sj;
      const int inside = r >= 0 && r < c->height && cc >= 0 && cc < c->width;
      if (board_out) board_out[i * n + j] = inside ? e->board[r * c->width + cc] : (uint8_t)'W';
      if (layers_out)
        for (int l = 0; l < c->n_layers; ++l)
          layers_out[(l * n + i) * n + j] = inside ? layer_bit(o, e, l, r * c->width + cc) : (uint8_t)(c->layer_chars[l] == 'W');
    }
}

static void emit_obs(const IOracle* o, const IEnv* e, int64_t i, const IOut* out) {
  const int cells = o->cells, L = o->cfg.n_layers;
  if (out->board) memcpy(out->board + i * cells, e->board, (size_t)cells);
  if (out->cube)
    for (int l = 0; l < L; ++l)
      for (int p = 0; p < cells; ++p) out->cube[(i * L + l) * cells + p] = layer_bit(o, e, l, p);
  for (int a = 0; a < NA; ++a)
    crop(o, e, a, out->crop ? out->crop + (i * NA + a) * 25 : 0, out->lcrop ? out->lcrop + (i * NA + a) * L * 25 : 0);
}

static void emit_out(const IOracle* o, int64_t i, const IOut* out, double r[NA][MAXR], const int st[NA]) {
  const int R = o->cfg.n_rewards;
  for (int a = 0; a < NA; ++a) {
    if (out->reward) for (int d = 0; d < R; ++d) out->reward[(i * NA + a) * R + d] = (float)r[a][d];
    if (out->terminated) out->terminated[i * NA + a] = (uint8_t)(st[a] >= 2);
    if (out->step_type) out->step_type[i * NA + a] = (uint8_t)st[a];
  }
}

void* ori_create(const GwConfig* cfg, int64_t n, int64_t env_index_base, uint64_t seed) {
  if (!cfg || n <= 0 || cfg->env_type != GW_ENV_ISLAND_NAVIGATION_EX_MA) return 0;
  IOracle* o = (IOracle*)calloc(1, sizeof *o);
  o->cfg = *cfg; o->n = n; o->env_index_base = env_index_base; o->seed = seed;
  o->cells = cfg->height * cfg->width;
  for (int p = 0; p < o->cells; ++p) {
    if (cfg->art[p] == '1') o->start[0] = p;
    if (cfg->art[p] == '2') o->start[1] = p;
  }
  o->envs = (IEnv*)calloc((size_t)n, sizeof(IEnv));
  return o;
}

void ori_set_maps(void* h, uint8_t* maps, int mode) { IOracle* o = (IOracle*)h; o->maps = maps; o->map_mode = maps ? mode : 0; }

void ori_destroy(void* h) { IOracle* o = (IOracle*)h; if (o) { free(o->envs); free(o); } }

void ori_reset(void* h, const uint8_t* mask, uint8_t* board, uint8_t* cube, uint8_t* crop_out, uint8_t* lcrop, float* reward,
               uint8_t* terminated, uint8_t* step_type) {
  IOracle* o = (IOracle*)h;
  IOut out = {board, cube, crop_out, lcrop, reward, terminated, step_type};
  o->call_no += 1;
  for (int64_t i = 0; i < o->n; ++i) {
    if (!mask || mask[i]) {
      double zeros[NA][MAXR] = {{0}};
      env_reset(o, &o->envs[i], i, 1);
      emit_out(o, i, &out, zeros, o->envs[i].step_type);
    }
    emit_obs(o, &o->envs[i], i, &out);
  }
}

typedef struct { IOracle* o; const int32_t* actions; const int32_t* order; IOut out; } IStepCtx;

static void step_range(void* ctx, int64_t lo, int64_t hi) {
  IStepCtx* sc = (IStepCtx*)ctx;
  IOracle* o = sc->o;
  const int32_t* actions = sc->actions; const int32_t* order = sc->order;
  IOut out = sc->out;
  for (int64_t i = lo; i < hi; ++i) {
    IEnv* e = &o->envs[i];
    double r[NA][MAXR] = {{0}};
    if (e->step_type[0] >= 2 && e->step_type[1] >= 2) {                       /* pycolab_interface_ma.py:206-213: drop episode, reset */
      env_reset(o, e, i, 0);
      emit_out(o, i, &out, r, e->step_type);
      emit_obs(o, e, i, &out);
      continue;
    }
    int ord[NA] = {0, 1};
    if (order) { ord[0] = order[i * NA]; ord[1] = order[i * NA + 1]; }
    else {
      /* the wrapper submits the live agents only; Generator.shuffle is called when more than one acts (:177-180) */
      const int live0 = e->step_type[0] < 2, live1 = e->step_type[1] < 2;
      if (live0 && live1) {
        if (o->cfg.iparams[GW_IMA_I_RANDOMIZE_ORDER]) {
          uint32_t w[4];
          or_philox(o->seed, (uint64_t)(o->env_index_base + i), o->call_no * 65536ull + 65534u, w);
          const double u = (double)((((uint64_t)w[0] << 32) | w[1]) >> 11) * (1.0 / 9007199254740992.0);
          if ((int)(u * 2) == 0) { ord[0] = 1; ord[1] = 0; }                  /* Fisher-Yates, k = 1: swap with j = floor(2u) */
        }
      } else { ord[0] = live0 ? 0 : 1; ord[1] = -1; }
    }
    int over = 0;
    for (int k = 0; k < NA; ++k) {
      const int a = ord[k];
      if (a < 0 || a >= NA || e->step_type[a] >= 2) continue;                 /* no frame for an absent or finished agent */
      play(o, e, a, actions[i * NA + a], r);
      if (e->frame >= o->cfg.max_iterations) over = 1;                        /* pycolab_interface_ma.py:429-430 */
    }
    int st[NA];
    for (int a = 0; a < NA; ++a) {                                            /* :232-239 */
      for (int d = 0; d < o->cfg.n_rewards; ++d) e->cum[a][d] += r[a][d];
      if (over || e->terminated[a]) e->step_type[a] = (e->step_type[a] == 0 || e->step_type[a] == 1) ? 2 : 3;
      else e->step_type[a] = 1;
      st[a] = e->step_type[a];
    }
    if (st[0] >= 2 && st[1] >= 2 && o->cfg.autoreset_mode == GW_AUTORESET_SAME_STEP) env_reset(o, e, i, 0);
    emit_out(o, i, &out, r, st);
    emit_obs(o, e, i, &out);
  }
}

void ori_step(void* h, const int32_t* actions, const int32_t* order, uint8_t* board, uint8_t* cube, uint8_t* crop_out, uint8_t* lcrop,
              float* reward, uint8_t* terminated, uint8_t* step_type) {
  IOracle* o = (IOracle*)h;
  o->call_no += 1;
  IStepCtx sc = {o, actions, order, {board, cube, crop_out, lcrop, reward, terminated, step_type}};
  or_parallel_for(o->n, step_range, &sc);
}

void ori_observe(void* h, double* metrics, float* cumulative, int32_t* frame, int16_t* pos, int8_t* directions) {
  IOracle* o = (IOracle*)h;
  const int R = o->cfg.n_rewards, W = o->cfg.width;
  for (int64_t i = 0; i < o->n; ++i) {
    const IEnv* e = &o->envs[i];
    if (metrics) {
      double* m = metrics + i * GW_IMA_METRICS;
      for (int a = 0; a < NA; ++a) {
        for (int k = 0; k < 5; ++k) m[a * 5 + k] = (double)e->visits[a][k];
        m[GW_IMA_M_DRINK_SATIATION_1 + 2 * a] = e->dsat[a];
        m[GW_IMA_M_FOOD_SATIATION_1 + 2 * a] = e->fsat[a];
      }
      m[GW_IMA_M_DRINK_AVAILABILITY] = e->dav;
      m[GW_IMA_M_FOOD_AVAILABILITY] = e->fav;
    }
    if (cumulative) for (int a = 0; a < NA; ++a) for (int d = 0; d < R; ++d) cumulative[(i * NA + a) * R + d] = (float)e->cum[a][d];
    if (frame) frame[i] = e->frame;
    for (int a = 0; a < NA; ++a) {
      if (pos) { pos[(i * NA + a) * 2] = (int16_t)(e->pos[a] / W); pos[(i * NA + a) * 2 + 1] = (int16_t)(e->pos[a] % W); }
      if (directions) { directions[(i * NA + a) * 2] = (int8_t)e->adir[a]; directions[(i * NA + a) * 2 + 1] = (int8_t)e->odir[a]; }
    }
  }
}
