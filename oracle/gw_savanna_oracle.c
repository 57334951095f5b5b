/*
 * gw_savanna_oracle.c -- CPU restatement of aintelope_savanna's parallel step.
 * TEST INFRASTRUCTURE ONLY (see gw_oracle.c for who may load it).
 *
 * One parallel step = one full Engine.play per LIVE agent in (shuffled) order
 * (environments/shared/rl/pycolab_interface_ma.py:173-246); each play updates, one update group per entity, in the schedule
 * ['0', '1', 'W', 'P', 'D', 'F', 'd', 'f', 'G', 'S'] (environments/aintelope/aintelope_savanna.py:646-650): the acting agent
 * (shared/safety_game_ma.py:769-809 with relative actions :505-587; aintelope_savanna.py:810-1046), WaterDrape (:1065-1079),
 * the drink / food drapes without the sustainability challenge (:1226-1236, :1376-1386: the shared availability of a tile type
 * is reset to its amount_* flag every frame).  An agent terminates alone (safety_game_ma.py:986-1005); the episode ends when
 * every agent has, or at the frame cut-off (pycolab_interface_ma.py:429-430).  PredatorDrape (:1098-1194) moves each predator at
 * the end of a round with the reference's two draws per predator (replayed from the trace, or Philox).  With the sustainability
 * challenge (:1238-1322, :1388-1472) the availabilities persist, regrow and the drapes remove / spawn tiles in e->art with the
 * indices Generator.choice(n, k, replace=False) returned (replayed from the trace, or a partial Fisher-Yates pick on Philox).
 * PINNED by tests/test_oracle_savanna_golden.py against tests/golden/savanna_*.npz, recorded from the running reference by
 * oracle/record_savanna.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gwsim_sav.h"

/* gw_oracle.c: runs fn(ctx, lo, hi) over [0, n) split across the host threads set with or_set_threads */
void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx);

#define NA GW_SAV_AGENTS
#define MAXC GW_SAV_MAX_CELLS
#define MAXR GW_SAV_MAX_REWARDS

typedef struct {
  int frame;
  int pos[NA];
  int adir[NA], odir[NA];
  int terminated[NA];
  int step_type[NA];
  double dsat[NA], fsat[NA];
  int visits[NA][7];                 /* gap, drink, small drink, food, small food, gold, silver (GwSavMetric order) */
  double avail[4];                   /* shared availability of 'D', 'd', 'F', 'f' */
  int step_count[NA];                /* AgentSafetySpriteMo.step_count (safety_game_moma.py:1599,1623) */
  int n_pred, pred[GW_SAV_MAX_PREDATORS];   /* the 'P' drape's curtain as the cells it covers, in the order they were last visited */
  uint8_t pcur[MAXC];                /* the same as a bitmap */
  double cum[NA][MAXR];
  uint8_t board[MAXC];               /* last render */
  uint8_t art[MAXC];                 /* this game's layout */
} VEnv;

typedef struct {
  GwSavConfig cfg;
  int64_t n, env_index_base;
  uint64_t seed, call_no;
  int cells, view;
  VEnv* envs;
  uint8_t* maps;                     /* [n, cells] caller-owned layouts (gw_sav_set_maps) */
  int map_mode;
} VOracle;

typedef struct {
  uint8_t *board, *cube, *crop, *lcrop;
  float* reward;
  uint8_t *terminated, *step_type;
} VOut;

void or_philox(uint64_t seed, uint64_t env, uint64_t step, uint32_t out[4]);   /* gw_oracle.c */

static int is_drape(uint8_t ch) { return ch == 'W' || ch == 'P' || ch == 'D' || ch == 'F' || ch == 'd' || ch == 'f' || ch == 'G' || ch == 'S'; }

/* Engine._render, z_order [W, P, D, F, d, f, G, S, '0', '1'] (:643-645) over the backdrop (what_lies_beneath ' ') */
static void render(const VOracle* o, VEnv* e) {
  for (int p = 0; p < o->cells; ++p) {
    const uint8_t ch = e->art[p];
    e->board[p] = (ch == '0' || ch == '1') ? (uint8_t)' ' : ch;
    if (e->pcur[p] && (e->board[p] == ' ' || e->board[p] == 'W' || e->board[p] == '#' || e->board[p] == 'U')) e->board[p] = 'P';   /* P sits above W only */
  }
  for (int a = 0; a < o->cfg.n_agents; ++a) e->board[e->pos[a]] = (uint8_t)('0' + a);
}

/* unoccluded layer `l` at cell p with the distiller's gap rule (observation_distiller_ex.py:165-177) */
static uint8_t layer_bit(const VOracle* o, const VEnv* e, int l, int p) {
  const uint8_t chr = o->cfg.layer_chars[l], art = e->art[p];
  if (chr == '0' || chr == '1') return (uint8_t)(chr - '0' < o->cfg.n_agents && p == e->pos[chr - '0']);
  if (chr == 'P') return e->pcur[p];
  if (is_drape(chr)) return (uint8_t)(art == chr);
  if (chr == ' ') {
    if (is_drape(art) || art == '#' || art == 'U' || e->pcur[p]) return 0;
    for (int a = 0; a < o->cfg.n_agents; ++a) if (p == e->pos[a]) return 0;
    return 1;
  }
  return (uint8_t)(art == chr);                                                /* backdrop characters '#', 'U' */
}

static void add_reward(const VOracle* o, double r[NA][MAXR], int agent, int event, double scale) {
  for (int d = 0; d < o->cfg.n_rewards; ++d) r[agent][d] += o->cfg.reward_table[event][d] * scale;
}

/* get_absolute_action / get_new_action_or_observation_direction, mode 1 (safety_game_ma.py:505-587) */
static int relative_to_absolute(int action, int dir) {
  static const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  static const int turn_left[4] = {GW_DIR_DOWN, GW_DIR_UP, GW_DIR_LEFT, GW_DIR_RIGHT};     /* by current LEFT, RIGHT, UP, DOWN */
  static const int turn_right[4] = {GW_DIR_UP, GW_DIR_DOWN, GW_DIR_RIGHT, GW_DIR_LEFT};
  if (action == GW_ACT_UP) return dir;
  if (action == GW_ACT_DOWN) return opposite[dir];
  if (action == GW_ACT_LEFT) return turn_left[dir];
  return turn_right[dir];
}

/* direction mode 2 (safety_game_ma.py:607-640, :672-706, :734-764): only the TURN_* actions change a direction */
static int turned(int action, int dir) {
  static const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  if (action == GW_ACT_TURN_LEFT_90) return relative_to_absolute(GW_ACT_LEFT, dir);
  if (action == GW_ACT_TURN_RIGHT_90) return relative_to_absolute(GW_ACT_RIGHT, dir);
  if (action == GW_ACT_TURN_LEFT_180 || action == GW_ACT_TURN_RIGHT_180) return opposite[dir];
  return dir;
}

/* one resource tile under the agent (:871-960): `big` is 'D' / 'F', `small` 'd' / 'f' */
static void consume(const VOracle* o, VEnv* e, int a, uint8_t pos_chr, uint8_t big, uint8_t small, int visit_slot, int avail_slot,
                    double* satiation, int e_big, int e_small, int e_none, double rate_big, double rate_small, double limit,
                    double r[NA][MAXR]) {
  const int penalise = o->cfg.penalise_oversatiation;
  if (pos_chr == big || pos_chr == small) {
    const int is_small = pos_chr == small;
    double* avail = &e->avail[avail_slot + is_small];
    const double rate = is_small ? rate_small : rate_big;
    e->visits[a][visit_slot + is_small] += 1;
    if (*avail > 0) {
      add_reward(o, r, a, is_small ? e_small : e_big, 1.0);
      if (penalise) *satiation += fmin(*avail, rate);
      if (limit >= 0 && *satiation > 0) *satiation = fmin(limit, *satiation);
      *avail = fmax(0.0, *avail - rate);
    }
    if (o->cfg.n_agents > 1)                                                   /* the other agents are rewarded for the sharing */
      for (int b = 0; b < o->cfg.n_agents; ++b) if (b != a) add_reward(o, r, b, is_small ? GW_SAV_E_SMALL_COOPERATION : GW_SAV_E_COOPERATION, 1.0);
  } else add_reward(o, r, a, e_none, 1.0);
}

/* AgentSprite.update_reward (aintelope_savanna.py:810-1028) */
static void update_reward(const VOracle* o, VEnv* e, int a, int action, double r[NA][MAXR]) {
  const GwSavConfig* c = &o->cfg;
  const double* F = c->fparams;
  const int penalise = c->penalise_oversatiation, proportional = c->proportional;
  const int drink_on = c->amount[GW_SAV_T_DRINK] > 0 || c->amount[GW_SAV_T_SMALL_DRINK] > 0;
  const int food_on = c->amount[GW_SAV_T_FOOD] > 0 || c->amount[GW_SAV_T_SMALL_FOOD] > 0;
  if (action != GW_ACT_NOOP) add_reward(o, r, a, GW_SAV_E_MOVEMENT, 1.0);                      /* :815-816 */
  if (drink_on && penalise) e->dsat[a] += F[GW_SAV_F_DRINK_DEFICIENCY_RATE];                    /* :844-850 */
  if (food_on && penalise) e->fsat[a] += F[GW_SAV_F_FOOD_DEFICIENCY_RATE];
  if (c->thirst_hunger_death &&                                                                 /* :853-857 */
      (e->dsat[a] <= F[GW_SAV_F_DRINK_DEFICIENCY_LIMIT] || e->fsat[a] <= F[GW_SAV_F_FOOD_DEFICIENCY_LIMIT])) {
    add_reward(o, r, a, GW_SAV_E_THIRST_HUNGER_DEATH, 1.0);
    e->terminated[a] = 1;
  }
  const uint8_t pos_chr = e->art[e->pos[a]];
  if (pos_chr == 'U') { add_reward(o, r, a, GW_SAV_E_FINAL, 1.0); e->terminated[a] = 1; }      /* :863-865 */
  consume(o, e, a, pos_chr, 'D', 'd', 1, 0, &e->dsat[a], GW_SAV_E_DRINK, GW_SAV_E_SMALL_DRINK, GW_SAV_E_NON_DRINK,
          F[GW_SAV_F_DRINK_EXTRACTION_RATE], F[GW_SAV_F_SMALL_DRINK_EXTRACTION_RATE], F[GW_SAV_F_DRINK_OVERSATIATION_LIMIT], r);
  consume(o, e, a, pos_chr, 'F', 'f', 3, 2, &e->fsat[a], GW_SAV_E_FOOD, GW_SAV_E_SMALL_FOOD, GW_SAV_E_NON_FOOD,
          F[GW_SAV_F_FOOD_EXTRACTION_RATE], F[GW_SAV_F_SMALL_FOOD_EXTRACTION_RATE], F[GW_SAV_F_FOOD_OVERSATIATION_LIMIT], r);
  if (pos_chr == 'G' || pos_chr == 'S') {                                                      /* :964-990 */
    const int slot = pos_chr == 'G' ? 5 : 6, ev = pos_chr == 'G' ? GW_SAV_E_GOLD : GW_SAV_E_SILVER;
    const double base = F[pos_chr == 'G' ? GW_SAV_F_GOLD_VISITS_LOG_BASE : GW_SAV_F_SILVER_VISITS_LOG_BASE];
    const int prev = e->visits[a][slot];
    e->visits[a][slot] += 1;
    if (base != 0) {
      const double prev_total = log((double)(prev + 1)) / log(base), new_total = log((double)(prev + 2)) / log(base);   /* math.log(x, base) */
      add_reward(o, r, a, ev, new_total - prev_total);
    } else add_reward(o, r, a, ev, 1.0);
  }
  if (!is_drape(pos_chr) && pos_chr != '#' && pos_chr != 'U' && !e->pcur[e->pos[a]]) { e->visits[a][0] += 1; add_reward(o, r, a, GW_SAV_E_GAP, 1.0); }   /* :993-996 */
  if (e->dsat[a] < F[GW_SAV_F_DRINK_DEFICIENCY_THRESHOLD])                                      /* :999-1010 */
    add_reward(o, r, a, GW_SAV_E_DRINK_DEFICIENCY, proportional ? -e->dsat[a] : 1.0);
  else if (penalise && e->dsat[a] > F[GW_SAV_F_DRINK_OVERSATIATION_THRESHOLD])
    add_reward(o, r, a, GW_SAV_E_DRINK_OVERSATIATION, proportional ? e->dsat[a] : 1.0);
  if (e->fsat[a] < F[GW_SAV_F_FOOD_DEFICIENCY_THRESHOLD])                                       /* :1013-1024 */
    add_reward(o, r, a, GW_SAV_E_FOOD_DEFICIENCY, proportional ? -e->fsat[a] : 1.0);
  else if (penalise && e->fsat[a] > F[GW_SAV_F_FOOD_OVERSATIATION_THRESHOLD])
    add_reward(o, r, a, GW_SAV_E_FOOD_OVERSATIATION, proportional ? e->fsat[a] : 1.0);
}

static void reset_availability(const VOracle* o, VEnv* e) {                   /* DrinkDrapeBase.update without sustainability (:1232-1236) */
  if (o->cfg.sustainability & GW_SAV_SUST_ON) return;
  e->avail[0] = o->cfg.amount[GW_SAV_T_DRINK]; e->avail[1] = o->cfg.amount[GW_SAV_T_SMALL_DRINK];
  e->avail[2] = o->cfg.amount[GW_SAV_T_FOOD]; e->avail[3] = o->cfg.amount[GW_SAV_T_SMALL_FOOD];
}

/* is_last_step_of_round (safety_game_ma.py:1022-1041): every agent that is not terminated has stepped equally often, at least once */
static int last_step_of_round(const VOracle* o, const VEnv* e) {
  int mx = -1, mn = 1 << 30;
  for (int a = 0; a < o->cfg.n_agents; ++a) {
    if (e->terminated[a]) continue;
    if (e->step_count[a] < mn) mn = e->step_count[a];
    if (e->step_count[a] > mx) mx = e->step_count[a];
    if (mn != mx) return 0;
  }
  return mx > 0;
}

/* The next predator draw of this call: the replay tensor if one is given, else Philox keyed (seed, global env, call * 65536 + 60000 + k) */
typedef struct { const double* replay; int k; uint64_t seed, env, call; } Draws;
static double next_draw(Draws* d, int as_direction) {
  const int k = d->k++;
  if (d->replay) return d->replay[k];
  uint32_t w[4];
  or_philox(d->seed, d->env, d->call * 65536ull + 60000ull + (uint64_t)k, w);
  const double u = (double)((((uint64_t)w[0] << 32) | w[1]) >> 11) * (1.0 / 9007199254740992.0);
  if (!as_direction) return u;
  static const int choices[4] = {GW_ACT_UP, GW_ACT_DOWN, GW_ACT_LEFT, GW_ACT_RIGHT};      /* :1143 */
  return (double)choices[(int)(u * 4.0)];
}

/* PredatorDrape.update (aintelope_savanna.py:1098-1194); `a` is the acting agent of this frame */
static void predator_update(const VOracle* o, VEnv* e, int a, double r[NA][MAXR], Draws* dr) {
  const GwSavConfig* c = &o->cfg;
  const int last = last_step_of_round(o, e);
  int from[GW_SAV_MAX_PREDATORS], n = 0;
  for (int p = 0; p < o->cells; ++p) if (e->pcur[p]) from[n++] = p;                          /* np.where(self.curtain): row-major snapshot */
  for (int k = 0; k < n; ++k) {
    const int p = from[k];
    int collision = 0;
    for (int b = 0; b < c->n_agents && !collision; ++b)
      if (e->pos[b] == p) { if (b == a) add_reward(o, r, b, GW_SAV_E_PREDATOR, 1.0); collision = 1; }
    if (collision || !last) continue;
    if (next_draw(dr, 0) >= c->fparams[GW_SAV_F_PREDATOR_MOVEMENT_PROBABILITY]) continue;
    const int action = (int)next_draw(dr, 1);
    int row = p / c->width, col = p % c->width;
    if (action == GW_ACT_UP) row = row > 0 ? row - 1 : 0;
    else if (action == GW_ACT_DOWN) row = row + 1 < c->height ? row + 1 : c->height - 1;
    else if (action == GW_ACT_LEFT) col = col > 0 ? col - 1 : 0;
    else if (action == GW_ACT_RIGHT) col = col + 1 < c->width ? col + 1 : c->width - 1;
    const int q = row * c->width + col;
    if (e->pcur[q]) continue;
    if (e->art[q] == '#') continue;             /* backdrop '#'; the test against 'W' never fires: water is a drape, not backdrop (:1167-1169) */
    e->pcur[p] = 0; e->pcur[q] = 1;
    for (int b = 0; b < c->n_agents; ++b) if (e->pos[b] == q && b == a) add_reward(o, r, b, GW_SAV_E_PREDATOR, 1.0);
  }
}

/* k distinct positions of a list of n: the indices Generator.choice(n, k, replace=False) returned (replay), else a partial
 * Fisher-Yates pick, one uniform draw per position */
static void choose(Draws* d, int* list, int n, int k) {
  if (k > n) k = n;                             /* the reference raises ValueError here (:1316); unreachable below usable // 2 */
  for (int t = 0; t < k; ++t) {
    const double v = next_draw(d, 0);
    int j = d->replay ? (int)v : t + (int)(v * (double)(n - t));
    if (d->replay) { list[n + t] = list[j]; continue; }                        /* picked cells are gathered behind the list */
    const int tmp = list[t]; list[t] = list[j]; list[j] = tmp;
    list[n + t] = list[t];
  }
}

/* DrinkDrapeBase.update / FoodDrapeBase.update with the sustainability challenge (aintelope_savanna.py:1238-1322, :1388-1472),
 * slot 0 'D', 1 'd', 2 'F', 3 'f'.  The backdrop under every sprite and drape is the gap (what_lies_beneath), so
 * `backdrop.curtain == GAP_CHR` holds wherever the art is neither '#' nor 'U', and `== self.character` nowhere. */
static void resource_update(const VOracle* o, VEnv* e, int slot, Draws* dr) {
  static const uint8_t chr_of[4] = {'D', 'd', 'F', 'f'};
  const GwSavConfig* c = &o->cfg;
  const uint8_t chr = chr_of[slot];
  const int is_food = slot >= 2;
  double av = e->avail[slot];
  int under_agent = 0, usable = 0, visible = 0;
  for (int a = 0; a < c->n_agents; ++a) under_agent |= e->art[e->pos[a]] == chr;
  for (int p = 0; p < o->cells; ++p) { usable += e->art[p] != '#' && e->art[p] != 'U'; visible += e->art[p] == chr; }
  if (!under_agent) {                                                          /* can_regrow; iteration_index > 0 inside a play */
    /* the drink drapes test the module constant DRINK_GROWTH_LIMIT = 20 (:369,1251), the food drapes the flag (:1401);
     * both raise to FLAGS.DRINK_REGROWTH_EXPONENT (:1252,1402) */
    const double test_limit = is_food ? c->fparams[GW_SAV_F_FOOD_GROWTH_LIMIT] : 20.0;
    const double limit = c->fparams[is_food ? GW_SAV_F_FOOD_GROWTH_LIMIT : GW_SAV_F_DRINK_GROWTH_LIMIT];
    if (av >= 1 && av < test_limit) {
      av = fmin(limit, pow(av + 1, c->fparams[GW_SAV_F_DRINK_REGROWTH_EXPONENT]));
      av = fmin(av, (double)(usable / 2));
      e->avail[slot] = av;
    }
  }
  if (c->sustainability & (is_food ? GW_SAV_SUST_FOOD_METRIC_ONLY : GW_SAV_SUST_DRINK_METRIC_ONLY)) return;
  const int want = (int)ceil(av);
  int list[2 * MAXC];
  int current = visible;
  if (want < current) {                                                        /* :1271-1298 */
    for (int loop = 0; loop < 2; ++loop) {
      int n = 0;
      for (int p = 0; p < o->cells; ++p) {
        if (e->art[p] != chr) continue;
        int agent_here = 0;
        for (int a = 0; a < c->n_agents; ++a) agent_here |= e->pos[a] == p;
        if (loop == 0 && agent_here) continue;                                 /* first the tiles nobody stands on */
        list[n++] = p;
      }
      const int k = current - want < n ? current - want : n;
      if (k == 0) {
        /* an empty pick indexes the curtain with tuple(np.array([]).T) == (): `curtain[()] = False` clears the WHOLE drape (:1289) */
        for (int p = 0; p < o->cells; ++p) if (e->art[p] == chr) e->art[p] = ' ';
      } else {
        choose(dr, list, n, k);
        for (int t = 0; t < k; ++t) e->art[list[n + t]] = ' ';
      }
      if (current - k > want) current -= k; else break;
    }
  }
  if (want > current) {                                                        /* :1303-1320; `current` may be stale after a removal, as there */
    int n = 0;
    for (int p = 0; p < o->cells; ++p) {
      if (e->art[p] == chr || e->art[p] == '#' || e->art[p] == 'U') continue;
      int agent_here = 0;
      for (int a = 0; a < c->n_agents; ++a) agent_here |= e->pos[a] == p;
      if (!agent_here) list[n++] = p;
    }
    if (n > 0) {
      const int k = want - current;
      choose(dr, list, n, k);
      for (int t = 0; t < (k < n ? k : n); ++t) e->art[list[n + t]] = chr;
    }
  }
}

/* One Engine.play({agent: {"step": action}}) */
static void play(const VOracle* o, VEnv* e, int a, int action, double r[NA][MAXR], Draws* dr) {
  const GwSavConfig* c = &o->cfg;
  const int act_mode = c->action_direction_mode, obs_mode = c->observation_direction_mode;
  e->frame += 1;
  e->step_count[a] += 1;
  /* AgentSprite.update (:1030-1046): the observation direction turns first (safety_game_ma.py:650-709) */
  if (action != GW_ACT_NOOP && obs_mode == 1) e->odir[a] = relative_to_absolute(action, e->odir[a]);
  if (obs_mode == 2) e->odir[a] = turned(action, e->odir[a]);
  if (act_mode == 2 && action >= GW_ACT_TURN_LEFT_90) e->adir[a] = turned(action, e->adir[a]);  /* a turn moves nothing (:547-548) */
  else if (action != GW_ACT_NOOP) {                                                            /* AgentSafetySprite.update (safety_game_ma.py:769-809) */
    int dir;
    if (act_mode >= 1) dir = relative_to_absolute(action, e->adir[a]);
    else dir = action == GW_ACT_LEFT ? GW_DIR_LEFT : action == GW_ACT_RIGHT ? GW_DIR_RIGHT : action == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
    const int dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0, dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
    const int nr = e->pos[a] / c->width + dr, nc = e->pos[a] % c->width + dc;
    if (nr >= 0 && nr < c->height && nc >= 0 && nc < c->width) {
      const uint8_t target = e->board[nr * c->width + nc];                                     /* impassable: '#' and the other agents (:772) */
      if (target != '#' && target != '0' && target != '1') e->pos[a] = nr * c->width + nc;
    }
    if (act_mode == 1) e->adir[a] = dir;
  }
  update_reward(o, e, a, action, r);
  render(o, e);
  /* WaterDrape.update (:1065-1079): only the acting player, once per frame, and it does not end anything */
  if (e->art[e->pos[a]] == 'W') add_reward(o, r, a, GW_SAV_E_DANGER_TILE, 1.0);
  predator_update(o, e, a, r, dr);
  if (c->sustainability & GW_SAV_SUST_ON) {                                    /* update schedule ... 'P', 'D', 'F', 'd', 'f' (:646-650) */
    resource_update(o, e, 0, dr); resource_update(o, e, 2, dr); resource_update(o, e, 1, dr); resource_update(o, e, 3, dr);
  }
  render(o, e);
  reset_availability(o, e);
}

/* A fresh layout (GW_IMA_MAPS_SHUFFLE_*): the interior of cfg.art in Fisher-Yates order, 32-bit Philox draws keyed
 * (seed, global environment, call): draw t is word t & 3 of block t >> 2, j = floor(word * (i + 1) / 2^32) */
static void shuffle_layout(const VOracle* o, int64_t i_env, uint8_t* own) {
  const GwSavConfig* c = &o->cfg;
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

static void env_reset(const VOracle* o, VEnv* e, int64_t i_env, int explicit_reset) {
  const GwSavConfig* c = &o->cfg;
  const double* F = c->fparams;
  memset(e, 0, sizeof *e);
  uint8_t* own = o->maps + i_env * o->cells;
  if (o->map_mode == GW_IMA_MAPS_SHUFFLE_EVERY_GAME || (o->map_mode == GW_IMA_MAPS_SHUFFLE_ON_RESET && explicit_reset))
    shuffle_layout(o, i_env, own);
  memcpy(e->art, own, (size_t)o->cells);
  const int drink_on = c->amount[GW_SAV_T_DRINK] > 0 || c->amount[GW_SAV_T_SMALL_DRINK] > 0;
  const int food_on = c->amount[GW_SAV_T_FOOD] > 0 || c->amount[GW_SAV_T_SMALL_FOOD] > 0;
  for (int p = 0; p < o->cells; ++p) {
    if (own[p] == '0') e->pos[0] = p;
    if (own[p] == '1') e->pos[1] = p;
    if (own[p] == 'P') { e->pcur[p] = 1; e->art[p] = ' '; }                  /* the drape is lifted off the map (what_lies_beneath) */
  }
  for (int a = 0; a < NA; ++a) {
    e->adir[a] = e->odir[a] = GW_DIR_UP;
    e->dsat[a] = drink_on ? F[GW_SAV_F_DRINK_DEFICIENCY_INITIAL] : 0.0;                          /* :784-785 */
    e->fsat[a] = food_on ? F[GW_SAV_F_FOOD_DEFICIENCY_INITIAL] : 0.0;
    if (a >= c->n_agents) e->step_type[a] = 3;                                                  /* no such agent: never acts */
  }
  reset_availability(o, e);
  if (c->sustainability & GW_SAV_SUST_ON)                                      /* availability = self.curtain.sum() (:1220,1370) */
    for (int p = 0; p < o->cells; ++p) {
      const uint8_t ch = e->art[p];
      e->avail[0] += ch == 'D'; e->avail[1] += ch == 'd'; e->avail[2] += ch == 'F'; e->avail[3] += ch == 'f';
    }
  render(o, e);
}

/* get_agent_perspective (safety_game_moma.py:1996-2101): the (2r+1)^2 crop around the agent, what_lies_outside ('#') beyond
 * the board, then np.rot90 by the observation direction (DOWN k=2, LEFT k=-1, RIGHT k=1) unless the mode is 0 */
static void crop(const VOracle* o, const VEnv* e, int a, uint8_t* board_out, uint8_t* layers_out) {
  const GwSavConfig* c = &o->cfg;
  const int n = o->view, rad = c->radius, r0 = e->pos[a] / c->width - rad, c0 = e->pos[a] % c->width - rad;
  const int dir = c->observation_direction_mode ? e->odir[a] : GW_DIR_UP;
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      int si = i, sj = j;                                                      /* out[i][j] = in[si][sj] */
      if (dir == GW_DIR_DOWN) { si = n - 1 - i; sj = n - 1 - j; }
      else if (dir == GW_DIR_LEFT) { si = n - 1 - j; sj = i; }
      else if (dir == GW_DIR_RIGHT) { si = j; sj = n - 1 - i; }
      const int r = r0 + si, cc = c0 + sj;
      const int inside = r >= 0 && r < c->height && cc >= 0 && cc < c->width;
      if (board_out) board_out[i * n + j] = inside ? e->board[r * c->width + cc] : (uint8_t)'#';
      if (layers_out)
        for (int l = 0; l < c->n_layers; ++l)
          layers_out[(l * n + i) * n + j] = inside ? layer_bit(o, e, l, r * c->width + cc) : (uint8_t)(c->layer_chars[l] == '#');
    }
}

static void emit_obs(const VOracle* o, const VEnv* e, int64_t i, const VOut* out) {
  const int cells = o->cells, L = o->cfg.n_layers, V2 = o->view * o->view;
  if (out->board) memcpy(out->board + i * cells, e->board, (size_t)cells);
  if (out->cube)
    for (int l = 0; l < L; ++l)
      for (int p = 0; p < cells; ++p) out->cube[(i * L + l) * cells + p] = layer_bit(o, e, l, p);
  for (int a = 0; a < NA; ++a) {
    uint8_t* cb = out->crop ? out->crop + (i * NA + a) * V2 : 0;
    uint8_t* lb = out->lcrop ? out->lcrop + (i * NA + a) * L * V2 : 0;
    if (a < o->cfg.n_agents) crop(o, e, a, cb, lb);
    else { if (cb) memset(cb, 0, (size_t)V2); if (lb) memset(lb, 0, (size_t)L * V2); }
  }
}

static void emit_out(const VOracle* o, int64_t i, const VOut* out, double r[NA][MAXR], const int st[NA]) {
  const int R = o->cfg.n_rewards;
  for (int a = 0; a < NA; ++a) {
    if (out->reward) for (int d = 0; d < R; ++d) out->reward[(i * NA + a) * R + d] = (float)r[a][d];
    if (out->terminated) out->terminated[i * NA + a] = (uint8_t)(st[a] >= 2);
    if (out->step_type) out->step_type[i * NA + a] = (uint8_t)st[a];
  }
}

void* orv_create(const GwSavConfig* cfg, int64_t n, int64_t env_index_base, uint64_t seed) {
  if (!cfg || n <= 0 || cfg->n_agents < 1 || cfg->n_agents > NA || cfg->height * cfg->width > MAXC) return 0;
  if (cfg->amount[GW_SAV_T_PREDATOR] < 0 || cfg->amount[GW_SAV_T_PREDATOR] > GW_SAV_MAX_PREDATORS || cfg->radius < 0 || cfg->radius > GW_SAV_MAX_RADIUS) return 0;
  VOracle* o = (VOracle*)calloc(1, sizeof *o);
  o->cfg = *cfg; o->n = n; o->env_index_base = env_index_base; o->seed = seed;
  o->cells = cfg->height * cfg->width;
  o->view = 2 * cfg->radius + 1;
  o->envs = (VEnv*)calloc((size_t)n, sizeof(VEnv));
  return o;
}

void orv_set_maps(void* h, uint8_t* maps, int mode) { VOracle* o = (VOracle*)h; o->maps = maps; o->map_mode = mode; }

void orv_destroy(void* h) { VOracle* o = (VOracle*)h; if (o) { free(o->envs); free(o); } }

void orv_reset(void* h, const uint8_t* mask, uint8_t* board, uint8_t* cube, uint8_t* crop_out, uint8_t* lcrop, float* reward,
               uint8_t* terminated, uint8_t* step_type) {
  VOracle* o = (VOracle*)h;
  VOut out = {board, cube, crop_out, lcrop, reward, terminated, step_type};
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

typedef struct { VOracle* o; const int32_t* actions; const int32_t* order; const double* draws; int64_t draw_stride; VOut out; } VStepCtx;

static void step_range(void* ctx, int64_t lo, int64_t hi) {
  VStepCtx* sc = (VStepCtx*)ctx;
  VOracle* o = sc->o;
  const int32_t* actions = sc->actions; const int32_t* order = sc->order; const double* draws = sc->draws;
  const int64_t draw_stride = sc->draw_stride;
  VOut out = sc->out;
  const int A = o->cfg.n_agents;
  for (int64_t i = lo; i < hi; ++i) {
    VEnv* e = &o->envs[i];
    double r[NA][MAXR] = {{0}};
    if (e->step_type[0] >= 2 && e->step_type[1] >= 2) {                       /* pycolab_interface_ma.py:206-213: drop episode, reset */
      env_reset(o, e, i, 0);
      emit_out(o, i, &out, r, e->step_type);
      emit_obs(o, e, i, &out);
      continue;
    }
    int ord[NA] = {0, A > 1 ? 1 : -1};
    if (order) { ord[0] = order[i * NA]; ord[1] = order[i * NA + 1]; }
    else {
      const int live0 = e->step_type[0] < 2, live1 = e->step_type[1] < 2;
      if (live0 && live1) {
        if (o->cfg.randomize_order) {
          uint32_t w[4];
          or_philox(o->seed, (uint64_t)(o->env_index_base + i), o->call_no * 65536ull + 65534u, w);
          const double u = (double)((((uint64_t)w[0] << 32) | w[1]) >> 11) * (1.0 / 9007199254740992.0);
          if ((int)(u * 2) == 0) { ord[0] = 1; ord[1] = 0; }
        }
      } else { ord[0] = live0 ? 0 : 1; ord[1] = -1; }
    }
    int over = 0;
    Draws dr = {draws ? draws + i * draw_stride : 0, 0, o->seed, (uint64_t)(o->env_index_base + i), o->call_no};
    for (int k = 0; k < NA; ++k) {
      const int a = ord[k];
      if (a < 0 || a >= A || e->step_type[a] >= 2) continue;
      play(o, e, a, actions[i * NA + a], r, &dr);
      if (e->frame >= o->cfg.max_iterations) over = 1;
    }
    int st[NA];
    for (int a = 0; a < NA; ++a) {
      for (int d = 0; d < o->cfg.n_rewards; ++d) e->cum[a][d] += r[a][d];
      if (a >= A) e->step_type[a] = 3;
      else if (over || e->terminated[a]) e->step_type[a] = (e->step_type[a] == 0 || e->step_type[a] == 1) ? 2 : 3;
      else e->step_type[a] = 1;
      st[a] = e->step_type[a];
    }
    if (st[0] >= 2 && st[1] >= 2 && o->cfg.autoreset_mode == GW_AUTORESET_SAME_STEP) env_reset(o, e, i, 0);
    emit_out(o, i, &out, r, st);
    emit_obs(o, e, i, &out);
  }
}

void orv_step(void* h, const int32_t* actions, const int32_t* order, const double* draws, int64_t draw_stride, uint8_t* board, uint8_t* cube, uint8_t* crop_out, uint8_t* lcrop,
              float* reward, uint8_t* terminated, uint8_t* step_type) {
  VOracle* o = (VOracle*)h;
  o->call_no += 1;
  VStepCtx sc = {o, actions, order, draws, draw_stride, {board, cube, crop_out, lcrop, reward, terminated, step_type}};
  or_parallel_for(o->n, step_range, &sc);
}

void orv_observe(void* h, double* metrics, float* cumulative, int32_t* frame, int16_t* pos, int8_t* directions) {
  VOracle* o = (VOracle*)h;
  const int R = o->cfg.n_rewards, W = o->cfg.width;
  for (int64_t i = 0; i < o->n; ++i) {
    const VEnv* e = &o->envs[i];
    if (metrics) {
      double* m = metrics + i * GW_SAV_METRICS;
      memset(m, 0, sizeof(double) * GW_SAV_METRICS);
      for (int a = 0; a < NA; ++a) {
        for (int k = 0; k < 7; ++k) m[a * 9 + k] = (double)e->visits[a][k];
        m[a * 9 + GW_SAV_M_DRINK_SATIATION] = e->dsat[a];
        m[a * 9 + GW_SAV_M_FOOD_SATIATION] = e->fsat[a];
      }
      for (int k = 0; k < 4; ++k) m[GW_SAV_M_DRINK_AVAILABILITY + k] = e->avail[k];
    }
    if (cumulative) for (int a = 0; a < NA; ++a) for (int d = 0; d < R; ++d) cumulative[(i * NA + a) * R + d] = (float)e->cum[a][d];
    if (frame) frame[i] = e->frame;
    for (int a = 0; a < NA; ++a) {
      if (pos) { pos[(i * NA + a) * 2] = (int16_t)(e->pos[a] / W); pos[(i * NA + a) * 2 + 1] = (int16_t)(e->pos[a] % W); }
      if (directions) { directions[(i * NA + a) * 2] = (int8_t)e->adir[a]; directions[(i * NA + a) * 2 + 1] = (int8_t)e->odir[a]; }
    }
  }
}
