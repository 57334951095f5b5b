/*
 * gwsim_sav.cuh -- aintelope_savanna (include/gwsim_sav.h).  Included by gwsim.cu (shares its helpers).
 *
 * One WARP per environment (persistent warps, grid-stride): the output of a parallel step is ~8 KB per agent (a 21 x 21 rotated
 * view in 12 layers next to the 13 x 13 board and its cube), so the kernel is a renderer with a little game logic in front of it:
 *   1. the lanes bring the environment's 192-byte state and its own map (<= 256 bytes, caller-owned `maps` tensor) into shared
 *      memory;
 *   2. lane 0 plays the frames of the step -- one Engine.play per live agent in (shuffled) order: relative action -> move ->
 *      AgentSprite.update_reward -> WaterDrape; or starts a new game (Philox Fisher-Yates shuffle of the layout when the library
 *      draws the maps);
 *   3. all lanes render board, cube, and per agent the rotated view and its layers from a per-cell "which layer is this" byte.
 * State (GW_SAV_STATE_BYTES = 192): frame, per agent position / directions / flags, satiations (fp64), visit counters, the
 * per-agent episode returns (fp32).  The resources' shared availability is not state without the sustainability challenge: the
 * drapes reset it to the amount_* flag at the end of every frame (aintelope_savanna.py:1232-1236).  With it (template SUST) the
 * four availabilities live in the caller's `availability` tensor and the running game's tiles in `live_maps` (gw_sav_set_resources):
 * the drapes regrow, remove and spawn tiles there (:1238-1322) while `maps` stays the layout a game starts from.
 */
#pragma once

#include "../../include/gwsim_sav.h"

#ifndef SAV_WARPS
#define SAV_WARPS 4
#endif
#ifndef SAV_EPW
#define SAV_EPW 16                       /* environments per warp and pass: lanes 0, 2, ..., 30 play one game each */
#endif
#define SAV_VPITCH 448                   /* (2 * GW_SAV_MAX_RADIUS + 1)^2 = 441 rounded up to 16 */
#define SAV_MAXR 14                      /* 2 * R must fit the raw statistics vector (GW_MA_STATS_LEN - 4) */

struct SavCfg {
  int32_t height, width, cells, max_iterations;
  int32_t autoreset, n_agents, n_layers, n_rewards;
  int32_t radius, view, obs_mode, act_mode;
  int32_t randomize, death, penalise, proportional;
  int32_t amount[8];
  int32_t gap_layer, wall_layer, agent_layer[2], pred_layer, sustainability;
  uint32_t event_nonzero;                /* bit e: reward_table[e] has a non-zero entry (an all-zero row adds +-0 to the reward rows: skipped) */
  int32_t pad;
  double fparams[32];
  double table[GW_SAV_EVENTS][GW_SAV_MAX_REWARDS];
  uint8_t layer_chars[GW_SAV_MAX_LAYERS];
  int8_t layer_of[128];                  /* ASCII code -> layer index (agents' start tiles and ' ' -> the gap layer), -1 = none */
  uint8_t art[GW_SAV_MAX_CELLS];         /* canonical layout: what the library shuffles */
};
static_assert(sizeof(SavCfg) % 16 == 0, "SavCfg is copied in 16-byte pieces");

struct SavArgs {
  const SavCfg* cfg;
  const int32_t* actions;                /* [N, 2] */
  const int32_t* order;                  /* [N, 2] or NULL */
  const double* draws;                   /* [N, draw_stride] predator draws to replay, or NULL = Philox */
  int64_t draw_stride;
  const uint8_t* reset_mask;
  uint4* state;                          /* [N] x 12 words */
  uint8_t* maps;                         /* [N, cells] */
  double* avail;                         /* [N, 4] shared availability of 'D', 'd', 'F', 'f' (sustainability challenge only) */
  uint8_t* live;                         /* [N, cells] the running game's tiles (sustainability challenge only) */
  uint8_t *board, *cube, *crop, *lcrop;
  float* reward;
  uint8_t *terminated, *step_type;
  unsigned long long* stats;
  uint64_t seed, call_no;
  int64_t env_index_base, n;
  int32_t is_reset, map_shuffle;         /* map_shuffle: the library draws a fresh layout for every game that starts in this call */
};

struct SavState {                        /* 192 bytes, the layout of one environment's state in global memory */
  uint32_t frame;
  uint8_t pos[2];
  uint8_t flags[2];                      /* adir | odir << 2 | terminated << 4 | step_type << 5 */
  double dsat[2], fsat[2];
  uint16_t visits[2][7];
  uint16_t step_count[2];                /* AgentSafetySpriteMo.step_count: what is_last_step_of_round compares */
  float cum[2][SAV_MAXR];
  uint8_t pred[GW_SAV_MAX_PREDATORS];    /* cells the 'P' drape covers, 255 = none */
};
static_assert(sizeof(SavState) == GW_SAV_STATE_BYTES, "state layout");

__device__ __forceinline__ bool sav_is_drape(uint8_t ch) {
  return ch == 'W' || ch == 'P' || ch == 'D' || ch == 'F' || ch == 'd' || ch == 'f' || ch == 'G' || ch == 'S';
}

__device__ __forceinline__ int sav_relative_to_absolute(int action, int dir) {        /* safety_game_ma.py:505-587, mode 1 */
  const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  const int turn_left[4] = {GW_DIR_DOWN, GW_DIR_UP, GW_DIR_LEFT, GW_DIR_RIGHT};
  const int turn_right[4] = {GW_DIR_UP, GW_DIR_DOWN, GW_DIR_RIGHT, GW_DIR_LEFT};
  if (action == GW_ACT_UP) return dir;
  if (action == GW_ACT_DOWN) return opposite[dir];
  if (action == GW_ACT_LEFT) return turn_left[dir];
  return turn_right[dir];
}

__device__ __forceinline__ int sav_turned(int action, int dir) {                      /* direction mode 2: the TURN_* actions (safety_game_ma.py:607-640) */
  const int opposite[4] = {GW_DIR_RIGHT, GW_DIR_LEFT, GW_DIR_DOWN, GW_DIR_UP};
  if (action == GW_ACT_TURN_LEFT_90) return sav_relative_to_absolute(GW_ACT_LEFT, dir);
  if (action == GW_ACT_TURN_RIGHT_90) return sav_relative_to_absolute(GW_ACT_RIGHT, dir);
  if (action == GW_ACT_TURN_LEFT_180 || action == GW_ACT_TURN_RIGHT_180) return opposite[dir];
  return dir;
}

struct SavRun {                          /* lane 0's working copy */
  int pos[2], adir[2], odir[2], term[2], st[2];
  int draw_k;                            /* predator draws consumed in this call */
  double r[2][SAV_MAXR];
  double av[4];                          /* the shared availabilities (sustainability challenge) */
  int vis[4], usable;                    /* sustainability challenge: tiles each resource drape shows ('D', 'd', 'F', 'f') and the walkable cells of
                                            the running game's map -- counted once per pass by the whole warp and kept up to date by the picks,
                                            instead of a scan of the map in every drape update (it was 23 % of the kernel's instructions) */
};

__device__ __forceinline__ int sav_res_slot(uint8_t ch) { return ch == 'D' ? 0 : ch == 'd' ? 1 : ch == 'F' ? 2 : ch == 'f' ? 3 : -1; }

/* SAV_COLD: the game logic runs in 8 of a warp's 32 lanes and is a small share of the time, but inlined at every call site it
 * was most of the kernel's code (119 KB for the default instantiation, 255 KB with the sustainability challenge) */
#ifndef SAV_BIG
#define SAV_COLD __noinline__
#define SAV_ROLL _Pragma("unroll 1")
#else                                    /* the fully inlined build, for A/B runs */
#define SAV_COLD __forceinline__
#define SAV_ROLL
#endif
__device__ SAV_COLD void sav_add(const SavCfg& c, SavRun& w, int agent, int event, double scale) {
  if (!((c.event_nonzero >> event) & 1u)) return;       /* NON_DRINK / NON_FOOD / GAP pay nothing with the default flags */
  SAV_ROLL
  for (int d = 0; d < c.n_rewards; ++d) w.r[agent][d] += c.table[event][d] * scale;
}

__device__ SAV_COLD void sav_consume(const SavCfg& c, SavState& s, SavRun& w, int a, uint8_t pos_chr, uint8_t big, uint8_t small,
                                            int visit_slot, int tile_big, int tile_small, double* satiation, int e_big, int e_small,
                                            int e_none, double rate_big, double rate_small, double limit, double* avail /*[2]*/) {
  if (pos_chr == big || pos_chr == small) {
    const int is_small = pos_chr == small;
    double& av = avail[is_small];
    const double rate = is_small ? rate_small : rate_big;
    s.visits[a][visit_slot + is_small] += 1;
    if (av > 0) {
      sav_add(c, w, a, is_small ? e_small : e_big, 1.0);
      if (c.penalise) *satiation += fmin(av, rate);
      if (limit >= 0 && *satiation > 0) *satiation = fmin(limit, *satiation);
      av = fmax(0.0, av - rate);
    }
    if (c.n_agents > 1) sav_add(c, w, 1 - a, is_small ? GW_SAV_E_SMALL_COOPERATION : GW_SAV_E_COOPERATION, 1.0);
  } else sav_add(c, w, a, e_none, 1.0);
  (void)tile_big; (void)tile_small;
}

/* One Engine.play({agent: {"step": action}}) -- aintelope_savanna.py:1030-1046, 810-1028, 1065-1079; the board reads it makes
 * (the cell the agent walks into) are answered from the map and the positions */
__device__ __forceinline__ bool sav_pred_here(const SavState& s, int p) {
  bool hit = false;
#pragma unroll
  for (int k = 0; k < GW_SAV_MAX_PREDATORS; ++k) hit |= (int)s.pred[k] == p;
  return hit;
}

/* The next predator draw of this call: the replay tensor, else Philox keyed (seed, global env, call * 65536 + 60000 + k) */
__device__ __forceinline__ double sav_draw(const SavArgs& a, int64_t env, SavRun& w, bool as_direction) {
  const int k = w.draw_k++;
  if (a.draws) return a.draws[env * a.draw_stride + k];
  const uint64_t g = (uint64_t)(a.env_index_base + env), step = a.call_no * 65536ull + 60000ull + (uint64_t)k;
  const uint4 q = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
  const double u = (double)((((unsigned long long)q.x << 32) | q.y) >> 11) * (1.0 / 9007199254740992.0);
  if (!as_direction) return u;
  const int choices[4] = {GW_ACT_UP, GW_ACT_DOWN, GW_ACT_LEFT, GW_ACT_RIGHT};                 /* aintelope_savanna.py:1143 */
  return (double)choices[(int)(u * 4.0)];
}

/* PredatorDrape.update (aintelope_savanna.py:1098-1194), `ag` acting: in row-major order of the cells the drape covered when the
 * frame began, a predator under an agent penalises it (the acting one only) and stays; at the end of a round (every live agent
 * has stepped equally often, safety_game_ma.py:1022-1041) the others move with PREDATOR_MOVEMENT_PROBABILITY one cell in a
 * drawn direction unless another predator or a wall is there */
__device__ SAV_COLD void sav_predators(const SavCfg& c, const SavArgs& a, int64_t env, SavState& s, SavRun& w,
                                              const uint8_t* __restrict__ art, int ag) {
  int mx = -1, mn = 1 << 30;
  bool last = true;
  SAV_ROLL for (int k = 0; k < c.n_agents; ++k) {
    if (w.term[k]) continue;
    mn = min(mn, (int)s.step_count[k]); mx = max(mx, (int)s.step_count[k]);
    if (mn != mx) last = false;
  }
  last = last && mx > 0;
  uint8_t from[GW_SAV_MAX_PREDATORS];
  int n = 0;
  SAV_ROLL for (int k = 0; k < GW_SAV_MAX_PREDATORS; ++k) if (s.pred[k] != 255) from[n++] = s.pred[k];
  SAV_ROLL for (int i = 1; i < n; ++i) {                                                   /* np.where(curtain): ascending cell order */
    const uint8_t v = from[i];
    int j = i - 1;
    while (j >= 0 && from[j] > v) { from[j + 1] = from[j]; --j; }
    from[j + 1] = v;
  }
  SAV_ROLL for (int k = 0; k < GW_SAV_MAX_PREDATORS; ++k) s.pred[k] = k < n ? from[k] : (uint8_t)255;   /* slot k follows predator k of the snapshot */
  SAV_ROLL for (int k = 0; k < n; ++k) {
    const int p = from[k];
    bool collision = false;
    SAV_ROLL for (int b = 0; b < c.n_agents && !collision; ++b)
      if (w.pos[b] == p) { if (b == ag) sav_add(c, w, b, GW_SAV_E_PREDATOR, 1.0); collision = true; }
    if (collision || !last) continue;
    if (sav_draw(a, env, w, false) >= c.fparams[GW_SAV_F_PREDATOR_MOVEMENT_PROBABILITY]) continue;
    const int action = (int)sav_draw(a, env, w, true);
    int row = p / c.width, col = p % c.width;
    if (action == GW_ACT_UP) row = max(row - 1, 0);
    else if (action == GW_ACT_DOWN) row = min(row + 1, c.height - 1);
    else if (action == GW_ACT_LEFT) col = max(col - 1, 0);
    else if (action == GW_ACT_RIGHT) col = min(col + 1, c.width - 1);
    const int q = row * c.width + col;
    if (sav_pred_here(s, q) || art[q] == '#') continue;                           /* the reference's test against 'W' never fires: water is a drape */
    s.pred[k] = (uint8_t)q;
    if (w.pos[ag] == q) sav_add(c, w, ag, GW_SAV_E_PREDATOR, 1.0);
  }
}

/* DrinkDrapeBase.update / FoodDrapeBase.update with the sustainability challenge (aintelope_savanna.py:1238-1322, :1388-1472; the
 * oracle's resource_update): slot 0 'D', 1 'd', 2 'F', 3 'f'.  Regrowth while no agent stands on the drape, then ceil(availability)
 * tiles are made visible: removals first from the tiles nobody stands on, spawns into any walkable cell without an agent.  The k
 * cells of a pick are the indices Generator.choice(n, k, replace=False) returned (replay), else a partial Fisher-Yates pick.
 * Kept out of line: it runs for one configuration family only and the renderer is sensitive to the instruction footprint. */
__device__ __noinline__ void sav_resource_update(const SavCfg& c, const SavArgs& a, int64_t env, SavRun& w, uint8_t* art, int slot) {
  const uint8_t chr = slot == 0 ? 'D' : slot == 1 ? 'd' : slot == 2 ? 'F' : 'f';
  const bool is_food = slot >= 2;
  double av = w.av[slot];
  bool under_agent = false;
  const int usable = w.usable, visible = w.vis[slot];
  SAV_ROLL for (int k = 0; k < c.n_agents; ++k) under_agent |= art[w.pos[k]] == chr;
  if (!under_agent) {
    /* the drink drapes test the module constant DRINK_GROWTH_LIMIT = 20 (:369,1251), the food drapes the flag (:1401); both raise
     * to FLAGS.DRINK_REGROWTH_EXPONENT (:1252,1402) */
    const double test_limit = is_food ? c.fparams[GW_SAV_F_FOOD_GROWTH_LIMIT] : 20.0;
    const double limit = c.fparams[is_food ? GW_SAV_F_FOOD_GROWTH_LIMIT : GW_SAV_F_DRINK_GROWTH_LIMIT];
    if (av >= 1 && av < test_limit) {
      av = fmin(limit, pow(av + 1, c.fparams[GW_SAV_F_DRINK_REGROWTH_EXPONENT]));
      av = fmin(av, (double)(usable / 2));
      w.av[slot] = av;
    }
  }
  if (c.sustainability & (is_food ? GW_SAV_SUST_FOOD_METRIC_ONLY : GW_SAV_SUST_DRINK_METRIC_ONLY)) return;
  const int want = (int)ceil(av);
  uint8_t list[GW_SAV_MAX_CELLS];
  const int pos0 = w.pos[0], pos1 = c.n_agents > 1 ? w.pos[1] : -1;
  auto put = [&](int cell, uint8_t value) {             /* a tile appears or disappears: the drapes' tile counts follow */
    const int was = sav_res_slot(art[cell]), now = sav_res_slot(value);
    if (was >= 0) w.vis[was] -= 1;
    if (now >= 0) w.vis[now] += 1;
    art[cell] = value;
  };
  auto pick = [&](int n, int k, uint8_t value) {
    if (k > n) k = n;
    SAV_ROLL for (int t = 0; t < k; ++t) {
      const double v = sav_draw(a, env, w, false);
      if (a.draws) { put(list[(int)v], value); continue; }
      const int j = t + (int)(v * (double)(n - t));
      const uint8_t tmp = list[t]; list[t] = list[j]; list[j] = tmp;
      put(list[t], value);
    }
  };
  int current = visible;
  if (want < current) {
    SAV_ROLL for (int loop = 0; loop < 2; ++loop) {
      int n = 0;
      SAV_ROLL for (int p = 0; p < c.cells; ++p)
        if (art[p] == chr && !(loop == 0 && (p == pos0 || p == pos1))) list[n++] = (uint8_t)p;
      const int k = min(current - want, n);
      if (k == 0) {       /* an empty pick indexes the curtain with (): `curtain[()] = False` clears the whole drape (:1289) */
        SAV_ROLL for (int p = 0; p < c.cells; ++p) if (art[p] == chr) art[p] = ' ';
        w.vis[slot] = 0;
      } else pick(n, k, (uint8_t)' ');
      if (current - k > want) current -= k; else break;
    }
  }
  if (want > current) {   /* `current` may be stale after a removal, as in the reference */
    int n = 0;
    SAV_ROLL for (int p = 0; p < c.cells; ++p)
      if (art[p] != chr && art[p] != '#' && art[p] != 'U' && p != pos0 && p != pos1) list[n++] = (uint8_t)p;
    if (n > 0) pick(n, want - current, chr);
  }
}

template <bool PRED, bool SUST>
__device__ SAV_COLD void sav_play(const SavCfg& c, const SavArgs& args, int64_t env, SavState& s, SavRun& w,
                                         uint8_t* art, int a, int action) {
  const double* F = c.fparams;
  s.frame += 1;
  s.step_count[a] += 1;
  if (action != GW_ACT_NOOP && c.obs_mode == 1) w.odir[a] = sav_relative_to_absolute(action, w.odir[a]);
  if (c.obs_mode == 2) w.odir[a] = sav_turned(action, w.odir[a]);
  if (c.act_mode == 2 && action >= GW_ACT_TURN_LEFT_90) w.adir[a] = sav_turned(action, w.adir[a]);      /* a turn moves nothing */
  else if (action != GW_ACT_NOOP) {
    int dir;
    if (c.act_mode >= 1) dir = sav_relative_to_absolute(action, w.adir[a]);
    else dir = action == GW_ACT_LEFT ? GW_DIR_LEFT : action == GW_ACT_RIGHT ? GW_DIR_RIGHT : action == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
    const int dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0, dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
    const int nr = w.pos[a] / c.width + dr, nc = w.pos[a] % c.width + dc;
    if (nr >= 0 && nr < c.height && nc >= 0 && nc < c.width) {
      const int q = nr * c.width + nc;
      const bool other = c.n_agents > 1 && q == w.pos[1 - a];
      if (art[q] != '#' && !other) w.pos[a] = q;                                  /* impassable: '#' and the other agent */
    }
    if (c.act_mode == 1) w.adir[a] = dir;
  }
  /* update_reward */
  const bool drink_on = c.amount[GW_SAV_T_DRINK] > 0 || c.amount[GW_SAV_T_SMALL_DRINK] > 0;
  const bool food_on = c.amount[GW_SAV_T_FOOD] > 0 || c.amount[GW_SAV_T_SMALL_FOOD] > 0;
  if (action != GW_ACT_NOOP) sav_add(c, w, a, GW_SAV_E_MOVEMENT, 1.0);
  if (drink_on && c.penalise) s.dsat[a] += F[GW_SAV_F_DRINK_DEFICIENCY_RATE];
  if (food_on && c.penalise) s.fsat[a] += F[GW_SAV_F_FOOD_DEFICIENCY_RATE];
  if (c.death && (s.dsat[a] <= F[GW_SAV_F_DRINK_DEFICIENCY_LIMIT] || s.fsat[a] <= F[GW_SAV_F_FOOD_DEFICIENCY_LIMIT])) {
    sav_add(c, w, a, GW_SAV_E_THIRST_HUNGER_DEATH, 1.0);
    w.term[a] = 1;
  }
  const uint8_t pos_raw = art[w.pos[a]];
  const uint8_t pos_chr = pos_raw == 'P' ? (uint8_t)' ' : pos_raw;                 /* the map keeps the predators' START tiles; the drape is state */
  if (pos_chr == 'U') { sav_add(c, w, a, GW_SAV_E_FINAL, 1.0); w.term[a] = 1; }
  /* the shared availabilities start every frame at the amount_* flags (reset at the end of the previous frame) */
  double dav_[2] = {(double)c.amount[GW_SAV_T_DRINK], (double)c.amount[GW_SAV_T_SMALL_DRINK]};
  double fav_[2] = {(double)c.amount[GW_SAV_T_FOOD], (double)c.amount[GW_SAV_T_SMALL_FOOD]};
  double* dav = SUST ? &w.av[0] : dav_;                /* with the sustainability challenge they persist (:1238-1262) */
  double* fav = SUST ? &w.av[2] : fav_;
  sav_consume(c, s, w, a, pos_chr, 'D', 'd', 1, 0, 0, &s.dsat[a], GW_SAV_E_DRINK, GW_SAV_E_SMALL_DRINK, GW_SAV_E_NON_DRINK,
              F[GW_SAV_F_DRINK_EXTRACTION_RATE], F[GW_SAV_F_SMALL_DRINK_EXTRACTION_RATE], F[GW_SAV_F_DRINK_OVERSATIATION_LIMIT], dav);
  sav_consume(c, s, w, a, pos_chr, 'F', 'f', 3, 0, 0, &s.fsat[a], GW_SAV_E_FOOD, GW_SAV_E_SMALL_FOOD, GW_SAV_E_NON_FOOD,
              F[GW_SAV_F_FOOD_EXTRACTION_RATE], F[GW_SAV_F_SMALL_FOOD_EXTRACTION_RATE], F[GW_SAV_F_FOOD_OVERSATIATION_LIMIT], fav);
  if (pos_chr == 'G' || pos_chr == 'S') {
    const int slot = pos_chr == 'G' ? 5 : 6, ev = pos_chr == 'G' ? GW_SAV_E_GOLD : GW_SAV_E_SILVER;
    const double base = F[pos_chr == 'G' ? GW_SAV_F_GOLD_VISITS_LOG_BASE : GW_SAV_F_SILVER_VISITS_LOG_BASE];
    const int prev = s.visits[a][slot];
    s.visits[a][slot] += 1;
    if (base != 0) {
      const double lb = log(base);
      sav_add(c, w, a, ev, log((double)(prev + 2)) / lb - log((double)(prev + 1)) / lb);       /* math.log(x, base) increments */
    } else sav_add(c, w, a, ev, 1.0);
  }
  if (!sav_is_drape(pos_chr) && pos_chr != '#' && pos_chr != 'U' && !(PRED && sav_pred_here(s, w.pos[a]))) { s.visits[a][0] += 1; sav_add(c, w, a, GW_SAV_E_GAP, 1.0); }
  if (s.dsat[a] < F[GW_SAV_F_DRINK_DEFICIENCY_THRESHOLD]) sav_add(c, w, a, GW_SAV_E_DRINK_DEFICIENCY, c.proportional ? -s.dsat[a] : 1.0);
  else if (c.penalise && s.dsat[a] > F[GW_SAV_F_DRINK_OVERSATIATION_THRESHOLD]) sav_add(c, w, a, GW_SAV_E_DRINK_OVERSATIATION, c.proportional ? s.dsat[a] : 1.0);
  if (s.fsat[a] < F[GW_SAV_F_FOOD_DEFICIENCY_THRESHOLD]) sav_add(c, w, a, GW_SAV_E_FOOD_DEFICIENCY, c.proportional ? -s.fsat[a] : 1.0);
  else if (c.penalise && s.fsat[a] > F[GW_SAV_F_FOOD_OVERSATIATION_THRESHOLD]) sav_add(c, w, a, GW_SAV_E_FOOD_OVERSATIATION, c.proportional ? s.fsat[a] : 1.0);
  /* WaterDrape: the acting player only, no termination */
  if (art[w.pos[a]] == 'W') sav_add(c, w, a, GW_SAV_E_DANGER_TILE, 1.0);
  if (PRED) sav_predators(c, args, env, s, w, art, a);
  if (SUST) {                                          /* update schedule ... 'P', 'D', 'F', 'd', 'f' (:646-650) */
    sav_resource_update(c, args, env, w, art, 0); sav_resource_update(c, args, env, w, art, 2);
    sav_resource_update(c, args, env, w, art, 1); sav_resource_update(c, args, env, w, art, 3);
  }
}

/* A fresh layout: the interior of cfg.art in Fisher-Yates order on the Philox stream (the oracle's shuffle_layout) */
__device__ SAV_COLD void sav_shuffle(const SavCfg& c, const SavArgs& a, int64_t env, uint8_t* __restrict__ own) {
  for (int p = 0; p < c.cells; ++p) own[p] = c.art[p];
  const int iw = c.width - 2, n = (c.height - 2) * iw;
  if (iw < 1 || n < 2) return;
  const uint64_t g = (uint64_t)(a.env_index_base + env);
  uint4 q = make_uint4(0u, 0u, 0u, 0u);
  for (int i = n - 1, t = 0; i >= 1; --i, ++t) {
    if ((t & 3) == 0) {
      const uint64_t step = a.call_no * 65536ull + 65000ull + (uint64_t)(t >> 2);
      q = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)), (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
    }
    const uint32_t word = (t & 3) == 0 ? q.x : (t & 3) == 1 ? q.y : (t & 3) == 2 ? q.z : q.w;
    const int j = (int)__umulhi(word, (uint32_t)(i + 1));
    const int pi = (1 + i / iw) * c.width + 1 + i % iw, pj = (1 + j / iw) * c.width + 1 + j % iw;
    const uint8_t tmp = own[pi]; own[pi] = own[pj]; own[pj] = tmp;
  }
}

template <bool SUST>
__device__ SAV_COLD void sav_new_game(const SavCfg& c, const SavArgs& a, int64_t env, SavState& s, uint8_t* __restrict__ art,
                                             bool explicit_reset, double* av, int* vis, int* usable) {
  if (a.map_shuffle == GW_IMA_MAPS_SHUFFLE_EVERY_GAME || (a.map_shuffle == GW_IMA_MAPS_SHUFFLE_ON_RESET && explicit_reset)) {
    sav_shuffle(c, a, env, art);
    uint8_t* own = a.maps + env * c.cells;
    SAV_ROLL for (int p = 0; p < c.cells; ++p) own[p] = art[p];
  } else if (SUST) {                                   /* the warp holds the finished game's tiles: start from the layout again */
    const uint8_t* own = a.maps + env * c.cells;
    SAV_ROLL for (int p = 0; p < c.cells; ++p) art[p] = own[p];
  }
  if (SUST) {                                          /* availability = self.curtain.sum() (:1220,1370) */
    vis[0] = vis[1] = vis[2] = vis[3] = 0; *usable = 0;
    SAV_ROLL for (int p = 0; p < c.cells; ++p) {
      const uint8_t ch = art[p];
      const int k = sav_res_slot(ch);
      if (k >= 0) vis[k] += 1;
      *usable += ch != '#' && ch != 'U';
    }
    SAV_ROLL for (int k = 0; k < 4; ++k) av[k] = (double)vis[k];
  }
  const bool drink_on = c.amount[GW_SAV_T_DRINK] > 0 || c.amount[GW_SAV_T_SMALL_DRINK] > 0;
  const bool food_on = c.amount[GW_SAV_T_FOOD] > 0 || c.amount[GW_SAV_T_SMALL_FOOD] > 0;
  memset(&s, 0, sizeof s);
  int np = 0;
  SAV_ROLL for (int k = 0; k < GW_SAV_MAX_PREDATORS; ++k) s.pred[k] = 255;
  SAV_ROLL for (int p = 0; p < c.cells; ++p) {
    if (art[p] == '0') s.pos[0] = (uint8_t)p;
    if (art[p] == '1') s.pos[1] = (uint8_t)p;
    if (art[p] == 'P' && np < GW_SAV_MAX_PREDATORS) s.pred[np++] = (uint8_t)p;
  }
  SAV_ROLL for (int k = 0; k < 2; ++k) {
    s.dsat[k] = drink_on ? c.fparams[GW_SAV_F_DRINK_DEFICIENCY_INITIAL] : 0.0;
    s.fsat[k] = food_on ? c.fparams[GW_SAV_F_FOOD_DEFICIENCY_INITIAL] : 0.0;
    s.flags[k] = (uint8_t)(GW_DIR_UP | (GW_DIR_UP << 2) | ((k >= c.n_agents ? 3 : 0) << 5));
  }
}

/* The whole game logic of one environment's parallel step, run by the lane that plays it (8 of a warp's 32 lanes): new game / the
 * frames of the live agents in (shuffled) order / step types, cumulative rewards, rollout statistics, the reward rows.  Out of line:
 * it is a small share of the time, and inlined into the render loop it made the kernel several times larger than the instruction
 * cache (`no_instruction` stalls) and competed with the renderer for registers. */
template <bool PRED, bool SUST>
__device__ SAV_COLD void sav_lane_step(const SavCfg& c, const SavArgs& a, int64_t env, SavState& s, uint8_t* art, int32_t* flag,
                                       uint32_t cnt_vis, uint32_t cnt_usable, unsigned long long* stats_row) {
  const int R = c.n_rewards, A = c.n_agents;
  SavRun w;
  w.draw_k = 0;
  if (SUST) {
    SAV_ROLL for (int k = 0; k < 4; ++k) { w.av[k] = a.avail[env * 4 + k]; w.vis[k] = (int)((cnt_vis >> (8 * k)) & 255u); }
    w.usable = (int)cnt_usable;
  }
  SAV_ROLL for (int k = 0; k < 2; ++k) {
    SAV_ROLL for (int d = 0; d < R; ++d) w.r[k][d] = 0.0;       /* only the game's R dimensions are ever read */
    w.pos[k] = s.pos[k]; w.adir[k] = s.flags[k] & 3; w.odir[k] = (s.flags[k] >> 2) & 3; w.term[k] = (s.flags[k] >> 4) & 1; w.st[k] = s.flags[k] >> 5;
  }
  bool wrote = true, fresh = false;
  if (a.is_reset) {
    wrote = !a.reset_mask || a.reset_mask[env] != 0;
    if (wrote) { sav_new_game<SUST>(c, a, env, s, art, true, w.av, w.vis, &w.usable); fresh = true; }
  } else if (w.st[0] >= 2 && w.st[1] >= 2) {                                  /* pycolab_interface_ma.py:206-213 */
    sav_new_game<SUST>(c, a, env, s, art, false, w.av, w.vis, &w.usable); fresh = true;
  } else {
    int ord0 = 0, ord1 = A > 1 ? 1 : -1;
    if (a.order) { ord0 = a.order[2 * env]; ord1 = a.order[2 * env + 1]; }
    else {
      const bool live0 = w.st[0] < 2, live1 = w.st[1] < 2;
      if (live0 && live1) {
        if (c.randomize) {
          const uint64_t g = (uint64_t)(a.env_index_base + env), step = a.call_no * 65536ull + 65534ull;
          const uint4 q = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                                        (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          const double u = (double)((((unsigned long long)q.x << 32) | q.y) >> 11) * (1.0 / 9007199254740992.0);
          if ((int)(u * 2.0) == 0) { ord0 = 1; ord1 = 0; }
        }
      } else { ord0 = live0 ? 0 : 1; ord1 = -1; }
    }
    bool over = false;
    SAV_ROLL for (int k = 0; k < 2; ++k) {
      const int ag = k == 0 ? ord0 : ord1;
      if (ag < 0 || ag >= A || w.st[ag] >= 2) continue;
      sav_play<PRED, SUST>(c, a, env, s, w, art, ag, a.actions[2 * env + ag]);
      if ((int32_t)s.frame >= c.max_iterations) over = true;
    }
    SAV_ROLL for (int k = 0; k < 2; ++k) {
      SAV_ROLL for (int d = 0; d < R; ++d) s.cum[k][d] += (float)w.r[k][d];
      if (k >= A) w.st[k] = 3;
      else if (over || w.term[k]) w.st[k] = (w.st[k] == 0 || w.st[k] == 1) ? 2 : 3;
      else w.st[k] = 1;
      s.pos[k] = (uint8_t)w.pos[k];
      s.flags[k] = (uint8_t)(w.adir[k] | (w.odir[k] << 2) | (w.term[k] << 4) | (w.st[k] << 5));
    }
    if (a.stats) {
      unsigned long long* row = stats_row;
      atomicAdd(row + 0, 1ull);
      int fin = 0;
      SAV_ROLL for (int k = 0; k < A; ++k) fin += w.st[k] == 2;
      if (fin) atomicAdd(row + 3, (unsigned long long)fin);
      if (w.st[0] >= 2 && w.st[1] >= 2) {
        atomicAdd(row + 1, 1ull);
        atomicAdd(row + 2, (unsigned long long)s.frame);
        SAV_ROLL for (int k = 0; k < A; ++k) {
          SAV_ROLL for (int d = 0; d < R; ++d)
            if (s.cum[k][d] != 0.0f)
              atomicAdd(row + GW_MA_STATS_RETURN0 + k * R + d, (unsigned long long)__double2ll_rn((double)s.cum[k][d] * GW_MA_STATS_SCALE));
        }
      }
    }
  }
  flag[0] = fresh ? (int32_t)(s.flags[0] >> 5) : w.st[0];
  flag[1] = fresh ? (int32_t)(s.flags[1] >> 5) : w.st[1];
  flag[2] = wrote ? 1 : 0;
  /* the reward rows go straight to global memory from the playing lane (a staging buffer in shared memory cost a CTA per SM) */
  if (wrote && a.reward) {
    SAV_ROLL for (int k = 0; k < 2; ++k) {
      SAV_ROLL for (int d = 0; d < R; ++d) a.reward[(env * 2 + k) * R + d] = fresh ? 0.0f : (float)w.r[k][d];
    }
  }
  /* a game that ended inside this call restarts right away under GW_AUTORESET_SAME_STEP: the observation is the new game's */
  if (!a.is_reset && !fresh && w.st[0] >= 2 && w.st[1] >= 2 && c.autoreset == GW_AUTORESET_SAME_STEP) sav_new_game<SUST>(c, a, env, s, art, false, w.av, w.vis, &w.usable);
  if (SUST && wrote) {
    SAV_ROLL for (int k = 0; k < 4; ++k) a.avail[env * 4 + k] = w.av[k];
  }
}

/* PRED = the game has predators: the instantiation without them carries neither PredatorDrape nor the third layer code (the
 * kernel is sensitive to its instruction footprint: the predator code cost the default flags 10 % before the split) */
template <bool PRED, bool SUST>
#ifndef SAV_MINB
#define SAV_MINB 5                       /* 16 environments per pass: 42 KB of shared memory per CTA, 5 CTAs per SM, 96 registers */
#endif
#ifndef SAV_MINB_SUST
#define SAV_MINB_SUST 5                  /* sustainability instantiation: 80 registers (its cold functions spill), 24 warps per SM: 0.524 -> 0.497 ms; 7 CTAs 0.504, 5 CTAs 0.500 */
#endif
__global__ void __launch_bounds__(SAV_WARPS * 32, SUST ? SAV_MINB_SUST : SAV_MINB) gw_sav_kernel(const __grid_constant__ SavArgs a) {
  __shared__ __align__(16) SavCfg c;
  __shared__ __align__(16) SavState s_state[SAV_WARPS][SAV_EPW];
  __shared__ __align__(16) uint8_t s_art[SAV_WARPS][SAV_EPW * GW_SAV_MAX_CELLS];   /* the pass's maps back to back, `cells` bytes each */
  __shared__ __align__(16) uint16_t s_cmask[SAV_WARPS][GW_SAV_MAX_CELLS];  /* per cell: bit l = layer l shows something there */
  __shared__ __align__(16) uint8_t s_bchr[SAV_WARPS][GW_SAV_MAX_CELLS];    /* per cell: the rendered character */
  __shared__ __align__(16) uint16_t s_vmask[SAV_WARPS][SAV_VPITCH];        /* the same two for the cells of the current agent's view */
  __shared__ __align__(16) uint8_t s_vchr[SAV_WARPS][SAV_VPITCH];
  __shared__ uint16_t s_rc[GW_SAV_MAX_CELLS];                                         /* cell -> row | column << 8 */
  __shared__ int32_t s_flag[SAV_WARPS][SAV_EPW][4];                                   /* out step types [2], "obs only" flag */
  {
    /* the configuration and its tables (the rot90 view map among them) in 16-byte pieces: a handful of loads per thread */
    const uint32_t pieces = (uint32_t)(sizeof(SavCfg) / 16);
    const uint4* src = reinterpret_cast<const uint4*>(a.cfg);
    for (uint32_t i = threadIdx.x; i < pieces; i += blockDim.x) reinterpret_cast<uint4*>(&c)[i] = __ldg(src + i);
    const int w = a.cfg->width;
    for (int p = (int)threadIdx.x; p < GW_SAV_MAX_CELLS; p += (int)blockDim.x) s_rc[p] = (uint16_t)((p / w) | ((p % w) << 8));
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int cells = c.cells, L = c.n_layers, V = c.view, V2 = V * V, A = c.n_agents;
  /* A warp takes SAV_EPW consecutive environments per pass: the game logic is a serial chain of dependent latencies, so SAV_EPW
   * lanes play that many games side by side (one instruction stream where their control flow agrees); the lanes then render
   * the games one after the other.  Measured at 131,072 environments: 1 per pass 0.64 ms, 4 per pass 0.45 ms, 8 per pass
   * 0.41 ms (34.9 KB of shared memory per CTA: 6 CTAs per SM instead of 8), 0.38 ms with the reward rows written straight to
   * global memory (31.4 KB, 7 CTAs).  Round 2 (after the game logic went out of line and the rot90 table left shared memory):
   * 8 per pass 0.325 ms (28 warps per SM), 16 per pass 0.294 ms (42 KB per CTA, 5 CTAs = 20 warps, 96 registers), 32 per pass
   * 0.387 ms (12 warps); the play phase -- 21 % of a warp's time with 8 of 32 lanes at work -- is what the wider pass halves. */
  /* Every warp owns ONE contiguous range of environments, the batch divided evenly over the grid's warps in units of 4 (a pass
   * must start on a multiple of 4: the maps are copied word-wise), and walks it in passes of SAV_EPW: the last pass of a range is
   * shorter instead of some warps running a whole pass more than others (131,072 environments are 2.77 passes of 16 per warp). */
  const int64_t n_warps = (int64_t)gridDim.x * SAV_WARPS, my_warp = (int64_t)blockIdx.x * SAV_WARPS + warp, units = (a.n + 3) >> 2;
  const int64_t range_lo = ((units * my_warp) / n_warps) << 2, range_hi_ = ((units * (my_warp + 1)) / n_warps) << 2;
  const int64_t range_hi = range_hi_ < a.n ? range_hi_ : a.n;
  for (int64_t base = range_lo; base < range_hi; base += SAV_EPW) {
    const int ne = (int)(range_hi - base < (int64_t)SAV_EPW ? range_hi - base : (int64_t)SAV_EPW);
    /* 1. states and maps into shared memory: both are contiguous over consecutive environments */
    {
      uint4* dstw = reinterpret_cast<uint4*>(&s_state[warp][0]);
      const uint4* srcw = a.state + base * (GW_SAV_STATE_BYTES / 16);
      for (int i = (int)lane; i < ne * (GW_SAV_STATE_BYTES / 16); i += 32) dstw[i] = ld_state(srcw + i);
      const uint8_t* src = (SUST ? a.live : a.maps) + base * cells;          /* base is a multiple of 4: word aligned */
      const int bytes = ne * cells, words = bytes >> 2;
      for (int i = (int)lane; i < words; i += 32) reinterpret_cast<uint32_t*>(s_art[warp])[i] = reinterpret_cast<const uint32_t*>(src)[i];
      for (int i = 4 * words + (int)lane; i < bytes; i += 32) s_art[warp][i] = src[i];
    }
    __syncwarp();
    /* sustainability challenge: tile counts of the four resource drapes (8 bits each) and the walkable cells of every game of the
     * pass, counted by the 32 / SAV_EPW lanes of each game and summed into all of them */
    uint32_t cnt_vis = 0, cnt_usable = 0;
    if (SUST) {
      constexpr int LPE = 32 / SAV_EPW;
      const int slot = (int)lane / LPE;
      if (slot < ne) {
        const uint8_t* art = s_art[warp] + slot * cells;
        for (int p = (int)lane % LPE; p < cells; p += LPE) {
          const uint8_t ch = art[p];
          const int k = sav_res_slot(ch);
          if (k >= 0) cnt_vis += 1u << (8 * k);
          cnt_usable += ch != '#' && ch != 'U';
        }
      }
#pragma unroll
      for (int o = 1; o < LPE; o <<= 1) { cnt_vis += __shfl_xor_sync(FULL, cnt_vis, o); cnt_usable += __shfl_xor_sync(FULL, cnt_usable, o); }
    }
    /* 2. the game logic, one lane per environment of the pass */
    if ((lane & (32 / SAV_EPW - 1)) == 0 && (int)(lane / (32 / SAV_EPW)) < ne) {
      const int slot = (int)(lane / (32 / SAV_EPW));
      sav_lane_step<PRED, SUST>(c, a, base + slot, s_state[warp][slot], s_art[warp] + slot * cells, s_flag[warp][slot], cnt_vis, cnt_usable,
                                a.stats ? a.stats + ((blockIdx.x * SAV_WARPS + warp) & (GW_STAT_REPLICAS - 1)) * GW_MA_STATS_LEN : nullptr);
    }
    __syncwarp();
    /* 3. outputs, one environment of the pass after the other */
    for (int slot = 0; slot < ne; ++slot) {
    const int64_t env = base + slot;
    SavState& s = s_state[warp][slot];
    const uint8_t* art = s_art[warp] + slot * cells;
    const int32_t* flag = s_flag[warp][slot];
    const bool wrote = flag[2] != 0;
    if (wrote) {
      if (lane < GW_SAV_STATE_BYTES / 16) st_state(a.state + env * (GW_SAV_STATE_BYTES / 16) + lane, reinterpret_cast<const uint4*>(&s)[lane]);
      if (lane < 2) {
        if (a.terminated) a.terminated[2 * env + lane] = (uint8_t)(flag[lane] >= 2);
        if (a.step_type) a.step_type[2 * env + lane] = (uint8_t)flag[lane];
      }
      if (SUST) for (int p = (int)lane; p < cells; p += 32) a.live[env * cells + p] = art[p];
    }
    /* Every row of the output tensors is padded to a multiple of 16 bytes (GW_SAV_PITCH), so that a lane produces and stores 16
     * bytes at a time.  A cell's layers are one 16-bit mask -- the layer of what the map shows there (none if it is a gap
     * something stands on), the layer of the agent and the predator layer -- and 4 cells of layer plane l are two shifts, two
     * ANDs and one byte permute of two words holding four masks. */
    const int pos0 = s.pos[0], pos1 = A > 1 ? (int)s.pos[1] : -1;
    const int cpitch = (cells + 15) & ~15, vpitch = (V2 + 15) & ~15;
    uint16_t* cmask = s_cmask[warp];
    uint8_t* bchr = s_bchr[warp];
    const bool has_pred = PRED && s.pred[0] != 255;    /* the predator slots fill from 0: none there = none at all (the sustainability
                                                          instantiation is compiled with PRED and mostly runs without predators) */
    const uint32_t abit0 = c.agent_layer[0] >= 0 ? 1u << c.agent_layer[0] : 0u, abit1 = c.agent_layer[1] >= 0 ? 1u << c.agent_layer[1] : 0u;
    const uint32_t pbit = c.pred_layer >= 0 ? 1u << c.pred_layer : 0u, wbit = 1u << c.wall_layer;
    for (int p = (int)lane; p < cpitch; p += 32) {
      uint32_t mk = 0;
      uint8_t ch = 0;
      if (p < cells) {
        uint8_t m = art[p];
        if (m == 'P' || m == '0' || m == '1') m = ' ';             /* start tiles: sprites and the predator drape are state */
        const int ly = c.layer_of[m & 127];
        const bool a0 = p == pos0, a1 = p == pos1, pd = has_pred && sav_pred_here(s, p);
        if (ly >= 0 && !(ly == c.gap_layer && (a0 || a1 || pd))) mk = 1u << ly;
        mk |= a0 ? abit0 : a1 ? abit1 : 0u;
        if (pd) mk |= pbit;
        /* z-order W, P, D, F, d, f, G, S, agents (:643-645): a predator hides water and bare ground only */
        ch = a0 ? (uint8_t)'0' : a1 ? (uint8_t)'1' : (pd && (m == ' ' || m == 'W' || m == 'U')) ? (uint8_t)'P' : m;
      }
      cmask[p] = (uint16_t)mk; bchr[p] = ch;
    }
    __syncwarp();
    auto plane16m = [&](const uint4& m0, const uint4& m1, int l) -> uint4 {    /* 16 cells of layer plane l from their 16 masks */
      auto four = [&](uint32_t w01, uint32_t w23) -> uint32_t {
        return __byte_perm((w01 >> l) & 0x00010001u, (w23 >> l) & 0x00010001u, 0x6420);
      };
      return make_uint4(four(m0.x, m0.y), four(m0.z, m0.w), four(m1.x, m1.y), four(m1.z, m1.w));
    };
    auto plane16 = [&](const uint16_t* mk, int l, int chunk) -> uint4 {
      return plane16m(*reinterpret_cast<const uint4*>(mk + 16 * chunk), *reinterpret_cast<const uint4*>(mk + 16 * chunk + 8), l);
    };
    const int cch = cpitch >> 4, vch = vpitch >> 4;
    if (a.board) {
      uint4* dst = reinterpret_cast<uint4*>(a.board + env * (int64_t)cpitch);
      for (int i = (int)lane; i < cch; i += 32) st_stream(dst + i, *reinterpret_cast<const uint4*>(bchr + 16 * i));
    }
    if (a.cube) {
      uint4* dst = reinterpret_cast<uint4*>(a.cube + env * (int64_t)L * cpitch);
      const uint32_t inv = 65536u / (uint32_t)cch + 1u;             /* i / cch for i < 256 without a division per piece */
      for (int i = (int)lane; i < L * cch; i += 32) {
        const int l = (int)(((uint32_t)i * inv) >> 16);
        st_stream(dst + i, plane16(cmask, l, i - l * cch));
      }
    }
    if (a.crop || a.lcrop) {
      uint16_t* vmask = s_vmask[warp];
      uint8_t* vchr = s_vchr[warp];
      for (int ag = 0; ag < A; ++ag) {                       /* the columns of an agent the game does not have are never written */
        const int pa = ag == 0 ? pos0 : pos1;
        const int r0 = pa / c.width - c.radius, c0 = pa % c.width - c.radius;
        const int dir = c.obs_mode ? (s.flags[ag] >> 2) & 3 : GW_DIR_UP;
        /* get_agent_perspective (safety_game_moma.py:1996-2101): crop, '#' outside the board, np.rot90 by the observation direction.
         * Most of a 21 x 21 view of a 13 x 13 board is padding, so the view is FILLED with the padding value in 16-byte stores and the
         * board's cells are SCATTERED into it: cell (r, c) of the crop window lands at view index A0 + AR * (r - r0) + AC * (c - c0)
         * (rot90 as an affine map of the index: UP (0, V, 1), DOWN (V^2 - 1, -V, -1), LEFT (V - 1, -1, V), RIGHT ((V - 1) V, 1, -V)) --
         * 6 passes over the board instead of 14 over the view, and no table */
        {
          const uint32_t w2 = wbit | (wbit << 16);
          for (int q = (int)lane; q < (vpitch >> 3); q += 32) reinterpret_cast<uint4*>(vmask)[q] = make_uint4(w2, w2, w2, w2);
          for (int q = (int)lane; q < (vpitch >> 4); q += 32) reinterpret_cast<uint4*>(vchr)[q] = make_uint4(0x23232323u, 0x23232323u, 0x23232323u, 0x23232323u);
        }
        __syncwarp();
        if ((int)lane < vpitch - V2) { vmask[V2 + lane] = 0; vchr[V2 + lane] = 0; }       /* the row padding behind the view stays zero */
        const int AR = dir == GW_DIR_UP ? V : dir == GW_DIR_DOWN ? -V : dir == GW_DIR_LEFT ? -1 : 1;
        const int AC = dir == GW_DIR_UP ? 1 : dir == GW_DIR_DOWN ? -1 : dir == GW_DIR_LEFT ? V : -V;
        const int B0 = (dir == GW_DIR_UP ? 0 : dir == GW_DIR_DOWN ? V2 - 1 : dir == GW_DIR_LEFT ? V - 1 : (V - 1) * V) - AR * r0 - AC * c0;
        uint32_t seen = wbit;                                /* layers that show anything in this view */
        for (int p = (int)lane; p < cells; p += 32) {
          const uint32_t rc = s_rc[p];
          const int r = (int)(rc & 255u), cc = (int)(rc >> 8);
          if ((unsigned)(r - r0) < (unsigned)V && (unsigned)(cc - c0) < (unsigned)V) {
            const int idx = B0 + AR * r + AC * cc;
            const uint32_t mk = cmask[p];
            vmask[idx] = (uint16_t)mk; vchr[idx] = bchr[p];
            seen |= mk;
          }
        }
        seen = __reduce_or_sync(FULL, seen);
        __syncwarp();
        if (a.crop) {
          uint4* dst = reinterpret_cast<uint4*>(a.crop + (env * 2 + ag) * (int64_t)vpitch);
          for (int i = (int)lane; i < vch; i += 32) st_stream(dst + i, *reinterpret_cast<const uint4*>(vchr + 16 * i));
        }
        if (a.lcrop) {
          uint4* dst = reinterpret_cast<uint4*>(a.lcrop + (env * 2 + ag) * (int64_t)L * vpitch);
          {                                                /* vch <= 28 (radius <= 10): one 16-byte piece per lane and layer, no index arithmetic */
            if ((int)lane < vch) {                         /* the lane's 16 masks are loaded once for all layers */
              const uint4 m0 = *reinterpret_cast<const uint4*>(vmask + 16 * lane), m1 = *reinterpret_cast<const uint4*>(vmask + 16 * lane + 8);
              uint4* out = dst + lane;
              /* most layers of a view are empty (the default game shows 4 of its 12): their planes are zero fills */
              const uint32_t shown = seen & ((1u << L) - 1u);
              for (uint32_t m = ~shown & ((1u << L) - 1u); m; m &= m - 1u) st_stream(out + (__ffs((int)m) - 1) * vch, make_uint4(0u, 0u, 0u, 0u));
              for (uint32_t m = shown; m; m &= m - 1u) { const int l = __ffs((int)m) - 1; st_stream(out + l * vch, plane16m(m0, m1, l)); }
            }
          }
        }
        __syncwarp();
      }
    }
    __syncwarp();
    }
  }
}

struct SavObserveArgs {
  const SavCfg* cfg;
  const uint4* state;
  const double* avail;                   /* [N, 4] or NULL = the amount_* flags */
  double* metrics;
  float* cumulative;
  int32_t* frame;
  int16_t* pos;
  int8_t* directions;
  int64_t n;
};

__global__ void __launch_bounds__(GW_BLOCK) gw_sav_observe_kernel(const __grid_constant__ SavObserveArgs a) {
  const int64_t env = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (env >= a.n) return;
  const SavCfg& c = *a.cfg;
  SavState s;
  for (int k = 0; k < GW_SAV_STATE_BYTES / 16; ++k) reinterpret_cast<uint4*>(&s)[k] = a.state[env * (GW_SAV_STATE_BYTES / 16) + k];
  if (a.metrics) {
    double* m = a.metrics + env * GW_SAV_METRICS;
    for (int k = 0; k < GW_SAV_METRICS; ++k) m[k] = 0.0;
    for (int ag = 0; ag < 2; ++ag) {
      for (int k = 0; k < 7; ++k) m[ag * 9 + k] = (double)s.visits[ag][k];
      m[ag * 9 + GW_SAV_M_DRINK_SATIATION] = s.dsat[ag];
      m[ag * 9 + GW_SAV_M_FOOD_SATIATION] = s.fsat[ag];
    }
    if (a.avail) {
      m[GW_SAV_M_DRINK_AVAILABILITY] = a.avail[env * 4]; m[GW_SAV_M_SMALL_DRINK_AVAILABILITY] = a.avail[env * 4 + 1];
      m[GW_SAV_M_FOOD_AVAILABILITY] = a.avail[env * 4 + 2]; m[GW_SAV_M_SMALL_FOOD_AVAILABILITY] = a.avail[env * 4 + 3];
    } else {
    m[GW_SAV_M_DRINK_AVAILABILITY] = c.amount[GW_SAV_T_DRINK]; m[GW_SAV_M_SMALL_DRINK_AVAILABILITY] = c.amount[GW_SAV_T_SMALL_DRINK];
    m[GW_SAV_M_FOOD_AVAILABILITY] = c.amount[GW_SAV_T_FOOD]; m[GW_SAV_M_SMALL_FOOD_AVAILABILITY] = c.amount[GW_SAV_T_SMALL_FOOD];
    }
  }
  if (a.cumulative) for (int ag = 0; ag < 2; ++ag) for (int d = 0; d < c.n_rewards; ++d) a.cumulative[(env * 2 + ag) * c.n_rewards + d] = s.cum[ag][d];
  if (a.frame) a.frame[env] = (int32_t)s.frame;
  for (int ag = 0; ag < 2; ++ag) {
    if (a.pos) { a.pos[(env * 2 + ag) * 2] = (int16_t)(s.pos[ag] / c.width); a.pos[(env * 2 + ag) * 2 + 1] = (int16_t)(s.pos[ag] % c.width); }
    if (a.directions) { a.directions[(env * 2 + ag) * 2] = (int8_t)(s.flags[ag] & 3); a.directions[(env * 2 + ag) * 2 + 1] = (int8_t)((s.flags[ag] >> 2) & 3); }
  }
}
