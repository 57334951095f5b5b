/*
 * gw_firemaker_oracle.c -- CPU restatement of firemaker_ex_ma's parallel step (BASELINE config 4).
 * TEST INFRASTRUCTURE ONLY (see gw_oracle.c for who may load it).
 *
 * One parallel step = one full Engine.play per agent in (shuffled) order
 * (environments/shared/rl/pycolab_interface_ma.py:173-246); each play updates, one update group per
 * entity in schedule order ['1','2','S','B','W','F','-'] (environments/firemaker_ex_ma.py:352-355; a
 * flat schedule = one group per entity, pycolab/ascii_art.py:236-240): the acting agent
 * (safety_game_ma.py:769-809, firemaker_ex_ma.py:430-476), StopButtonDrape (:656-673), WorkshopDrape
 * (:496-517), FireDrape (:539-629), WorkshopTerritoryDrape (:681-709).  PINNED by
 * tests/test_oracle_firemaker_golden.py against tests/golden/firemaker_*.npz, recorded from the
 * running reference by oracle/record_firemaker.py with every FireDrape draw and shuffle order.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gwsim_fm.h"
#include "../include/gwsim_ima.h"       /* GwDirection */

/* gw_oracle.c: runs fn(ctx, lo, hi) over [0, n) split across the host threads set with or_set_threads */
void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx);

#define S GW_FM_SIDE
#define CELLS GW_FM_CELLS
#define NA GW_FM_AGENTS

static const char AGENT_CHR[NA] = {'1', '2', 'S'};
static const char LAYER_ORDER[GW_FM_LAYERS] = {' ', '#', '-', '1', '2', 'B', 'F', 'S', 'W'};

typedef struct {
  int frame;
  int pos[NA];
  uint8_t fire[CELLS];
  int countdown, ext_fires;
  int at_workshop[NA];
  int adir[NA], odir[NA];            /* AgentSafetySprite.action_direction, AgentSprite.observation_direction (direction modes 1-2) */
  int visits[NA][5];                 /* external, internal, workshop, fire, stop button */
  int step_type[NA];
  double cum[7];                     /* worker 1 [ENERGY, WORKSHOP], worker 2, supervisor [ENERGY, EXTERNAL_FIRE, TRESPASSING] */
  uint8_t board[CELLS];              /* last render */
} FEnv;

typedef struct {
  GwFmConfig cfg;
  int64_t n, env_index_base;
  uint64_t seed, call_no;
  uint8_t backdrop[CELLS], workshop[CELLS], button[CELLS], territory[CELLS];
  int start[NA], static2;            /* static2: the '2' tile when amount_agents = 2 (backdrop character, layer only) */
  double spread_p[5][5];             /* spread probability by (dr + 2, dc + 2) */
  FEnv* envs;
} FOracle;

void or_philox(uint64_t seed, uint64_t env, uint64_t step, uint32_t out[4]);   /* gw_oracle.c */

/* The Philox streams the CUDA kernel shares (csrc/gwsim_fm.cuh): counter (global environment, call * 65536 + evaluation index).
 * philox_pair: the two 53-bit uniforms of an evaluation (the shuffle of the agents' order: draw k = half k & 1 of evaluation k >> 1);
 * philox_uniform: the FireDrape draws, one 32-bit word each (draw k = word k & 3 of evaluation k >> 2, u = word / 2^32). */
static double philox_pair(const FOracle* o, int64_t env, uint32_t k) {
  uint32_t r[4];
  or_philox(o->seed, (uint64_t)(o->env_index_base + env), o->call_no * 65536ull + (k >> 1), r);
  const uint32_t hi = (k & 1u) ? r[2] : r[0], lo = (k & 1u) ? r[3] : r[1];
  return (double)((((uint64_t)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}
static double philox_uniform(const FOracle* o, int64_t env, uint32_t k) {
  uint32_t r[4];
  or_philox(o->seed, (uint64_t)(o->env_index_base + env), o->call_no * 65536ull + (k >> 2), r);
  return (double)r[k & 3u] * (1.0 / 4294967296.0);
}

/* Engine._render, z_order ['-','W','F','B','1','2','S'] (firemaker_ex_ma.py:346-350) */
static void render(const FOracle* o, FEnv* e) {
  for (int p = 0; p < CELLS; ++p) {
    uint8_t ch = o->backdrop[p];
    if (o->territory[p]) ch = '-';
    if (o->workshop[p]) ch = 'W';
    if (e->fire[p]) ch = 'F';
    if (o->button[p]) ch = 'B';
    e->board[p] = ch;
  }
  for (int a = 0; a < NA; ++a) if (e->pos[a] >= 0) e->board[e->pos[a]] = (uint8_t)AGENT_CHR[a];
}

/* FireDrape.update (firemaker_ex_ma.py:539-629).  next_draw() supplies the uniform draws. */
typedef struct { const FOracle* o; int64_t env; const double* replay; int64_t replay_n; uint32_t k; } DrawSrc;
static double next_draw(DrawSrc* d) {
  const uint32_t k = d->k++;
  if (d->replay) return (int64_t)k < d->replay_n ? d->replay[k] : 2.0;
  return philox_uniform(d->o, d->env, k);
}

static int fire_update(const FOracle* o, FEnv* e, DrawSrc* d) {
  for (int a = 0; a < NA; ++a) if (e->pos[a] >= 0) e->fire[e->pos[a]] = 0;   /* :543-545 */
  int from[CELLS + NA], nfrom = 0;
  for (int p = 0; p < CELLS; ++p) if (e->fire[p]) from[nfrom++] = p;          /* np.where: row-major */
  if (e->countdown == 0)                                                      /* :555-559 */
    for (int a = 0; a < 2; ++a) if (e->at_workshop[a]) from[nfrom++] = e->pos[a];
  double cumulative[CELLS];
  memset(cumulative, 0, sizeof cumulative);
  for (int i = 0; i < nfrom; ++i) {                                           /* :567-606 */
    const int fr = from[i] / S, fc = from[i] % S;
    for (int tr = (fr - 2 > 0 ? fr - 2 : 0); tr < (fr + 3 < S ? fr + 3 : S); ++tr)
      for (int tc = (fc - 2 > 0 ? fc - 2 : 0); tc < (fc + 3 < S ? fc + 3 : S); ++tc) {
        const int t = tr * S + tc;
        if (e->fire[t]) continue;
        /* "fires cannot spread to under players": the reference's `continue` only leaves its inner loop (:579-581) */
        if (o->workshop[t] || o->button[t] || o->backdrop[t] == '#') continue;
        const int dr = fr - tr, dc = fc - tc;
        if (sqrt((double)(dr * dr + dc * dc)) < o->cfg.fire_spread_exclusive_max_distance) {
          const double p = o->spread_p[dr + 2][dc + 2];
          cumulative[t] = 1 - (1 - cumulative[t]) * (1 - p);
        }
      }
  }
  for (int t = 0; t < CELLS; ++t)                                             /* :612-615 */
    if (cumulative[t] > 0) e->fire[t] = next_draw(d) < cumulative[t];
  for (int i = 0; i < nfrom; ++i)                                             /* :619-621 */
    if (e->fire[from[i]]) e->fire[from[i]] = next_draw(d) < o->cfg.fire_continuation_probability;
  int ext = 0;
  for (int p = 0; p < CELLS; ++p) ext += e->fire[p] && !o->territory[p];      /* :624-625 */
  e->ext_fires = ext;
  return ext;
}

/* get_absolute_action / get_new_action_or_observation_direction, mode 1 (safety_game_ma.py:505-587): the action is relative to
 * `dir`; UP = forwards, DOWN = backwards, LEFT / RIGHT = a quarter turn.  Returns the absolute direction. */
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

/* one Engine.play({agent: action}); has_action = 0 is the frame-0 pass of its_showtime */
static void play(const FOracle* o, FEnv* e, int has_action, int agent, int action, double r[7], DrawSrc* d) {
  const double* R = o->cfg.rewards;
  static const int roff[NA] = {0, 2, 4};
  e->frame += 1;
  if (has_action) {                                                           /* the acting agent's sprite */
    int dr = 0, dc = 0;
    const int obs_mode = o->cfg.observation_direction_mode, act_mode = o->cfg.action_direction_mode;
    /* AgentSprite.update (firemaker_ex_ma.py:468-476): the observation direction turns first (safety_game_ma.py:640-698), then
     * AgentSafetySpriteMo.update maps the action through the action direction (:769-809) */
    if (action != GW_ACT_NOOP && obs_mode == 1 && act_mode == 1) e->odir[agent] = relative_to_absolute(action, e->odir[agent]);
    if (obs_mode == 2) e->odir[agent] = turned(action, e->odir[agent]);
    if (act_mode == 2 && action >= GW_ACT_TURN_LEFT_90) e->adir[agent] = turned(action, e->adir[agent]);   /* a turn moves nothing */
    else if (action >= GW_ACT_LEFT && action <= GW_ACT_DOWN) {
      int dir;
      if (act_mode >= 1) dir = relative_to_absolute(action, e->adir[agent]);
      else dir = action == GW_ACT_LEFT ? GW_DIR_LEFT : action == GW_ACT_RIGHT ? GW_DIR_RIGHT : action == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
      dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0; dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
      if (act_mode == 1) e->adir[agent] = dir;                                  /* map_action_to_action_direction (:724-766) */
    }
    if (dr || dc) {
      const int nr = e->pos[agent] / S + dr, nc = e->pos[agent] % S + dc;
      if (nr >= 0 && nr < S && nc >= 0 && nc < S) {
        const uint8_t ch = e->board[nr * S + nc];                             /* impassable: '#' and the other agents (:399-400) */
        int blocked = ch == '#';
        for (int a = 0; a < NA; ++a) if (a != agent && e->pos[a] >= 0 && ch == (uint8_t)AGENT_CHR[a]) blocked = 1;
        if (!blocked) e->pos[agent] = nr * S + nc;
      }
    }
    if (action != GW_ACT_NOOP) r[roff[agent]] += agent == 2 ? R[GW_FM_R_SUP_MOVEMENT] : R[GW_FM_R_AGENT_MOVEMENT];   /* :436-441 */
    const int p = e->pos[agent];                                              /* layers of the previous render (:449-463) */
    if (p != o->static2) e->visits[agent][0] += 1;                            /* the backdrop ' ' layer is True off the walls (and off a static '2' tile) */
    if (o->territory[p]) e->visits[agent][1] += 1;
    if (o->workshop[p]) e->visits[agent][2] += 1;
    if (e->fire[p]) e->visits[agent][3] += 1;
    if (o->button[p]) e->visits[agent][4] += 1;
  }
  render(o, e);
  for (int a = 0; a < NA; ++a)                                                /* StopButtonDrape.update :656-673 */
    if (e->pos[a] >= 0 && o->button[e->pos[a]]) {
      e->countdown = 1 + 1 + o->cfg.stop_button_duration;
      if (a == 2 && e->ext_fires == 0) r[6] += R[GW_FM_R_SUP_STOP_BUTTON];
    }
  e->countdown = e->countdown - 1 > 0 ? e->countdown - 1 : 0;
  for (int a = 0; a < NA; ++a) {                                              /* WorkshopDrape.update :496-517 */
    e->at_workshop[a] = e->pos[a] >= 0 && o->workshop[e->pos[a]];
    if (!e->at_workshop[a]) continue;
    if (a == 2 && e->ext_fires == 0) r[6] += R[GW_FM_R_SUP_WORKSHOP];
    else if (e->countdown == 0) {
      r[1] += R[GW_FM_R_WORKSHOP_WORK];
      if (o->cfg.amount_agents > 2) r[3] += R[GW_FM_R_WORKSHOP_WORK];           /* :512-513 */
      r[roff[a]] += R[GW_FM_R_WORKSHOP_ENERGY];
    }
  }
  const int ext = fire_update(o, e, d);
  r[5] += (double)ext * R[GW_FM_R_SUP_EXTERNAL_FIRE];                         /* :626-627 */
  if (o->territory[e->pos[2]] && e->ext_fires == 0) r[6] += R[GW_FM_R_SUP_TRESPASSING];   /* WorkshopTerritoryDrape :699-704 */
  render(o, e);
}

static void env_reset(const FOracle* o, FEnv* e) {
  memset(e, 0, sizeof *e);
  for (int a = 0; a < NA; ++a) { e->pos[a] = o->start[a]; e->adir[a] = e->odir[a] = GW_DIR_UP; }
  e->frame = -1;
  render(o, e);
  double r[7] = {0};
  DrawSrc d = {o, 0, 0, 0, 0};
  play(o, e, 0, 0, 0, r, &d);                                                 /* frame 0: no sources, no draws */
  for (int a = 0; a < NA; ++a) e->step_type[a] = 0;
  memset(e->cum, 0, sizeof e->cum);
}

typedef struct {
  uint8_t *board, *cube, *crop_w, *crop_s, *lcrop_w, *lcrop_s;
  float *reward_w, *reward_s;
  uint8_t *terminated, *step_type;
} FOut;

static uint8_t layer_bit(const FOracle* o, const FEnv* e, int l, int p) {
  const char ch = LAYER_ORDER[l];
  switch (ch) {
    case '#': return o->backdrop[p] == '#';
    case '-': return o->territory[p];
    case '1': return p == e->pos[0];
    case '2': return p == e->pos[1] || p == o->static2;
    case 'S': return p == e->pos[2];
    case 'B': return o->button[p];
    case 'F': return e->fire[p];
    case 'W': return o->workshop[p];
    default:                        /* ' ': backdrop gap AND NOT any other layer (observation_distiller_ex.py:165-178) */
      return o->backdrop[p] == ' ' && !o->territory[p] && !o->button[p] && !e->fire[p] && !o->workshop[p] &&
             p != e->pos[0] && p != e->pos[1] && p != e->pos[2];
  }
}

/* get_agent_perspective (safety_game_moma.py:1996-2101): crop around the agent, what_lies_outside ('#') beyond the board; a
 * layer pads with (layer chr == '#'); then np.rot90 by the observation direction (DOWN k=2, LEFT k=-1, RIGHT k=1) unless the
 * observation direction mode is 0 */
static void crop(const FOracle* o, const FEnv* e, int agent, int radius, uint8_t* board_out, uint8_t* layers_out) {
  if (e->pos[agent] < 0) return;                     /* amount_agents = 2: no worker '2', its columns stay as they are (zero) */
  const int side = 2 * radius + 1, r0 = e->pos[agent] / S - radius, c0 = e->pos[agent] % S - radius;
  const int dir = o->cfg.observation_direction_mode ? e->odir[agent] : GW_DIR_UP;
  for (int i = 0; i < side; ++i)
    for (int j = 0; j < side; ++j) {
      int si = i, sj = j;                                                      /* out[i][j] = in[si][sj] */
      if (dir == GW_DIR_DOWN) { si = side - 1 - i; sj = side - 1 - j; }
      else if (dir == GW_DIR_LEFT) { si = side - 1 - j; sj = i; }               /* rot90 k=-1 (clockwise) */
      else if (dir == GW_DIR_RIGHT) { si = j; sj = side - 1 - i; }              /* rot90 k=1 (counterclockwise) */
      const int r = r0 + si, c = c0 + sj;
      const int inside = r >= 0 && r < S && c >= 0 && c < S;
      if (board_out) board_out[i * side + j] = inside ? e->board[r * S + c] : (uint8_t)'#';
      if (layers_out)
        for (int l = 0; l < GW_FM_LAYERS; ++l)
          layers_out[(l * side + i) * side + j] = inside ? layer_bit(o, e, l, r * S + c) : (uint8_t)(LAYER_ORDER[l] == '#');
    }
}

static void emit_obs(const FOracle* o, const FEnv* e, int64_t i, const FOut* out) {
  if (out->board) memcpy(out->board + i * CELLS, e->board, CELLS);
  if (out->cube)
    for (int l = 0; l < GW_FM_LAYERS; ++l)
      for (int p = 0; p < CELLS; ++p) out->cube[(i * GW_FM_LAYERS + l) * CELLS + p] = layer_bit(o, e, l, p);
  for (int a = 0; a < 2; ++a)
    crop(o, e, a, 2, out->crop_w ? out->crop_w + (i * 2 + a) * 25 : 0, out->lcrop_w ? out->lcrop_w + (i * 2 + a) * GW_FM_LAYERS * 25 : 0);
  crop(o, e, 2, S - 1, out->crop_s ? out->crop_s + i * 1089 : 0, out->lcrop_s ? out->lcrop_s + i * GW_FM_LAYERS * 1089 : 0);
}

static void emit_out(const FEnv* e, int64_t i, const FOut* out, const double r[7], const int st[NA]) {
  if (out->reward_w) for (int k = 0; k < 4; ++k) out->reward_w[i * 4 + k] = (float)r[k];
  if (out->reward_s) for (int k = 0; k < 3; ++k) out->reward_s[i * 3 + k] = (float)r[4 + k];
  for (int a = 0; a < NA; ++a) {
    if (out->terminated) out->terminated[i * NA + a] = (uint8_t)(st[a] >= 2);
    if (out->step_type) out->step_type[i * NA + a] = (uint8_t)st[a];
  }
  (void)e;
}

void* orf_create(const GwFmConfig* cfg, int64_t n, int64_t env_index_base, uint64_t seed) {
  if (!cfg || n <= 0) return 0;
  FOracle* o = (FOracle*)calloc(1, sizeof *o);
  o->cfg = *cfg; o->n = n; o->env_index_base = env_index_base; o->seed = seed;
  /* ascii_art_to_game: sprites and drapes are lifted, what_lies_beneath ' ' fills under them */
  for (int p = 0; p < CELLS; ++p) {
    const uint8_t ch = cfg->art[p];
    o->backdrop[p] = ch == '#' ? '#' : ' ';
    o->workshop[p] = ch == 'W'; o->button[p] = ch == 'B'; o->territory[p] = ch == '-';
    for (int a = 0; a < NA; ++a) if (ch == (uint8_t)AGENT_CHR[a]) o->start[a] = p;
  }
  o->static2 = -1;
  if (cfg->amount_agents == 2) { o->static2 = o->start[1]; o->start[1] = -1; }   /* no sprite: '2' is a backdrop character */
  /* WorkshopTerritoryDrape.__init__ (:689-696): extend under agents; the scan sees its own earlier additions */
  for (int r = 0; r < S; ++r)
    for (int c = 0; c < S; ++c) {
      const uint8_t ob = cfg->art[r * S + c];
      int above = 0, below = 0, left = 0, right = 0;
      for (int rr = 0; rr < r; ++rr) above |= o->territory[rr * S + c];
      for (int rr = r + 1; rr < S; ++rr) below |= o->territory[rr * S + c];
      if (!o->territory[r * S + c] && above && below && ob != 'W' && ob != 'B') o->territory[r * S + c] = 1;
      for (int cc = 0; cc < c; ++cc) left |= o->territory[r * S + cc];
      for (int cc = c + 1; cc < S; ++cc) right |= o->territory[r * S + cc];
      if (!o->territory[r * S + c] && left && right && ob != 'W' && ob != 'B') o->territory[r * S + c] = 1;
    }
  for (int dr = -2; dr <= 2; ++dr)
    for (int dc = -2; dc <= 2; ++dc) {                                        /* :595-598 */
      const double dist = sqrt((double)(dr * dr + dc * dc));
      const double rel = (dist - 1) / (cfg->fire_spread_exclusive_max_distance - 1 + 1e-15);
      o->spread_p[dr + 2][dc + 2] = (1 - rel) * cfg->fire_spread_probability_at_distance_one;
    }
  o->envs = (FEnv*)calloc((size_t)n, sizeof(FEnv));
  return o;
}

void orf_destroy(void* h) { FOracle* o = (FOracle*)h; if (o) { free(o->envs); free(o); } }

void orf_reset(void* h, const uint8_t* mask, uint8_t* board, uint8_t* cube, uint8_t* crop_w, uint8_t* crop_s, uint8_t* lcrop_w,
               uint8_t* lcrop_s, float* reward_w, float* reward_s, uint8_t* terminated, uint8_t* step_type) {
  FOracle* o = (FOracle*)h;
  FOut out = {board, cube, crop_w, crop_s, lcrop_w, lcrop_s, reward_w, reward_s, terminated, step_type};
  o->call_no += 1;
  const double zeros[7] = {0};
  for (int64_t i = 0; i < o->n; ++i) {
    if (!mask || mask[i]) { env_reset(o, &o->envs[i]); emit_out(&o->envs[i], i, &out, zeros, o->envs[i].step_type); }
    emit_obs(o, &o->envs[i], i, &out);
  }
}

typedef struct { FOracle* o; const int32_t* actions; const int32_t* order; const double* draws; int64_t draw_stride; FOut out; } FStepCtx;

static void step_range(void* ctx, int64_t lo, int64_t hi) {
  FStepCtx* sc = (FStepCtx*)ctx;
  FOracle* o = sc->o;
  const int32_t* actions = sc->actions; const int32_t* order = sc->order; const double* draws = sc->draws;
  const int64_t draw_stride = sc->draw_stride;
  FOut out = sc->out;
  for (int64_t i = lo; i < hi; ++i) {
    FEnv* e = &o->envs[i];
    double r[7] = {0};
    int all_done = 1;
    for (int a = 0; a < NA; ++a) all_done &= e->step_type[a] >= 2 || o->start[a] < 0;      /* an absent agent has no step type */
    if (all_done) {                                                           /* pycolab_interface_ma.py:206-213: drop episode, reset */
      env_reset(o, e);
      emit_out(e, i, &out, r, e->step_type);
      emit_obs(o, e, i, &out);
      continue;
    }
    int ord[NA] = {0, 1, 2};
    const int two_agents = o->cfg.amount_agents == 2;
    if (two_agents) { ord[1] = 2; ord[2] = -1; }
    if (order) { for (int a = 0; a < NA; ++a) ord[a] = order[i * NA + a]; }
    else if (o->cfg.randomize_order) {                                        /* Fisher-Yates on the Philox stream */
      for (int k = (two_agents ? 1 : NA - 1); k >= 1; --k) {
        const int j = (int)(philox_pair(o, i, 65533u + (uint32_t)k) * (k + 1));
        const int t = ord[k]; ord[k] = ord[j]; ord[j] = t;
      }
    }
    DrawSrc d = {o, i, draws ? draws + i * draw_stride : 0, draw_stride, 0};
    int over = 0;
    for (int k = 0; k < NA; ++k) {
      if (ord[k] < 0 || e->pos[ord[k]] < 0) continue;                         /* AEC: step({agent: action}) plays one agent's frame only */
      play(o, e, 1, ord[k], actions[i * NA + ord[k]], r, &d);
      if (e->frame >= o->cfg.max_iterations) over = 1;                        /* pycolab_interface_ma.py:429-430 */
    }
    for (int k = 0; k < 7; ++k) e->cum[k] += r[k];
    int st[NA];
    for (int a = 0; a < NA; ++a) {                                            /* :232-239 */
      if (o->start[a] < 0) e->step_type[a] = 0;
      else if (over) e->step_type[a] = (e->step_type[a] == 0 || e->step_type[a] == 1) ? 2 : 3;
      else e->step_type[a] = 1;
      st[a] = e->step_type[a];
    }
    if (over && o->cfg.autoreset_mode == GW_AUTORESET_SAME_STEP) env_reset(o, e);
    emit_out(e, i, &out, r, st);
    emit_obs(o, e, i, &out);
  }
}

void orf_step(void* h, const int32_t* actions, const int32_t* order, const double* draws, int64_t draw_stride, uint8_t* board,
              uint8_t* cube, uint8_t* crop_w, uint8_t* crop_s, uint8_t* lcrop_w, uint8_t* lcrop_s, float* reward_w, float* reward_s,
              uint8_t* terminated, uint8_t* step_type) {
  FOracle* o = (FOracle*)h;
  o->call_no += 1;
  FStepCtx sc = {o, actions, order, draws, draw_stride,
                 {board, cube, crop_w, crop_s, lcrop_w, lcrop_s, reward_w, reward_s, terminated, step_type}};
  or_parallel_for(o->n, step_range, &sc);
}

void orf_directions(void* h, int8_t* directions) {          /* [N, 3, 2]: action direction, observation direction */
  FOracle* o = (FOracle*)h;
  for (int64_t i = 0; i < o->n; ++i)
    for (int a = 0; a < NA; ++a) { directions[(i * NA + a) * 2] = (int8_t)o->envs[i].adir[a]; directions[(i * NA + a) * 2 + 1] = (int8_t)o->envs[i].odir[a]; }
}

void orf_observe(void* h, double* metrics, float* cumulative, int32_t* frame, int16_t* pos, int32_t* ext_fires) {
  FOracle* o = (FOracle*)h;
  for (int64_t i = 0; i < o->n; ++i) {
    const FEnv* e = &o->envs[i];
    if (metrics) {
      for (int a = 0; a < NA; ++a) for (int k = 0; k < 5; ++k) metrics[i * GW_FM_METRICS + a * 5 + k] = e->visits[a][k];
      metrics[i * GW_FM_METRICS + 15] = e->countdown;
    }
    if (cumulative) for (int k = 0; k < 7; ++k) cumulative[i * 7 + k] = (float)e->cum[k];
    if (frame) frame[i] = e->frame;
    if (pos) for (int a = 0; a < NA; ++a) {
      pos[(i * NA + a) * 2] = (int16_t)(e->pos[a] < 0 ? -1 : e->pos[a] / S);
      pos[(i * NA + a) * 2 + 1] = (int16_t)(e->pos[a] < 0 ? -1 : e->pos[a] % S);
    }
    if (ext_fires) ext_fires[i] = e->ext_fires;
  }
}
