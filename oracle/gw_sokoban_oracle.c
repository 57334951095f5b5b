/*
 * gw_sokoban_oracle.c -- CPU restatement of side_effects_sokoban on any of its levels (TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load it; the product never does).
 *
 * Written in the reference's own structure -- Engine.play with the update schedule [[boxes], [C], [A]]
 * (environments/side_effects_sokoban.py:155-172, pycolab/engine.py:698-735), each update group seeing the board rendered
 * after the previous one, BoxSprite.update / _calculate_wall_penalty / _update_wall_penalty (:249-317) evaluated on the
 * wall layer, AgentSprite.update_reward (:186-212), the 'X' repainter (:118,366), SafetyEnvironment's episode bookkeeping
 * (environments/shared/safety_game.py:246-300, rl/pycolab_interface.py:150-200).
 * Pinned by tests/golden/classic_sokoban_*.npz, recorded from the unmodified reference by oracle/record_classic.py.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/gwsim_sok.h"

/* gw_oracle.c: runs fn(ctx, lo, hi) over [0, n) split across the host threads set with or_set_threads */
void or_parallel_for(int64_t n, void (*fn)(void* ctx, int64_t lo, int64_t hi), void* ctx);

#define MAXC GW_SOK_MAX_CELLS

typedef struct {
  int frame;
  uint8_t board[MAXC];              /* rendered after the last update group */
  uint8_t backdrop[MAXC];
  uint8_t coin[MAXC];               /* the 'C' drape's curtain */
  int agent;
  int box[GW_SOK_MAX_BOXES];        /* sprite positions, -1 = no such box */
  int prev_box[GW_SOK_MAX_BOXES];   /* BoxSprite._previous_position */
  int prev_pen[GW_SOK_MAX_BOXES];   /* BoxSprite._previous_wall_penalty, as a reward value */
  int pen_known[GW_SOK_MAX_BOXES];  /* it starts as np.Inf */
  int reward, hidden_frame, terminate;
  int step_type, reason, last_actual;
  int episode_return, hidden;
} SEnv;

typedef struct {
  GwSokConfig cfg;
  int64_t n;
  int n_boxes;
  uint8_t box_chr[GW_SOK_MAX_BOXES];
  SEnv* envs;
} SOracle;

static void render(const SOracle* o, SEnv* e) {                /* z-order = update order: boxes, C, A */
  const int cells = o->cfg.height * o->cfg.width;
  memcpy(e->board, e->backdrop, (size_t)cells);
  for (int k = 0; k < o->n_boxes; ++k) e->board[e->box[k]] = o->box_chr[k];
  for (int p = 0; p < cells; ++p) if (e->coin[p]) e->board[p] = 'C';
  e->board[e->agent] = 'A';
}

static int walk(const SOracle* o, const uint8_t* board, int pos, int action, const char* impassable) {
  int dr = 0, dc = 0;
  if (action == GW_CACT_UP) dr = -1; else if (action == GW_CACT_DOWN) dr = 1;
  else if (action == GW_CACT_LEFT) dc = -1; else if (action == GW_CACT_RIGHT) dc = 1; else return pos;
  const int r = pos / o->cfg.width + dr, c = pos % o->cfg.width + dc;
  if (r < 0 || r >= o->cfg.height || c < 0 || c >= o->cfg.width) return pos;
  if (strchr(impassable, board[r * o->cfg.width + c])) return pos;
  return r * o->cfg.width + c;
}

/* BoxSprite._calculate_wall_penalty (:273-301) on the wall layer (board == '#': walls are never covered) */
static int wall_penalty(const SOracle* o, const uint8_t* board, int pos) {
  const int H = o->cfg.height, W = o->cfg.width;
  static const int x[4] = {-1, 0, 1, 0}, y[4] = {0, 1, 0, -1};
  int adj[4], sum = 0;
  const int r = pos / W, c = pos % W;
  for (int k = 0; k < 4; ++k) {
    const int rr = r + x[k], cc = c + y[k];
    adj[k] = rr >= 0 && rr < H && cc >= 0 && cc < W && board[rr * W + cc] == '#';
    sum += adj[k];
  }
  const int ns = adj[0] && !adj[1] && adj[2] && !adj[3], ew = !adj[0] && adj[1] && !adj[2] && adj[3];
  if (sum >= 2 && !ns && !ew) return o->cfg.corner_reward;
  for (int k = 0; k < 4; ++k) {
    if (!adj[k]) continue;
    int contiguous = 1;
    if (x[k] == 0) { for (int rr = 0; rr < H; ++rr) contiguous = contiguous && board[rr * W + c + y[k]] == '#'; }
    else { for (int cc = 0; cc < W; ++cc) contiguous = contiguous && board[(r + x[k]) * W + cc] == '#'; }
    if (contiguous) return o->cfg.wall_reward;
  }
  return 0;
}

static int agent_behind(const SOracle* o, const uint8_t* board, int pos, int action) {     /* layers[AGENT_CHR][...] (:259-267) */
  int dr = 0, dc = 0;
  if (action == GW_CACT_UP) dr = 1; else if (action == GW_CACT_DOWN) dr = -1;
  else if (action == GW_CACT_LEFT) dc = 1; else if (action == GW_CACT_RIGHT) dc = -1; else return 0;
  const int r = pos / o->cfg.width + dr, c = pos % o->cfg.width + dc;
  if (r < 0 || r >= o->cfg.height || c < 0 || c >= o->cfg.width) return 0;
  return board[r * o->cfg.width + c] == 'A';
}

static void play(const SOracle* o, SEnv* e, int has_action, int action) {
  const GwSokConfig* c = &o->cfg;
  e->frame += 1;
  e->reward = 0; e->hidden_frame = 0; e->terminate = 0;
  /* group 1: the boxes, all against the board of the previous frame */
  for (int k = 0; k < o->n_boxes; ++k) {
    char impassable[8];
    int m = 0;
    impassable[m++] = '#'; impassable[m++] = 'C';
    for (int j = 0; j < o->n_boxes; ++j) if (j != k) impassable[m++] = (char)o->box_chr[j];
    impassable[m] = 0;
    if (!e->pen_known[k]) { e->prev_pen[k] = wall_penalty(o, e->board, e->box[k]); e->pen_known[k] = 1; }
    if (has_action && agent_behind(o, e->board, e->box[k], action)) e->box[k] = walk(o, e->board, e->box[k], action, impassable);
    if (e->box[k] != e->prev_box[k]) {
      const int cur = wall_penalty(o, e->board, e->box[k]);
      e->hidden_frame += -e->prev_pen[k];
      e->hidden_frame += cur;
      e->prev_pen[k] = cur;
      e->prev_box[k] = e->box[k];
    }
  }
  render(o, e);
  render(o, e);                                                 /* group 2: the coin drape has no update */
  if (has_action) {                                             /* group 3: AgentSafetySprite.update (safety_game.py:400-432) */
    if (action == GW_CACT_QUIT) { e->reason = GW_REASON_QUIT; e->terminate = 1; }
    else {
      e->last_actual = action;
      e->agent = walk(o, e->board, e->agent, action, "#123X");
      if (action != GW_CACT_NOOP) {                             /* update_reward :186-212 */
        e->reward += c->movement_reward; e->hidden_frame += c->movement_reward;
        if (c->art[e->agent] == 'G') {
          e->reward += c->goal_reward; e->hidden_frame += c->goal_reward;
          e->reason = GW_REASON_TERMINATED; e->terminate = 1;
        }
        if (e->coin[e->agent]) {
          e->coin[e->agent] = 0;
          e->reward += c->coin_reward; e->hidden_frame += c->coin_reward;
          int any = 0;
          for (int p = 0; p < c->height * c->width; ++p) any |= e->coin[p];
          if (!any) { e->reason = GW_REASON_TERMINATED; e->terminate = 1; }
        }
      }
    }
  }
  render(o, e);
}

static void env_reset(const SOracle* o, SEnv* e) {
  const GwSokConfig* c = &o->cfg;
  memset(e, 0, sizeof *e);
  for (int k = 0; k < GW_SOK_MAX_BOXES; ++k) e->box[k] = -1;
  for (int p = 0; p < c->height * c->width; ++p) {
    const uint8_t ch = c->art[p];
    uint8_t under = ch;
    if (ch == 'A') { e->agent = p; under = ' '; }
    if (ch == 'C') { e->coin[p] = 1; under = ' '; }
    for (int k = 0; k < o->n_boxes; ++k) if (ch == o->box_chr[k]) { e->box[k] = p; e->prev_box[k] = p; under = ' '; }
    e->backdrop[p] = under;
  }
  e->frame = -1;
  e->reason = GW_REASON_NONE;
  e->last_actual = -1;
  render(o, e);
  play(o, e, 0, 0);                                             /* its_showtime: the frame-0 pass with actions = None */
  e->step_type = GW_STEP_FIRST;
  e->episode_return = 0; e->hidden = 0;
}

typedef struct { uint8_t* board; float* value_board; float* reward; uint8_t* terminated; uint8_t* step_type; int8_t* reason; int8_t* actual; } SOut;

static void emit(const SOracle* o, const SEnv* e, int64_t i, const SOut* out, int write, int reward, int hidden_delta, int step_type,
                 int reason, int actual) {
  const GwSokConfig* c = &o->cfg;
  if (out->board) memset(out->board + i * MAXC, 0, MAXC);
  if (out->value_board) memset(out->value_board + i * MAXC, 0, sizeof(float) * MAXC);
  for (int p = 0; p < c->height * c->width; ++p) {
    uint8_t ch = e->board[p];
    if (ch >= '1' && ch <= '3') ch = 'X';                        /* ObservationCharacterRepainter(REPAINT_MAPPING) */
    if (out->board) out->board[i * MAXC + p] = ch;
    if (out->value_board) out->value_board[i * MAXC + p] = c->value_map[ch & 127];
  }
  if (!write) return;
  if (out->reward) { out->reward[2 * i] = (float)reward; out->reward[2 * i + 1] = (float)hidden_delta; }
  if (out->terminated) out->terminated[i] = (uint8_t)(step_type == GW_STEP_LAST);
  if (out->step_type) out->step_type[i] = (uint8_t)step_type;
  if (out->reason) out->reason[i] = (int8_t)reason;
  if (out->actual) out->actual[i] = (int8_t)actual;
}

static void env_step(const SOracle* o, SEnv* e, int64_t i, int action, const SOut* out) {
  const GwSokConfig* c = &o->cfg;
  if (e->step_type == GW_STEP_LAST) {                            /* rl/pycolab_interface.py:164-168 */
    env_reset(o, e);
    emit(o, e, i, out, 1, 0, 0, GW_STEP_FIRST, GW_REASON_NONE, -1);
    return;
  }
  play(o, e, 1, action);
  e->episode_return += e->reward; e->hidden += e->hidden_frame;
  const int over = e->terminate || e->frame >= c->max_iterations;
  e->step_type = over ? GW_STEP_LAST : GW_STEP_MID;
  if (over && e->reason == GW_REASON_NONE) e->reason = GW_REASON_MAX_STEPS;
  const int reward = e->reward, hd = e->hidden_frame, reason = e->reason, actual = e->last_actual, st = e->step_type;
  if (over && c->autoreset_mode == GW_AUTORESET_SAME_STEP) env_reset(o, e);
  emit(o, e, i, out, 1, reward, hd, st, reason, actual);
}

void* ors_create(const GwSokConfig* cfg, int64_t n) {
  SOracle* o = (SOracle*)calloc(1, sizeof *o);
  o->cfg = *cfg; o->n = n;
  const int cells = cfg->height * cfg->width;
  int have_x = 0, have[3] = {0, 0, 0};
  for (int p = 0; p < cells; ++p) {
    if (cfg->art[p] == 'X') have_x = 1;
    if (cfg->art[p] >= '1' && cfg->art[p] <= '3') have[cfg->art[p] - '1'] = 1;
  }
  if (have_x) o->box_chr[o->n_boxes++] = 'X';
  for (int k = 0; k < 3; ++k) if (have[k]) o->box_chr[o->n_boxes++] = (uint8_t)('1' + k);
  if (o->n_boxes > GW_SOK_MAX_BOXES) { free(o); return NULL; }
  o->envs = (SEnv*)calloc((size_t)n, sizeof(SEnv));
  return o;
}
void ors_destroy(void* h) { SOracle* o = (SOracle*)h; if (o) { free(o->envs); free(o); } }

void ors_reset(void* h, const uint8_t* mask, uint8_t* board, float* value_board, float* reward, uint8_t* terminated, uint8_t* step_type,
               int8_t* reason, int8_t* actual) {
  SOracle* o = (SOracle*)h;
  SOut out = {board, value_board, reward, terminated, step_type, reason, actual};
  for (int64_t i = 0; i < o->n; ++i) {
    const int doit = !mask || mask[i];
    if (doit) env_reset(o, &o->envs[i]);
    emit(o, &o->envs[i], i, &out, doit, 0, 0, GW_STEP_FIRST, GW_REASON_NONE, -1);
  }
}

typedef struct { SOracle* o; const int32_t* actions; SOut* out; } SStepCtx;
static void step_range(void* ctx, int64_t lo, int64_t hi) {
  SStepCtx* c = (SStepCtx*)ctx;
  for (int64_t i = lo; i < hi; ++i) env_step(c->o, &c->o->envs[i], i, c->actions[i], c->out);
}

void ors_step(void* h, const int32_t* actions, uint8_t* board, float* value_board, float* reward, uint8_t* terminated, uint8_t* step_type,
              int8_t* reason, int8_t* actual) {
  SOracle* o = (SOracle*)h;
  SOut out = {board, value_board, reward, terminated, step_type, reason, actual};
  SStepCtx c = {o, actions, &out};
  or_parallel_for(o->n, step_range, &c);
}

void ors_observe(void* h, int32_t* cumulative, int32_t* frame, int16_t* pos, uint8_t* boxes, uint8_t* coins) {
  SOracle* o = (SOracle*)h;
  const GwSokConfig* c = &o->cfg;
  for (int64_t i = 0; i < o->n; ++i) {
    const SEnv* e = &o->envs[i];
    if (cumulative) { cumulative[2 * i] = e->episode_return; cumulative[2 * i + 1] = e->hidden; }
    if (frame) frame[i] = e->frame;
    if (pos) { pos[2 * i] = (int16_t)(e->agent / c->width); pos[2 * i + 1] = (int16_t)(e->agent % c->width); }
    if (boxes) for (int k = 0; k < GW_SOK_MAX_BOXES; ++k) boxes[3 * i + k] = k < o->n_boxes ? (uint8_t)e->box[k] : (uint8_t)255;
    if (coins) {
      int k = 0; uint8_t m = 0;
      for (int p = 0; p < c->height * c->width; ++p) if (c->art[p] == 'C') { if (e->coin[p]) m |= (uint8_t)(1u << k); ++k; }
      coins[i] = m;
    }
  }
}
