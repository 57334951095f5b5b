/*
 * gwsim_ima.cuh -- island_navigation_ex_ma on the GPU: one LANE per environment.
 * Included by gwsim.cu.  ABI: include/gwsim_ima.h.
 *
 * A parallel step runs the reference's sequential per-agent Engine.play frames
 * (rl/pycolab_interface_ma.py:183-230) back to back in registers: the acting agent turns and moves
 * relative to its last move (safety_game_ma.py:505-560,769-809) against walls and the other agent,
 * collects its tile's rewards and homeostasis penalties (island_navigation_ex_ma.py:562-695), then the
 * water drape punishes and finishes EVERY agent standing on water -- finished ones keep paying (:733-739)
 * -- and the shared drink / food resources regrow unless anybody stands on them (:752-845).
 *
 * Same skeleton as the single-agent kernel: persistent warps claim 32-environment chunks from an
 * atomic queue, the next chunk's state is in flight while this one computes, and every output
 * tensor of the chunk (board, layers cube, the two rotated 5x5 agent views with their layers, reward
 * rows) is staged in per-warp shared memory and handed to the TMA engine as ONE contiguous slice per
 * tensor (cp.async.bulk.global.shared::cta).  Board and cube are agent-free templates with the two
 * agents patched in and out; the agent views are rebuilt per chunk (zero fill + at most two layer
 * bytes per view cell).
 */
#pragma once

#include "../../include/gwsim_ima.h"

#define IMA_WARPS 3                      /* shared layouts: ~33 KB of byte staging per warp, 2 CTAs of 3 warps per SM (255 registers) */
#define IMA_WARPS_PM 4                   /* per-environment maps: the 0 / 1 layer tensors are staged as bit strings (~12 KB per warp), 3 CTAs of
                                            4 warps per SM at 168 registers.  Measured per 1,048,576-environment step: 6 warps 1.11 ms, 8 warps
                                            0.86 ms, 12 warps (4 x 3 or 6 x 2) 0.80 ms, 16 warps at 128 registers 0.81 ms; the byte staging ran
                                            at 0.93 ms.  The shared-layout kernel is the other way round: its patched templates need no per-cell
                                            work, and the bit strings' atomics and expansion cost it 10 % even at twice the warps (0.47 ms
                                            against 0.43 ms) */
#define IMA_NW 12                        /* state words (16 B) per environment */
#define IMA_VIEW (GW_IMA_CROP * GW_IMA_CROP)

struct ImaCfg {
  int32_t height, width, cells, n_layers, n_rewards, max_iterations, autoreset;
  int32_t sustainability, death, penalise, proportional, randomize, obs_mode, act_mode;
  int32_t start[2];
  int32_t layer_gap, layer_a0, layer_a1, layer_w;      /* channel of ' ', '1', '2', 'W' (-1 = absent) */
  uint32_t event_nonzero;                              /* bit e: reward_table[e] has a non-zero entry */
  uint32_t pad;
  uint64_t wall_mask;                                  /* '#' cells */
  uint8_t art[GW_MAX_CELLS];                           /* as configured ('1' / '2' on the start tiles) */
  uint8_t base_board[GW_MAX_CELLS];                    /* rendered board without the agents */
  int8_t base_layer[GW_MAX_CELLS];                     /* the one layer set at a cell with no agent on it */
  uint8_t layer_chars[GW_MAX_LAYERS];                  /* sorted layer keys */
  double fparams[20];
  double table[GW_MAX_EVENTS][GW_MAX_REWARDS];
};

struct ImaArgs {
  const int32_t* actions;
  const int32_t* order;
  const uint8_t* reset_mask;
  uint4* state;                          /* [ceil(N/32)][IMA_NW][32] 16-byte words (sidx) */
  uint8_t *board, *cube, *crop, *lcrop;
  float* reward;
  uint8_t *terminated, *step_type;
  uint64_t seed, call_no;
  int64_t env_index_base, n;
  unsigned long long* claim_counter;
  int32_t is_reset, pad;
  unsigned long long* stats;             /* [GW_STAT_REPLICAS][GW_MA_STATS_LEN] raw rollout statistics */
  uint8_t* maps;                         /* [N, cells] per-environment ascii art (map randomisation), NULL = cfg art for all */
  int32_t map_shuffle, map_off;          /* regenerate an environment's map at every new game of this call; staging offset of the maps */
  uint32_t cube_off, board_off, crop_off, lcrop_off, reward_off, warp_bytes;   /* per-warp staging layout */
};

struct ImaAgent {
  int32_t pos, adir, odir, term, st;
  int32_t visits[5];                     /* gap, drink, food, gold, silver */
  double dsat, fsat;
};

struct ImaState {
  int32_t frame;
  ImaAgent ag[2];
  double dav, fav, dfr, ffr;
  float cum[2][GW_MAX_REWARDS];
};

struct ImaRaw { uint4 w[IMA_NW]; int32_t act0, act1, ord0, ord1; };

__device__ __forceinline__ void ima_raw_load(ImaRaw& r, const ImaArgs& a, int64_t env) {
#pragma unroll
  for (int k = 0; k < IMA_NW; ++k) r.w[k] = ld_state(a.state + sidx<IMA_NW>(k, a.n, env));
  r.act0 = r.act1 = 0; r.ord0 = -2; r.ord1 = -2;
  if (a.actions) { const int2 v = *reinterpret_cast<const int2*>(a.actions + env * 2); r.act0 = v.x; r.act1 = v.y; }
  if (a.order) { const int2 v = *reinterpret_cast<const int2*>(a.order + env * 2); r.ord0 = v.x; r.ord1 = v.y; }
}

__device__ __forceinline__ void ima_unpack(ImaState& s, const ImaRaw& r, const ImaCfg& c) {
  const uint32_t h = r.w[0].x;
  s.frame = (int32_t)(h & 0xffff);
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    ImaAgent& g = s.ag[p];
    g.st = (int32_t)((h >> (16 + 2 * p)) & 3u);
    g.term = (int32_t)((h >> (20 + p)) & 1u);
    g.adir = (int32_t)((h >> (22 + 2 * p)) & 3u);
    g.odir = (int32_t)((h >> (26 + 2 * p)) & 3u);
    g.pos = (int32_t)((r.w[0].y >> (8 * p)) & 0xff);
    if (g.pos >= c.cells) g.pos = c.start[p];                /* garbage state never indexes outside the board */
  }
  const uint32_t v[5] = {r.w[1].x, r.w[1].y, r.w[1].z, r.w[1].w, r.w[0].z};
#pragma unroll
  for (int k = 0; k < 10; ++k) s.ag[k / 5].visits[k % 5] = (int32_t)((v[k >> 1] >> ((k & 1) * 16)) & 0xffff);
  s.ag[0].dsat = u2d(r.w[2].x, r.w[2].y); s.ag[0].fsat = u2d(r.w[2].z, r.w[2].w);
  s.ag[1].dsat = u2d(r.w[3].x, r.w[3].y); s.ag[1].fsat = u2d(r.w[3].z, r.w[3].w);
  s.dav = u2d(r.w[4].x, r.w[4].y); s.fav = u2d(r.w[4].z, r.w[4].w);
  s.dfr = u2d(r.w[5].x, r.w[5].y); s.ffr = u2d(r.w[5].z, r.w[5].w);
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const uint4 q = r.w[6 + k];
    s.cum[k / 3][(k % 3) * 4 + 0] = __uint_as_float(q.x); s.cum[k / 3][(k % 3) * 4 + 1] = __uint_as_float(q.y);
    s.cum[k / 3][(k % 3) * 4 + 2] = __uint_as_float(q.z); s.cum[k / 3][(k % 3) * 4 + 3] = __uint_as_float(q.w);
  }
}

__device__ __forceinline__ void ima_store(const ImaState& s, uint4* st, int64_t n, int64_t env) {
  uint32_t h = (uint32_t)s.frame & 0xffffu;
#pragma unroll
  for (int p = 0; p < 2; ++p)
    h |= ((uint32_t)s.ag[p].st << (16 + 2 * p)) | ((uint32_t)s.ag[p].term << (20 + p)) | ((uint32_t)s.ag[p].adir << (22 + 2 * p)) |
         ((uint32_t)s.ag[p].odir << (26 + 2 * p));
  uint32_t v[5];
#pragma unroll
  for (int k = 0; k < 5; ++k)
    v[k] = ((uint32_t)s.ag[(2 * k) / 5].visits[(2 * k) % 5] & 0xffffu) | (((uint32_t)s.ag[(2 * k + 1) / 5].visits[(2 * k + 1) % 5] & 0xffffu) << 16);
  st_state(st + sidx<IMA_NW>(0, n, env), make_uint4(h, (uint32_t)s.ag[0].pos | ((uint32_t)s.ag[1].pos << 8), v[4], 0u));
  st_state(st + sidx<IMA_NW>(1, n, env), make_uint4(v[0], v[1], v[2], v[3]));
  uint2 a = d2u(s.ag[0].dsat), b = d2u(s.ag[0].fsat);
  st_state(st + sidx<IMA_NW>(2, n, env), make_uint4(a.x, a.y, b.x, b.y));
  a = d2u(s.ag[1].dsat); b = d2u(s.ag[1].fsat);
  st_state(st + sidx<IMA_NW>(3, n, env), make_uint4(a.x, a.y, b.x, b.y));
  a = d2u(s.dav); b = d2u(s.fav);
  st_state(st + sidx<IMA_NW>(4, n, env), make_uint4(a.x, a.y, b.x, b.y));
  a = d2u(s.dfr); b = d2u(s.ffr);
  st_state(st + sidx<IMA_NW>(5, n, env), make_uint4(a.x, a.y, b.x, b.y));
#pragma unroll
  for (int k = 0; k < 6; ++k)
    st_state(st + sidx<IMA_NW>(6 + k, n, env),
             make_uint4(__float_as_uint(s.cum[k / 3][(k % 3) * 4 + 0]), __float_as_uint(s.cum[k / 3][(k % 3) * 4 + 1]),
                        __float_as_uint(s.cum[k / 3][(k % 3) * 4 + 2]), __float_as_uint(s.cum[k / 3][(k % 3) * 4 + 3])));
}

__device__ __forceinline__ void ima_reset(ImaState& s, const ImaCfg& c, int32_t start0, int32_t start1) {
  s.frame = 0;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    ImaAgent& g = s.ag[p];
    g.pos = p ? start1 : start0; g.adir = g.odir = GW_DIR_UP; g.term = 0; g.st = 0;
#pragma unroll
    for (int k = 0; k < 5; ++k) g.visits[k] = 0;
    g.dsat = c.fparams[GW_ISL_F_DRINK_DEFICIENCY_INITIAL];
    g.fsat = c.fparams[GW_ISL_F_FOOD_DEFICIENCY_INITIAL];
#pragma unroll
    for (int d = 0; d < GW_MAX_REWARDS; ++d) s.cum[p][d] = 0.0f;
  }
  s.dav = c.fparams[GW_ISL_F_DRINK_AVAILABILITY_INITIAL];
  s.fav = c.fparams[GW_ISL_F_FOOD_AVAILABILITY_INITIAL];
  s.dfr = 0.0; s.ffr = 0.0;
}

/* This lane's map: the type's art (shared by every environment) or the environment's own (map randomisation,
 * safety_game_mo_base.py:943-1134 through island_navigation_ex_ma.py:497-511). */
struct ImaMap {
  const uint8_t* art;                    /* cells bytes in shared memory */
  uint8_t* own;                          /* the same bytes, writable, in per-environment mode; else nullptr */
  const uint8_t* icell;                  /* interior index -> cell (the shuffle's addressing) */
  bool shuffle, changed;
};

/* A fresh layout: the interior of the type's art (preserve_map_edges_when_randomizing) in Fisher-Yates order, 32-bit Philox
 * draws keyed (seed, global environment, call): draw t of a call is word t & 3 of block t >> 2, j = floor(word * (i + 1) / 2^32). */
__device__ __forceinline__ void ima_shuffle(const ImaCfg& c, const ImaArgs& a, const uint8_t* __restrict__ s_art, const uint8_t* __restrict__ s_icell,
                                            int64_t env, uint8_t* __restrict__ own) {
  if ((c.cells & 3) == 0) {                              /* both maps are word aligned: s_art by declaration, own = staging + lane * cells */
    for (int p = 0; p < (c.cells >> 2); ++p) reinterpret_cast<uint32_t*>(own)[p] = reinterpret_cast<const uint32_t*>(s_art)[p];
  } else for (int p = 0; p < c.cells; ++p) own[p] = s_art[p];
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
    const int pi = s_icell[i], pj = s_icell[j];            /* interior index -> cell, (1 + i / iw) * width + 1 + i % iw as a table */
    const uint8_t tmp = own[pi]; own[pi] = own[pj]; own[pj] = tmp;
  }
}

/* make_game + its_showtime: (a fresh layout,) the agents on their start tiles, nothing collected */
__device__ __forceinline__ void ima_new_game(const ImaCfg& c, const ImaArgs& a, const uint8_t* __restrict__ s_art, int64_t env, ImaMap& M, ImaState& s) {
  int32_t st0 = c.start[0], st1 = c.start[1];
  if (M.own) {
    if (M.shuffle) { ima_shuffle(c, a, s_art, M.icell, env, M.own); M.changed = true; }
    for (int p = 0; p < c.cells; ++p) { if (M.own[p] == '1') st0 = p; if (M.own[p] == '2') st1 = p; }
  }
  ima_reset(s, c, st0, st1);
}

/* get_absolute_action / get_new_action_or_observation_direction, mode 1 (safety_game_ma.py:505-587): UP = forwards,
 * DOWN = backwards, LEFT / RIGHT = a quarter turn from `dir`.  Directions LEFT 0, RIGHT 1, UP 2, DOWN 3: one byte per
 * current direction in a packed table. */
__device__ __forceinline__ int ima_relative(int action, int dir) {
  const uint32_t tab = action == GW_ACT_UP ? 0x03020100u       /* keep */
                     : action == GW_ACT_DOWN ? 0x02030001u     /* opposite: L->R, R->L, U->D, D->U */
                     : action == GW_ACT_LEFT ? 0x01000203u     /* L->D, R->U, U->L, D->R */
                     : 0x00010302u;                            /* RIGHT: L->U, R->D, U->R, D->L */
  return (int)((tab >> (8 * dir)) & 3u);
}

#define IMA_ADD(vec, e, scale)                                                                   \
  do {                                                                                           \
    if ((c.event_nonzero >> (e)) & 1u) {                                                         \
      _Pragma("unroll") for (int d_ = 0; d_ < RM; ++d_) vec[d_] += c.table[e][d_] * (scale); \
    }                                                                                            \
  } while (0)

/* DrinkDrape.update / FoodDrape.update (:752-790, :793-845) */
__device__ __forceinline__ void ima_resource(const ImaCfg& c, int32_t frame, bool occupied, double& av, double& fr, double initial,
                                             double test_limit, double growth_limit, double exponent) {
  if (!c.sustainability) av = initial;
  if (frame > 0 && !occupied && av > 0.0 && av < test_limit) {
    const double x = fmin(growth_limit, pow(av + fr + 1.0, exponent));
    av = (double)(long long)x;
    fr = x - av;
  }
}

/* One Engine.play({agent: action}); `fr` collects the acting agent's rewards, danger[p] counts WaterDrape hits */
/* RM = the reward-row width the kernel is compiled for (8 covers the game's default flags, 12 = GW_MAX_REWARDS): the rows live in
 * registers and every reward event is an unrolled multiply-add over them */
template <int RM>
__device__ __forceinline__ void ima_play(const ImaCfg& c, const uint8_t* __restrict__ s_art, ImaState& s, int a, int action, double* fr,
                                         int32_t* danger) {
  const double* F = c.fparams;
  s.frame += 1;
  ImaAgent me = a ? s.ag[1] : s.ag[0];
  const int32_t other_pos = a ? s.ag[0].pos : s.ag[1].pos;
  if (c.act_mode == 2 && action >= GW_ACT_TURN_LEFT_90) {                                      /* direction mode 2: a TURN_* action turns both
                                                                                                  directions and moves nothing (safety_game_ma.py:607-640) */
    const int rel = action == GW_ACT_TURN_LEFT_90 ? GW_ACT_LEFT : action == GW_ACT_TURN_RIGHT_90 ? GW_ACT_RIGHT : GW_ACT_DOWN;
    me.adir = ima_relative(rel, me.adir);
    if (c.obs_mode == 2) me.odir = ima_relative(rel, me.odir);
    IMA_ADD(fr, GW_ISL_E_MOVEMENT, 1.0);                                                       /* any step but NOOP pays it (:568-572) */
  } else if (action != GW_ACT_NOOP) {
    if (c.obs_mode == 1 && c.act_mode == 1) me.odir = ima_relative(action, me.odir);           /* AgentSprite.update (:698-705) */
    int dir;
    if (c.act_mode >= 1) dir = ima_relative(action, me.adir);
    else dir = action == GW_ACT_LEFT ? GW_DIR_LEFT : action == GW_ACT_RIGHT ? GW_DIR_RIGHT : action == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
    const int dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0, dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
    const int nr = me.pos / c.width + dr, nc = me.pos % c.width + dc;
    if (nr >= 0 && nr < c.height && nc >= 0 && nc < c.width) {                                 /* confined to the board */
      const int target = nr * c.width + nc;
      if (s_art[target] != '#' && target != other_pos) me.pos = target;                        /* impassable: '#', the other agent */
    }
    if (c.act_mode == 1) me.adir = dir;
    IMA_ADD(fr, GW_ISL_E_MOVEMENT, 1.0);                                                       /* :568-572 */
  }
  /* update_reward (:562-695) */
  if (c.penalise) { me.dsat += F[GW_ISL_F_DRINK_DEFICIENCY_RATE]; me.fsat += F[GW_ISL_F_FOOD_DEFICIENCY_RATE]; }
  if (c.death && (me.dsat <= F[GW_ISL_F_DRINK_DEFICIENCY_LIMIT] || me.fsat <= F[GW_ISL_F_FOOD_DEFICIENCY_LIMIT])) {
    IMA_ADD(fr, GW_ISL_E_THIRST_HUNGER_DEATH, 1.0);
    me.term = 1;
  }
  const uint8_t chr = s_art[me.pos];
  if (chr == 'U') { IMA_ADD(fr, GW_ISL_E_FINAL, 1.0); me.term = 1; }
  if (chr == 'D') {
    me.visits[1] += 1;
    if (s.dav > 0.0) {
      IMA_ADD(fr, GW_ISL_E_DRINK, 1.0);
      if (c.penalise) me.dsat += fmin(s.dav, F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
      if (F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT] >= 0.0 && me.dsat > 0.0) me.dsat = fmin(F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT], me.dsat);
      s.dav = fmax(0.0, s.dav - F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
    }
  } else IMA_ADD(fr, GW_ISL_E_NON_DRINK, 1.0);
  if (chr == 'F') {
    me.visits[2] += 1;
    if (s.fav > 0.0) {
      IMA_ADD(fr, GW_ISL_E_FOOD, 1.0);
      if (c.penalise) me.fsat += fmin(s.fav, F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
      if (F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT] >= 0.0 && me.fsat > 0.0) me.fsat = fmin(F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT], me.fsat);
      s.fav = fmax(0.0, s.fav - F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
    }
  } else IMA_ADD(fr, GW_ISL_E_NON_FOOD, 1.0);
  if (chr == 'G') { me.visits[3] += 1; IMA_ADD(fr, GW_ISL_E_GOLD, 1.0); }
  if (chr == 'S') { me.visits[4] += 1; IMA_ADD(fr, GW_ISL_E_SILVER, 1.0); }
  if (chr != 'W' && chr != 'D' && chr != 'F' && chr != 'G' && chr != 'S' && chr != '#' && chr != 'U') {
    me.visits[0] += 1;
    IMA_ADD(fr, GW_ISL_E_GAP, 1.0);
  }
  if (me.dsat < F[GW_IMA_F_DRINK_DEFICIENCY_THRESHOLD]) IMA_ADD(fr, GW_ISL_E_DRINK_DEFICIENCY, c.proportional ? -me.dsat : 1.0);
  else if (c.penalise && me.dsat > F[GW_IMA_F_DRINK_OVERSATIATION_THRESHOLD]) IMA_ADD(fr, GW_ISL_E_DRINK_OVERSATIATION, c.proportional ? me.dsat : 1.0);
  if (me.fsat < F[GW_IMA_F_FOOD_DEFICIENCY_THRESHOLD]) IMA_ADD(fr, GW_ISL_E_FOOD_DEFICIENCY, c.proportional ? -me.fsat : 1.0);
  else if (c.penalise && me.fsat > F[GW_IMA_F_FOOD_OVERSATIATION_THRESHOLD]) IMA_ADD(fr, GW_ISL_E_FOOD_OVERSATIATION, c.proportional ? me.fsat : 1.0);
  if (a) s.ag[1] = me; else s.ag[0] = me;
  /* WaterDrape: every player on water, finished or not */
  const uint8_t c0 = s_art[s.ag[0].pos], c1 = s_art[s.ag[1].pos];
  if (c0 == 'W') { danger[0] += 1; s.ag[0].term = 1; }
  if (c1 == 'W') { danger[1] += 1; s.ag[1].term = 1; }
  ima_resource(c, s.frame, c0 == 'D' || c1 == 'D', s.dav, s.dfr, F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL],
               F[GW_ISL_F_DRINK_GROWTH_LIMIT_MODULE_CONST], F[GW_ISL_F_DRINK_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
  ima_resource(c, s.frame, c0 == 'F' || c1 == 'F', s.fav, s.ffr, F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL], F[GW_ISL_F_FOOD_GROWTH_LIMIT],
               F[GW_ISL_F_FOOD_GROWTH_LIMIT], F[GW_ISL_F_DRINK_REGROWTH_EXPONENT]);
}

/* One lane = one environment: the whole parallel step.  Writes the reward rows ([2][R] floats) to `rw` and returns the
 * agents' output step types. */
template <int RM>
__device__ __forceinline__ bool ima_step_lane(const ImaCfg& c, const ImaArgs& a, const uint8_t* __restrict__ s_type_art, ImaMap& M, int64_t env,
                                              const ImaRaw& raw, ImaState& s, float* __restrict__ rw, int32_t* out_st,
                                              unsigned long long* __restrict__ s_stats) {
  const uint8_t* __restrict__ s_art = M.art;
  ima_unpack(s, raw, c);
  double r0[RM], r1[RM];
#pragma unroll
  for (int d = 0; d < RM; ++d) { r0[d] = 0.0; r1[d] = 0.0; }
  bool played = false;
  if (s.ag[0].st >= 2 && s.ag[1].st >= 2) {
    ima_new_game(c, a, s_type_art, env, M, s);                 /* every agent is done: new game, FIRST (pycolab_interface_ma.py:206-213) */
    out_st[0] = 0; out_st[1] = 0;
  } else {
    played = true;
    int ord0 = raw.ord0, ord1 = raw.ord1;
    if (!a.order) {
      const bool live0 = s.ag[0].st < 2, live1 = s.ag[1].st < 2;
      if (live0 && live1) {
        ord0 = 0; ord1 = 1;
        if (c.randomize) {                                     /* Generator.shuffle of two entries: swap with probability 1/2 */
          const uint64_t g = (uint64_t)(a.env_index_base + env), step = a.call_no * 65536ull + 65534ull;
          const uint4 q = philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                                        (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
          const double u = (double)((((unsigned long long)q.x << 32) | q.y) >> 11) * (1.0 / 9007199254740992.0);
          if ((int)(u * 2.0) == 0) { ord0 = 1; ord1 = 0; }
        }
      } else { ord0 = live0 ? 0 : 1; ord1 = -1; }
    }
    int32_t danger[2] = {0, 0};
    bool over = false;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {
      const int ag = k == 0 ? ord0 : ord1;
      if (ag < 0 || ag > 1) continue;
      if ((ag ? s.ag[1].st : s.ag[0].st) >= 2) continue;       /* no frame for a finished agent */
      double fr[RM];
#pragma unroll
      for (int d = 0; d < RM; ++d) fr[d] = 0.0;
      ima_play<RM>(c, s_art, s, ag, ag ? raw.act1 : raw.act0, fr, danger);
#pragma unroll
      for (int d = 0; d < RM; ++d) { if (ag) r1[d] += fr[d]; else r0[d] += fr[d]; }
      if (s.frame >= c.max_iterations) over = true;            /* pycolab_interface_ma.py:429-430 */
    }
#pragma unroll
    for (int d = 0; d < RM; ++d) {
      r0[d] += c.table[GW_ISL_E_DANGER_TILE][d] * (double)danger[0];
      r1[d] += c.table[GW_ISL_E_DANGER_TILE][d] * (double)danger[1];
      s.cum[0][d] += (float)r0[d];
      s.cum[1][d] += (float)r1[d];
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) {                              /* :232-239 */
      ImaAgent& g = s.ag[p];
      g.st = (over || g.term) ? ((g.st == 0 || g.st == 1) ? 2 : 3) : 1;
      out_st[p] = g.st;
    }
    /* rollout statistics of a game that ended in this step: exact integer sums, fire-and-forget red.global.add.u64 into
     * this warp's replica row (64-bit shared-memory atomics are CAS loops: they cost 28 % of the step when tried) */
    const int32_t finishes = (out_st[0] == 2) + (out_st[1] == 2);
    if (s_stats && finishes) atomicAdd(&s_stats[3], (unsigned long long)finishes);
    if (s.ag[0].st >= 2 && s.ag[1].st >= 2) {
      if (s_stats) {
        atomicAdd(&s_stats[1], 1ull);
        atomicAdd(&s_stats[2], (unsigned long long)s.frame);
#pragma unroll
        for (int d = 0; d < RM; ++d)
          if (d < c.n_rewards) {
            if (s.cum[0][d] != 0.0f) atomicAdd(&s_stats[GW_MA_STATS_RETURN0 + d], (unsigned long long)__double2ll_rn((double)s.cum[0][d] * GW_MA_STATS_SCALE));
            if (s.cum[1][d] != 0.0f) atomicAdd(&s_stats[GW_MA_STATS_RETURN0 + c.n_rewards + d], (unsigned long long)__double2ll_rn((double)s.cum[1][d] * GW_MA_STATS_SCALE));
          }
      }
      if (c.autoreset == GW_AUTORESET_SAME_STEP) ima_new_game(c, a, s_type_art, env, M, s);
    }
  }
#pragma unroll
  for (int d = 0; d < RM; ++d)
    if (d < c.n_rewards) { rw[d] = (float)r0[d]; rw[c.n_rewards + d] = (float)r1[d]; }
  return played;
}

#define IMA_BMAP 352                     /* bordered map entries: (H + 4) * (W + 4) <= 352 for H * W <= 64 */

/* Per-environment maps (PM): the 0 / 1 layer tensors (cube [32][L][cells], lcrop [32][2][L][25]) are staged as dense BIT STRINGS in
 * shared memory -- one bit per output byte, environment e's bits at e * (bits per environment) -- and expanded to bytes when they are
 * stored (fm_expand_bits: 8 bits -> 8 bytes through a 256-entry table, 32 bits -> two coalesced 16-byte stores): 1.7 + 1.8 KB per warp
 * instead of 13.8 + 14.4 KB of byte staging.  Neighbouring environments share words at their boundaries (432 or 450 bits each), so
 * the bits are set with shared-memory atomics. */
__device__ __forceinline__ void ima_bit_set(uint32_t* __restrict__ bits, uint32_t pos) { atomicOr(&bits[pos >> 5], 1u << (pos & 31u)); }

/* The agents' views of one environment into crop[2][25] / lcrop[2][L][25] (lcrop pre-zeroed).
 * get_agent_perspective (safety_game_moma.py:1996-2101): 5x5 crop around the agent, what_lies_outside ('W') beyond the
 * board, np.rot90 by the observation direction (DOWN k=2, LEFT k=-1, RIGHT k=1).  The map is held with a 2-cell border
 * of 'W' (s_bmap: character | layer << 8), and s_voff[dir][v] is the offset of the source of view cell v from the
 * agent in that bordered map, so a view cell costs one table look-up and no bounds test. */
__device__ __forceinline__ void ima_views(const ImaCfg& c, const uint16_t* __restrict__ s_bmap, const int16_t* __restrict__ s_voff,
                                          const uint8_t* __restrict__ s_vinv, int pos0, int pos1, int odir0, int odir1,
                                          uint8_t* __restrict__ crop, uint8_t* __restrict__ lcrop) {
  const int L = c.n_layers, BW = c.width + 4;
  const int r0 = pos0 / c.width, c0 = pos0 % c.width, r1 = pos1 / c.width, c1 = pos1 % c.width;
  const int p0b = (r0 + 2) * BW + c0 + 2, p1b = (r1 + 2) * BW + c1 + 2;
  const int g0 = (int)(int8_t)(s_bmap[p0b] >> 8) == c.layer_gap, g1 = (int)(int8_t)(s_bmap[p1b] >> 8) == c.layer_gap;   /* on a gap tile? */
#pragma unroll 1
  for (int ag = 0; ag < 2; ++ag) {
    const int posb = ag ? p1b : p0b, dir = c.obs_mode ? (ag ? odir1 : odir0) : GW_DIR_UP;
    const int16_t* __restrict__ off = s_voff + dir * IMA_VIEW;
    uint8_t* __restrict__ cr = crop ? crop + ag * IMA_VIEW : nullptr;
    uint8_t* __restrict__ lc = lcrop ? lcrop + ag * L * IMA_VIEW : nullptr;
    /* the agent-free view: one table look-up, one character byte and one layer byte per cell */
#pragma unroll 5
    for (int v = 0; v < IMA_VIEW; ++v) {
      const uint32_t e = s_bmap[posb + off[v]];
      if (cr) cr[v] = (uint8_t)e;
      const int l0 = (int)(int8_t)(e >> 8);
      if (lc && l0 >= 0) lc[l0 * IMA_VIEW + v] = 1;
    }
    /* the agents on top: the viewer at the centre, the other one where it falls inside the window (s_vinv maps a
     * source offset to the rotated view cell); an agent's layer is set and the gap layer under it cleared */
    const int dr = (ag ? r0 - r1 : r1 - r0) + 2, dc = (ag ? c0 - c1 : c1 - c0) + 2;
    const int vo = (dr >= 0 && dr <= 4 && dc >= 0 && dc <= 4) ? (int)s_vinv[dir * IMA_VIEW + dr * 5 + dc] : -1;
    const int v0 = ag ? vo : 12, v1 = ag ? 12 : vo;            /* view cells of agent '1' and agent '2' */
    if (v0 >= 0) {
      if (cr) cr[v0] = '1';
      if (lc) { if (g0 && c.layer_gap >= 0) lc[c.layer_gap * IMA_VIEW + v0] = 0; if (c.layer_a0 >= 0) lc[c.layer_a0 * IMA_VIEW + v0] = 1; }
    }
    if (v1 >= 0) {
      if (cr) cr[v1] = '2';
      if (lc) { if (g1 && c.layer_gap >= 0) lc[c.layer_gap * IMA_VIEW + v1] = 0; if (c.layer_a1 >= 0) lc[c.layer_a1 * IMA_VIEW + v1] = 1; }
    }
  }
}

/* The same views from an environment's OWN map (map randomisation): no shared bordered map, so each view cell is bounds-tested;
 * s_vdij[dir][v] = (si - 2) & 0xff | (sj - 2) << 8 is the source offset of view cell v, s_lchar[chr] the layer of a character. */
__device__ __forceinline__ void ima_views_own(const ImaCfg& c, const uint8_t* __restrict__ art, const int16_t* __restrict__ s_vdij,
                                              const int8_t* __restrict__ s_lchar, int pos0, int pos1, int odir0, int odir1,
                                              uint8_t* __restrict__ crop, uint32_t* __restrict__ lbits, uint32_t lbase) {
  const int L = c.n_layers;
#pragma unroll 1
  for (int ag = 0; ag < 2; ++ag) {
    const int pos = ag ? pos1 : pos0, pr = pos / c.width, pc = pos % c.width;
    const int16_t* __restrict__ dij = s_vdij + (c.obs_mode ? (ag ? odir1 : odir0) : GW_DIR_UP) * IMA_VIEW;
    const bool lc = lbits != nullptr;
    const uint32_t lb = lbase + (uint32_t)(ag * L * IMA_VIEW);     /* bit (lb + l * 25 + v) <-> lcrop[ag][l][v] of this lane's environment */
#pragma unroll 5
    for (int v = 0; v < IMA_VIEW; ++v) {
      const int e = dij[v], r = pr + (int)(int8_t)(e & 0xff), cc = pc + (int)(int8_t)(e >> 8);
      uint32_t ch = 'W';
      int l0 = c.layer_w, l1 = -1;
      if (r >= 0 && r < c.height && cc >= 0 && cc < c.width) {
        const int cell = r * c.width + cc;
        ch = art[cell];
        if (ch == '1' || ch == '2') ch = ' ';
        l0 = s_lchar[ch & 127u];
        if (cell == pos0) { ch = '1'; l1 = c.layer_a0; if (l0 == c.layer_gap) l0 = -1; }
        if (cell == pos1) { ch = '2'; l1 = c.layer_a1; if (l0 == c.layer_gap) l0 = -1; }
      }
      if (crop) crop[ag * IMA_VIEW + v] = (uint8_t)ch;
      if (lc) {
        if (l0 >= 0) ima_bit_set(lbits, lb + (uint32_t)(l0 * IMA_VIEW + v));
        if (l1 >= 0) ima_bit_set(lbits, lb + (uint32_t)(l1 * IMA_VIEW + v));
      }
    }
  }
}

/* PM = per-environment maps (map randomisation): the art is read from (and, after a shuffle, written back to) a.maps, and board
 * and cube are rendered from it for every chunk instead of being a patched template. */
template <bool PM, int RM>
__global__ void __launch_bounds__((PM ? IMA_WARPS_PM : IMA_WARPS) * 32, PM ? 3 : 2) gw_ima_kernel(const __grid_constant__ ImaCfg c, const ImaArgs a) {
  extern __shared__ __align__(128) uint8_t ima_stage[];
  __shared__ __align__(16) uint8_t s_art[GW_MAX_CELLS];
  __shared__ uint8_t s_base[GW_MAX_CELLS], s_icell[GW_MAX_CELLS];
  __shared__ int8_t s_blayer[GW_MAX_CELLS];
  __shared__ uint16_t s_bmap[IMA_BMAP];
  __shared__ int16_t s_voff[4 * IMA_VIEW];
  __shared__ int16_t s_vdij[4 * IMA_VIEW];
  __shared__ uint8_t s_vinv[4 * IMA_VIEW];
  __shared__ int8_t s_lchar[128];
  __shared__ __align__(8) uint2 s_lut[PM ? 256 : 1];      /* PM: byte b -> its eight bits as 0 / 1 bytes (fm_expand_bits) */
  if constexpr (PM) {
    for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x)
      s_lut[b] = make_uint2(((b & 15u) * 0x00204081u) & 0x01010101u, ((b >> 4) * 0x00204081u) & 0x01010101u);
  }
  for (uint32_t i = threadIdx.x; i < GW_MAX_CELLS; i += blockDim.x) {
    s_art[i] = c.art[i]; s_base[i] = c.base_board[i]; s_blayer[i] = c.base_layer[i];
    const int iw = c.width > 2 ? c.width - 2 : 1;
    s_icell[i] = (uint8_t)(((1 + (int)i / iw) * c.width + 1 + (int)i % iw) & 255);
  }
  {
    const int BW = c.width + 4, BH = c.height + 4;
    for (int i = (int)threadIdx.x; i < IMA_BMAP; i += (int)blockDim.x) {
      const int r = i / BW - 2, cc = i % BW - 2;
      uint32_t e = (uint32_t)'W' | (((uint32_t)c.layer_w & 0xffu) << 8);                  /* what_lies_outside; a layer pads with (chr == 'W') */
      if (i < BW * BH && r >= 0 && r < c.height && cc >= 0 && cc < c.width)
        e = (uint32_t)c.base_board[r * c.width + cc] | (((uint32_t)c.base_layer[r * c.width + cc] & 0xffu) << 8);
      s_bmap[i] = (uint16_t)e;
    }
    for (int i = (int)threadIdx.x; i < 4 * IMA_VIEW; i += (int)blockDim.x) {
      const int dir = i / IMA_VIEW, vi = (i % IMA_VIEW) / 5, vj = i % 5;
      int si = vi, sj = vj;                                                                /* out[vi][vj] = in[si][sj] */
      if (dir == GW_DIR_DOWN) { si = 4 - vi; sj = 4 - vj; }
      else if (dir == GW_DIR_LEFT) { si = 4 - vj; sj = vi; }                               /* rot90 k=-1 (clockwise) */
      else if (dir == GW_DIR_RIGHT) { si = vj; sj = 4 - vi; }                              /* rot90 k=1 (counterclockwise) */
      s_voff[i] = (int16_t)((si - 2) * BW + (sj - 2));
      s_vdij[i] = (int16_t)(((si - 2) & 0xff) | ((sj - 2) << 8));
      s_vinv[dir * IMA_VIEW + si * 5 + sj] = (uint8_t)(vi * 5 + vj);
    }
    for (int i = (int)threadIdx.x; i < 128; i += (int)blockDim.x) {
      int8_t l = -1;
      for (int k = 0; k < c.n_layers; ++k) if (c.layer_chars[k] == (uint8_t)i) l = (int8_t)k;
      s_lchar[i] = l;
    }
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long* s_stats = a.stats ? a.stats + ((blockIdx.x * IMA_WARPS + warp) & (GW_STAT_REPLICAS - 1)) * GW_MA_STATS_LEN : nullptr;
  const uint32_t cells = (uint32_t)c.cells, L = (uint32_t)c.n_layers, R2 = 2u * (uint32_t)c.n_rewards;
  const uint32_t Sc = L * cells, Sv = 2u * IMA_VIEW, Sl = 2u * L * IMA_VIEW;
  uint8_t* wbuf = ima_stage + warp * a.warp_bytes;
  uint8_t* s_cube = wbuf + a.cube_off;
  uint8_t* s_board = wbuf + a.board_off;
  uint8_t* s_crop = wbuf + a.crop_off;
  uint8_t* s_lcrop = wbuf + a.lcrop_off;
  uint32_t* s_cbits = reinterpret_cast<uint32_t*>(s_cube);      /* PM: the chunk's cube as 32 * Sc bits (+ one word of read-ahead) */
  uint32_t* s_lbits = reinterpret_cast<uint32_t*>(s_lcrop);     /* PM: the chunk's layer views as 32 * Sl bits (+ one word) */
  uint8_t* s_map = wbuf + a.map_off;             /* PM: the chunk's 32 maps, [32][cells] as in a.maps */
  uint32_t parity = 0;                           /* reward rows are double-buffered: the previous chunk's may still be in flight */
  if constexpr (PM) {
    if (lane == 0) { s_cbits[Sc] = 0u; s_lbits[Sl] = 0u; }      /* the read-ahead words of the expansion */
  } else {
    /* agent-free templates of this lane's staged environment */
    for (uint32_t p = 0; p < cells; ++p) {
      s_board[lane * cells + p] = s_base[p];
      for (uint32_t l = 0; l < L; ++l) s_cube[lane * Sc + l * cells + p] = (uint8_t)(s_blayer[p] == (int8_t)l);
    }
  }
  __syncwarp();
  int32_t staged0 = -1, staged1 = -1;            /* the cells this lane's staged environment shows the agents on */

  const int64_t nchunks = (a.n + 31) >> 5;
  auto claim = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = queue_claim(a.claim_counter, (unsigned long long)nchunks);
    const int64_t got = (int64_t)__shfl_sync(FULL, v, 0);
    return got < 0 ? ((int64_t)1 << 60) : got;           /* a corrupted counter counts as "queue exhausted", never as work */
  };
  int64_t chunk = claim();
  ImaRaw next;
  if (chunk < nchunks && (chunk << 5) + lane < a.n) ima_raw_load(next, a, (chunk << 5) + lane);

#pragma unroll 1
  while (chunk < nchunks) {
    const int64_t chunk_next = claim();
    const int64_t env0 = chunk << 5, env = env0 + lane;
    const uint32_t nvalid = (uint32_t)min((int64_t)32, a.n - env0);
    const ImaRaw raw = next;
    if (chunk_next < nchunks && (chunk_next << 5) + lane < a.n) ima_raw_load(next, a, (chunk_next << 5) + lane);

    float* s_rw = reinterpret_cast<float*>(wbuf + a.reward_off + parity * (128u * R2));
    parity ^= 1u;

    ImaMap M;
    M.art = s_art; M.own = nullptr; M.icell = s_icell; M.shuffle = false; M.changed = false;
    if constexpr (PM) {
      /* this chunk's maps: one coalesced copy (the TMA engine has finished reading the previous chunk's staging before the
       * observation phase; the map staging is only touched by generic loads and stores) */
      const uint4* src = reinterpret_cast<const uint4*>(a.maps + env0 * (int64_t)cells);
      const uint32_t nb = nvalid * cells;
      for (uint32_t q = lane; q < (nb >> 4); q += 32) reinterpret_cast<uint4*>(s_map)[q] = src[q];
      for (uint32_t i = (nb & ~15u) + lane; i < nb; i += 32) s_map[i] = a.maps[env0 * (int64_t)cells + i];
      __syncwarp();
      M.art = s_map + lane * cells; M.own = s_map + lane * cells; M.shuffle = a.map_shuffle != 0;
    }
    ImaState s;
    int32_t out_st[2] = {0, 0};
    bool wrote = true, played = false;
    if (lane < nvalid) {
      if (a.is_reset) {
        ima_unpack(s, raw, c);
        wrote = !a.reset_mask || a.reset_mask[env] != 0;
        if (wrote) ima_new_game(c, a, s_art, env, M, s);
        for (uint32_t d = 0; d < R2; ++d) s_rw[lane * R2 + d] = 0.0f;
      } else {
        played = ima_step_lane<RM>(c, a, s_art, M, env, raw, s, s_rw + lane * R2, out_st, s_stats);
      }
      if (wrote) {
        ima_store(s, a.state, a.n, env);
        if (a.terminated) *reinterpret_cast<uchar2*>(a.terminated + env * 2) = make_uchar2((uint8_t)(out_st[0] >= 2), (uint8_t)(out_st[1] >= 2));
        if (a.step_type) *reinterpret_cast<uchar2*>(a.step_type + env * 2) = make_uchar2((uint8_t)out_st[0], (uint8_t)out_st[1]);
      }
    } else {
      s.ag[0].pos = c.start[0]; s.ag[1].pos = c.start[1]; s.ag[0].odir = s.ag[1].odir = GW_DIR_UP;
    }

    {
      const uint32_t m = __ballot_sync(FULL, played);          /* parallel steps played in this chunk */
      if (lane == 0 && m && s_stats) atomicAdd(&s_stats[0], (unsigned long long)__popc(m));
    }
    /* ---- observations: patch the agents into the staged templates, rebuild the views ---- */
    /* the TMA engine may still be reading the observation staging of the previous chunk */
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
    const int32_t p0 = s.ag[0].pos, p1 = s.ag[1].pos;
    if constexpr (PM) {
      if (__any_sync(FULL, M.changed)) {                       /* fresh layouts go back to the caller's tensor, whole chunk coalesced */
        uint8_t* dst = a.maps + env0 * (int64_t)cells;
        const uint32_t nb = nvalid * cells;
        for (uint32_t q = lane; q < (nb >> 4); q += 32) reinterpret_cast<uint4*>(dst)[q] = reinterpret_cast<const uint4*>(s_map)[q];
        for (uint32_t i = (nb & ~15u) + lane; i < nb; i += 32) dst[i] = s_map[i];
      }
      /* render board and cube of this lane's environment from its own map */
      if (a.cube) {
        for (uint32_t w = lane; w < Sc; w += 32) s_cbits[w] = 0u;
        __syncwarp();
      }
      if (lane < nvalid) {
        const uint8_t* art = M.art;
        const uint32_t cb = lane * Sc;
        const bool wc = a.cube != nullptr;
        for (uint32_t p = 0; p < cells; ++p) {
          uint32_t ch = art[p];
          if (ch == '1' || ch == '2') ch = ' ';
          int l = s_lchar[ch & 127u];
          if ((int32_t)p == p0) { ch = '1'; if (l == c.layer_gap) l = -1; if (c.layer_a0 >= 0 && wc) ima_bit_set(s_cbits, cb + (uint32_t)c.layer_a0 * cells + p); }
          if ((int32_t)p == p1) { ch = '2'; if (l == c.layer_gap) l = -1; if (c.layer_a1 >= 0 && wc) ima_bit_set(s_cbits, cb + (uint32_t)c.layer_a1 * cells + p); }
          s_board[lane * cells + p] = (uint8_t)ch;
          if (l >= 0 && wc) ima_bit_set(s_cbits, cb + (uint32_t)l * cells + p);
        }
      }
    } else
    if (p0 != staged0 || p1 != staged1) {
      if (staged0 >= 0) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {              /* take the agents off the cells staged last time */
          const int32_t q = k ? staged1 : staged0;
          s_board[lane * cells + q] = s_base[q];
          const int32_t la = k ? c.layer_a1 : c.layer_a0;
          if (la >= 0) s_cube[lane * Sc + la * cells + q] = 0;
          if (c.layer_gap >= 0) s_cube[lane * Sc + c.layer_gap * cells + q] = (uint8_t)(s_blayer[q] == c.layer_gap);
        }
      }
      s_board[lane * cells + p0] = '1';
      s_board[lane * cells + p1] = '2';
      if (c.layer_a0 >= 0) s_cube[lane * Sc + c.layer_a0 * cells + p0] = 1;
      if (c.layer_a1 >= 0) s_cube[lane * Sc + c.layer_a1 * cells + p1] = 1;
      if (c.layer_gap >= 0) { s_cube[lane * Sc + c.layer_gap * cells + p0] = 0; s_cube[lane * Sc + c.layer_gap * cells + p1] = 0; }
      staged0 = p0; staged1 = p1;
    }
    if (a.lcrop) {
      if constexpr (PM) {
        for (uint32_t w = lane; w < Sl; w += 32) s_lbits[w] = 0u;
      } else {
        uint4* z = reinterpret_cast<uint4*>(s_lcrop);
        for (uint32_t q = lane; q < (32u * Sl) >> 4; q += 32) z[q] = make_uint4(0u, 0u, 0u, 0u);
      }
      __syncwarp();
    }
    if (a.crop || a.lcrop) {
      if constexpr (PM) {
        if (lane < nvalid)
          ima_views_own(c, M.art, s_vdij, s_lchar, p0, p1, s.ag[0].odir, s.ag[1].odir, a.crop ? s_crop + lane * Sv : nullptr,
                        a.lcrop ? s_lbits : nullptr, lane * Sl);
      } else {
        ima_views(c, s_bmap, s_voff, s_vinv, p0, p1, s.ag[0].odir, s.ag[1].odir, a.crop ? s_crop + lane * Sv : nullptr,
                  a.lcrop ? s_lcrop + lane * Sl : nullptr);
      }
    }

    if (nvalid == 32 && !(a.is_reset && a.reset_mask)) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* generic-proxy writes -> visible to the TMA */
      __syncwarp();
      if (lane == 0) {
        if constexpr (!PM) {
          if (a.cube) bulk_store(a.cube + env0 * (int64_t)Sc, s_cube, 32u * Sc);
          if (a.lcrop) bulk_store(a.lcrop + env0 * (int64_t)Sl, s_lcrop, 32u * Sl);
        }
        if (a.board) bulk_store(a.board + env0 * (int64_t)cells, s_board, 32u * cells);
        if (a.crop) bulk_store(a.crop + env0 * (int64_t)Sv, s_crop, 32u * Sv);
        if (a.reward) bulk_store(a.reward + env0 * (int64_t)R2, s_rw, 128u * R2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if constexpr (PM) {
        /* the bit strings leave through the table: 32 bits -> 32 bytes per lane and iteration, coalesced 16-byte stores */
        if (a.cube) fm_expand_bits(a.cube + env0 * (int64_t)Sc, s_cbits, (int)(32u * Sc), s_lut, lane);
        if (a.lcrop) fm_expand_bits(a.lcrop + env0 * (int64_t)Sl, s_lbits, (int)(32u * Sl), s_lut, lane);
        __syncwarp();                                /* the next chunk clears the strings */
      }
    } else {
      /* ragged last chunk / masked reset: plain stores, each lane its own environment */
      __syncwarp();
      if (lane < nvalid) {
        if (a.board) for (uint32_t i = 0; i < cells; ++i) a.board[env * cells + i] = s_board[lane * cells + i];
        if constexpr (PM) {
          if (a.cube) for (uint32_t i = 0; i < Sc; ++i) { const uint32_t b = lane * Sc + i; a.cube[env * (int64_t)Sc + i] = (uint8_t)((s_cbits[b >> 5] >> (b & 31u)) & 1u); }
          if (a.lcrop) for (uint32_t i = 0; i < Sl; ++i) { const uint32_t b = lane * Sl + i; a.lcrop[env * (int64_t)Sl + i] = (uint8_t)((s_lbits[b >> 5] >> (b & 31u)) & 1u); }
        } else {
          if (a.cube) for (uint32_t i = 0; i < Sc; ++i) a.cube[env * (int64_t)Sc + i] = s_cube[lane * Sc + i];
          if (a.lcrop) for (uint32_t i = 0; i < Sl; ++i) a.lcrop[env * (int64_t)Sl + i] = s_lcrop[lane * Sl + i];
        }
        if (a.crop) for (uint32_t i = 0; i < Sv; ++i) a.crop[env * (int64_t)Sv + i] = s_crop[lane * Sv + i];
        if (a.reward && wrote) for (uint32_t i = 0; i < R2; ++i) a.reward[env * (int64_t)R2 + i] = s_rw[lane * R2 + i];
      }
      __syncwarp();
    }
    chunk = chunk_next;
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}

struct ImaObserveArgs {
  const uint4* state;
  double* metrics;
  float* cumulative;
  int32_t* frame;
  int16_t* pos;
  int8_t* directions;
  int64_t n;
};

__global__ void __launch_bounds__(GW_BLOCK) gw_ima_observe_kernel(const __grid_constant__ ImaCfg c, const ImaObserveArgs a) {
  const int64_t env = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (env >= a.n) return;
  ImaRaw raw;
  ImaArgs la;
  la.state = const_cast<uint4*>(a.state); la.n = a.n; la.actions = nullptr; la.order = nullptr;
  ima_raw_load(raw, la, env);
  ImaState s;
  ima_unpack(s, raw, c);
  if (a.frame) a.frame[env] = s.frame;
#pragma unroll
  for (int p = 0; p < 2; ++p) {
    if (a.pos) { a.pos[(env * 2 + p) * 2] = (int16_t)(s.ag[p].pos / c.width); a.pos[(env * 2 + p) * 2 + 1] = (int16_t)(s.ag[p].pos % c.width); }
    if (a.directions) { a.directions[(env * 2 + p) * 2] = (int8_t)s.ag[p].adir; a.directions[(env * 2 + p) * 2 + 1] = (int8_t)s.ag[p].odir; }
    if (a.cumulative) for (int d = 0; d < c.n_rewards; ++d) a.cumulative[(env * 2 + p) * c.n_rewards + d] = s.cum[p][d];
    if (a.metrics) {
      double* m = a.metrics + env * GW_IMA_METRICS;
#pragma unroll
      for (int k = 0; k < 5; ++k) m[p * 5 + k] = (double)s.ag[p].visits[k];
      m[GW_IMA_M_DRINK_SATIATION_1 + 2 * p] = s.ag[p].dsat;
      m[GW_IMA_M_FOOD_SATIATION_1 + 2 * p] = s.ag[p].fsat;
    }
  }
  if (a.metrics) {
    a.metrics[env * GW_IMA_METRICS + GW_IMA_M_DRINK_AVAILABILITY] = s.dav;
    a.metrics[env * GW_IMA_METRICS + GW_IMA_M_FOOD_AVAILABILITY] = s.fav;
  }
}
