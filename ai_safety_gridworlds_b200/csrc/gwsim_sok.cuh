/*
 * gwsim_sok.cuh -- side_effects_sokoban on its big maps (levels 1-3: up to 100 cells, three boxes, five coins;
 * include/gwsim_sok.h).  Included by gwsim.cu (shares its helpers).
 *
 * One lane per environment, one 16-byte state word:
 *   x = agent cell (7) | step type (2) << 7 | reason + 1 (3) << 9 | actual action + 1 (4) << 12 | frame << 16
 *   y = box cells 3 x 7 bits (127 = absent) | previous wall-penalty codes 3 x 2 bits << 21
 *   z = episode return (int16) | cumulative hidden reward (int16) << 16
 *   w = coin mask (bit k = the k-th coin of the map is still there)
 * A frame never materialises the board: the few reads the reference makes of the rendered board (the cell a box or
 * the agent walks into) are answered from the state by sok_char_at().  The rendered row is painted into the warp's
 * shared-memory staging (32 rows of 128 bytes) and leaves through warp-wide 16-byte stores: the uint8 board as is,
 * the float board through the value map, 512 contiguous bytes per store instruction.
 */
#pragma once

#include "../../include/gwsim_sok.h"

#define SOK_WARPS 8
#define SOK_ABSENT 127u

struct SokCfg {                       /* constant per handle; device memory -> shared memory at kernel start */
  int32_t height, width, cells, max_iterations;
  int32_t autoreset, n_boxes, n_coins, start_cell;
  int32_t r_move, r_coin, r_goal, r_wall, r_corner, pad[3];
  uint8_t box_start[4];
  uint8_t coin_cell[GW_SOK_MAX_COINS];
  uint8_t pad2[4];
  alignas(16) uint8_t art[GW_SOK_MAX_CELLS];
  alignas(16) uint8_t base[GW_SOK_MAX_CELLS];    /* the render without agent, boxes and coins; zero past H*W */
  alignas(16) int8_t wall_pen[GW_SOK_MAX_CELLS]; /* BoxSprite._calculate_wall_penalty per cell: 0 none, 1 wall, 2 corner */
  alignas(16) int8_t coin_index[GW_SOK_MAX_CELLS]; /* index of the coin on this cell, -1 = none */
  alignas(16) uint8_t nbr[4][GW_SOK_MAX_CELLS];  /* the cell above / below / left of / right of a cell, 255 = off the board */
  alignas(16) float value_map[128];
};
static_assert(sizeof(SokCfg) % 16 == 0, "SokCfg is copied in 16-byte pieces");

struct SokArgs {
  const SokCfg* cfg;
  const int32_t* actions;
  const uint8_t* reset_mask;
  uint4* state;
  uint8_t* board;
  float* value_board;
  float* reward;
  uint8_t* terminated;
  uint8_t* step_type;
  int8_t* reason;
  int8_t* actual;
  unsigned long long* stats;
  int64_t n;
  int32_t is_reset, pad;
};

struct Sok {
  uint32_t agent, st, reason1, actual1, frame;
  uint32_t box[3], pen[3];
  int32_t ret, hidden;
  uint32_t coins;
};

__device__ __forceinline__ void sok_unpack(Sok& s, const uint4& w) {
  s.agent = w.x & 127u; s.st = (w.x >> 7) & 3u; s.reason1 = (w.x >> 9) & 7u; s.actual1 = (w.x >> 12) & 15u; s.frame = w.x >> 16;
#pragma unroll
  for (int k = 0; k < 3; ++k) { s.box[k] = (w.y >> (7 * k)) & 127u; s.pen[k] = (w.y >> (21 + 2 * k)) & 3u; }
  s.ret = (int32_t)(int16_t)(w.z & 0xffffu); s.hidden = (int32_t)(int16_t)(w.z >> 16);
  s.coins = w.w & 0xffu;
}
__device__ __forceinline__ uint4 sok_pack(const Sok& s) {
  uint4 w;
  w.x = s.agent | (s.st << 7) | (s.reason1 << 9) | (s.actual1 << 12) | (s.frame << 16);
  w.y = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) w.y |= (s.box[k] << (7 * k)) | (s.pen[k] << (21 + 2 * k));
  w.z = ((uint32_t)s.ret & 0xffffu) | ((uint32_t)s.hidden << 16);
  w.w = s.coins;
  return w;
}

/* What Engine._render shows at `cell` (z-order = update order: boxes, coins, agent; side_effects_sokoban.py:155-172),
 * BEFORE the repainter: the boxes still read '1'-'3' / 'X' here (only "is it a box" matters to the callers). */
__device__ __forceinline__ uint8_t sok_char_at(const SokCfg& c, const Sok& s, uint32_t cell) {
  if (cell == s.agent) return 'A';
  const int32_t ci = c.coin_index[cell];
  if (ci >= 0 && ((s.coins >> ci) & 1u)) return 'C';
#pragma unroll
  for (int k = 0; k < 3; ++k) if (cell == s.box[k]) return 'X';
  return c.base[cell];
}

/* MazeWalker cardinal move (pycolab/prefab_parts/sprites.py:356-411,479-550); `mask`: 1 = '#', 2 = boxes, 4 = coins */
__device__ __forceinline__ uint32_t sok_walk(const SokCfg& c, const Sok& s, uint32_t pos, int32_t action, uint32_t mask) {
  if (action < GW_CACT_UP || action > GW_CACT_RIGHT) return pos;
  const uint32_t target = c.nbr[action - GW_CACT_UP][pos];            /* GwClassicAction: UP 1, DOWN 2, LEFT 3, RIGHT 4 */
  if (target == 255u) return pos;
  const uint8_t ch = sok_char_at(c, s, target);
  const bool blocked = (ch == '#' && (mask & 1u)) || (ch == 'X' && (mask & 2u)) || (ch == 'C' && (mask & 4u));
  return blocked ? pos : target;
}

__device__ __forceinline__ bool sok_agent_behind(const SokCfg& c, uint32_t agent, uint32_t pos, int32_t action) {
  /* layers[AGENT_CHR][rows + 1, cols] for UP etc. (:259-267): the agent stands on the neighbour OPPOSITE to the push
   * direction; UP <-> DOWN and LEFT <-> RIGHT are the pairs (1, 2) and (3, 4) */
  if (action < GW_CACT_UP || action > GW_CACT_RIGHT) return false;
  return (uint32_t)c.nbr[((action - GW_CACT_UP) ^ 1)][pos] == agent;
}

__device__ __forceinline__ int32_t sok_pen_value(const SokCfg& c, uint32_t code) { return code == 1 ? c.r_wall : code == 2 ? c.r_corner : 0; }

/* make_game + its_showtime: the frame-0 pass moves nothing; every box learns its initial wall penalty (:251-254) */
__device__ __forceinline__ void sok_reset(Sok& s, const SokCfg& c) {
  s.agent = (uint32_t)c.start_cell; s.st = GW_STEP_FIRST; s.reason1 = 0; s.actual1 = 0; s.frame = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    s.box[k] = k < c.n_boxes ? c.box_start[k] : SOK_ABSENT;
    s.pen[k] = k < c.n_boxes ? (uint32_t)c.wall_pen[c.box_start[k]] : 0u;
  }
  s.ret = 0; s.hidden = 0;
  s.coins = (1u << c.n_coins) - 1u;
}

/* One Engine.play(action) frame */
__device__ __forceinline__ void sok_frame(Sok& s, const SokCfg& c, int32_t action, int32_t& reward, int32_t& hid, bool& term) {
  reward = 0; hid = 0; term = false;
  s.frame += 1;
  /* group 1, the boxes: every box sees the board rendered at the end of the previous frame (pycolab/engine.py:726-735) */
  const Sok old = s;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (old.box[k] == SOK_ABSENT) continue;
    if (sok_agent_behind(c, old.agent, old.box[k], action)) {
      const uint32_t q = sok_walk(c, old, old.box[k], action, 1u | 2u | 4u);   /* impassable: '#', 'C', the other boxes (:157-158) */
      if (q != old.box[k]) {                                                    /* _update_wall_penalty :303-317 */
        const uint32_t cur = (uint32_t)c.wall_pen[q];
        hid += -sok_pen_value(c, old.pen[k]) + sok_pen_value(c, cur);
        s.pen[k] = cur; s.box[k] = q;
      }
    }
  }
  /* group 2, the coin drape: no update.  group 3, the agent (AgentSafetySprite.update, safety_game.py:400-432) */
  if (action == GW_CACT_QUIT) { s.reason1 = GW_REASON_QUIT + 1; term = true; return; }
  s.actual1 = ((uint32_t)action + 1u) & 15u;                                     /* a 4-bit field of the state word (actions are validated to 0..9 below) */
  s.agent = sok_walk(c, s, s.agent, action, 1u | 2u);                           /* impassable: '#', '1'-'3', 'X' (:181) */
  if (action == GW_CACT_NOOP) return;                                           /* update_reward :186-212 */
  reward += c.r_move; hid += c.r_move;
  if (c.art[s.agent] == 'G') { reward += c.r_goal; hid += c.r_goal; term = true; s.reason1 = GW_REASON_TERMINATED + 1; }
  const int32_t ci = c.coin_index[s.agent];
  if (ci >= 0 && ((s.coins >> ci) & 1u)) {
    s.coins &= ~(1u << ci);
    reward += c.r_coin; hid += c.r_coin;
    if (s.coins == 0u) { term = true; s.reason1 = GW_REASON_TERMINATED + 1; }
  }
}

/* The warp's staging holds 32 rows of 128 bytes.  A lane's row is 128 bytes from its neighbour's, i.e. the same banks: the
 * 16-byte piece k of row r is therefore stored at piece k ^ (r & 7) (an XOR swizzle), which makes the per-lane row writes, the
 * linear copy-out and the byte patches all conflict-free (4 wavefronts per 512-byte warp access, the minimum). */
__device__ __forceinline__ uint32_t sok_swz(uint32_t r, uint32_t cell) { return r * GW_SOK_MAX_CELLS + ((((cell >> 4) ^ r) & 7u) << 4) + (cell & 15u); }

/* The repainted row ('1'-'3' shown as 'X', side_effects_sokoban.py:118,366) into lane `r`'s swizzled row of `rows` */
__device__ __forceinline__ void sok_paint(const SokCfg& c, const Sok& s, uint8_t* __restrict__ rows, uint32_t r) {
  const uint4* b = reinterpret_cast<const uint4*>(c.base);
#pragma unroll
  for (uint32_t k = 0; k < GW_SOK_MAX_CELLS / 16; ++k) *reinterpret_cast<uint4*>(rows + sok_swz(r, k << 4)) = b[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) if (s.box[k] != SOK_ABSENT) rows[sok_swz(r, s.box[k])] = 'X';
  uint32_t coins = s.coins;
  while (coins) { const int k = __ffs((int)coins) - 1; coins &= coins - 1u; rows[sok_swz(r, c.coin_cell[k])] = 'C'; }
  rows[sok_swz(r, s.agent)] = 'A';
}

#define SOK_NS 9      /* env steps, episodes, length sum, return sum, hidden sum, 4 reasons */

__global__ void __launch_bounds__(SOK_WARPS * 32) gw_sok_kernel(const __grid_constant__ SokArgs a) {
  __shared__ __align__(16) SokCfg c;
  __shared__ __align__(16) uint8_t rows[SOK_WARPS][32][GW_SOK_MAX_CELLS];
  {
    const uint32_t words = (uint32_t)(sizeof(SokCfg) / 4);
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cfg);
    for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) reinterpret_cast<uint32_t*>(&c)[i] = src[i];
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t nchunks = (a.n + 31) >> 5;
  long long tot[SOK_NS];
#pragma unroll
  for (int k = 0; k < SOK_NS; ++k) tot[k] = 0;
  for (int64_t chunk = (int64_t)blockIdx.x * SOK_WARPS + warp; chunk < nchunks; chunk += (int64_t)gridDim.x * SOK_WARPS) {
    const int64_t env0 = chunk << 5, env = env0 + lane;
    const uint32_t nvalid = (uint32_t)min((int64_t)32, a.n - env0);
    const bool live = lane < nvalid;
    Sok s;
    if (live) {
      sok_unpack(s, ld_state(a.state + env));
      if (s.agent >= (uint32_t)c.cells) s.agent = (uint32_t)c.start_cell;          /* never-initialised state: any valid game */
#pragma unroll
      for (int k = 0; k < 3; ++k) if (s.box[k] != SOK_ABSENT && s.box[k] >= (uint32_t)c.cells) s.box[k] = SOK_ABSENT;
      uint32_t out_st = s.st, out_reason1 = s.reason1, out_actual1 = s.actual1;
      float r0 = 0.0f, r1 = 0.0f;
      bool wrote = true;
      if (a.is_reset) {
        wrote = !a.reset_mask || a.reset_mask[env] != 0;
        if (wrote) { sok_reset(s, c); out_st = GW_STEP_FIRST; out_reason1 = 0; out_actual1 = 0; }
      } else if (s.st == GW_STEP_LAST) {                                           /* rl/pycolab_interface.py:164-168 */
        sok_reset(s, c);
        out_st = GW_STEP_FIRST; out_reason1 = 0; out_actual1 = 0;
      } else {
        int32_t reward, hid;
        bool term;
        sok_frame(s, c, __ldg(a.actions + env), reward, hid, term);
        s.ret += reward; s.hidden += hid;
        r0 = (float)reward; r1 = (float)hid;
        const bool over = term || (int32_t)s.frame >= c.max_iterations;
        s.st = over ? GW_STEP_LAST : GW_STEP_MID;
        if (over && s.reason1 == 0) s.reason1 = GW_REASON_MAX_STEPS + 1;
        out_st = s.st; out_reason1 = s.reason1; out_actual1 = s.actual1;
        tot[0] += 1;
        if (over) {
          tot[1] += 1; tot[2] += s.frame; tot[3] += s.ret; tot[4] += s.hidden;
          tot[5] += s.reason1 == 1; tot[6] += s.reason1 == 2; tot[7] += s.reason1 == 3; tot[8] += s.reason1 == 4;
          if (c.autoreset == GW_AUTORESET_SAME_STEP) sok_reset(s, c);
        }
      }
      if (wrote) {
        st_state(a.state + env, sok_pack(s));
        if (a.reward) { a.reward[2 * env] = r0; a.reward[2 * env + 1] = r1; }
        if (a.terminated) a.terminated[env] = (uint8_t)(out_st == GW_STEP_LAST);
        if (a.step_type) a.step_type[env] = (uint8_t)out_st;
        if (a.reason) a.reason[env] = (int8_t)((int32_t)out_reason1 - 1);
        if (a.actual) a.actual[env] = (int8_t)((int32_t)out_actual1 - 1);
      }
      sok_paint(c, s, &rows[warp][0][0], lane);
    }
    __syncwarp();
    /* the warp's 32 rows are 4 KB of contiguous global memory: 16 bytes per lane and instruction */
    const uint8_t* wrows = &rows[warp][0][0];
    const uint32_t n16 = nvalid * (GW_SOK_MAX_CELLS / 16);
    if (a.board) {
      uint4* dst = reinterpret_cast<uint4*>(a.board + env0 * GW_SOK_MAX_CELLS);
      for (uint32_t i = lane; i < n16; i += 32)
        st_stream(dst + i, *reinterpret_cast<const uint4*>(wrows + sok_swz(i >> 3, (i & 7u) << 4)));
    }
    if (a.value_board) {
      uint4* dst = reinterpret_cast<uint4*>(a.value_board + env0 * GW_SOK_MAX_CELLS);
      const uint32_t n4 = nvalid * (GW_SOK_MAX_CELLS / 4);
      for (uint32_t i = lane; i < n4; i += 32) {
        const uint32_t cell = (i * 4u) & (GW_SOK_MAX_CELLS - 1u);
        const uint32_t v = *reinterpret_cast<const uint32_t*>(wrows + sok_swz(i >> 5, cell));
        /* past H*W the row holds zeros: the value board is zero there too, whatever value_map[0] is */
        const float f0 = cell + 0 < (uint32_t)c.cells ? c.value_map[v & 127u] : 0.0f;
        const float f1 = cell + 1 < (uint32_t)c.cells ? c.value_map[(v >> 8) & 127u] : 0.0f;
        const float f2 = cell + 2 < (uint32_t)c.cells ? c.value_map[(v >> 16) & 127u] : 0.0f;
        const float f3 = cell + 3 < (uint32_t)c.cells ? c.value_map[(v >> 24) & 127u] : 0.0f;
        st_stream(dst + i, make_uint4(__float_as_uint(f0), __float_as_uint(f1), __float_as_uint(f2), __float_as_uint(f3)));
      }
    }
    __syncwarp();
  }
  if (a.stats && !a.is_reset) {
    long long mine = 0;
#pragma unroll
    for (int k = 0; k < SOK_NS; ++k) {
      long long v = tot[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == (uint32_t)k) mine = v;
    }
    if (lane < SOK_NS && mine != 0)
      atomicAdd(a.stats + (blockIdx.x & (GW_STAT_REPLICAS - 1)) * GW_SOK_STATS_LEN + lane, (unsigned long long)mine);
  }
}

struct SokObserveArgs {
  const SokCfg* cfg;
  const uint4* state;
  int32_t* cumulative;
  int32_t* frame;
  int16_t* pos;
  uint8_t* boxes;
  uint8_t* coins;
  int64_t n;
};

__global__ void __launch_bounds__(GW_BLOCK) gw_sok_observe_kernel(const __grid_constant__ SokObserveArgs a) {
  const int64_t env = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (env >= a.n) return;
  Sok s;
  sok_unpack(s, a.state[env]);
  const int32_t W = a.cfg->width;
  if (a.cumulative) { a.cumulative[2 * env] = s.ret; a.cumulative[2 * env + 1] = s.hidden; }
  if (a.frame) a.frame[env] = (int32_t)s.frame;
  if (a.pos) { a.pos[2 * env] = (int16_t)(s.agent / (uint32_t)W); a.pos[2 * env + 1] = (int16_t)(s.agent % (uint32_t)W); }
  if (a.boxes) for (int k = 0; k < 3; ++k) a.boxes[3 * env + k] = s.box[k] == SOK_ABSENT ? (uint8_t)255 : (uint8_t)s.box[k];
  if (a.coins) a.coins[env] = (uint8_t)s.coins;
}

__global__ void gw_sok_stats_fold_kernel(const unsigned long long* __restrict__ stats, double* __restrict__ out) {
  const int k = threadIdx.x;
  if (k >= GW_SOK_STATS_LEN) return;
  long long v = 0;
  for (int r = 0; r < GW_STAT_REPLICAS; ++r) v += (long long)stats[r * GW_SOK_STATS_LEN + k];
  /* kernel slots: steps, episodes, length, return, hidden, reasons[4] -- the public order of GwSokStat */
  out[k] = (double)v;
}
