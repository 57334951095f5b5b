/*
 * gwsim_fm.cuh -- firemaker_ex_ma (BASELINE config 4) on the GPU: one WARP per environment.
 * Included by gwsim.cu.  ABI: include/gwsim_fm.h.
 *
 * The 17x17 board does not fit a lane, and the fire update is a stencil, so the 32 lanes of a warp
 * cooperate on one environment: lane j owns cells j, j+32, ... (10 slots).  Scalar game state
 * (positions, countdown, counters, reward sums) is kept redundantly in every lane (uniform
 * registers); the fire curtain lives in a bordered 21x21 byte map in shared memory so that the 5x5
 * spread stencil needs no bounds checks.  A parallel step runs the reference's three sequential
 * per-agent Engine.play frames (rl/pycolab_interface_ma.py:183-230) back to back:
 *
 *   acting agent    MazeWalker move vs walls and the other agents, visit counters  (firemaker_ex_ma.py:430-476)
 *   StopButton      countdown                                                       (:656-673)
 *   Workshop        work / energy rewards                                           (:496-517)
 *   Fire            per target cell: P = 1-(1-P)(1-p) over burning cells within the 5x5 stencil in
 *                   row-major source order (bit-exact with the reference's accumulation order), then
 *                   one uniform draw per cell with P > 0 (row-major rank via ballots) and one per
 *                   previously burning cell                                         (:539-629)
 *   Territory       supervisor trespassing                                          (:699-704)
 *
 * Draws come from a caller-supplied trace (replay of a recorded reference run) or from Philox keyed
 * by (seed, global env, call, draw index).  Then the warp renders the global board, the layers cube
 * and the three agent-centred crops with their layers (safety_game_moma.py:1996-2101).
 */
#pragma once

#include "../../include/gwsim_fm.h"
#include "../../include/gwsim_ima.h"     /* GwDirection */

#define FM_S GW_FM_SIDE
#define FM_CELLS GW_FM_CELLS
#define FM_B 21                         /* bordered side: 17 + 2 + 2 */
#define FM_SLOTS 10                     /* ceil(289 / 32) */
#ifndef FM_WARPS
#define FM_WARPS 7                      /* 4 CTAs of 7 warps per SM: 72 registers per thread, no spills in the hot loops.  Measured per 262,144-game
                                           step (B200): 8 warps x 4 CTAs at 64 registers (168 B of spills) 1.26 ms, 6 x 5 and 5 x 6 (64 registers)
                                           1.27 ms, 6 x 4 (80 registers) 1.16 ms, 7 x 4 and 4 x 7 (72 registers) 1.15 ms */
#endif
#define FM_BATCH 16                     /* games per warp batch: the scalar game logic runs lane per game */
#define FM_BST 44                       /* words between the states of a batch in shared memory (40 + padding against bank conflicts) */

enum { FM_F_WALL = 1, FM_F_WORKSHOP = 2, FM_F_BUTTON = 4, FM_F_TERRITORY = 8 };

struct alignas(16) FmStatic {           /* per-handle tables: device memory -> shared memory per CTA */
  double spread_p[25];                  /* by (dr + 2) * 5 + (dc + 2), firemaker_ex_ma.py:595-598 */
  double spread_c[25];                  /* 1 - spread_p, rounded once as the reference's (1 - p) is */
  double cont_p;
  double rewards[8];
  int32_t start[4];
  int32_t max_iterations, autoreset, randomize, button_duration;
  int32_t two_workers, static2;         /* amount_agents == 3; with 2 there is no worker '2': start[1] = 0xffff, never on the board,
                                           and static2 = the art's '2' tile, a backdrop character whose layer reads 1 */
  int32_t obs_mode, act_mode;           /* direction modes (gw_fm_kernel<true> when either is not 0) */
  uint8_t base_chr[FM_CELLS + 15];      /* render without fire and agents: '#', ' ', '-', 'W', 'B' */
  uint8_t flags[FM_CELLS + 15];
  uint32_t lay_static[5][12];           /* flat 289-bit maps (bit cell & 31 of word cell >> 5): cells with no flag at all (the gap layer before
                                           fire and agents), walls, territory, stop button, workshop */
  uint32_t lay_static_t[5][12];         /* the same maps transposed (bit col * 17 + row): the supervisor's view turned by a quarter reads the
                                           board column by column (direction modes 1-2) */
};
enum { FM_SL_GAP = 0, FM_SL_WALL = 1, FM_SL_TERRITORY = 2, FM_SL_BUTTON = 3, FM_SL_WORKSHOP = 4 };

struct FmArgs {
  const FmStatic* st;
  const int32_t* actions;
  const int32_t* order;
  const double* draws;
  int64_t draw_stride;
  const uint8_t* reset_mask;
  uint4* state;                         /* [N][10] words, AoS: 160 contiguous bytes per environment */
  uint8_t *board, *cube, *crop_w, *crop_s, *lcrop_w, *lcrop_s;
  float *reward_w, *reward_s;
  uint8_t *terminated, *step_type;
  uint64_t seed, call_no;
  int64_t env_index_base, n;
  int32_t is_reset, pad;
  unsigned long long* claim_counter;    /* dynamic environment queue of the persistent kernel */
  unsigned long long* stats;            /* [GW_STAT_REPLICAS][GW_MA_STATS_LEN] raw rollout statistics */
};


/* Philox4x32-10 keyed by the seed, counter (global environment index, call * 65536 + evaluation index).  The FireDrape draws take one
 * 32-bit word each (draw d = word d & 3 of evaluation d >> 2, u = word / 2^32); the shuffle of the agents' order takes the two 53-bit
 * uniforms of evaluation 32767 (words (x, y) and (z, w)). */
__device__ __forceinline__ uint4 fm_philox(const FmArgs& a, int64_t env, uint32_t call_index) {
  const uint64_t g = (uint64_t)(a.env_index_base + env);
  const uint64_t step = a.call_no * 65536ull + call_index;
  return philox4x32_10(make_uint4((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                       (uint32_t)a.seed, (uint32_t)(a.seed >> 32));
}
__device__ __forceinline__ double fm_u53(uint32_t hi, uint32_t lo) {
  return (double)((((unsigned long long)hi << 32) | lo) >> 11) * (1.0 / 9007199254740992.0);
}

/* get_absolute_action / get_new_action_or_observation_direction, mode 1 (safety_game_ma.py:505-587): UP = forwards, DOWN =
 * backwards, LEFT / RIGHT = a quarter turn from `dir` (LEFT 0, RIGHT 1, UP 2, DOWN 3): one byte per current direction */
__device__ __forceinline__ int fm_relative(int action, int dir) {
  const uint32_t tab = action == GW_ACT_UP ? 0x03020100u       /* keep */
                     : action == GW_ACT_DOWN ? 0x02030001u     /* opposite: L->R, R->L, U->D, D->U */
                     : action == GW_ACT_LEFT ? 0x01000203u     /* L->D, R->U, U->L, D->R */
                     : 0x00010302u;                            /* RIGHT: L->U, R->D, U->R, D->L */
  return (int)((tab >> (8 * dir)) & 3u);
}
/* direction mode 2 (safety_game_ma.py:607-640): only the TURN_* actions change a direction */
__device__ __forceinline__ int fm_turned(int action, int dir) {
  if (action == GW_ACT_TURN_LEFT_90) return fm_relative(GW_ACT_LEFT, dir);
  if (action == GW_ACT_TURN_RIGHT_90) return fm_relative(GW_ACT_RIGHT, dir);
  if (action == GW_ACT_TURN_LEFT_180 || action == GW_ACT_TURN_RIGHT_180) return fm_relative(GW_ACT_DOWN, dir);
  return dir;
}
#define FM_DIRS_UP 0xaaau                 /* packed directions of a fresh game: action / observation direction UP (2) for the three agents */

/* per-warp scratch words of the fire update */
enum { FM_X_ROWMASK = 0, FM_X_NEAR = FM_B + 3, FM_X_CAND = FM_X_NEAR + FM_S + 1, FM_X_PRE = FM_X_CAND + FM_SLOTS,
       FM_X_NEW = FM_X_PRE + FM_SLOTS + 1, FM_X_PREB = FM_X_NEW + FM_SLOTS + 1, FM_X_WORDS = FM_X_PREB + FM_SLOTS + 2 };
#ifndef FM_UBUF
#define FM_UBUF 232                     /* draws of one frame: one per candidate + one per burning cell, both subsets of the cells a fire may
                                           occupy (not wall / workshop / button): <= 225 inside the 17 x 17 border; gw_fm_create checks the map */
#endif

/* k-th (0-based) set bit of the FM_SLOTS-word bitmap `bits`, given the exclusive prefix popcounts `pre`: returns the cell */
__device__ __forceinline__ int fm_select(const uint32_t* __restrict__ bits, const uint32_t* __restrict__ pre, uint32_t k) {
  int w = 0;
#pragma unroll
  for (int q = 1; q < FM_SLOTS; ++q) w += (k >= pre[q]) ? 1 : 0;
  return w * 32 + (int)__fns(bits[w], 0u, (int)(k - pre[w]) + 1);
}

/* exclusive prefix popcounts of a FM_SLOTS-word bitmap into pre[0..FM_SLOTS]; returns the total */
__device__ __forceinline__ uint32_t fm_prefix(const uint32_t* __restrict__ bits, uint32_t* __restrict__ pre, uint32_t lane) {
  uint32_t c = lane < FM_SLOTS ? (uint32_t)__popc(bits[lane]) : 0u, incl = c;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) { const uint32_t t = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += t; }
  if (lane <= FM_SLOTS) pre[lane] = incl - c;
  const uint32_t total = __shfl_sync(FULL, incl, FM_SLOTS - 1);
  __syncwarp();
  return total;
}

/* FireDrape.update for the whole warp (firemaker_ex_ma.py:539-629); returns the number of external fires.  The fire curtain is
 * the flat 289-bit map `fw` (bit cell & 31 of word cell >> 5: the state's own format).  Instead of visiting all 289 cells, the
 * warp derives the CANDIDATE targets (not burning, not wall / workshop / button, a burning cell or a working worker within the
 * 5x5 stencil) with row-mask arithmetic, compacts them, and spends one lane per candidate: ceil(candidates / 32) rounds of the
 * spread recurrence and the Philox draw instead of ten 32-cell slots.  Candidates are taken in row-major order, so the k-th
 * candidate consumes the k-th draw exactly as the reference's double loop does; then the cells that were burning draw for
 * their continuation, compacted the same way.  `k` is the running draw index of this environment within the call. */
__device__ __forceinline__ int fm_fire_update(const FmStatic& S, const FmArgs& a, int64_t env, uint32_t* __restrict__ fw,
                                              uint32_t* __restrict__ x, const uint32_t* __restrict__ s_allowed,
                                              const uint32_t* __restrict__ s_extw, const int32_t* pos, const bool* at_w,
                                              int32_t countdown, uint32_t& k, uint32_t lane, double* __restrict__ ub) {
  uint32_t* __restrict__ rowmask = x + FM_X_ROWMASK;   /* bordered: row r + 2, bit c + 2 */
  uint32_t* __restrict__ near = x + FM_X_NEAR;
  uint32_t* __restrict__ cand = x + FM_X_CAND;
  uint32_t* __restrict__ pre = x + FM_X_PRE;
  uint32_t* __restrict__ newf = x + FM_X_NEW;
  __syncwarp();
  if (lane < 3) {                                        /* fires under agents are put out (:543-545) */
    const int p = lane == 0 ? pos[0] : lane == 1 ? pos[1] : pos[2];
    if (p < FM_CELLS) atomicAnd(&fw[p >> 5], ~(1u << (p & 31)));
  }
  __syncwarp();
  if (lane < FM_B) {                                     /* row r = bits [17 r, 17 r + 17) of the flat map */
    uint32_t m = 0;
    if (lane >= 2 && lane < 2 + FM_S) {
      const uint32_t b = FM_S * (lane - 2), w = b >> 5;
      m = (__funnelshift_r(fw[w], w + 1 < FM_SLOTS ? fw[w + 1] : 0u, b & 31u) & 0x1ffffu) << 2;
    }
    rowmask[lane] = m;
  }
  if (lane < FM_SLOTS) newf[lane] = 0u;
  __syncwarp();
  const int vs0 = (countdown == 0 && at_w[0]) ? pos[0] : -1;        /* working workers are virtual fire sources (:555-559) */
  const int vs1 = (countdown == 0 && at_w[1]) ? pos[1] : -1;
  /* candidate row masks: near[tr] = OR of the five source rows; bit tc of (U | U>>1 | ... | U>>4) <=> a source within columns tc-2..tc+2 */
  uint32_t rc = 0;
  if (lane < FM_S) {
    const uint32_t U = rowmask[lane] | rowmask[lane + 1] | rowmask[lane + 2] | rowmask[lane + 3] | rowmask[lane + 4];
    near[lane] = U;
    uint32_t D = U | (U >> 1) | (U >> 2) | (U >> 3) | (U >> 4);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int vs = q ? vs1 : vs0;
      if (vs >= 0) {
        const int vr = vs / FM_S, vc = vs % FM_S;
        if (abs(vr - (int)lane) <= 2) D |= (0x1fu << (vc + 2)) >> 4;   /* columns vc-2..vc+2 (bits below 0 fall off) */
      }
    }
    rc = D & 0x1ffffu & ~(rowmask[lane + 2] >> 2) & s_allowed[lane];
  }
  const uint32_t any = __ballot_sync(FULL, rc != 0u);
  const uint32_t burning_any = __ballot_sync(FULL, lane < FM_SLOTS && fw[lane < FM_SLOTS ? lane : 0] != 0u);
  if (!any && !burning_any) return 0;                    /* nothing can ignite and nothing burns: no draws, no fires */
  /* row masks -> flat candidate words: lane s < FM_SLOTS assembles word s from rows r0 .. r0 + 2, fetched by shuffle from the
   * lanes that own them */
  {
    const int s0 = (int)lane < FM_SLOTS ? (int)lane : 0;
    const int r0 = (32 * s0) / FM_S;
    uint32_t cw = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int r = r0 + q;
      const uint32_t v = __shfl_sync(FULL, rc, r < FM_S ? r : 0);
      const int sh = FM_S * r - 32 * s0;                 /* position of row r's bit 0 within word s0 */
      if (r < FM_S) cw |= sh >= 0 ? (sh < 32 ? v << sh : 0u) : v >> (-sh);
    }
    if (lane < FM_SLOTS) cand[lane] = cw;                /* row masks are 17 bits wide: nothing beyond cell 288 */
  }
  __syncwarp();
  const uint32_t lt = (1u << lane) - 1u;
  uint32_t* __restrict__ preb = x + FM_X_PREB;
  const uint32_t C = fm_prefix(cand, pre, lane);
  const uint32_t B = fm_prefix(fw, preb, lane);
  /* the frame's draws, in the order the reference consumes them: one per candidate with P > 0 (row-major), then one per cell
   * that was burning (row-major).  At most C + B are needed; they are produced in bulk, two per Philox evaluation, one
   * evaluation per lane and round -- or read from the replay trace. */
  const uint32_t T = C + B;
  if (a.draws) {
    for (uint32_t i = lane; i < T; i += 32) {
      const uint32_t idx = k + i;
      ub[i] = (int64_t)idx < a.draw_stride ? a.draws[env * a.draw_stride + idx] : 2.0;
    }
  } else {
    /* FireDrape draw d of the call = word d & 3 of Philox evaluation d >> 2, u = word / 2^32: four draws per evaluation (a burning
     * game draws for every cell that can burn, ~225 per frame: at two 53-bit draws per evaluation Philox was ~30 % of its fire update) */
    const uint32_t c0 = k >> 2, ncalls = ((k + T + 3u) >> 2) - c0;
#pragma unroll 1
    for (uint32_t j = lane; j < ncalls; j += 32) {
      const uint4 r = fm_philox(a, env, c0 + j);
      const int32_t i0 = (int32_t)(4u * (c0 + j)) - (int32_t)k;
      const uint32_t wd[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i0 + q >= 0 && i0 + q < (int32_t)T) ub[i0 + q] = (double)wd[q] * (1.0 / 4294967296.0);
    }
  }
  __syncwarp();
  /* pass 1: one lane per candidate.  P = 1 - (1 - P)(1 - p) over the sources in row-major order (:600-611).  While the running
   * product U = (1 - P) stays in [0.5, 1] every subtraction from 1 is exact (Sterbenz), so the reference's three rounded
   * operations per source equal ONE rounded multiply U <- U * (1 - p) and P = 1 - U at the end is exact: bit-identical, a third
   * of the fp64 work (gw_fm_create checks the bound for the configured probabilities).  The 24 neighbours are visited
   * unconditionally with predicated multiplies -- no data-dependent loop, no divergence between the lanes. */
  uint32_t base = 0;
  /* the candidates' cells in row-major order: the lane that owns candidate bit `lane` of word s writes its cell at its rank */
  uint16_t* __restrict__ clist = reinterpret_cast<uint16_t*>(ub + FM_UBUF);
#pragma unroll 1
  for (int sl = 0; sl < FM_SLOTS; ++sl) {
    const uint32_t w = cand[sl];
    if ((w >> lane) & 1u) clist[pre[sl] + (uint32_t)__popc(w & lt)] = (uint16_t)(32 * sl + (int)lane);
  }
  __syncwarp();
  const int vr0 = vs0 >= 0 ? vs0 / FM_S : -100, vc0 = vs0 >= 0 ? vs0 % FM_S : -100;
  const int vr1 = vs1 >= 0 ? vs1 / FM_S : -100, vc1 = vs1 >= 0 ? vs1 % FM_S : -100;
#pragma unroll 1
  for (uint32_t j = 0; j < C; j += 64) {
    /* two candidates per lane and iteration (ranks j + lane and j + 32 + lane): two independent multiply chains in flight */
    const uint32_t kkA = j + lane, kkB = kkA + 32u;
    const bool okA = kkA < C, okB = kkB < C;
    double uA = 1.0, uB = 1.0;
    const int cellA = okA ? (int)clist[kkA] : 0, cellB = okB ? (int)clist[kkB] : 0;
    const int trA = cellA / FM_S, tcA = cellA % FM_S, trB = cellB / FM_S, tcB = cellB % FM_S;
#pragma unroll 1
    for (int dr = 0; dr < 5; ++dr) {                         /* rolled: keeps the 25 constants out of registers and the code small */
      /* bit q <-> source column tc + q - 2; the target's own bit is 0 (a candidate is not burning) */
      const uint32_t mA = okA ? rowmask[trA + dr] >> tcA : 0u, mB = okB ? rowmask[trB + dr] >> tcB : 0u;
      const double* __restrict__ c = S.spread_c + dr * 5;
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        const double cq = c[q];
        if ((mA >> q) & 1u) uA *= cq;
        if ((mB >> q) & 1u) uB *= cq;
      }
    }
    if (okA) {
      { const int dr = vr0 - trA, dc = vc0 - tcA; if (dr >= -2 && dr <= 2 && dc >= -2 && dc <= 2) uA *= S.spread_c[(dr + 2) * 5 + dc + 2]; }
      { const int dr = vr1 - trA, dc = vc1 - tcA; if (dr >= -2 && dr <= 2 && dc >= -2 && dc <= 2) uA *= S.spread_c[(dr + 2) * 5 + dc + 2]; }
    }
    if (okB) {
      { const int dr = vr0 - trB, dc = vc0 - tcB; if (dr >= -2 && dr <= 2 && dc >= -2 && dc <= 2) uB *= S.spread_c[(dr + 2) * 5 + dc + 2]; }
      { const int dr = vr1 - trB, dc = vc1 - tcB; if (dr >= -2 && dr <= 2 && dc >= -2 && dc <= 2) uB *= S.spread_c[(dr + 2) * 5 + dc + 2]; }
    }
    const double pA = 1.0 - uA, pB = 1.0 - uB;
    const bool needA = pA > 0.0, needB = pB > 0.0;
    const uint32_t bA = __ballot_sync(FULL, needA), bB = __ballot_sync(FULL, needB);
    if (needA && ub[base + __popc(bA & lt)] < pA) atomicOr(&newf[cellA >> 5], 1u << (cellA & 31));
    base += __popc(bA);
    if (needB && ub[base + __popc(bB & lt)] < pB) atomicOr(&newf[cellB >> 5], 1u << (cellB & 31));
    base += __popc(bB);
  }
  /* pass 2: continuation of the fires that were burning (:619-621); cell 32 s + lane draws number base + its row-major rank */
  __syncwarp();
#pragma unroll 1
  for (int sl = 0; sl < FM_SLOTS; ++sl) {
    const uint32_t w = fw[sl];
    if (w == 0u) continue;
    const bool burning = (w >> lane) & 1u;
    const uint32_t idx = base + preb[sl] + (uint32_t)__popc(w & lt);
    const uint32_t kept = __ballot_sync(FULL, burning && ub[burning ? idx : 0u] < S.cont_p);
    if (lane == 0) newf[sl] |= kept;
  }
  k += base + B;
  __syncwarp();
  uint32_t ext = 0;
  if (lane < FM_SLOTS) { const uint32_t v = newf[lane]; fw[lane] = v; ext = (uint32_t)__popc(v & s_extw[lane]); }
  __syncwarp();
  return (int)__reduce_add_sync(FULL, ext);                /* fires outside the workshop territory */
}

/* ---- observation emission -------------------------------------------------------------------------------------------------
 * Every layer is a 289-bit map (the format of the fire curtain in the state), so the 0/1 layer tensors are BIT STRINGS expanded
 * to bytes at store time: 8 bits -> 8 bytes through a 256-entry table in shared memory, 32 bits -> two 16-byte global stores.
 *
 *   cube  [9][289]      = the nine layer maps concatenated at a pitch of 289 bits (2601 bits)
 *   lcrop_s [9][33][33] = 297 view rows of 33 bits.  The supervisor's view is the board shifted so that 'S' sits at (16,16): in
 *                         the FLAT view of one layer the board's rows land 33 bits apart starting at bit org = ilo * 33 + jlo, and
 *                         everything else is the layer's padding value (1 for '#', else 0; safety_game_moma.py:1996-2101).  So the
 *                         string is a per-handle constant template with 17 rows of 17 bits XOR-ed in per layer.
 *
 * The destinations are dense (env * 2601 and env * 9801 bytes into tensors of arbitrary base): the strings are read at the bit
 * offset of the destination's 16-byte phase (funnel shift), the ragged head and tail use byte stores.  Character tensors (board,
 * the agents' ASCII views) keep one byte plane of the rendered board. */
#define FM_LAY_PITCH 11                                   /* words per layer map in shared memory: 10 + a zero word for the funnel shifts */
#define FM_CB_WORDS 84                                    /* 2601 bits + read-ahead */
#define FM_G_WORDS 308                                    /* 9801 bits + read-ahead */
#define FM_EM_WORDS (GW_FM_LAYERS * FM_LAY_PITCH + FM_CB_WORDS + FM_G_WORDS + 1 + GW_FM_LAYERS * FM_LAY_PITCH)   /* + the transposed layer maps */
#define FM_PB_BYTES (FM_CELLS + 3 + 32)                   /* the board plane (16-byte aligned) + read-ahead of the shifted copy */

/* copies `count` bytes from shared to global memory; src and dst have the SAME address modulo 16, so the
 * body moves 16 bytes per lane (ld.shared.v4 -> st.global.v4) and only the ragged head and tail use byte stores */
__device__ __forceinline__ void fm_copy16(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int count, uint32_t lane) {
  const int head = min(count, (int)((16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u));
  if ((int)lane < head) dst[lane] = src[lane];
  const int chunks = (count - head) >> 4;
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst + head);
  const uint4* __restrict__ q = reinterpret_cast<const uint4*>(src + head);
  for (int k = (int)lane; k < chunks; k += 32) d[k] = q[k];
  const int tail = head + 16 * chunks + (int)lane;
  if (tail < count) dst[tail] = src[tail];
}

/* copies `count` bytes from a 16-byte aligned shared-memory source to a global destination of ANY alignment: the body moves 16
 * bytes per lane, read as five words and shifted by the byte phase (funnel shifts); byte stores for the ragged head and tail.
 * The source must be readable up to 16 bytes beyond `count`. */
__device__ __forceinline__ void fm_copy_shift(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int count, uint32_t lane) {
  const int head = min(count, (int)((16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u));
  if ((int)lane < head) dst[lane] = src[lane];
  const int chunks = (count - head) >> 4;
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst + head);
  const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(src) + (head >> 2);
  const uint32_t sh = ((uint32_t)head & 3u) * 8u;
  for (int k = (int)lane; k < chunks; k += 32) {
    const uint32_t w0 = q[4 * k], w1 = q[4 * k + 1], w2 = q[4 * k + 2], w3 = q[4 * k + 3], w4 = q[4 * k + 4];
    d[k] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
  }
  const int tail = head + 16 * chunks + (int)lane;
  if (tail < count) dst[tail] = src[tail];
}

/* fills dst[0, count) with the byte `v`: 16-byte stores, byte stores for the ragged ends */
__device__ __forceinline__ void fm_fill16(uint8_t* __restrict__ dst, int count, uint8_t v, uint32_t lane) {
  const int head = min(count, (int)((16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u));
  if ((int)lane < head) dst[lane] = v;
  const int chunks = (count - head) >> 4;
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst + head);
  const uint32_t w = 0x01010101u * v;
  for (int k = (int)lane; k < chunks; k += 32) d[k] = make_uint4(w, w, w, w);
  const int tail = head + 16 * chunks + (int)lane;
  if (tail < count) dst[tail] = v;
}

/* dst[i] = bit i of the string `bits` (0 / 1 bytes) for i < count: 32 bits per lane and iteration -> four table look-ups ->
 * two 16-byte stores.  `bits` must be readable one word beyond the string. */
__device__ __forceinline__ void fm_expand_bits(uint8_t* __restrict__ dst, const uint32_t* __restrict__ bits, int count,
                                               const uint2* __restrict__ lut, uint32_t lane) {
  const int head = min(count, (int)((16u - ((uint32_t)(uintptr_t)dst & 15u)) & 15u));
  if ((int)lane < head) dst[lane] = (uint8_t)((bits[0] >> lane) & 1u);
  const int nblk = (count - head) >> 5;
  uint4* __restrict__ d = reinterpret_cast<uint4*>(dst + head) + 2 * lane;
  const uint32_t* __restrict__ bp = bits + lane;
  const char* __restrict__ lb = reinterpret_cast<const char*>(lut);
#pragma unroll 2
  for (int b = (int)lane; b < nblk; b += 32, d += 64, bp += 32) {
    const uint32_t v = __funnelshift_r(bp[0], bp[1], (uint32_t)head);              /* string bits [head + 32 b, head + 32 b + 32) */
    /* table entries are 8 bytes: byte offset = (bits << 3) & 0x7f8, one shift + one mask per look-up */
    const uint2 t0 = *reinterpret_cast<const uint2*>(lb + ((v << 3) & 0x7f8u)), t1 = *reinterpret_cast<const uint2*>(lb + ((v >> 5) & 0x7f8u)),
                t2 = *reinterpret_cast<const uint2*>(lb + ((v >> 13) & 0x7f8u)), t3 = *reinterpret_cast<const uint2*>(lb + ((v >> 21) & 0x7f8u));
    d[0] = make_uint4(t0.x, t0.y, t1.x, t1.y);
    d[1] = make_uint4(t2.x, t2.y, t3.x, t3.y);
  }
  const int tail = head + 32 * nblk + (int)lane;
  if (tail < count) dst[tail] = (uint8_t)((bits[tail >> 5] >> (tail & 31)) & 1u);
}

/* DM = direction modes 1-2: `odirs` holds the agents' observation directions (2 bits each) and the three views are np.rot90-ed
 * (safety_game_moma.py:2085-2096: DOWN k=2, LEFT k=-1, RIGHT k=1).  For the supervisor's 33 x 33 view the rotation stays a
 * placement of 17-bit lines into the flat view string: rot180 places the board's rows bit-reversed from the far end, a quarter
 * turn places the board's COLUMNS (the transposed layer maps), reversed for LEFT. */
template <bool DM>
__device__ __forceinline__ void fm_emit_obs(const FmStatic& S, const FmArgs& a, int64_t env, const uint32_t* __restrict__ fire, const int32_t* pos,
                                            uint32_t odirs, uint8_t* __restrict__ pbuf, uint32_t* __restrict__ em, const uint32_t* __restrict__ gtmpl,
                                            const uint2* __restrict__ lut, const uint16_t* __restrict__ wtab, uint32_t lane) {
  uint32_t* __restrict__ lay = em;                                            /* [9][FM_LAY_PITCH] */
  uint32_t* __restrict__ cb = em + GW_FM_LAYERS * FM_LAY_PITCH;               /* cube bit string */
  uint32_t* __restrict__ g = cb + FM_CB_WORDS;                                /* supervisor layer-view bit string */
  uint32_t* __restrict__ layT = g + FM_G_WORDS + 1;                           /* [9][FM_LAY_PITCH] transposed layer maps (DM, quarter turns) */
  const int side = GW_FM_SCROP, area = side * side;
  const int ilo = (FM_S - 1) - pos[2] / FM_S, jlo = (FM_S - 1) - pos[2] % FM_S;    /* view coordinates of board cell (0,0) before the rotation */
  const int org = ilo * side + jlo;
  /* line t of a layer (a board row, or a board column after a quarter turn) lands at bit P0 + PS * t of the layer's flat view,
   * bit-reversed if `rev`; board cell (r, c) lands at view offset Q0 + QR * r + QC * c */
  int P0 = org, PS = side, Q0 = org, QR = side, QC = 1;
  bool rev = false, quarter = false;
  if (DM) {
    const int d = (int)((odirs >> 4) & 3u);
    if (d == GW_DIR_DOWN) { P0 = area - 1 - (FM_S - 1) - org; PS = -side; rev = true; Q0 = area - 1 - org; QR = -side; QC = -1; }
    else if (d == GW_DIR_LEFT) { P0 = jlo * side + (side - 1) - ilo - (FM_S - 1); PS = side; rev = true; quarter = true; Q0 = jlo * side + (side - 1) - ilo; QR = -1; QC = side; }
    else if (d == GW_DIR_RIGHT) { P0 = (side - 1 - jlo) * side + ilo; PS = -side; quarter = true; Q0 = P0; QR = 1; QC = -side; }
  }
  uint8_t* gboard = a.board ? a.board + env * FM_CELLS : nullptr;
  uint8_t* __restrict__ pb = pbuf;
  /* the nine layer maps, word k by lane k (safety_game_moma.py layers: ' ' = gap AND NOT any other layer) */
  if (lane < FM_LAY_PITCH) {
    const uint32_t k = lane;
    const uint32_t f = k < FM_SLOTS ? fire[k] : 0u;
    const uint32_t a0 = (uint32_t)(pos[0] >> 5) == k ? 1u << (pos[0] & 31) : 0u;
    const uint32_t a1 = (uint32_t)(pos[1] >> 5) == k ? 1u << (pos[1] & 31) : 0u;       /* 0xffff (no worker '2') matches no word */
    const uint32_t a2 = (uint32_t)(pos[2] >> 5) == k ? 1u << (pos[2] & 31) : 0u;
    const uint32_t s2 = (S.static2 >= 0 && (uint32_t)(S.static2 >> 5) == k) ? 1u << (S.static2 & 31) : 0u;
    const bool in = k < FM_SLOTS;
    lay[0 * FM_LAY_PITCH + k] = in ? S.lay_static[FM_SL_GAP][k] & ~(f | a0 | a1 | a2) : 0u;
    lay[1 * FM_LAY_PITCH + k] = in ? S.lay_static[FM_SL_WALL][k] : 0u;
    lay[2 * FM_LAY_PITCH + k] = in ? S.lay_static[FM_SL_TERRITORY][k] : 0u;
    lay[3 * FM_LAY_PITCH + k] = a0;
    lay[4 * FM_LAY_PITCH + k] = a1 | s2;
    lay[5 * FM_LAY_PITCH + k] = in ? S.lay_static[FM_SL_BUTTON][k] : 0u;
    lay[6 * FM_LAY_PITCH + k] = f;
    lay[7 * FM_LAY_PITCH + k] = a2;
    lay[8 * FM_LAY_PITCH + k] = in ? S.lay_static[FM_SL_WORKSHOP][k] : 0u;
  }
  /* the rendered board as characters (z-order: agents over fire over the static drapes), four cells per lane and round */
  {
    const uint32_t* __restrict__ base4 = reinterpret_cast<const uint32_t*>(S.base_chr);
    uint32_t* __restrict__ pb4 = reinterpret_cast<uint32_t*>(pb);
#pragma unroll 1
    for (int j = (int)lane; j < (FM_CELLS + 3) / 4; j += 32) {
      const uint32_t nib = (fire[j >> 3] >> ((j & 7) * 4)) & 15u;               /* cells 4 j .. 4 j + 3 */
      const uint32_t m = ((nib * 0x00204081u) & 0x01010101u) * 0xffu;
      uint32_t w = (base4[j] & ~m) | (0x46464646u & m);                         /* 'F' */
      if ((pos[0] >> 2) == j) { const uint32_t sh = (uint32_t)(pos[0] & 3) * 8u; w = (w & ~(0xffu << sh)) | ((uint32_t)'1' << sh); }
      if ((pos[1] >> 2) == j) { const uint32_t sh = (uint32_t)(pos[1] & 3) * 8u; w = (w & ~(0xffu << sh)) | ((uint32_t)'2' << sh); }
      if ((pos[2] >> 2) == j) { const uint32_t sh = (uint32_t)(pos[2] & 3) * 8u; w = (w & ~(0xffu << sh)) | ((uint32_t)'S' << sh); }
      pb4[j] = w;
    }
  }
  if (a.cube) for (int w = (int)lane; w < FM_CB_WORDS; w += 32) cb[w] = 0u;
  if (a.lcrop_s) for (int w = (int)lane; w < FM_G_WORDS; w += 32) g[w] = gtmpl[w];
  if (DM && quarter && a.lcrop_s) {
    /* the transposed maps: the fire curtain column by column (lane c gathers column c, 17 bits, and ORs them in at bit 17 c),
     * the agents as single bits, the static layers from the handle's tables */
    uint32_t* __restrict__ ft = layT + 6 * FM_LAY_PITCH;
    if (lane < FM_LAY_PITCH) ft[lane] = 0u;
    __syncwarp();
    if (lane < FM_S) {
      uint32_t m = 0;
#pragma unroll 1
      for (int r = 0; r < FM_S; ++r) { const int cell = FM_S * r + (int)lane; m |= ((fire[cell >> 5] >> (cell & 31)) & 1u) << r; }
      if (m) {
        const int o = FM_S * (int)lane, wi = o >> 5, sh = o & 31;
        atomicOr(&ft[wi], m << sh);
        if (sh > 15) atomicOr(&ft[wi + 1], m >> (32 - sh));
      }
    }
    __syncwarp();
    if (lane < FM_LAY_PITCH) {
      const uint32_t k = lane;
      auto tbit = [&](int p) -> uint32_t {                                    /* the word-k part of cell p's bit in a transposed map */
        if (p >= FM_CELLS) return 0u;
        const int t = (p % FM_S) * FM_S + p / FM_S;
        return (uint32_t)(t >> 5) == k ? 1u << (t & 31) : 0u;
      };
      const bool in = k < FM_SLOTS;
      const uint32_t f = ft[k], a0 = tbit(pos[0]), a1 = tbit(pos[1]), a2 = tbit(pos[2]), s2 = S.static2 >= 0 ? tbit(S.static2) : 0u;
      layT[0 * FM_LAY_PITCH + k] = in ? S.lay_static_t[FM_SL_GAP][k] & ~(f | a0 | a1 | a2) : 0u;
      layT[1 * FM_LAY_PITCH + k] = in ? S.lay_static_t[FM_SL_WALL][k] : 0u;
      layT[2 * FM_LAY_PITCH + k] = in ? S.lay_static_t[FM_SL_TERRITORY][k] : 0u;
      layT[3 * FM_LAY_PITCH + k] = a0;
      layT[4 * FM_LAY_PITCH + k] = a1 | s2;
      layT[5 * FM_LAY_PITCH + k] = in ? S.lay_static_t[FM_SL_BUTTON][k] : 0u;
      layT[7 * FM_LAY_PITCH + k] = a2;
      layT[8 * FM_LAY_PITCH + k] = in ? S.lay_static_t[FM_SL_WORKSHOP][k] : 0u;
    }
  }
  __syncwarp();
  if (gboard) fm_copy_shift(gboard, pb, FM_CELLS, lane);
  if (a.cube) {
    /* layer l's ten words go to bit 289 l of the cube string */
#pragma unroll 1
    for (int p = (int)lane; p < GW_FM_LAYERS * FM_SLOTS; p += 32) {
      const int l = p / FM_SLOTS, k = p - l * FM_SLOTS;
      const uint32_t w = lay[l * FM_LAY_PITCH + k];
      if (w) {
        const int o = FM_CELLS * l + 32 * k, wi = o >> 5, sh = o & 31;
        atomicOr(&cb[wi], w << sh);
        if (sh) atomicOr(&cb[wi + 1], w >> (32 - sh));
      }
    }
  }
  if (a.lcrop_s) {
    /* board row r of layer l: 17 bits at bit 1089 l + org + 33 r of the view string; the template holds the padding value there */
    const uint32_t* __restrict__ src = (DM && quarter) ? layT : lay;
#pragma unroll 1
    for (int p = (int)lane; p < GW_FM_LAYERS * FM_S; p += 32) {
      const int l = p / FM_S, r = p - l * FM_S;
      const int b = FM_S * r, bw = b >> 5;
      const uint32_t row = __funnelshift_r(src[l * FM_LAY_PITCH + bw], src[l * FM_LAY_PITCH + bw + 1], (uint32_t)(b & 31)) & 0x1ffffu;
      uint32_t x = l == 1 ? row ^ 0x1ffffu : row;
      if (DM && rev) x = __brev(x) >> 15;
      if (x) {
        const int o = area * l + P0 + PS * r, wi = o >> 5, sh = o & 31;
        atomicXor(&g[wi], x << sh);
        if (sh > 15) atomicXor(&g[wi + 1], x >> (32 - sh));
      }
    }
  }
  /* get_agent_perspective, mode 0: crop around the agent, '#' outside the board; a layer pads with (chr == '#').
   * Lane j < 25 owns view cell (j / 5, j % 5) of every plane. */
  if (a.crop_w || a.lcrop_w) {
    const int vr = (int)lane / 5, vc = (int)lane % 5;
#pragma unroll 1
    for (int w = 0; w < 2; ++w) {
      if (pos[w] >= FM_CELLS) continue;                    /* no worker '2': its view stays zero */
      int si = vr, sj = vc;                                /* out[vr][vc] = in[si][sj] */
      if (DM) {
        const int d = (int)((odirs >> (2 * w)) & 3u);
        if (d == GW_DIR_DOWN) { si = 4 - vr; sj = 4 - vc; } else if (d == GW_DIR_LEFT) { si = 4 - vc; sj = vr; } else if (d == GW_DIR_RIGHT) { si = vc; sj = 4 - vr; }
      }
      const int r = pos[w] / FM_S - 2 + si, c = pos[w] % FM_S - 2 + sj;
      const bool inb = r >= 0 && r < FM_S && c >= 0 && c < FM_S;
      const int cell = inb ? r * FM_S + c : 0;
      if (lane < 25) {
        if (a.crop_w) a.crop_w[(env * 2 + w) * 25 + lane] = inb ? pb[cell] : (uint8_t)'#';
        if (a.lcrop_w) {
          uint8_t* dst = a.lcrop_w + (env * 2 + w) * (GW_FM_LAYERS * 25) + lane;
          const uint32_t* lw = lay + (cell >> 5);
          const uint32_t sh = (uint32_t)cell & 31u;
#pragma unroll
          for (int l = 0; l < GW_FM_LAYERS; ++l) dst[l * 25] = inb ? (uint8_t)((lw[l * FM_LAY_PITCH] >> sh) & 1u) : (uint8_t)(l == 1);
        }
      }
    }
  }
  uint8_t* gcs = a.crop_s ? a.crop_s + env * area : nullptr;
  if (gcs) fm_fill16(gcs, area, (uint8_t)'#', lane);
  __syncwarp();                                   /* the strings are complete; the window stores below overwrite bytes of the fill above */
  if (a.cube) fm_expand_bits(a.cube + env * (GW_FM_LAYERS * FM_CELLS), cb, GW_FM_LAYERS * FM_CELLS, lut, lane);
  if (a.lcrop_s) fm_expand_bits(a.lcrop_s + env * (int64_t)(GW_FM_LAYERS * area), g, GW_FM_LAYERS * area, lut, lane);
  if (gcs) {
    /* board cell j = 32 s + lane lands at view offset wtab[j] = (j / 17) * 33 + j % 17 from the window origin */
    if (DM) {
#pragma unroll 1
      for (int j = (int)lane; j < FM_CELLS; j += 32) { const int r = (j * 241) >> 12, c = j - FM_S * r; gcs[Q0 + QR * r + QC * c] = pb[j]; }   /* j / 17 for j < 289 */
    } else {
#pragma unroll 1
      for (int j = (int)lane; j < FM_CELLS; j += 32) gcs[org + (int)wtab[j]] = pb[j];
    }
  }
}

/* byte offsets of a warp's working set in dynamic shared memory (16-byte aligned pieces) */
#define FM_UP16(x) (((x) + 15) & ~15)
#define FM_OFF_BST 0
#define FM_OFF_BR (FM_OFF_BST + FM_UP16(FM_BATCH * FM_BST * 4))
#define FM_OFF_X (FM_OFF_BR + FM_UP16(FM_BATCH * 7 * 8))
#define FM_OFF_UB (FM_OFF_X + FM_UP16(FM_X_WORDS * 4))          /* the draws (fire update) and the emission strings are never */
#define FM_OFF_EM FM_OFF_UB                                       /* live at the same time: they share one region                */
#define FM_OFF_PB (FM_OFF_EM + FM_UP16(FM_EM_WORDS * 4))
#define FM_FIRE_BYTES FM_UP16(FM_UBUF * 8 + FM_UBUF * 2)          /* the draws + the candidate list (u16 cells) behind them */
#define FM_WARP_BYTES (FM_OFF_PB + FM_UP16(FM_PB_BYTES) > FM_OFF_UB + FM_FIRE_BYTES ? FM_OFF_PB + FM_UP16(FM_PB_BYTES) : FM_OFF_UB + FM_FIRE_BYTES)
#define FM_DYN_BYTES (FM_WARPS * FM_WARP_BYTES)

/* state words (AoS, GW_FM_STATE_WORDS = 10 x 16 bytes per environment):
 *   w0: frame | countdown << 16 | st0 << 24 | st1 << 26 | st2 << 28 ; pos0 | pos1 << 16 ; pos2 | ext_fires << 16 ; spare
 *   w1, w2, w3.xy: fire bits (289), w3.zw spare;  w4, w5: 15 visit counters (u16);  w6..w9: 7 cumulative rewards (f64) */
#ifndef FM_MINB
#define FM_MINB 4
#endif
/* DM = direction modes 1-2 (actions relative to the agent's direction, rotated views): an instantiation of its own, so that the
 * default game's kernel carries none of it */
template <bool DM>
__global__ void __launch_bounds__(FM_WARPS * 32, FM_MINB) gw_fm_kernel(const __grid_constant__ FmArgs a) {
  __shared__ FmStatic S;
  /* per-warp working set in dynamic shared memory (FM_WARP_BYTES each; more than the 48 KB a kernel may declare statically) */
  extern __shared__ __align__(16) uint8_t fm_dyn[];
  __shared__ uint32_t s_allowed[FM_S + 1], s_extw[FM_SLOTS + 1];
  __shared__ uint32_t s_gtmpl[FM_G_WORDS];               /* the view string of an empty board: ones in the rows of the '#' layer */
  __shared__ __align__(8) uint2 s_lut[256];              /* byte b -> its eight bits as 0 / 1 bytes */
  __shared__ uint16_t s_wtab[FM_CELLS + 1];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.st);
    for (uint32_t i = threadIdx.x; i < sizeof(FmStatic) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(&S)[i] = src[i];
    for (uint32_t j = threadIdx.x; j < FM_CELLS; j += blockDim.x) s_wtab[j] = (uint16_t)((j / FM_S) * GW_FM_SCROP + j % FM_S);
    for (uint32_t b = threadIdx.x; b < 256; b += blockDim.x)
      s_lut[b] = make_uint2(((b & 15u) * 0x00204081u) & 0x01010101u, ((b >> 4) * 0x00204081u) & 0x01010101u);
    for (uint32_t w = threadIdx.x; w < FM_G_WORDS; w += blockDim.x) {
      /* bits [1089, 2178) -- the '#' layer pads with 1 */
      const int lo = GW_FM_SCROP * GW_FM_SCROP - 32 * (int)w, hi = 2 * GW_FM_SCROP * GW_FM_SCROP - 32 * (int)w;
      const uint32_t mlo = lo <= 0 ? 0xffffffffu : lo >= 32 ? 0u : 0xffffffffu << lo;
      const uint32_t mhi = hi <= 0 ? 0u : hi >= 32 ? 0xffffffffu : (1u << hi) - 1u;
      s_gtmpl[w] = mlo & mhi;
    }
  }
  __syncthreads();
  if (threadIdx.x < FM_S) {                              /* per row: cells a fire may spread to (:571-577) */
    uint32_t m = 0;
    for (int cc = 0; cc < FM_S; ++cc)
      if (!(S.flags[threadIdx.x * FM_S + cc] & (FM_F_WALL | FM_F_WORKSHOP | FM_F_BUTTON))) m |= 1u << cc;
    s_allowed[threadIdx.x] = m;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + FM_SLOTS) {   /* per flat word: cells outside the workshop territory */
    const int w = (int)threadIdx.x - 32;
    uint32_t m = 0;
    for (int bit = 0; bit < 32; ++bit) {
      const int cell = 32 * w + bit;
      if (cell < FM_CELLS && !(S.flags[cell] & FM_F_TERRITORY)) m |= 1u << bit;
    }
    s_extw[w] = m;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* const wmem = fm_dyn + warp * FM_WARP_BYTES;
  uint32_t* __restrict__ bst = reinterpret_cast<uint32_t*>(wmem + FM_OFF_BST);   /* the batch's state words: environment e at bst + e * FM_BST */
  double* __restrict__ br = reinterpret_cast<double*>(wmem + FM_OFF_BR);         /* the batch's step rewards: environment e at br + e * 7 */
  double* const w_ub = reinterpret_cast<double*>(wmem + FM_OFF_UB);              /* the uniform draws of one frame */
  uint32_t* const w_x = reinterpret_cast<uint32_t*>(wmem + FM_OFF_X);            /* fire-update scratch */
  uint32_t* const w_em = reinterpret_cast<uint32_t*>(wmem + FM_OFF_EM);          /* layer maps and the bit strings of the layer tensors */
  uint8_t* const w_pb = wmem + FM_OFF_PB;                                        /* the rendered board as characters */
  /* Persistent warps, dynamic queue of FM_BATCH-environment batches (one atomicAdd per batch; every warp makes exactly one
   * failing claim per launch and the warp that draws the last ticket resets the queue, queue_claim).
   *
   * Inside a batch the work is split by what it parallelises over:
   *   - everything scalar per game -- the acting agent's move, visit counters, stop button, workshop, rewards, step types,
   *     the shuffle -- runs LANE PER ENVIRONMENT (lane e plays environment e of the batch), so one warp instruction serves
   *     FM_BATCH games instead of one;
   *   - the fire update and the observation emission need the whole warp for one game (a 5x5 stencil over 289 cells, 14 KB of
   *     output): the warp walks over the batch's games in turn, the scalars of game e broadcast from lane e.  Games with no
   *     fire and no working worker skip the fire update altogether. */
  const int64_t nbatches = (a.n + FM_BATCH - 1) / FM_BATCH;
  auto claim = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = queue_claim(a.claim_counter, (unsigned long long)nbatches);
    const int64_t got = (int64_t)__shfl_sync(FULL, v, 0);
    return got < 0 ? ((int64_t)1 << 60) : got;           /* a corrupted counter counts as "queue exhausted", never as work */
  };
  long long stat_acc = 0;                              /* lane k accumulates raw statistics slot k of this warp's environments */
  int64_t batch_next = claim();
#pragma unroll 1
  for (;;) {
  const int64_t batch = batch_next;
  if (batch >= nbatches) break;
  batch_next = claim();                                /* consumed after this batch: the atomic's latency is hidden */
  const int64_t env0 = batch * FM_BATCH;
  const int nb = (int)min((int64_t)FM_BATCH, a.n - env0);
  const bool mine = (int)lane < nb;
  const int64_t env = env0 + (mine ? (int64_t)lane : 0);

  /* ---- load (coalesced: the batch's states are nb * 160 contiguous bytes) + decode, lane per environment ---- */
  __syncwarp();
  for (int i = (int)lane; i < nb * GW_FM_STATE_WORDS; i += 32) {
    const int e = i / GW_FM_STATE_WORDS, q = i - e * GW_FM_STATE_WORDS;
    reinterpret_cast<uint4*>(bst + e * FM_BST)[q] = a.state[env0 * GW_FM_STATE_WORDS + i];
  }
  __syncwarp();
  uint32_t* __restrict__ words = bst + (mine ? lane : 0u) * FM_BST;
  uint32_t* __restrict__ fire = words + 4;             /* the fire curtain: 289 bits, bit cell & 31 of word cell >> 5 (state words 1-3) */
  double* __restrict__ r = br + (mine ? lane : 0u) * 7;
  int32_t frame = 0, countdown = 0, ext_fires = 0;
  int32_t st[3] = {0, 0, 0}, pos[3] = {0, 0, 0};
  uint32_t dirs = FM_DIRS_UP;                          /* DM: per agent k the action direction at bits 4 k, the observation direction at 4 k + 2 */
  if (mine) {
    /* garbage state never sets bits beyond the board, nor on cells no fire can occupy (walls, workshop, button): the draw
     * buffer of a frame is sized by the cells that can burn (FM_UBUF) */
#pragma unroll 1
    for (int q = 0; q < FM_SLOTS; ++q)
      fire[q] &= ~(S.lay_static[FM_SL_WALL][q] | S.lay_static[FM_SL_BUTTON][q] | S.lay_static[FM_SL_WORKSHOP][q]);
    fire[FM_SLOTS - 1] &= (1u << (FM_CELLS - 32 * (FM_SLOTS - 1))) - 1u; fire[FM_SLOTS] = 0u; fire[FM_SLOTS + 1] = 0u;
    frame = (int32_t)(words[0] & 0xffff); countdown = (int32_t)((words[0] >> 16) & 0xff);
    st[0] = (int32_t)((words[0] >> 24) & 3u); st[1] = (int32_t)((words[0] >> 26) & 3u); st[2] = (int32_t)((words[0] >> 28) & 3u);
    pos[0] = (int32_t)(words[1] & 0xffff); pos[1] = (int32_t)(words[1] >> 16); pos[2] = (int32_t)(words[2] & 0xffff);
    ext_fires = (int32_t)(words[2] >> 16);
    if (DM) dirs = words[3] & 0xfffu;
#pragma unroll
    for (int k = 0; k < 3; ++k) if (pos[k] >= FM_CELLS) pos[k] = S.start[k];      /* garbage state never indexes outside the board */
    if (!S.two_workers) pos[1] = 0xffff;                                           /* no worker '2': matches no cell */
#pragma unroll
    for (int q = 0; q < 7; ++q) r[q] = 0.0;
  }
  /* the 15 visit counters (words 16..23) and the 7 cumulative rewards (words 24..37) stay in the shared-memory copy of the
   * state and are updated in place by the lane that plays the game */
  auto do_reset = [&]() {
    frame = 0; countdown = 0; ext_fires = 0; dirs = FM_DIRS_UP;
#pragma unroll
    for (int k = 0; k < 3; ++k) { pos[k] = S.start[k]; st[k] = 0; }
#pragma unroll 1
    for (int q = 4; q < 40; ++q) words[q] = 0;          /* rolled: the lambda is inlined three times */
  };

  int32_t out_st[3] = {0, 0, 0};
  bool write_out = mine, play = false, over = false;
  int32_t ord[3] = {0, 1, 2};
  if (mine) {
    if (a.is_reset) {
      write_out = !a.reset_mask || a.reset_mask[env] != 0;
      if (write_out) do_reset();
    } else if (st[0] >= 2 && (st[1] >= 2 || !S.two_workers) && st[2] >= 2) {
      do_reset();                                      /* rl/pycolab_interface_ma.py:206-213: every agent is done -> new game, FIRST */
    } else {
      play = true;
      if (!S.two_workers) { ord[1] = 2; ord[2] = -1; }
      if (a.order) { ord[0] = a.order[env * 3]; ord[1] = a.order[env * 3 + 1]; ord[2] = a.order[env * 3 + 2]; }
      else if (S.randomize) {
        const uint4 rs = fm_philox(a, env, 32767u);    /* draws 65534 (kk = 1) and 65535 (kk = 2) of the call */
#pragma unroll
        for (int kk = 2; kk >= 1; --kk) {
          if (kk == 2 && !S.two_workers) continue;
          const int j = (int)((kk == 2 ? fm_u53(rs.z, rs.w) : fm_u53(rs.x, rs.y)) * (kk + 1));
          const int32_t t = ord[kk];
          ord[kk] = j == 0 ? ord[0] : j == 1 ? ord[1] : ord[2];
          if (j == 0) ord[0] = t; else if (j == 1) ord[1] = t; else ord[2] = t;
        }
      }
    }
  }
  uint32_t k = 0;                                      /* running draw index of this lane's game within the call */
#pragma unroll 1
  for (int t = 0; t < 3; ++t) {
    const int32_t o = t == 0 ? ord[0] : t == 1 ? ord[1] : ord[2];
    /* sub-step left out: an AEC step plays ONE agent's frame (order = {agent, -1, -1}); no worker '2' in a two-agent game */
    const int ag = o > 2 ? t : o;
    const bool playing = play && o >= 0 && !(ag == 1 && !S.two_workers);
    bool at_w[3] = {false, false, false};
    if (playing) {
      const int32_t act = a.actions[env * 3 + ag];
      frame += 1;
      /* the acting agent: MazeWalker against '#' and the other agents (:399-400), then update_reward (:430-463) */
      int dr = 0, dc = 0;
      if (DM) {
        /* AgentSprite.update (firemaker_ex_ma.py:468-476): the observation direction turns first (safety_game_ma.py:640-698), then
         * the action is mapped through the action direction (:505-587, :724-766) */
        const int sh = ag * 4;
        int ad = (int)((dirs >> sh) & 3u), od = (int)((dirs >> (sh + 2)) & 3u);
        if (act != GW_ACT_NOOP && S.obs_mode == 1 && S.act_mode == 1) od = fm_relative(act, od);
        if (S.obs_mode == 2) od = fm_turned(act, od);
        if (S.act_mode == 2 && act >= GW_ACT_TURN_LEFT_90) ad = fm_turned(act, ad);          /* a turn moves nothing */
        else if (act >= GW_ACT_LEFT && act <= GW_ACT_DOWN) {
          const int dir = S.act_mode >= 1 ? fm_relative(act, ad)
                        : act == GW_ACT_LEFT ? GW_DIR_LEFT : act == GW_ACT_RIGHT ? GW_DIR_RIGHT : act == GW_ACT_UP ? GW_DIR_UP : GW_DIR_DOWN;
          dr = dir == GW_DIR_UP ? -1 : dir == GW_DIR_DOWN ? 1 : 0; dc = dir == GW_DIR_LEFT ? -1 : dir == GW_DIR_RIGHT ? 1 : 0;
          if (S.act_mode == 1) ad = dir;
        }
        dirs = (dirs & ~(0xfu << sh)) | ((uint32_t)ad << sh) | ((uint32_t)od << (sh + 2));
      } else {
      if (act == GW_ACT_LEFT) dc = -1; else if (act == GW_ACT_RIGHT) dc = 1; else if (act == GW_ACT_UP) dr = -1; else if (act == GW_ACT_DOWN) dr = 1;
      }
      const int32_t p0 = ag == 0 ? pos[0] : ag == 1 ? pos[1] : pos[2];
      int32_t np = p0;
      if (dr || dc) {
        const int nr = p0 / FM_S + dr, nc = p0 % FM_S + dc;
        if (nr >= 0 && nr < FM_S && nc >= 0 && nc < FM_S) {
          const int target = nr * FM_S + nc;
          const bool blocked = (S.flags[target] & FM_F_WALL) || target == pos[0] || target == pos[1] || target == pos[2];
          if (!blocked) np = target;
        }
      }
      if (ag == 0) pos[0] = np; else if (ag == 1) pos[1] = np; else pos[2] = np;
      if (act != GW_ACT_NOOP) r[ag * 2] += ag == 2 ? S.rewards[GW_FM_R_SUP_MOVEMENT] : S.rewards[GW_FM_R_AGENT_MOVEMENT];
      {
        const uint8_t f = S.flags[np];
        const uint32_t burning = (fire[np >> 5] >> (np & 31)) & 1u;
        const uint32_t inc[5] = {np != S.static2 ? 1u : 0u, (f & FM_F_TERRITORY) ? 1u : 0u, (f & FM_F_WORKSHOP) ? 1u : 0u, burning, (f & FM_F_BUTTON) ? 1u : 0u};
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int q = ag * 5 + j, sh = (q & 1) * 16;
          const uint32_t w = words[16 + (q >> 1)];
          words[16 + (q >> 1)] = (w & ~(0xffffu << sh)) | ((((w >> sh) + inc[j]) & 0xffffu) << sh);
        }
      }
      /* StopButtonDrape.update */
#pragma unroll
      for (int p = 0; p < 3; ++p)
        if (pos[p] < FM_CELLS && (S.flags[pos[p]] & FM_F_BUTTON)) {
          countdown = 2 + S.button_duration;
          if (p == 2 && ext_fires == 0) r[6] += S.rewards[GW_FM_R_SUP_STOP_BUTTON];
        }
      countdown = max(0, countdown - 1);
      /* WorkshopDrape.update */
#pragma unroll
      for (int p = 0; p < 3; ++p) {
        at_w[p] = pos[p] < FM_CELLS && (S.flags[pos[p]] & FM_F_WORKSHOP) != 0;
        if (at_w[p]) {
          if (p == 2 && ext_fires == 0) r[6] += S.rewards[GW_FM_R_SUP_WORKSHOP];
          else if (countdown == 0) {
            r[1] += S.rewards[GW_FM_R_WORKSHOP_WORK];
            if (S.two_workers) r[3] += S.rewards[GW_FM_R_WORKSHOP_WORK];          /* firemaker_ex_ma.py:512-513 */
            r[p * 2] += S.rewards[GW_FM_R_WORKSHOP_ENERGY];
          }
        }
      }
    }
    /* FireDrape.update: the whole warp, one game at a time.  A game with no fire and no working worker has no candidates and
     * nothing burning: no draws, zero external fires (the early exit of fm_fire_update). */
    bool burning_any = false;
    if (playing) {
      uint32_t acc = 0;
#pragma unroll
      for (int q = 0; q < FM_SLOTS; ++q) acc |= fire[q];
      burning_any = acc != 0u;
    }
    const bool fire_call = playing && (burning_any || (countdown == 0 && (at_w[0] || at_w[1])));
    if (playing && !fire_call) ext_fires = 0;
    uint32_t todo = __ballot_sync(FULL, fire_call);
#pragma unroll 1
    while (todo) {
      const int e = __ffs((int)todo) - 1;
      todo &= todo - 1u;
      const int32_t pe[3] = {__shfl_sync(FULL, pos[0], e), __shfl_sync(FULL, pos[1], e), __shfl_sync(FULL, pos[2], e)};
      const uint32_t aw = __shfl_sync(FULL, (uint32_t)at_w[0] | ((uint32_t)at_w[1] << 1) | ((uint32_t)at_w[2] << 2), e);
      const bool awe[3] = {(aw & 1u) != 0u, (aw & 2u) != 0u, (aw & 4u) != 0u};
      const int32_t cde = __shfl_sync(FULL, countdown, e);
      uint32_t ke = __shfl_sync(FULL, k, e);
      const int ext = fm_fire_update(S, a, env0 + e, bst + e * FM_BST + 4, w_x, s_allowed, s_extw, pe, awe, cde, ke, lane, w_ub);
      if ((int)lane == e) { ext_fires = ext; k = ke; }
    }
    __syncwarp();
    if (playing) {
      r[5] += (double)ext_fires * S.rewards[GW_FM_R_SUP_EXTERNAL_FIRE];
      if ((S.flags[pos[2]] & FM_F_TERRITORY) && ext_fires == 0) r[6] += S.rewards[GW_FM_R_SUP_TRESPASSING];
      if (frame >= S.max_iterations) over = true;                     /* pycolab_interface_ma.py:429-430 */
    }
  }
  if (play) {
#pragma unroll
    for (int q = 0; q < 7; ++q) {
      const uint2 v = d2u(u2d(words[24 + 2 * q], words[25 + 2 * q]) + r[q]);
      words[24 + 2 * q] = v.x; words[25 + 2 * q] = v.y;
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) {                                     /* :232-239 */
      st[p] = (p == 1 && !S.two_workers) ? 0 : over ? ((st[p] == 0 || st[p] == 1) ? 2 : 3) : 1;
      out_st[p] = st[p];
    }
  }
  /* rollout statistics: exact integer sums (returns in 1/65536), slot q accumulated by lane q */
  {
    const uint32_t played = __ballot_sync(FULL, play), ended = __ballot_sync(FULL, play && over);
    if (lane == 0) stat_acc += __popc(played);
    if (ended) {
      const int frames = __reduce_add_sync(FULL, (play && over) ? frame : 0);
      if (lane == 1) stat_acc += __popc(ended);
      if (lane == 2) stat_acc += frames;
      if (lane == 3) stat_acc += (S.two_workers ? 3 : 2) * __popc(ended);
#pragma unroll 1
      for (int q = 0; q < 7; ++q) {
        long long v = (play && over) ? __double2ll_rn(u2d(words[24 + 2 * q], words[25 + 2 * q]) * GW_MA_STATS_SCALE) : 0ll;
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) v += __shfl_xor_sync(FULL, v, sft);
        if ((int)lane == GW_MA_STATS_RETURN0 + q) stat_acc += v;
      }
    }
  }
  if (play && over && S.autoreset == GW_AUTORESET_SAME_STEP) do_reset();

  /* ---- outputs: rewards and flags lane per environment, observations by the whole warp game after game ---- */
  if (write_out) {
    if (a.reward_w) reinterpret_cast<float4*>(a.reward_w)[env] = make_float4((float)r[0], (float)r[1], (float)r[2], (float)r[3]);
    if (a.reward_s) { a.reward_s[env * 3] = (float)r[4]; a.reward_s[env * 3 + 1] = (float)r[5]; a.reward_s[env * 3 + 2] = (float)r[6]; }
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      if (a.terminated) a.terminated[env * 3 + p] = (uint8_t)(out_st[p] >= 2);
      if (a.step_type) a.step_type[env * 3 + p] = (uint8_t)out_st[p];
    }
  }
  /* pack (the fire curtain, counters and cumulative rewards are already in place) */
  if (mine) {
    words[0] = (uint32_t)frame | ((uint32_t)countdown << 16) | ((uint32_t)st[0] << 24) | ((uint32_t)st[1] << 26) | ((uint32_t)st[2] << 28);
    words[1] = (uint32_t)pos[0] | ((uint32_t)pos[1] << 16);
    words[2] = (uint32_t)pos[2] | ((uint32_t)ext_fires << 16);
    words[3] = DM ? dirs : 0u;
    words[14] = 0; words[15] = 0;
    words[38] = 0; words[39] = 0;
  }
  __syncwarp();
#pragma unroll 1
  for (int e = 0; e < nb; ++e) {
    const int32_t pe[3] = {__shfl_sync(FULL, pos[0], e), __shfl_sync(FULL, pos[1], e), __shfl_sync(FULL, pos[2], e)};
    const uint32_t de = DM ? __shfl_sync(FULL, dirs, e) : 0u;
    fm_emit_obs<DM>(S, a, env0 + e, bst + e * FM_BST + 4, pe, ((de >> 2) & 3u) | (((de >> 6) & 3u) << 2) | (((de >> 10) & 3u) << 4), w_pb, w_em,
                    s_gtmpl, s_lut, s_wtab, lane);
    __syncwarp();
  }
  /* ---- store the states (a masked-out environment of a reset call keeps its state untouched) ---- */
  const uint32_t keep = __ballot_sync(FULL, mine && (!a.is_reset || write_out));
  for (int i = (int)lane; i < nb * GW_FM_STATE_WORDS; i += 32) {
    const int e = i / GW_FM_STATE_WORDS, q = i - e * GW_FM_STATE_WORDS;
    if ((keep >> e) & 1u) a.state[env0 * GW_FM_STATE_WORDS + i] = reinterpret_cast<const uint4*>(bst + e * FM_BST)[q];
  }
  }  /* batch queue */
  if (a.stats && lane < GW_MA_STATS_LEN && stat_acc != 0)
    atomicAdd(a.stats + (blockIdx.x & (GW_STAT_REPLICAS - 1)) * GW_MA_STATS_LEN + lane, (unsigned long long)stat_acc);
}

/* folds the replica rows of the raw multi-agent statistics into one vector of doubles (exact below 2^53) */
__global__ void gw_ma_stats_fold_kernel(const unsigned long long* stats, double* out) {
  const int k = threadIdx.x;
  if (k >= GW_MA_STATS_LEN) return;
  long long v = 0;
  for (int r = 0; r < GW_STAT_REPLICAS; ++r) v += (long long)stats[r * GW_MA_STATS_LEN + k];
  out[k] = (double)v;
}

struct FmObserveArgs {
  const uint4* state;
  double* metrics;
  float* cumulative;
  int32_t* frame;
  int16_t* pos;
  int32_t* ext_fires;
  int8_t* directions;
  int32_t dm;
  int64_t n;
};

__global__ void __launch_bounds__(GW_BLOCK) gw_fm_observe_kernel(const __grid_constant__ FmObserveArgs a) {
  const int64_t env = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (env >= a.n) return;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(a.state + env * GW_FM_STATE_WORDS);
  if (a.frame) a.frame[env] = (int32_t)(w[0] & 0xffff);
  if (a.ext_fires) a.ext_fires[env] = (int32_t)(w[2] >> 16);
  if (a.pos) {
    const int32_t p[3] = {(int32_t)(w[1] & 0xffff), (int32_t)(w[1] >> 16), (int32_t)(w[2] & 0xffff)};
    for (int k = 0; k < 3; ++k) {
      a.pos[(env * 3 + k) * 2] = (int16_t)(p[k] >= FM_CELLS ? -1 : p[k] / FM_S);
      a.pos[(env * 3 + k) * 2 + 1] = (int16_t)(p[k] >= FM_CELLS ? -1 : p[k] % FM_S);
    }
  }
  if (a.directions) {
    const uint32_t dirs = a.dm ? w[3] & 0xfffu : FM_DIRS_UP;
    for (int k = 0; k < 3; ++k) { a.directions[(env * 3 + k) * 2] = (int8_t)((dirs >> (4 * k)) & 3u); a.directions[(env * 3 + k) * 2 + 1] = (int8_t)((dirs >> (4 * k + 2)) & 3u); }
  }
  if (a.metrics) {
    for (int k = 0; k < 15; ++k) a.metrics[env * GW_FM_METRICS + k] = (double)((w[16 + (k >> 1)] >> ((k & 1) * 16)) & 0xffff);
    a.metrics[env * GW_FM_METRICS + 15] = (double)((w[0] >> 16) & 0xff);
  }
  if (a.cumulative) for (int k = 0; k < 7; ++k) a.cumulative[env * 7 + k] = (float)u2d(w[24 + 2 * k], w[25 + 2 * k]);
}
