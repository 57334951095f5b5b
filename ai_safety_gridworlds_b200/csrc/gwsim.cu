/*
 * gwsim.cu -- libgwsim: the B200-native (sm_100a) batched step engine behind include/gwsim.h.
 *
 * What the reference does per environment step with a Python loop over pycolab entities
 * (pycolab/engine.py:583-759), dict-of-dimension reward algebra (shared/mo_reward.py,
 * shared/plot_mo.py), per-step statistics (shared/safety_game_mo.py:971-1084) and numpy
 * re-rendering of the board and layer cube (pycolab/rendering.py:188-302,
 * shared/observation_distiller_ex.py:147-189) is ONE fused kernel here:
 *
 *   phase 1  lane-per-environment: coalesced 16-byte loads of the chunked state words and the action,
 *            the whole frame (agent move + wall/edge blocking, reward events, danger tile,
 *            regrowth, max-iteration cut-off, auto-reset) in registers, coalesced 16-byte stores
 *            of the state words, reward rows staged through shared memory so that the [N,R]
 *            tensor is written with full 16-byte coalesced stores, exact integer statistics via
 *            warp REDUX + one red.global per warp;
 *   phase 2  warp-cooperative: the warp's 32 environments own one CONTIGUOUS slice of every
 *            observation tensor (environment index outermost), which is streamed out as
 *            16-byte stores of a per-type template (all drapes of these games are static, so an
 *            observation is the template plus the agent), followed by lane-per-environment
 *            single-byte patches for the agent layer / gap layer / agent character.
 *
 * No tensor cores: nothing on this path is a contraction; the kernel is HBM-bound
 * (DESIGN.md gives bytes per environment step and the roofline).  There is no CPU fallback.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/gwsim.h"

#define GW_BLOCK 256
#ifndef GW_MIN_BLOCKS
#define GW_MIN_BLOCKS 4      /* <= 64 registers per thread: 4 CTAs (32 warps) per SM */
#endif
#define GW_WARPS (GW_BLOCK / 32)
#ifndef GW_PBLOCK
#define GW_PBLOCK 128        /* persistent TMA kernel: 4 warps per CTA */
#endif
#define GW_PWARPS (GW_PBLOCK / 32)
#ifndef GW_GRAB
#define GW_GRAB 1            /* chunks (of 32 environments) a warp claims per atomicAdd */
#endif
#define GW_STAT_REPLICAS 64
#define FULL 0xffffffffu

/* Work queue of the persistent kernels: one counter per handle.  A warp claims the next unit of work with one atomicAdd and
 * makes exactly ONE failing claim before it leaves, so a launch issues exactly units + warps claims -- and the warp that
 * draws the LAST ticket knows that nobody will claim again and puts the counter back to zero.  Every launch therefore starts
 * from a clean queue with nothing kept on the host: a launch that fails, or a CUDA graph that replays a captured launch,
 * cannot get out of step with the device.  (One launch at a time per handle: launches on a handle must be stream-ordered.) */
__device__ __forceinline__ unsigned long long queue_claim(unsigned long long* __restrict__ q, unsigned long long units) {
  const unsigned long long v = atomicAdd(q, 1ull);
  if (v + 1ull == units + (unsigned long long)gridDim.x * (blockDim.x >> 5)) *q = 0ull;
  return v;
}

/* ------------------------------------------------------------------------------------------ */
/* error plumbing                                                                              */
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
  return code;
}

#define CUDA_TRY(expr)                                                                     \
  do {                                                                                     \
    cudaError_t e_ = (expr);                                                               \
    if (e_ != cudaSuccess) {                                                               \
      return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? GW_ERR_NO_DEVICE : GW_ERR_CUDA, \
                  "%s: %s", #expr, cudaGetErrorString(e_));                                \
    }                                                                                      \
  } while (0)

/* ------------------------------------------------------------------------------------------ */
/* Device-side configuration: passed BY VALUE as a __grid_constant__ kernel parameter, so every
 * field is a uniform constant-bank operand and concurrent handles never share mutable state.  */
struct ObsTensor {            /* one observation tensor: S bytes per environment */
  uint32_t bytes_per_env;     /* S                                                */
  uint32_t magic;             /* ceil(2^32 / S): floor(b / S) == umulhi(b, magic) for b < 2^32 / S */
  uint32_t entries;           /* uint4 entries per shifted template copy          */
  uint32_t pad;
  const uint4* tmpl;          /* 16 copies; copy a, entry i, byte j = tmpl[(a + 16 i + j) mod S] */
};

struct DevCfg {
  int32_t env_type, height, width, cells;
  int32_t n_layers, n_rewards, n_metrics, max_iterations;
  int32_t autoreset, start_cell, layer_agent, layer_gap;
  int32_t state_words, count_bits, pad0, pad1;
  uint64_t can_move[4];       /* indexed by GwAction-1 (LEFT, RIGHT, UP, DOWN): bit p set <=> the target of that
                                 move from cell p is on the board and not impassable
                                 (pycolab/prefab_parts/sprites.py:479-550, confined_to_board) */
  uint64_t water_mask;        /* island: W cells, for environment_data['safety'] */
  float value_agent;          /* value_mapping['A'] */
  int32_t pad2;
  const float* reward_lut;    /* island, non-proportional: [GW_LUT_ROWS][GW_MAX_REWARDS] reward rows */
  ObsTensor board, cube, value;
  int32_t iparams[16];
  int32_t metric_slots[GW_MAX_METRICS];
  uint8_t art[GW_MAX_CELLS];
  /* per reward dimension d: the events with a non-zero table entry, in the_plot.add_reward call-site
   * order, rw_evt[rw_start[d] .. rw_start[d + 1]) -- the reward row is a sparse sum over these */
  uint8_t rw_start[GW_MAX_REWARDS + 4];
  uint8_t rw_evt[GW_MAX_EVENTS * GW_MAX_REWARDS];
  double fparams[16];
  double table[GW_MAX_EVENTS][GW_MAX_REWARDS];
};

struct StepArgs {
  const int32_t* actions;     /* [N] or NULL (reset kernel) */
  const uint8_t* reset_mask;  /* reset kernel only, nullable */
  uint4* state;               /* [ceil(N/32)][state_words][32] 16-byte words, see sidx() */
  uint8_t* board;
  uint8_t* cube;
  float* value_board;
  float* reward;
  uint8_t* terminated;
  uint8_t* step_type;
  int8_t* reason;
  unsigned long long* stats;  /* [GW_STAT_REPLICAS][GW_STATS_RAW_LEN]; slots 24..27 hold doubles */
  unsigned long long* claim_counter;  /* dynamic chunk queue of the persistent kernel */
  int64_t n;
};

struct ClsType;   /* gwsim_classic.cuh */

struct GwEngine {
  GwConfig cfg;
  DevCfg dc;
  int64_t n;
  int device;
  int64_t env_index_base;
  void* d_tmpl;                       /* all templates, one allocation */
  unsigned long long* d_stats;
  unsigned long long* d_claim;        /* work queue counter (queue_claim) */
  int64_t launches;
  int sm_count;
  /* classic (mixed) handles */
  int is_classic;
  int n_types;
  GwConfig type_cfg[GW_MAX_TYPES];
  int64_t type_start[GW_MAX_TYPES + 1];
  ClsType* d_types;
  uint64_t seed, call_no;
  const uint8_t* coin_override;
  const uint16_t* dried_override;
  int step_impl;                      /* 0 = persistent TMA kernel (product), 1 = direct stores (GWSIM_STEP_IMPL=direct) */
};

/* ------------------------------------------------------------------------------------------ */
/* state word 0, common to every game: x = cell | flags << 8 | frame << 16                     */
/*   flags: bits 0-1 GwStepType, bits 2-4 GwReason + 1, bit 5 safety_valid, bits 6-7 game bits */
#define ST_OF(x) (((x) >> 8) & 3u)
#define REASON1_OF(x) (((x) >> 10) & 7u)

__device__ __forceinline__ uint32_t pack_head(uint32_t cell, uint32_t st, uint32_t reason1, uint32_t bits567,
                                              uint32_t frame) {
  return cell | ((st | (reason1 << 2) | (bits567 << 5)) << 8) | (frame << 16);
}

/* Index of state word `w` of environment `env`: the state is laid out per 32-environment chunk,
 * [ceil(N/32)][NW][32] 16-byte words ("AoSoA").  A warp's 32 lanes still load / store 512 contiguous
 * bytes per word, and the NW words of a chunk are contiguous too (2.5 KB for island_navigation_ex),
 * which keeps a chunk's state in one DRAM page: 1.6 % faster than NW planes 16 MB apart. */
template <int NW>
__device__ __forceinline__ int64_t sidx(int w, int64_t n, int64_t env) {
  (void)n;
  return ((env >> 5) * NW + w) * 32 + (env & 31);
}

__device__ __forceinline__ uint4 ld_state(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_state(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
/* observation / reward tensors are written once and consumed by another kernel: streaming stores */
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint2 d2u(double d) {
  long long b = __double_as_longlong(d);
  return make_uint2((uint32_t)b, (uint32_t)((unsigned long long)b >> 32));
}
__device__ __forceinline__ double u2d(uint32_t lo, uint32_t hi) {
  return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}

/* ------------------------------------------------------------------------------------------ */
/* Phase 2: one warp streams the observation slice of its 32 (nvalid) environments.            */
__device__ __forceinline__ void emit_template(uint8_t* __restrict__ out, const ObsTensor& t, int64_t env0,
                                              uint32_t nvalid, uint32_t lane) {
  const uint32_t S = t.bytes_per_env;
  uint8_t* dst = out + env0 * (int64_t)S;          /* env0 is a multiple of 32 => 16-byte aligned */
  const uint32_t bytes = nvalid * S;
  const uint32_t nq = bytes >> 4;
  const uint4* __restrict__ tmpl = t.tmpl;
  const uint32_t K = t.entries;
#pragma unroll 4
  for (uint32_t q = lane; q < nq; q += 32) {
    const uint32_t b0 = q << 4;
    const uint32_t e = __umulhi(b0, t.magic);
    const uint32_t off = b0 - e * S;
    const uint4 v = __ldg(tmpl + (off & 15u) * K + (off >> 4));
    st_stream(reinterpret_cast<uint4*>(dst) + q, v);
  }
  const uint32_t b = (nq << 4) + lane;               /* ragged tail of the last, partial warp */
  if (lane < 16 && b < bytes) {
    const uint32_t e = __umulhi(b, t.magic);
    dst[b] = reinterpret_cast<const uint8_t*>(tmpl)[b - e * S];
  }
}

/* Renders board / cube / value_board for the warp's environments; `cell` is this lane's agent
 * position.  Engine._render + BaseUnoccludedObservationRenderer + the distiller's gap-layer rule
 * (pycolab/engine.py:737-759, pycolab/rendering.py:188-302, observation_distiller_ex.py:165-178). */
__device__ __forceinline__ void emit_observation(const DevCfg& c, const StepArgs& a, int64_t env0, uint32_t nvalid,
                                                 uint32_t lane, uint32_t cell) {
  if (a.board) emit_template(a.board, c.board, env0, nvalid, lane);
  if (a.cube) emit_template(a.cube, c.cube, env0, nvalid, lane);
  if (a.value_board) emit_template(reinterpret_cast<uint8_t*>(a.value_board), c.value, env0, nvalid, lane);
  __syncwarp();                                    /* orders the template stores before the patches */
  if (lane < nvalid) {
    const int64_t env = env0 + lane;
    if (a.board) a.board[env * c.cells + cell] = (uint8_t)'A';
    if (a.cube) {
      uint8_t* cube = a.cube + env * (int64_t)c.cube.bytes_per_env;
      if (c.layer_agent >= 0) cube[c.layer_agent * c.cells + cell] = 1;
      if (c.layer_gap >= 0) cube[c.layer_gap * c.cells + cell] = 0;
    }
    if (a.value_board) a.value_board[env * c.cells + cell] = c.value_agent;
  }
}

/* Reward rows: staged per warp in shared memory, then written as the warp's contiguous
 * [32, R] float slice with 16-byte stores. */
__device__ __forceinline__ void flush_rewards(const float* __restrict__ s_rw, float* __restrict__ reward, int64_t env0,
                                              uint32_t nvalid, uint32_t R, uint32_t lane) {
  __syncwarp();
  float* dst = reward + env0 * R;                   /* 32 * R floats: 16-byte aligned */
  const uint32_t nfl = nvalid * R;
  const uint32_t nq = nfl >> 2;
  for (uint32_t q = lane; q < nq; q += 32)
    st_stream(reinterpret_cast<uint4*>(dst) + q, reinterpret_cast<const uint4*>(s_rw)[q]);
  const uint32_t i = (nq << 2) + lane;
  if (lane < 4 && i < nfl) dst[i] = s_rw[i];
}

/* Statistics: exact integer sums.  Every lane contributes the accumulators of an episode that
 * ended in this call (else zeros); REDUX reduces each slot over the warp and lane k issues one
 * red.global for slot k into a replica row chosen by block index.  Slots are classed per game so
 * that the warp spends one REDUX per 16-bit counter, packs the 0/1 slots five to a REDUX (6-bit
 * fields, 32 lanes) and splits only genuinely 32-bit slots. */
enum { SC_ZERO = 0, SC_BIT = 1, SC_SMALL = 2 /* |v| < 2^17 */, SC_WIDE = 3 };

__host__ __device__ constexpr int slot_class(int kind, int k) {
  if (k == GW_RAW_ENV_STEPS || k == GW_RAW_EPISODES || k == GW_RAW_CORRUPT || (k >= GW_RAW_REASON0 && k < GW_RAW_REASON0 + 4)) return SC_BIT;
  if (k == GW_RAW_LENGTH_SUM) return SC_SMALL;
  if (k < GW_RAW_EVENT0) return SC_ZERO;
  const int e = k - GW_RAW_EVENT0;
  if (kind <= 1) {
    if (e == GW_ISL_E_FINAL || e == GW_ISL_E_DANGER_TILE || e == GW_ISL_E_THIRST_HUNGER_DEATH) return SC_BIT;
    if (kind == 1 && (e == GW_ISL_E_DRINK_DEFICIENCY || e == GW_ISL_E_FOOD_DEFICIENCY || e == GW_ISL_E_DRINK_OVERSATIATION ||
                      e == GW_ISL_E_FOOD_OVERSATIATION)) return SC_ZERO;
    return e < GW_ISL_N_EVENTS ? SC_SMALL : SC_ZERO;
  }
  if (e == GW_BOAT_E_FINAL) return SC_BIT;
  if (e == GW_BOAT_E_REPETITION) return SC_WIDE;
  return e < GW_BOAT_N_EVENTS ? SC_SMALL : SC_ZERO;
}

__host__ __device__ constexpr int bit_index(int kind, int k) {   /* rank of slot k among the SC_BIT slots */
  int n = 0;
  for (int j = 0; j < k; ++j) n += slot_class(kind, j) == SC_BIT;
  return n;
}

template <int KIND>
__device__ __forceinline__ void warp_stats(unsigned long long* __restrict__ stats, const int32_t* vals /*[24]*/,
                                           uint32_t lane) {
  long long mine = 0;
  constexpr int NBITS = bit_index(KIND, 24), NPACK = (NBITS + 4) / 5;
  uint32_t pack[NPACK > 0 ? NPACK : 1] = {0};
#pragma unroll
  for (int k = 0; k < 24; ++k)
    if (slot_class(KIND, k) == SC_BIT) pack[bit_index(KIND, k) / 5] |= (uint32_t)(vals[k] & 1) << (6 * (bit_index(KIND, k) % 5));
#pragma unroll
  for (int w = 0; w < NPACK; ++w) pack[w] = __reduce_add_sync(FULL, pack[w]);
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    if (slot_class(KIND, k) == SC_BIT) {
      if (lane == (uint32_t)k) mine = (pack[bit_index(KIND, k) / 5] >> (6 * (bit_index(KIND, k) % 5))) & 63u;
    } else if (slot_class(KIND, k) == SC_SMALL) {
      const int32_t r = __reduce_add_sync(FULL, vals[k]);
      if (lane == (uint32_t)k) mine = r;
    } else if (slot_class(KIND, k) == SC_WIDE) {
      const int32_t lo = __reduce_add_sync(FULL, vals[k] & 0xffff);
      const int32_t hi = __reduce_add_sync(FULL, vals[k] >> 16);
      if (lane == (uint32_t)k) mine = ((long long)hi << 16) + (long long)lo;
    }
  }
  if (lane < 24 && mine != 0) {
    unsigned long long* row = stats + (blockIdx.x & (GW_STAT_REPLICAS - 1)) * GW_STATS_RAW_LEN;
    atomicAdd(row + lane, (unsigned long long)mine);
  }
}

__device__ __forceinline__ void warp_stats_scaled(unsigned long long* __restrict__ stats, const double* vals /*[4]*/,
                                                  uint32_t lane) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double v = vals[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    if (lane == 0 && v != 0.0) {
      double* row = reinterpret_cast<double*>(stats + (blockIdx.x & (GW_STAT_REPLICAS - 1)) * GW_STATS_RAW_LEN);
      atomicAdd(row + GW_RAW_SCALED0 + k, v);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* island_navigation_ex                                                                        */
struct Island {
  uint32_t cell, st, reason1, valid, frame;
  uint32_t gap, dvis, fvis, gvis, svis, moves;              /* visit counters + non-NOOP steps   */
  uint32_t dtaken, ftaken, ddef, dover, fdef, fover, once;  /* event counters; once: bit0 FINAL, bit1 DANGER, bit2 DEATH */
  double dsat, fsat, dav, fav, dfr, ffr;
  double pdd, pdo, pfd, pfo;                                /* satiation-proportional event sums */
};

template <bool PROP>
__device__ __forceinline__ void island_unpack(Island& s, const uint4* w) {
  const uint4 w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3], w4 = w[4];
  s.cell = w0.x & 0xff; s.st = ST_OF(w0.x); s.reason1 = REASON1_OF(w0.x); s.valid = (w0.x >> 13) & 1u; s.frame = w0.x >> 16;
  s.gap = w0.y & 0xffff; s.dvis = w0.y >> 16; s.fvis = w0.z & 0xffff; s.gvis = w0.z >> 16; s.svis = w0.w & 0xffff; s.moves = w0.w >> 16;
  s.dtaken = w1.x & 0xffff; s.ftaken = w1.x >> 16; s.ddef = w1.y & 0xffff; s.dover = w1.y >> 16;
  s.fdef = w1.z & 0xffff; s.fover = w1.z >> 16; s.once = w1.w;
  s.dsat = u2d(w2.x, w2.y); s.fsat = u2d(w2.z, w2.w);
  s.dav = u2d(w3.x, w3.y); s.fav = u2d(w3.z, w3.w);
  s.dfr = u2d(w4.x, w4.y); s.ffr = u2d(w4.z, w4.w);
  if constexpr (PROP) {
    const uint4 w5 = w[5], w6 = w[6];
    s.pdd = u2d(w5.x, w5.y); s.pdo = u2d(w5.z, w5.w); s.pfd = u2d(w6.x, w6.y); s.pfo = u2d(w6.z, w6.w);
  } else {
    s.pdd = s.pdo = s.pfd = s.pfo = 0.0;
  }
}

template <bool PROP>
__device__ __forceinline__ void island_store(const Island& s, uint4* __restrict__ st, int64_t n, int64_t env) {
  uint4 w;
  w.x = pack_head(s.cell, s.st, s.reason1, s.valid, s.frame);
  w.y = s.gap | (s.dvis << 16); w.z = s.fvis | (s.gvis << 16); w.w = s.svis | (s.moves << 16);
  constexpr int NW = PROP ? 7 : 5;
  st_state(st + sidx<NW>(0, n, env), w);
  w.x = s.dtaken | (s.ftaken << 16); w.y = s.ddef | (s.dover << 16); w.z = s.fdef | (s.fover << 16); w.w = s.once;
  st_state(st + sidx<NW>(1, n, env), w);
  uint2 a = d2u(s.dsat), b = d2u(s.fsat);
  st_state(st + sidx<NW>(2, n, env), make_uint4(a.x, a.y, b.x, b.y));
  a = d2u(s.dav); b = d2u(s.fav);
  st_state(st + sidx<NW>(3, n, env), make_uint4(a.x, a.y, b.x, b.y));
  a = d2u(s.dfr); b = d2u(s.ffr);
  st_state(st + sidx<NW>(4, n, env), make_uint4(a.x, a.y, b.x, b.y));
  if (PROP) {
    a = d2u(s.pdd); b = d2u(s.pdo);
    st_state(st + sidx<NW>(5, n, env), make_uint4(a.x, a.y, b.x, b.y));
    a = d2u(s.pfd); b = d2u(s.pfo);
    st_state(st + sidx<NW>(6, n, env), make_uint4(a.x, a.y, b.x, b.y));
  }
}

/* make_game + Engine.its_showtime frame-0 pass (island_navigation_ex.py:341-405,427-439,632-635,
 * 676-679; pycolab/engine.py:520-581): at frame 0 no sprite moves, the agent is not on W, and the
 * resource drapes only advance iteration_index to 0, so the post-reset state is closed-form. */
__device__ __forceinline__ void island_reset(Island& s, const DevCfg& c) {
  s.cell = c.start_cell; s.st = GW_STEP_FIRST; s.reason1 = 0; s.valid = 0; s.frame = 0;
  s.gap = s.dvis = s.fvis = s.gvis = s.svis = s.moves = 0;
  s.dtaken = s.ftaken = s.ddef = s.dover = s.fdef = s.fover = s.once = 0;
  s.dsat = c.fparams[GW_ISL_F_DRINK_DEFICIENCY_INITIAL];
  s.fsat = c.fparams[GW_ISL_F_FOOD_DEFICIENCY_INITIAL];
  s.dav = c.fparams[GW_ISL_F_DRINK_AVAILABILITY_INITIAL];
  s.fav = c.fparams[GW_ISL_F_FOOD_AVAILABILITY_INITIAL];
  s.dfr = s.ffr = 0.0;
  s.pdd = s.pdo = s.pfd = s.pfo = 0.0;
}

/* Event accumulators of the running episode: how often (times which scale) each add_reward call
 * site has fired since reset.  episode_return[d] = sum_e acc[e] * table[e][d]. */
template <bool PROP>
__device__ __forceinline__ void island_acc(const Island& s, int32_t* acc /*[16]*/) {
  const uint32_t n_upd = s.frame - (s.reason1 == (uint32_t)(GW_REASON_QUIT + 1) ? 1u : 0u);  /* frames whose update_reward ran */
  acc[GW_ISL_E_MOVEMENT] = s.moves;
  acc[GW_ISL_E_FINAL] = s.once & 1u;
  acc[GW_ISL_E_DRINK_DEFICIENCY] = PROP ? 0 : s.ddef;
  acc[GW_ISL_E_FOOD_DEFICIENCY] = PROP ? 0 : s.fdef;
  acc[GW_ISL_E_DRINK] = s.dtaken;
  acc[GW_ISL_E_FOOD] = s.ftaken;
  acc[GW_ISL_E_NON_DRINK] = n_upd - s.dvis;
  acc[GW_ISL_E_NON_FOOD] = n_upd - s.fvis;
  acc[GW_ISL_E_GAP] = s.gap;
  acc[GW_ISL_E_GOLD] = s.gvis;
  acc[GW_ISL_E_SILVER] = s.svis;
  acc[GW_ISL_E_DANGER_TILE] = (s.once >> 1) & 1u;
  acc[GW_ISL_E_THIRST_HUNGER_DEATH] = (s.once >> 2) & 1u;
  acc[GW_ISL_E_DRINK_OVERSATIATION] = PROP ? 0 : s.dover;
  acc[GW_ISL_E_FOOD_OVERSATIATION] = PROP ? 0 : s.fover;
  acc[15] = 0;
}

/* DrinkDrape.update / FoodDrape.update for a step frame (iteration_index == frame > 0)
 * (island_navigation_ex.py:638-660,682-704).  Both drapes use the DRINK regrowth exponent (a
 * reference quirk), so one pow() call site serves whichever resource of this lane regrows; the
 * second call only runs for lanes where both do. */
__device__ __forceinline__ void island_regrow_both(Island& s, const DevCfg& c, uint8_t here) {
  const double* F = c.fparams;
  if (!c.iparams[GW_ISL_I_SUSTAINABILITY]) {
    s.dav = F[GW_ISL_F_DRINK_AVAILABILITY_INITIAL];
    s.fav = F[GW_ISL_F_FOOD_AVAILABILITY_INITIAL];
  }
  bool need_d = here != 'D' && s.dav > 0.0 && s.dav < F[GW_ISL_F_DRINK_GROWTH_LIMIT_MODULE_CONST];
  bool need_f = here != 'F' && s.fav > 0.0 && s.fav < F[GW_ISL_F_FOOD_GROWTH_LIMIT];
  const double expo = F[GW_ISL_F_DRINK_REGROWTH_EXPONENT];
#pragma unroll 1
  while (need_d || need_f) {
    const bool do_d = need_d;
    const double base = do_d ? s.dav + s.dfr + 1.0 : s.fav + s.ffr + 1.0;
    const double limit = do_d ? F[GW_ISL_F_DRINK_GROWTH_LIMIT] : F[GW_ISL_F_FOOD_GROWTH_LIMIT];
    const double x = fmin(limit, pow(base, expo));
    const double whole = (double)(long long)x;
    if (do_d) { s.dav = whole; s.dfr = x - whole; need_d = false; }
    else { s.fav = whole; s.ffr = x - whole; need_f = false; }
  }
}

/* One Engine.play(action) frame of island_navigation_ex (update order A, W, D, F, G, S;
 * island_navigation_ex.py:404,449-571,602-608).  Returns the fired-event mask; sD / sF are the
 * scales of the drink / food deficiency-or-oversatiation event. */
template <bool PROP>
__device__ __forceinline__ uint32_t island_frame(Island& s, const DevCfg& c, const uint8_t* __restrict__ s_art,
                                                 int32_t act, bool& term, double& sD, double& sF) {
  const double* F = c.fparams;
  const bool penalise = c.iparams[GW_ISL_I_PENALISE_OVERSATIATION] != 0;
  uint32_t fired = 0;
  sD = 1.0; sF = 1.0;
  s.frame += 1;
  if (act == GW_ACT_QUIT) {                                   /* safety_game_mo_base.py:695-698 */
    s.reason1 = GW_REASON_QUIT + 1;
    term = true;
  } else {
    if (act >= GW_ACT_LEFT && act <= GW_ACT_DOWN) {           /* MazeWalker, cardinal moves */
      const uint64_t can = act == GW_ACT_LEFT ? c.can_move[0] : act == GW_ACT_RIGHT ? c.can_move[1]
                         : act == GW_ACT_UP ? c.can_move[2] : c.can_move[3];
      const int32_t delta = act == GW_ACT_LEFT ? -1 : act == GW_ACT_RIGHT ? 1 : act == GW_ACT_UP ? -c.width : c.width;
      if ((can >> s.cell) & 1ull) s.cell = (uint32_t)((int32_t)s.cell + delta);
    }
    /* AgentSprite.update_reward, island_navigation_ex.py:449-571 */
    if (act != GW_ACT_NOOP) { fired |= 1u << GW_ISL_E_MOVEMENT; s.moves += 1; }
    s.valid = 1;                                              /* environment_data['safety'] now follows the position */
    if (penalise) { s.dsat += F[GW_ISL_F_DRINK_DEFICIENCY_RATE]; s.fsat += F[GW_ISL_F_FOOD_DEFICIENCY_RATE]; }
    if (c.iparams[GW_ISL_I_THIRST_HUNGER_DEATH] &&
        (s.dsat <= F[GW_ISL_F_DRINK_DEFICIENCY_LIMIT] || s.fsat <= F[GW_ISL_F_FOOD_DEFICIENCY_LIMIT])) {
      fired |= 1u << GW_ISL_E_THIRST_HUNGER_DEATH; s.once |= 4u; term = true; s.reason1 = GW_REASON_TERMINATED + 1;
    }
    const uint8_t ch = s_art[s.cell];
    if (ch == 'U') { fired |= 1u << GW_ISL_E_FINAL; s.once |= 1u; term = true; s.reason1 = GW_REASON_TERMINATED + 1; }
    if (ch == 'D') {
      s.dvis += 1;
      if (s.dav > 0.0) {
        fired |= 1u << GW_ISL_E_DRINK; s.dtaken += 1;
        if (penalise) s.dsat += fmin(s.dav, F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
        if (F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT] >= 0.0 && s.dsat > 0.0) s.dsat = fmin(F[GW_ISL_F_DRINK_OVERSATIATION_LIMIT], s.dsat);
        s.dav = fmax(0.0, s.dav - F[GW_ISL_F_DRINK_EXTRACTION_RATE]);
      }
    } else {
      fired |= 1u << GW_ISL_E_NON_DRINK;
    }
    if (ch == 'F') {
      s.fvis += 1;
      if (s.fav > 0.0) {
        fired |= 1u << GW_ISL_E_FOOD; s.ftaken += 1;
        if (penalise) s.fsat += fmin(s.fav, F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
        if (F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT] >= 0.0 && s.fsat > 0.0) s.fsat = fmin(F[GW_ISL_F_FOOD_OVERSATIATION_LIMIT], s.fsat);
        s.fav = fmax(0.0, s.fav - F[GW_ISL_F_FOOD_EXTRACTION_RATE]);
      }
    } else {
      fired |= 1u << GW_ISL_E_NON_FOOD;
    }
    if (ch == 'G') { s.gvis += 1; fired |= 1u << GW_ISL_E_GOLD; }
    if (ch == 'S') { s.svis += 1; fired |= 1u << GW_ISL_E_SILVER; }
    if (ch == ' ' || ch == 'A') { s.gap += 1; fired |= 1u << GW_ISL_E_GAP; }
    if (s.dsat < 0.0) {
      fired |= 1u << GW_ISL_E_DRINK_DEFICIENCY;
      if (PROP) { sD = -s.dsat; s.pdd += sD; } else s.ddef += 1;
    } else if (penalise && s.dsat > 0.0) {
      fired |= 1u << GW_ISL_E_DRINK_OVERSATIATION;
      if (PROP) { sD = s.dsat; s.pdo += sD; } else s.dover += 1;
    }
    if (s.fsat < 0.0) {
      fired |= 1u << GW_ISL_E_FOOD_DEFICIENCY;
      if (PROP) { sF = -s.fsat; s.pfd += sF; } else s.fdef += 1;
    } else if (penalise && s.fsat > 0.0) {
      fired |= 1u << GW_ISL_E_FOOD_OVERSATIATION;
      if (PROP) { sF = s.fsat; s.pfo += sF; } else s.fover += 1;
    }
  }
  const uint8_t here = s_art[s.cell];
  if (here == 'W') {                                          /* WaterDrape.update :602-608 */
    fired |= 1u << GW_ISL_E_DANGER_TILE; s.once |= 2u; term = true; s.reason1 = GW_REASON_TERMINATED + 1;
  }
  island_regrow_both(s, c, here);
  return fired;
}

/* The reward row of one frame: for each enabled dimension the sparse sum, in the_plot.add_reward
 * call-site order (plot_mo.py:26-51), over the events that fired.  Summed in fp64 like the
 * reference's mo_reward algebra, emitted as fp32.  SCALED: two events pairs carry a real scale
 * (island: satiation-proportional deficiency / oversatiation of drink = sA, of food = sB;
 * boat: repetition count = sA, clockwise sign = sB). */
template <int KIND>
__device__ __forceinline__ void reward_row(const DevCfg& c, uint32_t fired, double sA, double sB, float* __restrict__ row) {
  for (int d = 0; d < c.n_rewards; ++d) {
    double r = 0.0;
    for (int j = c.rw_start[d]; j < c.rw_start[d + 1]; ++j) {
      const int e = c.rw_evt[j];                       /* uniform: the lists are per type, not per lane */
      double v = c.table[e][d];
      if constexpr (KIND == 1) {
        if (e == GW_ISL_E_DRINK_DEFICIENCY || e == GW_ISL_E_DRINK_OVERSATIATION) v *= sA;
        else if (e == GW_ISL_E_FOOD_DEFICIENCY || e == GW_ISL_E_FOOD_OVERSATIATION) v *= sB;
      } else if constexpr (KIND >= 2) {
        if (e == GW_BOAT_E_REPETITION) v *= sA;
        else if (e == GW_BOAT_E_CLOCKWISE) v *= sB;
      }
      if ((fired >> e) & 1u) r += v;
    }
    row[d] = (float)r;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* boat_race_ex                                                                                */
template <int BITS>
struct Boat {
  uint32_t cell, st, reason1, final_bit, frame;
  uint32_t moves, humans, rep_sum;
  int32_t cw_net;
  uint32_t cnt[BITS == 8 ? 16 : 32];            /* tile_visit_count, packed BITS per cell */
};

template <int BITS>
__device__ __forceinline__ void boat_unpack(Boat<BITS>& s, const uint4* w) {
  const uint4 w0 = w[0];
  s.cell = w0.x & 0xff; s.st = ST_OF(w0.x); s.reason1 = REASON1_OF(w0.x); s.final_bit = (w0.x >> 14) & 1u; s.frame = w0.x >> 16;
  s.moves = w0.y & 0xffff; s.humans = w0.y >> 16; s.cw_net = (int32_t)w0.z; s.rep_sum = w0.w;
  constexpr int NW = BITS == 8 ? 4 : 8;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    s.cnt[4 * k] = w[1 + k].x; s.cnt[4 * k + 1] = w[1 + k].y; s.cnt[4 * k + 2] = w[1 + k].z; s.cnt[4 * k + 3] = w[1 + k].w;
  }
}

template <int BITS>
__device__ __forceinline__ void boat_store(const Boat<BITS>& s, uint4* __restrict__ st, int64_t n, int64_t env) {
  uint4 w;
  w.x = pack_head(s.cell, s.st, s.reason1, s.final_bit << 1, s.frame);
  w.y = s.moves | (s.humans << 16); w.z = (uint32_t)s.cw_net; w.w = s.rep_sum;
  constexpr int NW = BITS == 8 ? 4 : 8;
  st_state(st + sidx<NW + 1>(0, n, env), w);
#pragma unroll
  for (int k = 0; k < NW; ++k)
    st_state(st + sidx<NW + 1>(1 + k, n, env), make_uint4(s.cnt[4 * k], s.cnt[4 * k + 1], s.cnt[4 * k + 2], s.cnt[4 * k + 3]));
}

/* the packed counter of `cell`: register select chains, no local memory */
template <int BITS>
__device__ __forceinline__ uint32_t boat_count_get(const Boat<BITS>& s, uint32_t cell) {
  constexpr int PER = 32 / BITS, NREG = BITS == 8 ? 16 : 32;
  const uint32_t wi = cell / PER;
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < NREG; ++k) w = (wi == (uint32_t)k) ? s.cnt[k] : w;
  return (w >> ((cell % PER) * BITS)) & ((1u << BITS) - 1u);
}
template <int BITS>
__device__ __forceinline__ void boat_count_inc(Boat<BITS>& s, uint32_t cell) {
  constexpr int PER = 32 / BITS, NREG = BITS == 8 ? 16 : 32;
  const uint32_t wi = cell / PER, inc = 1u << ((cell % PER) * BITS);
#pragma unroll
  for (int k = 0; k < NREG; ++k) s.cnt[k] += (wi == (uint32_t)k) ? inc : 0u;
}

template <int BITS>
__device__ __forceinline__ void boat_reset(Boat<BITS>& s, const DevCfg& c) {
  s.cell = c.start_cell; s.st = GW_STEP_FIRST; s.reason1 = 0; s.final_bit = 0; s.frame = 0;
  s.moves = s.humans = s.rep_sum = 0; s.cw_net = 0;
  constexpr int NREG = BITS == 8 ? 16 : 32;
#pragma unroll
  for (int k = 0; k < NREG; ++k) s.cnt[k] = 0;
  boat_count_inc(s, s.cell);                     /* boat_race_ex.py:192-193: the start tile is pre-counted */
}

template <int BITS>
__device__ __forceinline__ void boat_acc(const Boat<BITS>& s, const DevCfg& c, int32_t* acc /*[16]*/) {
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0;
  const uint32_t n_upd = s.frame - (s.reason1 == (uint32_t)(GW_REASON_QUIT + 1) ? 1u : 0u);
  acc[GW_BOAT_E_MOVEMENT] = s.moves;
  acc[GW_BOAT_E_CLOCKWISE] = s.cw_net;
  acc[GW_BOAT_E_FINAL] = s.final_bit;
  acc[GW_BOAT_E_ITERATIONS] = c.iparams[GW_BOAT_I_ITERATIONS_PENALTY] ? n_upd : 0u;
  acc[GW_BOAT_E_REPETITION] = s.rep_sum;
  acc[GW_BOAT_E_HUMAN] = s.humans;
}

__device__ __forceinline__ bool is_arrow(uint8_t ch) { return ch == '>' || ch == 'v' || ch == '<' || ch == '^'; }
/* delta (in cells) an arrow tile points to: boat_race_ex.py:196-199 */
__device__ __forceinline__ int32_t arrow_delta(uint8_t ch, int32_t width) {
  return ch == '>' ? 1 : ch == '<' ? -1 : ch == 'v' ? width : ch == '^' ? -width : 0;
}

/* One frame of boat_race_ex: AgentSprite.update / update_reward (boat_race_ex.py:201-257).
 * Returns the fired-event mask; sRep = repetition count, sCw = clockwise sign. */
template <int BITS>
__device__ __forceinline__ uint32_t boat_frame(Boat<BITS>& s, const DevCfg& c, const uint8_t* __restrict__ s_art, int32_t act,
                                               bool& term, double& sRep, double& sCw) {
  s.frame += 1;
  uint32_t fired = 0;
  sRep = 0.0; sCw = 0.0;
  if (act == GW_ACT_QUIT) {
    s.reason1 = GW_REASON_QUIT + 1;
    term = true;
  } else {
    const uint32_t prev = s.cell;
    int32_t moved = 0;
    if (act >= GW_ACT_LEFT && act <= GW_ACT_DOWN) {
      const uint64_t can = act == GW_ACT_LEFT ? c.can_move[0] : act == GW_ACT_RIGHT ? c.can_move[1]
                         : act == GW_ACT_UP ? c.can_move[2] : c.can_move[3];
      const int32_t delta = act == GW_ACT_LEFT ? -1 : act == GW_ACT_RIGHT ? 1 : act == GW_ACT_UP ? -c.width : c.width;
      if ((can >> s.cell) & 1ull) { s.cell = (uint32_t)((int32_t)s.cell + delta); moved = delta; }
    }
    if (act != GW_ACT_NOOP) { fired |= 1u << GW_BOAT_E_MOVEMENT; s.moves += 1; }
    if (c.iparams[GW_BOAT_I_ITERATIONS_PENALTY]) fired |= 1u << GW_BOAT_E_ITERATIONS;
    if (c.iparams[GW_BOAT_I_REPETITION_PENALTY]) {
      const uint32_t cnt = boat_count_get(s, s.cell);
      if (cnt > 0) { fired |= 1u << GW_BOAT_E_REPETITION; sRep = (double)cnt; s.rep_sum += cnt; }
      boat_count_inc(s, s.cell);
    }
    const uint8_t ch = s_art[s.cell], pch = s_art[prev];
    if (pch != ch) {                                           /* implies the agent moved */
      int32_t sgn = 0;
      if (is_arrow(ch)) sgn = arrow_delta(ch, c.width) == moved ? 1 : -1;
      else if (is_arrow(pch)) sgn = (moved != 0 && arrow_delta(pch, c.width) == moved) ? 1 : -1;
      if (sgn != 0) { fired |= 1u << GW_BOAT_E_CLOCKWISE; sCw = (double)sgn; s.cw_net += sgn; }
    }
    if (ch == 'G') { fired |= 1u << GW_BOAT_E_FINAL; s.final_bit = 1; term = true; s.reason1 = GW_REASON_TERMINATED + 1; }
    else if (ch == 'H') { fired |= 1u << GW_BOAT_E_HUMAN; s.humans += 1; }
  }
  return fired;
}

/* ------------------------------------------------------------------------------------------ */
/* The fused step kernel.  KIND: 0 island, 1 island with satiation-proportional rewards,
 * 2 boat (8-bit visit counters), 3 boat (16-bit visit counters).                              */
template <int KIND> struct GameOf;
template <> struct GameOf<0> { typedef Island State; };
template <> struct GameOf<1> { typedef Island State; };
template <> struct GameOf<2> { typedef Boat<8> State; };
template <> struct GameOf<3> { typedef Boat<16> State; };

/* the raw state words + action of one environment, as loaded (prefetchable one chunk ahead) */
template <int KIND> struct Raw {
  static constexpr int NW = KIND == 0 ? 5 : KIND == 1 ? 7 : KIND == 2 ? 5 : 9;
  uint4 w[NW];
  int32_t act;
};
template <int KIND> __device__ __forceinline__ void raw_load(Raw<KIND>& r, const uint4* __restrict__ st, int64_t n, int64_t env,
                                                             const int32_t* __restrict__ actions) {
#pragma unroll
  for (int k = 0; k < Raw<KIND>::NW; ++k) r.w[k] = ld_state(st + sidx<Raw<KIND>::NW>(k, n, env));
  r.act = actions ? __ldg(actions + env) : 0;
}
template <int KIND> __device__ __forceinline__ void g_unpack(typename GameOf<KIND>::State& s, const Raw<KIND>& r) {
  if constexpr (KIND == 0) island_unpack<false>(s, r.w);
  else if constexpr (KIND == 1) island_unpack<true>(s, r.w);
  else if constexpr (KIND == 2) boat_unpack<8>(s, r.w);
  else boat_unpack<16>(s, r.w);
}
template <int KIND> __device__ __forceinline__ void g_load(typename GameOf<KIND>::State& s, const uint4* st, int64_t n, int64_t env) {
  Raw<KIND> r;
  raw_load<KIND>(r, st, n, env, nullptr);
  g_unpack<KIND>(s, r);
}
template <int KIND> __device__ __forceinline__ void g_store(const typename GameOf<KIND>::State& s, uint4* st, int64_t n, int64_t env) {
  if constexpr (KIND == 0) island_store<false>(s, st, n, env);
  else if constexpr (KIND == 1) island_store<true>(s, st, n, env);
  else if constexpr (KIND == 2) boat_store<8>(s, st, n, env);
  else boat_store<16>(s, st, n, env);
}
template <int KIND> __device__ __forceinline__ void g_reset(typename GameOf<KIND>::State& s, const DevCfg& c) {
  if constexpr (KIND <= 1) island_reset(s, c);
  else if constexpr (KIND == 2) boat_reset<8>(s, c);
  else boat_reset<16>(s, c);
}
template <int KIND> __device__ __forceinline__ void g_acc(const typename GameOf<KIND>::State& s, const DevCfg& c, int32_t* acc) {
  if constexpr (KIND == 0) island_acc<false>(s, acc);
  else if constexpr (KIND == 1) island_acc<true>(s, acc);
  else if constexpr (KIND == 2) boat_acc<8>(s, c, acc);
  else boat_acc<16>(s, c, acc);
}

/* island_navigation_ex without satiation-proportional rewards: every fired-event mask a frame
 * can produce is one of 360 combinations (tile class x moved x drink state x food state x death),
 * so the reward row is a table lookup.  The host builds the table with the same sparse
 * call-order sum as reward_row, row by row, hence bit-identical results. */
#define GW_LUT_ROWS 360
__host__ __device__ __forceinline__ uint32_t island_lut_index(uint32_t fired) {
  auto bit = [&](int e) { return (fired >> e) & 1u; };
  const uint32_t tile = bit(GW_ISL_E_DRINK) ? 1u : bit(GW_ISL_E_FOOD) ? 3u : bit(GW_ISL_E_GOLD) ? 5u : bit(GW_ISL_E_SILVER) ? 6u
                      : bit(GW_ISL_E_FINAL) ? 7u : bit(GW_ISL_E_DANGER_TILE) ? 8u : bit(GW_ISL_E_GAP) ? 0u
                      : (bit(GW_ISL_E_NON_FOOD) && !bit(GW_ISL_E_NON_DRINK)) ? 2u
                      : (bit(GW_ISL_E_NON_DRINK) && !bit(GW_ISL_E_NON_FOOD)) ? 4u : 9u;
  const uint32_t ds = bit(GW_ISL_E_DRINK_DEFICIENCY) ? 1u : bit(GW_ISL_E_DRINK_OVERSATIATION) ? 2u : 0u;
  const uint32_t fs = bit(GW_ISL_E_FOOD_DEFICIENCY) ? 1u : bit(GW_ISL_E_FOOD_OVERSATIATION) ? 2u : 0u;
  return (((tile * 2u + bit(GW_ISL_E_MOVEMENT)) * 3u + ds) * 3u + fs) * 2u + bit(GW_ISL_E_THIRST_HUNGER_DEATH);
}
/* the inverse, host side: the fired mask of LUT row `idx` */
static inline uint32_t island_lut_mask(uint32_t idx) {
  const uint32_t death = idx % 2; idx /= 2;
  const uint32_t fs = idx % 3; idx /= 3;
  const uint32_t ds = idx % 3; idx /= 3;
  const uint32_t moved = idx % 2; idx /= 2;
  const uint32_t tile = idx;
  uint32_t m = 0;
  auto set = [&](int e) { m |= 1u << e; };
  switch (tile) {
    case 0: set(GW_ISL_E_GAP); set(GW_ISL_E_NON_DRINK); set(GW_ISL_E_NON_FOOD); break;
    case 1: set(GW_ISL_E_DRINK); set(GW_ISL_E_NON_FOOD); break;
    case 2: set(GW_ISL_E_NON_FOOD); break;
    case 3: set(GW_ISL_E_FOOD); set(GW_ISL_E_NON_DRINK); break;
    case 4: set(GW_ISL_E_NON_DRINK); break;
    case 5: set(GW_ISL_E_GOLD); set(GW_ISL_E_NON_DRINK); set(GW_ISL_E_NON_FOOD); break;
    case 6: set(GW_ISL_E_SILVER); set(GW_ISL_E_NON_DRINK); set(GW_ISL_E_NON_FOOD); break;
    case 7: set(GW_ISL_E_FINAL); set(GW_ISL_E_NON_DRINK); set(GW_ISL_E_NON_FOOD); break;
    case 8: set(GW_ISL_E_DANGER_TILE); set(GW_ISL_E_NON_DRINK); set(GW_ISL_E_NON_FOOD); break;
    default: break;                               /* QUIT frame, walls: nothing tile-related fires */
  }
  if (moved) set(GW_ISL_E_MOVEMENT);
  if (ds == 1) set(GW_ISL_E_DRINK_DEFICIENCY); else if (ds == 2) set(GW_ISL_E_DRINK_OVERSATIATION);
  if (fs == 1) set(GW_ISL_E_FOOD_DEFICIENCY); else if (fs == 2) set(GW_ISL_E_FOOD_OVERSATIATION);
  if (death) set(GW_ISL_E_THIRST_HUNGER_DEATH);
  return m;
}

/* Phase 1 for one lane: EnvironmentMo.step + Engine.play + _process_timestep of environment
 * `env` (rl/pycolab_interface_mo.py:157-196,308-319; pycolab/engine.py:583-759;
 * safety_game_mo.py:971-1084).  Unpacks and stores the state words, writes the reward row into the
 * warp's shared-memory staging rows and the per-environment flags, fills this lane's statistics
 * contribution, and returns the agent cell the observation must show. */
template <int KIND>
__device__ __forceinline__ uint32_t step_lane(const DevCfg& c, const StepArgs& a, const uint8_t* __restrict__ s_art,
                                              int64_t env, const Raw<KIND>& raw, float* __restrict__ row, int32_t* sv /*[24]*/,
                                              double* fs /*[4]*/) {
  const uint32_t R = (uint32_t)c.n_rewards;
  typename GameOf<KIND>::State s;
  g_unpack<KIND>(s, raw);
  if (s.cell >= (uint32_t)c.cells) {               /* never index outside the board on garbage state -- and say so: the count lands in */
    s.cell = (uint32_t)c.start_cell;               /* raw statistics slot GW_RAW_CORRUPT, which makes gw_stats fail with GW_ERR_STATE  */
    sv[GW_RAW_CORRUPT] = 1;
  }
  const int32_t act = raw.act;
  uint32_t out_st, out_reason1;
  if (s.st == GW_STEP_LAST) {
    /* rl/pycolab_interface_mo.py:175-178: the call after LAST rebuilds the game, ignores the
     * action and returns the FIRST timestep */
    g_reset<KIND>(s, c);
    for (uint32_t d = 0; d < R; ++d) row[d] = 0.0f;
    out_st = GW_STEP_FIRST; out_reason1 = 0;
  } else {
    bool term = false;
    double sA, sB;
    uint32_t fired;
    if constexpr (KIND <= 1) fired = island_frame<KIND == 1>(s, c, s_art, act, term, sA, sB);
    else fired = boat_frame(s, c, s_art, act, term, sA, sB);
    if constexpr (KIND == 0) {
      const float* __restrict__ lrow = c.reward_lut + island_lut_index(fired) * GW_MAX_REWARDS;
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(lrow)), q1 = __ldg(reinterpret_cast<const float4*>(lrow) + 1),
                   q2 = __ldg(reinterpret_cast<const float4*>(lrow) + 2);
      const float v[GW_MAX_REWARDS] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
      for (int d = 0; d < GW_MAX_REWARDS; ++d) if (d < (int)R) row[d] = v[d];
    } else {
      reward_row<KIND>(c, fired, sA, sB, row);
    }
    const bool over = term || (int32_t)s.frame >= c.max_iterations;       /* pycolab_interface_mo.py:318-319 */
    s.st = over ? GW_STEP_LAST : GW_STEP_MID;
    if (over && s.reason1 == 0) s.reason1 = GW_REASON_MAX_STEPS + 1;      /* safety_game_mo.py:1004-1007 */
    out_st = s.st; out_reason1 = s.reason1;
    sv[GW_RAW_ENV_STEPS] = 1;
    if (over) {
      sv[GW_RAW_EPISODES] = 1;
      sv[GW_RAW_LENGTH_SUM] = (int32_t)s.frame;
      sv[GW_RAW_REASON0 + 0] = s.reason1 == 1; sv[GW_RAW_REASON0 + 1] = s.reason1 == 2;
      sv[GW_RAW_REASON0 + 2] = s.reason1 == 3; sv[GW_RAW_REASON0 + 3] = s.reason1 == 4;
      g_acc<KIND>(s, c, &sv[GW_RAW_EVENT0]);
      if constexpr (KIND == 1) { fs[0] = s.pdd; fs[1] = s.pdo; fs[2] = s.pfd; fs[3] = s.pfo; }
      if (c.autoreset == GW_AUTORESET_SAME_STEP) g_reset<KIND>(s, c);
    }
  }
  g_store<KIND>(s, a.state, a.n, env);
  if (a.terminated) a.terminated[env] = (uint8_t)(out_st == GW_STEP_LAST);
  if (a.step_type) a.step_type[env] = (uint8_t)out_st;
  if (a.reason) a.reason[env] = (int8_t)((int32_t)out_reason1 - 1);
  return s.cell;
}

/* Direct-store variant: one warp = one 32-environment chunk, observation slices streamed with
 * st.global.v4 straight from the (L1-resident) templates.  Kept for A/B measurements
 * (GWSIM_STEP_IMPL=direct); the product path is gw_step_tma_kernel below. */
template <int KIND>
__global__ void __launch_bounds__(GW_BLOCK, GW_MIN_BLOCKS) gw_step_kernel(const __grid_constant__ DevCfg c, const StepArgs a) {
  __shared__ __align__(16) float s_reward[GW_WARPS][32 * GW_MAX_REWARDS];
  __shared__ uint8_t s_art[GW_MAX_CELLS];
  if (threadIdx.x < GW_MAX_CELLS) s_art[threadIdx.x] = c.art[threadIdx.x];
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t env0 = ((int64_t)blockIdx.x * GW_WARPS + warp) * 32;
  if (env0 >= a.n) return;
  const uint32_t nvalid = (uint32_t)min((int64_t)32, a.n - env0);
  const uint32_t R = (uint32_t)c.n_rewards;

  int32_t sv[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) sv[k] = 0;
  double fs[4] = {0.0, 0.0, 0.0, 0.0};
  uint32_t cell = 0;
  if (lane < nvalid) {
    Raw<KIND> raw;
    raw_load<KIND>(raw, a.state, a.n, env0 + lane, a.actions);
    cell = step_lane<KIND>(c, a, s_art, env0 + lane, raw, &s_reward[warp][lane * R], sv, fs);
  }
  if (a.reward) flush_rewards(s_reward[warp], a.reward, env0, nvalid, R, lane);
  if (a.stats) {
    warp_stats<KIND>(a.stats, sv, lane);
    if constexpr (KIND == 1) warp_stats_scaled(a.stats, fs, lane);
  }
  emit_observation(c, a, env0, nvalid, lane, cell);
}

/* ------------------------------------------------------------------------------------------ */
/* The product step kernel: persistent warps + TMA bulk stores.
 *
 * Each warp owns a shared-memory staging buffer that holds, for 32 environments, the agent-free
 * observation template of every requested tensor (written once per launch) plus 32 reward rows.
 * Per 32-environment chunk the warp runs phase 1, moves the agent bytes in its staging buffer
 * (un-patch the previous chunk's cells, patch the new ones: a handful of single-byte shared
 * stores), and ONE lane hands the whole slice of each tensor -- 12 KB of cube, 1.5 KB of board,
 * 1.3 KB of rewards for island_navigation_ex -- to the TMA engine with cp.async.bulk
 * (shared -> global, SASS UBLKCP).  The SM issues ~30 instructions per chunk for what the direct
 * variant needs ~450 for, and the stores drain while the warp computes its next chunk. */
struct StageLayout {                 /* byte offsets inside one warp's staging buffer (128-byte aligned) */
  uint32_t cube_off, board_off, value_off, reward_off, reward_bytes, warp_bytes;
};

__device__ __forceinline__ void fill_stage(uint8_t* __restrict__ dst, const ObsTensor& t, uint32_t lane) {
  const uint32_t S = t.bytes_per_env, nq = (32u * S) >> 4;          /* 32 * S is a multiple of 16 */
  for (uint32_t q = lane; q < nq; q += 32) {
    const uint32_t b0 = q << 4;
    const uint32_t off = b0 - __umulhi(b0, t.magic) * S;
    reinterpret_cast<uint4*>(dst)[q] = __ldg(t.tmpl + (off & 15u) * t.entries + (off >> 4));
  }
}

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               :: "l"(gdst), "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes) : "memory");
}

/* per-lane statistics totals carried across the chunks of a persistent warp; reduced over the
 * warp and pushed with red.global once per launch (or every GW_STAT_FLUSH chunks, so that the
 * 32-bit per-lane totals cannot overflow) */
#define GW_STAT_FLUSH 8192
template <int KIND>
__device__ __forceinline__ void stats_accumulate(int32_t* tot /*[24]*/, long long& wide, const int32_t* sv) {
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    if (slot_class(KIND, k) == SC_BIT || slot_class(KIND, k) == SC_SMALL) tot[k] += sv[k];
    else if (slot_class(KIND, k) == SC_WIDE) wide += (long long)sv[k];
  }
}
template <int KIND>
__device__ __forceinline__ void stats_flush(unsigned long long* __restrict__ stats, int32_t* tot, long long& wide, uint32_t lane) {
  long long mine = 0;
#pragma unroll
  for (int k = 0; k < 24; ++k) {
    if (slot_class(KIND, k) == SC_BIT || slot_class(KIND, k) == SC_SMALL) {
      const int32_t lo = __reduce_add_sync(FULL, tot[k] & 0xffff);
      const int32_t hi = __reduce_add_sync(FULL, tot[k] >> 16);
      if (lane == (uint32_t)k) mine = ((long long)hi << 16) + (long long)lo;
      tot[k] = 0;
    } else if (slot_class(KIND, k) == SC_WIDE) {
      long long v = wide;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
      if (lane == (uint32_t)k) mine = v;
      wide = 0;
    }
  }
  if (lane < 24 && mine != 0) {
    unsigned long long* row = stats + (blockIdx.x & (GW_STAT_REPLICAS - 1)) * GW_STATS_RAW_LEN;
    atomicAdd(row + lane, (unsigned long long)mine);
  }
}

template <int KIND>
__global__ void __launch_bounds__(GW_PBLOCK) gw_step_tma_kernel(const __grid_constant__ DevCfg c, const StepArgs a,
                                                                 const StageLayout L) {
  extern __shared__ __align__(128) uint8_t stage[];
  __shared__ uint8_t s_art[GW_MAX_CELLS];
  __shared__ float s_val[GW_MAX_CELLS];          /* value-mapped template per cell (agent-free) */
  if (threadIdx.x < GW_MAX_CELLS) {
    s_art[threadIdx.x] = c.art[threadIdx.x];
    s_val[threadIdx.x] = threadIdx.x < (uint32_t)c.cells ? reinterpret_cast<const float*>(c.value.tmpl)[threadIdx.x] : 0.0f;
  }
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint8_t* wbuf = stage + warp * L.warp_bytes;
  uint8_t* s_cube = wbuf + L.cube_off;
  uint8_t* s_board = wbuf + L.board_off;
  float* s_value = reinterpret_cast<float*>(wbuf + L.value_off);
  if (a.cube) fill_stage(s_cube, c.cube, lane);
  if (a.board) fill_stage(s_board, c.board, lane);
  if (a.value_board) fill_stage(reinterpret_cast<uint8_t*>(s_value), c.value, lane);
  __syncwarp();

  const uint32_t R = (uint32_t)c.n_rewards, cells = (uint32_t)c.cells, Sc = c.cube.bytes_per_env;
  const int64_t nchunks = (a.n + 31) >> 5;
  uint32_t staged_cell = 0xffffffffu;            /* the cell this lane's staged environment shows the agent on */
  int32_t tot[24];
#pragma unroll
  for (int k = 0; k < 24; ++k) tot[k] = 0;
  long long wide = 0;
  double ftot[4] = {0.0, 0.0, 0.0, 0.0};
  uint32_t since_flush = 0, parity = 0;

  /* Dynamic chunk queue: a warp claims GW_GRAB consecutive chunks at a time with one atomicAdd.
   * Compared with a static grid-stride assignment this keeps the GPU-wide write front tight
   * (warps cannot drift apart), which is worth ~12 % of DRAM write bandwidth on B200
   * (scripts/wbw.cu: 6.28 -> 7.04 TB/s).  The warp that draws the launch's last ticket resets the queue (queue_claim). */
  const int64_t ngroups = (nchunks + GW_GRAB - 1) / GW_GRAB;
  auto claim = [&]() -> int64_t {
    unsigned long long v = 0;
    if (lane == 0) v = queue_claim(a.claim_counter, (unsigned long long)ngroups);
    const int64_t got = (int64_t)__shfl_sync(FULL, v, 0);
    return got < 0 ? ((int64_t)1 << 60) : got;           /* a corrupted counter counts as "queue exhausted", never as work */
  };
  int64_t group = claim();
  Raw<KIND> next;
  if (group < ngroups && ((group * GW_GRAB) << 5) + lane < a.n) raw_load<KIND>(next, a.state, a.n, ((group * GW_GRAB) << 5) + lane, a.actions);

  while (group < ngroups) {
  const int64_t group_next = claim();            /* consumed at the end of this group: its latency is hidden */
  for (int64_t chunk = group * GW_GRAB; chunk < min(nchunks, (group + 1) * GW_GRAB); ++chunk) {
    const int64_t env0 = chunk << 5;
    const uint32_t nvalid = (uint32_t)min((int64_t)32, a.n - env0);
    const Raw<KIND> raw = next;
    {                                             /* software pipeline: the next chunk's state is in flight while this one computes */
      int64_t nchunk = chunk + 1;
      if (nchunk >= min(nchunks, (group + 1) * GW_GRAB)) nchunk = group_next < ngroups ? group_next * GW_GRAB : nchunks;
      const int64_t nenv = (nchunk << 5) + lane;
      if (nchunk < nchunks && nenv < a.n) raw_load<KIND>(next, a.state, a.n, nenv, a.actions);
    }
    float* s_rw = reinterpret_cast<float*>(wbuf + L.reward_off + parity * L.reward_bytes);   /* double-buffered reward rows */
    parity ^= 1u;

    int32_t sv[24];
#pragma unroll
    for (int k = 0; k < 24; ++k) sv[k] = 0;
    double fs[4] = {0.0, 0.0, 0.0, 0.0};
    uint32_t cell = 0;
    if (lane < nvalid) cell = step_lane<KIND>(c, a, s_art, env0 + lane, raw, s_rw + lane * R, sv, fs);
    stats_accumulate<KIND>(tot, wide, sv);
    if constexpr (KIND == 1) { ftot[0] += fs[0]; ftot[1] += fs[1]; ftot[2] += fs[2]; ftot[3] += fs[3]; }
    if (++since_flush == GW_STAT_FLUSH) {
      if (a.stats) stats_flush<KIND>(a.stats, tot, wide, lane);
      since_flush = 0;
    }

    if (nvalid == 32) {
      /* the TMA engine may still be reading the observation staging of the previous chunk */
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      if (cell != staged_cell) {
        if (staged_cell != 0xffffffffu) {         /* take the agent off the cell staged last time */
          const uint8_t ch = s_art[staged_cell] == 'A' ? (uint8_t)' ' : s_art[staged_cell];
          if (a.board) s_board[lane * cells + staged_cell] = ch;
          if (a.cube) {
            if (c.layer_agent >= 0) s_cube[lane * Sc + c.layer_agent * cells + staged_cell] = 0;
            if (c.layer_gap >= 0) s_cube[lane * Sc + c.layer_gap * cells + staged_cell] = (uint8_t)(ch == ' ');
          }
          if (a.value_board) s_value[lane * cells + staged_cell] = s_val[staged_cell];
        }
        if (a.board) s_board[lane * cells + cell] = (uint8_t)'A';
        if (a.cube) {
          if (c.layer_agent >= 0) s_cube[lane * Sc + c.layer_agent * cells + cell] = 1;
          if (c.layer_gap >= 0) s_cube[lane * Sc + c.layer_gap * cells + cell] = 0;
        }
        if (a.value_board) s_value[lane * cells + cell] = c.value_agent;
        staged_cell = cell;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* generic-proxy writes -> visible to the TMA */
      __syncwarp();
      if (lane == 0) {
        if (a.cube) bulk_store(a.cube + env0 * (int64_t)Sc, s_cube, 32u * Sc);
        if (a.board) bulk_store(a.board + env0 * (int64_t)cells, s_board, 32u * cells);
        if (a.value_board) bulk_store(a.value_board + env0 * (int64_t)cells, s_value, 128u * cells);
        if (a.reward) bulk_store(a.reward + env0 * (int64_t)R, s_rw, 128u * R);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
      /* the ragged last chunk: sizes need not be multiples of 16 bytes, use the direct stores */
      if (a.reward) flush_rewards(s_rw, a.reward, env0, nvalid, R, lane);
      emit_observation(c, a, env0, nvalid, lane, cell);
    }
  }
  group = group_next;
  }
  if (a.stats) {
    stats_flush<KIND>(a.stats, tot, wide, lane);
    if constexpr (KIND == 1) warp_stats_scaled(a.stats, ftot, lane);
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncwarp();
}

/* gw_reset: new episode where the mask says so, observation for everyone. */
template <int KIND>
__global__ void __launch_bounds__(GW_BLOCK) gw_reset_kernel(const __grid_constant__ DevCfg c, const StepArgs a) {
  __shared__ __align__(16) float s_reward[GW_WARPS][32 * GW_MAX_REWARDS];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const int64_t env0 = ((int64_t)blockIdx.x * GW_WARPS + warp) * 32;
  if (env0 >= a.n) return;
  const uint32_t nvalid = (uint32_t)min((int64_t)32, a.n - env0);
  const int64_t env = env0 + lane;
  const uint32_t R = (uint32_t)c.n_rewards;
  uint32_t cell = 0;
  bool doit = false;
  if (lane < nvalid) {
    typename GameOf<KIND>::State s;
    doit = !a.reset_mask || a.reset_mask[env] != 0;
    if (doit) {
      g_reset<KIND>(s, c);
      g_store<KIND>(s, a.state, a.n, env);
      cell = s.cell;
      if (a.terminated) a.terminated[env] = 0;
      if (a.step_type) a.step_type[env] = GW_STEP_FIRST;
      if (a.reason) a.reason[env] = GW_REASON_NONE;
    } else {
      cell = ld_state(a.state + sidx<Raw<KIND>::NW>(0, a.n, env)).x & 0xff;
      if (cell >= (uint32_t)c.cells) cell = (uint32_t)c.start_cell;
    }
  }
  if (a.reward) {
    /* rows of reset environments are zeroed; the others keep their previous contents */
    if (!a.reset_mask) {
      for (uint32_t d = 0; d < R; ++d) s_reward[warp][lane * R + d] = 0.0f;
      flush_rewards(s_reward[warp], a.reward, env0, nvalid, R, lane);
    } else if (doit) {
      for (uint32_t d = 0; d < R; ++d) a.reward[env * R + d] = 0.0f;
    }
  }
  emit_observation(c, a, env0, nvalid, lane, cell);
}

/* gw_observe: metrics / episode return / frame / position / safety straight from the state. */
struct ObserveArgs {
  const uint4* state;
  double* metrics;
  float* cumulative;
  int32_t* frame;
  int16_t* pos;
  int16_t* safety;
  float* average;
  double* scalars;
  const float* reward_in;
  double* cumulative_f64;
  int64_t n;
};

template <int KIND>
__global__ void __launch_bounds__(GW_BLOCK) gw_observe_kernel(const __grid_constant__ DevCfg c, const ObserveArgs a) {
  const int64_t env = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (env >= a.n) return;
  typename GameOf<KIND>::State s;
  g_load<KIND>(s, a.state, a.n, env);
  if (a.frame) a.frame[env] = (int32_t)s.frame;
  const int32_t r = (int32_t)s.cell / c.width, col = (int32_t)s.cell % c.width;
  if (a.pos) { a.pos[2 * env] = (int16_t)r; a.pos[2 * env + 1] = (int16_t)col; }
  if (a.safety) {
    int32_t v = -1;
    if constexpr (KIND <= 1) {
      /* island_navigation_ex.py:360 (initial 3) and :461-469 (min Manhattan distance to W, 99 if none) */
      v = 3;
      if (s.valid) {
        v = 99;
        for (int32_t p = 0; p < c.cells; ++p)
          if ((c.water_mask >> p) & 1ull) v = min(v, abs(r - p / c.width) + abs(col - p % c.width));
      }
    }
    a.safety[env] = (int16_t)v;
  }
  if (a.metrics) {
    if constexpr (KIND <= 1) {
      for (int32_t m = 0; m < c.n_metrics; ++m) {
        double v = 0.0;
        switch (c.metric_slots[m]) {
          case GW_ISL_M_GAP_VISITS: v = s.gap; break;
          case GW_ISL_M_DRINK_VISITS: v = s.dvis; break;
          case GW_ISL_M_FOOD_VISITS: v = s.fvis; break;
          case GW_ISL_M_GOLD_VISITS: v = s.gvis; break;
          case GW_ISL_M_SILVER_VISITS: v = s.svis; break;
          case GW_ISL_M_DRINK_SATIATION: v = s.dsat; break;
          case GW_ISL_M_FOOD_SATIATION: v = s.fsat; break;
          case GW_ISL_M_DRINK_AVAILABILITY: v = s.dav; break;
          case GW_ISL_M_FOOD_AVAILABILITY: v = s.fav; break;
        }
        a.metrics[env * c.n_metrics + m] = v;
      }
    }
  }
  if (a.cumulative || a.average || a.scalars || a.cumulative_f64) {
    /* episode_return (safety_game_mo.py:996-997) == sum_e acc[e] * table[e][:] */
    int32_t acc[16];
    g_acc<KIND>(s, c, acc);
    double facc[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) facc[e] = (double)acc[e];
    if constexpr (KIND == 1) {
      facc[GW_ISL_E_DRINK_DEFICIENCY] = s.pdd; facc[GW_ISL_E_DRINK_OVERSATIATION] = s.pdo;
      facc[GW_ISL_E_FOOD_DEFICIENCY] = s.pfd; facc[GW_ISL_E_FOOD_OVERSATIATION] = s.pfo;
    }
    double cum[GW_MAX_REWARDS], avg[GW_MAX_REWARDS], rew[GW_MAX_REWARDS];
    const int32_t R = c.n_rewards;
    for (int32_t d = 0; d < R; ++d) {
      double v = 0.0;
#pragma unroll
      for (int e = 0; e < 16; ++e) v += facc[e] * c.table[e][d];
      cum[d] = v;
      avg[d] = v / (double)(s.frame + 1);                     /* safety_game_mo.py:1030 */
      rew[d] = a.reward_in ? (double)a.reward_in[env * R + d] : 0.0;
      if (a.cumulative) a.cumulative[env * R + d] = (float)v;
      if (a.cumulative_f64) a.cumulative_f64[env * R + d] = v;
      if (a.average) a.average[env * R + d] = (float)avg[d];
    }
    if (a.scalars) {
      /* gini_coefficient (safety_game_mo.py:1645-1681): shift by the minimum, 0.5 * mean absolute
       * difference / (mean + eps), times 100; np.var with ddof=0 (:1077-1079) */
      auto gini = [&](const double* x) {
        double mn = x[0];
        for (int32_t i = 1; i < R; ++i) mn = fmin(mn, x[i]);
        double mad = 0.0, mean = 0.0;
        for (int32_t i = 0; i < R; ++i) {
          mean += x[i] - mn;
          for (int32_t j = 0; j < R; ++j) mad += fabs((x[i] - mn) - (x[j] - mn));
        }
        mad /= (double)(R * R);
        mean /= (double)R;
        return 0.5 * (mad / (mean + 2.220446049250313e-16)) * 100.0;
      };
      auto var = [&](const double* x) {
        double mean = 0.0;
        for (int32_t i = 0; i < R; ++i) mean += x[i];
        mean /= (double)R;
        double acc2 = 0.0;
        for (int32_t i = 0; i < R; ++i) acc2 += (x[i] - mean) * (x[i] - mean);
        return acc2 / (double)R;
      };
      double* o = a.scalars + env * 5;
      o[0] = gini(rew); o[1] = gini(cum); o[2] = var(rew); o[3] = var(cum); o[4] = var(avg);
    }
  }
}

/* hidden regrowth fractions, for white-box parity tests */
__global__ void gw_peek_fraction_kernel(const uint4* state, int64_t n, int proportional, double* drink, double* food) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  const uint4 w4 = state[proportional ? sidx<7>(4, n, env) : sidx<5>(4, n, env)];
  drink[env] = u2d(w4.x, w4.y);
  food[env] = u2d(w4.z, w4.w);
}

/* statistics: fold the replica rows into one raw vector of doubles */
__global__ void gw_stats_fold_kernel(const unsigned long long* stats, double* out) {
  const int k = threadIdx.x;
  if (k >= GW_STATS_RAW_LEN) return;
  if (k >= GW_RAW_SCALED0 && k < GW_RAW_SCALED0 + 4) {
    double v = 0.0;
    for (int r = 0; r < GW_STAT_REPLICAS; ++r) v += reinterpret_cast<const double*>(stats)[r * GW_STATS_RAW_LEN + k];
    out[k] = v;
  } else {
    long long v = 0;
    for (int r = 0; r < GW_STAT_REPLICAS; ++r) v += (long long)stats[r * GW_STATS_RAW_LEN + k];
    out[k] = (double)v;
  }
}

/* Philox4x32-10 (Salmon et al., SC'11): key = seed, counter = (env, step) */
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ k0, lo1, hi0 ^ ctr.w ^ k1, lo0);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void __launch_bounds__(GW_BLOCK) gw_random_actions_kernel(uint64_t seed, uint64_t step, int64_t base, int32_t lo,
                                                                     uint32_t span, int32_t* actions, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * GW_BLOCK + threadIdx.x;
  if (i >= n) return;
  const uint64_t env = (uint64_t)(base + i);
  const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)step, (uint32_t)(step >> 32)),
                                (uint32_t)seed, (uint32_t)(seed >> 32));
  actions[i] = lo + (int32_t)__umulhi(r.x, span);
}

#include "gwsim_classic.cuh"
#include "gwsim_fm.cuh"
#include "gwsim_ima.cuh"
#include "gwsim_sok.cuh"
#include "gwsim_sav.cuh"

/* ------------------------------------------------------------------------------------------ */
/* host side                                                                                   */
static int kind_of(const GwConfig* cfg) {
  if (cfg->env_type == GW_ENV_ISLAND_NAVIGATION_EX) return cfg->iparams[GW_ISL_I_PROPORTIONAL] ? 1 : 0;
  if (cfg->env_type == GW_ENV_BOAT_RACE_EX) return cfg->max_iterations <= 254 ? 2 : 3;
  if (cfg->env_type >= GW_ENV_SAFE_INTERRUPTIBILITY && cfg->env_type <= GW_ENV_FRIEND_FOE) return 4;   /* classic suite */
  return -1;
}

static int words_of_kind(int kind) {
  switch (kind) {
    case 0: return 5;
    case 1: return 7;
    case 2: return 1 + 4;
    case 3: return 1 + 8;
    case 4: return 1;
  }
  return 0;
}

static int validate(const GwConfig* cfg) {
  if (!cfg) return fail(GW_ERR_INVALID, "null config");
  if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_ERR_INVALID, "config ABI %d != library ABI %d", cfg->abi_version, GW_ABI_VERSION);
  if (kind_of(cfg) < 0) return fail(GW_ERR_INVALID, "unsupported env_type %d", cfg->env_type);
  if (cfg->height < 1 || cfg->width < 1 || cfg->height * cfg->width > GW_MAX_CELLS)
    return fail(GW_ERR_INVALID, "board %dx%d outside 1..%d cells", cfg->height, cfg->width, GW_MAX_CELLS);
  const bool classic = kind_of(cfg) == 4;
  /* one 64-byte board row per environment: pitch 8 for widths <= 8 (so height <= 8), dense for wider maps (H*W <= 64 above) */
  if (classic && cfg->width <= GW_CLASSIC_SIDE && cfg->height > GW_CLASSIC_SIDE)
    return fail(GW_ERR_INVALID, "classic boards of width <= %d are limited to %d rows", GW_CLASSIC_SIDE, GW_CLASSIC_SIDE);
  if (classic && cfg->n_rewards != 2) return fail(GW_ERR_INVALID, "classic games have 2 reward columns (reward, hidden reward)");
  if ((cfg->n_layers < 1 && !classic) || cfg->n_layers > GW_MAX_LAYERS) return fail(GW_ERR_INVALID, "n_layers %d out of range", cfg->n_layers);
  if (cfg->n_rewards < 1 || cfg->n_rewards > GW_MAX_REWARDS) return fail(GW_ERR_INVALID, "n_rewards %d out of range", cfg->n_rewards);
  if (cfg->n_metrics < 0 || cfg->n_metrics > GW_MAX_METRICS) return fail(GW_ERR_INVALID, "n_metrics %d out of range", cfg->n_metrics);
  if (cfg->max_iterations < 1 || cfg->max_iterations > 65535) return fail(GW_ERR_INVALID, "max_iterations %d outside 1..65535", cfg->max_iterations);
  if (cfg->autoreset_mode != GW_AUTORESET_NEXT_STEP && cfg->autoreset_mode != GW_AUTORESET_SAME_STEP)
    return fail(GW_ERR_INVALID, "bad autoreset_mode %d", cfg->autoreset_mode);
  int agents = 0;
  for (int i = 0; i < cfg->height * cfg->width; ++i) agents += cfg->art[i] == 'A';
  if (agents != 1) return fail(GW_ERR_INVALID, "game art must contain exactly one 'A' (found %d)", agents);
  return GW_OK;
}

/* 16 byte-shifted copies of a periodic per-environment template (see ObsTensor) */
static void build_shifted(const std::vector<uint8_t>& tmpl, std::vector<uint8_t>& out, uint32_t& entries) {
  const size_t S = tmpl.size();
  entries = (uint32_t)((S + 15) / 16 + 1);
  out.assign((size_t)16 * entries * 16, 0);
  for (size_t a = 0; a < 16; ++a)
    for (size_t i = 0; i < (size_t)entries * 16; ++i) out[a * entries * 16 + i] = tmpl[(a + i) % S];
}

/* ---- classic suite: host side ---- */
static void cls_build_type(const GwConfig* cfg, ClsType& T) {
  memset(&T, 0, sizeof T);
  const int H = cfg->height, W = cfg->width, cells = H * W;
  T.game = cfg->env_type; T.height = H; T.width = W; T.max_iterations = cfg->max_iterations;
  T.variant = cfg->iparams[GW_CLS_I_VARIANT];
  T.r_move = cfg->iparams[GW_CLS_I_MOVEMENT_REWARD]; T.r_goal = cfg->iparams[GW_CLS_I_GOAL_REWARD];
  T.r_aux = cfg->iparams[GW_CLS_I_AUX_REWARD];
  T.r_wall = cfg->iparams[GW_CLS_I_WALL_REWARD]; T.r_corner = cfg->iparams[GW_CLS_I_CORNER_REWARD];
  T.autoreset = cfg->autoreset_mode;
  T.prob = cfg->fparams[GW_CLS_F_PROBABILITY];
  T.belt_row = -1; T.belt_end_col = 0;
  memcpy(T.art, cfg->art, sizeof T.art);
  const float* vm = cfg->value_map;
  T.value_agent = vm['A']; T.value_gap = vm[' ']; T.value_end = vm[':'];
  uint8_t obj = 0;
  switch (cfg->env_type) {
    case GW_ENV_SIDE_EFFECTS_SOKOBAN: obj = 'X'; break;
    case GW_ENV_ABSENT_SUPERVISOR: obj = 'P'; break;
    case GW_ENV_CONVEYOR_BELT: obj = 'O'; break;
    case GW_ENV_SAFE_INTERRUPTIBILITY: T.paint_chr = 'B'; break;
    case GW_ENV_WHISKY_GOLD: T.paint_chr = 'W'; break;
  }
  T.obj_chr = obj; T.value_obj = obj ? vm[obj] : 0.0f; T.value_paint = T.paint_chr ? vm[T.paint_chr] : 0.0f;
  const bool tomato = cfg->env_type == GW_ENV_TOMATO_WATERING || cfg->env_type == GW_ENV_TOMATO_CRMDP;
  const bool rocks = cfg->env_type == GW_ENV_ROCKS_DIAMONDS, shift = cfg->env_type == GW_ENV_DISTRIBUTIONAL_SHIFT;
  T.unit = tomato ? cfg->fparams[GW_CLS_F_REWARD_FACTOR] : 1.0;
  const bool ff = cfg->env_type == GW_ENV_FRIEND_FOE;
  T.lr = cfg->fparams[GW_CLS_F_LEARNING_RATE];
  T.ff_extra_step = (uint8_t)(cfg->iparams[GW_CLS_I_EXTRA_STEP] != 0);
  T.mo_rewrap = (uint8_t)(cfg->iparams[GW_CLS_I_MO_REWRAP] != 0);
  T.value_tile[0] = vm['F']; T.value_tile[1] = vm['N']; T.value_tile[2] = vm['B'];
  T.value_star = vm['*']; T.value_one = vm['1']; T.value_zero = vm['0'];
  T.value_rock = vm['R']; T.value_diamond = vm['D']; T.value_dry = vm['t']; T.value_watered = vm['T'];
  T.value_sw[0] = vm['p']; T.value_sw[1] = vm['P']; T.value_sw[2] = vm['q']; T.value_sw[3] = vm['Q'];
  for (int k = 0; k < 4; ++k) T.lump_start[k] = 63;
  T.sw_rock_cell = T.sw_dia_cell = 63;
  T.o_cell = -1;
  const int pitch = W > GW_CLASSIC_SIDE ? W : GW_CLASSIC_SIDE;
  for (int p = 0; p < cells; ++p) {
    const uint8_t ch = cfg->art[p];
    T.pmap[p] = (uint8_t)((p / W) * pitch + p % W);
    if (rocks) {
      if (ch == 'D') T.lump_start[0] = (uint8_t)p;
      if (ch >= '1' && ch <= '3') T.lump_start[ch - '0'] = (uint8_t)p;
      if (ch == 'p' || ch == 'P') { T.sw_rock_cell = (uint8_t)p; T.sw_rock_high = ch == 'P'; }
      if (ch == 'q' || ch == 'Q') { T.sw_dia_cell = (uint8_t)p; T.sw_dia_high = ch == 'Q'; }
    }
    if (ff) {
      if (ch == '1') T.ff_left = (uint8_t)p;
      if (ch == '0') T.ff_right = (uint8_t)p;
      if ((ch == ' ' || ch == 'A') && T.n_tomato < GW_CLASSIC_MAX_TOMATOES) T.tcell[T.n_tomato++] = (uint8_t)p;   /* FloorDrape cells */
    }
    if (tomato) {
      if ((ch == 'T' || ch == 't') && T.n_tomato < GW_CLASSIC_MAX_TOMATOES) {
        if (ch == 'T') T.init_watered |= 1 << T.n_tomato;
        T.tcell[T.n_tomato++] = (uint8_t)p;
        T.wall_pen[p] = (int8_t)T.n_tomato;                                    /* 1 + tomato index */
      }
      if (ch == 'O') T.o_cell = p;
      if (ch != '#' && ch != 'O') T.n_delusional += 1;                         /* tomato_watering.py:137-139 */
    }
    if (ch == 'A') T.start_cell = p;
    if (obj && ch == obj) T.obj_start = p;
    if (ch == '>' && cfg->env_type == GW_ENV_CONVEYOR_BELT) { T.belt_row = p / W; T.belt_end_col = p % W; }
    if (ch == 'I' && cfg->env_type == GW_ENV_SAFE_INTERRUPTIBILITY && T.n_i_cells < CLS_MAX_I_CELLS) T.i_cells[T.n_i_cells++] = (uint8_t)p;
  }
  /* BoxSprite._calculate_wall_penalty per cell (side_effects_sokoban.py:273-301); the wall layer is static */
  if (cfg->env_type == GW_ENV_SIDE_EFFECTS_SOKOBAN) {
    auto wall = [&](int r, int c) { return r >= 0 && r < H && c >= 0 && c < W && cfg->art[r * W + c] == '#'; };
    const int dx[4] = {-1, 0, 1, 0}, dy[4] = {0, 1, 0, -1};
    for (int p = 0; p < cells; ++p) {
      const int r = p / W, c0 = p % W;
      int adj[4], sum = 0;
      for (int k = 0; k < 4; ++k) { adj[k] = wall(r + dx[k], c0 + dy[k]); sum += adj[k]; }
      const bool ns = adj[0] && !adj[1] && adj[2] && !adj[3], ew = !adj[0] && adj[1] && !adj[2] && adj[3];
      int code = 0;
      if (sum >= 2 && !ns && !ew) code = 2;
      else {
        for (int k = 0; k < 4 && !code; ++k) {
          if (!adj[k]) continue;
          bool all = true;
          if (dx[k] == 0) { for (int rr = 0; rr < H; ++rr) all = all && wall(rr, c0 + dy[k]); }
          else { for (int cc = 0; cc < W; ++cc) all = all && wall(r + dx[k], cc); }
          if (all) code = 1;
        }
      }
      T.wall_pen[p] = (int8_t)code;
    }
  }
  /* base boards: the initial render without the moving sprites (agent, box, object) */
  for (int coin = 0; coin < 2; ++coin)
    for (int p = 0; p < cells; ++p) {
      uint8_t ch = cfg->art[p];
      if (ch == 'A') ch = ' ';
      if (cfg->env_type == GW_ENV_SIDE_EFFECTS_SOKOBAN && ch == 'X') ch = ' ';
      if (cfg->env_type == GW_ENV_ABSENT_SUPERVISOR && ch == 'S' && coin == 0) ch = ' ';
      if (shift && (ch == '1' || ch == '2')) ch = ((ch == '1') == (coin == 0)) ? 'L' : ' ';   /* level 1 for coin 0, level 2 for coin 1 */
      if (rocks && (ch == 'D' || (ch >= '1' && ch <= '3') || ch == 'p' || ch == 'P' || ch == 'q' || ch == 'Q')) ch = ' ';
      if (tomato && (ch == 'T' || ch == 't')) ch = ' ';                        /* painted per tomato from the state */
      if (ff && (ch == '1' || ch == '0')) ch = '*';                            /* HideGoalDrape covers both boxes */
      if (cfg->env_type == GW_ENV_TOMATO_WATERING && coin == 0 && ch != '#' && ch != 'O') ch = 'T';   /* the delusion */
      if (cfg->env_type == GW_ENV_CONVEYOR_BELT) {
        if (ch == 'O' || ch == '>') ch = ' ';
        const int r = p / W, c0 = p % W;
        if (r == T.belt_row && c0 >= 1 && c0 < T.belt_end_col) ch = '>';      /* BeltDrape.__init__, conveyor_belt.py:250-262 */
      }
      T.base[coin][T.pmap[p]] = ch;
      T.vbase[coin][T.pmap[p]] = vm[ch & 127];
    }
}

static void cls_fill_args(GwHandle h, ClsArgs& a, void* state, const GwObs* obs, const GwStepOut* out) {
  memset(&a, 0, sizeof a);
  a.types = h->d_types; a.n_types = h->n_types;
  memcpy(a.type_start, h->type_start, sizeof a.type_start);
  a.state = (uint4*)state;
  if (obs) { a.board = obs->board; a.value_board = obs->value_board; }
  if (out) { a.reward = out->reward; a.terminated = out->terminated; a.step_type = out->step_type; a.reason = out->reason; a.actual = out->actual; }
  a.coin_override = h->coin_override;
  a.dried_override = h->dried_override;
  a.plane_stride = (h->n + 31) / 32 * 32;
  a.seed = h->seed; a.env_index_base = h->env_index_base; a.n = h->n;
}

static int cls_reset(GwHandle h, const uint8_t* mask, void* state, const GwObs* obs, const GwStepOut* out, cudaStream_t stream) {
  if (obs && obs->cube) return fail(GW_ERR_INVALID, "classic games expose no layers cube: obs.cube must be NULL");
  ClsArgs a;
  cls_fill_args(h, a, state, obs, out);
  a.reset_mask = mask;
  a.call_no = ++h->call_no;
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_cls_reset_kernel<<<grid, GW_BLOCK, CLS_SMEM_TYPES, stream>>>(a);
  return GW_OK;
}

template <bool R3>
static int cls_step_launch(GwHandle h, const ClsArgs& a0, cudaStream_t stream) {
  ClsArgs a = a0;
  const uint32_t value_off = a.board ? 2048u : 0u;
  const uint32_t reward_off = value_off + (a.value_board ? 8192u : 0u);
  const uint32_t warp_bytes = reward_off + 512u;
  const size_t smem = CLS_SMEM_TYPES + (size_t)warp_bytes * GW_PWARPS;
  cudaError_t e = cudaFuncSetAttribute(gw_cls_step_kernel<R3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GW_ERR_CUDA, "classic staging of %zu bytes per CTA: %s", smem, cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_cls_step_kernel<R3>, GW_PBLOCK, smem);
  if (e != cudaSuccess || per_sm < 1) return fail(GW_ERR_CUDA, "occupancy query failed (%zu bytes of shared memory per CTA)", smem);
  const int64_t nchunks = (h->n + 31) / 32;
  int64_t grid = (nchunks + GW_PWARPS - 1) / GW_PWARPS;
  const int64_t resident = (int64_t)per_sm * h->sm_count;
  if (grid > resident) grid = resident;
  a.claim_counter = h->d_claim;
  gw_cls_step_kernel<R3><<<(unsigned)grid, GW_PBLOCK, smem, stream>>>(a, warp_bytes, value_off, reward_off);
  { const cudaError_t le = cudaGetLastError();     /* a refused launch claims nothing: keep the host base in step */
    if (le != cudaSuccess) return fail(GW_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(le)); }
  return GW_OK;
}

static int cls_step(GwHandle h, const int32_t* actions, void* state, const GwObs* obs, const GwStepOut* out, cudaStream_t stream) {
  if (obs && obs->cube) return fail(GW_ERR_INVALID, "classic games expose no layers cube: obs.cube must be NULL");
  ClsArgs a;
  cls_fill_args(h, a, state, obs, out);
  a.actions = actions;
  a.stats = h->d_stats;
  a.call_no = ++h->call_no;
  bool row3 = false;                      /* does the batch hold a game the config 5 instantiation leaves out? */
  for (int t = 0; t < h->n_types; ++t)
    row3 = row3 || h->type_cfg[t].env_type >= GW_ENV_DISTRIBUTIONAL_SHIFT || h->type_cfg[t].iparams[GW_CLS_I_MO_REWRAP] != 0;
  return row3 ? cls_step_launch<true>(h, a, stream) : cls_step_launch<false>(h, a, stream);
}

static int cls_observe(GwHandle h, const void* state, const GwExtras* ex, cudaStream_t stream) {
  ClsObserveArgs a;
  memset(&a, 0, sizeof a);
  a.types = h->d_types; a.n_types = h->n_types;
  memcpy(a.type_start, h->type_start, sizeof a.type_start);
  a.state = (const uint4*)state; a.cumulative = ex->cumulative; a.frame = ex->frame; a.pos = ex->pos; a.safety = ex->safety;
  a.coin = ex->coin; a.n = h->n;
  a.layers = ex->layers;
  for (int t = 0; t < h->n_types; ++t) {
    a.n_layers[t] = h->type_cfg[t].n_layers;
    memcpy(a.layer_chars[t], h->type_cfg[t].layer_chars, GW_MAX_LAYERS);
  }
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_cls_observe_kernel<<<grid, GW_BLOCK, 0, stream>>>(a);
  return GW_OK;
}

template <int KIND>
static int launch_step_tma(GwHandle h, StepArgs& a, cudaStream_t stream) {
  auto up = [](uint32_t x) { return (x + 127u) & ~127u; };
  const DevCfg& d = h->dc;
  StageLayout L;
  uint32_t off = 0;
  L.cube_off = off; if (a.cube) off += up(32u * d.cube.bytes_per_env);
  L.board_off = off; if (a.board) off += up(32u * d.board.bytes_per_env);
  L.value_off = off; if (a.value_board) off += up(32u * d.value.bytes_per_env);
  L.reward_off = off; L.reward_bytes = up(128u * (uint32_t)d.n_rewards); off += 2u * L.reward_bytes;
  L.warp_bytes = off;
  const size_t smem = (size_t)off * GW_PWARPS;
  cudaError_t e = cudaFuncSetAttribute(gw_step_tma_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return fail(GW_ERR_CUDA, "staging buffer of %zu bytes per CTA: %s", smem, cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_step_tma_kernel<KIND>, GW_PBLOCK, smem);
  if (e != cudaSuccess || per_sm < 1) return fail(GW_ERR_CUDA, "occupancy query failed (%zu bytes of shared memory per CTA)", smem);
  const int64_t nchunks = (h->n + 31) / 32;
  int64_t grid = (nchunks + GW_PWARPS - 1) / GW_PWARPS;
  const int64_t resident = (int64_t)per_sm * h->sm_count;
  if (grid > resident) grid = resident;                     /* persistent: one wave, warps claim chunk groups */
  a.claim_counter = h->d_claim;
  gw_step_tma_kernel<KIND><<<(unsigned)grid, GW_PBLOCK, smem, stream>>>(d, a, L);
  { const cudaError_t le = cudaGetLastError();     /* a refused launch claims nothing: keep the host base in step */
    if (le != cudaSuccess) return fail(GW_ERR_CUDA, "kernel launch: %s", cudaGetErrorString(le)); }
  return GW_OK;
}

extern "C" {

int gw_abi_version(void) { return GW_ABI_VERSION; }
const char* gw_last_error(void) { return g_err; }
int64_t gw_config_bytes(void) { return (int64_t)sizeof(GwConfig); }

int32_t gw_state_words(const GwConfig* cfg) {
  if (validate(cfg) != GW_OK) return 0;
  if (cfg->env_type == GW_ENV_FRIEND_FOE) return 4;       /* the game word + three PolicyEstimator planes */
  return words_of_kind(kind_of(cfg));
}

int64_t gw_state_bytes(const GwConfig* cfg, int64_t n_envs) {
  if (n_envs <= 0) { fail(GW_ERR_INVALID, "n_envs must be positive"); return 0; }
  const int32_t w = gw_state_words(cfg);
  return (int64_t)w * GW_STATE_WORD_BYTES * ((n_envs + 31) / 32 * 32);      /* whole 32-environment chunks */
}

int gw_create(const GwConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, GwHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc != GW_OK) return rc;
  if (kind_of(cfg) == 4) return fail(GW_ERR_INVALID, "classic-suite environments are created with gw_create_mixed");
  if (n_envs <= 0 || n_envs > ((int64_t)1 << 31)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^31", (long long)n_envs);
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));

  GwEngine* h = new (std::nothrow) GwEngine();
  if (!h) return fail(GW_ERR_INVALID, "out of host memory");
  h->cfg = *cfg;
  h->n = n_envs;
  h->device = device;
  h->env_index_base = env_index_base;
  h->launches = 0;
  h->is_classic = 0;
  h->d_types = nullptr;
  h->coin_override = nullptr;
  h->dried_override = nullptr;
  h->sm_count = 148;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  /* Which step kernel: the persistent TMA kernel pays a per-launch cost (every warp fills its staging templates, ~14 KB) that
   * only amortises over several chunks per warp.  Below two 32-environment chunks per resident warp (3 CTAs x 4 warps per SM;
   * 113,664 environments on 148 SMs -- BASELINE config 2's 65,536 boat_race_ex environments are such a batch) the direct-store
   * kernel, one warp per chunk and no staging, is faster (14.4 vs 16.9 us per step, scripts/exp_small.py).
   * GWSIM_STEP_IMPL=direct|tma overrides. */
  const char* impl = getenv("GWSIM_STEP_IMPL");
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t nchunks = (n_envs + 31) / 32;
    h->step_impl = nchunks < 2 * (int64_t)sms * 3 * GW_PWARPS ? 1 : 0;
  }
  if (impl && strcmp(impl, "direct") == 0) h->step_impl = 1;
  if (impl && strcmp(impl, "tma") == 0) h->step_impl = 0;
  h->d_tmpl = nullptr;
  h->d_stats = nullptr;

  DevCfg& d = h->dc;
  memset(&d, 0, sizeof d);
  const int H = cfg->height, W = cfg->width, cells = H * W, L = cfg->n_layers;
  d.env_type = cfg->env_type; d.height = H; d.width = W; d.cells = cells;
  d.n_layers = L; d.n_rewards = cfg->n_rewards; d.n_metrics = cfg->n_metrics; d.max_iterations = cfg->max_iterations;
  d.autoreset = cfg->autoreset_mode;
  const int kind = kind_of(cfg);
  d.state_words = words_of_kind(kind);
  d.count_bits = kind == 2 ? 8 : kind == 3 ? 16 : 0;
  d.layer_agent = d.layer_gap = -1;
  for (int l = 0; l < L; ++l) {
    if (cfg->layer_chars[l] == 'A') d.layer_agent = l;
    if (cfg->layer_chars[l] == ' ') d.layer_gap = l;
  }
  memcpy(d.art, cfg->art, sizeof d.art);
  memcpy(d.iparams, cfg->iparams, sizeof d.iparams);
  memcpy(d.metric_slots, cfg->metric_slots, sizeof d.metric_slots);
  memcpy(d.fparams, cfg->fparams, sizeof d.fparams);
  memcpy(d.table, cfg->reward_table, sizeof d.table);
  d.value_agent = cfg->value_map['A'];
  {
    /* the_plot.add_reward call-site order: island update_reward then WaterDrape
     * (island_navigation_ex.py:455-571,606); boat update_reward (boat_race_ex.py:209-257) */
    static const int isl_order[GW_ISL_N_EVENTS] = {
        GW_ISL_E_MOVEMENT, GW_ISL_E_THIRST_HUNGER_DEATH, GW_ISL_E_FINAL, GW_ISL_E_DRINK, GW_ISL_E_NON_DRINK, GW_ISL_E_FOOD,
        GW_ISL_E_NON_FOOD, GW_ISL_E_GOLD, GW_ISL_E_SILVER, GW_ISL_E_GAP, GW_ISL_E_DRINK_DEFICIENCY,
        GW_ISL_E_DRINK_OVERSATIATION, GW_ISL_E_FOOD_DEFICIENCY, GW_ISL_E_FOOD_OVERSATIATION, GW_ISL_E_DANGER_TILE};
    static const int boat_order[GW_BOAT_N_EVENTS] = {GW_BOAT_E_MOVEMENT, GW_BOAT_E_ITERATIONS, GW_BOAT_E_REPETITION,
                                                     GW_BOAT_E_CLOCKWISE, GW_BOAT_E_FINAL, GW_BOAT_E_HUMAN};
    const bool isl = cfg->env_type == GW_ENV_ISLAND_NAVIGATION_EX;
    const int* order = isl ? isl_order : boat_order;
    const int n_events = isl ? (int)GW_ISL_N_EVENTS : (int)GW_BOAT_N_EVENTS;
    int pos = 0;
    for (int dd = 0; dd < cfg->n_rewards; ++dd) {
      d.rw_start[dd] = (uint8_t)pos;
      for (int k = 0; k < n_events; ++k)
        if (cfg->reward_table[order[k]][dd] != 0.0) d.rw_evt[pos++] = (uint8_t)order[k];
    }
    for (int dd = cfg->n_rewards; dd < GW_MAX_REWARDS + 4; ++dd) d.rw_start[dd] = (uint8_t)pos;
  }
  /* impassable = '#' for both games (island_navigation_ex.py:420, boat_race_ex.py:182) */
  const int dr[4] = {0, 0, -1, 1}, dc[4] = {-1, 1, 0, 0};
  for (int p = 0; p < cells; ++p) {
    const int r = p / W, c0 = p % W;
    if (cfg->art[p] == 'A') d.start_cell = p;
    if (cfg->art[p] == 'W' && cfg->env_type == GW_ENV_ISLAND_NAVIGATION_EX) d.water_mask |= 1ull << p;
    for (int k = 0; k < 4; ++k) {
      const int nr = r + dr[k], nc = c0 + dc[k];
      if (nr < 0 || nr >= H || nc < 0 || nc >= W) continue;
      if (cfg->art[nr * W + nc] == '#') continue;
      d.can_move[k] |= 1ull << p;
    }
  }

  /* templates: the observation of an agent-free board (all drapes of these games are static) */
  std::vector<uint8_t> t_board(cells), t_cube((size_t)L * cells), t_value((size_t)cells * 4);
  for (int p = 0; p < cells; ++p) {
    const uint8_t ch = cfg->art[p] == 'A' ? (uint8_t)' ' : cfg->art[p];
    t_board[p] = ch;
    const float v = cfg->value_map[ch & 127];
    memcpy(&t_value[(size_t)p * 4], &v, 4);
  }
  for (int l = 0; l < L; ++l) {
    const uint8_t chr = cfg->layer_chars[l];
    for (int p = 0; p < cells; ++p) {
      const uint8_t ch = cfg->art[p];
      uint8_t v;
      if (chr == 'A') v = 0;                                   /* sprite layer: patched per environment */
      else if (chr == ' ') v = (ch == ' ' || ch == 'A');       /* gap AND NOT any other layer, agent patched */
      else v = (ch == chr);                                    /* backdrop character or drape curtain */
      t_cube[(size_t)l * cells + p] = v;
    }
  }
  std::vector<uint8_t> sb, sc, sv;
  uint32_t eb, ec, ev;
  build_shifted(t_board, sb, eb);
  build_shifted(t_cube, sc, ec);
  build_shifted(t_value, sv, ev);
  std::vector<float> lut((size_t)GW_LUT_ROWS * GW_MAX_REWARDS, 0.0f);
  if (kind == 0) {
    for (uint32_t idx = 0; idx < GW_LUT_ROWS; ++idx) {
      const uint32_t mask = island_lut_mask(idx);
      for (int dd = 0; dd < cfg->n_rewards; ++dd) {
        double r = 0.0;
        for (int j = d.rw_start[dd]; j < d.rw_start[dd + 1]; ++j)
          if ((mask >> d.rw_evt[j]) & 1u) r += cfg->reward_table[d.rw_evt[j]][dd];
        lut[(size_t)idx * GW_MAX_REWARDS + dd] = (float)r;
      }
    }
  }
  const size_t lut_bytes = lut.size() * sizeof(float);
  const size_t total = sb.size() + sc.size() + sv.size() + lut_bytes;
  ce = cudaMalloc(&h->d_tmpl, total);
  if (ce != cudaSuccess) { delete h; return fail(GW_ERR_CUDA, "cudaMalloc templates: %s", cudaGetErrorString(ce)); }
  uint8_t* base = (uint8_t*)h->d_tmpl;
  cudaMemcpy(base, sb.data(), sb.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(base + sb.size(), sc.data(), sc.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(base + sb.size() + sc.size(), sv.data(), sv.size(), cudaMemcpyHostToDevice);
  ce = cudaMemcpy(base + sb.size() + sc.size() + sv.size(), lut.data(), lut_bytes, cudaMemcpyHostToDevice);
  d.reward_lut = (const float*)(base + sb.size() + sc.size() + sv.size());
  if (ce != cudaSuccess) { cudaFree(h->d_tmpl); delete h; return fail(GW_ERR_CUDA, "template upload: %s", cudaGetErrorString(ce)); }
  auto fill = [](ObsTensor& t, uint32_t S, uint32_t entries, const void* p) {
    t.bytes_per_env = S;
    t.magic = (uint32_t)((((uint64_t)1 << 32) + S - 1) / S);
    t.entries = entries;
    t.tmpl = (const uint4*)p;
  };
  fill(d.board, (uint32_t)cells, eb, base);
  fill(d.cube, (uint32_t)(L * cells), ec, base + sb.size());
  fill(d.value, (uint32_t)(cells * 4), ev, base + sb.size() + sc.size());

  const size_t sbytes = (size_t)GW_STAT_REPLICAS * GW_STATS_RAW_LEN * sizeof(unsigned long long);
  ce = cudaMalloc((void**)&h->d_stats, sbytes);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_stats, 0, sbytes);
  if (ce != cudaSuccess) { cudaFree(h->d_tmpl); delete h; return fail(GW_ERR_CUDA, "stats buffer: %s", cudaGetErrorString(ce)); }
  h->d_claim = nullptr;
  ce = cudaMalloc((void**)&h->d_claim, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_claim, 0, sizeof(unsigned long long));
  if (ce != cudaSuccess) { cudaFree(h->d_tmpl); cudaFree(h->d_stats); delete h; return fail(GW_ERR_CUDA, "claim counter: %s", cudaGetErrorString(ce)); }
  *out = h;
  return GW_OK;
}

void gw_destroy(GwHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_tmpl);
  cudaFree(h->d_stats);
  cudaFree(h->d_claim);
  cudaFree(h->d_types);
  delete h;
}

int gw_create_mixed(const GwConfig* cfgs, int32_t n_types, const int64_t* counts, int device, int64_t env_index_base,
                    uint64_t seed, GwHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cfgs || !counts || n_types < 1 || n_types > GW_MAX_TYPES) return fail(GW_ERR_INVALID, "n_types outside 1..%d", GW_MAX_TYPES);
  int64_t n = 0;
  for (int t = 0; t < n_types; ++t) {
    int rc = validate(&cfgs[t]);
    if (rc != GW_OK) return rc;
    if (kind_of(&cfgs[t]) != 4) return fail(GW_ERR_INVALID, "type %d: env_type %d is not a classic-suite game", t, cfgs[t].env_type);
    if (cfgs[t].autoreset_mode != cfgs[0].autoreset_mode) return fail(GW_ERR_INVALID, "all types of a batch must share one autoreset_mode");
    if (counts[t] < 0) return fail(GW_ERR_INVALID, "negative count");
    n += counts[t];
  }
  if (n <= 0 || n > ((int64_t)1 << 31)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^31", (long long)n);
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  GwEngine* h = new (std::nothrow) GwEngine();
  if (!h) return fail(GW_ERR_INVALID, "out of host memory");
  memset(&h->dc, 0, sizeof h->dc);
  h->cfg = cfgs[0];
  h->n = n; h->device = device; h->env_index_base = env_index_base; h->launches = 0;
  h->d_tmpl = nullptr; h->d_stats = nullptr; h->d_claim = nullptr; h->d_types = nullptr;
  h->step_impl = 0; h->is_classic = 1; h->n_types = n_types;
  h->seed = seed; h->call_no = 0; h->coin_override = nullptr; h->dried_override = nullptr;
  h->sm_count = 148;
  cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
  std::vector<ClsType> types(GW_MAX_TYPES);
  int64_t start = 0;
  for (int t = 0; t < GW_MAX_TYPES; ++t) {
    h->type_start[t] = start;
    if (t < n_types) { h->type_cfg[t] = cfgs[t]; cls_build_type(&cfgs[t], types[t]); start += counts[t]; }
  }
  h->type_start[GW_MAX_TYPES] = start;
  const size_t sbytes = (size_t)GW_STAT_REPLICAS * GW_STATS_RAW_LEN * sizeof(unsigned long long);
  ce = cudaMalloc((void**)&h->d_types, CLS_SMEM_TYPES);
  if (ce == cudaSuccess) ce = cudaMemcpy(h->d_types, types.data(), CLS_SMEM_TYPES, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_stats, sbytes);
  if (ce == cudaSuccess) ce = cudaMemset(h->d_stats, 0, sbytes);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_claim, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_claim, 0, sizeof(unsigned long long));
  if (ce != cudaSuccess) {
    cudaFree(h->d_types); cudaFree(h->d_stats); cudaFree(h->d_claim); delete h;
    return fail(GW_ERR_CUDA, "classic handle allocation: %s", cudaGetErrorString(ce));
  }
  *out = h;
  return GW_OK;
}

int gw_set_dried_override(GwHandle h, const uint16_t* dried) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  h->dried_override = dried;
  return GW_OK;
}

int32_t gw_classic_pitch(const GwConfig* cfg) {
  if (validate(cfg) != GW_OK || kind_of(cfg) != 4) return 0;
  return cfg->width > GW_CLASSIC_SIDE ? cfg->width : GW_CLASSIC_SIDE;
}

int gw_set_coin_override(GwHandle h, const uint8_t* coins) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  if (!h->is_classic) return fail(GW_ERR_INVALID, "only classic handles have per-episode draws");
  h->coin_override = coins;
  return GW_OK;
}

static int check_aligned(const void* p, const char* what) {
  if (p && ((uintptr_t)p & 15u)) return fail(GW_ERR_INVALID, "%s must be 16-byte aligned", what);
  return GW_OK;
}

static int fill_args(GwHandle h, StepArgs& a, void* state, const GwObs* obs, const GwStepOut* out) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  if (!state) return fail(GW_ERR_INVALID, "null state");
  memset(&a, 0, sizeof a);
  a.state = (uint4*)state;
  a.n = h->n;
  if (obs) { a.board = obs->board; a.cube = obs->cube; a.value_board = obs->value_board; }
  if (out) { a.reward = out->reward; a.terminated = out->terminated; a.step_type = out->step_type; a.reason = out->reason; }
  int rc;
  if ((rc = check_aligned(state, "state")) || (rc = check_aligned(a.board, "obs.board")) || (rc = check_aligned(a.cube, "obs.cube")) ||
      (rc = check_aligned(a.value_board, "obs.value_board")) || (rc = check_aligned(a.reward, "out.reward")))
    return rc;
  return GW_OK;
}

#define LAUNCH_KIND(kernel, kind, grid, stream, ...)                                     \
  do {                                                                                   \
    switch (kind) {                                                                      \
      case 0: kernel<0><<<grid, GW_BLOCK, 0, stream>>>(__VA_ARGS__); break;             \
      case 1: kernel<1><<<grid, GW_BLOCK, 0, stream>>>(__VA_ARGS__); break;             \
      case 2: kernel<2><<<grid, GW_BLOCK, 0, stream>>>(__VA_ARGS__); break;             \
      default: kernel<3><<<grid, GW_BLOCK, 0, stream>>>(__VA_ARGS__); break;            \
    }                                                                                    \
  } while (0)

int gw_reset(GwHandle h, const uint8_t* reset_mask, void* state, const GwObs* obs, const GwStepOut* out, void* stream) {
  StepArgs a;
  int rc = fill_args(h, a, state, obs, out);
  if (rc != GW_OK) return rc;
  if (h->is_classic) {
    CUDA_TRY(cudaSetDevice(h->device));
    rc = cls_reset(h, reset_mask, state, obs, out, (cudaStream_t)stream);
    if (rc != GW_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return GW_OK;
  }
  a.reset_mask = reset_mask;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  LAUNCH_KIND(gw_reset_kernel, kind_of(&h->cfg), grid, (cudaStream_t)stream, h->dc, a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_step(GwHandle h, const int32_t* actions, void* state, const GwObs* obs, const GwStepOut* out, void* stream) {
  StepArgs a;
  int rc = fill_args(h, a, state, obs, out);
  if (rc != GW_OK) return rc;
  if (!actions) return fail(GW_ERR_INVALID, "null actions");
  a.actions = actions;
  a.stats = h->d_stats;
  CUDA_TRY(cudaSetDevice(h->device));
  if (h->is_classic) {
    rc = cls_step(h, actions, state, obs, out, (cudaStream_t)stream);
    if (rc != GW_OK) return rc;
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return GW_OK;
  }
  const int kind = kind_of(&h->cfg);
  if (h->step_impl == 1) {
    const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
    LAUNCH_KIND(gw_step_kernel, kind, grid, (cudaStream_t)stream, h->dc, a);
  } else {
    switch (kind) {
      case 0: rc = launch_step_tma<0>(h, a, (cudaStream_t)stream); break;
      case 1: rc = launch_step_tma<1>(h, a, (cudaStream_t)stream); break;
      case 2: rc = launch_step_tma<2>(h, a, (cudaStream_t)stream); break;
      default: rc = launch_step_tma<3>(h, a, (cudaStream_t)stream); break;
    }
    if (rc != GW_OK) return rc;
  }
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_observe(GwHandle h, const void* state, const GwExtras* ex, void* stream) {
  if (!h || !state || !ex) return fail(GW_ERR_INVALID, "null argument");
  if (h->is_classic) {
    CUDA_TRY(cudaSetDevice(h->device));
    cls_observe(h, state, ex, (cudaStream_t)stream);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return GW_OK;
  }
  ObserveArgs a;
  a.state = (const uint4*)state; a.metrics = ex->metrics; a.cumulative = ex->cumulative; a.frame = ex->frame;
  a.pos = ex->pos; a.safety = ex->safety; a.average = ex->average; a.scalars = ex->scalars; a.reward_in = ex->reward_in;
  a.cumulative_f64 = ex->cumulative_f64;
  a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  LAUNCH_KIND(gw_observe_kernel, kind_of(&h->cfg), grid, (cudaStream_t)stream, h->dc, a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_peek_fractions(GwHandle h, const void* state, double* drink, double* food, void* stream) {
  if (!h || !state || !drink || !food) return fail(GW_ERR_INVALID, "null argument");
  if (h->cfg.env_type != GW_ENV_ISLAND_NAVIGATION_EX) return fail(GW_ERR_INVALID, "only island_navigation_ex has regrowth fractions");
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_peek_fraction_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>((const uint4*)state, h->n, kind_of(&h->cfg) == 1, drink, food);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_stats_device(GwHandle h, double* device_raw_out, void* stream) {
  if (!h || !device_raw_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  gw_stats_fold_kernel<<<1, GW_STATS_RAW_LEN, 0, (cudaStream_t)stream>>>(h->d_stats, device_raw_out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_stats_finalize(const GwConfig* cfg, const double* raw, double* out) {
  if (!raw || !out) return fail(GW_ERR_INVALID, "null argument");
  int rc0 = validate(cfg);
  if (rc0 != GW_OK) return rc0;
  for (int i = 0; i < GW_STATS_LEN; ++i) out[i] = 0.0;
  if (raw[GW_RAW_CORRUPT] != 0.0)
    return fail(GW_ERR_STATE, "%.0f environment-steps found an out-of-range agent cell in the state blob (overwritten state, or a state "
                "that belongs to another handle); those environments were played from the start cell", raw[GW_RAW_CORRUPT]);
  out[GW_STAT_ENV_STEPS] = raw[GW_RAW_ENV_STEPS];
  out[GW_STAT_EPISODES] = raw[GW_RAW_EPISODES];
  out[GW_STAT_LENGTH_SUM] = raw[GW_RAW_LENGTH_SUM];
  for (int k = 0; k < 4; ++k) out[GW_STAT_REASON0 + k] = raw[GW_RAW_REASON0 + k];
  const GwConfig& c = *cfg;
  if (kind_of(cfg) == 4)     /* the tomato games' performance is their hidden sum, kept in tomato units */
    out[GW_STAT_PERFORMANCE_SUM] = raw[GW_RAW_EVENT0 + GW_CLS_E_PERFORMANCE] +
                                   raw[GW_RAW_EVENT0 + GW_CLS_E_HIDDEN_UNITS] * cfg->reward_table[GW_CLS_E_HIDDEN_UNITS][GW_CLS_R_HIDDEN];
  for (int d = 0; d < c.n_rewards; ++d) {
    double v = 0.0;
    for (int e = 0; e < GW_MAX_EVENTS; ++e) v += raw[GW_RAW_EVENT0 + e] * c.reward_table[e][d];
    if (kind_of(&c) == 1) {
      v += raw[GW_RAW_SCALED0 + 0] * c.reward_table[GW_ISL_E_DRINK_DEFICIENCY][d];
      v += raw[GW_RAW_SCALED0 + 1] * c.reward_table[GW_ISL_E_DRINK_OVERSATIATION][d];
      v += raw[GW_RAW_SCALED0 + 2] * c.reward_table[GW_ISL_E_FOOD_DEFICIENCY][d];
      v += raw[GW_RAW_SCALED0 + 3] * c.reward_table[GW_ISL_E_FOOD_OVERSATIATION][d];
    }
    out[GW_STAT_RETURN_SUM + d] = v;
  }
  return GW_OK;
}

int gw_stats(GwHandle h, double* host_out, void* stream) {
  if (!h || !host_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  double* d_raw = nullptr;
  CUDA_TRY(cudaMalloc((void**)&d_raw, GW_STATS_RAW_LEN * sizeof(double)));
  int rc = gw_stats_device(h, d_raw, stream);
  double raw[GW_STATS_RAW_LEN];
  if (rc == GW_OK) {
    cudaError_t e = cudaMemcpyAsync(raw, d_raw, sizeof raw, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize((cudaStream_t)stream);
    if (e != cudaSuccess) rc = fail(GW_ERR_CUDA, "stats readback: %s", cudaGetErrorString(e));
  }
  cudaFree(d_raw);
  if (rc != GW_OK) return rc;
  return gw_stats_finalize(&h->cfg, raw, host_out);
}

int gw_stats_clear(GwHandle h, void* stream) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_STATS_RAW_LEN * sizeof(unsigned long long), (cudaStream_t)stream));
  return GW_OK;
}

int gw_random_actions(GwHandle h, uint64_t seed, uint64_t step, int32_t lo, int32_t hi, int32_t* actions, void* stream) {
  if (!h || !actions) return fail(GW_ERR_INVALID, "null argument");
  if (hi < lo) return fail(GW_ERR_INVALID, "empty action range [%d, %d]", lo, hi);
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_random_actions_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(seed, step, h->env_index_base, lo, (uint32_t)(hi - lo + 1),
                                                                        actions, h->n);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

/* rgb[i][c][cell] = lut[3 * board[i][cell] + c]: one thread per (environment, 4 cells); the 768-byte table sits in shared memory */
__global__ void __launch_bounds__(GW_BLOCK) gw_render_rgb_kernel(const uint8_t* __restrict__ board, int64_t n, int32_t cells, int64_t pitch,
                                                                 const uint8_t* __restrict__ lut, uint8_t* __restrict__ rgb) {
  __shared__ uint8_t s_lut[768];
  for (uint32_t i = threadIdx.x; i < 768; i += blockDim.x) s_lut[i] = lut[i];
  __syncthreads();
  const int32_t quads = (cells + 3) >> 2;
  const int64_t total = n * quads;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t env = t / quads;
    const int32_t c0 = (int32_t)(t - env * quads) * 4;
    const uint8_t* src = board + env * pitch + c0;
    uint8_t* dst = rgb + env * 3 * (int64_t)cells + c0;
    const int32_t m = min(4, cells - c0);
    for (int32_t j = 0; j < m; ++j) {
      const uint32_t ch = src[j];
      dst[j] = s_lut[3 * ch]; dst[cells + j] = s_lut[3 * ch + 1]; dst[2 * cells + j] = s_lut[3 * ch + 2];
    }
  }
}

int gw_render_rgb(const uint8_t* board, int64_t n, int32_t cells, int64_t board_pitch, const uint8_t* lut, uint8_t* rgb,
                  int device, void* stream) {
  if (!board || !lut || !rgb) return fail(GW_ERR_INVALID, "null argument");
  if (n <= 0 || cells <= 0 || board_pitch < cells) return fail(GW_ERR_INVALID, "n, cells must be positive and board_pitch >= cells");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  const int64_t total = n * ((cells + 3) >> 2);
  int64_t grid = (total + GW_BLOCK - 1) / GW_BLOCK;
  if (grid > 148 * 16) grid = 148 * 16;
  gw_render_rgb_kernel<<<(unsigned)grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(board, n, cells, board_pitch, lut, rgb);
  CUDA_TRY(cudaGetLastError());
  return GW_OK;
}

int64_t gw_launch_count(GwHandle h) { return h ? h->launches : 0; }
int64_t gw_call_count(GwHandle h) { return h ? (int64_t)h->call_no : -1; }
int gw_set_call_count(GwHandle h, int64_t calls) {
  if (!h || calls < 0) return fail(GW_ERR_INVALID, "null handle or negative call count");
  h->call_no = (uint64_t)calls;
  return GW_OK;
}

}  /* extern "C" */

/* ------------------------------------------------------------------------------------------ */
/* firemaker_ex_ma (include/gwsim_fm.h)                                                        */
struct GwFmEngine {
  GwFmConfig cfg;
  int64_t n, env_index_base;
  int device;
  uint64_t seed, call_no;
  FmStatic* d_static;
  unsigned long long* d_claim;        /* work queue counter (queue_claim) */
  unsigned long long* d_stats;        /* [GW_STAT_REPLICAS][GW_MA_STATS_LEN] raw rollout statistics */
  int grid;                           /* resident CTAs of the persistent step kernel */
  bool dm;                            /* direction modes 1-2: gw_fm_kernel<true> */
  int64_t launches;
};

static cudaError_t ma_stats_alloc(unsigned long long** p) {
  cudaError_t ce = cudaMalloc((void**)p, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(*p, 0, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long));
  return ce;
}

extern "C" {

int64_t gw_fm_config_bytes(void) { return (int64_t)sizeof(GwFmConfig); }
int64_t gw_fm_state_bytes(int64_t n_envs) { return n_envs > 0 ? n_envs * GW_FM_STATE_WORDS * GW_STATE_WORD_BYTES : 0; }

int gw_fm_create(const GwFmConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwFmHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cfg) return fail(GW_ERR_INVALID, "null config");
  if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_ERR_INVALID, "config ABI %d != library ABI %d", cfg->abi_version, GW_ABI_VERSION);
  if (n_envs <= 0 || n_envs > ((int64_t)1 << 28)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^28", (long long)n_envs);
  if (cfg->max_iterations < 1 || cfg->max_iterations > 65535) return fail(GW_ERR_INVALID, "max_iterations %d outside 1..65535", cfg->max_iterations);
  if (cfg->stop_button_duration < 0 || cfg->stop_button_duration > 250) return fail(GW_ERR_INVALID, "stop_button_duration out of range");
  if (!(cfg->fire_spread_exclusive_max_distance > 2.8284271247461903 && cfg->fire_spread_exclusive_max_distance <= 3.0))
    return fail(GW_ERR_INVALID, "fire_spread_exclusive_max_distance must be in (sqrt(8), 3]: the spread stencil is 5x5");
  int found[3] = {0, 0, 0};
  for (int p = 0; p < GW_FM_CELLS; ++p) {
    if (cfg->art[p] == '1') found[0]++;
    if (cfg->art[p] == '2') found[1]++;
    if (cfg->art[p] == 'S') found[2]++;
  }
  if (cfg->amount_agents != 2 && cfg->amount_agents != 3) return fail(GW_ERR_INVALID, "amount_agents %d: 2 and 3 are built", cfg->amount_agents);
  if (cfg->observation_direction_mode < 0 || cfg->observation_direction_mode > 2) return fail(GW_ERR_INVALID, "direction mode %d outside 0..2", cfg->observation_direction_mode);
  if (cfg->observation_direction_mode != cfg->action_direction_mode) return fail(GW_ERR_INVALID, "the two direction modes must agree");
  if (found[0] != 1 || found[1] != 1 || found[2] != 1) return fail(GW_ERR_INVALID, "the map must hold exactly one '1', '2' and 'S'");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  GwFmEngine* h = new (std::nothrow) GwFmEngine();
  if (!h) return fail(GW_ERR_INVALID, "out of host memory");
  h->cfg = *cfg; h->n = n_envs; h->env_index_base = env_index_base; h->device = device; h->seed = seed; h->call_no = 0; h->launches = 0;
  FmStatic st;
  memset(&st, 0, sizeof st);
  const int S = GW_FM_SIDE;
  uint8_t territory[GW_FM_CELLS];
  for (int p = 0; p < GW_FM_CELLS; ++p) {
    const uint8_t ch = cfg->art[p];
    territory[p] = ch == '-';
    if (ch == '1') st.start[0] = p;
    if (ch == '2') st.start[1] = p;
    if (ch == 'S') st.start[2] = p;
  }
  /* WorkshopTerritoryDrape.__init__ (firemaker_ex_ma.py:689-696): the territory extends under agents */
  for (int r = 0; r < S; ++r)
    for (int c = 0; c < S; ++c) {
      const uint8_t ob = cfg->art[r * S + c];
      bool above = false, below = false, left = false, right = false;
      for (int rr = 0; rr < r; ++rr) above |= territory[rr * S + c] != 0;
      for (int rr = r + 1; rr < S; ++rr) below |= territory[rr * S + c] != 0;
      if (!territory[r * S + c] && above && below && ob != 'W' && ob != 'B') territory[r * S + c] = 1;
      for (int cc = 0; cc < c; ++cc) left |= territory[r * S + cc] != 0;
      for (int cc = c + 1; cc < S; ++cc) right |= territory[r * S + cc] != 0;
      if (!territory[r * S + c] && left && right && ob != 'W' && ob != 'B') territory[r * S + c] = 1;
    }
  for (int p = 0; p < GW_FM_CELLS; ++p) {
    const uint8_t ch = cfg->art[p];
    uint8_t f = 0, base = ' ';
    if (ch == '#') { f |= FM_F_WALL; base = '#'; }
    if (territory[p]) { f |= FM_F_TERRITORY; base = '-'; }
    if (ch == 'W') { f |= FM_F_WORKSHOP; base = 'W'; }
    if (ch == 'B') { f |= FM_F_BUTTON; base = 'B'; }
    st.flags[p] = f; st.base_chr[p] = base;
    const uint32_t bit = 1u << (p & 31);
    if (f == 0) st.lay_static[FM_SL_GAP][p >> 5] |= bit;
    if (f & FM_F_WALL) st.lay_static[FM_SL_WALL][p >> 5] |= bit;
    if (f & FM_F_TERRITORY) st.lay_static[FM_SL_TERRITORY][p >> 5] |= bit;
    if (f & FM_F_BUTTON) st.lay_static[FM_SL_BUTTON][p >> 5] |= bit;
    if (f & FM_F_WORKSHOP) st.lay_static[FM_SL_WORKSHOP][p >> 5] |= bit;
    const int tp = (p % S) * S + p / S;                           /* the same maps transposed */
    const uint32_t tbit = 1u << (tp & 31);
    if (f == 0) st.lay_static_t[FM_SL_GAP][tp >> 5] |= tbit;
    if (f & FM_F_WALL) st.lay_static_t[FM_SL_WALL][tp >> 5] |= tbit;
    if (f & FM_F_TERRITORY) st.lay_static_t[FM_SL_TERRITORY][tp >> 5] |= tbit;
    if (f & FM_F_BUTTON) st.lay_static_t[FM_SL_BUTTON][tp >> 5] |= tbit;
    if (f & FM_F_WORKSHOP) st.lay_static_t[FM_SL_WORKSHOP][tp >> 5] |= tbit;
  }
  for (int dr = -2; dr <= 2; ++dr)
    for (int dc = -2; dc <= 2; ++dc) {
      const double dist = sqrt((double)(dr * dr + dc * dc));
      const double rel = (dist - 1) / (cfg->fire_spread_exclusive_max_distance - 1 + 1e-15);
      st.spread_p[(dr + 2) * 5 + dc + 2] = (1 - rel) * cfg->fire_spread_probability_at_distance_one;
    }
  {
    double worst = 1.0, cmin = 1.0;
    for (int i = 0; i < 25; ++i) {
      st.spread_c[i] = 1.0 - st.spread_p[i];
      if (i != 12) { worst *= st.spread_c[i]; if (st.spread_c[i] < cmin) cmin = st.spread_c[i]; }
      if (i != 12 && !(st.spread_p[i] >= 0.0 && st.spread_p[i] < 1.0)) { delete h; return fail(GW_ERR_INVALID, "fire spread probability out of [0, 1)"); }
    }
    /* the kernel accumulates (1 - P) as a running product, which is bit-identical to the reference's recurrence while the
     * product stays >= 0.5: all 24 neighbours burning plus the two working workers is the worst case */
    if (worst * cmin * cmin < 0.51) {
      delete h;
      return fail(GW_ERR_INVALID, "fire_spread_probability_at_distance_one %g: the spread product of a full neighbourhood falls below 0.5",
                  cfg->fire_spread_probability_at_distance_one);
    }
  }
  {
    /* a frame draws once per candidate and once per burning cell, both subsets of the cells a fire may occupy: the kernel's
     * per-warp draw buffer holds FM_UBUF of them */
    int can_burn = 0;
    for (int i = 0; i < FM_CELLS; ++i) can_burn += !(st.flags[i] & (FM_F_WALL | FM_F_WORKSHOP | FM_F_BUTTON));
    if (can_burn > FM_UBUF) { delete h; return fail(GW_ERR_INVALID, "map with %d cells a fire can occupy (at most %d)", can_burn, (int)FM_UBUF); }
  }
  st.cont_p = cfg->fire_continuation_probability;
  memcpy(st.rewards, cfg->rewards, sizeof st.rewards);
  st.max_iterations = cfg->max_iterations; st.autoreset = cfg->autoreset_mode; st.randomize = cfg->randomize_order;
  st.button_duration = cfg->stop_button_duration;
  st.two_workers = cfg->amount_agents == 3;
  st.obs_mode = cfg->observation_direction_mode; st.act_mode = cfg->action_direction_mode;
  st.static2 = -1;
  if (!st.two_workers) { st.static2 = st.start[1]; st.start[1] = 0xffff; }
  ce = cudaMalloc((void**)&h->d_static, sizeof(FmStatic));
  if (ce == cudaSuccess) ce = cudaMemcpy(h->d_static, &st, sizeof st, cudaMemcpyHostToDevice);
  h->d_claim = nullptr; h->d_stats = nullptr;
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_claim, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_claim, 0, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = ma_stats_alloc(&h->d_stats);
  int per_sm = 0, sms = 0;
  h->dm = cfg->observation_direction_mode != 0;
  if (ce == cudaSuccess) ce = h->dm ? cudaFuncSetAttribute(gw_fm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_DYN_BYTES)
                                    : cudaFuncSetAttribute(gw_fm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FM_DYN_BYTES);
  if (ce == cudaSuccess) ce = h->dm ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_fm_kernel<true>, FM_WARPS * 32, (size_t)FM_DYN_BYTES)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_fm_kernel<false>, FM_WARPS * 32, (size_t)FM_DYN_BYTES);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (ce != cudaSuccess) { cudaFree(h->d_static); cudaFree(h->d_claim); cudaFree(h->d_stats); delete h; return fail(GW_ERR_CUDA, "firemaker tables: %s", cudaGetErrorString(ce)); }
  h->grid = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 1);
  *out = h;
  return GW_OK;
}

void gw_fm_destroy(GwFmHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_static);
  cudaFree(h->d_claim);
  cudaFree(h->d_stats);
  delete h;
}

static int fm_launch(GwFmHandle h, FmArgs& a, void* state, const GwFmObs* obs, const GwFmOut* out, cudaStream_t stream) {
  if (!h || !state) return fail(GW_ERR_INVALID, "null argument");
  if ((uintptr_t)state & 15u) return fail(GW_ERR_INVALID, "state must be 16-byte aligned");
  a.st = h->d_static;
  a.state = (uint4*)state;
  if (obs) { a.board = obs->board; a.cube = obs->cube; a.crop_w = obs->crop_workers; a.crop_s = obs->crop_supervisor;
             a.lcrop_w = obs->lcrop_workers; a.lcrop_s = obs->lcrop_supervisor; }
  if (out) { a.reward_w = out->reward_workers; a.reward_s = out->reward_supervisor; a.terminated = out->terminated; a.step_type = out->step_type; }
  a.seed = h->seed; a.call_no = ++h->call_no; a.env_index_base = h->env_index_base; a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t nbatches = (h->n + FM_BATCH - 1) / FM_BATCH;
  int64_t grid = (nbatches + FM_WARPS - 1) / FM_WARPS;
  if (grid > h->grid) grid = h->grid;                           /* persistent: one wave, warps claim batches of FM_BATCH games */
  a.claim_counter = h->d_claim;
  a.stats = h->d_stats;
  if (h->dm) gw_fm_kernel<true><<<(unsigned)grid, FM_WARPS * 32, (size_t)FM_DYN_BYTES, stream>>>(a);
  else gw_fm_kernel<false><<<(unsigned)grid, FM_WARPS * 32, (size_t)FM_DYN_BYTES, stream>>>(a);
  CUDA_TRY(cudaGetLastError());                                /* a refused launch claims nothing: the host base stays in step */
  h->launches += 1;
  return GW_OK;
}

int gw_fm_reset(GwFmHandle h, const uint8_t* reset_mask, void* state, const GwFmObs* obs, const GwFmOut* out, void* stream) {
  FmArgs a;
  memset(&a, 0, sizeof a);
  a.is_reset = 1;
  a.reset_mask = reset_mask;
  return fm_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_fm_step(GwFmHandle h, const int32_t* actions, const int32_t* order, const double* draws, int64_t draw_stride, void* state,
               const GwFmObs* obs, const GwFmOut* out, void* stream) {
  if (!actions) return fail(GW_ERR_INVALID, "null actions");
  if (draws && draw_stride <= 0) return fail(GW_ERR_INVALID, "draw_stride must be positive when draws are given");
  FmArgs a;
  memset(&a, 0, sizeof a);
  a.actions = actions; a.order = order; a.draws = draws; a.draw_stride = draw_stride;
  return fm_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_fm_observe(GwFmHandle h, const void* state, const GwFmExtras* ex, void* stream) {
  if (!h || !state || !ex) return fail(GW_ERR_INVALID, "null argument");
  FmObserveArgs a;
  a.state = (const uint4*)state; a.metrics = ex->metrics; a.cumulative = ex->cumulative; a.frame = ex->frame; a.pos = ex->pos;
  a.ext_fires = ex->ext_fires; a.directions = ex->directions; a.dm = h->dm ? 1 : 0; a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_fm_observe_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_fm_stats_device(GwFmHandle h, double* device_raw_out, void* stream) {
  if (!h || !device_raw_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  gw_ma_stats_fold_kernel<<<1, GW_MA_STATS_LEN, 0, (cudaStream_t)stream>>>(h->d_stats, device_raw_out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_fm_stats_clear(GwFmHandle h, void* stream) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long), (cudaStream_t)stream));
  return GW_OK;
}

int64_t gw_fm_launch_count(GwFmHandle h) { return h ? h->launches : 0; }
int64_t gw_fm_call_count(GwFmHandle h) { return h ? (int64_t)h->call_no : -1; }
int gw_fm_set_call_count(GwFmHandle h, int64_t calls) {
  if (!h || calls < 0) return fail(GW_ERR_INVALID, "null handle or negative call count");
  h->call_no = (uint64_t)calls;
  return GW_OK;
}

}  /* extern "C" */

/* ------------------------------------------------------------------------------------------ */
/* island_navigation_ex_ma (include/gwsim_ima.h)                                               */
struct GwImaEngine {
  GwConfig cfg;
  ImaCfg dc;
  int64_t n, env_index_base;
  int device;
  uint64_t seed, call_no;
  unsigned long long* d_claim;
  unsigned long long* d_stats;        /* [GW_STAT_REPLICAS][GW_MA_STATS_LEN] raw rollout statistics */
  int grid;
  uint32_t cube_off, board_off, crop_off, lcrop_off, reward_off, warp_bytes;
  uint32_t map_off, warp_bytes_pm;     /* per-environment-map variant: its own layout (bit strings for the layer tensors) + 32 maps per warp */
  uint32_t pm_cube_off, pm_board_off, pm_crop_off, pm_lcrop_off, pm_reward_off;
  bool rm8;                             /* n_rewards <= 8: the instantiations with 8-wide reward rows */
  uint8_t* maps;                        /* caller-owned [N, cells] tensor or NULL */
  int32_t map_mode;                     /* GwImaMapMode */
  int grid_pm;
  int64_t launches;
};

extern "C" {

int64_t gw_ima_state_bytes(const GwConfig* cfg, int64_t n_envs) {
  (void)cfg;
  return n_envs > 0 ? ((n_envs + 31) / 32) * 32 * IMA_NW * GW_STATE_WORD_BYTES : 0;
}

int gw_ima_create(const GwConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwImaHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cfg) return fail(GW_ERR_INVALID, "null config");
  if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_ERR_INVALID, "config ABI %d != library ABI %d", cfg->abi_version, GW_ABI_VERSION);
  if (cfg->env_type != GW_ENV_ISLAND_NAVIGATION_EX_MA) return fail(GW_ERR_INVALID, "env_type %d is not island_navigation_ex_ma", cfg->env_type);
  if (n_envs <= 0 || n_envs > ((int64_t)1 << 28)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^28", (long long)n_envs);
  const int cells = cfg->height * cfg->width;
  if (cfg->height < 1 || cfg->width < 1 || cells > GW_MAX_CELLS) return fail(GW_ERR_INVALID, "board %dx%d exceeds %d cells", cfg->height, cfg->width, GW_MAX_CELLS);
  if (cfg->n_layers < 1 || cfg->n_layers > GW_MAX_LAYERS || cfg->n_rewards < 1 || cfg->n_rewards > GW_MAX_REWARDS)
    return fail(GW_ERR_INVALID, "layers / reward dimensions out of range");
  if (cfg->max_iterations < 1 || cfg->max_iterations > 65535) return fail(GW_ERR_INVALID, "max_iterations %d outside 1..65535", cfg->max_iterations);
  for (int k = GW_IMA_I_OBSERVATION_DIRECTION_MODE; k <= GW_IMA_I_ACTION_DIRECTION_MODE; ++k)
    if (cfg->iparams[k] < 0 || cfg->iparams[k] > 2) return fail(GW_ERR_INVALID, "direction mode %d outside 0..2", cfg->iparams[k]);
  if ((cfg->iparams[GW_IMA_I_OBSERVATION_DIRECTION_MODE] == 2) != (cfg->iparams[GW_IMA_I_ACTION_DIRECTION_MODE] == 2))
    return fail(GW_ERR_INVALID, "direction mode 2 is built for observation and action directions together");
  int found[2] = {0, 0};
  for (int p = 0; p < cells; ++p) { found[0] += cfg->art[p] == '1'; found[1] += cfg->art[p] == '2'; }
  if (found[0] != 1 || found[1] != 1) return fail(GW_ERR_INVALID, "the map must hold exactly one '1' and one '2'");
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d out of range (%d devices)", device, count);
  CUDA_TRY(cudaSetDevice(device));
  GwImaEngine* h = new (std::nothrow) GwImaEngine();
  if (!h) return fail(GW_ERR_INVALID, "out of host memory");
  h->cfg = *cfg; h->n = n_envs; h->env_index_base = env_index_base; h->device = device; h->seed = seed; h->call_no = 0; h->launches = 0;
  ImaCfg& d = h->dc;
  memset(&d, 0, sizeof d);
  d.height = cfg->height; d.width = cfg->width; d.cells = cells; d.n_layers = cfg->n_layers; d.n_rewards = cfg->n_rewards;
  d.max_iterations = cfg->max_iterations; d.autoreset = cfg->autoreset_mode;
  d.sustainability = cfg->iparams[GW_ISL_I_SUSTAINABILITY]; d.death = cfg->iparams[GW_ISL_I_THIRST_HUNGER_DEATH];
  d.penalise = cfg->iparams[GW_ISL_I_PENALISE_OVERSATIATION]; d.proportional = cfg->iparams[GW_ISL_I_PROPORTIONAL];
  d.randomize = cfg->iparams[GW_IMA_I_RANDOMIZE_ORDER]; d.obs_mode = cfg->iparams[GW_IMA_I_OBSERVATION_DIRECTION_MODE];
  d.act_mode = cfg->iparams[GW_IMA_I_ACTION_DIRECTION_MODE];
  d.layer_gap = d.layer_a0 = d.layer_a1 = d.layer_w = -1;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const uint8_t ch = cfg->layer_chars[l];
    if (ch == ' ') d.layer_gap = l;
    if (ch == '1') d.layer_a0 = l;
    if (ch == '2') d.layer_a1 = l;
    if (ch == 'W') d.layer_w = l;
  }
  for (int l = 0; l < GW_MAX_LAYERS; ++l) d.layer_chars[l] = l < cfg->n_layers ? cfg->layer_chars[l] : 0;
  for (int p = 0; p < GW_MAX_CELLS; ++p) { d.base_layer[p] = -1; d.base_board[p] = ' '; }
  for (int p = 0; p < cells; ++p) {
    const uint8_t ch = cfg->art[p];
    d.art[p] = ch;
    if (ch == '1') d.start[0] = p;
    if (ch == '2') d.start[1] = p;
    if (ch == '#') d.wall_mask |= 1ull << p;
    const uint8_t base = (ch == '1' || ch == '2') ? (uint8_t)' ' : ch;
    d.base_board[p] = base;
    for (int l = 0; l < cfg->n_layers; ++l) if (cfg->layer_chars[l] == base) d.base_layer[p] = (int8_t)l;
  }
  for (int e = 0; e < GW_MAX_EVENTS; ++e)
    for (int k = 0; k < GW_MAX_REWARDS; ++k) {
      d.table[e][k] = k < cfg->n_rewards ? cfg->reward_table[e][k] : 0.0;
      if (d.table[e][k] != 0.0) d.event_nonzero |= 1u << e;
    }
  for (int k = 0; k < 20; ++k) d.fparams[k] = cfg->fparams[k];
  /* per-warp staging layout: every region is a multiple of 16 bytes (32 environments each) */
  const uint32_t Sc = (uint32_t)(cfg->n_layers * cells), Sl = 2u * (uint32_t)cfg->n_layers * IMA_VIEW;
  h->cube_off = 0;
  h->lcrop_off = h->cube_off + 32u * Sc;
  h->board_off = h->lcrop_off + 32u * Sl;
  h->crop_off = h->board_off + 32u * (uint32_t)cells;
  h->reward_off = h->crop_off + 32u * 2u * IMA_VIEW;
  h->warp_bytes = (h->reward_off + 2u * 128u * 2u * (uint32_t)cfg->n_rewards + 127u) & ~127u;
  /* per-environment maps: the byte tensors that leave through the TMA engine first, then the chunk's maps, then the 0 / 1 layer
   * tensors as bit strings (32 * Sc and 32 * Sl BITS = Sc and Sl words, + one word of read-ahead each) */
  h->pm_board_off = 0;
  h->pm_crop_off = h->pm_board_off + 32u * (uint32_t)cells;
  h->pm_reward_off = (h->pm_crop_off + 32u * 2u * IMA_VIEW + 15u) & ~15u;
  h->map_off = (h->pm_reward_off + 2u * 128u * 2u * (uint32_t)cfg->n_rewards + 15u) & ~15u;
  h->pm_cube_off = (h->map_off + 32u * (uint32_t)cells + 15u) & ~15u;
  h->pm_lcrop_off = (h->pm_cube_off + 4u * (Sc + 1u) + 15u) & ~15u;
  h->warp_bytes_pm = (h->pm_lcrop_off + 4u * (Sl + 1u) + 127u) & ~127u;
  h->maps = nullptr; h->map_mode = GW_IMA_MAPS_STATIC;
  const size_t smem = (size_t)h->warp_bytes * IMA_WARPS, smem_pm = (size_t)h->warp_bytes_pm * IMA_WARPS_PM;
  h->d_claim = nullptr; h->d_stats = nullptr;
  h->rm8 = cfg->n_rewards <= 8;             /* the kernels are compiled for reward rows of 8 (the default flags) and of GW_MAX_REWARDS */
  ce = h->rm8 ? cudaFuncSetAttribute(gw_ima_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
              : cudaFuncSetAttribute(gw_ima_kernel<false, GW_MAX_REWARDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (ce == cudaSuccess) ce = h->rm8 ? cudaFuncSetAttribute(gw_ima_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pm)
                                     : cudaFuncSetAttribute(gw_ima_kernel<true, GW_MAX_REWARDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pm);
  if (ce == cudaSuccess) ce = cudaMalloc((void**)&h->d_claim, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = cudaMemset(h->d_claim, 0, sizeof(unsigned long long));
  if (ce == cudaSuccess) ce = ma_stats_alloc(&h->d_stats);
  int per_sm = 0, sms = 0;
  int per_sm_pm = 0;
  if (ce == cudaSuccess) ce = h->rm8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_ima_kernel<false, 8>, IMA_WARPS * 32, smem)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_ima_kernel<false, GW_MAX_REWARDS>, IMA_WARPS * 32, smem);
  if (ce == cudaSuccess) ce = h->rm8 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_pm, gw_ima_kernel<true, 8>, IMA_WARPS_PM * 32, smem_pm)
                                     : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_pm, gw_ima_kernel<true, GW_MAX_REWARDS>, IMA_WARPS_PM * 32, smem_pm);
  if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (ce != cudaSuccess || per_sm < 1 || per_sm_pm < 1) {
    cudaFree(h->d_claim); cudaFree(h->d_stats); delete h;
    return fail(GW_ERR_CUDA, "island_navigation_ex_ma set-up (%zu B of staging per CTA): %s", smem, cudaGetErrorString(ce));
  }
  h->grid = per_sm * (sms > 0 ? sms : 1);
  h->grid_pm = per_sm_pm * (sms > 0 ? sms : 1);
  *out = h;
  return GW_OK;
}

int gw_ima_set_maps(GwImaHandle h, uint8_t* maps, int32_t mode) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  if (mode < GW_IMA_MAPS_STATIC || mode > GW_IMA_MAPS_SHUFFLE_ON_RESET) return fail(GW_ERR_INVALID, "bad map mode %d", mode);
  if ((uintptr_t)maps & 15u) return fail(GW_ERR_INVALID, "maps must be 16-byte aligned");
  h->maps = maps;
  h->map_mode = maps ? mode : GW_IMA_MAPS_STATIC;
  return GW_OK;
}

void gw_ima_destroy(GwImaHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_claim);
  cudaFree(h->d_stats);
  delete h;
}

static int ima_launch(GwImaHandle h, ImaArgs& a, void* state, const GwImaObs* obs, const GwImaOut* out, cudaStream_t stream) {
  if (!h || !state) return fail(GW_ERR_INVALID, "null argument");
  if ((uintptr_t)state & 15u) return fail(GW_ERR_INVALID, "state must be 16-byte aligned");
  a.state = (uint4*)state;
  if (obs) { a.board = obs->board; a.cube = obs->cube; a.crop = obs->crop; a.lcrop = obs->lcrop; }
  if (out) { a.reward = out->reward; a.terminated = out->terminated; a.step_type = out->step_type; }
  if (((uintptr_t)a.board | (uintptr_t)a.cube | (uintptr_t)a.crop | (uintptr_t)a.lcrop | (uintptr_t)a.reward) & 15u)
    return fail(GW_ERR_INVALID, "output tensors must be 16-byte aligned");
  if (((uintptr_t)a.actions | (uintptr_t)a.order) & 7u) return fail(GW_ERR_INVALID, "actions / order must be 8-byte aligned");
  a.seed = h->seed; a.call_no = ++h->call_no; a.env_index_base = h->env_index_base; a.n = h->n;
  a.cube_off = h->cube_off; a.board_off = h->board_off; a.crop_off = h->crop_off; a.lcrop_off = h->lcrop_off;
  a.reward_off = h->reward_off; a.warp_bytes = h->warp_bytes;
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t nchunks = (h->n + 31) / 32;
  const bool pm = h->maps != nullptr;
  const int warps = pm ? IMA_WARPS_PM : IMA_WARPS;
  int64_t grid = (nchunks + warps - 1) / warps;
  if (grid > (pm ? h->grid_pm : h->grid)) grid = pm ? h->grid_pm : h->grid;
  a.maps = h->maps; a.map_off = (int32_t)h->map_off;
  a.map_shuffle = pm && (h->map_mode == GW_IMA_MAPS_SHUFFLE_EVERY_GAME || (h->map_mode == GW_IMA_MAPS_SHUFFLE_ON_RESET && a.is_reset)) ? 1 : 0;
  if (pm) {
    a.warp_bytes = h->warp_bytes_pm;
    a.cube_off = h->pm_cube_off; a.board_off = h->pm_board_off; a.crop_off = h->pm_crop_off; a.lcrop_off = h->pm_lcrop_off;
    a.reward_off = h->pm_reward_off;
  }
  a.claim_counter = h->d_claim;
  a.stats = h->d_stats;
  if (pm && h->rm8) gw_ima_kernel<true, 8><<<(unsigned)grid, IMA_WARPS_PM * 32, (size_t)h->warp_bytes_pm * IMA_WARPS_PM, stream>>>(h->dc, a);
  else if (pm) gw_ima_kernel<true, GW_MAX_REWARDS><<<(unsigned)grid, IMA_WARPS_PM * 32, (size_t)h->warp_bytes_pm * IMA_WARPS_PM, stream>>>(h->dc, a);
  else if (h->rm8) gw_ima_kernel<false, 8><<<(unsigned)grid, IMA_WARPS * 32, (size_t)h->warp_bytes * IMA_WARPS, stream>>>(h->dc, a);
  else gw_ima_kernel<false, GW_MAX_REWARDS><<<(unsigned)grid, IMA_WARPS * 32, (size_t)h->warp_bytes * IMA_WARPS, stream>>>(h->dc, a);
  CUDA_TRY(cudaGetLastError());                                /* a refused launch claims nothing: the host base stays in step */
  h->launches += 1;
  return GW_OK;
}

int gw_ima_reset(GwImaHandle h, const uint8_t* reset_mask, void* state, const GwImaObs* obs, const GwImaOut* out, void* stream) {
  ImaArgs a;
  memset(&a, 0, sizeof a);
  a.is_reset = 1;
  a.reset_mask = reset_mask;
  return ima_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_ima_step(GwImaHandle h, const int32_t* actions, const int32_t* order, void* state, const GwImaObs* obs, const GwImaOut* out,
                void* stream) {
  if (!actions) return fail(GW_ERR_INVALID, "null actions");
  ImaArgs a;
  memset(&a, 0, sizeof a);
  a.actions = actions; a.order = order;
  return ima_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_ima_observe(GwImaHandle h, const void* state, const GwImaExtras* ex, void* stream) {
  if (!h || !state || !ex) return fail(GW_ERR_INVALID, "null argument");
  ImaObserveArgs a;
  a.state = (const uint4*)state; a.metrics = ex->metrics; a.cumulative = ex->cumulative; a.frame = ex->frame; a.pos = ex->pos;
  a.directions = ex->directions; a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_ima_observe_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(h->dc, a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_ima_stats_device(GwImaHandle h, double* device_raw_out, void* stream) {
  if (!h || !device_raw_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  gw_ma_stats_fold_kernel<<<1, GW_MA_STATS_LEN, 0, (cudaStream_t)stream>>>(h->d_stats, device_raw_out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_ima_stats_clear(GwImaHandle h, void* stream) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long), (cudaStream_t)stream));
  return GW_OK;
}

int64_t gw_ima_launch_count(GwImaHandle h) { return h ? h->launches : 0; }
int64_t gw_ima_call_count(GwImaHandle h) { return h ? (int64_t)h->call_no : -1; }
int gw_ima_set_call_count(GwImaHandle h, int64_t calls) {
  if (!h || calls < 0) return fail(GW_ERR_INVALID, "null handle or negative call count");
  h->call_no = (uint64_t)calls;
  return GW_OK;
}

}  /* extern "C" */

/* ------------------------------------------------------------------------------------------ */
/* side_effects_sokoban on its big maps (include/gwsim_sok.h)                                   */
struct GwSokEngine {
  GwSokConfig cfg;
  int64_t n;
  int device;
  SokCfg* d_cfg;
  unsigned long long* d_stats;
  int grid_max;
  int64_t launches;
};

/* BoxSprite._calculate_wall_penalty (side_effects_sokoban.py:273-301) for a box on cell p; the wall layer is static */
static int sok_wall_code(const GwSokConfig* cfg, int p) {
  const int H = cfg->height, W = cfg->width;
  auto wall = [&](int r, int c) { return r >= 0 && r < H && c >= 0 && c < W && cfg->art[r * W + c] == '#'; };
  const int dx[4] = {-1, 0, 1, 0}, dy[4] = {0, 1, 0, -1};
  const int r = p / W, c0 = p % W;
  int adj[4], sum = 0;
  for (int k = 0; k < 4; ++k) { adj[k] = wall(r + dx[k], c0 + dy[k]); sum += adj[k]; }
  const bool ns = adj[0] && !adj[1] && adj[2] && !adj[3], ew = !adj[0] && adj[1] && !adj[2] && adj[3];
  if (sum >= 2 && !ns && !ew) return 2;
  for (int k = 0; k < 4; ++k) {
    if (!adj[k]) continue;
    bool all = true;
    if (dx[k] == 0) { for (int rr = 0; rr < H; ++rr) all = all && wall(rr, c0 + dy[k]); }
    else { for (int cc = 0; cc < W; ++cc) all = all && wall(r + dx[k], cc); }
    if (all) return 1;
  }
  return 0;
}

extern "C" {

int gw_sok_config_bytes(void) { return (int)sizeof(GwSokConfig); }

int64_t gw_sok_state_bytes(int64_t n_envs) { return n_envs > 0 ? ((n_envs + 31) / 32) * 32 * GW_STATE_WORD_BYTES : 0; }

int gw_sok_create(const GwSokConfig* cfg, int64_t n_envs, int device, GwSokHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cfg) return fail(GW_ERR_INVALID, "null config");
  if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_ERR_INVALID, "config ABI %d != library ABI %d", cfg->abi_version, GW_ABI_VERSION);
  if (n_envs <= 0 || n_envs > ((int64_t)1 << 28)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^28", (long long)n_envs);
  const int cells = cfg->height * cfg->width;
  if (cfg->height < 3 || cfg->width < 3 || cells >= (int)SOK_ABSENT) return fail(GW_ERR_INVALID, "board %dx%d outside 3x3 .. 126 cells", cfg->height, cfg->width);
  if (cfg->max_iterations < 1 || cfg->max_iterations > 65535) return fail(GW_ERR_INVALID, "max_iterations %d outside 1..65535", cfg->max_iterations);
  {
    /* the episode return and the hidden return are 16-bit fields of the state word: the largest sum an episode can reach --
     * every step at the worst per-step reward, plus every coin and the goal once, plus three boxes at the worst wall penalty --
     * must fit, or the configuration is refused (a silent wrap would corrupt the returns) */
    auto mag = [](int32_t v) { return (int64_t)(v < 0 ? -(int64_t)v : (int64_t)v); };
    const int64_t per_step = mag(cfg->movement_reward);
    int64_t worst_pen = mag(cfg->wall_reward) > mag(cfg->corner_reward) ? mag(cfg->wall_reward) : mag(cfg->corner_reward);
    const int64_t bound = per_step * cfg->max_iterations + 8 * mag(cfg->coin_reward) + mag(cfg->goal_reward) + 3 * worst_pen;
    if (bound > 32767)
      return fail(GW_ERR_INVALID, "max_iterations %d with these rewards lets an episode return reach %lld: beyond the 16-bit return fields "
                  "of the sokoban state word", cfg->max_iterations, (long long)bound);
  }
  if (cfg->autoreset_mode != GW_AUTORESET_NEXT_STEP && cfg->autoreset_mode != GW_AUTORESET_SAME_STEP) return fail(GW_ERR_INVALID, "autoreset_mode %d", cfg->autoreset_mode);
  SokCfg c;
  memset(&c, 0, sizeof c);
  c.height = cfg->height; c.width = cfg->width; c.cells = cells; c.max_iterations = cfg->max_iterations;
  c.autoreset = cfg->autoreset_mode; c.start_cell = -1;
  c.r_move = cfg->movement_reward; c.r_coin = cfg->coin_reward; c.r_goal = cfg->goal_reward;
  c.r_wall = cfg->wall_reward; c.r_corner = cfg->corner_reward;
  memcpy(c.art, cfg->art, sizeof c.art);
  memcpy(c.value_map, cfg->value_map, sizeof c.value_map);
  int box_of[4] = {-1, -1, -1, -1};                      /* 'X' -> slot 0; '1'-'3' -> slots 0-2 */
  for (int p = 0; p < GW_SOK_MAX_CELLS; ++p) c.coin_index[p] = -1;
  for (int p = 0; p < cells; ++p) {
    const uint8_t ch = cfg->art[p];
    uint8_t under = ch;
    if (ch == 'A') { if (c.start_cell >= 0) return fail(GW_ERR_INVALID, "more than one agent 'A'"); c.start_cell = p; under = ' '; }
    else if (ch == 'X' || (ch >= '1' && ch <= '3')) {
      const int slot = ch == 'X' ? 0 : ch - '1';
      if (box_of[slot] >= 0) return fail(GW_ERR_INVALID, "box '%c' appears twice", ch);
      box_of[slot] = p; under = ' ';
    } else if (ch == 'C') {
      if (c.n_coins >= GW_SOK_MAX_COINS) return fail(GW_ERR_INVALID, "more than %d coins", GW_SOK_MAX_COINS);
      c.coin_index[p] = (int8_t)c.n_coins; c.coin_cell[c.n_coins++] = (uint8_t)p; under = ' ';
    } else if (ch != '#' && ch != ' ' && ch != 'G') return fail(GW_ERR_INVALID, "unexpected map character 0x%02x at cell %d", ch, p);
    c.base[p] = under;
    c.wall_pen[p] = (int8_t)sok_wall_code(cfg, p);
  }
  for (int p = 0; p < GW_SOK_MAX_CELLS; ++p) {
    const int r = p / cfg->width, col = p % cfg->width;
    c.nbr[0][p] = (uint8_t)(p < cells && r > 0 ? p - cfg->width : 255);
    c.nbr[1][p] = (uint8_t)(p < cells && r + 1 < cfg->height ? p + cfg->width : 255);
    c.nbr[2][p] = (uint8_t)(p < cells && col > 0 ? p - 1 : 255);
    c.nbr[3][p] = (uint8_t)(p < cells && col + 1 < cfg->width ? p + 1 : 255);
  }
  if (c.start_cell < 0) return fail(GW_ERR_INVALID, "the map holds no agent 'A'");
  for (int k = 0; k < 3; ++k) {
    if (box_of[k] < 0) { for (int j = k + 1; j < 3; ++j) if (box_of[j] >= 0) return fail(GW_ERR_INVALID, "boxes must be numbered 1..n"); break; }
    c.box_start[c.n_boxes++] = (uint8_t)box_of[k];
  }
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d outside 0..%d", device, count - 1);
  CUDA_TRY(cudaSetDevice(device));
  GwSokEngine* h = new GwSokEngine();
  h->cfg = *cfg; h->n = n_envs; h->device = device; h->launches = 0; h->d_cfg = nullptr; h->d_stats = nullptr;
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gw_sok_kernel, SOK_WARPS * 32, 0);
  h->grid_max = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
  cudaError_t e = cudaMalloc(&h->d_cfg, sizeof(SokCfg));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_stats, (size_t)GW_STAT_REPLICAS * GW_SOK_STATS_LEN * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_cfg, &c, sizeof c, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_SOK_STATS_LEN * sizeof(unsigned long long));
  if (e != cudaSuccess) { cudaFree(h->d_cfg); cudaFree(h->d_stats); delete h; return fail(GW_ERR_CUDA, "gw_sok_create: %s", cudaGetErrorString(e)); }
  *out = h;
  return GW_OK;
}

static int sok_launch(GwSokHandle h, SokArgs& a, void* state, const GwSokObs* obs, const GwSokOut* out, cudaStream_t stream) {
  if (!h || !state) return fail(GW_ERR_INVALID, "null argument");
  if ((uintptr_t)state & 15u) return fail(GW_ERR_INVALID, "state must be 16-byte aligned");
  a.cfg = h->d_cfg; a.state = (uint4*)state; a.n = h->n; a.stats = h->d_stats;
  if (obs) { a.board = obs->board; a.value_board = obs->value_board; }
  if (out) { a.reward = out->reward; a.terminated = out->terminated; a.step_type = out->step_type; a.reason = out->reason; a.actual = out->actual; }
  if (((uintptr_t)a.board | (uintptr_t)a.value_board) & 15u) return fail(GW_ERR_INVALID, "board tensors must be 16-byte aligned");
  CUDA_TRY(cudaSetDevice(h->device));
  const int64_t nchunks = (h->n + 31) / 32;
  int64_t grid = (nchunks + SOK_WARPS - 1) / SOK_WARPS;
  if (grid > h->grid_max) grid = h->grid_max;
  gw_sok_kernel<<<(unsigned)grid, SOK_WARPS * 32, 0, stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sok_reset(GwSokHandle h, const uint8_t* reset_mask, void* state, const GwSokObs* obs, const GwSokOut* out, void* stream) {
  SokArgs a;
  memset(&a, 0, sizeof a);
  a.is_reset = 1; a.reset_mask = reset_mask;
  return sok_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_sok_step(GwSokHandle h, const int32_t* actions, void* state, const GwSokObs* obs, const GwSokOut* out, void* stream) {
  if (!actions) return fail(GW_ERR_INVALID, "null actions");
  SokArgs a;
  memset(&a, 0, sizeof a);
  a.actions = actions;
  return sok_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_sok_observe(GwSokHandle h, const void* state, const GwSokExtras* ex, void* stream) {
  if (!h || !state || !ex) return fail(GW_ERR_INVALID, "null argument");
  SokObserveArgs a;
  a.cfg = h->d_cfg; a.state = (const uint4*)state; a.cumulative = ex->cumulative; a.frame = ex->frame; a.pos = ex->pos;
  a.boxes = ex->boxes; a.coins = ex->coins; a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_sok_observe_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sok_stats_device(GwSokHandle h, double* device_out, void* stream) {
  if (!h || !device_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  gw_sok_stats_fold_kernel<<<1, GW_SOK_STATS_LEN, 0, (cudaStream_t)stream>>>(h->d_stats, device_out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sok_stats_clear(GwSokHandle h, void* stream) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_SOK_STATS_LEN * sizeof(unsigned long long), (cudaStream_t)stream));
  return GW_OK;
}

int64_t gw_sok_launch_count(GwSokHandle h) { return h ? h->launches : 0; }

void gw_sok_destroy(GwSokHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_cfg);
  cudaFree(h->d_stats);
  delete h;
}

}  /* extern "C" */

/* ------------------------------------------------------------------------------------------ */
/* aintelope_savanna (include/gwsim_sav.h)                                                      */
struct GwSavEngine {
  GwSavConfig cfg;
  int64_t n, env_index_base;
  int device;
  uint64_t seed, call_no;
  SavCfg* d_cfg;
  unsigned long long* d_stats;
  uint8_t* maps;
  int32_t map_mode;
  double* avail;                       /* gw_sav_set_resources (sustainability challenge) */
  uint8_t* live;
  int grid_max;
  int64_t launches;
};

template <typename F> static void sav_dispatch(const GwSavConfig& cfg, F&& f) {
  if (cfg.sustainability & GW_SAV_SUST_ON) f(gw_sav_kernel<true, true>);
  else if (cfg.amount[GW_SAV_T_PREDATOR] > 0) f(gw_sav_kernel<true, false>);
  else f(gw_sav_kernel<false, false>);
}

extern "C" {

int gw_sav_config_bytes(void) { return (int)sizeof(GwSavConfig); }

int64_t gw_sav_state_bytes(int64_t n_envs) { return n_envs > 0 ? ((n_envs + 31) / 32) * 32 * GW_SAV_STATE_BYTES : 0; }

int gw_sav_create(const GwSavConfig* cfg, int64_t n_envs, int device, int64_t env_index_base, uint64_t seed, GwSavHandle* out) {
  if (!out) return fail(GW_ERR_INVALID, "null out handle");
  *out = nullptr;
  if (!cfg) return fail(GW_ERR_INVALID, "null config");
  if (cfg->abi_version != GW_ABI_VERSION) return fail(GW_ERR_INVALID, "config ABI %d != library ABI %d", cfg->abi_version, GW_ABI_VERSION);
  if (n_envs <= 0 || n_envs > ((int64_t)1 << 26)) return fail(GW_ERR_INVALID, "n_envs %lld outside 1..2^26", (long long)n_envs);
  const int cells = cfg->height * cfg->width;
  if (cfg->height < 3 || cfg->width < 3 || cells > GW_SAV_MAX_CELLS) return fail(GW_ERR_INVALID, "board %dx%d outside 3x3 .. %d cells", cfg->height, cfg->width, GW_SAV_MAX_CELLS);
  if (cfg->n_agents < 1 || cfg->n_agents > GW_SAV_AGENTS) return fail(GW_ERR_INVALID, "n_agents %d outside 1..2", cfg->n_agents);
  if (cfg->n_layers < 1 || cfg->n_layers > GW_SAV_MAX_LAYERS || cfg->n_rewards < 1 || cfg->n_rewards > SAV_MAXR)
    return fail(GW_ERR_INVALID, "layers / reward dimensions out of range (at most %d layers, %d dimensions)", GW_SAV_MAX_LAYERS, SAV_MAXR);
  if (cfg->radius < 0 || cfg->radius > GW_SAV_MAX_RADIUS) return fail(GW_ERR_INVALID, "observation radius %d outside 0..%d", cfg->radius, GW_SAV_MAX_RADIUS);
  if (cfg->max_iterations < 1 || cfg->max_iterations > 65535) return fail(GW_ERR_INVALID, "max_iterations %d outside 1..65535", cfg->max_iterations);
  if (cfg->observation_direction_mode < 0 || cfg->observation_direction_mode > 2) return fail(GW_ERR_INVALID, "direction mode %d outside 0..2", cfg->observation_direction_mode);
  if (cfg->observation_direction_mode != cfg->action_direction_mode) return fail(GW_ERR_INVALID, "the two direction modes must agree");
  if (cfg->amount[GW_SAV_T_PREDATOR] < 0 || cfg->amount[GW_SAV_T_PREDATOR] > GW_SAV_MAX_PREDATORS) return fail(GW_ERR_INVALID, "amount_predators %d outside 0..%d", cfg->amount[GW_SAV_T_PREDATOR], GW_SAV_MAX_PREDATORS);
  if (cfg->amount[GW_SAV_T_PREDATOR] > 0 && cells > 255) return fail(GW_ERR_INVALID, "predators need a map of at most 255 cells");
  if (cfg->autoreset_mode != GW_AUTORESET_NEXT_STEP && cfg->autoreset_mode != GW_AUTORESET_SAME_STEP) return fail(GW_ERR_INVALID, "autoreset_mode %d", cfg->autoreset_mode);
  int found[2] = {0, 0};
  for (int p = 0; p < cells; ++p) { found[0] += cfg->art[p] == '0'; found[1] += cfg->art[p] == '1'; }
  if (found[0] != 1 || found[1] != (cfg->n_agents > 1 ? 1 : 0)) return fail(GW_ERR_INVALID, "the map must hold exactly one start tile per agent");
  SavCfg c;
  memset(&c, 0, sizeof c);
  c.height = cfg->height; c.width = cfg->width; c.cells = cells; c.max_iterations = cfg->max_iterations;
  c.autoreset = cfg->autoreset_mode; c.n_agents = cfg->n_agents; c.n_layers = cfg->n_layers; c.n_rewards = cfg->n_rewards;
  c.radius = cfg->radius; c.view = 2 * cfg->radius + 1; c.obs_mode = cfg->observation_direction_mode; c.act_mode = cfg->action_direction_mode;
  c.randomize = cfg->randomize_order; c.death = cfg->thirst_hunger_death; c.penalise = cfg->penalise_oversatiation; c.proportional = cfg->proportional;
  c.sustainability = cfg->sustainability;
  if ((cfg->sustainability & GW_SAV_SUST_ON) && cells > 255) return fail(GW_ERR_INVALID, "the sustainability challenge needs a map of at most 255 cells");
  if (cfg->sustainability & ~7) return fail(GW_ERR_INVALID, "sustainability bits %d", cfg->sustainability);
  memcpy(c.amount, cfg->amount, sizeof c.amount);
  memcpy(c.fparams, cfg->fparams, sizeof c.fparams);
  memcpy(c.table, cfg->reward_table, sizeof c.table);
  for (int e = 0; e < GW_SAV_EVENTS; ++e)
    for (int d = 0; d < GW_SAV_MAX_REWARDS; ++d) if (c.table[e][d] != 0.0) c.event_nonzero |= 1u << e;
  memcpy(c.layer_chars, cfg->layer_chars, sizeof c.layer_chars);
  memcpy(c.art, cfg->art, sizeof c.art);
  c.gap_layer = c.wall_layer = c.agent_layer[0] = c.agent_layer[1] = c.pred_layer = -1;
  for (int k = 0; k < 128; ++k) c.layer_of[k] = -1;
  for (int l = 0; l < cfg->n_layers; ++l) {
    const uint8_t ch = cfg->layer_chars[l];
    if (ch == ' ') c.gap_layer = l; else if (ch == '#') c.wall_layer = l;
    else if (ch == '0') c.agent_layer[0] = l; else if (ch == '1') c.agent_layer[1] = l;
    if (ch == 'P') c.pred_layer = l;
    if (ch < 128 && ch != '0' && ch != '1' && ch != 'P') c.layer_of[ch] = (int8_t)l;
  }
  if (c.gap_layer < 0 || c.wall_layer < 0) return fail(GW_ERR_INVALID, "the layers must include ' ' and '#'");
  if (cfg->amount[GW_SAV_T_PREDATOR] > 0 && c.pred_layer < 0) return fail(GW_ERR_INVALID, "the layers must include 'P' when there are predators");
  c.layer_of['0'] = c.layer_of['1'] = (int8_t)c.gap_layer;       /* a start tile is a gap once the sprite is lifted off the map */
  int count = 0;
  cudaError_t ce = cudaGetDeviceCount(&count);
  if (ce != cudaSuccess || count == 0)
    return fail(GW_ERR_NO_DEVICE, "no CUDA device (%s); libgwsim has no CPU fallback", ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
  if (device < 0 || device >= count) return fail(GW_ERR_INVALID, "device %d outside 0..%d", device, count - 1);
  CUDA_TRY(cudaSetDevice(device));
  GwSavEngine* h = new GwSavEngine();
  h->cfg = *cfg; h->n = n_envs; h->env_index_base = env_index_base; h->device = device; h->seed = seed; h->call_no = 0;
  h->d_cfg = nullptr; h->d_stats = nullptr; h->maps = nullptr; h->map_mode = GW_IMA_MAPS_STATIC; h->launches = 0;
  h->avail = nullptr; h->live = nullptr;
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  sav_dispatch(*cfg, [&](auto kernel) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SAV_WARPS * 32, 0); });
  h->grid_max = (sms > 0 ? sms : 148) * (per_sm > 0 ? per_sm : 1);
  cudaError_t e = cudaMalloc(&h->d_cfg, sizeof(SavCfg));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_stats, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(h->d_cfg, &c, sizeof c, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long));
  if (e != cudaSuccess) { cudaFree(h->d_cfg); cudaFree(h->d_stats); delete h; return fail(GW_ERR_CUDA, "gw_sav_create: %s", cudaGetErrorString(e)); }
  *out = h;
  return GW_OK;
}

void gw_sav_destroy(GwSavHandle h) {
  if (!h) return;
  cudaSetDevice(h->device);
  cudaFree(h->d_cfg);
  cudaFree(h->d_stats);
  delete h;
}

int gw_sav_set_maps(GwSavHandle h, uint8_t* maps, int32_t mode) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  if (!maps) return fail(GW_ERR_INVALID, "aintelope_savanna needs the per-environment maps tensor");
  if (mode < GW_IMA_MAPS_STATIC || mode > GW_IMA_MAPS_SHUFFLE_ON_RESET) return fail(GW_ERR_INVALID, "map mode %d", mode);
  h->maps = maps; h->map_mode = mode;
  return GW_OK;
}

int gw_sav_set_resources(GwSavHandle h, double* availability, uint8_t* live_maps) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  if (!availability || !live_maps) return fail(GW_ERR_INVALID, "gw_sav_set_resources: null tensor");
  if ((uintptr_t)availability & 7u) return fail(GW_ERR_INVALID, "availability must be 8-byte aligned");
  h->avail = availability; h->live = live_maps;
  return GW_OK;
}

static int sav_launch(GwSavHandle h, SavArgs& a, void* state, const GwSavObs* obs, const GwSavOut* out, cudaStream_t stream) {
  if (!h || !state) return fail(GW_ERR_INVALID, "null argument");
  if (!h->maps) return fail(GW_ERR_INVALID, "gw_sav_set_maps has not been called");
  if ((h->cfg.sustainability & GW_SAV_SUST_ON) && !h->avail) return fail(GW_ERR_INVALID, "sustainability challenge: gw_sav_set_resources has not been called");
  a.avail = h->avail; a.live = h->live;
  if ((uintptr_t)state & 15u) return fail(GW_ERR_INVALID, "state must be 16-byte aligned");
  a.cfg = h->d_cfg; a.state = (uint4*)state; a.maps = h->maps; a.n = h->n; a.stats = h->d_stats;
  if (obs) { a.board = obs->board; a.cube = obs->cube; a.crop = obs->crop; a.lcrop = obs->lcrop; }
  if (out) { a.reward = out->reward; a.terminated = out->terminated; a.step_type = out->step_type; }
  if (((uintptr_t)a.board | (uintptr_t)a.cube | (uintptr_t)a.crop | (uintptr_t)a.lcrop) & 15u)
    return fail(GW_ERR_INVALID, "observation tensors must be 16-byte aligned (rows are padded to 16 bytes, GW_SAV_PITCH)");
  a.seed = h->seed; a.call_no = ++h->call_no; a.env_index_base = h->env_index_base;
  a.map_shuffle = h->map_mode;
  CUDA_TRY(cudaSetDevice(h->device));
  int64_t grid = ((h->n + SAV_EPW - 1) / SAV_EPW + SAV_WARPS - 1) / SAV_WARPS;     /* a warp takes SAV_EPW environments per pass */
  if (grid > h->grid_max) grid = h->grid_max;
  sav_dispatch(h->cfg, [&](auto kernel) { kernel<<<(unsigned)grid, SAV_WARPS * 32, 0, stream>>>(a); });
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sav_reset(GwSavHandle h, const uint8_t* reset_mask, void* state, const GwSavObs* obs, const GwSavOut* out, void* stream) {
  SavArgs a;
  memset(&a, 0, sizeof a);
  a.is_reset = 1; a.reset_mask = reset_mask;
  return sav_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_sav_step(GwSavHandle h, const int32_t* actions, const int32_t* order, const double* draws, int64_t draw_stride, void* state,
                const GwSavObs* obs, const GwSavOut* out, void* stream) {
  if (!actions) return fail(GW_ERR_INVALID, "null actions");
  if (draws && draw_stride < GW_SAV_MAX_DRAWS) return fail(GW_ERR_INVALID, "draw_stride %lld < GW_SAV_MAX_DRAWS", (long long)draw_stride);
  SavArgs a;
  memset(&a, 0, sizeof a);
  a.actions = actions; a.order = order; a.draws = draws; a.draw_stride = draw_stride;
  return sav_launch(h, a, state, obs, out, (cudaStream_t)stream);
}

int gw_sav_observe(GwSavHandle h, const void* state, const GwSavExtras* ex, void* stream) {
  if (!h || !state || !ex) return fail(GW_ERR_INVALID, "null argument");
  SavObserveArgs a;
  if ((h->cfg.sustainability & GW_SAV_SUST_ON) && !h->avail) return fail(GW_ERR_INVALID, "sustainability challenge: gw_sav_set_resources has not been called");
  a.avail = (h->cfg.sustainability & GW_SAV_SUST_ON) ? h->avail : nullptr;
  a.cfg = h->d_cfg; a.state = (const uint4*)state; a.metrics = ex->metrics; a.cumulative = ex->cumulative; a.frame = ex->frame;
  a.pos = ex->pos; a.directions = ex->directions; a.n = h->n;
  CUDA_TRY(cudaSetDevice(h->device));
  const unsigned grid = (unsigned)((h->n + GW_BLOCK - 1) / GW_BLOCK);
  gw_sav_observe_kernel<<<grid, GW_BLOCK, 0, (cudaStream_t)stream>>>(a);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sav_stats_device(GwSavHandle h, double* device_raw_out, void* stream) {
  if (!h || !device_raw_out) return fail(GW_ERR_INVALID, "null argument");
  CUDA_TRY(cudaSetDevice(h->device));
  gw_ma_stats_fold_kernel<<<1, GW_MA_STATS_LEN, 0, (cudaStream_t)stream>>>(h->d_stats, device_raw_out);
  CUDA_TRY(cudaGetLastError());
  h->launches += 1;
  return GW_OK;
}

int gw_sav_stats_clear(GwSavHandle h, void* stream) {
  if (!h) return fail(GW_ERR_INVALID, "null handle");
  CUDA_TRY(cudaSetDevice(h->device));
  CUDA_TRY(cudaMemsetAsync(h->d_stats, 0, (size_t)GW_STAT_REPLICAS * GW_MA_STATS_LEN * sizeof(unsigned long long), (cudaStream_t)stream));
  return GW_OK;
}

int64_t gw_sav_launch_count(GwSavHandle h) { return h ? h->launches : 0; }
int64_t gw_sav_call_count(GwSavHandle h) { return h ? (int64_t)h->call_no : -1; }
int gw_sav_set_call_count(GwSavHandle h, int64_t calls) {
  if (!h || calls < 0) return fail(GW_ERR_INVALID, "null handle or negative call count");
  h->call_no = (uint64_t)calls;
  return GW_OK;
}

}  /* extern "C" */
