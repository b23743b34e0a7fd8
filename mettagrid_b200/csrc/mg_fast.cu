// mg_fast.cu -- k_step_fast<G>: MettaGrid::_step (bindings/mettagrid_c.cpp:921-1102) for handler-free
// ("plain") programs in sparse environments: at most 32 objects per env, none of them ever created or
// removed.  This is the reference's own benchmark game (benchmarks/test_mettagrid_env_benchmark.py:21-29).
//
// Mapping: G = 8, 16 or 32 lanes own one environment (32 / G environments per warp).  Lane l plays two
// roles: agent l (actions, bookkeeping, observer) and object slot l + 1 (location, vibe, token cache).
// No phase runs on one lane only:
//   * state comes from the env's packed block (mg_state.h, MGFB_*) in ONE load wave; the program header is a
//     kernel argument (constant bank);
//   * the shuffle (:958-964) is decoded one draw per lane, and every lane finds the agent that acts at its step by
//     walking std::shuffle's transpositions backwards;
//   * every agent decodes its own actions; noop / change_vibe never interact across agents; moves resolve in the
//     shuffled order (:958-999) with one ballot per step: the target cell is free iff no object lane holds that
//     location -- the occupancy grid is never read (and not maintained: nothing reads it for such a handle;
//     mg_reset rebuilds it);
//   * observations (:665-824): each agent lane builds one sort key per object (Manhattan rank, packed offset, token
//     count, object), sorts them in registers with a Batcher network and hands every visible object its first
//     token position; the OBJECT lanes then write their cached tokens into each observer's row of a shared-memory
//     stage of the env's whole [A][T][3] block, which the group streams out with 16-byte stores;
//   * `cell.visited` (:787-796) goes to the lowest agent index that sees an object (one ballot per object);
//   * stat updates are decided in registers and written back once, in the reference's float-add order.
// k_fast_pack / k_fast_unpack move the packed block from / to the generic arrays (mg_capi.cu decides when).
//
// k_step_fast<G, true> -- the static layer.  In a handler-free game nothing ever changes a non-agent object (walls,
// blocks): no handler, event or AOE exists that could move, re-tag or refill it.  Maps with hundreds of them stay on
// this kernel by keeping them OUT of the object lanes: the lanes hold the agents only, and the env carries
//   * an occupancy bitmap of the padded grid (move test: one bit; observation: each window row is one funnel shift),
//   * the list of static cells with one `visited` stamp each (:787-796).
// An observer's window becomes a 256-bit mask indexed by the packed window offset; the position of any token is
// (global tokens) + (tokens of dynamic objects earlier in Manhattan order) + (static tokens) x popcount(mask & the
// precomputed set of offsets that come earlier).  The STATIC OBJECTS are then spread over all lanes: each finds its
// observers from per-row / per-column agent masks, stamps `visited` for the lowest one and writes its token(s) into
// every observer's row -- balanced over the lanes however unevenly the walls are spread over the windows.
#include <cuda_runtime.h>

#include "mg_state.h"

#ifndef MG_FAST_CTAS_PER_SM
#define MG_FAST_CTAS_PER_SM 16  // one-warp CTAs per SM the register budget allows (128 registers)
#endif
#ifndef MG_FS_WL
#define MG_FS_WL 8  // static variant: work-list entries per lane (16 measured equal; 8 lets a 16th one-warp CTA fit an SM on toy)
#endif
#ifndef MG_FAST_NO_BULK
#define MG_FAST_NO_BULK 0  // 1: always stream the observation block with vector stores (A/B against cp.async.bulk)
#endif

namespace {

#define FAST_INVALID 0xFFFFFFFFu

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
  uint32_t y = (cur & 0x80000000u) | (nxt & 0x7fffffffu);
  return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// one std::mt19937 output straight from / to the env's state in global memory
__device__ __forceinline__ uint32_t rng_direct(uint32_t* rng, int& idx) {
  int i = idx >= MG_RNG_WORDS ? 0 : idx;
  int i1 = i + 1 == MG_RNG_WORDS ? 0 : i + 1;
  int i2 = i + 397 >= MG_RNG_WORDS ? i + 397 - MG_RNG_WORDS : i + 397;
  uint32_t nw = mt_twist(rng[i], rng[i1], rng[i2]);
  rng[i] = nw;
  idx = i + 1;
  return mt_temper(nw);
}
// Lemire multiply-shift with rejection (bits/uniform_int_dist.h:252-282)
__device__ __forceinline__ uint32_t rng_below_direct(uint32_t* rng, int& idx, uint32_t range) {
  uint64_t prod = (uint64_t)rng_direct(rng, idx) * range;
  uint32_t low = (uint32_t)prod;
  if (low < range) {
    uint32_t thr = (0u - range) % range;
    while (low < thr) {
      prod = (uint64_t)rng_direct(rng, idx) * range;
      low = (uint32_t)prod;
    }
  }
  return (uint32_t)(prod >> 32);
}
// libstdc++ std::shuffle (bits/stl_algo.h:3719-3805), serial, for the rare tick where a draw is rejected or the
// draws straddle the end of the 624-word state
__device__ __noinline__ void shuffle_serial(uint32_t* rng, uint32_t* idx_word, uint8_t* v, int n) {
  int idx = (int)*idx_word;
  int i = 1;
  if ((n & 1) == 0) {
    int j = (int)rng_below_direct(rng, idx, 2);
    uint8_t t = v[i];
    v[i] = v[j], v[j] = t;
    i++;
  }
  while (i < n) {
    uint32_t sr = (uint32_t)i + 1;
    uint32_t x = rng_below_direct(rng, idx, sr * (sr + 1));
    int p0 = (int)(x / (sr + 1)), p1 = (int)(x % (sr + 1));
    uint8_t t = v[i];
    v[i] = v[p0], v[p0] = t;
    i++;
    t = v[i];
    v[i] = v[p1], v[p1] = t;
    i++;
  }
  *idx_word = (uint32_t)idx;
}

__device__ __forceinline__ void set_error(int32_t* E, int code, int info) {
  const int cur = E[MGEV_ERROR];
  if (!(cur & code)) E[MGEV_ERR_INFO] = info;
  E[MGEV_ERROR] = cur | code;
}

struct TokDims {
  const int32_t* P;
  int TW, ND, B;
  int cap, ftag, fvibe, fgroup, fagent, inv_feats, templates;  // header words, by value
};
// An object's observation tokens (core/grid_object.cpp:178-203, objects/agent.cpp:142-154) as
// (feature | value << 8) pairs, written to `tk`; same rules as rebuild_token_cache in mg_kernels.cu.
__device__ __noinline__ int build_tokens(const TokDims d, const uint32_t* o, uint32_t meta, uint16_t* tk, int32_t* E, bool live) {
  const int cap = d.cap;
  int n = 0;
  auto put = [&](int feat, int val) {
    if (n < cap) tk[n] = (uint16_t)((feat & 0xff) | ((val & 0xff) << 8));
    n++;
  };
  for (int k = 0; k < d.TW; k++) {
    uint32_t m = o[MGO_TAGS + k];
    while (m) {
      int b = __ffs(m) - 1;
      put(d.ftag, k * 32 + b);
      m &= m - 1;
    }
  }
  const int vibe = (int)((meta >> 16) & 0xffu);
  if (vibe) put(d.fvibe, vibe);
  const int fl = (int)(meta >> 24);
  if (fl & MGOF_OBS_INV) {
    const uint64_t ord = (uint64_t)o[MGO_INVORD_LO] | ((uint64_t)o[MGO_INVORD_HI] << 32);
    const int cnt = (int)(ord >> 60);
    const uint16_t* inv = (const uint16_t*)(o + MGO_TAGS + d.TW);
    const int32_t* feats = d.P + d.inv_feats;
    for (int i = 0; i < cnt; i++) {
      const int it = (int)((ord >> (4 * i)) & 15u);
      uint32_t amt = inv[it];
      int p = 0;
      do {
        put(__ldg(feats + it * d.ND + p), (int)(amt % (uint32_t)d.B));
        amt /= (uint32_t)d.B;
        p++;
      } while (amt > 0 && p < d.ND);
    }
  }
  if (fl & MGOF_AGENT) {
    const int ai = (int)o[MGO_AGENT];
    put(d.fgroup, __ldg(d.P + d.templates + (int)(meta & 0xffffu) * MG_TEMPLATE_WORDS + MGT_GROUP));
    put(d.fagent, ai >= 0 ? ai : 0);
  }
  if (n > cap) {
    if (live) set_error(E, MGERR_POOL_EXHAUSTED, 19);
    n = cap;
  }
  return n;
}

__device__ __forceinline__ void ce(uint32_t& a, uint32_t& b) {
  const uint32_t lo = min(a, b), hi = max(a, b);
  a = lo, b = hi;
}
// Sorting networks on registers, fully unrolled: Batcher's odd-even merge sort for 8 and 16 keys (19 and 63
// exchanges; tests/test_sort_network.py checks the lists below with the 0-1 principle), the bitonic network for 32
// keys (240 exchanges).
#define CE(a, b) ce(key[a], key[b]);
__device__ __forceinline__ void sort_net(uint32_t (&key)[8]) {
  /* SORT8 */
    CE(0, 1) CE(2, 3) CE(4, 5) CE(6, 7) CE(0, 2) CE(1, 3) CE(4, 6) CE(5, 7) CE(1, 2) CE(5, 6) CE(0, 4) CE(1, 5)
    CE(2, 6) CE(3, 7) CE(2, 4) CE(3, 5) CE(1, 2) CE(3, 4) CE(5, 6)
}
__device__ __forceinline__ void sort_net(uint32_t (&key)[16]) {
  /* SORT16 */
    CE(0, 1) CE(2, 3) CE(4, 5) CE(6, 7) CE(8, 9) CE(10, 11) CE(12, 13) CE(14, 15) CE(0, 2) CE(1, 3) CE(4, 6) CE(5, 7)
    CE(8, 10) CE(9, 11) CE(12, 14) CE(13, 15) CE(1, 2) CE(5, 6) CE(9, 10) CE(13, 14) CE(0, 4) CE(1, 5) CE(2, 6)
    CE(3, 7) CE(8, 12) CE(9, 13) CE(10, 14) CE(11, 15) CE(2, 4) CE(3, 5) CE(10, 12) CE(11, 13) CE(1, 2) CE(3, 4)
    CE(5, 6) CE(9, 10) CE(11, 12) CE(13, 14) CE(0, 8) CE(1, 9) CE(2, 10) CE(3, 11) CE(4, 12) CE(5, 13) CE(6, 14)
    CE(7, 15) CE(4, 8) CE(5, 9) CE(6, 10) CE(7, 11) CE(2, 4) CE(3, 5) CE(6, 8) CE(7, 9) CE(10, 12) CE(11, 13) CE(1, 2)
    CE(3, 4) CE(5, 6) CE(7, 8) CE(9, 10) CE(11, 12) CE(13, 14)
}
#undef CE
__device__ __forceinline__ void sort_net(uint32_t (&key)[32]) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll
      for (int i = 0; i < 32; i++) {
        const int l = i ^ j;
        if (l > i) {
          if ((i & k) == 0)
            ce(key[i], key[l]);
          else
            ce(key[l], key[i]);
        }
      }
    }
  }
}

// WPC = warps per CTA.  One-warp CTAs retire and refill SM slots env by env: 17.0 vs 18.4 us per tick at the headline
// 4096 envs x 16 agents (less than one wave) and within 1 % of four-warp CTAs at 32 768 and 262 144 envs, so the
// sorted-key variant always runs them; the static variant, whose CTAs share a 0xFF row besides the window table,
// gains a few per cent from four-warp CTAs once the grid is many waves deep (toy 16 384 envs: 175 vs 183 us).
// mg_fast_layout picks.
template <int G, bool S, int WPC>
__global__ void __launch_bounds__(WPC * 32, MG_FAST_CTAS_PER_SM / WPC) k_step_fast(const MgDev d, const MgFastLayout L, const MgFastHdr HD) {
  extern __shared__ __align__(16) unsigned char smem[];
  constexpr int GPW = 32 / G;  // environments per warp
  const int* const hdr = HD.v;  // the program header travels as a kernel argument: constant-bank reads, no load
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gl = lane & (G - 1), grp = lane / G;
  const uint32_t gmask = G == 32 ? MG_FULL : (((1u << (G & 31)) - 1u) << (grp * G));
  const int gshift = grp * G;

  int env = (blockIdx.x * WPC + warp) * GPW + grp;
  const bool live = env < d.num_envs;  // a dead group mirrors the last env but stores nothing
  if (!live) env = d.num_envs - 1;
  uint32_t* lut = (uint32_t*)smem;  // window table: packed offset -> rank << 24 | offset << 16 (0xFFFFFF00 outside)
  unsigned char* gb = smem + L.cta_bytes + (size_t)(warp * GPW + grp) * L.group_bytes;
  uint32_t* toks = (uint32_t*)(gb + L.tok_off);
  uint32_t* oloc = (uint32_t*)(gb + L.oloc_off);
  uint32_t* ontok = oloc + G;  // (oloc, ontok) are used as one uint2[G] table: location, count << 8 | lane
  uint32_t* stale = ontok + G;
  uint32_t* draw = stale + G;
  uint8_t* order = (uint8_t*)(draw + G);

  const int A = d.A, T = d.T;
  const size_t g0 = (size_t)env * A;
  const bool isA = gl < A;
  const size_t ga = g0 + (isA ? gl : 0);
  int32_t* E = d.env + (size_t)env * MGEV_WORDS;  // only the error words live here
  uint32_t* rng = d.rng + (size_t)env * MG_RNG_WORDS;
  uint32_t* blk = d.fast_blk + (size_t)env * d.fast_stride;  // the env's packed hot state (mg_state.h)
  uint32_t* bag = blk + MGFB_AGENT(G, gl);
  uint32_t* bob = blk + MGFB_OBJECT(G, gl);
  float* bst = (float*)(blk + MGFB_STAT(G, 0, gl));  // stat id at bst[id * G]
  uint32_t* const sb = (uint32_t*)(gb + L.sb_off);  // static variant: the env's static block (header + bitmap)
  if (S) {  // fire-and-forget copy of the block into shared memory; waited for before the moves
    const uint32_t* src = d.fast_static + (size_t)env * d.fast_sstride;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sb);
#pragma unroll 1
    for (int v = gl; v < (L.sb_words >> 2); v += G)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16u * v), "l"(src + 4 * v) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  uint32_t* cvrow = d.cover + ga * d.CW;
  const int TOKOFF = MG_TOKOFF(d.TW, d.R);
  const int tokw = L.tok_stride;  // words per object in the shared token table

  // ---- wave 1: every load whose address is known up front is issued before anything is consumed
  const uint4 bh0 = *(const uint4*)blk;        // step, rng index, objects, tokens_written
  const uint4 bh1 = *(const uint4*)(blk + 4);  // tokens_free, touched game stats
  const int ia_raw = d.actions[ga], iv_raw = d.vibe_actions[ga];
  const uint4 ag0 = *(const uint4*)bag;        // obj slot, spawn, prev_location, steps_without_motion
  const uint4 ag1 = *(const uint4*)(bag + 4);  // max_dist, unique cells, touched bits
  const uint4 orec = *(const uint4*)bob;       // loc, visited, meta, agent + 1 | ntok << 8 | dirty << 16
  const uint4 otok = *(const uint4*)(bob + 4); // first eight cached tokens
  const float s_cv0 = bst[hdr[MGH_ST_CELL_VISITED] * G];
  constexpr int LV = (64 + WPC * 32 - 1) / (WPC * 32);  // window-table vectors per thread
  uint4 lutv[LV];
#pragma unroll
  for (int k = 0; k < LV; k++) {
    const int i = tid + k * WPC * 32;
    lutv[k] = i < 64 ? __ldg((const uint4*)d.rank_lut + i) : make_uint4(0, 0, 0, 0);
  }
  if (gl < MGFB_STATS / 2) {  // the stat rows the end of the tick may update: pull them towards L2 now
    asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + MGFB_STAT(G, 2 * gl, 0)));
    if (G > 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + MGFB_STAT(G, 2 * gl + 1, 0)));
    if (G > 16) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + MGFB_STAT(G, 2 * gl, 16)));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(blk + MGFB_STAT(G, 2 * gl + 1, 16)));
    }
  }

  const uint32_t* slist = d.fast_slist + (size_t)env * d.fast_ns_cap;
  uint32_t* svis = d.fast_svis + (size_t)env * d.fast_ns_cap;
  if (S) {  // the static cells and their stamps are read late: pull their lines towards L2 now
#pragma unroll 1
    for (int i = gl * 32; i < d.fast_ns_cap; i += G * 32) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(slist + i));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(svis + i));
    }
  }
  uint8_t* gobs = d.obs + g0 * (size_t)(3 * T);
  const int nbytes = A * 3 * T;
#pragma unroll
  for (int k = 0; k < LV; k++) {
    const int i = tid + k * WPC * 32;
    if (i < 64) ((uint4*)lut)[i] = lutv[k];
  }
  const uint32_t step = bh0.x + 1u;  // :951
  int idx0 = (int)bh0.y;
  const int nobj = (int)bh0.z;
  const int ia = isA ? ia_raw : -1, iv = isA ? iv_raw : -1;
  // the generic record behind this lane's object (long token tails, cache rebuilds): slot l + 1, or -- static variant,
  // where the lanes hold the agents' objects -- the agent's own slot
  uint32_t* o = d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + (size_t)min(S ? (int)ag0.x : gl + 1, d.maxobj - 1)) * d.OS;
  const uint32_t a_slot = isA ? ag0.x : 1u, a_spawn = ag0.y, a_prev = ag0.z, a_swm = ag0.w, a_maxd = ag1.x;
  uint32_t a_unique = ag1.y;
  const bool isO = gl < nobj;
  uint32_t o_loc = orec.x, o_vis = orec.y, o_meta = orec.z;
  uint32_t o_ntok = (orec.w & MGFB_DIRTY) ? MG_TOK_DIRTY : ((orec.w >> 8) & 0xffu);
  const uint32_t ntok_raw = o_ntok;
  const int o_agent = isO ? (int)(orec.w & 0xffu) - 1 : -1;
  uint32_t tw0 = otok.x, tw1 = otok.y, tw2 = otok.z, tw3 = otok.w;
  const bool o_alive = isO && ((o_meta >> 24) & MGOF_ALIVE);
  if (!o_alive) o_loc = FAST_INVALID;
  const uint32_t o_loc0 = o_loc, o_vis0 = o_vis, o_meta0 = o_meta;
  // ---- wave 2: action rows, RNG words, the rest of long token caches
  const int NA = hdr[MGH_NUM_ACTIONS];
  const int32_t* acts = d.P + hdr[MGS_ACTIONS];
  const bool inv_p = ia < 0 || ia >= NA, inv_v = iv < 0 || iv >= NA;
  int4 ap = make_int4(0, 0, 0, 1), av = make_int4(0, 0, 0, 0);  // kind, arg, priority, is_vibe
  if (isA && !inv_p) ap = __ldg((const int4*)(acts + ia * MG_ACTION_WORDS));
  if (isA && !inv_v) av = __ldg((const int4*)(acts + iv * MG_ACTION_WORDS));
  const int ndraws = A < 2 ? 0 : ((A & 1) ? (A - 1) / 2 : A / 2);
  if (idx0 >= MG_RNG_WORDS) idx0 = 0;
  const bool window_ok = idx0 + ndraws <= MG_RNG_WORDS;
  uint32_t r_cur = 0, r_nxt = 0, r_far = 0;
  if (window_ok && gl < ndraws) {
    const int i = idx0 + gl;
    const int i1 = i + 1 == MG_RNG_WORDS ? 0 : i + 1;
    const int i2 = i + 397 >= MG_RNG_WORDS ? i + 397 - MG_RNG_WORDS : i + 397;
    r_cur = rng[i], r_nxt = rng[i1], r_far = rng[i2];
  }
  {
    const bool cached = o_alive && o_ntok != MG_TOK_DIRTY;
    const int nwords = cached ? (int)(o_ntok + 1) >> 1 : 0;
    uint32_t* row = toks + gl * tokw;
    row[0] = tw0, row[1] = tw1, row[2] = tw2, row[3] = tw3;  // tokw >= 5
#pragma unroll 1
    for (int k = 4; k < nwords; k++) row[k] = o[TOKOFF + k];  // long lists keep their tail in the object record
  }
  // independent work while wave 2 is in flight: the observation stage starts as all EmptyTokenByte (:940-942);
  // it shares the destination's 16-byte phase so that whole vectors can be streamed out
  // Static variant: only the first L.stage_tokens tokens of a row are staged (rows are mostly padding; the padding is
  // stored straight from registers and the rare longer row writes its tail to HBM directly), which keeps four times as
  // many environments resident per SM.
  uint8_t* stage = S ? gb : gb + ((uint32_t)(uintptr_t)gobs & 15u);
  const int SC = L.stage_tokens, RS = 3 * SC;  // static variant: staged tokens / bytes per row
  if (S) {  // the CTA's constant 0xFF buffer, source of every row's padding (read by the bulk-copy engine)
    uint4* f4 = (uint4*)(smem + 1024);
    for (int v = tid; v < ((3 * T + 15) >> 4); v += WPC * 32) f4[v] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (!S) {
    uint4* s4 = (uint4*)gb;
    const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
    const int nv = (nbytes + 15 + 16) >> 4;
    int v = gl;
#pragma unroll 1
    for (; v + 3 * G < nv; v += 4 * G) s4[v] = ff, s4[v + G] = ff, s4[v + 2 * G] = ff, s4[v + 3 * G] = ff;
#pragma unroll 1
    for (; v < nv; v += G) s4[v] = ff;
  }
  stale[gl] = 0;
  uint32_t* const rowm = (uint32_t*)(gb + L.rc_off);  // static variant: agents whose window covers row r / column c
  uint32_t* const colm = rowm + d.H;
  if (S) {
#pragma unroll 1
    for (int i = gl; i < d.H + d.W; i += G) rowm[i] = 0;
  }
  const bool act_p = isA && !inv_p && ap.w == 0;  // executed in the primary stream (:966-999)
  const bool act_v = isA && !inv_v && av.w == 1;  // executed in the vibe stream

  // ---- shuffle (:958-964): std::shuffle is a sequence of transpositions (pos, P[pos]), pos = 1..A-1.  Every draw
  // is decoded by its own lane; every lane then finds the element that ends at its position by walking the
  // transpositions backwards -- no serial swap loop.
  bool reject = !window_ok;
  uint32_t nw = 0;
  uint8_t* perm = order;  // P[pos]
  if (window_ok && gl < ndraws) {
    nw = mt_twist(r_cur, r_nxt, r_far);
    const uint32_t rnd = mt_temper(nw);
    const bool single = (A & 1) == 0 && gl == 0;
    const int i = single ? 1 : ((A & 1) ? 2 * gl + 1 : 2 * gl);
    const uint32_t sr = (uint32_t)i + 1;
    const uint32_t range = single ? 2u : sr * (sr + 1);
    const uint64_t prod = (uint64_t)rnd * range;
    const uint32_t low = (uint32_t)prod, x = (uint32_t)(prod >> 32);
    if (low < range && low < (0u - range) % range) reject = true;
    // x < sr * (sr + 1) <= 1056 and sr + 1 <= 33: the float quotient of (x + 0.5) is never within 0.015 of an
    // integer, far beyond the error of the fast division, so truncation gives the exact integer quotient
    const uint32_t q = __float2uint_rz(__fdividef((float)x + 0.5f, (float)(sr + 1)));
    perm[i] = (uint8_t)(single ? x : q);
    if (!single) perm[i + 1] = (uint8_t)(x - q * (sr + 1));
  }
  const uint32_t rej = __ballot_sync(MG_FULL, reject) & gmask;
  __syncwarp();
  int my_order = gl;  // the agent that acts at step gl
  if (rej) {        // rare: a rejected draw or the draws straddle the end of the state -- one lane replays serially
    __syncwarp(gmask);
    if (isA) order[gl] = (uint8_t)gl;
    __syncwarp(gmask);
    if (gl == 0 && A >= 2 && live) shuffle_serial(rng, blk + MGFB_RNG_IDX, order, A);  // a mirror group leaves the state alone
    __syncwarp(gmask);
    if (isA) my_order = order[gl];
  } else {
#pragma unroll 4
    for (int pos = A - 1; pos >= 1; pos--) {
      const int pp = perm[pos];
      my_order = my_order == pos ? pp : (my_order == pp ? pos : my_order);
    }
  }
  if (!rej && live && gl < ndraws) rng[idx0 + gl] = nw;
  if (!rej && live && gl == 0 && ndraws > 0) blk[MGFB_RNG_IDX] = (uint32_t)(idx0 + ndraws);
  if (S) asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();  // the window table (per CTA), the static block (per group)

  // ---- moves in shuffled order, highest priority first (actions/move.hpp:81-115 with the two default
  // handlers: relocate into an empty in-map cell, else fail).  Lane i first fetches what the agent acting at
  // step i wants; the loop then only carries the object locations from step to step.
  const int my_ol = S ? gl : (int)a_slot - 1;  // lane that plays this agent's object
  const uint32_t* const sbm = sb + MGFS_HDR_WORDS;  // static occupancy: bit c + PAD of row r + PAD
  const int SBW = d.fast_sbw;
  const uint32_t my_loc0 = __shfl_sync(MG_FULL, o_loc, my_ol, G);
  if (isA && my_loc0 != FAST_INVALID)  // the coverage word this agent most likely needs (it moves at most one cell)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(cvrow + (((my_loc0 >> 16) * d.W + (my_loc0 & 0xffffu)) >> 5)));
  uint32_t my_tgt = FAST_INVALID;
  const bool wants_move = act_p && ap.x == MGA_MOVE;
  if (wants_move) {  // actions/orientation.hpp:28-48
    const int arg = ap.y;
    const int dr = (arg == 0 || arg == 4 || arg == 5) ? -1 : (arg == 1 || arg == 6 || arg == 7) ? 1 : 0;
    const int dc = (arg == 2 || arg == 4 || arg == 6) ? -1 : (arg == 3 || arg == 5 || arg == 7) ? 1 : 0;
    const int tr = (int)(my_loc0 >> 16) + dr, tc = (int)(my_loc0 & 0xffffu) + dc;
    if (my_loc0 != FAST_INVALID && tr >= 0 && tr < d.H && tc >= 0 && tc < d.W) {
      my_tgt = ((uint32_t)tr << 16) | (uint32_t)tc;
      if (S) {  // a static object in the way: TargetLocEmpty fails, and nothing there has an on_use handler
        const int cb = tc + d.PAD;
        if ((sbm[(tr + d.PAD) * SBW + (cb >> 5)] >> (cb & 31)) & 1u) my_tgt = FAST_INVALID;
      }
    }
  }
  {
    // the step at which each agent acts (inverse of the shuffled order), handed to the lane that plays its object
    if (isA) order[my_order] = (uint8_t)gl;
    __syncwarp();
    const int a_step = isA ? (int)order[gl] : -1;
    const int oa = o_agent >= 0 && o_agent < A ? o_agent : 0;
    const int o_step = o_agent >= 0 && o_agent < A ? __shfl_sync(MG_FULL, a_step, oa, G) : (__shfl_sync(MG_FULL, a_step, oa, G), -1);
    const int maxp = hdr[MGH_MAX_PRIORITY], pmask = hdr[MGH_PRIORITY_MASK];
#pragma unroll 1
    for (int prio = maxp; prio >= 0; prio--) {
      if (!((pmask >> prio) & 1)) continue;
      const bool mine = wants_move && ap.z == prio;
      if (!__any_sync(MG_FULL, mine)) continue;
      const uint32_t tgt_a = mine ? my_tgt : FAST_INVALID;
      const uint32_t tgt_s = __shfl_sync(MG_FULL, tgt_a, my_order, G);  // target of the agent acting at step gl
      const uint32_t tgt_o = __shfl_sync(MG_FULL, tgt_a, oa, G);        // target of this lane's object
#pragma unroll
      for (int i = 0; i < G; i++) {  // steps beyond the agent count carry no target and change nothing
        if (S && i >= A) break;
        const uint32_t tgt = __shfl_sync(MG_FULL, tgt_s, i, G);
        const uint32_t occ = __ballot_sync(MG_FULL, o_loc == tgt) & gmask;
        if (o_step == i && tgt_o != FAST_INVALID && occ == 0) o_loc = tgt_o;  // TargetLocEmpty -> Relocate
      }
    }
  }
  const uint32_t my_loc = __shfl_sync(MG_FULL, o_loc, my_ol, G);  // observer position (:1049-1052)
  const bool move_ok = wants_move && my_loc != my_loc0;  // a successful move always changes the location

  // ---- per-agent outcome of both streams (actions/action_handler.hpp:78-105), in execution order
  const bool v_first = act_p && act_v && av.z > ap.z;  // the vibe-stream action has the higher priority
  const bool ok_p = act_p && (ap.x == MGA_MOVE ? move_ok : true);
  const bool ok_v = act_v && av.x != MGA_MOVE;
  int exec_idx = 0;
  int new_vibe = -1;
  {
    const bool cp = ok_p && ap.x == MGA_CHANGE_VIBE, cv = ok_v && av.x == MGA_CHANGE_VIBE;
    if (v_first) {
      if (ok_v) exec_idx = iv;
      if (ok_p) exec_idx = ia;
      if (cv) new_vibe = av.y;
      if (cp) new_vibe = ap.y;
    } else {
      if (ok_p) exec_idx = ia;
      if (ok_v) exec_idx = iv;
      if (cp) new_vibe = ap.y;
      if (cv) new_vibe = av.y;
    }
  }

  // ---- wave 3 loads: stats, touched bits, coverage word (prefetched into L2 by wave 1)
  const int npass = hdr[MGH_MAX_PRIORITY] + 1;
  // steps-without-motion chain over the executed actions (stream order, like k_step)
  uint32_t prev = a_prev, swm = a_swm;
  uint32_t swm_peak = 0;  // largest value steps_without_motion took this tick
  const int nacted = (act_p ? 1 : 0) + (act_v ? 1 : 0);
#pragma unroll
  for (int k = 0; k < 2; k++) {
    if (k < nacted) {
      if (my_loc == prev) {
        swm += 1;
        swm_peak = max(swm_peak, swm);
      } else {
        swm = 0;
      }
      prev = my_loc;
    }
  }
  const int id_p = act_p ? hdr[(ok_p ? MGH_ST_NOOP_SUCCESS : MGH_ST_NOOP_FAILED) + 2 * ap.x] : -1;
  const int id_v = act_v ? hdr[(ok_v ? MGH_ST_NOOP_SUCCESS : MGH_ST_NOOP_FAILED) + 2 * av.x] : -1;
  const int nfail = (act_p && !ok_p ? 1 : 0) + (act_v && !ok_v ? 1 : 0);
  const int ninv = (isA && inv_p ? 1 : 0) + (isA && inv_v ? 1 : 0);
  const int id_fail = hdr[MGH_ST_ACTION_FAILED], id_inv = hdr[MGH_ST_INVALID_INDEX], id_swm = hdr[MGH_ST_MAX_SWM];
  const int id_cv = hdr[MGH_ST_CELL_VISITED], id_un = hdr[MGH_ST_UNIQUE_VISITED], id_md = hdr[MGH_ST_MAX_DIST];
  float s_p = 0.f, s_v = 0.f, s_fail = 0.f, s_inv = 0.f, s_swm = 0.f;
  const float s_cv = s_cv0;
  if (id_p >= 0) s_p = bst[id_p * G];
  if (id_v >= 0) s_v = bst[id_v * G];
  if (nfail) s_fail = bst[id_fail * G];
  if (ninv) s_inv = bst[id_inv * G];
  if (swm_peak) s_swm = bst[id_swm * G];
  const uint32_t t0 = ag1.z;
  const int r0 = (int)(my_loc >> 16), c0 = (int)(my_loc & 0xffffu);
  const int cell = r0 * d.W + c0;
  uint32_t* cvp = cvrow + (isA ? (cell >> 5) : 0);
  uint32_t cvw = 0;
  if (isA) cvw = *cvp;
  float tw = __uint_as_float(bh0.w), tf = __uint_as_float(bh1.x);

  // ---- token caches: a changed vibe edits the cached list in place (tag tokens, then the vibe token, then
  // inventory / group / agent id: core/grid_object.cpp:178-203); never-built caches are built from the record
  {
    const int nv = __shfl_sync(MG_FULL, new_vibe, o_agent >= 0 ? o_agent : 0, G);
    const uint32_t oldv = (o_meta >> 16) & 0xffu, newv = (uint32_t)nv & 0xffu;
    bool rebuild = o_alive && o_ntok == MG_TOK_DIRTY;
    if (o_alive && o_agent >= 0 && o_agent < A && nv >= 0 && newv != oldv) {
      o_meta = (o_meta & 0xff00ffffu) | (newv << 16);
      uint16_t* tk = (uint16_t*)(toks + gl * tokw);
      const int cap = hdr[MGH_TOK_CAP];
      if (!rebuild && !(oldv == 0 && (int)o_ntok >= cap)) {
        int n = (int)o_ntok, p = 0;
        const uint32_t ftag = (uint32_t)hdr[MGH_FEAT_TAG] & 0xffu;
#pragma unroll 1
        while (p < n && (tk[p] & 0xffu) == ftag) p++;
        const uint16_t vt = (uint16_t)(((uint32_t)hdr[MGH_FEAT_VIBE] & 0xffu) | (newv << 8));
        if (oldv != 0 && newv != 0) {
          tk[p] = vt;
        } else if (oldv == 0) {
#pragma unroll 1
          for (int i = n; i > p; i--) tk[i] = tk[i - 1];
          tk[p] = vt;
          n++;
        } else {
#pragma unroll 1
          for (int i = p; i + 1 < n; i++) tk[i] = tk[i + 1];
          n--;
        }
        o_ntok = (uint32_t)n;
        if (live) {
#pragma unroll 1
          for (int k = 4; k < (n + 1) >> 1; k++) o[TOKOFF + k] = toks[gl * tokw + k];
        }
      } else {
        rebuild = true;
      }
    }
    if (rebuild) {
      const TokDims td = {d.P, d.TW, d.ND, d.B, hdr[MGH_TOK_CAP], hdr[MGH_FEAT_TAG], hdr[MGH_FEAT_VIBE], hdr[MGH_FEAT_GROUP],
                          hdr[MGH_FEAT_AGENT_ID], hdr[MGS_INV_FEATS], hdr[MGS_TEMPLATES]};
      o_ntok = (uint32_t)build_tokens(td, o, o_meta, (uint16_t*)(toks + gl * tokw), E, live);
      if (live) {
        const int nwords = (int)(o_ntok + 1) >> 1;
#pragma unroll 1
        for (int k = 4; k < nwords; k++) o[TOKOFF + k] = toks[gl * tokw + k];
      }
    }
  }
  // the first eight cached tokens stay in registers for the emission below
  if (o_ntok != ntok_raw || o_meta != o_meta0) {
    const uint32_t* rowp = toks + gl * tokw;
    tw0 = rowp[0], tw1 = rowp[1], tw2 = rowp[2], tw3 = rowp[3];
  }
  const int my_n = o_alive ? (int)o_ntok : 0;
  ((uint2*)oloc)[gl] = make_uint2(o_alive ? o_loc : 0x7FFF7FFFu,  // a dead slot lies outside every window
                                  ((uint32_t)my_n << 8) | (uint32_t)gl);
  __syncwarp();

  // ---- observations (:665-824)
  uint8_t* row = stage + gl * 3 * T;
  int pos = 0;
  // static variant: token p of agent a's row -- staged, or written in place when the row outgrows its stage
  auto emit = [&](int a, int p, uint32_t loc, uint32_t x) {
    if (p < SC) {
      uint8_t* q = stage + a * RS + p * 3;
      q[0] = (uint8_t)loc, q[1] = (uint8_t)x, q[2] = (uint8_t)(x >> 8);
    } else if (p < T && live) {
      uint8_t* q = gobs + (size_t)a * (3 * T) + p * 3;
      q[0] = (uint8_t)loc, q[1] = (uint8_t)x, q[2] = (uint8_t)(x >> 8);
    }
  };
  auto put = [&](uint32_t loc, uint32_t feat, uint32_t val) {
    if (S) {
      emit(gl, pos, loc, (feat & 0xffu) | ((val & 0xffu) << 8));
    } else if (pos < T) {
      row[pos * 3 + 0] = (uint8_t)loc;
      row[pos * 3 + 1] = (uint8_t)feat;
      row[pos * 3 + 2] = (uint8_t)val;
    }
    pos++;
  };
  auto global_tokens = [&]() {  // :700-742
    if (!isA) return;
    const int flags = hdr[MGH_GLOBAL_FLAGS];
    if (flags & MGG_EPISODE_PCT) {
      const int ms = hdr[MGH_MAX_STEPS];
      uint32_t val = 0;
      if (ms > 0) val = step >= (uint32_t)ms ? 255u : ((256u * step / (uint32_t)ms) & 0xffu);
      put(0xFE, hdr[MGH_FEAT_EPISODE_PCT], val);
    }
    if (flags & MGG_LAST_ACTION) put(0xFE, hdr[MGH_FEAT_LAST_ACTION], (uint32_t)exec_idx & 0xffu);
    if ((flags & MGG_LAST_ACTION_MOVE) && hdr[MGH_FEAT_LAST_ACTION_MOVE] != 0)
      put(0xFE, hdr[MGH_FEAT_LAST_ACTION_MOVE], my_loc != my_loc0);
    if (flags & MGG_LAST_REWARD) put(0xFE, hdr[MGH_FEAT_LAST_REWARD], 0);  // rewards are zero at this point (:937-938)
    if (flags & MGG_LOCAL_POSITION) {
      const int de = c0 - (int)(a_spawn & 0xffffu), dn = (int)(a_spawn >> 16) - r0;
      if (de != 0) put(0xFE, de > 0 ? hdr[MGH_FEAT_LP_EAST] : hdr[MGH_FEAT_LP_WEST], min(abs(de), 255));
      if (dn != 0) put(0xFE, dn > 0 ? hdr[MGH_FEAT_LP_NORTH] : hdr[MGH_FEAT_LP_SOUTH], min(abs(dn), 255));
    }
  };
  uint32_t* tab = (uint32_t*)(gb + L.key_off);
  if constexpr (!S) {
    // sort keys for the objects in this agent's window:
    // key = Manhattan rank << 24 | packed offset << 16 | token count << 8 | object lane
    uint32_t key[G];
    uint32_t col = 0;  // which agents see this lane's object
    {
      const uint32_t bias = (((uint32_t)(hdr[MGH_OBS_H] >> 1) << 16) | (uint32_t)(hdr[MGH_OBS_W] >> 1)) - my_loc;
#pragma unroll
      for (int j = 0; j < G; j++) {
        // both offsets in one subtraction; a negative column borrows into the row field but then fails the test
        const uint2 oi = ((const uint2*)oloc)[j];
        const uint32_t df = oi.x + bias;
        const uint32_t kk = lut[((df >> 12) & 0xf0u) | (df & 0xfu)] | oi.y;
        const bool vis = isA && (df & 0xfff0fff0u) == 0 && kk < 0xff000000u;
        key[j] = vis ? kk : FAST_INVALID;
        const uint32_t b = __ballot_sync(MG_FULL, vis);
        if (gl == j) col = b;
      }
      col = (col & gmask) >> gshift;
    }
    // cell staleness (:787-796) goes to the lowest agent index among the observers
    if (o_alive && col != 0 && o_vis < step) {
      atomicAdd(&stale[__ffs(col) - 1], step - o_vis);
      o_vis = step;
    }
    sort_net(key);
    global_tokens();
    // window tokens in Manhattan order (:756-811).  The observer walks its sorted keys once to give every visible
    // object its first token position; the object lanes then write their own tokens into each observer's row.
    // tab: [observer][object] -> first token position | packed offset << 16
#pragma unroll
    for (int k = 0; k < G; k++) {
      const uint32_t kk = key[k];
      if (kk != FAST_INVALID) {
        tab[gl * G + (int)(kk & 0xffu)] = (uint32_t)pos | (kk & 0x00ff0000u);
        pos += (int)((kk >> 8) & 0xffu);
      }
    }
    __syncwarp();
    while (col) {
      const int a = __ffs(col) - 1;
      col &= col - 1;
      const uint32_t e = tab[a * G + gl];
      const int start = (int)(e & 0xffffu);
      const uint32_t loc = e >> 16;
      const int m = min(my_n, T - start);  // tokens that fit the budget
      uint8_t* p = stage + a * 3 * T + start * 3;
      if (m > 0) {
        p[0] = (uint8_t)loc, p[1] = (uint8_t)tw0, p[2] = (uint8_t)(tw0 >> 8);
        if (m > 1) p[3] = (uint8_t)loc, p[4] = (uint8_t)(tw0 >> 16), p[5] = (uint8_t)(tw0 >> 24);
      }
      if (m > 2) {
        p[6] = (uint8_t)loc, p[7] = (uint8_t)tw1, p[8] = (uint8_t)(tw1 >> 8);
        if (m > 3) p[9] = (uint8_t)loc, p[10] = (uint8_t)(tw1 >> 16), p[11] = (uint8_t)(tw1 >> 24);
      }
      if (m > 4) {
        p[12] = (uint8_t)loc, p[13] = (uint8_t)tw2, p[14] = (uint8_t)(tw2 >> 8);
        if (m > 5) p[15] = (uint8_t)loc, p[16] = (uint8_t)(tw2 >> 16), p[17] = (uint8_t)(tw2 >> 24);
      }
      if (m > 6) {
        p[18] = (uint8_t)loc, p[19] = (uint8_t)tw3, p[20] = (uint8_t)(tw3 >> 8);
        if (m > 7) p[21] = (uint8_t)loc, p[22] = (uint8_t)(tw3 >> 16), p[23] = (uint8_t)(tw3 >> 24);
      }
      if (m > 8) {  // long token lists continue from the shared table
        const uint16_t* tk = (const uint16_t*)(toks + gl * tokw);
#pragma unroll 1
        for (int t = 8; t < m; t++) {
          const uint32_t x = tk[t];
          p[3 * t] = (uint8_t)loc, p[3 * t + 1] = (uint8_t)x, p[3 * t + 2] = (uint8_t)(x >> 8);
        }
      }
    }
  } else {
    // Static-layer variant.  No sort: an object at packed window offset `loc` of observer a starts at
    //   (a's global tokens) + s_ntok * popcount(a's static window set & offsets earlier than loc)
    //                       + (tokens of the dynamic objects a sees at an earlier rank),
    // so every (object, observer) pair can be placed on its own.  The pairs -- of the agents' objects first, then of the
    // static objects -- go to a work list that all lanes drain together, however unevenly they are spread over windows.
    global_tokens();
    uint32_t* const wm = (uint32_t*)(gb + L.wm_off);   // [agent][9] static window set, 256 bits by packed offset
    uint16_t* const dl = (uint16_t*)(gb + L.dl_off);   // [k][agent] dynamic objects the agent sees: rank << 8 | tokens
    uint32_t* const agw = (uint32_t*)(gb + L.ag_off);  // [agent] bias that turns a cell into its packed window offset
    uint32_t* const ngw = agw + G;                     // [agent] global tokens
    uint32_t* const nvw = ngw + G;                     // [agent] entries in its dl column
    uint32_t* const wl_count = nvw + G;
    uint32_t* const nvb = wl_count + 4;                // [agent] valid bytes of its row
    uint32_t* const wl = tab;                          // work list of (object, observer) pairs, see place()
    const int s_ntok = (int)sb[MGFS_NTOK];
    const int OH = hdr[MGH_OBS_H], OW = hdr[MGH_OBS_W], rr = OH >> 1, cr = OW >> 1;
    int n_static = 0;  // static objects in this agent's window
    {
      uint32_t M[8];
      const int rtop = r0 - rr + d.PAD, cbit = c0 - cr + d.PAD;  // >= 0: the bitmap has a PAD-wide frame
      const int kw = cbit >> 5, sh = cbit & 31;
      const uint32_t wmask = (1u << OW) - 1u;
#pragma unroll
      for (int wr = 0; wr < 15; wr++) {  // a window row = OW bits of a bitmap row; rows sit 16 bits apart like packed offsets
        uint32_t bits = 0;
        if (wr < OH && isA) {
          const uint32_t* rw = sbm + (rtop + wr) * SBW + kw;
          bits = __funnelshift_r(rw[0], rw[1], sh) & wmask;
        }
        if (wr & 1)
          M[wr >> 1] |= bits << 16;
        else
          M[wr >> 1] = bits;
      }
      const uint4* shape = (const uint4*)(d.fast_less + 256 * 8);  // offsets inside the observation shape
      const uint4 sa = __ldg(shape), sc = __ldg(shape + 1);
      M[0] &= sa.x, M[1] &= sa.y, M[2] &= sa.z, M[3] &= sa.w, M[4] &= sc.x, M[5] &= sc.y, M[6] &= sc.z, M[7] &= sc.w;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        wm[gl * 9 + k] = M[k];
        n_static += __popc(M[k]);
      }
    }
    agw[gl] = (uint32_t)(((rr - r0) << 4) + (cr - c0));  // + (r << 4) + c = packed window offset of cell (r, c)
    ngw[gl] = (uint32_t)pos;
    nvw[gl] = 0;
    if (gl == 0) *wl_count = 0;
    if (isA) {  // which agents' windows cover each row and each column (the bounding box of the observation shape)
      const uint32_t me = 1u << gl;
#pragma unroll 1
      for (int r = max(r0 - rr, 0); r <= min(r0 + rr, d.H - 1); r++) atomicOr(&rowm[r], me);
#pragma unroll 1
      for (int c = max(c0 - cr, 0); c <= min(c0 + cr, d.W - 1); c++) atomicOr(&colm[c], me);
    }
    // the first round of static cells and stamps: in flight while the agents' objects are handled
    const int NS = (int)sb[MGFS_COUNT];
    uint32_t e_nxt = 0, v_nxt = 0;
    if (gl < NS) e_nxt = __ldg(slist + gl), v_nxt = svis[gl];
    __syncwarp();
    // window-table entry (rank << 24 | packed offset << 16) of cell rc = (r << 4) + c for an observer whose bounding box
    // holds it, or 0xFFFFFFFF outside the observation shape
    auto entry_for = [&](int a, int rc) {
      const uint32_t lk = lut[(uint32_t)(rc + (int)agw[a]) & 0xffu];
      return lk >= 0xff000000u ? FAST_INVALID : lk;
    };
    // one (object, observer) pair: where the object's tokens start in the observer's row, then the tokens
    const uint16_t* stok = (const uint16_t*)(sb + MGFS_TOKENS);
    auto place = [&](uint32_t e) {  // e = rank << 24 | offset << 16 | observer << 8 | source lane (0xFF: static)
      const uint32_t loc = (e >> 16) & 0xffu, rank = e >> 24;
      const int a = (int)((e >> 8) & 0xffu), src = (int)(e & 0xffu);
      const uint4* lp = (const uint4*)(d.fast_less + loc * 8);  // offsets earlier in Manhattan order
      const uint4 la4 = __ldg(lp), lb4 = __ldg(lp + 1);
      const uint32_t* w = wm + a * 9;
      int before = __popc(w[0] & la4.x) + __popc(w[1] & la4.y) + __popc(w[2] & la4.z) + __popc(w[3] & la4.w) +
                   __popc(w[4] & lb4.x) + __popc(w[5] & lb4.y);
      if (OH > 12) before += __popc(w[6] & lb4.z) + __popc(w[7] & lb4.w);  // window rows 12-14
      int p = (int)ngw[a] + s_ntok * before;
      const int nva = (int)nvw[a];
#pragma unroll 1
      for (int k = 0; k < nva; k++) {
        const uint32_t kk = dl[k * G + a];
        if ((kk >> 8) < rank) p += (int)(kk & 0xffu);
      }
      const uint16_t* tk = stok;
      int nt = s_ntok;
      if (src != 0xff) tk = (const uint16_t*)(toks + src * tokw), nt = (int)(((const uint2*)oloc)[src].y >> 8);
#pragma unroll 1
      for (int t = 0; t < nt; t++) emit(a, p + t, loc, tk[t]);
    };
    // append to the work list; a full list places the pair at once instead
    auto push = [&](uint32_t e) {
      // (one atomic per lane: aggregating the lanes that arrive together -- ballot, leader, shuffle -- measured slower,
      // 59 vs 55 us per tick on the toy preset)
      const uint32_t slot = atomicAdd(wl_count, 1u);
      if (slot < (uint32_t)L.wl_cap)
        wl[slot] = e;
      else
        place(e);
    };
    // the agents' objects: observers, `visited` (:787-796) for the lowest one
    uint32_t seen = 0;
    if (o_alive) {
      const int r = (int)(o_loc >> 16), c = (int)(o_loc & 0xffffu);
      uint32_t cand = rowm[r] & colm[c];
      while (cand) {
        const int a = __ffs(cand) - 1;
        cand &= cand - 1;
        const uint32_t lk = entry_for(a, (r << 4) + c);
        if (lk == FAST_INVALID) continue;  // inside the bounding box but outside the shape
        seen |= 1u << a;
        dl[atomicAdd(&nvw[a], 1u) * G + a] = (uint16_t)(((lk >> 24) << 8) | (uint32_t)my_n);  // rank << 8 | tokens
      }
      if (seen && o_vis < step) {
        atomicAdd(&stale[__ffs(seen) - 1], step - o_vis);
        o_vis = step;
      }
    }
    __syncwarp();  // every agent's list of dynamic objects is complete
    if (isA) {     // every token this row attempts (:640-661)
      const int nva = (int)nvw[gl];
      pos += s_ntok * n_static;
#pragma unroll 1
      for (int k = 0; k < nva; k++) pos += (int)(dl[k * G + gl] & 0xffu);
    }
    nvb[gl] = (uint32_t)(3 * min(pos, T));  // valid bytes of the row, for the stream-out below
    if (o_alive) {
      const int r = (int)(o_loc >> 16), c = (int)(o_loc & 0xffffu);
      while (seen) {
        const int a = __ffs(seen) - 1;
        seen &= seen - 1;
        push(entry_for(a, (r << 4) + c) | ((uint32_t)a << 8) | (uint32_t)gl);
      }
    }
    // the static objects, G at a time: observers from the row / column masks, `visited` for the lowest one, entries
#pragma unroll 1
    for (int base = 0;; base += G) {
      __syncwarp(gmask);
      const int cnt = min((int)*wl_count, L.wl_cap);
      const bool last = base >= NS;
      __syncwarp(gmask);  // everyone has read the count before anyone appends again
      if (last || cnt > L.wl_cap - 2 * G) {  // drain the list: one pair per lane and turn
#pragma unroll 1
        for (int i = gl; i < cnt; i += G) place(wl[i]);
        __syncwarp(gmask);
        if (gl == 0) *wl_count = 0;
        __syncwarp(gmask);
      }
      if (last) break;
      const int idx = base + gl;
      const uint32_t e = e_nxt, v = v_nxt;
      if (idx + G < NS) e_nxt = __ldg(slist + idx + G), v_nxt = svis[idx + G];
      if (idx < NS) {
        const int r = (int)(e >> 16), c = (int)(e & 0xffffu), rc = (r << 4) + c;
        uint32_t cand = rowm[r] & colm[c];
        bool first = true;
        while (cand) {
          const int a = __ffs(cand) - 1;
          cand &= cand - 1;
          const uint32_t lk = entry_for(a, rc);
          if (lk == FAST_INVALID) continue;
          if (first) {
            first = false;
            if (v < step) {
              atomicAdd(&stale[a], step - v);
              if (live) svis[idx] = step;
            }
          }
          push(lk | ((uint32_t)a << 8) | 0xffu);
        }
      }
    }
  }
  const int attempted = pos;
  __syncwarp();

  // ---- stream the env's observation block out.  The block [A][T][3] is contiguous in the stage and in HBM: when both
  // ends are 16-byte aligned one elected lane hands the whole block to the bulk-copy engine (cp.async.bulk, shared ->
  // global; UBLKCP in the SASS) and the group goes on with its write-back while the copy drains; the wait sits at the
  // end of the kernel.  Unaligned blocks (odd T * A) take the vector loop.
  const bool bulk = !S && (((uint32_t)(uintptr_t)gobs | (uint32_t)nbytes) & 15u) == 0 && !MG_FAST_NO_BULK;
  const bool sbulk = S && (((uint32_t)(uintptr_t)gobs | (uint32_t)(3 * T)) & 7u) == 0;  // rows of whole 8-byte units
  if (S) {
    // Static variant: a row is its staged prefix, the tail it wrote in place (rare) and EmptyTokenByte padding
    // (:940-942).  Rows are whole 8-byte units when 3T is a multiple of 8.  Lane a stores row a's staged units; the
    // padding -- nine tenths of the block -- is handed to the bulk-copy engine row by row (cp.async.bulk from a constant
    // 0xFF buffer, 16-byte aligned on both ends; the odd 8-byte unit at either end is a plain store).
    const uint32_t* nvb = (const uint32_t*)(gb + L.ag_off) + 3 * G + 4;
    const int RB = 3 * T;
    if (sbulk) {
      if (isA && live) {
        uint8_t* grow = gobs + (size_t)gl * RB;
        const uint8_t* srow = stage + gl * RS;
        const int nv = (int)nvb[gl];
        const int fd = min(nv, RS) >> 3;
#pragma unroll 1
        for (int k = 0; k < fd; k++) *(uint2*)(grow + 8 * k) = *(const uint2*)(srow + 8 * k);
        if (nv & 7) {  // the unit where the row's tokens end
          const int kb = nv >> 3, n = nv & 7;
          if (nv <= RS) {
            uint2 h = *(const uint2*)(srow + 8 * kb);
            if (n < 4)
              h.x |= 0xffffffffu << (8 * n), h.y = 0xffffffffu;
            else
              h.y |= 0xffffffffu << (8 * (n - 4));
            *(uint2*)(grow + 8 * kb) = h;
          } else {
#pragma unroll 1
            for (int b = n; b < 8; b++) grow[8 * kb + b] = 0xff;
          }
        }
        int p0 = (nv + 7) & ~7, end = RB;
        const uint2 ff = make_uint2(0xffffffffu, 0xffffffffu);
        if (p0 < end && ((uint32_t)(uintptr_t)(grow + p0) & 8u)) *(uint2*)(grow + p0) = ff, p0 += 8;
        if (p0 < end && ((uint32_t)(uintptr_t)(grow + end) & 8u)) end -= 8, *(uint2*)(grow + end) = ff;
        if (end > p0) {  // every row's lane queues its own copy
          const uint32_t fsrc = (uint32_t)__cvta_generic_to_shared(smem + 1024);  // the CTA's 0xFF buffer
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(grow + p0), "r"(fsrc), "r"(end - p0) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else if (live) {  // odd shapes: byte by byte; bytes a row wrote in place are left alone
#pragma unroll 1
      for (int i = gl; i < nbytes; i += G) {
        const int a = i / RB, off = i - a * RB;
        if (off >= (int)nvb[a])
          gobs[i] = 0xff;
        else if (off < RS)
          gobs[i] = stage[a * RS + off];
      }
    }
  } else if (bulk) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the stage was written through the generic proxy
    __syncwarp();
    if (gl == 0 && live) {
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(stage);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gobs), "r"(saddr), "r"(nbytes) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  } else {
    const int head = min(nbytes, (int)((16u - ((uint32_t)(uintptr_t)gobs & 15u)) & 15u));
    if (live) {
#pragma unroll 1
      for (int i = gl; i < head; i += G) gobs[i] = stage[i];
    }
    const int body = (nbytes - head) >> 4;
    const uint4* s4 = (const uint4*)(stage + head);
    uint4* g4 = (uint4*)(gobs + head);
    if (live) {
      int v = gl;
#pragma unroll 1
      for (; v + 3 * G < body; v += 4 * G) {  // four vectors in flight per lane
        const uint4 x0 = s4[v], x1 = s4[v + G], x2 = s4[v + 2 * G], x3 = s4[v + 3 * G];
        __stcs(g4 + v, x0), __stcs(g4 + v + G, x1), __stcs(g4 + v + 2 * G, x2), __stcs(g4 + v + 3 * G, x3);
      }
#pragma unroll 1
      for (; v < body; v += G) __stcs(g4 + v, s4[v]);
    }
    const int done = head + (body << 4);
    if (live) {
#pragma unroll 1
      for (int i = done + gl; i < nbytes; i += G) gobs[i] = stage[i];
    }
  }

  // ---- token stats (:659-661, :640-642) and the token budget (:364-375)
  {
    const bool over = isA && attempted > T;
    const uint32_t ov = (__ballot_sync(MG_FULL, over) & gmask) >> gshift;
    int sw = (isA && !over) ? attempted : 0, sf = (isA && !over) ? T - attempted : 0;
#pragma unroll
    for (int off = G / 2; off >= 1; off >>= 1) {
      sw += __shfl_xor_sync(MG_FULL, sw, off, G);
      sf += __shfl_xor_sync(MG_FULL, sf, off, G);
    }
    // The reference adds one float per agent in agent order.  While the totals stay below 2^24 every such add
    // is exact, so the sum can be added at once; beyond that the adds are replayed in order.
    const bool slow = gl == 0 && !(tw + (float)sw < 16777216.0f && tf + (float)sf < 16777216.0f);
    if (__any_sync(MG_FULL, slow)) {
#pragma unroll 1
      for (int a = 0; a < A; a++) {
        const int at = __shfl_sync(MG_FULL, attempted, a, G);
        if (slow && at <= T) {
          tw = __fadd_rn(tw, (float)at);
          tf = __fadd_rn(tf, (float)(T - at));
        }
      }
    }
    if (gl == 0 && !slow) {
      tw = __fadd_rn(tw, (float)sw);
      tf = __fadd_rn(tf, (float)sf);
    }
    if (gl == 0 && live) {  // tokens_written, tokens_free_space, and the touched flags of the three token stats
      blk[MGFB_TOKENS_WRITTEN] = __float_as_uint(tw), blk[MGFB_TOKENS_FREE] = __float_as_uint(tf);
      if ((bh1.y & 7u) != 7u) blk[MGFB_GTOUCHED] = bh1.y | 7u;
    }
    if (__any_sync(MG_FULL, ov != 0)) {  // hard error in the reference; the first agent in index order is reported
      const int a = ov ? __ffs(ov) - 1 : 0;
      const int at = __shfl_sync(MG_FULL, attempted, a, G);
      if (gl == 0 && ov && live) set_error(E, MGERR_TOKEN_OVERFLOW, a | (min(at, 65535) << 16));
    }
  }

  // ---- per-agent write-back: stats, coverage (objects/agent.cpp:49-57), flags
  __syncwarp();
  if (isA && live) {
    uint32_t touched = 0;  // well-known stat ids are below 16 (checked by mg_create)
    const float fpass = (float)npass;
    if (ninv >= 1) s_inv = __fadd_rn(s_inv, fpass);  // once per priority pass
    if (ninv >= 2) s_inv = __fadd_rn(s_inv, fpass);
    if (ninv) bst[id_inv * G] = s_inv, touched |= 1u << id_inv;
    if (swm_peak && (float)swm_peak > s_swm) bst[id_swm * G] = (float)swm_peak, touched |= 1u << id_swm;
    if (id_p >= 0) bst[id_p * G] = __fadd_rn(s_p, 1.0f), touched |= 1u << id_p;
    if (id_v >= 0) bst[id_v * G] = __fadd_rn(s_v, 1.0f), touched |= 1u << id_v;
    if (nfail >= 1) s_fail = __fadd_rn(s_fail, 1.0f);
    if (nfail >= 2) s_fail = __fadd_rn(s_fail, 1.0f);
    if (nfail) bst[id_fail * G] = s_fail, touched |= 1u << id_fail;
    const uint32_t bit = 1u << (cell & 31);
    if (!(cvw & bit)) {
      *cvp = cvw | bit;
      ++a_unique;
    }
    bst[id_un * G] = (float)a_unique;
    const uint32_t dist = (uint32_t)(abs((int)(a_spawn >> 16) - r0) + abs(c0 - (int)(a_spawn & 0xffffu)));
    const uint32_t md = max(a_maxd, dist);
    bst[id_md * G] = (float)md;
    touched |= (1u << id_un) | (1u << id_md);
    const uint32_t ssum = stale[gl];
    if (ssum) bst[id_cv * G] = __fadd_rn(s_cv, (float)ssum), touched |= 1u << id_cv;
    const uint32_t nt0 = t0 | touched;
    if (prev != a_prev || swm != a_swm) *(uint2*)(bag + 2) = make_uint2(prev, swm);
    if (md != a_maxd || a_unique != ag1.y || nt0 != t0) *(uint4*)(bag + 4) = make_uint4(md, a_unique, nt0, 0);
    bool success = ok_p;  // same update order as handle_action's caller: primary stream, then vibe stream
    if (inv_v) success = false;
    if (ok_v) success = true;
    d.success[g0 + gl] = (uint8_t)success;
    d.rewards[g0 + gl] = 0.0f;  // no reward entries in a plain program (systems/reward.hpp:56-77)
    const int ms = hdr[MGH_MAX_STEPS];
    if (ms > 0 && step >= (uint32_t)ms) {  // :1090-1096
      if (hdr[MGH_EPISODE_TRUNCATES])
        d.truncations[g0 + gl] = 1;
      else
        d.terminals[g0 + gl] = 1;
    }
  }
  // ---- object write-back
  if (o_alive && live) {
    const uint32_t w3 = (orec.w & 0xffu) | ((o_ntok & 0xffu) << 8);
    if (o_loc != o_loc0 || o_vis != o_vis0 || o_meta != o_meta0 || w3 != orec.w) *(uint4*)bob = make_uint4(o_loc, o_vis, o_meta, w3);
    if (o_ntok != ntok_raw || o_meta != o_meta0) *(uint4*)(bob + 4) = make_uint4(tw0, tw1, tw2, tw3);
  }
  if (gl == 0 && live) blk[MGFB_STEP] = step;
  if (((bulk && gl == 0) || (sbulk && isA)) && live) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the stage must outlive the copy
}


// ---- generic arrays <-> packed hot state (layout in mg_state.h); one thread per (env, lane) --------------
__global__ void k_fast_pack(const MgDev d, const MgFastHdr HD, const int G, const int S, const uint8_t* __restrict__ mask) {
  const int* const hdr = HD.v;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int env = (int)(t / G), gl = (int)(t % G);
  if (env >= d.num_envs || (mask && !mask[env])) return;
  uint32_t* blk = d.fast_blk + (size_t)env * d.fast_stride;
  const int32_t* E = d.env + (size_t)env * MGEV_WORDS;
  const int nobj = S ? d.A : E[MGEV_NEXT_OBJ] - 1;  // static variant: the lanes hold the agents' objects
  if (gl == 0) {
    const float* gs = d.gstats + (size_t)env * d.SG;
    const uint32_t* gt = d.gtouched + (size_t)env * d.SGW;
    const int idw = hdr[MGH_GST_TOKENS_WRITTEN], idd = hdr[MGH_GST_TOKENS_DROPPED], idf = hdr[MGH_GST_TOKENS_FREE];
    blk[MGFB_STEP] = (uint32_t)E[MGEV_STEP];
    blk[MGFB_RNG_IDX] = (uint32_t)E[MGEV_RNG_IDX];
    blk[MGFB_NOBJ] = (uint32_t)nobj;
    blk[MGFB_TOKENS_WRITTEN] = __float_as_uint(gs[idw]);
    blk[MGFB_TOKENS_FREE] = __float_as_uint(gs[idf]);
    blk[MGFB_GTOUCHED] = ((gt[idw >> 5] >> (idw & 31)) & 1u) | (((gt[idd >> 5] >> (idd & 31)) & 1u) << 1) |
                         (((gt[idf >> 5] >> (idf & 31)) & 1u) << 2);
    blk[6] = blk[7] = 0;
  }
  uint4 a0 = make_uint4(1, 0, 0, 0), a1 = make_uint4(0, 0, 0, 0);
  if (gl < d.A) {
    const size_t ga = (size_t)env * d.A + gl;
    const uint32_t* ag = d.agents + ga * d.AS;
    a0 = make_uint4(ag[MGAG_OBJ], ag[MGAG_SPAWN], ag[MGAG_PREV_LOC], ag[MGAG_SWM]);
    a1 = make_uint4(ag[MGAG_MAX_DIST], ag[MGAG_UNIQUE], d.atouched[ga * d.SAW] & 0xffffu, 0);
    for (int id = 0; id < MGFB_STATS; id++) blk[MGFB_STAT(G, id, gl)] = id < d.SA ? __float_as_uint(d.astats[ga * d.SA + id]) : 0u;
  } else {
    for (int id = 0; id < MGFB_STATS; id++) blk[MGFB_STAT(G, id, gl)] = 0u;
  }
  *(uint4*)(blk + MGFB_AGENT(G, gl)) = a0;
  *(uint4*)(blk + MGFB_AGENT(G, gl) + 4) = a1;
  uint4 o0 = make_uint4(0, 0, 0, 0), o1 = make_uint4(0, 0, 0, 0);
  const int slot = S ? (int)a0.x : gl + 1;
  if (gl < nobj && slot < d.maxobj) {
    const uint32_t* o = d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + (size_t)slot) * d.OS;
    const int TOKOFF = MG_TOKOFF(d.TW, d.R);
    const uint32_t ntok = o[MGO_NTOK];
    const uint32_t w3 = (((uint32_t)((int)o[MGO_AGENT] + 1)) & 0xffu) | (ntok == MG_TOK_DIRTY ? MGFB_DIRTY : ((ntok & 0xffu) << 8));
    o0 = make_uint4(o[MGO_LOC], d.visited[(size_t)env * d.maxobj + slot], o[MGO_META], w3);  // the generic kernels keep the stamps apart
    const int nw = min(4, (hdr[MGH_TOK_CAP] + 1) / 2);
    uint32_t tk[4] = {0, 0, 0, 0};
    for (int k = 0; k < nw; k++) tk[k] = o[TOKOFF + k];
    o1 = make_uint4(tk[0], tk[1], tk[2], tk[3]);
  }
  *(uint4*)(blk + MGFB_OBJECT(G, gl)) = o0;
  *(uint4*)(blk + MGFB_OBJECT(G, gl) + 4) = o1;
}

// Static layer of an env (k_step_fast<G, true>): bitmap, cell list, slots and stamps of the non-agent objects, in slot
// order; one warp per env.  All of them share one template (mg_capi.cu checks the maps), so one token list serves all.
__global__ void k_fast_pack_static(const MgDev d, const MgFastHdr HD, const uint8_t* __restrict__ mask) {
  const int* const hdr = HD.v;
  const int env = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (env >= d.num_envs || (mask && !mask[env])) return;
  int32_t* E = d.env + (size_t)env * MGEV_WORDS;
  uint32_t* sblk = (uint32_t*)d.fast_static + (size_t)env * d.fast_sstride;
  uint32_t* slist = (uint32_t*)d.fast_slist + (size_t)env * d.fast_ns_cap;
  uint16_t* sslot = (uint16_t*)d.fast_sslot + (size_t)env * d.fast_ns_cap;
  uint32_t* svis = d.fast_svis + (size_t)env * d.fast_ns_cap;
  for (int i = lane; i < d.fast_sstride; i += 32) sblk[i] = 0;
  __syncwarp();
  const int nall = min(E[MGEV_NEXT_OBJ] - 1, d.maxobj - 1);
  uint32_t* bm = sblk + MGFS_HDR_WORDS;
  int count = 0, first_slot = 0;
  uint32_t first_meta = 0;
  bool mixed = false;
  for (int base = 1; base <= nall; base += 32) {
    const int slot = base + lane;
    uint32_t loc = 0, meta = 0;
    bool st = false;
    if (slot <= nall) {
      const uint32_t* o = d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + (size_t)slot) * d.OS;
      loc = o[MGO_LOC], meta = o[MGO_META];
      st = ((meta >> 24) & MGOF_ALIVE) && !((meta >> 24) & MGOF_AGENT);
    }
    const uint32_t b = __ballot_sync(MG_FULL, st);
    if (b && !first_slot) {
      const int src = __ffs(b) - 1;
      first_slot = base + src;
      first_meta = __shfl_sync(MG_FULL, meta, src);
    }
    if (st && (meta & 0x00ffffffu) != (first_meta & 0x00ffffffu)) mixed = true;  // template and vibe
    const int idx = count + __popc(b & ((1u << lane) - 1u));
    if (st && idx < d.fast_ns_cap) {
      const int r = (int)(loc >> 16), c = (int)(loc & 0xffffu), cb = c + d.PAD;
      slist[idx] = loc, sslot[idx] = (uint16_t)slot, svis[idx] = d.visited[(size_t)env * d.maxobj + slot];
      atomicOr(&bm[(r + d.PAD) * d.fast_sbw + (cb >> 5)], 1u << (cb & 31));
    }
    count += __popc(b);
  }
  mixed = __any_sync(MG_FULL, mixed);
  if (lane == 0) {
    int n = 0;
    if (first_slot) {
      const TokDims td = {d.P, d.TW, d.ND, d.B, hdr[MGH_TOK_CAP], hdr[MGH_FEAT_TAG], hdr[MGH_FEAT_VIBE], hdr[MGH_FEAT_GROUP],
                          hdr[MGH_FEAT_AGENT_ID], hdr[MGS_INV_FEATS], hdr[MGS_TEMPLATES]};
      uint16_t tk[128];  // MGH_TOK_CAP <= 126 on this path
      const uint32_t* o = d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + (size_t)first_slot) * d.OS;
      n = build_tokens(td, o, first_meta, tk, E, true);
      for (int t = 0; t < n && t < MGFS_MAX_TOKENS; t++) ((uint16_t*)(sblk + MGFS_TOKENS))[t] = tk[t];
    }
    if (count > d.fast_ns_cap || n > MGFS_MAX_TOKENS || mixed) set_error(E, MGERR_POOL_EXHAUSTED, 20);
    sblk[MGFS_COUNT] = (uint32_t)min(count, d.fast_ns_cap);
    sblk[MGFS_NTOK] = (uint32_t)min(n, MGFS_MAX_TOKENS);
  }
}

__global__ void k_fast_unpack(const MgDev d, const MgFastHdr HD, const int G, const int S) {
  const int* const hdr = HD.v;
  const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int env = (int)(t / G), gl = (int)(t % G);
  if (env >= d.num_envs) return;
  const uint32_t* blk = d.fast_blk + (size_t)env * d.fast_stride;
  int32_t* E = d.env + (size_t)env * MGEV_WORDS;
  const int nobj = (int)blk[MGFB_NOBJ];
  if (gl == 0) {
    float* gs = d.gstats + (size_t)env * d.SG;
    uint32_t* gt = d.gtouched + (size_t)env * d.SGW;
    const int idw = hdr[MGH_GST_TOKENS_WRITTEN], idd = hdr[MGH_GST_TOKENS_DROPPED], idf = hdr[MGH_GST_TOKENS_FREE];
    E[MGEV_STEP] = (int32_t)blk[MGFB_STEP];
    E[MGEV_RNG_IDX] = (int32_t)blk[MGFB_RNG_IDX];
    gs[idw] = __uint_as_float(blk[MGFB_TOKENS_WRITTEN]);
    gs[idf] = __uint_as_float(blk[MGFB_TOKENS_FREE]);
    const uint32_t g = blk[MGFB_GTOUCHED];
    if (g & 1u) gt[idw >> 5] |= 1u << (idw & 31);
    if (g & 2u) gt[idd >> 5] |= 1u << (idd & 31);
    if (g & 4u) gt[idf >> 5] |= 1u << (idf & 31);
  }
  if (gl < d.A) {
    const size_t ga = (size_t)env * d.A + gl;
    uint32_t* ag = d.agents + ga * d.AS;
    const uint4 a0 = *(const uint4*)(blk + MGFB_AGENT(G, gl)), a1 = *(const uint4*)(blk + MGFB_AGENT(G, gl) + 4);
    ag[MGAG_PREV_LOC] = a0.z, ag[MGAG_SWM] = a0.w, ag[MGAG_MAX_DIST] = a1.x, ag[MGAG_UNIQUE] = a1.y;
    d.atouched[ga * d.SAW] = (d.atouched[ga * d.SAW] & ~0xffffu) | (a1.z & 0xffffu);  // the block owns stat ids < 16
    for (int id = 0; id < MGFB_STATS && id < d.SA; id++) d.astats[ga * d.SA + id] = __uint_as_float(blk[MGFB_STAT(G, id, gl)]);
  }
  const int slot = S ? (int)blk[MGFB_AGENT(G, gl < d.A ? gl : 0)] : gl + 1;  // static variant: lane l holds agent l's object
  if (gl < nobj && slot < d.maxobj) {
    uint32_t* o = d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + (size_t)slot) * d.OS;
    const int TOKOFF = MG_TOKOFF(d.TW, d.R);
    const uint4 o0 = *(const uint4*)(blk + MGFB_OBJECT(G, gl)), o1 = *(const uint4*)(blk + MGFB_OBJECT(G, gl) + 4);
    o[MGO_LOC] = o0.x, d.visited[(size_t)env * d.maxobj + slot] = o0.y, o[MGO_META] = o0.z;
    if (o0.w & MGFB_DIRTY) {
      o[MGO_NTOK] = MG_TOK_DIRTY;
    } else {
      o[MGO_NTOK] = (o0.w >> 8) & 0xffu;
      const int nw = min(4, (hdr[MGH_TOK_CAP] + 1) / 2);
      const uint32_t tk[4] = {o1.x, o1.y, o1.z, o1.w};
      for (int k = 0; k < nw; k++) o[TOKOFF + k] = tk[k];
    }
  }
  // k_step_fast never reads or writes the occupancy grid; rebuild it for the generic consumers (k_init_buffers after a
  // mid-episode mg_set_buffers, the generic step kernels): objects are never created or removed on this path.
  // A group's G threads share one warp (G divides 32).
  uint16_t* cells = d.cells + (size_t)env * d.HWp;
  for (int i = gl; i < d.HWp; i += G) cells[i] = 0;
  __syncwarp();
  if (gl < nobj && slot < d.maxobj) {
    const uint32_t loc = blk[MGFB_OBJECT(G, gl)];
    cells[((int)(loc >> 16) + d.PAD) * d.WP + (int)(loc & 0xffffu) + d.PAD] = (uint16_t)slot;
  }
  if (S) {  // the static objects: their cells and their stamps
    const int NS = (int)d.fast_static[(size_t)env * d.fast_sstride + MGFS_COUNT];
    const uint32_t* slist = d.fast_slist + (size_t)env * d.fast_ns_cap;
    const uint16_t* sslot = d.fast_sslot + (size_t)env * d.fast_ns_cap;
    const uint32_t* svis = d.fast_svis + (size_t)env * d.fast_ns_cap;
    for (int i = gl; i < NS; i += G) {
      const uint32_t loc = slist[i];
      cells[((int)(loc >> 16) + d.PAD) * d.WP + (int)(loc & 0xffffu) + d.PAD] = sslot[i];
      d.visited[(size_t)env * d.maxobj + sslot[i]] = svis[i];
    }
  }
}


// Rebuild selected environments of a fast handle from the snapshot of the packed block taken right after a full
// reset: the same state k_reset + k_init_buffers + k_fast_pack would produce, for the price of a copy.  Used by the
// vectorised-env step, where a tick follows at once (the initial observations are not written).
__global__ void k_fast_restore(const MgDev d, const int G, const uint32_t* __restrict__ snapshot, const uint8_t* __restrict__ mask) {
  const int env = blockIdx.x;
  if (!mask[env]) return;
  const int t = threadIdx.x, nt = blockDim.x;
  uint32_t* blk = d.fast_blk + (size_t)env * d.fast_stride;
  const uint32_t* snap = snapshot + (size_t)env * d.fast_stride;
  for (int i = t; i < d.fast_stride; i += nt) blk[i] = snap[i];
  // the freshly seeded generator (k_reset keeps it per env); the block's header already says "index 624"
  const uint32_t* seeded = d.rng_seeded + (size_t)env * (MG_RNG_WORDS + 2);
  uint32_t* rng = d.rng + (size_t)env * MG_RNG_WORDS;
  for (int i = t; i < MG_RNG_WORDS; i += nt) rng[i] = seeded[i];
  // coverage (objects/agent.cpp:41-47): only the spawn cell is visited
  uint32_t* cover = d.cover + (size_t)env * d.A * d.CW;
  for (int i = t; i < d.A * d.CW; i += nt) {
    const int a = i / d.CW, wd = i - a * d.CW;
    const uint32_t spawn = snap[MGFB_AGENT(G, a) + 1];
    const int cell = (int)(spawn >> 16) * d.W + (int)(spawn & 0xffffu);
    cover[i] = (cell >> 5) == wd ? 1u << (cell & 31) : 0u;
  }
  if (d.fast_svis)  // static variant: nothing has been observed yet
    for (int i = t; i < d.fast_ns_cap; i += nt) d.fast_svis[(size_t)env * d.fast_ns_cap + i] = 0;
  for (int a = t; a < d.A; a += nt) {  // _init_buffers (:294-319)
    const size_t gi = (size_t)env * d.A + a;
    d.terminals[gi] = 0, d.truncations[gi] = 0, d.rewards[gi] = 0.0f, d.success[gi] = 0;
  }
  if (t == 0) {
    int32_t* E = d.env + (size_t)env * MGEV_WORDS;
    E[MGEV_ERROR] = 0, E[MGEV_ERR_INFO] = 0;
  }
}

}  // namespace

// ---- host side -------------------------------------------------------------------------------------
static inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

MgFastLayout mg_fast_layout(const MgDev& d, int G, int tok_cap, int statics) {
  MgFastLayout L{};
  L.G = G;
  L.statics = statics;
  L.rank_off = 0;
  L.cta_bytes = 1024 + (statics ? (int)al16((size_t)3 * d.T) : 0);  // the window table; static variant: + a row of 0xFF
  size_t n = al16((size_t)d.A * 3 * d.T + 16) + 16;  // stage (+ phase slack)
  if (statics) {  // only a prefix of every row is staged
    L.stage_tokens = d.T < 64 ? (d.T + 15) & ~15 : 64;
    n = (size_t)d.A * 3 * L.stage_tokens;  // a multiple of 16
  }
  L.tok_stride = ((tok_cap + 1) / 2) | 1;               // odd word stride: conflict-free columns
  if (L.tok_stride < 5) L.tok_stride = 5;               // the first four words are always staged
  L.tok_off = (int)n;
  n += al16((size_t)G * L.tok_stride * 4);
  L.oloc_off = (int)n;
  n += al16((size_t)G * 4 * 4 + G);  // oloc, ontok, stale, draw, order
  L.key_off = (int)n;
  L.wl_cap = MG_FS_WL * G;
  n += statics ? (size_t)L.wl_cap * 4 : (size_t)G * G * 4;  // first token positions [observer][object] / the work list
  if (statics) {
    L.sb_words = (d.fast_sstride + 3) & ~3;
    L.sb_off = (int)n;
    n += (size_t)L.sb_words * 4;
    L.rc_off = (int)n;
    n += al16((size_t)(d.H + d.W) * 4);
    L.wm_off = (int)n;
    n += al16((size_t)G * 9 * 4);
    L.dl_off = (int)n;
    n += al16((size_t)d.A * G * 2);  // an observer sees at most the A agents' objects
    L.ag_off = (int)n;
    n += (size_t)(G * 4 + 4) * 4;  // location, global tokens, list length, valid bytes per agent + the work-list counter
  }
  L.group_bytes = (int)al16(n);
  // warps per CTA: one while the grid is a few waves at most, four beyond (measured; see k_step_fast)
  const long long warps = ((long long)d.num_envs + 32 / G - 1) / (32 / G);
  L.warps = MG_FAST_WARPS ? MG_FAST_WARPS : (!statics || warps <= 8192) ? 1 : 4;
  L.smem_bytes = L.cta_bytes + (size_t)L.warps * (32 / G) * L.group_bytes;
  return L;
}

template <int G, bool S, int WPC>
static cudaError_t launch_fast(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, cudaStream_t st) {
  const int envs_per_cta = WPC * (32 / G);
  const int grid = (d.num_envs + envs_per_cta - 1) / envs_per_cta;
  k_step_fast<G, S, WPC><<<grid, WPC * 32, L.smem_bytes, st>>>(d, L, H);
  return cudaGetLastError();
}

template <int G, bool S, int WPC>
static cudaError_t configure_one(int bytes) {
  cudaError_t e = cudaFuncSetAttribute(k_step_fast<G, S, WPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  // prefer shared memory: the working set lives there, L1 only serves the program tables
  return cudaFuncSetAttribute(k_step_fast<G, S, WPC>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// the instantiation for (lanes per env, static layer, warps per CTA)
#define MG_FAST_DISPATCH(fn, L, ...)                                                             \
  do {                                                                                           \
    const int key_ = (L.G == 8 ? 0 : L.G == 16 ? 1 : 2) * 4 + (L.statics ? 2 : 0) + (L.warps == 1 ? 0 : 1); \
    switch (key_) {                                                                              \
      case 0: return fn<8, false, 1>(__VA_ARGS__);                                               \
      case 1: return fn<8, false, 4>(__VA_ARGS__);                                               \
      case 2: return fn<8, true, 1>(__VA_ARGS__);                                                \
      case 3: return fn<8, true, 4>(__VA_ARGS__);                                                \
      case 4: return fn<16, false, 1>(__VA_ARGS__);                                              \
      case 5: return fn<16, false, 4>(__VA_ARGS__);                                              \
      case 6: return fn<16, true, 1>(__VA_ARGS__);                                               \
      case 7: return fn<16, true, 4>(__VA_ARGS__);                                               \
      case 8: return fn<32, false, 1>(__VA_ARGS__);                                              \
      case 9: return fn<32, false, 4>(__VA_ARGS__);                                              \
      case 10: return fn<32, true, 1>(__VA_ARGS__);                                              \
      default: return fn<32, true, 4>(__VA_ARGS__);                                              \
    }                                                                                            \
  } while (0)

cudaError_t mg_fast_configure(const MgFastLayout& L) { MG_FAST_DISPATCH(configure_one, L, (int)L.smem_bytes); }

cudaError_t mg_launch_step_fast(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, cudaStream_t st) {
  MG_FAST_DISPATCH(launch_fast, L, d, L, H, st);
}

cudaError_t mg_launch_fast_pack(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, const uint8_t* mask, cudaStream_t st) {
  const size_t threads = (size_t)d.num_envs * L.G;
  k_fast_pack<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d, H, L.G, L.statics, mask);
  if (L.statics) k_fast_pack_static<<<(d.num_envs + 3) / 4, 128, 0, st>>>(d, H, mask);
  return cudaGetLastError();
}
cudaError_t mg_launch_fast_unpack(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, cudaStream_t st) {
  const size_t threads = (size_t)d.num_envs * L.G;
  k_fast_unpack<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(d, H, L.G, L.statics);
  return cudaGetLastError();
}

cudaError_t mg_launch_fast_restore(const MgDev& d, const MgFastLayout& L, const uint32_t* snapshot, const uint8_t* mask,
                                   cudaStream_t st) {
  k_fast_restore<<<d.num_envs, 128, 0, st>>>(d, L.G, snapshot, mask);
  return cudaGetLastError();
}
