// mg_kernels.cu -- sm_100a kernels for the batched MettaGrid tick.
//
//   k_reset        : build per-env state from the template map (the reference's constructor,
//                    bindings/mettagrid_c.cpp:42-191,200-269)
//   k_init_buffers : _init_buffers (:294-319): clear flags, reset coverage (+ k_observe / k_finish: initial observations)
//   one tick, MettaGrid::_step (:921-1102), is three launches on one stream:
//   k_world        : phases 1-12 -- actions, events, on_tick, AOE, territory, coverage.  One warp = one environment;
//                    lane = agent wherever the compiler's effect analysis allows (action passes, on_tick, AOE,
//                    territory, coverage), lane 0 only for what must stay in the reference's order.
//   k_observe      : phase 13, the observation pass.  One warp = one AGENT: rows are independent once the cell
//                    staleness owner is decided analytically; each row is composed in shared memory and streamed
//                    to HBM with 16-byte stores.  Its register budget is its own (the interpreter wants 64 registers
//                    and many resident warps, the window loop wants its loads in flight together).
//   k_finish       : phases 14-15 -- token stats in agent order, rewards, episode rewards, truncation.
#include "mg_device.cuh"
#include "mg_world.cuh"

#ifdef MG_PHASE_TIMING
// A/B build only (python -m mettagrid_b200.build --variant prof -DMG_PHASE_TIMING): lane 0 of every warp adds the cycles
// it spent in each phase of k_world; tools/phase_timing.py reads the totals through mg_debug_phase_cycles.
__device__ unsigned long long g_phase_cycles[16];
#define MG_PHASE(k)                                        \
  do {                                                     \
    __syncwarp();                                          \
    if (lane == 0) {                                       \
      const long long now_ = clock64();                    \
      atomicAdd(&g_phase_cycles[k], (unsigned long long)(now_ - ph_t0)); \
      ph_t0 = now_;                                        \
    }                                                      \
  } while (0)
#define MG_PHASE_BEGIN() long long ph_t0 = clock64()
extern "C" int mg_debug_phase_cycles(unsigned long long* out, int reset) {
  if (cudaMemcpyFromSymbol(out, g_phase_cycles, sizeof g_phase_cycles) != cudaSuccess) return -1;
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_phase_cycles, z, sizeof z);
  }
  return 0;
}
#else
#define MG_PHASE(k) do { } while (0)
#define MG_PHASE_BEGIN() do { } while (0)
#endif

namespace {

struct Smem {
  int32_t* hdr;
  // per warp
  uint16_t* cells;
  int* rs;
  uint32_t* rand;
  uint32_t* a_slot;   // agent -> object slot
  uint32_t* a_loc;    // location after the action phase (observer position)
  uint32_t* a_step;   // location at the start of the tick (_prev_agent_locations)
  int32_t* a_act;     // primary-stream action index
  int32_t* a_vact;    // vibe-stream action index
  uint32_t* a_res;    // MGR_* bits recorded by the serial action loop
  uint32_t* a_locp;   // location right after the primary-stream action
  uint32_t* a_locv;   // location right after the vibe-stream action
  int32_t* a_exec;    // executed action (last successful)
  uint16_t* a_order;  // shuffled agent order
  uint16_t* a_pos;    // agent -> position in the shuffled order
  uint32_t* fp;       // action footprints, MG_FP_WORDS per agent lane (action_pass)
  Wv* wv;             // this warp's env view, kept in shared memory
  Smem* self;         // shared-memory copy of this struct
};
#define MG_AGENT_WORD_ARRAYS 9
#define MG_FP_WORDS 8  // own cell + up to 3 line cells, victim box rows / cols, class as a target, spare
#ifndef MG_MIN_CTAS_PER_SM
#define MG_MIN_CTAS_PER_SM 8  // 64 registers per thread -> 32 resident warps per SM (measured best once the grid stopped
                              // occupying shared memory: profiles/README.md; 7 was best before)
#endif
#ifndef MG_FINISH_CTAS_PER_SM
#define MG_FINISH_CTAS_PER_SM MG_MIN_CTAS_PER_SM  // k_finish's own register budget (it mostly waits on dependent loads)
#endif

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

// bytes of dynamic shared memory per warp / per CTA (must match carve())
__host__ __device__ inline size_t smem_per_warp(int HWp, int T, int A) {  // HWp = 0: the grid is not staged
  size_t n = 0;
  n += align16((size_t)HWp * 2);
  n += align16(8 * sizeof(int) + 2 * MG_RNG_WINDOW * sizeof(uint32_t));
  n += align16((size_t)A * 4) * MG_AGENT_WORD_ARRAYS;
  n += 2 * align16((size_t)A * 2);
  n += align16((size_t)(A < 32 ? A : 32) * MG_FP_WORDS * 4);
  n += align16(sizeof(Wv)) + align16(sizeof(Smem));
  return n;
}
__host__ __device__ inline size_t smem_per_cta(int) { return align16(MGH_HEADER_WORDS * 4); }

__device__ __forceinline__ void carve(const MgDev& d, unsigned char* base, int warp, Smem& s) {
  s.hdr = (int32_t*)base;
  base += align16(MGH_HEADER_WORDS * 4);
  base += (size_t)warp * smem_per_warp(d.stage_grid ? d.HWp : 0, d.T, d.A);
  s.cells = (uint16_t*)base;
  base += align16((size_t)(d.stage_grid ? d.HWp : 0) * 2);
  s.rs = (int*)base;
  s.rand = (uint32_t*)(base + 8 * sizeof(int));
  base += align16(8 * sizeof(int) + 2 * MG_RNG_WINDOW * sizeof(uint32_t));
  size_t aw = align16((size_t)d.A * 4);
  s.a_slot = (uint32_t*)base, base += aw;
  s.a_loc = (uint32_t*)base, base += aw;
  s.a_step = (uint32_t*)base, base += aw;
  s.a_act = (int32_t*)base, base += aw;
  s.a_vact = (int32_t*)base, base += aw;
  s.a_res = (uint32_t*)base, base += aw;
  s.a_locp = (uint32_t*)base, base += aw;
  s.a_locv = (uint32_t*)base, base += aw;
  s.a_exec = (int32_t*)base, base += aw;
  s.a_order = (uint16_t*)base, base += align16((size_t)d.A * 2);
  s.a_pos = (uint16_t*)base, base += align16((size_t)d.A * 2);
  s.fp = (uint32_t*)base, base += align16((size_t)(d.A < 32 ? d.A : 32) * MG_FP_WORDS * 4);
  s.wv = (Wv*)base, base += align16(sizeof(Wv));
  s.self = (Smem*)base;
}

// CTA prologue: the program header into shared memory (the interpreter reads it through the env view)
__device__ __forceinline__ void load_cta_tables(const MgDev& d, Smem& s) {
  for (int i = threadIdx.x; i < MGH_HEADER_WORDS; i += blockDim.x) s.hdr[i] = __ldg(d.P + i);
  __syncthreads();
}

__device__ __forceinline__ void bind_env(const MgDev& d, const Smem& s, int env, Wv& w) {
  w.P = d.P;
  w.hdr = s.hdr;
  w.cells_g = d.cells + (size_t)env * d.HWp;
  w.cells = d.stage_grid ? s.cells : w.cells_g;  // read in place by default (mg_capi.cu: stage_grid)
  w.objs = d.objs + (size_t)env * (d.maxobj + d.NPROXY) * d.OS;
  w.agents = d.agents + (size_t)env * d.A * d.AS;
  w.astats = d.astats + (size_t)env * d.A * d.SA;
  w.atouched = d.atouched + (size_t)env * d.A * d.SAW;
  w.gstats = d.gstats + (size_t)env * d.SG;
  w.gtouched = d.gtouched + (size_t)env * d.SGW;
  w.cover = d.cover + (size_t)env * d.A * d.CW;
  w.rng = d.rng + (size_t)env * MG_RNG_WORDS;
  w.E = d.env + (size_t)env * MGEV_WORDS;
  w.logtab = d.logtab;
  w.H = d.H, w.W = d.W, w.A = d.A, w.R = d.R, w.TW = d.TW, w.OS = d.OS, w.AS = d.AS;
  w.SA = d.SA, w.SAW = d.SAW, w.T = d.T, w.B = d.B, w.ND = d.ND, w.NOFF = d.NOFF, w.CW = d.CW;
  w.obs = d.obs + (size_t)env * d.A * (size_t)(3 * d.T);
  w.maxobj = d.maxobj, w.NTERR = d.NTERR, w.NPROXY = d.NPROXY, w.PAD = d.PAD, w.WP = d.WP;
  w.TOKOFF = MG_TOKOFF(d.TW, d.R);
  w.ARENA = d.ARENA, w.AOECAP = d.AOECAP, w.AOEW = d.AOEW, w.PENDCAP = d.PENDCAP, w.TERRCAP = d.TERRCAP, w.NDYN = d.NDYN;
  w.arena = d.arena + (size_t)env * d.ARENA;
  w.aoe_src = d.aoe_src + (size_t)env * d.AOECAP * d.AOEW;
  w.aoe_pending = d.aoe_pending + (size_t)env * d.PENDCAP * 2;
  w.terr_src = d.terr_src + (size_t)env * d.TERRCAP * 4;
  w.terr_tab = d.terr_tab + (size_t)env * d.TERRCAP * 4;
  w.owner = d.owner_map + (size_t)env * d.NTERR * d.HW;
  w.inside_tag = d.inside_tag + (size_t)env * d.A * d.NTERR;
  w.dyn_stamp = d.dyn_stamp + (size_t)env * d.maxobj * d.NDYN;
  w.NTAGS = d.NTAGS;
  w.tag_lists = d.tag_lists + (size_t)env * d.NTAGS * MG_TAG_LIST_CAP;
  w.tag_state = d.tag_state + (size_t)env * d.NTAGS;
  s.rs[4] = 0;
  w.rs = s.rs;
  w.rand = s.rand;
  w.step = (uint32_t)w.E[MGEV_STEP];
}

// Publish the per-warp view structs to shared memory so that they cost no registers / local memory.
__device__ __forceinline__ void publish_warp(const MgDev& d, const Smem& s, int env, int lane) {
  if (lane == 0) {
    *s.self = s;
    bind_env(d, s, env, *s.wv);
  }
  __syncwarp();
}

__device__ __forceinline__ void stage_cells(const MgDev& d, const Wv& w, int lane) {
  if (!d.stage_grid) {
    __syncwarp();
    return;
  }
  const uint4* src = (const uint4*)w.cells_g;
  uint4* dst = (uint4*)w.cells;
  for (int i = lane; i < d.HWp / 8; i += 32) dst[i] = src[i];
  __syncwarp();
}


// Row of n bytes whose first `used` bytes are staged tokens and whose remainder is 0xFF.  The stage shares
// the destination's 16-byte phase, so the body moves as aligned 16-byte vectors; vectors that lie entirely
// behind the tokens are stored from a register without touching shared memory.
__device__ __forceinline__ void flush_row(uint8_t* src, uint8_t* g, int n, int used, int lane) {
  // pad the partial vector behind the last token (and the unaligned head) so whole vectors can be copied
  const int pad_to = min(n, ((used + 15) & ~15) + 16);
  if (used + lane < pad_to) src[used + lane] = 0xFF;
  __syncwarp();
  int head = (int)((16u - ((uint32_t)(uintptr_t)g & 15u)) & 15u);
  if (head > n) head = n;
  if (lane < head) g[lane] = lane < pad_to ? src[lane] : (uint8_t)0xFF;
  const int body = (n - head) >> 4;
  const uint4* s4 = (const uint4*)(src + head);
  uint4* g4 = (uint4*)(g + head);
  const uint4 ff = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
  const int tv = min(body, max(0, (pad_to - head) >> 4));  // vectors that lie entirely inside the staged (token + pad) bytes
  for (int i = lane; i < tv; i += 32) __stcs(g4 + i, s4[i]);
  for (int i = tv + lane; i < body; i += 32) __stcs(g4 + i, ff);
  const int done = head + (body << 4);
  if (lane < n - done) g[done + lane] = done + lane < pad_to ? src[done + lane] : (uint8_t)0xFF;
}


// An object's observation tokens (core/grid_object.cpp:178-203, objects/agent.cpp:142-154) only change
// when its tags, vibe or inventory change, so they are cached in the object record as (feature | value << 8)
// pairs and rebuilt lazily: agents by k_world (one lane per agent), anything else by whichever observer meets it
// first.  Concurrent observers may rebuild the same object: they write identical words, and the count is published
// behind a fence.
struct TokCtx {
  const int32_t* P;
  const int32_t* hdr;
  int32_t* E;
  int TW, B, ND, TOKOFF;
};
__device__ __forceinline__ TokCtx tok_ctx(const Wv& w) { return TokCtx{w.P, w.hdr, w.E, w.TW, w.B, w.ND, w.TOKOFF}; }
__device__ __noinline__ int rebuild_token_cache(const TokCtx c, uint32_t* o) {
  uint16_t* tk = (uint16_t*)(o + c.TOKOFF);
  const int cap = c.hdr[MGH_TOK_CAP];
  int n = 0;
  auto put = [&](int feat, int val) {
    if (n < cap) tk[n] = (uint16_t)((feat & 0xff) | ((val & 0xff) << 8));
    n++;
  };
  const int ftag = c.hdr[MGH_FEAT_TAG];
  for (int k = 0; k < c.TW; k++) {
    uint32_t m = o[MGO_TAGS + k];
    while (m) {
      int b = __ffs(m) - 1;
      put(ftag, k * 32 + b);
      m &= m - 1;
    }
  }
  int vibe = o_vibe(o);
  if (vibe) put(c.hdr[MGH_FEAT_VIBE], vibe);
  int fl = o_flags(o);
  if (fl & MGOF_OBS_INV) {
    uint64_t ord = o_order(o);
    int cnt = ord_count(ord);
    const uint16_t* inv = (const uint16_t*)(o + MGO_TAGS + c.TW);
    const int32_t* feats = c.P + c.hdr[MGS_INV_FEATS];
    for (int i = 0; i < cnt; i++) {
      int it = ord_item(ord, i);
      uint32_t amt = inv[it];
      int p = 0;
      do {
        put(__ldg(feats + it * c.ND + p), (int)(amt % (uint32_t)c.B));
        amt /= (uint32_t)c.B;
        p++;
      } while (amt > 0 && p < c.ND);
    }
  }
  if (fl & MGOF_AGENT) {
    int ai = o_agent(o);
    put(c.hdr[MGH_FEAT_GROUP], __ldg(c.P + c.hdr[MGS_TEMPLATES] + o_tmpl(o) * MG_TEMPLATE_WORDS + MGT_GROUP));
    put(c.hdr[MGH_FEAT_AGENT_ID], ai >= 0 ? ai : 0);
  }
  if (n > cap) {
    if (!(c.E[MGEV_ERROR] & MGERR_POOL_EXHAUSTED)) c.E[MGEV_ERR_INFO] = 19;
    atomicOr(&c.E[MGEV_ERROR], MGERR_POOL_EXHAUSTED);
    n = cap;
  }
  __threadfence();
  o[MGO_NTOK] = (uint32_t)n;
  return n;
}
__device__ __forceinline__ void put_token(uint8_t* out, int T, int pos, int loc, int feat, int val) {
  if (pos < T) {
    out[pos * 3 + 0] = (uint8_t)loc;
    out[pos * 3 + 1] = (uint8_t)feat;
    out[pos * 3 + 2] = (uint8_t)val;
  }
}

// =================================================================================================
// k_observe: the observation pass (bindings/mettagrid_c.cpp:665-912), ONE WARP PER AGENT.
//
// k_world (or k_init_buffers) leaves, per agent, {executed action, start-of-tick location, location, object slot}
// in d.obs_in and the tokens of the configured global game values in d.obsval; from there on an agent's row depends
// on nothing another agent's row computes -- with one exception, the reference's cell staleness (:787-796):
// an object's `visited` stamp is advanced by the FIRST agent, in index order, that sees it this tick, and the ticks
// since go to that agent's cell.visited stat.  Observers only CLAIM what they see -- one fire-and-forget
// atomicMax(claim[object], step << 32 | ~agent) each, so the lowest agent of the latest tick wins -- and k_finish, which
// runs when every row is written, advances the stamps of the objects claimed this tick and credits the winners.
//
// The program header travels as a kernel argument (constant bank) and the packed window offsets are a small
// read-only table (L1-resident), so a CTA has no prologue; blockIdx.x is the environment.
// =================================================================================================
#ifndef MG_OBS_MIN_WARPS
#define MG_OBS_MIN_WARPS 32  // resident warps per SM the register budget is chosen for (64 registers; A/B: profiles/README.md)
#endif
#define MG_OBS_MAX_WARPS 8   // warps (agents) per CTA: the launcher picks a divisor of A when it can

#define MG_OBS_MULTI 24      // multi-token objects per window listed in shared memory (more: copied lane-serially)
__host__ __device__ inline size_t obs_smem_bytes(int warps, int T) {  // per warp: the row's stage + the multi-token list
  return (size_t)warps * (align16((size_t)3 * T + 32) + (size_t)16 * MG_OBS_MULTI);
}

// NP = window passes of 32 cells whose loads are issued together: 4 (up to 11 x 11 windows) or 8 (up to 15 x 15)
template <bool PLAIN, int NP>
__global__ void __launch_bounds__(MG_OBS_MAX_WARPS * 32, MG_OBS_MIN_WARPS / MG_OBS_MAX_WARPS)
    k_observe(const MgDev d, const MgFastHdr HD, const uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int* const hdr = HD.v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int env = d.env0 + blockIdx.x, a = blockIdx.y * (blockDim.x >> 5) + warp;  // grid.x = environments (no 65535 limit)
  const int A = d.A, T = d.T;
  if (a >= A || (mask && !mask[env])) return;

  // 32-bit index arithmetic throughout (one widening multiply per base pointer): rows, slots and strides all fit
  const uint32_t row = (uint32_t)env * (uint32_t)A + (uint32_t)a;
  const uint4 me_in = d.obs_in[row];  // executed action, start-of-tick location, location, object slot
  const int r0 = (int)(me_in.z >> 16), c0 = (int)(me_in.z & 0xffffu);
  const uint32_t step = (uint32_t)d.env[(uint32_t)env * (uint32_t)MGEV_WORDS + MGEV_STEP];
  uint32_t* const objs = d.objs + (size_t)(uint32_t)env * (size_t)d.objs_stride;
  const uint16_t* const centre = d.cells + (size_t)(uint32_t)env * (size_t)(uint32_t)d.HWp + (uint32_t)((r0 + d.PAD) * d.WP + c0 + d.PAD);
  uint8_t* const g_row = d.obs + (size_t)row * (size_t)(uint32_t)(3 * T);
  const uint32_t stage_bytes = ((uint32_t)(3 * T + 32) + 15u) & ~15u;
  uint8_t* const out = smem_raw + (uint32_t)warp * stage_bytes + ((uint32_t)(uintptr_t)g_row & 15u);
  const int OS = d.OS, NOFF = d.NOFF, TOKOFF = MG_TOKOFF(d.TW, d.R);
  const uint32_t lt = (1u << lane) - 1u;

  // ---- window cells in Manhattan order, 32 per pass (:756-811).  The grid carries an empty frame as wide as the
  // window radius, so a cell outside the map reads as empty and needs no bounds test.  All cell ids are loaded
  // first, then the heads of the objects standing there (cached token count, first token pair): two memory round
  // trips per agent, issued before anything else is computed.
  uint32_t pk[NP], slot[NP], nt[NP], t0[NP];
  uint8_t own0[NP];  // territory 0's winner at the cell (aoe_mask tokens), loaded with the cell ids
  const int fmask = (!PLAIN && d.NTERR > 0) ? hdr[MGH_FEAT_AOE_MASK] : 0;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const int k = 32 * p + lane;
    pk[p] = k < NOFF ? __ldg(d.obs_offs + k) : 0u;
    slot[p] = k < NOFF ? (uint32_t)centre[(int)(short)(pk[p] >> 16)] : 0u;
    own0[p] = 0xFF;  // 0xFF: outside the map (no token)
    if (!PLAIN && fmask && k < NOFF) {
      const int r = r0 + (int)(pk[p] & 15u) - 8, c = c0 + (int)((pk[p] >> 4) & 15u) - 8;
      if (r >= 0 && c >= 0 && r < d.H && c < d.W) own0[p] = d.owner_map[(size_t)env * d.NTERR * d.HW + r * d.W + c];
    }
    if (MG_CHECKED && slot[p] >= (uint32_t)d.maxobj) {  // checked builds: a cell must name a slot of this env's pool
      atomicOr(&d.env[(size_t)env * MGEV_WORDS + MGEV_ERROR], MGERR_BOUNDS);
      slot[p] = 0;
    }
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {
    nt[p] = t0[p] = 0;
    if (slot[p]) {
      const uint32_t* o = objs + slot[p] * (uint32_t)OS;
      nt[p] = o[MGO_NTOK], t0[p] = o[TOKOFF];
    }
  }

  // ---- global tokens (:700-742), one candidate per lane, compacted in order
  const int flags = hdr[MGH_GLOBAL_FLAGS];
  int feat = 0, val = 0, have = 0;
  if (lane == 0 && (flags & MGG_EPISODE_PCT)) {
    const int ms = hdr[MGH_MAX_STEPS];
    have = 1, feat = hdr[MGH_FEAT_EPISODE_PCT];
    if (ms > 0) val = step >= (uint32_t)ms ? 255 : (int)((256u * step / (uint32_t)ms) & 0xffu);
  } else if (lane == 1 && (flags & MGG_LAST_ACTION)) {
    have = 1, feat = hdr[MGH_FEAT_LAST_ACTION], val = (int)me_in.x & 0xff;
  } else if (lane == 2 && (flags & MGG_LAST_ACTION_MOVE) && hdr[MGH_FEAT_LAST_ACTION_MOVE] != 0) {
    have = 1, feat = hdr[MGH_FEAT_LAST_ACTION_MOVE], val = me_in.z != me_in.y;
  } else if (lane == 3 && (flags & MGG_LAST_REWARD)) {
    // rewards are zeroed before and written after the observation pass (:937-938,1062,1070): always 0
    have = 1, feat = hdr[MGH_FEAT_LAST_REWARD], val = 0;
  } else if ((lane == 4 || lane == 5) && (flags & MGG_LOCAL_POSITION)) {
    const uint32_t sp = d.agents[(size_t)row * (size_t)(uint32_t)d.AS + MGAG_SPAWN];
    const int dd = lane == 4 ? c0 - (int)(sp & 0xffffu) : (int)(sp >> 16) - r0;
    if (dd != 0) {
      have = 1;
      val = min(abs(dd), 255);
      feat = lane == 4 ? (dd > 0 ? hdr[MGH_FEAT_LP_EAST] : hdr[MGH_FEAT_LP_WEST]) : (dd > 0 ? hdr[MGH_FEAT_LP_NORTH] : hdr[MGH_FEAT_LP_SOUTH]);
    }
  }
  const uint32_t gm = __ballot_sync(MG_FULL, have);
  if (have) put_token(out, T, __popc(gm & lt), 0xFE, feat, val);
  int base = __popc(gm);
  // ---- configured global game values (:1207-1238): evaluated by k_world, one (feature | value << 8) pair per token
  if (!PLAIN && d.OVW > 0) {
    const uint16_t* ov = d.obsval + (size_t)row * (size_t)(uint32_t)d.OVW;
    const int n = ov[0];
    for (int j = lane; j < n; j += 32) put_token(out, T, base + j, 0xFE, ov[1 + j] & 0xff, ov[1 + j] >> 8);
    base += n;
  }

  // ---- tokens.  The whole window is settled at once: lane l holds the cells of rank l, l + 32, ... (Manhattan order),
  // so "tokens before mine" = the cells of earlier passes + the lower lanes of my pass: one ballot and two popcounts
  // per pass for the cells that emit one token (every wall), plus a short loop over the few cells that emit more.
  const uint32_t* const me = objs + (size_t)me_in.w * OS;
  unsigned long long* const claims = d.claims + (size_t)(uint32_t)env * (size_t)(uint32_t)d.maxobj;
  const unsigned long long my_claim = ((unsigned long long)step << 32) | (unsigned long long)(0xffffffffu - (uint32_t)a);
  int n[NP], tm[NP], cnt[NP], pos[NP];
  uint32_t m_any[NP], m_multi[NP];
  bool dirty = false;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    tm[p] = 0;
    if (!PLAIN && fmask && own0[p] != 0xFF) {
      // :337-362, one aoe_mask token per in-map cell, before the cell's object tokens (ownership map: mg_world.cuh)
      const int r = r0 + (int)(pk[p] & 15u) - 8, c = c0 + (int)((pk[p] >> 4) & 15u) - 8;
      const uint8_t* own = d.owner_map + (size_t)env * d.NTERR * d.HW + r * d.W + c;
      for (int ti = 0; ti < d.NTERR && !tm[p]; ti++) {
        const int v = ti == 0 ? (int)own0[p] : (int)own[(size_t)ti * d.HW];
        if (v) {
          const int tag = __ldg(d.P + hdr[MGS_POOL] + __ldg(d.P + hdr[MGS_TERRITORIES] + ti * MG_TERR_WORDS) + v - 1);
          tm[p] = o_has_tag(me, tag) ? 1 : 2;
        }
      }
    }
    n[p] = 0;
    if (slot[p]) {
      atomicMax(claims + slot[p], my_claim);  // cell staleness (:787-796): settled by k_finish
      n[p] = (int)nt[p];
      dirty = dirty || nt[p] == MG_TOK_DIRTY;
    }
  }
  if (__any_sync(MG_FULL, dirty)) {  // a non-agent object changed since it was last seen: rebuild its cached tokens
#pragma unroll
    for (int p = 0; p < NP; p++)
      if (slot[p] && nt[p] == MG_TOK_DIRTY) {
        uint32_t* o = objs + (size_t)slot[p] * OS;
        n[p] = rebuild_token_cache(TokCtx{d.P, d.P, d.env + (size_t)env * MGEV_WORDS, d.TW, d.B, d.ND, TOKOFF}, o);  // header = start of P
        t0[p] = o[TOKOFF];
      }
  }
  int run = base;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    cnt[p] = n[p] + (tm[p] != 0);
    m_any[p] = __ballot_sync(MG_FULL, cnt[p] != 0);
    m_multi[p] = __ballot_sync(MG_FULL, cnt[p] > 1);
    pos[p] = run + __popc(m_any[p] & lt);
    run += __popc(m_any[p]);
  }
#pragma unroll
  for (int p = 0; p < NP; p++) {  // cells with more than one token push everything behind them
    uint32_t m = m_multi[p];
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      const int extra = __shfl_sync(MG_FULL, cnt[p], b) - 1;
#pragma unroll
      for (int q = p; q < NP; q++) pos[q] += (q > p || lane > b) ? extra : 0;
      run += extra;
    }
  }
  // single tokens straight from registers; multi-token objects go through a list in shared memory
  uint32_t* const mlist = (uint32_t*)(smem_raw + (blockDim.x >> 5) * stage_bytes) + warp * (4 * MG_OBS_MULTI);
  int nmulti = 0;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    const int loc = (int)((pk[p] >> 8) & 0xffu);
    if (tm[p]) put_token(out, T, pos[p], loc, fmask, tm[p]);
    const int tpos = pos[p] + (tm[p] != 0);
    if (n[p] == 1) put_token(out, T, tpos, loc, (int)(t0[p] & 0xffu), (int)((t0[p] >> 8) & 0xffu));
    const uint32_t mo = __ballot_sync(MG_FULL, n[p] > 1);
    if (n[p] > 1) {
      const int i = nmulti + __popc(mo & lt);
      if (i < MG_OBS_MULTI) {
        uint32_t* e = mlist + 4 * i;
        e[0] = slot[p], e[1] = (uint32_t)tpos, e[2] = (uint32_t)loc, e[3] = (uint32_t)n[p];
      }
    }
    nmulti += __popc(mo);
  }
  __syncwarp();
  // the whole warp copies each listed object's cached (feature, value) pairs behind its location byte, one lane per
  // token, the loads of four objects in flight together; anything beyond the list's capacity is copied by its own lane
  const int listed = min(nmulti, MG_OBS_MULTI);
  for (int i0 = 0; i0 < listed; i0 += 4) {
    uint32_t ev[4];
    uint4 en[4];
#pragma unroll
    for (int q = 0; q < 4; q++) {
      en[q] = make_uint4(0, 0, 0, 0);
      if (i0 + q < listed) {
        en[q] = *(const uint4*)(mlist + 4 * (i0 + q));
        if (lane < (int)en[q].w) ev[q] = ((const uint16_t*)(objs + en[q].x * (uint32_t)OS + TOKOFF))[lane];
      }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
      if (lane < (int)en[q].w) put_token(out, T, (int)en[q].y + lane, (int)en[q].z, (int)(ev[q] & 0xffu), (int)(ev[q] >> 8));
      if ((int)en[q].w > 32) {  // longer lists than lanes: the rest straight from the record
        const uint16_t* tk = (const uint16_t*)(objs + (size_t)en[q].x * OS + TOKOFF);
        for (int j = lane + 32; j < (int)en[q].w; j += 32) {
          const uint32_t e = tk[j];
          put_token(out, T, (int)en[q].y + j, (int)en[q].z, (int)(e & 0xffu), (int)(e >> 8));
        }
      }
    }
  }
  if (nmulti > MG_OBS_MULTI) {  // more multi-token objects in one window than the list holds (rare): each by its own lane
    int seen = 0;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const uint32_t mo = __ballot_sync(MG_FULL, n[p] > 1);
      if (n[p] > 1 && seen + __popc(mo & lt) >= MG_OBS_MULTI) {
        const uint16_t* tk = (const uint16_t*)(objs + (size_t)slot[p] * OS + TOKOFF);
        const int tpos = pos[p] + (tm[p] != 0), loc = (int)((pk[p] >> 8) & 0xffu);
        for (int j = 0; j < n[p]; j++) put_token(out, T, tpos + j, loc, (int)(tk[j] & 0xffu), (int)(tk[j] >> 8));
      }
      seen += __popc(mo);
    }
  }
  base = run;
  if (lane == 0) d.tok_attempted[row] = base;  // k_finish adds the env's token stats in agent order (:659-661)
  // ---- stream out: token bytes come from the stage, the rest of the row is 0xFF (EmptyTokenByte, :940-942)
  __syncwarp();
  flush_row(out, g_row, 3 * T, 3 * min(base, T), lane);
}

// What k_world / k_init_buffers leave for the observation pass: per agent {executed action, start-of-tick location,
// location, object slot}, the tokens of the configured global game values (:1207-1238, evaluated here because they need
// the interpreter), fresh token caches for the agents, a fresh territory ownership map.
__device__ __noinline__ void hand_over(const MgDev& d, const Wv& w, const Smem& s, int env, int lane, bool initial) {
  const int A = w.A;
  const TokCtx tc = tok_ctx(w);
  const int nov = w.hdr[MGH_NUM_OBS_VALUES];
  for (int a = lane; a < A; a += 32) {
    const uint32_t slot = s.a_slot[a];
    uint32_t* o = objp(w, (int)slot);
    const uint32_t loc = o[MGO_LOC];
    d.obs_in[(size_t)env * A + a] = make_uint4(initial ? 0u : (uint32_t)s.a_exec[a], initial ? loc : s.a_step[a], loc, slot);
    if (o[MGO_NTOK] == MG_TOK_DIRTY) rebuild_token_cache(tc, o);
    if (nov > 0) {
      uint16_t* ov = d.obsval + ((size_t)env * A + a) * d.OVW;
      const int32_t* cfg = sec(w, MGS_OBS_VALUES);
      int n = 0;
      for (int i = 0; i < nov; i++) {
        Ctx vc = make_ctx();
        vc.actor = vc.target = (uint16_t)slot;
        uint32_t enc = (uint32_t)eval_value(w, MG_DEPTH, __ldg(cfg + 2 * i + 1), vc, (int)slot);
        int f = __ldg(cfg + 2 * i);
        do {  // systems/encoding_utils.hpp:16-35: base-B digits on consecutive feature ids, at least one token
          if (1 + n < d.OVW) ov[1 + n] = (uint16_t)((f & 0xff) | ((enc % (uint32_t)w.B) << 8));
          n++, f++;
          enc /= (uint32_t)w.B;
        } while (enc > 0);
      }
      ov[0] = (uint16_t)min(n, d.OVW - 1);
    }
  }
  __syncwarp();
  terr_refresh(w, lane);  // the aoe_mask tokens read the ownership map
}

// =================================================================================================
// k_reset: the constructor.  Objects get ids 1.. in row-major map order, agents get ids in
// encounter order (bindings/mettagrid_c.cpp:222-267); each lane builds the object of one cell.
// =================================================================================================
__global__ void __launch_bounds__(MG_WARPS_PER_CTA * 32) k_reset(MgDev d, const uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mask) {  // a masked launch usually selects few environments: a CTA without any leaves before loading tables
    const int e = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
    if (!__syncthreads_or(e < d.num_envs && mask[e])) return;
  }
  {
    Smem cta;
    carve(d, smem_raw, warp, cta);
    load_cta_tables(d, cta);
  }
  const int env = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
  if (env >= d.num_envs) return;
  if (mask && !mask[env]) return;
  Smem s0;
  carve(d, smem_raw, warp, s0);
  publish_warp(d, s0, env, lane);
  Wv& w = *s0.wv;

  for (int i = lane; i < d.HWp; i += 32) w.cells_g[i] = 0;
  for (int i = lane; i < d.maxobj; i += 32) d.claims[(size_t)env * d.maxobj + i] = 0ull, d.visited[(size_t)env * d.maxobj + i] = 0u;
  for (int i = lane; i < d.A * d.SA; i += 32) w.astats[i] = 0.0f;
  for (int i = lane; i < d.A * d.SAW; i += 32) w.atouched[i] = 0;
  for (int i = lane; i < d.A * d.CW; i += 32) w.cover[i] = 0;
  for (int i = lane; i < d.A; i += 32) d.success[(size_t)env * d.A + i] = 0;
  for (int i = lane; i < d.SGW; i += 32) w.gtouched[i] = 0;
  __syncwarp();
  for (int i = lane; i < d.SG; i += 32) {
    float v = d.init_gstats ? d.init_gstats[(size_t)env * d.SG + i] : 0.0f;
    w.gstats[i] = v;
    if (v != 0.0f) atomicOr(&w.gtouched[i >> 5], 1u << (i & 31));
  }
  if (lane == 0) {
    for (int i = 0; i < MGEV_WORDS; i++) w.E[i] = 0;
    w.E[MGEV_RNG_IDX] = MG_RNG_WORDS;
  }
  {
    // std::mt19937(seed) (bits/random.tcc seed()): a 624-step serial recurrence.  The seeded state is kept per env
    // (word 624 = the seed it was made from, word 625 = valid), so that a reset with the same seed -- every
    // auto-reset of a vectorised env -- is a coalesced copy.
    uint32_t* cache = d.rng_seeded + (size_t)env * (MG_RNG_WORDS + 2);
    const uint32_t seed = d.seeds[env];
    const bool hit = cache[MG_RNG_WORDS + 1] == 1u && cache[MG_RNG_WORDS] == seed;
    __syncwarp();
    if (!hit && lane == 0) {
      uint32_t x = seed;
      cache[0] = x;
      for (int i = 1; i < MG_RNG_WORDS; i++) {
        x = 1812433253u * (x ^ (x >> 30)) + (uint32_t)i;
        cache[i] = x;
      }
      cache[MG_RNG_WORDS] = seed, cache[MG_RNG_WORDS + 1] = 1u;
    }
    __syncwarp();
    for (int i = lane; i < MG_RNG_WORDS; i += 32) w.rng[i] = cache[i];
  }
  __syncwarp();
  if (lane == 0) {  // stat keys the constructor creates (:134-136)
    atomicOr(&w.gtouched[w.hdr[MGH_GST_TOKENS_WRITTEN] >> 5], 1u << (w.hdr[MGH_GST_TOKENS_WRITTEN] & 31));
    atomicOr(&w.gtouched[w.hdr[MGH_GST_TOKENS_DROPPED] >> 5], 1u << (w.hdr[MGH_GST_TOKENS_DROPPED] & 31));
    atomicOr(&w.gtouched[w.hdr[MGH_GST_TOKENS_FREE] >> 5], 1u << (w.hdr[MGH_GST_TOKENS_FREE] & 31));
  }

  const int16_t* init = d.init_cells + (size_t)env * d.HW;
  int nobj = 0, nagent = 0;
  const uint32_t lt = (1u << lane) - 1u;
  for (int base = 0; base < d.HW; base += 32) {
    int i = base + lane;
    int t = i < d.HW ? (int)init[i] : -1;
    int kind = -1;
    const int32_t* tp = nullptr;
    if (t >= 0) {
      tp = tmpl(w, t);
      kind = __ldg(tp + MGT_KIND);
    }
    uint32_t m = __ballot_sync(MG_FULL, t >= 0), ma = __ballot_sync(MG_FULL, kind == 1);
    if (t >= 0) {
      int slot = nobj + __popc(m & lt) + 1;
      int aidx = kind == 1 ? nagent + __popc(ma & lt) : -1;
      int r = i / d.W, c = i - r * d.W;
      if (slot >= d.maxobj || aidx >= d.A) {
        set_error(w, MGERR_POOL_EXHAUSTED, slot);
      } else {
        init_object(w, slot, t, r, c, aidx, true, (uint32_t)slot);
        uint32_t* o = objp(w, slot);
        w.cells_g[cidx(w, r, c)] = (uint16_t)slot;
        if (kind == 1) {
          const int32_t* iv = pool(w, __ldg(tp + MGT_INIT_INV));
          int ni = __ldg(tp + MGT_INIT_INV_N);
          uint32_t* ag = w.agents + aidx * d.AS;
          for (int k = 0; k < d.AS; k++) ag[k] = 0;
          ag[MGAG_OBJ] = (uint32_t)slot;
          ag[MGAG_SPAWN] = o[MGO_LOC];
          ag[MGAG_PREV_LOC] = o[MGO_LOC];
          ag[MGAG_STEP_LOC] = o[MGO_LOC];
          // populate_initial_inventory sets "<res>.amount" for every configured item (agent.cpp:79-84)
          for (int k = 0; k < ni; k++)
            astat_set(w, aidx, __ldg(sec(w, MGS_RES_STATS) + __ldg(iv + 2 * k) * 4 + 2), (float)__ldg(iv + 2 * k + 1));
          // init_reward: a top-level StatValue entry creates its key (core/game_value.cpp:40-47)
          const int32_t* rw = pool(w, __ldg(tp + MGT_REWARDS));
          int nr = __ldg(tp + MGT_REWARDS_N);
          for (int k = 0; k < nr; k++) {
            const int32_t* v = sec(w, MGS_VALUES) + __ldg(rw + 2 * k) * MG_VALUE_WORDS;
            if (__ldg(v) == MGV_STAT) {
              if (__ldg(v + 1) == MGSC_GAME)
                atomicOr(&w.gtouched[__ldg(v + 2) >> 5], 1u << (__ldg(v + 2) & 31));
              else
                astat_touch(w, aidx, __ldg(v + 2));
            }
          }
          for (int k = 0; k < d.NTERR; k++) w.inside_tag[aidx * d.NTERR + k] = -1;
        }
      }
    }
    nobj += __popc(m);
    nagent += __popc(ma);
  }
  __syncwarp();
  stage_cells(d, w, lane);  // the serial set-up below (materialized queries) reads the staged grid
  if (lane == 0) {
    w.E[MGEV_NEXT_OBJ] = nobj + 1;
    w.E[MGEV_NEXT_ID] = nobj + 1;
    w.E[MGEV_TAG_SEQ] = nobj + 1;
    if (nagent != d.A) set_error(w, MGERR_POOL_EXHAUSTED, -nagent - 1);
    for (int tg = 0; tg < d.NTAGS; tg++) w.tag_state[tg] = 0;  // TagIndex::register_object in creation order (:247)
    for (int slot = 1; slot <= nobj && d.NTAGS > 0; slot++) tl_register_object(w, slot, true);
    // AOE / territory sources register in object-creation order (:249-257); proxies for territory handlers
    if (d.AOECAP > 0 || d.TERRCAP > 0) {
      for (int slot = 1; slot <= nobj; slot++) {
        const int32_t* tp = tmpl(w, o_tmpl(objp(w, slot)));
        for (int k = 0; k < __ldg(tp + MGT_AOES_N); k++) register_aoe(w, slot, __ldg(tp + MGT_AOES) + k);
        const int32_t* tc = pool(w, __ldg(tp + MGT_TERR));
        for (int k = 0; k < __ldg(tp + MGT_TERR_N); k++) {
          int n = w.E[MGEV_NUM_TERR];
          if (n >= d.TERRCAP) {
            set_error(w, MGERR_POOL_EXHAUSTED, 18);
            break;
          }
          uint32_t* ts = w.terr_src + (size_t)n * 4;
          ts[0] = (uint32_t)slot, ts[1] = (uint32_t)__ldg(tc + 3 * k), ts[2] = (uint32_t)__ldg(tc + 3 * k + 1);
          ts[3] = (uint32_t)__ldg(tc + 3 * k + 2);
          w.E[MGEV_NUM_TERR] = n + 1;
          objp(w, slot)[MGO_META] |= (uint32_t)MGOF_TERR_SRC << 24;
        }
      }
    }
    for (int ti = 0; ti < d.NPROXY; ti++) {  // one proxy per (territory, agent)
      uint32_t* po = objp(w, d.maxobj + ti);
      for (int k = 0; k < d.OS; k++) po[k] = 0;
      po[MGO_META] = (uint32_t)w.hdr[MGH_PROXY_TEMPLATE];  // not alive, not in the grid
      po[MGO_AGENT] = (uint32_t)-1;
    }
    w.E[MGEV_TERR_STALE] = 3;
    if (w.hdr[MGH_NUM_MQ] > 0) {  // QuerySystem::compute_all (core/query_system.cpp:91-114)
      Ctx g = make_ctx();
      const int32_t* mq = sec(w, MGS_MQ);
      for (int i = 0; i < w.hdr[MGH_NUM_MQ]; i++) recompute_mq(w, MG_DEPTH, __ldg(mq + 2 * i), g, false);
    }
  }
  __syncwarp();
  terr_refresh(w, lane);  // source table + ownership map (all lanes)
}

// load per-agent slot/location into shared memory (all lanes)
__device__ __forceinline__ void load_agents(const Wv& w, const Smem& s, int lane) {
  for (int a = lane; a < w.A; a += 32) {
    uint32_t slot = w.agents[a * w.AS + MGAG_OBJ];
    s.a_slot[a] = slot;
    uint32_t loc = objp(w, (int)slot)[MGO_LOC];
    s.a_loc[a] = loc;
    s.a_step[a] = loc;
  }
  __syncwarp();
}

// =================================================================================================
// k_init_buffers: _init_buffers (bindings/mettagrid_c.cpp:294-319) + Agent::init via set_buffers
// =================================================================================================
__global__ void __launch_bounds__(MG_WARPS_PER_CTA * 32) k_init_buffers(MgDev d, const uint8_t* __restrict__ mask) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mask) {  // a masked launch usually selects few environments: a CTA without any leaves before loading tables
    const int e = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
    if (!__syncthreads_or(e < d.num_envs && mask[e])) return;
  }
  {
    Smem cta;
    carve(d, smem_raw, warp, cta);
    load_cta_tables(d, cta);
  }
  const int env = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
  if (env >= d.num_envs) return;
  if (mask && !mask[env]) return;
  Smem s0;
  carve(d, smem_raw, warp, s0);
  publish_warp(d, s0, env, lane);
  const Smem& s = *s0.self;
  Wv& w = *s0.wv;
  stage_cells(d, w, lane);
  load_agents(w, s, lane);
  for (int a = lane; a < w.A; a += 32) {
    // reset_coverage_tracking (objects/agent.cpp:41-47)
    uint32_t* cv = w.cover + a * d.CW;
    for (int k = 0; k < d.CW; k++) cv[k] = 0;
    uint32_t loc = s.a_loc[a];
    int cell = (int)(loc >> 16) * w.W + (int)(loc & 0xffffu);
    cv[cell >> 5] = 1u << (cell & 31);
    uint32_t* ag = w.agents + a * w.AS;
    ag[MGAG_UNIQUE] = 1;
    ag[MGAG_MAX_DIST] = 0;
    ag[MGAG_EPISODE_REWARD] = 0;
    astat_set(w, a, w.hdr[MGH_ST_UNIQUE_VISITED], 1.0f);
    astat_set(w, a, w.hdr[MGH_ST_MAX_DIST], 0.0f);
    size_t gi = (size_t)env * d.A + a;
    d.terminals[gi] = 0;
    d.truncations[gi] = 0;
    d.rewards[gi] = 0.0f;
  }
  __syncwarp();
  hand_over(d, w, s, env, lane, true);  // k_observe + k_finish follow in the same stream
}

// =================================================================================================
// k_world: phases 1-12 of one tick for every environment
// =================================================================================================
#define MGR_ACTED_P 1u
#define MGR_OK_P 2u
#define MGR_ACTED_V 4u
#define MGR_OK_V 8u
#define MGR_KIND_P_SHIFT 4
#define MGR_KIND_V_SHIFT 6

// actions/move.hpp:81-115 + change_vibe.hpp:48-57 + noop.hpp:21-23 (serial)
// PLAIN programs (no handlers anywhere: movement, vibes, observations, stats -- the reference's benchmark
// game) never enter the interpreter: a move is "relocate into an empty cell or fail".
template <bool PLAIN>
__device__ __noinline__ bool do_action(const Wv& w, int slot, int kind, int arg) {
  if (kind == MGA_NOOP) return true;
  uint32_t* o = objp(w, slot);
  if (kind == MGA_CHANGE_VIBE) {
    o_set_vibe(o, arg);
    return true;
  }
  // actions/orientation.hpp:28-48
  const int dr = (arg == 0 || arg == 4 || arg == 5) ? -1 : (arg == 1 || arg == 6 || arg == 7) ? 1 : 0;
  const int dc = (arg == 2 || arg == 4 || arg == 6) ? -1 : (arg == 3 || arg == 5 || arg == 7) ? 1 : 0;
  if constexpr (PLAIN) {
    // [TargetLocEmpty -> Relocate] then [TargetIsUsable -> UseTarget]; without on_use handlers the second fails
    return move_object(w, slot, o_r(o) + dr, o_c(o) + dc);
  }
  const int32_t* chain = sec(w, MGS_MOVE_CHAIN);
  const int nh = w.hdr[MGH_NUM_MOVE_HANDLERS];
  for (int k = 0; k < nh; k++) {
    const int4 mh = __ldg((const int4*)(chain + k * MG_MOVEH_WORDS));  // handler, range, accepts_empty, builtin
    for (int i = 1; i <= mh.y; i++) {
      int tr = o_r(o) + dr * i, tc = o_c(o) + dc * i;
      if (!valid_loc(w, tr, tc)) break;
      int t = cell_at(w, tr, tc);
      if (!t && !mh.z) continue;
      if (mh.w == MGMB_RELOCATE) {  // [TargetLocEmpty -> Relocate]
        if (t == 0 && move_object(w, slot, tr, tc)) return true;
        break;
      }
      if (mh.w == MGMB_USE_TARGET && __ldg(tmpl(w, o_tmpl(objp(w, t))) + MGT_ON_USE) < 0) break;  // onUse -> false
      Ctx ctx = make_ctx();
      ctx.actor = slot;
      ctx.target = t;
      ctx.tr = tr;
      ctx.tc = tc;
      ctx.distance = i;
      ctx.move_dir = arg;
      if (handler_apply(w, MG_DEPTH, mh.x, ctx)) return true;
      break;
    }
  }
  return false;
}

// One agent's action of one stream, with the bookkeeping inputs the per-agent pass below needs (:966-999)
template <bool PLAIN>
__device__ __forceinline__ void run_action(const Wv& w, const Smem& s, int a, int stream, int idx, int kind, int arg) {
  const int slot = (int)s.a_slot[a];
  const bool ok = do_action<PLAIN>(w, slot, kind, arg);
  const uint32_t loc = objp(w, slot)[MGO_LOC];
  if (stream == 0) {
    s.a_res[a] |= MGR_ACTED_P | (ok ? MGR_OK_P : 0u) | ((uint32_t)kind << MGR_KIND_P_SHIFT);
    s.a_locp[a] = loc;
  } else {
    s.a_res[a] |= MGR_ACTED_V | (ok ? MGR_OK_V : 0u) | ((uint32_t)kind << MGR_KIND_V_SHIFT);
    s.a_locv[a] = loc;
  }
  if (ok) s.a_exec[a] = idx;
}

// lane 0: one (priority, stream) pass in the shuffled order -- the reference's loop as it is written
template <bool PLAIN>
__device__ __noinline__ void action_pass_serial(const Wv& w, const Smem& s, int stream, int prio) {
  const int NA = w.hdr[MGH_NUM_ACTIONS];
  const int32_t* acts = sec(w, MGS_ACTIONS);
  for (int i = 0; i < w.A; i++) {
    const int a = s.a_order[i];
    const int idx = stream ? s.a_vact[a] : s.a_act[a];
    if (idx < 0 || idx >= NA) continue;  // invalid indices are accounted per agent below
    const int4 act = __ldg((const int4*)(acts + idx * MG_ACTION_WORDS));  // kind, arg, priority, is_vibe
    if (act.w != stream || act.z != prio) continue;
    run_action<PLAIN>(w, s, a, stream, idx, act.x, act.y);
  }
}

// All lanes, lane = agent (A <= 32): one (priority, stream) pass with independent agents resolved concurrently.
//
// The reference runs the agents one after the other in the shuffled order.  An action reads and writes only the
// actor, the cells of the line its move scans (MGH_MOVE_REACH cells in its direction) and the objects standing there
// -- unless the compiler's effect analysis says otherwise (MGS_TMPL_CLASS: SHARED actions also update env-wide
// structures and keep their mutual order; SERIAL ones send the whole pass down the serial loop).  Two agents whose
// footprints (own cell + line cells) share no cell cannot observe each other, so any interleaving of them gives the
// sequential result.  Agents are grouped into components of the "footprints meet" relation (transitive closure over
// 32-bit rows); inside a component they run in shuffled order, one per round; different components run in the same
// rounds on different lanes.
//  * A swap moves its TARGET, whose own action then starts from the swapper's cell: every agent standing in the line of
//    an action that may swap gets the box around that cell added to its footprint, which keeps a component's region
//    closed under displacement.
//  * What an action may do depends on what it meets, and members of its component may walk (or be swapped, or spawn
//    something) into its line before its turn: an agent with company takes the highest class any member could offer
//    it as a target, and joins the SHARED chain if that makes it SHARED.
template <bool PLAIN>
__device__ __noinline__ void action_pass(const Wv& w, const Smem& s, int lane, int stream, int prio, int mypos) {
  const int A = w.A, NA = w.hdr[MGH_NUM_ACTIONS];
  int idx = -1, kind = 0, arg = 0;
  bool active = false;
  if (lane < A) {
    idx = stream ? s.a_vact[lane] : s.a_act[lane];
    if (idx >= 0 && idx < NA) {
      const int4 act = __ldg((const int4*)(sec(w, MGS_ACTIONS) + idx * MG_ACTION_WORDS));  // kind, arg, priority, is_vibe
      active = act.w == stream && act.z == prio;
      kind = act.x, arg = act.y;
    }
  }
  const uint32_t m_active = __ballot_sync(MG_FULL, active);
  if (!m_active) return;
  if (!__ballot_sync(MG_FULL, active && kind == MGA_MOVE)) {  // noop / change_vibe only touch the acting agent
    if (active) run_action<PLAIN>(w, s, lane, stream, idx, kind, arg);
    __syncwarp();
    return;
  }
  // ---- footprints
  const int R = w.hdr[MGH_MOVE_REACH];
  const bool mover = active && kind == MGA_MOVE;
  const uint32_t NONE = 0xffffffffu;
  int r = 0, c = 0, cls = MGC_LOCAL, as_target = MGC_LOCAL;
  uint32_t victims = 0, cell[4] = {NONE, NONE, NONE, NONE};
  const int32_t* tcls = sec(w, MGS_TMPL_CLASS);
  if (lane < A) {
    const uint32_t* me = objp(w, (int)s.a_slot[lane]);
    r = o_r(me), c = o_c(me);
    cell[0] = me[MGO_LOC];
    if (!PLAIN) as_target = __ldg(tcls + o_tmpl(me)) & 3;
  }
  int r0 = r, r1 = r, c0 = c, c1 = c;        // bounding box of everything below (cheap first test)
  int br0 = 1, br1 = 0, bc0 = 1, bc1 = 0;    // box part of the footprint: empty unless displaced / long lines
  if (mover) {
    const int dr = (arg == 0 || arg == 4 || arg == 5) ? -1 : (arg == 1 || arg == 6 || arg == 7) ? 1 : 0;
    const int dc = (arg == 2 || arg == 4 || arg == 6) ? -1 : (arg == 3 || arg == 5 || arg == 7) ? 1 : 0;
    if (!PLAIN) cls = w.hdr[MGH_CHAIN_EMPTY_CLASS];
    for (int k = 1; k <= R; k++) {
      const int rr = r + dr * k, cc = c + dc * k;
      if (!valid_loc(w, rr, cc)) break;
      r0 = min(r0, rr), r1 = max(r1, rr), c0 = min(c0, cc), c1 = max(c1, cc);
      if (k <= 3) cell[k] = ((uint32_t)rr << 16) | (uint32_t)cc;
      if (PLAIN) continue;
      const int t = cell_at(w, rr, cc);
      if (!t) continue;
      const uint32_t* to = objp(w, t);
      const int tc = __ldg(tcls + o_tmpl(to));
      cls = max(cls, tc & 3);
      if ((tc & MGC_DISPLACES) && o_is_agent(to) && o_agent(to) >= 0) victims |= 1u << o_agent(to);
    }
    if (R > 3) br0 = r0, br1 = r1, bc0 = c0, bc1 = c1;  // lines longer than the cell list: their box stands in
  }
  if (__ballot_sync(MG_FULL, cls >= MGC_SERIAL)) {
    if (lane == 0) action_pass_serial<PLAIN>(w, s, stream, prio);
    __syncwarp();
    return;
  }
  if (__ballot_sync(MG_FULL, victims != 0)) {
    for (int i = 0; i < A; i++) {
      const uint32_t vi = __shfl_sync(MG_FULL, victims, i);
      const int ri = __shfl_sync(MG_FULL, r, i), ci = __shfl_sync(MG_FULL, c, i);
      if ((vi >> lane) & 1u) {
        const int a0 = max(ri - R, 0), a1 = min(ri + R, w.H - 1), b0 = max(ci - R, 0), b1 = min(ci + R, w.W - 1);
        if (br0 > br1) br0 = a0, br1 = a1, bc0 = b0, bc1 = b1;
        else br0 = min(br0, a0), br1 = max(br1, a1), bc0 = min(bc0, b0), bc1 = max(bc1, b1);
      }
    }
    if (br0 <= br1) r0 = min(r0, br0), r1 = max(r1, br1), c0 = min(c0, bc0), c1 = max(c1, bc1);
  }
  if (lane >= A) r0 = 1, r1 = 0;  // an empty box meets nothing
  if (lane < A) {
    uint32_t* f = s.fp + lane * MG_FP_WORDS;
    f[0] = cell[0], f[1] = cell[1], f[2] = cell[2], f[3] = cell[3];
    f[4] = (uint32_t)br0 | ((uint32_t)br1 << 16), f[5] = (uint32_t)bc0 | ((uint32_t)bc1 << 16);
    f[6] = (uint32_t)as_target;
  }
  // ---- who meets whom: bounding boxes first (registers), exact cells for the few candidates (shared memory)
  uint32_t cand = 0;
  for (int j = 0; j < A; j++) {
    const int a0 = __shfl_sync(MG_FULL, r0, j), a1 = __shfl_sync(MG_FULL, r1, j);
    const int b0 = __shfl_sync(MG_FULL, c0, j), b1 = __shfl_sync(MG_FULL, c1, j);
    if (a0 <= r1 && r0 <= a1 && b0 <= c1 && c0 <= b1) cand |= 1u << j;
  }
  __syncwarp();
  uint32_t comp = lane < A ? 1u << lane : 0u;
  auto in_box = [](uint32_t cl, int x0, int x1, int y0, int y1) {
    const int rr = (int)(cl >> 16), cc = (int)(cl & 0xffffu);
    return cl != 0xffffffffu && rr >= x0 && rr <= x1 && cc >= y0 && cc <= y1;
  };
  if (lane < A) {
    uint32_t m = cand & ~(1u << lane);
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const uint32_t* f = s.fp + j * MG_FP_WORDS;
      const int jr0 = (int)(f[4] & 0xffffu), jr1 = (int)(f[4] >> 16), jc0 = (int)(f[5] & 0xffffu), jc1 = (int)(f[5] >> 16);
      bool meet = br0 <= br1 && jr0 <= jr1 && jr0 <= br1 && br0 <= jr1 && jc0 <= bc1 && bc0 <= jc1;
#pragma unroll
      for (int x = 0; x < 4; x++) {
        const uint32_t fj = f[x];
        meet = meet || in_box(fj, br0, br1, bc0, bc1) || in_box(cell[x], jr0, jr1, jc0, jc1);
#pragma unroll
        for (int y = 0; y < 4; y++) meet = meet || (fj != NONE && fj == cell[y]);
      }
      if (meet) comp |= 1u << j;
    }
  }
  // the relation is symmetric by construction; close it, then settle the classes
  const int spawn_cls = PLAIN ? 0 : w.hdr[MGH_SPAWN_CLASS];
  for (int iter = 0;; iter++) {
    for (int k = 0; k < A; k++) {  // Warshall: whoever reaches k also reaches what k reaches
      const uint32_t rk = __shfl_sync(MG_FULL, comp, k);
      if ((comp >> k) & 1u) comp |= rk;
    }
    if (PLAIN) break;
    int up = cls;
    if (mover) {
      uint32_t m = comp & ~(1u << lane);
      if (m) up = max(up, spawn_cls);
      while (m) {
        const int j = __ffs(m) - 1;
        m &= m - 1;
        up = max(up, (int)s.fp[j * MG_FP_WORDS + 6]);
      }
    }
    const bool changed = up != cls;
    cls = up;
    if (__ballot_sync(MG_FULL, cls >= MGC_SERIAL) || iter >= 6) {
      if (lane == 0) action_pass_serial<PLAIN>(w, s, stream, prio);
      __syncwarp();
      return;
    }
    const uint32_t m_shared = __ballot_sync(MG_FULL, mover && cls == MGC_SHARED);
    if (mover && cls == MGC_SHARED) comp |= m_shared;
    if (!__ballot_sync(MG_FULL, changed)) {
      if (iter == 0 && m_shared) continue;  // the SHARED chain was just linked: close once more
      break;
    }
  }
  // ---- rounds: an agent's turn = the number of acting members of its component that come earlier in the order
  const uint32_t peers = comp & m_active & ~(1u << lane);
  int turn = 0;
  if (__ballot_sync(MG_FULL, active && peers != 0)) {
    for (int j = 0; j < A; j++) {
      const int pj = __shfl_sync(MG_FULL, mypos, j);
      turn += (((peers >> j) & 1u) && pj < mypos) ? 1 : 0;
    }
  }
  const int last = __reduce_max_sync(MG_FULL, active ? turn : 0);
  for (int rd = 0; rd <= last; rd++) {
    if (active && turn == rd) run_action<PLAIN>(w, s, lane, stream, idx, kind, arg);
    __syncwarp();
  }
}

// Agent::apply_on_tick (objects/agent.cpp:63-67) of one agent
__device__ __forceinline__ void on_tick_one(const Wv& w, const Smem& s, int a) {
  const int slot = (int)s.a_slot[a];
  const int h = __ldg(tmpl(w, o_tmpl(objp(w, slot))) + MGT_ON_TICK);
  if (h >= 0) {
    Ctx c = make_ctx();
    c.actor = c.target = slot;
    handler_apply(w, MG_DEPTH, h, c);
  }
}

template <bool PLAIN>
__global__ void __launch_bounds__(MG_WARPS_PER_CTA * 32, MG_MIN_CTAS_PER_SM) k_world(MgDev d) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MG_PHASE_BEGIN();
  {
    Smem cta;
    carve(d, smem_raw, warp, cta);
    load_cta_tables(d, cta);
  }
  const int env = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
  if (env >= d.num_envs) return;
  Smem s0;
  carve(d, smem_raw, warp, s0);
  publish_warp(d, s0, env, lane);
  const Smem& s = *s0.self;
  Wv& w = *s0.wv;
  if (lane == 0) w.step += 1;  // :951
  __syncwarp();
  const int A = w.A;
  const size_t g0 = (size_t)env * A;

  // phase 1-2: stage state, read actions (:929-944; the obs buffer is rewritten whole below)
  stage_cells(d, w, lane);
  load_agents(w, s, lane);
  for (int a = lane; a < A; a += 32) {
    s.a_act[a] = d.actions[g0 + a];
    s.a_vact[a] = d.vibe_actions[g0 + a];
    s.a_res[a] = 0;
    s.a_exec[a] = 0;
    s.a_order[a] = (uint16_t)a;
  }
  rng_window_fill(w, lane);
  __syncwarp();
  MG_PHASE(0);

  // phase 4-5: shuffled action resolution (:958-999): by components of independent agents when every agent has a
  // lane, else (or when a SERIAL-class action shows up in a pass) in the reference's order on lane 0
  {
    const bool par = A <= 32 && (w.hdr[MGH_PAR_FLAGS] & MGP_ACTIONS);
    if (lane == 0) rng_shuffle(w, s.a_order, A);
    __syncwarp();
    int mypos = 0xffff;
    if (par) {
      if (lane < A) s.a_pos[s.a_order[lane]] = (uint16_t)lane;
      __syncwarp();
      if (lane < A) mypos = s.a_pos[lane];
    }
    const int maxp = w.hdr[MGH_MAX_PRIORITY];
    const int pmask = w.hdr[MGH_PRIORITY_MASK];
    for (int off = 0; off <= maxp; off++) {
      const int prio = maxp - off;
      if (!((pmask >> prio) & 1)) continue;  // no action has this priority: the pass only re-reports invalid indices
      for (int stream = 0; stream < 2; stream++) {
        if (par) {
          action_pass<PLAIN>(w, s, lane, stream, prio, mypos);
        } else {
          if (lane == 0) action_pass_serial<PLAIN>(w, s, stream, prio);
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  MG_PHASE(1);

  // per-agent bookkeeping of handle_action (actions/action_handler.hpp:78-105), one lane per agent.
  // It only touches the acting agent's own counters, so it is order-independent across agents.
  {
    const int NA = w.hdr[MGH_NUM_ACTIONS];
    const int npass = w.hdr[MGH_MAX_PRIORITY] + 1;
    for (int a = lane; a < A; a += 32) {
      uint32_t* ag = w.agents + a * w.AS;
      uint32_t prev = ag[MGAG_PREV_LOC], swm = ag[MGAG_SWM];
      const uint32_t res = s.a_res[a];
      const int ia = s.a_act[a], iv = s.a_vact[a];
      const bool inv_p = ia < 0 || ia >= NA, inv_v = iv < 0 || iv >= NA;
      bool success = false;
      for (int stream = 0; stream < 2; stream++) {
        const bool acted = res & (stream ? MGR_ACTED_V : MGR_ACTED_P);
        const bool ok = res & (stream ? MGR_OK_V : MGR_OK_P);
        if ((stream ? inv_v : inv_p)) {  // _handle_invalid_action, once per priority pass
          astat_add(w, a, w.hdr[MGH_ST_INVALID_INDEX], (float)npass);
          success = false;
        }
        if (!acted) continue;
        const uint32_t loc = stream ? s.a_locv[a] : s.a_locp[a];
        if (loc == prev) {
          swm += 1;
          if ((float)swm > astat_get(w, a, w.hdr[MGH_ST_MAX_SWM])) astat_set(w, a, w.hdr[MGH_ST_MAX_SWM], (float)swm);
        } else {
          swm = 0;
        }
        prev = loc;
        const int kind = (res >> (stream ? MGR_KIND_V_SHIFT : MGR_KIND_P_SHIFT)) & 3;
        if (ok) {
          astat_add(w, a, w.hdr[MGH_ST_NOOP_SUCCESS + 2 * kind], 1.0f);
          success = true;
        } else {
          astat_add(w, a, w.hdr[MGH_ST_NOOP_FAILED + 2 * kind], 1.0f);
          astat_add(w, a, w.hdr[MGH_ST_ACTION_FAILED], 1.0f);
        }
      }
      ag[MGAG_PREV_LOC] = prev;
      ag[MGAG_SWM] = swm;
      d.success[g0 + a] = success;
    }
  }
  __syncwarp();
  MG_PHASE(2);

  // phases 6-11 (:1009-1052).  Events and the game on_tick handler are serial.  The per-agent systems -- agent
  // on_tick, fixed AOE, territory handlers, mobile AOE -- run one lane per agent when the compiler proved that they
  // only write their own agent and read nothing another lane writes (MGH_PAR_FLAGS), else on lane 0 in agent order.
  if (!PLAIN) {
    const int parf = w.hdr[MGH_PAR_FLAGS];
    if (lane == 0 && w.hdr[MGH_NUM_EVENTS_SCHED] > 0) process_events(w);  // :1009-1011
    __syncwarp();
    if (parf & MGP_ON_TICK) {  // :1019-1024
      for (int a = lane; a < A; a += 32) on_tick_one(w, s, a);
    } else if (lane == 0) {
      for (int a = 0; a < A; a++) on_tick_one(w, s, a);
    }
    __syncwarp();
    MG_PHASE(3);
    const bool have_aoe = w.E[MGEV_NUM_AOE] > 0;
    const bool any_aoe = have_aoe || w.NTERR > 0 || w.E[MGEV_NUM_AOE_PENDING] > 0;
    if (any_aoe && (parf & MGP_AOE)) {
      terr_refresh(w, lane);  // events or actions may have moved a territory source
      for (int a = lane; a < A; a += 32) {  // :1032-1042
        if (have_aoe) aoe_apply_fixed(w, a, nullptr);
        if (w.NTERR > 0) terr_apply(w, a);
        if (have_aoe) aoe_apply_mobile_agent(w, a);
      }
      __syncwarp();
      if (lane == 0) aoe_flush_deferred(w);
    } else if (any_aoe) {
      // serial form: before each agent's turn all lanes test which sources can matter to it at its current cell
      for (int a = 0; a < A; a++) {
        uint32_t m[4];
        const bool masked = have_aoe && aoe_relevant_mask(w, a, lane, m);
        if (lane == 0) {
          if (have_aoe) aoe_apply_fixed(w, a, masked ? m : nullptr);
          if (w.NTERR > 0) terr_apply(w, a);
        }
        __syncwarp();
      }
      if (lane == 0) {
        aoe_apply_mobile(w);
        aoe_flush_deferred(w);
      }
    }
    if (lane == 0) {
      const int gh = w.hdr[MGH_GAME_ON_TICK];
      if (gh >= 0) {
        Ctx c = make_ctx();
        handler_apply(w, MG_DEPTH, gh, c);
      }
    }
  }
  __syncwarp();
  MG_PHASE(4);

  // phase 12: coverage (objects/agent.cpp:49-57), one lane per agent
  for (int a = lane; a < A; a += 32) {
    uint32_t* ag = w.agents + a * w.AS;
    const uint32_t loc = objp(w, (int)s.a_slot[a])[MGO_LOC];
    s.a_loc[a] = loc;
    const int r = (int)(loc >> 16), c = (int)(loc & 0xffffu);
    const int cell = r * w.W + c;
    uint32_t* cv = w.cover + a * d.CW + (cell >> 5);
    const uint32_t bit = 1u << (cell & 31);
    uint32_t unique = ag[MGAG_UNIQUE];
    if (!(*cv & bit)) {
      *cv |= bit;
      ag[MGAG_UNIQUE] = ++unique;
    }
    astat_set(w, a, w.hdr[MGH_ST_UNIQUE_VISITED], (float)unique);
    const uint32_t sp = ag[MGAG_SPAWN];
    const uint32_t dist = (uint32_t)(abs((int)(sp >> 16) - r) + abs(c - (int)(sp & 0xffffu)));
    const uint32_t md = max(ag[MGAG_MAX_DIST], dist);
    ag[MGAG_MAX_DIST] = md;
    astat_set(w, a, w.hdr[MGH_ST_MAX_DIST], (float)md);
  }
  __syncwarp();

  MG_PHASE(5);
  // phases 13-15 (observations, rewards, truncation) run in k_observe (one warp per agent) and k_finish
  rng_window_commit(w, lane);
  if (lane == 0) w.E[MGEV_STEP] = (int32_t)w.step;
  hand_over(d, w, s, env, lane, false);
  MG_PHASE(7);
}

// =================================================================================================
// k_finish: what the reference does after the observation pass (:640-661,1062-1096), one warp per env: the env's
// token stats in agent order, rewards (systems/reward.hpp:56-77), episode rewards, truncation / termination.
// =================================================================================================
template <bool PLAIN>
__global__ void __launch_bounds__(MG_WARPS_PER_CTA * 32, MG_FINISH_CTAS_PER_SM) k_finish(MgDev d, const uint8_t* __restrict__ mask, int initial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (mask) {
    const int e = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
    if (!__syncthreads_or(e < d.num_envs && mask[e])) return;
  }
  {
    Smem cta;
    carve(d, smem_raw, warp, cta);
    load_cta_tables(d, cta);
  }
  const int env = d.env0 + blockIdx.x * MG_WARPS_PER_CTA + warp;
  if (env >= d.num_envs) return;
  if (mask && !mask[env]) return;
  Smem s0;
  carve(d, smem_raw, warp, s0);
  publish_warp(d, s0, env, lane);
  Wv& w = *s0.wv;
  const Smem& s = *s0.self;
  const int A = w.A;
  const size_t g0 = (size_t)env * A;
  if (!initial) {
    // cell staleness (:787-796): every object some agent saw this tick was claimed by the lowest such agent
    // (k_observe); advance its stamp and give the ticks since to that agent's cell.visited -- one float add per agent
    for (int i = lane; i < A; i += 32) s.a_res[i] = 0;
    __syncwarp();
    const unsigned long long* claims = d.claims + (size_t)env * d.maxobj;
    uint32_t* visited = d.visited + (size_t)env * d.maxobj;  // GridObject::visited, kept apart from the records: this scan is coalesced
    const int nobj = w.E[MGEV_NEXT_OBJ];
    const uint32_t step = w.step;
    for (int s0_ = 1; s0_ < nobj; s0_ += 128) {  // four independent claim loads per lane in flight
      unsigned long long c[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int sl = s0_ + 32 * q + lane;
        c[q] = sl < nobj ? claims[sl] : 0ull;
      }
      uint32_t vis[4];
#pragma unroll
      for (int q = 0; q < 4; q++) {
        vis[q] = step;
        if ((uint32_t)(c[q] >> 32) == step) vis[q] = visited[s0_ + 32 * q + lane];
      }
#pragma unroll
      for (int q = 0; q < 4; q++) {
        if (vis[q] < step) {
          visited[s0_ + 32 * q + lane] = step;
          const uint32_t ag = 0xffffffffu - (uint32_t)c[q];
          if (ag < (uint32_t)A) atomicAdd(&s.a_res[ag], step - vis[q]);
        }
      }
    }
    __syncwarp();
    for (int i = lane; i < A; i += 32)
      if (s.a_res[i]) astat_add(w, i, w.hdr[MGH_ST_CELL_VISITED], (float)s.a_res[i]);
    __syncwarp();
  }
  {
    // token stats: one float add per agent in agent order like the reference (:659-661).  The counts are loaded 32 at a
    // time, one per lane, and handed to lane 0 by shuffles -- not one dependent load per agent.
    const int idw = w.hdr[MGH_GST_TOKENS_WRITTEN], idf = w.hdr[MGH_GST_TOKENS_FREE];
    float tw = 0.0f, tf = 0.0f;
    if (lane == 0) tw = w.gstats[idw], tf = w.gstats[idf];
    for (int base = 0; base < A; base += 32) {
      const int mine = base + lane < A ? d.tok_attempted[g0 + base + lane] : 0;
      const int n = min(32, A - base);
      for (int j = 0; j < n; j++) {
        const int attempted = __shfl_sync(MG_FULL, mine, j);
        if (lane != 0) continue;
        if (attempted > w.T) {  // hard error in the reference (:364-375)
          set_error(w, MGERR_TOKEN_OVERFLOW, (base + j) | (min(attempted, 65535) << 16));
        } else {
          tw = __fadd_rn(tw, (float)attempted);
          tf = __fadd_rn(tf, (float)(w.T - attempted));
        }
      }
    }
    if (lane == 0) {
      w.gstats[idw] = tw;
      w.gstats[idf] = tf;
      gstat_touch(w, idw);
      gstat_touch(w, w.hdr[MGH_GST_TOKENS_DROPPED]);
      gstat_touch(w, idf);
    }
  }
  if (initial) return;
  __syncwarp();
  // rewards (systems/reward.hpp:56-77), episode rewards, truncation (:1070-1096), one lane per agent
  const int ms = w.hdr[MGH_MAX_STEPS];
  const bool done = ms > 0 && w.step >= (uint32_t)ms;
  for (int a = lane; a < A; a += 32) {
    uint32_t* ag = w.agents + a * w.AS;
    float reward = 0.0f;
    if (!PLAIN) {
      const int slot = (int)ag[MGAG_OBJ];
      const int32_t* tp = tmpl(w, o_tmpl(objp(w, slot)));
      const int nr = __ldg(tp + MGT_REWARDS_N);
      if (nr > 0) {
        const int32_t* rw = pool(w, __ldg(tp + MGT_REWARDS));
        float total = 0.0f;
        Ctx rc = make_ctx();
        rc.actor = rc.target = (uint16_t)slot;
        for (int i = 0; i < nr; i++) {
          const float v = eval_value(w, MG_DEPTH, __ldg(rw + 2 * i), rc, slot);
          const float prevv = __uint_as_float(ag[MGAG_REWARD_PREV + i]);
          total = __ldg(rw + 2 * i + 1) ? __fadd_rn(total, v) : __fadd_rn(total, __fsub_rn(v, prevv));
          ag[MGAG_REWARD_PREV + i] = __float_as_uint(v);
        }
        if (total != 0.0f) reward = __fadd_rn(0.0f, total);
      }
    }
    d.rewards[g0 + a] = reward;
    ag[MGAG_EPISODE_REWARD] = __float_as_uint(__fadd_rn(__uint_as_float(ag[MGAG_EPISODE_REWARD]), reward));
    if (done) {
      if (w.hdr[MGH_EPISODE_TRUNCATES])
        d.truncations[g0 + a] = 1;
      else
        d.terminals[g0 + a] = 1;
    }
  }
}

// Agent::set_inventory (objects/agent.cpp:86-104) for one agent of one env: a single warp, lane 0 works
__global__ void k_set_inventory(MgDev d, int env, int agent, const int32_t* __restrict__ items,
                                const int32_t* __restrict__ amounts, int n) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  {
    Smem cta;
    carve(d, smem_raw, 0, cta);
    load_cta_tables(d, cta);
  }
  Smem s0;
  carve(d, smem_raw, 0, s0);
  publish_warp(d, s0, env, lane);
  Wv& w = *s0.wv;
  if (lane != 0) return;
  uint32_t* o = objp(w, (int)w.agents[agent * w.AS + MGAG_OBJ]);
  const uint64_t existing = o_order(o);
  for (int i = 0; i < ord_count(existing); i++) {
    const int it = ord_item(existing, i);
    inv_update(w, 2, o, it, -(int)o_inv(w, o)[it]);
    astat_set(w, agent, __ldg(sec(w, MGS_RES_STATS) + it * 4 + 2), 0.0f);
  }
  for (int i = 0; i < n; i++) inv_update(w, 2, o, items[i], amounts[i] - (int)o_inv(w, o)[items[i]]);
}

}  // namespace

cudaError_t mg_launch_set_inventory(const MgDev& d, int env, int agent, const int32_t* items, const int32_t* amounts, int n,
                                    cudaStream_t st) {
  size_t bytes = smem_per_cta(d.NOFF) + smem_per_warp(d.stage_grid ? d.HWp : 0, d.T, d.A);
  cudaError_t e = cudaFuncSetAttribute(k_set_inventory, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return e;
  k_set_inventory<<<1, 32, bytes, st>>>(d, env, agent, items, amounts, n);
  return cudaGetLastError();
}

// flags / rewards of a second buffer set (mg_step_host's staging) after a reset of the selected envs
__global__ void k_clear_outputs(int num_envs, int A, const uint8_t* __restrict__ mask, uint8_t* terminals, uint8_t* truncations,
                                float* rewards) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)num_envs * A || (mask && !mask[i / A])) return;
  terminals[i] = 0, truncations[i] = 0, rewards[i] = 0.0f;
}
cudaError_t mg_launch_clear_outputs(const MgDev& d, const uint8_t* mask, cudaStream_t st) {
  const size_t n = (size_t)d.num_envs * d.A;
  k_clear_outputs<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d.num_envs, d.A, mask, d.terminals, d.truncations, d.rewards);
  return cudaGetLastError();
}

// ---- host-side launchers (used by mg_capi.cu) --------------------------------------------------
size_t mg_smem_bytes(const MgDev& d) { return smem_per_cta(d.NOFF) + (size_t)MG_WARPS_PER_CTA * smem_per_warp(d.stage_grid ? d.HWp : 0, d.T, d.A); }

template <class K>
static cudaError_t set_smem(K k, size_t bytes) { return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes); }

cudaError_t mg_configure_kernels(const MgDev& d) {
  const size_t bytes = mg_smem_bytes(d), obytes = obs_smem_bytes(MG_OBS_MAX_WARPS, d.T);
  cudaError_t e;
  if ((e = set_smem(k_reset, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_init_buffers, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_world<false>, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_world<true>, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_finish<false>, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_finish<true>, bytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_observe<false, 4>, obytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_observe<false, 8>, obytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_observe<true, 4>, obytes)) != cudaSuccess) return e;
  if ((e = set_smem(k_observe<true, 8>, obytes)) != cudaSuccess) return e;
  // leave room for MG_MIN_CTAS_PER_SM CTAs' shared memory; the rest of the 256 KB stays L1
  int want_kb = (int)((bytes + 1024) * MG_MIN_CTAS_PER_SM / 1024) + 8;
  int pct = want_kb * 100 / 228 + 1;
  if (pct > 100) pct = 100;
  if ((e = cudaFuncSetAttribute(k_world<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_world<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct)) != cudaSuccess) return e;
  want_kb = (int)((bytes + 1024) * MG_FINISH_CTAS_PER_SM / 1024) + 8;
  pct = want_kb * 100 / 228 + 1;
  if (pct > 100) pct = 100;
  if ((e = cudaFuncSetAttribute(k_finish<true>, cudaFuncAttributePreferredSharedMemoryCarveout, pct)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_finish<false>, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}
// a launch covers the envs [d.env0, d.num_envs): the whole handle, or one chunk of it (mg_capi.cu: launch_step)
static inline int mg_grid(const MgDev& d) { return (d.num_envs - d.env0 + MG_WARPS_PER_CTA - 1) / MG_WARPS_PER_CTA; }
// the observation pass + what follows it, for a tick (initial = 0) or for _init_buffers (initial = 1)
static int obs_warps(int A) {  // agents per CTA: all of them, or the largest divisor of A that fits, so no warp idles
  if (A <= MG_OBS_MAX_WARPS) return A;
  for (int wpc = MG_OBS_MAX_WARPS; wpc >= 4; wpc--)
    if (A % wpc == 0) return wpc;
  return MG_OBS_MAX_WARPS;
}
static cudaError_t launch_observe_finish(const MgDev& d, const uint8_t* mask, int initial, cudaStream_t st) {
  const int wpc = obs_warps(d.A);
  const dim3 grid((unsigned)(d.num_envs - d.env0), (unsigned)((d.A + wpc - 1) / wpc));
  const size_t ob = obs_smem_bytes(wpc, d.T);
  const bool plain = d.plain || (d.NTERR == 0 && d.OVW == 0);
  const MgFastHdr& H = *d.hdr_host;
  if (d.NOFF <= 128) {
    if (plain) k_observe<true, 4><<<grid, wpc * 32, ob, st>>>(d, H, mask);
    else k_observe<false, 4><<<grid, wpc * 32, ob, st>>>(d, H, mask);
  } else {
    if (plain) k_observe<true, 8><<<grid, wpc * 32, ob, st>>>(d, H, mask);
    else k_observe<false, 8><<<grid, wpc * 32, ob, st>>>(d, H, mask);
  }
  if (d.plain)
    k_finish<true><<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d, mask, initial);
  else
    k_finish<false><<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d, mask, initial);
  return cudaGetLastError();
}
cudaError_t mg_launch_reset(const MgDev& d, const uint8_t* mask, cudaStream_t st) {
  k_reset<<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d, mask);
  return cudaGetLastError();
}
cudaError_t mg_launch_init_buffers(const MgDev& d, const uint8_t* mask, cudaStream_t st) {
  k_init_buffers<<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d, mask);
  return launch_observe_finish(d, mask, 1, st);
}
cudaError_t mg_launch_step(const MgDev& d, cudaStream_t st) {
  if (d.plain)
    k_world<true><<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d);
  else
    k_world<false><<<mg_grid(d), MG_WARPS_PER_CTA * 32, mg_smem_bytes(d), st>>>(d);
  return launch_observe_finish(d, nullptr, 0, st);
}
