// mg_device.cuh -- per-environment device state view and the game-program interpreter.
//
// One WARP owns one environment for a whole tick.  The reference's conflict
// semantics are sequential per env (bindings/mettagrid_c.cpp:958-999); the interpreter below is free of
// warp intrinsics, so the step kernel may run it on lane 0 only (what must keep the reference's order)
// or on one lane per agent (what the compiler's effect analysis proved independent).  Nothing here is
// shared with oracle/.
//
// The interpreter functions call each other recursively (handler -> mutation -> handler, filter ->
// query -> filter, value -> query ...).  It is REAL recursion: every function exists once and takes the
// remaining nesting budget D as an argument (MG_DEPTH at the roots); a program that nests deeper sets
// MGERR_UNSUPPORTED instead of misbehaving.  Round 1 instantiated every function once per nesting level
// (template depth): 83 k instructions for k_step, 26 % of the stall samples waiting for instructions.
// The stack is sized by mg_create (cudaLimitStackSize, MG_STACK_BYTES).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "mg_state.h"

#ifndef MG_DEPTH
#define MG_DEPTH 8
#endif

// Per-warp view of one environment (lives in shared memory).
struct Wv {
  const int32_t* P;
  const int32_t* hdr;  // shared-memory copy of the header
  uint16_t* cells;     // shared-memory staged grid
  uint16_t* cells_g;
  uint32_t* objs;
  uint32_t* agents;
  float* astats;
  uint32_t* atouched;
  float* gstats;
  uint32_t* gtouched;
  uint32_t* cover;
  uint32_t* rng;
  int32_t* E;
  const float* logtab;
  uint16_t* arena;
  uint32_t* aoe_src;
  int32_t* aoe_pending;
  uint32_t* terr_src;
  int16_t* inside_tag;
  uint32_t* dyn_stamp;
  uint16_t* tag_lists;
  int32_t* tag_state;
  uint32_t* terr_tab;  // per-tick table of scoring territory sources (mg_world.cuh)
  uint8_t* owner;      // [NTERR][H*W] winning prefix index + 1 per cell and territory (0 = nobody), mg_world.cuh
  uint32_t step;
  int H, W, A, R, TW, OS, AS, SA, SAW, T, B, ND, NOFF, CW, maxobj;
  int ARENA, AOECAP, AOEW, PENDCAP, TERRCAP, NDYN, NTERR, NPROXY, PAD, WP, TOKOFF, NTAGS;
  uint8_t* obs;  // this env's observation rows [A][T][3]
  // rng window + arena top (shared memory): [0]=consumed, [1]=count, [2]=direct mode, [3]=idx0, [4]=arena top
  int* rs;
  uint32_t* rand;
};

// deferred per-target ResourceDelta accumulator of AOETracker::apply_fixed (aoe_tracker.cpp:282-361)
struct Deferred {
  int n;
  int8_t order[16];
  int32_t delta[16];
  uint16_t seen;
};

// handler/handler_context.hpp:34-55
// Packed into 24 bytes: contexts live in local memory (they are passed by reference through the non-inlined
// interpreter) and one is built per handler invocation.  Object slots are < 65536 (mg_create), coordinates fit
// int16 (a line scan may step one cell outside the map).
struct alignas(8) Ctx {
  uint16_t actor, target, source;
  int16_t distance, tr, tc;
  int8_t move_dir;
  bool skip_trigger, failed;
  uint8_t pad_;
  Deferred* deferred;
};
static_assert(sizeof(Ctx) == 24, "Ctx is built and copied as three 8-byte words");
__device__ __forceinline__ Ctx make_ctx() {  // two 8-byte zero stores + the pointer instead of one store per field
  Ctx c;
  unsigned long long* q = reinterpret_cast<unsigned long long*>(&c);
  q[0] = 0ull, q[1] = 0ull;
  c.deferred = nullptr;
  return c;
}

// MG_CHECKED=1 (python -m mettagrid_b200.build --variant checked -DMG_CHECKED=1): every state accessor verifies its
// index and reports MGERR_BOUNDS through the env's error word, which the Python wrapper raises -- the pool's
// compute-sanitizer is closed, so the whole GPU test suite is run once against this build instead (profiles/README.md).
#ifndef MG_CHECKED
#define MG_CHECKED 0
#endif
#define MG_CHECK(w, cond, info)                                                   \
  do {                                                                            \
    if (MG_CHECKED && !(cond)) {                                                  \
      if (!((w).E[MGEV_ERROR] & MGERR_BOUNDS)) (w).E[MGEV_ERR_INFO] = (info);     \
      atomicOr(&(w).E[MGEV_ERROR], MGERR_BOUNDS);                                 \
    }                                                                             \
  } while (0)

// ---- program access -------------------------------------------------------------------------
__device__ __forceinline__ const int32_t* sec(const Wv& w, int k) { return w.P + w.hdr[k]; }
__device__ __forceinline__ const int32_t* pool(const Wv& w, int off) { return w.P + w.hdr[MGS_POOL] + off; }
__device__ __forceinline__ const int32_t* tmpl(const Wv& w, int t) { return sec(w, MGS_TEMPLATES) + t * MG_TEMPLATE_WORDS; }

// ---- object records ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t* objp(const Wv& w, int s) {
  MG_CHECK(w, (unsigned)s < (unsigned)(w.maxobj + w.NPROXY), 1000);
  if (MG_CHECKED && (unsigned)s >= (unsigned)(w.maxobj + w.NPROXY)) s = 0;
  return w.objs + (size_t)s * w.OS;
}
__device__ __forceinline__ int o_r(const uint32_t* o) { return (int)(o[MGO_LOC] >> 16); }
__device__ __forceinline__ int o_c(const uint32_t* o) { return (int)(o[MGO_LOC] & 0xffffu); }
__device__ __forceinline__ int o_tmpl(const uint32_t* o) { return (int)(o[MGO_META] & 0xffffu); }
__device__ __forceinline__ int o_vibe(const uint32_t* o) { return (int)((o[MGO_META] >> 16) & 0xffu); }
__device__ __forceinline__ int o_flags(const uint32_t* o) { return (int)(o[MGO_META] >> 24); }
__device__ __forceinline__ bool o_is_agent(const uint32_t* o) { return (o_flags(o) & MGOF_AGENT) != 0; }
__device__ __forceinline__ bool o_alive(const uint32_t* o) { return (o_flags(o) & MGOF_ALIVE) != 0; }
__device__ __forceinline__ int o_agent(const uint32_t* o) { return (int)o[MGO_AGENT]; }
__device__ __forceinline__ void o_set_vibe(uint32_t* o, int v) {
  o[MGO_META] = (o[MGO_META] & 0xff00ffffu) | ((uint32_t)(v & 0xff) << 16);
  o[MGO_NTOK] = MG_TOK_DIRTY;
}
__device__ __forceinline__ uint16_t* o_inv(const Wv& w, uint32_t* o) { return (uint16_t*)(o + MGO_TAGS + w.TW); }
__device__ __forceinline__ bool o_has_tag(const uint32_t* o, int t) { return t >= 0 && t < 256 && ((o[MGO_TAGS + (t >> 5)] >> (t & 31)) & 1u); }

// inventory iteration order: packed 4-bit ids, most recently inserted first
__device__ __forceinline__ uint64_t o_order(const uint32_t* o) { return (uint64_t)o[MGO_INVORD_LO] | ((uint64_t)o[MGO_INVORD_HI] << 32); }
__device__ __forceinline__ void o_set_order(uint32_t* o, uint64_t v) {
  o[MGO_INVORD_LO] = (uint32_t)v;
  o[MGO_INVORD_HI] = (uint32_t)(v >> 32);
}
__device__ __forceinline__ int ord_count(uint64_t v) { return (int)(v >> 60); }
__device__ __forceinline__ int ord_item(uint64_t v, int i) { return (int)((v >> (4 * i)) & 15u); }
__device__ __forceinline__ uint64_t ord_push_front(uint64_t v, int item) {
  uint64_t n = (uint64_t)(ord_count(v) + 1);
  uint64_t body = (v & 0x0fffffffffffffffull) << 4 | (uint64_t)item;
  return (body & 0x0fffffffffffffffull) | (n << 60);
}
__device__ __forceinline__ uint64_t ord_erase(uint64_t v, int item) {
  int n = ord_count(v);
  uint64_t body = v & 0x0fffffffffffffffull;
  for (int i = 0; i < n; i++)
    if (ord_item(body, i) == item) {
      uint64_t low = body & ((1ull << (4 * i)) - 1);
      uint64_t high = (body >> (4 * (i + 1))) << (4 * i);
      return (low | high) | ((uint64_t)(n - 1) << 60);
    }
  return v;
}

// ---- stats (systems/stats_tracker.hpp:57-98); lane-exclusive per (agent) or serial ----------
__device__ __forceinline__ void astat_touch(const Wv& w, int a, int id) {
  MG_CHECK(w, (unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA, 1001);
  if (MG_CHECKED && !((unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA)) return;
  uint32_t* p = w.atouched + a * w.SAW + (id >> 5);
  const uint32_t bit = 1u << (id & 31);
  if (!(*p & bit)) *p |= bit;
}
__device__ __forceinline__ void astat_add(const Wv& w, int a, int id, float v) {
  MG_CHECK(w, (unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA, 1002);
  if (MG_CHECKED && !((unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA)) return;
  float* p = w.astats + a * w.SA + id;
  *p = __fadd_rn(*p, v);
  astat_touch(w, a, id);
}
__device__ __forceinline__ void astat_set(const Wv& w, int a, int id, float v) {
  MG_CHECK(w, (unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA, 1003);
  if (MG_CHECKED && !((unsigned)a < (unsigned)w.A && (unsigned)id < (unsigned)w.SA)) return;
  w.astats[a * w.SA + id] = v;
  astat_touch(w, a, id);
}
__device__ __forceinline__ float astat_get(const Wv& w, int a, int id) { return w.astats[a * w.SA + id]; }
__device__ __forceinline__ void gstat_touch(const Wv& w, int id) {
  MG_CHECK(w, id >= 0, 1005);
  if (MG_CHECKED && id < 0) return;
  uint32_t* p = w.gtouched + (id >> 5);
  const uint32_t bit = 1u << (id & 31);
  if (!(*p & bit)) atomicOr(p, bit);  // agents of one env may report from different lanes
}
__device__ __forceinline__ void gstat_add(const Wv& w, int id, float v) {
  float* p = w.gstats + id;
  *p = __fadd_rn(*p, v);
  gstat_touch(w, id);
}
__device__ __forceinline__ float gstat_get(const Wv& w, int id) { return w.gstats[id]; }
__device__ __forceinline__ void gstat_set(const Wv& w, int id, float v) {
  w.gstats[id] = v;
  gstat_touch(w, id);
}

__device__ __forceinline__ void set_error(const Wv& w, int code, int info) {
  if (!(w.E[MGEV_ERROR] & code)) w.E[MGEV_ERR_INFO] = info;
  atomicOr(&w.E[MGEV_ERROR], code);  // several lanes of an env may report at once
}

// ---- MT19937.  Outputs are identical to std::mt19937: the generator is advanced
// one element at a time instead of 624 at once, which yields the same stream. -----------------
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
// all lanes: precompute the next <=32 outputs without committing them.
// rand[0..32) = tempered outputs, rand[32..64) = the new state words they came from.
__device__ __forceinline__ void rng_window_fill(const Wv& w, int lane) {
  int idx0 = w.E[MGEV_RNG_IDX];
  if (idx0 >= MG_RNG_WORDS) idx0 = 0;
  int count = min(MG_RNG_WINDOW, MG_RNG_WORDS - idx0);
  if (lane < count) {
    int i = idx0 + lane;
    int i1 = i + 1 == MG_RNG_WORDS ? 0 : i + 1;
    int i2 = i + 397 >= MG_RNG_WORDS ? i + 397 - MG_RNG_WORDS : i + 397;
    uint32_t nw = mt_twist(w.rng[i], w.rng[i1], w.rng[i2]);
    w.rand[lane] = mt_temper(nw);
    w.rand[MG_RNG_WINDOW + lane] = nw;
  }
  if (lane == 0) {
    w.rs[0] = 0;
    w.rs[1] = count;
    w.rs[2] = 0;
    w.rs[3] = idx0;
  }
}
// all lanes: commit what the serial code consumed from the window
__device__ __forceinline__ void rng_window_commit(const Wv& w, int lane) {
  if (w.rs[2]) return;  // direct mode already committed
  int used = w.rs[0];
  if (lane < used) w.rng[w.rs[3] + lane] = w.rand[MG_RNG_WINDOW + lane];
  if (lane == 0 && used > 0) w.E[MGEV_RNG_IDX] = w.rs[3] + used;
}
// serial: next raw 32-bit output once the window is exhausted
__device__ __noinline__ uint32_t rng_next_slow(const Wv& w) {
  if (!w.rs[2]) {  // leave window mode: commit the whole window, continue directly on global state
    int count = w.rs[1];
    for (int k = 0; k < count; k++) w.rng[w.rs[3] + k] = w.rand[MG_RNG_WINDOW + k];
    w.E[MGEV_RNG_IDX] = w.rs[3] + count;
    w.rs[2] = 1;
  }
  int i = w.E[MGEV_RNG_IDX];
  if (i >= MG_RNG_WORDS) i = 0;
  int i1 = i + 1 == MG_RNG_WORDS ? 0 : i + 1;
  int i2 = i + 397 >= MG_RNG_WORDS ? i + 397 - MG_RNG_WORDS : i + 397;
  uint32_t nw = mt_twist(w.rng[i], w.rng[i1], w.rng[i2]);
  w.rng[i] = nw;
  w.E[MGEV_RNG_IDX] = i + 1;
  return mt_temper(nw);
}
__device__ __forceinline__ uint32_t rng_next(const Wv& w) {
  int pos = w.rs[0];
  if (!w.rs[2] && pos < w.rs[1]) {
    w.rs[0] = pos + 1;
    return w.rand[pos];
  }
  return rng_next_slow(w);
}
// Lemire multiply-shift with rejection (bits/uniform_int_dist.h:252-282)
__device__ __forceinline__ uint32_t rng_below(const Wv& w, uint32_t range) {
  uint64_t prod = (uint64_t)rng_next(w) * range;
  uint32_t low = (uint32_t)prod;
  if (low < range) {
    uint32_t thr = (0u - range) % range;
    while (low < thr) {
      prod = (uint64_t)rng_next(w) * range;
      low = (uint32_t)prod;
    }
  }
  return (uint32_t)(prod >> 32);
}
// libstdc++ std::shuffle (bits/stl_algo.h:3719-3805): two swap positions per draw
template <class T>
__device__ __forceinline__ void rng_shuffle(const Wv& w, T* v, int n) {
  if (n < 2) return;
  int i = 1;
  if ((n & 1) == 0) {
    int j = (int)rng_below(w, 2);
    T t = v[i];
    v[i] = v[j];
    v[j] = t;
    i++;
  }
  while (i < n) {
    uint32_t sr = (uint32_t)i + 1;
    uint32_t x = rng_below(w, sr * (sr + 1));
    int p0 = (int)(x / (sr + 1)), p1 = (int)(x % (sr + 1));
    T t = v[i];
    v[i] = v[p0];
    v[p0] = t;
    i++;
    t = v[i];
    v[i] = v[p1];
    v[p1] = t;
    i++;
  }
}

// ---- inventory (objects/inventory.cpp:38-173, inventory.hpp:26-40) -----------------------------
__device__ __forceinline__ int limit_of(const Wv& w, const uint32_t* o, int item) { return __ldg(pool(w, __ldg(tmpl(w, o_tmpl(o)) + MGT_LIMIT_OF)) + item); }
__device__ __forceinline__ bool is_modifier(const Wv& w, const uint32_t* o, int item) { return (__ldg(tmpl(w, o_tmpl(o)) + MGT_MODIFIER_MASK) >> item) & 1; }
__device__ __forceinline__ int effective_limit(const Wv& w, uint32_t* o, int lim) {
  const int32_t* L = sec(w, MGS_LIMITS) + lim * MG_LIMIT_WORDS;
  const uint16_t* inv = o_inv(w, o);
  const int32_t* m = pool(w, __ldg(L + 2));
  int nm = __ldg(L + 3), sum = 0;
  for (int i = 0; i < nm; i++) sum += (int)inv[__ldg(m + 2 * i)] * __ldg(m + 2 * i + 1);
  int eff = min(__ldg(L + 1), max(__ldg(L + 0), sum));
  return min(max(eff, 0), 65535);
}
__device__ __forceinline__ int limit_amount(const Wv& w, uint32_t* o, int lim) {
  const int32_t* L = sec(w, MGS_LIMITS) + lim * MG_LIMIT_WORDS;
  const uint16_t* inv = o_inv(w, o);
  const int32_t* mem = pool(w, __ldg(L + 4));
  int n = __ldg(L + 5);
  uint32_t s = 0;
  for (int i = 0; i < n; i++) s += inv[__ldg(mem + i)];
  return (int)(s & 0xffffu);  // SharedInventoryLimit::amount is a uint16 running sum
}
// objects/agent.cpp:106-121
__device__ __forceinline__ void on_inventory_change(const Wv& w, uint32_t* o, int item, int delta) {
  if (!o_is_agent(o) || o_agent(o) < 0) return;
  int a = o_agent(o);
  const int32_t* rs = sec(w, MGS_RES_STATS) + item * 4;
  if (delta > 0)
    astat_add(w, a, __ldg(rs + 0), (float)delta);
  else
    astat_add(w, a, __ldg(rs + 1), (float)(-delta));
  int amount = o_inv(w, o)[item];
  astat_set(w, a, __ldg(rs + 2), (float)amount);
  if (amount == 0 && delta < 0 && item == w.hdr[MGH_HP_RESOURCE]) astat_add(w, a, w.hdr[MGH_ST_DEATH], 1.0f);
}
__device__ __noinline__ int inv_update(const Wv& w, int D, uint32_t* o, int item, int attempted, bool ignore_limits = false, bool notify = true);
__device__ __noinline__ void enforce_all_limits(const Wv& w, int D, uint32_t* o) {
  const int32_t* t = tmpl(w, o_tmpl(o));
  const int32_t* lo = pool(w, __ldg(t + MGT_LIMIT_ORDER));
  int nl = __ldg(t + MGT_LIMIT_ORDER_N);
  for (int i = 0; i < nl; i++) {
    int lim = __ldg(lo + i);
    int excess = limit_amount(w, o, lim) - effective_limit(w, o, lim);
    if (excess <= 0) continue;
    const int32_t* L = sec(w, MGS_LIMITS) + lim * MG_LIMIT_WORDS;
    const int32_t* mem = pool(w, __ldg(L + 4));
    int n = __ldg(L + 5);
    for (int k = 0; k < n; k++) {
      int it = __ldg(mem + k);
      int drop = min((int)o_inv(w, o)[it], excess);
      if (drop > 0) {
        if (D > 0)
          inv_update(w, D - 1, o, it, -drop);
        else
          set_error(w, MGERR_UNSUPPORTED, 1);
        excess = limit_amount(w, o, lim) - effective_limit(w, o, lim);
      }
      if (excess <= 0) break;
    }
  }
}
__device__ __noinline__ int inv_update(const Wv& w, int D, uint32_t* o, int item, int attempted, bool ignore_limits, bool notify) {
  uint16_t* inv = o_inv(w, o);
  int initial = inv[item];
  int na = initial + attempted;
  int mx = 65535;
  int lim = limit_of(w, o, item);
  if (!ignore_limits && lim >= 0) {
    int others = limit_amount(w, o, lim) - initial;
    if (others < 0) others = 0;
    int mi = effective_limit(w, o, lim) - others;
    mx = mi < 0 ? 0 : (mi & 0xffff);
  }
  int clamped = min(max(na, 0), mx);
  if (clamped == 0) {
    if (initial != 0) o_set_order(o, ord_erase(o_order(o), item));
  } else if (initial == 0) {
    o_set_order(o, ord_push_front(o_order(o), item));
  }
  inv[item] = (uint16_t)clamped;
  int d = clamped - initial;
  if (d != 0) o[MGO_NTOK] = MG_TOK_DIRTY;  // cached observation tokens are stale
  if (notify && d != 0) on_inventory_change(w, o, item, d);
  if (d < 0 && is_modifier(w, o, item)) enforce_all_limits(w, D, o);
  return d;
}
__device__ __forceinline__ int free_space(const Wv& w, uint32_t* o, int item) {
  int lim = limit_of(w, o, item);
  if (lim < 0) return 65535 - (int)o_inv(w, o)[item];
  int used = limit_amount(w, o, lim), eff = effective_limit(w, o, lim);
  return eff > used ? eff - used : 0;
}
// objects/has_inventory.cpp:76-108
__device__ __noinline__ int transfer(const Wv& w, uint32_t* src, uint32_t* dst, int item, int delta) {
  if (delta <= 0) return 0;
  int give = min((int)o_inv(w, src)[item], delta);
  int amount = min(give, free_space(w, dst, item));
  inv_update(w, 2, src, item, -amount);
  inv_update(w, 2, dst, item, amount);
  return amount;
}

// ---- grid (core/grid.hpp:31-130) -----------------------------------------------------------------
__device__ __forceinline__ bool valid_loc(const Wv& w, int r, int c) { return r >= 0 && c >= 0 && r < w.H && c < w.W; }
__device__ __forceinline__ int cidx(const Wv& w, int r, int c) {
  const int i = (r + w.PAD) * w.WP + c + w.PAD;
  MG_CHECK(w, (unsigned)i < (unsigned)((w.H + 2 * w.PAD) * w.WP), 1004);
  return (MG_CHECKED && (unsigned)i >= (unsigned)((w.H + 2 * w.PAD) * w.WP)) ? 0 : i;
}
__device__ __forceinline__ int cell_at(const Wv& w, int r, int c) { return w.cells[cidx(w, r, c)]; }
__device__ __forceinline__ void set_cell(const Wv& w, int r, int c, int s) {
  int i = cidx(w, r, c);
  w.cells[i] = (uint16_t)s;
  w.cells_g[i] = (uint16_t)s;  // write-through: the staged copy is never flushed
}
// the territory ownership map (mg_world.cuh) depends on where the territory sources stand and on their tags
__device__ __forceinline__ void terr_touch(const Wv& w, const uint32_t* o) {
  if (o_flags(o) & MGOF_TERR_SRC) w.E[MGEV_TERR_STALE] = 1;
}
__device__ __forceinline__ bool move_object(const Wv& w, int s, int r, int c) {
  if (!valid_loc(w, r, c) || cell_at(w, r, c) != 0) return false;
  uint32_t* o = objp(w, s);
  set_cell(w, r, c, s);
  set_cell(w, o_r(o), o_c(o), 0);
  o[MGO_LOC] = ((uint32_t)r << 16) | (uint32_t)c;
  terr_touch(w, o);
  return true;
}

// ---- tag index (core/tag_index.cpp:9-63): per tag, member slots in insertion order.  Kept only for
// programs with queries (NTAGS > 0); a tag with more than MG_TAG_LIST_CAP members is marked -1 and
// queried by scanning the object table instead.
__device__ __forceinline__ void tl_append(const Wv& w, int tag, int s) {
  if (tag >= w.NTAGS) return;
  int n = w.tag_state[tag];
  if (n < 0) return;
  if (n >= MG_TAG_LIST_CAP) {
    w.tag_state[tag] = -1;
    return;
  }
  w.tag_lists[tag * MG_TAG_LIST_CAP + n] = (uint16_t)s;
  w.tag_state[tag] = n + 1;
}
__device__ __forceinline__ void tl_erase(const Wv& w, int tag, int s) {
  if (tag >= w.NTAGS) return;
  int n = w.tag_state[tag];
  if (n <= 0) return;
  uint16_t* l = w.tag_lists + tag * MG_TAG_LIST_CAP;
  int k = 0;
  for (int i = 0; i < n; i++)
    if (l[i] != (uint16_t)s) l[k++] = l[i];
  w.tag_state[tag] = k;
}
__device__ __forceinline__ void tl_register_object(const Wv& w, int s, bool add) {
  if (w.NTAGS == 0) return;
  const uint32_t* o = objp(w, s);
  for (int k = 0; k < w.TW; k++) {
    uint32_t m = o[MGO_TAGS + k];
    while (m) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      if (add)
        tl_append(w, k * 32 + b, s);
      else
        tl_erase(w, k * 32 + b, s);
    }
  }
}

// ---- object creation (core/grid_object_factory.cpp:62-104); used by k_reset (one lane per object)
// and by the spawn mutations (serial) --------------------------------------------------------------
__device__ __forceinline__ void init_object(const Wv& w, int slot, int t, int r, int c, int aidx, bool obs_inv, uint32_t seq) {
  uint32_t* o = objp(w, slot);
  const int32_t* tp = tmpl(w, t);
  for (int k = 0; k < w.OS / 4; k++) ((uint4*)o)[k] = make_uint4(0, 0, 0, 0);  // records are 16-byte multiples (compiler.py)
  int kind = __ldg(tp + MGT_KIND);
  int flags = MGOF_ALIVE | (obs_inv ? MGOF_OBS_INV : 0) | (kind == 1 ? MGOF_AGENT : 0) | (kind == 0 ? MGOF_WALL : 0);
  o[MGO_LOC] = ((uint32_t)r << 16) | (uint32_t)c;
  o[MGO_META] = (uint32_t)t | ((uint32_t)(__ldg(tp + MGT_VIBE) & 0xff) << 16) | ((uint32_t)flags << 24);
  o[MGO_AGENT] = (uint32_t)aidx;
  o[MGO_ID] = (uint32_t)slot;
  o[MGO_NTOK] = MG_TOK_DIRTY;
  const int32_t* tg = pool(w, __ldg(tp + MGT_TAGS));
  for (int k = 0; k < w.TW; k++) o[MGO_TAGS + k] = (uint32_t)__ldg(tg + k);
  for (int k = 0; k < w.NDYN; k++) w.dyn_stamp[(size_t)slot * w.NDYN + k] = seq;  // tag-index registration order
  // initial inventory, stored in emission order; inserting back to front reproduces it
  const int32_t* iv = pool(w, __ldg(tp + MGT_INIT_INV));
  int ni = __ldg(tp + MGT_INIT_INV_N);
  for (int k = ni - 1; k >= 0; k--) inv_update(w, 0, o, __ldg(iv + 2 * k), __ldg(iv + 2 * k + 1), true, false);
}

// ==================================================================================================
// interpreter
// ==================================================================================================
struct QList {
  uint16_t* p;
  int n;
};
__device__ __noinline__ float eval_value(const Wv& w, int D, int node, const Ctx& ctx, int entity);
__device__ __noinline__ bool filter_pass(const Wv& w, int D, int fi, const Ctx& ctx);
__device__ __noinline__ QList query_eval(const Wv& w, int D, int qi, const Ctx& ctx);
__device__ __noinline__ void mutate(const Wv& w, int D, int mi, Ctx& ctx);
__device__ __noinline__ bool handler_apply(const Wv& w, int D, int h, Ctx& ctx);

// AND of a handler's filters (handler/handler.cpp:76-103).  Most filters are leaves -- a compare against the actor's or
// the target's record -- and most NOT / OR filters wrap leaves: those are evaluated right here from one row load each,
// so a typical filter list costs ONE call instead of one per filter (calls spill and reload registers through local
// memory, which the profile shows as the interpreter's main latency).  Anything else goes through filter_pass.
__device__ __forceinline__ int filter_leaf(const Wv& w, const int32_t* f, const Ctx& ctx) {  // 1 / 0, or -1: not a leaf
  const int2 r0 = __ldg((const int2*)f), r1 = __ldg((const int2*)f + 1);  // op, entity, a, b (rows are 8-byte aligned)
  const int op = r0.x, a = r1.x, b = r1.y;
  const int e = r0.y == MGE_ACTOR ? ctx.actor : r0.y == MGE_TARGET ? ctx.target : ctx.source;
  switch (op) {
    case MGF_VIBE:
      return e && o_vibe(objp(w, e)) == a;
    case MGF_RESOURCE:
      return e && (int)o_inv(w, objp(w, e))[a] >= b;
    case MGF_TAG_PREFIX: {
      if (!e) return 0;
      const uint32_t* x = objp(w, e);
      const int32_t* m = pool(w, a);
      for (int k = 0; k < w.TW; k++)
        if (x[MGO_TAGS + k] & (uint32_t)__ldg(m + k)) return 1;
      return 0;
    }
    case MGF_SHARED_TAG_PREFIX: {
      if (!ctx.actor || !ctx.target) return 0;
      const uint32_t *x = objp(w, ctx.actor), *y = objp(w, ctx.target);
      const int32_t* m = pool(w, a);
      for (int k = 0; k < w.TW; k++)
        if (x[MGO_TAGS + k] & y[MGO_TAGS + k] & (uint32_t)__ldg(m + k)) return 1;
      return 0;
    }
    case MGF_TARGET_LOC_EMPTY:
      return ctx.target == 0;
    case MGF_TARGET_IS_USABLE:
      return ctx.target != 0;
    case MGF_PERIODIC:
      return w.step >= (uint32_t)b && (w.step - (uint32_t)b) % (uint32_t)a == 0;
    default:
      return -1;
  }
}
__device__ __noinline__ bool filters_pass(const Wv& w, int D, int f0, int n, const Ctx& ctx) {
  const int32_t* f = sec(w, MGS_FILTERS) + f0 * MG_FILTER_WORDS;
  for (int i = 0; i < n; i++, f += MG_FILTER_WORDS) {
    int r = filter_leaf(w, f, ctx);
    if (r < 0) {
      const int op = __ldg(f);
      if ((op == MGF_NEG || op == MGF_OR) && D > 0) {  // one level of NOT(AND(leaves)) / OR(leaves) inline
        const int c0 = __ldg(f + 2), cn = __ldg(f + 3);
        const int32_t* cf = sec(w, MGS_FILTERS) + c0 * MG_FILTER_WORDS;
        bool all = true, any = false, flat = true;
        for (int k = 0; k < cn && flat; k++, cf += MG_FILTER_WORDS) {
          const int cr = filter_leaf(w, cf, ctx);
          if (cr < 0) {
            flat = false;
          } else {
            all = all && cr;
            any = any || cr;
            if (op == MGF_NEG ? !cr : cr) break;  // the reference stops at the first failing (NOT) / passing (OR) child
          }
        }
        if (flat) r = op == MGF_NEG ? !all : any;
      }
      if (r < 0) r = filter_pass(w, D, f0 + i, ctx);
    }
    if (!r) return false;
  }
  return true;
}
__device__ __forceinline__ int resolve_entity(const Ctx& c, int e) { return e == MGE_ACTOR ? c.actor : e == MGE_TARGET ? c.target : c.source; }
__device__ __forceinline__ int arena_top(const Wv& w) { return w.rs[4]; }
__device__ __forceinline__ void arena_set(const Wv& w, int t) { w.rs[4] = t; }

// ---- game values (core/game_value.cpp:14-148) -----------------------------------------------------
__device__ __forceinline__ float host_logf_plus1(const Wv& w, float term) {
  // std::log(term + 1.0f) through the host-libm table when the argument is an integer in range
  float x = __fadd_rn(term, 1.0f);
  if (x >= 1.0f && x <= 65536.0f && truncf(x) == x) return __ldg(w.logtab + (int)x - 1);
  return (float)log((double)x);  // non-integer argument: correctly rounded double log (<= 1 ulp vs glibc)
}
__device__ __noinline__ float eval_value(const Wv& w, int D, int node, const Ctx& ctx, int entity) {
  const int32_t* v = sec(w, MGS_VALUES) + node * MG_VALUE_WORDS;
  int op = __ldg(v), scope = __ldg(v + 1), a = __ldg(v + 2), b = __ldg(v + 3);
  switch (op) {
    case MGV_INVENTORY:
      if (entity) return (float)o_inv(w, objp(w, entity))[a];
      if (scope == MGSC_GAME) {
        int id = __ldg(sec(w, MGS_RES_GSTATS) + a);
        gstat_touch(w, id);
        return gstat_get(w, id);
      }
      return 0.0f;
    case MGV_STAT:
      if (scope == MGSC_GAME) {
        gstat_touch(w, a);
        return gstat_get(w, a);
      }
      if (entity) {
        const uint32_t* o = objp(w, entity);
        if (o_is_agent(o) && o_agent(o) >= 0) {
          astat_touch(w, o_agent(o), a);
          return astat_get(w, o_agent(o), a);
        }
      }
      return 0.0f;
    case MGV_CONST:
      return __int_as_float(a);
    default:
      break;
  }
  if (D > 0) {
    switch (op) {
      case MGV_QUERY_INVENTORY:
      case MGV_QUERY_COUNT: {
        if (op == MGV_QUERY_COUNT) {  // plain tag count: TagIndex::count_objects_with_tag
          const int32_t* q = sec(w, MGS_QUERIES) + b * MG_QUERY_WORDS;
          const int tag = __ldg(q + 3);
          if (__ldg(q) == MGQ_TAG && __ldg(q + 5) == 0 && __ldg(q + 1) < 0 && tag < w.NTAGS && w.tag_state[tag] >= 0)
            return (float)w.tag_state[tag];
        }
        Ctx c = ctx;
        c.actor = entity;  // HandlerContext::resolve_game_value: value_ctx.actor = entity
        int mark = arena_top(w);
        QList q = query_eval(w, D - 1, b, c);
        float total = 0.0f;
        if (op == MGV_QUERY_COUNT) {
          total = (float)q.n;
        } else {
          for (int i = 0; i < q.n; i++) total = __fadd_rn(total, (float)o_inv(w, objp(w, q.p[i]))[a]);
        }
        arena_set(w, mark);
        return total;
      }
      case MGV_SUM: {
        float total = 0.0f;
        const int32_t* kids = pool(w, a);
        int woff = __ldg(v + 4), lg = __ldg(v + 5);
        for (int i = 0; i < b; i++) {
          float term = eval_value(w, D - 1, __ldg(kids + i), ctx, entity);
          if (lg) term = host_logf_plus1(w, term);
          if (woff >= 0) term = __fmul_rn(term, __int_as_float(__ldg(pool(w, woff) + i)));
          total = __fadd_rn(total, term);
        }
        return total;
      }
      case MGV_RATIO: {
        float num = eval_value(w, D - 1, a, ctx, entity), den = eval_value(w, D - 1, b, ctx, entity);
        return den > 0.0f ? __fdiv_rn(num, den) : num;
      }
      case MGV_MAX:
      case MGV_MIN: {
        if (b == 0) return 0.0f;
        const int32_t* kids = pool(w, a);
        float best = op == MGV_MAX ? -3.402823466e+38f : 3.402823466e+38f;
        for (int i = 0; i < b; i++) {
          float x = eval_value(w, D - 1, __ldg(kids + i), ctx, entity);
          // std::max(a, b) = (a < b) ? b : a ; std::min(a, b) = (b < a) ? b : a
          if (op == MGV_MAX)
            best = (best < x) ? x : best;
          else
            best = (x < best) ? x : best;
        }
        return best;
      }
      default:
        break;
    }
  }
  set_error(w, MGERR_UNSUPPORTED, 2);
  return 0.0f;
}

// ---- filters (handler/filters/*.hpp) -------------------------------------------------------------
__device__ __noinline__ bool filter_pass(const Wv& w, int D, int fi, const Ctx& ctx) {
  const int32_t* f = sec(w, MGS_FILTERS) + fi * MG_FILTER_WORDS;
  int op = __ldg(f), a = __ldg(f + 2), b = __ldg(f + 3);
  int e = resolve_entity(ctx, __ldg(f + 1));
  switch (op) {
    case MGF_VIBE:
      return e && o_vibe(objp(w, e)) == a;
    case MGF_RESOURCE:
      return e && (int)o_inv(w, objp(w, e))[a] >= b;
    case MGF_SHARED_TAG_PREFIX: {
      if (!ctx.actor || !ctx.target) return false;
      const uint32_t *x = objp(w, ctx.actor), *y = objp(w, ctx.target);
      const int32_t* m = pool(w, a);
      for (int k = 0; k < w.TW; k++)
        if (x[MGO_TAGS + k] & y[MGO_TAGS + k] & (uint32_t)__ldg(m + k)) return true;
      return false;
    }
    case MGF_TAG_PREFIX: {
      if (!e) return false;
      const uint32_t* x = objp(w, e);
      const int32_t* m = pool(w, a);
      for (int k = 0; k < w.TW; k++)
        if (x[MGO_TAGS + k] & (uint32_t)__ldg(m + k)) return true;
      return false;
    }
    case MGF_TARGET_LOC_EMPTY:
      return ctx.target == 0;
    case MGF_TARGET_IS_USABLE:
      return ctx.target != 0;
    case MGF_PERIODIC:
      if (w.step < (uint32_t)b) return false;
      return (w.step - (uint32_t)b) % (uint32_t)a == 0;
    case MGF_MAX_DISTANCE:
      if (!e) return false;
      if (a < 0) {  // binary form (max_distance_filter.hpp:38-47)
        int ref = ctx.source ? ctx.source : ctx.actor;
        if (!ref) return false;
        if (b == 0) return true;
        const uint32_t *x = objp(w, e), *y = objp(w, ref);
        long long dr = o_r(x) - o_r(y), dc = o_c(x) - o_c(y);
        return dr * dr + dc * dc <= (long long)b * b;
      }
      break;
    default:
      break;
  }
  if (D > 0) {
    switch (op) {
      case MGF_GAME_VALUE:
        return eval_value(w, D - 1, a, ctx, e) >= eval_value(w, D - 1, b, ctx, e);
      case MGF_NEG:
        for (int i = 0; i < b; i++)
          if (!filter_pass(w, D - 1, a + i, ctx)) return true;
        return false;
      case MGF_OR:
        for (int i = 0; i < b; i++)
          if (filter_pass(w, D - 1, a + i, ctx)) return true;
        return false;
      case MGF_MAX_DISTANCE: {  // unary form: within radius of any query result (:49-62)
        int mark = arena_top(w);
        QList q = query_eval(w, D - 1, a, ctx);
        bool ok = false;
        if (b == 0) {
          ok = q.n > 0;
        } else {
          const uint32_t* x = objp(w, e);
          for (int i = 0; i < q.n && !ok; i++) {
            const uint32_t* y = objp(w, q.p[i]);
            long long dr = o_r(x) - o_r(y), dc = o_c(x) - o_c(y);
            ok = dr * dr + dc * dc <= (long long)b * b;
          }
        }
        arena_set(w, mark);
        return ok;
      }
      case MGF_QUERY_RESOURCE: {  // query_resource_filter.hpp:26-43
        int mark = arena_top(w);
        QList q = query_eval(w, D - 1, a, ctx);
        const int32_t* rq = pool(w, b);
        int nreq = __ldg(f + 4);
        bool ok = true;
        for (int i = 0; i < nreq && ok; i++) {
          uint32_t total = 0, need = (uint32_t)__ldg(rq + 2 * i + 1);
          int rid = __ldg(rq + 2 * i);
          for (int k = 0; k < q.n; k++) {
            total += o_inv(w, objp(w, q.p[k]))[rid];
            if (total >= need) break;
          }
          ok = total >= need;
        }
        arena_set(w, mark);
        return ok;
      }
      default:
        break;
    }
  }
  set_error(w, MGERR_UNSUPPORTED, 3);
  return false;
}

// ---- tag index + queries (core/tag_index.cpp, core/query_system.cpp:29-89,178-330) ------------------
// Objects carrying `tag`, in the reference's TagIndex order: registration order for tags fixed at
// creation (= slot order), insertion-stamp order for tags that can be added at run time.
__device__ __noinline__ QList collect_tag(const Wv& w, int tag) {
  int top = arena_top(w);
  uint16_t* out = w.arena + top;
  int n = 0, cap = w.ARENA - top;
  if (tag < w.NTAGS && w.tag_state[tag] >= 0) {  // the ordered member list is available
    n = min(w.tag_state[tag], cap);
    const uint16_t* l = w.tag_lists + tag * MG_TAG_LIST_CAP;
    for (int i = 0; i < n; i++) out[i] = l[i];
    arena_set(w, top + n);
    return QList{out, n};
  }
  const int last = w.E[MGEV_NEXT_OBJ];
  const int word = MGO_TAGS + (tag >> 5);
  const uint32_t bit = 1u << (tag & 31);
  for (int s = 1; s < last; s++) {
    const uint32_t* o = objp(w, s);
    if ((o[word] & bit) && o_alive(o)) {
      if (n >= cap) {
        set_error(w, MGERR_POOL_EXHAUSTED, 7);
        break;
      }
      out[n++] = (uint16_t)s;
    }
  }
  int ds = __ldg(sec(w, MGS_DYN_TAGS) + tag);
  if (ds >= 0) {  // insertion sort by stamp (stable: equal stamps keep slot order)
    for (int i = 1; i < n; i++) {
      uint16_t s = out[i];
      uint32_t key = w.dyn_stamp[(size_t)s * w.NDYN + ds];
      int j = i - 1;
      while (j >= 0 && w.dyn_stamp[(size_t)out[j] * w.NDYN + ds] > key) {
        out[j + 1] = out[j];
        j--;
      }
      out[j + 1] = s;
    }
  }
  arena_set(w, top + n);
  return QList{out, n};
}
__device__ __forceinline__ bool matches(const Wv& w, int D, int s, int f0, int n, const Ctx& ctx) {
  if (n == 0) return true;
  Ctx c = ctx;
  c.target = s;
  return filters_pass(w, D, f0, n, c);
}
__device__ __noinline__ QList query_eval(const Wv& w, int D, int qi, const Ctx& ctx) {
  const int32_t* q = sec(w, MGS_QUERIES) + qi * MG_QUERY_WORDS;
  const int kind = __ldg(q), mark = arena_top(w);
  uint16_t* res = w.arena + mark;
  int n = 0;
  if (D > 0) {
    if (kind == MGQ_TAG || kind == MGQ_FILTERED) {
      QList c = kind == MGQ_TAG ? collect_tag(w, __ldg(q + 3)) : query_eval(w, D - 1, __ldg(q + 3), ctx);
      int f0 = __ldg(q + 4), fn = __ldg(q + 5);
      for (int i = 0; i < c.n; i++) {  // in-place compaction; nested scratch lives above the list
        uint16_t s = c.p[i];
        if (matches(w, D - 1, s, f0, fn, ctx)) res[n++] = s;
      }
    } else if (kind == MGQ_CLOSURE) {
      QList roots = query_eval(w, D - 1, __ldg(q + 3), ctx);
      if (__ldg(q + 4) < 0) {
        n = roots.n;
      } else {
        QList cand = query_eval(w, D - 1, __ldg(q + 4), ctx);
        int top = arena_top(w);
        uint16_t* out = w.arena + top;
        int cap = w.ARENA - top, m = 0;
        if (cap < roots.n + cand.n) {
          set_error(w, MGERR_POOL_EXHAUSTED, 8);
        } else {
          arena_set(w, top + roots.n + cand.n);
          auto seen = [&](uint16_t s) {
            for (int k = 0; k < m; k++)
              if (out[k] == s) return true;
            return false;
          };
          for (int i = 0; i < roots.n; i++)
            if (!seen(roots.p[i])) out[m++] = roots.p[i];
          int e0 = __ldg(q + 5), en = __ldg(q + 6);
          for (int h = 0; h < m; h++) {  // BFS: the result list is its own frontier queue
            uint16_t cur = out[h];
            for (int i = 0; i < cand.n; i++) {
              uint16_t cd = cand.p[i];
              if (seen(cd)) continue;
              if (en > 0) {
                Ctx e = ctx;
                e.source = cur;
                e.target = cd;
                if (!filters_pass(w, D - 1, e0, en, e)) continue;
              }
              out[m++] = cd;
            }
          }
          int r0 = __ldg(q + 7), rn = __ldg(q + 8);
          for (int i = 0; i < m; i++)
            if (matches(w, D - 1, out[i], r0, rn, ctx)) res[n++] = out[i];  // res < out: forward copy is safe
        }
      }
    } else if (kind == MGQ_RAYCAST) {
      QList srcs = query_eval(w, D - 1, __ldg(q + 3), ctx);
      int top = arena_top(w);
      uint16_t* out = w.arena + top;
      int cap = w.ARENA - top, m = 0;
      const int nd = __ldg(q + 6) ? __ldg(q + 6) : 4;
      const int32_t* dirs = pool(w, __ldg(q + 5));
      const int b0 = __ldg(q + 7), bn = __ldg(q + 8), incl = __ldg(q + 9);
      for (int si = 0; si < srcs.n; si++) {
        int s = srcs.p[si];
        Ctx sc = ctx;
        sc.actor = s;
        sc.target = s;
        arena_set(w, top + m);
        int range = (int)eval_value(w, D - 1, __ldg(q + 4), sc, s);
        if (range <= 0) continue;
        const uint32_t* so = objp(w, s);
        for (int d = 0; d < nd; d++) {
          int dr, dc;
          if (__ldg(q + 6)) {
            dr = __ldg(dirs + 2 * d), dc = __ldg(dirs + 2 * d + 1);
          } else {  // CARDINALS {{-1,0},{1,0},{0,1},{0,-1}}
            dr = d == 0 ? -1 : d == 1 ? 1 : 0;
            dc = d == 2 ? 1 : d == 3 ? -1 : 0;
          }
          for (int dist = 1; dist <= range; dist++) {
            int r = o_r(so) + dr * dist, c = o_c(so) + dc * dist;
            if (!valid_loc(w, r, c)) break;
            int o = cell_at(w, r, c);
            if (!o) continue;
            bool blk = false;
            if (bn > 0) {
              Ctx bc = ctx;
              bc.target = o;
              arena_set(w, top + m);
              for (int i = 0; i < bn && !blk; i++) blk = filter_pass(w, D - 1, b0 + i, bc);
            }
            bool dup = false;
            for (int k = 0; k < m && !dup; k++) dup = out[k] == o;
            if ((!blk || incl) && !dup) {
              if (m < cap) out[m++] = (uint16_t)o;
              else set_error(w, MGERR_POOL_EXHAUSTED, 9);
            }
            if (blk) break;
          }
        }
      }
      for (int i = 0; i < m; i++) res[n++] = out[i];
    } else {
      set_error(w, MGERR_UNSUPPORTED, 10);
    }
    arena_set(w, mark + n);
    // apply_limits (query_system.cpp:74-89)
    if (__ldg(q + 2)) rng_shuffle(w, res, n);
    int mi = __ldg(q + 1);
    if (mi >= 0) {
      int mx = (int)eval_value(w, D - 1, mi, ctx, ctx.actor);
      if (mx >= 0 && n > mx) n = mx;
    }
    arena_set(w, mark + n);
    return QList{res, n};
  }
  set_error(w, MGERR_UNSUPPORTED, 11);
  return QList{res, 0};
}

// ---- tags (core/grid_object.cpp:91-141) ------------------------------------------------------------
__device__ __noinline__ void run_tag_handlers(const Wv& w, int D, int s, int tag, const Ctx& ctx) {
  const int32_t* t = tmpl(w, o_tmpl(objp(w, s)));
  const int32_t* pr = pool(w, __ldg(t + MGT_TAG_REMOVE));
  int n = __ldg(t + MGT_TAG_REMOVE_N);
  if (n == 0) return;
  Ctx h = ctx;
  h.actor = s;
  h.target = s;
  h.skip_trigger = false;
  for (int i = 0; i < n; i++)
    if (__ldg(pr + 2 * i) == tag) {
      if (D > 0)
        handler_apply(w, D - 1, __ldg(pr + 2 * i + 1), h);
      else
        set_error(w, MGERR_UNSUPPORTED, 12);
    }
}
__device__ __forceinline__ void add_tag(const Wv& w, int s, int tag) {
  uint32_t* o = objp(w, s);
  if (tag < 0 || tag >= 256 || o_has_tag(o, tag)) return;
  o[MGO_TAGS + (tag >> 5)] |= 1u << (tag & 31);
  o[MGO_NTOK] = MG_TOK_DIRTY;
  terr_touch(w, o);
  if (s < w.maxobj) tl_append(w, tag, s);  // territory proxy cells are not in the tag index
  int ds = __ldg(sec(w, MGS_DYN_TAGS) + tag);
  if (ds >= 0 && s < w.maxobj) w.dyn_stamp[(size_t)s * w.NDYN + ds] = (uint32_t)(w.E[MGEV_TAG_SEQ]++);
  // on_tag_add handlers cannot be configured from Python (no add_on_tag_add_handler call in the lowering)
}
__device__ __forceinline__ void remove_tag(const Wv& w, int D, int s, int tag, const Ctx& ctx) {
  uint32_t* o = objp(w, s);
  if (tag < 0 || tag >= 256 || !o_has_tag(o, tag)) return;
  o[MGO_TAGS + (tag >> 5)] &= ~(1u << (tag & 31));
  o[MGO_NTOK] = MG_TOK_DIRTY;
  terr_touch(w, o);
  if (s < w.maxobj) tl_erase(w, tag, s);
  if (!ctx.skip_trigger) run_tag_handlers(w, D, s, tag, ctx);
}

// ---- AOE bookkeeping shared by mutations (core/aoe_tracker.cpp:141-145,207-276) ----------------------
__device__ __forceinline__ uint32_t* aoe_rec(const Wv& w, int k) { return w.aoe_src + (size_t)k * w.AOEW; }
__device__ __forceinline__ const int32_t* aoe_cfg(const Wv& w, int cfg) { return sec(w, MGS_AOES) + cfg * MG_AOE_WORDS; }
__device__ __forceinline__ bool aoe_inside(const uint32_t* s, int ag) { return (s[4 + (ag >> 5)] >> (ag & 31)) & 1u; }
__device__ __forceinline__ void aoe_set_inside(uint32_t* s, int ag, bool v) {  // agents of one word may run on different lanes
  if (v)
    atomicOr(&s[4 + (ag >> 5)], 1u << (ag & 31));
  else
    atomicAnd(&s[4 + (ag >> 5)], ~(1u << (ag & 31)));
}
__device__ __noinline__ void apply_presence(const Wv& w, const int32_t* a, int target, int mult) {
  const int32_t* pd = pool(w, __ldg(a + 7));
  int n = __ldg(a + 8);
  uint32_t* o = objp(w, target);
  for (int i = 0; i < n; i++) inv_update(w, 2, o, __ldg(pd + 2 * i), __ldg(pd + 2 * i + 1) * mult);
}
__device__ __forceinline__ void register_aoe(const Wv& w, int slot, int cfg) {
  int k = w.E[MGEV_NUM_AOE];
  if (k >= w.AOECAP) {
    set_error(w, MGERR_POOL_EXHAUSTED, 13);
    return;
  }
  uint32_t* s = aoe_rec(w, k);
  s[0] = (uint32_t)slot, s[1] = (uint32_t)cfg, s[2] = objp(w, slot)[MGO_LOC], s[3] = 1;
  for (int i = 4; i < w.AOEW; i++) s[i] = 0;
  w.E[MGEV_NUM_AOE] = k + 1;
}
__device__ __noinline__ void remove_object(const Wv& w, int slot) {  // resource_mutation.hpp:88-97
  int na = w.E[MGEV_NUM_AOE];
  for (int k = 0; k < na; k++) {
    uint32_t* s = aoe_rec(w, k);
    if (!s[3] || (int)s[0] != slot) continue;
    const int32_t* a = aoe_cfg(w, (int)s[1]);
    for (int ag = 0; ag < w.A; ag++)
      if (aoe_inside(s, ag)) {
        apply_presence(w, a, (int)w.agents[ag * w.AS + MGAG_OBJ], -1);
        aoe_set_inside(s, ag, false);
      }
    s[3] = 0;
  }
  uint32_t* o = objp(w, slot);
  set_cell(w, o_r(o), o_c(o), 0);
  tl_register_object(w, slot, false);
  o[MGO_META] &= ~((uint32_t)MGOF_ALIVE << 24);  // leaves the tag index; territory sources persist like the reference
}
__device__ __noinline__ int spawn_object(const Wv& w, int t, int r, int c) {  // spawn_object_mutation.cpp:27-63
  int slot = w.E[MGEV_NEXT_OBJ];
  if (slot >= w.maxobj) {
    set_error(w, MGERR_POOL_EXHAUSTED, 14);
    return 0;
  }
  w.E[MGEV_NEXT_OBJ] = slot + 1;
  init_object(w, slot, t, r, c, -1, false, (uint32_t)(w.E[MGEV_TAG_SEQ]++));
  tl_register_object(w, slot, true);
  set_cell(w, r, c, slot);
  const int32_t* tp = tmpl(w, t);
  int na = __ldg(tp + MGT_AOES_N);
  for (int i = 0; i < na; i++) {
    int k = w.E[MGEV_NUM_AOE_PENDING];
    if (k >= w.PENDCAP) {
      set_error(w, MGERR_POOL_EXHAUSTED, 15);
      break;
    }
    w.aoe_pending[2 * k] = slot;
    w.aoe_pending[2 * k + 1] = __ldg(tp + MGT_AOES) + i;
    w.E[MGEV_NUM_AOE_PENDING] = k + 1;
  }
  return slot;
}

// ---- materialized queries (core/query_system.cpp:91-175) ---------------------------------------------
__device__ __noinline__ void recompute_mq(const Wv& w, int D, int tag, const Ctx& ctx, bool fire_handlers) {
  if (D > 1) {
    const int32_t* mq = sec(w, MGS_MQ);
    int nmq = w.hdr[MGH_NUM_MQ];
    Ctx tc = ctx;
    tc.skip_trigger = true;
    for (int i = 0; i < nmq; i++) {
      if (__ldg(mq + 2 * i) != tag) continue;
      int mark = arena_top(w);
      QList lost = collect_tag(w, tag);
      for (int k = 0; k < lost.n; k++) remove_tag(w, D - 1, lost.p[k], tag, tc);
      QList keep = query_eval(w, D - 1, __ldg(mq + 2 * i + 1), ctx);
      for (int k = 0; k < keep.n; k++) add_tag(w, keep.p[k], tag);
      if (fire_handlers) {
        tc.skip_trigger = false;
        for (int k = 0; k < lost.n; k++) {
          bool kept = false;
          for (int j = 0; j < keep.n && !kept; j++) kept = keep.p[j] == lost.p[k];
          if (!kept) run_tag_handlers(w, D - 1, lost.p[k], tag, tc);
        }
      }
      arena_set(w, mark);
      break;
    }
    return;
  }
  set_error(w, MGERR_UNSUPPORTED, 16);
}

// ---- mutations + handlers (handler/mutations/*.hpp, handler.cpp:76-103, multi_handler.cpp:8-21) ---
__device__ __noinline__ void mutate(const Wv& w, int D, int mi, Ctx& ctx) {
  const int32_t* m = sec(w, MGS_MUTATIONS) + mi * MG_MUTATION_WORDS;
  int op = __ldg(m), a = __ldg(m + 3), b = __ldg(m + 4), c = __ldg(m + 5), d = __ldg(m + 6);
  int e1 = resolve_entity(ctx, __ldg(m + 1)), e2 = resolve_entity(ctx, __ldg(m + 2));
  switch (op) {
    case MGM_RESOURCE_DELTA:  // resource_mutation.hpp:23-39
      if (ctx.deferred && __ldg(m + 1) == MGE_TARGET && ctx.target && !is_modifier(w, objp(w, ctx.target), a)) {
        Deferred& df = *ctx.deferred;
        if (!((df.seen >> a) & 1)) {
          df.seen |= (uint16_t)(1u << a);
          df.order[df.n++] = (int8_t)a;
          df.delta[a] = 0;
        }
        df.delta[a] += b;
        return;
      }
      if (e1) inv_update(w, 2, objp(w, e1), a, b);
      return;
    case MGM_RESOURCE_TRANSFER: {
      if (!e1 || !e2) return;
      uint32_t* src = objp(w, e1);
      int amount = b < 0 ? (int)o_inv(w, src)[a] : b;
      int moved = transfer(w, src, objp(w, e2), a, amount);
      if (moved > 0 && o_is_agent(src) && o_agent(src) >= 0) astat_add(w, o_agent(src), __ldg(sec(w, MGS_RES_STATS) + a * 4 + 3), (float)moved);
      if (c && ord_count(o_order(src)) == 0) remove_object(w, e1);
      return;
    }
    case MGM_CLEAR_INVENTORY: {
      if (!e1) return;
      uint32_t* o = objp(w, e1);
      if (b == 0) {
        uint64_t ord = o_order(o);
        int n = ord_count(ord);
        for (int i = 0; i < n; i++) {
          int it = ord_item(ord, i);
          inv_update(w, 2, o, it, -(int)o_inv(w, o)[it]);
        }
      } else {
        const int32_t* ids = pool(w, a);
        for (int i = 0; i < b; i++) {
          int it = __ldg(ids + i);
          inv_update(w, 2, o, it, -(int)o_inv(w, o)[it]);
        }
      }
      return;
    }
    case MGM_ATTACK: {
      if (!ctx.actor || !ctx.target) return;
      int weapon = o_inv(w, objp(w, ctx.actor))[a], armor = o_inv(w, objp(w, ctx.target))[b];
      int dmg = max(0, weapon * d / 100 - armor);
      if (dmg > 0) inv_update(w, 2, objp(w, ctx.target), c, -dmg);
      return;
    }
    case MGM_ADD_TAG:
      if (e1) add_tag(w, e1, a);
      return;
    case MGM_RELOCATE:
      if (ctx.actor && o_is_agent(objp(w, ctx.actor))) move_object(w, ctx.actor, ctx.tr, ctx.tc);
      return;
    case MGM_SWAP: {  // swap_mutation.hpp:15-22, core/grid.hpp:78-92
      if (!ctx.actor || !ctx.target) return;
      uint32_t *x = objp(w, ctx.actor), *y = objp(w, ctx.target);
      if (!o_is_agent(x) || !o_is_agent(y)) return;
      uint32_t lx = x[MGO_LOC], ly = y[MGO_LOC];
      set_cell(w, (int)(lx >> 16), (int)(lx & 0xffff), ctx.target);
      set_cell(w, (int)(ly >> 16), (int)(ly & 0xffff), ctx.actor);
      x[MGO_LOC] = ly;
      y[MGO_LOC] = lx;
      terr_touch(w, x), terr_touch(w, y);
      if (o_agent(x) >= 0) astat_add(w, o_agent(x), w.hdr[MGH_ST_ACTIONS_SWAP], 1.0f);
      return;
    }
    case MGM_CHANGE_VIBE:
      if (e1) o_set_vibe(objp(w, e1), a);
      return;
    case MGM_PUSH_OBJECT: {  // push_object_mutation.hpp:27-57
      if (!ctx.actor || !ctx.target) {
        ctx.failed = true;
        return;
      }
      const uint32_t *x = objp(w, ctx.actor), *t = objp(w, ctx.target);
      int dr = min(max(o_r(t) - o_r(x), -1), 1), dc = min(max(o_c(t) - o_c(x), -1), 1);
      if (!move_object(w, ctx.target, o_r(t) + dr, o_c(t) + dc)) ctx.failed = true;
      return;
    }
    case MGM_SPAWN_OBJECT: {
      if (a < 0 || !valid_loc(w, ctx.tr, ctx.tc) || cell_at(w, ctx.tr, ctx.tc) != 0) {
        ctx.failed = true;
        return;
      }
      int s = spawn_object(w, a, ctx.tr, ctx.tc);
      if (!s) {
        ctx.failed = true;
        return;
      }
      ctx.target = s;
      return;
    }
    default:
      break;
  }
  if (D > 0) {
    switch (op) {
      case MGM_STATS: {  // stats_mutation.hpp:21-41
        int ent = c ? ctx.actor : ctx.target;
        if (__ldg(m + 7)) {  // "stat := stat + integer": exact in any order below 2^24, so game stats take it atomically
          const float delta = __int_as_float(__ldg(m + 2));
          if (b == 0) {
            gstat_touch(w, a);
            atomicAdd(w.gstats + a, delta);
          } else if (ent) {
            const uint32_t* o = objp(w, ent);
            if (o_is_agent(o) && o_agent(o) >= 0) astat_add(w, o_agent(o), a, delta);
          }
          return;
        }
        float v = eval_value(w, D - 1, d, ctx, ent);
        if (b == 0) {
          gstat_set(w, a, v);
        } else if (ent) {
          const uint32_t* o = objp(w, ent);
          if (o_is_agent(o) && o_agent(o) >= 0) astat_set(w, o_agent(o), a, v);
        }
        return;
      }
      case MGM_GAME_VALUE: {  // game_value_mutation.hpp:18-25
        float delta = eval_value(w, D - 1, b, ctx, e1);
        const int32_t* v = sec(w, MGS_VALUES) + a * MG_VALUE_WORDS;
        int vop = __ldg(v), vscope = __ldg(v + 1), va = __ldg(v + 2);
        if (vop == MGV_INVENTORY) {
          if (e1)
            inv_update(w, 2, objp(w, e1), va, (int)delta);
          else if (vscope == MGSC_GAME)
            gstat_add(w, __ldg(sec(w, MGS_RES_GSTATS) + va), delta);
        } else if (vop == MGV_STAT) {
          if (vscope == MGSC_GAME) {
            gstat_add(w, va, delta);
          } else if (e1) {
            const uint32_t* o = objp(w, e1);
            if (o_is_agent(o) && o_agent(o) >= 0) astat_add(w, o_agent(o), va, delta);
          }
        }
        return;
      }
      case MGM_REMOVE_TAG:
        if (e1) remove_tag(w, D - 1, e1, a, ctx);
        return;
      case MGM_REMOVE_TAGS_PREFIX: {
        if (!e1) return;
        const int32_t* ids = pool(w, a);
        for (int i = 0; i < b; i++) remove_tag(w, D - 1, e1, __ldg(ids + i), ctx);
        return;
      }
      case MGM_RECOMPUTE_MQ:
        recompute_mq(w, D - 1, a, ctx, true);
        return;
      case MGM_QUERY_INVENTORY: {  // query_inventory_mutation.hpp:24-52
        int mark = arena_top(w);
        QList q = query_eval(w, D - 1, a, ctx);
        const int32_t* pr = pool(w, b);
        int stat_off = __ldg(m + 7);
        if (d) {
          if (e1) {
            for (int k = 0; k < q.n; k++)
              for (int i = 0; i < c; i++) {
                int rid = __ldg(pr + 2 * i), dl = __ldg(pr + 2 * i + 1), actual = 0;
                if (dl > 0)
                  actual = transfer(w, objp(w, e1), objp(w, q.p[k]), rid, dl);
                else if (dl < 0)
                  actual = transfer(w, objp(w, q.p[k]), objp(w, e1), rid, -dl);
                if (actual != 0 && stat_off >= 0 && __ldg(pool(w, stat_off) + rid) >= 0)
                  gstat_add(w, __ldg(pool(w, stat_off) + rid), (float)actual);
              }
          }
        } else {
          for (int k = 0; k < q.n; k++)
            for (int i = 0; i < c; i++) inv_update(w, 2, objp(w, q.p[k]), __ldg(pr + 2 * i), __ldg(pr + 2 * i + 1));
        }
        arena_set(w, mark);
        return;
      }
      case MGM_USE_TARGET: {  // use_target_mutation.hpp:17-30
        if (!ctx.target || !ctx.actor || !o_is_agent(objp(w, ctx.actor))) {
          ctx.failed = true;
          return;
        }
        int h = __ldg(tmpl(w, o_tmpl(objp(w, ctx.target))) + MGT_ON_USE);
        bool ok = false;
        if (h >= 0) {
          Ctx u = ctx;
          ok = handler_apply(w, D - 1, h, u);
        }
        if (!ok) {
          ctx.failed = true;
          return;
        }
        int after = __ldg(tmpl(w, o_tmpl(objp(w, ctx.actor))) + MGT_ON_AFTER_USE);
        if (after >= 0) handler_apply(w, D - 1, after, ctx);  // shares ctx (objects/agent.cpp:73-77)
        return;
      }
      case MGM_RAYCAST_SPAWN: {  // raycast_spawn_mutation.cpp:16-93
        if (!ctx.target || a < 0) {
          ctx.failed = true;
          return;
        }
        const uint32_t* org = objp(w, ctx.target);
        int orr = o_r(org), oc = o_c(org);
        int range = (int)eval_value(w, D - 1, d, ctx, ctx.target);
        if (range <= 0) return;
        const int32_t* dirs = pool(w, b);
        int b0 = __ldg(m + 7), bn = __ldg(m + 2);
        for (int di = 0; di < c; di++)
          for (int dist = 1; dist <= range; dist++) {
            int r = orr + __ldg(dirs + 2 * di) * dist, cc = oc + __ldg(dirs + 2 * di + 1) * dist;
            if (!valid_loc(w, r, cc)) break;
            int ex = cell_at(w, r, cc);
            if (ex) {
              bool blk = false;
              Ctx bc = ctx;
              bc.target = ex;
              for (int i = 0; i < bn && !blk; i++) blk = filter_pass(w, D - 1, b0 + i, bc);
              if (blk) break;
              continue;
            }
            spawn_object(w, a, r, cc);
          }
        return;
      }
      default:
        break;
    }
  }
  set_error(w, MGERR_UNSUPPORTED, 5 | (op << 8));
}

__device__ __noinline__ bool handler_apply(const Wv& w, int D, int h, Ctx& ctx) {
  const int32_t* hd = sec(w, MGS_HANDLERS) + h * MG_HANDLER_WORDS;
  int kind = __ldg(hd), a = __ldg(hd + 1), b = __ldg(hd + 2);
  if (kind == MGHK_SIMPLE) {
    if (!filters_pass(w, D, a, b, ctx)) return false;
    int m0 = __ldg(hd + 3), mn = __ldg(hd + 4);
    ctx.failed = false;
    for (int i = 0; i < mn; i++) {
      mutate(w, D, m0 + i, ctx);
      if (ctx.failed) return false;
    }
    return true;
  }
  if (D > 0) {
    bool any = false;
    const int32_t* kids = pool(w, a);
    for (int i = 0; i < b; i++)
      if (handler_apply(w, D - 1, __ldg(kids + i), ctx)) {
        any = true;
        if (kind == MGHK_FIRST_MATCH) return true;
      }
    return any;
  }
  set_error(w, MGERR_UNSUPPORTED, 6);
  return false;
}
