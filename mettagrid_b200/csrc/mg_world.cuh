// mg_world.cuh -- per-tick world systems: events, AOE, territory.  Events are serial (lane 0); the per-agent systems
// (fixed / mobile AOE, territory handlers) run one lane per agent when the compiler proved that their handlers only
// write the target agent (MGP_AOE, compiler.py: analyze_effects), else serially in the reference's order.
#pragma once
#include "mg_device.cuh"

// ---- events (handler/event.cpp:34-99, event_scheduler.cpp:36-53) -----------------------------------
__device__ __noinline__ int event_execute(const Wv& w, int ev, int depth) {
  const int32_t* e = sec(w, MGS_EVENTS) + ev * MG_EVENT_WORDS;
  const int maxt = __ldg(e + 1), f0 = __ldg(e + 2), fn = __ldg(e + 3), m0 = __ldg(e + 4), mn = __ldg(e + 5);
  Ctx g = make_ctx();
  int mark = arena_top(w);
  QList q = query_eval(w, MG_DEPTH, __ldg(e), g);
  if (maxt >= 0 && q.n > maxt) rng_shuffle(w, q.p, q.n);
  int applied = 0;
  for (int i = 0; i < q.n; i++) {
    if (maxt >= 0 && applied >= maxt) break;
    int t = q.p[i];
    Ctx c = make_ctx();
    c.actor = c.target = t;
    const uint32_t* o = objp(w, t);
    c.tr = o_r(o), c.tc = o_c(o);
    if (!filters_pass(w, MG_DEPTH, f0, fn, c)) continue;
    for (int k = 0; k < mn; k++) mutate(w, MG_DEPTH, m0 + k, c);
    applied++;
  }
  arena_set(w, mark);
  int fb = __ldg(e + 6);
  if (applied == 0 && fb >= 0) {
    if (depth >= 8) {
      set_error(w, MGERR_UNSUPPORTED, 17);
      return 0;
    }
    return event_execute(w, fb, depth + 1);
  }
  return applied;
}
__device__ __forceinline__ void process_events(const Wv& w) {
  const int n = w.hdr[MGH_NUM_EVENTS_SCHED];
  const int32_t* sch = sec(w, MGS_SCHEDULE);
  int cur = w.E[MGEV_EVENT_CURSOR];
  while (cur < n && __ldg(sch + 2 * cur) <= (int)w.step) {
    event_execute(w, __ldg(sch + 2 * cur + 1), 0);
    cur++;
  }
  w.E[MGEV_EVENT_CURSOR] = cur;
}

// ---- AOE (core/aoe_tracker.cpp:166-200,278-415) ------------------------------------------------------
__device__ __forceinline__ bool aoe_is_territory(const int32_t* a) { return __ldg(a + 6) == 0 && __ldg(a + 8) == 0 && __ldg(a) > 0; }
__device__ __forceinline__ bool aoe_covers(const uint32_t* s, const int32_t* a, int r, int c) {
  long long range = __ldg(a), dr = r - (int)(s[2] >> 16), dc = c - (int)(s[2] & 0xffffu);
  if (dr < -range || dr > range || dc < -range || dc > range) return false;
  long long d2 = dr * dr + dc * dc;
  if (d2 > range * range) return false;
  if (aoe_is_territory(a) && range >= 2 && d2 == range * range && (dr == 0 || dc == 0)) return false;
  return true;
}
// All lanes: which fixed sources can matter to agent `ag` right now -- alive, fixed, and either covering the agent's
// cell or still holding its "inside" bit (an exit is due).  One ballot per 32 sources replaces the serial scan over
// every source in aoe_apply_fixed, which then only visits the set bits.  Returns false when there are more than 128
// sources (the serial scan is used).
__device__ __forceinline__ bool aoe_relevant_mask(const Wv& w, int ag, int lane, uint32_t (&m)[4]) {
  const int na = w.E[MGEV_NUM_AOE];
  m[0] = m[1] = m[2] = m[3] = 0;
  if (na > 128) return false;
  const uint32_t* to = objp(w, (int)w.agents[ag * w.AS + MGAG_OBJ]);
  const int r = o_r(to), c = o_c(to);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (j * 32 >= na) break;
    const int k = j * 32 + lane;
    bool rel = false;
    if (k < na) {
      const uint32_t* s = aoe_rec(w, k);
      if (s[3]) {
        const int32_t* a = aoe_cfg(w, (int)s[1]);
        rel = __ldg(a + 1) && (aoe_inside(s, ag) || aoe_covers(s, a, r, c));
      }
    }
    m[j] = __ballot_sync(MG_FULL, rel);
  }
  return true;
}
// `m` (may be null): the sources aoe_relevant_mask selected for this agent at this moment
__device__ __noinline__ void aoe_apply_fixed(const Wv& w, int ag, const uint32_t* m) {
  const int target = (int)w.agents[ag * w.AS + MGAG_OBJ];
  Deferred df;
  df.n = 0;
  df.seen = 0;
  Ctx base = make_ctx();
  base.target = target;
  base.deferred = &df;
  const uint32_t* to = objp(w, target);
  const int na = w.E[MGEV_NUM_AOE];
  const uint32_t loc0 = to[MGO_LOC];  // the mask was computed for this cell; a handler that moves the target voids it
  // exits: sources the target was inside whose cells no longer cover it (registration order)
  for (int k = 0; k < na; k++) {
    if (m && !((m[k >> 5] >> (k & 31)) & 1u)) continue;
    uint32_t* s = aoe_rec(w, k);
    if (!s[3]) continue;
    const int32_t* a = aoe_cfg(w, (int)s[1]);
    if (!__ldg(a + 1)) continue;
    if (aoe_inside(s, ag) && !aoe_covers(s, a, o_r(to), o_c(to))) {
      aoe_set_inside(s, ag, false);
      apply_presence(w, a, target, -1);
    }
  }
  for (int k = 0; k < na; k++) {
    if (m && to[MGO_LOC] != loc0) m = nullptr;  // moved by an earlier source's handler: scan everything from here on
    if (m && !((m[k >> 5] >> (k & 31)) & 1u)) continue;
    uint32_t* s = aoe_rec(w, k);
    if (!s[3]) continue;
    const int32_t* a = aoe_cfg(w, (int)s[1]);
    const int mn = __ldg(a + 6), pn = __ldg(a + 8);
    if (!__ldg(a + 1) || !aoe_covers(s, a, o_r(to), o_c(to))) continue;
    if (mn == 0 && pn == 0) continue;
    Ctx c = base;
    c.actor = (int)s[0];
    const bool skip_self = !__ldg(a + 2) && (int)s[0] == target;
    const bool now = !skip_self && filters_pass(w, MG_DEPTH, __ldg(a + 3), __ldg(a + 4), c);
    const bool was = aoe_inside(s, ag);
    if (now && !was) {
      aoe_set_inside(s, ag, true);
      apply_presence(w, a, target, +1);
    } else if (!now && was) {
      aoe_set_inside(s, ag, false);
      apply_presence(w, a, target, -1);
    }
    if (now && mn > 0) {  // AOESource::try_apply re-checks the filters, then applies every mutation
      Ctx c2 = base;
      c2.actor = (int)s[0];
      if (filters_pass(w, MG_DEPTH, __ldg(a + 3), __ldg(a + 4), c2))
        for (int i = 0; i < mn; i++) mutate(w, MG_DEPTH, __ldg(a + 5) + i, c2);
    }
  }
  for (int i = 0; i < df.n; i++) {
    int rid = df.order[i];
    if (df.delta[rid] != 0) inv_update(w, 2, objp(w, target), rid, df.delta[rid]);
  }
}
// one mobile source against one agent (aoe_tracker.cpp:364-415)
__device__ __forceinline__ void aoe_mobile_one(const Wv& w, uint32_t* s, const int32_t* a, int ag) {
  const int so = (int)s[0], mn = __ldg(a + 6);
  const long long range = __ldg(a);
  const int t = (int)w.agents[ag * w.AS + MGAG_OBJ];
  if (!__ldg(a + 2) && so == t) return;
  const bool was = aoe_inside(s, ag);
  const uint32_t *x = objp(w, so), *y = objp(w, t);
  long long dr = o_r(x) - o_r(y), dc = o_c(x) - o_c(y);
  if (dr * dr + dc * dc > range * range) {
    if (was) {
      aoe_set_inside(s, ag, false);
      apply_presence(w, a, t, -1);
    }
    return;
  }
  Ctx c = make_ctx();
  c.actor = so;
  c.target = t;
  const bool now = filters_pass(w, MG_DEPTH, __ldg(a + 3), __ldg(a + 4), c);
  if (now) {
    if (!was) {
      aoe_set_inside(s, ag, true);
      apply_presence(w, a, t, +1);
    }
    if (mn > 0) {
      Ctx c2 = make_ctx();
      c2.actor = so;
      c2.target = t;
      if (filters_pass(w, MG_DEPTH, __ldg(a + 3), __ldg(a + 4), c2))
        for (int i = 0; i < mn; i++) mutate(w, MG_DEPTH, __ldg(a + 5) + i, c2);
    }
  } else if (was) {
    aoe_set_inside(s, ag, false);
    apply_presence(w, a, t, -1);
  }
}
// serial form: sources in registration order, agents in index order
__device__ __noinline__ void aoe_apply_mobile(const Wv& w) {
  const int na = w.E[MGEV_NUM_AOE];
  for (int k = 0; k < na; k++) {
    uint32_t* s = aoe_rec(w, k);
    if (!s[3]) continue;
    const int32_t* a = aoe_cfg(w, (int)s[1]);
    if (__ldg(a + 1)) continue;
    for (int ag = 0; ag < w.A; ag++) aoe_mobile_one(w, s, a, ag);
  }
}
// lane-per-agent form (MGP_AOE programs: the effects only touch the target agent): every source against agent `ag`
__device__ __noinline__ void aoe_apply_mobile_agent(const Wv& w, int ag) {
  const int na = w.E[MGEV_NUM_AOE];
  for (int k = 0; k < na; k++) {
    uint32_t* s = aoe_rec(w, k);
    if (!s[3]) continue;
    const int32_t* a = aoe_cfg(w, (int)s[1]);
    if (__ldg(a + 1)) continue;
    aoe_mobile_one(w, s, a, ag);
  }
}
__device__ __forceinline__ void aoe_flush_deferred(const Wv& w) {
  int n = w.E[MGEV_NUM_AOE_PENDING];
  for (int k = 0; k < n; k++) register_aoe(w, w.aoe_pending[2 * k], w.aoe_pending[2 * k + 1]);
  w.E[MGEV_NUM_AOE_PENDING] = 0;
}

// ---- territory (core/territory_tracker.cpp:20-50,215-346) -----------------------------------------------
#define MG_MAX_PREFIX_TAGS 16
// floor(sqrt(v)) for v < 2^62: a float estimate corrected with exact integer tests gives the same value as
// the reference's bit-by-bit floor_sqrt_u64 (core/territory_tracker.cpp:20-36)
__device__ __forceinline__ unsigned long long floor_sqrt_u64(unsigned long long v) {
  unsigned long long r = (unsigned long long)sqrt((double)v);
  while (r * r > v) r--;
  while ((r + 1) * (r + 1) <= v) r++;
  return r;
}
// Table of the territory sources that can own cells: (loc, range, strength | decay << 16, territory | prefix index << 8),
// and from it the OWNERSHIP MAP: per territory and cell the winning prefix index + 1.  Both only change when a
// territory source moves or changes tags (terr_touch sets MGEV_TERR_STALE: bit 0 = table, bit 1 = map), which in
// most games is never after the reset -- so the per-cell influence sums of compute_cell_ownership
// (territory_tracker.cpp:215-252) are paid once per episode instead of once per observed cell per tick.
__device__ __noinline__ void terr_build_table(const Wv& w) {
  const int nt = w.E[MGEV_NUM_TERR];
  int n = 0;
  for (int k = 0; k < nt; k++) {
    const uint32_t* s = w.terr_src + (size_t)k * 4;
    const int ti = (int)s[1];
    const int32_t* T_ = sec(w, MGS_TERRITORIES) + ti * MG_TERR_WORDS;
    const int32_t* pre = pool(w, __ldg(T_));
    const int np = min(__ldg(T_ + 1), MG_MAX_PREFIX_TAGS);
    const uint32_t* o = objp(w, (int)s[0]);
    int pi = -1;
    for (int i = 0; i < np; i++)
      if (o_has_tag(o, __ldg(pre + i))) {
        pi = i;
        break;
      }
    if (pi < 0) continue;  // find_matching_tag < 0: the source scores nothing
    const int strength = (int)s[2], decay = (int)s[3];
    uint32_t* e = w.terr_tab + (size_t)n * 4;
    e[0] = o[MGO_LOC];
    e[1] = (uint32_t)(decay > 0 ? strength / decay : strength);
    e[2] = (uint32_t)strength | ((uint32_t)decay << 16);
    e[3] = (uint32_t)ti | ((uint32_t)pi << 8);
    n++;
  }
  w.E[MGEV_RESERVED] = n;  // entries in the table
  w.E[MGEV_TERR_STALE] &= ~1;
}
__device__ __forceinline__ long long terr_score(const uint4 e, int r, int c) {
  const int range = (int)e.y;
  const int dr = r - (int)(e.x >> 16), dc = c - (int)(e.x & 0xffffu);
  if (dr < -range || dr > range || dc < -range || dc > range) return 0;
  const long long d2 = (long long)dr * dr + (long long)dc * dc;
  if (d2 > (long long)range * range) return 0;
  const long long strength = e.z & 0xffffu, decay = e.z >> 16;
  const long long sc = strength * 1024 - decay * (long long)floor_sqrt_u64((unsigned long long)d2 * 1024ull * 1024ull);
  return sc > 0 ? sc : 0;
}
// winning prefix index at (r, c) for territory ti, or -1 (read-only: callable from every lane)
__device__ __noinline__ int cell_owner_index(const Wv& w, int r, int c, int ti, const uint4* tab, int n) {
  const int32_t* T_ = sec(w, MGS_TERRITORIES) + ti * MG_TERR_WORDS;
  const int np = min(__ldg(T_ + 1), MG_MAX_PREFIX_TAGS);
  int win = -1;
  long long best = 0;
  bool tied = false;
  if (np <= 4) {  // the usual few teams: scores stay in registers
    long long s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int k = 0; k < n; k++) {
      const uint4 e = tab[k];
      if ((int)(e.w & 0xffu) != ti) continue;
      const long long sc = terr_score(e, r, c);
      const int pi = (int)((e.w >> 8) & 0xffu);
      s0 += pi == 0 ? sc : 0, s1 += pi == 1 ? sc : 0, s2 += pi == 2 ? sc : 0, s3 += pi == 3 ? sc : 0;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const long long si = i == 0 ? s0 : i == 1 ? s1 : i == 2 ? s2 : s3;
      if (i >= np || si <= 0) continue;
      if (si > best) {
        win = i;
        best = si;
        tied = false;
      } else if (si == best && win >= 0) {
        tied = true;
      }
    }
    return tied ? -1 : win;
  }
  long long score[MG_MAX_PREFIX_TAGS];
  for (int i = 0; i < np; i++) score[i] = 0;
  for (int k = 0; k < n; k++) {
    const uint4 e = tab[k];
    if ((int)(e.w & 0xffu) != ti) continue;
    score[(e.w >> 8) & 0xffu] += terr_score(e, r, c);
  }
  for (int i = 0; i < np; i++) {
    if (score[i] <= 0) continue;
    if (score[i] > best) {
      win = i;
      best = score[i];
      tied = false;
    } else if (score[i] == best && win >= 0) {
      tied = true;
    }
  }
  return tied ? -1 : win;
}
__device__ __forceinline__ int terr_prefix_tag(const Wv& w, int ti, int idx) {
  return idx < 0 ? -1 : __ldg(pool(w, __ldg(sec(w, MGS_TERRITORIES) + ti * MG_TERR_WORDS)) + idx);
}
// All lanes (convergent): bring the source table and the ownership map up to date.
__device__ __noinline__ void terr_refresh(const Wv& w, int lane) {
  if (w.NTERR == 0) return;
  const int stale = w.E[MGEV_TERR_STALE];
  __syncwarp();
  if (!stale) return;
  if (lane == 0 && (stale & 1)) terr_build_table(w);
  __syncwarp();
  const int n = w.E[MGEV_RESERVED], HW = w.H * w.W;
  const uint4* tab = (const uint4*)w.terr_tab;
  for (int ti = 0; ti < w.NTERR; ti++)
    for (int cell = lane; cell < HW; cell += 32) {
      const int r = cell / w.W, c = cell - r * w.W;
      w.owner[(size_t)ti * HW + cell] = (uint8_t)(cell_owner_index(w, r, c, ti, tab, n) + 1);
    }
  __syncwarp();
  if (lane == 0) w.E[MGEV_TERR_STALE] = 0;
  __syncwarp();
}
// winning tag at (r, c) or -1.  Serial callers may run while the map is stale (a handler just moved a source): they
// rebuild the table and evaluate the cell directly; the map itself is refreshed at the next convergent point.
__device__ __forceinline__ int terr_owner_tag(const Wv& w, int r, int c, int ti) {
  const int stale = w.E[MGEV_TERR_STALE];
  if (!stale) return terr_prefix_tag(w, ti, (int)w.owner[(size_t)ti * w.H * w.W + r * w.W + c] - 1);
  if (stale & 1) terr_build_table(w);
  return terr_prefix_tag(w, ti, cell_owner_index(w, r, c, ti, (const uint4*)w.terr_tab, w.E[MGEV_RESERVED]));
}
// aoe_mask token of one in-map cell (:254-273); the map must be fresh (terr_refresh)
__device__ __forceinline__ int territory_mask(const Wv& w, int cell, const uint32_t* observer) {
  const int HW = w.H * w.W;
  for (int ti = 0; ti < w.NTERR; ti++) {
    const int v = w.owner[(size_t)ti * HW + cell];
    if (!v) continue;
    return o_has_tag(observer, terr_prefix_tag(w, ti, v - 1)) ? 1 : 2;
  }
  return 0;
}
// territory handlers run with a proxy-cell object as actor (territory_tracker.cpp:103-107,275-346); every agent has
// its own proxy per territory so that agents can be processed on different lanes
__device__ __noinline__ void terr_run(const Wv& w, int ti, int list_off, int n, int tag, int target, int ag) {
  const int px = w.maxobj + ti * w.A + ag;
  uint32_t* po = objp(w, px);
  for (int k = 0; k < w.TW; k++) po[MGO_TAGS + k] = 0;
  po[MGO_TAGS + (tag >> 5)] = 1u << (tag & 31);
  po[MGO_LOC] = objp(w, target)[MGO_LOC];
  const int32_t* hs = pool(w, list_off);
  for (int i = 0; i < n; i++) {
    const int32_t* hd = sec(w, MGS_HANDLERS) + __ldg(hs + i) * MG_HANDLER_WORDS;
    Ctx c = make_ctx();
    c.actor = px;
    c.target = target;
    if (filters_pass(w, MG_DEPTH, __ldg(hd + 1), __ldg(hd + 2), c))
      for (int k = 0; k < __ldg(hd + 4); k++) mutate(w, MG_DEPTH, __ldg(hd + 3) + k, c);  // no mutation_failed check
  }
}
__device__ __noinline__ void terr_apply(const Wv& w, int ag) {
  const int target = (int)w.agents[ag * w.AS + MGAG_OBJ];
  for (int ti = 0; ti < w.NTERR; ti++) {
    const int32_t* T_ = sec(w, MGS_TERRITORIES) + ti * MG_TERR_WORDS;
    const uint32_t* to = objp(w, target);
    const int cur = terr_owner_tag(w, o_r(to), o_c(to), ti);
    const int prev = w.inside_tag[ag * w.NTERR + ti];
    if (prev != cur && prev >= 0) terr_run(w, ti, __ldg(T_ + 4), __ldg(T_ + 5), prev, target, ag);
    if (prev != cur && cur >= 0) terr_run(w, ti, __ldg(T_ + 2), __ldg(T_ + 3), cur, target, ag);
    w.inside_tag[ag * w.NTERR + ti] = (int16_t)cur;
    if (cur >= 0) terr_run(w, ti, __ldg(T_ + 6), __ldg(T_ + 7), cur, target, ag);
  }
}
