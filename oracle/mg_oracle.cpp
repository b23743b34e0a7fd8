// mg_oracle.cpp -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's per-timestep step
// (MettaGrid::_step and everything it calls), one environment per instance, strictly serial.
//
// It interprets the same compiled game program (include/mg_program.h) as the sm_100a kernels and is
// the checker the parity tests compare the CUDA path against.  It is NOT part of the product: only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
//
// Pinning: tests/test_oracle_vs_reference.py drives this file and the real reference build
// (oracle/_ref, compiled from /root/reference by oracle/Makefile.ref) with identical configs, seeds
// and action sequences and requires byte-equal observations/rewards/stats per step; the committed
// fixtures under tests/golden/ were produced by the real reference (tests/golden/make_golden.py).
//
// Every function cites the reference file:line it restates (paths relative to /root/reference/cpp).
// Build: g++ -O2 -ffp-contract=off -shared -fPIC  (no FMA contraction: the reference is built for
// baseline x86-64, bindings/../BUILD.bazel:14-20, so its float math is plain mul/add).
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../include/mg_program.h"

namespace {

// ------------------------------------------------------------------------------------------------
// MT19937 + libstdc++ std::shuffle (/usr/include/c++/13/bits/stl_algo.h:3719-3805,
// bits/uniform_int_dist.h:252-373).  Call sites: bindings/mettagrid_c.cpp:960,
// src/mettagrid/core/query_system.cpp:79, src/mettagrid/handler/event.cpp:43.
// ------------------------------------------------------------------------------------------------
struct Mt {
  uint32_t s[624];
  int idx;
  void seed(uint32_t v) {
    s[0] = v;
    for (int i = 1; i < 624; i++) s[i] = 1812433253u * (s[i - 1] ^ (s[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int i = 0; i < 624; i++) {
        uint32_t y = (s[i] & 0x80000000u) | (s[(i + 1) % 624] & 0x7fffffffu);
        s[i] = s[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = s[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  // Lemire multiply-shift with rejection, uniform in [0, range), range < 2^32
  uint32_t below(uint32_t range) {
    uint64_t prod = (uint64_t)next() * range;
    uint32_t low = (uint32_t)prod;
    if (low < range) {
      uint32_t thr = (uint32_t)(-range) % range;
      while (low < thr) {
        prod = (uint64_t)next() * range;
        low = (uint32_t)prod;
      }
    }
    return (uint32_t)(prod >> 32);
  }
  template <class T>
  void shuffle(std::vector<T>& v) {
    size_t n = v.size();
    if (n < 2) return;
    size_t i = 1;
    if (n % 2 == 0) {
      size_t j = below(2);
      std::swap(v[i], v[j]);
      i++;
    }
    while (i < n) {
      uint32_t sr = (uint32_t)i + 1;
      uint32_t x = below(sr * (sr + 1));
      std::swap(v[i], v[x / (sr + 1)]);
      i++;
      std::swap(v[i], v[x % (sr + 1)]);
      i++;
    }
  }
};

struct Obj {
  bool alive = false;
  bool in_grid = false;
  int tmpl = -1;
  int kind = 2;  // 0 wall, 1 agent, 2 object, 3 territory proxy
  int type_id = 0;
  int r = 0, c = 0;
  uint32_t id = 0;
  uint32_t visited = 0;
  int vibe = 0;
  int agent = -1;
  bool obs_inv = true;
  uint32_t tags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint16_t inv[16] = {0};
  std::vector<uint8_t> order;  // present resources, most recently inserted first
};

struct Agent {
  int obj = 0;
  int group = 0;
  int spawn_r = 0, spawn_c = 0;
  int prev_r = 0, prev_c = 0;  // Agent::prev_location
  int step_r = 0, step_c = 0;  // _prev_agent_locations
  uint32_t swm = 0;
  uint32_t max_dist = 0;
  std::vector<uint8_t> seen;
  uint32_t unique = 0;
  std::vector<float> reward_prev;
  std::vector<float> stats;
  std::vector<uint8_t> touched;
};

struct AoeSource {
  int obj, cfg, r, c;
  bool alive;
  std::vector<uint8_t> inside;  // per agent
};
struct TerrSource {
  int obj, terr, strength, decay;
};

struct Deferred {
  std::vector<int> order;
  int32_t delta[16];
  bool seen[16];
};

// handler/handler_context.hpp:34-55
struct Ctx {
  int actor = 0, target = 0, source = 0;
  int distance = 0;
  int tr = 0, tc = 0;
  int move_dir = 0;
  bool skip_trigger = false;
  bool failed = false;
  Deferred* deferred = nullptr;
};

struct Env {
  std::vector<int32_t> P;
  const int32_t* p;
  int H, W, A, T, R, TW, B, ND;
  std::vector<int> cells;
  std::vector<Obj> objs;
  std::vector<Agent> agents;
  std::vector<float> gstats;
  std::vector<uint8_t> gtouched;
  std::vector<std::vector<int>> tag_index;  // core/tag_index.cpp: insertion-ordered per tag
  Mt rng;
  uint32_t step = 0;
  uint32_t next_id = 1;
  int event_cursor = 0;
  int error = 0, err_info = 0;
  std::vector<AoeSource> aoe;
  std::vector<std::pair<int, int>> aoe_pending;
  std::vector<TerrSource> terr;
  std::vector<std::vector<int>> inside_tag;  // [agent][territory] -> tag or -1
  int proxy0 = 0;
  // buffers
  std::vector<uint8_t> obs;
  std::vector<float> rewards, episode_rewards;
  std::vector<uint8_t> terminals, truncations, success;
  std::vector<int32_t> last_actions;

  int hdr(int k) const { return p[k]; }
  const int32_t* sec(int k) const { return p + p[k]; }
  const int32_t* pool(int off) const { return p + p[MGS_POOL] + off; }
  const int32_t* tmpl(int t) const { return sec(MGS_TEMPLATES) + t * MG_TEMPLATE_WORDS; }

  // ---- stats (systems/stats_tracker.hpp:57-98) ------------------------------------------------
  void astat_add(int a, int id, float v) {
    agents[a].stats[id] += v;
    agents[a].touched[id] = 1;
  }
  void astat_set(int a, int id, float v) {
    agents[a].stats[id] = v;
    agents[a].touched[id] = 1;
  }
  void gstat_add(int id, float v) {
    gstats[id] += v;
    gtouched[id] = 1;
  }

  // ---- inventory (objects/inventory.cpp:38-173, inventory.hpp:26-40) ---------------------------
  int limit_of(const Obj& o, int item) const {
    if (o.tmpl < 0) return -1;
    return pool(tmpl(o.tmpl)[MGT_LIMIT_OF])[item];
  }
  bool is_modifier(const Obj& o, int item) const {
    if (o.tmpl < 0) return false;
    return (tmpl(o.tmpl)[MGT_MODIFIER_MASK] >> item) & 1;
  }
  uint16_t effective_limit(const Obj& o, int lim) const {
    const int32_t* L = sec(MGS_LIMITS) + lim * MG_LIMIT_WORDS;
    int sum = 0;
    const int32_t* m = pool(L[2]);
    for (int i = 0; i < L[3]; i++) sum += (int)o.inv[m[2 * i]] * m[2 * i + 1];
    int eff = std::min(L[1], std::max(L[0], sum));
    eff = std::min(std::max(eff, 0), 65535);
    return (uint16_t)eff;
  }
  uint16_t limit_amount(const Obj& o, int lim) const {
    const int32_t* L = sec(MGS_LIMITS) + lim * MG_LIMIT_WORDS;
    const int32_t* mem = pool(L[4]);
    uint32_t s = 0;
    for (int i = 0; i < L[5]; i++) s += o.inv[mem[i]];
    return (uint16_t)s;  // SharedInventoryLimit::amount is a uint16 running sum
  }
  // objects/agent.cpp:106-121
  void on_inventory_change(Obj& o, int item, int delta) {
    if (o.kind != 1 || o.agent < 0) return;
    const int32_t* rs = sec(MGS_RES_STATS) + item * 4;
    if (delta > 0)
      astat_add(o.agent, rs[0], (float)delta);
    else
      astat_add(o.agent, rs[1], (float)(-delta));
    astat_set(o.agent, rs[2], (float)o.inv[item]);
    if (o.inv[item] == 0 && delta < 0 && item == hdr(MGH_HP_RESOURCE)) astat_add(o.agent, hdr(MGH_ST_DEATH), 1.0f);
  }
  int inv_update(Obj& o, int item, int attempted, bool ignore_limits = false, bool notify = true) {
    int initial = o.inv[item];
    int na = initial + attempted;
    int mx = 65535;
    int lim = limit_of(o, item);
    if (!ignore_limits && lim >= 0) {
      int eff = effective_limit(o, lim);
      int others = (int)limit_amount(o, lim) - initial;
      if (others < 0) others = 0;
      int mi = eff - others;
      if (mi < 0) mi = 0;
      mx = (uint16_t)mi;
    }
    int clamped = std::min(std::max(na, 0), mx);
    if (clamped == 0) {
      if (initial != 0) o.order.erase(std::find(o.order.begin(), o.order.end(), (uint8_t)item));
    } else if (initial == 0) {
      o.order.insert(o.order.begin(), (uint8_t)item);
    }
    o.inv[item] = (uint16_t)clamped;
    int d = clamped - initial;
    if (notify && d != 0) on_inventory_change(o, item, d);
    if (d < 0 && is_modifier(o, item)) enforce_all_limits(o);
    return d;
  }
  uint16_t free_space(const Obj& o, int item) const {
    int lim = limit_of(o, item);
    if (lim < 0) return (uint16_t)(65535 - o.inv[item]);
    uint16_t used = limit_amount(o, lim), eff = effective_limit(o, lim);
    return eff > used ? (uint16_t)(eff - used) : 0;
  }
  void enforce_all_limits(Obj& o) {
    if (o.tmpl < 0) return;
    const int32_t* t = tmpl(o.tmpl);
    const int32_t* lo = pool(t[MGT_LIMIT_ORDER]);
    for (int i = 0; i < t[MGT_LIMIT_ORDER_N]; i++) {
      int lim = lo[i];
      int excess = (int)limit_amount(o, lim) - (int)effective_limit(o, lim);
      if (excess <= 0) continue;
      const int32_t* L = sec(MGS_LIMITS) + lim * MG_LIMIT_WORDS;
      const int32_t* mem = pool(L[4]);
      for (int k = 0; k < L[5]; k++) {
        int item = mem[k];
        int drop = std::min((int)o.inv[item], excess);
        if (drop > 0) {
          inv_update(o, item, -drop);
          excess = (int)limit_amount(o, lim) - (int)effective_limit(o, lim);
        }
        if (excess <= 0) break;
      }
    }
  }
  // objects/has_inventory.cpp:76-108
  int transfer(Obj& src, Obj& dst, int item, int delta) {
    if (delta <= 0) return 0;
    int give = std::min((int)src.inv[item], delta);
    int amount = std::min(give, (int)free_space(dst, item));
    inv_update(src, item, -amount);
    inv_update(dst, item, amount);
    return amount;
  }

  // ---- grid (core/grid.hpp:31-130) + territory re-registration is positional (see terr_*) ------
  bool valid(int r, int c) const { return r >= 0 && c >= 0 && r < H && c < W; }
  int at(int r, int c) const { return valid(r, c) ? cells[r * W + c] : 0; }
  bool move_object(int s, int r, int c) {
    if (!valid(r, c) || cells[r * W + c] != 0) return false;
    Obj& o = objs[s];
    cells[r * W + c] = s;
    cells[o.r * W + o.c] = 0;
    o.r = r;
    o.c = c;
    return true;
  }

  // ---- tags (core/grid_object.cpp:91-141, core/tag_index.cpp) -----------------------------------
  static bool has_tag(const Obj& o, int t) { return t >= 0 && t < 256 && ((o.tags[t >> 5] >> (t & 31)) & 1); }
  void run_tag_handlers(int s, int tag, const Ctx& ctx) {
    Obj& o = objs[s];
    if (o.tmpl < 0) return;
    const int32_t* t = tmpl(o.tmpl);
    const int32_t* pr = pool(t[MGT_TAG_REMOVE]);
    Ctx h = ctx;
    h.actor = s;
    h.target = s;
    h.skip_trigger = false;
    for (int i = 0; i < t[MGT_TAG_REMOVE_N]; i++)
      if (pr[2 * i] == tag) handler_apply(pr[2 * i + 1], h);
  }
  void add_tag(int s, int tag, const Ctx& ctx) {
    Obj& o = objs[s];
    if (tag < 0 || tag >= 256 || has_tag(o, tag)) return;
    o.tags[tag >> 5] |= 1u << (tag & 31);
    if (o.kind != 3) tag_index[tag].push_back(s);
    // on_tag_add handlers are never configured from Python (mettagrid_c_config.py has no add_on_tag_add call)
  }
  void remove_tag(int s, int tag, const Ctx& ctx) {
    Obj& o = objs[s];
    if (tag < 0 || tag >= 256 || !has_tag(o, tag)) return;
    o.tags[tag >> 5] &= ~(1u << (tag & 31));
    auto& v = tag_index[tag];
    v.erase(std::remove(v.begin(), v.end(), s), v.end());
    if (!ctx.skip_trigger) run_tag_handlers(s, tag, ctx);
  }

  // ---- game values (core/game_value.cpp:14-148, handler_context.cpp:8-27) ------------------------
  float value(int node, Ctx ctx, int entity) {
    ctx.actor = entity;
    const int32_t* v = sec(MGS_VALUES) + node * MG_VALUE_WORDS;
    switch (v[0]) {
      case MGV_INVENTORY:
        if (entity) return (float)objs[entity].inv[v[2]];
        if (v[1] == MGSC_GAME) {
          int id = sec(MGS_RES_GSTATS)[v[2]];
          gtouched[id] = 1;
          return gstats[id];
        }
        return 0.0f;
      case MGV_STAT:
        if (v[1] == MGSC_GAME) {
          gtouched[v[2]] = 1;
          return gstats[v[2]];
        }
        if (entity && objs[entity].kind == 1 && objs[entity].agent >= 0) {
          Agent& a = agents[objs[entity].agent];
          a.touched[v[2]] = 1;
          return a.stats[v[2]];
        }
        return 0.0f;
      case MGV_CONST: {
        float f;
        memcpy(&f, &v[2], 4);
        return f;
      }
      case MGV_QUERY_INVENTORY: {
        std::vector<int> res = query(v[3], ctx);
        float total = 0.0f;
        for (int s : res) total += (float)objs[s].inv[v[2]];
        return total;
      }
      case MGV_QUERY_COUNT:
        return (float)query(v[3], ctx).size();
      case MGV_SUM: {
        float total = 0.0f;
        const int32_t* kids = pool(v[2]);
        for (int i = 0; i < v[3]; i++) {
          float term = value(kids[i], ctx, entity);
          if (v[5]) term = logf(term + 1.0f);
          if (v[4] >= 0) {
            float w;
            memcpy(&w, pool(v[4]) + i, 4);
            term *= w;
          }
          total += term;
        }
        return total;
      }
      case MGV_RATIO: {
        float num = value(v[2], ctx, entity), den = value(v[3], ctx, entity);
        return den > 0.0f ? num / den : num;
      }
      case MGV_MAX: {
        if (v[3] == 0) return 0.0f;
        float best = -3.402823466e+38f;
        const int32_t* kids = pool(v[2]);
        for (int i = 0; i < v[3]; i++) best = std::max(best, value(kids[i], ctx, entity));
        return best;
      }
      case MGV_MIN: {
        if (v[3] == 0) return 0.0f;
        float best = 3.402823466e+38f;
        const int32_t* kids = pool(v[2]);
        for (int i = 0; i < v[3]; i++) best = std::min(best, value(kids[i], ctx, entity));
        return best;
      }
    }
    return 0.0f;
  }
  int resolve(const Ctx& ctx, int e) const { return e == MGE_ACTOR ? ctx.actor : e == MGE_TARGET ? ctx.target : ctx.source; }
  void value_update(int node, Ctx ctx, int entity, float delta) {  // resolved_game_value.hpp:50-62
    const int32_t* v = sec(MGS_VALUES) + node * MG_VALUE_WORDS;
    if (v[0] == MGV_INVENTORY) {
      if (entity) {
        inv_update(objs[entity], v[2], (int32_t)delta);
      } else if (v[1] == MGSC_GAME) {
        gstat_add(sec(MGS_RES_GSTATS)[v[2]], delta);
      }
    } else if (v[0] == MGV_STAT) {
      if (v[1] == MGSC_GAME)
        gstat_add(v[2], delta);
      else if (entity && objs[entity].kind == 1 && objs[entity].agent >= 0)
        astat_add(objs[entity].agent, v[2], delta);
    }
  }

  // ---- filters (handler/filters/*.hpp) -----------------------------------------------------------
  bool filter(int fi, const Ctx& ctx) {
    const int32_t* f = sec(MGS_FILTERS) + fi * MG_FILTER_WORDS;
    int e = resolve(ctx, f[1]);
    switch (f[0]) {
      case MGF_VIBE:
        return e && objs[e].vibe == f[2];
      case MGF_RESOURCE:
        return e && objs[e].inv[f[2]] >= f[3];
      case MGF_SHARED_TAG_PREFIX: {
        const int32_t* m = pool(f[2]);
        if (!ctx.actor || !ctx.target) return false;
        for (int w = 0; w < TW; w++)
          if (objs[ctx.actor].tags[w] & objs[ctx.target].tags[w] & (uint32_t)m[w]) return true;
        return false;
      }
      case MGF_TAG_PREFIX: {
        if (!e) return false;
        const int32_t* m = pool(f[2]);
        for (int w = 0; w < TW; w++)
          if (objs[e].tags[w] & (uint32_t)m[w]) return true;
        return false;
      }
      case MGF_GAME_VALUE:
        return value(f[2], ctx, e) >= value(f[3], ctx, e);
      case MGF_NEG:
        for (int i = 0; i < f[3]; i++)
          if (!filter(f[2] + i, ctx)) return true;
        return false;
      case MGF_OR:
        for (int i = 0; i < f[3]; i++)
          if (filter(f[2] + i, ctx)) return true;
        return false;
      case MGF_MAX_DISTANCE: {  // max_distance_filter.hpp:31-63
        if (!e) return false;
        int64_t rad = f[3];
        if (f[2] < 0) {
          int ref = ctx.source ? ctx.source : ctx.actor;
          if (!ref) return false;
          if (rad == 0) return true;
          int64_t dr = objs[e].r - objs[ref].r, dc = objs[e].c - objs[ref].c;
          return dr * dr + dc * dc <= rad * rad;
        }
        std::vector<int> src = query(f[2], ctx);
        if (rad == 0) return !src.empty();
        for (int s : src) {
          int64_t dr = objs[e].r - objs[s].r, dc = objs[e].c - objs[s].c;
          if (dr * dr + dc * dc <= rad * rad) return true;
        }
        return false;
      }
      case MGF_QUERY_RESOURCE: {  // query_resource_filter.hpp:26-43
        std::vector<int> res = query(f[2], ctx);
        const int32_t* rq = pool(f[3]);
        for (int i = 0; i < f[4]; i++) {
          uint32_t total = 0;
          for (int s : res) {
            total += objs[s].inv[rq[2 * i]];
            if (total >= (uint32_t)rq[2 * i + 1]) break;
          }
          if (total < (uint32_t)rq[2 * i + 1]) return false;
        }
        return true;
      }
      case MGF_TARGET_LOC_EMPTY:
        return ctx.target == 0;
      case MGF_TARGET_IS_USABLE:
        return ctx.target != 0;  // every GridObject derives from Usable (core/grid_object.hpp:93)
      case MGF_PERIODIC:
        if (step < (uint32_t)f[3]) return false;
        return (step - (uint32_t)f[3]) % (uint32_t)f[2] == 0;
    }
    return false;
  }
  bool filters_pass(int f0, int n, const Ctx& ctx) {
    for (int i = 0; i < n; i++)
      if (!filter(f0 + i, ctx)) return false;
    return true;
  }

  // ---- queries (core/query_system.cpp:29-89,178-330) ----------------------------------------------
  std::vector<int> apply_limits(std::vector<int> res, const int32_t* q, const Ctx& ctx) {
    if (q[2]) rng.shuffle(res);
    if (q[1] >= 0) {
      int mx = (int)value(q[1], ctx, ctx.actor);
      if (mx >= 0 && (int)res.size() > mx) res.resize(mx);
    }
    return res;
  }
  bool matches(int s, int f0, int n, const Ctx& ctx) {
    if (n == 0) return true;
    Ctx c = ctx;
    c.target = s;
    return filters_pass(f0, n, c);
  }
  std::vector<int> query(int qi, const Ctx& ctx) {
    const int32_t* q = sec(MGS_QUERIES) + qi * MG_QUERY_WORDS;
    std::vector<int> res;
    switch (q[0]) {
      case MGQ_TAG: {
        std::vector<int> cand = tag_index[q[3]];
        for (int s : cand)
          if (matches(s, q[4], q[5], ctx)) res.push_back(s);
        break;
      }
      case MGQ_FILTERED: {
        std::vector<int> cand = query(q[3], ctx);
        for (int s : cand)
          if (matches(s, q[4], q[5], ctx)) res.push_back(s);
        break;
      }
      case MGQ_CLOSURE: {
        std::vector<int> roots = query(q[3], ctx);
        if (q[4] < 0) return apply_limits(roots, q, ctx);
        std::vector<int> candp = query(q[4], ctx);
        std::vector<uint8_t> vis(objs.size(), 0);
        std::vector<int> frontier;
        for (int s : roots)
          if (!vis[s]) {
            vis[s] = 1;
            frontier.push_back(s);
            res.push_back(s);
          }
        for (size_t h = 0; h < frontier.size(); h++) {
          int cur = frontier[h];
          for (int cd : candp) {
            if (vis[cd]) continue;
            if (q[6] > 0) {
              Ctx e = ctx;
              e.source = cur;
              e.target = cd;
              if (!filters_pass(q[5], q[6], e)) continue;
            }
            vis[cd] = 1;
            frontier.push_back(cd);
            res.push_back(cd);
          }
        }
        if (q[8] > 0) {
          std::vector<int> out;
          for (int s : res)
            if (matches(s, q[7], q[8], ctx)) out.push_back(s);
          res = out;
        }
        break;
      }
      case MGQ_RAYCAST: {
        std::vector<int> srcs = query(q[3], ctx);
        static const int32_t card[8] = {-1, 0, 1, 0, 0, 1, 0, -1};
        const int32_t* dirs = q[6] ? pool(q[5]) : card;
        int nd = q[6] ? q[6] : 4;
        std::vector<uint8_t> seen(objs.size(), 0);
        for (int s : srcs) {
          Ctx sc = ctx;
          sc.actor = s;
          sc.target = s;
          int range = (int)value(q[4], sc, s);
          if (range <= 0) continue;
          for (int d = 0; d < nd; d++)
            for (int dist = 1; dist <= range; dist++) {
              int r = objs[s].r + dirs[2 * d] * dist, c = objs[s].c + dirs[2 * d + 1] * dist;
              if (!valid(r, c)) break;
              int o = cells[r * W + c];
              if (!o) continue;
              bool blk = false;
              if (q[8] > 0) {
                Ctx bc = ctx;
                bc.target = o;
                for (int i = 0; i < q[8]; i++)
                  if (filter(q[7] + i, bc)) {
                    blk = true;
                    break;
                  }
              }
              if (blk) {
                if (q[9] && !seen[o]) {
                  seen[o] = 1;
                  res.push_back(o);
                }
                break;
              }
              if (!seen[o]) {
                seen[o] = 1;
                res.push_back(o);
              }
            }
        }
        break;
      }
    }
    return apply_limits(res, q, ctx);
  }
  void recompute_mq(int tag, const Ctx& ctx) {  // query_system.cpp:116-175
    Ctx tc = ctx;
    tc.skip_trigger = true;
    std::vector<int> lost, keep;
    const int32_t* mq = sec(MGS_MQ);
    for (int i = 0; i < hdr(MGH_NUM_MQ); i++) {
      if (mq[2 * i] != tag) continue;
      lost = tag_index[tag];
      for (int s : lost) {
        tc.actor = s;
        tc.target = s;
        remove_tag(s, tag, tc);
      }
      std::vector<int> res = query(mq[2 * i + 1], ctx);
      for (int s : res) {
        if (std::find(keep.begin(), keep.end(), s) == keep.end()) keep.push_back(s);
        tc.actor = s;
        tc.target = s;
        add_tag(s, tag, tc);
      }
      break;
    }
    tc.skip_trigger = false;
    for (int s : lost)
      if (std::find(keep.begin(), keep.end(), s) == keep.end()) run_tag_handlers(s, tag, tc);
    // on_tag_add handlers: none can be configured
  }
  void compute_all_mq(const Ctx& ctx) {  // query_system.cpp:91-114
    Ctx tc = ctx;
    tc.skip_trigger = true;
    const int32_t* mq = sec(MGS_MQ);
    for (int i = 0; i < hdr(MGH_NUM_MQ); i++) {
      int tag = mq[2 * i];
      std::vector<int> tagged = tag_index[tag];
      for (int s : tagged) remove_tag(s, tag, tc);
      std::vector<int> res = query(mq[2 * i + 1], ctx);
      for (int s : res) add_tag(s, tag, tc);
    }
  }

  // ---- object creation / removal (core/grid_object_factory.cpp:62-104, mettagrid_c.cpp:222-267) ---
  int create_object(int t, int r, int c, bool with_obs_encoder) {
    const int32_t* tp = tmpl(t);
    int slot = (int)next_id;  // slots and ids coincide in the oracle (ids are never reused)
    if (slot >= proxy0) {
      error |= MGERR_POOL_EXHAUSTED;
      return 0;
    }
    Obj& o = objs[slot];
    o = Obj();
    o.alive = true;
    o.in_grid = true;
    o.tmpl = t;
    o.kind = tp[MGT_KIND];
    o.type_id = tp[MGT_TYPE_ID];
    o.r = r;
    o.c = c;
    o.id = next_id++;
    o.vibe = tp[MGT_VIBE];
    o.obs_inv = with_obs_encoder;
    for (int w = 0; w < TW; w++) o.tags[w] = (uint32_t)pool(tp[MGT_TAGS])[w];
    // initial inventory: stored in emission order (first = most recent); inserting back to front
    // reproduces it (objects/agent.cpp:79-84, grid_object_factory.cpp:83-87)
    const int32_t* iv = pool(tp[MGT_INIT_INV]);
    for (int i = tp[MGT_INIT_INV_N] - 1; i >= 0; i--) inv_update(o, iv[2 * i], iv[2 * i + 1], true, false);
    cells[r * W + c] = slot;
    for (int tg = 0; tg < hdr(MGH_NUM_TAGS); tg++)
      if (has_tag(o, tg)) tag_index[tg].push_back(slot);
    return slot;
  }
  void register_aoe(int slot, int cfg) {
    AoeSource s;
    s.obj = slot;
    s.cfg = cfg;
    s.r = objs[slot].r;
    s.c = objs[slot].c;
    s.alive = true;
    s.inside.assign(A, 0);
    aoe.push_back(s);
  }
  void apply_presence(const int32_t* a, Obj& tgt, int mult) {  // aoe_tracker.cpp:141-145
    const int32_t* pd = pool(a[7]);
    for (int i = 0; i < a[8]; i++) inv_update(tgt, pd[2 * i], pd[2 * i + 1] * mult);
  }
  void unregister_aoe(int slot) {  // aoe_tracker.cpp:207-276
    for (auto& s : aoe) {
      if (!s.alive || s.obj != slot) continue;
      const int32_t* a = sec(MGS_AOES) + s.cfg * MG_AOE_WORDS;
      for (int ag = 0; ag < A; ag++)
        if (s.inside[ag]) {
          apply_presence(a, objs[agents[ag].obj], -1);
          s.inside[ag] = 0;
        }
      s.alive = false;
    }
  }
  void remove_object(int slot) {  // resource_mutation.hpp:88-97
    unregister_aoe(slot);
    Obj& o = objs[slot];
    cells[o.r * W + o.c] = 0;
    o.in_grid = false;
    for (int tg = 0; tg < hdr(MGH_NUM_TAGS); tg++)
      if (has_tag(o, tg)) {
        auto& v = tag_index[tg];
        v.erase(std::remove(v.begin(), v.end(), slot), v.end());
      }
    o.alive = false;
  }
  int spawn(int t, int r, int c) {  // spawn_object_mutation.cpp:27-63
    int slot = create_object(t, r, c, false);
    if (!slot) return 0;
    const int32_t* tp = tmpl(t);
    for (int i = 0; i < tp[MGT_AOES_N]; i++) aoe_pending.push_back({slot, tp[MGT_AOES] + i});
    return slot;
  }

  // ---- mutations (handler/mutations/*.hpp) --------------------------------------------------------
  void mutate(int mi, Ctx& ctx) {
    const int32_t* m = sec(MGS_MUTATIONS) + mi * MG_MUTATION_WORDS;
    int e1 = resolve(ctx, m[1]), e2 = resolve(ctx, m[2]);
    switch (m[0]) {
      case MGM_RESOURCE_DELTA:  // resource_mutation.hpp:23-39
        if (ctx.deferred && m[1] == MGE_TARGET && ctx.target && !is_modifier(objs[ctx.target], m[3])) {
          Deferred& d = *ctx.deferred;
          if (!d.seen[m[3]]) {
            d.seen[m[3]] = true;
            d.order.push_back(m[3]);
          }
          d.delta[m[3]] += m[4];
          return;
        }
        if (e1) inv_update(objs[e1], m[3], m[4]);
        return;
      case MGM_RESOURCE_TRANSFER: {  // resource_mutation.hpp:52-99
        if (!e1 || !e2) return;
        int amount = m[4];
        if (amount < 0) amount = objs[e1].inv[m[3]];
        int moved = transfer(objs[e1], objs[e2], m[3], amount);
        if (moved > 0 && objs[e1].kind == 1 && objs[e1].agent >= 0)
          astat_add(objs[e1].agent, sec(MGS_RES_STATS)[m[3] * 4 + 3], (float)moved);
        if (m[5] && objs[e1].order.empty()) remove_object(e1);
        return;
      }
      case MGM_CLEAR_INVENTORY: {  // resource_mutation.hpp:111-130
        if (!e1) return;
        if (m[4] == 0) {
          std::vector<uint8_t> items = objs[e1].order;
          for (uint8_t it : items) inv_update(objs[e1], it, -(int)objs[e1].inv[it]);
        } else {
          const int32_t* ids = pool(m[3]);
          for (int i = 0; i < m[4]; i++) inv_update(objs[e1], ids[i], -(int)objs[e1].inv[ids[i]]);
        }
        return;
      }
      case MGM_ATTACK: {  // attack_mutation.hpp:20-38
        if (!ctx.actor || !ctx.target) return;
        int weapon = objs[ctx.actor].inv[m[3]], armor = objs[ctx.target].inv[m[4]];
        int dmg = std::max(0, weapon * m[6] / 100 - armor);
        if (dmg > 0) inv_update(objs[ctx.target], m[5], -dmg);
        return;
      }
      case MGM_STATS: {  // stats_mutation.hpp:21-41
        int ent = m[5] ? ctx.actor : ctx.target;
        float v = value(m[6], ctx, ent);
        if (m[4] == 0) {
          gstats[m[3]] = v;
          gtouched[m[3]] = 1;
        } else if (ent && objs[ent].kind == 1 && objs[ent].agent >= 0) {
          astat_set(objs[ent].agent, m[3], v);
        }
        return;
      }
      case MGM_ADD_TAG:
        if (e1) add_tag(e1, m[3], ctx);
        return;
      case MGM_REMOVE_TAG:
        if (e1) remove_tag(e1, m[3], ctx);
        return;
      case MGM_REMOVE_TAGS_PREFIX: {
        if (!e1) return;
        const int32_t* ids = pool(m[3]);
        for (int i = 0; i < m[4]; i++) remove_tag(e1, ids[i], ctx);
        return;
      }
      case MGM_GAME_VALUE: {  // game_value_mutation.hpp:18-25
        float delta = value(m[4], ctx, e1);
        value_update(m[3], ctx, e1, delta);
        return;
      }
      case MGM_RECOMPUTE_MQ:
        recompute_mq(m[3], ctx);
        return;
      case MGM_QUERY_INVENTORY: {  // query_inventory_mutation.hpp:24-52
        std::vector<int> res = query(m[3], ctx);
        const int32_t* pr = pool(m[4]);
        if (m[6]) {
          if (!e1) return;
          for (int s : res)
            for (int i = 0; i < m[5]; i++) {
              int rid = pr[2 * i], d = pr[2 * i + 1], actual = 0;
              if (d > 0)
                actual = transfer(objs[e1], objs[s], rid, d);
              else if (d < 0)
                actual = transfer(objs[s], objs[e1], rid, -d);
              if (actual != 0 && m[7] >= 0 && pool(m[7])[rid] >= 0) gstat_add(pool(m[7])[rid], (float)actual);
            }
        } else {
          for (int s : res)
            for (int i = 0; i < m[5]; i++) inv_update(objs[s], pr[2 * i], pr[2 * i + 1]);
        }
        return;
      }
      case MGM_RELOCATE:  // relocate_mutation.hpp:16-21
        if (ctx.actor && objs[ctx.actor].kind == 1) move_object(ctx.actor, ctx.tr, ctx.tc);
        return;
      case MGM_SWAP: {  // swap_mutation.hpp:15-22, core/grid.hpp:78-92
        if (!ctx.actor || !ctx.target) return;
        Obj &a = objs[ctx.actor], &b = objs[ctx.target];
        if (a.kind != 1 || b.kind != 1) return;
        cells[a.r * W + a.c] = ctx.target;
        cells[b.r * W + b.c] = ctx.actor;
        std::swap(a.r, b.r);
        std::swap(a.c, b.c);
        if (a.agent >= 0) astat_add(a.agent, hdr(MGH_ST_ACTIONS_SWAP), 1.0f);
        return;
      }
      case MGM_USE_TARGET: {  // use_target_mutation.hpp:17-30, core/grid_object.cpp:69-77
        if (!ctx.target || !ctx.actor || objs[ctx.actor].kind != 1) {
          ctx.failed = true;
          return;
        }
        Obj& t = objs[ctx.target];
        int h = t.tmpl >= 0 ? tmpl(t.tmpl)[MGT_ON_USE] : -1;
        bool ok = false;
        if (h >= 0) {
          Ctx u = ctx;
          ok = handler_apply(h, u);
        }
        if (!ok) {
          ctx.failed = true;
          return;
        }
        int after = tmpl(objs[ctx.actor].tmpl)[MGT_ON_AFTER_USE];
        if (after >= 0) handler_apply(after, ctx);  // shares ctx: may set ctx.failed (objects/agent.cpp:73-77)
        return;
      }
      case MGM_SPAWN_OBJECT: {
        if (m[3] < 0 || !valid(ctx.tr, ctx.tc) || cells[ctx.tr * W + ctx.tc] != 0) {
          ctx.failed = true;
          return;
        }
        int s = spawn(m[3], ctx.tr, ctx.tc);
        if (!s) {
          ctx.failed = true;
          return;
        }
        ctx.target = s;
        return;
      }
      case MGM_RAYCAST_SPAWN: {  // raycast_spawn_mutation.cpp:16-93
        if (!ctx.target || m[3] < 0) {
          ctx.failed = true;
          return;
        }
        int orr = objs[ctx.target].r, oc = objs[ctx.target].c;
        int range = (int)value(m[6], ctx, ctx.target);
        if (range <= 0) return;
        const int32_t* dirs = pool(m[4]);
        for (int d = 0; d < m[5]; d++)
          for (int dist = 1; dist <= range; dist++) {
            int r = orr + dirs[2 * d] * dist, c = oc + dirs[2 * d + 1] * dist;
            if (!valid(r, c)) break;
            int ex = cells[r * W + c];
            if (ex) {
              bool blk = false;
              Ctx bc = ctx;
              bc.target = ex;
              for (int i = 0; i < m[2]; i++)
                if (filter(m[7] + i, bc)) {
                  blk = true;
                  break;
                }
              if (blk) break;
              continue;
            }
            spawn(m[3], r, c);
          }
        return;
      }
      case MGM_CHANGE_VIBE:
        if (e1) objs[e1].vibe = m[3];
        return;
      case MGM_PUSH_OBJECT: {  // push_object_mutation.hpp:27-57
        if (!ctx.actor || !ctx.target) {
          ctx.failed = true;
          return;
        }
        Obj &a = objs[ctx.actor], &t = objs[ctx.target];
        int dr = std::min(std::max(t.r - a.r, -1), 1), dc = std::min(std::max(t.c - a.c, -1), 1);
        int nr = t.r + dr, nc = t.c + dc;
        if (!valid(nr, nc) || cells[nr * W + nc] != 0 || !move_object(ctx.target, nr, nc)) ctx.failed = true;
        return;
      }
    }
  }

  // ---- handlers (handler/handler.cpp:76-103, multi_handler.cpp:8-21) ------------------------------
  bool handler_apply(int h, Ctx& ctx) {
    const int32_t* hd = sec(MGS_HANDLERS) + h * MG_HANDLER_WORDS;
    if (hd[0] == MGHK_SIMPLE) {
      if (!filters_pass(hd[1], hd[2], ctx)) return false;
      ctx.failed = false;
      for (int i = 0; i < hd[4]; i++) {
        mutate(hd[3] + i, ctx);
        if (ctx.failed) return false;
      }
      return true;
    }
    bool any = false;
    const int32_t* kids = pool(hd[1]);
    for (int i = 0; i < hd[2]; i++)
      if (handler_apply(kids[i], ctx)) {
        any = true;
        if (hd[0] == MGHK_FIRST_MATCH) return true;
      }
    return any;
  }

  // ---- actions (actions/action_handler.hpp:78-105, move.hpp:81-115, change_vibe.hpp:48-57) ----------
  bool do_action(int a, const int32_t* act) {
    Agent& ag = agents[a];
    int s = ag.obj;
    switch (act[0]) {
      case MGA_NOOP:
        return true;
      case MGA_CHANGE_VIBE:
        objs[s].vibe = act[1];
        return true;
      case MGA_MOVE: {
        static const int DC[8] = {0, 0, -1, 1, -1, 1, -1, 1};  // actions/orientation.hpp:28-48
        static const int DR[8] = {-1, 1, 0, 0, -1, -1, 1, 1};
        int dr = DR[act[1]], dc = DC[act[1]];
        const int32_t* chain = sec(MGS_MOVE_CHAIN);
        for (int k = 0; k < hdr(MGH_NUM_MOVE_HANDLERS); k++) {
          const int32_t* mh = chain + k * MG_MOVEH_WORDS;
          for (int i = 1; i <= mh[1]; i++) {
            int tr = objs[s].r + dr * i, tc = objs[s].c + dc * i;
            if (!valid(tr, tc)) break;
            int t = cells[tr * W + tc];
            if (!t && !mh[2]) continue;
            Ctx ctx;
            ctx.actor = s;
            ctx.target = t;
            ctx.tr = tr;
            ctx.tc = tc;
            ctx.distance = i;
            ctx.move_dir = act[1];
            if (handler_apply(mh[0], ctx)) return true;
            break;
          }
        }
        return false;
      }
    }
    return false;
  }
  bool handle_action(int a, const int32_t* act) {
    Agent& ag = agents[a];
    bool ok = do_action(a, act);
    Obj& o = objs[ag.obj];
    if (o.r == ag.prev_r && o.c == ag.prev_c) {
      ag.swm += 1;
      int id = hdr(MGH_ST_MAX_SWM);
      if ((float)ag.swm > ag.stats[id]) astat_set(a, id, (float)ag.swm);  // stats.get() does not create the key
    } else {
      ag.swm = 0;
    }
    ag.prev_r = o.r;
    ag.prev_c = o.c;
    static const int SK[3] = {MGH_ST_NOOP_SUCCESS, MGH_ST_MOVE_SUCCESS, MGH_ST_VIBE_SUCCESS};
    static const int FK[3] = {MGH_ST_NOOP_FAILED, MGH_ST_MOVE_FAILED, MGH_ST_VIBE_FAILED};
    if (ok) {
      astat_add(a, hdr(SK[act[0]]), 1.0f);
    } else {
      astat_add(a, hdr(FK[act[0]]), 1.0f);
      astat_add(a, hdr(MGH_ST_ACTION_FAILED), 1.0f);
    }
    return ok;
  }

  // ---- events (handler/event.cpp:34-99, event_scheduler.cpp:36-53) ----------------------------------
  int event_execute(int ev) {
    const int32_t* e = sec(MGS_EVENTS) + ev * MG_EVENT_WORDS;
    Ctx g;
    std::vector<int> targets = query(e[0], g);
    if (e[1] >= 0 && (int)targets.size() > e[1]) rng.shuffle(targets);
    int applied = 0;
    for (int t : targets) {
      if (e[1] >= 0 && applied >= e[1]) break;
      Ctx c;
      c.actor = t;
      c.target = t;
      c.tr = objs[t].r;
      c.tc = objs[t].c;
      if (!filters_pass(e[2], e[3], c)) continue;
      for (int i = 0; i < e[5]; i++) mutate(e[4] + i, c);
      applied++;
    }
    if (applied == 0 && e[6] >= 0) return event_execute(e[6]);
    return applied;
  }

  // ---- AOE (core/aoe_tracker.cpp:166-200,278-415) ----------------------------------------------------
  bool aoe_is_territory(const int32_t* a) const { return a[6] == 0 && a[8] == 0 && a[0] > 0; }
  bool aoe_covers(const AoeSource& s, int r, int c) const {
    const int32_t* a = sec(MGS_AOES) + s.cfg * MG_AOE_WORDS;
    int64_t range = a[0], dr = r - s.r, dc = c - s.c;
    if (dr < -range || dr > range || dc < -range || dc > range) return false;
    int64_t d2 = dr * dr + dc * dc;
    if (d2 > range * range) return false;
    if (aoe_is_territory(a) && range >= 2 && d2 == range * range && (dr == 0 || dc == 0)) return false;
    return true;
  }
  bool aoe_passes(const AoeSource& s, const int32_t* a, int target, const Ctx& base) {
    Ctx c = base;
    c.actor = s.obj;
    c.target = target;
    return filters_pass(a[3], a[4], c);
  }
  void aoe_apply(const AoeSource& s, const int32_t* a, int target, const Ctx& base) {  // AOESource::try_apply
    Ctx c = base;
    c.actor = s.obj;
    c.target = target;
    if (!filters_pass(a[3], a[4], c)) return;
    for (int i = 0; i < a[6]; i++) mutate(a[5] + i, c);
  }
  void aoe_apply_fixed(int ag) {
    int target = agents[ag].obj;
    Deferred d;
    memset(d.delta, 0, sizeof d.delta);
    memset(d.seen, 0, sizeof d.seen);
    Ctx base;
    base.target = target;
    base.deferred = &d;
    int tr = objs[target].r, tc = objs[target].c;
    // exits: sources the target was inside that no longer cover its cell (registration order)
    for (auto& s : aoe) {
      if (!s.alive) continue;
      const int32_t* a = sec(MGS_AOES) + s.cfg * MG_AOE_WORDS;
      if (!a[1]) continue;
      if (s.inside[ag] && !aoe_covers(s, tr, tc)) {
        s.inside[ag] = 0;
        apply_presence(a, objs[target], -1);
      }
    }
    for (size_t k = 0; k < aoe.size(); k++) {
      if (!aoe[k].alive) continue;
      const int32_t* a = sec(MGS_AOES) + aoe[k].cfg * MG_AOE_WORDS;
      if (!a[1] || !aoe_covers(aoe[k], tr, tc)) continue;
      if (a[6] == 0 && a[8] == 0) continue;
      bool skip_self = !a[2] && aoe[k].obj == target;
      bool now = !skip_self && aoe_passes(aoe[k], a, target, base);
      bool was = aoe[k].inside[ag];
      if (now && !was) {
        aoe[k].inside[ag] = 1;
        apply_presence(a, objs[target], +1);
      } else if (!now && was) {
        aoe[k].inside[ag] = 0;
        apply_presence(a, objs[target], -1);
      }
      if (now && a[6] > 0) aoe_apply(aoe[k], a, target, base);
    }
    for (int rid : d.order)
      if (d.delta[rid] != 0) inv_update(objs[target], rid, d.delta[rid]);
  }
  void aoe_apply_mobile() {
    Ctx base;
    for (size_t k = 0; k < aoe.size(); k++) {
      if (!aoe[k].alive) continue;
      const int32_t* a = sec(MGS_AOES) + aoe[k].cfg * MG_AOE_WORDS;
      if (a[1]) continue;
      int so = aoe[k].obj;
      int64_t range = a[0];
      for (int ag = 0; ag < A; ag++) {
        int t = agents[ag].obj;
        if (!a[2] && so == t) continue;
        bool was = aoe[k].inside[ag];
        int64_t dr = objs[so].r - objs[t].r, dc = objs[so].c - objs[t].c;
        if (dr * dr + dc * dc > range * range) {
          if (was) {
            aoe[k].inside[ag] = 0;
            apply_presence(a, objs[t], -1);
          }
          continue;
        }
        bool now = aoe_passes(aoe[k], a, t, base);
        if (now) {
          if (!was) {
            aoe[k].inside[ag] = 1;
            apply_presence(a, objs[t], +1);
          }
          if (a[6] > 0) aoe_apply(aoe[k], a, t, base);
        } else if (was) {
          aoe[k].inside[ag] = 0;
          apply_presence(a, objs[t], -1);
        }
      }
    }
  }

  // ---- territory (core/territory_tracker.cpp:20-50,215-346) ----------------------------------------
  static uint64_t floor_sqrt_u64(uint64_t v) {
    uint64_t root = 0, bit = 1ULL << 62;
    while (bit > v) bit >>= 2;
    while (bit != 0) {
      if (v >= root + bit) {
        v -= root + bit;
        root = (root >> 1) + bit;
      } else {
        root >>= 1;
      }
      bit >>= 2;
    }
    return root;
  }
  int cell_owner(int r, int c, int ti) const {
    const int32_t* T_ = sec(MGS_TERRITORIES) + ti * MG_TERR_WORDS;
    const int32_t* pre = pool(T_[0]);
    int64_t score[256];
    bool used[256];
    memset(used, 0, sizeof used);
    for (const auto& s : terr) {
      if (s.terr != ti) continue;
      const Obj& o = objs[s.obj];
      int range = s.decay > 0 ? s.strength / s.decay : s.strength;
      int64_t dr = r - o.r, dc = c - o.c;
      if (dr < -range || dr > range || dc < -range || dc > range) continue;
      int64_t d2 = dr * dr + dc * dc;
      if (d2 > (int64_t)range * range) continue;
      int tag = -1;
      for (int i = 0; i < T_[1]; i++)
        if (has_tag(o, pre[i])) {
          tag = pre[i];
          break;
        }
      if (tag < 0) continue;
      int64_t sc = (int64_t)s.strength * 1024 - (int64_t)s.decay * (int64_t)floor_sqrt_u64((uint64_t)d2 * 1024 * 1024);
      if (sc <= 0) continue;
      if (!used[tag]) {
        used[tag] = true;
        score[tag] = 0;
      }
      score[tag] += sc;
    }
    int win = -1;
    int64_t best = 0;
    bool tied = false;
    for (int t = 0; t < 256; t++) {
      if (!used[t]) continue;
      if (score[t] > best) {
        win = t;
        best = score[t];
        tied = false;
      } else if (score[t] == best && win >= 0) {
        tied = true;
      }
    }
    return tied ? -1 : win;
  }
  int territory_mask(int r, int c, const Obj& observer) const {  // territory_tracker.cpp:254-273
    for (int ti = 0; ti < hdr(MGH_NUM_TERRITORIES); ti++) {
      int w = cell_owner(r, c, ti);
      if (w < 0) continue;
      return has_tag(observer, w) ? 1 : 2;
    }
    return 0;
  }
  void terr_run(int ti, int list_off, int n, int tag, int target) {
    Obj& px = objs[proxy0 + ti];
    memset(px.tags, 0, sizeof px.tags);
    px.tags[tag >> 5] |= 1u << (tag & 31);
    px.r = objs[target].r;
    px.c = objs[target].c;
    const int32_t* hs = pool(list_off);
    for (int i = 0; i < n; i++) {
      const int32_t* hd = sec(MGS_HANDLERS) + hs[i] * MG_HANDLER_WORDS;
      Ctx c;
      c.actor = proxy0 + ti;
      c.target = target;
      if (filters_pass(hd[1], hd[2], c))
        for (int k = 0; k < hd[4]; k++) mutate(hd[3] + k, c);
    }
  }
  void terr_apply(int ag) {
    int target = agents[ag].obj;
    for (int ti = 0; ti < hdr(MGH_NUM_TERRITORIES); ti++) {
      const int32_t* T_ = sec(MGS_TERRITORIES) + ti * MG_TERR_WORDS;
      int cur = cell_owner(objs[target].r, objs[target].c, ti);
      int prev = inside_tag[ag][ti];
      if (prev != cur && prev >= 0) terr_run(ti, T_[4], T_[5], prev, target);
      if (prev != cur && cur >= 0) terr_run(ti, T_[2], T_[3], cur, target);
      inside_tag[ag][ti] = cur;
      if (cur >= 0) terr_run(ti, T_[6], T_[7], cur, target);
    }
  }

  // ---- observations (bindings/mettagrid_c.cpp:665-824,1207-1238) ---------------------------------------
  int digits(uint32_t v) const {
    int n = 1;
    v /= (uint32_t)B;
    while (v > 0) {
      v /= (uint32_t)B;
      n++;
    }
    return n;
  }
  void observe(int a, int action) {
    Agent& ag = agents[a];
    const Obj& me = objs[ag.obj];
    uint8_t* out = &obs[(size_t)a * T * 3];
    size_t attempted = 0;
    auto emit = [&](int loc, int feat, int val) {
      if (attempted < (size_t)T) {
        out[attempted * 3 + 0] = (uint8_t)loc;
        out[attempted * 3 + 1] = (uint8_t)feat;
        out[attempted * 3 + 2] = (uint8_t)val;
      }
      attempted++;
    };
    int flags = hdr(MGH_GLOBAL_FLAGS);
    int max_steps = hdr(MGH_MAX_STEPS);
    if (flags & MGG_EPISODE_PCT) {
      uint8_t pct = 0;
      if (max_steps > 0) pct = step >= (uint32_t)max_steps ? 255 : (uint8_t)(256u * step / (uint32_t)max_steps);
      emit(0xFE, hdr(MGH_FEAT_EPISODE_PCT), pct);
    }
    if (flags & MGG_LAST_ACTION) emit(0xFE, hdr(MGH_FEAT_LAST_ACTION), (uint8_t)action);
    if ((flags & MGG_LAST_ACTION_MOVE) && hdr(MGH_FEAT_LAST_ACTION_MOVE) != 0)
      emit(0xFE, hdr(MGH_FEAT_LAST_ACTION_MOVE), (me.r != ag.step_r || me.c != ag.step_c) ? 1 : 0);
    if (flags & MGG_LAST_REWARD) emit(0xFE, hdr(MGH_FEAT_LAST_REWARD), (uint8_t)roundf(rewards[a] * 100.0f));
    if (flags & MGG_LOCAL_POSITION) {
      int dc = me.c - ag.spawn_c, dr = ag.spawn_r - me.r;
      if (dc > 0)
        emit(0xFE, hdr(MGH_FEAT_LP_EAST), std::min(dc, 255));
      else if (dc < 0)
        emit(0xFE, hdr(MGH_FEAT_LP_WEST), std::min(-dc, 255));
      if (dr > 0)
        emit(0xFE, hdr(MGH_FEAT_LP_NORTH), std::min(dr, 255));
      else if (dr < 0)
        emit(0xFE, hdr(MGH_FEAT_LP_SOUTH), std::min(-dr, 255));
    }
    const int32_t* ov = sec(MGS_OBS_VALUES);
    for (int i = 0; i < hdr(MGH_NUM_OBS_VALUES); i++) {
      Ctx c;
      c.actor = ag.obj;
      c.target = ag.obj;
      uint32_t enc = (uint32_t)value(ov[2 * i + 1], c, ag.obj);
      int feat = ov[2 * i];
      emit(0xFE, feat, enc % (uint32_t)B);
      enc /= (uint32_t)B;
      while (enc > 0) {
        feat++;
        emit(0xFE, feat, enc % (uint32_t)B);
        enc /= (uint32_t)B;
      }
    }
    int rr = hdr(MGH_OBS_H) >> 1, cr = hdr(MGH_OBS_W) >> 1;
    const int32_t* offs = sec(MGS_OFFSETS);
    const int32_t* invf = sec(MGS_INV_FEATS);
    for (int k = 0; k < hdr(MGH_NUM_OFFSETS); k++) {
      int r = me.r + offs[2 * k], c = me.c + offs[2 * k + 1];
      if (!valid(r, c)) continue;
      int loc = ((offs[2 * k] + rr) << 4) | ((offs[2 * k + 1] + cr) & 15);
      if (hdr(MGH_FEAT_AOE_MASK) != 0) {
        int m = territory_mask(r, c, me);
        if (m) emit(loc, hdr(MGH_FEAT_AOE_MASK), m);
      }
      int s = cells[r * W + c];
      if (!s) continue;
      Obj& o = objs[s];
      if (o.visited < step) {  // cell staleness (mettagrid_c.cpp:787-796)
        astat_add(a, hdr(MGH_ST_CELL_VISITED), (float)(step - o.visited));
        o.visited = step;
      }
      // core/grid_object.cpp:178-203, objects/agent.cpp:142-154
      for (int t = 0; t < hdr(MGH_NUM_TAGS); t++)
        if (has_tag(o, t)) emit(loc, hdr(MGH_FEAT_TAG), t);
      if (o.vibe != 0) emit(loc, hdr(MGH_FEAT_VIBE), o.vibe);
      if (o.obs_inv)
        for (uint8_t it : o.order) {
          uint32_t amt = o.inv[it];
          int nd = std::min(digits(amt), ND);
          for (int p = 0; p < nd; p++) {
            emit(loc, invf[it * ND + p], amt % (uint32_t)B);
            amt /= (uint32_t)B;
          }
        }
      if (o.kind == 1) {  // a spawned Agent is not in _agents: group from its config, agent_id stays 0
        emit(loc, hdr(MGH_FEAT_GROUP), o.agent >= 0 ? agents[o.agent].group : tmpl(o.tmpl)[MGT_GROUP]);
        emit(loc, hdr(MGH_FEAT_AGENT_ID), o.agent >= 0 ? o.agent : 0);
      }
    }
    if (attempted > (size_t)T) {  // hard error in the reference (mettagrid_c.cpp:364-375,813-819)
      if (!(error & MGERR_TOKEN_OVERFLOW)) err_info = a | ((int)std::min(attempted, (size_t)65535) << 16);
      error |= MGERR_TOKEN_OVERFLOW;
      return;
    }
    gstat_add(hdr(MGH_GST_TOKENS_WRITTEN), (float)attempted);
    gstat_add(hdr(MGH_GST_TOKENS_DROPPED), 0.0f);
    gstat_add(hdr(MGH_GST_TOKENS_FREE), (float)(T - attempted));
  }
  void observe_all(const std::vector<int32_t>& executed) {
    for (int a = 0; a < A; a++) observe(a, executed[a]);
  }

  // ---- coverage (objects/agent.cpp:41-57) ---------------------------------------------------------------
  void reset_coverage(int a) {
    Agent& ag = agents[a];
    std::fill(ag.seen.begin(), ag.seen.end(), 0);
    ag.max_dist = 0;
    const Obj& o = objs[ag.obj];
    ag.seen[o.r * W + o.c] = 1;
    ag.unique = 1;
    astat_set(a, hdr(MGH_ST_UNIQUE_VISITED), 1.0f);
    astat_set(a, hdr(MGH_ST_MAX_DIST), 0.0f);
  }
  void track_coverage(int a) {
    Agent& ag = agents[a];
    const Obj& o = objs[ag.obj];
    if (!ag.seen[o.r * W + o.c]) {
      ag.seen[o.r * W + o.c] = 1;
      ag.unique++;
    }
    astat_set(a, hdr(MGH_ST_UNIQUE_VISITED), (float)ag.unique);
    uint32_t d = (uint32_t)(abs(ag.spawn_r - o.r) + abs(o.c - ag.spawn_c));
    ag.max_dist = std::max(ag.max_dist, d);
    astat_set(a, hdr(MGH_ST_MAX_DIST), (float)ag.max_dist);
  }

  // ---- construction (bindings/mettagrid_c.cpp:42-191,200-319) --------------------------------------------
  void init_buffers() {  // _init_buffers :294-319 (re-run by every set_buffers call)
    for (int a = 0; a < A; a++) reset_coverage(a);  // Agent::init via set_buffers :1178-1180
    std::fill(terminals.begin(), terminals.end(), 0);
    std::fill(truncations.begin(), truncations.end(), 0);
    std::fill(episode_rewards.begin(), episode_rewards.end(), 0.0f);
    std::fill(rewards.begin(), rewards.end(), 0.0f);
    std::fill(obs.begin(), obs.end(), 0xFF);
    std::vector<int32_t> ex(A, 0);
    observe_all(ex);
  }
  bool create(const int32_t* blob, int nwords, const int16_t* init_cells, uint32_t seed, const float* init_gstats) {
    P.assign(blob, blob + nwords);
    p = P.data();
    if ((uint32_t)p[MGH_MAGIC] != MG_MAGIC || p[MGH_VERSION] != MG_VERSION) return false;
    H = hdr(MGH_H), W = hdr(MGH_W), A = hdr(MGH_NUM_AGENTS), T = hdr(MGH_NUM_TOKENS), R = hdr(MGH_NUM_RESOURCES);
    TW = hdr(MGH_TAG_WORDS), B = hdr(MGH_TOKEN_BASE), ND = hdr(MGH_INV_DIGITS);
    rng.seed(seed);
    cells.assign((size_t)H * W, 0);
    int NT = hdr(MGH_NUM_TERRITORIES);
    proxy0 = hdr(MGH_MAX_OBJECTS);
    objs.assign((size_t)proxy0 + NT, Obj());
    for (int ti = 0; ti < NT; ti++) objs[proxy0 + ti].kind = 3;
    tag_index.assign(256, {});
    int SG = hdr(MGH_NUM_GAME_STATS), SA = hdr(MGH_NUM_AGENT_STATS);
    gstats.assign(SG, 0.0f);
    gtouched.assign(SG, 0);
    gtouched[hdr(MGH_GST_TOKENS_WRITTEN)] = gtouched[hdr(MGH_GST_TOKENS_DROPPED)] = gtouched[hdr(MGH_GST_TOKENS_FREE)] = 1;
    if (init_gstats)
      for (int i = 0; i < SG; i++)
        if (init_gstats[i] != 0.0f) {
          gstats[i] = init_gstats[i];
          gtouched[i] = 1;
        }
    obs.assign((size_t)A * T * 3, 0xFF);
    rewards.assign(A, 0), episode_rewards.assign(A, 0), terminals.assign(A, 0), truncations.assign(A, 0);
    success.assign(A, 0), last_actions.assign(A, 0);
    // _init_grid: row-major, ids from 1, agent ids in encounter order
    for (int r = 0; r < H; r++)
      for (int c = 0; c < W; c++) {
        int t = init_cells[r * W + c];
        if (t < 0) continue;
        const int32_t* tp = tmpl(t);
        int slot = create_object(t, r, c, true);
        if (!slot) return false;
        for (int i = 0; i < tp[MGT_AOES_N]; i++) register_aoe(slot, tp[MGT_AOES] + i);
        const int32_t* tc = pool(tp[MGT_TERR]);
        for (int i = 0; i < tp[MGT_TERR_N]; i++) terr.push_back({slot, tc[3 * i], tc[3 * i + 1], tc[3 * i + 2]});
        if (tp[MGT_KIND] == 1) {
          if ((int)agents.size() >= A) return false;
          Agent ag;
          ag.obj = slot;
          ag.group = tp[MGT_GROUP];
          ag.spawn_r = ag.prev_r = ag.step_r = r;
          ag.spawn_c = ag.prev_c = ag.step_c = c;
          ag.seen.assign((size_t)H * W, 0);
          ag.stats.assign(SA, 0.0f);
          ag.touched.assign(SA, 0);
          ag.reward_prev.assign(tp[MGT_REWARDS_N], 0.0f);
          objs[slot].agent = (int)agents.size();
          agents.push_back(ag);
          int a = (int)agents.size() - 1;
          // populate_initial_inventory sets "<res>.amount" for every configured item (agent.cpp:79-84)
          const int32_t* iv = pool(tp[MGT_INIT_INV]);
          for (int i = 0; i < tp[MGT_INIT_INV_N]; i++)
            astat_set(a, sec(MGS_RES_STATS)[iv[2 * i] * 4 + 2], (float)iv[2 * i + 1]);
          reset_coverage(a);  // add_agent -> Agent::init (mettagrid_c.cpp:332-335)
        }
      }
    if ((int)agents.size() != A) return false;
    inside_tag.assign(A, std::vector<int>(NT, -1));
    for (auto& s : aoe) s.inside.assign(A, 0);
    Ctx g;
    compute_all_mq(g);
    // init_reward: top-level StatValue entries create their stat key (core/game_value.cpp:40-47)
    for (int a = 0; a < A; a++) {
      const int32_t* tp = tmpl(objs[agents[a].obj].tmpl);
      const int32_t* rw = pool(tp[MGT_REWARDS]);
      for (int i = 0; i < tp[MGT_REWARDS_N]; i++) {
        const int32_t* v = sec(MGS_VALUES) + rw[2 * i] * MG_VALUE_WORDS;
        if (v[0] == MGV_STAT) {
          if (v[1] == MGSC_GAME)
            gtouched[v[2]] = 1;
          else
            agents[a].touched[v[2]] = 1;
        }
      }
    }
    init_buffers();
    return true;
  }

  // ---- the tick (bindings/mettagrid_c.cpp:921-1102) ---------------------------------------------------------
  void tick(const int32_t* actions, const int32_t* vibe_actions) {
    for (int a = 0; a < A; a++) {
      agents[a].step_r = objs[agents[a].obj].r;
      agents[a].step_c = objs[agents[a].obj].c;
    }
    std::fill(rewards.begin(), rewards.end(), 0.0f);
    std::fill(obs.begin(), obs.end(), 0xFF);
    std::fill(success.begin(), success.end(), 0);
    step++;
    std::vector<int> order(A);
    for (int a = 0; a < A; a++) order[a] = a;
    rng.shuffle(order);
    std::vector<int32_t> executed(A, 0);
    int NA = hdr(MGH_NUM_ACTIONS);
    for (int off = 0; off <= hdr(MGH_MAX_PRIORITY); off++) {
      int prio = hdr(MGH_MAX_PRIORITY) - off;
      for (int stream = 0; stream < 2; stream++)
        for (int a : order) {
          int32_t idx = stream ? vibe_actions[a] : actions[a];
          if (idx < 0 || idx >= NA) {  // _handle_invalid_action :914-919 (once per priority pass)
            astat_add(a, hdr(MGH_ST_INVALID_INDEX), 1.0f);
            success[a] = 0;
            continue;
          }
          const int32_t* act = sec(MGS_ACTIONS) + idx * MG_ACTION_WORDS;
          if (act[3] != stream) continue;
          if (act[2] != prio) continue;
          if (handle_action(a, act)) {
            executed[a] = idx;
            success[a] = 1;
          }
        }
    }
    const int32_t* sch = sec(MGS_SCHEDULE);
    while (event_cursor < hdr(MGH_NUM_EVENTS_SCHED) && sch[2 * event_cursor] <= (int)step) {
      event_execute(sch[2 * event_cursor + 1]);
      event_cursor++;
    }
    for (int a = 0; a < A; a++) {
      int h = tmpl(objs[agents[a].obj].tmpl)[MGT_ON_TICK];
      if (h >= 0) {
        Ctx c;
        c.actor = c.target = agents[a].obj;
        handler_apply(h, c);
      }
    }
    for (int a = 0; a < A; a++) {
      aoe_apply_fixed(a);
      terr_apply(a);
    }
    aoe_apply_mobile();
    for (auto& pr : aoe_pending) register_aoe(pr.first, pr.second);
    aoe_pending.clear();
    if (hdr(MGH_GAME_ON_TICK) >= 0) {
      Ctx c;
      handler_apply(hdr(MGH_GAME_ON_TICK), c);
    }
    for (int a = 0; a < A; a++) track_coverage(a);
    last_actions = executed;
    observe_all(executed);
    for (int a = 0; a < A; a++) {  // systems/reward.hpp:56-77
      const int32_t* tp = tmpl(objs[agents[a].obj].tmpl);
      const int32_t* rw = pool(tp[MGT_REWARDS]);
      if (tp[MGT_REWARDS_N] > 0) {
        float total = 0.0f;
        Ctx c;
        c.actor = c.target = agents[a].obj;
        for (int i = 0; i < tp[MGT_REWARDS_N]; i++) {
          float v = value(rw[2 * i], c, agents[a].obj);
          if (rw[2 * i + 1])
            total += v;
          else
            total += v - agents[a].reward_prev[i];
          agents[a].reward_prev[i] = v;
        }
        if (total != 0.0f) rewards[a] += total;
      }
      episode_rewards[a] += rewards[a];
    }
    int ms = hdr(MGH_MAX_STEPS);
    if (ms > 0 && step >= (uint32_t)ms) {
      if (hdr(MGH_EPISODE_TRUNCATES))
        std::fill(truncations.begin(), truncations.end(), 1);
      else
        std::fill(terminals.begin(), terminals.end(), 1);
    }
  }
};

}  // namespace

extern "C" {
void* mgo_create(const int32_t* blob, int nwords, const int16_t* init_cells, uint32_t seed, const float* init_gstats) {
  Env* e = new Env();
  if (!e->create(blob, nwords, init_cells, seed, init_gstats)) {
    delete e;
    return nullptr;
  }
  return e;
}
void mgo_destroy(void* h) { delete (Env*)h; }
void mgo_step(void* h, const int32_t* actions, const int32_t* vibe_actions) { ((Env*)h)->tick(actions, vibe_actions); }
void mgo_reinit_buffers(void* h) { ((Env*)h)->init_buffers(); }
const uint8_t* mgo_obs(void* h) { return ((Env*)h)->obs.data(); }
const float* mgo_rewards(void* h) { return ((Env*)h)->rewards.data(); }
const float* mgo_episode_rewards(void* h) { return ((Env*)h)->episode_rewards.data(); }
const uint8_t* mgo_terminals(void* h) { return ((Env*)h)->terminals.data(); }
const uint8_t* mgo_truncations(void* h) { return ((Env*)h)->truncations.data(); }
const uint8_t* mgo_success(void* h) { return ((Env*)h)->success.data(); }
int mgo_error(void* h) { return ((Env*)h)->error; }
int mgo_error_info(void* h) { return ((Env*)h)->err_info; }
uint32_t mgo_current_step(void* h) { return ((Env*)h)->step; }
void mgo_agent_stats(void* h, float* vals, uint8_t* touched) {
  Env* e = (Env*)h;
  int S = e->hdr(MGH_NUM_AGENT_STATS);
  for (int a = 0; a < e->A; a++)
    for (int i = 0; i < S; i++) {
      vals[a * S + i] = e->agents[a].stats[i];
      touched[a * S + i] = e->agents[a].touched[i];
    }
}
void mgo_game_stats(void* h, float* vals, uint8_t* touched) {
  Env* e = (Env*)h;
  for (size_t i = 0; i < e->gstats.size(); i++) {
    vals[i] = e->gstats[i];
    touched[i] = e->gtouched[i];
  }
}
// Object dump, one row of (8 + R + R) int32 per live object in id order:
//   id, type_id, r, c, vibe, agent index (-1), tag word 0, number of present resources, inv[R], order[R] (-1 padded)
int mgo_dump_objects(void* h, int32_t* out, int max_rows) {
  Env* e = (Env*)h;
  int R = e->R, n = 0, stride = 8 + 2 * R + e->TW;
  for (int s = 1; s < e->proxy0 && n < max_rows; s++) {
    const Obj& o = e->objs[s];
    if (!o.alive) continue;
    int32_t* row = out + (size_t)n * stride;
    row[0] = (int32_t)o.id, row[1] = o.type_id, row[2] = o.r, row[3] = o.c, row[4] = o.vibe, row[5] = o.agent;
    row[6] = (int32_t)o.tags[0], row[7] = (int32_t)o.order.size();
    for (int i = 0; i < R; i++) row[8 + i] = o.inv[i];
    for (int i = 0; i < R; i++) row[8 + R + i] = i < (int)o.order.size() ? o.order[i] : -1;
    for (int k = 0; k < e->TW; k++) row[8 + 2 * R + k] = (int32_t)o.tags[k];
    n++;
  }
  return n;
}
// per agent: object id, group, steps_without_motion, RewardHelper::current_reward (systems/reward.hpp:36-42)
void mgo_agent_state(void* h, int32_t* out) {
  Env* e = (Env*)h;
  for (int a = 0; a < e->A; a++) {
    const Agent& ag = e->agents[a];
    float total = 0.0f;
    for (float v : ag.reward_prev) total += v;
    out[4 * a + 0] = (int32_t)e->objs[ag.obj].id;
    out[4 * a + 1] = ag.group;
    out[4 * a + 2] = (int32_t)ag.swm;
    memcpy(&out[4 * a + 3], &total, 4);
  }
}
// items/amounts in the iteration order of the pybind-built unordered_map (compiler.pybind_dict_order)
void mgo_set_inventory(void* h, int agent, const int32_t* items, const int32_t* amounts, int n) {  // objects/agent.cpp:86-104
  Env* e = (Env*)h;
  Obj& o = e->objs[e->agents[agent].obj];
  std::vector<uint8_t> existing = o.order;
  for (uint8_t it : existing) {
    e->inv_update(o, it, -(int)o.inv[it]);
    e->astat_set(agent, e->sec(MGS_RES_STATS)[it * 4 + 2], 0.0f);
  }
  for (int i = 0; i < n; i++) e->inv_update(o, items[i], amounts[i] - (int)o.inv[items[i]]);
}
// raw MT19937 + shuffle access for the H1 known-answer tests
void mgo_test_shuffle(uint32_t seed, int n, int32_t* out, uint32_t* next_raw) {
  Mt m;
  m.seed(seed);
  std::vector<int> v(n);
  for (int i = 0; i < n; i++) v[i] = i;
  m.shuffle(v);
  for (int i = 0; i < n; i++) out[i] = v[i];
  *next_raw = m.next();
}
}
