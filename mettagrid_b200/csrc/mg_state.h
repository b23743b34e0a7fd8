// mg_state.h -- device-state descriptor shared by the kernels and the C-ABI host code.
#pragma once
#include <stdint.h>

#include "../../include/mg_program.h"

#define MG_FULL 0xffffffffu
#ifndef MG_WARPS_PER_CTA
#define MG_WARPS_PER_CTA 4
#endif
#ifndef MG_FAST_WARPS
#define MG_FAST_WARPS 0  // warps per CTA of k_step_fast (mg_fast.cu): 0 = chosen by grid size (1 or 4), 1 / 4 = forced (A/B builds)
#endif
#define MG_RNG_WINDOW 32
#define MG_TAG_LIST_CAP 64

// the program header as a kernel argument (constant bank)
struct MgFastHdr {
  int v[MGH_HEADER_WORDS];
};

// Everything a kernel needs, passed by value.
struct MgDev {
  const int32_t* P;  // compiled program (global, read-only)
  int num_envs;
  int env0;  // first env of this launch (0 unless mg_step splits the batch into chunks on several streams)
  int H, W, HW, HWp, A, T, R, TW, OS, AS, SA, SAW, SG, SGW, CW, maxobj, NOFF, B, ND, NTERR;
  int NPROXY;  // territory proxy objects behind the object pool: one per territory and agent
  uint32_t objs_stride;  // words of object records per env = (maxobj + NPROXY) * OS
  uint32_t* fast_blk;    // [N][fast_stride] packed hot state of k_step_fast (layout below), or null
  int fast_stride;       // words per env
  // static layer of k_step_fast<G, true> (mg_fast.cu): every non-agent object of a handler-free game is immutable
  const uint32_t* fast_static;  // [N][fast_sstride] header (MGFS_*) + occupancy bitmap of the padded grid, bit c + PAD of row r + PAD
  const uint32_t* fast_slist;   // [N][fast_ns_cap] r << 16 | c of static object i
  const uint16_t* fast_sslot;   // [N][fast_ns_cap] its object slot (pack / unpack only)
  uint32_t* fast_svis;          // [N][fast_ns_cap] GridObject::visited of static object i
  const uint32_t* fast_less;    // [257][8] per packed window offset: the 256-bit set of offsets that come earlier in the
                                // reference's Manhattan order; row 256 = every offset inside the observation shape
  int fast_sstride, fast_ns_cap, fast_sbw;  // words per env, static objects per env (capacity), bitmap words per row
  const uint32_t* rank_lut;  // [256] packed window offset (dr + rr) << 4 | (dc + cr) -> rank << 24 | offset << 16, where rank is
                             // the position in Manhattan order; 0xFFFFFF00 outside the shape
  int plain;    // 1: the program has no handlers / rewards / world systems (k_step<PLAIN> applies)
  int stage_grid;  // 0 (default): cells are read in place; 1: each warp stages its env's grid in shared memory (A/B switch)
  int PAD, WP;  // the grid is stored with a PAD-wide empty frame (row pitch WP) so observation windows need no bounds tests
  // persistent state
  const int16_t* init_cells;  // [N][HW] template per cell
  const float* init_gstats;   // [N][SG]
  const uint32_t* seeds;      // [N]
  uint16_t* cells;            // [N][HWp] object slot per cell of the padded (H+2PAD) x WP grid (0 = empty)
  uint32_t* objs;             // [N][maxobj][OS]
  uint32_t* agents;           // [N][A][AS]
  float* astats;              // [N][A][SA]
  uint32_t* atouched;         // [N][A][SAW]
  float* gstats;              // [N][SG]
  uint32_t* gtouched;         // [N][SGW]
  uint32_t* cover;            // [N][A][CW]
  uint32_t* rng;              // [N][624]
  uint32_t* rng_seeded;       // [N][626] freshly seeded state of each env + the seed it came from + valid flag (k_reset)
  int32_t* env;               // [N][MGEV_WORDS]
  uint8_t* success;           // [N][A]
  // hand-over from k_world / k_init_buffers to k_observe and k_finish (mg_kernels.cu)
  uint4* obs_in;              // [N][A] executed action, start-of-tick location, location, object slot
  int32_t* tok_attempted;     // [N][A] tokens each agent's row attempted (k_finish adds the env's token stats in order)
  uint16_t* obsval;           // [N][A][OVW] word 0 = count, then (feature | value << 8) tokens of the configured global values
  int OVW;                    // 0 when the program has no global observation values
  unsigned long long* claims; // [N][maxobj] step << 32 | ~agent of the lowest agent that saw the object in the latest tick
  uint32_t* visited;          // [N][maxobj] GridObject::visited (the step an object was last observed); MGO_VISITED in the record is unused
  const uint32_t* obs_offs;   // [NOFF] packed window offsets in Manhattan order (mg_capi.cu), read-only
  const struct MgFastHdr* hdr_host;  // HOST copy of the program header, passed to kernels by value
  const float* logtab;        // logf(k + 1), k in [0, 65536), from the host libm
  // world systems (queries / AOE / territory / tags); sizes are 0 when the program does not use them
  uint16_t* arena;            // [N][ARENA] scratch for query result lists (stack discipline)
  uint32_t* aoe_src;          // [N][AOECAP][AOEW]: obj slot, cfg, registration loc, alive, inside bits per agent
  int32_t* aoe_pending;       // [N][PENDCAP][2]: (obj slot, cfg) registrations deferred to the end of the AOE phase
  uint32_t* terr_src;         // [N][TERRCAP][4]: obj slot, territory, strength, decay
  int16_t* inside_tag;        // [N][A][NTERR]: winning tag the agent stood in last tick, or -1
  uint32_t* dyn_stamp;        // [N][maxobj][NDYN]: insertion stamp of run-time-addable tags (tag-index order)
  uint16_t* tag_lists;        // [N][NTAGS][MG_TAG_LIST_CAP]: the reference's TagIndex, insertion-ordered slots per tag
  uint32_t* terr_tab;         // [N][TERRCAP][4]: per-tick table of scoring territory sources
  uint8_t* owner_map;         // [N][NTERR][HW]: winning prefix index + 1 per cell (mg_world.cuh: terr_refresh)
  int32_t* tag_state;         // [N][NTAGS]: member count, or -1 once a tag outgrew its list (then queries scan)
  int ARENA, AOECAP, AOEW, PENDCAP, TERRCAP, NDYN, NTAGS;
  // caller-owned buffers (aliased like the reference's numpy arrays)
  uint8_t* obs;               // [N][A][T][3]
  uint8_t* terminals;         // [N][A]
  uint8_t* truncations;       // [N][A]
  float* rewards;             // [N][A]
  const int32_t* actions;     // [N][A]
  const int32_t* vibe_actions;// [N][A]
};


// shared-memory layout of k_step_fast (mg_fast.cu), computed on the host
struct MgFastLayout {
  int G;           // lanes per environment: 8, 16 or 32
  int warps;       // warps per CTA: 1 or 4
  int statics;     // 1: the static-layer variant (walls and other immutable objects outside the object lanes)
  int rank_off, cta_bytes;
  int tok_off, tok_stride, oloc_off, key_off, group_bytes;
  int sb_off, sb_words, rc_off, wm_off, dl_off, ag_off;  // static variant: bitmap block, row / column observer masks,
                                                         // window sets [G][9], dynamic objects per observer [G][G], per-agent words
  int stage_tokens, wl_cap;  // static variant: tokens of a row that are staged; entries of the work list
  size_t smem_bytes;
};

// per-env header of the static block (MgDev::fast_static); the bitmap follows at word MGFS_HDR_WORDS
enum { MGFS_COUNT = 0, MGFS_NTOK, MGFS_TOKENS /* 4 words: up to eight (feature | value << 8) tokens */, MGFS_HDR_WORDS = 8 };
#define MGFS_MAX_TOKENS 8

// ---- packed hot state of k_step_fast: one contiguous block per env, G = lanes per env -------------------
// words [0, 8)                      : header (MGFB_*)
// words [8 + 8a, 8 + 8a + 8)        : agent a   = {obj slot, spawn, prev_location, steps_without_motion},
//                                                 {max_dist, unique cells, touched bits of stat ids < 16, 0}
// words [8 + 8G + 8j, ... + 8)      : object j+1 = {loc, visited, meta, (agent + 1) | ntok << 8 | dirty << 16},
//                                                 {first eight cached tokens}
// words [8 + 16G + id * G + a]      : agent stat `id` (< 16) of agent a, stat-major so a tick touches whole rows
// The block is the truth between ticks; k_fast_pack / k_fast_unpack move it from / to the generic arrays when an
// entry point other than mg_step needs them (mg_capi.cu).
enum { MGFB_STEP = 0, MGFB_RNG_IDX, MGFB_NOBJ, MGFB_TOKENS_WRITTEN, MGFB_TOKENS_FREE, MGFB_GTOUCHED, MGFB_HDR_WORDS = 8 };
#define MGFB_STATS 16
#define MGFB_AGENT(G, a) (MGFB_HDR_WORDS + 8 * (a))
#define MGFB_OBJECT(G, j) (MGFB_HDR_WORDS + 8 * (G) + 8 * (j))
#define MGFB_STAT(G, id, a) (MGFB_HDR_WORDS + 16 * (G) + (id) * (G) + (a))
#define MGFB_WORDS(G) ((MGFB_HDR_WORDS + 16 * (G) + MGFB_STATS * (G) + 31) & ~31)
#define MGFB_DIRTY (1u << 16)
