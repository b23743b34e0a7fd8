// mg_capi.cu -- the C ABI (include/mettagrid_b200.h): host-side handle, memory, launches.
#include <cuda_runtime.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/mettagrid_b200.h"
#include "mg_state.h"

size_t mg_smem_bytes(const MgDev& d);
cudaError_t mg_configure_kernels(const MgDev& d);
cudaError_t mg_launch_reset(const MgDev& d, const uint8_t* mask, cudaStream_t st);
cudaError_t mg_launch_init_buffers(const MgDev& d, const uint8_t* mask, cudaStream_t st);
cudaError_t mg_launch_step(const MgDev& d, cudaStream_t st);
cudaError_t mg_launch_clear_outputs(const MgDev& d, const uint8_t* mask, cudaStream_t st);
MgFastLayout mg_fast_layout(const MgDev& d, int G, int tok_cap, int statics);
cudaError_t mg_fast_configure(const MgFastLayout& L);
cudaError_t mg_launch_step_fast(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, cudaStream_t st);
cudaError_t mg_launch_fast_pack(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, const uint8_t* mask, cudaStream_t st);
cudaError_t mg_launch_fast_restore(const MgDev& d, const MgFastLayout& L, const uint32_t* snapshot, const uint8_t* mask,
                                   cudaStream_t st);
cudaError_t mg_launch_fast_unpack(const MgDev& d, const MgFastLayout& L, const MgFastHdr& H, cudaStream_t st);
cudaError_t mg_launch_obs_to_grid(const uint8_t* obs, float* grid, int rows, int T, int C, int H, int W, const float* scale,
                                  cudaStream_t st);
cudaError_t mg_launch_token_summary(const uint8_t* obs, int rows, int T, int hidden, int num_feat, const float* pos_x,
                                    const float* pos_y, const float* feat, const float* scale, float* out, cudaStream_t st);
cudaError_t mg_launch_vecenv_prepare(const void* act, int is_int64, int ncols, int num_envs, int A, int P, int V,
                                     const int32_t* vibe_ids, int32_t* actions, int32_t* vibe_actions, const uint8_t* term,
                                     const uint8_t* trunc, uint8_t* done, int64_t* steps, int* counters, cudaStream_t st);
cudaError_t mg_launch_vecenv_post(int num_envs, int A, int64_t* steps, int64_t* early, uint8_t* trunc, cudaStream_t st);
cudaError_t mg_launch_gather_info(const MgDev& d, int packed, int G, int n_game, const int32_t* game_ids, int n_agent,
                                  const int32_t* agent_ids, int idw, int idd, int idf, float* game_out, uint8_t* game_present,
                                  float* agent_out, uint8_t* agent_present, cudaStream_t st);
cudaError_t mg_launch_set_inventory(const MgDev& d, int env, int agent, const int32_t* items, const int32_t* amounts, int n,
                                    cudaStream_t st);

struct mg_handle {
  MgDev d;
  int device = 0;
  std::vector<int32_t> program;
  std::vector<void*> allocs;
  size_t bytes = 0;
  std::string err;
  bool buffers_set = false;
  bool fast = false;  // k_step_fast applies (mg_fast.cu): plain program, sparse environments
  int static_template = -1;  // static-layer variant: the one template every non-agent object is made from
  MgFastLayout fl{};
  MgFastHdr fh{};
  // which copy of the hot state is current on a fast handle: the generic arrays, the packed block, or both
  enum { BOTH = 0, GENERIC = 1, PACKED = 2 };
  int newest = GENERIC;
  // Snapshot of the packed block right after a full reset (+ _init_buffers): lets the vectorised-env step rebuild
  // finished environments with a copy.  `pristine` = nothing but a full reset / set_buffers happened since the
  // generic arrays were built, i.e. the next pack holds exactly that state; any edit of what a reset starts from
  // (seeds, maps, inventories) drops the snapshot until the next full reset.
  uint32_t* snapshot = nullptr;
  bool pristine = true, snap_valid = false;
  // device-side staging for mg_step_host
  int32_t* h_act = nullptr;
  int32_t* h_vact = nullptr;
  uint8_t* h_obs = nullptr;
  float* h_rew = nullptr;
  uint8_t* h_term = nullptr;
  uint8_t* h_trunc = nullptr;
  uint32_t* seeds_dev = nullptr;
  bool host_inited = false;     // staging set initialised (only needed when the caller never called mg_set_buffers)
  cudaStream_t last_stream = nullptr;  // the caller stream the most recent asynchronous entry point was given
  bool foreign_pending = true;  // work may be queued on a caller stream that own_stream is not ordered after
  // vectorised-env glue (mg_vecenv.cu)
  int ve_primary = 0, ve_vibes = 0;
  int32_t* ve_vibe_ids = nullptr;
  uint8_t* ve_done = nullptr;
  int64_t* ve_steps = nullptr;
  int64_t* ve_early = nullptr;
  int* ve_counters = nullptr;  // [0] episodes finished, [1] decoder error bits
  // The steady-state vec-env step of a fast handle (decode -> restore finished envs -> tick -> step counters) as ONE
  // graph launch; rebuilt when the action tensor or its layout changes, dropped when the handle's buffers do.
  cudaGraphExec_t ve_exec = nullptr;
  const void* ve_g_actions = nullptr;
  int ve_g_int64 = 0, ve_g_ncols = 0;
  int ve_g_misses = 0;  // rebuilds in a row: a caller that hands over a fresh tensor every step gets plain launches
  int32_t* info_game_ids = nullptr;   // step_info_keys: configured game / agent stat ids on the device
  int32_t* info_agent_ids = nullptr;
  int info_ngame = 0, info_nagent = 0;
  float* grid_scale = nullptr;  // device float[256], per-feature normalisation of the dense grid observations
  int grid_features = 0;
  cudaStream_t own_stream = nullptr;
  // mg_step of a generic handle splits the batch into chunks on these streams so that one chunk's latency-bound
  // k_world overlaps another chunk's issue-bound k_observe (launch_step)
  enum { MAX_CHUNKS = 8 };
  int chunks = 1;
  cudaStream_t chunk_stream[MAX_CHUNKS] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[MAX_CHUNKS] = {};
};

static std::string g_create_error;
#ifndef MG_STACK_BYTES
#define MG_STACK_BYTES 8192
#endif

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                                 \
      return MG_E_CUDA;                                                                            \
    }                                                                                              \
  } while (0)

// make the generic arrays current (before anything but mg_step reads or edits them)
static int sync_generic(mg_handle* h, cudaStream_t st) {
  if (h->fast && h->newest == mg_handle::PACKED) {
    CK(mg_launch_fast_unpack(h->d, h->fl, h->fh, st));
    h->newest = mg_handle::BOTH;
  }
  return MG_OK;
}
static int launch_step(mg_handle* h, const MgDev& dev, cudaStream_t st) {
  if (!h->fast) {
    if (h->chunks <= 1) {
      CK(mg_launch_step(dev, st));
      return MG_OK;
    }
    // k_world -> k_observe -> k_finish per chunk of envs, every chunk on its own stream, forked from and joined back
    // into the caller's stream: envs are independent, and the tick's three kernels stress different parts of an SM
    CK(cudaEventRecord(h->ev_fork, st));
    const int n = dev.num_envs, per = ((n + h->chunks - 1) / h->chunks + 3) & ~3;
    for (int c = 0; c < h->chunks; c++) {
      MgDev part = dev;
      part.env0 = c * per;
      part.num_envs = (c + 1) * per < n ? (c + 1) * per : n;
      if (part.env0 >= part.num_envs) break;
      CK(cudaStreamWaitEvent(h->chunk_stream[c], h->ev_fork, 0));
      CK(mg_launch_step(part, h->chunk_stream[c]));
      CK(cudaEventRecord(h->ev_join[c], h->chunk_stream[c]));
      CK(cudaStreamWaitEvent(st, h->ev_join[c], 0));
    }
    return MG_OK;
  }
  if (h->newest == mg_handle::GENERIC) {
    CK(mg_launch_fast_pack(dev, h->fl, h->fh, nullptr, st));
    if (h->pristine && h->buffers_set && h->snapshot) {
      CK(cudaMemcpyAsync(h->snapshot, dev.fast_blk, (size_t)dev.num_envs * dev.fast_stride * 4, cudaMemcpyDeviceToDevice, st));
      h->snap_valid = true;
    }
  }
  h->pristine = false;
  CK(mg_launch_step_fast(dev, h->fl, h->fh, st));
  h->newest = mg_handle::PACKED;
  return MG_OK;
}

// Wait for the work this handle queued: the caller stream of the latest asynchronous entry point and the handle's
// own stream.  Stream-scoped on purpose -- a getter must not stall unrelated streams of a training process.
static int sync_streams(mg_handle* h) {
  CK(cudaStreamSynchronize(h->last_stream));
  if (h->own_stream) CK(cudaStreamSynchronize(h->own_stream));
  return MG_OK;
}
// all queued work done and the generic arrays current: what the host-side getters read
static int sync_host_view(mg_handle* h) {
  if (int rc = sync_streams(h)) return rc;
  if (int rc = sync_generic(h, h->own_stream)) return rc;
  CK(cudaStreamSynchronize(h->own_stream));
  return MG_OK;
}

template <class T>
static int dev_alloc(mg_handle* h, T** p, size_t count) {
  size_t bytes = (count ? count : 1) * sizeof(T);
  bytes = (bytes + 255) & ~(size_t)255;
  CK(cudaMalloc((void**)p, bytes));
  CK(cudaMemset(*p, 0, bytes));  // legacy stream: callers synchronise the device before launching on own_stream
  h->foreign_pending = true;
  h->allocs.push_back(*p);
  h->bytes += bytes;
  return MG_OK;
}

// engine limits: anything beyond them is refused loudly rather than mis-simulated
// A "plain" program never needs the interpreter: only the two default move handlers, no on_use / on_tick /
// on_after_use / tag handlers, no rewards, obs values, events, AOE, territory or materialized queries.
static int program_is_plain(const int32_t* P) {
  if (P[MGH_NUM_MOVE_HANDLERS] != 2 || P[MGH_NUM_OBS_VALUES] || P[MGH_NUM_EVENTS_SCHED] || P[MGH_NUM_TERRITORIES] ||
      P[MGH_NUM_MQ] || P[MGH_GAME_ON_TICK] >= 0 || P[MGH_FEAT_AOE_MASK])
    return 0;
  const int32_t* ch = P + P[MGS_MOVE_CHAIN];
  if (ch[3] != MGMB_RELOCATE || ch[MG_MOVEH_WORDS + 3] != MGMB_USE_TARGET) return 0;
  const int32_t* T_ = P + P[MGS_TEMPLATES];
  for (int t = 0; t < P[MGH_NUM_TEMPLATES]; t++) {
    const int32_t* tp = T_ + t * MG_TEMPLATE_WORDS;
    if (tp[MGT_ON_USE] >= 0 || tp[MGT_ON_TICK] >= 0 || tp[MGT_ON_AFTER_USE] >= 0 || tp[MGT_REWARDS_N] || tp[MGT_AOES_N] ||
        tp[MGT_TERR_N] || tp[MGT_TAG_REMOVE_N])
      return 0;
  }
  return 1;
}

// k_step_fast (mg_fast.cu) needs: a plain program, agents and objects that fit one warp, primary-stream actions
// that are noop / move and vibe-stream actions that are change_vibe, well-known stat ids below 64, a window of at
// most 15 x 15 cells.  Returns the lanes per environment (8, 16, 32) or 0.  `max_objects_per_env` counts what has to
// sit in the object lanes: every object, or -- static-layer variant -- the agents only.
static int fast_group_size(const int32_t* P, const MgDev& d, int max_objects_per_env) {
  if (!d.plain || getenv("METTAGRID_B200_NO_FAST")) return 0;
  if ((d.AS & 3) || (d.OS & 3)) return 0;  // vector loads of the records
  if (d.A > 32 || max_objects_per_env > 32 || P[MGH_OBS_H] > 15 || P[MGH_OBS_W] > 15 || P[MGH_TOK_CAP] > 126) return 0;
  for (int k = MGH_ST_ACTION_FAILED; k <= MGH_ST_VIBE_FAILED; k++)
    if (P[k] < 0 || P[k] >= MGFB_STATS) return 0;
  const int32_t* acts = P + P[MGS_ACTIONS];
  for (int i = 0; i < P[MGH_NUM_ACTIONS]; i++) {
    const int kind = acts[i * MG_ACTION_WORDS], is_vibe = acts[i * MG_ACTION_WORDS + 3];
    if (is_vibe ? kind != MGA_CHANGE_VIBE : (kind != MGA_NOOP && kind != MGA_MOVE)) return 0;
  }
  const int need = d.A > max_objects_per_env ? d.A : max_objects_per_env;
  return need <= 8 ? 8 : need <= 16 ? 16 : 32;
}

// The static-layer variant applies when every non-agent object of every map comes from ONE template whose token list
// is short: then walls (or blocks) are a bitmap plus one shared token list.  Returns that template, -1 if the maps
// hold no non-agent object or none qualifies (-2).
static int static_layer_template(const int32_t* P, const MgDev& d, const int16_t* init_cells, size_t n_cells) {
  if (getenv("METTAGRID_B200_NO_FAST_STATIC")) return -2;
  const int32_t* TP = P + P[MGS_TEMPLATES];
  int tmpl = -1;
  for (size_t i = 0; i < n_cells; i++) {
    const int t = init_cells[i];
    if (t < 0 || TP[t * MG_TEMPLATE_WORDS + MGT_KIND] == 1) continue;
    if (tmpl >= 0 && t != tmpl) return -2;
    tmpl = t;
  }
  if (tmpl < 0) return -1;
  const int32_t* tp = TP + tmpl * MG_TEMPLATE_WORDS;
  int tokens = tp[MGT_VIBE] ? 1 : 0;
  for (int k = 0; k < d.TW; k++) tokens += __builtin_popcount((uint32_t)P[P[MGS_POOL] + tp[MGT_TAGS] + k]);
  tokens += tp[MGT_INIT_INV_N] * d.ND;  // upper bound: every digit of every held resource
  return tokens <= MGFS_MAX_TOKENS ? tmpl : -2;
}

static const char* unsupported_reason(const int32_t* P) {
  const int32_t* T_ = P + P[MGS_TERRITORIES];
  for (int i = 0; i < P[MGH_NUM_TERRITORIES]; i++)
    if (T_[i * MG_TERR_WORDS + 1] > 16) return "more than 16 tags under one territory prefix";
  return nullptr;
}

extern "C" {

int mg_create(const int32_t* program, size_t nwords, int num_envs, const int16_t* init_cells, const float* init_gstats,
              const uint32_t* seeds, int device, mg_handle** out) {
  if (!out) return MG_E_INVALID;
  *out = nullptr;
  if (!program || nwords < (size_t)MGH_HEADER_WORDS || (uint32_t)program[MGH_MAGIC] != MG_MAGIC ||
      program[MGH_VERSION] != MG_VERSION || (size_t)program[MGH_TOTAL_WORDS] != nwords) {
    g_create_error = "mg_create: not a compiled game program of this version";
    return MG_E_INVALID;
  }
  if (num_envs <= 0 || !init_cells || !seeds) {
    g_create_error = "mg_create: num_envs, init_cells and seeds are required";
    return MG_E_INVALID;
  }
  if (const char* why = unsupported_reason(program)) {
    g_create_error = std::string("mg_create: program uses ") + why + ", which this build does not run on the GPU";
    return MG_E_UNSUPPORTED;
  }
  mg_handle* h = new mg_handle();
  auto fail = [&](int code) {
    g_create_error = h->err;
    mg_destroy(h);
    return code;
  };
  h->device = device;
  h->program.assign(program, program + nwords);
  const int32_t* P = program;
  MgDev& d = h->d;
  memset(&d, 0, sizeof d);
  d.num_envs = num_envs;
  d.H = P[MGH_H], d.W = P[MGH_W], d.HW = d.H * d.W;
  d.PAD = (P[MGH_OBS_H] > P[MGH_OBS_W] ? P[MGH_OBS_H] : P[MGH_OBS_W]) / 2;
  d.WP = d.W + 2 * d.PAD;
  d.HWp = ((d.H + 2 * d.PAD) * d.WP + 7) & ~7;
  d.A = P[MGH_NUM_AGENTS], d.T = P[MGH_NUM_TOKENS], d.R = P[MGH_NUM_RESOURCES], d.TW = P[MGH_TAG_WORDS];
  d.OS = P[MGH_OBJ_STRIDE], d.AS = P[MGH_AGENT_STRIDE], d.SA = P[MGH_NUM_AGENT_STATS], d.SAW = (d.SA + 31) / 32;
  d.SG = P[MGH_NUM_GAME_STATS], d.SGW = (d.SG + 31) / 32, d.CW = P[MGH_COVER_WORDS], d.maxobj = P[MGH_MAX_OBJECTS];
  d.NOFF = P[MGH_NUM_OFFSETS], d.B = P[MGH_TOKEN_BASE], d.ND = P[MGH_INV_DIGITS], d.NTERR = P[MGH_NUM_TERRITORIES];
  d.NPROXY = d.NTERR * d.A;
  d.objs_stride = (uint32_t)(d.maxobj + d.NPROXY) * (uint32_t)d.OS;
  d.plain = program_is_plain(P);
  if (d.R > 13 || d.maxobj > 65535 || d.A > 4096) {
    h->err = "mg_create: program exceeds engine limits (R<=13, objects<=65535, agents<=4096)";
    return fail(MG_E_INVALID);
  }
  int rc;
  int max_objs = 0;  // most objects any env starts with
  int max_static = 0;  // ... most non-agent objects
#define TRY(x) \
  if ((rc = (x)) != MG_OK) return fail(rc)
  if (cudaSetDevice(device) != cudaSuccess) {
    h->err = "mg_create: cudaSetDevice failed (no CUDA device?)";
    return fail(MG_E_CUDA);
  }
  const size_t N = (size_t)num_envs;
  int32_t* Pd;
  int16_t* ic;
  float* ig = nullptr;
  float* lt;
  TRY(dev_alloc(h, &Pd, nwords));
  TRY(dev_alloc(h, &ic, N * d.HW));
  if (init_gstats) TRY(dev_alloc(h, &ig, N * d.SG));
  TRY(dev_alloc(h, &h->seeds_dev, N));
  TRY(dev_alloc(h, &d.cells, N * d.HWp));
  TRY(dev_alloc(h, &d.objs, N * (d.maxobj + d.NPROXY) * d.OS));  // records hold >= 4 token words (compiler.py: obj_stride)
  {
    // capacities of the world-system tables, from the program and the initial maps
    const int32_t* TP = P + P[MGS_TEMPLATES];
    size_t max_aoe = 0, max_terr = 0;
    for (size_t e = 0; e < N; e++) {
      size_t na = 0, nt = 0;
      int no = 0, ns = 0;
      const int16_t* cells = init_cells + e * d.HW;
      for (int i = 0; i < d.HW; i++)
        if (cells[i] >= 0) {
          if (cells[i] >= P[MGH_NUM_TEMPLATES]) {
            h->err = "mg_create: init_cells holds a template index outside the program";
            return fail(MG_E_INVALID);
          }
          no++;
          ns += TP[cells[i] * MG_TEMPLATE_WORDS + MGT_KIND] != 1;
          na += TP[cells[i] * MG_TEMPLATE_WORDS + MGT_AOES_N];
          nt += TP[cells[i] * MG_TEMPLATE_WORDS + MGT_TERR_N];
        }
      max_objs = no > max_objs ? no : max_objs;
      max_static = ns > max_static ? ns : max_static;
      max_aoe = na > max_aoe ? na : max_aoe;
      max_terr = nt > max_terr ? nt : max_terr;
    }
    const int spawn_aoes = P[MGH_SPAWN_AOES];
    d.AOECAP = (int)max_aoe + spawn_aoes * 1024;
    d.AOEW = 4 + (d.A + 31) / 32;
    d.PENDCAP = spawn_aoes > 0 ? 256 : 0;
    d.TERRCAP = (int)max_terr;
    d.NDYN = P[MGH_NUM_DYN_TAGS];
    const int nq = (P[MGS_TEMPLATES] - P[MGS_QUERIES]) / MG_QUERY_WORDS;
    d.ARENA = (nq > 0 || P[MGH_NUM_MQ] > 0) ? 4 * d.maxobj + 64 : 0;
    d.NTAGS = d.ARENA > 0 ? P[MGH_NUM_TAGS] : 0;
  }
  TRY(dev_alloc(h, &d.tag_lists, N * d.NTAGS * MG_TAG_LIST_CAP));
  TRY(dev_alloc(h, &d.tag_state, N * d.NTAGS));
  TRY(dev_alloc(h, &d.arena, N * d.ARENA));
  TRY(dev_alloc(h, &d.aoe_src, N * d.AOECAP * d.AOEW));
  TRY(dev_alloc(h, &d.aoe_pending, N * d.PENDCAP * 2));
  TRY(dev_alloc(h, &d.terr_src, N * d.TERRCAP * 4));
  TRY(dev_alloc(h, &d.terr_tab, N * d.TERRCAP * 4));
  TRY(dev_alloc(h, &d.inside_tag, N * d.A * d.NTERR));
  TRY(dev_alloc(h, &d.owner_map, N * d.NTERR * d.HW));
  TRY(dev_alloc(h, &d.dyn_stamp, N * d.maxobj * d.NDYN));
  TRY(dev_alloc(h, &d.agents, N * d.A * d.AS));
  TRY(dev_alloc(h, &d.astats, N * d.A * d.SA));
  TRY(dev_alloc(h, &d.atouched, N * d.A * d.SAW));
  TRY(dev_alloc(h, &d.gstats, N * d.SG));
  TRY(dev_alloc(h, &d.gtouched, N * d.SGW));
  TRY(dev_alloc(h, &d.cover, N * d.A * d.CW));
  TRY(dev_alloc(h, &d.rng, N * MG_RNG_WORDS));
  TRY(dev_alloc(h, &d.rng_seeded, N * (MG_RNG_WORDS + 2)));
  TRY(dev_alloc(h, &d.env, N * MGEV_WORDS));
  TRY(dev_alloc(h, &d.success, N * d.A));
  TRY(dev_alloc(h, &d.obs_in, N * d.A));
  TRY(dev_alloc(h, &d.claims, N * d.maxobj));
  TRY(dev_alloc(h, &d.visited, N * d.maxobj));
  memcpy(h->fh.v, P, sizeof h->fh.v);
  d.hdr_host = &h->fh;
  {
    // packed window offsets in the reference's Manhattan order (core/observation_shape.cpp:19-66) -- bits 0-7:
    // dr + 8 | (dc + 8) << 4; bits 8-15: packed location (packed_coordinate.hpp:50-56); bits 16-31: cell delta in the
    // padded grid
    const int32_t* po = P + P[MGS_OFFSETS];
    const int rr = P[MGH_OBS_H] >> 1, cr = P[MGH_OBS_W] >> 1;
    std::vector<uint32_t> tab((size_t)d.NOFF);
    for (int i = 0; i < d.NOFF; i++) {
      const int dr = po[2 * i], dc = po[2 * i + 1];
      const uint32_t loc = (uint32_t)(((dr + rr) << 4) | ((dc + cr) & 15));
      tab[i] = (uint32_t)(dr + 8) | ((uint32_t)(dc + 8) << 4) | (loc << 8) | ((uint32_t)((dr * d.WP + dc) & 0xffff) << 16);
    }
    uint32_t* dev;
    TRY(dev_alloc(h, &dev, tab.size()));
    if (cudaMemcpy(dev, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
      h->err = "mg_create: upload failed";
      return fail(MG_E_CUDA);
    }
    d.obs_offs = dev;
  }
  TRY(dev_alloc(h, &d.tok_attempted, N * d.A));
  {
    // tokens a configured global observation value can take: the base-B digits of a 32-bit value (encoding_utils.hpp:16-35)
    int digits = 1;
    for (unsigned long long v = (unsigned long long)(d.B > 1 ? d.B : 2); v <= 0xffffffffull; v *= (unsigned long long)(d.B > 1 ? d.B : 2)) digits++;
    d.OVW = P[MGH_NUM_OBS_VALUES] > 0 ? 1 + P[MGH_NUM_OBS_VALUES] * digits : 0;
    TRY(dev_alloc(h, &d.obsval, N * d.A * d.OVW));
  }
  TRY(dev_alloc(h, &lt, 65536));
  {
    // logf through the HOST libm, the same function the reference calls (core/game_value.cpp:91)
    std::vector<float> tab(65536);
    for (int k = 0; k < 65536; k++) tab[k] = logf((float)k + 1.0f);
    if (cudaMemcpy(lt, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      h->err = "mg_create: upload failed";
      return fail(MG_E_CUDA);
    }
  }
  if (cudaMemcpy(Pd, program, nwords * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(ic, init_cells, N * d.HW * 2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h->seeds_dev, seeds, N * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      (ig && cudaMemcpy(ig, init_gstats, N * d.SG * 4, cudaMemcpyHostToDevice) != cudaSuccess)) {
    h->err = "mg_create: upload failed";
    return fail(MG_E_CUDA);
  }
  d.P = Pd, d.init_cells = ic, d.init_gstats = ig, d.seeds = h->seeds_dev, d.logtab = lt;
  {
    // window-rank table: packed offset (dr + rr) << 4 | (dc + cr) -> position in the reference's Manhattan order
    // (core/observation_shape.cpp:19-66); 0xFF = outside the shape
    std::vector<uint32_t> lut(256, 0xFFFFFF00u);
    uint32_t* lut_dev;
    const int32_t* offs = P + P[MGS_OFFSETS];
    const int rr = P[MGH_OBS_H] >> 1, cr = P[MGH_OBS_W] >> 1;
    if (P[MGH_OBS_H] <= 15 && P[MGH_OBS_W] <= 15 && d.NOFF < 255)
      for (int i = 0; i < d.NOFF; i++) {
        const uint32_t loc = (uint32_t)(((offs[2 * i] + rr) << 4) | ((offs[2 * i + 1] + cr) & 15));
        lut[loc] = ((uint32_t)i << 24) | (loc << 16);
      }
    TRY(dev_alloc(h, &lut_dev, 256));
    if (cudaMemcpy(lut_dev, lut.data(), 1024, cudaMemcpyHostToDevice) != cudaSuccess) {
      h->err = "mg_create: upload failed";
      return fail(MG_E_CUDA);
    }
    d.rank_lut = lut_dev;
  }
  {
    // The env's grid is read in place (L1 / L2).  Staging it in shared memory every tick -- 3 to 12 KB per warp --
    // measured slower on every workload (profiles/README.md): C3 1259 -> 873 us, C4 6240 -> 3008 us, plain C2
    // 97.6 -> 95.7 us.  METTAGRID_B200_STAGE_GRID=1 brings the staging back for A/B runs.
    const char* f = getenv("METTAGRID_B200_STAGE_GRID");
    d.stage_grid = f && f[0] == '1';
  }
  {
    // The handler interpreter (mg_device.cuh) is recursive -- one copy of each function instead of one per nesting
    // level keeps the hot code inside the instruction cache -- so its stack cannot be sized statically: frames are at
    // most ~0.5 KB and the nesting is capped at MG_DEPTH = 8 levels of handlers / filters / values / queries.
    size_t cur = 0;
    if (cudaDeviceGetLimit(&cur, cudaLimitStackSize) == cudaSuccess && cur < MG_STACK_BYTES)
      cudaDeviceSetLimit(cudaLimitStackSize, MG_STACK_BYTES);
  }
  cudaError_t e = mg_configure_kernels(d);
  if (e != cudaSuccess) {
    h->err = std::string("mg_create: kernel configuration failed: ") + cudaGetErrorString(e) +
             " (shared memory per CTA = " + std::to_string(mg_smem_bytes(d)) + " bytes)";
    return fail(MG_E_CUDA);
  }
  // Sparse maps put every object into a lane; maps with more objects than lanes keep the agents there and carry the
  // rest as the static layer, when one template makes all of it (walls).
  int statics = 0;
  if (d.plain && (max_objs > 32 || getenv("METTAGRID_B200_FORCE_FAST_STATIC"))) {  // (the switch: A/B runs on sparse maps)
    h->static_template = static_layer_template(P, d, init_cells, N * d.HW);
    statics = h->static_template >= 0 || (h->static_template == -1 && getenv("METTAGRID_B200_FORCE_FAST_STATIC"));
  }
  if (const int G = fast_group_size(P, d, statics ? d.A : max_objs)) {
    if (statics) {
      d.fast_sbw = (d.W + 2 * d.PAD + 31) / 32 + 1;  // + 1: a window row is a funnel shift over two adjacent words
      d.fast_sstride = (MGFS_HDR_WORDS + (d.H + 2 * d.PAD) * d.fast_sbw + 3) & ~3;
      d.fast_ns_cap = (max_static + 31) & ~31;
    }
    h->fl = mg_fast_layout(d, G, P[MGH_TOK_CAP], statics);
    memcpy(h->fh.v, P, sizeof h->fh.v);
    if (h->fl.smem_bytes <= 200 * 1024 && mg_fast_configure(h->fl) == cudaSuccess) {
      d.fast_stride = MGFB_WORDS(G);
      TRY(dev_alloc(h, &d.fast_blk, N * d.fast_stride));
      TRY(dev_alloc(h, &h->snapshot, N * d.fast_stride));
      if (statics) {
        uint32_t *sb, *sl, *less;
        uint16_t* ss;
        TRY(dev_alloc(h, &sb, N * d.fast_sstride));
        TRY(dev_alloc(h, &sl, N * d.fast_ns_cap));
        TRY(dev_alloc(h, &ss, N * d.fast_ns_cap));
        TRY(dev_alloc(h, &d.fast_svis, N * d.fast_ns_cap));
        TRY(dev_alloc(h, &less, 257 * 8));
        // per packed window offset: the offsets that come earlier in the reference's Manhattan order
        // (core/observation_shape.cpp:52-66); row 256: every offset of the shape
        std::vector<uint32_t> tab(257 * 8, 0u);
        const int32_t* offs = P + P[MGS_OFFSETS];
        const int rr = P[MGH_OBS_H] >> 1, cr = P[MGH_OBS_W] >> 1;
        uint32_t acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < d.NOFF; i++) {
          const uint32_t loc = (uint32_t)(((offs[2 * i] + rr) << 4) | ((offs[2 * i + 1] + cr) & 15));
          memcpy(&tab[loc * 8], acc, sizeof acc);
          acc[loc >> 5] |= 1u << (loc & 31);
        }
        memcpy(&tab[256 * 8], acc, sizeof acc);
        if (cudaMemcpy(less, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
          h->err = "mg_create: upload failed";
          return fail(MG_E_CUDA);
        }
        d.fast_static = sb, d.fast_slist = sl, d.fast_sslot = ss, d.fast_less = less;
      }
      h->fast = true;
    } else {
      cudaGetLastError();  // too large for shared memory: the generic kernel runs instead
      d.fast_sbw = d.fast_sstride = d.fast_ns_cap = 0;
      h->fl.statics = 0;
    }
  }
  if (cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
    h->err = "mg_create: stream creation failed";
    return fail(MG_E_CUDA);
  }
  if (!h->fast) {
    // chunks per tick: METTAGRID_B200_CHUNKS, default 2 once the batch is large enough to fill the GPU several times
    // (C4 8192 envs: 1292 -> 1228 us per tick, C3 16 384 envs: 2126 -> 2097 us; profiles/README.md)
    const char* f = getenv("METTAGRID_B200_CHUNKS");
    int want = f ? atoi(f) : ((long long)num_envs * d.A >= 131072 ? 2 : 1);  // measured: 2 / 4 / 8 chunks within 1 % of each other
    if (want > mg_handle::MAX_CHUNKS) want = mg_handle::MAX_CHUNKS;
    if (want > 1) {
      bool ok = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) == cudaSuccess;
      for (int c = 0; c < want && ok; c++)
        ok = cudaStreamCreateWithFlags(&h->chunk_stream[c], cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreateWithFlags(&h->ev_join[c], cudaEventDisableTiming) == cudaSuccess;
      if (!ok) {
        h->err = "mg_create: stream creation failed";
        return fail(MG_E_CUDA);
      }
      h->chunks = want;
    }
  }
  e = cudaDeviceSynchronize();  // the zero fills above ran on the legacy stream, own_stream does not wait for it
  if (e == cudaSuccess) e = mg_launch_reset(d, nullptr, h->own_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
  if (e != cudaSuccess) {
    h->err = std::string("mg_create: reset kernel failed: ") + cudaGetErrorString(e);
    return fail(MG_E_CUDA);
  }
#undef TRY
  *out = h;
  return MG_OK;
}

static void drop_vecenv_graph(mg_handle* h) {
  if (h->ve_exec) cudaGraphExecDestroy(h->ve_exec);
  h->ve_exec = nullptr;
}

void mg_destroy(mg_handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  drop_vecenv_graph(h);
  for (void* p : h->allocs) cudaFree(p);
  if (h->own_stream) cudaStreamDestroy(h->own_stream);
  for (int c = 0; c < mg_handle::MAX_CHUNKS; c++) {
    if (h->chunk_stream[c]) cudaStreamDestroy(h->chunk_stream[c]);
    if (h->ev_join[c]) cudaEventDestroy(h->ev_join[c]);
  }
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  delete h;
}

const char* mg_last_error(const mg_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int mg_set_buffers(mg_handle* h, void* observations, void* terminals, void* truncations, void* rewards,
                   const void* actions, const void* vibe_actions, void* stream) {
  if (!h) return MG_E_INVALID;
  if (!observations || !terminals || !truncations || !rewards || !actions || !vibe_actions) {
    h->err = "mg_set_buffers: all six buffers are required";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  drop_vecenv_graph(h);
  MgDev& d = h->d;
  d.obs = (uint8_t*)observations, d.terminals = (uint8_t*)terminals, d.truncations = (uint8_t*)truncations;
  d.rewards = (float*)rewards, d.actions = (const int32_t*)actions, d.vibe_actions = (const int32_t*)vibe_actions;
  h->buffers_set = true;
  h->foreign_pending = true, h->last_stream = (cudaStream_t)stream;
  if (int rc = sync_generic(h, (cudaStream_t)stream)) return rc;
  CK(mg_launch_init_buffers(d, nullptr, (cudaStream_t)stream));
  if (h->fast) h->newest = mg_handle::GENERIC;
  return MG_OK;
}

int mg_step(mg_handle* h, void* stream) {
  if (!h) return MG_E_INVALID;
  if (!h->buffers_set) {
    h->err = "mg_step: call mg_set_buffers first";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  h->foreign_pending = true, h->last_stream = (cudaStream_t)stream;
  return launch_step(h, h->d, (cudaStream_t)stream);
}

// the handle's own staging set as a step descriptor: used by mg_step_host only, h->d keeps the caller's buffers
static MgDev host_dev(const mg_handle* h) {
  MgDev run = h->d;
  run.obs = h->h_obs, run.terminals = h->h_term, run.truncations = h->h_trunc, run.rewards = h->h_rew;
  run.actions = h->h_act, run.vibe_actions = h->h_vact;
  return run;
}

int mg_step_host(mg_handle* h, const int32_t* actions, const int32_t* vibe_actions, uint8_t* observations,
                 float* rewards, uint8_t* terminals, uint8_t* truncations) {
  if (!h || !actions) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  MgDev& d = h->d;
  const size_t NA = (size_t)d.num_envs * d.A;
  if (!h->h_act) {  // first use: device-side buffers owned by the handle
    int rc;
    if ((rc = dev_alloc(h, &h->h_act, NA)) || (rc = dev_alloc(h, &h->h_vact, NA)) ||
        (rc = dev_alloc(h, &h->h_obs, NA * d.T * 3)) || (rc = dev_alloc(h, &h->h_rew, NA)) ||
        (rc = dev_alloc(h, &h->h_term, NA)) || (rc = dev_alloc(h, &h->h_trunc, NA)))
      return rc;
  }
  cudaStream_t st = h->own_stream;
  if (h->foreign_pending) {  // own_stream is non-blocking: order it after whatever the caller queued elsewhere
    CK(cudaDeviceSynchronize());
    h->foreign_pending = false;
  }
  MgDev run = host_dev(h);
  // Observation rows are the bulk of the traffic (3T bytes per agent, 20 MB per step at C2) and the step kernels only
  // WRITE them, in whole 16-byte vectors.  When the caller's buffer is pinned host memory (cudaHostAlloc /
  // cudaHostRegister, what PufferLib-style numpy buffers are once pinned), the kernels store the rows straight into it
  // over PCIe: the transfer overlaps the tick instead of following it and the staging copy disappears.  Pageable memory
  // takes the staged path.  Measured on this pool (C2, 4096 envs, 20 MB of rows per step): 1.35e8 agent-steps/s direct vs
  // 1.41e8 through the copy engine -- SM stores over PCIe do not beat the DMA, both sit at the link's ~43 GB/s -- so the
  // direct path is opt-in: METTAGRID_B200_DIRECT_OBS=1.
  bool direct_obs = false;
  if (observations && getenv("METTAGRID_B200_DIRECT_OBS")) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, observations) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
      run.obs = (uint8_t*)at.devicePointer;
      direct_obs = true;
    } else {
      cudaGetLastError();
    }
  }
  if (!h->buffers_set && !h->host_inited) {
    // a caller that only ever uses host buffers: _init_buffers (coverage, flags) against the staging set.  A handle
    // whose buffers were set keeps its episode state untouched -- mg_set_buffers already ran _init_buffers.
    if (int rc = sync_generic(h, st)) return rc;
    CK(mg_launch_init_buffers(host_dev(h), nullptr, st));
    if (h->fast) h->newest = mg_handle::GENERIC;
    h->host_inited = true;
  }
  CK(cudaMemcpyAsync(h->h_act, actions, NA * 4, cudaMemcpyHostToDevice, st));
  if (vibe_actions)
    CK(cudaMemcpyAsync(h->h_vact, vibe_actions, NA * 4, cudaMemcpyHostToDevice, st));
  else
    CK(cudaMemsetAsync(h->h_vact, 0, NA * 4, st));
  {
    const int rc = launch_step(h, run, st);
    if (rc) return rc;
  }
  if (observations && !direct_obs) CK(cudaMemcpyAsync(observations, h->h_obs, NA * d.T * 3, cudaMemcpyDeviceToHost, st));
  if (rewards) CK(cudaMemcpyAsync(rewards, h->h_rew, NA * 4, cudaMemcpyDeviceToHost, st));
  if (terminals) CK(cudaMemcpyAsync(terminals, h->h_term, NA, cudaMemcpyDeviceToHost, st));
  if (truncations) CK(cudaMemcpyAsync(truncations, h->h_trunc, NA, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return MG_OK;
}

int mg_reset(mg_handle* h, const uint8_t* env_mask, const uint32_t* new_seeds, void* stream) {
  if (!h) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  h->foreign_pending = true, h->last_stream = st;
  if (new_seeds) {
    CK(cudaMemcpyAsync(h->seeds_dev, new_seeds, (size_t)h->d.num_envs * 4, cudaMemcpyHostToDevice, st));
    h->snap_valid = false;
    if (env_mask) h->pristine = false;  // the unmasked envs keep generator states seeded from their old seeds
  }
  if (!env_mask) h->pristine = true;  // every env is about to hold its post-reset state again
  CK(mg_launch_reset(h->d, env_mask, st));
  if (h->buffers_set) {
    CK(mg_launch_init_buffers(h->d, env_mask, st));
    if (h->h_act) CK(mg_launch_clear_outputs(host_dev(h), env_mask, st));  // the staging set's flags start over too
  } else if (h->host_inited) {
    CK(mg_launch_init_buffers(host_dev(h), env_mask, st));
  }
  if (h->fast) {
    // a full reset makes the generic arrays the truth; a masked one re-packs just the rebuilt environments so
    // that the others never leave the packed block
    if (!env_mask)
      h->newest = mg_handle::GENERIC;
    else if (h->newest != mg_handle::GENERIC)
      CK(mg_launch_fast_pack(h->d, h->fl, h->fh, env_mask, st));
  }
  return MG_OK;
}

int mg_set_map(mg_handle* h, int env, const int16_t* init_cells, const float* init_gstats) {
  if (!h || !init_cells || env < 0 || env >= h->d.num_envs) return MG_E_INVALID;
  const MgDev& d = h->d;
  const int32_t* P = h->program.data();
  int nobj = 0, nstatic = 0;
  bool foreign = false;  // a non-agent object made from another template than the handle's static layer
  for (int i = 0; i < d.HW; i++) {
    if (init_cells[i] >= P[MGH_NUM_TEMPLATES]) {
      h->err = "mg_set_map: init_cells holds a template index outside the program";
      return MG_E_INVALID;
    }
    nobj += init_cells[i] >= 0;
    if (init_cells[i] >= 0 && P[P[MGS_TEMPLATES] + init_cells[i] * MG_TEMPLATE_WORDS + MGT_KIND] != 1) {
      nstatic++;
      foreign |= init_cells[i] != h->static_template;
    }
  }
  const bool fits = !h->fast || (h->fl.statics ? nstatic <= d.fast_ns_cap && !foreign : nobj <= h->fl.G);
  if (nobj >= d.maxobj || !fits) {
    h->err = "mg_set_map: the map holds more objects (or other static objects) than this handle was created for";
    return MG_E_INVALID;
  }
  if (init_gstats && !d.init_gstats) {
    h->err = "mg_set_map: the handle was created without initial game stats";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  if (int rc = sync_streams(h)) return rc;
  CK(cudaMemcpy((void*)(d.init_cells + (size_t)env * d.HW), init_cells, (size_t)d.HW * 2, cudaMemcpyHostToDevice));
  if (init_gstats)
    CK(cudaMemcpy((void*)(d.init_gstats + (size_t)env * d.SG), init_gstats, (size_t)d.SG * 4, cudaMemcpyHostToDevice));
  h->snap_valid = false, h->pristine = false;
  return MG_OK;
}

int mg_poll_errors(mg_handle* h, int* env, int* code, int* info) {
  if (!h) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  std::vector<int32_t> e((size_t)h->d.num_envs * MGEV_WORDS);
  if (int rc = sync_streams(h)) return rc;
  CK(cudaMemcpy(e.data(), h->d.env, e.size() * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->d.num_envs; i++)
    if (e[(size_t)i * MGEV_WORDS + MGEV_ERROR]) {
      if (env) *env = i;
      if (code) *code = e[(size_t)i * MGEV_WORDS + MGEV_ERROR];
      if (info) *info = e[(size_t)i * MGEV_WORDS + MGEV_ERR_INFO];
      h->err = "environment " + std::to_string(i) + " raised error bits " +
               std::to_string(e[(size_t)i * MGEV_WORDS + MGEV_ERROR]);
      return MG_E_ENV;
    }
  return MG_OK;
}

int mg_get_episode_rewards(mg_handle* h, float* out) {
  if (!h || !out) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  const MgDev& d = h->d;
  std::vector<uint32_t> ag((size_t)d.num_envs * d.A * d.AS);
  if (int rc = sync_streams(h)) return rc;
  CK(cudaMemcpy(ag.data(), d.agents, ag.size() * 4, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < (size_t)d.num_envs * d.A; i++) memcpy(&out[i], &ag[i * d.AS + MGAG_EPISODE_REWARD], 4);
  return MG_OK;
}

int mg_get_action_success(mg_handle* h, uint8_t* out) {
  if (!h || !out) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (int rc = sync_streams(h)) return rc;
  CK(cudaMemcpy(out, h->d.success, (size_t)h->d.num_envs * h->d.A, cudaMemcpyDeviceToHost));
  return MG_OK;
}

int mg_get_current_steps(mg_handle* h, int32_t* out) {
  if (!h || !out) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  std::vector<int32_t> e((size_t)h->d.num_envs * MGEV_WORDS);
  if (int rc = sync_host_view(h)) return rc;
  CK(cudaMemcpy(e.data(), h->d.env, e.size() * 4, cudaMemcpyDeviceToHost));
  for (int i = 0; i < h->d.num_envs; i++) out[i] = e[(size_t)i * MGEV_WORDS + MGEV_STEP];
  return MG_OK;
}

int mg_get_agent_stats(mg_handle* h, int env, float* values, uint8_t* touched) {
  if (!h || env < 0 || env >= h->d.num_envs) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  const MgDev& d = h->d;
  if (int rc = sync_host_view(h)) return rc;
  if (values) CK(cudaMemcpy(values, d.astats + (size_t)env * d.A * d.SA, (size_t)d.A * d.SA * 4, cudaMemcpyDeviceToHost));
  if (touched) {
    std::vector<uint32_t> t((size_t)d.A * d.SAW);
    CK(cudaMemcpy(t.data(), d.atouched + (size_t)env * d.A * d.SAW, t.size() * 4, cudaMemcpyDeviceToHost));
    for (int a = 0; a < d.A; a++)
      for (int i = 0; i < d.SA; i++) touched[a * d.SA + i] = (t[a * d.SAW + (i >> 5)] >> (i & 31)) & 1;
  }
  return MG_OK;
}

int mg_get_game_stats(mg_handle* h, int env, float* values, uint8_t* touched) {
  if (!h || env < 0 || env >= h->d.num_envs) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  const MgDev& d = h->d;
  if (int rc = sync_host_view(h)) return rc;
  if (values) CK(cudaMemcpy(values, d.gstats + (size_t)env * d.SG, (size_t)d.SG * 4, cudaMemcpyDeviceToHost));
  if (touched) {
    std::vector<uint32_t> t(d.SGW);
    CK(cudaMemcpy(t.data(), d.gtouched + (size_t)env * d.SGW, t.size() * 4, cudaMemcpyDeviceToHost));
    for (int i = 0; i < d.SG; i++) touched[i] = (t[i >> 5] >> (i & 31)) & 1;
  }
  return MG_OK;
}

int mg_dump_objects(mg_handle* h, int env, int32_t* out, int max_rows) {
  if (!h || !out || env < 0 || env >= h->d.num_envs) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  const MgDev& d = h->d;
  if (int rc = sync_host_view(h)) return rc;
  int32_t E[MGEV_WORDS];
  CK(cudaMemcpy(E, d.env + (size_t)env * MGEV_WORDS, sizeof E, cudaMemcpyDeviceToHost));
  int nobj = E[MGEV_NEXT_OBJ];
  std::vector<uint32_t> objs((size_t)nobj * d.OS);
  CK(cudaMemcpy(objs.data(), d.objs + (size_t)env * (d.maxobj + d.NPROXY) * d.OS, objs.size() * 4, cudaMemcpyDeviceToHost));
  const int32_t* P = h->program.data();
  int n = 0, stride = 8 + 2 * d.R + d.TW;
  for (int s = 1; s < nobj && n < max_rows; s++) {
    const uint32_t* o = &objs[(size_t)s * d.OS];
    int flags = o[MGO_META] >> 24;
    if (!(flags & MGOF_ALIVE)) continue;
    int32_t* row = out + (size_t)n * stride;
    int t = o[MGO_META] & 0xffff;
    row[0] = (int32_t)o[MGO_ID];
    row[1] = P[P[MGS_TEMPLATES] + t * MG_TEMPLATE_WORDS + MGT_TYPE_ID];
    row[2] = o[MGO_LOC] >> 16, row[3] = o[MGO_LOC] & 0xffff, row[4] = (o[MGO_META] >> 16) & 0xff;
    row[5] = (int32_t)o[MGO_AGENT], row[6] = (int32_t)o[MGO_TAGS];
    uint64_t ord = (uint64_t)o[MGO_INVORD_LO] | ((uint64_t)o[MGO_INVORD_HI] << 32);
    int cnt = (int)(ord >> 60);
    row[7] = cnt;
    const uint16_t* inv = (const uint16_t*)(o + MGO_TAGS + d.TW);
    for (int i = 0; i < d.R; i++) row[8 + i] = inv[i];
    for (int i = 0; i < d.R; i++) row[8 + d.R + i] = i < cnt ? (int)((ord >> (4 * i)) & 15) : -1;
    for (int k = 0; k < d.TW; k++) row[8 + 2 * d.R + k] = (int32_t)o[MGO_TAGS + k];
    n++;
  }
  return n;
}

int mg_get_agent_state(mg_handle* h, int env, int32_t* out) {
  if (!h || !out || env < 0 || env >= h->d.num_envs) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  const MgDev& d = h->d;
  if (int rc = sync_host_view(h)) return rc;
  std::vector<uint32_t> ag((size_t)d.A * d.AS);
  CK(cudaMemcpy(ag.data(), d.agents + (size_t)env * d.A * d.AS, ag.size() * 4, cudaMemcpyDeviceToHost));
  const int32_t* P = h->program.data();
  for (int a = 0; a < d.A; a++) {
    const uint32_t* r = &ag[(size_t)a * d.AS];
    uint32_t obj[MGO_NTOK + 1];
    CK(cudaMemcpy(obj, d.objs + ((size_t)env * (d.maxobj + d.NPROXY) + r[MGAG_OBJ]) * d.OS, sizeof obj, cudaMemcpyDeviceToHost));
    const int t = obj[MGO_META] & 0xffff;
    float total = 0.0f;  // RewardHelper::current_reward (systems/reward.hpp:36-42): prev values added in entry order
    for (int i = MGAG_REWARD_PREV; i < d.AS; i++) {
      float v;
      memcpy(&v, &r[i], 4);
      total += v;
    }
    int32_t* o = out + 4 * a;
    o[0] = (int32_t)obj[MGO_ID];
    o[1] = P[P[MGS_TEMPLATES] + t * MG_TEMPLATE_WORDS + MGT_GROUP];
    o[2] = (int32_t)r[MGAG_SWM];
    memcpy(&o[3], &total, 4);
  }
  return MG_OK;
}

int mg_set_inventory(mg_handle* h, int env, int agent, const int32_t* items, const int32_t* amounts, int n) {
  if (!h || env < 0 || env >= h->d.num_envs || agent < 0 || agent >= h->d.A || n < 0 || n > h->d.R || (n && (!items || !amounts)))
    return MG_E_INVALID;
  for (int i = 0; i < n; i++)
    if (items[i] < 0 || items[i] >= h->d.R || amounts[i] < 0 || amounts[i] > 65535) {
      h->err = "mg_set_inventory: resource id or amount out of range";
      return MG_E_INVALID;
    }
  CK(cudaSetDevice(h->device));
  int32_t* buf = nullptr;
  CK(cudaMalloc((void**)&buf, (size_t)(2 * n + 2) * 4));
  cudaError_t e = cudaSuccess;
  if (n) {
    e = cudaMemcpy(buf, items, (size_t)n * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(buf + n, amounts, (size_t)n * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess && sync_host_view(h) != MG_OK) e = cudaErrorUnknown;
  if (e == cudaSuccess) e = mg_launch_set_inventory(h->d, env, agent, buf, buf + n, n, h->own_stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->own_stream);
  cudaFree(buf);
  CK(e);
  if (h->fast) h->newest = mg_handle::GENERIC;
  h->snap_valid = false, h->pristine = false;
  return MG_OK;
}

int mg_num_envs(const mg_handle* h) { return h ? h->d.num_envs : 0; }
int mg_num_agents(const mg_handle* h) { return h ? h->d.A : 0; }
int mg_num_tokens(const mg_handle* h) { return h ? h->d.T : 0; }
int mg_vecenv_configure(mg_handle* h, int num_primary, const int32_t* vibe_action_ids, int num_vibe_actions,
                        const int64_t* early_reset_steps) {
  if (!h || num_primary <= 0 || num_vibe_actions < 0 || (num_vibe_actions && !vibe_action_ids)) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  drop_vecenv_graph(h);
  const size_t N = (size_t)h->d.num_envs;
  int rc;
  if (!h->ve_done) {
    if ((rc = dev_alloc(h, &h->ve_done, N)) || (rc = dev_alloc(h, &h->ve_steps, N)) || (rc = dev_alloc(h, &h->ve_counters, 2)) ||
        (rc = dev_alloc(h, &h->ve_vibe_ids, (size_t)(num_vibe_actions > 0 ? num_vibe_actions : 1))))
      return rc;
  } else if (num_vibe_actions > h->ve_vibes) {
    h->err = "mg_vecenv_configure: the vibe action list cannot grow after the first call";
    return MG_E_INVALID;
  }
  h->ve_primary = num_primary, h->ve_vibes = num_vibe_actions;
  if (num_vibe_actions) CK(cudaMemcpy(h->ve_vibe_ids, vibe_action_ids, (size_t)num_vibe_actions * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(h->ve_steps, 0, N * 8));
  CK(cudaMemset(h->ve_counters, 0, 8));
  if (early_reset_steps) {
    if (!h->ve_early && (rc = dev_alloc(h, &h->ve_early, N))) return rc;
    CK(cudaMemcpy(h->ve_early, early_reset_steps, N * 8, cudaMemcpyHostToDevice));
  } else if (h->ve_early) {
    std::vector<int64_t> never(N, INT64_MAX);
    CK(cudaMemcpy(h->ve_early, never.data(), N * 8, cudaMemcpyHostToDevice));
  }
  return MG_OK;
}

int mg_vecenv_step(mg_handle* h, const void* actions, int is_int64, int ncols, int auto_reset, void* stream) {
  if (!h || !actions || (ncols != 1 && ncols != 2)) return MG_E_INVALID;
  if (!h->ve_done || !h->buffers_set) {
    h->err = "mg_vecenv_step: call mg_set_buffers and mg_vecenv_configure first";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  MgDev& d = h->d;
  h->foreign_pending = true, h->last_stream = st;
  static const bool no_graph = getenv("METTAGRID_B200_NO_GRAPH") != nullptr;
  if (auto_reset && h->fast && h->snap_valid && h->newest == mg_handle::PACKED && !no_graph && h->ve_g_misses < 4) {
    // steady state of a training loop: the same four launches every step -> one graph launch (the tick itself is
    // ~19 us at 4096 envs, so three extra launch gaps are a fifth of the step)
    if (h->ve_exec && h->ve_g_actions == actions && h->ve_g_int64 == is_int64 && h->ve_g_ncols == ncols) {
      h->ve_g_misses = 0;
    } else {
      if (h->ve_exec) h->ve_g_misses++;
      drop_vecenv_graph(h);
      cudaStream_t cs = h->own_stream;  // captured here (the caller's stream may be the legacy stream), launched on `st`
      CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
      cudaError_t e = mg_launch_vecenv_prepare(actions, is_int64, ncols, d.num_envs, d.A, h->ve_primary, h->ve_vibes, h->ve_vibe_ids,
                                               (int32_t*)d.actions, (int32_t*)d.vibe_actions, d.terminals, d.truncations,
                                               h->ve_done, h->ve_steps, h->ve_counters, cs);
      if (e == cudaSuccess) e = mg_launch_fast_restore(d, h->fl, h->snapshot, h->ve_done, cs);
      if (e == cudaSuccess) e = mg_launch_step_fast(d, h->fl, h->fh, cs);
      if (e == cudaSuccess) e = mg_launch_vecenv_post(d.num_envs, d.A, h->ve_steps, h->ve_early, d.truncations, cs);
      cudaGraph_t g = nullptr;
      const cudaError_t e2 = cudaStreamEndCapture(cs, &g);
      if (e == cudaSuccess) e = e2;
      if (e == cudaSuccess) e = cudaGraphInstantiate(&h->ve_exec, g, 0);
      if (g) cudaGraphDestroy(g);
      if (e != cudaSuccess) {
        h->ve_exec = nullptr;
        h->err = std::string("mg_vecenv_step: graph capture failed: ") + cudaGetErrorString(e);
        return MG_E_CUDA;
      }
      h->ve_g_actions = actions, h->ve_g_int64 = is_int64, h->ve_g_ncols = ncols;
    }
    CK(cudaGraphLaunch(h->ve_exec, st));
    h->pristine = false;
    return MG_OK;
  }
  CK(mg_launch_vecenv_prepare(actions, is_int64, ncols, d.num_envs, d.A, h->ve_primary, h->ve_vibes, h->ve_vibe_ids,
                              (int32_t*)d.actions, (int32_t*)d.vibe_actions, d.terminals, d.truncations, h->ve_done,
                              h->ve_steps, h->ve_counters, st));
  if (auto_reset) {  // rebuild what finished on the previous step
    if (h->fast && h->snap_valid && h->newest == mg_handle::PACKED) {
      CK(mg_launch_fast_restore(d, h->fl, h->snapshot, h->ve_done, st));  // a copy of the post-reset snapshot
    } else if (int rc = mg_reset(h, h->ve_done, nullptr, stream)) {  // CTAs without a selected env leave at once
      return rc;
    }
  }
  if (int rc = launch_step(h, d, st)) return rc;
  CK(mg_launch_vecenv_post(d.num_envs, d.A, h->ve_steps, h->ve_early, d.truncations, st));
  return MG_OK;
}

int mg_vecenv_poll(mg_handle* h, int* episodes_finished, int* error_bits) {
  if (!h || !h->ve_counters) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (int rc = sync_streams(h)) return rc;
  int c[2];
  CK(cudaMemcpy(c, h->ve_counters, sizeof c, cudaMemcpyDeviceToHost));
  if (episodes_finished) *episodes_finished = c[0];
  if (error_bits) *error_bits = c[1];
  if (c[1]) CK(cudaMemset(h->ve_counters + 1, 0, 4));  // reported once
  return MG_OK;
}

int mg_grid_obs_configure(mg_handle* h, int num_features, const float* scale) {
  if (!h || !scale || num_features <= 0 || num_features > 256) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  if (!h->grid_scale) {
    int rc = dev_alloc(h, &h->grid_scale, 256);
    if (rc) return rc;
  }
  float s[256];
  for (int i = 0; i < 256; i++) s[i] = scale[i] > 1.0f ? scale[i] : 1.0f;  // grid_obs_wrapper.py:43-45
  CK(cudaMemcpy(h->grid_scale, s, sizeof s, cudaMemcpyHostToDevice));
  h->grid_features = num_features;
  return MG_OK;
}

int mg_obs_to_grid(mg_handle* h, const void* observations, int rows, void* grid, void* stream) {
  if (!h || !grid || rows <= 0) return MG_E_INVALID;
  if (!h->grid_scale) {
    h->err = "mg_obs_to_grid: call mg_grid_obs_configure first";
    return MG_E_INVALID;
  }
  const uint8_t* obs = observations ? (const uint8_t*)observations : h->d.obs;
  if (!obs) {
    h->err = "mg_obs_to_grid: no observation buffer (pass one or call mg_set_buffers)";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  CK(mg_launch_obs_to_grid(obs, (float*)grid, rows, h->d.T, h->grid_features, h->program[MGH_OBS_H], h->program[MGH_OBS_W],
                           h->grid_scale, (cudaStream_t)stream));
  return MG_OK;
}

int mg_token_summary(mg_handle* h, const void* observations, int rows, const void* pos_x, const void* pos_y, const void* feat,
                     const void* scale, int hidden, int num_feat, void* out, void* stream) {
  if (!h || !pos_x || !pos_y || !feat || !scale || !out || rows <= 0 || num_feat <= 0) return MG_E_INVALID;
  if (hidden <= 0 || hidden > 256) {
    h->err = "mg_token_summary: hidden must be in [1, 256]";
    return MG_E_INVALID;
  }
  const uint8_t* obs = observations ? (const uint8_t*)observations : h->d.obs;
  if (!obs) {
    h->err = "mg_token_summary: no observation buffer (pass one or call mg_set_buffers)";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  CK(mg_launch_token_summary(obs, rows, h->d.T, hidden, num_feat, (const float*)pos_x, (const float*)pos_y, (const float*)feat,
                             (const float*)scale, (float*)out, (cudaStream_t)stream));
  return MG_OK;
}

int mg_info_configure(mg_handle* h, int n_game, const int32_t* game_stat_ids, int n_agent, const int32_t* agent_stat_ids) {
  if (!h || n_game < 0 || n_agent < 0 || (n_game && !game_stat_ids) || (n_agent && !agent_stat_ids)) return MG_E_INVALID;
  for (int i = 0; i < n_game; i++)
    if (game_stat_ids[i] < -1 || game_stat_ids[i] >= h->d.SG) return MG_E_INVALID;
  for (int i = 0; i < n_agent; i++)
    if (agent_stat_ids[i] < -2 || agent_stat_ids[i] >= h->d.SA) return MG_E_INVALID;
  CK(cudaSetDevice(h->device));
  int rc;
  if ((rc = dev_alloc(h, &h->info_game_ids, (size_t)(n_game ? n_game : 1))) || (rc = dev_alloc(h, &h->info_agent_ids, (size_t)(n_agent ? n_agent : 1))))
    return rc;
  if (n_game) CK(cudaMemcpy(h->info_game_ids, game_stat_ids, (size_t)n_game * 4, cudaMemcpyHostToDevice));
  if (n_agent) CK(cudaMemcpy(h->info_agent_ids, agent_stat_ids, (size_t)n_agent * 4, cudaMemcpyHostToDevice));
  h->info_ngame = n_game, h->info_nagent = n_agent;
  return MG_OK;
}

int mg_info_gather(mg_handle* h, void* game_values, void* game_present, void* agent_values, void* agent_present, void* stream) {
  if (!h) return MG_E_INVALID;
  if ((h->info_ngame && (!game_values || !game_present)) || (h->info_nagent && (!agent_values || !agent_present))) {
    h->err = "mg_info_gather: an output tensor is missing for the configured keys";
    return MG_E_INVALID;
  }
  if (!h->buffers_set) {
    h->err = "mg_info_gather: call mg_set_buffers first";
    return MG_E_INVALID;
  }
  CK(cudaSetDevice(h->device));
  const int32_t* P = h->program.data();
  h->last_stream = (cudaStream_t)stream;
  CK(mg_launch_gather_info(h->d, h->fast && h->newest == mg_handle::PACKED, h->fl.G, h->info_ngame, h->info_game_ids, h->info_nagent,
                           h->info_agent_ids, P[MGH_GST_TOKENS_WRITTEN], P[MGH_GST_TOKENS_DROPPED], P[MGH_GST_TOKENS_FREE],
                           (float*)game_values, (uint8_t*)game_present, (float*)agent_values, (uint8_t*)agent_present,
                           (cudaStream_t)stream));
  return MG_OK;
}

size_t mg_state_bytes(const mg_handle* h) { return h ? h->bytes : 0; }
int mg_step_kernel(const mg_handle* h) {
  return !h ? MG_E_INVALID : h->fast ? h->fl.G + (h->fl.statics ? 64 : 0) : h->d.plain ? 1 : 0;
}

}  // extern "C"
