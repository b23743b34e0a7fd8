// mg_vecenv.cu -- device-side glue of the vectorised env: what MettaGridPufferEnv.step does around
// the C++ step (python/src/mettagrid/envs/mettagrid_puffer_env.py:296-408), without a host round trip:
//   k_vecenv_prepare : which environments finished on the previous step (:299-302), and the combined-index /
//                      two-column action decoding (:313-394) straight into the handle's action buffers;
//   k_vecenv_post    : per-env step counters and the EarlyResetHandler truncation (envs/early_reset_handler.py:22-25).
#include <cuda_runtime.h>
#include <stdint.h>

#include "mg_state.h"

namespace {

// error bits raised by the decoder (the Python wrapper turns them into the reference's ValueErrors)
enum { VE_NEGATIVE = 1, VE_RANGE = 2, VE_VIBE_RANGE = 4, VE_NO_VIBES = 8, VE_CORE_RANGE = 16 };

template <class T>
__global__ void k_vecenv_prepare(const T* __restrict__ act, int ncols, int num_envs, int A, int P, int V,
                                 const int32_t* __restrict__ vibe_ids, int32_t* __restrict__ actions,
                                 int32_t* __restrict__ vibe_actions, const uint8_t* __restrict__ term,
                                 const uint8_t* __restrict__ trunc, uint8_t* __restrict__ done, int64_t* __restrict__ steps,
                                 int* __restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = num_envs * A;
  if (i < num_envs) {  // an env whose agents are all terminal or all truncated is rebuilt before this step
    bool all_term = true, all_trunc = true;
    for (int a = 0; a < A; a++) {
      all_term &= term[i * A + a] != 0;
      all_trunc &= trunc[i * A + a] != 0;
    }
    const bool dn = all_term || all_trunc;
    done[i] = dn;
    if (dn) {
      steps[i] = 0;
      atomicAdd(&counters[0], 1);
    }
  }
  if (i >= n) return;
  long long core, vibe = 0;
  int err = 0;
  if (ncols == 2) {
    core = (long long)act[2 * i];
    const long long raw = (long long)act[2 * i + 1];
    if (V <= 0)
      err |= VE_NO_VIBES;
    else if (raw < 0 || raw >= V)
      err |= VE_VIBE_RANGE;
    else
      vibe = vibe_ids[raw];
  } else {
    const long long a = (long long)act[i];
    core = a;
    if (a < 0) {
      err |= VE_NEGATIVE;
    } else if (a >= P) {
      if (V <= 0) {
        err |= VE_NO_VIBES;
      } else if (a >= (long long)P + (long long)P * V) {
        err |= VE_RANGE;
      } else {
        const long long off = a - P;
        core = off / V;
        vibe = vibe_ids[off % V];
      }
    }
  }
  if (!err && (core < 0 || core >= P)) err |= VE_CORE_RANGE;
  if (err) {
    atomicOr(&counters[1], err);
    core = 0, vibe = 0;
  }
  actions[i] = (int32_t)core;
  vibe_actions[i] = (int32_t)vibe;
}

__global__ void k_vecenv_post(int num_envs, int A, int64_t* __restrict__ steps, int64_t* __restrict__ early,
                              uint8_t* __restrict__ trunc) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= num_envs) return;
  const int64_t st = steps[e] + 1;
  steps[e] = st;
  if (early && st >= early[e]) {  // first episode only: end_episode() marks every agent truncated
    for (int a = 0; a < A; a++) trunc[e * A + a] = 1;
    early[e] = INT64_MAX;
  }
}

}  // namespace

cudaError_t mg_launch_vecenv_prepare(const void* act, int is_int64, int ncols, int num_envs, int A, int P, int V,
                                     const int32_t* vibe_ids, int32_t* actions, int32_t* vibe_actions, const uint8_t* term,
                                     const uint8_t* trunc, uint8_t* done, int64_t* steps, int* counters, cudaStream_t st) {
  const int n = num_envs * A, blocks = (n + 255) / 256;
  if (is_int64)
    k_vecenv_prepare<int64_t><<<blocks, 256, 0, st>>>((const int64_t*)act, ncols, num_envs, A, P, V, vibe_ids, actions,
                                                      vibe_actions, term, trunc, done, steps, counters);
  else
    k_vecenv_prepare<int32_t><<<blocks, 256, 0, st>>>((const int32_t*)act, ncols, num_envs, A, P, V, vibe_ids, actions,
                                                      vibe_actions, term, trunc, done, steps, counters);
  return cudaGetLastError();
}

cudaError_t mg_launch_vecenv_post(int num_envs, int A, int64_t* steps, int64_t* early, uint8_t* trunc, cudaStream_t st) {
  k_vecenv_post<<<(num_envs + 255) / 256, 256, 0, st>>>(num_envs, A, steps, early, trunc);
  return cudaGetLastError();
}

// ---- step_info_keys (mettagrid_puffer_env.py:132-282): the per-step stat export trainers read ---------------------
// One thread per (env, key) for the game keys and per (env, agent, key) for the agent keys gathers the configured stats
// into dense device tensors, so a training loop reads them without per-env host getters.  `present` mirrors the
// reference's "value is not None": a stat exists once it was touched.  A fast handle keeps the agent stats with ids
// below MGFB_STATS (and the token game stats) in its packed block while that is the newest copy.
namespace {
__global__ void k_gather_info(MgDev d, int packed, int G, int n_game, const int32_t* __restrict__ game_ids, int n_agent,
                              const int32_t* __restrict__ agent_ids, int idw, int idd, int idf, float* __restrict__ game_out,
                              uint8_t* __restrict__ game_present, float* __restrict__ agent_out, uint8_t* __restrict__ agent_present) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ng = (long long)d.num_envs * n_game, na = (long long)d.num_envs * d.A * n_agent;
  if (i < ng) {
    const int env = (int)(i / n_game), id = game_ids[i % n_game];
    const uint32_t* blk = packed ? d.fast_blk + (size_t)env * d.fast_stride : nullptr;
    float v = 0.0f;
    int have = 0;
    if (id == -1) {  // attributes/steps
      v = (float)(packed ? (int)blk[MGFB_STEP] : d.env[(size_t)env * MGEV_WORDS + MGEV_STEP]);
      have = 1;
    } else if (id >= 0 && id < d.SG) {
      if (packed && (id == idw || id == idd || id == idf)) {
        v = id == idw ? __uint_as_float(blk[MGFB_TOKENS_WRITTEN]) : id == idf ? __uint_as_float(blk[MGFB_TOKENS_FREE]) : 0.0f;
        have = (blk[MGFB_GTOUCHED] >> (id == idw ? 0 : id == idd ? 1 : 2)) & 1u;
      } else {
        v = d.gstats[(size_t)env * d.SG + id];
        have = (d.gtouched[(size_t)env * d.SGW + (id >> 5)] >> (id & 31)) & 1u;
      }
    }
    game_out[i] = have ? v : 0.0f;
    game_present[i] = (uint8_t)have;
    return;
  }
  const long long j = i - ng;
  if (j >= na) return;
  const int key = (int)(j % n_agent), id = agent_ids[key];
  const long long ga = j / n_agent;  // env * A + agent
  const int env = (int)(ga / d.A), a = (int)(ga % d.A);
  float v = 0.0f;
  int have = 0;
  if (id == -1) {  // reward_step
    v = d.rewards[ga], have = 1;
  } else if (id == -2) {  // reward_episode
    v = __uint_as_float(d.agents[(size_t)ga * d.AS + MGAG_EPISODE_REWARD]), have = 1;
  } else if (id >= 0 && id < d.SA) {
    if (packed && id < MGFB_STATS) {
      const uint32_t* blk = d.fast_blk + (size_t)env * d.fast_stride;
      v = __uint_as_float(blk[MGFB_STAT(G, id, a)]);
      have = (blk[MGFB_AGENT(G, a) + 6] >> id) & 1u;
    } else {
      v = d.astats[(size_t)ga * d.SA + id];
      have = (d.atouched[(size_t)ga * d.SAW + (id >> 5)] >> (id & 31)) & 1u;
    }
  }
  agent_out[j] = have ? v : 0.0f;
  agent_present[j] = (uint8_t)have;
}
}  // namespace

cudaError_t mg_launch_gather_info(const MgDev& d, int packed, int G, int n_game, const int32_t* game_ids, int n_agent,
                                  const int32_t* agent_ids, int idw, int idd, int idf, float* game_out, uint8_t* game_present,
                                  float* agent_out, uint8_t* agent_present, cudaStream_t st) {
  const long long n = (long long)d.num_envs * n_game + (long long)d.num_envs * d.A * n_agent;
  if (n <= 0) return cudaSuccess;
  k_gather_info<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, packed, G, n_game, game_ids, n_agent, agent_ids, idw, idd, idf, game_out,
                                                              game_present, agent_out, agent_present);
  return cudaGetLastError();
}

