/*
 * mettagrid_b200.h -- C ABI of the B200-native batched MettaGrid step.
 *
 * Drop-in boundary for the reference's pybind class `MettaGrid`
 * (/root/reference/cpp/bindings/mettagrid_c.hpp:58-320, bound in mettagrid_py.cpp:251-312): one handle
 * owns N independent environments that all run the same compiled game program
 * (include/mg_program.h).  Plain pointers and sizes only; every entry point returns 0 on success or
 * a negative MG_E_* code, with text available from mg_last_error().
 *
 * Buffers follow the reference's set_buffers contract (mettagrid_py.cpp:253-266,
 * mettagrid_c.cpp:1165-1184): the CALLER owns them, the library aliases them and re-runs
 * _init_buffers.  Layouts are the reference's with a leading env dimension:
 *   observations uint8 [N][A][T][3], terminals/truncations bool(uint8) [N][A], rewards float [N][A],
 *   actions / vibe_actions int32 [N][A].
 */
#ifndef METTAGRID_B200_H_
#define METTAGRID_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mg_handle mg_handle;

#define MG_OK 0
#define MG_E_INVALID -1     /* bad argument / bad program */
#define MG_E_CUDA -2        /* CUDA runtime error */
#define MG_E_UNSUPPORTED -3 /* program needs a feature this build does not run on the GPU */
#define MG_E_ENV -4         /* an environment raised an error (see mg_poll_errors) */

/* replaces MettaGrid::MettaGrid(GameConfig, map, seed) -- mettagrid_c.cpp:42-191.
 * program      : compiled game program (mettagrid_b200.compiler), nwords int32 words (host)
 * init_cells   : host int16 [num_envs][H][W], template index per cell, -1 = empty
 * init_gstats  : host float [num_envs][S_G] initial game stats ("objects.<cell>" counts) or NULL
 * seeds        : host uint32 [num_envs], one std::mt19937 seed per environment
 * device       : CUDA device ordinal */
int mg_create(const int32_t* program, size_t nwords, int num_envs, const int16_t* init_cells, const float* init_gstats,
              const uint32_t* seeds, int device, mg_handle** out);
void mg_destroy(mg_handle* h);
const char* mg_last_error(const mg_handle* h); /* h may be NULL: error of the last failed mg_create */

/* replaces MettaGrid::set_buffers -- mettagrid_c.cpp:1165-1184.  DEVICE pointers.  Runs
 * _init_buffers (initial observations) on `stream` (a cudaStream_t, may be NULL). */
int mg_set_buffers(mg_handle* h, void* observations, void* terminals, void* truncations, void* rewards,
                   const void* actions, const void* vibe_actions, void* stream);

/* replaces MettaGrid::step -- mettagrid_c.cpp:1186-1205.  Asynchronous on `stream`. */
int mg_step(mg_handle* h, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies actions in, steps, copies results out,
 * synchronises.  Any output pointer may be NULL. */
int mg_step_host(mg_handle* h, const int32_t* actions, const int32_t* vibe_actions, uint8_t* observations,
                 float* rewards, uint8_t* terminals, uint8_t* truncations);

/* reset = rebuild from the stored initial state, like constructing a new MettaGrid with the same
 * map and seed (envs/mettagrid_puffer_env.py:225-228).  env_mask: DEVICE uint8 [num_envs] or NULL
 * for all.  new_seeds: host uint32 [num_envs] or NULL to keep the seeds. */
int mg_reset(mg_handle* h, const uint8_t* env_mask, const uint32_t* new_seeds, void* stream);

/* Replace the stored initial state of one environment (SURVEY 8b "optional new initial state", 8f-1 map-instance
 * pool): the next mg_reset that selects `env` builds it from this map.  init_cells: host int16 [H][W] template
 * indices; init_gstats: host float [S_G] or NULL to keep the stored ones.  The map must hold the program's agent
 * count (else the reset raises the env's error word) and no more objects than the handle was sized for. */
int mg_set_map(mg_handle* h, int env, const int16_t* init_cells, const float* init_gstats);

/* first environment with a pending error: token-budget overflow (mettagrid_c.cpp:364-375) etc.
 * Returns MG_OK when none, MG_E_ENV otherwise.  code = MGERR_* bits, info = agent | attempted<<16. */
int mg_poll_errors(mg_handle* h, int* env, int* code, int* info);

/* introspection (mettagrid_py.cpp:161-239); all outputs are HOST buffers */
int mg_get_episode_rewards(mg_handle* h, float* out /*[N][A]*/);
int mg_get_action_success(mg_handle* h, uint8_t* out /*[N][A]*/);
int mg_get_current_steps(mg_handle* h, int32_t* out /*[N]*/);
int mg_get_agent_stats(mg_handle* h, int env, float* values /*[A][S_A]*/, uint8_t* touched /*[A][S_A]*/);
int mg_get_game_stats(mg_handle* h, int env, float* values /*[S_G]*/, uint8_t* touched /*[S_G]*/);
/* one row of (8 + 2R + TW) int32 per live object in id order:
 * id, type_id, r, c, vibe, agent (-1), tag word 0, #present resources, inv[R], order[R] (-1 padded), tags[TW] */
int mg_dump_objects(mg_handle* h, int env, int32_t* out, int max_rows);
/* per agent of one env, 4 int32: object id, group id, steps_without_motion, current_stat_reward (float bits;
 * RewardHelper::current_reward, systems/reward.hpp:36-42) -- with mg_dump_objects everything grid_objects()
 * (mettagrid_py.cpp:28-139) reports that the deterministic episode signature reads */
int mg_get_agent_state(mg_handle* h, int env, int32_t* out /*[A][4]*/);

/* replaces MettaGrid::set_inventory(agent_id, inventory) -- mettagrid_py.cpp:203-209, objects/agent.cpp:86-104.
 * items/amounts: host int32 [n], in the iteration order of the unordered_map pybind11 builds from the Python
 * dict (mettagrid_b200.compiler.pybind_dict_order). */
int mg_set_inventory(mg_handle* h, int env, int agent, const int32_t* items, const int32_t* amounts, int n);

/* Vectorised-env step (next row 8f-2): what MettaGridPufferEnv.step does around the C++ step
 * (python/src/mettagrid/envs/mettagrid_puffer_env.py:296-408), entirely on the device and asynchronous.
 * mg_vecenv_configure: the number of primary actions P, the env action ids of the V vibe actions (host int32 [V]),
 * and optionally one early-reset step per env (host int64 [N], EarlyResetHandler: envs/early_reset_handler.py).
 * mg_vecenv_step: `actions` is a DEVICE tensor of N * A combined indices (ncols = 1: values < P are primary actions,
 * values in [P, P + P*V) encode primary = off / V, vibe = off % V) or N * A (primary, vibe index) pairs (ncols = 2),
 * int32 or int64.  With auto_reset, environments whose agents were all terminal or all truncated after the
 * previous step are rebuilt first (:299-302).  Out-of-range actions raise error bits (1 negative, 2 out of range,
 * 4 vibe index out of range, 8 no vibe action space, 16 primary out of range) that mg_vecenv_poll returns with
 * the number of finished episodes. */
int mg_vecenv_configure(mg_handle* h, int num_primary, const int32_t* vibe_action_ids, int num_vibe_actions,
                        const int64_t* early_reset_steps);
int mg_vecenv_step(mg_handle* h, const void* actions, int is_int64, int ncols, int auto_reset, void* stream);
int mg_vecenv_poll(mg_handle* h, int* episodes_finished, int* error_bits);

/* step_info_keys (next row 8f-2): replaces MettaGridPufferEnv._build_step_info_payload --
 * python/src/mettagrid/envs/mettagrid_puffer_env.py:132-282 -- whose per-step get_game_stat / get_agent_stat calls a
 * batch cannot afford.  mg_info_configure fixes the stats to export: game stat ids (-1 = attributes/steps) and agent
 * stat ids (-1 = reward_step, -2 = reward_episode); mg_info_gather writes, asynchronously on `stream`, float32
 * [N][n_game] + uint8 presence and float32 [N][A][n_agent] + uint8 presence (all DEVICE; "present" = the reference's
 * "value is not None": the stat was touched this episode). */
int mg_info_configure(mg_handle* h, int n_game, const int32_t* game_stat_ids, int n_agent, const int32_t* agent_stat_ids);
int mg_info_gather(mg_handle* h, void* game_values, void* game_present, void* agent_values, void* agent_present, void* stream);

/* Dense grid observations (next row 8f-3): replaces GridObsWrapper._convert --
 * python/src/mettagrid/envs/grid_obs_wrapper.py:33-96.  mg_grid_obs_configure fixes the number of feature planes C
 * (max feature id + 1) and the per-feature normalisation (host float[256], entries < 1 are raised to 1 like the
 * wrapper does).  mg_obs_to_grid converts `rows` token rows uint8 [rows][T][3] (DEVICE; NULL = the handle's
 * observation buffer, rows = N * A) into float32 [rows][C][obs_height][obs_width] (DEVICE), asynchronously. */
int mg_grid_obs_configure(mg_handle* h, int num_features, const float* scale);
int mg_obs_to_grid(mg_handle* h, const void* observations, int rows, void* grid, void* stream);

/* Token-policy front end (next row 8f-3): replaces TokenPolicyNet._encode_tokens up to the pooled summary --
 * python/src/mettagrid/policy/token_encoder.py:89-113.  observations: uint8 [rows][T][3] (DEVICE; NULL = the handle's
 * buffer); pos_x / pos_y: float32 [>=16][hidden] (nn.Embedding weights; a nibble selects rows 0-15), feat: float32
 * [num_feat][hidden], scale: float32 [num_feat] (the module's _feature_scale), out: float32 [rows][hidden]; all DEVICE,
 * hidden <= 256; asynchronous.  Forward only: the token_mlp / heads behind it stay torch modules. */
int mg_token_summary(mg_handle* h, const void* observations, int rows, const void* pos_x, const void* pos_y, const void* feat,
                     const void* scale, int hidden, int num_feat, void* out, void* stream);

/* sizes */
int mg_num_envs(const mg_handle* h);
int mg_num_agents(const mg_handle* h);
int mg_num_tokens(const mg_handle* h);
size_t mg_state_bytes(const mg_handle* h); /* HBM held by the handle */
/* which step kernel mg_step launches for this handle: 0 = generic (handler interpreter), 1 = plain (handler-free
 * program), 8 / 16 / 32 = the sparse fast path with that many lanes per environment, 64 + lanes = the same with the
 * static layer (walls kept as a bitmap outside the lanes) */
int mg_step_kernel(const mg_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* METTAGRID_B200_H_ */
