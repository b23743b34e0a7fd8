// mg_gridobs.cu -- k_obs_to_grid: sparse observation tokens -> dense float32 (C, H, W) grids on the device
// (SURVEY 8f-3).  Replaces GridObsWrapper._convert (python/src/mettagrid/envs/grid_obs_wrapper.py:58-96):
// token [coord, feature, value]; coord 0xFF = padding (skipped), 0xFE = global token (placed at the window
// centre), otherwise y = high nibble, x = low nibble; grid[feature][y][x] += value / scale[feature].
//
// One warp per agent row.  The row is zeroed with 16-byte streaming stores, then the tokens are scattered 32 at a
// time.  numpy's add.at accumulates duplicates of one (feature, y, x) in token order; __match_any groups such
// duplicates inside a chunk and the lowest lane of a group adds them in lane order, chunks follow one another, so
// every sum is formed in the reference's order and the result is bit-exact.
#include <cuda_runtime.h>
#include <stdint.h>

#define FULL 0xffffffffu

namespace {

__global__ void __launch_bounds__(256) k_obs_to_grid(const uint8_t* __restrict__ obs, float* __restrict__ grid, int rows, int T,
                                                     int C, int H, int W, const float* __restrict__ scale) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t cells = (size_t)C * H * W;
  float* g = grid + (size_t)row * cells;
  {  // zero the row: scalar head up to 16-byte alignment, vector body, scalar tail
    const size_t head = min(cells, (size_t)(((16u - ((uint32_t)(uintptr_t)g & 15u)) & 15u) >> 2));
    if ((size_t)lane < head) g[lane] = 0.0f;
    float4* g4 = (float4*)(g + head);
    const size_t n4 = (cells - head) >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = lane; i < n4; i += 32) __stcs(g4 + i, z);
    const size_t done = head + (n4 << 2);
    if (done + lane < cells) g[done + lane] = 0.0f;
  }
  __syncwarp();
  const uint8_t* tok = obs + (size_t)row * T * 3;
  const int cy = H / 2, cx = W / 2;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int coord = 0xFF, fid = 0, val = 0;
    if (t < T) coord = tok[3 * t], fid = tok[3 * t + 1], val = tok[3 * t + 2];
    int y = (coord >> 4) & 15, x = coord & 15;
    if (coord == 0xFE) y = cy, x = cx;
    const bool valid = coord != 0xFF && y < H && x < W && fid < C;
    if (!__any_sync(FULL, valid)) continue;
    const uint32_t key = valid ? (uint32_t)((fid * H + y) * W + x) : (0x80000000u | (uint32_t)lane);
    const float v = valid ? __fdiv_rn((float)val, scale[fid]) : 0.0f;
    const uint32_t m = __match_any_sync(FULL, key);
    const bool leader = valid && (__ffs(m) - 1) == lane;
    const int cnt = __popc(m);
    const int maxc = __reduce_max_sync(FULL, valid ? cnt : 1);
    float acc = 0.0f;
    if (leader) acc = __fadd_rn(__ldcg(g + key), v);
    for (int r = 1; r < maxc; r++) {  // duplicates join in lane (= token) order
      const bool take = leader && r < cnt;
      const int src = take ? (int)__fns(m, 0, r + 1) : lane;
      const float vb = __shfl_sync(FULL, v, src);
      if (take) acc = __fadd_rn(acc, vb);
    }
    if (leader) g[key] = acc;
    __syncwarp();
  }
}

}  // namespace

cudaError_t mg_launch_obs_to_grid(const uint8_t* obs, float* grid, int rows, int T, int C, int H, int W, const float* scale,
                                  cudaStream_t st) {
  const int warps = 8;
  k_obs_to_grid<<<(rows + warps - 1) / warps, warps * 32, 0, st>>>(obs, grid, rows, T, C, H, W, scale);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// k_token_summary: the front end of the reference's token policy, TokenPolicyNet._encode_tokens up to the pooled
// summary (python/src/mettagrid/policy/token_encoder.py:89-113), fused over the observation rows where they lie:
//   summary[row, h] = sum over valid tokens of (pos_x[x][h] + pos_y[y][h] + feat[f][h]) * (value / (scale[f] + 1e-6))
//                     / sqrt(max(#valid, 1)),     valid = coord byte != 0xFF, x = low nibble, y = high nibble.
// The torch version materialises a [rows, T, hidden] float tensor (7.5 GB at C2) before pooling it; here a warp owns a
// row, lanes own hidden columns, the row's tokens are staged in shared memory, and only [rows, hidden] leaves the SM.
// The 16 rows of each position table a nibble can select are staged per CTA; the feature table is read through L1.
// Sums run in token order in fp32 (torch reduces the token axis in its own order: parity is a tolerance, 1e-5 relative).
// ---------------------------------------------------------------------------------------------------------------------
namespace {
#define MG_TS_WARPS 8
#define MG_TS_MAXH 8  // hidden columns per lane: hidden <= 256
__global__ void __launch_bounds__(MG_TS_WARPS * 32) k_token_summary(const uint8_t* __restrict__ obs, int rows, int T, int hidden,
                                                                   int num_feat, const float* __restrict__ pos_x,
                                                                   const float* __restrict__ pos_y, const float* __restrict__ feat,
                                                                   const float* __restrict__ scale, float* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char ts_smem[];
  float* spx = (float*)ts_smem;           // [16][hidden]
  float* spy = spx + 16 * hidden;         // [16][hidden]
  uint8_t* stok = (uint8_t*)(spy + 16 * hidden);  // [warps][3T rounded up to 16]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 16 * hidden; i += blockDim.x) spx[i] = pos_x[i], spy[i] = pos_y[i];
  __syncthreads();
  const int row = blockIdx.x * MG_TS_WARPS + warp;
  if (row >= rows) return;
  const int nb = 3 * T;
  uint8_t* tk = stok + (size_t)warp * ((nb + 15) & ~15);
  const uint8_t* src = obs + (size_t)row * nb;
  for (int i = lane; i < nb; i += 32) tk[i] = src[i];
  __syncwarp();
  float acc[MG_TS_MAXH];
#pragma unroll
  for (int j = 0; j < MG_TS_MAXH; j++) acc[j] = 0.0f;
  int count = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {  // 32 coordinate bytes per ballot: only valid tokens are visited
    const int tl = t0 + lane;
    uint32_t m = __ballot_sync(0xffffffffu, tl < T && tk[3 * tl] != 0xFF);
    count += __popc(m);
    while (m) {
      const int t = t0 + __ffs(m) - 1;
      m &= m - 1;
      const int coord = tk[3 * t];
      const int x = coord & 15, y = coord >> 4;
      const int f = min((int)tk[3 * t + 1], num_feat - 1);
      const float sv = __fdiv_rn((float)tk[3 * t + 2], __fadd_rn(__ldg(scale + f), 1e-6f));
      const float* fr = feat + (size_t)f * hidden;
#pragma unroll
      for (int j = 0; j < MG_TS_MAXH; j++) {
        const int h = lane + 32 * j;
        if (h < hidden) {
          const float e = __fadd_rn(__fadd_rn(spx[x * hidden + h], spy[y * hidden + h]), __ldg(fr + h));
          acc[j] = __fadd_rn(acc[j], __fmul_rn(e, sv));
        }
      }
    }
  }
  const float norm = sqrtf((float)max(count, 1));
#pragma unroll
  for (int j = 0; j < MG_TS_MAXH; j++) {
    const int h = lane + 32 * j;
    if (h < hidden) out[(size_t)row * hidden + h] = __fdiv_rn(acc[j], norm);
  }
}
}  // namespace

cudaError_t mg_launch_token_summary(const uint8_t* obs, int rows, int T, int hidden, int num_feat, const float* pos_x,
                                    const float* pos_y, const float* feat, const float* scale, float* out, cudaStream_t st) {
  const size_t smem = (size_t)32 * hidden * sizeof(float) + (size_t)MG_TS_WARPS * ((3 * T + 15) & ~15);
  cudaError_t e = cudaFuncSetAttribute(k_token_summary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_token_summary<<<(rows + MG_TS_WARPS - 1) / MG_TS_WARPS, MG_TS_WARPS * 32, smem, st>>>(obs, rows, T, hidden, num_feat, pos_x, pos_y,
                                                                                         feat, scale, out);
  return cudaGetLastError();
}
