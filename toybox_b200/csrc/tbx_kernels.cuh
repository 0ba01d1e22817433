/* tbx_kernels.cuh -- the sm_100a thread-per-env kernels over the word-major state planes: step (transition +
 * env-level bookkeeping + auto-reset), new_game, seeding, scalar reads, the synthetic action stream and the
 * record gather/scatter behind JSON import/export.  The render kernel lives in tbx_render.cuh.
 */
#ifndef TBX_KERNELS_CUH
#define TBX_KERNELS_CUH
#include <cuda_runtime.h>
#include "tbx_render.cuh"
#include "tbx_render_area.cuh"
#include "tbx_render_native.cuh"
#include "tbx_host.h"

namespace tbxk {

/* ------------------------------------------------------------------ step */
struct StepArgs {
  uint32_t *planes;
  int n, n_pad;
  const void *cfg, *tables;
  const int32_t *actions; /* ALE ids, or NULL */
  const uint8_t *inputs;  /* Input bitmasks, or NULL */
  /* both NULL: the benchmark's synthetic stream, generated in the kernel (no separate launch): env i takes
   * legal[tbx_action_index(syn_seed, syn_env0 + i, syn_t, n_legal)] */
  const int32_t *legal;
  int n_legal;
  uint64_t syn_seed, syn_env0, syn_t;
  int auto_reset;
  int32_t *reward, *score, *lives;
  uint8_t *done;
  unsigned long long *stats; /* episodes, sum return, sum length, max return */
  int *bad_actions;
};

/* Thread per env, two variants.
 * STAGED: the block's records are brought into shared memory first.  Between two steps the render kernel writes hundreds of MB of
 * observations, so the state planes are no longer in L2, and a transition that reads them field by field is a chain of dependent
 * DRAM round trips (Breakout, measured: 86 cycles per warp instruction, 11 % issue-active).  Staged, every state word of the block
 * travels as 16-byte asynchronous copies (cp.async, all in flight at once: word w of the block's envs is one contiguous run), the
 * transition runs on shared memory (column `tid`, stride EPB words: conflict-free) and the columns are stored back, coalesced.
 * LAZY: the transition reads and writes the planes directly.
 * Measured per 65,536 envs in steady state (profiles/r2_step_variants.md): Space Invaders -- whose transition walks its 36 invaders
 * several times -- 90 us staged against 178 us lazy; Breakout 39 / 37 us and Amidar 69 / 61 us gain nothing (their time is f64
 * latency and divergence, not memory), so only Space Invaders is staged by default. */
template <int GAME> struct StepGeom { static constexpr int EPB = Traits<GAME>::RW * 128 * 4 <= 48 * 1024 ? 128 : Traits<GAME>::RW * 64 * 4 <= 64 * 1024 ? 64 : 32; };

template <int GAME>
__device__ __forceinline__ void step_env(const StepArgs &a, const TbxAcc &S, int env) {
  typedef Traits<GAME> T;
  const typename T::Cfg &cfg = *(const typename T::Cfg *)a.cfg;
  const typename T::Table *tables = (const typename T::Table *)a.tables;
  int in = a.actions ? tbx_ale_action_to_input(a.actions[env])
           : a.inputs ? (int)a.inputs[env]
                      : tbx_ale_action_to_input(a.legal[tbx_action_index(a.syn_seed, a.syn_env0 + (uint64_t)env, a.syn_t, (uint32_t)a.n_legal)]);
  int lives_before = S.ldi(TBX_HW(lives));
  if (in < 0) atomicAdd(a.bad_actions, 1);
  else T::step(S, cfg, tables, in);
  TbxStepOut o = tbx_bookkeep(S, lives_before);
  if (a.reward) a.reward[env] = o.reward;
  if (a.done) a.done[env] = (uint8_t)o.done;
  if (a.score) a.score[env] = o.score;
  if (a.lives) a.lives[env] = o.lives;
  if (o.episode_ended) {
    atomicAdd(a.stats + 0, 1ull);
    atomicAdd(a.stats + 1, (unsigned long long)(long long)o.ep_return);
    atomicAdd(a.stats + 2, (unsigned long long)o.ep_len);
    atomicMax((long long *)(a.stats + 3), (long long)o.ep_return);
  }
  if (o.done && a.auto_reset) T::new_game(S, cfg, tables);
}

template <int GAME>
__global__ void __launch_bounds__(128) step_kernel(StepArgs a) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= a.n) return;
  TbxAcc S;
  S.p = a.planes + env;
  S.stride = (size_t)a.n_pad;
  step_env<GAME>(a, S, env);
}

template <int GAME>
__global__ void __launch_bounds__(StepGeom<GAME>::EPB) step_staged_kernel(StepArgs a) {
  constexpr int EPB = StepGeom<GAME>::EPB, RW = Traits<GAME>::RW, CPW = EPB / 4; /* 16-byte chunks per state word */
  extern __shared__ uint4 srec_raw[];
  uint32_t *srec = reinterpret_cast<uint32_t *>(srec_raw); /* [RW][EPB] */
  const int tid = threadIdx.x, env0 = blockIdx.x * EPB, env = env0 + tid;
  /* n_pad is a multiple of 32 and the planes are padded, so whole 16-byte chunks are readable; EPB divides n_pad or the grid's
   * last block is clipped to the padded range */
  const int n_chunks = min(EPB, a.n_pad - env0) / 4;
  for (int i = tid; i < RW * CPW; i += EPB) {
    const int w = i / CPW, c = i - w * CPW;
    if (c < n_chunks)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(srec + w * EPB + 4 * c)),
                   "l"(a.planes + (size_t)w * a.n_pad + env0 + 4 * c) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  if (env >= a.n) return;
  uint32_t *scol = srec + tid;
  TbxAcc S;
  S.p = scol;
  S.stride = (size_t)EPB;
  step_env<GAME>(a, S, env);
  uint32_t *gcol = a.planes + env;
#pragma unroll 8
  for (int w = 0; w < RW; w++) gcol[(size_t)w * a.n_pad] = scol[w * EPB];
}

template <int GAME>
__global__ void __launch_bounds__(128) new_game_kernel(uint32_t *planes, int n, int n_pad, const void *cfg_, const void *tables_, const uint8_t *mask) {
  typedef Traits<GAME> T;
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n || (mask && !mask[env])) return;
  TbxAcc S;
  S.p = planes + env;
  S.stride = (size_t)n_pad;
  T::new_game(S, *(const typename T::Cfg *)cfg_, (const typename T::Table *)tables_);
}

/* sim_rand <- the pool config's rand (pool creation), or <- seed mapping (tbx_seed) */
__global__ void set_sim_rand_kernel(uint32_t *planes, int n, int n_pad, uint64_t s0, uint64_t s1) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  TbxAcc S; S.p = planes + env; S.stride = (size_t)n_pad;
  S.st64(TBX_HW(sim_rand), s0); S.st64(TBX_HW(sim_rand) + 2, s1);
}
__global__ void seed_kernel(uint32_t *planes, int n_pad, const uint32_t *seeds, const int32_t *ids, int k) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k) return;
  int env = ids ? ids[i] : i;
  TbxAcc S; S.p = planes + env; S.stride = (size_t)n_pad;
  TbxRng g;
  tbx_rng_seed(g, seeds[i]);
  S.st64(TBX_HW(sim_rand), g.s0); S.st64(TBX_HW(sim_rand) + 2, g.s1);
}
__global__ void read_scalars_kernel(const uint32_t *planes, int n, int n_pad, int32_t *score, int32_t *lives, int32_t *level) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  if (score) score[env] = (int32_t)planes[(size_t)TBX_HW(score) * n_pad + env];
  if (lives) lives[env] = (int32_t)planes[(size_t)TBX_HW(lives) * n_pad + env];
  if (level) level[env] = (int32_t)planes[(size_t)TBX_HW(level) * n_pad + env];
}
__global__ void fill_actions_kernel(int32_t *actions, int n, uint64_t seed, uint64_t env0, uint64_t t, const int32_t *legal, int n_legal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  actions[i] = legal[tbx_action_index(seed, env0 + (uint64_t)i, t, (uint32_t)n_legal)];
}
/* the same with the frame counter read from device memory, so that a captured CUDA graph advances the stream on replay */
__global__ void fill_actions_at_kernel(int32_t *actions, int n, uint64_t seed, uint64_t env0, const uint64_t *t, const int32_t *legal, int n_legal) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  actions[i] = legal[tbx_action_index(seed, env0 + (uint64_t)i, *t, (uint32_t)n_legal)];
}
/* A scripted Breakout policy for benchmarking states deep into a game (the random stream rarely breaks a brick):
 * FIRE while waiting for a serve, otherwise move the paddle under the first ball with a slowly varying aim
 * offset so every paddle segment gets used. */
__global__ void breakout_tracking_actions_kernel(const uint32_t *planes, int n, int n_pad, int32_t *actions, uint64_t t) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  TbxAcc S; S.p = const_cast<uint32_t *>(planes) + env; S.stride = (size_t)n_pad;
  int a = 0;
  if (S.ldi(BRK_W(is_dead))) a = 1;
  else {
    double bx = S.ldd(BRK_W(ball)) + (double)((int)((env * 7 + t / 50) % 9) - 4), px = S.ldd(BRK_W(paddle_px));
    a = bx > px + 1.0 ? 3 : bx < px - 1.0 ? 4 : 0;
  }
  actions[env] = a;
}
/* One property of the state schema for every env (tbx_fields.h): the plane of its word.  kind: 0 i32, 1 f64, 2 bool,
 * 3 bit of a mask word, 4 Option<i32> (None <-> TBX_NONE).  out / values: int32[N] (kinds 0, 2, 3, 4) or f64[N] (kind 1);
 * mask (may be NULL): uint8[N], only flagged envs are written. */
__global__ void field_get_kernel(const uint32_t *planes, int n, int n_pad, int word, int kind, int bit, int32_t *out_i, double *out_d) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n) return;
  const uint32_t v = planes[(size_t)word * n_pad + env];
  if (kind == 1) {
    const uint64_t u = (uint64_t)v | ((uint64_t)planes[(size_t)(word + 1) * n_pad + env] << 32);
    out_d[env] = __longlong_as_double((long long)u);
  } else if (kind == 3) out_i[env] = (int32_t)((v >> bit) & 1u);
  else if (kind == 2) out_i[env] = v != 0;
  else out_i[env] = (int32_t)v;
}
__global__ void field_set_kernel(uint32_t *planes, int n, int n_pad, int word, int kind, int bit, const int32_t *val_i, const double *val_d,
                                 const uint8_t *mask, int also_word) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n || (mask && !mask[env])) return;
  uint32_t *p = planes + (size_t)word * n_pad + env;
  if (kind == 1) {
    const uint64_t u = (uint64_t)__double_as_longlong(val_d[env]);
    p[0] = (uint32_t)u;
    p[n_pad] = (uint32_t)(u >> 32);
  } else if (kind == 3) *p = val_i[env] ? (*p | (1u << bit)) : (*p & ~(1u << bit));
  else if (kind == 2) *p = val_i[env] != 0;
  else {
    *p = (uint32_t)val_i[env];
    if (also_word >= 0) planes[(size_t)also_word * n_pad + env] = (uint32_t)val_i[env]; /* score: prev_score follows, as write_state_json does */
  }
}

/* BreakoutIntervention.add_channel / fill_column (toybox/interventions/breakout.py:406-416) and ctoybox's channel count for every
 * (masked) env, straight on the alive-mask words: op 0 = every brick of column `col` dead, 1 = alive again, 2 = out[env] = number
 * of columns with no alive brick.  A brick's column is the `col` field of the env's OWN brick table (envs whose bricks were edited
 * through JSON carry their own), columns 0..63 as in query_state_json('count_channels'). */
__global__ void brk_column_kernel(uint32_t *planes, int n, int n_pad, const BrkTable *tables, int op, int col, const uint8_t *mask, int32_t *out) {
  int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= n || (op != 2 && mask && !mask[env])) return;
  const BrkTable &T = tables[(int32_t)planes[(size_t)TBX_HW(tbl) * n_pad + env]];
  uint32_t alive[5];
  for (int k = 0; k < 5; k++) alive[k] = planes[(size_t)(BRK_W(alive) + k) * n_pad + env];
  if (op == 2) {
    unsigned long long has = 0, live = 0;
    for (int i = 0; i < T.n_bricks; i++) {
      const int c = T.col[i];
      if (c < 0 || c > 63) continue;
      has |= 1ull << c;
      if ((alive[i >> 5] >> (i & 31)) & 1u) live |= 1ull << c;
    }
    out[env] = __popcll(has & ~live);
    return;
  }
  for (int i = 0; i < T.n_bricks; i++)
    if (T.col[i] == col) { if (op) alive[i >> 5] |= 1u << (i & 31); else alive[i >> 5] &= ~(1u << (i & 31)); }
  for (int k = 0; k < 5; k++) planes[(size_t)(BRK_W(alive) + k) * n_pad + env] = alive[k];
}

/* records (AoS, rw words each) <-> planes, for the JSON import/export of a few envs */
__global__ void gather_kernel(const uint32_t *planes, int n_pad, const int32_t *ids, int k, int rw, uint32_t *recs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * rw) return;
  int e = i / rw, w = i - e * rw;
  recs[i] = planes[(size_t)w * n_pad + ids[e]];
}
__global__ void scatter_kernel(uint32_t *planes, int n_pad, const int32_t *ids, int k, int rw, const uint32_t *recs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= k * rw) return;
  int e = i / rw, w = i - e * rw;
  planes[(size_t)w * n_pad + ids[e]] = recs[i];
}

} /* namespace tbxk */
#endif
