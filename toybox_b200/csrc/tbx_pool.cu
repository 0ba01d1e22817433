/* tbx_pool.cu -- the C ABI of include/toybox_b200.h: pool life cycle, launches, JSON import/export. */
#include "../../include/toybox_b200.h"
#include "tbx_kernels.cuh"
#include "tbx_wrap.cuh"
#include "tbx_fields.h"
#include "tbx_direct_launch.h"
#include <map>
#include <new>
#include <string>
#include <string.h>
#include <algorithm>
#include <thread>
#include <vector>

using namespace tbxk;
using tbxjson::Value;

static thread_local std::string g_err;
static int set_err(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess) return set_err(TBX_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

struct AreaRes {
  TbxAreaPlan *d_plan; TbxAreaPlan plan; uint8_t *d_base_out[2]; TbxDigitPatch *d_patches[2]; int dw, dh, tx, ty;
  void *d_direct, *d_direct2; int direct_ok; /* tables of the direct kernel (tbx_direct.h), NULL / 0 when the pair is not covered */
};
static void drop_render_cache(struct tbx_pool *p);

struct tbx_pool {
  int game, n, n_pad, device;
  const tbx::GameInfo *info;
  tbx::Config cfg;
  void *d_cfg;
  std::vector<BrkTable> brk_tables;
  std::vector<AmiTable> ami_tables;
  void *d_tables;
  int d_tables_n;
  uint32_t *planes;
  unsigned long long *d_stats;
  int *d_bad;
  int32_t *d_legal;
  /* render resources of the current config: static frame (gray, RGBA) and per output size the INTER_AREA plan */
  uint8_t *d_base_gray[2], *d_base_rgba[2], *d_base_rgb[2];
  std::vector<uint32_t> h_base_rgba[2];
  std::vector<uint8_t> h_base_gray[2];
  std::map<std::pair<int, int>, struct AreaRes> area;
  int32_t *d_dense; /* [0] = count, [8..] = env ids the patch kernel left to the canvas kernel */
  int32_t *d_fb;    /* [0] = count, [1] = finished CTAs, [8..] = env ids the direct INTER_AREA kernel left to the tile kernel */
  /* JSON import / export staging (grown on demand, kept): env ids and AoS records on the device, records in pinned host memory */
  int32_t *j_ids; uint32_t *j_recs, *j_host; size_t j_cap_ids, j_cap_words;
  /* tbx_step_host staging */
  cudaStream_t hs;
  int32_t *h_actions_dev, *h_reward_dev, *h_score_dev, *h_lives_dev;
  uint8_t *h_done_dev, *h_obs_dev;
  size_t h_obs_cap;
};

/* The JSON codec is host work per env (build / walk a document tree, print / parse ~20-40 KB of text): spread the envs of a
 * batched call over host threads.  f(i) may throw; the first message is re-thrown on the calling thread. */
template <class F> static void parallel_for(int n, F f) {
  int nt = (int)std::thread::hardware_concurrency();
  if (const char *env = getenv("TBX_JSON_THREADS")) nt = atoi(env);
  nt = std::max(1, std::min(std::min(nt, 32), n / 8));
  if (nt <= 1) { for (int i = 0; i < n; i++) f(i); return; }
  std::vector<std::string> errs(nt);
  std::vector<std::thread> th;
  for (int t = 0; t < nt; t++)
    th.emplace_back([&, t]() {
      try { for (int i = t; i < n; i += nt) f(i); } catch (const std::exception &e) { errs[t] = e.what(); if (errs[t].empty()) errs[t] = "error"; }
    });
  for (auto &x : th) x.join();
  for (auto &e : errs) if (!e.empty()) throw std::runtime_error(e);
}

static size_t cfg_bytes(int game) { return game == TBX_BREAKOUT ? sizeof(BrkCfg) : game == TBX_AMIDAR ? sizeof(AmiCfg) : sizeof(SiCfg); }
static const void *cfg_ptr(const tbx_pool *p) {
  return p->game == TBX_BREAKOUT ? (const void *)&p->cfg.brk : p->game == TBX_AMIDAR ? (const void *)&p->cfg.ami : (const void *)&p->cfg.si;
}
static int blocks(int n, int t) { return (n + t - 1) / t; }

static int upload_cfg(tbx_pool *p) {
  CK(cudaMemcpy(p->d_cfg, cfg_ptr(p), cfg_bytes(p->game), cudaMemcpyHostToDevice));
  return TBX_OK;
}
static int upload_tables(tbx_pool *p) {
  size_t bytes = 0;
  const void *src = 0;
  int n = 0;
  if (p->game == TBX_BREAKOUT) { n = (int)p->brk_tables.size(); bytes = n * sizeof(BrkTable); src = p->brk_tables.data(); }
  else if (p->game == TBX_AMIDAR) { n = (int)p->ami_tables.size(); bytes = n * sizeof(AmiTable); src = p->ami_tables.data(); }
  if (n == 0) return TBX_OK;
  if (n == p->d_tables_n) return TBX_OK;
  CK(cudaDeviceSynchronize());
  void *d = 0;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice));
  if (p->d_tables) cudaFree(p->d_tables);
  p->d_tables = d;
  p->d_tables_n = n;
  return TBX_OK;
}
/* index of an identical table, appending it if new */
template <class TT> static int intern(std::vector<TT> &v, const TT &t) {
  for (size_t i = 0; i < v.size(); i++) if (memcmp(&v[i], &t, sizeof t) == 0) return (int)i;
  v.push_back(t);
  return (int)v.size() - 1;
}
static void install_default_table(tbx_pool *p) {
  if (p->game == TBX_BREAKOUT) { BrkTable t; tbx::brk_default_table(p->cfg.brk, t); tbx::brk_mark_delta_ok(p->cfg, t); p->cfg.brk.default_tbl = intern(p->brk_tables, t); }
  else if (p->game == TBX_AMIDAR) { AmiTable t; tbx::ami_default_table(p->cfg.ami, t); p->cfg.ami.default_tbl = intern(p->ami_tables, t); }
}

template <int GAME> static void launch_new_game(tbx_pool *p, const uint8_t *mask, cudaStream_t s) {
  new_game_kernel<GAME><<<blocks(p->n, 128), 128, 0, s>>>(p->planes, p->n, p->n_pad, p->d_cfg, p->d_tables, mask);
}
template <int GAME> static int launch_step(tbx_pool *p, const StepArgs &a, cudaStream_t s) {
  /* staged (records through shared memory) where it measures faster; TBX_STEP_STAGED=0/1 overrides (tuning, tests) */
  bool staged = GAME == TBX_SPACE_INVADERS; /* measured, 65,536 envs in steady state: Space Invaders 90 vs 178 us; Breakout 39 vs 37; Amidar 69 vs 61 */
  if (const char *env = getenv("TBX_STEP_STAGED")) staged = atoi(env) != 0;
  if (staged) {
    constexpr int EPB = StepGeom<GAME>::EPB, smem = Traits<GAME>::RW * EPB * 4;
    CK((cudaFuncSetAttribute(step_staged_kernel<GAME>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
    step_staged_kernel<GAME><<<blocks(p->n, EPB), EPB, smem, s>>>(a);
  } else {
    step_kernel<GAME><<<blocks(p->n, 128), 128, 0, s>>>(a);
  }
  CK(cudaGetLastError());
  return TBX_OK;
}
static int do_new_game(tbx_pool *p, const uint8_t *mask, cudaStream_t s) {
  if (p->game == TBX_BREAKOUT) launch_new_game<TBX_BREAKOUT>(p, mask, s);
  else if (p->game == TBX_AMIDAR) launch_new_game<TBX_AMIDAR>(p, mask, s);
  else launch_new_game<TBX_SPACE_INVADERS>(p, mask, s);
  CK(cudaGetLastError());
  return TBX_OK;
}

extern "C" {

const char *tbx_last_error(void) { return g_err.c_str(); }
int tbx_version(void) { return 100; }

int tbx_pool_destroy(tbx_pool *p) {
  if (!p) return TBX_OK;
  cudaSetDevice(p->device);
  cudaDeviceSynchronize();
  cudaFree(p->d_cfg); cudaFree(p->d_tables); cudaFree(p->planes); cudaFree(p->d_stats); cudaFree(p->d_bad); cudaFree(p->d_legal); cudaFree(p->d_dense); cudaFree(p->d_fb); cudaFree(p->j_ids); cudaFree(p->j_recs); cudaFreeHost(p->j_host);
  drop_render_cache(p);
  cudaFree(p->h_actions_dev); cudaFree(p->h_reward_dev); cudaFree(p->h_score_dev); cudaFree(p->h_lives_dev); cudaFree(p->h_done_dev); cudaFree(p->h_obs_dev);
  if (p->hs) cudaStreamDestroy(p->hs);
  delete p;
  return TBX_OK;
}

int tbx_pool_create(const char *game, int n_envs, int device, const char *cfg_json, tbx_pool **out) {
  if (!out) return set_err(TBX_EINVAL, "out is NULL");
  *out = 0;
  int g = tbx::game_from_name(game);
  if (g < 0) return set_err(TBX_EINVAL, std::string("unknown game '") + (game ? game : "(null)") + "'");
  if (n_envs < 1) return set_err(TBX_EINVAL, "n_envs must be >= 1");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
    return set_err(TBX_ECUDA, "no CUDA device available: toybox_b200 has no CPU path");
  if (device < 0 || device >= ndev) return set_err(TBX_EINVAL, "device index out of range");
  tbx_pool *p = new (std::nothrow) tbx_pool();
  if (!p) return set_err(TBX_ENOMEM, "out of host memory");
  p->game = g; p->n = n_envs; p->n_pad = (n_envs + 31) & ~31; p->device = device; p->info = tbx::game_info(g);
  p->d_cfg = p->d_tables = 0; p->d_base_gray[0] = p->d_base_gray[1] = p->d_base_rgba[0] = p->d_base_rgba[1] = p->d_base_rgb[0] = p->d_base_rgb[1] = 0; p->d_tables_n = 0; p->planes = 0; p->d_stats = 0; p->d_bad = 0; p->d_legal = 0;
  p->hs = 0; p->d_dense = 0; p->d_fb = 0; p->j_ids = 0; p->j_recs = p->j_host = 0; p->j_cap_ids = p->j_cap_words = 0; p->h_actions_dev = p->h_reward_dev = p->h_score_dev = p->h_lives_dev = 0; p->h_done_dev = p->h_obs_dev = 0; p->h_obs_cap = 0;
  int rc = TBX_OK;
  try {
    tbx::default_config(g, p->cfg);
    if (cfg_json) tbx::config_from_json(p->cfg, tbxjson::parse(cfg_json));
    install_default_table(p);
  } catch (const std::exception &e) { delete p; return set_err(TBX_EJSON, e.what()); }
  auto body = [&]() -> int {
    CK(cudaSetDevice(device));
    CK(cudaMalloc(&p->d_cfg, cfg_bytes(g)));
    CK(cudaMalloc(&p->planes, (size_t)p->info->rec_words * p->n_pad * sizeof(uint32_t)));
    CK(cudaMemset(p->planes, 0, (size_t)p->info->rec_words * p->n_pad * sizeof(uint32_t)));
    CK(cudaMalloc(&p->d_stats, 4 * sizeof(unsigned long long)));
    CK(cudaMemset(p->d_stats, 0, 4 * sizeof(unsigned long long)));
    CK(cudaMalloc(&p->d_bad, sizeof(int)));
    CK(cudaMemset(p->d_bad, 0, sizeof(int)));
    CK(cudaMalloc(&p->d_legal, 18 * sizeof(int32_t)));
    CK(cudaMemcpy(p->d_legal, p->info->legal, 18 * sizeof(int32_t), cudaMemcpyHostToDevice));
    int r = upload_cfg(p);
    if (r) return r;
    r = upload_tables(p);
    if (r) return r;
    const uint64_t *rand = g == TBX_BREAKOUT ? p->cfg.brk.rand : g == TBX_AMIDAR ? p->cfg.ami.rand : p->cfg.si.rand;
    set_sim_rand_kernel<<<blocks(p->n, 256), 256>>>(p->planes, p->n, p->n_pad, rand[0], rand[1]);
    CK(cudaGetLastError());
    for (int k = 0; k < p->info->new_games_at_ctor; k++) { r = do_new_game(p, 0, 0); if (r) return r; }
    CK(cudaDeviceSynchronize());
    return TBX_OK;
  };
  rc = body();
  if (rc) { std::string keep = g_err; tbx_pool_destroy(p); g_err = keep; return rc; }
  *out = p;
  return TBX_OK;
}

int tbx_frame_width(const tbx_pool *p) { return p ? p->info->width : 0; }
int tbx_frame_height(const tbx_pool *p) { return p ? p->info->height : 0; }
int tbx_n_envs(const tbx_pool *p) { return p ? p->n : 0; }
int tbx_device(const tbx_pool *p) { return p ? p->device : -1; }
int tbx_legal_actions(const tbx_pool *p, int32_t *out, int cap) {
  if (!p) return 0;
  for (int i = 0; i < p->info->n_legal && i < cap; i++) out[i] = p->info->legal[i];
  return p->info->n_legal;
}
size_t tbx_obs_bytes(const tbx_pool *p, int mode, int out_w, int out_h) {
  if (!p) return 0;
  size_t px = (size_t)p->info->width * p->info->height;
  switch (mode) {
    case TBX_OBS_RGBA: return px * 4;
    case TBX_OBS_RGB: return px * 3;
    case TBX_OBS_GRAY: return px;
    case TBX_OBS_GRAY_AREA:
      if (out_w < 1 || out_h < 1 || out_w > p->info->width || out_h > p->info->height || out_w > TBX_RS_MAX_DST || out_h > TBX_RS_MAX_DST) return 0;
      return (size_t)out_w * out_h;
  }
  return 0;
}

int tbx_seed(tbx_pool *p, const uint32_t *seeds, const int32_t *ids, int n) {
  if (!p || !seeds || n < 0) return set_err(TBX_EINVAL, "bad arguments");
  if (!ids && n > p->n) return set_err(TBX_EINVAL, "more seeds than envs");
  if (n == 0) return TBX_OK;
  if (ids) for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= p->n) return set_err(TBX_EINVAL, "env id out of range");
  CK(cudaSetDevice(p->device));
  uint32_t *d_seeds = 0;
  int32_t *d_ids = 0;
  CK(cudaMalloc(&d_seeds, n * sizeof(uint32_t)));
  CK(cudaMemcpy(d_seeds, seeds, n * sizeof(uint32_t), cudaMemcpyHostToDevice));
  if (ids) { CK(cudaMalloc(&d_ids, n * sizeof(int32_t))); CK(cudaMemcpy(d_ids, ids, n * sizeof(int32_t), cudaMemcpyHostToDevice)); }
  seed_kernel<<<blocks(n, 256), 256>>>(p->planes, p->n_pad, d_seeds, d_ids, n);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  cudaFree(d_seeds); cudaFree(d_ids);
  return TBX_OK;
}

int tbx_new_game(tbx_pool *p, const uint8_t *mask, void *stream) {
  if (!p) return set_err(TBX_EINVAL, "pool is NULL");
  CK(cudaSetDevice(p->device));
  return do_new_game(p, mask, (cudaStream_t)stream);
}

struct SynStream { uint64_t seed, env0, t; };
static int do_step(tbx_pool *p, const int32_t *actions, const uint8_t *inputs, int auto_reset, int32_t *reward, uint8_t *done,
                   int32_t *score, int32_t *lives, cudaStream_t s, const SynStream *syn = 0) {
  StepArgs a;
  a.planes = p->planes; a.n = p->n; a.n_pad = p->n_pad; a.cfg = p->d_cfg; a.tables = p->d_tables;
  a.actions = actions; a.inputs = inputs; a.auto_reset = auto_reset;
  a.legal = p->d_legal; a.n_legal = p->info->n_legal;
  a.syn_seed = syn ? syn->seed : 0; a.syn_env0 = syn ? syn->env0 : 0; a.syn_t = syn ? syn->t : 0;
  a.reward = reward; a.done = done; a.score = score; a.lives = lives; a.stats = p->d_stats; a.bad_actions = p->d_bad;
  if (p->game == TBX_BREAKOUT) return launch_step<TBX_BREAKOUT>(p, a, s);
  if (p->game == TBX_AMIDAR) return launch_step<TBX_AMIDAR>(p, a, s);
  return launch_step<TBX_SPACE_INVADERS>(p, a, s);
}
int tbx_step(tbx_pool *p, const int32_t *actions, int auto_reset, int32_t *reward, uint8_t *done, int32_t *score, int32_t *lives, void *stream) {
  if (!p || !actions) return set_err(TBX_EINVAL, "pool/actions is NULL");
  CK(cudaSetDevice(p->device));
  return do_step(p, actions, 0, auto_reset, reward, done, score, lives, (cudaStream_t)stream);
}
int tbx_step_random(tbx_pool *p, uint64_t seed, uint64_t env0, uint64_t t, int auto_reset, int32_t *reward, uint8_t *done, int32_t *score, int32_t *lives,
                    void *stream) {
  if (!p) return set_err(TBX_EINVAL, "pool is NULL");
  CK(cudaSetDevice(p->device));
  SynStream syn = {seed, env0, t};
  return do_step(p, 0, 0, auto_reset, reward, done, score, lives, (cudaStream_t)stream, &syn);
}
int tbx_step_inputs(tbx_pool *p, const uint8_t *inputs, int auto_reset, int32_t *reward, uint8_t *done, int32_t *score, int32_t *lives, void *stream) {
  if (!p || !inputs) return set_err(TBX_EINVAL, "pool/inputs is NULL");
  CK(cudaSetDevice(p->device));
  return do_step(p, 0, inputs, auto_reset, reward, done, score, lives, (cudaStream_t)stream);
}
int tbx_check(tbx_pool *p, void *stream) {
  if (!p) return set_err(TBX_EINVAL, "pool is NULL");
  CK(cudaSetDevice(p->device));
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, p->d_bad, sizeof bad, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CK(cudaStreamSynchronize((cudaStream_t)stream));
  if (bad) {
    CK(cudaMemsetAsync(p->d_bad, 0, sizeof(int), (cudaStream_t)stream));
    return set_err(TBX_EACTION, "Expected to apply action, but failed: " + std::to_string(bad) + " invalid ALE action id(s)");
  }
  return TBX_OK;
}

} /* extern "C" */

/* ---- render */
static void drop_render_cache(tbx_pool *p) {
  for (int b = 0; b < 2; b++) { cudaFree(p->d_base_gray[b]); cudaFree(p->d_base_rgba[b]); cudaFree(p->d_base_rgb[b]); p->d_base_gray[b] = p->d_base_rgba[b] = p->d_base_rgb[b] = 0; }
  for (auto &kv : p->area) { cudaFree(kv.second.d_plan); cudaFree(kv.second.d_base_out[0]); cudaFree(kv.second.d_base_out[1]); cudaFree(kv.second.d_patches[0]); cudaFree(kv.second.d_patches[1]); cudaFree(kv.second.d_direct); cudaFree(kv.second.d_direct2); }
  p->area.clear();
}
/* base frames 0/1 of the current config (see tbx_render.cuh), gray and RGBA, on the device */
static int ensure_base(tbx_pool *p) {
  if (p->d_base_gray[0]) return TBX_OK;
  const int npix = p->info->width * p->info->height;
  const BrkTable *brk_default = p->game == TBX_BREAKOUT ? &p->brk_tables[p->cfg.brk.default_tbl] : 0;
  for (int b = 0; b < 2; b++) {
    p->h_base_rgba[b].resize(npix);
    p->h_base_gray[b].resize(npix);
    tbx::build_base_frame(p->cfg, brk_default, b, p->h_base_rgba[b].data());
    tbx::frame_to_gray(p->h_base_rgba[b].data(), npix, p->h_base_gray[b].data());
    CK(cudaMalloc(&p->d_base_gray[b], npix + 16)); /* slack: the direct kernel reads whole words around a pixel */
    CK(cudaMemset(p->d_base_gray[b] + npix, 0, 16));
    CK(cudaMalloc(&p->d_base_rgba[b], (size_t)npix * 4));
    CK(cudaMemcpy(p->d_base_gray[b], p->h_base_gray[b].data(), npix, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_base_rgba[b], p->h_base_rgba[b].data(), (size_t)npix * 4, cudaMemcpyHostToDevice));
    std::vector<uint8_t> rgb((size_t)npix * 3);
    for (int i = 0; i < npix; i++) { uint32_t c = p->h_base_rgba[b][i]; rgb[3 * i] = c & 255; rgb[3 * i + 1] = (c >> 8) & 255; rgb[3 * i + 2] = (c >> 16) & 255; }
    CK(cudaMalloc(&p->d_base_rgb[b], rgb.size()));
    CK(cudaMemcpy(p->d_base_rgb[b], rgb.data(), rgb.size(), cudaMemcpyHostToDevice));
  }
  return TBX_OK;
}
static int ensure_area(tbx_pool *p, int out_w, int out_h, AreaRes **out) {
  auto key = std::make_pair(out_w, out_h);
  auto it = p->area.find(key);
  if (it == p->area.end()) {
    tbx::ResizeTab rs;
    TbxAreaPlan plan;
    try { tbx::build_resize(p->info->width, p->info->height, out_w, out_h, rs); }
    catch (const std::exception &e) { return set_err(TBX_EINVAL, e.what()); }
    if (!tbx::build_area_plan(rs, plan)) return set_err(TBX_EINVAL, "resize: the fused kernel supports destinations up to 128x128 with at most 8 taps per axis");
    AreaRes r;
    r.plan = plan;
    r.dw = out_w; r.dh = out_h; r.tx = plan.tx; r.ty = plan.ty; r.d_plan = 0; r.d_base_out[0] = r.d_base_out[1] = 0; r.d_patches[0] = r.d_patches[1] = 0;
    r.d_direct = r.d_direct2 = 0; r.direct_ok = 0;
    CK(cudaMalloc(&r.d_plan, sizeof plan));
    CK(cudaMemcpy(r.d_plan, &plan, sizeof plan, cudaMemcpyHostToDevice));
    for (int b = 0; b < 2; b++) {
      std::vector<uint8_t> base_out(((size_t)out_w * out_h + 15) & ~(size_t)15, 0);
      tbx::area_resize(p->h_base_gray[b].data(), rs, base_out.data());
      CK(cudaMalloc(&r.d_base_out[b], base_out.size()));
      CK(cudaMemcpy(r.d_base_out[b], base_out.data(), base_out.size(), cudaMemcpyHostToDevice));
      std::vector<TbxDigitPatch> patches((size_t)TBX_DP_SLOTS * 10);
      tbx::build_digit_patches(p->cfg, p->game == TBX_BREAKOUT ? &p->brk_tables[p->cfg.brk.default_tbl] : 0, rs, plan, p->h_base_gray[b].data(), patches.data());
      CK(cudaMalloc(&r.d_patches[b], patches.size() * sizeof(TbxDigitPatch)));
      CK(cudaMemcpy(r.d_patches[b], patches.data(), patches.size() * sizeof(TbxDigitPatch), cudaMemcpyHostToDevice));
    }
    CK(tbx_direct_build(p->cfg, p->game == TBX_BREAKOUT ? &p->brk_tables[p->cfg.brk.default_tbl] : 0, rs, plan, p->h_base_gray[0].data(), &r.d_direct, &r.d_direct2));
    r.direct_ok = r.d_direct != 0;
    it = p->area.insert(std::make_pair(key, r)).first;
  }
  *out = &it->second;
  return TBX_OK;
}

static int align16(int v) { return (v + 15) & ~15; }

template <int GAME, int MODE, int TX, int TY> static int launch_render(const RenderArgs &a, const void *cfg_host, const TbxAreaPlan *plan_host, int smem, cudaStream_t s) {
  /* the opt-in is per device (a process may hold pools on several GPUs): set it on every launch, a cheap host-side call */
  CK((cudaFuncSetAttribute(render_kernel<GAME, MODE, TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem > 160 * 1024 ? smem : 160 * 1024)));
  const int H = Traits<GAME>::H;
  dim3 grid(blocks(a.n, TBX_EPC), ((MODE == TBX_OBS_GRAY_AREA ? a.out_h : H) + a.band_rows - 1) / a.band_rows);
  /* measured: Breakout's few primitives keep 4 warps busy, the sprite-heavy games use 8 (profiles/r1_native_layouts.md) */
  int threads = (MODE == TBX_OBS_GRAY_AREA || GAME == TBX_BREAKOUT) ? 128 : 256;
  if (const char *env = getenv("TBX_RENDER_THREADS")) threads = atoi(env); /* tuning: 32, 64, 128 or 256 */
  if (threads != 32 && threads != 64 && threads != 128 && threads != 256) threads = 128;
  static const TbxAreaPlan no_plan = TbxAreaPlan();
  render_kernel<GAME, MODE, TX, TY><<<grid, threads, smem, s>>>(a, *(const typename Traits<GAME>::Cfg *)cfg_host, plan_host ? *plan_host : no_plan);
  CK(cudaGetLastError());
  return TBX_OK;
}
/* native layouts as broadcast + patch (tbx_render_native.cuh) */
template <int GAME, int PIX> static int launch_native_pix(const tbx_pool *p, const RenderArgs &a, const void *cfg_host, int band_smem, cudaStream_t s) {
  typedef typename Traits<GAME>::Cfg Cfg;
  const int threads = 256;
  const int patch_smem = a.smem_canvas; /* the records of the CTA's 8 envs */
  CK((cudaFuncSetAttribute(base_fill_kernel<GAME, PIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)));
  CK((cudaFuncSetAttribute(native_patch_kernel<GAME, PIX>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)));
  if (band_smem > 64 * 1024 || patch_smem > 160 * 1024) return set_err(TBX_EINVAL, "frame too wide for the native render path");
  const int H = Traits<GAME>::H;
  dim3 grid(blocks(a.n, TBX_FILL_ENVS), (H + a.band_rows - 1) / a.band_rows);
  base_fill_kernel<GAME, PIX><<<grid, 128, band_smem, s>>>(a, *(const Cfg *)cfg_host);
  CK(cudaGetLastError());
  native_patch_kernel<GAME, PIX><<<blocks(a.n, TBX_EPC), threads, patch_smem, s>>>(a, *(const Cfg *)cfg_host);
  CK(cudaGetLastError());
  (void)p;
  return TBX_OK;
}
template <int GAME> static int launch_native_game(const tbx_pool *p, int mode, const RenderArgs &a, const void *c, int band_smem, cudaStream_t s) {
  if (mode == TBX_OBS_RGBA) return launch_native_pix<GAME, 4>(p, a, c, band_smem, s);
  if (mode == TBX_OBS_RGB) return launch_native_pix<GAME, 3>(p, a, c, band_smem, s);
  return launch_native_pix<GAME, 1>(p, a, c, band_smem, s);
}
static int launch_native(const tbx_pool *p, int mode, const RenderArgs &a, int band_smem, cudaStream_t s) {
  if (p->game == TBX_BREAKOUT) return launch_native_game<TBX_BREAKOUT>(p, mode, a, cfg_ptr(p), band_smem, s);
  if (p->game == TBX_AMIDAR) return launch_native_game<TBX_AMIDAR>(p, mode, a, cfg_ptr(p), band_smem, s);
  return launch_native_game<TBX_SPACE_INVADERS>(p, mode, a, cfg_ptr(p), band_smem, s);
}

/* INTER_AREA, one warp per env (tbx_render_area.cuh) */
template <int GAME, int TX, int TY, bool DUAL> static int launch_area_tile_dual(const RenderArgs &a, const void *cfg_host, const TbxAreaPlan *plan_host, int smem, int threads, cudaStream_t s) {
  CK((cudaFuncSetAttribute(area_tile_kernel<GAME, TX, TY, DUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)));
  /* env-list mode (the envs the direct kernel handed over, normally none): a small grid walks the list */
  const int grid = a.env_list ? (blocks(a.n, TBX_EPC) < 296 ? blocks(a.n, TBX_EPC) : 296) : blocks(a.n, TBX_EPC);
  area_tile_kernel<GAME, TX, TY, DUAL><<<grid, threads, smem, s>>>(a, *(const typename Traits<GAME>::Cfg *)cfg_host, *plan_host);
  CK(cudaGetLastError());
  return TBX_OK;
}
template <int GAME, int TX, int TY> static int launch_area_tile(const RenderArgs &a, const void *cfg_host, const TbxAreaPlan *plan_host, int smem, int threads, cudaStream_t s) {
  if (a.planes2) return launch_area_tile_dual<GAME, TX, TY, true>(a, cfg_host, plan_host, smem, threads, s);
  return launch_area_tile_dual<GAME, TX, TY, false>(a, cfg_host, plan_host, smem, threads, s);
}
template <int GAME> static int launch_area_tile_taps(int tx, int ty, const RenderArgs &a, const void *c, const TbxAreaPlan *pl, int smem, int threads, cudaStream_t s) {
  if (ty <= 3) {
    if (tx <= 3) return launch_area_tile<GAME, 3, 3>(a, c, pl, smem, threads, s);
    if (tx <= 4) return launch_area_tile<GAME, 4, 3>(a, c, pl, smem, threads, s);
    return launch_area_tile<GAME, 5, 3>(a, c, pl, smem, threads, s);
  }
  if (tx <= 3) return launch_area_tile<GAME, 3, 4>(a, c, pl, smem, threads, s);
  if (tx <= 4) return launch_area_tile<GAME, 4, 4>(a, c, pl, smem, threads, s);
  return launch_area_tile<GAME, 5, 4>(a, c, pl, smem, threads, s);
}
/* INTER_AREA: the smallest instantiated tap counts that cover the plan */
template <int GAME, int TY> static int launch_area_tx(int tx, const RenderArgs &a, const void *c, const TbxAreaPlan *pl, int smem, cudaStream_t s) {
  if (tx <= 3) return launch_render<GAME, TBX_OBS_GRAY_AREA, 3, TY>(a, c, pl, smem, s);
  if (tx <= 4) return launch_render<GAME, TBX_OBS_GRAY_AREA, 4, TY>(a, c, pl, smem, s);
  return launch_render<GAME, TBX_OBS_GRAY_AREA, 5, TY>(a, c, pl, smem, s);
}
template <int GAME> static int launch_render_mode(int mode, int tx, int ty, const RenderArgs &a, const void *c, const TbxAreaPlan *pl, int smem, cudaStream_t s) {
  switch (mode) {
    case TBX_OBS_RGBA: return launch_render<GAME, TBX_OBS_RGBA, 1, 1>(a, c, pl, smem, s);
    case TBX_OBS_RGB: return launch_render<GAME, TBX_OBS_RGB, 1, 1>(a, c, pl, smem, s);
    case TBX_OBS_GRAY: return launch_render<GAME, TBX_OBS_GRAY, 1, 1>(a, c, pl, smem, s);
    default:
      if (tx > 5 || ty > 4) return launch_render<GAME, TBX_OBS_GRAY_AREA, 8, 8>(a, c, pl, smem, s);
      if (ty <= 3) return launch_area_tx<GAME, 3>(tx, a, c, pl, smem, s);
      return launch_area_tx<GAME, 4>(tx, a, c, pl, smem, s);
  }
}

extern "C" {

struct DualRender { const uint32_t *planes2; const uint8_t *reset_flags; int stack_k, stack_slot, stack_mode; };
static int render_impl(tbx_pool *p, uint8_t *dst, int mode, int out_w, int out_h, void *stream, const DualRender *dual);

int tbx_render(tbx_pool *p, uint8_t *dst, int mode, int out_w, int out_h, void *stream) { return render_impl(p, dst, mode, out_w, out_h, stream, 0); }

} /* extern "C" */

static int render_impl(tbx_pool *p, uint8_t *dst, int mode, int out_w, int out_h, void *stream, const DualRender *dual) {
  if (!p || !dst) return set_err(TBX_EINVAL, "pool/dst is NULL");
  if (((uintptr_t)dst & 15) != 0) return set_err(TBX_EINVAL, "dst must be 16-byte aligned");
  size_t fb = tbx_obs_bytes(p, mode, out_w, out_h);
  if (fb == 0) return set_err(TBX_EINVAL, "unsupported observation layout or size");
  CK(cudaSetDevice(p->device));
  int r = ensure_base(p);
  if (r) return r;
  const int W = p->info->width, H = p->info->height;
  const int pix = mode == TBX_OBS_RGBA ? 4 : mode == TBX_OBS_RGB ? 3 : 1; /* canvas bytes per pixel */
  RenderArgs a;
  a.planes = p->planes; a.n = p->n; a.n_pad = p->n_pad; a.cfg = p->d_cfg; a.tables = p->d_tables;
  a.planes2 = dual ? dual->planes2 : 0; a.reset_flags = dual ? dual->reset_flags : 0;
  a.stack_k = dual ? dual->stack_k : 1; a.stack_slot = dual ? dual->stack_slot : 0; a.stack_mode = dual ? dual->stack_mode : 0; a.env_stride = fb * (size_t)a.stack_k; a.tile_bytes = 0;
  a.patches[0] = a.patches[1] = 0; a.dense_list = 0; a.dense_count = 0; a.dense_flag = 0; a.dense_threshold = 0; a.env_list = 0; a.env_count = 0;
  a.dst = dst; a.frame_bytes = fb; a.plan = 0; a.out_h = out_h; a.tile_stride = 0; a.warp_bytes = 0; a.list_cap = 0; a.tile_hshift = 0; a.max_run = 0; a.band_rows = 0; a.smem_rects = 0;
  for (int b = 0; b < 2; b++) {
    a.base[b] = pix == 4 ? p->d_base_rgba[b] : pix == 3 ? p->d_base_rgb[b] : p->d_base_gray[b];
    a.base_out[b] = 0; /* INTER_AREA: set below */
  }
  a.smem_canvas = align16(p->info->rec_words * TBX_EPC * 4) * (dual ? 2 : 1);
  int smem_total, tx = 1, ty = 1;
  bool patch_first = false;
  const TbxAreaPlan *host_plan = 0;
  if (mode == TBX_OBS_GRAY_AREA) {
    AreaRes *ar = 0;
    r = ensure_area(p, out_w, out_h, &ar);
    if (r) return r;
    a.base_out[0] = ar->d_base_out[0]; a.base_out[1] = ar->d_base_out[1]; a.plan = ar->d_plan;
    {
      const char *env = getenv("TBX_AREA_DIGIT_CACHE");
      const bool on = !(env && atoi(env) == 0);
      a.patches[0] = on ? ar->d_patches[0] : 0; a.patches[1] = on ? ar->d_patches[1] : 0;
    }
    tx = ar->tx; ty = ar->ty;
    host_plan = &ar->plan;
    /* Bands of output rows: each (chunk, band) CTA keeps only the canvas rows that feed its output rows, which
     * multiplies the number of independent CTAs per SM.  Surplus (zero-weight) taps may read up to TY-1 rows past
     * the band's last real row: the canvas allocation covers them. */
    const int ty_inst = (tx > 5 || ty > 4) ? 8 : (ty <= 3 ? 3 : 4);
    /* default: one warp per env over 16x4 output tiles; plans with many taps (or TBX_AREA_KERNEL=cta) use the
     * CTA-wide canvas kernel below */
    const char *ksel = getenv("TBX_AREA_KERNEL");
    if (tx <= 5 && ty <= 4 && !(ksel && !strcmp(ksel, "cta"))) {
      const int tx_inst = tx <= 3 ? 3 : tx <= 4 ? 4 : 5;
      /* tiles of 16 x 8 output pixels, runs of at most 2 tiles: the scratch holds the widest / tallest run window */
      int ths = 3, max_run = 2;
      if (const char *env = getenv("TBX_AREA_TILE_H")) ths = atoi(env) == 4 ? 2 : atoi(env) == 16 ? 4 : 3;
      if (const char *env = getenv("TBX_AREA_MAX_RUN")) max_run = atoi(env);
      if (max_run < 1) max_run = 1;
      if (max_run > 8) max_run = 8;
      int stride = 0, rows = 0;
      for (int dxA = 0; dxA < out_w; dxA += 16) {
        const int dxL = dxA + 16 * max_run - 1 < out_w ? dxA + 16 * max_run - 1 : out_w - 1;
        const int wdt = ((ar->plan.xs0[dxL] + tx_inst + 3) & ~3) - (ar->plan.xs0[dxA] & ~3);
        if (wdt > stride) stride = wdt;
      }
      for (int dyA = 0; dyA < out_h; dyA += 1 << ths) {
        const int dyL = dyA + (1 << ths) - 1 < out_h ? dyA + (1 << ths) - 1 : out_h - 1;
        const int rws = ar->plan.ys0[dyL] + ty_inst - ar->plan.ys0[dyA];
        if (rws > rows) rows = rws;
      }
      if (stride <= 512 && rows <= 96 && W % 4 == 0) {
        int threads = 256;
        if (const char *env = getenv("TBX_AREA_THREADS")) threads = atoi(env);
        if (threads != 32 && threads != 64 && threads != 128 && threads != 256) threads = 256;
        a.tile_stride = stride;
        a.tile_hshift = ths;
        a.max_run = max_run;
        a.list_cap = TBX_TILE_LCAP;
        if (const char *env = getenv("TBX_AREA_LCAP")) a.list_cap = atoi(env); /* tests: small lists force the sweep / per-tile paths */
        if (a.list_cap < 1 || a.list_cap > TBX_TILE_LCAP) a.list_cap = TBX_TILE_LCAP;
        a.tile_bytes = align16(stride * rows);
        a.warp_bytes = TBX_TILE_LCAP * 20 + 32 + a.tile_bytes * (dual ? 2 : 1);
        const int smem = a.smem_canvas + (threads / 32) * a.warp_bytes;
        cudaStream_t s = (cudaStream_t)stream;
        /* Direct kernel first (tbx_render_direct.cuh; TBX_AREA_KERNEL=tile turns it off): closed forms for the envs they
         * cover, the rest are listed and rendered by the tile kernel in env-list mode right after (usually an empty list) */
        if (ar->direct_ok && !dual && !(ksel && !strcmp(ksel, "tile"))) {
          if (!p->d_fb) { CK(cudaMalloc(&p->d_fb, ((size_t)p->n_pad + 8) * sizeof(int32_t))); CK(cudaMemsetAsync(p->d_fb, 0, 8 * sizeof(int32_t), s)); }
          DirectArgs da;
          da.aux = ar->d_direct; da.aux2 = ar->d_direct2; da.fb_list = p->d_fb + 8; da.fb_count = p->d_fb; da.sched = p->d_fb + 2;
          tbx_direct_geometry(p->game, out_w, out_h, p->game == TBX_BREAKOUT ? p->cfg.brk.n_rows : 0, da);
          CK(tbx_launch_direct(p->game, tx, ty, a, cfg_ptr(p), ar->plan, da, s));
          if (p->game == TBX_SPACE_INVADERS) return TBX_OK; /* covers every env: nothing is handed over */
          a.env_list = p->d_fb + 8;
          a.env_count = p->d_fb;
        }
        if (p->game == TBX_BREAKOUT) return launch_area_tile_taps<TBX_BREAKOUT>(tx, ty, a, cfg_ptr(p), host_plan, smem, threads, s);
        if (p->game == TBX_AMIDAR) return launch_area_tile_taps<TBX_AMIDAR>(tx, ty, a, cfg_ptr(p), host_plan, smem, threads, s);
        return launch_area_tile_taps<TBX_SPACE_INVADERS>(tx, ty, a, cfg_ptr(p), host_plan, smem, threads, s);
      }
    }
    if (dual) return set_err(TBX_EINVAL, "wrapped observations need an output size the tile kernel supports (at most 5 x 4 taps)");
    int nb = p->game == TBX_AMIDAR ? 1 : 2;
    if (const char *env = getenv("TBX_AREA_BANDS")) nb = atoi(env);
    if (nb < 1) nb = 1;
    if (nb > out_h) nb = out_h;
    a.band_rows = (out_h + nb - 1) / nb;
    int canvas_rows = 0;
    for (int d0 = 0; d0 < out_h; d0 += a.band_rows) {
      int d1 = d0 + a.band_rows < out_h ? d0 + a.band_rows : out_h;
      int rows = ar->plan.ys0[d1 - 1] + ty_inst - ar->plan.ys0[d0];
      if (rows > canvas_rows) canvas_rows = rows;
    }
    a.smem_rects = a.smem_canvas + align16(canvas_rows * W + 16);
  } else {
    /* Default: broadcast + patch (tbx_render_native.cuh) -- the base frame leaves through the TMA engine at HBM speed, what
     * differs from it is painted straight into the frame.  Envs that differ in many places (a Breakout wall with more
     * than 20 holes, an Amidar maze with more than 48 painted tiles / boxes; TBX_NATIVE_DENSE overrides) are cheaper
     * to repaint in shared memory: the patch kernel lists them and the canvas kernel below renders exactly those.
     * TBX_NATIVE_KERNEL=canvas renders everything with the canvas kernel. */
    const char *ksel = getenv("TBX_NATIVE_KERNEL");
    patch_first = !(ksel && !strcmp(ksel, "canvas")) && !dual;
    /* canvas bands of at most ~40 KB so that five CTAs stay resident per SM */
    int max_rows = (40 * 1024) / (W * pix);
    if (max_rows < 1) max_rows = 1;
    int nb = (H + max_rows - 1) / max_rows;
    a.band_rows = (H + nb - 1) / nb;
    a.smem_rects = a.smem_canvas + align16(a.band_rows * W * pix);
  }
  smem_total = a.smem_rects + 2 * TBX_MAX_RECTS * (int)sizeof(int4) + 128 + TBX_MAX_BIG * (int)sizeof(uint4); /* + list counts, per-env base / env ids, the big-primitive queue */
  if (smem_total > 220 * 1024) return set_err(TBX_EINVAL, "observation size needs more shared memory than one SM has");
  cudaStream_t s = (cudaStream_t)stream;
  if (patch_first) {
    int thr = p->game == TBX_BREAKOUT ? 20 : p->game == TBX_AMIDAR ? 48 : INT32_MAX;
    if (const char *env = getenv("TBX_NATIVE_DENSE")) thr = atoi(env) < 0 ? INT32_MAX : atoi(env);
    RenderArgs f = a; /* same bands for the broadcast */
    if (thr != INT32_MAX) {
      if (!p->d_dense) CK(cudaMalloc(&p->d_dense, ((size_t)p->n_pad + 8) * sizeof(int32_t) + (size_t)p->n_pad));
      f.dense_list = p->d_dense + 8;
      f.dense_count = p->d_dense;
      f.dense_flag = reinterpret_cast<uint8_t *>(p->d_dense + 8 + p->n_pad);
      f.dense_threshold = thr;
      CK(cudaMemsetAsync(p->d_dense, 0, sizeof(int32_t), s));
      if (p->game == TBX_BREAKOUT) dense_classify_kernel<TBX_BREAKOUT><<<blocks(p->n, 8), 256, 0, s>>>(f, p->cfg.brk);
      else if (p->game == TBX_AMIDAR) dense_classify_kernel<TBX_AMIDAR><<<blocks(p->n, 8), 256, 0, s>>>(f, p->cfg.ami);
      else dense_classify_kernel<TBX_SPACE_INVADERS><<<blocks(p->n, 8), 256, 0, s>>>(f, p->cfg.si);
      CK(cudaGetLastError());
    }
    r = launch_native(p, mode, f, align16(a.band_rows * W * pix), s);
    if (r || thr == INT32_MAX) return r;
    a.env_list = p->d_dense + 8;
    a.env_count = p->d_dense;
  }
  if (p->game == TBX_BREAKOUT) return launch_render_mode<TBX_BREAKOUT>(mode, tx, ty, a, cfg_ptr(p), host_plan, smem_total, s);
  if (p->game == TBX_AMIDAR) return launch_render_mode<TBX_AMIDAR>(mode, tx, ty, a, cfg_ptr(p), host_plan, smem_total, s);
  return launch_render_mode<TBX_SPACE_INVADERS>(mode, tx, ty, a, cfg_ptr(p), host_plan, smem_total, s);
}

/* ---- the fused wrapper stack (tbx_wrap.cuh) */
struct tbx_wrap {
  tbx_pool *pool;
  int skip, noop_max, episodic_life, fire_reset, clip_rewards, stack_k, out_w, out_h;
  uint64_t noop_seed, env0;
  uint32_t *planes_prev, *wstate;
  uint8_t *was_reset;
  int head; /* ring slot of the newest frame */
  int stack_mode; /* what a reset observation does to the other k-1 ring slots: 0 = fills them (FrameStack.reset), 1 = zeroes them (VecFrameStack) */
  int32_t *ep_return, *ep_length; /* caller-owned device outputs of Monitor's episode record, or NULL */
};

extern "C" {

int tbx_wrap_create(tbx_pool *p, int skip, int noop_max, int episodic_life, int fire_reset, int clip_rewards, int stack_k, int out_w, int out_h,
                    uint64_t noop_seed, uint64_t env0, tbx_wrap **out) {
  if (!p || !out) return set_err(TBX_EINVAL, "pool/out is NULL");
  *out = 0;
  if (skip < 1 || skip > 64 || noop_max < 0 || noop_max > 1000 || stack_k < 1 || stack_k > 16) return set_err(TBX_EINVAL, "skip / noop_max / stack_k out of range");
  if (tbx_obs_bytes(p, TBX_OBS_GRAY_AREA, out_w, out_h) == 0) return set_err(TBX_EINVAL, "unsupported observation size");
  CK(cudaSetDevice(p->device));
  tbx_wrap *w = new (std::nothrow) tbx_wrap();
  if (!w) return set_err(TBX_ENOMEM, "out of memory");
  w->pool = p; w->skip = skip; w->noop_max = noop_max; w->episodic_life = episodic_life; w->fire_reset = fire_reset; w->clip_rewards = clip_rewards;
  w->stack_k = stack_k; w->out_w = out_w; w->out_h = out_h; w->noop_seed = noop_seed; w->env0 = env0; w->head = 0; w->stack_mode = 0; w->ep_return = w->ep_length = 0;
  w->planes_prev = 0; w->wstate = 0; w->was_reset = 0;
  const size_t plane_bytes = (size_t)p->info->rec_words * p->n_pad * 4;
  cudaError_t e = cudaMalloc(&w->planes_prev, plane_bytes);
  if (e == cudaSuccess) e = cudaMalloc(&w->wstate, 5 * (size_t)p->n_pad * 4);
  if (e == cudaSuccess) e = cudaMalloc(&w->was_reset, (size_t)p->n_pad);
  if (e == cudaSuccess) e = cudaMemcpy(w->planes_prev, p->planes, plane_bytes, cudaMemcpyDeviceToDevice);
  if (e == cudaSuccess) e = cudaMemset(w->wstate, 0, 5 * (size_t)p->n_pad * 4);
  if (e == cudaSuccess) { /* EpisodicLifeEnv.__init__: lives = 0, was_real_done = True (atari_wrappers.py:155-156) */
    std::vector<uint32_t> ones((size_t)p->n_pad, 1u);
    e = cudaMemcpy(w->wstate + p->n_pad, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice);
  }
  if (e == cudaSuccess) e = cudaMemset(w->was_reset, 0, (size_t)p->n_pad);
  if (e != cudaSuccess) {
    cudaFree(w->planes_prev); cudaFree(w->wstate); cudaFree(w->was_reset);
    delete w;
    return set_err(TBX_ECUDA, std::string("tbx_wrap_create: ") + cudaGetErrorString(e));
  }
  *out = w;
  return TBX_OK;
}

int tbx_wrap_set_stack_mode(tbx_wrap *w, int mode) {
  if (!w || (mode != 0 && mode != 1)) return set_err(TBX_EINVAL, "wrap is NULL or unknown stack mode (0 = FrameStack, 1 = VecFrameStack)");
  w->stack_mode = mode;
  return TBX_OK;
}
int tbx_wrap_set_episode_outputs(tbx_wrap *w, int32_t *ep_return_dev, int32_t *ep_length_dev) {
  if (!w) return set_err(TBX_EINVAL, "wrap is NULL");
  w->ep_return = ep_return_dev; w->ep_length = ep_length_dev;
  return TBX_OK;
}

int tbx_wrap_destroy(tbx_wrap *w) {
  if (!w) return TBX_OK;
  cudaSetDevice(w->pool->device);
  cudaDeviceSynchronize();
  cudaFree(w->planes_prev); cudaFree(w->wstate); cudaFree(w->was_reset);
  delete w;
  return TBX_OK;
}

int tbx_wrap_step(tbx_wrap *w, const int32_t *actions, uint8_t *obs_ring, int32_t *reward, uint8_t *done, uint8_t *real_done, int32_t *score,
                  int32_t *lives, int *slot_out, void *stream) {
  if (!w) return set_err(TBX_EINVAL, "wrap is NULL");
  tbx_pool *p = w->pool;
  CK(cudaSetDevice(p->device));
  cudaStream_t s = (cudaStream_t)stream;
  WrapArgs a;
  a.planes = p->planes; a.planes_prev = w->planes_prev; a.wstate = w->wstate; a.n = p->n; a.n_pad = p->n_pad; a.cfg = p->d_cfg; a.tables = p->d_tables;
  a.legal = p->d_legal; a.n_legal = p->info->n_legal; a.actions = actions;
  a.skip = w->skip; a.noop_max = w->noop_max; a.episodic_life = w->episodic_life; a.fire_reset = w->fire_reset; a.clip_rewards = w->clip_rewards;
  a.noop_seed = w->noop_seed; a.env0 = w->env0;
  a.reward = reward; a.score = score; a.lives = lives; a.done = done; a.real_done = real_done; a.was_reset = w->was_reset;
  a.ep_return = w->ep_return; a.ep_length = w->ep_length;
  a.stats = p->d_stats; a.bad_actions = p->d_bad;
  if (p->game == TBX_BREAKOUT) wrap_step_kernel<TBX_BREAKOUT><<<blocks(p->n, 128), 128, 0, s>>>(a);
  else if (p->game == TBX_AMIDAR) wrap_step_kernel<TBX_AMIDAR><<<blocks(p->n, 128), 128, 0, s>>>(a);
  else wrap_step_kernel<TBX_SPACE_INVADERS><<<blocks(p->n, 128), 128, 0, s>>>(a);
  CK(cudaGetLastError());
  if (obs_ring) {
    w->head = (w->head + 1) % w->stack_k;
    DualRender d;
    d.planes2 = w->planes_prev; d.reset_flags = w->was_reset; d.stack_k = w->stack_k; d.stack_slot = w->head; d.stack_mode = w->stack_mode;
    int r = render_impl(p, obs_ring, TBX_OBS_GRAY_AREA, w->out_w, w->out_h, stream, &d);
    if (r) return r;
  }
  if (slot_out) *slot_out = w->head;
  return TBX_OK;
}

int tbx_field_lookup(const char *game, const char *path, int *word, int *kind, int *bit) {
  int g = tbx::game_from_name(game);
  if (g < 0 || !path) return set_err(TBX_EINVAL, "unknown game or NULL path");
  tbxfields::Field f;
  if (!tbxfields::lookup(g, path, &f)) return set_err(TBX_EINVAL, std::string("no scalar state property '") + path + "' for " + game);
  if (word) *word = f.word;
  if (kind) *kind = f.kind;
  if (bit) *bit = f.bit;
  return TBX_OK;
}
static int field_of(tbx_pool *p, const char *path, tbxfields::Field *f) {
  if (!p || !path) return set_err(TBX_EINVAL, "pool/path is NULL");
  if (!tbxfields::lookup(p->game, path, f)) return set_err(TBX_EINVAL, std::string("no scalar state property '") + path + "' for " + p->info->name);
  return TBX_OK;
}
int tbx_field_get(tbx_pool *p, const char *path, void *out, void *stream) {
  tbxfields::Field f;
  int r = field_of(p, path, &f);
  if (r) return r;
  if (!out) return set_err(TBX_EINVAL, "out is NULL");
  CK(cudaSetDevice(p->device));
  field_get_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(p->planes, p->n, p->n_pad, f.word, f.kind, f.bit < 0 ? 0 : f.bit,
                                                                         (int32_t *)out, (double *)out);
  CK(cudaGetLastError());
  return TBX_OK;
}
int tbx_field_set(tbx_pool *p, const char *path, const void *values, const uint8_t *mask, void *stream) {
  tbxfields::Field f;
  int r = field_of(p, path, &f);
  if (r) return r;
  if (!values) return set_err(TBX_EINVAL, "values is NULL");
  CK(cudaSetDevice(p->device));
  const int also = (f.kind == TBX_F_I32 && f.word == TBX_FW(TbxHdr, score)) ? TBX_FW(TbxHdr, prev_score) : -1;
  field_set_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(p->planes, p->n, p->n_pad, f.word, f.kind, f.bit < 0 ? 0 : f.bit,
                                                                         (const int32_t *)values, (const double *)values, mask, also);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_breakout_columns(tbx_pool *p, int op, int col, const uint8_t *mask, int32_t *out, void *stream) {
  if (!p || p->game != TBX_BREAKOUT) return set_err(TBX_EINVAL, "tbx_breakout_columns needs a breakout pool");
  if (op < 0 || op > 2 || (op == 2 && !out)) return set_err(TBX_EINVAL, "op is 0 (remove column), 1 (fill column) or 2 (count channels into out)");
  CK(cudaSetDevice(p->device));
  brk_column_kernel<<<blocks(p->n, 128), 128, 0, (cudaStream_t)stream>>>(p->planes, p->n, p->n_pad, (const BrkTable *)p->d_tables, op, col, mask, out);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_read_scalars(tbx_pool *p, int32_t *score, int32_t *lives, int32_t *level, void *stream) {
  if (!p) return set_err(TBX_EINVAL, "pool is NULL");
  CK(cudaSetDevice(p->device));
  read_scalars_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(p->planes, p->n, p->n_pad, score, lives, level);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_fill_actions(tbx_pool *p, int32_t *actions, uint64_t seed, uint64_t env0, uint64_t t, void *stream) {
  if (!p || !actions) return set_err(TBX_EINVAL, "pool/actions is NULL");
  CK(cudaSetDevice(p->device));
  fill_actions_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(actions, p->n, seed, env0, t, p->d_legal, p->info->n_legal);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_fill_actions_at(tbx_pool *p, int32_t *actions, uint64_t seed, uint64_t env0, const uint64_t *t_dev, void *stream) {
  if (!p || !actions || !t_dev) return set_err(TBX_EINVAL, "pool/actions/t_dev is NULL");
  CK(cudaSetDevice(p->device));
  fill_actions_at_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(actions, p->n, seed, env0, t_dev, p->d_legal, p->info->n_legal);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_fill_actions_policy(tbx_pool *p, int32_t *actions, int policy, uint64_t t, void *stream) {
  if (!p || !actions) return set_err(TBX_EINVAL, "pool/actions is NULL");
  if (policy != 1 || p->game != TBX_BREAKOUT) return set_err(TBX_EINVAL, "only policy 1 (Breakout ball tracking) is implemented");
  CK(cudaSetDevice(p->device));
  breakout_tracking_actions_kernel<<<blocks(p->n, 256), 256, 0, (cudaStream_t)stream>>>(p->planes, p->n, p->n_pad, actions, t);
  CK(cudaGetLastError());
  return TBX_OK;
}

int tbx_stats_read(tbx_pool *p, int64_t *out, int reset, void *stream) {
  if (!p || !out) return set_err(TBX_EINVAL, "pool/out is NULL");
  CK(cudaSetDevice(p->device));
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(out, p->d_stats, 4 * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (reset) CK(cudaMemsetAsync(p->d_stats, 0, 4 * sizeof(int64_t), s));
  return TBX_OK;
}

int tbx_stats_read_device(tbx_pool *p, int64_t *out_dev, void *stream) {
  if (!p || !out_dev) return set_err(TBX_EINVAL, "pool/out is NULL");
  CK(cudaSetDevice(p->device));
  CK(cudaMemcpyAsync(out_dev, p->d_stats, 4 * sizeof(int64_t), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return TBX_OK;
}

int tbx_step_host(tbx_pool *p, const int32_t *actions, int auto_reset, int mode, int out_w, int out_h, uint8_t *obs, int32_t *reward,
                  uint8_t *done, int32_t *score, int32_t *lives) {
  if (!p || !actions) return set_err(TBX_EINVAL, "pool/actions is NULL");
  CK(cudaSetDevice(p->device));
  const size_t n = (size_t)p->n;
  if (!p->hs) {
    CK(cudaStreamCreateWithFlags(&p->hs, cudaStreamNonBlocking));
    CK(cudaMalloc(&p->h_actions_dev, n * 4)); CK(cudaMalloc(&p->h_reward_dev, n * 4)); CK(cudaMalloc(&p->h_score_dev, n * 4));
    CK(cudaMalloc(&p->h_lives_dev, n * 4)); CK(cudaMalloc(&p->h_done_dev, n));
  }
  size_t fb = 0;
  if (obs) {
    fb = tbx_obs_bytes(p, mode, out_w, out_h);
    if (fb == 0) return set_err(TBX_EINVAL, "unsupported observation layout or size");
    if (p->h_obs_cap < fb * n) {
      CK(cudaStreamSynchronize(p->hs));
      cudaFree(p->h_obs_dev);
      p->h_obs_dev = 0; p->h_obs_cap = 0;
      CK(cudaMalloc(&p->h_obs_dev, fb * n));
      p->h_obs_cap = fb * n;
    }
  }
  CK(cudaMemcpyAsync(p->h_actions_dev, actions, n * 4, cudaMemcpyHostToDevice, p->hs));
  int r = do_step(p, p->h_actions_dev, 0, auto_reset, p->h_reward_dev, p->h_done_dev, p->h_score_dev, p->h_lives_dev, p->hs);
  if (r) return r;
  if (reward) CK(cudaMemcpyAsync(reward, p->h_reward_dev, n * 4, cudaMemcpyDeviceToHost, p->hs));
  if (done) CK(cudaMemcpyAsync(done, p->h_done_dev, n, cudaMemcpyDeviceToHost, p->hs));
  if (score) CK(cudaMemcpyAsync(score, p->h_score_dev, n * 4, cudaMemcpyDeviceToHost, p->hs));
  if (lives) CK(cudaMemcpyAsync(lives, p->h_lives_dev, n * 4, cudaMemcpyDeviceToHost, p->hs));
  if (obs) {
    r = tbx_render(p, p->h_obs_dev, mode, out_w, out_h, p->hs);
    if (r) return r;
    CK(cudaMemcpyAsync(obs, p->h_obs_dev, fb * n, cudaMemcpyDeviceToHost, p->hs));
  }
  CK(cudaStreamSynchronize(p->hs));
  int bad = 0;
  CK(cudaMemcpy(&bad, p->d_bad, sizeof bad, cudaMemcpyDeviceToHost));
  if (bad) { cudaMemset(p->d_bad, 0, sizeof(int)); return set_err(TBX_EACTION, "Expected to apply action, but failed: invalid ALE action id"); }
  return TBX_OK;
}

/* ---- JSON */
static char *dup_str(const std::string &s) {
  char *out = (char *)malloc(s.size() + 1);
  if (out) memcpy(out, s.c_str(), s.size() + 1);
  return out;
}
void tbx_free_str(char *s) { free(s); }

static int ensure_json_staging(tbx_pool *p, int n) {
  const size_t words = (size_t)n * p->info->rec_words;
  if (p->j_cap_ids < (size_t)n) {
    cudaFree(p->j_ids); p->j_ids = 0; p->j_cap_ids = 0;
    const size_t cap = std::max<size_t>(1024, (size_t)n * 2);
    CK(cudaMalloc(&p->j_ids, cap * sizeof(int32_t)));
    p->j_cap_ids = cap;
  }
  if (p->j_cap_words < words) {
    cudaFree(p->j_recs); cudaFreeHost(p->j_host); p->j_recs = p->j_host = 0; p->j_cap_words = 0;
    const size_t cap = std::max<size_t>((size_t)1024 * p->info->rec_words, words * 2);
    CK(cudaMalloc(&p->j_recs, cap * 4));
    CK(cudaMallocHost(&p->j_host, cap * 4));
    p->j_cap_words = cap;
  }
  return TBX_OK;
}
/* records of the listed envs -> p->j_host (AoS).  Runs on the legacy default stream, which is ordered after the work of
 * every blocking stream; callers on non-blocking streams synchronise first (the Python layer does). */
static int fetch_records(tbx_pool *p, const int32_t *ids, int n) {
  const int rw = p->info->rec_words;
  for (int i = 0; i < n; i++) if (ids[i] < 0 || ids[i] >= p->n) return set_err(TBX_EINVAL, "env id out of range");
  CK(cudaSetDevice(p->device));
  int r = ensure_json_staging(p, n);
  if (r) return r;
  if (p->hs) CK(cudaStreamSynchronize(p->hs));
  CK(cudaMemcpyAsync(p->j_ids, ids, n * sizeof(int32_t), cudaMemcpyHostToDevice, 0));
  gather_kernel<<<blocks(n * rw, 256), 256>>>(p->planes, p->n_pad, p->j_ids, n, rw, p->j_recs);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(p->j_host, p->j_recs, (size_t)n * rw * 4, cudaMemcpyDeviceToHost, 0));
  CK(cudaStreamSynchronize(0));
  return TBX_OK;
}
/* p->j_host (AoS) -> the listed envs; the ids are still on the device from fetch_records */
static int store_records(tbx_pool *p, int n) {
  const int rw = p->info->rec_words;
  CK(cudaMemcpyAsync(p->j_recs, p->j_host, (size_t)n * rw * 4, cudaMemcpyHostToDevice, 0));
  scatter_kernel<<<blocks(n * rw, 256), 256>>>(p->planes, p->n_pad, p->j_ids, n, rw, p->j_recs);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(0));
  return TBX_OK;
}

int tbx_state_to_json(tbx_pool *p, const int32_t *ids, int n, char **out) {
  if (!p || !ids || !out || n < 0) return set_err(TBX_EINVAL, "bad arguments");
  for (int i = 0; i < n; i++) out[i] = 0;
  if (n == 0) return TBX_OK;
  int r = fetch_records(p, ids, n);
  if (r) return r;
  const int rw = p->info->rec_words;
  const uint32_t *recs = p->j_host;
  try {
    parallel_for(n, [&](int i) {
      const uint32_t *R = recs + (size_t)i * rw;
      Value v;
      if (p->game == TBX_BREAKOUT) {
        const BrkRec &rec = *reinterpret_cast<const BrkRec *>(R);
        if (rec.hdr.tbl < 0 || rec.hdr.tbl >= (int)p->brk_tables.size()) throw std::runtime_error("corrupt table index");
        v = tbx::brk_state_to_json(rec, p->brk_tables[rec.hdr.tbl]);
      } else if (p->game == TBX_AMIDAR) {
        const AmiRec &rec = *reinterpret_cast<const AmiRec *>(R);
        if (rec.hdr.tbl < 0 || rec.hdr.tbl >= (int)p->ami_tables.size()) throw std::runtime_error("corrupt table index");
        v = tbx::ami_state_to_json(rec, p->ami_tables[rec.hdr.tbl]);
      } else v = tbx::si_state_to_json(*reinterpret_cast<const SiRec *>(R));
      out[i] = dup_str(tbxjson::dump(v));
      if (!out[i]) throw std::runtime_error("out of host memory");
    });
  } catch (const std::exception &e) {
    for (int i = 0; i < n; i++) { free(out[i]); out[i] = 0; }
    return set_err(TBX_EJSON, e.what());
  }
  return TBX_OK;
}

int tbx_state_from_json(tbx_pool *p, const int32_t *ids, int n, const char *const *json) {
  if (!p || !ids || !json || n < 0) return set_err(TBX_EINVAL, "bad arguments");
  if (n == 0) return TBX_OK;
  int r = fetch_records(p, ids, n); /* keeps the fields JSON does not carry (simulator rand, episode counters) */
  if (r) return r;
  const int rw = p->info->rec_words;
  uint32_t *recs = p->j_host;
  /* parse everything (in parallel) before touching the pool so a bad document leaves it unchanged */
  std::vector<BrkTable> new_brk(p->game == TBX_BREAKOUT ? n : 0);
  std::vector<AmiTable> new_ami(p->game == TBX_AMIDAR ? n : 0);
  try {
    parallel_for(n, [&](int i) {
      uint32_t *R = recs + (size_t)i * rw;
      Value v = tbxjson::parse(json[i]);
      if (p->game == TBX_BREAKOUT) { tbx::brk_state_from_json(v, *reinterpret_cast<BrkRec *>(R), new_brk[i]); tbx::brk_mark_delta_ok(p->cfg, new_brk[i]); }
      else if (p->game == TBX_AMIDAR) tbx::ami_state_from_json(v, *reinterpret_cast<AmiRec *>(R), new_ami[i]);
      else tbx::si_state_from_json(v, *reinterpret_cast<SiRec *>(R));
    });
  } catch (const std::exception &e) { return set_err(TBX_EJSON, e.what()); }
  for (int i = 0; i < n; i++) {
    uint32_t *R = recs + (size_t)i * rw;
    if (p->game == TBX_BREAKOUT) reinterpret_cast<BrkRec *>(R)->hdr.tbl = intern(p->brk_tables, new_brk[i]);
    else if (p->game == TBX_AMIDAR) reinterpret_cast<AmiRec *>(R)->hdr.tbl = intern(p->ami_tables, new_ami[i]);
  }
  r = upload_tables(p);
  if (r) return r;
  return store_records(p, n);
}

static uint64_t *cfg_rand(tbx::Config &c) { return c.game == TBX_BREAKOUT ? c.brk.rand : c.game == TBX_AMIDAR ? c.ami.rand : c.si.rand; }
int tbx_config_to_json(tbx_pool *p, char **out) {
  if (!p || !out) return set_err(TBX_EINVAL, "bad arguments");
  /* ctoybox's config_to_json shows the simulator as it is NOW: `rand` is the simulator rng that new_game has advanced.
   * A pool keeps one simulator rng per env; the pool-level document reports env 0's (the batch-1 shim's only env). */
  tbx::Config c = p->cfg;
  CK(cudaSetDevice(p->device));
  uint32_t w[4];
  CK(cudaMemcpy2D(w, sizeof(uint32_t), p->planes + (size_t)TBX_HW(sim_rand) * p->n_pad, (size_t)p->n_pad * sizeof(uint32_t), sizeof(uint32_t), 4,
                  cudaMemcpyDeviceToHost));
  cfg_rand(c)[0] = (uint64_t)w[0] | ((uint64_t)w[1] << 32);
  cfg_rand(c)[1] = (uint64_t)w[2] | ((uint64_t)w[3] << 32);
  try { *out = dup_str(tbxjson::dump(tbx::config_to_json(c))); } catch (const std::exception &e) { return set_err(TBX_EJSON, e.what()); }
  return *out ? TBX_OK : set_err(TBX_ENOMEM, "out of host memory");
}
int tbx_config_from_json(tbx_pool *p, const char *json) {
  if (!p || !json) return set_err(TBX_EINVAL, "bad arguments");
  tbx::Config c = p->cfg;
  try { tbx::config_from_json(c, tbxjson::parse(json)); } catch (const std::exception &e) { return set_err(TBX_EJSON, e.what()); }
  CK(cudaSetDevice(p->device));
  CK(cudaDeviceSynchronize());
  p->cfg = c;
  drop_render_cache(p);
  install_default_table(p);
  int r = upload_tables(p);
  if (r) return r;
  /* write_config_json replaces the simulator, its rng included: every env's simulator rng restarts from the document's */
  set_sim_rand_kernel<<<blocks(p->n, 256), 256>>>(p->planes, p->n, p->n_pad, cfg_rand(p->cfg)[0], cfg_rand(p->cfg)[1]);
  CK(cudaGetLastError());
  return upload_cfg(p);
}
int tbx_schema_for_state(const char *game, char **out) {
  int g = tbx::game_from_name(game);
  if (g < 0 || !out) return set_err(TBX_EINVAL, "unknown game");
  *out = dup_str(tbxjson::dump(tbx::schema_for_state(g)));
  return *out ? TBX_OK : set_err(TBX_ENOMEM, "out of host memory");
}
int tbx_schema_for_config(const char *game, char **out) {
  int g = tbx::game_from_name(game);
  if (g < 0 || !out) return set_err(TBX_EINVAL, "unknown game");
  try { *out = dup_str(tbxjson::dump(tbx::schema_for_config(g))); } catch (const std::exception &e) { return set_err(TBX_EJSON, e.what()); }
  return *out ? TBX_OK : set_err(TBX_ENOMEM, "out of host memory");
}
int tbx_query_json(tbx_pool *p, int env, const char *query, const char *args_json, char **out) {
  if (!p || !query || !out) return set_err(TBX_EINVAL, "bad arguments");
  *out = 0;
  if (env < 0 || env >= p->n) return set_err(TBX_EINVAL, "env id out of range");
  try {
    Value args = tbxjson::parse(args_json ? args_json : "null");
    std::string q(query);
    const uint32_t *rec = 0;
    const BrkTable *brk = 0;
    if (p->game == TBX_BREAKOUT) {
      int32_t id = env;
      int r = fetch_records(p, &id, 1);
      if (r) return r;
      rec = p->j_host;
      int tbl = reinterpret_cast<const BrkRec *>(rec)->hdr.tbl;
      if (tbl < 0 || tbl >= (int)p->brk_tables.size()) return set_err(TBX_EINVAL, "corrupt table index");
      brk = &p->brk_tables[tbl];
    }
    Value res = tbx::query_json(p->game, rec, brk, q, args);
    *out = dup_str(tbxjson::dump(res));
  } catch (const std::exception &e) { return set_err(TBX_EJSON, e.what()); }
  return *out ? TBX_OK : set_err(TBX_ENOMEM, "out of host memory");
}

} /* extern "C" */
