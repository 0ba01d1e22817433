/* tbx_direct.cu -- host side of the direct INTER_AREA kernels (tbx_render_direct.cuh): table upload and launchers.  A
 * translation unit of its own so that it compiles in parallel with, and rebuilds independently of, tbx_pool.cu. */
#include "tbx_render_direct.cuh"
#include <stdlib.h>
#include <vector>

using namespace tbxk;

static int align16(int v) { return (v + 15) & ~15; }

cudaError_t tbx_direct_build(const tbx::Config &c, const BrkTable *brk_default, const tbx::ResizeTab &rs, const TbxAreaPlan &plan, const uint8_t *base0_gray,
                             void **d_aux) {
  *d_aux = 0;
  if (c.game == TBX_BREAKOUT && brk_default) {
    std::vector<TbxBrkDirect> aux(1);
    tbx::build_brk_direct(c, *brk_default, rs, plan, base0_gray, aux[0]);
    if (!aux[0].ok) return cudaSuccess;
    cudaError_t e = cudaMalloc(d_aux, sizeof(TbxBrkDirect));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_aux, aux.data(), sizeof(TbxBrkDirect), cudaMemcpyHostToDevice);
  }
  return cudaSuccess;
}

void tbx_direct_geometry(int game, int out_w, int out_h, DirectArgs &d) {
  d.hstride = (out_w + 3) & ~3;
  d.smem_base = align16(out_w * out_h);
  /* per warp: its env's record, the wall's H rows, the movers' records */
  d.warp_bytes = game == TBX_BREAKOUT ? align16(TBX_WORDS(BrkRec) * 4) + align16(TBX_BRK_MAX_ROWS * d.hstride * (int)sizeof(float)) + 256 : 0;
  d.smem_total = d.smem_base + 2 * TBX_WORDS(BrkRec) * TBX_EPC * 4 + (TBX_DIRECT_THREADS / 32) * d.warp_bytes;
}

template <int TX, int TY> static cudaError_t launch_brk(const RenderArgs &a, const BrkCfg &cfg, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  /* the attribute is per device: set it on every launch (a cheap host-side call) rather than caching it per process */
  cudaError_t e = cudaFuncSetAttribute(brk_direct_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, d.smem_total);
  if (e != cudaSuccess) return e;
  /* persistent grid: as many CTAs as the device keeps resident (cached per device), each walks its share of the chunks */
  static int resident[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!resident[dev]) {
    int per_sm = 0, sms = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, brk_direct_kernel<TX, TY>, TBX_DIRECT_THREADS, d.smem_total);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident[dev] = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
  }
  const int n_chunks = (a.n + TBX_EPC - 1) / TBX_EPC;
  int grid = n_chunks < resident[dev] ? n_chunks : resident[dev];
  if (const char *env = getenv("TBX_DIRECT_GRID")) if (atoi(env) > 0) grid = atoi(env) < n_chunks ? atoi(env) : n_chunks; /* tuning / tests */
  brk_direct_kernel<TX, TY><<<grid, TBX_DIRECT_THREADS, d.smem_total, s>>>(a, cfg, plan, d);
  return cudaGetLastError();
}

cudaError_t tbx_launch_direct(int game, int tx, int ty, const RenderArgs &a, const void *cfg_host, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  if (game != TBX_BREAKOUT) return cudaErrorInvalidValue;
  const BrkCfg &cfg = *(const BrkCfg *)cfg_host;
  if (ty <= 3) {
    if (tx <= 3) return launch_brk<3, 3>(a, cfg, plan, d, s);
    if (tx <= 4) return launch_brk<4, 3>(a, cfg, plan, d, s);
    return launch_brk<5, 3>(a, cfg, plan, d, s);
  }
  if (tx <= 3) return launch_brk<3, 4>(a, cfg, plan, d, s);
  if (tx <= 4) return launch_brk<4, 4>(a, cfg, plan, d, s);
  return launch_brk<5, 4>(a, cfg, plan, d, s);
}
