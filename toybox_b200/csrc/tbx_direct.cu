/* tbx_direct.cu -- launchers of the direct INTER_AREA kernels (tbx_render_direct.cuh); a translation unit of its own so
 * that it compiles in parallel with tbx_pool.cu. */
#include "tbx_render_direct.cuh"
#include "tbx_direct_launch.h"

using namespace tbxk;

template <int TX, int TY> static cudaError_t launch_brk(const RenderArgs &a, const BrkCfg &cfg, const TbxAreaPlan &plan, const DirectArgs &d, int smem, cudaStream_t s) {
  /* the attribute is per device: set it on every launch (a cheap host-side call) rather than caching it per process */
  cudaError_t e = cudaFuncSetAttribute(brk_direct_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  brk_direct_kernel<TX, TY><<<(a.n + TBX_EPC - 1) / TBX_EPC, TBX_DIRECT_THREADS, smem, s>>>(a, cfg, plan, d);
  return cudaGetLastError();
}

cudaError_t tbx_launch_brk_direct(int tx, int ty, const RenderArgs &a, const BrkCfg &cfg, const TbxAreaPlan &plan, const DirectArgs &d, int smem, cudaStream_t s) {
  if (ty <= 3) {
    if (tx <= 3) return launch_brk<3, 3>(a, cfg, plan, d, smem, s);
    if (tx <= 4) return launch_brk<4, 3>(a, cfg, plan, d, smem, s);
    return launch_brk<5, 3>(a, cfg, plan, d, smem, s);
  }
  if (tx <= 3) return launch_brk<3, 4>(a, cfg, plan, d, smem, s);
  if (tx <= 4) return launch_brk<4, 4>(a, cfg, plan, d, smem, s);
  return launch_brk<5, 4>(a, cfg, plan, d, smem, s);
}
