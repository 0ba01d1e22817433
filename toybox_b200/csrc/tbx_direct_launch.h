/* tbx_direct_launch.h -- entry points of tbx_direct.cu for tbx_pool.cu */
#ifndef TBX_DIRECT_LAUNCH_H
#define TBX_DIRECT_LAUNCH_H
#include "tbx_render_direct.cuh"

/* Breakout, direct INTER_AREA kernel instantiated for at least tx x ty taps (tx <= 5, ty <= 4) */
cudaError_t tbx_launch_brk_direct(int tx, int ty, const tbxk::RenderArgs &a, const BrkCfg &cfg, const TbxAreaPlan &plan, const tbxk::DirectArgs &d, int smem,
                                  cudaStream_t s);
#endif
