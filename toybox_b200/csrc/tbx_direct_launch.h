/* tbx_direct_launch.h -- entry points of tbx_direct.cu (the direct INTER_AREA kernels, tbx_render_direct.cuh) for
 * tbx_pool.cu.  Kept free of the kernel headers so that the two translation units rebuild independently. */
#ifndef TBX_DIRECT_LAUNCH_H
#define TBX_DIRECT_LAUNCH_H
#include "tbx_render.cuh"
#include "tbx_host.h"

namespace tbxk {
struct DirectArgs {
  const void *aux;    /* the game's closed-form tables on the device (TbxBrkDirect, TbxSiDirect) */
  const void *aux2;   /* Space Invaders: the sprite patch tables (TbxSpritePatch[]) */
  int32_t *fb_list;   /* envs handed to the tile kernel */
  int *fb_count;
  int *sched;         /* [2]: next chunk id - grid size, finished CTAs (dynamic chunk scheduling of the persistent kernels) */
  int hstride;        /* floats per H row in shared memory */
  int warp_bytes;     /* shared memory per warp */
  int smem_base;      /* bytes of the staged base-frame down-sample at the start of shared memory */
  int smem_total;
};
}

/* closed-form tables of one (config, output size) pair on the device; *d_aux stays NULL when the pair is not covered */
cudaError_t tbx_direct_build(const tbx::Config &c, const BrkTable *brk_default, const tbx::ResizeTab &rs, const TbxAreaPlan &plan, const uint8_t *base0_gray,
                             void **d_aux, void **d_aux2);
/* shared memory per warp / floats per H row for an output width */
void tbx_direct_geometry(int game, int out_w, int out_h, int brk_rows, tbxk::DirectArgs &d); /* brk_rows: Breakout's brick rows (config) */
/* direct INTER_AREA kernel instantiated for at least tx x ty taps (tx <= 5, ty <= 4) */
cudaError_t tbx_launch_direct(int game, int tx, int ty, const tbxk::RenderArgs &a, const void *cfg_host, const TbxAreaPlan &plan, const tbxk::DirectArgs &d,
                              cudaStream_t s);
#endif
