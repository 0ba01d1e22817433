/* tbx_direct.cu -- host side of the direct INTER_AREA kernels (tbx_render_direct.cuh): table upload and launchers.  A
 * translation unit of its own so that it compiles in parallel with, and rebuilds independently of, tbx_pool.cu. */
#include <mutex>
#include "tbx_render_direct.cuh"
#include <stdlib.h>
#include <vector>

using namespace tbxk;

static int align16(int v) { return (v + 15) & ~15; }

cudaError_t tbx_direct_build(const tbx::Config &c, const BrkTable *brk_default, const tbx::ResizeTab &rs, const TbxAreaPlan &plan, const uint8_t *base0_gray,
                             void **d_aux, void **d_aux2) {
  *d_aux = 0;
  *d_aux2 = 0;
  cudaError_t e = cudaSuccess;
  if (c.game == TBX_BREAKOUT && brk_default) {
    std::vector<TbxBrkDirect> aux(1);
    tbx::build_brk_direct(c, *brk_default, rs, plan, base0_gray, aux[0]);
    if (!aux[0].ok) return cudaSuccess;
    e = cudaMalloc(d_aux, sizeof(TbxBrkDirect));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_aux, aux.data(), sizeof(TbxBrkDirect), cudaMemcpyHostToDevice);
  }
  if (c.game == TBX_AMIDAR) {
    std::vector<TbxAmiDirect> aux(1);
    tbx::build_ami_direct(c, rs, plan, aux[0]);
    if (!aux[0].ok) return cudaSuccess;
    e = cudaMalloc(d_aux, sizeof(TbxAmiDirect));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_aux, aux.data(), sizeof(TbxAmiDirect), cudaMemcpyHostToDevice);
  }
  if (c.game == TBX_SPACE_INVADERS) {
    std::vector<TbxSiDirect> aux(1);
    std::vector<TbxSpritePatch> patches;
    tbx::build_si_direct(c, rs, plan, base0_gray, aux[0], patches);
    if (!aux[0].ok) return cudaSuccess;
    e = cudaMalloc(d_aux, sizeof(TbxSiDirect));
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(*d_aux, aux.data(), sizeof(TbxSiDirect), cudaMemcpyHostToDevice);
    if (e != cudaSuccess || patches.empty()) return e;
    e = cudaMalloc(d_aux2, patches.size() * sizeof(TbxSpritePatch));
    if (e != cudaSuccess) return e;
    return cudaMemcpy(*d_aux2, patches.data(), patches.size() * sizeof(TbxSpritePatch), cudaMemcpyHostToDevice);
  }
  return cudaSuccess;
}

void tbx_direct_geometry(int game, int out_w, int out_h, int brk_rows, DirectArgs &d) {
  d.hstride = (out_w + 3) & ~3;
  d.smem_base = align16(out_w * out_h);
  if (game == TBX_BREAKOUT) {
    if (brk_rows < 1 || brk_rows > TBX_BRK_MAX_ROWS) brk_rows = TBX_BRK_MAX_ROWS;
    /* the kernel's small tables follow the staged frame: digit patches, base-frame row classes, division table, the wall's H look-up */
    d.smem_base += TBX_BRK_TAB_BYTES + align16(brk_rows * 4 * d.hstride * (int)sizeof(float));
    /* per warp: its env's record, the wall's H rows, the movers' records; two record stages */
    d.warp_bytes = align16(TBX_WORDS(BrkRec) * 4) + align16(brk_rows * d.hstride * (int)sizeof(float)) + 256;
    d.smem_total = d.smem_base + 2 * TBX_WORDS(BrkRec) * TBX_EPC * 4 + (TBX_DIRECT_THREADS / 32) * d.warp_bytes;
  } else if (game == TBX_AMIDAR) {
    /* per warp: its env's record, the tile rows as looks, the movers' records; two record stages */
    d.warp_bytes = align16(TBX_WORDS(AmiRec) * 4) + 256 + 448;
    d.smem_total = d.smem_base + 2 * TBX_WORDS(AmiRec) * TBX_EPC * 4 + (TBX_DIRECT_THREADS / 32) * d.warp_bytes;
  } else {
    /* per warp: its env's record, the entries, two coverage bitmaps, the evaluated-entry list; one record stage */
    const int ow = (out_w + 31) / 32;
    d.warp_bytes = align16(TBX_WORDS(SiRec) * 4) + TBX_SD_MAX_ENTRIES * 32 + align16(2 * out_h * ow * 4) + TBX_SD_MAX_ENTRIES * 8;
    d.smem_total = d.smem_base + TBX_WORDS(SiRec) * TBX_EPC * 4 + (TBX_DIRECT_THREADS / 32) * d.warp_bytes;
  }
}

/* persistent grid: as many CTAs as the device keeps resident, each walks its share of the chunks.  The residency is cached per
 * (kernel, shared memory, device): the kernels share function-pointer types, so the cache is keyed by the pointer's value */
template <class K> static cudaError_t persistent_grid(K kernel, int smem, int n_chunks, int *grid_out) {
  struct Entry { const void *k; int smem, dev, resident; };
  static Entry cache[64];
  static int n_cache = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  int resident = 0;
  {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < n_cache; i++)
      if (cache[i].k == (const void *)kernel && cache[i].smem == smem && cache[i].dev == dev) resident = cache[i].resident;
  }
  if (!resident) {
    int per_sm = 0, sms = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TBX_DIRECT_THREADS, smem);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    resident = (per_sm > 0 ? per_sm : 1) * (sms > 0 ? sms : 148);
    std::lock_guard<std::mutex> lock(mu);
    if (n_cache < 64) cache[n_cache++] = Entry{(const void *)kernel, smem, dev, resident};
  }
  int grid = n_chunks < resident ? n_chunks : resident;
  if (const char *env = getenv("TBX_DIRECT_GRID")) if (atoi(env) > 0) grid = atoi(env) < n_chunks ? atoi(env) : n_chunks; /* tuning / tests */
  *grid_out = grid;
  return cudaSuccess;
}

template <int TX, int TY> static cudaError_t launch_brk(const RenderArgs &a, const BrkCfg &cfg, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  /* the attribute is per device: set it on every launch (a cheap host-side call) rather than caching it per process */
  const int smem = d.smem_total + brk_plan_smem_bytes(TX, TY, plan.dw, plan.dh); /* the plan's per-pixel tables go behind the rest */
  cudaError_t e = cudaFuncSetAttribute(brk_direct_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = 1;
  e = persistent_grid(brk_direct_kernel<TX, TY>, smem, (a.n + TBX_EPC - 1) / TBX_EPC, &grid);
  if (e != cudaSuccess) return e;
  brk_direct_kernel<TX, TY><<<grid, TBX_DIRECT_THREADS, smem, s>>>(a, cfg, plan, d);
  return cudaGetLastError();
}
template <int TX, int TY> static cudaError_t launch_si(const RenderArgs &a, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  const int smem = d.smem_total + si_tab_smem_bytes(plan.dw, plan.dh); /* the plain-background map and the division table go behind the rest */
  cudaError_t e = cudaFuncSetAttribute(si_direct_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = 1;
  e = persistent_grid(si_direct_kernel<TX, TY>, smem, (a.n + TBX_EPC - 1) / TBX_EPC, &grid);
  if (e != cudaSuccess) return e;
  si_direct_kernel<TX, TY><<<grid, TBX_DIRECT_THREADS, smem, s>>>(a, plan, d);
  return cudaGetLastError();
}

template <int TX, int TY> static cudaError_t launch_ami(const RenderArgs &a, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  const int smem = d.smem_total + ami_tab_smem_bytes(plan.dh); /* three small per-pixel tables go behind the rest */
  cudaError_t e = cudaFuncSetAttribute(ami_direct_kernel<TX, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  int grid = 1;
  e = persistent_grid(ami_direct_kernel<TX, TY>, smem, (a.n + TBX_EPC - 1) / TBX_EPC, &grid);
  if (e != cudaSuccess) return e;
  ami_direct_kernel<TX, TY><<<grid, TBX_DIRECT_THREADS, smem, s>>>(a, plan, d);
  return cudaGetLastError();
}

cudaError_t tbx_launch_direct(int game, int tx, int ty, const RenderArgs &a, const void *cfg_host, const TbxAreaPlan &plan, const DirectArgs &d, cudaStream_t s) {
  if (game == TBX_AMIDAR) { /* the tables are built for tx <= 4 only */
    if (ty <= 3) return tx <= 3 ? launch_ami<3, 3>(a, plan, d, s) : launch_ami<4, 3>(a, plan, d, s);
    return tx <= 3 ? launch_ami<3, 4>(a, plan, d, s) : launch_ami<4, 4>(a, plan, d, s);
  }
  if (game == TBX_SPACE_INVADERS) {
    if (ty <= 3) {
      if (tx <= 3) return launch_si<3, 3>(a, plan, d, s);
      if (tx <= 4) return launch_si<4, 3>(a, plan, d, s);
      return launch_si<5, 3>(a, plan, d, s);
    }
    if (tx <= 3) return launch_si<3, 4>(a, plan, d, s);
    if (tx <= 4) return launch_si<4, 4>(a, plan, d, s);
    return launch_si<5, 4>(a, plan, d, s);
  }
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

#ifdef TBX_SI_STATS
extern "C" int tbx_debug_si_stats(unsigned long long *out, int reset) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out, tbxk::d_si_stats, sizeof(unsigned long long) * 48) != cudaSuccess) return 1;
  if (reset) { unsigned long long z[48] = {0}; cudaMemcpyToSymbol(tbxk::d_si_stats, z, sizeof(z)); }
  return 0;
}
#endif
