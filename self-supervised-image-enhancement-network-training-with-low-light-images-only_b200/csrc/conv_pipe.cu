// Persistent, pipelined tcgen05 gather GEMM for the stride-1 conv layers (forward convs and data gradients;
// model.py:33-47, 125-141 and their backward).  Same implicit GEMM and the same halo-window trick as
// conv_gather_halo_kernel (conv_umma.cu): M = 128 output pixels (a 16x8 tile), N = Cout, K = taps x 64-channel slabs, the
// A descriptor of tap (dh,dw) points INTO one TMA-loaded halo window.  What is new is the schedule:
//
//   * ONE CTA per SM walks the tiles  blockIdx.x, blockIdx.x + gridDim.x, ...   (barriers, TMEM, bias staged once);
//   * TWO LANES inside the CTA, each an MMA-issuing warp + four epilogue warps + two TMEM accumulators: lane g takes the
//     tiles i = g, g + 2, ...  Measured on B200: a single issuing thread cannot keep the tensor pipe busy - the command
//     queue behind tcgen05.mma is shallow, so the ~100 uniform-datapath instructions between two chunks of MMAs and the
//     barrier waits between two tiles show up as pipe idle time (3000 clk per 36-MMA tile instead of 1728).  With two
//     issuers the MMAs of one lane fill the gaps of the other, as two co-resident CTAs would, without a second prologue;
//   * up to FOUR halo buffers: the TMA producer loads the windows of the next tiles while the current ones are multiplied;
//   * weights stay RESIDENT in shared memory when the layer's packed slabs fit (every 3x3 64->64 layer: 72 KB) and are
//     loaded once per CTA instead of once per tile; otherwise they stream through a ring whose chunks are SHARED by the
//     two lanes (one chunk feeds the MMAs of two tiles: half the L2 -> SM weight traffic per tile);
//   * the MMAs of one weight chunk (G slabs, template parameter) are unrolled, their window offsets relative to the
//     chunk's first slab are loop invariant words in uniform registers: two uniform adds + one UTCHMMA per MMA;
//   * EPI_BF16 outputs leave through shared memory: each thread writes its pixel's 128-byte row into a SWIZZLE_128B
//     staging tile (conflict-free), a dedicated warp issues one TMA tensor store per 64-channel group - full 128-byte
//     lines instead of 32 partial lines per store instruction, rows beyond the image are clipped by the TMA unit, and
//     the store's issue latency (~600 clk) is off the epilogue warps' critical path.
//
// Warp roles (416 threads), ordered by the SM's issue priority (highest warp id first): warps 0..3 / 4..7 = epilogue of
// lane 0 / 1 (TMEM lane quarter = warp & 3), warp 8 = halo producer, warp 9 = weight producer, warp 10 = TMA store
// issuer, warps 11 / 12 = MMA issuers of lane 0 / 1 (warp 11 owns the TMEM allocation).  The single-thread roles wait
// politely (try_wait with a suspend hint + nanosleep): a spinning warp above the epilogue warps in issue priority
// slowed those to ~10 clk per instruction.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "common.cuh"
#include "epilogue.cuh"
#include "kernels.h"
#include "umma_ptx.cuh"

#define PIPE_THREADS 416
#define W_EPI0 0
#define W_EPI1 4
#define W_HALO 8
#define W_WGT 9
#define W_STORE 10
#define W_MMA0 11
#define W_MMA1 12
#define PIPE_MAX_WSLOTS 24
#define PIPE_MAX_ENT 4
#define PIPE_MAX_HB 4
#define PIPE_STG_TILE 16384               // one staging tile: 128 pixels x 64 bf16
#define PIPE_SMEM_BUDGET (222 * 1024)     // dynamic shared memory incl. 1 KB alignment slack

struct PipeArgs {
  int nh;                      // distinct (source, channel-slab) halo windows per tile
  int src[SS_MAX_WIN];
  int c0[SS_MAX_WIN];
  int pad, halo_bytes, tmem_cols, n_tiles;
  int nhb;                     // halo buffers (2..4)
  int lanes;                   // 1 or 2 issuer + epilogue lanes (2 needs 4 accumulators: 4 * Npad <= 512 TMEM columns)
  int G, n_iter, slots, resident;   // weight chunks of G slabs; `slots` chunk buffers; resident: slots == n_iter, loaded once
  int staged, n_ent, nsb;      // staged epilogue: nsb (1 or 2) staging buffers of n_ent tiles per lane
  int head;                    // EPI_HEAD: the fp32 (b, c, h, w) reflectance tile is staged as [c][16][8] (2 tiles = 32 KB) and
                               // leaves with ONE TMA store; the bf16 copies for the illumination net are written directly
  int ent_map[PIPE_MAX_ENT];   // which output map (0 = out, 1 = out_lo, 2 = out2) and channel offset of each staging tile
  int ent_c0[PIPE_MAX_ENT];
  int n_main, has_lo, n_split, n_store2, t_lo0, t_20;   // accumulator column -> staging tile (see epi_stage16)
  int nslabs, Npad, N, OH, OW, tiles_w, tiles_h;
  uint16_t aoff[SS_MAX_SLABS + 3];   // per slab: (byte offset of its A window inside a halo block) >> 4
  int adelta[12];              // slab j of ANY chunk: aoff[it * G + j] - aoff[it * G]  (checked on the host; G <= 9)
};
struct alignas(64) PipeOutMaps { CUtensorMap m[3]; };

SS_DEVINL void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
SS_DEVINL void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
SS_DEVINL void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
SS_DEVINL void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

SS_DEVINL void load16(const bf16* p, float* f) {
  unpack8(*reinterpret_cast<const uint4*>(p), f);
  unpack8(*reinterpret_cast<const uint4*>(p + 8), f + 8);
}
SS_DEVINL void pack16(const float* v, uint4& lo, uint4& hi) {
  lo.x = pack2(v[0], v[1]);  lo.y = pack2(v[2], v[3]);  lo.z = pack2(v[4], v[5]);   lo.w = pack2(v[6], v[7]);
  hi.x = pack2(v[8], v[9]);  hi.y = pack2(v[10], v[11]); hi.z = pack2(v[12], v[13]); hi.w = pack2(v[14], v[15]);
}

// tile blockIdx.x + i * step -> (tile column, tile row, image), advanced without divisions
struct TileIter { int twi, thi, b, sw, sh, sb, tiles_w, tiles_h; };
SS_DEVINL void tile_init(TileIter& t, int first, int step, int tiles_w, int tiles_h) {
  t.tiles_w = tiles_w; t.tiles_h = tiles_h;
  t.twi = first % tiles_w; int r = first / tiles_w; t.thi = r % tiles_h; t.b = r / tiles_h;
  t.sw = step % tiles_w; r = step / tiles_w; t.sh = r % tiles_h; t.sb = r / tiles_h;
}
SS_DEVINL void tile_next(TileIter& t) {
  t.twi += t.sw;
  int c = (t.twi >= t.tiles_w) ? 1 : 0;
  t.twi -= c ? t.tiles_w : 0;
  t.thi += t.sh + c;
  c = (t.thi >= t.tiles_h) ? 1 : 0;
  t.thi -= c ? t.tiles_h : 0;
  t.b += t.sb + c;
}

// 16 accumulator columns [c, c + 16) of one pixel -> bias / residual / ReLU / masks (the arithmetic of epi_apply16,
// EPI_BF16) -> this pixel's row of the staging tile(s).  Column -> staging tile without tables: main output columns
// [0, n_main) go to tile c / 64 (their bf16 residual to tile t_lo0 + c / 64), columns [n_split, n_split + n_store2) of a
// split output to tile t_20 + (c - n_split) / 64.
SS_DEVINL void epi_stage16(const Epi& e, const PipeArgs& pa, int c, bool ok, int b, int oh, int ow, float* v,
                           const float* bias_s, unsigned char* stg, int row) {
  const bool second = pa.n_split > 0 && c >= pa.n_split;
  const int cc = second ? c - pa.n_split : c;
  if (cc >= (second ? pa.n_store2 : pa.n_main)) return;
  {
    const float4* bp = reinterpret_cast<const float4*>(bias_s + c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bv = bp[q];
      v[4 * q] += bv.x; v[4 * q + 1] += bv.y; v[4 * q + 2] += bv.z; v[4 * q + 3] += bv.w;
    }
  }
  if (second) {
    if (e.mask2 && ok) {
      float m[16];
      load16(e.mask2 + b * e.m2B + oh * e.m2H + ow * e.m2W + cc, m);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (m[i] > 0.f) ? v[i] : 0.f;
    }
  } else {
    if (e.add && ok) {
      float a[16];
      load16(e.add + b * e.aB + oh * e.aH + ow * e.aW + c, a);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] += a[i];
    }
    if (e.relu) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (e.mask && ok) {
      float m[16];
      load16(e.mask + b * e.mB + oh * e.mH + ow * e.mW + c, m);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (m[i] > 0.f) ? v[i] : 0.f;
    }
  }
  const int t = (second ? pa.t_20 : 0) + (cc >> 6);
  const int j = (cc & 63) >> 3, sw = row & 7;
  uint4 lo, hi;
  pack16(v, lo, hi);
  unsigned char* rp = stg + t * PIPE_STG_TILE + row * 128;
  *reinterpret_cast<uint4*>(rp + ((j ^ sw) << 4)) = lo;
  *reinterpret_cast<uint4*>(rp + (((j + 1) ^ sw) << 4)) = hi;
  if (pa.has_lo && !second) {      // residual of the bf16 rounding (hi+lo pairs)
    float r[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = v[i] - bf2f(f2bf(v[i]));
    pack16(r, lo, hi);
    unsigned char* rl = stg + (pa.t_lo0 + (cc >> 6)) * PIPE_STG_TILE + row * 128;
    *reinterpret_cast<uint4*>(rl + ((j ^ sw) << 4)) = lo;
    *reinterpret_cast<uint4*>(rl + (((j + 1) ^ sw) << 4)) = hi;
  }
}

// sigmoid head (model.py:65-70), 16 accumulator columns [c, c + 16) of one pixel: R -> fp32 staging tile [band][h][w]
// (a warp's 32 pixels are 4 image rows x 8 columns = 128 contiguous bytes per band: conflict-free), I -> its fp32 plane,
// and the bf16 cat[R, I] (+ residual) the illumination net reads, exactly as epi_apply16 / EPI_HEAD writes them
SS_DEVINL void epi_head16_staged(const Epi& e, int c, bool ok, int b, int oh, int ow, float* v, const float* bias_s,
                                 float* stg32, int row) {
  {
    const float4* bp = reinterpret_cast<const float4*>(bias_s + c);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bv = bp[q];
      v[4 * q] += bv.x; v[4 * q + 1] += bv.y; v[4 * q + 2] += bv.z; v[4 * q + 3] += bv.w;
    }
  }
  const int64_t pix = ((int64_t)b * e.H + oh) * e.W + ow;
  float i_lo = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int n = c + i;
    const float s = sigmoidf_(v[i]);
    if (n < e.C) stg32[n * 128 + row] = s;
    else if (n == e.C && e.I32 && ok) e.I32[pix] = s;
    v[i] = (n <= e.C) ? s : ((n == e.C + 1 && e.ri_lo_off > 0) ? i_lo : 0.f);     // lane C + 1: I's bf16 residual
    if (n == e.C) i_lo = s - bf2f(f2bf(s));
  }
  if (e.RI && ok) {
    bf16* p = e.RI + pix * e.ri_c + c;
    uint4 lo, hi;
    pack16(v, lo, hi);
    *reinterpret_cast<uint4*>(p) = lo;
    *reinterpret_cast<uint4*>(p + 8) = hi;
    if (e.ri_lo_off > 0 && c < e.C) {
      float r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = (c + i < e.C) ? v[i] - bf2f(f2bf(v[i])) : 0.f;
      pack16(r, lo, hi);
      bf16* q = p + e.ri_lo_off;
      *reinterpret_cast<uint4*>(q) = lo;
      *reinterpret_cast<uint4*>(q + 8) = hi;
    }
  }
}


template <int GU>      // slabs per weight chunk: the MMA issue loop is unrolled over one chunk
__global__ void __launch_bounds__(PIPE_THREADS, 1)
conv_gather_pipe_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ PipeOutMaps omaps,
                        const __grid_constant__ Epi epi, const __grid_constant__ PipeArgs pa) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t halo_full[PIPE_MAX_HB];
  __shared__ __align__(8) uint64_t halo_empty[PIPE_MAX_HB];
  __shared__ __align__(8) uint64_t acc_full[4];
  __shared__ __align__(8) uint64_t acc_empty[4];
  __shared__ __align__(8) uint64_t stg_full[4];         // [lane * 2 + staging buffer]
  __shared__ __align__(8) uint64_t stg_empty[4];
  __shared__ __align__(8) uint64_t w_full[PIPE_MAX_WSLOTS];
  __shared__ __align__(8) uint64_t w_empty[PIPE_MAX_WSLOTS];
  __shared__ uint32_t tmem_base_smem;
  __shared__ __align__(16) float bias_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* dyn_ptr = smem_dyn + (dyn_base - smem_u32(smem_dyn));
  const int Npad = pa.Npad, nh = pa.nh, pad = pa.pad, n_iter = pa.n_iter, slots = pa.slots, nhb = pa.nhb;
  const int lanes = pa.lanes;
  constexpr int G = GU;
  const bool resident = pa.resident != 0;
  const uint32_t b_bytes = (uint32_t)Npad * 128u;
  const uint32_t halo_bytes = (uint32_t)pa.halo_bytes;
  const uint32_t hblock = (uint32_t)nh * halo_bytes;                 // one tile's windows
  const uint32_t ring_base = dyn_base + (uint32_t)nhb * hblock;
  const uint32_t chunk_bytes = (uint32_t)G * b_bytes;
  const uint32_t stg_off = (uint32_t)nhb * hblock + (uint32_t)slots * chunk_bytes;
  const uint32_t stg_buf = (uint32_t)pa.n_ent * PIPE_STG_TILE;       // one staging buffer; a lane owns pa.nsb (1 or 2) of them
  const uint32_t stg_lane = (uint32_t)pa.nsb * stg_buf;
  const int pitch = HALO_TW + 2 * pad;
  const int n_my = (pa.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int n_pairs = (n_my + lanes - 1) / lanes;                    // rounds of `lanes` concurrent tiles
  const uint32_t halo_tx = (uint32_t)nh * (uint32_t)((HALO_TH + 2 * pad) * pitch * 128);
  const int w_total = resident ? n_iter : n_pairs * n_iter;          // weight chunks this CTA loads

  // ---- prologue: each producer initialises its own barriers and puts its first requests in flight before the CTA-wide
  // synchronisation (TMEM allocation, staging clear) has finished
  TileIter ht;      // halo producer's position (tile order)
  if (warp == W_HALO) {
    if (lane == 0) {
      for (int s = 0; s < nhb; ++s) {
        mbar_init(smem_u32(&halo_full[s]), 1);
        mbar_init(smem_u32(&halo_empty[s]), 1);
      }
      for (int s = 0; s < 4; ++s) {
        mbar_init(smem_u32(&acc_full[s]), 1);
        mbar_init(smem_u32(&acc_empty[s]), 4);          // one arrival per epilogue warp of the lane
      }
      for (int s = 0; s < 4; ++s) {
        mbar_init(smem_u32(&stg_full[s]), 4);
        mbar_init(smem_u32(&stg_empty[s]), 1);
      }
      fence_barrier_init();
      tile_init(ht, (int)blockIdx.x, (int)gridDim.x, pa.tiles_w, pa.tiles_h);
      pdl_wait();                                       // the activations come from the previous kernel
      for (int i = 0; i < nhb && i < n_my; ++i) {
        const uint32_t hb = smem_u32(&halo_full[i]);
        mbar_expect_tx(hb, halo_tx);
        for (int h = 0; h < nh; ++h)
          tma_load_4d(dyn_base + (uint32_t)i * hblock + (uint32_t)h * halo_bytes, &maps.halo[pa.src[h]], hb, pa.c0[h],
                      ht.twi * HALO_TW - pad, ht.thi * HALO_TH - pad, ht.b);
        tile_next(ht);
      }
    }
    __syncwarp();
  } else if (warp == W_WGT) {
    if (lane == 0) {
      for (int s = 0; s < slots; ++s) {
        mbar_init(smem_u32(&w_full[s]), 1);
        mbar_init(smem_u32(&w_empty[s]), (uint32_t)lanes);     // a chunk is free when every lane's MMAs on it have retired
      }
      fence_barrier_init();
      pdl_wait();                                       // (single-layer calls: the weight packer is the previous kernel)
      for (int c = 0; c < slots && c < w_total; ++c) {
        const int s0 = (c % n_iter) * G;
        const uint32_t fb = smem_u32(&w_full[c]);
        mbar_expect_tx(fb, chunk_bytes);
        for (int q = 0; q < G; ++q)
          tma_load_2d(ring_base + (uint32_t)c * chunk_bytes + (uint32_t)q * b_bytes, &maps.w, fb, 0, (s0 + q) * Npad);
      }
    }
    __syncwarp();
  } else if (warp == W_MMA0) {
    tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)pa.tmem_cols);
  } else if (warp < 8) {
    for (int i = (int)threadIdx.x; i < 256; i += 256) bias_s[i] = (epi.bias && i < pa.N) ? __ldg(epi.bias + i) : 0.f;
    if (pa.staged) {      // pad chunks of partially used staging tiles stay zero for the whole kernel
      uint4* z = reinterpret_cast<uint4*>(dyn_ptr + stg_off);
      const int n16 = lanes * pa.nsb * pa.n_ent * (PIPE_STG_TILE / 16);
      for (int i = (int)threadIdx.x; i < n16; i += 256) z[i] = make_uint4(0, 0, 0, 0);
      fence_proxy_async();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == W_HALO) {
    // ===== halo producer =====
    if (lane == 0) {
      uint32_t hbuf = 0, k = 1;                         // tile i -> buffer i % nhb, use i / nhb
      for (int i = nhb; i < n_my; ++i) {
        mbar_wait_polite(smem_u32(&halo_empty[hbuf]), (k & 1u) ^ 1u);
        const uint32_t hb = smem_u32(&halo_full[hbuf]);
        mbar_expect_tx(hb, halo_tx);
        for (int h = 0; h < nh; ++h)
          tma_load_4d(dyn_base + hbuf * hblock + (uint32_t)h * halo_bytes, &maps.halo[pa.src[h]], hb, pa.c0[h],
                      ht.twi * HALO_TW - pad, ht.thi * HALO_TH - pad, ht.b);
        tile_next(ht);
        if (++hbuf == (uint32_t)nhb) { hbuf = 0; ++k; }
      }
    }
  } else if (warp == W_WGT) {
    // ===== weight producer (streaming mode: the ring runs across tile borders) =====
    if (lane == 0) {
      uint32_t slot = 0, use = 1;                     // chunk c -> slot c % slots, use c / slots
      int it = slots % n_iter;
      for (int c = slots; c < w_total; ++c) {
        mbar_wait_polite(smem_u32(&w_empty[slot]), (use & 1u) ^ 1u);
        const int s0 = it * G;
        const uint32_t fb = smem_u32(&w_full[slot]);
        mbar_expect_tx(fb, chunk_bytes);
        for (int q = 0; q < G; ++q)
          tma_load_2d(ring_base + slot * chunk_bytes + (uint32_t)q * b_bytes, &maps.w, fb, 0, (s0 + q) * Npad);
        if (++it == n_iter) it = 0;
        if (++slot == (uint32_t)slots) { slot = 0; ++use; }
      }
    }
  } else if (warp == W_MMA0 || warp == W_MMA1) {
    // ===== MMA issuer of lane g.  Everything in this loop is warp-uniform (kernel parameters, shuffled bases, loop
    // counters) so that descriptors and barrier addresses stay in uniform registers.
    const int g = warp - W_MMA0;
    if (g < lanes) {
      const uint32_t idesc = make_idesc(128, Npad, 0, 0);
      const uint32_t tm = uniform32(tmem_base);
      const uint32_t a_hi = (uint32_t)(make_sdesc(0, 16, (uint32_t)pitch * 128u) >> 32);
      const uint32_t b_hi = (uint32_t)(make_sdesc(0, 16, 1024) >> 32);
      const uint32_t a_lo0 = uniform32(((dyn_base >> 4) & 0x3FFFu) | (1u << 16));
      const uint32_t b_lo0 = uniform32(((ring_base >> 4) & 0x3FFFu) | (1u << 16));
      const uint32_t hf0 = uniform32(smem_u32(&halo_full[0])), he0 = uniform32(smem_u32(&halo_empty[0]));
      const uint32_t af0 = uniform32(smem_u32(&acc_full[0])), ae0 = uniform32(smem_u32(&acc_empty[0]));
      const uint32_t wf0 = uniform32(smem_u32(&w_full[0])), we0 = uniform32(smem_u32(&w_empty[0]));
      const uint32_t hstep = hblock >> 4, bstep = b_bytes >> 4, cstep = chunk_bytes >> 4;
      // loop-invariant words of one chunk's MMAs: window offset of slab j relative to the chunk's first slab (identical
      // for every chunk: checked on the host) and its weight-slab offset
      uint32_t dA[GU], dB[GU];
#pragma unroll
      for (int j = 0; j < GU; ++j) { dA[j] = (uint32_t)pa.adelta[j]; dB[j] = (uint32_t)j * bstep; }
      uint32_t slot = 0, wph = 0;
      uint32_t hbuf = (uint32_t)g, hph = 0;               // tile i = m * lanes + g -> halo buffer i % nhb (nhb >= lanes)
#pragma unroll 1
      for (int m = 0; m < n_pairs; ++m) {
        const bool real = m * lanes + g < n_my;
        if (!real && resident) break;                     // (streaming: a lane without a tile still releases the chunks)
        const uint32_t ab = (uint32_t)(2 * g + (m & 1)), aph = (uint32_t)((m >> 1) & 1);
        if (real) {
          mbar_wait_warp(ae0 + 8u * ab, aph ^ 1u, 0);       // the epilogue has drained this accumulator
          mbar_wait_warp(hf0 + 8u * hbuf, hph, 0);             // halo windows landed
          tc_fence_after();
        }
        const uint32_t a_base = a_lo0 + hbuf * hstep;
        const uint32_t tmd = tm + ab * (uint32_t)Npad;
#pragma unroll 1
        for (int it = 0; it < n_iter; ++it) {
          if (!resident || m == 0) {
            mbar_wait_warp(wf0 + 8u * slot, wph, 0);
            tc_fence_after();
          }
          const uint32_t a_c = a_base + (uint32_t)pa.aoff[it * GU];
          const uint32_t b_c = b_lo0 + slot * cstep;
          const uint32_t acc0 = (it > 0) ? 1u : 0u;
          if (elect_one()) {
            if (real) {
#pragma unroll
              for (int j = 0; j < GU; ++j)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_bf16(tmd, ((uint64_t)a_hi << 32) | (uint64_t)(a_c + dA[j] + 2u * kk),
                            ((uint64_t)b_hi << 32) | (uint64_t)(b_c + dB[j] + 2u * kk), idesc, (j > 0 || kk > 0) ? 1u : acc0);
            }
            if (!resident) umma_commit(we0 + 8u * slot);
            if (real && it == n_iter - 1) {
              umma_commit(he0 + 8u * hbuf);                 // halo buffer free for tile i + nhb
              umma_commit(af0 + 8u * ab);                   // accumulator ready for the epilogue
            }
          }
          __syncwarp();
          if (++slot == (uint32_t)slots) { slot = 0; if (!resident) wph ^= 1u; }
        }
        hbuf += (uint32_t)lanes;
        if (hbuf >= (uint32_t)nhb) { hbuf -= (uint32_t)nhb; hph ^= 1u; }
      }
    }
  } else if (warp < 8) {
    // ===== epilogue warps of lane g =====
    const int g = warp >> 2;
    if (g < lanes) {
      const int quarter = warp & 3;
      const int row = quarter * 32 + lane;
      const uint32_t af0 = smem_u32(&acc_full[0]), ae0 = smem_u32(&acc_empty[0]);
      const uint32_t sf0 = smem_u32(&stg_full[2 * g]), se0 = smem_u32(&stg_empty[2 * g]);
      unsigned char* stg0 = dyn_ptr + stg_off + (uint32_t)g * stg_lane;
      TileIter t;
      tile_init(t, (int)blockIdx.x + g * (int)gridDim.x, lanes * (int)gridDim.x, pa.tiles_w, pa.tiles_h);
      pdl_wait();     // the epilogue reads residual / mask tensors and overwrites buffers earlier kernels may still read
      for (int m = 0; m * lanes + g < n_my; ++m) {
        const uint32_t ab = (uint32_t)(2 * g + (m & 1)), aph = (uint32_t)((m >> 1) & 1);
        const int b = t.b;
        const int oh = t.thi * HALO_TH + (row >> 3), ow = t.twi * HALO_TW + (row & 7);
        const bool ok = oh < pa.OH && ow < pa.OW;      // partial tiles: zero-filled loads, skipped / clipped stores
        const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * (uint32_t)Npad;
        if (pa.staged) {
          // staging buffer sb of this lane, use u: the TMA stores that last read it have finished reading
          const uint32_t sb = (pa.nsb == 2) ? (uint32_t)(m & 1) : 0u, u = (pa.nsb == 2) ? (uint32_t)(m >> 1) : (uint32_t)m;
          const uint32_t sf = sf0 + 8u * sb, se = se0 + 8u * sb;
          unsigned char* stg = stg0 + sb * stg_buf;
          mbar_wait_warp_polite(se, (u & 1u) ^ 1u);
          mbar_wait_warp_polite(af0 + 8u * ab, aph);
          tc_fence_after();
          int n0 = 0;
          for (; n0 + 32 <= Npad; n0 += 32) {
            float v[32];
            tmem_ld32(trow + (uint32_t)n0, v);
            if (pa.head) {
              epi_head16_staged(epi, n0, ok, b, oh, ow, v, bias_s, reinterpret_cast<float*>(stg), row);
              epi_head16_staged(epi, n0 + 16, ok, b, oh, ow, v + 16, bias_s, reinterpret_cast<float*>(stg), row);
            } else {
              epi_stage16(epi, pa, n0, ok, b, oh, ow, v, bias_s, stg, row);
              epi_stage16(epi, pa, n0 + 16, ok, b, oh, ow, v + 16, bias_s, stg, row);
            }
          }
          if (n0 < Npad) {
            float v[16];
            tmem_ld16(trow + (uint32_t)n0, v);
            if (pa.head) epi_head16_staged(epi, n0, ok, b, oh, ow, v, bias_s, reinterpret_cast<float*>(stg), row);
            else epi_stage16(epi, pa, n0, ok, b, oh, ow, v, bias_s, stg, row);
          }
          tc_fence_before();
          fence_proxy_async();                                // staging writes -> visible to the TMA unit
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(ae0 + 8u * ab);                       // accumulator free: the MMAs of tile i + 2 * lanes may start
            mbar_arrive(sf);                                  // this warp's 32 rows are staged
          }
        } else {
          mbar_wait_warp_polite(af0 + 8u * ab, aph);
          tc_fence_after();
          int n0 = 0;
          for (; n0 + 32 <= Npad; n0 += 32) {
            float v[32];
            tmem_ld32(trow + (uint32_t)n0, v);
            if (ok) {
              epi_apply16(epi, b, oh, ow, n0, pa.N, v, bias_s);
              epi_apply16(epi, b, oh, ow, n0 + 16, pa.N, v + 16, bias_s);
            }
          }
          if (n0 < Npad) {
            float v[16];
            tmem_ld16(trow + (uint32_t)n0, v);
            if (ok) epi_apply16(epi, b, oh, ow, n0, pa.N, v, bias_s);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(ae0 + 8u * ab);
        }
        tile_next(t);
      }
    }
  } else if (warp == W_STORE) {
    // ===== TMA store issuer (staged epilogue): tiles in order, lane g = i % lanes =====
    if (lane == 0 && pa.staged) {
      TileIter t;
      tile_init(t, (int)blockIdx.x, (int)gridDim.x, pa.tiles_w, pa.tiles_h);
      int g = 0, m = 0;
      for (int i = 0; i < n_my; ++i) {
        const uint32_t sb = (pa.nsb == 2) ? (uint32_t)(m & 1) : 0u, u = (pa.nsb == 2) ? (uint32_t)(m >> 1) : (uint32_t)m;
        mbar_wait_polite(smem_u32(&stg_full[2 * g + sb]), u & 1u);
        {
          const uint32_t s_u32 = dyn_base + stg_off + (uint32_t)g * stg_lane + sb * stg_buf;
          if (pa.head) {      // fp32 (b, band, h, w) planes: box {8 w, 16 h, 64 bands}
            tma_store_4d(&omaps.m[0], s_u32, t.twi * HALO_TW, t.thi * HALO_TH, 0, t.b);
          } else {
            for (int q = 0; q < pa.n_ent; ++q)
              tma_store_4d(&omaps.m[pa.ent_map[q]], s_u32 + (uint32_t)q * PIPE_STG_TILE, pa.ent_c0[q], t.twi * HALO_TW,
                           t.thi * HALO_TH, t.b);
          }
          bulk_commit();
          bulk_wait_read0();                                  // the staging buffer has been read: the lane may refill it
        }
        mbar_arrive(smem_u32(&stg_empty[2 * g + sb]));
        tile_next(t);
        if (++g == lanes) { g = 0; ++m; }
      }
      bulk_wait_all();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == W_MMA0) tmem_dealloc(tmem_base, (uint32_t)pa.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
struct PipePlan {
  PipeArgs pa;
  PipeOutMaps om;
  int smem;
  int grid;
};
size_t ss_pipe_plan_size() { return sizeof(PipePlan); }

static int sm_count() {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      n <= 0)
    n = 148;
  return n;
}

static int encode_out(const bf16* base, int64_t sB, int64_t sH, int64_t sW, int OH, int OW, int B, CUtensorMap* out) {
  SrcView v;
  v.base = base; v.sB = sB; v.sH = sH; v.sW = sW; v.H = OH; v.W = OW;
  return ss_umma_encode_view(v, (int)sW, HALO_TW, HALO_TH, B, out);
}

// 0: the geometry / epilogue is not taken by this kernel
static int pipe_plan(const ConvGeom& g, const Epi& epi, PipePlan* out) {
  PipeArgs pa;
  memset(&pa, 0, sizeof(pa));
  if (!g.halo_ok || !ss_umma_supported(g) || g.Npad > 256) return 0;
  int pad = 0;
  for (int i = 0; i < g.nslabs; ++i) {
    const Slab& sl = g.slab[i];
    pad = std::max(pad, std::max(abs((int)sl.dh), abs((int)sl.dw)));
    int h = -1;
    for (int j = 0; j < pa.nh; ++j)
      if (pa.src[j] == sl.src && pa.c0[j] == sl.c0) h = j;
    if (h < 0) {
      if (pa.nh == SS_MAX_WIN) return 0;
      pa.src[pa.nh] = sl.src; pa.c0[pa.nh] = sl.c0; ++pa.nh;
    }
  }
  if (pad > 4) return 0;
  pa.pad = pad;
  const int rows = (HALO_TH + 2 * pad) * (HALO_TW + 2 * pad);
  pa.halo_bytes = (rows * 128 + 1023) / 1024 * 1024;
  pa.tiles_w = (g.OW + HALO_TW - 1) / HALO_TW;
  pa.tiles_h = (g.OH + HALO_TH - 1) / HALO_TH;
  pa.n_tiles = g.B * pa.tiles_h * pa.tiles_w;
  pa.nslabs = g.nslabs; pa.Npad = g.Npad; pa.N = g.N; pa.OH = g.OH; pa.OW = g.OW;
  const int pitch = HALO_TW + 2 * pad;
  for (int i = 0; i < g.nslabs; ++i) {
    const Slab& sl = g.slab[i];
    int h = 0;
    for (int j = 0; j < pa.nh; ++j)
      if (pa.src[j] == sl.src && pa.c0[j] == sl.c0) h = j;
    pa.aoff[i] = (uint16_t)((h * pa.halo_bytes + ((sl.dh + pad) * pitch + (sl.dw + pad)) * 128) >> 4);
  }

  // ---- staged epilogue: staging tiles = 64-channel groups of the main output [, of its bf16 residual] [, of the second
  // output of a column split]
  bool staged = epi.mode == EPI_BF16 && ss_env_int("SSHSLIE_PIPE_STAGED", 1) != 0 && epi.oW >= 64 && (epi.oW % 8) == 0 &&
                (epi.n_store % 16) == 0 && (epi.n_split % 64) == 0 && (epi.n_store2 % 16) == 0;
  if (staged && epi.n_split && (epi.o2W < 64 || (epi.o2W % 8))) staged = false;
  if (epi.mode == EPI_HEAD && ss_env_int("SSHSLIE_PIPE_STAGED", 1) != 0 && epi.C == 64 && epi.R32 && (g.OW % 4) == 0) {
    staged = true;
    pa.head = 1;
    pa.n_ent = 2;             // 64 bands x 128 pixels x fp32 = two 16 KB staging tiles
  } else if (staged) {
    const int n1 = epi.n_split ? std::min(epi.n_store, epi.n_split) : epi.n_store;
    const int n_main_t = (n1 + 63) / 64, n_2_t = epi.n_split ? (epi.n_store2 + 63) / 64 : 0;
    const int ne = n_main_t * (epi.out_lo ? 2 : 1) + n_2_t;
    if (ne > PIPE_MAX_ENT || ne < 1) staged = false;
    else {
      int e = 0;
      for (int q = 0; q < n_main_t; ++q) { pa.ent_map[e] = 0; pa.ent_c0[e] = q * 64; ++e; }
      pa.t_lo0 = e;
      if (epi.out_lo) for (int q = 0; q < n_main_t; ++q) { pa.ent_map[e] = 1; pa.ent_c0[e] = q * 64; ++e; }
      pa.t_20 = e;
      for (int q = 0; q < n_2_t; ++q) { pa.ent_map[e] = 2; pa.ent_c0[e] = q * 64; ++e; }
      pa.n_ent = ne;
      pa.n_main = n1; pa.has_lo = epi.out_lo ? 1 : 0; pa.n_split = epi.n_split; pa.n_store2 = epi.n_split ? epi.n_store2 : 0;
    }
  }

  // ---- slabs per weight chunk: the largest of {9, 6, 4, 3, 2, 1} that divides the slab count, keeps a chunk <= 48 KB and
  // has the SAME window offsets (relative to the chunk's first slab) in every chunk - the kernel unrolls one chunk
  const int b_bytes = g.Npad * 128;
  static const int kG[6] = {9, 6, 4, 3, 2, 1};
  const int g_cap = ss_env_int("SSHSLIE_PIPE_G", 9);
  int G = 1;
  for (int c = 0; c < 6; ++c) {
    const int gc = kG[c];
    if (gc > g_cap || (g.nslabs % gc) || (gc > 1 && gc * b_bytes > 48 * 1024)) continue;
    bool same = true;
    for (int it = 0; it < g.nslabs / gc && same; ++it)
      for (int j = 0; j < gc && same; ++j)
        same = ((int)pa.aoff[it * gc + j] - (int)pa.aoff[it * gc]) == ((int)pa.aoff[j] - (int)pa.aoff[0]);
    if (same) { G = gc; break; }
  }
  pa.G = G;
  for (int j = 0; j < G; ++j) pa.adelta[j] = (int)pa.aoff[j] - (int)pa.aoff[0];
  pa.n_iter = g.nslabs / G;

  // ---- shared memory: halo buffers + weights (resident if they fit, else a ring) + one staging buffer per lane.
  // Preference: two lanes > resident weights > staged epilogue > deeper halo prefetch.
  const int chunk = G * b_bytes;
  const int w_all = pa.n_iter * chunk;
  const int budget = PIPE_SMEM_BUDGET - 1024;
  const int allow_res = ss_env_int("SSHSLIE_PIPE_RESIDENT", 1);
  const int nhb_max = std::max(2, std::min(PIPE_MAX_HB, ss_env_int("SSHSLIE_PIPE_NHB", PIPE_MAX_HB)));
  const int lanes_max = std::max(1, std::min(2, ss_env_int("SSHSLIE_PIPE_LANES", 2)));
  const int hblock = pa.nh * pa.halo_bytes;
  bool found = false;
  for (int lanes = lanes_max; lanes >= 1 && !found; --lanes) {
    if (2 * lanes * g.Npad > 512) continue;
    for (int res = 1; res >= 0 && !found; --res) {
      if (res && (!allow_res || pa.n_iter > PIPE_MAX_WSLOTS)) continue;
      for (int st = staged ? 2 : 0; st >= 0 && !found; --st) {       // 2 / 1 staging buffers per lane, 0 = direct stores
        const int stg = st * lanes * pa.n_ent * PIPE_STG_TILE;
        if (st > ss_env_int("SSHSLIE_PIPE_NSB", 2)) continue;
        for (int nhb = nhb_max; nhb >= 2 && !found; --nhb) {
          if (st == 2 && nhb < std::min(3, nhb_max)) continue;           // a second staging buffer only if >= 3 halo buffers still fit
          const int fixed = nhb * hblock + stg;
          if (res) {
            if (fixed + w_all > budget) continue;
            pa.resident = 1; pa.slots = pa.n_iter;
          } else {
            int slots = (budget - fixed) / chunk;
            slots = std::min(slots, std::min(8, std::max(2, pa.n_iter)));
            if (slots < 2) continue;
            // streaming prefers ring depth over halo depth: at least 3 chunk buffers before a third / fourth halo buffer
            if (nhb > 2 && slots < 3) continue;
            pa.resident = 0; pa.slots = slots;
          }
          pa.lanes = lanes; pa.nhb = nhb; pa.staged = st ? 1 : 0; pa.nsb = st;
          found = true;
        }
      }
    }
  }
  if (!found) return 0;
  if (!pa.staged) { pa.n_ent = 0; pa.nsb = 0; }
  int cols = 32;
  while (cols < 2 * pa.lanes * g.Npad) cols <<= 1;
  pa.tmem_cols = cols;
  out->pa = pa;
  out->smem = pa.nhb * hblock + pa.slots * chunk + pa.lanes * pa.nsb * pa.n_ent * PIPE_STG_TILE + 1024;
  const int sms = sm_count();
  const int waves = (pa.n_tiles + sms - 1) / sms;
  out->grid = (pa.n_tiles + waves - 1) / waves;
  return 1;
}

int ss_umma_pipe_supported(const ConvGeom& g, const Epi& epi) {
  PipePlan p;
  return pipe_plan(g, epi, &p);
}

// plan_cache: ss_pipe_plan_size() bytes owned by the caller, zero-initialised; filled on the first launch of this
// (geom, epilogue) and reused afterwards (output tensor maps are built once)
int ss_launch_conv_gather_pipe(const ConvGeom& g, const UmmaMaps& maps, const Epi& epi, void* plan_cache, int* cache_valid,
                               cudaStream_t st) {
  PipePlan local;
  PipePlan* p = plan_cache ? reinterpret_cast<PipePlan*>(plan_cache) : &local;
  if (!plan_cache || !cache_valid || !*cache_valid) {
    if (!pipe_plan(g, epi, p)) {
      ss_set_error("conv_gather_pipe: geometry not eligible");
      return SSHSLIE_ERR_ARG;
    }
    memset(&p->om, 0, sizeof(p->om));
    if (p->pa.head) {
      const uint64_t dims[4] = {(uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)epi.C, (uint64_t)g.B};
      const uint64_t str[3] = {(uint64_t)g.OW * 4, (uint64_t)g.OW * g.OH * 4, (uint64_t)g.OW * g.OH * epi.C * 4};
      const uint32_t box[4] = {HALO_TW, HALO_TH, 64, 1};
      const int rc = ss_tma_encode_4d(&p->om.m[0], 1, epi.R32, dims, str, box, 0);
      if (rc) return rc;
    } else if (p->pa.staged) {
      int rc = encode_out(epi.out, epi.oB, epi.oH, epi.oW, g.OH, g.OW, g.B, &p->om.m[0]);
      if (!rc && epi.out_lo) rc = encode_out(epi.out_lo, epi.oB, epi.oH, epi.oW, g.OH, g.OW, g.B, &p->om.m[1]);
      if (!rc && epi.n_split) rc = encode_out(epi.out2, epi.o2B, epi.o2H, epi.o2W, g.OH, g.OW, g.B, &p->om.m[2]);
      if (rc) return rc;
    }
    if (cache_valid) *cache_valid = 1;
  }
  typedef void (*PipeKernel)(const UmmaMaps, const PipeOutMaps, const Epi, const PipeArgs);
  static const PipeKernel kernels[10] = {nullptr, conv_gather_pipe_kernel<1>, conv_gather_pipe_kernel<2>,
                                         conv_gather_pipe_kernel<3>, conv_gather_pipe_kernel<4>, nullptr,
                                         conv_gather_pipe_kernel<6>, nullptr, nullptr, conv_gather_pipe_kernel<9>};
  const PipeKernel kernel = (p->pa.G >= 1 && p->pa.G <= 9) ? kernels[p->pa.G] : nullptr;
  if (!kernel) {
    ss_set_error("conv_gather_pipe: no kernel for %d slabs per chunk", p->pa.G);
    return SSHSLIE_ERR_ARG;
  }
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    for (int gi = 1; gi <= 9; ++gi)
      if (kernels[gi] &&
          cudaFuncSetAttribute(kernels[gi], cudaFuncAttributeMaxDynamicSharedMemorySize, PIPE_SMEM_BUDGET) != cudaSuccess) {
        ss_set_error("conv_gather_pipe: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
        return SSHSLIE_ERR_CUDA;
      }
    attr_once.set();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(p->grid);
  cfg.blockDim = dim3(PIPE_THREADS);
  cfg.dynamicSmemBytes = (size_t)p->smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ss_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, maps, p->om, epi, p->pa);
  return ss_check_launch("conv_gather_pipe");
}

