// tcgen05 / TMEM / TMA executors of a ConvGeom (plan.h) for sm_100a.
//
// Gather GEMM (forward convs, transposed convs, data gradients):
//   one CTA = one tile of 128 output pixels (a th x tw block of the OH x OW grid) x Npad output channels.
//   K runs over the geom's slabs (tap x 64-channel slab).  Per slab the TMA producer issues
//     * one 4-D tiled load of the shifted activation window  box {64 ch, tw, th, 1}  -> 128 rows x 128 B,
//       SWIZZLE_128B, zero fill outside the image (this IS the conv padding; no im2col buffer exists),
//     * one 2-D load of the packed weight slab               box {64 k, Npad}        -> Npad rows x 128 B,
//   into a 4-stage shared-memory ring; one elected thread issues 4 x tcgen05.mma (M=128, N=Npad, K=16, bf16 -> fp32
//   accumulators in TMEM) per slab and releases the stage with tcgen05.commit; after the last slab the four
//   epilogue warps pull the accumulator with tcgen05.ld (32 lanes x 16 columns) and run the shared epilogue
//   (bias / residual / ReLU / ReLU-mask / sigmoid heads, epilogue.cuh) straight to global memory.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>

#include "common.cuh"
#include "epilogue.cuh"
#include "kernels.h"

#include "umma_ptx.cuh"

#define UM_STAGES 4
#define UM_THREADS 192

size_t ss_umma_maps_size() { return sizeof(UmmaMaps); }
static std::atomic<int> g_pdl{-1};
int ss_pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SSHSLIE_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v;
}


// ---------------------------------------------------------------------------------------------
// gather GEMM kernel
// ---------------------------------------------------------------------------------------------
// The geometry arrives as a __grid_constant__ kernel parameter (uniform constant-bank loads, no global->shared copy in
// the prologue); loop-invariant words of the MMA loop are hoisted (barrier addresses, descriptor bases, slab count).
SS_DEVINL void conv_gather_umma_body(const ConvGeom& g, const UmmaMaps& maps, const Epi& epi, int tmem_cols,
                                     int tile_index) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  // operand ring, 1024-byte aligned (SWIZZLE_128B atom)
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int Npad = g.Npad, nslabs = g.nslabs;
  const uint32_t b_bytes = (uint32_t)Npad * 128u;
  const uint32_t stage_bytes = UM_A_BYTES + b_bytes;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < UM_STAGES; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // tile -> (b, oh0, ow0)
  const int tiles_w = (g.OW + g.tw - 1) / g.tw, tiles_h = (g.OH + g.th - 1) / g.th;
  int t = tile_index;
  const int twi = t % tiles_w; t /= tiles_w;
  const int thi = t % tiles_h;
  const int b = t / tiles_h;
  const int oh0 = thi * g.th, ow0 = twi * g.tw;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      pdl_wait();      // everything below reads tensors produced by earlier kernels
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t st = 0, ph = 1;
      for (int s = 0; s < nslabs; ++s) {
        mbar_wait(empty0 + 8u * st, ph);
        const uint32_t a_dst = dyn_base + st * stage_bytes;
        const uint32_t fb = full0 + 8u * st;
        mbar_expect_tx(fb, stage_bytes);
        tma_load_4d(a_dst, &maps.src[g.slab[s].src], fb, g.slab[s].c0, ow0 + g.slab[s].dw, oh0 + g.slab[s].dh, b);
        tma_load_2d(a_dst + UM_A_BYTES, &maps.w, fb, 0, s * Npad);
        if (++st == UM_STAGES) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: whole warp converged, one elected lane issues =====
    const uint32_t idesc = make_idesc(128, Npad, 0, 0);
    const uint32_t tm = uniform32(tmem_base);
    const uint32_t hi = (uint32_t)(make_sdesc(0, 16, 1024) >> 32);
    const uint32_t a_lo0 = uniform32(((dyn_base >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t b_lo0 = a_lo0 + (UM_A_BYTES >> 4);
    const uint32_t sstep = stage_bytes >> 4;
    const uint32_t full0 = uniform32(smem_u32(&full_bar[0])), empty0 = uniform32(smem_u32(&empty_bar[0]));
    const uint32_t accb = uniform32(smem_u32(&accum_bar));
    uint32_t st = 0, ph = 0;
#pragma unroll 1
    for (int s = 0; s < nslabs; ++s) {
      mbar_wait_warp(full0 + 8u * st, ph, 0);
      tc_fence_after();
      const uint32_t a_lo = a_lo0 + st * sstep, b_lo = b_lo0 + st * sstep;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tm, ((uint64_t)hi << 32) | (uint64_t)(a_lo + 2u * k), ((uint64_t)hi << 32) | (uint64_t)(b_lo + 2u * k),
                    idesc, (s > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty0 + 8u * st);      // frees the stage when these MMAs retire
        if (s == nslabs - 1) umma_commit(accb);
      }
      __syncwarp();
      if (++st == UM_STAGES) { st = 0; ph ^= 1u; }
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int oh = oh0 + row / g.tw, ow = ow0 + row % g.tw;
    const bool ok = oh < g.OH && ow < g.OW;
    pdl_wait();        // the epilogue reads residual / mask tensors and overwrites buffers earlier kernels may still read
    mbar_wait_warp(smem_u32(&accum_bar), 0, 100);
    tc_fence_after();
    for (int n0 = 0; n0 < Npad; n0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)n0, v);
      if (ok) epi_apply16(epi, b, oh, ow, n0, g.N, v);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)tmem_cols);
}

__global__ void __launch_bounds__(UM_THREADS, 1)
conv_gather_umma_kernel(const __grid_constant__ ConvGeom g, const __grid_constant__ UmmaMaps maps,
                        const __grid_constant__ Epi epi, int tmem_cols) {
  conv_gather_umma_body(g, maps, epi, tmem_cols, blockIdx.x);
}

// the four output-parity classes of a transposed / strided-dgrad layer in ONE launch: blockIdx.y = class (qh, qw).
// The classes are four consecutive ConvGeoms with their own weight packs; their outputs interleave in the same tensor.
struct alignas(64) UmmaMaps4 { UmmaMaps m[4]; };
struct alignas(16) ConvGeom4 { ConvGeom g[4]; };
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_gather_umma4_kernel(const __grid_constant__ ConvGeom4 g4, const __grid_constant__ UmmaMaps4 maps,
                         const __grid_constant__ Epi epi, int tmem_cols) {
  const int q = blockIdx.y, qh = q >> 1, qw = q & 1;
  Epi e = epi;
  e.out += qh * (e.oH >> 1) + qw * (e.oW >> 1);
  if (e.add) e.add += qh * (e.aH >> 1) + qw * (e.aW >> 1);
  if (e.mask) e.mask += qh * (e.mH >> 1) + qw * (e.mW >> 1);
  conv_gather_umma_body(g4.g[q], maps.m[q], e, tmem_cols, blockIdx.x);
}

// ---------------------------------------------------------------------------------------------
// gather GEMM with HALO REUSE (stride-1 taps):  the per-tap activation windows of a tile overlap almost entirely,
// so instead of one TMA load per tap the CTA loads ONE halo window per (source, 64-channel slab) —
// (16+2p) x (8+2p) pixels x 64 ch — and points the UMMA A-descriptor of tap (dh,dw) INTO it:
//     start = halo + ((dh+p) * pitch + (dw+p)) * 128 B,   SBO = pitch * 128 B  (one 8-pixel output row per group)
// (the 128B swizzle is a function of the shared-memory address, for TMA writes and UMMA reads alike, so a start
// that is 128B- but not 1024B-aligned still addresses consistently; descriptor base_offset stays 0 — verified on B200).
// Only the weight slabs stream, HG slabs per ring stage: the MMA warp pays ONE barrier wait / fence / commit per
// 4*HG*T MMAs (the per-slab version of this kernel spent 3x the MMA time in that bookkeeping), and T pixel tiles per
// CTA share each weight slab.  L2->SM traffic per MMA drops from 6 KB (per-tap kernel) to 2 KB / T.
// ---------------------------------------------------------------------------------------------
#define HALO_MAX_T 4
#define HALO_MAX_STAGES 8

struct HaloArgs {
  int nh;                  // distinct (source, channel-slab) halo windows per tile
  int src[SS_MAX_WIN];
  int c0[SS_MAX_WIN];
  int pad;                 // p = max |dh|,|dw|
  int T;                   // pixel tiles per CTA (template parameter of the kernel)
  int halo_bytes;          // one window, rounded up to 1024
  int tmem_cols;
  int n_tiles;
  int stages;              // depth of the weight ring (the only operand that streams)
  int G;                   // weight slabs per ring stage
  int nslabs, Npad, N, OH, OW;
  uint16_t aoff[SS_MAX_SLABS + 3];   // per slab: (byte offset of its A window inside tile 0's halo block) >> 4.
                                     // Lives in the kernel's constant bank: the MMA loop reads it with uniform loads.
};


template <int T>
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_gather_halo_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ Epi epi,
                        const __grid_constant__ HaloArgs ha) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[HALO_MAX_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[HALO_MAX_STAGES];
  __shared__ __align__(8) uint64_t halo_bar;
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float bias_s[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int Npad = ha.Npad, nslabs = ha.nslabs;
  const uint32_t b_bytes = (uint32_t)Npad * 128u;
  const int nh = ha.nh, pad = ha.pad, stages = ha.stages, halo_bytes = ha.halo_bytes, G = ha.G;
  const uint32_t ring_base = dyn_base + (uint32_t)(T * nh * halo_bytes);
  const uint32_t stage_bytes = (uint32_t)G * b_bytes;
  const int pitch = HALO_TW + 2 * pad;
  const int tiles_w = (ha.OW + HALO_TW - 1) / HALO_TW, tiles_h = (ha.OH + HALO_TH - 1) / HALO_TH;
  const int t_first = blockIdx.x * T;            // the launcher guarantees n_tiles % T == 0
  const int n_iter = (nslabs + G - 1) / G;

  if (warp == 0) {
    if (lane == 0) {
      // this thread owns every barrier's initialisation AND the first requests, so the loads are in flight before the
      // rest of the CTA has finished its prologue (TMEM allocation, __syncthreads)
      for (int s = 0; s < stages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), 1);
      }
      mbar_init(smem_u32(&halo_bar), 1);
      mbar_init(smem_u32(&accum_bar), 1);
      fence_barrier_init();
      pdl_wait();      // halo windows (and, in single-layer calls, the packed weights) come from earlier kernels
      const uint32_t hb = smem_u32(&halo_bar);
      mbar_expect_tx(hb, (uint32_t)(T * nh) * (uint32_t)((HALO_TH + 2 * pad) * pitch * 128));
#pragma unroll
      for (int t = 0; t < T; ++t) {
        int ti = t_first + t;
        const int twi = ti % tiles_w; ti /= tiles_w;
        const int thi = ti % tiles_h;
        const int b = ti / tiles_h;
        for (int h = 0; h < nh; ++h)
          tma_load_4d(dyn_base + (uint32_t)((t * nh + h) * halo_bytes), &maps.halo[ha.src[h]], hb, ha.c0[h],
                      twi * HALO_TW - pad, thi * HALO_TH - pad, b);
      }
      {  // first weight stage
        const int cnt = min(G, nslabs);
        const uint32_t fb = smem_u32(&full_bar[0]);
        mbar_expect_tx(fb, (uint32_t)cnt * b_bytes);
        for (int q = 0; q < cnt; ++q) tma_load_2d(ring_base + (uint32_t)q * b_bytes, &maps.w, fb, 0, q * Npad);
      }
    }
    __syncwarp();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)ha.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      uint32_t st = (stages > 1) ? 1u : 0u, ph = (stages > 1) ? 1u : 0u;      // stage 0 was requested in the prologue
      for (int it = 1; it < n_iter; ++it) {
        mbar_wait(empty0 + 8u * st, ph);
        const int s0 = it * G;
        const int cnt = min(G, nslabs - s0);
        const uint32_t fb = full0 + 8u * st;
        mbar_expect_tx(fb, (uint32_t)cnt * b_bytes);
        for (int q = 0; q < cnt; ++q)
          tma_load_2d(ring_base + st * stage_bytes + (uint32_t)q * b_bytes, &maps.w, fb, 0, (s0 + q) * Npad);
        if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  Everything in this loop is warp-uniform (kernel parameters, shuffled bases, loop counters) so
    // that ptxas keeps descriptors and barrier addresses in uniform registers: ~10 instructions per slab of 4*T MMAs.
    const uint32_t idesc = make_idesc(128, Npad, 0, 0);
    const uint32_t tm = uniform32(tmem_base);
    const uint32_t a_hi = (uint32_t)(make_sdesc(0, 16, (uint32_t)pitch * 128u) >> 32);
    const uint32_t b_hi = (uint32_t)(make_sdesc(0, 16, 1024) >> 32);
    const uint32_t a_lo0 = uniform32(((dyn_base >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t b_lo0 = uniform32(((ring_base >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t full0 = uniform32(smem_u32(&full_bar[0])), empty0 = uniform32(smem_u32(&empty_bar[0]));
    const uint32_t accb = uniform32(smem_u32(&accum_bar));
    const uint32_t tstep = (uint32_t)((nh * halo_bytes) >> 4), bstep = b_bytes >> 4, sstep = stage_bytes >> 4;
    mbar_wait_warp(uniform32(smem_u32(&halo_bar)), 0, 0);
    uint32_t st = 0, ph = 0;
    int s = 0;
#pragma unroll 1
    for (int it = 0; it < n_iter; ++it) {
      mbar_wait_warp(full0 + 8u * st, ph, 0);
      tc_fence_after();
      uint32_t b_lo = b_lo0 + st * sstep;
      const int s_end = min(s + G, nslabs);
      if (elect_one()) {
#pragma unroll 1
        for (; s < s_end; ++s, b_lo += bstep) {
          const uint32_t a_lo = a_lo0 + (uint32_t)ha.aoff[s];
          const uint32_t acc0 = (s > 0) ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int t = 0; t < T; ++t)      // k outer, tile inner: consecutive MMAs hit different accumulators
              umma_bf16(tm + (uint32_t)(t * Npad), ((uint64_t)a_hi << 32) | (uint64_t)(a_lo + (uint32_t)t * tstep + 2u * k),
                        ((uint64_t)b_hi << 32) | (uint64_t)(b_lo + 2u * k), idesc, k > 0 ? 1u : acc0);
        }
        umma_commit(empty0 + 8u * st);
        if (it == n_iter - 1) umma_commit(accb);
      }
      s = s_end;
      __syncwarp();
      if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    // While the accumulator is being computed: stage the bias in shared memory (parameters, never written by a kernel
    // of this step) and pull this thread's rows of the residual / mask tensors towards the SM - otherwise the epilogue
    // pays two or three dependent L2 / HBM round trips after the last MMA has retired
    for (int i = (int)threadIdx.x - 64; i < Npad; i += 128) bias_s[i] = (epi.bias && i < ha.N) ? __ldg(epi.bias + i) : 0.f;
    pdl_wait();     // the epilogue reads residual / mask tensors and overwrites buffers earlier kernels may still read
    {
      int ti = t_first;
      const int twi = ti % tiles_w; ti /= tiles_w;
      const int thi = ti % tiles_h;
      const int b = ti / tiles_h;
      const int oh = thi * HALO_TH + (row >> 3), ow = twi * HALO_TW + (row & 7);
      if (oh < ha.OH) {
        if (epi.add) asm volatile("prefetch.global.L1 [%0];" ::"l"(epi.add + b * epi.aB + oh * epi.aH + ow * epi.aW));
        if (epi.mask) asm volatile("prefetch.global.L1 [%0];" ::"l"(epi.mask + b * epi.mB + oh * epi.mH + ow * epi.mW));
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");      // bias_s visible to the four epilogue warps
    mbar_wait_warp(smem_u32(&accum_bar), 0, 100u);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < T; ++t) {
      int ti = t_first + t;
      const int twi = ti % tiles_w; ti /= tiles_w;
      const int thi = ti % tiles_h;
      const int b = ti / tiles_h;
      const int oh = thi * HALO_TH + (row >> 3), ow = twi * HALO_TW + (row & 7);
      const bool ok = oh < ha.OH && ow < ha.OW;      // partial tiles: loads were zero-filled, stores are skipped
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * Npad);
      int n0 = 0;
      for (; n0 + 32 <= Npad; n0 += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)n0, v);
        if (ok) {
          epi_apply16(epi, b, oh, ow, n0, ha.N, v, bias_s);
          epi_apply16(epi, b, oh, ow, n0 + 16, ha.N, v + 16, bias_s);
        }
      }
      if (n0 < Npad) {
        float v[16];
        tmem_ld16(trow + (uint32_t)n0, v);
        if (ok) epi_apply16(epi, b, oh, ow, n0, ha.N, v, bias_s);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)ha.tmem_cols);
}

// ---------------------------------------------------------------------------------------------
// host side: eligibility, TMA descriptors, launch
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || !p) {
    ss_set_error("cuTensorMapEncodeTiled not available from the driver: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  fn = (EncodeTiledFn)p;
  return fn;
}

// generic 4-D tiled map (used for the fp32 NCHW tile stores of the sigmoid head)
int ss_tma_encode_4d(CUtensorMap* out, int fp32, void* base, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle128) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return SSHSLIE_ERR_CUDA;
  cuuint64_t d[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t s[3] = {strides_bytes[0], strides_bytes[1], strides_bytes[2]};
  cuuint32_t b[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, d, s, b, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled(4d) failed with CUresult %d", (int)r);
    return SSHSLIE_ERR_CUDA;
  }
  return SSHSLIE_OK;
}

int ss_umma_supported(const ConvGeom& g) {
  if (g.Npad < 16 || g.Npad > 256 || (g.Npad % 16)) return 0;
  if (g.tw < 8 || g.tw * g.th != 128) return 0;
  for (int s = 0; s < g.nsrc; ++s) {
    const SrcView& v = g.src[s];
    if (((uintptr_t)v.base & 15) || (v.sW % 8) || (v.sH % 8) || (v.sB % 8)) return 0;
  }
  return 1;
}

int ss_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}
#define HALO_SMEM_BUDGET (200 * 1024)
static int halo_args(const ConvGeom& g, HaloArgs* out) {
  HaloArgs ha;
  memset(&ha, 0, sizeof(ha));
  int pad = 0;
  for (int i = 0; i < g.nslabs; ++i) {
    const Slab& sl = g.slab[i];
    pad = std::max(pad, std::max(abs((int)sl.dh), abs((int)sl.dw)));
    int h = -1;
    for (int j = 0; j < ha.nh; ++j)
      if (ha.src[j] == sl.src && ha.c0[j] == sl.c0) h = j;
    if (h < 0) {
      if (ha.nh == SS_MAX_WIN) return 0;
      ha.src[ha.nh] = sl.src; ha.c0[ha.nh] = sl.c0; ++ha.nh;
    }
  }
  if (pad > 4) return 0;
  ha.pad = pad;
  const int rows = (HALO_TH + 2 * pad) * (HALO_TW + 2 * pad);
  ha.halo_bytes = (rows * 128 + 1023) / 1024 * 1024;
  ha.n_tiles = g.B * ((g.OH + HALO_TH - 1) / HALO_TH) * ((g.OW + HALO_TW - 1) / HALO_TW);
  const int b_bytes = g.Npad * 128;
  // slabs per ring stage: <= 32 KB of weights (16*T MMAs per barrier round at Npad = 64), balanced over the rounds, and
  // small enough that halo + two stages stay under half an SM's shared memory (two co-resident CTAs when T = 1)
  int G = std::max(1, (32 * 1024) / b_bytes);
  while (G > 1 && ha.nh * ha.halo_bytes + 2 * G * b_bytes > 108 * 1024) --G;
  G = std::min(G, g.nslabs);
  {
    const int rounds = (g.nslabs + G - 1) / G;
    G = (g.nslabs + rounds - 1) / rounds;
  }
  G = std::max(1, std::min(16, ss_env_int("SSHSLIE_HALO_G", G)));
  ha.G = G;
  const int min_ring = 2 * G * b_bytes;
  // tiles per CTA: small problems keep T = 1 (two CTAs per SM overlap each other's prologue / epilogue); when there
  // are more than ~4 waves of tiles, T = 2 halves the weight traffic per MMA
  int T = 1;      // measured: two co-resident T = 1 CTAs per SM beat one T = 2 CTA at every size of this network
  T = std::max(1, std::min(HALO_MAX_T, ss_env_int("SSHSLIE_HALO_T", T)));
  if (T > 2) T = 2;
  while (T > 1 && (T * ha.nh * ha.halo_bytes + min_ring > HALO_SMEM_BUDGET || T * g.Npad > 512 || (ha.n_tiles % T))) --T;
  if (T * ha.nh * ha.halo_bytes + min_ring > HALO_SMEM_BUDGET || T * g.Npad > 512) return 0;
  ha.T = T;
  const int n_iter = (g.nslabs + G - 1) / G;
  // ring depth: what fits next to the halo windows; with T = 1 stay under half an SM so that two CTAs are co-resident
  int budget = (T == 1) ? std::max(T * ha.nh * ha.halo_bytes + min_ring, 108 * 1024) : HALO_SMEM_BUDGET;
  // when halo + two stages fit a THIRD of an SM, stop there: the next kernel's CTAs (programmatic dependent launch) can
  // then become resident and finish their prologue while this kernel's last CTAs drain
  if (T == 1 && ss_env_int("SSHSLIE_HALO_OCC3", 1) && ha.nh * ha.halo_bytes + min_ring <= 73 * 1024)
    budget = ha.nh * ha.halo_bytes + min_ring;
  int stages = (budget - T * ha.nh * ha.halo_bytes) / (G * b_bytes);
  stages = std::max(2, std::min(stages, std::min(HALO_MAX_STAGES, std::max(2, n_iter))));
  stages = std::max(2, std::min(HALO_MAX_STAGES, ss_env_int("SSHSLIE_HALO_STAGES", stages)));
  if (T * ha.nh * ha.halo_bytes + stages * G * b_bytes > 220 * 1024 - 2048) return 0;
  ha.stages = stages;
  ha.nslabs = g.nslabs; ha.Npad = g.Npad; ha.N = g.N; ha.OH = g.OH; ha.OW = g.OW;
  const int pitch = HALO_TW + 2 * pad;
  for (int i = 0; i < g.nslabs; ++i) {
    const Slab& sl = g.slab[i];
    int h = 0;
    for (int j = 0; j < ha.nh; ++j)
      if (ha.src[j] == sl.src && ha.c0[j] == sl.c0) h = j;
    ha.aoff[i] = (uint16_t)((h * ha.halo_bytes + ((sl.dh + pad) * pitch + (sl.dw + pad)) * 128) >> 4);
  }
  int cols = 32;
  while (cols < T * g.Npad) cols <<= 1;
  ha.tmem_cols = cols;
  *out = ha;
  return 1;
}
int ss_umma_halo_supported(const ConvGeom& g) {
  HaloArgs ha;
  return ss_umma_supported(g) && halo_args(g, &ha);
}

int ss_umma_encode_view(const SrcView& v, int ld_extent, int tw, int th, int B, CUtensorMap* out) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return SSHSLIE_ERR_CUDA;
  cuuint64_t dims[4] = {(cuuint64_t)ld_extent, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)v.sW * 2, (cuuint64_t)v.sH * 2, (cuuint64_t)v.sB * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)tw, (cuuint32_t)th, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)v.base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled(src) failed with CUresult %d (W=%d H=%d sW=%lld sH=%lld tw=%d th=%d)", (int)r,
                 v.W, v.H, (long long)v.sW, (long long)v.sH, tw, th);
    return SSHSLIE_ERR_CUDA;
  }
  return SSHSLIE_OK;
}

int ss_umma_build_maps(const ConvGeom& g, UmmaMaps* maps) {
  // channel extent of each view = the largest c0 + 64 any slab reads from it (all tensors are padded to that)
  for (int s = 0; s < g.nsrc; ++s) {
    int ext = 64;
    for (int i = 0; i < g.nslabs; ++i)
      if (g.slab[i].src == s && g.slab[i].c0 + 64 > ext) ext = g.slab[i].c0 + 64;
    int rc = ss_umma_encode_view(g.src[s], ext, g.tw, g.th, g.B, &maps->src[s]);
    if (rc) return rc;
    HaloArgs ha;
    if (halo_args(g, &ha)) {
      rc = ss_umma_encode_view(g.src[s], ext, HALO_TW + 2 * ha.pad, HALO_TH + 2 * ha.pad, g.B, &maps->halo[s]);
      if (rc) return rc;
    }
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return SSHSLIE_ERR_CUDA;
  // packed weights are slab-major: [nslabs][Npad][64] bf16, so one slab is ONE contiguous Npad x 128 B block
  const cuuint64_t Ktot = (cuuint64_t)g.nslabs * SS_SLAB;
  cuuint64_t dims[2] = {64, (cuuint64_t)g.Npad * g.nslabs};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)g.Npad};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(&maps->w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)g.wp, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ss_set_error("cuTensorMapEncodeTiled(weights) failed with CUresult %d (Ktot=%llu Npad=%d)", (int)r,
                 (unsigned long long)Ktot, g.Npad);
    return SSHSLIE_ERR_CUDA;
  }
  return SSHSLIE_OK;
}

int ss_launch_conv_gather_umma(const ConvGeom* g_dev, const ConvGeom& g, const UmmaMaps& maps, const Epi& epi,
                               cudaStream_t st) {
  const int tiles = g.B * ((g.OH + g.th - 1) / g.th) * ((g.OW + g.tw - 1) / g.tw);
  int cols = 32;
  while (cols < g.Npad) cols <<= 1;
  const size_t smem = (size_t)UM_STAGES * (UM_A_BYTES + (size_t)g.Npad * 128) + 1024;
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(conv_gather_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      ss_set_error("conv_gather_umma: cannot raise dynamic shared memory: %s",
                   cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(tiles);
  cfg.blockDim = dim3(UM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ss_pdl_enabled() ? 1 : 0;
  (void)g_dev;
  cudaLaunchKernelEx(&cfg, conv_gather_umma_kernel, g, maps, epi, cols);
  return ss_check_launch("conv_gather_umma");
}

// ---------------------------------------------------------------------------------------------
// weight-gradient GEMM on tcgen05:   dW[n][slab s, channel j] += sum_pixels G[pixel, n] * A_s[pixel, j]
//
// The reduction (K) axis is the pixel axis, which is the ROW axis of both NHWC operands, so both are fed
// MN-major: the same 128-pixel x 64-channel SWIZZLE_128B tiles the gather kernel uses, read "transposed" by the
// tensor core (instruction-descriptor major bits 15/16 = 1).  To fill M = 128 the A operand is a PAIR of slab
// tiles (two 64-channel atoms, LBO = one tile), the B operand is the G tile (N = 64 or 128 gradient channels):
//     D_pair[(slab 2p | 2p+1, j)][n] += A_pair^T . G            8 x (K=16) MMAs per 128-pixel tile
// One CTA owns a group of up to 512/N slab pairs (all TMEM columns) and a contiguous range of pixel tiles, keeps
// the accumulators in TMEM for its whole range, and finally adds them into the flat fp32 gradient buffer with
// red.global.add.f32 at the weight tensor's own (n, c, kh, kw) strides.
// ---------------------------------------------------------------------------------------------
#define WG_STAGES 4
#define WG_STAGE_BYTES (2 * UM_A_BYTES)
#define WG_G_BYTES (2 * UM_A_BYTES)
#define WG_ONES_BYTES UM_A_BYTES          // an all-ones 128 x 64 bf16 tile: ones^T . G = column sums = bias gradient

struct WgradArgs {
  int slabs_per_group;     // even
  int tiles_per_cta;
  int n_tiles;             // B * tiles_h * tiles_w
  int N;                   // UMMA N: 64 or 128 (G channels loaded)
  int gN;                  // valid G channels
  int tmem_cols;
  long long bias_off;      // >= 0: also produce db[n] = sum_pixels G[pixel, n] (group 0), added at grads + bias_off
  int groups, splits;      // grid = (splits, groups)
  int blocks_per_cta;      // accumulator blocks a CTA writes: slabs_per_group/2 pairs + 1 bias block
};

__global__ void __launch_bounds__(UM_THREADS, 1)
conv_wgrad_umma_kernel(const __grid_constant__ ConvGeom g, const __grid_constant__ UmmaMaps maps,
                       const __grid_constant__ CUtensorMap gmap, const __grid_constant__ WgradArgs wa,
                       float* __restrict__ partial) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t a_full[WG_STAGES];
  __shared__ __align__(8) uint64_t a_empty[WG_STAGES];
  __shared__ __align__(8) uint64_t g_full[2];
  __shared__ __align__(8) uint64_t g_empty[2];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t g_base = dyn_base + WG_STAGES * WG_STAGE_BYTES;
  const uint32_t ones_base = g_base + 2 * WG_G_BYTES;
  {
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_dyn + (dyn_base - smem_u32(smem_dyn)) +
                                                 WG_STAGES * WG_STAGE_BYTES + 2 * WG_G_BYTES);
    for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) ones[i] = 0x3F803F80u;   // bf16 1.0 pairs
    fence_proxy_async();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  __syncthreads();

  const int group = blockIdx.y;
  const bool do_bias = (wa.bias_off >= 0) && (group == 0);
  const int s_begin = group * wa.slabs_per_group;
  const int s_end = min(g.nslabs, s_begin + wa.slabs_per_group);
  const int npairs = (s_end - s_begin + 1) / 2;
  const int t_begin = blockIdx.x * wa.tiles_per_cta;
  const int t_end = min(wa.n_tiles, t_begin + wa.tiles_per_cta);
  const int ntiles = t_end - t_begin;           // >= 1 by construction of the grid
  const int g_atoms = wa.N / 64;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nsrc; ++s) tma_prefetch_desc(&maps.src[s]);
    tma_prefetch_desc(&gmap);
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(smem_u32(&a_full[s]), 1);
      mbar_init(smem_u32(&a_empty[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&g_full[s]), 1);
      mbar_init(smem_u32(&g_empty[s]), 1);
    }
    mbar_init(smem_u32(&accum_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)wa.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const int tiles_w = (g.OW + g.tw - 1) / g.tw, tiles_h = (g.OH + g.th - 1) / g.th;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int twi = t % tiles_w; t /= tiles_w;
        const int thi = t % tiles_h;
        const int b = t / tiles_h;
        const int oh0 = thi * g.th, ow0 = twi * g.tw;
        {  // G tile of this pixel tile
          const int gs = ti & 1;
          mbar_wait(smem_u32(&g_empty[gs]), (((uint32_t)(ti >> 1)) & 1u) ^ 1u);
          const uint32_t fb = smem_u32(&g_full[gs]);
          mbar_expect_tx(fb, (uint32_t)g_atoms * UM_A_BYTES);
          for (int a = 0; a < g_atoms; ++a)
            tma_load_4d(g_base + gs * WG_G_BYTES + a * UM_A_BYTES, &gmap, fb, a * 64, ow0, oh0, b);
        }
        for (int p = 0; p < npairs; ++p, ++it) {
          const int st = it % WG_STAGES;
          mbar_wait(smem_u32(&a_empty[st]), (((uint32_t)(it / WG_STAGES)) & 1u) ^ 1u);
          const uint32_t fb = smem_u32(&a_full[st]);
          const int s0 = s_begin + 2 * p;
          const int cnt = (s0 + 1 < s_end) ? 2 : 1;
          mbar_expect_tx(fb, (uint32_t)cnt * UM_A_BYTES);
          for (int q = 0; q < cnt; ++q) {
            const Slab sl = g.slab[s0 + q];
            tma_load_4d(dyn_base + st * WG_STAGE_BYTES + q * UM_A_BYTES, &maps.src[sl.src], fb, sl.c0, ow0 + sl.dw,
                        oh0 + sl.dh, b);
          }
        }
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc(128, wa.N, 1, 1);
    const uint32_t tm = uniform32(tmem_base), base = uniform32(dyn_base), gb = uniform32(g_base),
                   ob = uniform32(ones_base);
    int it = 0;
    for (int ti = 0; ti < ntiles; ++ti) {
      const int gs = ti & 1;
      mbar_wait_warp(smem_u32(&g_full[gs]), ((uint32_t)(ti >> 1)) & 1u, 0);
      const uint64_t bd0 = make_sdesc(gb + gs * WG_G_BYTES, UM_A_BYTES, 1024);       // 1 or 2 G atoms
      for (int p = 0; p < npairs; ++p, ++it) {
        const int st = it % WG_STAGES;
        mbar_wait_warp(smem_u32(&a_full[st]), ((uint32_t)(it / WG_STAGES)) & 1u, 0);
        tc_fence_after();
        const uint64_t ad0 = make_sdesc(base + st * WG_STAGE_BYTES, UM_A_BYTES, 1024);  // 2 slab atoms, LBO = one tile
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 8; ++k)      // K step = 16 pixel rows = 2048 B = 128 descriptor units
            umma_bf16(tm + (uint32_t)(p * wa.N), ad0 + (uint64_t)(128 * k), bd0 + (uint64_t)(128 * k), idesc,
                      (ti > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&a_empty[st]));
        }
        __syncwarp();
      }
      if (elect_one()) {
        if (do_bias) {
          const uint64_t od0 = make_sdesc(ob, 0, 1024);                                // both M atoms = the ones tile
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_bf16(tm + (uint32_t)(npairs * wa.N), od0 + (uint64_t)(128 * k), bd0 + (uint64_t)(128 * k), idesc,
                      (ti > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&g_empty[gs]));
        if (ti == ntiles - 1) umma_commit(smem_u32(&accum_bar));
      }
      __syncwarp();
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait_warp(smem_u32(&accum_bar), 0, 200);
    tc_fence_after();
    // partial[split][group][block][n][row]: lanes = consecutive rows -> 128-byte coalesced stores, no atomics
    float* out = partial + ((size_t)(blockIdx.x * wa.groups + group) * wa.blocks_per_cta) * (size_t)wa.N * 128;
    const int nblocks = npairs + (do_bias ? 1 : 0);
    for (int p = 0; p < nblocks; ++p) {
      for (int n0 = 0; n0 < wa.N; n0 += 16) {
        float v[16];
        tmem_ld16(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p * wa.N + n0), v);
        float* o = out + ((size_t)p * wa.N + n0) * 128 + row;
#pragma unroll
        for (int i = 0; i < 16; ++i) o[(size_t)i * 128] = v[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)wa.tmem_cols);
}

int ss_launch_conv_gather_umma4(const ConvGeom* g_dev, const ConvGeom* g4, const UmmaMaps* maps4, const Epi& epi,
                                cudaStream_t st) {
  const ConvGeom& g = g4[0];
  const int tiles = g.B * ((g.OH + g.th - 1) / g.th) * ((g.OW + g.tw - 1) / g.tw);
  int cols = 32;
  while (cols < g.Npad) cols <<= 1;
  const size_t smem = (size_t)UM_STAGES * (UM_A_BYTES + (size_t)g.Npad * 128) + 1024;
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(conv_gather_umma4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      ss_set_error("conv_gather_umma4: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  UmmaMaps4 m4;
  ConvGeom4 gg;
  for (int q = 0; q < 4; ++q) { m4.m[q] = maps4[q]; gg.g[q] = g4[q]; }
  (void)g_dev;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(tiles, 4);
  cfg.blockDim = dim3(UM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ss_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, conv_gather_umma4_kernel, gg, m4, epi, cols);
  ss_count_launches(0);
  return ss_check_launch("conv_gather_umma4");
}

int ss_launch_conv_gather_halo(const ConvGeom* g_dev, const ConvGeom& g, const UmmaMaps& maps, const Epi& epi,
                               cudaStream_t st) {
  (void)g_dev;
  HaloArgs ha;
  if (!halo_args(g, &ha)) {
    ss_set_error("conv_gather_halo: geometry not eligible");
    return SSHSLIE_ERR_ARG;
  }
  const size_t smem = (size_t)ha.T * ha.nh * ha.halo_bytes + (size_t)ha.stages * ha.G * g.Npad * 128 + 1024;
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(conv_gather_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
            cudaSuccess ||
        cudaFuncSetAttribute(conv_gather_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
            cudaSuccess) {
      ss_set_error("conv_gather_halo: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(ha.n_tiles / ha.T);
  cfg.blockDim = dim3(UM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ss_pdl_enabled() ? 1 : 0;
  if (ha.T == 2) cudaLaunchKernelEx(&cfg, conv_gather_halo_kernel<2>, maps, epi, ha);
  else cudaLaunchKernelEx(&cfg, conv_gather_halo_kernel<1>, maps, epi, ha);
  return ss_check_launch("conv_gather_halo");
}

int ss_umma_wgrad_supported(const ConvGeom& g) { return ss_umma_supported(g) && g.N <= 128; }
// profiling aid: 0 = both kernels of a tcgen05 weight gradient (default), 1 = the GEMM kernel only, 2 = the split-K
// reduce only (sshslie_profile_step times them as separate rows)
static thread_local int g_wgrad_part = 0;
void ss_set_wgrad_part(int part) { g_wgrad_part = part; }

int ss_umma_build_gmap(const bf16* G, int64_t gB, int64_t gH, int64_t gW, int ld_extent, const ConvGeom& g,
                       void* out_map) {
  SrcView v;
  v.base = G; v.sB = gB; v.sH = gH; v.sW = gW; v.H = g.OH; v.W = g.OW;
  return ss_umma_encode_view(v, ld_extent, g.tw, g.th, g.B, reinterpret_cast<CUtensorMap*>(out_map));
}

// second stage: dW[n][slab, j] += sum over pixel splits of the partial accumulators (deterministic, no atomics)
// grid = (blocks_per_cta, groups, N/8); block = (128, 8): threadIdx.x = accumulator row, threadIdx.y strides the splits.
// Each thread sums every 8th split for 8 columns, the 8 partial sums per column meet in shared memory, and thread
// (row, q) finishes column n0 + q — 8x the parallelism and 1/8 the dependent-load depth of one thread per row.
__global__ void __launch_bounds__(1024) conv_wgrad_reduce_kernel(const ConvGeom* __restrict__ gp,
                                                                 const float* __restrict__ partial, WgradArgs wa,
                                                                 float* __restrict__ grads) {
  __shared__ float red[8][8][128];
  const int blk = blockIdx.x, group = blockIdx.y, n0 = blockIdx.z * 8;
  const int row = threadIdx.x, q = threadIdx.y;
  const int s_begin = group * wa.slabs_per_group;
  const int s_end = min(gp->nslabs, s_begin + wa.slabs_per_group);
  const int npairs = (s_end - s_begin + 1) / 2;
  const bool is_bias = (blk == npairs) && (group == 0) && (wa.bias_off >= 0);
  if (blk > npairs || (blk == npairs && !is_bias)) return;       // block-uniform
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const size_t cta_stride = (size_t)wa.groups * wa.blocks_per_cta * wa.N * 128;
  const float* p = partial + ((size_t)group * wa.blocks_per_cta + blk) * (size_t)wa.N * 128 + (size_t)n0 * 128 + row +
                   (size_t)q * cta_stride;
  // all loads of up to four splits in flight before their adds (the loop otherwise pays one L2 / HBM round trip per split)
  for (int sp = q; sp < wa.splits; sp += 32, p += 32 * cta_stride) {
    float v[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool valid = sp + 8 * u < wa.splits;
#pragma unroll
      for (int i = 0; i < 8; ++i) v[u][i] = valid ? __ldg(p + (size_t)u * 8 * cta_stride + (size_t)i * 128) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += v[u][i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[q][i][row] = acc[i];
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tot += red[j][q][row];              // fixed order -> deterministic
  const int n = n0 + q;
  if (is_bias) {
    if (row == 0 && n < wa.gN) grads[wa.bias_off + n] += tot;
    return;
  }
  const int s = s_begin + 2 * blk + (row >> 6);
  const int j = row & 63;
  if (s >= s_end) return;
  const Slab sl = gp->slab[s];
  if (j >= sl.wcn || sl.no_wgrad) return;
  if (n < gp->N && n < wa.gN) grads[gp->w_off + sl.woff + (int64_t)j * gp->w_sC + (int64_t)n * gp->w_sN] += tot;
}

static void wgrad_plan(const ConvGeom& g, int gN, long long bias_off, WgradArgs* out) {
  WgradArgs wa;
  wa.gN = gN;
  wa.bias_off = bias_off;
  wa.N = (gN > 64) ? 128 : 64;
  const int max_slabs = 2 * (512 / wa.N - 1);          // one accumulator block is reserved for the bias row
  const int ngroups = (g.nslabs + max_slabs - 1) / max_slabs;
  int per = (g.nslabs + ngroups - 1) / ngroups;
  per += per & 1;
  wa.slabs_per_group = per;
  wa.groups = (g.nslabs + per - 1) / per;
  wa.n_tiles = g.B * ((g.OH + g.th - 1) / g.th) * ((g.OW + g.tw - 1) / g.tw);
  // split the pixel axis until every CTA still owns ~4 pixel tiles or all SMs are covered
  int splits = (wa.n_tiles + 3) / 4;
  if (splits > 148 / wa.groups) splits = 148 / wa.groups;
  if (splits < 1) splits = 1;
  wa.tiles_per_cta = (wa.n_tiles + splits - 1) / splits;
  wa.splits = (wa.n_tiles + wa.tiles_per_cta - 1) / wa.tiles_per_cta;
  wa.blocks_per_cta = per / 2 + 1;
  int cols = 32;
  while (cols < wa.blocks_per_cta * wa.N) cols <<= 1;
  wa.tmem_cols = cols;
  *out = wa;
}

size_t ss_umma_wgrad_partial_floats(const ConvGeom& g, int gN) {
  WgradArgs wa;
  wgrad_plan(g, gN, 0, &wa);
  return (size_t)wa.splits * wa.groups * wa.blocks_per_cta * wa.N * 128;
}

int ss_launch_conv_wgrad_umma(const ConvGeom* g_dev, const ConvGeom& g, const UmmaMaps& maps, const void* gmap,
                              int gN, long long bias_off, float* partial, float* grads, cudaStream_t st) {
  WgradArgs wa;
  wgrad_plan(g, gN, bias_off, &wa);
  const size_t smem = (size_t)WG_STAGES * WG_STAGE_BYTES + 2 * WG_G_BYTES + WG_ONES_BYTES + 1024;
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(conv_wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
        cudaSuccess) {
      ss_set_error("conv_wgrad_umma: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  dim3 grid(wa.splits, wa.groups);
  int rc = SSHSLIE_OK;
  if (g_wgrad_part != 2) {
    conv_wgrad_umma_kernel<<<grid, UM_THREADS, smem, st>>>(g, maps, *reinterpret_cast<const CUtensorMap*>(gmap), wa,
                                                           partial);
    rc = ss_check_launch("conv_wgrad_umma");
    if (rc || g_wgrad_part == 1) return rc;
  }
  dim3 rgrid(wa.blocks_per_cta, wa.groups, wa.N / 8);
  conv_wgrad_reduce_kernel<<<rgrid, dim3(128, 8), 0, st>>>(g_dev, partial, wa, grads);
  return ss_check_launch("conv_wgrad_reduce");
}


// ---------------------------------------------------------------------------------------------
// weight gradient with HALO REUSE (stride-1 layers).  Same GEMM as conv_wgrad_umma_kernel — D[(tap pair, channel)][n]
// += A_pair^T . G over the pixels of a 16x8 tile, both operands MN-major — but the A operand of EVERY tap of the
// tile is read from ONE halo window (as in conv_gather_halo_kernel): per pixel tile a CTA loads the window(s) and the
// G tile once (65 KB for the 9x9 layer) and then issues all of its <= 7 tap pairs x 8 MMAs from shared memory, instead of
// streaming 32 KB per pair.  The two taps of a pair are the two 64-row halves of M: LBO = distance of their windows.
// One barrier round per pixel tile (56-64 MMAs); descriptor words come from the kernel's constant bank.
// ---------------------------------------------------------------------------------------------
#define WGH_MAX_PAIRS 48
#define WGH_ONES_BYTES 2048
#define WGH_ISSUERS 1                    // 2 = pairs split over two issuing warps: measured, no gain (see below)
#define WGH_MAX_GROUP_PAIRS 7          // 512 TMEM columns / N = 64, minus the bias block
struct WgHaloArgs {
  int nh;
  int src[SS_MAX_WIN];
  int c0[SS_MAX_WIN];
  int pad, halo_bytes;
  int N, gN, g_atoms;          // UMMA N (64 / 128), valid gradient channels, 64-channel atoms of G per tile
  int tmem_cols;
  long long bias_off;          // >= 0: group 0 also produces db (ones^T . G)
  int n_tiles, tiles_per_cta, splits, groups, pairs_per_group, npairs, blocks_per_cta;
  int stages, stage_bytes;
  int OH, OW;
  uint32_t pair_lo[WGH_MAX_PAIRS];   // (window offset of slab a) >> 4  |  ((offset of slab b - offset of slab a) >> 4) << 16
  uint8_t pair_a[WGH_MAX_PAIRS];     // slab indices of the two M halves (pair_b = 255: none)
  uint8_t pair_b[WGH_MAX_PAIRS];
  uint8_t ord[2 * WGH_MAX_PAIRS];    // per group (entries [2 p_begin, 2 p_end)): its slots 2 * (pair - p_begin) + half, sorted
                                     // by weight offset (the reduce kernel writes the gradient in memory order)
};

// body of the kernel (arguments in the constant bank): bx = pixel split, by = pair group
SS_DEVINL void wgrad_halo_body(const CUtensorMap* __restrict__ halo_maps, const CUtensorMap* __restrict__ gmap_p,
                               const WgHaloArgs& wa, float* __restrict__ partial, const int bx, const int by) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t full_bar[4];
  __shared__ __align__(8) uint64_t empty_bar[4];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const int nh = wa.nh, pad = wa.pad, stages = wa.stages;
  const uint32_t halo_bytes = (uint32_t)wa.halo_bytes, stage_bytes = (uint32_t)wa.stage_bytes;
  const uint32_t ones_base = dyn_base + (uint32_t)stages * stage_bytes;
  const int pitch = HALO_TW + 2 * pad;
  const int group = by;
  const bool do_bias = (wa.bias_off >= 0) && (group == 0);
  const int p_begin = group * wa.pairs_per_group;
  const int p_end = min(wa.npairs, p_begin + wa.pairs_per_group);
  const int t_begin = bx * wa.tiles_per_cta;
  const int t_end = min(wa.n_tiles, t_begin + wa.tiles_per_cta);
  const int ntiles = t_end - t_begin;           // >= 1 by construction of the grid
  const int tiles_w = (wa.OW + HALO_TW - 1) / HALO_TW, tiles_h = (wa.OH + HALO_TH - 1) / HALO_TH;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < stages; ++s) {
        mbar_init(smem_u32(&full_bar[s]), 1);
        mbar_init(smem_u32(&empty_bar[s]), WGH_ISSUERS);
      }
      mbar_init(smem_u32(&accum_bar), WGH_ISSUERS);
      fence_barrier_init();
    }
    __syncwarp();
  } else if (do_bias) {
    // an all-ones operand: two 8 x 128 B K-atoms (one MMA's K = 16); LBO = 0 makes both M atoms alias them and every
    // K step re-reads them
    uint32_t* ones = reinterpret_cast<uint32_t*>(smem_dyn + (ones_base - smem_u32(smem_dyn)));
    for (int i = threadIdx.x - 32; i < WGH_ONES_BYTES / 4; i += UM_THREADS - 32) ones[i] = 0x3F803F80u;   // bf16 1.0 pairs
    fence_proxy_async();       // generic-proxy writes -> visible to the tensor core's async-proxy reads
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), (uint32_t)wa.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t tx = (uint32_t)nh * (uint32_t)((HALO_TH + 2 * pad) * pitch * 128) + (uint32_t)wa.g_atoms * UM_A_BYTES;
      uint32_t st = 0, ph = 1;
      for (int ti = 0; ti < ntiles; ++ti) {
        int t = t_begin + ti;
        const int twi = t % tiles_w; t /= tiles_w;
        const int thi = t % tiles_h;
        const int b = t / tiles_h;
        mbar_wait(empty0 + 8u * st, ph);
        const uint32_t fb = full0 + 8u * st;
        const uint32_t dst = dyn_base + st * stage_bytes;
        mbar_expect_tx(fb, tx);
        for (int h = 0; h < nh; ++h)
          tma_load_4d(dst + (uint32_t)h * halo_bytes, &halo_maps[wa.src[h]], fb, wa.c0[h], twi * HALO_TW - pad,
                      thi * HALO_TH - pad, b);
        for (int a = 0; a < wa.g_atoms; ++a)
          tma_load_4d(dst + (uint32_t)nh * halo_bytes + (uint32_t)a * UM_A_BYTES, gmap_p, fb, a * 64, twi * HALO_TW,
                      thi * HALO_TH, b);
        if (++st == (uint32_t)stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1 || warp == WGH_ISSUERS) {
    // Measured on B200 (cycle breadcrumbs, round 1): this loop runs at ~50 clk per M128 N64 K16 MMA, the shared-memory
    // operand-read limit of SS-mode MMAs at N = 64 (tools/umma_probe.py: 48 clk, same for K- and MN-major operands).
    // Neither issuing from two warps (WGH_ISSUERS = 2) nor interleaving accumulators changed that.
    const int issuer = (warp == 1) ? 0 : 1;
    // (every loop-invariant word goes through a shuffle once, so that ptxas can keep descriptors in uniform registers
    // instead of wrapping each UTCHMMA in an R2UR waterfall)
    const uint32_t idesc = uniform32(make_idesc(128, wa.N, 1, 1));
    const uint32_t tm = uniform32(tmem_base);
    // A: two 64-channel atoms (the two taps), LBO per pair; K groups of 8 pixels = tile rows, SBO = pitch * 128 B
    const uint32_t a_hi = uniform32((uint32_t)(make_sdesc(0, 0, (uint32_t)pitch * 128u) >> 32));
    const uint32_t b_hi = (uint32_t)(make_sdesc(0, 0, 1024) >> 32);
    const uint32_t base_lo = uniform32((dyn_base >> 4) & 0x3FFFu);
    const uint32_t g_rel = uniform32((uint32_t)((nh * halo_bytes) >> 4) | ((uint32_t)(UM_A_BYTES >> 4) << 16));   // G atoms: LBO = one tile
    const uint32_t ones_lo = uniform32((ones_base >> 4) & 0x3FFFu);
    const uint32_t full0 = uniform32(smem_u32(&full_bar[0])), empty0 = uniform32(smem_u32(&empty_bar[0]));
    const uint32_t accb = uniform32(smem_u32(&accum_bar));
    const uint32_t sstep = uniform32(stage_bytes >> 4);
    const uint32_t kstep = uniform32((uint32_t)(2 * pitch * 128) >> 4);      // 16 pixels = two tile rows
    const uint32_t N = uniform32((uint32_t)wa.N);
    const uint32_t nstages = uniform32((uint32_t)stages);
    const int n_t = (int)uniform32((uint32_t)ntiles);
    const bool bias_u = uniform32(do_bias ? 1u : 0u) != 0u;
    // the group's pair words never change: keep them in (uniform) registers.  MMAs are issued k-outer / pair-inner so
    // that consecutive MMAs accumulate into different TMEM tiles
    const int np = (int)uniform32((uint32_t)(p_end - p_begin));
    uint32_t plo[WGH_MAX_GROUP_PAIRS];
#pragma unroll
    for (int p = 0; p < WGH_MAX_GROUP_PAIRS; ++p) plo[p] = uniform32(wa.pair_lo[min(p_begin + p, WGH_MAX_PAIRS - 1)]);
    uint32_t st = 0, ph = 0;
#pragma unroll 1
    for (int ti = 0; ti < n_t; ++ti) {
      mbar_wait_warp(full0 + 8u * st, ph, 0);
      tc_fence_after();
      const uint32_t s_lo = base_lo + st * sstep;
      const uint32_t acc0 = (ti > 0) ? 1u : 0u;
      if (elect_one()) {
        const uint64_t bd0 = ((uint64_t)b_hi << 32) | (uint64_t)(s_lo + g_rel);
        const uint64_t od0 = ((uint64_t)b_hi << 32) | (uint64_t)ones_lo;        // LBO = 0: both M atoms alias; same K atoms each step
#pragma unroll 1      // (fully unrolled, the 64 descriptor pairs overflow the uniform register file: spills in the issue loop)
        for (int k = 0; k < 8; ++k) {
          const uint32_t acc = k > 0 ? 1u : acc0;
#pragma unroll
          for (int p = 0; p < WGH_MAX_GROUP_PAIRS; ++p)
            if (p < np && (p % WGH_ISSUERS) == issuer)
              umma_bf16(tm + (uint32_t)p * N, ((uint64_t)a_hi << 32) | (uint64_t)(s_lo + plo[p] + (uint32_t)k * kstep),
                        bd0 + (uint64_t)(128 * k), idesc, acc);
          if (bias_u && issuer == WGH_ISSUERS - 1) umma_bf16(tm + (uint32_t)np * N, od0, bd0 + (uint64_t)(128 * k), idesc, acc);
        }
        umma_commit(empty0 + 8u * st);
        if (ti == n_t - 1) umma_commit(accb);
      }
      __syncwarp();
      if (++st == nstages) { st = 0; ph ^= 1u; }
    }
  }
  if (warp >= 2) {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    mbar_wait_warp(smem_u32(&accum_bar), 0, 200);
    tc_fence_after();
    // partial[split][group][block][n][row]: lanes = consecutive rows -> 128-byte coalesced stores, no atomics
    float* out = partial + ((size_t)(bx * wa.groups + group) * wa.blocks_per_cta) * (size_t)wa.N * 128;
    const int nblocks = (p_end - p_begin) + (do_bias ? 1 : 0);
    for (int p = 0; p < nblocks; ++p) {
      for (int n0 = 0; n0 < wa.N; n0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(p * wa.N + n0), v);
        float* o = out + ((size_t)p * wa.N + n0) * 128 + row;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[(size_t)i * 128] = v[i];
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)wa.tmem_cols);
}

__global__ void __launch_bounds__(UM_THREADS, 1)
conv_wgrad_halo_kernel(const __grid_constant__ UmmaMaps maps, const __grid_constant__ CUtensorMap gmap,
                       const __grid_constant__ WgHaloArgs wa, float* __restrict__ partial) {
  wgrad_halo_body(maps.halo, &gmap, wa, partial, (int)blockIdx.x, (int)blockIdx.y);
}

// second stage for the halo kernel: same partial layout as conv_wgrad_reduce_kernel, slab mapping through the pair table.
// One block per (output column n, pair group), one thread per (accumulator block, row): it adds the `splits` partials of
// its element in a fixed order (eight independent loads in flight) and parks the sum in shared memory; the block then
// adds its sums to the weight-gradient tensor IN MEMORY ORDER - for one (n, input channel) the taps of a group are
// adjacent words (the host sorts the group's slabs by weight offset), so a warp updates runs of consecutive words.
// (Writing straight from the accumulator layout - consecutive threads = consecutive input channels, K*K words apart -
//  made every thread's read-modify-write its own 32-byte sector: 19 of the 9x9 layer's 22 us.)
#define WGR_PITCH 65
__global__ void __launch_bounds__(1024) conv_wgrad_halo_reduce_kernel(const ConvGeom* __restrict__ gp,
                                                                      const float* __restrict__ partial,
                                                                      const __grid_constant__ WgHaloArgs wa,
                                                                      float* __restrict__ grads) {
  __shared__ float out[16 * WGR_PITCH];            // [slot t = 2 * block + M half][64 channels], pitch 65: conflict-free
  const int n = blockIdx.x, group = blockIdx.y;
  const int p_begin = group * wa.pairs_per_group;
  const int p_end = min(wa.npairs, p_begin + wa.pairs_per_group);
  const int np = p_end - p_begin;
  const int blk = threadIdx.x >> 7, row = threadIdx.x & 127;
  const bool is_bias = (blk == np) && (group == 0) && (wa.bias_off >= 0) && row == 0;
  if (blk < np || is_bias) {
    const size_t cta_stride = (size_t)wa.groups * wa.blocks_per_cta * wa.N * 128;
    const float* p = partial + ((size_t)group * wa.blocks_per_cta + blk) * (size_t)wa.N * 128 + (size_t)n * 128 + row;
    float tot = 0.f;
    int sp = 0;
    for (; sp + 8 <= wa.splits; sp += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(p + (size_t)(sp + u) * cta_stride);
#pragma unroll
      for (int u = 0; u < 8; ++u) tot += v[u];                     // fixed order -> deterministic
    }
    for (; sp < wa.splits; ++sp) tot += __ldg(p + (size_t)sp * cta_stride);
    if (is_bias) {
      if (n < wa.gN) grads[wa.bias_off + n] += tot;
    } else {
      out[(2 * blk + (row >> 6)) * WGR_PITCH + (row & 63)] = tot;
    }
  }
  __syncthreads();
  if (n >= gp->N || n >= wa.gN) return;
  const int nslots = 2 * np;
  for (int i = threadIdx.x; i < nslots * 64; i += 1024) {
    const int j = i / nslots, q = i - j * nslots;
    const int t = wa.ord[2 * p_begin + q];                          // slot of the q-th lowest weight offset of the group
    const int pr = p_begin + (t >> 1);
    const int sidx = (t & 1) ? (int)wa.pair_b[pr] : (int)wa.pair_a[pr];
    if (sidx == 255) continue;
    const Slab sl = gp->slab[sidx];
    if (j >= sl.wcn) continue;
    grads[gp->w_off + sl.woff + (int64_t)j * gp->w_sC + (int64_t)n * gp->w_sN] += out[t * WGR_PITCH + j];
  }
}

static int wgrad_halo_plan(const ConvGeom& g, int gN, long long bias_off, WgHaloArgs* out) {
  HaloArgs ha;
  if (g.halo_ok != 1 || !ss_umma_supported(g) || g.N > 128 || !halo_args(g, &ha)) return 0;   // 2 = gather kernels only
  WgHaloArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.nh = ha.nh;
  for (int i = 0; i < SS_MAX_WIN; ++i) { wa.src[i] = ha.src[i]; wa.c0[i] = ha.c0[i]; }
  wa.pad = ha.pad; wa.halo_bytes = ha.halo_bytes;
  wa.gN = gN; wa.bias_off = bias_off;
  wa.N = (gN > 64) ? 128 : 64;
  wa.g_atoms = wa.N / 64;
  wa.OH = g.OH; wa.OW = g.OW;
  // tap pairs over the slabs that carry weights of their own (residual "lo" slabs are skipped), lower window first
  int slabs[SS_MAX_SLABS], ns = 0;
  for (int i = 0; i < g.nslabs; ++i)
    if (!g.slab[i].no_wgrad) slabs[ns++] = i;
  wa.npairs = (ns + 1) / 2;
  if (wa.npairs > WGH_MAX_PAIRS) return 0;
  for (int p = 0; p < wa.npairs; ++p) {
    int a = slabs[2 * p], b = (2 * p + 1 < ns) ? slabs[2 * p + 1] : -1;
    if (b >= 0 && ha.aoff[b] < ha.aoff[a]) std::swap(a, b);
    wa.pair_a[p] = (uint8_t)a;
    wa.pair_b[p] = (uint8_t)(b >= 0 ? b : 255);
    const uint32_t lbo = (b >= 0) ? (uint32_t)(ha.aoff[b] - ha.aoff[a]) : 0u;
    if (lbo > 0x3FFFu) return 0;
    wa.pair_lo[p] = (uint32_t)ha.aoff[a] | (lbo << 16);
  }
  // TMEM budget of a CTA: all 512 columns (7 pair blocks + the bias block at N = 64).  A smaller budget would let a
  // forward / dgrad CTA of the main stream allocate tensor memory on the same SM, but it doubles the groups (and the halo
  // loads) of every small layer: measured at the training batch, 256 columns cost 2.6 % of the step (SSHSLIE_WGH_TMEM).
  int tmem_cap = ss_env_int("SSHSLIE_WGH_TMEM", 512);
  if (tmem_cap != 128 && tmem_cap != 256 && tmem_cap != 512) tmem_cap = 512;
  if (tmem_cap / wa.N < 2) tmem_cap = 2 * wa.N;
  const int max_pairs = tmem_cap / wa.N - 1;            // one accumulator block is reserved for the bias row
  wa.groups = (wa.npairs + max_pairs - 1) / max_pairs;
  wa.pairs_per_group = (wa.npairs + wa.groups - 1) / wa.groups;
  wa.groups = (wa.npairs + wa.pairs_per_group - 1) / wa.pairs_per_group;
  wa.blocks_per_cta = wa.pairs_per_group + 1;
  // per group: its accumulator slots (2 * pair + M half) in the order of their weight offsets
  for (int gidx = 0; gidx < wa.groups; ++gidx) {
    const int p_begin = gidx * wa.pairs_per_group, p_end = std::min(wa.npairs, p_begin + wa.pairs_per_group);
    const int nslots = 2 * (p_end - p_begin);
    std::pair<long long, int> key[2 * WGH_MAX_PAIRS];
    for (int t = 0; t < nslots; ++t) {
      const int pr = p_begin + (t >> 1);
      const int sidx = (t & 1) ? (int)wa.pair_b[pr] : (int)wa.pair_a[pr];
      key[t] = std::make_pair(sidx == 255 ? (1LL << 60) : (long long)g.slab[sidx].woff, t);
    }
    std::sort(key, key + nslots);
    for (int t = 0; t < nslots; ++t) wa.ord[2 * p_begin + t] = (uint8_t)key[t].second;
  }
  wa.n_tiles = ha.n_tiles;
  // split the pixel axis: every CTA should own >= 4 pixel tiles (the split-K partials cost 32 KB x blocks per CTA)
  const int min_tiles = std::max(1, ss_env_int("SSHSLIE_WGH_MIN_TILES", 8));
  int splits = (wa.n_tiles + min_tiles - 1) / min_tiles;
  splits = std::min(splits, std::max(1, 148 / wa.groups));
  splits = std::max(1, ss_env_int("SSHSLIE_WGH_SPLITS", splits));
  wa.tiles_per_cta = (wa.n_tiles + splits - 1) / splits;
  wa.splits = (wa.n_tiles + wa.tiles_per_cta - 1) / wa.tiles_per_cta;
  int cols = 32;
  while (cols < wa.blocks_per_cta * wa.N) cols <<= 1;
  wa.tmem_cols = cols;
  wa.stage_bytes = wa.nh * wa.halo_bytes + wa.g_atoms * UM_A_BYTES;
  // two stages when that keeps the CTA under half an SM's shared memory (a forward / dgrad CTA of the main stream can
  // then share the SM with a weight-gradient CTA of the side stream), else whatever fits, at most three
  int stages = (int)((216 * 1024 - WGH_ONES_BYTES) / wa.stage_bytes);
  stages = std::min(3, stages);
  if (2 * wa.stage_bytes + WGH_ONES_BYTES <= 108 * 1024) stages = 2;
  if (stages < 2 || stages * wa.stage_bytes + WGH_ONES_BYTES > 216 * 1024) return 0;
  wa.stages = stages;
  *out = wa;
  return 1;
}
int ss_umma_wgrad_halo_supported(const ConvGeom& g, int gN) {
  WgHaloArgs wa;
  return wgrad_halo_plan(g, gN, -1, &wa);
}
size_t ss_umma_wgrad_halo_partial_floats(const ConvGeom& g, int gN) {
  WgHaloArgs wa;
  if (!wgrad_halo_plan(g, gN, 0, &wa)) return 0;
  return (size_t)wa.splits * wa.groups * wa.blocks_per_cta * wa.N * 128;
}
int ss_umma_build_gmap_halo(const bf16* G, int64_t gB, int64_t gH, int64_t gW, int ld_extent, const ConvGeom& g,
                            void* out_map) {
  SrcView v;
  v.base = G; v.sB = gB; v.sH = gH; v.sW = gW; v.H = g.OH; v.W = g.OW;
  return ss_umma_encode_view(v, ld_extent, HALO_TW, HALO_TH, g.B, reinterpret_cast<CUtensorMap*>(out_map));
}
int ss_launch_conv_wgrad_halo(const ConvGeom* g_dev, const ConvGeom& g, const UmmaMaps& maps, const void* gmap, int gN,
                              long long bias_off, float* partial, float* grads, cudaStream_t st) {
  WgHaloArgs wa;
  if (!wgrad_halo_plan(g, gN, bias_off, &wa)) {
    ss_set_error("conv_wgrad_halo: geometry not eligible");
    return SSHSLIE_ERR_ARG;
  }
  size_t smem = (size_t)wa.stages * wa.stage_bytes + WGH_ONES_BYTES + 1024;
  // A weight-gradient CTA that owns all 512 TMEM columns must not share its SM with a forward / dgrad CTA of the main
  // stream: the block scheduler does not know about tensor memory, the co-located CTA would sit in tcgen05.alloc until
  // this one exits (up to 25 us added to a 7 us kernel of the critical path).  Asking for more shared memory than a
  // gather CTA leaves free keeps them apart.
  if (wa.tmem_cols > 256) smem = std::max(smem, (size_t)ss_env_int("SSHSLIE_WGH_SMEM_KB", 160) * 1024);
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
        cudaSuccess) {
      ss_set_error("conv_wgrad_halo: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  dim3 grid(wa.splits, wa.groups);
  int rc = SSHSLIE_OK;
  if (g_wgrad_part != 2) {
    conv_wgrad_halo_kernel<<<grid, UM_THREADS, smem, st>>>(maps, *reinterpret_cast<const CUtensorMap*>(gmap), wa, partial);
    rc = ss_check_launch("conv_wgrad_halo");
    if (rc || g_wgrad_part == 1) return rc;
  }
  conv_wgrad_halo_reduce_kernel<<<dim3(wa.N, wa.groups), 1024, 0, st>>>(g_dev, partial, wa, grads);
  return ss_check_launch("conv_wgrad_halo_reduce");
}


// ---------------------------------------------------------------------------------------------
// tcgen05.mma issue-rate probe (tools/umma_probe.py): a chain of n_mma bf16 MMAs (M=128, N, K=16) on operands already
// resident in shared memory (contents irrelevant), cycling over n_acc TMEM accumulators; reports cycles per MMA.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(int N, int n_mma, int n_acc, int commit_every,
                                                             long long* __restrict__ out, int a_off, int a_sbo,
                                                             int mn_major, int distinct) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t bar2;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  for (int i = threadIdx.x; i < (65536 + 256 * 128) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(smem_dyn + (base - smem_u32(smem_dyn)))[i] = 0x3C003C00u;
  fence_proxy_async();
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1);
    mbar_init(smem_u32(&bar2), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = uniform32(tmem_base_smem);
  if (warp == 0) {
    // mn_major = 1: both operands MN-major as in the weight-gradient kernels (A = two 64-wide atoms 16 KB apart, K steps of
    // 2 KB); distinct = 1: every group of 4 MMAs reads a different 8 KB of A (128 B further on), like the halo taps
    const uint32_t idesc = make_idesc(128, N, mn_major, mn_major);
    const uint64_t ad0 = mn_major ? make_sdesc(base + (uint32_t)a_off, 16384, (uint32_t)a_sbo)
                                  : make_sdesc(base + (uint32_t)a_off, 16, (uint32_t)a_sbo);
    const uint64_t bd0 = mn_major ? make_sdesc(base + 65536, 16384, 1024) : make_sdesc(base + 65536, 16, 1024);
    const uint64_t kadv = mn_major ? 128u : 2u;
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    if (elect_one()) {
      t0 = clock64();
      const uint32_t amask = (uint32_t)(n_acc - 1);          // n_acc is a power of two
      for (int gi = 0; gi < n_mma / 4; ++gi) {
        const uint32_t d = tm + ((uint32_t)gi & amask) * (uint32_t)N;
        const uint32_t accum = gi >= n_acc ? 1u : 0u;
        const uint64_t ad = ad0 + (distinct ? (uint64_t)(((uint32_t)gi & 63u) * 8u) : 0u);
        umma_bf16(d, ad, bd0, idesc, accum);
        umma_bf16(d, ad + kadv, bd0 + kadv, idesc, 1u);
        umma_bf16(d, ad + 2 * kadv, bd0 + 2 * kadv, idesc, 1u);
        umma_bf16(d, ad + 3 * kadv, bd0 + 3 * kadv, idesc, 1u);
        if (commit_every) umma_commit(smem_u32(&bar2));      // like a stage release: nobody waits on it
      }
      umma_commit(smem_u32(&bar));
      mbar_wait(smem_u32(&bar), phase);
      t1 = clock64();
      out[blockIdx.x] = t1 - t0;
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tm, 512);
}

extern "C" SSHSLIE_API int sshslie_umma_probe(int N, int n_mma, int n_acc, int commit_every, long long* out_cycles,
                                              int n_ctas, void* stream) {
  // A-operand placement experiments: SSHSLIE_PROBE_AOFF (byte offset of the window, multiple of 128) and
  // SSHSLIE_PROBE_SBO (byte stride between 8-row groups; 16 groups must fit in 64 KB)
  const char* ao = getenv("SSHSLIE_PROBE_AOFF");
  const char* sb = getenv("SSHSLIE_PROBE_SBO");
  const int a_off = ao ? atoi(ao) : 0, a_sbo = sb ? atoi(sb) : 1024;
  const int mn_major = ss_env_int("SSHSLIE_PROBE_MN", 0), distinct = ss_env_int("SSHSLIE_PROBE_DISTINCT", 0);
  if (N < 16 || N > 256 || (N % 16) || n_acc < 1 || n_acc * N > 512 || !out_cycles) {
    ss_set_error("sshslie_umma_probe: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
  umma_probe_kernel<<<n_ctas, 128, 65536 + 256 * 128 + 1024, (cudaStream_t)stream>>>(N, n_mma, n_acc, commit_every,
                                                                                    out_cycles, a_off, a_sbo, mn_major,
                                                                                    distinct);
  return ss_check_launch("umma_probe");
}

