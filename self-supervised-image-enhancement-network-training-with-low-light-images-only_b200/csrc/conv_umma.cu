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

#include "common.cuh"
#include "epilogue.cuh"
#include "kernels.h"

#define UM_STAGES 4
#define UM_A_BYTES (128 * 128)            // 128 rows x 64 bf16
#define UM_THREADS 192
#define UM_WAIT_CYCLES (2000000000LL)     // ~1 s: a stuck barrier traps instead of hanging the GPU

struct alignas(64) UmmaMaps {
  CUtensorMap src[SS_MAX_SRC];
  CUtensorMap w;
};
size_t ss_umma_maps_size() { return sizeof(UmmaMaps); }

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
SS_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SS_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
SS_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
SS_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SS_DEVINL void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SS_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
SS_DEVINL bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
SS_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > UM_WAIT_CYCLES) {
      printf("sshslie: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x,
             (int)threadIdx.x, bar, parity);
      __trap();
    }
  }
}
SS_DEVINL void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SS_DEVINL void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
SS_DEVINL void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}
SS_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SS_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SS_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SS_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SS_DEVINL void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
SS_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
SS_DEVINL void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// shared-memory matrix descriptor, K-major or MN-major operand in SWIZZLE_128B atoms (8 rows x 128 B = 1024 B)
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout (2 = SWIZZLE_128B)
SS_DEVINL uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor, kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1), majors (bit 15/16),
// N>>3 at bit 17, M>>4 at bit 24
SS_DEVINL uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// gather GEMM kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(UM_THREADS, 1)
conv_gather_umma_kernel(const ConvGeom* __restrict__ gp, const __grid_constant__ UmmaMaps maps, Epi epi,
                        int tmem_cols) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ ConvGeom g;
  __shared__ __align__(8) uint64_t full_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t empty_bar[UM_STAGES];
  __shared__ __align__(8) uint64_t accum_bar;
  __shared__ uint32_t tmem_base_smem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  {
    const int* src = reinterpret_cast<const int*>(gp);
    int* dst = reinterpret_cast<int*>(&g);
    for (int i = threadIdx.x; i < (int)(sizeof(ConvGeom) / 4); i += blockDim.x) dst[i] = src[i];
  }
  // operand ring, 1024-byte aligned (SWIZZLE_128B atom)
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  __syncthreads();
  const int Npad = g.Npad;
  const uint32_t b_bytes = (uint32_t)Npad * 128u;
  const uint32_t stage_bytes = UM_A_BYTES + b_bytes;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < g.nsrc; ++s) tma_prefetch_desc(&maps.src[s]);
    tma_prefetch_desc(&maps.w);
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
  int t = blockIdx.x;
  const int twi = t % tiles_w; t /= tiles_w;
  const int thi = t % tiles_h;
  const int b = t / tiles_h;
  const int oh0 = thi * g.th, ow0 = twi * g.tw;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int s = 0; s < g.nslabs; ++s) {
        const int st = s % UM_STAGES;
        const uint32_t ph = (uint32_t)(s / UM_STAGES) & 1u;
        mbar_wait(smem_u32(&empty_bar[st]), ph ^ 1u);
        const Slab sl = g.slab[s];
        const uint32_t a_dst = dyn_base + (uint32_t)st * stage_bytes;
        const uint32_t fb = smem_u32(&full_bar[st]);
        mbar_expect_tx(fb, stage_bytes);
        tma_load_4d(a_dst, &maps.src[sl.src], fb, sl.c0, ow0 + sl.dw, oh0 + sl.dh, b);
        tma_load_2d(a_dst + UM_A_BYTES, &maps.w, fb, s * SS_SLAB, 0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = make_idesc(128, Npad, 0, 0);
      for (int s = 0; s < g.nslabs; ++s) {
        const int st = s % UM_STAGES;
        const uint32_t ph = (uint32_t)(s / UM_STAGES) & 1u;
        mbar_wait(smem_u32(&full_bar[st]), ph);
        tc_fence_after();
        const uint32_t a_addr = dyn_base + (uint32_t)st * stage_bytes;
        const uint32_t b_addr = a_addr + UM_A_BYTES;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_sdesc(a_addr + k * 32, 16, 1024);
          const uint64_t bd = make_sdesc(b_addr + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (s > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(smem_u32(&empty_bar[st]));      // frees the stage when these MMAs retire
      }
      umma_commit(smem_u32(&accum_bar));            // accumulator complete
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter = warp % 4 =====
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int oh = oh0 + row / g.tw, ow = ow0 + row % g.tw;
    const bool ok = oh < g.OH && ow < g.OW;
    mbar_wait(smem_u32(&accum_bar), 0);
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

int ss_umma_supported(const ConvGeom& g) {
  if (g.Npad < 16 || g.Npad > 256 || (g.Npad % 16)) return 0;
  if (g.tw < 8 || g.tw * g.th != 128) return 0;
  if (g.OW % g.tw) return 0;                       // partial tiles along W are not handled (rows along H are)
  for (int s = 0; s < g.nsrc; ++s) {
    const SrcView& v = g.src[s];
    if (((uintptr_t)v.base & 15) || (v.sW % 8) || (v.sH % 8) || (v.sB % 8)) return 0;
  }
  return 1;
}

static int encode_src(const SrcView& v, int ld_extent, int tw, int th, int B, CUtensorMap* out) {
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
    const int rc = encode_src(g.src[s], ext, g.tw, g.th, g.B, &maps->src[s]);
    if (rc) return rc;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) return SSHSLIE_ERR_CUDA;
  const cuuint64_t Ktot = (cuuint64_t)g.nslabs * SS_SLAB;
  cuuint64_t dims[2] = {Ktot, (cuuint64_t)g.Npad};
  cuuint64_t strides[1] = {Ktot * 2};
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
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_gather_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      ss_set_error("conv_gather_umma: cannot raise dynamic shared memory: %s",
                   cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_set = true;
  }
  conv_gather_umma_kernel<<<tiles, UM_THREADS, smem, st>>>(g_dev, maps, epi, cols);
  return ss_check_launch("conv_gather_umma");
}

int ss_umma_wgrad_supported(const ConvGeom&) { return 0; }

int ss_launch_conv_wgrad_umma(const ConvGeom*, const ConvGeom&, const UmmaMaps&, const bf16*, int64_t, int64_t,
                              int64_t, int, float*, cudaStream_t) {
  ss_set_error("conv_wgrad_umma: not built in this revision");
  return SSHSLIE_ERR_ARG;
}
