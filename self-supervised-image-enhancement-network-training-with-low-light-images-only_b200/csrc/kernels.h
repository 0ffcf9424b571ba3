// Host-callable launchers of every kernel (internal header; the public ABI is include/sshslie_b200.h).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include "plan.h"
#include "../../include/sshslie_b200.h"

// conv_simt.cu

// "already done on this device" flag for the cudaFuncSetAttribute calls of the launchers: the attribute is per device, so a
// process that drives several GPUs sets it on each; two threads that race both set it (idempotent)
struct DeviceOnce {
  std::atomic<unsigned long long> mask{0};
  static int dev() { int d = 0; cudaGetDevice(&d); return d & 63; }
  bool done() const { return (mask.load(std::memory_order_acquire) >> dev()) & 1ull; }
  void set() { mask.fetch_or(1ull << dev(), std::memory_order_release); }
};

int ss_launch_conv_gather_simt(const ConvGeom* g_dev, const ConvGeom& g_host, const Epi& epi, cudaStream_t st);
int ss_launch_conv_wgrad_simt(const ConvGeom* g_dev, const ConvGeom& g_host, const bf16* G, int64_t gB, int64_t gH,
                              int64_t gW, int gN, float* grads, cudaStream_t st);
// db[n] += sum over pixels of G[pixel, n]; scratch: SS_BIAS_GRAD_MAX_BLOCKS x N floats (per-block partials, fixed-order sum)
int ss_launch_bias_grad(const bf16* G, int64_t npix, int ld, int N, float* db, float* scratch, cudaStream_t st);
int ss_launch_pack_weights(const ConvGeom* geoms_dev, const int* block_start_dev, int njobs, int total_blocks,
                           const float* params, cudaStream_t st, int first_block = 0, int n_blocks = -1);

// conv_umma.cu (tcgen05 / TMEM / TMA)
struct UmmaMaps;   // host-built CUtensorMaps of one geom (opaque here)
int ss_umma_supported(const ConvGeom& g);
int ss_launch_conv_gather_umma4(const ConvGeom* g_dev, const ConvGeom* g4_host, const UmmaMaps* maps4_host,
                                const Epi& epi_class0, cudaStream_t st);
int ss_umma_wgrad_supported(const ConvGeom& g);
int ss_umma_halo_supported(const ConvGeom& g);
int ss_launch_conv_gather_halo(const ConvGeom* g_dev, const ConvGeom& g_host, const UmmaMaps& maps, const Epi& epi,
                               cudaStream_t st);
int ss_umma_build_maps(const ConvGeom& g, UmmaMaps* maps);     // needs final device pointers
int ss_launch_conv_gather_umma(const ConvGeom* g_dev, const ConvGeom& g_host, const UmmaMaps& maps, const Epi& epi,
                               cudaStream_t st);
int ss_umma_build_gmap(const bf16* G, int64_t gB, int64_t gH, int64_t gW, int ld_extent, const ConvGeom& g,
                       void* out_map /* 128 bytes, 64-byte aligned */);
int ss_launch_conv_wgrad_umma(const ConvGeom* g_dev, const ConvGeom& g_host, const UmmaMaps& maps, const void* gmap,
                              int gN, long long bias_off, float* partial, float* grads, cudaStream_t st);
size_t ss_umma_wgrad_partial_floats(const ConvGeom& g, int gN);
size_t ss_umma_maps_size();
void ss_set_wgrad_part(int part);   // profiling: 0 both kernels, 1 GEMM only, 2 split-K reduce only
// halo-reuse weight gradient (stride-1 layers): G tiles are 16x8 like the halo tiles
int ss_umma_wgrad_halo_supported(const ConvGeom& g, int gN);
size_t ss_umma_wgrad_halo_partial_floats(const ConvGeom& g, int gN);
int ss_umma_build_gmap_halo(const bf16* G, int64_t gB, int64_t gH, int64_t gW, int ld_extent, const ConvGeom& g,
                            void* out_map);
int ss_launch_conv_wgrad_halo(const ConvGeom* g_dev, const ConvGeom& g_host, const UmmaMaps& maps, const void* gmap,
                              int gN, long long bias_off, float* partial, float* grads, cudaStream_t st);

// conv_pipe.cu: persistent pipelined gather (stride-1 layers), see the file header
size_t ss_pipe_plan_size();
int ss_umma_pipe_supported(const ConvGeom& g, const Epi& epi);
int ss_launch_conv_gather_pipe(const ConvGeom& g, const UmmaMaps& maps, const Epi& epi, void* plan_cache,
                               int* cache_valid, cudaStream_t st);

// elementwise.cu
int ss_launch_nchw32_to_nhwc16(const float* x, bf16* out, int B, int C, int H, int W, int ldo, cudaStream_t st);
int ss_launch_nhwc16_to_nchw32(const bf16* in, float* y, int B, int C, int H, int W, int ldi, cudaStream_t st);
// nearest resize (h, w) -> (ho, wo) of r (+ a), ATen index semantics
int ss_launch_upsample_add(const bf16* r, const bf16* a, bf16* out, int B, int h, int w, int ho, int wo, cudaStream_t st);
int ss_launch_upsample_add_pair(const bf16* r, const bf16* a, const bf16* a_lo, bf16* out, bf16* out_lo, int B, int h, int w,
                                int ho, int wo, cudaStream_t st);
int ss_launch_fuse_concat(const bf16* r1, const bf16* a2, const bf16* r2, const bf16* a1, const bf16* a1l, const bf16* r3,
                          const bf16* r3l, const bf16* a0, const bf16* a0l, bf16* fg, int B, int H, int W, int h2, int w2,
                          int h1, int w1, cudaStream_t st);
int ss_launch_pack_ri(const float* R, const float* I, bf16* RI, int B, int C, int H, int W, cudaStream_t st);
int ss_launch_make_s(const float* R, const float* I, const float* Id, float* S32, bf16* Sb, int B, int C, int H, int W,
                     cudaStream_t st);
int ss_launch_s_bwd(const float* dS32, const float* dSf32, const bf16* dSb, const float* R, const float* I,
                    const float* Id, float* dR32, float* dI32, float* dId32, int B, int C, int H, int W, cudaStream_t st);
int ss_fourier_loss(const float* x, const float* S, const float* mask, float* dS, float* partial_out, int n_img, int H,
                    int W, float grad_scale, int accumulate, float* work, cudaStream_t stream);
int64_t ss_fourier_work_floats(int n_img, int H, int W);   // workspace of the DFT path (0 for power-of-two planes <= 128)
int ss_launch_head_bwd(const float* dR32, const float* R32, const bf16* dRI, int ld_dri, const float* dI32,
                       const float* I32, bf16* dc8, int ld_out, int B, int C, int H, int W, cudaStream_t st);
int ss_launch_concat_bwd(const bf16* dfg, const bf16* r3, bf16* dr3, bf16* p2, bf16* p1, int B, int H, int W,
                         cudaStream_t st);
int ss_launch_pool2(const bf16* du, const bf16* addp, const bf16* maskr, bf16* out_sum, bf16* out_masked, float* out32,
                    int B, int h, int w, cudaStream_t st);
int ss_launch_finalize_losses(const float* pix_partials, int pix_rows, const float* four_partials, int four_rows,
                              const sshslie_loss_cfg* cfg, float* losses, int B, int C, int H, int W, cudaStream_t st);

// fixed-order reduction of per-block partial results into (segments of) the flat gradient buffer: column j of the
// partial rows belongs to the segment that contains it; dst[j] += sum over rows (row order fixed -> deterministic)
struct RedSegs { float* dst[10]; int len[10]; int n; };
int ss_launch_reduce_rows(const float* partials, int nrows, int ncols, const RedSegs& segs, cudaStream_t st);
#define SS_BIAS_GRAD_MAX_BLOCKS 512
#define SS_ATTN_WGRAD_MAX_BLOCKS 16
#define SS_ATTN_WGRAD_COLS (5 * (64 * 64 + 64))

// attention.cu   (tokens: T = B*L rows of 64 fp32)
struct AttnBuffers {
  float *x, *q, *k, *v, *o, *lse, *h, *t32;                  // forward (x = a3 as fp32)
  float *dq, *dk, *dv, *d_o, *dh, *dx, *Dv;                  // backward scratch
  bf16 *qp, *kvp;                                            // per-head bf16 hi+lo operands of the tensor-core core (L >= SS_ATTN_TC_MIN_L)
};
#define SS_ATTN_TC_MIN_L 1024
int ss_env_int(const char* name, int dflt);
int ss_attn_tc_min_l();      // SS_ATTN_TC_MIN_L unless SSHSLIE_ATTN_TC_MIN_L overrides it (tuning)
// attention_tc.cu: tcgen05 attention core (q, k, v fp32 [B*L][64] -> o fp32 [B*L][64], lse [B][4][L])
int ss_attention_core_tc(const float* q, const float* k, const float* v, bf16* qp, bf16* kvp, float* o, float* lse, int B,
                         int L, cudaStream_t st);
int ss_attention_forward(const bf16* a3, bf16* t_out, const float* params, const int64_t* poff, AttnBuffers bufs,
                         int B, int L, cudaStream_t st);
int ss_attention_backward(const float* dt, const bf16* a3, bf16* da3, const float* params, float* grads,
                          const int64_t* poff, AttnBuffers bufs, int B, int L, cudaStream_t st);

// scratch: SS_ATTN_WGRAD_MAX_BLOCKS x SS_ATTN_WGRAD_COLS floats
int ss_attention_backward_weights(const float* dt, float* grads, const int64_t* poff, AttnBuffers bufs, int B, int L,
                                  float* scratch, cudaStream_t st);

// loss.cu / fft_loss.cu / adam.cu : see include/sshslie_b200.h (exported directly)
int ss_pixel_losses(const float* x, const float* R, const float* I, const float* Id, const float* Re,
                    const sshslie_loss_cfg& cfg, int B, int C, int H, int W, float* partials, float* dR, float* dI,
                    float* dId, float* dS, float* dRe, cudaStream_t st);
int64_t ss_pixel_losses_scratch_floats(int B, int C, int H, int W);   // partial rows + edge-weight maps
int ss_pixel_losses_blocks(int B, int C, int H, int W);       // rows of 9 partial sums ss_pixel_losses writes
int ss_reduce_partials(const float* partials, int nrows, int ncols, float* out, int accumulate, cudaStream_t st);
