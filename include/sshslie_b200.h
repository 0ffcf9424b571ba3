/*
 * sshslie_b200 — C-ABI of the B200-native SS-HSLIE hot path.
 *
 * The reference (medemirhan/Self-supervised-Image-Enhancement-Network-Training-With-Low-Light-Images-Only)
 * has no FFI layer: its hot path is reached through `LowLightEnhance.forward` / `.compute_loss` /
 * `loss.backward()` / `optimizer.step()` (model.py:229-234, 544-575, 313-316).  These entry points are
 * what a binding for exactly those calls needs; the Python drop-in (`sshslie_b200/model.py`) binds them
 * with ctypes, and INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer except `sshslie_engine*`, `const char*` and host tables is a DEVICE pointer;
 *   - the caller owns every buffer; functions only ENQUEUE work on `stream` (a cudaStream_t passed as
 *     void*), never allocate, free or synchronise, and are CUDA-graph capturable
 *     (exception: sshslie_engine_bind, which uploads constant tables and synchronises `stream`);
 *   - return value: 0 = ok, negative = error; `sshslie_last_error()` describes the last failure of the
 *     calling thread; no C++ exception crosses the boundary;
 *   - fp32 tensors are NCHW contiguous (the reference's layout); parameters and gradients are ONE flat
 *     fp32 buffer in the reference's state_dict order (`sshslie_param_table`).
 */
#ifndef SSHSLIE_B200_H
#define SSHSLIE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(SSHSLIE_BUILD)
#define SSHSLIE_API __attribute__((visibility("default")))
#else
#define SSHSLIE_API
#endif

#define SSHSLIE_OK 0
#define SSHSLIE_ERR_ARG (-1)      /* bad shape / null pointer / unsupported configuration */
#define SSHSLIE_ERR_CUDA (-2)     /* a CUDA runtime or driver call failed (launch error, wrong arch) */
#define SSHSLIE_ERR_WORKSPACE (-3)/* workspace too small or engine not bound to this workspace */

#define SSHSLIE_NUM_PARAM_TENSORS 46
#define SSHSLIE_NUM_LOSSES 7      /* total, rec, R_fid, I_smooth_low, I_smooth_delta, fourier, spectral (model.py:566-574) */

/* engine flags */
#define SSHSLIE_FLAG_TRAIN 1      /* reserve workspace for compute_loss + backward */
#define SSHSLIE_FLAG_FORCE_SIMT 2 /* run every conv GEMM on the CUDA-core cross-check kernels (tests only) */

typedef struct sshslie_engine sshslie_engine;

/* loss weights: LowLightEnhance.__init__ (model.py:178-194), values from config/*.yml:24-31 */
typedef struct sshslie_loss_cfg {
  float c_loss_reconstruction;
  float c_loss_r_fidelity;
  float c_loss_i_smooth_low;
  float c_loss_i_smooth_delta;
  float c_loss_fourier;
  float c_loss_spectral_cons;
  float alpha_i_smooth_low;
  float alpha_i_smooth_delta;
} sshslie_loss_cfg;

SSHSLIE_API int sshslie_version(void);
SSHSLIE_API const char* sshslie_last_error(void);

/* Flat parameter layout.  Entry i of the reference's state_dict (46 tensors for any band count) starts
 * at offsets[i] and has sizes[i] floats.  Returns the total float count (1,141,922 for channels=64). */
SSHSLIE_API int64_t sshslie_param_table(int channels, int64_t* offsets, int64_t* sizes);

/* Engine = host-side launch plan for one (batch, channels, height, width).  channels == 64; height and width even and
 * >= 16 (the reference itself needs them even: stride-2 conv + ConvTranspose2d(output_padding=1), model.py:37-43; the
 * illumination pyramid is ceil(n/2) per level with ATen's nearest indices, model.py:127-129, 156-169).  With
 * SSHSLIE_FLAG_TRAIN height and width must be multiples of 8 (any: patch_size is a free config value). */
SSHSLIE_API int sshslie_engine_create(sshslie_engine** out, int batch, int channels, int height, int width, int flags);
SSHSLIE_API void sshslie_engine_destroy(sshslie_engine* e);
SSHSLIE_API int64_t sshslie_engine_workspace_bytes(const sshslie_engine* e);
/* Bind the engine to a caller-owned workspace (>= workspace_bytes, 1024-byte aligned): uploads the launch
 * plan, the Fourier mask (model.py:460-464) and FFT twiddles, zeroes padding lanes.  Synchronises. */
SSHSLIE_API int sshslie_engine_bind(sshslie_engine* e, void* workspace, int64_t workspace_bytes, void* stream);

/* LowLightEnhance.forward (model.py:229-234): x (B,C,H,W) -> R (B,C,H,W), I (B,1,H,W), I_delta (B,1,H,W),
 * S (B,C,H,W).  Output pointers may alias nothing in the workspace. */
SSHSLIE_API int sshslie_forward(sshslie_engine* e, const float* x, const float* params,
                    float* R, float* I, float* I_delta, float* S, void* stream);

/* IllumAdjustmentNet.forward(I, R) on its own (model.py:143-175): I (B,1,H,W), R (B,C,H,W) fp32 in, I_delta (B,1,H,W) fp32
 * out; the engine may be an inference or a training engine of that shape. */
SSHSLIE_API int sshslie_illum_forward(sshslie_engine* e, const float* I, const float* R, const float* params, float* I_delta,
                          void* stream);

/* LowLightEnhance.compute_loss + loss.backward() (model.py:544-575, 315): writes the seven loss values
 * (order of SSHSLIE_NUM_LOSSES) to losses[7], d(total_loss)/d(params) to grads (flat, overwritten),
 * and the four forward outputs when the pointers are non-null.  phase_mask selects sub-ranges of the
 * step so that a data-parallel caller can start the gradient all-reduce of finished buckets early:
 *   bit0: forward + loss + backward through the second DecompositionNet pass and IllumAdjustmentNet
 *         (illum_adjust_net gradients are final afterwards)
 *   bit1: backward through the first DecompositionNet pass (decomposition_net gradients final afterwards)
 */
SSHSLIE_API int sshslie_loss_and_grad(sshslie_engine* e, const float* x, const float* params,
                          const sshslie_loss_cfg* cfg, float* grads, float* losses,
                          float* R, float* I, float* I_delta, float* S, int phase_mask, void* stream);

/* torch.optim.Adam defaults as used by the reference (model.py:213, 316): one fused launch over the flat
 * buffers.  grad_scale multiplies the gradient first (1/world_size after an all-reduce(sum)). */
SSHSLIE_API int sshslie_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n,
                      float lr, float beta1, float beta2, float eps, int step, float grad_scale, void* stream);

/* Profiling aid (synchronises, allocates events): runs ONE step eagerly with a cudaEvent pair around every recorded
 * launch group and returns the number of rows (>0) or a negative status; rows are then read with
 * sshslie_profile_row: name ("phase/kind:layer[impl]"), device milliseconds, algorithmic FLOPs and bytes. */
SSHSLIE_API int sshslie_profile_step(sshslie_engine* e, const float* x, const float* params,
                         const sshslie_loss_cfg* cfg, float* grads, float* losses, void* stream);
/* number of kernels this library has enqueued since load (for launch accounting) */
SSHSLIE_API long long sshslie_launch_count(void);
SSHSLIE_API int sshslie_profile_row(int i, char* name, int name_cap, float* ms, double* flops, double* bytes);

/* On-device patch pipeline replacing the numpy crop + augmentation + H2D of train_model (model.py:301-312, utils.py:7-34).
 * cubes_dev: device array of B pointers to resident HWC fp32 cubes (one per sample); meta_dev: B x {h, w, x0, y0, mode}
 * int32 (x0, y0, mode drawn by the host in the reference's order); out: (B, C, ps, ps) fp32.  Bit-exact copy. */
SSHSLIE_API int sshslie_gather_patches(const float* const* cubes_dev, const int* meta_dev, float* out, int B, int C,
                           int patch_size, void* stream);

/* Result writer replacing the host-side permute + de-normalisation of test_model / evaluate_model (model.py:421-424,
 * 381-387): src (C,H,W) fp32 on the device -> dst (H,W,C) fp32; apply != 0: dst = src * scale + offset as two separately
 * rounded fp32 operations (numpy's S * (max - min) + min: bit-exact). */
SSHSLIE_API int sshslie_denorm_hwc(const float* src_chw, float* dst_hwc, int C, int H, int W, float scale, float offset,
                       int apply, void* stream);

/* PSNR / SAM sums of two (H,W,C) fp32 cubes as metrics.py:13-14,31-34 evaluates them through torchmetrics 1.6.2:
 * sums2[0] = sum of squared differences, sums2[1] = sum over pixels of the spectral angle (radians); device doubles. */
SSHSLIE_API int sshslie_psnr_sam(const float* pred_hwc, const float* target_hwc, int H, int W, int C, double* sums2,
                     void* stream);

/* SSIM of two (H,W,C) fp32 cubes as metrics.py:16-19 evaluates it through torchmetrics 1.6.2 (the cube unsqueezed to
 * (1,H,W,C): H is the channel axis, the 11x11 gaussian window, sigma 1.5, slides over the (W, C) plane; c1 = (0.01 range)^2,
 * c2 = (0.03 range)^2).  sum1[0] = sum of the index over the H x (W-10) x (C-10) averaged outputs; device double. */
SSHSLIE_API int sshslie_ssim_sum(const float* pred_hwc, const float* target_hwc, int H, int W, int C, float c1, float c2,
                     double* sum1, void* stream);

/* ---- single kernels, exported for kernel-level parity tests and profiling ---- */

/* Scratch of the two loss entry points below for a (B,C,H,W) problem: they keep NO floating-point atomics - every block
 * writes its partial sums to the scratch and a second launch adds them in a fixed order, so results are bit-repeatable
 * (the reference trains with cudnn.deterministic = True, main.py:165). */
SSHSLIE_API int64_t sshslie_loss_scratch_bytes(int B, int C, int H, int W);

/* fourier_spectrum_loss (model.py:456-473) forward + d/dS.  x,S,dS: (n_img,H,W) fp32 planes, H and W
 * in [2,1024] (powers of two up to 128: FFT with the plane in shared memory; anything else: DFT line passes over a
 * complex plane in the scratch); mask (H,W) fp32; adds sum_k mask*| |X|-|S| | to *sum_out (not zeroed);
 * dS += grad_scale * d(sum)/dS when dS != NULL.  scratch: >= sshslie_loss_scratch_bytes(1, n_img, H, W). */
SSHSLIE_API int sshslie_fourier_loss(const float* x, const float* S, const float* mask, float* dS, float* sum_out,
                         int n_img, int H, int W, float grad_scale, void* scratch, int64_t scratch_bytes, void* stream);

/* The five pixel-space loss terms (model.py:450-454, 475-481, 491-542, 551) and their gradients.
 * sums[9] (device) receives the raw term sums; gradients are written (not accumulated) already scaled by
 * c_loss_x / count.  Any gradient pointer may be NULL (forward only).  scratch: sshslie_loss_scratch_bytes(B,C,H,W). */
SSHSLIE_API int sshslie_pixel_losses(const float* x, const float* R, const float* I, const float* Idelta, const float* S,
                         const float* R_enh, const sshslie_loss_cfg* cfg, int B, int C, int H, int W,
                         float* sums, float* dR, float* dI, float* dIdelta, float* dS, float* dR_enh,
                         void* scratch, int64_t scratch_bytes, void* stream);

/* One conv layer through the implicit-GEMM executors, for kernel parity tests: x (B,Cin,H,W) fp32,
 * w (Cout,Cin,k,k) [or (Cin,Cout,k,k) when transposed], y (B,Cout,OH,OW) fp32.  Internally converts to the
 * engine's bf16 NHWC layout, runs the per-tap tcgen05 (impl=1), halo-reuse tcgen05 (impl=2, stride-1 layers), persistent
 * pipelined tcgen05 (impl=3, stride-1 layers, forward / dgrad) or CUDA-core (impl=0) kernel, converts back.
 * kind: 0 = forward, 1 = dgrad (x is dY, y is dX), 2 = wgrad (x is the layer input, w receives dW, y is dY; a non-NULL
 * `bias` then receives db = sum of dY over pixels, as the fused bias row of the tcgen05 weight gradient produces it).
 * scratch must hold sshslie_conv2d_scratch_bytes(...) bytes. */
SSHSLIE_API int64_t sshslie_conv2d_scratch_bytes(int B, int Cin, int Cout, int H, int W, int k, int stride);
SSHSLIE_API int sshslie_conv2d(int kind, int impl, int transposed, float* x, float* w, const float* bias, float* y,
                   int B, int Cin, int Cout, int H, int W, int k, int stride, int relu,
                   void* scratch, int64_t scratch_bytes, void* stream);

/* TransformerBlock alone (model.py:99-119), for kernel-level parity tests of the attention kernels (fp32 CUDA-core kernels
 * up to 1023 tokens, tcgen05 core from 1024 tokens).  x, y, dy, dx: fp32 (B, 64, H, W) as the reference block sees them
 * (tokens = the H*W grid; the kernels read / write them as bf16 like the engine does); params / dparams: the block's ten
 * tensors flat in state_dict order (q.w q.b k.w k.b v.w v.b ff1.w ff1.b ff2.w ff2.b = 20800 floats).  with_backward != 0:
 * also dparams = d/dparams of <y, dy> and dx = (d/dx of <y, dy>) masked by x > 0: in the network the block's input is the
 * output of a ReLU layer (model.py:128, 150-153) and the engine folds that layer's backward mask into this kernel.
 * Synchronises.  scratch: 1024-byte aligned. */
SSHSLIE_API int64_t sshslie_transformer_block_scratch_bytes(int B, int H, int W);
SSHSLIE_API int sshslie_transformer_block(int with_backward, const float* x, const float* params, float* y, const float* dy,
                              float* dx, float* dparams, int B, int H, int W, void* scratch, int64_t scratch_bytes,
                              void* stream);

/* with SSHSLIE_CONV2D_TIMING=n in the environment, sshslie_conv2d repeats the layer's launches n times between two
 * CUDA events; this returns the mean device time (ms) of the last such call (tools/conv_bench.py) */
SSHSLIE_API float sshslie_conv2d_last_ms(void);

/* tcgen05.mma issue-rate microbenchmark (tools/umma_probe.py); out_cycles[n_ctas] receives the elapsed SM cycles */
SSHSLIE_API int sshslie_umma_probe(int N, int n_mma, int n_acc, int commit_every, long long* out_cycles,
                       int n_ctas, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSHSLIE_B200_H */
