// Launch-plan structures shared by the host planner (engine.cu) and the conv kernels.
//
// Every convolution-like layer of the network (Conv2d stride 1/2, ConvTranspose2d stride 2, and their
// data- and weight-gradients; model.py:17-47, 125-141) is lowered to ONE description, a ConvGeom:
//
//   out[b, oh, ow, n] = sum over K-slabs s, j<64 of  A_s[b, oh+dh_s, ow+dw_s, c0_s + j] * Wp[n][s*64 + j]
//
// where A_s is a (possibly parity-strided) view of a bf16 NHWC activation tensor, zero outside its bounds,
// and Wp is the bf16 weight matrix packed slab-major ([slab][n][64]: one slab is one contiguous TMA box).  Strided convolutions read 4 parity views of their
// input; transposed convolutions are 4 geoms, one per output parity class.  The same ConvGeom drives
//   * the gather GEMM (forward / dgrad)                     — conv_umma.cu (tcgen05) and conv_simt.cu
//   * the weight-gradient GEMM  dW[n][s*64+j] = sum_pixels G[b,oh,ow,n] * A_s[...]   (same slabs)
//   * the fp32 -> bf16 weight packer and the wgrad scatter back into the flat fp32 gradient buffer.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

#define SS_MAX_SRC 8            // views of one geom: a stride-2 conv over a bf16 pair reads 2 x 4 parity views
#define SS_MAX_WIN 6           // distinct (source, 64-channel slab) halo windows of one tile (the fusion conv reads 5)
#define SS_MAX_SLABS 81       // 9x9 taps x one 64-channel slab
#define SS_SLAB 64            // K-slab width in channels (= 128 bytes of bf16 = one SWIZZLE_128B row)

struct SrcView {
  const bf16* base;           // element (b=0, h=0, w=0, c=0) of the view
  int64_t sB, sH, sW;         // element strides
  int H, W;                   // view extent; reads outside are zero
};

struct Slab {
  int8_t src;                 // index into ConvGeom::src
  int8_t dh, dw;              // pixel offset in view coordinates
  int8_t no_wgrad;            // 1: residual (lo) slab of a hi+lo pair — shares its weights with the hi slab, so the
                              //    weight gradient skips it (its term is 2^-9 of the hi term and would alias dW)
  int16_t c0;                 // channel offset inside the source tensor
  int16_t wcn;                // how many of the 64 channels carry real weights (rest are zero)
  int32_t woff;               // weight element offset for (tap, first inner channel): kh*sH + kw*sW + wc0*sC
};

struct ConvGeom {
  int B, OH, OW;              // GEMM M-space: one row per (b, oh, ow)
  int N, Npad;                // valid / padded (multiple of 16) output columns
  int nsrc, nslabs;
  SrcView src[SS_MAX_SRC];
  Slab slab[SS_MAX_SLABS];
  // fp32 master weight addressing inside the flat parameter / gradient buffer
  int64_t w_off;              // first element of the weight tensor
  int32_t w_sN, w_sC;         // element strides of the column (n) index and of the inner-channel index
  // packed bf16 weights, slab-major: Wp[nslabs][Npad][64]
  bf16* wp;
  // tcgen05 tiling: the 128 GEMM rows of a tile are a (th x tw) block of the OHxOW grid, tw*th == 128
  int tw, th;
  int dup_c;                  // >= 0: input channel dup_c (source channel index) is the bf16 residual of channel dup_c - 1 and
                              // shares its weight (the packer copies it; the weight gradient ignores the lane); -1: none
  int halo_ok;                // 1: stride-1 full views, the halo-reuse kernels (gather and weight gradient) may take this geom;
                              // 2: stride-2 parity views / transposed parity classes, the halo GATHER kernels may
};

// Epilogue applied to 16 consecutive columns of one GEMM row (one output pixel).
enum { EPI_BF16 = 0, EPI_HEAD = 1, EPI_PLANE32 = 2 };
struct Epi {
  const float* bias;          // [N] fp32 or null
  const bf16* add;            // optional tensor added before activation / mask
  int64_t aB, aH, aW;
  const bf16* mask;           // optional: zero the result where mask <= 0 (ReLU backward)
  int64_t mB, mH, mW;
  int relu;
  int mode;
  bf16* out;                  // EPI_BF16: NHWC (strided) bf16 output, columns [0, n_store)
  bf16* out_lo;               // optional: bf16(v - float(bf16(v))) at the same strides -> hi+lo carries ~16 mantissa bits
  int64_t oB, oH, oW;
  int n_store;
  // optional column split (EPI_BF16): columns [n_split, n_split + n_store2) go to a SECOND tensor with its own mask — the
  // data gradient of a layer whose input is a concat writes both halves of the concat from one GEMM
  int n_split;                // 0: no split
  int n_store2;
  bf16* out2;
  int64_t o2B, o2H, o2W;
  const bf16* mask2;
  int64_t m2B, m2H, m2W;
  // EPI_HEAD: sigmoid; columns [0,C) -> R32 (B,C,H,W) fp32 and RI (NHWC bf16, stride ri_c); column C -> I32, RI[C]
  float* R32;
  float* I32;
  bf16* RI;
  int ri_c;
  int ri_lo_off;              // > 0: also store the bf16 residual of R at RI[pix*ri_c + ri_lo_off + n]
  int C, H, W;
  // EPI_PLANE32: column 0 -> plane32[(b*H + oh)*W + ow]
  float* plane32;
};
