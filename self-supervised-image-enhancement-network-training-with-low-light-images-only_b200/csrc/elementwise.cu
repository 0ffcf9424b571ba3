// Layout, resampling and head/glue kernels around the conv GEMMs (all HBM-bound, fp32 math).
//   fp32 tensors: NCHW planes (the reference's layout, consumed by the loss/FFT kernels and returned to the caller)
//   bf16 tensors: NHWC (the GEMM operand layout), channel stride `ld`
#include <algorithm>
#include "common.cuh"
#include "kernels.h"

#define EW_CHECK(name) return ss_check_launch(name)

#define C64_PITCH 33
SS_DEVINL void tile_to_vec8(const float (*tile)[C64_PITCH], int i, int q, float* f) {
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] = tile[8 * q + k][i];
}
SS_DEVINL uint4 pack8(const float* f) {
  uint4 v;
  v.x = pack2(f[0], f[1]); v.y = pack2(f[2], f[3]); v.z = pack2(f[4], f[5]); v.w = pack2(f[6], f[7]);
  return v;
}


// A 32-pixel x 32-channel transposing tile.  blockDim = (32, 8).  HW is a multiple of 32 (H, W multiples of 8),
// so the 32 consecutive pixels of a tile never straddle two images.

// ---------------------------------------------------------------------------------------------
// x (B,C,H,W) fp32  ->  out (B,H,W,ldo) bf16        [decomposition_net input, model.py:51-52]
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// 64-channel fast paths (every tensor of this network): a block moves 32 pixels x ALL 64 channels through one padded tile.
//   plane side (NCHW fp32): thread (x = pixel, y) touches channels y, y + 8, ... : eight independent 128-byte row segments per
//                           warp and array, all loads issued before the first use
//   pixel side (NHWC bf16): thread t owns pixel t / 8 and the 8 channels [8 (t % 8), +8): ONE 16-byte access, 128 contiguous
//                           bytes per pixel; tile[c][i] with pitch 33 is conflict-free for both access patterns
// (the generic kernels below them walk the channels in chunks of 32 with two barriers per chunk and 2-byte bf16 accesses:
//  55-60 % of the time of each went into that.)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw32_to_nhwc16_c64_kernel(const float* __restrict__ x, bf16* __restrict__ out,
                                                                   int HW, int ldo) {
  SS_PDL_ENTRY();
  __shared__ float tile[64][C64_PITCH];
  const int b = blockIdx.y, hw0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool in = hw0 + tx < HW;
  const float* xp = x + (int64_t)b * 64 * HW + hw0 + tx;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = in ? __ldg(xp + (int64_t)(ty + 8 * j) * HW) : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tile[ty + 8 * j][tx] = v[j];
  __syncthreads();
  const int t = ty * 32 + tx, i = t >> 3, q = t & 7;
  if (hw0 + i < HW) {
    float f[8];
    tile_to_vec8(tile, i, q, f);
    *reinterpret_cast<uint4*>(out + ((int64_t)b * HW + hw0 + i) * ldo + 8 * q) = pack8(f);
  }
}

// blocks tile ONE image (grid.z = b): 32 pixels x 32 channels through a padded shared tile; any H*W (tail pixels guarded)
__global__ void nchw32_to_nhwc16_kernel(const float* __restrict__ x, bf16* __restrict__ out, int C, int HW, int ldo) {
  SS_PDL_ENTRY();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int hw0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, hw = hw0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && hw < HW) ? x[((int64_t)b * C + c) * HW + hw] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {            // i = pixel, threadIdx.x = channel
    const int c = c0 + threadIdx.x;
    if (c < C && hw0 + i < HW) out[((int64_t)b * HW + hw0 + i) * ldo + c] = f2bf(tile[threadIdx.x][i]);
  }
}
int ss_launch_nchw32_to_nhwc16(const float* x, bf16* out, int B, int C, int H, int W, int ldo, cudaStream_t st) {
  if (C == 64 && (ldo % 8) == 0 && ((uintptr_t)out & 15) == 0) {
    ss_launch_pdl(nchw32_to_nhwc16_c64_kernel, dim3((unsigned)((H * W + 31) / 32), B), dim3(32, 8), (size_t)0, st, x, out, H * W, ldo);
    EW_CHECK("nchw32_to_nhwc16_c64");
  }
  dim3 grid((unsigned)((H * W + 31) / 32), (C + 31) / 32, B);
  ss_launch_pdl(nchw32_to_nhwc16_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), st, x, out, C, H * W, ldo);
  EW_CHECK("nchw32_to_nhwc16");
}

__global__ void nhwc16_to_nchw32_kernel(const bf16* __restrict__ in, float* __restrict__ y, int C, int HW, int ldi) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int hw0 = blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {            // i = pixel, x = channel
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && hw0 + i < HW) ? bf2f(in[((int64_t)b * HW + hw0 + i) * ldi + c]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {            // i = channel, x = pixel
    const int c = c0 + i, hw = hw0 + threadIdx.x;
    if (c < C && hw < HW) y[((int64_t)b * C + c) * HW + hw] = tile[threadIdx.x][i];
  }
}
int ss_launch_nhwc16_to_nchw32(const bf16* in, float* y, int B, int C, int H, int W, int ldi, cudaStream_t st) {
  dim3 grid((unsigned)((H * W + 31) / 32), (C + 31) / 32, B);
  nhwc16_to_nchw32_kernel<<<grid, dim3(32, 8), 0, st>>>(in, y, C, H * W, ldi);
  EW_CHECK("nhwc16_to_nchw32");
}

// ---------------------------------------------------------------------------------------------
// out[b, y, x, :] = r[b, y/2, x/2, :] (+ a[b, y/2, x/2, :])      nearest x2 of (relu-out + skip), 64 channels
// [F.interpolate(mode='nearest') of deconv_k + conv_k, model.py:156-165]
// ---------------------------------------------------------------------------------------------
SS_DEVINL uint4 add8(const uint4& p, const uint4& q) {
  float f[8], g[8];
  unpack8(p, f);
  unpack8(q, g);
  uint4 v;
  v.x = pack2(f[0] + g[0], f[1] + g[1]);
  v.y = pack2(f[2] + g[2], f[3] + g[3]);
  v.z = pack2(f[4] + g[4], f[5] + g[5]);
  v.w = pack2(f[6] + g[6], f[7] + g[7]);
  return v;
}
__global__ void upsample2_add_kernel(const bf16* __restrict__ r, const bf16* __restrict__ a, bf16* __restrict__ out,
                                     int h, int w, int64_t total /* B*h*w*8 vectors of the SOURCE */) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int xs = (int)(pix % w);
  const int64_t t = pix / w;
  const int ys = (int)(t % h);
  const int64_t b = t / h;
  uint4 v = reinterpret_cast<const uint4*>(r)[i];
  if (a) {
    float f[8], g[8];
    unpack8(v, f);
    unpack8(reinterpret_cast<const uint4*>(a)[i], g);
    v.x = pack2(f[0] + g[0], f[1] + g[1]);
    v.y = pack2(f[2] + g[2], f[3] + g[3]);
    v.z = pack2(f[4] + g[4], f[5] + g[5]);
    v.w = pack2(f[6] + g[6], f[7] + g[7]);
  }
  const int W2 = 2 * w;
  uint4* o = reinterpret_cast<uint4*>(out) + ((b * 2 * h + 2 * ys) * W2 + 2 * xs) * 8 + q;
  o[0] = v;
  o[8] = v;
  o[(int64_t)W2 * 8] = v;
  o[(int64_t)W2 * 8 + 8] = v;
}
// F.interpolate(size=(ho, wo), mode='nearest') for any sizes: src = min(floor(dst * (float)in / out), in - 1), the index
// ATen computes (UpSampleKernel.cpp nearest_idx; equals dst >> 1 for out == 2 in).  One thread per DESTINATION vector.
SS_DEVINL int nearest_src(int d, float scale, int in) { return min((int)floorf((float)d * scale), in - 1); }
__global__ void upsample_nearest_add_kernel(const uint4* __restrict__ r, const uint4* __restrict__ a, uint4* __restrict__ out,
                                            int h, int w, int ho, int wo, float sh, float sw, int64_t total) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int x = (int)(pix % wo);
  const int64_t t = pix / wo;
  const int y = (int)(t % ho);
  const int64_t b = t / ho;
  const int64_t s = ((b * h + nearest_src(y, sh, h)) * w + nearest_src(x, sw, w)) * 8 + q;
  out[i] = a ? add8(r[s], a[s]) : r[s];
}
// the same resize of (r + a + a_lo), stored as a bf16 pair: hi = bf16(v), lo = bf16(v - hi)   [up(deconv2 + conv1), model.py:161-164]
__global__ void upsample_nearest_add_pair_kernel(const uint4* __restrict__ r, const uint4* __restrict__ a,
                                                 const uint4* __restrict__ al, uint4* __restrict__ out,
                                                 uint4* __restrict__ out_lo, int h, int w, int ho, int wo, float sh,
                                                 float sw, int64_t total) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int x = (int)(pix % wo);
  const int64_t t = pix / wo;
  const int y = (int)(t % ho);
  const int64_t b = t / ho;
  const int64_t s = ((b * h + nearest_src(y, sh, h)) * w + nearest_src(x, sw, w)) * 8 + q;
  float f[8], g[8], l[8], lo[8];
  unpack8(__ldg(r + s), f);
  unpack8(__ldg(a + s), g);
  unpack8(__ldg(al + s), l);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    f[j] = f[j] + g[j] + l[j];
    lo[j] = f[j] - bf2f(f2bf(f[j]));
  }
  out[i] = pack8(f);
  out_lo[i] = pack8(lo);
}
int ss_launch_upsample_add_pair(const bf16* r, const bf16* a, const bf16* a_lo, bf16* out, bf16* out_lo, int B, int h, int w,
                                int ho, int wo, cudaStream_t st) {
  const int64_t total = (int64_t)B * ho * wo * 8;
  ss_launch_pdl(upsample_nearest_add_pair_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st,
                (const uint4*)r, (const uint4*)a, (const uint4*)a_lo, (uint4*)out, (uint4*)out_lo, h, w, ho, wo,
                (float)h / (float)ho, (float)w / (float)wo, total);
  EW_CHECK("upsample_nearest_add_pair");
}

// (h, w) -> (ho, wo); the exact-doubling case keeps the one-read-four-writes kernel
int ss_launch_upsample_add(const bf16* r, const bf16* a, bf16* out, int B, int h, int w, int ho, int wo, cudaStream_t st) {
  if (ho == 2 * h && wo == 2 * w) {
    const int64_t total = (int64_t)B * h * w * 8;
    ss_launch_pdl(upsample2_add_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, r, a, out, h, w, total);
    EW_CHECK("upsample2_add");
  }
  const int64_t total = (int64_t)B * ho * wo * 8;
  ss_launch_pdl(upsample_nearest_add_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st,
                (const uint4*)r, (const uint4*)a, (uint4*)out, h, w, ho, wo, (float)h / (float)ho, (float)w / (float)wo, total);
  EW_CHECK("upsample_nearest_add");
}

// ---------------------------------------------------------------------------------------------
// fg[b,y,x,:] = [ (r1+a2)[y/4,x/4] | hi(r2+a1)[y/2,x/2] | hi(r3+a0)[y,x] | lo(r3+a0)[y,x] | lo(r2+a1)[y/2,x/2] ]   320 channels  [model.py:168-172]
// ---------------------------------------------------------------------------------------------
// one thread = one (pixel, 8-channel vector) of ALL four 64-channel groups: eight independent 16-byte loads in flight, four
// 16-byte stores; the eight threads of a pixel write 128 contiguous bytes per group (no divergence inside a warp)
__global__ void __launch_bounds__(256) fuse_concat_kernel(const uint4* __restrict__ r1, const uint4* __restrict__ a2,
                                                          const uint4* __restrict__ r2, const uint4* __restrict__ a1,
                                                          const uint4* __restrict__ a1l,
                                                          const uint4* __restrict__ r3, const uint4* __restrict__ r3l,
                                                          const uint4* __restrict__ a0, const uint4* __restrict__ a0l,
                                                          uint4* __restrict__ fg, int H, int W, int h2, int w2, int h1, int w1,
                                                          float sh2, float sw2, float sh1, float sw1,
                                                          int64_t total /* B*H*W*8 */) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int x = (int)(pix % W);
  const int64_t t = pix / W;
  const int y = (int)(t % H);
  const int64_t b = t / H;
  const int64_t s1 = ((b * h2 + nearest_src(y, sh2, h2)) * w2 + nearest_src(x, sw2, w2)) * 8 + q;
  const int64_t s2 = ((b * h1 + nearest_src(y, sh1, h1)) * w1 + nearest_src(x, sw1, w1)) * 8 + q;
  const uint4 v_r1 = __ldg(r1 + s1), v_a2 = __ldg(a2 + s1), v_r2 = __ldg(r2 + s2), v_a1 = __ldg(a1 + s2);
  const uint4 v_a1l = __ldg(a1l + s2);
  const uint4 v_r3 = __ldg(r3 + i), v_a0 = __ldg(a0 + i);
  uint4 v_r3l = make_uint4(0u, 0u, 0u, 0u), v_a0l = v_r3l;
  if (r3l) { v_r3l = __ldg(r3l + i); v_a0l = __ldg(a0l + i); }
  uint4* o = fg + pix * 40 + q;             // 320 channels: [d1 | d2 hi | d3 hi | d3 lo | d2 lo]
  o[0] = add8(v_r1, v_a2);
  {  // d2 = deconv2 + conv1 (with conv1's residual) as a bf16 pair
    float f2[8], g2[8], l2[8], lo2[8];
    unpack8(v_r2, f2);
    unpack8(v_a1, g2);
    unpack8(v_a1l, l2);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f2[j] = f2[j] + g2[j] + l2[j];
      lo2[j] = f2[j] - bf2f(f2bf(f2[j]));
    }
    o[8] = pack8(f2);
    o[32] = pack8(lo2);
  }
  // full-resolution block d3 = deconv3 + conv0 in ~16-bit mantissa: hi -> channels [128,192), residual -> [192,256)
  float f[8], g[8], hsum[8], r[8];
  unpack8(v_r3, f);
  unpack8(v_a0, g);
#pragma unroll
  for (int j = 0; j < 8; ++j) hsum[j] = f[j] + g[j];
  if (r3l) {
    unpack8(v_r3l, f);
    unpack8(v_a0l, g);
#pragma unroll
    for (int j = 0; j < 8; ++j) hsum[j] += f[j] + g[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = hsum[j] - bf2f(f2bf(hsum[j]));
  o[16] = pack8(hsum);
  o[24] = pack8(r);
}
int ss_launch_fuse_concat(const bf16* r1, const bf16* a2, const bf16* r2, const bf16* a1, const bf16* a1l, const bf16* r3,
                          const bf16* r3l, const bf16* a0, const bf16* a0l, bf16* fg, int B, int H, int W, int h2, int w2,
                          int h1, int w1, cudaStream_t st) {
  const int64_t total = (int64_t)B * H * W * 8;
  ss_launch_pdl(fuse_concat_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, 
      (const uint4*)r1, (const uint4*)a2, (const uint4*)r2, (const uint4*)a1, (const uint4*)a1l, (const uint4*)r3, (const uint4*)r3l,
      (const uint4*)a0, (const uint4*)a0l, (uint4*)fg, H, W, h2, w2, h1, w1, (float)h2 / (float)H, (float)w2 / (float)W,
      (float)h1 / (float)H, (float)w1 / (float)W, total);
  EW_CHECK("fuse_concat");
}

// ---------------------------------------------------------------------------------------------
// cat[R, I] (model.py:146) in the layout the illumination net reads: RI (B,H,W,192) bf16 =
//   [ bf16(R) (64) | bf16(I), bf16(I - bf16(I)), 62 zero lanes (never written) | bf16(R - bf16(R)) (64) ]   (hi + lo pairs)
// normally written by the sigmoid head's epilogue; this kernel builds it from caller-provided fp32 planes
// (IllumAdjustmentNet.forward on its own).  64 bands.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_ri_kernel(const float* __restrict__ R, const float* __restrict__ I,
                                                      bf16* __restrict__ RI, int HW) {
  __shared__ float tile[64][C64_PITCH];
  const int b = blockIdx.y, hw0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool in = hw0 + tx < HW;
  const float* rp = R + (int64_t)b * 64 * HW + hw0 + tx;
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = in ? __ldg(rp + (int64_t)(ty + 8 * j) * HW) : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) tile[ty + 8 * j][tx] = v[j];
  __syncthreads();
  const int t = ty * 32 + tx, i = t >> 3, q = t & 7;
  if (hw0 + i < HW) {
    float f[8], lo[8];
    tile_to_vec8(tile, i, q, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) lo[k] = f[k] - bf2f(f2bf(f[k]));
    bf16* o = RI + ((int64_t)b * HW + hw0 + i) * 192;
    *reinterpret_cast<uint4*>(o + 8 * q) = pack8(f);
    *reinterpret_cast<uint4*>(o + 128 + 8 * q) = pack8(lo);
    if (q == 0) {
      const float iv = __ldg(I + (int64_t)b * HW + hw0 + i);
      o[64] = f2bf(iv);
      o[65] = f2bf(iv - bf2f(f2bf(iv)));
    }
  }
}
int ss_launch_pack_ri(const float* R, const float* I, bf16* RI, int B, int C, int H, int W, cudaStream_t st) {
  if (C != 64) { ss_set_error("pack_ri: 64 bands only"); return SSHSLIE_ERR_ARG; }
  pack_ri_kernel<<<dim3((unsigned)((H * W + 31) / 32), B), dim3(32, 8), 0, st>>>(R, I, RI, H * W);
  EW_CHECK("pack_ri");
}

// ---------------------------------------------------------------------------------------------
// S = R*I_delta + R*I_low  (model.py:233) -> S32 (B,C,H,W) fp32 and Sb (B,H,W,C) bf16 (2nd decomposition input)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) make_s_c64_kernel(const float* __restrict__ R, const float* __restrict__ I,
                                                         const float* __restrict__ Id, float* __restrict__ S32,
                                                         bf16* __restrict__ Sb, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tile[64][C64_PITCH];
  const int b = blockIdx.y, hw0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const bool in = hw0 + tx < HW;
  const int64_t p = (int64_t)b * HW + hw0 + tx;
  const int64_t a0 = (int64_t)b * 64 * HW + hw0 + tx;
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = in ? __ldg(R + a0 + (int64_t)(ty + 8 * j) * HW) : 0.f;
  const float id = in ? __ldg(Id + p) : 0.f, il = in ? __ldg(I + p) : 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float sv = r[j] * id + r[j] * il;
    if (in) S32[a0 + (int64_t)(ty + 8 * j) * HW] = sv;
    tile[ty + 8 * j][tx] = sv;
  }
  if (!Sb) return;
  __syncthreads();
  const int t = ty * 32 + tx, i = t >> 3, q = t & 7;
  if (hw0 + i < HW) {
    float f[8];
    tile_to_vec8(tile, i, q, f);
    *reinterpret_cast<uint4*>(Sb + ((int64_t)b * HW + hw0 + i) * 64 + 8 * q) = pack8(f);
  }
}
__global__ void make_s_kernel(const float* __restrict__ R, const float* __restrict__ I, const float* __restrict__ Id,
                              float* __restrict__ S32, bf16* __restrict__ Sb, int C, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int hw0 = blockIdx.x * 32;
  const int64_t p0 = (int64_t)b * HW + hw0;
  const int c0 = blockIdx.y * 32;
  const bool in = hw0 + (int)threadIdx.x < HW;
  const float id = in ? Id[p0 + threadIdx.x] : 0.f, il = in ? I[p0 + threadIdx.x] : 0.f;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    if (c < C && in) {
      const int64_t a = ((int64_t)b * C + c) * HW + hw0 + threadIdx.x;
      const float r = R[a];
      const float s = r * id + r * il;
      S32[a] = s;
      tile[i][threadIdx.x] = s;
    }
  }
  __syncthreads();
  if (Sb) {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int c = c0 + threadIdx.x;
      if (c < C && hw0 + i < HW) Sb[(p0 + i) * C + c] = f2bf(tile[threadIdx.x][i]);
    }
  }
}
int ss_launch_make_s(const float* R, const float* I, const float* Id, float* S32, bf16* Sb, int B, int C, int H, int W,
                     cudaStream_t st) {
  if (C == 64 && ((uintptr_t)Sb & 15) == 0) {
    ss_launch_pdl(make_s_c64_kernel, dim3((unsigned)((H * W + 31) / 32), B), dim3(32, 8), (size_t)0, st, R, I, Id, S32, Sb, H * W);
    EW_CHECK("make_s_c64");
  }
  dim3 grid((unsigned)((H * W + 31) / 32), (C + 31) / 32, B);
  ss_launch_pdl(make_s_kernel, dim3(grid), dim3(dim3(32, 8)), (size_t)(0), st, R, I, Id, S32, Sb, C, H * W);
  EW_CHECK("make_s");
}

// ---------------------------------------------------------------------------------------------
// backward of S = R*(Id + I):  dS = dS32 (loss terms) + dSb (second decomposition pass, bf16 NHWC)
//   dR32 += dS*(Id+I) ;  t = sum_c dS*R ;  dId32 += t ;  dI32 += t        (block = 32 pixels x all channels)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) s_bwd_c64_kernel(const float* __restrict__ dS32, const float* __restrict__ dSf32,
                                                        const bf16* __restrict__ dSb, const float* __restrict__ R,
                                                        const float* __restrict__ I, const float* __restrict__ Id,
                                                        float* __restrict__ dR32, float* __restrict__ dI32,
                                                        float* __restrict__ dId32, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tile[64][C64_PITCH];
  __shared__ float part[8][32];
  const int b = blockIdx.y, hw0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t_ = ty * 32 + tx, i = t_ >> 3, q = t_ & 7;
  const int64_t p = (int64_t)b * HW + hw0 + tx;
  const int64_t a0 = (int64_t)b * 64 * HW + hw0 + tx;
  // every load of the thread first: the bf16 gradient of the second decomposition pass (pixel side), then the planes
  const uint4 dv = *reinterpret_cast<const uint4*>(dSb + ((int64_t)b * HW + hw0 + i) * 64 + 8 * q);
  float ds[8], rr[8], dr[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t a = a0 + (int64_t)(ty + 8 * j) * HW;
    ds[j] = __ldg(dS32 + a) + (dSf32 ? __ldg(dSf32 + a) : 0.f);
    rr[j] = __ldg(R + a);
    dr[j] = dR32[a];
  }
  const float gain = __ldg(Id + p) + __ldg(I + p);
  {
    float f[8];
    unpack8(dv, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[8 * q + k][i] = f[k];
  }
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float d = ds[j] + tile[ty + 8 * j][tx];
    dR32[a0 + (int64_t)(ty + 8 * j) * HW] = dr[j] + d * gain;
    t = fmaf(d, rr[j], t);
  }
  part[ty][tx] = t;
  __syncthreads();
  if (ty == 0) {
    float sm = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sm += part[k][tx];
    dId32[p] += sm;
    dI32[p] += sm;
  }
}
__global__ void s_bwd_kernel(const float* __restrict__ dS32, const float* __restrict__ dSf32,
                             const bf16* __restrict__ dSb, const float* __restrict__ R,
                             const float* __restrict__ I, const float* __restrict__ Id, float* __restrict__ dR32,
                             float* __restrict__ dI32, float* __restrict__ dId32, int C, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tile[32][33];
  __shared__ float part[8][32];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int b = (int)(p0 / HW);
  const int hw0 = (int)(p0 - (int64_t)b * HW);
  const float gain = Id[p0 + threadIdx.x] + I[p0 + threadIdx.x];
  float t = 0.f;
  for (int c0 = 0; c0 < C; c0 += 32) {
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {          // i = pixel, x = channel
      const int c = c0 + threadIdx.x;
      tile[threadIdx.x][i] = (c < C) ? bf2f(dSb[(p0 + i) * C + c]) : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {          // i = channel, x = pixel
      const int c = c0 + i;
      if (c < C) {
        const int64_t a = ((int64_t)b * C + c) * HW + hw0 + threadIdx.x;
        const float ds = dS32[a] + (dSf32 ? dSf32[a] : 0.f) + tile[i][threadIdx.x];
        dR32[a] += ds * gain;
        t = fmaf(ds, R[a], t);
      }
    }
  }
  part[threadIdx.y][threadIdx.x] = t;
  __syncthreads();
  if (threadIdx.y == 0) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][threadIdx.x];
    dId32[p0 + threadIdx.x] += s;
    dI32[p0 + threadIdx.x] += s;
  }
}
int ss_launch_s_bwd(const float* dS32, const float* dSf32, const bf16* dSb, const float* R, const float* I,
                    const float* Id, float* dR32, float* dI32, float* dId32, int B, int C, int H, int W, cudaStream_t st) {
  if (C == 64 && (H * W) % 32 == 0 && ((uintptr_t)dSb & 15) == 0) {
    ss_launch_pdl(s_bwd_c64_kernel, dim3((unsigned)(H * W / 32), B), dim3(32, 8), (size_t)0, st, dS32, dSf32, dSb, R, I, Id, dR32,
                  dI32, dId32, H * W);
    EW_CHECK("s_bwd_c64");
  }
  ss_launch_pdl(s_bwd_kernel, dim3((unsigned)((int64_t)B * H * W / 32)), dim3(dim3(32, 8)), (size_t)(0), st, dS32, dSf32, dSb, R, I, Id, dR32, dI32,
                                                                           dId32, C, H * W);
  EW_CHECK("s_bwd");
}

// ---------------------------------------------------------------------------------------------
// backward of the sigmoid heads (model.py:68-69):
//   dc8[pix, c] = (dR32[b,c,pix] + dRI[pix, c]) * R(1-R)      c < C
//   dc8[pix, C] = (dI32[pix] + dRI[pix, C]) * I(1-I)          (only when dI32 != NULL; else no column C)
//   columns above are left untouched (zeroed once at bind time)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) head_bwd_c64_kernel(const float* __restrict__ dR32, const float* __restrict__ R32,
                                                           const bf16* __restrict__ dRI, int ld_dri,
                                                           const float* __restrict__ dI32, const float* __restrict__ I32,
                                                           bf16* __restrict__ dc8, int ld_out, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tg[64][C64_PITCH];   // raw gradient
  __shared__ float tr[64][C64_PITCH];   // sigmoid'(.) = R(1-R)
  const int b = blockIdx.y, hw0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int t = ty * 32 + tx, i = t >> 3, q = t & 7;
  const int64_t a0 = (int64_t)b * 64 * HW + hw0 + tx;
  const int64_t pi = (int64_t)b * HW + hw0 + i;
  uint4 dv = make_uint4(0u, 0u, 0u, 0u);
  if (dRI) dv = *reinterpret_cast<const uint4*>(dRI + pi * ld_dri + 8 * q);
  float g[8], r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t a = a0 + (int64_t)(ty + 8 * j) * HW;
    g[j] = __ldg(dR32 + a);
    r[j] = __ldg(R32 + a);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    tg[ty + 8 * j][tx] = g[j];
    tr[ty + 8 * j][tx] = r[j] * (1.f - r[j]);
  }
  __syncthreads();
  {
    float f[8], d[8], o[8];
    unpack8(dv, f);
    tile_to_vec8(tg, i, q, o);
    tile_to_vec8(tr, i, q, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (o[k] + f[k]) * d[k];
    *reinterpret_cast<uint4*>(dc8 + pi * ld_out + 8 * q) = pack8(o);
  }
  if (dI32 && ty == 0) {
    const int64_t p = (int64_t)b * HW + hw0 + tx;
    float gi = dI32[p];
    if (dRI) gi += bf2f(dRI[p * ld_dri + 64]);
    const float il = I32[p];
    dc8[p * ld_out + 64] = f2bf(gi * il * (1.f - il));
  }
}
__global__ void head_bwd_kernel(const float* __restrict__ dR32, const float* __restrict__ R32,
                                const bf16* __restrict__ dRI, int ld_dri, const float* __restrict__ dI32,
                                const float* __restrict__ I32, bf16* __restrict__ dc8, int ld_out, int C, int HW) {
  SS_PDL_ENTRY();
  __shared__ float tg[32][33];   // raw gradient
  __shared__ float tr[32][33];   // sigmoid'(.) = R(1-R)
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int b = (int)(p0 / HW);
  const int hw0 = (int)(p0 - (int64_t)b * HW);
  for (int c0 = 0; c0 < C; c0 += 32) {
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {          // i = channel, x = pixel (coalesced NCHW reads)
      const int c = c0 + i;
      float g = 0.f, d = 0.f;
      if (c < C) {
        const int64_t a = ((int64_t)b * C + c) * HW + hw0 + threadIdx.x;
        const float r = R32[a];
        g = dR32[a];
        d = r * (1.f - r);
      }
      tg[i][threadIdx.x] = g;
      tr[i][threadIdx.x] = d;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {          // i = pixel, x = channel (coalesced NHWC writes)
      const int c = c0 + threadIdx.x;
      if (c < C) {
        float g = tg[threadIdx.x][i];
        if (dRI) g += bf2f(dRI[(p0 + i) * ld_dri + c]);
        dc8[(p0 + i) * ld_out + c] = f2bf(g * tr[threadIdx.x][i]);
      }
    }
  }
  if (dI32 && threadIdx.y == 0) {
    const int64_t p = p0 + threadIdx.x;
    float g = dI32[p];
    if (dRI) g += bf2f(dRI[p * ld_dri + C]);
    const float il = I32[p];
    dc8[p * ld_out + C] = f2bf(g * il * (1.f - il));
  }
}
int ss_launch_head_bwd(const float* dR32, const float* R32, const bf16* dRI, int ld_dri, const float* dI32,
                       const float* I32, bf16* dc8, int ld_out, int B, int C, int H, int W, cudaStream_t st) {
  if (C == 64 && (H * W) % 32 == 0 && (ld_out % 8) == 0 && ((uintptr_t)dc8 & 15) == 0 &&
      (!dRI || ((ld_dri % 8) == 0 && ((uintptr_t)dRI & 15) == 0))) {
    ss_launch_pdl(head_bwd_c64_kernel, dim3((unsigned)(H * W / 32), B), dim3(32, 8), (size_t)0, st, dR32, R32, dRI, ld_dri, dI32,
                  I32, dc8, ld_out, H * W);
    EW_CHECK("head_bwd_c64");
  }
  ss_launch_pdl(head_bwd_kernel, dim3((unsigned)((int64_t)B * H * W / 32)), dim3(dim3(32, 8)), (size_t)(0), st, dR32, R32, dRI, ld_dri, dI32, I32, dc8,
                                                                              ld_out, C, H * W);
  EW_CHECK("head_bwd");
}

// ---------------------------------------------------------------------------------------------
// backward of the concat/upsample that feeds feature_fusion (model.py:168-172), given dfg (B,H,W,192):
//   dr3 = dfg[...,128:192] * (r3 > 0)                          (B,H,W,64)
//   p2  = 2x2 sum-pool of dfg[..., 64:128]                     (B,H/2,W/2,64)
//   p1  = 4x4 sum-pool of dfg[...,  0: 64]                     (B,H/4,W/4,64)
// one thread = one (H/4 x W/4 pixel, 8-channel vector)
// ---------------------------------------------------------------------------------------------
__global__ void concat_bwd_kernel(const uint4* __restrict__ dfg, const uint4* __restrict__ r3, uint4* __restrict__ dr3,
                                  uint4* __restrict__ p2, uint4* __restrict__ p1, int H, int W, int64_t total) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int w4 = W / 4, h4 = H / 4;
  const int x4 = (int)(pix % w4);
  const int64_t t = pix / w4;
  const int y4 = (int)(t % h4);
  const int64_t b = t / h4;
  float s1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = 0.f;
  for (int yy = 0; yy < 2; ++yy)
    for (int xx = 0; xx < 2; ++xx) {
      float s2[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s2[j] = 0.f;
      for (int y1 = 0; y1 < 2; ++y1)
        for (int x1 = 0; x1 < 2; ++x1) {
          const int y = y4 * 4 + yy * 2 + y1, x = x4 * 4 + xx * 2 + x1;
          const int64_t p = (b * H + y) * W + x;
          float f[8], m[8];
          unpack8(dfg[p * 24 + q], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) s1[j] += f[j];
          unpack8(dfg[p * 24 + 8 + q], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) s2[j] += f[j];
          unpack8(dfg[p * 24 + 16 + q], f);
          unpack8(r3[p * 8 + q], m);
          uint4 o;
          o.x = pack2(m[0] > 0.f ? f[0] : 0.f, m[1] > 0.f ? f[1] : 0.f);
          o.y = pack2(m[2] > 0.f ? f[2] : 0.f, m[3] > 0.f ? f[3] : 0.f);
          o.z = pack2(m[4] > 0.f ? f[4] : 0.f, m[5] > 0.f ? f[5] : 0.f);
          o.w = pack2(m[6] > 0.f ? f[6] : 0.f, m[7] > 0.f ? f[7] : 0.f);
          dr3[p * 8 + q] = o;
        }
      uint4 o;
      o.x = pack2(s2[0], s2[1]); o.y = pack2(s2[2], s2[3]); o.z = pack2(s2[4], s2[5]); o.w = pack2(s2[6], s2[7]);
      p2[((b * (H / 2) + y4 * 2 + yy) * (W / 2) + x4 * 2 + xx) * 8 + q] = o;
    }
  uint4 o;
  o.x = pack2(s1[0], s1[1]); o.y = pack2(s1[2], s1[3]); o.z = pack2(s1[4], s1[5]); o.w = pack2(s1[6], s1[7]);
  p1[pix * 8 + q] = o;
}
int ss_launch_concat_bwd(const bf16* dfg, const bf16* r3, bf16* dr3, bf16* p2, bf16* p1, int B, int H, int W,
                         cudaStream_t st) {
  const int64_t total = (int64_t)B * (H / 4) * (W / 4) * 8;
  ss_launch_pdl(concat_bwd_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), (size_t)(0), st, (const uint4*)dfg, (const uint4*)r3, (uint4*)dr3,
                                                                    (uint4*)p2, (uint4*)p1, H, W, total);
  EW_CHECK("concat_bwd");
}

// ---------------------------------------------------------------------------------------------
// backward of nearest x2 upsampling of (relu-out + skip):  s = 2x2 sum-pool(du) (+ addp)
//   out_sum    = s                  (gradient of the skip tensor)               optional
//   out_masked = s * (maskr > 0)    (gradient of the ReLU output)               optional
//   out32      = s as fp32          (gradient of the transformer output)        optional
// du: (B,2h,2w,64) ; everything else (B,h,w,64)
// ---------------------------------------------------------------------------------------------
__global__ void pool2_kernel(const uint4* __restrict__ du, const uint4* __restrict__ addp,
                             const uint4* __restrict__ maskr, uint4* __restrict__ out_sum,
                             uint4* __restrict__ out_masked, float* __restrict__ out32, int h, int w, int64_t total) {
  SS_PDL_ENTRY();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int x = (int)(pix % w);
  const int64_t t = pix / w;
  const int y = (int)(t % h);
  const int64_t b = t / h;
  const int W2 = 2 * w;
  const uint4* s = du + ((b * 2 * h + 2 * y) * W2 + 2 * x) * 8 + q;
  float acc[8], f[8];
  unpack8(s[0], acc);
  unpack8(s[8], f);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += f[j];
  unpack8(s[(int64_t)W2 * 8], f);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += f[j];
  unpack8(s[(int64_t)W2 * 8 + 8], f);
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] += f[j];
  if (addp) {
    unpack8(addp[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += f[j];
  }
  if (out_sum) {
    uint4 o;
    o.x = pack2(acc[0], acc[1]); o.y = pack2(acc[2], acc[3]); o.z = pack2(acc[4], acc[5]); o.w = pack2(acc[6], acc[7]);
    out_sum[i] = o;
  }
  if (out_masked) {
    float m[8];
    unpack8(maskr[i], m);
    uint4 o;
    o.x = pack2(m[0] > 0.f ? acc[0] : 0.f, m[1] > 0.f ? acc[1] : 0.f);
    o.y = pack2(m[2] > 0.f ? acc[2] : 0.f, m[3] > 0.f ? acc[3] : 0.f);
    o.z = pack2(m[4] > 0.f ? acc[4] : 0.f, m[5] > 0.f ? acc[5] : 0.f);
    o.w = pack2(m[6] > 0.f ? acc[6] : 0.f, m[7] > 0.f ? acc[7] : 0.f);
    out_masked[i] = o;
  }
  if (out32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) out32[i * 8 + j] = acc[j];
  }
}
int ss_launch_pool2(const bf16* du, const bf16* addp, const bf16* maskr, bf16* out_sum, bf16* out_masked, float* out32,
                    int B, int h, int w, cudaStream_t st) {
  const int64_t total = (int64_t)B * h * w * 8;
  ss_launch_pdl(pool2_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), (size_t)(0), st, (const uint4*)du, (const uint4*)addp,
                                                               (const uint4*)maskr, (uint4*)out_sum,
                                                               (uint4*)out_masked, out32, h, w, total);
  EW_CHECK("pool2");
}

// ---------------------------------------------------------------------------------------------
// fixed-order reduction of per-block partials into the flat gradient buffer (deterministic replacement of fp32 atomics)
// block = (32 columns, 32 row parts): part q adds rows q, q + 32, ... (independent loads), the 32 parts of a column are then
// combined in a fixed order through shared memory.  (One thread per column walking all rows took 22 us on 512 rows.)
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const float* __restrict__ partials, int nrows, int ncols,
                                                           const __grid_constant__ RedSegs segs) {
  __shared__ float part[32][33];
  const int j = blockIdx.x * 32 + threadIdx.x, q = threadIdx.y;
  float t = 0.f;
  if (j < ncols) {
    int r = q;
    for (; r + 96 < nrows; r += 128) {
      const float a = partials[(size_t)r * ncols + j], b = partials[(size_t)(r + 32) * ncols + j];
      const float c = partials[(size_t)(r + 64) * ncols + j], d = partials[(size_t)(r + 96) * ncols + j];
      t += a; t += b; t += c; t += d;
    }
    for (; r < nrows; r += 32) t += partials[(size_t)r * ncols + j];
  }
  part[q][threadIdx.x] = t;
  __syncthreads();
  if (q != 0 || j >= ncols) return;
  t = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) t += part[i][threadIdx.x];
  int o = j;
  for (int s = 0; s < segs.n; ++s) {
    if (o < segs.len[s]) { segs.dst[s][o] += t; return; }
    o -= segs.len[s];
  }
}
int ss_launch_reduce_rows(const float* partials, int nrows, int ncols, const RedSegs& segs, cudaStream_t st) {
  reduce_rows_kernel<<<(ncols + 31) / 32, dim3(32, 32), 0, st>>>(partials, nrows, ncols, segs);
  EW_CHECK("reduce_rows");
}

// ---------------------------------------------------------------------------------------------
// raw term sums -> the seven loss values of model.py:557-574
//   sums: 0 rec | 1,2 I_smooth_low x,y | 3 |R-Re| | 4,5 grad fidelity x,y | 6,7 I_smooth_delta x,y | 8 spectral | 9 fourier
// ---------------------------------------------------------------------------------------------
// one warp per term sum: lanes stride the per-block partial sums of the loss kernels in a fixed order, fixed shuffle tree ->
// the seven loss values are bit-repeatable (the reference runs with cudnn.deterministic, main.py:165)
__global__ void __launch_bounds__(320) finalize_losses_kernel(const float* __restrict__ pix, int pix_rows,
                                                              const float* __restrict__ four, int four_rows,
                                                              sshslie_loss_cfg cfg, float* __restrict__ losses, int B, int C,
                                                              int H, int W) {
  SS_PDL_ENTRY();
  __shared__ float sums[10];
  const int term = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float t = 0.f;
  if (term < 9) {
    for (int r = lane; r < pix_rows; r += 32) t += pix[(size_t)r * 9 + term];
  } else {
    for (int r = lane; r < four_rows; r += 32) t += four[r];
  }
  t = warp_sum(t);
  if (lane == 0) sums[term] = t;
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double n0 = (double)B * C * H * W;
  const double nx1 = (double)B * H * (W - 1), ny1 = (double)B * (H - 1) * W;
  const double nxc = nx1 * C, nyc = ny1 * C;
  const double nsp = (double)B * (C - 1) * H * W;
  const double l_rec = sums[0] / n0;
  const double l_ilow = sums[1] / nx1 + sums[2] / ny1;
  const double l_rfid = sums[3] / n0 + 0.5 * (sums[4] / nxc + sums[5] / nyc);
  const double l_idel = sums[6] / nxc + sums[7] / nyc;
  const double l_spec = sums[8] / nsp;
  const double l_four = sums[9] / n0;
  const double total = cfg.c_loss_reconstruction * l_rec + cfg.c_loss_r_fidelity * l_rfid +
                       cfg.c_loss_i_smooth_low * l_ilow + cfg.c_loss_i_smooth_delta * l_idel +
                       cfg.c_loss_fourier * l_four + cfg.c_loss_spectral_cons * l_spec;
  losses[0] = (float)total;
  losses[1] = (float)l_rec;
  losses[2] = (float)l_rfid;
  losses[3] = (float)l_ilow;
  losses[4] = (float)l_idel;
  losses[5] = (float)l_four;
  losses[6] = (float)l_spec;
}
int ss_launch_finalize_losses(const float* pix_partials, int pix_rows, const float* four_partials, int four_rows,
                              const sshslie_loss_cfg* cfg, float* losses, int B, int C, int H, int W, cudaStream_t st) {
  ss_launch_pdl(finalize_losses_kernel, dim3(1), dim3(320), (size_t)(0), st, pix_partials, pix_rows, four_partials,
                four_rows, *cfg, losses, B, C, H, W);
  EW_CHECK("finalize_losses");
}


// ---------------------------------------------------------------------------------------------
// on-device patch pipeline: the reference crops + augments every training patch in numpy and ships
// 4 MiB per patch over PCIe (model.py:301-312).  Here the normalised cubes stay resident in HBM (HWC fp32, as load_hsi
// returns them) and one kernel does crop -> one of the 8 dihedral variants (utils.py:7-34: np.rot90(k) then optional
// np.flipud) -> HWC-to-NCHW for the whole batch.  meta[b] = {h, w, x0, y0, mode}: the host still draws x0, y0, mode with
// numpy in the reference's order, so runs are reproducible against it.  Pure copy: bit-exact.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gather_patches_kernel(const float* const* __restrict__ cubes,
                                                             const int* __restrict__ meta, float* __restrict__ out,
                                                             int C, int ps) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= ps) return;
  const int w = meta[b * 5 + 1], x0 = meta[b * 5 + 2], y0 = meta[b * 5 + 3], mode = meta[b * 5 + 4];
  const int k = mode >> 1, flip = mode & 1, n1 = ps - 1;
  const int ii = flip ? n1 - i : i;                 // undo flipud (applied last)
  int si, sj;                                       // source pixel inside the un-augmented patch
  if (k == 0) { si = ii; sj = j; }
  else if (k == 1) { si = j; sj = n1 - ii; }        // np.rot90(P)[i, j] = P[j, n-1-i]
  else if (k == 2) { si = n1 - ii; sj = n1 - j; }
  else { si = n1 - j; sj = ii; }
  const float* src = cubes[b] + ((int64_t)(x0 + si) * w + (y0 + sj)) * C;   // crop = cube[x0:x0+ps, y0:y0+ps, :]
  float* dst = out + (((int64_t)b * C) * ps + i) * ps + j;
  const int64_t plane = (int64_t)ps * ps;
  for (int c = 0; c < C; c += 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + c);
    dst[(c + 0) * plane] = v.x;
    dst[(c + 1) * plane] = v.y;
    dst[(c + 2) * plane] = v.z;
    dst[(c + 3) * plane] = v.w;
  }
}

extern "C" int sshslie_gather_patches(const float* const* cubes_dev, const int* meta_dev, float* out, int B, int C,
                                      int patch_size, void* stream) {
  if (!cubes_dev || !meta_dev || !out || B < 1 || C < 4 || (C % 4) || patch_size < 1) {
    ss_set_error("sshslie_gather_patches: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  dim3 grid((patch_size + 127) / 128, patch_size, B);
  gather_patches_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(cubes_dev, meta_dev, out, C, patch_size);
  return ss_check_launch("gather_patches");
}


// ---------------------------------------------------------------------------------------------
// result writer: the reference turns every output into an HWC numpy array on the host and, for S,
// de-normalises it there: S.squeeze(0).permute(1,2,0).cpu().numpy() * (max - min) + min (model.py:421-424).  Here one
// kernel does the NCHW -> HWC transposition and the de-normalisation on the device (two separately rounded fp32
// operations, as numpy does them: bit-exact), so the host only receives the finished cube.
// ---------------------------------------------------------------------------------------------
__global__ void denorm_hwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int64_t HW,
                                  float scale, float offset, int apply) {
  __shared__ float tile[32][33];
  const int64_t p0 = (int64_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {          // i = channel, x = pixel (coalesced plane reads)
    const int c = c0 + i;
    const int64_t pix = p0 + threadIdx.x;
    float v = 0.f;
    if (c < C && pix < HW) {
      v = src[(int64_t)c * HW + pix];
      if (apply) v = __fadd_rn(__fmul_rn(v, scale), offset);      // never contracted into an FMA
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {          // i = pixel, x = channel (coalesced HWC writes)
    const int c = c0 + threadIdx.x;
    const int64_t pix = p0 + i;
    if (c < C && pix < HW) dst[pix * C + c] = tile[threadIdx.x][i];
  }
}
extern "C" int sshslie_denorm_hwc(const float* src_chw, float* dst_hwc, int C, int H, int W, float scale, float offset,
                                  int apply, void* stream) {
  if (!src_chw || !dst_hwc || C < 1 || H < 1 || W < 1) {
    ss_set_error("sshslie_denorm_hwc: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  const int64_t HW = (int64_t)H * W;
  dim3 grid((unsigned)((HW + 31) / 32), (unsigned)((C + 31) / 32));
  denorm_hwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src_chw, dst_hwc, C, HW, scale, offset, apply);
  return ss_check_launch("denorm_hwc");
}

// ---------------------------------------------------------------------------------------------
// evaluation metrics on the device: PSNR and SAM as metrics.py:13-14,31-34 call them
// (torchmetrics 1.6.2 functional semantics, restated - the package is not available offline, parity unpinned):
//   sums[0] = sum (p - t)^2 over all elements          -> PSNR = 10 log10(range^2 / (sums[0] / n))
//   sums[1] = sum over pixels of acos(clamp(<p,t> / (|p| |t|), -1, 1))   -> SAM = sums[1] / pixels   (radians)
// pred / target are HWC fp32 cubes (what save_hsi wrote).  One warp per pixel; fp32 per pixel, fp64 across pixels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) psnr_sam_kernel(const float* __restrict__ pred, const float* __restrict__ tgt,
                                                       int64_t npix, int C, double* __restrict__ sums) {
  __shared__ double red[2][8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double sse = 0.0, ang = 0.0;
  for (int64_t pix = (int64_t)blockIdx.x * 8 + wid; pix < npix; pix += (int64_t)gridDim.x * 8) {
    float dot = 0.f, pp = 0.f, tt = 0.f, se = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float a = pred[pix * C + c], b = tgt[pix * C + c];
      dot = fmaf(a, b, dot);
      pp = fmaf(a, a, pp);
      tt = fmaf(b, b, tt);
      const float d = a - b;
      se = fmaf(d, d, se);
    }
    dot = warp_sum(dot); pp = warp_sum(pp); tt = warp_sum(tt); se = warp_sum(se);
    if (lane == 0) {
      sse += (double)se;
      const float cosv = fminf(fmaxf(dot / (sqrtf(pp) * sqrtf(tt)), -1.f), 1.f);
      ang += (double)acosf(cosv);
    }
  }
  if (lane == 0) { red[0][wid] = sse; red[1][wid] = ang; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += red[0][i]; b += red[1][i]; }
    atomicAdd(sums, a);
    atomicAdd(sums + 1, b);
  }
}
extern "C" int sshslie_psnr_sam(const float* pred_hwc, const float* target_hwc, int H, int W, int C, double* sums2,
                                void* stream) {
  if (!pred_hwc || !target_hwc || !sums2 || H < 1 || W < 1 || C < 1) {
    ss_set_error("sshslie_psnr_sam: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  const int64_t npix = (int64_t)H * W;
  cudaMemsetAsync(sums2, 0, 2 * sizeof(double), (cudaStream_t)stream);
  const unsigned blocks = (unsigned)std::min<int64_t>((npix + 7) / 8, 148 * 8);
  psnr_sam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred_hwc, target_hwc, npix, C, sums2);
  return ss_check_launch("psnr_sam");
}


// SSIM as metrics.py:16-19 evaluates it (torchmetrics on the (1,H,W,C) tensor: H plays the channel role, the 11x11 gaussian
// window slides over the (W, C) plane of every image row; reflect padding by 5 and the final crop by 5 cancel, so every
// averaged output is a plain window that lies inside the plane).  One thread per output (h, w, c), c fastest; fp32 per
// window, fp64 across outputs.  sum[0] += sum of the SSIM index over H x (W-10) x (C-10) outputs.
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, int H,
                                                   int W, int C, float c1, float c2, double* __restrict__ sum) {
  __shared__ float gw[11];
  __shared__ double red[8];
  if (threadIdx.x < 11) {
    float tot = 0.f;
    for (int i = 0; i < 11; ++i) tot += expf(-((float)(i - 5) / 1.5f) * ((float)(i - 5) / 1.5f) / 2.f);
    gw[threadIdx.x] = expf(-((float)((int)threadIdx.x - 5) / 1.5f) * ((float)((int)threadIdx.x - 5) / 1.5f) / 2.f) / tot;
  }
  __syncthreads();
  const int Wo = W - 10, Co = C - 10;
  const int64_t total = (int64_t)H * Wo * Co;
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % Co);
    const int64_t t = i / Co;
    const int w = (int)(t % Wo);
    const int h = (int)(t / Wo);
    const float* p0 = pred + ((int64_t)h * W + w) * C + c;
    const float* t0 = tgt + ((int64_t)h * W + w) * C + c;
    float mp = 0.f, mt = 0.f, epp = 0.f, ett = 0.f, ept = 0.f;
    for (int dw = 0; dw < 11; ++dw) {
      const float gy = gw[dw];
#pragma unroll
      for (int dc = 0; dc < 11; ++dc) {
        const float g = gy * gw[dc];
        const float a = p0[(int64_t)dw * C + dc], b = t0[(int64_t)dw * C + dc];
        mp = fmaf(g, a, mp);
        mt = fmaf(g, b, mt);
        epp = fmaf(g, a * a, epp);
        ett = fmaf(g, b * b, ett);
        ept = fmaf(g, a * b, ept);
      }
    }
    const float spp = fmaxf(epp - mp * mp, 0.f), stt = fmaxf(ett - mt * mt, 0.f), spt = ept - mp * mt;
    acc += (double)(((2.f * mp * mt + c1) * (2.f * spt + c2)) / ((mp * mp + mt * mt + c1) * (spp + stt + c2)));
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < 8; ++i) s += red[i];
    atomicAdd(sum, s);
  }
}
extern "C" int sshslie_ssim_sum(const float* pred_hwc, const float* target_hwc, int H, int W, int C, float c1, float c2,
                                double* sum1, void* stream) {
  if (!pred_hwc || !target_hwc || !sum1 || H < 1 || W < 11 || C < 11) {
    ss_set_error("sshslie_ssim_sum: need W >= 11 and C >= 11 (11x11 window over the (W, C) plane)");
    return SSHSLIE_ERR_ARG;
  }
  cudaMemsetAsync(sum1, 0, sizeof(double), (cudaStream_t)stream);
  const int64_t total = (int64_t)H * (W - 10) * (C - 10);
  const unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
  ssim_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(pred_hwc, target_hwc, H, W, C, c1, c2, sum1);
  return ss_check_launch("ssim");
}
