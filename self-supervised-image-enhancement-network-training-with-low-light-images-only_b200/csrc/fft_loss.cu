// fourier_spectrum_loss (model.py:456-473), forward value and d/dS in ONE kernel, one CTA per (b, band) image.
//
//   loss_sum += sum_k mask_k * | |X_k| - |S_k| |          X = fft2(x), S = fft2(s), unnormalised, full spectrum,
//                                                        mask on the UN-shifted grid
//   dS      += grad_scale * Re( IDFT_unnorm( -mask_k * sgn(|X_k| - |S_k|) * S_k / |S_k| ) )
//
// Both real images ride ONE complex transform: z = x + i*s, so X_k = (Z_k + conj(Z_-k))/2 and
// S_k = (Z_k - conj(Z_-k))/(2i).  The HxW complex tile lives in shared memory (128 KB at 128x128);
// forward = decimation-in-frequency along W then H, two radix-2 levels fused per sweep (natural in, bit-reversed out), the spectrum is
// consumed in bit-reversed positions, and the inverse runs decimation-in-time (bit-reversed in, natural out),
// so no reordering pass exists.  Algorithmic HBM traffic: read x, s (2 planes), read+write dS (1 plane each).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

#define FFT_THREADS 1024

SS_DEVINL float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SS_DEVINL int brev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }

// one radix-2 stage over the whole tile along the W axis (rows) or the H axis (columns)
//   dif:  a' = a + b ; b' = (a - b) * tw        dit:  b *= tw ; a' = a + b ; b' = a - b
template <bool ALONG_W, bool DIT, bool INV>
SS_DEVINL void fft_stage(float2* z, const float2* tw, int H, int W, int lgW, int lg_half, int lg_axis) {
  // all sizes are powers of two: index math is shifts and masks only (runtime div/mod was 80 % of the instructions)
  const int nbf = (H * W) >> 1;
  const int half = 1 << lg_half;
  const int lg_tstep = lg_axis - 1 - lg_half;
#pragma unroll 2
  for (int t = threadIdx.x; t < nbf; t += FFT_THREADS) {
    int i0, i1, pos;
    if (ALONG_W) {
      const int row = t >> (lgW - 1), j = t & ((W >> 1) - 1);
      const int grp = j >> lg_half;
      pos = j & (half - 1);
      i0 = (row << lgW) + (grp << (lg_half + 1)) + pos;
      i1 = i0 + half;
    } else {
      const int col = t & (W - 1), j = t >> lgW;      // consecutive lanes -> consecutive columns (conflict-free)
      const int grp = j >> lg_half;
      pos = j & (half - 1);
      i0 = (((grp << (lg_half + 1)) + pos) << lgW) + col;
      i1 = i0 + (half << lgW);
    }
    float2 w = tw[pos << lg_tstep];
    if (INV) w.y = -w.y;
    float2 a = z[i0], b = z[i1];
    if (DIT) {
      b = cmul(b, w);
      z[i0] = make_float2(a.x + b.x, a.y + b.y);
      z[i1] = make_float2(a.x - b.x, a.y - b.y);
    } else {
      z[i0] = make_float2(a.x + b.x, a.y + b.y);
      z[i1] = cmul(make_float2(a.x - b.x, a.y - b.y), w);
    }
  }
  __syncthreads();
}

// two radix-2 levels fused in registers (a radix-4 pass): halves the shared-memory sweeps and barriers.
//   DIF: levels (h, h/2) with lg_h = lg_first;   DIT: levels (h, 2h) with lg_h = lg_first.
template <bool ALONG_W, bool DIT, bool INV>
SS_DEVINL void fft_stage2(float2* z, const float2* tw, int H, int W, int lgW, int lg_first, int lg_axis) {
  const int nq = (H * W) >> 2;
  const int es = ALONG_W ? 1 : W;                       // element stride along the transformed axis
#pragma unroll 2
  for (int t = threadIdx.x; t < nq; t += FFT_THREADS) {
    int line, j;
    if (ALONG_W) { line = (t >> (lgW - 2)) << lgW; j = t & ((W >> 2) - 1); }
    else { line = t & (W - 1); j = t >> lgW; }
    if (!DIT) {
      const int h = 1 << lg_first, hh = h >> 1;         // levels h then h/2
      const int grp = j >> (lg_first - 1), pos = j & (hh - 1);
      const int base = line + ((grp << (lg_first + 1)) + pos) * es;
      const int lt1 = lg_axis - 1 - lg_first;           // twiddle stride of level h; level h/2 uses lt1 + 1
      float2 w1a = tw[pos << lt1], w1b = tw[(pos + hh) << lt1], w2 = tw[pos << (lt1 + 1)];
      if (INV) { w1a.y = -w1a.y; w1b.y = -w1b.y; w2.y = -w2.y; }
      const float2 e0 = z[base], e1 = z[base + hh * es], e2 = z[base + h * es], e3 = z[base + (h + hh) * es];
      const float2 a0 = make_float2(e0.x + e2.x, e0.y + e2.y);
      const float2 a2 = cmul(make_float2(e0.x - e2.x, e0.y - e2.y), w1a);
      const float2 a1 = make_float2(e1.x + e3.x, e1.y + e3.y);
      const float2 a3 = cmul(make_float2(e1.x - e3.x, e1.y - e3.y), w1b);
      z[base] = make_float2(a0.x + a1.x, a0.y + a1.y);
      z[base + hh * es] = cmul(make_float2(a0.x - a1.x, a0.y - a1.y), w2);
      z[base + h * es] = make_float2(a2.x + a3.x, a2.y + a3.y);
      z[base + (h + hh) * es] = cmul(make_float2(a2.x - a3.x, a2.y - a3.y), w2);
    } else {
      const int h = 1 << lg_first;                      // levels h then 2h
      const int grp = j >> lg_first, pos = j & (h - 1);
      const int base = line + ((grp << (lg_first + 2)) + pos) * es;
      const int lt1 = lg_axis - 1 - lg_first;           // level h; level 2h uses lt1 - 1
      float2 w1 = tw[pos << lt1], w2a = tw[pos << (lt1 - 1)], w2b = tw[(pos + h) << (lt1 - 1)];
      if (INV) { w1.y = -w1.y; w2a.y = -w2a.y; w2b.y = -w2b.y; }
      const float2 e0 = z[base], e1 = cmul(z[base + h * es], w1), e2 = z[base + 2 * h * es],
                   e3 = cmul(z[base + 3 * h * es], w1);
      const float2 a0 = make_float2(e0.x + e1.x, e0.y + e1.y), a1 = make_float2(e0.x - e1.x, e0.y - e1.y);
      const float2 a2 = cmul(make_float2(e2.x + e3.x, e2.y + e3.y), w2a);
      const float2 a3 = cmul(make_float2(e2.x - e3.x, e2.y - e3.y), w2b);
      z[base] = make_float2(a0.x + a2.x, a0.y + a2.y);
      z[base + 2 * h * es] = make_float2(a0.x - a2.x, a0.y - a2.y);
      z[base + h * es] = make_float2(a1.x + a3.x, a1.y + a3.y);
      z[base + 3 * h * es] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
  }
  __syncthreads();
}

// full 1-D transform along one axis of the tile: pairs of levels fused, one single level if the count is odd
template <bool ALONG_W, bool DIT, bool INV>
SS_DEVINL void fft_axis(float2* z, const float2* tw, int H, int W, int lgW, int lg_axis) {
  if (!DIT) {
    int lh = lg_axis - 1;
    for (; lh >= 1; lh -= 2) fft_stage2<ALONG_W, false, INV>(z, tw, H, W, lgW, lh, lg_axis);
    if (lh == 0) fft_stage<ALONG_W, false, INV>(z, tw, H, W, lgW, 0, lg_axis);
  } else {
    int lh = 0;
    for (; lh + 1 < lg_axis; lh += 2) fft_stage2<ALONG_W, true, INV>(z, tw, H, W, lgW, lh, lg_axis);
    if (lh < lg_axis) fft_stage<ALONG_W, true, INV>(z, tw, H, W, lgW, lh, lg_axis);
  }
}

__global__ void __launch_bounds__(FFT_THREADS, 1)
fourier_loss_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ mask,
                    float* __restrict__ dS, float* __restrict__ partial_out, int H, int W, int lgH, int lgW,
                    float grad_scale, int accumulate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);                    // H*W
  float2* twW = z + H * W;                                            // W/2
  float2* twH = twW + 64;                                             // H/2
  __shared__ float red[32];
  const int64_t img = blockIdx.x;
  const float* xp = x + img * H * W;
  const float* sp = s + img * H * W;

  for (int i = threadIdx.x; i < W / 2; i += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.f * (float)i / (float)W, &sn, &cs);
    twW[i] = make_float2(cs, sn);
  }
  for (int i = threadIdx.x; i < H / 2; i += FFT_THREADS) {
    float sn, cs;
    sincospif(-2.f * (float)i / (float)H, &sn, &cs);
    twH[i] = make_float2(cs, sn);
  }
  // both planes -> z = x + i s.  Batches of 8 + 8 loads in flight per thread (a load/store loop pays the HBM latency per
  // iteration), float4 per lane
  {
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    const float4* s4 = reinterpret_cast<const float4*>(sp);
    const int n4 = (H * W) >> 2;
    for (int i0 = threadIdx.x; i0 < n4; i0 += 4 * FFT_THREADS) {
      float4 xv[4], sv[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * FFT_THREADS;
        if (i < n4) { xv[k] = __ldg(x4 + i); sv[k] = __ldg(s4 + i); }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = i0 + k * FFT_THREADS;
        if (i < n4) {
          float4* zo = reinterpret_cast<float4*>(z + 4 * i);
          zo[0] = make_float4(xv[k].x, sv[k].x, xv[k].y, sv[k].y);
          zo[1] = make_float4(xv[k].z, sv[k].z, xv[k].w, sv[k].w);
        }
      }
    }
  }
  __syncthreads();

  // forward: DIF along W, then along H
  fft_axis<true, false, false>(z, twW, H, W, lgW, lgW);
  fft_axis<false, false, false>(z, twH, H, W, lgW, lgH);

  // spectrum pass: position (ph,pw) holds frequency (brev(ph), brev(pw)); pair it with -k
  float lsum = 0.f;
  for (int p = threadIdx.x; p < H * W; p += FFT_THREADS) {
    const int ph = p >> lgW, pw = p & (W - 1);
    const int ky = brev(ph, lgH), kx = brev(pw, lgW);
    const int nky = (H - ky) & (H - 1), nkx = (W - kx) & (W - 1);
    const int pn = brev(nky, lgH) * W + brev(nkx, lgW);
    if (pn < p) continue;                        // the pair is owned by its smaller position
    const float2 a = z[p], c = z[pn];
    // X_k = (Z_k + conj(Z_-k))/2 ; S_k = (Z_k - conj(Z_-k))/(2i)
    const float2 X = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
    const float2 S = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
    const float ax = sqrtf(X.x * X.x + X.y * X.y);
    const float as = sqrtf(S.x * S.x + S.y * S.y);
    const float mk = mask[ky * W + kx], mn = mask[nky * W + nkx];
    const float diff = ax - as;
    const float ad = fabsf(diff);
    lsum += (pn == p) ? mk * ad : (mk + mn) * ad;
    // dLoss/d|S_k| = -mask*sgn(diff);  Y_k = that * S_k/|S_k| ;  S_-k = conj(S_k)
    const float g = (as > 0.f) ? -sgnf(diff) / as : 0.f;
    z[p] = make_float2(mk * g * S.x, mk * g * S.y);
    if (pn != p) z[pn] = make_float2(mn * g * S.x, -mn * g * S.y);
  }
  __syncthreads();
  {
    const float t = block_sum(lsum, red);
    if (threadIdx.x == 0) partial_out[blockIdx.x] = t;        // one partial per plane, reduced in a fixed order later
  }
  if (dS == nullptr) return;

  // inverse: DIT along H then W with conjugated twiddles (bit-reversed in, natural out)
  fft_axis<false, true, true>(z, twH, H, W, lgW, lgH);
  fft_axis<true, true, true>(z, twW, H, W, lgW, lgW);

  float* dp = dS + img * H * W;
  if (accumulate) {
    for (int i = threadIdx.x; i < H * W; i += FFT_THREADS) dp[i] += grad_scale * z[i].x;
  } else {
    for (int i = threadIdx.x; i < H * W; i += FFT_THREADS) dp[i] = grad_scale * z[i].x;
  }
}

// ---------------------------------------------------------------------------------------------
// 128 x 128 planes (the training patch of every reference config): register FFT.
// A 128-point line transform is 16 x 8:  k = k1 + 16 k2,  n = j + 8 m
//     X[k1 + 16 k2] = sum_j W8^(j k2) [ W128^(j k1) sum_m W16^(m k1) x[j + 8 m] ]
//   step 1  thread (line, j):   16-point DFT over m in registers, twiddle W128^(j k1), stored in place at j + 8 k1
//   step 2  thread (line, k1):  8-point DFT over j (8 consecutive elements), stored at k1 + 16 k2  -> natural order
// Lanes run over LINES, so with a row pitch of 129 complex words both the row pass (line = image row, element stride 1)
// and the column pass (line = column, element stride = pitch) are free of bank conflicts, and the whole 2-D transform is
// 4 passes x 2 barriers instead of 16 radix-4 sweeps over shared memory.  Natural order in and out (no bit reversal).
// ---------------------------------------------------------------------------------------------
#define F128_N 128
#define F128_PITCH 129
#define F128_THREADS 1024

template <bool INV>
SS_DEVINL float2 tw_apply(float2 a, float c, float s_) {      // a * (c - i s) forward, a * (c + i s) inverse
  const float s2 = INV ? -s_ : s_;
  return make_float2(a.x * c + a.y * s2, a.y * c - a.x * s2);
}
// in-register radix-2 DIF of length N (power of two <= 16) with compile-time twiddles; output index k ends up at
// position bit-reverse(k): callers read v[BR(k)].
template <int N, bool INV>
SS_DEVINL void dft_regs(float2 (&v)[N]) {
  constexpr float C16[8] = {1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                            0.f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
  constexpr float S16[8] = {0.f, 0.38268343236508977f, 0.70710678118654752f, 0.92387953251128674f,
                            1.f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f};
#pragma unroll
  for (int half = N / 2; half >= 1; half >>= 1) {
#pragma unroll
    for (int g = 0; g < N; g += 2 * half) {
#pragma unroll
      for (int p = 0; p < half; ++p) {
        const float2 a = v[g + p], b = v[g + p + half];
        v[g + p] = make_float2(a.x + b.x, a.y + b.y);
        const float2 d = make_float2(a.x - b.x, a.y - b.y);
        const int tw = p * (8 / half);                 // W_(2 half)^p = W16^(p * 16 / (2 half))
        if (tw == 0) v[g + p + half] = d;
        else if (tw == 4) v[g + p + half] = INV ? make_float2(-d.y, d.x) : make_float2(d.y, -d.x);
        else v[g + p + half] = tw_apply<INV>(d, C16[tw], S16[tw]);
      }
    }
  }
}
SS_DEVINL constexpr int br4(int k) { return ((k & 1) << 3) | ((k & 2) << 1) | ((k & 4) >> 1) | ((k & 8) >> 3); }
SS_DEVINL constexpr int br3(int k) { return ((k & 1) << 2) | (k & 2) | ((k & 4) >> 2); }

// one pass over all 128 lines: es = element stride along the line, ls = stride between lines (in float2 words)
template <bool INV>
SS_DEVINL void fft128_lines(float2* __restrict__ z, const float2* __restrict__ tw128, int es, int ls) {
  const int line = threadIdx.x & (F128_N - 1), j = threadIdx.x >> 7;          // j = 0..7
  {
    float2* base = z + line * ls + j * es;
    float2 v[16];
#pragma unroll
    for (int m = 0; m < 16; ++m) v[m] = base[(8 * m) * es];
    dft_regs<16, INV>(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      float2 r = v[br4(k1)];
      if (k1 > 0 && j > 0) {
        const float2 w = tw128[(j * k1) & (F128_N - 1)];      // (cos, sin) of 2 pi j k1 / 128
        r = tw_apply<INV>(r, w.x, w.y);
      }
      base[(8 * k1) * es] = r;                                  // in place: position j + 8 k1
    }
  }
  __syncthreads();
  float2 u[2][8];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int k1 = j + 8 * it;                                  // this thread's two k1 values
    const float2* src = z + line * ls + (8 * k1) * es;
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) u[it][jj] = src[jj * es];
    dft_regs<8, INV>(u[it]);
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int k1 = j + 8 * it;
    float2* dst = z + line * ls + k1 * es;
#pragma unroll
    for (int k2 = 0; k2 < 8; ++k2) dst[(16 * k2) * es] = u[it][br3(k2)];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(F128_THREADS, 1)
fourier_loss128_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ mask,
                       float* __restrict__ dS, float* __restrict__ partial_out, float grad_scale, int accumulate) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* z = reinterpret_cast<float2*>(smem_raw);                    // [128][F128_PITCH]
  float2* tw128 = z + F128_N * F128_PITCH;                            // (cos, sin)(2 pi i / 128)
  __shared__ float red[32];
  constexpr int HW = F128_N * F128_N;
  const int64_t img = blockIdx.x;
  const float* xp = x + img * HW;
  const float* sp = s + img * HW;
  if (threadIdx.x < F128_N) {
    float sn, cs;
    sincospif(2.f * (float)threadIdx.x / (float)F128_N, &sn, &cs);
    tw128[threadIdx.x] = make_float2(cs, sn);
  }
  {  // z = x + i s, all loads in flight before the stores
    const float4* x4 = reinterpret_cast<const float4*>(xp);
    const float4* s4 = reinterpret_cast<const float4*>(sp);
    float4 xv[4], sv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { xv[k] = __ldg(x4 + threadIdx.x + k * F128_THREADS); sv[k] = __ldg(s4 + threadIdx.x + k * F128_THREADS); }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = (threadIdx.x + k * F128_THREADS) * 4;             // first of 4 consecutive pixels of a row
      float2* zo = z + (i >> 7) * F128_PITCH + (i & 127);
      zo[0] = make_float2(xv[k].x, sv[k].x); zo[1] = make_float2(xv[k].y, sv[k].y);
      zo[2] = make_float2(xv[k].z, sv[k].z); zo[3] = make_float2(xv[k].w, sv[k].w);
    }
  }
  __syncthreads();
  fft128_lines<false>(z, tw128, 1, F128_PITCH);        // rows (along W)
  fft128_lines<false>(z, tw128, F128_PITCH, 1);        // columns (along H)

  // spectrum pass, natural order: frequency (ky, kx) pairs with (-ky, -kx)
  float lsum = 0.f;
  for (int p = threadIdx.x; p < HW; p += F128_THREADS) {
    const int ky = p >> 7, kx = p & 127;
    const int nky = (F128_N - ky) & 127, nkx = (F128_N - kx) & 127;
    const int pn = nky * F128_N + nkx;
    if (pn < p) continue;                        // the pair is owned by its smaller index
    float2* za = z + ky * F128_PITCH + kx;
    float2* zc = z + nky * F128_PITCH + nkx;
    const float2 a = *za, c = *zc;
    // X_k = (Z_k + conj(Z_-k))/2 ; S_k = (Z_k - conj(Z_-k))/(2i)
    const float2 X = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
    const float2 S = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
    const float ax = sqrtf(X.x * X.x + X.y * X.y);
    const float as = sqrtf(S.x * S.x + S.y * S.y);
    const float mk = mask[p], mn = mask[pn];
    const float diff = ax - as;
    const float ad = fabsf(diff);
    lsum += (pn == p) ? mk * ad : (mk + mn) * ad;
    const float g = (as > 0.f) ? -sgnf(diff) / as : 0.f;
    *za = make_float2(mk * g * S.x, mk * g * S.y);
    if (pn != p) *zc = make_float2(mn * g * S.x, -mn * g * S.y);
  }
  __syncthreads();
  {
    const float t = block_sum(lsum, red);
    if (threadIdx.x == 0) partial_out[blockIdx.x] = t;        // one partial per plane, reduced in a fixed order later
  }
  if (dS == nullptr) return;
  fft128_lines<true>(z, tw128, F128_PITCH, 1);         // inverse (unnormalised): columns, then rows
  fft128_lines<true>(z, tw128, 1, F128_PITCH);
  float* dp = dS + img * HW;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = (threadIdx.x + k * F128_THREADS) * 4;
    const float2* zi = z + (i >> 7) * F128_PITCH + (i & 127);
    float4 o = make_float4(grad_scale * zi[0].x, grad_scale * zi[1].x, grad_scale * zi[2].x, grad_scale * zi[3].x);
    float4* d4 = reinterpret_cast<float4*>(dp + i);
    if (accumulate) {
      const float4 old = *d4;
      o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
    }
    *d4 = o;
  }
}

// ---------------------------------------------------------------------------------------------
// Any other plane size (patches of 96, 256, ... - the reference's patch_size is a free config value, config/*.yml:12):
// the same loss through plain DFTs, O(N) per output, with the complex plane in a global workspace.  Four line passes
// (rows, columns, spectrum, columns^-1, rows^-1) replace the shared-memory transform; the spectrum arithmetic is the same.
// A block takes 32 lines into shared memory (odd pitch: lanes run over LINES, conflict-free), each warp then produces
// DFT_KB outputs per sweep with a warp-uniform (broadcast) twiddle.  A correctness-first path: ~2 ms at B=2 x 256 x 256.
// ---------------------------------------------------------------------------------------------
#define DFT_LINES 32
#define DFT_WARPS 8
#define DFT_KB 4
// MODE 0: complex in (zin) -> complex out (zout, may alias zin: a block owns its lines)
// MODE 1: real pair (x, s) in -> complex out          MODE 2: complex in -> real out: dS (+)= scale * Re
// element (line, n) lives at (line / lpp) * plane + (line % lpp) * ls + n * es
template <bool INV, int MODE>
__global__ void __launch_bounds__(32 * DFT_WARPS) dft_lines_kernel(const float2* __restrict__ zin, const float* __restrict__ xr,
                                                                   const float* __restrict__ sr, float2* __restrict__ zout,
                                                                   float* __restrict__ dS, int n_lines, int N, int lpp,
                                                                   int64_t plane, int ls, int es, float scale,
                                                                   int accumulate) {
  extern __shared__ __align__(16) unsigned char dft_smem[];
  float2* tile = reinterpret_cast<float2*>(dft_smem);          // [DFT_LINES][pitch]
  const int pitch = N | 1;
  float2* tw = tile + DFT_LINES * pitch;                       // (cos, sin)(2 pi n / N)
  const int tid = threadIdx.x, lane = tid & 31, wj = tid >> 5;
  const int line0 = blockIdx.x * DFT_LINES;
  for (int n = tid; n < N; n += 32 * DFT_WARPS) {
    float sn, cs;
    sincospif(2.f * (float)n / (float)N, &sn, &cs);
    tw[n] = make_float2(cs, sn);
  }
  const bool lines_fast = (ls == 1);                           // adjacent lines are adjacent in memory (column pass)
  for (int i = tid; i < DFT_LINES * N; i += 32 * DFT_WARPS) {
    const int l = lines_fast ? (i & (DFT_LINES - 1)) : (i / N);
    const int n = lines_fast ? (i / DFT_LINES) : (i - l * N);
    const int line = line0 + l;
    float2 v = make_float2(0.f, 0.f);
    if (line < n_lines) {
      const int64_t a = (int64_t)(line / lpp) * plane + (int64_t)(line % lpp) * ls + (int64_t)n * es;
      v = (MODE == 1) ? make_float2(__ldg(xr + a), __ldg(sr + a)) : zin[a];
    }
    tile[l * pitch + n] = v;
  }
  __syncthreads();
  const int line = line0 + lane;
  const int64_t base = (int64_t)(line / lpp) * plane + (int64_t)(line % lpp) * ls;
  const float2* row = tile + lane * pitch;
  for (int kb = wj * DFT_KB; kb < N; kb += DFT_WARPS * DFT_KB) {
    float2 acc[DFT_KB];
    int idx[DFT_KB];
#pragma unroll
    for (int q = 0; q < DFT_KB; ++q) { acc[q] = make_float2(0.f, 0.f); idx[q] = 0; }
    for (int n = 0; n < N; ++n) {
      const float2 v = row[n];
#pragma unroll
      for (int q = 0; q < DFT_KB; ++q) {
        const float2 w = tw[idx[q]];                           // warp-uniform index: one broadcast read
        const float wy = INV ? w.y : -w.y;                     // forward: e^{-i t}, inverse: e^{+i t}
        acc[q].x += v.x * w.x - v.y * wy;
        acc[q].y += v.x * wy + v.y * w.x;
        idx[q] += kb + q;
        if (idx[q] >= N) idx[q] -= N;
      }
    }
    if (line < n_lines) {
#pragma unroll
      for (int q = 0; q < DFT_KB; ++q) {
        if (kb + q >= N) continue;
        const int64_t a = base + (int64_t)(kb + q) * es;
        if (MODE == 2) dS[a] = accumulate ? dS[a] + scale * acc[q].x : scale * acc[q].x;
        else zout[a] = acc[q];
      }
    }
  }
}

// spectrum pass of the DFT path, one block per plane, natural order (same arithmetic as the shared-memory kernels)
__global__ void __launch_bounds__(1024) dft_spectrum_kernel(float2* __restrict__ z, const float* __restrict__ mask,
                                                            float* __restrict__ partial_out, int H, int W) {
  __shared__ float red[32];
  float2* zp = z + (int64_t)blockIdx.x * H * W;
  float lsum = 0.f;
  for (int p = threadIdx.x; p < H * W; p += 1024) {
    const int ky = p / W, kx = p - ky * W;
    const int nky = ky ? H - ky : 0, nkx = kx ? W - kx : 0;
    const int pn = nky * W + nkx;
    if (pn < p) continue;                        // the pair is owned by its smaller index
    const float2 a = zp[p], c = zp[pn];
    const float2 X = make_float2(0.5f * (a.x + c.x), 0.5f * (a.y - c.y));
    const float2 S = make_float2(0.5f * (a.y + c.y), -0.5f * (a.x - c.x));
    const float ax = sqrtf(X.x * X.x + X.y * X.y);
    const float as = sqrtf(S.x * S.x + S.y * S.y);
    const float mk = mask[p], mn = mask[pn];
    const float diff = ax - as;
    const float ad = fabsf(diff);
    lsum += (pn == p) ? mk * ad : (mk + mn) * ad;
    const float g = (as > 0.f) ? -sgnf(diff) / as : 0.f;
    zp[p] = make_float2(mk * g * S.x, mk * g * S.y);
    if (pn != p) zp[pn] = make_float2(mn * g * S.x, -mn * g * S.y);
  }
  const float t = block_sum(lsum, red);
  if (threadIdx.x == 0) partial_out[blockIdx.x] = t;
}

static bool fft_in_smem(int H, int W) {
  return H >= 8 && W >= 8 && H <= 128 && W <= 128 && (H & (H - 1)) == 0 && (W & (W - 1)) == 0;
}
// floats of global workspace ss_fourier_loss needs for this plane size (0: the transform fits shared memory)
int64_t ss_fourier_work_floats(int n_img, int H, int W) {
  return fft_in_smem(H, W) ? 0 : (int64_t)2 * n_img * H * W;
}
#define DFT_MAX_N 1024
template <bool INV, int MODE>
static int dft_pass(const float2* zin, const float* xr, const float* sr, float2* zout, float* dS, int n_lines, int N, int lpp,
                    int64_t plane, int ls, int es, float scale, int accumulate, cudaStream_t st) {
  const size_t smem = ((size_t)DFT_LINES * (N | 1) + N) * sizeof(float2);
  static DeviceOnce once;
  if (!once.done()) {
    if (cudaFuncSetAttribute(dft_lines_kernel<INV, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024) !=
        cudaSuccess) {
      ss_set_error("sshslie_fourier_loss: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    once.set();
  }
  dft_lines_kernel<INV, MODE><<<(n_lines + DFT_LINES - 1) / DFT_LINES, 32 * DFT_WARPS, smem, st>>>(
      zin, xr, sr, zout, dS, n_lines, N, lpp, plane, ls, es, scale, accumulate);
  return ss_check_launch("dft_lines");
}
static int fourier_loss_dft(const float* x, const float* S, const float* mask, float* dS, float* partial_out, float2* z,
                            int n_img, int H, int W, float grad_scale, int accumulate, cudaStream_t st) {
  const int64_t plane = (int64_t)H * W;
  int rc = dft_pass<false, 1>(nullptr, x, S, z, nullptr, n_img * H, W, H, plane, W, 1, 0.f, 0, st);         // rows
  if (!rc) rc = dft_pass<false, 0>(z, nullptr, nullptr, z, nullptr, n_img * W, H, W, plane, 1, W, 0.f, 0, st);   // columns
  if (rc) return rc;
  dft_spectrum_kernel<<<n_img, 1024, 0, st>>>(z, mask, partial_out, H, W);
  rc = ss_check_launch("dft_spectrum");
  if (rc || !dS) return rc;
  rc = dft_pass<true, 0>(z, nullptr, nullptr, z, nullptr, n_img * W, H, W, plane, 1, W, 0.f, 0, st);
  if (!rc) rc = dft_pass<true, 2>(z, nullptr, nullptr, nullptr, dS, n_img * H, W, H, plane, W, 1, grad_scale, accumulate, st);
  return rc;
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

// accumulate = 1: dS += gradient (the exported entry point); 0: dS = gradient (the engine gives the term its own plane so
// that it can run beside the second decomposition pass instead of after the pixel-space terms)
// partial_out[n_img]: the masked-magnitude L1 sum of every plane (the caller adds them up in a fixed order)
int ss_fourier_loss(const float* x, const float* S, const float* mask, float* dS, float* partial_out, int n_img, int H,
                    int W, float grad_scale, int accumulate, float* work, cudaStream_t stream) {
  if (!x || !S || !mask || !partial_out || n_img < 1 || H < 2 || W < 2 || H > DFT_MAX_N || W > DFT_MAX_N) {
    ss_set_error("sshslie_fourier_loss: bad argument (planes of 2..%d pixels a side; got %dx%d)", DFT_MAX_N, H, W);
    return SSHSLIE_ERR_ARG;
  }
  if (!fft_in_smem(H, W)) {
    if (!work) {
      ss_set_error("sshslie_fourier_loss: %dx%d planes need ss_fourier_work_floats of workspace", H, W);
      return SSHSLIE_ERR_WORKSPACE;
    }
    return fourier_loss_dft(x, S, mask, dS, partial_out, reinterpret_cast<float2*>(work), n_img, H, W, grad_scale,
                            accumulate, stream);
  }
  const int lgH = ilog2_exact(H), lgW = ilog2_exact(W);
  static const bool reg_fft = !(getenv("SSHSLIE_FFT128") && getenv("SSHSLIE_FFT128")[0] == '0');
  if (reg_fft && H == F128_N && W == F128_N && ((uintptr_t)x & 15) == 0 && ((uintptr_t)S & 15) == 0 &&
      (!dS || ((uintptr_t)dS & 15) == 0)) {
    const size_t smem128 = (size_t)(F128_N * F128_PITCH + F128_N) * sizeof(float2);
    static DeviceOnce attr128;
    if (!attr128.done()) {
      if (cudaFuncSetAttribute(fourier_loss128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem128) !=
          cudaSuccess) {
        ss_set_error("sshslie_fourier_loss: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
        return SSHSLIE_ERR_CUDA;
      }
      attr128.set();
    }
    fourier_loss128_kernel<<<n_img, F128_THREADS, smem128, stream>>>(x, S, mask, dS, partial_out, grad_scale, accumulate);
    return ss_check_launch("fourier_loss128");
  }
  const size_t smem = (size_t)H * W * sizeof(float2) + 128 * sizeof(float2);
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(fourier_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      ss_set_error("sshslie_fourier_loss: cannot raise dynamic shared memory: %s",
                   cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  fourier_loss_kernel<<<n_img, FFT_THREADS, smem, stream>>>(x, S, mask, dS, partial_out, H, W, lgH, lgW, grad_scale,
                                                            accumulate);
  return ss_check_launch("fourier_loss");
}
extern "C" int sshslie_fourier_loss(const float* x, const float* S, const float* mask, float* dS, float* sum_out,
                                    int n_img, int H, int W, float grad_scale, void* scratch, int64_t scratch_bytes,
                                    void* stream) {
  // scratch: n_img partial sums (rounded up to 4 floats), then the complex planes of the DFT path when the size needs it
  const int64_t head = ((int64_t)n_img + 3) / 4 * 4;
  const int64_t need = (head + ss_fourier_work_floats(n_img, H, W)) * (int64_t)sizeof(float);
  if (!sum_out || !scratch || scratch_bytes < need) {
    ss_set_error("sshslie_fourier_loss: need sum_out and %lld bytes of scratch (sshslie_loss_scratch_bytes)", (long long)need);
    return SSHSLIE_ERR_WORKSPACE;
  }
  const int rc = ss_fourier_loss(x, S, mask, dS, (float*)scratch, n_img, H, W, grad_scale, 1, (float*)scratch + head,
                                 (cudaStream_t)stream);
  if (rc) return rc;
  return ss_reduce_partials((const float*)scratch, n_img, 1, sum_out, 1, (cudaStream_t)stream);
}
