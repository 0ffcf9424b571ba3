// Shared GEMM epilogue: 16 consecutive output columns of one GEMM row (= one output pixel).
// Used identically by the tcgen05 kernel (conv_umma.cu) and the CUDA-core cross-check kernel (conv_simt.cu).
#pragma once
#include "common.cuh"

// bias_s: optional shared-memory copy of the bias (zero beyond N), staged while the accumulator is still being computed
SS_DEVINL void epi_apply16(const Epi& e, int b, int oh, int ow, int n0, int N, float* v /*[16]*/,
                           const float* bias_s = nullptr) {
  if (bias_s) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += bias_s[n0 + i];
  } else if (e.bias) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += (n0 + i < N) ? __ldg(e.bias + n0 + i) : 0.f;
  }
  if (e.mode == EPI_BF16 && e.n_split > 0 && n0 >= e.n_split) {      // second half of a split output: own tensor, own mask
    const int n1 = n0 - e.n_split;
    if (n1 >= e.n_store2) return;
    if (e.mask2) {
      const bf16* p = e.mask2 + b * e.m2B + oh * e.m2H + ow * e.m2W + n1;
      float m[16];
      unpack8(*reinterpret_cast<const uint4*>(p), m);
      unpack8(*reinterpret_cast<const uint4*>(p + 8), m + 8);
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = (m[i] > 0.f) ? v[i] : 0.f;
    }
    bf16* p = e.out2 + b * e.o2B + oh * e.o2H + ow * e.o2W + n1;
    uint4 lo, hi;
    lo.x = pack2(v[0], v[1]);  lo.y = pack2(v[2], v[3]);  lo.z = pack2(v[4], v[5]);   lo.w = pack2(v[6], v[7]);
    hi.x = pack2(v[8], v[9]);  hi.y = pack2(v[10], v[11]); hi.z = pack2(v[12], v[13]); hi.w = pack2(v[14], v[15]);
    *reinterpret_cast<uint4*>(p) = lo;
    *reinterpret_cast<uint4*>(p + 8) = hi;
    return;
  }
  if (e.add) {
    const bf16* p = e.add + b * e.aB + oh * e.aH + ow * e.aW + n0;
    float a[16];
    unpack8(*reinterpret_cast<const uint4*>(p), a);
    unpack8(*reinterpret_cast<const uint4*>(p + 8), a + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += a[i];
  }
  if (e.relu) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
  }
  if (e.mask) {
    const bf16* p = e.mask + b * e.mB + oh * e.mH + ow * e.mW + n0;
    float m[16];
    unpack8(*reinterpret_cast<const uint4*>(p), m);
    unpack8(*reinterpret_cast<const uint4*>(p + 8), m + 8);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = (m[i] > 0.f) ? v[i] : 0.f;
  }
  if (e.mode == EPI_BF16) {
    if (n0 >= e.n_store) return;
    bf16* p = e.out + b * e.oB + oh * e.oH + ow * e.oW + n0;
    uint4 lo, hi;
    lo.x = pack2(v[0], v[1]);  lo.y = pack2(v[2], v[3]);  lo.z = pack2(v[4], v[5]);   lo.w = pack2(v[6], v[7]);
    hi.x = pack2(v[8], v[9]);  hi.y = pack2(v[10], v[11]); hi.z = pack2(v[12], v[13]); hi.w = pack2(v[14], v[15]);
    *reinterpret_cast<uint4*>(p) = lo;
    *reinterpret_cast<uint4*>(p + 8) = hi;
    if (e.out_lo) {   // residual of the bf16 rounding: consumers read hi and lo as two K-slabs with the same weights
      float r[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) r[i] = v[i] - bf2f(f2bf(v[i]));
      bf16* q = e.out_lo + b * e.oB + oh * e.oH + ow * e.oW + n0;
      uint4 l0, l1;
      l0.x = pack2(r[0], r[1]);  l0.y = pack2(r[2], r[3]);  l0.z = pack2(r[4], r[5]);   l0.w = pack2(r[6], r[7]);
      l1.x = pack2(r[8], r[9]);  l1.y = pack2(r[10], r[11]); l1.z = pack2(r[12], r[13]); l1.w = pack2(r[14], r[15]);
      *reinterpret_cast<uint4*>(q) = l0;
      *reinterpret_cast<uint4*>(q + 8) = l1;
    }
  } else if (e.mode == EPI_HEAD) {
    const int64_t pix = ((int64_t)b * e.H + oh) * e.W + ow;
    float i_lo = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int n = n0 + i;
      const float s = sigmoidf_(v[i]);
      if (n < e.C) {
        e.R32[(((int64_t)b * e.C + n) * e.H + oh) * e.W + ow] = s;
      } else if (n == e.C && e.I32) {
        e.I32[pix] = s;
      }
      // lane C + 1 of cat[R, I]: I's bf16 residual (static register indices only: the lane follows I's in the same chunk)
      v[i] = (n <= e.C) ? s : ((n == e.C + 1 && e.ri_lo_off > 0) ? i_lo : 0.f);
      if (n == e.C) i_lo = s - bf2f(f2bf(s));
    }
    if (e.RI) {
      bf16* p = e.RI + pix * e.ri_c + n0;
      uint4 lo, hi;
      lo.x = pack2(v[0], v[1]);  lo.y = pack2(v[2], v[3]);  lo.z = pack2(v[4], v[5]);   lo.w = pack2(v[6], v[7]);
      hi.x = pack2(v[8], v[9]);  hi.y = pack2(v[10], v[11]); hi.z = pack2(v[12], v[13]); hi.w = pack2(v[14], v[15]);
      *reinterpret_cast<uint4*>(p) = lo;
      *reinterpret_cast<uint4*>(p + 8) = hi;
      if (e.ri_lo_off > 0 && n0 < e.C) {
        float r[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) r[i] = (n0 + i < e.C) ? v[i] - bf2f(f2bf(v[i])) : 0.f;
        bf16* q = p + e.ri_lo_off;
        uint4 l0, l1;
        l0.x = pack2(r[0], r[1]);  l0.y = pack2(r[2], r[3]);  l0.z = pack2(r[4], r[5]);   l0.w = pack2(r[6], r[7]);
        l1.x = pack2(r[8], r[9]);  l1.y = pack2(r[10], r[11]); l1.z = pack2(r[12], r[13]); l1.w = pack2(r[14], r[15]);
        *reinterpret_cast<uint4*>(q) = l0;
        *reinterpret_cast<uint4*>(q + 8) = l1;
      }
    }
  } else {  // EPI_PLANE32
    if (n0 == 0) e.plane32[((int64_t)b * e.H + oh) * e.W + ow] = v[0];
  }
}
