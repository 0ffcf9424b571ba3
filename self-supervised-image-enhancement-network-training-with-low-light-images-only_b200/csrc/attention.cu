// TransformerBlock (model.py:87-119) forward and backward on the H/8 x W/8 token grid, fp32.
//   tokens: T = B*L rows of 64 features (the bf16 NHWC activation a3 IS the (B, L, 64) token matrix)
//   Q,K,V = Linear(64,64) ; 4 heads x 16 ; softmax(QK^T / 4) V ; y = x + W2 relu(W1 o + b1) + b2
// Streaming (flash-style) softmax: the (B,4,L,L) logits of model.py:111-113 are never materialised.
// FLOPs are negligible at train size (L = 256); at 512^2 inference L = 4096 and QK^T/PV is ~1% of the step.
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

#define AT_D 64
#define AT_HEADS 4
#define AT_HD 16

// ---------------------------------------------------------------------------------------------
// fused token kernels (16 tokens per block, 256 threads = 64 features x 4 token phases)
// ---------------------------------------------------------------------------------------------
#define TK 16
SS_DEVINL void stage_w(float (*Ws)[AT_D + 1], const float* __restrict__ Wt) {
  float r[16];          // all 16 loads of the thread in flight before the first store (L2 latency paid once, not 16x)
#pragma unroll
  for (int k = 0; k < 16; ++k) r[k] = __ldg(Wt + threadIdx.x + k * 256);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int i = threadIdx.x + k * 256;
    Ws[i >> 6][i & 63] = r[k];
  }
}

// h = relu(o W1^T + b1) ; t = x + h W2^T + b2  -> h (kept for backward), t as bf16     (model.py:115-118).  Stand-alone FFN
// behind the tensor-core attention core (attention_tc.cu); at the training size the FFN is fused into attn_core_ffn_kernel.
__global__ void __launch_bounds__(256) attn_ffn_kernel(const float* __restrict__ O, const float* __restrict__ X,
                                                       const float* __restrict__ P, int64_t o1, int64_t ob1, int64_t o2,
                                                       int64_t ob2, float* __restrict__ Hh, bf16* __restrict__ Tout,
                                                       int T) {
  SS_PDL_ENTRY();
  extern __shared__ float smf[];
  float (*W1)[AT_D + 1] = reinterpret_cast<float (*)[AT_D + 1]>(smf);
  float (*W2)[AT_D + 1] = W1 + AT_D;
  float (*Os)[AT_D] = reinterpret_cast<float (*)[AT_D]>(smf + 2 * AT_D * (AT_D + 1));
  float (*Hs)[AT_D] = Os + TK;
  const int t0 = blockIdx.x * TK;
  stage_w(W1, P + o1); stage_w(W2, P + o2);
  for (int i = threadIdx.x; i < TK * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    Os[i >> 6][i & 63] = (t < T) ? O[(int64_t)t * AT_D + (i & 63)] : 0.f;
  }
  __syncthreads();
  const int o = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    float acc = P[ob1 + o];
#pragma unroll 16
    for (int i = 0; i < AT_D; ++i) acc = fmaf(Os[tt][i], W1[o][i], acc);
    acc = fmaxf(acc, 0.f);
    Hs[tt][o] = acc;
    if (t0 + tt < T) Hh[(int64_t)(t0 + tt) * AT_D + o] = acc;
  }
  __syncthreads();
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc = P[ob2 + o];
#pragma unroll 16
    for (int i = 0; i < AT_D; ++i) acc = fmaf(Hs[tt][i], W2[o][i], acc);
    Tout[(int64_t)t * AT_D + o] = f2bf(acc + X[(int64_t)t * AT_D + o]);
  }
}

// backward of the FFN, data path only:  dh = dt W2 (kept pre-mask), dO = (dh * [h>0]) W1
__global__ void __launch_bounds__(256) attn_ffn_bwd_kernel(const float* __restrict__ dT, const float* __restrict__ Hh,
                                                           const float* __restrict__ P, int64_t o1, int64_t o2,
                                                           float* __restrict__ dH, float* __restrict__ dO, int T) {
  SS_PDL_ENTRY();
  extern __shared__ float smf[];
  float (*W1)[AT_D + 1] = reinterpret_cast<float (*)[AT_D + 1]>(smf);
  float (*W2)[AT_D + 1] = W1 + AT_D;
  float (*Ys)[AT_D] = reinterpret_cast<float (*)[AT_D]>(smf + 2 * AT_D * (AT_D + 1));
  float (*Ds)[AT_D] = Ys + TK;
  const int t0 = blockIdx.x * TK;
  stage_w(W1, P + o1); stage_w(W2, P + o2);
  for (int i = threadIdx.x; i < TK * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    Ys[i >> 6][i & 63] = (t < T) ? dT[(int64_t)t * AT_D + (i & 63)] : 0.f;
  }
  __syncthreads();
  const int ii = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    float acc = 0.f;
#pragma unroll 16
    for (int o = 0; o < AT_D; ++o) acc = fmaf(Ys[tt][o], W2[o][ii], acc);
    const int t = t0 + tt;
    float hm = 0.f;
    if (t < T) {
      dH[(int64_t)t * AT_D + ii] = acc;
      hm = Hh[(int64_t)t * AT_D + ii];
    }
    Ds[tt][ii] = (hm > 0.f) ? acc : 0.f;
  }
  __syncthreads();
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc = 0.f;
#pragma unroll 16
    for (int o = 0; o < AT_D; ++o) acc = fmaf(Ds[tt][o], W1[o][ii], acc);
    dO[(int64_t)t * AT_D + ii] = acc;
  }
}

// dx = dt + dq Wq + dk Wk + dv Wv ;  da3 = dx * [a3 > 0]  (ReLU of illum conv3, model.py:128) -> bf16
__global__ void __launch_bounds__(256) attn_dx_kernel(const float* __restrict__ dT, const float* __restrict__ dQ,
                                                      const float* __restrict__ dK, const float* __restrict__ dV,
                                                      const float* __restrict__ P, int64_t oq, int64_t ok, int64_t ov,
                                                      const bf16* __restrict__ a3, bf16* __restrict__ da3, int T) {
  SS_PDL_ENTRY();
  extern __shared__ float smf[];
  float (*Wq)[AT_D + 1] = reinterpret_cast<float (*)[AT_D + 1]>(smf);
  float (*Wk)[AT_D + 1] = Wq + AT_D;
  float (*Wv)[AT_D + 1] = Wk + AT_D;
  float (*Qs)[AT_D] = reinterpret_cast<float (*)[AT_D]>(smf + 3 * AT_D * (AT_D + 1));
  float (*Ks)[AT_D] = Qs + TK;
  float (*Vs)[AT_D] = Ks + TK;
  const int t0 = blockIdx.x * TK;
  stage_w(Wq, P + oq); stage_w(Wk, P + ok); stage_w(Wv, P + ov);
  for (int i = threadIdx.x; i < TK * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    const int64_t a = (int64_t)t * AT_D + (i & 63);
    Qs[i >> 6][i & 63] = (t < T) ? dQ[a] : 0.f;
    Ks[i >> 6][i & 63] = (t < T) ? dK[a] : 0.f;
    Vs[i >> 6][i & 63] = (t < T) ? dV[a] : 0.f;
  }
  __syncthreads();
  const int ii = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc = dT[(int64_t)t * AT_D + ii];
#pragma unroll 16
    for (int o = 0; o < AT_D; ++o) {
      acc = fmaf(Qs[tt][o], Wq[o][ii], acc);
      acc = fmaf(Ks[tt][o], Wk[o][ii], acc);
      acc = fmaf(Vs[tt][o], Wv[o][ii], acc);
    }
    da3[(int64_t)t * AT_D + ii] = f2bf(bf2f(a3[(int64_t)t * AT_D + ii]) > 0.f ? acc : 0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// attention proper: FOUR threads share one query (forward, dQ) or one key (dK/dV); each walks every 4th key / query of
// the shared-memory tile and the partial online-softmax states are merged with two xor-shuffles.
// block = 128 threads = 32 queries (keys) x 4 parts;  grid = (ceil(L/32), heads, B)
// ---------------------------------------------------------------------------------------------
#define AT_KT 64
#define AT_QB 32
// attention backward, query side: dQ_i = 0.25 * sum_j dS_ij K_j ;  also Dv_i = dO_i . O_i
__global__ void __launch_bounds__(128) attn_bwd_q_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                         const float* __restrict__ V, const float* __restrict__ O,
                                                         const float* __restrict__ dO, const float* __restrict__ LSE,
                                                         float* __restrict__ dQ, float* __restrict__ Dv, int L) {
  __shared__ float Ks[AT_KT][AT_HD + 1];
  __shared__ float Vs[AT_KT][AT_HD + 1];
  const int head = blockIdx.y, b = blockIdx.z;
  const int part = threadIdx.x & 3;
  const int qi = blockIdx.x * AT_QB + (threadIdx.x >> 2);
  const bool ok = qi < L;
  const int64_t rowbase = (int64_t)b * L;
  float q[AT_HD], go[AT_HD], dq[AT_HD];
  float Di = 0.f;
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) {
    const int64_t a = (rowbase + qi) * AT_D + head * AT_HD + d;
    q[d] = ok ? Q[a] * 0.25f : 0.f;
    go[d] = ok ? dO[a] : 0.f;
    Di += ok ? go[d] * O[a] : 0.f;
    dq[d] = 0.f;
  }
  const float lse = ok ? LSE[((int64_t)b * AT_HEADS + head) * L + qi] : 0.f;
  for (int k0 = 0; k0 < L; k0 += AT_KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_KT * AT_HD; i += 128) {
      const int kk = k0 + (i >> 4);
      const int64_t a = (rowbase + kk) * AT_D + head * AT_HD + (i & 15);
      Ks[i >> 4][i & 15] = (kk < L) ? K[a] : 0.f;
      Vs[i >> 4][i & 15] = (kk < L) ? V[a] : 0.f;
    }
    __syncthreads();
    const int kn = min(AT_KT, L - k0);
    for (int j = part; j < kn; j += 4) {
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) {
        sc = fmaf(q[d], Ks[j][d], sc);
        dp = fmaf(go[d], Vs[j][d], dp);
      }
      const float pj = __expf(sc - lse);
      const float ds = pj * (dp - Di);
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) dq[d] = fmaf(ds, Ks[j][d], dq[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) {
    dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], 1);
    dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], 2);
  }
  if (ok && part == 0) {
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) dQ[(rowbase + qi) * AT_D + head * AT_HD + d] = dq[d] * 0.25f;
    Dv[((int64_t)b * AT_HEADS + head) * L + qi] = Di;
  }
}

// attention backward, key side: dK_j = 0.25 * sum_i dS_ij Q_i ;  dV_j = sum_i P_ij dO_i
__global__ void __launch_bounds__(128) attn_bwd_kv_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                          const float* __restrict__ V, const float* __restrict__ dO,
                                                          const float* __restrict__ LSE, const float* __restrict__ Dv,
                                                          float* __restrict__ dK, float* __restrict__ dV, int L) {
  __shared__ float Qs[AT_KT][AT_HD + 1];
  __shared__ float Gs[AT_KT][AT_HD + 1];
  __shared__ float Ls[AT_KT];
  __shared__ float Ds[AT_KT];
  const int head = blockIdx.y, b = blockIdx.z;
  const int part = threadIdx.x & 3;
  const int kj = blockIdx.x * AT_QB + (threadIdx.x >> 2);
  const bool ok = kj < L;
  const int64_t rowbase = (int64_t)b * L;
  float k[AT_HD], v[AT_HD], dk[AT_HD], dv[AT_HD];
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) {
    const int64_t a = (rowbase + kj) * AT_D + head * AT_HD + d;
    k[d] = ok ? K[a] : 0.f;
    v[d] = ok ? V[a] : 0.f;
    dk[d] = 0.f;
    dv[d] = 0.f;
  }
  for (int q0 = 0; q0 < L; q0 += AT_KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < AT_KT * AT_HD; i += 128) {
      const int qq = q0 + (i >> 4);
      const int64_t a = (rowbase + qq) * AT_D + head * AT_HD + (i & 15);
      Qs[i >> 4][i & 15] = (qq < L) ? Q[a] * 0.25f : 0.f;
      Gs[i >> 4][i & 15] = (qq < L) ? dO[a] : 0.f;
    }
    if (threadIdx.x < AT_KT) {
      const int qq = q0 + threadIdx.x;
      Ls[threadIdx.x] = (qq < L) ? LSE[((int64_t)b * AT_HEADS + head) * L + qq] : 0.f;
      Ds[threadIdx.x] = (qq < L) ? Dv[((int64_t)b * AT_HEADS + head) * L + qq] : 0.f;
    }
    __syncthreads();
    const int qn = min(AT_KT, L - q0);
    for (int i = part; i < qn; i += 4) {
      float sc = 0.f, dp = 0.f;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) {
        sc = fmaf(Qs[i][d], k[d], sc);
        dp = fmaf(Gs[i][d], v[d], dp);
      }
      const float pj = __expf(sc - Ls[i]);
      const float ds = pj * (dp - Ds[i]);
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) {
        dk[d] = fmaf(ds, Qs[i][d], dk[d]);     // Qs already carries the 1/4 scale
        dv[d] = fmaf(pj, Gs[i][d], dv[d]);
      }
    }
  }
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) {
    dk[d] += __shfl_xor_sync(0xffffffffu, dk[d], 1);
    dk[d] += __shfl_xor_sync(0xffffffffu, dk[d], 2);
    dv[d] += __shfl_xor_sync(0xffffffffu, dv[d], 1);
    dv[d] += __shfl_xor_sync(0xffffffffu, dv[d], 2);
  }
  if (ok && part == 0) {
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      const int64_t a = (rowbase + kj) * AT_D + head * AT_HD + d;
      dK[a] = dk[d];
      dV[a] = dv[d];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward in TWO launches instead of three, both register-tiled (4 tokens x 4 outputs per thread, operands as float4 from
// feature-major shared-memory tiles; the older kernels above issue one shared-memory load per FMA and are LDS-bound):
//   attn_qkv4_kernel      16 tokens per CTA: x = fp32(a3), Q | K | V = x W^T + b
//   attn_core_ffn_kernel  one CTA per (image, 16 queries): per head, the K_h / V_h rows stream through shared memory in
//                         256-key tiles (one tile at the training size, 16 at 512x512 inference); two-pass softmax per
//                         tile over 16 keys per thread (16 threads per query) with a running maximum shared by the
//                         query's threads; then the FFN of the block
// A single fused kernel that recomputed K/V per CTA was measured at 67 us (issue-bound on 16 SMs); this split is ~4x
// less work per SM and uses 32 + 32 SMs.  Saves exactly what the older kernels save (x, q, k, v, o, lse, h).
// ---------------------------------------------------------------------------------------------
#define AF_QB 16
#define AF_LMAX 256
#define AF_WP (AT_D + 4)         // pitch of an input-major 64 x 64 weight tile
#define AF_W3P (3 * AT_D + 4)    // pitch of the input-major Wq | Wk | Wv tile
#define AF_TP (AF_QB + 4)        // pitch of a feature-major 16-token tile
#define AF_KP 20                 // pitch of K/V rows (16 + 4: float4 row reads by 8 lanes hit 8 distinct bank groups)

// acc[r][c] += sum_i A[i][t0 + r] * W[i][o0 + c]      (A: [64][pa] feature-major tokens, W: [64][pw] input-major)
SS_DEVINL void tile4x4(const float* __restrict__ A, int pa, int t0, const float* __restrict__ W, int pw, int o0,
                       float (&acc)[4][4]) {
#pragma unroll 8
  for (int i = 0; i < AT_D; ++i) {
    const float4 a = *reinterpret_cast<const float4*>(A + i * pa + t0);
    const float4 w = *reinterpret_cast<const float4*>(W + i * pw + o0);
    const float av[4] = {a.x, a.y, a.z, a.w}, wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], wv[c], acc[r][c]);
  }
}
// stage a (64 x 64, row = output) weight matrix transposed: Wt[i][c0 + o] = W[o][i].  All of a thread's global loads are
// issued before the first shared-memory store (a load-store loop serialises on the L2 latency: measured 5 us per matrix)
template <int NT>
SS_DEVINL void stage_wt(float* __restrict__ Wt, int pw, int c0, const float* __restrict__ W) {
  constexpr int PER = (AT_D * AT_D + NT - 1) / NT;
  float r[PER];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int idx = threadIdx.x + k * NT;
    r[k] = (idx < AT_D * AT_D) ? __ldg(W + idx) : 0.f;
  }
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const int idx = threadIdx.x + k * NT;
    if (idx < AT_D * AT_D) Wt[(idx & 63) * pw + c0 + (idx >> 6)] = r[k];
  }
}

__global__ void __launch_bounds__(192)
attn_qkv4_kernel(const bf16* __restrict__ a3, const float* __restrict__ P, int64_t oq, int64_t obq, int64_t ok,
                 int64_t obk, int64_t ov, int64_t obv, float* __restrict__ X, float* __restrict__ Q,
                 float* __restrict__ K, float* __restrict__ V, int T) {
  SS_PDL_ENTRY();
  extern __shared__ __align__(16) float smf[];
  float* Wt = smf;                               // [64][AF_W3P]
  float* Xt = Wt + AT_D * AF_W3P;                // [64][AF_TP]
  const int t0 = blockIdx.x * AF_QB, tid = threadIdx.x;
  stage_wt<192>(Wt, AF_W3P, 0, P + oq);
  stage_wt<192>(Wt, AF_W3P, AT_D, P + ok);
  stage_wt<192>(Wt, AF_W3P, 2 * AT_D, P + ov);
  for (int idx = tid; idx < AF_QB * (AT_D / 8); idx += 192) {
    const int tt = idx >> 3, c8 = (idx & 7) * 8, t = t0 + tt;
    float f[8];
    if (t < T) {
      unpack8(*reinterpret_cast<const uint4*>(a3 + (int64_t)t * AT_D + c8), f);
      float* xo = X + (int64_t)t * AT_D + c8;
      *reinterpret_cast<float4*>(xo) = make_float4(f[0], f[1], f[2], f[3]);
      *reinterpret_cast<float4*>(xo + 4) = make_float4(f[4], f[5], f[6], f[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) Xt[(c8 + i) * AF_TP + tt] = f[i];
  }
  __syncthreads();
  const int tg = tid & 3, og = tid >> 2;         // 4 token groups x 48 output groups
  float acc[4][4] = {};
  tile4x4(Xt, AF_TP, tg * 4, Wt, AF_W3P, og * 4, acc);
  const int which = og >> 4, o0 = (og & 15) * 4;
  float* dst = which == 0 ? Q : (which == 1 ? K : V);
  const int64_t ob = which == 0 ? obq : (which == 1 ? obk : obv);
  // (parameter offsets inside the flat buffer are not 16-byte aligned: scalar loads)
  const float4 bias = make_float4(P[ob + o0], P[ob + o0 + 1], P[ob + o0 + 2], P[ob + o0 + 3]);
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int t = t0 + tg * 4 + r;
    if (t < T)
      *reinterpret_cast<float4*>(dst + (int64_t)t * AT_D + o0) =
          make_float4(acc[r][0] + bias.x, acc[r][1] + bias.y, acc[r][2] + bias.z, acc[r][3] + bias.w);
  }
}

__global__ void __launch_bounds__(256)
attn_core_ffn_kernel(const float* __restrict__ X, const float* __restrict__ Q, const float* __restrict__ K,
                     const float* __restrict__ V, const float* __restrict__ P, int64_t o1, int64_t ob1, int64_t o2,
                     int64_t ob2, float* __restrict__ O, float* __restrict__ LSE, float* __restrict__ Hh,
                     bf16* __restrict__ Tout, int L) {
  SS_PDL_ENTRY();
  extern __shared__ __align__(16) float smf[];
  float* Ks = smf;                               // [256][AF_KP]   one head
  float* Vs = Ks + AF_LMAX * AF_KP;
  float* W1t = Vs + AF_LMAX * AF_KP;             // [64][AF_WP]
  float* W2t = W1t + AT_D * AF_WP;
  float* Ot = W2t + AT_D * AF_WP;                // [64][AF_TP]  attention output, feature-major
  float* Ht = Ot + AT_D * AF_TP;                 // [64][AF_TP]  relu(ff1), feature-major
  const int b = blockIdx.y, q0 = blockIdx.x * AF_QB, tid = threadIdx.x;
  const int64_t rowbase = (int64_t)b * L;
  const int qi = tid >> 4, part = tid & 15;      // 16 threads per query
  const int tq = q0 + qi;
  stage_wt<256>(W1t, AF_WP, 0, P + o1);          // consumed after the head loop
  stage_wt<256>(W2t, AF_WP, 0, P + o2);
  for (int head = 0; head < AT_HEADS; ++head) {
    float q[AT_HD];
#pragma unroll
    for (int d4 = 0; d4 < AT_HD; d4 += 4) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (tq < L) t = *reinterpret_cast<const float4*>(Q + (rowbase + tq) * AT_D + head * AT_HD + d4);
      q[d4] = 0.25f * t.x; q[d4 + 1] = 0.25f * t.y; q[d4 + 2] = 0.25f * t.z; q[d4 + 3] = 0.25f * t.w;   // model.py:110-111
    }
    // running softmax state of (query, head): the maximum is shared by the query's 16 threads (reduced per key tile), so
    // their partial sums l, o merge by plain addition at the end
    float m_run = -INFINITY, l = 0.f, o[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) o[d] = 0.f;
    for (int k0 = 0; k0 < L; k0 += AF_LMAX) {      // key tiles of 256 rows (one tile at the training size)
      __syncthreads();
      {  // K_h, V_h rows of the tile as float4: 4 + 4 loads per thread, all in flight before the first store
        float4 kv[4], vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = tid + k * 256;
          const int j = k0 + (idx >> 2), d4 = (idx & 3) * 4;
          kv[k] = vv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (j < L) {
            kv[k] = *reinterpret_cast<const float4*>(K + (rowbase + j) * AT_D + head * AT_HD + d4);
            vv[k] = *reinterpret_cast<const float4*>(V + (rowbase + j) * AT_D + head * AT_HD + d4);
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int idx = tid + k * 256;
          const int j = idx >> 2, d4 = (idx & 3) * 4;
          *reinterpret_cast<float4*>(Ks + j * AF_KP + d4) = kv[k];
          *reinterpret_cast<float4*>(Vs + j * AF_KP + d4) = vv[k];
        }
      }
      __syncthreads();
      // pass 1: the thread's 16 scores of the tile (keys part, part + 16, ...) and the tile maximum
      float sc[AF_LMAX / 16];
      float mx = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < AF_LMAX / 16; ++jj) {
        const int j = jj * 16 + part;
        float a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < AT_HD; d4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(Ks + j * AF_KP + d4);
          a = fmaf(q[d4], t.x, a); a = fmaf(q[d4 + 1], t.y, a); a = fmaf(q[d4 + 2], t.z, a); a = fmaf(q[d4 + 3], t.w, a);
        }
        sc[jj] = (k0 + j < L) ? a : -INFINITY;
        mx = fmaxf(mx, sc[jj]);
      }
#pragma unroll
      for (int sh = 1; sh <= 8; sh <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, sh));   // over the query's keys
      const float m_new = fmaxf(m_run, mx);           // finite: every tile holds at least one key
      const float corr = __expf(m_run - m_new);       // exp(-inf) = 0 on the first tile
      l *= corr;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) o[d] *= corr;
      m_run = m_new;
      // pass 2: p = exp(s - max), o += p v
#pragma unroll
      for (int jj = 0; jj < AF_LMAX / 16; ++jj) {
        const int j = jj * 16 + part;
        const float pj = __expf(sc[jj] - m_new);      // exp(-inf) = 0 for keys beyond L
        l += pj;
#pragma unroll
        for (int d4 = 0; d4 < AT_HD; d4 += 4) {
          const float4 t = *reinterpret_cast<const float4*>(Vs + j * AF_KP + d4);
          o[d4] = fmaf(pj, t.x, o[d4]); o[d4 + 1] = fmaf(pj, t.y, o[d4 + 1]);
          o[d4 + 2] = fmaf(pj, t.z, o[d4 + 2]); o[d4 + 3] = fmaf(pj, t.w, o[d4 + 3]);
        }
      }
    }
#pragma unroll
    for (int sh = 1; sh <= 8; sh <<= 1) {
      l += __shfl_xor_sync(0xffffffffu, l, sh);
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) o[d] += __shfl_xor_sync(0xffffffffu, o[d], sh);
    }
    if (part == 0) {
      const float inv = 1.f / l;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) {
        const float v = o[d] * inv;
        Ot[(head * AT_HD + d) * AF_TP + qi] = v;
        if (tq < L) O[(rowbase + tq) * AT_D + head * AT_HD + d] = v;
      }
      if (tq < L) LSE[((int64_t)b * AT_HEADS + head) * L + tq] = m_run + __logf(l);
    }
  }
  __syncthreads();
  // ---- FFN of the block's 16 tokens: 4 token groups x 16 output groups = 64 tiles
  if (tid < 64) {
    const int tg = tid & 3, og = tid >> 2;
    float acc[4][4] = {};
    tile4x4(Ot, AF_TP, tg * 4, W1t, AF_WP, og * 4, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int t = q0 + tg * 4 + r;
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        v[c] = fmaxf(acc[r][c] + P[ob1 + og * 4 + c], 0.f);
        Ht[(og * 4 + c) * AF_TP + tg * 4 + r] = v[c];
      }
      if (t < L) *reinterpret_cast<float4*>(Hh + (rowbase + t) * AT_D + og * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  __syncthreads();
  if (tid < 64) {
    const int tg = tid & 3, og = tid >> 2;
    float acc[4][4] = {};
    tile4x4(Ht, AF_TP, tg * 4, W2t, AF_WP, og * 4, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int t = q0 + tg * 4 + r;
      if (t >= L) continue;
      const float4 xr = *reinterpret_cast<const float4*>(X + (rowbase + t) * AT_D + og * 4);
      const float4 bb = make_float4(P[ob2 + og * 4], P[ob2 + og * 4 + 1], P[ob2 + og * 4 + 2], P[ob2 + og * 4 + 3]);
      uint2 pk;
      pk.x = pack2(acc[r][0] + bb.x + xr.x, acc[r][1] + bb.y + xr.y);
      pk.y = pack2(acc[r][2] + bb.z + xr.z, acc[r][3] + bb.w + xr.w);
      *reinterpret_cast<uint2*>(Tout + (rowbase + t) * AT_D + og * 4) = pk;
    }
  }
}
// ---------------------------------------------------------------------------------------------
// attention backward for the training token grid (L <= 256): same layout as attn_core_ffn_kernel - 16 threads per query
// (dQ, Dv) resp. per key (dK, dV), the head's K/V resp. Q/dO rows of the whole image staged once as float4 rows.
// grid = (ceil(L/16), heads, B), 256 threads.
// ---------------------------------------------------------------------------------------------
// stage rows [0, L) of two (B*L, 64) matrices, columns [c0, c0+16), into [256][AF_KP] tiles (scale applied to the first)
SS_DEVINL void stage_head_rows(float* __restrict__ As, float* __restrict__ Bs, const float* __restrict__ A,
                               const float* __restrict__ Bm, int64_t rowbase, int c0, int L, float scale_a) {
  float4 av[4], bv[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + k * 256;
    const int j = idx >> 2, d4 = (idx & 3) * 4;
    av[k] = bv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (j < L) {
      av[k] = *reinterpret_cast<const float4*>(A + (rowbase + j) * AT_D + c0 + d4);
      bv[k] = *reinterpret_cast<const float4*>(Bm + (rowbase + j) * AT_D + c0 + d4);
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int idx = threadIdx.x + k * 256;
    const int j = idx >> 2, d4 = (idx & 3) * 4;
    *reinterpret_cast<float4*>(As + j * AF_KP + d4) =
        make_float4(av[k].x * scale_a, av[k].y * scale_a, av[k].z * scale_a, av[k].w * scale_a);
    *reinterpret_cast<float4*>(Bs + j * AF_KP + d4) = bv[k];
  }
}
SS_DEVINL void load16(float* v, const float* __restrict__ p, bool ok, float scale) {
#pragma unroll
  for (int d4 = 0; d4 < AT_HD; d4 += 4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok) t = *reinterpret_cast<const float4*>(p + d4);
    v[d4] = t.x * scale; v[d4 + 1] = t.y * scale; v[d4 + 2] = t.z * scale; v[d4 + 3] = t.w * scale;
  }
}
SS_DEVINL float dot16(const float* a, const float* __restrict__ row) {
  float r = 0.f;
#pragma unroll
  for (int d4 = 0; d4 < AT_HD; d4 += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + d4);
    r = fmaf(a[d4], t.x, r); r = fmaf(a[d4 + 1], t.y, r); r = fmaf(a[d4 + 2], t.z, r); r = fmaf(a[d4 + 3], t.w, r);
  }
  return r;
}
SS_DEVINL void axpy16(float* acc, float s, const float* __restrict__ row) {
#pragma unroll
  for (int d4 = 0; d4 < AT_HD; d4 += 4) {
    const float4 t = *reinterpret_cast<const float4*>(row + d4);
    acc[d4] = fmaf(s, t.x, acc[d4]); acc[d4 + 1] = fmaf(s, t.y, acc[d4 + 1]);
    acc[d4 + 2] = fmaf(s, t.z, acc[d4 + 2]); acc[d4 + 3] = fmaf(s, t.w, acc[d4 + 3]);
  }
}

__global__ void __launch_bounds__(256)
attn_bwd_q16_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                    const float* __restrict__ O, const float* __restrict__ dO, const float* __restrict__ LSE,
                    float* __restrict__ dQ, float* __restrict__ Dv, int L) {
  SS_PDL_ENTRY();
  __shared__ __align__(16) float Ks[AF_LMAX * AF_KP];
  __shared__ __align__(16) float Vs[AF_LMAX * AF_KP];
  const int head = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * 16 + (threadIdx.x >> 4), part = threadIdx.x & 15;
  const bool ok = qi < L;
  const int64_t rowbase = (int64_t)b * L;
  stage_head_rows(Ks, Vs, K, V, rowbase, head * AT_HD, L, 1.f);
  float q[AT_HD], go[AT_HD], ov[AT_HD], dq[AT_HD];
  const int64_t a = (rowbase + qi) * AT_D + head * AT_HD;
  load16(q, Q + a, ok, 0.25f);
  load16(go, dO + a, ok, 1.f);
  load16(ov, O + a, ok, 1.f);
  float Di = 0.f;
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) { Di = fmaf(go[d], ov[d], Di); dq[d] = 0.f; }
  const float lse = ok ? LSE[((int64_t)b * AT_HEADS + head) * L + qi] : 0.f;
  __syncthreads();
#pragma unroll 4
  for (int j = part; j < L; j += 16) {
    const float sc = dot16(q, Ks + j * AF_KP);
    const float dp = dot16(go, Vs + j * AF_KP);
    const float ds = __expf(sc - lse) * (dp - Di);
    axpy16(dq, ds, Ks + j * AF_KP);
  }
#pragma unroll
  for (int sh = 1; sh <= 8; sh <<= 1)
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) dq[d] += __shfl_xor_sync(0xffffffffu, dq[d], sh);
  if (ok && part == 0) {
#pragma unroll
    for (int d4 = 0; d4 < AT_HD; d4 += 4)
      *reinterpret_cast<float4*>(dQ + a + d4) =
          make_float4(0.25f * dq[d4], 0.25f * dq[d4 + 1], 0.25f * dq[d4 + 2], 0.25f * dq[d4 + 3]);
    Dv[((int64_t)b * AT_HEADS + head) * L + qi] = Di;
  }
}

__global__ void __launch_bounds__(256)
attn_bwd_kv16_kernel(const float* __restrict__ Q, const float* __restrict__ K, const float* __restrict__ V,
                     const float* __restrict__ dO, const float* __restrict__ LSE, const float* __restrict__ Dv,
                     float* __restrict__ dK, float* __restrict__ dV, int L) {
  SS_PDL_ENTRY();
  __shared__ __align__(16) float Qs[AF_LMAX * AF_KP];      // 0.25 * Q_h
  __shared__ __align__(16) float Gs[AF_LMAX * AF_KP];      // dO_h
  __shared__ float Ls[AF_LMAX];
  __shared__ float Ds[AF_LMAX];
  const int head = blockIdx.y, b = blockIdx.z;
  const int kj = blockIdx.x * 16 + (threadIdx.x >> 4), part = threadIdx.x & 15;
  const bool ok = kj < L;
  const int64_t rowbase = (int64_t)b * L;
  stage_head_rows(Qs, Gs, Q, dO, rowbase, head * AT_HD, L, 0.25f);
  {
    const int i = threadIdx.x;
    Ls[i] = (i < L) ? LSE[((int64_t)b * AT_HEADS + head) * L + i] : 0.f;
    Ds[i] = (i < L) ? Dv[((int64_t)b * AT_HEADS + head) * L + i] : 0.f;
  }
  float k[AT_HD], v[AT_HD], dk[AT_HD], dv[AT_HD];
  const int64_t a = (rowbase + kj) * AT_D + head * AT_HD;
  load16(k, K + a, ok, 1.f);
  load16(v, V + a, ok, 1.f);
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) { dk[d] = 0.f; dv[d] = 0.f; }
  __syncthreads();
#pragma unroll 4
  for (int i = part; i < L; i += 16) {
    const float sc = dot16(k, Qs + i * AF_KP);
    const float dp = dot16(v, Gs + i * AF_KP);
    const float pj = __expf(sc - Ls[i]);
    const float ds = pj * (dp - Ds[i]);
    axpy16(dk, ds, Qs + i * AF_KP);          // Qs already carries the 1/4 scale
    axpy16(dv, pj, Gs + i * AF_KP);
  }
#pragma unroll
  for (int sh = 1; sh <= 8; sh <<= 1)
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      dk[d] += __shfl_xor_sync(0xffffffffu, dk[d], sh);
      dv[d] += __shfl_xor_sync(0xffffffffu, dv[d], sh);
    }
  if (ok && part == 0) {
#pragma unroll
    for (int d4 = 0; d4 < AT_HD; d4 += 4) {
      *reinterpret_cast<float4*>(dK + a + d4) = make_float4(dk[d4], dk[d4 + 1], dk[d4 + 2], dk[d4 + 3]);
      *reinterpret_cast<float4*>(dV + a + d4) = make_float4(dv[d4], dv[d4 + 1], dv[d4 + 2], dv[d4 + 3]);
    }
  }
}

// the five Linear weight gradients of the block in ONE launch: blockIdx.y = layer, blockIdx.x = chunk of 32 tokens.
//   dW[o, i] += sum_t dY[t,o] * X[t,i] ;  db[o] += sum_t dY[t,o]   (dY optionally masked by hmask > 0)
struct LinW5 {
  const float* dY[5];
  const float* hmask[5];
  const float* X[5];
  float* dW[5];
  float* db[5];
};
__global__ void __launch_bounds__(256) linear_bwd_weight5_kernel(const __grid_constant__ LinW5 a, int T, int nchunks,
                                                                 float* __restrict__ partials) {
  // block (bx, l) walks the 32-token chunks bx, bx + gridDim.x, ... of layer l and leaves its sums in row bx of
  // `partials` ([gridDim.x][5][64*64 + 64]); ss_launch_reduce_rows adds the rows in a fixed order (no atomics)
  __shared__ __align__(16) float Ys[32][AT_D + 4];
  __shared__ float Xs[32][AT_D + 1];
  const int l = blockIdx.y, tid = threadIdx.x;
  const float* __restrict__ dY = a.dY[l];
  const float* __restrict__ hm = a.hmask[l];
  const float* __restrict__ X = a.X[l];
  const int ii = tid & 63, o0 = (tid >> 6) * 16;
  float acc[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) acc[c] = 0.f;
  float sb = 0.f;
  for (int ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const int t0 = ch * 32;
    float y[8], xv[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {                 // all loads in flight before the first store
      const int i = tid + k * 256, t = t0 + (i >> 6);
      const int64_t g = (int64_t)t * AT_D + (i & 63);
      y[k] = (t < T) ? dY[g] : 0.f;
      xv[k] = (t < T) ? X[g] : 0.f;
      m[k] = (hm && t < T) ? hm[g] : 1.f;
    }
    __syncthreads();                              // the previous chunk's tiles have been consumed
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = tid + k * 256;
      Ys[i >> 6][i & 63] = (m[k] > 0.f) ? y[k] : 0.f;
      Xs[i >> 6][i & 63] = xv[k];
    }
    __syncthreads();
#pragma unroll 4
    for (int t = 0; t < 32; ++t) {
      const float xi = Xs[t][ii];
#pragma unroll
      for (int c4 = 0; c4 < 16; c4 += 4) {
        const float4 yv = *reinterpret_cast<const float4*>(&Ys[t][o0 + c4]);
        acc[c4] = fmaf(yv.x, xi, acc[c4]); acc[c4 + 1] = fmaf(yv.y, xi, acc[c4 + 1]);
        acc[c4 + 2] = fmaf(yv.z, xi, acc[c4 + 2]); acc[c4 + 3] = fmaf(yv.w, xi, acc[c4 + 3]);
      }
    }
    if (tid < AT_D) {
#pragma unroll 8
      for (int t = 0; t < 32; ++t) sb += Ys[t][tid];
    }
  }
  float* out = partials + ((size_t)blockIdx.x * 5 + l) * (AT_D * AT_D + AT_D);
#pragma unroll
  for (int c = 0; c < 16; ++c) out[(o0 + c) * AT_D + ii] = acc[c];
  if (tid < AT_D) out[AT_D * AT_D + tid] = sb;
}

static const size_t kSmemQkv4 = (AT_D * AF_W3P + AT_D * AF_TP) * sizeof(float);
static const size_t kSmemCore = (2 * AF_LMAX * AF_KP + 2 * AT_D * AF_WP + 2 * AT_D * AF_TP) * sizeof(float);

static const size_t kSmem3 = (3 * AT_D * (AT_D + 1) + 3 * TK * AT_D) * sizeof(float);   // 3 weight matrices + 3 token tiles
static const size_t kSmem2 = (2 * AT_D * (AT_D + 1) + 2 * TK * AT_D) * sizeof(float);
static int attn_attrs() {
  static DeviceOnce done;
  if (done.done()) return 0;
  cudaError_t e = cudaFuncSetAttribute(attn_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem3);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_qkv4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemQkv4);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_core_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemCore);
  if (e != cudaSuccess) {
    ss_set_error("attention: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    return SSHSLIE_ERR_CUDA;
  }
  done.set();
  return 0;
}

int ss_attn_tc_min_l() { return ss_env_int("SSHSLIE_ATTN_TC_MIN_L", SS_ATTN_TC_MIN_L); }

// parameter order inside the flat buffer: poff[0..9] = q.w q.b k.w k.b v.w v.b ff1.w ff1.b ff2.w ff2.b
int ss_attention_forward(const bf16* a3, bf16* t_out, const float* P, const int64_t* poff, AttnBuffers bf, int B, int L,
                         cudaStream_t st) {
  if (attn_attrs()) return SSHSLIE_ERR_CUDA;
  const int T = B * L;
  // register-tiled kernels; keys stream through 256-row tiles (one tile at the training size)
  ss_launch_pdl(attn_qkv4_kernel, dim3((T + AF_QB - 1) / AF_QB), dim3(192), (size_t)(kSmemQkv4), st, a3, P, poff[0], poff[1], poff[2], poff[3], poff[4],
                                                                   poff[5], bf.x, bf.q, bf.k, bf.v, T);
  static const bool tc_ok = !(getenv("SSHSLIE_ATTN_TC") && getenv("SSHSLIE_ATTN_TC")[0] == '0');
  if (tc_ok && bf.qp && bf.kvp && L >= ss_attn_tc_min_l()) {
    // large token grids (full-image inference): softmax(Q K^T) V on the tensor cores, then the FFN
    int rc = ss_check_launch("attention_qkv");
    if (!rc) rc = ss_attention_core_tc(bf.q, bf.k, bf.v, bf.qp, bf.kvp, bf.o, bf.lse, B, L, st);
    if (rc) return rc;
    ss_launch_pdl(attn_ffn_kernel, dim3((T + TK - 1) / TK), dim3(256), (size_t)(kSmem2), st, bf.o, bf.x, P, poff[6], poff[7], poff[8], poff[9],
                  bf.h, t_out, T);
    return ss_check_launch("attention_ffn");
  }
  dim3 g((L + AF_QB - 1) / AF_QB, B);
  ss_launch_pdl(attn_core_ffn_kernel, dim3(g), dim3(256), (size_t)(kSmemCore), st, bf.x, bf.q, bf.k, bf.v, P, poff[6], poff[7], poff[8], poff[9], bf.o,
                                                   bf.lse, bf.h, t_out, L);
  ss_count_launches(1);
  return ss_check_launch("attention_forward_fused");
}

int ss_attention_backward(const float* dt, const bf16* a3, bf16* da3, const float* P, float* G, const int64_t* poff,
                          AttnBuffers bf, int B, int L, cudaStream_t st) {
  (void)G;
  if (attn_attrs()) return SSHSLIE_ERR_CUDA;
  const int T = B * L;
  const int gl = (T + TK - 1) / TK;
  // data-gradient chain only; the five weight gradients run in ss_attention_backward_weights (side stream)
  ss_launch_pdl(attn_ffn_bwd_kernel, dim3(gl), dim3(256), (size_t)(kSmem2), st, dt, bf.h, P, poff[6], poff[8], bf.dh, bf.d_o, T);
  if (L <= AF_LMAX) {
    dim3 g16((L + 15) / 16, AT_HEADS, B);
    ss_launch_pdl(attn_bwd_q16_kernel, dim3(g16), dim3(256), (size_t)(0), st, bf.q, bf.k, bf.v, bf.o, bf.d_o, bf.lse, bf.dq, bf.Dv, L);
    ss_launch_pdl(attn_bwd_kv16_kernel, dim3(g16), dim3(256), (size_t)(0), st, bf.q, bf.k, bf.v, bf.d_o, bf.lse, bf.Dv, bf.dk, bf.dv, L);
  } else {
    dim3 ga((L + AT_QB - 1) / AT_QB, AT_HEADS, B);
    attn_bwd_q_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.o, bf.d_o, bf.lse, bf.dq, bf.Dv, L);
    attn_bwd_kv_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.d_o, bf.lse, bf.Dv, bf.dk, bf.dv, L);
  }
  ss_launch_pdl(attn_dx_kernel, dim3(gl), dim3(256), (size_t)(kSmem3), st, dt, bf.dq, bf.dk, bf.dv, P, poff[0], poff[2], poff[4], a3, da3, T);
  ss_count_launches(3);
  return ss_check_launch("attention_backward");
}

// dW, db of the five Linear layers (reads dt, dh, dq, dk, dv produced by ss_attention_backward)
int ss_attention_backward_weights(const float* dt, float* G, const int64_t* poff, AttnBuffers bf, int B, int L,
                                  float* scratch, cudaStream_t st) {
  if (!scratch) { ss_set_error("attention_backward_weights: scratch missing"); return SSHSLIE_ERR_WORKSPACE; }
  const int T = B * L;
  LinW5 a;
  const float* dys[5] = {dt, bf.dh, bf.dq, bf.dk, bf.dv};
  const float* hms[5] = {nullptr, bf.h, nullptr, nullptr, nullptr};
  const float* xs[5] = {bf.h, bf.o, bf.x, bf.x, bf.x};
  const int wi[5] = {8, 6, 0, 2, 4};          // ff_linear2, ff_linear1, q, k, v
  for (int i = 0; i < 5; ++i) {
    a.dY[i] = dys[i]; a.hmask[i] = hms[i]; a.X[i] = xs[i];
    a.dW[i] = G + poff[wi[i]]; a.db[i] = G + poff[wi[i] + 1];
  }
  const int nchunks = (T + 31) / 32;
  const int nb = nchunks < SS_ATTN_WGRAD_MAX_BLOCKS ? nchunks : SS_ATTN_WGRAD_MAX_BLOCKS;
  linear_bwd_weight5_kernel<<<dim3(nb, 5), 256, 0, st>>>(a, T, nchunks, scratch);
  int rc = ss_check_launch("attention_backward_weights");
  if (rc) return rc;
  RedSegs segs;
  memset(&segs, 0, sizeof(segs));
  segs.n = 10;
  for (int i = 0; i < 5; ++i) {                   // column layout of a partial row: [layer][dW (64*64) | db (64)]
    segs.dst[2 * i] = a.dW[i]; segs.len[2 * i] = AT_D * AT_D;
    segs.dst[2 * i + 1] = a.db[i]; segs.len[2 * i + 1] = AT_D;
  }
  return ss_launch_reduce_rows(scratch, nb, SS_ATTN_WGRAD_COLS, segs, st);
}
