// TransformerBlock (model.py:87-119) forward and backward on the H/8 x W/8 token grid, fp32.
//   tokens: T = B*L rows of 64 features (the bf16 NHWC activation a3 IS the (B, L, 64) token matrix)
//   Q,K,V = Linear(64,64) ; 4 heads x 16 ; softmax(QK^T / 4) V ; y = x + W2 relu(W1 o + b1) + b2
// Streaming (flash-style) softmax: the (B,4,L,L) logits of model.py:111-113 are never materialised.
// FLOPs are negligible at train size (L = 256); at 512^2 inference L = 4096 and QK^T/PV is ~1% of the step.
#include "common.cuh"
#include "kernels.h"

#define AT_D 64
#define AT_HEADS 4
#define AT_HD 16

// dW[o, i] += sum_t dY[t,o] * X[t,i] ;  db[o] += sum_t dY[t,o]   (dY optionally masked by hmask > 0)
// grid.x = token chunks of 64; block 256 = 64 (i) x 4 (o phase); atomics into the flat gradient buffer
__global__ void __launch_bounds__(256) linear_bwd_weight_kernel(const float* __restrict__ dY,
                                                                const float* __restrict__ hmask,
                                                                const float* __restrict__ X, float* __restrict__ dW,
                                                                float* __restrict__ db, int T) {
  __shared__ float Ys[64][AT_D + 1];
  __shared__ float Xs[64][AT_D + 1];
  const int t0 = blockIdx.x * 64;
  for (int i = threadIdx.x; i < 64 * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    float y = 0.f, xv = 0.f;
    if (t < T) {
      y = dY[(int64_t)t * AT_D + (i & 63)];
      if (hmask && !(hmask[(int64_t)t * AT_D + (i & 63)] > 0.f)) y = 0.f;
      xv = X[(int64_t)t * AT_D + (i & 63)];
    }
    Ys[i >> 6][i & 63] = y;
    Xs[i >> 6][i & 63] = xv;
  }
  __syncthreads();
  const int ii = threadIdx.x & 63;
  for (int o = threadIdx.x >> 6; o < AT_D; o += 4) {
    float acc = 0.f;
#pragma unroll 16
    for (int t = 0; t < 64; ++t) acc = fmaf(Ys[t][o], Xs[t][ii], acc);
    atomicAdd(dW + o * AT_D + ii, acc);
  }
  if (threadIdx.x < AT_D) {
    float acc = 0.f;
    for (int t = 0; t < 64; ++t) acc += Ys[t][threadIdx.x];
    atomicAdd(db + threadIdx.x, acc);
  }
}

// ---------------------------------------------------------------------------------------------
// fused token kernels (16 tokens per block, 256 threads = 64 features x 4 token phases)
// ---------------------------------------------------------------------------------------------
#define TK 16
SS_DEVINL void stage_w(float (*Ws)[AT_D + 1], const float* __restrict__ Wt) {
  for (int i = threadIdx.x; i < AT_D * AT_D; i += 256) Ws[i >> 6][i & 63] = Wt[i];
}

// x = fp32(a3);  q,k,v = x Wq^T + bq, ...                                   (model.py:103-106)
__global__ void __launch_bounds__(256) attn_qkv_kernel(const bf16* __restrict__ a3, const float* __restrict__ P,
                                                       const int64_t* __restrict__ poff_unused, int64_t oq, int64_t obq,
                                                       int64_t ok, int64_t obk, int64_t ov, int64_t obv,
                                                       float* __restrict__ X, float* __restrict__ Q,
                                                       float* __restrict__ K, float* __restrict__ V, int T) {
  extern __shared__ float smf[];
  float (*Wq)[AT_D + 1] = reinterpret_cast<float (*)[AT_D + 1]>(smf);
  float (*Wk)[AT_D + 1] = Wq + AT_D;
  float (*Wv)[AT_D + 1] = Wk + AT_D;
  float (*Xs)[AT_D] = reinterpret_cast<float (*)[AT_D]>(smf + 3 * AT_D * (AT_D + 1));
  const int t0 = blockIdx.x * TK;
  stage_w(Wq, P + oq); stage_w(Wk, P + ok); stage_w(Wv, P + ov);
  for (int i = threadIdx.x; i < TK * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    float v = 0.f;
    if (t < T) {
      v = bf2f(a3[(int64_t)t * AT_D + (i & 63)]);
      X[(int64_t)t * AT_D + (i & 63)] = v;
    }
    Xs[i >> 6][i & 63] = v;
  }
  __syncthreads();
  const int o = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < TK; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float aq = P[obq + o], ak = P[obk + o], av = P[obv + o];
#pragma unroll 16
    for (int i = 0; i < AT_D; ++i) {
      const float xv = Xs[tt][i];
      aq = fmaf(xv, Wq[o][i], aq);
      ak = fmaf(xv, Wk[o][i], ak);
      av = fmaf(xv, Wv[o][i], av);
    }
    Q[(int64_t)t * AT_D + o] = aq;
    K[(int64_t)t * AT_D + o] = ak;
    V[(int64_t)t * AT_D + o] = av;
  }
}

// h = relu(o W1^T + b1) ; t = x + h W2^T + b2  -> h (kept for backward), t as bf16     (model.py:115-118)
__global__ void __launch_bounds__(256) attn_ffn_kernel(const float* __restrict__ O, const float* __restrict__ X,
                                                       const float* __restrict__ P, int64_t o1, int64_t ob1, int64_t o2,
                                                       int64_t ob2, float* __restrict__ Hh, bf16* __restrict__ Tout,
                                                       int T) {
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
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                       const float* __restrict__ V, float* __restrict__ O,
                                                       float* __restrict__ LSE, int L) {
  __shared__ float Ks[AT_KT][AT_HD + 1];
  __shared__ float Vs[AT_KT][AT_HD + 1];
  const int head = blockIdx.y, b = blockIdx.z;
  const int part = threadIdx.x & 3;
  const int qi = blockIdx.x * AT_QB + (threadIdx.x >> 2);
  const bool ok = qi < L;
  const int64_t rowbase = (int64_t)b * L;
  float q[AT_HD], o[AT_HD];
#pragma unroll
  for (int d = 0; d < AT_HD; ++d) {
    q[d] = ok ? Q[(rowbase + qi) * AT_D + head * AT_HD + d] * 0.25f : 0.f;   // 1/sqrt(16), model.py:110-111
    o[d] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
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
    float sc[AT_KT / 4];
    float tmax = mx;
#pragma unroll
    for (int jj = 0; jj < AT_KT / 4; ++jj) {
      const int j = jj * 4 + part;
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) a = fmaf(q[d], Ks[j][d], a);
      sc[jj] = (j < kn) ? a : -INFINITY;
      tmax = fmaxf(tmax, sc[jj]);
    }
    if (tmax > -INFINITY) {
      const float corr = __expf(mx - tmax);
      l *= corr;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) o[d] *= corr;
#pragma unroll
      for (int jj = 0; jj < AT_KT / 4; ++jj) {
        const int j = jj * 4 + part;
        const float pj = __expf(sc[jj] - tmax);
        l += pj;
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) o[d] = fmaf(pj, Vs[j][d], o[d]);
      }
      mx = tmax;
    }
  }
  // merge the four partial states of the query
#pragma unroll
  for (int sh = 1; sh <= 2; sh <<= 1) {
    const float mo = __shfl_xor_sync(0xffffffffu, mx, sh);
    const float lo = __shfl_xor_sync(0xffffffffu, l, sh);
    const float mn = fmaxf(mx, mo);
    const float ca = (mx > -INFINITY) ? __expf(mx - mn) : 0.f, cb = (mo > -INFINITY) ? __expf(mo - mn) : 0.f;
    l = l * ca + lo * cb;
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      const float oo = __shfl_xor_sync(0xffffffffu, o[d], sh);
      o[d] = o[d] * ca + oo * cb;
    }
    mx = mn;
  }
  if (ok && part == 0) {
    const float inv = 1.f / l;
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) O[(rowbase + qi) * AT_D + head * AT_HD + d] = o[d] * inv;
    if (LSE) LSE[((int64_t)b * AT_HEADS + head) * L + qi] = mx + __logf(l);
  }
}

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

static const size_t kSmem3 = (3 * AT_D * (AT_D + 1) + 3 * TK * AT_D) * sizeof(float);   // 3 weight matrices + 3 token tiles
static const size_t kSmem2 = (2 * AT_D * (AT_D + 1) + 2 * TK * AT_D) * sizeof(float);
static int attn_attrs() {
  static bool done = false;
  if (done) return 0;
  cudaError_t e = cudaFuncSetAttribute(attn_qkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem3);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_dx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem3);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_ffn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(attn_ffn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
  if (e != cudaSuccess) {
    ss_set_error("attention: cannot raise dynamic shared memory: %s", cudaGetErrorString(e));
    return SSHSLIE_ERR_CUDA;
  }
  done = true;
  return 0;
}

// parameter order inside the flat buffer: poff[0..9] = q.w q.b k.w k.b v.w v.b ff1.w ff1.b ff2.w ff2.b
int ss_attention_forward(const bf16* a3, bf16* t_out, const float* P, const int64_t* poff, AttnBuffers bf, int B, int L,
                         cudaStream_t st) {
  if (attn_attrs()) return SSHSLIE_ERR_CUDA;
  const int T = B * L;
  const int gl = (T + TK - 1) / TK;
  attn_qkv_kernel<<<gl, 256, kSmem3, st>>>(a3, P, nullptr, poff[0], poff[1], poff[2], poff[3], poff[4], poff[5], bf.x, bf.q,
                                            bf.k, bf.v, T);
  dim3 ga((L + AT_QB - 1) / AT_QB, AT_HEADS, B);
  attn_fwd_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.o, bf.lse, L);
  attn_ffn_kernel<<<gl, 256, kSmem2, st>>>(bf.o, bf.x, P, poff[6], poff[7], poff[8], poff[9], bf.h, t_out, T);
  ss_count_launches(2);
  return ss_check_launch("attention_forward");
}

int ss_attention_backward(const float* dt, const bf16* a3, bf16* da3, const float* P, float* G, const int64_t* poff,
                          AttnBuffers bf, int B, int L, cudaStream_t st) {
  (void)G;
  if (attn_attrs()) return SSHSLIE_ERR_CUDA;
  const int T = B * L;
  const int gl = (T + TK - 1) / TK;
  // data-gradient chain only; the five weight gradients run in ss_attention_backward_weights (side stream)
  attn_ffn_bwd_kernel<<<gl, 256, kSmem2, st>>>(dt, bf.h, P, poff[6], poff[8], bf.dh, bf.d_o, T);
  dim3 ga((L + AT_QB - 1) / AT_QB, AT_HEADS, B);
  attn_bwd_q_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.o, bf.d_o, bf.lse, bf.dq, bf.Dv, L);
  attn_bwd_kv_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.d_o, bf.lse, bf.Dv, bf.dk, bf.dv, L);
  attn_dx_kernel<<<gl, 256, kSmem3, st>>>(dt, bf.dq, bf.dk, bf.dv, P, poff[0], poff[2], poff[4], a3, da3, T);
  ss_count_launches(3);
  return ss_check_launch("attention_backward");
}

// dW, db of the five Linear layers (reads dt, dh, dq, dk, dv produced by ss_attention_backward)
int ss_attention_backward_weights(const float* dt, float* G, const int64_t* poff, AttnBuffers bf, int B, int L,
                                  cudaStream_t st) {
  const int T = B * L;
  const int gw = (T + 63) / 64;
  linear_bwd_weight_kernel<<<gw, 256, 0, st>>>(dt, nullptr, bf.h, G + poff[8], G + poff[9], T);       // ff_linear2
  linear_bwd_weight_kernel<<<gw, 256, 0, st>>>(bf.dh, bf.h, bf.o, G + poff[6], G + poff[7], T);       // ff_linear1
  linear_bwd_weight_kernel<<<gw, 256, 0, st>>>(bf.dq, nullptr, bf.x, G + poff[0], G + poff[1], T);
  linear_bwd_weight_kernel<<<gw, 256, 0, st>>>(bf.dk, nullptr, bf.x, G + poff[2], G + poff[3], T);
  linear_bwd_weight_kernel<<<gw, 256, 0, st>>>(bf.dv, nullptr, bf.x, G + poff[4], G + poff[5], T);
  ss_count_launches(4);
  return ss_check_launch("attention_backward_weights");
}
