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

// ---------------------------------------------------------------------------------------------
// Y[t, o] = sum_i X[t, i] * W[o, i] + b[o]   (+relu) (+res[t,o]);  X given as fp32 or bf16
// block = 256 threads: 64 outputs x 4 tokens per pass, 16 tokens per block
// ---------------------------------------------------------------------------------------------
template <bool XBF16>
__global__ void __launch_bounds__(256) linear_fwd_kernel(const void* __restrict__ Xv, const float* __restrict__ Wt,
                                                         const float* __restrict__ bias, float* __restrict__ Y,
                                                         const float* __restrict__ res, float* __restrict__ xcopy,
                                                         int T, int relu) {
  __shared__ float Ws[AT_D][AT_D + 1];
  __shared__ float Xs[16][AT_D];
  const int t0 = blockIdx.x * 16;
  for (int i = threadIdx.x; i < AT_D * AT_D; i += 256) Ws[i >> 6][i & 63] = Wt[i];
  for (int i = threadIdx.x; i < 16 * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    float v = 0.f;
    if (t < T) {
      v = XBF16 ? bf2f(reinterpret_cast<const bf16*>(Xv)[(int64_t)t * AT_D + (i & 63)])
                : reinterpret_cast<const float*>(Xv)[(int64_t)t * AT_D + (i & 63)];
      if (xcopy) xcopy[(int64_t)t * AT_D + (i & 63)] = v;
    }
    Xs[i >> 6][i & 63] = v;
  }
  __syncthreads();
  const int o = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < 16; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc = bias[o];
#pragma unroll 16
    for (int i = 0; i < AT_D; ++i) acc = fmaf(Xs[tt][i], Ws[o][i], acc);
    if (relu) acc = fmaxf(acc, 0.f);
    if (res) acc += res[(int64_t)t * AT_D + o];
    Y[(int64_t)t * AT_D + o] = acc;
  }
}

// dX[t, i] (+)= sum_o dY[t, o] * W[o, i]     (optionally dY is first masked by Hmask > 0: ReLU backward)
__global__ void __launch_bounds__(256) linear_bwd_data_kernel(const float* __restrict__ dY, const float* __restrict__ Wt,
                                                              const float* __restrict__ hmask, float* __restrict__ dX,
                                                              int T, int accumulate) {
  __shared__ float Ws[AT_D][AT_D + 1];
  __shared__ float Ys[16][AT_D];
  const int t0 = blockIdx.x * 16;
  for (int i = threadIdx.x; i < AT_D * AT_D; i += 256) Ws[i >> 6][i & 63] = Wt[i];
  for (int i = threadIdx.x; i < 16 * AT_D; i += 256) {
    const int t = t0 + (i >> 6);
    float v = 0.f;
    if (t < T) {
      v = dY[(int64_t)t * AT_D + (i & 63)];
      if (hmask && !(hmask[(int64_t)t * AT_D + (i & 63)] > 0.f)) v = 0.f;
    }
    Ys[i >> 6][i & 63] = v;
  }
  __syncthreads();
  const int ii = threadIdx.x & 63;
  for (int tt = threadIdx.x >> 6; tt < 16; tt += 4) {
    const int t = t0 + tt;
    if (t >= T) break;
    float acc = 0.f;
#pragma unroll 16
    for (int o = 0; o < AT_D; ++o) acc = fmaf(Ys[tt][o], Ws[o][ii], acc);
    if (accumulate) acc += dX[(int64_t)t * AT_D + ii];
    dX[(int64_t)t * AT_D + ii] = acc;
  }
}

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
// attention forward: one thread = one query of one (b, head); K/V tiles of 64 keys staged in shared memory
// ---------------------------------------------------------------------------------------------
#define AT_KT 64
__global__ void __launch_bounds__(128) attn_fwd_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                       const float* __restrict__ V, float* __restrict__ O,
                                                       float* __restrict__ LSE, int L) {
  __shared__ float Ks[AT_KT][AT_HD];
  __shared__ float Vs[AT_KT][AT_HD];
  const int head = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * 128 + threadIdx.x;
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
    float sc[AT_KT];
    float tmax = mx;
#pragma unroll
    for (int j = 0; j < AT_KT; ++j) {
      float a = 0.f;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) a = fmaf(q[d], Ks[j][d], a);
      sc[j] = (j < kn) ? a : -INFINITY;
      tmax = fmaxf(tmax, sc[j]);
    }
    const float corr = __expf(mx - tmax);
    l *= corr;
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) o[d] *= corr;
#pragma unroll
    for (int j = 0; j < AT_KT; ++j) {
      const float pj = __expf(sc[j] - tmax);
      l += pj;
#pragma unroll
      for (int d = 0; d < AT_HD; ++d) o[d] = fmaf(pj, Vs[j][d], o[d]);
    }
    mx = tmax;
  }
  if (ok) {
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
  __shared__ float Ks[AT_KT][AT_HD];
  __shared__ float Vs[AT_KT][AT_HD];
  const int head = blockIdx.y, b = blockIdx.z;
  const int qi = blockIdx.x * 128 + threadIdx.x;
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
    for (int j = 0; j < kn; ++j) {
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
  if (ok) {
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
  __shared__ float Qs[AT_KT][AT_HD];
  __shared__ float Gs[AT_KT][AT_HD];
  __shared__ float Ls[AT_KT];
  __shared__ float Ds[AT_KT];
  const int head = blockIdx.y, b = blockIdx.z;
  const int kj = blockIdx.x * 128 + threadIdx.x;
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
    for (int i = 0; i < qn; ++i) {
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
  if (ok) {
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      const int64_t a = (rowbase + kj) * AT_D + head * AT_HD + d;
      dK[a] = dk[d];
      dV[a] = dv[d];
    }
  }
}

// t = x + y  -> bf16 token matrix (the NHWC activation consumed by deconv1's upsample)
__global__ void add_to_bf16_kernel(const float* __restrict__ a, bf16* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f2bf(a[i]);
}
// da3 = (dx) * (a3 > 0)   (ReLU of illum conv3, model.py:128) -> bf16
__global__ void mask_to_bf16_kernel(const float* __restrict__ dx, const bf16* __restrict__ a3, bf16* __restrict__ out,
                                    int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = f2bf(bf2f(a3[i]) > 0.f ? dx[i] : 0.f);
}

// parameter order inside the flat buffer: poff[0..9] = q.w q.b k.w k.b v.w v.b ff1.w ff1.b ff2.w ff2.b
int ss_attention_forward(const bf16* a3, bf16* t_out, const float* P, const int64_t* poff, AttnBuffers bf, int B, int L,
                         cudaStream_t st) {
  const int T = B * L;
  const int gl = (T + 15) / 16;
  linear_fwd_kernel<true><<<gl, 256, 0, st>>>(a3, P + poff[0], P + poff[1], bf.q, nullptr, bf.x, T, 0);
  linear_fwd_kernel<false><<<gl, 256, 0, st>>>(bf.x, P + poff[2], P + poff[3], bf.k, nullptr, nullptr, T, 0);
  linear_fwd_kernel<false><<<gl, 256, 0, st>>>(bf.x, P + poff[4], P + poff[5], bf.v, nullptr, nullptr, T, 0);
  dim3 ga((L + 127) / 128, AT_HEADS, B);
  attn_fwd_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.o, bf.lse, L);
  linear_fwd_kernel<false><<<gl, 256, 0, st>>>(bf.o, P + poff[6], P + poff[7], bf.h, nullptr, nullptr, T, 1);
  linear_fwd_kernel<false><<<gl, 256, 0, st>>>(bf.h, P + poff[8], P + poff[9], bf.t32, bf.x, nullptr, T, 0);
  const int64_t n = (int64_t)T * AT_D;
  add_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bf.t32, t_out, n);
  ss_count_launches(6);
  return ss_check_launch("attention_forward");
}

int ss_attention_backward(const float* dt, const bf16* a3, bf16* da3, const float* P, float* G, const int64_t* poff,
                          AttnBuffers bf, int B, int L, cudaStream_t st) {
  const int T = B * L;
  const int gl = (T + 15) / 16;
  // data-gradient chain only; the five weight gradients run in ss_attention_backward_weights (side stream)
  linear_bwd_data_kernel<<<gl, 256, 0, st>>>(dt, P + poff[8], nullptr, bf.dh, T, 0);          // dh (pre-mask)
  linear_bwd_data_kernel<<<gl, 256, 0, st>>>(bf.dh, P + poff[6], bf.h, bf.d_o, T, 0);         // dO (h = relu: mask)
  dim3 ga((L + 127) / 128, AT_HEADS, B);
  attn_bwd_q_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.o, bf.d_o, bf.lse, bf.dq, bf.Dv, L);
  attn_bwd_kv_kernel<<<ga, 128, 0, st>>>(bf.q, bf.k, bf.v, bf.d_o, bf.lse, bf.Dv, bf.dk, bf.dv, L);
  // dx = dt (residual) + dq Wq + dk Wk + dv Wv
  const int64_t n = (int64_t)T * AT_D;
  cudaMemcpyAsync(bf.dx, dt, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  linear_bwd_data_kernel<<<gl, 256, 0, st>>>(bf.dq, P + poff[0], nullptr, bf.dx, T, 1);
  linear_bwd_data_kernel<<<gl, 256, 0, st>>>(bf.dk, P + poff[2], nullptr, bf.dx, T, 1);
  linear_bwd_data_kernel<<<gl, 256, 0, st>>>(bf.dv, P + poff[4], nullptr, bf.dx, T, 1);
  mask_to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(bf.dx, a3, da3, n);
  ss_count_launches(8);
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
