// fourier_spectrum_loss (model.py:456-473), forward value and d/dS in ONE kernel, one CTA per (b, band) image.
//
//   loss_sum += sum_k mask_k * | |X_k| - |S_k| |          X = fft2(x), S = fft2(s), unnormalised, full spectrum,
//                                                        mask on the UN-shifted grid
//   dS      += grad_scale * Re( IDFT_unnorm( -mask_k * sgn(|X_k| - |S_k|) * S_k / |S_k| ) )
//
// Both real images ride ONE complex transform: z = x + i*s, so X_k = (Z_k + conj(Z_-k))/2 and
// S_k = (Z_k - conj(Z_-k))/(2i).  The HxW complex tile lives in shared memory (128 KB at 128x128);
// forward = radix-2 decimation-in-frequency along W then H (natural in, bit-reversed out), the spectrum is
// consumed in bit-reversed positions, and the inverse runs decimation-in-time (bit-reversed in, natural out),
// so no reordering pass exists.  Algorithmic HBM traffic: read x, s (2 planes), read+write dS (1 plane each).
#include "common.cuh"
#include "kernels.h"

#define FFT_THREADS 512

SS_DEVINL float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
SS_DEVINL int brev(int v, int bits) { return (int)(__brev((unsigned)v) >> (32 - bits)); }

// one radix-2 stage over the whole tile along the W axis (rows) or the H axis (columns)
//   dif:  a' = a + b ; b' = (a - b) * tw        dit:  b *= tw ; a' = a + b ; b' = a - b
template <bool ALONG_W, bool DIT, bool INV>
SS_DEVINL void fft_stage(float2* z, const float2* tw, int H, int W, int half, int n_axis) {
  const int nbf = H * W / 2;
  const int tstep = n_axis / (2 * half);
  for (int t = threadIdx.x; t < nbf; t += FFT_THREADS) {
    int i0, i1, pos;
    if (ALONG_W) {
      const int row = t / (W / 2), j = t - row * (W / 2);
      const int grp = j / half;
      pos = j - grp * half;
      i0 = row * W + grp * 2 * half + pos;
      i1 = i0 + half;
    } else {
      const int col = t % W, j = t / W;          // consecutive lanes -> consecutive columns (conflict-free)
      const int grp = j / half;
      pos = j - grp * half;
      i0 = (grp * 2 * half + pos) * W + col;
      i1 = i0 + half * W;
    }
    float2 w = tw[pos * tstep];
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

__global__ void __launch_bounds__(FFT_THREADS, 1)
fourier_loss_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ mask,
                    float* __restrict__ dS, float* __restrict__ sum_out, int H, int W, int lgH, int lgW,
                    float grad_scale) {
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
  for (int i = threadIdx.x; i < H * W; i += FFT_THREADS) z[i] = make_float2(xp[i], sp[i]);
  __syncthreads();

  // forward: DIF along W, then along H
  for (int half = W / 2; half >= 1; half >>= 1) fft_stage<true, false, false>(z, twW, H, W, half, W);
  for (int half = H / 2; half >= 1; half >>= 1) fft_stage<false, false, false>(z, twH, H, W, half, H);

  // spectrum pass: position (ph,pw) holds frequency (brev(ph), brev(pw)); pair it with -k
  float lsum = 0.f;
  for (int p = threadIdx.x; p < H * W; p += FFT_THREADS) {
    const int ph = p / W, pw = p - ph * W;
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
    if (threadIdx.x == 0) atomicAdd(sum_out, t);
  }
  if (dS == nullptr) return;

  // inverse: DIT along H then W with conjugated twiddles (bit-reversed in, natural out)
  for (int half = 1; half <= H / 2; half <<= 1) fft_stage<false, true, true>(z, twH, H, W, half, H);
  for (int half = 1; half <= W / 2; half <<= 1) fft_stage<true, true, true>(z, twW, H, W, half, W);

  float* dp = dS + img * H * W;
  for (int i = threadIdx.x; i < H * W; i += FFT_THREADS) dp[i] += grad_scale * z[i].x;
}

static int ilog2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

extern "C" int sshslie_fourier_loss(const float* x, const float* S, const float* mask, float* dS, float* sum_out,
                                    int n_img, int H, int W, float grad_scale, void* stream) {
  const int lgH = ilog2_exact(H), lgW = ilog2_exact(W);
  if (!x || !S || !mask || !sum_out || n_img < 1 || lgH < 3 || lgW < 3 || H > 128 || W > 128) {
    ss_set_error("sshslie_fourier_loss: H and W must be powers of two in [8,128] (got %dx%d)", H, W);
    return SSHSLIE_ERR_ARG;
  }
  const size_t smem = (size_t)H * W * sizeof(float2) + 128 * sizeof(float2);
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(fourier_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
        cudaSuccess) {
      ss_set_error("sshslie_fourier_loss: cannot raise dynamic shared memory: %s",
                   cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_set = true;
  }
  fourier_loss_kernel<<<n_img, FFT_THREADS, smem, (cudaStream_t)stream>>>(x, S, mask, dS, sum_out, H, W, lgH, lgW,
                                                                         grad_scale);
  return ss_check_launch("fourier_loss");
}
