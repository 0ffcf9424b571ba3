// Fused forward + gradient of the five pixel-space loss terms of LowLightEnhance.compute_loss
// (model.py:551-555 with helpers 445-454, 475-542), fp32, NCHW planes, one pass, no atomics anywhere: every block writes its
// nine term sums to a row of `partials`, which one warp per term adds up in a fixed order (deterministic, main.py:165).
//
//   0 L_reconstruction   mean |R*I - x|                                                  model.py:551
//   1,2 L_I_smooth_low   mean(wx |dx I|) + mean(wy |dy I|),  w = exp(-a1 * mean_c |d R|)   model.py:505-515
//   3,4,5 L_R_fidelity   mean|R-Re| + 0.5 (mean|dx(R-Re)| + mean|dy(R-Re)|)               model.py:521-534
//   6,7 L_I_smooth_delta mean(|dx Id| exp(-a2 |dx R_c|)) + same in y, Id broadcast over c  model.py:450-454
//   8 L_spectral_cons    mean |S[c+1] - S[c]|,  S = R*(Id + I)                            model.py:475-481, 233
//
// Eight threads own one pixel (b,h,w), each walking one eighth of the band axis twice: pass A builds the band-mean edge
// weights of term 1/2 for the four incident edges (combined through shared memory), pass B accumulates the term sums and
// writes every gradient by GATHER
// (each pixel collects the contributions of the <=4 forward-difference edges it takes part in).
// Gradients are written already multiplied by c_loss_x / count (d total_loss / d tensor).
#include <stdlib.h>
#include "common.cuh"
#include "kernels.h"

struct PixLossArgs {
  const float *x, *R, *I, *Id, *Re;
  float *partials, *dR, *dI, *dId, *dS, *dRe;      // partials[block][9]
  int B, C, H, W;
  float a1, a2;
  float k_rec, k_ilx, k_ily, k_rf, k_rfx, k_rfy, k_idx, k_idy, k_sp;   // c_loss / count per term
};

#define PL_CHUNKS 8      // the band axis is split over 8 threads per pixel: 8x the parallelism of one-thread-per-pixel
#define PL_WX 128        // pixels of one image row handled by a block

SS_DEVINL float block_sum_2d(float v, float* red, int tid, int nthreads) {
  v = warp_sum(v);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  if (tid < 32) {
    v = (tid < ((nthreads + 31) >> 5)) ? red[tid] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

// blockDim = (WX, 8): threadIdx.x = pixel along the row (coalesced plane reads), threadIdx.y = band chunk.
// grid = (ceil(W / WX), H, B)
__global__ void __launch_bounds__(PL_WX* PL_CHUNKS) pixel_losses_kernel(PixLossArgs p) {
  SS_PDL_ENTRY();
  __shared__ float part[PL_CHUNKS][4][PL_WX];
  __shared__ float red[32];
  const int W = p.W, H = p.H, C = p.C;
  const int HW = H * W;
  const int tx = threadIdx.x, k = threadIdx.y;
  const int tid = k * blockDim.x + tx, nthreads = blockDim.x * blockDim.y;
  const int w = blockIdx.x * blockDim.x + tx, h = blockIdx.y, b = blockIdx.z;
  const bool active = w < W;
  const int cpc = (C + PL_CHUNKS - 1) / PL_CHUNKS;
  const int c_begin = k * cpc, c_end = min(C, c_begin + cpc);
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = 0.f;

  const int hw = h * W + (active ? w : 0);
  const int64_t pix = (int64_t)b * HW + hw;
  const bool hasR = w + 1 < W, hasL = w > 0, hasD = h + 1 < H, hasU = h > 0;
  const float* Rb = p.R + (int64_t)b * C * HW + hw;
  const float* Eb = p.Re + (int64_t)b * C * HW + hw;
  const float* xb = p.x + (int64_t)b * C * HW + hw;
  const float* Ib = p.I + (int64_t)b * HW + hw;
  const float* Db = p.Id + (int64_t)b * HW + hw;

  // ---- pass A: band means of |dR| on the four incident edges (partial over this thread's chunk) -------------
  float mR = 0.f, mL = 0.f, mD = 0.f, mU = 0.f;
  if (active) {
    for (int c = c_begin; c < c_end; ++c) {
      const float* r = Rb + (int64_t)c * HW;
      const float r0 = r[0];
      if (hasR) mR += fabsf(r[1] - r0);
      if (hasL) mL += fabsf(r0 - r[-1]);
      if (hasD) mD += fabsf(r[W] - r0);
      if (hasU) mU += fabsf(r0 - r[-W]);
    }
  }
  part[k][0][tx] = mR; part[k][1][tx] = mL; part[k][2][tx] = mD; part[k][3][tx] = mU;
  __syncthreads();
  mR = mL = mD = mU = 0.f;
#pragma unroll
  for (int i = 0; i < PL_CHUNKS; ++i) {
    mR += part[i][0][tx]; mL += part[i][1][tx]; mD += part[i][2][tx]; mU += part[i][3][tx];
  }
  __syncthreads();

  float gI = 0.f, gId = 0.f;
  if (active) {
    const float invC = 1.f / (float)C;
    const float wR = __expf(-p.a1 * mR * invC), wL = __expf(-p.a1 * mL * invC);
    const float wD = __expf(-p.a1 * mD * invC), wU = __expf(-p.a1 * mU * invC);
    const float i0 = Ib[0], d0 = Db[0];
    const float iR = hasR ? Ib[1] - i0 : 0.f, iL = hasL ? i0 - Ib[-1] : 0.f;
    const float iD = hasD ? Ib[W] - i0 : 0.f, iU = hasU ? i0 - Ib[-W] : 0.f;
    const float dRt = hasR ? Db[1] - d0 : 0.f, dLt = hasL ? d0 - Db[-1] : 0.f;
    const float dDt = hasD ? Db[W] - d0 : 0.f, dUt = hasU ? d0 - Db[-W] : 0.f;
    if (k == 0) {   // band-independent parts, once per pixel
      if (hasR) s[1] += wR * fabsf(iR);
      if (hasD) s[2] += wD * fabsf(iD);
      gI = p.k_ilx * (wL * sgnf(iL) - wR * sgnf(iR)) + p.k_ily * (wU * sgnf(iU) - wD * sgnf(iD));
    }
    const float cR = p.k_ilx * wR * fabsf(iR) * p.a1 * invC, cL = p.k_ilx * wL * fabsf(iL) * p.a1 * invC;
    const float cD = p.k_ily * wD * fabsf(iD) * p.a1 * invC, cU = p.k_ily * wU * fabsf(iU) * p.a1 * invC;
    const float gain = d0 + i0;
    float s_prev = (c_begin > 0) ? Rb[(int64_t)(c_begin - 1) * HW] * gain : 0.f;
    float r_next = (c_begin < C) ? Rb[(int64_t)c_begin * HW] : 0.f;
    // ---- pass B over this thread's bands --------------------------------------------------------------------
    for (int c = c_begin; c < c_end; ++c) {
      const int64_t off = (int64_t)c * HW;
      const float* r = Rb + off;
      const float* e = Eb + off;
      const float r0 = r_next;
      if (c + 1 < C) r_next = r[HW];
      const float e0 = e[0];
      float gR = 0.f;
      const float u = r0 * i0 - xb[off];
      s[0] += fabsf(u);
      const float gu = p.k_rec * sgnf(u);
      gR += gu * i0;
      gI += gu * r0;
      const float q0 = r0 - e0;
      s[3] += fabsf(q0);
      float gq = p.k_rf * sgnf(q0);
      if (hasR) {
        const float dr = r[1] - r0;
        const float dq = (r[1] - e[1]) - q0;
        const float ex = __expf(-p.a2 * fabsf(dr));
        s[4] += fabsf(dq);
        s[6] += fabsf(dRt) * ex;
        gq -= p.k_rfx * sgnf(dq);
        gR += (cR + p.k_idx * fabsf(dRt) * p.a2 * ex) * sgnf(dr);
        gId -= p.k_idx * sgnf(dRt) * ex;
      }
      if (hasL) {
        const float dr = r0 - r[-1];
        const float dq = q0 - (r[-1] - e[-1]);
        const float ex = __expf(-p.a2 * fabsf(dr));
        gq += p.k_rfx * sgnf(dq);
        gR -= (cL + p.k_idx * fabsf(dLt) * p.a2 * ex) * sgnf(dr);
        gId += p.k_idx * sgnf(dLt) * ex;
      }
      if (hasD) {
        const float dr = r[W] - r0;
        const float dq = (r[W] - e[W]) - q0;
        const float ex = __expf(-p.a2 * fabsf(dr));
        s[5] += fabsf(dq);
        s[7] += fabsf(dDt) * ex;
        gq -= p.k_rfy * sgnf(dq);
        gR += (cD + p.k_idy * fabsf(dDt) * p.a2 * ex) * sgnf(dr);
        gId -= p.k_idy * sgnf(dDt) * ex;
      }
      if (hasU) {
        const float dr = r0 - r[-W];
        const float dq = q0 - (r[-W] - e[-W]);
        const float ex = __expf(-p.a2 * fabsf(dr));
        gq += p.k_rfy * sgnf(dq);
        gR -= (cU + p.k_idy * fabsf(dUt) * p.a2 * ex) * sgnf(dr);
        gId += p.k_idy * sgnf(dUt) * ex;
      }
      gR += gq;
      // spectral smoothness on S = R*(Id+I):  dS_c = k (sgn(S_c - S_{c-1}) - sgn(S_{c+1} - S_c))
      const float s0 = r0 * gain;
      float gS = 0.f;
      if (c > 0) gS += sgnf(s0 - s_prev);
      if (c + 1 < C) {
        const float sn = r_next * gain;
        s[8] += fabsf(sn - s0);
        gS -= sgnf(sn - s0);
      }
      s_prev = s0;
      const int64_t o = (int64_t)b * C * HW + hw + off;
      if (p.dR) p.dR[o] = gR;
      if (p.dRe) p.dRe[o] = -gq;
      if (p.dS) p.dS[o] = p.k_sp * gS;
    }
  }
  part[k][0][tx] = gI; part[k][1][tx] = gId;
  __syncthreads();
  if (active && k == 0) {
    float a = 0.f, d = 0.f;
#pragma unroll
    for (int i = 0; i < PL_CHUNKS; ++i) { a += part[i][0][tx]; d += part[i][1][tx]; }
    if (p.dI) p.dI[pix] = a;
    if (p.dId) p.dId[pix] = d;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float t = block_sum_2d(s[i], red, tid, nthreads);
    if (tid == 0) p.partials[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 9 + i] = t;
  }
}

// Tiled variant (C = 64, W % 16 == 0, H % 4 == 0): one block = a 4 x 16 pixel tile x 8 band chunks (512 threads).  The
// R and R_enh values of the tile and its one-pixel halo, ALL 64 bands, are staged in shared memory once (every global load
// of the thread issued before the first store); both band sweeps then read their 5-point stencils from shared memory
// instead of issuing 16 global loads per element.  grid = (W / 16, H / 4, B)
#define PT_TW 16
#define PT_TH 4
#define PT_PITCH (PT_TW + 2)
#define PT_PLANE ((PT_TH + 2) * PT_PITCH)
#define PT_PIX (PT_TW * PT_TH)
#define PT_C 64
__global__ void __launch_bounds__(PT_PIX* PL_CHUNKS) pixel_losses_tiled_kernel(PixLossArgs p) {
  SS_PDL_ENTRY();
  extern __shared__ __align__(16) float pt_smem[];
  float* Rs = pt_smem;                       // [64][PT_TH + 2][PT_PITCH]
  float* Es = pt_smem + PT_C * PT_PLANE;
  __shared__ float part[PL_CHUNKS][4][PT_PIX];
  const int W = p.W, H = p.H, C = p.C;
  const int HW = H * W;
  const int tx = threadIdx.x, k = threadIdx.y;
  const int tid = k * blockDim.x + tx, nthreads = blockDim.x * blockDim.y;
  const int w0 = blockIdx.x * PT_TW, h0 = blockIdx.y * PT_TH, b = blockIdx.z;
  const int lx = tx & (PT_TW - 1), ly = tx >> 4;
  const int w = w0 + lx, h = h0 + ly;
  const bool active = true;
  {  // stage R and R_enh: (band, row, col) of the haloed tile, zero outside the image
    constexpr int NEL = PT_C * PT_PLANE;
    constexpr int PER = (NEL + PT_PIX * PL_CHUNKS - 1) / (PT_PIX * PL_CHUNKS);
    const float* Rg = p.R + (int64_t)b * C * HW;
    const float* Eg = p.Re + (int64_t)b * C * HW;
    float rv[PER], ev[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tid + q * (PT_PIX * PL_CHUNKS);
      rv[q] = ev[q] = 0.f;
      if (i < NEL) {
        const int c = i / PT_PLANE, rem = i - c * PT_PLANE;
        const int gy = h0 - 1 + rem / PT_PITCH, gx = w0 - 1 + rem % PT_PITCH;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const int64_t o = (int64_t)c * HW + gy * W + gx;
          rv[q] = __ldg(Rg + o);
          ev[q] = __ldg(Eg + o);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) {
      const int i = tid + q * (PT_PIX * PL_CHUNKS);
      if (i < NEL) { Rs[i] = rv[q]; Es[i] = ev[q]; }
    }
  }
  __syncthreads();
  const int sidx = (ly + 1) * PT_PITCH + (lx + 1);      // this pixel inside a band plane of the staged tile
  const int cpc = (C + PL_CHUNKS - 1) / PL_CHUNKS;
  const int c_begin = k * cpc, c_end = min(C, c_begin + cpc);
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = 0.f;

  const int hw = h * W + (active ? w : 0);
  const int64_t pix = (int64_t)b * HW + hw;
  const bool hasR = w + 1 < W, hasL = w > 0, hasD = h + 1 < H, hasU = h > 0;
  const float* xb = p.x + (int64_t)b * C * HW + hw;
  const float* Ib = p.I + (int64_t)b * HW + hw;
  const float* Db = p.Id + (int64_t)b * HW + hw;

  // ---- pass A: band means of |dR| on the four incident edges (partial over this thread's chunk) -------------
  float mR = 0.f, mL = 0.f, mD = 0.f, mU = 0.f;
  if (active) {
    for (int c = c_begin; c < c_end; ++c) {
      const float* r = Rs + c * PT_PLANE + sidx;
      const float r0 = r[0];
      if (hasR) mR += fabsf(r[1] - r0);
      if (hasL) mL += fabsf(r0 - r[-1]);
      if (hasD) mD += fabsf(r[PT_PITCH] - r0);
      if (hasU) mU += fabsf(r0 - r[-PT_PITCH]);
    }
  }
  part[k][0][tx] = mR; part[k][1][tx] = mL; part[k][2][tx] = mD; part[k][3][tx] = mU;
  __syncthreads();
  mR = mL = mD = mU = 0.f;
#pragma unroll
  for (int i = 0; i < PL_CHUNKS; ++i) {
    mR += part[i][0][tx]; mL += part[i][1][tx]; mD += part[i][2][tx]; mU += part[i][3][tx];
  }
  __syncthreads();

  float gI = 0.f, gId = 0.f;
  if (active) {
    const float invC = 1.f / (float)C;
    const float wR = __expf(-p.a1 * mR * invC), wL = __expf(-p.a1 * mL * invC);
    const float wD = __expf(-p.a1 * mD * invC), wU = __expf(-p.a1 * mU * invC);
    const float i0 = Ib[0], d0 = Db[0];
    const float iR = hasR ? Ib[1] - i0 : 0.f, iL = hasL ? i0 - Ib[-1] : 0.f;
    const float iD = hasD ? Ib[W] - i0 : 0.f, iU = hasU ? i0 - Ib[-W] : 0.f;
    const float dRt = hasR ? Db[1] - d0 : 0.f, dLt = hasL ? d0 - Db[-1] : 0.f;
    const float dDt = hasD ? Db[W] - d0 : 0.f, dUt = hasU ? d0 - Db[-W] : 0.f;
    if (k == 0) {   // band-independent parts, once per pixel
      if (hasR) s[1] += wR * fabsf(iR);
      if (hasD) s[2] += wD * fabsf(iD);
      gI = p.k_ilx * (wL * sgnf(iL) - wR * sgnf(iR)) + p.k_ily * (wU * sgnf(iU) - wD * sgnf(iD));
    }
    const float cR = p.k_ilx * wR * fabsf(iR) * p.a1 * invC, cL = p.k_ilx * wL * fabsf(iL) * p.a1 * invC;
    const float cD = p.k_ily * wD * fabsf(iD) * p.a1 * invC, cU = p.k_ily * wU * fabsf(iU) * p.a1 * invC;
    const float gain = d0 + i0;
    float s_prev = (c_begin > 0) ? Rs[(c_begin - 1) * PT_PLANE + sidx] * gain : 0.f;
    float r_next = (c_begin < C) ? Rs[c_begin * PT_PLANE + sidx] : 0.f;
    // ---- pass B over this thread's bands --------------------------------------------------------------------
    for (int c = c_begin; c < c_end; ++c) {
      const int64_t off = (int64_t)c * HW;
      const float* r = Rs + c * PT_PLANE + sidx;
      const float* e = Es + c * PT_PLANE + sidx;
      const float r0 = r_next;
      if (c + 1 < C) r_next = r[PT_PLANE];
      const float e0 = e[0];
      float gR = 0.f;
      const float u = r0 * i0 - xb[off];
      s[0] += fabsf(u);
      const float gu = p.k_rec * sgnf(u);
      gR += gu * i0;
      gI += gu * r0;
      const float q0 = r0 - e0;
      s[3] += fabsf(q0);
      float gq = p.k_rf * sgnf(q0);
      if (hasR) {
        const float dr = r[1] - r0;
        const float dq = (r[1] - e[1]) - q0;
        const float ex = __expf(-p.a2 * fabsf(dr));
        s[4] += fabsf(dq);
        s[6] += fabsf(dRt) * ex;
        gq -= p.k_rfx * sgnf(dq);
        gR += (cR + p.k_idx * fabsf(dRt) * p.a2 * ex) * sgnf(dr);
        gId -= p.k_idx * sgnf(dRt) * ex;
      }
      if (hasL) {
        const float dr = r0 - r[-1];
        const float dq = q0 - (r[-1] - e[-1]);
        const float ex = __expf(-p.a2 * fabsf(dr));
        gq += p.k_rfx * sgnf(dq);
        gR -= (cL + p.k_idx * fabsf(dLt) * p.a2 * ex) * sgnf(dr);
        gId += p.k_idx * sgnf(dLt) * ex;
      }
      if (hasD) {
        const float dr = r[PT_PITCH] - r0;
        const float dq = (r[PT_PITCH] - e[PT_PITCH]) - q0;
        const float ex = __expf(-p.a2 * fabsf(dr));
        s[5] += fabsf(dq);
        s[7] += fabsf(dDt) * ex;
        gq -= p.k_rfy * sgnf(dq);
        gR += (cD + p.k_idy * fabsf(dDt) * p.a2 * ex) * sgnf(dr);
        gId -= p.k_idy * sgnf(dDt) * ex;
      }
      if (hasU) {
        const float dr = r0 - r[-PT_PITCH];
        const float dq = q0 - (r[-PT_PITCH] - e[-PT_PITCH]);
        const float ex = __expf(-p.a2 * fabsf(dr));
        gq += p.k_rfy * sgnf(dq);
        gR -= (cU + p.k_idy * fabsf(dUt) * p.a2 * ex) * sgnf(dr);
        gId += p.k_idy * sgnf(dUt) * ex;
      }
      gR += gq;
      // spectral smoothness on S = R*(Id+I):  dS_c = k (sgn(S_c - S_{c-1}) - sgn(S_{c+1} - S_c))
      const float s0 = r0 * gain;
      float gS = 0.f;
      if (c > 0) gS += sgnf(s0 - s_prev);
      if (c + 1 < C) {
        const float sn = r_next * gain;
        s[8] += fabsf(sn - s0);
        gS -= sgnf(sn - s0);
      }
      s_prev = s0;
      const int64_t o = (int64_t)b * C * HW + hw + off;
      if (p.dR) p.dR[o] = gR;
      if (p.dRe) p.dRe[o] = -gq;
      if (p.dS) p.dS[o] = p.k_sp * gS;
    }
  }
  part[k][0][tx] = gI; part[k][1][tx] = gId;
  __syncthreads();
  if (active && k == 0) {
    float a = 0.f, d = 0.f;
#pragma unroll
    for (int i = 0; i < PL_CHUNKS; ++i) { a += part[i][0][tx]; d += part[i][1][tx]; }
    if (p.dI) p.dI[pix] = a;
    if (p.dId) p.dId[pix] = d;
  }
  // the nine term sums of the block: warp sums -> one shared-memory exchange -> this block's row of `partials`
  __shared__ float wred[PT_PIX * PL_CHUNKS / 32][9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float t = warp_sum(s[i]);
    if ((tid & 31) == 0) wred[tid >> 5][i] = t;
  }
  __syncthreads();
  if (tid < 9) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < PT_PIX * PL_CHUNKS / 32; ++wv) t += wred[wv][tid];
    p.partials[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 9 + tid] = t;
  }
  (void)nthreads;
}

// ---------------------------------------------------------------------------------------------
// Row-streaming variant (W % 128 == 0): no shared-memory staging.  A warp owns 128 consecutive pixels of one image row
// (one float4 per lane: every global access is a full 512-byte row segment) and one eighth of the band axis; horizontal
// neighbours come from the adjacent lanes by shuffle, the rows above and below from global memory (the second row of the
// block / the L1 hold them).  Image borders are handled by CLAMPING the neighbour to the pixel itself: the difference is then
// exactly 0, sgn(0) = 0, and every term of the absent edge vanishes without a branch.
// The band-mean edge weights of L_I_smooth_low need all 64 bands before the main sweep can start, so they are produced by a
// small kernel of their own (edge_weights_rows_kernel: one more read of R, 2 maps of B*H*W floats out).
//   grid = (W / 128, ceil(H / RW_TH), B), block = (32, 8 band chunks, RW_TH rows)
// ---------------------------------------------------------------------------------------------
#define RW_TH 1
// k * sgn(v) with sgn(0) = 0 (torch.sign): the sign bit of v is XORed into k
SS_DEVINL float ksgn(float k, float v) {
  const float t = __int_as_float(__float_as_int(k) ^ (__float_as_int(v) & (int)0x80000000));
  return (v != 0.f) ? t : 0.f;
}
SS_DEVINL float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
SS_DEVINL void f4_to(const float4& v, float* a) { a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w; }

__global__ void __launch_bounds__(32 * PL_CHUNKS) edge_weights_rows_kernel(PixLossArgs p, float* __restrict__ wx,
                                                                          float* __restrict__ wy) {
  SS_PDL_ENTRY();
  __shared__ float4 part[PL_CHUNKS][2][32];
  const int W = p.W, H = p.H, C = p.C, HW = H * W;
  const int lane = threadIdx.x, k = threadIdx.y;
  const int w0 = blockIdx.x * 128 + lane * 4, h = blockIdx.y, b = blockIdx.z;
  const int cpc = (C + PL_CHUNKS - 1) / PL_CHUNKS;
  const int c_begin = k * cpc, c_end = min(C, c_begin + cpc);
  const int dn = (h + 1 < H) ? W : 0;                       // clamped: the last row is its own lower neighbour
  const bool edge_r = (lane == 31), has_next = (w0 + 4 < W);
  const float* base = p.R + ((int64_t)b * C * H + h) * W + w0;
  float mx[4] = {0.f, 0.f, 0.f, 0.f}, my[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int c = c_begin; c < c_end; ++c) {
    const float* q = base + (int64_t)c * HW;
    const float4 r0 = ldg4(q), rd = ldg4(q + dn);
    float rn = __shfl_down_sync(0xffffffffu, r0.x, 1);
    if (edge_r) rn = has_next ? __ldg(q + 4) : r0.w;
    mx[0] += fabsf(r0.y - r0.x); mx[1] += fabsf(r0.z - r0.y); mx[2] += fabsf(r0.w - r0.z); mx[3] += fabsf(rn - r0.w);
    my[0] += fabsf(rd.x - r0.x); my[1] += fabsf(rd.y - r0.y); my[2] += fabsf(rd.z - r0.z); my[3] += fabsf(rd.w - r0.w);
  }
  part[k][0][lane] = make_float4(mx[0], mx[1], mx[2], mx[3]);
  part[k][1][lane] = make_float4(my[0], my[1], my[2], my[3]);
  __syncthreads();
  if (k >= 2) return;
  float4 t = part[0][k][lane];
#pragma unroll
  for (int i = 1; i < PL_CHUNKS; ++i) {
    const float4 u = part[i][k][lane];
    t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
  }
  const float sc = -p.a1 / (float)C;
  t.x = __expf(sc * t.x); t.y = __expf(sc * t.y); t.z = __expf(sc * t.z); t.w = __expf(sc * t.w);
  float* o = (k == 0 ? wx : wy) + ((int64_t)b * H + h) * W + w0;
  *reinterpret_cast<float4*>(o) = t;
}

__global__ void __launch_bounds__(32 * PL_CHUNKS * RW_TH, 2 / RW_TH)
pixel_losses_rows_kernel(PixLossArgs p, const float* __restrict__ wx, const float* __restrict__ wy) {
  SS_PDL_ENTRY();
  __shared__ float4 gpart[RW_TH][PL_CHUNKS][2][32];
  __shared__ float wred[PL_CHUNKS * RW_TH][9];
  const int W = p.W, H = p.H, C = p.C, HW = H * W;
  const int lane = threadIdx.x, k = threadIdx.y, rz = threadIdx.z;
  const int w0 = blockIdx.x * 128 + lane * 4, b = blockIdx.z;
  const int h_raw = blockIdx.y * RW_TH + rz;
  const bool row_ok = h_raw < H;
  const int h = row_ok ? h_raw : H - 1;                     // (an odd H leaves the block's second row idle: it recomputes the
                                                            //  last row and its results are dropped)
  const int cpc = (C + PL_CHUNKS - 1) / PL_CHUNKS;
  const int c_begin = k * cpc, c_end = min(C, c_begin + cpc);
  const int up = (h > 0) ? -W : 0, dn = (h + 1 < H) ? W : 0;
  const bool first = (lane == 0), last = (lane == 31);
  const bool has_prev = (w0 > 0), has_next = (w0 + 4 < W);
  const int64_t pix = ((int64_t)b * H + h) * W + w0;

  // ---- band-independent quantities of the thread's four pixels ----
  // horizontal edges e = 0..4 of the thread (edge e joins pixels e-1 and e; 0 and 4 reach into the neighbouring lanes):
  //   ch = d L_I_smooth_low / d|dR_c|, Kh = k_idx |d I_delta| a2, gh = k_idx sgn(d I_delta), ah = |d I_delta|
  float i0[4], gain[4], ch[5], Kh[5], gh[5], ah[5], cU[4], cD[4], tU[4], tD[4];   // t* = forward differences of I_delta
  float gI[4] = {0.f, 0.f, 0.f, 0.f}, gId[4] = {0.f, 0.f, 0.f, 0.f};
  float s[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) s[i] = 0.f;
  {
    float ii[6], dd[6], iu[4], id_[4], du[4], dd_[4], wh[5], wd[4], wu[4];
    f4_to(ldg4(p.I + pix), ii + 1); f4_to(ldg4(p.Id + pix), dd + 1);
    f4_to(ldg4(p.I + pix + up), iu); f4_to(ldg4(p.I + pix + dn), id_);
    f4_to(ldg4(p.Id + pix + up), du); f4_to(ldg4(p.Id + pix + dn), dd_);
    f4_to(ldg4(wx + pix), wh + 1); f4_to(ldg4(wy + pix), wd); f4_to(ldg4(wy + pix + up), wu);
    ii[0] = __shfl_up_sync(0xffffffffu, ii[4], 1); ii[5] = __shfl_down_sync(0xffffffffu, ii[1], 1);
    dd[0] = __shfl_up_sync(0xffffffffu, dd[4], 1); dd[5] = __shfl_down_sync(0xffffffffu, dd[1], 1);
    wh[0] = __shfl_up_sync(0xffffffffu, wh[4], 1);
    if (first) {
      ii[0] = has_prev ? __ldg(p.I + pix - 1) : ii[1];
      dd[0] = has_prev ? __ldg(p.Id + pix - 1) : dd[1];
      wh[0] = has_prev ? __ldg(wx + pix - 1) : 1.f;
    }
    if (last) {
      ii[5] = has_next ? __ldg(p.I + pix + 4) : ii[4];
      dd[5] = has_next ? __ldg(p.Id + pix + 4) : dd[4];
    }
    const float invC = 1.f / (float)C;
    float sgi[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const float ih = ii[e + 1] - ii[e], th = dd[e + 1] - dd[e];
      ch[e] = p.k_ilx * wh[e] * fabsf(ih) * p.a1 * invC;
      Kh[e] = p.k_idx * fabsf(th) * p.a2;
      gh[e] = ksgn(p.k_idx, th);
      ah[e] = fabsf(th);
      sgi[e] = wh[e] * sgnf(ih);
      if (k == 0 && e > 0) s[1] += wh[e] * fabsf(ih);          // edge e is the right edge of pixel e-1
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = ii[j + 1], d0 = dd[j + 1];
      const float iD = id_[j] - v0, iU = v0 - iu[j];
      const float wD = wd[j], wU = wu[j];
      i0[j] = v0; gain[j] = d0 + v0;
      tD[j] = dd_[j] - d0; tU[j] = d0 - du[j];
      cD[j] = p.k_ily * wD * fabsf(iD) * p.a1 * invC; cU[j] = p.k_ily * wU * fabsf(iU) * p.a1 * invC;
      if (k == 0) {   // once per pixel
        s[2] += wD * fabsf(iD);
        gI[j] = p.k_ilx * (sgi[j] - sgi[j + 1]) + p.k_ily * (wU * sgnf(iU) - wD * sgnf(iD));
      }
    }
  }

  // ---- one sweep over this thread's bands ----
  const float* Rb = p.R + (int64_t)b * C * HW + (int64_t)h * W + w0;
  const float* Eb = p.Re + (int64_t)b * C * HW + (int64_t)h * W + w0;
  const float* Xb = p.x + (int64_t)b * C * HW + (int64_t)h * W + w0;
  float sprev[4] = {0.f, 0.f, 0.f, 0.f};
  if (c_begin > 0 && c_begin < C) {
    float t[4];
    f4_to(ldg4(Rb + (int64_t)(c_begin - 1) * HW), t);
#pragma unroll
    for (int j = 0; j < 4; ++j) sprev[j] = t[j] * gain[j];
  }
  float4 r_next = (c_begin < C) ? ldg4(Rb + (int64_t)c_begin * HW) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int64_t off = (int64_t)c * HW;
    float rr[6], qq[6], ru[4], rd[4], eu[4], ed[4], xx[4], rn[4];
    const float4 r4 = r_next;
    const float4 e4 = ldg4(Eb + off);
    const float4 ru4 = ldg4(Rb + off + up), rd4 = ldg4(Rb + off + dn);
    const float4 eu4 = ldg4(Eb + off + up), ed4 = ldg4(Eb + off + dn);
    const float4 x4 = ldg4(Xb + off);
    if (c + 1 < C) r_next = ldg4(Rb + off + HW);
    f4_to(r4, rr + 1); f4_to(ru4, ru); f4_to(rd4, rd); f4_to(eu4, eu); f4_to(ed4, ed); f4_to(x4, xx); f4_to(r_next, rn);
    qq[1] = r4.x - e4.x; qq[2] = r4.y - e4.y; qq[3] = r4.z - e4.z; qq[4] = r4.w - e4.w;
    rr[0] = __shfl_up_sync(0xffffffffu, rr[4], 1); rr[5] = __shfl_down_sync(0xffffffffu, rr[1], 1);
    qq[0] = __shfl_up_sync(0xffffffffu, qq[4], 1); qq[5] = __shfl_down_sync(0xffffffffu, qq[1], 1);
    if (first) {
      if (has_prev) { rr[0] = __ldg(Rb + off - 1); qq[0] = rr[0] - __ldg(Eb + off - 1); }
      else { rr[0] = rr[1]; qq[0] = qq[1]; }
    }
    if (last) {
      if (has_next) { rr[5] = __ldg(Rb + off + 4); qq[5] = rr[5] - __ldg(Eb + off + 4); }
      else { rr[5] = rr[4]; qq[5] = qq[4]; }
    }
    // horizontal edges e = 0..4 between rr[e] and rr[e+1]: evaluated once, applied with opposite signs to both end pixels
    float Eh[5], Gh[5], Dh[5];
#pragma unroll
    for (int e = 0; e < 5; ++e) {
      const float dr = rr[e + 1] - rr[e], dq = qq[e + 1] - qq[e];
      const float ex = __expf(-p.a2 * fabsf(dr));
      Eh[e] = ksgn(ch[e] + Kh[e] * ex, dr);
      Gh[e] = ksgn(p.k_rfx, dq);
      Dh[e] = gh[e] * ex;
      if (e > 0) {
        s[4] += fabsf(dq);
        s[6] += ah[e] * ex;
      }
    }
    float oR[4], oE[4], oS[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r0 = rr[j + 1], q0 = qq[j + 1];
      const float u = r0 * i0[j] - xx[j];
      s[0] += fabsf(u);
      const float gu = ksgn(p.k_rec, u);
      float gR = gu * i0[j] + (Eh[j + 1] - Eh[j]);
      gI[j] += gu * r0;
      s[3] += fabsf(q0);
      float gq = ksgn(p.k_rf, q0) + (Gh[j] - Gh[j + 1]);
      gId[j] += Dh[j] - Dh[j + 1];
      {  // lower edge
        const float dr = rd[j] - r0;
        const float dq = (rd[j] - ed[j]) - q0;
        const float ex = __expf(-p.a2 * fabsf(dr));
        s[5] += fabsf(dq);
        s[7] += fabsf(tD[j]) * ex;
        gq -= ksgn(p.k_rfy, dq);
        gR += ksgn(cD[j] + p.k_idy * fabsf(tD[j]) * p.a2 * ex, dr);
        gId[j] -= ksgn(p.k_idy, tD[j]) * ex;
      }
      {  // upper edge
        const float dr = r0 - ru[j];
        const float dq = q0 - (ru[j] - eu[j]);
        const float ex = __expf(-p.a2 * fabsf(dr));
        gq += ksgn(p.k_rfy, dq);
        gR -= ksgn(cU[j] + p.k_idy * fabsf(tU[j]) * p.a2 * ex, dr);
        gId[j] += ksgn(p.k_idy, tU[j]) * ex;
      }
      gR += gq;
      // spectral smoothness on S = R*(Id+I):  dS_c = k (sgn(S_c - S_{c-1}) - sgn(S_{c+1} - S_c))
      const float s0 = r0 * gain[j];
      float gS = 0.f;
      if (c > 0) gS = ksgn(p.k_sp, s0 - sprev[j]);
      if (c + 1 < C) {
        const float sn = rn[j] * gain[j];
        s[8] += fabsf(sn - s0);
        gS -= ksgn(p.k_sp, sn - s0);
      }
      sprev[j] = s0;
      oR[j] = gR; oE[j] = -gq; oS[j] = gS;
    }
    if (row_ok) {
      const int64_t o = (int64_t)b * C * HW + (int64_t)h * W + w0 + off;
      if (p.dR) *reinterpret_cast<float4*>(p.dR + o) = make_float4(oR[0], oR[1], oR[2], oR[3]);
      if (p.dRe) *reinterpret_cast<float4*>(p.dRe + o) = make_float4(oE[0], oE[1], oE[2], oE[3]);
      if (p.dS) *reinterpret_cast<float4*>(p.dS + o) = make_float4(oS[0], oS[1], oS[2], oS[3]);
    }
  }
  // ---- dI, dI_delta: sum over the eight band chunks in a fixed order ----
  gpart[rz][k][0][lane] = make_float4(gI[0], gI[1], gI[2], gI[3]);
  gpart[rz][k][1][lane] = make_float4(gId[0], gId[1], gId[2], gId[3]);
  if (!row_ok) {
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const float t = warp_sum(s[i]);
    if (lane == 0) wred[rz * PL_CHUNKS + k][i] = t;
  }
  __syncthreads();
  if (k < 2 && row_ok) {
    float4 t = gpart[rz][0][k][lane];
#pragma unroll
    for (int i = 1; i < PL_CHUNKS; ++i) {
      const float4 u = gpart[rz][i][k][lane];
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    float* o = (k == 0) ? p.dI : p.dId;
    if (o) *reinterpret_cast<float4*>(o + pix) = t;
  }
  const int tid = (rz * PL_CHUNKS + k) * 32 + lane;
  if (tid < 9) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < PL_CHUNKS * RW_TH; ++wv) t += wred[wv][tid];
    p.partials[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 9 + tid] = t;
  }
}

static bool pixel_rows(int C, int H, int W) {
  static const bool rows_ok = !(getenv("SSHSLIE_LOSS_ROWS") && getenv("SSHSLIE_LOSS_ROWS")[0] == '0');
  (void)H;
  return rows_ok && C >= 2 && (W % 128) == 0;
}
static bool pixel_tiled(int C, int H, int W) {
  static const bool tiled_ok = !(getenv("SSHSLIE_LOSS_TILED") && getenv("SSHSLIE_LOSS_TILED")[0] == '0');
  return tiled_ok && C == PT_C && (W % PT_TW) == 0 && (H % PT_TH) == 0;
}
// rows of 9 partial sums the kernel writes for this shape (= its grid size)
int ss_pixel_losses_blocks(int B, int C, int H, int W) {
  if (pixel_rows(C, H, W)) return (W / 128) * ((H + RW_TH - 1) / RW_TH) * B;
  if (pixel_tiled(C, H, W)) return (W / PT_TW) * (H / PT_TH) * B;
  const int wx = W < PL_WX ? W : PL_WX;
  return ((W + wx - 1) / wx) * H * B;
}

// floats of scratch ss_pixel_losses needs: the rows of partial sums, then (row-streaming kernel) the two edge-weight maps
int64_t ss_pixel_losses_scratch_floats(int B, int C, int H, int W) {
  int64_t n = (int64_t)ss_pixel_losses_blocks(B, C, H, W) * 9;
  n = (n + 3) & ~(int64_t)3;                                 // the maps are accessed as float4
  if (pixel_rows(C, H, W)) n += (int64_t)2 * B * H * W;
  return n;
}

// out[i] (+)= sum over rows of partials[row][i]: one warp per column, lanes stride the rows, fixed shuffle tree
__global__ void __launch_bounds__(32) reduce_partials_kernel(const float* __restrict__ partials, int nrows, int ncols,
                                                             float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x;
  float t = 0.f;
  for (int r = threadIdx.x; r < nrows; r += 32) t += partials[(size_t)r * ncols + c];
  t = warp_sum(t);
  if (threadIdx.x == 0) out[c] = accumulate ? out[c] + t : t;
}
int ss_reduce_partials(const float* partials, int nrows, int ncols, float* out, int accumulate, cudaStream_t st) {
  reduce_partials_kernel<<<ncols, 32, 0, st>>>(partials, nrows, ncols, out, accumulate);
  return ss_check_launch("reduce_partials");
}

int ss_pixel_losses(const float* x, const float* R, const float* I, const float* Id, const float* Re,
                    const sshslie_loss_cfg& cfg, int B, int C, int H, int W, float* partials, float* dR, float* dI,
                    float* dId, float* dS, float* dRe, cudaStream_t st) {
  PixLossArgs p;
  p.x = x; p.R = R; p.I = I; p.Id = Id; p.Re = Re;
  p.partials = partials; p.dR = dR; p.dI = dI; p.dId = dId; p.dS = dS; p.dRe = dRe;
  p.B = B; p.C = C; p.H = H; p.W = W;
  p.a1 = cfg.alpha_i_smooth_low;
  p.a2 = cfg.alpha_i_smooth_delta;
  const double n0 = (double)B * C * H * W;
  const double nx1 = (double)B * H * (W - 1), ny1 = (double)B * (H - 1) * W;
  p.k_rec = (float)(cfg.c_loss_reconstruction / n0);
  p.k_ilx = (float)(cfg.c_loss_i_smooth_low / nx1);
  p.k_ily = (float)(cfg.c_loss_i_smooth_low / ny1);
  p.k_rf = (float)(cfg.c_loss_r_fidelity / n0);
  p.k_rfx = (float)(cfg.c_loss_r_fidelity * 0.5 / (nx1 * C));
  p.k_rfy = (float)(cfg.c_loss_r_fidelity * 0.5 / (ny1 * C));
  p.k_idx = (float)(cfg.c_loss_i_smooth_delta / (nx1 * C));
  p.k_idy = (float)(cfg.c_loss_i_smooth_delta / (ny1 * C));
  p.k_sp = (float)(cfg.c_loss_spectral_cons / ((double)B * (C - 1) * H * W));
  if (pixel_rows(C, H, W)) {
    const int64_t rows9 = ((int64_t)ss_pixel_losses_blocks(B, C, H, W) * 9 + 3) & ~(int64_t)3;
    float* wx = partials + rows9;
    float* wy = wx + (int64_t)B * H * W;
    ss_launch_pdl(edge_weights_rows_kernel, dim3(W / 128, H, B), dim3(32, PL_CHUNKS), (size_t)0, st, p, wx, wy);
    int rc = ss_check_launch("edge_weights_rows");
    if (rc) return rc;
    ss_launch_pdl(pixel_losses_rows_kernel, dim3(W / 128, (H + RW_TH - 1) / RW_TH, B), dim3(32, PL_CHUNKS, RW_TH), (size_t)0,
                  st, p, (const float*)wx, (const float*)wy);
    return ss_check_launch("pixel_losses_rows");
  }
  if (pixel_tiled(C, H, W)) {
    const size_t smem = (size_t)2 * PT_C * PT_PLANE * sizeof(float);
    static DeviceOnce attr_once;
    if (!attr_once.done()) {
      if (cudaFuncSetAttribute(pixel_losses_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
          cudaSuccess) {
        ss_set_error("pixel_losses: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
        return SSHSLIE_ERR_CUDA;
      }
      attr_once.set();
    }
    dim3 grid(W / PT_TW, H / PT_TH, B), block(PT_PIX, PL_CHUNKS);
    ss_launch_pdl(pixel_losses_tiled_kernel, dim3(grid), dim3(block), (size_t)(smem), st, p);
    return ss_check_launch("pixel_losses_tiled");
  }
  const int wx = W < PL_WX ? W : PL_WX;
  dim3 grid((W + wx - 1) / wx, H, B), block(wx, PL_CHUNKS);
  ss_launch_pdl(pixel_losses_kernel, dim3(grid), dim3(block), (size_t)(0), st, p);
  return ss_check_launch("pixel_losses");
}

extern "C" int64_t sshslie_loss_scratch_bytes(int B, int C, int H, int W) {
  const int64_t pix = ss_pixel_losses_scratch_floats(B, C, H, W);
  // the Fourier kernels write one partial per (b, band) plane and, for sizes outside the shared-memory FFT, transform in a
  // global workspace behind them
  const int64_t four = ((int64_t)B * C + 3) / 4 * 4 + ss_fourier_work_floats(B * C, H, W);
  return (pix > four ? pix : four) * (int64_t)sizeof(float);
}

extern "C" int sshslie_pixel_losses(const float* x, const float* R, const float* I, const float* Idelta,
                                    const float* S, const float* R_enh, const sshslie_loss_cfg* cfg, int B, int C,
                                    int H, int W, float* sums, float* dR, float* dI, float* dIdelta, float* dS,
                                    float* dR_enh, void* scratch, int64_t scratch_bytes, void* stream) {
  (void)S;  // S = R*(Idelta + I) is recomputed in-kernel (model.py:233), the argument documents the dependency
  if (!x || !R || !I || !Idelta || !R_enh || !cfg || !sums || !scratch || B < 1 || C < 2 || H < 2 || W < 2) {
    ss_set_error("sshslie_pixel_losses: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  if (scratch_bytes < sshslie_loss_scratch_bytes(B, C, H, W)) {
    ss_set_error("sshslie_pixel_losses: scratch too small (need sshslie_loss_scratch_bytes)");
    return SSHSLIE_ERR_WORKSPACE;
  }
  const int rc = ss_pixel_losses(x, R, I, Idelta, R_enh, *cfg, B, C, H, W, (float*)scratch, dR, dI, dIdelta, dS, dR_enh,
                                 (cudaStream_t)stream);
  if (rc) return rc;
  return ss_reduce_partials((const float*)scratch, ss_pixel_losses_blocks(B, C, H, W), 9, sums, 0, (cudaStream_t)stream);
}
