// CUDA-core executors of a ConvGeom (see plan.h): the cross-check twin of the tcgen05 kernels in
// conv_umma.cu (same packed bf16 operands, fp32 accumulation, same epilogue), the fallback for tile
// shapes the tensor-core kernel does not take, plus the weight packer and the bias-gradient reduction.
#include "common.cuh"
#include "epilogue.cuh"
#include "kernels.h"

// ---------------------------------------------------------------------------------------------
// gather GEMM:  out[m, n] = sum_slabs sum_j A_s[m, j] * Wp[n][s*64+j]
// block = 128 threads = 128 GEMM rows; blockIdx.y selects a 16-column slab of the output.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) conv_gather_simt_kernel(const ConvGeom* __restrict__ gp, Epi epi) {
  __shared__ ConvGeom g;
  __shared__ __align__(16) bf16 wsm[16 * SS_SLAB];
  {
    const int* src = reinterpret_cast<const int*>(gp);
    int* dst = reinterpret_cast<int*>(&g);
    for (int i = threadIdx.x; i < (int)(sizeof(ConvGeom) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int M = g.B * g.OH * g.OW;
  const int m = blockIdx.x * 128 + threadIdx.x;
  const int n0 = blockIdx.y * 16;
  const bool row_ok = m < M;
  int b = 0, oh = 0, ow = 0;
  if (row_ok) {
    b = m / (g.OH * g.OW);
    const int r = m - b * g.OH * g.OW;
    oh = r / g.OW;
    ow = r - oh * g.OW;
  }
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int s = 0; s < g.nslabs; ++s) {
    __syncthreads();
    {  // 16 rows x 64 bf16 = 128 x 16 bytes
      const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
      reinterpret_cast<uint4*>(wsm)[threadIdx.x] =
          *reinterpret_cast<const uint4*>(g.wp + ((size_t)s * g.Npad + n0 + r) * SS_SLAB + q * 8);
    }
    __syncthreads();
    const Slab sl = g.slab[s];
    const SrcView& v = g.src[sl.src];
    const int ih = oh + sl.dh, iw = ow + sl.dw;
    if (row_ok && ih >= 0 && ih < v.H && iw >= 0 && iw < v.W) {
      const bf16* ap = v.base + b * v.sB + ih * v.sH + iw * v.sW + sl.c0;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float a[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(ap) + q), a);
#pragma unroll
        for (int n = 0; n < 16; ++n) {
          float w[8];
          unpack8(reinterpret_cast<const uint4*>(wsm)[n * 8 + q], w);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[n] = fmaf(a[j], w[j], acc[n]);
        }
      }
    }
  }
  if (row_ok) epi_apply16(epi, b, oh, ow, n0, g.N, acc);
}

int ss_launch_conv_gather_simt(const ConvGeom* g_dev, const ConvGeom& g_host, const Epi& epi, cudaStream_t st) {
  const int M = g_host.B * g_host.OH * g_host.OW;
  dim3 grid((M + 127) / 128, g_host.Npad / 16);
  conv_gather_simt_kernel<<<grid, 128, 0, st>>>(g_dev, epi);
  return ss_check_launch("conv_gather_simt");
}

// ---------------------------------------------------------------------------------------------
// weight-gradient GEMM:  dW[n][s*64+j] += sum_m G[m, n] * A_s[m, j]      (fp32 atomics into the flat grads: this is the
// CUDA-core CROSS-CHECK kernel, run only under SSHSLIE_FLAG_FORCE_SIMT in tests - the product path never launches it)
// grid = (nslabs, ceil(Npad/32), splitM); block = 256 threads: thread -> (n = t/8, 8 channels j = (t%8)*8..)
// ---------------------------------------------------------------------------------------------
#define WG_ROWS 32
__global__ void __launch_bounds__(256) conv_wgrad_simt_kernel(const ConvGeom* __restrict__ gp,
                                                              const bf16* __restrict__ G, int64_t gB, int64_t gH,
                                                              int64_t gW, int gN, float* __restrict__ grads,
                                                              int rows_per_part) {
  __shared__ ConvGeom g;
  __shared__ __align__(16) bf16 asm_[WG_ROWS * SS_SLAB];
  __shared__ float gsm[WG_ROWS * 32];
  {
    const int* src = reinterpret_cast<const int*>(gp);
    int* dst = reinterpret_cast<int*>(&g);
    for (int i = threadIdx.x; i < (int)(sizeof(ConvGeom) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int s = blockIdx.x;
  const int nb = blockIdx.y * 32;
  const Slab sl = g.slab[s];
  const SrcView v = g.src[sl.src];
  const int M = g.B * g.OH * g.OW;
  const int m_begin = blockIdx.z * rows_per_part;
  const int m_end = min(M, m_begin + rows_per_part);
  const int tn = threadIdx.x >> 3, tq = threadIdx.x & 7;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int m0 = m_begin; m0 < m_end; m0 += WG_ROWS) {
    __syncthreads();
    {  // A tile: 32 rows x 64 ch (256 x 16B); G tile: 32 rows x 32 cols
      const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
      const int m = m0 + r;
      uint4 val = make_uint4(0, 0, 0, 0);
      float gv[4] = {0.f, 0.f, 0.f, 0.f};
      if (m < m_end) {
        const int b = m / (g.OH * g.OW);
        const int rr = m - b * g.OH * g.OW;
        const int oh = rr / g.OW, ow = rr - (rr / g.OW) * g.OW;
        const int ih = oh + sl.dh, iw = ow + sl.dw;
        if (ih >= 0 && ih < v.H && iw >= 0 && iw < v.W)
          val = __ldg(reinterpret_cast<const uint4*>(v.base + b * v.sB + ih * v.sH + iw * v.sW + sl.c0) + q);
        const bf16* gp2 = G + b * gB + oh * gH + ow * gW + nb + q * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) gv[i] = (nb + q * 4 + i < gN) ? bf2f(gp2[i]) : 0.f;
      }
      reinterpret_cast<uint4*>(asm_)[threadIdx.x] = val;
#pragma unroll
      for (int i = 0; i < 4; ++i) gsm[r * 32 + q * 4 + i] = gv[i];
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WG_ROWS; ++r) {
      const float gval = gsm[r * 32 + tn];
      float a[8];
      unpack8(reinterpret_cast<const uint4*>(asm_)[r * 8 + tq], a);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(gval, a[j], acc[j]);
    }
  }
  const int n = nb + tn;
  if (n < g.N && !sl.no_wgrad) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = tq * 8 + j;
      if (c < sl.wcn) atomicAdd(grads + g.w_off + (int64_t)n * g.w_sN + sl.woff + (int64_t)c * g.w_sC, acc[j]);
    }
  }
}

int ss_launch_conv_wgrad_simt(const ConvGeom* g_dev, const ConvGeom& g_host, const bf16* G, int64_t gB, int64_t gH,
                              int64_t gW, int gN, float* grads, cudaStream_t st) {
  const int M = g_host.B * g_host.OH * g_host.OW;
  int parts = (M + 1023) / 1024;
  if (parts > 64) parts = 64;
  int rows = (M + parts - 1) / parts;
  rows = (rows + WG_ROWS - 1) / WG_ROWS * WG_ROWS;
  parts = (M + rows - 1) / rows;
  dim3 grid(g_host.nslabs, (g_host.Npad + 31) / 32, parts);
  conv_wgrad_simt_kernel<<<grid, 256, 0, st>>>(g_dev, G, gB, gH, gW, gN, grads, rows);
  return ss_check_launch("conv_wgrad_simt");
}

// ---------------------------------------------------------------------------------------------
// bias gradient: db[n] += sum over pixels of G[pixel, n]   (G: bf16, pixel stride `ld`, N <= 256 columns)
// Every block sums a contiguous pixel range into its row of `partials`; a second launch adds the rows in a fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_grad_kernel(const bf16* __restrict__ G, int64_t npix, int ld, int N,
                                                        float* __restrict__ partials, int pix_per_block) {
  // thread -> column n = t % 64 (+64 ...), row phase = t / 64
  __shared__ float red[256];
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = min(npix, p0 + pix_per_block);
  for (int nb = 0; nb < N; nb += 64) {
    const int n = nb + (threadIdx.x & 63);
    float s = 0.f;
    if (n < N)
      for (int64_t p = p0 + (threadIdx.x >> 6); p < p1; p += 4) s += bf2f(G[p * ld + n]);
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 64 && n < N)
      partials[(size_t)blockIdx.x * N + n] =
          red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192];
    __syncthreads();
  }
}

int ss_launch_bias_grad(const bf16* G, int64_t npix, int ld, int N, float* db, float* scratch, cudaStream_t st) {
  if (!scratch) { ss_set_error("bias_grad: scratch missing"); return SSHSLIE_ERR_WORKSPACE; }
  int64_t ppb = 64;
  while ((npix + ppb - 1) / ppb > SS_BIAS_GRAD_MAX_BLOCKS) ppb *= 2;
  const int nblk = (int)((npix + ppb - 1) / ppb);
  bias_grad_kernel<<<nblk, 256, 0, st>>>(G, npix, ld, N, scratch, (int)ppb);
  int rc = ss_check_launch("bias_grad");
  if (rc) return rc;
  RedSegs segs;
  memset(&segs, 0, sizeof(segs));
  segs.n = 1; segs.dst[0] = db; segs.len[0] = N;
  return ss_launch_reduce_rows(scratch, nblk, N, segs, st);
}

// ---------------------------------------------------------------------------------------------
// weight packer: fp32 master weights -> bf16 Wp[Npad][nslabs*64] for every geom of the plan, one launch.
// job table: block_start[j] .. block_start[j+1] blocks of 256 elements belong to geom j.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_weights_kernel(const ConvGeom* __restrict__ geoms,
                                                           const int* __restrict__ block_start, int njobs,
                                                           const float* __restrict__ params, int first_block) {
  SS_PDL_ENTRY();
  const int blk = (int)blockIdx.x + first_block;      // a launch may cover only blocks [first_block, ...) of the table
  int lo = 0, hi = njobs - 1;           // last job whose first block is <= blk (block_start is ascending)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (blk >= block_start[mid]) lo = mid; else hi = mid - 1;
  }
  const int j = lo;
  const ConvGeom* g = geoms + j;
  const int Ktot = g->nslabs * SS_SLAB;
  const int64_t idx = (int64_t)(blk - block_start[j]) * 256 + threadIdx.x;
  if (idx >= (int64_t)g->Npad * Ktot) return;
  const int n = (int)(idx / Ktot);
  const int k = (int)(idx - (int64_t)n * Ktot);
  const int s = k >> 6, c = k & 63;
  const Slab sl = g->slab[s];
  float w = 0.f;
  const int cw = (g->dup_c >= 0 && sl.c0 + c == g->dup_c) ? c - 1 : c;      // a residual lane reads its hi lane's weight
  if (n < g->N && cw >= 0 && cw < sl.wcn) w = __ldg(params + g->w_off + (int64_t)n * g->w_sN + sl.woff + (int64_t)cw * g->w_sC);
  g->wp[((size_t)s * g->Npad + n) * SS_SLAB + c] = f2bf(w);   // slab-major: one slab = Npad x 64 contiguous bf16
}

int ss_launch_pack_weights(const ConvGeom* geoms_dev, const int* block_start_dev, int njobs, int total_blocks,
                           const float* params, cudaStream_t st, int first_block, int n_blocks) {
  if (n_blocks < 0) n_blocks = total_blocks - first_block;
  if (n_blocks <= 0) return SSHSLIE_OK;
  ss_launch_pdl(pack_weights_kernel, dim3(n_blocks), dim3(256), (size_t)0, st, geoms_dev, block_start_dev, njobs, params, first_block);
  return ss_check_launch("pack_weights");
}
