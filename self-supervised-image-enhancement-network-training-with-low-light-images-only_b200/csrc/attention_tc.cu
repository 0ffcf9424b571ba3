// Tensor-core attention core of the TransformerBlock (model.py:107-114) for LARGE token grids (full-image inference:
// 512 x 512 -> L = 4096 tokens, where the fp32 CUDA-core kernel of attention.cu needs 0.37 ms for 1 % of the FLOPs).
//
//   o_i = sum_j softmax_j(q_i . k_j / 4) v_j        per (image, head), 4 heads x 16 features, never materialising (L, L)
//
// One CTA = (image, head, tile of 128 queries).  Keys / values stream through shared memory in tiles of 128 rows; per tile
//   S = Q K^T      ONE accumulator tile 128 x 128 in TMEM, three K = 16 tcgen05.mma: q_hi k_hi + q_lo k_hi + q_hi k_lo
//                  (bf16 hi + lo pairs: the logits are exact to ~2^-16, as the fp32 reference needs them to be)
//   P = exp(S - m) each of 128 threads owns one query row: tcgen05.ld -> exp2 -> bf16 hi + lo pair -> two SWIZZLE_128B
//                  tiles in shared memory (with plain bf16 probabilities o is only bf16-accurate, and the backward's
//                  D_i = dO_i . o_i amplifies that: 0.2 % of dx came out wrong by up to 6 % of its range)
//   O += P [V]     2 x eight K = 16 MMAs (P_hi, P_lo) with the SAME 128-row tile as MN-major B operand (N = 64: the tile's
//                  columns are k_hi | k_lo | v_hi | v_lo, the product's columns 32..63 are P v_hi and P v_lo)
// The row maximum m is exact, not running: pass A walks all key tiles computing only max_j S_ij (the MMAs are ~3 % of the
// pass), pass B recomputes S and accumulates.  No accumulator rescaling, no dependence of the result on the tile order.
// Softmax is MUFU-bound (16 exp2 / clk / SM): 128 x 128 exponentials per tile = 1024 clk; 4 x 4096^2 / (128 SMs x 16) ->
// ~30 us for the 512 x 512 cube.
//
// Operand layout (written by attn_pack_heads_kernel from the fp32 q, k, v of attn_qkv4_kernel): per (image, head) L rows
// of 64 bf16 = 128 bytes, Qp row = [0.25 q_hi | 0.25 q_lo | 0 | 0], KVp row = [k_hi | k_lo | v_hi | v_lo], so that one
// 16 KB TMA tile feeds both GEMMs of a key tile.
//
// Warp roles (576 threads), ordered by the SM's issue priority (highest warp id first): warps 0..7 / 8..15 = softmax group
// 0 / 1 (TMEM lane quarter = warp & 3; warps 4..7 of a group take keys 64..127 of the tile, warps 0..3 keys 0..63: four
// softmax warps per scheduler hide the TMEM-load and exp2 latencies that two could not - ncu: 35 % issue-active, average
// warp latency 7 clk per instruction), warp 16 = TMA producer, warp 17 = TMEM owner + MMA issuer.  Group g takes the key tiles with c & 1 == g (S buffer g, P buffer g); one warp per
// scheduler could not hide the TMEM-load latency (99 us for the 512 x 512 cube).  Row maxima / sums of the two groups are
// merged through shared memory in a fixed order.
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "common.cuh"
#include "kernels.h"
#include "umma_ptx.cuh"

#define TC_TILE 128
#define TC_SLOTS 4
#define TC_THREADS 576
#define TC_TILE_BYTES (TC_TILE * 128)
#define TC_P_BYTES (4 * TC_TILE_BYTES)          // probabilities of one key tile: bf16 hi (2 sub-tiles of 64 keys) + bf16 lo (2)
#define TC_SMEM (TC_TILE_BYTES + TC_SLOTS * TC_TILE_BYTES + 2 * TC_P_BYTES + 1024)
#define TC_LOG2E 1.4426950408889634f

SS_DEVINL float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// q, k, v: fp32 [B*L][64] (head h = columns 16h..16h+15)  ->  Qp, KVp: bf16 [(b*4 + h)*L + i][64]
__global__ void __launch_bounds__(256) attn_pack_heads_kernel(const float* __restrict__ Q, const float* __restrict__ K,
                                                              const float* __restrict__ V, bf16* __restrict__ Qp,
                                                              bf16* __restrict__ KVp, int B, int L) {
  SS_PDL_ENTRY();
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (token, head)
  if (idx >= (int64_t)B * L * 4) return;
  const int h = (int)(idx & 3);
  const int64_t t = idx >> 2;
  const int b = (int)(t / L), i = (int)(t - (int64_t)b * L);
  const int64_t src = t * 64 + h * 16, dst = (((int64_t)b * 4 + h) * L + i) * 64;
  float q[16], k[16], v[16];
#pragma unroll
  for (int c = 0; c < 16; c += 4) {
    const float4 a = *reinterpret_cast<const float4*>(Q + src + c);
    const float4 e = *reinterpret_cast<const float4*>(K + src + c);
    const float4 f = *reinterpret_cast<const float4*>(V + src + c);
    q[c] = 0.25f * a.x; q[c + 1] = 0.25f * a.y; q[c + 2] = 0.25f * a.z; q[c + 3] = 0.25f * a.w;   // 1/sqrt(16), exact
    k[c] = e.x; k[c + 1] = e.y; k[c + 2] = e.z; k[c + 3] = e.w;
    v[c] = f.x; v[c + 1] = f.y; v[c + 2] = f.z; v[c + 3] = f.w;
  }
  auto split = [](const float* x, uint4& h0, uint4& h1, uint4& l0, uint4& l1) {
    float lo[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) lo[c] = x[c] - bf2f(f2bf(x[c]));
    h0 = make_uint4(pack2(x[0], x[1]), pack2(x[2], x[3]), pack2(x[4], x[5]), pack2(x[6], x[7]));
    h1 = make_uint4(pack2(x[8], x[9]), pack2(x[10], x[11]), pack2(x[12], x[13]), pack2(x[14], x[15]));
    l0 = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
    l1 = make_uint4(pack2(lo[8], lo[9]), pack2(lo[10], lo[11]), pack2(lo[12], lo[13]), pack2(lo[14], lo[15]));
  };
  uint4 a0, a1, b0, b1;
  uint4* qo = reinterpret_cast<uint4*>(Qp + dst);
  uint4* ko = reinterpret_cast<uint4*>(KVp + dst);
  split(q, a0, a1, b0, b1);
  qo[0] = a0; qo[1] = a1; qo[2] = b0; qo[3] = b1;
  qo[4] = qo[5] = qo[6] = qo[7] = make_uint4(0, 0, 0, 0);
  split(k, a0, a1, b0, b1);
  ko[0] = a0; ko[1] = a1; ko[2] = b0; ko[3] = b1;
  split(v, a0, a1, b0, b1);
  ko[4] = a0; ko[5] = a1; ko[6] = b0; ko[7] = b1;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
attn_core_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kvmap,
                    float* __restrict__ O, float* __restrict__ LSE, int L) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t q_full;
  __shared__ __align__(8) uint64_t kv_full[TC_SLOTS];
  __shared__ __align__(8) uint64_t kv_empty[TC_SLOTS];
  __shared__ __align__(8) uint64_t s_full[3];          // THREE logit tiles in TMEM (columns 0 / 128 / 256), O at 384
  __shared__ __align__(8) uint64_t s_empty[3];
  __shared__ __align__(8) uint64_t p_full[2];
  __shared__ __align__(8) uint64_t p_empty[2];
  __shared__ __align__(8) uint64_t o_full;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[4][TC_TILE];                  // row maxima, then row sums, of the 2 groups x 2 key halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  const uint32_t dyn_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* dyn_ptr = smem_dyn + (dyn_base - smem_u32(smem_dyn));
  const uint32_t q_addr = dyn_base, kv_addr = dyn_base + TC_TILE_BYTES, p_addr = kv_addr + TC_SLOTS * TC_TILE_BYTES;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nk = (L + TC_TILE - 1) / TC_TILE;
  const int row_base = (b * 4 + head) * L;           // first row of this (image, head) in Qp / KVp
  const int n_tiles = 2 * nk;                        // pass A (row maxima) + pass B (softmax . V)

  if (warp == 16 && lane == 0) {
    mbar_init(smem_u32(&q_full), 1);
    for (int s = 0; s < TC_SLOTS; ++s) { mbar_init(smem_u32(&kv_full[s]), 1); mbar_init(smem_u32(&kv_empty[s]), 1); }
    for (int s = 0; s < 3; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), 8);           // one arrival per softmax warp of the group that read it
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&p_full[s]), 8);
      mbar_init(smem_u32(&p_empty[s]), 1);
    }
    mbar_init(smem_u32(&o_full), 1);
    fence_barrier_init();
  }
  if (warp == 17) tmem_alloc(smem_u32(&tmem_base_smem), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 16) {
    // ===== TMA producer: the query tile once, then the key/value tiles of both passes through a ring =====
    if (lane == 0) {
      pdl_wait();
      mbar_expect_tx(smem_u32(&q_full), TC_TILE_BYTES);
      tma_load_4d(q_addr, &qmap, smem_u32(&q_full), 0, row_base + qt * TC_TILE, 0, 0);
      uint32_t slot = 0, use = 0;
      for (int c = 0; c < n_tiles; ++c) {
        const int kt = (c < nk) ? c : c - nk;
        mbar_wait_polite(smem_u32(&kv_empty[slot]), (use & 1u) ^ 1u);
        mbar_expect_tx(smem_u32(&kv_full[slot]), TC_TILE_BYTES);
        tma_load_4d(kv_addr + slot * TC_TILE_BYTES, &kvmap, smem_u32(&kv_full[slot]), 0, row_base + kt * TC_TILE, 0, 0);
        if (++slot == TC_SLOTS) { slot = 0; ++use; }
      }
    }
  } else if (warp == 17) {
    // ===== MMA issuer =====
    const uint32_t idesc_qk = make_idesc(128, 128, 0, 0);          // A = Q (K-major), B = K rows (K-major), N = 128 keys
    const uint32_t idesc_pv = make_idesc(128, 64, 0, 1);           // A = P (K-major), B = the same tile MN-major, N = 64
    const uint32_t tm = uniform32(tmem_base);
    const uint32_t d_hi = (uint32_t)(make_sdesc(0, 16, 1024) >> 32);
    const uint32_t q_lo = uniform32(((q_addr >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t kv_lo0 = uniform32(((kv_addr >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t p_lo0 = uniform32(((p_addr >> 4) & 0x3FFFu) | (1u << 16));
    const uint32_t vb_hi = (uint32_t)(make_sdesc(0, TC_TILE_BYTES, 1024) >> 32);     // MN-major B: K atoms 1024 B apart
    const uint32_t vb_lo0 = uniform32((kv_addr >> 4) & 0x3FFFu);
    mbar_wait_warp(smem_u32(&q_full), 0, 0);
    auto issue_qk = [&](int c) {
      const uint32_t slot = (uint32_t)(c % TC_SLOTS), kph = (uint32_t)((c / TC_SLOTS) & 1);
      const uint32_t sb = (uint32_t)(c % 3), sph = (uint32_t)((c / 3) & 1);
      mbar_wait_warp(smem_u32(&kv_full[slot]), kph, 0);
      mbar_wait_warp(smem_u32(&s_empty[sb]), sph ^ 1u, 0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t k_lo = kv_lo0 + slot * (TC_TILE_BYTES >> 4);
        const uint32_t td = tm + sb * 128u;
        // q_hi k_hi + q_lo k_hi + q_hi k_lo   (K offsets inside the 128-byte rows: 32 bytes = 2 descriptor units)
        umma_bf16(td, ((uint64_t)d_hi << 32) | (uint64_t)(q_lo + 0u), ((uint64_t)d_hi << 32) | (uint64_t)(k_lo + 0u), idesc_qk, 0u);
        umma_bf16(td, ((uint64_t)d_hi << 32) | (uint64_t)(q_lo + 2u), ((uint64_t)d_hi << 32) | (uint64_t)(k_lo + 0u), idesc_qk, 1u);
        umma_bf16(td, ((uint64_t)d_hi << 32) | (uint64_t)(q_lo + 0u), ((uint64_t)d_hi << 32) | (uint64_t)(k_lo + 2u), idesc_qk, 1u);
        umma_commit(smem_u32(&s_full[sb]));
        if (c < nk) umma_commit(smem_u32(&kv_empty[slot]));        // pass A: the tile is not needed again
      }
      __syncwarp();
    };
    // the logits run TWO tiles ahead of the softmax (three S buffers for two softmax groups): when a group has finished
    // tile c, S of its next tile c + 2 is already in TMEM - with two buffers each group idled for QK + P V + barrier
    // latency (~1500 clk) per tile
    issue_qk(0);
    if (n_tiles > 1) issue_qk(1);
    uint32_t pcount[2] = {0u, 0u};                   // pass-B uses of each P buffer so far
#pragma unroll 1
    for (int c = 0; c < n_tiles; ++c) {
      if (c + 2 < n_tiles) issue_qk(c + 2);
      if (c >= nk) {
        const int j = c - nk;
        const uint32_t slot = (uint32_t)(c % TC_SLOTS), pb = (uint32_t)(c & 1), pph = pcount[c & 1] & 1u;
        ++pcount[c & 1];
        mbar_wait_warp(smem_u32(&p_full[pb]), pph, 0);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a0 = p_lo0 + pb * (TC_P_BYTES >> 4);
          const uint32_t b0 = vb_lo0 + slot * (TC_TILE_BYTES >> 4);
          const uint32_t td = tm + 384u;
#pragma unroll
          for (int part = 0; part < 2; ++part)    // P_hi, then P_lo (sub-tiles 2, 3 of the buffer)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)        // 16 keys per step: P columns 16 kk.. (two 64-key sub-tiles), tile rows 16 kk..
              umma_bf16(td, ((uint64_t)d_hi << 32) | (uint64_t)(a0 + (uint32_t)(2 * part + (kk >> 2)) * (TC_TILE_BYTES >> 4) + 2u * (kk & 3)),
                        ((uint64_t)vb_hi << 32) | (uint64_t)(b0 + 128u * kk), idesc_pv, (j > 0 || kk > 0 || part > 0) ? 1u : 0u);
          umma_commit(smem_u32(&p_empty[pb]));
          umma_commit(smem_u32(&kv_empty[slot]));
          if (c == n_tiles - 1) umma_commit(smem_u32(&o_full));
        }
        __syncwarp();
      }
    }
  } else {
    // ===== softmax warps: thread = query row; group g = key tiles with c & 1 == g =====
    const int g = warp >> 3, half = (warp >> 2) & 1;     // half: keys 0..63 / 64..127 of every tile
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int nlo = half * 64;
    const uint32_t pbuf = (uint32_t)g;
    const uint32_t trow0 = tmem_base + ((uint32_t)(quarter * 32) << 16);
    unsigned char* prow = dyn_ptr + (p_addr - dyn_base) + pbuf * TC_P_BYTES + half * TC_TILE_BYTES + row * 128;
    const int sw = row & 7;
    float m = -INFINITY, l = 0.f;
    uint32_t pu = 0;                                 // uses of this group's P buffer so far
    pdl_wait();
    // ---- pass A: row maxima
#pragma unroll 1
    for (int c = g; c < nk; c += 2) {
      const int valid = min(TC_TILE, L - c * TC_TILE);
      const uint32_t sb = (uint32_t)(c % 3), trow = trow0 + sb * 128u;
      mbar_wait_warp_polite(smem_u32(&s_full[sb]), (uint32_t)((c / 3) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int n0 = nlo; n0 < nlo + 64; n0 += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)n0, v);
        if (valid == TC_TILE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = fmaxf(m, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) m = (n0 + i < valid) ? fmaxf(m, v[i]) : m;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[sb]));
    }
    xch[2 * g + half][row] = m;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    m = fmaxf(fmaxf(xch[0][row], xch[1][row]), fmaxf(xch[2][row], xch[3][row]));
    const float ml2 = m * TC_LOG2E;
    asm volatile("bar.sync 1, 512;" ::: "memory");      // everybody has read the maxima before the sums reuse xch
    // ---- pass B: p = exp(s - m) -> bf16 hi + lo tiles, l += p
    const int c0 = nk + ((nk + g) & 1);                  // first pass-B tile with c & 1 == g
#pragma unroll 1
    for (int c = c0; c < n_tiles; c += 2, ++pu) {
      const int kt = c - nk;
      const int valid = min(TC_TILE, L - kt * TC_TILE);
      const uint32_t sb = (uint32_t)(c % 3), trow = trow0 + sb * 128u;
      mbar_wait_warp_polite(smem_u32(&s_full[sb]), (uint32_t)((c / 3) & 1));
      mbar_wait_warp_polite(smem_u32(&p_empty[pbuf]), (pu & 1u) ^ 1u);  // the MMAs that read this P buffer have retired
      tc_fence_after();
#pragma unroll 1
      for (int n0 = nlo; n0 < nlo + 64; n0 += 32) {
        float v[32];
        tmem_ld32(trow + (uint32_t)n0, v);
        if (valid == TC_TILE) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = ex2_approx(fmaf(v[i], TC_LOG2E, -ml2));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = (n0 + i < valid) ? ex2_approx(fmaf(v[i], TC_LOG2E, -ml2)) : 0.f;
        }
        {      // four independent partial sums (a single chain of 32 dependent adds stalls the warp), fixed order
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) { a0 += v[i]; a1 += v[i + 1]; a2 += v[i + 2]; a3 += v[i + 3]; }
          l += (a0 + a1) + (a2 + a3);
        }
        unsigned char* sub = prow;                                // this warp's 64-key K-major SWIZZLE_128B sub-tile
        const int j0 = (n0 & 63) >> 3;
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          // bf16 hi + lo pair by TRUNCATION with integer ops: hi = upper 16 bits of p, lo = upper 16 bits of (p - hi).
          // hi + lo carries 16 mantissa bits like the rounded pair, but costs no F2F conversions - those share the
          // 16-lane MUFU pipe with exp2, which is this kernel's bottleneck (measured: 3000 -> ~1000 clk per key tile)
          uint32_t hb[8], lb[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint32_t bits = __float_as_uint(v[8 * q4 + i]);
            hb[i] = bits;
            lb[i] = __float_as_uint(v[8 * q4 + i] - __uint_as_float(bits & 0xFFFF0000u));
          }
          uint4 u, w;
          u.x = __byte_perm(hb[0], hb[1], 0x7632); u.y = __byte_perm(hb[2], hb[3], 0x7632);
          u.z = __byte_perm(hb[4], hb[5], 0x7632); u.w = __byte_perm(hb[6], hb[7], 0x7632);
          w.x = __byte_perm(lb[0], lb[1], 0x7632); w.y = __byte_perm(lb[2], lb[3], 0x7632);
          w.z = __byte_perm(lb[4], lb[5], 0x7632); w.w = __byte_perm(lb[6], lb[7], 0x7632);
          *reinterpret_cast<uint4*>(sub + (((j0 + q4) ^ sw) << 4)) = u;
          *reinterpret_cast<uint4*>(sub + 2 * TC_TILE_BYTES + (((j0 + q4) ^ sw) << 4)) = w;      // the lo pair
        }
      }
      tc_fence_before();
      fence_proxy_async();                                        // P -> visible to the tensor core
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&s_empty[sb]));
        mbar_arrive(smem_u32(&p_full[pbuf]));
      }
    }
    xch[2 * g + half][row] = l;
    asm volatile("bar.sync 1, 512;" ::: "memory");
    l = (xch[0][row] + xch[1][row]) + (xch[2][row] + xch[3][row]);
    // ---- O columns 32..47 = P v_hi, 48..63 = P v_lo   (the first four warps write the tile's output)
    if (g == 0 && half == 0) {
      mbar_wait_warp_polite(smem_u32(&o_full), 0);
      tc_fence_after();
      float v[32];
      tmem_ld32(trow0 + 384u + 32u, v);
      const int i = qt * TC_TILE + row;
      if (i < L) {
        const float inv = 1.f / l;
        float* op = O + ((int64_t)b * L + i) * 64 + head * 16;
#pragma unroll
        for (int d = 0; d < 16; d += 4)
          *reinterpret_cast<float4*>(op + d) = make_float4((v[d] + v[16 + d]) * inv, (v[d + 1] + v[17 + d]) * inv,
                                                           (v[d + 2] + v[18 + d]) * inv, (v[d + 3] + v[19 + d]) * inv);
        if (LSE) LSE[((int64_t)b * 4 + head) * L + i] = m + __logf(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) tmem_dealloc(tmem_base, 512);
}

int ss_attention_core_tc(const float* q, const float* k, const float* v, bf16* qp, bf16* kvp, float* o, float* lse, int B,
                         int L, cudaStream_t st) {
  const int64_t rows = (int64_t)B * 4 * L;
  ss_launch_pdl(attn_pack_heads_kernel, dim3((unsigned)((rows + 255) / 256)), dim3(256), (size_t)0, st, q, k, v, qp, kvp, B, L);
  int rc = ss_check_launch("attn_pack_heads");
  if (rc) return rc;
  CUtensorMap qmap, kvmap;
  SrcView vq, vk;
  vq.base = qp; vq.sW = 64; vq.sH = 64 * rows; vq.sB = 64 * rows; vq.W = (int)rows; vq.H = 1;
  vk = vq; vk.base = kvp;
  rc = ss_umma_encode_view(vq, 64, TC_TILE, 1, 1, &qmap);
  if (!rc) rc = ss_umma_encode_view(vk, 64, TC_TILE, 1, 1, &kvmap);
  if (rc) return rc;
  static DeviceOnce attr_once;
  if (!attr_once.done()) {
    if (cudaFuncSetAttribute(attn_core_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM) != cudaSuccess) {
      ss_set_error("attn_core_tc: cannot raise dynamic shared memory: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
    attr_once.set();
  }
  dim3 grid((L + TC_TILE - 1) / TC_TILE, 4, B);
  ss_launch_pdl(attn_core_tc_kernel, grid, dim3(TC_THREADS), (size_t)TC_SMEM, st, qmap, kvmap, o, lse, L);
  return ss_check_launch("attn_core_tc");
}
