// Small device/host helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "plan.h"

#define SS_DEVINL __device__ __forceinline__

void ss_set_error(const char* fmt, ...);
int ss_check_launch(const char* what);
void ss_count_launches(int n);         // extra launches behind one ss_check_launch   // cudaGetLastError() -> SSHSLIE_ERR_CUDA + message

SS_DEVINL float bf2f(bf16 v) { return __bfloat162float(v); }
SS_DEVINL bf16 f2bf(float v) { return __float2bfloat16_rn(v); }

// unpack 8 bf16 (one 16-byte vector) into floats
SS_DEVINL void unpack8(const uint4& u, float* f) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
SS_DEVINL uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

SS_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum of `v` (blockDim.x multiple of 32, <= 1024); result valid in thread 0
SS_DEVINL float block_sum(float v, float* red /* >= 32 floats of smem */) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < (int)((blockDim.x + 31) >> 5)) ? red[lane] : 0.f;
    v = warp_sum(v);
  }
  return v;
}

SS_DEVINL float sigmoidf_(float v) { return 1.f / (1.f + __expf(-v)); }
SS_DEVINL float sgnf(float v) { return (v > 0.f) ? 1.f : ((v < 0.f) ? -1.f : 0.f); }


// ---------------------------------------------------------------------------------------------
// programmatic dependent launch for the small kernels between the conv GEMMs: every kernel of the main chain triggers
// its dependents at entry and waits for its predecessors before touching global memory, and is launched with the
// programmatic-stream-serialization attribute - the launch latency and prologue of kernel N+1 then overlap the tail of
// kernel N (measured on the 3x3 conv: 10.3 -> 7.4 us per dependent launch).  SSHSLIE_PDL=0 turns the attribute off.
// ---------------------------------------------------------------------------------------------
int ss_pdl_enabled();
#define SS_PDL_ENTRY()                                                   \
  do {                                                                   \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     \
    asm volatile("griddepcontrol.wait;" ::: "memory");                  \
  } while (0)
template <typename K, typename... A>
static inline void ss_launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, A... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ss_pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, args...);
}
