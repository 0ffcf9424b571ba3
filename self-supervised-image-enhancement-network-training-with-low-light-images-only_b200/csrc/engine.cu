// Host-side launch planner ("engine") and the C-ABI of include/sshslie_b200.h.
//
// An engine is built for one (batch, 64 bands, H, W).  Binding it to a caller-owned workspace lays out every
// activation / gradient tensor with a bump allocator, lowers each conv-like layer of the reference network
// (model.py:25-70, 121-175) to ConvGeoms (plan.h), builds the TMA descriptors of the tcgen05 path and records the
// forward and backward launch sequences.  Running a step only enqueues those launches on the caller's stream.
#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <array>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.cuh"
#include "kernels.h"

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void ss_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
// profiling notes: every launcher ends in ss_check_launch(name); the profiled entry point collects them per op
static thread_local bool g_prof_on = false;
static thread_local std::string g_prof_names;
static thread_local double g_prof_flops = 0, g_prof_bytes = 0;
static void prof_note(const std::string& label, double flops, double bytes) {
  if (!g_prof_on) return;
  g_prof_names = label;
  g_prof_flops += flops;
  g_prof_bytes += bytes;
}
static std::atomic<long long> g_launch_count{0};  // kernels of this library enqueued so far (bench.py's gpu_launches)
void ss_count_launches(int n) { g_launch_count += n; }
extern "C" SSHSLIE_API long long sshslie_launch_count(void) { return g_launch_count.load(); }
int ss_check_launch(const char* what) {
  g_launch_count += 1;
  if (g_prof_on && g_prof_names.find(':') == std::string::npos) {
    if (!g_prof_names.empty()) g_prof_names += "+";
    g_prof_names += what;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    ss_set_error("%s: %s", what, cudaGetErrorString(e));
    return SSHSLIE_ERR_CUDA;
  }
  return SSHSLIE_OK;
}
extern "C" const char* sshslie_last_error(void) { return g_err; }
extern "C" int sshslie_version(void) { return 100; }

// ---------------------------------------------------------------------------------------------
// parameter table (reference state_dict order)
// ---------------------------------------------------------------------------------------------
enum Layer {
  L_D_CONV0 = 0, L_D_SHALLOW, L_D_CONV1, L_D_CONV2, L_D_CONV3, L_D_DECONV, L_D_CONV5, L_D_CONV7, L_D_RECON,
  L_I_CONV0, L_I_CONV1, L_I_CONV2, L_I_CONV3, L_Q, L_K, L_V, L_FF1, L_FF2, L_I_DECONV1, L_I_DECONV2, L_I_DECONV3,
  L_I_FUSION, L_I_FINAL, L_COUNT
};
static const char* kLayerNames[] = {
  "d.conv0", "d.shallow9x9", "d.conv1", "d.conv2s2", "d.conv3", "d.deconv", "d.conv5", "d.conv7", "d.recon",
  "i.conv0", "i.conv1s2", "i.conv2s2", "i.conv3s2", "attn.q", "attn.k", "attn.v", "attn.ff1", "attn.ff2",
  "i.deconv1", "i.deconv2", "i.deconv3", "i.fusion1x1", "i.final"};
struct LayerShape { int d0, d1, k; bool transposed; };   // weight (d0, d1, k, k); k = 0 -> Linear (d0, d1)
static void layer_shapes(int C, LayerShape* s) {
  s[L_D_CONV0] = {32, C, 3, false};     s[L_D_SHALLOW] = {64, C, 9, false};  s[L_D_CONV1] = {64, 64, 3, false};
  s[L_D_CONV2] = {128, 64, 3, false};   s[L_D_CONV3] = {128, 128, 3, false}; s[L_D_DECONV] = {128, 64, 3, true};
  s[L_D_CONV5] = {64, 128, 3, false};   s[L_D_CONV7] = {64, 96, 3, false};   s[L_D_RECON] = {C + 1, 64, 3, false};
  s[L_I_CONV0] = {64, C + 1, 3, false}; s[L_I_CONV1] = {64, 64, 3, false};   s[L_I_CONV2] = {64, 64, 3, false};
  s[L_I_CONV3] = {64, 64, 3, false};    s[L_Q] = {64, 64, 0, false};         s[L_K] = {64, 64, 0, false};
  s[L_V] = {64, 64, 0, false};          s[L_FF1] = {64, 64, 0, false};       s[L_FF2] = {64, 64, 0, false};
  s[L_I_DECONV1] = {64, 64, 3, false};  s[L_I_DECONV2] = {64, 64, 3, false}; s[L_I_DECONV3] = {64, 64, 3, false};
  s[L_I_FUSION] = {64, 192, 1, false};  s[L_I_FINAL] = {1, 64, 3, false};
}
extern "C" int64_t sshslie_param_table(int channels, int64_t* offsets, int64_t* sizes) {
  LayerShape s[L_COUNT];
  layer_shapes(channels, s);
  int64_t off = 0;
  for (int l = 0; l < L_COUNT; ++l) {
    const int64_t kk = s[l].k ? (int64_t)s[l].k * s[l].k : 1;
    const int64_t wsz = (int64_t)s[l].d0 * s[l].d1 * kk;
    const int64_t bsz = s[l].transposed ? s[l].d1 : s[l].d0;
    if (offsets) { offsets[2 * l] = off; offsets[2 * l + 1] = off + wsz; }
    if (sizes) { sizes[2 * l] = wsz; sizes[2 * l + 1] = bsz; }
    off += wsz + bsz;
  }
  return off;
}

// ---------------------------------------------------------------------------------------------
// planner helpers
// ---------------------------------------------------------------------------------------------
struct Tens {          // bf16 NHWC tensor
  bf16* p = nullptr;
  int B = 0, H = 0, W = 0, ld = 0;
  int64_t pix() const { return (int64_t)B * H * W; }
};
struct SrcSpec { Tens t; int c_off; int c_cnt; int wc_off; int lo = 0; };
struct WAddr { int64_t w_off; int sN, sC, sKH, sKW; };

struct GeomOp {        // one lowered GEMM: geometry + (for gathers) epilogue template
  ConvGeom g;
  int geom_index = -1;
  bool use_umma = false;
};

struct alignas(64) GMapBox { unsigned char bytes[128]; };
#define SS_MAX_SIDE 8

struct sshslie_engine {
  int B, C, H, W, flags;
  std::map<std::pair<int, const void*>, GMapBox> gmaps;   // wgrad G-tensor TMA descriptors
  bool train, force_simt;
  bool skip_wgrad = false;
  bool halo_on = true;
  bool wgrad_halo = true;               // SSHSLIE_WGRAD_HALO=0 keeps the per-tap weight-gradient kernel
  int64_t ws_bytes = 0;
  unsigned char* ws = nullptr;
  bool bound = false;
  int64_t poff[2 * L_COUNT], psize[2 * L_COUNT], nparams;
  LayerShape shapes[L_COUNT];

  // bump allocator (dry run when base == nullptr)
  unsigned char* base = nullptr;
  int64_t cursor = 0;
  std::vector<std::pair<int64_t, int64_t>> zero_ranges;   // regions zeroed at bind (padding lanes)
  void* alloc(int64_t bytes, bool zero = false) {
    const int64_t off = cursor;
    cursor += (bytes + 1023) / 1024 * 1024;
    if (zero) zero_ranges.push_back({off, bytes});
    return base ? (void*)(base + off) : (void*)(uintptr_t)(off + 1024);   // non-null placeholder in dry runs
  }
  Tens talloc(int b, int h, int w, int ld, bool zero = false) {
    Tens t;
    t.B = b; t.H = h; t.W = w; t.ld = ld;
    t.p = (bf16*)alloc((int64_t)b * h * w * ld * sizeof(bf16), zero);
    return t;
  }
  float* falloc(int64_t n, bool zero = false) { return (float*)alloc(n * sizeof(float), zero); }

  // plan
  std::vector<ConvGeom> geoms;           // host copy, index = geom id
  std::vector<char> geom_umma;           // 1 if the tcgen05 kernel takes this geom
  std::vector<int> geom_layer;           // which layer's weights the geom reads (for profiles)
  std::vector<int> geom_role;            // 0 = forward-type addressing, 1 = dgrad-type addressing
  mutable int last_layer = -1, last_role = 0;
  std::vector<unsigned char> maps_blob;  // UmmaMaps per geom
  // persistent pipelined gather kernel (conv_pipe.cu): per geom, decided and planned at its first launch
  std::vector<unsigned char> pipe_blob;
  std::vector<int> pipe_valid;
  std::vector<char> pipe_use;            // 0 = undecided, 1 = pipelined kernel, 2 = halo kernel
  bool pipe_on = true;
  int fwd_pack_begin = -1, fwd_illum_begin = -1, fwd_illum_end = -1;   // op ranges of ops_fwd (sshslie_illum_forward)
  bf16* RI_ptr = nullptr;
  int s2_min_tiles = 512;                // stride-2 / transposed layers: halo-reuse kernels from this many tiles (per class)
  int pipe_min_tiles_head = 1024;        // same threshold for the sigmoid head (its staged fp32 stores are the gain)
  int pipe_min_tiles = 1024;             // below this the halo kernel (2-3 small co-resident CTAs per SM) has the lower latency
  int pipe_max_slabs = 36;               // the 9x9 layer (81 streamed slabs) is 8 % faster on the halo kernel
  ConvGeom* geoms_dev = nullptr;
  int* pack_start_dev = nullptr;
  std::vector<int> pack_start;
  int pack_blocks = 0, pack_split = 0;
  float* mask_dev = nullptr;
  // per-block partial sums of the loss kernels, of the thin weight-gradient kernels and of the bias gradient: everything
  // that used to be an fp32 atomicAdd is a fixed-order reduction over these (deterministic step, main.py:165)
  float *pix_partials = nullptr, *four_partials = nullptr, *final_partials = nullptr, *attn_partials = nullptr;
  float* bias_partials[SS_MAX_SIDE] = {};   // one per side stream
  int pix_rows = 0;
  // split-K partial accumulators of the tcgen05 wgrad (sized for the largest op), one buffer per side stream
  float* wg_partial[SS_MAX_SIDE] = {};
  size_t wg_partial_floats = 0;
  int64_t* attn_poff_dummy = nullptr;

  // side streams for weight gradients (created at bind; host objects only): consecutive wgrad launches rotate over
  // them, each with its own split-K partial buffer.  In the backward phases the weight-gradient kernels add up to more
  // device time than the data-gradient chain they hide behind, so two streams were not enough to keep up with it.
  int n_side = 4;
  cudaStream_t side[SS_MAX_SIDE] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[SS_MAX_SIDE] = {};
  bool side_dirty[SS_MAX_SIDE] = {}, use_side = true;
  int side_rr = 0;
  cudaStream_t fork(cudaStream_t main_st) {
    if (!use_side || !side[0]) return main_st;
    const int i = (side_rr++) % n_side;
    cudaEventRecord(ev_fork, main_st);
    cudaStreamWaitEvent(side[i], ev_fork, 0);
    side_dirty[i] = true;
    return side[i];
  }
  int join(cudaStream_t main_st) {
    for (int i = 0; i < n_side; ++i)
      if (side_dirty[i]) {
        cudaEventRecord(ev_join[i], side[i]);
        cudaStreamWaitEvent(main_st, ev_join[i], 0);
        side_dirty[i] = false;
      }
    side_rr = 0;
    return SSHSLIE_OK;
  }
  float* partial_for(cudaStream_t st) const {
    for (int i = 1; i < n_side; ++i)
      if (side[i] && st == side[i]) return wg_partial[i];
    return wg_partial[0];
  }
  float* bias_partial_for(cudaStream_t st) const {
    for (int i = 1; i < n_side; ++i)
      if (side[i] && st == side[i]) return bias_partials[i];
    return bias_partials[0];
  }

  // per-call state read by the recorded launches
  const float* x = nullptr;
  float *out_R = nullptr, *out_I = nullptr, *out_Id = nullptr, *out_S = nullptr;   // caller's output buffers (may be null)
  const float* params = nullptr;
  float* grads = nullptr;
  float* losses = nullptr;
  sshslie_loss_cfg cfg;

  typedef std::function<int(cudaStream_t)> OpFn;
  std::vector<OpFn> ops_fwd, ops_loss_bwd2_illum, ops_bwd1;

  // persistent tensors needed by the ABI
  float *R32 = nullptr, *I32 = nullptr, *Id32 = nullptr, *S32 = nullptr;

  int add_geom(const ConvGeom& g) {
    geoms.push_back(g);
    geom_umma.push_back(0);
    geom_layer.push_back(last_layer);
    geom_role.push_back(last_role);
    return (int)geoms.size() - 1;
  }
};

static bool s2_halo_enabled() {
  const char* v = getenv("SSHSLIE_S2_HALO");
  return !(v && v[0] == '0');
}
static ConvGeom geom_init(int B, int OH, int OW, int N, const WAddr& wa) {
  ConvGeom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.OH = OH; g.OW = OW;
  g.dup_c = -1;
  g.N = N; g.Npad = (N + 15) / 16 * 16;
  g.w_off = wa.w_off; g.w_sN = wa.sN; g.w_sC = wa.sC;
  // tcgen05 tile: th x tw output pixels = 128 GEMM rows
  int tw = 128;
  while (tw > 8 && (OW % tw) != 0) tw >>= 1;
  g.tw = tw; g.th = 128 / tw;
  return g;
}
static void geom_add_slabs(ConvGeom& g, int src, int dh, int dw, const SrcSpec& s, int kh, int kw, const WAddr& wa) {
  const int ns = (s.c_cnt + SS_SLAB - 1) / SS_SLAB;
  for (int i = 0; i < ns; ++i) {
    Slab& sl = g.slab[g.nslabs++];
    sl.src = (int8_t)src; sl.dh = (int8_t)dh; sl.dw = (int8_t)dw; sl.no_wgrad = (int8_t)s.lo;
    sl.c0 = (int16_t)(s.c_off + i * SS_SLAB);
    const int rem = s.c_cnt - i * SS_SLAB;
    sl.wcn = (int16_t)(rem < SS_SLAB ? rem : SS_SLAB);
    sl.woff = kh * wa.sKH + kw * wa.sKW + (s.wc_off + i * SS_SLAB) * wa.sC;
  }
}
static SrcView view_full(const Tens& t) {
  SrcView v;
  v.base = t.p; v.sB = (int64_t)t.H * t.W * t.ld; v.sH = (int64_t)t.W * t.ld; v.sW = t.ld; v.H = t.H; v.W = t.W;
  return v;
}
static SrcView view_parity(const Tens& t, int ph, int pw) {
  SrcView v;
  v.base = t.p + (int64_t)ph * t.W * t.ld + (int64_t)pw * t.ld;
  v.sB = (int64_t)t.H * t.W * t.ld; v.sH = 2LL * t.W * t.ld; v.sW = 2LL * t.ld;
  v.H = (t.H - ph + 1) / 2; v.W = (t.W - pw + 1) / 2;
  return v;
}

// gather geometry of a stride-1 / stride-2 "conv-like" read:  in_coord = out_coord*stride + sign*(k_idx) + off
//   conv forward:         sign=+1, off=-pad          (stride 1 or 2)
//   dgrad of s1 conv:     sign=-1, off=+pad          (stride 1)
static ConvGeom geom_conv(int B, int OH, int OW, const std::vector<SrcSpec>& srcs, int k, int stride, int pad, int sign,
                          int N, const WAddr& wa) {
  ConvGeom g = geom_init(B, OH, OW, N, wa);
  if (stride == 1) {
    g.halo_ok = 1;
    for (size_t s = 0; s < srcs.size(); ++s) g.src[g.nsrc++] = view_full(srcs[s].t);
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw)
        for (size_t s = 0; s < srcs.size(); ++s)
          geom_add_slabs(g, (int)s, sign * kh + (sign > 0 ? -pad : pad), sign * kw + (sign > 0 ? -pad : pad), srcs[s],
                         kh, kw, wa);
  } else {
    // stride 2 forward-type read: in = 2*out + k_idx - pad  ->  parity views of the single source.  Inside a view the taps
    // are a stride-1 stencil again (offsets -1 / 0), so the halo-reuse gather kernels apply with one window per view.
    g.halo_ok = s2_halo_enabled() ? 2 : 0;
    // (a bf16 pair = two sources = 2 x 4 views, the residual one flagged `lo`)
    for (size_t si = 0; si < srcs.size(); ++si)
      for (int ph = 0; ph < 2; ++ph)
        for (int pw = 0; pw < 2; ++pw) g.src[g.nsrc++] = view_parity(srcs[si].t, ph, pw);
    for (int kh = 0; kh < k; ++kh)
      for (int kw = 0; kw < k; ++kw) {
        const int eh = kh - pad, ew = kw - pad;
        const int ph = ((eh % 2) + 2) % 2, pw = ((ew % 2) + 2) % 2;
        for (size_t si = 0; si < srcs.size(); ++si)
          geom_add_slabs(g, (int)si * 4 + ph * 2 + pw, (eh - ph) / 2, (ew - pw) / 2, srcs[si], kh, kw, wa);
      }
  }
  return g;
}
// transposed (stride-2) gather for output parity class (qh,qw): in = out_class_coord + (q + pad - k_idx)/2
static ConvGeom geom_tconv_class(int B, int OHc, int OWc, const SrcSpec& s, int k, int pad, int qh, int qw, int N,
                                 const WAddr& wa) {
  ConvGeom g = geom_init(B, OHc, OWc, N, wa);
  g.halo_ok = s2_halo_enabled() ? 2 : 0;      // offsets 0 / +1 of a full view: a stride-1 stencil
  g.src[g.nsrc++] = view_full(s.t);
  for (int kh = 0; kh < k; ++kh) {
    if ((qh + pad - kh) & 1) continue;
    for (int kw = 0; kw < k; ++kw) {
      if ((qw + pad - kw) & 1) continue;
      geom_add_slabs(g, 0, (qh + pad - kh) / 2, (qw + pad - kw) / 2, s, kh, kw, wa);
    }
  }
  return g;
}

static Epi epi_bf16(const Tens& out, int n_store, int qh = 0, int qw = 0, int scale = 1) {
  Epi e;
  memset(&e, 0, sizeof(e));
  e.mode = EPI_BF16;
  e.out = out.p + (int64_t)qh * out.W * out.ld + (int64_t)qw * out.ld;
  e.oB = (int64_t)out.H * out.W * out.ld; e.oH = (int64_t)scale * out.W * out.ld; e.oW = (int64_t)scale * out.ld;
  e.n_store = n_store;
  return e;
}
static void epi_set_add(Epi& e, const Tens& t, int c_off = 0, int qh = 0, int qw = 0, int scale = 1) {
  e.add = t.p + c_off + (int64_t)qh * t.W * t.ld + (int64_t)qw * t.ld;
  e.aB = (int64_t)t.H * t.W * t.ld; e.aH = (int64_t)scale * t.W * t.ld; e.aW = (int64_t)scale * t.ld;
}
static void epi_set_split(Epi& e, int n_split, const Tens& out2, int n_store2, const Tens* mask2) {
  e.n_split = n_split; e.n_store2 = n_store2;
  e.out2 = out2.p;
  e.o2B = (int64_t)out2.H * out2.W * out2.ld; e.o2H = (int64_t)out2.W * out2.ld; e.o2W = out2.ld;
  if (mask2) {
    e.mask2 = mask2->p;
    e.m2B = (int64_t)mask2->H * mask2->W * mask2->ld; e.m2H = (int64_t)mask2->W * mask2->ld; e.m2W = mask2->ld;
  }
}
static void epi_set_mask(Epi& e, const Tens& t, int qh = 0, int qw = 0, int scale = 1) {
  e.mask = t.p + (int64_t)qh * t.W * t.ld + (int64_t)qw * t.ld;
  e.mB = (int64_t)t.H * t.W * t.ld; e.mH = (int64_t)scale * t.W * t.ld; e.mW = (int64_t)scale * t.ld;
}

// ---------------------------------------------------------------------------------------------
// thin kernels for final_conv (64 -> 1, model.py:141,174): its data- and weight-gradient are not GEMM shaped
// ---------------------------------------------------------------------------------------------
// dff[pix, c] = sum_taps dId[pix - tap] * w[c][tap]
__global__ void __launch_bounds__(256) final_dgrad_kernel(const float* __restrict__ dId, const float* __restrict__ w,
                                                          bf16* __restrict__ dff, int H, int W, int64_t total) {
  SS_PDL_ENTRY();
  __shared__ __align__(16) float ws[9 * 64];            // [tap][c]: a thread reads its 8 channels of a tap as two float4
  for (int i = threadIdx.x; i < 576; i += 256) ws[(i % 9) * 64 + i / 9] = w[i];
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i & 7);
  const int64_t pix = i >> 3;
  const int x = (int)(pix % W);
  const int64_t t = pix / W;
  const int y = (int)(t % H);
  const int64_t b = t / H;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
  for (int kh = 0; kh < 3; ++kh)
#pragma unroll
    for (int kw = 0; kw < 3; ++kw) {
      const int oy = y - kh + 1, ox = x - kw + 1;
      if (oy < 0 || oy >= H || ox < 0 || ox >= W) continue;
      const float g = dId[(b * H + oy) * W + ox];
      const float4 w0 = *reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * 64 + q * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + (kh * 3 + kw) * 64 + q * 8 + 4);
      acc[0] = fmaf(g, w0.x, acc[0]); acc[1] = fmaf(g, w0.y, acc[1]); acc[2] = fmaf(g, w0.z, acc[2]); acc[3] = fmaf(g, w0.w, acc[3]);
      acc[4] = fmaf(g, w1.x, acc[4]); acc[5] = fmaf(g, w1.y, acc[5]); acc[6] = fmaf(g, w1.z, acc[6]); acc[7] = fmaf(g, w1.w, acc[7]);
    }
  uint4 o;
  o.x = pack2(acc[0], acc[1]); o.y = pack2(acc[2], acc[3]); o.z = pack2(acc[4], acc[5]); o.w = pack2(acc[6], acc[7]);
  reinterpret_cast<uint4*>(dff)[i] = o;
}
// dw[c][tap] += sum_pix dId[pix] * ff[pix + tap, c] ; db += sum dId.   block = 576 threads (c = t%64, tap = t/64).
// Block blk walks the image rows blk, blk + gridDim.x, ... in segments of FW_XS pixels: the three ff rows around a segment
// (with a zero column either side) and the dId values are staged in shared memory with 16-byte loads, then every
// (channel, tap) thread runs over the segment from there.  (The first version read both operands straight from global
// memory, one dependent scalar load pair per pixel: 271 us at B=32 for 134 MB of input.)
// Row blk of `partials` ([gridDim.x][577]: 576 weights, then the bias) receives the block's sums; ss_launch_reduce_rows
// adds the rows in a fixed order (deterministic, no atomics).
#define FINAL_WGRAD_MAX_BLOCKS 592
#define FW_XS 64
__global__ void __launch_bounds__(576) final_wgrad_kernel(const float* __restrict__ dId, const bf16* __restrict__ ff,
                                                          float* __restrict__ partials, int B, int H, int W) {
  __shared__ __align__(16) bf16 ffs[3][FW_XS + 2][64];
  __shared__ float gs[FW_XS];
  const int c = threadIdx.x & 63, tap = threadIdx.x >> 6;
  const int kh = tap / 3, kw = tap - kh * 3;
  float acc = 0.f, accb = 0.f;
  for (int64_t r = blockIdx.x; r < (int64_t)B * H; r += gridDim.x) {      // image rows over B*H
    const int y = (int)(r % H);
    const int64_t b = r / H;
    for (int x0 = 0; x0 < W; x0 += FW_XS) {
      const int nx = min(FW_XS, W - x0);
      __syncthreads();                                                    // the previous segment has been consumed
      for (int i = threadIdx.x; i < 3 * (FW_XS + 2) * 8; i += 576) {      // 16-byte vectors: (row, pixel -1..FW_XS, 8 channels)
        const int v = i & 7, px = (i >> 3) % (FW_XS + 2), rr = (i >> 3) / (FW_XS + 2);
        const int iy = y + rr - 1, ix = x0 + px - 1;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W && px <= nx + 1)
          u = __ldg(reinterpret_cast<const uint4*>(ff + ((b * H + iy) * W + ix) * 64) + v);
        *reinterpret_cast<uint4*>(&ffs[rr][px][v * 8]) = u;
      }
      for (int i = threadIdx.x; i < FW_XS; i += 576) gs[i] = (i < nx) ? dId[r * W + x0 + i] : 0.f;
      __syncthreads();
#pragma unroll 8
      for (int x = 0; x < nx; ++x) acc = fmaf(gs[x], bf2f(ffs[kh][x + kw][c]), acc);
      if (threadIdx.x < 32) {                                             // bias gradient: warp 0 adds the segment
        float v = 0.f;
        for (int x = threadIdx.x; x < FW_XS; x += 32) v += gs[x];
        v = warp_sum(v);
        accb += v;
      }
    }
  }
  partials[(size_t)blockIdx.x * 577 + c * 9 + tap] = acc;
  if (threadIdx.x == 0) partials[(size_t)blockIdx.x * 577 + 576] = accb;
}

// ---------------------------------------------------------------------------------------------
// plan construction
// ---------------------------------------------------------------------------------------------
struct DecompBufs { Tens c0, sh, c1, c2, c3, dc, c5, c7; };

static double geom_flops(const ConvGeom& g) {
  double k = 0;
  for (int i = 0; i < g.nslabs; ++i) k += g.slab[i].wcn;
  return 2.0 * g.B * g.OH * g.OW * (double)g.N * k;
}
static std::string geom_label(const sshslie_engine* e, int gi, const char* kind) {
  const int l = gi < (int)e->geom_layer.size() ? e->geom_layer[gi] : -1;
  std::string s = std::string(kind) + ":" + (l >= 0 && l < L_COUNT ? kLayerNames[l] : "layer");
  const bool pipe = kind[0] != 'w' && e->geom_umma[gi] == 2 && gi < (int)e->pipe_use.size() && e->pipe_use[gi] == 1;
  s += pipe ? "[tcgen05-pipe]" : (e->geom_umma[gi] == 2 ? "[tcgen05-halo]" : (e->geom_umma[gi] ? "[tcgen05]" : "[simt]"));
  return s;
}
static int run_gather(sshslie_engine* e, int gi, Epi epi, int bias_layer, cudaStream_t st) {
  if (bias_layer >= 0) epi.bias = e->params + e->poff[2 * bias_layer + 1];
  const ConvGeom& g = e->geoms[gi];
  if (e->geom_umma[gi] == 2 && e->pipe_on) {
    if (e->pipe_use.size() < e->geoms.size()) {
      e->pipe_use.resize(e->geoms.size(), 0);
      e->pipe_valid.resize(e->geoms.size(), 0);
      e->pipe_blob.resize(e->geoms.size() * ss_pipe_plan_size(), 0);
    }
    if (e->pipe_use[gi] == 0) {
      const int n_tiles = g.B * ((g.OH + 15) / 16) * ((g.OW + 7) / 8);
      const int min_tiles = (epi.mode == EPI_HEAD) ? e->pipe_min_tiles_head : e->pipe_min_tiles;
      e->pipe_use[gi] = (n_tiles >= min_tiles && g.nslabs <= e->pipe_max_slabs && ss_umma_pipe_supported(g, epi)) ? 1 : 2;
    }
  }
  // stride-2 / transposed geoms below the tile threshold stay on the per-tap kernel (fewer, smaller windows per launch)
  const bool s2_small = g.halo_ok == 2 && g.B * ((g.OH + 15) / 16) * ((g.OW + 7) / 8) < e->s2_min_tiles;
  if (e->geom_umma[gi] == 2 && s2_small) {
    prof_note(std::string(e->geom_role[gi] ? "dgrad:" : "fwd:") +
                  (e->geom_layer[gi] >= 0 && e->geom_layer[gi] < L_COUNT ? kLayerNames[e->geom_layer[gi]] : "layer") + "[tcgen05]",
              geom_flops(g), 0);
    return ss_launch_conv_gather_umma(e->geoms_dev + gi, g,
                                      *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                      epi, st);
  }
  prof_note(geom_label(e, gi, e->geom_role[gi] ? "dgrad" : "fwd"), geom_flops(g), 0);
  if (e->geom_umma[gi] == 2 && e->pipe_on && e->pipe_use[gi] == 1)
    return ss_launch_conv_gather_pipe(g, *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                      epi, e->pipe_blob.data() + gi * ss_pipe_plan_size(), &e->pipe_valid[gi], st);
  if (e->geom_umma[gi] == 2)
    return ss_launch_conv_gather_halo(e->geoms_dev + gi, g,
                                      *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                      epi, st);
  if (e->geom_umma[gi])
    return ss_launch_conv_gather_umma(e->geoms_dev + gi, g,
                                      *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                      epi, st);
  return ss_launch_conv_gather_simt(e->geoms_dev + gi, g, epi, st);
}
// four parity classes (consecutive geoms gi0..gi0+3, same tile grid, same Npad) in one launch when all are on tcgen05
static int run_gather4(sshslie_engine* e, int gi0, const Epi* epis, int bias_layer, cudaStream_t st) {
  // one per-tap launch over the four classes (blockIdx.y = class) for small problems; above s2_min_tiles per class each
  // class goes through the halo-reuse kernels on its own (4 launches, each at 3-5x the per-tap kernel's rate)
  const ConvGeom& g0 = e->geoms[gi0];
  const int class_tiles = g0.B * ((g0.OH + 15) / 16) * ((g0.OW + 7) / 8);
  bool merged = true, halo4 = class_tiles >= e->s2_min_tiles;
  for (int q = 0; q < 4; ++q) {
    merged = merged && e->geom_umma[gi0 + q] >= 1 && e->geoms[gi0 + q].Npad == e->geoms[gi0].Npad;
    halo4 = halo4 && e->geom_umma[gi0 + q] == 2;
  }
  merged = merged && !halo4;
  if (!merged) {
    for (int q = 0; q < 4; ++q) {
      const int rc = run_gather(e, gi0 + q, epis[q], bias_layer, st);
      if (rc) return rc;
    }
    return SSHSLIE_OK;
  }
  Epi epi = epis[0];
  if (bias_layer >= 0) epi.bias = e->params + e->poff[2 * bias_layer + 1];
  double fl = 0;
  for (int q = 0; q < 4; ++q) fl += geom_flops(e->geoms[gi0 + q]);
  prof_note(std::string(e->geom_role[gi0] ? "dgrad:" : "fwd:") +
                (e->geom_layer[gi0] >= 0 && e->geom_layer[gi0] < L_COUNT ? kLayerNames[e->geom_layer[gi0]] : "layer") +
                "[tcgen05]x4", fl, 0);
  return ss_launch_conv_gather_umma4(e->geoms_dev + gi0, &e->geoms[gi0],
                                     reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi0 * ss_umma_maps_size()),
                                     epi, st);
}

static int run_wgrad(sshslie_engine* e, int gi, const Tens& G, int gN, int qh, int qw, int scale, cudaStream_t st,
                     int bias_layer = -1) {
  if (e->skip_wgrad) return SSHSLIE_OK;      // timing experiments only (SSHSLIE_SKIP_WGRAD=1): gradients are wrong
  const ConvGeom& g = e->geoms[gi];
  const bf16* gp = G.p + (int64_t)qh * G.W * G.ld + (int64_t)qw * G.ld;
  const int64_t gB = (int64_t)G.H * G.W * G.ld, gH = (int64_t)scale * G.W * G.ld, gW = (int64_t)scale * G.ld;
  {
    std::string lbl = geom_label(e, gi, "wgrad");
    if (!(e->geom_umma[gi] && ss_umma_wgrad_supported(g))) lbl.replace(lbl.find('['), std::string::npos, "[simt]");
    prof_note(lbl, geom_flops(g) * (double)gN / (double)g.N, 0);
  }
  if (e->geom_umma[gi] == 2 && e->wgrad_halo && ss_umma_wgrad_halo_supported(g, gN)) {
    const std::pair<int, const void*> key(gi | 0x10000, (const void*)gp);
    auto it = e->gmaps.find(key);
    if (it == e->gmaps.end()) {
      GMapBox box;
      const int rc = ss_umma_build_gmap_halo(gp, gB, gH, gW, G.ld, g, box.bytes);
      if (rc) return rc;
      it = e->gmaps.emplace(key, box).first;
    }
    return ss_launch_conv_wgrad_halo(e->geoms_dev + gi, g,
                                     *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                     it->second.bytes, gN,
                                     bias_layer >= 0 ? (long long)e->poff[2 * bias_layer + 1] : -1LL,
                                     e->partial_for(st), e->grads, st);
  }
  if (e->geom_umma[gi] && ss_umma_wgrad_supported(g)) {
    // TMA descriptor of the G tensor, built once per (geom, tensor) and cached on the host
    const std::pair<int, const void*> key(gi, (const void*)gp);
    auto it = e->gmaps.find(key);
    if (it == e->gmaps.end()) {
      GMapBox box;
      const int rc = ss_umma_build_gmap(gp, gB, gH, gW, G.ld, g, box.bytes);
      if (rc) return rc;
      it = e->gmaps.emplace(key, box).first;
    }
    return ss_launch_conv_wgrad_umma(e->geoms_dev + gi, g,
                                     *reinterpret_cast<const UmmaMaps*>(e->maps_blob.data() + gi * ss_umma_maps_size()),
                                     it->second.bytes, gN,
                                     bias_layer >= 0 ? (long long)e->poff[2 * bias_layer + 1] : -1LL,
                                     e->partial_for(st), e->grads, st);
  }
  if (bias_layer >= 0) {
    const int rc = ss_launch_bias_grad(G.p, G.pix(), G.ld, gN, e->grads + e->poff[2 * bias_layer + 1],
                                       e->bias_partial_for(st), st);
    if (rc) return rc;
  }
  return ss_launch_conv_wgrad_simt(e->geoms_dev + gi, g, gp, gB, gH, gW, gN, e->grads, st);
}

static WAddr waddr_conv_fwd(const sshslie_engine* e, int layer, int n_off = 0) {
  e->last_layer = layer; e->last_role = 0;
  const LayerShape& s = e->shapes[layer];
  const int kk = s.k * s.k;
  WAddr w;
  if (!s.transposed) { w.sN = s.d1 * kk; w.sC = kk; }       // W[co][ci][kh][kw], n = co, inner = ci
  else { w.sN = kk; w.sC = s.d1 * kk; }                     // Wt[i][o][kh][kw], n = o,  inner = i
  w.sKH = s.k; w.sKW = 1;
  w.w_off = e->poff[2 * layer] + (int64_t)n_off * w.sN;
  return w;
}
static WAddr waddr_conv_dgrad(const sshslie_engine* e, int layer, int n_off = 0) {
  e->last_layer = layer; e->last_role = 1;
  const LayerShape& s = e->shapes[layer];
  const int kk = s.k * s.k;
  WAddr w;
  if (!s.transposed) { w.sN = kk; w.sC = s.d1 * kk; }       // n = ci, inner = co
  else { w.sN = s.d1 * kk; w.sC = kk; }                     // n = i,  inner = o
  w.sKH = s.k; w.sKW = 1;
  w.w_off = e->poff[2 * layer] + (int64_t)n_off * w.sN;
  return w;
}

#define PUSH(vec, ...) (vec).push_back([=](cudaStream_t st) -> int { __VA_ARGS__ })
// weight-gradient work is off the critical path of backward (nothing reads the gradients before the step ends):
// it is forked onto the engine's side stream and joined at the end of each phase (captured as graph edges)
#define PUSH_SIDE(vec, ...) \
  (vec).push_back([=](cudaStream_t main_st) -> int { cudaStream_t st = e->fork(main_st); __VA_ARGS__ })
#define PUSH_JOIN(vec) (vec).push_back([=](cudaStream_t st) -> int { return e->join(st); })

// a layer's weight gradient, forked onto a side stream
static void queue_wgrad(sshslie_engine* e, std::vector<sshslie_engine::OpFn>& ops, int gi, const Tens& G, int gN,
                        int bias_layer) {
  PUSH_SIDE(ops, return run_wgrad(e, gi, G, gN, 0, 0, 1, st, bias_layer););
}

// forward of one DecompositionNet pass (model.py:49-70).  Returns the geom ids for reuse by the backward pass.
struct DecompGeoms {
  int conv0, shallow, conv1, conv2, conv3, deconv[4], conv5, conv7, recon;
};
static DecompGeoms plan_decomp_fwd(sshslie_engine* e, std::vector<sshslie_engine::OpFn>& ops, const Tens& in,
                                   DecompBufs& d, Epi head_epi, bool join_pack = false) {
  const int B = e->B, H = e->H, W = e->W;
  DecompGeoms G;
  {  // conv0: 64 -> 32, ReLU
    WAddr wa = waddr_conv_fwd(e, L_D_CONV0);
    G.conv0 = e->add_geom(geom_conv(B, H, W, {{in, 0, e->C, 0}}, 3, 1, 1, +1, 32, wa));
    Epi ep = epi_bf16(d.c0, 32); ep.relu = 1;
    const int gi = G.conv0;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV0, st););
  }
  {  // shallow 9x9: 64 -> 64, no activation
    WAddr wa = waddr_conv_fwd(e, L_D_SHALLOW);
    G.shallow = e->add_geom(geom_conv(B, H, W, {{in, 0, e->C, 0}}, 9, 1, 4, +1, 64, wa));
    Epi ep = epi_bf16(d.sh, 64);
    const int gi = G.shallow;
    PUSH(ops, return run_gather(e, gi, ep, L_D_SHALLOW, st););
    if (join_pack) PUSH_JOIN(ops);        // the rest of the packed weights (side stream) is needed from here on
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_D_CONV1);
    G.conv1 = e->add_geom(geom_conv(B, H, W, {{d.sh, 0, 64, 0}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(d.c1, 64); ep.relu = 1;
    const int gi = G.conv1;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV1, st););
  }
  {  // conv2: stride 2, 64 -> 128
    WAddr wa = waddr_conv_fwd(e, L_D_CONV2);
    G.conv2 = e->add_geom(geom_conv(B, H / 2, W / 2, {{d.c1, 0, 64, 0}}, 3, 2, 1, +1, 128, wa));
    Epi ep = epi_bf16(d.c2, 128); ep.relu = 1;
    const int gi = G.conv2;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV2, st););
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_D_CONV3);
    G.conv3 = e->add_geom(geom_conv(B, H / 2, W / 2, {{d.c2, 0, 128, 0}}, 3, 1, 1, +1, 128, wa));
    Epi ep = epi_bf16(d.c3, 128); ep.relu = 1;
    const int gi = G.conv3;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV3, st););
  }
  {  // ConvTranspose 3x3 s2 p1 op1: 128 -> 64, four output parity classes, one launch
    Epi eps[4];
    for (int q = 0; q < 4; ++q) {
      const int qh = q >> 1, qw = q & 1;
      WAddr wa = waddr_conv_fwd(e, L_D_DECONV);
      G.deconv[q] = e->add_geom(geom_tconv_class(B, H / 2, W / 2, {d.c3, 0, 128, 0}, 3, 1, qh, qw, 64, wa));
      eps[q] = epi_bf16(d.dc, 64, qh, qw, 2); eps[q].relu = 1;
    }
    const int gi0 = G.deconv[0];
    const std::array<Epi, 4> ea = {eps[0], eps[1], eps[2], eps[3]};
    PUSH(ops, return run_gather4(e, gi0, ea.data(), L_D_DECONV, st););
  }
  {  // conv5 on cat[deconv, conv1]
    WAddr wa = waddr_conv_fwd(e, L_D_CONV5);
    G.conv5 = e->add_geom(geom_conv(B, H, W, {{d.dc, 0, 64, 0}, {d.c1, 0, 64, 64}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(d.c5, 64); ep.relu = 1;
    const int gi = G.conv5;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV5, st););
  }
  {  // conv7 on cat[conv5, conv0(32)]
    WAddr wa = waddr_conv_fwd(e, L_D_CONV7);
    G.conv7 = e->add_geom(geom_conv(B, H, W, {{d.c5, 0, 64, 0}, {d.c0, 0, 32, 64}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(d.c7, 64);
    const int gi = G.conv7;
    PUSH(ops, return run_gather(e, gi, ep, L_D_CONV7, st););
  }
  {  // recon: 64 -> C+1, sigmoid heads
    WAddr wa = waddr_conv_fwd(e, L_D_RECON);
    G.recon = e->add_geom(geom_conv(B, H, W, {{d.c7, 0, 64, 0}}, 3, 1, 1, +1, e->C + 1, wa));
    const int gi = G.recon;
    PUSH(ops, return run_gather(e, gi, head_epi, L_D_RECON, st););
  }
  return G;
}

// backward of one DecompositionNet pass, given dc8 (gradient wrt the recon pre-activation, `dc8_n` valid columns)
struct DecompGrads { Tens dc7, dc5, dc0, ddc, dc1p, dc3, dc2, dc1, dsh, din; };
static void plan_decomp_bwd(sshslie_engine* e, std::vector<sshslie_engine::OpFn>& ops, const Tens& in,
                            const DecompBufs& d, const DecompGeoms& G, const Tens& dc8, int dc8_n, DecompGrads& g,
                            bool need_din) {
  const int B = e->B, H = e->H, W = e->W;
  float** grads = &e->grads;
  const int64_t* poff = e->poff;
  auto bias_grad = [&](const Tens& t, int n, int layer) {
    PUSH_SIDE(ops, return ss_launch_bias_grad(t.p, t.pix(), t.ld, n, *grads + poff[2 * layer + 1], e->bias_partial_for(st), st););
  };
  // recon
  queue_wgrad(e, ops, G.recon, dc8, dc8_n, L_D_RECON);
  {
    WAddr wa = waddr_conv_dgrad(e, L_D_RECON);
    const int gi = e->add_geom(geom_conv(B, H, W, {{dc8, 0, dc8_n, 0}}, 3, 1, 1, -1, 64, wa));
    Epi ep = epi_bf16(g.dc7, 64);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // conv7 (no activation): inputs [c5 | c0]
  queue_wgrad(e, ops, G.conv7, g.dc7, 64, L_D_CONV7);
  {  // both halves of the concat [c5 (64) | c0 (32)] from ONE GEMM (N = 96): columns 64.. go to dc0 with c0's ReLU mask
    WAddr wa = waddr_conv_dgrad(e, L_D_CONV7, 0);
    const int gi = e->add_geom(geom_conv(B, H, W, {{g.dc7, 0, 64, 0}}, 3, 1, 1, -1, 96, wa));
    Epi ep = epi_bf16(g.dc5, 64); epi_set_mask(ep, d.c5);
    epi_set_split(ep, 64, g.dc0, 32, &d.c0);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // conv5 (ReLU already folded into dc5): inputs [dc | c1]
  queue_wgrad(e, ops, G.conv5, g.dc5, 64, L_D_CONV5);
  {  // concat [deconv (64) | conv1 (64)]: ONE GEMM with N = 128, columns 64.. are the skip gradient dc1p (no mask)
    WAddr wa = waddr_conv_dgrad(e, L_D_CONV5, 0);
    const int gi = e->add_geom(geom_conv(B, H, W, {{g.dc5, 0, 64, 0}}, 3, 1, 1, -1, 128, wa));
    Epi ep = epi_bf16(g.ddc, 64); epi_set_mask(ep, d.dc);
    epi_set_split(ep, 64, g.dc1p, 64, nullptr);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // deconv: dgrad = stride-2 conv of ddc (n = 128 input channels); the same geom drives its wgrad with G = c3
  {
    WAddr wa = waddr_conv_dgrad(e, L_D_DECONV);
    const int gi = e->add_geom(geom_conv(B, H / 2, W / 2, {{g.ddc, 0, 64, 0}}, 3, 2, 1, +1, 128, wa));
    const Tens c3 = d.c3;
    queue_wgrad(e, ops, gi, c3, 128, -1);
    bias_grad(g.ddc, 64, L_D_DECONV);
    Epi ep = epi_bf16(g.dc3, 128); epi_set_mask(ep, d.c3);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // conv3
  queue_wgrad(e, ops, G.conv3, g.dc3, 128, L_D_CONV3);
  {
    WAddr wa = waddr_conv_dgrad(e, L_D_CONV3);
    const int gi = e->add_geom(geom_conv(B, H / 2, W / 2, {{g.dc3, 0, 128, 0}}, 3, 1, 1, -1, 128, wa));
    Epi ep = epi_bf16(g.dc2, 128); epi_set_mask(ep, d.c2);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // conv2 (stride 2): wgrad on its forward geom; dgrad = transposed gather per input parity class, + dc1p, ReLU mask
  queue_wgrad(e, ops, G.conv2, g.dc2, 128, L_D_CONV2);
  {
    Epi eps[4];
    int gi0 = -1;
    for (int q = 0; q < 4; ++q) {
      const int qh = q >> 1, qw = q & 1;
      WAddr wa = waddr_conv_dgrad(e, L_D_CONV2);
      const int gi = e->add_geom(geom_tconv_class(B, H / 2, W / 2, {g.dc2, 0, 128, 0}, 3, 1, qh, qw, 64, wa));
      if (q == 0) gi0 = gi;
      eps[q] = epi_bf16(g.dc1, 64, qh, qw, 2);
      epi_set_add(eps[q], g.dc1p, 0, qh, qw, 2);
      epi_set_mask(eps[q], d.c1, qh, qw, 2);
    }
    const std::array<Epi, 4> ea = {eps[0], eps[1], eps[2], eps[3]};
    PUSH(ops, return run_gather4(e, gi0, ea.data(), -1, st););
  }
  // conv1
  queue_wgrad(e, ops, G.conv1, g.dc1, 64, L_D_CONV1);
  {
    WAddr wa = waddr_conv_dgrad(e, L_D_CONV1);
    const int gi = e->add_geom(geom_conv(B, H, W, {{g.dc1, 0, 64, 0}}, 3, 1, 1, -1, 64, wa));
    Epi ep = epi_bf16(g.dsh, 64);
    PUSH(ops, return run_gather(e, gi, ep, -1, st););
  }
  // shallow 9x9 and conv0 read the pass input
  queue_wgrad(e, ops, G.shallow, g.dsh, 64, L_D_SHALLOW);
  queue_wgrad(e, ops, G.conv0, g.dc0, 32, L_D_CONV0);
  if (need_din) {  // d(input) = dgrad_shallow(dsh) + dgrad_conv0(dc0)
    {
      WAddr wa = waddr_conv_dgrad(e, L_D_SHALLOW);
      const int gi = e->add_geom(geom_conv(B, H, W, {{g.dsh, 0, 64, 0}}, 9, 1, 4, -1, e->C, wa));
      Epi ep = epi_bf16(g.din, e->C);
      PUSH(ops, return run_gather(e, gi, ep, -1, st););
    }
    {
      WAddr wa = waddr_conv_dgrad(e, L_D_CONV0);
      const int gi = e->add_geom(geom_conv(B, H, W, {{g.dc0, 0, 32, 0}}, 3, 1, 1, -1, e->C, wa));
      Epi ep = epi_bf16(g.din, e->C); epi_set_add(ep, g.din);
      PUSH(ops, return run_gather(e, gi, ep, -1, st););
    }
  }
  (void)in;
}

static int copy_outputs(sshslie_engine* e, float* R, float* I, float* Id, float* S, cudaStream_t st);
static int build_plan(sshslie_engine* e, unsigned char* base) {
  e->base = base;
  e->cursor = 0;
  e->zero_ranges.clear();
  e->geoms.clear();
  e->geom_umma.clear();
  e->geom_layer.clear();
  e->geom_role.clear();
  e->ops_fwd.clear();
  e->ops_loss_bwd2_illum.clear();
  e->ops_bwd1.clear();
  const int B = e->B, C = e->C, H = e->H, W = e->W;
  const int64_t n = (int64_t)B * H * W;
  const bool train = e->train;
  // the illumination net's pyramid: a 3x3 stride-2 pad-1 conv maps n -> ceil(n / 2) (model.py:127-129); H, W are even
  const int h1 = (H + 1) / 2, w1 = (W + 1) / 2, h2 = (h1 + 1) / 2, w2 = (w1 + 1) / 2, h3 = (h2 + 1) / 2, w3 = (w2 + 1) / 2;

  // device copies of the plan
  e->geoms_dev = (ConvGeom*)e->alloc(sizeof(ConvGeom) * 128);
  e->pack_start_dev = (int*)e->alloc(sizeof(int) * 130);
  e->mask_dev = e->falloc((int64_t)H * W);
  if (train) {
    e->pix_rows = ss_pixel_losses_blocks(B, C, H, W);
    e->pix_partials = e->falloc(ss_pixel_losses_scratch_floats(B, C, H, W));
    e->four_partials = e->falloc((int64_t)B * C);
    e->final_partials = e->falloc((int64_t)FINAL_WGRAD_MAX_BLOCKS * 577);
    e->attn_partials = e->falloc((int64_t)SS_ATTN_WGRAD_MAX_BLOCKS * SS_ATTN_WGRAD_COLS);
    for (int i = 0; i < SS_MAX_SIDE; ++i) e->bias_partials[i] = e->falloc((int64_t)SS_BIAS_GRAD_MAX_BLOCKS * 256);
  }

  // ---- tensors -----------------------------------------------------------------------------
  Tens X = e->talloc(B, H, W, 64);
  e->R32 = e->falloc(n * C); e->I32 = e->falloc(n); e->Id32 = e->falloc(n); e->S32 = e->falloc(n * C);
  // cat[R, I] as the illumination net reads it: [R hi (64) | I, 63 zero lanes | R lo (64)]  (hi+lo: DESIGN.md §4)
  Tens RI = e->talloc(B, H, W, 192, true);
  auto decomp_bufs = [&]() {
    DecompBufs d;
    d.c0 = e->talloc(B, H, W, 64, true);
    d.sh = e->talloc(B, H, W, 64); d.c1 = e->talloc(B, H, W, 64);
    d.c2 = e->talloc(B, H / 2, W / 2, 128); d.c3 = e->talloc(B, H / 2, W / 2, 128);
    d.dc = e->talloc(B, H, W, 64); d.c5 = e->talloc(B, H, W, 64); d.c7 = e->talloc(B, H, W, 64);
    return d;
  };
  DecompBufs d1 = decomp_bufs();
  Tens a0 = e->talloc(B, H, W, 64), a1 = e->talloc(B, h1, w1, 64), a2 = e->talloc(B, h2, w2, 64),
       a3 = e->talloc(B, h3, w3, 64), tt = e->talloc(B, h3, w3, 64);
  Tens a0l = e->talloc(B, H, W, 64), r3l = e->talloc(B, H, W, 64), ffl = e->talloc(B, H, W, 64);   // bf16 residuals
  // the half-resolution skip d2 = deconv2 + conv1 (model.py:161) feeds deconv3 and the concat at full resolution: its
  // rounding noise reaches I_delta directly, so conv1's output keeps its residual for the skip (a1l) and the up-sampled
  // sum is stored as a bf16 pair (u3 | u3l) - DESIGN.md 4
  Tens a1l = e->talloc(B, h1, w1, 64), u3l = e->talloc(B, H, W, 64);
  Tens u1 = e->talloc(B, h2, w2, 64), r1 = e->talloc(B, h2, w2, 64), u2 = e->talloc(B, h1, w1, 64),
       r2 = e->talloc(B, h1, w1, 64), u3 = e->talloc(B, H, W, 64), r3 = e->talloc(B, H, W, 64);
  Tens fg = e->talloc(B, H, W, 320), ff = e->talloc(B, H, W, 64);
  const int L = h3 * w3;
  const int64_t T64 = (int64_t)B * L * 64;
  AttnBuffers ab;
  memset(&ab, 0, sizeof(ab));
  ab.x = e->falloc(T64); ab.q = e->falloc(T64); ab.k = e->falloc(T64); ab.v = e->falloc(T64); ab.o = e->falloc(T64);
  ab.lse = e->falloc((int64_t)B * 4 * L); ab.h = e->falloc(T64); ab.t32 = e->falloc(T64);
  if (L >= ss_attn_tc_min_l()) {       // per-head bf16 operands of the tensor-core attention core
    ab.qp = (bf16*)e->alloc(T64 * 4 * (int64_t)sizeof(bf16));
    ab.kvp = (bf16*)e->alloc(T64 * 4 * (int64_t)sizeof(bf16));
  }
  if (train) {
    ab.dq = e->falloc(T64); ab.dk = e->falloc(T64); ab.dv = e->falloc(T64); ab.d_o = e->falloc(T64);
    ab.dh = e->falloc(T64); ab.dx = e->falloc(T64); ab.Dv = e->falloc((int64_t)B * 4 * L);
  }

  // ---- forward (model.py:229-234) ----------------------------------------------------------
  auto& F = e->ops_fwd;
  if (train) {
    // per-step zeroing first: a memset node between two kernels would break their programmatic-dependent-launch edge
    PUSH(F, if (!e->grads) return SSHSLIE_OK;
            return cudaMemsetAsync(e->grads, 0, e->nparams * sizeof(float), st) == cudaSuccess ? 0 : SSHSLIE_ERR_CUDA;);
  }
  PUSH(F, return ss_launch_nchw32_to_nhwc16(e->x, X.p, B, C, H, W, 64, st););
  e->fwd_pack_begin = (int)F.size();
  // fp32 master weights -> packed bf16: the first two layers' weights (conv0, 9x9) on the caller's stream, everything
  // else on a side stream, joined after the 9x9 layer (pack_split = first pack block of the third geom)
  PUSH(F, return ss_launch_pack_weights(e->geoms_dev, e->pack_start_dev, (int)e->geoms.size(), e->pack_blocks,
                                        e->params, st, 0, e->pack_split););
  PUSH_SIDE(F, return ss_launch_pack_weights(e->geoms_dev, e->pack_start_dev, (int)e->geoms.size(), e->pack_blocks,
                                             e->params, st, e->pack_split, -1););
  Epi head1;
  memset(&head1, 0, sizeof(head1));
  head1.mode = EPI_HEAD; head1.R32 = e->R32; head1.I32 = e->I32; head1.RI = RI.p; head1.ri_c = 192;
  head1.ri_lo_off = 128;
  head1.C = C; head1.H = H; head1.W = W;
  DecompGeoms G1 = plan_decomp_fwd(e, F, X, d1, head1, true);

  // IllumAdjustmentNet (model.py:143-175)
  e->fwd_illum_begin = (int)F.size();
  e->RI_ptr = RI.p;
  int g_i0, g_i1, g_i2, g_i3, g_d1, g_d2, g_d3, g_fus, g_fin;
  {
    WAddr wa = waddr_conv_fwd(e, L_I_CONV0);
    ConvGeom gi0 = geom_conv(B, H, W, {{RI, 0, C + 1, 0}, {RI, 128, C, 0, 1}}, 3, 1, 1, +1, 64, wa);
    gi0.dup_c = C + 1;                    // lane C + 1 of cat[R, I] holds I's bf16 residual (written by the sigmoid head)
    g_i0 = e->add_geom(gi0);
    Epi ep = epi_bf16(a0, 64); ep.out_lo = a0l.p;
    PUSH(F, return run_gather(e, g_i0, ep, L_I_CONV0, st););
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_I_CONV1);
    g_i1 = e->add_geom(geom_conv(B, h1, w1, {{a0, 0, 64, 0}, {a0l, 0, 64, 0, 1}}, 3, 2, 1, +1, 64, wa));
    Epi ep = epi_bf16(a1, 64); ep.relu = 1; ep.out_lo = a1l.p;
    PUSH(F, return run_gather(e, g_i1, ep, L_I_CONV1, st););
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_I_CONV2);
    g_i2 = e->add_geom(geom_conv(B, h2, w2, {{a1, 0, 64, 0}}, 3, 2, 1, +1, 64, wa));
    Epi ep = epi_bf16(a2, 64); ep.relu = 1;
    PUSH(F, return run_gather(e, g_i2, ep, L_I_CONV2, st););
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_I_CONV3);
    g_i3 = e->add_geom(geom_conv(B, h3, w3, {{a2, 0, 64, 0}}, 3, 2, 1, +1, 64, wa));
    Epi ep = epi_bf16(a3, 64); ep.relu = 1;
    PUSH(F, return run_gather(e, g_i3, ep, L_I_CONV3, st););
  }
  const int64_t* apoff = e->poff + 2 * L_Q;
  PUSH(F, return ss_attention_forward(a3.p, tt.p, e->params, apoff, ab, B, L, st););
  PUSH(F, return ss_launch_upsample_add(tt.p, nullptr, u1.p, B, h3, w3, h2, w2, st););
  {
    WAddr wa = waddr_conv_fwd(e, L_I_DECONV1);
    g_d1 = e->add_geom(geom_conv(B, h2, w2, {{u1, 0, 64, 0}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(r1, 64); ep.relu = 1;
    PUSH(F, return run_gather(e, g_d1, ep, L_I_DECONV1, st););
  }
  PUSH(F, return ss_launch_upsample_add(r1.p, a2.p, u2.p, B, h2, w2, h1, w1, st););
  {
    WAddr wa = waddr_conv_fwd(e, L_I_DECONV2);
    g_d2 = e->add_geom(geom_conv(B, h1, w1, {{u2, 0, 64, 0}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(r2, 64); ep.relu = 1;
    PUSH(F, return run_gather(e, g_d2, ep, L_I_DECONV2, st););
  }
  PUSH(F, return ss_launch_upsample_add_pair(r2.p, a1.p, a1l.p, u3.p, u3l.p, B, h1, w1, H, W, st););
  {
    WAddr wa = waddr_conv_fwd(e, L_I_DECONV3);
    g_d3 = e->add_geom(geom_conv(B, H, W, {{u3, 0, 64, 0}, {u3l, 0, 64, 0, 1}}, 3, 1, 1, +1, 64, wa));
    Epi ep = epi_bf16(r3, 64); ep.relu = 1; ep.out_lo = r3l.p;
    PUSH(F, return run_gather(e, g_d3, ep, L_I_DECONV3, st););
  }
  PUSH(F, return ss_launch_fuse_concat(r1.p, a2.p, r2.p, a1.p, a1l.p, r3.p, r3l.p, a0.p, a0l.p, fg.p, B, H, W, h2, w2, h1, w1, st););
  {
    WAddr wa = waddr_conv_fwd(e, L_I_FUSION);
    g_fus = e->add_geom(geom_conv(B, H, W, {{fg, 0, 192, 0}, {fg, 192, 64, 128, 1}, {fg, 256, 64, 64, 1}}, 1, 1, 0, +1, 64, wa));
    Epi ep = epi_bf16(ff, 64); ep.out_lo = ffl.p;
    PUSH(F, return run_gather(e, g_fus, ep, L_I_FUSION, st););
  }
  {
    WAddr wa = waddr_conv_fwd(e, L_I_FINAL);
    g_fin = e->add_geom(geom_conv(B, H, W, {{ff, 0, 64, 0}, {ffl, 0, 64, 0, 1}}, 3, 1, 1, +1, 1, wa));
    Epi ep;
    memset(&ep, 0, sizeof(ep));
    ep.mode = EPI_PLANE32; ep.plane32 = e->Id32; ep.H = H; ep.W = W;
    PUSH(F, return run_gather(e, g_fin, ep, L_I_FINAL, st););
  }
  e->fwd_illum_end = (int)F.size();
  Tens Sb;
  if (train) Sb = e->talloc(B, H, W, 64);
  {
    bf16* sbp = train ? Sb.p : nullptr;
    PUSH(F, return ss_launch_make_s(e->R32, e->I32, e->Id32, e->S32, sbp, B, C, H, W, st););
  }

  // the caller's copies of R, I, I_delta, S: final once make_s has run -> copied on a side stream, off the critical path
  // (joined at the end of the forward for inference, at the end of phase 1 for training)
  PUSH_SIDE(F, return copy_outputs(e, e->out_R, e->out_I, e->out_Id, e->out_S, st););
  if (!train) PUSH_JOIN(F);
  if (train) {
    // ---- second decomposition pass on S, losses, backward ------------------------------------
    auto& Lq = e->ops_loss_bwd2_illum;
    float* Re32 = e->falloc(n * C);
    float *dR32 = e->falloc(n * C), *dI32 = e->falloc(n), *dId32 = e->falloc(n), *dS32 = e->falloc(n * C),
          *dRe32 = e->falloc(n * C);
    DecompBufs d2 = decomp_bufs();
    Epi head2;
    memset(&head2, 0, sizeof(head2));
    head2.mode = EPI_HEAD; head2.R32 = Re32; head2.C = C; head2.H = H; head2.W = W;
    // The Fourier term needs only x and S: it runs on a side stream BESIDE the second decomposition pass, writes its
    // gradient into a plane of its own (dSf32, summed in s_bwd) and is joined before the loss values are finalised.
    float* dSf32 = e->falloc(n * C);
    const int64_t fwork_n = ss_fourier_work_floats(B * C, H, W);     // patches other than power-of-two <= 128: DFT path
    float* fwork = fwork_n ? e->falloc(fwork_n) : nullptr;
    PUSH_SIDE(Lq, prof_note("loss:fourier_fft+grad", 0, 12.0 * 1048576.0 * B);
                  return ss_fourier_loss(e->x, e->S32, e->mask_dev, dSf32, e->four_partials, B * C, H, W,
                                         (float)(e->cfg.c_loss_fourier / ((double)B * C * H * W)), 0, fwork, st););
    DecompGeoms G2 = plan_decomp_fwd(e, Lq, Sb, d2, head2);          // model.py:546

    // algorithmic HBM bytes: 24.25 MiB per patch for the 5-term loss + gradients, 12 MiB for the Fourier term
    PUSH(Lq, prof_note("loss:pixel_terms+grads", 0, 24.25 * 1048576.0 * B);
             return ss_pixel_losses(e->x, e->R32, e->I32, e->Id32, Re32, e->cfg, B, C, H, W, e->pix_partials, dR32, dI32,
                                    dId32, dS32, dRe32, st););
    PUSH_JOIN(Lq);
    PUSH(Lq, return ss_launch_finalize_losses(e->pix_partials, e->pix_rows, e->four_partials, B * C, &e->cfg, e->losses, B,
                                              C, H, W, st););

    // backward, pass 2 (R_enh branch).  I_enh is unused (model.py:546) -> only C gradient columns.
    DecompGrads gr;
    gr.dc7 = e->talloc(B, H, W, 64); gr.dc5 = e->talloc(B, H, W, 64); gr.dc0 = e->talloc(B, H, W, 64, true);
    gr.ddc = e->talloc(B, H, W, 64); gr.dc1p = e->talloc(B, H, W, 64); gr.dc3 = e->talloc(B, H / 2, W / 2, 128);
    gr.dc2 = e->talloc(B, H / 2, W / 2, 128); gr.dc1 = e->talloc(B, H, W, 64); gr.dsh = e->talloc(B, H, W, 64);
    gr.din = e->talloc(B, H, W, 64);
    Tens dc8 = e->talloc(B, H, W, 128, true);
    PUSH(Lq, return ss_launch_head_bwd(dRe32, Re32, nullptr, 0, nullptr, nullptr, dc8.p, 128, B, C, H, W, st););
    plan_decomp_bwd(e, Lq, Sb, d2, G2, dc8, C, gr, true);
    // S = R*(Id+I)
    PUSH(Lq, return ss_launch_s_bwd(dS32, dSf32, gr.din.p, e->R32, e->I32, e->Id32, dR32, dI32, dId32, B, C, H, W, st););

    // IllumAdjustmentNet backward
    Tens dff = e->talloc(B, H, W, 64), dfg = e->talloc(B, H, W, 192), dr3 = e->talloc(B, H, W, 64),
         p2 = e->talloc(B, H / 2, W / 2, 64), p1 = e->talloc(B, H / 4, W / 4, 64), du3 = e->talloc(B, H, W, 64),
         dr2 = e->talloc(B, H / 2, W / 2, 64), da1p = e->talloc(B, H / 2, W / 2, 64),
         du2 = e->talloc(B, H / 2, W / 2, 64), dr1 = e->talloc(B, H / 4, W / 4, 64),
         da2p = e->talloc(B, H / 4, W / 4, 64), du1 = e->talloc(B, H / 4, W / 4, 64),
         da3 = e->talloc(B, H / 8, W / 8, 64), da2 = e->talloc(B, H / 4, W / 4, 64),
         da1 = e->talloc(B, H / 2, W / 2, 64), da0 = e->talloc(B, H, W, 64), dRI = e->talloc(B, H, W, 128, true);
    float* dt32 = e->falloc(T64);
    {
      const int64_t total = n * 8;
      const int fw_blocks = (int)std::min<int64_t>(FINAL_WGRAD_MAX_BLOCKS, (int64_t)B * H);
      PUSH_SIDE(Lq, final_wgrad_kernel<<<fw_blocks, 576, 0, st>>>(dId32, ff.p, e->final_partials, B, H, W);
               int rc = ss_check_launch("final_wgrad");
               if (rc) return rc;
               RedSegs segs;
               memset(&segs, 0, sizeof(segs));
               segs.n = 2;
               segs.dst[0] = e->grads + e->poff[2 * L_I_FINAL]; segs.len[0] = 576;
               segs.dst[1] = e->grads + e->poff[2 * L_I_FINAL + 1]; segs.len[1] = 1;
               return ss_launch_reduce_rows(e->final_partials, fw_blocks, 577, segs, st););
      PUSH(Lq, ss_launch_pdl(final_dgrad_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), (size_t)(0), st, 
                   dId32, e->params + e->poff[2 * L_I_FINAL], dff.p, H, W, total);
               return ss_check_launch("final_dgrad"););
    }
    queue_wgrad(e, Lq, g_fus, dff, 64, L_I_FUSION);
    {
      WAddr wa = waddr_conv_dgrad(e, L_I_FUSION);
      const int gi = e->add_geom(geom_conv(B, H, W, {{dff, 0, 64, 0}}, 1, 1, 0, -1, 192, wa));
      Epi ep = epi_bf16(dfg, 192);
      PUSH(Lq, return run_gather(e, gi, ep, -1, st););
    }
    PUSH(Lq, return ss_launch_concat_bwd(dfg.p, r3.p, dr3.p, p2.p, p1.p, B, H, W, st););
    // deconv3
    queue_wgrad(e, Lq, g_d3, dr3, 64, L_I_DECONV3);
    {
      WAddr wa = waddr_conv_dgrad(e, L_I_DECONV3);
      const int gi = e->add_geom(geom_conv(B, H, W, {{dr3, 0, 64, 0}}, 3, 1, 1, -1, 64, wa));
      Epi ep = epi_bf16(du3, 64);
      PUSH(Lq, return run_gather(e, gi, ep, -1, st););
    }
    PUSH(Lq, return ss_launch_pool2(du3.p, p2.p, r2.p, da1p.p, dr2.p, nullptr, B, H / 2, W / 2, st););
    // deconv2
    queue_wgrad(e, Lq, g_d2, dr2, 64, L_I_DECONV2);
    {
      WAddr wa = waddr_conv_dgrad(e, L_I_DECONV2);
      const int gi = e->add_geom(geom_conv(B, H / 2, W / 2, {{dr2, 0, 64, 0}}, 3, 1, 1, -1, 64, wa));
      Epi ep = epi_bf16(du2, 64);
      PUSH(Lq, return run_gather(e, gi, ep, -1, st););
    }
    PUSH(Lq, return ss_launch_pool2(du2.p, p1.p, r1.p, da2p.p, dr1.p, nullptr, B, H / 4, W / 4, st););
    // deconv1
    queue_wgrad(e, Lq, g_d1, dr1, 64, L_I_DECONV1);
    {
      WAddr wa = waddr_conv_dgrad(e, L_I_DECONV1);
      const int gi = e->add_geom(geom_conv(B, H / 4, W / 4, {{dr1, 0, 64, 0}}, 3, 1, 1, -1, 64, wa));
      Epi ep = epi_bf16(du1, 64);
      PUSH(Lq, return run_gather(e, gi, ep, -1, st););
    }
    PUSH(Lq, return ss_launch_pool2(du1.p, nullptr, nullptr, nullptr, nullptr, dt32, B, H / 8, W / 8, st););
    PUSH(Lq, return ss_attention_backward(dt32, a3.p, da3.p, e->params, e->grads, apoff, ab, B, L, st););
    PUSH_SIDE(Lq, return ss_attention_backward_weights(dt32, e->grads, apoff, ab, B, L, e->attn_partials, st););
    // conv3 / conv2 / conv1 (stride 2): wgrad on the forward geom, dgrad per input parity class (+ skip gradient)
    struct S2 { int layer, gfwd; Tens dy, x_in, dx, addp; bool mask; };
    const S2 s2[3] = {{L_I_CONV3, g_i3, da3, a2, da2, da2p, true},
                      {L_I_CONV2, g_i2, da2, a1, da1, da1p, true},
                      {L_I_CONV1, g_i1, da1, a0, da0, Tens(), false}};
    for (int i = 0; i < 3; ++i) {
      const S2 s = s2[i];
      queue_wgrad(e, Lq, s.gfwd, s.dy, 64, s.layer);
      Epi eps[4];
      int gi0 = -1;
      for (int q = 0; q < 4; ++q) {
        const int qh = q >> 1, qw = q & 1;
        WAddr wa = waddr_conv_dgrad(e, s.layer);
        const int gi = e->add_geom(geom_tconv_class(B, s.dy.H, s.dy.W, {s.dy, 0, 64, 0}, 3, 1, qh, qw, 64, wa));
        if (q == 0) gi0 = gi;
        eps[q] = epi_bf16(s.dx, 64, qh, qw, 2);
        if (s.mask) {
          epi_set_add(eps[q], s.addp, 0, qh, qw, 2);
          epi_set_mask(eps[q], s.x_in, qh, qw, 2);
        } else {
          epi_set_add(eps[q], dfg, 128, qh, qw, 2);     // d(a0) also receives dfg[...,128:192] (deconv3 + conv0 skip)
        }
      }
      const std::array<Epi, 4> ea = {eps[0], eps[1], eps[2], eps[3]};
      PUSH(Lq, return run_gather4(e, gi0, ea.data(), -1, st););
    }
    // conv0 of the illumination net reads cat[R, I]
    queue_wgrad(e, Lq, g_i0, da0, 64, L_I_CONV0);
    {
      WAddr wa = waddr_conv_dgrad(e, L_I_CONV0);
      const int gi = e->add_geom(geom_conv(B, H, W, {{da0, 0, 64, 0}}, 3, 1, 1, -1, C + 1, wa));
      Epi ep = epi_bf16(dRI, (C + 1 + 15) / 16 * 16);
      PUSH(Lq, return run_gather(e, gi, ep, -1, st););
    }
    PUSH_JOIN(Lq);

    // ---- backward, pass 1 -------------------------------------------------------------------
    auto& B1 = e->ops_bwd1;
    PUSH(B1, return ss_launch_head_bwd(dR32, e->R32, dRI.p, 128, dI32, e->I32, dc8.p, 128, B, C, H, W, st););
    plan_decomp_bwd(e, B1, X, d1, G1, dc8, C + 1, gr, false);
    PUSH_JOIN(B1);
  }

  if ((int)e->geoms.size() > 128) {
    ss_set_error("internal: %d geoms exceed the plan table", (int)e->geoms.size());
    return SSHSLIE_ERR_ARG;
  }
  // packed weights + pack job table.  Geoms that read the same weights the same way (the two decomposition passes and
  // their two backward passes lower every layer twice) share ONE packed copy: only the first is a pack job.
  e->pack_start.assign(e->geoms.size() + 1, 0);
  for (size_t i = 0; i < e->geoms.size(); ++i) {
    ConvGeom& g = e->geoms[i];
    const int64_t elems = (int64_t)g.Npad * g.nslabs * SS_SLAB;
    int twin = -1;
    for (size_t j = 0; j < i && twin < 0; ++j) {
      const ConvGeom& o = e->geoms[j];
      if (o.w_off != g.w_off || o.w_sN != g.w_sN || o.w_sC != g.w_sC || o.N != g.N || o.Npad != g.Npad ||
          o.nslabs != g.nslabs)
        continue;
      bool same = true;
      for (int k = 0; k < g.nslabs && same; ++k)
        same = o.slab[k].woff == g.slab[k].woff && o.slab[k].wcn == g.slab[k].wcn;
      if (same) twin = (int)j;
    }
    if (twin >= 0) {
      g.wp = e->geoms[twin].wp;
      e->pack_start[i + 1] = e->pack_start[i];
    } else {
      g.wp = (bf16*)e->alloc(elems * sizeof(bf16));
      e->pack_start[i + 1] = e->pack_start[i] + (int)((elems + 255) / 256);
    }
  }
  e->pack_blocks = e->pack_start.back();
  e->pack_split = e->geoms.size() > 2 ? e->pack_start[2] : e->pack_blocks;
  if (e->train) {
    size_t mx = 0;
    for (size_t i = 0; i < e->geoms.size(); ++i) {
      const size_t a = ss_umma_wgrad_partial_floats(e->geoms[i], 128), b = ss_umma_wgrad_partial_floats(e->geoms[i], 64);
      mx = std::max(mx, std::max(a, b));
      if (e->geoms[i].halo_ok)
        mx = std::max(mx, std::max(ss_umma_wgrad_halo_partial_floats(e->geoms[i], 128),
                                   ss_umma_wgrad_halo_partial_floats(e->geoms[i], 64)));
    }
    e->wg_partial_floats = mx;
    for (int i = 0; i < SS_MAX_SIDE; ++i) e->wg_partial[i] = e->falloc((int64_t)mx);
  }
  e->ws_bytes = e->cursor;
  return SSHSLIE_OK;
}

// Fourier mask on the un-shifted grid (model.py:460-464): torch.linspace(-1,1,n) semantics in fp32
static void make_mask(std::vector<float>& m, int H, int W) {
  auto lin = [](int i, int n) -> float {
    const float step = 2.0f / (float)(n - 1);
    return (i < n / 2) ? (-1.0f + step * (float)i) : (1.0f - step * (float)(n - 1 - i));
  };
  m.resize((size_t)H * W);
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      const float Y = lin(y, H), X = lin(x, W);
      m[(size_t)y * W + x] = (sqrtf(X * X + Y * Y) >= 0.1f) ? 1.f : 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" int sshslie_engine_create(sshslie_engine** out, int batch, int channels, int height, int width, int flags) {
  // DecompositionNet halves and re-doubles the image (stride-2 conv, then a transposed conv with output_padding 1,
  // model.py:37-43), so the reference itself needs even H and W; nothing else is required of an inference shape
  if (!out || batch < 1 || channels != 64 || height < 16 || width < 16 || (height % 2) || (width % 2)) {
    ss_set_error("sshslie_engine_create: need batch>=1, channels==64, even H,W >= 16 (got B=%d C=%d %dx%d)",
                 batch, channels, height, width);
    return SSHSLIE_ERR_ARG;
  }
  if ((flags & SSHSLIE_FLAG_TRAIN)) {
    // the backward kernels of the illumination net (2x2 pooling of upsample gradients, concat split) assume exact
    // halving at every level; the Fourier term takes any size (shared-memory FFT for powers of two <= 128, DFT otherwise)
    if ((height % 8) || (width % 8) || height > 1024 || width > 1024) {
      ss_set_error("sshslie_engine_create: training patches must be multiples of 8, at most 1024 a side (got %dx%d)",
                   height, width);
      return SSHSLIE_ERR_ARG;
    }
  }
  sshslie_engine* e = new sshslie_engine();
  e->B = batch; e->C = channels; e->H = height; e->W = width; e->flags = flags;
  e->train = (flags & SSHSLIE_FLAG_TRAIN) != 0;
  e->force_simt = (flags & SSHSLIE_FLAG_FORCE_SIMT) != 0;
  {
    const char* we = getenv("SSHSLIE_WGRAD_HALO");
    e->wgrad_halo = !(we && we[0] == '0');
    const char* he0 = getenv("SSHSLIE_HALO");
    e->halo_on = !(he0 && he0[0] == '0');
    const char* pe = getenv("SSHSLIE_PIPE");
    e->pipe_on = !(pe && pe[0] == '0');
    const char* pm = getenv("SSHSLIE_PIPE_MIN_TILES");
    if (pm && pm[0]) e->pipe_min_tiles = atoi(pm);
    const char* s2m = getenv("SSHSLIE_S2_MIN_TILES");
    if (s2m && s2m[0]) e->s2_min_tiles = atoi(s2m);
    const char* pmh = getenv("SSHSLIE_PIPE_MIN_TILES_HEAD");
    if (pmh && pmh[0]) e->pipe_min_tiles_head = atoi(pmh);
    const char* px = getenv("SSHSLIE_PIPE_MAX_SLABS");
    if (px && px[0]) e->pipe_max_slabs = atoi(px);
    const char* sk = getenv("SSHSLIE_SKIP_WGRAD");
    e->skip_wgrad = (sk && sk[0] == '1');
    const char* ns = getenv("SSHSLIE_NO_SIDE");
    if (ns && ns[0] == '1') e->use_side = false;
    const char* nsd = getenv("SSHSLIE_SIDE_STREAMS");
    if (nsd && atoi(nsd) >= 1 && atoi(nsd) <= SS_MAX_SIDE) e->n_side = atoi(nsd);
  }
  layer_shapes(channels, e->shapes);
  e->nparams = sshslie_param_table(channels, e->poff, e->psize);
  memset(&e->cfg, 0, sizeof(e->cfg));
  const int rc = build_plan(e, nullptr);   // dry run: sizes only
  if (rc != SSHSLIE_OK) { delete e; return rc; }
  *out = e;
  return SSHSLIE_OK;
}
extern "C" void sshslie_engine_destroy(sshslie_engine* e) {
  if (!e) return;
  for (int i = 0; i < SS_MAX_SIDE; ++i) {
    if (e->side[i]) cudaStreamDestroy(e->side[i]);
    if (e->ev_join[i]) cudaEventDestroy(e->ev_join[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  delete e;
}
extern "C" int64_t sshslie_engine_workspace_bytes(const sshslie_engine* e) { return e ? e->ws_bytes : 0; }

extern "C" int sshslie_engine_bind(sshslie_engine* e, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!e || !workspace) { ss_set_error("sshslie_engine_bind: null argument"); return SSHSLIE_ERR_ARG; }
  if (workspace_bytes < e->ws_bytes || ((uintptr_t)workspace & 1023)) {
    ss_set_error("sshslie_engine_bind: workspace must be >= %lld bytes and 1024-byte aligned", (long long)e->ws_bytes);
    return SSHSLIE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  e->bound = false;
  e->gmaps.clear();
  e->pipe_use.clear(); e->pipe_valid.clear(); e->pipe_blob.clear();
  if (!e->side[0]) {
    bool ok = cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) == cudaSuccess;
    for (int i = 0; i < SS_MAX_SIDE && ok; ++i)
      ok = cudaStreamCreateWithFlags(&e->side[i], cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&e->ev_join[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
      ss_set_error("bind: cannot create the side streams: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
  }
  int rc = build_plan(e, (unsigned char*)workspace);
  if (rc != SSHSLIE_OK) return rc;
  e->ws = (unsigned char*)workspace;
  // tcgen05 eligibility + TMA descriptors
  const size_t msz = ss_umma_maps_size();
  e->maps_blob.assign(e->geoms.size() * msz, 0);
  for (size_t i = 0; i < e->geoms.size(); ++i) {
    e->geom_umma[i] = 0;
    if (!e->force_simt && ss_umma_supported(e->geoms[i])) {
      rc = ss_umma_build_maps(e->geoms[i], reinterpret_cast<UmmaMaps*>(e->maps_blob.data() + i * msz));
      if (rc != SSHSLIE_OK) return rc;
      const char* he = getenv("SSHSLIE_HALO");
      const bool halo_on = !(he && he[0] == '0');  // halo-reuse kernels for every stride-1 layer; SSHSLIE_HALO=0 = per-tap
      e->geom_umma[i] = (halo_on && e->geoms[i].halo_ok && ss_umma_halo_supported(e->geoms[i])) ? 2 : 1;
    }
  }
  for (auto& zr : e->zero_ranges)
    if (cudaMemsetAsync(e->ws + zr.first, 0, (size_t)zr.second, st) != cudaSuccess) {
      ss_set_error("bind: memset failed: %s", cudaGetErrorString(cudaGetLastError()));
      return SSHSLIE_ERR_CUDA;
    }
  std::vector<float> mask;
  make_mask(mask, e->H, e->W);
  cudaError_t ce = cudaMemcpyAsync(e->geoms_dev, e->geoms.data(), e->geoms.size() * sizeof(ConvGeom),
                                   cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(e->pack_start_dev, e->pack_start.data(), e->pack_start.size() * sizeof(int),
                         cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess)
    ce = cudaMemcpyAsync(e->mask_dev, mask.data(), mask.size() * sizeof(float), cudaMemcpyHostToDevice, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) {
    ss_set_error("bind: upload failed: %s", cudaGetErrorString(ce));
    return SSHSLIE_ERR_CUDA;
  }
  e->bound = true;
  return SSHSLIE_OK;
}

static int run_ops(std::vector<sshslie_engine::OpFn>& ops, cudaStream_t st) {
  for (auto& f : ops) {
    const int rc = f(st);
    if (rc != SSHSLIE_OK) return rc;
  }
  return SSHSLIE_OK;
}
static int copy_out(float* dst, const float* src, int64_t n, cudaStream_t st) {
  if (!dst) return SSHSLIE_OK;
  if (cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
    ss_set_error("output copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return SSHSLIE_ERR_CUDA;
  }
  return SSHSLIE_OK;
}
static int copy_outputs(sshslie_engine* e, float* R, float* I, float* Id, float* S, cudaStream_t st) {
  const int64_t n = (int64_t)e->B * e->H * e->W;
  int rc = copy_out(R, e->R32, n * e->C, st);
  if (!rc) rc = copy_out(I, e->I32, n, st);
  if (!rc) rc = copy_out(Id, e->Id32, n, st);
  if (!rc) rc = copy_out(S, e->S32, n * e->C, st);
  return rc;
}

extern "C" int sshslie_forward(sshslie_engine* e, const float* x, const float* params, float* R, float* I,
                               float* I_delta, float* S, void* stream) {
  if (!e || !x || !params) { ss_set_error("sshslie_forward: null argument"); return SSHSLIE_ERR_ARG; }
  if (!e->bound) { ss_set_error("sshslie_forward: engine not bound to a workspace"); return SSHSLIE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  e->x = x; e->params = params;
  e->out_R = R; e->out_I = I; e->out_Id = I_delta; e->out_S = S;
  if (e->train) e->grads = nullptr;      // forward only on a training engine: nothing to zero
  int rc = run_ops(e->ops_fwd, st);
  if (!rc && e->train) rc = e->join(st);
  return rc;
}

// IllumAdjustmentNet.forward(I, R) on its own (model.py:143-175): packs cat[R, I] the way the sigmoid head leaves it for the
// illumination net (bf16 hi | I | bf16 residual), then runs the illumination part of the forward plan.
extern "C" int sshslie_illum_forward(sshslie_engine* e, const float* I, const float* R, const float* params, float* I_delta,
                                     void* stream) {
  if (!e || !I || !R || !params || !I_delta) { ss_set_error("sshslie_illum_forward: null argument"); return SSHSLIE_ERR_ARG; }
  if (!e->bound) { ss_set_error("sshslie_illum_forward: engine not bound to a workspace"); return SSHSLIE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  e->params = params;
  const bool saved_side = e->use_side;
  e->use_side = false;                       // everything on the caller's stream: the weight packing has no 9x9 layer to hide behind
  int rc = SSHSLIE_OK;
  for (int i = e->fwd_pack_begin; i < e->fwd_pack_begin + 2 && !rc; ++i) rc = e->ops_fwd[i](st);
  if (!rc) rc = ss_launch_pack_ri(R, I, e->RI_ptr, e->B, e->C, e->H, e->W, st);
  for (int i = e->fwd_illum_begin; i < e->fwd_illum_end && !rc; ++i) rc = e->ops_fwd[i](st);
  e->use_side = saved_side;
  if (rc) return rc;
  return copy_out(I_delta, e->Id32, (int64_t)e->B * e->H * e->W, st);
}

extern "C" int sshslie_loss_and_grad(sshslie_engine* e, const float* x, const float* params,
                                     const sshslie_loss_cfg* cfg, float* grads, float* losses, float* R, float* I,
                                     float* I_delta, float* S, int phase_mask, void* stream) {
  if (!e || !x || !params || !cfg || !grads || !losses) {
    ss_set_error("sshslie_loss_and_grad: null argument");
    return SSHSLIE_ERR_ARG;
  }
  if (!e->train) { ss_set_error("sshslie_loss_and_grad: engine created without SSHSLIE_FLAG_TRAIN"); return SSHSLIE_ERR_ARG; }
  if (!e->bound) { ss_set_error("sshslie_loss_and_grad: engine not bound to a workspace"); return SSHSLIE_ERR_WORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  e->x = x; e->params = params; e->grads = grads; e->losses = losses; e->cfg = *cfg;
  int rc = SSHSLIE_OK;
  if (phase_mask & 1) {
    e->out_R = R; e->out_I = I; e->out_Id = I_delta; e->out_S = S;
    rc = run_ops(e->ops_fwd, st);
    if (!rc) rc = run_ops(e->ops_loss_bwd2_illum, st);
  }
  if (!rc && (phase_mask & 2)) rc = run_ops(e->ops_bwd1, st);
  return rc;
}

// ---------------------------------------------------------------------------------------------
// single-layer entry point for kernel-level parity tests (tests/test_gpu_conv.py)
// ---------------------------------------------------------------------------------------------
static int pad64(int c) { return (c + 63) / 64 * 64; }
static std::atomic<float> g_conv2d_last_ms{0.f};
extern "C" SSHSLIE_API float sshslie_conv2d_last_ms(void) { return g_conv2d_last_ms.load(); }

extern "C" int64_t sshslie_conv2d_scratch_bytes(int B, int Cin, int Cout, int H, int W, int k, int stride) {
  (void)k;
  const int64_t big = (int64_t)B * H * W * (stride == 2 ? 4 : 1);   // the larger of the two spatial sizes
  int64_t bytes = 0;
  bytes += big * pad64(Cin) * 2 + 1024;
  bytes += big * pad64(Cout) * 2 + 1024;
  bytes += (int64_t)sizeof(ConvGeom) * 8 + 4096;
  bytes += 4 * ((int64_t)(pad64(Cin) + pad64(Cout)) * k * k * pad64(Cin > Cout ? Cin : Cout) * 2 + 1024);
  bytes += (int64_t)148 * 512 * 128 * 4 + 1024;     // split-K partials of the tcgen05 wgrad: <= 148 CTAs x all TMEM
  return bytes + (1 << 16);
}

extern "C" int sshslie_conv2d(int kind, int impl, int transposed, float* x, float* w, const float* bias, float* y,
                              int B, int Cin, int Cout, int H, int W, int k, int stride, int relu, void* scratch,
                              int64_t scratch_bytes, void* stream) {
  if (!x || !w || !y || !scratch || kind < 0 || kind > 2 || (stride != 1 && stride != 2) || (k != 1 && k != 3 && k != 9) ||
      (transposed && (stride != 2 || k != 3)) || (H % 8) || (W % 8) || ((uintptr_t)scratch & 1023)) {
    ss_set_error("sshslie_conv2d: unsupported arguments");
    return SSHSLIE_ERR_ARG;
  }
  if (scratch_bytes < sshslie_conv2d_scratch_bytes(B, Cin, Cout, H, W, k, stride)) {
    ss_set_error("sshslie_conv2d: scratch too small");
    return SSHSLIE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  sshslie_engine E;
  sshslie_engine* e = &E;
  e->B = B; e->C = 64; e->H = H; e->W = W; e->flags = 0; e->train = false; e->force_simt = (impl == 0);
  e->pipe_on = (impl == 3);               // impl 3 = the persistent pipelined gather kernel (stride-1 layers)
  e->pipe_min_tiles = 0; e->pipe_min_tiles_head = 0; e->pipe_max_slabs = SS_MAX_SLABS;
  e->base = (unsigned char*)scratch; e->cursor = 0;
  memset(e->poff, 0, sizeof(e->poff));
  e->shapes[0] = {transposed ? Cin : Cout, transposed ? Cout : Cin, k, transposed != 0};
  e->params = w; e->grads = w;
  const int pad = (k - 1) / 2;
  // spatial sizes: "small" side is the strided-conv output / transposed-conv input
  const int Hin = H, Win = W;
  const int Hout = transposed ? 2 * H : (stride == 2 ? H / 2 : H), Wout = transposed ? 2 * W : (stride == 2 ? W / 2 : W);
  Tens tin = e->talloc(B, Hin, Win, pad64(Cin)), tout = e->talloc(B, Hout, Wout, pad64(Cout));
  e->geoms_dev = (ConvGeom*)e->alloc(sizeof(ConvGeom) * 8);
  e->pack_start_dev = (int*)e->alloc(sizeof(int) * 16);
  cudaMemsetAsync(tin.p, 0, (size_t)tin.pix() * tin.ld * 2, st);
  cudaMemsetAsync(tout.p, 0, (size_t)tout.pix() * tout.ld * 2, st);

  struct Job { int gi; Epi ep; };
  std::vector<Job> jobs;
  const Tens* gt = nullptr;   // wgrad: the "G" tensor
  int gN = 0;
  if (kind == 0) {
    int rc = ss_launch_nchw32_to_nhwc16(x, tin.p, B, Cin, Hin, Win, tin.ld, st);
    if (rc) return rc;
    WAddr wa = waddr_conv_fwd(e, 0);
    if (!transposed) {
      const int gi = e->add_geom(geom_conv(B, Hout, Wout, {{tin, 0, Cin, 0}}, k, stride, pad, +1, Cout, wa));
      Epi ep = epi_bf16(tout, (Cout + 15) / 16 * 16); ep.relu = relu; ep.bias = bias;
      jobs.push_back({gi, ep});
    } else {
      for (int q = 0; q < 4; ++q) {
        const int gi = e->add_geom(geom_tconv_class(B, Hin, Win, {tin, 0, Cin, 0}, 3, 1, q >> 1, q & 1, Cout, wa));
        Epi ep = epi_bf16(tout, (Cout + 15) / 16 * 16, q >> 1, q & 1, 2); ep.relu = relu; ep.bias = bias;
        jobs.push_back({gi, ep});
      }
    }
  } else if (kind == 1) {
    // x holds dY (B,Cout,Hout,Wout); y receives dX (B,Cin,Hin,Win)
    int rc = ss_launch_nchw32_to_nhwc16(x, tout.p, B, Cout, Hout, Wout, tout.ld, st);
    if (rc) return rc;
    WAddr wa = waddr_conv_dgrad(e, 0);
    if (!transposed && stride == 1) {
      const int gi = e->add_geom(geom_conv(B, Hin, Win, {{tout, 0, Cout, 0}}, k, 1, pad, -1, Cin, wa));
      jobs.push_back({gi, epi_bf16(tin, (Cin + 15) / 16 * 16)});
    } else if (!transposed) {
      for (int q = 0; q < 4; ++q) {
        const int gi = e->add_geom(geom_tconv_class(B, Hout, Wout, {tout, 0, Cout, 0}, 3, 1, q >> 1, q & 1, Cin, wa));
        jobs.push_back({gi, epi_bf16(tin, (Cin + 15) / 16 * 16, q >> 1, q & 1, 2)});
      }
    } else {
      const int gi = e->add_geom(geom_conv(B, Hin, Win, {{tout, 0, Cout, 0}}, 3, 2, 1, +1, Cin, wa));
      jobs.push_back({gi, epi_bf16(tin, (Cin + 15) / 16 * 16)});
    }
  } else {
    // wgrad: x = layer input (B,Cin,Hin,Win), y = dY (B,Cout,Hout,Wout), w <- dW
    int rc = ss_launch_nchw32_to_nhwc16(x, tin.p, B, Cin, Hin, Win, tin.ld, st);
    if (!rc) rc = ss_launch_nchw32_to_nhwc16(y, tout.p, B, Cout, Hout, Wout, tout.ld, st);
    if (rc) return rc;
    const int64_t wn = (int64_t)Cin * Cout * k * k;
    cudaMemsetAsync(w, 0, (size_t)wn * sizeof(float), st);
    if (bias) {   // kind 2: `bias` receives db (addressed relative to the gradient base like a layer's bias slice)
      cudaMemsetAsync(const_cast<float*>(bias), 0, (size_t)Cout * sizeof(float), st);
      float* lo = (bias < w) ? const_cast<float*>(bias) : w;      // offsets must be non-negative: rebase on the lower pointer
      e->grads = lo;
      e->poff[0] = (int64_t)(w - lo);
      e->poff[1] = (int64_t)(bias - lo);
    }
    if (!transposed) {
      WAddr wa = waddr_conv_fwd(e, 0);
      e->add_geom(geom_conv(B, Hout, Wout, {{tin, 0, Cin, 0}}, k, stride, pad, +1, Cout, wa));
      gt = &tout; gN = Cout;
    } else {
      WAddr wa = waddr_conv_dgrad(e, 0);
      e->add_geom(geom_conv(B, Hin, Win, {{tout, 0, Cout, 0}}, 3, 2, 1, +1, Cin, wa));
      gt = &tin; gN = Cin;
    }
  }
  // packed weights, plan upload, descriptors
  e->pack_start.assign(e->geoms.size() + 1, 0);
  for (size_t i = 0; i < e->geoms.size(); ++i) {
    ConvGeom& g = e->geoms[i];
    const int64_t elems = (int64_t)g.Npad * g.nslabs * SS_SLAB;
    g.wp = (bf16*)e->alloc(elems * sizeof(bf16));
    e->pack_start[i + 1] = e->pack_start[i] + (int)((elems + 255) / 256);
  }
  if (kind == 2 && impl >= 1) {
    e->wg_partial_floats = std::max(ss_umma_wgrad_partial_floats(e->geoms[0], gN),
                                    ss_umma_wgrad_halo_partial_floats(e->geoms[0], gN));
    e->wg_partial[0] = e->falloc((int64_t)e->wg_partial_floats);
  }
  if (e->cursor > scratch_bytes) { ss_set_error("sshslie_conv2d: scratch overflow"); return SSHSLIE_ERR_WORKSPACE; }
  const size_t msz = ss_umma_maps_size();
  e->maps_blob.assign(e->geoms.size() * msz, 0);
  for (size_t i = 0; i < e->geoms.size(); ++i) {
    if (impl >= 1) {
      if (!ss_umma_supported(e->geoms[i]) || (impl >= 2 && !ss_umma_halo_supported(e->geoms[i]))) { ss_set_error("sshslie_conv2d: shape not taken by the tcgen05 kernel"); return SSHSLIE_ERR_ARG; }
      int rc = ss_umma_build_maps(e->geoms[i], reinterpret_cast<UmmaMaps*>(e->maps_blob.data() + i * msz));
      if (rc) return rc;
      e->geom_umma[i] = (char)(impl == 3 ? 2 : impl);
    }
  }
  cudaMemcpyAsync(e->geoms_dev, e->geoms.data(), e->geoms.size() * sizeof(ConvGeom), cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(e->pack_start_dev, e->pack_start.data(), e->pack_start.size() * sizeof(int), cudaMemcpyHostToDevice, st);
  int rc = SSHSLIE_OK;
  if (kind != 2) {
    rc = ss_launch_pack_weights(e->geoms_dev, e->pack_start_dev, (int)e->geoms.size(), e->pack_start.back(), w, st);
    for (size_t j = 0; j < jobs.size() && !rc; ++j) rc = run_gather(e, jobs[j].gi, jobs[j].ep, -1, st);
    if (!rc) {
      if (kind == 0) rc = ss_launch_nhwc16_to_nchw32(tout.p, y, B, Cout, Hout, Wout, tout.ld, st);
      else rc = ss_launch_nhwc16_to_nchw32(tin.p, y, B, Cin, Hin, Win, tin.ld, st);
    }
  } else {
    rc = run_wgrad(e, 0, *gt, gN, 0, 0, 1, st, (bias && !transposed) ? 0 : -1);
  }
  // SSHSLIE_CONV2D_TIMING=n: repeat the layer's own launches n times between two events (tools/conv_bench.py)
  if (const char* te = getenv("SSHSLIE_CONV2D_TIMING")) {
    const int reps = atoi(te);
    if (reps > 0 && !rc) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
      for (int r = 0; r < reps && !rc; ++r) {
        if (kind != 2) for (size_t j = 0; j < jobs.size() && !rc; ++j) rc = run_gather(e, jobs[j].gi, jobs[j].ep, -1, st);
        else rc = run_wgrad(e, 0, *gt, gN, 0, 0, 1, st, (bias && !transposed) ? 0 : -1);
      }
      cudaEventRecord(b, st);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      g_conv2d_last_ms = ms / (float)reps;
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) {   // the plan lives on this stack frame: finish before returning
    ss_set_error("sshslie_conv2d: %s", cudaGetErrorString(cudaGetLastError()));
    return SSHSLIE_ERR_CUDA;
  }
  return rc;
}


// ---------------------------------------------------------------------------------------------
// TransformerBlock alone (model.py:99-119), for kernel-level parity tests of the attention kernels: x, y, dy, dx are
// fp32 (B, 64, H, W) as the reference block sees them (tokens = the H*W grid); params / dparams are the block's ten tensors
// flat in state_dict order (q.w q.b k.w k.b v.w v.b ff1.w ff1.b ff2.w ff2.b = 20800 floats).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) nchw32_to_tokens32_kernel(const float* __restrict__ x, float* __restrict__ t, int C,
                                                                 int L, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bl = i / C;
  const int l = (int)(bl % L);
  const int64_t b = bl / L;
  t[i] = x[(b * C + c) * L + l];
}
// (B, 64, L) fp32 <-> (B, L, 64) bf16 for any L (the engine's layout converters assume its 8-pixel tiling)
__global__ void __launch_bounds__(256) nchw32_to_tokens16_kernel(const float* __restrict__ x, bf16* __restrict__ t, int C,
                                                                 int L, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bl = i / C;
  t[i] = f2bf(x[((bl / L) * C + c) * L + (bl % L)]);
}
__global__ void __launch_bounds__(256) tokens16_to_nchw32_kernel(const bf16* __restrict__ t, float* __restrict__ y, int C,
                                                                 int L, int64_t total) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  const int64_t bl = i / C;
  y[((bl / L) * C + c) * L + (bl % L)] = bf2f(t[i]);
}
extern "C" int64_t sshslie_transformer_block_scratch_bytes(int B, int H, int W) {
  const int64_t T64 = (int64_t)B * H * W * 64;
  int64_t bytes = 3 * (T64 * 2 + 1024);                                   // a3, t, da3 (bf16)
  bytes += 16 * (T64 * 4 + 1024) + 2 * ((int64_t)B * 4 * H * W * 4 + 1024);   // fp32 token buffers, lse, Dv
  bytes += 2 * (T64 * 4 * 2 + 1024);                                      // qp, kvp
  bytes += (int64_t)SS_ATTN_WGRAD_MAX_BLOCKS * SS_ATTN_WGRAD_COLS * 4 + 1024;
  return bytes + (1 << 16);
}
extern "C" int sshslie_transformer_block(int with_backward, const float* x, const float* params, float* y, const float* dy,
                                         float* dx, float* dparams, int B, int H, int W, void* scratch,
                                         int64_t scratch_bytes, void* stream) {
  if (!x || !params || !y || !scratch || B < 1 || H < 1 || W < 1 || ((uintptr_t)scratch & 1023) ||
      (with_backward && (!dy || !dx || !dparams))) {
    ss_set_error("sshslie_transformer_block: bad argument");
    return SSHSLIE_ERR_ARG;
  }
  if (scratch_bytes < sshslie_transformer_block_scratch_bytes(B, H, W)) {
    ss_set_error("sshslie_transformer_block: scratch too small");
    return SSHSLIE_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  sshslie_engine E;
  sshslie_engine* e = &E;
  e->base = (unsigned char*)scratch; e->cursor = 0;
  const int L = H * W;
  const int64_t T64 = (int64_t)B * L * 64;
  bf16* a3 = (bf16*)e->alloc(T64 * 2);
  bf16* tt = (bf16*)e->alloc(T64 * 2);
  bf16* da3 = (bf16*)e->alloc(T64 * 2);
  AttnBuffers ab;
  memset(&ab, 0, sizeof(ab));
  ab.x = e->falloc(T64); ab.q = e->falloc(T64); ab.k = e->falloc(T64); ab.v = e->falloc(T64); ab.o = e->falloc(T64);
  ab.lse = e->falloc((int64_t)B * 4 * L); ab.h = e->falloc(T64); ab.t32 = e->falloc(T64);
  ab.dq = e->falloc(T64); ab.dk = e->falloc(T64); ab.dv = e->falloc(T64); ab.d_o = e->falloc(T64);
  ab.dh = e->falloc(T64); ab.dx = e->falloc(T64); ab.Dv = e->falloc((int64_t)B * 4 * L);
  float* dt32 = e->falloc(T64);
  if (L >= ss_attn_tc_min_l()) {
    ab.qp = (bf16*)e->alloc(T64 * 4 * 2);
    ab.kvp = (bf16*)e->alloc(T64 * 4 * 2);
  }
  float* partials = e->falloc((int64_t)SS_ATTN_WGRAD_MAX_BLOCKS * SS_ATTN_WGRAD_COLS);
  int64_t poff[10];
  for (int i = 0; i < 5; ++i) { poff[2 * i] = (int64_t)i * 4160; poff[2 * i + 1] = (int64_t)i * 4160 + 4096; }
  const unsigned cgrid = (unsigned)((T64 + 255) / 256);
  nchw32_to_tokens16_kernel<<<cgrid, 256, 0, st>>>(x, a3, 64, L, T64);
  int rc = ss_check_launch("nchw32_to_tokens16");
  if (!rc) rc = ss_attention_forward(a3, tt, params, poff, ab, B, L, st);
  if (!rc) {
    tokens16_to_nchw32_kernel<<<cgrid, 256, 0, st>>>(tt, y, 64, L, T64);
    rc = ss_check_launch("tokens16_to_nchw32");
  }
  if (!rc && with_backward) {
    nchw32_to_tokens32_kernel<<<(unsigned)((T64 + 255) / 256), 256, 0, st>>>(dy, dt32, 64, L, T64);
    rc = ss_check_launch("nchw32_to_tokens32");
    if (!rc && cudaMemsetAsync(dparams, 0, 20800 * sizeof(float), st) != cudaSuccess) rc = SSHSLIE_ERR_CUDA;
    if (!rc) rc = ss_attention_backward(dt32, a3, da3, params, dparams, poff, ab, B, L, st);
    if (!rc) rc = ss_attention_backward_weights(dt32, dparams, poff, ab, B, L, partials, st);
    if (!rc) {
      tokens16_to_nchw32_kernel<<<cgrid, 256, 0, st>>>(da3, dx, 64, L, T64);
      rc = ss_check_launch("tokens16_to_nchw32");
    }
  }
  if (cudaStreamSynchronize(st) != cudaSuccess) {
    ss_set_error("sshslie_transformer_block: %s", cudaGetErrorString(cudaGetLastError()));
    return SSHSLIE_ERR_CUDA;
  }
  return rc;
}

// ---------------------------------------------------------------------------------------------
// profiling entry point: one eager step with a cudaEvent pair around every recorded op (synchronises)
// ---------------------------------------------------------------------------------------------
struct ProfRow { std::string name; float ms; double flops, bytes; };
static thread_local std::vector<ProfRow> g_prof_rows;

static int time_op(sshslie_engine::OpFn& f, cudaStream_t st, int reps, float* ms_out) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a, st);
  int rc = SSHSLIE_OK;
  for (int r = 0; r < reps && rc == SSHSLIE_OK; ++r) {
    g_prof_names.clear();
    g_prof_flops = g_prof_bytes = 0;
    rc = f(st);
  }
  cudaEventRecord(b, st);
  if (rc == SSHSLIE_OK) {
    cudaEventSynchronize(b);
    cudaEventElapsedTime(ms_out, a, b);
    *ms_out /= (float)reps;
  }
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  return rc;
}
// Every recorded op is enqueued `reps` times back to back between two events and the time divided: one eager launch +
// event pair costs ~6-8 us of host/driver latency that the CUDA-graph replay of the real step does not pay (the repeats
// only disturb values of this profiling pass: gradients accumulate, in-place adds repeat).  A tcgen05 weight gradient
// is two kernels (GEMM + split-K reduce): they are timed separately and reported as two rows.
static int run_ops_profiled(std::vector<sshslie_engine::OpFn>& ops, cudaStream_t st, const char* phase) {
  static const int reps = []() { const char* r = getenv("SSHSLIE_PROFILE_REPS"); return (r && atoi(r) > 0) ? atoi(r) : 4; }();
  for (auto& f : ops) {
    float ms = 0;
    int rc = time_op(f, st, reps, &ms);
    if (rc != SSHSLIE_OK) return rc;
    const std::string name = g_prof_names.empty() ? "memop" : g_prof_names;
    if (name.compare(0, 6, "wgrad:") == 0 && name.find("[tcgen05") != std::string::npos) {
      const double fl = g_prof_flops;
      float ms_main = 0, ms_red = 0;
      ss_set_wgrad_part(1);
      rc = time_op(f, st, reps, &ms_main);
      ss_set_wgrad_part(2);
      if (rc == SSHSLIE_OK) rc = time_op(f, st, reps, &ms_red);
      ss_set_wgrad_part(0);
      if (rc != SSHSLIE_OK) return rc;
      g_prof_rows.push_back({std::string(phase) + "/" + name, ms_main, fl, 0});
      g_prof_rows.push_back({std::string(phase) + "/" + "splitk_reduce:" + name.substr(6), ms_red, 0, 0});
      continue;
    }
    g_prof_rows.push_back({std::string(phase) + "/" + name, ms, g_prof_flops, g_prof_bytes});
  }
  return SSHSLIE_OK;
}

extern "C" SSHSLIE_API int sshslie_profile_step(sshslie_engine* e, const float* x, const float* params,
                                                const sshslie_loss_cfg* cfg, float* grads, float* losses,
                                                void* stream) {
  if (!e || !x || !params || !e->bound) { ss_set_error("sshslie_profile_step: bad argument"); return SSHSLIE_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  e->x = x; e->params = params;
  g_prof_rows.clear();
  g_prof_on = true;
  const bool saved_side = e->use_side;
  e->use_side = false;
  int rc = run_ops_profiled(e->ops_fwd, st, "fwd");
  if (!rc && e->train && cfg && grads && losses) {
    e->grads = grads; e->losses = losses; e->cfg = *cfg;
    rc = run_ops_profiled(e->ops_loss_bwd2_illum, st, "pass2+loss+illum_bwd");
    if (!rc) rc = run_ops_profiled(e->ops_bwd1, st, "pass1_bwd");
  }
  g_prof_on = false;
  e->use_side = saved_side;
  return rc ? rc : (int)g_prof_rows.size();
}
extern "C" SSHSLIE_API int sshslie_profile_row(int i, char* name, int name_cap, float* ms, double* flops,
                                               double* bytes) {
  if (i < 0 || i >= (int)g_prof_rows.size()) return SSHSLIE_ERR_ARG;
  const ProfRow& r = g_prof_rows[i];
  if (name && name_cap > 0) { strncpy(name, r.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
  if (ms) *ms = r.ms;
  if (flops) *flops = r.flops;
  if (bytes) *bytes = r.bytes;
  return SSHSLIE_OK;
}
