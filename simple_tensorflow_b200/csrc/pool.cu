// MaxPool / MaxPoolGrad, NHWC, HBM-bound.
//
// Reference kernels replaced (relative to tensorflow/core/kernels/):
//   MaxPoolForwardNHWC          maxpooling_op_gpu.cu.cc:93-129  (one thread per output scalar)
//   MaxPoolBackwardNoMaskNHWC   maxpooling_op_gpu.cu.cc:133-177 (SetZero + atomicAdd scatter)
// Semantics follow the CPU kernels the oracle restates: pooling_ops_common.h:204-238 (forward;
// output starts at lowest(), padded cells never participate) and maxpooling_op.cc:52-188
// (grad; first maximum in row-major window scan wins, strict '<' comparison).
//
// Forward: one thread per (n, oh, ow, 16-byte channel vector); consecutive threads walk the
// channel dimension so every load/store is a coalesced 128-bit access.
// Grad: gather formulation, one thread per INPUT element vector: for each window covering the
// element the window's argmax is recomputed and the gradient added when it is this element.
// No atomics, no zero-fill pass, deterministic order (ascending output index, as the reference's
// `in_backprop[argmax] += grad` loop, maxpooling_op.cc:165-178).
#include <atomic>
#include <cfloat>
#include <cuda_bf16.h>

#include "b200_internal.h"
#include "ordered_reduce.cuh"

namespace b200 {

struct PoolGeom {
  int N, H, W, C, OH, OW;
  int wh, ww, sh, sw, pt, pl;
};

template <typename T, int V>
struct PoolVec;
template <>
struct PoolVec<float, 4> {
  __device__ static void load(const float* p, float (&f)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x;
    f[1] = v.y;
    f[2] = v.z;
    f[3] = v.w;
  }
  __device__ static void store(float* p, const float (&f)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(f[0], f[1], f[2], f[3]);
  }
};
template <>
struct PoolVec<float, 1> {
  __device__ static void load(const float* p, float (&f)[1]) { f[0] = __ldg(p); }
  __device__ static void store(float* p, const float (&f)[1]) { *p = f[0]; }
};
template <>
struct PoolVec<__nv_bfloat16, 8> {
  __device__ static void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
template <>
struct PoolVec<__nv_bfloat16, 1> {
  __device__ static void load(const __nv_bfloat16* p, float (&f)[1]) {
    f[0] = __bfloat162float(*p);
  }
  __device__ static void store(__nv_bfloat16* p, const float (&f)[1]) {
    *p = __float2bfloat16_rn(f[0]);
  }
};

template <typename T, int V>
__global__ void __launch_bounds__(256)
max_pool_fwd_kernel(const T* __restrict__ in, T* __restrict__ out, PoolGeom g, long long total) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int CV = g.C / V;
  const int cv = (int)(idx % CV);
  long long r = idx / CV;
  const int ow = (int)(r % g.OW);
  r /= g.OW;
  const int oh = (int)(r % g.OH);
  const int n = (int)(r / g.OH);
  int h0 = oh * g.sh - g.pt, w0 = ow * g.sw - g.pl;
  const int h1 = min(h0 + g.wh, g.H), w1 = min(w0 + g.ww, g.W);
  h0 = max(h0, 0);
  w0 = max(w0, 0);
  float best[V];
#pragma unroll
  for (int j = 0; j < V; ++j) best[j] = -FLT_MAX;  // NumTraits<T>::lowest()
  const T* base = in + ((long long)n * g.H * g.W) * g.C + (long long)cv * V;
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) {
      float v[V];
      PoolVec<T, V>::load(base + ((long long)h * g.W + w) * g.C, v);
#pragma unroll
      for (int j = 0; j < V; ++j) best[j] = best[j] > v[j] ? best[j] : v[j];  // cwiseMax
    }
  PoolVec<T, V>::store(out + idx * V, best);
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
max_pool_grad_kernel(const T* __restrict__ in, const T* __restrict__ grad, T* __restrict__ dx,
                     PoolGeom g, long long total) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int CV = g.C / V;
  const int cv = (int)(idx % CV);
  long long r = idx / CV;
  const int w = (int)(r % g.W);
  r /= g.W;
  const int h = (int)(r % g.H);
  const int n = (int)(r / g.H);
  const T* ibase = in + ((long long)n * g.H * g.W) * g.C + (long long)cv * V;
  const T* gbase = grad + ((long long)n * g.OH * g.OW) * g.C + (long long)cv * V;
  float me[V], acc[V];
  PoolVec<T, V>::load(ibase + ((long long)h * g.W + w) * g.C, me);
#pragma unroll
  for (int j = 0; j < V; ++j) acc[j] = 0.f;
  // output windows that contain (h, w): pooling_ops_common.h:215-224 / maxpooling_op.cc:124-131
  const int hpad = h + g.pt, wpad = w + g.pl;
  const int ph0 = hpad < g.wh ? 0 : (hpad - g.wh) / g.sh + 1;
  const int ph1 = min(hpad / g.sh + 1, g.OH);
  const int pw0 = wpad < g.ww ? 0 : (wpad - g.ww) / g.sw + 1;
  const int pw1 = min(wpad / g.sw + 1, g.OW);
  for (int ph = ph0; ph < ph1; ++ph)
    for (int pw = pw0; pw < pw1; ++pw) {
      int hs = ph * g.sh - g.pt, ws = pw * g.sw - g.pl;
      const int he = min(hs + g.wh, g.H), we = min(ws + g.ww, g.W);
      hs = max(hs, 0);
      ws = max(ws, 0);
      // Recompute the window's argmax in row-major scan order.  I win channel j iff no earlier
      // cell is >= me (an earlier equal value keeps the slot: strict '<' in maxpooling_op.cc:140)
      // and no later cell is > me.
      bool win[V];
#pragma unroll
      for (int j = 0; j < V; ++j) win[j] = true;
      for (int hh = hs; hh < he; ++hh)
        for (int wc = ws; wc < we; ++wc) {
          if (hh == h && wc == w) continue;
          float v[V];
          PoolVec<T, V>::load(ibase + ((long long)hh * g.W + wc) * g.C, v);
          const bool earlier = hh < h || (hh == h && wc < w);
#pragma unroll
          for (int j = 0; j < V; ++j) {
            // earlier cell keeps the slot unless incumbent < me; later cell takes it if me < v
            if (earlier ? !(v[j] < me[j]) : (me[j] < v[j])) win[j] = false;
          }
        }
      float gv[V];
      PoolVec<T, V>::load(gbase + ((long long)ph * g.OW + pw) * g.C, gv);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (win[j]) acc[j] += gv[j];
    }
  PoolVec<T, V>::store(dx + idx * V, acc);
}

// Non-overlapping windows (stride >= window, the LeNet 2x2/2 case): every input cell belongs to at
// most one window, so one thread per OUTPUT cell reads its window once, finds the first maximum and
// writes the whole window's gradients (winner = grad, others = 0).  Each input is read exactly
// once and nothing is recomputed; cells covered by no window are zeroed by a memset beforehand
// (only needed when the windows do not tile the input).
template <typename T, int V>
__global__ void __launch_bounds__(256)
max_pool_grad_disjoint_kernel(const T* __restrict__ in, const T* __restrict__ grad,
                              T* __restrict__ dx, PoolGeom g, long long total) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int CV = g.C / V;
  const int cv = (int)(idx % CV);
  long long r = idx / CV;
  const int ow = (int)(r % g.OW);
  r /= g.OW;
  const int oh = (int)(r % g.OH);
  const int n = (int)(r / g.OH);
  int h0 = oh * g.sh - g.pt, w0 = ow * g.sw - g.pl;
  const int h1 = min(h0 + g.wh, g.H), w1 = min(w0 + g.ww, g.W);
  h0 = max(h0, 0);
  w0 = max(w0, 0);
  const long long ibase = ((long long)n * g.H * g.W) * g.C + (long long)cv * V;
  float best[V], gv[V];
  int arg[V];
#pragma unroll
  for (int j = 0; j < V; ++j) {
    best[j] = -FLT_MAX;
    arg[j] = -1;
  }
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) {
      float v[V];
      PoolVec<T, V>::load(in + ibase + ((long long)h * g.W + w) * g.C, v);
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (best[j] < v[j] || arg[j] == -1) {  // maxpooling_op.cc:139-144: first maximum wins
          best[j] = v[j];
          arg[j] = h * g.W + w;
        }
    }
  PoolVec<T, V>::load(grad + idx * V, gv);
  for (int h = h0; h < h1; ++h)
    for (int w = w0; w < w1; ++w) {
      float o[V];
#pragma unroll
      for (int j = 0; j < V; ++j) o[j] = (arg[j] == h * g.W + w) ? gv[j] : 0.f;
      PoolVec<T, V>::store(dx + ibase + ((long long)h * g.W + w) * g.C, o);
    }
}

// MaxPoolGrad -> ReluGrad -> BiasAddGrad in one pass (the backward tail of a conv -> bias -> relu ->
// pool block: LeNet's two, VGG's five).  The pool's input IS the Relu output, i.e. the `features`
// of the ReluGrad, so one read of it decides both the window's winner and the relu mask:
//   dx[cell]  = (cell is the window's first maximum) ? grad[window] : 0     (maxpooling_op.cc:124-144)
//   dy[cell]  = features[cell] > 0 ? dx[cell] : dx[cell] * 0                (relu_op_functor.h:54-55)
//   db[c]    += dy[cell]                                                     (bias_op.cc:210-226)
// Unfused this is 115 + 154 MB of traffic for LeNet's pool1 (dx written, then read again with the
// features); fused it is 115 MB.  Windows must tile the input (stride == window, no padding) and
// the 16-byte channel vectors of a pixel must divide a warp (G = C / V in {1, 2, .., 32}), so a
// thread meets one channel group only: the partial sums fold like flat_bias_grad_kernel's (shuffles,
// shared memory, one partial row per CTA, ordered last-ticket pass; no atomics on the data).
constexpr int kPoolFusedSlots = 32;
__device__ unsigned int g_pool_fused_tickets[kPoolFusedSlots];

template <typename T, int V>
__global__ void __launch_bounds__(256)
max_pool_grad_relu_bias_kernel(const T* __restrict__ in, const T* __restrict__ grad,
                               T* __restrict__ dy, float* __restrict__ partial,
                               T* __restrict__ bias_grad, PoolGeom g, long long total, int G,
                               unsigned int* __restrict__ ticket) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float bsum[V];
#pragma unroll
  for (int j = 0; j < V; ++j) bsum[j] = 0.f;
  // the grid stride is a multiple of G (256 is), so idx % G -- the channel group -- never changes
  for (long long idx = (long long)blockIdx.x * 256 + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * 256) {
    const int cv = (int)(idx % G);
    long long r = idx / G;
    const int ow = (int)(r % g.OW);
    r /= g.OW;
    const int oh = (int)(r % g.OH);
    const int n = (int)(r / g.OH);
    const int h0 = oh * g.sh, w0 = ow * g.sw;
    const int h1 = min(h0 + g.wh, g.H), w1 = min(w0 + g.ww, g.W);
    const long long ibase = ((long long)n * g.H * g.W) * g.C + (long long)cv * V;
    float best[V], feat[V], gv[V];
    int arg[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      best[j] = -FLT_MAX;
      arg[j] = -1;
    }
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        float v[V];
        PoolVec<T, V>::load(in + ibase + ((long long)h * g.W + w) * g.C, v);
#pragma unroll
        for (int j = 0; j < V; ++j)
          if (best[j] < v[j] || arg[j] == -1) {  // first maximum wins
            best[j] = v[j];
            arg[j] = h * g.W + w;
          }
      }
    PoolVec<T, V>::load(grad + idx * V, gv);
    // the winner's feature value is `best`: only the winner can carry a non-zero gradient
#pragma unroll
    for (int j = 0; j < V; ++j) {
      feat[j] = best[j];
      gv[j] = feat[j] > 0.f ? gv[j] : gv[j] * 0.f;
      bsum[j] += gv[j];
    }
    for (int h = h0; h < h1; ++h)
      for (int w = w0; w < w1; ++w) {
        float o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = (arg[j] == h * g.W + w) ? gv[j] : 0.f;
        PoolVec<T, V>::store(dy + ibase + ((long long)h * g.W + w) * g.C, o);
      }
  }
  // ---- bias gradient: lanes with the same lane % G hold the same channel group
  for (int off = G; off < 32; off <<= 1) {
#pragma unroll
    for (int j = 0; j < V; ++j) bsum[j] += __shfl_xor_sync(0xffffffffu, bsum[j], off);
  }
  __shared__ float sm[8][32][V];
  if (lane < G) {
#pragma unroll
    for (int j = 0; j < V; ++j) sm[warp][lane][j] = bsum[j];
  }
  __syncthreads();
  const int C = G * V;
  if ((int)threadIdx.x < C) {
    const int cg = threadIdx.x / V, j = threadIdx.x % V;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += sm[w][cg][j];
    partial[(long long)blockIdx.x * C + threadIdx.x] = t;
    __threadfence();
  }
  __shared__ bool is_last;
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // ordered final pass over the per-CTA partial rows (ordered_reduce.cuh)
  __shared__ uint4 slots[8 * 256];
  __shared__ float flat[256 * 4];
  const float tot = ordered_partial_sum(partial, (int)gridDim.x, C, slots, flat);
  if ((int)threadIdx.x < C) {
    if (sizeof(T) == 4)
      reinterpret_cast<float*>(bias_grad)[threadIdx.x] = tot;
    else
      reinterpret_cast<__nv_bfloat16*>(bias_grad)[threadIdx.x] = __float2bfloat16_rn(tot);
  }
  if (threadIdx.x == 0) *ticket = 0;  // ready for the next launch that uses this slot
}

struct PoolFusedPlan {
  bool ok;
  int G, V, blocks;
  long long total;
};
static PoolFusedPlan plan_pool_fused(int dtype, const PoolGeom& g, int64_t batch) {
  PoolFusedPlan p{false, 0, 0, 0, 0};
  p.V = dtype == B200_DT_FLOAT ? 4 : 8;
  if (g.C % p.V != 0) return p;
  const int G = g.C / p.V;
  if (G < 1 || G > 32 || (G & (G - 1)) != 0) return p;
  // the windows tile the input exactly: every input cell belongs to exactly one window
  if (g.sh != g.wh || g.sw != g.ww || g.pt != 0 || g.pl != 0 || g.OH <= 0 || g.OW <= 0 ||
      (long long)g.OH * g.sh < g.H || (long long)g.OW * g.sw < g.W)
    return p;
  p.G = G;
  p.total = (long long)batch * g.OH * g.OW * G;
  long long blocks = (p.total + 255) / 256;
  const long long cap = 4LL * sm_count();  // the ordered tail reads one partial row per CTA
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  p.blocks = (int)blocks;
  p.ok = true;
  return p;
}

static int check_pool_args(const char* what, int dtype, int64_t batch, int64_t in_h, int64_t in_w,
                           int64_t channels, int64_t out_h, int64_t out_w, int window_h,
                           int window_w, int stride_h, int stride_w, int pad_top, int pad_left,
                           PoolGeom* g) {
  if (dtype != B200_DT_FLOAT && dtype != B200_DT_BFLOAT16) {
    set_last_error("%s: unsupported dtype %d", what, dtype);
    return B200_UNIMPLEMENTED;
  }
  if (batch < 0 || in_h < 0 || in_w < 0 || channels < 0 || out_h < 0 || out_w < 0 ||
      window_h <= 0 || window_w <= 0 || stride_h <= 0 || stride_w <= 0 || pad_top < 0 ||
      pad_left < 0) {
    set_last_error("%s: invalid geometry", what);
    return B200_INVALID_ARGUMENT;
  }
  if (in_h > INT32_MAX || in_w > INT32_MAX || channels > INT32_MAX || batch > INT32_MAX) {
    set_last_error("%s: dimension exceeds int32", what);
    return B200_INVALID_ARGUMENT;
  }
  g->N = (int)batch;
  g->H = (int)in_h;
  g->W = (int)in_w;
  g->C = (int)channels;
  g->OH = (int)out_h;
  g->OW = (int)out_w;
  g->wh = window_h;
  g->ww = window_w;
  g->sh = stride_h;
  g->sw = stride_w;
  g->pt = pad_top;
  g->pl = pad_left;
  return B200_OK;
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace b200

using namespace b200;

extern "C" {

int b200_max_pool(int dtype, const void* in, void* out, int64_t batch, int64_t in_h, int64_t in_w,
                  int64_t channels, int64_t out_h, int64_t out_w, int window_h, int window_w,
                  int stride_h, int stride_w, int pad_top, int pad_left, void* stream) {
  PoolGeom g;
  int rc = check_pool_args("b200_max_pool", dtype, batch, in_h, in_w, channels, out_h, out_w,
                           window_h, window_w, stride_h, stride_w, pad_top, pad_left, &g);
  if (rc) return rc;
  const long long nout = (long long)batch * out_h * out_w * channels;
  if (nout == 0) return B200_OK;
  rc = require_device("b200_max_pool");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const bool al = aligned16(in) && aligned16(out);
  if (dtype == B200_DT_FLOAT) {
    if (al && channels % 4 == 0) {
      const long long t = nout / 4;
      launch_pdl(max_pool_fwd_kernel<float, 4>, dim3((unsigned)((t + 255) / 256)), dim3(256), 0, s, 
          static_cast<const float*>(in), static_cast<float*>(out), g, t);
    } else {
      launch_pdl(max_pool_fwd_kernel<float, 1>, dim3((unsigned)((nout + 255) / 256)), dim3(256), 0, s, 
          static_cast<const float*>(in), static_cast<float*>(out), g, nout);
    }
  } else {
    if (al && channels % 8 == 0) {
      const long long t = nout / 8;
      launch_pdl(max_pool_fwd_kernel<__nv_bfloat16, 8>, dim3((unsigned)((t + 255) / 256)), dim3(256), 0, s, 
          static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), g, t);
    } else {
      launch_pdl(max_pool_fwd_kernel<__nv_bfloat16, 1>, dim3((unsigned)((nout + 255) / 256)), dim3(256), 0, s, 
          static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), g, nout);
    }
  }
  note_launch();
  return check_launch("b200_max_pool");
}

int b200_max_pool_grad(int dtype, const void* orig_in, const void* orig_out, const void* grad,
                       void* in_backprop, int64_t batch, int64_t in_h, int64_t in_w,
                       int64_t channels, int64_t out_h, int64_t out_w, int window_h, int window_w,
                       int stride_h, int stride_w, int pad_top, int pad_left, void* stream) {
  (void)orig_out;  // the reference CPU kernel recomputes the forward pass too (maxpooling_op.cc:259-262)
  PoolGeom g;
  int rc = check_pool_args("b200_max_pool_grad", dtype, batch, in_h, in_w, channels, out_h, out_w,
                           window_h, window_w, stride_h, stride_w, pad_top, pad_left, &g);
  if (rc) return rc;
  const long long nin = (long long)batch * in_h * in_w * channels;
  if (nin == 0) return B200_OK;
  rc = require_device("b200_max_pool_grad");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const bool al = aligned16(orig_in) && aligned16(grad) && aligned16(in_backprop);
  if (stride_h >= window_h && stride_w >= window_w && out_h > 0 && out_w > 0) {
    // disjoint windows: do they tile the whole input?
    const bool tiles = stride_h == window_h && stride_w == window_w && pad_top == 0 &&
                       pad_left == 0 && out_h * stride_h >= in_h && out_w * stride_w >= in_w;
    const size_t es = dtype == B200_DT_FLOAT ? 4 : 2;
    if (!tiles) {
      rc = b200_memset_async(in_backprop, 0, (size_t)nin * es, stream);
      if (rc) return rc;
    }
    const long long nout = (long long)batch * out_h * out_w * channels;
#define POOLG(T, V)                                                                            \
  do {                                                                                         \
    const long long t = nout / V;                                                              \
    launch_pdl(max_pool_grad_disjoint_kernel<T, V>, dim3((unsigned)((t + 255) / 256)), dim3(256), 0, s,            \
        static_cast<const T*>(orig_in), static_cast<const T*>(grad), static_cast<T*>(in_backprop), \
        g, t);                                                                                 \
  } while (0)
    if (dtype == B200_DT_FLOAT) {
      if (al && channels % 4 == 0) POOLG(float, 4); else POOLG(float, 1);
    } else {
      if (al && channels % 8 == 0) POOLG(__nv_bfloat16, 8); else POOLG(__nv_bfloat16, 1);
    }
#undef POOLG
    note_launch();
    return check_launch("b200_max_pool_grad");
  }
  if (dtype == B200_DT_FLOAT) {
    if (al && channels % 4 == 0) {
      const long long t = nin / 4;
      launch_pdl(max_pool_grad_kernel<float, 4>, dim3((unsigned)((t + 255) / 256)), dim3(256), 0, s, 
          static_cast<const float*>(orig_in), static_cast<const float*>(grad),
          static_cast<float*>(in_backprop), g, t);
    } else {
      launch_pdl(max_pool_grad_kernel<float, 1>, dim3((unsigned)((nin + 255) / 256)), dim3(256), 0, s, 
          static_cast<const float*>(orig_in), static_cast<const float*>(grad),
          static_cast<float*>(in_backprop), g, nin);
    }
  } else {
    if (al && channels % 8 == 0) {
      const long long t = nin / 8;
      launch_pdl(max_pool_grad_kernel<__nv_bfloat16, 8>, dim3((unsigned)((t + 255) / 256)), dim3(256), 0, s, 
          static_cast<const __nv_bfloat16*>(orig_in), static_cast<const __nv_bfloat16*>(grad),
          static_cast<__nv_bfloat16*>(in_backprop), g, t);
    } else {
      launch_pdl(max_pool_grad_kernel<__nv_bfloat16, 1>, dim3((unsigned)((nin + 255) / 256)), dim3(256), 0, s, 
          static_cast<const __nv_bfloat16*>(orig_in), static_cast<const __nv_bfloat16*>(grad),
          static_cast<__nv_bfloat16*>(in_backprop), g, nin);
    }
  }
  note_launch();
  return check_launch("b200_max_pool_grad");
}

size_t b200_max_pool_grad_relu_bias_grad_workspace_bytes(int dtype, int64_t batch, int64_t in_h,
                                                         int64_t in_w, int64_t channels,
                                                         int64_t out_h, int64_t out_w, int window_h,
                                                         int window_w, int stride_h, int stride_w,
                                                         int pad_top, int pad_left) {
  PoolGeom g;
  if (check_pool_args("b200_max_pool_grad_relu_bias_grad_workspace_bytes", dtype, batch, in_h, in_w,
                      channels, out_h, out_w, window_h, window_w, stride_h, stride_w, pad_top,
                      pad_left, &g) != B200_OK)
    return 0;
  const PoolFusedPlan p = plan_pool_fused(dtype, g, batch);
  return p.ok ? (size_t)p.blocks * (size_t)channels * sizeof(float) : 0;
}

int b200_max_pool_grad_relu_bias_grad(int dtype, const void* orig_in, const void* grad,
                                      void* backprops, void* bias_grad, int64_t batch,
                                      int64_t in_h, int64_t in_w, int64_t channels, int64_t out_h,
                                      int64_t out_w, int window_h, int window_w, int stride_h,
                                      int stride_w, int pad_top, int pad_left, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  PoolGeom g;
  int rc = check_pool_args("b200_max_pool_grad_relu_bias_grad", dtype, batch, in_h, in_w, channels,
                           out_h, out_w, window_h, window_w, stride_h, stride_w, pad_top, pad_left,
                           &g);
  if (rc) return rc;
  const PoolFusedPlan p = plan_pool_fused(dtype, g, batch);
  static const bool off = getenv("B200TF_NO_POOL_GRAD_FUSION") != nullptr;
  if (!p.ok || off || batch == 0 || !aligned16(orig_in) || !aligned16(grad) ||
      !aligned16(backprops)) {
    // the caller composes b200_max_pool_grad + b200_relu_grad_bias_grad instead
    set_last_error("b200_max_pool_grad_relu_bias_grad: geometry outside the fused kernel (windows "
                   "must tile the input, C / (16 B) a power of two <= 32)");
    return B200_UNIMPLEMENTED;
  }
  const size_t need = (size_t)p.blocks * (size_t)channels * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_last_error("b200_max_pool_grad_relu_bias_grad: workspace too small (%zu < %zu bytes)",
                   workspace_bytes, need);
    return B200_INVALID_ARGUMENT;
  }
  rc = require_device("b200_max_pool_grad_relu_bias_grad");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  static std::atomic<unsigned> next_slot{0};
  unsigned int* base = nullptr;
  if (cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_pool_fused_tickets) != cudaSuccess)
    return check_launch("b200_max_pool_grad_relu_bias_grad");
  unsigned int* ticket = base + next_slot.fetch_add(1) % kPoolFusedSlots;
  float* partial = static_cast<float*>(workspace);
  if (dtype == B200_DT_FLOAT)
    launch_pdl(max_pool_grad_relu_bias_kernel<float, 4>, dim3(p.blocks), dim3(256), 0, s,
               static_cast<const float*>(orig_in), static_cast<const float*>(grad),
               static_cast<float*>(backprops), partial, static_cast<float*>(bias_grad), g, p.total,
               p.G, ticket);
  else
    launch_pdl(max_pool_grad_relu_bias_kernel<__nv_bfloat16, 8>, dim3(p.blocks), dim3(256), 0, s,
               static_cast<const __nv_bfloat16*>(orig_in), static_cast<const __nv_bfloat16*>(grad),
               static_cast<__nv_bfloat16*>(backprops), partial,
               static_cast<__nv_bfloat16*>(bias_grad), g, p.total, p.G, ticket);
  note_launch();
  return check_launch("b200_max_pool_grad_relu_bias_grad");
}

}  // extern "C"
