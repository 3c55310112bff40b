// Conv2D / Conv2DBackpropInput / Conv2DBackpropFilter, NHWC x HWIO -> NHWC.
//
// Replaces LaunchConv2DOp<GPUDevice,T>::launch (tensorflow/core/kernels/conv_ops.cc:433-720),
// Conv2DSlowBackpropInputOp<GPUDevice,T> (conv_grad_input_ops.cc:533-917) and
// Conv2DSlowBackpropFilterOp<GPUDevice,T> (conv_grad_filter_ops.cc:361-738): cuDNN calls wrapped
// in NHWC<->NCHW and HWIO->OIHW shuffles (conv_ops_gpu_3.cu.cc).  Here everything stays NHWC:
//
//   forward   out[N*OH*OW, K]  = patches[N*OH*OW, R*S*C] . filter[R*S*C, K]
//   dFilter   dW [R*S*C, K]    = patches^T               . dY[N*OH*OW, K]      (split-K, ordered)
//   dInput    cols[N*OH*OW, R*S*C] = dY . filter^T,  then a GATHER col2im (no atomics)
//
// i.e. the im2col-GEMM formulation of the reference's CPU kernels
// (eigen_spatial_convolutions.h:1050-1067, conv_grad_filter_ops.cc:55-82,
// conv_grad_input_ops.cc:57-87) with the GEMMs on the tcgen05 kernel (gemm_tcgen05.cu).  The
// filter needs no reshuffle: HWIO is already the row-major [R*S*C, K] GEMM operand.  1x1/stride-1
// convolutions skip the patch matrix entirely (same shortcut as conv_ops.cc:454-480).
// Data paths, fastest first: unit-stride convolutions whose channel counts are whole 128-byte
// blocks run the halo-tile kernels (conv_halo.cu: forward and, with the flipped filter, dInput;
// conv_halo_wgrad.cu: the filter gradient); other strides use TMA im2col-mode loads on the
// tcgen05 GEMM (no patch matrix either); first layers use CUDA-core kernels (C = 1, unit stride:
// the lane-per-filter conv_c1_* kernels below; other C <= 4: conv_small_cin_*); everything else
// materialises the patch matrix in the caller-provided workspace.
#include <cuda_bf16.h>
#include <cstdlib>

#include "b200_internal.h"

namespace b200 {

struct ConvG {
  int N, H, W, C, R, S, K, OH, OW, sh, sw, pt, pl;
  int ldk;  // leading dimension of the patch matrix (R*S*C rounded up to 16 bytes)
};

// Patch-matrix builder.  Thread (lane-in-row) owns ONE vector slot j of the patch row, so its filter
// tap (r, s, c) is decoded once; the CTA then walks `rows_per_cta` output pixels, decoding each
// pixel's (n, oh, ow) once per row-group instead of once per element (the first version spent its
// time in ~10 integer divisions per 16 bytes).  Rows are written fully coalesced.
template <typename T, int V>
__global__ void __launch_bounds__(256)
im2col_kernel(const T* __restrict__ in, T* __restrict__ col, ConvG g, long long rows,
              int threads_per_row, int rows_per_cta) {
  pdl_prologue();
  const int lane_in_row = threadIdx.x % threads_per_row;
  const int row_in_group = threadIdx.x / threads_per_row;
  const int groups = 256 / threads_per_row;  // pixel rows handled concurrently by the CTA
  const int vecs_per_row = g.ldk / V;
  const int rsc = g.R * g.S * g.C;
  const long long row_begin = (long long)blockIdx.x * rows_per_cta;
  long long row_end = row_begin + rows_per_cta;
  if (row_end > rows) row_end = rows;
  for (int jv = lane_in_row; jv < vecs_per_row; jv += threads_per_row) {
    const int j = jv * V;
    const bool pad_slot = j >= rsc;
    const int c = pad_slot ? 0 : j % g.C;
    const int tap = pad_slot ? 0 : j / g.C;
    const int fs = tap % g.S, fr = tap / g.S;
    // (n, oh, ow) of the first pixel by division once, then advanced incrementally
    long long p = row_begin + row_in_group;
    int ow = (int)(p % g.OW);
    int oh = (int)((p / g.OW) % g.OH);
    int n = (int)(p / ((long long)g.OW * g.OH));
    for (; p < row_end; p += groups) {
      const int ih = oh * g.sh - g.pt + fr, iw = ow * g.sw - g.pl + fs;
      T* dst = col + p * g.ldk + j;
      const bool inside = !pad_slot && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
      if (V * sizeof(T) == 16) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (inside)
          v = __ldg(reinterpret_cast<const uint4*>(
              in + (((long long)n * g.H + ih) * g.W + iw) * g.C + c));
        *reinterpret_cast<uint4*>(dst) = v;
      } else {
        const T* src = in + (((long long)n * g.H + ih) * g.W + iw) * g.C + c;
#pragma unroll
        for (int e = 0; e < V; ++e) dst[e] = inside ? src[e] : T(0.f);
      }
      ow += groups;
      while (ow >= g.OW) {
        ow -= g.OW;
        if (++oh == g.OH) {
          oh = 0;
          ++n;
        }
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ float cvt_in(T v);
template <>
__device__ __forceinline__ float cvt_in<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ float cvt_in<__nv_bfloat16>(__nv_bfloat16 v) {
  return __bfloat162float(v);
}
template <typename T>
__device__ __forceinline__ T cvt_out(float v);
template <>
__device__ __forceinline__ float cvt_out<float>(float v) {
  return v;
}
template <>
__device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// Gather col2im: thread per input element (n, h, w, c); sums the patch-matrix entries that map
// to it, visiting output pixels in ascending (oh, ow) order like Col2im
// (conv_grad_input_ops.cc:57-87) does.
template <typename T>
__global__ void __launch_bounds__(256)
col2im_gather_kernel(const T* __restrict__ col, T* __restrict__ dx, ConvG g, long long total) {
  pdl_prologue();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % g.C);
  long long q = idx / g.C;
  const int w = (int)(q % g.W);
  q /= g.W;
  const int h = (int)(q % g.H);
  const int n = (int)(q / g.H);
  float acc = 0.f;
  // oh * sh - pt + r == h  with 0 <= r < R
  const int hp = h + g.pt, wp = w + g.pl;
  int oh_lo = hp - (g.R - 1);
  oh_lo = oh_lo <= 0 ? 0 : (oh_lo + g.sh - 1) / g.sh;
  const int oh_hi = min(hp / g.sh, g.OH - 1);
  int ow_lo = wp - (g.S - 1);
  ow_lo = ow_lo <= 0 ? 0 : (ow_lo + g.sw - 1) / g.sw;
  const int ow_hi = min(wp / g.sw, g.OW - 1);
  for (int oh = oh_lo; oh <= oh_hi; ++oh) {
    const int r = hp - oh * g.sh;
    for (int ow = ow_lo; ow <= ow_hi; ++ow) {
      const int s = wp - ow * g.sw;
      const long long p = ((long long)n * g.OH + oh) * g.OW + ow;
      acc += cvt_in<T>(col[p * g.ldk + (r * g.S + s) * g.C + c]);
    }
  }
  dx[idx] = cvt_out<T>(acc);
}

// dInput as a forward convolution of dY: filter'[r', s', k, c] = filter[R-1-r', S-1-s', c, k]
// (spatially flipped, in/out channels swapped).  R*S*C*K elements: negligible next to the conv.
template <typename T>
__global__ void __launch_bounds__(256)
flip_filter_kernel(const T* __restrict__ w, T* __restrict__ wt, int R, int S, int C, int K) {
  pdl_prologue();
  const int i = blockIdx.x * 256 + threadIdx.x;  // index into wt [R, S, K, C]
  if (i >= R * S * C * K) return;
  const int c = i % C;
  int t = i / C;
  const int k = t % K;
  t /= K;
  const int s2 = t % S, r2 = t / S;
  wt[i] = w[(((R - 1 - r2) * S + (S - 1 - s2)) * C + c) * K + k];
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline size_t esize_of(int dtype) { return dtype == B200_DT_FLOAT ? 4 : 2; }
static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static int to_geom(const char* what, int dtype, const b200_conv2d_geometry* in, ConvG* g) {
  if (dtype != B200_DT_FLOAT && dtype != B200_DT_BFLOAT16) {
    set_last_error("%s: unsupported dtype %d", what, dtype);
    return B200_UNIMPLEMENTED;
  }
  if (!in) {
    set_last_error("%s: null geometry", what);
    return B200_INVALID_ARGUMENT;
  }
  if (in->batch < 0 || in->in_h < 0 || in->in_w < 0 || in->in_c < 0 || in->filter_h <= 0 ||
      in->filter_w <= 0 || in->out_c < 0 || in->out_h < 0 || in->out_w < 0 || in->stride_h <= 0 ||
      in->stride_w <= 0 || in->pad_top < 0 || in->pad_left < 0) {
    set_last_error("%s: invalid geometry", what);
    return B200_INVALID_ARGUMENT;
  }
  const int64_t lim = INT32_MAX;
  if (in->batch > lim || in->in_h > lim || in->in_w > lim || in->in_c > lim || in->out_c > lim ||
      in->filter_h * in->filter_w * in->in_c > lim) {
    set_last_error("%s: dimension exceeds int32", what);
    return B200_INVALID_ARGUMENT;
  }
  g->N = (int)in->batch;
  g->H = (int)in->in_h;
  g->W = (int)in->in_w;
  g->C = (int)in->in_c;
  g->R = (int)in->filter_h;
  g->S = (int)in->filter_w;
  g->K = (int)in->out_c;
  g->OH = (int)in->out_h;
  g->OW = (int)in->out_w;
  g->sh = in->stride_h;
  g->sw = in->stride_w;
  g->pt = in->pad_top;
  g->pl = in->pad_left;
  const int a = (int)(16 / esize_of(dtype));
  const int rsc = g->R * g->S * g->C;
  g->ldk = (rsc + a - 1) / a * a;
  return B200_OK;
}

static bool is_pointwise(const ConvG& g) {
  return g.R == 1 && g.S == 1 && g.sh == 1 && g.sw == 1 && g.pt == 0 && g.pl == 0 &&
         g.OH == g.H && g.OW == g.W;
}

template <typename T>
static int run_im2col(const void* in, void* col, const ConvG& g, cudaStream_t s) {
  constexpr int V16 = 16 / sizeof(T);
  const long long rows = (long long)g.N * g.OH * g.OW;
  auto plan = [&](int vecs_per_row, int* tpr, int* rpc) {
    int t = 32;  // threads per patch row: power of two in [32, 256]
    while (t < vecs_per_row && t < 256) t <<= 1;
    *tpr = t;
    const int groups = 256 / t;
    // ~4 waves of CTAs over the SMs, at least one row per group
    long long r = (rows + 4LL * sm_count() - 1) / (4LL * sm_count());
    r = (r + groups - 1) / groups * groups;
    if (r < groups) r = groups;
    *rpc = (int)r;
  };
  int tpr, rpc;
  if (g.C % V16 == 0 && aligned16(in) && aligned16(col)) {
    plan(g.ldk / V16, &tpr, &rpc);
    launch_pdl(im2col_kernel<T, V16>, dim3((unsigned)((rows + rpc - 1) / rpc)), dim3(256), 0, s, 
        static_cast<const T*>(in), static_cast<T*>(col), g, rows, tpr, rpc);
  } else {
    plan(g.ldk, &tpr, &rpc);
    launch_pdl(im2col_kernel<T, 1>, dim3((unsigned)((rows + rpc - 1) / rpc)), dim3(256), 0, s, 
        static_cast<const T*>(in), static_cast<T*>(col), g, rows, tpr, rpc);
  }
  note_launch();
  return check_launch("im2col");
}


// ------------------------------------------------------------------ first layers (C_in <= 4)
// A network's first convolution (LeNet conv1: C=1, VGG conv1_1: C=3) has K_gemm = R*S*C of a few
// dozen: as a GEMM it would pad K to the MMA granularity AND need a patch matrix because TMA
// im2col wants whole 128-byte channel blocks.  It is an HBM-bound streaming problem instead
// (conv1 of LeNet at batch 512: 1.6 MB in, 51 MB out, 0.64 GFLOP), done on the CUDA cores in
// IEEE fp32: the filter lives in shared memory, 8 consecutive threads produce the K outputs of
// one pixel (4 each, so a pixel's row of K floats is written as one coalesced run) and the taps
// of a pixel are broadcast loads out of L1.
constexpr int kSmallCinThreads = 256;
struct SmallCinFit {
  bool ok;
  int taps;  // R * S * C
};
static SmallCinFit small_cin_fit(int dtype, const ConvG& g) {
  SmallCinFit f{false, g.R * g.S * g.C};
  f.ok = dtype == B200_DT_FLOAT && g.C >= 1 && g.C <= 4 && g.K % 4 == 0 && g.K >= 4 && g.K <= 128 &&
         (size_t)f.taps * g.K * sizeof(float) <= 40 * 1024 && f.taps <= 200 &&
         (long long)g.N * g.OH * g.OW < (1LL << 30) && (long long)g.N * g.H * g.W * g.C < (1LL << 31);
  return f;
}

// kR/kS/kC > 0: compile-time filter geometry (taps fully unrolled); 0: read it from `g`.
template <int kR, int kS, int kC, int kKPerThread>
__global__ void __launch_bounds__(kSmallCinThreads)
conv_small_cin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w,
                          float* __restrict__ y, ConvG g, int pixels,
                          const float* __restrict__ bias, int relu) {
  pdl_prologue();
  extern __shared__ float wsm[];  // [taps][K]
  const int R = kR ? kR : g.R, S = kS ? kS : g.S, C = kC ? kC : g.C;
  const int taps = R * S * C;
  for (int i = threadIdx.x; i < taps * g.K; i += kSmallCinThreads) wsm[i] = __ldg(w + i);
  __syncthreads();
  constexpr int KP = kKPerThread;            // filters per thread: one input load feeds KP FMAs
  const int kgroups = g.K / KP;
  const int ppc = kSmallCinThreads / kgroups;  // pixels per CTA pass
  const int kg = threadIdx.x % kgroups, pl = threadIdx.x / kgroups;
  if (pl >= ppc) return;
  const float* wk = wsm + kg * KP;
  for (int p = blockIdx.x * ppc + pl; p < pixels; p += gridDim.x * ppc) {
    const int ow = p % g.OW;
    const int t = p / g.OW;
    const int oh = t % g.OH;
    const int n = t / g.OH;
    float acc[KP];
#pragma unroll
    for (int j = 0; j < KP; ++j) acc[j] = 0.f;
    const float* xn = x + (long long)n * g.H * g.W * C;
    const int ih0 = oh * g.sh - g.pt, iw0 = ow * g.sw - g.pl;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int ih = ih0 + r;
      const bool row_in = ih >= 0 && ih < g.H;
#pragma unroll
      for (int sx = 0; sx < S; ++sx) {
        const int iw = iw0 + sx;
        const bool in = row_in && iw >= 0 && iw < g.W;
        const float* xp = xn + (ih * g.W + iw) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float v = in ? __ldg(xp + c) : 0.f;
          const float* wt = wk + ((r * S + sx) * C + c) * g.K;
#pragma unroll
          for (int q = 0; q < KP / 4; ++q) {
            const float4 wv = *reinterpret_cast<const float4*>(wt + 4 * q);
            acc[4 * q] = fmaf(v, wv.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(v, wv.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, wv.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, wv.w, acc[4 * q + 3]);
          }
        }
      }
    }
    if (bias != nullptr) {  // fused BiasAdd (+ Relu): same fp32 add / max as the separate kernels
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        acc[j] += __ldg(bias + kg * KP + j);
        if (relu) acc[j] = acc[j] > 0.f ? acc[j] : 0.f;
      }
    }
    float4* out = reinterpret_cast<float4*>(y + (long long)p * g.K + kg * KP);
#pragma unroll
    for (int q = 0; q < KP / 4; ++q)
      out[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
  }
}

// Filter gradient of a first layer: dW[tap, k] = sum over pixels of patch[p, tap] * dY[p, k].
// The kernel is bound by bytes in flight on the dY stream (51 MB for LeNet conv1), so dY is
// staged: a 64-thread CTA cp.async's tiles of 64 pixels x 32 filters (8 KB, double-buffered) into
// shared memory; thread (tap = tid / 2, half = tid % 2) then owns 16 accumulators (one tap x 16
// filters) for the whole kernel -- per pixel one gathered input value, four broadcast 16-byte
// shared loads, 16 FMAs -- so no reduction is needed inside the CTA and the pixel order is fixed.
// CTAs write partial [taps][K] tiles that a second kernel adds in CTA order: no atomics,
// reproducible.  K % 32 == 0; more than 32 taps / 32 filters are further passes over dY.
constexpr int kWgradThreads = 64, kWgradTilePix = 64;
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__global__ void __launch_bounds__(kWgradThreads)
conv_small_cin_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                            float* __restrict__ partial, ConvG g, int pixels) {
  pdl_prologue();
  __shared__ __align__(16) float tile[2][kWgradTilePix][32];
  const int taps = g.R * g.S * g.C;
  const int t = threadIdx.x >> 1, half = threadIdx.x & 1;
  const int ntiles = (pixels + kWgradTilePix - 1) / kWgradTilePix;
  for (int tb = 0; tb * 32 < taps; ++tb) {
    const int tap = tb * 32 + t;
    const bool tap_ok = tap < taps;
    const int c = tap % g.C, rs = tap / g.C;
    const int sx = rs % g.S, r = rs / g.S;
    for (int kb = 0; kb * 32 < g.K; ++kb) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      auto load_tile = [&](int ti, int buf) {
        for (int i = threadIdx.x; i < kWgradTilePix * 8; i += kWgradThreads) {
          const int pix = i >> 3, q = i & 7;
          const int p = ti * kWgradTilePix + pix;
          float* dst = &tile[buf][pix][q * 4];
          if (p < pixels)
            cp_async16(dst, dy + (long long)p * g.K + kb * 32 + q * 4);
          else
            *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      int buf = 0;
      int ti = blockIdx.x;
      if (ti < ntiles) load_tile(ti, 0);
      for (; ti < ntiles; ti += gridDim.x) {
        const int nxt = ti + gridDim.x;
        if (nxt < ntiles) {
          load_tile(nxt, buf ^ 1);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        const int p0 = ti * kWgradTilePix;
        int ow = p0 % g.OW, tt = p0 / g.OW;
        int oh = tt % g.OH, n = tt / g.OH;
        const int npix = min(kWgradTilePix, pixels - p0);
        for (int pix = 0; pix < npix; ++pix) {
          const int ih = oh * g.sh - g.pt + r, iw = ow * g.sw - g.pl + sx;
          const bool in = tap_ok && ih >= 0 && ih < g.H && iw >= 0 && iw < g.W;
          const float v = in ? __ldg(x + (((long long)n * g.H + ih) * g.W + iw) * g.C + c) : 0.f;
          const float4* d = reinterpret_cast<const float4*>(&tile[buf][pix][half * 16]);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 dq = d[q];
            acc[4 * q] = fmaf(v, dq.x, acc[4 * q]);
            acc[4 * q + 1] = fmaf(v, dq.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, dq.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, dq.w, acc[4 * q + 3]);
          }
          if (++ow == g.OW) {  // next pixel, without divisions
            ow = 0;
            if (++oh == g.OH) {
              oh = 0;
              ++n;
            }
          }
        }
        __syncthreads();  // tile consumed: the next trip's prefetch may overwrite the other buffer
        buf ^= 1;
      }
      if (tap_ok) {
        float4* dst = reinterpret_cast<float4*>(partial + ((long long)blockIdx.x * taps + tap) * g.K +
                                                kb * 32 + half * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          dst[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
      }
    }
  }
}
// dW[i] = sum over CTAs (fixed order) of partial[cta][i]: a CTA of (32, strands) threads owns 32
// elements; strand q adds every strands-th partial in ascending order (four loads in flight), the
// strands are merged in strand order.  blockDim.y = 8 or 32 strands.
__global__ void __launch_bounds__(1024)
conv_small_cin_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                   int blocks, int elems) {
  pdl_prologue();
  __shared__ float strands[32][33];
  const int x = threadIdx.x & 31;
  const int st = blockDim.y == 1 ? (int)(threadIdx.x >> 5) : (int)threadIdx.y;
  const int nst = blockDim.y == 1 ? (int)(blockDim.x >> 5) : (int)blockDim.y;
  const int e = blockIdx.x * 32 + x;
  float sum = 0.f;
  if (e < elems) {
    int b = st;
    for (; b + 3 * nst < blocks; b += 4 * nst) {
      const float v0 = __ldcg(partial + (long long)b * elems + e);
      const float v1 = __ldcg(partial + (long long)(b + nst) * elems + e);
      const float v2 = __ldcg(partial + (long long)(b + 2 * nst) * elems + e);
      const float v3 = __ldcg(partial + (long long)(b + 3 * nst) * elems + e);
      sum += v0;
      sum += v1;
      sum += v2;
      sum += v3;
    }
    for (; b < blocks; b += nst) sum += __ldcg(partial + (long long)b * elems + e);
  }
  strands[st][x] = sum;
  __syncthreads();
  if (st == 0 && e < elems) {
    float t = 0.f;
    for (int q = 0; q < nst; ++q) t += strands[q][x];
    dw[e] = t;
  }
}
static int small_cin_wgrad_blocks() { return 8 * sm_count(); }

// ------------------------------------------------------------------ one input channel, unit stride
// LeNet conv1 (28x28x1 * 5x5x1x32, batch 512: 1.6 MB in, 51 MB out / dY, 0.64 GFLOP) is an HBM
// streaming problem; the generic small-C_in kernels above were bound by shared-memory loads instead
// (16-byte filter / dY broadcasts: 4 LDS.128 per 16 FMAs).  Here the lane IS the filter: a lane
// keeps its filter's R*S taps (forward) or its R*S accumulators (filter gradient) in registers, a
// warp walks one output row, and the R x S input window -- the same for every lane -- slides over
// a zero-padded copy of the image in shared memory: S..R new broadcast 4-byte loads per pixel, the
// window rotates through registers (pixel loop unrolled by S so every index is static).
// Per pixel and warp: R loads + R*S FMAs + one coalesced 128-byte store (forward) or one
// conflict-free shared load of the staged dY row (gradient).  Same tap order and fp32 FMAs as the
// generic kernels (bit-identical forward).
constexpr int kC1Threads = 256, kC1Warps = 8;

// Tile row stride (floats): OW + S - 1 columns rounded up to a multiple of 4; everything outside
// the image is zero.
__host__ __device__ inline int c1_row_stride(int OW) { return (OW + 4 + 3) / 4 * 4; }

template <int R, int S>
__device__ __forceinline__ void c1_stage_image(float* xs, const float* __restrict__ xn,
                                               const ConvG& g, int HP, int WPs) {
  for (int i = threadIdx.x; i < HP * WPs; i += kC1Threads) {
    const int ih = i / WPs - g.pt, iw = i % WPs - g.pl;
    xs[i] = (ih >= 0 && ih < g.H && iw >= 0 && iw < g.W) ? __ldg(xn + ih * g.W + iw) : 0.f;
  }
}

template <int R, int S>
__global__ void __launch_bounds__(kC1Threads)
conv_c1_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, float* __restrict__ y,
                   ConvG g, const float* __restrict__ bias, int relu, int parts) {
  pdl_prologue();
  extern __shared__ __align__(16) float c1_smem[];
  float* xs = c1_smem;  // [HP][WP], zero padded
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HP = g.OH + R - 1, WPs = c1_row_stride(g.OW);
  // work item = (image, band of output rows): `parts` bands per image keep the last wave of CTAs
  // short when the batch is only a few images per SM
  for (int item = blockIdx.x; item < g.N * parts; item += gridDim.x) {
    const int n = item / parts, part = item - n * parts;
    const int oh_begin = g.OH * part / parts, oh_end = g.OH * (part + 1) / parts;
    __syncthreads();  // the previous item's readers are done
    c1_stage_image<R, S>(xs, x + (long long)n * g.H * g.W, g, HP, WPs);
    __syncthreads();
    for (int kb = 0; kb < g.K; kb += 32) {
      float wr[R * S];
#pragma unroll
      for (int t = 0; t < R * S; ++t) wr[t] = __ldg(w + t * g.K + kb + lane);
      const float bv = bias != nullptr ? __ldg(bias + kb + lane) : 0.f;
      for (int oh = oh_begin + warp; oh < oh_end; oh += kC1Warps) {
        // rp[r]: the window's NEW column (tile column ow + S - 1) in row r; one increment per row
        // and group of S pixels, everything else is an immediate offset
        const float* rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rp[r] = xs + (oh + r) * WPs + (S - 1);
        float win[R][S];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < S - 1; ++c) win[r][c] = rp[r][c - (S - 1)];
        float* yp = y + (((long long)n * g.OH + oh) * g.OW) * g.K + kb + lane;
        const int Kst = g.K;
        auto pixel = [&](int j) {  // j = ow % S, a compile-time constant at every call site
#pragma unroll
          for (int r = 0; r < R; ++r) win[r][(j + S - 1) % S] = rp[r][j];
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int sx = 0; sx < S; ++sx) acc = fmaf(win[r][(j + sx) % S], wr[r * S + sx], acc);
          if (bias != nullptr) {  // fused BiasAdd (+ Relu), as in conv_small_cin_fwd_kernel
            acc += bv;
            if (relu) acc = acc > 0.f ? acc : 0.f;
          }
          yp[j * Kst] = acc;
        };
        int ow0 = 0;
        for (; ow0 + S <= g.OW; ow0 += S) {  // whole groups: no per-pixel predicate
#pragma unroll
          for (int j = 0; j < S; ++j) pixel(j);
#pragma unroll
          for (int r = 0; r < R; ++r) rp[r] += S;
          yp += S * Kst;
        }
#pragma unroll
        for (int j = 0; j < S - 1; ++j)
          if (ow0 + j < g.OW) pixel(j);
      }
    }
  }
}

// Filter gradient: dW[r, s, k] = sum over pixels of x[oh + r - pt, ow + s - pl] * dY[pixel, k].
// Every warp stages the dY rows it walks through its own double-buffered cp.async ring (a whole
// output row of 32 filters in flight per warp), accumulates R*S taps per lane over all its rows and
// images, the warps are merged in warp order and each CTA writes one partial [R*S][K] tile that
// conv_small_cin_wgrad_reduce_kernel adds in CTA order.  No atomics; reproducible.
template <int R, int S>
__global__ void __launch_bounds__(kC1Threads)
conv_c1_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                     float* __restrict__ partial, ConvG g, int parts) {
  pdl_prologue();
  extern __shared__ __align__(16) float c1_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int HP = g.OH + R - 1, WPs = c1_row_stride(g.OW);
  const int xs_floats = HP * WPs;
  const int row_floats = g.OW * 32;
  float* xs = c1_smem;
  float* my = c1_smem + xs_floats + warp * 2 * row_floats;  // this warp's ring: 2 dY rows
  for (int kb = 0; kb < g.K; kb += 32) {
    float acc[R * S];
#pragma unroll
    for (int t = 0; t < R * S; ++t) acc[t] = 0.f;
    for (int item = blockIdx.x; item < g.N * parts; item += gridDim.x) {
      const int n = item / parts, part = item - n * parts;
      const int oh_begin = g.OH * part / parts, oh_end = g.OH * (part + 1) / parts;
      __syncthreads();
      c1_stage_image<R, S>(xs, x + (long long)n * g.H * g.W, g, HP, WPs);
      __syncthreads();
      const float* dyn = dy + (long long)n * g.OH * g.OW * g.K + kb;
      auto fetch = [&](int oh, int st) {
        float* dst = my + st * row_floats;
        const float* src = dyn + (long long)oh * g.OW * g.K;
        for (int i = lane; i < g.OW * 8; i += 32)
          cp_async16(dst + (i >> 3) * 32 + (i & 7) * 4, src + (long long)(i >> 3) * g.K + (i & 7) * 4);
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      int st = 0;
      if (oh_begin + warp < oh_end) fetch(oh_begin + warp, 0);
      for (int oh = oh_begin + warp; oh < oh_end; oh += kC1Warps) {
        if (oh + kC1Warps < oh_end) {
          fetch(oh + kC1Warps, st ^ 1);
          asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
          asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        const float* d = my + st * row_floats + lane;
        const float* rp[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rp[r] = xs + (oh + r) * WPs + (S - 1);
        float win[R][S];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int c = 0; c < S - 1; ++c) win[r][c] = rp[r][c - (S - 1)];
        auto pixel = [&](int j) {
#pragma unroll
          for (int r = 0; r < R; ++r) win[r][(j + S - 1) % S] = rp[r][j];
          const float dv = d[j * 32];
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int sx = 0; sx < S; ++sx)
              acc[r * S + sx] = fmaf(win[r][(j + sx) % S], dv, acc[r * S + sx]);
        };
        int ow0 = 0;
        for (; ow0 + S <= g.OW; ow0 += S) {
#pragma unroll
          for (int j = 0; j < S; ++j) pixel(j);
#pragma unroll
          for (int r = 0; r < R; ++r) rp[r] += S;
          d += S * 32;
        }
#pragma unroll
        for (int j = 0; j < S - 1; ++j)
          if (ow0 + j < g.OW) pixel(j);
        __syncwarp();  // row consumed: the next trip's prefetch may overwrite this stage
        st ^= 1;
      }
    }
    __syncthreads();
    float* red = c1_smem;  // [warps][R*S][32] over the (now idle) image and rings
#pragma unroll
    for (int t = 0; t < R * S; ++t) red[(warp * R * S + t) * 32 + lane] = acc[t];
    __syncthreads();
    for (int i = threadIdx.x; i < R * S * 32; i += kC1Threads) {
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < kC1Warps; ++q) sum += red[q * R * S * 32 + i];
      partial[((long long)blockIdx.x * R * S + (i >> 5)) * g.K + kb + (i & 31)] = sum;
    }
    __syncthreads();
  }
}

struct C1Fit {
  bool fwd, wgrad;
  size_t smem_fwd, smem_wgrad;
};
static C1Fit c1_fit(int dtype, const ConvG& g) {
  C1Fit f{false, false, 0, 0};
  if (dtype != B200_DT_FLOAT || g.C != 1 || g.sh != 1 || g.sw != 1 || g.K % 32 != 0 || g.K <= 0)
    return f;
  if (!((g.R == 5 && g.S == 5) || (g.R == 3 && g.S == 3))) return f;
  if (g.OH < 1 || g.OW < 1 || (long long)g.N * g.OH * g.OW * g.K >= (1LL << 40)) return f;
  const size_t image = (size_t)(g.OH + g.R - 1) * c1_row_stride(g.OW);
  f.smem_fwd = image * sizeof(float);
  f.fwd = f.smem_fwd <= 64 * 1024;
  const size_t red = (size_t)kC1Warps * g.R * g.S * 32 * sizeof(float);
  f.smem_wgrad = std::max(image * sizeof(float) +
                              (size_t)kC1Warps * 2 * g.OW * 32 * sizeof(float), red);
  f.wgrad = f.smem_wgrad <= 100 * 1024;
  return f;
}

static GemmArgs base_gemm(int dtype) {
  GemmArgs a{};
  a.dtype = dtype;
  a.batch = 1;
  return a;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_conv2d_workspace_bytes(int dtype, const b200_conv2d_geometry* geom, int which) {
  ConvG g;
  if (to_geom("b200_conv2d_workspace_bytes", dtype, geom, &g) != B200_OK) return 0;
  const size_t es = esize_of(dtype);
  const long long rows = (long long)g.N * g.OH * g.OW;
  const long long rsc = (long long)g.R * g.S * g.C;
  const size_t col = is_pointwise(g) ? 0 : align256((size_t)rows * g.ldk * es);
  const SmallCinFit small = small_cin_fit(dtype, g);
  static const bool no_direct_ws = getenv("B200TF_CONV_NO_DIRECT") != nullptr;
  if (small.ok && !no_direct_ws && !is_pointwise(g)) {
    if (which == 0) return 0;
    if (which == 2 && g.K % 32 == 0)
      return align256((size_t)small_cin_wgrad_blocks() * small.taps * g.K * sizeof(float));
  }
  if (which == 0) {
    // implicit GEMM needs no patch matrix (decided again at launch from the real pointers)
    ConvAOperand ca{reinterpret_cast<const void*>(16), g.N, g.H, g.W, g.C, g.R, g.S, g.OH, g.OW,
                    g.sh, g.sw, g.pt, g.pl};
    if (!is_pointwise(g) && getenv("B200TF_CONV_EXPLICIT") == nullptr &&
        b200_get_matmul_precision() == 0 && (g.K % (16 / (int)es)) == 0 &&
        conv_a_supported(dtype, ca))
      return 0;
    return col + align256(gemm_workspace_bytes(dtype, rows, g.K, rsc, 1));
  }
  if (which == 1) {
    // stride-1 dInput = implicit forward conv of dY with the flipped filter: scratch = that filter
    ConvAOperand cd{reinterpret_cast<const void*>(16), g.N, g.OH, g.OW, g.K, g.R, g.S, g.H, g.W,
                    1, 1, g.R - 1 - g.pt, g.S - 1 - g.pl};
    if (!is_pointwise(g) && g.sh == 1 && g.sw == 1 && getenv("B200TF_CONV_EXPLICIT") == nullptr &&
        b200_get_matmul_precision() == 0 && (g.C % (16 / (int)es)) == 0 &&
        g.R - 1 - g.pt >= 0 && g.S - 1 - g.pl >= 0 && conv_a_supported(dtype, cd))
      return align256((size_t)rsc * g.K * es);
    return col + align256(gemm_workspace_bytes(dtype, rows, rsc, g.K, 1));
  }
  if (which == 2) {
    if (!is_pointwise(g) && g.sh == 1 && g.sw == 1 && getenv("B200TF_CONV_EXPLICIT") == nullptr &&
        b200_get_matmul_precision() == 0) {
      // halo-tile filter gradient: one fp32 partial per CTA (conv_halo_wgrad.cu)
      ConvHaloArgs h{reinterpret_cast<const void*>(16), reinterpret_cast<const void*>(16),
                     reinterpret_cast<void*>(16), nullptr, false, g.N, g.H, g.W, g.C, g.K,
                     g.R, g.S, g.pt, g.pl, g.OH, g.OW};
      if (conv_halo_wgrad_supported(dtype, h))
        return align256(conv_halo_wgrad_workspace_bytes(dtype, h));
    }
    ConvAOperand ca{reinterpret_cast<const void*>(16), g.N, g.H, g.W, g.C, g.R, g.S, g.OH, g.OW,
                    g.sh, g.sw, g.pt, g.pl};
    if (!is_pointwise(g) && getenv("B200TF_CONV_EXPLICIT") == nullptr &&
        b200_get_matmul_precision() == 0 && (g.K % (16 / (int)es)) == 0 &&
        conv_a_supported(dtype, ca))
      return align256(gemm_workspace_bytes(dtype, rsc, g.K, rows, 1));  // split-K partials only
    return col + align256(gemm_workspace_bytes(dtype, rsc, g.K, rows, 1));
  }
  return 0;
}

// Forward convolution with an optional fused tail  out = [relu](conv + bias[k]).  The halo-tile
// kernel and the first-layer kernel apply the tail in their epilogues; every other path runs the
// convolution and then the library's own BiasAdd / Relu kernels in place (same arithmetic).
static int conv2d_impl(int dtype, const void* input, const void* filter, void* output,
                       const b200_conv2d_geometry* geom, void* workspace, size_t workspace_bytes,
                       void* stream, const void* bias, bool relu);

int b200_conv2d(int dtype, const void* input, const void* filter, void* output,
                const b200_conv2d_geometry* geom, void* workspace, size_t workspace_bytes,
                void* stream) {
  return conv2d_impl(dtype, input, filter, output, geom, workspace, workspace_bytes, stream,
                     nullptr, false);
}

int b200_fused_conv2d(int dtype, const void* input, const void* filter, const void* bias, int relu,
                      void* output, const b200_conv2d_geometry* geom, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (bias == nullptr && relu) {
    set_last_error("b200_fused_conv2d: relu without bias is not a fusion the executor produces");
    return B200_INVALID_ARGUMENT;
  }
  return conv2d_impl(dtype, input, filter, output, geom, workspace, workspace_bytes, stream, bias,
                     relu != 0);
}

static int conv2d_tail(int dtype, void* output, long long rows, int K, const void* bias, bool relu,
                       void* stream) {
  if (bias == nullptr) return B200_OK;
  int rc = b200_bias_add(dtype, output, bias, output, rows, K, stream);
  if (rc == B200_OK && relu) rc = b200_relu(dtype, output, output, rows * K, stream);
  return rc;
}

static int conv2d_impl(int dtype, const void* input, const void* filter, void* output,
                       const b200_conv2d_geometry* geom, void* workspace, size_t workspace_bytes,
                       void* stream, const void* bias, bool relu) {
  ConvG g;
  int rc = to_geom("b200_conv2d", dtype, geom, &g);
  if (rc) return rc;
  const long long rows = (long long)g.N * g.OH * g.OW;
  if (rows == 0 || g.K == 0) return B200_OK;  // conv_ops.cc:357-359
  rc = require_device("b200_conv2d");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const size_t es = esize_of(dtype);
  const long long rsc = (long long)g.R * g.S * g.C;
  if (rsc == 0) {
    rc = b200_memset_async(output, 0, (size_t)rows * g.K * es, stream);
    return rc ? rc : conv2d_tail(dtype, output, rows, g.K, bias, relu, stream);
  }
  const size_t need = b200_conv2d_workspace_bytes(dtype, geom, 0);
  if (need > 0 && (!workspace || workspace_bytes < need)) {
    set_last_error("b200_conv2d: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return B200_INVALID_ARGUMENT;
  }
  static const bool no_direct = getenv("B200TF_CONV_NO_DIRECT") != nullptr;
  static const bool no_c1 = getenv("B200TF_CONV_NO_C1") != nullptr;
  const SmallCinFit small = small_cin_fit(dtype, g);
  const C1Fit c1 = c1_fit(dtype, g);
  if (c1.fwd && !no_direct && !no_c1) {
    // one image per CTA trip; 3 CTAs per SM by registers
    const int parts = (g.N < 8 * sm_count() && g.OH >= 2 * kC1Warps) ? 2 : 1;
    const int blocks = std::min(g.N * parts, 8 * sm_count());
    const float* xin = static_cast<const float*>(input);
    const float* win = static_cast<const float*>(filter);
    float* yout = static_cast<float*>(output);
#define C1_FWD(R_, S_)                                                                          \
  do {                                                                                          \
    static const cudaError_t attr = cudaFuncSetAttribute(                                       \
        conv_c1_fwd_kernel<R_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);    \
    (void)attr;                                                                                 \
    launch_pdl(conv_c1_fwd_kernel<R_, S_>, dim3((unsigned)blocks), dim3(kC1Threads),            \
               c1.smem_fwd, s, xin, win, yout, g, static_cast<const float*>(bias),              \
               relu ? 1 : 0, parts);                                                            \
  } while (0)
    if (g.R == 5)
      C1_FWD(5, 5);
    else
      C1_FWD(3, 3);
#undef C1_FWD
    note_launch();
    return check_launch("b200_conv2d");
  }
  if (small.ok && !no_direct && !is_pointwise(g) && aligned16(output)) {
    // 16 filters per thread when K allows it (one input load feeds 16 FMAs), else 4
    const int kp = g.K % 16 == 0 ? 16 : 4;
    const int ppc = kSmallCinThreads / (g.K / kp);
    long long blocks = (rows + ppc - 1) / ppc;
    if (blocks > 8LL * sm_count()) blocks = 8LL * sm_count();
    const size_t smem = (size_t)small.taps * g.K * sizeof(float);
    const float* xin = static_cast<const float*>(input);
    const float* win = static_cast<const float*>(filter);
    float* yout = static_cast<float*>(output);
    const int pixels = (int)rows;
#define SMALL_FWD(R_, S_, C_, KP_)                                                                \
  launch_pdl(conv_small_cin_fwd_kernel<R_, S_, C_, KP_>, dim3((unsigned)blocks),                  \
             dim3(kSmallCinThreads), smem, s, xin, win, yout, g, pixels,                          \
             static_cast<const float*>(bias), relu ? 1 : 0)
    if (kp == 16 && g.R == 5 && g.S == 5 && g.C == 1)
      SMALL_FWD(5, 5, 1, 16);
    else if (kp == 16 && g.R == 3 && g.S == 3 && g.C == 3)
      SMALL_FWD(3, 3, 3, 16);
    else if (kp == 16)
      SMALL_FWD(0, 0, 0, 16);
    else
      SMALL_FWD(0, 0, 0, 4);
#undef SMALL_FWD
    note_launch();
    return check_launch("b200_conv2d");
  }
  GemmArgs a = base_gemm(dtype);
  a.b = filter;  // [R*S*C, K] row-major == HWIO
  a.c = output;
  a.M = rows;
  a.N = g.K;
  a.K = rsc;
  a.ldb = g.K;
  a.ldc = g.K;
  a.a_mn_major = false;
  a.b_mn_major = true;
  ConvAOperand ca{input, g.N, g.H, g.W, g.C, g.R, g.S, g.OH, g.OW, g.sh, g.sw, g.pt, g.pl};
  static const bool no_implicit = getenv("B200TF_CONV_EXPLICIT") != nullptr;
  if (!is_pointwise(g) && g.sh == 1 && g.sw == 1 && !no_implicit &&
      b200_get_matmul_precision() == 0) {
    // unit stride: halo tile in shared memory, every tap a shifted view of it (conv_halo.cu)
    ConvHaloArgs h{input, filter, output, bias, relu, g.N, g.H, g.W, g.C, g.K,
                   g.R, g.S, g.pt, g.pl, g.OH, g.OW};
    if (conv_halo_supported(dtype, h)) return conv_halo(dtype, h, s);
  }
  if (is_pointwise(g)) {
    a.a = input;
    a.lda = g.C;
    a.workspace = workspace;
    a.workspace_bytes = workspace ? workspace_bytes : 0;
  } else if (!no_implicit && b200_get_matmul_precision() == 0 && (g.K % (16 / (int)es)) == 0 &&
             aligned16(filter) && conv_a_supported(dtype, ca)) {
    // implicit GEMM: the TMA producer gathers patch rows straight from the NHWC input
    a.a = input;  // unused by the kernel, kept non-null for validation
    a.lda = g.C;
    a.conv_a = &ca;
    a.workspace = nullptr;
    a.workspace_bytes = 0;
    rc = gemm_tcgen05(a, s);
    return rc ? rc : conv2d_tail(dtype, output, rows, g.K, bias, relu, stream);
  } else {
    const size_t col_bytes = align256((size_t)rows * g.ldk * es);
    if (!workspace || workspace_bytes < col_bytes) {  // e.g. unaligned pointers ruled out TMA im2col
      set_last_error("b200_conv2d: the patch-matrix path needs %zu bytes of workspace, got %zu",
                     col_bytes, workspace_bytes);
      return B200_INVALID_ARGUMENT;
    }
    rc = dtype == B200_DT_FLOAT ? run_im2col<float>(input, workspace, g, s)
                                : run_im2col<__nv_bfloat16>(input, workspace, g, s);
    if (rc) return rc;
    a.a = workspace;
    a.lda = g.ldk;
    a.workspace = static_cast<char*>(workspace) + col_bytes;
    a.workspace_bytes = workspace_bytes - col_bytes;
    if (a.workspace_bytes == 0) a.workspace = nullptr;
  }
  rc = gemm_dispatch(a, s);
  return rc ? rc : conv2d_tail(dtype, output, rows, g.K, bias, relu, stream);
}

int b200_conv2d_backprop_filter(int dtype, const void* input, const void* out_backprop,
                                void* filter_backprop, const b200_conv2d_geometry* geom,
                                void* workspace, size_t workspace_bytes, void* stream) {
  ConvG g;
  int rc = to_geom("b200_conv2d_backprop_filter", dtype, geom, &g);
  if (rc) return rc;
  const long long rows = (long long)g.N * g.OH * g.OW;
  const long long rsc = (long long)g.R * g.S * g.C;
  if (rsc == 0 || g.K == 0) return B200_OK;
  rc = require_device("b200_conv2d_backprop_filter");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const size_t es = esize_of(dtype);
  if (rows == 0) return b200_memset_async(filter_backprop, 0, (size_t)rsc * g.K * es, stream);
  const size_t need = b200_conv2d_workspace_bytes(dtype, geom, 2);
  if (need > 0 && (!workspace || workspace_bytes < need)) {
    set_last_error("b200_conv2d_backprop_filter: workspace too small (%zu < %zu bytes)",
                   workspace_bytes, need);
    return B200_INVALID_ARGUMENT;
  }
  static const bool no_direct = getenv("B200TF_CONV_NO_DIRECT") != nullptr;
  const SmallCinFit small = small_cin_fit(dtype, g);
  static const bool no_c1 = getenv("B200TF_CONV_NO_C1") != nullptr;
  const C1Fit c1 = c1_fit(dtype, g);
  if (c1.wgrad && small.ok && !no_direct && !no_c1 && aligned16(out_backprop)) {
    const int parts = (g.N < 8 * sm_count() && g.OH >= 2 * kC1Warps) ? 2 : 1;
    const int blocks = std::min(g.N * parts, small_cin_wgrad_blocks());  // the workspace holds that many
    float* partial = static_cast<float*>(workspace);
    const float* xin = static_cast<const float*>(input);
    const float* dyin = static_cast<const float*>(out_backprop);
#define C1_WGRAD(R_, S_)                                                                        \
  do {                                                                                          \
    static const cudaError_t attr = cudaFuncSetAttribute(                                       \
        conv_c1_wgrad_kernel<R_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); \
    (void)attr;                                                                                 \
    launch_pdl(conv_c1_wgrad_kernel<R_, S_>, dim3((unsigned)blocks), dim3(kC1Threads),          \
               c1.smem_wgrad, s, xin, dyin, partial, g, parts);                                 \
  } while (0)
    if (g.R == 5)
      C1_WGRAD(5, 5);
    else
      C1_WGRAD(3, 3);
#undef C1_WGRAD
    const int elems = small.taps * g.K;
    launch_pdl(conv_small_cin_wgrad_reduce_kernel, dim3((elems + 31) / 32), dim3(32, 32), 0, s,
               static_cast<const float*>(partial), static_cast<float*>(filter_backprop), blocks,
               elems);
    note_launch(2);
    return check_launch("b200_conv2d_backprop_filter");
  }
  if (small.ok && g.K % 32 == 0 && !no_direct && !is_pointwise(g) && aligned16(out_backprop)) {
    const int blocks = small_cin_wgrad_blocks();
    float* partial = static_cast<float*>(workspace);
    launch_pdl(conv_small_cin_wgrad_kernel, dim3(blocks), dim3(kWgradThreads), 0, s,
               static_cast<const float*>(input), static_cast<const float*>(out_backprop), partial, g,
               (int)rows);
    const int elems = small.taps * g.K;
    launch_pdl(conv_small_cin_wgrad_reduce_kernel, dim3((elems + 31) / 32), dim3(256), 0, s,
               static_cast<const float*>(partial), static_cast<float*>(filter_backprop), blocks,
               elems);
    note_launch(2);
    return check_launch("b200_conv2d_backprop_filter");
  }
  if (!is_pointwise(g) && g.sh == 1 && g.sw == 1 && getenv("B200TF_CONV_EXPLICIT") == nullptr &&
      b200_get_matmul_precision() == 0) {
    ConvHaloArgs h{input, out_backprop, filter_backprop, nullptr, false, g.N, g.H, g.W, g.C, g.K,
                   g.R, g.S, g.pt, g.pl, g.OH, g.OW};
    if (conv_halo_wgrad_supported(dtype, h) &&
        workspace_bytes >= conv_halo_wgrad_workspace_bytes(dtype, h))
      return conv_halo_wgrad(dtype, h, workspace, workspace_bytes, s);
  }
  GemmArgs a = base_gemm(dtype);
  a.b = out_backprop;  // [rows, K]
  a.c = filter_backprop;
  a.M = rsc;
  a.N = g.K;
  a.K = rows;
  a.ldb = g.K;
  a.ldc = g.K;
  a.a_mn_major = true;  // A^T is stored: patches [rows, R*S*C]
  a.b_mn_major = true;
  ConvAOperand ca{input, g.N, g.H, g.W, g.C, g.R, g.S, g.OH, g.OW, g.sh, g.sw, g.pt, g.pl};
  static const bool no_implicit = getenv("B200TF_CONV_EXPLICIT") != nullptr;
  if (!is_pointwise(g) && !no_implicit && b200_get_matmul_precision() == 0 &&
      (g.K % (16 / (int)es)) == 0 && aligned16(out_backprop) && conv_a_supported(dtype, ca)) {
    // implicit GEMM: im2col boxes feed the MN-major A operand directly (no patch matrix)
    a.a = input;
    a.lda = g.C;
    a.conv_a = &ca;
    a.workspace = workspace;
    a.workspace_bytes = workspace ? workspace_bytes : 0;
    return gemm_tcgen05(a, s);
  }
  if (is_pointwise(g)) {
    a.a = input;
    a.lda = g.C;
    a.workspace = workspace;
    a.workspace_bytes = workspace ? workspace_bytes : 0;
  } else {
    const size_t col_bytes = align256((size_t)rows * g.ldk * es);
    if (!workspace || workspace_bytes < col_bytes) {
      set_last_error("b200_conv2d_backprop_filter: the patch-matrix path needs %zu bytes of "
                     "workspace, got %zu", col_bytes, workspace_bytes);
      return B200_INVALID_ARGUMENT;
    }
    rc = dtype == B200_DT_FLOAT ? run_im2col<float>(input, workspace, g, s)
                                : run_im2col<__nv_bfloat16>(input, workspace, g, s);
    if (rc) return rc;
    a.a = workspace;
    a.lda = g.ldk;
    a.workspace = static_cast<char*>(workspace) + col_bytes;
    a.workspace_bytes = workspace_bytes - col_bytes;
    if (a.workspace_bytes == 0) a.workspace = nullptr;
  }
  return gemm_dispatch(a, s);
}

int b200_conv2d_backprop_input(int dtype, const void* filter, const void* out_backprop,
                               void* in_backprop, const b200_conv2d_geometry* geom,
                               void* workspace, size_t workspace_bytes, void* stream) {
  ConvG g;
  int rc = to_geom("b200_conv2d_backprop_input", dtype, geom, &g);
  if (rc) return rc;
  const long long rows = (long long)g.N * g.OH * g.OW;
  const long long rsc = (long long)g.R * g.S * g.C;
  const long long nin = (long long)g.N * g.H * g.W * g.C;
  if (nin == 0) return B200_OK;
  rc = require_device("b200_conv2d_backprop_input");
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const size_t es = esize_of(dtype);
  if (rows == 0 || g.K == 0) return b200_memset_async(in_backprop, 0, (size_t)nin * es, stream);
  const size_t need = b200_conv2d_workspace_bytes(dtype, geom, 1);
  if (need > 0 && (!workspace || workspace_bytes < need)) {
    set_last_error("b200_conv2d_backprop_input: workspace too small (%zu < %zu bytes)",
                   workspace_bytes, need);
    return B200_INVALID_ARGUMENT;
  }
  {
    ConvAOperand cd{out_backprop, g.N, g.OH, g.OW, g.K, g.R, g.S, g.H, g.W,
                    1, 1, g.R - 1 - g.pt, g.S - 1 - g.pl};
    static const bool no_implicit = getenv("B200TF_CONV_EXPLICIT") != nullptr;
    if (!is_pointwise(g) && g.sh == 1 && g.sw == 1 && !no_implicit &&
        b200_get_matmul_precision() == 0 && (g.C % (16 / (int)es)) == 0 && cd.pt >= 0 &&
        cd.pl >= 0 && aligned16(in_backprop) && conv_a_supported(dtype, cd)) {
      const int nflt = g.R * g.S * g.C * g.K;
      if (dtype == B200_DT_FLOAT)
        launch_pdl(flip_filter_kernel<float>, dim3((nflt + 255) / 256), dim3(256), 0, s, 
            static_cast<const float*>(filter), static_cast<float*>(workspace), g.R, g.S, g.C, g.K);
      else
        launch_pdl(flip_filter_kernel<__nv_bfloat16>, dim3((nflt + 255) / 256), dim3(256), 0, s, 
            static_cast<const __nv_bfloat16*>(filter), static_cast<__nv_bfloat16*>(workspace), g.R,
            g.S, g.C, g.K);
      note_launch();
      {
        // dInput = unit-stride convolution of dY with the flipped filter [R, S, K, C]
        ConvHaloArgs h{out_backprop, workspace, in_backprop, nullptr, false, g.N, g.OH, g.OW, g.K,
                       g.C, g.R, g.S, cd.pt, cd.pl, g.H, g.W};
        if (conv_halo_supported(dtype, h)) return conv_halo(dtype, h, s);
      }
      GemmArgs d = base_gemm(dtype);
      d.a = out_backprop;
      d.lda = g.K;
      d.conv_a = &cd;
      d.b = workspace;  // [R*S*K, C] row-major
      d.ldb = g.C;
      d.b_mn_major = true;
      d.a_mn_major = false;
      d.c = in_backprop;
      d.ldc = g.C;
      d.M = (long long)g.N * g.H * g.W;
      d.N = g.C;
      d.K = (long long)g.R * g.S * g.K;
      return gemm_tcgen05(d, s);
    }
  }
  GemmArgs a = base_gemm(dtype);
  a.a = out_backprop;  // [rows, K]
  a.lda = g.K;
  a.b = filter;        // [R*S*C, K]: logical B[K, R*S*C] stored transposed => K-major
  a.ldb = g.K;
  a.M = rows;
  a.N = rsc;
  a.K = g.K;
  a.a_mn_major = false;
  a.b_mn_major = false;
  if (is_pointwise(g)) {
    a.c = in_backprop;
    a.ldc = g.C;
    a.workspace = workspace;
    a.workspace_bytes = workspace ? workspace_bytes : 0;
    return gemm_dispatch(a, s);
  }
  const size_t col_bytes = align256((size_t)rows * g.ldk * es);
  if (!workspace || workspace_bytes < col_bytes) {
    set_last_error("b200_conv2d_backprop_input: the patch-matrix path needs %zu bytes of "
                   "workspace, got %zu", col_bytes, workspace_bytes);
    return B200_INVALID_ARGUMENT;
  }
  a.c = workspace;
  a.ldc = g.ldk;
  a.workspace = static_cast<char*>(workspace) + col_bytes;
  a.workspace_bytes = workspace_bytes - col_bytes;
  if (a.workspace_bytes == 0) a.workspace = nullptr;
  rc = gemm_dispatch(a, s);
  if (rc) return rc;
  if (dtype == B200_DT_FLOAT)
    launch_pdl(col2im_gather_kernel<float>, dim3((unsigned)((nin + 255) / 256)), dim3(256), 0, s, 
        static_cast<const float*>(workspace), static_cast<float*>(in_backprop), g, nin);
  else
    launch_pdl(col2im_gather_kernel<__nv_bfloat16>, dim3((unsigned)((nin + 255) / 256)), dim3(256), 0, s, 
        static_cast<const __nv_bfloat16*>(workspace), static_cast<__nv_bfloat16*>(in_backprop), g,
        nin);
  note_launch();
  return check_launch("col2im_gather");
}

}  // extern "C"
