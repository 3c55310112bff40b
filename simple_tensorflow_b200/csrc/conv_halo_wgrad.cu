// Halo-tile filter gradient for unit-stride convolutions on sm_100a.
//
//   dW[r, s, c, k] = sum_{n, oh, ow} x[n, oh + r - pt, ow + s - pl, c] * dy[n, oh, ow, k]
//
// Replaces the cuDNN backward-filter call of Conv2DSlowBackpropFilterOp<GPUDevice,T>
// (tensorflow/core/kernels/conv_grad_filter_ops.cc:361-738) for unit-stride NHWC convolutions.
//
// Like conv_halo.cu, a CTA loads the input halo of a band of output rows ONCE and addresses every
// filter tap as the same shared-memory tile shifted by whole pixel rows (tcgen05 applies the
// swizzle to absolute shared-memory address bits: tools/desc_probe.cu).  The contraction runs
// over output pixels in padded-row order p' = r' * WP + c' (WP = OW + S - 1):
//
//   D_g[(j, c), k] += sum_{p'} x_halo[p' + shift_g + j * stride_g][c] * dy_pad[p'][k]
//
//   * A = the x halo tile read as an MN-major operand whose M dimension is "G taps x one
//     128-byte channel block": chunk j of the operand is the SAME tile shifted by
//     shift_g + j * stride_g pixel rows, expressed through the descriptor's leading byte offset
//     (overlapping chunks; verified by the probe's third test).  G = 4 taps (fp32) / 2 (bf16)
//     fill the M = 128 rows of one MMA: horizontal runs of taps use stride 1 row, vertical runs
//     stride WP rows (a 5x5 filter = 5 horizontal groups + 1 vertical + 1 single).
//   * B = the dy band, MN-major, loaded by TMA into the same padded enumeration: the box is WP
//     pixels wide, so the S - 1 junk columns per row are out-of-bounds ZEROS and contribute
//     nothing; rows beyond the band are zeroed once.
//   * D = every group's 128 x BN fp32 accumulator stays in TMEM for the whole kernel
//     (groups x channel blocks x BN <= 512 columns); a CTA streams its share of the images and
//     writes ONE partial filter gradient, which an ordered second pass sums (deterministic).
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "b200_internal.h"
#include "b200_ptx.cuh"

namespace b200 {

namespace {

constexpr int kWgIssueWarps = 4;  // MMA-issuing warps: warp 1 and warps 6..8
constexpr int kWgThreads = 192 + 32 * (kWgIssueWarps - 1);
constexpr int kRowB = 128;
constexpr int kMaxGroups = 16;

struct TapGroup {
  int shift;    // pixel-row shift of chunk 0: r * WP + s
  int stride;   // pixel-row distance between consecutive chunks (1 = horizontal run, WP = vertical)
  int ntaps;    // valid chunks (<= G); the rest of the M rows are junk and never stored
  int tap0;     // filter tap index (r * S + s) of chunk 0
  int tap_step; // tap index distance between chunks (1 horizontal, S vertical)
};

struct WgradShape {
  int N, H, W, C, K;
  int R, S, pt, pl, OH, OW;
  int bh, tiles_h, WP, HP;
  int ksteps;            // MMA K steps per item: ceil(bh * WP / kUmmaK)
  int cblocks;           // channel blocks handled by one CTA (all of C)
  int x_rows, dy_rows;   // rows allocated per channel block / per dy column chunk (multiples of 8)
  int ngroups;
  TapGroup groups[kMaxGroups];
  int nblocks;           // K / BN partitions (each CTA handles one)
  int slices;            // CTAs per partition; each streams items slice, slice + slices, ...
  long long items;       // N * tiles_h
  float* partial;        // [slices][R*S*C][K] fp32
};

template <typename T>
struct WgTraits;
template <>
struct WgTraits<float> {
  static constexpr int kChunk = 32;
  static constexpr int kUmmaK = 8;
  static constexpr uint32_t kFormat = 2;
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  static constexpr bool kMn32 = true;
  static constexpr CUtensorMapSwizzle kSwz = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
};
template <>
struct WgTraits<__nv_bfloat16> {
  static constexpr int kChunk = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kFormat = 1;
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static constexpr bool kMn32 = false;
  static constexpr CUtensorMapSwizzle kSwz = CU_TENSOR_MAP_SWIZZLE_128B;
};

template <typename TIn, int BN>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_halo_wgrad_kernel(const __grid_constant__ CUtensorMap tmapX,
                       const __grid_constant__ CUtensorMap tmapDY, const WgradShape s) {
  pdl_launch_dependents();
  using Tr = WgTraits<TIn>;
  constexpr int kChunk = Tr::kChunk;
  constexpr int kNChunks = BN / kChunk;
  constexpr int kG = 128 / kChunk;  // taps per group (chunks of the M = 128 operand)
  constexpr uint32_t kIdesc = make_idesc(Tr::kFormat, true, true, 128, BN);
  constexpr uint32_t kSbo = Tr::kMn32 ? 512 : 1024;
  constexpr uint32_t kLayout = Tr::kMn32 ? 1 : 2;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int x_bytes = s.cblocks * s.x_rows * kRowB;
  const int dy_bytes = kNChunks * s.dy_rows * kRowB;
  uint8_t* smX = smem;                      // [2][cblocks][x_rows][128]
  uint8_t* smDY = smem + 2 * x_bytes;       // [2][kNChunks][dy_rows][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smDY + 2 * dy_bytes);
  uint64_t* full = bars;        // [2]
  uint64_t* empty = bars + 2;   // [2]
  uint64_t* done = bars + 4;    // [1] all MMAs of this CTA retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int nb = blockIdx.x / s.slices;           // K partition of this CTA
  const int slice = blockIdx.x - nb * s.slices;   // its share of the items

  // Rows the TMA boxes never write are read by the MMAs (the K tail of the dy band, the halo
  // rows behind the last shifted window): zero them once so that they hold finite values / add 0.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (2 * x_bytes + 2 * dy_bytes) / 16;
    for (int i = threadIdx.x; i < n16; i += kWgThreads) z[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async_smem();
  }
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapX);
    tma_prefetch_desc(&tmapDY);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&full[i], 1);
        mbar_init(&empty[i], kWgIssueWarps);
      }
      mbar_init(&done[0], kWgIssueWarps);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // (the shuffle makes the value provably warp-uniform: uniform-register descriptor math)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();

  const uint32_t item_bytes = (uint32_t)(s.cblocks * s.HP * s.WP * kRowB + kNChunks * s.bh * s.WP * kRowB);
  if (warp == 0) {
    if (lane == 0) {
      uint32_t buf = 0, phase = 0;
      for (long long it = slice; it < s.items; it += s.slices) {
        const int n = (int)(it / s.tiles_h);
        const int oh0 = (int)(it - (long long)n * s.tiles_h) * s.bh;
        mbar_wait(&empty[buf], phase ^ 1);
        mbar_expect_tx(&full[buf], item_bytes);
        for (int cb = 0; cb < s.cblocks; ++cb)
          tma_load_4d(smX + buf * x_bytes + cb * s.x_rows * kRowB, &tmapX, &full[buf], cb * kChunk,
                      -s.pl, oh0 - s.pt, n);
        for (int c = 0; c < kNChunks; ++c)
          tma_load_4d(smDY + buf * dy_bytes + c * s.dy_rows * kRowB, &tmapDY, &full[buf],
                      (nb * kNChunks + c) * kChunk, 0, oh0, n);
        if (++buf == 2) {
          buf = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 || warp >= 6) {
    // Four issuing warps (a short MMA costs its issuer ~130 cycles of R2UR moves, conv_halo.cu):
    // issuer iw owns the tap groups g = iw, iw + 4, ... and with them their accumulators.
    const int iw = warp == 1 ? 0 : warp - 5;
    {  // the whole warp runs the loop; one elected lane issues each tcgen05 instruction
      // per-group descriptor of K step 0 in buffer 0 / channel block 0 (start, LBO = tap stride)
      uint64_t gdesc[kMaxGroups];
#pragma unroll
      for (int g = 0; g < kMaxGroups; ++g)
        gdesc[g] = g < s.ngroups
                       ? make_smem_desc_sw128(smem_u32(smX) + (uint32_t)s.groups[g].shift * kRowB,
                                              (uint32_t)s.groups[g].stride * kRowB, kSbo, kLayout)
                       : 0ull;
      const uint64_t bdesc0 =
          make_smem_desc_sw128(smem_u32(smDY), (uint32_t)s.dy_rows * kRowB, kSbo, kLayout);
      constexpr uint32_t kStepK = (Tr::kUmmaK * kRowB) >> 4;  // kUmmaK pixel rows per K step
      uint32_t buf = 0, phase = 0;
      bool first_item = true;
      for (long long it = slice; it < s.items; it += s.slices) {
        mbar_wait(&full[buf], phase);
        __syncwarp();
        tc_fence_after();
        const uint64_t xoff = (uint64_t)((buf * x_bytes) >> 4);
        const uint64_t bdesc_b = bdesc0 + (uint64_t)((buf * dy_bytes) >> 4);
        const uint32_t b_hi = (uint32_t)(bdesc_b >> 32);
        for (int ks = 0; ks < s.ksteps; ++ks) {
          const uint32_t b_lo = (uint32_t)bdesc_b + (uint32_t)ks * kStepK;
          const uint32_t accum = (first_item && ks == 0) ? 0u : 1u;
          for (int cb = 0; cb < s.cblocks; ++cb) {
            const uint32_t aoff = (uint32_t)xoff + (uint32_t)((cb * s.x_rows * kRowB) >> 4) +
                                  (uint32_t)ks * kStepK;
#pragma unroll
            for (int g = 0; g < kMaxGroups; ++g) {
              if (g < s.ngroups && (g % kWgIssueWarps) == iw) {
                const uint32_t d_tmem = tmem_base + (uint32_t)((cb * s.ngroups + g) * BN);
                if (sizeof(TIn) == 4)
                  umma_tf32_elect_lohi(d_tmem, (uint32_t)gdesc[g] + aoff, (uint32_t)(gdesc[g] >> 32),
                                       b_lo, b_hi, kIdesc, accum);
                else
                  umma_f16_elect_lohi(d_tmem, (uint32_t)gdesc[g] + aoff, (uint32_t)(gdesc[g] >> 32),
                                      b_lo, b_hi, kIdesc, accum);
              }
            }
          }
        }
        umma_commit_elect(&empty[buf]);  // buffers free once these MMAs retire
        first_item = false;
        if (++buf == 2) {
          buf = 0;
          phase ^= 1;
        }
      }
      umma_commit_elect(&done[0]);
    }
  } else {
    // ===================== epilogue: TMEM -> this CTA's partial filter gradient =====================
    const int quad = warp & 3;
    mbar_wait(&done[0], 0);
    tc_fence_after();
    const bool any = slice < s.items;  // a CTA without items holds stale accumulators: write zeros
    const int lg = quad * 32 + lane;   // accumulator row = (chunk j, channel c)
    const int j = lg / kChunk, c = lg - j * kChunk;
    float* part = s.partial + (long long)slice * ((long long)s.R * s.S * s.C * s.K);
    for (int cb = 0; cb < s.cblocks; ++cb) {
      for (int g = 0; g < s.ngroups; ++g) {
        const TapGroup& tg = s.groups[g];
        const bool valid = j < tg.ntaps;
        const int tap = tg.tap0 + j * tg.tap_step;
        float* row = part + ((long long)tap * s.C + cb * kChunk + c) * s.K + nb * BN;
#pragma unroll 1
        for (int col = 0; col < BN; col += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (cb * s.ngroups + g) * BN + col, v);
          tmem_ld_wait();
          if (valid) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float4 f = any ? make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]))
                             : make_float4(0.f, 0.f, 0.f, 0.f);
              *reinterpret_cast<float4*>(row + col + 4 * q) = f;
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// out[i] = sum over slices of partial[s][i], in a fixed order (deterministic): block (32, 8), thread
// row y adds slices y, y + 8, ... (independent 16-byte loads, all in flight), then the 8 row sums are
// added in order through shared memory.
template <typename TOut>
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, TOut* __restrict__ out, int slices,
                    long long n4) {
  pdl_prologue();
  __shared__ float4 sm[8][32];
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n4) {
    const float4* p = reinterpret_cast<const float4*>(partial) + i;
#pragma unroll 4
    for (int sl = y; sl < slices; sl += 8) {
      const float4 v = __ldg(p + (long long)sl * n4);
      a.x += v.x;
      a.y += v.y;
      a.z += v.z;
      a.w += v.w;
    }
  }
  sm[y][x] = a;
  __syncthreads();
  if (y == 0 && i < n4) {
    float4 t = sm[0][x];
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      t.x += sm[k][x].x;
      t.y += sm[k][x].y;
      t.z += sm[k][x].z;
      t.w += sm[k][x].w;
    }
    if (sizeof(TOut) == 4) {
      reinterpret_cast<float4*>(out)[i] = t;
    } else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y), hi = __floats2bfloat162_rn(t.z, t.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      reinterpret_cast<uint2*>(out)[i] = pk;
    }
  }
}

struct WgradPlan {
  WgradShape s;
  size_t smem;
  int bn;
  bool ok;
};

constexpr size_t kWgSmemLimit = 227 * 1024;

// Tap groups: full horizontal runs of G taps per filter row, the leftover columns as vertical
// runs, what remains as partial groups.
static int plan_groups(int R, int S, int WP, int G, TapGroup* out) {
  int n = 0;
  const int full_cols = S / G * G;
  for (int r = 0; r < R; ++r)
    for (int s0 = 0; s0 < full_cols; s0 += G) {
      if (n == kMaxGroups) return -1;
      out[n++] = TapGroup{r * WP + s0, 1, G, r * S + s0, 1};
    }
  const int rem = S - full_cols;
  if (rem > 0) {
    const int full_rows = R / G * G;
    for (int sx = full_cols; sx < S; ++sx)
      for (int r0 = 0; r0 < full_rows; r0 += G) {
        if (n == kMaxGroups) return -1;
        out[n++] = TapGroup{r0 * WP + sx, WP, G, r0 * S + sx, S};
      }
    // leftover rows: horizontal partial groups of `rem` taps
    for (int r = full_rows; r < R; ++r) {
      if (n == kMaxGroups) return -1;
      out[n++] = TapGroup{r * WP + full_cols, 1, rem, r * S + full_cols, 1};
    }
  }
  return n;
}

static WgradPlan plan_wgrad(int dtype, const ConvHaloArgs& a) {
  WgradPlan p{};
  p.ok = false;
  const int es = dtype == B200_DT_FLOAT ? 4 : 2;
  const int chunk = kRowB / es, G = 128 / chunk, ummak = 32 / es;
  if (a.C % chunk || a.K % chunk || a.C <= 0 || a.K <= 0) return p;
  const int WP = a.OW + a.S - 1;
  if (WP > 256) return p;
  WgradShape& s = p.s;
  s.ngroups = plan_groups(a.R, a.S, WP, G, s.groups);
  if (s.ngroups <= 0) return p;
  s.cblocks = a.C / chunk;
  // BN: the widest K partition whose accumulators fit the 512 TMEM columns
  int bn = 0;
  for (int cand : {256, 128, 64, 32}) {
    if (cand < chunk || a.K % cand) continue;
    if ((long long)s.ngroups * s.cblocks * cand <= 512) {
      bn = cand;
      break;
    }
  }
  if (bn == 0) return p;
  p.bn = bn;
  s.nblocks = a.K / bn;
  int max_reach = 0;  // furthest pixel row any chunk of any group reaches beyond the K window
  for (int g = 0; g < s.ngroups; ++g)
    max_reach = std::max(max_reach, s.groups[g].shift + (G - 1) * s.groups[g].stride);
  const int units = std::max(1, sm_count() / s.nblocks);
  double best = -1;
  for (int bh = 1; bh <= std::min(a.OH, 256 - a.R + 1); ++bh) {
    const int HP = bh + a.R - 1;
    const int ksteps = (bh * WP + ummak - 1) / ummak;
    const int x_rows = (std::max(HP * WP, ksteps * ummak + max_reach) + 7) / 8 * 8;
    const int dy_rows = (std::max(bh * WP, ksteps * ummak) + 7) / 8 * 8;
    const size_t smem = 2 * ((size_t)s.cblocks * x_rows + (size_t)(bn / chunk) * dy_rows) * kRowB +
                        64 + 1024;
    if (smem > kWgSmemLimit) break;
    const int th = (a.OH + bh - 1) / bh;
    const long long items = (long long)a.N * th;
    const long long rounds = (items + units - 1) / units;
    // cost ~ rounds x MMAs per item; fewer, fuller rounds win
    const double cost = (double)rounds * ksteps;
    const double score = 1.0 / cost;
    if (score > best) {
      best = score;
      p.ok = true;
      s.bh = bh; s.tiles_h = th; s.HP = HP; s.WP = WP; s.ksteps = ksteps; s.x_rows = x_rows;
      s.dy_rows = dy_rows; s.items = items;
      p.smem = smem;
    }
  }
  if (!p.ok) return p;
  s.N = a.N; s.H = a.H; s.W = a.W; s.C = a.C; s.K = a.K; s.R = a.R; s.S = a.S; s.pt = a.pt;
  s.pl = a.pl; s.OH = a.OH; s.OW = a.OW;
  s.slices = (int)std::min<long long>(units, s.items);
  if (s.slices < 1) s.slices = 1;
  return p;
}

template <typename TIn, int BN>
static int launch_wgrad(const WgradPlan& p, const ConvHaloArgs& a, void* workspace,
                        cudaStream_t stream) {
  using Tr = WgTraits<TIn>;
  WgradShape s = p.s;
  constexpr int es = (int)sizeof(TIn);
  s.partial = static_cast<float*>(workspace);
  CUtensorMap mx, mdy;
  cuuint32_t estr[4] = {1, 1, 1, 1};
  {
    cuuint64_t gdim[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
    cuuint64_t gstr[3] = {(cuuint64_t)s.C * es, (cuuint64_t)s.W * s.C * es,
                          (cuuint64_t)s.H * s.W * s.C * es};
    cuuint32_t box[4] = {(cuuint32_t)Tr::kChunk, (cuuint32_t)s.WP, (cuuint32_t)s.HP, 1};
    if (driver().cuTensorMapEncodeTiled(&mx, Tr::kTmaType, 4, const_cast<void*>(a.input), gdim,
                                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, Tr::kSwz,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_last_error("conv_halo_wgrad: input tensor map failed");
      return B200_INTERNAL;
    }
  }
  {
    cuuint64_t gdim[4] = {(cuuint64_t)s.K, (cuuint64_t)s.OW, (cuuint64_t)s.OH, (cuuint64_t)s.N};
    cuuint64_t gstr[3] = {(cuuint64_t)s.K * es, (cuuint64_t)s.OW * s.K * es,
                          (cuuint64_t)s.OH * s.OW * s.K * es};
    cuuint32_t box[4] = {(cuuint32_t)Tr::kChunk, (cuuint32_t)s.WP, (cuuint32_t)s.bh, 1};
    if (driver().cuTensorMapEncodeTiled(&mdy, Tr::kTmaType, 4, const_cast<void*>(a.filter), gdim,
                                        gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, Tr::kSwz,
                                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
      set_last_error("conv_halo_wgrad: out_backprop tensor map failed");
      return B200_INTERNAL;
    }
  }
  auto kern = conv_halo_wgrad_kernel<TIn, BN>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)kWgSmemLimit) != cudaSuccess) {
      set_last_error("conv_halo_wgrad: cudaFuncSetAttribute failed");
      cudaGetLastError();
      return B200_INTERNAL;
    }
    attr_set = true;
  }
  const bool prof = profile_enabled();
  if (prof) profile_gemm_launch_begin(stream);
  cudaError_t e = launch_pdl(kern, dim3((unsigned)(s.slices * s.nblocks)), dim3(kWgThreads), p.smem,
                             stream, mx, mdy, s);
  if (e != cudaSuccess) {
    set_last_error("conv_halo_wgrad launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  if (prof)
    profile_gemm_launch_end(stream, 2.0 * (double)s.N * s.OH * s.OW * (double)s.K * s.R * s.S * s.C);
  const long long n4 = (long long)s.R * s.S * s.C * s.K / 4;
  const long long blocks = (n4 + 31) / 32;
  e = launch_pdl(wgrad_reduce_kernel<TIn>, dim3((unsigned)blocks), dim3(256), 0, stream,
                 static_cast<const float*>(workspace), static_cast<TIn*>(a.output), s.slices, n4);
  if (e != cudaSuccess) {
    set_last_error("conv_halo_wgrad reduce launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  note_launch(2);
  return check_launch("conv_halo_wgrad");
}

}  // namespace

// ConvHaloArgs here: input = x [N,H,W,C], filter = out_backprop dy [N,OH,OW,K], output = dW.
size_t conv_halo_wgrad_workspace_bytes(int dtype, const ConvHaloArgs& a) {
  const WgradPlan p = plan_wgrad(dtype, a);
  if (!p.ok) return 0;
  return (size_t)p.s.slices * a.R * a.S * a.C * a.K * sizeof(float);
}

bool conv_halo_wgrad_supported(int dtype, const ConvHaloArgs& a) {
  if (!driver().cuTensorMapEncodeTiled) return false;
  if (dtype != B200_DT_FLOAT && dtype != B200_DT_BFLOAT16) return false;
  static const bool off = getenv("B200TF_CONV_NO_HALO") != nullptr ||
                          getenv("B200TF_CONV_NO_HALO_WGRAD") != nullptr;
  if (off) return false;
  if ((reinterpret_cast<uintptr_t>(a.input) & 15) || (reinterpret_cast<uintptr_t>(a.filter) & 15) ||
      (reinterpret_cast<uintptr_t>(a.output) & 15))
    return false;
  if (a.N <= 0 || a.OH <= 0 || a.OW <= 0 || a.pt < 0 || a.pl < 0) return false;
  if (((long long)a.R * a.S * a.C * a.K) % 4) return false;
  return plan_wgrad(dtype, a).ok;
}

int conv_halo_wgrad(int dtype, const ConvHaloArgs& a, void* workspace, size_t workspace_bytes,
                    cudaStream_t stream) {
  const WgradPlan p = plan_wgrad(dtype, a);
  if (!p.ok) {
    set_last_error("conv_halo_wgrad: unsupported geometry");
    return B200_UNIMPLEMENTED;
  }
  const size_t need = (size_t)p.s.slices * a.R * a.S * a.C * a.K * sizeof(float);
  if (!workspace || workspace_bytes < need) {
    set_last_error("conv_halo_wgrad: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    return B200_INVALID_ARGUMENT;
  }
#define WG_CASE(T, BN_) \
  if (p.bn == BN_) return launch_wgrad<T, BN_>(p, a, workspace, stream)
  if (dtype == B200_DT_FLOAT) {
    WG_CASE(float, 32);
    WG_CASE(float, 64);
    WG_CASE(float, 128);
    WG_CASE(float, 256);
  } else {
    WG_CASE(__nv_bfloat16, 64);
    WG_CASE(__nv_bfloat16, 128);
    WG_CASE(__nv_bfloat16, 256);
  }
#undef WG_CASE
  set_last_error("conv_halo_wgrad: no kernel for BN = %d", p.bn);
  return B200_UNIMPLEMENTED;
}

}  // namespace b200
