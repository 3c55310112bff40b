// MatMul / BatchMatMul core for sm_100a: persistent, warp-specialised tcgen05 GEMM.
//
//   C[b][M,N] = op(A[b]) * op(B[b])       fp32 I/O on kind::tf32, bf16 I/O on kind::f16,
//                                          fp32 accumulation in TMEM in both cases.
//
// Replaces the reference's GPU path  LaunchMatMul<GPUDevice,T,true>::launch -> Stream::ThenBlasGemm
// (tensorflow/core/kernels/matmul_op.cc:162-203) and LaunchBatchMatMul<GPUDevice>
// (tensorflow/core/kernels/batch_matmul_op_impl.h:297-363).  All four transpose_a/transpose_b
// combinations are served without any data movement: a row-major [rows,K] operand is "K-major",
// a row-major [K,rows] operand is "MN-major"; both are legal tcgen05 shared-memory layouts for
// tf32 and bf16 (instruction-descriptor bits 15/16).
//
// CTA layout (192 threads, one CTA per SM, persistent over output tiles):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier tx)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld TMEM -> registers -> 128-bit global stores)
// Pipelines: smem full/empty ring (kStages), TMEM accumulator full/empty (2 stages) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include "b200_ptx.cuh"
#include "b200_internal.h"

#include <cuda_bf16.h>

namespace b200 {

constexpr int kBM = 128;          // tile rows  (UMMA M, cta_group::1)
constexpr int kSwizzleBytes = 128;
constexpr int kGemmThreads = 192;

template <typename T>
struct GemmTraits;
template <>
struct GemmTraits<float> {
  static constexpr int kBK = 32;       // 128 B of K per smem row
  static constexpr int kUmmaK = 8;     // tf32: 32 B per instruction
  static constexpr uint32_t kFormat = 2;  // TF32
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
};
template <>
struct GemmTraits<__nv_bfloat16> {
  static constexpr int kBK = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kFormat = 1;  // BF16
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
};

template <int BN>
constexpr int gemm_stages() {
  return BN == 256 ? 4 : (BN == 128 ? 6 : 8);
}
template <int BN>
constexpr size_t gemm_smem_bytes() {
  // A stage + B stage, + 1 KiB alignment slack + barriers.
  return static_cast<size_t>(gemm_stages<BN>()) * (kBM * kSwizzleBytes + BN * kSwizzleBytes) +
         1024 + 256;
}

static int plan_splits(long long tiles, int num_kb);

struct GemmShape {
  int M, N, K, batch;
  int ldc;            // elements
  long long strideC;  // elements between batches
  int splits;         // split-K factor (1 = none)
  int kb_per_split;   // K blocks per split
  float* partial;     // [splits][batch][M][N] fp32 partial sums when splits > 1
};

__device__ __forceinline__ void store_row32(float* dst, const uint32_t (&v)[32], int ncols,
                                            bool vec_ok) {
  if (vec_ok && ncols == 32) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 f = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                             __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      reinterpret_cast<float4*>(dst)[j] = f;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = __uint_as_float(v[j]);
  }
}
__device__ __forceinline__ void store_row32(__nv_bfloat16* dst, const uint32_t (&v)[32],
                                            int ncols, bool vec_ok) {
  if (vec_ok && ncols == 32) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 p;
      __nv_bfloat162 h;
      h = __floats2bfloat162_rn(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
      p.x = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
      p.y = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
      p.z = *reinterpret_cast<uint32_t*>(&h);
      h = __floats2bfloat162_rn(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
      p.w = *reinterpret_cast<uint32_t*>(&h);
      reinterpret_cast<uint4*>(dst)[j] = p;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < ncols) dst[j] = __float2bfloat16_rn(__uint_as_float(v[j]));
  }
}

// TIn: operand element type (float -> tf32 MMA, bf16 -> f16-kind MMA); TOut: stored type.
template <typename TIn, typename TOut, bool kAMN, bool kBMN, int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmapA,
                    const __grid_constant__ CUtensorMap tmapB, TOut* __restrict__ C,
                    GemmShape s) {
  using Tr = GemmTraits<TIn>;
  constexpr int BK = Tr::kBK;
  constexpr int kStages = gemm_stages<BN>();
  constexpr int kABytes = kBM * kSwizzleBytes;  // 16 KiB
  constexpr int kBBytes = BN * kSwizzleBytes;
  constexpr int kChunk = kSwizzleBytes / sizeof(TIn);  // MN elements per 128-B swizzle row
  constexpr int kTmemCols = 2 * BN;                    // double-buffered fp32 accumulator
  static_assert(kTmemCols <= 512 && (kTmemCols & (kTmemCols - 1)) == 0, "TMEM cols");
  constexpr uint32_t kIdesc = make_idesc(Tr::kFormat, kAMN, kBMN, kBM, BN);

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* smA = smem;
  uint8_t* smB = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (s.M + kBM - 1) / kBM;
  const int tiles_n = (s.N + BN - 1) / BN;
  const int tiles_per_batch = tiles_m * tiles_n;
  const int num_tiles = tiles_per_batch * s.batch;
  const int num_kb = (s.K + BK - 1) / BK;
  // Work item = (split, tile): consecutive CTAs take different tiles of the same K split.
  const int num_work = num_tiles * s.splits;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapA);
    tma_prefetch_desc(&tmapB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStages; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tfull_bar[i], 1);
        mbar_init(&tempty_bar[i], 4);  // one arrive per epilogue warp
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc<kTmemCols>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int split = work / num_tiles;
        const int tile = work - split * num_tiles;
        const int b = tile / tiles_per_batch;
        const int t = tile - b * tiles_per_batch;
        const int m0 = (t % tiles_m) * kBM;
        const int n0 = (t / tiles_m) * BN;
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, num_kb);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_expect_tx(&full_bar[stage], kABytes + kBBytes);
          const int k0 = kb * BK;
          uint8_t* a_dst = smA + stage * kABytes;
          uint8_t* b_dst = smB + stage * kBBytes;
          if (!kAMN) {
            tma_load_3d(a_dst, &tmapA, &full_bar[stage], k0, m0, b);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / kChunk; ++c)
              tma_load_3d(a_dst + c * (BK * kSwizzleBytes), &tmapA, &full_bar[stage],
                          m0 + c * kChunk, k0, b);
          }
          if (!kBMN) {
            tma_load_3d(b_dst, &tmapB, &full_bar[stage], k0, n0, b);
          } else {
#pragma unroll
            for (int c = 0; c < BN / kChunk; ++c)
              tma_load_3d(b_dst + c * (BK * kSwizzleBytes), &tmapB, &full_bar[stage],
                          n0 + c * kChunk, k0, b);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
        const int split = work / num_tiles;
        const int kb0 = split * s.kb_per_split;
        const int kb1 = min(kb0 + s.kb_per_split, num_kb);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smA + stage * kABytes);
          const uint32_t b_addr = smem_u32(smB + stage * kBBytes);
#pragma unroll
          for (int k = 0; k < BK / Tr::kUmmaK; ++k) {
            // K-major: advance 32 B inside the 128-B swizzle row; MN-major: advance kUmmaK rows.
            const uint32_t a_off = kAMN ? k * Tr::kUmmaK * kSwizzleBytes
                                        : k * Tr::kUmmaK * (int)sizeof(TIn);
            const uint32_t b_off = kBMN ? k * Tr::kUmmaK * kSwizzleBytes
                                        : k * Tr::kUmmaK * (int)sizeof(TIn);
            // MN-major tf32 must use the 32-byte-atom 128B swizzle: 4-row (512 B) K groups.
            constexpr bool kMn32 = sizeof(TIn) == 4;
            const uint64_t adesc = kAMN ? make_smem_desc_sw128(a_addr + a_off, BK * kSwizzleBytes,
                                                               kMn32 ? 512 : 1024, kMn32 ? 1 : 2)
                                        : make_smem_desc_sw128(a_addr + a_off, 16, 1024);
            const uint64_t bdesc = kBMN ? make_smem_desc_sw128(b_addr + b_off, BK * kSwizzleBytes,
                                                               kMn32 ? 512 : 1024, kMn32 ? 1 : 2)
                                        : make_smem_desc_sw128(b_addr + b_off, 16, 1024);
            if (sizeof(TIn) == 4)
              umma_tf32(d_tmem, adesc, bdesc, kIdesc, ((kb - kb0) | k) != 0);
            else
              umma_f16(d_tmem, adesc, bdesc, kIdesc, ((kb - kb0) | k) != 0);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[acc]);  // accumulator ready for the epilogue
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    uint32_t acc = 0, acc_phase = 0;
    const bool vec_ok = (s.ldc % (16 / (int)sizeof(TOut)) == 0) &&
                        (s.strideC % (16 / (int)sizeof(TOut)) == 0) &&
                        ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
    for (int work = blockIdx.x; work < num_work; work += gridDim.x) {
      const int split = work / num_tiles;
      const int tile = work - split * num_tiles;
      const int b = tile / tiles_per_batch;
      const int t = tile - b * tiles_per_batch;
      const int m0 = (t % tiles_m) * kBM;
      const int n0 = (t / tiles_m) * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const int row = m0 + quad * 32 + lane;
      TOut* crow = C + (long long)b * s.strideC + (long long)row * s.ldc;
      // split-K: fp32 partial tile, dense [split][batch][M][N]
      float* prow = s.partial + (((long long)split * s.batch + b) * s.M + row) * (long long)s.N;
      const bool pvec = (s.N & 3) == 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32, v);
        tmem_ld_wait();
        const int col = n0 + c * 32;
        int ncols = s.N - col;
        ncols = ncols > 32 ? 32 : ncols;
        if (row < s.M && ncols > 0) {
          if (s.splits > 1)
            store_row32(prow + col, v, ncols, pvec);
          else
            store_row32(crow + col, v, ncols, vec_ok);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// Split-K reduction: out = sum over splits (ascending, deterministic) of the fp32 partials.
template <typename TOut>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ partial, TOut* __restrict__ C, int splits,
                     long long batch, int M, int N, int ldc, long long strideC) {
  const long long per_split = batch * (long long)M * N;
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;  // index of a 4-column group
  const int n4 = (N + 3) / 4;
  if (i >= batch * (long long)M * n4) return;
  const int c4 = (int)(i % n4);
  const long long r = i / n4;  // b * M + row
  const long long b = r / M;
  const int row = (int)(r - b * M);
  const int col = c4 * 4;
  const int nc = min(4, N - col);
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float* p = partial + r * N + col;
  for (int s = 0; s < splits; ++s) {
    if (nc == 4 && (N & 3) == 0) {
      const float4 v = *reinterpret_cast<const float4*>(p + s * per_split);
      acc[0] += v.x;
      acc[1] += v.y;
      acc[2] += v.z;
      acc[3] += v.w;
    } else {
      for (int j = 0; j < nc; ++j) acc[j] += p[s * per_split + j];
    }
  }
  TOut* dst = C + b * strideC + (long long)row * ldc + col;
  for (int j = 0; j < nc; ++j) {
    if (sizeof(TOut) == 4)
      reinterpret_cast<float*>(dst)[j] = acc[j];
    else
      reinterpret_cast<__nv_bfloat16*>(dst)[j] = __float2bfloat16_rn(acc[j]);
  }
}

// ------------------------------------------------------------------ host side

// Operand description: row-major matrix [rows, cols] per batch, leading dimension ld (elements).
//   K-major  operand: rows = MN extent, cols = K.
//   MN-major operand: rows = K,         cols = MN extent.
static int encode_operand_map(CUtensorMap* map, CUtensorMapDataType dt, size_t esize,
                              const void* ptr, long long rows, long long cols, long long ld,
                              long long batch, long long batch_stride, int box_inner,
                              int box_rows, bool mn_major) {
  cuuint64_t gdim[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t gstride[2] = {(cuuint64_t)(ld * esize),
                           (cuuint64_t)((batch > 1 ? batch_stride : rows * ld) * esize)};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = driver().cuTensorMapEncodeTiled(
      map, dt, 3, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
      // MN-major fp32 (tf32) tiles need 32-byte swizzle atoms (UMMA SWIZZLE_128B_BASE32B)
      (mn_major && esize == 4) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
      CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)",
                   (int)r, rows, cols, ld);
    return B200_INTERNAL;
  }
  return B200_OK;
}

template <typename TIn, typename TOut, bool kAMN, bool kBMN, int BN>
static int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  using Tr = GemmTraits<TIn>;
  constexpr int kChunk = kSwizzleBytes / sizeof(TIn);
  CUtensorMap ma, mb;
  int rc;
  // A: logical [M,K]; stored [M,K] (K-major) or [K,M] (MN-major).
  if (!kAMN)
    rc = encode_operand_map(&ma, Tr::kTmaType, sizeof(TIn), g.a, g.M, g.K, g.lda, g.batch,
                            g.strideA, Tr::kBK, kBM, false);
  else
    rc = encode_operand_map(&ma, Tr::kTmaType, sizeof(TIn), g.a, g.K, g.M, g.lda, g.batch,
                            g.strideA, kChunk, Tr::kBK, true);
  if (rc) return rc;
  // B: logical [K,N]; stored [K,N] (MN-major) or [N,K] (K-major).
  if (!kBMN)
    rc = encode_operand_map(&mb, Tr::kTmaType, sizeof(TIn), g.b, g.N, g.K, g.ldb, g.batch,
                            g.strideB, Tr::kBK, BN, false);
  else
    rc = encode_operand_map(&mb, Tr::kTmaType, sizeof(TIn), g.b, g.K, g.N, g.ldb, g.batch,
                            g.strideB, kChunk, Tr::kBK, true);
  if (rc) return rc;

  auto kern = gemm_tcgen05_kernel<TIn, TOut, kAMN, kBMN, BN>;
  static bool attr_set = false;  // per template instantiation
  constexpr size_t smem = gemm_smem_bytes<BN>();
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_last_error("cudaFuncSetAttribute(smem=%zu): %s", smem, cudaGetErrorString(e));
      return B200_INTERNAL;
    }
    attr_set = true;
  }
  GemmShape s;
  s.M = (int)g.M;
  s.N = (int)g.N;
  s.K = (int)g.K;
  s.batch = (int)g.batch;
  s.ldc = (int)g.ldc;
  s.strideC = g.strideC;
  const long long tiles = ((g.M + kBM - 1) / kBM) * ((g.N + BN - 1) / BN) * g.batch;
  // split-K when the output tiles alone cannot fill the SMs and scratch was provided
  const int num_kb = (int)((g.K + Tr::kBK - 1) / Tr::kBK);
  int splits = 1;
  if (g.workspace) {
    splits = plan_splits(tiles, num_kb);
    while (splits > 1 &&
           (size_t)splits * g.batch * g.M * g.N * sizeof(float) > g.workspace_bytes)
      --splits;
    if (splits < 1) splits = 1;
  }
  s.kb_per_split = (num_kb + splits - 1) / splits;
  splits = (num_kb + s.kb_per_split - 1) / s.kb_per_split;  // no empty split
  s.splits = splits;
  s.partial = static_cast<float*>(g.workspace);
  const long long work = tiles * splits;
  const int grid = (int)(work < sm_count() ? work : sm_count());
  const bool prof = profile_enabled();
  if (prof) profile_gemm_launch_begin(stream);
  kern<<<grid, kGemmThreads, smem, stream>>>(ma, mb, static_cast<TOut*>(g.c), s);
  if (prof) profile_gemm_launch_end(stream, 2.0 * (double)g.M * (double)g.N * (double)g.K * g.batch);
  note_launch();
  if (splits > 1) {
    const long long groups = g.batch * g.M * ((g.N + 3) / 4);
    splitk_reduce_kernel<TOut><<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(
        s.partial, static_cast<TOut*>(g.c), splits, g.batch, s.M, s.N, s.ldc, s.strideC);
    note_launch();
  }
  return check_launch("gemm_tcgen05");
}

template <typename TIn, typename TOut, int BN>
static int dispatch_major(const GemmArgs& g, cudaStream_t stream) {
  if (!g.a_mn_major && !g.b_mn_major) return launch_gemm<TIn, TOut, false, false, BN>(g, stream);
  if (!g.a_mn_major && g.b_mn_major) return launch_gemm<TIn, TOut, false, true, BN>(g, stream);
  if (g.a_mn_major && !g.b_mn_major) return launch_gemm<TIn, TOut, true, false, BN>(g, stream);
  return launch_gemm<TIn, TOut, true, true, BN>(g, stream);
}

// Split-K plan shared by the launcher and gemm_workspace_bytes(): how many K splits a 128-wide
// tiling would use for this shape (1 = none).
static int plan_splits(long long tiles, int num_kb) {
  if (tiles * 2 > sm_count() || num_kb < 8) return 1;
  long long splits = sm_count() / tiles;
  if (splits > num_kb / 4) splits = num_kb / 4;  // >= 4 K blocks per split
  if (splits > 16) splits = 16;
  return splits < 1 ? 1 : (int)splits;
}

size_t gemm_workspace_bytes(int dtype, long long M, long long N, long long K, long long batch) {
  if (M <= 0 || N <= 0 || K <= 0 || batch <= 0) return 0;
  const int bk = dtype == B200_DT_FLOAT ? GemmTraits<float>::kBK : GemmTraits<__nv_bfloat16>::kBK;
  const int bn = N <= 64 ? 64 : 128;
  const long long tiles = ((M + kBM - 1) / kBM) * ((N + bn - 1) / bn) * batch;
  const int splits = plan_splits(tiles, (int)((K + bk - 1) / bk));
  return splits > 1 ? (size_t)splits * batch * M * N * sizeof(float) : 0;
}

// Tile-N choice: fill the 148 SMs.  128x128 tiles; when those cannot fill the machine either
// split K (scratch available) or fall back to 128x64 tiles.
static int choose_bn(const GemmArgs& g) {
  if (g.force_bn == 64 || g.force_bn == 128 || g.force_bn == 256) return g.force_bn;
  if (g.N <= 64) return 64;
  const long long tm = (g.M + kBM - 1) / kBM;
  const long long t128 = tm * ((g.N + 127) / 128) * g.batch;
  if (t128 >= sm_count()) return 128;
  const int bk = g.dtype == B200_DT_FLOAT ? GemmTraits<float>::kBK : GemmTraits<__nv_bfloat16>::kBK;
  if (g.workspace && plan_splits(t128, (int)((g.K + bk - 1) / bk)) > 1) return 128;
  return 64;
}

bool gemm_tcgen05_supported(const GemmArgs& g) {
  const int e = g.dtype == B200_DT_FLOAT ? 4 : 2;
  const int align = 16 / e;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) return false;
  if (g.lda % align || g.ldb % align) return false;
  if (g.batch > 1 && (g.strideA % align || g.strideB % align)) return false;
  if ((reinterpret_cast<uintptr_t>(g.a) & 15) || (reinterpret_cast<uintptr_t>(g.b) & 15))
    return false;
  if (g.M > 0x7fffffffLL || g.N > 0x7fffffffLL || g.K > 0x7fffffffLL) return false;
  return true;
}

int gemm_tcgen05(const GemmArgs& g, cudaStream_t stream) {
  const int bn = choose_bn(g);
  if (g.dtype == B200_DT_FLOAT) {
    if (bn == 64) return dispatch_major<float, float, 64>(g, stream);
    if (bn == 256) return dispatch_major<float, float, 256>(g, stream);
    return dispatch_major<float, float, 128>(g, stream);
  } else if (g.dtype == B200_DT_BFLOAT16) {
    if (bn == 64) return dispatch_major<__nv_bfloat16, __nv_bfloat16, 64>(g, stream);
    if (bn == 256) return dispatch_major<__nv_bfloat16, __nv_bfloat16, 256>(g, stream);
    return dispatch_major<__nv_bfloat16, __nv_bfloat16, 128>(g, stream);
  }
  set_last_error("gemm_tcgen05: unsupported dtype %d", g.dtype);
  return B200_UNIMPLEMENTED;
}

}  // namespace b200
