// Hardware probe (measurement tool, not product code): does a tcgen05 shared-memory matrix
// descriptor whose start address is NOT aligned to the swizzle pattern need the "matrix base
// offset" field (bits 49-51), and with which value?  The halo-tile convolution addresses every
// filter tap as the same smem input tile shifted by whole pixel rows (128 B), so the answer decides
// its A / B descriptors.
//
//   test K : A K-major SWIZZLE_128B, rows = pixels; D_n = A[n : n+128, :] * B^T       (M128 N64 K32)
//   test MN: B MN-major SWIZZLE_128B_ATOM_32B (tf32), rows = pixels (K); D_n = A * B[n : n+32, :]
// for every row shift n and base-offset convention; the host compares with a CPU product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/desc_probe tools/desc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>

#include "../simple_tensorflow_b200/csrc/b200_ptx.cuh"

using namespace b200;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int kShifts = 20;
constexpr int kModes = 3;     // base offset = 0, n & 7, n & 3
constexpr int kARows = 160;   // rows of the shifted operand resident in smem

__device__ __forceinline__ uint64_t with_base_offset(uint64_t d, uint32_t bo) {
  return d | (static_cast<uint64_t>(bo & 7) << 49);
}

// test 0: K-major A shifted.  test 1: MN-major B shifted.
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap mapShift, const __grid_constant__ CUtensorMap mapFix,
             float* out, int test) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* smem = raw + (base - smem_u32(raw));
  uint8_t* smShift = smem;                        // kARows x 128 B
  uint8_t* smFix = smem + 24 * 1024;              // 128 x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 48 * 1024 + 64);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<64>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;
  const int fix_rows = test == 0 ? 64 : (test == 2 ? 32 : 128);
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar[0], kARows * 128 + fix_rows * 128);
    tma_load_2d(smShift, &mapShift, &bar[0], 0, 0);
    tma_load_2d(smFix, &mapFix, &bar[0], 0, 0);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  uint32_t phase = 0;
  const int N = test == 0 ? 64 : 32;
  for (int n = 0; n < kShifts; ++n) {
    for (int mode = 0; mode < kModes; ++mode) {
      const uint32_t bo = mode == 0 ? 0u : (mode == 1 ? (uint32_t)(n & 7) : (uint32_t)(n & 3));
      if (threadIdx.x == 0) {
        for (int k = 0; k < 4; ++k) {
          uint64_t ad, bd;
          uint32_t idesc;
          if (test == 0) {
            // A: K-major rows n.., +32 B per k step inside the swizzle row
            ad = with_base_offset(
                make_smem_desc_sw128(smem_u32(smShift) + n * 128 + k * 32, 16, 1024), bo);
            bd = make_smem_desc_sw128(smem_u32(smFix) + k * 32, 16, 1024);
            idesc = make_idesc(2, false, false, 128, 64);
          } else if (test == 2) {
            // A: MN-major, M = 4 chunks x 32 channels; chunk j = the shifted tile at rows n + j
            // (leading byte offset = ONE pixel row: overlapping chunks), K = 32 pixel rows.
            // B: MN-major [K = 32 pixel rows][N = 32], fixed.
            ad = with_base_offset(
                make_smem_desc_sw128(smem_u32(smShift) + n * 128 + k * 8 * 128, 128, 512, 1), bo);
            bd = make_smem_desc_sw128(smem_u32(smFix) + k * 8 * 128, 32 * 128, 512, 1);
            idesc = make_idesc(2, true, true, 128, 32);
          } else {
            // B: MN-major (N = 32 channels contiguous), K rows n.., 8 rows per k step
            ad = make_smem_desc_sw128(smem_u32(smFix) + k * 32, 16, 1024);
            bd = with_base_offset(
                make_smem_desc_sw128(smem_u32(smShift) + n * 128 + k * 8 * 128, 32 * 128, 512, 1),
                bo);
            idesc = make_idesc(2, false, true, 128, 32);
          }
          umma_tf32(tmem, ad, bd, idesc, k != 0);
        }
        umma_commit(&bar[1]);
      }
      mbar_wait(&bar[1], phase);
      phase ^= 1;
      tc_fence_after();
      uint32_t v[32];
      float* dst = out + ((size_t)(n * kModes + mode) * 128 + warp * 32 + lane) * 64;
      for (int c = 0; c < N; c += 32) {
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) dst[c + j] = __uint_as_float(v[j]);
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc<64>(tmem);
}

static float tf32(float x) {  // truncation of the low 13 mantissa bits, like kind::tf32
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

int main() {
  EncodeTiledFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) !=
          cudaSuccess || q != cudaDriverEntryPointSuccess) {
    printf("no cuTensorMapEncodeTiled\n");
    return 1;
  }
  std::vector<float> hShift(kARows * 32), hFix(128 * 32);
  srand(7);
  for (auto& v : hShift) v = tf32((rand() % 2001 - 1000) / 1000.0f);
  for (auto& v : hFix) v = tf32((rand() % 2001 - 1000) / 1000.0f);
  float *dShift, *dFix, *dOut;
  const size_t out_elems = (size_t)kShifts * kModes * 128 * 64;
  cudaMalloc(&dShift, hShift.size() * 4);
  cudaMalloc(&dFix, hFix.size() * 4);
  cudaMalloc(&dOut, out_elems * 4);
  cudaMemcpy(dShift, hShift.data(), hShift.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dFix, hFix.data(), hFix.size() * 4, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 52 * 1024);
  for (int test = 0; test < 3; ++test) {
    CUtensorMap mShift, mFix;
    cuuint64_t gdim[2] = {32, (cuuint64_t)kARows};
    cuuint64_t gstr[1] = {128};
    cuuint32_t box[2] = {32, (cuuint32_t)kARows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&mShift, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dShift, gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        test == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    const int fix_rows = test == 0 ? 64 : (test == 2 ? 32 : 128);
    cuuint64_t gdim2[2] = {32, (cuuint64_t)fix_rows};
    cuuint32_t box2[2] = {32, (cuuint32_t)fix_rows};
    CUresult r2 = encode(&mFix, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dFix, gdim2, gstr, box2, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         test == 2 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) {
      printf("encode failed %d %d\n", (int)r, (int)r2);
      return 1;
    }
    cudaMemset(dOut, 0, out_elems * 4);
    probe_kernel<<<1, 128, 52 * 1024>>>(mShift, mFix, dOut, test);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("test %d: kernel failed: %s\n", test, cudaGetErrorString(e));
      return 1;
    }
    std::vector<float> hOut(out_elems);
    cudaMemcpy(hOut.data(), dOut, out_elems * 4, cudaMemcpyDeviceToHost);
    printf("== test %s: max abs error per (row shift n, base-offset mode 0 | n&7 | n&3)\n",
           test == 0 ? "K-major A, SWIZZLE_128B"
                     : (test == 1 ? "MN-major B, SWIZZLE_128B_ATOM_32B"
                                  : "MN-major A, M = 4 overlapping chunks (LBO = 128 B), ATOM_32B"));
    for (int n = 0; n < kShifts; ++n) {
      printf("n=%2d:", n);
      for (int mode = 0; mode < kModes; ++mode) {
        double worst = 0;
        const float* D = hOut.data() + (size_t)(n * kModes + mode) * 128 * 64;
        const int N = test == 0 ? 64 : 32;
        for (int i = 0; i < 128; ++i)
          for (int j = 0; j < N; ++j) {
            double ref = 0;
            for (int k = 0; k < 32; ++k) {
              if (test == 0)
                ref += (double)hShift[(n + i) * 32 + k] * hFix[j * 32 + k];
              else if (test == 2)  // row i = (chunk i / 32, channel i % 32); K index = pixel row
                ref += (double)hShift[(n + i / 32 + k) * 32 + (i % 32)] * hFix[k * 32 + j];
              else
                ref += (double)hFix[i * 32 + k] * hShift[(n + k) * 32 + j];
            }
            worst = fmax(worst, fabs(ref - D[i * 64 + j]));
          }
        printf("  %s(%.3g)", worst < 1e-3 ? "OK " : "BAD", worst);
      }
      printf("\n");
    }
  }
  return 0;
}
