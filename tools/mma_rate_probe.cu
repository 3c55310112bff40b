// Hardware probe (measurement tool, not product code): how long does one tcgen05.mma last as a
// function of its shape?  A single CTA issues `reps` back-to-back MMAs with CONSTANT descriptors
// (so the issue path is one UTCHMMA per iteration) from one elected lane, commits, waits, and
// reports cycles per MMA.  Shapes: M in {64, 128}, N in {32, 64, 128, 256}, tf32 (K = 8) and
// bf16 (K = 16), operands in shared memory (SS mode), K-major SWIZZLE_128B tiles.
// Also dumps where the rows of an M = 64 accumulator live in TMEM (which lanes).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o build/mma_rate_probe tools/mma_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../simple_tensorflow_b200/csrc/b200_ptx.cuh"

using namespace b200;

__global__ void __launch_bounds__(128, 1)
rate_kernel(long long* out, int M, int N, int fmt, int reps, int nacc) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* smem = raw + (base - smem_u32(raw));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 64);
  const int warp = uniform_warp_idx();
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
  if (warp == 0) {
    const uint32_t idesc = make_idesc((uint32_t)fmt, false, false, (uint32_t)M, (uint32_t)N);
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem) + 32 * 1024, 16, 1024);
    __syncwarp();
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const uint32_t d = tmem + (uint32_t)((i % nacc) * N);
      if (fmt == 2)
        umma_tf32_elect(d, ad, bd, idesc, 1u);
      else
        umma_f16_elect(d, ad, bd, idesc, 1u);
    }
    umma_commit_elect(&bar[0]);
    const long long t1 = clock64();
    mbar_wait(&bar[0], 0);
    const long long t2 = clock64();
    if ((threadIdx.x & 31) == 0) {
      out[0] = t1 - t0;  // issue time
      out[1] = t2 - t0;  // until all MMAs retired
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

// M = 64 layout: A row i (K-major) holds the value i + 1 in its first K element, B row j holds 1 in
// its first K element -> D[i][j] = i + 1.  Every warp then reads its 32 TMEM lanes, column 0.
__global__ void __launch_bounds__(128, 1) layout_kernel(float* out, int M) {
  extern __shared__ uint8_t raw[];
  const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  uint8_t* smem = raw + (base - smem_u32(raw));
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 96 * 1024);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + 96 * 1024 + 64);
  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  // K-major SW128 tile: row r at r * 128 bytes; element (r, k = 0) sits in 16-byte chunk (0 ^ (r & 7))
  if (threadIdx.x < 128) {
    const int r = threadIdx.x;
    float* a = reinterpret_cast<float*>(smem + r * 128 + ((0 ^ (r & 7)) << 4));
    a[0] = (float)(r + 1);
    if (r < 64) {
      float* b = reinterpret_cast<float*>(smem + 32 * 1024 + r * 128 + ((0 ^ (r & 7)) << 4));
      b[0] = 1.0f;
    }
  }
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tslot, 0);
  // clear all lanes' column range first (M = 128 MMA with zero operands would do, simpler: st via ld? skip)
  if (warp == 0) {
    const uint32_t idesc = make_idesc(2, false, false, (uint32_t)M, 64);
    const uint64_t ad = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
    const uint64_t bd = make_smem_desc_sw128(smem_u32(smem) + 32 * 1024, 16, 1024);
    umma_tf32_elect(tmem, ad, bd, idesc, 0u);
    umma_commit_elect(&bar[0]);
  }
  mbar_wait(&bar[0], 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  out[threadIdx.x * 2] = __uint_as_float(v[0]);
  out[threadIdx.x * 2 + 1] = __uint_as_float(v[1]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(layout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int reps = 512;
  printf("cycles per tcgen05.mma (SS mode, %d back-to-back, 1 CTA); nacc = accumulators rotated\n", reps);
  for (int fmt : {2, 1}) {
    for (int M : {128, 64}) {
      for (int N : {32, 64, 128, 256}) {
        for (int nacc : {1, 2}) {
          if (nacc * N > 512) continue;
          rate_kernel<<<1, 128, 100 * 1024>>>(d, M, N, fmt, 8, nacc);  // warm-up
          rate_kernel<<<1, 128, 100 * 1024>>>(d, M, N, fmt, reps, nacc);
          long long h[2];
          if (cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost) != cudaSuccess) {
            printf("failed: %s\n", cudaGetErrorString(cudaGetLastError()));
            return 1;
          }
          const int K = fmt == 2 ? 8 : 16;
          printf("%s M=%3d N=%3d nacc=%d: issue %.1f cyc/MMA, retire %.1f cyc/MMA  -> %.0f flop/cyc\n",
                 fmt == 2 ? "tf32" : "bf16", M, N, nacc, (double)h[0] / reps, (double)h[1] / reps,
                 2.0 * M * N * K * reps / (double)h[1]);
        }
      }
    }
  }
  for (int M : {64, 128}) {
    float* o;
    cudaMalloc(&o, 256 * 4);
    cudaMemset(o, 0, 256 * 4);
    layout_kernel<<<1, 128, 100 * 1024>>>(o, M);
    float h[256];
    cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
    printf("M=%d accumulator: TMEM lane -> D row (column 0 value - 1; column 1 should be equal), -1 = untouched/other\n", M);
    for (int l = 0; l < 128; ++l) printf("%s%d", l % 32 == 0 ? "\n  " : " ", (int)h[l * 2] - 1);
    printf("\n");
  }
  return 0;
}
