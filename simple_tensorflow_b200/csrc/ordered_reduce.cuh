// Ordered cross-CTA sum of per-CTA partial rows (the tail of the single-launch bias-gradient
// kernels in reduce.cu and pool.cu).  Called by ONE CTA of 256 threads -- the one that took the last
// ticket, after its __threadfence() -- with partial[nb][C] fp32 rows in global memory, C a multiple
// of 4 and C <= 256.
//
// Thread t owns the 16-byte channel group t % (C/4) of the rows part, part + P, ... (P = 256 / (C/4)
// parts); a batch of eight rows is fetched with cp.async into private shared-memory slots (one
// memory latency per batch: ptxas keeps only ~3 register loads in flight, profiles/r02_notes.md
// 3.5) and added in ascending row order; the parts are then combined in part order.  Every
// association is fixed by (nb, C) alone: the result does not depend on which CTA runs this or on
// timing.  `slots` is 8 x 256 x 16 B of shared memory, `flat` at least 256 x 4 floats; both may
// alias buffers the caller no longer needs.
#pragma once
#include <cstdint>

namespace b200 {

__device__ __forceinline__ void ordered_cp_async16_zfill(void* smem_dst, const void* gsrc,
                                                         int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(
                   static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc), "r"(src_bytes)
               : "memory");
}

// Returns, in threads c < C, the sum over rows of partial[row][c]; other threads return 0.
__device__ __forceinline__ float ordered_partial_sum(const float* __restrict__ partial, int nb, int C,
                                                     uint4* slots /* [8][256] */,
                                                     float* flat /* [P][C] */) {
  const int t = threadIdx.x;
  const int G4 = C >> 2;          // 16-byte channel groups per row
  const int P = 256 / G4;         // row strands
  const int cg = t % G4, part = t / G4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (part < P) {
    for (int b0 = part; b0 < nb; b0 += 8 * P) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int b = b0 + u * P;
        ordered_cp_async16_zfill(&slots[u * 256 + t],
                                 partial + (long long)(b < nb ? b : nb - 1) * C + cg * 4,
                                 b < nb ? 16 : 0);
      }
      asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#pragma unroll
      for (int u = 0; u < 8; ++u) {  // own slots only; rows past the end were zero-filled
        const uint4 v = slots[u * 256 + t];
        acc.x += __uint_as_float(v.x);
        acc.y += __uint_as_float(v.y);
        acc.z += __uint_as_float(v.z);
        acc.w += __uint_as_float(v.w);
      }
    }
  }
  __syncthreads();  // slots may alias flat
  if (part < P) {
    float* dst = flat + part * C + cg * 4;
    dst[0] = acc.x;
    dst[1] = acc.y;
    dst[2] = acc.z;
    dst[3] = acc.w;
  }
  __syncthreads();
  float tot = 0.f;
  if (t < C)
    for (int q = 0; q < P; ++q) tot += flat[q * C + t];
  return tot;
}

}  // namespace b200
