// Halo-tile implicit-GEMM convolution for sm_100a (stride 1): forward and, through the flipped
// filter, the input gradient.
//
//   out[n, oh, ow, k] = sum_{r, s, c} in[n, oh + r - pt, ow + s - pl, c] * w[r, s, c, k]
//
// Replaces the cuDNN forward / backward-data calls of LaunchConv2DOp<GPUDevice,T>::launch
// (tensorflow/core/kernels/conv_ops.cc:433-720) and Conv2DSlowBackpropInputOp<GPUDevice,T>
// (conv_grad_input_ops.cc:533-917) for unit-stride convolutions, NHWC x HWIO -> NHWC, with no
// layout shuffles and no patch matrix.
//
// Why it exists (profiles/r01_notes.md): the im2col-mode TMA formulation fetches
// every input pixel once per filter tap -- 12-25x the algorithmic bytes over the L2->SM crossbar,
// tensor pipe 9-14 % active on LeNet conv2.  Here a CTA loads the (bh + R - 1) x (bw + S - 1)
// input halo of a bh x bw output tile ONCE (one tiled TMA box per 128-byte channel block, zero
// fill supplies the padding) and addresses every filter tap as the same shared-memory tile shifted
// by whole pixel rows:
//
//   * output pixels are enumerated in padded-row coordinates p' = r' * WP + c' (WP = bw + S - 1),
//     so the A rows of tap (r, s) are rows p' + r * WP + s of the halo tile: an affine shift, i.e.
//     the SAME K-major SWIZZLE_128B tile with the descriptor start address advanced by
//     (r * WP + s) * 128 bytes.  tcgen05 applies the 128-byte swizzle to absolute shared-memory
//     address bits, so a descriptor may start at any 128-byte row (tools/desc_probe.cu,
//     profiles/r02_desc_probe.txt); the S - 1 junk columns per row cost (S - 1) / WP extra MMA
//     work and are never stored.
//   * the filter tap tiles [C-block rows, BN columns] (HWIO is already the MN-major B operand)
//     stream through a TMA ring; a cluster of up to 4 CTAs loads each tile once and multicasts it,
//     so the per-SM filter traffic drops by the cluster size.
//   * accumulators: n_mtiles x BN fp32 columns of TMEM, double buffered (2 x 256 columns) so the
//     epilogue of one tile overlaps the MMAs of the next.
//
// CTA layout (224 threads, persistent, one CTA per SM): warp 0 = TMA producer, warp 1 = TMEM
// allocator + tcgen05.mma issuer, warps 2..5 = epilogue (TMEM -> registers -> [bias, relu] ->
// swizzled smem -> coalesced 16-byte global stores of the valid pixel rows), warp 6 = second
// tcgen05.mma issuer.  Two issuers because an M128 x N64 x K8 MMA lasts 32 cycles but costs its
// issuing warp ~130 (five R2UR moves of descriptor words into uniform registers, measured): each
// issuer owns the M tiles mt = issuer, issuer + 2, ... and therefore its own TMEM accumulators, so
// no two warps ever accumulate into the same columns and the summation order stays fixed.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "b200_internal.h"
#include "b200_ptx.cuh"

namespace b200 {

namespace {

constexpr int kIssueWarps = 2;   // MMA-issuing warps: warp 1 and warp 6 (see the kernel comment)
constexpr int kHaloThreads = 192 + 32 * (kIssueWarps - 1);
constexpr int kMaxBStages = 8;
constexpr int kRowBytes = 128;            // one pixel row of one channel block
constexpr int kAccCols = 256;             // TMEM columns per accumulator stage
constexpr int kEpiBuf = 32 * kRowBytes;   // 32 rows x 128 B staging buffer

struct HaloShape {
  int N, H, W, C, K;     // input NHWC, output channels
  int R, S, pt, pl;      // filter, leading pads (stride 1)
  int OH, OW;
  int bh, bw, tiles_h, tiles_w;
  int WP, HP;            // halo tile: bw + S - 1, bh + R - 1
  int n_mtiles;          // ceil(bh * WP / 128)
  int cblocks;           // C / (128 B of channels)
  int nblocks;           // ceil(K / BN)
  int halo_rows;         // rows allocated per channel block (multiple of 8, covers the shifts)
  int cl;                // cluster size: filter tiles are multicast across it
  int b_stages;          // filter ring depth
  long long items_per_nb, padded_per_nb, work_padded;
  const void* bias;      // optional fused BiasAdd (+ bias[k]), element type TOut
  int relu;              // optional fused Relu
};

template <typename T>
struct HaloTraits;
template <>
struct HaloTraits<float> {
  static constexpr int kChunk = 32;       // elements per 128 bytes
  static constexpr int kUmmaK = 8;
  static constexpr uint32_t kFormat = 2;  // TF32
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  static constexpr bool kMn32 = true;     // MN-major tf32: 32-byte swizzle atoms
};
template <>
struct HaloTraits<__nv_bfloat16> {
  static constexpr int kChunk = 64;
  static constexpr int kUmmaK = 16;
  static constexpr uint32_t kFormat = 1;  // BF16
  static constexpr CUtensorMapDataType kTmaType = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  static constexpr bool kMn32 = false;
};

__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                               int c0, int c1, int c2, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      ".multicast::cluster [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64"
      " [%0], %1;" ::"r"(smem_u32(bar)),
      "h"(mask)
      : "memory");
}

struct WorkItem {
  int n, oh0, ow0, nb;
  bool active;
};
__device__ __forceinline__ WorkItem decode_work(const HaloShape& s, long long w) {
  WorkItem it;
  const long long nb = w / s.padded_per_nb;
  long long sp = w - nb * s.padded_per_nb;
  it.nb = (int)nb;
  it.active = sp < s.items_per_nb;
  if (!it.active) sp = 0;
  const int tw = (int)(sp % s.tiles_w);
  sp /= s.tiles_w;
  const int th = (int)(sp % s.tiles_h);
  it.n = (int)(sp / s.tiles_h);
  it.oh0 = th * s.bh;
  it.ow0 = tw * s.bw;
  return it;
}

// kPair: the two CTAs of a cluster form a tcgen05 CTA pair (cta_group::2).  A tcgen05.mma occupies
// the tensor pipe ~160 cycles whatever its N (tools/mma_rate_probe.cu: M128 x N32..256 all retire in
// 160 cycles), so with N = K_out <= 64 the only way to more work per instruction is M: the leader
// issues M = 256 MMAs over both CTAs' halo tiles (each CTA its own image, same smem offsets) and
// each CTA stages half of the filter tile's columns.  Halves the MMA count per SM.
template <typename TIn, typename TOut, int BN, bool kPair>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmapX, const __grid_constant__ CUtensorMap tmapW,
                 TOut* __restrict__ out, const HaloShape s) {
  pdl_launch_dependents();
  using Tr = HaloTraits<TIn>;
  constexpr int kChunk = Tr::kChunk;               // channels per block = B rows per stage
  constexpr int kNChunks = BN / kChunk;            // 128-byte column chunks of a filter tile
  constexpr int kMyChunks = kPair ? kNChunks / 2 : kNChunks;  // filter chunks staged by this CTA
  constexpr int kBStageBytes = kMyChunks * kChunk * kRowBytes;
  constexpr uint32_t kIdesc = make_idesc(Tr::kFormat, false, true, kPair ? 256 : 128, BN);
  static_assert(!kPair || kNChunks % 2 == 0, "pair mode splits the filter chunks between the CTAs");
  constexpr int kEpiCols = kRowBytes / (int)sizeof(TOut);  // output columns per staged 128-B row
  constexpr int kLdPerIter = kEpiCols / 32;
  static_assert(BN % kChunk == 0 && BN <= kAccCols, "BN");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
  const int halo_bytes = s.cblocks * s.halo_rows * kRowBytes;  // one buffer, all channel blocks
  uint8_t* smHalo = smem;                                      // [2][cblocks][halo_rows][128]
  uint8_t* smB = smem + 2 * halo_bytes;                        // [b_stages][kBStageBytes]
  uint8_t* smEpi = smB + s.b_stages * kBStageBytes;            // [4 warps][2][32 x 128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smEpi + 4 * 2 * kEpiBuf);
  uint64_t* halo_full = bars;                  // [2]
  uint64_t* halo_empty = bars + 2;             // [2]
  uint64_t* b_full = bars + 4;                 // [kMaxBStages]
  uint64_t* b_empty = bars + 4 + kMaxBStages;  // [kMaxBStages]
  uint64_t* tfull = bars + 4 + 2 * kMaxBStages;   // [2]
  uint64_t* tempty = bars + 6 + 2 * kMaxBStages;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * kMaxBStages);

  const int warp = uniform_warp_idx(), lane = threadIdx.x & 31;
  const int cl = s.cl;
  const uint32_t rank = cl > 1 ? cluster_ctarank() : 0;
  const long long first = (long long)(blockIdx.x / cl) * cl + rank;
  const long long stride = (long long)gridDim.x;  // clusters * cl
  const uint16_t mask = (uint16_t)((1u << cl) - 1u);
  const int taps = s.R * s.S;
  const bool leader = !kPair || rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmapX);
    tma_prefetch_desc(&tmapW);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < 2; ++i) {
        mbar_init(&halo_full[i], kPair ? 2 : 1);  // pair: one producer arrive per CTA (leader's is used)
        mbar_init(&halo_empty[i], kIssueWarps);
        mbar_init(&tfull[i], kIssueWarps);
        mbar_init(&tempty[i], kPair ? 8 : 4);     // pair: both CTAs' epilogue warps (leader's is used)
      }
      for (int i = 0; i < kMaxBStages; ++i) {
        mbar_init(&b_full[i], kPair ? 2 : 1);
        // one commit per issuing warp: of every CTA (multicast filter tiles) or of the leader (pair)
        mbar_init(&b_empty[i], (kPair ? 1 : cl) * kIssueWarps);
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (kPair)
      tmem_alloc_2cta<512>(tmem_slot);
    else
      tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (cl > 1)
    cluster_sync_all();  // peers' barriers exist before any multicast / remote commit
  else
    __syncthreads();
  tc_fence_after();
  // (the shuffle makes the value provably warp-uniform: uniform-register descriptor math)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();  // everything above overlapped the previous kernel's tail

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, hb = 0, hphase = 0;
      // this CTA's share of every filter tile: pieces = chunk x row-part
      const int row_parts = cl > kNChunks ? cl / kNChunks : 1;
      const int pieces = kNChunks * row_parts;
      const int per_cta = pieces / cl;  // >= 1 (host guarantees divisibility)
      const int rows_per_piece = kChunk / row_parts;
      // The halo tile of item i + 1 is requested BEFORE the filter tiles of item i: the filter
      // ring throttles the producer to a few taps ahead of the MMAs, so a halo load issued after
      // it would start only when item i is almost finished and its latency would be exposed.
      // pair mode: `h` is this CTA's item, `peer_active` whether the other CTA of the pair loads too
      auto issue_halo = [&](const WorkItem& h, bool peer_active) {
        mbar_wait(&halo_empty[hb], hphase ^ 1);
        const uint32_t bytes = (uint32_t)(s.cblocks * s.HP * s.WP * kRowBytes);
        if (!kPair) {
          mbar_expect_tx(&halo_full[hb], bytes);
        } else if (leader) {
          mbar_expect_tx(&halo_full[hb], (h.active ? bytes : 0u) + (peer_active ? bytes : 0u));
        } else {
          mbar_arrive_remote(&halo_full[hb], 0);
        }
        if (h.active) {
          for (int cb = 0; cb < s.cblocks; ++cb) {
            uint8_t* dst = smHalo + hb * halo_bytes + cb * s.halo_rows * kRowBytes;
            if (kPair)
              tma_load_4d_2cta(dst, &tmapX, &halo_full[hb], cb * kChunk, h.ow0 - s.pl,
                               h.oh0 - s.pt, h.n);
            else
              tma_load_4d(dst, &tmapX, &halo_full[hb], cb * kChunk, h.ow0 - s.pl, h.oh0 - s.pt,
                          h.n);
          }
        }
        if (++hb == 2) {
          hb = 0;
          hphase ^= 1;
        }
      };
      // the pair advances in lockstep: the leader's item is w, the peer's w + 1 (both or neither
      // of a pair may be padding, never the leader alone)
      auto pair_active = [&](long long w) {
        const long long base_w = kPair ? w - (long long)rank : w;
        return decode_work(s, base_w).active;  // the leader's item is active <=> the halo is needed
      };
      auto request_halo = [&](long long w) {
        const WorkItem h = decode_work(s, w);
        if (!kPair) {
          if (h.active) issue_halo(h, false);
        } else if (pair_active(w)) {
          const long long peer_w = rank == 0 ? w + 1 : w - 1;
          issue_halo(h, decode_work(s, peer_w).active);
        }
      };
      if (first < s.work_padded) request_halo(first);
      for (long long w = first; w < s.work_padded; w += stride) {
        const WorkItem it = decode_work(s, w);
        if (w + stride < s.work_padded) request_halo(w + stride);
        for (int cb = 0; cb < s.cblocks; ++cb) {
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_empty[stage], phase ^ 1);
            const int krow = tap * s.C + cb * kChunk;      // HWIO row of this (tap, channel block)
            uint8_t* dst = smB + stage * kBStageBytes;
            if (kPair) {
              // each CTA stages its half of the tile's column chunks; bytes credit the leader
              if (leader)
                mbar_expect_tx(&b_full[stage], 2 * kBStageBytes);
              else
                mbar_arrive_remote(&b_full[stage], 0);
              for (int c = 0; c < kMyChunks; ++c)
                tma_load_3d_2cta(dst + c * (kChunk * kRowBytes), &tmapW, &b_full[stage], 0, krow,
                                 it.nb * kNChunks + (int)rank * kMyChunks + c);
            } else {
              mbar_expect_tx(&b_full[stage], kBStageBytes);  // the whole tile lands here (multicast)
              for (int p = 0; p < per_cta; ++p) {
                const int piece = (int)rank * per_cta + p;
                const int chunk = piece / row_parts, part = piece - chunk * row_parts;
                uint8_t* d = dst + chunk * (kChunk * kRowBytes) + part * rows_per_piece * kRowBytes;
                const int r0 = krow + part * rows_per_piece;
                const int nchunk = it.nb * kNChunks + chunk;
                if (cl > 1)
                  tma_load_3d_mc(d, &tmapW, &b_full[stage], 0, r0, nchunk, mask);
                else
                  tma_load_3d(d, &tmapW, &b_full[stage], 0, r0, nchunk);
              }
            }
            if (++stage == (uint32_t)s.b_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1 || warp >= 6) {
    // ===================== MMA issuers =====================
    // the whole warp runs this loop; one elected lane issues each tcgen05 instruction
    const int iw = warp == 1 ? 0 : warp - 5;  // issuer index: owns M tiles iw, iw + kIssueWarps, ...
    if (leader) {  // pair mode: the leader CTA issues for both CTAs
      uint32_t stage = 0, phase = 0, hb = 0, hphase = 0, acc = 0, acc_phase = 0;
      for (long long w = first; w < s.work_padded; w += stride) {
        const WorkItem it = decode_work(s, w);
        if (it.active) {
          mbar_wait(&halo_full[hb], hphase);
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          tc_fence_after();
        }
        // Descriptors differ only in their start-address field (bits 0-13, address >> 4): build
        // them once and add offsets -- one thread issues every MMA and an M128 x N64 x K8 MMA
        // lasts 32 cycles, so the per-MMA instruction count is on the critical path.
        const uint64_t adesc_hb =
            make_smem_desc_sw128(smem_u32(smHalo + hb * halo_bytes), 16, 1024);
        const uint64_t bdesc_0 = make_smem_desc_sw128(smem_u32(smB), kChunk * kRowBytes,
                                                      Tr::kMn32 ? 512 : 1024, Tr::kMn32 ? 1 : 2);
        constexpr uint32_t kAStepK = (Tr::kUmmaK * (int)sizeof(TIn)) >> 4;   // +32 B per k step
        constexpr uint32_t kAStepMt = (128 * kRowBytes) >> 4;                 // next 128 rows
        constexpr uint32_t kBStepK = (Tr::kUmmaK * kRowBytes) >> 4;           // kUmmaK rows
        for (int cb = 0; cb < s.cblocks; ++cb) {
          const uint64_t adesc_cb = adesc_hb + (uint64_t)((cb * s.halo_rows * kRowBytes) >> 4);
          int r = 0, sx = 0;
          for (int tap = 0; tap < taps; ++tap) {
            mbar_wait(&b_full[stage], phase);
            __syncwarp();  // the spin loop above is per-lane: tell ptxas the warp is converged
            tc_fence_after();
            if (it.active) {
              const uint64_t adesc_tap = adesc_cb + (uint64_t)(((r * s.WP + sx) * kRowBytes) >> 4);
              const uint64_t bdesc_st = bdesc_0 + (uint64_t)((stage * kBStageBytes) >> 4);
              const uint32_t a_hi = (uint32_t)(adesc_tap >> 32), b_hi = (uint32_t)(bdesc_st >> 32);
              const uint32_t a_lo0 = (uint32_t)adesc_tap, b_lo0 = (uint32_t)bdesc_st;
              const uint32_t d0 = tmem_base + acc * kAccCols;
              const uint32_t acc_first = (cb | tap) == 0 ? 0u : 1u;
              // k outer, M tile inner: the B descriptor changes once per k step, the address
              // fields never carry into the high words (smem addresses < 256 KB)
#pragma unroll
              for (int k = 0; k < kChunk / Tr::kUmmaK; ++k) {
                const uint32_t b_lo = b_lo0 + k * kBStepK;
                const uint32_t accum = k == 0 ? acc_first : 1u;
                for (int mt = iw; mt < s.n_mtiles; mt += kIssueWarps) {
                  const uint32_t a_lo = a_lo0 + mt * kAStepMt + k * kAStepK;
                  if (kPair) {
                    if (sizeof(TIn) == 4)
                      umma_tf32_elect_lohi_2cta(d0 + mt * BN, a_lo, a_hi, b_lo, b_hi, kIdesc, accum);
                    else
                      umma_f16_elect_lohi_2cta(d0 + mt * BN, a_lo, a_hi, b_lo, b_hi, kIdesc, accum);
                  } else if (sizeof(TIn) == 4) {
                    umma_tf32_elect_lohi(d0 + mt * BN, a_lo, a_hi, b_lo, b_hi, kIdesc, accum);
                  } else {
                    umma_f16_elect_lohi(d0 + mt * BN, a_lo, a_hi, b_lo, b_hi, kIdesc, accum);
                  }
                }
              }
            }
            if (++sx == s.S) {
              sx = 0;
              ++r;
            }
            // the filter slot is free (in every CTA of the cluster) once these MMAs retire
            if (kPair)
              umma_commit_elect_2cta(&b_empty[stage]);
            else if (cl > 1)
              umma_commit_mc_elect(&b_empty[stage], mask);
            else
              umma_commit_elect(&b_empty[stage]);
            if (++stage == (uint32_t)s.b_stages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (it.active) {
          if (kPair) {
            umma_commit_elect_2cta(&tfull[acc]);      // accumulators complete -> both epilogues
            umma_commit_elect_2cta(&halo_empty[hb]);  // halo tiles consumed -> both producers
          } else {
            umma_commit_elect(&tfull[acc]);        // accumulators complete -> epilogue
            umma_commit_elect(&halo_empty[hb]);    // halo tile consumed -> producer may refill it
          }
          if (++acc == 2) {
            acc = 0;
            acc_phase ^= 1;
          }
          if (++hb == 2) {
            hb = 0;
            hphase ^= 1;
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps =====================
    const int quad = warp & 3;
    uint32_t acc = 0, acc_phase = 0, ebuf = 0;
    uint8_t* my_stage = smEpi + quad * 2 * kEpiBuf;
    const int sub = lane >> 3, c16 = lane & 7;  // copy-out: 4 rows x 8 x 16 B per instruction
    for (long long w = first; w < s.work_padded; w += stride) {
      const WorkItem it = decode_work(s, w);
      // pair mode: the peer of an active leader takes part in the TMEM hand-shake even when its
      // own item is padding (the leader counts both CTAs' epilogue warps)
      const bool engaged = kPair ? decode_work(s, w - (long long)rank).active : it.active;
      if (!engaged) continue;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (!it.active) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(&tempty[acc], 0);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }
      const int vh = min(s.bh, s.OH - it.oh0), vw = min(s.bw, s.OW - it.ow0);  // valid extent
      const int n0 = it.nb * BN;
      for (int mt = 0; mt < s.n_mtiles; ++mt) {
        const int prow0 = mt * 128 + quad * 32;  // first padded pixel of this warp's 32 rows
        if (prow0 >= vh * s.WP) continue;        // warp-uniform: nothing valid in these rows
        // copy-out geometry of this lane's 8 pixel rows, once per M tile (integer divisions)
        long long poff[8];
#pragma unroll
        for (int itr = 0; itr < 8; ++itr) {
          const int pp = prow0 + itr * 4 + sub;
          const int rr = pp / s.WP, cc = pp - rr * s.WP;
          poff[itr] = (rr < vh && cc < vw)
                          ? (((long long)it.n * s.OH + it.oh0 + rr) * s.OW + it.ow0 + cc) * s.K
                          : -1;
        }
#pragma unroll 1
        for (int c = 0; c < BN / kEpiCols; ++c) {
          const int col = n0 + c * kEpiCols;
          if (col >= s.K) break;
          uint32_t v[32 * kLdPerIter];
          {
            uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
            const uint32_t taddr =
                tmem_base + ((uint32_t)(quad * 32) << 16) + acc * kAccCols + mt * BN + c * kEpiCols;
            tmem_ld_32x32(taddr, v0);
            if (kLdPerIter == 2) {
              uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32 * (kLdPerIter - 1)]);
              tmem_ld_32x32(taddr + 32, v1);
            }
            tmem_ld_wait();
          }
          if (s.bias != nullptr) {
            const TOut* bp = static_cast<const TOut*>(s.bias) + col;
#pragma unroll
            for (int j = 0; j < kEpiCols; ++j)
              if (col + j < s.K) {
                float b;
                if (sizeof(TOut) == 4)
                  b = __ldg(reinterpret_cast<const float*>(bp) + j);
                else
                  b = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(bp)[j]);
                v[j] = __float_as_uint(__uint_as_float(v[j]) + b);
              }
          }
          if (s.relu) {
#pragma unroll
            for (int j = 0; j < kEpiCols; ++j) {
              const float f = __uint_as_float(v[j]);
              v[j] = __float_as_uint(f > 0.f ? f : 0.f);
            }
          }
          uint32_t wv[32];
          if (sizeof(TOut) == 4) {
#pragma unroll
            for (int j = 0; j < 32; ++j) wv[j] = v[j];
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(v[(2 * j) % (32 * kLdPerIter)]),
                                                       __uint_as_float(v[(2 * j + 1) % (32 * kLdPerIter)]));
              wv[j] = *reinterpret_cast<uint32_t*>(&h);
            }
          }
          uint8_t* buf = my_stage + ebuf * kEpiBuf;
          const uint32_t rbase = smem_u32(buf) + lane * kRowBytes;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t addr = rbase + (uint32_t)((j ^ (lane & 7)) << 4);  // 128B swizzle
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(wv[4 * j]),
                         "r"(wv[4 * j + 1]), "r"(wv[4 * j + 2]), "r"(wv[4 * j + 3])
                         : "memory");
          }
          __syncwarp();
          // copy-out: each 8-lane group writes one pixel's 128 contiguous bytes; junk pixels
          // (columns >= valid width, rows >= valid height of the tile) are skipped
          const bool vec_ok = (s.K * (int)sizeof(TOut)) % 16 == 0 &&
                              (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                              col + kEpiCols <= s.K;
#pragma unroll
          for (int itr = 0; itr < 8; ++itr) {
            const int row = itr * 4 + sub;
            if (poff[itr] >= 0) {
              uint32_t x0, x1, x2, x3;
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                           : "r"(smem_u32(buf) + row * kRowBytes +
                                 (uint32_t)((c16 ^ (row & 7)) << 4)));
              TOut* dst = out + poff[itr] + col + c16 * (16 / (int)sizeof(TOut));
              if (vec_ok) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(x0, x1, x2, x3);
              } else {
                const uint32_t xs[4] = {x0, x1, x2, x3};
                constexpr int kPer = 16 / (int)sizeof(TOut);
#pragma unroll
                for (int e = 0; e < kPer; ++e) {
                  const int cj = col + c16 * kPer + e;
                  if (cj < s.K) {
                    if (sizeof(TOut) == 4) {
                      reinterpret_cast<uint32_t*>(dst)[e] = xs[e];
                    } else {
                      const uint32_t wd = xs[e / 2];
                      reinterpret_cast<uint16_t*>(dst)[e] =
                          (uint16_t)((e & 1) ? (wd >> 16) : (wd & 0xFFFFu));
                    }
                  }
                }
              }
            }
          }
          __syncwarp();  // the buffer may be overwritten two chunks from now
          ebuf ^= 1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (kPair && !leader)
          mbar_arrive_remote(&tempty[acc], 0);  // the leader's MMA warps wait on it
        else
          mbar_arrive(&tempty[acc]);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  if (cl > 1)
    cluster_sync_all();  // no CTA exits while a peer may still multicast into it / signal it
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (kPair)
      tmem_dealloc_2cta<512>(tmem_base);
    else
      tmem_dealloc<512>(tmem_base);
  }
}

// ------------------------------------------------------------------ host side
struct HaloPlan {
  HaloShape s;
  size_t smem;
  int bn;
  bool ok;
};

constexpr size_t kSmemLimit = 227 * 1024;

static size_t halo_smem_bytes(const HaloShape& s, int bn, int es) {
  const int chunk = kRowBytes / es;
  const size_t bstage = (size_t)(bn / chunk) * chunk * kRowBytes;
  return 2 * (size_t)s.cblocks * s.halo_rows * kRowBytes + (size_t)s.b_stages * bstage +
         4 * 2 * kEpiBuf + (8 + 2 * kMaxBStages) * 8 + 16 + 1024;
}

// Pick the output tile (bh x bw): every work item is n_mtiles x 128 padded pixels with
// n_mtiles * BN <= 256 TMEM columns; prefer little junk (row padding S - 1, tile tails) and, for
// equal efficiency, more pixels per filter pass.
static HaloPlan plan_halo(int dtype, int N, int H, int W, int C, int K, int R, int S, int pt,
                          int pl, int OH, int OW) {
  HaloPlan best{};
  best.ok = false;
  const int es = dtype == B200_DT_FLOAT ? 4 : 2;
  const int chunk = kRowBytes / es;
  if (C % chunk != 0 || K % chunk != 0 || C <= 0 || K <= 0) return best;
  int bn = K >= 256 ? 256 : K;             // K is a multiple of chunk (32 / 64)
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) {
    bn = K % 128 == 0 ? 128 : (K % 64 == 0 ? 64 : chunk);
  }
  if (bn < chunk) return best;
  // A tcgen05.mma costs the same ~160 cycles for N = 32 as for N = 256, so a one-chunk N is padded
  // to two chunks (the filter map zero-fills the missing columns, the epilogue never stores them):
  // CTA pairs then split the chunks and issue M = 256 MMAs.
  static const bool no_pad_n = getenv("B200TF_CONV_HALO_NO_PAD_N") != nullptr;  // comparison knob
  if (bn == chunk && 2 * chunk <= 256 && !no_pad_n) bn = 2 * chunk;
  const int max_mt = kAccCols / bn;
  double best_score = -1.0;
  const int bw_cands[6] = {OW, 126, 62, 30, 14, 6};
  for (int bi = 0; bi < 6; ++bi) {
    const int bw = std::min(OW, bw_cands[bi]);
    const int WP = bw + S - 1;
    if (WP > 256) continue;
    for (int bh = 1; bh <= std::min(OH, 256 - R + 1); ++bh) {
      const int HP = bh + R - 1;
      const int mt = (bh * WP + 127) / 128;
      if (mt > max_mt) break;
      HaloShape s{};
      s.cblocks = C / chunk;
      s.halo_rows = ((mt * 128 + (R - 1) * WP + (S - 1)) + 7) / 8 * 8;
      if (s.halo_rows < HP * WP) s.halo_rows = (HP * WP + 7) / 8 * 8;
      s.b_stages = 3;
      if (halo_smem_bytes(s, bn, es) > kSmemLimit) continue;
      const int th = (OH + bh - 1) / bh, tw = (OW + bw - 1) / bw;
      const double useful = (double)OH * OW;
      const double done = (double)th * tw * mt * 128;
      double score = useful / done + 1e-3 * mt;  // tie-break: more rows per filter pass
      if (score > best_score) {
        best_score = score;
        best.ok = true;
        s.bh = bh; s.bw = bw; s.WP = WP; s.HP = HP; s.n_mtiles = mt; s.tiles_h = th; s.tiles_w = tw;
        best.s = s;
      }
    }
  }
  if (!best.ok) return best;
  HaloShape& s = best.s;
  s.N = N; s.H = H; s.W = W; s.C = C; s.K = K; s.R = R; s.S = S; s.pt = pt; s.pl = pl;
  s.OH = OH; s.OW = OW;
  s.nblocks = (K + bn - 1) / bn;
  best.bn = bn;
  s.items_per_nb = (long long)N * s.tiles_h * s.tiles_w;
  // cluster: multicast the filter tiles when there is enough work for whole clusters
  static const int force_cl = [] {
    const char* v = getenv("B200TF_CONV_HALO_CLUSTER");
    return v ? atoi(v) : 0;
  }();
  // pairs tile all 148 SMs; clusters of 4 fit only 33 x 4 = 132 CTAs on this part (measured)
  int cl = force_cl ? force_cl : 2;
  while (cl > 1 && s.items_per_nb < (long long)cl * 8) cl /= 2;
  if (cl != 1 && cl != 2 && cl != 4) cl = 1;
  s.cl = cl;
  s.padded_per_nb = (s.items_per_nb + cl - 1) / cl * cl;
  s.work_padded = s.padded_per_nb * s.nblocks;
  // deepest filter ring that fits
  s.b_stages = kMaxBStages;
  while (s.b_stages > 2 && halo_smem_bytes(s, bn, es) > kSmemLimit) --s.b_stages;
  best.smem = halo_smem_bytes(s, bn, es);
  best.bn = bn;
  if (best.smem > kSmemLimit) best.ok = false;
  return best;
}

template <typename TIn, int BN, bool kPair>
static int launch_halo(const HaloPlan& p, const void* input, const void* filter, void* output,
                       cudaStream_t stream) {
  using Tr = HaloTraits<TIn>;
  const HaloShape& s = p.s;
  constexpr int es = (int)sizeof(TIn);
  CUtensorMap mx, mw;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)s.C, (cuuint64_t)s.W, (cuuint64_t)s.H, (cuuint64_t)s.N};
    cuuint64_t gstr[3] = {(cuuint64_t)s.C * es, (cuuint64_t)s.W * s.C * es,
                          (cuuint64_t)s.H * s.W * s.C * es};
    cuuint32_t box[4] = {(cuuint32_t)Tr::kChunk, (cuuint32_t)s.WP, (cuuint32_t)s.HP, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = driver().cuTensorMapEncodeTiled(
        &mx, Tr::kTmaType, 4, const_cast<void*>(input), gdim, gstr, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv_halo: input tensor map failed (CUresult %d)", (int)r);
      return B200_INTERNAL;
    }
  }
  {
    // filter [R*S*C, K] row-major viewed as (128-B column chunk, row, chunk index)
    const int row_parts = (!kPair && s.cl > BN / Tr::kChunk) ? s.cl / (BN / Tr::kChunk) : 1;
    cuuint64_t gdim[3] = {(cuuint64_t)Tr::kChunk, (cuuint64_t)s.R * s.S * s.C,
                          (cuuint64_t)(s.K / Tr::kChunk)};
    cuuint64_t gstr[2] = {(cuuint64_t)s.K * es, (cuuint64_t)Tr::kChunk * es};
    cuuint32_t box[3] = {(cuuint32_t)Tr::kChunk, (cuuint32_t)(Tr::kChunk / row_parts), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = driver().cuTensorMapEncodeTiled(
        &mw, Tr::kTmaType, 3, const_cast<void*>(filter), gdim, gstr, box, estr,
        CU_TENSOR_MAP_INTERLEAVE_NONE,
        es == 4 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_last_error("conv_halo: filter tensor map failed (CUresult %d)", (int)r);
      return B200_INTERNAL;
    }
  }
  auto kern = conv_halo_kernel<TIn, TIn, BN, kPair>;
  static size_t attr_smem = 0;  // per instantiation
  if (attr_smem < p.smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kSmemLimit);
    if (e != cudaSuccess) {
      set_last_error("conv_halo: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
      cudaGetLastError();
      return B200_INTERNAL;
    }
    attr_smem = kSmemLimit;
  }
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kHaloThreads);
  cfg.dynamicSmemBytes = p.smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (s.cl > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = s.cl;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  // Persistent static schedule: the grid must not exceed what is co-resident, or the surplus
  // clusters run as a second wave.  A GPC holds a whole number of clusters, so fewer than
  // SMs / cl clusters may fit: ask the occupancy calculator (cached per cluster size).
  static int max_clusters[5] = {0, 0, 0, 0, 0};
  if (max_clusters[s.cl] == 0) {
    int n = 0;
    cfg.gridDim = dim3((unsigned)(sm_count() / s.cl * s.cl));
    cfg.attrs = attr;
    cfg.numAttrs = na;
    if (s.cl > 1 && cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0)
      max_clusters[s.cl] = n;
    else
      max_clusters[s.cl] = sm_count() / s.cl;
    cudaGetLastError();
  }
  const long long clusters_wanted = s.work_padded / s.cl;
  const int clusters = (int)std::min<long long>(clusters_wanted, max_clusters[s.cl]);
  cfg.gridDim = dim3((unsigned)(clusters * s.cl));
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  const bool prof = profile_enabled();
  if (prof) profile_gemm_launch_begin(stream);
  void* ktok = kernel_times_enabled() ? kernel_times_begin(stream) : nullptr;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mx, mw, static_cast<TIn*>(output), s);
  if (ktok) kernel_times_end(ktok, stream, reinterpret_cast<const void*>(kern));
  if (e != cudaSuccess) {
    set_last_error("conv_halo launch: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return B200_INTERNAL;
  }
  if (prof)
    profile_gemm_launch_end(stream, 2.0 * (double)s.N * s.OH * s.OW * (double)s.K * s.R * s.S * s.C);
  note_launch();
  return check_launch("conv_halo");
}

}  // namespace

bool conv_halo_supported(int dtype, const ConvHaloArgs& a) {
  if (!driver().cuTensorMapEncodeTiled) return false;
  if (dtype != B200_DT_FLOAT && dtype != B200_DT_BFLOAT16) return false;
  if ((reinterpret_cast<uintptr_t>(a.input) & 15) || (reinterpret_cast<uintptr_t>(a.filter) & 15) ||
      (reinterpret_cast<uintptr_t>(a.output) & 15))
    return false;
  if (a.R * a.S > 1024 || a.N <= 0 || a.OH <= 0 || a.OW <= 0) return false;
  // every output pixel's window must start inside the zero-filled halo the TMA box provides
  if (a.pt < 0 || a.pl < 0 || a.pt >= a.R + 128 || a.pl >= a.S + 128) return false;
  static const bool off = getenv("B200TF_CONV_NO_HALO") != nullptr;
  if (off) return false;
  return plan_halo(dtype, a.N, a.H, a.W, a.C, a.K, a.R, a.S, a.pt, a.pl, a.OH, a.OW).ok;
}

int conv_halo(int dtype, const ConvHaloArgs& a, cudaStream_t stream) {
  HaloPlan p = plan_halo(dtype, a.N, a.H, a.W, a.C, a.K, a.R, a.S, a.pt, a.pl, a.OH, a.OW);
  if (!p.ok) {
    set_last_error("conv_halo: unsupported geometry");
    return B200_UNIMPLEMENTED;
  }
  p.s.bias = a.bias;
  p.s.relu = a.relu ? 1 : 0;
  // CTA pairs (cta_group::2) whenever the cluster is a pair and the filter tile has an even number
  // of 128-byte column chunks; B200TF_CONV_HALO_NO_PAIR=1 keeps independent CTAs (comparison).
  static const bool no_pair = getenv("B200TF_CONV_HALO_NO_PAIR") != nullptr;
  const int chunk = dtype == B200_DT_FLOAT ? 32 : 64;
  const bool pair = !no_pair && p.s.cl == 2 && (p.bn / chunk) % 2 == 0;
#define HALO_CASE(T, BN_)                                                                    \
  if (p.bn == BN_)                                                                           \
    return pair ? launch_halo<T, BN_, true>(p, a.input, a.filter, a.output, stream)          \
                : launch_halo<T, BN_, false>(p, a.input, a.filter, a.output, stream)
#define HALO_CASE_SINGLE(T, BN_) \
  if (p.bn == BN_) return launch_halo<T, BN_, false>(p, a.input, a.filter, a.output, stream)
  if (dtype == B200_DT_FLOAT) {
    HALO_CASE_SINGLE(float, 32);
    HALO_CASE(float, 64);
    HALO_CASE(float, 128);
    HALO_CASE(float, 256);
  } else {
    HALO_CASE_SINGLE(__nv_bfloat16, 64);
    HALO_CASE(__nv_bfloat16, 128);
    HALO_CASE(__nv_bfloat16, 256);
  }
#undef HALO_CASE
#undef HALO_CASE_SINGLE
  set_last_error("conv_halo: no kernel for BN = %d", p.bn);
  return B200_UNIMPLEMENTED;
}

}  // namespace b200
