/* oracle.c -- CPU restatement of the reference's CPU (Eigen) kernels for the hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (simple_tensorflow_b200/, libb200tf.so,
 * libb200tf_framework.so) may include, link or call this file.  Only tests/, the smoke check in
 * __graft_entry__.py and the cpu_baseline / --impl reference legs of bench.py load it.
 *
 * Why a restatement: the reference (DengZhuangSouthRd/simple_tensorflow @ 83d7422) cannot be
 * compiled here -- no bazel, no protoc, and the arithmetic lives in Eigen (bitbucket eigen/eigen
 * commit 174e09eed96c, tensorflow/workspace.bzl:75-83) which is not vendored.  Each function
 * below restates the algorithm of the cited reference file:line in plain C with fp32
 * accumulation, and is pinned against the golden vectors of the reference's own tests
 * (tests/golden/ JSON files, extracted from tensorflow/python/kernel_tests; see
 * tests/test_oracle_golden.py).  All paths are relative to /root/reference/tensorflow/.
 *
 * Layouts: dense row-major (core/framework/tensor_types.h:25-28); images NHWC, filters HWIO.
 * Threading: OpenMP over the same unit the reference shards on (batch / rows), thread count =
 * omp default = sched_getaffinity count, as NumSchedulableCPUs (core/platform/posix/port.cc:50-55).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
/* Thread count of the following calls (bench.py pins both CPU arms to the same number). */
API void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * GetWindowedOutputSizeVerbose -- core/framework/common_shape_fns.cc:19-47.
 * padding_same: 0 = VALID, 1 = SAME.  Returns 0, or -1 for the InvalidArgument cases. */
API int oracle_windowed_output_size(int64_t input_size, int64_t filter_size, int64_t stride,
                                    int padding_same, int64_t* output_size,
                                    int64_t* padding_before, int64_t* padding_after) {
  if (stride <= 0) return -1;
  if (!padding_same) {
    *output_size = (input_size - filter_size + stride) / stride;
    *padding_before = *padding_after = 0;
  } else {
    *output_size = (input_size + stride - 1) / stride;
    int64_t needed = (*output_size - 1) * stride + filter_size - input_size;
    if (needed < 0) needed = 0;
    *padding_before = needed / 2; /* odd total: the extra cell goes after */
    *padding_after = needed - *padding_before;
  }
  if (*output_size < 0) return -1;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Blocked, register-tiled fp32 GEMM used by MatMul, BatchMatMul and the im2col convolutions.
 * C[M,N] (+)= A * B with generic element strides, fp32 accumulate (what Eigen's gebp kernel does,
 * matmul_op.h:35-40: out = in0.contract(in1, dim_pair)).  Summation order over k is ascending
 * within a K block and blocks are added in ascending order. */
/* Register-blocked inner kernel: MR rows x NR columns of C stay in registers over a K block, so
 * a row of B is loaded once per MR rows (the job of Eigen's gebp micro-kernel).  Per C element
 * the arithmetic is exactly `c += a * b` for k ascending -- separate multiply and add
 * (-ffp-contract=off), same order as the plain triple loop it replaces, so results are
 * bit-identical to it; only the speed of the CPU baseline changes. */
#if defined(ORACLE_FAST) && defined(__AVX512F__)
/* CPU-arm build on an AVX-512 host: 12 x 32 = 24 zmm accumulators + 2 B vectors + 1 broadcast use
 * 27 of the 32 registers and halve the B loads per FMA of the 6 x 32 tile (measured +10-20 %). */
enum { GEMM_MR = 12, GEMM_NR = 32 };
#else
enum { GEMM_MR = 6, GEMM_NR = 32 };
#endif
static inline void sgemm_micro(const float* A, int64_t a_rs, int64_t a_cs, const float* Bp,
                               int64_t brs, float* C, int64_t ldc, int64_t kn, int zero_first) {
  float acc[GEMM_MR][GEMM_NR];
  for (int r = 0; r < GEMM_MR; ++r)
    for (int j = 0; j < GEMM_NR; ++j) acc[r][j] = zero_first ? 0.f : C[r * ldc + j];
  for (int64_t k = 0; k < kn; ++k) {
    const float* b = Bp + k * brs;
    for (int r = 0; r < GEMM_MR; ++r) {
      const float a = A[r * a_rs + k * a_cs];
      for (int j = 0; j < GEMM_NR; ++j) acc[r][j] += a * b[j];
    }
  }
  for (int r = 0; r < GEMM_MR; ++r)
    for (int j = 0; j < GEMM_NR; ++j) C[r * ldc + j] = acc[r][j];
}

static void sgemm_strided(const float* A, int64_t a_rs, int64_t a_cs, const float* B,
                          int64_t b_rs, int64_t b_cs, float* C, int64_t ldc, int64_t M, int64_t N,
                          int64_t K, int accumulate, int parallel) {
  /* work items = (row block, column panel): 66 x 256 tiles give 62 x 4 items for the MLP's
   * 4096 x 1024 products and 16 x 4 for its 1024 x 1024 weight gradients, enough for 64 threads */
  enum { KB = 256, MB = GEMM_MR == 12 ? 72 : 66, NB = 256 };
  const int64_t mblocks = (M + MB - 1) / MB, nblocks = (N + NB - 1) / NB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1) if (parallel)
  for (int64_t mb = 0; mb < mblocks; ++mb) {
    for (int64_t nb = 0; nb < nblocks; ++nb) {
      const int64_t m0 = mb * MB, m1 = m0 + MB < M ? m0 + MB : M;
      const int64_t n0 = nb * NB, n1 = n0 + NB < N ? n0 + NB : N;
      const int64_t nw = n1 - n0;
      const int64_t nfull = nw / GEMM_NR * GEMM_NR, ntail = nw - nfull;
      /* The K x nw block of B is always packed, PANEL-major: panel p holds columns
       * [p * NR, (p + 1) * NR) as kn contiguous rows of NR floats, so the micro-kernel streams 32 KB
       * of consecutive memory that stays in L1 across the row groups of the block.  (Reading B in
       * place, rows a power-of-two stride apart, every k of a panel falls into the same few L1
       * sets: the MLP's X.W products ran 7x slower than its X.W^T ones.)  Columns past the last
       * full panel follow as kn rows of ntail floats.  Packing copies values; the arithmetic and
       * its order are unchanged (Eigen's gebp packs its rhs the same way, GeneralBlockPanelKernel.h). */
      float* packB = (float*)malloc(sizeof(float) * KB * (size_t)nw);
      /* ... and so is the MB x kn block of A (Eigen packs its lhs too): row group g holds rows
       * [g * MR, (g + 1) * MR) as kn consecutive MR-vectors, so the broadcasts of a k step share
       * one cache line whatever A's strides are (A^T products jump 4 KB per k otherwise). */
      float* packA = (float*)malloc(sizeof(float) * KB * (size_t)MB);
      for (int64_t k0 = 0; k0 < K; k0 += KB) {
        const int64_t k1 = k0 + KB < K ? k0 + KB : K;
        const int64_t kn = k1 - k0;
        float* tailB = packB + kn * nfull;
        if (b_cs == 1) { /* rows of B are contiguous: walk them */
          for (int64_t k = 0; k < kn; ++k) {
            const float* brow = B + (k0 + k) * b_rs + n0;
            for (int64_t j = 0; j < nfull; ++j)
              packB[(j / GEMM_NR) * kn * GEMM_NR + k * GEMM_NR + (j % GEMM_NR)] = brow[j];
            for (int64_t j = 0; j < ntail; ++j) tailB[k * ntail + j] = brow[nfull + j];
          }
        } else { /* B^T is stored: its columns are the contiguous direction */
          for (int64_t j = 0; j < nw; ++j) {
            const float* bcol = B + k0 * b_rs + (n0 + j) * b_cs;
            float* dst = j < nfull ? packB + (j / GEMM_NR) * kn * GEMM_NR + (j % GEMM_NR)
                                   : tailB + (j - nfull);
            const int64_t ds = j < nfull ? GEMM_NR : ntail;
            for (int64_t k = 0; k < kn; ++k) dst[k * ds] = bcol[k * b_rs];
          }
        }
        const int zero_first = k0 == 0 && !accumulate;
        const int64_t mfull = m0 + (m1 - m0) / GEMM_MR * GEMM_MR;
        for (int64_t i = m0; i < mfull; ++i) {
          const float* arow = A + i * a_rs + k0 * a_cs;
          float* dst = packA + ((i - m0) / GEMM_MR) * kn * GEMM_MR + (i - m0) % GEMM_MR;
          for (int64_t k = 0; k < kn; ++k) dst[k * GEMM_MR] = arow[k * a_cs];
        }
        for (int64_t j = 0; j < nfull; j += GEMM_NR)
          for (int64_t i = m0; i < mfull; i += GEMM_MR)
            sgemm_micro(packA + ((i - m0) / GEMM_MR) * kn * GEMM_MR, 1, GEMM_MR,
                        packB + (j / GEMM_NR) * kn * GEMM_NR, GEMM_NR, C + i * ldc + n0 + j, ldc, kn,
                        zero_first);
        /* edges: the tail columns of the full row groups, then every column of the remaining rows */
        for (int64_t i = m0; i < m1; ++i) {
          float* c = C + i * ldc + n0;
          const float* arow = A + i * a_rs + k0 * a_cs;
          if (i >= mfull) {
            for (int64_t p = 0; p < nfull; p += GEMM_NR) {
              const float* bp = packB + (p / GEMM_NR) * kn * GEMM_NR;
              if (zero_first)
                for (int64_t j = 0; j < GEMM_NR; ++j) c[p + j] = 0.f;
              for (int64_t k = 0; k < kn; ++k) {
                const float a = arow[k * a_cs];
                for (int64_t j = 0; j < GEMM_NR; ++j) c[p + j] += a * bp[k * GEMM_NR + j];
              }
            }
          }
          if (ntail > 0) {
            if (zero_first)
              for (int64_t j = nfull; j < nw; ++j) c[j] = 0.f;
            for (int64_t k = 0; k < kn; ++k) {
              const float a = arow[k * a_cs];
              for (int64_t j = 0; j < ntail; ++j) c[nfull + j] += a * tailB[k * ntail + j];
            }
          }
        }
      }
      free(packB);
      free(packA);
    }
  }
}

/* MatMulOp<CPUDevice,float>::Compute -- core/kernels/matmul_op.cc:215-256.
 * a: [m,k] or [k,m] when transpose_a; b: [k,n] or [n,k] when transpose_b; out [m,n].
 * Zero-size rules (:240-253): empty output -> nothing; k == 0 -> zero fill. */
API void oracle_matmul_f32(const float* a, const float* b, float* out, int64_t m, int64_t n,
                           int64_t k, int transpose_a, int transpose_b) {
  if (m == 0 || n == 0) return;
  if (k == 0) {
    memset(out, 0, sizeof(float) * (size_t)(m * n));
    return;
  }
  sgemm_strided(a, transpose_a ? 1 : k, transpose_a ? m : 1, b, transpose_b ? 1 : n,
                transpose_b ? k : 1, out, n, m, n, k, 0, 1);
}

/* BatchMatMul<CPUDevice,float>::Compute -- core/kernels/batch_matmul_op_impl.h:367-434 with the
 * real-type adjoint == transpose (:91-94).  x [batch,m,k] / [batch,k,m]; y [batch,k,n] / [batch,n,k]. */
API void oracle_batch_matmul_f32(const float* x, const float* y, float* out, int64_t batch,
                                 int64_t m, int64_t n, int64_t k, int adj_x, int adj_y) {
  if (batch == 0 || m == 0 || n == 0) return;
  if (k == 0) {
    memset(out, 0, sizeof(float) * (size_t)(batch * m * n));
    return;
  }
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b)
    sgemm_strided(x + b * m * k, adj_x ? 1 : k, adj_x ? m : 1, y + b * k * n, adj_y ? 1 : n,
                  adj_y ? k : 1, out + b * m * n, n, m, n, k, 0, 0);
}

/* ---------------------------------------------------------------------------------------------
 * functor::Bias -- core/kernels/bias_op.h:27-53 (op bias_op.cc:62-117, NHWC only on CPU). */
API void oracle_bias_add_f32(const float* in, const float* bias, float* out, int64_t rows,
                             int64_t channels) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < channels; ++c) out[r * channels + c] = in[r * channels + c] + bias[c];
}

/* BiasGradOp<CPUDevice,float>::Compute -- core/kernels/bias_op.cc:185-227: reshape to
 * [rows, channels], sum over axis 0 in AccumulatorType<float> = float. */
API void oracle_bias_add_grad_f32(const float* out_backprop, float* out, int64_t rows,
                                  int64_t channels) {
  for (int64_t c = 0; c < channels; ++c) out[c] = 0.f;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t c = 0; c < channels; ++c) out[c] += out_backprop[r * channels + c];
}

/* functor::Relu / functor::ReluGrad -- core/kernels/relu_op_functor.h:28-60.
 * cwiseMax(0); gradients * (features > 0): a zero activation passes no gradient. */
API void oracle_relu_f32(const float* features, float* activations, int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) activations[i] = features[i] > 0.f ? features[i] : 0.f;
}
API void oracle_relu_grad_f32(const float* gradients, const float* features, float* backprops,
                              int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) backprops[i] = gradients[i] * (features[i] > 0.f ? 1.f : 0.f);
}

/* SoftmaxEigenImpl::Compute -- core/kernels/softmax_op_functor.h:43-99.
 * shifted = logits - rowmax; softmax = exp(shifted) * (1 / sum exp(shifted));
 * log-softmax = shifted - log(sum exp(shifted)). */
API void oracle_softmax_f32(const float* logits, float* out, int64_t rows, int64_t cols, int log) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* x = logits + r * cols;
    float* y = out + r * cols;
    float mx = -FLT_MAX; /* Eigen maximum() reducer starts from lowest() */
    for (int64_t c = 0; c < cols; ++c) mx = x[c] > mx ? x[c] : mx;
    float sum = 0.f;
    for (int64_t c = 0; c < cols; ++c) sum += expf(x[c] - mx);
    if (log) {
      const float ls = logf(sum);
      for (int64_t c = 0; c < cols; ++c) y[c] = (x[c] - mx) - ls;
    } else {
      const float inv = 1.f / sum;
      for (int64_t c = 0; c < cols; ++c) y[c] = expf(x[c] - mx) * inv;
    }
  }
}

/* XentEigenImpl::Compute -- core/kernels/xent_op.h:47-113.
 * loss[r] = sum_c labels * (log(sum exp(shifted)) - shifted); backprop = exp(shifted)/sum - labels. */
API void oracle_softmax_xent_f32(const float* logits, const float* labels, float* loss,
                                 float* backprop, int64_t rows, int64_t cols) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float* x = logits + r * cols;
    const float* l = labels + r * cols;
    float* bp = backprop + r * cols;
    float mx = -FLT_MAX;
    for (int64_t c = 0; c < cols; ++c) mx = x[c] > mx ? x[c] : mx;
    float sum = 0.f;
    for (int64_t c = 0; c < cols; ++c) sum += expf(x[c] - mx);
    const float ls = logf(sum);
    float acc = 0.f;
    for (int64_t c = 0; c < cols; ++c) acc += l[c] * (ls - (x[c] - mx));
    loss[r] = acc;
    for (int64_t c = 0; c < cols; ++c) bp[c] = expf(x[c] - mx) / sum - l[c];
  }
}

/* ---------------------------------------------------------------------------------------------
 * MaxPoolingOp<CPUDevice,float>::SpatialMaxPool -- core/kernels/pooling_ops_common.h:204-238.
 * Scatter formulation: output initialised to lowest(), every input pixel is max-merged into
 * each output window it projects to; padded cells never participate. */
API void oracle_max_pool_f32(const float* in, float* out, int64_t batch, int64_t in_rows,
                             int64_t in_cols, int64_t depth, int64_t out_height,
                             int64_t out_width, int window_rows, int window_cols, int row_stride,
                             int col_stride, int pad_rows, int pad_cols) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    float* ob = out + b * out_height * out_width * depth;
    for (int64_t i = 0; i < out_height * out_width * depth; ++i) ob[i] = -FLT_MAX;
    for (int64_t h = 0; h < in_rows; ++h)
      for (int64_t w = 0; w < in_cols; ++w) {
        const int64_t hpad = h + pad_rows, wpad = w + pad_cols;
        const int64_t h_start = hpad < window_rows ? 0 : (hpad - window_rows) / row_stride + 1;
        int64_t h_end = hpad / row_stride + 1;
        if (h_end > out_height) h_end = out_height;
        const int64_t w_start = wpad < window_cols ? 0 : (wpad - window_cols) / col_stride + 1;
        int64_t w_end = wpad / col_stride + 1;
        if (w_end > out_width) w_end = out_width;
        const float* ip = in + ((b * in_rows + h) * in_cols + w) * depth;
        for (int64_t ph = h_start; ph < h_end; ++ph)
          for (int64_t pw = w_start; pw < w_end; ++pw) {
            float* op = ob + (ph * out_width + pw) * depth;
            for (int64_t d = 0; d < depth; ++d) op[d] = op[d] > ip[d] ? op[d] : ip[d];
          }
      }
  }
}

/* MaxPoolingGradOp<CPUDevice,float> via SpatialMaxPoolWithArgMaxHelper --
 * core/kernels/maxpooling_op.cc:52-188.  Forward is recomputed with an argmax per output cell:
 * an input replaces the incumbent when (output < input || argmax == -1), so the FIRST maximum
 * in row-major input scan order wins; then in_backprop[argmax] += grad. */
API void oracle_max_pool_grad_f32(const float* orig_in, const float* grad, float* in_backprop,
                                  int64_t batch, int64_t in_rows, int64_t in_cols, int64_t depth,
                                  int64_t out_height, int64_t out_width, int window_rows,
                                  int window_cols, int row_stride, int col_stride, int pad_rows,
                                  int pad_cols) {
  const int64_t out_img = out_height * out_width * depth;
  const int64_t in_img = in_rows * in_cols * depth;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    float* best = (float*)malloc(sizeof(float) * (size_t)out_img);
    int64_t* arg = (int64_t*)malloc(sizeof(int64_t) * (size_t)out_img);
    for (int64_t i = 0; i < out_img; ++i) {
      best[i] = -FLT_MAX;
      arg[i] = -1;
    }
    const float* ib = orig_in + b * in_img;
    for (int64_t h = 0; h < in_rows; ++h)
      for (int64_t w = 0; w < in_cols; ++w) {
        const int64_t hpad = h + pad_rows, wpad = w + pad_cols;
        const int64_t h_start = hpad < window_rows ? 0 : (hpad - window_rows) / row_stride + 1;
        int64_t h_end = hpad / row_stride + 1;
        if (h_end > out_height) h_end = out_height;
        const int64_t w_start = wpad < window_cols ? 0 : (wpad - window_cols) / col_stride + 1;
        int64_t w_end = wpad / col_stride + 1;
        if (w_end > out_width) w_end = out_width;
        const int64_t in_index = h * in_cols + w;
        for (int64_t ph = h_start; ph < h_end; ++ph)
          for (int64_t pw = w_start; pw < w_end; ++pw) {
            const int64_t o = (ph * out_width + pw) * depth;
            for (int64_t d = 0; d < depth; ++d) {
              const float v = ib[in_index * depth + d];
              if (best[o + d] < v || arg[o + d] == -1) {
                best[o + d] = v;
                arg[o + d] = in_index * depth + d;
              }
            }
          }
      }
    float* gb = in_backprop + b * in_img;
    for (int64_t i = 0; i < in_img; ++i) gb[i] = 0.f;
    const float* g = grad + b * out_img;
    for (int64_t i = 0; i < out_img; ++i) gb[arg[i]] += g[i];
    free(best);
    free(arg);
  }
}

/* ---------------------------------------------------------------------------------------------
 * Cast float <-> bfloat16 -- core/framework/bfloat16.cc:20-50 (scalar_cast_op, cast_op.h:119-141):
 * float -> bfloat16 keeps the upper 16 bits (TRUNCATION, no rounding); bfloat16 -> float shifts. */
API void oracle_cast_f32_to_bf16(const float* in, uint16_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, &in[i], 4);
    out[i] = (uint16_t)(u >> 16);
  }
}
API void oracle_cast_bf16_to_f32(const uint16_t* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u = ((uint32_t)in[i]) << 16;
    memcpy(&out[i], &u, 4);
  }
}
/* CastOp for the numeric pairs the hot path's graphs use: C++ static_cast (cast_op.h:93-117). */
API void oracle_cast_f32_to_i32(const float* in, int32_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)in[i];
}
API void oracle_cast_i32_to_f32(const int32_t* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (float)in[i];
}
API void oracle_cast_i64_to_f32(const int64_t* in, float* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (float)in[i];
}
API void oracle_cast_f32_to_i64(const float* in, int64_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)in[i];
}
API void oracle_cast_i32_to_i64(const int32_t* in, int64_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (int64_t)in[i];
}
API void oracle_cast_i64_to_i32(const int64_t* in, int32_t* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)in[i];
}

/* ArgOp<CPUDevice,T,ArgMax> -- core/kernels/argmax_op.cc:44-98, argmax_op.h:29-42:
 * input viewed as [outer, axis, inner]; Eigen's tuple reducer keeps the LOWEST index on ties;
 * output int64. */
API void oracle_argmax_f32(const float* in, int64_t* out, int64_t outer, int64_t axis_size,
                           int64_t inner) {
  for (int64_t o = 0; o < outer; ++o)
    for (int64_t i = 0; i < inner; ++i) {
      int64_t best = 0;
      float bv = in[(o * axis_size) * inner + i];
      for (int64_t a = 1; a < axis_size; ++a) {
        const float v = in[(o * axis_size + a) * inner + i];
        if (v > bv) {
          bv = v;
          best = a;
        }
      }
      out[o * inner + i] = best;
    }
}
API void oracle_argmax_i32(const int32_t* in, int64_t* out, int64_t outer, int64_t axis_size,
                           int64_t inner) {
  for (int64_t o = 0; o < outer; ++o)
    for (int64_t i = 0; i < inner; ++i) {
      int64_t best = 0;
      int32_t bv = in[(o * axis_size) * inner + i];
      for (int64_t a = 1; a < axis_size; ++a) {
        const int32_t v = in[(o * axis_size + a) * inner + i];
        if (v > bv) {
          bv = v;
          best = a;
        }
      }
      out[o * inner + i] = best;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Convolutions.  Geometry struct mirrors include/b200_ops.h b200_conv2d_geometry. */
typedef struct {
  int64_t batch, in_h, in_w, in_c;
  int64_t filter_h, filter_w, out_c;
  int64_t out_h, out_w;
  int32_t stride_h, stride_w;
  int32_t pad_top, pad_left;
} conv_geom;

/* Im2col -- core/kernels/conv_grad_filter_ops.cc:55-82: patches in (out_h*out_w, fh, fw, depth)
 * order, zero for padded cells. */
static void im2col_image(const float* img, const conv_geom* g, float* col) {
  const int64_t d = g->in_c;
  for (int64_t oh = 0; oh < g->out_h; ++oh)
    for (int64_t ow = 0; ow < g->out_w; ++ow) {
      const int64_t h0 = oh * g->stride_h - g->pad_top, w0 = ow * g->stride_w - g->pad_left;
      for (int64_t ih = h0; ih < h0 + g->filter_h; ++ih)
        for (int64_t iw = w0; iw < w0 + g->filter_w; ++iw) {
          if (ih >= 0 && ih < g->in_h && iw >= 0 && iw < g->in_w)
            memcpy(col, img + (ih * g->in_w + iw) * d, sizeof(float) * (size_t)d);
          else
            memset(col, 0, sizeof(float) * (size_t)d);
          col += d;
        }
    }
}

/* Col2im -- core/kernels/conv_grad_input_ops.cc:57-87: scatter-add patches back into a
 * zero-initialised image. */
static void col2im_image(const float* col, const conv_geom* g, float* img) {
  const int64_t d = g->in_c;
  for (int64_t oh = 0; oh < g->out_h; ++oh)
    for (int64_t ow = 0; ow < g->out_w; ++ow) {
      const int64_t h0 = oh * g->stride_h - g->pad_top, w0 = ow * g->stride_w - g->pad_left;
      for (int64_t ih = h0; ih < h0 + g->filter_h; ++ih)
        for (int64_t iw = w0; iw < w0 + g->filter_w; ++iw) {
          if (ih >= 0 && ih < g->in_h && iw >= 0 && iw < g->in_w) {
            float* p = img + (ih * g->in_w + iw) * d;
            for (int64_t i = 0; i < d; ++i) p[i] += col[i];
          }
          col += d;
        }
    }
}

/* Conv2DOp<CPUDevice,float> -> LaunchGeneric -> Eigen::SpatialConvolution --
 * core/kernels/conv_ops.cc:59-110, core/kernels/eigen_spatial_convolutions.h:1050-1067:
 * out[N*OH*OW, K] = extract_image_patches(input)[N*OH*OW, R*S*C] . filter[R*S*C, K]. */
API void oracle_conv2d_f32(const float* input, const float* filter, float* output,
                           const conv_geom* g) {
  const int64_t patch = g->filter_h * g->filter_w * g->in_c;
  const int64_t opix = g->out_h * g->out_w;
  if (g->batch == 0 || opix == 0 || g->out_c == 0) return;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < g->batch; ++b) {
    float* col = (float*)malloc(sizeof(float) * (size_t)(opix * patch));
    im2col_image(input + b * g->in_h * g->in_w * g->in_c, g, col);
    sgemm_strided(col, patch, 1, filter, g->out_c, 1, output + b * opix * g->out_c, g->out_c,
                  opix, g->out_c, patch, 0, 0);
    free(col);
  }
}

/* Conv2DCustomBackpropInputOp<CPUDevice,float>::Compute --
 * core/kernels/conv_grad_input_ops.cc:266-501: per image
 * col[OH*OW, R*S*C] = out_backprop[OH*OW, K] . filter[R*S*C, K]^T, then Col2im into zeroed dX. */
API void oracle_conv2d_backprop_input_f32(const float* filter, const float* out_backprop,
                                          float* in_backprop, const conv_geom* g) {
  const int64_t patch = g->filter_h * g->filter_w * g->in_c;
  const int64_t opix = g->out_h * g->out_w;
  const int64_t in_img = g->in_h * g->in_w * g->in_c;
  if (g->batch == 0 || in_img == 0) return;
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t b = 0; b < g->batch; ++b) {
    float* img = in_backprop + b * in_img;
    memset(img, 0, sizeof(float) * (size_t)in_img);
    if (opix == 0 || g->out_c == 0) continue;
    float* col = (float*)malloc(sizeof(float) * (size_t)(opix * patch));
    /* B = filter^T: element (k, j) = filter[j * out_c + k] */
    sgemm_strided(out_backprop + b * opix * g->out_c, g->out_c, 1, filter, 1, g->out_c, col, patch,
                  opix, patch, g->out_c, 0, 0);
    col2im_image(col, g, img);
    free(col);
  }
}

/* Conv2DCustomBackpropFilterOp<CPUDevice,float>::Compute --
 * core/kernels/conv_grad_filter_ops.cc:155-329: dW[R*S*C, K] += im2col(X)^T . dY, accumulated
 * image by image in ascending batch order (the reference accumulates shard by shard, :311-317). */
API void oracle_conv2d_backprop_filter_f32(const float* input, const float* out_backprop,
                                           float* filter_backprop, const conv_geom* g) {
  const int64_t patch = g->filter_h * g->filter_w * g->in_c;
  const int64_t opix = g->out_h * g->out_w;
  if (patch * g->out_c == 0) return;
  memset(filter_backprop, 0, sizeof(float) * (size_t)(patch * g->out_c));
  if (g->batch == 0 || opix == 0) return;
  float* col = (float*)malloc(sizeof(float) * (size_t)(opix * patch));
  for (int64_t b = 0; b < g->batch; ++b) {
    im2col_image(input + b * g->in_h * g->in_w * g->in_c, g, col);
    /* A = col^T: element (i, k) = col[k * patch + i] */
    sgemm_strided(col, 1, patch, out_backprop + b * opix * g->out_c, g->out_c, 1, filter_backprop,
                  g->out_c, patch, g->out_c, opix, 1, 1);
  }
  free(col);
}

/* ---------------------------------------------------------------------------------------------
 * Graph glue.
 * ApplyGradientDescent -- core/kernels/training_ops.cc:410-412: var -= alpha * delta. */
API void oracle_apply_gradient_descent_f32(float* var, float alpha, const float* delta,
                                           int64_t n) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) var[i] -= alpha * delta[i];
}
