/*
 * ofspmm_oracle.c — CPU restatement of the SpMM / A^T·dY / SDDMM path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product path (of-spmm_b200/) never links, imports or calls
 * it; it fails loudly when the CUDA library is missing instead.
 *
 * PARITY UNPINNED: the mounted reference (/root/reference, stock OneFlow v0.8.1-dev) contains no
 * SpMM / SDDMM / CSR operator, no golden vectors and no test for one, and
 * OneFlow cannot be built offline.  This oracle therefore restates the
 * *mathematical definition* of the op in the idiom of OneFlow's CPU kernels and is triangulated
 * against two independent implementations (scipy.sparse and torch.sparse_csr) by
 * tests/golden/make_golden.py and tests/test_oracle.py.
 *
 * Reference idiom followed (paths relative to /root/reference):
 *   - CPU kernels are plain sequential loops in storage order that add one dense row into
 *     another with std::transform(..., std::plus<T>()):
 *       oneflow/user/kernels/unsorted_segment_sum_kernel_util.cpp:28-45
 *       oneflow/user/kernels/gather_kernel_util.cpp:72-95
 *   - indices are CHECK_GE(idx, 0)'d and out-of-range ones silently skipped:
 *       oneflow/user/kernels/unsorted_segment_sum_kernel_util.cpp:35-39
 *   - half/bf16 data accumulate in an fp32 buffer and are cast once at the end:
 *       oneflow/user/kernels/unsorted_segment_sum_kernel.cpp:145-189
 *       oneflow/core/ep/cuda/primitive/broadcast_matmul.cpp:74-82 (fp32 compute type)
 *   - multi-threaded CPU work = equal-count contiguous blocks (BalancedSplitter) over a pool:
 *       oneflow/core/common/balanced_splitter.cpp:20-39
 *       oneflow/core/thread/thread_manager.h:53-73
 *
 * Notation: A is M×K CSR (crow[M+1], col[nnz], val[nnz]); B is K×N row-major; C is M×N row-major.
 * Index arrays are int32 or int64 (idx64 flag), mirroring INDEX_DATA_TYPE_SEQ
 * (oneflow/core/common/data_type_seq.h:50-52).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define IDX(arr, i, idx64) ((idx64) ? ((const int64_t*)(arr))[(i)] : (int64_t)((const int32_t*)(arr))[(i)])

/* ---------------------------------------------------------------- bf16 helpers */

/* bf16 -> fp32 is exact (bf16 is the top 16 bits of an IEEE fp32). */
void oracle_bf16_to_f32(const uint16_t* src, float* dst, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u = ((uint32_t)src[i]) << 16;
    memcpy(&dst[i], &u, 4);
  }
}

/* fp32 -> bf16, round-to-nearest-even, NaN preserved (same as __float2bfloat16_rn). */
void oracle_f32_to_bf16(const float* src, uint16_t* dst, int64_t n) {
  for (int64_t i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, &src[i], 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) {
      dst[i] = (uint16_t)((u >> 16) | 0x0040u);
    } else {
      uint32_t lsb = (u >> 16) & 1u;
      u += 0x7fffu + lsb;
      dst[i] = (uint16_t)(u >> 16);
    }
  }
}

/* ---------------------------------------------------------------- forward  C = A·B */

/* oracle-A ("reference-style"): fp32 accumulate, rows in order, non-zeros in storage order.
 * C[i,:] = sum_p val[p] * B[col[p],:]; out-of-range columns are skipped like the reference's
 * segment-sum does (unsorted_segment_sum_kernel_util.cpp:37). */
static void spmm_rows_f32(const void* crow, const void* col, const float* val, const float* B,
                          int64_t K, int64_t N, float* C, int idx64, int64_t r0, int64_t r1) {
  for (int64_t i = r0; i < r1; ++i) {
    float* to = C + i * N;
    for (int64_t j = 0; j < N; ++j) to[j] = 0.0f;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const float a = val[p];
      const float* from = B + c * N;
      for (int64_t j = 0; j < N; ++j) to[j] += a * from[j];
    }
  }
}

void oracle_spmm_f32(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                     const float* val, const float* B, float* C, int idx64) {
  spmm_rows_f32(crow, col, val, B, K, N, C, idx64, 0, M);
}

/* oracle-B (ground truth): same loop, fp64 accumulate and fp64 output. */
void oracle_spmm_f64(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                     const float* val, const float* B, double* C, int idx64) {
  for (int64_t i = 0; i < M; ++i) {
    double* to = C + i * N;
    for (int64_t j = 0; j < N; ++j) to[j] = 0.0;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const double a = (double)val[p];
      const float* from = B + c * N;
      for (int64_t j = 0; j < N; ++j) to[j] += a * (double)from[j];
    }
  }
}

/* Largest |term| per output element: amax[i,j] = max_p |val[p]·B[col[p],j]|.  Feeds the
 * row-length-scaled atol of SURVEY.md §8c: atol_ij = 2^-23 · len_i · amax[i,j]. */
void oracle_spmm_absmax(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                        const float* val, const float* B, float* amax, int idx64) {
  for (int64_t i = 0; i < M; ++i) {
    float* to = amax + i * N;
    for (int64_t j = 0; j < N; ++j) to[j] = 0.0f;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const float a = val[p];
      const float* from = B + c * N;
      for (int64_t j = 0; j < N; ++j) {
        const float t = fabsf(a * from[j]);
        if (t > to[j]) to[j] = t;
      }
    }
  }
}

/* Equal-row-count multi-thread version of oracle-A: the stand-in for a OneFlow CPU kernel under
 * MultiThreadLoop + BalancedSplitter (thread_manager.h:53-73, balanced_splitter.cpp:20-39). */
typedef struct {
  const void *crow, *col;
  const float *val, *B;
  float* C;
  int64_t K, N, r0, r1;
  int idx64;
} spmm_job_t;

static void* spmm_job_main(void* arg) {
  spmm_job_t* j = (spmm_job_t*)arg;
  spmm_rows_f32(j->crow, j->col, j->val, j->B, j->K, j->N, j->C, j->idx64, j->r0, j->r1);
  return NULL;
}

/* BalancedSplitter::At(idx) restated (balanced_splitter.cpp:26-39). */
void oracle_balanced_split(int64_t total, int64_t parts, int64_t idx, int64_t* begin, int64_t* end) {
  const int64_t base = total / parts, rem = total % parts;
  if (idx < rem) {
    *begin = (base + 1) * idx;
    *end = *begin + base + 1;
  } else {
    *begin = (base + 1) * rem + base * (idx - rem);
    *end = *begin + base;
  }
}

int oracle_spmm_f32_mt(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                       const float* val, const float* B, float* C, int idx64, int threads) {
  if (threads < 1) threads = 1;
  if ((int64_t)threads > M) threads = (int)(M > 0 ? M : 1);
  if (threads == 1) {
    spmm_rows_f32(crow, col, val, B, K, N, C, idx64, 0, M);
    return 1;
  }
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  spmm_job_t* jobs = (spmm_job_t*)malloc(sizeof(spmm_job_t) * threads);
  for (int t = 0; t < threads; ++t) {
    spmm_job_t j = {crow, col, val, B, C, K, N, 0, 0, idx64};
    oracle_balanced_split(M, threads, t, &j.r0, &j.r1);
    jobs[t] = j;
    pthread_create(&tid[t], NULL, spmm_job_main, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
  free(tid);
  free(jobs);
  return threads;
}

/* ---------------------------------------------------------------- backward  dB = A^T·dY */

/* The literal scatter loop of SURVEY.md §8a4, in row / storage order (the CPU analogue of
 * embedding_grad / unsorted_segment_sum: memset then add rows,
 * oneflow/user/kernels/unsorted_segment_sum_kernel.cpp:101-117). */
void oracle_spmm_t_f32(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                       const float* val, const float* dY, float* dB, int idx64) {
  memset(dB, 0, sizeof(float) * (size_t)(K * N));
  for (int64_t i = 0; i < M; ++i) {
    const float* from = dY + i * N;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const float a = val[p];
      float* to = dB + c * N;
      for (int64_t j = 0; j < N; ++j) to[j] += a * from[j];
    }
  }
}

void oracle_spmm_t_f64(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                       const float* val, const float* dY, double* dB, int idx64) {
  memset(dB, 0, sizeof(double) * (size_t)(K * N));
  for (int64_t i = 0; i < M; ++i) {
    const float* from = dY + i * N;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const double a = (double)val[p];
      double* to = dB + c * N;
      for (int64_t j = 0; j < N; ++j) to[j] += a * (double)from[j];
    }
  }
}

/* Per-column entry count and largest |term| for the A^T·dY tolerance (the "row length" of the
 * transposed product is the column's in-degree). */
void oracle_spmm_t_absmax(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                          const float* val, const float* dY, float* amax, int64_t* colcnt,
                          int idx64) {
  memset(amax, 0, sizeof(float) * (size_t)(K * N));
  memset(colcnt, 0, sizeof(int64_t) * (size_t)K);
  for (int64_t i = 0; i < M; ++i) {
    const float* from = dY + i * N;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      if (c < 0 || c >= K) continue;
      const float a = val[p];
      float* to = amax + c * N;
      colcnt[c] += 1;
      for (int64_t j = 0; j < N; ++j) {
        const float t = fabsf(a * from[j]);
        if (t > to[j]) to[j] = t;
      }
    }
  }
}

/* Multi-thread A^T·dY for the timed CPU baseline: rows split in equal-count blocks
 * (BalancedSplitter), each thread scatters into a private K×N buffer, then the buffers are summed
 * in thread order (column blocks in parallel).  Deterministic for a fixed thread count. */
typedef struct {
  const void *crow, *col;
  const float *val, *dY;
  float* priv;   /* this thread's K*N accumulator (phase 1) */
  float** all;   /* all accumulators (phase 2) */
  float* dB;
  int64_t K, N, r0, r1, c0, c1;
  int idx64, threads;
} spmmt_job_t;

static void* spmmt_scatter_main(void* arg) {
  spmmt_job_t* j = (spmmt_job_t*)arg;
  memset(j->priv, 0, sizeof(float) * (size_t)(j->K * j->N));
  for (int64_t i = j->r0; i < j->r1; ++i) {
    const float* from = j->dY + i * j->N;
    const int64_t pe = IDX(j->crow, i + 1, j->idx64);
    for (int64_t p = IDX(j->crow, i, j->idx64); p < pe; ++p) {
      const int64_t c = IDX(j->col, p, j->idx64);
      if (c < 0 || c >= j->K) continue;
      const float a = j->val[p];
      float* to = j->priv + c * j->N;
      for (int64_t q = 0; q < j->N; ++q) to[q] += a * from[q];
    }
  }
  return NULL;
}

static void* spmmt_reduce_main(void* arg) {
  spmmt_job_t* j = (spmmt_job_t*)arg;
  for (int64_t e = j->c0 * j->N; e < j->c1 * j->N; ++e) {
    float s = j->all[0][e];
    for (int t = 1; t < j->threads; ++t) s += j->all[t][e];
    j->dB[e] = s;
  }
  return NULL;
}

int oracle_spmm_t_f32_mt(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                         const float* val, const float* dY, float* dB, int idx64, int threads) {
  if (threads < 1) threads = 1;
  if ((int64_t)threads > M) threads = (int)(M > 0 ? M : 1);
  if (threads == 1) {
    oracle_spmm_t_f32(M, K, N, crow, col, val, dY, dB, idx64);
    return 1;
  }
  pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  spmmt_job_t* jobs = (spmmt_job_t*)malloc(sizeof(spmmt_job_t) * threads);
  float** all = (float**)malloc(sizeof(float*) * threads);
  for (int t = 0; t < threads; ++t) all[t] = (float*)malloc(sizeof(float) * (size_t)(K * N > 0 ? K * N : 1));
  for (int t = 0; t < threads; ++t) {
    spmmt_job_t j = {crow, col, val, dY, all[t], all, dB, K, N, 0, 0, 0, 0, idx64, threads};
    oracle_balanced_split(M, threads, t, &j.r0, &j.r1);
    oracle_balanced_split(K, threads, t, &j.c0, &j.c1);
    jobs[t] = j;
    pthread_create(&tid[t], NULL, spmmt_scatter_main, &jobs[t]);
  }
  for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
  for (int t = 0; t < threads; ++t) pthread_create(&tid[t], NULL, spmmt_reduce_main, &jobs[t]);
  for (int t = 0; t < threads; ++t) pthread_join(tid[t], NULL);
  for (int t = 0; t < threads; ++t) free(all[t]);
  free(all);
  free(tid);
  free(jobs);
  return threads;
}

/* ---------------------------------------------------------------- SDDMM value gradient */

/* dval[p] = <dY[i,:], B[col[p],:]>, sequential dot product in column order. */
void oracle_sddmm_f32(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                      const float* dY, const float* B, float* dval, int idx64) {
  for (int64_t i = 0; i < M; ++i) {
    const float* y = dY + i * N;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      float acc = 0.0f;
      if (c >= 0 && c < K) {
        const float* b = B + c * N;
        for (int64_t j = 0; j < N; ++j) acc += y[j] * b[j];
      }
      dval[p] = acc;
    }
  }
}

/* fp64 ground truth plus sum_j |dY[i,j]·B[c,j]| (abs-sum) for a dot-product error bound. */
void oracle_sddmm_f64(int64_t M, int64_t K, int64_t N, const void* crow, const void* col,
                      const float* dY, const float* B, double* dval, double* abssum, int idx64) {
  for (int64_t i = 0; i < M; ++i) {
    const float* y = dY + i * N;
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      double acc = 0.0, aabs = 0.0;
      if (c >= 0 && c < K) {
        const float* b = B + c * N;
        for (int64_t j = 0; j < N; ++j) {
          const double t = (double)y[j] * (double)b[j];
          acc += t;
          aabs += fabs(t);
        }
      }
      dval[p] = acc;
      if (abssum) abssum[p] = aabs;
    }
  }
}

/* ---------------------------------------------------------------- partitioner + histogram */

/* Merge-path split of the (row-end, non-zero) merge list.
 *
 * List A = row-end offsets crow[1..M] (M items), list B = the naturals 0..nnz-1 (nnz items).
 * A row-end item is consumed as soon as every non-zero of its row has been.  Diagonal d splits
 * the merged list after d items; (row, nz) with row + nz = d is found by binary search.
 * Worker k of P owns diagonals [k·ipw, (k+1)·ipw) with ipw = ceil((M+nnz)/P), clamped to M+nnz.
 * Outputs P+1 split points.  This is the host reference the device partitioner must match
 * bit-for-bit.  (Contrast: the reference's only splitter is equal-count,
 * oneflow/core/common/balanced_splitter.cpp:20-39.) */
static void merge_path_search(const void* crow, int idx64, int64_t M, int64_t nnz, int64_t d,
                              int64_t* row, int64_t* nz) {
  int64_t lo = d > nnz ? d - nnz : 0;
  int64_t hi = d < M ? d : M;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    /* row-end of row `mid` (= crow[mid+1]) precedes non-zero number d-mid-1 ? */
    if (IDX(crow, mid + 1, idx64) <= d - mid - 1) lo = mid + 1; else hi = mid;
  }
  *row = lo;
  *nz = d - lo;
}

void oracle_merge_path_partition(const void* crow, int idx64, int64_t M, int64_t nnz, int64_t P,
                                 int64_t* out_row, int64_t* out_nz) {
  const int64_t total = M + nnz;
  const int64_t ipw = P > 0 ? (total + P - 1) / P : total;
  for (int64_t k = 0; k <= P; ++k) {
    int64_t d = k * ipw;
    if (d > total) d = total;
    merge_path_search(crow, idx64, M, nnz, d, &out_row[k], &out_nz[k]);
  }
}

/* Whole-row, nnz-balanced row blocks for P devices: block k = rows
 * [bounds[k], bounds[k+1]) where bounds[k] = the merge-path split row of diagonal k·ipw, i.e. a
 * row cut by a diagonal goes to the block in which it ends.  bounds[0]=0, bounds[P]=M. */
void oracle_row_blocks(const void* crow, int idx64, int64_t M, int64_t nnz, int64_t P,
                       int64_t* bounds) {
  int64_t* nz = (int64_t*)malloc(sizeof(int64_t) * (size_t)(P + 1));
  oracle_merge_path_partition(crow, idx64, M, nnz, P, bounds, nz);
  bounds[0] = 0;
  bounds[P] = M;
  free(nz);
}

/* Row-length histogram in log2 buckets: hist[0] counts empty rows, hist[b] (1<=b<=31) counts rows
 * with 2^(b-1) <= len < 2^b.  32 counters. */
void oracle_row_hist(const void* crow, int idx64, int64_t M, int64_t* hist32) {
  for (int b = 0; b < 32; ++b) hist32[b] = 0;
  for (int64_t i = 0; i < M; ++i) {
    int64_t len = IDX(crow, i + 1, idx64) - IDX(crow, i, idx64);
    int b = 0;
    while (len > 0) { ++b; len >>= 1; }
    if (b > 31) b = 31;
    hist32[b] += 1;
  }
}

/* CSR -> CSC (= CSR of A^T) with entries of each column in ascending row order (stable counting
 * sort).  Host reference for the device transpose used by the A^T·dY plan. */
void oracle_csr_transpose(int64_t M, int64_t K, const void* crow, const void* col, const float* val,
                          int idx64, int64_t* t_crow, int64_t* t_col, float* t_val,
                          int64_t* t_perm) {
  /* Entries whose column lies outside [0, K) are skipped by every product (segment-sum idiom,
   * oneflow/user/kernels/unsorted_segment_sum_kernel_util.cpp:35-39).  In the transpose they keep
   * their source order behind the last row of A^T, marked with column -1 and value 0, and
   * t_crow[K] stays nnz so that t_crow[rows] == nnz holds for the transposed CSR too. */
  const int64_t nnz = IDX(crow, M, idx64);
  for (int64_t c = 0; c <= K; ++c) t_crow[c] = 0;
  for (int64_t p = 0; p < nnz; ++p) {
    const int64_t c = IDX(col, p, idx64);
    if (c >= 0 && c < K) t_crow[c + 1] += 1;
  }
  for (int64_t c = 0; c < K; ++c) t_crow[c + 1] += t_crow[c];
  int64_t* cursor = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K > 0 ? K : 1));
  for (int64_t c = 0; c < K; ++c) cursor[c] = t_crow[c];
  int64_t tail = t_crow[K];
  for (int64_t i = 0; i < M; ++i) {
    const int64_t pe = IDX(crow, i + 1, idx64);
    for (int64_t p = IDX(crow, i, idx64); p < pe; ++p) {
      const int64_t c = IDX(col, p, idx64);
      const int ok = c >= 0 && c < K;
      const int64_t q = ok ? cursor[c]++ : tail++;
      t_col[q] = ok ? i : -1;
      if (t_val) t_val[q] = ok ? val[p] : 0.0f;
      if (t_perm) t_perm[q] = p;
    }
  }
  if (K > 0) t_crow[K] = nnz;
  free(cursor);
}
