// spmm_kernels.cuh — merge-path CSR·dense kernels for sm_100a (forward SpMM; also run on the
// transposed CSR for A^T·dY).
//
// Scheduling: the merged list of (row-end, non-zero) items is cut into tasks
// of ITEMS consecutive items; `part[k]` = (first row, first non-zero) of task k comes from the
// device merge-path partitioner.  One warp executes one task at a time (persistent grid, tasks
// interleaved over warps so that concurrently active rows are neighbours → B-row gathers share
// L2 lines).  A task therefore never holds more than ITEMS non-zeros no matter how skewed the row
// lengths are: hub rows of a power-law graph are spread over many warps and stitched together by
// the deterministic segmented fix-up pass below; short rows are packed many-to-a-task.
//
// Data movement per task: the task's col / val / crow slices are staged into shared memory with
// three 1-D TMA bulk copies (cp.async.bulk + mbarrier complete_tx, L2 evict-first since the CSR
// stream is read exactly once); every non-zero then costs one LDS broadcast of (col, val) and one
// coalesced 16-byte-per-lane gather of the dense row (float4 or 8 x bf16), fp32 FMA accumulate.
// LPR lanes cover one dense row (vector-per-row when LPR < 32: 32/LPR non-zeros in flight per
// warp instruction, combined with __shfl_xor at the row end; warp-per-row when LPR == 32).
//
// Output: rows that start and end inside a task are streamed to C once (st.global.cs).  The
// trailing partial row of a task goes to carry[k] (fp32); a row that ends in task k but started
// earlier goes to C (fp32 output) or head[k] (bf16 output, to round only once); fixup_kernel adds
// carries in ascending task order — no atomics, bitwise reproducible.
#pragma once
#include "common.cuh"

namespace ofspmm {

struct FwdParams {
  const void* crow;
  const void* col;
  const void* val;
  const void* B;
  void* C;
  const int2* part;  // P+1 split points (row, nz)
  float* carry;      // P x n fp32
  float* head;       // P x n fp32 (only for non-fp32 outputs)
  long long cols;    // K
  int rows;          // M
  int nnz;
  int n;             // dense width
  int P;             // number of tasks
};

template <typename IdxT, typename ValT, int ITEMS>
struct alignas(16) TaskStage {
  static constexpr int kPadI = 2 * (16 / sizeof(IdxT));
  static constexpr int kPadV = 2 * (16 / sizeof(ValT));
  IdxT col[ITEMS + kPadI];
  IdxT crow[ITEMS + 2 * kPadI];
  ValT val[ITEMS + kPadV];
};

// Where element e of a `count`-element global slice starting at `src` lands in the staging array
// (dst[pre + e]) and which part of it is a 16-byte aligned body a TMA bulk copy can move.
template <typename T>
__device__ __forceinline__ void seg_plan(const T* src, int count, int& pre, int& head, int& body) {
  constexpr int EPV = 16 / sizeof(T);
  pre = static_cast<int>((reinterpret_cast<uintptr_t>(src) & 15) / sizeof(T));
  head = (EPV - pre) & (EPV - 1);
  if (head > count) head = count;
  body = ((count - head) / EPV) * EPV;
}

// Ragged ends (< 16 bytes each side) are moved by lanes with ordinary loads.
template <typename T>
__device__ __forceinline__ void seg_copy_edges(T* dst, const T* src, int count, int pre, int head,
                                               int body, int lane) {
  const int edge = count - body;  // head + tail elements, at most 2*(EPV-1)
  if (lane < edge) {
    const int e = lane < head ? lane : body + lane;
    dst[pre + e] = src[e];
  }
}

template <typename DT, typename ValT, typename IdxT, int VEC, int LPR, int CH, int ITEMS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spmm_merge_kernel(const FwdParams p) {
  constexpr int G = 32 / LPR;                       // non-zeros processed per warp step
  constexpr int U = (CH * VEC >= 16) ? 2 : 4;       // unroll: independent gathers in flight
  constexpr bool kF32Out = sizeof(DT) == 4;
  using Stage = TaskStage<IdxT, ValT, ITEMS>;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Stage* stages = reinterpret_cast<Stage*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;  // which of the G concurrent non-zeros
  const int lig = lane % LPR;  // lane in group -> column chunk
  Stage& st = stages[warp];
  uint64_t* bar = &bars[warp];

  if (lane == 0) mbar_init(bar, 1);
  fence_mbar_init();
  __syncwarp();

  const uint64_t pol_stream = l2_policy_evict_first();
  const IdxT* __restrict__ crow = static_cast<const IdxT*>(p.crow);
  const IdxT* __restrict__ col = static_cast<const IdxT*>(p.col);
  const ValT* __restrict__ val = static_cast<const ValT*>(p.val);
  const int n = p.n;

  // this lane's columns: panel base + chunk ch * LPR*VEC + lig*VEC
  const int col0 = blockIdx.y * (LPR * VEC * CH) + lig * VEC;
  unsigned chmask = 0;
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
    if (col0 + ch * LPR * VEC < n) chmask |= 1u << ch;
  const DT* __restrict__ Bl = static_cast<const DT*>(p.B) + col0;
  DT* __restrict__ Cl = static_cast<DT*>(p.C) + col0;
  float* __restrict__ carry = p.carry + col0;
  float* __restrict__ headbuf = p.head + col0;

  const int total_warps = gridDim.x * WARPS;
  uint32_t phase = 0;

  for (int k = blockIdx.x * WARPS + warp; k < p.P; k += total_warps) {
    const int2 ps = __ldg(&p.part[k]);
    const int2 pe = __ldg(&p.part[k + 1]);
    const int rs = ps.x, ns = ps.y, re = pe.x, ne = pe.y;
    const int cnt_nz = ne - ns;
    const int cnt_row = re - rs + 1;  // crow[rs .. re]

    // ---- stage the task's CSR slices: TMA bulk copies for the aligned bodies
    int pre_c, head_c, body_c, pre_v, head_v, body_v, pre_r, head_r, body_r;
    seg_plan(col + ns, cnt_nz, pre_c, head_c, body_c);
    seg_plan(val + ns, cnt_nz, pre_v, head_v, body_v);
    seg_plan(crow + rs, cnt_row, pre_r, head_r, body_r);
    const uint32_t tx = static_cast<uint32_t>(body_c * sizeof(IdxT) + body_v * sizeof(ValT) +
                                              body_r * sizeof(IdxT));
    __syncwarp();  // every lane is done reading the previous task's stage
    if (lane == 0 && tx != 0) {
      mbar_arrive_expect_tx(bar, tx);
      if (body_c) tma_bulk_g2s(st.col + pre_c + head_c, col + ns + head_c, body_c * sizeof(IdxT), bar, pol_stream);
      if (body_v) tma_bulk_g2s(st.val + pre_v + head_v, val + ns + head_v, body_v * sizeof(ValT), bar, pol_stream);
      if (body_r) tma_bulk_g2s(st.crow + pre_r + head_r, crow + rs + head_r, body_r * sizeof(IdxT), bar, pol_stream);
    }
    seg_copy_edges(st.col, col + ns, cnt_nz, pre_c, head_c, body_c, lane);
    seg_copy_edges(st.val, val + ns, cnt_nz, pre_v, head_v, body_v, lane);
    seg_copy_edges(st.crow, crow + rs, cnt_row, pre_r, head_r, body_r, lane);
    if (tx != 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    __syncwarp();

    const IdxT* scol = st.col + pre_c;   // scol[e]  = col[ns + e]
    const ValT* sval = st.val + pre_v;   // sval[e]  = val[ns + e]
    const IdxT* srow = st.crow + pre_r;  // srow[i]  = crow[rs + i]

    float acc[CH][VEC];

    // acc += sum over elements e in [e0, e1) handled by this lane's group
    auto accumulate = [&](int e0, int e1) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
#pragma unroll
        for (int i = 0; i < VEC; ++i) acc[ch][i] = 0.f;
      int e = e0 + grp;
      for (; e + (U - 1) * G < e1; e += U * G) {
        IdxT c[U];
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          c[u] = scol[e + u * G];
          v[u] = to_float(sval[e + u * G]);
          if (static_cast<unsigned long long>(c[u]) >= static_cast<unsigned long long>(p.cols)) { c[u] = 0; v[u] = 0.f; }
        }
        float x[U][CH][VEC];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch))
              RowVec<DT, VEC>::load(Bl + static_cast<size_t>(c[u]) * n + ch * LPR * VEC, x[u][ch]);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch))
#pragma unroll
              for (int i = 0; i < VEC; ++i) acc[ch][i] = fmaf(v[u], x[u][ch][i], acc[ch][i]);
      }
      for (; e < e1; e += G) {
        IdxT c = scol[e];
        float v = to_float(sval[e]);
        if (static_cast<unsigned long long>(c) >= static_cast<unsigned long long>(p.cols)) { c = 0; v = 0.f; }
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (chmask & (1u << ch)) {
            float x[VEC];
            RowVec<DT, VEC>::load(Bl + static_cast<size_t>(c) * n + ch * LPR * VEC, x);
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[ch][i] = fmaf(v, x[i], acc[ch][i]);
          }
      }
      if constexpr (G > 1) {  // combine the G concurrent partial rows (fixed xor tree)
#pragma unroll
        for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[ch][i] += __shfl_xor_sync(0xffffffffu, acc[ch][i], off);
      }
    };

    const bool started_earlier = static_cast<int>(srow[0]) < ns;
    int e = 0;
    for (int r = rs; r < re; ++r) {
      const int e_end = static_cast<int>(srow[r - rs + 1]) - ns;
      accumulate(e, e_end);
      e = e_end;
      if (grp == 0) {
        if (!kF32Out && r == rs && started_earlier) {
          float* dst = headbuf + static_cast<size_t>(k) * n;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch)) store_f32<VEC>(dst + ch * LPR * VEC, acc[ch]);
        } else {
          DT* dst = Cl + static_cast<size_t>(r) * n;
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch)) RowVec<DT, VEC>::store_stream(dst + ch * LPR * VEC, acc[ch]);
        }
      }
    }
    if (cnt_nz > e) {  // trailing partial row `re`
      accumulate(e, cnt_nz);
      if (grp == 0) {
        float* dst = carry + static_cast<size_t>(k) * n;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (chmask & (1u << ch)) store_f32<VEC>(dst + ch * LPR * VEC, acc[ch]);
      }
    }
  }
}

// Adds the carries of every row that spans several tasks, in ascending task order, on top of the
// row's final segment.  One warp per task k that *ends* a row which started earlier.
template <typename DT, typename IdxT, int VEC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) spmm_fixup_kernel(const FwdParams p) {
  constexpr bool kF32Out = sizeof(DT) == 4;
  const int k = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (k <= 0 || k >= p.P) return;
  const int2 ps = __ldg(&p.part[k]);
  const int rs = ps.x, ns = ps.y;
  const int re = __ldg(&p.part[k + 1]).x;
  if (re <= rs) return;  // no row ends here
  const int cr = static_cast<int>(static_cast<const IdxT*>(p.crow)[rs]);
  if (cr >= ns) return;  // the row started in this task: already complete
  int j = k - 1;         // tasks j..k-1 each hold >= 1 non-zero of row rs
  while (j > 0 && __ldg(&p.part[j]).y > cr) --j;
  const int n = p.n;
  DT* crow_out = static_cast<DT*>(p.C) + static_cast<size_t>(rs) * n;
  for (int c0 = lane * VEC; c0 < n; c0 += 32 * VEC) {
    float sum[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) sum[i] = 0.f;
    for (int t = j; t < k; ++t) {
      float x[VEC];
      load_f32<VEC>(p.carry + static_cast<size_t>(t) * n + c0, x);
#pragma unroll
      for (int i = 0; i < VEC; ++i) sum[i] += x[i];
    }
    float h[VEC];
    if constexpr (kF32Out) {
      load_f32<VEC>(reinterpret_cast<const float*>(crow_out) + c0, h);
    } else {
      load_f32<VEC>(p.head + static_cast<size_t>(k) * n + c0, h);
    }
#pragma unroll
    for (int i = 0; i < VEC; ++i) sum[i] += h[i];
    RowVec<DT, VEC>::store_stream(crow_out + c0, sum);
  }
}

// part[k] = merge-path split of diagonal min(k*items, rows+nnz), k = 0..P.
template <typename IdxT>
__global__ void task_partition_kernel(const IdxT* __restrict__ crow, int rows, int nnz, int items,
                                      int P, int2* __restrict__ part) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > P) return;
  const long long total = static_cast<long long>(rows) + nnz;
  long long d = static_cast<long long>(k) * items;
  if (d > total) d = total;
  const long long r = merge_path_search<IdxT>(crow, rows, nnz, d);
  part[k] = make_int2(static_cast<int>(r), static_cast<int>(d - r));
}

}  // namespace ofspmm
