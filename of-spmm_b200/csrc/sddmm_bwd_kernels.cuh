// sddmm_bwd_kernels.cuh — SDDMM value gradient and the atomic A^T·dY scatter, both on the same
// merge-path task list as the forward kernel (see spmm_kernels.cuh for the scheduling notes).
#pragma once
#include "spmm_kernels.cuh"

namespace ofspmm {

struct SddmmParams {
  const void* crow;
  const void* col;
  const void* dY;   // rows x n
  const void* B;    // cols x n
  void* dval;       // nnz
  const int2* part;
  long long cols;
  int rows;
  int nnz;
  int n;
  int P;
};

// dval[p] = <dY[i,:], B[col[p],:]>.  LPR lanes span one dense row (CH chunks of VEC each; CH == 0
// means "loop over n", used when the row does not fit the register tile).  The dY row chunk
// stays in registers while the task walks the row's non-zeros; each non-zero costs one coalesced
// gather of its B row, VEC*CH FMAs and a log2(LPR) xor-shuffle reduction.
template <typename DT, typename ValT, typename IdxT, int VEC, int LPR, int CH, int ITEMS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) sddmm_merge_kernel(const SddmmParams p) {
  constexpr int G = 32 / LPR;
  constexpr int CHR = CH > 0 ? CH : 1;
  using Stage = TaskStage<IdxT, float, ITEMS>;  // val array reused as the output staging buffer

  extern __shared__ __align__(16) unsigned char smem_raw[];
  Stage* stages = reinterpret_cast<Stage*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * WARPS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int lig = lane % LPR;
  Stage& st = stages[warp];
  uint64_t* bar = &bars[warp];
  if (lane == 0) mbar_init(bar, 1);
  fence_mbar_init();
  __syncwarp();

  const uint64_t pol_stream = l2_policy_evict_first();
  const IdxT* __restrict__ crow = static_cast<const IdxT*>(p.crow);
  const IdxT* __restrict__ col = static_cast<const IdxT*>(p.col);
  const DT* __restrict__ B = static_cast<const DT*>(p.B);
  const DT* __restrict__ dY = static_cast<const DT*>(p.dY);
  ValT* __restrict__ dval = static_cast<ValT*>(p.dval);
  const int n = p.n;
  const int c_lane = lig * VEC;
  unsigned chmask = 0;
#pragma unroll
  for (int ch = 0; ch < CHR; ++ch)
    if (c_lane + ch * LPR * VEC < n) chmask |= 1u << ch;

  const int total_warps = gridDim.x * WARPS;
  uint32_t phase = 0;

  for (int k = blockIdx.x * WARPS + warp; k < p.P; k += total_warps) {
    const int2 ps = __ldg(&p.part[k]);
    const int2 pe = __ldg(&p.part[k + 1]);
    const int rs = ps.x, ns = ps.y, re = pe.x, ne = pe.y;
    const int cnt_nz = ne - ns;
    const int cnt_row = re - rs + 1;
    if (cnt_nz == 0) continue;

    int pre_c, head_c, body_c, pre_r, head_r, body_r;
    seg_plan(col + ns, cnt_nz, pre_c, head_c, body_c);
    seg_plan(crow + rs, cnt_row, pre_r, head_r, body_r);
    const uint32_t tx = static_cast<uint32_t>((body_c + body_r) * sizeof(IdxT));
    __syncwarp();
    if (lane == 0 && tx != 0) {
      mbar_arrive_expect_tx(bar, tx);
      if (body_c) tma_bulk_g2s(st.col + pre_c + head_c, col + ns + head_c, body_c * sizeof(IdxT), bar, pol_stream);
      if (body_r) tma_bulk_g2s(st.crow + pre_r + head_r, crow + rs + head_r, body_r * sizeof(IdxT), bar, pol_stream);
    }
    seg_copy_edges(st.col, col + ns, cnt_nz, pre_c, head_c, body_c, lane);
    seg_copy_edges(st.crow, crow + rs, cnt_row, pre_r, head_r, body_r, lane);
    if (tx != 0) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    __syncwarp();
    const IdxT* scol = st.col + pre_c;
    const IdxT* srow = st.crow + pre_r;
    float* sout = st.val;  // sout[e] = result for non-zero ns + e

    int e = 0;
    for (int r = rs; r <= re; ++r) {
      const int e_end = r < re ? static_cast<int>(srow[r - rs + 1]) - ns : cnt_nz;
      if (e_end <= e) continue;
      const DT* yrow = dY + static_cast<size_t>(r) * n + c_lane;
      float y[CHR][VEC];
      if constexpr (CH > 0) {
#pragma unroll
        for (int ch = 0; ch < CH; ++ch) {
          if (chmask & (1u << ch)) {
            RowVec<DT, VEC>::load(yrow + ch * LPR * VEC, y[ch]);
          } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) y[ch][i] = 0.f;
          }
        }
      }
      for (int q = e + grp; q < e_end + grp; q += G) {  // uniform trip count across groups
        const bool live = q < e_end;
        IdxT c = live ? scol[q] : 0;
        const bool ok = live && static_cast<unsigned long long>(c) < static_cast<unsigned long long>(p.cols);
        if (!ok) c = 0;
        const DT* brow = B + static_cast<size_t>(c) * n + c_lane;
        float dot = 0.f;
        if constexpr (CH > 0) {
          float x[CH][VEC];
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch)) RowVec<DT, VEC>::load(brow + ch * LPR * VEC, x[ch]);
#pragma unroll
          for (int ch = 0; ch < CH; ++ch)
            if (chmask & (1u << ch))
#pragma unroll
              for (int i = 0; i < VEC; ++i) dot = fmaf(y[ch][i], x[ch][i], dot);
        } else {
          for (int c0 = 0; c_lane + c0 < n; c0 += LPR * VEC) {
            float x[VEC], yy[VEC];
            RowVec<DT, VEC>::load(brow + c0, x);
            RowVec<DT, VEC>::load(yrow + c0, yy);
#pragma unroll
            for (int i = 0; i < VEC; ++i) dot = fmaf(yy[i], x[i], dot);
          }
        }
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
        if (live && lig == 0) sout[q] = ok ? dot : 0.f;
      }
      e = e_end;
    }
    __syncwarp();
    // coalesced write-back of the task's results
    for (int q = lane; q < cnt_nz; q += 32) dval[ns + q] = from_float<ValT>(sout[q]);
  }
}

struct BwdAtomicParams {
  const void* crow;
  const void* col;
  const void* val;
  const void* dY;  // rows x n
  float* acc;      // cols x n fp32 accumulator (zeroed by the caller)
  const int2* part;
  long long cols;
  int rows;
  int nnz;
  int n;
  int P;
};

__device__ __forceinline__ void red_add_f32(float* p, const float (&v)[1]) { atomicAdd(p, v[0]); }
__device__ __forceinline__ void red_add_f32(float* p, const float (&v)[4]) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3])
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, const float (&v)[8]) {
  const float a[4] = {v[0], v[1], v[2], v[3]};
  const float b[4] = {v[4], v[5], v[6], v[7]};
  red_add_f32(p, a);
  red_add_f32(p + 4, b);
}

// acc[col[p], :] += val[p] * dY[i, :] with 16-byte vector reductions (red.global.add.v4.f32).
// Route (2) of ofspmm_bwd_b: needs no transposed copy of A, order-nondeterministic like the
// reference's atomic scatter (oneflow/user/kernels/unsorted_segment_sum_kernel_util.cu:96-117).
template <typename DT, typename ValT, typename IdxT, int VEC, int LPR, int CH, int ITEMS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) bwd_atomic_kernel(const BwdAtomicParams p) {
  constexpr int G = 32 / LPR;
  using Stage = TaskStage<IdxT, ValT, ITEMS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Stage* stages = reinterpret_cast<Stage*>(smem_raw);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + sizeof(Stage) * WARPS);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int grp = lane / LPR;
  const int lig = lane % LPR;
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
  const int col0 = blockIdx.y * (LPR * VEC * CH) + lig * VEC;
  unsigned chmask = 0;
#pragma unroll
  for (int ch = 0; ch < CH; ++ch)
    if (col0 + ch * LPR * VEC < n) chmask |= 1u << ch;
  const DT* __restrict__ dYl = static_cast<const DT*>(p.dY) + col0;
  float* __restrict__ accl = p.acc + col0;

  const int total_warps = gridDim.x * WARPS;
  uint32_t phase = 0;
  for (int k = blockIdx.x * WARPS + warp; k < p.P; k += total_warps) {
    const int2 ps = __ldg(&p.part[k]);
    const int2 pe = __ldg(&p.part[k + 1]);
    const int rs = ps.x, ns = ps.y, re = pe.x, ne = pe.y;
    const int cnt_nz = ne - ns;
    const int cnt_row = re - rs + 1;
    if (cnt_nz == 0) continue;
    int pre_c, head_c, body_c, pre_v, head_v, body_v, pre_r, head_r, body_r;
    seg_plan(col + ns, cnt_nz, pre_c, head_c, body_c);
    seg_plan(val + ns, cnt_nz, pre_v, head_v, body_v);
    seg_plan(crow + rs, cnt_row, pre_r, head_r, body_r);
    const uint32_t tx = static_cast<uint32_t>(body_c * sizeof(IdxT) + body_v * sizeof(ValT) +
                                              body_r * sizeof(IdxT));
    __syncwarp();
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
    const IdxT* scol = st.col + pre_c;
    const ValT* sval = st.val + pre_v;
    const IdxT* srow = st.crow + pre_r;

    int e = 0;
    for (int r = rs; r <= re; ++r) {
      const int e_end = r < re ? static_cast<int>(srow[r - rs + 1]) - ns : cnt_nz;
      if (e_end <= e) continue;
      float y[CH][VEC];
#pragma unroll
      for (int ch = 0; ch < CH; ++ch)
        if (chmask & (1u << ch)) RowVec<DT, VEC>::load(dYl + static_cast<size_t>(r) * n + ch * LPR * VEC, y[ch]);
      for (int q = e + grp; q < e_end; q += G) {
        const IdxT c = scol[q];
        if (static_cast<unsigned long long>(c) >= static_cast<unsigned long long>(p.cols)) continue;
        const float v = to_float(sval[q]);
        float* dst = accl + static_cast<size_t>(c) * n;
#pragma unroll
        for (int ch = 0; ch < CH; ++ch)
          if (chmask & (1u << ch)) {
            float t[VEC];
#pragma unroll
            for (int i = 0; i < VEC; ++i) t[i] = v * y[ch][i];
            red_add_f32(dst + ch * LPR * VEC, t);
          }
      }
      e = e_end;
    }
  }
}

// out[i] = DT(acc[i]) — the cast that follows an fp32 accumulation for bf16 outputs
// (reference idiom: oneflow/user/kernels/unsorted_segment_sum_kernel.cpp:162-186).
template <typename DT>
__global__ void cast_from_f32_kernel(const float* __restrict__ in, DT* __restrict__ out, long long count) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < count; i += stride)
    out[i] = from_float<DT>(in[i]);
}

}  // namespace ofspmm
