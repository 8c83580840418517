// k smallest squared-L2 distances per query row from a Gram matrix (retrieval after the descriptor
// all-gather; reference evaluate/global_eval/evaluation_retrieval.py:37-40 uses a host cKDTree).
//   d2[q,r] = qn[q] + rn[r] - 2*gram[q,r];  one warp per query: each lane keeps the k smallest of its
//   strided share in registers (sorted insert), then k rounds of warp-argmin pop the global order
//   (ties: smaller index first).
// The Gram form cancels catastrophically for near-identical descriptors (|d2| ~ 1e-6 against rounding noise of
// ~1e-7: measured, 12 % of 4096 random-weight descriptors did not find THEMSELVES first).  topk_rerank_kernel
// therefore recomputes the selected candidates' distances exactly, sum_d (q_d - r_d)^2, and re-sorts them; the
// Gram pass selects K + 8 candidates so that a mis-ranked boundary neighbour is still inside the set.
#include "common.cuh"

namespace dh3d {

template <int KC>
__global__ void __launch_bounds__(128)
topk_l2_kernel(const float* __restrict__ gram, const float* __restrict__ qn, const float* __restrict__ rn,
               int Q, int R, int ldg, int K, int32_t* __restrict__ idx, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const float* g = gram + (long long)q * ldg;
  const float nq = __ldg(qn + q);
  float v[KC];
  int id[KC];
#pragma unroll
  for (int j = 0; j < KC; ++j) { v[j] = CUDART_INF_F; id[j] = 0x7fffffff; }
  for (int r = lane; r < R; r += 32) {
    const float d = fmaxf(nq + __ldg(rn + r) - 2.f * __ldg(g + r), 0.f);
    if (d < v[KC - 1]) {
#pragma unroll
      for (int j = KC - 1; j > 0; --j) {
        if (d < v[j - 1]) { v[j] = v[j - 1]; id[j] = id[j - 1]; }
        else if (d < v[j]) { v[j] = d; id[j] = r; }
      }
      if (d < v[0]) { v[0] = d; id[0] = r; }
    }
  }
  for (int out = 0; out < K; ++out) {
    // lane with the smallest head (value, index)
    float bv = v[0];
    int bi = id[0];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov < bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { idx[(long long)q * K + out] = bi; val[(long long)q * K + out] = bv; }
    if (id[0] == bi) {  // pop
#pragma unroll
      for (int j = 0; j < KC - 1; ++j) { v[j] = v[j + 1]; id[j] = id[j + 1]; }
      v[KC - 1] = CUDART_INF_F;
      id[KC - 1] = 0x7fffffff;
    }
  }
}

// one warp per query: exact squared distances of its C candidates (lane-strided over the D features, warp sum),
// then lane j (j < C) ranks candidate j by counting the candidates ahead of it in (distance, index) order
__global__ void __launch_bounds__(128)
topk_rerank_kernel(const float* __restrict__ qd, const float* __restrict__ rd, int Q, int D, int C, int K,
                   const int32_t* __restrict__ cand, int32_t* __restrict__ idx, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const float* qv = qd + (long long)q * D;
  float my_d = CUDART_INF_F;
  int my_i = 0x7fffffff;
  for (int c = 0; c < C; ++c) {
    const int r = __ldg(cand + (long long)q * C + c);
    float acc = 0.f;
    if (r >= 0 && r != 0x7fffffff) {
      const float* rv = rd + (long long)r * D;
      for (int d = lane; d < D; d += 32) {
        const float t = __ldg(qv + d) - __ldg(rv + d);
        acc = fmaf(t, t, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == c) { my_d = (r >= 0 && r != 0x7fffffff) ? acc : CUDART_INF_F; my_i = r; }
  }
  int rank = 0;
  for (int c = 0; c < C; ++c) {
    const float od = __shfl_sync(0xffffffffu, my_d, c);
    const int oi = __shfl_sync(0xffffffffu, my_i, c);
    if (od < my_d || (od == my_d && oi < my_i)) ++rank;
  }
  if (lane < C && rank < K) {
    idx[(long long)q * K + rank] = my_i;
    val[(long long)q * K + rank] = my_d;
  }
}

int topk_l2_exact_launch(const float* gram, int ldg, const float* qn, const float* rn, const float* qd,
                         const float* rd, int Q, int R, int D, int K, int32_t* idx, float* val, int32_t* cand,
                         float* cand_val, cudaStream_t st);

int topk_l2_launch(const float* gram, int ldg, const float* qn, const float* rn, int Q, int R, int K,
                   int32_t* idx, float* val, cudaStream_t st) {
  if (!gram || !qn || !rn || !idx || !val) return DH3D_ERR_NULL;
  if (Q <= 0 || R <= 0 || K <= 0 || K > R || ldg < R) return DH3D_ERR_DIM;
  if (K > 32) return DH3D_ERR_UNSUPPORTED;
  const int blocks = ceil_div(Q * 32, 128);
  if (K <= 8) topk_l2_kernel<8><<<blocks, 128, 0, st>>>(gram, qn, rn, Q, R, ldg, K, idx, val);
  else if (K <= 16) topk_l2_kernel<16><<<blocks, 128, 0, st>>>(gram, qn, rn, Q, R, ldg, K, idx, val);
  else topk_l2_kernel<32><<<blocks, 128, 0, st>>>(gram, qn, rn, Q, R, ldg, K, idx, val);
  return launch_status();
}

// Gram selection of C = min(K + 8, 32, R) candidates, then the exact re-rank.  cand / cand_val: [Q, C] scratch.
int topk_l2_exact_launch(const float* gram, int ldg, const float* qn, const float* rn, const float* qd,
                         const float* rd, int Q, int R, int D, int K, int32_t* idx, float* val, int32_t* cand,
                         float* cand_val, cudaStream_t st) {
  if (!qd || !rd || !cand || !cand_val) return DH3D_ERR_NULL;
  if (D <= 0 || K > 24) return K > 24 ? DH3D_ERR_UNSUPPORTED : DH3D_ERR_DIM;
  int C = K + 8;
  if (C > 32) C = 32;
  if (C > R) C = R;
  int rc = topk_l2_launch(gram, ldg, qn, rn, Q, R, C, cand, cand_val, st);
  if (rc != DH3D_OK) return rc;
  topk_rerank_kernel<<<ceil_div(Q * 32, 128), 128, 0, st>>>(qd, rd, Q, D, C, K, cand, idx, val);
  return launch_status();
}

}  // namespace dh3d
