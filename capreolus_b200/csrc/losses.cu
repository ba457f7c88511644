// Pairwise losses of the training loop.
// Pairwise hinge loss and its gradient (capreolus/reranker/common.py:7,101-103):
//   MarginRankingLoss(margin=1, reduction="mean")(pos, neg, +1) = mean_b max(0, 1 - (pos_b - neg_b)).
// One CTA; B is a training batch (tens of pairs).  Fixed-order reduction -> deterministic loss.
#include "common.cuh"

namespace capr {

__global__ void __launch_bounds__(256) pair_hinge_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int B,
                                                         float* loss, float* gpos, float* gneg) {
  __shared__ float part[8];
  const int tid = threadIdx.x;
  const float invB = 1.0f / (float)B;
  float acc = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float m = 1.0f - (pos[i] - neg[i]);
    const bool active = m > 0.f;
    acc += active ? m : 0.f;
    if (gpos) gpos[i] = active ? -invB : 0.f;
    if (gneg) gneg[i] = active ? invB : 0.f;
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    loss[0] = t * invB;
  }
}

// Pairwise softmax loss (capreolus/reranker/common.py:96-98): mean_b (1 - softmax([pos_b, neg_b])[0]) = mean_b sigmoid(neg_b - pos_b).
// d/dpos_b = -p(1-p)/B, d/dneg_b = +p(1-p)/B with p = softmax[0].
__global__ void __launch_bounds__(256) pair_softmax_kernel(const float* __restrict__ pos, const float* __restrict__ neg, int B,
                                                           float* loss, float* gpos, float* gneg) {
  __shared__ float part[8];
  const int tid = threadIdx.x;
  const float invB = 1.0f / (float)B;
  float acc = 0.f;
  for (int i = tid; i < B; i += blockDim.x) {
    const float a = pos[i], b = neg[i], m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    const float p = ea / (ea + eb);  // softmax(dim=1)[:, 0], max-subtracted like torch
    acc += 1.0f - p;
    const float gr = p * (1.0f - p) * invB;
    if (gpos) gpos[i] = -gr;
    if (gneg) gneg[i] = gr;
  }
  acc = warp_sum(acc);
  if ((tid & 31) == 0) part[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
    loss[0] = t * invB;
  }
}

}  // namespace capr

extern "C" int capr_pair_softmax(const float* pos, const float* neg, int B, float* loss, float* grad_pos, float* grad_neg,
                                 capr_stream_t stream) {
  capr::DeviceGuard device_guard(pos);
  CAPR_REQUIRE(B > 0, CAPR_ERR_BAD_SHAPE, "capr_pair_softmax: B=%d must be positive", B);
  CAPR_REQUIRE(pos && neg && loss, CAPR_ERR_BAD_POINTER, "capr_pair_softmax: null pointer");
  capr::pair_softmax_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pos, neg, B, loss, grad_pos, grad_neg);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}

extern "C" int capr_pair_hinge(const float* pos, const float* neg, int B, float* loss, float* grad_pos, float* grad_neg,
                               capr_stream_t stream) {
  capr::DeviceGuard device_guard(pos);  // act on the device that owns the caller's buffers
  CAPR_REQUIRE(B > 0, CAPR_ERR_BAD_SHAPE, "capr_pair_hinge: B=%d must be positive", B);
  CAPR_REQUIRE(pos && neg && loss, CAPR_ERR_BAD_POINTER, "capr_pair_hinge: null pointer");
  capr::pair_hinge_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pos, neg, B, loss, grad_pos, grad_neg);
  CAPR_CHECK_CUDA(cudaGetLastError());
  return CAPR_OK;
}
