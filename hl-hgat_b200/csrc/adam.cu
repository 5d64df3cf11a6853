// Adam over ONE flat parameter / gradient / moment buffer (hl_adam_flat).
//
// The reference scripts call torch.optim.Adam(model.parameters(), lr, weight_decay) (main_zinc_HL_HGCNN_dense_int3_pyr.py:213);
// on 134 small tensors even its fused multi-tensor form is five launches of ~31 us (0.16 ms, 3 % of the 4.7 ms ZINC step,
// running alone at the tail of the step graph).  The gradients already live in one flat bucket for the all-reduce; with the
// parameters and both moments flat as well the whole optimizer step is one streaming pass: 28 bytes per parameter.
// Semantics are torch.optim.Adam's (L2 weight decay added to the gradient, bias-corrected moments, eps outside the root);
// `grad_scale` folds the 1 / world_size of the data-parallel average in.  The step counter lives on the device (capturable):
// a one-thread tick kernel advances it and publishes the two bias-correction factors, then the streaming kernel runs.
#include "common.cuh"

namespace hl {

// state[0] = step (as float), state[1] = 1 / (1 - beta1^t), state[2] = 1 / sqrt(1 - beta2^t)
__global__ void adam_tick_kernel(float* __restrict__ state, float beta1, float beta2) {
  const float t = state[0] + 1.f;
  state[0] = t;
  state[1] = 1.f / (1.f - powf(beta1, t));
  state[2] = rsqrtf(1.f - powf(beta2, t));
}

__global__ void __launch_bounds__(256)
adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ state, float lr, float beta1, float beta2, float eps, float weight_decay,
                 float grad_scale) {
  const float c1 = state[1], c2 = state[2];
  const float step_size = lr * c1;
  const int64_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = p4[i], gg = g4[i], mm = m4[i], vv = v4[i];
    float* P = reinterpret_cast<float*>(&pp);
    float* G = reinterpret_cast<float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm);
    float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gr = G[k] * grad_scale + weight_decay * P[k];
      M[k] = beta1 * M[k] + (1.f - beta1) * gr;
      V[k] = beta2 * V[k] + (1.f - beta2) * gr * gr;
      P[k] -= step_size * M[k] / (sqrtf(V[k]) * c2 + eps);
    }
    p4[i] = pp; m4[i] = mm; v4[i] = vv;
  }
  // tail (n % 4 elements)
  const int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const float gr = g[i] * grad_scale + weight_decay * p[i];
    const float mk = beta1 * m[i] + (1.f - beta1) * gr;
    const float vk = beta2 * v[i] + (1.f - beta2) * gr * gr;
    m[i] = mk; v[i] = vk;
    p[i] -= step_size * mk / (sqrtf(vk) * c2 + eps);
  }
}

}  // namespace hl

extern "C" int hl_adam_flat(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float* state,
                            float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                            hl_stream_t stream) {
  using namespace hl;
  if (n < 0 || !state) return HL_ERR_INVALID;
  if (n == 0) return HL_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq) return HL_ERR_INVALID;
  if (!aligned_to(params, 16) || !aligned_to(grads, 16) || !aligned_to(exp_avg, 16) || !aligned_to(exp_avg_sq, 16)) return HL_ERR_ALIGN;
  cudaStream_t st = as_stream(stream);
  adam_tick_kernel<<<1, 1, 0, st>>>(state, beta1, beta2);
  HL_LAUNCH_CHECK("adam_tick_kernel");
  int64_t blocks = ((n >> 2) + 255) / 256;
  const int64_t cap = (int64_t)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_flat_kernel<<<(int)blocks, 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, n, state, lr, beta1, beta2, eps, weight_decay,
                                                grad_scale);
  HL_LAUNCH_CHECK("adam_flat_kernel");
  return HL_OK;
}
