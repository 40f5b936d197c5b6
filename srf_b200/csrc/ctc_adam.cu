// ctc_adam.cu -- the steps either side of the routing stack in a training / decoding step
// (SURVEY.md 8f "next-2" and "next-3"), sm_100a:
//
//   ctc_greedy_kernel   greedy CTC decode used by the parity criterion: argmax per routing frame
//                       for s < input_length, collapse repeats, drop blank (= class_n - 1,
//                       tfsr/trainer_sr.py:133-134; the reference itself decodes with beam 100,
//                       trainer_sr.py:110-112).
//   ctc_loss_kernel     CTC negative log-likelihood and its gradient w.r.t. the logits
//                       (tf.nn.ctc_loss(labels, logits, label_length, logit_length,
//                       logits_time_major=False, blank_index=blank), trainer_sr.py:64-66):
//                       log-softmax inside, alpha/beta recursions in log space, one CTA per
//                       utterance, one thread per extended-label state.
//   adam_kernel         Keras Adam (beta1 .9, beta2 .98, eps 1e-9 in egs/conf/*.conf) on a flat
//                       parameter buffer with the learning rate of the warm-up schedule
//                       tfsr/helper/train_helper.py:32-56 supplied by the host.

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "routing_kernels.h"

namespace srf {

namespace {
__device__ __forceinline__ float log_add(float a, float b) {
  if (a == -CUDART_INF_F) return b;
  if (b == -CUDART_INF_F) return a;
  const float m = fmaxf(a, b);
  return m + log1pf(expf(-fabsf(a - b)));
}
}  // namespace

// ---------------------------------------------------------------------------------------
// greedy decode: one warp per utterance
// ---------------------------------------------------------------------------------------
__global__ void ctc_greedy_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lens,
                                  int B, int S, int C, int blank, int32_t* __restrict__ out_ids,
                                  int32_t* __restrict__ out_lens) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  int n = lens[b];
  n = n < 0 ? 0 : (n > S ? S : n);
  int prev = -1, cnt = 0;
  for (int s = 0; s < n; ++s) {
    const float* row = logits + ((size_t)b * S + s) * C;
    float best = -CUDART_INF_F;
    int arg = 0x7fffffff;
    for (int c = lane; c < C; c += 32) {
      const float v = row[c];
      if (v > best) {  // first maximum wins inside a lane (ascending c)
        best = v;
        arg = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ov > best || (ov == best && oa < arg)) {  // ties -> lowest class index (argmax rule)
        best = ov;
        arg = oa;
      }
    }
    if (arg != prev && arg != blank) {
      if (lane == 0) out_ids[(size_t)b * S + cnt] = arg;
      ++cnt;
    }
    prev = arg;
  }
  if (lane == 0) out_lens[b] = cnt;
}

void launch_ctc_greedy(const float* logits, const int32_t* lens, int B, int S, int C, int blank,
                       int32_t* out_ids, int32_t* out_lens, cudaStream_t stream) {
  const int warps = 4;
  ctc_greedy_kernel<<<(B + warps - 1) / warps, warps * 32, 0, stream>>>(logits, lens, B, S, C, blank,
                                                                          out_ids, out_lens);
}

// ---------------------------------------------------------------------------------------
// CTC loss + gradient.  grid = B, block = CTC_THREADS; dynamic smem = 2*(2*Lmax+1) floats + C floats.
// scratch: alpha [B][S][2*Lmax+1] floats (global), lsm = log-softmax kept in d_logits' storage
// until it is overwritten by the gradient.
// ---------------------------------------------------------------------------------------
constexpr int CTC_THREADS = 256;

__global__ void __launch_bounds__(CTC_THREADS)
ctc_loss_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                const int32_t* __restrict__ in_lens, const int32_t* __restrict__ lab_lens, int S, int C,
                int Lmax, int blank, float scale, float* __restrict__ loss, float* __restrict__ d_logits,
                float* __restrict__ alpha_ws) {
  extern __shared__ float sm[];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int NS = 2 * Lmax + 1;
  float* cur = sm;             // [NS]
  float* nxt = sm + NS;        // [NS]
  float* cls = sm + 2 * NS;    // [C] per-frame class accumulator
  __shared__ float s_logp;
  int T = in_lens[b];
  T = T < 0 ? 0 : (T > S ? S : T);
  int L = lab_lens[b];
  L = L < 0 ? 0 : (L > Lmax ? Lmax : L);
  const int ns = 2 * L + 1;
  const int32_t* lab = labels + (size_t)b * Lmax;
  float* lsm = d_logits + (size_t)b * S * C;     // log-softmax, later the gradient
  float* alpha = alpha_ws + (size_t)b * S * NS;
  auto ext = [&](int s) { return (s & 1) ? lab[s >> 1] : blank; };
  // labels outside [0, C) would index the log-softmax rows and the class accumulator out of bounds: such an
  // utterance is treated like an infeasible alignment (loss 0, zero gradient)
  {
    int bad = 0;
    for (int l = tid; l < L; l += CTC_THREADS) bad |= (lab[l] < 0 || lab[l] >= C);
    if (__syncthreads_or(bad)) {
      for (size_t e = tid; e < (size_t)S * C; e += CTC_THREADS) lsm[e] = 0.f;
      if (tid == 0) loss[b] = 0.f;
      return;
    }
  }

  // log-softmax of every frame (one warp per frame)
  for (int t = tid >> 5; t < S; t += CTC_THREADS >> 5) {
    const float* row = logits + ((size_t)b * S + t) * C;
    const int lane = tid & 31;
    float m = -CUDART_INF_F;
    for (int c = lane; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float z = 0.f;
    for (int c = lane; c < C; c += 32) z += expf(row[c] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
    const float lz = m + logf(z);
    for (int c = lane; c < C; c += 32) lsm[(size_t)t * C + c] = row[c] - lz;
  }
  __syncthreads();
  if (T == 0) {  // empty input: loss 0 if the label is empty, else infinite -> zeroed (zero_infinity)
    for (int e = tid; e < S * C; e += CTC_THREADS) lsm[e] = 0.f;
    if (tid == 0) loss[b] = 0.f;
    return;
  }
  // ---- alpha ----
  for (int s = tid; s < NS; s += CTC_THREADS) {
    float v = -CUDART_INF_F;
    if (s == 0) v = lsm[blank];
    if (s == 1 && ns > 1) v = lsm[ext(1)];
    cur[s] = v;
    alpha[s] = v;
  }
  __syncthreads();
  for (int t = 1; t < T; ++t) {
    for (int s = tid; s < ns; s += CTC_THREADS) {
      float v = cur[s];
      if (s >= 1) v = log_add(v, cur[s - 1]);
      if (s >= 2 && (s & 1) && ext(s) != ext(s - 2)) v = log_add(v, cur[s - 2]);
      v = (v == -CUDART_INF_F) ? v : v + lsm[(size_t)t * C + ext(s)];
      nxt[s] = v;
      alpha[(size_t)t * NS + s] = v;
    }
    __syncthreads();
    float* tmp = cur;
    cur = nxt;
    nxt = tmp;
  }
  if (tid == 0) {
    float lp = cur[ns - 1];
    if (ns > 1) lp = log_add(lp, cur[ns - 2]);
    s_logp = lp;
    loss[b] = (lp == -CUDART_INF_F) ? 0.f : -lp * 1.0f;  // infeasible alignment -> 0 (zero_infinity)
  }
  __syncthreads();
  const float logp = s_logp;
  const bool feasible = logp != -CUDART_INF_F;
  // ---- beta (backward in time) fused with the gradient ----
  // beta_t(s) here EXCLUDES the emission at t, so alpha_t(s) + beta_t(s) is the log-probability of
  // all alignments through state s at time t.
  for (int s = tid; s < NS; s += CTC_THREADS)
    cur[s] = (s < ns && (s == ns - 1 || s == ns - 2)) ? 0.f : -CUDART_INF_F;
  __syncthreads();
  for (int t = T - 1; t >= 0; --t) {
    for (int c = tid; c < C; c += CTC_THREADS) cls[c] = 0.f;
    __syncthreads();
    if (feasible) {
      for (int s = tid; s < ns; s += CTC_THREADS) {
        const float ab = alpha[(size_t)t * NS + s] + cur[s];
        if (ab != -CUDART_INF_F) atomicAdd(&cls[ext(s)], expf(ab - logp));
      }
    }
    __syncthreads();
    for (int c = tid; c < C; c += CTC_THREADS) {
      const float l = lsm[(size_t)t * C + c];
      lsm[(size_t)t * C + c] = feasible ? scale * (expf(l) - cls[c]) : 0.f;
    }
    // beta_{t-1}(s) = logsum over successors s' of beta_t(s') + lsm_t(ext(s'))  (lsm_t read BEFORE
    // it was overwritten: recompute the emission from cls-free data kept in nxt)
    __syncthreads();
    if (t > 0) {
      // emissions at time t are needed; they were just overwritten by the gradient, so they are
      // reconstructed from the logits row (log-softmax again for the few extended labels)
      const float* row = logits + ((size_t)b * S + t) * C;
      __shared__ float s_lz;
      if (tid < 32) {
        float m = -CUDART_INF_F;
        for (int c = tid; c < C; c += 32) m = fmaxf(m, row[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float z = 0.f;
        for (int c = tid; c < C; c += 32) z += expf(row[c] - m);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
        if (tid == 0) s_lz = m + logf(z);
      }
      __syncthreads();
      const float lz = s_lz;
      for (int s = tid; s < ns; s += CTC_THREADS) {
        float v = cur[s] == -CUDART_INF_F ? cur[s] : cur[s] + (row[ext(s)] - lz);
        if (s + 1 < ns) {
          const float w = cur[s + 1];
          if (w != -CUDART_INF_F) v = log_add(v, w + (row[ext(s + 1)] - lz));
        }
        if (s + 2 < ns && (s & 1) && ext(s) != ext(s + 2)) {
          const float w = cur[s + 2];
          if (w != -CUDART_INF_F) v = log_add(v, w + (row[ext(s + 2)] - lz));
        }
        nxt[s] = v;
      }
      __syncthreads();
      float* tmp = cur;
      cur = nxt;
      nxt = tmp;
    }
  }
  // frames beyond the input length carry no gradient
  for (int e = tid + T * C; e < S * C; e += CTC_THREADS) lsm[e] = 0.f;
}

cudaError_t launch_ctc_loss(const float* logits, const int32_t* labels, const int32_t* in_lens,
                            const int32_t* lab_lens, int B, int S, int C, int Lmax, int blank,
                            float scale, float* loss, float* d_logits, float* alpha_ws,
                            cudaStream_t stream) {
  const size_t smem = sizeof(float) * (2 * (size_t)(2 * Lmax + 1) + C);
  cudaError_t e = cudaFuncSetAttribute(ctc_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem);
  if (e != cudaSuccess) return e;
  ctc_loss_kernel<<<B, CTC_THREADS, smem, stream>>>(logits, labels, in_lens, lab_lens, S, C, Lmax, blank,
                                                    scale, loss, d_logits, alpha_ws);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// Adam on a flat buffer (tf.keras.optimizers.Adam semantics)
// ---------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, long long n, float lr_t, float beta1, float beta2,
                            float eps) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gi = g[i];
    const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

void launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                 float beta2, float eps, long long step, cudaStream_t stream) {
  // Keras: lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
  const double t = (double)step;
  const float lr_t = (float)((double)lr * sqrt(1.0 - pow((double)beta2, t)) / (1.0 - pow((double)beta1, t)));
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<blocks, 256, 0, stream>>>(p, g, m, v, n, lr_t, beta1, beta2, eps);
}

}  // namespace srf
