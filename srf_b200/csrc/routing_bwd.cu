// routing_bwd.cu -- backward of one routing layer (training mode of
// tfsr/model/sequence_router_naive.py:145-193; the reference differentiates the TF graph with
// tf.GradientTape, tfsr/trainer_sr.py:62-71).  Correctness-first FP32 version:
//
//   K_b1  ln_head_bwd_kernel   per frame (one warp, lane = output capsule): head
//         ln_o(length(.)) backward, dropout mask, LayerNorm(O*D) backward
//         -> dL/d(v_raw) + parameter gradients (atomics).
//   K_b2  route_layer_bwd_kernel  per chain (utterance for SDR, frame for DR): recomputes
//         u_hat in FP32 from the layer input and the packed weights (nothing of size
//         I*O*D is saved by the forward), re-runs the routing passes of the frame to recover
//         t_r / Vacc_r, then walks the passes backwards.  SDR frames are visited in reverse
//         time order and the gradient w.r.t. the carried output (Vacc_0 = previous frame's v)
//         is handed to the previous frame (BPTT).
//
// Per pass r (Vacc_{r-1} = v_0 + ... + v_{r-1}, v_0 = previous frame's output | 0):
//   a = u.Vacc_{r-1};  c = softmax_j a;  t_r = sum_i c u;  v_r = squash(t_r)
// backward, r = R..1, G_R = 0:
//   g_v = [r==R] g_out + G_r;   g_t = squash'(t_r) g_v
//   g_c[i,j] = g_t[j].u[i,j];   g_a = c (g_c - sum_j c g_c)
//   g_u[i,j] = c g_t[j] + g_a Vacc_{r-1}[j];      G_{r-1} = G_r + sum_i g_a u[i,j]
//   dbias += g_u;  dW[i,j,k,l] += g_u[j,k] x[i,l];  dx[i,l] += sum_{j,k} g_u[j,k] W[i,j,k,l]
// (SURVEY.md appendix B for R = 1.)  Parameter gradients are accumulated with fp32 atomics.

#include <cuda_runtime.h>
#include <math_constants.h>

#include "routing_kernels.h"

namespace srf {

namespace {
__device__ __forceinline__ float bw_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float bw_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
constexpr int BW_MAX_ITERS = 8;
}  // namespace

// ---------------------------------------------------------------------------------------
// K_b1: head + dropout + LayerNorm backward.  grid = frames / 8, block = 256 (warp per frame)
// ---------------------------------------------------------------------------------------
__global__ void ln_head_bwd_kernel(const BwdParams p) {
  const int lane = threadIdx.x & 31;
  const long long frame = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (frame >= (long long)p.B * p.S) return;
  const int O = p.O, D = p.D;
  const int n = O * D;
  const float* v = p.v_raw + frame * n;
  const bool do_ln = p.ln_gamma != nullptr;
  // pass 1: LayerNorm statistics of v_raw
  float mean = 0.f, rstd = 1.f;
  if (do_ln) {
    float s = 0.f;
    for (int e = lane; e < n; e += 32) s += v[e];
    mean = bw_sum(s) / (float)n;
    float q = 0.f;
    for (int e = lane; e < n; e += 32) {
      const float dv = v[e] - mean;
      q = fmaf(dv, dv, q);
    }
    rstd = 1.0f / sqrtf(bw_sum(q) / (float)n + p.ln_eps);
  }
  auto yval = [&](int e, float& xhat) {
    float y = v[e];
    xhat = 0.f;
    if (do_ln) {
      xhat = (y - mean) * rstd;
      y = xhat * p.ln_gamma[e] + p.ln_beta[e];
    }
    if (p.dropout_mask) y *= p.dropout_mask[frame * n + e];
    return y;
  };
  // head backward: d_len[j] (lane-strided over j)
  float hm = 0.f, hr = 0.f, mdz = 0.f, mdzz = 0.f;
  const bool head = p.d_logits != nullptr;
  if (head) {
    float s = 0.f;
    for (int j = lane; j < O; j += 32) {
      float l2 = 0.f, xh;
      for (int k = 0; k < D; ++k) {
        const float y = yval(j * D + k, xh);
        l2 = fmaf(y, y, l2);
      }
      s += sqrtf(l2 + p.length_eps);
    }
    hm = bw_sum(s) / (float)O;
    float q = 0.f;
    for (int j = lane; j < O; j += 32) {
      float l2 = 0.f, xh;
      for (int k = 0; k < D; ++k) {
        const float y = yval(j * D + k, xh);
        l2 = fmaf(y, y, l2);
      }
      const float dv = sqrtf(l2 + p.length_eps) - hm;
      q = fmaf(dv, dv, q);
    }
    hr = 1.0f / sqrtf(bw_sum(q) / (float)O + p.ln_eps);
    float a = 0.f, b = 0.f;
    for (int j = lane; j < O; j += 32) {
      float l2 = 0.f, xh;
      for (int k = 0; k < D; ++k) {
        const float y = yval(j * D + k, xh);
        l2 = fmaf(y, y, l2);
      }
      const float zhat = (sqrtf(l2 + p.length_eps) - hm) * hr;
      const float dl = p.d_logits[frame * O + j];
      atomicAdd(p.dhead_gamma + j, dl * zhat);
      atomicAdd(p.dhead_beta + j, dl);
      const float dz = dl * p.head_gamma[j];
      a += dz;
      b = fmaf(dz, zhat, b);
    }
    mdz = bw_sum(a) / (float)O;
    mdzz = bw_sum(b) / (float)O;
  }
  // d_y -> d_yln -> LayerNorm backward.  Two sweeps: the means of dxhat and dxhat*xhat first.
  auto dy_of = [&](int j, int k, float y, float len_j, float zhat_j) {
    float g = p.d_out ? p.d_out[frame * n + j * D + k] : 0.f;
    if (head) {
      const float dz = p.d_logits[frame * O + j] * p.head_gamma[j];
      const float dlen = hr * (dz - mdz - zhat_j * mdzz);
      g += dlen * y / len_j;
    }
    return g;
  };
  float m1 = 0.f, m2 = 0.f;
  for (int j = lane; j < O; j += 32) {
    float len_j = 1.f, zhat_j = 0.f;
    if (head) {
      float l2 = 0.f, xh;
      for (int k = 0; k < D; ++k) {
        const float y = yval(j * D + k, xh);
        l2 = fmaf(y, y, l2);
      }
      len_j = sqrtf(l2 + p.length_eps);
      zhat_j = (len_j - hm) * hr;
    }
    for (int k = 0; k < D; ++k) {
      const int e = j * D + k;
      float xhat;
      const float y = yval(e, xhat);
      float g = dy_of(j, k, y, len_j, zhat_j);
      if (p.dropout_mask) g *= p.dropout_mask[frame * n + e];
      if (do_ln) {
        atomicAdd(p.dgamma + e, g * xhat);
        atomicAdd(p.dbeta + e, g);
        const float dxh = g * p.ln_gamma[e];
        m1 += dxh;
        m2 = fmaf(dxh, xhat, m2);
      } else {
        p.d_raw[frame * n + e] = g;
      }
    }
  }
  if (!do_ln) return;
  m1 = bw_sum(m1) / (float)n;
  m2 = bw_sum(m2) / (float)n;
  for (int j = lane; j < O; j += 32) {
    float len_j = 1.f, zhat_j = 0.f;
    if (head) {
      float l2 = 0.f, xh;
      for (int k = 0; k < D; ++k) {
        const float y = yval(j * D + k, xh);
        l2 = fmaf(y, y, l2);
      }
      len_j = sqrtf(l2 + p.length_eps);
      zhat_j = (len_j - hm) * hr;
    }
    for (int k = 0; k < D; ++k) {
      const int e = j * D + k;
      float xhat;
      const float y = yval(e, xhat);
      float g = dy_of(j, k, y, len_j, zhat_j);
      if (p.dropout_mask) g *= p.dropout_mask[frame * n + e];
      const float dxh = g * p.ln_gamma[e];
      p.d_raw[frame * n + e] = rstd * (dxh - m1 - xhat * m2);
    }
  }
}

void launch_ln_head_bwd(const BwdParams& p, cudaStream_t stream) {
  const long long frames = (long long)p.B * p.S;
  const int warps = 8;
  ln_head_bwd_kernel<<<(unsigned)((frames + warps - 1) / warps), warps * 32, 0, stream>>>(p);
}

// ---------------------------------------------------------------------------------------
// K_b2: routing backward.  One CTA per chain; warps stride over the input capsules;
// lane = output capsule (j = q*32 + lane).
// ---------------------------------------------------------------------------------------
template <int T, int OPL, int NW>
__global__ void __launch_bounds__(NW * 32) route_layer_bwd_kernel(const BwdParams p) {
  constexpr int OP = 32 * OPL;
  constexpr int T4 = T / 4;
  constexpr int NT = NW * 32;
  constexpr int E = OPL * T * 32;  // (q*T+k)*32+lane
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chain = blockIdx.x;
  const int I = p.I, O = p.O, D = p.D, R = p.iters;
  float* xs = smem;                 // [I][T] window-gathered input of the frame
  float* red = xs + (size_t)I * T;  // [NW][E]
  float* tr = red + NW * E;         // [R][E]   t_r
  float* vacc = tr + BW_MAX_ITERS * E;  // [R+1][E] Vacc_0 .. Vacc_R
  float* gout = vacc + (BW_MAX_ITERS + 1) * E;  // [E] dL/dv of this frame (incl. BPTT carry)
  float* gt = gout + E;             // [E] g_t of the current pass
  float* gacc = gt + E;             // [E] G_r

  const float4* __restrict__ Wp = reinterpret_cast<const float4*>(p.Wp);
  const float* __restrict__ Bp = p.Bp;

  for (int e = tid; e < E; e += NT) gacc[e] = 0.f;  // BPTT carry into the last frame is zero
  __syncthreads();

  for (int step = p.nsteps - 1; step >= 0; --step) {
    const int b = p.sdr ? chain : chain / p.S;
    const int sf = p.sdr ? step : chain % p.S;
    const long long frame = (long long)b * p.S + sf;
    // window gather
    for (int idx = tid; idx < I * T; idx += NT) {
      const int l = idx % T, i = idx / T;
      float v = 0.f;
      if (l < p.d) {
        const int w = i / p.H, hc = i - w * p.H;
        const int src = sf - p.lpad + w;
        if (src >= 0 && src < p.S) v = p.emb[(((long long)b * p.S + src) * p.H + hc) * p.d + l];
      }
      xs[i * T + l] = v;
    }
    // g_out = dL/d v_raw[frame] + carry; Vacc_0 = previous frame's output (SDR) or 0
    for (int e = tid; e < E; e += NT) {
      const int ln = e & 31, qk = e >> 5, q = qk / T, k = qk % T;
      const int j = q * 32 + ln;
      const bool ok = j < O && k < D;
      float g = ok ? p.d_raw[(frame * O + j) * D + k] : 0.f;
      if (p.sdr) g += gacc[e];
      gout[e] = g;
      float v0 = 0.f;
      if (p.sdr && sf > 0 && ok) v0 = p.v_raw[((frame - 1) * O + j) * D + k];
      vacc[e] = v0;
    }
    __syncthreads();

    // ---------------- forward recompute: t_r, Vacc_r ----------------
    for (int r = 0; r < R; ++r) {
      float va[OPL][T], ta[OPL][T];
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) {
          va[q][k] = vacc[r * E + (q * T + k) * 32 + lane];
          ta[q][k] = 0.f;
        }
      for (int i = warp; i < I; i += NW) {
        float u[OPL][T], a[OPL];
        const float4* xrow = reinterpret_cast<const float4*>(xs) + (size_t)i * T4;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
#pragma unroll
          for (int k = 0; k < T; ++k) {
            float acc = Bp[(size_t)(i * T + k) * OP + jp];
#pragma unroll
            for (int c = 0; c < T4; ++c) {
              const float4 w4 = Wp[((size_t)(i * T + k) * T4 + c) * OP + jp];
              const float4 x4 = xrow[c];
              acc = fmaf(w4.x, x4.x, acc);
              acc = fmaf(w4.y, x4.y, acc);
              acc = fmaf(w4.z, x4.z, acc);
              acc = fmaf(w4.w, x4.w, acc);
            }
            u[q][k] = acc;
          }
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(u[q][k], va[q][k], acc);
          a[q] = ((jp < O) && !(p.mask0 && jp == 0)) ? acc : -CUDART_INF_F;
        }
        float m = a[0];
#pragma unroll
        for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[q]);
        m = bw_max(m);
        float ex[OPL], z = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          ex[q] = exp2f((a[q] - m) * LOG2E);
          z += ex[q];
        }
        const float inv = 1.0f / bw_sum(z);
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k) ta[q][k] = fmaf(ex[q] * inv, u[q][k], ta[q][k]);
      }
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) red[warp * E + (q * T + k) * 32 + lane] = ta[q][k];
      __syncthreads();
      for (int e = tid; e < E; e += NT) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += red[w * E + e];
        tr[r * E + e] = acc;
      }
      __syncthreads();
      for (int idx = tid; idx < OPL * 32; idx += NT) {
        const int ln = idx & 31, q = idx >> 5;
        float n2 = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const float t = tr[r * E + (q * T + k) * 32 + ln];
          n2 = fmaf(t, t, n2);
        }
        const float scale = (n2 / (1.0f + n2)) / sqrtf(n2 + 1e-7f);
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int e = (q * T + k) * 32 + ln;
          vacc[(r + 1) * E + e] = vacc[r * E + e] + tr[r * E + e] * scale;
        }
      }
      __syncthreads();
    }

    // ---------------- backward over the passes ----------------
    for (int e = tid; e < E; e += NT) gacc[e] = 0.f;  // G_R = 0
    __syncthreads();
    for (int r = R - 1; r >= 0; --r) {
      // g_v = [last] g_out + G_{r+1};  g_t = squash'(t_r) g_v
      for (int idx = tid; idx < OPL * 32; idx += NT) {
        const int ln = idx & 31, q = idx >> 5;
        float n2 = 0.f, dot = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int e = (q * T + k) * 32 + ln;
          const float t = tr[r * E + e];
          const float gv = (r == R - 1 ? gout[e] : 0.f) + gacc[e];
          n2 = fmaf(t, t, n2);
          dot = fmaf(gv, t, dot);
        }
        const float sq = sqrtf(n2 + 1e-7f);
        const float f = n2 / ((1.0f + n2) * sq);
        // f'(n2) = 1/((1+n2) sq) - f/(1+n2) - f/(2 (n2+eps))
        const float fp = 1.0f / ((1.0f + n2) * sq) - f / (1.0f + n2) - 0.5f * f / (n2 + 1e-7f);
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int e = (q * T + k) * 32 + ln;
          const float t = tr[r * E + e];
          const float gv = (r == R - 1 ? gout[e] : 0.f) + gacc[e];
          gt[e] = f * gv + 2.0f * fp * dot * t;
        }
      }
      __syncthreads();
      if (p.split) {
        for (int e = tid; e < E; e += NT) {
          const int ln = e & 31, qk = e >> 5, q = qk / T, k = qk % T;
          const int j = q * 32 + ln;
          if (j < O) {
            const size_t o = (((size_t)frame * R + r) * O + j) * T + k;
            p.gtT[o] = gt[e];
            p.vaT[o] = vacc[r * E + e];
          }
        }
      }
      float va[OPL][T], gtr[OPL][T], gv_acc[OPL][T];
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) {
          va[q][k] = vacc[r * E + (q * T + k) * 32 + lane];
          gtr[q][k] = gt[(q * T + k) * 32 + lane];
          gv_acc[q][k] = 0.f;
        }
      for (int i = warp; i < I; i += NW) {
        float u[OPL][T], a[OPL];
        const float4* xrow = reinterpret_cast<const float4*>(xs) + (size_t)i * T4;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
#pragma unroll
          for (int k = 0; k < T; ++k) {
            float acc = Bp[(size_t)(i * T + k) * OP + jp];
#pragma unroll
            for (int c = 0; c < T4; ++c) {
              const float4 w4 = Wp[((size_t)(i * T + k) * T4 + c) * OP + jp];
              const float4 x4 = xrow[c];
              acc = fmaf(w4.x, x4.x, acc);
              acc = fmaf(w4.y, x4.y, acc);
              acc = fmaf(w4.z, x4.z, acc);
              acc = fmaf(w4.w, x4.w, acc);
            }
            u[q][k] = acc;
          }
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(u[q][k], va[q][k], acc);
          a[q] = ((jp < O) && !(p.mask0 && jp == 0)) ? acc : -CUDART_INF_F;
        }
        float m = a[0];
#pragma unroll
        for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[q]);
        m = bw_max(m);
        float c[OPL], z = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          c[q] = exp2f((a[q] - m) * LOG2E);
          z += c[q];
        }
        const float inv = 1.0f / bw_sum(z);
        float gc[OPL], cg = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          c[q] *= inv;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(gtr[q][k], u[q][k], acc);
          gc[q] = acc;
          cg = fmaf(c[q], acc, cg);
        }
        cg = bw_sum(cg);
        // per-lane g_u, dbias, dW; dx needs a sum over output capsules (lanes)
        float dx[T];
#pragma unroll
        for (int l = 0; l < T; ++l) dx[l] = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
          const float ga = c[q] * (gc[q] - cg);
          if (p.split) {
            const size_t o = (((size_t)frame * R + r) * I + i) * OP + jp;
            p.cbuf[o] = c[q];
            p.gabuf[o] = ga;
          }
#pragma unroll
          for (int k = 0; k < T; ++k) {
            const float gu = c[q] * gtr[q][k] + ga * va[q][k];
            gv_acc[q][k] = fmaf(ga, u[q][k], gv_acc[q][k]);
            if (jp < O && k < D) {
              float* dWrow = p.dW + (((size_t)i * O + jp) * D + k) * p.d;
              if (!p.split) atomicAdd(p.dbias + ((size_t)i * O + jp) * D + k, gu);
#pragma unroll
              for (int c4 = 0; c4 < T4; ++c4) {
                const float4 x4 = xrow[c4];
                const float4 w4 = Wp[((size_t)(i * T + k) * T4 + c4) * OP + jp];
                const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int li = 0; li < 4; ++li) {
                  const int l = c4 * 4 + li;
                  if (l < p.d) {
                    if (!p.split) atomicAdd(dWrow + l, gu * xv[li]);
                    dx[l] = fmaf(gu, wv[li], dx[l]);
                  }
                }
              }
            }
          }
        }
        if (p.split) {
          // sum dx over the lanes with a transposing butterfly: 31 shuffles, lane l ends up with
          // the total of dx[l]
          float v32[32];
#pragma unroll
          for (int l = 0; l < 32; ++l) v32[l] = l < T ? dx[l] : 0.f;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int m = 0; m < o; ++m) {
              const float send = up ? v32[m] : v32[m + o];
              const float keep = up ? v32[m + o] : v32[m];
              v32[m] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
          }
          if (lane < T) {
            float* dst = p.dxw + ((size_t)frame * I + i) * T + lane;
            *dst = (r == R - 1) ? v32[0] : *dst + v32[0];
          }
        } else if (p.d_emb != nullptr) {
          const int w = i / p.H, hc = i - w * p.H;
          const int src = sf - p.lpad + w;
#pragma unroll
          for (int l = 0; l < T; ++l) {
            const float s = bw_sum(dx[l]);
            if (lane == 0 && l < p.d && src >= 0 && src < p.S)
              atomicAdd(p.d_emb + (((long long)b * p.S + src) * p.H + hc) * p.d + l, s);
          }
        }
      }
      // G_r = G_{r+1} + sum_i g_a u
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) red[warp * E + (q * T + k) * 32 + lane] = gv_acc[q][k];
      __syncthreads();
      for (int e = tid; e < E; e += NT) {
        float acc = gacc[e];
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += red[w * E + e];
        gacc[e] = acc;
      }
      __syncthreads();
    }
    // gacc now holds dL/dVacc_0 = the BPTT carry into the previous frame (SDR)
  }
}

// ---------------------------------------------------------------------------------------
// split mode, phase B: dW[i,j,k,l] = sum_f sum_r (c g_t[k] + g_a Vacc[k]) x[l],
// dbias[i,j,k] = sum_f sum_r (c g_t[k] + g_a Vacc[k]).  One CTA per (i, j), one thread per (k, l),
// the sum over all frames stays in a register: no atomics.
// ---------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(T* T) dw_from_saved_kernel(const BwdParams p) {
  constexpr int FT = 32;  // frames per shared-memory tile
  extern __shared__ float sm[];
  const int R = p.iters, O = p.O, I = p.I;
  const int OP = p.OP;
  float* gts = sm;                 // [FT][R][T]
  float* vas = gts + FT * R * T;   // [FT][R][T]
  float* xs = vas + FT * R * T;    // [FT][T]
  float* cs = xs + FT * T;         // [FT][R]
  float* gas = cs + FT * R;        // [FT][R]
  const int j = blockIdx.x, i = blockIdx.y;
  const int tid = threadIdx.x, k = tid / T, l = tid % T;
  const int w = i / p.H, hc = i - w * p.H;
  const long long frames = (long long)p.B * p.S;
  float acc = 0.f, accb = 0.f;
  for (long long f0 = 0; f0 < frames; f0 += FT) {
    const int nf = (int)((frames - f0) < FT ? (frames - f0) : FT);
    for (int e = tid; e < nf * R * T; e += T * T) {
      const int kk = e % T, fr = e / T;  // fr = ft*R + r
      const size_t o = (((size_t)f0 * R + fr) * O + j) * T + kk;
      gts[e] = p.gtT[o];
      vas[e] = p.vaT[o];
    }
    for (int e = tid; e < nf * T; e += T * T) {
      const int ll = e % T, ft = e / T;
      const long long f = f0 + ft;
      const int b = (int)(f / p.S), s = (int)(f % p.S);
      const int src = s - p.lpad + w;
      float v = 0.f;
      if (ll < p.d && src >= 0 && src < p.S) v = p.emb[(((long long)b * p.S + src) * p.H + hc) * p.d + ll];
      xs[e] = v;
    }
    for (int e = tid; e < nf * R; e += T * T) {
      const size_t o = (((size_t)f0 * R + e) * I + i) * OP + j;
      cs[e] = p.cbuf[o];
      gas[e] = p.gabuf[o];
    }
    __syncthreads();
    for (int fr = 0; fr < nf * R; ++fr) {
      const float gu = fmaf(cs[fr], gts[fr * T + k], gas[fr] * vas[fr * T + k]);
      acc = fmaf(gu, xs[(fr / R) * T + l], acc);
      accb += gu;
    }
    __syncthreads();
  }
  if (k < p.D && l < p.d) p.dW[(((size_t)i * O + j) * p.D + k) * p.d + l] += acc;
  if (k < p.D && l == 0) p.dbias[((size_t)i * O + j) * p.D + k] += accb;
}

void launch_dw_from_saved(const BwdParams& p, int T, cudaStream_t stream) {
  const int FT = 32, R = p.iters;
  const size_t smem = sizeof(float) * ((size_t)2 * FT * R * T + FT * T + 2 * FT * R);
  dim3 grid(p.O, p.I);
  if (T == 8) dw_from_saved_kernel<8><<<grid, 64, smem, stream>>>(p);
  else if (T == 16) dw_from_saved_kernel<16><<<grid, 256, smem, stream>>>(p);
  else if (T == 20) dw_from_saved_kernel<20><<<grid, 400, smem, stream>>>(p);
  else dw_from_saved_kernel<32><<<grid, 1024, smem, stream>>>(p);
}

// d_emb[b,s',h,l] += sum_w dxw[(b, s'+lpad-w), w*H+h, l]   (fold the window back)
__global__ void fold_dx_kernel(const BwdParams p, int T) {
  const long long n = (long long)p.B * p.S * p.H * p.d;
  const int window = p.I / p.H;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(e % p.d);
    long long q = e / p.d;
    const int hc = (int)(q % p.H);
    q /= p.H;
    const int s = (int)(q % p.S);
    const int b = (int)(q / p.S);
    float acc = 0.f;
    for (int w = 0; w < window; ++w) {
      const int sf = s + p.lpad - w;
      if (sf >= 0 && sf < p.S)
        acc += p.dxw[(((long long)b * p.S + sf) * p.I + w * p.H + hc) * T + l];
    }
    p.d_emb[e] += acc;
  }
}

void launch_fold_dx(const BwdParams& p, int T, cudaStream_t stream) {
  const long long n = (long long)p.B * p.S * p.H * p.d;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  fold_dx_kernel<<<blocks, 256, 0, stream>>>(p, T);
}

template <int T, int OPL>
static cudaError_t launch_bwd_variant(const BwdParams& p, int nchains, cudaStream_t stream) {
  constexpr int NW = 8;
  const size_t E = (size_t)OPL * T * 32;
  const size_t smem = sizeof(float) * ((size_t)p.I * T + NW * E + (2 * BW_MAX_ITERS + 1) * E + 3 * E);
  auto kern = route_layer_bwd_kernel<T, OPL, NW>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  kern<<<nchains, NW * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

#define SRF_BWD(T_, OPL_) \
  if (T == T_ && OPL == OPL_) return launch_bwd_variant<T_, OPL_>(p, nchains, stream);

cudaError_t launch_route_layer_bwd(const BwdParams& p, int T, int OPL, int nchains,
                                   cudaStream_t stream) {
  if (p.iters > BW_MAX_ITERS) return cudaErrorInvalidValue;
  SRF_BWD(8, 1)
  SRF_BWD(8, 2)
  SRF_BWD(8, 4)
  SRF_BWD(16, 1)
  SRF_BWD(16, 2)
  SRF_BWD(20, 1)
  SRF_BWD(20, 2)
  SRF_BWD(32, 1)
  return cudaErrorInvalidValue;
}

}  // namespace srf
