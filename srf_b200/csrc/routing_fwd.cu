// routing_fwd.cu -- fused routing-layer forward for sm_100a (FP32 CUDA-core u_hat variant).
//
// One launch = one routing layer of the reference's stack
// (tfsr/model/sequence_router_naive.py:145-193): window gather, u_hat = W.x + b, SDR or DR
// routing iterations, LayerNorm(O*D) + dropout mask, optional head ln_o(length(.)).
// u_hat, the routing logits and the coupling coefficients never touch HBM.
//
// Work decomposition
//   chain      = one utterance (SDR: the frame loop runs inside the kernel and carries v)
//                or one routing frame (DR).
//   cluster    = C CTAs that share F chains and split the input capsules i between them;
//                the partial weighted sums are all-reduced through distributed shared memory.
//   warp       = a strided subset of the CTA's input capsules.
//   lane       = output capsule j (j = q*32 + lane, q < OPL); a thread keeps u_hat[i,j,:],
//                the accumulated outputs Vacc[j,:] and the running sum t[j,:] in registers.
//   softmax over output capsules = warp shuffles; reductions over k are in-thread.
//
// Routing logits are never stored: b_r[i,j] = u_hat[i,j,:] . (v_0 + ... + v_{r-1})[j,:]
// (v_0 = previous frame's output for SDR, 0 for DR) by linearity of the agreement update
// (naive:205, :223, :240), so only Vacc = sum of the squashed outputs is kept.

#include "routing_kernels.h"

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math_constants.h>

namespace cg = cooperative_groups;

namespace srf {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------
// weight packing: canonical W[I,O,D,d], bias[I,O,D]  ->  lane-major padded layout
//   Wp float4[((i*T + k)*T/4 + c)*OP + j] = W[i,j,k,4c..4c+3]   (zero padded to T, OP)
//   Bp float [(i*T + k)*OP + j]          = bias[i,j,k]
// so that a warp's load of one (i,k,c) row is one coalesced 512-byte request.
// ---------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                    float* __restrict__ Wp, float* __restrict__ Bp, int I, int O,
                                    int D, int d, int T, int OP) {
  const long long nW = (long long)I * T * T * OP;
  const long long nB = (long long)I * T * OP;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nW + nB; e += stride) {
    if (e < nW) {
      const int r = (int)(e & 3);
      long long q = e >> 2;
      const int j = (int)(q % OP);
      q /= OP;
      const int c = (int)(q % (T / 4));
      q /= (T / 4);
      const int k = (int)(q % T);
      const int i = (int)(q / T);
      const int l = 4 * c + r;
      float v = 0.f;
      if (j < O && k < D && l < d) v = W[(((long long)i * O + j) * D + k) * d + l];
      Wp[e] = v;
    } else {
      long long q = e - nW;
      const int j = (int)(q % OP);
      q /= OP;
      const int k = (int)(q % T);
      const int i = (int)(q / T);
      float v = 0.f;
      if (j < O && k < D) v = bias[((long long)i * O + j) * D + k];
      Bp[e - nW] = v;
    }
  }
}

void launch_pack_weights(const float* W, const float* bias, float* Wp, float* Bp, int I, int O,
                         int D, int d, int T, int OP, cudaStream_t stream) {
  const long long n = (long long)I * T * OP * (T + 1);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_kernel<<<blocks, 256, 0, stream>>>(W, bias, Wp, Bp, I, O, D, d, T, OP);
}

// ---------------------------------------------------------------------------------------
// the fused layer kernel
// ---------------------------------------------------------------------------------------
// UM = 0: u_hat computed here in FP32 from emb and the packed weights;
// UM = 1 / 2: u_hat was materialised by the tcgen05 GEMM (uhat_gemm.cu) as bf16 / fp32 in the
//             streaming layout [frame pair][i][q*(T/4)+k4][lane*4+k%4][2] and is only loaded
//             here (F must be 2 = one frame pair per warp).
template <int T, int OPL, int F, int NW, int UM>
__global__ void __launch_bounds__(NW * 32) route_layer_kernel(const RouteParams p) {
  static_assert(UM == 0 || F == 2, "u_hat streaming mode works on frame pairs");
  constexpr int OP = 32 * OPL;
  constexpr int T4 = T / 4;
  constexpr int NT = NW * 32;
  constexpr int E = F * OPL * T * 32;  // one t tile; index ((f*OPL+q)*T+k)*32+lane
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C;
  int rank = 0;
  if (C > 1) rank = (int)cg::this_cluster().block_rank();
  const int group = blockIdx.x / C;
  const int i_lo = rank * p.Ic;
  const int i_hi = min(p.I, i_lo + p.Ic);
  const int ni = max(0, i_hi - i_lo);

  float* xs = smem;                // [F][Ic][T]   window-gathered inputs of this CTA's i range
  float* red = xs + (UM == 0 ? F * p.Ic * T : 0);  // [NW][E] per-warp partial t, later the total
  float* tsum = red + NW * E;      // [2][E]       CTA partial for the cluster exchange
  float* vacc = tsum + 2 * E;      // [E]          Vacc = sum of squashed outputs so far
  float* vlast = vacc + E;         // [E]          last squashed output

  const float4* __restrict__ Wp = reinterpret_cast<const float4*>(p.Wp);
  const float* __restrict__ Bp = p.Bp;
  const int O = p.O, D = p.D;
  int par = 0;

  for (int s = 0; s < p.nsteps; ++s) {
    // ---- (1) window gather (naive:150-151) of this CTA's input capsules into smem --------
    for (int idx = tid; UM == 0 && idx < F * ni * T; idx += NT) {
      const int l = idx % T;
      const int ii = (idx / T) % ni;
      const int f = idx / (T * ni);
      const int chain = group * F + f;
      float v = 0.f;
      if (chain < p.nchains && l < p.d) {
        const int b = p.sdr ? chain : chain / p.S;
        const int sf = p.sdr ? s : chain % p.S;
        const int i = i_lo + ii;
        const int w = i / p.H, hcap = i - w * p.H;
        const int src = sf - p.lpad + w;
        if (src >= 0 && src < p.S)
          v = __ldg(p.emb + (((long long)b * p.S + src) * p.H + hcap) * p.d + l);
      }
      xs[(f * p.Ic + ii) * T + l] = v;
    }
    // ---- (2) Vacc: previous frame's output (SDR, naive:164,167) or zero (DR, naive:172) ---
    for (int e = tid; e < E; e += NT) vacc[e] = (p.sdr && s > 0) ? vlast[e] : 0.f;
    __syncthreads();

    for (int pass = 0; pass < p.iters; ++pass) {
      float va[F][OPL][T], ta[F][OPL][T];
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k) {
            va[f][q][k] = vacc[((f * OPL + q) * T + k) * 32 + lane];
            ta[f][q][k] = 0.f;
          }

      // streaming mode: frame pair index of this group at this step, and a one-capsule-ahead
      // register prefetch of the materialised u_hat
      constexpr int RAWN = (UM == 1) ? OPL * T4 : (UM == 2 ? 2 * OPL * T4 : 1);
      uint4 raw[RAWN];
      const long long gg = p.sdr ? ((long long)s * p.halfB + group) : (long long)group;
      const size_t ustride_i = (size_t)OPL * T4 * 128 * 2;  // elements per (pair, i)
      auto load_raw = [&](int i, uint4(&dst)[RAWN]) {
        if (UM == 1) {
          const uint4* src = reinterpret_cast<const uint4*>(
              reinterpret_cast<const uint16_t*>(p.u) + ((size_t)gg * p.I + i) * ustride_i);
#pragma unroll
          for (int m = 0; m < OPL * T4; ++m) dst[m] = __ldg(src + m * 32 + lane);
        } else if (UM == 2) {
          const uint4* src = reinterpret_cast<const uint4*>(
              reinterpret_cast<const float*>(p.u) + ((size_t)gg * p.I + i) * ustride_i);
#pragma unroll
          for (int m = 0; m < OPL * T4; ++m) {
            dst[2 * m] = __ldg(src + (m * 32 + lane) * 2);
            dst[2 * m + 1] = __ldg(src + (m * 32 + lane) * 2 + 1);
          }
        }
      };
      if (UM != 0 && i_lo + warp < i_hi) load_raw(i_lo + warp, raw);

      for (int i = i_lo + warp; i < i_hi; i += NW) {
        float u[F][OPL][T];
        float a[F][OPL];
        const float4* xrow = reinterpret_cast<const float4*>(xs) + (size_t)(i - i_lo) * T4;
        if (UM != 0) {
          // unpack the prefetched registers: element (k_in, f) of chunk m=(q,k4) sits at 2*k_in+f
#pragma unroll
          for (int q = 0; q < OPL; ++q)
#pragma unroll
            for (int k4 = 0; k4 < T4; ++k4) {
              const int m = q * T4 + k4;
              if (UM == 1) {
                const uint32_t w[4] = {raw[m].x, raw[m].y, raw[m].z, raw[m].w};
#pragma unroll
                for (int kin = 0; kin < 4; ++kin) {
                  u[0][q][k4 * 4 + kin] = __uint_as_float(w[kin] << 16);
                  u[F - 1][q][k4 * 4 + kin] = __uint_as_float(w[kin] & 0xffff0000u);
                }
              } else {
                const uint32_t w[8] = {raw[(2 * m) % RAWN].x, raw[(2 * m) % RAWN].y,
                                       raw[(2 * m) % RAWN].z, raw[(2 * m) % RAWN].w,
                                       raw[(2 * m + 1) % RAWN].x, raw[(2 * m + 1) % RAWN].y,
                                       raw[(2 * m + 1) % RAWN].z, raw[(2 * m + 1) % RAWN].w};
#pragma unroll
                for (int kin = 0; kin < 4; ++kin) {
                  u[0][q][k4 * 4 + kin] = __uint_as_float(w[2 * kin]);
                  u[F - 1][q][k4 * 4 + kin] = __uint_as_float(w[2 * kin + 1]);
                }
              }
            }
          if (i + NW < i_hi) load_raw(i + NW, raw);
        }
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
#pragma unroll
          for (int k = 0; UM == 0 && k < T; ++k) {
            float4 w4[T4];
#pragma unroll
            for (int c = 0; c < T4; ++c)
              w4[c] = __ldg(Wp + ((size_t)(i * T + k) * T4 + c) * OP + jp);
            const float bk = __ldg(Bp + (size_t)(i * T + k) * OP + jp);
#pragma unroll
            for (int f = 0; f < F; ++f) {
              float acc = bk;
#pragma unroll
              for (int c = 0; c < T4; ++c) {
                const float4 x4 = xrow[(size_t)f * p.Ic * T4 + c];
                acc = fmaf(w4[c].x, x4.x, acc);
                acc = fmaf(w4[c].y, x4.y, acc);
                acc = fmaf(w4[c].z, x4.z, acc);
                acc = fmaf(w4[c].w, x4.w, acc);
              }
              u[f][q][k] = acc;
            }
          }
          // agreement with the accumulated outputs (naive:205 / :223 / :240)
          const bool valid = (jp < O) && !(p.mask0 && jp == 0);
#pragma unroll
          for (int f = 0; f < F; ++f) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < T; ++k) acc = fmaf(u[f][q][k], va[f][q][k], acc);
            a[f][q] = valid ? acc : -CUDART_INF_F;
          }
        }
        // coupling softmax over output capsules (naive:202 / :225 / :241) + weighted sum
#pragma unroll
        for (int f = 0; f < F; ++f) {
          float m = a[f][0];
#pragma unroll
          for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[f][q]);
          m = warp_max(m);
          float ex[OPL];
          float z = 0.f;
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            ex[q] = exp2f((a[f][q] - m) * LOG2E);
            z += ex[q];
          }
          z = warp_sum(z);
          const float inv = 1.0f / z;
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            const float c = ex[q] * inv;
#pragma unroll
            for (int k = 0; k < T; ++k) ta[f][q][k] = fmaf(c, u[f][q][k], ta[f][q][k]);
          }
        }
      }

      // ---- reduce t over warps, then over the cluster's CTAs (DSMEM) ---------------------
#pragma unroll
      for (int f = 0; f < F; ++f)
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k)
            red[warp * E + ((f * OPL + q) * T + k) * 32 + lane] = ta[f][q][k];
      __syncthreads();
      for (int e = tid; e < E; e += NT) {
        float acc = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) acc += red[w * E + e];
        tsum[par * E + e] = acc;
      }
      if (C > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        for (int e = tid; e < E; e += NT) {
          float acc = 0.f;
          for (int r = 0; r < C; ++r) acc += cluster.map_shared_rank(tsum, r)[par * E + e];
          red[e] = acc;
        }
      } else {
        __syncthreads();
        for (int e = tid; e < E; e += NT) red[e] = tsum[par * E + e];
      }
      par ^= 1;
      __syncthreads();
      // ---- squash (naive:248-253) and Vacc update ----------------------------------------
      for (int idx = tid; idx < F * OPL * 32; idx += NT) {
        const int ln = idx & 31;
        const int fq = idx >> 5;
        float n2 = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const float t = red[(fq * T + k) * 32 + ln];
          n2 = fmaf(t, t, n2);
        }
        const float scale = (n2 / (1.0f + n2)) / sqrtf(n2 + 1e-7f);
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int e = (fq * T + k) * 32 + ln;
          const float v = red[e] * scale;
          vlast[e] = v;
          vacc[e] += v;
        }
      }
      __syncthreads();
    }  // passes

    // ---- (3) epilogue for this step: LayerNorm(O*D) (+dropout mask), optional head --------
    if (warp < F && rank == 0) {
      const int f = warp;
      const int chain = group * F + f;
      if (chain < p.nchains) {
        long long frame = p.sdr ? ((long long)chain * p.S + s) : chain;
        bool frame_ok = true;
        if (UM != 0 && !p.sdr) {  // DR on frame pairs: group -> (s, utterance pair)
          const int ss = group / p.halfB, bb = 2 * (group % p.halfB) + f;
          frame = (long long)bb * p.S + ss;
          frame_ok = bb < p.B;
        }
        if (UM != 0 && p.sdr) frame_ok = chain < p.B;
        const float* vf = vlast + (size_t)f * OPL * T * 32;
        float mean = 0.f, rstd = 1.f;
        const bool do_ln = p.ln_gamma != nullptr;
        if (do_ln) {
          float sum = 0.f;
          for (int q = 0; q < OPL; ++q) {
            const int j = q * 32 + lane;
            if (j < O)
              for (int k = 0; k < D; ++k) sum += vf[(q * T + k) * 32 + lane];
          }
          mean = warp_sum(sum) / (float)(O * D);
          float sq = 0.f;
          for (int q = 0; q < OPL; ++q) {
            const int j = q * 32 + lane;
            if (j < O)
              for (int k = 0; k < D; ++k) {
                const float dv = vf[(q * T + k) * 32 + lane] - mean;
                sq = fmaf(dv, dv, sq);
              }
          }
          const float var = warp_sum(sq) / (float)(O * D);
          rstd = 1.0f / sqrtf(var + p.ln_eps);
        }
        float len[OPL];
        for (int q = 0; q < OPL; ++q) {
          const int j = q * 32 + lane;
          float l2 = 0.f;
          if (j < O) {
            for (int k = 0; k < D; ++k) {
              float y = vf[(q * T + k) * 32 + lane];
              if (p.out_raw && frame_ok) p.out_raw[(frame * O + j) * D + k] = y;
              if (do_ln) y = (y - mean) * rstd * __ldg(p.ln_gamma + j * D + k) + __ldg(p.ln_beta + j * D + k);
              if (p.dropout_mask && frame_ok) y *= __ldg(p.dropout_mask + (frame * O + j) * D + k);
              if (p.out_caps && frame_ok) p.out_caps[(frame * O + j) * D + k] = y;
              l2 = fmaf(y, y, l2);
            }
          }
          len[q] = sqrtf(l2 + p.length_eps);  // naive:256-258
        }
        if (p.head_gamma != nullptr) {  // ln_output over the class capsule lengths (naive:193)
          float sum = 0.f;
          for (int q = 0; q < OPL; ++q)
            if (q * 32 + lane < O) sum += len[q];
          const float hm = warp_sum(sum) / (float)O;
          float sq = 0.f;
          for (int q = 0; q < OPL; ++q)
            if (q * 32 + lane < O) {
              const float dv = len[q] - hm;
              sq = fmaf(dv, dv, sq);
            }
          const float hr = 1.0f / sqrtf(warp_sum(sq) / (float)O + p.ln_eps);
          for (int q = 0; q < OPL; ++q) {
            const int j = q * 32 + lane;
            if (j < O && frame_ok)
              p.out_logits[frame * O + j] =
                  (len[q] - hm) * hr * __ldg(p.head_gamma + j) + __ldg(p.head_beta + j);
          }
        }
      }
    }
    __syncthreads();
  }  // steps

  // a CTA must not exit while peers may still read its shared memory
  if (C > 1) cg::this_cluster().sync();
}

// ---------------------------------------------------------------------------------------
// host-side dispatch
// ---------------------------------------------------------------------------------------
template <int T, int OPL, int F, int NW, int UM>
static cudaError_t launch_variant(const RouteParams& p, int groups, size_t smem_bytes,
                                  cudaStream_t stream) {
  auto kern = route_layer_kernel<T, OPL, F, NW, UM>;
  cudaError_t err =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * p.C));
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)p.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

size_t route_layer_smem_bytes(int T, int OPL, int F, int NW, int Ic, int um) {
  const size_t E = (size_t)F * OPL * T * 32;
  return sizeof(float) * ((um == 0 ? (size_t)F * Ic * T : 0) + (size_t)NW * E + 4 * E);
}

#define SRF_VARIANT(T_, OPL_, F_)                                                    \
  if (T == T_ && OPL == OPL_ && F == F_)                                             \
    return launch_variant<T_, OPL_, F_, SRF_NW, 0>(p, groups, smem_bytes, stream);
#define SRF_VARIANT_U(T_, OPL_)                                                      \
  if (T == T_ && OPL == OPL_ && F == 2 && um == 1)                                   \
    return launch_variant<T_, OPL_, 2, SRF_NW, 1>(p, groups, smem_bytes, stream);    \
  if (T == T_ && OPL == OPL_ && F == 2 && um == 2)                                   \
    return launch_variant<T_, OPL_, 2, SRF_NW, 2>(p, groups, smem_bytes, stream);

cudaError_t launch_route_layer(const RouteParams& p, int T, int OPL, int F, int groups,
                               size_t smem_bytes, int um, cudaStream_t stream) {
  if (um != 0) {
    SRF_VARIANT_U(8, 1)
    SRF_VARIANT_U(8, 2)
    SRF_VARIANT_U(8, 4)
    SRF_VARIANT_U(16, 1)
    SRF_VARIANT_U(16, 2)
    SRF_VARIANT_U(20, 1)
    SRF_VARIANT_U(20, 2)
    SRF_VARIANT_U(32, 1)
    return cudaErrorInvalidValue;
  }
  SRF_VARIANT(8, 1, 4)
  SRF_VARIANT(8, 1, 2)
  SRF_VARIANT(8, 1, 1)
  SRF_VARIANT(8, 2, 2)
  SRF_VARIANT(8, 2, 1)
  SRF_VARIANT(8, 4, 1)
  SRF_VARIANT(16, 1, 2)
  SRF_VARIANT(16, 1, 1)
  SRF_VARIANT(16, 2, 1)
  SRF_VARIANT(16, 4, 1)
  SRF_VARIANT(20, 1, 2)
  SRF_VARIANT(20, 1, 1)
  SRF_VARIANT(20, 2, 1)
  SRF_VARIANT(20, 4, 1)
  SRF_VARIANT(32, 1, 1)
  SRF_VARIANT(32, 2, 1)
  SRF_VARIANT(32, 4, 1)
  return cudaErrorInvalidValue;
}

int route_layer_max_F(int T, int OPL) {
  if (T == 8) return OPL == 1 ? 4 : (OPL == 2 ? 2 : 1);
  if (T == 16 || T == 20) return OPL == 1 ? 2 : 1;
  return 1;
}

}  // namespace srf
