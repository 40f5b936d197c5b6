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

#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cuda.h>

#include "routing_kernels.h"
#include "sm100_ptx.cuh"

namespace cg = cooperative_groups;

namespace srf {

namespace {
__device__ __forceinline__ uint32_t bw_map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
// remote store of 4 floats into a peer CTA's shared memory that completes (by byte count) on that
// peer's mbarrier: data and signal in one instruction (the exchange of routing_stream.cu)
__device__ __forceinline__ void bw_st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          remote_addr),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
      : "memory");
}
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
// tensor modes: both softmax reductions as single REDUX instructions (the max on an
// order-preserving integer image of the floats, the sum on an unsigned Q26 image of exp(a - max)
// in (0, 1]; same construction as routing_stream.cu)
__device__ __forceinline__ float bw_max_redux(float v) {
  int mi = __float_as_int(v);
  mi ^= (mi >> 31) & 0x7fffffff;
  mi = __reduce_max_sync(0xffffffffu, mi);
  mi ^= (mi >> 31) & 0x7fffffff;
  return __int_as_float(mi);
}
template <int OPL>
__device__ __forceinline__ float bw_sum_redux(float z) {
  constexpr float QS = (float)(1 << 26) / (float)OPL;
  const unsigned zi = __reduce_add_sync(0xffffffffu, __float2uint_rn(z * QS));
  return (float)zi * (1.0f / QS);
}
constexpr int BW_MAX_ITERS = 8;
#define SRF_BWD_NW_BF16 12
#define SRF_BWD_NW_F32 12
}  // namespace

// ---------------------------------------------------------------------------------------
// K_b1: head + dropout + LayerNorm backward.  grid = frames / 8, block = 256 (warp per frame)
// ---------------------------------------------------------------------------------------
__global__ void ln_head_bwd_kernel(const BwdParams p) {
  extern __shared__ float lsm[];
  const int lane = threadIdx.x & 31;
  const int O = p.O, D = p.D;
  const int n = O * D;
  // parameter gradients are summed per CTA in shared memory over its strip of frames and leave
  // with one global atomic per element and CTA
  float* sg = lsm;          // [n] dgamma
  float* sb = sg + n;       // [n] dbeta
  float* shg = sb + n;      // [O] dhead_gamma
  float* shb = shg + O;     // [O] dhead_beta
  for (int e = threadIdx.x; e < 2 * n + 2 * O; e += blockDim.x) lsm[e] = 0.f;
  __syncthreads();
  const long long frames = (long long)p.B * p.S;
  const int wpb = blockDim.x >> 5;
  for (long long frame = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); frame < frames;
       frame += (long long)gridDim.x * wpb) {
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
      atomicAdd(shg + j, dl * zhat);
      atomicAdd(shb + j, dl);
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
        atomicAdd(sg + e, g * xhat);
        atomicAdd(sb + e, g);
        const float dxh = g * p.ln_gamma[e];
        m1 += dxh;
        m2 = fmaf(dxh, xhat, m2);
      } else {
        p.d_raw[frame * n + e] = g;
      }
    }
  }
  if (!do_ln) continue;
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
  }  // frames of this warp
  __syncthreads();
  if (p.ln_gamma != nullptr)
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
      atomicAdd(p.dgamma + e, sg[e]);
      atomicAdd(p.dbeta + e, sb[e]);
    }
  if (p.d_logits != nullptr)
    for (int j = threadIdx.x; j < O; j += blockDim.x) {
      atomicAdd(p.dhead_gamma + j, shg[j]);
      atomicAdd(p.dhead_beta + j, shb[j]);
    }
}

void launch_ln_head_bwd(const BwdParams& p, cudaStream_t stream) {
  const long long frames = (long long)p.B * p.S;
  const int warps = 8;
  long long blocks = (frames + warps - 1) / warps;
  if (blocks > 148 * 4) blocks = 148 * 4;
  const size_t smem = sizeof(float) * (2 * (size_t)p.O * p.D + 2 * (size_t)p.O);
  ln_head_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, stream>>>(p);
}

// ---------------------------------------------------------------------------------------
// K_b2: routing backward (BPTT sweep).  One CTA per chain; warps stride over the input
// capsules; lane = output capsule (j = q*32 + lane).
//   UM = 0: u_hat recomputed in FP32 from the packed weights (exact mode);
//   UM = 1/2: u_hat streamed from the tcgen05 GEMM's output (bf16 / fp32 storage, the layout of
//             uhat_gemm.cu), one-capsule-ahead register prefetch.
//   SPLIT: store c, g_a, g_t, Vacc for phase B (dwdx_from_saved_kernel) instead of forming
//          dW / dbias / dx with atomics in this kernel.
// ---------------------------------------------------------------------------------------
template <int T, int OPL, int NW, int UM, bool SPLIT>
__global__ void __launch_bounds__(NW * 32) route_layer_bwd_kernel(const BwdParams p) {
  constexpr int OP = 32 * OPL;
  constexpr int T4 = T / 4;
  constexpr int NT = NW * 32;
  constexpr int E = OPL * T * 32;  // (q*T+k)*32+lane
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr int RAWN = (UM == 1) ? OPL * T4 : (UM == 2 ? 2 * OPL * T4 : 1);
  static_assert(UM == 0 || SPLIT, "streamed u_hat needs split mode");

  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C;
  int rank = 0;
  if (C > 1) rank = (int)cg::this_cluster().block_rank();
  const int chain = blockIdx.x / C;
  const int I = p.I, O = p.O, D = p.D, R = p.iters;
  const int i_lo = rank * p.Ic;
  const int i_hi = min(I, i_lo + p.Ic);
  constexpr int PER = (E + NT - 1) / NT;
  uint64_t* xfull = reinterpret_cast<uint64_t*>(smem);  // [2] exchange barriers (16 bytes)
  float* xs = smem + 4;             // [Ic][T] window-gathered input of the frame (UM == 0)
  float* red = xs + (UM == 0 ? (size_t)p.Ic * T : 0);  // [NW][E]
  float* xbuf = red + NW * E;       // [2][C][E] CTA partials of the cluster exchange (own + pushed)
  float* tr = xbuf + (size_t)2 * C * E;  // [R][E]   t_r
  float* vacc = tr + BW_MAX_ITERS * E;  // [R+1][E] Vacc_0 .. Vacc_R
  float* gout = vacc + (BW_MAX_ITERS + 1) * E;  // [E] dL/dv of this frame (incl. BPTT carry)
  float* gt = gout + E;             // [E] g_t of the current pass
  float* gacc = gt + E;             // [E] G_r
  float* ndb = gacc + E;            // [E] next frame's dL/dv_raw (cp.async landing zone)
  float* nvb = ndb + E;             // [E] next frame's carried output

  const float4* __restrict__ Wp = reinterpret_cast<const float4*>(p.Wp);
  const float* __restrict__ Bp = p.Bp;

  for (int e = tid; e < E; e += NT) gacc[e] = 0.f;  // BPTT carry into the last frame is zero
  if (tid == 0) {
    ptx::mbar_init(&xfull[0], 1 + NW);  // expect_tx arrive + one arrive per warp (own copy stored)
    ptx::mbar_init(&xfull[1], 1 + NW);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (C > 1) cg::this_cluster().sync();  // peers' barriers are initialised before anybody pushes
  int par = 0;
  uint32_t nexch = 0;
  // Sum the per-warp partials in `red` over the warps and over the cluster's CTAs.  Every CTA
  // pushes its partial into every peer's xbuf[par][rank] with st.async (completes on the peer's
  // mbarrier: no cluster barrier, no remote loads) and sums the C partials in rank order, so all
  // CTAs hold bit-identical totals.  Call after a __syncthreads(); `sink(e4, total)` stores.
  auto exchange = [&](auto sink) {
    const float4* red4 = reinterpret_cast<const float4*>(red);
    if (C > 1) {
      uint64_t* xf = &xfull[par];
      if (tid == 0) ptx::mbar_arrive_expect_tx(xf, (uint32_t)(C - 1) * E * 4);
      float* mine = xbuf + ((size_t)par * C + rank) * E;
      const uint32_t mine_a = ptx::smem_u32(mine), bar_a = ptx::smem_u32(xf);
      for (int e4 = tid; e4 < E / 4; e4 += NT) {
        float4 acc = red4[e4];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          const float4 x = red4[(size_t)w * (E / 4) + e4];
          acc.x += x.x;
          acc.y += x.y;
          acc.z += x.z;
          acc.w += x.w;
        }
        reinterpret_cast<float4*>(mine)[e4] = acc;
        for (int r2 = 1; r2 < C; ++r2) {
          const int peer = (rank + r2) & (C - 1);  // C is a power of two
          bw_st_async_v4(bw_map_to_rank(mine_a + e4 * 16, peer), acc, bw_map_to_rank(bar_a, peer));
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(xf);
      ptx::mbar_wait(xf, (nexch >> 1) & 1);
      ++nexch;
      const float4* xb4 = reinterpret_cast<const float4*>(xbuf + (size_t)par * C * E);
      for (int e4 = tid; e4 < E / 4; e4 += NT) {
        float4 acc = xb4[e4];
        for (int r2 = 1; r2 < C; ++r2) {
          const float4 x = xb4[(size_t)r2 * (E / 4) + e4];
          acc.x += x.x;
          acc.y += x.y;
          acc.z += x.z;
          acc.w += x.w;
        }
        sink(e4, acc);
      }
      par ^= 1;
    } else {
      for (int e4 = tid; e4 < E / 4; e4 += NT) {
        float4 acc = red4[e4];
#pragma unroll
        for (int w = 1; w < NW; ++w) {
          const float4 x = red4[(size_t)w * (E / 4) + e4];
          acc.x += x.x;
          acc.y += x.y;
          acc.z += x.z;
          acc.w += x.w;
        }
        sink(e4, acc);
      }
    }
  };
  // dL/dv_raw of a frame and the carried output of the frame before it are fetched one step
  // ahead with cp.async into shared memory (no registers held across the frame; every thread
  // fetches exactly the elements it consumes itself, so a per-thread wait is enough)
  auto fetch_frame = [&](int step_) {
    const int b_ = p.sdr ? chain : chain / p.S;
    const int sf_ = p.sdr ? step_ : chain % p.S;
    const long long fr_ = (long long)b_ * p.S + sf_;
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = tid + u * NT;
      if (e < E) {
        const int ln = e & 31, qk = e >> 5, q = qk / T, k = qk % T;
        const int j = q * 32 + ln;
        const bool ok = j < O && k < D;
        const bool okv = ok && p.sdr && sf_ > 0;
        const float* sd = p.d_raw + (ok ? (fr_ * O + j) * D + k : 0);
        const float* sv = p.v_raw + (okv ? ((fr_ - 1) * O + j) * D + k : 0);
        // src-size 0 zero-fills the destination
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ptx::smem_u32(ndb + e)), "l"(sd),
                     "r"(ok ? 4 : 0));
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(ptx::smem_u32(nvb + e)), "l"(sv),
                     "r"(okv ? 4 : 0));
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch_frame(p.nsteps - 1);

  // phase timers cost 18 live registers: compiled in only with -DSRF_BWD_PHASE_TIMERS
#ifdef SRF_BWD_PHASE_TIMERS
  const bool timing = p.dbg != nullptr && tid == 0;
  long long tk[9];
#define SRF_TK(n) if (timing) tk[n] = clock64();
#else
#define SRF_TK(n)
#endif
  for (int step = p.nsteps - 1; step >= 0; --step) {
    const int b = p.sdr ? chain : chain / p.S;
    const int sf = p.sdr ? step : chain % p.S;
    const long long frame = (long long)b * p.S + sf;
    SRF_TK(0)
    // streamed u_hat: frame pair index and which member of the pair this chain is
    const long long gg = (long long)sf * p.halfB + (b >> 1);
    const bool mem1 = (b & 1) != 0;
    // bf16 -> fp32 of this chain's pair member: bytes (0, 0, lo, hi) of the selected half
    const uint32_t usel = mem1 ? 0x3244u : 0x1044u;
    const size_t ustride_i = (size_t)OPL * T4 * 128 * 2;  // elements per (pair, i)
    auto load_raw_at = [&](long long gq, int i, uint4(&dst)[RAWN]) {
      if (UM == 1) {
        const uint4* src = reinterpret_cast<const uint4*>(
            reinterpret_cast<const uint16_t*>(p.u) + ((size_t)gq * I + i) * ustride_i);
#pragma unroll
        for (int m = 0; m < OPL * T4; ++m) dst[m] = __ldg(src + m * 32 + lane);
      } else if (UM == 2) {
        const uint4* src = reinterpret_cast<const uint4*>(
            reinterpret_cast<const float*>(p.u) + ((size_t)gq * I + i) * ustride_i);
#pragma unroll
        for (int m = 0; m < OPL * T4; ++m) {
          dst[(2 * m) % RAWN] = __ldg(src + (m * 32 + lane) * 2);
          dst[(2 * m + 1) % RAWN] = __ldg(src + (m * 32 + lane) * 2 + 1);
        }
      }
    };
    auto load_raw = [&](int i, uint4(&dst)[RAWN]) { load_raw_at(gg, i, dst); };
    // element (k_in, member) of chunk m = (q, k4) sits at 2*k_in + member
    auto unpack = [&](const uint4(&raw)[RAWN], float(&u)[OPL][T]) {
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k4 = 0; k4 < T4; ++k4) {
          const int m = q * T4 + k4;
          if (UM == 1) {
            const uint32_t w[4] = {raw[m % RAWN].x, raw[m % RAWN].y, raw[m % RAWN].z, raw[m % RAWN].w};
#pragma unroll
            for (int kin = 0; kin < 4; ++kin)
              u[q][k4 * 4 + kin] = __uint_as_float(__byte_perm(w[kin], 0u, usel));  // one PRMT
          } else {
            const uint4 r0 = raw[(2 * m) % RAWN], r1 = raw[(2 * m + 1) % RAWN];
            u[q][k4 * 4 + 0] = __uint_as_float(mem1 ? r0.y : r0.x);
            u[q][k4 * 4 + 1] = __uint_as_float(mem1 ? r0.w : r0.z);
            u[q][k4 * 4 + 2] = __uint_as_float(mem1 ? r1.y : r1.x);
            u[q][k4 * 4 + 3] = __uint_as_float(mem1 ? r1.w : r1.z);
          }
        }
    };
    auto compute_u = [&](int i, float(&u)[OPL][T]) {
      const float4* xrow = reinterpret_cast<const float4*>(xs) + (size_t)(i - i_lo) * T4;
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
      }
    };

    // window gather
    for (int idx = tid; UM == 0 && idx < (i_hi - i_lo) * T; idx += NT) {
      const int l = idx % T, i = i_lo + idx / T;
      float v = 0.f;
      if (l < p.d) {
        const int w = i / p.H, hc = i - w * p.H;
        const int src = sf - p.lpad + w;
        if (src >= 0 && src < p.S) v = p.emb[(((long long)b * p.S + src) * p.H + hc) * p.d + l];
      }
      xs[idx] = v;
    }
    // g_out = dL/d v_raw[frame] + carry; Vacc_0 = previous frame's output (SDR) or 0
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int e = tid + u * NT;
      if (e < E) {
        gout[e] = ndb[e] + (p.sdr ? gacc[e] : 0.f);
        gacc[e] = 0.f;  // G_R = 0 for the backward passes of this frame
        vacc[e] = nvb[e];
      }
    }
    __syncthreads();
    if (step > 0) fetch_frame(step - 1);
    SRF_TK(1)
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
      uint4 raw[RAWN];
      if (UM != 0 && i_lo + warp < i_hi) load_raw(i_lo + warp, raw);
      for (int i = i_lo + warp; i < i_hi; i += NW) {
        float u[OPL][T], a[OPL];
        if (UM != 0) {
          unpack(raw, u);
          if (i + NW < i_hi) load_raw(i + NW, raw);
        } else {
          compute_u(i, u);
        }
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(u[q][k], va[q][k], acc);
          a[q] = ((jp < O) && !(p.mask0 && jp == 0)) ? acc : -CUDART_INF_F;
        }
        float m = a[0];
#pragma unroll
        for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[q]);
        m = UM != 0 ? bw_max_redux(m) : bw_max(m);
        float ex[OPL], z = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          ex[q] = exp2f((a[q] - m) * LOG2E);
          z += ex[q];
        }
        const float inv = 1.0f / (UM != 0 ? bw_sum_redux<OPL>(z) : bw_sum(z));
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k) ta[q][k] = fmaf(ex[q] * inv, u[q][k], ta[q][k]);
      }
      SRF_TK(2)
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) red[warp * E + (q * T + k) * 32 + lane] = ta[q][k];
      __syncthreads();
      exchange([&](int e4, float4 tot) { reinterpret_cast<float4*>(tr + r * E)[e4] = tot; });
      __syncthreads();
      SRF_TK(3)
      // Vacc_{r+1} is only needed by a later pass (the backward of pass r uses Vacc_r and t_r)
      if (r + 1 < R) {
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
    }

    SRF_TK(4)
    // ---------------- backward over the passes ----------------
    for (int r = R - 1; r >= 0; --r) {
      // g_v = [last] g_out + G_{r+1};  g_t = squash'(t_r) g_v -- every warp for its own lanes,
      // straight into registers (no serial phase, no barrier); rank 0 / warp 0 saves g_t, Vacc_r
      SRF_TK(5)
      float va[OPL][T], gtr[OPL][T], gv_acc[OPL][T];
#pragma unroll
      for (int q = 0; q < OPL; ++q) {
        float tt[T], gv[T];
        float n2 = 0.f, dot = 0.f;
#pragma unroll
        for (int k = 0; k < T; ++k) {
          const int e = (q * T + k) * 32 + lane;
          tt[k] = tr[r * E + e];
          gv[k] = (r == R - 1 ? gout[e] : 0.f) + gacc[e];
          va[q][k] = vacc[r * E + e];
          gv_acc[q][k] = 0.f;
          n2 = fmaf(tt[k], tt[k], n2);
          dot = fmaf(gv[k], tt[k], dot);
        }
        // f = n2 / ((1+n2) sqrt(n2+eps));  f'(n2) = 1/((1+n2) sq) - f/(1+n2) - f/(2 (n2+eps))
        float f, fp;
        if (UM != 0) {  // approximate rsqrt / rcp: the tensor modes' tolerance class
          const float rsq = rsqrtf(n2 + 1e-7f), r1 = __frcp_rn(1.0f + n2);
          f = n2 * r1 * rsq;
          fp = r1 * rsq - f * r1 - 0.5f * f * rsq * rsq;
        } else {
          const float sq = sqrtf(n2 + 1e-7f);
          f = n2 / ((1.0f + n2) * sq);
          fp = 1.0f / ((1.0f + n2) * sq) - f / (1.0f + n2) - 0.5f * f / (n2 + 1e-7f);
        }
#pragma unroll
        for (int k = 0; k < T; ++k) gtr[q][k] = f * gv[k] + 2.0f * fp * dot * tt[k];
        if (SPLIT && rank == 0 && warp == 0) {
          const int j = q * 32 + lane;
          if (j < O) {
            float4* gd = reinterpret_cast<float4*>(p.gtT + (((size_t)frame * R + r) * O + j) * T);
            float4* vd = reinterpret_cast<float4*>(p.vaT + (((size_t)frame * R + r) * O + j) * T);
#pragma unroll
            for (int k4 = 0; k4 < T4; ++k4) {
              gd[k4] = make_float4(gtr[q][k4 * 4], gtr[q][k4 * 4 + 1], gtr[q][k4 * 4 + 2], gtr[q][k4 * 4 + 3]);
              vd[k4] = make_float4(va[q][k4 * 4], va[q][k4 * 4 + 1], va[q][k4 * 4 + 2], va[q][k4 * 4 + 3]);
            }
          }
        }
      }
      uint4 raw[RAWN];
      if (UM != 0 && i_lo + warp < i_hi) load_raw(i_lo + warp, raw);
      for (int i = i_lo + warp; i < i_hi; i += NW) {
        float u[OPL][T], a[OPL];
        if (UM != 0) {
          unpack(raw, u);
          if (i + NW < i_hi) load_raw(i + NW, raw);
        } else {
          compute_u(i, u);
        }
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int jp = q * 32 + lane;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < T; ++k) acc = fmaf(u[q][k], va[q][k], acc);
          a[q] = ((jp < O) && !(p.mask0 && jp == 0)) ? acc : -CUDART_INF_F;
        }
        float m = a[0];
#pragma unroll
        for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[q]);
        m = UM != 0 ? bw_max_redux(m) : bw_max(m);
        float c[OPL], z = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          c[q] = exp2f((a[q] - m) * LOG2E);
          z += c[q];
        }
        const float inv = 1.0f / (UM != 0 ? bw_sum_redux<OPL>(z) : bw_sum(z));
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
        if (SPLIT) {
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            const int jp = q * 32 + lane;
            const float ga = c[q] * (gc[q] - cg);
            const size_t o = (((size_t)frame * R + r) * I + i) * OP + jp;
            p.cbuf[o] = c[q];
            p.gabuf[o] = ga;
#pragma unroll
            for (int k = 0; k < T; ++k) gv_acc[q][k] = fmaf(ga, u[q][k], gv_acc[q][k]);
          }
        } else {
          // fused variant: per-lane g_u, dbias, dW with atomics; dx summed over the lanes
          const float4* xrow = reinterpret_cast<const float4*>(xs) + (size_t)(i - i_lo) * T4;
          float dx[T];
#pragma unroll
          for (int l = 0; l < T; ++l) dx[l] = 0.f;
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            const int jp = q * 32 + lane;
            const float ga = c[q] * (gc[q] - cg);
#pragma unroll
            for (int k = 0; k < T; ++k) {
              const float gu = c[q] * gtr[q][k] + ga * va[q][k];
              gv_acc[q][k] = fmaf(ga, u[q][k], gv_acc[q][k]);
              if (jp < O && k < D) {
                float* dWrow = p.dW + (((size_t)i * O + jp) * D + k) * p.d;
                atomicAdd(p.dbias + ((size_t)i * O + jp) * D + k, gu);
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
                      atomicAdd(dWrow + l, gu * xv[li]);
                      dx[l] = fmaf(gu, wv[li], dx[l]);
                    }
                  }
                }
              }
            }
          }
          if (p.d_emb != nullptr) {
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
      }
      SRF_TK(6)
      // G_r = G_{r+1} + sum_i g_a u
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) red[warp * E + (q * T + k) * 32 + lane] = gv_acc[q][k];
      __syncthreads();
      exchange([&](int e4, float4 tot) {
        float4 g4 = reinterpret_cast<float4*>(gacc)[e4];
        g4.x += tot.x;
        g4.y += tot.y;
        g4.z += tot.z;
        g4.w += tot.w;
        reinterpret_cast<float4*>(gacc)[e4] = g4;
      });
      __syncthreads();
    }
    // gacc now holds dL/dVacc_0 = the BPTT carry into the previous frame (SDR)
#ifdef SRF_BWD_PHASE_TIMERS
    if (timing) {
      const long long t7 = clock64();
      unsigned long long* dd = p.dbg + (size_t)(blockIdx.x & 1023) * 8;
      for (int n = 0; n < 6; ++n) dd[n] += (unsigned long long)(tk[n + 1] - tk[n]);
      dd[6] += (unsigned long long)(t7 - tk[6]);
      dd[7] += 1;
    }
#endif
  }
#undef SRF_TK
  if (C > 1) cg::this_cluster().sync();  // no CTA exits while a peer may still push into it
}

// ---------------------------------------------------------------------------------------
// split mode, phase B: frame-parallel dW, dbias and dx from the saved coefficients.
//   g_u[f,i,j,k] = sum_r c[f,r,i,j] g_t[f,r,j,k] + g_a[f,r,i,j] Vacc[f,r,j,k]   (rank 2 per (i,j))
//   dW[i,j,k,l] = sum_f g_u x[f,i,l];  dbias[i,j,k] = sum_f g_u;  dx[f,i,l] = sum_jk g_u W[i,j,k,l]
// grid (I, FS frame splits, 32-capsule chunks of j).  Per tile of FT frames the CTA builds g_u in
// shared memory once ([ft][k][j], k rows padded to 36 floats so both the float4 row reads of
// the dW phase and the per-frame reads of the dx phase are conflict-free), then
//   dW phase: thread (k, l) keeps the sums for all 32 j in registers (l == d is the bias column:
//             x augmented with a 1), 9 shared loads per 32 FMAs;
//   dx phase: thread (ft, l quad, part) walks (k, 4 j) units against W[i] held in shared memory.
// Partial sums per frame split go to dwp and are folded by reduce_dwp_kernel: no atomics, the
// result is deterministic.
// ---------------------------------------------------------------------------------------
namespace {
struct DwdxPlan {
  int FT, NT, NC, P;  // frames per tile, threads, threads of the dW role (warp multiple), dx parts
  size_t smem;
};
__host__ __device__ inline int round4(int v) { return (v + 3) & ~3; }
}  // namespace

size_t dwdx_smem_bytes(int D, int d, int P, int FT) {
  const int dp = (d + 3) & ~3, xrow = (d + 2) & ~1;
  return sizeof(float) * ((size_t)32 * D * dp + (size_t)FT * (D * 36 + 4) + round4(FT * xrow) +
                          (size_t)P * FT * dp);
}

static bool dwdx_plan(const BwdParams& p, int max_smem, DwdxPlan* pl) {
  const int D = p.D, d = p.d, dp = (d + 3) & ~3, LQ = dp / 4;
  const int ntc = D * ((d + 2) / 2);  // dW role: thread = (k, pair of l; l == d is the bias column)
  pl->NC = (ntc + 31) & ~31;
  for (int FT = 32; FT >= 4; FT >>= 1) {
    const int grp = (FT / 2) * LQ;  // dx role: thread = (pair of frames, 4 consecutive l, part)
    // 384 threads leave 170 registers per thread (64 dW sums + the load groups of phase (b))
    int P = (384 - pl->NC) / grp;
    if (P < 1) P = (1024 - pl->NC) / grp;
    if (P > 4) P = 4;
    if (P < 1) continue;
    const int NT = pl->NC + ((grp * P + 31) & ~31);
    if (NT > 1024) continue;
    const size_t smem = dwdx_smem_bytes(D, d, P, FT);
    if (smem <= (size_t)max_smem) {
      pl->FT = FT;
      pl->P = P;
      pl->NT = NT;
      pl->smem = smem;
      return true;
    }
  }
  return false;
}

// number of frame splits phase B will use (the caller sizes dwp with it)
bool dwdx_uses_tensor_cores(const BwdParams& p, int max_smem);

int dwdx_frame_splits(const BwdParams& p, int max_smem, int num_sms) {
  DwdxPlan pl;
  if (dwdx_uses_tensor_cores(p, max_smem))
    pl.FT = 32;
  else if (!dwdx_plan(p, max_smem, &pl))
    return 0;
  const long long frames = (long long)p.B * p.S;
  const long long tiles = (frames + pl.FT - 1) / pl.FT;
  const int per = p.I * (p.OP / 32);
  // one CTA per SM at a time: aim for ~16 waves so that the ragged last wave costs a few percent
  // (4 waves lost 20 %), but keep at least 12 tiles per CTA to amortise the W[i] load and the
  // partial-sum write
  long long FS = (16LL * num_sms + per - 1) / per;
  if (FS > tiles / 12) FS = tiles / 12;
  if (FS > 32) FS = 32;
  if (FS < 1) FS = 1;
  return (int)FS;
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT, 1) dwdx_from_saved_kernel(const BwdParams p, int FT, int NC, int P) {
  constexpr int GB = 4;  // frames per load group of phase (b)
  extern __shared__ __align__(16) float sm[];
  const int D = p.D, d = p.d, dp = p.dp, R = p.iters, O = p.O, I = p.I, OP = p.OP, Tu = p.Tu;
  const int d1 = d + 1, LQ = dp >> 2, GS = D * 36 + 4, NT = blockDim.x;
  const int xrow = (d + 2) & ~1;           // x row: d values, the bias column of ones, zero pad
  const int L2n = (d + 2) >> 1;            // l pairs per k
  float* Ws = sm;                          // [32][D][dp]
  float* gus = Ws + 32 * D * dp;           // [FT][GS]: g_u[ft][k][j] at k*36 + j
  float* xs = gus + FT * GS;               // [FT][xrow]
  float* dxs = xs + round4(FT * xrow);     // [P][FT][dp]
  const int i = blockIdx.x, fs = blockIdx.y, q = blockIdx.z, OPL = gridDim.z;
  const int jbase = q * 32;
  const int nj = (O - jbase) < 32 ? (O - jbase) : 32;
  const int tid = threadIdx.x;
  const int w = i / p.H, hc = i - w * p.H;
  const long long frames = (long long)p.B * p.S;
  const long long f_lo = (long long)fs * p.fps;
  const long long f_hi = (f_lo + p.fps) < frames ? (f_lo + p.fps) : frames;

  for (int e = tid; e < 32 * D * dp; e += NT) {
    const int l = e % dp, jk = e / dp, k = jk % D, j = jk / D;
    float v = 0.f;
    if (j < nj && l < d) v = p.W[(((size_t)i * O + jbase + j) * D + k) * d + l];
    Ws[e] = v;
  }
  // dW role: threads [0, NC): (k, l pair), the sums over the 32 output capsules in registers
  const bool dw_thread = tid < D * L2n;
  const int k_dw = tid / L2n, lp = tid - k_dw * L2n;
  float acc0[32], acc1[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc0[j] = acc1[j] = 0.f;
  // dx role: threads [NC, NC + FT/2*LQ*P): (pair of frames, 4 consecutive l, part of the (k, j) range)
  const int td = tid - NC, FH = FT >> 1;
  const int ftp = td % FH, rest = td / FH;
  const int part = rest / LQ, lq = rest - part * LQ;
  const bool dx_thread = td >= 0 && part < P;

  for (long long f0 = f_lo; f0 < f_hi; f0 += FT) {
    const int nf = (int)((f_hi - f0) < FT ? (f_hi - f0) : FT);
    // (a) window-gathered x of capsule i for the tile (+ the bias column of ones)
    for (int e = tid; e < FT * xrow; e += NT) {
      const int ft = e / xrow, l = e - ft * xrow;
      float v = 0.f;
      if (ft < nf) {
        if (l == d) {
          v = 1.f;
        } else if (l < d) {
          const long long f = f0 + ft;
          const int b = (int)(f / p.S), s = (int)(f - (long long)b * p.S);
          const int src = s - p.lpad + w;
          if (src >= 0 && src < p.S) v = p.emb[(((long long)b * p.S + src) * p.H + hc) * d + l];
        }
      }
      xs[e] = v;
    }
    // (b) g_u of the tile: a work item is (group of GB frames, j, 4 consecutive k); g_t / Vacc
    // rows are read with 16-byte loads, all loads of an item in flight together
    {
      const int D4 = (D + 3) >> 2, per = 32 * D4, NG = (FT + GB - 1) / GB;
      const size_t cstride = (size_t)R * I * OP, gstride = (size_t)R * O * Tu;
      for (int it = tid; it < NG * per; it += NT) {
        const int grp = it / per, rem = it - grp * per;
        const int j = rem / D4, k4 = rem - j * D4;
        const bool jok = j < nj;
        const int jc = jok ? j : 0;
        const int ft0 = grp * GB;
        const float* cb = p.cbuf + ((size_t)f0 * R * I + i) * OP + jbase + jc;
        const float* gab = p.gabuf + ((size_t)f0 * R * I + i) * OP + jbase + jc;
        const float* gtb = p.gtT + ((size_t)f0 * R * O + jbase + jc) * Tu + k4 * 4;
        const float* vab = p.vaT + ((size_t)f0 * R * O + jbase + jc) * Tu + k4 * 4;
        float4 g[GB];
#pragma unroll
        for (int u = 0; u < GB; ++u) g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < R; ++r) {
          float cv[GB], gav[GB];
          float4 gtv[GB], vav[GB];
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            // frames past the end of the tile re-read the last valid one (zeroed at the store)
            const int ft = (ft0 + u) < nf ? (ft0 + u) : (nf - 1);
            const size_t oc = ft * cstride + (size_t)r * I * OP;
            const size_t og = ft * gstride + (size_t)r * O * Tu;
            cv[u] = __ldg(cb + oc);
            gav[u] = __ldg(gab + oc);
            gtv[u] = __ldg(reinterpret_cast<const float4*>(gtb + og));
            vav[u] = __ldg(reinterpret_cast<const float4*>(vab + og));
          }
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            g[u].x = fmaf(cv[u], gtv[u].x, fmaf(gav[u], vav[u].x, g[u].x));
            g[u].y = fmaf(cv[u], gtv[u].y, fmaf(gav[u], vav[u].y, g[u].y));
            g[u].z = fmaf(cv[u], gtv[u].z, fmaf(gav[u], vav[u].z, g[u].z));
            g[u].w = fmaf(cv[u], gtv[u].w, fmaf(gav[u], vav[u].w, g[u].w));
          }
        }
#pragma unroll
        for (int u = 0; u < GB; ++u) {
          const int ft = ft0 + u;
          if (ft < FT) {
            const bool ok = jok && ft < nf;
            float* dst = gus + ft * GS + (k4 * 4) * 36 + j;
            const float gv[4] = {g[u].x, g[u].y, g[u].z, g[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (k4 * 4 + c < D) dst[c * 36] = ok ? gv[c] : 0.f;
          }
        }
      }
    }
    __syncthreads();
    if (dw_thread) {
      // (c) dW / dbias: per frame one 8-byte x load and eight 16-byte g_u loads feed 64 FMAs
      const float* xp = xs + 2 * lp;
      const float* gp = gus + k_dw * 36;
      for (int ft = 0; ft < FT; ++ft) {
        const float2 xv = *reinterpret_cast<const float2*>(xp + ft * xrow);
        const float4* row = reinterpret_cast<const float4*>(gp + ft * GS);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 g4 = row[j4];
          acc0[j4 * 4 + 0] = fmaf(g4.x, xv.x, acc0[j4 * 4 + 0]);
          acc0[j4 * 4 + 1] = fmaf(g4.y, xv.x, acc0[j4 * 4 + 1]);
          acc0[j4 * 4 + 2] = fmaf(g4.z, xv.x, acc0[j4 * 4 + 2]);
          acc0[j4 * 4 + 3] = fmaf(g4.w, xv.x, acc0[j4 * 4 + 3]);
          acc1[j4 * 4 + 0] = fmaf(g4.x, xv.y, acc1[j4 * 4 + 0]);
          acc1[j4 * 4 + 1] = fmaf(g4.y, xv.y, acc1[j4 * 4 + 1]);
          acc1[j4 * 4 + 2] = fmaf(g4.z, xv.y, acc1[j4 * 4 + 2]);
          acc1[j4 * 4 + 3] = fmaf(g4.w, xv.y, acc1[j4 * 4 + 3]);
        }
      }
    } else if (dx_thread) {
      // (d) dx partials for two frames: per (k, 4 j) unit two g_u loads and four W loads feed 32 FMAs
      float4 aA = make_float4(0.f, 0.f, 0.f, 0.f), aB = aA;
      const float* gA = gus + ftp * GS;
      const float* gB = gA + FH * GS;
      const float* wbase = Ws + lq * 4;
      const int wj = D * dp;  // floats per output capsule in Ws
      for (int un = part; un < 8 * D; un += P) {
        const int k = un >> 3, j4 = un & 7;
        const float4 a4 = *reinterpret_cast<const float4*>(gA + k * 36 + j4 * 4);
        const float4 b4 = *reinterpret_cast<const float4*>(gB + k * 36 + j4 * 4);
        const float ga[4] = {a4.x, a4.y, a4.z, a4.w};
        const float gb[4] = {b4.x, b4.y, b4.z, b4.w};
        const float* wp = wbase + j4 * 4 * wj + k * dp;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float4 w4 = *reinterpret_cast<const float4*>(wp + jj * wj);
          aA.x = fmaf(ga[jj], w4.x, aA.x);
          aA.y = fmaf(ga[jj], w4.y, aA.y);
          aA.z = fmaf(ga[jj], w4.z, aA.z);
          aA.w = fmaf(ga[jj], w4.w, aA.w);
          aB.x = fmaf(gb[jj], w4.x, aB.x);
          aB.y = fmaf(gb[jj], w4.y, aB.y);
          aB.z = fmaf(gb[jj], w4.z, aB.z);
          aB.w = fmaf(gb[jj], w4.w, aB.w);
        }
      }
      *reinterpret_cast<float4*>(dxs + ((size_t)part * FT + ftp) * dp + lq * 4) = aA;
      *reinterpret_cast<float4*>(dxs + ((size_t)part * FT + ftp + FH) * dp + lq * 4) = aB;
    }
    __syncthreads();
    for (int e = tid; e < nf * dp; e += NT) {
      const int ft = e / dp, l = e - ft * dp;
      float s = 0.f;
      for (int pp = 0; pp < P; ++pp) s += dxs[((size_t)pp * FT + ft) * dp + l];
      p.dxw[((((size_t)(f0 + ft)) * I + i) * OPL + q) * dp + l] = s;
    }
  }
  if (dw_thread) {
    const int l0 = 2 * lp, l1 = l0 + 1;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < nj) {
        const size_t o = ((((size_t)fs * I + i) * O + jbase + j) * D + k_dw) * d1;
        p.dwp[o + l0] = acc0[j];
        if (l1 < d1) p.dwp[o + l1] = acc1[j];
      }
  }
}

// ---------------------------------------------------------------------------------------
// phase B on the tensor cores (uhat_mode TF32 / F16 / BF16: the gradient carries the rounding class
// of the mode anyway).  Same grid, same shared-memory tiles and the same deterministic partial-sum
// scheme as dwdx_from_saved_kernel; the two contractions of a 32-frame tile run as warp-level
// mma.sync.m16n8k8 TF32 with fp32 accumulation in registers:
//   dW[(k,j)][l]  += g_u^T [(k,j)] [ft] . x  [ft] [l]      M = 32 D rows, N = d+1 (bias column), K = 32 frames
//   dx [ft] [l]    = g_u   [ft] [(k,j)] . W_i [(k,j)] [l]  M = 32 frames, N = d,   K = 32 D
// 16 warps; warp w owns the 16-row blocks w, w+16, ... of dW (all tiles of the CTA's frame split
// accumulate in its registers) and the K steps w, w+8, ... of dx (partials folded through shared memory).
// g_u, x and W_i are rounded to TF32 (rna) when they are written to shared memory.
// ---------------------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float to_tf32(float v) {
  uint32_t t;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
  return __uint_as_float(t);
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const float (&a)[4], float b0, float b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
        "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
// 16 warps: the kernel is bound by the latency of the g_u rebuild (dependent L2 loads, then shared-memory
// stores), not by a pipe -- 16 instead of 8 warps: 9.9 -> 8.8 ms per WSJ layer at 64 x 375 frames.  (Also
// measured and dropped: two input capsules per CTA on 16-frame tiles sharing one pass over g_t / Vacc,
// 10.7 ms; lane = output capsule rows with conflict-free stores, 9.7 ms with 8 warps, 9.2 with 16.)
#ifndef SRF_DWM_WARPS
#define SRF_DWM_WARPS 16
#endif
constexpr int DWM_FT = 32, DWM_WARPS = SRF_DWM_WARPS;
}  // namespace

size_t dwdx_mma_smem_bytes(int D, int d) {
  const int dp = (d + 3) & ~3, xrow = (d + 2) & ~1;
  return sizeof(float) * ((size_t)32 * D * dp + (size_t)DWM_FT * (D * 36 + 4) + round4(DWM_FT * xrow) +
                          (size_t)DWM_WARPS * DWM_FT * dp);
}

// MBW = 16-row blocks of dW per warp (2 D / 8 rounded up), NB = 8-column blocks of l
template <int MBW, int NB>
__global__ void __launch_bounds__(DWM_WARPS * 32, 1) dwdx_from_saved_mma_kernel(const BwdParams p) {
  constexpr int GB = 4, FT = DWM_FT, NT = DWM_WARPS * 32;
  extern __shared__ __align__(16) float sm[];
  const int D = p.D, d = p.d, dp = p.dp, R = p.iters, O = p.O, I = p.I, OP = p.OP, Tu = p.Tu;
  const int d1 = d + 1, GS = D * 36 + 4;
  const int xrow = (d + 2) & ~1;
  float* Ws = sm;                          // [32][D][dp]
  float* gus = Ws + 32 * D * dp;           // [FT][GS]: g_u[ft][k][j] at k*36 + j
  float* xs = gus + FT * GS;               // [FT][xrow]
  float* dxs = xs + round4(FT * xrow);     // [warps][FT][dp]
  const int i = blockIdx.x, fs = blockIdx.y, q = blockIdx.z, OPL = gridDim.z;
  const int jbase = q * 32;
  const int nj = (O - jbase) < 32 ? (O - jbase) : 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int w = i / p.H, hc = i - w * p.H;
  const long long frames = (long long)p.B * p.S;
  const long long f_lo = (long long)fs * p.fps;
  const long long f_hi = (f_lo + p.fps) < frames ? (f_lo + p.fps) : frames;

  for (int e = tid; e < 32 * D * dp; e += NT) {
    const int l = e % dp, jk = e / dp, k = jk % D, j = jk / D;
    float v = 0.f;
    if (j < nj && l < d) v = p.W[(((size_t)i * O + jbase + j) * D + k) * d + l];
    Ws[e] = to_tf32(v);
  }
  float acc[MBW][NB][4];
#pragma unroll
  for (int a = 0; a < MBW; ++a)
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[a][b][c] = 0.f;

  for (long long f0 = f_lo; f0 < f_hi; f0 += FT) {
    const int nf = (int)((f_hi - f0) < FT ? (f_hi - f0) : FT);
    // (a) window-gathered x of capsule i for the tile (+ the bias column of ones)
    for (int e = tid; e < FT * xrow; e += NT) {
      const int ft = e / xrow, l = e - ft * xrow;
      float v = 0.f;
      if (ft < nf) {
        if (l == d) {
          v = 1.f;
        } else if (l < d) {
          // (frame indices fit 32 bits; a 64-bit division here was 8 % of the kernel's stall samples)
          const unsigned f = (unsigned)(f0 + ft);
          const int b = (int)(f / (unsigned)p.S), s_ = (int)(f - (unsigned)b * (unsigned)p.S);
          const int src = s_ - p.lpad + w;
          if (src >= 0 && src < p.S) v = p.emb[(((long long)b * p.S + src) * p.H + hc) * d + l];
        }
      }
      xs[e] = to_tf32(v);
    }
    // (b) g_u of the tile (as in dwdx_from_saved_kernel), rounded to TF32
    {
      const int D4 = (D + 3) >> 2, per = 32 * D4, NG = (FT + GB - 1) / GB;
      const size_t cstride = (size_t)R * I * OP, gstride = (size_t)R * O * Tu;
      for (int it = tid; it < NG * per; it += NT) {
        const int grp = it / per, rem = it - grp * per;
        const int j = rem / D4, k4 = rem - j * D4;
        const bool jok = j < nj;
        const int jc = jok ? j : 0;
        const int ft0 = grp * GB;
        const float* cb = p.cbuf + ((size_t)f0 * R * I + i) * OP + jbase + jc;
        const float* gab = p.gabuf + ((size_t)f0 * R * I + i) * OP + jbase + jc;
        const float* gtb = p.gtT + ((size_t)f0 * R * O + jbase + jc) * Tu + k4 * 4;
        const float* vab = p.vaT + ((size_t)f0 * R * O + jbase + jc) * Tu + k4 * 4;
        float4 gg[GB];
#pragma unroll
        for (int u = 0; u < GB; ++u) gg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < R; ++r) {
          float cv[GB], gav[GB];
          float4 gtv[GB], vav[GB];
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            const int ft = (ft0 + u) < nf ? (ft0 + u) : (nf - 1);
            const size_t oc = ft * cstride + (size_t)r * I * OP;
            const size_t og = ft * gstride + (size_t)r * O * Tu;
            cv[u] = __ldg(cb + oc);
            gav[u] = __ldg(gab + oc);
            gtv[u] = __ldg(reinterpret_cast<const float4*>(gtb + og));
            vav[u] = __ldg(reinterpret_cast<const float4*>(vab + og));
          }
#pragma unroll
          for (int u = 0; u < GB; ++u) {
            gg[u].x = fmaf(cv[u], gtv[u].x, fmaf(gav[u], vav[u].x, gg[u].x));
            gg[u].y = fmaf(cv[u], gtv[u].y, fmaf(gav[u], vav[u].y, gg[u].y));
            gg[u].z = fmaf(cv[u], gtv[u].z, fmaf(gav[u], vav[u].z, gg[u].z));
            gg[u].w = fmaf(cv[u], gtv[u].w, fmaf(gav[u], vav[u].w, gg[u].w));
          }
        }
#pragma unroll
        for (int u = 0; u < GB; ++u) {
          const int ft = ft0 + u;
          if (ft < FT) {
            const bool ok = jok && ft < nf;
            float* dst = gus + ft * GS + (k4 * 4) * 36 + j;
            const float gv[4] = {gg[u].x, gg[u].y, gg[u].z, gg[u].w};
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (k4 * 4 + c < D) dst[c * 36] = ok ? to_tf32(gv[c]) : 0.f;
          }
        }
      }
    }
    __syncthreads();
    // (c) dW / dbias: A[(k, j0 + row)][ft] = g_u, B[ft][l] = x
#pragma unroll
    for (int ks = 0; ks < FT / 8; ++ks) {
      float b0[NB], b1[NB];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        const int l = nb * 8 + g;
        b0[nb] = l < xrow ? xs[(ks * 8 + t) * xrow + l] : 0.f;
        b1[nb] = l < xrow ? xs[(ks * 8 + t + 4) * xrow + l] : 0.f;
      }
#pragma unroll
      for (int mi = 0; mi < MBW; ++mi) {
        const int mb = warp + DWM_WARPS * mi;
        if (mb < 2 * D) {
          const float* ap = gus + (ks * 8 + t) * GS + (mb >> 1) * 36 + (mb & 1) * 16 + g;
          const float a[4] = {ap[0], ap[8], ap[4 * GS], ap[4 * GS + 8]};
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) mma_tf32_16x8x8(acc[mi][nb], a, b0[nb], b1[nb]);
        }
      }
    }
    // (d) dx: A[ft][(k, j0 + col)] = g_u, B[(k, j)][l] = W_i; this warp's share of the K steps
    {
      float ax[2][NB][4];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int c = 0; c < 4; ++c) ax[a][b][c] = 0.f;
      for (int ks = warp; ks < 4 * D; ks += DWM_WARPS) {
        const int k = ks >> 2, j0 = (ks & 3) * 8;
        float b0[NB], b1[NB];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int l = nb * 8 + g;
          b0[nb] = l < dp ? Ws[((j0 + t) * D + k) * dp + l] : 0.f;
          b1[nb] = l < dp ? Ws[((j0 + t + 4) * D + k) * dp + l] : 0.f;
        }
#pragma unroll
        for (int mbx = 0; mbx < 2; ++mbx) {
          const float* ap = gus + (mbx * 16 + g) * GS + k * 36 + j0 + t;
          const float a[4] = {ap[0], ap[8 * GS], ap[4], ap[8 * GS + 4]};
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) mma_tf32_16x8x8(ax[mbx][nb], a, b0[nb], b1[nb]);
        }
      }
#pragma unroll
      for (int mbx = 0; mbx < 2; ++mbx)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int l = nb * 8 + 2 * t;
          float* o0 = dxs + ((size_t)warp * FT + mbx * 16 + g) * dp + l;
          float* o1 = o0 + 8 * dp;
          if (l < dp) {
            o0[0] = ax[mbx][nb][0];
            o1[0] = ax[mbx][nb][2];
          }
          if (l + 1 < dp) {
            o0[1] = ax[mbx][nb][1];
            o1[1] = ax[mbx][nb][3];
          }
        }
    }
    __syncthreads();
    for (int e = tid; e < nf * dp; e += NT) {
      const int ft = e / dp, l = e - ft * dp;
      float s_ = 0.f;
#pragma unroll
      for (int pp = 0; pp < DWM_WARPS; ++pp) s_ += dxs[((size_t)pp * FT + ft) * dp + l];
      p.dxw[((((size_t)(f0 + ft)) * I + i) * OPL + q) * dp + l] = s_;
    }
    // the next tile's (a) / (b) overwrite xs / gus: every warp must be done with (c) and (d), which the
    // barrier above guarantees; dxs is rewritten only after the next tile's first barrier
  }
#pragma unroll
  for (int mi = 0; mi < MBW; ++mi) {
    const int mb = warp + DWM_WARPS * mi;
    if (mb >= 2 * D) continue;
    const int k = mb >> 1, j0 = (mb & 1) * 16;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int j = j0 + g + (c >> 1) * 8, l = nb * 8 + 2 * t + (c & 1);
        if (j < nj && l < d1)
          p.dwp[((((size_t)fs * I + i) * O + jbase + j) * D + k) * d1 + l] = acc[mi][nb][c];
      }
  }
}

// dW += sum_fs dwp[fs][..][l < d],  dbias += sum_fs dwp[fs][..][d]
__global__ void reduce_dwp_kernel(const BwdParams p) {
  const int d1 = p.d + 1;
  const long long n = (long long)p.I * p.O * p.D * d1;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int fs = 0; fs < p.FS; ++fs) s += p.dwp[(size_t)fs * n + e];
    const long long row = e / d1;
    const int l = (int)(e - row * d1);
    if (l < p.d) p.dW[row * p.d + l] += s;
    else p.dbias[row] += s;
  }
}

// tensor-core phase B variant for the shape, or null: MBW covers 2 D row blocks over 8 warps, NB the
// d+1 columns
typedef void (*DwdxMmaKernel)(const BwdParams);
static DwdxMmaKernel dwdx_mma_variant(int D, int d) {
  const int mbw = (2 * D + DWM_WARPS - 1) / DWM_WARPS, nb = (d + 1 + 7) / 8;
  if (nb > 3 || mbw > 5) return nullptr;
  const int MB = mbw <= 2 ? 2 : (mbw <= 4 ? 4 : 5), NBc = nb <= 2 ? 2 : 3;
  if (MB == 2) return NBc == 2 ? dwdx_from_saved_mma_kernel<2, 2> : dwdx_from_saved_mma_kernel<2, 3>;
  if (MB == 4) return NBc == 2 ? dwdx_from_saved_mma_kernel<4, 2> : dwdx_from_saved_mma_kernel<4, 3>;
  return NBc == 2 ? dwdx_from_saved_mma_kernel<5, 2> : dwdx_from_saved_mma_kernel<5, 3>;
}

bool dwdx_uses_tensor_cores(const BwdParams& p, int max_smem) {
  return p.tc_phase_b && dwdx_mma_variant(p.D, p.d) != nullptr &&
         dwdx_mma_smem_bytes(p.D, p.d) <= (size_t)max_smem;
}

cudaError_t launch_dwdx_from_saved(const BwdParams& p, int max_smem, cudaStream_t stream) {
  dim3 grid(p.I, p.FS, p.OP / 32);
  cudaError_t e;
  if (dwdx_uses_tensor_cores(p, max_smem)) {
    DwdxMmaKernel kern = dwdx_mma_variant(p.D, p.d);
    const size_t smem = dwdx_mma_smem_bytes(p.D, p.d);
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, DWM_WARPS * 32, smem, stream>>>(p);
  } else {
    DwdxPlan pl;
    if (!dwdx_plan(p, max_smem, &pl)) return cudaErrorInvalidValue;
    auto go = [&](auto kern) -> cudaError_t {
      cudaError_t ee = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem);
      if (ee != cudaSuccess) return ee;
      kern<<<grid, pl.NT, pl.smem, stream>>>(p, pl.FT, pl.NC, pl.P);
      return cudaSuccess;
    };
    if (pl.NT <= 384) e = go(dwdx_from_saved_kernel<384>);
    else if (pl.NT <= 512) e = go(dwdx_from_saved_kernel<512>);
    else e = go(dwdx_from_saved_kernel<1024>);
    if (e != cudaSuccess) return e;
  }
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const long long n = (long long)p.I * p.O * p.D * (p.d + 1);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  reduce_dwp_kernel<<<blocks, 256, 0, stream>>>(p);
  return cudaGetLastError();
}

// d_emb[b,s',h,l] += sum_w sum_q dxw[(b, s'+lpad-w), w*H+h, q, l]   (fold the window back)
__global__ void fold_dx_kernel(const BwdParams p) {
  const long long n = (long long)p.B * p.S * p.H * p.d;
  const int window = p.I / p.H, OPL = p.OP / 32;
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
      if (sf >= 0 && sf < p.S) {
        const float* src = p.dxw + ((((long long)b * p.S + sf) * p.I + w * p.H + hc) * OPL) * p.dp + l;
        for (int c = 0; c < OPL; ++c) acc += src[c * p.dp];
      }
    }
    p.d_emb[e] += acc;
  }
}

void launch_fold_dx(const BwdParams& p, cudaStream_t stream) {
  const long long n = (long long)p.B * p.S * p.H * p.d;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  fold_dx_kernel<<<blocks, 256, 0, stream>>>(p);
}

template <int T, int OPL, int NW, int UM, bool SPLIT>
static cudaError_t launch_bwd_variant(const BwdParams& p, int nchains, cudaStream_t stream) {
  const size_t E = (size_t)OPL * T * 32;
  const size_t smem = sizeof(float) * (4 + (UM == 0 ? (size_t)p.Ic * T : 0) + (NW + 2 * p.C) * E +
                                       (2 * BW_MAX_ITERS + 1) * E + 5 * E);
  auto kern = route_layer_bwd_kernel<T, OPL, NW, UM, SPLIT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(nchains * p.C));
  cfg.blockDim = dim3(NW * 32);
  cfg.dynamicSmemBytes = smem;
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

size_t route_layer_bwd_smem_bytes(int T, int OPL, int um, int Ic, int C) {
  const size_t E = (size_t)OPL * T * 32;
  const size_t NW = (size_t)route_layer_bwd_warps(um, T, OPL);
  return sizeof(float) * (4 + (um == 0 ? (size_t)Ic * T : 0) + (NW + 2 * (size_t)C) * E +
                          (2 * BW_MAX_ITERS + 1) * E + 5 * E);
}

// warps per CTA of the sweep: 16 / 12 while the per-lane state (OPL*T values per array) leaves
// room under the register budget, else 8
int route_layer_bwd_warps(int um, int T, int OPL) {
  if (um == 0 || T * OPL > 20) return 8;
  return um == 1 ? SRF_BWD_NW_BF16 : SRF_BWD_NW_F32;
}

#define SRF_BWD(T_, OPL_)                                                                      \
  if (T == T_ && OPL == OPL_) {                                                                \
    if (um == 1)                                                                               \
      return launch_bwd_variant<T_, OPL_, (T_ * OPL_ <= 20 ? SRF_BWD_NW_BF16 : 8), 1, true>(p, nchains, stream); \
    if (um == 2)                                                                               \
      return launch_bwd_variant<T_, OPL_, (T_ * OPL_ <= 20 ? SRF_BWD_NW_F32 : 8), 2, true>(p, nchains, stream);  \
    if (p.split) return launch_bwd_variant<T_, OPL_, 8, 0, true>(p, nchains, stream);         \
    return launch_bwd_variant<T_, OPL_, 8, 0, false>(p, nchains, stream);                     \
  }

cudaError_t launch_route_layer_bwd(const BwdParams& p, int T, int OPL, int um, int nchains,
                                   cudaStream_t stream) {
  if (p.iters > BW_MAX_ITERS) return cudaErrorInvalidValue;
  if (um != 0 && !p.split) return cudaErrorInvalidValue;
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
