// routing_stream.cu -- routing iterations over a MATERIALISED u_hat (tensor-core path), sm_100a.
//
// u_hat of the whole layer was produced by the tcgen05 GEMM (uhat_gemm.cu) in the layout
//   [frame pair g][input capsule i][chunk m = q*(T/4)+k4][lane*4 + k%4][pair member f]   (bf16|fp32)
// so that for one (g, i) every lane (= output capsule j) owns 16-byte chunks and a warp's read is
// one contiguous slab.  This kernel is the sequential part of the layer
// (tfsr/model/sequence_router_naive.py:162-191): SDR frame recurrence or DR iterations, squash,
// LayerNorm + dropout, head.  It is a streaming kernel:
//
//   warp NW     : producer.  One lane issues cp.async.bulk (TMA 1-D) copies of the CTA's slice of
//                 u_hat into a shared-memory ring, NW capsules per stage, running ahead of the
//                 consumers across frames (u_hat does not depend on the recurrence).
//   warps 0..NW-1: consumers.  lane = output capsule; u_hat chunk -> registers, agreement with
//                 Vacc, softmax over output capsules by warp shuffles, t += c*u_hat in registers.
//   warps NW+1,+2: output.  LayerNorm(O*D), dropout mask, head and the global stores of frame s run
//                 off a double-buffered copy of v while the consumers already work on frame s+1.
//
// The partial sums of the C CTAs of a cluster (input capsules are split between them) are
// exchanged by pushing into the peers' shared memory (st.shared::cluster) and signalling the
// peers' mbarriers -- no cluster-wide barrier on the per-frame critical path.

#include <cuda_runtime.h>

#include <cstdlib>
#include <math_constants.h>

#include "routing_kernels.h"
#include "sm100_ptx.cuh"

namespace srf {

namespace {

__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}

enum { BAR_COMPUTE = 1 };

__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}

// remote store of 4 floats into a peer CTA's shared memory that completes (by byte count) on
// that peer's mbarrier: data visibility and signalling in one instruction, no fences.
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile(
      "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
          remote_addr),
      "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar)
      : "memory");
}

}  // namespace

// NSLOT = input capsules per ring stage; FPW = pair members (frames) per consumer warp.
// FPW = 1: 2*NSLOT consumer warps, warp w works on capsule slot w>>1 for pair member w&1;
// FPW = 2: NSLOT consumer warps, each handles both members (for variants with a large per-lane
// state O*D/32 that would not fit the register budget of 2*NSLOT+3 warps).
template <int T, int OPL, int NSLOT, int FPW, bool BF16>
__global__ void __launch_bounds__((NSLOT * (2 / FPW) + 3) * 32, 1)
route_stream_kernel(const RouteParams p) {
  constexpr int T4 = T / 4;
  constexpr int NCW = NSLOT * (2 / FPW);       // consumer warps
  constexpr int NCT = NCW * 32;                // consumer threads
  constexpr int EF = OPL * T * 32;             // t tile of ONE pair member: (q*T+k)*32+lane
  constexpr int E = 2 * EF;                    // both members: f*EF + ...
  // FPW == 1: the warps of pair member 0 and of pair member 1 form two independent TEAMS (the
  // members are different utterances): own named barrier, own exchange buffers and mbarriers.
  // The teams share only the u_hat ring, so they drift out of phase and one team's
  // reduce/exchange/squash tail overlaps the other team's capsule loop.
  constexpr int NTEAM = FPW == 1 ? 2 : 1;
  constexpr int TW = NCW / NTEAM;              // warps per team
  constexpr int TT = TW * 32;                  // threads per team
  constexpr int ET = E / NTEAM;                // t elements a team owns
  constexpr int SLAB = OPL * T4 * 128 * 2 * (BF16 ? 2 : 4);  // bytes of u_hat per (pair, capsule)
  constexpr int STAGE = NSLOT * SLAB;
  constexpr int RAWN = SLAB / 512;             // uint4 per lane per capsule
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int C = p.C, NSTAGE = p.nstage;
  const int rank = C > 1 ? (int)cluster_rank() : 0;
  const int group = blockIdx.x / C;
  const int i_lo = rank * p.Ic;
  const int i_hi = min(p.I, i_lo + p.Ic);
  const int O = p.O, D = p.D;

  uint8_t* ring = smem_raw;                                        // [NSTAGE][STAGE]
  float* red = reinterpret_cast<float*>(ring + (size_t)NSTAGE * STAGE);  // [NSLOT][E] warp partials
  float* xbuf = red + NSLOT * E;                                   // [2][C][E] CTA partial sums
  float* tot = xbuf + 2 * C * E;                                   // [E] cluster-summed t
  float* vout = tot + E;                                           // [2][E]
  uint64_t* full = reinterpret_cast<uint64_t*>(vout + 2 * E);      // [NSTAGE]
  uint64_t* empty = full + NSTAGE;                                 // [NSTAGE]
  uint64_t* xfull = empty + NSTAGE;                                // [team][2]
  uint64_t* vready = xfull + 4;                                    // [member][2] v of frame s is in vout[s&1]
  uint64_t* vfree = vready + 4;                                    // [member][2] the output warp is done with it

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], NCW);
    }
    for (int b = 0; b < 4; ++b) {
      ptx::mbar_init(&xfull[b], 1 + TW);  // expect_tx arrive + one arrive per warp of the team
      ptx::mbar_init(&vready[b], 1);
      ptx::mbar_init(&vfree[b], 1);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (C > 1) cluster_sync_all();  // peers' barriers are initialised before anybody pushes

  const size_t pair_stride = (size_t)p.I * SLAB;  // bytes per frame pair

  if (warp == NCW) {
    // ============================== producer ==============================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int s = 0; s < p.nsteps; ++s) {
        const long long gg = p.sdr ? ((long long)s * p.halfB + group) : (long long)group;
        const uint8_t* src_pair = reinterpret_cast<const uint8_t*>(p.u) + (size_t)gg * pair_stride;
        for (int pass = 0; pass < p.iters; ++pass) {
          for (int base = i_lo; base < i_hi; base += NSLOT) {
            const int cnt = min(NSLOT, i_hi - base);
            ptx::mbar_wait(&empty[st], ph ^ 1);
            ptx::mbar_arrive_expect_tx(&full[st], (uint32_t)cnt * SLAB);
            ptx::bulk_g2s(ring + (size_t)st * STAGE, src_pair + (size_t)base * SLAB,
                          (uint32_t)cnt * SLAB, &full[st]);
            if (++st == NSTAGE) {
              st = 0;
              ph ^= 1;
            }
          }
        }
      }
    }
  } else if (warp < NCW) {
    // ============================== consumers ==============================
    const int slot = FPW == 1 ? (warp >> 1) : warp;   // capsule slot inside a stage
    const int f0 = FPW == 1 ? (warp & 1) : 0;         // first pair member of this warp
    const int team = f0;                              // (0 when FPW == 2)
    const int ttid = FPW == 1 ? slot * 32 + lane : tid;   // thread index inside the team
    const int tb = team * ET;                         // the team's offset inside E-sized buffers
    const int bar_id = BAR_COMPUTE + team;
    // bf16 -> fp32 of pair member f0 in one PRMT: bytes (0, 0, lo, hi) of the selected half
    const uint32_t usel = f0 ? 0x3244u : 0x1044u;
    const uint32_t ring_a = ptx::smem_u32(ring) + (uint32_t)slot * SLAB + (uint32_t)lane * (BF16 ? 16u : 32u);
    const uint32_t full_a = ptx::smem_u32(full), empty_a = ptx::smem_u32(empty);
    int st = 0;
    uint32_t ph = 0;
    uint32_t npass = 0;   // exchange counter
    // Vacc = sum of the squashed outputs the logits are linear in (previous frame's output for
    // SDR) -- kept in registers across passes and frames; lane = output capsule.
    float va[FPW][OPL][T];
#pragma unroll
    for (int f = 0; f < FPW; ++f)
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) va[f][q][k] = 0.f;
    for (int s = 0; s < p.nsteps; ++s) {
      for (int pass = 0; pass < p.iters; ++pass) {
        const bool last_pass = pass == p.iters - 1;
        float ta[FPW][OPL][T];
#pragma unroll
        for (int f = 0; f < FPW; ++f)
#pragma unroll
          for (int q = 0; q < OPL; ++q)
#pragma unroll
            for (int k = 0; k < T; ++k) ta[f][q][k] = 0.f;
        // phase timers hold 12 live registers: compiled in only with -DSRF_STREAM_PHASE_TIMERS
#ifdef SRF_STREAM_PHASE_TIMERS
        long long tk0 = 0, tk1 = 0, tk2 = 0, tk3 = 0, tk4 = 0, tk5 = 0;
        const bool timing = p.dbg != nullptr && tid == 0;
#define SRF_STK(v) if (timing) v = clock64();
#else
#define SRF_STK(v)
#endif
        SRF_STK(tk0)

        int rel_st = -1;  // stage still to be released
        for (int base = i_lo; base < i_hi; base += NSLOT) {
          const int i = base + slot;
          uint4 raw[RAWN];
          while (!mbar_try_wait_a(full_a + st * 8, ph)) {
          }
          if (i < i_hi) {
            const uint32_t a0 = ring_a + (uint32_t)st * STAGE;
#pragma unroll
            for (int m = 0; m < RAWN; ++m)
              raw[m] = BF16 ? lds128(a0 + m * 512) : lds128(a0 + (m >> 1) * 1024 + (m & 1) * 16);
          }
          // The stage is handed back to the producer one round LATE (at the top of the next
          // round, or after the loop): by then every register loaded from it has been consumed
          // by a full round of math, so no shared-memory load of this warp can still be in
          // flight when the TMA engine overwrites the stage.  (Arriving right after issuing the
          // loads let the refill race with loads delayed by bank-conflict replays.)
          __syncwarp();
          if (lane == 0 && rel_st >= 0) mbar_arrive_a(empty_a + rel_st * 8);
          rel_st = st;
          if (++st == NSTAGE) {
            st = 0;
            ph ^= 1;
          }
          if (i >= i_hi) continue;

          float u[FPW][OPL][T];
#pragma unroll
          for (int q = 0; q < OPL; ++q)
#pragma unroll
            for (int k4 = 0; k4 < T4; ++k4) {
              const int m = q * T4 + k4;
              if (BF16) {
                const uint32_t w[4] = {raw[m].x, raw[m].y, raw[m].z, raw[m].w};
#pragma unroll
                for (int kin = 0; kin < 4; ++kin) {
                  if (FPW == 2) {
                    u[0][q][k4 * 4 + kin] = __uint_as_float(w[kin] << 16);
                    u[FPW - 1][q][k4 * 4 + kin] = __uint_as_float(w[kin] & 0xffff0000u);
                  } else {
                    u[0][q][k4 * 4 + kin] = __uint_as_float(__byte_perm(w[kin], 0u, usel));
                  }
                }
              } else {
                const uint4 r0 = raw[(2 * m) % RAWN], r1 = raw[(2 * m + 1) % RAWN];
                const uint32_t w[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                for (int kin = 0; kin < 4; ++kin) {
                  const float lo = __uint_as_float(w[2 * kin]);
                  const float hi = __uint_as_float(w[2 * kin + 1]);
                  if (FPW == 2) {
                    u[0][q][k4 * 4 + kin] = lo;
                    u[FPW - 1][q][k4 * 4 + kin] = hi;
                  } else {
                    u[0][q][k4 * 4 + kin] = f0 ? hi : lo;
                  }
                }
              }
            }
          // agreement with the accumulated outputs (naive:205 / :223 / :240)
          float a[FPW][OPL];
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            const int jp = q * 32 + lane;
            const bool valid = (jp < O) && !(p.mask0 && jp == 0);
#pragma unroll
            for (int f = 0; f < FPW; ++f) {
              float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
#pragma unroll
              for (int k = 0; k < T; k += 4) {
                acc0 = fmaf(u[f][q][k], va[f][q][k], acc0);
                acc1 = fmaf(u[f][q][k + 1], va[f][q][k + 1], acc1);
                acc2 = fmaf(u[f][q][k + 2], va[f][q][k + 2], acc2);
                acc3 = fmaf(u[f][q][k + 3], va[f][q][k + 3], acc3);
              }
              a[f][q] = valid ? (acc0 + acc1) + (acc2 + acc3) : -CUDART_INF_F;
            }
          }
          // coupling softmax over output capsules (naive:202 / :225 / :241) + weighted sum.
          // Both warp reductions are single REDUX instructions: the max on an order-preserving
          // integer image of the floats, the sum on a Q26 fixed-point image of exp(a - max) in
          // (0, 1] (32 terms <= 2^31, unsigned; the normaliser keeps ~2^-21 relative accuracy, the
          // coefficients themselves stay fp32).
#pragma unroll
          for (int f = 0; f < FPW; ++f) {
            float m = a[f][0];
#pragma unroll
            for (int q = 1; q < OPL; ++q) m = fmaxf(m, a[f][q]);
            int mi = __float_as_int(m);
            mi ^= (mi >> 31) & 0x7fffffff;
            mi = __reduce_max_sync(0xffffffffu, mi);
            mi ^= (mi >> 31) & 0x7fffffff;
            m = __int_as_float(mi);
            float ex[OPL];
            float zl = 0.f;
#pragma unroll
            for (int q = 0; q < OPL; ++q) {
              ex[q] = fast_ex2((a[f][q] - m) * LOG2E);
              zl += ex[q];
            }
            // OPL <= 4 terms per lane, each <= 1: Q(26 - log2 OPL) keeps the warp sum below 2^31
            constexpr float QS = (float)(1 << 26) / (float)OPL;
            // unsigned: 32 lanes x 2^26 = 2^31 exactly when the coupling is uniform
            const unsigned zi = __reduce_add_sync(0xffffffffu, __float2uint_rn(zl * QS));
            const float inv = fast_rcp((float)zi * (1.0f / QS));
#pragma unroll
            for (int q = 0; q < OPL; ++q) {
              const float c = ex[q] * inv;
#pragma unroll
              for (int k = 0; k < T; ++k) ta[f][q][k] = fmaf(c, u[f][q][k], ta[f][q][k]);
            }
          }
        }

        SRF_STK(tk1)
        // ---- reduce t over the capsule slots of this CTA -------------------------------------
#pragma unroll
        for (int f = 0; f < FPW; ++f)
#pragma unroll
          for (int q = 0; q < OPL; ++q)
#pragma unroll
            for (int k = 0; k < T; ++k)
              red[slot * E + (f0 + f) * EF + (q * T + k) * 32 + lane] = ta[f][q][k];
        // release the last stage of the pass (its data went into the ta just stored)
        __syncwarp();
        if (lane == 0 && rel_st >= 0) mbar_arrive_a(empty_a + rel_st * 8);
        const int par = npass & 1;
        uint64_t* xf = &xfull[team * 2 + par];
        if (C > 1 && ttid == 0) ptx::mbar_arrive_expect_tx(xf, (uint32_t)(C - 1) * ET * 4);
        named_sync(bar_id, TT);
        SRF_STK(tk2)
        {
          // sum the NSLOT warp partials (float4 = 4 consecutive lanes) and publish the CTA partial:
          // own copy with a plain store, the peers' copies with st.async into their xbuf[par][rank]
          float* mine = xbuf + ((size_t)par * C + rank) * E + tb;
          const uint32_t mine_a = ptx::smem_u32(mine);
          const uint32_t bar_a = ptx::smem_u32(xf);
          const float4* red4 = reinterpret_cast<const float4*>(red + tb);
          for (int e4 = ttid; e4 < ET / 4; e4 += TT) {
            float4 acc = red4[e4];
#pragma unroll
            for (int w = 1; w < NSLOT; ++w) {
              const float4 x = red4[(size_t)w * (E / 4) + e4];
              acc.x += x.x;
              acc.y += x.y;
              acc.z += x.z;
              acc.w += x.w;
            }
            reinterpret_cast<float4*>(mine)[e4] = acc;  // own copy: plain shared store
            if (C > 1) {
              for (int r = 1; r < C; ++r) {
                const int peer = (rank + r) & (C - 1);  // C is a power of two
                st_async_v4(map_to_rank(mine_a + e4 * 16, peer), acc, map_to_rank(bar_a, peer));
              }
            }
          }
          if (C > 1) {
            // the local stores are published to the team by one mbarrier arrive per warp
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(xf);
          }
        }
        SRF_STK(tk3)
        if (C > 1) {
          ptx::mbar_wait(xf, (npass >> 1) & 1);
        } else {
          named_sync(bar_id, TT);
        }
        ++npass;
        SRF_STK(tk4)

        // ---- cluster sum (cooperative, float4) -> tot ------------------------------------------
        if (C > 1) {
          const float4* xb4 = reinterpret_cast<const float4*>(xbuf + (size_t)par * C * E + tb);
          for (int e4 = ttid; e4 < ET / 4; e4 += TT) {
            float4 acc = xb4[e4];
            for (int r = 1; r < C; ++r) {
              const float4 x = xb4[(size_t)r * (E / 4) + e4];
              acc.x += x.x;
              acc.y += x.y;
              acc.z += x.z;
              acc.w += x.w;
            }
            reinterpret_cast<float4*>(tot + tb)[e4] = acc;
          }
          named_sync(bar_id, TT);
        }
        const float* total = C > 1 ? tot : xbuf + (size_t)par * E;
        SRF_STK(tk5)
        // ---- squash (naive:248-253) + Vacc update: every consumer warp does this for its own
        // (member, lane) columns, the result feeds its registers directly; the slot-0 warps also
        // hand v to the output warps.
        const bool writer = last_pass && slot == 0;
        if (writer && s >= 2) {
#pragma unroll
          for (int f = 0; f < FPW; ++f) ptx::mbar_wait(&vfree[(f0 + f) * 2 + (s & 1)], ((s >> 1) - 1) & 1);
        }
#pragma unroll
        for (int f = 0; f < FPW; ++f)
#pragma unroll
          for (int q = 0; q < OPL; ++q) {
            const float* xb = total + (size_t)(f0 + f) * EF + q * T * 32 + lane;
            float t[T];
            float n2 = 0.f;
#pragma unroll
            for (int k = 0; k < T; ++k) {
              t[k] = xb[k * 32];
              n2 = fmaf(t[k], t[k], n2);
            }
            // n2/(1+n2) / sqrt(n2+eps) with approximate rsqrt / rcp (tensor-path tolerance class)
            const float scale = n2 * rsqrtf(n2 + 1e-7f) * fast_rcp(1.0f + n2);
            float* vo = vout + (size_t)(s & 1) * E + (size_t)(f0 + f) * EF + q * T * 32 + lane;
#pragma unroll
            for (int k = 0; k < T; ++k) {
              const float v = t[k] * scale;
              if (last_pass) {
                if (writer) vo[k * 32] = v;
                va[f][q][k] = p.sdr ? v : 0.f;  // SDR: next frame starts from this output (naive:167)
              } else {
                va[f][q][k] += v;
              }
            }
          }
        if (writer) {
          __syncwarp();
          if (lane == 0) {
#pragma unroll
            for (int f = 0; f < FPW; ++f) ptx::mbar_arrive(&vready[(f0 + f) * 2 + (s & 1)]);
          }
        }
#ifdef SRF_STREAM_PHASE_TIMERS
        if (timing) {
          const long long tk6 = clock64();
          unsigned long long* d = p.dbg + (size_t)blockIdx.x * 8;
          d[0] += (unsigned long long)(tk1 - tk0);  // capsule loop
          d[1] += (unsigned long long)(tk2 - tk1);  // store partials + CTA barrier
          d[2] += (unsigned long long)(tk3 - tk2);  // local sum + push
          d[3] += (unsigned long long)(tk4 - tk3);  // wait for the peers' partials
          d[4] += (unsigned long long)(tk5 - tk4);  // cluster sum + barrier
          d[5] += (unsigned long long)(tk6 - tk5);  // squash + hand-off
          d[6] += 1;
        }
#endif
#undef SRF_STK
      }
    }
  } else if (warp <= NCW + 2) {
    // ============================== output warps (one per pair member) ==============================
    const int f = warp - (NCW + 1);
    const bool do_ln = p.ln_gamma != nullptr;
    const bool do_head = p.head_gamma != nullptr;
    float gam[OPL][T], bet[OPL][T], hg[OPL], hb[OPL];
#pragma unroll
    for (int q = 0; q < OPL; ++q) {
      const int j = q * 32 + lane;
      hg[q] = (do_head && j < O) ? __ldg(p.head_gamma + j) : 1.f;
      hb[q] = (do_head && j < O) ? __ldg(p.head_beta + j) : 0.f;
#pragma unroll
      for (int k = 0; k < T; ++k) {
        const bool ok = do_ln && j < O && k < D;
        gam[q][k] = ok ? __ldg(p.ln_gamma + j * D + k) : 1.f;
        bet[q][k] = ok ? __ldg(p.ln_beta + j * D + k) : 0.f;
      }
    }
    const float inv_n = 1.0f / (float)(O * D);
    for (int s = 0; s < p.nsteps; ++s) {
      ptx::mbar_wait(&vready[f * 2 + (s & 1)], (s >> 1) & 1);
      const float* vf = vout + (size_t)(s & 1) * E + (size_t)f * EF;
      float y[OPL][T];
#pragma unroll
      for (int q = 0; q < OPL; ++q)
#pragma unroll
        for (int k = 0; k < T; ++k) y[q][k] = vf[(q * T + k) * 32 + lane];
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&vfree[f * 2 + (s & 1)]);

      long long frame;
      bool ok;
      if (p.sdr) {
        const int b = group * 2 + f;
        ok = b < p.B;
        frame = (long long)b * p.S + s;
      } else {
        const int ss = group / p.halfB, bb = 2 * (group % p.halfB) + f;
        ok = bb < p.B;
        frame = (long long)bb * p.S + ss;
      }
      if (!ok) continue;
      if (p.out_raw) {
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k)
            if (q * 32 + lane < O && k < D) p.out_raw[(frame * O + q * 32 + lane) * D + k] = y[q][k];
      }
      if (do_ln) {
        // padded capsules / dims hold exact zeros (zero weights -> t = 0 -> v = 0), so the sums
        // need no predicates; variance in one pass: E[y^2] - mean^2
        float sum = 0.f, sq = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k) {
            sum += y[q][k];
            sq = fmaf(y[q][k], y[q][k], sq);
          }
        const float mean = wsum(sum) * inv_n;
        const float var = fmaxf(wsum(sq) * inv_n - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.ln_eps);
#pragma unroll
        for (int q = 0; q < OPL; ++q)
#pragma unroll
          for (int k = 0; k < T; ++k) y[q][k] = (y[q][k] - mean) * rstd * gam[q][k] + bet[q][k];
      }
      float len[OPL];
#pragma unroll
      for (int q = 0; q < OPL; ++q) {
        const int j = q * 32 + lane;
        float l2 = 0.f;
        if (j < O) {
#pragma unroll
          for (int k = 0; k < T; ++k)
            if (k < D) {
              float v = y[q][k];
              if (p.dropout_mask) v *= __ldg(p.dropout_mask + (frame * O + j) * D + k);
              if (p.out_caps) p.out_caps[(frame * O + j) * D + k] = v;
              l2 = fmaf(v, v, l2);
            }
        }
        len[q] = sqrtf(l2 + p.length_eps);  // naive:256-258
      }
      if (do_head) {  // ln_output over the capsule lengths (naive:193)
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q)
          if (q * 32 + lane < O) sum += len[q];
        const float hm = wsum(sum) / (float)O;
        float sq = 0.f;
#pragma unroll
        for (int q = 0; q < OPL; ++q)
          if (q * 32 + lane < O) {
            const float dv = len[q] - hm;
            sq = fmaf(dv, dv, sq);
          }
        const float hr = 1.0f / sqrtf(wsum(sq) / (float)O + p.ln_eps);
#pragma unroll
        for (int q = 0; q < OPL; ++q) {
          const int j = q * 32 + lane;
          if (j < O) p.out_logits[frame * O + j] = (len[q] - hm) * hr * hg[q] + hb[q];
        }
      }
    }
  }

  // nobody leaves while a peer may still push into this CTA's shared memory
  __syncthreads();
  if (C > 1) cluster_sync_all();
}

// ---------------------------------------------------------------------------------------
template <int T, int OPL, int NSLOT, int FPW, bool BF16>
static cudaError_t launch_stream_variant(const RouteParams& p, int groups, size_t smem_bytes,
                                         cudaStream_t stream) {
  auto kern = route_stream_kernel<T, OPL, NSLOT, FPW, BF16>;
  cudaError_t err =
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (err != cudaSuccess) return err;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * p.C));
  cfg.blockDim = dim3((NSLOT * (2 / FPW) + 3) * 32);
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

// fixed (non-ring) shared memory of the streaming kernel
// capsule slots per ring stage (= consumer warps per pair member).  With one member per warp,
// 6 slots (15 warps, 136 registers, more ring stages) beat 8 slots (19 warps, 104 registers):
// measured on cfg-3 with 4 / 5 / 6 / 7 / 8 slots: 28.6 / 26.9 / 25.4 / 29.1 / 26.4 ms per step in
// bf16 storage, and 42.5 vs 54.9 ms (6 vs 8) in fp32 storage, where the 8-slot build spills.
// SRF_STREAM_NSLOT_BF16 / SRF_STREAM_NSLOT_F32 = 8 select the wider build.
int route_stream_nslot(int T, int OPL, bool bf16) {
  static const int f32_slots = [] {
    const char* e = getenv("SRF_STREAM_NSLOT_F32");
    const int v = e ? atoi(e) : 6;
    return v == 8 ? 8 : 6;
  }();
  static const int bf16_slots = [] {
    const char* e = getenv("SRF_STREAM_NSLOT_BF16");
    const int v = e ? atoi(e) : 6;
    return v == 8 ? 8 : 6;
  }();
  // two members per warp (large per-lane state): 4 slots make the D = 32 variant spill-free
  // (cfg-5 DIM=32: routing 53.0 -> 33.7 ms); the O > 32 variants measure the same or slightly
  // worse with 4 slots and stay at 8.  SRF_STREAM_NSLOT_FPW2 = 4 | 8 overrides.
  static const int fpw2_slots = [] {
    const char* e = getenv("SRF_STREAM_NSLOT_FPW2");
    const int v = e ? atoi(e) : 0;
    return (v == 4 || v == 8) ? v : 0;
  }();
  const bool one_member_per_warp = T * OPL <= 20;
  if (!one_member_per_warp) return fpw2_slots ? fpw2_slots : (T == 32 ? 4 : 8);
  if (bf16) return (one_member_per_warp && T * OPL > 8) ? bf16_slots : SRF_NW;
  return (one_member_per_warp && T * OPL > 8) ? f32_slots : SRF_NW;
}
size_t route_stream_fixed_smem(int T, int OPL, bool bf16, int C, int max_stages) {
  const size_t E = (size_t)2 * OPL * T * 32;
  return sizeof(float) * ((size_t)route_stream_nslot(T, OPL, bf16) * E + (size_t)2 * C * E + 3 * E) +
         sizeof(uint64_t) * (2 * (size_t)max_stages + 12) + 128;
}
size_t route_stream_stage_bytes(int T, int OPL, bool bf16) {
  return (size_t)route_stream_nslot(T, OPL, bf16) * OPL * (T / 4) * 128 * 2 * (bf16 ? 2 : 4);
}

// variants with a per-lane state of <= 20 floats per pair member run one member per warp
#define SRF_STREAM(T_, OPL_, FPW_)                                                                  \
  if (T == T_ && OPL == OPL_) {                                                                     \
    if (route_stream_nslot(T_, OPL_, bf16) == 4 && FPW_ == 2)                                       \
      return bf16 ? launch_stream_variant<T_, OPL_, 4, 2, true>(p, groups, smem_bytes, stream)      \
                  : launch_stream_variant<T_, OPL_, 4, 2, false>(p, groups, smem_bytes, stream);    \
    if (bf16 && route_stream_nslot(T_, OPL_, true) == 6)                                            \
      return launch_stream_variant<T_, OPL_, 6, FPW_, true>(p, groups, smem_bytes, stream);         \
    if (bf16) return launch_stream_variant<T_, OPL_, SRF_NW, FPW_, true>(p, groups, smem_bytes, stream); \
    if (route_stream_nslot(T_, OPL_, false) == 6)                                                   \
      return launch_stream_variant<T_, OPL_, 6, FPW_, false>(p, groups, smem_bytes, stream);        \
    return launch_stream_variant<T_, OPL_, SRF_NW, FPW_, false>(p, groups, smem_bytes, stream);     \
  }

cudaError_t launch_route_stream(const RouteParams& p, int T, int OPL, bool bf16, int groups,
                                size_t smem_bytes, cudaStream_t stream) {
  SRF_STREAM(8, 1, 1)
  SRF_STREAM(8, 2, 1)
  SRF_STREAM(8, 4, 2)
  SRF_STREAM(16, 1, 1)
  SRF_STREAM(16, 2, 2)
  SRF_STREAM(20, 1, 1)
  SRF_STREAM(20, 2, 2)
  SRF_STREAM(32, 1, 2)
  return cudaErrorInvalidValue;
}

}  // namespace srf
