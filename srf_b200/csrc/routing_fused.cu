// routing_fused.cu -- the fused routing kernel of round 2 (sm_100a): prediction vectors u_hat are
// produced by tcgen05.mma straight into TMEM and consumed from there by the routing math; u_hat
// never exists in HBM (reference: tfsr/model/sequence_router_naive.py:150-193 -- window gather,
// u_hat = W.x + bias, SDR scan / DR iterations, squash, LayerNorm + dropout, head).
//
// Decomposition.  A GROUP is 32 frames that share every MMA: for SDR the
// same time step of 32 utterances, for DR a (utterances x time steps) tile.  A UNIT is one group of
// one layer; it is owned by C CTAs that split the layer's input capsules.  One persistent launch
// runs, for SDR, ALL layers of the stack as a wavefront (unit (l, g) at step s needs layer l-1 up
// to step s+RPAD: per-frame progress flags in global memory), so with 64 utterances and 10 layers
// 148 SMs are busy although every chain is strictly sequential in time.
//
// Per CTA and input capsule i:
//   warp 0   TMA producer: W[i] arrives tile by tile (128 rows x K, pre-packed in the shared-memory
//            operand image) through a ring of bulk copies from L2 (the weights are re-streamed every
//            time step: 60 B/clk/SM measured, profiles/r2_ubench.txt); the x tile (32 frames x K)
//            comes by a 5-D tensor-map load whose out-of-bounds zero fill is the window padding
//            (naive:150).
//   warp 1   MMA issuer: D[tile m] = W[i] tile m (A, M = 128 rows) . x^T (B, N = 32 frames), kind::tf32,
//            accumulators in TMEM, NT tiles per capsule, 2-4 capsules in flight.  The bias rides
//            in the K padding (x carries a constant 1 in column d).  X3: 3 x TF32 split
//            (W_hi x_hi + W_lo x_hi + W_hi x_lo) for fp32-class u_hat.
//   warp 2   X3 only: splits each x tile into x_hi / x_lo in shared memory.
//   warps 4-11  routing math, two TEAMS of 4 warps x 16 frames.  Row r of tile m holds
//            (j = jb*32 + r%32, k = 4*k4 + r/32) with m = jb*T4 + k4, so warp q of a team (TMEM lane
//            quarter q) owns k = q mod 4 for every output capsule j = lane, and a thread keeps
//            t[jb][k4][16 frames] and Vacc[jb][k4][16 frames] in registers.  Per capsule:
//            partial logits over the warp's own k -> exchanged through shared memory -> each warp
//            does the softmax over j (lanes, REDUX) for 4 of the 16 frames -> coefficients back
//            through shared memory -> t += c * u_hat with u_hat re-read from TMEM.
// Per step (SDR) / pass (ITER > 1, DR): the C partial t of a unit meet in an L2-resident exchange
// buffer; one warp per frame (lane = j) sums them, squashes (naive:248-253), publishes v for the
// next pass / step and, after the last pass, applies LayerNorm + dropout (naive:188-191), the head
// (naive:193) and stores the frame.  DSMEM was measured at 17 B/clk/CTA -- too slow for this.
//
// Every wait is bounded: on a timeout the kernel sets an abort flag, all waits stop blocking, the
// launch drains and the host reports an error instead of hanging the GPU.

#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "routing_kernels.h"
#include "sm100_ptx.cuh"

namespace srf {

// ---------------------------------------------------------------------------------------
// weight packing:  W[I,O,D,d], bias[I,O,D] -> Wf float[i][part][m][c][r][4]
//   m = jb*T4 + k4, r = (k%4)*32 + j%32 (j = jb*32 + r%32, k = 4*k4 + r/32), c = 16-byte K chunk;
//   element (c, li) is l = 4c + li: l < d -> W[i,j,k,l];  l == d -> bias[i,j,k];  else 0.
//   part 0 = rna_tf32(value), part 1 (X3) = rna_tf32(value - part0)
// ---------------------------------------------------------------------------------------
__global__ void pack_weights_fused_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                          float* __restrict__ Wf, int I, int O, int D, int d, int T4,
                                          int OPL, int KC, int parts) {
  const int NT = OPL * T4;
  const long long n = (long long)I * parts * NT * KC * 512;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int li = (int)(e & 3);
    long long t = e >> 2;
    const int r = (int)(t & 127);
    t >>= 7;
    const int c = (int)(t % KC);
    t /= KC;
    const int m = (int)(t % NT);
    t /= NT;
    const int part = (int)(t % parts);
    const int i = (int)(t / parts);
    const int jb = m / T4, k4 = m - jb * T4;
    const int j = jb * 32 + (r & 31), k = 4 * k4 + (r >> 5), l = 4 * c + li;
    float v = 0.f;
    if (j < O && k < D) {
      if (l < d)
        v = W[(((long long)i * O + j) * D + k) * d + l];
      else if (l == d)
        v = bias[((long long)i * O + j) * D + k];
    }
    uint32_t tf;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tf) : "f"(v));
    if (part == 1) {
      const float lo = v - __uint_as_float(tf);
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tf) : "f"(lo));
    }
    Wf[e] = __uint_as_float(tf);
  }
}

void launch_pack_weights_fused(const float* W, const float* bias, float* Wf, int I, int O, int D, int d,
                               int T4, int OPL, int KC, int parts, cudaStream_t stream) {
  const long long n = (long long)I * parts * OPL * T4 * KC * 512;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  pack_weights_fused_kernel<<<blocks, 256, 0, stream>>>(W, bias, Wf, I, O, D, d, T4, OPL, KC, parts);
}

namespace {

constexpr int FZ_N = 32;                 // frames per group = MMA N
constexpr int FZ_TF = 16;                // frames per team
constexpr int FZ_MATH_WARPS = 8;
constexpr int FZ_THREADS = 384;          // 4 service warps + 8 math warps
constexpr int FZ_XST = 4;                // x-tile ring depth
constexpr long long FZ_TIMEOUT = 6000000000ll;  // ~3 s of SM clocks

enum { BAR_TEAM0 = 1, BAR_MATH = 5 };

__device__ __forceinline__ void named_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
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
__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() {
  asm volatile("fence.proxy.async;" ::: "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}
// TMEM -> registers, 16 consecutive columns of this thread's lane; the wait takes the registers as
// read-write operands so that no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
        "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]),
        "=f"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait16(float (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]),
                 "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]),
                 "+f"(r[14]), "+f"(r[15])::"memory");
}

// bounded waits.  `dead` is the CTA-local copy of the abort state: once set nothing blocks any more.
struct Waiter {
  int* abort_flag;      // global
  volatile int* host_abort;  // mapped host memory: read by the host without a synchronise
  volatile int* dead;   // shared
  __device__ __forceinline__ bool is_dead() const { return *dead != 0; }
  __device__ __noinline__ void fail(int code) {
    *dead = 1;
    if (atomicCAS(abort_flag, 0, code) == 0 && host_abort) {
      *host_abort = code;
      __threadfence_system();
    }
  }
  __device__ __forceinline__ void mbar(uint64_t* bar, uint32_t parity, int code) {
    if (ptx::mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    unsigned n = 0;
    while (!ptx::mbar_try_wait(bar, parity)) {
      if ((++n & 255u) == 0) {
        if (*dead) return;
        if (ld_acquire(abort_flag) != 0) {
          *dead = 1;
          return;
        }
        if (clock64() - t0 > FZ_TIMEOUT) {
          fail(code);
          return;
        }
      }
    }
  }
  // wait until *p >= target (global counter written by other CTAs)
  __device__ __forceinline__ void counter(const int* p, int target, int code) {
    if (ld_acquire(p) >= target) return;
    const long long t0 = clock64();
    unsigned n = 0;
    while (ld_acquire(p) < target) {
      if ((++n & 63u) == 0) {
        if (*dead) return;
        if (ld_acquire(abort_flag) != 0) {
          *dead = 1;
          return;
        }
        if (clock64() - t0 > FZ_TIMEOUT) {
          fail(code);
          return;
        }
      }
    }
  }
};

}  // namespace

// T4 = ceil(D/4) tiles per block of 32 output capsules, OPL = blocks of 32 output capsules the
// build holds (a layer may use fewer: FusedLayer::opl), X3 = 3 x TF32 split
template <int T4, int OPL, bool X3>
__global__ void __launch_bounds__(FZ_THREADS, 1) route_fused_kernel(const FusedParams p) {
  constexpr int NT = T4 * OPL;        // M tiles per input capsule (at most)
  constexpr int T = 4 * T4;           // padded output capsule dim
  constexpr int OP = 32 * OPL;        // padded output capsules
  constexpr int TCOLS = NT * FZ_N;    // TMEM columns per capsule buffer
  constexpr int NBUF = (512 / TCOLS) > 4 ? 4 : (512 / TCOLS);
  constexpr float LOG2E = 1.4426950408889634f;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // a CTA serves ONE layer for the whole launch (host invariant): its operand geometry is fixed
  int my_layer = -1;
  for (int it = 0; it < p.rounds && my_layer < 0; ++it) my_layer = p.items[(size_t)it * gridDim.x + blockIdx.x].layer;
  const int KC = my_layer >= 0 ? p.layers[my_layer].KC : 2;   // 16-byte K chunks per operand row
  const int KX = my_layer >= 0 ? p.layers[my_layer].KX : 1;   // of which the TMA box fills KX
  const uint32_t wtile = (uint32_t)KC * 2048u;           // one image of one W tile
  const uint32_t wstage = X3 ? 2u * wtile : wtile;       // hi (+ lo)
  const uint32_t xtile = (uint32_t)KC * FZ_N * 16u;      // one image of one x tile
  const int NWST = p.nwst;

  uint8_t* sW = smem_raw;                                     // [NWST][wstage]
  uint8_t* sX = sW + (size_t)NWST * wstage;                   // [XST][xtile]  (x, or x_hi)
  uint8_t* sXlo = sX + (size_t)FZ_XST * xtile;                // [XST][xtile]  (X3)
  float* sP = reinterpret_cast<float*>(sXlo + (X3 ? (size_t)FZ_XST * xtile : 0));  // [2 teams][4][OPL][32][16]
  float* sC = sP + 2 * 4 * OPL * 32 * 16;                     // [2 teams][OPL][32][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + 2 * OPL * 32 * 16);
  uint64_t* w_full = bars;                   // [NWST]
  uint64_t* w_empty = w_full + NWST;         // [NWST]
  uint64_t* x_full = w_empty + NWST;         // [XST]
  uint64_t* x_empty = x_full + FZ_XST;       // [XST]
  uint64_t* x_split = x_empty + FZ_XST;      // [XST]
  uint64_t* t_full = x_split + FZ_XST;       // [NBUF]
  uint64_t* t_empty = t_full + NBUF;         // [NBUF]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(t_empty + NBUF);
  volatile int* s_dead = reinterpret_cast<volatile int*>(tmem_ptr + 1);

  Waiter wt{p.abort_flag, p.host_abort, s_dead};

  if (tid == 0) {
    *s_dead = 0;
    for (int s = 0; s < NWST; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < FZ_XST; ++s) {
      ptx::mbar_init(&x_full[s], 1);
      ptx::mbar_init(&x_empty[s], 1);
      ptx::mbar_init(&x_split[s], 1);
    }
    for (int s = 0; s < NBUF; ++s) {
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], FZ_MATH_WARPS);
    }
    ptx::fence_barrier_init();
  }
  // constant part of the x tiles: chunk KX carries the 1 that multiplies the bias column of W
  // (x_hi image), everything above the TMA box is zero; the TMA never writes these chunks
  for (int e = tid; e < FZ_XST * (KC - KX) * FZ_N; e += FZ_THREADS) {
    const int st = e / ((KC - KX) * FZ_N), rem = e - st * ((KC - KX) * FZ_N);
    const int c = KX + rem / FZ_N, f = rem % FZ_N;
    float4* dst = reinterpret_cast<float4*>(sX + (size_t)st * xtile + ((size_t)c * FZ_N + f) * 16);
    *dst = make_float4(c == KX ? 1.f : 0.f, 0.f, 0.f, 0.f);
    if (X3) *reinterpret_cast<float4*>(sXlo + (size_t)st * xtile + ((size_t)c * FZ_N + f) * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ptx::fence_proxy_async();
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int iters = p.iters;
  const int nsteps = p.sdr ? p.S : 1;

  // register budget: the routing warps hold 2 x NT x 16 floats of state per thread
  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (warp == 0) {
    // =========================== TMA producer ===========================
    uint32_t n_w = 0, n_x = 0;
    for (int it = 0; it < p.rounds; ++it) {
      const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
      if (item.layer < 0) continue;
      const FusedLayer* L = p.layers + item.layer;
      const int H = L->H, lpad = L->lpad;
      const int nt = L->opl * T4;
      const int b0 = item.b0, s00 = item.s0;
      const int* prog = L->dep_layer >= 0
                            ? p.progress + ((size_t)L->dep_layer * p.ngroups + item.group) * FZ_N
                            : nullptr;
      const bool my_frame = ((item.vmask >> lane) & 1u) != 0;
      // the tensor map lives in global memory (written by the host before the launch)
      asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(&L->tmap) : "memory");
      for (int step = 0; step < nsteps; ++step) {
        const int s0 = p.sdr ? step : s00;
        if (prog) {
          // the producing layer must have stored every frame of the window of this step
          int need = step + L->rpad + 1;
          if (need > p.S) need = p.S;
          if (my_frame) wt.counter(prog + lane, need, 101);
          __syncwarp();
          fence_proxy_async_all();
        }
        if (lane == 0) {
          for (int pass = 0; pass < iters; ++pass) {
            for (int i = item.i_lo; i < item.i_hi; ++i) {
              const int xs = n_x % FZ_XST;
              wt.mbar(&x_empty[xs], ((n_x / FZ_XST) & 1) ^ 1, 102);
              ptx::mbar_arrive_expect_tx(&x_full[xs], (uint32_t)KX * FZ_N * 16u);
              const int w = i / H, h = i - w * H;
              ptx::tma_load_5d(sX + (size_t)xs * xtile, &L->tmap, &x_full[xs], 0, b0, s0 - lpad + w, h, 0);
              ++n_x;
              const uint8_t* src = reinterpret_cast<const uint8_t*>(L->Wf) + (size_t)i * nt * wstage;
              for (int m = 0; m < nt; ++m) {
                const int ws = n_w % NWST;
                wt.mbar(&w_empty[ws], ((n_w / NWST) & 1) ^ 1, 103);
                ptx::mbar_arrive_expect_tx(&w_full[ws], wstage);
                uint8_t* dst = sW + (size_t)ws * wstage;
                if (X3) {
                  // packed as [i][part][m]: the hi and lo images of tile m are NT tiles apart
                  ptx::bulk_g2s(dst, src + (size_t)m * wtile, wtile, &w_full[ws]);
                  ptx::bulk_g2s(dst + wtile, src + (size_t)(nt + m) * wtile, wtile, &w_full[ws]);
                } else {
                  ptx::bulk_g2s(dst, src + (size_t)m * wtile, wtile, &w_full[ws]);
                }
                ++n_w;
              }
            }
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_tf32(128, FZ_N);
      const uint32_t sW_a = ptx::smem_u32(sW), sX_a = ptx::smem_u32(sX), sXlo_a = ptx::smem_u32(sXlo);
      uint32_t n_w = 0, n_x = 0, n_t = 0;
      for (int it = 0; it < p.rounds; ++it) {
        const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
        if (item.layer < 0) continue;
        const long long ncaps = (long long)nsteps * iters * (item.i_hi - item.i_lo);
        const int nt = p.layers[item.layer].opl * T4;
        for (long long cc = 0; cc < ncaps; ++cc) {
          const int xs = n_x % FZ_XST;
          wt.mbar(X3 ? &x_split[xs] : &x_full[xs], (n_x / FZ_XST) & 1, 104);
          const int buf = n_t % NBUF;
          wt.mbar(&t_empty[buf], ((n_t / NBUF) & 1) ^ 1, 105);
          ptx::tc_fence_after();
          for (int m = 0; m < nt; ++m) {
            const int ws = n_w % NWST;
            wt.mbar(&w_full[ws], (n_w / NWST) & 1, 106);
            ptx::tc_fence_after();
            const uint32_t d_addr = tmem_base + (uint32_t)(buf * TCOLS + m * FZ_N);
            const uint32_t a0 = sW_a + (uint32_t)ws * wstage;
            for (int ks = 0; ks < KC / 2; ++ks) {
              const uint64_t adesc = ptx::make_smem_desc(a0 + (uint32_t)ks * 4096u, 2048u, 128u);
              const uint64_t bdesc = ptx::make_smem_desc(
                  sX_a + (uint32_t)xs * xtile + (uint32_t)ks * (2u * FZ_N * 16u), FZ_N * 16u, 128u);
              ptx::mma_tf32_ss(d_addr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
              if (X3) {
                const uint64_t adesc_lo =
                    ptx::make_smem_desc(a0 + wtile + (uint32_t)ks * 4096u, 2048u, 128u);
                const uint64_t bdesc_lo = ptx::make_smem_desc(
                    sXlo_a + (uint32_t)xs * xtile + (uint32_t)ks * (2u * FZ_N * 16u), FZ_N * 16u, 128u);
                ptx::mma_tf32_ss(d_addr, adesc_lo, bdesc, idesc, 1u);
                ptx::mma_tf32_ss(d_addr, adesc, bdesc_lo, idesc, 1u);
              }
            }
            ptx::mma_commit(&w_empty[ws]);
            ++n_w;
          }
          ptx::mma_commit(&t_full[buf]);
          ptx::mma_commit(&x_empty[xs]);
          ++n_t;
          ++n_x;
        }
      }
    }
  } else if (warp == 2) {
    // =========================== x splitter (X3) ===========================
    if (X3) {
      uint32_t n_x = 0;
      const int n4 = KX * FZ_N;  // float4 per tile inside the TMA box
      for (int it = 0; it < p.rounds; ++it) {
        const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
        if (item.layer < 0) continue;
        const long long ncaps = (long long)nsteps * iters * (item.i_hi - item.i_lo);
        for (long long cc = 0; cc < ncaps; ++cc) {
          const int xs = n_x % FZ_XST;
          wt.mbar(&x_full[xs], (n_x / FZ_XST) & 1, 107);
          uint4* px = reinterpret_cast<uint4*>(sX + (size_t)xs * xtile);
          float4* pl = reinterpret_cast<float4*>(sXlo + (size_t)xs * xtile);
          for (int e = lane; e < n4; e += 32) {
            const uint4 v = px[e];
            uint4 hi;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi.x) : "f"(__uint_as_float(v.x)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi.y) : "f"(__uint_as_float(v.y)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi.z) : "f"(__uint_as_float(v.z)));
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi.w) : "f"(__uint_as_float(v.w)));
            px[e] = hi;
            pl[e] = make_float4(__uint_as_float(v.x) - __uint_as_float(hi.x),
                                __uint_as_float(v.y) - __uint_as_float(hi.y),
                                __uint_as_float(v.z) - __uint_as_float(hi.z),
                                __uint_as_float(v.w) - __uint_as_float(hi.w));
          }
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&x_split[xs]);
          ++n_x;
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =========================== routing math ===========================
    const int mw = warp - 4;          // 0..7
    const int team = mw >> 2;         // frames team*16 .. +15
    const int q = mw & 3;             // TMEM lane quarter == warp % 4; owns k = q mod 4
    const int mtid = tid - 128;       // 0..255
    const int bar_a = BAR_TEAM0 + team * 2, bar_b = bar_a + 1;
    // swizzled exchange rows: 16 floats (64 B) per (jb, lane); the 16-byte chunk fq is stored at
    // position fq ^ ((lane >> 1) & 3) so that 8 lanes of a store / load phase hit 32 distinct banks
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);
    const uint32_t sP_team = ptx::smem_u32(sP) + (uint32_t)team * (4 * OPL * 32 * 64);
    const uint32_t sC_team = ptx::smem_u32(sC) + (uint32_t)team * (OPL * 32 * 64);
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)team * FZ_TF;

    uint32_t n_t = 0;
    int epoch = 0;          // passes completed by this CTA (all CTAs of a slot agree)
    int v_target = 0;       // running target of this team's frame counter

    for (int it = 0; it < p.rounds; ++it) {
      const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
      if (item.layer < 0) continue;
      const FusedLayer* L = p.layers + item.layer;
      const int O = L->O, D = L->D;
      const int nt = L->opl * T4;
      const int C = item.C, c = item.c;
      const uint32_t team_mask = (item.vmask >> (team * FZ_TF)) & 0xffffu;
      const int team_frames = __popc(team_mask);
      float* const Pbuf = p.xP + (size_t)item.slot * p.maxC * FZ_N * T * OP;   // [C][32][T][OP]
      float* const Vbuf = p.xV + (size_t)item.slot * FZ_N * T * OP;            // [32][T][OP]
      int* const cnt_p = p.cnt_p + item.slot;
      int* const cnt_v = p.cnt_v + item.slot * 2;
      const bool do_ln = L->ln_gamma != nullptr;
      const bool do_head = L->head_gamma != nullptr;

      float va[OPL][T4][FZ_TF];
#pragma unroll
      for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
        for (int k4 = 0; k4 < T4; ++k4)
#pragma unroll
          for (int f = 0; f < FZ_TF; ++f) va[jb][k4][f] = 0.f;

      for (int step = 0; step < nsteps; ++step) {
        for (int pass = 0; pass < iters; ++pass) {
          const bool last_pass = pass == iters - 1;
          float ta[OPL][T4][FZ_TF];
#pragma unroll
          for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
            for (int k4 = 0; k4 < T4; ++k4)
#pragma unroll
              for (int f = 0; f < FZ_TF; ++f) ta[jb][k4][f] = 0.f;

          for (int i = item.i_lo; i < item.i_hi; ++i) {
            const int buf = n_t % NBUF;
            wt.mbar(&t_full[buf], (n_t / NBUF) & 1, 110);
            ptx::tc_fence_after();
            const uint32_t tb = tmem_lane + (uint32_t)(buf * TCOLS);
            // ---- phase A: partial logits over this warp's k (naive:205 / :223 / :240) ----
            float pl[OPL][FZ_TF];
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int f = 0; f < FZ_TF; ++f) pl[jb][f] = 0.f;
            {
              float u0[16], u1[16];
              tmem_ld16f(tb, u0);
#pragma unroll
              for (int m = 0; m < NT; ++m) {
                if (m >= nt) break;
                const int jb = m / T4, k4 = m % T4;
                if (m & 1) tmem_wait16(u1); else tmem_wait16(u0);
                if (m + 1 < nt) {
                  if (m & 1) tmem_ld16f(tb + (uint32_t)((m + 1) * FZ_N), u0);
                  else tmem_ld16f(tb + (uint32_t)((m + 1) * FZ_N), u1);
                }
#pragma unroll
                for (int f = 0; f < FZ_TF; ++f) {
                  const float uu = (m & 1) ? u1[f] : u0[f];
                  pl[jb][f] = fmaf(uu, va[jb][k4][f], pl[jb][f]);
                }
              }
            }
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb) {
              const uint32_t row = sP_team + (uint32_t)(((q * OPL + jb) * 32 + lane) * 64);
#pragma unroll
              for (int fq = 0; fq < 4; ++fq)
                sts128f(row + (((uint32_t)fq ^ swz) << 4), pl[jb][4 * fq], pl[jb][4 * fq + 1],
                        pl[jb][4 * fq + 2], pl[jb][4 * fq + 3]);
            }
            named_sync(bar_a, 128);
            // ---- coupling softmax over the output capsules for frames 4q .. 4q+3 of the team ----
            {
              float a4[OPL][4];
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                  const float4 x = lds128f(sP_team + (uint32_t)(((qq * OPL + jb) * 32 + lane) * 64) +
                                           (((uint32_t)q ^ swz) << 4));
                  acc.x += x.x;
                  acc.y += x.y;
                  acc.z += x.z;
                  acc.w += x.w;
                }
                const int j = jb * 32 + lane;
                const bool valid = (j < O) && !(L->mask0 && j == 0);
                a4[jb][0] = valid ? acc.x : -CUDART_INF_F;
                a4[jb][1] = valid ? acc.y : -CUDART_INF_F;
                a4[jb][2] = valid ? acc.z : -CUDART_INF_F;
                a4[jb][3] = valid ? acc.w : -CUDART_INF_F;
              }
              float c4[OPL][4];
#pragma unroll
              for (int ff = 0; ff < 4; ++ff) {
                float m = a4[0][ff];
#pragma unroll
                for (int jb = 1; jb < OPL; ++jb) m = fmaxf(m, a4[jb][ff]);
                int mi = __float_as_int(m);
                mi ^= (mi >> 31) & 0x7fffffff;
                mi = __reduce_max_sync(0xffffffffu, mi);
                mi ^= (mi >> 31) & 0x7fffffff;
                m = __int_as_float(mi);
                float ex[OPL];
                float zl = 0.f;
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb) {
                  ex[jb] = fast_ex2((a4[jb][ff] - m) * LOG2E);
                  zl += ex[jb];
                }
                // fixed-point normaliser (OPL terms <= 1 per lane, 32 lanes: < 2^31 in Q(26 - log2 OPL))
                constexpr float QS = (float)(1 << 26) / (float)OPL;
                const unsigned zi = __reduce_add_sync(0xffffffffu, __float2uint_rn(zl * QS));
                const float inv = fast_rcp((float)zi * (1.0f / QS));
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb) c4[jb][ff] = ex[jb] * inv;
              }
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
                sts128f(sC_team + (uint32_t)((jb * 32 + lane) * 64) + (((uint32_t)q ^ swz) << 4),
                        c4[jb][0], c4[jb][1], c4[jb][2], c4[jb][3]);
            }
            named_sync(bar_b, 128);
            // ---- phase B: t += c * u_hat, u_hat re-read from TMEM (naive:203 / :226 / :242) ----
            {
              float cf[OPL][FZ_TF];
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int fq = 0; fq < 4; ++fq) {
                  const float4 x =
                      lds128f(sC_team + (uint32_t)((jb * 32 + lane) * 64) + (((uint32_t)fq ^ swz) << 4));
                  cf[jb][4 * fq] = x.x;
                  cf[jb][4 * fq + 1] = x.y;
                  cf[jb][4 * fq + 2] = x.z;
                  cf[jb][4 * fq + 3] = x.w;
                }
              float u0[16], u1[16];
              tmem_ld16f(tb, u0);
#pragma unroll
              for (int m = 0; m < NT; ++m) {
                if (m >= nt) break;
                const int jb = m / T4, k4 = m % T4;
                if (m & 1) tmem_wait16(u1); else tmem_wait16(u0);
                if (m + 1 < nt) {
                  if (m & 1) tmem_ld16f(tb + (uint32_t)((m + 1) * FZ_N), u0);
                  else tmem_ld16f(tb + (uint32_t)((m + 1) * FZ_N), u1);
                } else {
                  // last read of this capsule buffer: hand it back to the MMA issuer
                  ptx::tc_fence_before();
                  __syncwarp();
                  if (lane == 0) ptx::mbar_arrive(&t_empty[buf]);
                }
#pragma unroll
                for (int f = 0; f < FZ_TF; ++f) {
                  const float uu = (m & 1) ? u1[f] : u0[f];
                  ta[jb][k4][f] = fmaf(cf[jb][f], uu, ta[jb][k4][f]);
                }
              }
            }
            ++n_t;
          }

          // ================= end of pass: exchange the partial sums through L2 =================
          ++epoch;
          v_target += team_frames;
          {
            float* mine = Pbuf + (size_t)c * FZ_N * T * OP;
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int k4 = 0; k4 < T4; ++k4)
#pragma unroll
                for (int f = 0; f < FZ_TF; ++f)
                  mine[((size_t)(team * FZ_TF + f) * T + (4 * k4 + q)) * OP + jb * 32 + lane] = ta[jb][k4][f];
          }
          named_sync(BAR_MATH, 256);
          if (mtid == 0) {
            __threadfence();
            atomicAdd(cnt_p, 1);
          }
          // ---- frame owners: one warp per frame, lane = output capsule ----
          for (int f = c + C * mw; f < FZ_N; f += C * FZ_MATH_WARPS) {
            if (!((item.vmask >> f) & 1u)) continue;
            if (lane == 0) wt.counter(cnt_p, epoch * C, 111);
            __syncwarp();
            float y[OPL][T];
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int k = 0; k < T; ++k) y[jb][k] = 0.f;
            for (int cc = 0; cc < C; ++cc) {
              const float* src = Pbuf + ((size_t)cc * FZ_N + f) * T * OP + lane;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) y[jb][k] += __ldcg(src + k * OP + jb * 32);
            }
            // squash (naive:248-253)
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb) {
              float n2 = 0.f;
#pragma unroll
              for (int k = 0; k < T; ++k) n2 = fmaf(y[jb][k], y[jb][k], n2);
              const float scale = X3 ? (n2 / (1.0f + n2)) / sqrtf(n2 + 1e-7f)
                                     : n2 * rsqrtf(n2 + 1e-7f) * fast_rcp(1.0f + n2);
#pragma unroll
              for (int k = 0; k < T; ++k) y[jb][k] *= scale;
            }
            if (!(last_pass && !p.sdr)) {
              // the next pass / step needs v (DR's last pass does not: its Vacc is reset)
              float* dst = Vbuf + (size_t)f * T * OP + lane;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) dst[k * OP + jb * 32] = y[jb][k];
            }
            __syncwarp();
            if (lane == 0) {
              __threadfence();
              atomicAdd(cnt_v + (f >> 4), 1);
            }
            if (!last_pass) continue;
            // ---- LayerNorm + dropout (naive:188-191), head (naive:193), stores ----
            const int fb = f % p.NB, fs = f / p.NB;
            const int b = item.b0 + fb, s = (p.sdr ? step : item.s0) + fs;
            const long long frame = (long long)b * p.S + s;
            if (L->out_raw) {
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k)
                  if (jb * 32 + lane < O && k < D) L->out_raw[(frame * O + jb * 32 + lane) * D + k] = y[jb][k];
            }
            if (do_ln) {
              float sum = 0.f, sq = 0.f;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) {
                  sum += y[jb][k];
                  sq = fmaf(y[jb][k], y[jb][k], sq);
                }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) {
                sum += __shfl_xor_sync(0xffffffffu, sum, o);
                sq += __shfl_xor_sync(0xffffffffu, sq, o);
              }
              const float inv_n = 1.0f / (float)(O * D);
              const float mean = sum * inv_n;
              float var;
              if (X3) {
                // exact class: two-pass variance (padded entries are exact zeros: remove their share)
                float s2 = 0.f;
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                  for (int k = 0; k < T; ++k) {
                    const float dv = (jb * 32 + lane < O && k < D) ? y[jb][k] - mean : 0.f;
                    s2 = fmaf(dv, dv, s2);
                  }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                var = s2 * inv_n;
              } else {
                var = fmaxf(sq * inv_n - mean * mean, 0.f);
              }
              const float rstd = X3 ? 1.0f / sqrtf(var + L->ln_eps) : rsqrtf(var + L->ln_eps);
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) {
                  const int j = jb * 32 + lane;
                  const bool ok = j < O && k < D;
                  const float g = ok ? __ldg(L->ln_gamma + j * D + k) : 0.f;
                  const float be = ok ? __ldg(L->ln_beta + j * D + k) : 0.f;
                  y[jb][k] = (y[jb][k] - mean) * rstd * g + be;
                }
            }
            float len[OPL];
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb) {
              const int j = jb * 32 + lane;
              float l2 = 0.f;
              if (j < O) {
#pragma unroll
                for (int k = 0; k < T; ++k)
                  if (k < D) {
                    float v = y[jb][k];
                    if (L->dropout_mask) v *= __ldg(L->dropout_mask + (frame * O + j) * D + k);
                    if (L->out_caps) L->out_caps[(frame * O + j) * D + k] = v;
                    l2 = fmaf(v, v, l2);
                  }
              }
              len[jb] = sqrtf(l2 + L->length_eps);  // naive:256-258
            }
            if (do_head) {  // ln_output over the capsule lengths (naive:193)
              float sum = 0.f;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
                if (jb * 32 + lane < O) sum += len[jb];
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
              const float hm = sum / (float)O;
              float sq = 0.f;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
                if (jb * 32 + lane < O) {
                  const float dv = len[jb] - hm;
                  sq = fmaf(dv, dv, sq);
                }
#pragma unroll
              for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
              const float hr = 1.0f / sqrtf(sq / (float)O + L->ln_eps);
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb) {
                const int j = jb * 32 + lane;
                if (j < O)
                  L->out_logits[frame * O + j] =
                      (len[jb] - hm) * hr * __ldg(L->head_gamma + j) + __ldg(L->head_beta + j);
              }
            }
            if (p.progress && p.sdr) {
              // publish: frame f of this unit is stored up to and including `step`
              __syncwarp();
              if (lane == 0) {
                __threadfence();
                fence_proxy_async_all();
                st_release(p.progress + ((size_t)item.layer * p.ngroups + item.group) * FZ_N + f, step + 1);
              }
            }
          }
          // ---- everybody: fetch v of the team's frames, update Vacc ----
          // (always wait: the owners must be done with this team's rows of the exchange buffer
          // before the next pass overwrites them)
          if (lane == 0) wt.counter(cnt_v + team, v_target, 112);
          __syncwarp();
          if (last_pass && !p.sdr) {
            // DR: frames are independent, the next item starts from Vacc = 0
          } else {
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int k4 = 0; k4 < T4; ++k4)
#pragma unroll
                for (int f = 0; f < FZ_TF; ++f) {
                  const float v =
                      __ldcg(Vbuf + ((size_t)(team * FZ_TF + f) * T + (4 * k4 + q)) * OP + jb * 32 + lane);
                  // SDR: the next frame starts from this output (naive:167); ITER > 1: logits are
                  // linear in the accumulated outputs
                  va[jb][k4][f] = last_pass ? v : va[jb][k4][f] + v;
                }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
size_t route_fused_smem_bytes(int OPL, int KC, int x3, int nwst) {
  const size_t wstage = (size_t)KC * 2048 * (x3 ? 2 : 1);
  const size_t xt = (size_t)KC * FZ_N * 16;
  return 1024 + (size_t)nwst * wstage + (size_t)FZ_XST * xt * (x3 ? 2 : 1) +
         sizeof(float) * (2 * 4 * OPL * 32 * 16 + 2 * OPL * 32 * 16) +
         sizeof(uint64_t) * (2 * (size_t)nwst + 3 * FZ_XST + 8) + 64;
}

template <int T4, int OPL, bool X3>
static cudaError_t launch_fused_variant(const FusedParams& p, int grid, size_t smem, cudaStream_t stream) {
  auto kern = route_fused_kernel<T4, OPL, X3>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(FZ_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait on one another
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

bool route_fused_supported(int T4, int OPL) {
  return (T4 == 2 && (OPL == 1 || OPL == 2)) || (T4 == 4 && OPL == 1) || (T4 == 5 && OPL == 1);
}

cudaError_t launch_route_fused(const FusedParams& p, int T4, int OPL, int x3, int grid, size_t smem,
                               cudaStream_t stream) {
#define SRF_FUSED(T4_, OPL_)                                                              \
  if (T4 == T4_ && OPL == OPL_)                                                           \
    return x3 ? launch_fused_variant<T4_, OPL_, true>(p, grid, smem, stream)              \
              : launch_fused_variant<T4_, OPL_, false>(p, grid, smem, stream);
  SRF_FUSED(2, 1)
  SRF_FUSED(2, 2)
  SRF_FUSED(4, 1)
  SRF_FUSED(5, 1)
#undef SRF_FUSED
  return cudaErrorInvalidValue;
}

}  // namespace srf
