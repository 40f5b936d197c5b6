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
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include "routing_kernels.h"
#include "sm100_ptx.cuh"

namespace srf {

// ---------------------------------------------------------------------------------------
// weight packing:  W[I,O,D,d], bias[I,O,D] -> Wf float[i][m][part][c][r][4]
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
    const int part = (int)(t % parts);
    t /= parts;
    const int m = (int)(t % NT);
    const int i = (int)(t / NT);
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

// FP16 operand images (SRF_UHAT_F16): Wf half[i][m][c][r][8], chunk c holds l = 8c .. 8c+7, values
// rounded to nearest-even fp16 (the same 11-bit significand as TF32) and clamped to the finite range
__global__ void pack_weights_fused_f16_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                              __half* __restrict__ Wf, int I, int O, int D, int d, int T4,
                                              int OPL, int KC) {
  const int NT = OPL * T4;
  const long long n = (long long)I * NT * KC * 1024;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const int li = (int)(e & 7);
    long long t = e >> 3;
    const int r = (int)(t & 127);
    t >>= 7;
    const int c = (int)(t % KC);
    t /= KC;
    const int m = (int)(t % NT);
    const int i = (int)(t / NT);
    const int jb = m / T4, k4 = m - jb * T4;
    const int j = jb * 32 + (r & 31), k = 4 * k4 + (r >> 5), l = 8 * c + li;
    float v = 0.f;
    if (j < O && k < D) {
      if (l < d)
        v = W[(((long long)i * O + j) * D + k) * d + l];
      else if (l == d)
        v = bias[((long long)i * O + j) * D + k];
    }
    Wf[e] = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
  }
}

// parts: 1 = TF32 image, 2 = hi + lo images (3 x TF32), 0 = FP16 image
void launch_pack_weights_fused(const float* W, const float* bias, float* Wf, int I, int O, int D, int d,
                               int T4, int OPL, int KC, int parts, cudaStream_t stream) {
  const long long n = (long long)I * (parts ? parts : 1) * OPL * T4 * KC * 512;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  if (parts == 0)
    pack_weights_fused_f16_kernel<<<blocks, 256, 0, stream>>>(W, bias, reinterpret_cast<__half*>(Wf), I, O, D, d,
                                                              T4, OPL, KC);
  else
    pack_weights_fused_kernel<<<blocks, 256, 0, stream>>>(W, bias, Wf, I, O, D, d, T4, OPL, KC, parts);
}

namespace {

constexpr int FZ_N = 32;                 // frames per group = MMA N
constexpr int FZ_TF = 16;                // frames per team
constexpr int FZ_MATH_WARPS = 8;
constexpr int FZ_THREADS = 384;          // 4 service warps + 8 math warps
constexpr int FZ_CAP_RING = 8;           // capsule-completion barriers (>= x ring depth and TMEM buffers)
constexpr int FZ_STG = 4;                // fp32 staging slots of the FP16-image x loader (3 capsules of lookahead)
#ifndef SRF_FZ_XST
#define SRF_FZ_XST 8
#endif
constexpr int FZ_XST_MAX = SRF_FZ_XST;            // x-tile ring depth (tf32; the 3 x TF32 build keeps two images: 4)
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
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
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
#ifdef SRF_EXP_NOLDTM
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(taddr + i) * 1e-30f;
  return;
#endif
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

// Cross-CTA synchronisation goes through L2 only: flags are polled with ld.relaxed.gpu and bumped
// with red.release.gpu / st.release.gpu, the data behind them is read with ld.global.cg / cp.async.cg.
// (ld.acquire.gpu and __threadfence() make ptxas emit CCTL.IVALL: every poll would wipe the L1 that
// holds the routing warps' spill slots and the LayerNorm parameters.)
__device__ __forceinline__ int ld_relaxed(const int* p) {
  int v;
  asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// mbarriers are addressed by their 32-bit shared-memory address (half the registers of a pointer)
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(20000u)   // may sleep up to 20 us, wakes on completion
      : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may park the thread for an implementation-defined time)
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bulk copy global -> shared with an L2 eviction policy: the packed weights are re-read every time
// step by every frame group and must stay resident in L2 while capsules and exchange buffers
// stream through it (without the hint ncu showed 38 GB of DRAM reads per cfg-3 step)
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                           uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], TF32 (or FP16 operands: kind::f16); descriptors given as (low word,
// shared high word)
template <bool F16>
__device__ __forceinline__ void mma_tf32_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                            uint32_t idesc, bool accumulate) {
  if (F16) {
    if (accumulate)
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
          "setp.ne.b32 p, 1, 0;\n\t"
          "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
          "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
          : "memory");
    else
      asm volatile(
          "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
          "setp.ne.b32 p, 0, 0;\n\t"
          "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
          "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
          : "memory");
    return;
  }
  if (accumulate)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
}
// all MMAs of one capsule: NTA tiles x NKS K steps (x 3 in the 3 x TF32 build), no run-time
// predicates between the instructions
template <int NTA, int NKS, bool X3, bool F16>
__device__ __forceinline__ void mma_issue_capsule(uint32_t d_base, const uint32_t (&a_lo)[NTA], uint32_t b_lo,
                                                  uint32_t blo_lo, uint32_t wtile16, uint32_t desc_hi,
                                                  uint32_t idesc) {
#pragma unroll
  for (int m = 0; m < NTA; ++m) {
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) {
      mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo[m] + ks * 256, b_lo + ks * (2 * FZ_N), desc_hi, idesc, ks > 0);
      if (X3) {
        mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo[m] + wtile16 + ks * 256, b_lo + ks * (2 * FZ_N), desc_hi, idesc, true);
        mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo[m] + ks * 256, blo_lo + ks * (2 * FZ_N), desc_hi, idesc, true);
      }
    }
  }
}
__device__ __forceinline__ void mma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// bounded waits.  `dead` is the CTA-local copy of the abort state: once set nothing blocks any more.
struct Waiter {
  int* abort_flag;           // global
  volatile int* host_abort;  // mapped host memory: read by the host without a synchronise
  volatile int* dead;        // shared
  // cold path, called every few thousand polls: true = stop waiting
  __device__ __forceinline__ bool give_up(long long t0, int code) {
    if (*dead) return true;
    if (ld_relaxed(abort_flag) != 0) {
      *dead = 1;
      return true;
    }
    if (clock64() - t0 > FZ_TIMEOUT) {
      *dead = 1;
      if (atomicCAS(abort_flag, 0, code) == 0 && host_abort) {
        *host_abort = code;
        __threadfence_system();
      }
      return true;
    }
    return false;
  }
  __device__ __forceinline__ void mbar(uint32_t bar, uint32_t parity, int code) {
    if (mbar_try_wait_a(bar, parity)) return;
    long long t0 = 0;
    unsigned n = 0;
    while (!mbar_try_wait_hint(bar, parity)) {
      if ((++n & 1023u) == 0) {
        if (t0 == 0) t0 = clock64();
        if (give_up(t0, code)) return;
      }
    }
  }
  // wait until *p >= target (global counter bumped by other CTAs with release semantics)
  __device__ __forceinline__ void counter(const int* p, int target, int code) {
    if (ld_relaxed(p) >= target) return;
    long long t0 = 0;
    unsigned n = 0;
    while (ld_relaxed(p) < target) {
      __nanosleep(40);
      if ((++n & 4095u) == 0) {
        if (t0 == 0) t0 = clock64();
        if (give_up(t0, code)) return;
      }
    }
  }
};

}  // namespace

// clock64 phase timers of the routing warps (thread 0 of team 0), compiled in with
// -DSRF_FUSED_TIMERS (SRF_NVCC_EXTRA) and read by tools/dev_fused_timers.py
#ifdef SRF_FUSED_TIMERS
#define FZ_TK(slot)                                    \
  if (timing) {                                        \
    const long long now_ = clock64();                  \
    tacc[slot] += (unsigned long long)(now_ - tlast);  \
    tlast = now_;                                      \
  }
#else
#define FZ_TK(slot)
#endif

// compiler-only dependency on 16 registers that a previous tcgen05.wait::ld already covered
__device__ __forceinline__ void reg_fence16(float (&r)[16]) {
  asm volatile(""
               : "+f"(r[0]), "+f"(r[1]), "+f"(r[2]), "+f"(r[3]), "+f"(r[4]), "+f"(r[5]), "+f"(r[6]),
                 "+f"(r[7]), "+f"(r[8]), "+f"(r[9]), "+f"(r[10]), "+f"(r[11]), "+f"(r[12]), "+f"(r[13]),
                 "+f"(r[14]), "+f"(r[15])::"memory");
}
// acc[f] += a[f] * b[f] over 16 frames with packed fma.rn.f32x2: half the issue slots of
// scalar FFMA (measured on cfg-3: 15.0 -> 14.3 ms, bit-identical results)
__device__ __forceinline__ void fma16(float (&acc)[FZ_TF], const float (&a)[16], const float (&b)[FZ_TF]) {
#pragma unroll
  for (int f = 0; f < FZ_TF; f += 2) {
    const float2 r = __ffma2_rn(make_float2(a[f], a[f + 1]), make_float2(b[f], b[f + 1]), make_float2(acc[f], acc[f + 1]));
    acc[f] = r.x;
    acc[f + 1] = r.y;
  }
}
// second stage of the coupling softmax of 4 frames: exp(a - max), lane-partial normaliser, REDUX
// add in Q26 fixed point (every term is <= 1; the normaliser keeps ~2^-21 relative accuracy)
template <int OA>
__device__ __forceinline__ void fz_softmax_exp(const float (&a4)[OA][4], const int (&mi)[4], float (&ex)[OA][4],
                                               unsigned (&zi)[4]) {
  constexpr float LOG2E = 1.4426950408889634f;
  constexpr float QS = (float)(1 << 26) / (float)OA;
#pragma unroll
  for (int ff = 0; ff < 4; ++ff) {
    int x = mi[ff];
    x ^= (x >> 31) & 0x7fffffff;
    const float mx = __int_as_float(x);
    float zl = 0.f;
#pragma unroll
    for (int jb = 0; jb < OA; ++jb) {
      ex[jb][ff] = fast_ex2((a4[jb][ff] - mx) * LOG2E);
      zl += ex[jb][ff];
    }
    zi[ff] = __reduce_add_sync(0xffffffffu, __float2uint_rn(zl * QS));
  }
}

// ---------------------------------------------------------------------------------------
// The routing math of a team (4 warps x 16 frames), software-pipelined over two capsules:
//   iteration n, first half  : S(n-1) softmax over the output capsules for 4 of the team's 16 frames
//                              (reads the partial logits P[(n-1) & 1], writes the coefficients C)
//                              A(n)   partial logits of capsule n over this warp's k -> P[n & 1]
//   --- named barrier ---
//   iteration n, second half : B(n-1) t += c * u_hat of capsule n-1 (reads C, re-reads u_hat from TMEM)
// S and A are independent, so the long latencies (TMEM loads, REDUX / EX2 / RCP chains: 90-150 clk
// each, profiles/r2_ubench.txt) overlap.  Only capsules n and n-1 are live in TMEM: the third buffer
// is free for the MMAs of capsule n+1 during the whole iteration (with B two capsules behind, the
// MMAs of capsule n+1 could only start once B had released its buffer, and iteration n+1 had to
// wait for them: a serial chain).  OA = blocks of 32 output capsules the layer uses.
// ---------------------------------------------------------------------------------------
template <int T4, int OPL, int OA, bool DOA, bool DOS>
__device__ __forceinline__ void fz_step_sa(const float (&va)[OPL][T4][FZ_TF], uint32_t tbA, uint32_t sP_w,
                                           uint32_t sP_r, uint32_t sC_w, int q, int lane, uint32_t swz, int O,
                                           int mask0) {
  constexpr int NTA = OA * T4;
  float a4[OA][4], ex[OA][4];
  int mi[4];
  unsigned zi[4];
  float u0[16], u1[16];
  if (DOA) {
    tmem_ld16f(tbA, u0);
    if (NTA > 1) tmem_ld16f(tbA + (uint32_t)FZ_N, u1);
  }
  if (DOS) {
    // logits of frames 4q .. 4q+3: sum of the four warps' partial dot products (naive:205/:223/:240)
#pragma unroll
    for (int jb = 0; jb < OA; ++jb) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const float4 x = lds128f(sP_r + (uint32_t)(((qq * OPL + jb) * 32 + lane) * 64) + (((uint32_t)q ^ swz) << 4));
        acc.x += x.x;
        acc.y += x.y;
        acc.z += x.z;
        acc.w += x.w;
      }
      const int j = jb * 32 + lane;
      const bool valid = (j < O) && !(mask0 && j == 0);
      a4[jb][0] = valid ? acc.x : -CUDART_INF_F;
      a4[jb][1] = valid ? acc.y : -CUDART_INF_F;
      a4[jb][2] = valid ? acc.z : -CUDART_INF_F;
      a4[jb][3] = valid ? acc.w : -CUDART_INF_F;
    }
    // max over the output capsules: one REDUX on an order-preserving integer image
#pragma unroll
    for (int ff = 0; ff < 4; ++ff) {
      float m = a4[0][ff];
#pragma unroll
      for (int jb = 1; jb < OA; ++jb) m = fmaxf(m, a4[jb][ff]);
      int x = __float_as_int(m);
      x ^= (x >> 31) & 0x7fffffff;
      mi[ff] = __reduce_max_sync(0xffffffffu, x);
    }
  }
  float pl[OA][FZ_TF];
  if (DOA) {
#pragma unroll
    for (int jb = 0; jb < OA; ++jb)
#pragma unroll
      for (int f = 0; f < FZ_TF; ++f) pl[jb][f] = 0.f;
#pragma unroll
    for (int m = 0; m < NTA; m += 2) {
      tmem_wait16(u0);
      if (m + 1 < NTA) reg_fence16(u1);   // arrived with the same wait
      {
        const int jb = m / T4, k4 = m % T4;
        fma16(pl[jb], u0, va[jb][k4]);
      }
      if (m + 2 < NTA) tmem_ld16f(tbA + (uint32_t)((m + 2) * FZ_N), u0);
      if (m + 1 < NTA) {
        const int jb = (m + 1) / T4, k4 = (m + 1) % T4;
        fma16(pl[jb], u1, va[jb][k4]);
        if (m + 3 < NTA) tmem_ld16f(tbA + (uint32_t)((m + 3) * FZ_N), u1);
      }
      if (DOS && m == 0) fz_softmax_exp<OA>(a4, mi, ex, zi);
    }
  }
  if (DOS) {
    if (!DOA) fz_softmax_exp<OA>(a4, mi, ex, zi);
    constexpr float QS = (float)(1 << 26) / (float)OA;
#pragma unroll
    for (int ff = 0; ff < 4; ++ff) {
      const float inv = fast_rcp((float)zi[ff] * (1.0f / QS));
#pragma unroll
      for (int jb2 = 0; jb2 < OA; ++jb2) ex[jb2][ff] *= inv;
    }
#pragma unroll
    for (int jb2 = 0; jb2 < OA; ++jb2)
      sts128f(sC_w + (uint32_t)((jb2 * 32 + lane) * 64) + (((uint32_t)q ^ swz) << 4), ex[jb2][0], ex[jb2][1],
              ex[jb2][2], ex[jb2][3]);
  }
  if (DOA) {
#pragma unroll
    for (int jb = 0; jb < OA; ++jb) {
      const uint32_t row = sP_w + (uint32_t)(((q * OPL + jb) * 32 + lane) * 64);
#pragma unroll
      for (int fq = 0; fq < 4; ++fq)
        sts128f(row + (((uint32_t)fq ^ swz) << 4), pl[jb][4 * fq], pl[jb][4 * fq + 1], pl[jb][4 * fq + 2],
                pl[jb][4 * fq + 3]);
    }
  }
}

// second half: t += c * u_hat of the previous capsule; two tiles in flight per tcgen05.wait::ld
template <int T4, int OPL, int OA>
__device__ __forceinline__ void fz_step_b(float (&ta)[OPL][T4][FZ_TF], uint32_t tbB, uint32_t sC_r, int lane,
                                          uint32_t swz, uint32_t t_empty_bar) {
  constexpr int NTA = OA * T4;
  float u0[16], u1[16];
  tmem_ld16f(tbB, u0);
  if (NTA > 1) tmem_ld16f(tbB + (uint32_t)FZ_N, u1);
  float cf[OA][FZ_TF];
#pragma unroll
  for (int jb = 0; jb < OA; ++jb)
#pragma unroll
    for (int fq = 0; fq < 4; ++fq) {
      const float4 x = lds128f(sC_r + (uint32_t)((jb * 32 + lane) * 64) + (((uint32_t)fq ^ swz) << 4));
      cf[jb][4 * fq] = x.x;
      cf[jb][4 * fq + 1] = x.y;
      cf[jb][4 * fq + 2] = x.z;
      cf[jb][4 * fq + 3] = x.w;
    }
#pragma unroll
  for (int m = 0; m < NTA; m += 2) {
    tmem_wait16(u0);
    if (m + 1 < NTA) reg_fence16(u1);   // arrived with the same wait
    if (m + 2 >= NTA) {
      // the last tiles of the capsule are in registers: hand its TMEM buffer back to the MMA issuers
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(t_empty_bar);
    }
    {
      const int jb = m / T4, k4 = m % T4;
      fma16(ta[jb][k4], u0, cf[jb]);
    }
    if (m + 2 < NTA) tmem_ld16f(tbB + (uint32_t)((m + 2) * FZ_N), u0);
    if (m + 1 < NTA) {
      const int jb = (m + 1) / T4, k4 = (m + 1) % T4;
      fma16(ta[jb][k4], u1, cf[jb]);
      if (m + 3 < NTA) tmem_ld16f(tbB + (uint32_t)((m + 3) * FZ_N), u1);
    }
  }
}

// T4 = ceil(D/4) tiles per block of 32 output capsules, OPL = blocks of 32 output capsules the
// build holds (a layer may use fewer: FusedLayer::opl), X3 = 3 x TF32 split
template <int T4, int OPL, int MODE>
__global__ void __launch_bounds__(FZ_THREADS, 1) route_fused_kernel(const FusedParams p) {
  constexpr bool X3 = MODE == 1;      // 3 x TF32 split
  constexpr bool F16 = MODE == 2;     // FP16 operand images (8 elements per 16-byte chunk, K = 16 per MMA)
  constexpr int NT = T4 * OPL;        // M tiles per input capsule (at most)
  constexpr int T = 4 * T4;           // padded output capsule dim
  constexpr int OP = 32 * OPL;        // padded output capsules
  constexpr int TCOLS = NT * FZ_N;    // TMEM columns per capsule buffer
  constexpr int NBUF = (512 / TCOLS) > 4 ? 4 : (512 / TCOLS);
  static_assert(NBUF >= 3, "the math pipeline keeps three capsules in TMEM");
  constexpr int XST = X3 ? FZ_XST_MAX / 2 : FZ_XST_MAX;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31;
  // the shuffle makes the warp index warp-uniform FOR THE COMPILER: everything the MMA issuers derive
  // from it (their role, the smem descriptors) then stays on ptxas' uniform datapath instead of
  // paying five R2UR per tcgen05.mma
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  // a CTA serves ONE layer for the whole launch and a fixed slice of its input capsules; all of
  // it follows from blockIdx and the kernel parameters (warp-uniform by construction)
  int my_layer = 0;
  const int cta_in_group = (int)(blockIdx.x % (unsigned)p.per_group);
#pragma unroll 1
  for (int l = 0; l < p.n_layers - 1; ++l)
    if (cta_in_group >= p.cta_end[l]) my_layer = l + 1;
  const int my_C = p.cta_end[my_layer] - (my_layer > 0 ? p.cta_end[my_layer - 1] : 0);
  const int my_c = p.sdr ? cta_in_group - (my_layer > 0 ? p.cta_end[my_layer - 1] : 0) : (my_C > 1 ? cta_in_group : 0);
  const int my_I = p.layer_I[my_layer];
  const int u_ilo = (int)((long long)my_I * my_c / my_C), u_ihi = (int)((long long)my_I * (my_c + 1) / my_C);
  const int u_ncap = u_ihi - u_ilo;                 // input capsules per pass (same for every item)
  const int u_nt = p.layer_opl[my_layer] * T4;      // M tiles per capsule
  const int KC = p.layer_KC[my_layer];              // 16-byte K chunks per operand row
  const int KX = p.layer_KX[my_layer];              // of which the x loader fills KX
  const uint32_t wtile = (uint32_t)KC * 2048u;           // one image of one W tile
  const uint32_t wpair = X3 ? 2u * wtile : wtile;        // hi (+ lo) of one tile
  const uint32_t xtile = (uint32_t)KC * FZ_N * 16u;      // one image of one x tile
  const int NWST = p.nwst;                               // ring stages
  const int G = p.gtiles;                                // tiles per stage (one bulk copy)
  const uint32_t wstage = (uint32_t)p.wstage_bytes;      // stage stride (sized for the widest layer)

  uint8_t* sW = smem_raw;                                     // [NWST][wstage]
  uint8_t* sX = sW + (size_t)NWST * wstage;                   // [XST][xtile]  (x, or x_hi)
  uint8_t* sXlo = sX + (size_t)XST * p.xtile_bytes;        // [XST][xtile]  (X3)
  // FP16 images: fp32 staging ring of the x loader [FZ_STG][4-float chunk][frame][16 B]
  uint8_t* sStg = sXlo + (X3 ? (size_t)XST * p.xtile_bytes : 0);
  float* sP = reinterpret_cast<float*>(sStg + (F16 ? (size_t)FZ_STG * p.xstg_bytes : 0));  // [2][2 teams][4][OPL][32][16]
  float* sC = sP + 2 * 2 * 4 * OPL * 32 * 16;                 // [2][2 teams][OPL][32][16]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sC + 2 * 2 * OPL * 32 * 16);
  const uint32_t w_full = ptx::smem_u32(bars);               // [NWST]   (8 bytes each)
  const uint32_t w_empty = w_full + 8u * (uint32_t)NWST;     // [NWST]
  const uint32_t x_full = w_empty + 8u * (uint32_t)NWST;     // [XST]
  // cap_done[n % 8]: the MMAs of capsule n have retired -- ONE tcgen05.commit per capsule tells the
  // routing warps that its TMEM buffer is full and the x loader that its x slot is free (a commit
  // costs the issuing thread ~350 clk)
  const uint32_t cap_done = x_full + 8u * XST;            // [FZ_CAP_RING]
  const uint32_t t_empty = cap_done + 8u * FZ_CAP_RING;   // [NBUF]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * NWST + XST + FZ_CAP_RING + NBUF);
  volatile int* s_dead = reinterpret_cast<volatile int*>(tmem_ptr + 1);

  Waiter wt{p.abort_flag, p.host_abort, s_dead};

  if (tid == 0) {
    *s_dead = 0;
    for (int s = 0; s < NWST; ++s) {
      ptx::mbar_init(bars + s, 1);
      ptx::mbar_init(bars + NWST + s, 2);   // w_empty: one commit per MMA issuer
    }
    for (int s = 0; s < XST; ++s) ptx::mbar_init(bars + 2 * NWST + s, 32);   // x_full: every lane of the loader warp
    for (int s = 0; s < FZ_CAP_RING; ++s) ptx::mbar_init(bars + 2 * NWST + XST + s, 1);
    for (int s = 0; s < NBUF; ++s) ptx::mbar_init(bars + 2 * NWST + XST + FZ_CAP_RING + s, FZ_MATH_WARPS);
    ptx::fence_barrier_init();
  }
  // constant part of the x tiles: chunk KX carries the 1 that multiplies the bias column of W
  // (x_hi image), everything above the TMA box is zero; the TMA never writes these chunks
  // (FP16 images: the loader writes the chunks up to and including the one that holds the 1; the
  // chunks above it are zero)
  const int KXW = F16 ? (4 * KX + 1 + 7) / 8 : KX;   // chunks the x loader (re)writes per tile
  for (int e = tid; e < XST * (KC - KXW) * FZ_N; e += FZ_THREADS) {
    const int st = e / ((KC - KXW) * FZ_N), rem = e - st * ((KC - KXW) * FZ_N);
    const int c = KXW + rem / FZ_N, f = rem % FZ_N;
    float4* dst = reinterpret_cast<float4*>(sX + (size_t)st * xtile + ((size_t)c * FZ_N + f) * 16);
    *dst = make_float4((!F16 && c == KX) ? 1.f : 0.f, 0.f, 0.f, 0.f);
    if (X3) *reinterpret_cast<float4*>(sXlo + (size_t)st * xtile + ((size_t)c * FZ_N + f) * 16) =
        make_float4(0.f, 0.f, 0.f, 0.f);
  }
  ptx::fence_proxy_async();
  // warps 0-7: routing math; warps 8-11: W producer, MMA issuer, two x loaders.  The service
  // warps sit at the HIGH warp ids: the warp arbiter prefers high ids (B300_MICROARCH.md), and
  // the single MMA-issuing thread must never wait behind eight FMA-bound warps.
  const int swarp = warp - 8;   // service role
  if (swarp == 1) {
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
  if (swarp >= 0) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
  if (swarp < 2) {
    // =========================== MMA issuers (warps 8 and 9, alternate capsules) ===========================
    // One thread issues the 3 * NT MMAs of a capsule.  With ptxas 12.9 every tcgen05.mma issued
    // from divergent code costs ~120-250 clk of R2UR / ELECT bookkeeping (tools/ubench mma2), far
    // more than the 47 clk the tensor core needs for M=128 x N=32 x K=8 -- so two threads take
    // alternate capsules.  Both walk the whole tile stream: a W stage is handed back to the
    // producer by a tcgen05.commit of EACH issuer (w_empty counts 2), which arrives once that
    // thread's own MMAs on the stage have retired (at once if it had none).
    // ONE elected thread runs the issuer: inside an elect.sync-guarded region ptxas knows a single thread
    // is active and emits plain UTCHMMA instructions; under `lane == 0` it wraps every tcgen05.mma in an
    // ELECT / BRA.U.ANY serialisation loop (~100 clk of dependent uniform-datapath latency per MMA)
    if (p.capstage) {
      // ---- capsule-sized ring stages (the layout the host picks when three of them fit) ----
      // warp 8: the ONE MMA issuer.  Per capsule: x tile full, TMEM buffer free, W stage full, NT x K/16
      // MMAs from straight-line code, ONE tcgen05.commit (cap_done) that tells the routing warps "TMEM
      // buffer full", the x loader "x slot free" and the W producer "ring stage free".  No second issuer
      // whose books must be kept, no commit per ring stage.
      // warp 9: the W producer.  Stage n % NWST is refilled with capsule n's tiles (one bulk copy from L2)
      // as soon as capsule n - NWST has retired; weights do not depend on the recurrence, so the ring
      // runs ahead across time steps.
      if (swarp == 0) {
        if (ptx::elect_one()) {
          const uint32_t idesc = F16 ? ptx::make_idesc_f16(128, FZ_N) : ptx::make_idesc_tf32(128, FZ_N);
          constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
          const uint32_t a_lo0 = ((ptx::smem_u32(sW) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
          const uint32_t b_lo0 = ((ptx::smem_u32(sX) >> 4) & 0x3FFFu) | (((FZ_N * 16u) >> 4) << 16);
          const uint32_t blo_lo0 = ((ptx::smem_u32(sXlo) >> 4) & 0x3FFFu) | (((FZ_N * 16u) >> 4) << 16);
          const uint32_t wstage16 = wstage >> 4, wpair16 = wpair >> 4, wtile16 = wtile >> 4, xtile16 = xtile >> 4;
          const int nks = KC / 2;
          const int nt = u_nt;
          uint32_t n = 0, ws = 0, w_par = 0, buf = 0, t_par = 0;
#ifdef SRF_FUSED_TIMERS
          unsigned* trace = nullptr;
          if (p.dbg && (blockIdx.x == 0 || blockIdx.x == 80))
            trace = reinterpret_cast<unsigned*>(p.dbg) + 6000 + (blockIdx.x == 0 ? 0 : 3072);
          unsigned ncap_tr = 0;
#define FZ_MK1(slot) if (trace && ncap_tr < 512) trace[ncap_tr * 6 + (slot)] = (unsigned)clock();
#else
#define FZ_MK1(slot)
#endif
          for (int it = 0; it < p.rounds; ++it) {
            const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
            if (item.layer < 0) continue;
            const int total = nsteps * iters * u_ncap;
            for (int cc = 0; cc < total; ++cc, ++n) {
              // ring positions are carried, not divided out (NWST is a run-time value); the operands are
              // checked first, the TMEM buffer last: it is what the issuer normally waits for, and the MMAs
              // must go out as soon as the routing warps hand it back
              const uint32_t xs = n % (uint32_t)XST;
              FZ_MK1(0)
              wt.mbar(x_full + 8u * xs, (n / XST) & 1, 104);
              ptx::fence_proxy_async();   // cp.async / st.shared wrote the tile through the generic proxy
              FZ_MK1(1)
              wt.mbar(w_full + 8u * ws, w_par, 106);
              FZ_MK1(2)
              wt.mbar(t_empty + 8u * buf, t_par ^ 1u, 105);
              FZ_MK1(5)
              ptx::tc_fence_after();
              const uint32_t d_base = tmem_base + buf * (uint32_t)TCOLS;
              const uint32_t b_lo = b_lo0 + xs * xtile16, blo_lo = blo_lo0 + xs * xtile16;
              const uint32_t a_base = a_lo0 + ws * wstage16;
              if (nt == NT && (nks == 2 || nks == 3 || nks == 5 || (F16 && nks == 1))) {
                uint32_t al[NT];
#pragma unroll
                for (int m = 0; m < NT; ++m) al[m] = a_base + (uint32_t)m * wpair16;
                if (nks == 3) mma_issue_capsule<NT, 3, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else if (nks == 2) mma_issue_capsule<NT, 2, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else if (F16 && nks == 1) mma_issue_capsule<NT, 1, false, true>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else mma_issue_capsule<NT, 5, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
              } else {
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                  if (m < nt) {
                    const uint32_t a_lo = a_base + (uint32_t)m * wpair16;
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks) {
                      if (ks < nks) {
                        mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + ks * 256, b_lo + ks * (2 * FZ_N), DESC_HI, idesc, ks > 0);
                        if (X3) {
                          mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + wtile16 + ks * 256, b_lo + ks * (2 * FZ_N), DESC_HI, idesc, true);
                          mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + ks * 256, blo_lo + ks * (2 * FZ_N), DESC_HI, idesc, true);
                        }
                      }
                    }
                  }
                }
              }
              FZ_MK1(3)
              mma_commit_a(cap_done + 8u * (n % (uint32_t)FZ_CAP_RING));
              FZ_MK1(4)
              if (++ws == (uint32_t)NWST) {
                ws = 0;
                w_par ^= 1u;
              }
              if (++buf == (uint32_t)NBUF) {
                buf = 0;
                t_par ^= 1u;
              }
#ifdef SRF_FUSED_TIMERS
              ++ncap_tr;
#endif
            }
          }
        }
      } else if (ptx::elect_one()) {
        const uint32_t sW_a = ptx::smem_u32(sW);
        const uint64_t w_policy = l2_policy_evict_last();
        uint32_t n = 0, ws = 0;
        for (int it = 0; it < p.rounds; ++it) {
          const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
          if (item.layer < 0) continue;
          const uint32_t bytes = (uint32_t)u_nt * wpair;
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.layers[item.layer].Wf) + (size_t)item.i_lo * bytes;
          const int npass = nsteps * iters;
          for (int ps = 0; ps < npass; ++ps)
            for (int cc = 0; cc < u_ncap; ++cc, ++n) {
              if (n >= (uint32_t)NWST) {
                const uint32_t prev = n - (uint32_t)NWST;   // the stage's previous tenant has retired
                wt.mbar(cap_done + 8u * (prev % FZ_CAP_RING), (prev / FZ_CAP_RING) & 1, 103);
              }
              mbar_expect_tx_a(w_full + 8u * ws, bytes);
              bulk_g2s_a(sW_a + ws * wstage, wsrc + (size_t)cc * bytes, bytes, w_full + 8u * ws, w_policy);
              if (++ws == (uint32_t)NWST) ws = 0;
            }
        }
      }
    } else if (ptx::elect_one()) {
      const int me = swarp;
      const uint32_t idesc = F16 ? ptx::make_idesc_f16(128, FZ_N) : ptx::make_idesc_tf32(128, FZ_N);
      // K-major no-swizzle descriptors: low word = start address >> 4 | LBO >> 4 << 16, high word =
      // SBO >> 4 | version 1; only the 14-bit address field changes from MMA to MMA
      constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
      const uint32_t a_lo0 = ((ptx::smem_u32(sW) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
      const uint32_t b_lo0 = ((ptx::smem_u32(sX) >> 4) & 0x3FFFu) | (((FZ_N * 16u) >> 4) << 16);
      const uint32_t blo_lo0 = ((ptx::smem_u32(sXlo) >> 4) & 0x3FFFu) | (((FZ_N * 16u) >> 4) << 16);
      const uint32_t wstage16 = wstage >> 4, wpair16 = wpair >> 4, wtile16 = wtile >> 4, xtile16 = xtile >> 4;
      const int nks = KC / 2;
      uint32_t n_x = 0, n_t = 0;
#ifdef SRF_FUSED_TIMERS
      // event trace of the first 512 capsules of CTA 0 and CTA 80: 32-bit clock per event
      unsigned* trace = nullptr;
      if (p.dbg && me == 0 && (blockIdx.x == 0 || blockIdx.x == 80))
        trace = reinterpret_cast<unsigned*>(p.dbg) + 6000 + (blockIdx.x == 0 ? 0 : 3072);
      unsigned ncap_tr = 0;
#define FZ_MK(slot) if (trace && ncap_tr < 512) trace[ncap_tr * 6 + (slot)] = (unsigned)clock();
#else
#define FZ_MK(slot)
#endif
      // The MMAs of a capsule are issued from straight-line code: NT tiles x up to 5 K steps with
      // compile-time offsets from a handful of per-capsule bases, so ptxas keeps the descriptors in
      // uniform registers (52 clk per MMA instead of 140+ with per-MMA R2UR chains, tools/ubench
      // mma3).  A capsule's tiles may straddle up to three ring stages: base[s] is chosen per tile.
      uint32_t gs = 0;        // ring stages consumed so far (global count; slot = gs % NWST)
      uint32_t w_seen = 0;    // stages this thread has already seen full
      for (int it = 0; it < p.rounds; ++it) {
        const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
        if (item.layer < 0) continue;
        const int ncap = u_ncap, nt = u_nt;
        const int npass = nsteps * iters;
        for (int ps = 0; ps < npass; ++ps) {
          int tin = 0;          // position of the capsule's first tile inside its ring stage
          for (int cc = 0; cc < ncap; ++cc, ++n_t, ++n_x) {
            const bool mine = (int)(n_t & 1u) == me;
            const int tend = tin + nt;                       // one past the capsule's last tile, stage-relative
            const int nst = tend <= G ? 1 : (tend <= 2 * G ? 2 : 3);
            if (mine) {
              const int xs = n_x % XST;
              const int buf = n_t % NBUF;
              FZ_MK(0)
              wt.mbar(x_full + 8u * xs, (n_x / XST) & 1, 104);
              ptx::fence_proxy_async();   // cp.async wrote the tile through the generic proxy
              FZ_MK(1)
              wt.mbar(t_empty + 8u * buf, ((n_t / NBUF) & 1) ^ 1, 105);
              FZ_MK(2)
              for (int sidx = 0; sidx < nst; ++sidx) {
                const uint32_t g = gs + (uint32_t)sidx;
                if (g >= w_seen) {
                  wt.mbar(w_full + 8u * (g % NWST), (g / NWST) & 1, 106);
                  w_seen = g + 1;
                }
              }
              FZ_MK(5)
              ptx::tc_fence_after();
              const uint32_t d_base = tmem_base + (uint32_t)(buf * TCOLS);
              const uint32_t b_lo = b_lo0 + (uint32_t)xs * xtile16, blo_lo = blo_lo0 + (uint32_t)xs * xtile16;
              // base[s] + m * wpair16 is the descriptor low word of tile m when it lies in stage gs + s
              const uint32_t base0 = a_lo0 + (gs % NWST) * wstage16 + (uint32_t)tin * wpair16;
              const uint32_t base1 = a_lo0 + ((gs + 1) % NWST) * wstage16 + (uint32_t)(tin - G) * wpair16;
              const uint32_t base2 = a_lo0 + ((gs + 2) % NWST) * wstage16 + (uint32_t)(tin - 2 * G) * wpair16;
#ifndef SRF_FUSED_NOMMA
              if (nt == NT && (nks == 2 || nks == 3 || nks == 5 || (F16 && nks == 1))) {
                uint32_t al[NT];
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                  const int t = tin + m;
                  al[m] = (t < G ? base0 : (t < 2 * G ? base1 : base2)) + (uint32_t)m * wpair16;
                }
                if (nks == 3) mma_issue_capsule<NT, 3, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else if (nks == 2) mma_issue_capsule<NT, 2, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else if (F16 && nks == 1) mma_issue_capsule<NT, 1, false, true>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
                else mma_issue_capsule<NT, 5, X3, F16>(d_base, al, b_lo, blo_lo, wtile16, DESC_HI, idesc);
              } else {
#pragma unroll
                for (int m = 0; m < NT; ++m) {
                  if (m < nt) {
                    const int t = tin + m;
                    const uint32_t a_lo = (t < G ? base0 : (t < 2 * G ? base1 : base2)) + (uint32_t)m * wpair16;
#pragma unroll
                    for (int ks = 0; ks < 5; ++ks) {
                      if (ks < nks) {
                        mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + ks * 256, b_lo + ks * (2 * FZ_N), DESC_HI, idesc, ks > 0);
                        if (X3) {
                          mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + wtile16 + ks * 256, b_lo + ks * (2 * FZ_N), DESC_HI, idesc, true);
                          mma_tf32_lo<F16>(d_base + m * FZ_N, a_lo + ks * 256, blo_lo + ks * (2 * FZ_N), DESC_HI, idesc, true);
                        }
                      }
                    }
                  }
                }
              }
#endif
              FZ_MK(3)
            }
            // stages whose last tile belongs to this capsule (or that end with the pass) go back to
            // the producer: the commit arrives once MY MMAs on them have retired
            const bool pass_end = cc == ncap - 1;
            {
              int left = tend;
              while (left >= G || (pass_end && left > 0)) {
                // never hand a stage back before having seen it full: an issuer that only keeps
                // the books for a capsule could otherwise arrive a whole ring revolution early and
                // complete the barrier phase the other issuer's MMAs still depend on
                if (gs >= w_seen) {
                  wt.mbar(w_full + 8u * (gs % NWST), (gs / NWST) & 1, 107);
                  w_seen = gs + 1;
                }
                mma_commit_a(w_empty + 8u * (gs % NWST));
                ++gs;
                left -= G;
              }
              tin = left > 0 ? left : 0;
              if (pass_end) tin = 0;
            }
            if (mine) {
              mma_commit_a(cap_done + 8u * (uint32_t)(n_t % FZ_CAP_RING));
              FZ_MK(4)
#ifdef SRF_FUSED_TIMERS
              ++ncap_tr;
#endif
            }
          }
        }
      }
    }
  } else if (swarp == 2) {
    // =========================== x loader ===========================
    // lane = frame of the group.  The frame's input capsule (d floats, contiguous) goes to the
    // K-major operand image [16-byte chunk][frame][16 B]; frames outside the utterance / batch and
    // window positions outside [0, S) are zeros (naive:150).  (A 5-D tensor-map load of the same
    // box costs ~15 clk per 16-byte row on the TMA unit: 2400 clk per tile, measured.)
    // tf32: cp.async (16 B, zero fill when the source is out of range) straight into shared
    // memory, completion counted on the tile's mbarrier -- no registers, tiles of several
    // capsules in flight.  X3: through registers, where x is split into x_hi + x_lo.  For layers
    // fed by another layer of the same launch every lane first waits for ITS utterance's progress.
    // Lane 0 also feeds the W ring: one bulk copy per stage of G tiles, always after the x tiles of
    // every capsule that has a tile in the stage (so the issuers can never starve on x).
    uint32_t n_x = 0, n_w = 0;
    const uint32_t sW_a = ptx::smem_u32(sW);
    const uint64_t w_policy = l2_policy_evict_last();
    for (int it = 0; it < p.rounds; ++it) {
      const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
      if (item.layer < 0) continue;
      const FusedLayer* L = p.layers + item.layer;
      const int H = L->H, lpad = L->lpad, d = 4 * KX;
      const int nt = L->opl * T4;
      const int ntiles = (item.i_hi - item.i_lo) * nt;   // one pass: a contiguous range of the packed weights
      const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(L->Wf) + (size_t)item.i_lo * nt * wpair;
      const int fb = lane % p.NB, fs = lane / p.NB;
      const int b = item.b0 + fb;
      const bool frame_ok = ((item.vmask >> lane) & 1u) != 0;
      const int* prog = L->dep_layer >= 0
                            ? p.progress + ((size_t)L->dep_layer * p.ngroups + item.group) * FZ_N + lane
                            : nullptr;
      const int rpad = L->rpad;
      const float* base = L->emb + (size_t)(frame_ok ? b : 0) * p.S * H * d;
      for (int step = 0; step < nsteps; ++step) {
        const int s0 = (p.sdr ? step : item.s0) + fs;
        if (prog && frame_ok) {
          // the producing layer must have stored this utterance's frames of the whole window
          int need = step + rpad + 1;
          if (need > p.S) need = p.S;
          wt.counter(prog, need, 101);
        }
        // FP16 images: the frame's fp32 capsule is staged by cp.async FZ_STG - 1 capsules ahead of its
        // conversion (the x of a step is complete once the progress wait above has passed)
        const int ncapP = item.i_hi - item.i_lo, qtotal = iters * ncapP;
        const uint32_t stg_a = ptx::smem_u32(sStg) + (uint32_t)lane * 16u;
        auto stage_x = [&](int q) {
          if (q < qtotal) {
            const int iq = item.i_lo + q % ncapP;
            const int wq = iq / H, hq = iq - wq * H;
            const int sq = s0 - lpad + wq;
            const bool okq = frame_ok && sq >= 0 && sq < p.S;
            const float* srcq = base + ((size_t)(okq ? sq : 0) * H + hq) * d;
            const uint32_t dq = stg_a + (uint32_t)(q % FZ_STG) * (uint32_t)p.xstg_bytes;
            const uint32_t nb = okq ? 16u : 0u;
            for (int c = 0; c < KX; ++c)
              asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dq + (uint32_t)c * (FZ_N * 16u)),
                           "l"(srcq + 4 * c), "r"(nb)
                           : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        };
        int q = 0;
        if (F16)
          for (int qq = 0; qq < FZ_STG - 1; ++qq) stage_x(qq);
        for (int pass = 0; pass < iters; ++pass) {
          int i = item.i_lo;
          for (int t0 = 0; t0 < ntiles; t0 += G) {
          const int cnt = ntiles - t0 < G ? ntiles - t0 : G;
          const int i_end = item.i_lo + (t0 + cnt - 1) / nt + 1;   // capsules with a tile in this stage
          for (; i < i_end; ++i, ++n_x, ++q) {
            const int w = i / H, h = i - w * H;
            const int s = s0 - lpad + w;
            const bool ok = frame_ok && s >= 0 && s < p.S;
            const float* src = base + ((size_t)(ok ? s : 0) * H + h) * d;
            const int xs = n_x % XST;
            // the slot's previous tenant was capsule n_x - XST
            if (n_x >= (uint32_t)XST) {
              const uint32_t prev = n_x - (uint32_t)XST;
              wt.mbar(cap_done + 8u * (prev % FZ_CAP_RING), (prev / FZ_CAP_RING) & 1, 102);
            }
            const uint32_t dst = ptx::smem_u32(sX) + (uint32_t)xs * xtile + (uint32_t)lane * 16u;
            if (F16) {
              // staged fp32 -> fp16 (rn; clamped to the finite fp16 range): chunk c holds the elements
              // 8c .. 8c+7, element d is the 1 that multiplies the bias column of W
              stage_x(q + FZ_STG - 1);
              asm volatile("cp.async.wait_group %0;" ::"n"(FZ_STG - 1) : "memory");
              const uint32_t sq_a = stg_a + (uint32_t)(q % FZ_STG) * (uint32_t)p.xstg_bytes;
              for (int c0 = 0; c0 < KXW; c0 += 2) {
                float4 v[4];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                  const int qi = 2 * c0 + q4;
                  v[q4] = make_float4(qi == KX ? 1.f : 0.f, 0.f, 0.f, 0.f);
                  if (qi < KX) v[q4] = lds128f(sq_a + (uint32_t)qi * (FZ_N * 16u));
                }
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                  if (c0 + cc < KXW) {
                    uint32_t h[4];
                    const float4 a = v[2 * cc], b2 = v[2 * cc + 1];
                    const float e[8] = {a.x, a.y, a.z, a.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                      const float lo = fminf(fmaxf(e[2 * q4], -65504.f), 65504.f);
                      const float hi = fminf(fmaxf(e[2 * q4 + 1], -65504.f), 65504.f);
                      asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[q4]) : "f"(hi), "f"(lo));
                    }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + (uint32_t)(c0 + cc) * (FZ_N * 16u)),
                                 "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3])
                                 : "memory");
                  }
                }
              }
              ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's reads
              mbar_arrive_a(x_full + 8u * xs);
            } else if (!X3) {
              const uint32_t nbytes = ok ? 16u : 0u;
              for (int c = 0; c < KX; ++c)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst + (uint32_t)c * (FZ_N * 16u)),
                             "l"(src + 4 * c), "r"(nbytes)
                             : "memory");
              asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(x_full + 8u * xs) : "memory");
            } else {
              const uint32_t dlo = ptx::smem_u32(sXlo) + (uint32_t)xs * xtile + (uint32_t)lane * 16u;
              for (int c0 = 0; c0 < KX; c0 += 4) {
                float4 v[4];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  v[cc] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (c0 + cc < KX && ok) v[cc] = __ldcg(reinterpret_cast<const float4*>(src) + c0 + cc);
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                  const int c = c0 + cc;
                  if (c < KX) {
                    float hx, hy, hz, hw;
                    uint32_t t;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[cc].x));
                    hx = __uint_as_float(t);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[cc].y));
                    hy = __uint_as_float(t);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[cc].z));
                    hz = __uint_as_float(t);
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[cc].w));
                    hw = __uint_as_float(t);
                    sts128f(dst + (uint32_t)c * (FZ_N * 16u), hx, hy, hz, hw);
                    sts128f(dlo + (uint32_t)c * (FZ_N * 16u), v[cc].x - hx, v[cc].y - hy, v[cc].z - hz, v[cc].w - hw);
                  }
                }
              }
              ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's reads
              mbar_arrive_a(x_full + 8u * xs);
            }
          }
          if (lane == 0 && !p.capstage) {
            const int ws = n_w % NWST;
            wt.mbar(w_empty + 8u * ws, ((n_w / NWST) & 1) ^ 1, 103);
            const uint32_t bytes = (uint32_t)cnt * wpair;
#ifdef SRF_EXP_NOW
            mbar_arrive_a(w_full + 8u * ws);
#else
            mbar_expect_tx_a(w_full + 8u * ws, bytes);
            bulk_g2s_a(sW_a + (uint32_t)ws * wstage, wsrc + (size_t)t0 * wpair, bytes, w_full + 8u * ws, w_policy);
#endif
          }
          ++n_w;
          __syncwarp();
          }
        }
      }
    }
  } else {
    // =========================== output warp ===========================
    // After the last pass of a step the frames this CTA owns leave the routing stack here:
    // LayerNorm + dropout (naive:188-191), the head (naive:193), the stores, and the progress flag
    // the next layer's x loader waits on -- off the critical path of the recurrence (the routing
    // warps only publish the squashed v).  lane = output capsule.
    int epoch = 0;
    int v_target[2] = {0, 0};
    for (int it = 0; it < p.rounds; ++it) {
      const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
      if (item.layer < 0) continue;
      const FusedLayer* L = p.layers + item.layer;
      const int O = L->O, D = L->D;
      const float* const ln_gamma = L->ln_gamma;
      const float* const ln_beta = L->ln_beta;
      const float* const dropout_mask = L->dropout_mask;
      const float* const head_gamma = L->head_gamma;
      const float* const head_beta = L->head_beta;
      float* const out_caps = L->out_caps;
      float* const out_logits = L->out_logits;
      float* const out_raw = L->out_raw;
      const float ln_eps = L->ln_eps, length_eps = L->length_eps;
      const bool vec4 = (D & 3) == 0 && (((uintptr_t)out_caps | (uintptr_t)out_raw) & 15) == 0;
      const float* const Vbuf = p.xV + (size_t)item.slot * FZ_N * T * OP;
      const int* const cnt_v = p.cnt_v + item.slot * 2;
      int* const oflag = p.oflag + item.slot * FZ_N;
      int* const prog = (p.progress && p.sdr)
                            ? p.progress + ((size_t)item.layer * p.ngroups + item.group) * FZ_N
                            : nullptr;
      const int tf0 = __popc(item.vmask & 0xffffu), tf1 = __popc(item.vmask >> 16);
      for (int step = 0; step < nsteps; ++step) {
        for (int pass = 0; pass < iters; ++pass) {
          ++epoch;
          v_target[0] += tf0;
          v_target[1] += tf1;
          if (pass != iters - 1) continue;
          for (int f = item.c; f < FZ_N; f += item.C) {
            if (!((item.vmask >> f) & 1u)) {
              // no such frame in this group: keep the row's epoch moving (a later group of this
              // slot may have the frame, and its owner waits on the flag)
              if (lane == 0) st_release(oflag + f, epoch);
              continue;
            }
            if (lane == 0) wt.counter(cnt_v + (f >> 4), v_target[f >> 4], 113);
            __syncwarp();
            float y[OPL][T];
            {
              const float* src = Vbuf + (size_t)f * T * OP + lane;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) y[jb][k] = __ldcg(src + k * OP + jb * 32);
            }
            // v is in registers: the owner may overwrite this row in the next pass
            __syncwarp();
            if (lane == 0) st_release(oflag + f, epoch);
            const int fb = f % p.NB, fs = f / p.NB;
            const int b = item.b0 + fb, s = (p.sdr ? step : item.s0) + fs;
            const long long frame = (long long)b * p.S + s;
            if (out_raw) {
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb) {
                const int j = jb * 32 + lane;
                if (j >= O) continue;
                float* dst = out_raw + (frame * O + j) * D;
                if (vec4) {
#pragma unroll
                  for (int k = 0; k < T; k += 4)
                    if (k < D) *reinterpret_cast<float4*>(dst + k) = make_float4(y[jb][k], y[jb][k + 1], y[jb][k + 2], y[jb][k + 3]);
                } else {
#pragma unroll
                  for (int k = 0; k < T; ++k)
                    if (k < D) dst[k] = y[jb][k];
                }
              }
            }
            if (ln_gamma) {
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
              const float rstd = X3 ? 1.0f / sqrtf(var + ln_eps) : rsqrtf(var + ln_eps);
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) {
                  const int j = jb * 32 + lane;
                  const bool okp = j < O && k < D;
                  const float g = okp ? __ldg(ln_gamma + j * D + k) : 0.f;
                  const float be = okp ? __ldg(ln_beta + j * D + k) : 0.f;
                  y[jb][k] = (y[jb][k] - mean) * rstd * g + be;
                }
            }
            float len[OPL];
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb) {
              const int j = jb * 32 + lane;
              float l2 = 0.f;
              if (j < O) {
                if (dropout_mask) {
#pragma unroll
                  for (int k = 0; k < T; ++k)
                    if (k < D) y[jb][k] *= __ldg(dropout_mask + (frame * O + j) * D + k);
                }
#pragma unroll
                for (int k = 0; k < T; ++k)
                  if (k < D) l2 = fmaf(y[jb][k], y[jb][k], l2);
                if (out_caps) {
                  float* dst = out_caps + (frame * O + j) * D;
                  if (vec4) {
#pragma unroll
                    for (int k = 0; k < T; k += 4)
                      if (k < D) *reinterpret_cast<float4*>(dst + k) = make_float4(y[jb][k], y[jb][k + 1], y[jb][k + 2], y[jb][k + 3]);
                  } else {
#pragma unroll
                    for (int k = 0; k < T; ++k)
                      if (k < D) dst[k] = y[jb][k];
                  }
                }
              }
              len[jb] = sqrtf(l2 + length_eps);  // naive:256-258
            }
            if (head_gamma) {  // ln_output over the capsule lengths (naive:193)
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
              const float hr = 1.0f / sqrtf(sq / (float)O + ln_eps);
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb) {
                const int j = jb * 32 + lane;
                if (j < O)
                  out_logits[frame * O + j] = (len[jb] - hm) * hr * __ldg(head_gamma + j) + __ldg(head_beta + j);
              }
            }
            if (prog) {
              // publish: frame f of this unit is stored up to and including `step`
              __syncwarp();
              if (lane == 0) st_release(prog + f, step + 1);
            }
          }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // =========================== routing math ===========================
    const int mw = warp;              // 0..7
    const int team = mw >> 2;         // frames team*16 .. +15
    const int q = mw & 3;             // TMEM lane quarter == warp % 4; owns k = q mod 4
    const int mtid = tid;             // 0..255
    const int bar_team = BAR_TEAM0 + team;
    // swizzled exchange rows: 16 floats (64 B) per (jb, lane); the 16-byte chunk fq is stored at
    // position fq ^ ((lane >> 1) & 3) so that 8 lanes of a store / load phase hit 32 distinct banks
    const uint32_t swz = (uint32_t)((lane >> 1) & 3);
    constexpr uint32_t P_BYTES = 4 * OPL * 32 * 64, C_BYTES = OPL * 32 * 64;   // per team (and parity for P)
    const uint32_t sP_team = ptx::smem_u32(sP) + (uint32_t)team * (2 * P_BYTES);
    const uint32_t sC_team = ptx::smem_u32(sC) + (uint32_t)team * C_BYTES;
    const uint32_t tmem_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)team * FZ_TF;

#ifdef SRF_FUSED_TIMERS
    const bool timing = p.dbg != nullptr && mtid == 0;
    unsigned long long tacc[16];
#pragma unroll
    for (int z = 0; z < 16; ++z) tacc[z] = 0;
    long long tlast = clock64();
#endif
    uint32_t n_t = 0;       // capsules consumed (all passes)
    int epoch = 0;          // passes completed by this CTA (all CTAs of a slot agree)
    int v_target = 0;       // running target of this team's frame counter

    for (int it = 0; it < p.rounds; ++it) {
      const FusedItem item = p.items[(size_t)it * gridDim.x + blockIdx.x];
      if (item.layer < 0) continue;
      const FusedLayer* L = p.layers + item.layer;
      const int O = L->O;
      const int opl = L->opl;
      const int C = item.C, c = item.c;
      const int ncap = item.i_hi - item.i_lo;
      const uint32_t team_mask = (item.vmask >> (team * FZ_TF)) & 0xffffu;
      const int team_frames = __popc(team_mask);
      float* const Pbuf = p.xP + (size_t)item.slot * p.maxC * FZ_N * T * OP;   // [C][32][T][OP]
      float* const Vbuf = p.xV + (size_t)item.slot * FZ_N * T * OP;            // [32][T][OP]
      int* const cnt_p = p.cnt_p + item.slot;
      int* const cnt_v = p.cnt_v + item.slot * 2;
      const int* const oflag = p.oflag + item.slot * FZ_N;
      const int mask0 = L->mask0;

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

          FZ_TK(0)
          // a team none of whose 16 frames exists (batches of <= 16 utterances) only keeps the TMEM ring
          // moving: its issue slots and TMEM bandwidth go to the other team
          if (team_frames == 0) {
            for (int n = 0; n < ncap; ++n) {
              const uint32_t capA = n_t + (uint32_t)n;
              wt.mbar(cap_done + 8u * (capA % FZ_CAP_RING), (capA / FZ_CAP_RING) & 1, 110);
              __syncwarp();
              if (lane == 0) mbar_arrive_a(t_empty + 8u * (uint32_t)(capA % NBUF));
            }
          } else
          // ---- the capsule loop, software-pipelined over two capsules ----
          for (int n = 0; n <= ncap; ++n) {
            const bool doA = n < ncap, doSB = n >= 1;
            const uint32_t capA = n_t + (uint32_t)n, capB = capA - 1u;
            const int bufA = capA % NBUF, bufB = capB % NBUF;
            if (doA) {
              wt.mbar(cap_done + 8u * (capA % FZ_CAP_RING), (capA / FZ_CAP_RING) & 1, 110);
              ptx::tc_fence_after();
            }
            FZ_TK(1)
            named_sync(bar_team, 128);   // P(n-1) is complete, C may be rewritten
            FZ_TK(2)
            const uint32_t tbA = tmem_lane + (uint32_t)(bufA * TCOLS), tbB = tmem_lane + (uint32_t)(bufB * TCOLS);
            const uint32_t pw = sP_team + (uint32_t)(n & 1) * P_BYTES, pr = sP_team + (uint32_t)((n + 1) & 1) * P_BYTES;
#define FZ_SA(OA_, A_, S_) fz_step_sa<T4, OPL, OA_, A_, S_>(va, tbA, pw, pr, sC_team, q, lane, swz, O, mask0)
#define FZ_DISPATCH(OA_)                  \
  if (doA && doSB) FZ_SA(OA_, true, true);     \
  else if (doA) FZ_SA(OA_, true, false);       \
  else FZ_SA(OA_, false, true);                \
  if (doSB) {                                  \
    FZ_TK(3)                                   \
    named_sync(bar_team, 128);                 \
    FZ_TK(13)                                  \
    fz_step_b<T4, OPL, OA_>(ta, tbB, sC_team, lane, swz, t_empty + 8u * (uint32_t)bufB); \
  }
            if (OPL == 1 || opl == OPL) {
              FZ_DISPATCH(OPL)
            } else {
              FZ_DISPATCH(1)
            }
#undef FZ_DISPATCH
#undef FZ_SA
            FZ_TK(14)
          }
          n_t += (uint32_t)ncap;

          // ================= end of pass: exchange the partial sums through L2 =================
          ++epoch;
          v_target += team_frames;
          if (team_frames != 0) {
            float* mine = Pbuf + (size_t)c * FZ_N * T * OP;
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int k4 = 0; k4 < T4; ++k4)
#pragma unroll
                for (int f = 0; f < FZ_TF; ++f)
                  mine[((size_t)(team * FZ_TF + f) * T + (4 * k4 + q)) * OP + jb * 32 + lane] = ta[jb][k4][f];
          }
          FZ_TK(7)
          named_sync(BAR_MATH, 256);
          if (mtid == 0) red_release_add(cnt_p, 1);
          FZ_TK(8)
          // ---- frame owners: one warp per frame, lane = output capsule ----
          for (int f = c + C * mw; f < FZ_N; f += C * FZ_MATH_WARPS) {
            if (!((item.vmask >> f) & 1u)) continue;
            if (lane == 0) wt.counter(cnt_p, epoch * C, 111);
            __syncwarp();
            FZ_TK(9)
            float y[OPL][T];
#pragma unroll
            for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
              for (int k = 0; k < T; ++k) y[jb][k] = 0.f;
            {
              // two partial sums per round: 2 x T x OPL loads in flight (four per round was measured: it helps
              // units of 27 CTAs by 1 % and costs the cfg-3 shape 2 % through register allocation)
              int cc = 0;
              for (; cc + 1 < C; cc += 2) {
                const float* s0 = Pbuf + ((size_t)cc * FZ_N + f) * T * OP + lane;
                const float* s1 = s0 + (size_t)FZ_N * T * OP;
                float t0[OPL][T], t1[OPL][T];
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                  for (int k = 0; k < T; ++k) {
                    t0[jb][k] = __ldcg(s0 + k * OP + jb * 32);
                    t1[jb][k] = __ldcg(s1 + k * OP + jb * 32);
                  }
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                  for (int k = 0; k < T; ++k) y[jb][k] += t0[jb][k] + t1[jb][k];
              }
              if (cc < C) {
                const float* s0 = Pbuf + ((size_t)cc * FZ_N + f) * T * OP + lane;
#pragma unroll
                for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                  for (int k = 0; k < T; ++k) y[jb][k] += __ldcg(s0 + k * OP + jb * 32);
              }
            }
            FZ_TK(10)
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
            {
              // the output warp of this CTA must have read the row it still needs (the v of the
              // previous last pass) before it is overwritten
              const int prev_last = ((epoch - 1) / iters) * iters;
              if (prev_last > 0 && lane == 0) wt.counter(oflag + f, prev_last, 114);
              __syncwarp();
              float* dst = Vbuf + (size_t)f * T * OP + lane;
#pragma unroll
              for (int jb = 0; jb < OPL; ++jb)
#pragma unroll
                for (int k = 0; k < T; ++k) dst[k * OP + jb * 32] = y[jb][k];
            }
            __syncwarp();
            if (lane == 0) red_release_add(cnt_v + (f >> 4), 1);
            FZ_TK(4)
          }
          FZ_TK(6)
          // ---- everybody: fetch v of the team's frames, update Vacc ----
          // (always wait: the owners must be done with this team's rows of the exchange buffer
          // before the next pass overwrites them)
          if (lane == 0) wt.counter(cnt_v + team, v_target, 112);
          __syncwarp();
          FZ_TK(11)
          if ((last_pass && !p.sdr) || team_frames == 0) {
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
          FZ_TK(12)
#ifdef SRF_FUSED_TIMERS
          if (timing) tacc[15] += 1;
#endif
        }
      }
    }
#ifdef SRF_FUSED_TIMERS
    if (timing) {
#pragma unroll
      for (int z = 0; z < 16; ++z) p.dbg[(size_t)blockIdx.x * 16 + z] += tacc[z];
    }
#endif
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (swarp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------
size_t route_fused_smem_bytes(int OPL, int KC, int x3, int nwst, size_t wstage_bytes) {
  // mode 2 (FP16 images) has one image like tf32 plus the loader's fp32 staging ring: KC chunks of 8
  // halfs hold at most 2 * KC chunks of 4 floats
  const size_t stg = x3 == 2 ? (size_t)FZ_STG * (2 * (size_t)KC) * FZ_N * 16 : 0;
  x3 = x3 == 1;
  const size_t xt = (size_t)KC * FZ_N * 16;
  const size_t XST = x3 ? FZ_XST_MAX / 2 : FZ_XST_MAX;
  return 1024 + stg + (size_t)nwst * wstage_bytes + XST * xt * (x3 ? 2 : 1) +
         sizeof(float) * 2 * (2 * 4 * OPL * 32 * 16 + 2 * OPL * 32 * 16) +
         sizeof(uint64_t) * (2 * (size_t)nwst + 2 * XST + 16) + 64;
}

template <int T4, int OPL, int MODE>
static cudaError_t launch_fused_variant(const FusedParams& p, int grid, size_t smem, cudaStream_t stream,
                                        const void* l2_window, size_t l2_window_bytes, float l2_hit_ratio) {
  auto kern = route_fused_kernel<T4, OPL, MODE>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(FZ_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait on one another
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (l2_window && l2_window_bytes > 0) {
    attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
    attr[1].val.accessPolicyWindow.base_ptr = const_cast<void*>(l2_window);
    attr[1].val.accessPolicyWindow.num_bytes = l2_window_bytes;
    attr[1].val.accessPolicyWindow.hitRatio = l2_hit_ratio;
    attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cfg.numAttrs = 2;
  }
  return cudaLaunchKernelEx(&cfg, kern, p);
}

bool route_fused_supported(int T4, int OPL) {
  return (T4 == 2 && (OPL == 1 || OPL == 2)) || (T4 == 4 && OPL == 1) || (T4 == 5 && OPL == 1);
}

cudaError_t launch_route_fused(const FusedParams& p, int T4, int OPL, int x3, int grid, size_t smem,
                               cudaStream_t stream, const void* l2_window, size_t l2_window_bytes,
                               float l2_hit_ratio) {
#define SRF_FUSED(T4_, OPL_)                                                              \
  if (T4 == T4_ && OPL == OPL_)                                                           \
    return x3 == 1 ? launch_fused_variant<T4_, OPL_, 1>(p, grid, smem, stream, l2_window, l2_window_bytes, l2_hit_ratio) \
         : x3 == 2 ? launch_fused_variant<T4_, OPL_, 2>(p, grid, smem, stream, l2_window, l2_window_bytes, l2_hit_ratio) \
                   : launch_fused_variant<T4_, OPL_, 0>(p, grid, smem, stream, l2_window, l2_window_bytes, l2_hit_ratio);
  SRF_FUSED(2, 1)
  SRF_FUSED(2, 2)
  SRF_FUSED(4, 1)
  SRF_FUSED(5, 1)
#undef SRF_FUSED
  return cudaErrorInvalidValue;
}

}  // namespace srf
