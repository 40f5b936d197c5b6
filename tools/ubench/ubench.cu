// ubench.cu -- B200 micro-benchmarks behind the design of the fused routing kernel:
//   l2     : L2 -> shared-memory ingest with cp.async.bulk (per-SM and chip-wide), also multicast x2
//   tmem   : tcgen05.ld throughput per SM for 4 / 8 / 16 reading warps
//   fma    : scalar FFMA vs packed fma.rn.f32x2 throughput per SM
//   dsmem  : st.async pushes and bulk smem->peer-smem copies inside a cluster
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu ; run: ./ubench [which]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
// bounded wait: returns false after ~2^26 polls (never hang the box)
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t par) {
  for (long long n = 0; n < (1ll << 26); ++n) if (mbar_try(b, par)) return true;
  return false;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }

// ------------------------------------------------------------------ l2 ingest
// one warp per CTA; lane 0 drives an NST-stage ring of STAGE bytes; no consumer math.
// mode 0: every CTA walks its own slice of the buffer; mode 1: CTA pairs (2c, 2c+1) walk the SAME slice
// (two frame groups reading the same weights); mode 2: cluster of 2, each CTA loads half a stage and
// multicasts it to both.
template <int MODE>
__global__ void __launch_bounds__(32, 1) l2_kernel(const uint8_t* buf, size_t total, int stage, int nst, int iters, unsigned long long* clk, int* err, int piece) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint8_t* ring = sm + 128;
  const int lane = threadIdx.x;
  if (lane == 0) { for (int s = 0; s < nst; ++s) mbar_init(&full[s], 1); fence_init(); }
  __syncwarp();
  if (MODE == 2) cluster_sync();
  const int walker = MODE == 0 ? blockIdx.x : blockIdx.x / 2;
  const int nwalk = MODE == 0 ? gridDim.x : gridDim.x / 2;
  const size_t slice = (total / nwalk) / stage * stage;
  const uint8_t* base = buf + (size_t)walker * slice;
  const uint32_t rank = MODE == 2 ? cluster_rank() : 0;
  long long t0 = clock64();
  size_t off = 0;
  bool ok = true;
  for (int it = 0; it < iters + nst && ok; ++it) {
    const int st = it % nst;
    if (it >= nst) {  // wait for the copy issued nst iterations ago, then reuse the stage
      if (lane == 0) ok = mbar_wait(&full[st], ((it / nst) - 1) & 1);
      ok = __shfl_sync(0xffffffffu, ok, 0);
      if (MODE == 2 && st == nst - 1) cluster_sync();  // both CTAs drained the whole ring round
    }
    if (it < iters && lane == 0) {
      mbar_expect(&full[st], (uint32_t)stage);
      if (MODE == 2) {
        const uint32_t half = stage / 2;
        for (uint32_t o = 0; o < half; o += 16384u) {
          uint32_t n = half - o < 16384u ? half - o : 16384u;
          bulk_g2s_mc(ring + (size_t)st * stage + rank * half + o, base + off + rank * half + o, n, &full[st], (uint16_t)3);
        }
      } else {
        for (uint32_t o = 0; o < (uint32_t)stage; o += (uint32_t)piece) {
          uint32_t n = stage - o < (uint32_t)piece ? stage - o : (uint32_t)piece;
          bulk_g2s(ring + (size_t)st * stage + o, base + off + o, n, &full[st]);
        }
      }
    }
    off += stage;
    if (off + stage > slice) off = 0;
  }
  long long t1 = clock64();
  if (lane == 0) { clk[blockIdx.x] = (unsigned long long)(t1 - t0); if (!ok) *err = 1; }
  if (MODE == 2) cluster_sync();
}

template <int MODE>
void run_l2(const uint8_t* buf, size_t total, int grid, int stage, int nst, int iters, const char* tag, int piece = 16384) {
  unsigned long long* clk; int* err;
  CK(cudaMalloc(&clk, sizeof(unsigned long long) * grid)); CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  size_t smem = 128 + (size_t)stage * nst;
  CK(cudaFuncSetAttribute(l2_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(32); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = MODE == 2 ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, l2_kernel<MODE>, buf, total, stage, nst, iters, clk, err, piece));
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
  }
  std::vector<unsigned long long> h(grid); int herr;
  CK(cudaMemcpy(h.data(), clk, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  double bytes_per_cta = (double)stage * iters;
  printf("l2 %-10s grid %3d piece %6d stage %6d x%2d iters %5d : %.3f ms  per-SM ingest %.1f B/clk (median CTA), chip ingest %.0f GB/s%s\n", tag, grid, piece, stage, nst, iters, best,
         bytes_per_cta / (double)h[grid / 2], bytes_per_cta * grid / (best * 1e6), herr ? "  [TIMEOUT]" : "");
  CK(cudaFree(clk)); CK(cudaFree(err));
}

// ------------------------------------------------------------------ tmem
template <int X>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tmem_ld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tmem_ld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                 "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
}
template <int X>
__global__ void __launch_bounds__(512, 1) tmem_kernel(int reps, unsigned long long* clk, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    uint32_t v[X];
    tmem_ld<X>(base + (uint32_t)((r * X) & (512 - X)), v);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < X; ++i) acc ^= v[i];
  }
  long long t1 = clock64();
  if (lane == 0) clk[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory"); }
}
// two loads in flight before the wait
template <int X>
__global__ void __launch_bounds__(512, 1) tmem2_kernel(int reps, unsigned long long* clk, uint32_t* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; r += 2) {
    uint32_t v[X], w[X];
    tmem_ld<X>(base + (uint32_t)((r * X) & (512 - X)), v);
    tmem_ld<X>(base + (uint32_t)(((r + 1) * X) & (512 - X)), w);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < X; ++i) acc ^= v[i] + w[i];
  }
  long long t1 = clock64();
  if (lane == 0) clk[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) sink[0] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512) : "memory"); }
}

template <int X, bool TWO>
void run_tmem(int warps) {
  unsigned long long* clk; uint32_t* sink;
  CK(cudaMalloc(&clk, sizeof(unsigned long long) * 148 * 16)); CK(cudaMalloc(&sink, 4));
  const int reps = 4096;
  for (int rep = 0; rep < 2; ++rep) {
    if (TWO) tmem2_kernel<X><<<148, warps * 32>>>(reps, clk, sink); else tmem_kernel<X><<<148, warps * 32>>>(reps, clk, sink);
    CK(cudaDeviceSynchronize());
  }
  std::vector<unsigned long long> h(148 * 16);
  CK(cudaMemcpy(h.data(), clk, sizeof(unsigned long long) * 148 * 16, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int w = 0; w < warps; ++w) mx = std::max(mx, h[w]);
  double bytes = (double)reps * X * 32 * 4 * warps;
  printf("tmem ld 32x32b.x%-2d %s warps %2d : %.1f clk per ld per warp, %.1f B/clk/SM\n", X, TWO ? "2-in-flight" : "1-in-flight", warps, (double)mx / reps, bytes / (double)mx);
  CK(cudaFree(clk)); CK(cudaFree(sink));
}

// ------------------------------------------------------------------ fma
template <bool PACKED>
__global__ void __launch_bounds__(512, 1) fma_kernel(int reps, float seed, unsigned long long* clk, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed * (i + 1) + lane;
  float m0 = seed * 0.5f, m1 = seed * 0.25f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (PACKED) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) {
        unsigned long long d, x, y;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a[i]), "f"(a[i + 1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(m0), "f"(m1));
        asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(m1), "f"(m0));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(x), "l"(y));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(a[i]), "=f"(a[i + 1]) : "l"(d));
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(m0), "f"(m1));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (lane == 0) clk[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
  if (s == 1.2345f) sink[0] = s;
}
template <bool PACKED>
void run_fma(int warps) {
  unsigned long long* clk; float* sink;
  CK(cudaMalloc(&clk, sizeof(unsigned long long) * 148 * 16)); CK(cudaMalloc(&sink, 4));
  const int reps = 8192;
  for (int rep = 0; rep < 2; ++rep) { fma_kernel<PACKED><<<148, warps * 32>>>(reps, 1.0001f, clk, sink); CK(cudaDeviceSynchronize()); }
  std::vector<unsigned long long> h(148 * 16);
  CK(cudaMemcpy(h.data(), clk, sizeof(unsigned long long) * 148 * 16, cudaMemcpyDeviceToHost));
  unsigned long long mx = 0; for (int w = 0; w < warps; ++w) mx = std::max(mx, h[w]);
  printf("fma %s warps %2d : %.1f FMA/clk/SM\n", PACKED ? "f32x2 " : "scalar", warps, (double)reps * 16 * 32 * warps / (double)mx);
  CK(cudaFree(clk)); CK(cudaFree(sink));
}

// ------------------------------------------------------------------ dsmem
// cluster of CS CTAs, 256 threads.  mode 0: every thread pushes float4 with st.async to the next rank;
// mode 1: one thread issues bulk smem -> peer smem copies of 16 KB.  BYTES per CTA per round.
template <int MODE>
__global__ void __launch_bounds__(256, 1) dsmem_kernel(int bytes, int rounds, unsigned long long* clk, int* err) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm);
  uint8_t* src = sm + 128;
  uint8_t* dst = src + bytes;
  uint32_t cs; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(cs));
  const uint32_t rank = cluster_rank(), peer = (rank + 1) % cs;
  if (threadIdx.x == 0) { mbar_init(bar, 1); fence_init(); }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) reinterpret_cast<float*>(src)[i] = (float)i;
  __syncthreads();
  cluster_sync();
  const uint32_t rdst = mapa(smem_u32(dst), peer), rbar = mapa(smem_u32(bar), peer);
  bool ok = true;
  long long t0 = clock64();
  for (int r = 0; r < rounds && ok; ++r) {
    if (threadIdx.x == 0) mbar_expect(bar, (uint32_t)bytes);
    if (MODE == 0) {
      for (int i = threadIdx.x; i < bytes / 16; i += blockDim.x) {
        float4 v = reinterpret_cast<float4*>(src)[i];
        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(rdst + i * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar) : "memory");
      }
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      if (threadIdx.x == 0) {
        for (int o = 0; o < bytes; o += 16384) {
          int n = bytes - o < 16384 ? bytes - o : 16384;
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rdst + o), "r"(smem_u32(src + o)), "r"(n), "r"(rbar) : "memory");
        }
      }
    }
    if (threadIdx.x == 0) ok = mbar_wait(bar, r & 1);
    ok = __syncthreads_and(ok);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { clk[blockIdx.x] = (unsigned long long)(t1 - t0); if (!ok) *err = 1; }
  cluster_sync();
}
template <int MODE>
void run_dsmem(int cs, int bytes) {
  unsigned long long* clk; int* err;
  const int grid = (148 / cs) * cs >= 8 * cs ? 8 * cs : cs;
  CK(cudaMalloc(&clk, sizeof(unsigned long long) * grid)); CK(cudaMalloc(&err, 4)); CK(cudaMemset(err, 0, 4));
  size_t smem = 128 + 2 * (size_t)bytes;
  CK(cudaFuncSetAttribute(dsmem_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const int rounds = 64;
  for (int rep = 0; rep < 2; ++rep) { CK(cudaLaunchKernelEx(&cfg, dsmem_kernel<MODE>, bytes, rounds, clk, err)); CK(cudaDeviceSynchronize()); }
  std::vector<unsigned long long> h(grid); int herr;
  CK(cudaMemcpy(h.data(), clk, sizeof(unsigned long long) * grid, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost));
  std::sort(h.begin(), h.end());
  printf("dsmem %s cluster %d, %6d B per round: %.0f clk per round, %.1f B/clk per CTA (send = receive)%s\n", MODE == 0 ? "st.async.v4" : "bulk s2s   ", cs, bytes,
         (double)h[grid / 2] / rounds, (double)bytes * rounds / (double)h[grid / 2], herr ? "  [TIMEOUT]" : "");
  CK(cudaFree(clk)); CK(cudaFree(err));
}

// ------------------------------------------------------------------ tcgen05.mma issue rate
// one thread issues `reps` MMAs (M=128, N, K=8 tf32, both operands from shared memory) back to back
// and commits; layout: 0 = no swizzle (LBO/SBO core-matrix strides), 2 = 128B swizzle, 6 = 32B swizzle.
__global__ void __launch_bounds__(384, 1) mma_kernel(int N, int layout, int reps, int ntiles, unsigned long long* clk, int noise) {
  __shared__ __align__(8) uint64_t never;
  __shared__ volatile int done;
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&never, 1); done = 0; fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a_addr = smem_u32(sm), b_addr = smem_u32(sm) + 32768;
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      d |= (uint64_t)layout << 61;
      return d;
    };
    uint64_t ads[4], bds[4];
    for (uint32_t ks = 0; ks < 4; ++ks) {
      if (layout == 0) { ads[ks] = desc(a_addr + ks * 4096u, 2048u, 128u); bds[ks] = desc(b_addr + ks * (2u * N * 16u), (uint32_t)N * 16u, 128u); }
      else if (layout == 2) { ads[ks] = desc(a_addr + ks * 32u, 16u, 1024u); bds[ks] = desc(b_addr + ks * 32u, 16u, 1024u); }
      else { ads[ks] = desc(a_addr + ks * 4096u, 16u, 256u); bds[ks] = desc(b_addr + ks * ((uint32_t)N * 32u), 16u, 256u); }
    }
    const uint32_t tmask = (uint32_t)ntiles - 1u;   // ntiles is a power of two
    long long t0 = clock64();
    for (int r = 0; r < reps; r += 4) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t d_addr = tm + (((uint32_t)r >> 2) & tmask) * (uint32_t)N;
        if (ks == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(ads[ks]), "l"(bds[ks]), "r"(idesc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(ads[ks]), "l"(bds[ks]), "r"(idesc) : "memory");
      }
    }
    long long t1 = clock64();
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    clk[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    clk[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
    done = 1;
  } else if (warp >= 4 && noise) {
    uint32_t acc = 0; float f = 1.0f + threadIdx.x;
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16);
    while (!done) {
      if (noise == 1) { acc += mbar_try(&never, 0) ? 1u : 0u; }
      else if (noise == 2) { uint32_t v[16]; tmem_ld<16>(base + 256, v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc ^= v[3]; }
      else if (noise == 4) { uint32_t v[16]; tmem_ld<16>(base + ((warp & 4) ? 16 : 0) + 32 * (acc & 3), v); asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); acc += 1 + (v[3] & 0); }
      else { for (int i = 0; i < 64; ++i) f = fmaf(f, 1.0001f, 0.5f); }
    }
    if (acc == 0x1234567u || f == 1.2345f) clk[1000] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory"); }
}
void run_mma() {
  unsigned long long* clk; CK(cudaMalloc(&clk, 16 * 1024));
  CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  const int reps = 600;
  for (int layout : {0})
    for (int N : {32}) {
      const int ntiles = 512 / N > 4 ? 4 : 512 / N;
      for (int noise = 0; noise < 5; ++noise) {
      for (int rep = 0; rep < 2; ++rep) { mma_kernel<<<148, 384, 65536>>>(N, layout, reps, ntiles, clk, noise); CK(cudaDeviceSynchronize()); }
      unsigned long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
      printf("mma tf32 M=128 N=%3d K=8 layout %d noise %d (0 none, 1 try_wait spin, 2 tcgen05.ld other cols, 3 fma, 4 tcgen05.ld same cols; 8 warps): issue %.1f clk/MMA, issue+retire %.1f clk/MMA  (%.0f MAC/clk/SM)\n", N, layout, noise,
             (double)h[0] / reps, (double)h[1] / reps, 128.0 * N * 8 * reps / (double)h[1]);
      }
    }
  CK(cudaFree(clk));
}

// ------------------------------------------------------------------ tcgen05.mma issue, closer to the routing kernel
// flags: 1 = descriptors advanced per MMA (not loop-invariant), 2 = commit + fresh tile every 3 MMAs,
//        4 = another thread streams bulk copies into other shared memory, 8 = random operand data,
//        16 = 8 warps do FMA + LDS/STS traffic, 32 = wait on a (ready) mbarrier before every tile
__global__ void __launch_bounds__(384, 1) mma2_kernel(int flags, int reps, const uint8_t* gsrc, unsigned long long* clk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar, bar2, ready, cbar[4];
  __shared__ volatile int done;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x)
    reinterpret_cast<float*>(sm)[i] = (flags & 8) ? (float)((i * 2654435761u) >> 8) * (1.0f / 16777216.0f) - 0.5f : 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&ready, 1); for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1); done = 0; fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&ready)) : "memory");
  __syncthreads();
  if ((flags & 128) && warp == 0 && lane != 0) return;   // the issuing warp keeps a single thread
  if ((flags & 64) ? (warp == 0) : (threadIdx.x == 0)) {
    const bool uniform = (flags & 64) != 0;
    const int N = 32;
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    auto desc = [&](uint32_t addr, uint32_t lbo, uint32_t sbo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr >> 4) & 0x3FFF);
      d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
      d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
      d |= (uint64_t)1 << 46;
      return d;
    };
    const uint64_t a0 = desc(smem_u32(sm), 2048u, 128u), b0 = desc(smem_u32(sm) + 65536, 512u, 128u);
    long long t0 = clock64();
    uint32_t ncommit = 0;
    for (int r = 0; r < reps; ++r) {     // one "tile" = 3 MMAs
      if (flags & 32) mbar_wait(&ready, 0);
      const uint32_t tile = (uint32_t)r % 5u;
      uint64_t ad = a0, bd = b0;
      if (flags & 1) { ad += (uint64_t)(tile * 768u); }
      const uint32_t d_addr = tm + ((uint32_t)r & 3u) * 32u;
      uint32_t el = 1;
      if (uniform) asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n" : "=r"(el));
#pragma unroll
      for (int ks = 0; ks < 3; ++ks) {
        if (el) {
        if (ks == 0)
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        else
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_addr), "l"(ad), "l"(bd), "r"(idesc) : "memory");
        }
        if (flags & 1) { ad += 256u; bd += 64u; }
      }
      if (flags & 2) {
        if (el) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&cbar[r & 3])) : "memory");
        ++ncommit;
      }
      if (uniform) __syncwarp();
    }
    long long t1 = clock64();
    if (lane == 0) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    clk[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    clk[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
    done = 1;
    }
  } else if (warp == 1 && lane == 0 && (flags & 4)) {
    // bulk copies into smem [32 KB .. 64 KB) region (not read by the MMAs), 16 KB each, back to back
    uint32_t ph = 0;
    while (!done) {
      mbar_expect(&bar2, 16384u);
      bulk_g2s(sm + 32768 + 0, gsrc, 16384u, &bar2);
      mbar_wait(&bar2, ph);
      ph ^= 1;
    }
  } else if (warp >= 4 && (flags & 16)) {
    float f = 1.0f + threadIdx.x;
    float* sp = reinterpret_cast<float*>(sm + 81920) + threadIdx.x;
    while (!done) {
      for (int i = 0; i < 32; ++i) f = fmaf(f, 1.0001f, 0.5f);
      sp[0] = f;
      f += sp[384];
    }
    if (f == 1.2345f) clk[1000] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory"); }
}
// mma3: per capsule two runtime bases, 15 MMAs fully unrolled with compile-time offsets, lo/hi split descriptors
__device__ __forceinline__ void mma_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc, bool acc) {
  if (acc)
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 0, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(384, 1) mma3_kernel(int variant, int ncaps, unsigned long long* clk) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint32_t tptr;
  __shared__ __align__(8) uint64_t bar, cbar[4];
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 98304 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&cbar[i], 1); fence_init(); }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tptr;
  __shared__ volatile int done3;
  if (threadIdx.x == 0) done3 = 0;
  __syncthreads();
  const int wq = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  if ((threadIdx.x & 31) == 0 && (wq == 0 || (wq == 1 && (variant & 2)))) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t hi = (128u >> 4) | (1u << 14);
    const uint32_t a_lo0 = (((smem_u32(sm) + (uint32_t)wq * 16384u) >> 4) & 0x3FFFu) | ((2048u >> 4) << 16);
    const uint32_t b_lo0 = (((smem_u32(sm) + 65536) >> 4) & 0x3FFFu) | ((512u >> 4) << 16);
    long long t0 = clock64();
    for (int cap = 0; cap < ncaps; ++cap) {
      uint32_t a_base = a_lo0 + (uint32_t)(cap % 3) * 256u;   // varies per capsule
      if (variant & 8) a_base = a_lo0 + (uint32_t)(cap & 1) * 3840u * 0u;   // same base; tiles below are 12 KB apart (streaming 60 KB)
      const uint32_t b_base = b_lo0 + (uint32_t)(cap & 7) * 192u;
      const uint32_t d_base = tm + (uint32_t)(cap % 3) * 160u;
      if (variant & 8) {
        // real tile geometry: tile m at +12 KB (768 x 16 B), K step at +4 KB: 60 KB of A per capsule
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            mma_lo(d_base + m * 32, a_base + m * 768 + ks * 256, b_base + ks * 64, hi, idesc, ks > 0);
      } else if ((variant & 1) == 0) {
#pragma unroll
        for (int m = 0; m < 5; ++m)
#pragma unroll
          for (int ks = 0; ks < 3; ++ks)
            mma_lo(d_base + m * 32, a_base + (m & 0) * 768 + (ks & 0) * 256, b_base + ks * 64, hi, idesc, ks > 0);
      } else {
        // runtime split of the capsule's tiles over two ring stages (select per tile)
        const int split = 1 + (cap & 3);
        const uint32_t a_base2 = a_lo0 + 3072u - (uint32_t)split * 768u;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
          const uint32_t ab = (m < split ? a_base : a_base2) + m * 768;
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) mma_lo(d_base + m * 32, ab + ks * 256, b_base + ks * 64, hi, idesc, ks > 0);
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&cbar[(cap & 1) + 2 * wq])) : "memory");
    }
    long long t1 = clock64();
    if (wq == 0) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    clk[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
    clk[blockIdx.x * 2 + 1] = (unsigned long long)(t2 - t0);
    done3 = 1;
    }
  } else if (wq >= 4 && (variant & 4)) {
    float f = 1.0f + threadIdx.x, g = 0.5f;
    const uint32_t base = tm + ((uint32_t)((wq & 3) * 32) << 16) + ((wq & 4) ? 16u : 0u);
    while (!done3) {
      uint32_t v[16], w[16];
      tmem_ld<16>(base + 32, v); tmem_ld<16>(base + 64, w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 16; ++i) { f = fmaf(__uint_as_float(v[i]), g, f); g = fmaf(__uint_as_float(w[i]), f, g); }
    }
    if (f == 1.2345f) clk[1000] = 1;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory"); }
}
void run_mma3() {
  unsigned long long* clk; CK(cudaMalloc(&clk, 16 * 1024));
  CK(cudaFuncSetAttribute(mma3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  const int ncaps = 100;
  for (int variant : {0, 8, 10}) {
    for (int rep = 0; rep < 2; ++rep) { mma3_kernel<<<148, 384, 98304>>>(variant, ncaps, clk); CK(cudaDeviceSynchronize()); }
    unsigned long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    printf("mma3 variant %d (15 MMAs per capsule; 0 = every MMA re-reads the SAME 4 KB of A, 8 = A streams over 60 KB like the real tiles, +2 = second issuing warp): issue %.1f clk/MMA, issue+retire %.1f clk/MMA%s\n", variant,
           "", (double)h[0] / (15 * ncaps), (double)h[1] / (15 * ncaps));
  }
  CK(cudaFree(clk));
}

void run_mma2() {
  unsigned long long* clk; CK(cudaMalloc(&clk, 16 * 1024));
  uint8_t* gsrc; CK(cudaMalloc(&gsrc, 1 << 20)); CK(cudaMemset(gsrc, 0, 1 << 20));
  CK(cudaFuncSetAttribute(mma2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
  const int reps = 300;
  for (int flags : {0, 3, 128, 131, 131 + 16}) {
    for (int rep = 0; rep < 2; ++rep) { mma2_kernel<<<148, 384, 98304>>>(flags, reps, gsrc, clk); CK(cudaDeviceSynchronize()); }
    unsigned long long h[2]; CK(cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost));
    printf("mma2 flags %2d: issue %.1f clk/MMA, issue+retire %.1f clk/MMA\n", flags, (double)h[0] / (3 * reps), (double)h[1] / (3 * reps));
  }
  CK(cudaFree(clk)); CK(cudaFree(gsrc));
}

// ------------------------------------------------------------------ redux / ex2 dependent-chain latency
__global__ void __launch_bounds__(256, 1) chain_kernel(int reps, int which, unsigned long long* clk, float* sink) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int v = lane + 1;
  float f = 1.0f + lane * 0.001f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    if (which == 0) v = __reduce_max_sync(0xffffffffu, v + r);
    else if (which == 1) v = (int)__reduce_add_sync(0xffffffffu, (unsigned)(v & 0xffff));
    else if (which == 2) f += __shfl_xor_sync(0xffffffffu, f, 16);
    else if (which == 3) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f)); f = y * 0.5f; }
    else { float y; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(f)); f = y + 1.5f; }
  }
  long long t1 = clock64();
  if (lane == 0) clk[blockIdx.x * 16 + warp] = (unsigned long long)(t1 - t0);
  if (v == 123456789 || f == 1.2345f) sink[0] = f + v;
}
void run_chain() {
  unsigned long long* clk; float* sink;
  CK(cudaMalloc(&clk, sizeof(unsigned long long) * 16)); CK(cudaMalloc(&sink, 4));
  const char* names[5] = {"REDUX.MAX", "REDUX.SUM", "SHFL+FADD", "EX2+FMUL", "RCP+FADD"};
  for (int warps : {1, 8})
    for (int w = 0; w < 5; ++w) {
      chain_kernel<<<1, warps * 32>>>(4096, w, clk, sink); CK(cudaDeviceSynchronize());
      chain_kernel<<<1, warps * 32>>>(4096, w, clk, sink); CK(cudaDeviceSynchronize());
      unsigned long long h[16]; CK(cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost));
      printf("chain %-10s warps %d: %.1f clk per dependent op\n", names[w], warps, (double)h[0] / 4096);
    }
  CK(cudaFree(clk)); CK(cudaFree(sink));
}

int main(int argc, char** argv) {
  const char* which = argc > 1 ? argv[1] : "all";
  auto want = [&](const char* s) { return !strcmp(which, "all") || !strcmp(which, s); };
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
  if (want("fma")) {
    for (int w : {4, 8, 16}) { run_fma<false>(w); run_fma<true>(w); }
  }
  if (want("tmem")) {
    for (int w : {4, 8, 16}) { run_tmem<16, false>(w); run_tmem<32, false>(w); run_tmem<32, true>(w); }
  }
  if (want("l2")) {
    const size_t total = 84ull << 20;
    uint8_t* buf; CK(cudaMalloc(&buf, total)); CK(cudaMemset(buf, 1, total));
    for (int grid : {1, 16, 37, 74, 148}) run_l2<0>(buf, total, grid, 49152, 4, 2000, "own-slice");
    run_l2<0>(buf, total, 148, 16384, 8, 6000, "own-slice");
    run_l2<0>(buf, total, 148, 12288, 16, 8000, "own-slice", 12288);
    run_l2<0>(buf, total, 148, 12288, 8, 8000, "own-slice", 12288);
    run_l2<0>(buf, total, 148, 24576, 8, 4000, "own-slice", 24576);
    run_l2<0>(buf, total, 148, 49152, 4, 2000, "own-slice", 49152);
    run_l2<0>(buf, total, 148, 61440, 3, 2000, "own-slice", 61440);
    run_l2<0>(buf, total, 148, 61440, 3, 2000, "own-slice", 12288);
    run_l2<0>(buf, total, 148, 4096, 32, 16000, "own-slice", 4096);
    run_l2<1>(buf, total, 148, 49152, 4, 2000, "pair-same");
    run_l2<2>(buf, total, 148, 49152, 4, 2000, "multicast2");
    run_l2<2>(buf, total, 74, 49152, 4, 2000, "multicast2");
    // working set larger than L2 for comparison (HBM-bound)
    const size_t big = 1024ull << 20;
    uint8_t* buf2; CK(cudaMalloc(&buf2, big)); CK(cudaMemset(buf2, 1, big));
    run_l2<0>(buf2, big, 148, 49152, 4, 100, "hbm");
    CK(cudaFree(buf)); CK(cudaFree(buf2));
  }
  if (want("chain")) run_chain();
  if (want("mma")) run_mma();
  if (want("mma2")) run_mma2();
  if (want("mma3")) run_mma3();
  if (want("dsmem")) {
    for (int cs : {2, 4, 8}) { run_dsmem<0>(cs, 32768); run_dsmem<1>(cs, 32768); run_dsmem<1>(cs, 65536); }
  }
  return 0;
}
