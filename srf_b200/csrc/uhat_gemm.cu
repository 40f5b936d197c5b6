// uhat_gemm.cu -- prediction vectors u_hat = W.x + bias for ALL frames of a layer as a batched
// tensor-core GEMM on sm_100a (reference: tfsr/model/sequence_router_naive.py:150-159,
// einsum form sequence_router_einsum.py:156-157).
//
//   batch = input capsule i (own weight matrix W[i]),  M = rows of W[i] = (j,k) pairs,
//   N = frames,  K = d (input capsule dim).
//
// The contraction has no recurrence, so it runs time-parallel even for SDR; only the routing
// kernel that consumes u_hat is sequential.  Blackwell mapping:
//   * A = W[i] tile (128 rows x K) and B = x tile (64 frames x K) live in shared memory in the
//     canonical K-major no-swizzle UMMA layout [K-chunk of 16 B][row][16 B];
//   * x tiles are fetched by TMA (cp.async.bulk.tensor.5d) straight from emb[B,S,H,d]; the box is
//     (4 floats, NB utterances, NS time steps, 1 capsule, K chunks) and the tensor-map's
//     out-of-bounds ZERO FILL implements the window zero padding of naive:150 (coordinate
//     s - LPAD + w may be negative or >= S) and the K padding;
//   * W[i] arrives with one 1-D bulk copy (pre-packed in the shared-memory image order);
//   * tcgen05.mma kind::tf32 reads the fp32 tiles directly (TF32 truncation in the tensor core),
//     M=128, N=64, K=8 per instruction, fp32 accumulators in TMEM (8 slots of 64 columns);
//   * 8 epilogue warps drain TMEM with tcgen05.ld, add the bias and store u_hat in the layout
//     the routing kernel streams: [frame pair][i][M tile][row][2] (bf16 or fp32), 128 contiguous
//     bytes per warp store.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..9 = epilogue
// (two per TMEM lane quarter, 32 of the 64 columns each).
// Pipelines: x tiles (XSTAGES-deep mbarrier ring), W tile (single buffer, full/empty), TMEM slots
// (8-deep full/empty ring between the MMA issuer and the epilogue).

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "routing_kernels.h"
#include "sm100_ptx.cuh"

namespace srf {

// ---------------------------------------------------------------------------------------
// weight packing for the MMA:  W[I,O,D,d], bias[I,O,D] ->
//   Wm float[i][part][mt][c][r][4]   mt = jb*(T/4) + k4, r = (j%32)*4 + k%4, c = 16-byte K chunk
//   Bm float[i][mt][r]
// part: one tile (W rounded to TF32) or, for the 3 x TF32 split, two: hi = rn_tf32(W) and
// lo = rn_tf32(W - hi)
// ---------------------------------------------------------------------------------------
__global__ void pack_weights_mma_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                        float* __restrict__ Wm, float* __restrict__ Bm, int I, int O,
                                        int D, int d, int T, int OPL, int KC, int x3) {
  const int MT = OPL * (T / 4);
  const int P = x3 ? 2 : 1;
  const long long nW = (long long)I * P * MT * KC * 128 * 4;
  const long long nB = (long long)I * MT * 128;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < nW + nB; e += stride) {
    if (e < nW) {
      const int li = (int)(e & 3);
      long long q = e >> 2;
      const int r = (int)(q % 128);
      q /= 128;
      const int c = (int)(q % KC);
      q /= KC;
      const int mt = (int)(q % MT);
      q /= MT;
      const int part = (int)(q % P);
      const int i = (int)(q / P);
      const int jb = mt / (T / 4), k4 = mt % (T / 4);
      const int j = jb * 32 + r / 4, k = k4 * 4 + (r & 3), l = c * 4 + li;
      float v = 0.f;
      if (j < O && k < D && l < d) v = W[(((long long)i * O + j) * D + k) * d + l];
      // the tensor core truncates fp32 operands to TF32; round to nearest here instead so the
      // weight operand carries no truncation bias
      uint32_t tf;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tf) : "f"(v));
      if (part == 1) {
        const float lo = v - __uint_as_float(tf);  // exact in fp32
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(tf) : "f"(lo));
      }
      Wm[e] = __uint_as_float(tf);
    } else {
      long long q = e - nW;
      const int r = (int)(q % 128);
      q /= 128;
      const int mt = (int)(q % MT);
      const int i = (int)(q / MT);
      const int jb = mt / (T / 4), k4 = mt % (T / 4);
      const int j = jb * 32 + r / 4, k = k4 * 4 + (r & 3);
      float v = 0.f;
      if (j < O && k < D) v = bias[((long long)i * O + j) * D + k];
      Bm[e - nW] = v;
    }
  }
}

void launch_pack_weights_mma(const float* W, const float* bias, float* Wm, float* Bm, int I, int O,
                             int D, int d, int T, int OPL, int KC, int x3, cudaStream_t stream) {
  const long long n = (long long)I * OPL * (T / 4) * 128 * (KC * 4 * (x3 ? 2 : 1) + 1);
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_weights_mma_kernel<<<blocks, 256, 0, stream>>>(W, bias, Wm, Bm, I, O, D, d, T, OPL, KC, x3);
}

// ---------------------------------------------------------------------------------------
// the GEMM
// ---------------------------------------------------------------------------------------
constexpr int UH_N = 64;         // frames per MMA (TMEM columns per slot)
constexpr int UH_SLOTS = 8;      // 8 x 64 = 512 TMEM columns
constexpr int UH_XSTAGES = 4;    // x-tile ring depth
constexpr int UH_EPI_WARPS = 8;   // 2 per TMEM lane quarter, 32 columns each
constexpr int UH_THREADS = (3 + UH_EPI_WARPS) * 32;  // + one splitter warp (3 x TF32 mode)

// X3: 3 x TF32 split.  The A operand arrives as two tiles (hi, lo); a splitter warp rewrites each
// x tile in place as x_hi = x with the 13 low mantissa bits cleared (exactly TF32-representable)
// and writes x_lo = x - x_hi (exact) to a second buffer; the MMA issuer accumulates
// W_hi x_hi + W_lo x_hi + W_hi x_lo in the same TMEM slot: u_hat carries fp32-class error.
// GRP: W[i] resident in groups of M tiles (only where its images exceed shared memory); the
// common GRP = false build has NG = 1 folded at compile time (the epilogue is close to
// instruction bound: the per-item index arithmetic of the grouped form costs it 5 %).
template <bool BF16, bool X3, bool GRP>
__global__ void __launch_bounds__(UH_THREADS, 1)
uhat_gemm_kernel(const __grid_constant__ CUtensorMap tmap_x, const UhatParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int MT = p.MT, KC = p.KC;
  // W[i] is resident MTG M tiles at a time (all of them unless two images would not fit)
  const int MTG = GRP ? p.MTG : MT, NG = GRP ? p.NG : 1;
  const uint32_t a_tile = (uint32_t)MTG * KC * 2048u;   // one image of a group of M tiles
  const uint32_t a_bytes = X3 ? 2u * a_tile : a_tile;   // hi (+ lo)
  const uint32_t x_bytes = (uint32_t)KC * UH_N * 16u;   // one x tile

  uint8_t* sA = smem_raw;
  uint8_t* sX = sA + a_bytes;
  uint8_t* sXlo = sX + (size_t)UH_XSTAGES * x_bytes;    // X3 only
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXlo + (X3 ? (size_t)UH_XSTAGES * x_bytes : 0));
  uint64_t* x_full = bars;                       // [XSTAGES]
  uint64_t* x_empty = x_full + UH_XSTAGES;       // [XSTAGES]
  uint64_t* t_full = x_empty + UH_XSTAGES;       // [SLOTS]
  uint64_t* t_empty = t_full + UH_SLOTS;         // [SLOTS]
  uint64_t* w_full = t_empty + UH_SLOTS;         // [1]
  uint64_t* w_empty = w_full + 1;                // [1]
  uint64_t* x_split = w_empty + 1;               // [XSTAGES] x_hi / x_lo ready (X3)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(x_split + UH_XSTAGES);

  const long long per = (p.items + gridDim.x - 1) / gridDim.x;
  const long long item_lo = (long long)blockIdx.x * per;
  const long long item_hi = item_lo + per < p.items ? item_lo + per : p.items;
  const int ntiles = p.NBT * p.NST;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&tmap_x);
    for (int s = 0; s < UH_XSTAGES; ++s) {
      ptx::mbar_init(&x_full[s], 1);
      ptx::mbar_init(&x_empty[s], 1);
      ptx::mbar_init(&x_split[s], 1);
    }
    for (int s = 0; s < UH_SLOTS; ++s) {
      ptx::mbar_init(&t_full[s], 1);
      ptx::mbar_init(&t_empty[s], UH_EPI_WARPS);  // one arrive per epilogue warp
    }
    ptx::mbar_init(w_full, 1);
    ptx::mbar_init(w_empty, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_ptr, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int cur_ig = -1;
      uint32_t n_w = 0, n_x = 0;
      for (long long item = item_lo; item < item_hi; ++item) {
        const int ig = (int)(item / ntiles), tile = (int)(item % ntiles);
        const int i = NG == 1 ? ig : ig / NG, grp = NG == 1 ? 0 : ig - i * NG;  // no division on the common path
        if (ig != cur_ig) {
          const int mt0 = grp * MTG, mtn = min(MTG, MT - mt0);
          const uint32_t part = (uint32_t)mtn * KC * 2048u;       // bytes of this group's image
          ptx::mbar_wait(w_empty, (n_w & 1) ^ 1);  // previous weights no longer read by the MMAs
          ptx::mbar_arrive_expect_tx(w_full, X3 ? 2u * part : part);
          for (int im = 0; im < (X3 ? 2 : 1); ++im) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.Wm) +
                                 (((size_t)i * (X3 ? 2 : 1) + im) * MT + mt0) * KC * 2048u;
            uint8_t* dst = sA + (size_t)im * a_tile;
            for (uint32_t off = 0; off < part; off += 16384u) {
              const uint32_t n = part - off < 16384u ? part - off : 16384u;
              ptx::bulk_g2s(dst + off, src + off, n, w_full);
            }
          }
          ++n_w;
          cur_ig = ig;
        }
        const int st = n_x % UH_XSTAGES;
        ptx::mbar_wait(&x_empty[st], ((n_x / UH_XSTAGES) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&x_full[st], x_bytes);
        const int w = i / p.H, h = i - w * p.H;
        const int b0 = (tile % p.NBT) * p.NB, s0 = (tile / p.NBT) * p.NS;
        // box (4 floats, NB utterances, NS steps, 1 capsule, KC chunks); OOB -> zeros
        ptx::tma_load_5d(sX + (size_t)st * x_bytes, &tmap_x, &x_full[st], 0, b0, s0 - p.lpad + w, h, 0);
        ++n_x;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_tf32(128, UH_N);
      const uint32_t sA_addr = ptx::smem_u32(sA), sX_addr = ptx::smem_u32(sX);
      int cur_ig = -1;
      uint32_t n_w = 0, n_x = 0, n_t = 0;
      for (long long item = item_lo; item < item_hi; ++item) {
        const int ig = (int)(item / ntiles);
        const int mtn = NG == 1 ? MT : min(MTG, MT - (ig % NG) * MTG);
        if (ig != cur_ig) {
          ptx::mbar_wait(w_full, n_w & 1);
          ++n_w;
          cur_ig = ig;
        }
        const int st = n_x % UH_XSTAGES;
        ptx::mbar_wait(X3 ? &x_split[st] : &x_full[st], (n_x / UH_XSTAGES) & 1);
        ptx::tc_fence_after();
        for (int mt = 0; mt < mtn; ++mt) {
          const int slot = n_t % UH_SLOTS;
          ptx::mbar_wait(&t_empty[slot], ((n_t / UH_SLOTS) & 1) ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_addr = tmem_base + (uint32_t)slot * UH_N;
          for (int ks = 0; ks < KC / 2; ++ks) {
            const uint64_t adesc = ptx::make_smem_desc(
                sA_addr + (uint32_t)mt * KC * 2048u + (uint32_t)ks * 4096u, 2048u, 128u);
            const uint64_t bdesc = ptx::make_smem_desc(
                sX_addr + (uint32_t)st * x_bytes + (uint32_t)ks * (2u * UH_N * 16u), UH_N * 16u, 128u);
            ptx::mma_tf32_ss(d_addr, adesc, bdesc, idesc, ks > 0 ? 1u : 0u);
            if (X3) {
              const uint64_t adesc_lo = ptx::make_smem_desc(
                  sA_addr + a_tile + (uint32_t)mt * KC * 2048u + (uint32_t)ks * 4096u, 2048u, 128u);
              const uint64_t bdesc_lo = ptx::make_smem_desc(
                  ptx::smem_u32(sXlo) + (uint32_t)st * x_bytes + (uint32_t)ks * (2u * UH_N * 16u),
                  UH_N * 16u, 128u);
              ptx::mma_tf32_ss(d_addr, adesc_lo, bdesc, idesc, 1u);
              ptx::mma_tf32_ss(d_addr, adesc, bdesc_lo, idesc, 1u);
            }
          }
          ptx::mma_commit(&t_full[slot]);
          ++n_t;
        }
        ptx::mma_commit(&x_empty[st]);
        ++n_x;
        const bool last_of_group = (item + 1 == item_hi) || ((int)((item + 1) / ntiles) != ig);
        if (last_of_group) ptx::mma_commit(w_empty);
      }
    }
  } else if (warp >= 2 + UH_EPI_WARPS) {
    // ================= splitter (X3): x -> x_hi (in place), x_lo =================
    if (X3) {
      uint32_t n_x = 0;
      const int n4 = (int)(x_bytes / 16u);
      for (long long item = item_lo; item < item_hi; ++item) {
        const int st = n_x % UH_XSTAGES;
        ptx::mbar_wait(&x_full[st], (n_x / UH_XSTAGES) & 1);
        uint4* px = reinterpret_cast<uint4*>(sX + (size_t)st * x_bytes);
        float4* pl = reinterpret_cast<float4*>(sXlo + (size_t)st * x_bytes);
        for (int e = lane; e < n4; e += 32) {
          const uint4 v = px[e];
          const uint4 hi = make_uint4(v.x & 0xffffe000u, v.y & 0xffffe000u, v.z & 0xffffe000u,
                                      v.w & 0xffffe000u);
          px[e] = hi;
          pl[e] = make_float4(__uint_as_float(v.x) - __uint_as_float(hi.x),
                              __uint_as_float(v.y) - __uint_as_float(hi.y),
                              __uint_as_float(v.z) - __uint_as_float(hi.z),
                              __uint_as_float(v.w) - __uint_as_float(hi.w));
        }
        ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&x_split[st]);
        ++n_x;
      }
    }
  } else {
    // ================= epilogue: TMEM -> (+bias) -> u_hat in HBM =================
    const int q = warp & 3;             // TMEM lane quarter this warp may access
    const int ch0 = ((warp - 2) >> 2);  // which 32-column half of the slot
    const int row = q * 32 + lane;
    uint32_t n_t = 0;
    const int halfB = p.Bpad >> 1;
    const int lgNB = 31 - __clz(p.NB);  // NB is a power of two
    constexpr int ES = BF16 ? 2 : 4;    // bytes per stored element
    const long long gstride = (long long)p.I * MT * 256 * ES;  // bytes between frame pairs
    uint8_t* const ubytes = reinterpret_cast<uint8_t*>(p.u);
    for (long long item = item_lo; item < item_hi; ++item) {
      const int ig = (int)(item / ntiles), tile = (int)(item % ntiles);
      // the epilogue is close to instruction bound: keep the integer divisions off the common path
      const int i = NG == 1 ? ig : ig / NG;
      const int mt0 = NG == 1 ? 0 : (ig - i * NG) * MTG;
      const int mtn = NG == 1 ? MT : min(MTG, MT - mt0);
      const int b0 = (tile % p.NBT) * p.NB, s0 = (tile / p.NBT) * p.NS;
      // byte offset of the frame pair and validity of the 16 column pairs this warp owns
      long long gofs[16];
      uint32_t okmask = 0;
#pragma unroll
      for (int pr = 0; pr < 16; ++pr) {
        const int n = ch0 * 32 + pr * 2;
        const int s = s0 + (n >> lgNB), b = b0 + (n & (p.NB - 1));
        gofs[pr] = ((long long)s * halfB + (b >> 1)) * gstride;
        if (s < p.S && b < p.B) okmask |= 1u << pr;
      }
      const bool all_ok = okmask == 0xffffu;
      for (int ml = 0; ml < mtn; ++ml) {
        const int mt = mt0 + ml;
        const int slot = n_t % UH_SLOTS;
        const float bias = __ldg(p.Bm + ((size_t)i * MT + mt) * 128 + row);
        uint8_t* const prow = ubytes + (((size_t)i * MT + mt) * 128 + row) * 2 * ES;
        ptx::mbar_wait(&t_full[slot], (n_t / UH_SLOTS) & 1);
        ptx::tc_fence_after();
        const uint32_t taddr =
            tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)slot * UH_N + (uint32_t)ch0 * 32;
        uint32_t r0[16], r1[16];
        ptx::tmem_ld16(taddr, r0);
        ptx::tmem_ld16(taddr + 16, r1);
        ptx::tmem_ld_wait();
        // the slot can be refilled as soon as its values sit in registers
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&t_empty[slot]);
        ++n_t;
        if (all_ok) {
#pragma unroll
          for (int pr = 0; pr < 16; ++pr) {
            const float v0 = __uint_as_float(pr < 8 ? r0[pr * 2] : r1[(pr - 8) * 2]) + bias;
            const float v1 = __uint_as_float(pr < 8 ? r0[pr * 2 + 1] : r1[(pr - 8) * 2 + 1]) + bias;
            if (BF16)
              *reinterpret_cast<__nv_bfloat162*>(prow + gofs[pr]) = __floats2bfloat162_rn(v0, v1);
            else
              *reinterpret_cast<float2*>(prow + gofs[pr]) = make_float2(v0, v1);
          }
        } else {
#pragma unroll
          for (int pr = 0; pr < 16; ++pr) {
            if (okmask & (1u << pr)) {
              const float v0 = __uint_as_float(pr < 8 ? r0[pr * 2] : r1[(pr - 8) * 2]) + bias;
              const float v1 = __uint_as_float(pr < 8 ? r0[pr * 2 + 1] : r1[(pr - 8) * 2 + 1]) + bias;
              if (BF16)
                *reinterpret_cast<__nv_bfloat162*>(prow + gofs[pr]) = __floats2bfloat162_rn(v0, v1);
              else
                *reinterpret_cast<float2*>(prow + gofs[pr]) = make_float2(v0, v1);
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

size_t uhat_gemm_smem_bytes(int MTG, int KC, int x3) {
  const size_t mul = x3 ? 2 : 1;
  return mul * ((size_t)MTG * KC * 2048 + (size_t)UH_XSTAGES * KC * UH_N * 16) +
         sizeof(uint64_t) * (3 * UH_XSTAGES + 2 * UH_SLOTS + 2) + 16;
}

cudaError_t launch_uhat_gemm(const CUtensorMap& tmap, const UhatParams& p, int num_sms,
                             cudaStream_t stream) {
  const size_t smem = uhat_gemm_smem_bytes(p.MTG, p.KC, p.x3);
  if (p.NG > 1 && !p.x3) return cudaErrorInvalidValue;  // grouped form is built for fp32x3 only
  auto kern = p.x3 ? (p.NG > 1 ? uhat_gemm_kernel<false, true, true> : uhat_gemm_kernel<false, true, false>)
                   : (p.store_bf16 ? uhat_gemm_kernel<true, false, false>
                                   : uhat_gemm_kernel<false, false, false>);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  long long grid = p.items < num_sms ? p.items : num_sms;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, UH_THREADS, smem, stream>>>(tmap, p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------
// unpack u_hat from the streaming layout to the reference's [B,S,I,O,D] (naive:158) --
// used by srf_uhat_fwd (tests / debugging), not by the routing path.
// ---------------------------------------------------------------------------------------
__global__ void unpack_uhat_kernel(const void* __restrict__ u, float* __restrict__ out, int B, int S,
                                   int I, int O, int D, int T, int OPL, int Bpad, int is_bf16) {
  const int MT = OPL * (T / 4);
  const long long total = (long long)B * S * I * O * D;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    long long q = e;
    const int k = (int)(q % D);
    q /= D;
    const int j = (int)(q % O);
    q /= O;
    const int i = (int)(q % I);
    q /= I;
    const int s = (int)(q % S);
    const int b = (int)(q / S);
    const long long g = (long long)s * (Bpad >> 1) + (b >> 1);
    const int mt = (j / 32) * (T / 4) + k / 4;
    const int r = (j % 32) * 4 + (k & 3);
    const size_t src = ((((size_t)g * I + i) * MT + mt) * 128 + r) * 2 + (b & 1);
    out[e] = is_bf16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(u)[src])
                     : reinterpret_cast<const float*>(u)[src];
  }
}

void launch_unpack_uhat(const void* u, float* out, int B, int S, int I, int O, int D, int T, int OPL,
                        int Bpad, int is_bf16, cudaStream_t stream) {
  const long long n = (long long)B * S * I * O * D;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (blocks < 1) blocks = 1;
  unpack_uhat_kernel<<<blocks, 256, 0, stream>>>(u, out, B, S, I, O, D, T, OPL, Bpad, is_bf16);
}

}  // namespace srf
