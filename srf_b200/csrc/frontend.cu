// frontend.cu -- the capsulation front-end (SURVEY.md 8f "next-1"), forward, sm_100a: fbank features
// -> primary capsules emb [B,S,PH,PD], i.e. naive:129-142 / sequence_router.py:44-82 of the reference,
// in four kernels (FP32 CUDA cores: the 1e-4 parity class; the stage-1 convolution is the only part
// with real arithmetic, 2 x 9 x C x C MACs per output position):
//
//   fe_conv_maxout_kernel  one CNN-FE stage: two Conv2D(3x3, stride 2, 'same') paths + bias
//                          (+ training dropout masks) -> maximum -> feat_mask -> (inference:
//                          BatchNormalization with the moving statistics folded to scale/shift ->
//                          feat_mask).  Implicit GEMM on a shared-memory input patch (zero padding
//                          built in), weights staged per (tap, Cin chunk), 4 positions x 4 channels
//                          x 2 paths per thread.
//   fe_bn_stats_kernel /   training only: deterministic per-channel batch statistics over
//   fe_bn_apply_kernel     (B, time, freq) (fixed-order two-level sums, no atomics), normalise in
//                          place, feat_mask, moving-average update.
//   fe_dense_kernel        Dense(PH) on the flattened [Fq*C] frame (+ the einsum variant's sqrt(PH)
//                          scale and positional encoding, einsum:130-131, model_helper.py:30-58).
//   fe_encaps_kernel       per routing frame: two Conv2D(3x3, stride 1, 'same', 1 -> PD) paths over
//                          (time, PH) (+ dropout) -> maximum -> feat_mask -> squash over PD ->
//                          LayerNormalization(PH*PD) -> input dropout -> emb.
//
// The stage-0 activations [B,T/2,F/2,C] are the only large intermediate (written once, read once).

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/srf_b200.h"
#include "routing_kernels.h"

namespace srf {

namespace {

constexpr int FE_THREADS = 256;
constexpr int FE_TR = 2;      // output time rows per CTA of the conv kernel
constexpr int FE_CCH = 32;    // Cin chunk staged per tap
constexpr int FE_DM = 32;     // frames per CTA of the dense kernel
constexpr int FE_DK = 32;     // K chunk of the dense kernel
constexpr int FE_DN = 16;     // output columns per thread of the dense kernel (PH <= 8 * FE_DN)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
// TF 'same' padding before the first element for kernel 3 / stride 2
__host__ __device__ inline int pad_before_s2(int in) {
  const int out = ceil_div(in, 2);
  int total = (out - 1) * 2 + 3 - in;
  if (total < 0) total = 0;
  return total / 2;
}

struct ConvStage {
  const float* x;       // [B,Tin,Fin,Cin]
  const float* k0;      // [3,3,Cin,C]
  const float* k1;
  const float* b0;
  const float* b1;
  const float* drop0;   // [B,Tout,Fout,C] or null
  const float* drop1;
  const float* bn_scale_src[4];  // gamma, beta, mean, var (inference fold) or null when training
  const int32_t* lengths;
  float* y;             // [B,Tout,Fout,C]
  int B, Tin, Fin, Cin, Tout, Fout, C;
  int padT, padF, div;  // feat_mask divisor stride^(stage+1)
  int fold_bn;
  float bn_eps;
};

// ---------------------------------------------------------------------------------------
// one CNN-FE stage (sequence_router.py:71-81)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FE_THREADS) fe_conv_maxout_kernel(const ConvStage p) {
  extern __shared__ __align__(16) float smem[];
  const int b = blockIdx.y;
  const int to0 = blockIdx.x * FE_TR;
  const int rows = 2 * FE_TR + 1;
  const int Fp = 2 * p.Fout + 1;
  const int C4 = (p.C + 3) & ~3;
  float* patch = smem;                                  // [rows][Fp][Cin]
  float* w0 = patch + ((rows * Fp * p.Cin + 3) & ~3);   // [FE_CCH][C4]
  float* w1 = w0 + FE_CCH * C4;

  // input patch with the zero padding of Conv2D(padding='same')
  const int npatch = rows * Fp * p.Cin;
  for (int e = threadIdx.x; e < npatch; e += FE_THREADS) {
    const int ci = e % p.Cin;
    const int pf = (e / p.Cin) % Fp;
    const int r = e / (p.Cin * Fp);
    const int ti = 2 * to0 - p.padT + r;
    const int fi = pf - p.padF;
    float v = 0.f;
    if (ti >= 0 && ti < p.Tin && fi >= 0 && fi < p.Fin)
      v = p.x[(((size_t)b * p.Tin + ti) * p.Fin + fi) * p.Cin + ci];
    patch[e] = v;
  }

  const int P = FE_TR * p.Fout;
  const int npt = ceil_div(P, 4), nct = C4 / 4;
  const int ntiles = npt * nct;
  const int len = p.lengths[b];
  const int nvalid = ceil_div(len < 0 ? 0 : len, p.div);  // feat_mask: frames < ceil(len/div) survive

  for (int tile0 = 0; tile0 < ntiles; tile0 += FE_THREADS) {
    const int tile = tile0 + threadIdx.x;
    const bool live = tile < ntiles;
    const int ct = live ? tile % nct : 0, pt = live ? tile / nct : 0;
    int poff[4];   // patch offset of the position's (kh=0,kw=0,ci=0) element
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int pos = pt * 4 + i;
      if (pos >= P) pos = 0;
      const int tr = pos / p.Fout, fo = pos % p.Fout;
      poff[i] = ((2 * tr) * Fp + 2 * fo) * p.Cin;
    }
    float a0[4][4], a1[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) a0[i][j] = a1[i][j] = 0.f;

    for (int tap = 0; tap < 9; ++tap) {
      const int kh = tap / 3, kw = tap % 3;
      const int toff = (kh * Fp + kw) * p.Cin;
      for (int c0 = 0; c0 < p.Cin; c0 += FE_CCH) {
        const int cc = min(FE_CCH, p.Cin - c0);
        __syncthreads();   // previous chunk consumed (and, first time, the patch is complete)
        for (int e = threadIdx.x; e < cc * C4; e += FE_THREADS) {
          const int c = e % C4, ci = e / C4;
          const size_t g = ((size_t)tap * p.Cin + c0 + ci) * p.C + c;
          w0[e] = c < p.C ? p.k0[g] : 0.f;
          w1[e] = c < p.C ? p.k1[g] : 0.f;
        }
        __syncthreads();
        if (live) {
          for (int ci = 0; ci < cc; ++ci) {
            const float4 wa = *reinterpret_cast<const float4*>(w0 + ci * C4 + ct * 4);
            const float4 wb = *reinterpret_cast<const float4*>(w1 + ci * C4 + ct * 4);
            const float wav[4] = {wa.x, wa.y, wa.z, wa.w};
            const float wbv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float xv = patch[poff[i] + toff + c0 + ci];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                a0[i][j] = fmaf(xv, wav[j], a0[i][j]);
                a1[i][j] = fmaf(xv, wbv[j], a1[i][j]);
              }
            }
          }
        }
      }
    }
    if (!live) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pos = pt * 4 + i;
      if (pos >= P) continue;
      const int to = to0 + pos / p.Fout, fo = pos % p.Fout;
      if (to >= p.Tout) continue;
      const size_t row = (((size_t)b * p.Tout + to) * p.Fout + fo) * p.C;
      const bool keep = to < nvalid;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = ct * 4 + j;
        if (c >= p.C) continue;
        float v0 = a0[i][j] + p.b0[c], v1 = a1[i][j] + p.b1[c];
        if (p.drop0) v0 *= p.drop0[row + c];
        if (p.drop1) v1 *= p.drop1[row + c];
        float v = keep ? fmaxf(v0, v1) : 0.f;
        if (p.fold_bn) {   // BatchNormalization with the moving statistics, then feat_mask again
          const float inv = rsqrtf(p.bn_scale_src[3][c] + p.bn_eps);
          v = (v - p.bn_scale_src[2][c]) * inv * p.bn_scale_src[0][c] + p.bn_scale_src[1][c];
          if (!keep) v = 0.f;
        }
        p.y[row + c] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// training BatchNormalization: per-channel batch statistics, fixed summation order
// ---------------------------------------------------------------------------------------
// partial[blk][0][c] = sum, partial[blk][1][c] = sum of squares over the CTA's strip of rows (double)
__global__ void __launch_bounds__(FE_THREADS) fe_bn_stats_kernel(const float* __restrict__ y, size_t nrows,
                                                                 int C, double* __restrict__ partial) {
  extern __shared__ double shd[];   // [groups][2][C]
  const size_t per = (nrows + gridDim.x - 1) / gridDim.x;
  const size_t r0 = (size_t)blockIdx.x * per, r1 = min(nrows, r0 + per);
  const int groups = C <= FE_THREADS ? FE_THREADS / C : 1;   // row lanes per channel
  if (C <= FE_THREADS) {
    const int c = threadIdx.x % C, g = threadIdx.x / C;
    if (g < groups) {
      double s = 0.0, q = 0.0;
      for (size_t r = r0 + g; r < r1; r += groups) {
        const double v = (double)y[r * C + c];
        s += v;
        q += v * v;
      }
      shd[(g * 2 + 0) * C + c] = s;
      shd[(g * 2 + 1) * C + c] = q;
    }
  } else {
    for (int c = threadIdx.x; c < C; c += FE_THREADS) {
      double s = 0.0, q = 0.0;
      for (size_t r = r0; r < r1; ++r) {
        const double v = (double)y[r * C + c];
        s += v;
        q += v * v;
      }
      shd[c] = s;
      shd[C + c] = q;
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += FE_THREADS) {
    double s = 0.0, q = 0.0;
    for (int g = 0; g < groups; ++g) {
      s += shd[(g * 2 + 0) * C + c];
      q += shd[(g * 2 + 1) * C + c];
    }
    partial[((size_t)blockIdx.x * 2 + 0) * C + c] = s;
    partial[((size_t)blockIdx.x * 2 + 1) * C + c] = q;
  }
}

// finalises the statistics (every CTA, same order -> same bits), normalises its strip in place,
// re-applies feat_mask; CTA 0 also updates the moving statistics
__global__ void __launch_bounds__(FE_THREADS) fe_bn_apply_kernel(
    float* __restrict__ y, size_t nrows, int rows_per_utt_frame, int Tout, int C,
    const double* __restrict__ partial, int nblk, const float* __restrict__ gamma,
    const float* __restrict__ beta, float* __restrict__ mov_mean, float* __restrict__ mov_var,
    const int32_t* __restrict__ lengths, int div, float eps, float momentum) {
  extern __shared__ float sh[];   // scale[C], shift[C]
  const double n = (double)nrows;
  for (int c = threadIdx.x; c < C; c += FE_THREADS) {
    double s = 0.0, q = 0.0;
    for (int k = 0; k < nblk; ++k) {
      s += partial[((size_t)k * 2 + 0) * C + c];
      q += partial[((size_t)k * 2 + 1) * C + c];
    }
    const double mean = s / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float inv = (float)(1.0 / sqrt(var + (double)eps));
    sh[c] = inv * gamma[c];
    sh[C + c] = beta[c] - (float)mean * inv * gamma[c];
    if (blockIdx.x == 0) {
      mov_mean[c] = mov_mean[c] * momentum + (float)mean * (1.f - momentum);
      mov_var[c] = mov_var[c] * momentum + (float)var * (1.f - momentum);
    }
  }
  __syncthreads();
  const size_t total = nrows * (size_t)C;
  for (size_t e = (size_t)blockIdx.x * FE_THREADS + threadIdx.x; e < total;
       e += (size_t)gridDim.x * FE_THREADS) {
    const int c = (int)(e % C);
    const size_t row = e / C;                       // (b, t, f)
    const size_t bt = row / rows_per_utt_frame;     // (b, t)
    const int t = (int)(bt % Tout), b = (int)(bt / Tout);
    const int len = lengths[b];
    const bool keep = t < ceil_div(len < 0 ? 0 : len, div);
    y[e] = keep ? fmaf(y[e], sh[c], sh[C + c]) : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// Dense(PH) "flatten" (naive:131-132) + einsum variant's scale and positional encoding
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FE_THREADS) fe_dense_kernel(const float* __restrict__ x, int M, int K,
                                                              int N, const float* __restrict__ W,
                                                              const float* __restrict__ bias, int S,
                                                              int pos_enc, float* __restrict__ out) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;                        // [FE_DM][FE_DK + 1]
  float* ws = smem + FE_DM * (FE_DK + 1);  // [FE_DK][N]
  const int m0 = blockIdx.x * FE_DM;
  const int r = threadIdx.x >> 3, cg = threadIdx.x & 7;
  float acc[FE_DN];
#pragma unroll
  for (int j = 0; j < FE_DN; ++j) acc[j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += FE_DK) {
    const int kc = min(FE_DK, K - k0);
    __syncthreads();
    for (int e = threadIdx.x; e < FE_DM * FE_DK; e += FE_THREADS) {
      const int kk = e % FE_DK, rr = e / FE_DK;
      xs[rr * (FE_DK + 1) + kk] = (m0 + rr < M && kk < kc) ? x[(size_t)(m0 + rr) * K + k0 + kk] : 0.f;
    }
    for (int e = threadIdx.x; e < kc * N; e += FE_THREADS) ws[e] = W[(size_t)k0 * N + e];
    __syncthreads();
    for (int kk = 0; kk < kc; ++kk) {
      const float xv = xs[r * (FE_DK + 1) + kk];
#pragma unroll
      for (int j = 0; j < FE_DN; ++j) {
        const int c = cg + 8 * j;
        if (c < N) acc[j] = fmaf(xv, ws[kk * N + c], acc[j]);
      }
    }
  }
  const int m = m0 + r;
  if (m >= M) return;
  const float scale = pos_enc ? sqrtf((float)N) : 1.f;
  const int nts = N / 2;
  // model_helper.py:49-56: inv_timescales = exp(-i * log(1e4) / (nts - 1)); signal = [sin | cos]
  const float inc = nts > 1 ? 9.210340371976184f / ((float)nts - 1.f) : 0.f;
  const float pos = (float)(m % S);
#pragma unroll
  for (int j = 0; j < FE_DN; ++j) {
    const int c = cg + 8 * j;
    if (c >= N) continue;
    float v = (acc[j] + bias[c]) * scale;
    if (pos_enc && c < 2 * nts) {
      const int i = c < nts ? c : c - nts;
      const float st = pos * expf((float)i * -inc);
      v += c < nts ? sinf(st) : cosf(st);
    }
    out[(size_t)m * N + c] = v;
  }
}

// ---------------------------------------------------------------------------------------
// encaps convolutions -> maxout -> feat_mask -> squash -> ln_input -> input dropout (naive:133-142)
// one CTA per routing frame
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) fe_encaps_kernel(
    const float* __restrict__ proj, const int32_t* __restrict__ lengths, const float* __restrict__ k0,
    const float* __restrict__ k1, const float* __restrict__ b0, const float* __restrict__ b1,
    const float* __restrict__ drop0, const float* __restrict__ drop1, const float* __restrict__ gamma,
    const float* __restrict__ beta, const float* __restrict__ inp_drop, int S, int PH, int PD, int div,
    float ln_eps, float squash_eps, float* __restrict__ emb) {
  extern __shared__ float sh[];
  float* pr = sh;                    // [3][PH + 2] projected rows s-1, s, s+1 with zero padding
  float* kk = pr + 3 * (PH + 2);     // [2][9][PD]
  float* val = kk + 18 * PD;         // [PH*PD]
  float* red = val + PH * PD;        // [PH] squared norms, then [8] reduction scratch
  const int s = blockIdx.x % S, b = blockIdx.x / S;
  const int n = PH * PD;
  const size_t base = (size_t)blockIdx.x * n;
  for (int e = threadIdx.x; e < 3 * (PH + 2); e += blockDim.x) {
    const int c = e % (PH + 2) - 1, r = s - 1 + e / (PH + 2);
    pr[e] = (c >= 0 && c < PH && r >= 0 && r < S) ? proj[((size_t)b * S + r) * PH + c] : 0.f;
  }
  for (int e = threadIdx.x; e < 9 * PD; e += blockDim.x) {
    kk[e] = k0[e];
    kk[9 * PD + e] = k1[e];
  }
  __syncthreads();
  const int len = lengths[b];
  const bool keep = s < ceil_div(len < 0 ? 0 : len, div);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int pd = e % PD, ph = e / PD;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      const float xv = pr[(t / 3) * (PH + 2) + ph + t % 3];
      v0 = fmaf(xv, kk[t * PD + pd], v0);
      v1 = fmaf(xv, kk[(9 + t) * PD + pd], v1);
    }
    v0 += b0[pd];
    v1 += b1[pd];
    if (drop0) v0 *= drop0[base + e];
    if (drop1) v1 *= drop1[base + e];
    val[e] = keep ? fmaxf(v0, v1) : 0.f;
  }
  __syncthreads();
  for (int ph = threadIdx.x; ph < PH; ph += blockDim.x) {   // squash, naive:248-253
    float n2 = 0.f;
    for (int k = 0; k < PD; ++k) n2 = fmaf(val[ph * PD + k], val[ph * PD + k], n2);
    red[ph] = (n2 / (1.f + n2)) / sqrtf(n2 + squash_eps);
  }
  __syncthreads();
  float s1 = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const float v = val[e] * red[e / PD];
    val[e] = v;
    s1 += v;
  }
  __syncthreads();
  // LayerNormalization over PH*PD (two-pass: mean, then centred variance)
  float* scratch = red;   // PH >= 1; needs 4 + 1 floats: guaranteed by the host (red has max(PH,8))
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s1;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mean += scratch[w];
  mean /= (float)n;
  __syncthreads();
  float s2 = 0.f;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const float dlt = val[e] - mean;
    s2 = fmaf(dlt, dlt, s2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = s2;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) var += scratch[w];
  const float inv = rsqrtf(var / (float)n + ln_eps);
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    float v = (val[e] - mean) * inv * gamma[e] + beta[e];
    if (inp_drop) v *= inp_drop[base + e];
    emb[base + e] = v;
  }
}

}  // namespace

size_t frontend_workspace_bytes(const srf_frontend_desc& d) {
  const size_t T1 = ceil_div(d.T, 2), S = ceil_div((int)T1, 2);
  const size_t F1 = ceil_div(d.F, 2), Fq = ceil_div((int)F1, 2);
  size_t fl = (size_t)d.B * T1 * F1 * d.C + (size_t)d.B * S * Fq * d.C + (size_t)d.B * S * d.PH;
  fl = (fl + 1) & ~(size_t)1;
  fl += (size_t)4 * 1024 * d.C;   // BatchNormalization partial sums (training, double)
  return fl * sizeof(float) + 256;
}

// returns cudaSuccess, or cudaErrorInvalidValue when a shape does not fit the kernels' shared memory
cudaError_t launch_frontend(const srf_frontend_desc& d, float* ws, int max_smem, cudaStream_t stream,
                            int* launches, const char** why) {
  const int T1 = ceil_div(d.T, 2), S = ceil_div(T1, 2);
  const int F1 = ceil_div(d.F, 2), Fq = ceil_div(F1, 2);
  float* y1 = ws;
  float* y2 = y1 + (size_t)d.B * T1 * F1 * d.C;
  float* proj = y2 + (size_t)d.B * S * Fq * d.C;
  size_t off = (size_t)d.B * T1 * F1 * d.C + (size_t)d.B * S * Fq * d.C + (size_t)d.B * S * d.PH;
  off = (off + 1) & ~(size_t)1;
  double* partial = reinterpret_cast<double*>(ws + off);
  *launches = 0;
  for (int st = 0; st < 2; ++st) {
    ConvStage p{};
    p.x = st == 0 ? d.feats : y1;
    p.k0 = d.cnn_kernel[0][st];
    p.k1 = d.cnn_kernel[1][st];
    p.b0 = d.cnn_bias[0][st];
    p.b1 = d.cnn_bias[1][st];
    p.drop0 = d.training ? d.cnn_dropout[0][st] : nullptr;
    p.drop1 = d.training ? d.cnn_dropout[1][st] : nullptr;
    p.bn_scale_src[0] = d.bn_gamma[st];
    p.bn_scale_src[1] = d.bn_beta[st];
    p.bn_scale_src[2] = d.bn_mean[st];
    p.bn_scale_src[3] = d.bn_var[st];
    p.lengths = d.lengths;
    p.y = st == 0 ? y1 : y2;
    p.B = d.B;
    p.Tin = st == 0 ? d.T : T1;
    p.Fin = st == 0 ? d.F : F1;
    p.Cin = st == 0 ? 1 : d.C;
    p.Tout = st == 0 ? T1 : S;
    p.Fout = st == 0 ? F1 : Fq;
    p.C = d.C;
    p.padT = pad_before_s2(p.Tin);
    p.padF = pad_before_s2(p.Fin);
    p.div = st == 0 ? 2 : 4;
    p.fold_bn = d.training ? 0 : 1;
    p.bn_eps = d.bn_eps;
    const int C4 = (d.C + 3) & ~3;
    const size_t patch = ((size_t)(2 * FE_TR + 1) * (2 * p.Fout + 1) * p.Cin + 3) & ~(size_t)3;
    const size_t smem = (patch + (size_t)2 * FE_CCH * C4) * sizeof(float);
    if (smem > (size_t)max_smem) {
      *why = "front-end convolution stage does not fit shared memory (feature dim x filters too large)";
      return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute(fe_conv_maxout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    fe_conv_maxout_kernel<<<dim3(ceil_div(p.Tout, FE_TR), d.B), FE_THREADS, smem, stream>>>(p);
    ++*launches;
    if (d.training) {
      const size_t nrows = (size_t)d.B * p.Tout * p.Fout;
      const int nblk = (int)(nrows < 1024 ? (nrows ? nrows : 1) : 1024);
      const int groups = FE_THREADS / d.C > 0 ? FE_THREADS / d.C : 1;
      const size_t sm_stats = (size_t)groups * 2 * d.C * sizeof(double);
      if (sm_stats > 48 * 1024 || (size_t)2 * d.C * sizeof(float) > 48 * 1024) {
        *why = "too many convolution filters for the BatchNormalization kernels";
        return cudaErrorInvalidValue;
      }
      fe_bn_stats_kernel<<<nblk, FE_THREADS, sm_stats, stream>>>(p.y, nrows, d.C, partial);
      fe_bn_apply_kernel<<<592, FE_THREADS, 2 * d.C * sizeof(float), stream>>>(
          p.y, nrows, p.Fout, p.Tout, d.C, partial, nblk, d.bn_gamma[st], d.bn_beta[st], d.bn_mean[st],
          d.bn_var[st], d.lengths, p.div, d.bn_eps, d.bn_momentum);
      *launches += 2;
    }
  }
  {
    const int M = d.B * S, K = Fq * d.C, N = d.PH;
    if (N > 8 * FE_DN) {
      *why = "model_caps_primary_num > 128 is not supported by the front-end dense kernel";
      return cudaErrorInvalidValue;
    }
    const size_t smem = ((size_t)FE_DM * (FE_DK + 1) + (size_t)FE_DK * N) * sizeof(float);
    fe_dense_kernel<<<ceil_div(M, FE_DM), FE_THREADS, smem, stream>>>(y2, M, K, N, d.dense_kernel,
                                                                     d.dense_bias, S, d.pos_enc, proj);
    ++*launches;
  }
  {
    const int redn = d.PH > 8 ? d.PH : 8;
    const size_t smem = ((size_t)3 * (d.PH + 2) + 18 * d.PD + (size_t)d.PH * d.PD + redn) * sizeof(float);
    if (smem > 48 * 1024) {
      *why = "primary capsule layer too large for the encaps kernel (PH*PD floats must fit 48 KB)";
      return cudaErrorInvalidValue;
    }
    fe_encaps_kernel<<<d.B * S, 128, smem, stream>>>(
        proj, d.lengths, d.encaps_kernel[0], d.encaps_kernel[1], d.encaps_bias[0], d.encaps_bias[1],
        d.training ? d.encaps_dropout[0] : nullptr, d.training ? d.encaps_dropout[1] : nullptr, d.ln_gamma,
        d.ln_beta, d.training ? d.inp_dropout : nullptr, S, d.PH, d.PD, 4, d.ln_eps, d.squash_eps, d.out_emb);
    ++*launches;
  }
  return cudaGetLastError();
}

}  // namespace srf
