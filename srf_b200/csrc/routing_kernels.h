// routing_kernels.h -- internal interface between the C-ABI (capi.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#ifndef SRF_NW
#define SRF_NW 8  // warps per CTA of the fused FP32 layer kernel
#endif

struct srf_frontend_desc;

namespace srf {

struct RouteParams {
  const float* emb;           // [B,S,H,d]
  const float* Wp;            // packed weights (see pack_weights_kernel)
  const float* Bp;            // packed bias
  const float* ln_gamma;      // [O*D] or null
  const float* ln_beta;
  const float* dropout_mask;  // [B,S,O,D] or null
  const float* head_gamma;    // [O] or null
  const float* head_beta;
  float* out_caps;            // [B,S,O,D] or null
  float* out_logits;          // [B,S,O] or null
  float* out_raw;             // [B,S,O,D] pre-LayerNorm capsules or null
  int B, S, H, d, O, D, I;
  int lpad, iters, sdr, mask0;
  int C;        // cluster size (CTAs that split the input capsules of one chain group)
  int Ic;       // input capsules per CTA = ceil(I / C)
  int nchains;  // SDR: B, DR: B*S
  int nsteps;   // SDR: S, DR: 1
  float ln_eps, length_eps;
  const void* u;  // materialised u_hat (streaming modes), else null
  int halfB;      // frame pairs per time step = ceil(B/2)
  int nstage;     // streaming kernel: depth of the shared-memory u_hat ring
  unsigned long long* dbg;  // optional phase timers [CTA][8] (clock64 sums), null = off
};

// backward (routing_bwd.cu)
struct BwdParams {
  const float* emb;
  const float* Wp;
  const float* Bp;
  const float* ln_gamma;
  const float* ln_beta;
  const float* dropout_mask;
  const float* head_gamma;
  const float* v_raw;
  const float* d_out;
  const float* d_logits;
  float* d_raw;
  float* dW;
  float* dbias;
  float* dgamma;
  float* dbeta;
  float* dhead_gamma;
  float* dhead_beta;
  float* d_emb;
  int B, S, H, d, O, D, I;
  int lpad, iters, sdr, mask0, nsteps;
  float ln_eps, length_eps;
  // split mode (no per-element atomics): the BPTT sweep stores, per frame and pass, the coupling
  // coefficients c[i,j], the logit gradients g_a[i,j] and the vectors g_t[j,:], Vacc[j,:]; the
  // frame-parallel dW kernel rebuilds g_u = c g_t + g_a Vacc from them.
  int split;
  int OP;        // padded output capsules (32 * OPL of the BPTT kernel)
  float* cbuf;   // [frames][R][I][OP]
  float* gabuf;  // [frames][R][I][OP]
  float* gtT;    // [frames][R][O][T]
  float* vaT;    // [frames][R][O][T]
  float* dxw;    // [frames][I][OPL][dp]   dL/dx of the windowed input, per 32-capsule chunk
  // phase B (dwdx_from_saved_kernel)
  const float* W;  // canonical [I][O][D][d] weights
  int Tu;          // row length of gtT / vaT (the BPTT kernel's T)
  int dp;          // d rounded up to a multiple of 4 (row length of dxw)
  int FS;          // frame splits of phase B; partial sums go to dwp
  int fps;         // frames per split
  float* dwp;      // [FS][I][O][D][d+1]   partial dW (l < d) and dbias (l == d)
  int tc_phase_b;  // 1: phase B contractions on the tensor cores (TF32 mma.sync; the reduced-precision modes)
  // streamed u_hat (uhat_mode TF32 / BF16): the BPTT kernel reads the tcgen05 GEMM's output
  const void* u;
  int halfB;
  // cluster split of the BPTT sweep: C CTAs per chain, Ic input capsules each
  int C, Ic;
  unsigned long long* dbg;  // optional phase timers [CTA][8] (clock64 sums), null = off
};
int route_layer_bwd_warps(int um, int T, int OPL);
size_t route_layer_bwd_smem_bytes(int T, int OPL, int um, int Ic, int C);
size_t dwdx_smem_bytes(int D, int d, int P, int FT);
int dwdx_frame_splits(const BwdParams& p, int max_smem, int num_sms);
cudaError_t launch_dwdx_from_saved(const BwdParams& p, int max_smem, cudaStream_t stream);
void launch_fold_dx(const BwdParams& p, cudaStream_t stream);
void launch_ln_head_bwd(const BwdParams& p, cudaStream_t stream);
cudaError_t launch_route_layer_bwd(const BwdParams& p, int T, int OPL, int um, int nchains,
                                   cudaStream_t stream);

// CTC + Adam (ctc_adam.cu)
void launch_ctc_greedy(const float* logits, const int32_t* lens, int B, int S, int C, int blank,
                       int32_t* out_ids, int32_t* out_lens, cudaStream_t stream);
cudaError_t launch_ctc_loss(const float* logits, const int32_t* labels, const int32_t* in_lens,
                            const int32_t* lab_lens, int B, int S, int C, int Lmax, int blank,
                            float scale, float* loss, float* d_logits, float* alpha_ws,
                            cudaStream_t stream);
void launch_adam(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                 float beta2, float eps, long long step, cudaStream_t stream);

// u_hat GEMM (uhat_gemm.cu)
struct UhatParams {
  const float* Wm;  // packed A operand  [I][MT][KC][128][4]
  const float* Bm;  // packed bias       [I][MT][128]
  void* u;          // u_hat out         [S*Bpad/2 frame pairs][I][MT][128][2]  bf16 or fp32
  int I, MT, KC;    // KC = 16-byte K chunks per row (2 per tf32 MMA)
  int MTG, NG;      // M tiles resident in shared memory at a time, number of such groups per capsule
  int B, S, H, lpad;
  int NB, NS;       // frame tile = NB utterances x NS time steps (NB*NS = 64)
  int NBT, NST;     // number of tiles along b and s
  int Bpad;         // B rounded up to even
  int store_bf16;
  int x3;           // 3 x TF32 split: Wm holds [hi tile][lo tile] per capsule, x is split in-kernel
  long long items;  // I * NG * NBT * NST
};

void launch_pack_weights(const float* W, const float* bias, float* Wp, float* Bp, int I, int O,
                         int D, int d, int T, int OP, cudaStream_t stream);

size_t route_layer_smem_bytes(int T, int OPL, int F, int NW, int Ic, int um);
int route_layer_max_F(int T, int OPL);
cudaError_t launch_route_layer(const RouteParams& p, int T, int OPL, int F, int groups,
                               size_t smem_bytes, int um, cudaStream_t stream);

// streaming routing kernel over a materialised u_hat (routing_stream.cu)
int route_stream_nslot(int T, int OPL, bool bf16);
size_t route_stream_fixed_smem(int T, int OPL, bool bf16, int C, int max_stages);
size_t route_stream_stage_bytes(int T, int OPL, bool bf16);
cudaError_t launch_route_stream(const RouteParams& p, int T, int OPL, bool bf16, int groups,
                                size_t smem_bytes, cudaStream_t stream);

void launch_pack_weights_mma(const float* W, const float* bias, float* Wm, float* Bm, int I, int O,
                             int D, int d, int T, int OPL, int KC, int x3, cudaStream_t stream);
size_t uhat_gemm_smem_bytes(int MTG, int KC, int x3);
void launch_unpack_uhat(const void* u, float* out, int B, int S, int I, int O, int D, int T, int OPL,
                        int Bpad, int is_bf16, cudaStream_t stream);

// fused routing kernel (routing_fused.cu): packed operand images of W (+ bias in the K padding)
void launch_pack_weights_fused(const float* W, const float* bias, float* Wf, int I, int O, int D, int d,
                               int T4, int OPL, int KC, int parts, cudaStream_t stream);
bool route_fused_supported(int T4, int OPL);
size_t route_fused_smem_bytes(int OPL, int KC, int x3, int nwst, size_t wstage_bytes);

}  // namespace srf

// capsulation front-end (frontend.cu)
namespace srf {
size_t frontend_workspace_bytes(const srf_frontend_desc& d);
cudaError_t launch_frontend(const srf_frontend_desc& d, float* ws, int max_smem, cudaStream_t stream,
                            int* launches, const char** why);
}  // namespace srf

// needs <cuda.h> for CUtensorMap; declared separately so plain users of this header need not include it
#ifdef CUDA_VERSION
namespace srf {
cudaError_t launch_uhat_gemm(const CUtensorMap& tmap, const UhatParams& p, int num_sms,
                             cudaStream_t stream);

// ---- fused routing kernel: device-side descriptors ----
constexpr int FZ_MAX_LAYERS = 16;
struct FusedLayer {
  const float* emb;           // this layer's input capsules [B,S,H,d]
  const float* Wf;            // [i][m][part][KC][128][4]
  const float* ln_gamma;
  const float* ln_beta;
  const float* dropout_mask;
  const float* head_gamma;
  const float* head_beta;
  float* out_caps;
  float* out_logits;
  float* out_raw;
  int H, O, D, opl;           // opl = blocks of 32 output capsules this layer uses
  int KC, KX;                 // 16-byte K chunks per operand row (incl. bias column) / fetched by TMA
  int lpad, rpad, mask0;
  int dep_layer;              // layer whose stored frames this one reads (SDR wavefront), -1 = none
  float ln_eps, length_eps;
};
// one unit of work of one CTA: group `group` of layer `layer`, input capsules [i_lo, i_hi)
struct FusedItem {
  int layer;                  // -1 = idle
  int group, slot;            // frame group; exchange-buffer / counter slot shared by the C CTAs
  int c, C;
  int i_lo, i_hi;
  int b0, s0;                 // first utterance / time step of the group's frame tile
  unsigned vmask;             // bit f: frame f of the tile exists
};
struct FusedParams {
  const FusedLayer* layers;
  const FusedItem* items;     // [rounds][grid]
  int rounds;
  float* xP;                  // [slot][maxC][32][T][OP] partial sums
  float* xV;                  // [slot][32][T][OP]       squashed outputs
  int* cnt_p;                 // [slot]
  int* cnt_v;                 // [slot][2]
  int* oflag;                 // [slot][32] last epoch whose v row the output warp has taken
  int* progress;              // [layer][group][32] frames stored (SDR wavefront), or null
  int* abort_flag;
  int* host_abort;            // mapped host copy of the abort code
  int maxC, ngroups;
  int B, S, sdr, iters, NB;
  int nwst;                   // W ring stages
  int gtiles;                 // W tiles (hi + lo images) per ring stage = one bulk copy
  int wstage_bytes;           // ring stage stride
  int capstage;               // 1: a ring stage holds exactly ONE capsule's tiles; one MMA issuer (warp 8), a
                              // dedicated W producer (warp 9), stages released by the capsule's own commit
  int xtile_bytes;            // x ring stage stride (widest layer)
  int xstg_bytes;             // FP16 images: stride of the x loader's fp32 staging slots
  unsigned long long* dbg;    // optional phase timers [CTA][16] (clock64 sums), null = off
  // The geometry of every CTA's work, as kernel parameters: the MMA issuers derive their smem
  // descriptors from it, and only values that are provably warp-uniform (parameters, blockIdx)
  // stay on ptxas' uniform datapath.  CTA b serves layer l with cta_end[l-1] <= b % per_group <
  // cta_end[l] (SDR; DR has one layer) as the c-th of C_l CTAs of its unit.
  int per_group;
  int n_layers;
  int cta_end[FZ_MAX_LAYERS];
  int layer_I[FZ_MAX_LAYERS];
  int layer_opl[FZ_MAX_LAYERS];
  int layer_KC[FZ_MAX_LAYERS];
  int layer_KX[FZ_MAX_LAYERS];
};
cudaError_t launch_route_fused(const FusedParams& p, int T4, int OPL, int x3, int grid, size_t smem,
                               cudaStream_t stream, const void* l2_window, size_t l2_window_bytes,
                               float l2_hit_ratio);
}
#endif
