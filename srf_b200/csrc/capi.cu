// capi.cu -- the C-ABI of include/srf_b200.h: handle, argument validation, packed-weight
// cache, kernel-variant dispatch and the per-layer stack driver.  No torch types here.

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/srf_b200.h"
#include "routing_kernels.h"

namespace {

struct PackedWeights {
  const float* W = nullptr;
  const float* bias = nullptr;
  int I = 0, O = 0, D = 0, d = 0, T = 0, OP = 0;
  uint64_t version = 0;
  int x3 = 0;
  float* Wp = nullptr;
  float* Bp = nullptr;
  size_t bytes = 0;
};

}  // namespace

struct srf_handle {
  int device = 0;
  int num_sms = 148;
  int max_smem = 227 * 1024;
  std::string error;
  std::string last_kernel;
  int64_t launches = 0;
  std::vector<PackedWeights> packed;
  float* ws[2] = {nullptr, nullptr};
  size_t ws_bytes = 0;
  int force_F = 0, force_C = 0, no_stream = 0, max_stages = 0;
  float* bwd_ws = nullptr;  // split-mode backward scratch
  size_t bwd_ws_bytes = 0;
  int bwd_atomics = 0;
  float* ctc_ws = nullptr;  // alpha workspace of srf_ctc_loss
  size_t ctc_ws_bytes = 0;
  float* fe_ws = nullptr;   // front-end intermediates of srf_capsulate_fwd
  size_t fe_ws_bytes = 0;
  unsigned long long* dbg = nullptr;  // SRF_PHASE_TIMERS=1: per-CTA phase timers of the streaming kernel
  // tensor-core path
  std::vector<PackedWeights> packed_mma;
  void* ubuf = nullptr;  // materialised u_hat of one layer
  size_t ubuf_bytes = 0;
  void* encode_tiled = nullptr;  // cuTensorMapEncodeTiled
  // fused routing kernel (routing_fused.cu)
  std::vector<PackedWeights> packed_fused;
  // packed weights of the last multi-layer fused stack, in ONE allocation: the kernel re-streams them
  // from L2 every time step, an access-policy window keeps them resident there
  float* fz_arena = nullptr;
  size_t fz_arena_bytes = 0;
  std::vector<PackedWeights> fz_arena_keys;
  size_t l2_persist_max = 0, l2_window_max = 0;
  bool l2_carved = false;   // the persisting-L2 carve-out is currently reserved (fused inference path)
  // the stream of the last compute call (see enter_stream)
  cudaStream_t last_stream = nullptr;
  bool have_stream = false;
  cudaEvent_t xev = nullptr;
  int no_fused = 0, force_fused = 0;
  void* fz_tab = nullptr;        // device: FusedLayer[] + FusedItem[] + counters + progress
  size_t fz_tab_bytes = 0;
  float* fz_x = nullptr;         // device: exchange buffers (partial sums, squashed outputs)
  size_t fz_x_bytes = 0;
  std::vector<float*> fz_inter;  // inter-layer capsule buffers of a wavefront launch
  size_t fz_inter_bytes = 0;
  int* fz_host_abort = nullptr;  // mapped host word: abort code of a timed-out fused launch
  int* fz_host_abort_dev = nullptr;
  // per-kernel timing (srf_profile_begin/end)
  bool profiling = false;
  struct Span {
    int kind;
    cudaEvent_t a, b;
  };
  std::vector<Span> spans;
};

namespace {
// brackets one kernel launch with events when profiling is on
struct KernelSpan {
  srf_handle* h;
  cudaStream_t stream;
  cudaEvent_t a = nullptr, b = nullptr;
  int kind;
  KernelSpan(srf_handle* h_, int kind_, cudaStream_t s) : h(h_), stream(s), kind(kind_) {
    if (!h->profiling) return;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, stream);
  }
  ~KernelSpan() {
    if (!a) return;
    cudaEventRecord(b, stream);
    h->spans.push_back({kind, a, b});
  }
};
}  // namespace

static std::string g_create_error;

static int fail(srf_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h)
    h->error = buf;
  else
    g_create_error = buf;
  return code;
}

static int cuda_fail(srf_handle* h, cudaError_t e, const char* what) {
  return fail(h, (int)e, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
}

namespace {
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    cudaGetDevice(&cur);
    if (prev >= 0 && cur != prev) cudaSetDevice(prev);
  }
};
}  // namespace

extern "C" int srf_version(void) { return SRF_B200_VERSION; }

extern "C" int srf_create(int device, srf_handle** out) {
  if (!out) return fail(nullptr, -1, "srf_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "srf_create: cudaGetDeviceCount");
  if (device < 0 || device >= ndev)
    return fail(nullptr, -2, "srf_create: device %d out of range (%d CUDA devices)", device, ndev);
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "srf_create: cudaGetDeviceProperties");
  if (prop.major != 10)
    return fail(nullptr, -3,
                "srf_create: device %d is sm_%d%d; this library is built for sm_100a (B200) only "
                "and has no fallback path",
                device, prop.major, prop.minor);
  srf_handle* h = new srf_handle();
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  h->max_smem = (int)prop.sharedMemPerBlockOptin;
  if (const char* s = getenv("SRF_FORCE_F")) h->force_F = atoi(s);
  if (const char* s = getenv("SRF_FORCE_C")) h->force_C = atoi(s);
  if (const char* s = getenv("SRF_NO_STREAM")) h->no_stream = atoi(s);
  if (const char* s = getenv("SRF_STREAM_STAGES")) h->max_stages = atoi(s);
  if (const char* s = getenv("SRF_BWD_ATOMICS")) h->bwd_atomics = atoi(s);
  if (const char* s = getenv("SRF_NO_FUSED")) h->no_fused = atoi(s);
  if (const char* s = getenv("SRF_FORCE_FUSED")) h->force_fused = atoi(s);
  h->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
  h->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
  // the persisting carve-out takes its share of L2 away from ALL normal traffic of the device (measured:
  // the two-kernel fp32x3 path fell from 42 to 65 ms per cfg-3 step with it set), so it is reserved lazily
  // by the fused inference path (l2_carve) and given back by every other path; SRF_L2_WINDOW=0: never
  if (const char* s = getenv("SRF_L2_WINDOW"))
    if (s[0] == '0') h->l2_persist_max = 0;
  if (cudaHostAlloc((void**)&h->fz_host_abort, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
    *h->fz_host_abort = 0;
    if (cudaHostGetDevicePointer((void**)&h->fz_host_abort_dev, h->fz_host_abort, 0) != cudaSuccess)
      h->fz_host_abort_dev = nullptr;
  } else {
    cudaGetLastError();
    h->fz_host_abort = nullptr;
  }
  if (const char* s = getenv("SRF_PHASE_TIMERS")) {
    if (atoi(s) > 0 && cudaMalloc((void**)&h->dbg, 1024 * 8 * sizeof(unsigned long long)) == cudaSuccess)
      cudaMemset(h->dbg, 0, 1024 * 8 * sizeof(unsigned long long));
  }
  *out = h;
  return 0;
}

static void l2_release(srf_handle* h);
// One stream at a time per handle: the scratch buffers (u_hat, workspaces, packed-weight caches, exchange
// buffers) are reused, regrown and freed stream-ordered.  A call on another stream than the previous one first
// waits for everything the handle enqueued there, so two streams can share a handle without overwriting or
// freeing each other's scratch (they serialise); concurrent use from two host threads still needs two handles.
static void enter_stream(srf_handle* h, cudaStream_t s) {
  if (h->have_stream && h->last_stream != s) {
    if (!h->xev && cudaEventCreateWithFlags(&h->xev, cudaEventDisableTiming) != cudaSuccess) h->xev = nullptr;
    if (h->xev && cudaEventRecord(h->xev, h->last_stream) == cudaSuccess) cudaStreamWaitEvent(s, h->xev, 0);
    cudaGetLastError();   // a stream the caller has destroyed meanwhile has nothing left to wait for
  }
  h->last_stream = s;
  h->have_stream = true;
}
extern "C" int srf_destroy(srf_handle* h) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  l2_release(h);
  if (h->xev) cudaEventDestroy(h->xev);
  for (auto& pw : h->packed) {
    if (pw.Wp) cudaFree(pw.Wp);
  }
  for (auto& pw : h->packed_mma) {
    if (pw.Wp) cudaFree(pw.Wp);
  }
  if (h->ubuf) cudaFree(h->ubuf);
  for (auto& pw : h->packed_fused) {
    if (pw.Wp) cudaFree(pw.Wp);
  }
  if (h->fz_arena) cudaFree(h->fz_arena);
  if (h->fz_tab) cudaFree(h->fz_tab);
  if (h->fz_x) cudaFree(h->fz_x);
  for (float* b : h->fz_inter)
    if (b) cudaFree(b);
  if (h->fz_host_abort) cudaFreeHost(h->fz_host_abort);
  if (h->dbg) cudaFree(h->dbg);
  if (h->ctc_ws) cudaFree(h->ctc_ws);
  if (h->fe_ws) cudaFree(h->fe_ws);
  if (h->bwd_ws) cudaFree(h->bwd_ws);
  for (int i = 0; i < 2; ++i)
    if (h->ws[i]) cudaFree(h->ws[i]);
  delete h;
  return 0;
}

extern "C" const char* srf_last_error(const srf_handle* h) {
  return h ? h->error.c_str() : g_create_error.c_str();
}

extern "C" int srf_ctc_greedy_decode(srf_handle* h, const float* logits, const int32_t* lens, int32_t B,
                                     int32_t S, int32_t C, int32_t blank, int32_t* out_ids,
                                     int32_t* out_lens, void* stream) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (B == 0) return 0;
  if (!logits || !lens || !out_ids || !out_lens) return fail(h, -1, "NULL argument");
  if (B < 0 || S <= 0 || C <= 0 || blank < 0 || blank >= C) return fail(h, -2, "bad shape or blank index");
  DeviceGuard g(h->device);
  srf::launch_ctc_greedy(logits, lens, B, S, C, blank, out_ids, out_lens, (cudaStream_t)stream);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "ctc_greedy launch");
  return 0;
}

extern "C" int srf_ctc_loss(srf_handle* h, const float* logits, const int32_t* labels,
                            const int32_t* in_lens, const int32_t* lab_lens, int32_t B, int32_t S,
                            int32_t C, int32_t Lmax, int32_t blank, float grad_scale, float* loss,
                            float* d_logits, void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (B == 0) return 0;
  if (!logits || !labels || !in_lens || !lab_lens || !loss || !d_logits) return fail(h, -1, "NULL argument");
  if (B < 0 || S <= 0 || C <= 1 || Lmax <= 0 || blank < 0 || blank >= C)
    return fail(h, -2, "bad shape or blank index");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  enter_stream(h, stream);
  const size_t need = (size_t)B * S * (2 * (size_t)Lmax + 1) * sizeof(float);
  if (need > h->ctc_ws_bytes) {
    if (h->ctc_ws) cudaFreeAsync(h->ctc_ws, stream);
    h->ctc_ws = nullptr;
    h->ctc_ws_bytes = 0;
    cudaError_t e = cudaMallocAsync((void**)&h->ctc_ws, need, stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "CTC workspace allocation");
    h->ctc_ws_bytes = need;
  }
  cudaError_t e = srf::launch_ctc_loss(logits, labels, in_lens, lab_lens, B, S, C, Lmax, blank, grad_scale,
                                       loss, d_logits, h->ctc_ws, stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "ctc_loss launch");
  h->launches++;
  return 0;
}

extern "C" int srf_adam_step(srf_handle* h, float* param, const float* grad, float* m, float* v,
                             int64_t n, float lr, float beta1, float beta2, float eps, int64_t step,
                             void* stream) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (n == 0) return 0;
  if (!param || !grad || !m || !v) return fail(h, -1, "NULL argument");
  if (n < 0 || step < 1) return fail(h, -2, "n must be >= 0 and step >= 1");
  DeviceGuard g(h->device);
  srf::launch_adam(param, grad, m, v, n, lr, beta1, beta2, eps, step, (cudaStream_t)stream);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "adam launch");
  return 0;
}

extern "C" int srf_capsulate_fwd(srf_handle* h, const srf_frontend_desc* fe, void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!fe) return fail(h, -1, "front-end descriptor is NULL");
  const srf_frontend_desc& d = *fe;
  if (d.B == 0) return 0;
  if (d.B < 0 || d.T <= 0 || d.F <= 0 || d.C <= 0 || d.PH <= 0 || d.PD <= 0)
    return fail(h, -2, "front-end: bad shape (B=%d T=%d F=%d C=%d PH=%d PD=%d)", d.B, d.T, d.F, d.C, d.PH, d.PD);
  if (!d.feats || !d.lengths || !d.out_emb || !d.dense_kernel || !d.dense_bias || !d.ln_gamma || !d.ln_beta)
    return fail(h, -1, "front-end: NULL argument");
  for (int a = 0; a < 2; ++a) {
    if (!d.encaps_kernel[a] || !d.encaps_bias[a] || !d.bn_gamma[a] || !d.bn_beta[a] || !d.bn_mean[a] ||
        !d.bn_var[a])
      return fail(h, -1, "front-end: NULL argument");
    for (int b = 0; b < 2; ++b)
      if (!d.cnn_kernel[a][b] || !d.cnn_bias[a][b]) return fail(h, -1, "front-end: NULL argument");
  }
  if (d.pos_enc && (d.PH & 1))
    return fail(h, -2, "front-end: the positional encoding needs an even model_caps_primary_num");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  enter_stream(h, stream);
  const size_t need = srf::frontend_workspace_bytes(d);
  if (need > h->fe_ws_bytes) {
    if (h->fe_ws) cudaFreeAsync(h->fe_ws, stream);
    h->fe_ws = nullptr;
    h->fe_ws_bytes = 0;
    cudaError_t e = cudaMallocAsync((void**)&h->fe_ws, need, stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "front-end workspace allocation");
    h->fe_ws_bytes = need;
  }
  int launches = 0;
  const char* why = "";
  cudaError_t e = srf::launch_frontend(d, h->fe_ws, h->max_smem, stream, &launches, &why);
  h->launches += launches;
  if (e == cudaErrorInvalidValue && why[0]) {
    cudaGetLastError();
    return fail(h, -3, "front-end: %s", why);
  }
  if (e != cudaSuccess) return cuda_fail(h, e, "front-end launch");
  h->last_kernel = "fe_conv_maxout_kernel x2 + fe_dense_kernel + fe_encaps_kernel";
  return 0;
}

extern "C" int srf_profile_begin(srf_handle* h) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  for (auto& sp : h->spans) {
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  h->spans.clear();
  h->profiling = true;
  return 0;
}

extern "C" int srf_profile_end(srf_handle* h, float* ms, int32_t* launches) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!ms || !launches) return fail(h, -1, "ms / launches is NULL");
  DeviceGuard g(h->device);
  h->profiling = false;
  for (int k = 0; k < 3; ++k) {
    ms[k] = 0.f;
    launches[k] = 0;
  }
  int rc = 0;
  for (auto& sp : h->spans) {
    cudaError_t e = cudaEventSynchronize(sp.b);
    float t = 0.f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&t, sp.a, sp.b);
    if (e != cudaSuccess && rc == 0) rc = cuda_fail(h, e, "srf_profile_end");
    if (sp.kind >= 0 && sp.kind < 3) {
      ms[sp.kind] += t;
      launches[sp.kind]++;
    }
    cudaEventDestroy(sp.a);
    cudaEventDestroy(sp.b);
  }
  h->spans.clear();
  return rc;
}

// debug aid (not part of the public header): copy the phase timers to the host and clear them
extern "C" int srf_debug_phase_timers(srf_handle* h, unsigned long long* out, int n_cta) {
  if (!h || !h->dbg) return -1;
  DeviceGuard g(h->device);
  cudaDeviceSynchronize();
  if (n_cta > 1024) n_cta = 1024;
  cudaMemcpy(out, h->dbg, (size_t)n_cta * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  cudaMemset(h->dbg, 0, 1024 * 8 * sizeof(unsigned long long));
  return 0;
}

extern "C" int64_t srf_launch_count(const srf_handle* h) { return h ? h->launches : 0; }

extern "C" const char* srf_last_kernel(const srf_handle* h) {
  return h ? h->last_kernel.c_str() : "";
}

// the checks of validate_layer that decide eligibility for a fused stack, without an error message
static bool validate_layer_quiet(const srf_layer_desc* L, bool need_emb) {
  if (!L || !L->W || !L->bias || (need_emb && !L->emb)) return false;
  if (L->B < 0 || L->S < 0 || L->H <= 0 || L->d <= 0 || L->O <= 0 || L->D <= 0) return false;
  if (L->lpad < 0 || L->rpad < 0 || L->iters < 1) return false;
  return true;
}

static int validate_layer(srf_handle* h, const srf_layer_desc* L, bool fwd = true) {
  if (!L) return fail(h, -1, "layer descriptor is NULL");
  if (!L->emb || !L->W || !L->bias) return fail(h, -1, "emb, W and bias must be non-NULL");
  if (L->B < 0 || L->S < 0) return fail(h, -2, "negative B or S");
  if (L->H <= 0 || L->d <= 0 || L->O <= 0 || L->D <= 0)
    return fail(h, -2, "H, d, O, D must be positive (got %d %d %d %d)", L->H, L->d, L->O, L->D);
  if (L->lpad < 0 || L->rpad < 0) return fail(h, -2, "lpad/rpad must be >= 0");
  if (L->iters < 1) return fail(h, -2, "iters must be >= 1 (got %d)", L->iters);
  if (L->O < 2 && L->mask_class0)
    return fail(h, -2, "mask_class0 needs at least 2 output capsules");
  if ((L->ln_gamma == nullptr) != (L->ln_beta == nullptr))
    return fail(h, -1, "ln_gamma and ln_beta must both be given or both be NULL");
  if ((L->head_gamma == nullptr) != (L->head_beta == nullptr))
    return fail(h, -1, "head_gamma and head_beta must both be given or both be NULL");
  if (fwd && L->head_gamma && !L->out_logits)
    return fail(h, -1, "head requested but out_logits is NULL");
  if (L->uhat_mode != SRF_UHAT_FP32 && L->uhat_mode != SRF_UHAT_TF32 &&
      L->uhat_mode != SRF_UHAT_BF16 && L->uhat_mode != SRF_UHAT_FP32X3 && L->uhat_mode != SRF_UHAT_F16)
    return fail(h, -4, "unknown uhat_mode %d", L->uhat_mode);
  if (L->O > 128) return fail(h, -3, "O = %d output capsules > 128 is not supported", L->O);
  if (L->D > 32 || L->d > 32)
    return fail(h, -3, "capsule dims D=%d, d=%d > 32 are not supported", L->D, L->d);
  return 0;
}

static int get_packed(srf_handle* h, const srf_layer_desc* L, int I, int T, int OP,
                      cudaStream_t stream, const PackedWeights** out) {
  PackedWeights* hit = nullptr;
  for (auto& pw : h->packed)
    if (pw.W == L->W && pw.bias == L->bias) {
      hit = &pw;
      break;
    }
  const size_t nW = (size_t)I * T * T * OP, nB = (size_t)I * T * OP;
  const size_t bytes = (nW + nB) * sizeof(float);
  bool same = hit && hit->I == I && hit->O == L->O && hit->D == L->D && hit->d == L->d &&
              hit->T == T && hit->OP == OP;
  if (same && L->weights_version != 0 && hit->version == L->weights_version) {
    *out = hit;
    return 0;
  }
  if (!hit) {
    if (h->packed.size() >= 64) {  // bound the cache: drop everything (stream-ordered frees)
      for (auto& pw : h->packed)
        if (pw.Wp) cudaFreeAsync(pw.Wp, stream);
      h->packed.clear();
    }
    h->packed.emplace_back();
    hit = &h->packed.back();
  }
  if (hit->bytes < bytes) {
    if (hit->Wp) cudaFreeAsync(hit->Wp, stream);
    hit->Wp = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&hit->Wp, bytes, stream);
    if (e != cudaSuccess) {
      hit->bytes = 0;
      return cuda_fail(h, e, "packed-weight allocation");
    }
    hit->bytes = bytes;
  }
  hit->Bp = hit->Wp + nW;
  hit->W = L->W;
  hit->bias = L->bias;
  hit->I = I;
  hit->O = L->O;
  hit->D = L->D;
  hit->d = L->d;
  hit->T = T;
  hit->OP = OP;
  hit->version = L->weights_version;
  {
    KernelSpan span(h, 0, stream);
    srf::launch_pack_weights(L->W, L->bias, hit->Wp, hit->Bp, I, L->O, L->D, L->d, T, OP, stream);
  }
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "pack_weights launch");
  *out = hit;
  return 0;
}

// ---------------------------------------------------------------------------------------
// tensor-core u_hat (uhat_gemm.cu): packed MMA weights, TMA tensor map over emb, launch
// ---------------------------------------------------------------------------------------
namespace {
struct UhatGeom {
  int I, T, OPL, MT, KC, NB, NS, NBT, NST, Bpad, MTG, NG;
  bool bf16, x3;
  size_t bytes;
};
}  // namespace

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int uhat_geometry(srf_handle* h, const srf_layer_desc* L, UhatGeom* g) {
  // SRF_UHAT_F16 shapes the fused kernel does not take run here as TF32 (the same significand width)
  if (L->uhat_mode != SRF_UHAT_TF32 && L->uhat_mode != SRF_UHAT_BF16 &&
      L->uhat_mode != SRF_UHAT_FP32X3 && L->uhat_mode != SRF_UHAT_F16)
    return fail(h, -4, "tensor-core u_hat needs uhat_mode TF32, F16, BF16 or FP32X3");
  if (L->d % 4 != 0)
    return fail(h, -3, "tensor-core u_hat needs d %% 4 == 0 (got d=%d); use SRF_UHAT_FP32", L->d);
  if ((reinterpret_cast<uintptr_t>(L->emb) & 15) != 0)
    return fail(h, -3, "tensor-core u_hat needs a 16-byte aligned emb pointer");
  const int window = L->lpad + L->rpad + 1;
  g->I = window * L->H;
  g->T = L->D <= 8 ? 8 : (L->D <= 16 ? 16 : (L->D <= 20 ? 20 : 32));  // = routing kernel's T
  g->OPL = L->O <= 32 ? 1 : (L->O <= 64 ? 2 : 4);  // = the routing kernels' output capsules per lane
  g->MT = g->OPL * (g->T / 4);
  g->KC = 2 * ((L->d + 7) / 8);
  g->Bpad = (L->B + 1) & ~1;
  // utterances per 64-frame tile: the power of two that wastes the fewest tile slots on the batch
  // edge (B = 43: NB = 4 covers 44 slots, NB = 64 would cover 64), larger on ties
  int nb = 2, best_cover = 1 << 30;
  for (int c = 64; c >= 2; c >>= 1) {
    const int cover = (g->Bpad + c - 1) / c * c;
    if (cover < best_cover) {
      best_cover = cover;
      nb = c;
    }
  }
  g->NB = nb;
  g->NS = 64 / nb;
  g->NBT = (L->B + g->NB - 1) / g->NB;
  g->NST = (L->S + g->NS - 1) / g->NS;
  g->bf16 = L->uhat_mode == SRF_UHAT_BF16;
  g->x3 = L->uhat_mode == SRF_UHAT_FP32X3;
  g->bytes = (size_t)L->S * (g->Bpad / 2) * g->I * g->MT * 128 * 2 * (g->bf16 ? 2 : 4);
  // M tiles of W[i] resident at a time: all of them, or as many as fit (the x tiles are then
  // streamed once per group)
  g->MTG = g->MT;
  while (g->x3 && g->MTG > 1 &&
         srf::uhat_gemm_smem_bytes(g->MTG, g->KC, 1) > (size_t)h->max_smem)
    g->MTG = (g->MTG + 1) / 2;
  g->NG = (g->MT + g->MTG - 1) / g->MTG;
  if (srf::uhat_gemm_smem_bytes(g->MTG, g->KC, g->x3 ? 1 : 0) > (size_t)h->max_smem)
    return fail(h, -3, "u_hat GEMM tile does not fit in shared memory (O=%d, D=%d, d=%d)", L->O,
                L->D, L->d);
  return 0;
}

static int get_packed_mma(srf_handle* h, const srf_layer_desc* L, const UhatGeom& g,
                          cudaStream_t stream, const PackedWeights** out) {
  PackedWeights* hit = nullptr;
  for (auto& pw : h->packed_mma)
    if (pw.W == L->W && pw.bias == L->bias && pw.x3 == (g.x3 ? 1 : 0)) {
      hit = &pw;
      break;
    }
  const size_t nW = (size_t)g.I * g.MT * g.KC * 512 * (g.x3 ? 2 : 1), nB = (size_t)g.I * g.MT * 128;
  const size_t bytes = (nW + nB) * sizeof(float);
  const bool same = hit && hit->I == g.I && hit->O == L->O && hit->D == L->D && hit->d == L->d &&
                    hit->x3 == (g.x3 ? 1 : 0);
  if (same && L->weights_version != 0 && hit->version == L->weights_version) {
    *out = hit;
    return 0;
  }
  if (!hit) {
    if (h->packed_mma.size() >= 64) {
      for (auto& pw : h->packed_mma)
        if (pw.Wp) cudaFreeAsync(pw.Wp, stream);
      h->packed_mma.clear();
    }
    h->packed_mma.emplace_back();
    hit = &h->packed_mma.back();
  }
  if (hit->bytes < bytes) {
    if (hit->Wp) cudaFreeAsync(hit->Wp, stream);
    hit->Wp = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&hit->Wp, bytes, stream);
    if (e != cudaSuccess) {
      hit->bytes = 0;
      return cuda_fail(h, e, "packed MMA weight allocation");
    }
    hit->bytes = bytes;
  }
  hit->Bp = hit->Wp + nW;
  hit->W = L->W;
  hit->bias = L->bias;
  hit->I = g.I;
  hit->O = L->O;
  hit->D = L->D;
  hit->d = L->d;
  hit->version = L->weights_version;
  hit->x3 = g.x3 ? 1 : 0;
  {
    KernelSpan span(h, 0, stream);
    srf::launch_pack_weights_mma(L->W, L->bias, hit->Wp, hit->Bp, g.I, L->O, L->D, L->d, g.T, g.OPL,
                                 g.KC, g.x3 ? 1 : 0, stream);
  }
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "pack_weights_mma launch");
  *out = hit;
  return 0;
}

// runs the GEMM; on return h->ubuf holds u_hat of this layer in the streaming layout
static int compute_uhat(srf_handle* h, const srf_layer_desc* L, const UhatGeom& g,
                        cudaStream_t stream) {
  const PackedWeights* pw = nullptr;
  int rc = get_packed_mma(h, L, g, stream, &pw);
  if (rc) return rc;
  if (h->ubuf_bytes < g.bytes) {
    if (h->ubuf) cudaFreeAsync(h->ubuf, stream);
    h->ubuf = nullptr;
    h->ubuf_bytes = 0;
    cudaError_t e = cudaMallocAsync(&h->ubuf, g.bytes, stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "u_hat buffer allocation");
    h->ubuf_bytes = g.bytes;
  }
  if (!h->encode_tiled) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !fn || qres != cudaDriverEntryPointSuccess)
      return fail(h, e != cudaSuccess ? (int)e : 999, "cuTensorMapEncodeTiled is not available");
    h->encode_tiled = fn;
  }
  // emb[B,S,H,d] viewed as 5-D (l_in=4, b, s, h, l4=d/4); box (4, NB, NS, 1, KC) lands in
  // shared memory as [K chunk][s][b][4 floats] = the canonical K-major UMMA operand layout.
  CUtensorMap tmap;
  const cuuint64_t gdim[5] = {4, (cuuint64_t)L->B, (cuuint64_t)L->S, (cuuint64_t)L->H,
                              (cuuint64_t)(L->d / 4)};
  const cuuint64_t gstr[4] = {(cuuint64_t)L->S * L->H * L->d * 4, (cuuint64_t)L->H * L->d * 4,
                              (cuuint64_t)L->d * 4, 16};
  const cuuint32_t box[5] = {4, (cuuint32_t)g.NB, (cuuint32_t)g.NS, 1, (cuuint32_t)g.KC};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = ((EncodeTiledFn)h->encode_tiled)(
      &tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(L->emb), gdim, gstr, box, estr,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(h, 900 + (int)cr, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)cr);

  srf::UhatParams p;
  p.Wm = pw->Wp;
  p.Bm = pw->Bp;
  p.u = h->ubuf;
  p.I = g.I;
  p.MT = g.MT;
  p.KC = g.KC;
  p.B = L->B;
  p.S = L->S;
  p.H = L->H;
  p.lpad = L->lpad;
  p.NB = g.NB;
  p.NS = g.NS;
  p.NBT = g.NBT;
  p.NST = g.NST;
  p.Bpad = g.Bpad;
  p.store_bf16 = g.bf16 ? 1 : 0;
  p.x3 = g.x3 ? 1 : 0;
  p.MTG = g.MTG;
  p.NG = g.NG;
  p.items = (long long)g.I * g.NG * g.NBT * g.NST;
  cudaError_t e;
  {
    KernelSpan span(h, 1, stream);
    e = srf::launch_uhat_gemm(tmap, p, h->num_sms, stream);
  }
  if (e != cudaSuccess) return cuda_fail(h, e, "uhat_gemm launch");
  h->launches++;
  return 0;
}

static void l2_release(srf_handle* h);
extern "C" int srf_uhat_fwd(srf_handle* h, const srf_layer_desc* L, float* out_uhat, void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!L || !out_uhat) return fail(h, -1, "layer descriptor or output is NULL");
  if (L->B == 0 || L->S == 0) return 0;
  if (!L->emb || !L->W || !L->bias) return fail(h, -1, "emb, W and bias must be non-NULL");
  if (L->H <= 0 || L->d <= 0 || L->O <= 0 || L->D <= 0 || L->lpad < 0 || L->rpad < 0)
    return fail(h, -2, "bad shape");
  if (L->O > 128 || L->D > 32 || L->d > 32) return fail(h, -3, "O > 128 or capsule dim > 32");
  DeviceGuard guard(h->device);
  l2_release(h);
  cudaStream_t stream = (cudaStream_t)stream_;
  enter_stream(h, stream);
  UhatGeom g;
  int rc = uhat_geometry(h, L, &g);
  if (rc) return rc;
  rc = compute_uhat(h, L, g, stream);
  if (rc) return rc;
  srf::launch_unpack_uhat(h->ubuf, out_uhat, L->B, L->S, g.I, L->O, L->D, g.T, g.OPL, g.Bpad,
                          g.bf16 ? 1 : 0, stream);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "unpack_uhat launch");
  h->last_kernel = "uhat_gemm_kernel(tcgen05 tf32)";
  return 0;
}

// ---------------------------------------------------------------------------------------
// fused routing kernel (routing_fused.cu): eligibility, packed operand images, work plan, launch
// ---------------------------------------------------------------------------------------
namespace {
struct FusedGeom {
  int T4, opl, KC, KX, I;
};
}  // namespace

// a previous fused launch that timed out left its code in the mapped host word
static int fused_check_abort(srf_handle* h) {
  if (h->fz_host_abort && *h->fz_host_abort != 0) {
    const int code = *h->fz_host_abort;
    *h->fz_host_abort = 0;
    return fail(h, 700 + code, "a fused routing launch timed out waiting (wait site %d); its results are invalid", code);
  }
  return 0;
}

// fused_forward's answer for a shape it does not take: the caller runs the two-kernel path
enum { FUSED_FALLBACK = 1 << 20 };

static bool fused_geometry(const srf_handle* h, const srf_layer_desc* L, FusedGeom* g) {
  if (h->no_fused) return false;
  if (L->uhat_mode != SRF_UHAT_TF32 && L->uhat_mode != SRF_UHAT_FP32X3 && L->uhat_mode != SRF_UHAT_F16)
    return false;
  if (L->d % 4 != 0 || L->d > 32 || L->D > 20 || L->O > 64 || L->iters > 64) return false;
  if (L->emb && (reinterpret_cast<uintptr_t>(L->emb) & 15) != 0) return false;
  const int t4 = (L->D + 3) / 4;
  g->T4 = t4 <= 2 ? 2 : (t4 <= 4 ? 4 : 5);
  g->opl = (L->O + 31) / 32;
  if (!srf::route_fused_supported(g->T4, g->opl)) return false;
  g->KX = L->d / 4;
  // 16-byte K chunks per operand row, incl. the bias column, a whole number of MMA K steps (2 chunks):
  // 4 fp32 / tf32 elements per chunk, 8 fp16 elements
  g->KC = L->uhat_mode == SRF_UHAT_F16 ? 2 * ((L->d + 1 + 15) / 16) : 2 * ((L->d + 1 + 7) / 8);
  g->I = (L->lpad + L->rpad + 1) * L->H;
  return true;
}

static int get_packed_fused(srf_handle* h, const srf_layer_desc* L, const FusedGeom& g, int T4, int parts,
                            cudaStream_t stream, const PackedWeights** out) {
  PackedWeights* hit = nullptr;
  for (auto& pw : h->packed_fused)
    if (pw.W == L->W && pw.bias == L->bias && pw.x3 == parts && pw.T == T4) {
      hit = &pw;
      break;
    }
  const size_t nW = (size_t)g.I * (parts ? parts : 1) * g.opl * T4 * g.KC * 512;   // parts 0 = FP16 image
  const size_t bytes = nW * sizeof(float);
  const bool same = hit && hit->I == g.I && hit->O == L->O && hit->D == L->D && hit->d == L->d;
  if (same && L->weights_version != 0 && hit->version == L->weights_version) {
    *out = hit;
    return 0;
  }
  if (!hit) {
    if (h->packed_fused.size() >= 64) {
      for (auto& pw : h->packed_fused)
        if (pw.Wp) cudaFreeAsync(pw.Wp, stream);
      h->packed_fused.clear();
    }
    h->packed_fused.emplace_back();
    hit = &h->packed_fused.back();
  }
  if (hit->bytes < bytes) {
    if (hit->Wp) cudaFreeAsync(hit->Wp, stream);
    hit->Wp = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&hit->Wp, bytes, stream);
    if (e != cudaSuccess) {
      hit->bytes = 0;
      return cuda_fail(h, e, "packed fused weight allocation");
    }
    hit->bytes = bytes;
  }
  hit->W = L->W;
  hit->bias = L->bias;
  hit->I = g.I;
  hit->O = L->O;
  hit->D = L->D;
  hit->d = L->d;
  hit->T = T4;
  hit->version = L->weights_version;
  hit->x3 = parts;
  {
    KernelSpan span(h, 0, stream);
    srf::launch_pack_weights_fused(L->W, L->bias, hit->Wp, g.I, L->O, L->D, L->d, T4, g.opl, g.KC, parts,
                                   stream);
  }
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "pack_weights_fused launch");
  *out = hit;
  return 0;
}

// One fused launch: an SDR stack of n layers as a wavefront, or ONE DR layer (n == 1).
// `layers` are validated, all fused-eligible, and agree on B, S, sdr, iters and uhat_mode.
// Persisting-L2 carve-out for the packed weights of a fused multi-layer stack.  Without it the share of
// the weight stream that misses L2 swings from run to run with the placement of the CTAs on the two dies
// (ncu, cfg-3 f16: 10.4 and 21.3 GB of DRAM reads per step in two captures of the same build); with the
// arena marked persisting it is 2.7 GB at the same speed (13.9 ms).  The carve-out is device-wide, so it
// is reserved when the fused inference path runs and dropped again by every other path.
static bool l2_carve(srf_handle* h) {
  if (h->l2_persist_max == 0) return false;
  if (!h->l2_carved) {
    if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, h->l2_persist_max) != cudaSuccess) {
      cudaGetLastError();
      h->l2_persist_max = 0;
      return false;
    }
    h->l2_carved = true;
  }
  return true;
}
static void l2_release(srf_handle* h) {
  if (!h->l2_carved) return;
  cudaCtxResetPersistingL2Cache();
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
  cudaGetLastError();
  h->l2_carved = false;
}

static int fused_forward(srf_handle* h, const srf_layer_desc* layers, int n, cudaStream_t stream) {
  int rc = fused_check_abort(h);
  if (rc) return rc;
  const srf_layer_desc& L0 = layers[0];
  const int B = L0.B, S = L0.S, sdr = L0.sdr ? 1 : 0;
  // operand images per W tile: 1 = TF32, 2 = hi + lo (3 x TF32), 0 = one FP16 image; kmode = the
  // kernel's MODE template argument
  const int parts = L0.uhat_mode == SRF_UHAT_FP32X3 ? 2 : (L0.uhat_mode == SRF_UHAT_F16 ? 0 : 1);
  const int nimg = parts ? parts : 1;
  const int kmode = parts == 2 ? 1 : (parts == 0 ? 2 : 0);
  std::vector<FusedGeom> geo(n);
  int T4 = 0, OPLM = 0, KCmax = 0;
  for (int l = 0; l < n; ++l) {
    if (!fused_geometry(h, &layers[l], &geo[l])) return fail(h, -3, "layer %d is not fused-eligible", l);
    if (geo[l].T4 > T4) T4 = geo[l].T4;
    if (geo[l].opl > OPLM) OPLM = geo[l].opl;
    if (geo[l].KC > KCmax) KCmax = geo[l].KC;
  }
  if (!srf::route_fused_supported(T4, OPLM)) return fail(h, -3, "fused variant T4=%d OPL=%d not built", T4, OPLM);
  if (n > srf::FZ_MAX_LAYERS) return fail(h, -3, "fused launch holds at most %d layers", srf::FZ_MAX_LAYERS);
  const int T = 4 * T4, OP = 32 * OPLM;

  // ---- work plan ----
  int NB = 32, NS = 1, ngroups, grid, rounds, nslots, maxC = 1;
  std::vector<int> Cl(n, 1);
  int Gc = 1, NBT = 1;
  if (sdr) {
    ngroups = (B + 31) / 32;
    Gc = h->num_sms / n;
    if (Gc < 1) return fail(h, -3, "more layers than SMs");
    if (Gc > ngroups) Gc = ngroups;
    rounds = (ngroups + Gc - 1) / Gc;
    int budget = h->num_sms / Gc - n;  // CTAs to hand out on top of one per layer
    while (budget > 0) {
      int best = -1;
      double worst = 0.0;
      for (int l = 0; l < n; ++l) {
        if (Cl[l] >= geo[l].I) continue;
        const double load = (double)geo[l].I / Cl[l];
        if (load > worst) {
          worst = load;
          best = l;
        }
      }
      if (best < 0) break;
      Cl[best]++;
      budget--;
    }
    // TPC alignment.  Blocks 2k and 2k+1 of the cooperative launch are the two SMs of one TPC
    // (profiles/r2d_block_to_sm.txt).  Measured on cfg-3 (bench.py, ms per step; the layers 1-9 get seven CTAs
    // each except two that get six and carry the most capsules per CTA): the two 6-CTA units on blocks that start
    // at an even offset -- whole TPCs -- 13.28-13.33 (layers 2,3 / 4,5 / 6,7); starting at an odd offset, i.e.
    // sharing a TPC with a unit of another layer at both ends, 13.69-13.79 (layers 3,4 / 5,6 / 3,6); the greedy
    // order (last two layers) 13.69-13.74.  So, among layers of equal size (interchangeable for the greedy
    // count), even-sized units are dealt to even block offsets and odd-sized ones to odd offsets.
    if (!getenv("SRF_FUSED_NO_ALIGN")) {
      std::vector<int> done(n, 0);
      for (int l0 = 0; l0 < n; ++l0) {
        if (done[l0]) continue;
        std::vector<int> cls, cnt;
        for (int l = l0; l < n; ++l)
          if (!done[l] && geo[l].I == geo[l0].I) {
            cls.push_back(l);
            cnt.push_back(Cl[l]);
            done[l] = 1;
          }
        if (cls.size() < 2) continue;
        // offsets depend on everything before a layer: walk all layers, re-deal this class on the way
        int off = 0;
        size_t k = 0;
        for (int l = 0; l < n; ++l) {
          if (k < cls.size() && l == cls[k]) {
            int pick = -1;
            for (size_t c = 0; c < cnt.size(); ++c)
              if (cnt[c] > 0 && ((cnt[c] ^ off) & 1) == 0) {   // same parity as the offset: even on even, odd on odd
                pick = (int)c;
                break;
              }
            if (pick < 0)
              for (size_t c = 0; c < cnt.size(); ++c)
                if (cnt[c] > 0) {
                  pick = (int)c;
                  break;
                }
            Cl[l] = cnt[pick];
            cnt[pick] = 0;
            ++k;
          }
          off += Cl[l];
        }
      }
    }
    if (h->force_C > 0)
      for (int l = 0; l < n; ++l) Cl[l] = h->force_C < geo[l].I ? h->force_C : geo[l].I;
    if (const char* ev = getenv("SRF_FUSED_PLAN")) {
      // developer override: CTAs per layer, ':'-separated
      int l = 0;
      for (const char* q = ev; *q && l < n; ++l) {
        const int c = atoi(q);
        Cl[l] = c < 1 ? 1 : (c > geo[l].I ? geo[l].I : c);
        while (*q && *q != ':') ++q;
        if (*q == ':') ++q;
      }
    }
    int per_group = 0;
    for (int l = 0; l < n; ++l) {
      per_group += Cl[l];
      if (Cl[l] > maxC) maxC = Cl[l];
    }
    if (per_group * Gc > h->num_sms) return fail(h, -3, "fused plan exceeds the SM count");
    grid = per_group * Gc;
    nslots = Gc * n;
  } else {
    NB = 1;
    while (NB < 32 && NB < B) NB *= 2;
    NS = 32 / NB;
    NBT = (B + NB - 1) / NB;
    const int NST = (S + NS - 1) / NS;
    ngroups = NBT * NST;
    if (ngroups >= h->num_sms) {
      grid = h->num_sms;
      rounds = (ngroups + grid - 1) / grid;
      nslots = grid;
    } else {
      int C = h->num_sms / ngroups;
      if (C > geo[0].I) C = geo[0].I;
      if (h->force_C > 0 && h->force_C <= C) C = h->force_C;
      Cl[0] = maxC = C;
      grid = ngroups * C;
      rounds = 1;
      nslots = ngroups;
    }
  }
  // Policy (measured, profiles/r2_modes.md): a pass costs the fused kernel a fixed ~35k clk (the
  // L2 exchange of the partial sums) plus ~3k clk per input capsule of the CTA's slice (tf32;
  // 3 x TF32 issues three times the MMAs: ~9k).  With few capsules per CTA (TIMIT-sized SDR layers
  // spread over 148 SMs) or in the 3 x TF32 mode on long slices the two-kernel path is faster.
  {
    int ncap_min = 1 << 30;
    for (int l = 0; l < n; ++l) {
      const int c = geo[l].I / Cl[sdr ? l : 0];
      if (c < ncap_min) ncap_min = c;
    }
    const bool x3 = parts == 2;
    bool take = x3 ? (sdr && ncap_min >= 8 && ncap_min <= 16) : !(sdr && ncap_min < 8);
    if (h->force_fused > 0) take = true;
    if (!take) return FUSED_FALLBACK;
  }
  std::vector<srf::FusedItem> items((size_t)rounds * grid);
  for (auto& itm : items) itm.layer = -1;
  auto vmask_of = [&](int b0, int s0) {
    unsigned m = 0;
    for (int f = 0; f < 32; ++f)
      if (b0 + f % NB < B && s0 + f / NB < S) m |= 1u << f;
    return m;
  };
  if (sdr) {
    for (int r = 0; r < rounds; ++r) {
      int cta = 0;
      for (int gs = 0; gs < Gc; ++gs)
        for (int l = 0; l < n; ++l)
          for (int c = 0; c < Cl[l]; ++c, ++cta) {
            const int group = r * Gc + gs;
            if (group >= ngroups) continue;
            srf::FusedItem& itm = items[(size_t)r * grid + cta];
            itm.layer = l;
            itm.group = group;
            itm.slot = gs * n + l;
            itm.c = c;
            itm.C = Cl[l];
            itm.i_lo = (int)((long long)geo[l].I * c / Cl[l]);
            itm.i_hi = (int)((long long)geo[l].I * (c + 1) / Cl[l]);
            itm.b0 = group * 32;
            itm.s0 = 0;
            itm.vmask = vmask_of(itm.b0, 0);
          }
    }
  } else {
    const int C = Cl[0];
    for (int r = 0; r < rounds; ++r)
      for (int cta = 0; cta < grid; ++cta) {
        const int group = C > 1 ? cta / C : r * grid + cta;
        if (group >= ngroups) continue;
        srf::FusedItem& itm = items[(size_t)r * grid + cta];
        itm.layer = 0;
        itm.group = group;
        itm.slot = C > 1 ? group : cta;
        itm.c = C > 1 ? cta % C : 0;
        itm.C = C;
        itm.i_lo = (int)((long long)geo[0].I * itm.c / C);
        itm.i_hi = (int)((long long)geo[0].I * (itm.c + 1) / C);
        itm.b0 = (group % NBT) * NB;
        itm.s0 = (group / NBT) * NS;
        itm.vmask = vmask_of(itm.b0, itm.s0);
      }
  }

  // ---- packed weights of a multi-layer stack: one arena, repacked when any layer's tag changed ----
  std::vector<size_t> arena_off(n, 0);
  size_t arena_total = 0;
  if (n > 1) {
    bool same = (int)h->fz_arena_keys.size() == n;
    for (int l = 0; l < n; ++l) {
      arena_off[l] = arena_total;
      arena_total += (((size_t)geo[l].I * nimg * geo[l].opl * T4 * geo[l].KC * 512 * sizeof(float)) + 255) & ~(size_t)255;
      if (same) {
        const PackedWeights& k = h->fz_arena_keys[l];
        const srf_layer_desc& L = layers[l];
        same = k.W == L.W && k.bias == L.bias && k.version == L.weights_version && L.weights_version != 0 &&
               k.T == T4 && k.x3 == parts && k.I == geo[l].I && k.O == L.O && k.D == L.D && k.d == L.d;
      }
    }
    if (!same) {
      if (arena_total > h->fz_arena_bytes) {
        if (h->fz_arena) cudaFreeAsync(h->fz_arena, stream);
        h->fz_arena = nullptr;
        h->fz_arena_bytes = 0;
        cudaError_t ea = cudaMallocAsync((void**)&h->fz_arena, arena_total, stream);
        if (ea != cudaSuccess) return cuda_fail(h, ea, "packed-weight arena allocation");
        h->fz_arena_bytes = arena_total;
      }
      h->fz_arena_keys.assign(n, PackedWeights());
      for (int l = 0; l < n; ++l) {
        const srf_layer_desc& L = layers[l];
        {
          KernelSpan span(h, 0, stream);
          srf::launch_pack_weights_fused(L.W, L.bias, h->fz_arena + arena_off[l] / sizeof(float), geo[l].I, L.O, L.D,
                                         L.d, T4, geo[l].opl, geo[l].KC, parts, stream);
        }
        h->launches++;
        PackedWeights& k = h->fz_arena_keys[l];
        k.W = L.W;
        k.bias = L.bias;
        k.version = L.weights_version;
        k.T = T4;
        k.x3 = parts;
        k.I = geo[l].I;
        k.O = L.O;
        k.D = L.D;
        k.d = L.d;
      }
      cudaError_t ep = cudaGetLastError();
      if (ep != cudaSuccess) return cuda_fail(h, ep, "pack_weights_fused launch");
    }
  }

  // ---- inter-layer buffers, packed weights, device descriptors ----
  std::vector<srf::FusedLayer> fl(n);
  memset(fl.data(), 0, sizeof(srf::FusedLayer) * n);
  const float* prev_out = nullptr;
  for (int l = 0; l < n; ++l) {
    const srf_layer_desc& L = layers[l];
    const bool is_final = l == n - 1;
    float* out_caps = L.out_caps;
    if (!out_caps && !(is_final && L.head_gamma)) {
      const size_t bytes = (size_t)B * S * L.O * L.D * sizeof(float);
      if ((int)h->fz_inter.size() < n) h->fz_inter.resize(n, nullptr);
      if (bytes > h->fz_inter_bytes) {
        for (float*& b : h->fz_inter) {
          if (b) cudaFreeAsync(b, stream);
          b = nullptr;
        }
        h->fz_inter_bytes = bytes;
      }
      if (!h->fz_inter[l]) {
        cudaError_t e = cudaMallocAsync((void**)&h->fz_inter[l], h->fz_inter_bytes, stream);
        if (e != cudaSuccess) return cuda_fail(h, e, "inter-layer buffer allocation");
      }
      out_caps = h->fz_inter[l];
    }
    const float* emb = L.emb ? L.emb : prev_out;
    if (!emb) return fail(h, -1, "layer %d: emb is NULL", l);
    if ((reinterpret_cast<uintptr_t>(emb) & 15) != 0) return fail(h, -3, "layer %d: emb is not 16-byte aligned", l);
    const float* Wf = nullptr;
    if (n > 1) {
      Wf = h->fz_arena + arena_off[l] / sizeof(float);
    } else {
      const PackedWeights* pw = nullptr;
      rc = get_packed_fused(h, &L, geo[l], T4, parts, stream, &pw);
      if (rc) return rc;
      Wf = pw->Wp;
    }
    srf::FusedLayer& F = fl[l];
    F.emb = emb;
    F.Wf = Wf;
    F.ln_gamma = L.ln_gamma;
    F.ln_beta = L.ln_beta;
    F.dropout_mask = L.dropout_mask;
    F.head_gamma = L.head_gamma;
    F.head_beta = L.head_beta;
    F.out_caps = out_caps;
    F.out_logits = L.out_logits;
    F.out_raw = L.out_raw;
    F.H = L.H;
    F.O = L.O;
    F.D = L.D;
    F.opl = geo[l].opl;
    F.KC = geo[l].KC;
    F.KX = geo[l].KX;
    F.lpad = L.lpad;
    F.rpad = L.rpad;
    F.mask0 = L.mask_class0 ? 1 : 0;
    F.dep_layer = (l > 0 && emb == prev_out) ? l - 1 : -1;
    F.ln_eps = L.ln_eps;
    F.length_eps = L.length_eps;
    prev_out = out_caps;
  }

  // device table: layers | items | cnt_p | cnt_v | progress | abort
  const size_t off_items = sizeof(srf::FusedLayer) * n;
  const size_t off_cnt = (off_items + sizeof(srf::FusedItem) * items.size() + 63) & ~(size_t)63;
  const size_t n_prog = sdr ? (size_t)n * ngroups * 32 : 0;
  const size_t cnt_ints = (size_t)nslots * (3 + 32) + n_prog + 16;
  const size_t tab_bytes = off_cnt + cnt_ints * sizeof(int);
  if (tab_bytes > h->fz_tab_bytes) {
    if (h->fz_tab) cudaFreeAsync(h->fz_tab, stream);
    h->fz_tab = nullptr;
    h->fz_tab_bytes = 0;
    cudaError_t e = cudaMallocAsync(&h->fz_tab, tab_bytes, stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "fused table allocation");
    h->fz_tab_bytes = tab_bytes;
  }
  const size_t x_floats = (size_t)nslots * (maxC + 1) * 32 * T * OP;
  if (x_floats * sizeof(float) > h->fz_x_bytes) {
    if (h->fz_x) cudaFreeAsync(h->fz_x, stream);
    h->fz_x = nullptr;
    h->fz_x_bytes = 0;
    cudaError_t e = cudaMallocAsync((void**)&h->fz_x, x_floats * sizeof(float), stream);
    if (e != cudaSuccess) return cuda_fail(h, e, "fused exchange buffer allocation");
    h->fz_x_bytes = x_floats * sizeof(float);
  }
  uint8_t* tab = reinterpret_cast<uint8_t*>(h->fz_tab);
  // pageable source: the runtime stages the bytes before returning, the vectors may die after this
  cudaError_t e = cudaMemcpyAsync(tab, fl.data(), sizeof(srf::FusedLayer) * n, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(tab + off_items, items.data(), sizeof(srf::FusedItem) * items.size(),
                        cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(tab + off_cnt, 0, cnt_ints * sizeof(int), stream);
  if (e != cudaSuccess) return cuda_fail(h, e, "fused table upload");

  srf::FusedParams p;
  p.layers = reinterpret_cast<const srf::FusedLayer*>(tab);
  p.items = reinterpret_cast<const srf::FusedItem*>(tab + off_items);
  p.rounds = rounds;
  int* ints = reinterpret_cast<int*>(tab + off_cnt);
  p.cnt_p = ints;
  p.cnt_v = ints + nslots;
  p.oflag = ints + 3 * (size_t)nslots;
  p.progress = sdr && n > 1 ? ints + 35 * (size_t)nslots : nullptr;
  p.abort_flag = ints + 35 * (size_t)nslots + n_prog;
  p.host_abort = h->fz_host_abort_dev;
  p.xP = h->fz_x;
  p.xV = h->fz_x + (size_t)nslots * maxC * 32 * T * OP;
  p.maxC = maxC;
  p.ngroups = ngroups;
  p.B = B;
  p.S = S;
  p.sdr = sdr;
  p.iters = L0.iters;
  p.NB = NB;
  // W ring: a stage is G consecutive tiles fetched by ONE bulk copy (the copy rate is set by the
  // bytes issued per barrier round trip: single 12 KB copies reach 19 B/clk/SM, 48 KB copies 70,
  // profiles/r2_ubench.txt); at least 3 stages, G as large as fits up to ~48 KB
  const size_t wpair = (size_t)KCmax * 2048 * nimg;
  int G = (int)(49152 / wpair);
  if (G < 1) G = 1;
  int nwst = 0;
  for (; G >= 1; --G) {
    nwst = 8;
    while (nwst > 3 && srf::route_fused_smem_bytes(OPLM, KCmax, kmode, nwst, G * wpair) > (size_t)h->max_smem) --nwst;
    if (srf::route_fused_smem_bytes(OPLM, KCmax, kmode, nwst, G * wpair) <= (size_t)h->max_smem) break;
  }
  // Capsule-sized stages: when three stages of one capsule's tiles (NT x wpair) fit, the kernel runs ONE MMA
  // issuer and a dedicated W producer, and a stage is released by the capsule's own commit (no second
  // issuer's book-keeping, no commit per stage).  SRF_FUSED_CAPSTAGE=0 keeps the tile-granular ring.
  int capstage = 0;
  {
    const char* ev = getenv("SRF_FUSED_CAPSTAGE");
    const size_t cb = (size_t)T4 * OPLM * wpair;
    // (small capsules keep the tile-granular ring: its stages bundle several capsules into ~48 KB copies,
    // and the L2 -> shared-memory rate follows the bytes per barrier round trip; cfg-2: 16 KB per capsule)
    const bool big = cb >= 32768 || (ev && ev[0] == '1');
    if (big && !(ev && ev[0] == '0') && srf::route_fused_smem_bytes(OPLM, KCmax, kmode, 3, cb) <= (size_t)h->max_smem) {
      capstage = 1;
      G = T4 * OPLM;
      nwst = 3;
      while (nwst < 6 && srf::route_fused_smem_bytes(OPLM, KCmax, kmode, nwst + 1, cb) <= (size_t)h->max_smem) ++nwst;
    }
  }
  // the MMA issuers address a capsule's tiles through at most three ring stages
  if (G < 1 || (!capstage && T4 * OPLM >= 2 * G + 2)) return FUSED_FALLBACK;   // the two-kernel path takes this shape
  const size_t smem = srf::route_fused_smem_bytes(OPLM, KCmax, kmode, nwst, G * wpair);
  p.capstage = capstage;
  p.gtiles = G;
  p.wstage_bytes = (int)(G * wpair);
  p.xtile_bytes = KCmax * 32 * 16;
  p.xstg_bytes = 2 * KCmax * 32 * 16;
  p.nwst = nwst;
  p.dbg = h->dbg;
  p.n_layers = n;
  p.per_group = 0;
  for (int l = 0; l < srf::FZ_MAX_LAYERS; ++l) {
    if (l < n) {
      p.per_group += sdr ? Cl[l] : Cl[0];
      p.layer_I[l] = geo[l].I;
      p.layer_opl[l] = geo[l].opl;
      p.layer_KC[l] = geo[l].KC;
      p.layer_KX[l] = geo[l].KX;
    } else {
      p.layer_I[l] = p.layer_opl[l] = p.layer_KC[l] = p.layer_KX[l] = 0;
    }
    p.cta_end[l] = p.per_group;
  }
  {
    KernelSpan span(h, 2, stream);
    // inference launches of a multi-layer stack whose packed weights fit the persisting carve-out (the
    // FP16 images of cfg-3: 67.6 MB) mark the arena as persisting in L2 (see l2_carve); training forwards
    // (out_raw requested) leave the L2 to the backward that follows.  (TF32 images, 101 MB, were measured
    // in the first session: 39.0 -> 31.9 GB of DRAM reads but 19.1 -> 20.0 ms: not worth it.)
    bool infer = n > 1 && arena_total <= h->l2_persist_max && arena_total <= h->l2_window_max;
    for (int l = 0; l < n && infer; ++l)
      if (layers[l].out_raw) infer = false;
    const void* win = nullptr;
    if (infer && l2_carve(h)) win = h->fz_arena;
    else l2_release(h);
    size_t win_bytes = win ? arena_total : 0;
    const float hit = win_bytes > h->l2_persist_max ? (float)h->l2_persist_max / (float)win_bytes : 1.0f;
    e = srf::launch_route_fused(p, T4, OPLM, kmode, grid, smem, stream, win, win_bytes, hit);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(h, e, "route_fused launch");
  }
  h->launches++;
  char nm[240];
  snprintf(nm, sizeof(nm),
           "route_fused_kernel<T4=%d,OPL=%d,%s> %s layers=%d grid=%d groups=%d rounds=%d maxC=%d wstages=%dx%d "
           "tiles smem=%zu",
           T4, OPLM, parts == 2 ? "3xTF32" : (parts == 0 ? "f16" : "tf32"), sdr ? "SDR-wavefront" : "DR", n, grid, ngroups, rounds,
           maxC, nwst, G, smem);
  if (capstage) strncat(nm, " capstage", sizeof(nm) - strlen(nm) - 1);
  h->last_kernel = nm;
  return 0;
}

static int pow2_floor(int v) {
  int p = 1;
  while (p * 2 <= v) p *= 2;
  return p;
}

static int route_layer_impl(srf_handle* h, const srf_layer_desc* L, cudaStream_t stream) {
  if (L && (L->B == 0 || L->S == 0)) return 0;  // empty batch: nothing to do
  int rc = validate_layer(h, L);
  if (rc) return rc;
  if (!L->out_caps && !L->out_logits && !L->out_raw) return fail(h, -1, "no output requested");

  {
    FusedGeom fg;
    if (fused_geometry(h, L, &fg)) {
      const int frc = fused_forward(h, L, 1, stream);
      if (frc != FUSED_FALLBACK) return frc;
    }
  }
  l2_release(h);   // the two-kernel paths want the whole L2 for the materialised u_hat
  const int window = L->lpad + L->rpad + 1;
  const int I = window * L->H;
  const int um = L->uhat_mode == SRF_UHAT_FP32 ? 0 : (L->uhat_mode == SRF_UHAT_BF16 ? 1 : 2);
  const int OPL = L->O <= 32 ? 1 : (L->O <= 64 ? 2 : 4);
  const int OP = 32 * OPL;
  int T;
  const PackedWeights* pw = nullptr;
  UhatGeom g;
  if (um == 0) {
    const int m = L->D > L->d ? L->D : L->d;
    T = m <= 8 ? 8 : (m <= 16 ? 16 : (m <= 20 ? 20 : 32));
    rc = get_packed(h, L, I, T, OP, stream, &pw);
    if (rc) return rc;
  } else {
    // tensor-core path: u_hat of the whole layer by the tcgen05 GEMM, then stream it
    rc = uhat_geometry(h, L, &g);
    if (rc) return rc;
    T = g.T;
    if ((T == 16 && OPL > 2) || (T == 20 && OPL > 2) || (T == 32 && OPL > 1))
      return fail(h, -3, "tensor-core path: O=%d with D=%d is not instantiated", L->O, L->D);
    rc = compute_uhat(h, L, g, stream);
    if (rc) return rc;
  }

  const int halfB = (L->B + 1) / 2;
  long long nchains = L->sdr ? L->B : (long long)L->B * L->S;
  if (um != 0 && !L->sdr) nchains = 2LL * halfB * L->S;
  if (nchains > (1LL << 30)) return fail(h, -2, "too many frames");
  int F = srf::route_layer_max_F(T, OPL);
  while (F > 1 && ((nchains + F - 1) / F) * 8 < h->num_sms) F /= 2;
  if (h->force_F > 0 && h->force_F <= srf::route_layer_max_F(T, OPL)) F = h->force_F;
  if (um != 0) F = 2;
  int groups = (int)((nchains + F - 1) / F);
  int C = pow2_floor(h->num_sms / groups > 0 ? h->num_sms / groups : 1);
  if (C > 8) C = 8;
  if (h->force_C > 0) C = h->force_C;
  if (C > I) C = pow2_floor(I);
  // shared-memory fit: widen the cluster, then narrow the chain group
  for (;;) {
    const int Ic = (I + C - 1) / C;
    const size_t smem = srf::route_layer_smem_bytes(T, OPL, F, SRF_NW, Ic, um);
    if (smem <= (size_t)h->max_smem) break;
    if (C < 8 && C * 2 <= I) {
      C *= 2;
      continue;
    }
    if (F > 1 && um == 0) {
      F /= 2;
      groups = (int)((nchains + F - 1) / F);
      continue;
    }
    return fail(h, -3, "layer does not fit in shared memory (I=%d, T=%d, O=%d)", I, T, L->O);
  }
  const int Ic = (I + C - 1) / C;
  const size_t smem = srf::route_layer_smem_bytes(T, OPL, F, SRF_NW, Ic, um);

  srf::RouteParams p;
  p.emb = L->emb;
  p.Wp = pw ? pw->Wp : nullptr;
  p.Bp = pw ? pw->Bp : nullptr;
  p.ln_gamma = L->ln_gamma;
  p.ln_beta = L->ln_beta;
  p.dropout_mask = L->dropout_mask;
  p.head_gamma = L->head_gamma;
  p.head_beta = L->head_beta;
  p.out_caps = L->out_caps;
  p.out_logits = L->out_logits;
  p.out_raw = L->out_raw;
  p.B = L->B;
  p.S = L->S;
  p.H = L->H;
  p.d = L->d;
  p.O = L->O;
  p.D = L->D;
  p.I = I;
  p.lpad = L->lpad;
  p.iters = L->iters;
  p.sdr = L->sdr ? 1 : 0;
  p.mask0 = L->mask_class0 ? 1 : 0;
  p.C = C;
  p.Ic = Ic;
  p.nchains = (int)nchains;
  p.nsteps = L->sdr ? L->S : 1;
  p.ln_eps = L->ln_eps;
  p.length_eps = L->length_eps;
  p.u = um != 0 ? h->ubuf : nullptr;
  p.halfB = halfB;

  p.nstage = 0;
  p.dbg = h->dbg;
  if (um != 0 && !h->no_stream) {
    // streaming kernel: TMA-fed ring + warp-specialised output; needs >= 2 ring stages
    int Cs = C;
    if (Cs * groups > h->num_sms) Cs = pow2_floor(h->num_sms / groups > 0 ? h->num_sms / groups : 1);
    const size_t stage = srf::route_stream_stage_bytes(T, OPL, um == 1);
    const size_t fixed = srf::route_stream_fixed_smem(T, OPL, um == 1, Cs, 16);
    if (fixed + 2 * stage <= (size_t)h->max_smem) {
      int nstage = (int)(((size_t)h->max_smem - fixed) / stage);
      if (nstage > 16) nstage = 16;
      if (h->max_stages >= 2 && nstage > h->max_stages) nstage = h->max_stages;
      p.C = Cs;
      p.Ic = (I + Cs - 1) / Cs;
      p.nstage = nstage;
      const size_t smem_s = fixed + (size_t)nstage * stage;
      cudaError_t es;
      {
        KernelSpan span(h, 2, stream);
        es = srf::launch_route_stream(p, T, OPL, um == 1, groups, smem_s, stream);
      }
      if (es != cudaSuccess) {
        cudaGetLastError();
        return cuda_fail(h, es, "route_stream launch");
      }
      h->launches++;
      char nm[200];
      snprintf(nm, sizeof(nm),
               "uhat_gemm_kernel(tcgen05 tf32) + route_stream_kernel<T=%d,OPL=%d,NW=%d,%s> C=%d "
               "groups=%d stages=%d smem=%zu",
               T, OPL, srf::route_stream_nslot(T, OPL, um == 1), um == 1 ? "bf16" : "fp32", Cs, groups,
               nstage, smem_s);
      h->last_kernel = nm;
      return 0;
    }
  }

  cudaError_t e;
  {
    KernelSpan span(h, 2, stream);
    e = srf::launch_route_layer(p, T, OPL, F, groups, smem, um, stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(h, e, "route_layer launch");
  }
  h->launches++;
  char name[160];
  snprintf(name, sizeof(name),
           "%sroute_layer_kernel<T=%d,OPL=%d,F=%d,NW=%d,UM=%d> C=%d groups=%d smem=%zu",
           um != 0 ? "uhat_gemm_kernel(tcgen05 tf32) + " : "", T, OPL, F, SRF_NW, um, C, groups, smem);
  h->last_kernel = name;
  return 0;
}

extern "C" int srf_route_layer_bwd(srf_handle* h, const srf_layer_desc* L, const srf_layer_grads* G,
                                   void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!L || !G) return fail(h, -1, "layer descriptor or grads is NULL");
  if (L->B == 0 || L->S == 0) return 0;
  DeviceGuard guard(h->device);
  enter_stream(h, (cudaStream_t)stream_);
  l2_release(h);   // the backward's kernels want the whole L2
  cudaStream_t stream = (cudaStream_t)stream_;
  int rc = validate_layer(h, L, false);
  if (rc) return rc;
  if (!G->v_raw || !G->d_raw || !G->dW || !G->dbias)
    return fail(h, -1, "v_raw, d_raw, dW and dbias must be non-NULL");
  if (!G->d_out && !G->d_logits) return fail(h, -1, "neither d_out nor d_logits given");
  if (G->d_logits && (!L->head_gamma || !G->dhead_gamma || !G->dhead_beta))
    return fail(h, -1, "d_logits needs head_gamma and dhead_gamma/dhead_beta");
  if (L->ln_gamma && (!G->dgamma || !G->dbeta)) return fail(h, -1, "LayerNorm needs dgamma/dbeta");
  if (L->iters > 8) return fail(h, -3, "backward supports iters <= 8");
  const int window = L->lpad + L->rpad + 1;
  const int I = window * L->H;
  const int um = L->uhat_mode == SRF_UHAT_FP32 ? 0 : (L->uhat_mode == SRF_UHAT_BF16 ? 1 : 2);
  const int OPL = L->O <= 32 ? 1 : (L->O <= 64 ? 2 : 4);
  if (um != 0 && h->bwd_atomics)
    return fail(h, -4, "SRF_BWD_ATOMICS=1 supports uhat_mode FP32 only");
  int T;
  const PackedWeights* pw = nullptr;
  if (um == 0) {
    // exact mode: the BPTT sweep recomputes u_hat in FP32 from the packed weights
    const int m = L->D > L->d ? L->D : L->d;
    T = m <= 8 ? 8 : (m <= 16 ? 16 : (m <= 20 ? 20 : 32));
    rc = get_packed(h, L, I, T, 32 * OPL, stream, &pw);
    if (rc) return rc;
  } else {
    // tensor path: u_hat of the whole layer by the tcgen05 GEMM, streamed by the BPTT sweep
    UhatGeom g;
    rc = uhat_geometry(h, L, &g);
    if (rc) return rc;
    T = g.T;
    rc = compute_uhat(h, L, g, stream);
    if (rc) return rc;
  }
  if ((T >= 16 && OPL > 2) || (T == 32 && OPL > 1))
    return fail(h, -3, "backward: O=%d with D=%d is not instantiated", L->O, L->D);
  srf::BwdParams p;
  p.emb = L->emb;
  p.Wp = pw ? pw->Wp : nullptr;
  p.Bp = pw ? pw->Bp : nullptr;
  p.W = L->W;
  p.u = um != 0 ? h->ubuf : nullptr;
  p.halfB = (L->B + 1) / 2;
  p.ln_gamma = L->ln_gamma;
  p.ln_beta = L->ln_beta;
  p.dropout_mask = L->dropout_mask;
  p.head_gamma = L->head_gamma;
  p.v_raw = G->v_raw;
  p.d_out = G->d_out;
  p.d_logits = G->d_logits;
  p.d_raw = G->d_raw;
  p.dW = G->dW;
  p.dbias = G->dbias;
  p.dgamma = G->dgamma;
  p.dbeta = G->dbeta;
  p.dhead_gamma = G->dhead_gamma;
  p.dhead_beta = G->dhead_beta;
  p.d_emb = G->d_emb;
  p.B = L->B;
  p.S = L->S;
  p.H = L->H;
  p.d = L->d;
  p.O = L->O;
  p.D = L->D;
  p.I = I;
  p.lpad = L->lpad;
  p.iters = L->iters;
  p.sdr = L->sdr ? 1 : 0;
  p.mask0 = L->mask_class0 ? 1 : 0;
  p.nsteps = L->sdr ? L->S : 1;
  p.ln_eps = L->ln_eps;
  p.length_eps = L->length_eps;
  // split mode (default): the BPTT sweep saves c / g_a / g_t / Vacc, the frame-parallel phase B
  // builds dW, dbias and dx without atomics.  SRF_BWD_ATOMICS=1 selects the fused atomics variant.
  p.split = h->bwd_atomics ? 0 : 1;
  p.OP = 32 * OPL;
  p.Tu = T;
  p.dp = (L->d + 3) & ~3;
  p.FS = 1;
  p.fps = 0;
  // the reduced-precision modes differentiate a TF32-class function: their phase B runs on the tensor
  // cores; FP32 and FP32X3 (the 1e-4 class) keep the FP32 CUDA-core contraction
  p.tc_phase_b = (L->uhat_mode == SRF_UHAT_TF32 || L->uhat_mode == SRF_UHAT_BF16 || L->uhat_mode == SRF_UHAT_F16) &&
                 !getenv("SRF_BWD_NO_TC");
  p.cbuf = p.gabuf = p.gtT = p.vaT = p.dxw = p.dwp = nullptr;
  if (p.split) {
    p.FS = srf::dwdx_frame_splits(p, h->max_smem, h->num_sms);
    if (p.FS <= 0) return fail(h, -3, "backward phase B does not fit in shared memory (D=%d, d=%d)", L->D, L->d);
    const size_t frames = (size_t)L->B * L->S, R = (size_t)L->iters;
    p.fps = (int)((frames + p.FS - 1) / p.FS);
    const size_t n_c = frames * R * I * p.OP, n_g = frames * R * L->O * T;
    const size_t n_x = frames * I * OPL * p.dp;
    const size_t n_w = (size_t)p.FS * I * L->O * L->D * (L->d + 1);
    const size_t need = (2 * n_c + 2 * n_g + n_x + n_w) * sizeof(float);
    if (need > h->bwd_ws_bytes) {
      if (h->bwd_ws) cudaFreeAsync(h->bwd_ws, stream);
      h->bwd_ws = nullptr;
      h->bwd_ws_bytes = 0;
      cudaError_t ea = cudaMallocAsync((void**)&h->bwd_ws, need, stream);
      if (ea != cudaSuccess) return cuda_fail(h, ea, "backward workspace allocation");
      h->bwd_ws_bytes = need;
    }
    p.cbuf = h->bwd_ws;
    p.gabuf = p.cbuf + n_c;
    p.gtT = p.gabuf + n_c;
    p.vaT = p.gtT + n_g;
    p.dxw = p.vaT + n_g;
    p.dwp = p.dxw + n_x;
  }
  {
    KernelSpan span(h, 2, stream);
    srf::launch_ln_head_bwd(p, stream);
  }
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "ln_head_bwd launch");
  const long long nchains = L->sdr ? L->B : (long long)L->B * L->S;
  {
    // cluster split over the input capsules when there are fewer chains than SMs (SDR)
    const int nw = srf::route_layer_bwd_warps(um, T, OPL);
    int C = pow2_floor(h->num_sms / nchains > 0 ? (int)(h->num_sms / nchains) : 1);
    if (C > 8) C = 8;
    while (C > 1 && (I + C - 1) / C < nw / 2) C /= 2;
    if (h->force_C > 0 && h->force_C <= 8) C = h->force_C;
    if (!p.split) C = 1;
    // the exchange buffers grow with the cluster size: narrow the cluster until the CTA fits
    while (C > 1 && srf::route_layer_bwd_smem_bytes(T, OPL, um, (I + C - 1) / C, C) > (size_t)h->max_smem)
      C /= 2;
    if (srf::route_layer_bwd_smem_bytes(T, OPL, um, (I + C - 1) / C, C) > (size_t)h->max_smem)
      return fail(h, -3, "backward sweep does not fit in shared memory (I=%d, O=%d, D=%d)", I, L->O, L->D);
    p.C = C;
    p.Ic = (I + C - 1) / C;
    p.dbg = h->dbg;
  }
  {
    KernelSpan span(h, 2, stream);
    e = srf::launch_route_layer_bwd(p, T, OPL, um, (int)nchains, stream);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(h, e, "route_layer_bwd launch");
  }
  h->launches++;
  if (p.split) {
    {
      KernelSpan span(h, 2, stream);
      e = srf::launch_dwdx_from_saved(p, h->max_smem, stream);
    }
    if (e != cudaSuccess) {
      cudaGetLastError();
      return cuda_fail(h, e, "dwdx_from_saved launch");
    }
    h->launches += 2;
    if (p.d_emb) {
      KernelSpan span(h, 2, stream);
      srf::launch_fold_dx(p, stream);
      h->launches++;
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(h, e, "fold_dx launch");
  }
  h->last_kernel = "ln_head_bwd_kernel + route_layer_bwd_kernel";
  return 0;
}

extern "C" int srf_route_layer_fwd(srf_handle* h, const srf_layer_desc* layer, void* stream) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  DeviceGuard g(h->device);
  enter_stream(h, (cudaStream_t)stream);
  return route_layer_impl(h, layer, (cudaStream_t)stream);
}

extern "C" int srf_route_stack_fwd(srf_handle* h, const srf_layer_desc* layers, int32_t n_layers,
                                   void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!layers || n_layers <= 0) return fail(h, -1, "no layers");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
  enter_stream(h, stream);
  const int B = layers[0].B, S = layers[0].S;
  size_t need = 0;
  for (int n = 0; n < n_layers; ++n) {
    const srf_layer_desc& L = layers[n];
    if (L.B != B || L.S != S) return fail(h, -2, "layer %d: B,S differ from layer 0", n);
    if (n > 0 && (L.H != layers[n - 1].O || L.d != layers[n - 1].D))
      return fail(h, -2, "layer %d: input capsules (%d x %d) do not match layer %d output (%d x %d)",
                  n, L.H, L.d, n - 1, layers[n - 1].O, layers[n - 1].D);
    if (n == 0 && !L.emb) return fail(h, -1, "layer 0: emb is NULL");
    const bool is_final = n == n_layers - 1;
    const bool needs_ws = !L.out_caps && !(is_final && L.head_gamma);
    if (needs_ws) {
      const size_t bytes = (size_t)B * S * L.O * L.D * sizeof(float);
      if (bytes > need) need = bytes;
    }
  }
  // SDR stacks whose layers all qualify run as ONE wavefront launch of the fused kernel
  if (layers[0].sdr && n_layers >= 1 && n_layers <= srf::FZ_MAX_LAYERS) {
    bool all = true;
    for (int n = 0; n < n_layers && all; ++n) {
      FusedGeom fg;
      const srf_layer_desc& L = layers[n];
      all = validate_layer_quiet(&L, n == 0) && fused_geometry(h, &L, &fg) && L.sdr &&
            L.iters == layers[0].iters && L.uhat_mode == layers[0].uhat_mode &&
            (L.out_caps || L.out_logits || L.out_raw || n < n_layers - 1);
    }
    if (all) {
      for (int n = 0; n < n_layers; ++n) {
        srf_layer_desc L = layers[n];
        if (n > 0 && !L.emb) L.emb = layers[0].emb;  // placeholder for the validator only
        int rc = validate_layer(h, &L);
        if (rc) return rc;
      }
      if (B == 0 || S == 0) return 0;
      const int frc = fused_forward(h, layers, n_layers, stream);
      if (frc != FUSED_FALLBACK) return frc;
    }
  }
  if (need > h->ws_bytes) {
    for (int i = 0; i < 2; ++i) {
      if (h->ws[i]) cudaFreeAsync(h->ws[i], stream);
      h->ws[i] = nullptr;
    }
    h->ws_bytes = 0;
    for (int i = 0; i < 2; ++i) {
      cudaError_t e = cudaMallocAsync((void**)&h->ws[i], need, stream);
      if (e != cudaSuccess) return cuda_fail(h, e, "stack workspace allocation");
    }
    h->ws_bytes = need;
  }
  const float* prev = nullptr;
  for (int n = 0; n < n_layers; ++n) {
    srf_layer_desc L = layers[n];
    if (n > 0 && !L.emb) L.emb = prev;
    const bool is_final = n == n_layers - 1;
    if (!L.out_caps && !(is_final && L.head_gamma)) L.out_caps = h->ws[n & 1];
    int rc = route_layer_impl(h, &L, stream);
    if (rc) return rc;
    prev = L.out_caps;
  }
  return 0;
}

// the whole routing stack, backward: srf_route_layer_bwd from the last layer down, the gradient
// w.r.t. a layer's input feeding the next call (reference: tape.gradient through the stack,
// tfsr/trainer_sr.py:62-71)
extern "C" int srf_route_stack_bwd(srf_handle* h, const srf_layer_desc* layers, const srf_layer_grads* grads,
                                   int32_t n_layers, void* stream) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!layers || !grads || n_layers <= 0) return fail(h, -1, "layers / grads is NULL or n_layers <= 0");
  for (int n = n_layers - 1; n >= 0; --n) {
    srf_layer_desc L = layers[n];
    srf_layer_grads G = grads[n];
    if (!L.emb) {
      if (n == 0) return fail(h, -1, "layers[0].emb is NULL");
      L.emb = layers[n - 1].out_caps;
      if (!L.emb) return fail(h, -1, "layer %d: the input (layers[%d].out_caps of the forward) is NULL", n, n - 1);
    }
    if (n < n_layers - 1) {
      if (!G.d_out) G.d_out = grads[n + 1].d_emb;
      if (!G.d_out) return fail(h, -1, "layer %d: grads[%d].d_emb is needed to continue the backward", n, n + 1);
      G.d_logits = nullptr;
    }
    if (n > 0 && !G.d_emb) return fail(h, -1, "layer %d: d_emb is required for every layer but the first", n);
    const int rc = srf_route_layer_bwd(h, &L, &G, stream);
    if (rc) return rc;
  }
  return 0;
}

