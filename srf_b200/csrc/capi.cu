// capi.cu -- the C-ABI of include/srf_b200.h: handle, argument validation, packed-weight
// cache, kernel-variant dispatch and the per-layer stack driver.  No torch types here.

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
  int force_F = 0, force_C = 0;
};

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
  *out = h;
  return 0;
}

extern "C" int srf_destroy(srf_handle* h) {
  if (!h) return 0;
  DeviceGuard g(h->device);
  for (auto& pw : h->packed) {
    if (pw.Wp) cudaFree(pw.Wp);
  }
  for (int i = 0; i < 2; ++i)
    if (h->ws[i]) cudaFree(h->ws[i]);
  delete h;
  return 0;
}

extern "C" const char* srf_last_error(const srf_handle* h) {
  return h ? h->error.c_str() : g_create_error.c_str();
}

extern "C" int64_t srf_launch_count(const srf_handle* h) { return h ? h->launches : 0; }

extern "C" const char* srf_last_kernel(const srf_handle* h) {
  return h ? h->last_kernel.c_str() : "";
}

static int validate_layer(srf_handle* h, const srf_layer_desc* L) {
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
  if (L->head_gamma && !L->out_logits) return fail(h, -1, "head requested but out_logits is NULL");
  if (L->uhat_mode != SRF_UHAT_FP32)
    return fail(h, -4, "uhat_mode %d is not available in this build (only SRF_UHAT_FP32)",
                L->uhat_mode);
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
  srf::launch_pack_weights(L->W, L->bias, hit->Wp, hit->Bp, I, L->O, L->D, L->d, T, OP, stream);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(h, e, "pack_weights launch");
  *out = hit;
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
  if (!L->out_caps && !L->out_logits) return fail(h, -1, "no output requested");

  const int window = L->lpad + L->rpad + 1;
  const int I = window * L->H;
  const int m = L->D > L->d ? L->D : L->d;
  const int T = m <= 8 ? 8 : (m <= 16 ? 16 : (m <= 20 ? 20 : 32));
  const int OPL = L->O <= 32 ? 1 : (L->O <= 64 ? 2 : 4);
  const int OP = 32 * OPL;

  const PackedWeights* pw = nullptr;
  rc = get_packed(h, L, I, T, OP, stream, &pw);
  if (rc) return rc;

  const long long nchains = L->sdr ? L->B : (long long)L->B * L->S;
  if (nchains > (1LL << 30)) return fail(h, -2, "too many frames");
  int F = srf::route_layer_max_F(T, OPL);
  while (F > 1 && ((nchains + F - 1) / F) * 8 < h->num_sms) F /= 2;
  if (h->force_F > 0 && h->force_F <= srf::route_layer_max_F(T, OPL)) F = h->force_F;
  int groups = (int)((nchains + F - 1) / F);
  int C = pow2_floor(h->num_sms / groups > 0 ? h->num_sms / groups : 1);
  if (C > 8) C = 8;
  if (h->force_C > 0) C = h->force_C;
  while (C > 1 && (I + C - 1) / C < 1) C /= 2;
  if (C > I) C = pow2_floor(I);
  // shared-memory fit: widen the cluster, then narrow the chain group
  for (;;) {
    const int Ic = (I + C - 1) / C;
    const size_t smem = srf::route_layer_smem_bytes(T, OPL, F, SRF_NW, Ic);
    if (smem <= (size_t)h->max_smem) break;
    if (C < 8 && C * 2 <= I) {
      C *= 2;
      continue;
    }
    if (F > 1) {
      F /= 2;
      groups = (int)((nchains + F - 1) / F);
      continue;
    }
    return fail(h, -3, "layer does not fit in shared memory (I=%d, T=%d, O=%d)", I, T, L->O);
  }
  const int Ic = (I + C - 1) / C;
  const size_t smem = srf::route_layer_smem_bytes(T, OPL, F, SRF_NW, Ic);

  srf::RouteParams p;
  p.emb = L->emb;
  p.Wp = pw->Wp;
  p.Bp = pw->Bp;
  p.ln_gamma = L->ln_gamma;
  p.ln_beta = L->ln_beta;
  p.dropout_mask = L->dropout_mask;
  p.head_gamma = L->head_gamma;
  p.head_beta = L->head_beta;
  p.out_caps = L->out_caps;
  p.out_logits = L->out_logits;
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

  cudaError_t e = srf::launch_route_layer(p, T, OPL, F, groups, smem, stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return cuda_fail(h, e, "route_layer launch");
  }
  h->launches++;
  char name[128];
  snprintf(name, sizeof(name), "route_layer_kernel<T=%d,OPL=%d,F=%d,NW=%d> C=%d groups=%d smem=%zu",
           T, OPL, F, SRF_NW, C, groups, smem);
  h->last_kernel = name;
  return 0;
}

extern "C" int srf_route_layer_fwd(srf_handle* h, const srf_layer_desc* layer, void* stream) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  DeviceGuard g(h->device);
  return route_layer_impl(h, layer, (cudaStream_t)stream);
}

extern "C" int srf_route_stack_fwd(srf_handle* h, const srf_layer_desc* layers, int32_t n_layers,
                                   void* stream_) {
  if (!h) return fail(nullptr, -1, "handle is NULL");
  if (!layers || n_layers <= 0) return fail(h, -1, "no layers");
  DeviceGuard g(h->device);
  cudaStream_t stream = (cudaStream_t)stream_;
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
