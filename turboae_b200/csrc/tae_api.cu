// C ABI of libturboae_b200.so (see include/turboae_b200.h): argument validation, the error
// convention and dispatch to the fp32 (tae_f32.cu) and bf16 tcgen05 (tae_dec_pair.cu) paths.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "tae_common.cuh"

namespace tae {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- device-mapped host slots for the barrier-wait codes (one per device) --------------------------------------
static std::mutex g_slot_mu;
static int* g_slot_host[TAE_MAX_DEVICES] = {};
static int* g_slot_dev[TAE_MAX_DEVICES] = {};

int* wait_code_slot(void* fallback) {
  int* fb = reinterpret_cast<int*>(align_up(reinterpret_cast<uintptr_t>(fallback), 16));
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TAE_MAX_DEVICES) return fb;
  std::lock_guard<std::mutex> lk(g_slot_mu);
  if (!g_slot_dev[dev]) {
    void* h = nullptr;
    void* d = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped) != cudaSuccess) { (void)cudaGetLastError(); return fb; }
    memset(h, 0, 64);
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { (void)cudaGetLastError(); cudaFreeHost(h); return fb; }
    g_slot_host[dev] = reinterpret_cast<int*>(h);
    g_slot_dev[dev] = reinterpret_cast<int*>(d);
  }
  return g_slot_dev[dev];
}

static int pending_wait_code() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= TAE_MAX_DEVICES) return 0;
  std::lock_guard<std::mutex> lk(g_slot_mu);
  return g_slot_host[dev] ? *reinterpret_cast<volatile int*>(g_slot_host[dev]) : 0;
}

int require_sm100(int dev, const char* who) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { set_error("%s: cudaGetDeviceProperties: %s", who, cudaGetErrorString(e)); return TAE_ECUDA; }
  if (prop.major != 10) { set_error("%s needs an sm_100a device (found sm_%d%d)", who, prop.major, prop.minor); return TAE_EUNSUPPORTED; }
  return TAE_OK;
}

int after_launch(const char* kernel_name) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    const int code = pending_wait_code();
    if (code) set_error("%s launch failed: %s (an earlier kernel trapped in a bounded barrier wait, code %d: see tae_umma.cuh)",
                        kernel_name, cudaGetErrorString(e), code);
    else set_error("%s launch failed: %s", kernel_name, cudaGetErrorString(e));
    return TAE_ECUDA;
  }
  return TAE_OK;
}

static bool bf16_supported(const TaeDecConfig& c, const char** why) { return dec_pair_supported(c, why); }

int check_dec_config(const TaeDecConfig* c) {
  if (!c) { set_error("TaeDecConfig is NULL"); return TAE_EINVAL; }
  if (c->block_len < 1 || c->num_iteration < 1 || c->num_iter_ft < 1 || c->num_layer < 1 || c->num_unit < 1) {
    set_error("TaeDecConfig: non-positive dimension (L=%d I=%d F=%d layers=%d units=%d)", c->block_len,
              c->num_iteration, c->num_iter_ft, c->num_layer, c->num_unit);
    return TAE_EINVAL;
  }
  if (c->num_iteration > 32 || c->num_layer > 16) {
    set_error("TaeDecConfig: num_iteration <= 32 and num_layer <= 16 supported (got %d, %d)", c->num_iteration,
              c->num_layer);
    return TAE_EUNSUPPORTED;
  }
  if (c->kernel_size < 1 || c->kernel_size > 9 || (c->kernel_size & 1) == 0) {
    set_error("TaeDecConfig: kernel_size %d unsupported (odd sizes 1..9)", c->kernel_size);
    return TAE_EUNSUPPORTED;
  }
  return TAE_OK;
}

int check_enc_config(const TaeEncConfig* c) {
  if (!c) { set_error("TaeEncConfig is NULL"); return TAE_EINVAL; }
  if (c->block_len < 1 || c->num_layer < 1 || c->num_unit < 1) {
    set_error("TaeEncConfig: non-positive dimension (L=%d layers=%d units=%d)", c->block_len, c->num_layer,
              c->num_unit);
    return TAE_EINVAL;
  }
  if (c->num_layer > 16) { set_error("TaeEncConfig: num_layer <= 16 supported (got %d)", c->num_layer); return TAE_EUNSUPPORTED; }
  if (c->kernel_size < 1 || c->kernel_size > 9 || (c->kernel_size & 1) == 0) {
    set_error("TaeEncConfig: kernel_size %d unsupported (odd sizes 1..9)", c->kernel_size);
    return TAE_EUNSUPPORTED;
  }
  return TAE_OK;
}

}  // namespace tae

using namespace tae;

#define TAE_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      set_error(__VA_ARGS__);             \
      return TAE_EINVAL;                  \
    }                                     \
  } while (0)

extern "C" {

int tae_version(void) { return 100; }   // 0.1.0

const char* tae_last_error(void) { return g_err; }

uint64_t tae_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int tae_interleave_f32(const float* in, float* out, const int32_t* perm, int32_t B, int32_t L, int32_t F,
                       void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && F >= 1, "tae_interleave_f32: bad shape B=%d L=%d F=%d", B, L, F);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(in && out && perm, "tae_interleave_f32: NULL pointer");
  TAE_REQUIRE(in != out, "tae_interleave_f32: in-place permutation is not supported");
  return launch_interleave_f32(in, out, perm, B, L, F, (cudaStream_t)stream);
}

size_t tae_conv1d_workspace_bytes(int32_t Cin, int32_t Cout, int32_t K) {
  if (Cin < 1 || Cout < 1 || K < 1) return 0;
  return conv_packed_floats(Cin, Cout, K) * sizeof(float) + 256;
}

int tae_conv1d_elu_f32(const float* in, float* out, const float* weight, const float* bias, int32_t B, int32_t L,
                       int32_t Cin, int32_t Cout, int32_t K, int32_t apply_elu, void* workspace,
                       size_t workspace_bytes, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && Cin >= 1 && Cout >= 1, "tae_conv1d_elu_f32: bad shape B=%d L=%d Cin=%d Cout=%d", B,
              L, Cin, Cout);
  if (K < 1 || K > 9 || (K & 1) == 0) {
    set_error("tae_conv1d_elu_f32: kernel_size %d unsupported (odd sizes 1..9)", K);
    return TAE_EUNSUPPORTED;
  }
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(in && out && weight && bias && workspace, "tae_conv1d_elu_f32: NULL pointer");
  if (workspace_bytes < tae_conv1d_workspace_bytes(Cin, Cout, K)) {
    set_error("tae_conv1d_elu_f32: workspace %zu < %zu bytes", workspace_bytes, tae_conv1d_workspace_bytes(Cin, Cout, K));
    return TAE_EWORKSPACE;
  }
  float* packed = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  int rc = launch_pack_conv_f32(weight, packed, Cin, Cout, K, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_conv_f32(in, out, packed, bias, B, L, Cin, Cout, K, apply_elu, (cudaStream_t)stream);
}

size_t tae_conv1d_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t K) {
  if (Cin < 1 || Cout < 1 || K < 1) return 0;
  return conv_bwd_packed_floats(Cin, Cout, K) * sizeof(float) + 256;
}

int tae_conv1d_elu_bwd_f32(const float* x, const float* y, const float* dy, const float* weight, float* dx, float* dweight,
                           float* dbias, int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t K, int32_t apply_elu,
                           void* workspace, size_t workspace_bytes, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && Cin >= 1 && Cout >= 1, "tae_conv1d_elu_bwd_f32: bad shape B=%d L=%d Cin=%d Cout=%d", B, L, Cin, Cout);
  if (K < 1 || K > 9 || (K & 1) == 0) { set_error("tae_conv1d_elu_bwd_f32: kernel_size %d unsupported (odd sizes 1..9)", K); return TAE_EUNSUPPORTED; }
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(x && y && dy && weight && workspace, "tae_conv1d_elu_bwd_f32: NULL pointer");
  TAE_REQUIRE(dx || dweight, "tae_conv1d_elu_bwd_f32: nothing to compute (dx and dweight are NULL)");
  if (workspace_bytes < tae_conv1d_bwd_workspace_bytes(Cin, Cout, K)) {
    set_error("tae_conv1d_elu_bwd_f32: workspace %zu < %zu bytes", workspace_bytes, tae_conv1d_bwd_workspace_bytes(Cin, Cout, K));
    return TAE_EWORKSPACE;
  }
  float* packed = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  return launch_conv_bwd_f32(x, y, dy, weight, dx, dweight, dbias, B, L, Cin, Cout, K, apply_elu, packed, (cudaStream_t)stream);
}

size_t tae_dec_param_count(const TaeDecConfig* cfg) {
  if (check_dec_config(cfg)) return 0;
  return dec_layout(*cfg, nullptr);
}

size_t tae_dec_packed_bytes(const TaeDecConfig* cfg) {
  if (check_dec_config(cfg)) return 0;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return 0; }
  return dec_pair_packed_bytes(*cfg);
}

int tae_dec_pack_bf16(const TaeDecConfig* cfg, const float* params, void* packed, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed, "tae_dec_pack_bf16: NULL pointer");
  return dec_pair_pack(*cfg, params, packed, (cudaStream_t)stream);
}

size_t tae_dec_workspace_bytes(const TaeDecConfig* cfg, int32_t B, int32_t precision) {
  if (check_dec_config(cfg) || B < 0) return 0;
  if (precision == TAE_PRECISION_FP32) return dec_workspace_bytes_f32(*cfg, B);
  if (precision == TAE_PRECISION_BF16) {
    const char* why = nullptr;
    if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return 0; }
    return 256;
  }
  if (precision == TAE_PRECISION_F16X3) {
    const char* why = nullptr;
    if (!dec_x3_supported(*cfg, &why)) { set_error("f16x3 path: %s", why); return 0; }
    return 256;
  }
  set_error("unknown precision %d", precision);
  return 0;
}

int tae_dec_forward(const TaeDecConfig* cfg, const float* params, const void* packed, const float* received,
                    const int32_t* perm, const int32_t* inv_perm, float* out, float* trace, int32_t B,
                    int32_t precision, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  TAE_REQUIRE(B >= 0, "tae_dec_forward: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(params && received && perm && inv_perm && out && workspace, "tae_dec_forward: NULL pointer");
  if (precision == TAE_PRECISION_FP32)
    return dec_forward_f32(*cfg, params, received, perm, inv_perm, out, trace, B, workspace, workspace_bytes,
                           (cudaStream_t)stream);
  if (precision == TAE_PRECISION_BF16) {
    const char* why = nullptr;
    if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
    TAE_REQUIRE(packed, "tae_dec_forward(bf16): packed weight image is NULL (call tae_dec_pack_bf16)");
    return dec_forward_pair(*cfg, packed, received, perm, inv_perm, out, trace, B, workspace, workspace_bytes,
                            (cudaStream_t)stream);
  }
  if (precision == TAE_PRECISION_F16X3) {
    const char* why = nullptr;
    if (!dec_x3_supported(*cfg, &why)) { set_error("f16x3 path: %s", why); return TAE_EUNSUPPORTED; }
    TAE_REQUIRE(packed, "tae_dec_forward(f16x3): packed weight image is NULL (call tae_dec_pack_f16x3)");
    return dec_forward_x3(*cfg, params, packed, received, perm, inv_perm, out, trace, B, workspace, workspace_bytes,
                          (cudaStream_t)stream);
  }
  set_error("tae_dec_forward: unknown precision %d", precision);
  return TAE_EINVAL;
}

size_t tae_dec_packed_bytes_x3(const TaeDecConfig* cfg) {
  if (check_dec_config(cfg)) return 0;
  const char* why = nullptr;
  if (!dec_x3_supported(*cfg, &why)) { set_error("f16x3 path: %s", why); return 0; }
  return dec_x3_packed_bytes(*cfg);
}

int tae_dec_pack_f16x3(const TaeDecConfig* cfg, const float* params, void* packed, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!dec_x3_supported(*cfg, &why)) { set_error("f16x3 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed, "tae_dec_pack_f16x3: NULL pointer");
  return dec_x3_pack(*cfg, params, packed, (cudaStream_t)stream);
}

// ---- host-buffer decode: chunked H2D / decode / D2H pipeline ---------------------------------------------------
static constexpr int HOST_CHUNKS = 4;

static int host_chunk(int B) {
  if (B < 4000) return B;                                   // too small to be worth splitting
  const int c = (B + HOST_CHUNKS - 1) / HOST_CHUNKS;
  return (c + 9) / 10 * 10;                                 // whole CTA-pair groups (10 codewords at L = 100)
}

size_t tae_dec_host_workspace_bytes(const TaeDecConfig* cfg, int32_t B, int32_t precision) {
  if (check_dec_config(cfg) || B < 0) return 0;
  const size_t chunk = (size_t)host_chunk(B);
  const size_t kernel_ws = tae_dec_workspace_bytes(cfg, (int32_t)chunk, precision);
  if (B > 0 && kernel_ws == 0) return 0;
  return 2 * align_up(chunk * cfg->block_len * 3 * sizeof(float), 256) + 2 * align_up(chunk * cfg->block_len * sizeof(float), 256) +
         align_up(kernel_ws, 256) + 256;
}

int tae_dec_forward_host(const TaeDecConfig* cfg, const float* params, const void* packed, const float* received_host,
                         const int32_t* perm, const int32_t* inv_perm, float* out_host, int32_t B, int32_t precision,
                         void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  TAE_REQUIRE(B >= 0, "tae_dec_forward_host: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(params && received_host && out_host && perm && inv_perm && workspace, "tae_dec_forward_host: NULL pointer");
  if (workspace_bytes < tae_dec_host_workspace_bytes(cfg, B, precision)) {
    set_error("tae_dec_forward_host: workspace %zu < %zu bytes", workspace_bytes, tae_dec_host_workspace_bytes(cfg, B, precision));
    return TAE_EWORKSPACE;
  }
  // per-device copy streams and events, created once
  struct Pipe { cudaStream_t in = nullptr, out = nullptr; cudaEvent_t h2d[2], dec[2], d2h[2], start; bool ok = false; };
  static Pipe pipes[TAE_MAX_DEVICES];
  static std::mutex pipes_mu;
  int dev = 0;
  cudaGetDevice(&dev);
  TAE_REQUIRE(dev >= 0 && dev < TAE_MAX_DEVICES, "tae_dec_forward_host: device index %d out of range", dev);
  Pipe& P = pipes[dev];
  std::unique_lock<std::mutex> pipes_lk(pipes_mu);      // lazy creation is serialised; the streams are then used unlocked
  if (!P.ok) {
    cudaError_t e = cudaStreamCreateWithFlags(&P.in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P.out, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
      e = cudaEventCreateWithFlags(&P.h2d[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.dec[i], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.d2h[i], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.start, cudaEventDisableTiming);
    if (e != cudaSuccess) { set_error("tae_dec_forward_host: stream/event setup: %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    P.ok = true;
  }
  pipes_lk.unlock();
  cudaStream_t s = (cudaStream_t)stream;
  const int L = cfg->block_len;
  const size_t chunk = (size_t)host_chunk(B);
  const size_t in_b = align_up(chunk * L * 3 * sizeof(float), 256), out_b = align_up(chunk * L * sizeof(float), 256);
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<uintptr_t>(workspace), 256));
  float* d_in[2] = {reinterpret_cast<float*>(base), reinterpret_cast<float*>(base + in_b)};
  float* d_out[2] = {reinterpret_cast<float*>(base + 2 * in_b), reinterpret_cast<float*>(base + 2 * in_b + out_b)};
  void* kws = base + 2 * in_b + 2 * out_b;
  const size_t kws_b = tae_dec_workspace_bytes(cfg, (int32_t)chunk, precision);
#define TAE_CU(call)                                                                        \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess) { set_error("tae_dec_forward_host: %s: %s", #call, cudaGetErrorString(e_)); return TAE_ECUDA; } \
  } while (0)
  // the copy streams start after whatever precedes this call on `stream` (e.g. a weight update / pack)
  TAE_CU(cudaEventRecord(P.start, s));
  TAE_CU(cudaStreamWaitEvent(P.in, P.start, 0));
  TAE_CU(cudaStreamWaitEvent(P.out, P.start, 0));
  int i = 0;
  for (size_t b0 = 0; b0 < (size_t)B; b0 += chunk, ++i) {
    const int nb = (int)std::min(chunk, (size_t)B - b0);
    const int k = i & 1;
    if (i >= 2) TAE_CU(cudaStreamWaitEvent(P.in, P.dec[k], 0));          // staging buffer k: its previous decode is done
    TAE_CU(cudaMemcpyAsync(d_in[k], received_host + b0 * L * 3, (size_t)nb * L * 3 * sizeof(float), cudaMemcpyHostToDevice, P.in));
    TAE_CU(cudaEventRecord(P.h2d[k], P.in));
    TAE_CU(cudaStreamWaitEvent(s, P.h2d[k], 0));
    if (i >= 2) TAE_CU(cudaStreamWaitEvent(s, P.d2h[k], 0));             // output buffer k has been read back
    rc = tae_dec_forward(cfg, params, packed, d_in[k], perm, inv_perm, d_out[k], nullptr, nb, precision, kws, kws_b, s);
    if (rc) return rc;
    TAE_CU(cudaEventRecord(P.dec[k], s));
    TAE_CU(cudaStreamWaitEvent(P.out, P.dec[k], 0));
    TAE_CU(cudaMemcpyAsync(out_host + b0 * L, d_out[k], (size_t)nb * L * sizeof(float), cudaMemcpyDeviceToHost, P.out));
    TAE_CU(cudaEventRecord(P.d2h[k], P.out));
  }
  TAE_CU(cudaStreamWaitEvent(s, P.d2h[(i - 1) & 1], 0));                 // `stream` completes when out_host is complete
  if (i >= 2) TAE_CU(cudaStreamWaitEvent(s, P.d2h[i & 1], 0));
#undef TAE_CU
  return TAE_OK;
}

size_t tae_enc_param_count(const TaeEncConfig* cfg) {
  if (check_enc_config(cfg)) return 0;
  return enc_layout(*cfg, nullptr);
}

size_t tae_enc_workspace_bytes(const TaeEncConfig* cfg, int32_t B) {
  if (check_enc_config(cfg) || B < 0) return 0;
  return enc_workspace_bytes_f32(*cfg, B);
}

int tae_enc_forward(const TaeEncConfig* cfg, const float* params, const float* u, const int32_t* perm, float* x_tx,
                    double* stats, int32_t B, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  TAE_REQUIRE(B >= 0, "tae_enc_forward: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(params && u && perm && x_tx && stats && workspace, "tae_enc_forward: NULL pointer");
  return enc_forward_f32(*cfg, params, u, perm, x_tx, stats, B, workspace, workspace_bytes, (cudaStream_t)stream);
}

size_t tae_enc_packed_bytes(const TaeEncConfig* cfg) {
  if (check_enc_config(cfg)) return 0;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return 0; }
  return enc_pair_packed_bytes(*cfg);
}

int tae_enc_pack_bf16(const TaeEncConfig* cfg, const float* params, void* packed, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed, "tae_enc_pack_bf16: NULL pointer");
  return enc_pair_pack(*cfg, params, packed, (cudaStream_t)stream);
}

int tae_enc_forward_bf16(const TaeEncConfig* cfg, const void* packed, const float* u, const int32_t* perm,
                         const int32_t* inv_perm, float* x_tx, double* stats, int32_t B, void* workspace,
                         size_t workspace_bytes, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_enc_forward_bf16: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed && u && perm && inv_perm && x_tx && stats && workspace, "tae_enc_forward_bf16: NULL pointer");
  rc = enc_forward_pair(*cfg, packed, u, perm, inv_perm, x_tx, stats, B, workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_add_count(stats, (double)B * cfg->block_len * 3, (cudaStream_t)stream);
}

size_t tae_enc_packed_bytes_x3(const TaeEncConfig* cfg) {
  if (check_enc_config(cfg)) return 0;
  const char* why = nullptr;
  if (!enc_x3_supported(*cfg, &why)) { set_error("f16x3 encoder path: %s", why); return 0; }
  return enc_x3_packed_bytes(*cfg);
}

int tae_enc_pack_f16x3(const TaeEncConfig* cfg, const float* params, void* packed, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_x3_supported(*cfg, &why)) { set_error("f16x3 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed, "tae_enc_pack_f16x3: NULL pointer");
  return enc_x3_pack(*cfg, params, packed, (cudaStream_t)stream);
}

int tae_enc_forward_f16x3(const TaeEncConfig* cfg, const float* params, const void* packed, const float* u, const int32_t* perm,
                           const int32_t* inv_perm, float* x_tx, double* stats, int32_t B, void* workspace,
                           size_t workspace_bytes, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_x3_supported(*cfg, &why)) { set_error("f16x3 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_enc_forward_f16x3: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(params && packed && u && perm && inv_perm && x_tx && stats && workspace, "tae_enc_forward_f16x3: NULL pointer");
  rc = enc_forward_x3(*cfg, params, packed, u, perm, inv_perm, x_tx, stats, B, workspace, workspace_bytes, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_add_count(stats, (double)B * cfg->block_len * 3, (cudaStream_t)stream);
}

int tae_power_norm_f32(const float* x, float* codes, size_t n, const double* stats, float* mean_std, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(x && codes && stats, "tae_power_norm_f32: NULL pointer");
  return launch_power_norm_f32(x, codes, n, stats, mean_std, 1.f, 0.f, (cudaStream_t)stream);
}

int tae_power_norm_given_f32(const float* x, float* codes, size_t n, const float* mean_std, float value_limit, float quantize_level,
                             void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(x && codes && mean_std, "tae_power_norm_given_f32: NULL pointer");
  TAE_REQUIRE(quantize_level == 0.f || (value_limit > 0.f && quantize_level >= 2.f),
              "tae_power_norm_given_f32: quantize_level must be 0 (no quantiser) or >= 2 with value_limit > 0");
  return launch_power_norm_f32(x, codes, n, nullptr, const_cast<float*>(mean_std), quantize_level == 0.f ? 1.f : value_limit,
                               quantize_level, (cudaStream_t)stream);
}

int tae_power_norm_ste_f32(const float* x, float* codes, size_t n, const double* stats, float* mean_std, float value_limit,
                           float quantize_level, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(x && codes && stats, "tae_power_norm_ste_f32: NULL pointer");
  TAE_REQUIRE(value_limit > 0.f && quantize_level >= 2.f, "tae_power_norm_ste_f32: need value_limit > 0 and quantize_level >= 2");
  return launch_power_norm_f32(x, codes, n, stats, mean_std, value_limit, quantize_level, (cudaStream_t)stream);
}

int tae_dec_out_backward_f32(const float* d_out, const float* out, const int32_t* perm, float* d_out_last, int32_t B, int32_t L, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1, "tae_dec_out_backward_f32: bad shape B=%d L=%d", B, L);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(d_out && out && perm && d_out_last, "tae_dec_out_backward_f32: NULL pointer");
  return launch_dec_out_bwd_f32(d_out, out, perm, d_out_last, B, L, (cudaStream_t)stream);
}

int tae_dec_input_grad_f32(const float* dxin_all, const int32_t* inv_perm, float* d_received, int32_t n_stacks, int32_t B, int32_t L,
                           void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && n_stacks >= 1, "tae_dec_input_grad_f32: bad shape stacks=%d B=%d L=%d", n_stacks, B, L);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(dxin_all && inv_perm && d_received, "tae_dec_input_grad_f32: NULL pointer");
  return launch_dec_input_grad_f32(dxin_all, inv_perm, d_received, n_stacks, B, L, (cudaStream_t)stream);
}

int tae_enc_out_backward_f32(const float* d_x_tx, const float* x_tx, float* dlin, int32_t B, int32_t L, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1, "tae_enc_out_backward_f32: bad shape B=%d L=%d", B, L);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(d_x_tx && x_tx && dlin, "tae_enc_out_backward_f32: NULL pointer");
  return launch_enc_out_bwd_f32(d_x_tx, x_tx, dlin, B, L, (cudaStream_t)stream);
}

int tae_power_stats_f32(const float* x, size_t n, double* stats, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(x && stats, "tae_power_stats_f32: NULL pointer");
  int rc = launch_power_sums_f32(x, nullptr, n, stats, (cudaStream_t)stream);
  if (rc) return rc;
  return launch_add_count(stats, (double)n, (cudaStream_t)stream);
}

int tae_power_norm_bwd_sums_f32(const float* g, const float* codes, size_t n, double* sums, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(g && codes && sums, "tae_power_norm_bwd_sums_f32: NULL pointer");
  return launch_power_sums_f32(g, codes, n, sums, (cudaStream_t)stream);
}

int tae_power_norm_bwd_f32(const float* g, const float* codes, float* dx, size_t n, const double* sums, const double* stats,
                           const float* mean_std, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(g && codes && dx && sums && stats && mean_std, "tae_power_norm_bwd_f32: NULL pointer");
  return launch_power_norm_bwd_f32(g, codes, dx, n, sums, stats, mean_std, (cudaStream_t)stream);
}

int32_t tae_train_groups(int32_t block_len, int32_t B) {
  if (block_len < 1 || block_len > 512 || B < 0) return 0;
  return train_groups(block_len, B);
}

int32_t tae_train_units(int32_t block_len, int32_t B) { return (tae_train_groups(block_len, B) + 1) / 2; }

int tae_dec_forward_train_bf16(const TaeDecConfig* cfg, const void* packed, const float* received, const int32_t* perm,
                               const int32_t* inv_perm, float* out, float* trace, int32_t B, void* stash_y, void* stash_x,
                               void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_dec_forward_train_bf16: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed && received && perm && inv_perm && out && workspace && stash_y && stash_x, "tae_dec_forward_train_bf16: NULL pointer");
  return dec_forward_pair(*cfg, packed, received, perm, inv_perm, out, trace, B, workspace, workspace_bytes, (cudaStream_t)stream,
                          stash_y, stash_x);
}

size_t tae_dec_bwd_packed_bytes(const TaeDecConfig* cfg) {
  if (check_dec_config(cfg)) return 0;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return 0; }
  return dec_pair_bwd_packed_bytes(*cfg);
}

int tae_dec_pack_bwd_bf16(const TaeDecConfig* cfg, const float* params, void* packed_bwd, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed_bwd, "tae_dec_pack_bwd_bf16: NULL pointer");
  return dec_pair_pack_bwd(*cfg, params, packed_bwd, (cudaStream_t)stream);
}

int tae_dec_backward_bf16(const TaeDecConfig* cfg, const void* packed_bwd, const float* d_out_last, const int32_t* perm,
                          const int32_t* inv_perm, const void* stash_y, void* stash_g, void* stash_d, float* dxin_all, float* dlin_all,
                          float* grad_flat, int32_t B, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_dec_backward_bf16: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed_bwd && d_out_last && perm && inv_perm && stash_y && stash_g && stash_d && dxin_all && dlin_all && workspace,
              "tae_dec_backward_bf16: NULL pointer");
  return dec_backward_pair(*cfg, packed_bwd, d_out_last, perm, inv_perm, stash_y, stash_g, stash_d, dxin_all, dlin_all, grad_flat, B, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

int tae_dec_backward_range_bf16(const TaeDecConfig* cfg, const void* packed_bwd, const float* d_out_last, const int32_t* perm,
                                const int32_t* inv_perm, const void* stash_y, void* stash_g, void* stash_d, float* dxin_all, float* dlin_all,
                                float* grad_flat, int32_t B, int32_t unit_begin, int32_t unit_end, void* workspace, size_t workspace_bytes,
                                void* stream) {
  int rc = check_dec_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!bf16_supported(*cfg, &why)) { set_error("bf16 path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_dec_backward_range_bf16: negative batch %d", B);
  TAE_REQUIRE(unit_begin >= 0 && unit_end >= unit_begin && unit_end <= tae_train_units(cfg->block_len, B),
              "tae_dec_backward_range_bf16: work units [%d, %d) outside [0, %d)", unit_begin, unit_end, tae_train_units(cfg->block_len, B));
  if (B == 0 || unit_begin == unit_end) return TAE_OK;
  TAE_REQUIRE(packed_bwd && d_out_last && perm && inv_perm && stash_y && stash_g && stash_d && dxin_all && dlin_all && workspace,
              "tae_dec_backward_range_bf16: NULL pointer");
  return dec_backward_pair(*cfg, packed_bwd, d_out_last, perm, inv_perm, stash_y, stash_g, stash_d, dxin_all, dlin_all, grad_flat, B, workspace,
                           workspace_bytes, (cudaStream_t)stream, unit_begin, unit_end);
}

int tae_enc_forward_train_bf16(const TaeEncConfig* cfg, const void* packed, const float* u, const int32_t* perm, const int32_t* inv_perm,
                               float* x_tx, double* stats, int32_t B, void* stash_y, void* stash_x, void* workspace,
                               size_t workspace_bytes, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_enc_forward_train_bf16: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed && u && perm && inv_perm && x_tx && stats && workspace && stash_y && stash_x, "tae_enc_forward_train_bf16: NULL pointer");
  rc = enc_forward_pair(*cfg, packed, u, perm, inv_perm, x_tx, stats, B, workspace, workspace_bytes, (cudaStream_t)stream, stash_y, stash_x);
  if (rc) return rc;
  return launch_add_count(stats, (double)B * cfg->block_len * 3, (cudaStream_t)stream);
}

size_t tae_enc_bwd_packed_bytes(const TaeEncConfig* cfg) {
  if (check_enc_config(cfg)) return 0;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return 0; }
  return enc_pair_bwd_packed_bytes(*cfg);
}

int tae_enc_pack_bwd_bf16(const TaeEncConfig* cfg, const float* params, void* packed_bwd, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(params && packed_bwd, "tae_enc_pack_bwd_bf16: NULL pointer");
  return enc_pair_pack_bwd(*cfg, params, packed_bwd, (cudaStream_t)stream);
}

int tae_enc_backward_bf16(const TaeEncConfig* cfg, const void* packed_bwd, const float* dlin, const void* stash_y, void* stash_g, void* stash_d,
                          float* dxin_all, float* grad_flat, int32_t B, void* workspace, size_t workspace_bytes, void* stream) {
  int rc = check_enc_config(cfg);
  if (rc) return rc;
  const char* why = nullptr;
  if (!enc_pair_supported(*cfg, &why)) { set_error("bf16 encoder path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0, "tae_enc_backward_bf16: negative batch %d", B);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed_bwd && dlin && stash_y && stash_g && stash_d && dxin_all && workspace, "tae_enc_backward_bf16: NULL pointer");
  return enc_backward_pair(*cfg, packed_bwd, dlin, stash_y, stash_g, stash_d, dxin_all, grad_flat, B, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

int tae_wgrad_bf16(const TaeWgradJob* jobs_host, int32_t n_jobs, const void* jobs_dev, void* workspace, size_t workspace_bytes, void* stream) {
  TAE_REQUIRE(n_jobs >= 0, "tae_wgrad_bf16: negative job count");
  if (n_jobs == 0) return TAE_OK;
  TAE_REQUIRE(jobs_host && workspace, "tae_wgrad_bf16: NULL pointer");
  return launch_wgrad(jobs_host, n_jobs, jobs_dev, workspace, workspace_bytes, (cudaStream_t)stream);
}

int tae_gru_direction_f32(const float* xproj, const float* w_hh, const float* b_hh, float* out, int32_t B, int32_t L, int32_t H,
                          int32_t out_stride, int32_t out_offset, int32_t reverse, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && H >= 1, "tae_gru_direction_f32: bad shape B=%d L=%d H=%d", B, L, H);
  TAE_REQUIRE(out_stride >= H && out_offset >= 0 && out_offset + H <= out_stride, "tae_gru_direction_f32: bad output window");
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(xproj && w_hh && b_hh && out, "tae_gru_direction_f32: NULL pointer");
  return launch_gru_direction(xproj, w_hh, b_hh, out, B, L, H, out_stride, out_offset, reverse, (cudaStream_t)stream);
}

int tae_gru_direction_bwd_f32(const float* xproj, const float* w_hh, const float* b_hh, const float* hout, const float* dout,
                              float* dgi, float* dghn, int32_t B, int32_t L, int32_t H, int32_t io_stride, int32_t io_offset,
                              int32_t reverse, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && H >= 1, "tae_gru_direction_bwd_f32: bad shape B=%d L=%d H=%d", B, L, H);
  TAE_REQUIRE(io_stride >= H && io_offset >= 0 && io_offset + H <= io_stride, "tae_gru_direction_bwd_f32: bad hidden-state window");
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(xproj && w_hh && b_hh && hout && dout && dgi && dghn, "tae_gru_direction_bwd_f32: NULL pointer");
  return launch_gru_direction_bwd(xproj, w_hh, b_hh, hout, dout, dgi, dghn, B, L, H, io_stride, io_offset, reverse, (cudaStream_t)stream);
}

void tae_debug_gru_timeline(long long* dev) { gru_tc_set_timeline(dev); }

int32_t tae_gru_rows_per_block(int32_t B) { return gru_tc_rows_per_block(B < 0 ? 0 : B); }

size_t tae_gru_tile_bytes(int32_t B, int32_t L, int32_t n_chunks, int32_t R) {
  if (B < 0 || L < 1 || n_chunks < 1 || !(R == 16 || (R >= 32 && R <= 128 && R % 32 == 0))) return 0;
  const size_t n_blk = 2 * ((((size_t)B + R - 1) / R + 1) / 2);
  return n_blk * (size_t)L * n_chunks * R * 16;
}

size_t tae_gru_packed_bytes(int32_t H, int32_t in_ch, int32_t grp_valid) {
  const char* why = nullptr;
  if (!gru_tc_supported(H, in_ch, grp_valid, &why)) { set_error("bf16 GRU path: %s", why); return 0; }
  return gru_tc_packed_bytes(H, in_ch, grp_valid);
}

int tae_gru_pack_bf16(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* packed, int32_t H, int32_t in_ch,
                      int32_t grp_valid, void* stream) {
  const char* why = nullptr;
  if (!gru_tc_supported(H, in_ch, grp_valid, &why)) { set_error("bf16 GRU path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(w_ih && w_hh && b_ih && b_hh && packed, "tae_gru_pack_bf16: NULL pointer");
  return gru_tc_pack(w_ih, w_hh, b_ih, b_hh, packed, H, in_ch, grp_valid, (cudaStream_t)stream);
}

int tae_gru_direction_bf16(const void* packed, const void* x_tiles, void* out_tiles, int32_t B, int32_t L, int32_t H, int32_t in_ch,
                           int32_t grp_valid, int32_t R, int32_t out_chunks, int32_t out_chunk0, int32_t reverse, void* workspace,
                           size_t workspace_bytes, void* stream) {
  const char* why = nullptr;
  if (!gru_tc_supported(H, in_ch, grp_valid, &why)) { set_error("bf16 GRU path: %s", why); return TAE_EUNSUPPORTED; }
  TAE_REQUIRE(B >= 0 && L >= 1, "tae_gru_direction_bf16: bad shape B=%d L=%d", B, L);
  TAE_REQUIRE(R == 16 || (R >= 32 && R <= 128 && R % 32 == 0), "tae_gru_direction_bf16: rows per block %d (16, 32, 64, 96 or 128)", R);
  TAE_REQUIRE(out_chunk0 >= 0 && out_chunk0 + (H + 7) / 8 <= out_chunks, "tae_gru_direction_bf16: bad output chunk window");
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(packed && x_tiles && out_tiles && workspace, "tae_gru_direction_bf16: NULL pointer");
  return gru_tc_direction(packed, x_tiles, out_tiles, B, L, H, in_ch, grp_valid, R, out_chunks, out_chunk0, reverse, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

int tae_gru_tiles_from_f32(const float* x, void* tiles, int32_t B, int32_t L, int32_t C, int32_t R, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && C >= 1 && (R == 16 || (R >= 32 && R <= 128 && R % 32 == 0)), "tae_gru_tiles_from_f32: bad shape");
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(x && tiles, "tae_gru_tiles_from_f32: NULL pointer");
  return gru_tc_tiles_from_f32(x, tiles, B, L, C, R, (cudaStream_t)stream);
}

int tae_gru_linear_f32(const void* tiles, const float* weight, const float* bias, float* out, int32_t B, int32_t L, int32_t in_ch,
                       int32_t grp_valid, int32_t F, int32_t R, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1 && F >= 1 && F <= 8 && (R == 16 || (R >= 32 && R <= 128 && R % 32 == 0)), "tae_gru_linear_f32: bad shape (F <= 8)");
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(tiles && weight && bias && out, "tae_gru_linear_f32: NULL pointer");
  return gru_tc_linear(tiles, weight, bias, out, B, L, in_ch, grp_valid, F, R, (cudaStream_t)stream);
}

int tae_awgn_f32(const float* codes, float* received, size_t n, float sigma, uint64_t seed, uint64_t offset, void* stream) {
  if (n == 0) return TAE_OK;
  TAE_REQUIRE(codes && received, "tae_awgn_f32: NULL pointer");
  TAE_REQUIRE(sigma >= 0.f, "tae_awgn_f32: negative sigma");
  return launch_awgn(codes, received, n, sigma, seed, offset, (cudaStream_t)stream);
}

int tae_error_count_f32(const float* y_true, const float* y_pred, int32_t B, int32_t L, unsigned long long* counts, void* stream) {
  TAE_REQUIRE(B >= 0 && L >= 1, "tae_error_count_f32: bad shape B=%d L=%d", B, L);
  if (B == 0) return TAE_OK;
  TAE_REQUIRE(y_true && y_pred && counts, "tae_error_count_f32: NULL pointer");
  return launch_error_count(y_true, y_pred, B, L, counts, (cudaStream_t)stream);
}

}  // extern "C"
