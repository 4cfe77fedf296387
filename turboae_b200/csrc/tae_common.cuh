// Shared helpers of libturboae_b200.so (error convention, launch accounting, layouts).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <algorithm>
#include <mutex>

#include "turboae_b200.h"

namespace tae {

// Thread-local error message behind tae_last_error().
void set_error(const char* fmt, ...);
// Records a launch (bench.py's "gpu_launches") and converts a launch failure into TAE_ECUDA.
int after_launch(const char* kernel_name);
void count_launch();

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One-time setup PER DEVICE (cudaFuncSetAttribute, capability check and SM count are properties of a device, and one
// process may drive several devices from several threads): `fn(dev)` runs once for each device under the mutex.
constexpr int TAE_MAX_DEVICES = 64;
struct DeviceOnce {
  std::mutex mu;
  bool done[TAE_MAX_DEVICES] = {};
  int n_sm[TAE_MAX_DEVICES] = {};
};
template <class F>
int device_once(DeviceOnce& st, const char* who, F&& fn, int* n_sm_out = nullptr) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess || dev < 0 || dev >= TAE_MAX_DEVICES) {
    set_error("%s: cudaGetDevice: %s (device %d)", who, cudaGetErrorString(e), dev);
    return TAE_ECUDA;
  }
  std::lock_guard<std::mutex> lk(st.mu);
  if (!st.done[dev]) {
    int n_sm = 0;
    e = cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { set_error("%s: cudaDeviceGetAttribute: %s", who, cudaGetErrorString(e)); return TAE_ECUDA; }
    int rc = fn(dev);
    if (rc) return rc;
    st.n_sm[dev] = n_sm;
    st.done[dev] = true;
  }
  if (n_sm_out) *n_sm_out = st.n_sm[dev];
  return TAE_OK;
}
// sm_100 check shared by the tcgen05 paths
int require_sm100(int dev, const char* who);
// Where a kernel's bounded barrier wait records its code before it traps: a page of pinned, device-mapped HOST memory per
// device, so that the code survives the context error (tae_last_error() reports it).  Falls back to `fallback` (the caller's
// workspace) when the mapping cannot be had.
int* wait_code_slot(void* fallback);

// ---- canonical flat-parameter layouts (see include/turboae_b200.h) ---------------------
struct ConvLayer {
  size_t w_off;   // offset (floats) of weight (Cout, Cin, K) in the flat parameter buffer
  size_t b_off;   // offset of bias (Cout)
  int cin, cout;
};

struct DecStackLayout {
  ConvLayer conv[16];
  size_t lin_w_off, lin_b_off;
  int fout;
};

// stack index = 2*idx + s  (s = 0: dec1, s = 1: dec2)
inline size_t dec_layout(const TaeDecConfig& c, DecStackLayout* out /* 2*I entries or nullptr */) {
  size_t off = 0;
  for (int idx = 0; idx < c.num_iteration; ++idx) {
    for (int s = 0; s < 2; ++s) {
      DecStackLayout st{};
      for (int j = 0; j < c.num_layer; ++j) {
        int cin = (j == 0) ? (2 + c.num_iter_ft) : c.num_unit;
        st.conv[j].cin = cin;
        st.conv[j].cout = c.num_unit;
        st.conv[j].w_off = off; off += (size_t)c.num_unit * cin * c.kernel_size;
        st.conv[j].b_off = off; off += (size_t)c.num_unit;
      }
      st.fout = (s == 1 && idx == c.num_iteration - 1) ? 1 : c.num_iter_ft;
      st.lin_w_off = off; off += (size_t)st.fout * c.num_unit;
      st.lin_b_off = off; off += (size_t)st.fout;
      if (out) out[2 * idx + s] = st;
    }
  }
  return off;
}

struct EncBranchLayout {
  ConvLayer conv[16];
  size_t lin_w_off, lin_b_off;
};

inline size_t enc_layout(const TaeEncConfig& c, EncBranchLayout* out /* 3 entries or nullptr */) {
  size_t off = 0;
  for (int br = 0; br < 3; ++br) {
    EncBranchLayout st{};
    for (int j = 0; j < c.num_layer; ++j) {
      int cin = (j == 0) ? 1 : c.num_unit;
      st.conv[j].cin = cin;
      st.conv[j].cout = c.num_unit;
      st.conv[j].w_off = off; off += (size_t)c.num_unit * cin * c.kernel_size;
      st.conv[j].b_off = off; off += (size_t)c.num_unit;
    }
    st.lin_w_off = off; off += (size_t)c.num_unit;
    st.lin_b_off = off; off += 1;
    if (out) out[br] = st;
  }
  return off;
}

int check_dec_config(const TaeDecConfig* cfg);
int check_enc_config(const TaeEncConfig* cfg);

// ---- fp32 CUDA-core path (tae_f32.cu) -----------------------------------------------------
size_t conv_packed_floats(int cin, int cout, int k);
int launch_pack_conv_f32(const float* w, float* packed, int cin, int cout, int k, cudaStream_t s);
int launch_conv_f32(const float* in, float* out, const float* packed, const float* bias, int B, int L,
                    int cin, int cout, int k, int apply_elu, cudaStream_t s);
size_t conv_bwd_packed_floats(int cin, int cout, int k);
int launch_conv_bwd_f32(const float* x, const float* y, const float* dy, const float* w, float* dx, float* dw, float* db, int B, int L,
                        int cin, int cout, int k, int apply_elu, float* packed_ws, cudaStream_t s);
int launch_interleave_f32(const float* in, float* out, const int32_t* perm, int B, int L, int F, cudaStream_t s);
size_t dec_workspace_bytes_f32(const TaeDecConfig& c, int B);
int dec_forward_f32(const TaeDecConfig& c, const float* params, const float* received, const int32_t* perm,
                    const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes,
                    cudaStream_t s);
size_t enc_workspace_bytes_f32(const TaeEncConfig& c, int B);
int enc_forward_f32(const TaeEncConfig& c, const float* params, const float* u, const int32_t* perm, float* x_tx,
                    double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_power_norm_f32(const float* x, float* codes, size_t n, const double* stats, float* mean_std, float limit, float q,
                          cudaStream_t s);

int launch_dec_out_bwd_f32(const float* d_out, const float* out, const int32_t* perm, float* d_last, int B, int L, cudaStream_t s);
int launch_dec_input_grad_f32(const float* dxin_all, const int32_t* inv_perm, float* d_received, int n_stacks, int B, int L, cudaStream_t s);
int launch_enc_out_bwd_f32(const float* d_x, const float* x_tx, float* dlin, int B, int L, cudaStream_t s);
int launch_power_sums_f32(const float* a, const float* y, size_t n, double* out, cudaStream_t s);
int launch_power_norm_bwd_f32(const float* g, const float* y, float* dx, size_t n, const double* sums, const double* stats,
                              const float* mean_std, cudaStream_t s);

// ---- bf16 tcgen05 path: fused CTA-pair kernel (tae_dec_pair.cu) -----------------------------
bool dec_pair_supported(const TaeDecConfig& c, const char** why);
size_t dec_pair_packed_bytes(const TaeDecConfig& c);
int dec_pair_pack(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s);
int dec_forward_pair(const TaeDecConfig& c, const void* packed, const float* received, const int32_t* perm,
                     const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes, cudaStream_t s,
                     void* stash_y = nullptr, void* stash_x = nullptr);
// training on the tensor cores (tae_dec_pair.cu MODE 1, tae_wgrad.cu)
int train_groups(int block_len, int B);
size_t dec_pair_bwd_packed_bytes(const TaeDecConfig& c);
int dec_pair_pack_bwd(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s);
int dec_backward_pair(const TaeDecConfig& c, const void* packed_bwd, const float* d_out_last, const int32_t* perm, const int32_t* inv_perm,
                      const void* stash_y, void* stash_g, void* stash_d, float* dxin_all, float* dlin_all, float* grad_flat, int B, void* ws,
                      size_t ws_bytes, cudaStream_t s, int pair_begin = 0, int pair_end = -1);
int launch_wgrad(const TaeWgradJob* jobs_host, int n_jobs, const void* jobs_dev, void* ws, size_t ws_bytes, cudaStream_t s);
// ENC_interCNN on the same CTA-pair kernel (three branches as three stacks)
bool enc_pair_supported(const TaeEncConfig& c, const char** why);
size_t enc_pair_packed_bytes(const TaeEncConfig& c);
int enc_pair_pack(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s);
int enc_forward_pair(const TaeEncConfig& c, const void* packed, const float* u, const int32_t* perm, const int32_t* inv_perm,
                     float* x_tx, double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s,
                     void* stash_y = nullptr, void* stash_x = nullptr);
size_t enc_pair_bwd_packed_bytes(const TaeEncConfig& c);
int enc_pair_pack_bwd(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s);
int enc_backward_pair(const TaeEncConfig& c, const void* packed_bwd, const float* dlin, const void* stash_y, void* stash_g, void* stash_d,
                      float* dxin_all, float* grad_flat, int B, void* ws, size_t ws_bytes, cudaStream_t s);
int launch_add_count(double* stats, double n, cudaStream_t s);
// ---- f16x3 tcgen05 path: split-operand kernel with fp32-class accuracy (tae_x3.cu) ----------
bool dec_x3_supported(const TaeDecConfig& c, const char** why);
size_t dec_x3_packed_bytes(const TaeDecConfig& c);
int dec_x3_pack(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s);
int dec_forward_x3(const TaeDecConfig& c, const float* params, const void* packed, const float* received, const int32_t* perm,
                   const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes, cudaStream_t s);
bool enc_x3_supported(const TaeEncConfig& c, const char** why);
size_t enc_x3_packed_bytes(const TaeEncConfig& c);
int enc_x3_pack(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s);
int enc_forward_x3(const TaeEncConfig& c, const float* params, const void* packed, const float* u, const int32_t* perm,
                   const int32_t* inv_perm, float* x_tx, double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s);
// ---- DEC_LargeRNN recurrence (tae_gru.cu) ----------------------------------------------------
int launch_gru_direction_bwd(const float* xproj, const float* w_hh, const float* b_hh, const float* hout, const float* dout, float* dgi,
                             float* dghn, int B, int L, int H, int io_stride, int io_offset, int reverse, cudaStream_t s);
int launch_gru_direction(const float* xproj, const float* w_hh, const float* b_hh, float* out, int B, int L, int H, int out_stride,
                         int out_offset, int reverse, cudaStream_t s);
// tensor-core recurrence (tae_gru_tc.cu)
bool gru_tc_supported(int H, int in_ch, int grp_valid, const char** why);
void gru_tc_set_timeline(long long* dev);
int gru_tc_rows_per_block(int B);
size_t gru_tc_packed_bytes(int H, int in_ch, int grp_valid);
int gru_tc_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* packed, int H, int in_ch, int grp_valid,
                cudaStream_t s);
int gru_tc_direction(const void* packed, const void* x_tiles, void* out_tiles, int B, int L, int H, int in_ch, int grp_valid, int R,
                     int out_chunks, int out_c0, int reverse, void* ws, size_t ws_bytes, cudaStream_t s);
int gru_tc_tiles_from_f32(const float* x, void* tiles, int B, int L, int C, int R, cudaStream_t s);
int gru_tc_linear(const void* tiles, const float* w, const float* bias, float* out, int B, int L, int in_ch, int grp_valid, int F, int R,
                  cudaStream_t s);
// ---- channel + metrics (tae_channel.cu) ------------------------------------------------------
int launch_awgn(const float* codes, float* received, size_t n, float sigma, uint64_t seed, uint64_t offset, cudaStream_t s);
int launch_error_count(const float* y_true, const float* y_pred, int B, int L, unsigned long long* counts, cudaStream_t s);

}  // namespace tae
