// DEC_LargeRNN building block (BASELINE config 5): one direction of one GRU layer.
//
// The reference uses torch.nn.GRU (decoders.py:43-52: 2 layers, bidirectional, batch_first); its arithmetic, restated in
// oracle/turboae_oracle.py::gru_direction, is PyTorch's documented GRU cell
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
// The input projections W_i* x + b_i* of all time steps are ONE pointwise GEMM (tae_conv1d_elu_f32 with K = 1); this kernel
// is the sequential part: a persistent batch-parallel recurrence with W_h* resident in shared memory.
//   CTA = 16 codewords x H units; thread (j, q) owns unit j of codewords 4q..4q+3, so every W_h* element read from shared
//   memory feeds 4 FMAs and the 4 hidden values of a step come as one float4.
#include "tae_common.cuh"

namespace tae {

namespace {

constexpr int GRU_NB = 16;      // codewords per CTA
constexpr int GRU_RB = 4;       // codewords per thread

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__global__ void __launch_bounds__(512, 1)
gru_direction_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                     float* __restrict__ out, int B, int L, int H, int out_stride, int out_offset, int reverse) {
  extern __shared__ __align__(16) float sm[];
  float* W_s = sm;                                   // [k][3H]  (transposed: consecutive j are consecutive addresses)
  float* h_s = W_s + (size_t)H * 3 * H;              // [2][k][GRU_NB]
  const int nthr = blockDim.x;
  for (int i = threadIdx.x; i < 3 * H * H; i += nthr) {
    const int g = i / H, k = i % H;                  // w_hh is (3H, H) row-major: row g (gate*H + j), column k
    W_s[(size_t)k * 3 * H + g] = w_hh[i];
  }
  const int j = threadIdx.x % H, q = threadIdx.x / H;       // blockDim.x == H * (GRU_NB / GRU_RB)
  const bool active = q < GRU_NB / GRU_RB;
  const float bhr = b_hh[j], bhz = b_hh[H + j], bhn = b_hh[2 * H + j];
  const int n_groups = (B + GRU_NB - 1) / GRU_NB;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int b0 = grp * GRU_NB + q * GRU_RB;
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H * GRU_NB; i += nthr) h_s[i] = 0.f;            // h_0 = 0
    __syncthreads();
    float hp[GRU_RB] = {0.f, 0.f, 0.f, 0.f};
    int cur = 0;
    for (int s = 0; s < L; ++s) {
      const int t = reverse ? L - 1 - s : s;
      // this step's input projections (global; issued before the dot products so that their latency is hidden)
      float xr[GRU_RB], xz[GRU_RB], xn[GRU_RB];
#pragma unroll
      for (int c = 0; c < GRU_RB; ++c) {
        const bool ok = active && (b0 + c < B);
        const float* xp = xproj + ((size_t)(b0 + c) * L + t) * 3 * H;
        xr[c] = ok ? __ldg(xp + j) : 0.f;
        xz[c] = ok ? __ldg(xp + H + j) : 0.f;
        xn[c] = ok ? __ldg(xp + 2 * H + j) : 0.f;
      }
      float ar[GRU_RB] = {0.f, 0.f, 0.f, 0.f}, az[GRU_RB] = {0.f, 0.f, 0.f, 0.f}, an[GRU_RB] = {0.f, 0.f, 0.f, 0.f};
      if (active) {
        const float* hc = h_s + (size_t)cur * H * GRU_NB + q * GRU_RB;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
          const float4 hv = *reinterpret_cast<const float4*>(hc + (size_t)k * GRU_NB);
          const float wr = W_s[(size_t)k * 3 * H + j], wz = W_s[(size_t)k * 3 * H + H + j], wn = W_s[(size_t)k * 3 * H + 2 * H + j];
          ar[0] = fmaf(wr, hv.x, ar[0]); ar[1] = fmaf(wr, hv.y, ar[1]); ar[2] = fmaf(wr, hv.z, ar[2]); ar[3] = fmaf(wr, hv.w, ar[3]);
          az[0] = fmaf(wz, hv.x, az[0]); az[1] = fmaf(wz, hv.y, az[1]); az[2] = fmaf(wz, hv.z, az[2]); az[3] = fmaf(wz, hv.w, az[3]);
          an[0] = fmaf(wn, hv.x, an[0]); an[1] = fmaf(wn, hv.y, an[1]); an[2] = fmaf(wn, hv.z, an[2]); an[3] = fmaf(wn, hv.w, an[3]);
        }
        float* hn = h_s + (size_t)(cur ^ 1) * H * GRU_NB + (size_t)j * GRU_NB + q * GRU_RB;
        float hnew[GRU_RB];
#pragma unroll
        for (int c = 0; c < GRU_RB; ++c) {
          const float r = sigmoidf_(xr[c] + ar[c] + bhr);
          const float z = sigmoidf_(xz[c] + az[c] + bhz);
          const float n = tanhf(xn[c] + r * (an[c] + bhn));
          hnew[c] = (1.f - z) * n + z * hp[c];
          hp[c] = hnew[c];
          if (b0 + c < B) out[((size_t)(b0 + c) * L + t) * out_stride + out_offset + j] = hnew[c];
        }
        *reinterpret_cast<float4*>(hn) = make_float4(hnew[0], hnew[1], hnew[2], hnew[3]);
      }
      cur ^= 1;
      __syncthreads();
    }
  }
}

// Backward of the recurrence (training of DEC_LargeRNN: reference trainer.py:74 backpropagates through torch.nn.GRU,
// decoders.py:43-52).  Walks the steps in the opposite order of the forward kernel.  Per step and unit j, with the gates
// recomputed from the stored input projections and h_{t-1} (nothing but h is kept by the forward pass):
//     dh    = dout_t + carry                       dn = dh (1 - z)            dz = dh (h_{t-1} - n)
//     dn'   = dn (1 - n^2)                          dr = dn' (W_hn h + b_hn)
//     dgi_t = [dr r (1 - r), dz z (1 - z), dn']     (gradient at W_i x + b_i: the caller's GEMMs turn it into dW_ih, db_ih, dx)
//     dgh_t = [dgi_r, dgi_z, dn' r]                 (gradient at W_h h + b_h: dW_hh, db_hh; only its n part is stored, dghn)
//     carry = dh z + dgh_t W_hh                     (gradient flowing to h_{t-1})
// Same thread mapping as the forward kernel; W_hh is kept ONCE in shared memory as [3H][H + 1] (odd row pitch): phase A
// (unit j sums over k, rows j of the three gates) and phase B (unit k sums over the 3H rows) both read it conflict-free.
__global__ void __launch_bounds__(512, 1)
gru_direction_bwd_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                         const float* __restrict__ hout, const float* __restrict__ dout, float* __restrict__ dgi,
                         float* __restrict__ dghn, int B, int L, int H, int io_stride, int io_offset, int reverse) {
  extern __shared__ __align__(16) float sm[];
  const int P = H + 1;
  float* W_p = sm;                                   // [3H][H + 1]
  float* h_s = W_p + (size_t)3 * H * P;              // [k][GRU_NB]      h_{t-1}
  float* g_s = h_s + (size_t)H * GRU_NB;             // [3][j][GRU_NB]   dgh of this step
  const int nthr = blockDim.x;
  for (int i = threadIdx.x; i < 3 * H * H; i += nthr) W_p[(size_t)(i / H) * P + (i % H)] = w_hh[i];
  const int j = threadIdx.x % H, q = threadIdx.x / H;
  const bool active = q < GRU_NB / GRU_RB;
  const float bhr = b_hh[j], bhz = b_hh[H + j], bhn = b_hh[2 * H + j];
  const int n_groups = (B + GRU_NB - 1) / GRU_NB;
  for (int grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
    const int b0 = grp * GRU_NB + q * GRU_RB;
    float carry[GRU_RB] = {0.f, 0.f, 0.f, 0.f};
    for (int s = L - 1; s >= 0; --s) {
      const int t = reverse ? L - 1 - s : s;
      const int tp = reverse ? t + 1 : t - 1;          // time index of h_{t-1} in the direction of the recurrence (s > 0)
      float hp[GRU_RB], xr[GRU_RB], xz[GRU_RB], xn[GRU_RB], dh[GRU_RB];
#pragma unroll
      for (int c = 0; c < GRU_RB; ++c) {
        const bool ok = active && (b0 + c < B);
        const size_t row = (size_t)(b0 + c) * L;
        hp[c] = (ok && s > 0) ? __ldg(hout + (row + tp) * io_stride + io_offset + j) : 0.f;
        const float* xp = xproj + (row + t) * 3 * H;
        xr[c] = ok ? __ldg(xp + j) : 0.f;
        xz[c] = ok ? __ldg(xp + H + j) : 0.f;
        xn[c] = ok ? __ldg(xp + 2 * H + j) : 0.f;
        dh[c] = (ok ? __ldg(dout + (row + t) * io_stride + io_offset + j) : 0.f) + carry[c];
      }
      __syncthreads();                                 // phase B of the previous step has read g_s; phase A of it has read h_s
      if (active) *reinterpret_cast<float4*>(h_s + (size_t)j * GRU_NB + q * GRU_RB) = make_float4(hp[0], hp[1], hp[2], hp[3]);
      __syncthreads();
      if (active) {
        // ---- phase A: recompute the gates of unit j, gradients at the pre-activations
        float ar[GRU_RB] = {0.f, 0.f, 0.f, 0.f}, az[GRU_RB] = {0.f, 0.f, 0.f, 0.f}, an[GRU_RB] = {0.f, 0.f, 0.f, 0.f};
        const float* hc = h_s + q * GRU_RB;
        const float* wr_ = W_p + (size_t)j * P;
        const float* wz_ = W_p + (size_t)(H + j) * P;
        const float* wn_ = W_p + (size_t)(2 * H + j) * P;
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
          const float4 hv = *reinterpret_cast<const float4*>(hc + (size_t)k * GRU_NB);
          const float wr = wr_[k], wz = wz_[k], wn = wn_[k];
          ar[0] = fmaf(wr, hv.x, ar[0]); ar[1] = fmaf(wr, hv.y, ar[1]); ar[2] = fmaf(wr, hv.z, ar[2]); ar[3] = fmaf(wr, hv.w, ar[3]);
          az[0] = fmaf(wz, hv.x, az[0]); az[1] = fmaf(wz, hv.y, az[1]); az[2] = fmaf(wz, hv.z, az[2]); az[3] = fmaf(wz, hv.w, az[3]);
          an[0] = fmaf(wn, hv.x, an[0]); an[1] = fmaf(wn, hv.y, an[1]); an[2] = fmaf(wn, hv.z, an[2]); an[3] = fmaf(wn, hv.w, an[3]);
        }
        float gr[GRU_RB], gz[GRU_RB], gn[GRU_RB];
#pragma unroll
        for (int c = 0; c < GRU_RB; ++c) {
          const float r = sigmoidf_(xr[c] + ar[c] + bhr);
          const float z = sigmoidf_(xz[c] + az[c] + bhz);
          const float hl = an[c] + bhn;
          const float n = tanhf(xn[c] + r * hl);
          const float dn_pre = dh[c] * (1.f - z) * (1.f - n * n);
          const float dz_pre = dh[c] * (hp[c] - n) * z * (1.f - z);
          const float dr_pre = dn_pre * hl * r * (1.f - r);
          gr[c] = dr_pre; gz[c] = dz_pre; gn[c] = dn_pre * r;
          carry[c] = dh[c] * z;
          if (b0 + c < B) {
            float* o = dgi + ((size_t)(b0 + c) * L + t) * 3 * H;
            o[j] = dr_pre; o[H + j] = dz_pre; o[2 * H + j] = dn_pre;
            dghn[((size_t)(b0 + c) * L + t) * H + j] = gn[c];
          }
        }
        float* gs = g_s + (size_t)j * GRU_NB + q * GRU_RB;
        *reinterpret_cast<float4*>(gs) = make_float4(gr[0], gr[1], gr[2], gr[3]);
        *reinterpret_cast<float4*>(gs + (size_t)H * GRU_NB) = make_float4(gz[0], gz[1], gz[2], gz[3]);
        *reinterpret_cast<float4*>(gs + (size_t)2 * H * GRU_NB) = make_float4(gn[0], gn[1], gn[2], gn[3]);
      }
      __syncthreads();
      if (active && s > 0) {
        // ---- phase B: carry_k += sum over the 3H rows of dgh * W_hh[:, k]   (k = this thread's unit)
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* gc = g_s + q * GRU_RB;
#pragma unroll 4
        for (int r3 = 0; r3 < 3 * H; ++r3) {
          const float4 gv = *reinterpret_cast<const float4*>(gc + (size_t)r3 * GRU_NB);
          const float w = W_p[(size_t)r3 * P + j];
          a0 = fmaf(w, gv.x, a0); a1 = fmaf(w, gv.y, a1); a2 = fmaf(w, gv.z, a2); a3 = fmaf(w, gv.w, a3);
        }
        carry[0] += a0; carry[1] += a1; carry[2] += a2; carry[3] += a3;
      }
    }
    __syncthreads();
  }
}

}  // namespace

int launch_gru_direction_bwd(const float* xproj, const float* w_hh, const float* b_hh, const float* hout, const float* dout, float* dgi,
                             float* dghn, int B, int L, int H, int io_stride, int io_offset, int reverse, cudaStream_t s) {
  if (B == 0) return TAE_OK;
  const int threads = H * (GRU_NB / GRU_RB);
  if (H < 1 || H > 128 || threads > 512) { set_error("tae_gru_direction_bwd_f32: hidden size %d unsupported (1..128)", H); return TAE_EUNSUPPORTED; }
  const size_t smem = ((size_t)3 * H * (H + 1) + 4 * (size_t)H * GRU_NB) * sizeof(float);
  static DeviceOnce once;
  int n_sm = 0;
  {
    int rc = device_once(once, "gru_direction_bwd_kernel", [](int) -> int {
      cudaError_t e = cudaFuncSetAttribute(gru_direction_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gru_direction_bwd_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
      return TAE_OK;
    }, &n_sm);
    if (rc) return rc;
  }
  if (smem > 220 * 1024) { set_error("tae_gru_direction_bwd_f32: hidden size %d needs %zu B of shared memory", H, smem); return TAE_EUNSUPPORTED; }
  const int n_groups = (B + GRU_NB - 1) / GRU_NB;
  gru_direction_bwd_kernel<<<std::min(n_groups, n_sm), threads, smem, s>>>(xproj, w_hh, b_hh, hout, dout, dgi, dghn, B, L, H, io_stride,
                                                                          io_offset, reverse);
  return after_launch("gru_direction_bwd_kernel");
}

int launch_gru_direction(const float* xproj, const float* w_hh, const float* b_hh, float* out, int B, int L, int H, int out_stride,
                         int out_offset, int reverse, cudaStream_t s) {
  if (B == 0) return TAE_OK;
  const int threads = H * (GRU_NB / GRU_RB);
  if (H < 1 || H > 128 || threads > 512) { set_error("tae_gru_direction_f32: hidden size %d unsupported (1..128)", H); return TAE_EUNSUPPORTED; }
  const size_t smem = ((size_t)3 * H * H + 2 * (size_t)H * GRU_NB) * sizeof(float);
  static DeviceOnce once;
  int n_sm = 0;
  {
    int rc = device_once(once, "gru_direction_kernel", [](int) -> int {
      cudaError_t e = cudaFuncSetAttribute(gru_direction_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gru_direction_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
      return TAE_OK;
    }, &n_sm);
    if (rc) return rc;
  }
  const int n_groups = (B + GRU_NB - 1) / GRU_NB;
  gru_direction_kernel<<<std::min(n_groups, n_sm), threads, smem, s>>>(xproj, w_hh, b_hh, out, B, L, H, out_stride, out_offset, reverse);
  return after_launch("gru_direction_kernel");
}

}  // namespace tae
