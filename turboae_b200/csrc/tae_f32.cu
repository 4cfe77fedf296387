// fp32 CUDA-core path of the TurboAE hot path (TAE_PRECISION_FP32).
//
// This is the elementwise-parity anchor (<= 1e-4 vs the reference's fp32 forward, measured
// ~1e-6): layer-at-a-time kernels with fp32 FMA accumulation, activations in HBM between
// layers.  The throughput path is the fused tcgen05 kernel in tae_dec_pair.cu.
//
// Reference lines restated here (paths relative to the reference checkout):
//   conv layer + ELU ............ cnn_utils.py:15-22, 36-46
//   Interleaver / DeInterleaver . interleavers.py:15-21, 43-48
//   DEC_LargeCNN.forward ........ decoders.py:219-269
//   ENC_interCNN.forward ........ encoders.py:362-375 ; power_constraint encoders.py:107-116
#include <algorithm>

#include "tae_common.cuh"

namespace tae {

namespace {

constexpr int CONV_TN = 128;  // output channels per CTA (4 per lane)
constexpr int CONV_CC = 16;   // input channels staged per shared-memory chunk

__device__ __forceinline__ float elu1(float z) { return z > 0.f ? z : expm1f(z); }

// ---------------------------------------------------------------------------------------
// weight re-layout: torch Conv1d (Cout, Cin, K)  ->  [Cin][K][CoutPad] (zero padded), so that
// one (c, t) row of 128 output channels is a contiguous, 16-byte aligned run.
// ---------------------------------------------------------------------------------------
__global__ void pack_conv_f32_kernel(const float* __restrict__ w, float* __restrict__ packed, int cin, int cout,
                                     int k, int cout_pad) {
  size_t n = (size_t)cin * k * cout_pad;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int o = (int)(idx % cout_pad);
    int t = (int)((idx / cout_pad) % k);
    int c = (int)(idx / ((size_t)cout_pad * k));
    packed[idx] = (o < cout) ? w[((size_t)o * cin + c) * k + t] : 0.f;
  }
}

// ---------------------------------------------------------------------------------------
// One Conv1d(K, stride 1, pad K/2) + bias (+ ELU) on a channel-last (B, L, Cin) tensor.
// CTA = one tile of TM (= 8 * warps) consecutive positions of ONE codeword x 128 output
// channels; zero padding at the codeword edges is materialised while staging, so the inner
// loop has no masks.  Warp w owns rows 8w..8w+7, lane owns channels 4*lane..4*lane+3:
// per input channel a lane keeps a sliding window of 8+K-1 inputs in registers and issues
// 8*K*4 FMAs against K float4 weight loads.
// ---------------------------------------------------------------------------------------
// BWD: `in` is dy and `yfwd` the forward output of the same layer; the staged value is g = dy * ELU'(z) with
// ELU'(z) = 1 for y > 0 and y + 1 otherwise (y = e^z - 1  =>  e^z = y + 1), i.e. the backward-data pass of a layer is this same
// kernel run with transposed, tap-flipped weights on g.
template <int K, bool BWD>
__global__ void __launch_bounds__(512)
conv1d_f32_kernel(const float* __restrict__ in, const float* __restrict__ yfwd, float* __restrict__ out,
                  const float* __restrict__ wp, const float* __restrict__ bias, int L, int cin, int cout, int cout_pad,
                  int tiles_per_cw, int TM, int AS, int apply_elu) {
  extern __shared__ __align__(16) float smem[];
  float* W_s = smem;                          // [CC][K][128]
  float* A_s = smem + CONV_CC * K * CONV_TN;  // [CC][AS]
  constexpr int PAD = K / 2;
  constexpr int NWIN = 8 + K - 1;
  constexpr int NV = (NWIN + 3) / 4;

  const int b = blockIdx.x / tiles_per_cw;
  const int l0 = (blockIdx.x % tiles_per_cw) * TM;
  const int n0 = blockIdx.y * CONV_TN;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* in_cw = in + (size_t)b * L * cin;
  const float* y_cw = BWD ? yfwd + (size_t)b * L * cin : nullptr;
  const int rows = TM + K - 1;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int c0 = 0; c0 < cin; c0 += CONV_CC) {
    const int cc_n = min(CONV_CC, cin - c0);
    for (int i = threadIdx.x; i < cc_n * K * 32; i += blockDim.x) {
      const int v = i & 31, ct = i >> 5;
      const float4 val = *reinterpret_cast<const float4*>(wp + ((size_t)c0 * K + ct) * cout_pad + n0 + v * 4);
      *reinterpret_cast<float4*>(W_s + ct * CONV_TN + v * 4) = val;
    }
    for (int i = threadIdx.x; i < rows * cc_n; i += blockDim.x) {
      const int cc = i % cc_n, r = i / cc_n;
      const int l = l0 - PAD + r;
      float v = 0.f;
      if (l >= 0 && l < L) {
        v = in_cw[(size_t)l * cin + c0 + cc];
        if (BWD && apply_elu) {
          const float y = y_cw[(size_t)l * cin + c0 + cc];
          v *= (y > 0.f) ? 1.f : (y + 1.f);
        }
      }
      A_s[cc * AS + r] = v;
    }
    __syncthreads();
    for (int cc = 0; cc < cc_n; ++cc) {
      float win[4 * NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 a = *reinterpret_cast<const float4*>(A_s + cc * AS + warp * 8 + v * 4);
        win[4 * v + 0] = a.x; win[4 * v + 1] = a.y; win[4 * v + 2] = a.z; win[4 * v + 3] = a.w;
      }
#pragma unroll
      for (int t = 0; t < K; ++t) {
        const float4 w = *reinterpret_cast<const float4*>(W_s + (cc * K + t) * CONV_TN + lane * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float a = win[i + t];
          acc[i][0] = fmaf(a, w.x, acc[i][0]);
          acc[i][1] = fmaf(a, w.y, acc[i][1]);
          acc[i][2] = fmaf(a, w.z, acc[i][2]);
          acc[i][3] = fmaf(a, w.w, acc[i][3]);
        }
      }
    }
    __syncthreads();
  }

  const int o = n0 + lane * 4;
  if (o >= cout) return;
  float bv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) bv[j] = (!BWD && o + j < cout) ? bias[o + j] : 0.f;
  const bool vec = ((cout & 3) == 0) && (o + 3 < cout);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int l = l0 + warp * 8 + i;
    if (l >= L) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bv[j];
      if (!BWD && apply_elu) v[j] = elu1(v[j]);
    }
    float* dst = out + ((size_t)b * L + l) * cout + o;
    if (vec) {
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      for (int j = 0; j < 4 && o + j < cout; ++j) dst[j] = v[j];
    }
  }
}


// ---------------------------------------------------------------------------------------
// Backward of one Conv1d(K, pad K/2) + ELU layer (training).
//   g        = dy * ELU'(z)                                   (fused into both kernels' loads)
//   dx[l,c]  = sum_o sum_t g[l - t + K/2, o] W[o,c,t]          = conv(g) with W'[c][o][t'] = W[o][c][K-1-t']
//   dW[o,c,t]= sum_{b,l} g[b,l,o] x[b,l+t-K/2,c]   db[o] = sum_{b,l} g[b,l,o]
// ---------------------------------------------------------------------------------------
__global__ void pack_conv_bwd_f32_kernel(const float* __restrict__ w, float* __restrict__ packed, int cin, int cout, int k,
                                         int cin_pad) {
  // packed as a forward weight of a (Cout -> Cin) layer: [o][t'][c_pad]
  size_t n = (size_t)cout * k * cin_pad;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    int c = (int)(idx % cin_pad);
    int t = (int)((idx / cin_pad) % k);
    int o = (int)(idx / ((size_t)cin_pad * k));
    packed[idx] = (c < cin) ? w[((size_t)o * cin + c) * k + (k - 1 - t)] : 0.f;
  }
}

constexpr int WG_TY = 26, WG_TX = 13;   // 26 x 13 threads, 4 x 8 accumulators each: up to 104 x 104 (Cout x Cin) per tap
constexpr int WG_R = 50;                // rows staged per step (block_len 100 = 2 full steps)
constexpr int WG_LD = 104 + 4;          // padded row stride of the staged tiles (floats)

__global__ void __launch_bounds__(WG_TY * WG_TX, 2)
conv1d_wgrad_f32_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ dy,
                        float* __restrict__ dw, float* __restrict__ db, int B, int L, int cin, int cout, int K, int apply_elu) {
  __shared__ __align__(16) float G_s[WG_R][WG_LD];
  __shared__ __align__(16) float X_s[WG_R][WG_LD];
  const int t = blockIdx.y, pad = K / 2;
  const int o0 = blockIdx.z / ((cin + 103) / 104) * 104, c0 = blockIdx.z % ((cin + 103) / 104) * 104;
  const int ty = threadIdx.x / WG_TX, tx = threadIdx.x % WG_TX;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float gsum[4] = {0.f, 0.f, 0.f, 0.f};

  const int chunks_per_cw = (L + WG_R - 1) / WG_R;
  const long long n_chunks = (long long)B * chunks_per_cw;
  for (long long ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const int b = (int)(ch / chunks_per_cw), l0 = (int)(ch % chunks_per_cw) * WG_R;
    // stage 48 rows of g = dy * ELU'(z) and of the tap-shifted input: all global loads of a thread are issued before the
    // first use (15 x 3 loads in flight), otherwise this phase is latency-bound and dominates the kernel
    constexpr int NT = WG_TY * WG_TX, NIT = (WG_R * 104 + NT - 1) / NT;
    float vd[NIT], vy[NIT], vx[NIT];
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      const int i = threadIdx.x + j * NT;
      const int r = i / 104, cc = i % 104, l = l0 + r;
      const int o = o0 + cc, ls = l + t - pad, c = c0 + cc;
      const bool okg = (i < WG_R * 104) && (l < L) && (o < cout);
      const bool okx = (i < WG_R * 104) && (l < L) && (ls >= 0) && (ls < L) && (c < cin);
      vd[j] = okg ? __ldg(dy + ((size_t)b * L + l) * cout + o) : 0.f;
      vy[j] = (okg && apply_elu) ? __ldg(y + ((size_t)b * L + l) * cout + o) : 1.f;
      vx[j] = okx ? __ldg(x + ((size_t)b * L + ls) * cin + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NIT; ++j) {
      const int i = threadIdx.x + j * NT;
      if (i < WG_R * 104) {
        const int r = i / 104, cc = i % 104;
        G_s[r][cc] = vd[j] * ((vy[j] > 0.f) ? 1.f : (vy[j] + 1.f));
        X_s[r][cc] = vx[j];
      }
    }
    __syncthreads();
#pragma unroll 4
    for (int r = 0; r < WG_R; ++r) {
      const float4 g0 = *reinterpret_cast<const float4*>(&G_s[r][4 * ty]);
      // a thread's 8 input channels are two groups of 4, 52 apart: consecutive lanes read consecutive 16-byte words
      const float4 x0 = *reinterpret_cast<const float4*>(&X_s[r][4 * tx]), x1 = *reinterpret_cast<const float4*>(&X_s[r][52 + 4 * tx]);
      const float gv[4] = {g0.x, g0.y, g0.z, g0.w};
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(gv[i], xv[j], acc[i][j]);
        if (tx == 0) gsum[i] += gv[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + 4 * ty + i;
    if (o >= cout) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + (j < 4 ? 4 * tx + j : 52 + 4 * tx + (j - 4));
      if (c < cin) atomicAdd(dw + ((size_t)o * cin + c) * K + t, acc[i][j]);
    }
    if (tx == 0 && t == pad && c0 == 0 && db) atomicAdd(db + o, gsum[i]);
  }
}

// ---------------------------------------------------------------------------------------
// Interleaver / DeInterleaver gather (reference interleavers.py:15-21, 43-48).
// ---------------------------------------------------------------------------------------
__global__ void interleave_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                      const int32_t* __restrict__ perm, size_t n, int L, int F) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(idx % F);
    const size_t bi = idx / F;
    const int i = (int)(bi % L);
    const size_t b = bi / L;
    out[idx] = in[(b * L + perm[i]) * F + f];
  }
}

// ---------------------------------------------------------------------------------------
// Decoder glue.  x_in of a stack is (B, L, 2+F) = [sys, parity, prior_0..F-1] (decoders.py:230,240).
// ---------------------------------------------------------------------------------------
__global__ void dec_init_kernel(const float* __restrict__ received, float* __restrict__ xin, size_t n_rows, int F) {
  const int C = 2 + F;
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n_rows; g += (size_t)gridDim.x * blockDim.x) {
    xin[g * C + 0] = received[g * 3 + 0];
    xin[g * C + 1] = received[g * 3 + 1];
    for (int q = 0; q < F; ++q) xin[g * C + 2 + q] = 0.f;   // prior = zeros (decoders.py:227)
  }
}

// One warp per OUTPUT row (b, i); source row s = map[i] (perm: interleave, inv_perm: de-interleave).
//   lin[q]  = lb[q] + sum_o lw[q,o] * h[b,s,o]                         (dec*_outputs Linear)
//   middle:  ext[q] = lin[q] - x_in_prev[b,s,2+q]   (extrinsic, decoders.py:235-236, 246-247)
//            x_in_next[b,i,:] = [sys, parity, ext]  (interleave/deinterleave + cat, :238-240, :249 + :230)
//   final :  out[b,i] = sigmoid(lin[0])             (decoders.py:267)
__global__ void dec_tail_kernel(const float* __restrict__ h, const float* __restrict__ lw,
                                const float* __restrict__ lb, const float* __restrict__ xin_prev,
                                const float* __restrict__ received, const int32_t* __restrict__ map,
                                float* __restrict__ xin_next, float* __restrict__ out_final,
                                float* __restrict__ trace, size_t n_rows, int L, int units, int F, int fout,
                                int extrinsic, int to_dec2, int final_mode) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  const int C = 2 + F;
  for (size_t g = warp0; g < n_rows; g += nwarp) {
    const size_t b = g / L;
    const int i = (int)(g % L);
    const int s = map[i];
    const size_t src = b * L + s;
    const float* hrow = h + src * units;
    for (int q = 0; q < fout; ++q) {
      float part = 0.f;
      for (int o = lane; o < units; o += 32) part = fmaf(lw[q * units + o], hrow[o], part);
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
      if (lane == 0) {
        const float lin = part + lb[q];
        if (trace) trace[src * F + q] = lin;
        if (final_mode) {
          out_final[g] = 1.f / (1.f + expf(-lin));
        } else {
          const float prior = extrinsic ? xin_prev[src * C + 2 + q] : 0.f;
          xin_next[g * C + 2 + q] = lin - prior;
        }
      }
    }
    if (!final_mode && lane == 0) {
      xin_next[g * C + 0] = received[(to_dec2 ? src : g) * 3 + 0];   // r_sys_int (decoders.py:222) or r_sys
      xin_next[g * C + 1] = received[g * 3 + (to_dec2 ? 2 : 1)];     // r_par2 or r_par1
    }
  }
}

// ---------------------------------------------------------------------------------------
// Encoder glue.
// ---------------------------------------------------------------------------------------
__global__ void enc_prep_kernel(const float* __restrict__ u, const int32_t* __restrict__ perm, float* __restrict__ x,
                                float* __restrict__ x_int, size_t n_rows, int L) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n_rows; g += (size_t)gridDim.x * blockDim.x) {
    const size_t b = g / L;
    const int i = (int)(g % L);
    x[g] = 2.0f * u[g] - 1.0f;                               // encoders.py:362
    x_int[g] = 2.0f * u[b * L + perm[i]] - 1.0f;             // encoders.py:369
  }
}

// warp per row: x_tx[g, branch] = ELU(lb + <lw, h[g,:]>) (encoders.py:364,367,371) and the running
// (sum, sum of squares) of everything written, for power_constraint.
__global__ void enc_tail_kernel(const float* __restrict__ h, const float* __restrict__ lw,
                                const float* __restrict__ lb, float* __restrict__ x_tx, size_t n_rows, int units,
                                int branch, double* __restrict__ stats) {
  __shared__ double red[2][32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5;
  const size_t nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  double s1 = 0.0, s2 = 0.0;
  for (size_t g = warp0; g < n_rows; g += nwarp) {
    const float* hrow = h + g * units;
    float part = 0.f;
    for (int o = lane; o < units; o += 32) part = fmaf(lw[o], hrow[o], part);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if (lane == 0) {
      const float v = elu1(part + lb[0]);
      x_tx[g * 3 + branch] = v;
      s1 += (double)v;
      s2 += (double)v * (double)v;
    }
  }
  if (lane == 0) { red[0][wib] = s1; red[1][wib] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += red[0][w]; c += red[1][w]; }
    atomicAdd(stats + 0, a);
    atomicAdd(stats + 1, c);
  }
}

__global__ void add_count_kernel(double* stats, double n) { stats[2] += n; }

// codes = (x - mean) / std with unbiased std over everything counted in stats (encoders.py:107-116).
// quantize_level 0: plain normalisation; 2: sign(clamp(.)) ; q > 2: q uniform levels on [-limit, limit]
// (STEQuantize.forward, reference encoders.py:20-37, applied after the normalisation as in encoders.py:118-120)
// stats != NULL: mean / unbiased std derived from (sum, sum of squares, count), reported in mean_std when given;
// stats == NULL: mean_std holds the (mean, std) to normalise WITH (running statistics, encoders.py:110-114).
__global__ void power_norm_kernel(const float* __restrict__ x, float* __restrict__ codes, size_t n,
                                  const double* __restrict__ stats, float* __restrict__ mean_std, float limit, float q) {
  float meanf, stdf;
  if (stats) {
    const double cnt = stats[2];
    const double mean = stats[0] / cnt;
    const double var = (stats[1] - cnt * mean * mean) / (cnt - 1.0);
    meanf = (float)mean;
    stdf = (float)sqrt(var > 0.0 ? var : 0.0);
    if (mean_std && blockIdx.x == 0 && threadIdx.x == 0) { mean_std[0] = meanf; mean_std[1] = stdf; }
  } else {
    meanf = mean_std[0];
    stdf = mean_std[1];
  }
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    float v = (x[idx] - meanf) / stdf;
    if (q >= 2.f) {
      v = fminf(fmaxf(v, -limit), limit);
      if (q == 2.f) v = (v > 0.f) ? 1.f : (v < 0.f ? -1.f : 0.f);                               // torch.sign
      else v = rintf((v + limit) * ((q - 1.f) / (2.f * limit))) * (2.f * limit) / (q - 1.f) - limit;
    }
    codes[idx] = v;
  }
}

// ---- power_constraint under autograd (trainer.py:74 through encoders.py:107-116) ---------------------------------------------
// Block-wide sum of two doubles per thread, added into out[0], out[1] by thread 0 (256 threads).
__device__ __forceinline__ void block_add2(double a, double b, double* __restrict__ out) {
  __shared__ double red[2][8];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, d);
    b += __shfl_xor_sync(0xffffffffu, b, d);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = a; red[1][threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double sa = 0.0, sb = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { sa += red[0][w]; sb += red[1][w]; }
    atomicAdd(out + 0, sa);
    atomicAdd(out + 1, sb);
  }
}

// y == NULL: (sum x, sum x^2) -- the forward statistics of an x_tx that did not come out of one of the encoder kernels;
// y != NULL: (sum g, sum g*y) -- the two sums of the backward.
__global__ void __launch_bounds__(256) power_sums_kernel(const float* __restrict__ a, const float* __restrict__ y, size_t n,
                                                         double* __restrict__ out) {
  double s1 = 0.0, s2 = 0.0;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const double v = (double)a[idx];
    s1 += v;
    s2 += v * (y ? (double)y[idx] : v);
  }
  block_add2(s1, s2, out);
}

// y = (x - mean) / std  =>  dx = (g - sum(g) / N - y * sum(g*y) / (N - 1)) / std     (N = stats[2], std = mean_std[1])
__global__ void power_norm_bwd_kernel(const float* __restrict__ g, const float* __restrict__ y, float* __restrict__ dx, size_t n,
                                      const double* __restrict__ sums, const double* __restrict__ stats,
                                      const float* __restrict__ mean_std) {
  const double cnt = stats[2];
  const float mg = (float)(sums[0] / cnt), cy = (float)(sums[1] / (cnt - 1.0)), stdf = mean_std[1];
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
    dx[idx] = (g[idx] - mg - y[idx] * cy) / stdf;
}

// ---- glue of the decoder's / encoder's backward around the fused kernels (loss.backward() of trainer.py:74) -----------------
// out = sigmoid(deinterleave(o_last)) (decoders.py:263-267)  =>  d o_last[b, i] = (d_out * out * (1 - out))[b, perm[i]]
__global__ void dec_out_bwd_kernel(const float* __restrict__ d_out, const float* __restrict__ out, const int32_t* __restrict__ perm,
                                   float* __restrict__ d_last, size_t n, int L) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t src = (idx / L) * L + perm[idx % L];
    const float y = out[src];
    d_last[idx] = d_out[src] * y * (1.f - y);
  }
}

// d received (B, L, 3) from the gradients w.r.t. the stack inputs, dxin_all (2I, B, L, 8): even stacks read [r_sys, r_par1, ...]
// (decoders.py:230), odd stacks [interleave(r_sys), r_par2, ...] (:222, :240), so
//   d r_sys[b, l] = sum_even dxin[s, b, l, 0] + sum_odd dxin[s, b, inv_perm[l], 0],  d r_par1 = sum_even [.., 1],  d r_par2 = sum_odd [.., 1]
__global__ void dec_input_grad_kernel(const float* __restrict__ dxin, const int32_t* __restrict__ inv_perm, float* __restrict__ d_rec,
                                      size_t n_rows, int L, int n_stacks) {
  for (size_t g = blockIdx.x * (size_t)blockDim.x + threadIdx.x; g < n_rows; g += (size_t)gridDim.x * blockDim.x) {
    const size_t gi = (g / L) * L + inv_perm[g % L];
    float sys = 0.f, p1 = 0.f, sys_i = 0.f, p2 = 0.f;
    for (int st = 0; st < n_stacks; st += 2) {
      const float2 e = *reinterpret_cast<const float2*>(dxin + ((size_t)st * n_rows + g) * 8);
      sys += e.x;
      p1 += e.y;
      if (st + 1 < n_stacks) {
        sys_i += dxin[((size_t)(st + 1) * n_rows + gi) * 8];
        p2 += dxin[((size_t)(st + 1) * n_rows + g) * 8 + 1];
      }
    }
    d_rec[g * 3 + 0] = sys + sys_i;
    d_rec[g * 3 + 1] = p1;
    d_rec[g * 3 + 2] = p2;
  }
}

// x_tx = ELU(Linear(h)) (encoders.py:364-371)  =>  dlin[branch, b, l] = d x_tx[b, l, branch] * (x_tx > 0 ? 1 : x_tx + 1)
__global__ void enc_out_bwd_kernel(const float* __restrict__ d_x, const float* __restrict__ x_tx, float* __restrict__ dlin, size_t n_rows) {
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < 3 * n_rows; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t br = idx / n_rows, g = idx - br * n_rows;
    const float x = x_tx[g * 3 + br];
    dlin[idx] = d_x[g * 3 + br] * (x > 0.f ? 1.f : x + 1.f);
  }
}

inline int grid_for(size_t n, int block, int max_blocks = 148 * 16) {
  size_t g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > (size_t)max_blocks) g = max_blocks;
  return (int)g;
}

template <int K, bool BWD>
int launch_conv_k(const float* in, const float* yfwd, float* out, const float* packed, const float* bias, int B, int L, int cin,
                  int cout, int apply_elu, cudaStream_t s) {
  const int cout_pad = (int)align_up(cout, CONV_TN);
  const int tiles_per_cw = (L + 127) / 128;
  int TM = (L + tiles_per_cw - 1) / tiles_per_cw;
  TM = (int)align_up(TM, 8);
  int AS = TM + 8;
  while (AS % 32 != 4) AS += 4;
  const size_t smem = (size_t)(CONV_CC * K * CONV_TN + CONV_CC * AS) * sizeof(float);
  static DeviceOnce once;          // per instantiation
  {
    int rc = device_once(once, "conv1d_f32_kernel", [](int) -> int {
      cudaError_t e = cudaFuncSetAttribute(conv1d_f32_kernel<K, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(conv1d_f32): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
      return TAE_OK;
    });
    if (rc) return rc;
  }
  if (smem > 96 * 1024) { set_error("conv1d_f32: kernel_size %d needs %zu B shared memory", K, smem); return TAE_EUNSUPPORTED; }
  dim3 grid((unsigned)((size_t)B * tiles_per_cw), (unsigned)(cout_pad / CONV_TN));
  conv1d_f32_kernel<K, BWD><<<grid, 32 * (TM / 8), smem, s>>>(in, yfwd, out, packed, bias, L, cin, cout, cout_pad,
                                                              tiles_per_cw, TM, AS, apply_elu);
  return after_launch("conv1d_f32_kernel");
}

}  // namespace

size_t conv_packed_floats(int cin, int cout, int k) { return (size_t)cin * k * align_up(cout, CONV_TN); }

int launch_pack_conv_f32(const float* w, float* packed, int cin, int cout, int k, cudaStream_t s) {
  const size_t n = conv_packed_floats(cin, cout, k);
  pack_conv_f32_kernel<<<grid_for(n, 256), 256, 0, s>>>(w, packed, cin, cout, k, (int)align_up(cout, CONV_TN));
  return after_launch("pack_conv_f32_kernel");
}

int launch_conv_f32(const float* in, float* out, const float* packed, const float* bias, int B, int L, int cin,
                    int cout, int k, int apply_elu, cudaStream_t s) {
  if (B == 0) return TAE_OK;
  switch (k) {
    case 1: return launch_conv_k<1, false>(in, nullptr, out, packed, bias, B, L, cin, cout, apply_elu, s);
    case 3: return launch_conv_k<3, false>(in, nullptr, out, packed, bias, B, L, cin, cout, apply_elu, s);
    case 5: return launch_conv_k<5, false>(in, nullptr, out, packed, bias, B, L, cin, cout, apply_elu, s);
    case 7: return launch_conv_k<7, false>(in, nullptr, out, packed, bias, B, L, cin, cout, apply_elu, s);
    case 9: return launch_conv_k<9, false>(in, nullptr, out, packed, bias, B, L, cin, cout, apply_elu, s);
    default:
      set_error("kernel_size %d unsupported (odd sizes 1..9 only: SameShapeConv1d pads K/2, cnn_utils.py:16)", k);
      return TAE_EUNSUPPORTED;
  }
}


size_t conv_bwd_packed_floats(int cin, int cout, int k) { return (size_t)cout * k * align_up(cin, CONV_TN); }

int launch_conv_bwd_f32(const float* x, const float* y, const float* dy, const float* w, float* dx, float* dw, float* db, int B, int L,
                        int cin, int cout, int k, int apply_elu, float* packed_ws, cudaStream_t s) {
  if (B == 0) return TAE_OK;
  if (dw) {
    const int tiles = ((cout + 103) / 104) * ((cin + 103) / 104);
    const long long n_chunks = (long long)B * ((L + WG_R - 1) / WG_R);
    dim3 grid((unsigned)std::min<long long>(n_chunks, 148 * 2), (unsigned)k, (unsigned)tiles);
    conv1d_wgrad_f32_kernel<<<grid, WG_TY * WG_TX, 0, s>>>(x, y, dy, dw, db, B, L, cin, cout, k, apply_elu);
    int rc = after_launch("conv1d_wgrad_f32_kernel");
    if (rc) return rc;
  }
  if (dx) {
    const int cin_pad = (int)align_up(cin, CONV_TN);
    const size_t n = (size_t)cout * k * cin_pad;
    pack_conv_bwd_f32_kernel<<<grid_for(n, 256), 256, 0, s>>>(w, packed_ws, cin, cout, k, cin_pad);
    int rc = after_launch("pack_conv_bwd_f32_kernel");
    if (rc) return rc;
    switch (k) {
      case 1: return launch_conv_k<1, true>(dy, y, dx, packed_ws, nullptr, B, L, cout, cin, apply_elu, s);
      case 3: return launch_conv_k<3, true>(dy, y, dx, packed_ws, nullptr, B, L, cout, cin, apply_elu, s);
      case 5: return launch_conv_k<5, true>(dy, y, dx, packed_ws, nullptr, B, L, cout, cin, apply_elu, s);
      case 7: return launch_conv_k<7, true>(dy, y, dx, packed_ws, nullptr, B, L, cout, cin, apply_elu, s);
      case 9: return launch_conv_k<9, true>(dy, y, dx, packed_ws, nullptr, B, L, cout, cin, apply_elu, s);
      default: set_error("kernel_size %d unsupported", k); return TAE_EUNSUPPORTED;
    }
  }
  return TAE_OK;
}

int launch_interleave_f32(const float* in, float* out, const int32_t* perm, int B, int L, int F, cudaStream_t s) {
  const size_t n = (size_t)B * L * F;
  if (n == 0) return TAE_OK;
  interleave_f32_kernel<<<grid_for(n, 256, 148 * 32), 256, 0, s>>>(in, out, perm, n, L, F);
  return after_launch("interleave_f32_kernel");
}

// ---------------------------------------------------------------------------------------
// DEC_LargeCNN.forward, fp32 (reference decoders.py:219-269)
// ---------------------------------------------------------------------------------------
static constexpr int DEC_F32_CHUNK = 8192;  // codewords per pass: bounds the activation workspace

static size_t dec_packed_floats_f32(const TaeDecConfig& c) {
  size_t n = 0;
  for (int j = 0; j < c.num_layer; ++j)
    n += conv_packed_floats(j == 0 ? 2 + c.num_iter_ft : c.num_unit, c.num_unit, c.kernel_size);
  return n * 2 * c.num_iteration;
}

size_t dec_workspace_bytes_f32(const TaeDecConfig& c, int B) {
  const size_t chunk = (size_t)std::min(B, DEC_F32_CHUNK);
  const size_t rows = chunk * c.block_len;
  size_t fl = align_up(dec_packed_floats_f32(c), 64);
  fl += 2 * align_up(rows * (2 + c.num_iter_ft), 64);
  fl += 2 * align_up(rows * c.num_unit, 64);
  return fl * sizeof(float) + 256;
}

int dec_forward_f32(const TaeDecConfig& c, const float* params, const float* received, const int32_t* perm,
                    const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes,
                    cudaStream_t s) {
  if (ws_bytes < dec_workspace_bytes_f32(c, B)) {
    set_error("tae_dec_forward(fp32): workspace %zu < %zu bytes", ws_bytes, dec_workspace_bytes_f32(c, B));
    return TAE_EWORKSPACE;
  }
  const int L = c.block_len, F = c.num_iter_ft, U = c.num_unit, K = c.kernel_size, I = c.num_iteration;
  DecStackLayout lay[64];
  dec_layout(c, lay);

  float* base = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  const size_t chunk = (size_t)std::min(B, DEC_F32_CHUNK);
  const size_t rows_max = chunk * L;
  float* wpk = base;
  float* xin[2];
  xin[0] = wpk + align_up(dec_packed_floats_f32(c), 64);
  xin[1] = xin[0] + align_up(rows_max * (2 + F), 64);
  float* hbuf[2];
  hbuf[0] = xin[1] + align_up(rows_max * (2 + F), 64);
  hbuf[1] = hbuf[0] + align_up(rows_max * U, 64);

  // re-layout every conv weight once per call (2.4 M floats: negligible next to the convs)
  const float* wp_ptr[64][16];
  {
    float* p = wpk;
    for (int st = 0; st < 2 * I; ++st)
      for (int j = 0; j < c.num_layer; ++j) {
        const ConvLayer& cl = lay[st].conv[j];
        int rc = launch_pack_conv_f32(params + cl.w_off, p, cl.cin, cl.cout, K, s);
        if (rc) return rc;
        wp_ptr[st][j] = p;
        p += conv_packed_floats(cl.cin, cl.cout, K);
      }
  }

  for (size_t b0 = 0; b0 < (size_t)B; b0 += chunk) {
    const int nb = (int)std::min(chunk, (size_t)B - b0);
    const size_t rows = (size_t)nb * L;
    const float* rec = received + b0 * L * 3;
    dec_init_kernel<<<grid_for(rows, 256), 256, 0, s>>>(rec, xin[0], rows, F);
    int rc = after_launch("dec_init_kernel");
    if (rc) return rc;
    for (int st = 0; st < 2 * I; ++st) {
      const int sidx = st & 1;           // 0: dec1 stack (natural order), 1: dec2 stack (interleaved order)
      const float* cur = xin[sidx];
      int hb = 0;
      for (int j = 0; j < c.num_layer; ++j) {
        const ConvLayer& cl = lay[st].conv[j];
        rc = launch_conv_f32(cur, hbuf[hb], wp_ptr[st][j], params + cl.b_off, nb, L, cl.cin, cl.cout, K, 1, s);
        if (rc) return rc;
        cur = hbuf[hb];
        hb ^= 1;
      }
      const bool final_mode = (st == 2 * I - 1);
      float* tr = trace ? trace + ((size_t)st * B + b0) * L * F : nullptr;
      const int threads = 256;
      const int blocks = grid_for(rows * 32, threads);
      dec_tail_kernel<<<blocks, threads, 0, s>>>(cur, params + lay[st].lin_w_off, params + lay[st].lin_b_off,
                                                 xin[sidx], rec, sidx == 0 ? perm : inv_perm, xin[sidx ^ 1],
                                                 out + b0 * L, tr, rows, L, U, F, lay[st].fout, c.extrinsic,
                                                 sidx == 0 ? 1 : 0, final_mode ? 1 : 0);
      rc = after_launch("dec_tail_kernel");
      if (rc) return rc;
    }
  }
  return TAE_OK;
}

// ---------------------------------------------------------------------------------------
// ENC_interCNN.forward up to the concat, fp32 (reference encoders.py:362-373)
// ---------------------------------------------------------------------------------------
static constexpr int ENC_F32_CHUNK = 8192;

static size_t enc_packed_floats_f32(const TaeEncConfig& c) {
  size_t n = 0;
  for (int j = 0; j < c.num_layer; ++j) n += conv_packed_floats(j == 0 ? 1 : c.num_unit, c.num_unit, c.kernel_size);
  return n * 3;
}

size_t enc_workspace_bytes_f32(const TaeEncConfig& c, int B) {
  const size_t rows = (size_t)std::min(B, ENC_F32_CHUNK) * c.block_len;
  size_t fl = align_up(enc_packed_floats_f32(c), 64);
  fl += 2 * align_up(rows, 64);
  fl += 2 * align_up(rows * c.num_unit, 64);
  return fl * sizeof(float) + 256;
}

int enc_forward_f32(const TaeEncConfig& c, const float* params, const float* u, const int32_t* perm, float* x_tx,
                    double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < enc_workspace_bytes_f32(c, B)) {
    set_error("tae_enc_forward: workspace %zu < %zu bytes", ws_bytes, enc_workspace_bytes_f32(c, B));
    return TAE_EWORKSPACE;
  }
  const int L = c.block_len, U = c.num_unit, K = c.kernel_size;
  EncBranchLayout lay[3];
  enc_layout(c, lay);
  float* base = reinterpret_cast<float*>(align_up(reinterpret_cast<uintptr_t>(ws), 256));
  const size_t chunk = (size_t)std::min(B, ENC_F32_CHUNK);
  const size_t rows_max = chunk * L;
  float* wpk = base;
  float* x = wpk + align_up(enc_packed_floats_f32(c), 64);
  float* x_int = x + align_up(rows_max, 64);
  float* hbuf[2];
  hbuf[0] = x_int + align_up(rows_max, 64);
  hbuf[1] = hbuf[0] + align_up(rows_max * U, 64);

  const float* wp_ptr[3][16];
  {
    float* p = wpk;
    for (int br = 0; br < 3; ++br)
      for (int j = 0; j < c.num_layer; ++j) {
        const ConvLayer& cl = lay[br].conv[j];
        int rc = launch_pack_conv_f32(params + cl.w_off, p, cl.cin, cl.cout, K, s);
        if (rc) return rc;
        wp_ptr[br][j] = p;
        p += conv_packed_floats(cl.cin, cl.cout, K);
      }
  }
  for (size_t b0 = 0; b0 < (size_t)B; b0 += chunk) {
    const int nb = (int)std::min(chunk, (size_t)B - b0);
    const size_t rows = (size_t)nb * L;
    enc_prep_kernel<<<grid_for(rows, 256), 256, 0, s>>>(u + b0 * L, perm, x, x_int, rows, L);
    int rc = after_launch("enc_prep_kernel");
    if (rc) return rc;
    for (int br = 0; br < 3; ++br) {
      const float* cur = (br == 2) ? x_int : x;
      int hb = 0;
      for (int j = 0; j < c.num_layer; ++j) {
        const ConvLayer& cl = lay[br].conv[j];
        rc = launch_conv_f32(cur, hbuf[hb], wp_ptr[br][j], params + cl.b_off, nb, L, cl.cin, cl.cout, K, 1, s);
        if (rc) return rc;
        cur = hbuf[hb];
        hb ^= 1;
      }
      enc_tail_kernel<<<grid_for(rows * 32, 256, 148 * 8), 256, 0, s>>>(cur, params + lay[br].lin_w_off,
                                                                         params + lay[br].lin_b_off,
                                                                         x_tx + b0 * L * 3, rows, U, br, stats);
      rc = after_launch("enc_tail_kernel");
      if (rc) return rc;
    }
  }
  add_count_kernel<<<1, 1, 0, s>>>(stats, (double)B * L * 3);
  return after_launch("add_count_kernel");
}

int launch_add_count(double* stats, double n, cudaStream_t s) {
  add_count_kernel<<<1, 1, 0, s>>>(stats, n);
  return after_launch("add_count_kernel");
}

int launch_power_norm_f32(const float* x, float* codes, size_t n, const double* stats, float* mean_std, float limit, float q,
                          cudaStream_t s) {
  if (n == 0) return TAE_OK;
  power_norm_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, codes, n, stats, mean_std, limit, q);
  return after_launch("power_norm_kernel");
}

int launch_dec_out_bwd_f32(const float* d_out, const float* out, const int32_t* perm, float* d_last, int B, int L, cudaStream_t s) {
  const size_t n = (size_t)B * L;
  if (n == 0) return TAE_OK;
  dec_out_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(d_out, out, perm, d_last, n, L);
  return after_launch("dec_out_bwd_kernel");
}

int launch_dec_input_grad_f32(const float* dxin_all, const int32_t* inv_perm, float* d_received, int n_stacks, int B, int L, cudaStream_t s) {
  const size_t n = (size_t)B * L;
  if (n == 0) return TAE_OK;
  dec_input_grad_kernel<<<grid_for(n, 256), 256, 0, s>>>(dxin_all, inv_perm, d_received, n, L, n_stacks);
  return after_launch("dec_input_grad_kernel");
}

int launch_enc_out_bwd_f32(const float* d_x, const float* x_tx, float* dlin, int B, int L, cudaStream_t s) {
  const size_t n = (size_t)B * L;
  if (n == 0) return TAE_OK;
  enc_out_bwd_kernel<<<grid_for(3 * n, 256), 256, 0, s>>>(d_x, x_tx, dlin, n);
  return after_launch("enc_out_bwd_kernel");
}

int launch_power_sums_f32(const float* a, const float* y, size_t n, double* out, cudaStream_t s) {
  if (n == 0) return TAE_OK;
  power_sums_kernel<<<grid_for(n, 256 * 8, 148 * 2), 256, 0, s>>>(a, y, n, out);
  return after_launch("power_sums_kernel");
}

int launch_power_norm_bwd_f32(const float* g, const float* y, float* dx, size_t n, const double* sums, const double* stats,
                              const float* mean_std, cudaStream_t s) {
  if (n == 0) return TAE_OK;
  power_norm_bwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(g, y, dx, n, sums, stats, mean_std);
  return after_launch("power_norm_bwd_kernel");
}

}  // namespace tae
