// DEC_LargeRNN recurrence on 5th-gen tensor cores (BASELINE config 5), sm_100a only.
//
// Reference arithmetic restated: torch.nn.GRU as the reference builds it (decoders.py:43-52: 2 layers, bidirectional,
// batch_first), one direction of one layer per launch:
//     r = sigmoid(W_ir x + b_ir + W_hr h + b_hr)      z = sigmoid(W_iz x + b_iz + W_hz h + b_hz)
//     n = tanh(W_in x + b_in + r * (W_hn h + b_hn))   h' = (1 - z) * n + z * h
//
// Execution model
//   * cluster of 2 CTAs, persistent; each CTA owns 128 codewords = the 128 rows of an MMA tile.  Per time step the pair
//     issues tcgen05.mma.cta_group::2 (M = 256) chains into three accumulators in TMEM:
//         D_rz (N = 224: r | z)  = [x_t | 1 | h_{t-1}] . [W_i{r,z} | b | W_h{r,z}]^T
//         D_nx (N = 112)         = [x_t | 1] . [W_in | b_in]^T          D_nh (N = 112) = [h_{t-1} | 1] . [W_hn | b_hn]^T
//     The input projection is part of the chain (K = in + H), so x is read once as bf16 and no (B, L, 3H) projection
//     tensor exists.  Biases enter through a constant-one chunk (bf16 hi + lo split), no epilogue add.
//   * all weights of the layer-direction stay resident in shared memory (each CTA stages its half of the N columns);
//     h_{t-1} lives in shared memory as the bf16 A operand (canonical K-major [H/8][128 rows][8]) and, as the fp32
//     state, in the REGISTERS of the epilogue thread that owns (codeword, unit slice) for the whole sequence.
//   * 16 epilogue warps (4 TMEM lane quadrants x 4 unit slices): tcgen05.ld -> gates (r, z through one packed
//     tanh.approx.f16x2, n through tanh.approx.f32) -> h' -> bf16 -> shared memory (next step's operand) + HBM.
//   * activations between launches travel as TIME-MAJOR TILES in HBM: [block of R codewords][t][chunk][R][8] bf16, i.e.
//     the A-operand chunks of one time step are contiguous: one loader thread fetches x_{t+1} with bulk copies
//     (cp.async.bulk, no register staging) while the epilogue of step t runs, and the epilogue's stores are coalesced.
//     R = rows per CTA (32..128, a multiple of 32): small batches use fewer rows per CTA so that all SMs work.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>

#include "tae_common.cuh"
#include "tae_umma.cuh"

namespace tae {

namespace {

constexpr int G_ROWS = 128;
constexpr uint32_t G_CHUNK_B = G_ROWS * 16;          // one 8-channel chunk of the A operand
constexpr int G_HCH = 13;                            // hidden chunks (H <= 104)
constexpr int G_HKS = 7;                             // k-steps of the hidden part: (h0,h1) .. (h10,h11), (h12, ones)
constexpr int G_XCH_MAX = 26;                        // input chunks (two directions of 13 chunks)
// The hidden units are processed in two halves, U0 = units [0, 64) and U1 = units [64, 112), each with its own accumulators
// and its own commit: the epilogue of U0 overlaps the MMAs of U1 (the epilogue is MUFU-bound, the MMAs use the tensor pipe).
constexpr int G_NU0 = 64, G_NU1 = 48;                // units per half (UMMA N: 2*NU for r|z, NU for n; all multiples of 16)
constexpr int G_NG = G_NU0 + G_NU1;                  // 112 columns per gate
constexpr int G_CH0 = G_NU0 / 8;                     // chunks of U0 (8); U1 = chunks 8..12
constexpr uint32_t G_RZ_KS_B = 2 * G_NG * 16;        // bytes of one k-step of W_rz per CTA, both halves (2 chunks x 112 columns x 16 B)
constexpr uint32_t G_N_KS_B = 2 * (G_NG / 2) * 16;   // bytes of one k-step of W_nx / W_nh per CTA, both halves (56 columns)
constexpr uint32_t G_TM_RZ0 = 0, G_TM_NX0 = 128, G_TM_NH0 = 192, G_TM_RZ1 = 256, G_TM_NX1 = 352, G_TM_NH1 = 400;
constexpr int G_EPI_WARP0 = 4, G_EPI_WARPS = 16, G_LOAD_WARP = 1;
constexpr int G_THREADS = 32 * (G_EPI_WARP0 + G_EPI_WARPS);      // 640
constexpr int G_RDY_COUNT = 2 * (G_EPI_WARPS + 1);

struct GruGeom {
  int n_xch, xks;                  // input chunks, k-steps of the input part (the last one pairs with the ones chunk)
  uint32_t off_x, off_ones, off_w, off_bars, total;
  uint32_t w_rz_b, w_nx_b, w_nh_b; // bytes per CTA half
};
__host__ __device__ inline GruGeom gru_geom(int n_xch) {
  GruGeom g{};
  g.n_xch = n_xch;
  g.xks = (g.n_xch + 2) / 2;                         // chunks x0..x_{n-1}, ones (+ a repeat of ones with zero weights if needed)
  g.off_x = 2 * G_HCH * G_CHUNK_B;                   // two hidden-state buffers (step parity): see the epilogue
  g.off_ones = g.off_x + (uint32_t)g.n_xch * G_CHUNK_B;
  g.off_w = g.off_ones + G_CHUNK_B;
  g.w_rz_b = (uint32_t)(g.xks + G_HKS) * G_RZ_KS_B;
  g.w_nx_b = (uint32_t)g.xks * G_N_KS_B;
  g.w_nh_b = (uint32_t)G_HKS * G_N_KS_B;
  g.off_bars = g.off_w + g.w_rz_b + g.w_nx_b + g.w_nh_b;
  g.total = g.off_bars + 64;
  return g;
}

struct GruArgs {
  const uint8_t* wimg;             // [2 halves][W_rz | W_nx | W_nh]
  const uint8_t* x;                // input tiles  [block][L][n_xch][R][8] bf16
  uint8_t* out;                    // output tiles [block][L][out_chunks][R][8] bf16; this direction writes chunks [out_c0, out_c0 + 13)
  int* err;
  int B, L, H, hch, n_xch, R, out_chunks, out_c0, reverse, n_pairs;
  long long* tl;                   // optional timeline (tae_debug_gru_timeline): clock64 stamps of cluster 0, steps 0..63, 8 per step
};

// ---- weight image -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float g_bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// element (k-step of the input part, chunk 0/1, e8) of gate row `row` (0 .. 3H-1)
// The input tile holds `in_groups` groups of grp_pad channels of which grp_valid are real (layer 0: one group of 2 + F in 8;
// layer 1: the two directions of the layer below, H of 104 each): padded channel -> column of weight_ih, or -1.
struct InMap { int in_ch, grp_valid, grp_pad; };
__device__ __forceinline__ int in_col(const InMap& m, int c_pad) {
  const int gidx = c_pad / m.grp_pad, w = c_pad - gidx * m.grp_pad;
  const int c = gidx * m.grp_valid + w;
  return (w < m.grp_valid && c < m.in_ch) ? c : -1;
}
__device__ __forceinline__ float gru_x_elem(const float* w_ih, const float* bias, const InMap& im, int n_xch, int row, bool valid, int ks,
                                            int ch, int e8) {
  if (!valid) return 0.f;
  const int xc = 2 * ks + ch;
  if (xc < n_xch) {
    const int c = in_col(im, 8 * xc + e8);
    return c >= 0 ? w_ih[(size_t)row * im.in_ch + c] : 0.f;
  }
  if (xc == n_xch) {                                  // the ones chunk: bias as bf16 hi + lo
    if (e8 == 0) return g_bf16_round(bias[0]);
    if (e8 == 1) return bias[0] - g_bf16_round(bias[0]);
  }
  return 0.f;
}
__device__ __forceinline__ float gru_h_elem(const float* w_hh, const float* bias, int H, int row, bool valid, int hk, int ch, int e8) {
  if (!valid) return 0.f;
  const int hc = 2 * hk + ch;
  if (hc < G_HCH) {
    const int c = 8 * hc + e8;
    return c < H ? w_hh[(size_t)row * H + c] : 0.f;
  }
  if (bias) {                                         // (h12, ones)
    if (e8 == 0) return g_bf16_round(bias[0]);
    if (e8 == 1) return bias[0] - g_bf16_round(bias[0]);
  }
  return 0.f;
}

__global__ void gru_pack_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                                const float* __restrict__ b_hh, __nv_bfloat16* __restrict__ img, int H, const InMap im, int n_xch) {
  // per CTA half: [RZ of U0 | RZ of U1 | NX of U0 | NX of U1 | NH of U0 | NH of U1]; every section is [k-step][2 chunks][columns][8]
  const GruGeom g = gru_geom(n_xch);
  const uint32_t half_elems = (g.w_rz_b + g.w_nx_b + g.w_nh_b) / 2;
  const int ks_rz = g.xks + G_HKS;
  const uint32_t sec[6] = {(uint32_t)ks_rz * 2 * G_NU0 * 8,       (uint32_t)ks_rz * 2 * G_NU1 * 8,
                           (uint32_t)g.xks * 2 * (G_NU0 / 2) * 8, (uint32_t)g.xks * 2 * (G_NU1 / 2) * 8,
                           (uint32_t)G_HKS * 2 * (G_NU0 / 2) * 8, (uint32_t)G_HKS * 2 * (G_NU1 / 2) * 8};
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 2 * half_elems; idx += gridDim.x * blockDim.x) {
    const int hf = idx / half_elems;
    uint32_t r = idx % half_elems;
    int sc = 0;
    while (r >= sec[sc]) { r -= sec[sc]; ++sc; }
    const int hu = sc & 1, kind = sc >> 1;                 // kind 0: r|z, 1: n (input part), 2: n (hidden part)
    const int NU = hu ? G_NU1 : G_NU0, u0 = hu ? G_NU0 : 0;
    const int cols = kind == 0 ? NU : NU / 2;              // columns this CTA stages
    const int ks = r / (2 * cols * 8), ch = (r / (cols * 8)) & 1, n = (r / 8) % cols, e8 = r % 8;
    float v;
    if (kind == 0) {
      const int u = u0 + n, row = hf * H + u;              // CTA 0 = r gate, CTA 1 = z gate
      const bool valid = u < H;
      if (ks < g.xks) {
        const float bsum = valid ? b_ih[row] + b_hh[row] : 0.f;
        v = gru_x_elem(w_ih, &bsum, im, g.n_xch, row, valid, ks, ch, e8);
      } else {
        v = gru_h_elem(w_hh, nullptr, H, row, valid, ks - g.xks, ch, e8);
      }
    } else {
      const int u = u0 + hf * (NU / 2) + n;
      const bool valid = u < H;
      if (kind == 1) v = gru_x_elem(w_ih, valid ? b_ih + 2 * H + u : nullptr, im, g.n_xch, 2 * H + u, valid, ks, ch, e8);
      else v = gru_h_elem(w_hh, valid ? b_hh + 2 * H + u : nullptr, H, 2 * H + u, valid, ks, ch, e8);
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

// ---- small device helpers ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (sigmoid(a), sigmoid(b)) through ONE packed tanh: sigmoid(v) = 0.5 + 0.5 tanh(0.5 v)
__device__ __forceinline__ float2 sigmoid2_fast(float a, float b) {
  const __half2 hx = __floats2half2_rn(0.5f * a, 0.5f * b);
  uint32_t in = *reinterpret_cast<const uint32_t*>(&hx), outv;
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(outv) : "r"(in));
  const float2 t = __half22float2(*reinterpret_cast<const __half2*>(&outv));
  return make_float2(fmaf(0.5f, t.x, 0.5f), fmaf(0.5f, t.y, 0.5f));
}
// ---- the recurrence kernel -------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1) gru_pair_kernel(const GruArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const GruGeom g = gru_geom(a.n_xch);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t bar_w = sbase + g.off_bars, bar_rdy = bar_w + 8, bar_acc = bar_w + 16, bar_x = bar_w + 24, tptr = bar_w + 32, bar_acc1 = bar_w + 40,
                 bar_xfree = bar_w + 48;
  const int L = a.L;

  // ---- setup: zero the operand region, ones chunk, barriers, TMEM, weights ------------------------------------------
  for (uint32_t i = threadIdx.x * 16; i < g.off_w; i += G_THREADS * 16) st_shared_v4(sbase + i, 0u, 0u, 0u, 0u);
  __syncthreads();
  for (int r = threadIdx.x; r < G_ROWS; r += G_THREADS) st_shared_v4(sbase + g.off_ones + (uint32_t)r * 16, 0x3F803F80u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    mbar_init(bar_w, 1);
    mbar_init(bar_rdy, rank == 0 ? G_RDY_COUNT : 1);
    mbar_init(bar_acc, 1);
    mbar_init(bar_acc1, 1);
    mbar_init(bar_xfree, 1);
    mbar_init(bar_x, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<2>(tptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  const uint32_t w_bytes = g.w_rz_b + g.w_nx_b + g.w_nh_b;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar_w, w_bytes);
    const uint8_t* src = a.wimg + (size_t)rank * w_bytes;
    for (uint32_t o = 0; o < w_bytes; o += 16384) bulk_g2s(sbase + g.off_w + o, src + o, (w_bytes - o < 16384u) ? (w_bytes - o) : 16384u, bar_w);
  }
  mbar_wait_cluster(bar_w, 0, a.err, 31);
  cluster_sync_all();                       // both CTAs: barriers initialised, weights resident
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr) : "memory");

  const int pair0 = (int)cluster_id_x(), pair_stride = (int)n_clusters_x();

  if (warp == 0) {
    // ================= MMA issuer (leader CTA) ========================================================================
    if (rank == 0) {
      const uint32_t h0 = sbase, x0 = sbase + g.off_x, ones = sbase + g.off_ones;
      const int ks_rz = g.xks + G_HKS;
      // weight sections of this CTA (gru_pack_kernel): [RZ0 | RZ1 | NX0 | NX1 | NH0 | NH1]
      uint32_t w_sec[6];
      {
        uint32_t o = sbase + g.off_w;
        const uint32_t bytes[6] = {(uint32_t)ks_rz * 2 * G_NU0 * 16,       (uint32_t)ks_rz * 2 * G_NU1 * 16,
                                   (uint32_t)g.xks * 2 * (G_NU0 / 2) * 16, (uint32_t)g.xks * 2 * (G_NU1 / 2) * 16,
                                   (uint32_t)G_HKS * 2 * (G_NU0 / 2) * 16, (uint32_t)G_HKS * 2 * (G_NU1 / 2) * 16};
        for (int i = 0; i < 6; ++i) { w_sec[i] = o; o += bytes[i]; }
      }
      uint32_t step = 0;
      for (int pr = pair0; pr < a.n_pairs; pr += pair_stride)
        for (int s = 0; s < L; ++s, ++step) {
          const bool stamp = a.tl && pair0 == 0 && step < 64 && lane == 0;
          if (stamp) a.tl[step * 8 + 0] = clock64();
          mbar_wait(bar_rdy, step & 1u, a.err, 32);
          tc_fence_after();
          if (stamp) a.tl[step * 8 + 1] = clock64();
          if (elect_one()) {
            const uint32_t hbuf = h0 + (step & 1u) * (uint32_t)(G_HCH * G_CHUNK_B);      // h_{t-1}: the buffer of this step's parity
            // per half: input part, hidden part, commit -- so U0's accumulators complete while U1's MMAs still run and the
            // (MUFU-bound) epilogue of U0 overlaps them.  The x chunks are free once U1's input part has been read.
#pragma unroll
            for (int hu = 0; hu < 2; ++hu) {
              const uint32_t NU = hu ? G_NU1 : G_NU0;
              const uint32_t idesc_rz = make_idesc(256, 2 * (int)NU), idesc_n = make_idesc(256, (int)NU);
              const uint32_t d_rz = tmem_base + (hu ? G_TM_RZ1 : G_TM_RZ0), d_nx = tmem_base + (hu ? G_TM_NX1 : G_TM_NX0),
                             d_nh = tmem_base + (hu ? G_TM_NH1 : G_TM_NH0);
              const uint32_t w_rz = w_sec[hu], w_nx = w_sec[2 + hu], w_nh = w_sec[4 + hu];
              const uint32_t rz_ks_b = 2 * NU * 16, n_ks_b = NU * 16;         // bytes per k-step (2 chunks x columns x 16 B)
              for (int ks = 0; ks < g.xks; ++ks) {      // chunk pairs (x0,x1) ..; the last pair ends on the ones chunk
                const int c0 = 2 * ks;
                const uint32_t a_addr = c0 < g.n_xch ? x0 + (uint32_t)c0 * G_CHUNK_B : ones;
                const uint32_t a_lbo = (c0 + 1 < g.n_xch) ? G_CHUNK_B : (c0 < g.n_xch ? ones - a_addr : 0u);
                const uint64_t ad = make_desc(a_addr, a_lbo);
                umma_bf16<2>(d_rz, ad, make_desc(w_rz + (uint32_t)ks * rz_ks_b, NU * 16), idesc_rz, ks > 0);
                umma_bf16<2>(d_nx, ad, make_desc(w_nx + (uint32_t)ks * n_ks_b, (NU / 2) * 16), idesc_n, ks > 0);
              }
              if (hu == 1) umma_commit_pair(bar_xfree, 3);
#pragma unroll
              for (int hk = 0; hk < G_HKS; ++hk) {       // (h0,h1) .. (h10,h11), (h12, ones)
                const uint32_t a_addr = hbuf + (uint32_t)(2 * hk) * G_CHUNK_B;
                const uint64_t ad = make_desc(a_addr, hk < G_HKS - 1 ? G_CHUNK_B : ones - a_addr);
                umma_bf16<2>(d_rz, ad, make_desc(w_rz + (uint32_t)(g.xks + hk) * rz_ks_b, NU * 16), idesc_rz, 1);
                umma_bf16<2>(d_nh, ad, make_desc(w_nh + (uint32_t)hk * n_ks_b, (NU / 2) * 16), idesc_n, hk > 0);
              }
              umma_commit_pair(hu ? bar_acc1 : bar_acc, 3);
            }
          }
          __syncwarp();
          if (stamp) a.tl[step * 8 + 2] = clock64();
        }
    }
  } else if (warp == G_LOAD_WARP) {
    // ================= loader: one thread, bulk copies of the time step's input chunks (R rows each) =====================
    if (lane == 0) {
      const uint32_t chunk_bytes = (uint32_t)a.R * 16u;
      uint32_t step = 0;
      for (int pr = pair0; pr < a.n_pairs; pr += pair_stride) {
        const size_t blk = (size_t)(2 * pr + (int)rank);
        for (int s = 0; s < L; ++s, ++step) {
          const int t = a.reverse ? L - 1 - s : s;
          if (step > 0) mbar_wait(bar_xfree, (step - 1) & 1u, a.err, 33);     // the input-part MMAs of the previous step have read the x chunks
          const bool stamp = a.tl && pair0 == 0 && rank == 0 && step < 64;
          if (stamp) a.tl[step * 8 + 3] = clock64();
          const uint8_t* src = a.x + ((blk * L + t) * a.n_xch) * chunk_bytes;
          mbar_arrive_expect_tx(bar_x, (uint32_t)a.n_xch * chunk_bytes);
          if (a.R == G_ROWS) bulk_g2s(sbase + g.off_x, src, (uint32_t)a.n_xch * chunk_bytes, bar_x);      // a full tile is contiguous on both sides
          else for (int c = 0; c < a.n_xch; ++c) bulk_g2s(sbase + g.off_x + (uint32_t)c * G_CHUNK_B, src + (size_t)c * chunk_bytes, chunk_bytes, bar_x);
          if (s + 1 < L) {            // next step's tile towards L2 while this one lands and the step computes
            const int tn = a.reverse ? t - 1 : t + 1;
            const uint8_t* nsrc = a.x + ((blk * L + tn) * a.n_xch) * chunk_bytes;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(nsrc), "r"((uint32_t)a.n_xch * chunk_bytes) : "memory");
          }
          mbar_wait(bar_x, step & 1u, a.err, 35);
          mbar_arrive_leader(bar_rdy, rank);                                   // written by the async proxy: no fence needed
          if (stamp) a.tl[step * 8 + 4] = clock64();
        }
      }
    }
  } else if (warp >= G_EPI_WARP0) {
    // ================= epilogue: gates, new hidden state =============================================================
    const int ew = warp - G_EPI_WARP0, q = warp & 3, part = ew >> 2;
    // chunks of this thread: U0 -> {part, part + 4}; U1 -> {8 + part} and, for part 0, chunk 12.  Slot i of hp[] <-> chunk_of[i].
    const int chunk_of[4] = {part, part + 4, 8 + part, 12};
    const int n_slots = part == 0 ? 4 : 3;
    const int row = 32 * q + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    uint32_t step = 0;
    for (int pr = pair0; pr < a.n_pairs; pr += pair_stride) {
      const size_t blk = (size_t)(2 * pr + (int)rank);
      const bool ok = row < a.R;                       // rows beyond R are padding of the MMA tile (their x rows stay zero)
      const bool warp_ok = 32 * q < a.R;               // warp-uniform: tcgen05.ld is .aligned
      float hp[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) hp[i] = 0.f;
      // h_0 = 0 in the buffer the first step of this pair reads (its previous readers completed before bar_acc1 of that step)
      for (int i = 0; i < n_slots; ++i)
        st_shared_v4(sbase + (step & 1u) * (uint32_t)(G_HCH * G_CHUNK_B) + (uint32_t)chunk_of[i] * G_CHUNK_B + (uint32_t)row * 16, 0u, 0u, 0u, 0u);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar_rdy, rank);
      for (int s = 0; s < L; ++s, ++step) {
        const int t = a.reverse ? L - 1 - s : s;
        uint8_t* otile = a.out + (((blk * L + t) * a.out_chunks + a.out_c0) * a.R + row) * 16;
        const bool stamp = a.tl && pair0 == 0 && rank == 0 && step < 64 && ew == 0 && lane == 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (i == 0) {
            mbar_wait(bar_acc, step & 1u, a.err, 34);       // accumulators of U0 complete (the MMAs of U1 are still running)
            tc_fence_after();
            if (stamp) a.tl[step * 8 + 5] = clock64();
          }
          if (i == 2) {
            mbar_wait(bar_acc1, step & 1u, a.err, 36);      // accumulators of U1 complete
            tc_fence_after();
            if (stamp) a.tl[step * 8 + 7] = clock64();
          }
          const int cidx = chunk_of[i];
          if (i < n_slots && warp_ok && cidx < a.hch) {
            const bool u1 = i >= 2;
            const uint32_t NU = u1 ? G_NU1 : G_NU0;
            const uint32_t col = (uint32_t)(8 * (cidx - (u1 ? G_CH0 : 0)));
            uint32_t dr[8], dz[8], dx[8], dh[8];
            tmem_ld8(lane_addr + (u1 ? G_TM_RZ1 : G_TM_RZ0) + col, dr);
            tmem_ld8(lane_addr + (u1 ? G_TM_RZ1 : G_TM_RZ0) + NU + col, dz);
            tmem_ld8(lane_addr + (u1 ? G_TM_NX1 : G_TM_NX0) + col, dx);
            tmem_ld8(lane_addr + (u1 ? G_TM_NH1 : G_TM_NH0) + col, dh);
            tmem_ld_wait();
            uint32_t pk[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              float hn[2];
#pragma unroll
              for (int k = 0; k < 2; ++k) {
                const float2 rz = sigmoid2_fast(__uint_as_float(dr[j + k]), __uint_as_float(dz[j + k]));
                const float n = tanh_fast(fmaf(rz.x, __uint_as_float(dh[j + k]), __uint_as_float(dx[j + k])));
                hn[k] = fmaf(rz.y, hp[8 * i + j + k] - n, n);                     // (1 - z) n + z h
                hp[8 * i + j + k] = hn[k];
              }
              pk[j >> 1] = pack_bf16x2(hn[0], hn[1]);
            }
            // h_t goes to the OTHER buffer: the hidden-part MMAs of U1 may still be reading h_{t-1} while U0's epilogue runs
            st_shared_v4(sbase + ((step + 1) & 1u) * (uint32_t)(G_HCH * G_CHUNK_B) + (uint32_t)cidx * G_CHUNK_B + (uint32_t)row * 16, pk[0], pk[1],
                         pk[2], pk[3]);
            if (ok) *reinterpret_cast<uint4*>(otile + (size_t)cidx * a.R * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
        if (stamp) a.tl[step * 8 + 6] = clock64();
        if (s < L - 1) {
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_leader(bar_rdy, rank);
        } else {
          tc_fence_before();
          __syncwarp();
        }
      }
    }
  }

  // ---- teardown -------------------------------------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, 512);
  }
}

long long* g_gru_tl = nullptr;

}  // namespace

void gru_tc_set_timeline(long long* dev) { g_gru_tl = dev; }

// ---- tile helpers: fp32 (B, L, C) -> tiles, and the Linear after the GRU stack straight from tiles ---------------------
namespace {

__global__ void tiles_from_f32_kernel(const float* __restrict__ x, uint8_t* __restrict__ tiles, int B, int L, int C, int R, int n_blk) {
  const int nch = (C + 7) / 8;
  const int r = threadIdx.x;
  for (long long item = blockIdx.x; item < (long long)n_blk * L; item += gridDim.x) {
    const int blk = (int)(item / L), t = (int)(item % L);
    const int cw = blk * R + r;
    if (r >= R) continue;
    for (int c = 0; c < nch; ++c) {
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (cw < B && 8 * c + e < C) ? x[((size_t)cw * L + t) * C + 8 * c + e] : 0.f;
      *reinterpret_cast<uint4*>(tiles + ((((size_t)blk * L + t) * nch + c) * R + r) * 16) =
          make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    }
  }
}

// out[b, t, f] = bias[f] + sum_c W[f, c] * h[b, t, c] with h read from tiles (decoders.py:104-105, 118-119: dec*_outputs Linear).
// One thread = one codeword row and TWO time steps (every weight fetched from shared memory feeds two rows).
__global__ void __launch_bounds__(128) tiles_linear_kernel(const uint8_t* __restrict__ tiles, const float* __restrict__ w,
                                                           const float* __restrict__ bias, float* __restrict__ out, int B, int L, InMap im,
                                                           int n_xch, int F, int R, int n_blk) {
  extern __shared__ __align__(16) float w_s[];   // [n_xch * 8][8]: feature-padded rows, zero where the padded channel has no weight
  for (int i = threadIdx.x; i < n_xch * 8 * 8; i += blockDim.x) {
    const int cp = i / 8, f = i % 8, c = in_col(im, cp);
    w_s[i] = (c >= 0 && f < F) ? w[(size_t)f * im.in_ch + c] : 0.f;
  }
  __syncthreads();
  const int r = threadIdx.x;
  const int Lp = (L + 1) / 2;
  for (long long item = blockIdx.x; item < (long long)n_blk * Lp; item += gridDim.x) {
    const int blk = (int)(item / Lp), t0 = 2 * (int)(item % Lp);
    const int cw = blk * R + r;
    if (r >= R || cw >= B) continue;
    const bool two = t0 + 1 < L;
    float acc[2][8];
#pragma unroll
    for (int f = 0; f < 8; ++f) acc[0][f] = acc[1][f] = f < F ? bias[f] : 0.f;
    const size_t step_b = (size_t)n_xch * R * 16;
    const uint8_t* base = tiles + (((size_t)blk * L + t0) * n_xch) * R * 16 + (size_t)r * 16;
    for (int c = 0; c < n_xch; ++c) {
      const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(base + (size_t)c * R * 16));
      const uint4 v1 = two ? __ldg(reinterpret_cast<const uint4*>(base + step_b + (size_t)c * R * 16)) : make_uint4(0u, 0u, 0u, 0u);
      const uint32_t a0[4] = {v0.x, v0.y, v0.z, v0.w}, a1[4] = {v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float h0 = __uint_as_float((e & 1) ? (a0[e >> 1] & 0xFFFF0000u) : (a0[e >> 1] << 16));
        const float h1 = __uint_as_float((e & 1) ? (a1[e >> 1] & 0xFFFF0000u) : (a1[e >> 1] << 16));
        const float4 wa = *reinterpret_cast<const float4*>(w_s + (size_t)(8 * c + e) * 8);
        const float4 wb = *reinterpret_cast<const float4*>(w_s + (size_t)(8 * c + e) * 8 + 4);
        const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
        for (int f = 0; f < 8; ++f) {
          acc[0][f] = fmaf(wv[f], h0, acc[0][f]);
          acc[1][f] = fmaf(wv[f], h1, acc[1][f]);
        }
      }
    }
    float* o = out + ((size_t)cw * L + t0) * F;
    for (int f = 0; f < F; ++f) o[f] = acc[0][f];
    if (two)
      for (int f = 0; f < F; ++f) o[F + f] = acc[1][f];
  }
}

inline int grp_pad_of(int grp_valid) { return 8 * ((grp_valid + 7) / 8); }
inline int n_xch_of(int in_ch, int grp_valid) { return ((in_ch + grp_valid - 1) / grp_valid) * grp_pad_of(grp_valid) / 8; }

}  // namespace

bool gru_tc_supported(int H, int in_ch, int grp_valid, const char** why) {
  static thread_local char msg[128];
  *why = msg;
  if (H < 4 || H > 104 || H % 4) { snprintf(msg, sizeof msg, "hidden size %d (tensor path: multiples of 4 up to 104)", H); return false; }
  if (in_ch < 1 || grp_valid < 1 || n_xch_of(in_ch, grp_valid) > G_XCH_MAX) {
    snprintf(msg, sizeof msg, "input size %d in groups of %d (tensor path: at most %d chunks of 8)", in_ch, grp_valid, G_XCH_MAX);
    return false;
  }
  *why = nullptr;
  return true;
}

int gru_tc_rows_per_block(int B) {
  // blocks of R codewords, two per cluster: the largest R (<= 128, multiple of 32) that still gives every SM a block
  int R = 32 * ((B + 148 * 32 - 1) / (148 * 32));
  return std::max(32, std::min(128, R));
}

size_t gru_tc_packed_bytes(int /*H*/, int in_ch, int grp_valid) {
  const GruGeom g = gru_geom(n_xch_of(in_ch, grp_valid));
  return (size_t)2 * (g.w_rz_b + g.w_nx_b + g.w_nh_b);
}

int gru_tc_pack(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* packed, int H, int in_ch, int grp_valid,
                cudaStream_t s) {
  const size_t elems = gru_tc_packed_bytes(H, in_ch, grp_valid) / 2;
  const InMap im{in_ch, grp_valid, grp_pad_of(grp_valid)};
  gru_pack_kernel<<<(int)std::min<size_t>((elems + 255) / 256, 148 * 8), 256, 0, s>>>(w_ih, w_hh, b_ih, b_hh,
                                                                                    reinterpret_cast<__nv_bfloat16*>(packed), H, im,
                                                                                    n_xch_of(in_ch, grp_valid));
  return after_launch("gru_pack_kernel");
}

int gru_tc_direction(const void* packed, const void* x_tiles, void* out_tiles, int B, int L, int H, int in_ch, int grp_valid, int R,
                     int out_chunks, int out_c0, int reverse, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < 256) { set_error("tae_gru_direction_bf16: workspace %zu < 256 bytes", ws_bytes); return TAE_EWORKSPACE; }
  static DeviceOnce once;
  int n_sm = 0;
  {
    int rc = device_once(once, "gru_pair_kernel", [](int dev) -> int {
      int rc2 = require_sm100(dev, "the bf16 GRU path");
      if (rc2) return rc2;
      cudaError_t e = cudaFuncSetAttribute(gru_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gru_geom(G_XCH_MAX).total);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(gru_pair_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
      return TAE_OK;
    }, &n_sm);
    if (rc) return rc;
  }
  GruArgs a{};
  a.wimg = reinterpret_cast<const uint8_t*>(packed);
  a.x = reinterpret_cast<const uint8_t*>(x_tiles);
  a.out = reinterpret_cast<uint8_t*>(out_tiles);
  a.err = wait_code_slot(ws);
  a.B = B; a.L = L; a.H = H; a.hch = (H + 7) / 8; a.n_xch = n_xch_of(in_ch, grp_valid); a.R = R; a.out_chunks = out_chunks; a.out_c0 = out_c0; a.reverse = reverse;
  const int n_blk = (B + R - 1) / R;
  a.n_pairs = (n_blk + 1) / 2;
  a.tl = g_gru_tl;
  const int n_clusters = std::min(a.n_pairs, n_sm / 2);
  gru_pair_kernel<<<2 * n_clusters, G_THREADS, gru_geom(a.n_xch).total, s>>>(a);
  return after_launch("gru_pair_kernel");
}

int gru_tc_tiles_from_f32(const float* x, void* tiles, int B, int L, int C, int R, cudaStream_t s) {
  const int n_blk = 2 * (((B + R - 1) / R + 1) / 2);          // whole pairs: the second block of the last pair may be all padding
  tiles_from_f32_kernel<<<(int)std::min<long long>((long long)n_blk * L, 148 * 16), 128, 0, s>>>(x, reinterpret_cast<uint8_t*>(tiles), B, L, C, R, n_blk);
  return after_launch("tiles_from_f32_kernel");
}

int gru_tc_linear(const void* tiles, const float* w, const float* bias, float* out, int B, int L, int in_ch, int grp_valid, int F, int R,
                  cudaStream_t s) {
  const int n_blk = (B + R - 1) / R, n_xch = n_xch_of(in_ch, grp_valid);
  const InMap im{in_ch, grp_valid, grp_pad_of(grp_valid)};
  tiles_linear_kernel<<<(int)std::min<long long>((long long)n_blk * ((L + 1) / 2), 148 * 16), 128, (size_t)n_xch * 64 * sizeof(float), s>>>(
      reinterpret_cast<const uint8_t*>(tiles), w, bias, out, B, L, im, n_xch, F, R, n_blk);
  return after_launch("tiles_linear_kernel");
}

}  // namespace tae
