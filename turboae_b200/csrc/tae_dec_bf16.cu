// Fused DEC_LargeCNN.forward on 5th-gen tensor cores (TAE_PRECISION_BF16), sm_100a only.
//
// Reference arithmetic restated (paths relative to the reference checkout):
//   decoders.py:219-269 (turbo schedule), cnn_utils.py:36-46 (conv + ELU stack),
//   interleavers.py:15-21, 43-48 (row permutations).
//
// One persistent CTA per SM decodes "groups" of codewords end to end: all 2*I conv stacks,
// the Linear projections, the extrinsic subtractions, the (de)interleaves and the final
// sigmoid run out of shared memory / TMEM; HBM sees `received` once and `out` once, the
// weights stream from L2 through a bulk-copy (TMA unit, UBLKCP) ring.
//
// Mapping of one conv layer onto tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM):
//   * rows (M)      = positions.  A group is a buffer of 512 rows holding floor(514/(L+2))
//                     codewords, each followed by 2 all-zero separator rows, so the k=5
//                     receptive field never reads a neighbouring codeword (=> zero padding
//                     of cnn_utils.py:16 for free).  4 MMA tiles of M = 128.
//   * columns (N)   = output channels, 100 padded to 112.
//   * K             = input channels, 100 padded to 112 (7 x UMMA_K=16); the 5 taps are 5
//                     accumulating MMAs whose A descriptor start address is shifted by one
//                     16-byte row each: the activation buffer is the canonical no-swizzle
//                     K-major layout [K/8][rows][8] with SBO = 128 B, so "row + 1" is
//                     "address + 16".
//   * bias          = two constant-one input channels (bf16 hi/lo split of the fp32 bias in
//                     the centre tap), so the epilogue is ELU + bf16 pack only.
//   * Linear 100->F = one more MMA pass with N = 16.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "tae_common.cuh"

namespace tae {

namespace {

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
constexpr int GROUP_ROWS = 512;              // rows per group = 4 MMA tiles of 128
constexpr int BUF_ROWS = GROUP_ROWS + 4;     // + 2 zero rows in front and behind (tap halo)
constexpr int N_TILES = 4;
constexpr int KPAD = 112;                    // padded channel count (K and N of the 100->100 layers)
constexpr int KCH = KPAD / 8;                // 14 chunks of 8 channels (16 B)
constexpr int XCH = 2;                       // stack-input buffer: 16 channels = 2 chunks
constexpr int LIN_N = 16;                    // padded N of the Linear pass
constexpr int TAPS = 5;
constexpr uint32_t ROW_B = 16;                               // bytes per (row, chunk) entry
constexpr uint32_t LBO_ACT = BUF_ROWS * ROW_B;               // 8256: chunk stride of activation buffers
constexpr uint32_t ACT_BYTES = KCH * LBO_ACT;                // 115584
constexpr uint32_t XIN_BYTES = XCH * LBO_ACT;                // 16512
constexpr uint32_t LBO_W = KPAD * ROW_B;                     // 1792: chunk stride of a weight tap image
constexpr uint32_t TAP_BYTES = KCH * LBO_W;                  // 25088 (layers >= 1, one tap)
constexpr uint32_t L0_TAP_BYTES = XCH * LBO_W;               // 3584  (layer 0, one tap)
constexpr uint32_t L0_BYTES = TAPS * L0_TAP_BYTES;           // 17920 (layer 0, all taps = one stage)
constexpr uint32_t LBO_LIN = LIN_N * ROW_B;                  // 256
constexpr uint32_t LIN_BYTES = KCH * LBO_LIN;                // 3584
constexpr int W_STAGES = 2;
constexpr int N_EPI_WARPS = 8;
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
constexpr int WARP_PRODUCER = 8;
constexpr int WARP_MMA = 9;
constexpr int N_THREADS = 320;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_LIN_COL = N_TILES * KPAD;            // 448

// instruction descriptor (kind::f16): D fp32, A/B bf16, both K-major, M = 128
constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
constexpr uint32_t IDESC_N112 = make_idesc(KPAD);
constexpr uint32_t IDESC_N16 = make_idesc(LIN_N);

struct SmemLayout {
  uint32_t act, xin[2], wst[W_STAGES], pri[2], perm, inv_perm, bars, tmem_ptr, total;
};

__host__ __device__ inline SmemLayout make_smem_layout(int L, int F) {
  SmemLayout s{};
  uint32_t o = 0;
  s.act = o; o += ACT_BYTES;
  s.xin[0] = o; o += XIN_BYTES;
  s.xin[1] = o; o += XIN_BYTES;
  for (int i = 0; i < W_STAGES; ++i) { s.wst[i] = o; o += TAP_BYTES; }
  s.pri[0] = o; o += (uint32_t)F * BUF_ROWS * 4;
  s.pri[1] = o; o += (uint32_t)F * BUF_ROWS * 4;
  s.perm = o; o += (uint32_t)((L * 2 + 15) / 16 * 16);
  s.inv_perm = o; o += (uint32_t)((L * 2 + 15) / 16 * 16);
  s.bars = o; o += 128;
  s.tmem_ptr = o; o += 16;
  s.total = o;
  return s;
}

enum { BAR_W_FULL = 0, BAR_W_EMPTY = 2, BAR_ACC_FULL = 4, BAR_ACT_READY = 5 };  // 8-byte slots in SmemLayout::bars

struct DecKernelArgs {
  const uint8_t* wimg;
  const float* received;
  float* out;
  float* trace;
  float* dbg;          // debug: (2I * n_layer, 512, 112) bf16-rounded activations of group 0
  int* err;
  const int32_t* perm;
  const int32_t* inv_perm;
  int B, L, F, I, n_layer, extrinsic, n_groups, cw_per_group;
  uint32_t stack_bytes;
  uint32_t flags;      // bit0: swap LBO/SBO (probe only)
};

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug traps (context error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      if (err) atomicExch(err, code);
      __threadfence_system();
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// Shared-memory matrix descriptor, no swizzle, K-major: 8-row x 16-byte core matrices;
// SBO = byte distance between 8-row groups, LBO = byte distance between the two 8-element
// K chunks of one UMMA_K=16 slice.  Bits [46,48) = 1 is the Blackwell descriptor version.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// Arrives on the mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float elu_fast(float v) {
  const float e = fast_exp2(v * 1.4426950408889634f) - 1.0f;
  return v > 0.f ? v : e;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint16_t v) {
  asm volatile("st.shared.b16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint16_t ld_shared_u16(uint32_t addr) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint16_t bf16_bits(float v) {
  __nv_bfloat16 t = __float2bfloat16_rn(v);
  return *reinterpret_cast<uint16_t*>(&t);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory"); }

// ------------------------------------------------------------------------------------------
// weight image
//   per stack:  [layer 0: 5 taps x [2][112][8]] [layers 1..n-1: 5 taps x [14][112][8]] [Linear: [14][16][8]]
//   element (kc, n, e) = W[o = n][c = 8 kc + e][tap]; channel `cin` / `cin+1` of the centre tap carry
//   bias_hi / bias_lo, and output channels `units`, `units+1` regenerate the constant one.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__global__ void pack_dec_bf16_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ img,
                                     const DecStackLayout* __restrict__ lay, int n_stacks, int n_layer, int units,
                                     int F, uint32_t stack_elems) {
  const size_t total = (size_t)n_stacks * stack_elems;
  const uint32_t l0_elems = L0_BYTES / 2, tap_elems = TAP_BYTES / 2, layer_elems = TAPS * tap_elems;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int st = (int)(idx / stack_elems);
    uint32_t r = (uint32_t)(idx % stack_elems);
    const DecStackLayout& S = lay[st];
    float v = 0.f;
    if (r < l0_elems) {                                   // layer 0
      const int t = r / (L0_TAP_BYTES / 2);
      r %= (L0_TAP_BYTES / 2);
      const int kc = r / (KPAD * 8), n = (r / 8) % KPAD, e = r % 8;
      const int c = kc * 8 + e, cin = 2 + F;
      const float* w = params + S.conv[0].w_off;
      const float* b = params + S.conv[0].b_off;
      if (n < units && c < cin) v = w[((size_t)n * cin + c) * TAPS + t];
      else if (t == TAPS / 2 && c == cin) v = (n < units) ? bf16_hi(b[n]) : ((n == units || n == units + 1) ? 1.f : 0.f);
      else if (t == TAPS / 2 && c == cin + 1) v = (n < units) ? (b[n] - bf16_hi(b[n])) : 0.f;
    } else if (r < l0_elems + (uint32_t)(n_layer - 1) * layer_elems) {   // layers 1..n-1
      r -= l0_elems;
      const int j = 1 + r / layer_elems;
      r %= layer_elems;
      const int t = r / tap_elems;
      r %= tap_elems;
      const int kc = r / (KPAD * 8), n = (r / 8) % KPAD, e = r % 8;
      const int c = kc * 8 + e;
      const float* w = params + S.conv[j].w_off;
      const float* b = params + S.conv[j].b_off;
      if (n < units && c < units) v = w[((size_t)n * units + c) * TAPS + t];
      else if (t == TAPS / 2 && c == units) v = (n < units) ? bf16_hi(b[n]) : ((n == units || n == units + 1) ? 1.f : 0.f);
      else if (t == TAPS / 2 && c == units + 1) v = (n < units) ? (b[n] - bf16_hi(b[n])) : 0.f;
    } else {                                              // Linear
      r -= l0_elems + (uint32_t)(n_layer - 1) * layer_elems;
      const int kc = r / (LIN_N * 8), n = (r / 8) % LIN_N, e = r % 8;
      const int c = kc * 8 + e;
      const float* w = params + S.lin_w_off;
      const float* b = params + S.lin_b_off;
      if (n < S.fout && c < units) v = w[(size_t)n * units + c];
      else if (c == units) v = (n < S.fout) ? bf16_hi(b[n]) : 0.f;
      else if (c == units + 1) v = (n < S.fout) ? (b[n] - bf16_hi(b[n])) : 0.f;
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// the fused decoder kernel
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(N_THREADS, 1) dec_bf16_kernel(const DecKernelArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemLayout S = make_smem_layout(a.L, a.F);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int L = a.L, F = a.F, CW_ROWS = a.L + 2;
  const int n_stacks = 2 * a.I;
  const int one_ch = 2 + F;                        // first constant-one channel of the stack input

  const uint32_t bar_w_full = sbase + S.bars + 8 * BAR_W_FULL;
  const uint32_t bar_w_empty = sbase + S.bars + 8 * BAR_W_EMPTY;
  const uint32_t bar_acc_full = sbase + S.bars + 8 * BAR_ACC_FULL;
  const uint32_t bar_act_ready = sbase + S.bars + 8 * BAR_ACT_READY;

  // ---- one-time setup ----------------------------------------------------------------------
  for (uint32_t i = threadIdx.x * 16; i < S.bars; i += N_THREADS * 16)       // zero ACT, XIN, weight ring, PRI, perms
    st_shared_v4(sbase + i, 0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < W_STAGES; ++i) { mbar_init(bar_w_full + 8 * i, 1); mbar_init(bar_w_empty + 8 * i, 1); }
    mbar_init(bar_acc_full, 1);
    mbar_init(bar_act_ready, N_EPI_THREADS);
    fence_barrier_init();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < L; i += N_THREADS) {
    st_shared_u16(sbase + S.perm + 2 * i, (uint16_t)a.perm[i]);
    st_shared_u16(sbase + S.inv_perm + 2 * i, (uint16_t)a.inv_perm[i]);
  }
  if (warp == WARP_PRODUCER) tmem_alloc(sbase + S.tmem_ptr, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + S.tmem_ptr) : "memory");

  if (warp == WARP_PRODUCER) {
    // ================= weight producer: bulk copies through a W_STAGES ring ===================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
        for (int st = 0; st < n_stacks; ++st) {
          const uint8_t* src = a.wimg + (size_t)st * a.stack_bytes;
          const int n_stage = 2 + TAPS * (a.n_layer - 1);
          for (int i = 0; i < n_stage; ++i) {
            const uint32_t bytes = (i == 0) ? L0_BYTES : (i == n_stage - 1 ? LIN_BYTES : TAP_BYTES);
            mbar_wait(bar_w_empty + 8 * stage, phase ^ 1, a.err, 1);
            mbar_arrive_expect_tx(bar_w_full + 8 * stage, bytes);
            bulk_g2s(sbase + S.wst[stage], src, bytes, bar_w_full + 8 * stage);
            src += bytes;
            if (++stage == W_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ================= MMA issuer (one thread) ==================================================
    if (lane == 0) {
      uint32_t stage = 0, wphase = 0, act_phase = 0;
      const uint32_t sbo = 128;
      for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
        for (int st = 0; st < n_stacks; ++st) {
          const uint32_t xin = sbase + S.xin[st & 1];
          for (int layer = 0; layer <= a.n_layer; ++layer) {
            mbar_wait(bar_act_ready, act_phase, a.err, 2);
            act_phase ^= 1;
            tc_fence_after();
            if (layer == 0) {
              mbar_wait(bar_w_full + 8 * stage, wphase, a.err, 3);
              tc_fence_after();
              const uint32_t wb = sbase + S.wst[stage];
#pragma unroll 1
              for (int t = 0; t < TAPS; ++t) {
                const uint64_t bdesc = make_desc(wb + t * L0_TAP_BYTES, LBO_W, sbo);
#pragma unroll
                for (int m = 0; m < N_TILES; ++m) {
                  const uint64_t adesc = make_desc(xin + (uint32_t)(128 * m + t) * ROW_B, LBO_ACT, sbo);
                  umma_bf16(tmem_base + m * KPAD, adesc, bdesc, IDESC_N112, t > 0);
                }
              }
              umma_commit(bar_w_empty + 8 * stage);
              if (++stage == W_STAGES) { stage = 0; wphase ^= 1; }
            } else if (layer < a.n_layer) {
              const uint32_t act = sbase + S.act;
#pragma unroll 1
              for (int t = 0; t < TAPS; ++t) {
                mbar_wait(bar_w_full + 8 * stage, wphase, a.err, 4);
                tc_fence_after();
                const uint32_t wb = sbase + S.wst[stage];
#pragma unroll 1
                for (int m = 0; m < N_TILES; ++m) {
                  const uint32_t arow = act + (uint32_t)(128 * m + t) * ROW_B;
#pragma unroll
                  for (int ks = 0; ks < KPAD / 16; ++ks) {
                    const uint64_t adesc = make_desc(arow + 2 * ks * LBO_ACT, LBO_ACT, sbo);
                    const uint64_t bdesc = make_desc(wb + 2 * ks * LBO_W, LBO_W, sbo);
                    umma_bf16(tmem_base + m * KPAD, adesc, bdesc, IDESC_N112, (t > 0 || ks > 0));
                  }
                }
                umma_commit(bar_w_empty + 8 * stage);
                if (++stage == W_STAGES) { stage = 0; wphase ^= 1; }
              }
            } else {
              mbar_wait(bar_w_full + 8 * stage, wphase, a.err, 5);
              tc_fence_after();
              const uint32_t wb = sbase + S.wst[stage];
              const uint32_t act = sbase + S.act;
#pragma unroll 1
              for (int m = 0; m < N_TILES; ++m) {
                const uint32_t arow = act + (uint32_t)(128 * m + 2) * ROW_B;   // centre "tap": no shift
#pragma unroll
                for (int ks = 0; ks < KPAD / 16; ++ks) {
                  const uint64_t adesc = make_desc(arow + 2 * ks * LBO_ACT, LBO_ACT, sbo);
                  const uint64_t bdesc = make_desc(wb + 2 * ks * LBO_LIN, LBO_LIN, sbo);
                  umma_bf16(tmem_base + TMEM_LIN_COL + m * LIN_N, adesc, bdesc, IDESC_N16, ks > 0);
                }
              }
              umma_commit(bar_w_empty + 8 * stage);
              if (++stage == W_STAGES) { stage = 0; wphase ^= 1; }
            }
            umma_commit(bar_acc_full);
          }
        }
      }
    }
  } else {
    // ================= epilogue warps (8): TMEM -> ELU -> bf16 -> shared memory =================
    const int q = warp & 3;          // TMEM lane quadrant this warp may read
    const int half = warp >> 2;      // tiles {half, half + 2}
    const int tid = threadIdx.x;     // 0..255
    uint32_t acc_phase = 0;
    int g_row[2], g_cw[2], g_l[2];
    for (int k = 0; k < 2; ++k) {
      g_row[k] = 128 * (half + 2 * k) + 32 * q + lane;
      g_cw[k] = g_row[k] / CW_ROWS;
      g_l[k] = g_row[k] % CW_ROWS;
    }
    for (int grp = blockIdx.x; grp < a.n_groups; grp += gridDim.x) {
      const int cw0 = grp * a.cw_per_group;
      const int n_cw = min(a.cw_per_group, a.B - cw0);
      bool valid[2];
      for (int k = 0; k < 2; ++k) valid[k] = (g_l[k] < L) && (g_cw[k] < n_cw);

      // ---- load this group's received symbols into the two stack-input buffers ---------------
      for (uint32_t i = tid * 16; i < 2 * XIN_BYTES; i += N_EPI_THREADS * 16) st_shared_v4(sbase + S.xin[0] + i, 0u, 0u, 0u, 0u);
      for (uint32_t i = tid * 16; i < (uint32_t)F * BUF_ROWS * 4; i += N_EPI_THREADS * 16)
        st_shared_v4(sbase + S.pri[0] + i, 0u, 0u, 0u, 0u);
      epi_bar_sync();
      for (int i = tid; i < n_cw * L; i += N_EPI_THREADS) {
        const int c = i / L, l = i % L;
        const float* r = a.received + ((size_t)(cw0 + c) * L + l) * 3;
        const float r0 = r[0], r1 = r[1], r2 = r[2];
        const uint32_t row = (uint32_t)(c * CW_ROWS + l + 2) * ROW_B;
        const uint32_t row_i = (uint32_t)(c * CW_ROWS + ld_shared_u16(sbase + S.inv_perm + 2 * l) + 2) * ROW_B;
        st_shared_u16(sbase + S.xin[0] + row + 0, bf16_bits(r0));        // r_sys          (decoders.py:221)
        st_shared_u16(sbase + S.xin[0] + row + 2, bf16_bits(r1));        // r_par1         (decoders.py:223)
        st_shared_u16(sbase + S.xin[1] + row_i + 0, bf16_bits(r0));      // r_sys_int[i] = r_sys[p[i]]  (:222)
        st_shared_u16(sbase + S.xin[1] + row + 2, bf16_bits(r2));        // r_par2         (decoders.py:224)
        for (int o = 0; o < 2; ++o) {                                    // constant-one (bias) channels
          const int ch = one_ch + o;
          const uint32_t off = (uint32_t)(ch >> 3) * LBO_ACT + row + 2 * (ch & 7);
          st_shared_u16(sbase + S.xin[0] + off, 0x3F80);
          st_shared_u16(sbase + S.xin[1] + off, 0x3F80);
        }
      }
      fence_proxy_async();
      mbar_arrive(bar_act_ready);

      for (int st = 0; st < n_stacks; ++st) {
        for (int layer = 0; layer <= a.n_layer; ++layer) {
          mbar_wait(bar_acc_full, acc_phase, a.err, 6);
          acc_phase ^= 1;
          tc_fence_after();
          if (layer < a.n_layer) {
            // conv layer epilogue: ELU, bf16, in-place store (all MMAs of this layer have completed)
            for (int k = 0; k < 2; ++k) {
              const int m = half + 2 * k;
              const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(m * KPAD);
              const uint32_t dst = sbase + S.act + (uint32_t)(g_row[k] + 2) * ROW_B;
#pragma unroll 1
              for (int cb = 0; cb < KPAD / 16; ++cb) {
                uint32_t r[16];
                tmem_ld16(taddr + cb * 16, r);
                tmem_ld_wait();
                uint32_t p[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float v0 = __uint_as_float(r[2 * j]), v1 = __uint_as_float(r[2 * j + 1]);
                  v0 = valid[k] ? elu_fast(v0) : 0.f;
                  v1 = valid[k] ? elu_fast(v1) : 0.f;
                  p[j] = pack_bf16x2(v0, v1);
                }
                st_shared_v4(dst + (2 * cb) * LBO_ACT, p[0], p[1], p[2], p[3]);
                st_shared_v4(dst + (2 * cb + 1) * LBO_ACT, p[4], p[5], p[6], p[7]);
                if (a.dbg && grp == 0) {
                  float* d = a.dbg + ((size_t)(st * a.n_layer + layer) * GROUP_ROWS + g_row[k]) * KPAD + cb * 16;
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    d[2 * j] = __uint_as_float(p[j] << 16);
                    d[2 * j + 1] = __uint_as_float(p[j] & 0xFFFF0000u);
                  }
                }
              }
            }
            fence_proxy_async();
          } else {
            // Linear epilogue: extrinsic subtraction + (de)interleave into the next stack's input
            const bool last = (st == n_stacks - 1);
            const int fout = last ? 1 : F;
            const uint32_t pri_cur = sbase + S.pri[st & 1], pri_nxt = sbase + S.pri[(st & 1) ^ 1];
            const uint32_t xin_nxt = sbase + S.xin[(st & 1) ^ 1];
            const uint32_t map = sbase + ((st & 1) ? S.perm : S.inv_perm);   // where position l lands
            for (int k = 0; k < 2; ++k) {
              const int m = half + 2 * k;
              uint32_t r[16];
              tmem_ld16(tmem_base + ((uint32_t)(32 * q) << 16) + TMEM_LIN_COL + (uint32_t)(m * LIN_N), r);
              tmem_ld_wait();
              if (!valid[k]) continue;
              const int cw = cw0 + g_cw[k], l = g_l[k];
              if (a.trace) {
                float* tr = a.trace + (((size_t)st * a.B + cw) * L + l) * F;
                for (int f = 0; f < fout; ++f) tr[f] = __uint_as_float(r[f]);
              }
              if (last) {
                const int dl = ld_shared_u16(sbase + S.perm + 2 * l);         // deinterleave: out[p[l]] = x[l]
                a.out[(size_t)cw * L + dl] = 1.f / (1.f + __expf(-__uint_as_float(r[0])));   // decoders.py:267
              } else {
                const int dl = ld_shared_u16(map + 2 * l);
                const uint32_t drow = (uint32_t)(g_cw[k] * CW_ROWS + dl + 2);
                for (int f = 0; f < F; ++f) {
                  const float prior = a.extrinsic ? ld_shared_f32(pri_cur + ((uint32_t)f * BUF_ROWS + g_row[k] + 2) * 4) : 0.f;
                  const float ext = __uint_as_float(r[f]) - prior;             // decoders.py:235-236, 246-247
                  st_shared_f32(pri_nxt + ((uint32_t)f * BUF_ROWS + drow) * 4, ext);
                  const int ch = 2 + f;
                  st_shared_u16(xin_nxt + (uint32_t)(ch >> 3) * LBO_ACT + drow * ROW_B + 2 * (ch & 7), bf16_bits(ext));
                }
              }
            }
            fence_proxy_async();
          }
          tc_fence_before();
          if (!(st == n_stacks - 1 && layer == a.n_layer)) mbar_arrive(bar_act_ready);
        }
      }
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == WARP_PRODUCER) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// UMMA probe (debug / self-test): D[128 x N] = A[rows shift..shift+127][K] * B[N][K]^T with the exact
// descriptor scheme of the decoder kernel (no-swizzle K-major, row shift by start address).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D,
                  int R, int K, int N, int shift, uint32_t flags, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t lbo_a = (uint32_t)R * 16, lbo_b = (uint32_t)N * 16;
  const uint32_t a_off = 0, b_off = (uint32_t)(K / 8) * lbo_a;
  const uint32_t bar_off = (b_off + (uint32_t)(K / 8) * lbo_b + 15) / 16 * 16, tptr_off = bar_off + 16;
  for (int i = threadIdx.x; i < R * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    reinterpret_cast<__nv_bfloat16*>(smem + a_off)[(size_t)(c / 8) * R * 8 + r * 8 + (c % 8)] = A[i];
  }
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) {
    const int n = i / K, c = i % K;
    reinterpret_cast<__nv_bfloat16*>(smem + b_off)[(size_t)(c / 8) * N * 8 + n * 8 + (c % 8)] = Bm[i];
  }
  if (threadIdx.x == 0) { mbar_init(sbase + bar_off, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(sbase + tptr_off, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + tptr_off) : "memory");
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int ks = 0; ks < K / 16; ++ks) {
      uint64_t ad, bd;
      if (flags & 1u) {   // hypothesis B: LBO/SBO roles swapped
        ad = make_desc(sbase + a_off + 2 * ks * lbo_a + shift * 16, 128, lbo_a);
        bd = make_desc(sbase + b_off + 2 * ks * lbo_b, 128, lbo_b);
      } else {
        ad = make_desc(sbase + a_off + 2 * ks * lbo_a + shift * 16, lbo_a, 128);
        bd = make_desc(sbase + b_off + 2 * ks * lbo_b, lbo_b, 128);
      }
      if (flags & 2u) { ad &= ~(1ull << 46); bd &= ~(1ull << 46); }
      umma_bf16(tmem_base, ad, bd, idesc, ks > 0);
    }
    umma_commit(sbase + bar_off);
  }
  mbar_wait(sbase + bar_off, 0, err, 7);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < N / 16; ++cb) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + cb * 16, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(32 * warp + lane) * N + cb * 16 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem_base, 128); }
}

static float* g_debug_dump = nullptr;   // set through tae_debug_set_dump()

static uint32_t stack_image_bytes(const TaeDecConfig& c) {
  return L0_BYTES + (uint32_t)(c.num_layer - 1) * TAPS * TAP_BYTES + LIN_BYTES;
}

}  // namespace

bool dec_bf16_supported(const TaeDecConfig& c, const char** why) {
  static thread_local char msg[160];
  *why = msg;
  if (c.kernel_size != TAPS) { snprintf(msg, sizeof msg, "kernel_size %d (only 5 is built for the tensor path)", c.kernel_size); return false; }
  if (c.num_unit + 2 > KPAD) { snprintf(msg, sizeof msg, "num_unit %d > %d", c.num_unit, KPAD - 2); return false; }
  if (c.num_iter_ft + 4 > 16) { snprintf(msg, sizeof msg, "num_iter_ft %d > 12", c.num_iter_ft); return false; }
  if (c.num_layer < 2) { snprintf(msg, sizeof msg, "num_layer %d < 2", c.num_layer); return false; }
  if (c.block_len + 2 > GROUP_ROWS + 2) { snprintf(msg, sizeof msg, "block_len %d > %d (one codeword must fit a 512-row group)", c.block_len, GROUP_ROWS); return false; }
  if (make_smem_layout(c.block_len, c.num_iter_ft).total > 227 * 1024) { snprintf(msg, sizeof msg, "shared memory budget exceeded (num_iter_ft %d, block_len %d)", c.num_iter_ft, c.block_len); return false; }
  *why = nullptr;
  return true;
}

size_t dec_packed_bytes_bf16(const TaeDecConfig& c) { return (size_t)2 * c.num_iteration * stack_image_bytes(c); }

int dec_pack_bf16(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s) {
  DecStackLayout lay[64];
  dec_layout(c, lay);
  const int n_stacks = 2 * c.num_iteration;
  // the layout table is tiny: stage it in a stream-ordered temporary
  DecStackLayout* d_lay = nullptr;
  cudaError_t e = cudaMallocAsync(&d_lay, sizeof(DecStackLayout) * n_stacks, s);
  if (e != cudaSuccess) { set_error("cudaMallocAsync: %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  // pageable source is copied to a staging buffer before the call returns, so `lay` may die.
  e = cudaMemcpyAsync(d_lay, lay, sizeof(DecStackLayout) * n_stacks, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) { set_error("cudaMemcpyAsync: %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  const uint32_t stack_elems = stack_image_bytes(c) / 2;
  const size_t total = (size_t)n_stacks * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  pack_dec_bf16_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<__nv_bfloat16*>(packed), d_lay, n_stacks,
                                              c.num_layer, c.num_unit, c.num_iter_ft, stack_elems);
  int rc = after_launch("pack_dec_bf16_kernel");
  cudaFreeAsync(d_lay, s);
  return rc;
}

size_t dec_workspace_bytes_bf16(const TaeDecConfig&, int) { return 256; }

int dec_forward_bf16(const TaeDecConfig& c, const float*, const void* packed, const float* received,
                     const int32_t* perm, const int32_t* inv_perm, float* out, float* trace, int B, void* ws,
                     size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < 256) { set_error("tae_dec_forward(bf16): workspace %zu < 256 bytes", ws_bytes); return TAE_EWORKSPACE; }
  static int n_sm = 0;
  static bool attr_done = false;
  const SmemLayout S = make_smem_layout(c.block_len, c.num_iter_ft);
  if (!attr_done) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    cudaError_t e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    if (prop.major != 10) { set_error("bf16 path needs an sm_100a device (found sm_%d%d)", prop.major, prop.minor); return TAE_EUNSUPPORTED; }
    n_sm = prop.multiProcessorCount;
    e = cudaFuncSetAttribute(dec_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(dec_bf16_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    attr_done = true;
  }
  DecKernelArgs a{};
  a.wimg = reinterpret_cast<const uint8_t*>(packed);
  a.received = received;
  a.out = out;
  a.trace = trace;
  a.dbg = g_debug_dump;
  a.err = reinterpret_cast<int*>(align_up(reinterpret_cast<uintptr_t>(ws), 16));
  a.perm = perm;
  a.inv_perm = inv_perm;
  a.B = B; a.L = c.block_len; a.F = c.num_iter_ft; a.I = c.num_iteration; a.n_layer = c.num_layer;
  a.extrinsic = c.extrinsic;
  a.cw_per_group = (GROUP_ROWS + 2) / (c.block_len + 2);
  a.n_groups = (B + a.cw_per_group - 1) / a.cw_per_group;
  a.stack_bytes = stack_image_bytes(c);
  a.flags = 0;
  const int grid = std::min(a.n_groups, n_sm);
  dec_bf16_kernel<<<grid, N_THREADS, S.total, s>>>(a);
  return after_launch("dec_bf16_kernel");
}

}  // namespace tae

// ---- debug / self-test entry points (not part of the drop-in surface) --------------------------
extern "C" {

// When non-NULL, every later bf16 tae_dec_forward dumps the bf16-rounded activations of group 0
// (first cw_per_group codewords) after each conv layer: (2I * num_layer, 512, 112) floats.
void tae_debug_set_dump(float* device_ptr) { tae::g_debug_dump = device_ptr; }

// D (128, N) = A[shift : shift+128, :K] @ Bm[:N, :K]^T on the tensor cores with the decoder's
// descriptor scheme.  A: (R, K) bf16 row-major, Bm: (N, K) bf16 row-major, K % 16 == 0, N % 16 == 0.
int tae_debug_umma_probe(const void* A, const void* Bm, float* D, int32_t R, int32_t K, int32_t N, int32_t shift,
                         uint32_t flags, int* err, void* stream) {
  using namespace tae;
  if (K % 16 || N % 16 || N > 128 || shift < 0 || shift + 128 > R) { set_error("umma_probe: bad shape"); return TAE_EINVAL; }
  const size_t smem = (size_t)(K / 8) * R * 16 + (size_t)(K / 8) * N * 16 + 64;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(umma_probe): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  if (smem > 200 * 1024) { set_error("umma_probe: too large"); return TAE_EINVAL; }
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                                          reinterpret_cast<const __nv_bfloat16*>(Bm), D, R, K, N, shift,
                                                          flags, err);
  return after_launch("umma_probe_kernel");
}

}  // extern "C"
