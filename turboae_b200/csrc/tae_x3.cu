// ENC_interCNN.forward and DEC_LargeCNN.forward on the 5th-gen tensor cores with fp32-class accuracy (TAE_PRECISION_F16X3),
// sm_100a only.
//
// Reference arithmetic restated (paths relative to the reference checkout):
//   encoders.py:362-373 (three branches, Linear + ELU), decoders.py:219-269 (turbo schedule),
//   cnn_utils.py:36-46 (conv + ELU stack), interleavers.py:15-21, 43-48 (row permutations).
//
// Why: the plain bf16 tensor path rounds activations and weights to 8 mantissa bits (codes off by up to 3e-2, posteriors by
// 3e-2 on seeded batches): it meets the BER gate, not north_star's elementwise 1e-4.  Here every operand of a conv layer is
// split into two fp16 terms, x = x_hi + x_lo and W = W_hi + W_lo (22 mantissa bits together), and the layer is the three
// tcgen05.mma chains  x_hi W_hi + x_lo W_hi + x_hi W_lo  accumulated in fp32 TMEM (the x_lo W_lo term is below 2^-22).
// Bias, ELU, the Linear projections, the extrinsic subtraction and the priors stay in fp32 on the CUDA cores.  Measured
// (scripts/x3_accuracy.py, profiles/r02_x3_accuracy.md): codes within 6e-6 of the oracle; posteriors within 6e-6 at 2 dB and,
// at 0 dB on 300 000 posteriors, 99.99 % within 4.5e-5 with a maximum of 1e-4.  What is left is the tensor core's fp32
// accumulation: a CPU emulation of the scheme with exact accumulation gives 4e-6, with round-toward-zero accumulation per
// k-step 8e-5.  A bf16 split (-DTAE_X3_FP16=0; 16 bits) measures 2x worse.
//
// Execution model
//   * clusters of 2 CTAs (one per SM of a TPC), persistent.  The two CTAs' 256-row activation buffers form ONE 512-row space in
//     which the codewords of a work unit are laid out back to back, each followed by 2 all-zero separator rows (the zero padding
//     of cnn_utils.py:16): 5 codewords of block length 100 (97.7 % of the MMA rows are real positions; without the shared space
//     2 + 2).  A codeword may straddle the CTA boundary: the two rows either side of it are MIRRORED into the neighbour's halo
//     rows by the threads that own them -- st.async through distributed shared memory, completing a transaction barrier in the
//     receiving CTA, so the writer needs no fence -- and the stack-input scatter writes to whichever CTA owns the row.
//   * every layer is tcgen05.mma.cta_group::2, M = 256 = tile m of both CTAs, N = 112; each CTA stages its half of the weight
//     columns.  Activations live in shared memory twice (hi and lo images), 16-bit, canonical no-swizzle K-major layout
//     [13 chunks of 8 channels][264 rows][8]: tap t of the convolution is the same buffer addressed 16*t bytes later.
//     K of a units->units layer = 33 k-steps of 16, CHUNK-major: 6 pairs of chunks x 5 taps (LBO = one chunk), then chunk 12
//     with two taps per k-step (LBO = 16 bytes: taps (0,1), (2,3), (4,-)).  The first layer ((2+F) or 1 -> units) is 3
//     k-steps of its one chunk.
//   * weights stream from L2 through a bulk-copy ring: per layer 11 slots of W_hi (each used by the x_hi and the x_lo chain)
//     then 11 slots of W_lo (x_hi chain); two MMA issuer warps in the leader CTA, one per tile (two interleaved issue streams);
//     the peer relays "my half of the slot has landed" and "my halo rows have landed".
//   * the epilogue (8 warps per CTA, one row per thread: tcgen05.ld, + bias, ELU, split into hi / lo, st.shared in place)
//     publishes its output in 7 chunk stages; the next layer's MMAs start on the stages that are there (accumulators are
//     double-buffered by layer parity), so the epilogue of layer j overlaps the MMAs of layer j+1.  The last layer of a stack
//     keeps its output in registers and applies the Linear there (fp32): nothing is rounded between the last conv layer and the
//     stack output.
//   * decoder: the stack inputs (received values and priors) are kept as an fp32 master copy [row][8] next to their hi / lo
//     operand chunks; the extrinsic subtraction uses the fp32 prior, and (de)interleave is the row index of the store.
//   * -DTAE_X3_STRADDLE=0: one group of floor(258/(L+2)) codewords per CTA, no cross-CTA rows; -DTAE_X3_PAIR=0: one CTA per
//     group, cta_group::1.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>

#include "tae_common.cuh"
#include "tae_umma.cuh"

namespace tae {

// Operand format of the split: 1 (default) = fp16 hi + fp16 lo (22 mantissa bits together; values are clamped to +-65504, fp16's
// finite range), 0 = bf16 hi + bf16 lo (16 bits, full fp32 range).  Same tcgen05.mma kind::f16, same rate.
#ifndef TAE_X3_FP16
#define TAE_X3_FP16 1
#endif

namespace {
namespace x3 {

constexpr bool FP16 = TAE_X3_FP16 != 0;
// instruction descriptor (kind::f16): D fp32, A/B fp16 (format 0) or bf16 (format 1), both K-major
__host__ __device__ constexpr uint32_t make_idesc_x3(int m, int n) {
  return (1u << 4) | (FP16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
constexpr float F16_MAX = 65504.f;
// Optional power-of-two scaling of the stored operands (the epilogue undoes both factors with the multiply of its bias FMA):
// it would keep the lo terms of small weights out of fp16's subnormal range (a weight of 0.03 has a lo term of ~7e-6, quantised
// to 6e-8).  Measured with 256 / 8: no change of the output error (the tensor core's accumulation dominates), so both stay 1 and
// the full fp16 range is available to the activations.
#ifndef TAE_X3_SCALE_W
#define TAE_X3_SCALE_W 1.f
#endif
#ifndef TAE_X3_SCALE_X
#define TAE_X3_SCALE_X 1.f
#endif
constexpr float SCALE_W = FP16 ? TAE_X3_SCALE_W : 1.f, SCALE_X = FP16 ? TAE_X3_SCALE_X : 1.f;
constexpr float ACC_INV = 1.f / (SCALE_W * SCALE_X);

// 1 (default): clusters of 2 CTAs, tcgen05.mma.cta_group::2 with M = 256 = tile m of BOTH CTAs' groups; each CTA stages only its half
// of the weight columns (56 of 112), which halves the weight traffic from L2 and takes 1.75 KB per MMA off the shared-memory port
// that bounds the single-CTA version (0: one CTA per group, cta_group::1).
#ifndef TAE_X3_PAIR
#define TAE_X3_PAIR 1
#endif
constexpr bool PAIR = TAE_X3_PAIR != 0;
constexpr int CG = PAIR ? 2 : 1;
// 1 (default with the pair): the two CTAs' 256-row buffers form ONE 512-row space in which codewords are laid out back to back
// (5 codewords of block length 100 instead of 2 x 2: 97.7 % instead of 80 % of the MMA rows are real positions).  A codeword may
// straddle the CTA boundary: the two rows either side of it are MIRRORED into the neighbour's halo rows through distributed shared
// memory (st.shared::cluster) by the threads that own them, and the stack-input scatter writes to whichever CTA owns the row.
#ifndef TAE_X3_STRADDLE
#define TAE_X3_STRADDLE TAE_X3_PAIR
#endif
constexpr bool STRADDLE = PAIR && (TAE_X3_STRADDLE != 0);
constexpr int UNIT_ROWS = STRADDLE ? 2 * 256 : 256;     // rows of the space codewords are packed into

constexpr int GROUP_ROWS = 256;
constexpr int N_TILES = 2;
constexpr int HALO_LO = 2, HALO_HI = 6;                 // rows in front / behind (the two-taps-per-k-step scheme reads "tap 5")
constexpr int BUF_ROWS = GROUP_ROWS + HALO_LO + HALO_HI;   // 264
constexpr uint32_t ROW_B = 16;
constexpr uint32_t CHUNK_B = BUF_ROWS * ROW_B;          // 4224
constexpr int NPAD = 112;                               // UMMA N
constexpr int N_CHUNKS = 13;                            // 104 channels
constexpr int UNITS_MAX = 104;
constexpr int TAPS = 5;
constexpr int KS_CONV = 33, KS_L0 = 3, KS_PER_SLOT = 3;
constexpr int SLOTS_PASS = KS_CONV / KS_PER_SLOT;       // 11 slots of W_hi, 11 of W_lo per units->units layer
constexpr int NCTA = NPAD / CG;                         // weight columns staged per CTA
constexpr uint32_t WCHUNK_B = NCTA * ROW_B;             // 8 K elements of this CTA's columns (1792 / 896 bytes)
constexpr uint32_t KSTEP_B = 2 * WCHUNK_B;
constexpr uint32_t SLOT_B = KS_PER_SLOT * KSTEP_B;      // bytes of a slot in ONE CTA (10752 / 5376)
constexpr uint32_t SLOT_IMG_B = CG * SLOT_B;            // bytes of a slot in the weight image: [cta half][k-step][2 chunks][columns][8]
constexpr int NS = PAIR ? 8 : 6;                        // ring slots
constexpr int MAX_LAYER = 8, MAX_F = 5;
constexpr int TAB_BIAS = 0, TAB_V = MAX_LAYER * NPAD, TAB_C = TAB_V + MAX_F * NPAD, TAB_FLOATS = TAB_C + 8;   // one stack's tables (x 2: stack parity)
constexpr int N_EPI_WARPS = 8, N_EPI_THREADS = 256;     // warp w: tile w >> 2, TMEM lane quadrant w & 3
constexpr int WARP_MMA = 8, WARP_PRODUCER = 10;         // warps 8, 9: one MMA issuer per tile (two interleaved issue streams keep the
constexpr int N_THREADS = 352;                          // tensor pipe fed: a single issuing thread leaves a gap after every MMA)
constexpr uint32_t TMEM_COLS = 512;                     // 2 accumulator buffers (layer parity) x 2 tiles x 112 columns
constexpr uint32_t TMEM_BUF_COLS = N_TILES * NPAD;      // 224
constexpr int N_STAGES = 7;                             // epilogue stages = 16-column blocks = chunk pairs 0..5, then chunk 12

struct Smem {
  uint32_t act_hi, act_lo, xin_hi[2], xin_lo[2], master[2], wslot, tab, perm, inv_perm, bars, tmem_ptr, total;
};
// B_WFULL / B_WEMPTY[p]: ring slot p landed / its MMAs have completed.  B_ACC: all MMAs of a layer have completed (one phase per
// layer; waited for by every epilogue warp).  B_ACT: the inputs of a stack's FIRST layer are in place (group start; Linear
// epilogue of the previous stack), one phase per stack.  B_STAGE[c]: every epilogue warp has written chunk stage c of the layer
// output (one phase per non-final layer epilogue): the next layer's MMAs start on the chunks that are there while the epilogue is
// still writing the rest -- the K order of a units->units layer is chunk-major for that reason.
// B_ZERO (straddle): both CTAs have zeroed their stack-input buffers (cluster scope), so remote rows may be written.
// B_HALO[c] (straddle, every CTA): the neighbour's mirror rows of chunk stage c have landed in this CTA's halo rows (st.async
// complete_tx; armed by the waiter).  B_HREL[c] (leader): the peer's relay reports the peer's B_HALO[c].
enum { B_WFULL = 0, B_WEMPTY = NS, B_ACC = 2 * NS, B_ACT = 2 * NS + 1, B_STAGE = 2 * NS + 2, B_ZERO = 2 * NS + 2 + N_STAGES,
       B_HALO = 2 * NS + 3 + N_STAGES, B_HREL = 2 * NS + 3 + 2 * N_STAGES, N_BARS = 2 * NS + 3 + 3 * N_STAGES };
// bytes the two mirror rows of one chunk stage carry: 2 rows x (2 chunks, or chunk 12 alone) x (hi, lo) x 16 bytes
__host__ __device__ constexpr uint32_t halo_stage_bytes(int c) { return c < 6 ? 128u : 64u; }
// the last epilogue stage whose chunks slot s of the W_hi pass reads (k-steps 3s .. 3s+2, chunk-major: k-step 5 cp + t)
__host__ __device__ constexpr int stage_of_slot(int s) { return (3 * s + 2) < 30 ? (3 * s + 2) / 5 : 6; }

__host__ __device__ inline Smem make_smem() {
  Smem s{};
  uint32_t o = 0;
  s.act_hi = o; o += N_CHUNKS * CHUNK_B;
  s.act_lo = o; o += N_CHUNKS * CHUNK_B;
  s.xin_hi[0] = o; o += CHUNK_B;
  s.xin_hi[1] = o; o += CHUNK_B;
  s.xin_lo[0] = o; o += CHUNK_B;
  s.xin_lo[1] = o; o += CHUNK_B;
  s.master[0] = o; o += BUF_ROWS * 32;
  s.master[1] = o; o += BUF_ROWS * 32;
  s.wslot = o; o += NS * SLOT_B;
  s.tab = o; o += 2 * TAB_FLOATS * 4;
  s.perm = o; o += 1024;                            // u16 x block length (<= 510)
  s.inv_perm = o; o += 1024;
  s.bars = o; o += N_BARS * 8;
  s.tmem_ptr = o; o += 16;
  s.total = o;
  return s;
}

// flat-parameter layout of a sequence of conv stacks + Linear (same arithmetic as dec_layout / enc_layout of tae_common.cuh)
struct Layout {
  int n_stacks, n_layer, units, cin0, f_regular, f_last;
  __host__ __device__ size_t l0() const { return (size_t)units * cin0 * TAPS + units; }
  __host__ __device__ size_t lj() const { return (size_t)units * units * TAPS + units; }
  __host__ __device__ size_t base(int st) const { return (size_t)st * (l0() + (size_t)(n_layer - 1) * lj() + (size_t)f_regular * units + f_regular); }
  __host__ __device__ int fout(int st) const { return st == n_stacks - 1 ? f_last : f_regular; }
  __host__ __device__ size_t conv_w(int st, int j) const { return base(st) + (j == 0 ? 0 : l0() + (size_t)(j - 1) * lj()); }
  __host__ __device__ size_t conv_b(int st, int j) const { return conv_w(st, j) + (size_t)units * (j == 0 ? cin0 : units) * TAPS; }
  __host__ __device__ size_t lin_w(int st) const { return base(st) + l0() + (size_t)(n_layer - 1) * lj(); }
  __host__ __device__ size_t lin_b(int st) const { return lin_w(st) + (size_t)fout(st) * units; }
};

struct Args {
  const uint8_t* wimg;
  const float* params;
  const float* received;       // dec: (B, L, 3)
  float* out;                  // dec: (B, L, 1)
  float* trace;                // dec: NULL or (n_stacks, B, L, F)
  const float* u;              // enc: bits (B, L, 1)
  float* x_tx;                 // enc: un-normalised codes (B, L, 3)
  double* stats;               // enc: running (sum, sum of squares)
  int* err;
  const int32_t* perm;
  const int32_t* inv_perm;
  int B, L, F, n_stacks, n_layer, extrinsic, n_groups, n_units, cw_per_group, enc;   // cw_per_group: codewords per 256-row group (per 512-row unit when straddling)
  uint32_t stack_bytes;
  Layout lay;
};

// the 16-bit pattern of the split format nearest to x, and its value
__device__ __forceinline__ uint16_t half_bits(float x) {
  if (FP16) { const __half h = __float2half_rn(fminf(fmaxf(x, -F16_MAX), F16_MAX)); return *reinterpret_cast<const uint16_t*>(&h); }
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  return *reinterpret_cast<const uint16_t*>(&h);
}
__device__ __forceinline__ float half_value(uint16_t b) {
  if (FP16) return __half2float(*reinterpret_cast<const __half*>(&b));
  return __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(&b));
}

// K element e (0..15) of k-step ks of a units->units layer -> weight W[o, c, t] (0 for padding)
__device__ __forceinline__ float conv_w_elem(const float* __restrict__ w, int units, int o, int ks, int e) {
  if (o >= units) return 0.f;
  int c, t;
  if (ks < 30) { t = ks % 5; c = 16 * (ks / 5) + e; }
  else { t = 2 * (ks - 30) + (e >> 3); c = 96 + (e & 7); }
  if (t >= TAPS || c >= units) return 0.f;
  return w[((size_t)o * units + c) * TAPS + t];
}
__device__ __forceinline__ float l0_w_elem(const float* __restrict__ w, int units, int cin, int o, int ks, int e) {
  if (o >= units) return 0.f;
  const int t = 2 * ks + (e >> 3), c = e & 7;
  if (t >= TAPS || c >= cin) return 0.f;
  return w[((size_t)o * cin + c) * TAPS + t];
}

// weight image (bf16), per stack: [L0 hi slot][L0 lo slot][(n_layer-1) x (11 hi slots, 11 lo slots)]; one slot = 3 k-steps,
// one k-step = 2 chunks [112 columns][8 K elements]
__global__ void pack_x3_kernel(const float* __restrict__ params, uint16_t* __restrict__ img, const Layout lay, uint32_t stack_elems) {
  const size_t total = (size_t)lay.n_stacks * stack_elems;
  constexpr uint32_t slot_elems = SLOT_IMG_B / 2;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int st = (int)(idx / stack_elems);
    uint32_t r = (uint32_t)(idx % stack_elems);
    const int sl = (int)(r / slot_elems);
    r %= slot_elems;
    const int half = (int)(r / (SLOT_B / 2));          // which CTA of the pair stages these columns
    r %= (SLOT_B / 2);
    const int k3 = (int)(r / (KSTEP_B / 2)), h = (int)(r / (NCTA * 8)) & 1, n = half * NCTA + (int)(r / 8) % NCTA, e = h * 8 + (int)(r % 8);
    float v;
    int part;
    if (sl < 2) {
      part = sl;
      v = l0_w_elem(params + lay.conv_w(st, 0), lay.units, lay.cin0, n, k3, e);
    } else {
      const int q = sl - 2, j = 1 + q / (2 * SLOTS_PASS), within = q % (2 * SLOTS_PASS);
      part = within / SLOTS_PASS;
      v = conv_w_elem(params + lay.conv_w(st, j), lay.units, n, (within % SLOTS_PASS) * KS_PER_SLOT + k3, e);
    }
    v *= SCALE_W;
    const uint16_t hi = half_bits(v);
    img[idx] = part == 0 ? hi : half_bits(v - half_value(hi));
  }
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory"); }
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// ELU (alpha 1, cnn_utils.py:24-25): ex2.approx is accurate to ~2^-22 of its result, so e^v - 1 carries an absolute error of
// ~2.4e-7, far inside the 1e-4 gate
__device__ __forceinline__ float elu_x3(float v) { return v > 0.f ? v : fast_exp2(v * 1.4426950408889634f) - 1.0f; }
// (a, b) -> packed hi pair and packed lo pair with hi + lo = value to 16 (bf16) / 22 (fp16) mantissa bits
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (FP16) {
    a *= SCALE_X;
    b *= SCALE_X;
    const __half2 h = __floats2half2_rn(fminf(fmaxf(a, -F16_MAX), F16_MAX), fminf(fmaxf(b, -F16_MAX), F16_MAX));
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - hf.x, b - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
    return;
  }
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// distributed shared memory: address of the same offset in CTA `r` of the cluster, and stores through it
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t addr, uint32_t r) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(r));
  return ra;
}
__device__ __forceinline__ void st_cluster_b32(uint32_t ra, uint32_t v) { asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(ra), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_b16(uint32_t ra, uint16_t v) { asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(ra), "h"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_v2(uint32_t ra, uint32_t a, uint32_t b) { asm volatile("st.shared::cluster.v2.b32 [%0], {%1,%2};" ::"r"(ra), "r"(a), "r"(b) : "memory"); }
__device__ __forceinline__ void st_cluster_v4(uint32_t ra, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(ra), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 16 bytes into the peer's shared memory through the async proxy; completes 16 bytes of the transaction count of the mbarrier
// `rbar` (an address in the SAME peer CTA) once the data has been written: no fence on the writer's side
__device__ __forceinline__ void st_async_v4(uint32_t ra, uint32_t rbar, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1,%2,%3,%4}, [%5];" ::"r"(ra), "r"(a), "r"(b), "r"(c),
               "r"(d), "r"(rbar)
               : "memory");
}
// generic-proxy writes (to this CTA's or the peer's shared memory) before later async-proxy (tensor core) reads
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async.shared::cluster;" ::: "memory"); }

#if TAE_X3_PAIR
#define X3_CLUSTER __cluster_dims__(2, 1, 1)
#else
#define X3_CLUSTER
#endif
// commit of all previously issued MMAs: to this CTA's barrier, or (pair) to the barrier at the same offset in both CTAs
__device__ __forceinline__ void commit_x3(uint32_t bar) {
  if (PAIR) umma_commit_pair(bar, 3); else umma_commit_1(bar);
}

__global__ void X3_CLUSTER __launch_bounds__(N_THREADS, 1) x3_kernel(const Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem S = make_smem();
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // pair: the cluster's CTAs own groups 2 u and 2 u + 1 of work unit u; single: one group per unit
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int unit0 = PAIR ? (int)cluster_id_x() : (int)blockIdx.x, unit_stride = PAIR ? (int)n_clusters_x() : (int)gridDim.x;
  const int n_units = a.n_units;
  // arrive on the barrier the MMA issuers wait on: the leader CTA's (a local arrive when this IS the leader / the only CTA)
  auto arrive_issuer = [&](uint32_t b) { if (PAIR) mbar_arrive_leader(b, rank); else mbar_arrive_local(b); };
  const int L = a.L, F = a.F, CW_ROWS = a.L + 2;
  const int n_stacks = a.n_stacks, n_layer = a.n_layer;
  const int slots_per_stack = 2 + 2 * SLOTS_PASS * (n_layer - 1);
  auto bar = [&](int i) { return sbase + S.bars + 8u * (uint32_t)i; };

  // ---- one-time setup ----------------------------------------------------------------------
  for (uint32_t i = threadIdx.x * 16; i < S.bars; i += N_THREADS * 16) st_shared_v4(sbase + i, 0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x == 0) {
    // pair: the leader's B_WFULL also counts the peer's "my half has landed" (relay), its B_ACT / B_STAGE the peer's epilogue warps
    for (int i = 0; i < NS; ++i) { mbar_init(bar(B_WFULL + i), (PAIR && rank == 0) ? 2 : 1); mbar_init(bar(B_WEMPTY + i), N_TILES); }
    mbar_init(bar(B_ACC), N_TILES);
    mbar_init(bar(B_ACT), CG * N_EPI_WARPS);
    for (int i = 0; i < N_STAGES; ++i) mbar_init(bar(B_STAGE + i), CG * N_EPI_WARPS);
    mbar_init(bar(B_ZERO), CG * N_EPI_WARPS);
    for (int i = 0; i < N_STAGES; ++i) { mbar_init(bar(B_HALO + i), 1); mbar_init(bar(B_HREL + i), 1); }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < L; i += N_THREADS) {
    st_shared_u16(sbase + S.perm + 2 * i, (uint16_t)a.perm[i]);
    st_shared_u16(sbase + S.inv_perm + 2 * i, (uint16_t)a.inv_perm[i]);
  }
  if (warp == WARP_MMA) tmem_alloc<CG>(sbase + S.tmem_ptr, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();             // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + S.tmem_ptr) : "memory");

  if (warp == WARP_PRODUCER) {
    // ================= weight producer ==================================================================
    if (lane == 0) {
      uint32_t pos = 0, phase = 0;
      for (int un = unit0; un < n_units; un += unit_stride)
        for (int st = 0; st < n_stacks; ++st) {
          const uint8_t* src = a.wimg + (size_t)st * a.stack_bytes + rank * SLOT_B;      // this CTA's half of every slot
          for (int i = 0; i < slots_per_stack; ++i) {
            mbar_wait(bar(B_WEMPTY + pos), phase ^ 1, a.err, 21);
            mbar_arrive_expect_tx(bar(B_WFULL + pos), SLOT_B);
            bulk_g2s(sbase + S.wslot + pos * SLOT_B, src, SLOT_B, bar(B_WFULL + pos));
            src += SLOT_IMG_B;
            if (++pos == NS) { pos = 0; phase ^= 1; }
          }
        }
    }
  } else if (warp >= WARP_MMA && warp < WARP_MMA + N_TILES && rank != 0) {
    // ================= peer CTA: tell the leader that this CTA's half of a slot has landed (written by the async proxy: no fence) ====
    if (warp == WARP_MMA) {
      uint32_t pos = 0, phase = 0;
      for (int un = unit0; un < n_units; un += unit_stride)
        for (int i = 0; i < n_stacks * slots_per_stack; ++i) {
          mbar_wait(bar(B_WFULL + pos), phase, a.err, 26);
          if (elect_one()) mbar_arrive_remote(bar(B_WFULL + pos), 0);
          __syncwarp();
          if (++pos == NS) { pos = 0; phase ^= 1; }
        }
    } else if (STRADDLE) {
      // halo relay: the leader's rows 254, 255 land in this CTA's front halo rows chunk stage by chunk stage (st.async); arm, wait,
      // tell the leader's tile-0 issuer (data written by the async proxy into THIS CTA's shared memory: no fence, like the weights)
      uint32_t ph = 0;
      for (int un = unit0; un < n_units; un += unit_stride)
        for (int i = 0; i < n_stacks * (n_layer - 1); ++i, ph ^= 1)
          for (int c = 0; c < N_STAGES; ++c) {
            if (elect_one()) mbar_arrive_expect_tx(bar(B_HALO + c), halo_stage_bytes(c));
            __syncwarp();
            mbar_wait(bar(B_HALO + c), ph, a.err, 30);
            if (elect_one()) mbar_arrive_remote(bar(B_HREL + c), 0);
            __syncwarp();
          }
    }
  } else if (warp >= WARP_MMA && warp < WARP_MMA + N_TILES) {
    // ================= MMA issuers (leader CTA): warp 8 + m owns tile m (of both CTAs' groups in the pair version); each warp runs
    // converged, one elected lane issues ===========================================================================================
    constexpr uint32_t IDESC = make_idesc_x3(128 * CG, NPAD);
    const int m = warp - WARP_MMA;
    uint32_t pos = 0, wphase = 0, n_act = 0, n_stage = 0;
    // descriptor low words of this tile's operands (start address + LBO); a k-step adds a compile-time constant
    const uint32_t rowoff = (uint32_t)(128 * m) * ROW_B;
    const uint32_t ahi_c = dlo(sbase + S.act_hi + rowoff, CHUNK_B), alo_c = dlo(sbase + S.act_lo + rowoff, CHUNK_B);              // chunk pairs
    const uint32_t ahi_t = dlo(sbase + S.act_hi + 12 * CHUNK_B + rowoff, ROW_B), alo_t = dlo(sbase + S.act_lo + 12 * CHUNK_B + rowoff, ROW_B);   // chunk 12, two taps
    for (int un = unit0; un < n_units; un += unit_stride)
      for (int st = 0; st < n_stacks; ++st) {
        const uint32_t xsel = (uint32_t)(a.enc ? (st == 2) : (st & 1));     // enc: branch 3 reads the interleaved bits
        const uint32_t xhi = dlo(sbase + S.xin_hi[0] + xsel * CHUNK_B + rowoff, ROW_B), xlo = dlo(sbase + S.xin_lo[0] + xsel * CHUNK_B + rowoff, ROW_B);
        for (int layer = 0; layer < n_layer; ++layer) {
          const uint32_t d_tmem = tmem_base + (uint32_t)(layer & 1) * TMEM_BUF_COLS + (uint32_t)(m * NPAD);     // accumulators alternate by layer parity
          if (layer == 0) {
            // the stack input is in place (straddle: rows may have been written by the peer's threads: cluster-scope acquire)
            if (STRADDLE) mbar_wait_cluster(bar(B_ACT), n_act & 1, a.err, 22); else mbar_wait(bar(B_ACT), n_act & 1, a.err, 22);
            ++n_act;
            tc_fence_after();
            // slots: W0_hi (x_hi and x_lo chains), W0_lo (x_hi chain)
#pragma unroll
            for (int part = 0; part < 2; ++part) {
              mbar_wait(bar(B_WFULL + pos), wphase, a.err, 23);
              tc_fence_after();
              const uint32_t wlo = dlo(sbase + S.wslot + pos * SLOT_B, WCHUNK_B);
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < KS_L0; ++ks) {
                  const uint64_t bdesc = dfull(wlo + (uint32_t)(ks * KSTEP_B) / 16);
                  umma_bf16<CG>(d_tmem, dfull(xhi + (uint32_t)(2 * ks)), bdesc, IDESC, (part | ks) != 0);
                  if (part == 0) umma_bf16<CG>(d_tmem, dfull(xlo + (uint32_t)(2 * ks)), bdesc, IDESC, 1);
                }
                commit_x3(bar(B_WEMPTY + pos));
                if (part == 1) commit_x3(bar(B_ACC));
              }
              __syncwarp();
              if (++pos == NS) { pos = 0; wphase ^= 1; }
            }
          } else {
            const uint32_t sphase = n_stage & 1;
            ++n_stage;
#pragma unroll
            for (int part = 0; part < 2; ++part) {
#pragma unroll
              for (int s = 0; s < SLOTS_PASS; ++s) {
                // the chunks this slot's k-steps read have been written by every epilogue warp of the previous layer
                if (part == 0) {
#pragma unroll
                  for (int c = (s == 0 ? 0 : stage_of_slot(s - 1) + 1); c <= stage_of_slot(s); ++c) {
                    mbar_wait(bar(B_STAGE + c), sphase, a.err, 25);
                    if (STRADDLE) {
                      // + the halo rows the neighbour mirrors into the CTA whose tile reads them: tile 1 reads the leader's rows behind
                      // row 255 (armed and waited for here), tile 0 the peer's rows in front of its row 0 (reported by the peer's relay)
                      if (m == 1) {
                        if (elect_one()) mbar_arrive_expect_tx(bar(B_HALO + c), halo_stage_bytes(c));
                        __syncwarp();
                        mbar_wait(bar(B_HALO + c), sphase, a.err, 28);
                      } else {
                        mbar_wait(bar(B_HREL + c), sphase, a.err, 29);
                      }
                    }
                  }
                }
                mbar_wait(bar(B_WFULL + pos), wphase, a.err, 23);
                tc_fence_after();
                const uint32_t wlo = dlo(sbase + S.wslot + pos * SLOT_B, WCHUNK_B);
                if (elect_one()) {
#pragma unroll
                  for (int k3 = 0; k3 < KS_PER_SLOT; ++k3) {
                    const int ks = s * KS_PER_SLOT + k3;
                    // A operand, chunk-major: (chunk pair ks / 5, tap ks % 5) for ks < 30, else chunk 12 with two taps in one k-step
                    const uint32_t add = ks < 30 ? ((uint32_t)(2 * (ks / 5)) * CHUNK_B + (uint32_t)(ks % 5) * ROW_B) / 16 : (uint32_t)(2 * (ks - 30));
                    const uint64_t bdesc = dfull(wlo + (uint32_t)(k3 * KSTEP_B) / 16);
                    umma_bf16<CG>(d_tmem, dfull((ks < 30 ? ahi_c : ahi_t) + add), bdesc, IDESC, (part | s | k3) != 0);
                    if (part == 0) umma_bf16<CG>(d_tmem, dfull((ks < 30 ? alo_c : alo_t) + add), bdesc, IDESC, 1);
                  }
                  commit_x3(bar(B_WEMPTY + pos));
                  if (part == 1 && s == SLOTS_PASS - 1) commit_x3(bar(B_ACC));
                }
                __syncwarp();
                if (++pos == NS) { pos = 0; wphase ^= 1; }
              }
            }
          }
        }
      }
  } else {
    // ================= epilogue warps: one activation row per thread =====================================
    const int tile = warp >> 2, q = warp & 3;
    const int tid = threadIdx.x;                      // 0..255
    const int g_row = 128 * tile + 32 * q + lane;     // row of the group this thread owns in every epilogue
    const uint32_t brow = (uint32_t)(g_row + HALO_LO);
    const uint32_t taddr0 = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(tile * NPAD);
    const uint32_t tab_base = sbase + S.tab;
    uint32_t n_acc = 0, n_zero = 0;
    // straddle: this thread's row in the unit's 512-row space; the two rows either side of the CTA boundary are mirrored into the
    // neighbour's halo rows (rank 0 rows 254, 255 -> peer buffer rows 0, 1; rank 1 rows 0, 1 -> leader buffer rows 258, 259)
    const int u_row = STRADDLE ? 256 * (int)rank + g_row : g_row;
    const bool mirror = STRADDLE && (rank == 0 ? g_row >= 254 : g_row <= 1);
    const uint32_t mirror_brow = rank == 0 ? (uint32_t)(g_row - 254) : (uint32_t)(258 + g_row);
    // store into the stack-input buffers of whichever CTA owns unit row R (+ the neighbour's halo row when R is at the boundary)
    auto owner_addr = [&](uint32_t buf, int R, uint32_t stride, uint32_t off) {
      return dsmem_addr(sbase + buf + (uint32_t)((R & 255) + HALO_LO) * stride + off, (uint32_t)(R >> 8));
    };
    auto halo_addr = [&](uint32_t buf, int R, uint32_t off) {          // only for R in {254, 255, 256, 257}
      return R < 256 ? dsmem_addr(sbase + buf + (uint32_t)(R - 254) * ROW_B + off, 1u) : dsmem_addr(sbase + buf + (uint32_t)(R + HALO_LO) * ROW_B + off, 0u);
    };
    auto put16 = [&](uint32_t buf, int R, uint32_t off, uint16_t v) {
      if (!STRADDLE) { st_shared_u16(sbase + buf + (uint32_t)(R + HALO_LO) * ROW_B + off, v); return; }
      st_cluster_b16(owner_addr(buf, R, ROW_B, off), v);
      if (R >= 254 && R <= 257) st_cluster_b16(halo_addr(buf, R, off), v);
    };
    auto put32 = [&](uint32_t buf, int R, uint32_t off, uint32_t v) {
      if (!STRADDLE) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + buf + (uint32_t)(R + HALO_LO) * ROW_B + off), "r"(v) : "memory"); return; }
      st_cluster_b32(owner_addr(buf, R, ROW_B, off), v);
      if (R >= 254 && R <= 257) st_cluster_b32(halo_addr(buf, R, off), v);
    };
    auto put64 = [&](uint32_t buf, int R, uint32_t off, uint32_t v0, uint32_t v1) {
      if (!STRADDLE) { st_shared_v2(sbase + buf + (uint32_t)(R + HALO_LO) * ROW_B + off, v0, v1); return; }
      st_cluster_v2(owner_addr(buf, R, ROW_B, off), v0, v1);
      if (R >= 254 && R <= 257) st_cluster_v2(halo_addr(buf, R, off), v0, v1);
    };
    auto put_master = [&](uint32_t buf, int R, uint32_t off, float v) {      // fp32 master copy [row][8]: read by the row's own thread only
      if (!STRADDLE) { st_shared_f32(sbase + buf + (uint32_t)(R + HALO_LO) * 32 + off, v); return; }
      st_cluster_b32(owner_addr(buf, R, 32, off), __float_as_uint(v));
    };
    // arrive on the issuers' barrier after writes that may have gone to the PEER's shared memory: cluster-scope release
    auto arrive_issuer_cluster = [&](uint32_t b) { if (STRADDLE) mbar_arrive_cluster(b, 0); else arrive_issuer(b); };
    for (int un = unit0; un < n_units; un += unit_stride) {
      // non-straddle pair: (the odd group of the last pair may not exist: n_cw = 0, nothing is read or written)
      const int cw0 = (STRADDLE ? un : (PAIR ? 2 * un + (int)rank : un)) * a.cw_per_group;
      const int n_cw = max(0, min(a.cw_per_group, a.B - cw0));
      const int g_cw = u_row / CW_ROWS, g_l = u_row - g_cw * CW_ROWS;
      const bool valid = (g_l < L) && (g_cw < n_cw);
      const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;

      // ---- group start: stack inputs (prior channels = 0, decoders.py:227) ------------------------------
      epi_bar_sync();                                 // (every thread is done with the previous group's buffers)
#pragma unroll 1
      for (uint32_t i = tid * 16; i < S.wslot - S.xin_hi[0]; i += N_EPI_THREADS * 16) st_shared_v4(sbase + S.xin_hi[0] + i, 0u, 0u, 0u, 0u);
      epi_bar_sync();
      if (STRADDLE) {
        // both CTAs have zeroed before either writes rows into the other (cluster-scope release / acquire)
        if (lane == 0) { mbar_arrive_cluster(bar(B_ZERO), 0); mbar_arrive_cluster(bar(B_ZERO), 1); }
        mbar_wait_cluster(bar(B_ZERO), n_zero & 1, a.err, 27);
        ++n_zero;
      }
      {
        const int idx = STRADDLE ? 256 * (int)rank + tid : tid;      // one (codeword, position) of the unit per thread
        const int sc = idx / L, sl = idx - sc * L;
        if (idx < n_cw * L) {
          const int R = sc * CW_ROWS + sl;
          const int Ri = sc * CW_ROWS + (int)ld_shared_u16(sbase + S.inv_perm + 2 * sl);
          if (a.enc) {
            const uint16_t x = half_bits(SCALE_X * (2.0f * a.u[(size_t)cw0 * L + idx] - 1.0f));              // encoders.py:362 (+-1: exact)
            put16(S.xin_hi[0], R, 0, x);           // branches 1, 2
            put16(S.xin_hi[1], Ri, 0, x);          // branch 3: x_int[i] = x[p[i]]   (encoders.py:369)
          } else {
            const float* rsrc = a.received + ((size_t)cw0 * L + idx) * 3;
            const float r0 = __ldg(rsrc), r1 = __ldg(rsrc + 1), r2 = __ldg(rsrc + 2);
            uint32_t h01, l01, h2, l2;
            split2(r0, r1, h01, l01);
            split2(r2, 0.f, h2, l2);
            // stack "dec1" input: [r_sys, r_par1, prior...]  (decoders.py:221, 223, 230)
            put_master(S.master[0], R, 0, r0);
            put_master(S.master[0], R, 4, r1);
            put32(S.xin_hi[0], R, 0, h01);
            put32(S.xin_lo[0], R, 0, l01);
            // stack "dec2" input: [r_sys_int, r_par2, x_plr_int...], r_sys_int[i] = r_sys[p[i]]   (decoders.py:222, 224, 240)
            put_master(S.master[1], Ri, 0, r0);
            put_master(S.master[1], R, 4, r2);
            put16(S.xin_hi[1], Ri, 0, (uint16_t)(h01 & 0xFFFFu));
            put16(S.xin_lo[1], Ri, 0, (uint16_t)(l01 & 0xFFFFu));
            put16(S.xin_hi[1], R, 2, (uint16_t)(h2 & 0xFFFFu));
            put16(S.xin_lo[1], R, 2, (uint16_t)(l2 & 0xFFFFu));
          }
        }
      }

      for (int st = 0; st < n_stacks; ++st) {
        const int fout = a.lay.fout(st);
        const bool last_stack = (st == n_stacks - 1);
        // ---- fp32 tables (conv biases, Linear weights and bias), double-buffered by stack parity: this stack's were staged
        //      one stack ago (the first one at the group start), the next stack's are staged now, while layer 0's MMAs run ------
        const uint32_t tab = tab_base + (uint32_t)(st & 1) * (TAB_FLOATS * 4);
        auto stage_tables = [&](int ts) {
          const uint32_t tt = tab_base + (uint32_t)(ts & 1) * (TAB_FLOATS * 4);
          const int tf = a.lay.fout(ts);
          for (int i = tid; i < n_layer * NPAD; i += N_EPI_THREADS) {
            const int j = i / NPAD, c = i - j * NPAD;
            st_shared_f32(tt + 4u * (uint32_t)(TAB_BIAS + i), c < a.lay.units ? __ldg(a.params + a.lay.conv_b(ts, j) + c) : 0.f);
          }
          for (int i = tid; i < MAX_F * NPAD; i += N_EPI_THREADS) {
            const int f = i / NPAD, c = i - f * NPAD;
            st_shared_f32(tt + 4u * (uint32_t)(TAB_V + i), (f < tf && c < a.lay.units) ? __ldg(a.params + a.lay.lin_w(ts) + (size_t)f * a.lay.units + c) : 0.f);
          }
          if (tid < 8) st_shared_f32(tt + 4u * (uint32_t)(TAB_C + tid), tid < tf ? __ldg(a.params + a.lay.lin_b(ts) + tid) : 0.f);
        };
        if (st == 0) {
          stage_tables(0);
          if (STRADDLE) fence_proxy_async_all(); else fence_proxy_async();      // the stack inputs were written with generic stores
          epi_bar_sync();
          if (lane == 0) arrive_issuer_cluster(bar(B_ACT));
        } else {
          epi_bar_sync();                              // every warp is done with stack st-1 (its tables' buffer is reused below) and sees this stack's tables
        }
        if (st + 1 < n_stacks) stage_tables(st + 1);

        for (int layer = 0; layer < n_layer; ++layer) {
          const bool lin_layer = (layer == n_layer - 1);
          mbar_wait(bar(B_ACC), n_acc & 1, a.err, 24);
          ++n_acc;
          tc_fence_after();
          const uint32_t btab = tab + 4u * (uint32_t)(TAB_BIAS + layer * NPAD);
          const uint32_t taddr = taddr0 + (uint32_t)(layer & 1) * TMEM_BUF_COLS;
          float lin[MAX_F] = {0.f, 0.f, 0.f, 0.f, 0.f};
          uint32_t ra[16], rb[16];
          tmem_ld16(taddr, ra);
#pragma unroll
          for (int cb = 0; cb < 7; ++cb) {
            tmem_ld_wait();
            uint32_t* cur = (cb & 1) ? rb : ra;
            if (cb < 6) tmem_ld16(taddr + 16 * (cb + 1), (cb & 1) ? ra : rb);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int c = 2 * cb + half;
              if (c >= N_CHUNKS) break;
              const float4 b0 = ld_shared_f4(btab + (uint32_t)(c * 32)), b1 = ld_shared_f4(btab + (uint32_t)(c * 32 + 16));
              float v[8];
              v[0] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 0]), ACC_INV, b0.x));
              v[1] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 1]), ACC_INV, b0.y));
              v[2] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 2]), ACC_INV, b0.z));
              v[3] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 3]), ACC_INV, b0.w));
              v[4] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 4]), ACC_INV, b1.x));
              v[5] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 5]), ACC_INV, b1.y));
              v[6] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 6]), ACC_INV, b1.z));
              v[7] = elu_x3(fmaf(__uint_as_float(cur[8 * half + 7]), ACC_INV, b1.w));
              if (!lin_layer) {
                uint32_t h[4], l[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) split2(v[2 * j], v[2 * j + 1], h[j], l[j]);
                st_shared_v4(sbase + S.act_hi + (uint32_t)c * CHUNK_B + brow * ROW_B, h[0] & keep, h[1] & keep, h[2] & keep, h[3] & keep);
                st_shared_v4(sbase + S.act_lo + (uint32_t)c * CHUNK_B + brow * ROW_B, l[0] & keep, l[1] & keep, l[2] & keep, l[3] & keep);
                if (mirror) {       // the neighbour CTA's taps read this row as a halo row: async-proxy store + complete_tx on ITS B_HALO[cb]
                  const uint32_t rbar = dsmem_addr(bar(B_HALO + cb), rank ^ 1u);
                  st_async_v4(dsmem_addr(sbase + S.act_hi + (uint32_t)c * CHUNK_B + mirror_brow * ROW_B, rank ^ 1u), rbar, h[0] & keep, h[1] & keep, h[2] & keep, h[3] & keep);
                  st_async_v4(dsmem_addr(sbase + S.act_lo + (uint32_t)c * CHUNK_B + mirror_brow * ROW_B, rank ^ 1u), rbar, l[0] & keep, l[1] & keep, l[2] & keep, l[3] & keep);
                }
              } else {
                // Linear of the stack output, in fp32 straight from the accumulators (decoders.py:232, 243; encoders.py:364)
#pragma unroll
                for (int f = 0; f < MAX_F; ++f) {
                  if (f >= fout) break;
                  const float4 w0 = ld_shared_f4(tab + 4u * (uint32_t)(TAB_V + f * NPAD + c * 8)), w1 = ld_shared_f4(tab + 4u * (uint32_t)(TAB_V + f * NPAD + c * 8 + 4));
                  float s = lin[f];
                  s = fmaf(w0.x, v[0], s); s = fmaf(w0.y, v[1], s); s = fmaf(w0.z, v[2], s); s = fmaf(w0.w, v[3], s);
                  s = fmaf(w1.x, v[4], s); s = fmaf(w1.y, v[5], s); s = fmaf(w1.z, v[6], s); s = fmaf(w1.w, v[7], s);
                  lin[f] = s;
                }
              }
            }
            if (!lin_layer) {
              // stage cb of this layer's output is complete for this warp: the next layer's MMAs on these chunks may start
              fence_proxy_async();
              if (cb == 6) tc_fence_before();           // (the accumulators this warp read are drained before the last arrival)
              __syncwarp();
              if (lane == 0) arrive_issuer(bar(B_STAGE + cb));
            }
          }
          if (lin_layer) {
            const float4 c0 = ld_shared_f4(tab + 4u * (uint32_t)TAB_C);
            const float c4 = ld_shared_f32(tab + 4u * (uint32_t)(TAB_C + 4));
            lin[0] += c0.x; lin[1] += c0.y; lin[2] += c0.z; lin[3] += c0.w; lin[4] += c4;
            if (a.enc) {
              // x_tx[:, :, branch] = ELU(Linear(h)) (encoders.py:364, 367, 371) + the power sums
              double s1 = 0.0, s2 = 0.0;
              if (valid) {
                const float z = lin[0];
                const float v = z > 0.f ? z : expm1f(z);
                a.x_tx[((size_t)(cw0 + g_cw) * L + g_l) * 3 + st] = v;
                s1 = (double)v;
                s2 = (double)v * (double)v;
              }
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, d);
                s2 += __shfl_xor_sync(0xffffffffu, s2, d);
              }
              if (lane == 0 && (s1 != 0.0 || s2 != 0.0)) { atomicAdd(a.stats + 0, s1); atomicAdd(a.stats + 1, s2); }
            } else if (valid) {
              const int cw = cw0 + g_cw;
              if (a.trace) {
                float* tr = a.trace + (((size_t)st * a.B + cw) * L + g_l) * F;
                const int nf = last_stack ? 1 : F;
#pragma unroll
                for (int f = 0; f < MAX_F; ++f)
                  if (f < nf) tr[f] = lin[f];
              }
              // where position l lands: interleave after dec1 (x_int[i] = x[p[i]]: l -> rp[l]), de-interleave after dec2 (l -> p[l])
              const uint32_t dl = ld_shared_u16(sbase + ((st & 1) ? S.perm : S.inv_perm) + 2 * g_l);
              if (last_stack) {
                a.out[(size_t)cw * L + dl] = 1.f / (1.f + __expf(-lin[0]));                     // decoders.py:267
              } else {
                const uint32_t cur_m = sbase + S.master[st & 1] + brow * 32, nxt = (uint32_t)((st & 1) ^ 1);
                const int dR = g_cw * CW_ROWS + (int)dl;             // destination row in the unit's row space
                float e[6];
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                  const float prior = a.extrinsic ? ld_shared_f32(cur_m + 8 + 4 * f) : 0.f;     // decoders.py:235-236, 246-247
                  e[f] = f < F ? lin[f] - prior : 0.f;
                  put_master(S.master[nxt], dR, 8 + 4 * f, e[f]);
                }
                e[5] = 0.f;
                uint32_t h0, l0, h1, l1, h2, l2;
                split2(e[0], e[1], h0, l0);
                split2(e[2], e[3], h1, l1);
                split2(e[4], e[5], h2, l2);
                put32(S.xin_hi[nxt], dR, 4, h0);
                put64(S.xin_hi[nxt], dR, 8, h1, h2);
                put32(S.xin_lo[nxt], dR, 4, l0);
                put64(S.xin_lo[nxt], dR, 8, l1, l2);
              }
            }
          }
          // the next stack's first layer may start (its inputs were scattered by ALL warps): every warp reports on its own; the
          // group's very last epilogue is followed by the next group's start, which reports instead
          if (lin_layer && !last_stack) {
            if (STRADDLE) fence_proxy_async_all(); else fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_issuer_cluster(bar(B_ACT));
          }
        }
      }
      tc_fence_before();
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();             // the peer's shared memory and barriers stay alive until the leader's last commit has landed
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, TMEM_COLS);
  }
}

uint32_t stack_bytes(int n_layer) { return (uint32_t)(2 + 2 * SLOTS_PASS * (n_layer - 1)) * SLOT_IMG_B; }

bool supported(int L, int n_layer, int units, int k, int F, const char** why) {
  static thread_local char msg[160];
  *why = msg;
  if (k != TAPS) { snprintf(msg, sizeof msg, "kernel_size %d (only 5 is built for the tensor paths)", k); return false; }
  if (units < 1 || units > UNITS_MAX) { snprintf(msg, sizeof msg, "num_unit %d > %d", units, UNITS_MAX); return false; }
  if (F < 1 || F > MAX_F) { snprintf(msg, sizeof msg, "num_iter_ft %d > %d", F, MAX_F); return false; }
  if (n_layer < 1 || n_layer > MAX_LAYER) { snprintf(msg, sizeof msg, "num_layer %d outside 1..%d", n_layer, MAX_LAYER); return false; }
  if (L < 1 || L + 2 > UNIT_ROWS) { snprintf(msg, sizeof msg, "block_len %d > %d (one codeword must fit the %d-row space of a work unit)", L, UNIT_ROWS - 2, UNIT_ROWS); return false; }
  *why = nullptr;
  return true;
}

int launch_setup(int* n_sm_out) {
  static DeviceOnce once;
  return device_once(once, "x3_kernel", [](int dev) -> int {
    int rc = require_sm100(dev, "the f16x3 tensor path");
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)make_smem().total);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(x3_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    return TAE_OK;
  }, n_sm_out);
}

int pack(const Layout& lay, const float* params, void* packed, cudaStream_t s) {
  const uint32_t stack_elems = stack_bytes(lay.n_layer) / 2;
  const size_t total = (size_t)lay.n_stacks * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  pack_x3_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<uint16_t*>(packed), lay, stack_elems);
  return after_launch("pack_x3_kernel");
}

int launch(Args& a, void* ws, size_t ws_bytes, cudaStream_t s, const char* who) {
  if (ws_bytes < 256) { set_error("%s: workspace %zu < 256 bytes", who, ws_bytes); return TAE_EWORKSPACE; }
  int n_sm = 0;
  int rc = launch_setup(&n_sm);
  if (rc) return rc;
  a.err = wait_code_slot(ws);
  a.cw_per_group = (UNIT_ROWS + 2) / (a.L + 2);       // codewords per 256-row group, or per 512-row unit when straddling
  a.n_groups = (a.B + a.cw_per_group - 1) / a.cw_per_group;
  a.n_units = (PAIR && !STRADDLE) ? (a.n_groups + 1) / 2 : a.n_groups;
  a.stack_bytes = stack_bytes(a.n_layer);
  const int grid = PAIR ? 2 * std::min(a.n_units, n_sm / 2) : std::min(a.n_units, n_sm);
  x3_kernel<<<grid, N_THREADS, make_smem().total, s>>>(a);
  return after_launch(who);
}

}  // namespace x3
}  // namespace

bool dec_x3_supported(const TaeDecConfig& c, const char** why) {
  return x3::supported(c.block_len, c.num_layer, c.num_unit, c.kernel_size, c.num_iter_ft, why);
}
size_t dec_x3_packed_bytes(const TaeDecConfig& c) { return (size_t)2 * c.num_iteration * x3::stack_bytes(c.num_layer); }
int dec_x3_pack(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s) {
  const x3::Layout lay{2 * c.num_iteration, c.num_layer, c.num_unit, 2 + c.num_iter_ft, c.num_iter_ft, 1};
  return x3::pack(lay, params, packed, s);
}
int dec_forward_x3(const TaeDecConfig& c, const float* params, const void* packed, const float* received, const int32_t* perm,
                   const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes, cudaStream_t s) {
  x3::Args a{};
  a.wimg = reinterpret_cast<const uint8_t*>(packed);
  a.params = params; a.received = received; a.out = out; a.trace = trace; a.perm = perm; a.inv_perm = inv_perm;
  a.B = B; a.L = c.block_len; a.F = c.num_iter_ft; a.n_stacks = 2 * c.num_iteration; a.n_layer = c.num_layer;
  a.extrinsic = c.extrinsic; a.enc = 0;
  a.lay = x3::Layout{a.n_stacks, c.num_layer, c.num_unit, 2 + c.num_iter_ft, c.num_iter_ft, 1};
  return x3::launch(a, ws, ws_bytes, s, "x3_kernel(dec)");
}

bool enc_x3_supported(const TaeEncConfig& c, const char** why) {
  return x3::supported(c.block_len, c.num_layer, c.num_unit, c.kernel_size, 1, why);
}
size_t enc_x3_packed_bytes(const TaeEncConfig& c) { return (size_t)3 * x3::stack_bytes(c.num_layer); }
int enc_x3_pack(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s) {
  const x3::Layout lay{3, c.num_layer, c.num_unit, 1, 1, 1};
  return x3::pack(lay, params, packed, s);
}
int enc_forward_x3(const TaeEncConfig& c, const float* params, const void* packed, const float* u, const int32_t* perm,
                   const int32_t* inv_perm, float* x_tx, double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s) {
  x3::Args a{};
  a.wimg = reinterpret_cast<const uint8_t*>(packed);
  a.params = params; a.u = u; a.x_tx = x_tx; a.stats = stats; a.perm = perm; a.inv_perm = inv_perm;
  a.B = B; a.L = c.block_len; a.F = 1; a.n_stacks = 3; a.n_layer = c.num_layer; a.extrinsic = 0; a.enc = 1;
  a.lay = x3::Layout{3, c.num_layer, c.num_unit, 1, 1, 1};
  return x3::launch(a, ws, ws_bytes, s, "x3_kernel(enc)");
}

}  // namespace tae
