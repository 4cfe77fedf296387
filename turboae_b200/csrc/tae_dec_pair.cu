// Fused DEC_LargeCNN.forward on 5th-gen tensor cores, CTA-pair version (TAE_PRECISION_BF16), sm_100a only.
//
// Reference arithmetic restated (paths relative to the reference checkout):
//   decoders.py:219-269 (turbo schedule), cnn_utils.py:36-46 (conv + ELU stack),
//   interleavers.py:15-21, 43-48 (row permutations).
//
// Execution model
//   * cluster of 2 CTAs (one per SM of a TPC), persistent; each CTA owns one "group" = a 512-row activation
//     buffer holding floor(514/(L+2)) codewords, each followed by 2 all-zero separator rows (= the zero
//     padding of cnn_utils.py:16 for free).  4 MMA tiles of 128 rows per CTA.
//   * every conv layer is tcgen05.mma.cta_group::2, M = 256 (tile m of both CTAs), N = 112 (100 output
//     channels, padded).  Each CTA stages only ITS half of the weight columns (56 of 112), so a whole layer
//     (57 KB per CTA) is resident while the four tiles run through it.  One issuer warp per tile (leader CTA); the
//     units->units tiles are issued in program order with two in flight (B_ISS chain: one issuing thread alone reaches
//     ~67 cycles per MMA, two interleaved streams the nominal 56), so the epilogue of tile m overlaps the MMAs of tiles
//     m+1, m+2 and the next layer's tile 0 starts as soon as tiles 0 and 1 of this layer are written back.
//   * K is packed to exactly 32 k-steps of 16: 5 taps x 96 channels = 30 k-steps from the canonical
//     no-swizzle K-major activation layout [C/8][rows][8] (tap t = descriptor start address + 16*t bytes), and
//     channels 96..99 live in a "combined" chunk whose row r holds [x[r][96..99], x[r+1][96..99]], so that one
//     16-byte chunk carries two taps; the last k-step pairs tap 4 of those channels with a constant-one chunk
//     that injects the bias (bf16 hi + lo split) -- no output channel, no epilogue add.
//   * activations are updated IN PLACE: the epilogue of tile m may not touch the two rows that tile m+1's MMAs
//     still read (its last two), so the two lanes owning them park their packed outputs in a 416-byte shared-memory
//     scratch and move them into place at the start of the next tile's epilogue (B_DEF tells the Linear of tile m
//     that they are there).  Epilogue warps report per warp (no CTA-wide barrier).
//   * inference (MODE 0) keeps every activation scaled by log2(e) (the weight image absorbs the factor), so the ELU of
//     the epilogue is max(z', fma(ex2(-|z'|), log2e, -log2e)): MUFU + FFMA + FMNMX.
//   * priors stay in shared memory as the bf16 channels 2.. of the stack inputs for all 2*I half-iterations;
//     (de)interleave is the row index of the Linear epilogue's store (one tile per 4 warps: column part t owns tile t),
//     and the first layer of the next stack starts on the tiles whose codewords are complete.  HBM sees `received` once
//     (pulled into L2 one stack ahead) and `out` once; weights stream from L2 once per CTA pair and layer through a
//     bulk-copy (UBLKCP) ring.
//   * what bounds it (profiles/r02_dec_schedule_variants.md): inside a layer both the tensor pipe (7 168 cycles) and the
//     shared-memory port (105 B/clk of MMA operands + weight ring + epilogue stores = 7 056 cycles at 128 B/clk) are
//     saturated; the stack boundary (Linear -> scatter -> first layer) is a chain of barrier hops of ~9 k cycles.
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "tae_common.cuh"
#include "tae_umma.cuh"

#ifndef TAE_TIMELINE
#define TAE_TIMELINE 0      // 1: record clock64 stamps (scripts/dec_timeline.py); costs code size, keep off in production
#endif

namespace tae {

namespace {

// ------------------------------------------------------------------------------------------
// geometry
// ------------------------------------------------------------------------------------------
constexpr int GROUP_ROWS = 512;
constexpr int BUF_ROWS = GROUP_ROWS + 4;          // 2 halo rows in front and behind
constexpr int N_TILES = 4;
constexpr int NPAD = 112;                         // UMMA N of a conv layer (both CTAs together)
constexpr int NHALF = NPAD / 2;                   // weight columns staged per CTA
constexpr int LIN_N = 16, LIN_NHALF = 8;
constexpr int N_REG_CH = 96;                      // channels in regular chunks
constexpr int N_REG_CHUNKS = N_REG_CH / 8;        // 12
constexpr int UNITS_MAX = 100;
constexpr int TAPS = 5;
constexpr int KS_CONV = 32;                       // k-steps of a units->units layer
constexpr int KS_L0 = 3;                          // k-steps of the (2+F)->units layer
constexpr int KS_LIN = 7;                         // k-steps of the Linear
constexpr int KS_PER_SLOT = 4;
constexpr int SLOTS_CONV = KS_CONV / KS_PER_SLOT; // 8
constexpr uint32_t ROW_B = 16;
constexpr uint32_t CHUNK_B = BUF_ROWS * ROW_B;                  // 8256
constexpr uint32_t WCHUNK_B = NHALF * ROW_B;                    // 896: one 8-channel K chunk of 56 weight columns
constexpr uint32_t SLOT_B = KS_PER_SLOT * 2 * WCHUNK_B;         // 7168
constexpr uint32_t L0_B = KS_L0 * 2 * WCHUNK_B;                 // 5376
constexpr uint32_t LIN_WCHUNK_B = LIN_NHALF * ROW_B;            // 128
constexpr uint32_t LIN_B = KS_LIN * 2 * LIN_WCHUNK_B;           // 1792
constexpr int KS_FIN_SLOT = 16;                                 // backward: the units -> (2+F) transposed conv, 2 slots of 16 k-steps
constexpr uint32_t FIN_B = KS_FIN_SLOT * 2 * LIN_WCHUNK_B;      // 4096
constexpr int IMG_CHUNKS = 13;                                  // chunks of a stashed group image (channels 0..103)
constexpr int NS = 12;                                          // weight ring slots (a layer uses 8; 4 are prefetch headroom)
#ifndef TAE_EPI_WARPS
#define TAE_EPI_WARPS 16
#endif
constexpr int N_EPI_WARPS = TAE_EPI_WARPS;                      // 4 lane quadrants x PARTS column parts (8 or 16)
constexpr int PARTS = N_EPI_WARPS / 4;                          // column parts per lane quadrant
static_assert(PARTS == N_TILES, "the per-tile Linear epilogue maps tile t to the warps of column part t");
constexpr int CPP = N_REG_CHUNKS / PARTS;                       // regular 8-channel chunks per part
constexpr int N_EPI_THREADS = N_EPI_WARPS * 32;
// Warp roles by index.  The SM sub-partition arbiter prefers the highest warp id among the eligible warps: the MMA issuers
// and the weight producer (few instructions, all of them on a critical path) get the highest ids, the epilogue warps
// (instruction-heavy) the lowest.
constexpr int EPI_WARP0 = 0;                                    // warps 0..15: warp & 3 = TMEM lane quadrant
constexpr int WARP_MMA = N_EPI_WARPS;                           // 4 warps of the leader: one MMA issuer per tile
constexpr int WARP_PRODUCER = N_EPI_WARPS + N_TILES;
constexpr int N_THREADS = 32 * (1 + N_TILES + N_EPI_WARPS);      // 672
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t TMEM_LIN_COL = N_TILES * NPAD;               // 448


struct Smem {
  uint32_t act, comb, xin[2], ones, wslot, perm, inv_perm, defer, bars, tmem_ptr, total;
};
// barrier slots (8 bytes each).  Waits are parity waits, so every barrier below has ONE class of waiters that observes EVERY
// one of its phases, and a phase cannot complete before those waiters have seen the previous one (a waiter that skipped a phase
// could take "two phases behind" for "complete", one that fell two phases back would wait forever):
// B_WFULL[p]: ring slot p resident in BOTH CTAs (leader: own bulk copy + one arrival forwarded by the peer's relay, so an
//   MMA issuer does ONE wait per slot).  B_WEMPTY[p]: the MMAs of all four tiles reading slot p have completed (one commit
//   per tile issuer, multicast to both CTAs).
// B_ACC[m]: accumulators of tile m complete (commit, multicast), one phase per step whose epilogue ALL epilogue warps run
//   (conv-type layers; also the encoder's Linear and the backward's last step).  B_LACC[m]: the decoder's Linear of tile m
//   (one phase per stack; waited for by the 4 warps of column part m only).
// B_ACT[m] (leader): tile m written back by all 16 epilogue warps of the pair, one phase per conv-type epilogue; waited for by
//   the issuer of tile m before every step that reads the activations.  B_HALO[m] (leader): same arrivals, for the issuer of
//   tile m-1 (whose taps 3, 4 read the first two rows of tile m): one phase per epilogue whose output a tap-shifted step reads.
// B_LE[m] (leader): the input rows of tile m of the NEXT stack's first layer are in place (group start; Linear epilogue of
//   tile m; backward: stack prologue), one phase per stack; waited for by the issuers of the tiles whose codewords live in m.
// B_DEF[m] (leader): the two deferred rows of tile m have been stored (start of the epilogue of tile m+1 in the last conv-type
//   layer): the Linear of tile m (centre tap only) waits for this instead of the whole epilogue of tile m+1.
// B_ISS[m] (leader): the issuer of tile m has issued the first half of its MMAs of a units->units step: those tiles are
//   issued in program order, two in flight (tile 0..3 of a layer, then the next layer).
enum { B_WFULL = 0, B_WEMPTY = NS, B_ACC = 2 * NS, B_LACC = 2 * NS + 4, B_ACT = 2 * NS + 8, B_HALO = 2 * NS + 12, B_LE = 2 * NS + 16,
       B_DEF = 2 * NS + 20, B_ISS = 2 * NS + 24, N_BARS = 2 * NS + 28 };

__host__ __device__ inline Smem make_smem(int F) {
  Smem s{};
  uint32_t o = 0;
  s.act = o; o += N_REG_CHUNKS * CHUNK_B;
  s.comb = o; o += CHUNK_B;
  s.xin[0] = o; o += CHUNK_B;
  s.xin[1] = o; o += CHUNK_B;
  s.ones = o; o += CHUNK_B;                        // after comb and xin: positive LBO towards it
  (void)F;
  s.wslot = o; o += NS * SLOT_B;
  s.perm = o; o += 1024;
  s.inv_perm = o; o += 1024;
  s.defer = o; o += 2 * (N_REG_CHUNKS + 1) * ROW_B;     // the two deferred rows of a tile, [lane 30 | 31][chunk][16 B] (see the epilogue)
  s.bars = o; o += N_BARS * 8;
  s.tmem_ptr = o; o += 16;
  s.total = o;
  return s;
}

// Flat-parameter layout of a sequence of conv stacks + Linear (dec_layout / enc_layout of tae_common.cuh), passed BY VALUE:
// only the last stack's Linear may have a different width, so every offset is arithmetic (no device table, no copy).
struct PackLayout {
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

struct PairArgs {
  const uint8_t* wimg;
  const float* received;
  float* out;
  float* trace;
  int* err;
  const int32_t* perm;
  const int32_t* inv_perm;
  int B, L, F, n_stacks, n_layer, extrinsic, n_groups, n_pairs, cw_per_group;
  int pair_begin;              // this launch walks the work units [pair_begin, n_pairs) (0 except for a range-split backward)
  int enc;                     // 0: DEC_LargeCNN schedule; 1: ENC_interCNN (three branches, Linear(units,1) + ELU, power sums)
  const float* u;              // enc: bits (B, L, 1)
  float* x_tx;                 // enc: un-normalised codes (B, L, 3)
  double* stats;               // enc: running (sum, sum of squares)
  uint32_t stack_bytes;        // bytes of one stack's image (both halves)
  unsigned long long* tl;      // optional timeline buffer (tae_debug_set_timeline): clock64 stamps of cluster 0, leader CTA
  // training (group images in HBM, bf16 [group][chunk][516 rows][8 channels]; see tae_wgrad.cu)
  uint8_t* stash_y;            // MODE 0: written, [stack][layer][group][13 chunks]; MODE 1: read, this stack's [layer][group][13]
  uint8_t* stash_x;            // MODE 0: the stack inputs, [stack][group][1 chunk]; MODE 1: the dlin image of this stack, [group][1]
  uint8_t* stash_g;            // MODE 1: written, gradients at the pre-activations, [layer][group][13 chunks]
  const float* dlin;           // MODE 1: gradient w.r.t. the Linear output (B, L, fin)
  float* dxin;                 // MODE 1: gradient w.r.t. every stack's input, (n_stacks, B, L, 8)
  // MODE 1 runs the stacks n_stacks-1 .. 0 in one launch.  chain == 0 (encoder branches): stack s reads dlin + s*B*L*fout(s).
  // chain == 1 (turbo schedule): the last stack reads dlin (B, L, fout); every other stack s derives its dlin from the outputs of
  // stack s+1 -- dlin_s[b, l, f] = dxin_{s+1}[b, idx[l], 2 + f] - (sub ? dlin_{s+1}[b, idx[l], f] : 0), the extrinsic subtraction
  // and the (de)interleaver backwards (decoders.py:235-249; idx = inv_perm for odd s+1, perm for even; sub = extrinsic and s+1
  // not the last stack) -- and records it in dlin_all (n_stacks, B, L, F) for stack s-1.
  int chain;
  float* dlin_all;
  float* grad_flat;            // nullptr or the flat gradient buffer: the Linear bias gradients (sum of dlin) are added at lay.lin_b(s)
  PackLayout lay;              // flat-parameter layout (fout per stack, bias offsets)
};

// ELU'(z) from the bf16 forward output y packed two per word: y + 1 where y < 0, else 1 (cnn_utils.py:24-25 backward)
__device__ __forceinline__ float elu_grad_lo(uint32_t ypair) {
  const float y = __uint_as_float(ypair << 16);
  return y < 0.f ? y + 1.f : 1.f;
}
__device__ __forceinline__ float elu_grad_hi(uint32_t ypair) {
  const float y = __uint_as_float(ypair & 0xFFFF0000u);
  return y < 0.f ? y + 1.f : 1.f;
}
constexpr float LOG2E_F = 1.4426950408889634f, LN2_F = 0.6931471805599453f;

// ELU of the epilogue.  SCALED (inference image): input z' = log2e * z, output log2e * ELU(z) = max(z', log2e * (2^-|z'| - 1)):
// for z' > 0 the second term is negative, for z' < 0 it is >= z' (e^z - 1 >= z).  Plain: elu_fast (cnn_utils.py:24-25).
template <bool SCALED>
__device__ __forceinline__ float elu_act(float v) {
  if (SCALED) return fmaxf(v, fmaf(fast_exp2(-fabsf(v)), LOG2E_F, -LOG2E_F));
  return elu_fast(v);
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_THREADS) : "memory"); }

// ------------------------------------------------------------------------------------------
// weight image (bf16), per stack:  [slot][cta half][bytes];  slot 0 = layer 0 (3 k-steps), then
// (n_layer-1) x 8 slots of 4 k-steps, then the Linear (7 k-steps, 8 columns per half).
// One k-step of a half = 2 chunks [columns of this half][8 K elements].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// K element e (0..15) of k-step ks of a units->units layer  ->  (channel, tap) or bias / zero
// Weight scaling of the inference image (`scaled`): every activation is kept as a' = log2(e) * a, so that the ELU of the
// epilogue is max(z', fma(ex2(-|z'|), log2e, -log2e)) -- one MUFU, one FFMA, one FMNMX, no multiply by log2(e) and no select:
//   first layer   z' = log2e * (W x + b)            -> W * log2e, b * log2e
//   other layers  z' = W a' + log2e * b             -> W,          b * log2e
//   Linear        y  = (ln2 * V) a' + c             -> V * ln2,    c
__device__ __forceinline__ float conv_w_elem(const float* __restrict__ w, const float* __restrict__ b, int units, int o,
                                             int ks, int e, float sb) {
  if (o >= units) return 0.f;
  if (ks < 30) {
    const int t = ks / 6, c = 16 * (ks % 6) + e;
    return c < units ? w[((size_t)o * units + c) * TAPS + t] : 0.f;
  }
  if (ks == 30) {
    const int t = e >> 2, c = N_REG_CH + (e & 3);
    return c < units ? w[((size_t)o * units + c) * TAPS + t] : 0.f;
  }
  if (e < 4) {
    const int c = N_REG_CH + e;
    return c < units ? w[((size_t)o * units + c) * TAPS + 4] : 0.f;
  }
  if (e == 8) return bf16_round(sb * b[o]);
  if (e == 9) return sb * b[o] - bf16_round(sb * b[o]);
  return 0.f;
}
__device__ __forceinline__ float l0_w_elem(const float* __restrict__ w, const float* __restrict__ b, int units, int cin,
                                           int o, int ks, int e, float sw) {
  if (o >= units) return 0.f;
  const int t = 2 * ks + (e >> 3), c = e & 7;
  if (t < TAPS) return c < cin ? sw * w[((size_t)o * cin + c) * TAPS + t] : 0.f;
  if (e == 8) return bf16_round(sw * b[o]);
  if (e == 9) return sw * b[o] - bf16_round(sw * b[o]);
  return 0.f;
}
__device__ __forceinline__ float lin_w_elem(const float* __restrict__ w, const float* __restrict__ b, int units, int fout,
                                            int o, int ks, int e, float sw) {
  if (o >= fout) return 0.f;
  if (ks < 6) {
    const int c = 16 * ks + e;
    return c < units ? sw * w[(size_t)o * units + c] : 0.f;
  }
  if (e < 4) {
    const int c = N_REG_CH + e;
    return c < units ? sw * w[(size_t)o * units + c] : 0.f;
  }
  if (e == 8) return bf16_round(b[o]);
  if (e == 9) return b[o] - bf16_round(b[o]);
  return 0.f;
}

__global__ void pack_pair_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ img,
                                 const PackLayout lay, int n_stacks, int n_layer, int units, int cin0,
                                 uint32_t stack_elems, int scaled) {
  const float s_act = scaled ? LOG2E_F : 1.f, s_lin = scaled ? LN2_F : 1.f;
  const size_t total = (size_t)n_stacks * stack_elems;
  const uint32_t l0_elems = 2 * L0_B / 2;
  const uint32_t conv_elems = (uint32_t)(n_layer - 1) * SLOTS_CONV * (2 * SLOT_B / 2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int st = (int)(idx / stack_elems);
    uint32_t r = (uint32_t)(idx % stack_elems);
    float v;
    if (r < l0_elems) {
      const int half = r / (L0_B / 2);
      r %= (L0_B / 2);
      const int ks = r / (2 * NHALF * 8), ch = (r / (NHALF * 8)) & 1, n = (r / 8) % NHALF, e8 = r % 8;
      v = l0_w_elem(params + lay.conv_w(st, 0), params + lay.conv_b(st, 0), units, cin0, half * NHALF + n, ks, ch * 8 + e8, s_act);
    } else if (r < l0_elems + conv_elems) {
      r -= l0_elems;
      const uint32_t slot_elems = 2 * SLOT_B / 2;      // both halves of one slot
      const int sl = r / slot_elems;                   // slot within the stack's units->units layers
      r %= slot_elems;
      const int half = r / (SLOT_B / 2);
      r %= (SLOT_B / 2);
      const int j = 1 + sl / SLOTS_CONV;
      const int ks = (sl % SLOTS_CONV) * KS_PER_SLOT + r / (2 * NHALF * 8);
      const int ch = (r / (NHALF * 8)) & 1, n = (r / 8) % NHALF, e8 = r % 8;
      v = conv_w_elem(params + lay.conv_w(st, j), params + lay.conv_b(st, j), units, half * NHALF + n, ks, ch * 8 + e8, s_act);
    } else {
      r -= l0_elems + conv_elems;
      const int half = r / (LIN_B / 2);
      r %= (LIN_B / 2);
      const int ks = r / (2 * LIN_NHALF * 8), ch = (r / (LIN_NHALF * 8)) & 1, n = (r / 8) % LIN_NHALF, e8 = r % 8;
      v = lin_w_elem(params + lay.lin_w(st), params + lay.lin_b(st), units, lay.fout(st), half * LIN_NHALF + n, ks, ch * 8 + e8, s_lin);
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

// K element e (0..15) of k-step ks of a units->units operand sequence -> (channel, tap); false: bias / padding element
__device__ __forceinline__ bool conv_k_map(int ks, int e, int* c, int* t) {
  if (ks < 30) { *t = ks / 6; *c = 16 * (ks % 6) + e; return true; }
  if (ks == 30) { *t = e >> 2; *c = N_REG_CH + (e & 3); return true; }
  if (e < 4) { *t = 4; *c = N_REG_CH + e; return true; }
  return false;
}

// Backward weight image of one stack (MODE 1 of the fused kernel), same slot structure as the forward image:
//   slot 0   : the transposed Linear as a "layer 0" (F -> units): only the centre tap is non-zero, no bias;
//   8 slots per units->units layer, layers in REVERSE order, weights transposed and tap-flipped:
//              dx[l, c] = sum_o sum_t W[o, c, 4 - t] g[l + t - 2, o]                     (backward of cnn_utils.py:42-44)
//   2 slots  : the transposed first layer (units -> cin0), 16 k-steps each, N = 16 (8 columns per CTA half).
__global__ void pack_bwd_kernel(const float* __restrict__ params, __nv_bfloat16* __restrict__ img,
                                const PackLayout lay, int n_stacks, int n_layer, int units, int cin0,
                                uint32_t stack_elems) {
  const size_t total = (size_t)n_stacks * stack_elems;
  const uint32_t l0_elems = 2 * L0_B / 2;
  const uint32_t conv_elems = (uint32_t)(n_layer - 1) * SLOTS_CONV * (2 * SLOT_B / 2);
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int st = (int)(idx / stack_elems);
    uint32_t r = (uint32_t)(idx % stack_elems);
    float v = 0.f;
    if (r < l0_elems) {
      const int half = r / (L0_B / 2);
      r %= (L0_B / 2);
      const int ks = r / (2 * NHALF * 8), ch = (r / (NHALF * 8)) & 1, n = (r / 8) % NHALF, e8 = r % 8;
      const int o = half * NHALF + n, t = 2 * ks + ch, f = e8;
      if (o < units && t == 2 && f < lay.fout(st)) v = params[lay.lin_w(st) + (size_t)f * units + o];
    } else if (r < l0_elems + conv_elems) {
      r -= l0_elems;
      const uint32_t slot_elems = 2 * SLOT_B / 2;
      const int sl = r / slot_elems;
      r %= slot_elems;
      const int half = r / (SLOT_B / 2);
      r %= (SLOT_B / 2);
      const int j = n_layer - 1 - sl / SLOTS_CONV;                     // forward layer whose gradient this step propagates
      const int ks = (sl % SLOTS_CONV) * KS_PER_SLOT + r / (2 * NHALF * 8);
      const int ch = (r / (NHALF * 8)) & 1, n = (r / 8) % NHALF, e8 = r % 8;
      const int cdst = half * NHALF + n;                               // output channel of the backward step = input channel of layer j
      int c, t;
      if (cdst < units && conv_k_map(ks, ch * 8 + e8, &c, &t) && c < units)
        v = params[lay.conv_w(st, j) + ((size_t)c * units + cdst) * TAPS + (TAPS - 1 - t)];
    } else {
      r -= l0_elems + conv_elems;
      const uint32_t slot_elems = 2 * FIN_B / 2;
      const int sl = r / slot_elems;
      r %= slot_elems;
      const int half = r / (FIN_B / 2);
      r %= (FIN_B / 2);
      const int ks = sl * KS_FIN_SLOT + r / (2 * LIN_NHALF * 8);
      const int ch = (r / (LIN_NHALF * 8)) & 1, n = (r / 8) % LIN_NHALF, e8 = r % 8;
      const int qdst = half * LIN_NHALF + n;                           // input channel of the first layer
      int c, t;
      if (qdst < cin0 && conv_k_map(ks, ch * 8 + e8, &c, &t) && c < units)
        v = params[lay.conv_w(st, 0) + ((size_t)c * cin0 + qdst) * TAPS + (TAPS - 1 - t)];
    }
    img[idx] = __float2bfloat16_rn(v);
  }
}

// ------------------------------------------------------------------------------------------
// the fused decoder kernel (cluster of 2)
// ------------------------------------------------------------------------------------------
// MODE 0: forward (DEC_LargeCNN / ENC_interCNN schedule; optionally stashes every layer's output for training).
// MODE 1: backward of ONE conv stack + Linear: the same pipeline run on gradients -- "layer 0" is the transposed Linear
//         (F -> units, centre tap only), the units -> units layers use the transposed, tap-flipped weights, every epilogue
//         multiplies by ELU'(y) taken from the stashed forward output (ELU'(z) = y + 1 for y < 0, else 1) and writes the
//         pre-activation gradient g to shared memory (next layer's operand) and to HBM (weight-gradient kernel); the last
//         step is the transposed first layer (units -> 2+F, 5 taps, N = 16).
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(N_THREADS, 1) dec_pair_kernel(const PairArgs a) {
  constexpr bool FWD = (MODE != 1);          // MODE 2 = MODE 0 + stash of every layer's output (training forward)
  constexpr bool STASH = (MODE == 2);
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem S = make_smem(a.F);
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int L = a.L, F = a.F, CW_ROWS = a.L + 2;
  const int n_stacks = a.n_stacks;
  // weight transfers of a stack: layer 0, 8 per units->units layer, then the Linear (MODE 0) or 2 slots of the transposed
  // first layer (MODE 1)
  const int slots_per_stack = (FWD ? 2 : 3) + SLOTS_CONV * (a.n_layer - 1);

  auto bar = [&](int i) { return sbase + S.bars + 8u * (uint32_t)i; };

  // ---- one-time setup ----------------------------------------------------------------------
  for (uint32_t i = threadIdx.x * 16; i < S.bars; i += N_THREADS * 16) st_shared_v4(sbase + i, 0u, 0u, 0u, 0u);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(bar(B_WFULL + i), rank == 0 ? 2 : 1); mbar_init(bar(B_WEMPTY + i), N_TILES); }
    for (int m = 0; m < N_TILES; ++m) {
      mbar_init(bar(B_ACC + m), 1); mbar_init(bar(B_LACC + m), 1);
      mbar_init(bar(B_ACT + m), 2 * N_EPI_WARPS); mbar_init(bar(B_HALO + m), 2 * N_EPI_WARPS); mbar_init(bar(B_LE + m), 2 * N_EPI_WARPS);
      mbar_init(bar(B_ISS + m), 1); mbar_init(bar(B_DEF + m), 2 * PARTS);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < L && FWD; i += N_THREADS) {
    st_shared_u16(sbase + S.perm + 2 * i, (uint16_t)a.perm[i]);
    st_shared_u16(sbase + S.inv_perm + 2 * i, (uint16_t)a.inv_perm[i]);
  }
  if (warp == WARP_PRODUCER) tmem_alloc<2>(sbase + S.tmem_ptr, TMEM_COLS);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + S.tmem_ptr) : "memory");

  const int pair0 = a.pair_begin + (int)cluster_id_x(), pair_stride = (int)n_clusters_x();
  // timeline: the third group of cluster 0 (steady state: weights in L2, clocks settled), else its first
  const int tl_pr = (a.n_pairs > 2 * pair_stride) ? 2 * pair_stride : 0;
  (void)tl_pr;
  unsigned long long t_start_ns = 0, t_start_clk = 0;
  if (TAE_TIMELINE && a.tl && threadIdx.x == 0) {
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start_ns));
    t_start_clk = clock64();
  }

  if (warp == WARP_PRODUCER) {
    // ================= weight producer: this CTA's half of every slot =================================
    if (lane == 0) {
      uint32_t pos = 0, phase = 0;
      for (int pr = pair0; pr < a.n_pairs; pr += pair_stride) {
        for (int st = 0; st < n_stacks; ++st) {
          const uint8_t* src = a.wimg + (size_t)(MODE == 1 ? n_stacks - 1 - st : st) * a.stack_bytes;
          for (int i = 0; i < slots_per_stack; ++i) {
            const uint32_t bytes = (i == 0) ? L0_B : (i <= SLOTS_CONV * (a.n_layer - 1) ? SLOT_B : (FWD ? LIN_B : FIN_B));
            mbar_wait(bar(B_WEMPTY + pos), phase ^ 1, a.err, 1);
            mbar_arrive_expect_tx(bar(B_WFULL + pos), bytes);
            bulk_g2s(sbase + S.wslot + pos * SLOT_B, src + rank * bytes, bytes, bar(B_WFULL + pos));
            src += 2 * bytes;
            if (++pos == NS) { pos = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp >= WARP_MMA && warp < WARP_MMA + N_TILES) {
    if (rank == 1) {
      if (warp == WARP_MMA) {
      // ================= peer relay: tell the leader that this CTA's half of a slot has landed ==========
      uint32_t pos = 0, phase = 0;
      for (int pr = pair0; pr < a.n_pairs; pr += pair_stride)
        for (int i = 0; i < n_stacks * slots_per_stack; ++i) {
          mbar_wait(bar(B_WFULL + pos), phase, a.err, 2);
          if (elect_one()) mbar_arrive_remote(bar(B_WFULL + pos), 0);   // data was written by the async proxy: no fence needed
          __syncwarp();
          if (++pos == NS) { pos = 0; phase ^= 1; }
        }
      }
    } else {
      // ================= MMA issuers: warp 1+m of the leader CTA owns tile m; the warp runs converged and one elected
      // lane issues.  tcgen05.mma issue is nearly synchronous (the queue is about one instruction deep), so a single
      // issuer would expose every barrier wait as tensor-pipe idle time; with one issuer per tile the waits of tile
      // m+1 (inputs written back, weight slots resident) overlap the MMAs of tile m.  Tiles use disjoint accumulators,
      // so the order in which the tensor core interleaves the four streams does not matter.
      const int m = warp - WARP_MMA;
      constexpr uint32_t IDESC_CONV = make_idesc(256, NPAD), IDESC_LIN = make_idesc(256, LIN_N);
      uint32_t pos = 0, wphase = 0, step = 0, csteps = 0, n_act = 0, n_halo = 0, n_le = 0, n_def = 0;
      const uint32_t act = sbase + S.act, comb = sbase + S.comb, ones = sbase + S.ones;
      for (int pr = pair0; pr < a.n_pairs; pr += pair_stride) {
        for (int st = 0; st < n_stacks; ++st) {
          const uint32_t xin = sbase + S.xin[0] + (uint32_t)(MODE == 1 ? 0 : (a.enc ? (st == 2) : (st & 1))) * CHUNK_B;   // enc: branch 3 reads the interleaved bits; backward: one operand chunk
          for (int layer = 0; layer <= a.n_layer; ++layer, ++step) {
            const bool conv = (layer > 0 && layer < a.n_layer);
            {
              const bool stamp = TAE_TIMELINE && a.tl && pr == tl_pr && lane == 0;
              const uint32_t sidx = (uint32_t)(st * (a.n_layer + 1) + layer);
              (void)sidx;
              if (stamp) a.tl[(sidx * 4 + m) * 8 + 0] = clock64();
              if (layer == 0) {
                // the stack input rows this tile reads (its own + 2 halo rows each side) belong to whole codewords, which the
                // per-tile Linear epilogues of the previous stack scatter: wait for the tiles those codewords live in, and
                // always for the neighbours (the same set for every stack: each of these barriers is seen once per stack)
                int t_lo = 0, t_hi = N_TILES - 1;
                if (MODE == 0 && !a.enc) {
                  const int c_lo = max(128 * m - 2, 0) / CW_ROWS;
                  const int c_hi = min(min(128 * m + 129, GROUP_ROWS - 1) / CW_ROWS, a.cw_per_group - 1);
                  t_lo = max(m - 1, 0); t_hi = min(m + 1, N_TILES - 1);
                  if (c_lo <= c_hi) { t_lo = min(t_lo, (c_lo * CW_ROWS) / 128); t_hi = max(t_hi, min(N_TILES - 1, (c_hi * CW_ROWS + L - 1) / 128)); }
                }
                for (int t = t_lo; t <= t_hi; ++t) mbar_wait(bar(B_LE + t), n_le & 1, a.err, 5);
                ++n_le;
              } else if (FWD && layer == a.n_layer) {
                // Linear (centre tap): this tile's own rows, including its two deferred rows
                mbar_wait(bar(B_ACT + m), n_act & 1, a.err, 5);
                ++n_act;
                if (m < N_TILES - 1) mbar_wait(bar(B_DEF + m), n_def & 1, a.err, 10);
                ++n_def;
              } else {
                // tap-shifted step: own rows (and the last rows of tile m-1, stored by this tile's epilogue) + the first two
                // rows of tile m+1 and this tile's deferred rows (stored by the epilogue of tile m+1)
                mbar_wait(bar(B_ACT + m), n_act & 1, a.err, 5);
                ++n_act;
                if (m < N_TILES - 1) mbar_wait(bar(B_HALO + m + 1), n_halo & 1, a.err, 5);
                ++n_halo;
              }
              tc_fence_after();
              if (stamp) a.tl[(sidx * 4 + m) * 8 + 1] = clock64();
              const uint32_t rowoff = (uint32_t)(128 * m) * ROW_B;
              uint32_t p = pos, ph = wphase;
              // descriptors = per-tile base words + compile-time constants (everything below is fully unrolled)
              uint32_t wlo = dlo(sbase + S.wslot + p * SLOT_B, WCHUNK_B);
              const uint32_t d_tmem = tmem_base + m * NPAD;
              if (layer == 0) {
                mbar_wait(bar(B_WFULL + p), ph, a.err, 3);
                tc_fence_after();
                const uint32_t x_lo = dlo(xin + rowoff, ROW_B);
                const uint32_t x2_lo = dlo(xin + rowoff + 4 * ROW_B, (ones + 2 * ROW_B) - (xin + 4 * ROW_B));
                if (elect_one()) {
                  umma_bf16<2>(d_tmem, dfull(x_lo), dfull(wlo), IDESC_CONV, 0);                                 // taps 0,1
                  umma_bf16<2>(d_tmem, dfull(x_lo + 2), dfull(wlo + ((2 * WCHUNK_B) >> 4)), IDESC_CONV, 1);     // taps 2,3
                  umma_bf16<2>(d_tmem, dfull(x2_lo), dfull(wlo + ((4 * WCHUNK_B) >> 4)), IDESC_CONV, 1);        // tap 4, bias
                  umma_commit_pair(bar(B_WEMPTY + p), 3);
                  umma_commit_pair(bar(B_ACC + m), 3);
                }
                __syncwarp();
              } else if (conv) {
                const uint32_t act_lo = dlo(act + rowoff, CHUNK_B);
                const uint32_t c30_lo = dlo(comb + rowoff, 2 * ROW_B);             // [taps 0,1 | taps 2,3] of channels 96..99
                const uint32_t c31_lo = dlo(comb + rowoff + 4 * ROW_B, (ones + 2 * ROW_B) - (comb + 4 * ROW_B));   // [tap 4 | ones (bias)]
                // which of the 8 weight slots are resident already: ONE probe (lane s tests slot s) instead of a wait per slot on
                // the issue path (each satisfied wait still costs ~90 cycles during which the queue of ~1 MMA drains)
                uint32_t rdy;
                {
                  uint32_t pp = p + (uint32_t)(lane & (SLOTS_CONV - 1)), phh = ph;
                  if (pp >= NS) { pp -= NS; phh ^= 1; }
                  rdy = __ballot_sync(0xffffffffu, mbar_test_wait(bar(B_WFULL + pp), phh) != 0) & ((1u << SLOTS_CONV) - 1u);
                }
                // issue order = program order of the units->units tiles, two in flight: wait until the previous tile (tile 3 of
                // the previous units->units step for tile 0) has issued its first half.  One issuing thread alone feeds the
                // tensor pipe at ~67 cycles per MMA (56 nominal: the per-slot descriptor set-up is on its critical path); two
                // interleaved streams reach the nominal rate, and with more than two a tile would complete later than needed
                if (m > 0) mbar_wait(bar(B_ISS + m - 1), csteps & 1, a.err, 11);
                else if (csteps > 0) mbar_wait(bar(B_ISS + N_TILES - 1), (csteps - 1) & 1, a.err, 11);
                ++csteps;
                tc_fence_after();
#pragma unroll
                for (int s = 0; s < SLOTS_CONV; ++s) {
                  if (!((rdy >> s) & 1u)) {
                    mbar_wait(bar(B_WFULL + p), ph, a.err, 3);
                    tc_fence_after();
                  }
                  if (elect_one()) {
#pragma unroll
                    for (int k4 = 0; k4 < KS_PER_SLOT; ++k4) {
                      const int ks = s * KS_PER_SLOT + k4;
                      const uint32_t alo = ks < 30 ? act_lo + ((uint32_t)(2 * (ks % 6)) * CHUNK_B + (uint32_t)(ks / 6) * ROW_B) / 16
                                                   : (ks == 30 ? c30_lo : c31_lo);
                      umma_bf16<2>(d_tmem, dfull(alo), dfull(wlo + (uint32_t)(k4 * 2 * WCHUNK_B) / 16), IDESC_CONV, ks > 0);
                    }
                    umma_commit_pair(bar(B_WEMPTY + p), 3);
                    if (s == SLOTS_CONV / 2 - 1) mbar_arrive_local(bar(B_ISS + m));
                    if (s == SLOTS_CONV - 1) umma_commit_pair(bar(B_ACC + m), 3);
                  }
                  __syncwarp();
                  wlo += SLOT_B / 16;
                  if (++p == NS) { p = 0; ph ^= 1; wlo -= NS * SLOT_B / 16; }
                }
              } else if (MODE == 1) {
                // transposed first layer: all 5 taps x units channels (the conv layers' A sequence) onto N = 16 columns
                const uint32_t act_lo = dlo(act + rowoff, CHUNK_B);
                const uint32_t c30_lo = dlo(comb + rowoff, 2 * ROW_B);
                const uint32_t c31_lo = dlo(comb + rowoff + 4 * ROW_B, (ones + 2 * ROW_B) - (comb + 4 * ROW_B));
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                  mbar_wait(bar(B_WFULL + p), ph, a.err, 3);
                  tc_fence_after();
                  const uint32_t wl = dlo(sbase + S.wslot + p * SLOT_B, LIN_WCHUNK_B);
                  if (elect_one()) {
#pragma unroll
                    for (int k16 = 0; k16 < KS_FIN_SLOT; ++k16) {
                      const int ks = s * KS_FIN_SLOT + k16;
                      const uint32_t alo = ks < 30 ? act_lo + ((uint32_t)(2 * (ks % 6)) * CHUNK_B + (uint32_t)(ks / 6) * ROW_B) / 16
                                                   : (ks == 30 ? c30_lo : c31_lo);
                      umma_bf16<2>(tmem_base + TMEM_LIN_COL + m * LIN_N, dfull(alo), dfull(wl + (uint32_t)(k16 * 2 * LIN_WCHUNK_B) / 16),
                                   IDESC_LIN, ks > 0);
                    }
                    umma_commit_pair(bar(B_WEMPTY + p), 3);
                    if (s == 1) umma_commit_pair(bar(B_ACC + m), 3);
                  }
                  __syncwarp();
                  if (++p == NS) { p = 0; ph ^= 1; }
                }
              } else {
                mbar_wait(bar(B_WFULL + p), ph, a.err, 3);
                tc_fence_after();
                const uint32_t act_lo = dlo(act + rowoff + 2 * ROW_B, CHUNK_B);                      // centre row only
                const uint32_t c_lo = dlo(comb + rowoff + 2 * ROW_B, ones - comb);
                const uint32_t wl = dlo(sbase + S.wslot + p * SLOT_B, LIN_WCHUNK_B);
                if (elect_one()) {
#pragma unroll
                  for (int ks = 0; ks < KS_LIN; ++ks)
                    umma_bf16<2>(tmem_base + TMEM_LIN_COL + m * LIN_N, dfull(ks < 6 ? act_lo + (uint32_t)(2 * ks) * CHUNK_B / 16 : c_lo),
                                 dfull(wl + (uint32_t)(ks * 2 * LIN_WCHUNK_B) / 16), IDESC_LIN, ks > 0);
                  umma_commit_pair(bar(B_WEMPTY + p), 3);
                  umma_commit_pair(bar((a.enc ? B_ACC : B_LACC) + m), 3);      // decoder: its own barrier (only column part m waits for it)
                }
                __syncwarp();
              }
              if (stamp) a.tl[(sidx * 4 + m) * 8 + 2] = clock64();
            }
            // advance the ring past this layer's slots
            const int n_adv = conv ? SLOTS_CONV : ((MODE == 1 && layer == a.n_layer) ? 2 : 1);
            for (int s = 0; s < n_adv; ++s)
              if (++pos == NS) { pos = 0; wphase ^= 1; }
          }
        }
      }
    }
  } else {
    // ================= epilogue warps: TMEM -> ELU -> bf16 -> shared memory ======================
    const int ew = warp - EPI_WARP0;
    const int q = warp & 3;                    // TMEM lane quadrant this warp may read
    const int part = ew >> 2;                  // column part: channels [8*CPP*part, 8*CPP*(part+1)); the last part adds 96..99
    const int tid = threadIdx.x - EPI_WARP0 * 32;
    uint32_t step = 0, n_acc = 0, n_lacc = 0;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * q) << 16);
    // decoder input of the NEXT group: pulled into L2 a whole stack before it is staged (no registers held)
    auto prefetch_received = [&](int pr_) {
      if (MODE == 1 || a.enc || pr_ >= a.n_pairs) return;
      const int cw0_ = (2 * pr_ + (int)rank) * a.cw_per_group;
      const int n_ = max(0, min(a.cw_per_group, a.B - cw0_));
      if (tid * 32 < n_ * L * 3)                       // one 128-byte line per thread
        asm volatile("prefetch.global.L2 [%0];" ::"l"(a.received + (size_t)cw0_ * L * 3 + (size_t)tid * 32) : "memory");
    };
    for (int pr = pair0; pr < a.n_pairs; pr += pair_stride) {
      const int grp = 2 * pr + (int)rank;
      const int cw0 = grp * a.cw_per_group;
      const int n_cw = max(0, min(a.cw_per_group, a.B - cw0));
      uint32_t vmask = 0;      // bit m: this thread's row of tile m holds a real codeword position
#pragma unroll
      for (int m = 0; m < N_TILES; ++m) {
        const int g = 128 * m + 32 * q + lane;
        vmask |= ((g % CW_ROWS < L) && (g / CW_ROWS < n_cw)) ? (1u << m) : 0u;
      }

      if (TAE_TIMELINE && a.tl && pair0 == 0 && rank == 0 && tid == 0) a.tl[72 * 4 * 8 + 148 * 4 + (pr / pair_stride)] = clock64();
      // ---- group start: stack inputs (prior channels = 0, decoders.py:227), ones chunk ----------------------------------------
#pragma unroll 1
      for (uint32_t i = tid * 16; i < 3 * CHUNK_B; i += N_EPI_THREADS * 16) st_shared_v4(sbase + S.xin[0] + i, 0u, 0u, 0u, 0u);
      epi_bar_sync();
      const bool grp_ok = grp < a.n_groups;               // the odd group of the last pair does not exist: no stash traffic
      // a group has at most GROUP_ROWS = N_EPI_THREADS positions: one (codeword, position) per epilogue thread
      const int sc = tid / L, sl = tid - sc * L;
      const bool s_ok = tid < n_cw * L;
      const uint32_t s_row = (uint32_t)(sc * CW_ROWS + sl + 2) * ROW_B;
      if (MODE == 1) {
        // the stack loop below fills the operand chunk of each stack's first step
      } else if (a.enc) {
        if (s_ok) {
          const uint16_t x = bf16_bits(2.0f * a.u[(size_t)cw0 * L + tid] - 1.0f);                      // encoders.py:362
          const uint32_t row_i = (uint32_t)(sc * CW_ROWS + ld_shared_u16(sbase + S.inv_perm + 2 * sl) + 2) * ROW_B;
          st_shared_u16(sbase + S.xin[0] + s_row, x);          // branches 1, 2
          st_shared_u16(sbase + S.xin[1] + row_i, x);          // branch 3: x_int[i] = x[p[i]]          (encoders.py:369)
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + S.ones + s_row), "r"(0x3F803F80u) : "memory");
        }
      } else if (s_ok) {
        // (`received` of this group was pulled into L2 during the last stack of the previous group)
        const float* rsrc = a.received + ((size_t)cw0 * L + tid) * 3;
        const float pre0 = __ldg(rsrc), pre1 = __ldg(rsrc + 1), pre2 = __ldg(rsrc + 2);
        const uint32_t row_i = (uint32_t)(sc * CW_ROWS + ld_shared_u16(sbase + S.inv_perm + 2 * sl) + 2) * ROW_B;
        st_shared_u16(sbase + S.xin[0] + s_row + 0, bf16_bits(pre0));      // r_sys          (decoders.py:221)
        st_shared_u16(sbase + S.xin[0] + s_row + 2, bf16_bits(pre1));      // r_par1         (decoders.py:223)
        st_shared_u16(sbase + S.xin[1] + row_i + 0, bf16_bits(pre0));      // r_sys_int[i] = r_sys[p[i]]  (:222)
        st_shared_u16(sbase + S.xin[1] + s_row + 2, bf16_bits(pre2));      // r_par2         (decoders.py:224)
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + S.ones + s_row), "r"(0x3F803F80u) : "memory");   // bias inputs
      }
      if (FWD) {
        fence_proxy_async();
        epi_bar_sync();
        if (lane == 0)
          for (int m = 0; m < N_TILES; ++m) mbar_arrive_leader(bar(B_LE + m), rank);       // inputs of the first stack
      }

      // Deferred rows: the last two rows of tile m are still read by the MMAs of tile m+1 (taps 0, 1), so the two
      // lanes that own them park their packed outputs in shared memory (S.defer) and store them at the start of the next tile.
      bool le_pending = false;       // this warp's Linear epilogue (tile == part) of stack le_stack (step le_step) is due
      int le_stack = 0;
      uint32_t le_step = 0;

      // -- Linear epilogue of ONE tile of decoder stack `ls` (run by the 4 warps of column part == tile: one row per lane):
      //    extrinsic subtraction + (de)interleave into the next stack's input (decoders.py:235-249), or the final sigmoid.
      //    The prior that is subtracted is the bf16 value the stack actually saw (channels 2.. of its own input row), so no
      //    separate fp32 prior buffer exists, and the extrinsic values of a row go out as one 4-byte and one 8-byte store.
      //    `lstep` = step index of that stack's Linear.
      auto lin_tile = [&](int ls, int t, uint32_t lstep) {
        const bool last = (ls == n_stacks - 1);
        const uint32_t xin_cur = sbase + S.xin[0] + (uint32_t)(ls & 1) * CHUNK_B, xin_nxt = sbase + S.xin[0] + (uint32_t)((ls & 1) ^ 1) * CHUNK_B;
        const uint32_t map = sbase + (last ? S.perm : ((ls & 1) ? S.perm : S.inv_perm));   // where position l lands
        const bool stamp = TAE_TIMELINE && a.tl && pr == tl_pr && rank == 0 && lane == 0 && q == 0;
        const uint32_t sidx_l = (uint32_t)(ls * (a.n_layer + 1) + a.n_layer);
        (void)sidx_l;
        if (stamp) a.tl[(sidx_l * 4 + t) * 8 + 3] = clock64();
        (void)lstep;
        mbar_wait(bar(B_LACC + t), n_lacc & 1, a.err, 6);
        ++n_lacc;
        tc_fence_after();
        if (stamp) a.tl[(sidx_l * 4 + t) * 8 + 4] = clock64();
        const int g_row = 128 * t + 32 * q + lane;
        const int g_cw = g_row / CW_ROWS, g_l = g_row - g_cw * CW_ROWS;
        const bool valid = (vmask >> t) & 1u;
        uint32_t r[8];
        tmem_ld8(lane_addr + TMEM_LIN_COL + (uint32_t)(t * LIN_N), r);
        uint32_t x0 = 0, x1 = 0, x2 = 0, x3 = 0, dl = 0;
        if (valid) {
          dl = ld_shared_u16(map + 2 * g_l);
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3)
                       : "r"(xin_cur + (uint32_t)(g_row + 2) * ROW_B) : "memory");
        }
        tmem_ld_wait();
        if (valid) {
          const int cw = cw0 + g_cw, l = g_l;
          if (a.trace) {
            float* tr = a.trace + (((size_t)ls * a.B + cw) * L + l) * F;
            const int nf = last ? 1 : F;
#pragma unroll
            for (int f = 0; f < 5; ++f)
              if (f < nf) tr[f] = __uint_as_float(r[f]);
          }
          if (last) {
            // deinterleave: out[p[l]] = sigmoid(x[l])                                       (decoders.py:267)
            a.out[(size_t)cw * L + dl] = 1.f / (1.f + __expf(-__uint_as_float(r[0])));
          } else {
            // input channels of this row: [sys, par, prior_0..4, 0] as bf16 pairs (x0 = ch0,1  x1 = ch2,3 ...)
            const bool ex = a.extrinsic != 0;
            float e[6];
            e[0] = __uint_as_float(r[0]) - (ex ? __uint_as_float(x1 << 16) : 0.f);            // decoders.py:235-236, 246-247
            e[1] = __uint_as_float(r[1]) - (ex ? __uint_as_float(x1 & 0xFFFF0000u) : 0.f);
            e[2] = __uint_as_float(r[2]) - (ex ? __uint_as_float(x2 << 16) : 0.f);
            e[3] = __uint_as_float(r[3]) - (ex ? __uint_as_float(x2 & 0xFFFF0000u) : 0.f);
            e[4] = __uint_as_float(r[4]) - (ex ? __uint_as_float(x3 << 16) : 0.f);
            e[5] = 0.f;
#pragma unroll
            for (int f = 0; f < 5; ++f)
              if (f >= F) e[f] = 0.f;
            const uint32_t drow = (uint32_t)(g_cw * CW_ROWS) + dl + 2;
            const uint32_t dst = xin_nxt + drow * ROW_B;
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + 4), "r"(pack_bf16x2(e[0], e[1])) : "memory");
            st_shared_v2(dst + 8, pack_bf16x2(e[2], e[3]), pack_bf16x2(e[4], e[5]));
          }
        }
        (void)x0;
        if (!last) {
          // 4 warps of each CTA report a tile: 4 arrivals each make up the 2 * N_EPI_WARPS the barrier expects
          fence_proxy_async();
          tc_fence_before();
          __syncwarp();
          if (lane < N_EPI_WARPS / 4) mbar_arrive_leader(bar(B_LE + t), rank);
        }
        if (stamp) a.tl[(sidx_l * 4 + t) * 8 + 5] = clock64();
      };

      for (int st = 0; st < n_stacks; ++st) {
        const int sr = MODE == 1 ? n_stacks - 1 - st : st;          // MODE 1 walks the schedule backwards
        float* dxin_s = nullptr;
        if (MODE == 1) {
          // ---- gradient w.r.t. this stack's Linear output -> the 8-channel operand chunk of its first (transposed Linear) step ------
          const int fin = a.lay.fout(sr);
          const size_t BL = (size_t)a.B * L;
          dxin_s = a.dxin + (size_t)sr * BL * 8;
          const bool direct = a.chain == 0 || sr == n_stacks - 1;
          const float* src_d = a.dlin + (a.chain == 0 ? (size_t)sr * BL * fin : 0);
          const float* px_b = a.dxin + (size_t)(sr + 1) * BL * 8 + 2;
          const float* pd_b = a.dlin_all + (size_t)(sr + 1) * BL * F;
          const int32_t* idx = ((sr + 1) & 1) ? a.inv_perm : a.perm;
          const bool sub = a.extrinsic && (sr + 1 != n_stacks - 1);
          float* out_d = (a.chain && a.dlin_all) ? a.dlin_all + (size_t)sr * BL * F : nullptr;
          if (st > 0) epi_bar_sync();          // the previous stack's dxin rows (written by this CTA) are visible to every thread
          float bsum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
          for (int i = tid; i < n_cw * L; i += N_EPI_THREADS) {
            const int c = i / L, l = i % L;
            float v[8];
            if (direct) {
              const float* d = src_d + ((size_t)(cw0 + c) * L + l) * fin;
#pragma unroll
              for (int f = 0; f < 8; ++f) v[f] = f < fin ? d[f] : 0.f;
            } else {
              const size_t src = (size_t)(cw0 + c) * L + idx[l];
              const float* px = px_b + src * 8;
              const float* pd = pd_b + src * F;
#pragma unroll
              for (int f = 0; f < 8; ++f) v[f] = f < fin ? px[f] - (sub ? pd[f] : 0.f) : 0.f;
            }
            if (out_d) {
              float* o = out_d + ((size_t)(cw0 + c) * L + l) * F;
#pragma unroll
              for (int f = 0; f < 5; ++f)
                if (f < fin) o[f] = v[f];
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) bsum[f] += v[f];
            st_shared_v4(sbase + S.xin[0] + (uint32_t)(c * CW_ROWS + l + 2) * ROW_B, pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                         pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
          }
          if (a.grad_flat) {
            float* bg = a.grad_flat + a.lay.lin_b(sr);
#pragma unroll
            for (int f = 0; f < 5; ++f) {
              float t = bsum[f];
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
              if (lane == 0 && f < fin && t != 0.f) atomicAdd(bg + f, t);
            }
          }
          fence_proxy_async();
          epi_bar_sync();
          if (lane == 0)
            for (int m = 0; m < N_TILES; ++m) mbar_arrive_leader(bar(B_LE + m), rank);
          if (a.stash_x && grp_ok)          // the dlin image of this stack (B operand of the Linear's weight gradient)
            for (int i = tid; i < BUF_ROWS; i += N_EPI_THREADS) {
              uint32_t x0, x1, x2, x3;
              asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(sbase + S.xin[0] + (uint32_t)i * ROW_B) : "memory");
              *reinterpret_cast<uint4*>(a.stash_x + (((size_t)sr * a.n_groups + grp) * BUF_ROWS + i) * ROW_B) = make_uint4(x0, x1, x2, x3);
            }
        }
        for (int layer = 0; layer <= a.n_layer; ++layer, ++step) {
          // B_ACC phases this warp has seen: every step except the decoder's Linear (B_LACC, column part = tile only)
          const bool acc_step = !(FWD && !a.enc && layer == a.n_layer);
          const uint32_t par = n_acc & 1;
          if (acc_step) ++n_acc;
          if (FWD && st == n_stacks - 1 && layer == 1) prefetch_received(pr + pair_stride);
          const uint32_t sidx = (uint32_t)(st * (a.n_layer + 1) + layer);      // timeline row
          (void)sidx;
          const bool last_step = (st == n_stacks - 1 && layer == a.n_layer);
          if (MODE == 1 && layer == a.n_layer) {
            // -- gradient w.r.t. the stack input (transposed first layer): one (8-float) row per position -----------------
#pragma unroll
            for (int m = 0; m < N_TILES; ++m) mbar_wait(bar(B_ACC + m), par, a.err, 6);
            tc_fence_after();
            if (part == 0) {
#pragma unroll 1
              for (int m = 0; m < N_TILES; ++m) {
                uint32_t r[8];
                tmem_ld8(lane_addr + TMEM_LIN_COL + (uint32_t)(m * LIN_N), r);
                tmem_ld_wait();
                if (!((vmask >> m) & 1u)) continue;
                const int g_row = 128 * m + 32 * q + lane;
                const int g_cw = g_row / CW_ROWS, g_l = g_row - g_cw * CW_ROWS;
                float4* dst = reinterpret_cast<float4*>(dxin_s + ((size_t)(cw0 + g_cw) * L + g_l) * 8);
                dst[0] = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]), __uint_as_float(r[2]), __uint_as_float(r[3]));
                dst[1] = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]), __uint_as_float(r[6]), __uint_as_float(r[7]));
              }
            }
            tc_fence_before();              // (the next stack's prologue, or the next group's, arrives on B_ACT)
            __syncwarp();
            continue;
          }
          if (FWD && layer == a.n_layer && a.enc) {
            // -- ENC_interCNN tail: x_tx[:, :, branch] = ELU(Linear(h)) (encoders.py:364,367,371) + the power sums ------
#pragma unroll
            for (int m = 0; m < N_TILES; ++m) mbar_wait(bar(B_ACC + m), par, a.err, 6);
            tc_fence_after();
            double s1 = 0.0, s2 = 0.0;
            if (part == 0) {
#pragma unroll 1
              for (int m = 0; m < N_TILES; ++m) {
                const uint32_t r0 = tmem_ld1(lane_addr + TMEM_LIN_COL + (uint32_t)(m * LIN_N));
                tmem_ld_wait();
                if (!((vmask >> m) & 1u)) continue;
                const int g_row = 128 * m + 32 * q + lane;
                const int g_cw = g_row / CW_ROWS, g_l = g_row - g_cw * CW_ROWS;
                const float z = __uint_as_float(r0);
                const float v = z > 0.f ? z : expm1f(z);
                a.x_tx[((size_t)(cw0 + g_cw) * L + g_l) * 3 + st] = v;
                s1 += (double)v;
                s2 += (double)v * (double)v;
              }
#pragma unroll
              for (int d = 16; d > 0; d >>= 1) {
                s1 += __shfl_xor_sync(0xffffffffu, s1, d);
                s2 += __shfl_xor_sync(0xffffffffu, s2, d);
              }
              if (lane == 0) { atomicAdd(a.stats + 0, s1); atomicAdd(a.stats + 1, s2); }
            }
            if (!last_step) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0)
                for (int m = 0; m < N_TILES; ++m) mbar_arrive_leader(bar(B_LE + m), rank);     // the next branch's input is in place since the group start
            }
            continue;
          }
          if (layer == a.n_layer) continue;     // decoder Linear: its per-tile epilogues are interleaved with the units->units ones (lin_tile)
          // training: this layer's group image in HBM (MODE 0: stash of the outputs; MODE 1: forward outputs in, gradients out)
          uint8_t* img_out = nullptr;
          const uint8_t* img_y = nullptr;
          if (FWD) {
            if (STASH && a.stash_y && grp_ok) img_out = a.stash_y + ((size_t)(st * a.n_layer + layer) * a.n_groups + grp) * (IMG_CHUNKS * CHUNK_B);
          } else if (grp_ok) {
            const size_t off = ((size_t)(sr * a.n_layer + a.n_layer - 1 - layer) * a.n_groups + grp) * (IMG_CHUNKS * CHUNK_B);
            img_y = a.stash_y + off;
            img_out = a.stash_g + off;
          }
          const int m_end = N_TILES + ((FWD && !a.enc && st == n_stacks - 1 && layer == a.n_layer - 1) ? 1 : 0);   // + a pass that only drains the Linear epilogues
#pragma unroll 1
          for (int m = 0; m < m_end; ++m) {
            if (FWD && !a.enc && le_pending) {
              // Linear epilogue of the decoder: tile t belongs to the 4 warps of column part t and is due once the last
              // units->units layer has written tile t completely (the epilogue of tile t+1 has begun: B_DEF), i.e. right
              // here, before this warp's next tile.  Tile 3 alone is postponed past tile 0 of the next stack's first layer when
              // that tile's input codewords do not reach into tile 3 (same rule as the issuer's; training stashes the whole
              // stack input there, so it waits for all tiles): the first layer of the next stack then starts on tiles 0, 1
              // while tile 3 of this one is still in its epilogue.
              // (Measured alternatives, profiles/r02_dec_schedule_variants.md: running it only once its accumulators have
              // arrived -- a non-blocking probe, or one tile later -- avoids the 0.7 - 2 k cycle wait here but starts the next
              // stack's first layer later and is slower overall.)
              bool now = true;
              if (!STASH && part == N_TILES - 1 && layer == 0 && m == 0) {
                const int c_hi = min(129 / CW_ROWS, a.cw_per_group - 1);
                now = min(N_TILES - 1, (c_hi * CW_ROWS + L - 1) / 128) >= N_TILES - 1;
              }
              if (now) { lin_tile(le_stack, part, le_step); le_pending = false; }
            }
            if (m == N_TILES) break;
            const bool stamp = TAE_TIMELINE && a.tl && pr == tl_pr && rank == 0 && lane == 0 && ew == 0;
            if (stamp && ew == 0) a.tl[(sidx * 4 + m) * 8 + 3] = clock64();
            const int g_row = 128 * m + 32 * q + lane;
            const uint32_t brow = (uint32_t)(g_row + 2);
            const bool has_comb = (part == PARTS - 1);               // this part also owns channels 96..99
            uint4 yv[CPP];
            uint2 yc = make_uint2(0u, 0u);
            if (MODE == 1) {
              // forward outputs of this row (ELU' comes from them), requested before the accumulators are waited for
#pragma unroll
              for (int c = 0; c < CPP; ++c)
                yv[c] = img_y ? __ldg(reinterpret_cast<const uint4*>(img_y + (size_t)(part * CPP + c) * CHUNK_B + brow * ROW_B)) : make_uint4(0u, 0u, 0u, 0u);
              if (has_comb && img_y) yc = __ldg(reinterpret_cast<const uint2*>(img_y + (size_t)N_REG_CHUNKS * CHUNK_B + brow * ROW_B));
            }
            mbar_wait(bar(B_ACC + m), par, a.err, 6);
            tc_fence_after();
            if (stamp) a.tl[(sidx * 4 + m) * 8 + 4] = clock64();
            // per-warp stamps of one stack (both CTAs): [layer][tile][cta][warp][acc seen, reported]
            const bool wstamp = TAE_TIMELINE && a.tl && pr == tl_pr && lane == 0 && st == 6 && layer < 5;
            unsigned long long* wtl = a.tl + 72 * 4 * 8 + 148 * 4 + 128 + ((((size_t)layer * 4 + m) * 2 + rank) * 16 + ew) * 2;
            (void)wtl;
            if (wstamp) wtl[0] = clock64();
            if (STASH && layer == 0 && m == 0 && a.stash_x && grp_ok) {
              // the stack input (sys, parity, priors as the stack saw them): complete since the MMAs of layer 0 were released
              const uint32_t xsrc = sbase + S.xin[0] + (uint32_t)(a.enc ? (st == 2) : (st & 1)) * CHUNK_B;
              for (int i = tid; i < BUF_ROWS; i += N_EPI_THREADS) {
                uint32_t x0, x1, x2, x3;
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(xsrc + (uint32_t)i * ROW_B) : "memory");
                *reinterpret_cast<uint4*>(a.stash_x + (((size_t)st * a.n_groups + grp) * BUF_ROWS + i) * ROW_B) = make_uint4(x0, x1, x2, x3);
              }
            }
            const bool valid = (vmask >> m) & 1u;
            {
              const bool owns_tail = (q == 3) && (lane >= 30);
              const bool defer = owns_tail && (m < N_TILES - 1);
              const uint32_t col0 = (uint32_t)(m * NPAD + part * CPP * 8);
              const uint32_t act_part = sbase + S.act + (uint32_t)(part * CPP) * CHUNK_B;
              uint32_t r[CPP * 8 + 8];
#pragma unroll
              for (int c = 0; c + 1 < CPP; c += 2) tmem_ld16(lane_addr + col0 + 8 * c, r + 8 * c);
              if (CPP & 1) tmem_ld8(lane_addr + col0 + 8 * (CPP - 1), r + 8 * (CPP - 1));
              if (has_comb) tmem_ld8(lane_addr + (uint32_t)(m * NPAD + N_REG_CH), r + 8 * CPP);
              // -- store the rows deferred from tile m-1 (their readers, the MMAs of tile m, have completed) -----
              if (owns_tail && m > 0) {
                // (parked in shared memory by this same thread during tile m-1: program order is all the ordering it needs)
                const uint32_t prow = brow - 128;
                const uint32_t park = sbase + S.defer + (uint32_t)(lane - 30) * (N_REG_CHUNKS + 1) * ROW_B + (uint32_t)(part * CPP) * ROW_B;
#pragma unroll
                for (int c = 0; c < CPP; ++c) {
                  uint32_t d0, d1, d2, d3;
                  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(d0), "=r"(d1), "=r"(d2), "=r"(d3) : "r"(park + (uint32_t)c * ROW_B) : "memory");
                  st_shared_v4(act_part + (uint32_t)c * CHUNK_B + prow * ROW_B, d0, d1, d2, d3);
                }
                if (has_comb) {
                  uint32_t d0, d1;
                  asm volatile("ld.shared.v2.b32 {%0,%1}, [%2];" : "=r"(d0), "=r"(d1) : "r"(park + (uint32_t)CPP * ROW_B) : "memory");
                  st_shared_v2(sbase + S.comb + prow * ROW_B, d0, d1);
                  st_shared_v2(sbase + S.comb + (prow - 1) * ROW_B + 8, d0, d1);
                }
              }
              if (q == 3 && m > 0 && FWD && layer == a.n_layer - 1) {
                // tile m-1 is complete now (B_DEF: what its Linear waits for instead of this tile's whole epilogue)
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(bar(B_DEF + m - 1), rank);
              }
              tmem_ld_wait();
              if (stamp && ew == 0) a.tl[(sidx * 4 + m) * 8 + 6] = clock64();
              const uint32_t keep = valid ? 0xFFFFFFFFu : 0u;
#pragma unroll
              for (int c = 0; c < CPP; ++c) {
                uint32_t p[4];
                if (FWD) {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    p[j] = pack_bf16x2(elu_act<MODE == 0>(__uint_as_float(r[8 * c + 2 * j])), elu_act<MODE == 0>(__uint_as_float(r[8 * c + 2 * j + 1]))) & keep;
                } else {
                  const uint32_t yw[4] = {yv[c].x, yv[c].y, yv[c].z, yv[c].w};
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    p[j] = pack_bf16x2(__uint_as_float(r[8 * c + 2 * j]) * elu_grad_lo(yw[j]), __uint_as_float(r[8 * c + 2 * j + 1]) * elu_grad_hi(yw[j])) & keep;
                }
                if (img_out) *reinterpret_cast<uint4*>(img_out + (size_t)(part * CPP + c) * CHUNK_B + brow * ROW_B) = make_uint4(p[0], p[1], p[2], p[3]);
                if (defer) {
                  st_shared_v4(sbase + S.defer + (uint32_t)(lane - 30) * (N_REG_CHUNKS + 1) * ROW_B + (uint32_t)(part * CPP + c) * ROW_B, p[0], p[1], p[2], p[3]);
                } else {
                  st_shared_v4(act_part + (uint32_t)c * CHUNK_B + brow * ROW_B, p[0], p[1], p[2], p[3]);
                }
              }
              if (has_comb) {
                uint32_t p0, p1;
                if (FWD) {
                  p0 = pack_bf16x2(elu_act<MODE == 0>(__uint_as_float(r[8 * CPP])), elu_act<MODE == 0>(__uint_as_float(r[8 * CPP + 1]))) & keep;
                  p1 = pack_bf16x2(elu_act<MODE == 0>(__uint_as_float(r[8 * CPP + 2])), elu_act<MODE == 0>(__uint_as_float(r[8 * CPP + 3]))) & keep;
                } else {
                  p0 = pack_bf16x2(__uint_as_float(r[8 * CPP]) * elu_grad_lo(yc.x), __uint_as_float(r[8 * CPP + 1]) * elu_grad_hi(yc.x)) & keep;
                  p1 = pack_bf16x2(__uint_as_float(r[8 * CPP + 2]) * elu_grad_lo(yc.y), __uint_as_float(r[8 * CPP + 3]) * elu_grad_hi(yc.y)) & keep;
                }
                if (img_out) *reinterpret_cast<uint4*>(img_out + (size_t)N_REG_CHUNKS * CHUNK_B + brow * ROW_B) = make_uint4(p0, p1, 0u, 0u);
                if (defer) {
                  st_shared_v2(sbase + S.defer + (uint32_t)(lane - 30) * (N_REG_CHUNKS + 1) * ROW_B + (uint32_t)N_REG_CHUNKS * ROW_B, p0, p1);
                } else {
                  st_shared_v2(sbase + S.comb + brow * ROW_B, p0, p1);              // x[r][96..99]  -> comb[r][0:4]
                  st_shared_v2(sbase + S.comb + (brow - 1) * ROW_B + 8, p0, p1);    //               -> comb[r-1][4:8]
                }
              }
            }
            {
              // every warp reports on its own: no CTA-wide barrier on the tile-to-tile critical path
              if (stamp && ew == 0) a.tl[(sidx * 4 + m) * 8 + 7] = clock64();
              fence_proxy_async();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) {
                mbar_arrive_leader(bar(B_ACT + m), rank);
                // the issuer of tile m-1 reads this tile's first rows in the next step if that step is tap-shifted: every
                // conv-type layer but the last of a forward stack (whose reader is the centre-tap Linear)
                if (m > 0 && (!FWD || layer < a.n_layer - 1)) mbar_arrive_leader(bar(B_HALO + m), rank);
              }
            }
            if (stamp) a.tl[(sidx * 4 + m) * 8 + 5] = clock64();
            if (wstamp) wtl[1] = clock64();
            if (FWD && !a.enc && layer == a.n_layer - 1 && (m == part + 1 || (m == N_TILES - 1 && part == N_TILES - 1))) {
              le_pending = true; le_stack = st; le_step = step + 1;
            }
          }
        }
      }
      tc_fence_before();
      epi_bar_sync();        // every epilogue thread is done with this group before its buffers are reloaded
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (TAE_TIMELINE && a.tl && threadIdx.x == 0) {
    unsigned long long t_end_ns;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end_ns));
    unsigned long long* o = a.tl + 72 * 4 * 8 + (size_t)blockIdx.x * 4;
    o[0] = t_start_ns; o[1] = t_end_ns; o[2] = clock64() - t_start_clk;
    unsigned int smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    o[3] = smid;
  }
  cluster_sync_all();
  if (warp == WARP_PRODUCER) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<2>(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// UMMA probes (self-tests of the operand schemes the decoder relies on)
// ------------------------------------------------------------------------------------------
// (1) one CTA: D[128 x N] = A_eff[128 x K] * B[N x K]^T where A is ONE 8-column chunk X[R][8] and K chunk j of
//     A_eff is X shifted down by j*lbo_rows rows (LBO = 16*lbo_rows bytes): the "two taps in one k-step" scheme.
__global__ void __launch_bounds__(128, 1)
probe_lbo_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D, int R,
                 int N, int shift, int lbo_rows, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t b_off = (uint32_t)R * 16, lbo_b = (uint32_t)N * 16;
  const uint32_t bar_off = (b_off + 2 * lbo_b + 15) / 16 * 16, tptr_off = bar_off + 16;
  for (int i = threadIdx.x; i < R * 8; i += blockDim.x) reinterpret_cast<__nv_bfloat16*>(smem)[i] = X[i];
  for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
    const int n = i / 16, c = i % 16;
    reinterpret_cast<__nv_bfloat16*>(smem + b_off)[(size_t)(c / 8) * N * 8 + n * 8 + (c % 8)] = Bm[i];
  }
  if (threadIdx.x == 0) { mbar_init(sbase + bar_off, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<1>(sbase + tptr_off, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + tptr_off) : "memory");
  if (threadIdx.x == 0) {
    umma_bf16<1>(tmem_base, make_desc(sbase + shift * 16, (uint32_t)lbo_rows * 16), make_desc(sbase + b_off, lbo_b),
                 make_idesc(128, N), 0);
    umma_commit_1(sbase + bar_off);
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
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<1>(tmem_base, 128); }
}

// (2) CTA pair: D[256 x N] = A[256 x K] * B[N x K]^T with cta_group::2; CTA r stages rows 128r.. of A and
//     columns r*N/2.. of B (the decoder's assumption), reads back its own 128 rows of D.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
probe_pair_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ Bm, float* __restrict__ D, int K,
                  int N, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t rank = cluster_ctarank();
  const int NH = N / 2;
  const uint32_t lbo_a = 128 * 16, lbo_b = (uint32_t)NH * 16;
  const uint32_t b_off = (uint32_t)(K / 8) * lbo_a;
  const uint32_t bar_off = (b_off + (uint32_t)(K / 8) * lbo_b + 15) / 16 * 16, tptr_off = bar_off + 16;
  for (int i = threadIdx.x; i < 128 * K; i += blockDim.x) {
    const int r = i / K, c = i % K;
    reinterpret_cast<__nv_bfloat16*>(smem)[(size_t)(c / 8) * 128 * 8 + r * 8 + (c % 8)] = A[(size_t)(128 * rank + r) * K + c];
  }
  for (int i = threadIdx.x; i < NH * K; i += blockDim.x) {
    const int n = i / K, c = i % K;
    reinterpret_cast<__nv_bfloat16*>(smem + b_off)[(size_t)(c / 8) * NH * 8 + n * 8 + (c % 8)] = Bm[(size_t)(NH * rank + n) * K + c];
  }
  if (threadIdx.x == 0) { mbar_init(sbase + bar_off, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<2>(sbase + tptr_off, 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + tptr_off) : "memory");
  if (threadIdx.x == 0 && rank == 0) {
    for (int ks = 0; ks < K / 16; ++ks)
      umma_bf16<2>(tmem_base, make_desc(sbase + 2 * ks * lbo_a, lbo_a), make_desc(sbase + b_off + 2 * ks * lbo_b, lbo_b),
                   make_idesc(256, N), ks > 0);
    umma_commit_pair(sbase + bar_off, 3);
  }
  mbar_wait(sbase + bar_off, 0, err, 8);
  tc_fence_after();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int cb = 0; cb < N / 16; ++cb) {
    uint32_t r[16];
    tmem_ld16(tmem_base + ((uint32_t)(32 * warp) << 16) + cb * 16, r);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) D[(size_t)(128 * rank + 32 * warp + lane) * N + cb * 16 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<2>(tmem_base, 128); }
}


// (3) MMA issue-rate probe: one CTA (CG = 1) or CTA pair (CG = 2) per launch streams `n_mma` accumulating MMAs whose
//     operands rotate through shared memory (A: 512 rows x 96 ch layout of the decoder, B: 8 k-step slots), with no
//     epilogue; reports clock64 cycles from first issue to completion.  Measures the operand-fetch-limited MMA rate.
template <int CG>
__device__ __forceinline__ void probe_rate_body(int N, int n_mma, int same_operands, long long* cycles, const uint8_t* src, int tma_on, int sts_warps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
  const uint32_t a_off = 0, a_bytes = 12 * CHUNK_B;                 // activation-like region
  const uint32_t nh = (uint32_t)N / CG;
  const uint32_t b_off = a_bytes, b_bytes = 16 * 2 * nh * 16;        // 16 k-steps of B
  const uint32_t tma_off = (b_off + b_bytes + 1023) / 1024 * 1024;   // 4 x 7168 B landing slots for concurrent bulk copies
  const uint32_t sts_off = tma_off + 4 * SLOT_B;                     // 16 KB scratch for concurrent st.shared traffic
  const uint32_t bar_off = sts_off + 16384, tptr_off = bar_off + 64;
  for (uint32_t i = threadIdx.x * 16; i < bar_off; i += blockDim.x * 16) st_shared_v4(sbase + i, 0u, 0u, 0u, 0u);
  if (threadIdx.x == 0) { for (int i = 0; i < 6; ++i) mbar_init(sbase + bar_off + 8 * i, 1); fence_barrier_init(); }
  volatile uint32_t* stop = reinterpret_cast<volatile uint32_t*>(smem + bar_off + 48);
  if (threadIdx.x == 0) *stop = 0;
  if (threadIdx.x < 32) tmem_alloc<CG>(sbase + tptr_off, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sbase + tptr_off) : "memory");
  long long t0 = 0;
  if (threadIdx.x < 32 && rank == 0) {
    t0 = clock64();
    const uint32_t idesc = make_idesc(128 * CG, N);
    const uint32_t act_lo = dlo(sbase + a_off, CHUNK_B);
    const uint32_t w_lo = dlo(sbase + b_off, nh * 16);
    for (int i = 0; i < n_mma; i += 32) {
      const uint32_t m = (uint32_t)(i >> 5) & 3u;
      const uint32_t d_tmem = tmem_base + (N <= 128 ? m * 128u : 0u);
      const uint32_t a_m = act_lo + (same_operands ? 0u : m * 128u);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 32; ++ks) {
          const uint32_t alo = a_m + (same_operands ? 0u : ((uint32_t)(2 * (ks % 6)) * CHUNK_B + (uint32_t)(ks / 6) * ROW_B) / 16);
          const uint32_t blo = w_lo + (same_operands ? 0u : (uint32_t)(ks % 16) * 2 * nh);
          umma_bf16<CG>(d_tmem, dfull(alo), dfull(blo), idesc, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) {
      if (CG == 2) umma_commit_pair(sbase + bar_off, 3); else umma_commit_1(sbase + bar_off);
    }
    __syncwarp();
  }
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && tma_on && (threadIdx.x & 31) == 0) {
    // concurrent bulk copies global -> shared, 4 in flight, until the MMA stream has drained
    uint32_t n = 0, ph = 0;
    for (int i = 0; i < 4; ++i) { mbar_arrive_expect_tx(sbase + bar_off + 8 * (1 + i), SLOT_B); bulk_g2s(sbase + tma_off + i * SLOT_B, src + (size_t)((blockIdx.x * 4 + i) % 64) * SLOT_B, SLOT_B, sbase + bar_off + 8 * (1 + i)); }
    while (!*stop) {
      for (int i = 0; i < 4; ++i) {
        mbar_wait(sbase + bar_off + 8 * (1 + i), ph, nullptr, 9);
        ++n;
        mbar_arrive_expect_tx(sbase + bar_off + 8 * (1 + i), SLOT_B);
        bulk_g2s(sbase + tma_off + i * SLOT_B, src + (size_t)((n + blockIdx.x) % 64) * SLOT_B, SLOT_B, sbase + bar_off + 8 * (1 + i));
      }
      ph ^= 1;
    }
    for (int i = 0; i < 4; ++i) mbar_wait(sbase + bar_off + 8 * (1 + i), ph, nullptr, 9);
    if (rank == 0) cycles[128 + blockIdx.x / CG] = (long long)n * SLOT_B;
  } else if (warp >= 2 && warp < 2 + sts_warps) {
    uint32_t n = 0;
    const uint32_t dst = sbase + sts_off + (uint32_t)(warp - 2) * 2048 + (threadIdx.x & 31) * 16;
    while (!*stop) {
#pragma unroll
      for (int u = 0; u < 4; ++u) st_shared_v4(dst + u * 512, n, n, n, n);
      ++n;
    }
    if (rank == 0 && (threadIdx.x & 31) == 0 && warp == 2) cycles[192 + blockIdx.x / CG] = (long long)n * 2048 * sts_warps;
  }
  if (warp == 0) {
    mbar_wait(sbase + bar_off, 0, nullptr, 9);
    if (threadIdx.x == 0) { if (rank == 0) cycles[blockIdx.x / CG] = clock64() - t0; *stop = 1; }
  }
  __syncthreads();
  tc_fence_after();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<CG>(tmem_base, 512); }
}
__global__ void __launch_bounds__(256, 1) probe_rate_kernel_1(int N, int n_mma, int same, long long* cycles, const uint8_t* src, int tma_on, int sts_warps) { probe_rate_body<1>(N, n_mma, same, cycles, src, tma_on, sts_warps); }
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) probe_rate_kernel_2(int N, int n_mma, int same, long long* cycles, const uint8_t* src, int tma_on, int sts_warps) { probe_rate_body<2>(N, n_mma, same, cycles, src, tma_on, sts_warps); }

static unsigned long long* g_timeline = nullptr;

static uint32_t stack_image_bytes(const TaeDecConfig& c) {
  return 2 * (L0_B + (uint32_t)(c.num_layer - 1) * SLOTS_CONV * SLOT_B + LIN_B);
}
static uint32_t stack_bwd_image_bytes(const TaeDecConfig& c) {
  return 2 * (L0_B + (uint32_t)(c.num_layer - 1) * SLOTS_CONV * SLOT_B + 2 * FIN_B);
}

}  // namespace

bool dec_pair_supported(const TaeDecConfig& c, const char** why) {
  static thread_local char msg[160];
  *why = msg;
  if (c.kernel_size != TAPS) { snprintf(msg, sizeof msg, "kernel_size %d (only 5 is built for the tensor path)", c.kernel_size); return false; }
  if (c.num_unit > UNITS_MAX || c.num_unit < 1) { snprintf(msg, sizeof msg, "num_unit %d > %d", c.num_unit, UNITS_MAX); return false; }
  if (c.num_iter_ft > 5) { snprintf(msg, sizeof msg, "num_iter_ft %d > 5", c.num_iter_ft); return false; }
  if (c.num_layer < 2) { snprintf(msg, sizeof msg, "num_layer %d < 2", c.num_layer); return false; }
  if (c.block_len > GROUP_ROWS) { snprintf(msg, sizeof msg, "block_len %d > %d (one codeword must fit a 512-row group)", c.block_len, GROUP_ROWS); return false; }
  if (make_smem(c.num_iter_ft).total > 227 * 1024) { snprintf(msg, sizeof msg, "shared memory budget exceeded"); return false; }
  *why = nullptr;
  return true;
}

// The packed buffer holds TWO images: [0] the inference image (activations scaled by log2(e), see conv_w_elem) and [1] the
// plain image of the training forward (whose stashed activations feed the backward and weight-gradient kernels unscaled).
size_t dec_pair_packed_bytes(const TaeDecConfig& c) { return (size_t)2 * 2 * c.num_iteration * stack_image_bytes(c); }

int dec_pair_pack(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s) {
  const int n_stacks = 2 * c.num_iteration;
  const PackLayout lay{n_stacks, c.num_layer, c.num_unit, 2 + c.num_iter_ft, c.num_iter_ft, 1};
  const uint32_t stack_elems = stack_image_bytes(c) / 2;
  const size_t total = (size_t)n_stacks * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  for (int scaled = 1; scaled >= 0; --scaled) {
    pack_pair_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<__nv_bfloat16*>(packed) + (scaled ? 0 : total), lay, n_stacks,
                                            c.num_layer, c.num_unit, 2 + c.num_iter_ft, stack_elems, scaled);
    int rc = after_launch("pack_pair_kernel");
    if (rc) return rc;
  }
  return TAE_OK;
}

static int pair_launch_setup(const TaeDecConfig&, int* n_sm_out) {
  static DeviceOnce once;
  return device_once(once, "dec_pair_kernel", [](int dev) -> int {
    int rc = require_sm100(dev, "the bf16 tensor path");
    if (rc) return rc;
    const int smem = (int)make_smem(5).total;
    cudaError_t e = cudaFuncSetAttribute(dec_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dec_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(dec_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(dec_pair_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    return TAE_OK;
  }, n_sm_out);
}

int dec_forward_pair(const TaeDecConfig& c, const void* packed, const float* received, const int32_t* perm,
                     const int32_t* inv_perm, float* out, float* trace, int B, void* ws, size_t ws_bytes, cudaStream_t s,
                     void* stash_y, void* stash_x) {
  if (ws_bytes < 256) { set_error("tae_dec_forward(bf16): workspace %zu < 256 bytes", ws_bytes); return TAE_EWORKSPACE; }
  const Smem S = make_smem(c.num_iter_ft);
  int n_sm = 0;
  {
    int rc = pair_launch_setup(c, &n_sm);
    if (rc) return rc;
  }
  PairArgs a{};
  a.wimg = reinterpret_cast<const uint8_t*>(packed) + (stash_y ? dec_pair_packed_bytes(c) / 2 : 0);
  a.received = received;
  a.out = out;
  a.trace = trace;
  a.err = wait_code_slot(ws);
  a.perm = perm;
  a.inv_perm = inv_perm;
  a.B = B; a.L = c.block_len; a.F = c.num_iter_ft; a.n_stacks = 2 * c.num_iteration; a.n_layer = c.num_layer;
  a.enc = 0;
  a.extrinsic = c.extrinsic;
  a.cw_per_group = (GROUP_ROWS + 2) / (c.block_len + 2);
  a.n_groups = (B + a.cw_per_group - 1) / a.cw_per_group;
  a.n_pairs = (a.n_groups + 1) / 2;
  a.stack_bytes = stack_image_bytes(c);
  a.tl = g_timeline;
  a.stash_y = reinterpret_cast<uint8_t*>(stash_y);
  a.stash_x = reinterpret_cast<uint8_t*>(stash_x);
  const int n_clusters = std::min(a.n_pairs, n_sm / 2);
  if (stash_y) dec_pair_kernel<2><<<2 * n_clusters, N_THREADS, S.total, s>>>(a);
  else dec_pair_kernel<0><<<2 * n_clusters, N_THREADS, S.total, s>>>(a);
  return after_launch("dec_pair_kernel");
}

// ---- training: group-image geometry, backward weight image, backward of one stack ---------------------------------
int train_groups(int block_len, int B) {
  const int cpg = (GROUP_ROWS + 2) / (block_len + 2);
  return (B + cpg - 1) / cpg;
}

size_t dec_pair_bwd_packed_bytes(const TaeDecConfig& c) { return (size_t)2 * c.num_iteration * stack_bwd_image_bytes(c); }

int dec_pair_pack_bwd(const TaeDecConfig& c, const float* params, void* packed, cudaStream_t s) {
  const int n_stacks = 2 * c.num_iteration;
  const PackLayout lay{n_stacks, c.num_layer, c.num_unit, 2 + c.num_iter_ft, c.num_iter_ft, 1};
  const uint32_t stack_elems = stack_bwd_image_bytes(c) / 2;
  const size_t total = (size_t)n_stacks * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  pack_bwd_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<__nv_bfloat16*>(packed), lay, n_stacks, c.num_layer,
                                         c.num_unit, 2 + c.num_iter_ft, stack_elems);
  return after_launch("pack_bwd_kernel");
}

// Backward of all 2I stacks in ONE launch (schedule walked backwards; the glue between stacks runs in the kernel's per-stack
// prologue): d_out_last (B, L, fout of the last stack) -> dxin_all (2I, B, L, 8); dlin_all (2I, B, L, F) is scratch for the chain;
// the Linear bias gradients are added into grad_flat (flat parameter layout) when it is not NULL.
static int backward_pair(const TaeDecConfig& c, const PackLayout& lay, int n_stacks, int chain, const void* packed_bwd, const float* dlin,
                         const int32_t* perm, const int32_t* inv_perm, const void* stash_y, void* stash_g, void* stash_d, float* dxin_all,
                         float* dlin_all, float* grad_flat, int B, void* ws, size_t ws_bytes, cudaStream_t s, const char* who,
                         int pair_begin = 0, int pair_end = -1) {
  if (ws_bytes < 256) { set_error("%s: workspace %zu < 256 bytes", who, ws_bytes); return TAE_EWORKSPACE; }
  const Smem S = make_smem(c.num_iter_ft);
  int n_sm = 0;
  {
    int rc = pair_launch_setup(c, &n_sm);
    if (rc) return rc;
  }
  PairArgs a{};
  a.B = B; a.L = c.block_len; a.F = c.num_iter_ft; a.n_stacks = n_stacks; a.n_layer = c.num_layer; a.extrinsic = c.extrinsic;
  a.cw_per_group = (GROUP_ROWS + 2) / (c.block_len + 2);
  a.n_groups = (B + a.cw_per_group - 1) / a.cw_per_group;
  a.n_pairs = (a.n_groups + 1) / 2;
  a.stack_bytes = stack_bwd_image_bytes(c);
  a.wimg = reinterpret_cast<const uint8_t*>(packed_bwd);
  a.err = wait_code_slot(ws);
  a.stash_y = const_cast<uint8_t*>(reinterpret_cast<const uint8_t*>(stash_y));
  a.stash_g = reinterpret_cast<uint8_t*>(stash_g);
  a.stash_x = reinterpret_cast<uint8_t*>(stash_d);
  a.dlin = dlin; a.dxin = dxin_all; a.dlin_all = dlin_all; a.grad_flat = grad_flat; a.chain = chain; a.lay = lay;
  a.perm = perm; a.inv_perm = inv_perm;
  if (pair_end < 0 || pair_end > a.n_pairs) pair_end = a.n_pairs;
  if (pair_begin < 0 || pair_begin > pair_end) { set_error("%s: bad work-unit range [%d, %d) of %d", who, pair_begin, pair_end, a.n_pairs); return TAE_EINVAL; }
  if (pair_begin == pair_end) return TAE_OK;
  a.pair_begin = pair_begin;
  a.n_pairs = pair_end;                          // the kernel's unit loops end here
  const int n_clusters = std::min(pair_end - pair_begin, n_sm / 2);
  dec_pair_kernel<1><<<2 * n_clusters, N_THREADS, S.total, s>>>(a);
  return after_launch("dec_pair_kernel<1>");
}

int dec_backward_pair(const TaeDecConfig& c, const void* packed_bwd, const float* d_out_last, const int32_t* perm, const int32_t* inv_perm,
                      const void* stash_y, void* stash_g, void* stash_d, float* dxin_all, float* dlin_all, float* grad_flat, int B, void* ws,
                      size_t ws_bytes, cudaStream_t s, int pair_begin, int pair_end) {
  const int n_stacks = 2 * c.num_iteration;
  const PackLayout lay{n_stacks, c.num_layer, c.num_unit, 2 + c.num_iter_ft, c.num_iter_ft, 1};
  return backward_pair(c, lay, n_stacks, 1, packed_bwd, d_out_last, perm, inv_perm, stash_y, stash_g, stash_d, dxin_all, dlin_all, grad_flat, B,
                       ws, ws_bytes, s, "tae_dec_backward_bf16", pair_begin, pair_end);
}

// ---- ENC_interCNN on the same kernel ----------------------------------------------------------------------------
static TaeDecConfig enc_as_dec(const TaeEncConfig& c) {
  TaeDecConfig d{};
  d.block_len = c.block_len; d.num_iteration = 2; d.num_iter_ft = 1; d.num_layer = c.num_layer; d.num_unit = c.num_unit;
  d.kernel_size = c.kernel_size; d.extrinsic = 0;
  return d;
}

bool enc_pair_supported(const TaeEncConfig& c, const char** why) {
  static thread_local char msg[160];
  const TaeDecConfig d = enc_as_dec(c);
  if (!dec_pair_supported(d, why)) return false;
  *why = msg;
  if (c.num_layer < 2) { snprintf(msg, sizeof msg, "enc_num_layer %d < 2", c.num_layer); return false; }
  *why = nullptr;
  return true;
}

size_t enc_pair_packed_bytes(const TaeEncConfig& c) { return (size_t)2 * 3 * stack_image_bytes(enc_as_dec(c)); }   // [inference | training]

int enc_pair_pack(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s) {
  const PackLayout lay{3, c.num_layer, c.num_unit, 1, 1, 1};
  const TaeDecConfig d = enc_as_dec(c);
  const uint32_t stack_elems = stack_image_bytes(d) / 2;
  const size_t total = (size_t)3 * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  for (int scaled = 1; scaled >= 0; --scaled) {
    pack_pair_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<__nv_bfloat16*>(packed) + (scaled ? 0 : total), lay, 3, c.num_layer,
                                            c.num_unit, 1, stack_elems, scaled);
    int rc = after_launch("pack_pair_kernel");
    if (rc) return rc;
  }
  return TAE_OK;
}

size_t enc_pair_bwd_packed_bytes(const TaeEncConfig& c) { return (size_t)3 * stack_bwd_image_bytes(enc_as_dec(c)); }

int enc_pair_pack_bwd(const TaeEncConfig& c, const float* params, void* packed, cudaStream_t s) {
  const PackLayout lay{3, c.num_layer, c.num_unit, 1, 1, 1};
  const TaeDecConfig d = enc_as_dec(c);
  const uint32_t stack_elems = stack_bwd_image_bytes(d) / 2;
  const size_t total = (size_t)3 * stack_elems;
  const int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 16);
  pack_bwd_kernel<<<blocks, 256, 0, s>>>(params, reinterpret_cast<__nv_bfloat16*>(packed), lay, 3, c.num_layer, c.num_unit, 1, stack_elems);
  return after_launch("pack_bwd_kernel");
}

// Backward of the three encoder branches in one launch: dlin (3, B, L, 1) -> dxin_all (3, B, L, 8); independent stacks (no chain).
int enc_backward_pair(const TaeEncConfig& c, const void* packed_bwd, const float* dlin, const void* stash_y, void* stash_g, void* stash_d,
                      float* dxin_all, float* grad_flat, int B, void* ws, size_t ws_bytes, cudaStream_t s) {
  const PackLayout lay{3, c.num_layer, c.num_unit, 1, 1, 1};
  return backward_pair(enc_as_dec(c), lay, 3, 0, packed_bwd, dlin, nullptr, nullptr, stash_y, stash_g, stash_d, dxin_all, nullptr, grad_flat, B,
                       ws, ws_bytes, s, "tae_enc_backward_bf16");
}

int enc_forward_pair(const TaeEncConfig& c, const void* packed, const float* u, const int32_t* perm, const int32_t* inv_perm,
                     float* x_tx, double* stats, int B, void* ws, size_t ws_bytes, cudaStream_t s, void* stash_y, void* stash_x) {
  if (ws_bytes < 256) { set_error("tae_enc_forward_bf16: workspace %zu < 256 bytes", ws_bytes); return TAE_EWORKSPACE; }
  const TaeDecConfig d = enc_as_dec(c);
  // reuse the decoder launcher's one-time setup by going through the same code path
  PairArgs a{};
  int n_sm = 0;
  int rc = pair_launch_setup(d, &n_sm);
  if (rc) return rc;
  a.wimg = reinterpret_cast<const uint8_t*>(packed) + (stash_y ? enc_pair_packed_bytes(c) / 2 : 0);
  a.u = u; a.x_tx = x_tx; a.stats = stats; a.enc = 1;
  a.err = wait_code_slot(ws);
  a.perm = perm; a.inv_perm = inv_perm;
  a.B = B; a.L = c.block_len; a.F = 1; a.n_stacks = 3; a.n_layer = c.num_layer; a.extrinsic = 0;
  a.cw_per_group = (GROUP_ROWS + 2) / (c.block_len + 2);
  a.n_groups = (B + a.cw_per_group - 1) / a.cw_per_group;
  a.n_pairs = (a.n_groups + 1) / 2;
  a.stack_bytes = stack_image_bytes(d);
  a.tl = nullptr;
  a.stash_y = reinterpret_cast<uint8_t*>(stash_y);
  a.stash_x = reinterpret_cast<uint8_t*>(stash_x);
  const int n_clusters = std::min(a.n_pairs, n_sm / 2);
  if (stash_y) dec_pair_kernel<2><<<2 * n_clusters, N_THREADS, make_smem(1).total, s>>>(a);
  else dec_pair_kernel<0><<<2 * n_clusters, N_THREADS, make_smem(1).total, s>>>(a);
  return after_launch("dec_pair_kernel(enc)");
}

}  // namespace tae

// ---- self-test entry points (not part of the drop-in surface) ------------------------------------------
extern "C" {

// When non-NULL, later CTA-pair decodes record clock64 stamps of cluster 0's leader CTA: (steps*4, 8) uint64
// [mma wait start, mma wait done, mma issued+committed, epi warp0 wait start, acc ready, done, epi warp7 acc ready, done].
void tae_debug_set_timeline(unsigned long long* dev) { tae::g_timeline = dev; }

// D (128, N) = A_eff @ Bm^T, A_eff[r, 8j + e] = X[shift + r + j*lbo_rows, e] (j = 0, 1).  X: (R, 8) bf16, Bm: (N, 16) bf16.
int tae_debug_probe_lbo(const void* X, const void* Bm, float* D, int32_t R, int32_t N, int32_t shift, int32_t lbo_rows,
                        int* err, void* stream) {
  using namespace tae;
  if (N % 16 || N > 128 || shift < 0 || lbo_rows < 1 || shift + lbo_rows + 128 > R) { set_error("probe_lbo: bad shape"); return TAE_EINVAL; }
  const size_t smem = (size_t)R * 16 + 2 * (size_t)N * 16 + 64;
  cudaError_t e = cudaFuncSetAttribute(probe_lbo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(probe_lbo): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  if (smem > 100 * 1024) { set_error("probe_lbo: too large"); return TAE_EINVAL; }
  probe_lbo_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(X),
                                                         reinterpret_cast<const __nv_bfloat16*>(Bm), D, R, N, shift, lbo_rows, err);
  return after_launch("probe_lbo_kernel");
}

// Issue-rate probe: `grid` CTAs (CG = 1) or CTA pairs (CG = 2), each streaming n_mma MMAs of shape (128*CG, N, 16).
int tae_debug_probe_rate(int32_t cg, int32_t N, int32_t n_mma, int32_t same_operands, int32_t grid, long long* cycles, const void* src,
                         int32_t tma_on, int32_t sts_warps, void* stream) {
  using namespace tae;
  if ((cg != 1 && cg != 2) || N % 16 || N < 16 || N > 256 || n_mma % 4) { set_error("probe_rate: bad arguments"); return TAE_EINVAL; }
  const size_t smem = 12 * CHUNK_B + 16 * 2 * (size_t)(N / cg) * 16 + 1024 + 4 * SLOT_B + 16384 + 128;
  if (sts_warps < 0 || sts_warps > 6 || (tma_on && !src)) { set_error("probe_rate: bad arguments"); return TAE_EINVAL; }
  cudaError_t e = cg == 1 ? cudaFuncSetAttribute(probe_rate_kernel_1, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                          : cudaFuncSetAttribute(probe_rate_kernel_2, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(probe_rate): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  const uint8_t* sp = reinterpret_cast<const uint8_t*>(src);
  if (cg == 1) probe_rate_kernel_1<<<grid, 256, smem, (cudaStream_t)stream>>>(N, n_mma, same_operands, cycles, sp, tma_on, sts_warps);
  else probe_rate_kernel_2<<<2 * grid, 256, smem, (cudaStream_t)stream>>>(N, n_mma, same_operands, cycles, sp, tma_on, sts_warps);
  return after_launch("probe_rate_kernel");
}

// D (256, N) = A (256, K) @ Bm (N, K)^T with one cta_group::2 MMA chain.
int tae_debug_probe_pair(const void* A, const void* Bm, float* D, int32_t K, int32_t N, int* err, void* stream) {
  using namespace tae;
  if (K % 16 || N % 16 || N > 128 || K > 128) { set_error("probe_pair: bad shape"); return TAE_EINVAL; }
  const size_t smem = (size_t)(K / 8) * 128 * 16 + (size_t)(K / 8) * (N / 2) * 16 + 64;
  cudaError_t e = cudaFuncSetAttribute(probe_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(probe_pair): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
  probe_pair_kernel<<<2, 128, smem, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(A),
                                                          reinterpret_cast<const __nv_bfloat16*>(Bm), D, K, N, err);
  return after_launch("probe_pair_kernel");
}

}  // extern "C"
