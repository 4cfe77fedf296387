/*
 * turboae_b200.h -- C ABI of libturboae_b200.so
 *
 * B200 (sm_100a) implementation of ONE hot path of yihanjiang/turboae: the rate-1/3 CNN
 * encoder ENC_interCNN, the iterative CNN turbo decoder DEC_LargeCNN and the interleaver
 * gather they share.  The reference is pure Python/PyTorch and has no FFI; every entry
 * point below names the reference lines (paths relative to the reference checkout) whose
 * arithmetic it replaces.  Host language above this boundary is Python
 * (turboae_b200/*.py, loaded with ctypes); INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - tensors are dense, row-major, float32, "channel-last" exactly as the reference's
 *    nn.Module.forward() sees them: (B, L, C);
 *  - calls only ENQUEUE work on `stream` (a cudaStream_t passed as void*) and never
 *    synchronise; inputs/outputs/workspace are borrowed for the duration of that work;
 *  - return value: 0 (TAE_OK) or a negative TAE_E* code; tae_last_error() returns a
 *    thread-local message for the last failing call.  There is no CPU fallback.
 */
#ifndef TURBOAE_B200_H_
#define TURBOAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TAE_OK            0
#define TAE_EINVAL       -1   /* bad pointer / shape / size                            */
#define TAE_EUNSUPPORTED -2   /* configuration outside what the kernels implement      */
#define TAE_ECUDA        -3   /* CUDA runtime error (message has cudaGetErrorString)   */
#define TAE_EWORKSPACE   -4   /* workspace smaller than tae_*_workspace_bytes() asks   */

#define TAE_PRECISION_FP32 0  /* CUDA-core fp32 FMA path: elementwise parity (<=1e-4)  */
#define TAE_PRECISION_BF16 1  /* tcgen05 bf16 operands, fp32 TMEM accumulation          */
#define TAE_PRECISION_F16X3 2 /* tcgen05, every operand split into fp16 hi + lo (three MMA chains per layer), bias / ELU /
                                  Linear / priors in fp32: elementwise parity (<=1e-4) at tensor-core speed; activations
                                  beyond +-65504 (fp16's range) are clamped                                               */

/* Shape of a DEC_LargeCNN (reference decoders.py:158-192; get_args.py:83-84,89-100,122). */
typedef struct TaeDecConfig {
  int32_t block_len;      /* L            args.block_len                                 */
  int32_t num_iteration;  /* I            args.num_iteration                             */
  int32_t num_iter_ft;    /* F            args.num_iter_ft                               */
  int32_t num_layer;      /*              args.dec_num_layer                             */
  int32_t num_unit;       /*              args.dec_num_unit                              */
  int32_t kernel_size;    /* odd          args.dec_kernel_size                           */
  int32_t extrinsic;      /* 0/1          args.extrinsic (decoders.py:235,246,257)       */
} TaeDecConfig;

/* Shape of an ENC_interCNN (reference encoders.py:307-337). */
typedef struct TaeEncConfig {
  int32_t block_len;      /* L                                                           */
  int32_t num_layer;      /* args.enc_num_layer                                          */
  int32_t num_unit;       /* args.enc_num_unit                                           */
  int32_t kernel_size;    /* args.enc_kernel_size (odd)                                  */
} TaeEncConfig;

/* ---- library ------------------------------------------------------------------------ */
int          tae_version(void);
const char*  tae_last_error(void);
/* Number of kernels this library has launched in this process (bench.py "gpu_launches"). */
uint64_t     tae_launch_count(void);

/* ---- a1/a2: Interleaver.forward / DeInterleaver.forward ----------------------------------
 * reference interleavers.py:15-21 and :43-48.  out[b,i,f] = in[b,perm[i],f] for an int32
 * device permutation of length L; pass the inverse permutation (rp[p[i]] = i,
 * interleavers.py:29-33) to de-interleave.  Pure data movement: bit-exact.              */
int tae_interleave_f32(const float* in, float* out, const int32_t* perm,
                       int32_t B, int32_t L, int32_t F, void* stream);

/* ---- a3: one layer of SameShapeConv1d.forward ---------------------------------------------
 * reference cnn_utils.py:36-46 (conv built at :15-22): channel-last
 * y[b,l,o] = act(bias[o] + sum_c sum_t W[o,c,t] * x[b,l+t-K/2,c]), zero padded, act = ELU
 * (alpha 1) when apply_elu != 0.  `weight` is torch's Conv1d layout (Cout, Cin, K).
 * `workspace` needs tae_conv1d_workspace_bytes(Cin, Cout, K) bytes.                      */
size_t tae_conv1d_workspace_bytes(int32_t Cin, int32_t Cout, int32_t K);
int tae_conv1d_elu_f32(const float* in, float* out, const float* weight, const float* bias,
                       int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t K,
                       int32_t apply_elu, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of that layer (training, reference trainer.py:74 `loss.backward()` through cnn_utils.py:36-46):
 *   x (B,L,Cin) = the layer's input, y (B,L,Cout) = its forward output, dy = gradient w.r.t. y.
 *   dx (B,L,Cin) or NULL; dweight (Cout,Cin,K) and dbias (Cout) are ACCUMULATED into (caller zeroes), NULL to skip.
 *   With apply_elu the ELU derivative is taken from y (ELU'(z) = y > 0 ? 1 : y + 1).                              */
size_t tae_conv1d_bwd_workspace_bytes(int32_t Cin, int32_t Cout, int32_t K);
int tae_conv1d_elu_bwd_f32(const float* x, const float* y, const float* dy, const float* weight,
                           float* dx, float* dweight, float* dbias,
                           int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t K, int32_t apply_elu,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- a7/a8: DEC_LargeCNN -------------------------------------------------------------------
 * Parameters travel as ONE flat float32 buffer in canonical order:
 *   for idx in 0..I-1: for s in (dec1, dec2):
 *       for j in 0..num_layer-1: cnns[j].weight (Cout,Cin,K) then cnns[j].bias (Cout)
 *       outputs[idx].weight (Fout, num_unit) then outputs[idx].bias (Fout)
 * with Cin = 2+F for j == 0 else num_unit, Fout = F except dec2 of the last iteration
 * (Fout = 1) -- the order of reference decoders.py:178-192.
 */
size_t tae_dec_param_count(const TaeDecConfig* cfg);
/* Size of / conversion into the bf16 tensor-core weight image (UMMA canonical K-major
 * layout, zero padded), consumed by TAE_PRECISION_BF16.  Rebuild after every weight update. */
size_t tae_dec_packed_bytes(const TaeDecConfig* cfg);
int    tae_dec_pack_bf16(const TaeDecConfig* cfg, const float* params, void* packed, void* stream);
size_t tae_dec_workspace_bytes(const TaeDecConfig* cfg, int32_t B, int32_t precision);
/* DEC_LargeCNN.forward, reference decoders.py:206-269.
 *   received (B,L,3) -> out (B,L,1) = sigmoid posteriors.
 *   perm / inv_perm: int32[L] device arrays (Interleaver / DeInterleaver index).
 *   packed: result of tae_dec_pack_bf16 (may be NULL for TAE_PRECISION_FP32).
 *   trace: NULL, or (2I, B, L, F) floats receiving the output of every dec{1,2}_outputs
 *          Linear before the extrinsic subtraction (last one has 1 feature, stored in [..,0]). */
int tae_dec_forward(const TaeDecConfig* cfg, const float* params, const void* packed,
                    const float* received, const int32_t* perm, const int32_t* inv_perm,
                    float* out, float* trace, int32_t B, int32_t precision,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Weight image of TAE_PRECISION_F16X3 (W_hi and W_lo slots per layer; pass it as `packed` to tae_dec_forward, which
 * also reads the fp32 biases and Linear weights from `params`).  Rebuild after every weight update.               */
size_t tae_dec_packed_bytes_x3(const TaeDecConfig* cfg);
int    tae_dec_pack_f16x3(const TaeDecConfig* cfg, const float* params, void* packed, void* stream);

/* DEC_LargeCNN.forward for HOST buffers (reference decoders.py:219 moves `received` to the device itself, and
 * trainer.py:176-177 reads the result back): `received_host` / `out_host` are host pointers (pinned memory makes the
 * copies asynchronous).  The batch is cut into chunks whose H2D copy, decode and D2H copy overlap on internal
 * streams; on return everything is ENQUEUED and `stream` waits for the last D2H copy, so synchronising `stream`
 * makes `out_host` valid.  `workspace` (device) needs tae_dec_host_workspace_bytes().                          */
size_t tae_dec_host_workspace_bytes(const TaeDecConfig* cfg, int32_t B, int32_t precision);
int tae_dec_forward_host(const TaeDecConfig* cfg, const float* params, const void* packed,
                         const float* received_host, const int32_t* perm, const int32_t* inv_perm,
                         float* out_host, int32_t B, int32_t precision,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- a4/a5/a6: ENC_interCNN ----------------------------------------------------------------
 * Flat parameter order: for branch in 1..3: for j: enc_cnn_b.cnns[j].weight, .bias; then
 * enc_linear_b.weight (1,num_unit), .bias (1)   (reference encoders.py:314-335).
 */
size_t tae_enc_param_count(const TaeEncConfig* cfg);
size_t tae_enc_workspace_bytes(const TaeEncConfig* cfg, int32_t B);
/* Branches + concat, reference encoders.py:362-373: u (B,L,1) in {0,1} -> x_tx (B,L,3),
 * NOT yet normalised.  Adds (sum x, sum x^2, count) of this call's x_tx into stats[0..2]
 * (device doubles; caller zeroes them, and all-reduces them across ranks when the batch is
 * sharded, because power_constraint normalises over the WHOLE batch).                     */
int tae_enc_forward(const TaeEncConfig* cfg, const float* params, const float* u,
                    const int32_t* perm, float* x_tx, double* stats, int32_t B,
                    void* workspace, size_t workspace_bytes, void* stream);
/* The same forward on the tensor cores (bf16 operands, fp32 accumulation; the fused CTA-pair kernel of the decoder with
 * the three branches as three conv stacks).  `packed` from tae_enc_pack_bf16 (rebuild after every weight update);
 * needs perm AND inverse perm; workspace >= 256 bytes.  Same x_tx / stats contract as tae_enc_forward.            */
size_t tae_enc_packed_bytes(const TaeEncConfig* cfg);
int    tae_enc_pack_bf16(const TaeEncConfig* cfg, const float* params, void* packed, void* stream);
int tae_enc_forward_bf16(const TaeEncConfig* cfg, const void* packed, const float* u, const int32_t* perm,
                         const int32_t* inv_perm, float* x_tx, double* stats, int32_t B,
                         void* workspace, size_t workspace_bytes, void* stream);
/* The same forward with split fp16 operands (x_hi W_hi + x_lo W_hi + x_hi W_lo on the tensor cores, fp32 bias / ELU /
 * Linear): codes within 1e-5 of the fp32 path (fp16 hi + lo terms).  `packed` from tae_enc_pack_f16x3; `params` supplies the fp32 biases and
 * Linear weights.  Same x_tx / stats contract as tae_enc_forward; workspace >= 256 bytes.                           */
size_t tae_enc_packed_bytes_x3(const TaeEncConfig* cfg);
int    tae_enc_pack_f16x3(const TaeEncConfig* cfg, const float* params, void* packed, void* stream);
int tae_enc_forward_f16x3(const TaeEncConfig* cfg, const float* params, const void* packed, const float* u,
                           const int32_t* perm, const int32_t* inv_perm, float* x_tx, double* stats, int32_t B,
                           void* workspace, size_t workspace_bytes, void* stream);
/* ENCBase.power_constraint default branch, reference encoders.py:107-116:
 * codes = (x - mean) / std, mean/std (unbiased, N-1) derived on the device from stats[0..2].
 * mean_std: NULL or 2 device floats receiving (mean, std).  x may alias codes.            */
int tae_power_norm_f32(const float* x, float* codes, size_t n, const double* stats,
                       float* mean_std, void* stream);

/* ENCBase.power_constraint with precompute_norm_stats, reference encoders.py:110-114: codes = (x - mean) / std with the
 * GIVEN running (mean, std) -- 2 device floats the caller keeps up to date -- followed, when quantize_level >= 2, by the
 * STE quantiser's forward (encoders.py:118-120).  quantize_level 0: no quantiser.  x may alias codes.                  */
int tae_power_norm_given_f32(const float* x, float* codes, size_t n, const float* mean_std, float value_limit,
                             float quantize_level, void* stream);
/* The same followed by STEQuantize.forward (reference encoders.py:20-37, applied at :118-120 when
 * train_channel_mode == 'block_norm_ste'): clamp to +-value_limit, then sign() for quantize_level == 2 or
 * quantize_level uniform levels otherwise.  (The straight-through backward is torch glue in the Python layer.)     */
int tae_power_norm_ste_f32(const float* x, float* codes, size_t n, const double* stats, float* mean_std,
                           float value_limit, float quantize_level, void* stream);

/* ENCBase.power_constraint under autograd (reference trainer.py:74 `loss.backward()` through encoders.py:107-116), for an x_tx
 * that is already on the device:
 *   tae_power_stats_f32          adds (sum x, sum x^2, n) into the 3 device doubles `stats` (the encoder kernels above deliver the
 *                                same triple; this entry point serves an x_tx produced elsewhere).  Sharded batch: all-reduce
 *                                `stats` over the ranks, then tae_power_norm_f32 with mean_std != NULL.
 *   tae_power_norm_bwd_sums_f32  adds (sum g, sum g * codes) into the 2 device doubles `sums` (g = gradient w.r.t. codes);
 *                                all-reduced over the ranks when the batch is sharded.
 *   tae_power_norm_bwd_f32       dx = (g - sums[0] / N - codes * sums[1] / (N - 1)) / std, N = stats[2], std = mean_std[1]:
 *                                the exact gradient of codes = (x - mean(x)) / std_unbiased(x) w.r.t. x.  dx may alias g.   */
int tae_power_stats_f32(const float* x, size_t n, double* stats, void* stream);
int tae_power_norm_bwd_sums_f32(const float* g, const float* codes, size_t n, double* sums, void* stream);
int tae_power_norm_bwd_f32(const float* g, const float* codes, float* dx, size_t n, const double* sums, const double* stats,
                           const float* mean_std, void* stream);

/* ---- next row f1 on the tensor cores: training of DEC_LargeCNN (reference trainer.py:33-76: forward, loss.backward()
 * through decoders.py:219-269 and cnn_utils.py:36-46; bf16 operands, fp32 accumulation, fp32 gradients) ------------------
 * Activations travel between the three kernels as "group images" in HBM: bf16 [group][chunk][516 rows][8 channels], the
 * layout the fused kernel keeps in shared memory (rows: 2 halo rows, then per codeword L positions + 2 zero separator
 * rows; a group holds floor(514 / (L + 2)) codewords).  All image buffers must be ZERO-INITIALISED once by the caller.
 *   stash_y : [2I stacks][num_layer][groups][13 chunks]   forward output of every conv layer (after ELU)
 *   stash_x : [2I stacks][groups][1 chunk]                 the 2+F input channels of every stack
 *   stash_g : like stash_y                                  dL/dz (gradient at the pre-activation) of every conv layer
 *   stash_d : [2I stacks][groups][1 chunk]                 gradient w.r.t. every Linear output
 */
#define TAE_IMG_CHUNK_BYTES 8256
#define TAE_IMG_CHUNKS 13
int32_t tae_train_groups(int32_t block_len, int32_t B);
/* Work units of the fused kernels for a batch: one unit = the two groups 2p, 2p+1 = one pass of a CTA pair; (groups + 1) / 2. */
int32_t tae_train_units(int32_t block_len, int32_t B);
/* tae_dec_forward(TAE_PRECISION_BF16) that also writes stash_y / stash_x. */
int tae_dec_forward_train_bf16(const TaeDecConfig* cfg, const void* packed, const float* received, const int32_t* perm,
                               const int32_t* inv_perm, float* out, float* trace, int32_t B, void* stash_y, void* stash_x,
                               void* workspace, size_t workspace_bytes, void* stream);
/* Backward weight image (transposed, tap-flipped conv weights; transposed Linear).  Rebuild after every weight update. */
size_t tae_dec_bwd_packed_bytes(const TaeDecConfig* cfg);
int    tae_dec_pack_bwd_bf16(const TaeDecConfig* cfg, const float* params, void* packed_bwd, void* stream);
/* Backward of all 2I conv stacks + Linears in ONE launch, the turbo schedule walked backwards:
 *   d_out_last (B, L, 1) = gradient w.r.t. the last stack's Linear output (before deinterleave + sigmoid)
 *   dxin_all (2I, B, L, 8) receives the gradient w.r.t. every stack's 2+F inputs (columns 2+F.. are zero); the caller sums the
 *   sys / parity columns into d received.  Between stacks the kernel itself applies the backward of the extrinsic subtraction
 *   and the (de)interleaver (reference decoders.py:235-249): dlin_s[b,l,f] = dxin_{s+1}[b, idx[l], 2+f] - dlin_{s+1}[b, idx[l], f];
 *   dlin_all (2I, B, L, F) is scratch for that chain.  Reads stash_y, writes stash_g and stash_d.  grad_flat: NULL, or the flat
 *   gradient buffer (layout of the flat parameter buffer) into which the Linear bias gradients (sums of dlin) are ADDED.      */
int tae_dec_backward_bf16(const TaeDecConfig* cfg, const void* packed_bwd, const float* d_out_last, const int32_t* perm,
                          const int32_t* inv_perm, const void* stash_y, void* stash_g, void* stash_d, float* dxin_all,
                          float* dlin_all, float* grad_flat, int32_t B, void* workspace, size_t workspace_bytes, void* stream);
/* The same for the work units [unit_begin, unit_end) only (all buffers as for the whole batch).  Units are independent, so the
 * backward of a batch may be split over several launches: with more units than CTA pairs on the device (batch 1000: 100 units on
 * 74 pairs) the last, partly filled wave goes into a launch of its own and tae_wgrad_bf16 over the groups of the earlier units runs
 * beside it on the SMs that wave leaves idle (turboae_b200/train_tc.py).                                                        */
int tae_dec_backward_range_bf16(const TaeDecConfig* cfg, const void* packed_bwd, const float* d_out_last, const int32_t* perm,
                                const int32_t* inv_perm, const void* stash_y, void* stash_g, void* stash_d, float* dxin_all,
                                float* dlin_all, float* grad_flat, int32_t B, int32_t unit_begin, int32_t unit_end, void* workspace,
                                size_t workspace_bytes, void* stream);
/* Glue of loss.backward() (reference trainer.py:74) on either side of tae_dec_backward_bf16:
 *   tae_dec_out_backward_f32  out = sigmoid(deinterleave(o_last)) (decoders.py:263-267): d_out_last[b, i] =
 *                             (d_out * out * (1 - out))[b, perm[i]], all (B, L, 1).
 *   tae_dec_input_grad_f32    d_received (B, L, 3) from dxin_all (n_stacks, B, L, 8): even stacks read [r_sys, r_par1, priors]
 *                             (decoders.py:230), odd stacks [interleave(r_sys), r_par2, priors] (:222, :240).                  */
int tae_dec_out_backward_f32(const float* d_out, const float* out, const int32_t* perm, float* d_out_last, int32_t B, int32_t L,
                             void* stream);
int tae_dec_input_grad_f32(const float* dxin_all, const int32_t* inv_perm, float* d_received, int32_t n_stacks, int32_t B, int32_t L,
                           void* stream);
/* ... and in front of tae_enc_backward_bf16: x_tx = ELU(Linear(h)) (encoders.py:364-371), so
 *   dlin[branch, b, l] = d_x_tx[b, l, branch] * (x_tx[b, l, branch] > 0 ? 1 : x_tx[b, l, branch] + 1);  dlin is (3, B, L, 1).   */
int tae_enc_out_backward_f32(const float* d_x_tx, const float* x_tx, float* dlin, int32_t B, int32_t L, void* stream);

/* The same three steps for ENC_interCNN (reference encoders.py:362-373 under trainer.py:74): 3 branches = 3 independent stacks
 * with one input channel and Linear(units, 1); x_tx / stats as tae_enc_forward_bf16; dlin (3, B, L, 1) = gradient w.r.t. each
 * branch's Linear output (i.e. d x_tx[:, :, branch] * ELU'), dxin_all (3, B, L, 8) = gradient w.r.t. the +-1 input in column 0. */
int tae_enc_forward_train_bf16(const TaeEncConfig* cfg, const void* packed, const float* u, const int32_t* perm,
                               const int32_t* inv_perm, float* x_tx, double* stats, int32_t B, void* stash_y, void* stash_x,
                               void* workspace, size_t workspace_bytes, void* stream);
size_t tae_enc_bwd_packed_bytes(const TaeEncConfig* cfg);
int    tae_enc_pack_bwd_bf16(const TaeEncConfig* cfg, const float* params, void* packed_bwd, void* stream);
int tae_enc_backward_bf16(const TaeEncConfig* cfg, const void* packed_bwd, const float* dlin, const void* stash_y, void* stash_g,
                          void* stash_d, float* dxin_all, float* grad_flat, int32_t B, void* workspace, size_t workspace_bytes,
                          void* stream);
/* Weight gradients as tensor-core GEMMs over group images (one CTA per job):
 *   grad[m*s_m + (n0+n)*s_n + t*s_t] += sum_{group in [g0,g1)} sum_rows A[row, m] * B[row + t - taps/2 + tap_shift, b_c0*8 + n]
 * for m < m_valid, n < n_valid, t < taps (row offsets t - taps/2 + tap_shift must stay within [-2, 2]: a job may cover a SUBSET of a
 * 5-tap kernel's taps, e.g. taps = 3, tap_shift = -1 for taps 0..2 and taps = 2, tap_shift = 2, grad advanced by 3*s_t for taps 3..4); bias_grad[m] += sum_rows A[row, m] (NULL to skip; needs 8*b_nc < n_cols).
 * a_img has 13 chunks per group, b_img has b_chunks; the job reads chunks [b_c0, b_c0 + b_nc) of B (b_nc <= 13).
 * n_cols = UMMA N: a multiple of 16, 8*b_nc <= n_cols <= 8*(b_nc+1), taps * n_cols <= 512.
 * Conv layer: A = stash_g layer, B = its input image (stash_y of the layer below, or stash_x), taps = 5, grad = dW
 * (Cout, Cin, 5): s_m = 5*Cin, s_n = 5, s_t = 1.  Linear: A = last stash_y, B = stash_d, taps = 1, grad = dV (F, units):
 * s_m = 1, s_n = units.  `jobs_host` is a HOST array (validated on every call); `jobs_dev` is NULL or a device copy of it
 * the caller uploaded once (job lists are reused from step to step: no per-call copy, no host synchronisation);
 * workspace (device) >= 256 bytes, + n_jobs * sizeof(TaeWgradJob) when jobs_dev is NULL.                             */
typedef struct TaeWgradJob {
  const void* a_img;
  const void* b_img;
  float* grad;
  float* bias_grad;
  int32_t b_chunks, b_c0, b_nc;
  int32_t taps, n_cols;
  int32_t m_valid, n_valid, n0;
  int32_t s_m, s_n, s_t;
  int32_t g0, g1;
  int32_t tap_shift;
} TaeWgradJob;
int tae_wgrad_bf16(const TaeWgradJob* jobs_host, int32_t n_jobs, const void* jobs_dev, void* workspace, size_t workspace_bytes,
                   void* stream);

/* ---- next row f2: DEC_LargeRNN (reference decoders.py:16-149, torch.nn.GRU 2 layers bidirectional) -----------------
 * One direction of one GRU layer over a whole batch: xproj (B, L, 3H) = W_ih x + b_ih for every time step (gate order r,
 * z, n; computed with tae_conv1d_elu_f32, K = 1), w_hh (3H, H), b_hh (3H); writes h_t into out[b, t, out_offset .. +H) of
 * a (B, L, out_stride) tensor (out_offset = 0 forward, H reverse).  reverse != 0 runs t = L-1 .. 0.               */
int tae_gru_direction_f32(const float* xproj, const float* w_hh, const float* b_hh, float* out,
                          int32_t B, int32_t L, int32_t H, int32_t out_stride, int32_t out_offset, int32_t reverse, void* stream);
/* Backward of the same recurrence (training, reference trainer.py:74 through torch.nn.GRU): hout / dout are (B, L, io_stride)
 * tensors holding this direction's hidden states / their gradient at [io_offset, io_offset + H).  Writes dgi (B, L, 3H), the
 * gradient at the input-side pre-activations W_ih x + b_ih (gate order r, z, n), and dghn (B, L, H), the n part of the
 * gradient at the hidden-side pre-activations W_hh h + b_hh (its r and z parts equal dgi's).  The parameter and input
 * gradients are then plain GEMMs / sums over (B, L): dW_ih = dgi^T x, dx = dgi W_ih, dW_hh = [dgi_r, dgi_z, dghn]^T h_prev. */
int tae_gru_direction_bwd_f32(const float* xproj, const float* w_hh, const float* b_hh, const float* hout, const float* dout,
                              float* dgi, float* dghn, int32_t B, int32_t L, int32_t H, int32_t io_stride, int32_t io_offset,
                              int32_t reverse, void* stream);

/* The same recurrence on the tensor cores (bf16 operands, fp32 accumulation, fp32 hidden state): the input projection is
 * part of the per-step MMA chain, so no (B, L, 3H) projection tensor exists.  Activations travel between the launches as
 * TIME-MAJOR TILES: bf16 [block of R codewords][t][chunk of 8 channels][R][8] -- the operand chunks of one time step are
 * contiguous (bulk copies in, coalesced stores out).  R = rows per block: 16, 32, 64, 96 or 128; tae_gru_rows_per_block(B) picks 32 .. 128 (small
 * batches use fewer rows per CTA so that every SM gets a block); a tile buffer holds an even number of blocks
 * (tae_gru_tile_bytes).  Channels are stored in groups of grp_valid real channels padded to a multiple of 8: the stack
 * input is one group (2 + F channels in 8), a layer's output is two groups (forward | reverse, H channels each in 13
 * chunks), which is the next layer's input (in_ch = 2H, grp_valid = H).
 *   tae_gru_tiles_from_f32 : (B, L, C) fp32 -> tiles with ceil(C/8) chunks
 *   tae_gru_pack_bf16      : weight_ih (3H, in_ch), weight_hh (3H, H), bias_ih, bias_hh of one layer-direction -> `packed`
 *                            (rebuild after a weight update)
 *   tae_gru_direction_bf16 : x_tiles -> h_t for all t into chunks [out_chunk0, out_chunk0 + 13) of out_tiles (out_chunks
 *                            chunks per time step: 26 for a bidirectional layer); reverse != 0 runs t = L-1 .. 0
 *   tae_gru_linear_f32     : the dec*_outputs Linear (reference decoders.py:104-105, 118-119) straight from tiles:
 *                            out (B, L, F) fp32 = weight (F, in_ch) . h + bias, F <= 8
 * H: multiple of 4, <= 104; at most 26 input chunks; workspace >= 256 bytes.                                          */
int32_t tae_gru_rows_per_block(int32_t B);
size_t tae_gru_tile_bytes(int32_t B, int32_t L, int32_t n_chunks, int32_t R);
size_t tae_gru_packed_bytes(int32_t H, int32_t in_ch, int32_t grp_valid);
int    tae_gru_pack_bf16(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, void* packed,
                         int32_t H, int32_t in_ch, int32_t grp_valid, void* stream);
int tae_gru_direction_bf16(const void* packed, const void* x_tiles, void* out_tiles, int32_t B, int32_t L, int32_t H,
                           int32_t in_ch, int32_t grp_valid, int32_t R, int32_t out_chunks, int32_t out_chunk0,
                           int32_t reverse, void* workspace, size_t workspace_bytes, void* stream);
int tae_gru_tiles_from_f32(const float* x, void* tiles, int32_t B, int32_t L, int32_t C, int32_t R, void* stream);
int tae_gru_linear_f32(const void* tiles, const float* weight, const float* bias, float* out, int32_t B, int32_t L,
                       int32_t in_ch, int32_t grp_valid, int32_t F, int32_t R, void* stream);

/* ---- around the path: on-device channel and metrics -------------------------------------
 * AWGN channel, reference channel_ae.py:41-42 with channels.py:21-35: received = codes + sigma * N(0,1).
 * The reference draws torch.randn on the CPU (unseeded); this stream is Philox4x32-10 + Box-Muller, element i uses
 * counter (offset + i/4), word i%4, key = seed -- reproducible, restated in oracle/turboae_oracle.py.  codes may alias
 * received.                                                                                                      */
int tae_awgn_f32(const float* codes, float* received, size_t n, float sigma, uint64_t seed, uint64_t offset, void* stream);
/* reference utils.py:6-18 (errors_ber) and :49-66 (errors_bler): adds sum(round(y_true) != round(y_pred)) to counts[0]
 * and the number of codewords with at least one such error to counts[1] (device uint64[2], caller zeroes).        */
int tae_error_count_f32(const float* y_true, const float* y_pred, int32_t B, int32_t L, unsigned long long* counts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TURBOAE_B200_H_ */
