// Weight gradients of the conv stacks on 5th-gen tensor cores (training path).
//
// Reference arithmetic restated: trainer.py:74 (loss.backward()) through cnn_utils.py:36-46, i.e. for one Conv1d layer
//     dW[o, c, t] = sum_{b, l} g[b, l, o] * x[b, l + t - K/2, c]          db[o] = sum_{b, l} g[b, l, o]
// with g = dL/dz (gradient at the pre-activation) and x the layer input, and for the Linear after the stack
//     dV[f, o] = sum_{b, l} dlin[b, l, f] * h[b, l, o].
//
// Both operands are "group images": the bf16 activation layout the fused forward / backward kernels keep in shared
// memory and stash in HBM, [group][chunk = C/8][516 rows][8 channels] (rows = 2 halo rows, then per codeword L positions
// + 2 all-zero separator rows).  The reduction runs over ROWS, so both operands are MN-major for tcgen05.mma (8 channels
// contiguous, consecutive rows 16 bytes apart, 8-row blocks 128 bytes apart = LBO, 8-channel blocks one chunk apart = SBO):
// the images are consumed exactly as they were written, no transpose.  Tap t of the convolution is the B operand's start
// address moved by 16*t bytes; the zero separator rows make the shifted products vanish across codeword borders.
//     D_t[m = o][n = c]  +=  A[rows, o]^T  B[rows + t - 2, c]        (M = 128, N = 16..64, K = 16 rows per MMA)
// One CTA = one job (a layer, a slab of input channels or -- for layers wider than 64 channels -- a SUBSET OF THE TAPS over all
// input channels, a range of groups): accumulators of the job's taps stay in TMEM for the whole range (5 x 64 or 3 x 112
// columns), then go out as fp32 atomics.  The MMAs are bound by operand fetch through the 128 B/clk shared-memory port: an
// M128 x N64 MMA reads 6 KB for 32 tensor cycles (port: 48), an M128 x N112 one 7.5 KB for 56 (port: 59), which is why a
// 100-channel layer runs as taps {0,1,2} and {3,4} at N = 112 rather than as channel slabs 64 + 36 at all five taps.  Half groups stream through a 2-stage bulk-copy
// pipeline (the copies of half i+1 overlap the MMAs of half i).  A constant all-ones chunk appended to B yields db in
// a spare column.
#include <cuda_bf16.h>

#include <cstdio>

#include "tae_common.cuh"
#include "tae_umma.cuh"

namespace tae {

namespace {

constexpr uint32_t W_ROWS = 516, W_CHUNK_B = W_ROWS * 16, W_A_CHUNKS = 13, W_B_CHUNKS_MAX = 13;
// One pipeline stage holds HALF (or a quarter) of a group: a 260-row window (256 reduction rows + 2 halo rows each side for the taps) of the
// 13 A chunks and of up to 13 B chunks, followed by the constant ones chunk; chunks are W_WIN_B apart (= SBO).  The three
// A chunks that M = 128 reads past channel 103 fall into the stage's own B region (their output rows are never stored).
#ifndef TAE_WGRAD_STAGES
#define TAE_WGRAD_STAGES 2
#endif
// 2 stages of half a group (shipped) or 4 stages of a quarter group (-DTAE_WGRAD_STAGES=4: three windows = 168 KB in flight instead
// of one = 112 KB; measured SLOWER, 3.39 vs 3.16 ms per training step: the stage loads are not what limits the kernel, and twice as
// many hand-overs cost more than the deeper pipeline buys).
constexpr int W_STAGES = TAE_WGRAD_STAGES;                                      // 2 or 4 windows per group, one stage each
constexpr uint32_t W_PART_ROWS = 512 / W_STAGES;                                // reduction rows of one window
constexpr uint32_t W_WIN_ROWS = W_PART_ROWS + 4, W_WIN_B = W_WIN_ROWS * 16;     // + 2 halo rows each side: 260 rows = 4160 B (132 rows = 2112 B)
constexpr uint32_t W_STAGE_B = (W_A_CHUNKS + W_B_CHUNKS_MAX + 1) * W_WIN_B;     // 112 320 (57 024)
constexpr uint32_t W_BAR_OFF = W_STAGES * W_STAGE_B;
constexpr uint32_t W_SMEM = W_BAR_OFF + 128;                                    // full[4] empty[4] final tmem-ptr
constexpr int W_KSTEPS = W_PART_ROWS / 16;
static_assert(W_STAGES == 2 || W_STAGES == 4, "2 or 4 stages");
static_assert(W_SMEM <= 232448, "shared memory of a CTA");

__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
         (1ull << 46);
}

__global__ void __launch_bounds__(128, 1) wgrad_kernel(const TaeWgradJob* __restrict__ jobs, int* err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const TaeWgradJob J = jobs[blockIdx.x];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto bar_full = [&](int s) { return sbase + W_BAR_OFF + 8u * (uint32_t)s; };
  auto bar_empty = [&](int s) { return sbase + W_BAR_OFF + 32u + 8u * (uint32_t)s; };
  const uint32_t bar_final = sbase + W_BAR_OFF + 64, tptr = sbase + W_BAR_OFF + 72;
  const int n_it = W_STAGES * (J.g1 - J.g0);            // windows (half or quarter groups)

  // B regions of both stages: zeros, then the ones chunk right behind the slab (channel 0 of every row = 1.0)
  for (int st = 0; st < W_STAGES; ++st) {
    const uint32_t b_off = sbase + (uint32_t)st * W_STAGE_B + W_A_CHUNKS * W_WIN_B;
    for (uint32_t i = threadIdx.x * 16; i < (W_B_CHUNKS_MAX + 1) * W_WIN_B; i += blockDim.x * 16) st_shared_v4(b_off + i, 0u, 0u, 0u, 0u);
  }
  __syncthreads();
  for (int st = 0; st < W_STAGES; ++st)
    for (uint32_t r = threadIdx.x; r < W_WIN_ROWS; r += blockDim.x)
      st_shared_v4(sbase + (uint32_t)st * W_STAGE_B + (W_A_CHUNKS + (uint32_t)J.b_nc) * W_WIN_B + r * 16, 0x00003F80u, 0u, 0u, 0u);
  if (threadIdx.x == 0) {
    for (int st = 0; st < W_STAGES; ++st) { mbar_init(bar_full(st), 1); mbar_init(bar_empty(st), 1); }
    mbar_init(bar_final, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<1>(tptr, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tptr) : "memory");

  if (warp == 0) {
    if (lane == 0) {
      // producer: half h of group g = rows [256 h, 256 h + 260) of every chunk
      const uint32_t bytes = (W_A_CHUNKS + (uint32_t)J.b_nc) * W_WIN_B;
      for (int it = 0; it < n_it; ++it) {
        const int st = it % W_STAGES;
        const size_t g = (size_t)(J.g0 + it / W_STAGES);
        const uint32_t row_off = (uint32_t)st * W_PART_ROWS * 16u;
        if (it >= W_STAGES) mbar_wait(bar_empty(st), (uint32_t)(it / W_STAGES - 1) & 1u, err, 21);    // the MMAs that read this stage are done
        mbar_arrive_expect_tx(bar_full(st), bytes);
        const uint32_t dst = sbase + (uint32_t)st * W_STAGE_B;
        const uint8_t* a = reinterpret_cast<const uint8_t*>(J.a_img) + g * W_A_CHUNKS * W_CHUNK_B + row_off;
        for (uint32_t c = 0; c < W_A_CHUNKS; ++c) bulk_g2s(dst + c * W_WIN_B, a + (size_t)c * W_CHUNK_B, W_WIN_B, bar_full(st));
        const uint8_t* b = reinterpret_cast<const uint8_t*>(J.b_img) + (g * (size_t)J.b_chunks + (size_t)J.b_c0) * W_CHUNK_B + row_off;
        for (uint32_t c = 0; c < (uint32_t)J.b_nc; ++c)
          bulk_g2s(dst + (W_A_CHUNKS + c) * W_WIN_B, b + (size_t)c * W_CHUNK_B, W_WIN_B, bar_full(st));
      }
    }
  } else if (warp == 1) {
    // MMA issuer.  Both operands MN-major: LBO = 128 B between 8-row blocks of the reduction, SBO = one chunk window.
    const uint32_t idesc = make_idesc(128, J.n_cols) | (1u << 15) | (1u << 16);
    const int half = J.taps / 2;
#ifdef TAE_WGRAD_PROBE
    long long t_wait = 0, t_first = 0;
    const long long t_begin = clock64();
#endif
    for (int it = 0; it < n_it; ++it) {
      const int st = it % W_STAGES;
#ifdef TAE_WGRAD_PROBE
      const long long tw = clock64();
#endif
      mbar_wait(bar_full(st), (uint32_t)(it / W_STAGES) & 1u, err, 22);
#ifdef TAE_WGRAD_PROBE
      if (it == 0) t_first = clock64() - tw; else t_wait += clock64() - tw;
#endif
      tc_fence_after();
      if (elect_one()) {
        const uint32_t stage = sbase + (uint32_t)st * W_STAGE_B;
        const uint64_t a0 = mn_desc(stage + 2 * 16, 128u, W_WIN_B);
        for (int t = 0; t < J.taps; ++t) {
          const uint64_t b0 = mn_desc(stage + W_A_CHUNKS * W_WIN_B + (uint32_t)(2 + t - half + J.tap_shift) * 16, 128u, W_WIN_B);
          const uint32_t d = tmem_base + (uint32_t)(t * J.n_cols);
#pragma unroll
          for (int ks = 0; ks < W_KSTEPS; ++ks)
            umma_bf16<1>(d, a0 + (uint64_t)(ks * 16), b0 + (uint64_t)(ks * 16), idesc, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit_1(bar_empty(st));
        if (it == n_it - 1) umma_commit_1(bar_final);      // its own barrier: the drain threads do not follow the per-stage phases
      }
      __syncwarp();
    }
#ifdef TAE_WGRAD_PROBE
    // (issue-side view: the last window's MMAs are still running when the loop ends)
    if (lane == 0 && (blockIdx.x % 37) == 0)
      printf("wgrad job %4d taps %d N %3d windows %3d: first load %6lld  waits for loads %8lld  loop %8lld cycles (%.0f per MMA issued)\n", (int)blockIdx.x, J.taps,
             J.n_cols, n_it, t_first, t_wait, clock64() - t_begin, (double)(clock64() - t_begin) / ((double)n_it * J.taps * W_KSTEPS));
#endif
  }
  __syncwarp();
  // ---- drain: every thread owns TMEM lane m = output row -------------------------------------------------------------
  if (n_it > 0) {
    mbar_wait(bar_final, 0u, err, 23);
    tc_fence_after();
#ifdef TAE_WGRAD_PROBE
    const long long t_drain = clock64();
#endif
    const int m = 32 * warp + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)(32 * warp) << 16);
    const uint32_t s_row = (uint32_t)(J.taps * J.n_cols) | 1u;
    if (J.s_m == 1 || 512u * s_row > W_BAR_OFF) {
      // Linear (dV is (F, units): consecutive output rows m are consecutive floats): straight from the registers, lanes across m
      // (also the fallback for a job whose transposed tile would not fit the stage memory)
      for (int t = 0; t < J.taps; ++t)
        for (int cb = 0; cb < J.n_cols / 16; ++cb) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(t * J.n_cols + cb * 16), r);
          tmem_ld_wait();
          if (m < J.m_valid) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = cb * 16 + j;
              if (n < J.n_valid) atomicAdd(J.grad + (size_t)m * J.s_m + (size_t)(J.n0 + n) * J.s_n + (size_t)t * J.s_t, __uint_as_float(r[j]));
              else if (J.bias_grad && t == 0 && n == 8 * J.b_nc) atomicAdd(J.bias_grad + m, __uint_as_float(r[j]));
            }
          }
        }
    } else {
      // Conv (dW is (Cout, Cin, taps): one output row m owns Cin * taps CONTIGUOUS floats, this job a (tap subset x channel range) of
      // them).  A thread holds one row in TMEM, so atomics issued from the registers hit 32 different rows per warp instruction
      // (32 sectors; measured: the drain took as long as the MMA loop).  Transpose through the now idle stage memory instead:
      // S[m][t * n_cols + n] (row stride odd: conflict-free both ways), then each warp walks whole rows with its lanes across
      // (n, t), t fastest -- runs of `taps` adjacent floats every s_n: ~7 sectors per warp instruction.
      for (int t = 0; t < J.taps; ++t)
        for (int cb = 0; cb < J.n_cols / 16; ++cb) {
          uint32_t r[16];
          tmem_ld16(lane_addr + (uint32_t)(t * J.n_cols + cb * 16), r);
          tmem_ld_wait();
          const uint32_t dst = sbase + 4u * ((uint32_t)m * s_row + (uint32_t)(t * J.n_cols + cb * 16));
#pragma unroll
          for (int j = 0; j < 16; ++j) st_shared_f32(dst + 4u * (uint32_t)j, __uint_as_float(r[j]));
        }
      __syncthreads();
      const int items = J.n_valid * J.taps;
      for (int mm = warp; mm < J.m_valid; mm += 4) {
        const uint32_t src = sbase + 4u * (uint32_t)mm * s_row;
        float* g_row = J.grad + (size_t)mm * J.s_m + (size_t)J.n0 * J.s_n;
        for (int idx = lane; idx < items; idx += 32) {
          const int n = idx / J.taps, t = idx - n * J.taps;
          atomicAdd(g_row + (size_t)n * J.s_n + (size_t)t * J.s_t, ld_shared_f32(src + 4u * (uint32_t)(t * J.n_cols + n)));
        }
        if (J.bias_grad && lane == 0) atomicAdd(J.bias_grad + mm, ld_shared_f32(src + 4u * (uint32_t)(8 * J.b_nc)));
      }
    }
#ifdef TAE_WGRAD_PROBE
    __syncthreads();
    if (threadIdx.x == 0 && (blockIdx.x % 37) == 0)
      printf("wgrad job %4d taps %d N %3d: drain %8lld cycles\n", (int)blockIdx.x, J.taps, J.n_cols, clock64() - t_drain);
#endif
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tmem_base, 512); }
}

}  // namespace

int launch_wgrad(const TaeWgradJob* jobs_host, int n_jobs, const void* jobs_dev, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (n_jobs == 0) return TAE_OK;
  const size_t need = 256 + (jobs_dev ? 0 : sizeof(TaeWgradJob) * (size_t)n_jobs);
  if (ws_bytes < need) { set_error("tae_wgrad_bf16: workspace %zu < %zu bytes", ws_bytes, need); return TAE_EWORKSPACE; }
  for (int i = 0; i < n_jobs; ++i) {
    const TaeWgradJob& J = jobs_host[i];
    if (!J.a_img || !J.b_img || !J.grad || J.b_nc < 1 || J.b_nc > (int)W_B_CHUNKS_MAX || J.b_c0 < 0 || J.b_c0 + J.b_nc > J.b_chunks ||
        J.taps < 1 || J.taps > 5 || 2 - J.taps / 2 + J.tap_shift < 0 || 2 - J.taps / 2 + J.tap_shift + J.taps - 1 > 4 || J.n_cols % 16 || J.n_cols < 16 || J.n_cols > 8 * (J.b_nc + 1) || J.n_cols < 8 * J.b_nc || J.taps * J.n_cols > 512 ||
        J.m_valid < 1 || J.m_valid > 104 || J.n_valid < 0 || J.n_valid > 8 * J.b_nc || J.g1 < J.g0) {
      set_error("tae_wgrad_bf16: job %d is malformed", i);
      return TAE_EINVAL;
    }
  }
  static DeviceOnce once;
  {
    int rc = device_once(once, "wgrad_kernel", [](int dev) -> int {
      int rc2 = require_sm100(dev, "the tensor-core weight-gradient kernel");
      if (rc2) return rc2;
      cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W_SMEM);
      if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(wgrad_kernel): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
      return TAE_OK;
    });
    if (rc) return rc;
  }
  int* err = reinterpret_cast<int*>(align_up(reinterpret_cast<uintptr_t>(ws), 16));      // the job upload below lives in the same workspace
  int* wait_code = wait_code_slot(ws);
  const TaeWgradJob* d_jobs = reinterpret_cast<const TaeWgradJob*>(jobs_dev);
  if (!d_jobs) {      // no resident copy: upload (a pageable source makes this call wait for the copy)
    TaeWgradJob* up = reinterpret_cast<TaeWgradJob*>(reinterpret_cast<uint8_t*>(err) + 128);
    cudaError_t e = cudaMemcpyAsync(up, jobs_host, sizeof(TaeWgradJob) * (size_t)n_jobs, cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) { set_error("cudaMemcpyAsync(jobs): %s", cudaGetErrorString(e)); return TAE_ECUDA; }
    d_jobs = up;
  }
  wgrad_kernel<<<n_jobs, 128, W_SMEM, s>>>(d_jobs, wait_code);
  return after_launch("wgrad_kernel");
}

}  // namespace tae
