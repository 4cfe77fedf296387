// On-device channel and metrics around the hot path:
//   * AWGN:  received = codes + sigma * N(0,1)        (reference channels.py:21-35 `generate_noise`, channel_ae.py:41-42)
//     The reference draws torch.randn on the CPU, unseeded; here the stream is Philox4x32-10 (counter = element index / 4,
//     key = seed) + Box-Muller, reproducible and restated bit-for-bit (integers) in oracle/turboae_oracle.py.
//   * error counting: bit errors  sum(round(y_true) != round(y_pred))          (reference utils.py:6-18  errors_ber)
//                     block errors = codewords with at least one bit error      (reference utils.py:49-66 errors_bler)
//     torch.round is round-half-to-even = rintf.
#include "tae_common.cuh"

namespace tae {

namespace {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// two uniforms in (0,1) -> two standard normals
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
  const float u1 = ((float)a + 0.5f) * 2.3283064365386963e-10f;       // (a + 0.5) / 2^32, never 0
  const float u2 = ((float)b + 0.5f) * 2.3283064365386963e-10f;
  const float r = sqrtf(-2.0f * logf(u1));
  float s, c;
  sincospif(2.0f * u2, &s, &c);
  z0 = r * c;
  z1 = r * s;
}

__global__ void awgn_kernel(const float* __restrict__ codes, float* __restrict__ received, size_t n, float sigma, uint32_t seed_lo,
                            uint32_t seed_hi, uint64_t offset) {
  const size_t n4 = (n + 3) / 4;
  for (size_t q = blockIdx.x * (size_t)blockDim.x + threadIdx.x; q < n4; q += (size_t)gridDim.x * blockDim.x) {
    const uint64_t ctr = offset + q;
    uint32_t x[4];
    philox4x32_10((uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u, seed_lo, seed_hi, x);
    float z[4];
    box_muller(x[0], x[1], z[0], z[1]);
    box_muller(x[2], x[3], z[2], z[3]);
    const size_t i0 = q * 4;
    if (i0 + 3 < n && (((uintptr_t)(codes + i0) | (uintptr_t)(received + i0)) & 15) == 0) {
      const float4 c = *reinterpret_cast<const float4*>(codes + i0);
      *reinterpret_cast<float4*>(received + i0) = make_float4(fmaf(sigma, z[0], c.x), fmaf(sigma, z[1], c.y), fmaf(sigma, z[2], c.z),
                                                               fmaf(sigma, z[3], c.w));
    } else {
      for (int j = 0; j < 4 && i0 + j < n; ++j) received[i0 + j] = fmaf(sigma, z[j], codes[i0 + j]);
    }
  }
}

// one warp per codeword
__global__ void error_count_kernel(const float* __restrict__ y_true, const float* __restrict__ y_pred, int B, int L,
                                   unsigned long long* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const size_t warp0 = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5, nwarp = ((size_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long bits = 0, blocks = 0;
  for (size_t b = warp0; b < (size_t)B; b += nwarp) {
    int e = 0;
    for (int l = lane; l < L; l += 32) e += (rintf(y_true[b * L + l]) != rintf(y_pred[b * L + l])) ? 1 : 0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) e += __shfl_xor_sync(0xffffffffu, e, d);
    bits += (unsigned long long)e;
    blocks += e > 0 ? 1ull : 0ull;
  }
  if (lane == 0 && (bits | blocks)) {
    atomicAdd(counts + 0, bits);
    atomicAdd(counts + 1, blocks);
  }
}

}  // namespace

int launch_awgn(const float* codes, float* received, size_t n, float sigma, uint64_t seed, uint64_t offset, cudaStream_t s) {
  if (n == 0) return TAE_OK;
  const size_t n4 = (n + 3) / 4;
  const int blocks = (int)std::min<size_t>((n4 + 255) / 256, 148 * 16);
  awgn_kernel<<<blocks, 256, 0, s>>>(codes, received, n, sigma, (uint32_t)seed, (uint32_t)(seed >> 32), offset);
  return after_launch("awgn_kernel");
}

int launch_error_count(const float* y_true, const float* y_pred, int B, int L, unsigned long long* counts, cudaStream_t s) {
  if (B == 0) return TAE_OK;
  const int blocks = (int)std::min<size_t>(((size_t)B * 32 + 255) / 256, 148 * 16);
  error_count_kernel<<<blocks, 256, 0, s>>>(y_true, y_pred, B, L, counts);
  return after_launch("error_count_kernel");
}

}  // namespace tae
