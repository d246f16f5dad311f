// Shapes of the device MT19937 stream kernel (resynthesizer_b200/csrc/rs_kernels.cu: k_mt19937_raw), timed with CUDA events
// on one B200 and checked word for word against a host MT19937.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mt_bench mt_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t lds(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t temper(uint32_t y) {
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; return y ^ (y >> 18);
}
__device__ __forceinline__ uint32_t twist(uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
}
template <int N> __device__ __forceinline__ void bar_named() { asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory"); }

// NT maker threads (rounded up to whole warps), WPL words per maker lane per step, WT writer threads (0: makers temper and store)
template <int NT, int WPL, int WT>
__global__ void __launch_bounds__(((NT + 31) / 32) * 32 + WT, 1) mt_kernel(uint32_t seed, uint32_t n_words, uint32_t *__restrict__ out) {
  constexpr int NTP = ((NT + 31) / 32) * 32, ALL = NTP + WT;
  __shared__ __align__(16) uint32_t ring[2048];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t t = threadIdx.x;
  if (t == 0) {
    uint32_t x = seed;
    ring[0] = x;
    for (uint32_t i = 1; i < 624; i++) { x = 1812433253u * (x ^ (x >> 30)) + i; ring[i] = x; }
  }
  __syncthreads();
  const uint32_t rounds = (n_words + 453u) / 454u;
  const uint32_t last = rounds + (WT ? 1u : 0u);
  if (t < (uint32_t)NTP) {
    uint32_t c[WPL];
#pragma unroll
    for (int k = 0; k < WPL; k++) c[k] = (t + k * NT < 227u && t < (uint32_t)NT) ? ring[397u + t + k * NT] : 0u;
    uint32_t pb = t * 4u;  // byte offset in the ring of x[i0 - 624]
    for (uint32_t r = 0; r < last; r++) {
      if (r < rounds && t < (uint32_t)NT) {
        uint32_t a0[WPL], b0[WPL], a1[WPL], b1[WPL];
#pragma unroll
        for (int k = 0; k < WPL; k++)
          if (t + k * NT < 227u) {
            const uint32_t q = pb + k * NT * 4u;
            a0[k] = lds(sb + (q & 8188u)); b0[k] = lds(sb + ((q + 4u) & 8188u));
            a1[k] = lds(sb + ((q + 908u) & 8188u)); b1[k] = lds(sb + ((q + 912u) & 8188u));
          }
#pragma unroll
        for (int k = 0; k < WPL; k++)
          if (t + k * NT < 227u) {
            const uint32_t q = pb + k * NT * 4u;
            const uint32_t v0 = twist(a0[k], b0[k], c[k]), v1 = twist(a1[k], b1[k], v0);
            c[k] = v1;
            sts(sb + ((q + 2496u) & 8188u), v0);
            sts(sb + ((q + 3404u) & 8188u), v1);
            if (WT == 0) {
              const uint32_t i0 = r * 454u + t + k * NT;
              if (i0 < n_words) out[i0] = temper(v0);
              if (i0 + 227u < n_words) out[i0 + 227u] = temper(v1);
            }
          }
        pb = (pb + 1816u) & 8188u;
      }
      if (WT) bar_named<ALL>(); else __syncthreads();
    }
  } else {
    const uint32_t u = t - NTP;
    for (uint32_t r = 0; r < last; r++) {
      if (r > 0) {
        const uint32_t base = (r - 1u) * 454u;
#pragma unroll
        for (int k = 0; k < (454 + (WT ? WT : 1) - 1) / (WT ? WT : 1); k++) {
          const uint32_t o = u + k * WT, i = base + o;
          if (o < 454u && i < n_words) out[i] = temper(lds(sb + (((i + 624u) & 2047u) << 2)));
        }
      }
      bar_named<ALL>();
    }
  }
}


// Rounds of 623 words: x[i] = x[i-227] ^ f(x[i-624], x[i-623]) -- every f of a 623-word round reads words older than the
// round, and the x[i-227] chain is one XOR per word, kept inside a lane: lane t makes words t, t + 227 and (t < 169) t + 454.
template <int WT>
__global__ void __launch_bounds__(256 + WT, 1) mt3_kernel(uint32_t seed, uint32_t n_words, uint32_t *__restrict__ out) {
  constexpr int ALL = 256 + WT;
  __shared__ __align__(16) uint32_t ring[2048];
  const uint32_t sb = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t t = threadIdx.x;
  if (t == 0) {
    uint32_t x = seed;
    ring[0] = x;
    for (uint32_t i = 1; i < 624; i++) { x = 1812433253u * (x ^ (x >> 30)) + i; ring[i] = x; }
  }
  __syncthreads();
  const uint32_t rounds = (n_words + 622u) / 623u;
  const uint32_t last = rounds + (WT ? 1u : 0u);
  auto f = [](uint32_t a, uint32_t b) {
    const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
    return (y >> 1) ^ ((0u - (y & 1u)) & 0x9908b0dfu);
  };
  if (t < 256u) {
    uint32_t pb = t * 4u;  // byte offset in the ring of x[i0 - 624], i0 = 623 r + t
    const bool mk = t < 227u, third = t < 169u;
    for (uint32_t r = 0; r < last; r++) {
      if (r < rounds && mk) {
        const uint32_t c = lds(sb + ((pb + 1588u) & 8188u));  // x[i0 - 227]
        const uint32_t a0 = lds(sb + pb), b0 = lds(sb + ((pb + 4u) & 8188u));
        const uint32_t a1 = lds(sb + ((pb + 908u) & 8188u)), b1 = lds(sb + ((pb + 912u) & 8188u));
        uint32_t a2 = 0, b2 = 0;
        if (third) { a2 = lds(sb + ((pb + 1816u) & 8188u)); b2 = lds(sb + ((pb + 1820u) & 8188u)); }
        const uint32_t f0 = f(a0, b0), f1 = f(a1, b1), f2 = f(a2, b2);
        const uint32_t v0 = c ^ f0, v1 = v0 ^ f1, v2 = v1 ^ f2;
        sts(sb + ((pb + 2496u) & 8188u), v0);
        sts(sb + ((pb + 3404u) & 8188u), v1);
        if (third) sts(sb + ((pb + 4312u) & 8188u), v2);
        if (WT == 0) {
          const uint32_t i0 = r * 623u + t;
          if (i0 < n_words) out[i0] = temper(v0);
          if (i0 + 227u < n_words) out[i0 + 227u] = temper(v1);
          if (third && i0 + 454u < n_words) out[i0 + 454u] = temper(v2);
        }
        pb = (pb + 2492u) & 8188u;
      }
      if (WT) bar_named<ALL>(); else __syncthreads();
    }
  } else {
    const uint32_t u = t - 256u;
    constexpr int K = (623 + (WT ? WT : 1) - 1) / (WT ? WT : 1);
    for (uint32_t r = 0; r < last; r++) {
      if (r > 0) {
        const uint32_t base = (r - 1u) * 623u;
        uint32_t y[K];
#pragma unroll
        for (int k = 0; k < K; k++) { const uint32_t o = u + k * WT; y[k] = o < 623u ? lds(sb + (((base + o + 624u) & 2047u) << 2)) : 0u; }
#pragma unroll
        for (int k = 0; k < K; k++) { const uint32_t o = u + k * WT, i = base + o; if (o < 623u && i < n_words) out[i] = temper(y[k]); }
      }
      bar_named<ALL>();
    }
  }
}
template <int WT>
void run3(const char *name, uint32_t n, uint32_t *d, const std::vector<uint32_t> &ref, std::vector<uint32_t> &got) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemset(d, 0, (size_t)n * 4);
  float best = 1e9f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    mt3_kernel<WT><<<1, 256 + WT>>>(1198472u, n, d);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaMemcpy(got.data(), d, (size_t)n * 4, cudaMemcpyDeviceToHost);
  const bool ok = memcmp(got.data(), ref.data(), (size_t)n * 4) == 0;
  printf("%-34s threads %4d  %8.3f ms for %u words  (%.1f ns per 623-word round)  %s %s\n", name, 256 + WT, best, n, best * 1e6 / ((n + 622) / 623),
         ok ? "EQUAL" : "DIFFERENT", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

struct HostMT { uint32_t mt[624]; int mti;
  explicit HostMT(uint32_t s) { mt[0] = s; for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i-1] ^ (mt[i-1] >> 30)) + i; mti = 624; }
  uint32_t next() { if (mti >= 624) { for (int k = 0; k < 624; k++) { uint32_t y = (mt[k] & 0x80000000u) | (mt[(k+1)%624] & 0x7fffffffu); mt[k] = mt[(k+397)%624] ^ (y >> 1) ^ ((y & 1) ? 0x9908b0dfu : 0); } mti = 0; }
    uint32_t y = mt[mti++]; y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18; return y; } };

template <int NT, int WPL, int WT>
void run(const char *name, uint32_t n, uint32_t *d, const std::vector<uint32_t> &ref, std::vector<uint32_t> &got) {
  constexpr int TH = ((NT + 31) / 32) * 32 + WT;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaMemset(d, 0, (size_t)n * 4);
  float best = 1e9f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    mt_kernel<NT, WPL, WT><<<1, TH>>>(1198472u, n, d);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  cudaError_t e = cudaGetLastError();
  cudaMemcpy(got.data(), d, (size_t)n * 4, cudaMemcpyDeviceToHost);
  const bool ok = memcmp(got.data(), ref.data(), (size_t)n * 4) == 0;
  printf("%-34s threads %4d  %8.3f ms for %u words  (%.1f ns per 454-word round)  %s %s\n", name, TH, best, n, best * 1e6 / ((n + 453) / 454),
         ok ? "EQUAL" : "DIFFERENT", e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  const uint32_t n = 4390912u;  // cfg3: n + n/32 + 65536 for n = 4 Mi target points
  std::vector<uint32_t> ref(n), got(n);
  HostMT h(1198472u); for (uint32_t i = 0; i < n; i++) ref[i] = h.next();
  uint32_t *d; cudaMalloc(&d, (size_t)n * 4);
  run<227, 1, 0>("227x1 makers store", n, d, ref, got);
  run<128, 2, 0>("128x2 makers store", n, d, ref, got);
  run<114, 2, 0>("114x2 makers store", n, d, ref, got);
  run<64, 4, 0>("64x4 makers store", n, d, ref, got);
  run<32, 8, 0>("32x8 makers store", n, d, ref, got);
  run<227, 1, 256>("227x1 + 256 writers", n, d, ref, got);
  run<227, 1, 128>("227x1 + 128 writers", n, d, ref, got);
  run<128, 2, 128>("128x2 + 128 writers", n, d, ref, got);
  run<128, 2, 256>("128x2 + 256 writers", n, d, ref, got);
  run<64, 4, 64>("64x4 + 64 writers", n, d, ref, got);
  run<64, 4, 128>("64x4 + 128 writers", n, d, ref, got);
  run<32, 8, 32>("32x8 + 32 writers", n, d, ref, got);
  run<32, 8, 96>("32x8 + 96 writers", n, d, ref, got);
  run3<0>("623-word rounds, makers store", n, d, ref, got);
  run3<128>("623-word rounds + 128 writers", n, d, ref, got);
  run3<224>("623-word rounds + 224 writers", n, d, ref, got);
  run3<320>("623-word rounds + 320 writers", n, d, ref, got);
  run3<640>("623-word rounds + 640 writers", n, d, ref, got);
  for (uint32_t m : {1u, 622u, 623u, 624u, 1246u, 100000u}) {  // lengths around the round size
    cudaMemset(d, 0xEE, (size_t)(m + 8) * 4);
    mt3_kernel<320><<<1, 576>>>(1198472u, m, d);
    cudaMemcpy(got.data(), d, (size_t)(m + 8) * 4, cudaMemcpyDeviceToHost);
    bool ok = memcmp(got.data(), ref.data(), (size_t)m * 4) == 0;
    for (int k = 0; k < 8; k++) ok = ok && got[m + k] == 0xEEEEEEEEu;
    printf("n = %u: %s\n", m, ok ? "EQUAL" : "DIFFERENT");
  }
  return 0;
}
