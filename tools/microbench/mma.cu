// Legacy warp-level tensor-core (mma.sync -> SASS HMMA) issue rates on B200 (sm_100a), for the design of the
// reverse-pass kernel: its three 32x32x32 products per warp-pass are candidates for the tensor pipe with an
// fp32-accurate operand split (3 x TF32 or 3 x BF16).  Measures, per SM and per clock:
//   tf32_m16n8k8      mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32   (1024 MAC per instruction)
//   bf16_m16n8k16     mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32  (2048 MAC per instruction)
//   f16_m16n8k16      ... f16.f16.f32
//   *_plus_ffma2      the same with one independent FFMA2 per MMA in the loop (do the two pipes overlap?)
//   cvt_split         the operand split itself: cvt.rna.tf32 + sub + cvt.rna.tf32 per element
// Independent accumulator tiles per warp (NACC) hide the MMA latency; warps per SM are swept.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma mma.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 2048;
constexpr int NACC = 8;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// KIND: 0 tf32, 1 bf16, 2 f16, 3 tf32 + ffma2, 4 bf16 + ffma2, 5 cvt split, 6 tf32 + 2 ffma2
template <int KIND>
__global__ void bench(float* out, long long* cyc, uint32_t seed) {
  float c[NACC][4];
  uint32_t a[4], b[2];
#pragma unroll
  for (int i = 0; i < 4; ++i) a[i] = seed * (i + 3) + threadIdx.x;
  b[0] = seed ^ 0x3f800000u; b[1] = seed + 7u;
#pragma unroll
  for (int u = 0; u < NACC; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i) c[u][i] = (float)(u + i);
  unsigned long long p[NACC], pa = 0x3f8000003f800000ull ^ seed, pb = 0x3c0000003c000000ull ^ (seed << 1);
#pragma unroll
  for (int u = 0; u < NACC; ++u) p[u] = pa + u;
  float f[NACC];
#pragma unroll
  for (int u = 0; u < NACC; ++u) f[u] = 1.0f + 1e-3f * (u + threadIdx.x) + __uint_as_float(seed & 0xff);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < NACC; ++u) {
      if constexpr (KIND == 0 || KIND == 3 || KIND == 6) mma_tf32(c[u], a, b);
      if constexpr (KIND == 1 || KIND == 4) mma_bf16(c[u], a, b);
      if constexpr (KIND == 2) mma_f16(c[u], a, b);
      if constexpr (KIND == 3 || KIND == 4 || KIND == 6) fma2(p[u], pa, pb);
      if constexpr (KIND == 6) fma2(p[(u + 1) % NACC], pb, pa);
      if constexpr (KIND == 5) {
        uint32_t hi, lo;
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(f[u]));
        const float r = f[u] - __uint_as_float(hi);
        asm volatile("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
        f[u] = __uint_as_float(hi ^ lo) + 1.0f;
      }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < NACC; ++u) {
    s += c[u][0] + c[u][1] + c[u][2] + c[u][3] + f[u];
    s += __uint_as_float((uint32_t)(p[u] >> 32)) + __uint_as_float((uint32_t)p[u]);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(const char* name, int sms, int warps_per_sm, double mac_per_instr, int instr_per_iter_unit) {
  const int threads = warps_per_sm * 32 > 1024 ? 1024 : warps_per_sm * 32;
  const int blocks_per_sm = (warps_per_sm * 32 + threads - 1) / threads;
  const int grid = sms * blocks_per_sm;
  float* out; long long* cyc;
  CK(cudaMalloc(&out, (size_t)grid * threads * 4));
  CK(cudaMalloc(&cyc, (size_t)grid * 8));
  bench<KIND><<<grid, threads>>>(out, cyc, 1u);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<KIND><<<grid, threads>>>(out, cyc, 2u);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long* h = (long long*)malloc((size_t)grid * 8);
  CK(cudaMemcpy(h, cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
  const double instr = (double)ITERS * NACC * instr_per_iter_unit * warps_per_sm;   // warp-instructions of the counted kind per SM
  printf("{\"test\": \"%s\", \"warps_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"mac_per_clk_per_sm\": %.1f, "
         "\"chip_tflops\": %.1f, \"ms\": %.4f, \"mean_block_cycles\": %.0f}\n",
         name, warps_per_sm, instr / mean, instr * mac_per_instr / mean,
         2.0 * instr * mac_per_instr * sms / (ms * 1e-3) / 1e12, ms, mean);
  fflush(stdout);
  free(h); cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", pr.name, pr.multiProcessorCount, pr.clockRate);
  const int sms = pr.multiProcessorCount;
  for (int w : {4, 8, 16, 32}) {
    run<0>("tf32_m16n8k8", sms, w, 1024, 1);
    run<1>("bf16_m16n8k16", sms, w, 2048, 1);
    run<2>("f16_m16n8k16", sms, w, 2048, 1);
    run<3>("tf32_m16n8k8_plus_ffma2 (mma counted)", sms, w, 1024, 1);
    run<6>("tf32_m16n8k8_plus_2ffma2 (mma counted)", sms, w, 1024, 1);
    run<4>("bf16_m16n8k16_plus_ffma2 (mma counted)", sms, w, 2048, 1);
    run<5>("cvt_split_tf32 (per element: 2 cvt + sub + xor + add)", sms, w, 0, 1);
  }
  return 0;
}
