// Pipe-rate microbenchmark for the rollout kernel design (B200, sm_100a).
//
// Measures, per SM and per clock, the issue rate of the instruction forms the
// fused policy-MLP + Euler-Maruyama step is made of:
//   ffma_rrr   FFMA with three register sources
//   ffma_rcr   FFMA with one constant-bank source (weights as kernel params)
//   ffma2_rrr  packed fma.rn.f32x2 (two FMAs per lane per instruction)
//   ffma2_lds  packed FMA fed by a broadcast LDS.128 per two instructions
//   mufu_*     ex2 / rcp / tanh / lg2 / rsq / sin / cos
//   mix_*      FFMA(const) : MUFU at the ratio of one MLP layer (16 : 1, 8 : 1)
//   imad_wide  32x32->64 multiply used by Philox
//   dfma       FP64 FMA (table builder CDF arithmetic)
// Each kernel runs ITERS iterations of an unrolled body of independent
// chains; cycles are taken from clock64() per block and wall time from CUDA
// events.  Output: one JSON line per test.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int ITERS = 4096;
constexpr int NACC = 16;

struct Weights { float w[256]; };

__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

template <int KIND>
__global__ void __launch_bounds__(256) bench(float* out, long long* cyc, const __grid_constant__ Weights W,
                                            float seed_a, float seed_b) {
  __shared__ __align__(16) float sw[256];
  sw[threadIdx.x] = W.w[threadIdx.x & 255] + seed_b;
  __syncthreads();
  float acc[NACC];
#pragma unroll
  for (int u = 0; u < NACC; ++u) acc[u] = seed_a * (u + 1) + threadIdx.x * 1e-6f;
  float a = seed_a, b = seed_b;
  unsigned long long p[NACC / 2 > 0 ? NACC : 1];
#pragma unroll
  for (int u = 0; u < NACC; ++u) p[u] = ((unsigned long long)__float_as_uint(acc[u]) << 32) | __float_as_uint(acc[u] + 1.f);
  double dacc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) dacc[u] = (double)acc[u];
  unsigned int ia = threadIdx.x * 2654435761u + 12345u, ib = 0xD2511F53u;
  unsigned long long iacc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) iacc[u] = ia + u;

  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    if constexpr (KIND == 0) {          // FFMA R,R,R
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < NACC; ++u) acc[u] = fmaf(a, b, acc[u]), a = a;  // a,b loop-invariant regs
    } else if constexpr (KIND == 1) {   // FFMA R,c[],R   (h_i * W[j][i] + acc_j pattern)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < NACC; ++u) acc[u] = fmaf(a, W.w[r * NACC + u], acc[u]);
    } else if constexpr (KIND == 2) {   // packed FFMA2 R,R,R
      unsigned long long pa = p[0] ^ 0x1000100010001ull, pb = p[1] ^ 0x2000200020002ull;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 2; u < NACC; ++u) fma2(p[u], pa, pb);
    } else if constexpr (KIND == 3) {   // packed FFMA2 fed by broadcast LDS.128 (one per two FFMA2)
      unsigned long long pa = p[0] ^ 0x1000100010001ull;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int u = 0; u < NACC; u += 2) {
          const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(&sw[((r * NACC + u) * 2) & 255]);
          fma2(p[u], pa, wv.x);
          fma2(p[u + 1], pa, wv.y);
        }
      }
    } else if constexpr (KIND >= 10 && KIND < 20) {  // MUFU family
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int u = 0; u < NACC; ++u) {
          float x = acc[u], y;
          if constexpr (KIND == 10) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 11) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 12) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 13) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 14) asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 15) asm volatile("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 16) asm volatile("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          if constexpr (KIND == 17) asm volatile("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
          acc[u] = y;
        }
    } else if constexpr (KIND == 20 || KIND == 21 || KIND == 22) {  // FFMA(const) : MUFU mixes
      constexpr int RATIO = (KIND == 20) ? 16 : (KIND == 21 ? 8 : 4);
      float m[4] = {acc[0], acc[1], acc[2], acc[3]};
#pragma unroll
      for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int q = 0; q < RATIO; ++q)
          acc[4 + (q % 12)] = fmaf(a, W.w[(r * RATIO + q) & 255], acc[4 + (q % 12)]);
        float y;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(m[r & 3]));
        m[r & 3] = y;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = m[u];
    } else if constexpr (KIND == 30) {  // IMAD.WIDE.U32 (Philox multiply)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) iacc[u] = (unsigned long long)(unsigned int)iacc[u] * ib + (iacc[u] >> 32);
    } else if constexpr (KIND == 31) {  // DFMA
      double da = (double)a, db = (double)b;
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < 8; ++u) dacc[u] = fma(da, db, dacc[u]);
    } else if constexpr (KIND == 32) {  // double erfc (table builder)
#pragma unroll
      for (int u = 0; u < 8; ++u) dacc[u] = erfc(dacc[u] * 0.3) + 0.1 * u;
    } else if constexpr (KIND == 33) {  // FFMA(const) + independent IADD3/LOP3 stream (alu pipe co-issue)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int u = 0; u < NACC; ++u) {
          acc[u] = fmaf(a, W.w[r * NACC + u], acc[u]);
          if ((u & 3) == 0) ia = (ia ^ ib) + (ia >> 3);
        }
    }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < NACC; ++u) s += acc[u] + __uint_as_float((unsigned)(p[u] >> 32)) + __uint_as_float((unsigned)p[u]);
#pragma unroll
  for (int u = 0; u < 8; ++u) s += (float)dacc[u] + (float)iacc[u];
  s += (float)ia;
  if (s == 123.456f) out[0] = s;  // keep everything live
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Case { const char* name; int kind; double ops_per_iter; const char* unit; };

template <int KIND>
static void run(const Case& c, int nsm, int blocks_per_sm, float* d_out, long long* d_cyc, const Weights& W) {
  const int grid = nsm * blocks_per_sm;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  bench<KIND><<<grid, 256>>>(d_out, d_cyc, W, 1.0001f, 0.9999f);  // warm-up
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  bench<KIND><<<grid, 256>>>(d_out, d_cyc, W, 1.0001f, 0.9999f);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)h[i]; mean /= grid;
  free(h);
  // warp-instructions per SM per clock: blocks_per_sm * 8 warps * ops_per_iter * ITERS / cycles
  const double winstr = (double)blocks_per_sm * 8.0 * c.ops_per_iter * ITERS;
  const double per_clk_sm = winstr / mean;
  const double lane_ops_per_s = winstr * 32.0 * nsm / (ms * 1e-3);
  printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"warp_instr_per_clk_per_sm\": %.3f, \"lane_ops_per_clk_per_sm\": %.1f, "
         "\"chip_lane_ops_per_s\": %.4e, \"ms\": %.4f, \"mean_block_cycles\": %.0f, \"implied_mhz\": %.0f}\n",
         c.name, blocks_per_sm, per_clk_sm, per_clk_sm * 32.0, lane_ops_per_s, ms, mean, mean / (ms * 1e3));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"cc\": \"%d.%d\"}\n", prop.name, nsm, prop.clockRate, prop.major, prop.minor);
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, 64)); CK(cudaMalloc(&d_cyc, sizeof(long long) * nsm * 16));
  Weights W; for (int i = 0; i < 256; ++i) W.w[i] = 1.0f + 1e-3f * i;
  for (int bps : {1, 2, 4}) {
    run<0>({"ffma_rrr", 0, 64, ""}, nsm, bps, d_out, d_cyc, W);
    run<1>({"ffma_rcr", 1, 64, ""}, nsm, bps, d_out, d_cyc, W);
    run<2>({"ffma2_rrr(x2 fma/lane)", 2, 56, ""}, nsm, bps, d_out, d_cyc, W);
    run<3>({"ffma2_lds128(ffma2 only counted)", 3, 64, ""}, nsm, bps, d_out, d_cyc, W);
    run<10>({"mufu_ex2", 10, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<11>({"mufu_rcp", 11, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<12>({"mufu_tanh", 12, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<13>({"mufu_lg2", 13, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<14>({"mufu_rsqrt", 14, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<15>({"mufu_sin", 15, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<16>({"mufu_cos", 16, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<17>({"mufu_sqrt", 17, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<20>({"mix_ffma16_mufu1(ffma counted)", 20, 8 * 16, ""}, nsm, bps, d_out, d_cyc, W);
    run<21>({"mix_ffma8_mufu1(ffma counted)", 21, 8 * 8, ""}, nsm, bps, d_out, d_cyc, W);
    run<22>({"mix_ffma4_mufu1(ffma counted)", 22, 8 * 4, ""}, nsm, bps, d_out, d_cyc, W);
    run<30>({"imad_wide_u32", 30, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<31>({"dfma", 31, 32, ""}, nsm, bps, d_out, d_cyc, W);
    run<32>({"erfc_f64(calls)", 32, 8, ""}, nsm, bps, d_out, d_cyc, W);
    run<33>({"ffma_rcr+alu(ffma counted)", 33, 64, ""}, nsm, bps, d_out, d_cyc, W);
  }
  return 0;
}
