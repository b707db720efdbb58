// "One hidden layer" microbenchmark: h <- tanh(W h + b), H = 32, one (or two)
// trajectories per thread, repeated ITERS times.  Candidate operand designs for
// the 32x32 GEMV of the fused rollout kernel:
//   A  weights as FFMA constant-bank operands (kernel parameter, __grid_constant__)
//   B  weights in shared memory, broadcast LDS.128 (4 weights / load), scalar FFMA
//   C  packed fma.rn.f32x2, one trajectory / thread, acc pairs (j,j+1), LDS.128 = 2 FFMA2
//   D  packed fma.rn.f32x2, two trajectories / thread share each weight load, LDS.128 = 4 FFMA2
// TANH: 0 = MUFU.TANH (fast), 1 = 1 - 2/(ex2(c*x)+1) (precise: MUFU.EX2 + MUFU.RCP)
// Reports trajectory-layers per second and FMA-lane utilisation vs 128 lanes/clk/SM.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int H = 32;
constexpr int ITERS = 2048;

struct LayerW { float W[H][H]; float b[H]; };  // W[i][j]: input index major, output index contiguous

template <int TANH>
__device__ __forceinline__ float act(float x) {
  float y;
  if constexpr (TANH == 0) {
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  } else {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.885390081777927f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    y = fmaf(-2.0f, r, 1.0f);
  }
  return y;
}

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ float4 lds128(const float* p) {
  float4 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ ulonglong2 lds128u(const float* p) {
  ulonglong2 v; unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ld.shared.v2.b64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void fma2(unsigned long long& d, unsigned long long a, unsigned long long b) {
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// ---- A: constant-bank weights -------------------------------------------------
template <int TANH>
__global__ void __launch_bounds__(128) layer_const(float* out, long long* cyc, const __grid_constant__ LayerW P, float x0) {
  float h[H];
#pragma unroll
  for (int i = 0; i < H; ++i) h[i] = x0 * (i + 1) + threadIdx.x * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; ++j) acc[j] = P.b[j];
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int j = 0; j < H; ++j) acc[j] = fmaf(h[i], P.W[i][j], acc[j]);
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = act<TANH>(acc[j]);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < H; ++i) s += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- B: smem weights, broadcast LDS.128, scalar FFMA --------------------------
template <int TANH>
__global__ void __launch_bounds__(128) layer_smem(float* out, long long* cyc, const LayerW* __restrict__ Pg, float x0) {
  __shared__ __align__(16) float sW[H][H];  // sW[i][j] = W[j][i]
  __shared__ __align__(16) float sb[H];
  for (int t = threadIdx.x; t < H * H; t += blockDim.x) sW[t / H][t % H] = Pg->W[t / H][t % H];
  if (threadIdx.x < H) sb[threadIdx.x] = Pg->b[threadIdx.x];
  __syncthreads();
  float h[H];
#pragma unroll
  for (int i = 0; i < H; ++i) h[i] = x0 * (i + 1) + threadIdx.x * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    float acc[H];
#pragma unroll
    for (int j = 0; j < H; j += 4) {
      const float4 b4 = *reinterpret_cast<const float4*>(&sb[j]);
      acc[j] = b4.x; acc[j + 1] = b4.y; acc[j + 2] = b4.z; acc[j + 3] = b4.w;
    }
#pragma unroll
    for (int i = 0; i < H; ++i)
#pragma unroll
      for (int j = 0; j < H; j += 4) {
        const float4 w = lds128(&sW[i][j]);
        acc[j] = fmaf(h[i], w.x, acc[j]);
        acc[j + 1] = fmaf(h[i], w.y, acc[j + 1]);
        acc[j + 2] = fmaf(h[i], w.z, acc[j + 2]);
        acc[j + 3] = fmaf(h[i], w.w, acc[j + 3]);
      }
#pragma unroll
    for (int j = 0; j < H; ++j) h[j] = act<TANH>(acc[j]);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < H; ++i) s += h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- C / D: packed FFMA2, NT trajectories per thread --------------------------
template <int TANH, int NT>
__global__ void __launch_bounds__(128) layer_ffma2(float* out, long long* cyc, const LayerW* __restrict__ Pg, float x0) {
  __shared__ __align__(16) float sW[H][H];  // sW[i][j] = W[j][i]
  __shared__ __align__(16) float sb[H];
  for (int t = threadIdx.x; t < H * H; t += blockDim.x) sW[t / H][t % H] = Pg->W[t / H][t % H];
  if (threadIdx.x < H) sb[threadIdx.x] = Pg->b[threadIdx.x];
  __syncthreads();
  float h[NT][H];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < H; ++i) h[t][i] = x0 * (i + 1) + (threadIdx.x + 128 * t) * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    unsigned long long acc[NT][H / 2];
#pragma unroll
    for (int j = 0; j < H / 2; j += 2) {
      const ulonglong2 b4 = *reinterpret_cast<const ulonglong2*>(&sb[2 * j]);
#pragma unroll
      for (int t = 0; t < NT; ++t) { acc[t][j] = b4.x; acc[t][j + 1] = b4.y; }
    }
#pragma unroll
    for (int i = 0; i < H; ++i) {
      unsigned long long hh[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) hh[t] = pack2(h[t][i], h[t][i]);
#pragma unroll
      for (int j = 0; j < H / 2; j += 2) {
        const ulonglong2 w = lds128u(&sW[i][2 * j]);
#pragma unroll
        for (int t = 0; t < NT; ++t) { fma2(acc[t][j], hh[t], w.x); fma2(acc[t][j + 1], hh[t], w.y); }
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < H / 2; ++j) {
        float lo, hi; unpack2(acc[t][j], lo, hi);
        h[t][2 * j] = act<TANH>(lo); h[t][2 * j + 1] = act<TANH>(hi);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < H; ++i) s += h[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- E: packed FFMA2 with constant-bank weight pairs, NT trajectories / thread ----
template <int TANH, int NT>
__global__ void __launch_bounds__(128) layer_ffma2_const(float* out, long long* cyc, const __grid_constant__ LayerW P, float x0) {
  float h[NT][H];
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < H; ++i) h[t][i] = x0 * (i + 1) + (threadIdx.x + 128 * t) * 1e-4f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
    unsigned long long acc[NT][H / 2];
#pragma unroll
    for (int j = 0; j < H / 2; ++j)
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t][j] = pack2(P.b[2 * j], P.b[2 * j + 1]);
#pragma unroll
    for (int i = 0; i < H; ++i) {
      unsigned long long hh[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) hh[t] = pack2(h[t][i], h[t][i]);
#pragma unroll
      for (int j = 0; j < H / 2; ++j) {
        const unsigned long long w = pack2(P.W[i][2 * j], P.W[i][2 * j + 1]);
#pragma unroll
        for (int t = 0; t < NT; ++t) fma2(acc[t][j], hh[t], w);
      }
    }
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
      for (int j = 0; j < H / 2; ++j) {
        float lo, hi; unpack2(acc[t][j], lo, hi);
        h[t][2 * j] = act<TANH>(lo); h[t][2 * j + 1] = act<TANH>(hi);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int t = 0; t < NT; ++t)
#pragma unroll
    for (int i = 0; i < H; ++i) s += h[t][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
static void time_it(const char* name, int traj_per_thread, int nsm, int bps, long long* d_cyc, F launch) {
  const int grid = nsm * bps;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(grid); CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0)); launch(grid); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* hc = (long long*)malloc(sizeof(long long) * grid);
  CK(cudaMemcpy(hc, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double mean = 0; for (int i = 0; i < grid; ++i) mean += (double)hc[i]; mean /= grid; free(hc);
  const double layers = (double)grid * 128 * traj_per_thread * ITERS;
  const double fma_per_clk_sm = (double)bps * 128 * traj_per_thread * ITERS * (H * H) / mean;
  printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"traj_layers_per_s\": %.4e, \"fma_lanes_per_clk_per_sm\": %.1f, "
         "\"frac_of_128\": %.3f, \"ms\": %.3f, \"implied_mhz\": %.0f}\n",
         name, bps, layers / (ms * 1e-3), fma_per_clk_sm, fma_per_clk_sm / 128.0, ms, mean / (ms * 1e3));
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  LayerW hw;
  for (int j = 0; j < H; ++j) { hw.b[j] = 0.01f * j; for (int i = 0; i < H; ++i) hw.W[j][i] = 0.17f * (((j * 7 + i * 13) % 11) - 5) / 5.0f; }
  LayerW* dW; CK(cudaMalloc(&dW, sizeof(LayerW))); CK(cudaMemcpy(dW, &hw, sizeof(LayerW), cudaMemcpyHostToDevice));
  float* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(float) * nsm * 16 * 128)); CK(cudaMalloc(&d_cyc, sizeof(long long) * nsm * 16));
  for (int bps : {1, 2, 3, 4, 6, 8}) {
    time_it("A_const_tanhfast", 1, nsm, bps, d_cyc, [&](int g) { layer_const<0><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
    time_it("A_const_tanhprecise", 1, nsm, bps, d_cyc, [&](int g) { layer_const<1><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
    time_it("B_smem_tanhfast", 1, nsm, bps, d_cyc, [&](int g) { layer_smem<0><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
    time_it("B_smem_tanhprecise", 1, nsm, bps, d_cyc, [&](int g) { layer_smem<1><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
    time_it("C_ffma2x1_tanhfast", 1, nsm, bps, d_cyc, [&](int g) { layer_ffma2<0, 1><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
    time_it("C_ffma2x1_tanhprecise", 1, nsm, bps, d_cyc, [&](int g) { layer_ffma2<1, 1><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
    time_it("E_ffma2const_x1_tanhfast", 1, nsm, bps, d_cyc, [&](int g) { layer_ffma2_const<0, 1><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
    time_it("E_ffma2const_x1_tanhprecise", 1, nsm, bps, d_cyc, [&](int g) { layer_ffma2_const<1, 1><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
    if (bps <= 4) {
      time_it("E_ffma2const_x2_tanhfast", 2, nsm, bps, d_cyc, [&](int g) { layer_ffma2_const<0, 2><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
      time_it("E_ffma2const_x2_tanhprecise", 2, nsm, bps, d_cyc, [&](int g) { layer_ffma2_const<1, 2><<<g, 128>>>(d_out, d_cyc, hw, 0.01f); });
      time_it("D_ffma2x2_tanhfast", 2, nsm, bps, d_cyc, [&](int g) { layer_ffma2<0, 2><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
      time_it("D_ffma2x2_tanhprecise", 2, nsm, bps, d_cyc, [&](int g) { layer_ffma2<1, 2><<<g, 128>>>(d_out, d_cyc, dW, 0.01f); });
    }
  }
  return 0;
}
