// tcgen05 bring-up test for the wide-policy forward kernel (rl_sde_is_b200/csrc/rollout_umma.cuh): one CTA computes
//   D[128 x N] = A[128 x K] * B[N x K]^T      (f16 operands, fp32 accumulation in TMEM, N = K = 256)
// with the exact building blocks the kernel uses -- K-major SWIZZLE_NONE core-matrix layouts written by ordinary threads
// (A) and by cp.async.bulk from a pre-packed global image (B, one 16 KB chunk per 16-wide k-step through a ring of
// mbarrier-guarded stages), single-thread tcgen05.mma issue, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b epilogue --
// and checks the result against the host.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_test umma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int M = 128, N = 256, K = 256, KSTEPS = K / 16, STAGES = 4;
constexpr int CHUNK_BYTES = N * 16 * 2;            // one k-step of B: N rows x 16 halfs

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor: 8 x 16-byte core matrices, LBO = byte distance between the two
// core matrices of a k-step, SBO = byte distance between 8-row groups (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
  return d;                                     // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D f32, A / B f16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// A in shared memory: core(kc, rg) at (kc * 16 + rg) * 128 bytes, row r % 8 at +16 bytes, 8 halfs along k
__device__ __forceinline__ uint32_t a_offset(int row, int kc) { return (uint32_t)((kc * (M / 8) + (row >> 3)) * 128 + (row & 7) * 16); }

__global__ void __launch_bounds__(192, 1) umma_test_kernel(const __half* __restrict__ Ag, const uint8_t* __restrict__ Bimg, float* __restrict__ Dg) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sA = smem;                                     // 128 x 256 halfs = 64 KB
  uint8_t* sB = smem + M * K * 2;                          // STAGES x 8 KB
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES], a_ready, mma_done;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&a_ready, 128);
    mbar_init(&mma_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp < 4) {
    // ---- "trajectory" threads: thread t writes row t of A (as the rollout kernel's layer 1 will), then waits for D
    const int row = tid;
    for (int kc = 0; kc < K / 8; ++kc) {
      const uint4 v = *reinterpret_cast<const uint4*>(Ag + (size_t)row * K + 8 * kc);
      *reinterpret_cast<uint4*>(sA + a_offset(row, kc)) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
    mbar_arrive(&a_ready);
    mbar_wait(&mma_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int cb = 0; cb < N / 32; ++cb) {
      uint32_t r[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(32 * cb);
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
            "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
            "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
            "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int c = 0; c < 32; ++c) Dg[(size_t)row * N + 32 * cb + c] = __uint_as_float(r[c]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  } else if (warp == 4) {
    // ---- MMA issuer (one thread)
    if (lane == 0) {
      mbar_wait(&a_ready, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t idesc = make_idesc(M, N);
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int s = ks % STAGES;
        mbar_wait(&full[s], (ks / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t ad = make_desc(smem_u32(sA) + (uint32_t)ks * 2 * (M / 8) * 128, (M / 8) * 128, 128);
        const uint64_t bd = make_desc(smem_u32(sB) + (uint32_t)s * CHUNK_BYTES, (N / 8) * 128, 128);
        umma_f16(tmem_base, ad, bd, idesc, ks > 0 ? 1u : 0u);
        umma_commit(&empty[s]);                 // frees the stage when the MMAs issued so far have read it
      }
      umma_commit(&mma_done);
    }
  } else {
    // ---- B producer (one thread): chunk ks -> stage ks % STAGES
    if (lane == 0) {
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const int s = ks % STAGES;
        if (ks >= STAGES) mbar_wait(&empty[s], ((ks / STAGES) - 1) & 1);
        mbar_expect_tx(&full[s], CHUNK_BYTES);
        bulk_g2s(sB + (size_t)s * CHUNK_BYTES, Bimg + (size_t)ks * CHUNK_BYTES, CHUNK_BYTES, &full[s]);
      }
    }
  }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(N) : "memory");
}

int main() {
  std::vector<__half> A((size_t)M * K), B((size_t)N * K);
  std::vector<float> Af((size_t)M * K), Bf((size_t)N * K);
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 2001 - 1000) / 1000.f; A[i] = __float2half(v); Af[i] = __half2float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 2001 - 1000) / 4000.f; B[i] = __float2half(v); Bf[i] = __half2float(B[i]); }
  // B image: per k-step a chunk of N x 16 halfs in the core-matrix layout: core(kc, ng) at (kc * N/8 + ng) * 128 bytes
  std::vector<__half> img((size_t)N * K);
  for (int ks = 0; ks < KSTEPS; ++ks)
    for (int kc = 0; kc < 2; ++kc)
      for (int n = 0; n < N; ++n)
        for (int e = 0; e < 8; ++e)
          img[(size_t)ks * N * 16 + ((size_t)(kc * (N / 8) + (n >> 3)) * 64) + (n & 7) * 8 + e] = B[(size_t)n * K + 16 * ks + 8 * kc + e];
  __half *dA; uint8_t* dB; float* dD;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, img.size() * 2)); CK(cudaMalloc(&dD, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0, (size_t)M * N * 4));
  const size_t smem = (size_t)M * K * 2 + (size_t)STAGES * CHUNK_BYTES;
  CK(cudaFuncSetAttribute(umma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_test_kernel<<<1, 192, smem>>>(dA, dB, dD);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D((size_t)M * N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)m * K + k] * (double)Bf[(size_t)n * K + k];
      maxerr = fmax(maxerr, fabs(ref - (double)D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("{\"test\": \"umma 128x256x256 f16\", \"max_abs_err\": %.3e, \"max_abs_ref\": %.3e, \"ok\": %s}\n", maxerr, maxref,
         maxerr < 1e-3 * maxref ? "true" : "false");
  return maxerr < 1e-3 * maxref ? 0 : 2;
}
