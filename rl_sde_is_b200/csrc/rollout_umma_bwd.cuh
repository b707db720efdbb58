// K2u: reverse pass (pathwise gradient of the REINFORCE loss) for wide policies (hidden width 128 / 256) on the
// 5th-generation tensor cores.  Same recursion, arguments and outputs as K2x (rollout_wide.cuh: rollout_bwd_wide_kernel;
// the reference's eff_loss.backward(), reinforce_deterministic_core.py:240, through a policy of the width reinforce()
// defaults to, :102); ckpt_every must be 1 (every state is read from the stored path).
//
// The three H x H products of a pass of a 128-trajectory tile,
//     Z2  = H1  W2^T   (recomputed forward)     M = trajectory, K = in,  N = out
//     dH1 = dZ2 W2                               M = trajectory, K = out, N = in
//     dW2 += dZ2^T H1  (sum over trajectories)   M = out,        K = trajectory, N = in
// do not fit one SM: the first two need 2 H accumulator columns of tensor memory, and the running H x H gradient block
// (256 KB at H = 256) is exactly one SM's tensor memory on its own.  So the grid is split into two ROLES:
//   * PRODUCER CTAs (about four fifths of the SMs) run the recursion.  Threads 0-127 own one trajectory each (= one TMEM
//     lane), exactly as in K1u: layer 1 into the A-operand ring (float16 hi / lo), accumulator read back row by row.
//     Sweep 1 reads Z2, applies bias + tanh, sums the head and parks h2 in the same tensor-memory columns; the owner then
//     forms its a_j (K2's formula) and sweep 2 turns h2 into dz2 = (W3^T a)(1 - h2^2), which goes -- scaled by a power of
//     two 2^esh chosen per launch from max |G| (ub_scale_kernel) so that float16 holds it, hi / lo -- into the A ring again; sweep 3 reads
//     dH1, forms dz1 and dX and updates the adjoint.  Warp 4 issues the MMAs, warp 5 streams both weight images
//     (W2 as [out][in] and as [in][out]) through the mbarrier ring, as in K1u.
//     The small gradient blocks (b1, W1, b2, W3) are column sums over trajectories: 32-lane butterfly reductions of each
//     32-column chunk, accumulated per warp in shared memory and flushed to float64 partials per tile.
//   * Every pass the owners also write h1 and dz2 (the same hi / lo halves) to an EXCHANGE buffer in global memory, laid
//     out as the MN-major operands of the third product, one 16-trajectory k-step per contiguous chunk, and publish the
//     pass with a release store.  A ring of UB_RING passes per producer keeps the buffer L2-sized.
//   * CONSUMER CTAs (the remaining SMs; one serves up to four producers, strictly round-robin so that the summation order
//     -- and the result -- is deterministic) copy the chunks into shared memory and accumulate dW2 += dZ2^T H1 in their
//     whole tensor memory (H/128 blocks of 128 x H fp32), draining it into a float64 partial every UB_FLUSH passes.
// All CTAs must be co-resident (producers wait for consumers and vice versa): the launcher sizes the grid to the SM
// count, and either role needs more than half an SM's shared memory or tensor memory, so one CTA per SM.
#pragma once
#include "rollout_umma.cuh"

namespace rlsde {

constexpr int UB_THREADS = 320;          // producer: 8 owner warps + MMA issuer + weight loader
#ifndef UB_RING_N
#define UB_RING_N 4
#endif
constexpr int UB_RING = UB_RING_N;             // passes per producer in the exchange buffer
#ifndef UB_CSTAGES_N
#define UB_CSTAGES_N 6
#endif
constexpr int UB_CSTAGES = UB_CSTAGES_N;          // consumer: k-step chunks in shared memory
#ifndef UB_FLUSH_N
#define UB_FLUSH_N 16
#endif
constexpr int UB_FLUSH = UB_FLUSH_N;           // consumer: passes between drains of the tensor-memory accumulator
#ifndef UB_PPC_N
#define UB_PPC_N 2
#endif
constexpr int UB_PPC = UB_PPC_N;              // producers per consumer
constexpr int UB_ASLOTS = 4;           // producer: A-operand ring slots

template <int H> __host__ __device__ constexpr size_t ub_chunk_bytes() { return (size_t)H * 128; }      // A_hi | A_lo | B_hi | B_lo of 16 trajectories
template <int H> __host__ __device__ constexpr size_t ub_pass_bytes() { return 8 * ub_chunk_bytes<H>(); }
template <int D> __host__ __device__ constexpr int ub_nq() { return 2 + 2 * D; }                          // b1, b2, W1 [D], W3 [D] column blocks
template <int D, int H> __host__ __device__ constexpr int ub_small_count() { return ub_nq<D>() * H + D; }  // + b3
template <int D, int H>
__host__ __device__ constexpr size_t ub_smem_bytes() {
  const size_t prod = (size_t)UB_ASLOTS * UMMA_ACHUNK + (size_t)UMMA_STAGES * umma_chunk_bytes<H>() +
                      (size_t)(2 * D + 2) * H * sizeof(float) + (size_t)4 * ub_nq<D>() * H * sizeof(float);
  const size_t cons = (size_t)UB_CSTAGES * ub_chunk_bytes<H>();
  const size_t m = prod > cons ? prod : cons;
  return m > (size_t)116 * 1024 ? m : (size_t)116 * 1024;          // more than half an SM: one CTA per SM
}
// control words per producer (global, zeroed by the launcher): [0] passes published, [1] passes consumed, [2] finished
constexpr int UB_CTL_WORDS = 16;        // ([4..11]: cycle counters of the UB_PROFILE build; [12]: esh, the launch's scale exponent)

#ifdef UB_PROFILE
#define UB_T(i) do { if (tid == 0) { const long long t_ = clock64(); prof[i] += t_ - tprev; tprev = t_; } } while (0)
#else
#define UB_T(i) do { } while (0)
#endif

namespace umma {
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_cg(const void* p) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void bar_owners() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// v[c], c = 0..31, one value per column held by every lane  ->  returns on lane l the sum over the 32 lanes of v[l]
__device__ __forceinline__ float lane_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}
}  // namespace umma

// partial layout (float64): [n_cons][H * H] (dW2 in 2^esh units, row = out unit), then [n_prod][ub_small_count]
template <int D, int H, bool FAST>
__global__ void __launch_bounds__(UB_THREADS, 1)
rollout_bwd_umma_kernel(const float* __restrict__ Wp, const uint8_t* __restrict__ Bimg1, const uint8_t* __restrict__ Bimg2,
                        const __grid_constant__ FwdArgs A, uint8_t* __restrict__ xbuf, unsigned* __restrict__ ctl,
                        double* __restrict__ partial, int n_prod, int n_cons) {
  using namespace umma;
  typedef WideParams<D, H> L;
  constexpr int KSTEPS = H / 16;
  constexpr int NQ = ub_nq<D>();
  constexpr int MB = H / 128;                                  // 128-row blocks of dW2
  constexpr uint32_t CHUNK = (uint32_t)umma_chunk_bytes<H>();  // one k-step of a weight image (hi + lo)
  constexpr uint32_t XCH = (uint32_t)ub_chunk_bytes<H>();      // one k-step (16 trajectories) of the exchange
  constexpr uint32_t PLANE = (uint32_t)H * 32;                 // one of the four planes of an exchange chunk
  extern __shared__ __align__(128) uint8_t umma_smem[];
  __shared__ __align__(8) uint64_t full[UMMA_STAGES], empty[UMMA_STAGES], a_full[UB_ASLOTS], a_empty[UB_ASLOTS], accZ, accG;
  __shared__ __align__(8) uint64_t cfull[UB_CSTAGES], cempty[UB_CSTAGES], acc_ready, flushed;
  __shared__ uint32_t tmem_base_s;
  __shared__ int s_np[4], s_cmd[UB_CSTAGES], s_have;
  __shared__ float s_b3[4][D];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool is_prod = (int)blockIdx.x < n_prod;
  const uint32_t tmem_cols = is_prod ? (uint32_t)(2 * H) : (uint32_t)(MB * H);

  if (tid == 0) {
    for (int s = 0; s < UMMA_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < UB_ASLOTS; ++s) { mbar_init(&a_full[s], UMMA_M); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < UB_CSTAGES; ++s) { mbar_init(&cfull[s], 1); mbar_init(&cempty[s], 1); }
    mbar_init(&accZ, 1); mbar_init(&accG, 1); mbar_init(&acc_ready, 1); mbar_init(&flushed, UMMA_M);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  if (!is_prod) {
    // =========================================================================================== consumer
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int c = (int)blockIdx.x - n_prod;
    uint8_t* const stage0 = umma_smem;
    double* const myW2 = partial + (size_t)c * H * H;
    enum { CMD_DATA = 1, CMD_FLUSH = 2, CMD_END = 3, CMD_END_NODRAIN = 4 };
    if (warp == 5 && lane == 0) {
      // ---- loader: polls the producers' sequence numbers and moves their passes, one bulk copy per k-step chunk
      unsigned cc = 0;                         // running chunk index (data and command chunks)
      unsigned fin = 0;                        // bit i: producer i of this consumer has no more passes
      int since = 0;
      int npp = 0;
      for (int i = 0; i < UB_PPC; ++i) if (c + i * n_cons < n_prod) npp = i + 1;
      // a pass's slot goes back to its producer once all its chunks have landed; that is known for free when the stage
      // of its last chunk comes round again, so the release trails by up to UB_CSTAGES chunks (two passes can be pending)
      unsigned* pend_ctl[2] = {nullptr, nullptr};
      unsigned pend_seq[2] = {0, 0}, pend_last[2] = {0, 0};
      int pend_n = 0;
      auto release_landed = [&](bool all) {
        while (pend_n > 0 && (all || pend_last[0] + UB_CSTAGES < cc + 1u)) {
          if (all)         // wait for the pass's chunks whose stage has not been re-armed since (the others were multiplied)
            for (unsigned q = pend_last[0] - 7u; q <= pend_last[0]; ++q)
              if (q + UB_CSTAGES >= cc) mbar_wait(&cfull[q % UB_CSTAGES], (q / UB_CSTAGES) & 1u);
          st_release(pend_ctl[0] + 1, pend_seq[0]);
          pend_ctl[0] = pend_ctl[1]; pend_seq[0] = pend_seq[1]; pend_last[0] = pend_last[1];
          --pend_n;
        }
      };
      auto post = [&](int cmd) {
        const int st = cc % UB_CSTAGES;
        if (cc >= UB_CSTAGES) mbar_wait(&cempty[st], ((cc / UB_CSTAGES) - 1u) & 1u);
        release_landed(false);
        *(volatile int*)&s_cmd[st] = cmd;
        mbar_arrive(&cfull[st]);
        ++cc;
      };
      for (unsigned n = 0; fin != (1u << npp) - 1u; ++n) {
        for (int i = 0; i < npp; ++i) {
          if (fin & (1u << i)) continue;
          const int p = c + i * n_cons;
          unsigned* const pc = ctl + (size_t)p * UB_CTL_WORDS;
          bool have;
          for (;;) {
            if (ld_acquire(pc) > n) { have = true; break; }
            if (ld_acquire(pc + 2)) { have = ld_acquire(pc) > n; break; }
            __nanosleep(32);
          }
          if (!have) { fin |= 1u << i; continue; }
          asm volatile("fence.proxy.async;" ::: "memory");         // the producer's generic-proxy stores -> this thread's bulk copies
          const uint8_t* src = xbuf + ((size_t)p * UB_RING + (n % UB_RING)) * ub_pass_bytes<H>();
#pragma unroll 1
          for (int g = 0; g < 8; ++g) {
            const int st = cc % UB_CSTAGES;
            if (cc >= UB_CSTAGES) mbar_wait(&cempty[st], ((cc / UB_CSTAGES) - 1u) & 1u);
            release_landed(false);              // chunk cc - UB_CSTAGES has been multiplied, so it and all before it have landed
            *(volatile int*)&s_cmd[st] = CMD_DATA;
            mbar_expect_tx(&cfull[st], XCH);
            bulk_g2s(stage0 + (size_t)st * XCH, src + (size_t)g * XCH, XCH, &cfull[st]);
            ++cc;
          }
          if (pend_n == 2) release_landed(true);
          pend_ctl[pend_n] = pc; pend_seq[pend_n] = n + 1; pend_last[pend_n] = cc - 1u; ++pend_n;
          if (++since == UB_FLUSH) { post(CMD_FLUSH); since = 0; }
        }
      }
      release_landed(true);
      post(since > 0 ? CMD_END : CMD_END_NODRAIN);
    } else if (warp < 4) {
      // ---- drain crew: tensor-memory accumulator -> float64 partial, on the MMA thread's request
      for (unsigned nflush = 0;; ++nflush) {
        mbar_wait(&acc_ready, nflush & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int cmd = *(volatile int*)&s_have;
        for (int mb = 0; cmd != CMD_END_NODRAIN && mb < MB; ++mb) {
          double* row = myW2 + (size_t)(128 * mb + 32 * warp + lane) * H;
#pragma unroll 1
          for (int cb = 0; cb < H / 32; ++cb) {
            uint32_t r[32];
            ld32(tmem_base + ((uint32_t)(32 * warp) << 16) + (uint32_t)(mb * H + 32 * cb), r);
#pragma unroll
            for (int k = 0; k < 32; ++k) row[32 * cb + k] += (double)__uint_as_float(r[k]);
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&flushed);
        if (cmd != CMD_FLUSH) break;
      }
    } else if (warp == 4 && lane == 0) {
      constexpr uint32_t idesc = make_idesc(UMMA_M, H) | (1u << 15) | (1u << 16);     // A and B MN-major
      bool first = true;
      unsigned nflush = 0;
      for (unsigned k = 0;; ++k) {
        const int st = k % UB_CSTAGES;
        mbar_wait(&cfull[st], (k / UB_CSTAGES) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int cmd = *(volatile int*)&s_cmd[st];
        if (cmd == CMD_DATA) {
          const uint32_t base = smem_u32(stage0) + (uint32_t)st * XCH;
          // MN-major, no swizzle: 8 x 8 core matrices (8 trajectories x 16 bytes of units); LBO = next 8 trajectories
          // (128 B), SBO = next 8 units (256 B)
          const uint64_t bh = make_desc(base + 2 * PLANE, 128, 256);
          const uint64_t bl = make_desc(base + 3 * PLANE, 128, 256);
#pragma unroll
          for (int mb = 0; mb < MB; ++mb) {
            const uint64_t ah = make_desc(base + (uint32_t)mb * 16u * 256u, 128, 256);
            const uint64_t al = make_desc(base + PLANE + (uint32_t)mb * 16u * 256u, 128, 256);
            const uint32_t d = tmem_base + (uint32_t)(mb * H);
            mma_f16(d, ah, bh, idesc, first ? 0u : 1u);
            mma_f16(d, al, bh, idesc, 1u);
            mma_f16(d, ah, bl, idesc, 1u);
          }
          first = false;
          commit(&cempty[st]);
        } else {
          // every MMA so far must have landed before the crew reads the accumulator: wait for the commit here, then hand
          // over with an ordinary (release) arrive
          commit(&cempty[st]);
          commit(&accG);
          mbar_wait(&accG, nflush & 1u);
          *(volatile int*)&s_have = cmd;
          mbar_arrive(&acc_ready);
          mbar_wait(&flushed, nflush & 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          ++nflush;
          first = true;
          if (cmd != CMD_FLUSH) break;
        }
      }
    }
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    return;
  }

  // ============================================================================================= producer
  // Threads 0-255 are the owners: TWO threads per trajectory (tr = tid & 127 is the row = TMEM lane; warps w and w + 4
  // reach the same 32 lanes), each taking every other 32-unit chunk of the hidden layers (cb = 2 i + half).  With one
  // thread per row a scheduler holds a single warp and every MUFU / tensor-memory latency is exposed (measured: 24 us of
  // owner work per pass); two warps per scheduler overlap them.  Partial head sums and partial dX meet in shared memory.
  uint8_t* sA = umma_smem;                                           // UB_ASLOTS x [hi 4 KB | lo 4 KB]
  uint8_t* sB = sA + UB_ASLOTS * UMMA_ACHUNK;                        // UMMA_STAGES weight chunks
  float* sW1 = reinterpret_cast<float*>(sB + UMMA_STAGES * CHUNK);   // [D][H]
  float* sb1 = sW1 + D * H;
  float* sb2 = sb1 + H;
  float* sW3 = sb2 + H;                                              // [D][H]
  float* sAcc = sW3 + D * H;                                         // [4 row groups][NQ][H]: b1, b2, W1[D], W3[D] column sums
  __shared__ float s_u[2][UMMA_M][D], s_dx[2][UMMA_M][D];
  for (int i = tid; i < H; i += blockDim.x) {
    sb1[i] = __ldg(Wp + L::o_b1 + i);
    sb2[i] = __ldg(Wp + L::o_b2 + i);
#pragma unroll
    for (int k = 0; k < D; ++k) { sW1[k * H + i] = __ldg(Wp + L::o_W1t + (size_t)k * H + i); sW3[k * H + i] = __ldg(Wp + L::o_W3 + (size_t)k * H + i); }
  }
  for (int i = tid; i < 4 * NQ * H; i += blockDim.x) sAcc[i] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmemZ = tmem_base, tmemG = tmem_base + (uint32_t)H;
  const int p = (int)blockIdx.x;
  unsigned* const pc = ctl + (size_t)p * UB_CTL_WORDS;
  double* const mySmall = partial + (size_t)n_cons * H * H + (size_t)p * ub_small_count<D, H>();
  const bool owner = warp < 8;
  const int tr = tid & (UMMA_M - 1), half = (tid >> 7) & 1, rgrp = warp & 3;
  const bool inject = (A.flags & RLSDE_F_NOISE_INJECTED) != 0;
  const bool s_exact = (A.flags & RLSDE_F_STOCH_INT_EXACT) != 0;
  const float inv_s = FAST ? 1.0f : (float)(1.0 / RLSDE_TWO_LOG2E);
  const int esh = (int)pc[12];                                                // set by ub_scale_kernel for this launch
  const float dscale = __uint_as_float((unsigned)(127 + esh) << 23);          // 2^esh
  const float dunscale = __uint_as_float((unsigned)(127 - esh) << 23);
  float b3[D];
#pragma unroll
  for (int k = 0; k < D; ++k) b3[k] = __ldg(Wp + L::o_b3 + k);
#ifdef UB_PROFILE
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
  unsigned kc = 0;            // running k-step index (A slots and weight stages advance together)
  unsigned pcount = 0;        // running pass index = exchange sequence number
  unsigned known_consumed = 0;   // thread 0: passes the consumer is known to have taken
  const long long n_tiles = (A.K + UMMA_M - 1) / UMMA_M;

  // tiles (length-sorted by the caller) are dealt out boustrophedon: round r goes p = 0 .. n_prod-1 when r is even and
  // back when it is odd, so every producer's total is close to the mean (static, hence deterministic)
  for (long long rnd = 0; rnd * n_prod < n_tiles; ++rnd) {
    const long long tile = rnd * n_prod + ((rnd & 1) ? (n_prod - 1 - p) : p);
    if (tile >= n_tiles) continue;
    // ---- owners pick up their trajectory; the tile runs max (T + 1) passes
    bool alive = false;
    long long traj = 0;
    int kstar = 0, j = -1;
    float Gk = 0.f, lam[D], gb3[D], xn[D], xf[D];
    NoiseCache<D> nc;
    nc.reset();
#pragma unroll
    for (int i = 0; i < D; ++i) { lam[i] = 0.f; gb3[i] = 0.f; xn[i] = A.x0_f[i]; xf[i] = A.x0_f[i]; }
    if (owner) {
      const long long slot = tile * UMMA_M + tr;
      if (slot < A.K) {
        traj = A.order ? A.order[slot] : slot;
        const int t = A.T[traj];
        if (t >= 0) { alive = true; kstar = t; j = t; Gk = ((const float*)A.G)[traj]; }
      }
      const int np = __reduce_max_sync(0xffffffffu, alive ? j + 1 : 0);
      if (lane == 0 && half == 0) s_np[rgrp] = np;
      if (alive) {
#pragma unroll
        for (int i = 0; i < D; ++i) xn[i] = A.path[((long long)traj * A.ckpt_stride + j) * D + i];
      }
    }
    __syncthreads();
    const int n_pass = max(max(s_np[0], s_np[1]), max(s_np[2], s_np[3]));

    if (owner) {
      const uint32_t row_off = (uint32_t)((tr >> 3) * 128 + (tr & 7) * 16);                       // K-major A image (MMA 1, 2)
      const uint32_t xoff = (uint32_t)(tr >> 4) * XCH + (uint32_t)((tr >> 3) & 1) * 128u + (uint32_t)(tr & 7) * 16u;   // exchange
      const uint32_t tlane = (uint32_t)(32 * rgrp) << 16;
      float* const myAcc = sAcc + (size_t)rgrp * NQ * H;
      bool live = false;
#pragma unroll 1
      for (int ps = 0; ps < n_pass; ++ps, ++pcount, kc += 2 * KSTEPS) {
        const uint32_t par = pcount & 1u;
        UB_T(7);
        // ---- the exchange slot of this pass must have been consumed
        // (thread 0 remembers the last count it read: the acquire load is an L2 round trip, and a count read k passes ago
        // usually still proves the slot free)
        if (tid == 0 && known_consumed + (unsigned)UB_RING <= pcount) {
          while ((known_consumed = ld_acquire(pc + 1)) + (unsigned)UB_RING <= pcount) __nanosleep(32);
        }
        bar_owners();
        UB_T(0);
        // ---- adjoint state after the previous pass (both halves keep a copy; the partial dX are added in a fixed order)
        if (ps > 0 && live) {
#pragma unroll
          for (int i = 0; i < D; ++i) {
            const float dxs = (s_dx[0][tr][i] + s_dx[1][tr][i]) * inv_s;
            const float hess = A.c4a_f[i] * fmaf(3.0f * xf[i], xf[i], -1.0f);
            lam[i] = fmaf(lam[i], fmaf(-A.dt_f, hess, 1.0f), dxs);
          }
          --j;
        }
        uint8_t* const xs = xbuf + ((size_t)p * UB_RING + (pcount % UB_RING)) * ub_pass_bytes<H>() + xoff;
        live = alive && j >= 0;
#pragma unroll
        for (int i = 0; i < D; ++i) xf[i] = xn[i];
        if (live && j >= 1) {                      // next pass's state: issued now, used in a pass's time
#pragma unroll
          for (int i = 0; i < D; ++i) xn[i] = A.path[((long long)traj * A.ckpt_stride + (j - 1)) * D + i];
        }

        // ---- layer 1 -> A ring (K-major, MMA 1) and exchange B planes (MN-major operand of dW2); this half's chunks
#pragma unroll 1
        for (int ci = 0; ci < H / 64; ++ci) {
          const int cb = 2 * ci + half;
#pragma unroll 1
          for (int hs = 0; hs < 2; ++hs) {
            const int ks = 2 * cb + hs;
            const unsigned c = kc + ks;
            const int slot = c % UB_ASLOTS;
            if (c >= UB_ASLOTS) mbar_wait(&a_empty[slot], ((c / UB_ASLOTS) - 1u) & 1u);
            uint8_t* dst = sA + (size_t)slot * UMMA_ACHUNK + row_off;
#pragma unroll
            for (int kc2 = 0; kc2 < 2; ++kc2) {
              const int ub = 2 * ks + kc2;
              float h[8];
#pragma unroll
              for (int q4 = 0; q4 < 2; ++q4) {
                const float4 bb = *reinterpret_cast<const float4*>(sb1 + 8 * ub + 4 * q4);
                float zz[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int k = 0; k < D; ++k) {
                  const float4 ww = *reinterpret_cast<const float4*>(sW1 + k * H + 8 * ub + 4 * q4);
                  zz[0] = fmaf(xf[k], ww.x, zz[0]); zz[1] = fmaf(xf[k], ww.y, zz[1]);
                  zz[2] = fmaf(xf[k], ww.z, zz[2]); zz[3] = fmaf(xf[k], ww.w, zz[3]);
                }
                tanh_pair<FAST>(pack2(zz[0], zz[1]), h[4 * q4], h[4 * q4 + 1]);
                tanh_pair<FAST>(pack2(zz[2], zz[3]), h[4 * q4 + 2], h[4 * q4 + 3]);
              }
              uint4 hi, lo;
              split_pair(h[0], h[1], hi.x, lo.x);
              split_pair(h[2], h[3], hi.y, lo.y);
              split_pair(h[4], h[5], hi.z, lo.z);
              split_pair(h[6], h[7], hi.w, lo.w);
              const uint32_t off = (uint32_t)kc2 * (UMMA_M / 8) * 128;
              *reinterpret_cast<uint4*>(dst + off) = hi;
              *reinterpret_cast<uint4*>(dst + UMMA_ACHUNK / 2 + off) = lo;
              *reinterpret_cast<uint4*>(xs + 2 * PLANE + (uint32_t)ub * 256u) = hi;
              *reinterpret_cast<uint4*>(xs + 3 * PLANE + (uint32_t)ub * 256u) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&a_full[slot]);
          }
        }

        // ---- sweep 1: h2 = tanh(z2 + b2) parked in the accumulator's columns, head summed (this half's chunks)
        UB_T(1);
        mbar_wait(&accZ, par);
        UB_T(2);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float u[D];
#pragma unroll
        for (int k = 0; k < D; ++k) u[k] = 0.f;
#pragma unroll 1
        for (int ci = 0; ci < H / 64; ++ci) {
          const int cb = 2 * ci + half;
          uint32_t r[32];
          const uint32_t ta = tmemZ + tlane + (uint32_t)(32 * cb);
          ld32(ta, r);
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 bb = *reinterpret_cast<const float4*>(sb2 + 32 * cb + 4 * c4);
            float h2[4];
            tanh_pair<FAST>(pack2(__uint_as_float(r[4 * c4]) + bb.x, __uint_as_float(r[4 * c4 + 1]) + bb.y), h2[0], h2[1]);
            tanh_pair<FAST>(pack2(__uint_as_float(r[4 * c4 + 2]) + bb.z, __uint_as_float(r[4 * c4 + 3]) + bb.w), h2[2], h2[3]);
#pragma unroll
            for (int k = 0; k < D; ++k) {
              const float4 w3 = *reinterpret_cast<const float4*>(sW3 + k * H + 32 * cb + 4 * c4);
              u[k] = fmaf(w3.x, h2[0], u[k]); u[k] = fmaf(w3.y, h2[1], u[k]);
              u[k] = fmaf(w3.z, h2[2], u[k]); u[k] = fmaf(w3.w, h2[3], u[k]);
            }
            r[4 * c4] = __float_as_uint(h2[0]); r[4 * c4 + 1] = __float_as_uint(h2[1]);
            r[4 * c4 + 2] = __float_as_uint(h2[2]); r[4 * c4 + 3] = __float_as_uint(h2[3]);
          }
          st32(ta, r);
        }
#pragma unroll
        for (int k = 0; k < D; ++k) s_u[half][tr][k] = u[k];
        bar_owners();
        UB_T(3);

        // ---- a_j (K2's formula; zero for rows that are not on a trajectory this pass)
        float a[D], dB[D];
        nc.get(A, inject, live, traj, j, dB);
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float v = 0.f;
          if (live) {
            const bool incl = s_exact ? (j < kstar) : true;
            const float ui = (s_u[0][tr][i] + s_u[1][tr][i]) + b3[i];
            v = (j < kstar ? ui * A.dt_f : 0.f) - (incl ? Gk * dB[i] : 0.f) + A.sigma_f * A.dt_f * lam[i];
          }
          a[i] = v;
          if (half == 0) gb3[i] += v;
        }

        // ---- sweep 2: dz2 = (W3^T a)(1 - h2^2) -> A ring (MMA 2) and exchange A planes, scaled by 2^esh; db2, dW3
#pragma unroll 1
        for (int ci = 0; ci < H / 64; ++ci) {
          const int cb = 2 * ci + half;
          uint32_t r[32];
          ld32(tmemZ + tlane + (uint32_t)(32 * cb), r);
          float dz[32];
#pragma unroll
          for (int cI = 0; cI < 32; ++cI) {
            const float hv = __uint_as_float(r[cI]);
            float dh = 0.f;
#pragma unroll
            for (int i = 0; i < D; ++i) dh = fmaf(a[i], sW3[i * H + 32 * cb + cI], dh);
            dz[cI] = dh * fmaf(-hv, hv, 1.0f);
          }
#pragma unroll
          for (int hs = 0; hs < 2; ++hs) {
            const int ks = 2 * cb + hs;
            const unsigned c = kc + KSTEPS + ks;
            const int slot = c % UB_ASLOTS;
            mbar_wait(&a_empty[slot], ((c / UB_ASLOTS) - 1u) & 1u);
            uint8_t* dst = sA + (size_t)slot * UMMA_ACHUNK + row_off;
#pragma unroll
            for (int kc2 = 0; kc2 < 2; ++kc2) {
              const int ub = 2 * ks + kc2;
              const int o = 16 * hs + 8 * kc2;
              uint4 hi, lo;
              split_pair(dz[o] * dscale, dz[o + 1] * dscale, hi.x, lo.x);
              split_pair(dz[o + 2] * dscale, dz[o + 3] * dscale, hi.y, lo.y);
              split_pair(dz[o + 4] * dscale, dz[o + 5] * dscale, hi.z, lo.z);
              split_pair(dz[o + 6] * dscale, dz[o + 7] * dscale, hi.w, lo.w);
              const uint32_t off = (uint32_t)kc2 * (UMMA_M / 8) * 128;
              *reinterpret_cast<uint4*>(dst + off) = hi;
              *reinterpret_cast<uint4*>(dst + UMMA_ACHUNK / 2 + off) = lo;
              *reinterpret_cast<uint4*>(xs + (uint32_t)ub * 256u) = hi;
              *reinterpret_cast<uint4*>(xs + PLANE + (uint32_t)ub * 256u) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&a_full[slot]);
          }
          // column sums over the warp's 32 trajectories
#pragma unroll
          for (int i = 0; i < D; ++i) {
            float v[32];
#pragma unroll
            for (int cI = 0; cI < 32; ++cI) v[cI] = a[i] * __uint_as_float(r[cI]);
            myAcc[(2 + D + i) * H + 32 * cb + lane] += lane_transpose_sum(v, lane);       // dW3[i][unit]
          }
          myAcc[1 * H + 32 * cb + lane] += lane_transpose_sum(dz, lane);                   // db2
        }

        UB_T(4);
        // ---- publish the pass to the consumer (which reads it with bulk copies: async proxy)
        // (every owner's stores are ordered before the barrier; the publishing thread's fence + release store then covers
        // them by cumulativity -- the usual "barrier, one thread fences and flags" pattern)
        asm volatile("fence.proxy.async.global;" ::: "memory");
        bar_owners();
        if (tid == 0) { __threadfence(); st_release(pc, pcount + 1); }

        // ---- sweep 3: dz1 = dh1 (1 - h1^2) / s, dX = dz1 W1 / s; db1, dW1
        UB_T(5);
        mbar_wait(&accG, par);
        UB_T(6);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float dx[D];
#pragma unroll
        for (int i = 0; i < D; ++i) dx[i] = 0.f;
        const float un = inv_s * dunscale;
#pragma unroll 1
        for (int ci = 0; ci < H / 64; ++ci) {
          const int cb = 2 * ci + half;
          uint32_t r[32];
          ld32(tmemG + tlane + (uint32_t)(32 * cb), r);
          float dz[32];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int ub = 4 * cb + q;
            const uint4 hi = *reinterpret_cast<const uint4*>(xs + 2 * PLANE + (uint32_t)ub * 256u);     // this thread's own h1
            const uint4 lo = *reinterpret_cast<const uint4*>(xs + 3 * PLANE + (uint32_t)ub * 256u);
            const uint32_t hw[4] = {hi.x, hi.y, hi.z, hi.w}, lw[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 fh = __half22float2(*reinterpret_cast<const __half2*>(&hw[e]));
              const float2 fl = __half22float2(*reinterpret_cast<const __half2*>(&lw[e]));
              const float h0 = fh.x + fl.x, h1v = fh.y + fl.y;
              dz[8 * q + 2 * e] = __uint_as_float(r[8 * q + 2 * e]) * un * fmaf(-h0, h0, 1.0f);
              dz[8 * q + 2 * e + 1] = __uint_as_float(r[8 * q + 2 * e + 1]) * un * fmaf(-h1v, h1v, 1.0f);
            }
          }
#pragma unroll
          for (int i = 0; i < D; ++i) {
            float v[32];
#pragma unroll
            for (int cI = 0; cI < 32; ++cI) {
              dx[i] = fmaf(sW1[i * H + 32 * cb + cI], dz[cI], dx[i]);
              v[cI] = dz[cI] * xf[i];
            }
            myAcc[(2 + i) * H + 32 * cb + lane] += lane_transpose_sum(v, lane);             // dW1[unit][i]
          }
          myAcc[0 * H + 32 * cb + lane] += lane_transpose_sum(dz, lane);                     // db1
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");     // the next pass's MMAs overwrite both accumulators
#pragma unroll
        for (int i = 0; i < D; ++i) s_dx[half][tr][i] = dx[i];                // met at the top of the next pass
      }
      // ---- tile done: small blocks to the float64 partial (row groups in index order)
      if (half == 0) {
#pragma unroll
        for (int i = 0; i < D; ++i) {
          float v = gb3[i];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
          if (lane == 0) s_b3[rgrp][i] = v;
        }
      }
      bar_owners();
      for (int e = tid; e < NQ * H; e += 2 * UMMA_M) {
        const float v = ((sAcc[e] + sAcc[(size_t)NQ * H + e]) + sAcc[(size_t)2 * NQ * H + e]) + sAcc[(size_t)3 * NQ * H + e];
        mySmall[e] += (double)v;
        sAcc[e] = 0.f; sAcc[(size_t)NQ * H + e] = 0.f; sAcc[(size_t)2 * NQ * H + e] = 0.f; sAcc[(size_t)3 * NQ * H + e] = 0.f;
      }
      if (tid < D) mySmall[NQ * H + tid] += (double)(((s_b3[0][tid] + s_b3[1][tid]) + s_b3[2][tid]) + s_b3[3][tid]);
    } else if (warp == 8) {
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(UMMA_M, H);
        unsigned kk = kc, pp = pcount;
#pragma unroll 1
        for (int ps = 0; ps < n_pass; ++ps, ++pp) {
#pragma unroll 1
          for (int prod = 0; prod < 2; ++prod) {
            const uint32_t dt = prod == 0 ? tmemZ : tmemG;
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks, ++kk) {
              const int s = kk % UMMA_STAGES, slot = kk % UB_ASLOTS;
              mbar_wait(&a_full[slot], (kk / UB_ASLOTS) & 1u);
              mbar_wait(&full[s], (kk / UMMA_STAGES) & 1u);
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
              const uint32_t a_base = smem_u32(sA) + (uint32_t)slot * UMMA_ACHUNK;
              const uint64_t ah = make_desc(a_base, (UMMA_M / 8) * 128, 128);
              const uint64_t al = make_desc(a_base + UMMA_ACHUNK / 2, (UMMA_M / 8) * 128, 128);
              const uint64_t bh = make_desc(smem_u32(sB) + (uint32_t)s * CHUNK, (H / 8) * 128, 128);
              const uint64_t bl = make_desc(smem_u32(sB) + (uint32_t)s * CHUNK + CHUNK / 2, (H / 8) * 128, 128);
              mma_f16(dt, ah, bh, idesc, ks > 0 ? 1u : 0u);
              mma_f16(dt, al, bh, idesc, 1u);
              mma_f16(dt, ah, bl, idesc, 1u);
              commit(&a_empty[slot]);
              commit(&empty[s]);
            }
            commit(prod == 0 ? &accZ : &accG);
          }
        }
      }
      kc += (unsigned)n_pass * 2 * KSTEPS;
      pcount += (unsigned)n_pass;
    } else {
      if (lane == 0) {
        unsigned kk = kc;
#pragma unroll 1
        for (int ps = 0; ps < n_pass; ++ps) {
#pragma unroll 1
          for (int prod = 0; prod < 2; ++prod) {
            const uint8_t* img = prod == 0 ? Bimg1 : Bimg2;
#pragma unroll 1
            for (int ks = 0; ks < KSTEPS; ++ks, ++kk) {
              const int s = kk % UMMA_STAGES;
              if (kk >= UMMA_STAGES) mbar_wait(&empty[s], ((kk / UMMA_STAGES) - 1u) & 1u);
              mbar_expect_tx(&full[s], CHUNK);
              bulk_g2s(sB + (size_t)s * CHUNK, img + (size_t)ks * CHUNK, CHUNK, &full[s]);
            }
          }
        }
      }
      kc += (unsigned)n_pass * 2 * KSTEPS;
      pcount += (unsigned)n_pass;
    }
    __syncthreads();                    // s_np is rewritten by the next tile
  }
  // ---- no more passes from this producer
#ifdef UB_PROFILE
  if (tid == 0) for (int i = 0; i < 8; ++i) pc[4 + i] = (unsigned)(prof[i] >> 8);
#endif
  if (tid == 0) { __threadfence(); st_release(pc + 2, 1u); }
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
}

// The launch's scale exponent: dz2 enters the tensor cores as float16 hi / lo of dz2 2^esh.  |dz2| <= |a|_1 max |W3| and
// a_j = u_j dt - G dB_j + sigma dt lambda_j is dominated by G dB: |a|_1 <~ D (max |G| 8 sqrt(dt) + 1).  esh leaves three binades
// of head-room above that bound (float16 tops out at 2^16; below it the hi / lo pair keeps an absolute 2^-25, so head-room
// costs nothing until ~20 binades).  One block; written to word 12 of every producer's control record.
static __global__ void ub_scale_kernel(const float* __restrict__ G, const int* __restrict__ T, long long K, float w3max, float sqrt_dt,
                                       int D, int n_prod, unsigned* __restrict__ ctl) {
  __shared__ float s_max[32];
  float m = 0.f;
  for (long long i = threadIdx.x; i < K; i += blockDim.x) if (T[i] >= 0) m = fmaxf(m, fabsf(G[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? s_max[threadIdx.x] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float bound = w3max * (float)D * (m * 8.0f * sqrt_dt + 1.0f);
    int e = 0;
    if (bound > 0.f && isfinite(bound)) frexpf(bound, &e);          // bound < 2^e
    int esh = 15 - 3 - e;
    esh = esh > 60 ? 60 : (esh < -60 ? -60 : esh);
    for (int p = threadIdx.x; p < n_prod; p += 32) ctl[(size_t)p * UB_CTL_WORDS + 12] = (unsigned)esh;
  }
}

// grad[...] (+)= scale * (partials in index order); the H x H block carries the 2^-esh unscale
template <int D, int H>
static __global__ void ub_reduce_kernel(const double* __restrict__ partial, int n_prod, int n_cons, float scale,
                                        const unsigned* __restrict__ ctl, float* __restrict__ grad, int accumulate) {
  const int esh = (int)ctl[12];
  constexpr int NQ = ub_nq<D>();
  constexpr int P = D * H + H + H * H + H + H * D + D;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= P) return;
  // state_dict order: W1 (H, D), b1 (H), W2 (H, H), b2 (H), W3 (D, H), b3 (D)
  const int oW1 = 0, ob1 = oW1 + H * D, oW2 = ob1 + H, ob2 = oW2 + H * H, oW3 = ob2 + H, ob3 = oW3 + D * H;
  double acc = 0.0;
  if (e >= oW2 && e < ob2) {
    for (int c = 0; c < n_cons; ++c) acc += partial[(size_t)c * H * H + (e - oW2)];
    acc = ldexp(acc, -esh);
  } else {
    int idx;
    if (e < ob1) { const int unit = e / D, i = e % D; idx = (2 + i) * H + unit; }
    else if (e < oW2) idx = 0 * H + (e - ob1);
    else if (e < oW3) idx = 1 * H + (e - ob2);
    else if (e < ob3) { const int i = (e - oW3) / H, unit = (e - oW3) % H; idx = (2 + D + i) * H + unit; }
    else idx = NQ * H + (e - ob3);
    const double* base = partial + (size_t)n_cons * H * H;
    for (int p = 0; p < n_prod; ++p) acc += base[(size_t)p * ub_small_count<D, H>() + idx];
  }
  const float g = (float)(acc * (double)scale);
  grad[e] = accumulate ? grad[e] + g : g;
}

// scratch the launcher needs behind the fixed workspace (exchange ring, control words, float64 partials, second image)
template <int D, int H>
size_t bwd_umma_scratch_bytes(int sm_count);
template <int D, int H>
int launch_rollout_bwd_umma(const float* params_host, float* params_dev, uint8_t* image_dev, const FwdArgs& args, float scale,
                            float* grad, uint8_t* scratch, size_t scratch_bytes, int sm_count, cudaStream_t stream);

}  // namespace rlsde
