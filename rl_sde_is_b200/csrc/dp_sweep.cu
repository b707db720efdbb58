// K6: Bellman sweeps over the device-resident transition tensor (the consumer of K4's output).
//
// Replaces the contraction inside q_table_update_vect (tabular_dp_qvalue_iteration.py:35-43),
// v_table_update_vect (tabular_dp_value_iteration.py:41-52) and policy_update_vect
// (tabular_dp_policy_iteration.py:37-49):
//     values[s, a] = R[s, a] + (1 - d[s]) * gamma * sum_{s'} P[s', s, a] * v[s']
// which NumPy evaluates as one (Na, Ns, Ns') x (Ns',) matmul over the 773 MB tensor per sweep, followed by a
// max / argmax over actions.  The kernel is a streaming FP64 GEMV: HBM-read bound (8 Ns^2 Na bytes per sweep,
// 0.118 ms at the measured copy bandwidth); keeping P resident removes the tensor's host round trip entirely.
//
// Mapping: the tensor is [s'][s][a] with the action innermost, so consecutive lanes take consecutive (s, a)
// columns (coalesced 256-byte reads per warp and s') and each thread walks s' with UNROLL independent loads in
// flight.  The s' axis is additionally split over blockIdx.y (partial sums combined in a fixed order by the
// epilogue kernel, so results are deterministic).
#include "aux_kernels.cuh"

namespace rlsde {

constexpr int SWEEP_UNROLL = 8;
constexpr int SWEEP_SPLIT_MAX = 8;

__global__ void __launch_bounds__(256) dp_sweep_partial_kernel(const double* __restrict__ P, long long n_sp, long long cols,
                                                               const double* __restrict__ v, int n_split,
                                                               double* __restrict__ partial) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const long long per = (n_sp + n_split - 1) / n_split;
  const long long lo = (long long)blockIdx.y * per;
  const long long hi = lo + per < n_sp ? lo + per : n_sp;
  const double* p = P + lo * cols + c;
  double acc[SWEEP_UNROLL];
#pragma unroll
  for (int u = 0; u < SWEEP_UNROLL; ++u) acc[u] = 0.0;
  long long sp = lo;
  for (; sp + SWEEP_UNROLL <= hi; sp += SWEEP_UNROLL) {
    double x[SWEEP_UNROLL];
#pragma unroll
    for (int u = 0; u < SWEEP_UNROLL; ++u) x[u] = __ldcs(p + (long long)u * cols);     // streamed once per sweep
#pragma unroll
    for (int u = 0; u < SWEEP_UNROLL; ++u) acc[u] = fma(x[u], __ldg(v + sp + u), acc[u]);
    p += (long long)SWEEP_UNROLL * cols;
  }
  for (; sp < hi; ++sp, p += cols) acc[0] = fma(__ldcs(p), __ldg(v + sp), acc[0]);
  double s = 0.0;
#pragma unroll
  for (int u = 0; u < SWEEP_UNROLL; ++u) s += acc[u];
  partial[(long long)blockIdx.y * cols + c] = s;
}

// values[s, a] = R[s, a] + (1 - d[s]) gamma * (sum of the split partials, in split order)
__global__ void dp_sweep_epilogue_kernel(const double* __restrict__ partial, int n_split, long long Ns, long long Na,
                                         const double* __restrict__ R, const unsigned char* __restrict__ in_ts, double gamma,
                                         double* __restrict__ values) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cols = Ns * Na;
  if (c >= cols) return;
  double w = 0.0;
  for (int k = 0; k < n_split; ++k) w += partial[(long long)k * cols + c];
  const long long s = c / Na;
  const double live = in_ts[s] ? 0.0 : 1.0;
  values[c] = R[c] + live * gamma * w;
}

// v[s] = max_a values[s, a], arg[s] = first maximiser (np.max / np.argmax over axis 1); one warp per state
__global__ void dp_rowmax_kernel(const double* __restrict__ values, long long Ns, long long Na, double* __restrict__ vmax,
                                 long long* __restrict__ arg) {
  const long long s = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= Ns) return;
  double best = -INFINITY;
  long long bi = 0x7fffffffffffffffLL;
  for (long long a = lane; a < Na; a += 32) {
    const double x = values[s * Na + a];
    if (x > best) { best = x; bi = a; }          // strict: keeps the first maximiser within a lane
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) {
    if (vmax) vmax[s] = best;
    if (arg) arg[s] = bi;
  }
}

int launch_dp_sweep(const double* P, long long Ns, long long Na, const double* R, const unsigned char* in_ts,
                    const double* v, double gamma, double* values, double* scratch, cudaStream_t stream) {
  const long long cols = Ns * Na;
  // enough (column-block x split) blocks to cover the SMs a few times over
  int n_split = 1;
  while (n_split < SWEEP_SPLIT_MAX && ((cols + 255) / 256) * n_split < 148 * 8 && Ns / (n_split * 2) >= 4 * SWEEP_UNROLL) n_split *= 2;
  dim3 grid((unsigned)((cols + 255) / 256), (unsigned)n_split);
  dp_sweep_partial_kernel<<<grid, 256, 0, stream>>>(P, Ns, cols, v, n_split, scratch);
  dp_sweep_epilogue_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>(scratch, n_split, Ns, Na, R, in_ts, gamma, values);
  return (int)cudaGetLastError();
}

size_t dp_sweep_scratch_bytes(long long Ns, long long Na) { return (size_t)SWEEP_SPLIT_MAX * Ns * Na * sizeof(double); }

int launch_dp_rowmax(const double* values, long long Ns, long long Na, double* vmax, long long* arg, cudaStream_t stream) {
  dp_rowmax_kernel<<<(unsigned)((Ns + 7) / 8), 256, 0, stream>>>(values, Ns, Na, vmax, arg);
  return (int)cudaGetLastError();
}

}  // namespace rlsde
