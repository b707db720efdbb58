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
// columns (coalesced 512-byte reads per warp and s') and each thread walks s' with several independent 16-byte
// loads in flight.  The s' axis is additionally split over blockIdx.y (partial sums combined in a fixed order by
// the epilogue kernel, so results are deterministic).
#include "aux_kernels.cuh"

namespace rlsde {

// Work decomposition.  The tensor is one flat array of Ns' * cols doubles (cols = Ns * Na); row s' starts at flat
// offset s' * cols.  A thread owns one 16-byte-aligned PAIR of flat elements per row and reads it with one LDG.128.
// When cols is odd (config 3: 401 * 601) the rows alternate between two alignments: on rows with an even offset
// the pair is columns (2t, 2t+1), on rows with an odd offset it is (2t+1, 2t+2).  The two alignments accumulate
// into separate partial arrays ("even" / "odd"), which the epilogue adds in a fixed order; the single column each
// alignment cannot pair (the last one on even rows, the first one on odd rows) is walked with scalar loads by the
// one thread just past the last pair.  The s' axis is cut into n_split chunks (blockIdx.y) so that the grid fills
// the resident block slots of the GPU almost exactly (see pick_split): equal-sized items, no ragged last wave.
// Each thread keeps SWEEP_ROWS rows (16 B each) in flight.
#ifndef SWEEP_ROWS_N
#define SWEEP_ROWS_N 4
#endif
#ifndef SWEEP_THREADS_N
#define SWEEP_THREADS_N 128
#endif
constexpr int SWEEP_ROWS = SWEEP_ROWS_N;           // rows per unrolled step (even)
constexpr int SWEEP_THREADS = SWEEP_THREADS_N;
constexpr int SWEEP_SPLIT_MAX = 16;

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 v;
  asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}

// partial layout: [split][2 (alignment)][cols]
template <bool ODD_COLS>
__global__ void __launch_bounds__(SWEEP_THREADS) dp_sweep_partial_kernel(const double* __restrict__ P, long long n_sp,
                                                                         long long cols, const double* __restrict__ v,
                                                                         long long rows_per_split,
                                                                         double* __restrict__ partial) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long n_pairs = ODD_COLS ? (cols - 1) / 2 : cols / 2;
  const long long lo = (long long)blockIdx.y * rows_per_split;                    // even by construction
  const long long hi = lo + rows_per_split < n_sp ? lo + rows_per_split : n_sp;
  double* out_even = partial + (long long)blockIdx.y * 2 * cols;
  double* out_odd = out_even + cols;
  if (t > n_pairs || (t == n_pairs && !ODD_COLS)) return;
  if (t == n_pairs) {
    // the unpaired columns: cols - 1 on even-offset rows, 0 on odd-offset rows
    double se = 0.0, so = 0.0;
    for (long long sp = lo; sp < hi; ++sp) {
      if (sp & 1) so = fma(__ldcs(P + sp * cols), __ldg(v + sp), so);
      else se = fma(__ldcs(P + sp * cols + cols - 1), __ldg(v + sp), se);
    }
    out_even[cols - 1] = se;
    out_odd[0] = so;
    return;
  }
  double e0 = 0.0, e1 = 0.0, o0 = 0.0, o1 = 0.0;
  long long sp = lo;
  // flat element index of this thread's pair on row sp:  sp * cols + 2t + (row offset parity)
  for (; sp + SWEEP_ROWS <= hi; sp += SWEEP_ROWS) {
    double2 x[SWEEP_ROWS];
#pragma unroll
    for (int u = 0; u < SWEEP_ROWS; ++u) {
      const long long row = sp + u;
      const long long shift = ODD_COLS ? (row & 1) : 0;      // lo is even, so the parity of u is the parity of the row
      x[u] = ld_stream2(P + row * cols + 2 * t + shift);
    }
#pragma unroll
    for (int u = 0; u < SWEEP_ROWS; ++u) {
      const double w = __ldg(v + sp + u);
      if (ODD_COLS && (u & 1)) { o0 = fma(x[u].x, w, o0); o1 = fma(x[u].y, w, o1); }
      else { e0 = fma(x[u].x, w, e0); e1 = fma(x[u].y, w, e1); }
    }
  }
  for (; sp < hi; ++sp) {
    const long long shift = ODD_COLS ? (sp & 1) : 0;
    const double2 x = ld_stream2(P + sp * cols + 2 * t + shift);
    const double w = __ldg(v + sp);
    if (shift) { o0 = fma(x.x, w, o0); o1 = fma(x.y, w, o1); }
    else { e0 = fma(x.x, w, e0); e1 = fma(x.y, w, e1); }
  }
  *reinterpret_cast<double2*>(out_even + 2 * t) = make_double2(e0, e1);
  if (ODD_COLS) { out_odd[2 * t + 1] = o0; out_odd[2 * t + 2] = o1; }
}

// values[s, a] = R[s, a] + (1 - d[s]) gamma * (sum of the partials: split order, even alignment before odd)
__global__ void dp_sweep_epilogue_kernel(const double* __restrict__ partial, int n_split, int n_align, long long Ns,
                                         long long Na, const double* __restrict__ R,
                                         const unsigned char* __restrict__ in_ts, double gamma,
                                         double* __restrict__ values) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long cols = Ns * Na;
  if (c >= cols) return;
  double w = 0.0;
  for (int k = 0; k < n_split; ++k) {
    w += partial[(long long)k * 2 * cols + c];
    if (n_align == 2) w += partial[(long long)k * 2 * cols + cols + c];
  }
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

// number of s' chunks: the one whose block count fills the resident slots (SMs x blocks per SM) most evenly
static int pick_split(long long col_blocks, long long n_sp, int slots) {
  int best = 1;
  double best_eff = 0.0;
  for (int sp = 1; sp <= SWEEP_SPLIT_MAX; ++sp) {
    long long rows = (n_sp + sp - 1) / sp;
    rows += rows & 1;                                     // chunks start on even rows
    if (rows < 2 * SWEEP_ROWS && sp > 1) break;
    const long long used = (n_sp + rows - 1) / rows;      // chunks that actually hold rows
    if (used != sp) continue;
    const long long items = col_blocks * sp;
    const long long waves = (items + slots - 1) / slots;
    const double eff = (double)items / (double)(waves * slots) - 0.004 * sp;   // small penalty: partials written per split
    if (eff > best_eff) { best_eff = eff; best = sp; }
  }
  return best;
}

int launch_dp_sweep(const double* P, long long Ns, long long Na, const double* R, const unsigned char* in_ts,
                    const double* v, double gamma, double* values, double* scratch, cudaStream_t stream) {
  const long long cols = Ns * Na;
  const bool odd = (cols & 1) != 0;
  const long long n_threads = (odd ? (cols - 1) / 2 + 1 : cols / 2);
  const long long col_blocks = (n_threads + SWEEP_THREADS - 1) / SWEEP_THREADS;
  int sm = 148, per_sm = 8;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
  if (odd) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_sweep_partial_kernel<true>, SWEEP_THREADS, 0);
  else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, dp_sweep_partial_kernel<false>, SWEEP_THREADS, 0);
  if (per_sm < 1) per_sm = 1;
  const int n_split = pick_split(col_blocks, Ns, sm * per_sm);
  long long rows = (Ns + n_split - 1) / n_split;
  rows += rows & 1;
  dim3 grid((unsigned)col_blocks, (unsigned)n_split);
  if (odd) dp_sweep_partial_kernel<true><<<grid, SWEEP_THREADS, 0, stream>>>(P, Ns, cols, v, rows, scratch);
  else dp_sweep_partial_kernel<false><<<grid, SWEEP_THREADS, 0, stream>>>(P, Ns, cols, v, rows, scratch);
  dp_sweep_epilogue_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>(scratch, n_split, odd ? 2 : 1, Ns, Na, R, in_ts,
                                                                                gamma, values);
  note_kernel_launches(2);
  return (int)cudaGetLastError();
}

size_t dp_sweep_scratch_bytes(long long Ns, long long Na) { return (size_t)SWEEP_SPLIT_MAX * 2 * Ns * Na * sizeof(double); }

int launch_dp_rowmax(const double* values, long long Ns, long long Na, double* vmax, long long* arg, cudaStream_t stream) {
  dp_rowmax_kernel<<<(unsigned)((Ns + 7) / 8), 256, 0, stream>>>(values, Ns, Na, vmax, arg);
  note_kernel_launches(1);
  return (int)cudaGetLastError();
}

}  // namespace rlsde
