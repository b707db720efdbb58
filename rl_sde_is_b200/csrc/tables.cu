// K4: tabular transition tensor P[s', s, a] and reward table R[s, a]  (float64).
//
// Replaces compute_p_tensor_batch / compute_r_table (dynamic_programming.py:3-36), i.e. 180 300
// Python iterations x 4 scipy.stats.norm.cdf calls at h = 0.01, and the per-column arithmetic of
// state_action_transition_function (environments.py:87-102):
//     mu = x_s + (-gradV(x_s) + sigma a) dt,  sd = sigma sqrt(dt)
//     P[s', s, a] = Phi((x_s' + h - mu) / sd) - Phi((x_s' - h - mu) / sd)       s outside the target set
//     P[0,   s, a] += Phi((x_0 - h - mu) / sd);   P[N-1, s, a] += 1 - Phi((x_{N-1} + h - mu) / sd)
//     P[s', s, a] = 1/|TS| if s' in TS else 0                                     s in the target set
//
// Layout / mapping: the tensor is C-ordered with the action index innermost, so a warp's 32 lanes
// take 32 consecutive actions (each store instruction writes 256 contiguous bytes) and every thread
// walks TILE consecutive next-states, carrying the CDF of the shared cell edge from one cell to the
// next (TILE+1 CDF evaluations per TILE entries instead of 2 TILE).  The kernel is bound by the FP64
// pipe (erf/erfc), not by HBM: see DESIGN.md.
#include "aux_kernels.cuh"

namespace rlsde {

// scipy.special.ndtr (cephes ndtr.c) branch structure, with CUDA's erf / erfc
__device__ __forceinline__ double ndtr(double a) {
  const double kSqrtH = 0.70710678118654752440;
  const double x = a * kSqrtH;
  const double z = fabs(x);
  double y;
  if (z < kSqrtH) {
    y = 0.5 + 0.5 * erf(x);
  } else {
    y = 0.5 * erfc(z);
    if (x > 0) y = 1.0 - y;
  }
  return y;
}

constexpr int TABLE_TILE = 8;

__global__ void __launch_bounds__(128) tables_kernel(const double* __restrict__ sgrid, long long Ns,
                                                     const double* __restrict__ agrid, long long Na,
                                                     const unsigned char* __restrict__ in_ts, double inv_nts,
                                                     double alpha, double sigma, double dt, double h,
                                                     long long sp_begin, long long sp_end, double* __restrict__ P) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = blockIdx.y;
  const long long sp0 = sp_begin + (long long)blockIdx.z * TABLE_TILE;
  if (a >= Na || sp0 >= sp_end) return;
  const long long sp1 = sp0 + TABLE_TILE < sp_end ? sp0 + TABLE_TILE : sp_end;
  double* out = P + ((sp0 - sp_begin) * Ns + s) * Na + a;
  const long long stride = Ns * Na;
  if (in_ts[s]) {
    for (long long sp = sp0; sp < sp1; ++sp, out += stride) *out = in_ts[sp] ? inv_nts : 0.0;
    return;
  }
  const double xs = sgrid[s];
  const double act = agrid[a];
  // mu = state + (-gradient(state) + sigma * action) * dt,  gradient = 4 alpha x (x^2 - 1)
  const double grad = __dmul_rn(__dmul_rn(__dmul_rn(4.0, alpha), xs), __dsub_rn(__dmul_rn(xs, xs), 1.0));
  const double mu = __dadd_rn(xs, __dmul_rn(__dadd_rn(-grad, __dmul_rn(sigma, act)), dt));
  const double sd = __dmul_rn(sigma, sqrt(dt));
  double lo = ndtr(__ddiv_rn(__dsub_rn(__dsub_rn(sgrid[sp0], h), mu), sd));
  for (long long sp = sp0; sp < sp1; ++sp, out += stride) {
    // upper edge of cell sp = lower edge of cell sp+1 (x_{sp+1} - h), so each cell boundary has ONE value
    // whatever the tiling / slab; only the last cell of the grid uses its own x_sp + h
    const double edge = (sp + 1 < Ns) ? __dsub_rn(sgrid[sp + 1], h) : __dadd_rn(sgrid[sp], h);
    const double hi = ndtr(__ddiv_rn(__dsub_rn(edge, mu), sd));
    double p = hi - lo;
    if (sp == 0) p += lo;                 // left tail folded into the first row
    if (sp == Ns - 1) p += 1.0 - hi;      // right tail folded into the last row
    *out = p;
    lo = hi;
  }
}

// ---- fast path: uniform state grid with fine cells (cell width <= 0.25 sd) -------------------------------
// P_cell = integral of the N(0,1) density over [t, t + dcell]: 4-point Gauss-Legendre (error < 1e-13 at
// dcell = 0.25) with the density at the nodes advanced from cell to cell by the multiplicative recurrence
//     phi(t + ds) = phi(t) * r,   r(t) = exp(-t ds - ds^2 / 2),   r(t + ds) = r(t) * exp(-ds^2)
// and re-anchored with exact exp() every GL_ANCHOR cells (measured |P - reference formula| <= 5e-15 at config 3,
// tests/test_gpu_parity.py holds 1e-13).  ~14 FP64 operations per entry instead of ~100 for erf/erfc, which moves
// the kernel from the FP64 pipe towards the HBM write roofline.  One thread owns one (s, a) column and walks s'.
constexpr int GL_ANCHOR = 64;

__global__ void __launch_bounds__(128) tables_gl_kernel(const double* __restrict__ sgrid, long long Ns,
                                                        const double* __restrict__ agrid, long long Na,
                                                        const unsigned char* __restrict__ in_ts, double inv_nts,
                                                        double alpha, double sigma, double dt, double h,
                                                        long long sp_begin, long long sp_end, double* __restrict__ P) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long s = blockIdx.y;
  if (a >= Na) return;
  // blockIdx.z picks one anchor-aligned chunk of GL_ANCHOR next-states: many equal work items instead of one
  // long column per thread, so the grid spreads evenly over the SMs (2005 long blocks ran as 1.4 waves)
  const long long chunk_lo = (sp_begin / GL_ANCHOR + blockIdx.z) * GL_ANCHOR;
  const long long slab_begin = sp_begin;
  sp_begin = chunk_lo > sp_begin ? chunk_lo : sp_begin;
  sp_end = chunk_lo + GL_ANCHOR < sp_end ? chunk_lo + GL_ANCHOR : sp_end;
  if (sp_begin >= sp_end) return;
  const long long stride = Ns * Na;
  double* out = P + (sp_begin - slab_begin) * stride + s * Na + a;     // P points at row slab_begin
  if (in_ts[s]) {
    for (long long sp = sp_begin; sp < sp_end; ++sp, out += stride) *out = in_ts[sp] ? inv_nts : 0.0;
    return;
  }
  const double xs = sgrid[s];
  const double act = agrid[a];
  const double grad = __dmul_rn(__dmul_rn(__dmul_rn(4.0, alpha), xs), __dsub_rn(__dmul_rn(xs, xs), 1.0));
  const double mu = __dadd_rn(xs, __dmul_rn(__dadd_rn(-grad, __dmul_rn(sigma, act)), dt));
  const double sd = __dmul_rn(sigma, sqrt(dt));
  const double inv_sd = 1.0 / sd;
  // anchors sit at absolute multiples of GL_ANCHOR, so a slab holds exactly the entries of the full tensor
  const long long sp_first = (sp_begin / GL_ANCHOR) * GL_ANCHOR;
  const double dcell = 2.0 * h * inv_sd;                                   // integration width of a cell
  const double dstep = (sgrid[Ns - 1] - sgrid[0]) / (double)(Ns - 1) * inv_sd;   // advance between cells
  const double rho = exp(-dstep * dstep);
  const double c0 = 0.5 - 0.4305681557970263, c1 = 0.5 - 0.1699905217924281;
  const double c2 = 0.5 + 0.1699905217924281, c3 = 0.5 + 0.4305681557970263;
  const double w0 = 0.1739274225687269 * dcell, w1 = 0.3260725774312731 * dcell;
  const double kInvSqrt2Pi = 0.3989422804014327;
  double p0 = 0, p1 = 0, p2 = 0, p3 = 0, r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  for (long long sp = sp_first; sp < sp_end; ++sp) {
    if ((sp % GL_ANCHOR) == 0) {
      const double tl = (sgrid[sp] - h - mu) * inv_sd;
      const double t0 = fma(c0, dcell, tl), t1 = fma(c1, dcell, tl), t2 = fma(c2, dcell, tl), t3 = fma(c3, dcell, tl);
      const double hs = -0.5 * dstep * dstep;
      p0 = kInvSqrt2Pi * exp(-0.5 * t0 * t0); r0 = exp(fma(-t0, dstep, hs));
      p1 = kInvSqrt2Pi * exp(-0.5 * t1 * t1); r1 = exp(fma(-t1, dstep, hs));
      p2 = kInvSqrt2Pi * exp(-0.5 * t2 * t2); r2 = exp(fma(-t2, dstep, hs));
      p3 = kInvSqrt2Pi * exp(-0.5 * t3 * t3); r3 = exp(fma(-t3, dstep, hs));
    }
    if (sp >= sp_begin) {
      double p = fma(w0, p0 + p3, w1 * (p1 + p2));
      if (sp == 0) p += ndtr((sgrid[0] - h - mu) * inv_sd);                   // left tail folded into the first row
      if (sp == Ns - 1) p += 1.0 - ndtr((sgrid[Ns - 1] + h - mu) * inv_sd);   // right tail folded into the last row
      *out = p;
      out += stride;
    }
    p0 *= r0; p1 *= r1; p2 *= r2; p3 *= r3;
    r0 *= rho; r1 *= rho; r2 *= rho; r3 *= rho;
  }
}

__global__ void rtable_kernel(long long Ns, const double* __restrict__ agrid, long long Na,
                              const unsigned char* __restrict__ in_ts, double dt, double* __restrict__ R) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Ns * Na) return;
  const long long s = e / Na, a = e % Na;
  const double act = agrid[a];
  // where(done, -g(x) = -0.0, -(f + 0.5 |a|^2) dt)   environments.py:104-110
  R[e] = in_ts[s] ? -0.0 : -__dmul_rn(__dadd_rn(1.0, __dmul_rn(0.5, __dmul_rn(act, act))), dt);
}

int launch_tables(const double* state_grid, long long Ns, const double* action_grid, long long Na,
                  const unsigned char* in_ts, long long n_ts, double alpha, double sigma, double dt, double h_half,
                  double lb, double rb, long long sprime_begin, long long sprime_end, double* P, double* R,
                  int uniform_grid, cudaStream_t stream) {
  (void)lb; (void)rb;
  const double dcell = 2.0 * h_half / (sigma * sqrt(dt));
  if (P && sprime_end > sprime_begin && uniform_grid && Ns >= 2 && dcell <= 0.25) {
    const long long n_chunks = (sprime_end - 1) / GL_ANCHOR - sprime_begin / GL_ANCHOR + 1;
    dim3 grid((unsigned)((Na + 127) / 128), (unsigned)Ns, (unsigned)n_chunks);
    tables_gl_kernel<<<grid, 128, 0, stream>>>(state_grid, Ns, action_grid, Na, in_ts, n_ts > 0 ? 1.0 / (double)n_ts : 0.0,
                                               alpha, sigma, dt, h_half, sprime_begin, sprime_end,
                                               P);
  } else if (P && sprime_end > sprime_begin) {
    const long long nsp = sprime_end - sprime_begin;
    dim3 grid((unsigned)((Na + 127) / 128), (unsigned)Ns, (unsigned)((nsp + TABLE_TILE - 1) / TABLE_TILE));
    tables_kernel<<<grid, 128, 0, stream>>>(state_grid, Ns, action_grid, Na, in_ts, n_ts > 0 ? 1.0 / (double)n_ts : 0.0,
                                            alpha, sigma, dt, h_half, sprime_begin, sprime_end, P);
  }
  if (R) {
    const long long n = Ns * Na;
    rtable_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(Ns, action_grid, Na, in_ts, dt, R);
  }
  return (int)cudaGetLastError();
}

// colsum[s, a] += sum over the slab's next-states, in index order (deterministic)
__global__ void colsum_kernel(const double* __restrict__ P, long long n_sprime, long long cols, double* __restrict__ colsum) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double a = colsum[c];
  for (long long sp = 0; sp < n_sprime; ++sp) a += P[sp * cols + c];
  colsum[c] = a;
}

int launch_tables_colsum(const double* P, long long n_sprime, long long Ns, long long Na, double* colsum,
                         cudaStream_t stream) {
  const long long cols = Ns * Na;
  colsum_kernel<<<(unsigned)((cols + 127) / 128), 128, 0, stream>>>(P, n_sprime, cols, colsum);
  return (int)cudaGetLastError();
}

}  // namespace rlsde
