// Device-side preparation of the cosmology input: what the reference redoes on the host for every
// cosmology of a forecast -- FITPACK fits of the tabulated functions, the comoving-distance integral, the
// Savitzky-Golay no-wiggle spectrum, the velocity moments -- as batched kernels over all cosmologies of a job.
//
//   cfp_bank_fit            RectBivariateSpline(z, k, table, kx=3, ky=3) of every (cosmology, table)
//                           cosmicfishpie/cosmology/cosmology.py:1463-1532
//   cfp_background_fit      InterpolatedUnivariateSpline of H, sigma8; comoving / angular distance
//                           cosmology.py:36-55, 1447-1475, 1500-1506
//   cfp_bank_eval_many      f(z, k), D(z, k) of every cosmology at given points      cosmology.py:1859-1949
//   cfp_bank_nowiggle       Savitzky-Golay smoothing of ln P(ln k)                   cosmology.py:1760-1813
//   cfp_bank_theta_moments  (1 / 6 pi^2) int P f^m dk                                LSSsurvey/spectro_obs.py:724-755
//
// An interpolating cubic spline on FITPACK's knots (the data abscissae without the second and the second-to-last:
// "not-a-knot") is unique, so the coefficients need not come from FITPACK's Givens rotations: the collocation matrix
// B_j(x_i) is tridiagonal except for its second and second-to-last rows (four entries), totally positive, and is
// eliminated here in natural order without pivoting.  The factorisation depends on the abscissae only and is shared
// by all right-hand sides of a grid; a surface is fitted by solving along z for every k column, then along k for
// every z row.  Measured against scipy on the toy tables: coefficients and evaluated values agree to 3e-15 relative.
#include <cmath>
#include <cstdarg>

#include "cfp_common.cuh"

namespace {

// ---------------------------------------------------------------------------------- B-spline pieces (any address space)
__device__ __forceinline__ void bspl3_any(const double* t, int l, double x, double h[4]) {
  double hh[3];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= 3; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 0; i < j; ++i) {
      const double tli = t[l + i + 1];
      const double tlj = t[l + i + 1 - j];
      if (tli == tlj) {
        h[i + 1] = 0.0;
      } else {  // (no FMA contraction: the order of fpbspl.f)
        const double f = __ddiv_rn(hh[i], __dsub_rn(tli, tlj));
        h[i] = __dadd_rn(h[i], __dmul_rn(f, __dsub_rn(tli, x)));
        h[i + 1] = __dmul_rn(f, __dsub_rn(x, tlj));
      }
    }
  }
}

__device__ __forceinline__ int find_interval_any(const double* t, int n, double x) {
  int lo = 3, hi = n - 5;  // splev.f: the end pieces extrapolate
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (t[mid] <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// splev with ext = 0 (InterpolatedUnivariateSpline.__call__): t has n + 4 knots, c n coefficients
__device__ __forceinline__ double splev_any(const double* t, int n, const double* c, double x) {
  const int l = find_interval_any(t, n + 4, x);
  double h[4];
  bspl3_any(t, l, x, h);
  double sp = 0.0;
#pragma unroll
  for (int j = 0; j < 4; ++j) sp = __dadd_rn(sp, __dmul_rn(c[l - 3 + j], h[j]));
  return sp;
}

// Knots and LU factors of the collocation matrix of the abscissae x[0..n): t[n + 4]; fac = L[n] | E[n] | R[n] | {h0, f1,
// m1, g3} with R = 1 / pivot.  Rows 2..n-3 sit on knots (three entries a, b, c on the unknowns i-1, i, i+1); row 1 and
// row n-2 lie inside the first and last knot interval (four entries).  c_0 and c_{n-1} are the end values themselves.
// Called by all threads of a block: the knots and the basis values of every row are computed in parallel (the divisions
// of the de Boor recurrence are the expensive part), the elimination itself -- one division per row -- by thread 0.
// `work` holds 3 n doubles (shared or global); all three arrays may be read after the closing barrier.
__device__ void nak_factor_block(const double* x, int n, double* t, double* fac, double* work) {
  double *L = fac, *E = fac + n, *R = fac + 2 * n, *misc = fac + 3 * n;
  double *ha = work, *hb = work + n, *hc = work + 2 * n;
  const int tid = threadIdx.x, nt = blockDim.x;
  for (int i = tid; i < n + 4; i += nt) t[i] = i < 4 ? x[0] : (i >= n ? x[n - 1] : x[i - 2]);
  __syncthreads();
  for (int i = tid + 2; i < n - 2; i += nt) {
    double h[4];
    bspl3_any(t, i + 2, x[i], h);
    ha[i] = h[0], hb[i] = h[1], hc[i] = h[2];
  }
  __syncthreads();
  if (tid == 0) {
    double h[4];
    bspl3_any(t, 3, x[1], h);
    const double h0 = h[0], f1 = h[3];
    L[0] = L[1] = E[0] = R[0] = 0.0;
    R[1] = 1.0 / h[1];
    E[1] = h[2];
    for (int i = 2; i < n - 2; ++i) {
      const double m = ha[i] * R[i - 1];
      L[i] = m;
      R[i] = 1.0 / (hb[i] - m * E[i - 1]);
      E[i] = i == 2 ? hc[i] - m * f1 : hc[i];
    }
    bspl3_any(t, n - 1, x[n - 2], h);
    const double m1 = h[0] * R[n - 4];
    const double g1 = h[1] - m1 * E[n - 4];
    const double g2 = n - 4 == 1 ? h[2] - m1 * f1 : h[2];
    const double m2 = g1 * R[n - 3];
    L[n - 2] = m2;
    R[n - 2] = 1.0 / (g2 - m2 * E[n - 3]);
    E[n - 2] = E[n - 1] = L[n - 1] = R[n - 1] = 0.0;
    misc[0] = h0, misc[1] = f1, misc[2] = m1, misc[3] = h[3];
  }
  __syncthreads();
}

// In-place solve of one right-hand side y[i * stride], scaled by `scale` on the way in.
__device__ __forceinline__ void nak_solve(const double* __restrict__ fac, int n, double* y, size_t stride, double scale) {
  const double *L = fac, *E = fac + n, *R = fac + 2 * n, *misc = fac + 3 * n;
  const double h0 = misc[0], f1 = misc[1], m1 = misc[2], g3 = misc[3];
  const double c0 = y[0] * scale, cl = y[(size_t)(n - 1) * stride] * scale;
  y[0] = c0;
  y[(size_t)(n - 1) * stride] = cl;
  double prev = y[stride] * scale - h0 * c0, prev2 = 0.0;  // r_{i-1}, r_{i-2}
  y[stride] = prev;
  for (int i = 2; i < n - 2; ++i) {
    const double r = y[(size_t)i * stride] * scale - L[i] * prev;
    y[(size_t)i * stride] = r;
    prev2 = prev, prev = r;
  }
  // row n-2: entries on c_{n-4}, c_{n-3}, c_{n-2}, c_{n-1}; prev = r_{n-3}, prev2 = r_{n-4}
  double cn = (y[(size_t)(n - 2) * stride] * scale - g3 * cl - m1 * prev2 - L[n - 2] * prev) * R[n - 2];
  y[(size_t)(n - 2) * stride] = cn;
  for (int i = n - 3; i >= 2; --i) {
    cn = (y[(size_t)i * stride] - E[i] * cn) * R[i];
    y[(size_t)i * stride] = cn;
  }
  y[stride] = (y[stride] - E[1] * y[2 * stride] - f1 * y[3 * stride]) * R[1];
}

// ---------------------------------------------------------------------------------- bicubic fit
// one block per grid: g < NC the k grid of cosmology g, g == NC the shared z grid; work[(NC + 1)][3 max(Nz, Nk)]
__global__ void __launch_bounds__(128) prep_factor_kernel(int NC, int Nz, int Nk, const double* __restrict__ z,
                                                          const double* __restrict__ k, double* tz, double* fz, double* tk,
                                                          double* fk, double* work) {
  const int g = blockIdx.x;
  double* w = work + (size_t)g * 3 * max(Nz, Nk);
  if (g == NC)
    nak_factor_block(z, Nz, tz, fz, w);
  else
    nak_factor_block(k + (size_t)g * Nk, Nk, tk + (size_t)g * (Nk + 4), fk + (size_t)g * (3 * Nk + 4), w);
}

// jobs[j] = {offset of the slot's tx in the blob, cosmology}; scale[j]
__global__ void prep_knots_kernel(int njobs, const int64_t* __restrict__ jobs, int Nz, int Nk, const double* __restrict__ tz,
                                  const double* __restrict__ tk, double* __restrict__ blob) {
  const int j = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs || i >= Nz + 4 + Nk + 4) return;
  const int64_t off = jobs[2 * j];
  const int c = (int)jobs[2 * j + 1];
  blob[off + i] = i < Nz + 4 ? tz[i] : tk[(size_t)c * (Nk + 4) + (i - Nz - 4)];
}

__global__ void prep_solve_z_kernel(int njobs, const int64_t* __restrict__ jobs, const double* __restrict__ scale, int Nz, int Nk,
                                    const double* __restrict__ fz, double* __restrict__ blob) {
  const int j = blockIdx.y;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= njobs || col >= Nk) return;
  double* cc = blob + jobs[2 * j] + (Nz + 4) + (Nk + 4);
  nak_solve(fz, Nz, cc + col, (size_t)Nk, scale[j]);
}

// Along k every row of a surface is one right-hand side.  A block stages `rpb` rows in shared memory with coalesced
// loads (row pitch odd: conflict-free), one thread per row runs the serial elimination there, and the block writes the
// coefficients back coalesced -- a thread walking its own row in global memory would wait on every element.
__global__ void __launch_bounds__(256) prep_solve_k_kernel(int njobs, const int64_t* __restrict__ jobs, int Nz, int Nk, int rpb,
                                                           int pitch, const double* __restrict__ fk, double* __restrict__ blob) {
  extern __shared__ double tile[];
  const int j = blockIdx.y, tid = threadIdx.x;
  const int row0 = blockIdx.x * rpb;
  if (j >= njobs || row0 >= Nz) return;
  const int nrows = min(rpb, Nz - row0);
  double* cc = blob + jobs[2 * j] + (Nz + 4) + (Nk + 4) + (size_t)row0 * Nk;
  const int c = (int)jobs[2 * j + 1];
  for (int idx = tid; idx < nrows * Nk; idx += blockDim.x) {
    const int r = idx / Nk;
    tile[r * pitch + (idx - r * Nk)] = cc[idx];
  }
  __syncthreads();
  if (tid < nrows) nak_solve(fk + (size_t)c * (3 * Nk + 4), Nk, tile + (size_t)tid * pitch, 1, 1.0);
  __syncthreads();
  for (int idx = tid; idx < nrows * Nk; idx += blockDim.x) {
    const int r = idx / Nk;
    cc[idx] = tile[r * pitch + (idx - r * Nk)];
  }
}

// ---------------------------------------------------------------------------------- background
// One block per cosmology.  Shared memory: t[Nz+4] | fac[3Nz+4] | coef[(3 + nx)][Nz] | work[3 Nz] | ybuf[warps][ntrap].
// coef rows: 0 H, 1 comoving, 2 angular, 3.. the extra functions of z (sigma8 ...).
__global__ void __launch_bounds__(512) prep_background_kernel(int Nz, const double* __restrict__ z, const double* __restrict__ hub,
                                                              int nx, const double* __restrict__ extra, int ntrap, int nq,
                                                              const double* __restrict__ zq, double* __restrict__ out,
                                                              double* __restrict__ coef_out) {
  extern __shared__ double sm[];
  double* t = sm;
  double* fac = t + (Nz + 4);
  double* coef = fac + (3 * Nz + 4);
  double* work = coef + (size_t)(3 + nx) * Nz;
  const int c = blockIdx.x, tid = threadIdx.x;
  for (int i = tid; i < Nz; i += blockDim.x) {
    coef[i] = hub[(size_t)c * Nz + i];
    for (int e = 0; e < nx; ++e) coef[(size_t)(3 + e) * Nz + i] = extra[((size_t)c * nx + e) * Nz + i];
  }
  __syncthreads();
  nak_factor_block(z, Nz, t, fac, work);
  if (tid == 0) nak_solve(fac, Nz, coef, 1, 1.0);
  if (tid >= 1 && tid <= nx) nak_solve(fac, Nz, coef + (size_t)(2 + tid) * Nz, 1, 1.0);
  __syncthreads();
  // comoving distance to every tabulated redshift: trapezoid of 1/H on numpy.linspace(0, z_i, ntrap); a warp per
  // redshift, the integrand samples through shared memory ybuf[warp][ntrap]
  {
    const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
    double* yb = work + (size_t)3 * Nz + (size_t)warp * ntrap;
    for (int i = warp; i < Nz; i += nw) {
      const double zi = z[i];
      const double step = zi / (double)(ntrap - 1);
      for (int s = lane; s < ntrap; s += 32) yb[s] = 1.0 / splev_any(t, Nz, coef, s == ntrap - 1 ? zi : (double)s * step);
      __syncwarp();
      double acc = 0.0;
      for (int s = lane + 1; s < ntrap; s += 32) {
        const double x0 = (double)(s - 1) * step, x1 = s == ntrap - 1 ? zi : (double)s * step;
        acc += (x1 - x0) * (yb[s] + yb[s - 1]) / 2.0;
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0) {
        coef[(size_t)Nz + i] = acc;
        coef[(size_t)2 * Nz + i] = acc / (1.0 + zi);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (tid < 2) nak_solve(fac, Nz, coef + (size_t)(1 + tid) * Nz, 1, 1.0);
  __syncthreads();
  const int nf = 3 + nx;
  for (int i = tid; i < nf * nq; i += blockDim.x) {
    const int f = i / nq, q = i - f * nq;
    out[((size_t)c * nf + f) * nq + q] = splev_any(t, Nz, coef + (size_t)f * Nz, zq[q]);
  }
  if (coef_out)
    for (int i = tid; i < nf * Nz; i += blockDim.x) coef_out[(size_t)c * nf * Nz + i] = coef[i];
}

// ---------------------------------------------------------------------------------- point evaluation, all cosmologies
__global__ void prep_eval_many_kernel(const double* __restrict__ blob, const int64_t* __restrict__ desc, int n_tables, int table,
                                      const double* __restrict__ z, const double* __restrict__ k, size_t n,
                                      double* __restrict__ out) {
  const int c = blockIdx.y;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t* d = desc + ((size_t)c * n_tables + table) * CFP_BANK_DESC;
  out[(size_t)c * n + i] = cfp_bispev_point(blob + d[2], (int)d[0], blob + d[3], (int)d[1], blob + d[4], z[i], k[i]);
}

// ---------------------------------------------------------------------------------- no-wiggle spectrum
// One block per (cosmology, redshift).  x[s] = ln P(z, k_s) on the nsamp samples, then the Savitzky-Golay filter as
// a linear operator: interior rows are `win` taps starting at column s + off, the first and last `ne` rows are dense
// (the polynomial edge fit of scipy's mode="interp") over the first / last `wc` samples; out = exp(filtered).
__global__ void __launch_bounds__(256) prep_nowiggle_kernel(const double* __restrict__ blob, const int64_t* __restrict__ desc,
                                                            int n_tables, int table, int nzq, const double* __restrict__ zq,
                                                            int nsamp, const double* __restrict__ ks,
                                                            const int32_t* __restrict__ filt_of_c,
                                                            const int32_t* __restrict__ fdesc, const int64_t* __restrict__ foff,
                                                            const double* __restrict__ fblob, double* __restrict__ out) {
  extern __shared__ double sm[];
  double* x = sm;
  const int c = blockIdx.y, q = blockIdx.x, tid = threadIdx.x;
  const int64_t* d = desc + ((size_t)c * n_tables + table) * CFP_BANK_DESC;
  const double zz = zq[q];
  for (int s = tid; s < nsamp; s += blockDim.x)
    x[s] = log(cfp_bispev_point(blob + d[2], (int)d[0], blob + d[3], (int)d[1], blob + d[4], zz, ks[(size_t)c * nsamp + s]));
  __syncthreads();
  const int f = filt_of_c[c];
  const int win = fdesc[4 * f], off = fdesc[4 * f + 1], ne = fdesc[4 * f + 2], wc = fdesc[4 * f + 3];
  const double* taps = fblob + foff[f];
  const double* EL = taps + win;
  const double* ER = EL + (size_t)ne * wc;
  for (int s = tid; s < nsamp; s += blockDim.x) {
    double acc = 0.0;
    if (s < ne) {
      const double* e = EL + (size_t)s * wc;
      for (int u = 0; u < wc; ++u) acc = fma(e[u], x[u], acc);
    } else if (s >= nsamp - ne) {
      const double* e = ER + (size_t)(s - (nsamp - ne)) * wc;
      const double* xx = x + (nsamp - wc);
      for (int u = 0; u < wc; ++u) acc = fma(e[u], xx[u], acc);
    } else {
      const double* xx = x + (s + off);
      for (int u = 0; u < win; ++u) acc = fma(taps[u], xx[u], acc);
    }
    out[((size_t)c * nzq + q) * nsamp + s] = exp(acc);
  }
}

// ---------------------------------------------------------------------------------- velocity moments
// One block per (cosmology, redshift): I_m = (1 / 6 pi^2) trapezoid(P f^m, k), m = 0, 1, 2.
__global__ void __launch_bounds__(256) prep_moments_kernel(const double* __restrict__ blob, const int64_t* __restrict__ desc,
                                                           int n_tables, int tab_p, int tab_f, int nzq,
                                                           const double* __restrict__ zq, int nk, const double* __restrict__ k,
                                                           double* __restrict__ out) {
  __shared__ double red[3][8];
  const int c = blockIdx.y, q = blockIdx.x, tid = threadIdx.x;
  const int64_t* dp = desc + ((size_t)c * n_tables + tab_p) * CFP_BANK_DESC;
  const int64_t* df = desc + ((size_t)c * n_tables + tab_f) * CFP_BANK_DESC;
  const double zz = zq[q];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  // interval s joins samples s and s + 1; a thread evaluates both ends (the tables are L1-resident)
  for (int s = tid; s < nk - 1; s += blockDim.x) {
    const double k0 = k[s], k1 = k[s + 1];
    const double p0 = cfp_bispev_point(blob + dp[2], (int)dp[0], blob + dp[3], (int)dp[1], blob + dp[4], zz, k0);
    const double p1 = cfp_bispev_point(blob + dp[2], (int)dp[0], blob + dp[3], (int)dp[1], blob + dp[4], zz, k1);
    const double f0 = cfp_bispev_point(blob + df[2], (int)df[0], blob + df[3], (int)df[1], blob + df[4], zz, k0);
    const double f1 = cfp_bispev_point(blob + df[2], (int)df[0], blob + df[3], (int)df[1], blob + df[4], zz, k1);
    const double dk = k1 - k0;
    a0 += dk * (p1 + p0) / 2.0;
    a1 += dk * (p1 * f1 + p0 * f0) / 2.0;
    a2 += dk * (p1 * (f1 * f1) + p0 * (f0 * f0)) / 2.0;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((tid & 31) == 0) red[0][tid >> 5] = a0, red[1][tid >> 5] = a1, red[2][tid >> 5] = a2;
  __syncthreads();
  if (tid < 3) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[tid][w];
    out[((size_t)c * nzq + q) * 3 + tid] = s * (1.0 / (6.0 * 3.141592653589793 * 3.141592653589793));
  }
}

struct Scratch {  // device buffers of one call, released on every exit path
  cfp_ctx* ctx;
  std::vector<void*> p;
  explicit Scratch(cfp_ctx* c) : ctx(c) {}
  ~Scratch() {
    for (void* q : p) cfp_dev_free(ctx, q);
  }
  template <typename T>
  int up(const T* src_h, size_t n, T** dst) {
    int rc = cfp_upload(ctx, src_h, n, dst);
    if (rc == CFP_OK && *dst) p.push_back(*dst);
    return rc;
  }
};

int finish(cfp_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cfp_set_error("%s: %s", what, cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}

}  // namespace

#define CFP_TRY(expr)            \
  do {                           \
    int _rc = (expr);            \
    if (_rc != CFP_OK) return _rc; \
  } while (0)

extern "C" int cfp_bank_fit(cfp_ctx* ctx, int n_cosmo, int n_tables, int Nz, int Nk, const double* z_h, const double* k_h,
                            const double* const* tab_h, const double* scale_h, cfp_bank** out) {
  CFP_REQUIRE(ctx && z_h && k_h && tab_h && scale_h && out, "NULL argument");
  CFP_REQUIRE(n_cosmo > 0 && n_tables > 0, "empty bank");
  CFP_REQUIRE(Nz >= 5 && Nk >= 5, "a bicubic fit needs at least 5 samples per axis");
  *out = nullptr;
  for (int i = 1; i < Nz; ++i) CFP_REQUIRE(z_h[i] > z_h[i - 1], "z grid must increase strictly");
  for (int c = 0; c < n_cosmo; ++c)
    for (int i = 1; i < Nk; ++i) CFP_REQUIRE(k_h[(size_t)c * Nk + i] > k_h[(size_t)c * Nk + i - 1], "k grid must increase strictly");
  CFP_CUDA(cudaSetDevice(ctx->device));
  const size_t slot_len = (size_t)(Nz + 4) + (Nk + 4) + (size_t)Nz * Nk;
  cfp_bank* b = new cfp_bank();
  b->ctx = ctx;
  b->n_cosmo = n_cosmo;
  b->n_tables = n_tables;
  b->desc_h.assign((size_t)n_cosmo * n_tables * CFP_BANK_DESC, 0);
  // every cosmology's row has the same length and layout (cfp_bank_combine relies on that): slots that are empty in
  // cosmology 0 are empty everywhere
  std::vector<int64_t> jobs;
  std::vector<int> job_tab;
  std::vector<double> scale;
  size_t total = 0;
  for (int c = 0; c < n_cosmo; ++c)
    for (int t = 0; t < n_tables; ++t) {
      const bool have = tab_h[(size_t)c * n_tables + t] != nullptr;
      if (have != (tab_h[t] != nullptr)) {
        delete b;
        CFP_REQUIRE(false, "every cosmology must provide the same table slots");
      }
      if (!have) continue;
      int64_t* d = &b->desc_h[((size_t)c * n_tables + t) * CFP_BANK_DESC];
      d[0] = Nz + 4, d[1] = Nk + 4, d[2] = (int64_t)total, d[3] = (int64_t)total + Nz + 4;
      d[4] = (int64_t)total + (Nz + 4) + (Nk + 4);
      jobs.push_back((int64_t)total);
      jobs.push_back(c);
      job_tab.push_back(t);
      scale.push_back(scale_h[(size_t)c * n_tables + t]);
      total += slot_len;
    }
  const int njobs = (int)scale.size();
  if (njobs == 0) {
    delete b;
    CFP_REQUIRE(false, "no table given");
  }
  b->n = total;
  int rc = CFP_OK;
  {
    Scratch s(ctx);
    double *z_d = nullptr, *k_d = nullptr, *tz = nullptr, *fz = nullptr, *tk = nullptr, *fk = nullptr, *sc_d = nullptr;
    double* work = nullptr;
    int64_t* jobs_d = nullptr;
    cudaError_t e = cfp_dev_alloc(ctx, (void**)&b->blob_d, total * sizeof(double));
    if (e != cudaSuccess) rc = CFP_ERR_NOMEM;
    if (rc == CFP_OK) rc = cfp_upload(ctx, b->desc_h.data(), b->desc_h.size(), &b->desc_d);
    if (rc == CFP_OK) rc = s.up(z_h, (size_t)Nz, &z_d);
    if (rc == CFP_OK) rc = s.up(k_h, (size_t)n_cosmo * Nk, &k_d);
    if (rc == CFP_OK) rc = s.up<double>(nullptr, (size_t)Nz + 4, &tz);
    if (rc == CFP_OK) rc = s.up<double>(nullptr, (size_t)3 * Nz + 4, &fz);
    if (rc == CFP_OK) rc = s.up<double>(nullptr, (size_t)n_cosmo * (Nk + 4), &tk);
    if (rc == CFP_OK) rc = s.up<double>(nullptr, (size_t)n_cosmo * (3 * Nk + 4), &fk);
    if (rc == CFP_OK) rc = s.up<double>(nullptr, (size_t)(n_cosmo + 1) * 3 * (Nz > Nk ? Nz : Nk), &work);
    if (rc == CFP_OK) rc = s.up(scale.data(), scale.size(), &sc_d);
    if (rc == CFP_OK) rc = s.up(jobs.data(), jobs.size(), &jobs_d);
    if (rc == CFP_OK) {
      prep_factor_kernel<<<n_cosmo + 1, 128, 0, ctx->stream>>>(n_cosmo, Nz, Nk, z_d, k_d, tz, fz, tk, fk, work);
      // the raw tables go straight into the coefficient areas they are solved in
      for (int j = 0; j < njobs && e == cudaSuccess; ++j) {
        const int c = (int)jobs[2 * j + 1];
        e = cudaMemcpyAsync(b->blob_d + jobs[2 * j] + (Nz + 4) + (Nk + 4), tab_h[(size_t)c * n_tables + job_tab[j]],
                            (size_t)Nz * Nk * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      }
      if (e != cudaSuccess) {
        cfp_set_error("cfp_bank_fit: %s", cudaGetErrorString(e));
        rc = CFP_ERR_CUDA;
      }
    }
    if (rc == CFP_OK) {
      const int nkn = Nz + 4 + Nk + 4;
      prep_knots_kernel<<<dim3((nkn + 127) / 128, njobs), 128, 0, ctx->stream>>>(njobs, jobs_d, Nz, Nk, tz, tk, b->blob_d);
      prep_solve_z_kernel<<<dim3((Nk + 63) / 64, njobs), 64, 0, ctx->stream>>>(njobs, jobs_d, sc_d, Nz, Nk, fz, b->blob_d);
      const int pitch = Nk | 1;
      int rpb = (int)((size_t)160 * 1024 / ((size_t)pitch * sizeof(double)));
      rpb = rpb > 8 ? 8 : rpb;  // few rows per block: the row solves are serial, the blocks are what runs in parallel
      if (rpb < 1) {
        cfp_set_error("cfp_bank_fit: %d wavenumbers per table exceed the shared-memory tile", Nk);
        rc = CFP_ERR_INVALID;
      } else {
        const size_t smem_k = (size_t)rpb * pitch * sizeof(double);
        cudaFuncSetAttribute(prep_solve_k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k);
        prep_solve_k_kernel<<<dim3((Nz + rpb - 1) / rpb, njobs), 256, smem_k, ctx->stream>>>(njobs, jobs_d, Nz, Nk, rpb, pitch,
                                                                                          fk, b->blob_d);
      }
      ctx->launches += 4;
    }
    if (rc == CFP_OK) {
      rc = finish(ctx, "cfp_bank_fit");
    }
  }
  if (rc != CFP_OK) {
    cfp_dev_free(ctx, b->blob_d);
    cfp_dev_free(ctx, b->desc_d);
    delete b;
    return rc;
  }
  *out = b;
  return CFP_OK;
}

extern "C" int cfp_background_fit(cfp_ctx* ctx, int n_cosmo, int Nz, const double* z_h, const double* hub_h, int n_extra,
                                  const double* extra_h, int ntrap, int nq, const double* zq_h, double* out_h,
                                  double* coef_h) {
  CFP_REQUIRE(ctx && z_h && hub_h && (nq == 0 || (zq_h && out_h)) && (n_extra == 0 || extra_h), "NULL argument");
  CFP_REQUIRE(n_cosmo > 0 && Nz >= 5 && Nz <= 2048 && nq >= 0 && ntrap >= 2, "bad size");
  CFP_REQUIRE(n_extra >= 0 && n_extra <= 8, "at most 8 extra functions of z");
  for (int i = 1; i < Nz; ++i) CFP_REQUIRE(z_h[i] > z_h[i - 1], "z grid must increase strictly");
  CFP_CUDA(cudaSetDevice(ctx->device));
  Scratch s(ctx);
  double *z_d = nullptr, *h_d = nullptr, *x_d = nullptr, *q_d = nullptr, *o_d = nullptr, *c_d = nullptr;
  const int nf = 3 + n_extra;
  CFP_TRY(s.up(z_h, (size_t)Nz, &z_d));
  CFP_TRY(s.up(hub_h, (size_t)n_cosmo * Nz, &h_d));
  CFP_TRY(s.up(extra_h, (size_t)n_cosmo * n_extra * Nz, &x_d));
  CFP_TRY(s.up(zq_h, (size_t)nq, &q_d));
  CFP_TRY(s.up<double>(nullptr, (size_t)n_cosmo * nf * nq, &o_d));
  if (coef_h) CFP_TRY(s.up<double>(nullptr, (size_t)n_cosmo * nf * Nz, &c_d));
  const int bg_threads = 512;
  CFP_REQUIRE(ntrap <= 4096, "at most 4096 trapezoid samples");
  const size_t smem = ((size_t)(Nz + 4) + (3 * Nz + 4) + (size_t)(nf + 3) * Nz + (size_t)(bg_threads / 32) * ntrap) * sizeof(double);
  CFP_REQUIRE(smem <= 200 * 1024, "redshift grid too long for the background kernel");
  CFP_CUDA(cudaFuncSetAttribute(prep_background_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  prep_background_kernel<<<n_cosmo, bg_threads, smem, ctx->stream>>>(Nz, z_d, h_d, n_extra, x_d, ntrap, nq, q_d, o_d, c_d);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  if (nq > 0)
    CFP_CUDA(cudaMemcpyAsync(out_h, o_d, (size_t)n_cosmo * nf * nq * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (coef_h)
    CFP_CUDA(cudaMemcpyAsync(coef_h, c_d, (size_t)n_cosmo * nf * Nz * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx, "cfp_background_fit");
}

extern "C" int cfp_bank_eval_many(cfp_ctx* ctx, const cfp_bank* bank, int table, size_t n, const double* z_h, const double* k_h,
                                  double* out_h) {
  CFP_REQUIRE(ctx && bank && z_h && k_h && out_h, "NULL argument");
  CFP_REQUIRE(table >= 0 && table < bank->n_tables, "table out of range");
  for (int c = 0; c < bank->n_cosmo; ++c) CFP_REQUIRE(bank->desc(c, table)[0] >= 8, "empty table slot");
  if (n == 0) return CFP_OK;
  CFP_CUDA(cudaSetDevice(ctx->device));
  Scratch s(ctx);
  double *z_d = nullptr, *k_d = nullptr, *o_d = nullptr;
  CFP_TRY(s.up(z_h, n, &z_d));
  CFP_TRY(s.up(k_h, n, &k_d));
  CFP_TRY(s.up<double>(nullptr, n * bank->n_cosmo, &o_d));
  prep_eval_many_kernel<<<dim3((unsigned)((n + 127) / 128), bank->n_cosmo), 128, 0, ctx->stream>>>(
      bank->blob_d, bank->desc_d, bank->n_tables, table, z_d, k_d, n, o_d);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  CFP_CUDA(cudaMemcpyAsync(out_h, o_d, n * bank->n_cosmo * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx, "cfp_bank_eval_many");
}

extern "C" int cfp_bank_nowiggle(cfp_ctx* ctx, const cfp_bank* bank, int table, int nzq, const double* zq_h, int nsamp,
                                 const double* k_h, int n_filt, const int32_t* filt_of_c_h, const int32_t* fdesc_h,
                                 const int64_t* foff_h, const double* fblob_h, size_t fblob_len, double* out_h) {
  CFP_REQUIRE(ctx && bank && zq_h && k_h && filt_of_c_h && fdesc_h && foff_h && fblob_h && out_h, "NULL argument");
  CFP_REQUIRE(table >= 0 && table < bank->n_tables && nzq > 0 && nsamp > 0 && n_filt > 0, "bad size");
  CFP_REQUIRE(nsamp <= 4096, "at most 4096 samples");
  const int NC = bank->n_cosmo;
  for (int c = 0; c < NC; ++c) {
    CFP_REQUIRE(bank->desc(c, table)[0] >= 8, "empty table slot");
    CFP_REQUIRE(filt_of_c_h[c] >= 0 && filt_of_c_h[c] < n_filt, "filter index out of range");
  }
  for (int f = 0; f < n_filt; ++f) {
    const int win = fdesc_h[4 * f], off = fdesc_h[4 * f + 1], ne = fdesc_h[4 * f + 2], wc = fdesc_h[4 * f + 3];
    CFP_REQUIRE(win >= 1 && ne >= 0 && wc >= 0 && wc <= nsamp && 2 * ne <= nsamp, "bad filter shape");
    // interior rows ne .. nsamp-ne-1 read x[s + off .. s + off + win)
    CFP_REQUIRE(ne + off >= 0 && (nsamp - ne - 1) + off + win <= nsamp, "filter taps leave the sample range");
    CFP_REQUIRE(foff_h[f] >= 0 && (size_t)foff_h[f] + win + 2 * (size_t)ne * wc <= fblob_len, "filter blob too short");
  }
  CFP_CUDA(cudaSetDevice(ctx->device));
  Scratch s(ctx);
  double *q_d = nullptr, *k_d = nullptr, *fb_d = nullptr, *o_d = nullptr;
  int32_t *fc_d = nullptr, *fd_d = nullptr;
  int64_t* fo_d = nullptr;
  CFP_TRY(s.up(zq_h, (size_t)nzq, &q_d));
  CFP_TRY(s.up(k_h, (size_t)NC * nsamp, &k_d));
  CFP_TRY(s.up(filt_of_c_h, (size_t)NC, &fc_d));
  CFP_TRY(s.up(fdesc_h, (size_t)4 * n_filt, &fd_d));
  CFP_TRY(s.up(foff_h, (size_t)n_filt, &fo_d));
  CFP_TRY(s.up(fblob_h, fblob_len, &fb_d));
  CFP_TRY(s.up<double>(nullptr, (size_t)NC * nzq * nsamp, &o_d));
  prep_nowiggle_kernel<<<dim3(nzq, NC), 256, (size_t)nsamp * sizeof(double), ctx->stream>>>(
      bank->blob_d, bank->desc_d, bank->n_tables, table, nzq, q_d, nsamp, k_d, fc_d, fd_d, fo_d, fb_d, o_d);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  CFP_CUDA(cudaMemcpyAsync(out_h, o_d, (size_t)NC * nzq * nsamp * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx, "cfp_bank_nowiggle");
}

extern "C" int cfp_bank_theta_moments(cfp_ctx* ctx, const cfp_bank* bank, int tab_p, int tab_f, int nzq, const double* zq_h,
                                      int nk, const double* k_h, double* out_h) {
  CFP_REQUIRE(ctx && bank && zq_h && k_h && out_h, "NULL argument");
  CFP_REQUIRE(tab_p >= 0 && tab_p < bank->n_tables && tab_f >= 0 && tab_f < bank->n_tables && nzq > 0 && nk >= 2, "bad size");
  const int NC = bank->n_cosmo;
  for (int c = 0; c < NC; ++c)
    CFP_REQUIRE(bank->desc(c, tab_p)[0] >= 8 && bank->desc(c, tab_f)[0] >= 8, "empty table slot");
  CFP_CUDA(cudaSetDevice(ctx->device));
  Scratch s(ctx);
  double *q_d = nullptr, *k_d = nullptr, *o_d = nullptr;
  CFP_TRY(s.up(zq_h, (size_t)nzq, &q_d));
  CFP_TRY(s.up(k_h, (size_t)nk, &k_d));
  CFP_TRY(s.up<double>(nullptr, (size_t)NC * nzq * 3, &o_d));
  prep_moments_kernel<<<dim3(nzq, NC), 256, 0, ctx->stream>>>(bank->blob_d, bank->desc_d, bank->n_tables, tab_p, tab_f, nzq,
                                                               q_d, nk, k_d, o_d);
  ctx->launches++;
  CFP_CUDA(cudaGetLastError());
  CFP_CUDA(cudaMemcpyAsync(out_h, o_d, (size_t)NC * nzq * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return finish(ctx, "cfp_bank_theta_moments");
}
