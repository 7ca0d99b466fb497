// Legendre-multipole likelihood of the spectroscopic probe (reference: likelihood/spectro_like.py:48-140).
// Small, latency-trivial kernels; everything FP64.
#include <cmath>

#include "cfp_common.cuh"

namespace {

// pl[s][k][z][l] = sum_mu lw[l][mu] * lat[s][z][mu][k]
__global__ void legendre_project_kernel(int S, int Z, int Nmu, int Nk, const double* __restrict__ lat,
                                        const double* __restrict__ lw, double* __restrict__ pl) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t tot = (size_t)S * Z * Nk;
  if (idx >= tot) return;
  const int k = (int)(idx % Nk);
  const int z = (int)((idx / Nk) % Z);
  const int s = (int)(idx / ((size_t)Nk * Z));
  const double* p = lat + (((size_t)s * Z + z) * Nmu) * Nk + k;
  double a0 = 0.0, a2 = 0.0, a4 = 0.0;
  for (int m = 0; m < Nmu; ++m) {
    const double v = p[(size_t)m * Nk];
    a0 += lw[m] * v;
    a2 += lw[Nmu + m] * v;
    a4 += lw[2 * Nmu + m] * v;
  }
  double* o = pl + (((size_t)s * Nk + k) * Z + z) * 3;
  o[0] = a0;
  o[1] = a2;
  o[2] = a4;
}

// mode-by-mode covariance times k^3 at node i (spectro_like.py:80-106); m = m00,m22,m44,m02,m24,m04
__device__ void mode_cov(const double* __restrict__ pl, const double* __restrict__ m, double k, double vol, int Z, int i,
                         int z, double c[9]) {
  const double* P = pl + ((size_t)i * Z + z) * 3;
  const double p0 = P[0], p2 = P[1], p4 = P[2];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      const int e = a * 3 + b;
      const double sum = p0 * p0 * m[e] + p2 * p2 * m[9 + e] + p4 * p4 * m[18 + e] + 2.0 * (p0 * p2 * m[27 + e]) +
                         2.0 * (p2 * p4 * m[36 + e]) + 2.0 * (p0 * p4 * m[45 + e]);
      c[e] = 2.0 * (4.0 * a + 1.0) * (4.0 * b + 1.0) / vol * sum * (k * k * k);
    }
}

// integral over [a, b] of the line through (x0, y0), (x1, y1)
__device__ __forceinline__ double line_int(double x0, double y0, double x1, double y1, double a, double b) {
  const double sl = (y1 - y0) / (x1 - x0);
  const double ya = y0 + sl * (a - x0), yb = y0 + sl * (b - x0);
  return 0.5 * (ya + yb) * (b - a);
}

__global__ void legendre_cov_kernel(int Nk, int Z, const double* __restrict__ pl, const double* __restrict__ kk,
                                    const double* __restrict__ vol, const double* __restrict__ m,
                                    double* __restrict__ cov, double* __restrict__ icov) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Nk * Z) return;
  const int z = idx % Z, i = idx / Z;
  // ln k bin edges: midpoints, mirrored at both ends (spectro_like.py:96-104)
  const double li = log(kk[i]);
  const double lm = (i > 0) ? log(kk[i - 1]) : 0.0, lp = (i < Nk - 1) ? log(kk[i + 1]) : 0.0;
  double ea, eb;
  if (i == 0) {
    eb = 0.5 * (lp + li);
    ea = eb - (lp - li);
  } else if (i == Nk - 1) {
    ea = 0.5 * (li + lm);
    eb = ea + (li - lm);
  } else {
    ea = 0.5 * (li + lm);
    eb = 0.5 * (lp + li);
  }
  const double kvol = 4.0 * M_PI / 3.0 * (exp(3.0 * eb) - exp(3.0 * ea));
  double c0[9], cm[9], cp[9];
  mode_cov(pl, m, kk[i], vol[z], Z, i, z, c0);
  if (i > 0) mode_cov(pl, m, kk[i - 1], vol[z], Z, i - 1, z, cm);
  if (i < Nk - 1) mode_cov(pl, m, kk[i + 1], vol[z], Z, i + 1, z, cp);
  const double norm = 2.0 * pow(2.0 * M_PI, 4.0) / (kvol * kvol);
  double C[9];
  for (int e = 0; e < 9; ++e) {
    double v;
    // degree-1 spline in ln k through the nodes, extended linearly beyond the first / last node
    if (i == 0)
      v = line_int(li, c0[e], lp, cp[e], ea, eb);
    else if (i == Nk - 1)
      v = line_int(lm, cm[e], li, c0[e], ea, eb);
    else
      v = line_int(lm, cm[e], li, c0[e], ea, li) + line_int(li, c0[e], lp, cp[e], li, eb);
    C[e] = v * norm;
  }
  double* oc = cov + (size_t)idx * 9;
  for (int e = 0; e < 9; ++e) oc[e] = C[e];
  // 3x3 inverse by cofactors
  const double a = C[0], b = C[1], c = C[2], d = C[3], e_ = C[4], f = C[5], g = C[6], h = C[7], j = C[8];
  const double A = e_ * j - f * h, B = -(d * j - f * g), Cc = d * h - e_ * g;
  const double det = a * A + b * B + c * Cc;
  const double r = 1.0 / det;
  double* oi = icov + (size_t)idx * 9;
  oi[0] = A * r;
  oi[1] = -(b * j - c * h) * r;
  oi[2] = (b * f - c * e_) * r;
  oi[3] = B * r;
  oi[4] = (a * j - c * g) * r;
  oi[5] = -(a * f - c * d) * r;
  oi[6] = Cc * r;
  oi[7] = -(a * h - b * g) * r;
  oi[8] = (a * e_ - b * d) * r;
}

// one CTA per point: chi2[s] = sum_{k,z} d^T icov d, fixed-order block reduction
__global__ void __launch_bounds__(256) legendre_chi2_kernel(int Nk, int Z, const double* __restrict__ pd,
                                                            const double* __restrict__ pt, const double* __restrict__ icov,
                                                            double* __restrict__ chi2) {
  __shared__ double red[8];
  const int s = blockIdx.x, tid = threadIdx.x;
  const double* T = pt + (size_t)s * Nk * Z * 3;
  double acc = 0.0;
  for (int i = tid; i < Nk * Z; i += blockDim.x) {
    const double d0 = pd[i * 3] - T[i * 3], d1 = pd[i * 3 + 1] - T[i * 3 + 1], d2 = pd[i * 3 + 2] - T[i * 3 + 2];
    const double* M = icov + (size_t)i * 9;
    acc += d0 * (M[0] * d0 + M[1] * d1 + M[2] * d2) + d1 * (M[3] * d0 + M[4] * d1 + M[5] * d2) +
           d2 * (M[6] * d0 + M[7] * d1 + M[8] * d2);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((tid & 31) == 0) red[tid >> 5] = acc;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    chi2[s] = t;
  }
}

struct DevBufs {
  std::vector<void*> p;
  cfp_ctx* owner = nullptr;
  ~DevBufs() {
    for (void* q : p) cfp_dev_free(owner, q);
  }
  template <typename T>
  int up(cfp_ctx* ctx, const T* h, size_t n, T** d) {
    owner = ctx;
    int rc = cfp_upload<T>(ctx, h, n, d);
    if (rc == CFP_OK) p.push_back((void*)*d);
    return rc;
  }
};

}  // namespace

#define LK_TRY(x)                 \
  do {                            \
    int _rc = (x);                \
    if (_rc != CFP_OK) return _rc; \
  } while (0)

static int lk_finish(cfp_ctx* ctx, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cfp_set_error("%s: %s", what, cudaGetErrorString(e));
    return CFP_ERR_CUDA;
  }
  return CFP_OK;
}

extern "C" int cfp_legendre_project(cfp_ctx* ctx, int S, int Z, int Nmu, int Nk, const double* lat_h, const double* lw_h,
                                    double* pl_h) {
  CFP_REQUIRE(ctx && lat_h && lw_h && pl_h, "NULL argument");
  CFP_REQUIRE(S >= 1 && Z >= 1 && Nmu >= 1 && Nk >= 1, "bad dimensions");
  CFP_CUDA(cudaSetDevice(ctx->device));
  DevBufs B;
  double *lat_d, *lw_d, *pl_d;
  LK_TRY(B.up(ctx, lat_h, (size_t)S * Z * Nmu * Nk, &lat_d));
  LK_TRY(B.up(ctx, lw_h, (size_t)3 * Nmu, &lw_d));
  LK_TRY(B.up<double>(ctx, nullptr, (size_t)S * Nk * Z * 3, &pl_d));
  const size_t tot = (size_t)S * Z * Nk;
  legendre_project_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, ctx->stream>>>(S, Z, Nmu, Nk, lat_d, lw_d, pl_d);
  ctx->launches++;
  CFP_CUDA(cudaMemcpyAsync(pl_h, pl_d, (size_t)S * Nk * Z * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return lk_finish(ctx, "cfp_legendre_project");
}

extern "C" int cfp_legendre_covariance(cfp_ctx* ctx, int Nk, int Z, const double* pl_h, const double* k_h,
                                       const double* vol_h, const double* m_h, double* cov_h, double* icov_h) {
  CFP_REQUIRE(ctx && pl_h && k_h && vol_h && m_h && cov_h && icov_h, "NULL argument");
  CFP_REQUIRE(Nk >= 2 && Z >= 1, "bad dimensions");
  CFP_CUDA(cudaSetDevice(ctx->device));
  DevBufs B;
  double *pl_d, *k_d, *vol_d, *m_d, *cov_d, *icov_d;
  LK_TRY(B.up(ctx, pl_h, (size_t)Nk * Z * 3, &pl_d));
  LK_TRY(B.up(ctx, k_h, (size_t)Nk, &k_d));
  LK_TRY(B.up(ctx, vol_h, (size_t)Z, &vol_d));
  LK_TRY(B.up(ctx, m_h, (size_t)54, &m_d));
  LK_TRY(B.up<double>(ctx, nullptr, (size_t)Nk * Z * 9, &cov_d));
  LK_TRY(B.up<double>(ctx, nullptr, (size_t)Nk * Z * 9, &icov_d));
  legendre_cov_kernel<<<(Nk * Z + 127) / 128, 128, 0, ctx->stream>>>(Nk, Z, pl_d, k_d, vol_d, m_d, cov_d, icov_d);
  ctx->launches++;
  CFP_CUDA(cudaMemcpyAsync(cov_h, cov_d, (size_t)Nk * Z * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CFP_CUDA(cudaMemcpyAsync(icov_h, icov_d, (size_t)Nk * Z * 9 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return lk_finish(ctx, "cfp_legendre_covariance");
}

extern "C" int cfp_legendre_chi2(cfp_ctx* ctx, int S, int Nk, int Z, const double* pl_data_h, const double* pl_theory_h,
                                 const double* icov_h, double* chi2_h) {
  CFP_REQUIRE(ctx && pl_data_h && pl_theory_h && icov_h && chi2_h, "NULL argument");
  CFP_REQUIRE(S >= 1 && Nk >= 1 && Z >= 1, "bad dimensions");
  CFP_CUDA(cudaSetDevice(ctx->device));
  DevBufs B;
  double *pd_d, *pt_d, *ic_d, *c_d;
  LK_TRY(B.up(ctx, pl_data_h, (size_t)Nk * Z * 3, &pd_d));
  LK_TRY(B.up(ctx, pl_theory_h, (size_t)S * Nk * Z * 3, &pt_d));
  LK_TRY(B.up(ctx, icov_h, (size_t)Nk * Z * 9, &ic_d));
  LK_TRY(B.up<double>(ctx, nullptr, (size_t)S, &c_d));
  legendre_chi2_kernel<<<S, 256, 0, ctx->stream>>>(Nk, Z, pd_d, pt_d, ic_d, c_d);
  ctx->launches++;
  CFP_CUDA(cudaMemcpyAsync(chi2_h, c_d, (size_t)S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return lk_finish(ctx, "cfp_legendre_chi2");
}
