// Batched Fisher-matrix post-processing on the device: sums over parameter unions, marginal 1-sigma
// bounds, marginalised sub-matrices and figures of merit for thousands of Fisher matrices in ONE launch.
// Reference semantics: cosmicfishpie/analysis/fisher_matrix.py:462-548 (__add__), :590-733 (inverse, confidence
// bounds), analysis/fisher_operations.py:177-272 (marginalise: invert, keep the rows/columns of the wanted parameters,
// invert back).  One CTA of 64 threads per matrix, n <= 64, everything in shared memory; the matrices are symmetric
// positive definite, so both inversions go through Cholesky factors (the reference calls LAPACK's LU inverse).
#include <cmath>

#include <math_constants.h>

#include "cfp_common.cuh"

namespace {

constexpr int FO_THREADS = 64;

// in-place Cholesky of the n x n matrix A (row stride ld), lower triangle; returns false (for every thread) if a pivot
// is not positive
__device__ bool fo_cholesky(double* A, int n, int ld, int tid) {
  __shared__ int bad;
  if (tid == 0) bad = 0;
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double akk = A[k * ld + k];
    __syncthreads();
    if (!(akk > 0.0)) {
      if (tid == 0) bad = 1;
      __syncthreads();
      break;
    }
    const double d = sqrt(akk), rd = 1.0 / d;
    if (tid == 0) A[k * ld + k] = d;
    for (int i = k + 1 + tid; i < n; i += FO_THREADS) A[i * ld + k] *= rd;
    __syncthreads();
    for (int i = k + 1 + tid; i < n; i += FO_THREADS) {
      const double lik = A[i * ld + k];
      for (int j = k + 1; j <= i; ++j) A[i * ld + j] -= lik * A[j * ld + k];
    }
    __syncthreads();
  }
  __syncthreads();
  return bad == 0;
}

// Inv = (L L^T)^-1 from the Cholesky factor L (lower triangle of A): column j of L^-1 by forward substitution
// (thread j), then Inv = L^-T L^-1.  W is n x n scratch (row stride ld).
__device__ void fo_inverse_from_chol(const double* L, double* W, double* Inv, int n, int ld, int tid) {
  for (int j = tid; j < n; j += FO_THREADS) {  // W[i][j] = (L^-1)[i][j], i >= j
    for (int i = 0; i < j; ++i) W[i * ld + j] = 0.0;
    W[j * ld + j] = 1.0 / L[j * ld + j];
    for (int i = j + 1; i < n; ++i) {
      double s = 0.0;
      for (int k = j; k < i; ++k) s += L[i * ld + k] * W[k * ld + j];
      W[i * ld + j] = -s / L[i * ld + i];
    }
  }
  __syncthreads();
  for (int e = tid; e < n * n; e += FO_THREADS) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    double s = 0.0;
    for (int k = i; k < n; ++k) s += W[k * ld + i] * W[k * ld + j];  // k >= max(i, j) = i
    Inv[i * ld + j] = s;
    Inv[j * ld + i] = s;
  }
  __syncthreads();
}

// per matrix b: sigma[b][i] = coef sqrt((F^-1)_ii); marg[b] = ((F^-1)[keep][keep])^-1; fom[b] = sqrt(det marg[b])
__global__ void __launch_bounds__(FO_THREADS) fisher_post_kernel(int B, int n, const double* __restrict__ F, int m,
                                                                 const int* __restrict__ keep, double coef,
                                                                 double* __restrict__ sigma, double* __restrict__ marg,
                                                                 double* __restrict__ fom) {
  extern __shared__ double fo_sm[];
  const int b = blockIdx.x, tid = threadIdx.x, ld = n + 1;
  double* A = fo_sm;           // [n][ld]
  double* W = A + n * ld;      // [n][ld]
  double* Inv = W + n * ld;    // [n][ld]
  const double* Fb = F + (size_t)b * n * n;
  for (int e = tid; e < n * n; e += FO_THREADS) A[(e / n) * ld + (e % n)] = Fb[e];
  __syncthreads();
  const bool ok = fo_cholesky(A, n, ld, tid);
  if (!ok) {  // not positive definite: NaN, like the square root of a negative variance
    for (int i = tid; i < n; i += FO_THREADS) sigma[(size_t)b * n + i] = CUDART_NAN;
    for (int e = tid; e < m * m; e += FO_THREADS) marg[(size_t)b * m * m + e] = CUDART_NAN;
    if (tid == 0 && m > 0) fom[b] = CUDART_NAN;
    return;
  }
  fo_inverse_from_chol(A, W, Inv, n, ld, tid);
  for (int i = tid; i < n; i += FO_THREADS) sigma[(size_t)b * n + i] = coef * sqrt(Inv[i * ld + i]);
  if (m <= 0) return;
  __syncthreads();
  // the covariance block of the wanted parameters, inverted back (fisher_operations.py:208-218)
  for (int e = tid; e < m * m; e += FO_THREADS) A[(e / m) * ld + (e % m)] = Inv[keep[e / m] * ld + keep[e % m]];
  __syncthreads();
  const bool ok2 = fo_cholesky(A, m, ld, tid);
  if (!ok2) {
    for (int e = tid; e < m * m; e += FO_THREADS) marg[(size_t)b * m * m + e] = CUDART_NAN;
    if (tid == 0) fom[b] = CUDART_NAN;
    return;
  }
  if (tid == 0) {  // det(marg) = 1 / det(block) = 1 / prod(diag L)^2
    double p = 1.0;
    for (int i = 0; i < m; ++i) p *= A[i * ld + i];
    fom[b] = 1.0 / p;
  }
  fo_inverse_from_chol(A, W, Inv, m, ld, tid);
  for (int e = tid; e < m * m; e += FO_THREADS) marg[(size_t)b * m * m + e] = Inv[(e / m) * ld + (e % m)];
}

// out[b] (n_out x n_out, zero-initialised by the caller) += scatter of a[b] and of bmat[b] through the index maps
__global__ void fisher_add_kernel(int B, int na, int nb, int n_out, const int* __restrict__ map_a,
                                  const int* __restrict__ map_b, const double* __restrict__ a,
                                  const double* __restrict__ bmat, double* __restrict__ out) {
  const int b = blockIdx.x;
  double* o = out + (size_t)b * n_out * n_out;
  for (int e = threadIdx.x; e < n_out * n_out; e += blockDim.x) o[e] = 0.0;
  __syncthreads();
  for (int e = threadIdx.x; e < na * na; e += blockDim.x)
    o[map_a[e / na] * n_out + map_a[e % na]] += a[(size_t)b * na * na + e];
  __syncthreads();
  for (int e = threadIdx.x; e < nb * nb; e += blockDim.x)
    o[map_b[e / nb] * n_out + map_b[e % nb]] += bmat[(size_t)b * nb * nb + e];
}

}  // namespace

extern "C" int cfp_fisher_batch_post(cfp_ctx* ctx, int B, int n, const double* fisher_h, int m, const int32_t* keep_h,
                                     double coef, double* sigma_h, double* marg_h, double* fom_h) {
  CFP_REQUIRE(ctx && fisher_h && sigma_h, "NULL argument");
  CFP_REQUIRE(B >= 1 && n >= 1 && n <= 64 && m >= 0 && m <= n, "bad dimensions (at most 64 parameters)");
  CFP_REQUIRE(m == 0 || (keep_h && marg_h && fom_h), "keep / marg / fom are required when m > 0");
  for (int i = 0; i < m; ++i) CFP_REQUIRE(keep_h[i] >= 0 && keep_h[i] < n, "kept parameter index out of range");
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  double *F_d = nullptr, *s_d = nullptr, *m_d = nullptr, *f_d = nullptr;
  int* k_d = nullptr;
  int rc = cfp_upload<double>(ctx, fisher_h, (size_t)B * n * n, &F_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)B * n, &s_d);
  if (rc == CFP_OK && m > 0) rc = cfp_upload<int>(ctx, keep_h, (size_t)m, &k_d);
  if (rc == CFP_OK && m > 0) rc = cfp_upload<double>(ctx, nullptr, (size_t)B * m * m, &m_d);
  if (rc == CFP_OK && m > 0) rc = cfp_upload<double>(ctx, nullptr, (size_t)B, &f_d);
  if (rc == CFP_OK) {
    const size_t smem = (size_t)3 * n * (n + 1) * sizeof(double);
    cudaError_t e = cudaFuncSetAttribute(fisher_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) {
      fisher_post_kernel<<<B, FO_THREADS, smem, st>>>(B, n, F_d, m, k_d, coef, s_d, m_d, f_d);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(sigma_h, s_d, (size_t)B * n * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && m > 0)
      e = cudaMemcpyAsync(marg_h, m_d, (size_t)B * m * m * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && m > 0) e = cudaMemcpyAsync(fom_h, f_d, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_fisher_batch_post: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  for (void* q : {(void*)F_d, (void*)s_d, (void*)m_d, (void*)f_d, (void*)k_d}) cfp_dev_free(ctx, q);
  return rc;
}

extern "C" int cfp_fisher_batch_add(cfp_ctx* ctx, int B, int na, int nb, int n_out, const int32_t* map_a_h,
                                    const int32_t* map_b_h, const double* a_h, const double* b_h, double* out_h) {
  CFP_REQUIRE(ctx && map_a_h && map_b_h && a_h && b_h && out_h, "NULL argument");
  CFP_REQUIRE(B >= 1 && na >= 1 && nb >= 1 && n_out >= 1, "bad dimensions");
  for (int i = 0; i < na; ++i) CFP_REQUIRE(map_a_h[i] >= 0 && map_a_h[i] < n_out, "index map out of range");
  for (int i = 0; i < nb; ++i) CFP_REQUIRE(map_b_h[i] >= 0 && map_b_h[i] < n_out, "index map out of range");
  CFP_CUDA(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  double *a_d = nullptr, *b_d = nullptr, *o_d = nullptr;
  int *ma_d = nullptr, *mb_d = nullptr;
  int rc = cfp_upload<double>(ctx, a_h, (size_t)B * na * na, &a_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, b_h, (size_t)B * nb * nb, &b_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, (size_t)B * n_out * n_out, &o_d);
  if (rc == CFP_OK) rc = cfp_upload<int>(ctx, map_a_h, (size_t)na, &ma_d);
  if (rc == CFP_OK) rc = cfp_upload<int>(ctx, map_b_h, (size_t)nb, &mb_d);
  if (rc == CFP_OK) {
    fisher_add_kernel<<<B, 256, 0, st>>>(B, na, nb, n_out, ma_d, mb_d, a_d, b_d, o_d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(out_h, o_d, (size_t)B * n_out * n_out * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_fisher_batch_add: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  for (void* q : {(void*)a_d, (void*)b_d, (void*)o_d, (void*)ma_d, (void*)mb_d}) cfp_dev_free(ctx, q);
  return rc;
}
