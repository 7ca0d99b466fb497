// Shared internals of libcfp_b200.so (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "cfp_b200.h"

void cfp_set_error(const char* fmt, ...);

#define CFP_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      cfp_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CFP_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

#define CFP_REQUIRE(cond, msg)                               \
  do {                                                       \
    if (!(cond)) {                                           \
      cfp_set_error("%s:%d: %s", __FILE__, __LINE__, msg);   \
      return CFP_ERR_INVALID;                                \
    }                                                        \
  } while (0)

struct cfp_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t aux = nullptr;  // second stream for callers that overlap a running batch with the next one's uploads
  cudaStream_t aux2 = nullptr; // and a third: consecutive batches on alternating streams overlap each other's kernels
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::atomic<int64_t> launches{0};  // plans of one context may be driven from two host threads (pipelined batches)
  double last_ms = 0.0;
  cudaMemPool_t pool = nullptr;  // stream-ordered arena: plan create/destroy never reach the driver allocator
  // page-locked result buffers, recycled: cudaHostAlloc / cudaFreeHost synchronise the device, which would serialise a
  // pipeline of batches that allocates one per plan
  std::mutex pin_mu;
  std::vector<std::pair<size_t, void*>> pin_free;
};

inline void* cfp_pinned_acquire(cfp_ctx* ctx, size_t bytes) {
  {
    std::lock_guard<std::mutex> g(ctx->pin_mu);
    for (size_t i = 0; i < ctx->pin_free.size(); ++i)
      if (ctx->pin_free[i].first >= bytes) {
        void* p = ctx->pin_free[i].second;
        ctx->pin_free[i] = ctx->pin_free.back();
        ctx->pin_free.pop_back();
        return p;  // (capacity >= bytes; returned to the list with its real capacity by the matching release)
      }
  }
  void* p = nullptr;
  const size_t cap = std::max<size_t>(bytes, (size_t)1 << 16);
  if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) return nullptr;
  return p;
}
inline void cfp_pinned_release(cfp_ctx* ctx, void* p, size_t bytes) {
  if (!p) return;
  std::lock_guard<std::mutex> g(ctx->pin_mu);
  ctx->pin_free.emplace_back(std::max<size_t>(bytes, (size_t)1 << 16), p);
}

// Device memory comes from the context's pool, ordered on the context stream.  Work that runs on a caller's
// stream synchronises with the context stream after allocating and before releasing (see the plan code).
inline cudaError_t cfp_dev_alloc(cfp_ctx* ctx, void** p, size_t bytes) {
  if (ctx->pool) return cudaMallocFromPoolAsync(p, bytes, ctx->pool, ctx->stream);
  return cudaMalloc(p, bytes);
}
inline void cfp_dev_free(cfp_ctx* ctx, void* p) {
  if (!p) return;
  if (ctx && ctx->pool)
    cudaFreeAsync(p, ctx->stream);
  else
    cudaFree(p);
}

struct cfp_bank {
  cfp_ctx* ctx = nullptr;
  int n_cosmo = 0, n_tables = 0;
  size_t n = 0;
  double* blob_d = nullptr;
  int64_t* desc_d = nullptr;
  std::vector<int64_t> desc_h;  // [n_cosmo][n_tables][CFP_BANK_DESC]
  const int64_t* desc(int c, int t) const { return &desc_h[((size_t)c * n_tables + t) * CFP_BANK_DESC]; }
};

// Upload helper: pool allocation + async copy on the context stream.
template <typename T>
int cfp_upload(cfp_ctx* ctx, const T* src_h, size_t count, T** dst_d) {
  *dst_d = nullptr;
  if (count == 0) return CFP_OK;
  CFP_CUDA(cfp_dev_alloc(ctx, (void**)dst_d, count * sizeof(T)));
  if (src_h) {
    CFP_CUDA(cudaMemcpyAsync(*dst_d, src_h, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CFP_CUDA(cudaMemsetAsync(*dst_d, 0, count * sizeof(T), ctx->stream));
  }
  return CFP_OK;
}

// ------------------------------------------------------------------ device spline helpers
// Interval search with FITPACK semantics (fpbisp.f): the largest l in [3, n-5] with
// t[l] <= x, for x already clamped to [t[3], t[n-4]].
__device__ __forceinline__ int cfp_find_interval(const double* __restrict__ t, int n, double x) {
  int lo = 3, hi = n - 5;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(t + mid) <= x)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// Non-zero cubic B-spline basis values at x in [t[l], t[l+1]) -- de Boor / fpbspl.f recurrence.
__device__ __forceinline__ void cfp_bspl3(const double* __restrict__ t, int l, double x, double h[4]) {
  double hh[3];
  h[0] = 1.0;
#pragma unroll
  for (int j = 1; j <= 3; ++j) {
#pragma unroll
    for (int i = 0; i < j; ++i) hh[i] = h[i];
    h[0] = 0.0;
#pragma unroll
    for (int i = 0; i < j; ++i) {
      const double tli = __ldg(t + l + i + 1);
      const double tlj = __ldg(t + l + i + 1 - j);
      if (tli == tlj) {
        h[i + 1] = 0.0;
      } else {
        // explicit round-to-nearest intrinsics: no FMA contraction, so the result is
        // bit-identical to the (non-FMA) FITPACK build scipy ships
        const double f = __ddiv_rn(hh[i], __dsub_rn(tli, tlj));
        h[i] = __dadd_rn(h[i], __dmul_rn(f, __dsub_rn(tli, x)));
        h[i + 1] = __dmul_rn(f, __dsub_rn(x, tlj));
      }
    }
  }
}

// Point-wise bicubic evaluation (bispeu / fpbisp.f order of accumulation).
__device__ __forceinline__ double cfp_bispev_point(const double* __restrict__ tx, int nx,
                                                   const double* __restrict__ ty, int ny,
                                                   const double* __restrict__ c, double x, double y) {
  x = fmin(fmax(x, __ldg(tx + 3)), __ldg(tx + nx - 4));
  y = fmin(fmax(y, __ldg(ty + 3)), __ldg(ty + ny - 4));
  const int lx = cfp_find_interval(tx, nx, x);
  const int ly = cfp_find_interval(ty, ny, y);
  double hx[4], hy[4];
  cfp_bspl3(tx, lx, x, hx);
  cfp_bspl3(ty, ly, y, hy);
  const int ncy = ny - 4;
  const double* cc = c + (size_t)(lx - 3) * ncy + (ly - 3);
  double sp = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      sp = __dadd_rn(sp, __dmul_rn(__dmul_rn(__ldg(cc + (size_t)i * ncy + j), hx[i]), hy[j]));
  }
  return sp;
}
