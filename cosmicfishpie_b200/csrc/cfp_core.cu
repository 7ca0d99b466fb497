// Context, error handling and the cosmology table bank of libcfp_b200.so.
#include <algorithm>
#include <cstdarg>

#include "cfp_common.cuh"

static thread_local std::string g_cfp_error;

void cfp_set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_cfp_error = buf;
}

extern "C" int cfp_abi_version(void) { return CFP_ABI_VERSION; }

extern "C" const char* cfp_last_error(void) { return g_cfp_error.c_str(); }

extern "C" int cfp_ctx_create(int device, cfp_ctx** out) {
  CFP_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  int ndev = 0;
  CFP_CUDA(cudaGetDeviceCount(&ndev));
  CFP_REQUIRE(device >= 0 && device < ndev, "device index out of range");
  CFP_CUDA(cudaSetDevice(device));
  cfp_ctx* ctx = new cfp_ctx();
  ctx->device = device;
  cudaDeviceProp prop;
  CFP_CUDA(cudaGetDeviceProperties(&prop, device));
  ctx->sm_count = prop.multiProcessorCount;
  CFP_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  CFP_CUDA(cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
  CFP_CUDA(cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking));
  CFP_CUDA(cudaEventCreate(&ctx->ev0));
  CFP_CUDA(cudaEventCreate(&ctx->ev1));
  // Arena for every device buffer of banks and plans: a private stream-ordered pool that keeps what it
  // has been given (release threshold = max), so create/destroy cycles cost no driver allocation.
  const char* no_pool = getenv("CFP_NO_POOL");
  if (!(no_pool && atoi(no_pool) != 0)) {
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CFP_CUDA(cudaMemPoolCreate(&ctx->pool, &props));
    uint64_t keep = UINT64_MAX;
    CFP_CUDA(cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  *out = ctx;
  return CFP_OK;
}

extern "C" int cfp_ctx_destroy(cfp_ctx* ctx) {
  if (!ctx) return CFP_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaEventDestroy(ctx->ev0);
  cudaEventDestroy(ctx->ev1);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  cudaStreamSynchronize(ctx->aux);
  cudaStreamDestroy(ctx->aux);
  cudaStreamSynchronize(ctx->aux2);
  cudaStreamDestroy(ctx->aux2);
  cudaStreamDestroy(ctx->stream);
  for (auto& pr : ctx->pin_free) cudaFreeHost(pr.second);
  delete ctx;
  return CFP_OK;
}

extern "C" int cfp_ctx_sync(cfp_ctx* ctx) {
  CFP_REQUIRE(ctx != nullptr, "ctx is NULL");
  CFP_CUDA(cudaSetDevice(ctx->device));
  CFP_CUDA(cudaStreamSynchronize(ctx->stream));
  return CFP_OK;
}

extern "C" void* cfp_ctx_aux_stream(cfp_ctx* ctx) { return ctx ? (void*)ctx->aux : nullptr; }

extern "C" void* cfp_ctx_aux_stream2(cfp_ctx* ctx) { return ctx ? (void*)ctx->aux2 : nullptr; }

extern "C" int64_t cfp_ctx_launch_count(const cfp_ctx* ctx) { return ctx ? ctx->launches.load() : -1; }

extern "C" double cfp_ctx_last_run_ms(const cfp_ctx* ctx) { return ctx ? ctx->last_ms : -1.0; }

// ------------------------------------------------------------------------------ bank
extern "C" int cfp_bank_upload(cfp_ctx* ctx, const double* blob_h, size_t n_doubles,
                               const int64_t* desc_h, int n_cosmo, int n_tables, cfp_bank** out) {
  CFP_REQUIRE(ctx && blob_h && desc_h && out, "NULL argument");
  CFP_REQUIRE(n_cosmo > 0 && n_tables > 0 && n_doubles > 0, "empty bank");
  *out = nullptr;
  const size_t nd = (size_t)n_cosmo * n_tables * CFP_BANK_DESC;
  for (int c = 0; c < n_cosmo; ++c)
    for (int t = 0; t < n_tables; ++t) {
      const int64_t* d = desc_h + ((size_t)c * n_tables + t) * CFP_BANK_DESC;
      if (d[0] == 0 && d[1] == 0) continue;  // empty slot
      CFP_REQUIRE(d[0] >= 8 && d[1] >= 8, "bicubic table needs >= 8 knots per axis");
      CFP_REQUIRE(d[2] >= 0 && (size_t)(d[2] + d[0]) <= n_doubles, "tx out of range");
      CFP_REQUIRE(d[3] >= 0 && (size_t)(d[3] + d[1]) <= n_doubles, "ty out of range");
      CFP_REQUIRE(d[4] >= 0 && (size_t)(d[4] + (d[0] - 4) * (d[1] - 4)) <= n_doubles, "c out of range");
    }
  CFP_CUDA(cudaSetDevice(ctx->device));
  cfp_bank* b = new cfp_bank();
  b->ctx = ctx;
  b->n_cosmo = n_cosmo;
  b->n_tables = n_tables;
  b->n = n_doubles;
  b->desc_h.assign(desc_h, desc_h + nd);
  int rc = cfp_upload(ctx, blob_h, n_doubles, &b->blob_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, desc_h, nd, &b->desc_d);
  if (rc == CFP_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = CFP_ERR_CUDA;
  if (rc != CFP_OK) {
    cfp_dev_free(ctx, b->blob_d);
    cfp_dev_free(ctx, b->desc_d);
    delete b;
    return rc;
  }
  *out = b;
  return CFP_OK;
}

extern "C" int cfp_host_alloc(cfp_ctx* ctx, size_t bytes, void** out) {
  CFP_REQUIRE(ctx && out && bytes > 0, "bad argument");
  CFP_CUDA(cudaSetDevice(ctx->device));
  CFP_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocPortable));
  return CFP_OK;
}

extern "C" int cfp_host_free(cfp_ctx* ctx, void* p) {
  (void)ctx;  // page-locked memory is not tied to a device; the handle may already be gone at interpreter exit
  if (!p) return CFP_OK;
  CFP_CUDA(cudaFreeHost(p));
  return CFP_OK;
}

extern "C" int cfp_dev_upload(cfp_ctx* ctx, const double* src_h, size_t n_doubles, double** out_d) {
  CFP_REQUIRE(ctx && src_h && out_d && n_doubles > 0, "bad argument");
  CFP_CUDA(cudaSetDevice(ctx->device));
  int rc = cfp_upload(ctx, src_h, n_doubles, out_d);
  if (rc == CFP_OK) CFP_CUDA(cudaStreamSynchronize(ctx->stream));
  return rc;
}

extern "C" int cfp_dev_release(cfp_ctx* ctx, double* p_d) {
  CFP_REQUIRE(ctx != nullptr, "ctx is NULL");
  CFP_CUDA(cudaSetDevice(ctx->device));
  CFP_CUDA(cudaStreamSynchronize(ctx->stream));
  cfp_dev_free(ctx, p_d);
  return CFP_OK;
}

extern "C" int cfp_bank_upload_rows(cfp_ctx* ctx, const double* const* rows_h, const int64_t* row_len,
                                    const int64_t* desc_h, int n_cosmo, int n_tables, cfp_bank** out) {
  CFP_REQUIRE(ctx && rows_h && row_len && desc_h && out, "NULL argument");
  CFP_REQUIRE(n_cosmo > 0 && n_tables > 0, "empty bank");
  *out = nullptr;
  const size_t nd = (size_t)n_cosmo * n_tables * CFP_BANK_DESC;
  std::vector<int64_t> desc(desc_h, desc_h + nd);
  std::vector<size_t> start(n_cosmo);
  size_t total = 0;
  for (int c = 0; c < n_cosmo; ++c) {
    CFP_REQUIRE(rows_h[c] != nullptr && row_len[c] > 0, "empty bank row");
    start[c] = total;
    for (int t = 0; t < n_tables; ++t) {
      int64_t* d = &desc[((size_t)c * n_tables + t) * CFP_BANK_DESC];
      if (d[0] == 0 && d[1] == 0) continue;  // empty slot
      CFP_REQUIRE(d[0] >= 8 && d[1] >= 8, "bicubic table needs >= 8 knots per axis");
      CFP_REQUIRE(d[2] >= 0 && d[2] + d[0] <= row_len[c], "tx out of range");
      CFP_REQUIRE(d[3] >= 0 && d[3] + d[1] <= row_len[c], "ty out of range");
      CFP_REQUIRE(d[4] >= 0 && d[4] + (d[0] - 4) * (d[1] - 4) <= row_len[c], "c out of range");
      d[2] += (int64_t)total, d[3] += (int64_t)total, d[4] += (int64_t)total;  // row-local -> bank offsets
    }
    total += (size_t)row_len[c];
  }
  CFP_CUDA(cudaSetDevice(ctx->device));
  cfp_bank* b = new cfp_bank();
  b->ctx = ctx;
  b->n_cosmo = n_cosmo;
  b->n_tables = n_tables;
  b->n = total;
  b->desc_h = desc;
  cudaError_t e = cfp_dev_alloc(ctx, (void**)&b->blob_d, total * sizeof(double));
  for (int c = 0; c < n_cosmo && e == cudaSuccess; ++c)
    e = cudaMemcpyAsync(b->blob_d + start[c], rows_h[c], (size_t)row_len[c] * sizeof(double), cudaMemcpyHostToDevice,
                        ctx->stream);
  int rc = e == cudaSuccess ? cfp_upload(ctx, b->desc_h.data(), nd, &b->desc_d) : CFP_ERR_CUDA;
  // rows may be pageable or pinned (cfp_host_alloc); either way the copies are complete on return
  if (rc == CFP_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = CFP_ERR_CUDA;
  if (rc != CFP_OK) {
    if (e != cudaSuccess) cfp_set_error("cfp_bank_upload_rows: %s", cudaGetErrorString(e));
    cfp_dev_free(ctx, b->blob_d);
    cfp_dev_free(ctx, b->desc_d);
    delete b;
    return rc;
  }
  *out = b;
  return CFP_OK;
}

// out[m][e] = sum_c W[m][c] base[c][e] over the coefficient ranges of a row; knot ranges are copied from row 0.
// is_knot[e] flags the knot elements of a row (all rows share one layout).
__global__ void cfp_bank_combine_kernel(int NC, int M, size_t row_len, const double* __restrict__ base,
                                        const double* __restrict__ W, const unsigned char* __restrict__ is_knot,
                                        double* __restrict__ out) {
  const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (e >= row_len) return;
  double acc;
  if (is_knot[e]) {
    acc = base[e];
  } else {
    acc = 0.0;
    for (int c = 0; c < NC; ++c) {
      const double w = W[(size_t)m * NC + c];
      if (w != 0.0) acc = fma(w, base[(size_t)c * row_len + e], acc);
    }
  }
  out[(size_t)m * row_len + e] = acc;
}

static int bank_combine_impl(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out, bool keep_base);

extern "C" int cfp_bank_combine(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out) {
  return bank_combine_impl(ctx, base, M, W_h, out, true);
}

extern "C" int cfp_bank_combine_rows(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out) {
  return bank_combine_impl(ctx, base, M, W_h, out, false);
}

// keep_base: the new bank starts with the base rows (cfp_bank_combine); otherwise it holds the M combined rows only
static int bank_combine_impl(cfp_ctx* ctx, const cfp_bank* base, int M, const double* W_h, cfp_bank** out, bool keep_base) {
  CFP_REQUIRE(ctx && base && W_h && out && M >= 1, "bad argument");
  *out = nullptr;
  const int NC = base->n_cosmo, NT = base->n_tables;
  const int NB = keep_base ? NC : 0;  // rows carried over in front of the combined ones
  CFP_REQUIRE(base->n % (size_t)NC == 0, "the rows of the base bank differ in length");
  const size_t row_len = base->n / (size_t)NC;
  // every row must have the layout of row 0 (same tables on the same knots): descriptors equal up to the row offset
  std::vector<unsigned char> is_knot(row_len, 0);
  for (int c = 0; c < NC; ++c)
    for (int t = 0; t < NT; ++t) {
      const int64_t* d = base->desc(c, t);
      const int64_t* d0 = base->desc(0, t);
      const int64_t shift = (int64_t)((size_t)c * row_len);
      CFP_REQUIRE(d[0] == d0[0] && d[1] == d0[1], "the rows of the base bank differ in table shapes");
      if (d0[0] == 0 && d0[1] == 0) continue;
      CFP_REQUIRE(d[2] - shift == d0[2] && d[3] - shift == d0[3] && d[4] - shift == d0[4],
                  "the rows of the base bank differ in layout");
      if (c == 0) {
        for (int64_t i = 0; i < d0[0]; ++i) is_knot[(size_t)(d0[2] + i)] = 1;
        for (int64_t i = 0; i < d0[1]; ++i) is_knot[(size_t)(d0[3] + i)] = 1;
      }
    }
  CFP_CUDA(cudaSetDevice(ctx->device));
  cfp_bank* b = new cfp_bank();
  b->ctx = ctx;
  b->n_cosmo = NB + M;
  b->n_tables = NT;
  b->n = (size_t)(NB + M) * row_len;
  b->desc_h.resize((size_t)(NB + M) * NT * CFP_BANK_DESC);
  for (int c = 0; c < NB + M; ++c)
    for (int t = 0; t < NT; ++t) {
      const int64_t* d0 = base->desc(0, t);
      int64_t* d = &b->desc_h[((size_t)c * NT + t) * CFP_BANK_DESC];
      d[0] = d0[0], d[1] = d0[1];
      const int64_t shift = (int64_t)((size_t)c * row_len);
      d[2] = d0[2] + shift, d[3] = d0[3] + shift, d[4] = d0[4] + shift;
      if (d0[0] == 0 && d0[1] == 0) d[2] = d[3] = d[4] = 0;
    }
  double* W_d = nullptr;
  unsigned char* k_d = nullptr;
  cudaError_t e = cfp_dev_alloc(ctx, (void**)&b->blob_d, b->n * sizeof(double));
  int rc = e == cudaSuccess ? cfp_upload(ctx, W_h, (size_t)M * NC, &W_d) : CFP_ERR_CUDA;
  if (rc == CFP_OK) rc = cfp_upload(ctx, is_knot.data(), row_len, &k_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, b->desc_h.data(), b->desc_h.size(), &b->desc_d);
  if (rc == CFP_OK) {
    if (keep_base)
      e = cudaMemcpyAsync(b->blob_d, base->blob_d, base->n * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream);
    if (e == cudaSuccess) {
      cfp_bank_combine_kernel<<<dim3((unsigned)((row_len + 255) / 256), M), 256, 0, ctx->stream>>>(
          NC, M, row_len, base->blob_d, W_d, k_d, b->blob_d + (size_t)NB * row_len);
      ctx->launches++;
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_bank_combine: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  cfp_dev_free(ctx, W_d);
  cfp_dev_free(ctx, k_d);
  if (rc != CFP_OK) {
    cfp_dev_free(ctx, b->blob_d);
    cfp_dev_free(ctx, b->desc_d);
    delete b;
    return rc;
  }
  *out = b;
  return CFP_OK;
}

extern "C" int cfp_bank_free(cfp_ctx* ctx, cfp_bank* bank) {
  if (!bank) return CFP_OK;
  cfp_ctx* owner = bank->ctx ? bank->ctx : ctx;
  if (owner) cudaSetDevice(owner->device);
  cfp_dev_free(owner, bank->blob_d);
  cfp_dev_free(owner, bank->desc_d);
  delete bank;
  return CFP_OK;
}

__global__ void cfp_bank_eval_kernel(const double* __restrict__ blob, int64_t nx, int64_t ny, int64_t otx,
                                     int64_t oty, int64_t oc, const double* __restrict__ z,
                                     const double* __restrict__ k, size_t n, double* __restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = cfp_bispev_point(blob + otx, (int)nx, blob + oty, (int)ny, blob + oc, z[i], k[i]);
}

extern "C" int cfp_bank_eval(cfp_ctx* ctx, const cfp_bank* bank, int cosmo, int table, const double* z_h,
                             const double* k_h, size_t n, double* out_h) {
  CFP_REQUIRE(ctx && bank && z_h && k_h && out_h, "NULL argument");
  CFP_REQUIRE(cosmo >= 0 && cosmo < bank->n_cosmo && table >= 0 && table < bank->n_tables, "index out of range");
  if (n == 0) return CFP_OK;
  const int64_t* d = bank->desc(cosmo, table);
  CFP_REQUIRE(d[0] >= 8 && d[1] >= 8, "empty table slot");
  CFP_CUDA(cudaSetDevice(ctx->device));
  double *z_d = nullptr, *k_d = nullptr, *o_d = nullptr;
  int rc = cfp_upload(ctx, z_h, n, &z_d);
  if (rc == CFP_OK) rc = cfp_upload(ctx, k_h, n, &k_d);
  if (rc == CFP_OK) rc = cfp_upload<double>(ctx, nullptr, n, &o_d);
  if (rc == CFP_OK) {
    const int bs = 128;
    cfp_bank_eval_kernel<<<(unsigned)((n + bs - 1) / bs), bs, 0, ctx->stream>>>(bank->blob_d, d[0], d[1], d[2], d[3],
                                                                              d[4], z_d, k_d, n, o_d);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_h, o_d, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      cfp_set_error("cfp_bank_eval: %s", cudaGetErrorString(e));
      rc = CFP_ERR_CUDA;
    }
  }
  cfp_dev_free(ctx, z_d);
  cfp_dev_free(ctx, k_d);
  cfp_dev_free(ctx, o_d);
  return rc;
}

// ------------------------------------------------------------------------------ FP64 peaks
__global__ void __launch_bounds__(256) cfp_dfma_peak_kernel(double* out, int iters, double x) {
  double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double y = x * 0.5;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
    a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

__global__ void __launch_bounds__(256) cfp_dmma_peak_kernel(double* out, int iters, double x) {
  double c[8][2];
#pragma unroll
  for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
  const double a = x + threadIdx.x * 1e-9, b = x - threadIdx.x * 1e-9;
#pragma unroll 2
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int t = 0; t < 8; ++t)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[t][0]), "+d"(c[t][1])
                   : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

extern "C" int cfp_peak_fp64(cfp_ctx* ctx, int kind, double* tflops_out) {
  CFP_REQUIRE(ctx && tflops_out, "NULL argument");
  CFP_REQUIRE(kind == 0 || kind == 1, "kind must be 0 (DFMA) or 1 (DMMA)");
  CFP_CUDA(cudaSetDevice(ctx->device));
  const int blocks = ctx->sm_count * 8, threads = 256;
  const int iters = kind == 0 ? 8192 : 2048;
  double* out = nullptr;
  CFP_CUDA(cfp_dev_alloc(ctx, (void**)&out, (size_t)blocks * threads * sizeof(double)));
  double best = 0.0;
  for (int rep = 0; rep < 4; ++rep) {  // first rep warms up
    CFP_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    if (kind == 0)
      cfp_dfma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 0.999999);
    else
      cfp_dmma_peak_kernel<<<blocks, threads, 0, ctx->stream>>>(out, iters, 0.999999);
    ctx->launches++;
    CFP_CUDA(cudaGetLastError());
    CFP_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    CFP_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.f;
    CFP_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    const double flops = kind == 0 ? (double)blocks * threads * iters * 8.0 * 2.0
                                   : (double)blocks * (threads / 32) * iters * 8.0 * 512.0;
    if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cfp_dev_free(ctx, out);
  *tflops_out = best;
  return CFP_OK;
}
