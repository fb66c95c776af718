// Context, columns and vectors of libvdl_cuda; the synthetic column generator.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <chrono>

#include "vdl_internal.h"

int vdl_fail(vdl_ctx *ctx, int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

int vdl_cuda_fail(vdl_ctx *ctx, cudaError_t e, const char *what) {
  return vdl_fail(ctx, e == cudaErrorMemoryAllocation ? VDL_ENOMEM : VDL_ECUDA, "%s: %s", what, cudaGetErrorString(e));
}

static char g_noctx_err[256] = "no context";

extern "C" int vdl_abi_version(void) { return VDL_ABI_VERSION; }
extern "C" int vdl_abi_sizeof_fused_desc(void) { return (int)sizeof(vdl_fused_desc); }

extern "C" int vdl_ctx_create(int device, vdl_ctx **out) {
  if (!out) return VDL_EINVAL;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || device < 0 || device >= n) {
    snprintf(g_noctx_err, sizeof g_noctx_err, "vdl_ctx_create: device %d not available (%s, %d devices)", device,
             cudaGetErrorString(e), n);
    return VDL_ECUDA;
  }
  vdl_ctx *ctx = new vdl_ctx();
  ctx->device = device;
  ctx->vecs.resize(1);
  if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreateWithFlags(&ctx->copy_event, cudaEventDisableTiming)) != cudaSuccess ||
      (e = cudaMalloc(&ctx->d_errflag, sizeof(int))) != cudaSuccess || (e = cudaMemset(ctx->d_errflag, 0, sizeof(int))) != cudaSuccess) {
    snprintf(g_noctx_err, sizeof g_noctx_err, "vdl_ctx_create: %s", cudaGetErrorString(e));
    delete ctx;
    return VDL_ECUDA;
  }
  {   // vectors come from the device's stream-ordered pool (cudaMallocAsync): allocation and release are stream operations,
      // never a device synchronisation; keep freed memory in the pool instead of returning it to the driver
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&ctx->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  *out = ctx;
  return VDL_OK;
}

extern "C" int vdl_ctx_destroy(vdl_ctx *ctx) {
  if (!ctx) return VDL_EINVAL;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto &v : ctx->vecs)
    if (v.live && v.owned && v.ptr) cudaFreeAsync(v.ptr, ctx->stream);
  cudaStreamSynchronize(ctx->stream);
  vdl_jit_destroy(ctx);
  if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->d_errflag) cudaFree(ctx->d_errflag);
  if (ctx->copy_event) cudaEventDestroy(ctx->copy_event);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return VDL_OK;
}

extern "C" const char *vdl_last_error(vdl_ctx *ctx) { return ctx ? ctx->err.c_str() : g_noctx_err; }
extern "C" void *vdl_ctx_stream(vdl_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t vdl_ctx_launch_count(vdl_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" int vdl_ctx_synchronize(vdl_ctx *ctx) {
  if (!ctx) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- vectors
static int vec_slot(vdl_ctx *ctx) {
  for (size_t i = 1; i < ctx->vecs.size(); i++)
    if (!ctx->vecs[i].live) return (int)i;
  ctx->vecs.emplace_back();
  return (int)ctx->vecs.size() - 1;
}

static i64 padded_bytes(int dtype, i64 rows) {
  // Bulk (TMA) copies move whole 16-byte units and tiles of up to 2048 rows may be requested for the
  // tail, so keep every allocation readable up to the next 256-byte boundary past the logical end.
  i64 b = rows * (i64)dtype;
  return ((b + 255) / 256) * 256 + 256;
}

int vec_new(vdl_ctx *ctx, int dtype, i64 len, vdl_vec *out) {
  if (dtype != VDL_I32 && dtype != VDL_I64 && dtype != VDL_U8) return vdl_fail(ctx, VDL_EINVAL, "dtype must be VDL_I32, VDL_I64 or VDL_U8 (string heap), got %d", dtype);
  if (len < 0) return vdl_fail(ctx, VDL_EINVAL, "negative length %lld", (long long)len);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  void *p = nullptr;
  i64 bytes = padded_bytes(dtype, len);
  VDL_CUDA(ctx, cudaMallocAsync(&p, (size_t)bytes, ctx->stream));
  int h = vec_slot(ctx);
  Vec &v = ctx->vecs[h];
  v = Vec();
  v.ptr = p;
  v.dtype = dtype;
  v.len = len;
  v.cap_rows = bytes / dtype;
  v.owned = true;
  v.live = true;
  v.gen = ++ctx->gen_counter;
  *out = h;
  return VDL_OK;
}

void vec_written(vdl_ctx *ctx, Vec *v) {
  v->gen = ++ctx->gen_counter;
  v->has_stats = false;
  v->is_perm = false;
  v->narrow32 = false;
}

bool vec_identity(vdl_ctx *ctx, vdl_vec h, u64 *gen) {
  if (!ctx || h <= 0 || (size_t)h >= ctx->vecs.size() || !ctx->vecs[h].live) return false;
  *gen = ctx->vecs[h].gen;
  return true;
}

int vec_new_range(vdl_ctx *ctx, i64 from, i64 step, i64 len, vdl_vec *out) {
  if (len < 0) return vdl_fail(ctx, VDL_EINVAL, "negative length %lld", (long long)len);
  int h = vec_slot(ctx);
  Vec &v = ctx->vecs[h];
  v = Vec();
  v.live = true;
  v.is_range = true;
  v.from = from;
  v.step = step;
  v.len = len;
  v.gen = ++ctx->gen_counter;
  if (from == 0 && step == 1) v.domain = len;
  *out = h;
  return VDL_OK;
}

Vec *vec_get_any(vdl_ctx *ctx, vdl_vec h) {
  if (!ctx || h <= 0 || (size_t)h >= ctx->vecs.size() || !ctx->vecs[h].live) {
    vdl_fail(ctx, VDL_EINVAL, "invalid vector handle %d", (int)h);
    return nullptr;
  }
  return &ctx->vecs[h];
}

// numeric vectors only: a string heap (VDL_U8) is nothing but the dictionary argument of vdl_op_like
Vec *vec_get(vdl_ctx *ctx, vdl_vec h) {
  Vec *v = vec_get_any(ctx, h);
  if (v && v->dtype == VDL_U8) {
    vdl_fail(ctx, VDL_EINVAL, "vector %d (%s) is a string heap: it can only be the dictionary of a Like", (int)h, v->name.c_str());
    return nullptr;
  }
  return v;
}

Operand operand_of(const Vec &v) {
  Operand o;
  o.p = v.ptr;
  o.kind = v.is_range ? 2 : (v.dtype == VDL_I32 ? 1 : 0);
  o.from = v.from;
  o.step = v.step;
  return o;
}

int scratch_reserve(vdl_ctx *ctx, size_t bytes) {
  if (bytes <= ctx->scratch_bytes) return VDL_OK;
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->scratch) cudaFree(ctx->scratch);
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  size_t want = bytes + bytes / 4 + 4096;
  VDL_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
  ctx->scratch_bytes = want;
  return VDL_OK;
}

__global__ void scalar_mail_kernel(const unsigned char *src, unsigned char *dst, int bytes) {
  for (int i = 0; i < bytes; i++) dst[i] = src[i];
}

int read_scalar(vdl_ctx *ctx, const void *device_src, void *host_dst, int bytes) {
  if (bytes < 0 || bytes > 64) return vdl_fail(ctx, VDL_EINVAL, "read_scalar: %d bytes", bytes);
  if (!ctx->h_mail) {
    VDL_CUDA(ctx, cudaHostAlloc(&ctx->h_mail, 64, cudaHostAllocMapped));
    VDL_CUDA(ctx, cudaHostGetDevicePointer(&ctx->d_mail, ctx->h_mail, 0));
  }
  scalar_mail_kernel<<<1, 1, 0, ctx->stream>>>((const unsigned char *)device_src, (unsigned char *)ctx->d_mail, bytes);
  VDL_CUDA(ctx, cudaGetLastError());
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(host_dst, ctx->h_mail, (size_t)bytes);
  return VDL_OK;
}

int wait_published(vdl_ctx *ctx, const volatile i64 *word, i64 seq) {
  if (!getenv("VDL_NO_SPIN_WAIT")) {
    const auto t0 = std::chrono::steady_clock::now();
    for (int it = 0;; it++) {
      if (*word == seq) { std::atomic_thread_fence(std::memory_order_acquire); return VDL_OK; }
      if ((it & 1023) == 1023 && std::chrono::steady_clock::now() - t0 > std::chrono::milliseconds(20)) break;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VDL_OK;
}

int check_errflag(vdl_ctx *ctx, const char *what) {
  int flag = 0;
  VDL_TRY(read_scalar(ctx, ctx->d_errflag, &flag, sizeof(int)));
  if (flag) {
    cudaMemsetAsync(ctx->d_errflag, 0, sizeof(int), ctx->stream);
    return vdl_fail(ctx, VDL_ERANGE, "%s: %d positions out of range", what, flag);
  }
  return VDL_OK;
}

extern "C" int vdl_vec_len(vdl_ctx *ctx, vdl_vec h, int64_t *len) {
  Vec *v = vec_get_any(ctx, h);
  if (!v || !len) return VDL_EINVAL;
  *len = v->len;
  return VDL_OK;
}
extern "C" int vdl_vec_dtype(vdl_ctx *ctx, vdl_vec h, int *dtype) {
  Vec *v = vec_get_any(ctx, h);
  if (!v || !dtype) return VDL_EINVAL;
  *dtype = v->dtype;
  return VDL_OK;
}
extern "C" int vdl_vec_index_space(vdl_ctx *ctx, vdl_vec h, int64_t *len) {
  Vec *v = vec_get_any(ctx, h);
  if (!v || !len) return VDL_EINVAL;
  *len = v->domain;
  return VDL_OK;
}
extern "C" void *vdl_vec_device_ptr(vdl_ctx *ctx, vdl_vec h) {
  Vec *v = vec_get_any(ctx, h);
  return v ? v->ptr : nullptr;
}
extern "C" int vdl_vec_free(vdl_ctx *ctx, vdl_vec h) {
  Vec *v = vec_get_any(ctx, h);
  if (!v) return VDL_EINVAL;
  if (!v->name.empty()) ctx->columns.erase(v->name);
  if (v->owned && v->ptr) VDL_CUDA(ctx, cudaFreeAsync(v->ptr, ctx->stream));   // ordered after the stream's pending readers
  *v = Vec();
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- columns
extern "C" int vdl_column_alloc(vdl_ctx *ctx, const char *name, int dtype, int64_t rows, vdl_vec *out) {
  if (!ctx || !name || !out) return VDL_EINVAL;
  if (ctx->columns.count(name)) return vdl_fail(ctx, VDL_EINVAL, "column %s already registered", name);
  VDL_TRY(vec_new(ctx, dtype, rows, out));
  ctx->vecs[*out].name = name;
  ctx->columns[name] = *out;
  return VDL_OK;
}

extern "C" int vdl_column_bind(vdl_ctx *ctx, const char *name, int dtype, int64_t rows, int64_t capacity_rows,
                               void *device_ptr, vdl_vec *out) {
  if (!ctx || !name || !out) return VDL_EINVAL;
  if (dtype != VDL_I32 && dtype != VDL_I64 && dtype != VDL_U8) return vdl_fail(ctx, VDL_EINVAL, "dtype must be VDL_I32, VDL_I64 or VDL_U8");
  if (((uintptr_t)device_ptr & 15) != 0) return vdl_fail(ctx, VDL_EINVAL, "column %s: device pointer must be 16-byte aligned", name);
  if (rows < 0 || capacity_rows < rows) return vdl_fail(ctx, VDL_EINVAL, "column %s: capacity %lld < rows %lld", name, (long long)capacity_rows, (long long)rows);
  if (ctx->columns.count(name)) return vdl_fail(ctx, VDL_EINVAL, "column %s already registered", name);
  int h = vec_slot(ctx);
  Vec &v = ctx->vecs[h];
  v = Vec();
  v.ptr = device_ptr;
  v.dtype = dtype;
  v.len = rows;
  v.cap_rows = capacity_rows;
  v.live = true;
  v.gen = ++ctx->gen_counter;
  v.name = name;
  ctx->columns[name] = h;
  *out = h;
  return VDL_OK;
}

// The caller wrote to the memory of a column (caller-owned memory registered with vdl_column_bind, or the pointer of
// vdl_vec_device_ptr): cached statistics are dropped and every prepared scan / probe over it re-proves its
// assumptions before its next launch.
extern "C" int vdl_column_touch(vdl_ctx *ctx, vdl_vec col) {
  Vec *v = vec_get_any(ctx, col);
  if (!v) return VDL_EINVAL;
  if (v->is_range) return vdl_fail(ctx, VDL_EINVAL, "cannot touch a range vector");
  vec_written(ctx, v);
  return VDL_OK;
}

extern "C" int vdl_vec_generation(vdl_ctx *ctx, vdl_vec h, uint64_t *gen) {
  Vec *v = vec_get_any(ctx, h);
  if (!v || !gen) return VDL_EINVAL;
  *gen = v->gen;
  return VDL_OK;
}

extern "C" int vdl_column_lookup(vdl_ctx *ctx, const char *name, vdl_vec *out) {
  if (!ctx || !name || !out) return VDL_EINVAL;
  auto it = ctx->columns.find(name);
  if (it == ctx->columns.end()) return vdl_fail(ctx, VDL_ENOTFOUND, "Load: column %s is not registered", name);
  *out = it->second;
  return VDL_OK;
}

extern "C" int vdl_column_drop(vdl_ctx *ctx, const char *name) {
  vdl_vec h;
  VDL_TRY(vdl_column_lookup(ctx, name, &h));
  return vdl_vec_free(ctx, h);
}

extern "C" int vdl_column_upload(vdl_ctx *ctx, vdl_vec col, const void *host, int64_t rows) {
  Vec *v = vec_get_any(ctx, col);
  if (!v || !host) return VDL_EINVAL;
  if (v->is_range || rows != v->len) return vdl_fail(ctx, VDL_EINVAL, "upload of %lld rows into a vector of %lld", (long long)rows, (long long)v->len);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  vec_written(ctx, v);
  VDL_CUDA(ctx, cudaMemcpyAsync(v->ptr, host, (size_t)(rows * v->dtype), cudaMemcpyHostToDevice, ctx->stream));
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VDL_OK;
}

extern "C" int vdl_column_download(vdl_ctx *ctx, vdl_vec col, void *host, int64_t rows) {
  Vec *v = vec_get_any(ctx, col);
  if (!v || !host) return VDL_EINVAL;
  if (v->is_range || rows != v->len) return vdl_fail(ctx, VDL_EINVAL, "download of %lld rows from a vector of %lld", (long long)rows, (long long)v->len);
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  VDL_CUDA(ctx, cudaMemcpyAsync(host, v->ptr, (size_t)(rows * v->dtype), cudaMemcpyDeviceToHost, ctx->stream));
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- download
__global__ void widen_i32_kernel(const int32_t *__restrict__ in, i64 *__restrict__ out, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void range_fill_kernel(i64 *__restrict__ out, i64 from, i64 step, i64 n) {
  i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (i64)((u64)from + (u64)i * (u64)step);
}

extern "C" int vdl_vec_download(vdl_ctx *ctx, vdl_vec h, int64_t *host, int64_t capacity) {
  Vec *v = vec_get(ctx, h);
  if (!v || (!host && v->len)) return VDL_EINVAL;
  if (capacity < v->len) return vdl_fail(ctx, VDL_EINVAL, "download: capacity %lld < length %lld", (long long)capacity, (long long)v->len);
  if (v->len == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  if (v->is_range) {
    for (i64 i = 0; i < v->len; i++) host[i] = (i64)((u64)v->from + (u64)i * (u64)v->step);
    return VDL_OK;
  }
  const void *src = v->ptr;
  if (v->dtype == VDL_I32) {
    VDL_TRY(scratch_reserve(ctx, (size_t)v->len * 8));
    unsigned blocks = (unsigned)((v->len + 255) / 256);
    widen_i32_kernel<<<blocks, 256, 0, ctx->stream>>>((const int32_t *)v->ptr, (i64 *)ctx->scratch, v->len);
    ctx->launches++;
    src = ctx->scratch;
  }
  VDL_CUDA(ctx, cudaMemcpyAsync(host, src, (size_t)v->len * 8, cudaMemcpyDeviceToHost, ctx->stream));
  VDL_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- column statistics
template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T *__restrict__ in, i64 n, i64 *out /* [min, max] */) {
  i64 lo = INT64_MAX, hi = INT64_MIN;
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
    i64 v = (i64)in[i];
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    i64 a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0 && n > 0) {
    atomicMin((long long *)&out[0], (long long)lo);
    atomicMax((long long *)&out[1], (long long)hi);
  }
}

extern "C" int vdl_column_analyze(vdl_ctx *ctx, vdl_vec col, int64_t *vmin, int64_t *vmax) {
  Vec *v = vec_get(ctx, col);
  if (!v) return VDL_EINVAL;
  if (v->is_range) return vdl_fail(ctx, VDL_EINVAL, "cannot analyze a range vector");
  if (!v->has_stats) {
    VDL_CUDA(ctx, cudaSetDevice(ctx->device));
    VDL_TRY(scratch_reserve(ctx, 16));
    v = vec_get(ctx, col);
    i64 init[2] = {INT64_MAX, INT64_MIN}, res[2];
    VDL_CUDA(ctx, cudaMemcpyAsync(ctx->scratch, init, 16, cudaMemcpyHostToDevice, ctx->stream));
    if (v->len > 0) {
      int blocks = ctx->sm_count * 8;
      if (v->dtype == VDL_I32) minmax_kernel<int32_t><<<blocks, 256, 0, ctx->stream>>>((const int32_t *)v->ptr, v->len, (i64 *)ctx->scratch);
      else minmax_kernel<i64><<<blocks, 256, 0, ctx->stream>>>((const i64 *)v->ptr, v->len, (i64 *)ctx->scratch);
      ctx->launches++;
    }
    VDL_TRY(read_scalar(ctx, ctx->scratch, res, 16));
    v->vmin = res[0];
    v->vmax = res[1];
    v->has_stats = true;
  }
  if (vmin) *vmin = v->vmin;
  if (vmax) *vmax = v->vmax;
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- synthetic
// Counter-based recipe (specification shared with the host generators; mplan2vdl_b200/synth.py).
__device__ __forceinline__ u64 splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

template <typename T>
__global__ void synth_fill_kernel(T *__restrict__ out, i64 rows, u64 base, int kind, i64 vmin, i64 stride, u64 p0, u64 p1, u64 row_offset) {
  for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += (i64)gridDim.x * blockDim.x) {
    u64 row = row_offset + (u64)i;
    u64 idx;
    if (kind == VDL_SYNTH_UNIFORM) idx = __umul64hi(splitmix64(base + row), p0);
    else if (kind == VDL_SYNTH_SEQ) idx = row;
    else idx = (row * p0) / p1;
    out[i] = (T)(i64)((u64)vmin + (u64)stride * idx);
  }
}

static u64 host_splitmix64(u64 x) {
  x += 0x9E3779B97F4A7C15ULL;
  u64 z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

extern "C" int vdl_column_fill_synthetic(vdl_ctx *ctx, vdl_vec col, uint64_t seed, uint64_t stream, int kind, int64_t vmin,
                                         int64_t stride, int64_t p0, int64_t p1, int64_t row_offset) {
  Vec *v = vec_get(ctx, col);
  if (!v) return VDL_EINVAL;
  if (v->is_range) return vdl_fail(ctx, VDL_EINVAL, "cannot fill a range vector");
  if (kind < 0 || kind > 2 || (kind == VDL_SYNTH_FKDENSE && p1 <= 0) || (kind == VDL_SYNTH_UNIFORM && p0 <= 0))
    return vdl_fail(ctx, VDL_EINVAL, "bad synthetic spec kind=%d p0=%lld p1=%lld", kind, (long long)p0, (long long)p1);
  vec_written(ctx, v);
  if (v->len == 0) return VDL_OK;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  u64 base = host_splitmix64(seed ^ (stream * 0x9E3779B97F4A7C15ULL));
  int blocks = ctx->sm_count * 8;
  if (v->dtype == VDL_I32)
    synth_fill_kernel<int32_t><<<blocks, 256, 0, ctx->stream>>>((int32_t *)v->ptr, v->len, base, kind, vmin, stride, (u64)p0, (u64)p1, (u64)row_offset);
  else
    synth_fill_kernel<i64><<<blocks, 256, 0, ctx->stream>>>((i64 *)v->ptr, v->len, base, kind, vmin, stride, (u64)p0, (u64)p1, (u64)row_offset);
  ctx->launches++;
  VDL_CUDA(ctx, cudaGetLastError());
  return VDL_OK;
}

// ---------------------------------------------------------------------------------- peer-addressable buffers
extern "C" int vdl_ipc_alloc(vdl_ctx *ctx, int64_t bytes, void **device_ptr) {
  if (!ctx || !device_ptr || bytes <= 0) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  VDL_CUDA(ctx, cudaMalloc(device_ptr, (size_t)bytes));
  VDL_CUDA(ctx, cudaMemset(*device_ptr, 0, (size_t)bytes));
  return VDL_OK;
}
extern "C" int vdl_ipc_export(vdl_ctx *ctx, void *device_ptr, unsigned char handle[VDL_IPC_HANDLE_BYTES]) {
  if (!ctx || !device_ptr || !handle) return VDL_EINVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == VDL_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size");
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  VDL_CUDA(ctx, cudaIpcGetMemHandle(&h, device_ptr));
  memcpy(handle, &h, sizeof h);
  return VDL_OK;
}
extern "C" int vdl_ipc_open(vdl_ctx *ctx, const unsigned char handle[VDL_IPC_HANDLE_BYTES], void **device_ptr) {
  if (!ctx || !device_ptr || !handle) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof h);
  VDL_CUDA(ctx, cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return VDL_OK;
}
extern "C" int vdl_ipc_close(vdl_ctx *ctx, void *device_ptr) {
  if (!ctx || !device_ptr) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  VDL_CUDA(ctx, cudaIpcCloseMemHandle(device_ptr));
  return VDL_OK;
}
extern "C" int vdl_ipc_free(vdl_ctx *ctx, void *device_ptr) {
  if (!ctx || !device_ptr) return VDL_EINVAL;
  VDL_CUDA(ctx, cudaSetDevice(ctx->device));
  VDL_CUDA(ctx, cudaFree(device_ptr));
  return VDL_OK;
}
