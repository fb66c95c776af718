// Streaming-read microbenchmark used to pick the data path of the fused scan:
//   ldg   : grid-stride 128-bit loads, UNROLL independent loads in flight per thread
//   bulk  : persistent CTAs, one producer thread issuing cp.async.bulk (1-D TMA) copies into an
//           mbarrier-tracked shared-memory ring, consumer warps summing the staged words
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/membench tools/membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

template <int UNROLL>
__global__ void __launch_bounds__(256) ldg_kernel(const int4 *__restrict__ in, size_t n16, unsigned long long *out) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  long long acc = 0;
  for (; i + (UNROLL - 1) * stride < n16; i += UNROLL * stride) {
    int4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) v[u] = __ldcs(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  for (; i < n16; i += stride) { int4 v = in[i]; acc += v.x + v.y + v.z + v.w; }
  if (acc == 0x123456789LL) atomicAdd(out, 1ULL);
}

// ring of `stages` stages of `stage_bytes`, each filled by `ncopies` bulk copies (separate streams like columns)
__global__ void bulk_kernel(const char *__restrict__ in, size_t bytes, int stage_bytes, int stages, int ncopies, int hint, int consume,
                            unsigned long long *out) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t *full = (uint64_t *)(smem + (size_t)stages * stage_bytes);
  uint64_t *empty = full + stages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ncons = blockDim.x - 32;
  if (tid == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], ncons / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  size_t ntiles = bytes / stage_bytes;
  size_t stream_bytes = bytes / ncopies;           // stream c covers [c*stream_bytes, (c+1)*stream_bytes)
  int chunk = stage_bytes / ncopies;
  if (warp == 0) {
    if (lane == 0) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
      int st = 0; uint32_t ph = 0;
      for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], (uint32_t)stage_bytes);
        for (int c = 0; c < ncopies; c++) {
          const char *src = in + (size_t)c * stream_bytes + t * chunk;
          if (hint) bulk_g2s_hint(smem + (size_t)st * stage_bytes + c * chunk, src, chunk, &full[st], pol);
          else bulk_g2s(smem + (size_t)st * stage_bytes + c * chunk, src, chunk, &full[st]);
        }
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
    return;
  }
  const int ctid = tid - 32;
  long long acc = 0;
  int st = 0; uint32_t ph = 0;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    mbar_wait(&full[st], ph);
    if (consume) {
      const int4 *p = (const int4 *)(smem + (size_t)st * stage_bytes);
      for (int i = ctid; i < stage_bytes / 16; i += ncons) { int4 v = p[i]; acc += v.x + v.y + v.z + v.w; }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[st]);
    if (++st == stages) { st = 0; ph ^= 1; }
  }
  if (acc == 0x123456789LL) atomicAdd(out, 1ULL);
}

int main(int argc, char **argv) {
  size_t bytes = (size_t)(argc > 1 ? atof(argv[1]) : 8.0) * (1ull << 30);
  char *buf; unsigned long long *out;
  CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&out, 8));
  CK(cudaMemset(buf, 1, bytes)); CK(cudaMemset(out, 0, 8));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_it = [&](const char *name, auto launch) {
    for (int w = 0; w < 2; w++) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e9, sum = 0;
    for (int r = 0; r < 5; r++) {
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; sum += ms;
    }
    CK(cudaGetLastError());
    printf("%-58s best %7.3f ms  %7.1f GB/s   mean %7.1f GB/s\n", name, best, bytes / best / 1e6, bytes / (sum / 5) / 1e6);
  };
  char name[128];
  size_t n16 = bytes / 16;
  for (int bps : {4, 8, 16}) {
    snprintf(name, sizeof name, "ldg128 unroll4 blocks/SM=%d", bps);
    time_it(name, [&] { ldg_kernel<4><<<sms * bps, 256>>>((const int4 *)buf, n16, out); });
    snprintf(name, sizeof name, "ldg128 unroll8 blocks/SM=%d", bps);
    time_it(name, [&] { ldg_kernel<8><<<sms * bps, 256>>>((const int4 *)buf, n16, out); });
  }
  CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  struct Cfg { int stage_kb, stages, ncopies, hint, consume, threads, ctas_per_sm; };
  Cfg cfgs[] = {
      {28, 7, 4, 1, 1, 544, 1}, {28, 7, 4, 0, 1, 544, 1}, {28, 7, 4, 0, 0, 544, 1}, {28, 7, 1, 0, 1, 544, 1},
      {32, 6, 1, 0, 1, 288, 1}, {32, 6, 4, 0, 1, 288, 1}, {16, 12, 4, 0, 1, 288, 1}, {8, 24, 4, 0, 1, 288, 1},
      {64, 3, 4, 0, 1, 288, 1}, {16, 6, 4, 0, 1, 160, 2}, {16, 6, 1, 0, 1, 160, 2}, {8, 6, 4, 0, 1, 160, 4},
      {16, 3, 4, 0, 1, 160, 4}, {32, 3, 4, 0, 1, 288, 2}, {28, 7, 4, 0, 1, 160, 1}, {28, 7, 4, 0, 1, 96, 1},
  };
  for (auto &c : cfgs) {
    int stage_bytes = c.stage_kb * 1024;
    size_t smem = (size_t)stage_bytes * c.stages + 16 * c.stages + 64;
    snprintf(name, sizeof name, "bulk stage=%dKB x%d copies/stage=%d hint=%d consume=%d thr=%d cta/SM=%d", c.stage_kb, c.stages, c.ncopies,
             c.hint, c.consume, c.threads, c.ctas_per_sm);
    time_it(name, [&] { bulk_kernel<<<sms * c.ctas_per_sm, c.threads, smem>>>(buf, bytes, stage_bytes, c.stages, c.ncopies, c.hint, c.consume, out); });
  }
  return 0;
}
