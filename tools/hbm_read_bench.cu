// Speed-of-light probe for the decode GEMVs: how fast can one B200 READ a weight-sized region?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/hbm_read_bench tools/hbm_read_bench.cu
// Variants: (a) per-thread 128-bit non-allocating loads, U loads in flight per thread, CTA-contiguous
// or grid-interleaved walk; (b) cp.async.bulk ring into shared memory (one producer lane).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// each CTA walks a contiguous block (like the GEMV row blocks)
template <int U>
__global__ void read_block(const uint4* __restrict__ src, size_t n_vec, unsigned* sink) {
  const size_t per = (n_vec + gridDim.x - 1) / gridDim.x;
  const size_t b0 = per * blockIdx.x;
  const size_t b1 = b0 + per < n_vec ? b0 + per : n_vec;
  unsigned acc = 0;
  for (size_t i = b0 + threadIdx.x; i < b1; i += (size_t)blockDim.x * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t j = i + (size_t)u * blockDim.x;
      v[u] = j < b1 ? ld_nc(src + j) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// whole grid walks the region together (grid-stride)
template <int U>
__global__ void read_stride(const uint4* __restrict__ src, size_t n_vec, unsigned* sink) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  unsigned acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride * U) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t j = i + (size_t)u * stride;
      v[u] = j < n_vec ? ld_nc(src + j) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// cp.async.bulk ring: SLOTS x CHUNK bytes in flight per CTA, consumers only touch one word per chunk
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void read_bulk(const uint8_t* __restrict__ src, size_t bytes, int chunk, int slots, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const size_t n_chunks = bytes / chunk;
  const size_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const size_t c0 = per * blockIdx.x, c1 = c0 + per < n_chunks ? c0 + per : n_chunks;
  if (threadIdx.x == 0) {
    for (int s = 0; s < slots; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  unsigned acc = 0;
  auto issue = [&](size_t c, int s) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(chunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)s * chunk)),
                 "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(bars + s))
                 : "memory");
  };
  size_t c = c0;
  for (int s = 0; s < slots && c < c1; ++s, ++c) issue(c, s);
  int s = 0;
  uint32_t par = 0;
  for (size_t d = c0; d < c1; ++d) {
    uint32_t ok = 0;
    long long spins = 0;
    while (!ok && ++spins < (1ll << 26))
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(bars + s)), "r"(par) : "memory");
    if (!ok) { *sink = 0xdeadbeefu; return; }
    acc ^= *reinterpret_cast<volatile unsigned*>(ring + (size_t)s * chunk);
    if (c < c1) { issue(c, s); ++c; }
    if (++s == slots) { s = 0; par ^= 1; }
  }
  if (acc == 0x12345678u) *sink = acc;
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  const size_t bytes = (size_t)2 << 30;  // 2 GiB region, far larger than L2
  uint8_t* buf; unsigned* sink;
  CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(buf, 1, bytes));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t n_vec = bytes / 16;
  auto time = [&](const char* name, auto launch) {
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int r = 0; r < 3; ++r) launch();
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("%-44s %8.1f GB/s\n", name, 3.0 * bytes / (ms * 1e-3) / 1e9);
  };
  char name[128];
#define RUN_LD(KERN, U, TH, PER_SM) \
  snprintf(name, sizeof name, #KERN " U=%d threads=%d ctas/sm=%d (%d KB/SM)", U, TH, PER_SM, U * TH * 16 * PER_SM / 1024); \
  time(name, [&] { KERN<U><<<sms * PER_SM, TH>>>(reinterpret_cast<const uint4*>(buf), n_vec, sink); });
  RUN_LD(read_block, 4, 512, 2) RUN_LD(read_block, 8, 512, 2) RUN_LD(read_block, 16, 512, 2)
  RUN_LD(read_block, 8, 1024, 1) RUN_LD(read_block, 16, 1024, 1) RUN_LD(read_block, 8, 512, 4)
  RUN_LD(read_block, 8, 256, 8) RUN_LD(read_block, 16, 256, 4) RUN_LD(read_block, 32, 256, 2)
  RUN_LD(read_stride, 4, 512, 2) RUN_LD(read_stride, 8, 512, 2) RUN_LD(read_stride, 16, 512, 2)
  RUN_LD(read_stride, 8, 1024, 2) RUN_LD(read_stride, 16, 256, 4) RUN_LD(read_stride, 8, 256, 8)
  CK(cudaFuncSetAttribute(read_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  const int cfgs[][3] = {{2048, 16, 1}, {2048, 32, 1}, {2048, 64, 1}, {2048, 96, 1}, {8192, 8, 1}, {8192, 16, 1},
                         {8192, 24, 1}, {16384, 12, 1}, {32768, 6, 1}, {8192, 8, 2}, {8192, 12, 2}, {4096, 24, 2},
                         {2048, 48, 2}, {8192, 6, 4}, {4096, 12, 4}};
  for (auto& c : cfgs) {
    snprintf(name, sizeof name, "read_bulk chunk=%d slots=%d ctas/sm=%d (%d KB/SM)", c[0], c[1], c[2], c[0] * c[1] * c[2] / 1024);
    time(name, [&] { read_bulk<<<sms * c[2], 32, 1024 + (size_t)c[0] * c[1]>>>(buf, bytes, c[0], c[1], sink); });
  }
  // the driver's own copy kernel for reference (read + write)
  uint8_t* dst; CK(cudaMalloc(&dst, bytes / 2));
  time("cudaMemcpy D2D 1 GiB x3 (read+write bytes)", [&] { CK(cudaMemcpyAsync(dst, buf, bytes / 2, cudaMemcpyDeviceToDevice)); });
  return 0;
}
