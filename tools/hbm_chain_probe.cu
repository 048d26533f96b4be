// How fast can the decode step's weight stream be read when it is cut into the real per-layer
// matrices (7B: qkv 100.7 MB, o 33.6 MB, gate/up 180.4 MB, down 90.2 MB per layer, x32, + 262 MB
// head = 13.2 GB) and each matrix is one short kernel? Compares access patterns and launch modes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/hbm_chain_probe tools/hbm_chain_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint4 ld_nc(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void pdl_go() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// The matrix is a sequence of 8 KB pieces (512 threads x 16 B). PPU pieces form a unit; units are
// dealt to CTAs either in contiguous blocks (interleave = 0) or round-robin (interleave = 1).
// U pieces are in flight per thread, software-pipelined in two halves like the GEMV.
template <int U>
__global__ void __launch_bounds__(512) read_pieces(const uint4* __restrict__ src, int n_pieces, int ppu, int interleave,
                                                   int pdl, unsigned* sink) {
  const int n_units = (n_pieces + ppu - 1) / ppu;
  const int G = gridDim.x, c = blockIdx.x;
  int my_units, first_unit;
  if (interleave) { my_units = (n_units - c + G - 1) / G; first_unit = c; }
  else { const int per = (n_units + G - 1) / G; first_unit = c * per; my_units = max(0, min(per, n_units - first_unit)); }
  const int my_pieces = my_units * ppu;
  auto piece_of = [&](int p) {
    const int u = p / ppu, w = p % ppu;
    return (interleave ? (first_unit + u * G) : (first_unit + u)) * ppu + w;
  };
  constexpr int H = U / 2;
  uint4 a[H], b[H];
  auto load = [&](uint4 (&v)[H], int p0) {
#pragma unroll
    for (int i = 0; i < H; ++i) {
      const int p = p0 + i;
      const int gp = p < my_pieces ? piece_of(p) : n_pieces;
      v[i] = gp < n_pieces ? ld_nc(src + (size_t)gp * 512 + threadIdx.x) : make_uint4(0, 0, 0, 0);
    }
  };
  unsigned acc = 0;
  auto use = [&](const uint4 (&v)[H]) {
#pragma unroll
    for (int i = 0; i < H; ++i) acc ^= v[i].x ^ v[i].y ^ v[i].z ^ v[i].w;
  };
  load(a, 0); load(b, H);
  if (pdl) { pdl_go(); pdl_wait(); }
  for (int p0 = 0; p0 < my_pieces; p0 += U) {
    use(a); load(a, p0 + U);
    use(b); load(b, p0 + U + H);
  }
  if (acc == 0x12345678u) *sink = acc;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// one lane per CTA drives a ring of cp.async.bulk copies; chunks contiguous per CTA or round-robin
__global__ void read_bulk(const uint8_t* __restrict__ src, size_t bytes, int chunk, int slots, int interleave, int pdl,
                          unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* ring = smem + 1024;
  const int n_chunks = (int)(bytes / chunk);
  const int G = gridDim.x, c = blockIdx.x;
  int mine, first;
  if (interleave) { mine = (n_chunks - c + G - 1) / G; first = c; }
  else { const int per = (n_chunks + G - 1) / G; first = c * per; mine = max(0, min(per, n_chunks - first)); }
  if (threadIdx.x != 0) return;
  for (int s = 0; s < slots; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + s)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  auto issue = [&](int i, int s) {
    const size_t ci = interleave ? (size_t)first + (size_t)i * G : (size_t)first + i;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + s)), "r"(chunk) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(ring + (size_t)s * chunk)),
                 "l"(src + ci * chunk), "r"(chunk), "r"(smem_u32(bars + s))
                 : "memory");
  };
  int nxt = 0;
  for (int s = 0; s < slots && nxt < mine; ++s, ++nxt) issue(nxt, s);
  if (pdl) { pdl_go(); pdl_wait(); }
  unsigned acc = 0;
  int s = 0; uint32_t par = 0;
  for (int d = 0; d < mine; ++d) {
    uint32_t ok = 0; long long spins = 0;
    while (!ok && ++spins < (1ll << 24))
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(smem_u32(bars + s)), "r"(par) : "memory");
    if (!ok) { *sink = 0xdeadbeefu; return; }
    acc ^= *reinterpret_cast<volatile unsigned*>(ring + (size_t)s * chunk);
    if (nxt < mine) { issue(nxt, s); ++nxt; }
    if (++s == slots) { s = 0; par ^= 1; }
  }
  if (acc == 0x12345678u) *sink = acc;
}

struct Mat { size_t off, bytes; };

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  const int L = 32;
  const size_t D = 4096, F = 11008, V = 32003;
  std::vector<Mat> mats;
  size_t off = 0;
  auto add = [&](size_t rows, size_t k) { size_t b = rows * k * 2; b = (b + 8191) / 8192 * 8192; mats.push_back({off, b}); off += b; };
  for (int l = 0; l < L; ++l) { add(3 * D, D); add(D, D); add(2 * F, D); add(D, F); }
  add(V, D);
  const size_t total = off;
  uint8_t* buf; unsigned* sink;
  CK(cudaMalloc(&buf, total)); CK(cudaMalloc(&sink, 4)); CK(cudaMemset(buf, 1, total));
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaFuncSetAttribute(read_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  printf("decode-step weight stream: %.3f GB in %zu kernels\n", total / 1e9, mats.size());

  auto run = [&](const char* name, int pdl, auto launch_one) {
    // capture one step as a graph
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (auto& m : mats) launch_one(m, pdl);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st)); CK(cudaStreamSynchronize(st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 4; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    unsigned hs = 0; CK(cudaMemcpy(&hs, sink, 4, cudaMemcpyDeviceToHost));
    printf("%-64s pdl=%d  %7.3f ms/step  %7.1f GB/s%s\n", name, pdl, ms / 4, 4.0 * total / (ms * 1e-3) / 1e9,
           hs == 0xdeadbeefu ? "  (WATCHDOG)" : "");
    CK(cudaGraphExecDestroy(ge)); CK(cudaGraphDestroy(g));
  };
  auto cfg_launch = [&](cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl;
    cfg.attrs = attr; cfg.numAttrs = 1; cfg.stream = st;
  };
  char name[160];
  for (int pdl = 0; pdl < 2; ++pdl) {
    struct P { int u, per_sm, ppu, inter; };
    const P ps[] = {{8, 2, 1, 0}, {8, 2, 4, 0}, {16, 2, 1, 0}, {8, 2, 1, 1}, {8, 2, 2, 1}, {8, 2, 4, 1}, {16, 2, 1, 1}, {16, 2, 4, 1},
                    {8, 3, 1, 1}, {8, 4, 1, 1}, {4, 4, 1, 1}, {8, 1, 1, 1}};
    for (auto& q : ps) {
      snprintf(name, sizeof name, "ldg U=%d ctas/sm=%d unit=%dKB %s", q.u, q.per_sm, q.ppu * 8, q.inter ? "round-robin" : "cta-contiguous");
      run(name, pdl, [&](const Mat& m, int pdl_) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute attr[1]; cfg_launch(cfg, attr, pdl_);
        cfg.gridDim = dim3(sms * q.per_sm); cfg.blockDim = dim3(512);
        const uint4* src = reinterpret_cast<const uint4*>(buf + m.off);
        const int n_pieces = (int)(m.bytes / 8192);
        if (q.u == 4) CK(cudaLaunchKernelEx(&cfg, read_pieces<4>, src, n_pieces, q.ppu, q.inter, pdl_, sink));
        else if (q.u == 8) CK(cudaLaunchKernelEx(&cfg, read_pieces<8>, src, n_pieces, q.ppu, q.inter, pdl_, sink));
        else CK(cudaLaunchKernelEx(&cfg, read_pieces<16>, src, n_pieces, q.ppu, q.inter, pdl_, sink));
      });
    }
    struct Bk { int chunk, slots, per_sm, inter; };
    const Bk bs[] = {{8192, 8, 1, 0}, {8192, 8, 1, 1}, {8192, 16, 1, 1}, {8192, 8, 2, 1}, {8192, 8, 2, 0}, {16384, 6, 2, 1}, {32768, 6, 1, 1}, {8192, 6, 4, 1}};
    for (auto& q : bs) {
      snprintf(name, sizeof name, "bulk chunk=%d slots=%d ctas/sm=%d %s", q.chunk, q.slots, q.per_sm, q.inter ? "round-robin" : "cta-contiguous");
      run(name, pdl, [&](const Mat& m, int pdl_) {
        cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute attr[1]; cfg_launch(cfg, attr, pdl_);
        cfg.gridDim = dim3(sms * q.per_sm); cfg.blockDim = dim3(32);
        cfg.dynamicSmemBytes = 1024 + (size_t)q.chunk * q.slots;
        CK(cudaLaunchKernelEx(&cfg, read_bulk, (const uint8_t*)(buf + m.off), m.bytes, q.chunk, q.slots, q.inter, pdl_, sink));
      });
    }
  }
  // one kernel over the whole stream, for reference (no per-matrix boundaries)
  {
    cudaLaunchConfig_t cfg = {}; cudaLaunchAttribute attr[1]; cfg_launch(cfg, attr, 0);
    cfg.gridDim = dim3(sms * 2); cfg.blockDim = dim3(512);
    const uint4* src = reinterpret_cast<const uint4*>(buf);
    const int n_pieces = (int)(total / 8192);
    for (int inter = 0; inter < 2; ++inter) {
      CK(cudaLaunchKernelEx(&cfg, read_pieces<8>, src, n_pieces, 1, inter, 0, sink)); CK(cudaStreamSynchronize(st));
      CK(cudaEventRecord(e0, st));
      CK(cudaLaunchKernelEx(&cfg, read_pieces<8>, src, n_pieces, 1, inter, 0, sink));
      CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      printf("single kernel over the whole stream, %s: %7.3f ms  %7.1f GB/s\n", inter ? "round-robin" : "cta-contiguous", ms,
             total / (ms * 1e-3) / 1e9);
    }
  }
  return 0;
}
