// One cached decoding step of a single clip (all layers + head + arg-max) as ONE persistent kernel.
//
// Decode at B = 1 is pure weight streaming: 13.2 GB per step for the 7B model, 6 dependent phases
// per layer (qkv, attention x2, o_proj, gate/up, down) whose inputs are tiny vectors. Launching one
// kernel per phase leaves HBM idle for ~3 us at every one of the ~160 boundaries per step (ramp-up,
// tail, launch latency). Here every SM runs one persistent CTA:
//
//   warp 16 (producer)  streams this CTA's slice of EVERY weight matrix of the step, in order, into
//                       a shared-memory ring (5 slots of 16 rows x 1024 k = 32 KB) with
//                       cp.async.bulk + mbarrier complete_tx. Weights do not depend on activations,
//                       so the producer never waits for a phase to finish: it runs ahead across
//                       phase boundaries and HBM keeps streaming while the consumers sit in a grid
//                       barrier.
//   warps 0-15          consume the ring with mma.sync (m16n8k16, bf16 -> fp32): a slot is a 16-row
//                       A tile, lane (g, q) reads one 16-byte chunk of row g and one of row g+8 per
//                       32-wide K block (rows are padded by 16 B in the slot, so these reads are
//                       bank-conflict free), the activation vector sits in shared memory as bf16 and
//                       supplies column 0 of the B fragment. ~5 instructions per KB of weights per
//                       warp: the first version of this kernel used CUDA-core dot products and was
//                       bound by its consumer loop (51 instructions per 16-byte chunk, ncu).
//                       Per-row partial sums of the 16 warps meet in shared memory, the first
//                       threads run the fused epilogues (RoPE + KV append, SwiGLU, residual, logits),
//                       then a two-level grid barrier (global atomics) publishes the phase's output.
//
// Attention runs between the qkv and o_proj phases as two small phases (4 KV splits per head):
// scores + local softmax statistics, barrier, probabilities (normalised with the GLOBAL max / sum,
// then rounded to bf16 exactly like the eager reference) times V, barrier; the o_proj phase adds
// the 4 partial outputs while loading its input vector.
//
// Arithmetic and rounding points are identical to gemv.cu / decode_attention.cu (reference:
// transformers/models/llama/modeling_llama.py:53-67,124-168,171-184,199-222,292-331 and
// video_chatgpt/model/video_chatgpt.py:225-226).
#include "common.cuh"
#include "kernels.h"

#include <stdio.h>
#include <stdlib.h>

#include <vector>

namespace vcl {

namespace {

constexpr int MG_CWARPS = 16;
constexpr int MG_CONSUMERS = MG_CWARPS * 32;       // 512
constexpr int MG_THREADS = MG_CONSUMERS + 32;      // + producer warp
constexpr int MG_KC = 1024;                        // k elements per slot
constexpr int MG_ROWB = MG_KC * 2 + 64;            // padded row pitch inside a slot (bank spread)
constexpr int MG_SLOT_BYTES = 16 * MG_ROWB;        // 33024
constexpr int MG_NSLOT = 5;
constexpr int MG_RMAX = 240;                       // max rows of one phase owned by a CTA (x16)
constexpr int MG_KMAX = 14336;                     // max K (activation vector kept in smem)
constexpr int MG_SPLIT = 4;                        // KV splits per head
constexpr int MG_MAX_ITEMS = 2;                    // attention work items per CTA (B = 1)
constexpr int MG_OFF_X = MG_NSLOT * MG_SLOT_BYTES;                 // bf16 x[K]
constexpr int MG_OFF_PART = MG_OFF_X + MG_KMAX * 2;                // float part[16][RMAX]
constexpr int MG_OFF_SC = MG_OFF_PART + MG_CWARPS * MG_RMAX * 4;   // float sc[items][128]
constexpr int MG_OFF_RED = MG_OFF_SC + MG_MAX_ITEMS * 128 * 4;     // float red[64]
constexpr int MG_OFF_BAR = MG_OFF_RED + 64 * 4;
constexpr int MG_SMEM = MG_OFF_BAR + 2 * MG_NSLOT * 8 + 64;

enum { MODE_RES = 0, MODE_SWIGLU = 1, MODE_QKV = 2, MODE_LOGITS = 3 };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const void* p) {       // L2-coherent load (no stale L1 lines)
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_cg_f32(const float* p) {
  float v;
  asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_cg_bf16(const bf16* p) {
  unsigned short v;
  asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return __uint_as_float((uint32_t)v << 16);
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void cbar(int id) {      // barrier among the 512 consumer threads
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(MG_CONSUMERS) : "memory");
}
__device__ __forceinline__ void mg_mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                       uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ long long qkv_row(int v) {   // RoPE pairs (d, d+64) made adjacent
  return (long long)(v >> 7) * 128 + ((v & 127) >> 1) + (((v & 127) & 1) << 6);
}
__device__ __forceinline__ int rows_per_cta(int N) {    // multiple of 16 (one mma A tile)
  const int R = (N + gridDim.x - 1) / gridDim.x;
  return (R + 15) & ~15;
}

__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const MegaParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  bf16* xs = reinterpret_cast<bf16*>(smem + MG_OFF_X);
  float* part = reinterpret_cast<float*>(smem + MG_OFF_PART);
  float* sc = reinterpret_cast<float*>(smem + MG_OFF_SC);
  float* red = reinterpret_cast<float*>(smem + MG_OFF_RED);
  const uint32_t ring0 = smem_u32(smem);
  const uint32_t bar0 = smem_u32(smem + MG_OFF_BAR);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (MG_NSLOT + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < MG_NSLOT; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MG_CWARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  const int D = p.D, F = p.F, H = p.H, V = p.V, L = p.L;
  int slot = 0;
  uint32_t par = 0;
  auto advance = [&]() { if (++slot == MG_NSLOT) { slot = 0; par ^= 1u; } };
  // optional timeline: 8 timestamps per (CTA, phase); phases per layer: qkv, attA, attB, o, gu, down
  const int n_phase = L * 6 + 1;
  auto trace = [&](int phase, int ev) {
    if (p.trace != nullptr) p.trace[((size_t)blockIdx.x * n_phase + phase) * 8 + ev] = globaltimer_ns();
  };

  if (warp == MG_CWARPS) {
    // =============================== producer ===============================
    // The fill sequence (matrix, 16-row group, k chunk) is walked by two cursors: `ld` feeds the ring,
    // `pf` runs p.l2_slots fills ahead of it and only issues L2 prefetches (fire and forget, no
    // shared memory), so that HBM keeps streaming while the ring is full, i.e. while the consumers sit
    // in a grid barrier or in the attention phases.
    struct Cursor { int m, g0, kc, r1, nkc, K; const bf16* W; bool qkv, done; };
    auto open = [&](Cursor& c, int m) {
      for (; m <= 4 * L; ++m) {
        int N;
        if (m == 4 * L) { c.W = p.lm_head; N = V; c.K = D; c.qkv = false; }
        else {
          const MegaLayer w = p.layers[m >> 2];
          const int which = m & 3;
          c.W = which == 0 ? w.wqkv : which == 1 ? w.wo : which == 2 ? w.wgu : w.wd;
          N = which == 0 ? 3 * D : which == 1 ? D : which == 2 ? 2 * F : D;
          c.K = which == 3 ? F : D;
          c.qkv = which == 0;
        }
        const int R = rows_per_cta(N);
        const int r0 = blockIdx.x * R;
        c.r1 = min(N, r0 + R);
        if (r0 < c.r1) { c.m = m; c.g0 = r0; c.kc = 0; c.nkc = (c.K + MG_KC - 1) / MG_KC; return; }
      }
      c.done = true;
    };
    auto step = [&](Cursor& c) {
      if (++c.kc == c.nkc) {
        c.kc = 0; c.g0 += 16;
        if (c.g0 >= c.r1) open(c, c.m + 1);
      }
    };
    auto src_of = [&](const Cursor& c) {
      const long long row = c.qkv ? qkv_row(c.g0 + lane) : (long long)(c.g0 + lane);
      return c.W + row * c.K + (long long)c.kc * MG_KC;
    };
    Cursor ld, pf;
    ld.done = false; open(ld, 0);
    pf = ld;
    for (int i = 0; i < MG_NSLOT && !pf.done; ++i) step(pf);          // the ring itself covers these
    auto prefetch_one = [&]() {
      if (pf.done) return;
      const uint32_t seg = (uint32_t)min(MG_KC, pf.K - pf.kc * MG_KC) * 2u;
      if (lane < min(16, pf.r1 - pf.g0))
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_of(pf)), "r"(seg) : "memory");
      step(pf);
    };
    for (int i = 0; i < p.l2_slots; ++i) prefetch_one();
    while (!ld.done) {
      if (p.l2_slots > 0) prefetch_one();
      const int nr = min(16, ld.r1 - ld.g0);
      const uint32_t seg = (uint32_t)min(MG_KC, ld.K - ld.kc * MG_KC) * 2u;
      mbar_wait(empty_bar(slot), par ^ 1u);
      const uint32_t dst = ring0 + slot * MG_SLOT_BYTES;
      if (lane == 0) mbar_arrive_expect_tx(full_bar(slot), nr * seg);
      __syncwarp();
      if (lane < nr) bulk_g2s(dst + lane * MG_ROWB, src_of(ld), seg, full_bar(slot));
      __syncwarp();
      advance();
      const int m_done = ld.m;
      step(ld);
      if (lane == 0 && (ld.done || ld.m != m_done)) {
        const int l = m_done >> 2, which = m_done & 3;
        trace(m_done == 4 * L ? L * 6 : l * 6 + (which == 0 ? 0 : which + 2), 5);
      }
    }
    return;
  }

  // =============================== consumers ===============================
  // Grid barrier, two levels to keep atomic contention low: CTAs arrive on their group's counter
  // (16 CTAs per group, one 128-byte line each); the last arrival of a group bumps the top counter;
  // everybody polls the top counter. Counters are monotonic and zeroed by the launcher.
  unsigned epoch = 0;
  const unsigned n_groups = (gridDim.x + 15) / 16;
  const unsigned my_group = blockIdx.x / 16;
  const unsigned group_size = min(16u, gridDim.x - my_group * 16);
  unsigned* top_cnt = p.barrier;
  unsigned* grp_cnt = p.barrier + 32 * (1 + my_group);
  auto grid_sync = [&]() {
    cbar(2);
    if (tid == 0) {
      ++epoch;
      __threadfence();                               // publish this CTA's phase output (cumulative)
      unsigned old;
      asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(grp_cnt) : "memory");
      if (old + 1 == group_size * epoch) {
        __threadfence();
        asm volatile("red.add.relaxed.gpu.global.u32 [%0], 1;" ::"l"(top_cnt) : "memory");
      }
      unsigned seen;
      do {                                           // plain polling, ONE fence after the loop
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(top_cnt) : "memory");
      } while (seen < n_groups * epoch);
      __threadfence();
    }
    cbar(2);
  };

  const int g = lane >> 2, q = lane & 3;

  // ---- one GEMV phase. XL(c) returns the 8 bf16 (packed) of chunk c of the input vector ----
  auto gemv = [&](int mode, int N, int K, const bf16* norm_w, auto XL, int l, int phase) {
    if (tid == 0) trace(phase, 0);
    const int nch = K >> 3;
    const int R = rows_per_cta(N);
    const int r0 = blockIdx.x * R, r1 = min(N, r0 + R);
    const int n_rows = max(0, r1 - r0);
    const int nkc = (K + MG_KC - 1) / MG_KC;
    // activation vector -> shared memory (bf16), RMS-normalised
    {
      float ss = 0.f;
      for (int c = tid; c < nch; c += MG_CONSUMERS) {
        const uint4 u = XL(c);
        *reinterpret_cast<uint4*>(xs + c * 8) = u;
        const float f0 = bf16lo(u.x), f1 = bf16hi(u.x), f2 = bf16lo(u.y), f3 = bf16hi(u.y);
        const float f4 = bf16lo(u.z), f5 = bf16hi(u.z), f6 = bf16lo(u.w), f7 = bf16hi(u.w);
        ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3 + f4 * f4 + f5 * f5 + f6 * f6 + f7 * f7;
      }
      if (norm_w != nullptr) {
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        cbar(3);
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < MG_CWARPS; ++w) tot += red[w];
        const float rstd = rsqrtf(tot / (float)K + p.eps);
        for (int c = tid; c < nch; c += MG_CONSUMERS) {     // each thread re-reads its own chunks
          const uint4 u = *reinterpret_cast<const uint4*>(xs + c * 8);
          const uint4 gw = *reinterpret_cast<const uint4*>(norm_w + c * 8);
          uint4 o;
          o.x = bf16x2_mul(gw.x, pack_bf16x2(bf16lo(u.x) * rstd, bf16hi(u.x) * rstd));
          o.y = bf16x2_mul(gw.y, pack_bf16x2(bf16lo(u.y) * rstd, bf16hi(u.y) * rstd));
          o.z = bf16x2_mul(gw.z, pack_bf16x2(bf16lo(u.z) * rstd, bf16hi(u.z) * rstd));
          o.w = bf16x2_mul(gw.w, pack_bf16x2(bf16lo(u.w) * rstd, bf16hi(u.w) * rstd));
          *reinterpret_cast<uint4*>(xs + c * 8) = o;
        }
      }
      cbar(3);
    }
    if (tid == 0) trace(phase, 1);
    // stream the 16-row groups of this CTA out of the ring; warp w owns K blocks w, w+16 of a slot
    for (int g0 = 0; g0 < n_rows; g0 += 16) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      for (int kc = 0; kc < nkc; ++kc) {
        const int kb_n = min(MG_KC, K - kc * MG_KC) >> 5;        // 32-wide K blocks in this slot
        mbar_wait(full_bar(slot), par);
        const uint8_t* base = smem + slot * MG_SLOT_BYTES;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          const int kb = warp + 16 * t;
          if (kb < kb_n) {
            const uint4 wa = *reinterpret_cast<const uint4*>(base + g * MG_ROWB + kb * 64 + q * 16);
            const uint4 wb = *reinterpret_cast<const uint4*>(base + (g + 8) * MG_ROWB + kb * 64 + q * 16);
            uint4 xq = make_uint4(0, 0, 0, 0);
            if (g == 0) xq = *reinterpret_cast<const uint4*>(xs + kc * MG_KC + kb * 32 + q * 8);
            mg_mma(c, wa.x, wb.x, wa.y, wb.y, xq.x, xq.y);
            mg_mma(c, wa.z, wb.z, wa.w, wb.w, xq.z, xq.w);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar(slot));
        advance();
      }
      // c[0]: (row g, column 0) on lanes with q == 0; c[2]: (row g+8, column 0)
      if (q == 0) {
        part[warp * MG_RMAX + g0 + g] = c[0];
        part[warp * MG_RMAX + g0 + g + 8] = c[2];
      }
    }
    if (tid == 0) trace(phase, 2);
    cbar(3);
    // epilogue
    const bool pairs = (mode == MODE_SWIGLU || mode == MODE_QKV);
    const int n_items = pairs ? n_rows / 2 : n_rows;
    for (int it = tid; it < n_items; it += MG_CONSUMERS) {
      const int rr = pairs ? 2 * it : it;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < MG_CWARPS; ++w) {
        v0 += part[w * MG_RMAX + rr];
        if (pairs) v1 += part[w * MG_RMAX + rr + 1];
      }
      const int vrow = r0 + rr;
      if (mode == MODE_RES) {
        const float y = bf16r(v0) + ld_cg_bf16(p.h + vrow);
        p.h[vrow] = __float2bfloat16_rn(y);
      } else if (mode == MODE_LOGITS) {
        p.logits[vrow] = bf16r(v0);
      } else if (mode == MODE_SWIGLU) {
        const float gt = bf16r(v0);
        const float sg = bf16r(__fdividef(gt, 1.0f + __expf(-gt)));
        p.act[vrow >> 1] = __float2bfloat16_rn(sg * bf16r(v1));
      } else {  // MODE_QKV
        const int hr = vrow >> 7;
        const int which = hr / H, head = hr - which * H;
        const int d = (vrow & 127) >> 1;
        const float lo = bf16r(v0), hi = bf16r(v1);
        const long long coff = (long long)l * p.cache_layer_elems + ((long long)head * p.s_max + p.pos) * 128;
        if (which == 2) {
          p.vcache[coff + d] = __float2bfloat16_rn(lo);
          p.vcache[coff + d + 64] = __float2bfloat16_rn(hi);
        } else {
          const float cs = __bfloat162float(p.cos_t[(long long)p.pos * 64 + d]);
          const float sn = __bfloat162float(p.sin_t[(long long)p.pos * 64 + d]);
          const float olo = bf16r(lo * cs) + bf16r(-hi * sn);
          const float ohi = bf16r(hi * cs) + bf16r(lo * sn);
          if (which == 0) {
            p.q[head * 128 + d] = __float2bfloat16_rn(olo);
            p.q[head * 128 + d + 64] = __float2bfloat16_rn(ohi);
          } else {
            p.kcache[coff + d] = __float2bfloat16_rn(olo);
            p.kcache[coff + d + 64] = __float2bfloat16_rn(ohi);
          }
        }
      }
    }
    if (tid == 0) trace(phase, 3);
  };

  const int kv_len = p.pos + 1;
  const int per = ((kv_len + MG_SPLIT - 1) / MG_SPLIT + 15) / 16 * 16;
  const int n_att = H * MG_SPLIT;
  constexpr float LOG2E = 1.4426950408889634f;

  // ---- step start: h = embed[tok] ----
  if (blockIdx.x == 0) {
    int id = p.tok_in[0];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    for (int c = tid; c < (D >> 3); c += MG_CONSUMERS)
      *reinterpret_cast<uint4*>(p.h + c * 8) = *reinterpret_cast<const uint4*>(p.embed + (long long)id * D + c * 8);
  }
  grid_sync();

  auto x_from = [&](const bf16* src) { return [=](int c) { return ld_cg_v4(src + c * 8); }; };

  for (int l = 0; l < L; ++l) {
    const MegaLayer w = p.layers[l];
    gemv(MODE_QKV, 3 * D, D, w.ln1, x_from(p.h), l, l * 6);
    grid_sync();
    if (tid == 0) { trace(l * 6, 4); trace(l * 6 + 1, 0); }
    // ---------------- attention A: scores + local statistics ----------------
    const bf16* kc_l = p.kcache + (long long)l * p.cache_layer_elems;
    const bf16* vc_l = p.vcache + (long long)l * p.cache_layer_elems;
    {
      int k = 0;
      for (int it = blockIdx.x; it < n_att; it += gridDim.x, ++k) {
        const int split = it % MG_SPLIT, head = it / MG_SPLIT;
        const int lo = split * per, n_loc = max(0, min(kv_len - lo, per));
        const bf16* kc = kc_l + ((long long)head * p.s_max + lo) * 128;
        float* s_it = sc + k * 128;
        const int kq = tid >> 4, dl = tid & 15;      // 32 keys per pass, 16 lanes per key
        const uint4 qu = ld_cg_v4(p.q + head * 128 + dl * 8);
        const float qf[8] = {bf16lo(qu.x), bf16hi(qu.x), bf16lo(qu.y), bf16hi(qu.y),
                             bf16lo(qu.z), bf16hi(qu.z), bf16lo(qu.w), bf16hi(qu.w)};
        uint4 ku[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = kq + 32 * u;
          ku[u] = (j < n_loc) ? ld_cg_v4(kc + (long long)j * 128 + dl * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = kq + 32 * u;
          float d = qf[0] * bf16lo(ku[u].x) + qf[1] * bf16hi(ku[u].x) + qf[2] * bf16lo(ku[u].y) +
                    qf[3] * bf16hi(ku[u].y) + qf[4] * bf16lo(ku[u].z) + qf[5] * bf16hi(ku[u].z) +
                    qf[6] * bf16lo(ku[u].w) + qf[7] * bf16hi(ku[u].w);
          d += __shfl_xor_sync(0xffffffffu, d, 8);
          d += __shfl_xor_sync(0xffffffffu, d, 4);
          d += __shfl_xor_sync(0xffffffffu, d, 2);
          d += __shfl_xor_sync(0xffffffffu, d, 1);
          if (dl == 0 && j < n_loc) s_it[j] = bf16r(bf16r(d) * p.scale);
        }
        cbar(3);
        if (warp == 0) {
          float mx = -INFINITY;
          for (int j = lane; j < n_loc; j += 32) mx = fmaxf(mx, s_it[j]);
          mx = warp_max(mx);
          float sum = 0.f;
          for (int j = lane; j < n_loc; j += 32) sum += exp2f((s_it[j] - mx) * LOG2E);
          sum = warp_sum(sum);
          if (lane == 0) {
            p.att_stats[it * 2 + 0] = mx;
            p.att_stats[it * 2 + 1] = n_loc > 0 ? sum : 0.f;
          }
        }
      }
    }
    if (tid == 0) trace(l * 6 + 1, 3);
    grid_sync();
    if (tid == 0) { trace(l * 6 + 1, 4); trace(l * 6 + 2, 0); }
    // ---------------- attention B: probabilities (global max / sum) x V ----------------
    {
      int k = 0;
      for (int it = blockIdx.x; it < n_att; it += gridDim.x, ++k) {
        const int split = it % MG_SPLIT, head = it / MG_SPLIT;
        const int lo = split * per, n_loc = max(0, min(kv_len - lo, per));
        const bf16* vc = vc_l + ((long long)head * p.s_max + lo) * 128;
        const float* s_it = sc + k * 128;
        float m = -INFINITY;
#pragma unroll
        for (int r = 0; r < MG_SPLIT; ++r) m = fmaxf(m, ld_cg_f32(p.att_stats + (head * MG_SPLIT + r) * 2));
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < MG_SPLIT; ++r) {
          const float lm = ld_cg_f32(p.att_stats + (head * MG_SPLIT + r) * 2);
          const float ls = ld_cg_f32(p.att_stats + (head * MG_SPLIT + r) * 2 + 1);
          tot += (ls > 0.f) ? ls * exp2f((lm - m) * LOG2E) : 0.f;
        }
        const float inv = 1.0f / tot;
        const int kg = tid >> 4, dc = tid & 15;      // 32 key groups x 16 dim chunks
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint4 vu[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = kg + 32 * u;
          vu[u] = (j < n_loc) ? ld_cg_v4(vc + (long long)j * 128 + dc * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = kg + 32 * u;
          const float pj = (j < n_loc) ? bf16r(exp2f((s_it[j] - m) * LOG2E) * inv) : 0.f;
          acc[0] += pj * bf16lo(vu[u].x); acc[1] += pj * bf16hi(vu[u].x);
          acc[2] += pj * bf16lo(vu[u].y); acc[3] += pj * bf16hi(vu[u].y);
          acc[4] += pj * bf16lo(vu[u].z); acc[5] += pj * bf16hi(vu[u].z);
          acc[6] += pj * bf16lo(vu[u].w); acc[7] += pj * bf16hi(vu[u].w);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
        cbar(3);                                      // previous use of `part` is over
        if (lane < 16) {
#pragma unroll
          for (int e = 0; e < 8; ++e) part[warp * 128 + dc * 8 + e] = acc[e];
        }
        cbar(3);
        if (tid < 128) {
          float v = 0.f;
#pragma unroll
          for (int w = 0; w < MG_CWARPS; ++w) v += part[w * 128 + tid];
          p.att_part[(long long)it * 128 + tid] = v;
        }
      }
    }
    if (tid == 0) trace(l * 6 + 2, 3);
    grid_sync();
    if (tid == 0) trace(l * 6 + 2, 4);
    // ---------------- o_proj: input = bf16(sum of the 4 partial attention outputs) ----------------
    auto x_att = [&](int c) {
      const int head = c >> 4, d0 = (c & 15) * 8;
      const float* src = p.att_part + ((long long)head * MG_SPLIT) * 128 + d0;
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int r = 0; r < MG_SPLIT; ++r) {
        const uint4 a = ld_cg_v4(src + r * 128), bq = ld_cg_v4(src + r * 128 + 4);
        f[0] += __uint_as_float(a.x); f[1] += __uint_as_float(a.y); f[2] += __uint_as_float(a.z); f[3] += __uint_as_float(a.w);
        f[4] += __uint_as_float(bq.x); f[5] += __uint_as_float(bq.y); f[6] += __uint_as_float(bq.z); f[7] += __uint_as_float(bq.w);
      }
      return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    };
    gemv(MODE_RES, D, D, nullptr, x_att, l, l * 6 + 3);
    grid_sync();
    if (tid == 0) trace(l * 6 + 3, 4);
    gemv(MODE_SWIGLU, 2 * F, D, w.ln2, x_from(p.h), l, l * 6 + 4);
    grid_sync();
    if (tid == 0) trace(l * 6 + 4, 4);
    gemv(MODE_RES, D, F, nullptr, x_from(p.act), l, l * 6 + 5);
    grid_sync();
    if (tid == 0) trace(l * 6 + 5, 4);
  }
  gemv(MODE_LOGITS, V, D, p.norm_w, x_from(p.h), 0, L * 6);
  grid_sync();
  if (tid == 0) trace(L * 6, 4);
  // ---- arg-max (lowest index wins ties) ----
  if (blockIdx.x == 0) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += MG_CONSUMERS) {
      const float x = ld_cg_f32(p.logits + i);
      if (x > best) { best = x; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    int* redi = reinterpret_cast<int*>(red + 32);
    if (lane == 0) { red[warp] = best; redi[warp] = bi; }
    cbar(3);
    if (tid == 0) {
      for (int w = 1; w < MG_CWARPS; ++w)
        if (red[w] > best || (red[w] == best && redi[w] < bi)) { best = red[w]; bi = redi[w]; }
      p.tok_out[0] = bi;
    }
  }
}

}  // namespace

int init_decode_mega_kernels() {
  VCL_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MG_SMEM));
  return 0;
}

bool decode_mega_supported(int B, int D, int F, int V) {
  if (B != 1) return false;
  const int sms = device_num_sms();
  auto rmax = [&](int n) { return (((n + sms - 1) / sms) + 15) & ~15; };
  return D % 128 == 0 && F % 32 == 0 && D <= MG_KMAX && F <= MG_KMAX && rmax(V) <= MG_RMAX &&
         rmax(3 * D) <= MG_RMAX && rmax(2 * F) <= MG_RMAX;
}

int launch_decode_mega(const MegaParams& p, int B, cudaStream_t stream) {
  VCL_REQUIRE(decode_mega_supported(B, p.D, p.F, p.V), "decode megakernel: unsupported shape B=%d D=%d F=%d", B, p.D, p.F);
  VCL_REQUIRE(p.pos >= 0 && p.pos < p.s_max && p.pos + 1 <= MG_SPLIT * 128,
              "decode megakernel: position %d outside the supported range", p.pos);
  VCL_REQUIRE(p.H * MG_SPLIT <= MG_MAX_ITEMS * device_num_sms(), "decode megakernel: too many attention items");
  VCL_CUDA_OK(cudaMemsetAsync(p.barrier, 0, 32 * 17 * sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(device_num_sms());
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = MG_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  // VCL_MEGA_TRACE=<file>: record per-CTA phase timestamps of this step and dump them (debug aid,
  // eager launches only; layout [grid][L*6+1][8] uint64 nanoseconds)
  static unsigned long long* trace_buf = nullptr;
  const char* trace_path = getenv("VCL_MEGA_TRACE");
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  VCL_CUDA_OK(cudaStreamIsCapturing(stream, &cap));
  const bool tracing = trace_path != nullptr && cap == cudaStreamCaptureStatusNone;
  const size_t trace_n = (size_t)device_num_sms() * (p.L * 6 + 1) * 8;
  MegaParams q = p;
  static const int l2_slots = getenv("VCL_MEGA_L2_SLOTS") ? atoi(getenv("VCL_MEGA_L2_SLOTS")) : 0;   // measured: any L2 look-ahead is slower (86 -> 140 ms)
  q.l2_slots = l2_slots;
  if (tracing) {
    if (trace_buf == nullptr) VCL_CUDA_OK(cudaMalloc(&trace_buf, trace_n * sizeof(unsigned long long)));
    VCL_CUDA_OK(cudaMemsetAsync(trace_buf, 0, trace_n * sizeof(unsigned long long), stream));
    q.trace = trace_buf;
  }
  VCL_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_mega_kernel, q));
  count_launches(1);
  if (tracing) {
    std::vector<unsigned long long> host(trace_n);
    VCL_CUDA_OK(cudaStreamSynchronize(stream));
    VCL_CUDA_OK(cudaMemcpy(host.data(), trace_buf, trace_n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    FILE* f = fopen(trace_path, "wb");
    if (f != nullptr) { fwrite(host.data(), sizeof(unsigned long long), trace_n, f); fclose(f); }
  }
  return 0;
}

}  // namespace vcl
