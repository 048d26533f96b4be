// One cached decoding step (all layers + head + arg-max) as ONE persistent kernel.
//
// Decode at B <= 4 is pure weight streaming: 13.2 GB per step for the 7B model, 6 dependent phases
// per layer (qkv, attention, o_proj, gate/up, down) whose inputs are tiny vectors. Launching one
// kernel per phase leaves HBM idle for ~3 us at every one of the ~160 boundaries per step (ramp-up,
// tail, launch latency). Here every SM runs one persistent CTA:
//
//   warp 16 (producer)  streams this CTA's slice of EVERY weight matrix of the step, in order, into
//                       a shared-memory ring (6 x 32 KB) with cp.async.bulk + mbarrier
//                       complete_tx. Weights do not depend on activations, so the producer never
//                       waits for a phase to finish: it runs ahead across phase boundaries and HBM
//                       keeps streaming while the consumers sit in a grid barrier.
//   warps 0-15          consume the ring: K is dealt across the 512 lanes (activations live in
//                       registers), per-row partial sums meet in shared memory, the first threads
//                       run the fused epilogues (RMSNorm prologue, RoPE + KV append, SwiGLU,
//                       residual, logits), then a grid-wide barrier (global atomic) publishes the
//                       phase's output vector to all CTAs.
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

#include <stdlib.h>

#include <type_traits>

namespace vcl {

namespace {

constexpr int MG_CWARPS = 16;
constexpr int MG_CONSUMERS = MG_CWARPS * 32;       // 512
constexpr int MG_THREADS = MG_CONSUMERS + 32;      // + producer warp
constexpr int MG_SLOT_BYTES = 32768;
constexpr int MG_RMAX = 224;                       // max rows of one phase owned by a CTA
constexpr int MG_SPLIT = 4;                        // KV splits per (clip, head)
constexpr int MG_MAX_ITEMS = 4;                    // attention work items per CTA

enum { MODE_RES = 0, MODE_SWIGLU = 1, MODE_QKV = 2, MODE_LOGITS = 3 };

template <int NB>
struct MegaCfg {
  static constexpr int NSLOT = NB <= 2 ? 6 : 5;
  static constexpr int PART_FLOATS = MG_CWARPS * MG_RMAX * NB;
  static constexpr int SC_FLOATS = MG_MAX_ITEMS * 128;
  static constexpr int SMEM = NSLOT * MG_SLOT_BYTES + (PART_FLOATS + SC_FLOATS + 64) * 4 + 2 * NSLOT * 8 + 64;
};

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
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void cbar(int id) {      // barrier among the 512 consumer threads
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(MG_CONSUMERS) : "memory");
}
__device__ __forceinline__ float mdot8(const uint4& w, const uint4& x, float s) {
  s = fmaf(bf16lo(w.x), bf16lo(x.x), s); s = fmaf(bf16hi(w.x), bf16hi(x.x), s);
  s = fmaf(bf16lo(w.y), bf16lo(x.y), s); s = fmaf(bf16hi(w.y), bf16hi(x.y), s);
  s = fmaf(bf16lo(w.z), bf16lo(x.z), s); s = fmaf(bf16hi(w.z), bf16hi(x.z), s);
  s = fmaf(bf16lo(w.w), bf16lo(x.w), s); s = fmaf(bf16hi(w.w), bf16hi(x.w), s);
  return s;
}
__device__ __forceinline__ long long qkv_row(int v) {   // RoPE pairs (d, d+64) made adjacent
  return (long long)(v >> 7) * 128 + ((v & 127) >> 1) + (((v & 127) & 1) << 6);
}
__device__ __forceinline__ int rows_per_cta(int N) {
  int R = (N + gridDim.x - 1) / gridDim.x;
  return (R + 1) & ~1;
}
__device__ __forceinline__ int rows_per_slot(int K) {
  const int x = MG_SLOT_BYTES / (K * 2);
  return x >= 4 ? 4 : (x >= 2 ? 2 : 1);
}

template <int NB, int JD, int JF>
__global__ void __launch_bounds__(MG_THREADS, 1) decode_mega_kernel(const MegaParams p) {
  using Cfg = MegaCfg<NB>;
  constexpr int NSLOT = Cfg::NSLOT;
  extern __shared__ __align__(1024) uint8_t smem[];
  float* part = reinterpret_cast<float*>(smem + NSLOT * MG_SLOT_BYTES);
  float* sc = part + Cfg::PART_FLOATS;
  float* red = sc + Cfg::SC_FLOATS;
  const uint32_t ring0 = smem_u32(smem);
  const uint32_t bar0 = smem_u32(red + 64);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (NSLOT + s); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < NSLOT; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), MG_CWARPS); }
    mbar_fence_init();
  }
  __syncthreads();

  const int D = p.D, F = p.F, H = p.H, V = p.V, L = p.L;
  int slot = 0;
  uint32_t par = 0;
  auto advance = [&]() { if (++slot == NSLOT) { slot = 0; par ^= 1u; } };

  if (warp == MG_CWARPS) {
    // =============================== producer ===============================
    auto stream = [&](const bf16* W, int N, int K, bool qkv) {
      const int R = rows_per_cta(N);
      const int r0 = blockIdx.x * R, r1 = min(N, r0 + R);
      const int rps = rows_per_slot(K);
      const uint32_t rb = (uint32_t)K * 2u;
      for (int r = r0; r < r1; r += rps) {
        const int nr = min(rps, r1 - r);
        mbar_wait(empty_bar(slot), par ^ 1u);
        if (lane == 0) {
          mbar_arrive_expect_tx(full_bar(slot), nr * rb);
          const uint32_t dst = ring0 + slot * MG_SLOT_BYTES;
          if (!qkv) {
            bulk_g2s(dst, W + (long long)r * K, nr * rb, full_bar(slot));
          } else {
            for (int i = 0; i < nr; ++i) bulk_g2s(dst + i * rb, W + qkv_row(r + i) * K, rb, full_bar(slot));
          }
        }
        __syncwarp();
        advance();
      }
    };
    for (int l = 0; l < L; ++l) {
      const MegaLayer w = p.layers[l];
      stream(w.wqkv, 3 * D, D, true);
      stream(w.wo, D, D, false);
      stream(w.wgu, 2 * F, D, false);
      stream(w.wd, D, F, false);
    }
    stream(p.lm_head, V, D, false);
    return;
  }

  // =============================== consumers ===============================
  // Grid barrier, two levels to keep atomic contention low (148 atomics on one address serialise in
  // L2 for ~2 us): CTAs arrive on their group's counter (16 CTAs per group, one 128-byte line each);
  // the last arrival of a group bumps the top counter; everybody polls the top counter.
  // Counters are monotonic and zeroed by the launcher. release/acquire at gpu scope publish the
  // phase's global writes (made by other threads of the CTA before the CTA-level barrier).
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
      unsigned old;
      asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(grp_cnt) : "memory");
      if (old + 1 == group_size * epoch)
        asm volatile("red.add.release.gpu.global.u32 [%0], 1;" ::"l"(top_cnt) : "memory");
      while (ld_acquire_u32(top_cnt) < n_groups * epoch) {}
    }
    cbar(2);
  };

  // ---- one GEMV phase. XL(b, c) returns the 8 bf16 of chunk c of the input row b (packed) ----
  // Row -> "quad" mapping: a weight row is reduced by ONE group of 128 lanes (4 warps), each lane
  // owning JQ = ceil(K/1024) 16-byte chunks of it, so the warp-shuffle reduction is paid once per
  // 4*JQ chunks instead of once per chunk; consecutive rows go to the 4 quads round-robin.
  auto gemv = [&](auto jtag, int mode, int N, int K, const bf16* norm_w, auto XL, int l) {
    constexpr int JQ = decltype(jtag)::value;
    constexpr bool XF = (NB == 1 && JQ <= 5);      // activations held as fp32 (no unpack in the loop)
    const int nch = K >> 3;
    const int R = rows_per_cta(N);
    const int r0 = blockIdx.x * R, r1 = min(N, r0 + R);
    const int n_rows = max(0, r1 - r0);
    const int rps = rows_per_slot(K);
    const uint32_t rb = (uint32_t)K * 2u;
    const int tq = tid & 127, qd = tid >> 7, qw = warp & 3;
    // activations -> registers (RMS-normalised); every quad holds the whole vector
    uint4 xv[NB][JQ];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < JQ; ++j) {
        const int c = j * 128 + tq;
        xv[b][j] = (c < nch) ? XL(b, c) : make_uint4(0, 0, 0, 0);
        const uint4 u = xv[b][j];
        const float f0 = bf16lo(u.x), f1 = bf16hi(u.x), f2 = bf16lo(u.y), f3 = bf16hi(u.y);
        const float f4 = bf16lo(u.z), f5 = bf16hi(u.z), f6 = bf16lo(u.w), f7 = bf16hi(u.w);
        ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3 + f4 * f4 + f5 * f5 + f6 * f6 + f7 * f7;
      }
      if (norm_w != nullptr) {
        ss = warp_sum(ss);
        if (lane == 0) red[b * MG_CWARPS + warp] = ss;
      }
    }
    if (norm_w != nullptr) {
      cbar(3);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float tot = red[b * MG_CWARPS + qd * 4] + red[b * MG_CWARPS + qd * 4 + 1] +
                          red[b * MG_CWARPS + qd * 4 + 2] + red[b * MG_CWARPS + qd * 4 + 3];
        const float rstd = rsqrtf(tot / (float)K + p.eps);
#pragma unroll
        for (int j = 0; j < JQ; ++j) {
          const int c = j * 128 + tq;
          if (c < nch) {
            const uint4 u = xv[b][j];
            const uint4 g = *reinterpret_cast<const uint4*>(norm_w + c * 8);
            uint4 o;
            o.x = bf16x2_mul(g.x, pack_bf16x2(bf16lo(u.x) * rstd, bf16hi(u.x) * rstd));
            o.y = bf16x2_mul(g.y, pack_bf16x2(bf16lo(u.y) * rstd, bf16hi(u.y) * rstd));
            o.z = bf16x2_mul(g.z, pack_bf16x2(bf16lo(u.z) * rstd, bf16hi(u.z) * rstd));
            o.w = bf16x2_mul(g.w, pack_bf16x2(bf16lo(u.w) * rstd, bf16hi(u.w) * rstd));
            xv[b][j] = o;
          }
        }
      }
    }
    float xf[XF ? JQ : 1][8];
    if (XF) {
#pragma unroll
      for (int j = 0; j < (XF ? JQ : 1); ++j) {
        const uint4 u = xv[0][j];
        xf[j][0] = bf16lo(u.x); xf[j][1] = bf16hi(u.x); xf[j][2] = bf16lo(u.y); xf[j][3] = bf16hi(u.y);
        xf[j][4] = bf16lo(u.z); xf[j][5] = bf16hi(u.z); xf[j][6] = bf16lo(u.w); xf[j][7] = bf16hi(u.w);
      }
    }
    // stream the rows of this CTA out of the ring (every warp walks every slot in order; only the
    // quad that owns a row computes it)
    for (int r = 0; r < n_rows; r += rps) {
      const int nr = min(rps, n_rows - r);
      mbar_wait(full_bar(slot), par);
      const uint8_t* base = smem + slot * MG_SLOT_BYTES;
#pragma unroll 1
      for (int i = 0; i < nr; ++i) {
        if (((r + i) & 3) != qd) continue;
        const uint8_t* wrow = base + (size_t)i * rb + (size_t)tq * 16;
        float acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.f;
        constexpr int JB = 4;                      // weight chunks in flight per lane (register budget)
#pragma unroll
        for (int j0 = 0; j0 < JQ; j0 += JB) {
          uint4 wv[JB];
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            const int j = j0 + jj;
            wv[jj] = (j < JQ && j * 128 + tq < nch) ? *reinterpret_cast<const uint4*>(wrow + (size_t)j * 2048)
                                                     : make_uint4(0, 0, 0, 0);
          }
#pragma unroll
          for (int jj = 0; jj < JB; ++jj) {
            const int j = j0 + jj;
            if (j < JQ) {
              if (XF) {
                const uint4 w = wv[jj];
                float s = acc[0];
                s = fmaf(bf16lo(w.x), xf[XF ? j : 0][0], s); s = fmaf(bf16hi(w.x), xf[XF ? j : 0][1], s);
                s = fmaf(bf16lo(w.y), xf[XF ? j : 0][2], s); s = fmaf(bf16hi(w.y), xf[XF ? j : 0][3], s);
                s = fmaf(bf16lo(w.z), xf[XF ? j : 0][4], s); s = fmaf(bf16hi(w.z), xf[XF ? j : 0][5], s);
                s = fmaf(bf16lo(w.w), xf[XF ? j : 0][6], s); s = fmaf(bf16hi(w.w), xf[XF ? j : 0][7], s);
                acc[0] = s;
              } else {
#pragma unroll
                for (int b = 0; b < NB; ++b) acc[b] = mdot8(wv[jj], xv[b][j], acc[b]);
              }
            }
          }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float s = warp_sum(acc[b]);
          if (lane == 0) part[(qw * R + r + i) * NB + b] = s;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(slot));
      advance();
    }
    cbar(3);
    // epilogue
    const bool pairs = (mode == MODE_SWIGLU || mode == MODE_QKV);
    const int n_items = (pairs ? n_rows / 2 : n_rows) * NB;
    for (int it = tid; it < n_items; it += MG_CONSUMERS) {
      const int b = it % NB, u = it / NB;
      const int rr = pairs ? 2 * u : u;
      float v0 = 0.f, v1 = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        v0 += part[(w * R + rr) * NB + b];
        if (pairs) v1 += part[(w * R + rr + 1) * NB + b];
      }
      const int vrow = r0 + rr;
      if (mode == MODE_RES) {
        const float y = bf16r(v0) + ld_cg_bf16(p.h + (long long)b * D + vrow);
        p.h[(long long)b * D + vrow] = __float2bfloat16_rn(y);
      } else if (mode == MODE_LOGITS) {
        p.logits[(long long)b * V + vrow] = bf16r(v0);
      } else if (mode == MODE_SWIGLU) {
        const float g = bf16r(v0);
        const float sg = bf16r(__fdividef(g, 1.0f + __expf(-g)));
        p.act[(long long)b * F + (vrow >> 1)] = __float2bfloat16_rn(sg * bf16r(v1));
      } else {  // MODE_QKV
        const int hr = vrow >> 7;
        const int which = hr / H, head = hr - which * H;
        const int d = (vrow & 127) >> 1;
        const float lo = bf16r(v0), hi = bf16r(v1);
        const long long coff = (long long)l * p.cache_layer_elems + (((long long)b * H + head) * p.s_max + p.pos) * 128;
        if (which == 2) {
          p.vcache[coff + d] = __float2bfloat16_rn(lo);
          p.vcache[coff + d + 64] = __float2bfloat16_rn(hi);
        } else {
          const float c = __bfloat162float(p.cos_t[(long long)p.pos * 64 + d]);
          const float s = __bfloat162float(p.sin_t[(long long)p.pos * 64 + d]);
          const float olo = bf16r(lo * c) + bf16r(-hi * s);
          const float ohi = bf16r(hi * c) + bf16r(lo * s);
          if (which == 0) {
            p.q[(long long)b * D + head * 128 + d] = __float2bfloat16_rn(olo);
            p.q[(long long)b * D + head * 128 + d + 64] = __float2bfloat16_rn(ohi);
          } else {
            p.kcache[coff + d] = __float2bfloat16_rn(olo);
            p.kcache[coff + d + 64] = __float2bfloat16_rn(ohi);
          }
        }
      }
    }
  };

  const int kv_len = p.pos + 1;
  const int per = ((kv_len + MG_SPLIT - 1) / MG_SPLIT + 15) / 16 * 16;
  const int n_att = NB * H * MG_SPLIT;
  constexpr float LOG2E = 1.4426950408889634f;

  // ---- step start: h = embed[tok] ----
  for (int b = blockIdx.x; b < NB; b += gridDim.x) {
    int id = p.tok_in[(long long)b * p.tok_in_stride];
    id = id < 0 ? 0 : (id >= V ? V - 1 : id);
    for (int c = tid; c < (D >> 3); c += MG_CONSUMERS)
      *reinterpret_cast<uint4*>(p.h + (long long)b * D + c * 8) =
          *reinterpret_cast<const uint4*>(p.embed + (long long)id * D + c * 8);
  }
  grid_sync();

  auto x_from = [&](const bf16* src, int ld) {
    return [=](int b, int c) { return ld_cg_v4(src + (long long)b * ld + c * 8); };
  };
  std::integral_constant<int, JD> jd;
  std::integral_constant<int, JF> jf;

  for (int l = 0; l < L; ++l) {
    const MegaLayer w = p.layers[l];
    gemv(jd, MODE_QKV, 3 * D, D, w.ln1, x_from(p.h, D), l);
    grid_sync();
    // ---------------- attention A: scores + local statistics ----------------
    const bf16* kc_l = p.kcache + (long long)l * p.cache_layer_elems;
    const bf16* vc_l = p.vcache + (long long)l * p.cache_layer_elems;
    {
      int k = 0;
      for (int it = blockIdx.x; it < n_att; it += gridDim.x, ++k) {
        const int split = it % MG_SPLIT, bh = it / MG_SPLIT;
        const int head = bh % H, b = bh / H;
        const int lo = split * per, n_loc = max(0, min(kv_len - lo, per));
        const bf16* kc = kc_l + (((long long)b * H + head) * p.s_max + lo) * 128;
        float* s_it = sc + k * 128;
        const int kq = tid >> 4, dl = tid & 15;      // 32 keys per pass, 16 lanes per key
        const uint4 qu = ld_cg_v4(p.q + (long long)b * D + head * 128 + dl * 8);
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
        // local max and sum of exp(s - local max) (first warp)
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
    grid_sync();
    // ---------------- attention B: probabilities (global max / sum) x V ----------------
    {
      int k = 0;
      for (int it = blockIdx.x; it < n_att; it += gridDim.x, ++k) {
        const int split = it % MG_SPLIT, bh = it / MG_SPLIT;
        const int head = bh % H, b = bh / H;
        const int lo = split * per, n_loc = max(0, min(kv_len - lo, per));
        const bf16* vc = vc_l + (((long long)b * H + head) * p.s_max + lo) * 128;
        const float* s_it = sc + k * 128;
        float m = -INFINITY;
#pragma unroll
        for (int r = 0; r < MG_SPLIT; ++r) m = fmaxf(m, ld_cg_f32(p.att_stats + (bh * MG_SPLIT + r) * 2));
        float tot = 0.f;
#pragma unroll
        for (int r = 0; r < MG_SPLIT; ++r) {
          const float lm = ld_cg_f32(p.att_stats + (bh * MG_SPLIT + r) * 2);
          const float ls = ld_cg_f32(p.att_stats + (bh * MG_SPLIT + r) * 2 + 1);
          tot += (ls > 0.f) ? ls * exp2f((lm - m) * LOG2E) : 0.f;
        }
        const float inv = 1.0f / tot;
        const int g = tid >> 4, dc = tid & 15;       // 32 key groups x 16 dim chunks
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        uint4 vu[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = g + 32 * u;
          vu[u] = (j < n_loc) ? ld_cg_v4(vc + (long long)j * 128 + dc * 8) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int j = g + 32 * u;
          const float pj = (j < n_loc) ? bf16r(exp2f((s_it[j] - m) * LOG2E) * inv) : 0.f;
          acc[0] += pj * bf16lo(vu[u].x); acc[1] += pj * bf16hi(vu[u].x);
          acc[2] += pj * bf16lo(vu[u].y); acc[3] += pj * bf16hi(vu[u].y);
          acc[4] += pj * bf16lo(vu[u].z); acc[5] += pj * bf16hi(vu[u].z);
          acc[6] += pj * bf16lo(vu[u].w); acc[7] += pj * bf16hi(vu[u].w);
        }
        // the two key groups of a warp (lanes 0-15 / 16-31) first, then the 16 warps through smem
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
    grid_sync();
    // ---------------- o_proj: input = bf16(sum of the 4 partial attention outputs) ----------------
    auto x_att = [&](int b, int c) {
      const int head = c >> 4, d0 = (c & 15) * 8;
      const float* src = p.att_part + ((long long)(b * H + head) * MG_SPLIT) * 128 + d0;
      float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int r = 0; r < MG_SPLIT; ++r) {
        const uint4 a = ld_cg_v4(src + r * 128), bq = ld_cg_v4(src + r * 128 + 4);
        f[0] += __uint_as_float(a.x); f[1] += __uint_as_float(a.y); f[2] += __uint_as_float(a.z); f[3] += __uint_as_float(a.w);
        f[4] += __uint_as_float(bq.x); f[5] += __uint_as_float(bq.y); f[6] += __uint_as_float(bq.z); f[7] += __uint_as_float(bq.w);
      }
      return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
    };
    gemv(jd, MODE_RES, D, D, nullptr, x_att, l);
    grid_sync();
    gemv(jd, MODE_SWIGLU, 2 * F, D, w.ln2, x_from(p.h, D), l);
    grid_sync();
    gemv(jf, MODE_RES, D, F, nullptr, x_from(p.act, F), l);
    grid_sync();
  }
  gemv(jd, MODE_LOGITS, V, D, p.norm_w, x_from(p.h, D), 0);
  grid_sync();
  // ---- arg-max (lowest index wins ties), one CTA per clip ----
  for (int b = blockIdx.x; b < NB; b += gridDim.x) {
    const float* lg = p.logits + (long long)b * V;
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += MG_CONSUMERS) {
      const float x = ld_cg_f32(lg + i);
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
      p.tok_out[(long long)b * p.tok_out_stride] = bi;
    }
  }
}

template <int NB, int JD, int JF>
int launch_t(const MegaParams& p, cudaStream_t stream) {
  using Cfg = MegaCfg<NB>;
  VCL_CUDA_OK(cudaMemsetAsync(p.barrier, 0, 32 * 17 * sizeof(unsigned), stream));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(device_num_sms());
  cfg.blockDim = dim3(MG_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;   // all CTAs co-resident: the grid barrier cannot deadlock
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VCL_CUDA_OK(cudaLaunchKernelEx(&cfg, decode_mega_kernel<NB, JD, JF>, p));
  count_launches(1);
  return 0;
}

template <int JD, int JF>
int launch_nb(int B, const MegaParams& p, cudaStream_t stream) {
  switch (B) {
    case 1: return launch_t<1, JD, JF>(p, stream);
    case 2: return launch_t<2, JD, JF>(p, stream);
    case 3: return launch_t<3, JD, JF>(p, stream);
    case 4: return launch_t<4, JD, JF>(p, stream);
  }
  return -1;
}

template <int NB, int JD, int JF>
int init_t() {
  VCL_CUDA_OK(cudaFuncSetAttribute(decode_mega_kernel<NB, JD, JF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   MegaCfg<NB>::SMEM));
  return 0;
}
template <int JD, int JF>
int init_j() {
  if (init_t<1, JD, JF>() || init_t<2, JD, JF>() || init_t<3, JD, JF>() || init_t<4, JD, JF>()) return -2;
  return 0;
}

}  // namespace

int init_decode_mega_kernels() {
  if (init_j<1, 1>() || init_j<4, 11>() || init_j<5, 14>()) return -2;
  return 0;
}

bool decode_mega_supported(int B, int D, int F, int V) {
  if (B < 1 || B > 4) return false;
  const int jd = (D / 8 + 127) / 128, jf = (F / 8 + 127) / 128;     // chunks per lane of a 128-lane quad
  const bool combo = (jd == 1 && jf == 1) || (jd == 4 && jf == 11) || (jd == 5 && jf == 14);
  const int sms = device_num_sms();
  const int rmax = (((V > 3 * D ? V : 3 * D) + sms - 1) / sms + 1) & ~1;
  const int rgu = ((2 * F + sms - 1) / sms + 1) & ~1;
  return combo && rmax <= MG_RMAX && rgu <= MG_RMAX && D % 128 == 0;
}

int launch_decode_mega(const MegaParams& p, int B, cudaStream_t stream) {
  VCL_REQUIRE(decode_mega_supported(B, p.D, p.F, p.V), "decode megakernel: unsupported shape B=%d D=%d F=%d", B, p.D, p.F);
  VCL_REQUIRE(p.pos >= 0 && p.pos < p.s_max && p.pos + 1 <= MG_SPLIT * 128,
              "decode megakernel: position %d outside the supported range", p.pos);
  VCL_REQUIRE(B * p.H * MG_SPLIT <= MG_MAX_ITEMS * device_num_sms(), "decode megakernel: too many attention items");
  const int jd = (p.D / 8 + 127) / 128, jf = (p.F / 8 + 127) / 128;
  if (jd == 1 && jf == 1) return launch_nb<1, 1>(B, p, stream);
  if (jd == 4 && jf == 11) return launch_nb<4, 11>(B, p, stream);
  return launch_nb<5, 14>(B, p, stream);
}

}  // namespace vcl
