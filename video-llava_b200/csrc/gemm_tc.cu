// Dense bf16 GEMM on the 5th-generation tensor cores:  C[M,N] = epilogue(A[M,K] . W[N,K]^T)
//
// This one kernel serves every dense contraction on the hot path (SURVEY.md section 2.3 rows
// V1, V3, V5-V7, J1, L1, L4-L6): both operands are K-major exactly as nn.Linear stores them
// (activations [rows, K], weights [out, in]), so no transposes are ever materialised.
//
//   warp 0      TMA producer : cp.async.bulk.tensor 2-D tiles (128B swizzle) -> smem ring
//   warp 1      MMA issuer   : one thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) x4 per stage
//   warp 2      TMEM allocator (2 accumulator buffers of BLOCK_N fp32 columns)
//   warps 4-11  epilogue     : tcgen05.ld 32 lanes x 32 columns -> bias / activation / residual
//                              -> bf16 -> a 2 KB per-warp staging block in shared memory (a thread owns
//                              64 contiguous bytes of ONE row; the block is read back transposed so that
//                              a store instruction writes 8 rows x 64 contiguous bytes) -> global; two
//                              warps per TMEM lane quarter, software-pipelined over the column chunks
//
// Persistent: grid = min(#tiles, #SMs); the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the mainloop of tile i+1.
//
// Rounding points reproduce the reference's eager bf16 path (each nn.Linear output is rounded to
// bf16 before the following elementwise op):
//   CLIP quick_gelu  x*sigmoid(1.702x)   transformers/activations.py:117-123
//   CLIP residual    hidden + out        transformers/models/clip/modeling_clip.py:377,382
//   LLaMA SwiGLU     silu(gate)*up       transformers/models/llama/modeling_llama.py:182-184
//   LLaMA residual                       transformers/models/llama/modeling_llama.py:325,331
//   LLaMA RoPE + KV-cache write (ACT_ROPE, the prefill's q|k|v projection)
//                                        transformers/models/llama/modeling_llama.py:124-168
#include "common.cuh"
#include "kernels.h"

#include <cudaTypedefs.h>
#include <stdlib.h>

namespace vcl {

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int BLOCK_M = 128;
  static constexpr int BLOCK_K = 64;  // 64 bf16 = one 128-byte swizzle span
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int OUT_STAGE_BYTES = 8 * 2048;   // per epilogue warp: a 32 x 32 bf16 block on its way to global memory
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + OUT_STAGE_BYTES + 1024;  // +1024: manual align
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "tmem");
  static_assert((2 * STAGES + 4) * 8 + 8 <= BAR_BYTES, "barrier area");
};

__device__ __forceinline__ float act_quick_gelu(float x) {
  // three bf16 tensors are materialised by the reference: 1.702*x, sigmoid(.), x*sigmoid(.)
  x = bf16r(x);
  float t = bf16r(1.702f * x);
  float s = bf16r(1.0f / (1.0f + __expf(-t)));
  return x * s;
}
__device__ __forceinline__ float act_gelu_erf(float x) {
  x = bf16r(x);
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float act_silu(float g) {
  g = bf16r(g);
  return bf16r(g / (1.0f + __expf(-g)));
}

// ACT_ROPE epilogue (kernels.h: RopeEpilogue). A 128-column block of the q|k|v output is one head; the RoPE
// partner of column d is d + 64, which sits 64 TMEM columns further in the same lane, so a thread reads BOTH
// 32-column chunks of its pairs and the rotation is thread-local: with 128-wide tiles the two warps of a lane
// quarter split the head's 64 pairs (d in [32 hf, 32 hf + 32)), with 256-wide tiles each takes one head.
// Arithmetic = rope_kv_prefill_kernel (elementwise.cu): x rounded to bf16, every product rounded, fp32 sum, one
// final rounding (transformers/models/llama/modeling_llama.py:124-168).
template <int BLOCK_N, class WaitFn, class ArriveFn>
__device__ __forceinline__ void gemm_epilogue_rope(uint32_t tmem_acc, int q4, int hf, int lane, int m_blk, int n_blk,
                                                   bf16* C, long long ldc, int M, const RopeEpilogue& rp,
                                                   uint8_t* stg, WaitFn wait_full, ArriveFn arrive_empty) {
  if constexpr (BLOCK_N == 128 || BLOCK_N == 256) {
    constexpr int UNITS = BLOCK_N / 128;
    const int row = m_blk * 128 + q4 * 32 + lane;
    const bool row_ok = row < M;
    const int b = row / rp.S, s = row - b * rp.S;
    const int pos = rp.start_pos + s;
    wait_full();
    const uint32_t tl = tmem_acc + ((uint32_t)(q4 * 32) << 16);
#pragma unroll
    for (int u = 0; u < UNITS; ++u) {
      const int hh = BLOCK_N == 128 ? 0 : hf;                 // head inside the tile
      const int d0 = 32 * (BLOCK_N == 128 ? hf : u);          // first pair of this unit
      uint32_t lo[32], hi[32];
      __syncwarp();
      tmem_ld_32x32(tl + hh * 128 + d0, lo);
      tmem_ld_32x32(tl + hh * 128 + 64 + d0, hi);
      tc_wait_ld();
      if (u == UNITS - 1) arrive_empty();                     // every TMEM read of this accumulator has completed
      const int gh = n_blk * UNITS + hh;                      // head index over q | k | v (the same for the warp)
      const int which = gh / rp.H, head = gh - which * rp.H;
      uint32_t olo[16], ohi[16];
      if (row_ok && which == 2) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          olo[j] = pack_bf16x2(__uint_as_float(lo[2 * j]), __uint_as_float(lo[2 * j + 1]));
          ohi[j] = pack_bf16x2(__uint_as_float(hi[2 * j]), __uint_as_float(hi[2 * j + 1]));
        }
      } else if (row_ok) {
        const bf16* ct = rp.cos_t + (long long)pos * 64 + d0;
        const bf16* st = rp.sin_t + (long long)pos * 64 + d0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(ct + q * 8));
          const uint4 s4 = __ldg(reinterpret_cast<const uint4*>(st + q * 8));
          const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w}, sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = q * 4 + e;                        // pair of columns 2j, 2j + 1
            const uint32_t l2 = pack_bf16x2(__uint_as_float(lo[2 * j]), __uint_as_float(lo[2 * j + 1]));
            const uint32_t h2 = pack_bf16x2(__uint_as_float(hi[2 * j]), __uint_as_float(hi[2 * j + 1]));
            const float l0 = bf16lo(l2), l1 = bf16hi(l2), h0 = bf16lo(h2), h1 = bf16hi(h2);
            const float c0 = bf16lo(cw[e]), c1 = bf16hi(cw[e]), s0 = bf16lo(sw[e]), s1 = bf16hi(sw[e]);
            olo[j] = pack_bf16x2(bf16r(l0 * c0) + bf16r(-h0 * s0), bf16r(l1 * c1) + bf16r(-h1 * s1));
            ohi[j] = pack_bf16x2(bf16r(h0 * c0) + bf16r(l0 * s0), bf16r(h1 * c1) + bf16r(l1 * s1));
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) { olo[j] = 0u; ohi[j] = 0u; }
      }
      // both 64-byte halves of the row leave through the warp's staging block (gemm_epilogue_tile): 8 rows x 64
      // contiguous bytes per store instruction; the destination row is recomputed for the row a lane stores
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int swz = (lane >> 1) & 3;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t* src = half == 0 ? olo : ohi;
          *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ swz) << 4)) =
              make_uint4(src[4 * q], src[4 * q + 1], src[4 * q + 2], src[4 * q + 3]);
        }
        __syncwarp();
        const int r0 = lane >> 2, cc = lane & 3;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int r = 8 * k + r0;
          const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
          const int grow = m_blk * 128 + q4 * 32 + r;
          if (grow < M) {
            const int gb = grow / rp.S, gs = grow - gb * rp.S;
            bf16* dst = which == 0
                ? C + (long long)grow * ldc + gh * 128 + d0
                : (which == 1 ? rp.kcache : rp.vcache) + (((long long)gb * rp.H + head) * rp.s_max + rp.start_pos + gs) * 128 + d0;
            *reinterpret_cast<uint4*>(dst + half * 64 + cc * 8) = val;
          }
        }
        __syncwarp();
      }
    }
  } else {
    wait_full();               // narrower tiles are never launched with ACT_ROPE (launch_gemm_bf16_tn)
    arrive_empty();
  }
}

// One output tile of the epilogue, for the 32 rows x (BLOCK_N / 2) columns this warp owns: TMEM -> bias /
// activation / residual -> bf16 -> global. warp w reads TMEM lanes 32*(w%4).. (hardware restriction) and
// owns one half (hf) of the tile's 32-column chunks; the tcgen05.ld and the residual loads of chunk i+1
// are issued before the math of chunk i so their latency hides behind it. wait_full() blocks until the
// accumulator is complete, arrive_empty() hands it back to the MMA issuer once it has been read out.
template <int BLOCK_N, int ACT, class WaitFn, class ArriveFn>
__device__ __forceinline__ void gemm_epilogue_tile(uint32_t tmem_acc, int q4, int hf, int lane, int m_blk, int n_blk,
                                                   bf16* C, long long ldc, const bf16* __restrict__ bias,
                                                   const bf16* residual, long long ldr, int M, int N,
                                                   const RopeEpilogue& rope, uint8_t* stg, WaitFn wait_full,
                                                   ArriveFn arrive_empty) {
  if constexpr (ACT == ACT_ROPE) {
    gemm_epilogue_rope<BLOCK_N>(tmem_acc, q4, hf, lane, m_blk, n_blk, C, ldc, M, rope, stg, wait_full, arrive_empty);
    return;
  }
  constexpr int CH = BLOCK_N / 32;
  constexpr int PER = (CH + 1) / 2;
  const int c_begin = hf * PER;
  const int n_mine = (c_begin + PER <= CH) ? PER : (CH > c_begin ? CH - c_begin : 0);
  const int row = m_blk * 128 + q4 * 32 + lane;
  const bool row_ok = row < M;
  const int colbase = n_blk * BLOCK_N + c_begin * 32;
  const bf16* rrow = residual + (long long)row * ldr + colbase;
  const bool has_res = (residual != nullptr) && row_ok;
  uint32_t v[2][32];
  uint4 rr[2][4];
  if (has_res && n_mine > 0 && colbase < N) {
#pragma unroll
    for (int q = 0; q < 4; ++q) rr[0][q] = *reinterpret_cast<const uint4*>(rrow + q * 8);
  }
  wait_full();
  const uint32_t taddr = tmem_acc + ((uint32_t)(q4 * 32) << 16) + c_begin * 32;
  __syncwarp();
  if (n_mine > 0) {
    tmem_ld_32x32(taddr, v[0]);
    tc_wait_ld();
  }
  if (n_mine <= 1) {
    arrive_empty();
  }
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    if (i < n_mine) {
      const int cur = i & 1, nxt = cur ^ 1;
      const int col0 = colbase + i * 32;
      if (i + 1 < n_mine) {
        if (has_res && col0 + 32 < N) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            rr[nxt][q] = *reinterpret_cast<const uint4*>(rrow + (i + 1) * 32 + q * 8);
        }
        __syncwarp();
        tmem_ld_32x32(taddr + (i + 1) * 32, v[nxt]);
      }
      auto load_f = [&](float (&f)[32]) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[cur][j]);
        if (bias != nullptr) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 b = *reinterpret_cast<const uint4*>(bias + col0 + q * 8);
            f[q * 8 + 0] += bf16lo(b.x); f[q * 8 + 1] += bf16hi(b.x);
            f[q * 8 + 2] += bf16lo(b.y); f[q * 8 + 3] += bf16hi(b.y);
            f[q * 8 + 4] += bf16lo(b.z); f[q * 8 + 5] += bf16hi(b.z);
            f[q * 8 + 6] += bf16lo(b.w); f[q * 8 + 7] += bf16hi(b.w);
          }
        }
      };
      if constexpr (ACT == ACT_SWIGLU) {
        if (col0 < N) {                            // warp-uniform
          // weight rows are interleaved (2j = gate_j, 2j+1 = up_j): 32 columns -> 16 outputs (32 bytes per row)
          uint32_t o[8];
          if (row_ok) {
            float f[32];
            load_f(f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t g2 = pack_bf16x2(f[4 * j + 0], f[4 * j + 2]);   // gate pair (bf16)
              const uint32_t u2 = pack_bf16x2(f[4 * j + 1], f[4 * j + 3]);   // up pair (bf16)
              const float g0 = bf16lo(g2), g1 = bf16hi(g2);
              const uint32_t s2 = pack_bf16x2(__fdividef(g0, 1.0f + __expf(-g0)),
                                              __fdividef(g1, 1.0f + __expf(-g1)));   // silu (bf16)
              o[j] = bf16x2_mul(s2, u2);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = 0u;
          }
          // staged like the plain path below: 16 rows x 32 contiguous bytes (one sector each) per instruction
          const int sw = (lane >> 2) & 1;
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((0 ^ sw) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<uint4*>(stg + lane * 32 + ((1 ^ sw) << 4)) = make_uint4(o[4], o[5], o[6], o[7]);
          __syncwarp();
          const int r0 = lane >> 1, cc = lane & 1;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int r = 16 * k + r0;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 32 + ((cc ^ ((r >> 2) & 1)) << 4));
            const int grow = m_blk * 128 + q4 * 32 + r;
            if (grow < M) *reinterpret_cast<uint4*>(C + (long long)grow * ldc + (col0 >> 1) + cc * 8) = val;
          }
          __syncwarp();
        }
      } else if (col0 < N) {                       // warp-uniform: every lane takes part in the staged store
        uint32_t o[16];
        if (row_ok) {
          float f[32];
          load_f(f);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if constexpr (ACT == ACT_QGELU) {
              // x*sigmoid(1.702x): the three bf16 tensors of the reference are materialised
              const uint32_t x2 = pack_bf16x2(f[2 * j], f[2 * j + 1]);
              const uint32_t t2 = pack_bf16x2(1.702f * bf16lo(x2), 1.702f * bf16hi(x2));
              const uint32_t s2 = pack_bf16x2(__fdividef(1.0f, 1.0f + __expf(-bf16lo(t2))),
                                              __fdividef(1.0f, 1.0f + __expf(-bf16hi(t2))));
              o[j] = bf16x2_mul(x2, s2);
            } else if constexpr (ACT == ACT_GELU) {
              o[j] = pack_bf16x2(act_gelu_erf(f[2 * j]), act_gelu_erf(f[2 * j + 1]));
            } else {
              o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
            }
          }
          if (has_res) {
            // bf16(linear output) + residual, one more bf16 rounding (packed bf16x2 add)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              o[q * 4 + 0] = bf16x2_add(o[q * 4 + 0], rr[cur][q].x);
              o[q * 4 + 1] = bf16x2_add(o[q * 4 + 1], rr[cur][q].y);
              o[q * 4 + 2] = bf16x2_add(o[q * 4 + 2], rr[cur][q].z);
              o[q * 4 + 3] = bf16x2_add(o[q * 4 + 3], rr[cur][q].w);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = 0u;
        }
        // A thread holds 64 contiguous bytes of ONE row: storing them directly makes every store instruction
        // touch 32 rows with 16 bytes each (half sectors; the stores of a 128 x 256 tile cost ~20 % of the ViT
        // GEMMs: 130 -> 102 us without them). The 32 x 32 block goes through this warp's 2 KB of shared memory
        // (XOR-swizzled, conflict-free both ways) and leaves as 8 rows x 64 contiguous bytes per instruction.
        {
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((q ^ sw) << 4)) =
                make_uint4(o[q * 4 + 0], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
          __syncwarp();
          const int r0 = lane >> 2, cc = lane & 3;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int r = 8 * k + r0;
            const uint4 val = *reinterpret_cast<const uint4*>(stg + r * 64 + ((cc ^ ((r >> 1) & 3)) << 4));
            const int grow = m_blk * 128 + q4 * 32 + r;
            if (grow < M) *reinterpret_cast<uint4*>(C + (long long)grow * ldc + col0 + cc * 8) = val;
          }
          __syncwarp();
        }
      }
      if (i + 1 < n_mine) {
        __syncwarp();
        tc_wait_ld();
        if (i + 2 >= n_mine) {
          // every TMEM read of this accumulator has completed: hand it back to the MMA warp
          arrive_empty();
        }
      }
    }
  }
}

// CL = thread-block-cluster size along M (1, 2 or 4). The CL CTAs of a cluster work on CL
// consecutive M tiles of the SAME N tile: every CTA fetches 1/CL of the weight tile and TMA-multicasts
// it into all CL shared memories, so a weight byte crosses L2->SM once per cluster instead of once
// per CTA (the 128x256 tile is L2-bandwidth-limited otherwise, above all for the short-M prefill).
template <int BLOCK_N, int ACT, int CL>
__global__ void __launch_bounds__(384, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_b, bf16* C, long long ldc,
                    const bf16* __restrict__ bias, const bf16* residual, long long ldr, int M,
                    int N, int K, int pf_ahead, const RopeEpilogue rope) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;                 // 1024-byte aligned (SWIZZLE_128B atoms)
  const uint32_t smem_base = raw_addr + pad;

  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), CL);      // one tcgen05.commit arrival from every CTA of the cluster
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 256);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), Cfg::TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_m = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;
  // cluster-tile = (N tile, group of CL M tiles); this CTA takes M tile `group*CL + rank`
  const int cta_rank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int num_tiles = ((num_m + CL - 1) / CL) * num_n;
  constexpr uint16_t MC_MASK = (uint16_t)((1u << CL) - 1u);
  constexpr int B_SLICE_ROWS = BLOCK_N / CL;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // L2 look-ahead for the weight operand: with few row tiles (prefill, M = 448) the weights come from
      // HBM and the ring alone holds few of this CTA's own weight bytes in flight (4 stages x an 8 KB multicast
      // slice). A prefetch cursor runs pf_ahead k-blocks in front of the loads, across tile boundaries; the
      // launcher turns it on only where it measured faster (weight_prefetch_depth).
      int pf_tile = cluster_id, pf_kb = 0;
      auto prefetch_next = [&]() {
        if (pf_tile < num_tiles) {
          tma_prefetch_2d(&tmap_b, pf_kb * Cfg::BLOCK_K, (pf_tile % num_n) * BLOCK_N + cta_rank * B_SLICE_ROWS);
          if (++pf_kb == num_k) { pf_kb = 0; pf_tile += n_clusters; }
        }
      };
      for (int i = 0; i < pf_ahead; ++i) prefetch_next();
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_blk = (tile / num_n) * CL + cta_rank, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb) {
          if (pf_ahead > 0) prefetch_next();
          mbar_wait(empty_bar(stage), phase ^ 1u);    // slot free in EVERY CTA of the cluster
          mbar_arrive_expect_tx(full_bar(stage), Cfg::STAGE_BYTES);
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_2d(a_dst, &tmap_a, full_bar(stage), kb * Cfg::BLOCK_K, m_blk * Cfg::BLOCK_M);
          if (CL == 1) {
            tma_load_2d(a_dst + Cfg::A_BYTES, &tmap_b, full_bar(stage), kb * Cfg::BLOCK_K,
                        n_blk * BLOCK_N);
          } else {
            // my slice of the weight tile, delivered to the same offset in all CL shared memories
            tma_load_2d_mc(a_dst + Cfg::A_BYTES + cta_rank * B_SLICE_ROWS * 128, &tmap_b,
                           full_bar(stage), kb * Cfg::BLOCK_K, n_blk * BLOCK_N + cta_rank * B_SLICE_ROWS,
                           MC_MASK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(Cfg::BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);   // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb) {
          mbar_wait(full_bar(stage), phase);          // TMA bytes have landed
          tc_fence_after();
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < Cfg::BLOCK_K / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the swizzle span: +2 in the (addr >> 4) field
            tc_mma_bf16(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // smem slot free once these MMAs retire; with a cluster, tell every CTA that multicasts here
          if (CL == 1) tc_commit(empty_bar(stage));
          else tc_commit_mc(empty_bar(stage), MC_MASK);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit(tfull_bar(acc));                    // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (8 warps) ------------------------------
    // warp w reads TMEM lanes 32*(w%4).. (hardware restriction) and owns one half of the tile's
    // 32-column chunks; the tcgen05.ld and the residual loads of chunk i+1 are issued before the
    // math of chunk i so their latency hides behind it.
    const int q4 = warp & 3;
    const int hf = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int m_blk = (tile / num_n) * CL + cta_rank, n_blk = tile % num_n;
      gemm_epilogue_tile<BLOCK_N, ACT>(
          tmem_base + acc * BLOCK_N, q4, hf, lane, m_blk, n_blk, C, ldc, bias, residual, ldr, M, N, rope,
          smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + (warp - 4) * 2048,
          [&]() { mbar_wait(tfull_bar(acc), acc_phase); tc_fence_after(); },
          [&]() { tc_fence_before(); mbar_arrive(tempty_bar(acc)); });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();      // nobody exits while a peer may still multicast / signal into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): the two CTAs of a cluster work on ONE 256 x BLOCK_N tile. Each
// CTA stages its own 128 rows of A and only HALF of the weight tile (BLOCK_N / 2 rows); the leader CTA
// issues M = 256 MMAs that read both shared memories, and each CTA finds its 128 accumulator rows in its
// own TMEM. Against the multicast clusters above (same L2 traffic) this halves the shared-memory footprint
// and the shared-memory reads of the weight operand per CTA: 32 KB stages, six of them.
//   full barrier   lives in the leader, which expects the TMA bytes of BOTH CTAs; the peer's loads complete on
//                  it too but the peer does not arrive (a remote mbarrier.arrive.release.cluster per stage
//                  cost the peer's producer ~0.5 us and paced the whole kernel at 0.75 us per k-block).
//                  The peer cannot run a ring turn ahead: its slot is released by the MMAs of the previous
//                  turn, which waited for that turn's phase of this barrier
//   empty barrier  in both CTAs, released by the leader's tcgen05.commit (multicast to the pair)
//   tmem full      in both CTAs (multicast commit); tmem empty in the leader, 2 x 256 epilogue arrivals
// ---------------------------------------------------------------------------------------------
template <int BLOCK_N>
struct Gemm2Cfg {
  static constexpr int STAGES = 6;
  static constexpr int A_BYTES = 128 * 64 * 2;
  static constexpr int B_BYTES = (BLOCK_N / 2) * 64 * 2;     // this CTA's half of the weight tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int OUT_STAGE_BYTES = 8 * 2048;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + OUT_STAGE_BYTES + 1024;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static_assert(BLOCK_N == 256 || BLOCK_N == 128, "pair tiles are 256 x 256 or 256 x 128");
};

template <int BLOCK_N, int ACT>
__global__ void __launch_bounds__(384, 1)
gemm2_bf16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, bf16* C,
                     long long ldc, const bf16* __restrict__ bias, const bf16* residual, long long ldr, int M, int N,
                     int K, int pf_ahead, unsigned long long* __restrict__ trace, const RopeEpilogue rope) {
  // debug timeline (tools/gemm_pair_trace.py): [cta < 2][k-block < 128, counted over the tiles][4] globaltimer stamps; null in production
  auto stamp = [&](int kbi, int ev) {
    if (trace != nullptr && blockIdx.x < 2 && kbi < 128) {
      unsigned long long t_;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));
      trace[((size_t)blockIdx.x * 128 + kbi) * 4 + ev] = t_;
    }
  };
  using Cfg = Gemm2Cfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t pad = ((raw_addr + 1023u) & ~1023u) - raw_addr;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_addr + pad;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  volatile uint32_t* tmem_ptr_smem =
      reinterpret_cast<volatile uint32_t*>(smem + STAGES * Cfg::STAGE_BYTES + 8 * (2 * STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                  // 0 = leader: issues the MMAs, owns the full barriers
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);            // the leader's arrive.expect_tx (bytes of both CTAs)
      mbar_init(empty_bar(s), 1);           // one multicast commit
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 512);        // the epilogue threads of both CTAs
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc2(smem_u32(const_cast<uint32_t*>(tmem_ptr_smem)), Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int num_pairs = (M + 255) / 256;
  const int num_n = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_k = (K + 63) / 64;
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const int num_tiles = num_pairs * num_n;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs) ------------------------------
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int pf_tile = cluster_id, pf_kb = 0;                   // L2 look-ahead for this CTA's half of the weight tiles
      auto prefetch_next = [&]() {
        if (pf_tile < num_tiles) {
          tma_prefetch_2d(&tmap_b, pf_kb * 64, (pf_tile % num_n) * BLOCK_N + (int)rank * (BLOCK_N / 2));
          if (++pf_kb == num_k) { pf_kb = 0; pf_tile += n_clusters; }
        }
      };
      for (int i = 0; i < pf_ahead; ++i) prefetch_next();
      int gkb = 0;                                           // k-blocks issued by this CTA so far (trace index)
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        const int m_blk = (tile / num_n) * 2 + (int)rank, n_blk = tile % num_n;
        for (int kb = 0; kb < num_k; ++kb, ++gkb) {
          if (pf_ahead > 0) prefetch_next();
          stamp(gkb, 0);
          mbar_wait_safe(empty_bar(stage), phase ^ 1u);
          stamp(gkb, 1);
          const uint32_t lead_full = mapa_cluster(full_bar(stage), 0);
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * Cfg::STAGE_BYTES);   // the bytes of BOTH CTAs
          const uint32_t a_dst = smem_base + stage * Cfg::STAGE_BYTES;
          tma_load_2d_2cta(a_dst, &tmap_a, lead_full, kb * 64, m_blk * 128);
          tma_load_2d_2cta(a_dst + Cfg::A_BYTES, &tmap_b, lead_full, kb * 64, n_blk * BLOCK_N + (int)rank * (BLOCK_N / 2));
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader only) ------------------------------
    if (lane == 0 && rank == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(256, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int gkb = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
        mbar_wait_safe(tempty_bar(acc), acc_phase ^ 1u);     // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < num_k; ++kb, ++gkb) {
          stamp(gkb, 2);
          mbar_wait_safe(full_bar(stage), phase);
          tc_fence_after();
          stamp(gkb, 3);
          const uint32_t a_addr = smem_base + stage * Cfg::STAGE_BYTES;
          const uint64_t a_desc = umma_desc_k_sw128(a_addr);
          const uint64_t b_desc = umma_desc_k_sw128(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc_mma_bf16_2cta(d_tmem, a_desc + 2u * k, b_desc + 2u * k, idesc, (kb | k) != 0 ? 1u : 0u);
          tc_commit_2cta(empty_bar(stage), 3);               // the slot is free in both CTAs
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        tc_commit_2cta(tfull_bar(acc), 3);                   // accumulator complete -> both epilogues
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    // ------------------------------ epilogue (8 warps per CTA, own 128 rows) ------------------------------
    const int q4 = warp & 3;
    const int hf = (warp - 4) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += n_clusters) {
      const int m_blk = (tile / num_n) * 2 + (int)rank, n_blk = tile % num_n;
      const uint32_t lead_tempty = mapa_cluster(tempty_bar(acc), 0);
      gemm_epilogue_tile<BLOCK_N, ACT>(
          tmem_base + acc * BLOCK_N, q4, hf, lane, m_blk, n_blk, C, ldc, bias, residual, ldr, M, N, rope,
          smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + (warp - 4) * 2048,
          [&]() { mbar_wait_safe(tfull_bar(acc), acc_phase); tc_fence_after(); },
          [&]() { tc_fence_before(); mbar_arrive_cluster(lead_tempty); });
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                  // nobody exits (or frees TMEM) while the peer may still read / signal it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static unsigned long long* g_gemm_trace = nullptr;     // debug: vcl_debug_set_gemm_trace

static int resolve_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_last_error("cuTensorMapEncodeTiled is not available from the driver (%s)",
                   cudaGetErrorString(e));
    return -2;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

// 2-D bf16 tensor [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle.
int make_tmap_2d(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld,
                 int box_rows) {
  if (resolve_encode() != 0) return -2;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%lld cols=%lld ld=%lld box=%d",
                   (int)r, ptr, rows, cols, ld, box_rows);
    return -2;
  }
  return 0;
}

static int g_num_sms = 0;
int device_num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// k-blocks of L2 look-ahead for the weight operand (see the producer), measured on the prefill shapes
// (M = 448, tools/sweep_gemm.py with VCL_GEMM_PF = 0 / 8 / 16 / 32): it pays only where one CTA of a 4-cluster
// streams its quarter of a 256-wide weight tile -- gate|up 95 -> 88 us at depth 8, q|k|v 71 -> 65 us (level with
// the 128-wide tiles it then uses anyway) -- is neutral for the CTA pairs and costs 5-10 % where every row tile
// prefetches the same weights for itself (no cluster). VCL_GEMM_PF=<k-blocks> overrides (0 = off) for A/B runs.
// The ACTIVATION operand of the ViT GEMMs was tried too: the per-k-block timeline (tools/gemm_pair_trace.py,
// PAIR_SHAPE=25700,3072,1024) shows a ~2.9 us bubble at every tile start (the first loads of rows nobody has touched
// take 2.6-4.4 us against 0.33 us per k-block), yet pulling the next tile's rows into L2 ahead of time made every
// shape SLOWER, through the TMA (q|k|v 132 -> 144 us, fc2 153 -> 188) as well as with plain prefetch.global.L2
// issued by the idle warp (128 -> 143, 153 -> 172). Not kept.
static int weight_prefetch_depth(const GemmArgs& g, int block_n, int cl) {
  static const int forced = getenv("VCL_GEMM_PF") ? atoi(getenv("VCL_GEMM_PF")) : -1;
  if (forced >= 0) return forced;
  return (cl == 4 && block_n == 256 && (g.M + 127) / 128 <= 16) ? 8 : 0;
}

template <int BLOCK_N, int ACT>
static int launch_pair(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = Gemm2Cfg<BLOCK_N>;
  auto kern = gemm2_bf16_tn_kernel<BLOCK_N, ACT>;
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, g.A, g.M, g.K, g.lda, 128) != 0) return -2;
  if (make_tmap_2d(&tb, g.W, g.N, g.K, g.ldw, BLOCK_N / 2) != 0) return -2;
  const int num_tiles = ((g.M + 255) / 256) * ((g.N + BLOCK_N - 1) / BLOCK_N);
  int clusters = device_num_sms() / 2;
  if (clusters > num_tiles) clusters = num_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * 2);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VCL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, g.C, (long long)g.ldc, g.bias, g.residual, (long long)g.ldr, g.M,
                                 g.N, g.K, weight_prefetch_depth(g, BLOCK_N, -2), g_gemm_trace, g.rope));
  count_launches(1);
  return 0;
}

template <int BLOCK_N>
static int launch_pair_act(const GemmArgs& g, cudaStream_t stream) {
  switch (g.act) {
    case ACT_NONE: return launch_pair<BLOCK_N, ACT_NONE>(g, stream);
    case ACT_QGELU: return launch_pair<BLOCK_N, ACT_QGELU>(g, stream);
    case ACT_GELU: return launch_pair<BLOCK_N, ACT_GELU>(g, stream);
    case ACT_SWIGLU: return launch_pair<BLOCK_N, ACT_SWIGLU>(g, stream);
    case ACT_ROPE: return launch_pair<BLOCK_N, ACT_ROPE>(g, stream);
  }
  set_last_error("gemm: unknown activation %d", g.act);
  return -1;
}

template <int BLOCK_N, int ACT, int CL>
static int launch_one(const GemmArgs& g, cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_bf16_tn_kernel<BLOCK_N, ACT, CL>;
  CUtensorMap ta, tb;
  if (make_tmap_2d(&ta, g.A, g.M, g.K, g.lda, Cfg::BLOCK_M) != 0) return -2;
  if (make_tmap_2d(&tb, g.W, g.N, g.K, g.ldw, BLOCK_N / CL) != 0) return -2;
  const int num_m = (g.M + 127) / 128;
  const int num_tiles = ((num_m + CL - 1) / CL) * ((g.N + BLOCK_N - 1) / BLOCK_N);   // cluster-tiles
  int clusters = device_num_sms() / CL;
  if (CL == 4) clusters -= 4;            // GPC boundaries strand ~16 SMs for 4-CTA clusters
  if (clusters > num_tiles) clusters = num_tiles;
  if (g.max_ctas > 0 && clusters * CL > g.max_ctas) clusters = g.max_ctas / CL > 0 ? g.max_ctas / CL : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CL);
  cfg.blockDim = dim3(384);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CL > 1 ? 1 : 0;
  VCL_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, g.C, (long long)g.ldc, g.bias, g.residual,
                                 (long long)g.ldr, g.M, g.N, g.K, weight_prefetch_depth(g, BLOCK_N, CL), g.rope));
  count_launches(1);
  return 0;
}

template <int BLOCK_N, int ACT>
static int launch_cl(const GemmArgs& g, int cl, cudaStream_t stream) {
  if (BLOCK_N >= 128) {   // multicast slices must stay whole 8-row swizzle groups and >= 32 rows
    if (cl == 4) return launch_one<BLOCK_N, ACT, (BLOCK_N >= 128 ? 4 : 1)>(g, stream);
    if (cl == 2) return launch_one<BLOCK_N, ACT, (BLOCK_N >= 128 ? 2 : 1)>(g, stream);
  }
  return launch_one<BLOCK_N, ACT, 1>(g, stream);
}

template <int BLOCK_N>
static int launch_act(const GemmArgs& g, int cl, cudaStream_t stream) {
  switch (g.act) {
    case ACT_NONE: return launch_cl<BLOCK_N, ACT_NONE>(g, cl, stream);
    case ACT_QGELU: return launch_cl<BLOCK_N, ACT_QGELU>(g, cl, stream);
    case ACT_GELU: return launch_cl<BLOCK_N, ACT_GELU>(g, cl, stream);
    case ACT_SWIGLU: return launch_cl<BLOCK_N, ACT_SWIGLU>(g, cl, stream);
    case ACT_ROPE:
      if constexpr (BLOCK_N >= 128) return launch_cl<BLOCK_N, ACT_ROPE>(g, cl, stream);
      break;
  }
  set_last_error("gemm: unknown activation %d", g.act);
  return -1;
}

template <int BLOCK_N, int ACT>
static int init_one() {
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, ACT, 1>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   GemmCfg<BLOCK_N>::SMEM_BYTES));
  if (BLOCK_N >= 128) {
    VCL_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, ACT, (BLOCK_N >= 128 ? 2 : 1)>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GemmCfg<BLOCK_N>::SMEM_BYTES));
    VCL_CUDA_OK(cudaFuncSetAttribute(gemm_bf16_tn_kernel<BLOCK_N, ACT, (BLOCK_N >= 128 ? 4 : 1)>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     GemmCfg<BLOCK_N>::SMEM_BYTES));
  }
  return 0;
}
template <int BLOCK_N>
static int init_bn() {
  if (init_one<BLOCK_N, ACT_NONE>() || init_one<BLOCK_N, ACT_QGELU>() ||
      init_one<BLOCK_N, ACT_GELU>() || init_one<BLOCK_N, ACT_SWIGLU>()) return -2;
  if constexpr (BLOCK_N >= 128) {
    if (init_one<BLOCK_N, ACT_ROPE>()) return -2;
  }
  return 0;
}
// Opt every instantiation into its dynamic shared memory size (done once, outside any capture).
template <int BLOCK_N>
static int init_pair() {
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BLOCK_N, ACT_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BLOCK_N>::SMEM_BYTES));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BLOCK_N, ACT_QGELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BLOCK_N>::SMEM_BYTES));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BLOCK_N, ACT_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BLOCK_N>::SMEM_BYTES));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BLOCK_N, ACT_SWIGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BLOCK_N>::SMEM_BYTES));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemm2_bf16_tn_kernel<BLOCK_N, ACT_ROPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Gemm2Cfg<BLOCK_N>::SMEM_BYTES));
  return 0;
}

int init_gemm_kernels() {
  if (resolve_encode() != 0) return -2;
  if (init_pair<256>() || init_pair<128>()) return -2;
  if (init_bn<256>() || init_bn<128>() || init_bn<64>() || init_bn<32>()) return -2;
  return 0;
}

extern "C" void vcl_debug_set_gemm_trace(void* dev_buffer) {
  g_gemm_trace = reinterpret_cast<unsigned long long*>(dev_buffer);
}

int launch_gemm_bf16_tn(const GemmArgs& g, cudaStream_t stream) {
  VCL_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  VCL_REQUIRE(g.K % 64 == 0, "gemm: K=%d must be a multiple of 64", g.K);
  VCL_REQUIRE(g.N % 32 == 0, "gemm: N=%d must be a multiple of 32", g.N);
  VCL_REQUIRE(g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldc % 8 == 0, "gemm: pitches must be x8");
  VCL_REQUIRE(((uintptr_t)g.A % 16) == 0 && ((uintptr_t)g.W % 16) == 0 &&
                  ((uintptr_t)g.C % 16) == 0, "gemm: pointers must be 16-byte aligned");
  VCL_REQUIRE(g.residual == nullptr || (g.ldr % 8 == 0 && ((uintptr_t)g.residual % 16) == 0),
              "gemm: residual must be 16-byte aligned with pitch x8");
  VCL_REQUIRE(g.bias == nullptr || ((uintptr_t)g.bias % 16) == 0, "gemm: bias alignment");
  VCL_REQUIRE(!(g.act == ACT_SWIGLU && g.residual != nullptr), "gemm: swiglu takes no residual");
  if (g.act == ACT_ROPE) {
    const RopeEpilogue& r = g.rope;
    VCL_REQUIRE(r.cos_t && r.sin_t && r.kcache && r.vcache && r.S > 0 && r.H > 0, "gemm: ACT_ROPE needs the RoPE tables and the cache");
    VCL_REQUIRE(g.N == 3 * r.H * 128 && g.M % r.S == 0 && r.start_pos + r.S <= r.s_max && g.bias == nullptr && g.residual == nullptr,
                "gemm: ACT_ROPE shape (N=%d, H=%d, M=%d, S=%d, start %d, cache %d)", g.N, r.H, g.M, r.S, r.start_pos, r.s_max);
    VCL_REQUIRE(((uintptr_t)r.kcache % 16) == 0 && ((uintptr_t)r.vcache % 16) == 0 && ((uintptr_t)r.cos_t % 16) == 0 &&
                    ((uintptr_t)r.sin_t % 16) == 0, "gemm: ACT_ROPE pointers must be 16-byte aligned");
  }
  int bn = g.block_n;
  int cl = g.cluster;
  if (bn == 0) {
    // Tile / cluster choice, from the block_n x cluster sweep on B200 (profiles/r01_gemm_sweep.txt):
    //  * many M tiles (ViT, batched prefill): 128x256 tiles, pairs of CTAs sharing each weight tile;
    //  * few M tiles (single-clip prefill, M = 448): the weight stream dominates L2->SM traffic, so
    //    all M tiles of an N tile form one cluster (multicast x4) when there are enough tiles to fill
    //    the machine twice, else narrower tiles without a cluster to get more CTAs in flight;
    //  * one M tile (decode with B > 4): the narrowest tile that still gives every SM work.
    const int sms = device_num_sms();
    const long long mt = (g.M + 127) / 128;
    int auto_cl = 1;
    if (mt >= 16 && g.N % 256 == 0) {
      // long-K tiles (ViT fc2, K = 4096): CTA pairs (cta_group::2, 256 x 256 per pair) are 2-3 % ahead of the
      // multicast pairs; with K = 1024 the pair's longer fill / drain per tile costs more than it saves
      // (profiles/r02_gemm_sweep.txt: q|k|v 154 vs 129 us, fc1 188 vs 166, out 57 vs 53, fc2 156 vs 160)
      bn = 256; auto_cl = g.K >= 4096 ? -2 : 2;
    } else if (mt >= 2) {
      const int mcl = mt >= 4 ? 4 : 2;
      if (g.N % 256 == 0 && mt * (g.N / 256) >= 2 * sms) { bn = 256; auto_cl = mcl; }
      else if (g.N % 128 == 0 && mt * (g.N / 128) >= 2 * sms) { bn = 128; auto_cl = mcl; }
      else if (g.N % 128 == 0) { bn = 128; }
      else { bn = (g.N % 64 == 0) ? 64 : 32; }
    } else {
      bn = 256;
      while (bn > 32 && (g.N % bn != 0 || mt * (g.N / bn) < sms)) bn >>= 1;
      if (g.N % bn != 0) bn = 32;
    }
    static const int forced_cl = getenv("VCL_GEMM_CLUSTER") ? atoi(getenv("VCL_GEMM_CLUSTER")) : 0;   // A/B switch
    if (cl == 0) cl = forced_cl ? forced_cl : auto_cl;
    if (forced_cl == 1 && mt >= 2 && mt < 16 && bn == 128 && g.N % 256 == 0 && mt * (g.N / 256) >= sms) bn = 256;
  }
  if (cl == 0) cl = 1;
  if (g.act == ACT_ROPE && g.block_n == 0 && bn < 128) bn = 128;      // one head (128 columns) per tile at least
  VCL_REQUIRE(g.act != ACT_ROPE || bn == 128 || bn == 256, "gemm: ACT_ROPE needs 128- or 256-wide tiles (one head = 128 columns), got %d", bn);
  // cluster = -2 (or VCL_GEMM_PAIR=1 for every launch with at least two row tiles): CTA pairs, cta_group::2
  static const bool pair_all = getenv("VCL_GEMM_PAIR") != nullptr;
  if ((cl == -2 || (pair_all && g.M > 128)) && (bn == 256 || bn == 128) && g.N % bn == 0)
    return bn == 256 ? launch_pair_act<256>(g, stream) : launch_pair_act<128>(g, stream);
  if (cl == -2) cl = 1;
  VCL_REQUIRE(cl == 1 || cl == 2 || cl == 4, "gemm: cluster must be 1, 2 or 4 (got %d)", cl);
  switch (bn) {
    case 256: return launch_act<256>(g, cl, stream);
    case 128: return launch_act<128>(g, cl, stream);
    case 64: return launch_act<64>(g, 1, stream);
    case 32: return launch_act<32>(g, 1, stream);
  }
  set_last_error("gemm: unsupported block_n %d", bn);
  return -1;
}

}  // namespace vcl
