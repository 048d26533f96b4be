// Decode-time projections for SMALL BATCHES (5 <= B <= 16; 2..4 when gemv_tc.cu does not apply):
// out[b, n] = x[b, :] . W[n, :].
//
// Still pure weight streaming (every weight byte is used once per step), but B dot products per
// weight row on the CUDA cores cost ~50 instructions per 16-byte chunk and per batch row of
// registers, which is what made the register-resident GEMV (gemv.cu) collapse for B >= 3. Here the
// dot products run on the legacy tensor path (mma.sync m16n8k16, bf16 -> fp32) straight out of
// REGISTERS -- no shared-memory staging and no layout shuffle:
//
//   * a warp owns 16 weight rows x a K slice; lane (g = lane/4, q = lane%4) loads, per 32-wide K
//     block, one 16-byte chunk of row g and one of row g+8 (k = 8q..8q+7) with non-allocating loads;
//   * a dot product does not care about the order of k, so those 8 consecutive elements are fed to
//     two MMAs under a fixed permutation (elements 0-3 -> MMA 1, 4-7 -> MMA 2) and the activation
//     fragment is loaded with the SAME 16-byte pattern (x[n = g][8q..8q+7], L1-resident), which makes
//     every load a full 128-bit access and every MMA operand a plain register;
//   * ~5 instructions per 32 B of weights per lane instead of ~100; 8 independent 128-bit loads per
//     lane in flight; the 8 warps of a CTA split K and meet in shared memory; fused epilogues as in
//     gemv.cu (residual, SwiGLU over interleaved rows, RoPE + KV append, logits).
//
// The input vector must already be normalised (launch_rmsnorm): with B up to 16 rows the norm is a
// separate 10 KB-sized kernel rather than a per-CTA prologue.
#include "common.cuh"
#include "kernels.h"

namespace vcl {

namespace {

enum { MODE_RES = 0, MODE_SWIGLU = 1, MODE_QKV = 2, MODE_LOGITS = 3 };
constexpr int GM_THREADS = 256;
constexpr int GM_WARPS = 8;

struct GmParams {
  const bf16* x; long long ldx;     // [B, K] (already normalised)
  const bf16* W;                    // [N, K]
  int B, N, K;
  int rows_per_cta;                 // multiple of 16
  bf16* out; long long ldo;
  const bf16* res; long long ldr;
  bf16* q_out; long long ldq;
  bf16* kcache; bf16* vcache;
  const bf16* cos_t; const bf16* sin_t;
  int H, s_max, pos;
  const int* pos_dev;
  float* logits; long long ldl;
};

template <int MODE>
__device__ __forceinline__ long long gm_map_row(int v) {
  if (MODE == MODE_QKV) {
    const int within = v & 127;
    return (long long)(v >> 7) * 128 + (within >> 1) + ((within & 1) << 6);
  }
  return v;
}

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// NBLK = number of 8-row batch blocks (1: B <= 8, 2: B <= 16)
template <int NBLK, int MODE>
__global__ void __launch_bounds__(GM_THREADS, 2) gemv_mma_kernel(const GmParams p) {
  extern __shared__ __align__(16) float part[];   // [GM_WARPS][rows_per_cta][NBLK*8]
  constexpr int NC = NBLK * 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, q = lane & 3;
  const int K = p.K, R = p.rows_per_cta;
  const int row_begin = blockIdx.x * R;
  const int n_rows = max(0, min(p.N, row_begin + R) - row_begin);
  const int n_groups = (n_rows + 15) / 16;
  // this warp's K slice, in 32-element blocks
  const int kb_total = K / 32;
  const int kb_per = (kb_total + GM_WARPS - 1) / GM_WARPS;
  const int kb0 = warp * kb_per, kb1 = min(kb_total, kb0 + kb_per);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const bf16* xr0 = p.x + (long long)g * p.ldx + q * 8;           // batch row g
  const bf16* xr1 = p.x + (long long)(g + 8) * p.ldx + q * 8;     // batch row g + 8 (NBLK == 2)
  const bool x0_ok = g < p.B, x1_ok = (g + 8) < p.B;

  for (int grp = 0; grp < n_groups; ++grp) {
    const int ra = grp * 16 + g, rb = ra + 8;                      // rows inside the CTA block
    const bool ra_ok = ra < n_rows, rb_ok = rb < n_rows;
    const bf16* wa = p.W + gm_map_row<MODE>(row_begin + (ra_ok ? ra : 0)) * K + q * 8;
    const bf16* wb = p.W + gm_map_row<MODE>(row_begin + (rb_ok ? rb : 0)) * K + q * 8;
    float c[NBLK][4];
#pragma unroll
    for (int nb = 0; nb < NBLK; ++nb) c[nb][0] = c[nb][1] = c[nb][2] = c[nb][3] = 0.f;
    constexpr int U = 4;                                           // K blocks in flight per lane
    for (int kb = kb0; kb < kb1; kb += U) {
      // weights (HBM) and activation fragments (L1/L2) of U K-blocks are all issued before any MMA
      uint4 w0[U], w1[U], xa[U], xb[NBLK == 2 ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = (kb + u) * 32;
        const bool ok = kb + u < kb1;
        w0[u] = (ok && ra_ok) ? ld_nc_v4(wa + k) : make_uint4(0, 0, 0, 0);
        w1[u] = (ok && rb_ok) ? ld_nc_v4(wb + k) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int k = (kb + u) * 32;
        const bool ok = kb + u < kb1;
        xa[u] = (ok && x0_ok) ? __ldg(reinterpret_cast<const uint4*>(xr0 + k)) : make_uint4(0, 0, 0, 0);
        if (NBLK == 2)
          xb[NBLK == 2 ? u : 0] = (ok && x1_ok) ? __ldg(reinterpret_cast<const uint4*>(xr1 + k)) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        mma16816(c[0], w0[u].x, w1[u].x, w0[u].y, w1[u].y, xa[u].x, xa[u].y);
        mma16816(c[0], w0[u].z, w1[u].z, w0[u].w, w1[u].w, xa[u].z, xa[u].w);
        if (NBLK == 2) {
          const uint4 xq = xb[NBLK == 2 ? u : 0];
          mma16816(c[NBLK - 1], w0[u].x, w1[u].x, w0[u].y, w1[u].y, xq.x, xq.y);
          mma16816(c[NBLK - 1], w0[u].z, w1[u].z, w0[u].w, w1[u].w, xq.z, xq.w);
        }
      }
    }
    // c[nb]: (row g, n = 8nb + 2q, +1), (row g+8, same n)
#pragma unroll
    for (int nb = 0; nb < NBLK; ++nb) {
      float* pa = part + ((size_t)warp * R + grp * 16 + g) * NC + nb * 8 + 2 * q;
      float* pb = pa + 8 * NC;
      *reinterpret_cast<float2*>(pa) = make_float2(c[nb][0], c[nb][1]);
      *reinterpret_cast<float2*>(pb) = make_float2(c[nb][2], c[nb][3]);
    }
  }
  __syncthreads();

  // ---------------- epilogue ----------------
  constexpr bool PAIRS = (MODE == MODE_SWIGLU || MODE == MODE_QKV);
  const int units = PAIRS ? n_rows / 2 : n_rows;
  for (int it = tid; it < units * p.B; it += GM_THREADS) {
    const int b = it % p.B, u = it / p.B;
    const int r0 = PAIRS ? 2 * u : u;
    float v0 = 0.f, v1 = 0.f;
#pragma unroll
    for (int w = 0; w < GM_WARPS; ++w) {
      v0 += part[((size_t)w * R + r0) * NC + b];
      if (PAIRS) v1 += part[((size_t)w * R + r0 + 1) * NC + b];
    }
    const int vrow = row_begin + r0;
    if (MODE == MODE_RES) {
      float y = bf16r(v0);
      if (p.res != nullptr) y += __bfloat162float(p.res[(long long)b * p.ldr + vrow]);
      p.out[(long long)b * p.ldo + vrow] = __float2bfloat16_rn(y);
    } else if (MODE == MODE_LOGITS) {
      p.logits[(long long)b * p.ldl + vrow] = bf16r(v0);
    } else if (MODE == MODE_SWIGLU) {
      const float gt = bf16r(v0);
      const float sg = bf16r(__fdividef(gt, 1.0f + __expf(-gt)));
      p.out[(long long)b * p.ldo + (vrow >> 1)] = __float2bfloat16_rn(sg * bf16r(v1));
    } else {
      const int hr = vrow >> 7;
      const int which = hr / p.H, head = hr - which * p.H;
      const int d = (vrow & 127) >> 1;
      const float lo = bf16r(v0), hi = bf16r(v1);
      const int pos = p.pos + (p.pos_dev != nullptr ? __ldg(p.pos_dev) : 0);
      const long long coff = (((long long)b * p.H + head) * p.s_max + pos) * 128;
      if (which == 2) {
        p.vcache[coff + d] = __float2bfloat16_rn(lo);
        p.vcache[coff + d + 64] = __float2bfloat16_rn(hi);
      } else {
        const float cs = __bfloat162float(p.cos_t[(long long)pos * 64 + d]);
        const float sn = __bfloat162float(p.sin_t[(long long)pos * 64 + d]);
        const float olo = bf16r(lo * cs) + bf16r(-hi * sn);
        const float ohi = bf16r(hi * cs) + bf16r(lo * sn);
        if (which == 0) {
          p.q_out[(long long)b * p.ldq + head * 128 + d] = __float2bfloat16_rn(olo);
          p.q_out[(long long)b * p.ldq + head * 128 + d + 64] = __float2bfloat16_rn(ohi);
        } else {
          p.kcache[coff + d] = __float2bfloat16_rn(olo);
          p.kcache[coff + d + 64] = __float2bfloat16_rn(ohi);
        }
      }
    }
  }
}

template <int NBLK, int MODE>
int gm_launch(GmParams p, cudaStream_t stream) {
  const int slots = 2 * device_num_sms();
  int R = ((p.N + slots - 1) / slots + 15) / 16 * 16;
  const int grid = (p.N + R - 1) / R;
  p.rows_per_cta = R;
  const size_t smem = (size_t)GM_WARPS * R * NBLK * 8 * sizeof(float);
  VCL_REQUIRE(smem <= 100 * 1024, "gemv_mma: %d rows per CTA need %zu bytes of shared memory", R, smem);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(GM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  VCL_CUDA_OK(cudaLaunchKernelEx(&cfg, gemv_mma_kernel<NBLK, MODE>, p));
  count_launches(1);
  return 0;
}

template <int MODE>
int gm_dispatch(const GmParams& p, cudaStream_t stream) {
  VCL_REQUIRE(p.B >= 1 && p.B <= 16, "gemv_mma: batch %d outside 1..16", p.B);
  VCL_REQUIRE(p.K % 32 == 0 && p.ldx % 8 == 0, "gemv_mma: K must be a multiple of 32");
  VCL_REQUIRE(((uintptr_t)p.x % 16) == 0 && ((uintptr_t)p.W % 16) == 0, "gemv_mma: 16-byte alignment required");
  return p.B <= 8 ? gm_launch<1, MODE>(p, stream) : gm_launch<2, MODE>(p, stream);
}

template <int MODE>
int gm_init() {
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_mma_kernel<1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_mma_kernel<2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  return 0;
}

GmParams gm_base(const GemvArgs& g) {
  GmParams p = {};
  p.x = g.x; p.ldx = g.ldx; p.W = g.W; p.B = g.B; p.N = g.N; p.K = g.K;
  return p;
}

}  // namespace

int init_gemv_mma_kernels() {
  if (gm_init<MODE_RES>() || gm_init<MODE_SWIGLU>() || gm_init<MODE_QKV>() || gm_init<MODE_LOGITS>()) return -2;
  return 0;
}

int launch_gemv_mma_residual(const GemvArgs& g, bf16* out, long long ldo, const bf16* res, long long ldr,
                             cudaStream_t stream) {
  VCL_REQUIRE(g.norm_w == nullptr, "gemv_mma: the input must be normalised beforehand");
  GmParams p = gm_base(g);
  p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr;
  return gm_dispatch<MODE_RES>(p, stream);
}

int launch_gemv_mma_swiglu(const GemvArgs& g, bf16* out, long long ldo, cudaStream_t stream) {
  VCL_REQUIRE(g.norm_w == nullptr && g.N % 2 == 0, "gemv_mma swiglu: bad arguments");
  GmParams p = gm_base(g);
  p.out = out; p.ldo = ldo;
  return gm_dispatch<MODE_SWIGLU>(p, stream);
}

int launch_gemv_mma_qkv_rope(const GemvArgs& g, bf16* q_out, long long ldq, bf16* kcache, bf16* vcache,
                             const bf16* cos_t, const bf16* sin_t, int H, int head_dim, int s_max, int pos,
                             cudaStream_t stream, const int* pos_dev) {
  VCL_REQUIRE(g.norm_w == nullptr && head_dim == 128 && g.N == 3 * H * 128, "gemv_mma qkv: bad arguments");
  VCL_REQUIRE(pos >= 0 && pos < s_max, "gemv_mma qkv: position %d outside the cache (%d)", pos, s_max);
  GmParams p = gm_base(g);
  p.q_out = q_out; p.ldq = ldq; p.kcache = kcache; p.vcache = vcache;
  p.cos_t = cos_t; p.sin_t = sin_t; p.H = H; p.s_max = s_max; p.pos = pos; p.pos_dev = pos_dev;
  return gm_dispatch<MODE_QKV>(p, stream);
}

int launch_gemv_mma_logits(const GemvArgs& g, float* logits, long long ldl, cudaStream_t stream) {
  VCL_REQUIRE(g.norm_w == nullptr, "gemv_mma: the input must be normalised beforehand");
  GmParams p = gm_base(g);
  p.logits = logits; p.ldl = ldl;
  return gm_dispatch<MODE_LOGITS>(p, stream);
}

}  // namespace vcl
