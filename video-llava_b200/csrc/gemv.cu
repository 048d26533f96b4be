// Decode-time projections: out[b, n] = x[b, :] . W[n, :] for B <= 4 new tokens.
//
// With one token per clip every weight byte is used once per step (13.2 GB per step for the 7B
// model), so these kernels are pure HBM streaming: each warp owns whole
// weight rows, reads them with 128-bit non-allocating loads (U of them in flight per lane) and
// keeps the activations in shared memory as fp32. Tensor cores are deliberately not used here
// (M = B <= 4 would waste >96 % of an MMA tile and the bound is HBM either way); B > 4 goes
// through the tcgen05 GEMM with a narrow N tile instead.
//
// Fusions (each removes a launch and an HBM round trip from the 32-layer decode step):
//   prologue  LlamaRMSNorm of x (transformers/models/llama/modeling_llama.py:53-67)
//   epilogue  RES     bf16(bf16(acc) + residual)               (modeling_llama.py:325,331)
//             SWIGLU  silu(gate) * up on interleaved rows      (modeling_llama.py:182-184)
//             QKV     RoPE on q,k + KV-cache append            (modeling_llama.py:124-168,262-270)
//             LOGITS  bf16-rounded logits kept as fp32 for the arg-max
#include "common.cuh"
#include "kernels.h"

namespace vcl {

namespace {

enum { MODE_RES = 0, MODE_SWIGLU = 1, MODE_QKV = 2, MODE_LOGITS = 3 };
constexpr int GEMV_THREADS = 512;

struct GemvParams {
  const bf16* x; long long ldx;
  const bf16* W;
  int N, K;
  int n_tasks;
  const bf16* norm_w; float eps;
  // RES / SWIGLU
  bf16* out; long long ldo;
  const bf16* res; long long ldr;
  // QKV
  bf16* q_out; long long ldq;
  bf16* kcache; bf16* vcache;
  const bf16* cos_t; const bf16* sin_t;
  int H, s_max, pos;
  // LOGITS
  float* logits; long long ldl;
};

__device__ __forceinline__ float silu_bf16(float g) {
  g = bf16r(g);
  return bf16r(g / (1.0f + __expf(-g)));
}

template <int NB, int MODE>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_kernel(const GemvParams p) {
  constexpr int R = (MODE == MODE_SWIGLU || MODE == MODE_QKV) ? 2 : 1;
  constexpr int U = (R == 2) ? 4 : 8;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float red[GEMV_THREADS / 32];
  const int K = p.K;
  const int nch = K >> 3;
  // per batch row: plane 0 holds elements 0..3 of every 8-chunk, plane 1 elements 4..7
  float4* xs = reinterpret_cast<float4*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---------------- prologue: stage (optionally RMS-normalised) activations ----------------
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    const bf16* xr = p.x + (long long)b * p.ldx;
    float rstd = 1.f;
    if (p.norm_w != nullptr) {
      float ss = 0.f;
      for (int c = tid; c < nch; c += GEMV_THREADS) {
        const uint4 u = *reinterpret_cast<const uint4*>(xr + c * 8);
        const float f0 = bf16lo(u.x), f1 = bf16hi(u.x), f2 = bf16lo(u.y), f3 = bf16hi(u.y);
        const float f4 = bf16lo(u.z), f5 = bf16hi(u.z), f6 = bf16lo(u.w), f7 = bf16hi(u.w);
        ss += f0 * f0 + f1 * f1 + f2 * f2 + f3 * f3 + f4 * f4 + f5 * f5 + f6 * f6 + f7 * f7;
      }
      ss = warp_sum(ss);
      __syncthreads();
      if (lane == 0) red[warp] = ss;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < GEMV_THREADS / 32; ++w) tot += red[w];
      rstd = rsqrtf(tot / (float)K + p.eps);
    }
    for (int c = tid; c < nch; c += GEMV_THREADS) {
      const uint4 u = *reinterpret_cast<const uint4*>(xr + c * 8);
      float f[8] = {bf16lo(u.x), bf16hi(u.x), bf16lo(u.y), bf16hi(u.y),
                    bf16lo(u.z), bf16hi(u.z), bf16lo(u.w), bf16hi(u.w)};
      if (p.norm_w != nullptr) {
        const uint4 wu = *reinterpret_cast<const uint4*>(p.norm_w + c * 8);
        const float g[8] = {bf16lo(wu.x), bf16hi(wu.x), bf16lo(wu.y), bf16hi(wu.y),
                            bf16lo(wu.z), bf16hi(wu.z), bf16lo(wu.w), bf16hi(wu.w)};
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = bf16r(g[j] * bf16r(f[j] * rstd));
      }
      xs[(b * 2 + 0) * nch + c] = make_float4(f[0], f[1], f[2], f[3]);
      xs[(b * 2 + 1) * nch + c] = make_float4(f[4], f[5], f[6], f[7]);
    }
  }
  __syncthreads();

  // ---------------- main loop: one task (R weight rows) per warp iteration ----------------
  const int gw = blockIdx.x * (GEMV_THREADS / 32) + warp;
  const int tw = gridDim.x * (GEMV_THREADS / 32);
  const int HD2 = 64;
#pragma unroll 1
  for (int task = gw; task < p.n_tasks; task += tw) {
    long long row0, row1 = 0;
    int which = 0, head = 0, d = 0;
    if (MODE == MODE_SWIGLU) {
      row0 = 2LL * task; row1 = row0 + 1;
    } else if (MODE == MODE_QKV) {
      const int per = p.H * HD2;
      which = task / per;
      const int rem = task - which * per;
      head = rem / HD2; d = rem - head * HD2;
      row0 = (long long)which * p.H * 128 + head * 128 + d;
      row1 = row0 + 64;
    } else {
      row0 = task;
    }
    const bf16* w0 = p.W + row0 * K;
    const bf16* w1 = p.W + row1 * K;
    float acc[R][NB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;

    for (int c0 = lane; c0 < nch; c0 += 32 * U) {
      uint4 wv[R][U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 32;
        if (c < nch) {
          wv[0][u] = ld_nc_v4(w0 + c * 8);
          if (R == 2) wv[R - 1][u] = ld_nc_v4(w1 + c * 8);
        } else {
          wv[0][u] = make_uint4(0, 0, 0, 0);
          if (R == 2) wv[R - 1][u] = make_uint4(0, 0, 0, 0);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int c = c0 + u * 32;
        if (c < nch) {
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            const float4 xa = xs[(b * 2 + 0) * nch + c];
            const float4 xb = xs[(b * 2 + 1) * nch + c];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const uint4 w = wv[r][u];
              float s = acc[r][b];
              s = fmaf(bf16lo(w.x), xa.x, s); s = fmaf(bf16hi(w.x), xa.y, s);
              s = fmaf(bf16lo(w.y), xa.z, s); s = fmaf(bf16hi(w.y), xa.w, s);
              s = fmaf(bf16lo(w.z), xb.x, s); s = fmaf(bf16hi(w.z), xb.y, s);
              s = fmaf(bf16lo(w.w), xb.z, s); s = fmaf(bf16hi(w.w), xb.w, s);
              acc[r][b] = s;
            }
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);

    if (lane == 0) {
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        if (MODE == MODE_RES) {
          float v = bf16r(acc[0][b]);
          if (p.res != nullptr) v += __bfloat162float(p.res[(long long)b * p.ldr + row0]);
          p.out[(long long)b * p.ldo + row0] = __float2bfloat16_rn(v);
        } else if (MODE == MODE_SWIGLU) {
          const float y = silu_bf16(acc[0][b]) * bf16r(acc[R - 1][b]);
          p.out[(long long)b * p.ldo + task] = __float2bfloat16_rn(y);
        } else if (MODE == MODE_LOGITS) {
          p.logits[(long long)b * p.ldl + row0] = bf16r(acc[0][b]);
        } else {  // MODE_QKV
          const float lo = bf16r(acc[0][b]), hi = bf16r(acc[R - 1][b]);
          const long long coff = (((long long)b * p.H + head) * p.s_max + p.pos) * 128;
          if (which == 2) {
            p.vcache[coff + d] = __float2bfloat16_rn(lo);
            p.vcache[coff + d + 64] = __float2bfloat16_rn(hi);
          } else {
            const float c = __bfloat162float(p.cos_t[(long long)p.pos * 64 + d]);
            const float s = __bfloat162float(p.sin_t[(long long)p.pos * 64 + d]);
            const float olo = bf16r(lo * c) + bf16r(-hi * s);
            const float ohi = bf16r(hi * c) + bf16r(lo * s);
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
  }
}

template <int NB, int MODE>
int launch_nb(const GemvParams& p, cudaStream_t stream) {
  const size_t smem = (size_t)NB * p.K * sizeof(float);
  auto kern = gemv_kernel<NB, MODE>;
  const int wpc = GEMV_THREADS / 32;
  int grid = device_num_sms();
  const int need = (p.n_tasks + wpc - 1) / wpc;
  if (grid > need) grid = need;
  kern<<<grid, GEMV_THREADS, smem, stream>>>(p);
  VCL_CUDA_OK(cudaGetLastError());
  count_launches(1);
  return 0;
}

template <int MODE>
int launch_mode(int B, const GemvParams& p, cudaStream_t stream) {
  VCL_REQUIRE(B >= 1 && B <= 4, "gemv: batch %d outside 1..4 (larger batches use the tcgen05 GEMM)", B);
  VCL_REQUIRE(p.K % 8 == 0 && p.ldx % 8 == 0, "gemv: K and pitch must be multiples of 8");
  VCL_REQUIRE((size_t)B * p.K * 4 <= 226 * 1024, "gemv: B*K*4 = %zu exceeds shared memory",
              (size_t)B * p.K * 4);
  switch (B) {
    case 1: return launch_nb<1, MODE>(p, stream);
    case 2: return launch_nb<2, MODE>(p, stream);
    case 3: return launch_nb<3, MODE>(p, stream);
    default: return launch_nb<4, MODE>(p, stream);
  }
}

template <int MODE>
int init_mode() {
  const int cap = 227 * 1024 - 1024;  // static smem (reduction scratch) counts against the limit
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_kernel<1, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_kernel<2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_kernel<3, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  VCL_CUDA_OK(cudaFuncSetAttribute(gemv_kernel<4, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  return 0;
}

GemvParams base_params(const GemvArgs& g) {
  GemvParams p = {};
  p.x = g.x; p.ldx = g.ldx; p.W = g.W; p.N = g.N; p.K = g.K;
  p.norm_w = g.norm_w; p.eps = g.eps;
  return p;
}

}  // namespace

int init_gemv_kernels() {
  if (init_mode<MODE_RES>() || init_mode<MODE_SWIGLU>() || init_mode<MODE_QKV>() ||
      init_mode<MODE_LOGITS>()) return -2;
  return 0;
}

int launch_gemv_residual(const GemvArgs& g, bf16* out, long long ldo, const bf16* res,
                         long long ldr, cudaStream_t stream) {
  GemvParams p = base_params(g);
  p.n_tasks = g.N; p.out = out; p.ldo = ldo; p.res = res; p.ldr = ldr;
  return launch_mode<MODE_RES>(g.B, p, stream);
}

int launch_gemv_swiglu(const GemvArgs& g, bf16* out, long long ldo, cudaStream_t stream) {
  VCL_REQUIRE(g.N % 2 == 0, "gemv swiglu: N must be even (interleaved gate/up rows)");
  GemvParams p = base_params(g);
  p.n_tasks = g.N / 2; p.out = out; p.ldo = ldo;
  return launch_mode<MODE_SWIGLU>(g.B, p, stream);
}

int launch_gemv_qkv_rope(const GemvArgs& g, bf16* q_out, long long ldq, bf16* kcache, bf16* vcache,
                         const bf16* cos_t, const bf16* sin_t, int H, int head_dim, int s_max,
                         int pos, cudaStream_t stream) {
  VCL_REQUIRE(head_dim == 128, "gemv qkv: head_dim must be 128");
  VCL_REQUIRE(g.N == 3 * H * 128, "gemv qkv: N=%d != 3*H*128", g.N);
  VCL_REQUIRE(pos >= 0 && pos < s_max, "gemv qkv: position %d outside the cache (%d)", pos, s_max);
  GemvParams p = base_params(g);
  p.n_tasks = 3 * H * 64; p.q_out = q_out; p.ldq = ldq; p.kcache = kcache; p.vcache = vcache;
  p.cos_t = cos_t; p.sin_t = sin_t; p.H = H; p.s_max = s_max; p.pos = pos;
  return launch_mode<MODE_QKV>(g.B, p, stream);
}

int launch_gemv_logits(const GemvArgs& g, float* logits, long long ldl, cudaStream_t stream) {
  GemvParams p = base_params(g);
  p.n_tasks = g.N; p.logits = logits; p.ldl = ldl;
  return launch_mode<MODE_LOGITS>(g.B, p, stream);
}

}  // namespace vcl
