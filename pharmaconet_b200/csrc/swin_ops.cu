// swin_ops.cu - the non-GEMM part of a 3-D Swin-V2 block in two kernels (the GEMMs stay plain library GEMMs).
//
//  pmnet_window_attention  cyclic shift + window partition + cosine attention + continuous position bias + shift
//                          mask + softmax + P.V + window reverse + un-shift, one read of qkv and one write of the
//                          result per token (src/pmnet/network/backbones/swinv2.py:114-158, 272-298,
//                          swin.py:46-96). Reference quirk kept: the shift rolls only the first two spatial axes
//                          (swinv2.py:277,296).
//  pmnet_ln_residual       x = shortcut + LayerNorm(h) (res-post-norm, swinv2.py:300-303), one warp per token.
//
// Both are memory-bound: the reference's eager sequence makes ~25 passes over the token tensor per block, these
// make 4 (qkv read, attention write, two LayerNorm passes). fp32 arithmetic; storage fp32 or bf16.

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pmnet_b200.h"

extern void pmnet_set_error(const char* msg);

namespace {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return __ldg(p); }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
// 32 contiguous values of a token-head row, as 128-bit accesses
__device__ __forceinline__ void load_row32(const float* p, float (&v)[32]) {
  const float4* q = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(q + i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load_row32(const __nv_bfloat16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 t = __ldg(q + i);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __bfloat1622float2(h[j]);
      v[8 * i + 2 * j] = f.x;
      v[8 * i + 2 * j + 1] = f.y;
    }
  }
}
__device__ __forceinline__ void store_row32(float* p, const float (&v)[32]) {
  float4* q = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
// two-term bf16 split of 32 values: hi = bf16(v), lo = bf16(v - hi) (the operand pair of a split-precision GEMM)
__device__ __forceinline__ void store_row32_split(__nv_bfloat16* hi, __nv_bfloat16* lo, const float (&v)[32]) {
  uint32_t h[16], l[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&h2);
    const float2 f = __bfloat1622float2(h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - f.x, v[2 * i + 1] - f.y);
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  uint4* qh = reinterpret_cast<uint4*>(hi);
  uint4* ql = reinterpret_cast<uint4*>(lo);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    qh[i] = make_uint4(h[4 * i], h[4 * i + 1], h[4 * i + 2], h[4 * i + 3]);
    ql[i] = make_uint4(l[4 * i], l[4 * i + 1], l[4 * i + 2], l[4 * i + 3]);
  }
}

constexpr int kHeadDim = 32;
constexpr int kTokens = 64;  // 4^3 window

// grid (windows, heads), block 64: thread t = token t of the window.
template <typename T>
__global__ void __launch_bounds__(kTokens) window_attention_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                   const float* __restrict__ scale,
                                                                   const float* __restrict__ rel_bias,
                                                                   const float* __restrict__ mask, int res, int shift,
                                                                   int heads, __nv_bfloat16* __restrict__ out_hi,
                                                                   __nv_bfloat16* __restrict__ out_lo) {
  __shared__ __align__(16) float ks[kTokens][kHeadDim];  // every thread reads the same row: broadcast, no conflicts
  __shared__ __align__(16) float vs[kTokens][kHeadDim];
  const int t = threadIdx.x, h = blockIdx.y;
  const int nw1 = res / 4, nw = nw1 * nw1 * nw1;
  const int win = blockIdx.x % nw, b = blockIdx.x / nw;
  const int wd = win / (nw1 * nw1), wh = (win / nw1) % nw1, ww = win % nw1;
  // position inside the shifted volume -> source token (roll by -shift on D and H only)
  int d = wd * 4 + (t >> 4), hh = wh * 4 + ((t >> 2) & 3);
  const int w = ww * 4 + (t & 3);
  d = (d + shift) % res;
  hh = (hh + shift) % res;
  const size_t tok = (((size_t)b * res + d) * res + hh) * res + w;
  const int C = heads * kHeadDim;
  const T* src = qkv + tok * 3 * C + h * kHeadDim;
  float q[kHeadDim], kk[kHeadDim], vv[kHeadDim];
  load_row32(src, q);
  load_row32(src + C, kk);
  load_row32(src + 2 * C, vv);
  float qn = 0.f, kn = 0.f;
#pragma unroll
  for (int i = 0; i < kHeadDim; ++i) {
    qn = fmaf(q[i], q[i], qn);
    kn = fmaf(kk[i], kk[i], kn);
    vs[t][i] = vv[i];
  }
  // F.normalize(dim=-1, eps=1e-12) on q and k, then the per-head logit scale (folded into q)
  const float qi = scale[h] / fmaxf(sqrtf(qn), 1e-12f), ki = 1.0f / fmaxf(sqrtf(kn), 1e-12f);
#pragma unroll
  for (int i = 0; i < kHeadDim; ++i) {
    q[i] *= qi;
    ks[t][i] = kk[i] * ki;
  }
  __syncthreads();
  float s[kTokens];
  const float* bias = rel_bias + ((size_t)h * kTokens + t) * kTokens;
  const float* mk = mask ? mask + ((size_t)win * kTokens + t) * kTokens : nullptr;
  float mx = -3.0e38f;
#pragma unroll
  for (int j = 0; j < kTokens; ++j) {
    // four independent partial sums: the 32-long dot product is otherwise one dependent FMA chain
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const float4* kr = reinterpret_cast<const float4*>(ks[j]);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      const float4 k4 = kr[i];
      a0 = fmaf(q[4 * i + 0], k4.x, a0);
      a1 = fmaf(q[4 * i + 1], k4.y, a1);
      a2 = fmaf(q[4 * i + 2], k4.z, a2);
      a3 = fmaf(q[4 * i + 3], k4.w, a3);
    }
    float a = (a0 + a1) + (a2 + a3);
    a += __ldg(bias + j);
    if (mk) a += __ldg(mk + j);
    s[j] = a;
    mx = fmaxf(mx, a);
  }
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < kTokens; ++j) {
    s[j] = __expf(s[j] - mx);
    sum += s[j];
  }
  const float inv = 1.0f / sum;
  float o[kHeadDim];
#pragma unroll
  for (int i = 0; i < kHeadDim; ++i) o[i] = 0.f;
#pragma unroll
  for (int j = 0; j < kTokens; ++j) {
    const float p = s[j] * inv;
    const float4* vr = reinterpret_cast<const float4*>(vs[j]);
#pragma unroll
    for (int i = 0; i < kHeadDim / 4; ++i) {
      const float4 v4 = vr[i];
      o[4 * i + 0] = fmaf(p, v4.x, o[4 * i + 0]);
      o[4 * i + 1] = fmaf(p, v4.y, o[4 * i + 1]);
      o[4 * i + 2] = fmaf(p, v4.z, o[4 * i + 2]);
      o[4 * i + 3] = fmaf(p, v4.w, o[4 * i + 3]);
    }
  }
  if (out) store_row32(out + tok * C + h * kHeadDim, o);
  if (out_hi) store_row32_split(out_hi + tok * C + h * kHeadDim, out_lo + tok * C + h * kHeadDim, o);
}


// ------------------------------------------------------------------ tensor-core version for bf16 storage
// One warp per (window, head): S = Q K^T (64 x 64 x 32) and O = P V (64 x 32 x 64) with mma.sync m16n8k16 (bf16 in,
// fp32 accumulate), 16 query rows at a time so that S stays in 32 registers; P is re-used from the accumulator
// registers as the A operand of the second product. (tcgen05 needs M >= 64 tiles resident in TMEM and a
// TMEM round trip for the softmax; for 64-token windows that is slower than staying in registers.)
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

constexpr int kQStride = 20;  // words per Q/K row (16 data + 4 pad: conflict-free fragment reads)
constexpr int kVStride = 36;  // words per V^T row (32 data + 4 pad)

__global__ void __launch_bounds__(32) window_attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                  __nv_bfloat16* __restrict__ out,
                                                                  const float* __restrict__ scale,
                                                                  const float* __restrict__ rel_bias,
                                                                  const float* __restrict__ mask, int res, int shift,
                                                                  int heads) {
  __shared__ uint32_t Qs[kTokens * kQStride];
  __shared__ uint32_t Ks[kTokens * kQStride];
  __shared__ uint32_t Vt[kHeadDim * kVStride];
  __shared__ unsigned long long tok_of[kTokens];
  const int lane = threadIdx.x, h = blockIdx.y;
  const int nw1 = res / 4, nw = nw1 * nw1 * nw1;
  const int win = blockIdx.x % nw, b = blockIdx.x / nw;
  const int wd = win / (nw1 * nw1), wh = (win / nw1) % nw1, ww = win % nw1;
  const int C = heads * kHeadDim;
  __nv_bfloat16* vt16 = reinterpret_cast<__nv_bfloat16*>(Vt);
  // ---- load, normalise q and k (fp32), stage as bf16; lane handles tokens lane and lane + 32
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int t = lane + 32 * half;
    int d = wd * 4 + (t >> 4), hh = wh * 4 + ((t >> 2) & 3);
    const int w = ww * 4 + (t & 3);
    d = (d + shift) % res;
    hh = (hh + shift) % res;
    const size_t tok = (((size_t)b * res + d) * res + hh) * res + w;
    tok_of[t] = tok;
    const __nv_bfloat16* src = qkv + tok * 3 * C + h * kHeadDim;
    float q[kHeadDim], kk[kHeadDim], vv[kHeadDim];
    load_row32(src, q);
    load_row32(src + C, kk);
    load_row32(src + 2 * C, vv);
    float qn = 0.f, kn = 0.f;
#pragma unroll
    for (int i = 0; i < kHeadDim; ++i) {
      qn = fmaf(q[i], q[i], qn);
      kn = fmaf(kk[i], kk[i], kn);
    }
    const float qi = 1.0f / fmaxf(sqrtf(qn), 1e-12f), ki = 1.0f / fmaxf(sqrtf(kn), 1e-12f);
#pragma unroll
    for (int i = 0; i < kHeadDim / 2; ++i) {
      Qs[t * kQStride + i] = pack_bf16(q[2 * i] * qi, q[2 * i + 1] * qi);
      Ks[t * kQStride + i] = pack_bf16(kk[2 * i] * ki, kk[2 * i + 1] * ki);
    }
#pragma unroll
    for (int i = 0; i < kHeadDim; ++i) vt16[i * (2 * kVStride) + t] = __float2bfloat16_rn(vv[i]);
  }
  __syncwarp();
  const int g = lane >> 2, tq = lane & 3;
  const float sc = scale[h];
#pragma unroll 1
  for (int mt = 0; mt < 4; ++mt) {
    const int r0 = 16 * mt + g, r1 = r0 + 8;
    float S[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) S[nt][0] = S[nt][1] = S[nt][2] = S[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const uint32_t a0 = Qs[r0 * kQStride + ks * 8 + tq], a1 = Qs[r1 * kQStride + ks * 8 + tq];
      const uint32_t a2 = Qs[r0 * kQStride + ks * 8 + 4 + tq], a3 = Qs[r1 * kQStride + ks * 8 + 4 + tq];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b0 = Ks[(8 * nt + g) * kQStride + ks * 8 + tq], b1 = Ks[(8 * nt + g) * kQStride + ks * 8 + 4 + tq];
        mma16816(S[nt], a0, a1, a2, a3, b0, b1);
      }
    }
    // logits = scale * cos + bias (+ mask); row maxima
    const float* bias0 = rel_bias + ((size_t)h * kTokens + r0) * kTokens + 2 * tq;
    const float* bias1 = rel_bias + ((size_t)h * kTokens + r1) * kTokens + 2 * tq;
    const float* m0 = mask ? mask + ((size_t)win * kTokens + r0) * kTokens + 2 * tq : nullptr;
    const float* m1 = mask ? mask + ((size_t)win * kTokens + r1) * kTokens + 2 * tq : nullptr;
    float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 bz0 = __ldg(reinterpret_cast<const float2*>(bias0 + 8 * nt));
      const float2 bz1 = __ldg(reinterpret_cast<const float2*>(bias1 + 8 * nt));
      S[nt][0] = fmaf(S[nt][0], sc, bz0.x);
      S[nt][1] = fmaf(S[nt][1], sc, bz0.y);
      S[nt][2] = fmaf(S[nt][2], sc, bz1.x);
      S[nt][3] = fmaf(S[nt][3], sc, bz1.y);
      if (mask) {
        const float2 k0 = __ldg(reinterpret_cast<const float2*>(m0 + 8 * nt));
        const float2 k1 = __ldg(reinterpret_cast<const float2*>(m1 + 8 * nt));
        S[nt][0] += k0.x; S[nt][1] += k0.y; S[nt][2] += k1.x; S[nt][3] += k1.y;
      }
      mx0 = fmaxf(mx0, fmaxf(S[nt][0], S[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(S[nt][2], S[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      S[nt][0] = __expf(S[nt][0] - mx0); S[nt][1] = __expf(S[nt][1] - mx0);
      S[nt][2] = __expf(S[nt][2] - mx1); S[nt][3] = __expf(S[nt][3] - mx1);
      sum0 += S[nt][0] + S[nt][1];
      sum1 += S[nt][2] + S[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // O = P V with the un-normalised probabilities, normalised at the end
    float O[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) O[nt][0] = O[nt][1] = O[nt][2] = O[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const uint32_t a0 = pack_bf16(S[2 * ks][0], S[2 * ks][1]), a1 = pack_bf16(S[2 * ks][2], S[2 * ks][3]);
      const uint32_t a2 = pack_bf16(S[2 * ks + 1][0], S[2 * ks + 1][1]), a3 = pack_bf16(S[2 * ks + 1][2], S[2 * ks + 1][3]);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const uint32_t b0 = Vt[(8 * nt + g) * kVStride + ks * 8 + tq], b1 = Vt[(8 * nt + g) * kVStride + ks * 8 + 4 + tq];
        mma16816(O[nt], a0, a1, a2, a3, b0, b1);
      }
    }
    const float i0 = 1.0f / sum0, i1 = 1.0f / sum1;
    __nv_bfloat16* d0 = out + tok_of[r0] * C + h * kHeadDim + 2 * tq;
    __nv_bfloat16* d1 = out + tok_of[r1] * C + h * kHeadDim + 2 * tq;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<uint32_t*>(d0 + 8 * nt) = pack_bf16(O[nt][0] * i0, O[nt][1] * i0);
      *reinterpret_cast<uint32_t*>(d1 + 8 * nt) = pack_bf16(O[nt][2] * i1, O[nt][3] * i1);
    }
  }
}

// one warp per row: y = shortcut + LayerNorm(h) * gamma + beta   (C = 32 * PER)
template <typename T, int PER>
__global__ void __launch_bounds__(256) ln_residual_kernel(const float* __restrict__ shortcut, const T* __restrict__ hsrc,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float* __restrict__ y, int64_t rows, float eps,
                                                          __nv_bfloat16* __restrict__ y_hi,
                                                          __nv_bfloat16* __restrict__ y_lo) {
  constexpr int C = 32 * PER;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* hp = hsrc + row * C;
  float v[PER];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = ldf(hp + lane + 32 * i);
    sum += v[i];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum / (float)C;
  float var = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float dlt = v[i] - mean;
    var = fmaf(dlt, dlt, var);
  }
  for (int o = 16; o > 0; o >>= 1) var += __shfl_xor_sync(0xffffffffu, var, o);
  const float rstd = rsqrtf(var / (float)C + eps);
  const float* sp = shortcut ? shortcut + row * C : nullptr;
  float* yp = y + row * C;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = lane + 32 * i;
    float r = (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
    if (sp) r += sp[c];
    yp[c] = r;
    if (y_hi) {  // the GEMM operand(s) of the next linear layer
      const __nv_bfloat16 hb = __float2bfloat16_rn(r);
      y_hi[row * C + c] = hb;
      if (y_lo) y_lo[row * C + c] = __float2bfloat16_rn(r - __bfloat162float(hb));
    }
  }
}

template <typename T>
cudaError_t launch_ln(const float* shortcut, const T* h, const float* gamma, const float* beta, float* y, int64_t rows,
                      int C, float eps, cudaStream_t stream, __nv_bfloat16* y_hi = nullptr, __nv_bfloat16* y_lo = nullptr) {
  const unsigned blocks = (unsigned)((rows + 7) / 8);
  switch (C / 32) {
    case 3: ln_residual_kernel<T, 3><<<blocks, 256, 0, stream>>>(shortcut, h, gamma, beta, y, rows, eps, y_hi, y_lo); break;
    case 6: ln_residual_kernel<T, 6><<<blocks, 256, 0, stream>>>(shortcut, h, gamma, beta, y, rows, eps, y_hi, y_lo); break;
    case 12: ln_residual_kernel<T, 12><<<blocks, 256, 0, stream>>>(shortcut, h, gamma, beta, y, rows, eps, y_hi, y_lo); break;
    case 24: ln_residual_kernel<T, 24><<<blocks, 256, 0, stream>>>(shortcut, h, gamma, beta, y, rows, eps, y_hi, y_lo); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace

extern "C" {

int pmnet_window_attention(const void* qkv, void* out, const float* logit_scale, const float* rel_bias,
                           const float* attn_mask, int32_t B, int32_t res, int32_t shift, int32_t heads,
                           int32_t is_bf16, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!qkv || !out || !logit_scale || !rel_bias) {
    pmnet_set_error("pmnet_window_attention: null argument");
    return PMNET_EINVAL;
  }
  if (B <= 0 || res < 4 || (res & 3) || heads <= 0 || shift < 0 || shift >= 4) {
    pmnet_set_error("pmnet_window_attention: resolution must be a multiple of the 4^3 window");
    return PMNET_EINVAL;
  }
  const int nw = (res / 4) * (res / 4) * (res / 4);
  dim3 grid((unsigned)(B * nw), (unsigned)heads);
  if (is_bf16)  // bf16 storage: tensor-core kernel; fp32 storage: the CUDA-core kernel keeps fp32 parity
    window_attention_mma_kernel<<<grid, 32, 0, stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, logit_scale,
                                                         rel_bias, attn_mask, res, shift, heads);
  else
    window_attention_kernel<float><<<grid, kTokens, 0, stream>>>((const float*)qkv, (float*)out, logit_scale, rel_bias,
                                                                 attn_mask, res, shift, heads, nullptr, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

// fp32 qkv in, result as the two-term bf16 split the projection GEMM consumes (split-precision backbone)
int pmnet_window_attention_split(const float* qkv, void* out_hi, void* out_lo, const float* logit_scale,
                                 const float* rel_bias, const float* attn_mask, int32_t B, int32_t res, int32_t shift,
                                 int32_t heads, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!qkv || !out_hi || !out_lo || !logit_scale || !rel_bias) {
    pmnet_set_error("pmnet_window_attention_split: null argument");
    return PMNET_EINVAL;
  }
  if (B <= 0 || res < 4 || (res & 3) || heads <= 0 || shift < 0 || shift >= 4) {
    pmnet_set_error("pmnet_window_attention_split: resolution must be a multiple of the 4^3 window");
    return PMNET_EINVAL;
  }
  const int nw = (res / 4) * (res / 4) * (res / 4);
  dim3 grid((unsigned)(B * nw), (unsigned)heads);
  window_attention_kernel<float><<<grid, kTokens, 0, stream>>>(qkv, nullptr, logit_scale, rel_bias, attn_mask, res, shift,
                                                               heads, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

int pmnet_ln_residual(const float* shortcut, const void* h, int32_t h_is_bf16, const float* gamma, const float* beta,
                      float* y, int64_t rows, int32_t C, float eps, void* stream_) {
  return pmnet_ln_residual_split(shortcut, h, h_is_bf16, gamma, beta, y, nullptr, nullptr, rows, C, eps, stream_);
}

int pmnet_ln_residual_split(const float* shortcut, const void* h, int32_t h_is_bf16, const float* gamma,
                            const float* beta, float* y, void* y_hi, void* y_lo, int64_t rows, int32_t C, float eps,
                            void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!h || !gamma || !beta || !y || (y_lo && !y_hi)) {
    pmnet_set_error("pmnet_ln_residual: null argument");
    return PMNET_EINVAL;
  }
  if (rows <= 0 || (C != 96 && C != 192 && C != 384 && C != 768)) {
    pmnet_set_error("pmnet_ln_residual: C must be 96, 192, 384 or 768 (the Swin stage widths)");
    return PMNET_EINVAL;
  }
  cudaError_t e = h_is_bf16 ? launch_ln<__nv_bfloat16>(shortcut, (const __nv_bfloat16*)h, gamma, beta, y, rows, C, eps, stream,
                                                       (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo)
                            : launch_ln<float>(shortcut, (const float*)h, gamma, beta, y, rows, C, eps, stream,
                                               (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo);
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

}  // extern "C"
