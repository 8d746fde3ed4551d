// pointwise.cu - the memory-bound glue around the tcgen05 convolutions of the CNN forward, one pass each over the
// 8-channel-chunk ("c8") bf16 activations. None of these is GEMM-heavy enough to be worth tensor cores; they are
// written for coalesced 16 B accesses and are HBM/L2 bound.
//
//  pmnet_lateral_c96      FPNDecoder lateral: 1x1 conv (C_in -> 96) + BatchNorm(eval) + ReLU, then
//                         "+ nearest-upsampled coarser level"      src/pmnet/network/decoders/fpn_decoder.py:100-111
//  pmnet_box_combine_c96  MaskHead.get_box_features folded through the (linear) lateral 1x1 conv of the mask-head
//                         decoder: the conv of the shared feature map is computed once per pocket, each box adds
//                         its background / point vectors                 src/pmnet/network/mask_head.py:170-196
//  pmnet_density_post     sigmoid -> mask -> 5^3 Gaussian (sigma 0.5 voxel, zero pad) -> mask -> threshold, with the
//                         spherical box area computed in place          src/pmnet/module.py:277-288,
//                         src/pmnet/utils/smoothing.py:17-71, src/pmnet/data/token_inference.py:118-146

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pmnet_b200.h"

extern void pmnet_set_error(const char* msg);

namespace {

__device__ __forceinline__ void unpack8(const uint4& p, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 p;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&p);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}

// y -> (hi, lo) with hi = bf16(y), lo = bf16(y - hi): the two-term split the convolution's extra passes consume
__device__ __forceinline__ void pack8_split(const float (&f)[8], uint4& hi, uint4& lo) {
  hi = pack8(f);
  float h[8], r[8];
  unpack8(hi, h);
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = f[i] - h[i];
  lo = pack8(r);
}
__device__ __forceinline__ void unpack8_sum(const uint4* __restrict__ hi, const uint4* __restrict__ lo, size_t i,
                                            float (&f)[8]) {
  unpack8(__ldg(hi + i), f);
  if (lo) {
    float g[8];
    unpack8(__ldg(lo + i), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] += g[j];
  }
}

constexpr int kLatVox = 128;   // voxels per block tile
constexpr int kLatTiles = 4;   // tiles per block (the weight tile is loaded once per 512 voxels)

// ------------------------------------------------------------------ lateral 1x1 conv
// grid: (ceil(V / 512), B); block 128 = 4 warps. Warp q owns output channels [24 q, 24 q + 24), lane l owns voxels
// l, l + 32, l + 64, l + 96 of the tile: 96 accumulators per thread, and every weight float4 read from shared memory
// (a warp-wide broadcast) feeds 16 FMAs. Weights [C_in][96] fp32 in smem.
template <int IN_C8>
__global__ void __launch_bounds__(128) lateral_kernel(const void* __restrict__ x, int cin, const float* __restrict__ wt,
                                                      const float* __restrict__ scale, const float* __restrict__ bias,
                                                      int relu, const uint4* __restrict__ up, uint4* __restrict__ out,
                                                      int D, int H, int W, const uint4* __restrict__ x_lo,
                                                      const uint4* __restrict__ up_lo, uint4* __restrict__ out_lo) {
  extern __shared__ __align__(16) float sw[];  // [cin][96]
  const int V = D * H * W;
  for (int i = threadIdx.x; i < cin * 96; i += blockDim.x) sw[i] = wt[i];
  __syncthreads();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, quarter = threadIdx.x >> 5;
  const int Vu = (D / 2) * (H / 2) * (W / 2);
  for (int tile = 0; tile < kLatTiles; ++tile) {
    const int v0 = (blockIdx.x * kLatTiles + tile) * kLatVox;
    if (v0 >= V) break;
    float acc[4][24];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 24; ++c) acc[i][c] = 0.0f;
    if (IN_C8) {
      const uint4* xp = reinterpret_cast<const uint4*>(x) + (size_t)b * (cin / 8) * V;
      for (int q = 0; q < cin / 8; ++q) {
        float f[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int v = v0 + lane + 32 * i;
          if (v < V) {
            unpack8_sum(xp, x_lo ? x_lo + (size_t)b * (cin / 8) * V : nullptr, (size_t)q * V + v, f[i]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[i][j] = 0.0f;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4* wr = reinterpret_cast<const float4*>(sw + (q * 8 + j) * 96 + quarter * 24);
#pragma unroll
          for (int c4 = 0; c4 < 6; ++c4) {
            const float4 w4 = wr[c4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[i][4 * c4 + 0] = fmaf(f[i][j], w4.x, acc[i][4 * c4 + 0]);
              acc[i][4 * c4 + 1] = fmaf(f[i][j], w4.y, acc[i][4 * c4 + 1]);
              acc[i][4 * c4 + 2] = fmaf(f[i][j], w4.z, acc[i][4 * c4 + 2]);
              acc[i][4 * c4 + 3] = fmaf(f[i][j], w4.w, acc[i][4 * c4 + 3]);
            }
          }
        }
      }
    } else {
      const float* xp = reinterpret_cast<const float*>(x) + (size_t)b * cin * V;
      for (int k0 = 0; k0 < cin; k0 += 4) {
        float xv[4][4];  // [channel of the group][voxel]: 16 independent loads in flight
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int v = v0 + lane + 32 * i;
            xv[j][i] = (k0 + j < cin && v < V) ? __ldg(xp + (size_t)(k0 + j) * V + v) : 0.0f;
          }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (k0 + j < cin) {
            const float4* wr = reinterpret_cast<const float4*>(sw + (k0 + j) * 96 + quarter * 24);
#pragma unroll
            for (int c4 = 0; c4 < 6; ++c4) {
              const float4 w4 = wr[c4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[i][4 * c4 + 0] = fmaf(xv[j][i], w4.x, acc[i][4 * c4 + 0]);
                acc[i][4 * c4 + 1] = fmaf(xv[j][i], w4.y, acc[i][4 * c4 + 1]);
                acc[i][4 * c4 + 2] = fmaf(xv[j][i], w4.z, acc[i][4 * c4 + 2]);
                acc[i][4 * c4 + 3] = fmaf(xv[j][i], w4.w, acc[i][4 * c4 + 3]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int v = v0 + lane + 32 * i;
      if (v >= V) continue;
      const int w = v % W, h = (v / W) % H, d = v / (W * H);
      const int vu = ((d >> 1) * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        const int chunk = quarter * 3 + q;
        float y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = chunk * 8 + j;
          float t = acc[i][q * 8 + j];
          if (scale) t = fmaf(t, __ldg(scale + c), __ldg(bias + c));
          if (relu) t = fmaxf(t, 0.0f);
          y[j] = t;
        }
        if (up) {
          float u[8];
          unpack8_sum(up, up_lo, ((size_t)b * 12 + chunk) * Vu + vu, u);
#pragma unroll
          for (int j = 0; j < 8; ++j) y[j] += u[j];
        }
        if (out_lo) {
          uint4 ph, pl;
          pack8_split(y, ph, pl);
          out[((size_t)b * 12 + chunk) * V + v] = ph;
          out_lo[((size_t)b * 12 + chunk) * V + v] = pl;
        } else {
          out[((size_t)b * 12 + chunk) * V + v] = pack8(y);
        }
      }
    }
  }
}

// ------------------------------------------------------------------ lateral 1x1 conv on the tensor cores (bf16 mode)
// The two large FPN laterals (64^3 x 33 -> 96 and 32^3 x 96 -> 96, fp32 NCDHW inputs from the image / the backbone) are
// 13 + 5 GFLOP per 8 pockets: on CUDA cores (lateral_kernel<0>) they ran at ~20 % of the fp32 peak, 1.1 ms of the
// 8.5 ms forward. Here a 128-voxel tile of x is transposed to bf16 [voxel][k] in shared memory and multiplied by the
// bf16 weights with mma.sync m16n8k16 (fp32 accumulate); an n8 tile is exactly one 8-channel chunk of the c8 output,
// so the accumulator fragments are stored (scale / bias / ReLU / + upsampled coarser level) without any shuffle. The
// kernel is bound by its 0.7 GB of traffic, not by the 72 HMMAs per warp tile. Used when neither input nor output
// carries a low part: the split-precision mode keeps the fp32 CUDA-core kernel.
constexpr int kLmVox = 128;  // voxels per tile = 4 warps x 32

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], const uint32_t b0, const uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// IN_C8: x is a bf16 c8 activation (C_in = 16 KT exactly): the A fragments are read straight from global memory - 8
// consecutive voxels of one chunk are 128 contiguous bytes, 4 lanes per voxel - and nothing is staged.
template <int KT, bool IN_C8 = false>  // K tiles of 16: C_in <= 16 KT
__global__ void __launch_bounds__(128, 4) lateral_mma_kernel(const float* __restrict__ x, int cin, const float* __restrict__ wt,
                                                          const float* __restrict__ scale, const float* __restrict__ bias,
                                                          int relu, const uint32_t* __restrict__ up,
                                                          uint32_t* __restrict__ out, int D, int H, int W, int tiles_per_block) {
  constexpr int KP = 16 * KT + 2;  // row pitch in bf16: KP / 2 odd -> conflict-free transposing stores
  __shared__ __align__(16) __nv_bfloat16 Wt[96 * KP];     // [n][k]
  __shared__ __align__(16) __nv_bfloat16 Xs[IN_C8 ? 8 : kLmVox * KP];  // [voxel][k] (fp32 NCDHW input only)
  const int V = D * H * W, Vu = (D / 2) * (H / 2) * (W / 2);
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, c = lane & 3;
  for (int i = threadIdx.x; i < 96 * 16 * KT; i += blockDim.x) {  // (coalesced reads of wt[k][n])
    const int k = i / 96, n = i % 96;
    Wt[n * KP + k] = __float2bfloat16(k < cin ? wt[i] : 0.0f);
  }
  const float* xb = x + (size_t)b * cin * V;
  for (int tile = 0; tile < tiles_per_block; ++tile) {
    const int v0 = (blockIdx.x * tiles_per_block + tile) * kLmVox;
    if (v0 >= V) break;
    __syncthreads();  // Wt written / the previous tile's Xs consumed
    if (!IN_C8) {
      const int v = v0 + threadIdx.x;
      __nv_bfloat162* row = reinterpret_cast<__nv_bfloat162*>(Xs + threadIdx.x * KP);
#pragma unroll 4
      for (int k = 0; k < 16 * KT; k += 2) {
        const float f0 = (k < cin && v < V) ? __ldg(xb + (size_t)k * V + v) : 0.0f;
        const float f1 = (k + 1 < cin && v < V) ? __ldg(xb + (size_t)(k + 1) * V + v) : 0.0f;
        row[k >> 1] = __floats2bfloat162_rn(f0, f1);
      }
    }
    if (!IN_C8) __syncthreads();
    float acc[2][12][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 12; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.0f;
    const uint32_t* Xw = reinterpret_cast<const uint32_t*>(Xs);
    const uint32_t* Ww = reinterpret_cast<const uint32_t*>(Wt);
#pragma unroll
    for (int kt = 0; kt < KT; ++kt) {
      uint32_t a[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r = warp * 32 + mt * 16 + g;
        if (IN_C8) {
          const uint32_t* x32 = reinterpret_cast<const uint32_t*>(x) + ((size_t)b * (2 * KT) + 2 * kt) * V * 4;
          const int va = v0 + r, vb = va + 8;
          a[mt][0] = va < V ? __ldg(x32 + (size_t)va * 4 + c) : 0u;
          a[mt][1] = vb < V ? __ldg(x32 + (size_t)vb * 4 + c) : 0u;
          a[mt][2] = va < V ? __ldg(x32 + ((size_t)V + va) * 4 + c) : 0u;
          a[mt][3] = vb < V ? __ldg(x32 + ((size_t)V + vb) * 4 + c) : 0u;
          continue;
        }
        a[mt][0] = Xw[(r * KP + kt * 16 + 2 * c) >> 1];
        a[mt][1] = Xw[((r + 8) * KP + kt * 16 + 2 * c) >> 1];
        a[mt][2] = Xw[(r * KP + kt * 16 + 8 + 2 * c) >> 1];
        a[mt][3] = Xw[((r + 8) * KP + kt * 16 + 8 + 2 * c) >> 1];
      }
#pragma unroll
      for (int nt = 0; nt < 12; ++nt) {
        const int n = nt * 8 + g;
        const uint32_t b0 = Ww[(n * KP + kt * 16 + 2 * c) >> 1];
        const uint32_t b1 = Ww[(n * KP + kt * 16 + 8 + 2 * c) >> 1];
        mma_bf16_16816(acc[0][nt], a[0], b0, b1);
        mma_bf16_16816(acc[1][nt], a[1], b0, b1);
      }
    }
    // epilogue: fragment (row g / g + 8, columns 2c, 2c + 1 of n tile nt) -> channels 8 nt + 2c, + 1 of that voxel
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int v = v0 + warp * 32 + mt * 16 + g + 8 * half;
        if (v >= V) continue;
        const int w = v % W, h = (v / W) % H, d = v / (W * H);
        const int vu = ((d >> 1) * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
#pragma unroll
        for (int nt = 0; nt < 12; ++nt) {
          const int ch = nt * 8 + 2 * c;
          float y0 = acc[mt][nt][2 * half], y1 = acc[mt][nt][2 * half + 1];
          if (scale) {
            y0 = fmaf(y0, __ldg(scale + ch), __ldg(bias + ch));
            y1 = fmaf(y1, __ldg(scale + ch + 1), __ldg(bias + ch + 1));
          }
          if (relu) {
            y0 = fmaxf(y0, 0.0f);
            y1 = fmaxf(y1, 0.0f);
          }
          if (up) {
            const uint32_t u = __ldg(up + (((size_t)b * 12 + nt) * Vu + vu) * 4 + c);
            const float2 uf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u));
            y0 += uf.x;
            y1 += uf.y;
          }
          const __nv_bfloat162 o = __floats2bfloat162_rn(y0, y1);
          out[(((size_t)b * 12 + nt) * V + v) * 4 + c] = *reinterpret_cast<const uint32_t*>(&o);
        }
      }
  }
}

// ------------------------------------------------------------------ per-box combine
// out[j] = act(scale * (S + u_j + [v in P_j] p_j) + bias) + up_j ; thread = (voxel, chunk), loops over boxes.
// P_j = up to 4 voxel ids per box (-1 padded): the token voxels of the box's group of 4 (mask_head.py:190-194).
__global__ void __launch_bounds__(256) box_combine_kernel(const uint4* __restrict__ S, const float* __restrict__ u,
                                                          const float* __restrict__ pvec, const int4* __restrict__ pvox,
                                                          const float* __restrict__ scale,
                                                          const float* __restrict__ bias, int relu,
                                                          const uint4* __restrict__ up, uint4* __restrict__ out,
                                                          int nbox, int D, int H, int W, const uint4* __restrict__ S_lo,
                                                          const uint4* __restrict__ up_lo, uint4* __restrict__ out_lo) {
  const int V = D * H * W;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;  // chunk * V + v
  if (idx >= (size_t)12 * V) return;
  const int chunk = (int)(idx / V), v = (int)(idx % V);
  float s[8];
  unpack8_sum(S, S_lo, idx, s);
  float sc[8], bi[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    sc[k] = scale ? __ldg(scale + chunk * 8 + k) : 1.0f;
    bi[k] = scale ? __ldg(bias + chunk * 8 + k) : 0.0f;
  }
  const int w = v % W, h = (v / W) % H, d = v / (W * H);
  const int Vu = (D / 2) * (H / 2) * (W / 2);
  const int vu = ((d >> 1) * (H / 2) + (h >> 1)) * (W / 2) + (w >> 1);
  for (int j = 0; j < nbox; ++j) {
    const int4 pv = __ldg(pvox + j);
    const bool at_point = (pv.x == v) | (pv.y == v) | (pv.z == v) | (pv.w == v);
    float y[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = chunk * 8 + k;
      float t = s[k] + __ldg(u + j * 96 + c);
      if (at_point) t += __ldg(pvec + j * 96 + c);
      t = fmaf(t, sc[k], bi[k]);
      if (relu) t = fmaxf(t, 0.0f);
      y[k] = t;
    }
    if (up) {
      float uu[8];
      unpack8_sum(up, up_lo, ((size_t)j * 12 + chunk) * Vu + vu, uu);
#pragma unroll
      for (int k = 0; k < 8; ++k) y[k] += uu[k];
    }
    if (out_lo) {
      uint4 ph, pl;
      pack8_split(y, ph, pl);
      out[((size_t)j * 12 + chunk) * V + v] = ph;
      out_lo[((size_t)j * 12 + chunk) * V + v] = pl;
    } else {
      out[((size_t)j * 12 + chunk) * V + v] = pack8(y);
    }
  }
}

// ------------------------------------------------------------------ density-map post-processing
// INTERACTION_DIST (src/pmnet/data/constant.py:28-39) -> ceil((dist + 1.0) / 0.5) voxels
__constant__ int kBoxRadius[10] = {11, 14, 14, 15, 15, 11, 11, 14, 14, 11};

__device__ __forceinline__ bool available(int x, int y, int z, int tx, int ty, int tz, int r2, const uint8_t* prot,
                                          const uint8_t* cav, int S) {
  const int dx = x - tx, dy = y - ty, dz = z - tz;
  if (dx * dx + dy * dy + dz * dz >= r2) return false;  // |g - t| < threshold (token_inference.py:144-145)
  const int v = (x * S + y) * S + z;
  return prot[v] && cav[v];
}

// grid: (S^3 / 256, n maps). logits [n][S^3]; tokens int32 [n][4] (x, y, z, type); masks uint8 [S^3]; w5[5] = 1-D
// normalised Gaussian taps (the 3-D kernel of smoothing.py is their outer product).
__global__ void __launch_bounds__(256) density_post_kernel(const float* __restrict__ logits, const int* __restrict__ tokens,
                                                           const uint8_t* __restrict__ prot, const uint8_t* __restrict__ cav,
                                                           float w0, float w1, float w2, float threshold,
                                                           float* __restrict__ out, int S) {
  const int n = blockIdx.y;
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  const int V = S * S * S;
  if (v >= V) return;
  const int tx = tokens[n * 4 + 0], ty = tokens[n * 4 + 1], tz = tokens[n * 4 + 2], tt = tokens[n * 4 + 3];
  const int r = kBoxRadius[tt], r2 = r * r;
  const int z = v % S, y = (v / S) % S, x = v / (S * S);
  float res = 0.0f;
  if (available(x, y, z, tx, ty, tz, r2, prot, cav, S)) {
    const float wk[5] = {w0, w1, w2, w1, w0};
    const float* lg = logits + (size_t)n * V;
    float acc = 0.0f;
    for (int a = -2; a <= 2; ++a) {
      const int xa = x + a;
      if (xa < 0 || xa >= S) continue;
      for (int bq = -2; bq <= 2; ++bq) {
        const int yb = y + bq;
        if (yb < 0 || yb >= S) continue;
        const float wab = wk[a + 2] * wk[bq + 2];
        for (int c = -2; c <= 2; ++c) {
          const int zc = z + c;
          if (zc < 0 || zc >= S) continue;
          if (!available(xa, yb, zc, tx, ty, tz, r2, prot, cav, S)) continue;
          const float p = 1.0f / (1.0f + expf(-lg[(xa * S + yb) * S + zc]));
          acc = fmaf(wab * wk[c + 2], p, acc);
        }
      }
    }
    res = (acc < threshold) ? 0.0f : acc;
  }
  out[(size_t)n * V + v] = res;
}

}  // namespace

extern "C" {

int pmnet_lateral_c96(const void* x, int32_t x_is_c8, int32_t c_in, const float* w_t, const float* scale,
                      const float* bias, int32_t relu, const void* up_c8, void* out_c8, int32_t B, int32_t D,
                      int32_t H, int32_t W, void* stream_) {
  return pmnet_lateral_c96_split(x, nullptr, x_is_c8, c_in, w_t, scale, bias, relu, up_c8, nullptr, out_c8, nullptr, B, D, H,
                                 W, stream_);
}

int pmnet_lateral_c96_split(const void* x, const void* x_lo, int32_t x_is_c8, int32_t c_in, const float* w_t,
                            const float* scale, const float* bias, int32_t relu, const void* up_c8, const void* up_lo_c8,
                            void* out_c8, void* out_lo_c8, int32_t B, int32_t D, int32_t H, int32_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if ((x_lo && !x_is_c8) || (up_lo_c8 && !up_c8)) {
    pmnet_set_error("pmnet_lateral_c96_split: a low part needs its high part in c8 layout");
    return PMNET_EINVAL;
  }
  if (!x || !w_t || !out_c8 || (scale && !bias)) {
    pmnet_set_error("pmnet_lateral_c96: null argument");
    return PMNET_EINVAL;
  }
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || c_in <= 0 || (x_is_c8 && (c_in & 7)) || (up_c8 && ((D | H | W) & 1))) {
    pmnet_set_error("pmnet_lateral_c96: bad shape");
    return PMNET_EINVAL;
  }
  const size_t smem = (size_t)c_in * 96 * 4;
  if (smem > 200 * 1024) {
    pmnet_set_error("pmnet_lateral_c96: C_in too large for the shared-memory weight tile");
    return PMNET_ELIMIT;
  }
  const int V = D * H * W;
  dim3 grid((V + kLatVox * kLatTiles - 1) / (kLatVox * kLatTiles), B);
  cudaError_t e;
  if (x_is_c8 && !x_lo && !out_lo_c8 && !up_lo_c8 && c_in == 96 && V >= 8192) {
    // bf16 mode, c8 input (the mask head's shared laterals): tensor cores, fragments straight from global memory
    const int tiles = V >= (1 << 18) ? 2 : 1;
    dim3 mgrid((V + kLmVox * tiles - 1) / (kLmVox * tiles), B);
    e = cudaSuccess;
    lateral_mma_kernel<6, true><<<mgrid, 128, 0, stream>>>((const float*)x, c_in, w_t, scale, bias, relu,
                                                           (const uint32_t*)up_c8, (uint32_t*)out_c8, D, H, W, tiles);
  } else if (x_is_c8) {
    e = cudaFuncSetAttribute(lateral_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      lateral_kernel<1><<<grid, 128, smem, stream>>>(x, c_in, w_t, scale, bias, relu, (const uint4*)up_c8, (uint4*)out_c8,
                                                     D, H, W, (const uint4*)x_lo, (const uint4*)up_lo_c8,
                                                     (uint4*)out_lo_c8);
  } else if (!out_lo_c8 && !up_lo_c8 && c_in <= 96 && V >= 8192) {
    // bf16 mode, large level: tensor cores (lateral_mma_kernel)
    const int tiles = V >= (1 << 18) ? 2 : 1;  // enough blocks for several waves at every level
    dim3 mgrid((V + kLmVox * tiles - 1) / (kLmVox * tiles), B);
    const float* xf = (const float*)x;
    e = cudaSuccess;
    if (c_in <= 48)
      lateral_mma_kernel<3><<<mgrid, 128, 0, stream>>>(xf, c_in, w_t, scale, bias, relu, (const uint32_t*)up_c8,
                                                       (uint32_t*)out_c8, D, H, W, tiles);
    else
      lateral_mma_kernel<6><<<mgrid, 128, 0, stream>>>(xf, c_in, w_t, scale, bias, relu, (const uint32_t*)up_c8,
                                                       (uint32_t*)out_c8, D, H, W, tiles);
  } else {
    e = cudaFuncSetAttribute(lateral_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      lateral_kernel<0><<<grid, 128, smem, stream>>>(x, c_in, w_t, scale, bias, relu, (const uint4*)up_c8, (uint4*)out_c8,
                                                     D, H, W, nullptr, (const uint4*)up_lo_c8, (uint4*)out_lo_c8);
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

int pmnet_box_combine_c96(const void* s_c8, const float* u, const float* pvec, const int32_t* pvox,
                          const float* scale, const float* bias, int32_t relu, const void* up_c8, void* out_c8,
                          int32_t nbox, int32_t D, int32_t H, int32_t W, void* stream_) {
  return pmnet_box_combine_c96_split(s_c8, nullptr, u, pvec, pvox, scale, bias, relu, up_c8, nullptr, out_c8, nullptr, nbox,
                                     D, H, W, stream_);
}

int pmnet_box_combine_c96_split(const void* s_c8, const void* s_lo_c8, const float* u, const float* pvec,
                                const int32_t* pvox, const float* scale, const float* bias, int32_t relu,
                                const void* up_c8, const void* up_lo_c8, void* out_c8, void* out_lo_c8, int32_t nbox,
                                int32_t D, int32_t H, int32_t W, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!s_c8 || !u || !out_c8 || !pvec || !pvox || (scale && !bias)) {
    pmnet_set_error("pmnet_box_combine_c96: null argument");
    return PMNET_EINVAL;
  }
  if (nbox <= 0 || D <= 0 || H <= 0 || W <= 0 || (up_c8 && ((D | H | W) & 1))) {
    pmnet_set_error("pmnet_box_combine_c96: bad shape");
    return PMNET_EINVAL;
  }
  const size_t total = (size_t)12 * D * H * W;
  box_combine_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
      (const uint4*)s_c8, u, pvec, (const int4*)pvox, scale, bias, relu, (const uint4*)up_c8, (uint4*)out_c8, nbox, D, H,
      W, (const uint4*)s_lo_c8, (const uint4*)up_lo_c8, (uint4*)out_lo_c8);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

int pmnet_density_post(const float* logits, const int32_t* tokens, const uint8_t* protein_mask,
                       const uint8_t* cavity_mask, const float* taps3, float threshold, float* out, int32_t n,
                       int32_t size, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!logits || !tokens || !protein_mask || !cavity_mask || !taps3 || !out) {
    pmnet_set_error("pmnet_density_post: null argument");
    return PMNET_EINVAL;
  }
  if (n <= 0 || size <= 0) {
    pmnet_set_error("pmnet_density_post: bad shape");
    return PMNET_EINVAL;
  }
  const int V = size * size * size;
  dim3 grid((V + 255) / 256, n);
  density_post_kernel<<<grid, 256, 0, stream>>>(logits, tokens, protein_mask, cavity_mask, taps3[0], taps3[1], taps3[2],
                                                threshold, out, size);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}

}  // extern "C"
