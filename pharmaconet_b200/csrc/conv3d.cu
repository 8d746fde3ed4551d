// conv3d.cu - 3x3x3 convolution, 96 -> 96 channels, stride 1, zero padding 1, with the folded BatchNorm(eval) scale /
// bias and ReLU in the epilogue, as an implicit GEMM on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// This one shape is >85 % of the FLOPs of the reference's CNN forward:
//   BaseConv3d.forward = act(norm(conv(x)))                src/pmnet/network/nn/layers.py:45-46
//   FPNDecoder.fpn_convs_list (levels 64^3, 32^3, 16^3)    src/pmnet/network/decoders/fpn_decoder.py:54-66, 112
//   CavityHead short/long heads (+ 1x1 -> 1 logits)        src/pmnet/network/cavity_head.py:18-37, 57-59
//   MaskHead decoder (same FPNDecoder)                     src/pmnet/network/mask_head.py:38-80
//
// Layouts
//   activations  bf16 [B][C/8][D][H][W][8]   ("c8": 8-channel chunks are the unit, 16 B per voxel and chunk)
//   weights      bf16 [27 taps][C_in/8][C_out][8]   = the UMMA no-swizzle K-major image of B for every tap
// Implicit GEMM: M = 128 output voxels = one 16 (h) x 8 (w) tile of one d-plane, N = 96, K = 27 taps x 96.
// A halo plane (18 x 10 voxels x 96 ch = 34 560 B) is fetched by ONE 5-D TMA box; in shared memory it is
// [chunk][h 18][w 10][16 B], so for any tap (kd,kh,kw) the A operand of a K = 16 step is a plain no-swizzle K-major
// UMMA matrix: 8 consecutive w voxels are one 128 B core matrix, SBO = 160 B (next h row), LBO = 2880 B (next
// chunk) - shifted windows need no data movement. A persistent CTA marches along d: every input plane is loaded once
// per (h, w) tile column and lives in a 4-slot ring; two consecutive output planes share every weight tap
// (G = 2) to halve the weight stream from L2. Warp roles: warp 0 TMA producer, warps 1-2 MMA issuers (one thread
// each, one output plane each), warps 3-6 epilogue (TMEM -> registers -> scale/bias/ReLU -> bf16 -> global). Accumulators are double buffered in
// TMEM (4 x 128 columns) so the epilogue of a plane pair overlaps the MMAs of the next.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pmnet_b200.h"

namespace pmconv {

constexpr int kC = 96;               // channels in and out
constexpr int kChunks = kC / 8;      // 12
constexpr int kTileH = 16, kTileW = 8;
constexpr int kHaloH = kTileH + 2, kHaloW = kTileW + 2;
constexpr int kRowBytes = kHaloW * 16;                       // 160: one halo row of one chunk
constexpr int kChunkBytes = kHaloH * kRowBytes;              // 2880
constexpr int kPlaneBytes = kChunks * kChunkBytes;           // 34560
constexpr int kPlaneSlots = 4;
constexpr int kTapBytes = kChunks * kC * 16;                 // 18432: weights of one tap
constexpr int kWStages = 4;
constexpr int kTaps = 27;
constexpr int kKSteps = kC / 16;                             // 6 MMAs (K = 16) per tap and accumulator
constexpr int kThreads = 224;  // TMA producer, 2 MMA issuers, 4 epilogue warps
constexpr uint32_t kTmemCols = 512;                          // 2 buffers x 2 planes x 128-column slots

constexpr int kSmemPlanes = 0;
constexpr int kSmemWeights = kSmemPlanes + kPlaneSlots * kPlaneBytes;   // 138240
constexpr int kSmemBars = kSmemWeights + kWStages * kTapBytes;          // 211968
constexpr int kNumBars = 2 * kPlaneSlots + 2 * kWStages + 4;
constexpr int kSmemTmemPtr = kSmemBars + kNumBars * 8;
constexpr int kSmemBytes = kSmemTmemPtr + 16;

struct Params {
  int B, D, H, W;
  int dc;                 // output planes per work item (even)
  int n_hb, n_wb, n_dc;   // tile grid
  int n_items;
  const __nv_bfloat16* wpacked;
  const float* scale;     // [96] folded BN scale (1 if none)
  const float* bias;      // [96] folded BN bias / conv bias
  __nv_bfloat16* out;     // c8 layout, may be null when only the fused head is wanted
  int relu;
  const float* head_w;    // optional fused 1x1 conv to one channel (cavity logits), fp32 [96]
  float head_b;
  float* head_out;        // [B][D][H][W] fp32
  // split-precision passes (x = x_hi + x_lo, w = w_hi + w_lo; y = x_hi w_hi + x_lo w_hi + x_hi w_lo as three launches):
  const float* acc_in;    // fp32 [B][D][H][W][96] partial sums of the earlier passes, added to the accumulator; or null
  float* acc_out;         // not null: store the raw sum here (may alias acc_in) and skip activation / outputs
  __nv_bfloat16* out_lo;  // optional second output, c8: bf16(y - float(bf16(y))), the low part of the activation
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
      "selp.b32 %0, 1, 0, P1;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t addr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(addr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// no-swizzle K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// [0,14) start >> 4, [16,30) leading (K-direction) byte offset >> 4, [32,46) stride (8-row group) byte offset >> 4,
// [46,48) version = 1, [61,64) layout type = 0 (SWIZZLE_NONE)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kC >> 3) << 17) | ((128u >> 4) << 24);

struct Item {
  int b, h0, w0, d_begin, d_end;
};
__device__ __forceinline__ Item decode_item(const Params& p, int it) {
  Item r;
  const int wb = it % p.n_wb;
  it /= p.n_wb;
  const int hb = it % p.n_hb;
  it /= p.n_hb;
  const int dcx = it % p.n_dc;
  r.b = it / p.n_dc;
  r.h0 = hb * kTileH;
  r.w0 = wb * kTileW;
  r.d_begin = dcx * p.dc;
  r.d_end = min(p.D, r.d_begin + p.dc);
  return r;
}

__global__ void __launch_bounds__(kThreads, 1) conv3d_k3_c96_kernel(const __grid_constant__ CUtensorMap tmap,
                                                                    const Params p) {
  extern __shared__ __align__(128) unsigned char smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // barriers
  auto bar = [&](int i) { return sbase + kSmemBars + 8u * i; };
  const int BAR_PFULL = 0, BAR_PEMPTY = kPlaneSlots, BAR_WFULL = 2 * kPlaneSlots, BAR_WEMPTY = 2 * kPlaneSlots + kWStages,
            BAR_AFULL = 2 * kPlaneSlots + 2 * kWStages, BAR_AEMPTY = BAR_AFULL + 2;
  volatile uint32_t* tmem_ptr_smem = (volatile uint32_t*)(smem + kSmemTmemPtr);

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPlaneSlots; ++i) {
      mbar_init(bar(BAR_PFULL + i), 1);
      mbar_init(bar(BAR_PEMPTY + i), 2);  // both MMA issuers
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(bar(BAR_WFULL + i), 1);
      mbar_init(bar(BAR_WEMPTY + i), 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(BAR_AFULL + i), 2);   // both MMA issuers
      mbar_init(bar(BAR_AEMPTY + i), 4);  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + kSmemTmemPtr),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // =============================================================== TMA producer
    if (lane == 0) {
      uint32_t pseq = 0, wseq = 0;  // planes / weight taps issued so far
      auto load_plane = [&](const Item& it, int d) {
        const uint32_t slot = pseq % kPlaneSlots, round = pseq / kPlaneSlots;
        mbar_wait(bar(BAR_PEMPTY + slot), (round & 1) ^ 1);
        mbar_expect_tx(bar(BAR_PFULL + slot), kPlaneBytes);
        tma_load_5d(sbase + kSmemPlanes + slot * kPlaneBytes, &tmap, bar(BAR_PFULL + slot), (it.w0 - 1) * 8, it.h0 - 1, d,
                    0, it.b);
        ++pseq;
      };
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        for (int d0 = it.d_begin; d0 < it.d_end; d0 += 2) {
          if (d0 == it.d_begin) {
            load_plane(it, d0 - 1);
            load_plane(it, d0);
          }
          load_plane(it, d0 + 1);
          load_plane(it, d0 + 2);
          for (int tap = 0; tap < kTaps; ++tap) {
            const uint32_t st = wseq % kWStages, round = wseq / kWStages;
            mbar_wait(bar(BAR_WEMPTY + st), (round & 1) ^ 1);
            mbar_expect_tx(bar(BAR_WFULL + st), kTapBytes);
            bulk_load(sbase + kSmemWeights + st * kTapBytes, (const unsigned char*)p.wpacked + (size_t)tap * kTapBytes,
                      kTapBytes, bar(BAR_WFULL + st));
            ++wseq;
          }
        }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // =============================================================== MMA issuers: warp 1 -> output plane d0 (g = 0),
    // warp 2 -> output plane d0 + 1 (g = 1); one thread each. Two issuers because one thread cannot feed the tensor
    // pipe: per tap it pays a barrier poll (~90 cycles), a fence, a commit and ~30 cycles per MMA for 12 MMAs of 48.
    if (lane == 0) {
      const int g = warp - 1;
      uint32_t pseq0 = 0;  // sequence number of plane (d_begin - 1) of the current item
      uint32_t wseq = 0, gseq = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        const Item it = decode_item(p, item);
        auto plane_seq = [&](int d) { return pseq0 + (uint32_t)(d - (it.d_begin - 1)); };
        auto wait_plane = [&](int d) {
          const uint32_t s = plane_seq(d);
          mbar_wait(bar(BAR_PFULL + s % kPlaneSlots), (s / kPlaneSlots) & 1);
        };
        // a plane slot is free again when BOTH issuers have retired their last MMA on it (barrier count 2)
        auto release_plane = [&](int d) { tc_commit(bar(BAR_PEMPTY + plane_seq(d) % kPlaneSlots)); };
        for (int d0 = it.d_begin; d0 < it.d_end; d0 += 2, ++gseq) {
          const uint32_t buf = gseq & 1;
          mbar_wait(bar(BAR_AEMPTY + buf), ((gseq >> 1) & 1) ^ 1);
          tc_fence_after();
          // descriptors are {hi, lo}: hi is constant per operand, lo = start >> 4 | (LBO >> 4) << 16, so moving the
          // window by a tap or a K step is one 32-bit add with an immediate
          constexpr uint32_t kAHi = (uint32_t)(kRowBytes >> 4) | (1u << 14);
          constexpr uint32_t kBHi = (uint32_t)(128 >> 4) | (1u << 14);
          const uint32_t tacc = tmem_base + (buf * 2 + g) * 128;
          if (g == 1) release_plane(d0 - 1);  // never read by this issuer
#pragma unroll 1
          for (int kd = 0; kd < 3; ++kd) {
            const int dplane = d0 + g + kd - 1;  // input plane of this issuer's output plane for this kd
            wait_plane(dplane);
            const uint32_t slot = plane_seq(dplane) % kPlaneSlots;
            const uint32_t a_lo = ((sbase + kSmemPlanes + slot * kPlaneBytes) >> 4) | ((uint32_t)(kChunkBytes >> 4) << 16);
#pragma unroll
            for (int khw = 0; khw < 9; ++khw) {
              const int kh = khw / 3, kw = khw % 3;
              const uint32_t st = wseq & (kWStages - 1);
              mbar_wait(bar(BAR_WFULL + st), (wseq / kWStages) & 1);
              ++wseq;
              tc_fence_after();
              const uint32_t b_lo = ((sbase + kSmemWeights + st * kTapBytes) >> 4) | ((uint32_t)((kC * 16) >> 4) << 16);
#pragma unroll
              for (int ks = 0; ks < kKSteps; ++ks) {
                const uint32_t al = a_lo + (uint32_t)((kh * kRowBytes + kw * 16 + 2 * ks * kChunkBytes) >> 4);
                const uint32_t bl = b_lo + (uint32_t)((2 * ks * kC * 16) >> 4);
                const uint64_t ad = ((uint64_t)kAHi << 32) | al;
                const uint64_t bd = ((uint64_t)kBHi << 32) | bl;
                tc_mma(tacc, ad, bd, kIdesc, (kd | khw | ks) != 0);
              }
              tc_commit(bar(BAR_WEMPTY + st));
            }
            // input planes d0-1 and d0 are not needed by later groups: g = 0 read them at kd = 0 / 1, g = 1 read d0 at
            // kd = 0. Planes d0+1 and d0+2 stay for the next group unless this is the item's last group.
            if (g == 0 && kd <= 1) release_plane(d0 - 1 + kd);
            if (g == 1 && kd == 0) release_plane(d0);
          }
          if (d0 + 2 >= it.d_end) {
            release_plane(d0 + 1);
            release_plane(d0 + 2);
          }
          tc_commit(bar(BAR_AFULL + buf));
        }
        pseq0 += (uint32_t)(it.d_end - it.d_begin + 2);
      }
    }
  } else {
    // =============================================================== epilogue warps (TMEM lane quadrant = warp % 4)
    const int quad = warp & 3;
    const int m = quad * 32 + lane;       // accumulator row = voxel of the tile
    const int th = m >> 3, tw = m & 7;
    uint32_t gseq = 0;
    const size_t plane_vox = (size_t)p.H * p.W;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Item it = decode_item(p, item);
      const int h = it.h0 + th, w = it.w0 + tw;
      const bool inside = (h < p.H) && (w < p.W);
      for (int d0 = it.d_begin; d0 < it.d_end; d0 += 2, ++gseq) {
        const uint32_t buf = gseq & 1;
        mbar_wait(bar(BAR_AFULL + buf), (gseq >> 1) & 1);
        tc_fence_after();
#pragma unroll 1
        for (int g = 0; g < 2; ++g) {
          const int d = d0 + g;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (buf * 2 + g) * 128;
          float head = 0.0f;
          const size_t vox96 = ((((size_t)it.b * p.D + d) * p.H + h) * p.W + w) * kC;  // fp32 partial-sum layout
#pragma unroll 1
          for (int c32 = 0; c32 < 3; ++c32) {
            uint32_t v[32];
            tmem_ld32(taddr + c32 * 32, v);
            if (p.acc_in && inside) {
              const float4* a4 = reinterpret_cast<const float4*>(p.acc_in + vox96 + c32 * 32);
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                const float4 t = a4[q];
                v[4 * q + 0] = __float_as_uint(__uint_as_float(v[4 * q + 0]) + t.x);
                v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + t.y);
                v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + t.z);
                v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + t.w);
              }
            }
            if (p.acc_out) {
              if (inside) {
                float4* o4 = reinterpret_cast<float4*>(p.acc_out + vox96 + c32 * 32);
#pragma unroll
                for (int q = 0; q < 8; ++q)
                  o4[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                      __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
              }
              continue;
            }
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = c32 * 32 + j;
              float t = fmaf(__uint_as_float(v[j]), __ldg(p.scale + c), __ldg(p.bias + c));
              if (p.relu) t = fmaxf(t, 0.0f);
              y[j] = t;
            }
            if (p.head_out) {
#pragma unroll
              for (int j = 0; j < 32; ++j) head = fmaf(y[j], __ldg(p.head_w + c32 * 32 + j), head);
            }
            if (p.out && inside) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int chunk = c32 * 4 + q;
                __nv_bfloat162 b0 = __floats2bfloat162_rn(y[q * 8 + 0], y[q * 8 + 1]);
                __nv_bfloat162 b1 = __floats2bfloat162_rn(y[q * 8 + 2], y[q * 8 + 3]);
                __nv_bfloat162 b2 = __floats2bfloat162_rn(y[q * 8 + 4], y[q * 8 + 5]);
                __nv_bfloat162 b3 = __floats2bfloat162_rn(y[q * 8 + 6], y[q * 8 + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&b0);
                pk.y = *reinterpret_cast<uint32_t*>(&b1);
                pk.z = *reinterpret_cast<uint32_t*>(&b2);
                pk.w = *reinterpret_cast<uint32_t*>(&b3);
                const size_t vox = (((size_t)it.b * kChunks + chunk) * p.D + d) * plane_vox + (size_t)h * p.W + w;
                *reinterpret_cast<uint4*>(p.out + vox * 8) = pk;
                if (p.out_lo) {
                  // residual of the bf16 rounding: the next layer's x_lo
                  const float2 f0 = __bfloat1622float2(b0), f1 = __bfloat1622float2(b1), f2 = __bfloat1622float2(b2),
                               f3 = __bfloat1622float2(b3);
                  __nv_bfloat162 l0 = __floats2bfloat162_rn(y[q * 8 + 0] - f0.x, y[q * 8 + 1] - f0.y);
                  __nv_bfloat162 l1 = __floats2bfloat162_rn(y[q * 8 + 2] - f1.x, y[q * 8 + 3] - f1.y);
                  __nv_bfloat162 l2 = __floats2bfloat162_rn(y[q * 8 + 4] - f2.x, y[q * 8 + 5] - f2.y);
                  __nv_bfloat162 l3 = __floats2bfloat162_rn(y[q * 8 + 6] - f3.x, y[q * 8 + 7] - f3.y);
                  uint4 pl;
                  pl.x = *reinterpret_cast<uint32_t*>(&l0);
                  pl.y = *reinterpret_cast<uint32_t*>(&l1);
                  pl.z = *reinterpret_cast<uint32_t*>(&l2);
                  pl.w = *reinterpret_cast<uint32_t*>(&l3);
                  *reinterpret_cast<uint4*>(p.out_lo + vox * 8) = pl;
                }
              }
            }
          }
          if (p.head_out && !p.acc_out && inside)
            p.head_out[((size_t)it.b * p.D + d) * plane_vox + (size_t)h * p.W + w] = head + p.head_b;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(BAR_AEMPTY + buf));
      }
    }
  }
  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

}  // namespace pmconv

extern void pmnet_set_error(const char* msg);

extern "C" int pmnet_conv3d_k3_c96_pass(const void* x_c8, const void* w_packed, const float* scale, const float* bias,
                                        void* y_c8, void* y_lo_c8, const float* acc_in, float* acc_out,
                                        const float* head_w, float head_b, float* head_out, int32_t B, int32_t D,
                                        int32_t H, int32_t W, int32_t relu, int32_t planes_per_item, int32_t max_ctas,
                                        void* stream_);

extern "C" int pmnet_conv3d_k3_c96(const void* x_c8, const void* w_packed, const float* scale, const float* bias,
                                   void* y_c8, const float* head_w, float head_b, float* head_out, int32_t B,
                                   int32_t D, int32_t H, int32_t W, int32_t relu, int32_t planes_per_item,
                                   int32_t max_ctas, void* stream_) {
  return pmnet_conv3d_k3_c96_pass(x_c8, w_packed, scale, bias, y_c8, nullptr, nullptr, nullptr, head_w, head_b, head_out,
                                  B, D, H, W, relu, planes_per_item, max_ctas, stream_);
}

extern "C" int pmnet_conv3d_k3_c96_pass(const void* x_c8, const void* w_packed, const float* scale, const float* bias,
                                        void* y_c8, void* y_lo_c8, const float* acc_in, float* acc_out,
                                        const float* head_w, float head_b, float* head_out, int32_t B, int32_t D,
                                        int32_t H, int32_t W, int32_t relu, int32_t planes_per_item, int32_t max_ctas,
                                        void* stream_) {
  using namespace pmconv;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!x_c8 || !w_packed || !scale || !bias || (!y_c8 && !head_out && !acc_out) || (head_out && !head_w) ||
      (y_lo_c8 && !y_c8)) {
    pmnet_set_error("pmnet_conv3d_k3_c96: null argument");
    return PMNET_EINVAL;
  }
  if (B <= 0 || D <= 0 || H <= 0 || W <= 0 || (D & 1)) {
    pmnet_set_error("pmnet_conv3d_k3_c96: sizes must be positive and D even");
    return PMNET_EINVAL;
  }
  EncodeTiledFn enc = get_encode();
  if (!enc) {
    pmnet_set_error("pmnet_conv3d_k3_c96: cuTensorMapEncodeTiled unavailable");
    return PMNET_ECUDA;
  }
  CUtensorMap tmap;
  const cuuint64_t gdim[5] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)kChunks, (cuuint64_t)B};
  const cuuint64_t gstr[4] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16, (cuuint64_t)D * H * W * 16,
                              (cuuint64_t)kChunks * D * H * W * 16};
  const cuuint32_t box[5] = {kHaloW * 8, kHaloH, 1, kChunks, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x_c8), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) {
    pmnet_set_error("pmnet_conv3d_k3_c96: cuTensorMapEncodeTiled failed");
    return PMNET_ECUDA;
  }
  Params p;
  p.B = B; p.D = D; p.H = H; p.W = W;
  int dc = planes_per_item > 0 ? planes_per_item : 16;
  dc = (dc + 1) & ~1;
  if (dc > D) dc = D;
  p.dc = dc;
  p.n_hb = (H + kTileH - 1) / kTileH;
  p.n_wb = (W + kTileW - 1) / kTileW;
  p.n_dc = (D + dc - 1) / dc;
  p.n_items = B * p.n_dc * p.n_hb * p.n_wb;
  p.wpacked = (const __nv_bfloat16*)w_packed;
  p.scale = scale; p.bias = bias;
  p.out = (__nv_bfloat16*)y_c8;
  p.relu = relu;
  p.head_w = head_w; p.head_b = head_b; p.head_out = head_out;
  p.acc_in = acc_in; p.acc_out = acc_out; p.out_lo = (__nv_bfloat16*)y_lo_c8;
  int sms = 148;
  {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  int grid = p.n_items < sms ? p.n_items : sms;
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  cudaError_t e = cudaFuncSetAttribute(conv3d_k3_c96_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  conv3d_k3_c96_kernel<<<grid, kThreads, kSmemBytes, stream>>>(tmap, p);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}
