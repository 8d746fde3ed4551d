// gemm.cu - C[M, N] = act(A[M, K] . W[N, K]^T + bias) on the sm_100a tensor cores (tcgen05 + TMEM + TMA), the linear
// layers of the reference network that are not 3x3x3 convolutions:
//   WindowAttention.qkv / .proj, Mlp.fc1 / .fc2          src/pmnet/network/backbones/swinv2.py:114-158, swin.py:19-44
//   PatchMerging.reduction, PatchEmbed.proj (as a GEMM)    swinv2.py:346-363, 484-500
//   FPNDecoder top-level conv (im2col) and small laterals  decoders/fpn_decoder.py:86-115
//   TokenHead / MaskHead MLPs                               token_head.py:50-86, mask_head.py:128-168
//
// Operands are bf16, K-major (row-major with K contiguous), fetched by TMA as 64-element (128 B) wide boxes with the
// 128-byte swizzle; accumulation is fp32 in TMEM. One CTA computes one 128 x BN output tile (BN = 96 or 192: every
// layer width of the network is a multiple of 96) through a 4-stage TMA -> tcgen05.mma pipeline:
//   warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warps 2.. = epilogue (12 or 24 of them) (TMEM -> registers -> bias / GELU ->
//   fp32 and / or bf16 stores; TMEM lane quadrant = warp % 4, column group = (warp - 2) / 4).
// Split precision: with a_lo / w_lo given the K loop runs three times over (a_hi, w_hi), (a_lo, w_hi), (a_hi, w_lo) into
// the SAME accumulator - x w = x_hi w_hi + x_lo w_hi + x_hi w_lo to 2^-16 relative - and the epilogue can emit the
// (hi, lo) pair of its result for the next layer. Out-of-range rows / columns / K tails are zero-filled by TMA.

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/pmnet_b200.h"

extern void pmnet_set_error(const char* msg);

namespace pmgemm {

constexpr int BM = 128, BK = 64;
// epilogue warps: one per (TMEM lane quadrant, 32-column chunk of the tile) = 12 for BN = 96, 24 for BN = 192
__host__ __device__ constexpr int epi_warps(int bn) { return 4 * (bn / 32); }
__host__ __device__ constexpr int threads(int bn) { return 64 + 32 * epi_warps(bn); }  // + TMA producer warp + MMA issuer warp
constexpr int kABytes = BM * BK * 2;  // 16384

struct Params {
  int M, N, K;
  int npass;
  int pa[3], pb[3];        // operand index (0 = hi, 1 = lo) of A and W in each pass
  const float* bias;       // [N] or null
  float* out_f32;          // [M][N] or null
  __nv_bfloat16* out_hi;   // [M][N] or null: bf16(y)
  __nv_bfloat16* out_lo;   // [M][N] or null: bf16(y - bf16(y))
  int act;                 // 0 none, 1 GELU (erf), 2 ReLU, 3 SiLU
};

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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
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

// K-major shared-memory matrix descriptor for a tile of 128-byte rows written by TMA with CU_TENSOR_MAP_SWIZZLE_128B
// (cute::UMMA::SmemDescriptor): start >> 4, leading byte offset (unused for swizzled K-major layouts, 1), stride byte
// offset = 8 rows x 128 B = 1024, version 1, layout type 2 = SWIZZLE_128B. A K step of 16 elements moves the start by 32 B.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}

// erf to 1.5e-7 absolute (Abramowitz & Stegun 7.1.26): one reciprocal, one exponential and five FMAs instead of the
// branchy library erff - the epilogue of the fc1 GEMMs is bound by instruction issue, not by the tensor pipe
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __frcp_rn(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float r = 1.0f - poly * t * __expf(-ax * ax);
  return copysignf(r, x);
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752440f));  // nn.GELU (swin.py:32)
  if (act == 2) return fmaxf(x, 0.0f);
  if (act == 3) return x / (1.0f + __expf(-x));
  return x;
}

// Persistent: the grid is one CTA per SM (two for the short-K shape, whose two pipeline stages leave room) and every CTA
// walks output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (n tile fastest, so that the CTAs working at the same
// time share their A rows in L2). Barrier init, TMEM allocation and the tensor-map fetch are paid once per CTA, and the
// accumulator is double buffered in TMEM: the epilogue of tile i overlaps the loads and MMAs of tile i + 1.
// kStages = 4 for long K loops; 2 for the short ones (K = 96 ... 256 with up to three passes).
template <int BN, int kStages>
__global__ void __launch_bounds__(threads(BN)) gemm_kernel(const __grid_constant__ CUtensorMap tmA0,
                                                        const __grid_constant__ CUtensorMap tmA1,
                                                        const __grid_constant__ CUtensorMap tmB0,
                                                        const __grid_constant__ CUtensorMap tmB1, const Params p) {
  constexpr int kBBytes = BN * BK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  constexpr uint32_t kTmemCols = 2 * BN <= 256 ? 256 : 512;  // two accumulators of BN columns
  // instruction descriptor, kind::f16: D = f32 (bit 4), A = B = bf16 (bits 7, 10), K-major, N >> 3 at 17, M >> 4 at 24
  constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // the swizzled tiles need 1024-byte alignment in the shared window: align explicitly (1 KB of slack is allocated)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + kStages * kStageBytes;  // full[kStages], empty[kStages], accfull[2], accempty[2]
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (kStages + s); };
  auto bar_accfull = [&](int b) { return bars + 8u * (2 * kStages + b); };
  auto bar_accempty = [&](int b) { return bars + 8u * (2 * kStages + 2 + b); };
  volatile uint32_t* tmem_ptr_smem = (volatile uint32_t*)(smem + kStages * kStageBytes + 8 * (2 * kStages + 4));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kblocks = (p.K + BK - 1) / BK;
  const int per_tile = kblocks * p.npass;
  const int n_tiles_n = (p.N + BN - 1) / BN;
  const int n_tiles = ((p.M + BM - 1) / BM) * n_tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(bar_full(s), 1);
      mbar_init(bar_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_accfull(b), 1);
      mbar_init(bar_accempty(b), epi_warps(BN));  // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((void*)tmem_ptr_smem)),
                 "r"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;  // k iterations issued so far, across tiles
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
        for (int j = 0; j < per_tile; ++j, ++it) {
          const int pass = j / kblocks, kb = j - pass * kblocks;
          const uint32_t s = it % kStages, round = it / kStages;
          mbar_wait(bar_empty(s), (round & 1) ^ 1);
          mbar_expect_tx(bar_full(s), kStageBytes);
          const CUtensorMap* ma = p.pa[pass] ? &tmA1 : &tmA0;
          const CUtensorMap* mb = p.pb[pass] ? &tmB1 : &tmB0;
          tma_load_2d(sbase + s * kStageBytes, ma, bar_full(s), kb * BK, m0);
          tma_load_2d(sbase + s * kStageBytes + kABytes, mb, bar_full(s), kb * BK, n0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(bar_accempty(buf), ((tcount >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * BN;
        for (int j = 0; j < per_tile; ++j, ++it) {
          const uint32_t s = it % kStages, round = it / kStages;
          mbar_wait(bar_full(s), round & 1);
          tc_fence_after();
          const uint64_t ad = make_desc_sw128(sbase + s * kStageBytes);
          const uint64_t bd = make_desc_sw128(sbase + s * kStageBytes + kABytes);
#pragma unroll
          for (int ks = 0; ks < BK / 16; ++ks)
            tc_mma(tacc, ad + (uint64_t)(ks * 2), bd + (uint64_t)(ks * 2), kIdesc, (j | ks) != 0);
          tc_commit(bar_empty(s));  // the stage is free again when these MMAs have read it
        }
        tc_commit(bar_accfull(buf));
      }
    }
  } else {
    // ---- epilogue: warp w reads TMEM lanes [32 (w % 4), 32 (w % 4) + 32) = rows of the tile and ONE 32-column chunk:
    // a single warp per quadrant could not keep up with the tensor pipe (bias + GELU + bf16 split of 128 x 192 values is
    // ~10^4 dependent instructions per thread; measured 968 -> 437 us for the stage-0 fc1 with 4 -> 12 warps)
    const int quad = warp & 3;
    const int cgroup = (warp - 2) >> 2;
    uint32_t tcount = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int m0 = (tile / n_tiles_n) * BM, n0 = (tile % n_tiles_n) * BN;
      const uint32_t buf = tcount & 1;
      const int row = m0 + quad * 32 + lane;
      mbar_wait(bar_accfull(buf), (tcount >> 1) & 1);
      tc_fence_after();
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int c32 = cgroup; c32 < BN / 32; c32 += epi_warps(BN) / 4) {
        uint32_t v[32];
        tmem_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * BN + c32 * 32, v);
        const int n = n0 + c32 * 32;
        if (!row_ok || n >= p.N) continue;
        float y[32];
        if (p.bias) {
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + n);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(b4 + j);
            y[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
            y[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
            y[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
            y[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = __uint_as_float(v[j]);
        }
        if (p.act) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = apply_act(y[j], p.act);
        }
        const size_t o = (size_t)row * p.N + n;
        if (p.out_f32) {
          float4* q = reinterpret_cast<float4*>(p.out_f32 + o);
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] = make_float4(y[4 * j], y[4 * j + 1], y[4 * j + 2], y[4 * j + 3]);
        }
        if (p.out_hi) {
          uint32_t hi[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * j], y[2 * j + 1]);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          uint4* qh = reinterpret_cast<uint4*>(p.out_hi + o);
#pragma unroll
          for (int j = 0; j < 4; ++j) qh[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          if (p.out_lo) {  // the residual of the bf16 rounding: the low operand of the next layer
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hi[j]));
              const __nv_bfloat162 l2 = __floats2bfloat162_rn(y[2 * j] - f.x, y[2 * j + 1] - f.y);
              hi[j] = *reinterpret_cast<const uint32_t*>(&l2);
            }
            uint4* ql = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
            for (int j = 0; j < 4; ++j) ql[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_accempty(buf));
    }
  }
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

// bf16 [rows][K] row-major, box = 64 (K) x box_rows, 128-byte swizzle
static bool make_map(CUtensorMap* map, const void* base, int64_t rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstr[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int kStages>
static cudaError_t launch(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& b0, const CUtensorMap& b1,
                          const Params& p, cudaStream_t stream) {
  constexpr int smem = kStages * (kABytes + BN * BK * 2) + 8 * (2 * kStages + 4) + 16 + 1024;  // + alignment slack
  cudaError_t e = cudaFuncSetAttribute(gemm_kernel<BN, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        sms <= 0)
      sms = 148;
  }
  // CTAs per SM: limited by shared memory and by the 512 TMEM columns (two accumulators per CTA)
  constexpr int tmem_cols = 2 * BN <= 256 ? 256 : 512;
  int per_sm = (227 * 1024) / smem;
  if (per_sm > 512 / tmem_cols) per_sm = 512 / tmem_cols;
  if (per_sm < 1) per_sm = 1;
  const long tiles = (long)((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
  long grid = (long)sms * per_sm;
  if (grid > tiles) grid = tiles;
  gemm_kernel<BN, kStages><<<(unsigned)grid, threads(BN), smem, stream>>>(a0, a1, b0, b1, p);
  return cudaGetLastError();
}

}  // namespace pmgemm

extern "C" int pmnet_gemm_bf16(const void* a_hi, const void* a_lo, const void* w_hi, const void* w_lo, const float* bias,
                               float* out_f32, void* out_hi, void* out_lo, int64_t M, int32_t N, int32_t K, int32_t act,
                               void* stream_) {
  using namespace pmgemm;
  cudaStream_t stream = (cudaStream_t)stream_;
  if (!a_hi || !w_hi || (!out_f32 && !out_hi) || (out_lo && !out_hi) || ((a_lo == nullptr) != (w_lo == nullptr))) {
    pmnet_set_error("pmnet_gemm_bf16: null / inconsistent argument");
    return PMNET_EINVAL;
  }
  if (M <= 0 || N <= 0 || K <= 0 || (K & 7) || (N % 32) || M > 0x7fffffffLL) {
    pmnet_set_error("pmnet_gemm_bf16: K must be a multiple of 8 and N a multiple of 32");
    return PMNET_EINVAL;
  }
  const int BN = (N % 192 == 0) ? 192 : 96;
  if (N % 96 != 0 && N > 96) {
    pmnet_set_error("pmnet_gemm_bf16: N must be <= 96 or a multiple of 96");
    return PMNET_EINVAL;
  }
  CUtensorMap a0, a1, b0, b1;
  bool ok = make_map(&a0, a_hi, M, K, BM) && make_map(&b0, w_hi, N, K, BN);
  if (ok && a_lo) ok = make_map(&a1, a_lo, M, K, BM) && make_map(&b1, w_lo, N, K, BN);
  if (!ok) {
    pmnet_set_error("pmnet_gemm_bf16: cuTensorMapEncodeTiled failed (operands must be 16-byte aligned)");
    return PMNET_ECUDA;
  }
  if (!a_lo) {
    a1 = a0;
    b1 = b0;
  }
  Params p;
  p.M = (int)M; p.N = N; p.K = K;
  p.npass = a_lo ? 3 : 1;
  p.pa[0] = 0; p.pb[0] = 0;
  p.pa[1] = 1; p.pb[1] = 0;
  p.pa[2] = 0; p.pb[2] = 1;
  p.bias = bias;
  p.out_f32 = out_f32;
  p.out_hi = (__nv_bfloat16*)out_hi;
  p.out_lo = (__nv_bfloat16*)out_lo;
  p.act = act;
  const bool short_k = ((K + BK - 1) / BK) * p.npass <= 8;
  cudaError_t e;
  if (BN == 192) e = short_k ? launch<192, 2>(a0, a1, b0, b1, p, stream) : launch<192, 4>(a0, a1, b0, b1, p, stream);
  else e = short_k ? launch<96, 2>(a0, a1, b0, b1, p, stream) : launch<96, 4>(a0, a1, b0, b1, p, stream);
  if (e != cudaSuccess) {
    pmnet_set_error(cudaGetErrorString(e));
    return PMNET_ECUDA;
  }
  return PMNET_OK;
}
