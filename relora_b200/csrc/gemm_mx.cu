// Block-scaled MXFP8 path for the frozen weights: tcgen05.mma.kind::mxf8f6f4.block_scale with the ue8m0 scale
// factors staged shared memory -> tensor memory by tcgen05.cp, plus the quantisation / requantisation kernels.
//
// Replaces the reference's bitsandbytes storage (peft_pretraining/relora.py:222-238 Params4bit / Int8Params, :314-317 matmul_4bit /
// bnb.matmul, :277-299 dequantise -> add -> requantise merge).  Formats (OCP MX): E4M3 elements, one UE8M0 (power of two) scale per
// 32 elements of the reduction dimension.
//
//   * activations / gradients  [M, K]: scales per (row, 32 columns)                              mx_quantize_rows
//   * frozen weights           [N, K]: ONE scale per 32 x 32 tile, so the same bytes serve        mx_quantize_weight_2d
//       - the forward GEMM  y = x · Wᵀ   (B operand K-major,  scale of (n, k/32) = tile scale)
//       - the backward GEMM dx = dy · W  (B operand MN-major, scale of (k, n/32) = the same tile scale)
//     -> 1 byte per parameter + two expanded scale arrays (1/32 byte each) instead of two fp8 copies.
//
// Scale-factor layout in global memory (the layout tcgen05.cp / the MMA expect, cf. CUTLASS Sm1xxBlockScaledBasicChunk): for a
// block of 128 rows x 4 scale columns (= 128 reduction elements) 512 contiguous bytes,
//     byte[(r % 32) * 16 + (r / 32) * 4 + j]  =  scale(row r, scale column j)
// blocks ordered [row block][k group].  One such block is one TMEM "column quad" after  tcgen05.cp.32x128b.warpx4 .
#include <cuda.h>
#include <cuda_fp8.h>

#include <cstdlib>
#include <stdexcept>

#include "common.cuh"
#include "kernels.h"
#include "sm100.cuh"
#include "tensormap.h"

namespace rb {

using namespace sm100;

namespace {

constexpr int BM = 128, BN = 128;
constexpr int KB8 = 128;  // fp8 elements per k-block (128 bytes = one swizzle row)
constexpr int KB16 = 64;  // bf16 elements per k-block of the optional second (LoRA) segment
constexpr int kStages = 4;
constexpr int kTileBytes = BM * 128;           // 16 KB per operand tile
constexpr int kStageBytes = 2 * kTileBytes;    // A + B
constexpr int kSfBytes = 512;                  // one 128-row x 4-scale block
constexpr int kSlabBytes = BM * 128;           // 128 x 64 bf16 output slab
constexpr int kSlabs = 2;
constexpr int kSmemTiles = kStages * kStageBytes;             // 128 KB
constexpr int kSmemSf = kStages * 2 * kSfBytes;               // 4 KB
constexpr int kSmemTotal = kSmemTiles + kSmemSf + 1024 + kSlabs * kSlabBytes + 1024;
constexpr uint32_t kTmemCols = 512;  // 2 x 128 accumulator columns + 2 x 8 scale columns -> next power of two
constexpr uint32_t kSfCol0 = 256;    // scale factors live behind the two accumulator stages

// ---- instruction descriptor of kind::mxf8f6f4.block_scale (cute::UMMA::InstrDescriptorBlockScaled):
//   [4,6) b_sf_id  [7,10) a_format (0 = E4M3)  [10,13) b_format  [15] a_major  [16] b_major  [17,23) N>>3  [23] scale_format (1 = UE8M0)
//   [24,29) M>>4  [29,31) a_sf_id
__device__ __forceinline__ uint32_t make_idesc_mx(int M, int N, int b_mn_major, uint32_t a_sf_id, uint32_t b_sf_id) {
  return (b_sf_id << 4) | (uint32_t(b_mn_major) << 16) | (uint32_t(N >> 3) << 17) | (1u << 23) | (uint32_t(M >> 4) << 24) | (a_sf_id << 29);
}
__device__ __forceinline__ void umma_mx_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t tmem_sfa, uint32_t tmem_sfb,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
      : "memory");
}
// shared memory -> tensor memory: 32 rows x 128 bits, replicated into the four 32-lane quadrants (the scale-factor copy)
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t tmem_dst, uint64_t smem_desc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}
// K-major, no swizzle: 8-row groups of 16-byte rows are 128 bytes apart (SBO); one 16-byte column -> LBO unused
__device__ __forceinline__ uint64_t make_desc_sf(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= uint64_t((smem_addr >> 4) & 0x3FFF);
  d |= uint64_t(1) << 16;            // LBO (unused: a single 16-byte column)
  d |= uint64_t(128 >> 4) << 32;     // SBO
  d |= uint64_t(1) << 46;            // descriptor version (sm_100)
  return d;                          // layout type 0 = no swizzle ("interleave")
}
// fp8 operand descriptors (128-byte swizzle).  K-major: rows of 128 B = 128 elements, 32 elements (32 B) per MMA.
// MN-major: rows of 128 B = 128 MN elements, 8-row (K) groups 1024 B apart, 32 K rows (4096 B) per MMA.
__device__ __forceinline__ uint64_t desc8_k(uint32_t smem_addr, int kstep) { return make_desc_sw128(smem_addr + kstep * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc8_mn(uint32_t smem_addr, int kstep) { return make_desc_sw128(smem_addr + kstep * 4096, 16384, 1024); }
__device__ __forceinline__ uint64_t desc16_k(uint32_t smem_addr, int kstep) { return make_desc_sw128(smem_addr + kstep * 32, 16, 1024); }

__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(reinterpret_cast<uint64_t>(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct MxArgs {
  int M, N, K8, K2;                 // K8: fp8 reduction length (multiple of 128), K2: bf16 LoRA segment (multiple of 64, may be 0)
  const uint8_t* sfa;               // [m blocks][sfa_kg][512]
  const uint8_t* sfb;               // [n blocks][sfb_kg][512]
  int sfa_kg, sfb_kg;               // 128-element k groups per row block in the arrays
  int num_m_tiles, num_n_tiles;
  int has_res;
  float alpha2;                     // unused (the LoRA operand already carries its scale)
};

template <bool B_MN>
__global__ void __launch_bounds__(256, 1)
gemm_mx_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_a2,
               const __grid_constant__ CUtensorMap map_b2, const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
               const MxArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sf_base = smem + kSmemTiles;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sf_base + kSmemSf);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* res_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + kSlabs);
  uint8_t* slab_base = sf_base + kSmemSf + 1024;

  const uint32_t warp = warp_id(), lane = lane_id();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    tma_prefetch_desc(&map_out);
    if (p.K2 > 0) {
      tma_prefetch_desc(&map_a2);
      tma_prefetch_desc(&map_b2);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    for (int a = 0; a < kSlabs; ++a) mbar_init(&res_bar[a], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);  // known warp-uniform

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kb8 = p.K8 / KB8, kb16 = p.K2 / KB16;

  if (warp == 0) {
    // ===================================================================== TMA producer: the whole warp walks the loop, one elected lane
    // issues (uniform control flow lets ptxas keep addresses / descriptors in uniform registers instead of wrapping every
    // UTMALDG / UTC*MMA in an ELECT + R2UR + BRA.U.ANY loop; see gemm_tcgen05.cu)
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int tm = tile / p.num_n_tiles, tn = tile % p.num_n_tiles;
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < kb8 + kb16; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * kStageBytes;
          uint8_t* sb = sa + kTileBytes;
          if (!elect_one()) {
          } else if (kb < kb8) {
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes + 2 * kSfBytes);
            tma_load_2d(&map_a, &full_bar[stage], sa, kb * KB8, m0, kEvictNormal);
            if constexpr (B_MN) tma_load_2d(&map_b, &full_bar[stage], sb, n0, kb * KB8, kEvictLast);  // box {128 (MN), 128 (K rows)}
            else tma_load_2d(&map_b, &full_bar[stage], sb, kb * KB8, n0, kEvictLast);
            bulk_copy_g2s(sf_base + stage * 2 * kSfBytes, p.sfa + ((long long)tm * p.sfa_kg + kb) * kSfBytes, kSfBytes, &full_bar[stage]);
            bulk_copy_g2s(sf_base + stage * 2 * kSfBytes + kSfBytes, p.sfb + ((long long)tn * p.sfb_kg + kb) * kSfBytes, kSfBytes,
                          &full_bar[stage]);
          } else {  // bf16 LoRA segment: [x | u]·[W | B]ᵀ shares the accumulator
            const int k = (kb - kb8) * KB16;
            mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
            tma_load_2d(&map_a2, &full_bar[stage], sa, k, m0, kEvictNormal);
            tma_load_2d(&map_b2, &full_bar[stage], sb, k, n0, kEvictLast);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t idesc16 = make_idesc_bf16(BM, BN, 0, 0);
      int stage = 0, acc = 0, sfbuf = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kb8 + kb16; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem) + stage * kStageBytes;
          const uint32_t sb = sa + kTileBytes;
          if (!elect_one()) {
          } else if (kb < kb8) {
            // scale factors of this k-block: shared memory -> 4 + 4 tensor-memory columns (tcgen05.cp and tcgen05.mma execute in
            // issue order, so the copies are complete before the MMAs below read them; two buffers alternate anyway)
            const uint32_t sfa_t = tmem_base + kSfCol0 + sfbuf * 8, sfb_t = sfa_t + 4;
            const uint32_t sf_s = smem_u32(sf_base + stage * 2 * kSfBytes);
            tmem_cp_32x128b_warpx4(sfa_t, make_desc_sf(sf_s));
            tmem_cp_32x128b_warpx4(sfb_t, make_desc_sf(sf_s + kSfBytes));
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {  // K = 32 per instruction; scale column k of the block (sf id, also in bits 30-31 of the address)
              const uint64_t da = desc8_k(sa, k);
              const uint64_t db = B_MN ? desc8_mn(sb, k) : desc8_k(sb, k);
              umma_mx_ss(d_tmem, da, db, make_idesc_mx(BM, BN, B_MN ? 1 : 0, k, k), sfa_t | (k << 30), sfb_t | (k << 30), (kb | (int)k) != 0);
            }
            umma_commit(&empty_bar[stage]);
          } else {
#pragma unroll
            for (int k = 0; k < KB16 / 16; ++k) umma_f16_ss(d_tmem, desc16_k(sa, k), desc16_k(sb, k), idesc16, 1u);
            umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
          if (kb < kb8) sfbuf ^= 1;
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) umma_commit(&tmem_full[acc]);
        __syncwarp();
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================================================================== epilogue: TMEM -> bf16 slab (+ residual by TMA) -> TMA store
    const uint32_t quad = warp & 3;
    const uint32_t rloc = quad * 32 + lane;
    const bool issuer = (threadIdx.x == 128);
    int acc = 0, slab_counter = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * BM, n0 = (tile % p.num_n_tiles) * BN;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int sl = 0; sl < BN / 64; ++sl) {
        const int sb_i = slab_counter % kSlabs;
        uint8_t* slab = slab_base + sb_i * kSlabBytes;
        uint8_t* rowp = slab + rloc * 128;
        if (p.has_res) {
          if (issuer) {
            mbar_arrive_expect_tx(&res_bar[sb_i], kSlabBytes);
            tma_load_2d(&map_res, &res_bar[sb_i], slab, n0 + sl * 64, m0, kEvictNormal);
          }
          mbar_wait(&res_bar[sb_i], (slab_counter / kSlabs) & 1);
        }
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, acc * BN + sl * 64), r0);
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, acc * BN + sl * 64 + 32), r1);
        tmem_ld_wait();
        if (sl == BN / 64 - 1) {
          tc_fence_before();
          mbar_arrive(&tmem_empty[acc]);
        }
        uint4 packed[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float((q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i]);
          if (p.has_res) {
            float a[8];
            unpack8(*reinterpret_cast<const uint4*>(rowp + ((q ^ (rloc & 7)) << 4)), a);
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] += a[i];
          }
          packed[q] = pack8(f);
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(rowp + ((q ^ (rloc & 7)) << 4)) = packed[q];
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (issuer) {
          if (n0 + sl * 64 < p.N) {
            tma_store_2d(&map_out, slab, n0 + sl * 64, m0);
            tma_store_commit();
          }
          tma_store_wait_read<kSlabs - 1>();  // the other slab (the next one to be written) has been read by its store
        }
        named_bar_sync(1, 128);
        ++slab_counter;
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------- quantisation
__device__ __forceinline__ uint32_t sf_offset(long long row, int sfcol, int kg_per_block) {
  // byte offset of scale (row, scale column) in the [row block][k group][512] layout
  const long long rb = row >> 7;
  const int r = int(row & 127), kg = sfcol >> 2, j = sfcol & 3;
  return uint32_t(((rb * kg_per_block + kg) << 9) + ((r & 31) << 4) + ((r >> 5) << 2) + j);
}
// smallest power of two s = 2^e with amax / s <= 448 (E4M3 max); returns the biased UE8M0 byte and 1/s
__device__ __forceinline__ uint8_t ue8m0_for(float amax, float& inv_scale) {
  int e = -127;
  if (amax > 0.f) {
    e = (int)ceilf(log2f(amax * (1.f / 448.f)));
    e = max(-127, min(127, e));
  }
  inv_scale = exp2f((float)-e);
  return (uint8_t)(e + 127);
}

// rows: one warp per (row, 4 blocks of 32 columns): lane l holds 4 consecutive elements of column block l / 8
__global__ void __launch_bounds__(256) mx_quantize_rows_kernel(const bf16* __restrict__ x, long long ldx, uint8_t* __restrict__ q, long long ldq,
                                                               uint8_t* __restrict__ sf, int M, int K, int Kpad, int kg_per_block) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int groups_per_row = Kpad / 128;
  const long long total = (long long)((M + 127) / 128 * 128) * groups_per_row;
  for (long long w = warp_global; w < total; w += ((long long)gridDim.x * blockDim.x) >> 5) {
    const long long row = w / groups_per_row;
    const int kg = int(w % groups_per_row);
    const int col = kg * 128 + lane * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < M && col < K) {  // K is a multiple of 8: a 4-element group is entirely inside or outside
      const uint2 raw = *reinterpret_cast<const uint2*>(x + row * ldx + col);
      const float2 a = unpack_bf16x2(raw.x), b = unpack_bf16x2(raw.y);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
    float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));  // 8 lanes = one 32-column block
    float inv;
    const uint8_t e = ue8m0_for(amax, inv);
    if ((lane & 7) == 0) sf[sf_offset(row, kg * 4 + (lane >> 3), kg_per_block)] = e;
    if (row < M && col < Kpad) {
      const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[0] * inv, v[1] * inv), __NV_SATFINITE, __NV_E4M3);
      const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[2] * inv, v[3] * inv), __NV_SATFINITE, __NV_E4M3);
      *reinterpret_cast<uint32_t*>(q + row * ldq + col) = (uint32_t)lo | ((uint32_t)hi << 16);
    }
  }
}

// weights: one warp per 32 x 32 tile (lane = row of the tile, 32 columns each); optional fp32 `delta` is added first (merge)
__global__ void __launch_bounds__(256) mx_quantize_weight_2d_kernel(const bf16* __restrict__ w, long long ldw, const uint8_t* __restrict__ q_old,
                                                                    const uint8_t* __restrict__ sf_old, const float* __restrict__ delta,
                                                                    long long ldd, uint8_t* __restrict__ q, long long ldq,
                                                                    uint8_t* __restrict__ sf_fwd, uint8_t* __restrict__ sf_bwd, int N, int K,
                                                                    int Npad, int Kpad, int kg_fwd, int kg_bwd) {
  const int lane = threadIdx.x & 31;
  const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int tiles_k = Kpad / 32;
  const long long total = (long long)(Npad / 32) * tiles_k;
  for (long long t = warp_global; t < total; t += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int tn = int(t / tiles_k), tk = int(t % tiles_k);
    const long long n = (long long)tn * 32 + lane;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    if (n < N) {
      if (w != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          const int k = tk * 32 + j;
          if (k < K) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4*>(w + n * ldw + k), f);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j + i] = f[i];
          }
        }
      } else {  // requantisation: start from the packed weight and its (forward-layout) scale
        const float s_old = exp2f((float)sf_old[sf_offset(n, tk, kg_fwd)] - 127.f);
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int k = tk * 32 + j;
          if (k < Kpad) {
            const uint32_t raw = *reinterpret_cast<const uint32_t*>(q_old + n * ldq + k);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)((raw >> (8 * i)) & 0xff), __NV_E4M3);
              v[j + i] = __half2float(*reinterpret_cast<const __half*>(&h)) * s_old;
            }
          }
        }
      }
      if (delta != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int k = tk * 32 + j;
          if (k < K) v[j] += delta[n * ldd + k];
        }
      }
    }
    float amax = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(v[j]));
    amax = warp_max(amax);
    float inv;
    const uint8_t e = ue8m0_for(amax, inv);
    // forward layout: scale of (row n, k block tk); backward layout: scale of (row k, n block tn) for the 32 k of this tile
    if (n < Npad) sf_fwd[sf_offset(n, tk, kg_fwd)] = e;
    {
      const long long k = (long long)tk * 32 + lane;
      if (k < Kpad) sf_bwd[sf_offset(k, tn, kg_bwd)] = e;
    }
    if (n < Npad) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const int k = tk * 32 + j;
        if (k < Kpad) {
          const __nv_fp8x2_storage_t lo = __nv_cvt_float2_to_fp8x2(make_float2(v[j] * inv, v[j + 1] * inv), __NV_SATFINITE, __NV_E4M3);
          const __nv_fp8x2_storage_t hi = __nv_cvt_float2_to_fp8x2(make_float2(v[j + 2] * inv, v[j + 3] * inv), __NV_SATFINITE, __NV_E4M3);
          *reinterpret_cast<uint32_t*>(q + n * ldq + k) = (uint32_t)lo | ((uint32_t)hi << 16);
        }
      }
    }
  }
}

// dense bf16 copy of a packed weight (checkpoints, numerics oracle)
__global__ void __launch_bounds__(256) mx_dequantize_weight_kernel(const uint8_t* __restrict__ q, long long ldq, const uint8_t* __restrict__ sf,
                                                                   bf16* __restrict__ out, long long ldo, int N, int K, int kg) {
  const long long total = (long long)N * (K / 4);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long n = i / (K / 4);
    const int k = int(i % (K / 4)) * 4;
    const float s = exp2f((float)sf[sf_offset(n, k >> 5, kg)] - 127.f);
    const uint32_t raw = *reinterpret_cast<const uint32_t*>(q + n * ldq + k);
    float f[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __half_raw h = __nv_cvt_fp8_to_halfraw((__nv_fp8_storage_t)((raw >> (8 * j)) & 0xff), __NV_E4M3);
      f[j] = __half2float(*reinterpret_cast<const __half*>(&h)) * s;
    }
    uint2 o;
    o.x = pack_bf16x2(f[0], f[1]);
    o.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(out + n * ldo + k) = o;
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------- host
long long mx_sf_bytes(long long rows, long long k_elems) { return ((rows + 127) / 128) * ((k_elems + 127) / 128) * 512; }

void mx_quantize_rows(const void* x, long long ldx, void* q, long long ldq, void* sf, int M, int K, cudaStream_t s) {
  if (K % 8) throw std::runtime_error("mx_quantize_rows: K must be a multiple of 8");
  const int Kpad = (K + 127) / 128 * 128;
  if (ldq < Kpad) throw std::runtime_error("mx_quantize_rows: the fp8 buffer needs a row pitch >= K rounded up to 128");
  const long long warps = (long long)((M + 127) / 128 * 128) * (Kpad / 128);
  const int grid = (int)std::min<long long>((warps * 32 + 255) / 256, (long long)num_sms() * 16);
  if (grid <= 0) return;
  launch_k(mx_quantize_rows_kernel, grid, 256, 0, s, (const bf16*)x, ldx, (uint8_t*)q, ldq, (uint8_t*)sf, M, K, Kpad, Kpad / 128);
  RB_CHECK_LAUNCH("mx_quantize_rows");
}

void mx_quantize_weight_2d(const void* w, long long ldw, const void* q_old, const void* sf_old, const float* delta, long long ldd, void* q,
                           long long ldq, void* sf_fwd, void* sf_bwd, int N, int K, cudaStream_t s) {
  if (K % 8) throw std::runtime_error("mx_quantize_weight_2d: K must be a multiple of 8");
  const int Npad = (N + 127) / 128 * 128, Kpad = (K + 127) / 128 * 128;
  if (ldq < Kpad) throw std::runtime_error("mx_quantize_weight_2d: the fp8 buffer needs a row pitch >= K rounded up to 128");
  const long long warps = (long long)(Npad / 32) * (Kpad / 32);
  const int grid = (int)std::min<long long>((warps * 32 + 255) / 256, (long long)num_sms() * 16);
  launch_k(mx_quantize_weight_2d_kernel, grid, 256, 0, s, (const bf16*)w, ldw, (const uint8_t*)q_old, (const uint8_t*)sf_old, delta, ldd,
           (uint8_t*)q, ldq, (uint8_t*)sf_fwd, (uint8_t*)sf_bwd, N, K, Npad, Kpad, Kpad / 128, Npad / 128);
  RB_CHECK_LAUNCH("mx_quantize_weight_2d");
}

void mx_dequantize_weight(const void* q, long long ldq, const void* sf_fwd, void* out, long long ldo, int N, int K, cudaStream_t s) {
  if (K % 4) throw std::runtime_error("mx_dequantize_weight: K must be a multiple of 4");
  const int Kpad = (K + 127) / 128 * 128;
  const long long total = (long long)N * (K / 4);
  const int grid = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
  if (grid <= 0) return;
  launch_k(mx_dequantize_weight_kernel, grid, 256, 0, s, (const uint8_t*)q, ldq, (const uint8_t*)sf_fwd, (bf16*)out, ldo, N, K, Kpad / 128);
  RB_CHECK_LAUNCH("mx_dequantize_weight");
}

void gemm_mx(const MxGemmDesc& d, cudaStream_t stream) {
  if (d.M <= 0 || d.N <= 0) return;
  const int K8 = (d.K + 127) / 128 * 128;
  if (d.K2 % KB16) throw std::runtime_error("gemm_mx: the bf16 segment must be a multiple of 64");
  if (d.ldc % 8 || (reinterpret_cast<uintptr_t>(d.out) & 15)) throw std::runtime_error("gemm_mx: output must be 16-byte aligned");
  MxArgs p;
  p.M = d.M; p.N = d.N; p.K8 = K8; p.K2 = d.K2;
  p.sfa = reinterpret_cast<const uint8_t*>(d.sfa); p.sfb = reinterpret_cast<const uint8_t*>(d.sfb);
  p.sfa_kg = K8 / 128; p.sfb_kg = K8 / 128;
  p.num_m_tiles = ceil_div(d.M, BM); p.num_n_tiles = ceil_div(d.N, BN);
  p.has_res = d.residual != nullptr ? 1 : 0;
  p.alpha2 = 1.f;
  // operands: fp8 bytes; A [M, K8] K-major (pitch lda); B K-major [N, K8] (pitch ldb) or MN-major [K8 rows, N] (pitch ldb)
  CUtensorMap ma = make_map_2d_sw128(d.a, K8, d.M, d.lda, KB8, BM, 1);
  CUtensorMap mb = d.b_mn_major ? make_map_2d_sw128(d.b, d.N, K8, d.ldb, 128, KB8, 1) : make_map_2d_sw128(d.b, K8, d.N, d.ldb, KB8, BN, 1);
  CUtensorMap mo = make_map_2d_sw128(d.out, d.N, d.M, d.ldc, 64, BM, 2);
  CUtensorMap ma2 = ma, mb2 = mb, mr = mo;
  if (d.K2 > 0) {
    ma2 = make_map_2d_sw128(d.a2, d.K2, d.M, d.lda2, KB16, BM, 2);
    mb2 = make_map_2d_sw128(d.b2, d.K2, d.N, d.ldb2, KB16, BN, 2);
  }
  if (d.residual != nullptr) {
    if (d.ldr % 8 || (reinterpret_cast<uintptr_t>(d.residual) & 15)) throw std::runtime_error("gemm_mx: residual must be 16-byte aligned");
    mr = make_map_2d_sw128(d.residual, d.N, d.M, d.ldr, 64, BM, 2);
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (d.b_mn_major) {
    static bool cfg = false;
    if (!cfg) { check(cudaFuncSetAttribute(gemm_mx_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal), "attr(gemm_mx)"); cfg = true; }
    launch_k(gemm_mx_kernel<true>, grid, 256, kSmemTotal, stream, ma, mb, ma2, mb2, mo, mr, p);
  } else {
    static bool cfg = false;
    if (!cfg) { check(cudaFuncSetAttribute(gemm_mx_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal), "attr(gemm_mx)"); cfg = true; }
    launch_k(gemm_mx_kernel<false>, grid, 256, kSmemTotal, stream, ma, mb, ma2, mb2, mo, mr, p);
  }
  RB_CHECK_LAUNCH("gemm_mx");
}

}  // namespace rb
