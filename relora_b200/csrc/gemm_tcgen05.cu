// Persistent warp-specialised bf16 GEMM for sm_100a: TMA -> shared memory (128B swizzle) -> tcgen05.mma
// with fp32 accumulators in tensor memory -> tcgen05.ld epilogue.
//
//   D[M,N] = alpha * ( A1[M,K1]·B1[N,K1]ᵀ + A2[M,K2]·B2[N,K2]ᵀ ) (+ residual) (+ D)
//
// The second product is the fused LoRA up-projection: it simply extends the K loop ("concat-K"), so the frozen
// weight GEMM and the low-rank branch share one accumulator, one read of x and one write of y.
// Reference behaviour being replaced: relora.py:319-322 (F.linear + dropout + two small GEMMs + mul + add).
//
// Roles (256 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0   TMA producer      (one elected lane)         smem ring: full[s] / empty[s]
//   warp 1   MMA issuer        (one elected lane)         tmem ring: tmem_full[a] / tmem_empty[a]
//   warp 2   TMEM allocator
//   warps 4-7 epilogue         (128 threads == 128 TMEM lanes == 128 rows of the tile)
#include <cuda.h>

#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "gemm.h"
#include "tensormap.h"
#include "sm100.cuh"

namespace rb {

#define RB_TR(slot) do { if (p.trace != nullptr && blockIdx.x == 0 && issuer && sl == 0 && tr_e < 512) p.trace[(slot) * 512 + tr_e] = clock64(); } while (0)

long long g_launch_count = 0;
void* g_trace_ptr = nullptr;  // diagnostics: device buffer of 6 x 512 int64 (gemm_set_trace)
void gemm_set_trace(void* p) { g_trace_ptr = p; }
int g_pair_clusters = 0;  // cudaOccupancyMaxActiveClusters of the CTA-pair GEMM (diagnostics)
int gemm_pair_clusters() { return g_pair_clusters; }

using namespace sm100;

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kNumThreads = 256;
constexpr int kEpilogueWarp0 = 4;

struct KernelArgs {
  int M, N, K1, K2;
  int n_per_group, a1_group_kofs, a2_group_kofs;
  int b1_group_kofs, b1_local_n;        // B1: K-window offset per N-group; MN coordinate relative to the group
  int m_per_group, b1_mn_ofs_per_mgroup; // B1: extra MN offset selected by the M-tile's group (weight gradients)
  void* out;
  long long ldc;
  const bf16* residual;
  long long ldr;
  const bf16* bias;  // [N], added in fp32 (Pythia projections carry biases)
  float alpha;
  const float* alpha_dev;  // optional device scalar multiplied into alpha (fp8 dequantisation scales live on the device)
  int fp8;                 // segment 1 (A1, B1) holds fp8 bytes: k-blocks of 128 elements, kind::f8f6f4 MMAs (2: A1 is E5M2)
  int out_f32, accumulate;
  int num_m_tiles, num_n_tiles;
  int tiles_per_group;  // N-tiles per output-column group (the last one of a group may be ragged)
  int use_tma_store;  // bf16 output written through swizzled smem slabs + TMA store
  long long* trace;   // diagnostics: clock64 stamps of CTA 0 (see bench/gemm_trace.py), normally null
  int debug;          // RB_GEMM_DEBUG (diagnostics only): 1 = skip TMA store, 2 = skip TMEM loads, 4 = skip slab writes
  int res_tma;        // ... and the residual is fetched into the same slab by TMA (coalesced) instead of per-row loads
  int split_k;  // >1: each output tile is computed by split_k CTAs over disjoint K ranges, combined with fp32 atomics
};

template <int BLOCK_N, bool PAIR = false>
struct SmemLayout {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;  // 16 KB
  static constexpr int kBRows = PAIR ? BLOCK_N / 2 : BLOCK_N;  // CTA pair: each CTA stages half of the B tile
  static constexpr int kBBytes = kBRows * BLOCK_K * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kStageBytes <= 32768) ? 5 : 4;  // load->full latency is 1000-2200 cycles, one k-block is 512
  static constexpr int kTileBytes = kStages * kStageBytes;
  static constexpr int kBarrierBytes = 1024;                          // barriers + tmem slot (keeps the slabs 1024-aligned)
  static constexpr int kSlabBytes = BLOCK_M * 128;                    // one 128 x 64 bf16 output slab (128B swizzle)
  static constexpr int kSlabs = (kStageBytes <= 32768) ? 4 : 2;       // rotating output slabs; the rest of smem feeds the MMA
  static constexpr int kTotal = kTileBytes + kBarrierBytes + kSlabs * kSlabBytes + 1024;  // +1024 for manual alignment
};

// Issue the TMA loads of one operand tile (BLOCK_MN x BLOCK_K) into `dst`; `bar` is the 32-bit shared address of the
// mbarrier (for CTA pairs: the shared::cluster address of the leader's barrier).
template <int BLOCK_MN, bool MN_MAJOR, bool PAIR = false>
__device__ __forceinline__ void load_operand(const CUtensorMap* map, uint32_t bar, uint8_t* dst, int mn0, int k0, uint64_t hint) {
  auto ld = [&](uint8_t* d, int c0, int c1) {
    if constexpr (PAIR) {
      tma_load_2d_pair(map, bar, d, c0, c1, hint);
    } else {
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cta.global.mbarrier::complete_tx::bytes.L2::cache_hint"
          " [%0], [%1, {%3, %4}], [%2], %5;"
          :
          : "r"(smem_u32(d)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
          : "memory");
    }
  };
  if constexpr (!MN_MAJOR) {
    ld(dst, k0, mn0);  // box {64 (K), BLOCK_MN}
  } else {
#pragma unroll
    for (int j = 0; j < BLOCK_MN / 64; ++j) ld(dst + j * 8192, mn0 + j * 64, k0);  // box {64 (MN), 64 (K)} = 8 KB each
  }
}
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t operand_desc(uint32_t smem_addr, int kstep) {
  if constexpr (!MN_MAJOR) return make_desc_sw128(smem_addr + kstep * (UMMA_K * 2), 16, 1024);
  return make_desc_sw128(smem_addr + kstep * (UMMA_K * 128), 8192, 1024);
}

__device__ __forceinline__ void add_bias8(const bf16* bias, int col, int N, float (&f)[8]) {
  if (bias == nullptr) return;
  if (col + 8 <= N && ((reinterpret_cast<uintptr_t>(bias + col) & 15) == 0)) {
    float b[8];
    unpack8(*reinterpret_cast<const uint4*>(bias + col), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] += b[i];
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (col + i < N) f[i] += __bfloat162float(bias[col + i]);
  }
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool PAIR = false>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a1, const __grid_constant__ CUtensorMap map_b1,
            const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b2,
            const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res, const KernelArgs p) {
  using L = SmemLayout<BLOCK_N, PAIR>;
  constexpr int kStages = L::kStages;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;  // two accumulator stages (256 or 512 columns)
  constexpr int kTileM = PAIR ? 2 * BLOCK_M : BLOCK_M;  // rows of the output tile owned by one CTA / CTA pair
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int work0 = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int work_step = PAIR ? (int)cluster_count_x() : (int)gridDim.x;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* res_bar = tmem_empty_bar + 2;  // residual slab landed (TMA load into the output slab, see epilogue)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bar + L::kSlabs);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (p.trace != nullptr && threadIdx.x == 0) {  // rows 12-15: CTA 0 entry / setup / exit clocks; grid-wide first entry / last exit (ns)
    if (blockIdx.x == 0) { p.trace[12 * 512 + 0] = clock64(); p.trace[13 * 512 + 0] = (long long)global_timer_ns(); }
    atomicMin(reinterpret_cast<unsigned long long*>(p.trace + 14 * 512), global_timer_ns());
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a1);
    tma_prefetch_desc(&map_b1);
    if (p.K2 > 0) {
      tma_prefetch_desc(&map_a2);
      tma_prefetch_desc(&map_b2);
    }
    if (p.use_tma_store) tma_prefetch_desc(&map_out);
    if (p.res_tma) tma_prefetch_desc(&map_res);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], PAIR ? 256 : 128);  // pair: the epilogues of both CTAs release the leader's barrier
    }
    for (int a = 0; a < L::kSlabs; ++a) mbar_init(&res_bar[a], 1);
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_base_slot, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_base_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);  // known warp-uniform: stays in a uniform register
  pdl_wait();  // everything above is CTA-local: it overlaps the tail of the previous kernel in the stream (PDL)
  if (p.trace != nullptr && threadIdx.x == 0 && blockIdx.x == 0) p.trace[12 * 512 + 1] = clock64();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_work = num_tiles * p.split_k;
  const int kstep1 = p.fp8 ? 2 * BLOCK_K : BLOCK_K;  // elements per k-block of segment 1 (always 128 bytes per row)
  const int kb1 = (p.K1 + kstep1 - 1) / kstep1;
  const int kb2 = (p.K2 + BLOCK_K - 1) / BLOCK_K;
  const int num_kb = kb1 + kb2;
  const int kb_per_split = (num_kb + p.split_k - 1) / p.split_k;

  // Producer and MMA issuer: the WHOLE warp walks the loop with uniform control flow and one elected lane issues.  Under an
  // `if (lane == 0)` branch ptxas cannot prove addresses / descriptors warp-uniform and wraps every UTMALDG / UTCHMMA in an
  // ELECT + R2UR + BRA.U.ANY loop: ~100-130 dependent instructions per k-block on one warp = a ~600-cycle floor per k-block
  // (what capped 128-wide tiles at ~45 % and 256-wide tiles at ~83 % of the tensor peak).  Converged, the same loop is ~50
  // uniform-datapath instructions with the four UTCHMMA back to back.
  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    int tr_p = 0;
    const uint32_t full_addr0 = PAIR ? mapa_shared(smem_u32(&full_bar[0]), 0) : smem_u32(&full_bar[0]);
    const int b_half = PAIR ? (int)cta_rank * (BLOCK_N / 2) : 0;  // this CTA's share of the B tile
    for (int work = work0; work < num_work; work += work_step) {
      const int tile = work / p.split_k, split = work % p.split_k;
      const int m0 = (tile / p.num_n_tiles) * kTileM + (int)cta_rank * BLOCK_M;
      const int tn = tile % p.num_n_tiles;
      const int g = tn / p.tiles_per_group;
      const int nl = (tn % p.tiles_per_group) * BLOCK_N;  // column offset inside the group
      const int n0 = g * p.n_per_group + nl;
      const int a1_k = g * p.a1_group_kofs;
      const int a2_k = g * p.a2_group_kofs;
      const int b1_k = g * p.b1_group_kofs;
      const int b1_n = (p.b1_local_n ? nl : n0) + (m0 / p.m_per_group) * p.b1_mn_ofs_per_mgroup;
      const int kb_begin = split * kb_per_split, kb_end = min(num_kb, kb_begin + kb_per_split);
      for (int kb = kb_begin; kb < kb_end; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * L::kStageBytes;
          uint8_t* sb = sa + L::kABytes;
          if (p.trace != nullptr && blockIdx.x == 0 && tr_p < 512) p.trace[0 * 512 + tr_p] = clock64();
          const uint32_t fb = full_addr0 + stage * 8;
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], (PAIR ? 2 : 1) * L::kStageBytes);  // both CTAs' bytes land here
          if (kb < kb1) {
            load_operand<BLOCK_M, A_MN, PAIR>(&map_a1, fb, sa, m0, a1_k + kb * kstep1, kEvictNormal);
            load_operand<L::kBRows, B_MN, PAIR>(&map_b1, fb, sb, b1_n + b_half, b1_k + kb * kstep1, kEvictLast);
          } else {
            const int k = (kb - kb1) * BLOCK_K;
            load_operand<BLOCK_M, false, PAIR>(&map_a2, fb, sa, m0, a2_k + k, kEvictNormal);
            load_operand<L::kBRows, false, PAIR>(&map_b2, fb, sb, n0 + b_half, k, kEvictLast);
          }
        }
        __syncwarp();
        ++tr_p;
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    if (leader) {
      constexpr uint32_t idesc1 = make_idesc_bf16(kTileM, BLOCK_N, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t idesc2 = make_idesc_bf16(kTileM, BLOCK_N, 0, 0);
      const uint32_t idesc8 = p.fp8 == 2 ? make_idesc_e4m3(kTileM, BLOCK_N, 1) : make_idesc_e4m3(kTileM, BLOCK_N, 0);
      // descriptor of k-step 0 + (byte offset >> 4) in the 14-bit start-address field: 32 B per K = 16 step of a K-major
      // tile (also E4M3, K = 32), 2048 B (16 rows of 128 B) per step of an MN-major tile; shared memory is < 256 KB: no carry
      constexpr uint64_t kStepA = A_MN ? 128 : 2, kStepB = B_MN ? 128 : 2;
      auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc_flag) {
        if constexpr (PAIR) umma_f16_ss_pair(d, da, db, idesc, acc_flag);
        else umma_f16_ss(d, da, db, idesc, acc_flag);
      };
      auto mma8 = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t acc_flag) {
        if constexpr (PAIR) umma_f8_ss_pair(d, da, db, idesc8, acc_flag);
        else umma_f8_ss(d, da, db, idesc8, acc_flag);
      };
      auto commit = [&](uint64_t* bar) {
        if constexpr (PAIR) umma_commit_pair(bar, 3);
        else umma_commit(bar);
      };
      const uint32_t smem0 = smem_u32(smem);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tr_m = 0, tr_t = 0;
      for (int work = work0; work < num_work; work += work_step) {
        const int split = work % p.split_k;
        const int kb_begin = split * kb_per_split, kb_end = min(num_kb, kb_begin + kb_per_split);
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_t < 512) p.trace[2 * 512 + tr_t] = clock64();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_m < 512) p.trace[1 * 512 + tr_m] = clock64();
          ++tr_m;
          const uint32_t sa = smem0 + stage * L::kStageBytes;
          const uint32_t sb = sa + L::kABytes;
          if (elect_one()) {
            if (kb < kb1 && p.fp8) {  // E4M3 rows of 128 bytes: four K = 32 steps, same byte offsets as bf16
              const uint64_t da = operand_desc<false>(sa, 0), db = operand_desc<false>(sb, 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) mma8(d_tmem, da + 2 * k, db + 2 * k, ((kb - kb_begin) | k) != 0);
            } else if (kb < kb1) {
              const uint64_t da = operand_desc<A_MN>(sa, 0), db = operand_desc<B_MN>(sb, 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) mma(d_tmem, da + kStepA * k, db + kStepB * k, idesc1, ((kb - kb_begin) | k) != 0);
            } else {
              const uint64_t da = operand_desc<false>(sa, 0), db = operand_desc<false>(sb, 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) mma(d_tmem, da + 2 * k, db + 2 * k, idesc2, ((kb - kb_begin) | k) != 0);
            }
            commit(&empty_bar[stage]);  // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (elect_one()) commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue(s)
        __syncwarp();
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_t < 512) p.trace[3 * 512 + tr_t] = clock64();
        ++tr_t;
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ===================================================================== epilogue
    // TMEM -> registers (two 32-column loads in flight) -> fp32 math -> either
    //   (a) bf16, 128B-swizzled shared-memory slab -> TMA store (full-line, asynchronous, clips ragged edges), or
    //   (b) direct 128-bit global accesses (fp32 outputs, accumulation, split-K atomics).
    const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may access
    const uint32_t et = threadIdx.x - kEpilogueWarp0 * 32;  // 0..127 == row inside the tile
    const bool issuer = (et == 0);
    uint8_t* stage_base = smem + L::kTileBytes + L::kBarrierBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool res_vec = p.residual != nullptr && (p.ldr % 8 == 0) && ((reinterpret_cast<uintptr_t>(p.residual) & 15) == 0);
    const bool out_vec = (p.ldc % (p.out_f32 ? 4 : 8) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0);
    const bool lean = p.bias == nullptr && (p.residual == nullptr || p.res_tma);  // epilogue is alpha (+ TMA residual) only
    const float alpha_eff = p.alpha * (p.alpha_dev != nullptr ? *p.alpha_dev : 1.0f);
    int slab_counter = 0;
    int tr_e = 0;
    // residual slabs travel by TMA into the (swizzled) output slab one slab ahead of their use; each thread then
    // reads back exactly the 16-byte chunks it is about to overwrite
    auto slab_coords = [&](int work_i, int sl_i, int& c_col, int& c_row) {
      const int tile_i = work_i / p.split_k;
      const int tn_i = tile_i % p.num_n_tiles;
      c_row = (tile_i / p.num_n_tiles) * kTileM + (int)cta_rank * BLOCK_M;
      c_col = (tn_i / p.tiles_per_group) * p.n_per_group + (tn_i % p.tiles_per_group) * BLOCK_N + sl_i * 64;
    };
    const uint32_t tmem_empty_addr0 = PAIR ? mapa_shared(smem_u32(&tmem_empty_bar[0]), 0) : 0u;
    if (p.res_tma && issuer && work0 < num_work) {
      int c_col, c_row;
      slab_coords(work0, 0, c_col, c_row);
      mbar_arrive_expect_tx(&res_bar[0], L::kSlabBytes);
      tma_load_2d(&map_res, &res_bar[0], stage_base, c_col, c_row, kEvictNormal);
    }
    for (int work = work0; work < num_work; work += work_step) {
      const int tile = work / p.split_k, split = work % p.split_k;
      const int m0 = (tile / p.num_n_tiles) * kTileM + (int)cta_rank * BLOCK_M;
      const int tn = tile % p.num_n_tiles;
      const int nl = (tn % p.tiles_per_group) * BLOCK_N;
      const int n0 = (tn / p.tiles_per_group) * p.n_per_group + nl;
      const int n_lim = min(p.N, n0 + min(BLOCK_N, p.n_per_group - nl));  // a group's last tile may be ragged
      const bool empty_split = split * kb_per_split >= num_kb;  // nothing was accumulated for this work item
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      if (p.trace != nullptr && blockIdx.x == 0 && issuer && tr_e < 512) p.trace[4 * 512 + tr_e] = clock64();
      const int row = m0 + quad * 32 + lane;
      const bool row_ok = row < p.M;
#pragma unroll 1
      for (int sl = 0; sl < BLOCK_N / 64; ++sl) {
        uint32_t r0[32], r1[32];
        if (!(p.debug & 2)) {
          tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, acc * BLOCK_N + sl * 64), r0);
          tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, acc * BLOCK_N + sl * 64 + 32), r1);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r0[i] = r1[i] = 0x3f800000u;
        }
        RB_TR(6);  // TMEM loads returned
        if (sl == BLOCK_N / 64 - 1) {  // accumulator fully read: hand it back to the MMA warp early
          tc_fence_before();
          if constexpr (PAIR) mbar_arrive_cluster(tmem_empty_addr0 + acc * 8);
          else mbar_arrive(&tmem_empty_bar[acc]);
        }
        const int col0 = n0 + sl * 64;
        if (p.use_tma_store) {
          // ---- (a) bf16 via swizzled smem + TMA store
          // kSlabs buffers rotate; the issuer's wait_read<kSlabs-2> after every commit (two iterations ago, ordered by
          // the barrier of the previous iteration) guarantees the store that last used this buffer has read it
          const int sb_i = slab_counter % L::kSlabs;
          uint8_t* slab = stage_base + sb_i * L::kSlabBytes;
          const uint32_t rloc = quad * 32 + lane;
          uint8_t* rowp = slab + rloc * 128;
          if (p.res_tma) mbar_wait(&res_bar[sb_i], (slab_counter / L::kSlabs) & 1);  // residual slab landed
          uint4 packed[8];
          // Three separately compiled variants behind uniform branches.  A single loop with per-element `if (bias)` /
          // `if (residual)` tests is if-converted by the compiler into ~1600 predicated instructions that are issued
          // (and cost ~1800 cycles per slab) even when every predicate is false.
          if (!lean) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t raw = (q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i];
                f[i] = __uint_as_float(raw) * alpha_eff;
              }
              add_bias8(p.bias, col0 + q * 8, n_lim, f);
              if (p.res_tma) {
                float a[8];
                unpack8(*reinterpret_cast<const uint4*>(rowp + ((q ^ (rloc & 7)) << 4)), a);
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] += a[i];
              } else if (p.residual != nullptr && row_ok) {
                const bf16* rp = p.residual + (long long)row * p.ldr + col0 + q * 8;
                if (res_vec && col0 + q * 8 + 8 <= n_lim) {
                  float a[8];
                  unpack8(*reinterpret_cast<const uint4*>(rp), a);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += a[i];
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (col0 + q * 8 + i < n_lim) f[i] += __bfloat162float(rp[i]);
                }
              }
              packed[q] = pack8(f);
            }
          } else if (p.res_tma) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float f[8], a[8];
              unpack8(*reinterpret_cast<const uint4*>(rowp + ((q ^ (rloc & 7)) << 4)), a);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t raw = (q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i];
                f[i] = fmaf(__uint_as_float(raw), alpha_eff, a[i]);
              }
              packed[q] = pack8(f);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const uint32_t raw = (q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i];
                f[i] = __uint_as_float(raw) * alpha_eff;
              }
              packed[q] = pack8(f);
            }
          }
          RB_TR(7);  // math done
          if (!(p.debug & 4)) {
#pragma unroll
            for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(rowp + ((q ^ (rloc & 7)) << 4)) = packed[q];
          } else if (packed[0].x == 0x12345u) {
            *reinterpret_cast<uint4*>(rowp) = packed[7];
          }
          RB_TR(8);  // slab written
          fence_proxy_async_smem();
          RB_TR(9);  // proxy fence
          named_bar_sync(1, 128);
          RB_TR(10);  // barrier
          if (issuer && !empty_split && col0 < n_lim && !(p.debug & 1)) {
            tma_store_2d(&map_out, slab, col0, m0);
            tma_store_commit();
          }
          if (issuer) tma_store_wait_read<L::kSlabs - 2>();  // frees the buffer of the NEXT slab (see above)
          RB_TR(11);  // store issued + wait_read
          if (p.res_tma && issuer) {  // prefetch the next slab's residual into the next buffer
            int nwork = work, nsl = sl + 1;
            if (nsl == BLOCK_N / 64) {
              nsl = 0;
              nwork = work + work_step;
            }
            if (nwork < num_work) {
              int c_col, c_row;
              slab_coords(nwork, nsl, c_col, c_row);
              const int nb = (slab_counter + 1) % L::kSlabs;
              mbar_arrive_expect_tx(&res_bar[nb], L::kSlabBytes);
              tma_load_2d(&map_res, &res_bar[nb], stage_base + nb * L::kSlabBytes, c_col, c_row, kEvictNormal);
            }
          }
          ++slab_counter;
        } else if (row_ok && !empty_split && col0 < n_lim) {
          // ---- (b) direct global path
          if (p.split_k > 1 && out_vec && col0 + 64 <= n_lim) {
            float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float a0 = __uint_as_float(q < 8 ? r0[q * 4 + 0] : r1[(q - 8) * 4 + 0]) * alpha_eff;
              const float a1 = __uint_as_float(q < 8 ? r0[q * 4 + 1] : r1[(q - 8) * 4 + 1]) * alpha_eff;
              const float a2 = __uint_as_float(q < 8 ? r0[q * 4 + 2] : r1[(q - 8) * 4 + 2]) * alpha_eff;
              const float a3 = __uint_as_float(q < 8 ? r0[q * 4 + 3] : r1[(q - 8) * 4 + 3]) * alpha_eff;
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + q * 4), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
            }
          } else if (p.split_k > 1) {
            float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              const float a0 = __uint_as_float(q < 8 ? r0[q * 4 + 0] : r1[(q - 8) * 4 + 0]) * alpha_eff;
              const float a1 = __uint_as_float(q < 8 ? r0[q * 4 + 1] : r1[(q - 8) * 4 + 1]) * alpha_eff;
              const float a2 = __uint_as_float(q < 8 ? r0[q * 4 + 2] : r1[(q - 8) * 4 + 2]) * alpha_eff;
              const float a3 = __uint_as_float(q < 8 ? r0[q * 4 + 3] : r1[(q - 8) * 4 + 3]) * alpha_eff;
              if (out_vec && col0 + q * 4 + 4 <= n_lim) {  // one 16-byte vector reduction instead of four scalar atomics
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op + q * 4), "f"(a0), "f"(a1), "f"(a2), "f"(a3)
                             : "memory");
              } else {
                const float e[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (col0 + q * 4 + i < n_lim) atomicAdd(op + q * 4 + i, e[i]);
              }
            }
          } else if (p.out_f32 && out_vec && col0 + 64 <= n_lim) {
            float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
            if (p.accumulate) {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                float4 o = *reinterpret_cast<const float4*>(op + q * 4);
                o.x = fmaf(__uint_as_float(q < 8 ? r0[q * 4 + 0] : r1[(q - 8) * 4 + 0]), alpha_eff, o.x);
                o.y = fmaf(__uint_as_float(q < 8 ? r0[q * 4 + 1] : r1[(q - 8) * 4 + 1]), alpha_eff, o.y);
                o.z = fmaf(__uint_as_float(q < 8 ? r0[q * 4 + 2] : r1[(q - 8) * 4 + 2]), alpha_eff, o.z);
                o.w = fmaf(__uint_as_float(q < 8 ? r0[q * 4 + 3] : r1[(q - 8) * 4 + 3]), alpha_eff, o.w);
                *reinterpret_cast<float4*>(op + q * 4) = o;
              }
            } else {
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                float4 o;
                o.x = __uint_as_float(q < 8 ? r0[q * 4 + 0] : r1[(q - 8) * 4 + 0]) * alpha_eff;
                o.y = __uint_as_float(q < 8 ? r0[q * 4 + 1] : r1[(q - 8) * 4 + 1]) * alpha_eff;
                o.z = __uint_as_float(q < 8 ? r0[q * 4 + 2] : r1[(q - 8) * 4 + 2]) * alpha_eff;
                o.w = __uint_as_float(q < 8 ? r0[q * 4 + 3] : r1[(q - 8) * 4 + 3]) * alpha_eff;
                *reinterpret_cast<float4*>(op + q * 4) = o;
              }
            }
          } else if (p.out_f32) {
            float* op = reinterpret_cast<float*>(p.out) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float4 o;
              o.x = __uint_as_float(q < 8 ? r0[q * 4 + 0] : r1[(q - 8) * 4 + 0]) * alpha_eff;
              o.y = __uint_as_float(q < 8 ? r0[q * 4 + 1] : r1[(q - 8) * 4 + 1]) * alpha_eff;
              o.z = __uint_as_float(q < 8 ? r0[q * 4 + 2] : r1[(q - 8) * 4 + 2]) * alpha_eff;
              o.w = __uint_as_float(q < 8 ? r0[q * 4 + 3] : r1[(q - 8) * 4 + 3]) * alpha_eff;
              if (out_vec && col0 + q * 4 + 4 <= n_lim) {
                if (p.accumulate) {
                  const float4 old = *reinterpret_cast<const float4*>(op + q * 4);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(op + q * 4) = o;
              } else {
                const float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (col0 + q * 4 + i < n_lim) op[q * 4 + i] = p.accumulate ? op[q * 4 + i] + e[i] : e[i];
              }
            }
          } else {
            bf16* op = reinterpret_cast<bf16*>(p.out) + (long long)row * p.ldc + col0;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float f[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(q < 4 ? r0[q * 8 + i] : r1[(q - 4) * 8 + i]) * alpha_eff;
              const bool fullv = col0 + q * 8 + 8 <= n_lim;
              add_bias8(p.bias, col0 + q * 8, n_lim, f);
              if (p.residual != nullptr) {
                const bf16* rp = p.residual + (long long)row * p.ldr + col0 + q * 8;
                if (res_vec && fullv) {
                  float a[8];
                  unpack8(*reinterpret_cast<const uint4*>(rp), a);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += a[i];
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    if (col0 + q * 8 + i < n_lim) f[i] += __bfloat162float(rp[i]);
                }
              }
              if (out_vec && fullv) {
                if (p.accumulate) {
                  float a[8];
                  unpack8(*reinterpret_cast<const uint4*>(op + q * 8), a);
#pragma unroll
                  for (int i = 0; i < 8; ++i) f[i] += a[i];
                }
                *reinterpret_cast<uint4*>(op + q * 8) = pack8(f);
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                  if (col0 + q * 8 + i < n_lim) {
                    const float o = p.accumulate ? __bfloat162float(op[q * 8 + i]) + f[i] : f[i];
                    op[q * 8 + i] = __float2bfloat16_rn(o);
                  }
              }
            }
          }
        }
      }
      if (p.trace != nullptr && blockIdx.x == 0 && issuer && tr_e < 512) p.trace[5 * 512 + tr_e++] = clock64();
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (p.use_tma_store && issuer) tma_store_wait<0>();  // all output bytes are globally visible before exit
  }

  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
  if (p.trace != nullptr && threadIdx.x == 0) {
    if (blockIdx.x == 0) { p.trace[12 * 512 + 2] = clock64(); p.trace[13 * 512 + 1] = (long long)global_timer_ns(); }
    atomicMax(reinterpret_cast<unsigned long long*>(p.trace + 15 * 512), global_timer_ns());
  }
}

// =============================================================================================
// Fused backward of a stacked LoRA group (input gradient):
//
//   dx[M,N] = dy[M,Kb] · W[Kb,N]  +  inv_keep · Σ_g keep_g(row, col) ⊙ ( du_g[M,r] · A_g[r,N] )
//
// (Kb = G·Ng: the group's stacked output width; N: the group's input width.)  The dropout mask sits on the LoRA
// *input*, i.e. on the output of this backward product, so the low-rank terms cannot share the accumulator of the
// frozen path.  Each CTA therefore keeps 1+G accumulators per tile in tensor memory: the G short LoRA products are
// issued first, and while the long frozen-path reduction runs on the tensor core the epilogue warps already read
// the LoRA accumulators, evaluate the counter-based masks and hold the combined term packed in registers.  This
// replaces base-GEMM + parts-GEMM + dropout_combine (two extra [M, (G+1)·N] round trips through HBM).
// Reference math: backward of relora.py:319-322 with nn.Dropout on the LoRA input.
struct LoraDxArgs {
  int M, N, Kb, r;
  const bf16* base;  // Kb == 0: the frozen-path product was computed by a separate (256-wide / CTA-pair) GEMM
  long long ld_base;
  int num_m_tiles, num_n_tiles;
  uint32_t thr16;
  float inv_keep;
  const uint32_t* seed_ptr;
  uint32_t keys[3];
};

// One-kernel form (frozen-path product in the same kernel, Kb > 0) with TWO epilogue warpgroups, one per 64-column half of the tile.
template <int G>
__global__ void __launch_bounds__(384, 1)
lora_dx_fused2_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_a,
               const __grid_constant__ CUtensorMap map_out, const LoraDxArgs p) {
  constexpr int BLOCK_N = 128;
  using L = SmemLayout<BLOCK_N>;
  constexpr int kStages = L::kStages;
  constexpr int kBaseStages = (G <= 2) ? 2 : 1;  // TMEM budget: (kBaseStages + kLoraStages·G) · 128 <= 512 columns
  constexpr int kLoraStages = (G == 1) ? 2 : 1;
  constexpr uint32_t kTmemCols = 512;
  static_assert((kBaseStages + kLoraStages * G) * BLOCK_N <= 512, "tensor memory budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* lora_full = empty_bar + kStages;
  uint64_t* lora_empty = lora_full + 2;
  uint64_t* base_full = lora_empty + 2;
  uint64_t* base_empty = base_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(base_empty + 2);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&lora_full[a], 1);
      mbar_init(&lora_empty[a], 256);
      mbar_init(&base_full[a], 1);
      mbar_init(&base_empty[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kb_lora = p.r / BLOCK_K;                     // r is a multiple of 64
  const int kb_base = (p.Kb + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
        for (int kb = 0; kb < G * kb_lora; ++kb) {  // du [M, G·r] K-major ; A [G·r, N] MN-major
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            load_operand<BLOCK_M, false>(&map_du, smem_u32(&full_bar[stage]), sa, m0, kb * BLOCK_K, kEvictNormal);
            load_operand<BLOCK_N, true>(&map_a, smem_u32(&full_bar[stage]), sa + L::kABytes, n0, kb * BLOCK_K, kEvictLast);
          }
          __syncwarp();
          advance();
        }
        for (int kb = 0; kb < kb_base; ++kb) {      // dy [M, Kb] K-major ; W [Kb, N] MN-major
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            load_operand<BLOCK_M, false>(&map_dy, smem_u32(&full_bar[stage]), sa, m0, kb * BLOCK_K, kEvictNormal);
            load_operand<BLOCK_N, true>(&map_w, smem_u32(&full_bar[stage]), sa + L::kABytes, n0, kb * BLOCK_K, kEvictLast);
          }
          __syncwarp();
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      int ls = 0, bs = 0;
      uint32_t ls_phase = 0, bs_phase = 0;
      auto mma_block = [&](uint32_t d_tmem, bool first) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem) + stage * L::kStageBytes;
        const uint32_t sb = sa + L::kABytes;
        if (elect_one()) {
          const uint64_t da = operand_desc<false>(sa, 0), db = operand_desc<true>(sb, 0);  // + (byte offset >> 4) per K = 16 step
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_f16_ss(d_tmem, da + 2 * k, db + 128 * k, idesc, !(first && k == 0));
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&lora_empty[ls], ls_phase ^ 1);
        tc_fence_after();
        for (int g = 0; g < G; ++g) {
          const uint32_t d_tmem = tmem_base + (kBaseStages + ls * G + g) * BLOCK_N;
          for (int kb = 0; kb < kb_lora; ++kb) mma_block(d_tmem, kb == 0);
        }
        if (elect_one()) umma_commit(&lora_full[ls]);
        __syncwarp();
        if (kb_base > 0) {
          mbar_wait(&base_empty[bs], bs_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + bs * BLOCK_N;
          for (int kb = 0; kb < kb_base; ++kb) mma_block(d_tmem, kb == 0);
          if (elect_one()) umma_commit(&base_full[bs]);
          __syncwarp();
        }
        if (++ls == kLoraStages) {
          ls = 0;
          ls_phase ^= 1;
        }
        if (++bs == kBaseStages) {
          bs = 0;
          bs_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ===================================================================== epilogue: warps 4-7 -> columns 0-63, warps 8-11 -> 64-127
    // (the single-warpgroup epilogue was ~3x longer than the tile's MMAs at K = 768: ncu, profiles/ncu/lora_dx_k2560_g1_round2)
    const uint32_t quad = warp & 3;
    const uint32_t half = (warp - kEpilogueWarp0) >> 2;
    const bool issuer = ((warp & 3) == 0) && lane == 0;
    uint8_t* stage_base = smem + L::kTileBytes + L::kBarrierBytes + half * 2 * L::kSlabBytes;  // two private rotating slabs
    const uint32_t seed0 = p.seed_ptr ? *p.seed_ptr : 0u;
    uint32_t seeds[G];
#pragma unroll
    for (int g = 0; g < G; ++g) seeds[g] = mix_seed(seed0, p.keys[g]);
    int ls = 0, bs = 0;
    uint32_t ls_phase = 0, bs_phase = 0;
    int slab_counter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N + half * 64;
      const uint32_t rloc = quad * 32 + lane;
      const uint32_t row = m0 + rloc;
      const uint32_t rowmix = row * 0x9E3779B1u;
      // ---- phase 1: masked sum of this half's LoRA accumulator columns, packed to bf16x2 (overlaps the frozen-path MMAs)
      uint32_t cpk[32];
      mbar_wait(&lora_full[ls], ls_phase);
      tc_fence_after();
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float cf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = 0.f;
        uint32_t rr[G][16];
#pragma unroll
        for (int g = 0; g < G; ++g)
          tmem_ld_32x32b_x16(tmem_addr(tmem_base, quad * 32, (kBaseStages + ls * G + g) * BLOCK_N + half * 64 + c4 * 16), rr[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t sg = rowmix ^ seeds[g];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {  // one hash per column pair (common.cuh:keep_drop)
            const uint32_t cp = (uint32_t)(n0 + c4 * 16 + i) >> 1;
            const uint32_t hsh = lowbias32(sg ^ (cp * 0x85EBCA77u));
            if ((hsh & 0xFFFFu) >= p.thr16) cf[i] += __uint_as_float(rr[g][i]);
            if ((hsh >> 16) >= p.thr16) cf[i + 1] += __uint_as_float(rr[g][i + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cpk[c4 * 8 + i] = pack_bf16x2(cf[2 * i] * p.inv_keep, cf[2 * i + 1] * p.inv_keep);
      }
      tc_fence_before();
      mbar_arrive(&lora_empty[ls]);
      // ---- phase 2: frozen-path accumulator (this half's 64 columns) + combined LoRA term -> bf16 slab -> TMA store
      mbar_wait(&base_full[bs], bs_phase);
      tc_fence_after();
      uint8_t* slab = stage_base + (slab_counter & 1) * L::kSlabBytes;
      uint8_t* rowp = slab + rloc * 128;
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, bs * BLOCK_N + half * 64), r0);
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, bs * BLOCK_N + half * 64 + 32), r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&base_empty[bs]);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float2 c2 = unpack_bf16x2(cpk[q * 4 + i / 2]);
            const uint32_t raw0 = (q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i];
            const uint32_t raw1 = (q < 4) ? r0[q * 8 + i + 1] : r1[(q - 4) * 8 + i + 1];
            f[i] = __uint_as_float(raw0) + c2.x;
            f[i + 1] = __uint_as_float(raw1) + c2.y;
          }
          *reinterpret_cast<uint4*>(rowp + ((q ^ (rloc & 7)) << 4)) = pack8(f);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (issuer) {
        if (n0 < p.N) {
          tma_store_2d(&map_out, slab, n0, m0);
          tma_store_commit();
        }
        tma_store_wait_read<1>();
      }
      named_bar_sync(1 + half, 128);
      ++slab_counter;
      if (++ls == kLoraStages) {
        ls = 0;
        ls_phase ^= 1;
      }
      if (++bs == kBaseStages) {
        bs = 0;
        bs_phase ^= 1;
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

// CTA-pair form of the one-kernel LoRA input gradient (cta_group::2): a pair owns a 256 x 128 tile, each CTA stages its own 128
// rows of dy / du and only HALF (64 columns) of the W / A tile: 24 KB instead of 32 KB per k-block, and EIGHT stages instead of five.
// The 128-wide kernels are bound by load latency, not by bandwidth or the epilogue: a k-block costs ~640 cycles whatever the tile does
// (ncu: tensor pipe 35 %, L2 -> SM 35 % of peak; pair tiles with 5 stages, or 4 epilogue warpgroups, change nothing), which is
// 5 stages / ~3 k cycles of loaded TMA latency.  More k-blocks in flight is what moves it.  MMAs are issued by the leader CTA; both
// CTAs run their own producer and their own two epilogue warpgroups (one output slab each) on their own rows of the accumulator.
struct LoraPairSmem {
  static constexpr int kStages = 8;
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;        // 16 KB: this CTA's 128 rows of dy / du
  static constexpr int kBBytes = 64 * BLOCK_K * 2;             // 8 KB: this CTA's 64 columns of W / A
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTileBytes = kStages * kStageBytes;     // 192 KB
  static constexpr int kBarrierBytes = 1024;
  static constexpr int kSlabBytes = BLOCK_M * 128;
  static constexpr int kTotal = kTileBytes + kBarrierBytes + 2 * kSlabBytes + 1024;
};
static_assert(LoraPairSmem::kTotal <= 232448, "shared memory budget");
template <int G>
__global__ void __launch_bounds__(384, 1)
lora_dx_pair_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_w,
               const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_a,
               const __grid_constant__ CUtensorMap map_out, const LoraDxArgs p) {
  constexpr int BLOCK_N = 128;
  using L = LoraPairSmem;
  constexpr int kStages = L::kStages;
  constexpr int kBaseStages = (G <= 2) ? 2 : 1;  // TMEM budget: (kBaseStages + kLoraStages·G) · 128 <= 512 columns
  constexpr int kLoraStages = (G == 1) ? 2 : 1;
  constexpr uint32_t kTmemCols = 512;
  static_assert((kBaseStages + kLoraStages * G) * BLOCK_N <= 512, "tensor memory budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* lora_full = empty_bar + kStages;
  uint64_t* lora_empty = lora_full + 2;
  uint64_t* base_full = lora_empty + 2;
  uint64_t* base_empty = base_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(base_empty + 2);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int work0 = (int)cluster_id_x(), work_step = (int)cluster_count_x();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_dy);
    tma_prefetch_desc(&map_w);
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&lora_full[a], 1);
      mbar_init(&lora_empty[a], 512);  // the epilogue threads of BOTH CTAs release the leader's accumulators
      mbar_init(&base_full[a], 1);
      mbar_init(&base_empty[a], 512);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_base_slot, kTmemCols);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kb_lora = p.r / BLOCK_K;                     // r is a multiple of 64
  const int kb_base = (p.Kb + BLOCK_K - 1) / BLOCK_K;

  if (warp == 0) {
    // ===================================================================== TMA producer (whole warp, one elected lane issues)
    {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t full_addr0 = mapa_shared(smem_u32(&full_bar[0]), 0);  // loads of both CTAs complete on the leader's barrier
      const int b_half = (int)cta_rank * (BLOCK_N / 2);
      auto advance = [&]() {
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int tile = work0; tile < num_tiles; tile += work_step) {
        const int m0 = (tile / p.num_n_tiles) * 2 * BLOCK_M + (int)cta_rank * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N + b_half;
        for (int kb = 0; kb < G * kb_lora; ++kb) {  // du [M, G·r] K-major ; A [G·r, N] MN-major
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            load_operand<BLOCK_M, false, true>(&map_du, full_addr0 + stage * 8, sa, m0, kb * BLOCK_K, kEvictNormal);
            load_operand<BLOCK_N / 2, true, true>(&map_a, full_addr0 + stage * 8, sa + L::kABytes, n0, kb * BLOCK_K, kEvictLast);
          }
          __syncwarp();
          advance();
        }
        for (int kb = 0; kb < kb_base; ++kb) {      // dy [M, Kb] K-major ; W [Kb, N] MN-major
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2 * L::kStageBytes);
            load_operand<BLOCK_M, false, true>(&map_dy, full_addr0 + stage * 8, sa, m0, kb * BLOCK_K, kEvictNormal);
            load_operand<BLOCK_N / 2, true, true>(&map_w, full_addr0 + stage * 8, sa + L::kABytes, n0, kb * BLOCK_K, kEvictLast);
          }
          __syncwarp();
          advance();
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (whole warp of the leader CTA, one elected lane issues)
    if (leader) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, BLOCK_N, 0, 1);
      int stage = 0;
      uint32_t phase = 0;
      int ls = 0, bs = 0;
      uint32_t ls_phase = 0, bs_phase = 0;
      auto mma_block = [&](uint32_t d_tmem, bool first) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem) + stage * L::kStageBytes;
        const uint32_t sb = sa + L::kABytes;
        if (elect_one()) {
          const uint64_t da = operand_desc<false>(sa, 0), db = operand_desc<true>(sb, 0);  // + (byte offset >> 4) per K = 16 step
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_f16_ss_pair(d_tmem, da + 2 * k, db + 128 * k, idesc, !(first && k == 0));
          umma_commit_pair(&empty_bar[stage], 3);  // frees the slot in both CTAs
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int tile = work0; tile < num_tiles; tile += work_step) {
        mbar_wait(&lora_empty[ls], ls_phase ^ 1);
        tc_fence_after();
        for (int g = 0; g < G; ++g) {
          const uint32_t d_tmem = tmem_base + (kBaseStages + ls * G + g) * BLOCK_N;
          for (int kb = 0; kb < kb_lora; ++kb) mma_block(d_tmem, kb == 0);
        }
        if (elect_one()) umma_commit_pair(&lora_full[ls], 3);
        __syncwarp();
        if (kb_base > 0) {
          mbar_wait(&base_empty[bs], bs_phase ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + bs * BLOCK_N;
          for (int kb = 0; kb < kb_base; ++kb) mma_block(d_tmem, kb == 0);
          if (elect_one()) umma_commit_pair(&base_full[bs], 3);
          __syncwarp();
        }
        if (++ls == kLoraStages) {
          ls = 0;
          ls_phase ^= 1;
        }
        if (++bs == kBaseStages) {
          bs = 0;
          bs_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ===================================================================== epilogue: warps 4-7 -> columns 0-63, warps 8-11 -> 64-127
    // (the single-warpgroup epilogue was ~3x longer than the tile's MMAs at K = 768: ncu, profiles/ncu/lora_dx_k2560_g1_round2)
    const uint32_t quad = warp & 3;
    const uint32_t half = (warp - kEpilogueWarp0) >> 2;
    const bool issuer = ((warp & 3) == 0) && lane == 0;
    uint8_t* slab = smem + L::kTileBytes + L::kBarrierBytes + half * L::kSlabBytes;  // one private slab per warpgroup
    const uint32_t seed0 = p.seed_ptr ? *p.seed_ptr : 0u;
    uint32_t seeds[G];
#pragma unroll
    for (int g = 0; g < G; ++g) seeds[g] = mix_seed(seed0, p.keys[g]);
    int ls = 0, bs = 0;
    uint32_t ls_phase = 0, bs_phase = 0;
    const uint32_t lora_empty0 = mapa_shared(smem_u32(&lora_empty[0]), 0), base_empty0 = mapa_shared(smem_u32(&base_empty[0]), 0);
    for (int tile = work0; tile < num_tiles; tile += work_step) {
      const int m0 = (tile / p.num_n_tiles) * 2 * BLOCK_M + (int)cta_rank * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N + half * 64;
      const uint32_t rloc = quad * 32 + lane;
      const uint32_t row = m0 + rloc;
      const uint32_t rowmix = row * 0x9E3779B1u;
      // ---- phase 1: masked sum of this half's LoRA accumulator columns, packed to bf16x2 (overlaps the frozen-path MMAs)
      uint32_t cpk[32];
      mbar_wait(&lora_full[ls], ls_phase);
      tc_fence_after();
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        float cf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = 0.f;
        uint32_t rr[G][16];
#pragma unroll
        for (int g = 0; g < G; ++g)
          tmem_ld_32x32b_x16(tmem_addr(tmem_base, quad * 32, (kBaseStages + ls * G + g) * BLOCK_N + half * 64 + c4 * 16), rr[g]);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t sg = rowmix ^ seeds[g];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {  // one hash per column pair (common.cuh:keep_drop)
            const uint32_t cp = (uint32_t)(n0 + c4 * 16 + i) >> 1;
            const uint32_t hsh = lowbias32(sg ^ (cp * 0x85EBCA77u));
            if ((hsh & 0xFFFFu) >= p.thr16) cf[i] += __uint_as_float(rr[g][i]);
            if ((hsh >> 16) >= p.thr16) cf[i + 1] += __uint_as_float(rr[g][i + 1]);
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) cpk[c4 * 8 + i] = pack_bf16x2(cf[2 * i] * p.inv_keep, cf[2 * i + 1] * p.inv_keep);
      }
      tc_fence_before();
      mbar_arrive_cluster(lora_empty0 + ls * 8);
      // ---- phase 2: frozen-path accumulator (this half's 64 columns) + combined LoRA term -> bf16 slab -> TMA store
      mbar_wait(&base_full[bs], bs_phase);
      tc_fence_after();
      uint8_t* rowp = slab + rloc * 128;
      {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, bs * BLOCK_N + half * 64), r0);
        tmem_ld_32x32b_x32(tmem_addr(tmem_base, quad * 32, bs * BLOCK_N + half * 64 + 32), r1);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive_cluster(base_empty0 + bs * 8);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; i += 2) {
            const float2 c2 = unpack_bf16x2(cpk[q * 4 + i / 2]);
            const uint32_t raw0 = (q < 4) ? r0[q * 8 + i] : r1[(q - 4) * 8 + i];
            const uint32_t raw1 = (q < 4) ? r0[q * 8 + i + 1] : r1[(q - 4) * 8 + i + 1];
            f[i] = __uint_as_float(raw0) + c2.x;
            f[i + 1] = __uint_as_float(raw1) + c2.y;
          }
          *reinterpret_cast<uint4*>(rowp + ((q ^ (rloc & 7)) << 4)) = pack8(f);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (issuer) {
        if (n0 < p.N) {
          tma_store_2d(&map_out, slab, n0, m0);
          tma_store_commit();
        }
        tma_store_wait_read<0>();  // the slab is rewritten for the next tile after the barrier below
      }
      named_bar_sync(1 + half, 128);
      if (++ls == kLoraStages) {
        ls = 0;
        ls_phase ^= 1;
      }
      if (++bs == kBaseStages) {
        bs = 0;
        bs_phase ^= 1;
      }
    }
    if (issuer) tma_store_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();  // the peer may still read this CTA's shared memory / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, kTmemCols);
  }
}

// =============================================================================================
// Two-kernel form of the LoRA input gradient with the frozen-path product supplied (`base`):
//
//   dx[M,N] = base[M,N] + inv_keep · Σ_g keep_g(row, col) ⊙ ( du_g[M,r] · A_g[r,N] )
//
// The products are tiny (K = r per group); the work is the epilogue (G TMEM reads + one mask hash per column pair and
// group).  ncu on the round-1 single-epilogue-warpgroup version: ~3950 instructions per thread and tile,
// one warp per scheduler, 7.3 stall cycles per issue (barrier 2.1, long scoreboard 1.4, instruction fetch 1.3) -> 41.7 us for
// M 12288 x N 768, G = 3 (0.17 of the HBM roofline).  Here TWO epilogue warpgroups split every 128-column tile into its two
// 64-column slabs (two warps per scheduler, half the instructions each, independent slabs / named barriers / TMA stores).
template <int G>
__global__ void __launch_bounds__(384, 1)
lora_dx_base_kernel(const __grid_constant__ CUtensorMap map_du, const __grid_constant__ CUtensorMap map_a,
                    const __grid_constant__ CUtensorMap map_out, const LoraDxArgs p) {
  constexpr int BLOCK_N = 128;
  using L = SmemLayout<BLOCK_N>;
  constexpr int kStages = L::kStages;
  constexpr int kLoraStages = (G <= 2) ? 2 : 1;  // TMEM budget: kLoraStages · G · 128 <= 512 columns
  constexpr uint32_t kTmemCols = 512;
  static_assert(kLoraStages * G * BLOCK_N <= 512, "tensor memory budget");
  static_assert(L::kSlabs >= 4, "two rotating slabs per epilogue warpgroup");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::kTileBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* lora_full = empty_bar + kStages;
  uint64_t* lora_empty = lora_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(lora_empty + 2);

  const uint32_t warp = warp_id();
  const uint32_t lane = lane_id();
  if (threadIdx.x == 0) pdl_launch_dependents();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_du);
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&lora_full[a], 1);
      mbar_init(&lora_empty[a], 256);  // both epilogue warpgroups release an accumulator stage
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_base_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_base_slot, 0);
  pdl_wait();

  const int num_tiles = p.num_m_tiles * p.num_n_tiles;
  const int kb_lora = p.r / BLOCK_K;

  if (warp == 0) {
    {  // TMA producer (whole warp, one elected lane issues): du [M, G·r] K-major ; A [G·r, N] MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
        const int n0 = (tile % p.num_n_tiles) * BLOCK_N;
        for (int kb = 0; kb < G * kb_lora; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* sa = smem + stage * L::kStageBytes;
            mbar_arrive_expect_tx(&full_bar[stage], L::kStageBytes);
            load_operand<BLOCK_M, false>(&map_du, smem_u32(&full_bar[stage]), sa, m0, kb * BLOCK_K, kEvictNormal);
            load_operand<BLOCK_N, true>(&map_a, smem_u32(&full_bar[stage]), sa + L::kABytes, n0, kb * BLOCK_K, kEvictLast);
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
    {  // MMA issuer (whole warp, one elected lane issues)
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, 1);
      int stage = 0, ls = 0;
      uint32_t phase = 0, ls_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&lora_empty[ls], ls_phase ^ 1);
        tc_fence_after();
        for (int g = 0; g < G; ++g) {
          const uint32_t d_tmem = tmem_base + (ls * G + g) * BLOCK_N;
          for (int kb = 0; kb < kb_lora; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem) + stage * L::kStageBytes;
            const uint32_t sb = sa + L::kABytes;
            if (elect_one()) {
              const uint64_t da = operand_desc<false>(sa, 0), db = operand_desc<true>(sb, 0);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) umma_f16_ss(d_tmem, da + 2 * k, db + 128 * k, idesc, !(kb == 0 && k == 0));
              umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        if (elect_one()) umma_commit(&lora_full[ls]);
        __syncwarp();
        if (++ls == kLoraStages) {
          ls = 0;
          ls_phase ^= 1;
        }
      }
    }
  } else if (warp >= kEpilogueWarp0) {
    // ===================================================================== epilogue: warps 4-7 -> slab 0, warps 8-11 -> slab 1
    const uint32_t quad = warp & 3;                         // TMEM lane quadrant this warp may access
    const uint32_t half = (warp - kEpilogueWarp0) >> 2;     // 64-column slab of the tile handled by this warpgroup
    const bool issuer = ((warp & 3) == 0) && lane == 0;     // one TMA-store issuer per warpgroup
    uint8_t* stage_base = smem + L::kTileBytes + L::kBarrierBytes + half * 2 * L::kSlabBytes;  // two private rotating slabs
    const uint32_t seed0 = p.seed_ptr ? *p.seed_ptr : 0u;
    uint32_t seeds[G];
#pragma unroll
    for (int g = 0; g < G; ++g) seeds[g] = mix_seed(seed0, p.keys[g]);
    int ls = 0;
    uint32_t ls_phase = 0;
    int slab_counter = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / p.num_n_tiles) * BLOCK_M;
      const int n0 = (tile % p.num_n_tiles) * BLOCK_N + half * 64;
      const uint32_t rloc = quad * 32 + lane;
      const uint32_t row = m0 + rloc;
      const uint32_t rowmix = row * 0x9E3779B1u;
      // this thread's 64 columns of the frozen-path product are requested now: the ~1 us global latency hides behind the MMAs
      uint4 bpre[8];
      {
        const bool row_in = (int)row < p.M;
        const bf16* bp = p.base + (long long)row * p.ld_base + n0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          bpre[q] = (row_in && n0 + q * 8 + 8 <= p.N) ? *reinterpret_cast<const uint4*>(bp + q * 8) : make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&lora_full[ls], ls_phase);
      tc_fence_after();
      uint8_t* slab = stage_base + (slab_counter & 1) * L::kSlabBytes;
      uint8_t* rowp = slab + rloc * 128;
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        uint32_t rr[G][16];
#pragma unroll
        for (int g = 0; g < G; ++g)
          tmem_ld_32x32b_x16(tmem_addr(tmem_base, quad * 32, (ls * G + g) * BLOCK_N + half * 64 + c4 * 16), rr[g]);
        tmem_ld_wait();
        if (c4 == 3) {  // last read of the accumulators: hand them back to the MMA warp
          tc_fence_before();
          mbar_arrive(&lora_empty[ls]);
        }
        float cf[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) cf[i] = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint32_t sg = rowmix ^ seeds[g];
#pragma unroll
          for (int i = 0; i < 16; i += 2) {  // one hash per column pair (common.cuh:keep_drop)
            const uint32_t cp = (uint32_t)(n0 + c4 * 16 + i) >> 1;
            const uint32_t hsh = lowbias32(sg ^ (cp * 0x85EBCA77u));
            if ((hsh & 0xFFFFu) >= p.thr16) cf[i] += __uint_as_float(rr[g][i]);
            if ((hsh >> 16) >= p.thr16) cf[i + 1] += __uint_as_float(rr[g][i + 1]);
          }
        }
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          float f[8];
          unpack8(bpre[c4 * 2 + h2], f);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = fmaf(cf[h2 * 8 + i], p.inv_keep, f[i]);
          const int q = c4 * 2 + h2;
          *reinterpret_cast<uint4*>(rowp + ((q ^ (rloc & 7)) << 4)) = pack8(f);
        }
      }
      fence_proxy_async_smem();
      named_bar_sync(1 + half, 128);
      if (issuer) {
        if (n0 < p.N) {
          tma_store_2d(&map_out, slab, n0, m0);
          tma_store_commit();
        }
        tma_store_wait_read<1>();  // the slab used two tiles ago (the next one) has been read by its store
      }
      named_bar_sync(1 + half, 128);  // nobody overwrites the next slab before the issuer's wait returned
      ++slab_counter;
      if (++ls == kLoraStages) {
        ls = 0;
        ls_phase ^= 1;
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

// ---------------------------------------------------------------------------------------------
// host: tensor-map construction (driver entry point resolved at run time; no link against libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    check(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q), "cudaGetDriverEntryPoint");
    if (q != cudaDriverEntryPointSuccess || p == nullptr) throw std::runtime_error("cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

struct MapKey {
  const void* ptr;
  long long inner, outer, ld;
  int box_inner, box_outer, esize;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer && esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&](long long v) { h ^= std::hash<long long>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.box_inner); mix(k.box_outer); mix(k.esize);
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

void gemm_clear_descriptor_cache() {
  std::lock_guard<std::mutex> lk(g_maps_mu);
  g_maps.clear();
}

// 2-D bf16 tensor, `inner` contiguous elements per row, `outer` rows `ld` elements apart, 128B swizzle.
static CUtensorMap make_map_2d(const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer,
                               int esize = 2) {
  MapKey key{ptr, inner, outer, ld, box_inner, box_outer, esize};
  {
    std::lock_guard<std::mutex> lk(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) return it->second;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) throw std::runtime_error("gemm operand pointer must be 16-byte aligned");
  if ((ld * esize) % 16 != 0) throw std::runtime_error("gemm operand leading dimension must be a multiple of 16 bytes");
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  const CUtensorMapDataType dt = esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_ERROR_INVALID_CONTEXT) {
    // the driver call needs the primary context current on THIS thread; autograd worker threads only bind it lazily.
    // cudaSetDevice binds it and, unlike cudaFree(0), is legal while a stream is being captured.
    int dev = 0;
    check(cudaGetDevice(&dev), "cudaGetDevice");
    check(cudaSetDevice(dev), "cudaSetDevice");
    r = get_encode()(&m, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  std::lock_guard<std::mutex> lk(g_maps_mu);
  if (g_maps.size() > 8192) g_maps.clear();  // eager runs see fresh activation pointers every step: bound the cache
  g_maps.emplace(key, m);
  return m;
}

// 3-D bf16 tensor map (128B swizzle, zero fill outside the tensor): dims d0 (contiguous) x d1 x d2 with strides s1 / s2 elements.
// Attention views the packed [rows, 3*hidden] qkv buffer as [head_dim, heads, rows]; a 64-wide box over a 48-wide head is
// zero padded by the TMA unit.
CUtensorMap make_map_3d_bf16(const void* ptr, long long d0, long long d1, long long d2, long long s1, long long s2, int b0, int b1,
                             int b2) {
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (s1 * 2) % 16 != 0 || (s2 * 2) % 16 != 0)
    throw std::runtime_error("tensor map: base and strides must be 16-byte aligned");
  CUtensorMap m;
  cuuint64_t dims[3] = {(cuuint64_t)d0, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)s1 * 2, (cuuint64_t)s2 * 2};
  cuuint32_t box[3] = {(cuuint32_t)b0, (cuuint32_t)b1, (cuuint32_t)b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = CUDA_SUCCESS;
  for (int attempt = 0; attempt < 2; ++attempt) {
    r = get_encode()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_ERROR_INVALID_CONTEXT) break;
    int dev = 0;  // bind the primary context on this thread (legal during stream capture), then retry
    check(cudaGetDevice(&dev), "cudaGetDevice");
    check(cudaSetDevice(dev), "cudaSetDevice");
  }
  if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled(3d) failed with code " + std::to_string((int)r));
  return m;
}

// exported for the other translation units (gemm_mx.cu): same cache, same checks
CUtensorMap make_map_2d_sw128(const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int esize) {
  return make_map_2d(ptr, inner, outer, ld, box_inner, box_outer, esize);
}

// Operand covering `mn` rows/cols of the output dimension and `k` of the reduction dimension.
static CUtensorMap operand_map(const Operand& o, long long mn, long long k, int block_mn, bool fp8 = false) {
  if (fp8) return make_map_2d(o.ptr, k, mn, o.ld, 2 * BLOCK_K, block_mn, 1);  // E4M3 bytes, K-major only
  if (!o.mn_major) return make_map_2d(o.ptr, k, mn, o.ld, BLOCK_K, block_mn);
  return make_map_2d(o.ptr, mn, k, o.ld, 64, BLOCK_K);
}

template <int BLOCK_N, bool A_MN, bool B_MN, bool PAIR = false>
static void launch(const GemmDesc& d, cudaStream_t stream) {
  using L = SmemLayout<BLOCK_N, PAIR>;
  constexpr int kTileM = PAIR ? 2 * BLOCK_M : BLOCK_M;
  KernelArgs p;
  p.M = d.M; p.N = d.N; p.K1 = d.K1; p.K2 = d.K2;
  p.n_per_group = d.n_per_group > 0 ? d.n_per_group : (d.N > 0 ? d.N : 1);
  p.a1_group_kofs = d.a1_group_kofs; p.a2_group_kofs = d.a2_group_kofs;
  p.b1_group_kofs = d.b1_group_kofs; p.b1_local_n = d.b1_local_n ? 1 : 0;
  p.m_per_group = d.m_per_group > 0 ? d.m_per_group : (1 << 30); p.b1_mn_ofs_per_mgroup = d.b1_mn_ofs_per_mgroup;
  p.out = d.out; p.ldc = d.ldc; p.residual = reinterpret_cast<const bf16*>(d.residual); p.ldr = d.ldr;
  p.bias = reinterpret_cast<const bf16*>(d.bias);
  if (d.bias != nullptr && d.out_f32) throw std::runtime_error("gemm: bias is only fused for bf16 outputs");
  p.alpha = d.alpha; p.alpha_dev = d.alpha_dev; p.fp8 = d.fp8 ? (d.fp8_a_e5m2 ? 2 : 1) : 0; p.out_f32 = d.out_f32 ? 1 : 0; p.accumulate = d.accumulate ? 1 : 0;
  p.num_m_tiles = ceil_div(d.M, kTileM);
  const int groups = ceil_div(d.N, p.n_per_group);
  // tiles never straddle a group: the last tile of a group may be ragged (its tail columns are computed but not
  // stored), which is what lets llama_1b's 5504-wide gate/up groups run with 256-wide tiles
  if (groups > 1 && (p.n_per_group % 64) != 0) throw std::runtime_error("gemm: n_per_group must be a multiple of 64");
  p.tiles_per_group = ceil_div(p.n_per_group, BLOCK_N);
  p.num_n_tiles = groups * p.tiles_per_group;

  // K extents of the global tensors include the per-group windows
  const long long a1_k_total = (long long)d.K1 + (long long)(groups - 1) * d.a1_group_kofs;
  if (d.fp8 && (A_MN || B_MN)) throw std::runtime_error("gemm: fp8 operands must be K-major");
  CUtensorMap ma1 = operand_map(d.a1, d.M, a1_k_total, BLOCK_M, d.fp8);
  const int mgroups = d.m_per_group > 0 ? ceil_div(d.M, d.m_per_group) : 1;
  const long long b1_mn_total = (d.b1_local_n ? (long long)p.n_per_group : (long long)d.N) + (long long)(mgroups - 1) * d.b1_mn_ofs_per_mgroup;
  const long long b1_k_total = (long long)d.K1 + (long long)(groups - 1) * d.b1_group_kofs;
  if (d.m_per_group > 0 && (d.m_per_group % kTileM) != 0) throw std::runtime_error("gemm: m_per_group must be a multiple of the M tile");
  CUtensorMap mb1 = operand_map(d.b1, b1_mn_total, b1_k_total, L::kBRows, d.fp8);
  CUtensorMap ma2 = ma1, mb2 = mb1;
  if (d.K2 > 0) {
    if (d.a2.mn_major || d.b2.mn_major) throw std::runtime_error("gemm: the LoRA (A2/B2) operands must be K-major");
    const long long a2_k_total = (long long)d.K2 + (long long)(groups - 1) * d.a2_group_kofs;
    ma2 = operand_map(d.a2, d.M, a2_k_total, BLOCK_M);
    mb2 = operand_map(d.b2, d.N, d.K2, L::kBRows);
  }
  auto kern = gemm_kernel<BLOCK_N, A_MN, B_MN, PAIR>;
  static bool configured = false;
  if (!configured) {
    check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal), "cudaFuncSetAttribute(gemm)");
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int num_kb = ceil_div(d.K1, d.fp8 ? 2 * BLOCK_K : BLOCK_K) + ceil_div(d.K2, BLOCK_K);
  int split = d.split_k;
  if (split == 0) {  // auto: fill the machine when there are few output tiles and a long reduction
    split = 1;
    if (d.out_f32 && d.accumulate && d.residual == nullptr && tiles * 2 <= num_sms() && num_kb >= 16)
      split = std::min(std::min(num_sms() / tiles, num_kb / 4), 64);
  }
  if (split > 1 && !(d.out_f32 && d.accumulate && d.residual == nullptr))
    throw std::runtime_error("gemm: split_k needs an fp32 accumulate output without residual");
  if (split < 1) split = 1;
  // an MMA with no k-blocks would leave the epilogue waiting: every split must own >= 1 k-block or be skipped
  p.split_k = split;
  const int work = tiles * split;
  int grid = work < num_sms() ? work : num_sms();
  if (PAIR) {
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(num_sms() / 2 * 2); qc.blockDim = dim3(kNumThreads); qc.dynamicSmemBytes = L::kTotal;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &qc) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms() / 2;
      }
      max_clusters = n;
      g_pair_clusters = n;
    }
    grid = 2 * (work < max_clusters ? work : max_clusters);  // one CTA pair per work item
  }
  if (grid <= 0) return;
  p.use_tma_store = (!d.out_f32 && !d.accumulate && split == 1 && (d.ldc % 8 == 0) &&
                     (reinterpret_cast<uintptr_t>(d.out) & 15) == 0) ? 1 : 0;
  CUtensorMap mout = ma1, mres = ma1;
  if (p.use_tma_store) mout = make_map_2d(d.out, d.N, d.M, d.ldc, 64, BLOCK_M);
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("RB_GEMM_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
    p.trace = reinterpret_cast<long long*>(g_trace_ptr);
  }
  p.res_tma = (p.use_tma_store && d.residual != nullptr && (d.ldr % 8 == 0) && (reinterpret_cast<uintptr_t>(d.residual) & 15) == 0) ? 1 : 0;
  if (p.res_tma) mres = make_map_2d(d.residual, d.N, d.M, d.ldr, 64, BLOCK_M);
  if (PAIR) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kNumThreads); cfg.dynamicSmemBytes = L::kTotal; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    check(cudaLaunchKernelEx(&cfg, kern, ma1, mb1, ma2, mb2, mout, mres, p), "cudaLaunchKernelEx(gemm pair)");
  } else {
    launch_k(kern, grid, kNumThreads, L::kTotal, stream, ma1, mb1, ma2, mb2, mout, mres, p);
  }
  RB_CHECK_LAUNCH("gemm_kernel");
}

template <int BLOCK_N, bool PAIR = false>
static void dispatch_major(const GemmDesc& d, cudaStream_t s) {
  if (d.a1.mn_major) {
    if (d.b1.mn_major) launch<BLOCK_N, true, true, PAIR>(d, s);
    else launch<BLOCK_N, true, false, PAIR>(d, s);
  } else {
    if (d.b1.mn_major) launch<BLOCK_N, false, true, PAIR>(d, s);
    else launch<BLOCK_N, false, false, PAIR>(d, s);
  }
}

void gemm_bf16(const GemmDesc& d, cudaStream_t stream) {
  if (d.n_lora_acc != 0) throw std::runtime_error("gemm: dropout-combine epilogue not built into this kernel variant");
  if (d.M <= 0 || d.N <= 0) return;
  int bn = d.block_n;
  if (bn == 0) {
    // wide tiles halve the shared-memory bandwidth per MMA; keep 128 when the group size demands it or N is small
    const int npg = d.n_per_group > 0 ? d.n_per_group : d.N;
    // (measured on B200, M = 12288: N=768 K=768 27 -> 23 us, N=768 K=2560 64 -> 49 us, N=2304 K=768 64 -> 45 us)
    const bool groups_ok = (npg % 256 == 0) || (npg >= d.N) || (npg % 64 == 0 && npg >= 2048);  // ragged last tile: <= 6 % waste
    bn = (d.N >= 512 && groups_ok && ceil_div(d.M, BLOCK_M) * ceil_div(d.N, 256) >= num_sms() / 2) ? 256 : 128;
  }
  // CTA pairs (cta_group::2, M = 256 per instruction): each CTA stages only half of the B tile, which cuts the
  // L2 -> shared-memory traffic per FLOP by a third (these GEMMs sit at the ~6.3 KB/clk L2 limit with single CTAs)
  int pair = d.cta_pair;
  // measured (M 12288, N 2304): equal at K = 768, pairs 3-6 % faster from K = 1536 up, single CTAs faster below
  if (pair < 0) pair = (bn == 256 && d.M >= 512 && d.K1 + d.K2 >= 1024 && (d.m_per_group == 0 || d.m_per_group % 256 == 0)) ? 1 : 0;
  if (pair && bn != 256) throw std::runtime_error("gemm: CTA pairs need block_n = 256");
  if (bn == 256) {
    if (pair) dispatch_major<256, true>(d, stream);
    else dispatch_major<256>(d, stream);
  } else {
    dispatch_major<128>(d, stream);
  }
}

template <int G>
static void launch_lora_dx(const LoraDxDesc& d, cudaStream_t stream) {
  using L = SmemLayout<128>;
  LoraDxArgs p;
  p.M = d.M; p.N = d.N; p.Kb = d.Kb; p.r = d.r;
  p.base = reinterpret_cast<const bf16*>(d.base); p.ld_base = d.ld_base;
  p.num_m_tiles = ceil_div(d.M, BLOCK_M);
  p.num_n_tiles = ceil_div(d.N, 128);
  p.thr16 = d.drop_threshold16; p.inv_keep = d.inv_keep; p.seed_ptr = d.seed_ptr;
  for (int g = 0; g < 3; ++g) p.keys[g] = d.seed_key[g];
  CUtensorMap m_du = make_map_2d(d.du, (long long)G * d.r, d.M, d.ld_du, BLOCK_K, BLOCK_M);
  CUtensorMap m_a = make_map_2d(d.a, d.N, (long long)G * d.r, d.ld_a, 64, BLOCK_K);
  CUtensorMap m_out = make_map_2d(d.out, d.N, d.M, d.ldc, 64, BLOCK_M);
  CUtensorMap m_dy = m_du, m_w = m_a;  // unused when the frozen-path product is supplied (Kb == 0)
  if (d.Kb > 0) {
    m_dy = make_map_2d(d.dy, d.Kb, d.M, d.ld_dy, BLOCK_K, BLOCK_M);
    m_w = make_map_2d(d.w, d.N, d.Kb, d.ld_w, 64, BLOCK_K);
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (grid <= 0) return;
  static const bool pair_ok = [] {
    const char* e = getenv("RB_LORA_DX_PAIR");  // 0: single-CTA 128 x 128 tiles, 5 stages (lora_dx_fused2_kernel; A/B timing)
    return e == nullptr || atoi(e) != 0;
  }();
  // measured with the converged issue loops (bench/lora_dx_bench.py): pairs win from a reduction of ~2 K up (1b down dx 145 vs 156 us,
  // 250m qkv dx 45.5 vs 48.7), single CTAs below (250m down dx, Kb = 768: 55.8 vs 66.2 us)
  if (d.Kb >= 1536 && pair_ok && d.M > BLOCK_M) {
    using LP = LoraPairSmem;
    auto kernp = lora_dx_pair_kernel<G>;
    static int max_clusters = 0;
    if (max_clusters == 0) {
      check(cudaFuncSetAttribute(kernp, cudaFuncAttributeMaxDynamicSharedMemorySize, LP::kTotal), "cudaFuncSetAttribute(lora_dx_pair)");
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(num_sms() / 2 * 2); qc.blockDim = dim3(384); qc.dynamicSmemBytes = LP::kTotal;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      qc.attrs = qa; qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kernp, &qc) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        n = num_sms() / 2;
      }
      max_clusters = n;
    }
    p.num_m_tiles = ceil_div(d.M, 2 * BLOCK_M);  // one 256-row tile per CTA pair
    const int ptiles = p.num_m_tiles * p.num_n_tiles;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (ptiles < max_clusters ? ptiles : max_clusters)); cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = LP::kTotal; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
    check(cudaLaunchKernelEx(&cfg, kernp, m_dy, m_w, m_du, m_a, m_out, p), "cudaLaunchKernelEx(lora_dx pair)");
    RB_CHECK_LAUNCH("lora_dx_pair_kernel");
    return;
  }
  {
    auto kern2 = lora_dx_fused2_kernel<G>;
    static bool configured2 = false;
    if (!configured2) {
      check(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal), "cudaFuncSetAttribute(lora_dx_fused2)");
      configured2 = true;
    }
    launch_k(kern2, grid, 384, L::kTotal, stream, m_dy, m_w, m_du, m_a, m_out, p);
    RB_CHECK_LAUNCH("lora_dx_fused2_kernel");
  }
}

template <int G>
static void launch_lora_dx_base(const LoraDxDesc& d, cudaStream_t stream) {
  using L = SmemLayout<128>;
  LoraDxArgs p;
  p.M = d.M; p.N = d.N; p.Kb = 0; p.r = d.r;
  p.base = reinterpret_cast<const bf16*>(d.base); p.ld_base = d.ld_base;
  p.num_m_tiles = ceil_div(d.M, BLOCK_M);
  p.num_n_tiles = ceil_div(d.N, 128);
  p.thr16 = d.drop_threshold16; p.inv_keep = d.inv_keep; p.seed_ptr = d.seed_ptr;
  for (int g = 0; g < 3; ++g) p.keys[g] = d.seed_key[g];
  CUtensorMap m_du = make_map_2d(d.du, (long long)G * d.r, d.M, d.ld_du, BLOCK_K, BLOCK_M);
  CUtensorMap m_a = make_map_2d(d.a, d.N, (long long)G * d.r, d.ld_a, 64, BLOCK_K);
  CUtensorMap m_out = make_map_2d(d.out, d.N, d.M, d.ldc, 64, BLOCK_M);
  auto kern = lora_dx_base_kernel<G>;
  static bool configured = false;
  if (!configured) {
    check(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal), "cudaFuncSetAttribute(lora_dx_base)");
    configured = true;
  }
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (grid <= 0) return;
  launch_k(kern, grid, 384, L::kTotal, stream, m_du, m_a, m_out, p);
  RB_CHECK_LAUNCH("lora_dx_base_kernel");
}

void lora_dx(const LoraDxDesc& d, cudaStream_t stream) {
  if (d.M <= 0 || d.N <= 0) return;
  if (d.groups < 1 || d.groups > 3) throw std::runtime_error("lora_dx: 1..3 stacked LoRA groups");
  if (d.r <= 0 || d.r % BLOCK_K != 0) throw std::runtime_error("lora_dx: the LoRA rank must be a multiple of 64");
  if (d.Kb == 0 && (d.base == nullptr || d.ld_base % 8 != 0 || (reinterpret_cast<uintptr_t>(d.base) & 15) != 0))
    throw std::runtime_error("lora_dx: Kb == 0 needs a 16-byte aligned base product");
  if (d.ldc % 8 != 0 || (reinterpret_cast<uintptr_t>(d.out) & 15) != 0) throw std::runtime_error("lora_dx: output must be 16-byte aligned");
  if (d.Kb == 0) {  // frozen-path product supplied: mask-and-add kernel
    if (d.groups == 1) launch_lora_dx_base<1>(d, stream);
    else if (d.groups == 2) launch_lora_dx_base<2>(d, stream);
    else launch_lora_dx_base<3>(d, stream);
    return;
  }
  if (d.groups == 1) launch_lora_dx<1>(d, stream);
  else if (d.groups == 2) launch_lora_dx<2>(d, stream);
  else launch_lora_dx<3>(d, stream);
}

}  // namespace rb
