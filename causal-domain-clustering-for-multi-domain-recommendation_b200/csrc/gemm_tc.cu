// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma with the
// fp32 accumulator in TMEM -> tcgen05.ld epilogue (bias / ReLU / mask / dropout / bf16-or-fp32 split / split-K) ->
// swizzled shared-memory staging -> TMA store.
// Replaces the nn.Linear / F.linear / autograd matmul call sites of the reference's expert, gate and tower layers
// (layer.py:185,193; ple.py:83-94; mmoe.py:36-40) on the bf16 path.
//
// One persistent CTA per SM, 576 threads: warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..17 = epilogue (warp w may touch TMEM lanes 32*(w%4)..+31; the four warps of a lane quarter take the tile's
// 32-column chunks in turn).  Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, two
// accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1).  Tile = 128 (M) x block_n (runtime,
// multiple of 16, <= 256) x 64 (K); UMMA 128 x block_n x 16.  The number of smem stages is whatever fits next to the
// 32 KB of epilogue staging (4 at block_n = 256).  CTA pairs (cta_group::2, 256 x block_n tiles) for long K loops.
// Operands may be K-major (reduction index contiguous: activations / weights in the forward and input-gradient GEMMs)
// or MN-major (reduction index is the row: weight-gradient GEMMs reduce over the batch) - both are read in place.
//
// Why the epilogue looks the way it does (B200 probes of 2026-10-18, M=65536 N=2560 K=368): with K this short a tile's MMAs
// take ~3k cycles, so the epilogue is on the critical path.  Per-thread row stores, 32 scalar bias loads per chunk and a 64-bit
// dropout hash per element cost 190 / 430 / 1045 us; 8 epilogue warps with 64-column chunks were still busy ~7k cycles per tile
// (cdcmdr_gemm_bf16_tc_profile: the MMA issuer waited 25 % of the kernel for a drained accumulator).  Hence: 16 epilogue warps,
// a per-warp bias slice, an LCG step per column for dropout, 64-byte-row TMA stores from swizzled staging: 167 us, the MMA
// issuer now waits < 3 % for the epilogue and the tile is bound by the tensor pipe's operand reads competing with the TMA fills
// and the staging traffic for shared-memory bandwidth.
#include "tc_ptx.cuh"

namespace cdcmdr {

constexpr int TC_BLOCK_M = 128;
constexpr int TC_BLOCK_K = 64;                       // 64 bf16 = 128 bytes = one swizzle row
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_MAX_N = 256;
constexpr int TC_A_BYTES = TC_BLOCK_M * TC_BLOCK_K * 2;          // 16 KB
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;               // 576
constexpr int TC_STAGING_WARP_BYTES = 32 * 64;                   // 32 rows x 32 bf16 (or 16 fp32), 64B-swizzled
constexpr int TC_STAGING_BYTES = TC_EPI_WARPS * TC_STAGING_WARP_BYTES;
constexpr int TC_BIAS_BYTES = 2 * TC_MAX_N * 4;                  // one bias tile per accumulator stage
constexpr int TC_BAR_BYTES = 512;
constexpr int TC_MASK_BYTES = TC_EPI_WARPS * TC_STAGING_WARP_BYTES;  // ReLU-mask tiles (one 32 x 32 bf16 box per epilogue warp), only with p.tma_mask
constexpr int TC_SMEM_LIMIT = 232448;                            // 227 KB opt-in maximum per CTA
constexpr int TC_TMEM_COLS = 512;                    // 2 accumulator stages x 256 fp32 columns

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_EPI_WARPS) : "memory"); }

struct TcParams {
  int64_t M, N, K;
  int32_t G;
  int32_t a_mn_major, b_mn_major;
  int64_t a_gmn, a_gk, b_gmn, b_gk;          // per-group offsets (elements) along the MN / K index of each operand
  int32_t block_n, n_tiles_n, n_tiles_m, split_k, kb_per_split, num_kb;
  int32_t stages, stage_bytes;
  const float* bias; int64_t bias_gs;
  int64_t n_main;
  uint16_t* out_main; int64_t ld_main, main_gn;
  float* out_aux; int64_t ld_aux, aux_gn, aux_split_stride;
  int32_t act;
  const uint16_t* mask; int64_t ld_mask, mask_gn; float mask_scale;
  float drop_p; const uint64_t* seed_dev; uint32_t salt;
  int32_t accumulate;
  int32_t tma_store;                         // out_main is written through map_c (alignment checked by the host)
  int32_t tma_aux;                           // out_aux (fp32) is written through map_d
  int32_t tma_mask;                          // the ReLU mask (forward activation) is read through map_m into shared memory
  int32_t a_res;                             // A-resident sweep: see the kernel comment
  int32_t epi_subs;                          // epilogue warps per TMEM lane quarter: 4 (576 threads) or 1 (lean: 192 threads)
  int64_t tiles_lo; int32_t tiles_rem;       // a_res: CTA c owns tiles [c*lo + min(c, rem), +lo + (c < rem))
  unsigned long long* prof;                  // diagnostics: cycles spent in each pipeline wait, summed over CTAs (NULL = off)
  const uint16_t* cross_x0; const uint16_t* cross_x; int64_t ld_cross;   // CrossNetV2 epilogue (see cdcmdr.h): bf16 x0 / x boxes through map_m / map_x
  int32_t debug;                             // probe only (cdcmdr_gemm_bf16_tc_mode bits 4..6): 16 = epilogue drains nothing, 64 = tcgen05.ld only,
                                             // 32 = everything but the TMA store.  Results are garbage; never set by the product path.
};

// Epilogue math for 32 consecutive columns [col0, col0+32) of one row; f in/out.  Returns packed bf16 pairs in o[16].
// cycle counters of cdcmdr_gemm_bf16_tc_profile: [0] producer waits for a free stage, [1] MMA waits for a drained accumulator,
// [2] MMA waits for operands, [3] epilogue warp 0 waits for an accumulator, [4] epilogue warp 0 waits for the staging buffer,
// [5] epilogue warp 0 busy (total loop time), [6] CTA lifetime, [7] number of CTAs
#define TC_PROF_T0() const long long prof_t0 = p.prof ? clock64() : 0
#define TC_PROF_ADD(i) do { if (p.prof) atomicAdd(p.prof + (i), (unsigned long long)(clock64() - prof_t0)); } while (0)

struct EpiCtx {
  int act; float mask_scale; bool has_mask, has_drop, has_old; uint32_t s0, thr32; float keep_scale;
};

__device__ __forceinline__ void epi_math32(const EpiCtx& c, float (&f)[32], const float* bias_s /* smem, 32 floats, or null */,
                                           const uint4* mask_row /* 4 x uint4 = 32 bf16, or null */,
                                           const uint4* old_row /* previous output values (accumulate), 32 bf16 */, uint32_t row,
                                           uint32_t col0, uint32_t (&o)[16]) {
  // with dropout the keep scale s = 1/(1-p) > 0 is folded into the accumulator: relu((acc + b)) * s == relu(acc*s + b*s);
  // the bias tile in shared memory is pre-multiplied by s (see the tile prologue)
  if (bias_s) {
    const float4* b4 = reinterpret_cast<const float4*>(bias_s);
    if (c.has_drop) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = b4[j];
        f[4 * j] = fmaf(f[4 * j], c.keep_scale, b.x); f[4 * j + 1] = fmaf(f[4 * j + 1], c.keep_scale, b.y);
        f[4 * j + 2] = fmaf(f[4 * j + 2], c.keep_scale, b.z); f[4 * j + 3] = fmaf(f[4 * j + 3], c.keep_scale, b.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float4 b = b4[j]; f[4 * j] += b.x; f[4 * j + 1] += b.y; f[4 * j + 2] += b.z; f[4 * j + 3] += b.w; }
    }
  } else if (c.has_drop) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] *= c.keep_scale;
  }
  if (c.act == 1) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (c.has_mask) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 mk = mask_row[j];
      const uint32_t w[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // bf16 > 0  <=>  sign bit clear and not zero
        const uint32_t lo = w[k] & 0xFFFFu, hi = w[k] >> 16;
        f[8 * j + 2 * k] = (lo != 0u && lo < 0x8000u) ? f[8 * j + 2 * k] * c.mask_scale : 0.f;
        f[8 * j + 2 * k + 1] = (hi != 0u && hi < 0x8000u) ? f[8 * j + 2 * k + 1] * c.mask_scale : 0.f;
      }
    }
  }
  if (c.has_drop) {
    uint32_t x = drop_lcg_seed(c.s0, row, col0 >> 5);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      x = x * kLcgA + kLcgC;
      f[j] = x >= c.thr32 ? f[j] : 0.f;
    }
  }
  if (c.has_old) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 ov = old_row[j];
      const uint32_t w[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) { f[8 * j + 2 * k] += __uint_as_float(w[k] << 16); f[8 * j + 2 * k + 1] += __uint_as_float(w[k] & 0xFFFF0000u); }
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
}

// CTA2 = true: launched as clusters of two CTAs (one SM pair).  The pair owns a 256 x block_n tile: each CTA loads its own 128
// rows of A and HALF of the B tile (block_n/2 rows of N), rank 0 issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs'
// shared memory and writes each CTA's 128 accumulator rows into that CTA's TMEM; both CTAs run their own epilogue.  Per CTA and
// k-block that is 16 + 16 KB of operands instead of 16 + 32 KB - the level-0 GEMMs are bound by exactly that L2 -> SM traffic.
// Cross-CTA signalling: both CTAs' TMA loads count their bytes on rank 0's "stage full" barrier (cp.async.bulk.tensor
// .cta_group::2), rank 0's tcgen05.commit multicasts "stage free" / "accumulator ready" to both CTAs, both epilogues arrive on
// rank 0's "accumulator drained" barrier.
template <bool CTA2>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_d,
                    const __grid_constant__ CUtensorMap map_m, const __grid_constant__ CUtensorMap map_x, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];               // SWIZZLE_128B tiles need 1024-byte alignment
  // A-resident sweep (p.a_res; short K loops with several column tiles): a CTA owns a CONTIGUOUS range of tiles, column tile
  // fastest, so consecutive tiles share their 128 rows of A.  All k-blocks of that A panel sit in dedicated slots, loaded once per
  // row tile; only B streams through the ring.  At M=65536 N=2560 K=368 the operand fill drops from 288 KB to ~200 KB per tile
  // (1.45 GB -> 1.0 GB per launch): the probes showed the main loop waiting for operands half of its time at the L2 -> SM rate.
  const uint32_t a_region = p.a_res ? (uint32_t)p.num_kb * TC_A_BYTES : 0u;      // [num_kb][128 x 64] resident A panel
  const uint32_t b_in_stage = p.a_res ? 0u : (uint32_t)TC_A_BYTES;              // B's offset inside a ring stage
  uint8_t* ring = smem + a_region;
  uint8_t* staging = ring + p.stages * p.stage_bytes;
  float* bias_s = (float*)(staging + 4 * p.epi_subs * TC_STAGING_WARP_BYTES);
  uint64_t* bars = (uint64_t*)((uint8_t*)bias_s + TC_BIAS_BYTES);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + TC_MAX_STAGES);
  const uint32_t tfull0 = smem_u32(bars + 2 * TC_MAX_STAGES), tempty0 = smem_u32(bars + 2 * TC_MAX_STAGES + 2);
  const uint32_t pfull0 = smem_u32(bars + 2 * TC_MAX_STAGES + 4);            // rank 0 only: "rank 1's stage is full"
  uint32_t* tmem_slot = (uint32_t*)(bars + 3 * TC_MAX_STAGES + 4);
  const uint32_t afull0 = smem_u32(bars + 48), aempty0 = smem_u32(bars + 56);   // a_res: per k-block "A slot full / free"
  const uint32_t mbar0 = smem_u32(bars + 32);                                 // one "mask tile landed" barrier per epilogue warp
  uint8_t* mask_s = (uint8_t*)bars + TC_BAR_BYTES;                             // [TC_EPI_WARPS][32 rows x 64 B], 64B-swizzled
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  const long long prof_cta0 = p.prof ? clock64() : 0;
  const int64_t cta_id = CTA2 ? (int64_t)(blockIdx.x >> 1) : (int64_t)blockIdx.x;
  const int64_t tile0 = p.a_res ? cta_id * p.tiles_lo + min(cta_id, (int64_t)p.tiles_rem) : cta_id;
  const int64_t tile_step = p.a_res ? 1 : (CTA2 ? (int64_t)(gridDim.x >> 1) : (int64_t)gridDim.x);

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();                          // the swizzle atoms below assume it
    for (int i = 0; i < p.stages; ++i) { mbar_init(full0 + 8 * i, 1); mbar_init(empty0 + 8 * i, 1); mbar_init(pfull0 + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, (CTA2 ? 8 : 4) * p.epi_subs); }
    for (int i = 0; i < TC_EPI_WARPS; ++i) mbar_init(mbar0 + 8 * i, 1);
    for (int i = 0; i < 8; ++i) { mbar_init(afull0 + 8 * i, 1); mbar_init(aempty0 + 8 * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    if (p.tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    if (p.tma_aux) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_d) : "memory");
    if (p.tma_mask || p.cross_x0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_m) : "memory");
    if (p.cross_x0) asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();                        // both CTAs' barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (CTA2 && tmem_base != 0u && threadIdx.x == 0) __trap();       // the pair's MMA addresses both CTAs' TMEM with one address

  // p.n_tiles_m counts 128-row tiles (CTA2: 256-row pair tiles; this CTA's rows are tile (2*mt + rank))
  const int64_t tiles_per_group = (int64_t)p.n_tiles_m * p.n_tiles_n * p.split_k;
  const int64_t all_tiles = tiles_per_group * p.G;
  const int64_t total_tiles = p.a_res ? min(all_tiles, tile0 + p.tiles_lo + (cta_id < p.tiles_rem ? 1 : 0)) : all_tiles;   // end of this CTA's tiles
  const int b_cols = CTA2 ? (p.block_n >> 1) : p.block_n;          // N columns of the B tile held by this CTA
  const uint32_t stage_tx = (uint32_t)((p.a_res ? 0 : TC_A_BYTES) + (p.b_mn_major ? ((b_cols + 63) / 64) * 64 : b_cols) * TC_BLOCK_K * 2);

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0, a_phase = 0;
      const uint32_t full_pair0 = CTA2 ? mapa_u32(full0, 0) : full0;               // rank 0's "stage full" barriers
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        const int g = (int)(tile / tiles_per_group);
        int64_t r = tile % tiles_per_group;
        const int z = (int)(r % p.split_k); r /= p.split_k;
        const int nt = (int)(r % p.n_tiles_n);
        const int64_t mt = CTA2 ? 2 * (r / p.n_tiles_n) + rank : r / p.n_tiles_n;
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const int32_t a_mn = (int32_t)(g * p.a_gmn + mt * TC_BLOCK_M);
        const int32_t b_mn = (int32_t)(g * p.b_gmn + (int64_t)nt * p.block_n + (CTA2 ? (int64_t)rank * b_cols : 0));
        const bool new_run = p.a_res && (tile == tile0 || nt == 0);          // first tile of a row tile in this CTA's range
        for (int kb = kb0; kb < kb1; ++kb) {
          if (new_run) {                                                     // this k-block of the new A panel, once per row tile
            mbar_wait(aempty0 + 8 * kb, a_phase ^ 1);
            mbar_expect_tx(afull0 + 8 * kb, TC_A_BYTES);
            tma_load_2d(smem_u32(smem + kb * TC_A_BYTES), &map_a, afull0 + 8 * kb, (int32_t)(g * p.a_gk + (int64_t)kb * TC_BLOCK_K), a_mn);
          }
          { TC_PROF_T0();
            if (CTA2) mbar_wait_cluster(empty0 + 8 * stage, phase ^ 1); else mbar_wait(empty0 + 8 * stage, phase ^ 1);
            TC_PROF_ADD(0); }
          const uint32_t sa = smem_u32(ring + stage * p.stage_bytes), sb = sa + b_in_stage;
          const int32_t ak = (int32_t)(g * p.a_gk + (int64_t)kb * TC_BLOCK_K), bk = (int32_t)(g * p.b_gk + (int64_t)kb * TC_BLOCK_K);
          if (CTA2) {
            // both CTAs' bytes are counted on rank 0's barrier of this stage; rank 0 announces the total
            const uint32_t bar = full_pair0 + 8 * stage;
            if (rank == 0) mbar_expect_tx(full0 + 8 * stage, 2 * stage_tx);
            if (!p.a_mn_major) tma_load_2d_pair(sa, &map_a, bar, ak, a_mn);
            else { tma_load_2d_pair(sa, &map_a, bar, a_mn, ak); tma_load_2d_pair(sa + 8192, &map_a, bar, a_mn + 64, ak); }
            if (!p.b_mn_major) tma_load_2d_pair(sb, &map_b, bar, bk, b_mn);
            else for (int j = 0; j * 64 < b_cols; ++j) tma_load_2d_pair(sb + 8192 * j, &map_b, bar, b_mn + 64 * j, bk);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          const uint32_t bar = full0 + 8 * stage;
          mbar_expect_tx(bar, stage_tx);
          if (p.a_res) {}                                                               // A panel is resident
          else if (!p.a_mn_major) tma_load_2d(sa, &map_a, bar, ak, a_mn);              // box [128 rows (m) x 64 (k)]
          else { tma_load_2d(sa, &map_a, bar, a_mn, ak); tma_load_2d(sa + 8192, &map_a, bar, a_mn + 64, ak); }   // 2 x [64 (k) x 64 (m)]
          if (!p.b_mn_major) tma_load_2d(sb, &map_b, bar, bk, b_mn);                    // box [b_cols rows (n) x 64 (k)]
          else for (int j = 0; j * 64 < b_cols; ++j) tma_load_2d(sb + 8192 * j, &map_b, bar, b_mn + 64 * j, bk);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (new_run) a_phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    if (CTA2 && rank == 1) {
      // rank 1 issues no MMAs; its operand loads are counted on rank 0's barriers by the TMA unit itself
    } else {
      // The whole warp walks the loop converged (every lane polls the barriers) and ONE elected lane issues: everything the
      // tcgen05 instructions take is warp-uniform, so it lives in uniform registers - a `lane == 0` branch made the compiler
      // rebuild both 64-bit descriptors and broadcast five registers per MMA (~150 instructions per k-block on a scheduler
      // the issuer shares with four busy epilogue warps).  Descriptors are a constant high word + a low word that advances by
      // a constant per k-step and per stage.
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=F32 [4,6)=1, A=BF16 [7,10)=1, B=BF16 [10,13)=1,
      // a_major bit 15, b_major bit 16, N>>3 [17,23), M>>4 [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.a_mn_major ? 1 : 0) << 15) |
                             ((uint32_t)(p.b_mn_major ? 1 : 0) << 16) | ((uint32_t)(p.block_n >> 3) << 17) |
                             ((uint32_t)((CTA2 ? 2 * TC_BLOCK_M : TC_BLOCK_M) >> 4) << 24);
      // shared-memory descriptor low word: start >> 4 [0,14) | LBO >> 4 [16,30).  K-major: LBO 16 B, a k-step of 16 bf16 is 32 bytes
      // further inside the 128-byte swizzle row; MN-major: LBO 8192 B, 16 k-rows are 2048 bytes further.  High word (SBO 1024 B,
      // version 1, SWIZZLE_128B) is the constant kDescHi inside the wrappers.
      const uint32_t a_flag = p.a_mn_major ? (512u << 16) : (1u << 16), b_flag = p.b_mn_major ? (512u << 16) : (1u << 16);
      const uint32_t a_kstep = p.a_mn_major ? 128u : 2u, b_kstep = p.b_mn_major ? 128u : 2u;
      // a_res: A k-block kb sits in slot kb of the resident panel (16 KB apart); otherwise at the head of ring stage `stage`
      const uint32_t a_lo0 = (smem_u32(smem) >> 4) | a_flag, b_lo0 = ((smem_u32(ring) + b_in_stage) >> 4) | b_flag;
      uint32_t a_phase = 0;
      const uint32_t stage_step = (uint32_t)p.stage_bytes >> 4;
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
        int64_t r = tile % tiles_per_group;
        const int z = (int)(r % p.split_k);
        const int kb0 = z * p.kb_per_split;
        const int kb1 = min(kb0 + p.kb_per_split, p.num_kb);
        const int nt_i = (int)((r / p.split_k) % p.n_tiles_n);
        const bool new_run = p.a_res && (tile == tile0 || nt_i == 0);
        const bool end_run = p.a_res && (tile + 1 == total_tiles || nt_i == p.n_tiles_n - 1);
        { TC_PROF_T0();
          if (CTA2) mbar_wait_cluster(tempty0 + 8 * acc, acc_phase ^ 1); else mbar_wait(tempty0 + 8 * acc, acc_phase ^ 1);
          if (lane == 0) TC_PROF_ADD(1); }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_MAX_N);
        for (int kb = kb0; kb < kb1; ++kb) {
          { TC_PROF_T0();
            if (new_run) mbar_wait(afull0 + 8 * kb, a_phase);
            if (CTA2) mbar_wait_cluster(full0 + 8 * stage, phase); else mbar_wait(full0 + 8 * stage, phase);
            if (lane == 0) TC_PROF_ADD(2); }
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (p.a_res ? (uint32_t)kb * (TC_A_BYTES >> 4) : (uint32_t)stage * stage_step);
          const uint32_t b_lo = b_lo0 + (uint32_t)stage * stage_step;
          __syncwarp();
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < TC_BLOCK_K / 16; ++k) {
              if (CTA2) tc_mma2_bf16_lo(tmem_d, a_lo + k * a_kstep, b_lo + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else tc_mma_bf16_lo(tmem_d, a_lo + k * a_kstep, b_lo + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (CTA2) tc_commit2_mc(empty0 + 8 * stage); else tc_commit(empty0 + 8 * stage);   // frees the smem slot(s) when these MMAs retire
            if (end_run) tc_commit(aempty0 + 8 * kb);       // last column tile of this row tile: the A slot may take the next panel
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) {
          if (CTA2) tc_commit2_mc(tfull0 + 8 * acc); else tc_commit(tfull0 + 8 * acc);         // accumulator complete -> epilogue(s)
        }
        __syncwarp();
        if (new_run) a_phase ^= 1;
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =============================== epilogue (warps 2..17) ===============================
    // 16 warps: warp w may touch TMEM lanes 32*(w%4)..+31; the four warps of a lane quarter take the tile's 32-column chunks in
    // turn.  The epilogue is SIMT work at one CTA per SM: it is bound by how many warps can hide each other's latencies (8 warps
    // with 64-column chunks spent ~7 k cycles per tile against 3 k cycles of MMAs), hence many narrow chunks rather than few wide ones.
    const int ew = warp - 2;                             // 0..15
    const int n_chunks_c = (p.block_n + 31) >> 5;
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int sub = ew >> 2;                             // which of the four warps of the quarter
    uint8_t* my_stage = staging + ew * TC_STAGING_WARP_BYTES;       // 32 rows x 64 B, 64B-swizzled
    float* bias_t = bias_s + ew * 32;                    // this warp's bias slice: the 32 columns of the chunk in flight
    const uint32_t my_stage_u32 = smem_u32(my_stage);
    // ReLU mask through TMA: the warp's next 32 x 32 mask box streams into its private buffer while the previous chunk is still in
    // the math / store phase (per-lane 64-byte global reads of 32 different rows cost ~5 k cycles per tile in the masked dgrad)
    const uint8_t* my_mask = mask_s + ew * (p.cross_x0 ? 2 * TC_STAGING_WARP_BYTES : TC_STAGING_WARP_BYTES);   // cross: x0 box, then x box
    const uint32_t my_mask_u32 = smem_u32(my_mask), my_mbar = mbar0 + 8 * ew;
    uint32_t mask_phase = 0;
    // chunk (n0 + 32*ci) of this tile takes the TMA-store fast path AND carries a mask
    auto mask_by_tma = [&](int64_t n0_, int ci_) -> bool {
      const int64_t nb_ = n0_ + 32 * ci_;
      return p.tma_mask && ci_ < n_chunks_c && 32 * ci_ + 32 <= p.block_n && nb_ + 32 <= p.n_main;
    };
    // CrossNetV2 epilogue: the chunk's x0 and x boxes (32 rows x 32 bf16 each) land in the warp's buffer behind one barrier
    auto cross_chunk = [&](int64_t n0_, int ci_) -> bool { return p.cross_x0 && ci_ < n_chunks_c && n0_ + 32 * ci_ < p.N; };
    auto issue_cross = [&](int64_t mt_, int64_t nb_) {
      if (lane == 0) {
        mbar_expect_tx(my_mbar, 2 * TC_STAGING_WARP_BYTES);
        tma_load_2d(my_mask_u32, &map_m, my_mbar, (int32_t)nb_, (int32_t)(mt_ * TC_BLOCK_M + q * 32));
        tma_load_2d(my_mask_u32 + TC_STAGING_WARP_BYTES, &map_x, my_mbar, (int32_t)nb_, (int32_t)(mt_ * TC_BLOCK_M + q * 32));
      }
    };
    auto issue_mask = [&](int g_, int64_t mt_, int64_t nb_) {
      if (lane == 0) {
        mbar_expect_tx(my_mbar, TC_STAGING_WARP_BYTES);
        tma_load_2d(my_mask_u32, &map_m, my_mbar, (int32_t)(g_ * p.mask_gn + nb_), (int32_t)(mt_ * TC_BLOCK_M + q * 32));
      }
    };
    int acc = 0; uint32_t acc_phase = 0;
    EpiCtx ec;
    ec.act = p.act; ec.mask_scale = p.mask_scale; ec.has_mask = p.mask != nullptr; ec.has_drop = p.drop_p > 0.f;
    ec.has_old = p.accumulate != 0;
    ec.thr32 = drop_thr32(p.drop_p);
    ec.keep_scale = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
    ec.s0 = p.drop_p > 0.f ? drop_s0(*p.seed_dev, p.salt) : 0u;
    const int n_chunks = (p.block_n + 31) >> 5;
    const int swz = (lane >> 1) & 3;                     // SWIZZLE_64B: 16-byte chunk index XOR ((row / 2) % 4)
    bool store_pending = false;
    const uint32_t tempty_target0 = (CTA2 && rank == 1) ? mapa_u32(tempty0, 0) : tempty0;
    const long long prof_epi0 = p.prof ? clock64() : 0;
    for (int64_t tile = tile0; tile < total_tiles; tile += tile_step) {
      const int g = (int)(tile / tiles_per_group);
      int64_t r = tile % tiles_per_group;
      const int z = (int)(r % p.split_k); r /= p.split_k;
      const int nt = (int)(r % p.n_tiles_n);
      const int64_t mt = CTA2 ? 2 * (r / p.n_tiles_n) + rank : r / p.n_tiles_n;
      const int64_t m = mt * TC_BLOCK_M + q * 32 + lane;
      const int64_t n0 = (int64_t)nt * p.block_n;
      const bool use_bias = p.bias != nullptr && z == 0;
      if (mask_by_tma(n0, sub)) issue_mask(g, mt, n0 + 32 * sub);
      if (cross_chunk(n0, sub)) issue_cross(mt, n0 + 32 * sub);
      { TC_PROF_T0();
        if (CTA2) mbar_wait_cluster(tfull0 + 8 * acc, acc_phase); else mbar_wait(tfull0 + 8 * acc, acc_phase);
        if (ew == 0 && lane == 0) TC_PROF_ADD(3); }
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_MAX_N);
      for (int ci = sub; ci < n_chunks; ci += p.epi_subs) {
        if (p.debug & 16) break;
        const int c0 = ci * 32;
        const int64_t nb = n0 + c0;                      // first global column of the chunk
        if (nb >= p.N) break;
        const int cw = (int)min((int64_t)32, min((int64_t)p.block_n - c0, p.N - nb));     // valid columns in this chunk
        uint32_t v[32];
        tc_ld32_nowait(trow + c0, v);                    // in flight while the bias slice / mask / old values are fetched
        if (use_bias) {
          // this warp's 32 bias values of the chunk, in its private slice.  Columns that take the dropout fold (bf16 main part)
          // hold bias * keep_scale; aux columns (fp32, no dropout) the plain bias.
          __syncwarp();
          const int64_t col = nb + lane;
          float bv = col < p.N ? __ldg(p.bias + g * p.bias_gs + col) : 0.f;
          if (ec.has_drop && col < p.n_main) bv *= ec.keep_scale;
          bias_t[lane] = bv;
          __syncwarp();
        }
        if (p.cross_x0) {
          // ---------------- CrossNetV2 layer (layer.py:339-343): y = x0 * acc + b + x ----------------
          // x0 / x arrive as bf16 boxes by TMA (requested one chunk ahead), y leaves as bf16 and the raw product acc as fp32 (kept for
          // the backward) by TMA stores from the warp's swizzled staging tile: no per-lane row accesses to global memory (a first cut
          // with 128-byte per-lane loads / stores spent ~80 k cycles per tile in the LSU: 32 distinct lines per instruction).
          mbar_wait(my_mbar, mask_phase);
          mask_phase ^= 1;
          uint4 xa[4], xb[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            xa[j] = *reinterpret_cast<const uint4*>(my_mask + lane * 64 + ((j ^ swz) * 16));
            xb[j] = *reinterpret_cast<const uint4*>(my_mask + TC_STAGING_WARP_BYTES + lane * 64 + ((j ^ swz) * 16));
          }
          __syncwarp();                                  // every lane has its rows: the buffer may take the next boxes
          if (cross_chunk(n0, ci + p.epi_subs)) issue_cross(mt, n0 + 32 * (ci + p.epi_subs));
          tc_ld_wait();
          const int32_t row0 = (int32_t)(mt * TC_BLOCK_M + q * 32);
          if (p.out_aux) {                               // acc, fp32: two 16-column boxes
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (store_pending) { if (lane == 0) tma_store_wait_read0(); store_pending = false; }
              __syncwarp();
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int e = 16 * hh + 4 * j;
                *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((j ^ swz) * 16)) = make_uint4(v[e], v[e + 1], v[e + 2], v[e + 3]);
              }
              fence_async_smem();
              __syncwarp();
              if (lane == 0) { tma_store_2d(&map_d, my_stage_u32, (int32_t)(nb + 16 * hh), row0); tma_store_commit(); }
              store_pending = true;
            }
          }
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t wa[4] = {xa[j].x, xa[j].y, xa[j].z, xa[j].w}, wb[4] = {xb[j].x, xb[j].y, xb[j].z, xb[j].w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int e = 8 * j + 2 * k;
              const float y0 = fmaf(__uint_as_float(wa[k] << 16), __uint_as_float(v[e]), bias_t[e]) + __uint_as_float(wb[k] << 16);
              const float y1 = fmaf(__uint_as_float(wa[k] & 0xFFFF0000u), __uint_as_float(v[e + 1]), bias_t[e + 1]) + __uint_as_float(wb[k] & 0xFFFF0000u);
              o[4 * j + k] = pack_bf16x2(y0, y1);
            }
          }
          if (store_pending) { if (lane == 0) tma_store_wait_read0(); store_pending = false; }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((j ^ swz) * 16)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) { tma_store_2d(&map_c, my_stage_u32, (int32_t)nb, row0); tma_store_commit(); }
          store_pending = true;
          continue;
        }
        const bool full_main = nb + 32 <= p.n_main && cw == 32;
        // a chunk cut short by the END OF THE MATRIX may still go through TMA: the store map clips columns >= N
        const bool clip_ok = p.G == 1 && p.n_main == p.N && !p.mask && !p.accumulate && nb + cw == p.N;
        if (p.tma_store && (full_main || clip_ok)) {
          // ---------------- fast path: 32 bf16 columns per row -> swizzled staging -> TMA store ----------------
          uint4 mk[4], old[4];
          if (p.tma_mask) {
            // full_main holds here (clip_ok excludes masks), i.e. mask_by_tma(n0, ci) was true when this box was requested
            mbar_wait(my_mbar, mask_phase);
            mask_phase ^= 1;
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[j] = *reinterpret_cast<const uint4*>(my_mask + lane * 64 + ((j ^ swz) * 16));
            __syncwarp();                                // every lane has its row: the buffer may take the next box
            if (mask_by_tma(n0, ci + p.epi_subs)) issue_mask(g, mt, n0 + 32 * (ci + p.epi_subs));
          } else if (ec.has_mask && m < p.M) {
            const uint4* mp = reinterpret_cast<const uint4*>(p.mask + m * p.ld_mask + g * p.mask_gn + nb);
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[j] = __ldg(mp + j);
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[j] = make_uint4(0, 0, 0, 0);
          }
          if (ec.has_old && m < p.M) {
            const uint4* op4 = reinterpret_cast<const uint4*>(p.out_main + m * p.ld_main + g * p.main_gn + nb);
#pragma unroll
            for (int j = 0; j < 4; ++j) old[j] = op4[j];
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) old[j] = make_uint4(0, 0, 0, 0);
          }
          tc_ld_wait();
          if (p.debug & 64) continue;
          const uint32_t gcol = (uint32_t)(g * p.main_gn + nb);
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          uint32_t o[16];
          epi_math32(ec, f, use_bias ? bias_t : nullptr, mk, old, (uint32_t)m, gcol, o);
          // the epilogue math above overlapped the bulk engine still reading the previous chunk out of the staging buffer
          if (store_pending) {
            if (lane == 0) { TC_PROF_T0(); tma_store_wait_read0(); if (ew == 0) TC_PROF_ADD(4); }
            store_pending = false;
          }
          __syncwarp();
          // row `lane` of the warp's [32 x 64 B] staging tile
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(my_stage + lane * 64 + ((j ^ swz) * 16)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          fence_async_smem();
          __syncwarp();
          if (p.debug & 32) continue;
          if (lane == 0) {
            tma_store_2d(&map_c, my_stage_u32, (int32_t)gcol, (int32_t)(mt * TC_BLOCK_M + q * 32));
            tma_store_commit();
          }
          store_pending = true;
          continue;
        }
        // ---------------- fast fp32 path: 2 x (16 fp32 columns = 64 B per row) -> swizzled staging -> TMA store ----------------
        if (p.tma_aux && nb >= p.n_main && ((cw & 15) == 0 || (p.G == 1 && nb + cw == p.N && !p.accumulate))) {
          tc_ld_wait();
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (16 * hh >= cw) break;
            const int64_t nbb = nb + 16 * hh;
            float4 old[4];
            if (p.accumulate && m < p.M) {
              const float4* op4 = reinterpret_cast<const float4*>(p.out_aux + m * p.ld_aux + g * p.aux_gn + (nbb - p.n_main));
#pragma unroll
              for (int j = 0; j < 4; ++j) old[j] = op4[j];
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) old[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (store_pending) { if (lane == 0) tma_store_wait_read0(); store_pending = false; }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = 16 * hh + 4 * j;
              float4 o = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
              if (use_bias) {
                const float4 b = *reinterpret_cast<const float4*>(bias_t + e);
                o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
              }
              o.x += old[j].x; o.y += old[j].y; o.z += old[j].z; o.w += old[j].w;
              *reinterpret_cast<float4*>(my_stage + lane * 64 + ((j ^ swz) * 16)) = o;
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&map_d, my_stage_u32, (int32_t)(g * p.aux_gn + (nbb - p.n_main)), (int32_t)(mt * TC_BLOCK_M + q * 32));
              tma_store_commit();
            }
            store_pending = true;
          }
          continue;
        }
        // ---------------- general path: direct global stores ----------------
        tc_ld_wait();
        if (m < p.M) {
          const int ncols = cw;
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
          // columns [nb, nb+ncols): main part (bf16) below n_main, aux part (fp32) at/after n_main
          const int n_mainc = (int)max((int64_t)0, min((int64_t)ncols, p.n_main - nb));
          if (ec.has_drop) {                             // same fold as the fast path: main columns carry the keep scale
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < n_mainc) f[j] *= ec.keep_scale;
          }
          if (use_bias) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] += bias_t[j];
          }
          if (n_mainc > 0) {
            if (p.act == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < n_mainc) f[j] = fmaxf(f[j], 0.f);
            }
            uint16_t* op = p.out_main + m * p.ld_main + g * p.main_gn + nb;
            if (p.mask) {
              const uint16_t* mp = p.mask + m * p.ld_mask + g * p.mask_gn + nb;
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < n_mainc) f[j] = bf16_to_f32(mp[j]) > 0.f ? f[j] * p.mask_scale : 0.f;
            }
            if (ec.has_drop) {
              const uint32_t gcol = (uint32_t)(g * p.main_gn + nb);
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < n_mainc) f[j] = drop_keep_lcg(ec.s0, ec.thr32, (uint32_t)m, gcol + j) ? f[j] : 0.f;
            }
            if (p.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < n_mainc) f[j] += bf16_to_f32(op[j]);
            }
            if (n_mainc == 32 && (((uintptr_t)op) & 15) == 0) {
              uint4* o4 = reinterpret_cast<uint4*>(op);
#pragma unroll
              for (int j = 0; j < 4; ++j)
                o4[j] = make_uint4(pack_bf16x2(f[8 * j], f[8 * j + 1]), pack_bf16x2(f[8 * j + 2], f[8 * j + 3]),
                                   pack_bf16x2(f[8 * j + 4], f[8 * j + 5]), pack_bf16x2(f[8 * j + 6], f[8 * j + 7]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j < n_mainc) op[j] = f32_to_bf16(f[j]);
            }
          }
          if (n_mainc < ncols) {
            float* ap = p.out_aux + (int64_t)z * p.aux_split_stride + m * p.ld_aux + g * p.aux_gn + (nb - p.n_main);
            if (p.accumulate && p.split_k == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j >= n_mainc && j < ncols) f[j] += ap[j];
            }
            if (n_mainc == 0 && ncols == 32 && (((uintptr_t)ap) & 15) == 0) {
              float4* a4 = reinterpret_cast<float4*>(ap);
#pragma unroll
              for (int j = 0; j < 8; ++j) a4[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (j >= n_mainc && j < ncols) ap[j] = f[j];
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (CTA2) mbar_arrive_cluster(tempty_target0 + 8 * acc); else mbar_arrive(tempty0 + 8 * acc); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (store_pending && lane == 0) tma_store_wait_all();   // global writes complete before the kernel ends
    if (p.prof && ew == 0 && lane == 0) atomicAdd(p.prof + 5, (unsigned long long)(clock64() - prof_epi0));
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync_all();                        // neither CTA leaves while its partner may still signal it or read its smem
  if (p.prof && threadIdx.x == 0) { atomicAdd(p.prof + 6, (unsigned long long)(clock64() - prof_cta0)); atomicAdd(p.prof + 7, 1ull); }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_TMEM_COLS) : "memory");
  }
}

// sum split-K partials into the destination (fp32), fixed order
__global__ void tc_splitk_reduce_kernel(const float* __restrict__ part, int64_t stride, int splits, float* __restrict__ out, int64_t rows,
                                        int64_t cols, int64_t ld_part, int64_t ld_out, int accumulate) {
  const int64_t total = rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols, r, c);
    float v = 0.f;
    for (int z = 0; z < splits; ++z) v += part[(int64_t)z * stride + r * ld_part + c];
    float* o = out + r * ld_out + c;
    *o = accumulate ? *o + v : v;
  }
}

// the same sums (same order per element) four columns per thread, the slices' loads issued four at a time before their adds: the
// scalar kernel walked the slices one dependent 4-byte load after the other (24 us for the 27 MB of the level-0 weight gradient)
__global__ void __launch_bounds__(256)
tc_splitk_reduce_vec4_kernel(const float* __restrict__ part, int64_t stride, int splits, float* __restrict__ out, int64_t rows,
                             int64_t cols4, int64_t ld_part, int64_t ld_out, int accumulate) {
  const int64_t total = rows * cols4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r, c; split_idx(i, cols4, r, c);
    const float* src = part + r * ld_part + 4 * c;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int z = 0;
    for (; z + 4 <= splits; z += 4) {
      float4 t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)(z + u) * stride));
#pragma unroll
      for (int u = 0; u < 4; ++u) { v.x += t[u].x; v.y += t[u].y; v.z += t[u].z; v.w += t[u].w; }
    }
    for (; z < splits; ++z) {
      const float4 t = __ldcs(reinterpret_cast<const float4*>(src + (int64_t)z * stride));
      v.x += t.x; v.y += t.y; v.z += t.z; v.w += t.w;
    }
    float4* o = reinterpret_cast<float4*>(out + r * ld_out + 4 * c);
    if (accumulate) { const float4 a = *o; v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w; }
    *o = v;
  }
}

// tiled transpose of a bf16 matrix: dst[c, r] = src[r, c]
__global__ void transpose_bf16_kernel(const uint16_t* __restrict__ src, int64_t lds, uint16_t* __restrict__ dst, int64_t ldd, int64_t rows, int64_t cols) {
  __shared__ uint16_t tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? src[r * lds + c] : (uint16_t)0;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) dst[c * ldd + r] = tile[threadIdx.x][i];
  }
}

// bit 0: single-CTA tiles only; bit 1: no A-resident sweep; bit 2: CTA pairs whenever legal (default 0: pairs only for long K loops);
// bit 3: no lean epilogue for long split-K slices
static std::atomic<int> g_tc_mode{0};
static std::atomic<unsigned long long*> g_tc_prof{nullptr};

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" int cdcmdr_gemm_bf16_tc(const cdcmdr_gemm_bf16_t* p, cdcmdr_stream_t s) {
  CDC_REQUIRE(p && p->G >= 1 && p->M >= 0 && p->N >= 0 && p->K >= 1, "bad gemm shape");
  if (p->M == 0 || p->N == 0) return 0;
  CDC_REQUIRE(p->A && p->Bt, "null operand");
  CDC_REQUIRE(((uintptr_t)p->A % 16) == 0 && ((uintptr_t)p->Bt % 16) == 0 && p->lda % 8 == 0 && p->ldb % 8 == 0,
              "TMA needs 16-byte aligned bf16 operands (pointer and leading dimension)");
  CDC_REQUIRE(p->n_main >= 0 && p->n_main <= p->N, "n_main out of range");
  CDC_REQUIRE(p->n_main == 0 || p->out_main, "main output missing");
  CDC_REQUIRE(p->n_main == p->N || p->out_aux, "aux output missing");
  CDC_REQUIRE(p->drop_p <= 0.f || p->seed_dev, "dropout needs a device seed");
  const int split = p->split_k > 1 ? p->split_k : 1;
  CDC_REQUIRE(split == 1 || p->n_main == 0, "split-K writes fp32 partials only");
  CDC_REQUIRE(p->G == 1 || (p->a_gk == 0 && p->b_gk == 0) || p->K % TC_BLOCK_K == 0, "grouped K offsets need K to be a multiple of 64");

  TcParams q{};
  q.M = p->M; q.N = p->N; q.K = p->K; q.G = p->G;
  q.a_mn_major = p->a_mn_major ? 1 : 0; q.b_mn_major = p->b_mn_major ? 1 : 0;
  q.a_gmn = p->a_gm; q.a_gk = p->a_gk; q.b_gmn = p->b_gn; q.b_gk = p->b_gk;
  int bn = p->block_n;
  if (bn <= 0) {
    const int64_t nt = ceil_div(p->N, TC_MAX_N);
    bn = (int)(ceil_div(ceil_div(p->N, nt), 16) * 16);
  }
  if (p->cross_x0 && p->block_n <= 0) bn = (int)(ceil_div(bn, 32) * 32);   // whole 32-column chunks (the x0 / x boxes)
  if (q.b_mn_major) bn = (int)(ceil_div(bn, 64) * 64);            // whole 64-column boxes of the stored [K, N] matrix
  CDC_REQUIRE(bn >= 16 && bn <= TC_MAX_N && bn % 16 == 0, "block_n must be a multiple of 16 in [16, 256]");
  q.block_n = bn;
  q.n_tiles_n = (int)ceil_div(p->N, bn);
  q.n_tiles_m = (int)ceil_div(p->M, TC_BLOCK_M);
  q.num_kb = (int)ceil_div(p->K, TC_BLOCK_K);
  q.split_k = split > q.num_kb ? q.num_kb : split;
  q.kb_per_split = (int)ceil_div(q.num_kb, q.split_k);
  q.split_k = (int)ceil_div(q.num_kb, q.kb_per_split);            // no empty slices
  q.bias = p->bias; q.bias_gs = p->bias_gs; q.n_main = p->n_main;
  q.out_main = p->out_main; q.ld_main = p->ld_main; q.main_gn = p->main_gn;
  q.out_aux = p->out_aux; q.ld_aux = p->ld_aux; q.aux_gn = p->aux_gn; q.aux_split_stride = p->aux_split_stride;
  q.act = p->act; q.mask = p->mask; q.ld_mask = p->ld_mask; q.mask_gn = p->mask_gn; q.mask_scale = p->mask_scale;
  q.drop_p = p->drop_p; q.seed_dev = p->seed_dev; q.salt = p->salt; q.accumulate = p->accumulate;
  q.cross_x0 = p->cross_x0; q.cross_x = p->cross_x; q.ld_cross = p->ld_cross;
  const bool cross = p->cross_x0 != nullptr;
  if (cross) {
    CDC_REQUIRE(p->cross_x && p->out_main && p->bias && p->G == 1 && split == 1 && !p->mask && p->drop_p <= 0.f && !p->accumulate && p->act == 0,
                "cross epilogue: needs x, a bf16 output and a bias; no groups / split-K / mask / dropout / accumulate / activation");
    CDC_REQUIRE(p->N % 32 == 0 && p->n_main == p->N && p->ld_cross % 8 == 0 && p->ld_main % 8 == 0 && ((uintptr_t)p->cross_x0 % 16) == 0 &&
                ((uintptr_t)p->cross_x % 16) == 0 && ((uintptr_t)p->out_main % 16) == 0 && p->M < (int64_t)1 << 31 &&
                (!p->out_aux || (p->ld_aux % 4 == 0 && ((uintptr_t)p->out_aux % 16) == 0)),
                "cross epilogue: N must be a multiple of 32 and every stream TMA-able (16-byte aligned base and pitch)");
  }
  CDC_REQUIRE(split == 1 || q.split_k == 1 || p->aux_split_stride > 0, "split-K needs aux_split_stride");
  if (q.split_k != split && split > 1) {
    // the caller sized its partial buffer for `split` slices; unused slices must read as zero
    CDC_REQUIRE(false, "split_k exceeds the number of non-empty K slices; use cdcmdr_gemm_bf16_tc_splits() to size it");
  }

  // CTA pairs (cta_group::2) when the tile shape allows halving B between the two CTAs and there are at least two row tiles
  q.prof = g_tc_prof.load(std::memory_order_relaxed);
  const int mode = g_tc_mode.load(std::memory_order_relaxed);
  q.debug = mode & 0x70;
  const bool pair_ok = p->M > TC_BLOCK_M && (q.b_mn_major ? bn % 128 == 0 : bn % 32 == 0);
  // auto: pairs pay off once a tile's K loop is long enough to hide the cross-CTA hand-offs (probe: +8 % at 128 k-blocks, -20 % at 6)
  const bool cta2 = (mode & 1) ? false : ((mode & 4) ? pair_ok : (pair_ok && q.kb_per_split >= 32));
  if (cta2) q.n_tiles_m = (int)ceil_div(p->M, 2 * TC_BLOCK_M);
  const int b_cols = cta2 ? bn / 2 : bn;
  CUtensorMap ma, mb, mc;
  if (int rc = make_map(&ma, p->A, p->a_rows, p->a_cols, p->lda, q.a_mn_major ? 64u : (uint32_t)TC_BLOCK_M)) return rc;
  if (int rc = make_map(&mb, p->Bt, p->b_rows, p->b_cols, p->ldb, q.b_mn_major ? 64u : (uint32_t)b_cols)) return rc;
  // bf16 output through TMA stores when its layout allows a tensor map (16-byte aligned base / pitch / group offsets)
  q.tma_store = (p->n_main > 0 && ((uintptr_t)p->out_main % 16) == 0 && p->ld_main % 8 == 0 && p->main_gn % 8 == 0 &&
                 (p->G - 1) * p->main_gn + p->n_main <= p->ld_main && p->M < (int64_t)1 << 31) ? 1 : 0;
  if (q.tma_store) {
    if (int rc = make_map(&mc, p->out_main, p->M, (p->G - 1) * p->main_gn + p->n_main, p->ld_main, 32u, false, true)) return rc;
  } else {
    mc = ma;
  }
  CUtensorMap md = ma, mm = ma;
  const int64_t n_aux = p->N - p->n_main;
  q.tma_aux = (n_aux > 0 && q.split_k == 1 && ((uintptr_t)p->out_aux % 16) == 0 && p->ld_aux % 4 == 0 && p->aux_gn % 4 == 0 &&
               p->n_main % 16 == 0 && (p->G - 1) * p->aux_gn + n_aux <= p->ld_aux && p->M < (int64_t)1 << 31) ? 1 : 0;
  if (q.tma_aux) {
    if (int rc = make_map(&md, p->out_aux, p->M, (p->G - 1) * p->aux_gn + n_aux, p->ld_aux, 32u, true, true)) return rc;
  }
  // ReLU mask through TMA when every chunk that carries it is a whole 32-column TMA-store chunk of a TMA-able matrix
  q.tma_mask = (p->mask && q.tma_store && p->n_main == p->N && p->N % 32 == 0 && bn % 32 == 0 && ((uintptr_t)p->mask % 16) == 0 &&
                p->ld_mask % 8 == 0 && p->mask_gn % 8 == 0 && (p->G - 1) * p->mask_gn + p->n_main <= p->ld_mask && !p->accumulate) ? 1 : 0;
  if (q.tma_mask) {
    if (int rc = make_map(&mm, p->mask, p->M, (p->G - 1) * p->mask_gn + p->n_main, p->ld_mask, 32u, false, true)) return rc;
  }
  CUtensorMap mx = ma;
  if (cross) {
    CDC_REQUIRE(q.tma_store && bn % 32 == 0, "cross epilogue: the bf16 output must be TMA-able");
    if (int rc = make_map(&mm, p->cross_x0, p->M, p->N, p->ld_cross, 32u, false, true)) return rc;
    if (int rc = make_map(&mx, p->cross_x, p->M, p->N, p->ld_cross, 32u, false, true)) return rc;
    if (p->out_aux) { if (int rc = make_map(&md, p->out_aux, p->M, p->N, p->ld_aux, 32u, true, true)) return rc; }
  }
  const int b_bytes = (q.b_mn_major ? (int)ceil_div(b_cols, 64) * 64 : b_cols) * TC_BLOCK_K * 2;
  // Lean epilogue: a split-K slice with a long K loop (weight gradients: K = batch) spends > 90 % of its time in the main loop, so
  // one epilogue warp per TMEM lane quarter drains it as well as four.  192 threads and 24 KB less staging leave registers and
  // shared memory for ~4 small CTAs per SM next to this one: the HBM-bound kernels of the step's side branch (embedding backward)
  // run UNDER the weight-gradient GEMM instead of after it (a 576-thread CTA with 227 KB fills the SM on its own).  Mode bit 3 off.
  q.epi_subs = (!cta2 && !(mode & 8) && q.split_k > 1 && q.kb_per_split >= 48 && !cross && !q.tma_mask) ? 1 : 4;
  const int fixed = 4 * q.epi_subs * TC_STAGING_WARP_BYTES + TC_BIAS_BYTES + TC_BAR_BYTES +
                    (cross ? 2 * TC_MASK_BYTES : (q.tma_mask ? TC_MASK_BYTES : 0));
  // A-resident sweep: K-major A, no split, a short K loop whose whole A panel fits next to >= 3 ring stages of B, and at least
  // two column tiles to sweep.  Mode bit 1 switches it off.
  const int64_t total = (int64_t)q.n_tiles_m * q.n_tiles_n * q.split_k * q.G;
  q.a_res = (!cta2 && !(mode & 2) && !q.a_mn_major && q.split_k == 1 && q.n_tiles_n >= 2 && q.num_kb <= 8 &&
             q.num_kb * TC_A_BYTES + 3 * b_bytes + fixed <= TC_SMEM_LIMIT) ? 1 : 0;
  const int a_region = q.a_res ? q.num_kb * TC_A_BYTES : 0;
  q.stage_bytes = (q.a_res ? 0 : TC_A_BYTES) + b_bytes;
  const int smem_limit = q.epi_subs == 1 ? TC_SMEM_LIMIT - 24576 : TC_SMEM_LIMIT;      // lean: room for the co-resident CTAs
  int stages = (smem_limit - fixed - a_region) / q.stage_bytes;
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  CDC_REQUIRE(stages >= 2, "shared memory budget too small for a 2-stage pipeline");
  q.stages = stages;
  const int smem_bytes = a_region + stages * q.stage_bytes + fixed;
  {
    const int64_t ctas = total < kNumSMs ? total : kNumSMs;
    q.tiles_lo = total / ctas; q.tiles_rem = (int32_t)(total % ctas);
  }

  static bool attr_set = false;
  if (!attr_set) {
    CDC_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    CDC_CHECK(cudaFuncSetAttribute(gemm_bf16_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_LIMIT));
    attr_set = true;
  }
  if (!cta2) {
    const int grid = (int)(total < kNumSMs ? total : kNumSMs);
    gemm_bf16_tc_kernel<false><<<grid, 64 + 128 * q.epi_subs, smem_bytes, to_stream(s)>>>(ma, mb, mc, md, mm, mx, q);
    CDC_LAUNCHED();
    return 0;
  }
  const int64_t clusters = total < kNumSMs / 2 ? total : kNumSMs / 2;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * clusters), 1, 1);
  cfg.blockDim = dim3(TC_THREADS, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = to_stream(s);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  CDC_CHECK(cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<true>, ma, mb, mc, md, mm, mx, q));
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_gemm_bf16_tc_profile(uint64_t* counters8) {
  g_tc_prof.store(reinterpret_cast<unsigned long long*>(counters8), std::memory_order_relaxed);
  return 0;
}

extern "C" int cdcmdr_gemm_bf16_tc_mode(int mode) {
  const int old = g_tc_mode.load(std::memory_order_relaxed);
  if (mode >= 0) g_tc_mode.store(mode & 0x7f, std::memory_order_relaxed);
  return old;
}

extern "C" int cdcmdr_gemm_bf16_tc_splits(int64_t K, int32_t want) {
  // number of split-K slices the kernel will actually use for `want` requested slices
  const int num_kb = (int)ceil_div(K, TC_BLOCK_K);
  int s = want < 1 ? 1 : want;
  if (s > num_kb) s = num_kb;
  const int per = (int)ceil_div(num_kb, s);
  return (int)ceil_div(num_kb, per);
}

extern "C" int cdcmdr_splitk_reduce(const float* part, int64_t stride, int32_t splits, float* out, int64_t rows, int64_t cols,
                                    int64_t ld_part, int64_t ld_out, int32_t accumulate, cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  int64_t g = ceil_div(rows * cols, 256);
  if (g > 8 * kNumSMs) g = 8 * kNumSMs;
  if (cols % 4 == 0 && ld_part % 4 == 0 && ld_out % 4 == 0 && stride % 4 == 0 && ((uintptr_t)part % 16) == 0 && ((uintptr_t)out % 16) == 0) {
    g = ceil_div(rows * (cols / 4), 256);
    if (g > 8 * kNumSMs) g = 8 * kNumSMs;
    tc_splitk_reduce_vec4_kernel<<<(int)g, 256, 0, to_stream(s)>>>(part, stride, splits, out, rows, cols / 4, ld_part, ld_out, accumulate);
  } else {
    tc_splitk_reduce_kernel<<<(int)g, 256, 0, to_stream(s)>>>(part, stride, splits, out, rows, cols, ld_part, ld_out, accumulate);
  }
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_transpose_bf16(const uint16_t* src, int64_t lds, uint16_t* dst, int64_t ldd, int64_t rows, int64_t cols, cdcmdr_stream_t s) {
  if (rows <= 0 || cols <= 0) return 0;
  dim3 grid((unsigned)ceil_div(cols, 32), (unsigned)ceil_div(rows, 32));
  CDC_REQUIRE(grid.y <= 65535, "transpose: too many rows");
  transpose_bf16_kernel<<<grid, dim3(32, 8), 0, to_stream(s)>>>(src, lds, dst, ldd, rows, cols);
  CDC_LAUNCHED();
  return 0;
}
