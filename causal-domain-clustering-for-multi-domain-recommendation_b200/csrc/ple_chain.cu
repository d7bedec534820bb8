// R1 (north_star 2; reference model/ple.py:96-124 + model/layer.py:184-190): the first CGC level of PLE as ONE kernel.
//
// At level 0 every expert MLP and every gate reads the same gathered embeddings X [B, K0] (ple.py:54).  The generic path runs
// layer 0 of all experts as one concatenated-N GEMM, writes its [B, nE*d0] activation to HBM, and reads it back for the grouped
// layer-1 GEMM.  Here a CTA owns 128 rows of X (resident in shared memory for the whole tile) and walks the experts:
//
//   for e:  Z0 = X W0_e^T          tcgen05.mma, 128 x 128 accumulator units in TMEM (two per expert at d0 = 256)
//           A0_e = drop(relu(Z0+b0)) epilogue -> bf16, written STRAIGHT into shared memory in the 128-byte-swizzled K-major layout
//                                     the next MMA reads (and, in training, TMA-stored to HBM from the same tile for the backward)
//           Z1 = A0_e W1_e^T        tcgen05.mma, A operand = the tile the epilogue just wrote
//           H_e = drop(relu(Z1+b1)) epilogue -> bf16 staging tile -> TMA store
//   gate / wide-linear logits = X Wg^T ride along as one more small MMA job per tile (fp32 out).
//
// In inference the [B, nE*d0] activation never exists in HBM.  The gate softmax + expert mixing stay in cdcmdr_gate_mix_fwd: the
// mixed outputs of all gates of a 128-row tile are 5 x 128 x 128 fp32 accumulators (320 KB) - more than the register file and the
// TMEM columns left next to the pipeline below can hold ("why the mix is not in the chain kernel").
//
// Roles (576 threads, one persistent CTA per SM): warp 0 = TMA producer (X panel once per tile, then every weight block in the
// order the MMA warp consumes them, through a ring of 16 KB slots), warp 1 = MMA issuer (all lanes walk the loop converged, one
// elected lane issues), warps 2..17 = epilogue (warp w owns TMEM lane quarter w % 4; the four warps of a quarter take the four
// 32-column chunks of a 128-column accumulator unit).
// TMEM: four 128-column accumulator slots used as ONE in-order ring by the job sequence
//   [gates], M1(0,0..), then per expert e:  M1(e+1, 0), M2(e), M1(e+1, 1..)
// which both the MMA warp and the epilogue follow, so the layer-0 MMAs of expert e+1 run under the epilogues of expert e and the
// layer-1 MMA of expert e under the epilogue of M1(e+1, 0).
// Shared memory: X panel (<= 6 x 16 KB) | weight ring (4 x 16 KB) | activation blocks (4 x 16 KB: block kb = columns 64kb.. of
// A0_e; the H staging tiles reuse blocks 2, 3 between M2(e) and the epilogue of M1(e+1, 1)) | per-warp bias slices | barriers.
#include "tc_ptx.cuh"

namespace cdcmdr {

constexpr int CH_THREADS = 576;
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_BLK = 16384;                        // [128 rows x 64 bf16], 128-byte swizzle
constexpr int CH_W_SLOTS = 4;
constexpr int CH_A_SLOTS = 4;
constexpr int CH_MAX_KB0 = 6;
constexpr int CH_SMEM_LIMIT = 232448;
constexpr int CH_TMEM_COLS = 512;

struct ChainParams {
  int64_t B;
  int32_t num_kb0, nE, d0, d1, n_g, n_g16, nU, KB1;
  const float* b0; const float* b1;
  float* Lg; int64_t ldg;
  int32_t store_a0;
  float drop_p; const uint64_t* seed_dev; uint32_t salt0, salt1;
  unsigned long long* prof;                  // cdcmdr_ple_chain_profile: cycles per pipeline wait, summed over CTAs (NULL = off)
};

// counters: [0] producer waits for a free weight slot, [1] MMA waits for a free accumulator slot, [2] MMA waits for weights,
// [3] MMA waits for activation blocks, [4] epilogue warp 0 waits for an accumulator, [5] epilogue warp 0 waits for a free
// activation block, [6] epilogue warp 0 lifetime, [7] CTAs, [8] MMA warp lifetime, [9] producer lifetime
// The counters are compiled in only with -DCDCMDR_CHAIN_PROF (tools/probe_chain.py --prof builds that variant): even the untaken
// `if (prof)` branches around every barrier wait of the MMA warp cost 40 us per launch at the C4 shape.
#ifdef CDCMDR_CHAIN_PROF
#define CH_PROF(ptr) (ptr)
#else
#define CH_PROF(ptr) ((unsigned long long*)nullptr)
#endif
#define CH_T0() const long long ch_t0 = CH_PROF(p.prof) ? clock64() : 0
#define CH_ADD(i) do { if (CH_PROF(p.prof)) atomicAdd(p.prof + (i), (unsigned long long)(clock64() - ch_t0)); } while (0)

__device__ __forceinline__ void pair_bar_sync(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

__global__ void __launch_bounds__(CH_THREADS, 1)
ple_chain_fwd_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w0,
                     const __grid_constant__ CUtensorMap map_wg, const __grid_constant__ CUtensorMap map_w1,
                     const __grid_constant__ CUtensorMap map_a0, const __grid_constant__ CUtensorMap map_h, const ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* xs = smem;                                              // [num_kb0][128 x 64]
  uint8_t* wr = xs + p.num_kb0 * CH_BLK;                           // weight ring
  uint8_t* ar = wr + CH_W_SLOTS * CH_BLK;                          // activation blocks
  float* bias_s = (float*)(ar + CH_A_SLOTS * CH_BLK);              // [16 warps][32]
  uint64_t* bars = (uint64_t*)(bias_s + CH_EPI_WARPS * 32);
  const uint32_t xfull0 = smem_u32(bars), xempty = smem_u32(bars + 8);
  const uint32_t wfull0 = smem_u32(bars + 9), wempty0 = smem_u32(bars + 13);
  const uint32_t afull0 = smem_u32(bars + 17), aempty0 = smem_u32(bars + 21);
  const uint32_t tfull0 = smem_u32(bars + 25), tempty0 = smem_u32(bars + 29);
  uint32_t* tmem_slot = (uint32_t*)(bars + 33);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int i = 0; i < 8; ++i) mbar_init(xfull0 + 8 * i, 1);
    mbar_init(xempty, 1);
    for (int i = 0; i < CH_W_SLOTS; ++i) { mbar_init(wfull0 + 8 * i, 1); mbar_init(wempty0 + 8 * i, 1); }
    for (int i = 0; i < CH_A_SLOTS; ++i) { mbar_init(afull0 + 8 * i, 4); mbar_init(aempty0 + 8 * i, 1); }   // full: one arrival per lane quarter
    for (int i = 0; i < 4; ++i) { mbar_init(tfull0 + 8 * i, 1); mbar_init(tempty0 + 8 * i, CH_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_h) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(CH_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t n_tiles = ceil_div(p.B, 128);
  const int nkb = p.num_kb0, nE = p.nE, nU = p.nU, KB1 = p.KB1;
  // Every CTA walks the experts of a tile in the same cyclic order but STARTS at a different one: all 148 CTAs run in near lockstep,
  // and without the rotation they all request the same 16 KB weight block (the same 128 L2 lines) at the same moment.
  const int rot = (int)(blockIdx.x % (unsigned)nE);
  auto pe = [&](int e) -> int { const int r = e + rot; return r >= nE ? r - nE : r; };
  // ... and the CTAs that share a starting expert walk the k-blocks of a layer-0 unit from different starting points (the fp32
  // accumulation order of a row then depends on the CTA that owns its tile - fixed for a given B, bit-repeatable)
  const int krot = (int)((blockIdx.x / (unsigned)nE) % (unsigned)nkb);
  auto kbo = [&](int kb) -> int { const int r = kb + krot; return r >= nkb ? r - nkb : r; };

  if (warp == 0) {
    // =============================== TMA producer ===============================
    if (lane == 0) {
      int ws = 0; uint32_t wph = 0, xph = 0;
      const long long prod_t0 = CH_PROF(p.prof) ? clock64() : 0;
      auto load_w = [&](const CUtensorMap* map, int32_t col, int32_t row, uint32_t bytes) {
        { CH_T0(); mbar_wait(wempty0 + 8 * ws, wph ^ 1); CH_ADD(0); }
        mbar_expect_tx(wfull0 + 8 * ws, bytes);
        tma_load_2d(smem_u32(wr + ws * CH_BLK), map, wfull0 + 8 * ws, col, row);
        if (++ws == CH_W_SLOTS) { ws = 0; wph ^= 1; }
      };
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int32_t m0 = (int32_t)(tile * 128);
        mbar_wait(xempty, xph ^ 1);                                // the previous tile's layer-0 MMAs have read the panel
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_expect_tx(xfull0 + 8 * kb, CH_BLK);
          tma_load_2d(smem_u32(xs + kb * CH_BLK), &map_x, xfull0 + 8 * kb, kb * 64, m0);
        }
        xph ^= 1;
        if (p.n_g > 0)
          for (int kb = 0; kb < nkb; ++kb) load_w(&map_wg, kbo(kb) * 64, nE * p.d0, (uint32_t)p.n_g16 * 128u);
        for (int u = 0; u < nU; ++u)
          for (int kb = 0; kb < nkb; ++kb) load_w(&map_w0, kbo(kb) * 64, pe(0) * p.d0 + u * 128, CH_BLK);
        for (int e = 0; e < nE; ++e) {
          if (e + 1 < nE)
            for (int kb = 0; kb < nkb; ++kb) load_w(&map_w0, kbo(kb) * 64, pe(e + 1) * p.d0, CH_BLK);
          for (int kb = 0; kb < KB1; ++kb) load_w(&map_w1, kb * 64, pe(e) * p.d1, (uint32_t)p.d1 * 128u);
          if (e + 1 < nE)
            for (int u = 1; u < nU; ++u)
              for (int kb = 0; kb < nkb; ++kb) load_w(&map_w0, kbo(kb) * 64, pe(e + 1) * p.d0 + u * 128, CH_BLK);
        }
      }
      if (CH_PROF(p.prof)) atomicAdd(p.prof + 9, (unsigned long long)(clock64() - prod_t0));
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // instruction descriptor: D = F32 [4,6) = 1, A = B = BF16 ([7,10) = 1, [10,13) = 1), K-major operands, N >> 3 [17,23), M >> 4 [24,29)
    const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc_m1 = idesc0 | ((uint32_t)(128 >> 3) << 17);
    const uint32_t idesc_m2 = idesc0 | ((uint32_t)(p.d1 >> 3) << 17);
    const uint32_t idesc_g = idesc0 | ((uint32_t)(p.n_g16 >> 3) << 17);
    const uint32_t kflag = 1u << 16;                               // LBO 16 B (K-major, 128-byte swizzle); high word = kDescHi
    const uint32_t x_lo0 = (smem_u32(xs) >> 4) | kflag, w_lo0 = (smem_u32(wr) >> 4) | kflag, a_lo0 = (smem_u32(ar) >> 4) | kflag;
    int ws = 0; uint32_t wph = 0, xph = 0;
    uint32_t job = 0, ae = 0;                                      // accumulator-ring position; experts done (parity of the activation blocks)
    uint32_t slot = 0;
    const long long mma_t0 = CH_PROF(p.prof) ? clock64() : 0;
    auto begin_job = [&]() -> uint32_t {
      slot = job & 3u;
      { CH_T0(); mbar_wait(tempty0 + 8 * slot, ((job >> 2) & 1u) ^ 1u); if (lane == 0) CH_ADD(1); }
      tc_fence_after();
      return tmem_base + slot * 128u;
    };
    auto end_job = [&]() {
      if (elect_one()) tc_commit(tfull0 + 8 * slot);
      __syncwarp();
      ++job;
    };
    // one 64-wide k-block: four UMMAs of K = 16 (32 bytes further inside the swizzle row each)
    auto mma_block = [&](uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, bool first, uint32_t release_bar, uint32_t release_bar2) {
      __syncwarp();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_bf16_lo(d, a_lo + 2u * k, b_lo + 2u * k, idesc, (!first || k > 0) ? 1u : 0u);
        tc_commit(release_bar);
        if (release_bar2) tc_commit(release_bar2);
      }
      __syncwarp();
    };
    // layer-0 style job: A = X panel, B = num_kb0 ring blocks
    auto x_job = [&](uint32_t idesc, bool last_x_use) {
      const uint32_t d = begin_job();
      for (int kb = 0; kb < nkb; ++kb) {
        const int xb = kbo(kb);
        mbar_wait(xfull0 + 8 * xb, xph);
        { CH_T0(); mbar_wait(wfull0 + 8 * ws, wph); if (lane == 0) CH_ADD(2); }
        tc_fence_after();
        mma_block(d, x_lo0 + (uint32_t)xb * (CH_BLK >> 4), w_lo0 + (uint32_t)ws * (CH_BLK >> 4), idesc, kb == 0, wempty0 + 8 * ws,
                  (last_x_use && kb == nkb - 1) ? xempty : 0u);
        if (++ws == CH_W_SLOTS) { ws = 0; wph ^= 1; }
      }
      end_job();
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      if (p.n_g > 0) x_job(idesc_g, false);
      for (int u = 0; u < nU; ++u) x_job(idesc_m1, nE == 1 && u == nU - 1);
      for (int e = 0; e < nE; ++e) {
        if (e + 1 < nE) x_job(idesc_m1, e + 2 == nE && nU == 1);
        {                                                          // M2(e): A = the activation blocks the epilogue wrote
          const uint32_t d = begin_job();
          for (int kb = 0; kb < KB1; ++kb) {
            { CH_T0(); mbar_wait(afull0 + 8 * kb, ae & 1u); if (lane == 0) CH_ADD(3); }
            { CH_T0(); mbar_wait(wfull0 + 8 * ws, wph); if (lane == 0) CH_ADD(2); }
            tc_fence_after();
            mma_block(d, a_lo0 + (uint32_t)kb * (CH_BLK >> 4), w_lo0 + (uint32_t)ws * (CH_BLK >> 4), idesc_m2, kb == 0, wempty0 + 8 * ws,
                      aempty0 + 8 * kb);
            if (++ws == CH_W_SLOTS) { ws = 0; wph ^= 1; }
          }
          end_job();
          ++ae;
        }
        if (e + 1 < nE)
          for (int u = 1; u < nU; ++u) x_job(idesc_m1, e + 2 == nE && u == nU - 1);
      }
      xph ^= 1;
    }
    if (CH_PROF(p.prof) && lane == 0) atomicAdd(p.prof + 8, (unsigned long long)(clock64() - mma_t0));
  } else {
    // =============================== epilogue (warps 2..17) ===============================
    const int ew = warp - 2;
    const int q = warp & 3;                                        // TMEM lane quarter
    const int sub = ew >> 2;                                       // 32-column chunk of a 128-column unit
    const int pj = sub >> 1, half = sub & 1;                       // block pair within the unit, 64-byte half of the block's rows
    const int bar_id = 2 + 2 * q + pj;                             // the two warps (q, 2pj), (q, 2pj+1) share one 32-row x 128-byte region
    const bool issuer = half == 0;
    float* bias_t = bias_s + ew * 32;
    const bool has_drop = p.drop_p > 0.f;
    const float keep = has_drop ? 1.f / (1.f - p.drop_p) : 1.f;
    const uint32_t thr32 = drop_thr32(p.drop_p);
    const uint32_t s00 = has_drop ? drop_s0(*p.seed_dev, p.salt0) : 0u, s01 = has_drop ? drop_s0(*p.seed_dev, p.salt1) : 0u;
    uint32_t ejob = 0, ae = 0;
    uint32_t slot = 0;
    bool store_pending = false;
    const long long epi_t0 = CH_PROF(p.prof) ? clock64() : 0;
    auto wait_job = [&]() -> uint32_t {
      slot = ejob & 3u;
      { CH_T0(); mbar_wait(tfull0 + 8 * slot, (ejob >> 2) & 1u); if (ew == 0 && lane == 0) CH_ADD(4); }
      tc_fence_after();
      return tmem_base + ((uint32_t)(q * 32) << 16) + slot * 128u;
    };
    auto done_job = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + 8 * slot);
      ++ejob;
    };
    // bias + ReLU + dropout of 32 accumulator columns -> 16 packed bf16 pairs
    auto act32 = [&](const uint32_t (&v)[32], const float* __restrict__ bias_g, uint32_t s0, uint32_t row, uint32_t gcol, uint32_t (&o)[16]) {
      __syncwarp();
      bias_t[lane] = __ldg(bias_g + lane) * keep;
      __syncwarp();
      float f[32];
      const float4* b4 = reinterpret_cast<const float4*>(bias_t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b = b4[j];
        f[4 * j] = fmaxf(fmaf(__uint_as_float(v[4 * j]), keep, b.x), 0.f);
        f[4 * j + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j + 1]), keep, b.y), 0.f);
        f[4 * j + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j + 2]), keep, b.z), 0.f);
        f[4 * j + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j + 3]), keep, b.w), 0.f);
      }
      if (has_drop) {
        uint32_t x = drop_lcg_seed(s0, row, gcol >> 5);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          x = x * kLcgA + kLcgC;
          f[j] = x >= thr32 ? f[j] : 0.f;
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) o[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
    };
    // the pair's region of activation block `blk`: rows q*32.., this warp's 64-byte half of every 128-byte row, 128-byte swizzle
    auto write_block = [&](int blk, const uint32_t (&o)[16], const CUtensorMap* map, int32_t gcol, int32_t grow, uint32_t signal) {
      uint8_t* region = ar + blk * CH_BLK + q * 4096;
      if (issuer && lane == 0 && store_pending) { tma_store_wait_read0(); }
      store_pending = false;
      pair_bar_sync(bar_id);                                       // the region's previous TMA store has been read
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<uint4*>(region + lane * 128 + (((4 * half + j) ^ (lane & 7)) * 16)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      fence_async_smem();
      pair_bar_sync(bar_id);                                       // both halves are in place (and visible to the async proxy)
      if (issuer && lane == 0) {
        if (map) { tma_store_2d(map, smem_u32(region), gcol, grow); tma_store_commit(); }
        if (signal) mbar_arrive(signal);
      }
      if (issuer && map) store_pending = true;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t m0 = tile * 128;
      const int64_t m = m0 + q * 32 + lane;
      const int32_t grow = (int32_t)(m0 + q * 32);
      if (p.n_g > 0) {                                             // gate / wide-linear logits: fp32, a few columns
        const uint32_t trow = wait_job();
        if (32 * sub < p.n_g16) {
          uint32_t v[32];
          tc_ld32_nowait(trow + 32 * sub, v);
          tc_ld_wait();
          if (m < p.B) {
            const float* bg = p.b0 + (int64_t)nE * p.d0 + 32 * sub;
            float* lg = p.Lg + m * p.ldg + 32 * sub;
            const int nv = min(32, p.n_g - 32 * sub);
#pragma unroll
            for (int j = 0; j < 32; ++j) if (j < nv) lg[j] = __uint_as_float(v[j]) + __ldg(bg + j);
          }
        }
        done_job();
      }
      // E1: layer-0 unit u of expert e -> activation blocks 2u, 2u+1
      auto e1 = [&](int e, int u, uint32_t gi) {                  // gi: running index of expert e over all tiles of this CTA
        const uint32_t trow = wait_job();
        uint32_t v[32];
        tc_ld32_nowait(trow + 32 * sub, v);
        const int kb = 2 * u + pj;
        const int gcol = e * p.d0 + 128 * u + 32 * sub;
        tc_ld_wait();
        uint32_t o[16];
        act32(v, p.b0 + gcol, s00, (uint32_t)m, (uint32_t)gcol, o);
        { CH_T0(); mbar_wait(aempty0 + 8 * kb, (gi & 1u) ^ 1u); if (ew == 0 && lane == 0) CH_ADD(5); }   // the previous expert's layer-1 MMAs have read block kb
        write_block(kb, o, p.store_a0 ? &map_a0 : nullptr, (int32_t)(e * p.d0 + 64 * kb), grow, afull0 + 8 * kb);
        done_job();
      };
      auto e2 = [&](int e) {
        const uint32_t trow = wait_job();                          // Z1 complete: every layer-1 MMA of e has retired, blocks 2, 3 are free
        if (32 * sub < p.d1) {
          uint32_t v[32];
          tc_ld32_nowait(trow + 32 * sub, v);
          const int gcol = e * p.d1 + 32 * sub;
          tc_ld_wait();
          uint32_t o[16];
          act32(v, p.b1 + gcol, s01, (uint32_t)m, (uint32_t)gcol, o);
          write_block(2 + pj, o, &map_h, (int32_t)(e * p.d1 + 64 * pj), grow, 0u);
        }
        done_job();
      };
      for (int u = 0; u < nU; ++u) e1(pe(0), u, ae);
      for (int e = 0; e < nE; ++e) {
        if (e + 1 < nE) e1(pe(e + 1), 0, ae + 1);
        e2(pe(e));
        ++ae;
        if (e + 1 < nE)
          for (int u = 1; u < nU; ++u) e1(pe(e + 1), u, ae);
      }
    }
    if (issuer && lane == 0) tma_store_wait_all();
    if (CH_PROF(p.prof) && ew == 0 && lane == 0) { atomicAdd(p.prof + 6, (unsigned long long)(clock64() - epi_t0)); atomicAdd(p.prof + 7, 1ull); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
  }
}

}  // namespace cdcmdr

using namespace cdcmdr;

static std::atomic<unsigned long long*> g_chain_prof{nullptr};
extern "C" int cdcmdr_ple_chain_profile(uint64_t* counters16) {
  g_chain_prof.store(reinterpret_cast<unsigned long long*>(counters16), std::memory_order_relaxed);
  return 0;
}

extern "C" int cdcmdr_ple_chain_ok(int32_t K0, int32_t d0, int32_t d1, int32_t n_g) {
  return (K0 >= 8 && K0 % 8 == 0 && ceil_div(K0, 64) <= CH_MAX_KB0 && (d0 == 128 || d0 == 256) && (d1 == 64 || d1 == 128) && n_g >= 0 &&
          n_g <= 128) ? 1 : 0;
}

extern "C" int cdcmdr_ple_chain_fwd(const cdcmdr_ple_chain_t* c, cdcmdr_stream_t s) {
  CDC_REQUIRE(c && c->X && c->W0 && c->b0 && c->W1 && c->b1 && c->H, "chain: null operand");
  CDC_REQUIRE(cdcmdr_ple_chain_ok(c->K0, c->d0, c->d1, c->n_g), "chain: unsupported geometry (cdcmdr_ple_chain_ok)");
  CDC_REQUIRE(c->nE >= 1 && c->B >= 0 && c->B < ((int64_t)1 << 31), "chain: bad sizes");
  if (c->B == 0) return 0;
  CDC_REQUIRE(c->ldx % 8 == 0 && c->ldh % 8 == 0 && ((uintptr_t)c->X % 16) == 0 && ((uintptr_t)c->W0 % 16) == 0 && ((uintptr_t)c->W1 % 16) == 0 &&
              ((uintptr_t)c->H % 16) == 0 && (!c->A0 || (c->lda0 % 8 == 0 && ((uintptr_t)c->A0 % 16) == 0)),
              "chain: TMA needs 16-byte aligned bases and pitches");
  CDC_REQUIRE(c->n_g == 0 || c->Lg, "chain: logits output missing");
  CDC_REQUIRE(c->drop_p <= 0.f || c->seed_dev, "chain: dropout needs a device seed");
  ChainParams q{};
  q.B = c->B; q.num_kb0 = (int)ceil_div(c->K0, 64); q.nE = c->nE; q.d0 = c->d0; q.d1 = c->d1; q.n_g = c->n_g;
  q.n_g16 = (int)(ceil_div(c->n_g > 0 ? c->n_g : 1, 16) * 16);
  q.nU = c->d0 / 128; q.KB1 = c->d0 / 64;
  q.b0 = c->b0; q.b1 = c->b1; q.Lg = c->Lg; q.ldg = c->ldg; q.store_a0 = c->A0 ? 1 : 0;
  q.drop_p = c->drop_p; q.seed_dev = c->seed_dev; q.salt0 = c->salt0; q.salt1 = c->salt1;
  q.prof = g_chain_prof.load(std::memory_order_relaxed);
  const int64_t w0_rows = (int64_t)c->nE * c->d0 + c->n_g;
  CUtensorMap mx, mw0, mwg, mw1, ma0, mh;
  if (int rc = make_map_sw128(&mx, c->X, c->B, c->K0, c->ldx, 128u)) return rc;
  if (int rc = make_map_sw128(&mw0, c->W0, w0_rows, c->K0, c->K0, 128u)) return rc;
  if (int rc = make_map_sw128(&mwg, c->W0, w0_rows, c->K0, c->K0, (uint32_t)q.n_g16)) return rc;
  if (int rc = make_map_sw128(&mw1, c->W1, (int64_t)c->nE * c->d1, c->d0, c->d0, (uint32_t)c->d1)) return rc;
  if (int rc = make_map_sw128(&mh, c->H, c->B, (int64_t)c->nE * c->d1, c->ldh, 32u)) return rc;
  ma0 = mh;
  if (c->A0) { if (int rc = make_map_sw128(&ma0, c->A0, c->B, (int64_t)c->nE * c->d0, c->lda0, 32u)) return rc; }
  const int smem_bytes = (q.num_kb0 + CH_W_SLOTS + CH_A_SLOTS) * CH_BLK + CH_EPI_WARPS * 32 * 4 + 512;
  CDC_REQUIRE(smem_bytes <= CH_SMEM_LIMIT, "chain: shared memory budget exceeded");
  static bool attr_set = false;
  if (!attr_set) {
    CDC_CHECK(cudaFuncSetAttribute(ple_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_LIMIT));
    attr_set = true;
  }
  const int64_t tiles = ceil_div(c->B, 128);
  const int grid = (int)(tiles < kNumSMs ? tiles : kNumSMs);
  ple_chain_fwd_kernel<<<grid, CH_THREADS, smem_bytes, to_stream(s)>>>(mx, mw0, mwg, mw1, ma0, mh, q);
  CDC_LAUNCHED();
  return 0;
}
