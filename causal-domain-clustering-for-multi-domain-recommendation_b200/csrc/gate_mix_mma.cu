// Gate softmax + expert-weighted sum (CGC.forward / MMoE.forward, ple.py:106-123, mmoe.py:56-60) and its backward for bf16
// activations, one warp per batch row, with the per-row products on mma.sync:
//
//   forward    out[j, :]  = sum_e P[j, e] * H[e, :]                  [gates x experts] . [experts x h]
//   backward   S[j, e]    = <dOut[j, :], H[e, :]>                    [gates x h] . [h x experts]   (exact: bf16 x bf16 -> fp32)
//              dz         = softmax backward of S against P          (pair lanes)
//              dH[e, :]   = (sum_j P[j, e] * dOut[j, :]) * relu'(H)  [experts x gates] . [gates x h]
//
// Every row has its own operands, so there is no operand to share across rows: tcgen05 (128-row tiles against one B) does not
// apply; the warp-level m16n8k16 / m16n8k8 shapes hold one row's [<=8 gates] x [<=16 experts] problem in one fragment.  The
// fp32 probabilities enter as a bf16 hi + lo pair (two MMAs, ~2^-17 relative), the activations are bf16 already, all
// accumulation is fp32.  ncu on the SIMT kernels this replaces (profiles/r1_ncu_gate_mix.md): 2 405 / 1 272 warp instructions
// per row (bwd / fwd), issue slots 69-77 % busy - instruction-bound at ~3x their HBM time.  Here a row costs ~400 / ~200.
//
// Rows are staged with cp.async (16 B per lane, double-buffered per warp: row i+1 streams in while row i is mixed) into a
// 16-byte-chunk XOR swizzle (chunk ^ (row & 7)), so that ldmatrix over 8 expert rows of a 128/256-byte pitch is conflict-free
// without padding; outputs leave through the same swizzle with 128-bit coalesced stores.
#include "common.cuh"

namespace cdcmdr {

struct MixM {
  int n_gates, n_experts, max_sel, n_pairs;
  const int32_t* gate_col; const int32_t* gate_n; const int32_t* gate_sel;
};

namespace {

constexpr int kMmaMaxPairs = 32;
constexpr int kMmaMaxNp = 64;

__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
// D(16x8, fp32) += A(16x16, bf16, row) * B(16x8, bf16, col)
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// D(16x8, fp32) += A(16x8, bf16, row) * B(8x8, bf16, col)
__device__ __forceinline__ void mma_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(b0));
}
// fp32 pair -> bf16x2 hi + bf16x2 lo (hi + lo carries ~16 mantissa bits)
__device__ __forceinline__ void split_hi_lo(float p0, float p1, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(p0, p1);
  lo = pack_bf16x2(p0 - __uint_as_float(hi << 16), p1 - __uint_as_float(hi & 0xffff0000u));
}

struct PairTables {
  int gate[kMmaMaxPairs], e[kMmaMaxPairs], slot[kMmaMaxPairs], col[kMmaMaxPairs], first[kMmaMaxPairs], cnt[kMmaMaxPairs];
  int n_pairs, max_cnt;
};

// thread 0 of the CTA: the compact gate-major list of (gate, expert) pairs
__device__ __forceinline__ void build_tables(const MixM& d, PairTables& t) {
  int n = 0, mc = 0;
  uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < d.n_gates; ++j) {
    const int first = n, cnt = d.gate_n[j], col = d.gate_col[j];
    mc = cnt > mc ? cnt : mc;
    for (int s = 0; s < cnt; ++s, ++n) {
      const int e = d.gate_sel[j * d.max_sel + s];
      if (n >= kMmaMaxPairs || e < 0 || e >= 16 || (seen[j] >> e & 1u)) __trap();     // launcher contract violated
      seen[j] |= 1u << e;
      t.gate[n] = j; t.e[n] = e; t.slot[n] = j * d.max_sel + s; t.col[n] = col + s; t.first[n] = first; t.cnt[n] = cnt;
    }
  }
  t.n_pairs = n; t.max_cnt = mc;
}

// Swizzled row buffers.  A buffer holds `rows` blocks of HH bf16 (ROWB = 2*HH bytes each, buffer base aligned to ROWB); 16-byte
// chunk c of block r lives at chunk (c ^ (r & 7)).  Because the chunk bits of a block's base address are zero, the swizzle is an
// XOR on the ADDRESS: every unrolled access below is `lane_base ^ compile-time constant` - one LOP3, no per-access index math.
//
// Lane l moves chunk (l % CPR) of block (l / CPR) + RPI*it in iteration it (RPI = 32 / CPR blocks per warp-wide pass; 512
// contiguous global bytes per pass).  (r & 7) = (l / CPR) ^ ((RPI*it) & 7) since the two never share a bit.
template <int HH> struct Swz {
  static constexpr int CPR = HH / 8, ROWB = HH * 2, RPI = 32 / CPR, MAXIT = 16 / RPI;
  static __device__ __forceinline__ uint32_t lane_off(int lane) {          // offset of (block l/CPR, chunk l%CPR) inside the buffer
    const int r0 = lane / CPR, c = lane % CPR;
    return (uint32_t)(r0 * ROWB + ((c ^ r0) << 4));
  }
  static __device__ __forceinline__ uint32_t at(uint32_t lane_base, int it) {
    return (lane_base + (uint32_t)(it * 512)) ^ (uint32_t)(((RPI * it) & 7) << 4);
  }
};

template <int HH>
__device__ __forceinline__ void stage_rows(uint32_t buf, const uint16_t* src, int rows, int lane) {
  using S = Swz<HH>;
  const uint32_t lb = buf + S::lane_off(lane);
  const uint8_t* s8 = reinterpret_cast<const uint8_t*>(src) + lane * 16;
  const int r0 = lane / S::CPR;
#pragma unroll
  for (int it = 0; it < S::MAXIT; ++it)
    if (r0 + S::RPI * it < rows) cp_async16(S::at(lb, it), s8 + it * 512);
}

template <int HH>
__device__ __forceinline__ void unstage_rows(uint8_t* dst, uint32_t buf, int rows, int lane) {
  using S = Swz<HH>;
  const uint32_t lb = buf + S::lane_off(lane);
  uint8_t* d8 = dst + lane * 16;
  const int r0 = lane / S::CPR;
#pragma unroll
  for (int it = 0; it < S::MAXIT; ++it)
    if (r0 + S::RPI * it < rows) {
      uint4 v;
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(S::at(lb, it)) : "memory");
      *reinterpret_cast<uint4*>(d8 + it * 512) = v;
    }
}

__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// per-warp shared memory: [row buffers ...][floats ...]; the CTA's region starts with a 256-byte zero block that stands in for
// gates / experts beyond the descriptor's (ldmatrix reads zeros there, at any XOR-ed chunk offset)
constexpr int kZeroBytes = 256;

// ------------------------------------------------------------------------------------------------ forward
template <int HH>
__global__ void __launch_bounds__(256, 3)
gate_mix_fwd_mma_kernel(MixM d, const uint16_t* __restrict__ H, int64_t ldh, const float* __restrict__ logits, int64_t ldl,
                        uint16_t* __restrict__ out, int64_t ldo, float* __restrict__ probs, int64_t B, int per_warp) {
  extern __shared__ __align__(128) uint8_t dyn_raw[];
  __shared__ PairTables T;
  constexpr int ROWB = HH * 2;
  uint8_t* dyn = dyn_raw + ((256u - (sm_u32(dyn_raw) & 255u)) & 255u);      // 256-byte aligned (the launcher adds the slack)
  if (threadIdx.x == 0) build_tables(d, T);
  if (threadIdx.x < kZeroBytes / 4) reinterpret_cast<uint32_t*>(dyn)[threadIdx.x] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nE = d.n_experts, ng = d.n_gates, np = ng * d.max_sel;
  const int hb = nE * ROWB, ob = ng * ROWB;
  uint8_t* base = dyn + kZeroBytes + (size_t)warp * per_warp;
  const uint32_t hbuf0 = sm_u32(base), ostu = hbuf0 + 2 * hb, zrow = sm_u32(dyn);
  float* sprob = reinterpret_cast<float*>(base + 2 * hb + ob);     // [kMmaMaxNp] probabilities by slot (padding slots stay 0)
  float* Pm = sprob + kMmaMaxNp;                         // [8][16]  P[j][e]
  float* lx = Pm + 128;                                  // [2][32] this row's / the next row's logits (cp.async)
  for (int i = lane; i < 128; i += 32) Pm[i] = 0.f;
  for (int i = lane; i < kMmaMaxNp; i += 32) sprob[i] = 0.f;
  lx[lane] = 0.f; lx[32 + lane] = 0.f;
  __syncthreads();
  const int n_pairs = T.n_pairs, max_cnt = T.max_cnt;
  const bool is_pair = lane < n_pairs;
  const int pj = is_pair ? T.gate[lane] : 0, pe = is_pair ? T.e[lane] : 0, pslot = is_pair ? T.slot[lane] : 0;
  const int pcol = is_pair ? T.col[lane] : 0, pfirst = is_pair ? T.first[lane] : 0, pcnt = is_pair ? T.cnt[lane] : 0;
  const int j0 = lane >> 2, k0 = (lane & 3) * 2;
  // ldmatrix.x4.trans row of this lane: matrix m = lane/8 holds experts (m&1)*8 .. +7 of chunk 2g + (m>>1)
  const int le = ((lane >> 3) & 1) * 8 + (lane & 7), lch = lane >> 4;
  const bool le_ok = le < nE;
  const uint32_t h_off = (uint32_t)(le * ROWB + ((lch ^ (le & 7)) << 4));            // chunk 2g+lch == (2g) ^ lch
  const uint32_t o_base = ostu + (uint32_t)(j0 * ROWB + ((j0 & 7) << 4) + (lane & 3) * 4);
  const bool j_ok = j0 < ng;

  const int64_t stride = (int64_t)gridDim.x * 8;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  // no thread touches a buffer before its own cp.async group has landed AND the warp has synchronised
  if (row < B) {
    stage_rows<HH>(hbuf0, H + row * ldh, nE, lane);
    if (is_pair) cp_async4(sm_u32(lx + lane), logits + row * ldl + pcol);
  }
  cp_async_commit();
  int b = 0;
  for (; row < B; row += stride, b ^= 1) {
    const int64_t next = row + stride;
    if (next < B) {
      stage_rows<HH>(hbuf0 + (b ^ 1) * hb, H + next * ldh, nE, lane);
      if (is_pair) cp_async4(sm_u32(lx + (b ^ 1) * 32 + lane), logits + next * ldl + pcol);
    }
    cp_async_commit();
    cp_async_wait1();                                    // this row (H and logits) has landed; the next one streams in behind it
    // ---- softmax over each gate's logits, one (gate, expert) pair per lane; a gate's pairs sit in consecutive lanes
    const float x = lx[b * 32 + lane];
    float mx = -INFINITY;
    for (int i = 0; i < max_cnt; ++i) { const float v = __shfl_sync(0xffffffffu, x, pfirst + i); if (i < pcnt) mx = fmaxf(mx, v); }
    const float ex = is_pair ? expf(x - mx) : 0.f;
    float sum = 0.f;
    for (int i = 0; i < max_cnt; ++i) { const float v = __shfl_sync(0xffffffffu, ex, pfirst + i); if (i < pcnt) sum += v; }
    if (is_pair) { const float p = ex / sum; sprob[pslot] = p; Pm[pj * 16 + pe] = p; }
    __syncwarp();
    for (int k = lane; k < np; k += 32) probs[row * np + k] = sprob[k];
    // A fragments (gates x experts): a0 = P[j0][k0, k0+1], a2 = P[j0][k0+8, k0+9]; rows 8..15 are zero
    uint32_t ah0, al0, ah2, al2;
    {
      const float2 p0 = *reinterpret_cast<const float2*>(Pm + j0 * 16 + k0);
      const float2 p1 = *reinterpret_cast<const float2*>(Pm + j0 * 16 + k0 + 8);
      split_hi_lo(p0.x, p0.y, ah0, al0);
      split_hi_lo(p1.x, p1.y, ah2, al2);
    }
    const uint32_t h_base = le_ok ? hbuf0 + b * hb + h_off : zrow;
#pragma unroll
    for (int g = 0; g < HH / 16; ++g) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(h_base ^ (uint32_t)((2 * g) << 4), b0, b1, b2, b3);
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(d0, ah0, 0u, ah2, 0u, b0, b1);
      mma_16816(d0, al0, 0u, al2, 0u, b0, b1);
      mma_16816(d1, ah0, 0u, ah2, 0u, b2, b3);
      mma_16816(d1, al0, 0u, al2, 0u, b2, b3);
      if (j_ok) {
        sts32(o_base ^ (uint32_t)((2 * g) << 4), pack_bf16x2(d0[0], d0[1]));
        sts32(o_base ^ (uint32_t)((2 * g + 1) << 4), pack_bf16x2(d1[0], d1[1]));
      }
    }
    __syncwarp();
    unstage_rows<HH>(reinterpret_cast<uint8_t*>(out + row * ldo), ostu, ng, lane);
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int HH, bool RELU>
__global__ void __launch_bounds__(256, 3)
gate_mix_bwd_mma_kernel(MixM d, const uint16_t* __restrict__ H, int64_t ldh, const float* __restrict__ probs,
                        const uint16_t* __restrict__ dOut, int64_t ldo, uint16_t* __restrict__ dH, int64_t lddh, float relu_scale,
                        float* __restrict__ dlogits, int64_t lddl, int64_t B, int per_warp) {
  extern __shared__ __align__(128) uint8_t dyn_raw[];
  __shared__ PairTables T;
  constexpr int ROWB = HH * 2;
  uint8_t* dyn = dyn_raw + ((256u - (sm_u32(dyn_raw) & 255u)) & 255u);
  if (threadIdx.x == 0) build_tables(d, T);
  if (threadIdx.x < kZeroBytes / 4) reinterpret_cast<uint32_t*>(dyn)[threadIdx.x] = 0u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nE = d.n_experts, ng = d.n_gates, np = ng * d.max_sel;
  const int hb = nE * ROWB, ob = ng * ROWB;
  uint8_t* base = dyn + kZeroBytes + (size_t)warp * per_warp;
  const uint32_t buf0 = sm_u32(base), zrow = sm_u32(dyn);           // each of the two buffers: H row | dOut row
  float* sprob2 = reinterpret_cast<float*>(base + 2 * (hb + ob));  // [2][kMmaMaxNp] this row's / the next row's probabilities
  float* Pm = sprob2 + 2 * kMmaMaxNp;                    // [8][16]  P[j][e] (x relu_scale)
  float* Ss = Pm + 128;                                  // [8][16]  S[j][e] = <dOut_j, H_e>
  for (int i = lane; i < 128; i += 32) Pm[i] = 0.f;
  __syncthreads();
  const int n_pairs = T.n_pairs, max_cnt = T.max_cnt;
  const bool is_pair = lane < n_pairs;
  const int pj = is_pair ? T.gate[lane] : 0, pe = is_pair ? T.e[lane] : 0, pslot = is_pair ? T.slot[lane] : 0;
  const int pcol = is_pair ? T.col[lane] : 0, pfirst = is_pair ? T.first[lane] : 0, pcnt = is_pair ? T.cnt[lane] : 0;
  const int e0 = lane >> 2, k0 = (lane & 3) * 2;
  // S: A = dOut via ldmatrix.x2 (lanes 0..15: gate lane%8, chunk 2ks + (lane/8)&1); B = H via ldmatrix.x4 (expert
  // (lane/16)*8 + lane%8, chunk 2ks + (lane/8)&1)
  const int aj = lane & 7, ach = (lane >> 3) & 1;
  const bool aj_ok = aj < ng;
  const uint32_t a_off = (uint32_t)(hb + aj * ROWB + ((ach ^ aj) << 4));
  const int be = (lane >> 4) * 8 + (lane & 7);
  const bool be_ok = be < nE;
  const uint32_t b_off = (uint32_t)(be * ROWB + ((ach ^ (be & 7)) << 4));
  // dH: B = dOut via ldmatrix.x4.trans (gate lane%8, chunk 4g + lane/8)
  const int tch = lane >> 3;
  const uint32_t t_off = (uint32_t)(hb + aj * ROWB + ((tch ^ aj) << 4));
  const bool has_hi = nE > 8;                            // experts 8..15 exist
  const bool lo_ok = e0 < nE, hi_ok = e0 + 8 < nE;
  const uint32_t h_off = (uint32_t)(e0 * ROWB + ((e0 & 7) << 4) + (lane & 3) * 4);   // this lane's (expert e0, columns k0..k0+1) of chunk 0
  const float pscale = RELU ? relu_scale : 1.f;

  const int64_t stride = (int64_t)gridDim.x * 8;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row < B) {
    stage_rows<HH>(buf0, H + row * ldh, nE, lane);
    stage_rows<HH>(buf0 + hb, dOut + row * ldo, ng, lane);
    for (int k = lane; k < np; k += 32) cp_async4(sm_u32(sprob2 + k), probs + row * np + k);
  }
  cp_async_commit();
  int b = 0;
  for (; row < B; row += stride, b ^= 1) {
    const int64_t next = row + stride;
    if (next < B) {
      const uint32_t nb = buf0 + (b ^ 1) * (hb + ob);
      stage_rows<HH>(nb, H + next * ldh, nE, lane);
      stage_rows<HH>(nb + hb, dOut + next * ldo, ng, lane);
      for (int k = lane; k < np; k += 32) cp_async4(sm_u32(sprob2 + (b ^ 1) * kMmaMaxNp + k), probs + next * np + k);
    }
    cp_async_commit();
    const float* sprob = sprob2 + b * kMmaMaxNp;
    cp_async_wait1();
    __syncwarp();
    const uint32_t hs = buf0 + b * (hb + ob);
    // ---- S = dOut . H^T
    {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t a_base = aj_ok ? hs + a_off : zrow, b_base = be_ok ? hs + b_off : zrow;
#pragma unroll
      for (int ks = 0; ks < HH / 16; ++ks) {
        uint32_t a0, a2, b00, b01, b10, b11;
        ldsm_x2(a_base ^ (uint32_t)((2 * ks) << 4), a0, a2);
        ldsm_x4(b_base ^ (uint32_t)((2 * ks) << 4), b00, b01, b10, b11);
        mma_16816(s0, a0, 0u, a2, 0u, b00, b01);
        if (has_hi) mma_16816(s1, a0, 0u, a2, 0u, b10, b11);
      }
      // rows 0..7 of the 16 x 16 result: S[j = lane/4][e = k0, k0+1 (+8)]
      *reinterpret_cast<float2*>(Ss + e0 * 16 + k0) = make_float2(s0[0], s0[1]);
      *reinterpret_cast<float2*>(Ss + e0 * 16 + 8 + k0) = make_float2(s1[0], s1[1]);
    }
    __syncwarp();
    // ---- softmax backward per (gate, expert) pair: dz = p * (dp - <p, dp>); a gate's pairs sit in consecutive lanes
    {
      const float pr = is_pair ? sprob[pslot] : 0.f;
      const float dp = is_pair ? Ss[pj * 16 + pe] : 0.f;
      const float pd = pr * dp;
      float dot = 0.f;
      for (int i = 0; i < max_cnt; ++i) { const float v = __shfl_sync(0xffffffffu, pd, pfirst + i); if (i < pcnt) dot += v; }
      if (is_pair) {
        dlogits[row * lddl + pcol] = pr * (dp - dot);
        Pm[pj * 16 + pe] = pr * pscale;
      }
    }
    __syncwarp();
    // ---- dH = (relu_scale * P)^T . dOut:  A[m = e][k = j] = P[j][e]
    uint32_t ah0, al0, ah1, al1;
    split_hi_lo(Pm[k0 * 16 + e0], Pm[(k0 + 1) * 16 + e0], ah0, al0);
    split_hi_lo(Pm[k0 * 16 + e0 + 8], Pm[(k0 + 1) * 16 + e0 + 8], ah1, al1);
    const uint32_t t_base = aj_ok ? hs + t_off : zrow;
    const uint32_t h_base = hs + h_off;
#pragma unroll
    for (int g = 0; g < HH / 32; ++g) {
      uint32_t bb[4];
      ldsm_x4_t(t_base ^ (uint32_t)((4 * g) << 4), bb[0], bb[1], bb[2], bb[3]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float dd[4] = {0.f, 0.f, 0.f, 0.f};
        mma_1688(dd, ah0, ah1, bb[t]);
        mma_1688(dd, al0, al1, bb[t]);
        const uint32_t a_lo = h_base ^ (uint32_t)((4 * g + t) << 4), a_hi = a_lo + 8 * ROWB;
        if (RELU) {
          // bf16 > 0: low half as a signed value after << 16, high half: the word itself >= 0x10000 as signed
          const uint32_t hl = lo_ok ? lds32(a_lo) : 0u, hh = hi_ok ? lds32(a_hi) : 0u;
          dd[0] = (int32_t)(hl << 16) > 0 ? dd[0] : 0.f; dd[1] = (int32_t)hl > 0xFFFF ? dd[1] : 0.f;
          dd[2] = (int32_t)(hh << 16) > 0 ? dd[2] : 0.f; dd[3] = (int32_t)hh > 0xFFFF ? dd[3] : 0.f;
        }
        if (lo_ok) sts32(a_lo, pack_bf16x2(dd[0], dd[1]));
        if (hi_ok) sts32(a_hi, pack_bf16x2(dd[2], dd[3]));
      }
    }
    __syncwarp();
    unstage_rows<HH>(reinterpret_cast<uint8_t*>(dH + row * lddh), hs, nE, lane);
    __syncwarp();
  }
}

constexpr int kMmaSmemPerSM = 227 * 1024;
constexpr int kMmaSmemPerCta = 100 * 1024;

bool mma_shape_ok(const cdcmdr_mix_desc_t* d, int is_bf16) {
  return is_bf16 && (d->h == 64 || d->h == 128) && d->n_gates <= 8 && d->n_experts <= 16 && d->n_pairs > 0 && d->n_pairs <= kMmaMaxPairs &&
         d->n_gates * d->max_sel <= kMmaMaxNp;
}
bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

int pick_grid(int64_t B, size_t smem) {
  int per_sm = (int)((kMmaSmemPerSM + 1024) / (smem + 1024));
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  int64_t g = ceil_div(B, 8);
  const int64_t cap = (int64_t)kNumSMs * per_sm;
  return (int)(g > cap ? cap : (g < 1 ? 1 : g));
}

}  // namespace

// Both return -1 when the shape / layout is outside what the MMA kernels cover (the caller falls through to the SIMT kernels).
int gate_mix_fwd_mma(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* logits, int64_t ldl, void* out, int64_t ldo,
                     float* probs, int64_t B, int is_bf16, cudaStream_t st) {
  if (!mma_shape_ok(d, is_bf16) || !al16(H) || !al16(out) || ldh % 8 || ldo % 8) return -1;
  const size_t per_warp = (size_t)2 * d->n_experts * d->h * 2 + (size_t)d->n_gates * d->h * 2 + (kMmaMaxNp + 128 + 64) * sizeof(float);
  const size_t pw = (per_warp + 255) & ~(size_t)255;
  const size_t smem = 8 * pw + kZeroBytes + 256;
  if (smem > (size_t)kMmaSmemPerCta) return -1;
  MixM k{d->n_gates, d->n_experts, d->max_sel, d->n_pairs, d->gate_col, d->gate_n, d->gate_sel};
  const int grid = pick_grid(B, smem);
  if (d->h == 128) {
    CDC_CHECK(cudaFuncSetAttribute(gate_mix_fwd_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemPerCta));
    gate_mix_fwd_mma_kernel<128><<<grid, 256, smem, st>>>(k, (const uint16_t*)H, ldh, logits, ldl, (uint16_t*)out, ldo, probs, B, (int)pw);
  } else {
    CDC_CHECK(cudaFuncSetAttribute(gate_mix_fwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemPerCta));
    gate_mix_fwd_mma_kernel<64><<<grid, 256, smem, st>>>(k, (const uint16_t*)H, ldh, logits, ldl, (uint16_t*)out, ldo, probs, B, (int)pw);
  }
  CDC_LAUNCHED();
  return 0;
}

int gate_mix_bwd_mma(const cdcmdr_mix_desc_t* d, const void* H, int64_t ldh, const float* probs, const void* dOut, int64_t ldo, void* dH,
                     int64_t lddh, float relu_scale, float* dlogits, int64_t lddl, int64_t B, int is_bf16, cudaStream_t st) {
  if (!mma_shape_ok(d, is_bf16) || !al16(H) || !al16(dOut) || !al16(dH) || ldh % 8 || ldo % 8 || lddh % 8) return -1;
  const size_t per_warp = (size_t)2 * (d->n_experts + d->n_gates) * d->h * 2 + (2 * kMmaMaxNp + 128 + 128) * sizeof(float);
  const size_t pw = (per_warp + 255) & ~(size_t)255;
  const size_t smem = 8 * pw + kZeroBytes + 256;
  if (smem > (size_t)kMmaSmemPerCta) return -1;
  MixM k{d->n_gates, d->n_experts, d->max_sel, d->n_pairs, d->gate_col, d->gate_n, d->gate_sel};
  const int grid = pick_grid(B, smem);
#define MIXB(HH, R) do { \
    CDC_CHECK(cudaFuncSetAttribute(gate_mix_bwd_mma_kernel<HH, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemPerCta)); \
    gate_mix_bwd_mma_kernel<HH, R><<<grid, 256, smem, st>>>(k, (const uint16_t*)H, ldh, probs, (const uint16_t*)dOut, ldo, (uint16_t*)dH, lddh, \
                                                           relu_scale, dlogits, lddl, B, (int)pw); } while (0)
  const bool relu = relu_scale > 0.f;
  if (d->h == 128) { if (relu) MIXB(128, true); else MIXB(128, false); }
  else { if (relu) MIXB(64, true); else MIXB(64, false); }
#undef MIXB
  CDC_LAUNCHED();
  return 0;
}

}  // namespace cdcmdr
