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
  int n_pairs;
};

// thread 0 of the CTA: the compact gate-major list of (gate, expert) pairs
__device__ __forceinline__ void build_tables(const MixM& d, PairTables& t) {
  int n = 0;
  uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < d.n_gates; ++j) {
    const int first = n, cnt = d.gate_n[j], col = d.gate_col[j];
    for (int s = 0; s < cnt; ++s, ++n) {
      const int e = d.gate_sel[j * d.max_sel + s];
      if (n >= kMmaMaxPairs || e < 0 || e >= 16 || (seen[j] >> e & 1u)) __trap();     // launcher contract violated
      seen[j] |= 1u << e;
      t.gate[n] = j; t.e[n] = e; t.slot[n] = j * d.max_sel + s; t.col[n] = col + s; t.first[n] = first; t.cnt[n] = cnt;
    }
  }
  t.n_pairs = n;
}

// issue the cp.async copies of one row of `rows` blocks of HH bf16 into the swizzled buffer at `dst`
template <int HH>
__device__ __forceinline__ void stage_swizzled(uint32_t dst, const uint16_t* src, int rows, int lane) {
  constexpr int CPR = HH / 8;                            // 16-byte chunks per block row
  const int n = rows * CPR;
  const uint8_t* s8 = reinterpret_cast<const uint8_t*>(src);
  for (int i = lane; i < n; i += 32) {
    const int r = i / CPR, c = i % CPR;
    cp_async16(dst + r * (HH * 2) + ((c ^ (r & 7)) << 4), s8 + (size_t)i * 16);
  }
}

// ------------------------------------------------------------------------------------------------ forward
template <int HH>
__global__ void __launch_bounds__(256, 3)
gate_mix_fwd_mma_kernel(MixM d, const uint16_t* __restrict__ H, int64_t ldh, const float* __restrict__ logits, int64_t ldl,
                        uint16_t* __restrict__ out, int64_t ldo, float* __restrict__ probs, int64_t B, int per_warp) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ PairTables T;
  __shared__ __align__(16) uint4 zero16;
  constexpr int CPR = HH / 8;
  if (threadIdx.x == 0) { build_tables(d, T); zero16 = make_uint4(0, 0, 0, 0); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nE = d.n_experts, ng = d.n_gates, np = ng * d.max_sel;
  const int hb = nE * HH * 2, ob = ng * HH * 2;
  uint8_t* base = dyn + (size_t)warp * per_warp;
  uint8_t* Hbuf[2] = {base, base + hb};
  uint8_t* Ost = base + 2 * hb;
  float* sprob = reinterpret_cast<float*>(Ost + ob);     // [kMmaMaxNp]
  float* Pm = sprob + kMmaMaxNp;                         // [8][16]  P[j][e]
  float* sx = Pm + 128;                                  // [32] logits, then [32] exponentials
  for (int i = lane; i < 128; i += 32) Pm[i] = 0.f;
  for (int i = lane; i < kMmaMaxNp; i += 32) sprob[i] = 0.f;
  __syncthreads();
  const int n_pairs = T.n_pairs;
  const bool is_pair = lane < n_pairs;
  const int pj = is_pair ? T.gate[lane] : 0, pe = is_pair ? T.e[lane] : 0, pslot = is_pair ? T.slot[lane] : 0;
  const int pcol = is_pair ? T.col[lane] : 0, pfirst = is_pair ? T.first[lane] : 0, pcnt = is_pair ? T.cnt[lane] : 0;
  const uint32_t z16 = sm_u32(&zero16);
  const int j0 = lane >> 2, k0 = (lane & 3) * 2;
  // ldmatrix.x4.trans row of this lane: matrix m = lane/8: expert (m&1)*8 + lane%8, chunk 2g + (m>>1)
  const int le = ((lane >> 3) & 1) * 8 + (lane & 7), lch = lane >> 4;
  const bool le_ok = le < nE;

  const int64_t stride = (int64_t)gridDim.x * 8;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row < B) stage_swizzled<HH>(sm_u32(Hbuf[0]), H + row * ldh, nE, lane);
  cp_async_commit();
  int b = 0;
  for (; row < B; row += stride, b ^= 1) {
    const int64_t next = row + stride;
    if (next < B) stage_swizzled<HH>(sm_u32(Hbuf[b ^ 1]), H + next * ldh, nE, lane);
    cp_async_commit();
    // ---- softmax over each gate's logits, one (gate, expert) pair per lane
    const float x = is_pair ? logits[row * ldl + pcol] : 0.f;
    sx[lane] = x;
    __syncwarp();
    float mx = -INFINITY;
    for (int i = 0; i < pcnt; ++i) mx = fmaxf(mx, sx[pfirst + i]);
    const float ex = is_pair ? expf(x - mx) : 0.f;
    sx[32 + lane] = ex;
    __syncwarp();
    float sum = 0.f;
    for (int i = 0; i < pcnt; ++i) sum += sx[32 + pfirst + i];
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
    cp_async_wait1();
    __syncwarp();
    const uint32_t hrow = le_ok ? sm_u32(Hbuf[b]) + le * (HH * 2) : z16;
    const uint32_t hsw = le_ok ? (uint32_t)(le & 7) : 0u;
    const uint32_t ost = sm_u32(Ost) + j0 * (HH * 2) + (lane & 3) * 4;
#pragma unroll
    for (int g = 0; g < HH / 16; ++g) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(le_ok ? hrow + ((((uint32_t)(2 * g + lch)) ^ hsw) << 4) : z16, b0, b1, b2, b3);
      float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
      mma_16816(d0, ah0, 0u, ah2, 0u, b0, b1);
      mma_16816(d0, al0, 0u, al2, 0u, b0, b1);
      mma_16816(d1, ah0, 0u, ah2, 0u, b2, b3);
      mma_16816(d1, al0, 0u, al2, 0u, b2, b3);
      if (j0 < ng) {
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ost + ((((uint32_t)(2 * g)) ^ (uint32_t)(j0 & 7)) << 4)), "r"(pack_bf16x2(d0[0], d0[1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(ost + ((((uint32_t)(2 * g + 1)) ^ (uint32_t)(j0 & 7)) << 4)), "r"(pack_bf16x2(d1[0], d1[1])) : "memory");
      }
    }
    __syncwarp();
    {
      uint8_t* orow = reinterpret_cast<uint8_t*>(out + row * ldo);
      const int n = ng * CPR;
      for (int i = lane; i < n; i += 32) {
        const int r = i / CPR, c = i % CPR;
        *reinterpret_cast<uint4*>(orow + (size_t)i * 16) = *reinterpret_cast<const uint4*>(Ost + r * (HH * 2) + ((c ^ (r & 7)) << 4));
      }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ backward
template <int HH>
__global__ void __launch_bounds__(256, 3)
gate_mix_bwd_mma_kernel(MixM d, const uint16_t* __restrict__ H, int64_t ldh, const float* __restrict__ probs,
                        const uint16_t* __restrict__ dOut, int64_t ldo, uint16_t* __restrict__ dH, int64_t lddh, float relu_scale,
                        float* __restrict__ dlogits, int64_t lddl, int64_t B, int per_warp) {
  extern __shared__ __align__(128) uint8_t dyn[];
  __shared__ PairTables T;
  __shared__ __align__(16) uint4 zero16;
  constexpr int CPR = HH / 8;
  if (threadIdx.x == 0) { build_tables(d, T); zero16 = make_uint4(0, 0, 0, 0); }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nE = d.n_experts, ng = d.n_gates, np = ng * d.max_sel;
  const int hb = nE * HH * 2, ob = ng * HH * 2;
  uint8_t* base = dyn + (size_t)warp * per_warp;
  uint8_t* Hbuf[2] = {base, base + hb + ob};             // each buffer: H row | dOut row
  float* sprob = reinterpret_cast<float*>(base + 2 * (hb + ob));   // [kMmaMaxNp]
  float* Pm = sprob + kMmaMaxNp;                         // [8][16]  P[j][e]
  float* Ss = Pm + 128;                                  // [8][16]  S[j][e] = <dOut_j, H_e>
  for (int i = lane; i < 128; i += 32) Pm[i] = 0.f;
  __syncthreads();
  const int n_pairs = T.n_pairs;
  const bool is_pair = lane < n_pairs;
  const int pj = is_pair ? T.gate[lane] : 0, pe = is_pair ? T.e[lane] : 0, pslot = is_pair ? T.slot[lane] : 0;
  const int pcol = is_pair ? T.col[lane] : 0, pfirst = is_pair ? T.first[lane] : 0, pcnt = is_pair ? T.cnt[lane] : 0;
  const uint32_t z16 = sm_u32(&zero16);
  const int e0 = lane >> 2, k0 = (lane & 3) * 2;
  // dots: A = dOut via ldmatrix.x2 (lanes 0..15: gate lane%8, chunk 2ks + (lane/8)&1); B = H via ldmatrix.x4 (expert
  // (lane/16)*8 + lane%8, chunk 2ks + (lane/8)&1)
  const int aj = lane & 7, ach = (lane >> 3) & 1;
  const bool aj_ok = aj < ng;
  const int be = (lane >> 4) * 8 + (lane & 7), bch = (lane >> 3) & 1;
  const bool be_ok = be < nE;
  // dH: B = dOut via ldmatrix.x4.trans (gate lane%8, chunk 4g + lane/8)
  const int tj = lane & 7, tch = lane >> 3;
  const bool tj_ok = tj < ng;
  const bool has_hi = nE > 8;                            // experts 8..15 exist

  const int64_t stride = (int64_t)gridDim.x * 8;
  int64_t row = (int64_t)blockIdx.x * 8 + warp;
  if (row < B) {
    stage_swizzled<HH>(sm_u32(Hbuf[0]), H + row * ldh, nE, lane);
    stage_swizzled<HH>(sm_u32(Hbuf[0] + hb), dOut + row * ldo, ng, lane);
  }
  cp_async_commit();
  int b = 0;
  for (; row < B; row += stride, b ^= 1) {
    const int64_t next = row + stride;
    if (next < B) {
      stage_swizzled<HH>(sm_u32(Hbuf[b ^ 1]), H + next * ldh, nE, lane);
      stage_swizzled<HH>(sm_u32(Hbuf[b ^ 1] + hb), dOut + next * ldo, ng, lane);
    }
    cp_async_commit();
    for (int k = lane; k < np; k += 32) sprob[k] = probs[row * np + k];
    cp_async_wait1();
    __syncwarp();
    uint8_t* Hs = Hbuf[b];
    const uint32_t hs = sm_u32(Hs), os = hs + hb;
    // ---- S = dOut . H^T
    {
      float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t arow = os + aj * (HH * 2), brow = hs + be * (HH * 2);
      const uint32_t asw = (uint32_t)(aj & 7), bsw = (uint32_t)(be & 7);
#pragma unroll
      for (int ks = 0; ks < HH / 16; ++ks) {
        uint32_t a0, a2, b00, b01, b10, b11;
        ldsm_x2(aj_ok ? arow + ((((uint32_t)(2 * ks + ach)) ^ asw) << 4) : z16, a0, a2);
        ldsm_x4(be_ok ? brow + ((((uint32_t)(2 * ks + bch)) ^ bsw) << 4) : z16, b00, b01, b10, b11);
        mma_16816(s0, a0, 0u, a2, 0u, b00, b01);
        if (has_hi) mma_16816(s1, a0, 0u, a2, 0u, b10, b11);
      }
      // rows 0..7 of the 16 x 16 result: S[j = lane/4][e = k0, k0+1 (+8)]
      *reinterpret_cast<float2*>(Ss + e0 * 16 + k0) = make_float2(s0[0], s0[1]);
      *reinterpret_cast<float2*>(Ss + e0 * 16 + 8 + k0) = make_float2(s1[0], s1[1]);
    }
    __syncwarp();
    // ---- softmax backward per (gate, expert) pair: dz = p * (dp - <p, dp>)
    if (is_pair) {
      const float pr = sprob[pslot];
      const float dp = Ss[pj * 16 + pe];
      float dot = 0.f;
      for (int i = 0; i < pcnt; ++i) dot = fmaf(sprob[T.slot[pfirst + i]], Ss[pj * 16 + T.e[pfirst + i]], dot);
      dlogits[row * lddl + pcol] = pr * (dp - dot);
      Pm[pj * 16 + pe] = pr;
    }
    __syncwarp();
    // ---- dH = P^T . dOut:  A[m = e][k = j] = P[j][e]
    uint32_t ah0, al0, ah1, al1;
    split_hi_lo(Pm[k0 * 16 + e0], Pm[(k0 + 1) * 16 + e0], ah0, al0);
    split_hi_lo(Pm[k0 * 16 + e0 + 8], Pm[(k0 + 1) * 16 + e0 + 8], ah1, al1);
    const uint32_t trow = os + tj * (HH * 2);
    const uint32_t tsw = (uint32_t)(tj & 7);
    const bool lo_ok = e0 < nE, hi_ok = e0 + 8 < nE;
    const uint32_t hlo = hs + e0 * (HH * 2) + (lane & 3) * 4, hhi = hlo + 8 * (HH * 2);
    const uint32_t esw = (uint32_t)(e0 & 7);
#pragma unroll
    for (int g = 0; g < HH / 32; ++g) {
      uint32_t bb[4];
      ldsm_x4_t(tj_ok ? trow + ((((uint32_t)(4 * g + tch)) ^ tsw) << 4) : z16, bb[0], bb[1], bb[2], bb[3]);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float dd[4] = {0.f, 0.f, 0.f, 0.f};
        mma_1688(dd, ah0, ah1, bb[t]);
        mma_1688(dd, al0, al1, bb[t]);
        const uint32_t coff = (((uint32_t)(4 * g + t)) ^ esw) << 4;
        if (lo_ok) {
          uint32_t o;
          if (relu_scale > 0.f) {
            uint32_t hv;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv) : "r"(hlo + coff) : "memory");
            const uint32_t l = hv & 0xFFFFu, u = hv >> 16;
            o = pack_bf16x2((l != 0u && l < 0x8000u) ? dd[0] * relu_scale : 0.f, (u != 0u && u < 0x8000u) ? dd[1] * relu_scale : 0.f);
          } else {
            o = pack_bf16x2(dd[0], dd[1]);
          }
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(hlo + coff), "r"(o) : "memory");
        }
        if (hi_ok) {
          uint32_t o;
          if (relu_scale > 0.f) {
            uint32_t hv;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(hv) : "r"(hhi + coff) : "memory");
            const uint32_t l = hv & 0xFFFFu, u = hv >> 16;
            o = pack_bf16x2((l != 0u && l < 0x8000u) ? dd[2] * relu_scale : 0.f, (u != 0u && u < 0x8000u) ? dd[3] * relu_scale : 0.f);
          } else {
            o = pack_bf16x2(dd[2], dd[3]);
          }
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(hhi + coff), "r"(o) : "memory");
        }
      }
    }
    __syncwarp();
    {
      uint8_t* drow = reinterpret_cast<uint8_t*>(dH + row * lddh);
      const int n = nE * CPR;
      for (int i = lane; i < n; i += 32) {
        const int r = i / CPR, c = i % CPR;
        *reinterpret_cast<uint4*>(drow + (size_t)i * 16) = *reinterpret_cast<const uint4*>(Hs + r * (HH * 2) + ((c ^ (r & 7)) << 4));
      }
    }
    __syncwarp();
  }
}

constexpr int kMmaSmemPerSM = 216 * 1024;
constexpr int kMmaSmemPerCta = 100 * 1024;

bool mma_shape_ok(const cdcmdr_mix_desc_t* d, int is_bf16) {
  return is_bf16 && (d->h == 64 || d->h == 128) && d->n_gates <= 8 && d->n_experts <= 16 && d->n_pairs > 0 && d->n_pairs <= kMmaMaxPairs &&
         d->n_gates * d->max_sel <= kMmaMaxNp;
}
bool al16(const void* p) { return ((uintptr_t)p % 16) == 0; }

int pick_grid(int64_t B, size_t smem) {
  int per_sm = (int)(kMmaSmemPerSM / (smem + 1024));
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
  const size_t pw = (per_warp + 127) & ~(size_t)127;
  const size_t smem = 8 * pw;
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
  const size_t per_warp = (size_t)2 * (d->n_experts + d->n_gates) * d->h * 2 + (kMmaMaxNp + 128 + 128) * sizeof(float);
  const size_t pw = (per_warp + 127) & ~(size_t)127;
  const size_t smem = 8 * pw;
  if (smem > (size_t)kMmaSmemPerCta) return -1;
  MixM k{d->n_gates, d->n_experts, d->max_sel, d->n_pairs, d->gate_col, d->gate_n, d->gate_sel};
  const int grid = pick_grid(B, smem);
  if (d->h == 128) {
    CDC_CHECK(cudaFuncSetAttribute(gate_mix_bwd_mma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemPerCta));
    gate_mix_bwd_mma_kernel<128><<<grid, 256, smem, st>>>(k, (const uint16_t*)H, ldh, probs, (const uint16_t*)dOut, ldo, (uint16_t*)dH, lddh,
                                                          relu_scale, dlogits, lddl, B, (int)pw);
  } else {
    CDC_CHECK(cudaFuncSetAttribute(gate_mix_bwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMmaSmemPerCta));
    gate_mix_bwd_mma_kernel<64><<<grid, 256, smem, st>>>(k, (const uint16_t*)H, ldh, probs, (const uint16_t*)dOut, ldo, (uint16_t*)dH, lddh,
                                                         relu_scale, dlogits, lddl, B, (int)pw);
  }
  CDC_LAUNCHED();
  return 0;
}

}  // namespace cdcmdr
