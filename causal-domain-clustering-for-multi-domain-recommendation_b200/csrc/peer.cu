// the small, latency-bound exchanges of the data-parallel step over NVLink peer memory instead of NCCL calls.
//
// cdcmdr_peer_allreduce_f64: cross-replica BatchNorm statistics (double[2*C] per layer and direction) and the loss sums are a few
// KB each and sit on the critical path of the model program: an NCCL all-reduce costs 15-25 us of fixed latency apiece.  Here every
// rank owns one SYMMETRIC buffer (torch's symmetric-memory allocator supplies the peer mappings - plumbing) laid out as
//   [2 parities][world slots][max_n doubles] | flags: [2 parities][world] uint64
// and ONE kernel does the whole exchange: store my n values into slot[rank] of every peer's buffer over NVLink, system-scope
// fence, release-store the call's sequence number into flag[rank] of every peer, spin (acquire) until all `world` flags of MY
// buffer carry the sequence number, then sum the world slots in rank order - the same order on every rank, so replicas stay bit
// identical.  Calls alternate between the two parities: a rank can only reach call k+2 (which reuses call k's slots) after every
// rank has signalled call k+1, i.e. after every rank finished reading call k.  The sequence counter lives in device memory, so a
// captured CUDA graph replays the exchange without host involvement.
#include "common.cuh"

namespace cdcmdr {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(256)
peer_allreduce_f64_kernel(double* const* __restrict__ bufs, int rank, int world, const double* __restrict__ in, double* __restrict__ out,
                          int n, int64_t max_n, unsigned long long* __restrict__ seq_p) {
  __shared__ unsigned long long seq_s;
  if (threadIdx.x == 0) seq_s = *seq_p + 1ull;
  __syncthreads();
  const unsigned long long seq = seq_s;
  const int par = (int)(seq & 1ull);
  const int64_t par_stride = (int64_t)world * max_n;
  for (int r = 0; r < world; ++r) {                               // my contribution -> slot[par][rank] of every rank (mine included)
    double* dst = bufs[r] + par * par_stride + (int64_t)rank * max_n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = in[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(bufs[threadIdx.x] + 2 * par_stride) + par * world + rank;
    st_release_sys(flag, seq);
  }
  if ((int)threadIdx.x < world) {
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(bufs[rank] + 2 * par_stride) + par * world + threadIdx.x;
    while (ld_acquire_sys(mine) < seq) { }
  }
  __syncthreads();
  const double* slots = bufs[rank] + par * par_stride;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double a = 0.0;
    for (int r = 0; r < world; ++r) a += __ldcg(slots + (int64_t)r * max_n + i);     // written by peers: L2 is the coherence point
    out[i] = a;
  }
  if (threadIdx.x == 0) *seq_p = seq;
}

// ---- the dense gradient arena (a few MB of fp32) between replicas: two-shot all-reduce over peer memory --------------------------
// push: chunk c of my vector -> slot[rank] of rank c's inbox (peer stores) | barrier | rank c sums its inbox slots in rank order
// (the same order everywhere; ONE rank computes each element, so replicas stay bit identical) and stores the sums into every rank's
// outbox (peer stores) | barrier | outbox -> out.  Each rank moves 2 (N-1)/N of the vector over NVLink, as a ring does, but in two
// hops instead of 2 (N-1): at 5.6 MB the NCCL ring took 50-85 us of the step's tail, this takes ~20.
__global__ void __launch_bounds__(256)
peer_ar_push_kernel(const float* __restrict__ in, int64_t n, int64_t chunk, float* const* __restrict__ inbox, int rank) {
  const int c = blockIdx.y;
  const int64_t lo = (int64_t)c * chunk;
  const int64_t len = min(chunk, n - lo);
  if (len <= 0) return;
  const float* src = in + lo;
  float* dst = inbox[c] + (int64_t)rank * chunk;
  const int64_t n4 = len >> 2;
  const bool vec = ((uintptr_t)src % 16) == 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < (vec ? n4 : 0); i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
  for (int64_t i = (vec ? 4 * n4 : 0) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void __launch_bounds__(256)
peer_ar_reduce_kernel(const float* __restrict__ my_inbox, int64_t n, int64_t chunk, float* const* __restrict__ outbox, int rank, int world) {
  const int64_t lo = (int64_t)rank * chunk;
  const int64_t len = min(chunk, n - lo);
  if (len <= 0) return;
  const int64_t n4 = len >> 2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {
      const float4 t = __ldcg(reinterpret_cast<const float4*>(my_inbox + (int64_t)r * chunk) + i);      // written by peers: read at L2
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    for (int q = 0; q < world; ++q) reinterpret_cast<float4*>(outbox[q] + lo)[i] = a;
  }
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x) {
    float a = 0.f;
    for (int r = 0; r < world; ++r) a += __ldcg(my_inbox + (int64_t)r * chunk + i);
    for (int q = 0; q < world; ++q) outbox[q][lo + i] = a;
  }
}

__global__ void __launch_bounds__(256)
peer_ar_copy_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const bool vec = ((uintptr_t)dst % 16) == 0;
  const int64_t n4 = vec ? n >> 2 : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = __ldcg(reinterpret_cast<const float4*>(src) + i);
  for (int64_t i = 4 * n4 + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldcg(src + i);
}

__global__ void __launch_bounds__(32)
peer_flag_barrier_kernel(unsigned long long* const* __restrict__ flags, int rank, int world, int slot, unsigned long long* __restrict__ seqs) {
  const unsigned long long seq = seqs[slot] + 1ull;
  __syncwarp();
  __threadfence_system();
  for (int r = threadIdx.x; r < world; r += 32) st_release_sys(flags[r] + (int64_t)slot * world + rank, seq);
  for (int r = threadIdx.x; r < world; r += 32) {
    const unsigned long long* mine = flags[rank] + (int64_t)slot * world + r;
    while (ld_acquire_sys(mine) < seq) { }
  }
  __syncwarp();
  if (threadIdx.x == 0) seqs[slot] = seq;
}

}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" int64_t cdcmdr_peer_allreduce_f32_chunk(int world, int64_t n) {
  if (world < 1 || n < 1) return 0;
  const int64_t c = (n + world - 1) / world;
  return (c + 3) & ~(int64_t)3;
}

extern "C" int cdcmdr_peer_allreduce_f32(float* const* inbox, float* const* outbox, uint64_t* const* peer_flags, float* my_inbox,
                                         const float* my_outbox, int rank, int world, const float* in, float* out, int64_t n,
                                         uint64_t* seqs2, cdcmdr_stream_t s) {
  CDC_REQUIRE(inbox && outbox && peer_flags && my_inbox && my_outbox && in && out && seqs2, "peer all-reduce (f32): null argument");
  CDC_REQUIRE(world >= 1 && world <= 256 && rank >= 0 && rank < world && n >= 0, "peer all-reduce (f32): bad rank / world");
  if (n == 0) return 0;
  const int64_t chunk = cdcmdr_peer_allreduce_f32_chunk(world, n);
  cudaStream_t st = to_stream(s);
  auto blocks = [](int64_t work) {
    int64_t g = ceil_div(work, (int64_t)256);
    const int64_t cap = 4 * (int64_t)kNumSMs;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
  };
  unsigned long long* const* fl = reinterpret_cast<unsigned long long* const*>(peer_flags);
  unsigned long long* sq = reinterpret_cast<unsigned long long*>(seqs2);
  peer_ar_push_kernel<<<dim3(blocks(chunk / 4), (unsigned)world), 256, 0, st>>>(in, n, chunk, inbox, rank);
  CDC_LAUNCHED();
  peer_flag_barrier_kernel<<<1, 32, 0, st>>>(fl, rank, world, 0, sq);
  CDC_LAUNCHED();
  peer_ar_reduce_kernel<<<blocks(chunk / 4), 256, 0, st>>>(my_inbox, n, chunk, outbox, rank, world);
  CDC_LAUNCHED();
  peer_flag_barrier_kernel<<<1, 32, 0, st>>>(fl, rank, world, 1, sq);
  CDC_LAUNCHED();
  peer_ar_copy_kernel<<<blocks(n / 4), 256, 0, st>>>(my_outbox, out, n);
  CDC_LAUNCHED();
  return 0;
}
extern "C" size_t cdcmdr_peer_allreduce_bytes(int world, int64_t max_n) {
  if (world < 1 || max_n < 1) return 0;
  return (size_t)2 * world * max_n * sizeof(double) + (size_t)2 * world * sizeof(unsigned long long);
}

extern "C" int cdcmdr_peer_allreduce_f64(double* const* peer_bufs, int rank, int world, const double* in, double* out, int64_t n,
                                         int64_t max_n, uint64_t* seq, cdcmdr_stream_t s) {
  CDC_REQUIRE(peer_bufs && in && out && seq, "peer all-reduce: null argument");
  CDC_REQUIRE(world >= 1 && world <= 256 && rank >= 0 && rank < world, "peer all-reduce: bad rank / world");
  CDC_REQUIRE(n >= 0 && n <= max_n && n < ((int64_t)1 << 31), "peer all-reduce: n exceeds the symmetric buffer's slot size");
  if (n == 0) return 0;
  peer_allreduce_f64_kernel<<<1, 256, 0, to_stream(s)>>>(peer_bufs, rank, world, in, out, (int)n, max_n,
                                                         reinterpret_cast<unsigned long long*>(seq));
  CDC_LAUNCHED();
  return 0;
}
