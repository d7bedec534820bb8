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

}  // namespace cdcmdr

using namespace cdcmdr;

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
