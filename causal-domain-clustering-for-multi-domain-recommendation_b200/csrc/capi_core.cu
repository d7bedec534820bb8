// Library-wide state: error string, launch counter, and the device-resident step state that lets a
// whole training step (dropout seeds, Adam bias corrections) be replayed from one CUDA graph.
#include "common.cuh"

namespace cdcmdr {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

__global__ void step_state_init_kernel(cdcmdr_step_state_t* st, int64_t step) {
  st->step = step; st->seed = 0; st->lr_t = 0.f; st->beta1 = 0.f; st->beta2 = 0.f; st->eps = 0.f;
  st->weight_decay = 0.f; st->bc2_sqrt = 1.f; st->pad[0] = st->pad[1] = 0.f;
}

// torch.optim.Adam (_single_tensor_adam): bias_correction1 = 1 - beta1^t ; step_size = lr / bias_correction1 ;
// bias_correction2_sqrt = sqrt(1 - beta2^t).  Computed in double like the reference's Python floats.
__global__ void step_tick_kernel(cdcmdr_step_state_t* st, float lr, float b1, float b2, float eps, float wd,
                                 uint64_t base_seed) {
  const int64_t t = st->step + 1;
  st->step = t;
  st->seed = ((uint64_t)mix_hash(base_seed, 0x51ed270bu, (uint64_t)t) << 32) | mix_hash(base_seed, 0x2545f491u, (uint64_t)t);
  const double bc1 = 1.0 - pow((double)b1, (double)t);
  const double bc2 = 1.0 - pow((double)b2, (double)t);
  st->lr_t = (float)((double)lr / bc1);
  st->beta1 = b1; st->beta2 = b2; st->eps = eps; st->weight_decay = wd;
  st->bc2_sqrt = (float)sqrt(bc2);
}
}  // namespace cdcmdr

using namespace cdcmdr;

extern "C" const char* cdcmdr_last_error(void) { return cdcmdr::g_err; }
extern "C" int cdcmdr_version(void) { return 101; }
extern "C" int64_t cdcmdr_launch_count(void) { return cdcmdr::g_launches.load(); }
extern "C" void cdcmdr_launch_count_reset(void) { cdcmdr::g_launches.store(0); }

extern "C" int cdcmdr_step_state_init(cdcmdr_step_state_t* st, int64_t step, cdcmdr_stream_t s) {
  CDC_REQUIRE(st, "null step state");
  step_state_init_kernel<<<1, 1, 0, to_stream(s)>>>(st, step);
  CDC_LAUNCHED();
  return 0;
}

extern "C" int cdcmdr_step_tick(cdcmdr_step_state_t* st, float lr, float beta1, float beta2, float eps, float weight_decay,
                                uint64_t base_seed, cdcmdr_stream_t s) {
  CDC_REQUIRE(st, "null step state");
  step_tick_kernel<<<1, 1, 0, to_stream(s)>>>(st, lr, beta1, beta2, eps, weight_decay, base_seed);
  CDC_LAUNCHED();
  return 0;
}
