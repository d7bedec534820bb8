#include "common.cuh"
namespace cdcmdr {
thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
}
extern "C" const char* cdcmdr_last_error(void) { return cdcmdr::g_err; }
extern "C" int cdcmdr_version(void) { return 100; }
extern "C" int64_t cdcmdr_launch_count(void) { return cdcmdr::g_launches.load(); }
extern "C" void cdcmdr_launch_count_reset(void) { cdcmdr::g_launches.store(0); }
