// Library-wide state of libb200e2tts.so (error string, launch counter, version).
#include "common.cuh"

namespace b200 {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};
std::atomic<const unsigned long long*> g_seed_dev{nullptr};
}  // namespace b200

extern "C" int b200_set_dropout_seed_device(const uint64_t* dev_seed) {
    b200::g_seed_dev.store(reinterpret_cast<const unsigned long long*>(dev_seed));
    return 0;
}

extern "C" const char* b200_last_error(void) { return b200::g_err; }
extern "C" int b200_version(void) { return 100; }
extern "C" uint64_t b200_launch_count(void) { return b200::g_launches.load(); }
