// Shared host-side helpers of libb200e2tts.so: error reporting, launch accounting, launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/b200_e2tts.h"

namespace b200 {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
// Optional device-resident addend of every dropout seed (b200_set_dropout_seed_device): lets a captured CUDA graph draw fresh
// dropout masks on every replay — the host-side seed is a kernel argument and therefore frozen at capture time.
extern std::atomic<const unsigned long long*> g_seed_dev;
inline const unsigned long long* seed_dev_ptr() { return g_seed_dev.load(std::memory_order_relaxed); }

#define B200_FAIL(...)                                   \
    do {                                                 \
        snprintf(b200::g_err, sizeof(b200::g_err), __VA_ARGS__); \
        return -1;                                       \
    } while (0)

#define B200_REQUIRE(cond, ...) \
    do {                        \
        if (!(cond)) B200_FAIL(__VA_ARGS__); \
    } while (0)

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}

inline int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace b200
