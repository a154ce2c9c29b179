// gpu.hh -- scoped libclumpy_cuda context for the command TUs.  There is no CPU path: when no B200
// is usable the command prints the library's error and fails (exit status 1).
#pragma once

#include "clumpy_cuda.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

// CLUMPY_DEVICES="0,1,2,3" or "0-7": the GPUs a command may spread over (advect_points does; one process drives
// them all).  Empty when unset or naming a single device (then CLUMPY_DEVICE / device 0 is used as before).
inline std::vector<int> clumpy_devices() {
    std::vector<int> out;
    const char* s = getenv("CLUMPY_DEVICES");
    if (!s) return out;
    while (*s) {
        char* end = nullptr;
        const long a = strtol(s, &end, 10);
        if (end == s) break;
        long b = a;
        s = end;
        if (*s == '-') { b = strtol(s + 1, &end, 10); s = end; }
        for (long d = a; d <= b && out.size() < 8; ++d) out.push_back((int)d);
        if (*s == ',') ++s; else break;
    }
    if (out.size() < 2) out.clear();
    return out;
}

struct Gpu {
    clumpy_cuda_ctx* ctx = nullptr;
    Gpu() {
        const char* dev = getenv("CLUMPY_DEVICE");
        if (clumpy_cuda_create(dev ? atoi(dev) : 0, nullptr, &ctx) != CLUMPY_OK) {
            printf("clumpy_cuda: %s\n", clumpy_cuda_last_error());
            ctx = nullptr;
        }
    }
    ~Gpu() { clumpy_cuda_destroy(ctx); }
    explicit operator bool() const { return ctx != nullptr; }
    static bool report(int rc) {
        if (rc == CLUMPY_OK) return true;
        printf("clumpy_cuda: %s\n", clumpy_cuda_last_error());
        return false;
    }
};
