// gpu.hh -- scoped libclumpy_cuda context for the command TUs.  There is no CPU path: when no B200
// is usable the command prints the library's error and fails (exit status 1).
#pragma once

#include "clumpy_cuda.h"

#include <cstdio>
#include <cstdlib>

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
