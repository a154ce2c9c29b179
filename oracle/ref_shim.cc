// ref_shim.cc -- plain-C doors into the UNMODIFIED reference objects (oracle/_ref/*.o).
//
// TEST INFRASTRUCTURE ONLY.  This file is our own code; it is linked with objects compiled from
// /root/reference (see oracle/Makefile) into oracle/_ref/libclumpy_ref.so so that tests can call
// the reference's routines in-process, without going through .npy files:
//
//   splat_disks / splat_points   external linkage, commands/splat_points.cc:108,118
//   advect_wrapx                 global, commands/splat_points.cc:20
//   open_simplex_noise{,2,_free} external linkage, commands/generate_simplex.cc:241,286,298
//   ClumpyCommand::registry()    clumpy_command.hh:21-24 (run any command's exec in-process)
//
// glm and clumpy_command.hh are found through -I/root/reference at build time.
#include "clumpy_command.hh"

#include <glm/vec2.hpp>
#include <glm/gtc/type_precision.hpp>

#include <cstdint>
#include <random>
#include <string>
#include <vector>

void splat_points(glm::vec2 const* ptlist, uint32_t npts, glm::u32vec2 size, uint8_t* dstimg);
void splat_disks(glm::vec2 const* ptlist, uint32_t npts, glm::u32vec2 dims, uint8_t* dstimg,
                 float alpha, int kernel_size);
extern bool advect_wrapx;

struct osn_context { // layout of commands/generate_simplex.cc:114-117
    int16_t* perm;
    int16_t* permGradIndex3D;
};
int open_simplex_noise(int64_t seed, struct osn_context** ctx);
void open_simplex_noise_free(struct osn_context* ctx);
double open_simplex_noise2(struct osn_context* ctx, double x, double y);

#define SHIM_API extern "C" __attribute__((visibility("default")))

SHIM_API void ref_splat_disks(const float* pts, uint32_t npts, uint32_t width, uint32_t height,
                              uint8_t* img, float alpha, int kernel_size, int wrapx) {
    advect_wrapx = wrapx != 0;
    splat_disks(reinterpret_cast<glm::vec2 const*>(pts), npts, glm::u32vec2(width, height), img,
                alpha, kernel_size);
    advect_wrapx = false;
}

SHIM_API void ref_splat_points(const float* pts, uint32_t npts, uint32_t width, uint32_t height,
                               uint8_t* img) {
    splat_points(reinterpret_cast<glm::vec2 const*>(pts), npts, glm::u32vec2(width, height), img);
}

SHIM_API void ref_simplex_perm(int64_t seed, int16_t* perm_out) {
    osn_context* ctx = nullptr;
    open_simplex_noise(seed, &ctx);
    for (int i = 0; i < 256; ++i) perm_out[i] = ctx->perm[i];
    open_simplex_noise_free(ctx);
}

SHIM_API void ref_simplex2_many(int64_t seed, const double* xy, uint32_t n, double* out) {
    osn_context* ctx = nullptr;
    open_simplex_noise(seed, &ctx);
    for (uint32_t i = 0; i < n; ++i) out[i] = open_simplex_noise2(ctx, xy[2 * i], xy[2 * i + 1]);
    open_simplex_noise_free(ctx);
}

// The reset offsets exactly as commands/advect_points.cc:127-135 draws them: <random> itself (the
// distribution's algorithm is the host library's, not the standard's), for pinning the restatements.
SHIM_API void ref_std_age_offsets(uint32_t nframes, uint32_t npts, uint32_t* out) {
    std::mt19937 generator;
    generator.seed(0);
    std::uniform_int_distribution<> get_age_offset(0, (int)(nframes - 1));
    for (uint32_t i = 0; i < npts; ++i) out[i] = (uint32_t)get_age_offset(generator);
}

// Runs `clumpy <name> argv...` in-process; returns the exit code main_clumpy.cc:59 would produce,
// or 2 when the command does not exist.  advect_wrapx is reset afterwards (the reference CLI is
// one command per process, so it never needed to).
SHIM_API int ref_exec(const char* name, int argc, const char** argv) {
    auto& reg = ClumpyCommand::registry();
    auto it = reg.find(name);
    if (it == reg.end()) return 2;
    ClumpyCommand* cmd = it->second();
    std::vector<std::string> args(argv, argv + argc);
    bool ok = cmd->exec(args);
    delete cmd;
    advect_wrapx = false;
    return ok ? 0 : 1;
}
