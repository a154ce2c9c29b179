// simplex.cu -- 2-D OpenSimplex noise field (clumpy `generate_simplex`) for sm_100a.
//
// Replaces the per-pixel loop of commands/generate_simplex.cc:59-82 and the noise evaluation it
// calls (:298-417).  One thread evaluates PX consecutive pixels of a row-major band.
//
// Precision split (DESIGN.md "simplex"): the reference evaluates in FP64 on an FP32 coordinate
// product.  Here the lattice placement -- skew, floor, un-skew, position inside the cell, the region
// tests: the part whose error grows with |coordinate| -- runs in FP64 exactly as the reference
// writes it; the four vertex contributions, whose inputs are O(1), run on the FP32 pipe.  Measured
// against the FP64 reference this stays below 6e-7 absolute for amplitude 1 at any frequency
// (tolerance 1e-5).  The 256-entry permutation lives in shared memory as bytes.
#include "device_utils.cuh"
#include "internal.h"

#include <algorithm>

namespace clumpy {

constexpr int SIMPLEX_THREADS = 256;
constexpr int SIMPLEX_PX = 4; // pixels per thread (one float4 store when aligned)

// Gradient of generate_simplex.cc:124-126 for hash h = perm[...] & 0x0E:
// pair index p = h >> 1; bit0 of p swaps the magnitudes (5,2)->(2,5), bit1 negates x, bit2 negates y.
__device__ __forceinline__ float vertex_term(const uint8_t* __restrict__ perm, int xv, int yv,
                                             float dx, float dy) {
    float attn = 2.0f - dx * dx - dy * dy;
    if (attn <= 0.0f) return 0.0f;
    uint32_t h = perm[(perm[xv & 0xFF] + yv) & 0xFF];
    float gx = (h & 2u) ? 2.0f : 5.0f;
    float gy = (h & 2u) ? 5.0f : 2.0f;
    gx = __uint_as_float(__float_as_uint(gx) | ((h & 4u) << 29));
    gy = __uint_as_float(__float_as_uint(gy) | ((h & 8u) << 28));
    attn *= attn;
    return attn * attn * (gx * dx + gy * dy);
}

// generate_simplex.cc:298-417; u, v are the float coordinates widened to double (:75-76).
__device__ __forceinline__ float simplex2(const uint8_t* __restrict__ perm, float uf, float vf) {
    const double kStretch = -0.211324865405187;
    const double kSquish = 0.366025403784439;
    const float kSquishF = 0.366025403784439f;
    const double u = (double)uf, v = (double)vf;
    // placement in FP64, operation for operation (no contraction)
    const double so = __dmul_rn(__dadd_rn(u, v), kStretch);
    const double xs = __dadd_rn(u, so), ys = __dadd_rn(v, so);
    const int xsb = __double2int_rd(xs), ysb = __double2int_rd(ys);
    const double qo = __dmul_rn((double)(xsb + ysb), kSquish);
    const double xins = __dsub_rn(xs, (double)xsb), yins = __dsub_rn(ys, (double)ysb);
    const double in_sum = __dadd_rn(xins, yins);
    const float d0x = (float)__dsub_rn(u, __dadd_rn((double)xsb, qo));
    const float d0y = (float)__dsub_rn(v, __dadd_rn((double)ysb, qo));

    // third vertex (0,0) or (1,1); extra vertex from the region tests of :351-399
    int ti, ei, ej;
    if (in_sum <= 1.0) {
        const double zins = 1.0 - in_sum;
        ti = 0;
        if (zins > xins || zins > yins) {
            const bool xg = xins > yins;
            ei = xg ? 1 : -1;
            ej = xg ? -1 : 1;
        } else { ei = 1; ej = 1; }
    } else {
        const double zins = 2.0 - in_sum;
        ti = 1;
        if (zins < xins || zins < yins) {
            const bool xg = xins > yins;
            ei = xg ? 2 : 0;
            ej = xg ? 0 : 2;
        } else { ei = 0; ej = 0; }
    }
    // displacement of vertex (i, j): d0 - (i, j) - (i + j) * squish
    float value = vertex_term(perm, xsb + 1, ysb, d0x - 1.0f - kSquishF, d0y - kSquishF);
    value += vertex_term(perm, xsb, ysb + 1, d0x - kSquishF, d0y - 1.0f - kSquishF);
    {
        const float t = (float)ti, off = 2.0f * kSquishF * t;
        value += vertex_term(perm, xsb + ti, ysb + ti, d0x - t - off, d0y - t - off);
    }
    {
        const float off = (float)(ei + ej) * kSquishF;
        value += vertex_term(perm, xsb + ei, ysb + ej, d0x - (float)ei - off, d0y - (float)ej - off);
    }
    return value;
}

// Band rows [row0, row0 + nrows) of a `width`-wide image; out is nrows x width.
struct PermArg { uint8_t p[256]; }; // passed by value: lands in the constant bank, no allocation

// Pixel indexing shared by the two kernels: a CTA covers SIMPLEX_THREADS * SIMPLEX_PX consecutive pixels of ONE
// row (blockIdx.x strides over the row, blockIdx.y over the rows), so no 64-bit division is needed per pixel
// group (round 1 indexed a flat pixel range: ~15 of its ~215 instructions per pixel were that division).
template <typename Eval>
__device__ __forceinline__ void simplex_rows(uint32_t width, uint32_t nrows, float* __restrict__ out,
                                             uint32_t* __restrict__ minmax, int vec_ok, Eval eval) {
    float lo = 0.f, hi = 0.f;
    bool any = false;
    for (uint32_t row = blockIdx.y; row < nrows; row += gridDim.y) {
        float* const orow = out + (size_t)row * width;
        for (uint32_t col0 = (blockIdx.x * SIMPLEX_THREADS + threadIdx.x) * SIMPLEX_PX; col0 < width;
             col0 += gridDim.x * SIMPLEX_THREADS * SIMPLEX_PX) {
            float vals[SIMPLEX_PX];
            const int n = (width - col0 < (uint32_t)SIMPLEX_PX) ? (int)(width - col0) : SIMPLEX_PX;
#pragma unroll
            for (int k = 0; k < SIMPLEX_PX; ++k) {
                if (k < n) {
                    const float s = eval(col0 + (uint32_t)k, row);
                    vals[k] = s;
                    if (!any) { lo = hi = s; any = true; }
                    else { lo = (s < lo) ? s : lo; hi = (hi < s) ? s : hi; }
                }
            }
            if (vec_ok && n == SIMPLEX_PX) {
                *reinterpret_cast<float4*>(orow + col0) = make_float4(vals[0], vals[1], vals[2], vals[3]);
            } else {
#pragma unroll
                for (int k = 0; k < SIMPLEX_PX; ++k)
                    if (k < n) orow[col0 + k] = vals[k];
            }
        }
    }
    block_minmax_commit(lo, any, hi, any, minmax);
}

__global__ void __launch_bounds__(SIMPLEX_THREADS)
simplex_kernel(const PermArg perm_arg, uint32_t width, uint32_t row0, uint32_t nrows,
               float step, float scale, float* __restrict__ out, uint32_t* __restrict__ minmax,
               int vec_ok) {
    __shared__ uint8_t perm[256];
    perm[threadIdx.x] = perm_arg.p[threadIdx.x]; // blockDim.x == 256
    __syncthreads();
    simplex_rows(width, nrows, out, minmax, vec_ok, [&](uint32_t col, uint32_t row) {
        // generate_simplex.cc:75-77: FP32 product step*index, then amplitude * noise / 47
        const float uf = __fmul_rn(step, (float)col);
        const float vf = __fmul_rn(step, (float)(row0 + row));
        return simplex2(perm, uf, vf) * scale;
    });
}

// Several octaves in one pass (README.md:28-36, extras/test.py:18-21: the reference runs generate_simplex once
// per octave and sums the .npy files with numpy, i.e. in FP32, left to right).  Each octave's value is computed
// exactly as simplex_kernel computes it (same FP32 coordinate product, same FP64 placement, same scale), and the
// octaves are added in FP32 in the order given, so the field equals the sum of the single-octave fields bit for
// bit -- without writing and re-reading one 1 GB image per octave at the C3 size.
constexpr int SIMPLEX_MAX_OCTAVES = 8;
struct OctavePerms { uint8_t p[SIMPLEX_MAX_OCTAVES][256]; };
struct OctaveArgs { int n; float step[SIMPLEX_MAX_OCTAVES]; float scale[SIMPLEX_MAX_OCTAVES]; };

__global__ void __launch_bounds__(SIMPLEX_THREADS)
simplex_octaves_kernel(const OctavePerms perms_arg, const OctaveArgs oa, uint32_t width, uint32_t row0, uint32_t nrows,
                       float* __restrict__ out, uint32_t* __restrict__ minmax, int vec_ok) {
    __shared__ uint8_t perm[SIMPLEX_MAX_OCTAVES][256];
    // The per-octave step and scale go to shared memory too: indexed by the loop variable straight out of the
    // kernel parameters they became a chain of 16 compare + predicated constant loads per pixel and octave (a quarter
    // of the kernel's instructions, round 2 SASS).
    __shared__ float oct_step[SIMPLEX_MAX_OCTAVES], oct_scale[SIMPLEX_MAX_OCTAVES];
    for (int o = 0; o < oa.n; ++o) perm[o][threadIdx.x] = perms_arg.p[o][threadIdx.x]; // blockDim.x == 256
    if (threadIdx.x < SIMPLEX_MAX_OCTAVES) { oct_step[threadIdx.x] = oa.step[threadIdx.x]; oct_scale[threadIdx.x] = oa.scale[threadIdx.x]; }
    __syncthreads();
    const int noct = oa.n;
    simplex_rows(width, nrows, out, minmax, vec_ok, [&](uint32_t col, uint32_t row) {
        float total = 0.f;
        for (int o = 0; o < noct; ++o) {
            const float step = oct_step[o];
            const float uf = __fmul_rn(step, (float)col);
            const float vf = __fmul_rn(step, (float)(row0 + row));
            const float s = simplex2(perm[o], uf, vf) * oct_scale[o];
            total = o == 0 ? s : __fadd_rn(total, s);
        }
        return total;
    });
}

// A grid for simplex_rows: the row's pixel groups along x, as many rows along y as fill a few waves.
static dim3 simplex_grid(const clumpy_cuda_ctx* ctx, uint32_t width, uint32_t nrows) {
    const uint32_t gx = std::max<uint32_t>(1u, ceil_div(width, SIMPLEX_THREADS * SIMPLEX_PX));
    const uint32_t want_y = std::max<uint32_t>(1u, ((uint32_t)ctx->sm_count * 8 * 4 + gx - 1) / gx);
    return dim3(gx, std::min<uint32_t>(std::min<uint32_t>(nrows, want_y), 65535u));
}

static int run_simplex(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height, uint32_t row0,
                       uint32_t nrows, double amplitude, double frequency, int64_t seed,
                       float* d_out, float* minval, float* maxval) {
    int16_t perm16[256];
    clumpy_cuda_simplex_perm(seed, perm16);
    PermArg perm8;
    for (int i = 0; i < 256; ++i) perm8.p[i] = (uint8_t)perm16[i];
    // generate_simplex.cc:59-67: float step from a double divide by the shorter side
    const float step = (width > height) ? (float)(frequency / height) : (float)(frequency / width);
    const float scale = (float)(amplitude / 47.0);
    int rc = minmax_reset(ctx);
    if (rc) return rc;
    if (nrows > 0) {
        const int vec_ok = ((uintptr_t)d_out % 16 == 0 && width % 4 == 0) ? 1 : 0;
        simplex_kernel<<<simplex_grid(ctx, width, nrows), SIMPLEX_THREADS, 0, ctx->stream>>>(perm8, width, row0, nrows, step,
                                                                                              scale, d_out, ctx->d_minmax, vec_ok);
        CK_LAUNCH(ctx);
    }
    if (minval || maxval) return minmax_fetch(ctx, minval, maxval);
    return CLUMPY_OK;
}

} // namespace clumpy

using namespace clumpy;

// commands/generate_simplex.cc:241-280: 64-bit LCG, three warm-up steps, then draw perm[255..0]
// from a shrinking pool.
CLUMPY_API int clumpy_cuda_simplex_perm(int64_t seed, int16_t* perm256) {
    REQUIRE(perm256, "perm256 is NULL");
    const uint64_t mul = 6364136223846793005ULL, inc = 1442695040888963407ULL;
    int16_t pool[256];
    for (int i = 0; i < 256; ++i) pool[i] = (int16_t)i;
    uint64_t s = (uint64_t)seed;
    for (int w = 0; w < 3; ++w) s = s * mul + inc;
    for (int i = 255; i >= 0; --i) {
        s = s * mul + inc;
        int r = (int)((int64_t)(s + 31ULL) % (int64_t)(i + 1));
        if (r < 0) r += i + 1;
        perm256[i] = pool[r];
        pool[r] = pool[i];
    }
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_generate_simplex_dev(clumpy_cuda_ctx* ctx, uint32_t width,
                                                uint32_t height, uint32_t row0, uint32_t nrows,
                                                double amplitude, double frequency, int64_t seed,
                                                float* out_dev, float* minval, float* maxval) {
    REQUIRE(ctx, "ctx is NULL");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE((uint64_t)row0 + nrows <= height, "row band exceeds the image");
    REQUIRE(out_dev || nrows == 0, "out_dev is NULL");
    CK(cudaSetDevice(ctx->device));
    return run_simplex(ctx, width, height, row0, nrows, amplitude, frequency, seed, out_dev, minval,
                       maxval);
}

CLUMPY_API int clumpy_cuda_generate_simplex_octaves_dev(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                                        uint32_t row0, uint32_t nrows, uint32_t noctaves,
                                                        const double* amplitudes, const double* frequencies,
                                                        const int64_t* seeds, float* out_dev, float* minval, float* maxval) {
    REQUIRE(ctx && amplitudes && frequencies && seeds, "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    REQUIRE((uint64_t)row0 + nrows <= height, "row band exceeds the image");
    REQUIRE(out_dev || nrows == 0, "out_dev is NULL");
    REQUIRE(noctaves >= 1 && noctaves <= (uint32_t)SIMPLEX_MAX_OCTAVES, "1 .. 8 octaves");
    CK(cudaSetDevice(ctx->device));
    OctavePerms perms;
    OctaveArgs oa;
    oa.n = (int)noctaves;
    for (uint32_t o = 0; o < noctaves; ++o) {
        int16_t perm16[256];
        clumpy_cuda_simplex_perm(seeds[o], perm16);
        for (int i = 0; i < 256; ++i) perms.p[o][i] = (uint8_t)perm16[i];
        // generate_simplex.cc:59-67 and :77, per octave
        oa.step[o] = (width > height) ? (float)(frequencies[o] / height) : (float)(frequencies[o] / width);
        oa.scale[o] = (float)(amplitudes[o] / 47.0);
    }
    int rc = minmax_reset(ctx);
    if (rc) return rc;
    if (nrows > 0) {
        const int vec_ok = ((uintptr_t)out_dev % 16 == 0 && width % 4 == 0) ? 1 : 0;
        simplex_octaves_kernel<<<simplex_grid(ctx, width, nrows), SIMPLEX_THREADS, 0, ctx->stream>>>(
            perms, oa, width, row0, nrows, out_dev, ctx->d_minmax, vec_ok);
        CK_LAUNCH(ctx);
    }
    if (minval || maxval) return minmax_fetch(ctx, minval, maxval);
    return CLUMPY_OK;
}

CLUMPY_API int clumpy_cuda_generate_simplex(clumpy_cuda_ctx* ctx, uint32_t width, uint32_t height,
                                            double amplitude, double frequency, int64_t seed,
                                            float* out_host, float* minval, float* maxval) {
    REQUIRE(ctx && out_host, "null argument");
    REQUIRE(width > 0 && height > 0, "image dimensions must be positive");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)width * height * sizeof(float);
    float* d_out = nullptr;
    CK(cudaMalloc(&d_out, bytes));
    int rc = run_simplex(ctx, width, height, 0, height, amplitude, frequency, seed, d_out, nullptr,
                         nullptr);
    if (rc == CLUMPY_OK) {
        cudaError_t e = cudaMemcpyAsync(out_host, d_out, bytes, cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { set_error("D2H of the noise field failed: %s", cudaGetErrorString(e)); rc = CLUMPY_ERR_CUDA; }
    }
    if (rc == CLUMPY_OK && (minval || maxval)) rc = minmax_fetch(ctx, minval, maxval);
    cudaFree(d_out);
    return rc;
}
