/*
 * clumpy_oracle.c -- CPU restatement of clumpy's point-advection hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under clumpy_b200/ may include, link or call this file.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.
 *
 * Parity status: PINNED.  The restatement is checked bit-for-bit against the reference compiled
 * from its own sources (oracle/Makefile -> oracle/_ref/clumpy_ref) on the fixtures in
 * tests/golden/ (see tests/golden/make_golden.py) and against the md5 hashes SURVEY.md 8(c)
 * records for README example5.  The reference itself ships no result-pinning tests.
 *
 * Every function cites the reference file:line it restates (paths relative to the clumpy root).
 * Arithmetic rules that matter (all verified against the compiled reference):
 *   - no FMA contraction anywhere (build with -ffp-contract=off, plain x86-64 SSE2);
 *   - float -> uint32_t goes through a 64-bit signed truncation, out of range / NaN -> 0;
 *   - float -> int32_t truncates, out of range / NaN -> INT32_MIN;
 *   - the framebuffer is uint8_t end to end: decay and every blend truncate to u8.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------
 * Conversions as gcc/x86-64 compiles them in the reference binary.
 * advect_points.cc:35 (float -> uint32_t argument of texel_fetch), splat_points.cc:110-111:
 *   cvttss2si r64 then the low 32 bits; "integer indefinite" (0x8000000000000000) when the value
 *   does not fit, whose low half is 0.
 * splat_points.cc:143-144 ((int32_t) cast): cvttss2si r32, indefinite = INT32_MIN.
 */
static uint32_t f2u32_x86(float f)
{
    if (!(f > -9223372036854775808.0f && f < 9223372036854775808.0f)) return 0u;
    return (uint32_t)(uint64_t)(int64_t)f;
}

static int32_t f2i32_x86(float f)
{
    if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
    return (int32_t)f;
}

/* int32 arithmetic with wrap-around, as the compiled reference behaves */
static int32_t add_wrap(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }

/* =============================================================================================
 * generate_simplex
 * ========================================================================================== */

/* generate_simplex.cc:241-280 -- permutation from a 64-bit LCG (three warm-up steps, then a
 * Fisher-Yates style draw from the shrinking source array, top index first). */
ORACLE_API void oracle_simplex_perm(int64_t seed, int16_t perm[256])
{
    const uint64_t mul = 6364136223846793005ULL, inc = 1442695040888963407ULL;
    int16_t pool[256];
    uint64_t s = (uint64_t)seed;
    for (int i = 0; i < 256; ++i) pool[i] = (int16_t)i;
    for (int w = 0; w < 3; ++w) s = s * mul + inc;
    for (int i = 255; i >= 0; --i) {
        s = s * mul + inc;
        int64_t t = (int64_t)(s + 31ULL);
        int r = (int)(t % (int64_t)(i + 1));
        if (r < 0) r += i + 1;
        perm[i] = pool[r];
        pool[r] = pool[i];
    }
}

/* generate_simplex.cc:124-126,162-168 -- the eight octagon gradients (+-5,+-2),(+-2,+-5);
 * pair p = hash>>1: bit0 swaps the magnitudes, bit1 negates x, bit2 negates y. */
static double lattice_gradient(const int16_t* perm, int xv, int yv, double dx, double dy)
{
    int hash = perm[(perm[xv & 0xFF] + yv) & 0xFF] & 0x0E;
    int p = hash >> 1;
    double gx = (p & 1) ? 2.0 : 5.0;
    double gy = (p & 1) ? 5.0 : 2.0;
    if (p & 2) gx = -gx;
    if (p & 4) gy = -gy;
    return gx * dx + gy * dy;
}

/* generate_simplex.cc:194-198 */
static int floor_to_int(double x)
{
    int xi = (int)x;
    return x < xi ? xi - 1 : xi;
}

/* One lattice vertex at integer offset (i, j) from the cell origin.  Its displacement from the
 * sample is (d0x - i) - c*SQUISH with c = i + j, which reproduces every per-vertex expression of
 * generate_simplex.cc:325-399 bit for bit (c*SQUISH is 0, SQUISH or the exact double 2*SQUISH).
 * Contribution: generate_simplex.cc:327-331 (attn = 2 - dx^2 - dy^2, fourth power times gradient). */
static double vertex_term(const int16_t* perm, int xsb, int ysb, double d0x, double d0y, int i, int j)
{
    const double squish = 0.366025403784439;
    double off = (double)(i + j) * squish;
    double dx = d0x - i - off;
    double dy = d0y - j - off;
    double attn = 2 - dx * dx - dy * dy;
    if (attn > 0) {
        attn *= attn;
        return attn * attn * lattice_gradient(perm, xsb + i, ysb + j, dx, dy);
    }
    return 0.0;
}

/* generate_simplex.cc:298-417 -- 2-D OpenSimplex.  Vertices (1,0) and (0,1) always contribute;
 * the third is (0,0) or (1,1) by which triangle of the rhombus holds the sample; the fourth
 * ("extra") vertex is picked from the region tests of :351-399.  Terms are summed in the
 * reference's order: (1,0), (0,1), third, extra; then / 47. */
ORACLE_API double oracle_simplex2(const int16_t* perm, double x, double y)
{
    const double stretch = -0.211324865405187;
    const double squish = 0.366025403784439;
    double so = (x + y) * stretch;
    double xs = x + so, ys = y + so;
    int xsb = floor_to_int(xs), ysb = floor_to_int(ys);
    double qo = (xsb + ysb) * squish;
    double xins = xs - xsb, yins = ys - ysb;
    double in_sum = xins + yins;
    double d0x = x - (xsb + qo), d0y = y - (ysb + qo);

    int ti, tj, ei, ej; /* third and extra vertex offsets */
    if (in_sum <= 1) {
        double zins = 1 - in_sum;
        ti = 0; tj = 0;
        if (zins > xins || zins > yins) {
            if (xins > yins) { ei = 1; ej = -1; } else { ei = -1; ej = 1; }
        } else { ei = 1; ej = 1; }
    } else {
        double zins = 2 - in_sum;
        ti = 1; tj = 1;
        if (zins < xins || zins < yins) {
            if (xins > yins) { ei = 2; ej = 0; } else { ei = 0; ej = 2; }
        } else { ei = 0; ej = 0; }
    }
    double value = 0;
    value += vertex_term(perm, xsb, ysb, d0x, d0y, 1, 0);
    value += vertex_term(perm, xsb, ysb, d0x, d0y, 0, 1);
    value += vertex_term(perm, xsb, ysb, d0x, d0y, ti, tj);
    value += vertex_term(perm, xsb, ysb, d0x, d0y, ei, ej);
    return value / 47.0;
}

/* generate_simplex.cc:59-82 -- float step = frequency / min(W,H) (double divide rounded to
 * float), coordinate = FP32 product step*index widened to double, sample = (float)(amp*noise). */
ORACLE_API void oracle_generate_simplex(uint32_t width, uint32_t height, double amplitude,
                                        double frequency, int64_t seed, float* out,
                                        float* minval, float* maxval)
{
    int16_t perm[256];
    oracle_simplex_perm(seed, perm);
    float step = (width > height) ? (float)(frequency / height) : (float)(frequency / width);
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    for (int row = 0; row < (int)height; ++row) {
        for (int col = 0; col < (int)width; ++col) {
            double u = (double)(step * (float)col);
            double v = (double)(step * (float)row);
            float s = (float)(amplitude * oracle_simplex2(perm, u, v));
            out[(size_t)row * width + col] = s;
            if (s < lo) lo = s;
            if (s > hi) hi = s;
        }
    }
    if (minval) *minval = lo;
    if (maxval) *maxval = hi;
}

/* Rows [row0, row0 + nrows) of the same image (the band a GPU of a row-sharded run computes; also what lets a
 * check at 16384 x 16384 look at a few rows without evaluating 268 M pixels on the CPU). */
ORACLE_API void oracle_generate_simplex_rows(uint32_t width, uint32_t height, uint32_t row0, uint32_t nrows,
                                             double amplitude, double frequency, int64_t seed, float* out)
{
    int16_t perm[256];
    oracle_simplex_perm(seed, perm);
    float step = (width > height) ? (float)(frequency / height) : (float)(frequency / width);
    for (uint32_t r = 0; r < nrows; ++r) {
        for (int col = 0; col < (int)width; ++col) {
            double u = (double)(step * (float)col);
            double v = (double)(step * (float)(int)(row0 + r));
            out[(size_t)r * width + col] = (float)(amplitude * oracle_simplex2(perm, u, v));
        }
    }
}

/* =============================================================================================
 * curl_2d
 * ========================================================================================== */

/* curl_2d.cc:57-71 -- forward differences with edge clamp; output texel = (dpdy, dpdx).
 * The printed range keeps the reference's quirk: min over dpdx only, max over both (:64-65). */
ORACLE_API void oracle_curl_2d(const float* src, uint32_t width, uint32_t height, float* dst,
                               float* minval, float* maxval)
{
    float lo = 3.402823466e+38f, hi = -3.402823466e+38f;
    for (uint32_t row = 0; row < height; ++row) {
        uint32_t nrow = (row < height - 1) ? row + 1 : row;
        for (uint32_t col = 0; col < width; ++col) {
            uint32_t ncol = (col < width - 1) ? col + 1 : col;
            float p = src[(size_t)row * width + col];
            float dpdx = src[(size_t)row * width + ncol] - p;
            float dpdy = p - src[(size_t)nrow * width + col];
            if (dpdx < lo) lo = dpdx;
            float m = dpdx > dpdy ? dpdx : dpdy;
            if (m > hi) hi = m;
            dst[2 * ((size_t)row * width + col) + 0] = dpdy;
            dst[2 * ((size_t)row * width + col) + 1] = dpdx;
        }
    }
    if (minval) *minval = lo;
    if (maxval) *maxval = hi;
}

/* =============================================================================================
 * splat_points / splat_disks
 * ========================================================================================== */

/* glm detail/func_common.inl:562-568 (smoothstep, float) */
static float smoothstep_f(float e0, float e1, float x)
{
    float t = (x - e0) / (e1 - e0);
    if (t < 0.0f) t = 0.0f;
    if (t > 1.0f) t = 1.0f;
    return t * t * (3.0f - 2.0f * t);
}

/* splat_points.cc:126-136 -- k x k anti-aliased sprite, sprite[i + j*k]. */
ORACLE_API void oracle_sprite(int kernel_size, float alpha, float* sprite)
{
    int middle = kernel_size / 2;
    float r2 = (float)(middle * middle);
    for (int i = 0; i < kernel_size; ++i)
        for (int j = 0; j < kernel_size; ++j) {
            int x = i - middle, y = j - middle;
            float d2 = (float)(x * x + y * y);
            sprite[i + j * kernel_size] = alpha * smoothstep_f(r2 + 5.0f, r2 - 5.0f, d2);
        }
}

/* splat_points.cc:153-155 -- one src-over hit with a white source, truncated to u8. */
static uint8_t blend_hit(uint8_t dst, float a)
{
    float v = (1.0f - a) * (float)(uint32_t)dst + a * 255.0f;
    return (uint8_t)(int32_t)v;
}

/* splat_points.cc:108-116 -- the k=1, alpha=1 fast path of the splat_points command. */
ORACLE_API void oracle_splat_points(const float* pts, uint32_t npts, uint32_t width,
                                    uint32_t height, uint8_t* img)
{
    for (uint32_t i = 0; i < npts; ++i) {
        uint32_t x = f2u32_x86(pts[2 * (size_t)i]);
        uint32_t y = f2u32_x86(pts[2 * (size_t)i + 1]);
        if (x < width && y < height) img[(size_t)width * y + x] = 255;
    }
}

/* splat_points.cc:118-161 -- points in index order, sprite rows outer / columns inner,
 * optional x wrap (advect_wrapx global, splat_points.cc:20), bounds test, truncating blend.
 * Returns -1 for an even kernel (the reference prints and exit(1)s, :122-125). */
ORACLE_API int oracle_splat_disks(const float* pts, uint32_t npts, uint32_t width, uint32_t height,
                                  uint8_t* img, float alpha, int kernel_size, int wrapx)
{
    if (kernel_size % 2 == 0) return -1;
    float* sprite = (float*)malloc(sizeof(float) * (size_t)kernel_size * kernel_size);
    oracle_sprite(kernel_size, alpha, sprite);
    const int32_t w = (int32_t)width, hgt = (int32_t)height, h = kernel_size / 2;
    for (uint32_t i = 0; i < npts; ++i) {
        int32_t x = f2i32_x86(pts[2 * (size_t)i]);
        int32_t y = f2i32_x86(pts[2 * (size_t)i + 1]);
        const float* sv = sprite;
        for (int32_t dy = -h; dy <= h; ++dy)
            for (int32_t dx = -h; dx <= h; ++dx, ++sv) {
                int32_t y0 = add_wrap(y, dy), x0 = add_wrap(x, dx);
                int32_t xx = x0;
                if (wrapx) xx = add_wrap(w, x0 % w) % w;
                if (xx >= 0 && y0 >= 0 && xx < w && y0 < hgt) {
                    size_t at = (size_t)w * (size_t)y0 + (size_t)xx;
                    img[at] = blend_hit(img[at], *sv);
                }
            }
    }
    free(sprite);
    return 0;
}

/* =============================================================================================
 * advect_points
 * ========================================================================================== */

/* std::mt19937 (public algorithm, Matsumoto & Nishimura) -- advect_points.cc:129-130 */
typedef struct { uint32_t mt[624]; int idx; } mt19937_t;

static void mt_seed(mt19937_t* g, uint32_t seed)
{
    g->mt[0] = seed;
    for (int i = 1; i < 624; ++i)
        g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
    g->idx = 624;
}

static uint32_t mt_next(mt19937_t* g)
{
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i] = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}

/* advect_points.cc:127-135 -- std::uniform_int_distribution<>(0, nframes-1) over mt19937(0) as
 * libstdc++ (>= 11, bits/uniform_int_dist.h) implements it for a 32-bit generator range:
 * Lemire's multiply-shift with rejection of the biased low products. */
ORACLE_API void oracle_age_offsets(uint32_t nframes, uint32_t npts, uint32_t* out)
{
    mt19937_t g;
    mt_seed(&g, 0u);
    const uint32_t range = nframes; /* (nframes-1) - 0 + 1 */
    for (uint32_t i = 0; i < npts; ++i) {
        uint64_t product = (uint64_t)mt_next(&g) * range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            uint32_t threshold = (0u - range) % range;
            while (low < threshold) {
                product = (uint64_t)mt_next(&g) * range;
                low = (uint32_t)product;
            }
        }
        out[i] = (uint32_t)(product >> 32);
    }
}

typedef struct {
    uint32_t npts, width, height, nframes, simframe;
    int kernel_size, wrapx;
    float step_size, decay;
    const float* velocities; /* borrowed, (H, W, 2) */
    float* orig;             /* owned copies */
    float* pts;
    float* age;
    uint32_t* age_offset;
    uint8_t* img;
} oracle_advect_t;

/* advect_points.cc:70-81,119-135 -- set-up.  A negative decay selects x-wrap mode (:78-81). */
ORACLE_API oracle_advect_t* oracle_advect_new(const float* pts, uint32_t npts, const float* vel,
                                              uint32_t width, uint32_t height, float step_size,
                                              int kernel_size, float decay, uint32_t nframes)
{
    oracle_advect_t* st = (oracle_advect_t*)calloc(1, sizeof(*st));
    st->npts = npts; st->width = width; st->height = height; st->nframes = nframes;
    st->kernel_size = kernel_size; st->step_size = step_size;
    st->wrapx = decay < 0;
    st->decay = decay < 0 ? -decay : decay;
    st->velocities = vel;
    st->orig = (float*)malloc(sizeof(float) * 2 * (size_t)npts + 8);
    st->pts = (float*)malloc(sizeof(float) * 2 * (size_t)npts + 8);
    memcpy(st->orig, pts, sizeof(float) * 2 * (size_t)npts);
    memcpy(st->pts, pts, sizeof(float) * 2 * (size_t)npts);
    st->age = (float*)calloc((size_t)npts + 1, sizeof(float));
    st->age_offset = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)npts + 1));
    oracle_age_offsets(nframes, npts, st->age_offset);
    st->img = (uint8_t*)calloc((size_t)width * height + 1, 1);
    return st;
}

ORACLE_API void oracle_advect_free(oracle_advect_t* st)
{
    if (!st) return;
    free(st->orig); free(st->pts); free(st->age); free(st->age_offset); free(st->img); free(st);
}

/* advect_points.cc:23-37 -- nearest-texel fetch; x wraps in uint32 when wrap mode is on;
 * outside the image the velocity is zero. */
static void texel_fetch(const oracle_advect_t* st, float fx, float fy, float* vx, float* vy)
{
    uint32_t x = f2u32_x86(fx), y = f2u32_x86(fy);
    if (st->wrapx) x = (st->width + (x % st->width)) % st->width;
    if (x >= st->width || y >= st->height) { *vx = 0.0f; *vy = 0.0f; return; }
    const float* t = st->velocities + 2 * ((size_t)x + (size_t)st->width * y);
    *vx = t[0]; *vy = t[1];
}

/* advect_points.cc:139-179 -- ONE simulation frame: Euler step of every point (multiply and add
 * rounded separately), age / reset bookkeeping exactly as written (:144-153 warm-up, :163-171
 * recorded), u8 decay of the whole framebuffer, splat.  Returns 1 if this frame is one the
 * reference writes to disk (the second nframes frames), 0 for a warm-up frame, -1 when done. */
ORACLE_API int oracle_advect_step(oracle_advect_t* st)
{
    const uint32_t s = st->simframe;
    if (s >= 2u * st->nframes) return -1;
    const int warm = s < st->nframes;
    for (uint32_t i = 0; i < st->npts; ++i) {
        float* p = st->pts + 2 * (size_t)i;
        float vx, vy;
        texel_fetch(st, p[0], p[1], &vx, &vy);
        float mx = st->step_size * vx, my = st->step_size * vy;
        p[0] = p[0] + mx;
        p[1] = p[1] + my;
        st->age[i] += 1.0f;
        if (warm) {
            if (s >= st->age_offset[i]) {
                p[0] = st->orig[2 * (size_t)i]; p[1] = st->orig[2 * (size_t)i + 1];
                st->age[i] = 0.0f;
                st->age_offset[i] = st->nframes;
            }
        } else if (st->age[i] >= (float)st->nframes) {
            p[0] = st->orig[2 * (size_t)i]; p[1] = st->orig[2 * (size_t)i + 1];
            st->age[i] = 0.0f;
        }
    }
    st->simframe = s + 1;
    if (warm && st->decay == 0.0f) return 0; /* :154 -- no decay, no splat */
    const size_t npix = (size_t)st->width * st->height;
    for (size_t k = 0; k < npix; ++k) {
        float v = (float)(int32_t)st->img[k] * st->decay;
        st->img[k] = (uint8_t)(int32_t)v; /* `v *= decay` on a uint8_t& (:155,174) */
    }
    oracle_splat_disks(st->pts, st->npts, st->width, st->height, st->img, 1.0f, st->kernel_size,
                       st->wrapx);
    return warm ? 0 : 1;
}

ORACLE_API const float* oracle_advect_points(const oracle_advect_t* st) { return st->pts; }
ORACLE_API const uint8_t* oracle_advect_image(const oracle_advect_t* st) { return st->img; }
ORACLE_API uint32_t oracle_advect_simframe(const oracle_advect_t* st) { return st->simframe; }

/* =============================================================================================
 * The producers either side of the path (SURVEY.md 8f "next": N1 bridson_points, N2 pendulum_phase,
 * N4 cull_points).  Pinned like the rest: bit-identical to oracle/_ref/clumpy_ref on the fixtures of
 * tests/golden/next_small.npz and on the C1 / C2 point lists and the C2 field (md5 in golden.json).
 * ========================================================================================== */

/* bridson_points.cc:60-66 -- integer hash that turns 0,1,2,... into random numbers */
static uint32_t randhash(uint32_t seed)
{
    uint32_t i = (seed ^ 12345391u) * 2654435769u;
    i ^= (i << 6) ^ (i >> 26);
    i *= 2654435769u;
    i += (i << 5) ^ (i >> 12);
    return i;
}

/* bridson_points.cc:68-70: every operand is float; (float)UINT32_MAX rounds to 2^32 */
static float randhashf(uint32_t seed, float a, float b)
{
    return (b - a) * (float)randhash(seed) / 4294967296.0f + a;
}

/* bridson_points.cc:72-86: a point of the annulus 1 < |r| <= 2 by rejection, scaled and moved */
static void sample_annulus(float radius, float cx, float cy, int* seedptr, float* ox, float* oy)
{
    uint32_t seed = (uint32_t)*seedptr;
    const float rscale = 1.0f / 4294967296.0f; /* 1.0f / UINT_MAX, the int converted to float */
    float rx, ry;
    for (;;) {
        rx = 4 * rscale * (float)randhash(seed++) - 2;
        ry = 4 * rscale * (float)randhash(seed++) - 2;
        const float r2 = rx * rx + ry * ry;
        if (r2 > 1 && r2 <= 4) break;
    }
    *seedptr = (int)seed;
    *ox = rx * radius + cx;
    *oy = ry * radius + cy;
}

static int clamp_i(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* bridson_points.cc:93-177 (generate_pts): Bridson's Poisson-disk sampling on a background grid of
 * cell size r / sqrt(2), 30 attempts per active sample, all randomness from randhash(seed++).
 * `out` must hold 2 * ceil(w / cell) * ceil(h / cell) floats (oracle_bridson_capacity); returns the
 * number of points. */
ORACLE_API uint32_t oracle_bridson_capacity(uint32_t w, uint32_t h, float radius)
{
    const float cellsize = radius / sqrtf(2);
    const float invcell = 1.0f / cellsize;
    const int ncols = (int)ceilf((float)w * invcell), nrows = (int)ceilf((float)h * invcell);
    return (uint32_t)(ncols * nrows);
}

ORACLE_API uint32_t oracle_bridson_points(uint32_t w, uint32_t h, float radius, int seed, float* out)
{
    const float width = (float)w, height = (float)h;
    const int maxattempts = 30;
    const float rscale = 1.0f / 4294967296.0f;
    const float r2 = radius * radius;
    const float cellsize = radius / sqrtf(2);
    const float invcell = 1.0f / cellsize;
    const int ncols = (int)ceilf(width * invcell), nrows = (int)ceilf(height * invcell);
    const int maxcol = ncols - 1, maxrow = nrows - 1, ncells = ncols * nrows;
    if (ncells <= 0) return 0;
    int* grid = (int*)malloc((size_t)ncells * sizeof(int));
    int* actives = (int*)malloc((size_t)ncells * sizeof(int));
    for (int i = 0; i < ncells; ++i) grid[i] = -1;
    int nactives = 0, nsamples = 0;
    float px = width * (float)randhash((uint32_t)seed++) * rscale;
    float py = height * (float)randhash((uint32_t)seed++) * rscale;
    grid[(int)(px * invcell) + ncols * (int)(py * invcell)] = actives[nactives++] = nsamples;
    out[2 * nsamples] = px; out[2 * nsamples + 1] = py; ++nsamples;
    while (nsamples < ncells) {
        const float pick = randhashf((uint32_t)seed++, 0, (float)nactives);
        const float cap = (float)nactives - 1.0f;
        const int aindex = (int)(cap < pick ? cap : pick); /* min(a, b) = b < a ? b : a */
        const int sindex = actives[aindex];
        int found = 0;
        for (int attempt = 0; attempt < maxattempts && !found; ++attempt) {
            sample_annulus(radius, out[2 * sindex], out[2 * sindex + 1], &seed, &px, &py);
            if (px < 0 || px >= width || py < 0 || py >= height) continue;
            /* the window of grid cells within `radius`; its bounds pass through float on the way */
            const float maxjx = (float)clamp_i((int)((px + radius) * invcell), 0, maxcol);
            const float maxjy = (float)clamp_i((int)((py + radius) * invcell), 0, maxrow);
            const float minjx = (float)clamp_i((int)((px - radius) * invcell), 0, maxcol);
            const float minjy = (float)clamp_i((int)((py - radius) * invcell), 0, maxrow);
            int reject = 0;
            for (float jy = minjy; jy <= maxjy && !reject; jy++)
                for (float jx = minjx; jx <= maxjx && !reject; jx++) {
                    const int entry = grid[(int)jy * ncols + (int)jx];
                    if (entry > -1 && entry != sindex) {
                        const float dx = out[2 * entry] - px, dy = out[2 * entry + 1] - py;
                        if (dx * dx + dy * dy < r2) reject = 1;
                    }
                }
            if (reject) continue;
            found = 1;
        }
        if (found) {
            grid[(int)(px * invcell) + ncols * (int)(py * invcell)] = actives[nactives++] = nsamples;
            out[2 * nsamples] = px; out[2 * nsamples + 1] = py; ++nsamples;
        } else {
            if (--nactives <= 0) break;
            actives[aindex] = actives[nactives];
        }
    }
    free(grid);
    free(actives);
    return (uint32_t)nsamples;
}

/* pendulum_phase.cc:38-71: out[row][col] = (omega, -friction * omega - g / L * sin(theta)) with
 * theta = sx * (col / W - 0.5), omega = sy * (row / H - 0.5); the subtraction of 0.5 and the product
 * with the scale run in double and round to float, sin is libm's sinf (the call resolves to it in the
 * compiled reference), sx = (float)(scale_x * M_PI / 0.5). */
ORACLE_API void oracle_pendulum_phase(uint32_t width, uint32_t height, float friction, float scalex,
                                      float scaley, float* out)
{
    const float g = 1.0f, L = 1.0f;
    const float gsx = (float)((double)scalex * M_PI / 0.5), gsy = scaley * 1;
    for (uint32_t row = 0; row < height; ++row)
        for (uint32_t col = 0; col < width; ++col) {
            const float x = (float)((double)gsx * ((double)((float)col / (float)width) - 0.5));
            const float y = (float)((double)gsy * ((double)((float)row / (float)height) - 0.5));
            const float omega_dot = -friction * y - g / L * sinf(x);
            out[2 * ((size_t)row * width + col)] = y;
            out[2 * ((size_t)row * width + col) + 1] = omega_dot;
        }
}

/* cull_points.cc:97-111: keep the points whose nearest mask texel (float -> uint32_t like texel_fetch,
 * outside the image reads 0) is positive, in their original order; returns how many were kept. */
ORACLE_API uint32_t oracle_cull_points(const float* pts, uint32_t npts, const float* mask,
                                       uint32_t width, uint32_t height, float* out)
{
    uint32_t kept = 0;
    for (uint32_t i = 0; i < npts; ++i) {
        const uint32_t x = f2u32_x86(pts[2 * (size_t)i]), y = f2u32_x86(pts[2 * (size_t)i + 1]);
        const float m = (x >= width || y >= height) ? 0.0f : mask[x + (size_t)width * y];
        if (m > 0) {
            out[2 * (size_t)kept] = pts[2 * (size_t)i];
            out[2 * (size_t)kept + 1] = pts[2 * (size_t)i + 1];
            ++kept;
        }
    }
    return kept;
}
