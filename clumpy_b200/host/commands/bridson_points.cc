// bridson_points -- same command name, arguments, messages and .npy output as the reference
// (commands/bridson_points.cc:19-182 there).
//
// Poisson-disk sampling by Bridson's algorithm.  This command stays on the HOST on purpose: every
// sample depends on all the samples before it (one active list, one stream of random numbers drawn
// from an integer hash of a running counter), so the reference's point list -- which advect_points
// and its golden frames are pinned to -- can only be reproduced by throwing the darts in the same
// order; a parallel Poisson-disk sampler is a different algorithm with a different output.  It is the
// producer in front of the accelerated path (README example5 / example6), runs once per pipeline, and
// takes ~10 ms for the 12 392 points of those examples.
#include "clumpy_command.hh"
#include "npy.hh"

#include <cmath>
#include <cstdio>
#include <cstdlib>

using std::string;
using std::vector;

namespace {

// the reference's integer hash (:60-66): 0, 1, 2, ... in, well-mixed 32-bit values out
uint32_t mix(uint32_t v) {
    v = (v ^ 12345391u) * 2654435769u;
    v ^= (v << 6) ^ (v >> 26);
    v *= 2654435769u;
    v += (v << 5) ^ (v >> 12);
    return v;
}

struct Darts {
    // Arithmetic follows the reference operation by operation, all in float (x86-64 SSE, no contraction;
    // this file is built with -ffp-contract=off): (float)UINT32_MAX is 2^32.
    static constexpr float kUnit = 1.0f / 4294967296.0f;
    int counter;
    explicit Darts(int seed) : counter(seed) {}
    uint32_t next() { return mix((uint32_t)counter++); }
    float below(float top) { return top * (float)next() / 4294967296.0f + 0.0f; } // randhashf(seed, 0, top) (:68-70)
    // a point with 1 < |r| <= 2 by rejection from [-2, 2)^2 (:72-86)
    void annulus(float& rx, float& ry) {
        for (;;) {
            rx = 4 * kUnit * (float)next() - 2;
            ry = 4 * kUnit * (float)next() - 2;
            const float len2 = rx * rx + ry * ry;
            if (len2 > 1 && len2 <= 4) return;
        }
    }
};

int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// generate_pts (:93-177): background grid of cell size radius / sqrt(2) holding at most one sample per
// cell, an active list, 30 candidates per pick.
vector<float> poisson_disk(float width, float height, float radius, int seed) {
    const int tries = 30;
    const float radius2 = radius * radius;
    const float inv_cell = 1.0f / (radius / sqrtf(2));
    const int ncols = (int)ceilf(width * inv_cell), nrows = (int)ceilf(height * inv_cell);
    const int ncells = ncols * nrows;
    vector<float> pts;
    if (ncells <= 0) return pts;
    pts.resize((size_t)ncells * 2);
    vector<int> cell(ncells, -1), active(ncells);
    int nactive = 0, n = 0;
    Darts rng(seed);
    auto cell_of = [&](float x, float y) -> int& { return cell[(int)(x * inv_cell) + ncols * (int)(y * inv_cell)]; };
    auto accept = [&](float x, float y) {
        cell_of(x, y) = active[nactive++] = n;
        pts[2 * (size_t)n] = x;
        pts[2 * (size_t)n + 1] = y;
        ++n;
    };
    {
        const float x = width * (float)rng.next() * Darts::kUnit;
        const float y = height * (float)rng.next() * Darts::kUnit;
        accept(x, y);
    }
    while (n < ncells) {
        const float pick = rng.below((float)nactive), last = (float)nactive - 1.0f;
        const int slot = (int)(last < pick ? last : pick);
        const int parent = active[slot];
        bool found = false;
        float x = 0, y = 0;
        for (int t = 0; t < tries && !found; ++t) {
            float rx, ry;
            rng.annulus(rx, ry);
            x = rx * radius + pts[2 * (size_t)parent];
            y = ry * radius + pts[2 * (size_t)parent + 1];
            if (x < 0 || x >= width || y < 0 || y >= height) continue;
            // cells within `radius` of the candidate; the reference keeps the bounds in floats (:136-144)
            const float x1 = (float)clampi((int)((x + radius) * inv_cell), 0, ncols - 1);
            const float y1 = (float)clampi((int)((y + radius) * inv_cell), 0, nrows - 1);
            const float x0 = (float)clampi((int)((x - radius) * inv_cell), 0, ncols - 1);
            const float y0 = (float)clampi((int)((y - radius) * inv_cell), 0, nrows - 1);
            bool clash = false;
            for (float cy = y0; cy <= y1 && !clash; cy++)
                for (float cx = x0; cx <= x1 && !clash; cx++) {
                    const int other = cell[(int)cy * ncols + (int)cx];
                    if (other < 0 || other == parent) continue;
                    const float dx = pts[2 * (size_t)other] - x, dy = pts[2 * (size_t)other + 1] - y;
                    if (dx * dx + dy * dy < radius2) clash = true;
                }
            found = !clash;
        }
        if (found) {
            accept(x, y);
        } else {
            if (--nactive <= 0) break;
            active[slot] = active[nactive];
        }
    }
    pts.resize((size_t)n * 2);
    return pts;
}

struct BridsonPoints : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "generate list of 2D points"; }
    string usage() const override { return "<dim> <minradius> <seed> <output_pts>"; }
    string example() const override { return "500x250 10 987 bridson.npy"; }
};

ClumpyCommand::Register registrar("bridson_points", [] { return new BridsonPoints(); });

bool BridsonPoints::exec(vector<string> args) {
    if (args.size() != 4) {
        printf("This command takes 4 arguments.\n");
        return false;
    }
    const string dims = args[0];
    const uint32_t width = atoi(dims.c_str());
    const uint32_t height = atoi(dims.substr(dims.find('x') + 1).c_str());
    const float minradius = atof(args[1].c_str());
    const int seed = atoi(args[2].c_str());
    const vector<float> pts = poisson_disk((float)width, (float)height, minradius, seed);
    const size_t npts = pts.size() / 2;
    printf("Generated %zu points.\n", npts);
    return npy::save_f32(args[3], pts.data(), {npts, 2});
}

} // namespace
