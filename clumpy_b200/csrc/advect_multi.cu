// advect_multi.cu -- the whole `advect_points` command (commands/advect_points.cc:83-179) on SEVERAL GPUs of one
// box from ONE process: the C++ driver behind `clumpy advect_points` when CLUMPY_DEVICES names more than one device.
//
// It is written against the library's own C ABI (include/clumpy_cuda.h), exactly as an application would be:
//   * one context per device, peer access enabled between every pair (the inboxes are plain allocations, so a
//     peer's pointer is usable as it is: no IPC handles inside one process);
//   * points sharded by HOME ROW into strips of equal point count (clumpy_cuda_row_histogram on device 0, the
//     shards themselves selected on each device from the whole list: desc.shard_row_lo / _hi), reset offsets drawn
//     once (advect_points.cc:127-135) and shared, velocity field replicated;
//   * frames run in chunks: every rank enqueues its route / drain / composite launches for the chunk
//     (clumpy_cuda_advect_route_record: nothing blocks, the ranks synchronise with each other on the device through
//     the epoch flags in peer memory) and copies its row band of every recorded frame over its own PCIe link into
//     pinned memory; the host then stitches the bands into whole frames and hands them to the caller in order while
//     the GPUs run the next chunk.
//
// The same device may be listed several times (tests: the ranks then share a GPU, like tests/test_gpu_routed.py).
#include "internal.h"

#include <algorithm>
#include <cstring>
#include <thread>
#include <vector>

using namespace clumpy;

namespace {

struct Rank {
    clumpy_cuda_ctx* ctx = nullptr;
    clumpy_advect* adv = nullptr;
    void* inbox = nullptr;
    uint32_t row0 = 0, nrows = 0;
    uint8_t* pinned[2] = {nullptr, nullptr}; // two chunks of band frames
    cudaEvent_t landed[2] = {nullptr, nullptr};
    size_t pinned_bytes = 0;
    int rc = CLUMPY_OK;
    std::string error;
};

// strips of equal point count: edges[r] = first row of strip r (clumpy_b200/multigpu.py strip_edges)
std::vector<uint32_t> strip_edges(const std::vector<uint32_t>& rows, int world) {
    const uint32_t height = (uint32_t)rows.size();
    uint64_t n = 0;
    for (uint32_t c : rows) n += c;
    std::vector<uint64_t> below(height);
    uint64_t acc = 0;
    for (uint32_t y = 0; y < height; ++y) { acc += rows[y]; below[y] = acc; }
    std::vector<uint32_t> edges(world + 1, height);
    edges[0] = 0;
    for (int r = 1; r < world; ++r) {
        const uint64_t want = n * (uint64_t)r / (uint64_t)world;
        edges[r] = (uint32_t)(std::upper_bound(below.begin(), below.end(), want) - below.begin());
    }
    return edges;
}

} // namespace

CLUMPY_API int clumpy_cuda_advect_points_multi(const int* devices, int ndevices, const float* pts_host, uint32_t npts,
                                               const float* vel_host, uint32_t width, uint32_t height,
                                               float step_size, int kernel_size, float decay, int wrapx,
                                               uint32_t nframes, const uint32_t* age_offset,
                                               clumpy_frame_fn on_frame, void* user) {
    REQUIRE(devices && ndevices >= 1 && ndevices <= 8, "1 .. 8 devices");
    REQUIRE(on_frame && vel_host && (pts_host || npts == 0), "null argument");
    REQUIRE(width > 0 && height > 0 && nframes > 0, "bad dimensions");
    const int world = ndevices;
    std::vector<Rank> ranks(world);
    int rc = CLUMPY_OK;
    auto cleanup = [&]() {
        for (auto& r : ranks) {
            if (r.adv) clumpy_cuda_advect_end(r.adv);
            r.adv = nullptr;
        }
        for (auto& r : ranks) {
            for (uint8_t* p : r.pinned) if (p && r.ctx) ctx_pinned_put(r.ctx, p, r.pinned_bytes);
            for (cudaEvent_t e : r.landed) if (e) cudaEventDestroy(e);
            if (r.ctx) clumpy_cuda_destroy(r.ctx);
            r.ctx = nullptr;
        }
    };
    for (int r = 0; r < world && rc == CLUMPY_OK; ++r) rc = clumpy_cuda_create(devices[r], nullptr, &ranks[r].ctx);
    if (rc) { cleanup(); return rc; }
    // every GPU stores into every other GPU's inbox
    for (int a = 0; a < world; ++a)
        for (int b = 0; b < world; ++b) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, devices[a], devices[b]);
            if (!can) { cleanup(); set_error("device %d cannot access device %d's memory (no NVLink / P2P)", devices[a], devices[b]); return CLUMPY_ERR_CUDA; }
            cudaSetDevice(devices[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { cleanup(); set_error("cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e)); return CLUMPY_ERR_CUDA; }
            cudaGetLastError();
        }
    // reset offsets (they follow the ORIGINAL point index): every rank draws the whole list on its own device when
    // the caller has none (age_offset == NULL goes through to clumpy_cuda_advect_begin_ex)
    // home-row strips of equal point count
    std::vector<uint32_t> rows(height, 0);
    rc = clumpy_cuda_row_histogram(ranks[0].ctx, pts_host, npts, 0, height, rows.data());
    if (rc) { cleanup(); return rc; }
    const std::vector<uint32_t> edges = strip_edges(rows, world);
    uint32_t capacity = 1;
    for (int r = 0; r < world; ++r) {
        uint64_t mine = 0;
        for (uint32_t y = edges[r]; y < edges[r + 1]; ++y) mine += rows[y];
        capacity = std::max<uint32_t>(capacity, (uint32_t)mine);
    }
    // per-rank set-up in parallel (each uploads over its own PCIe link and sorts its shard)
    {
        std::vector<std::thread> workers;
        for (int r = 0; r < world; ++r)
            workers.emplace_back([&, r] {
                Rank& k = ranks[r];
                clumpy_advect_desc d;
                memset(&d, 0, sizeof(d));
                d.size = sizeof(d);
                d.pts = pts_host; d.npts = npts; d.vel = vel_host; d.width = width; d.height = height;
                d.step_size = step_size; d.kernel_size = kernel_size; d.decay = decay; d.wrapx = wrapx; d.nframes = nframes;
                d.age_offset = age_offset;
                d.shard_row_lo = edges[r]; d.shard_row_hi = edges[r + 1];
                if (d.shard_row_hi <= d.shard_row_lo) { d.shard_row_lo = height; d.shard_row_hi = height + 1; } // an empty strip
                d.cells = CLUMPY_CELLS_COUNT;
                k.rc = clumpy_cuda_advect_begin_ex(k.ctx, &d, &k.adv);
                size_t ibytes = 0;
                if (k.rc == CLUMPY_OK) k.rc = clumpy_cuda_advect_route_begin(k.adv, r, world, capacity, &k.inbox, &ibytes);
                if (k.rc == CLUMPY_OK) k.rc = clumpy_cuda_advect_owned_rows(k.adv, &k.row0, &k.nrows);
                if (k.rc) k.error = clumpy_cuda_last_error(); // (the message is per thread)
            });
        for (auto& t : workers) t.join();
    }
    for (auto& k : ranks)
        if (k.rc) { rc = k.rc; set_error("%s", k.error.c_str()); break; }
    std::vector<void*> inboxes(world);
    for (int r = 0; r < world; ++r) inboxes[r] = ranks[r].inbox;
    for (int r = 0; r < world && rc == CLUMPY_OK; ++r) rc = clumpy_cuda_advect_route_peers(ranks[r].adv, inboxes.data());
    // recorded frames travel in chunks of CHUNK frames per rank, double-buffered
    constexpr uint32_t CHUNK = 8;
    const size_t npix = (size_t)width * height;
    for (auto& k : ranks) {
        if (rc) break;
        k.pinned_bytes = std::max<size_t>((size_t)CHUNK * k.nrows * width, 1);
        for (int b = 0; b < 2; ++b) {
            k.pinned[b] = static_cast<uint8_t*>(ctx_pinned_get(k.ctx, k.pinned_bytes));
            if (!k.pinned[b]) { rc = CLUMPY_ERR_NOMEM; break; }
            cudaSetDevice(k.ctx->device);
            if (cudaEventCreateWithFlags(&k.landed[b], cudaEventDisableTiming) != cudaSuccess) { set_error("event creation failed"); rc = CLUMPY_ERR_CUDA; break; }
        }
    }
    std::vector<uint8_t> frame(rc == CLUMPY_OK ? npix : 0);
    // warm-up half (advect_points.cc:139-158).  Nothing here waits on the host, the GPUs pace each other through
    // the epoch flags -- but a rank's drain kernels wait (bounded) for launches of the OTHER ranks, which this same
    // thread has yet to enqueue, and a stream only queues so much before a launch call blocks: enqueue in turns.
    constexpr uint32_t TURN = 64;
    for (uint32_t s = 0; s < nframes && rc == CLUMPY_OK; s += TURN)
        for (int r = 0; r < world && rc == CLUMPY_OK; ++r) rc = clumpy_cuda_advect_route_steps(ranks[r].adv, std::min(TURN, nframes - s));
    // recorded half
    auto deliver = [&](uint32_t first, uint32_t count, int buf) -> int {
        for (auto& k : ranks)
            if (cudaEventSynchronize(k.landed[buf]) != cudaSuccess) { set_error("frame copy failed"); return CLUMPY_ERR_CUDA; }
        for (uint32_t i = 0; i < count; ++i) {
            for (auto& k : ranks)
                if (k.nrows) memcpy(frame.data() + (size_t)k.row0 * width, k.pinned[buf] + (size_t)i * k.nrows * width, (size_t)k.nrows * width);
            if (on_frame(user, first + i, frame.data()) != 0) { set_error("frame callback aborted the run"); return CLUMPY_ERR_STATE; }
        }
        return CLUMPY_OK;
    };
    uint32_t issued = 0, pending_first = 0, pending_count = 0;
    int buf = 0;
    while (rc == CLUMPY_OK && issued < nframes) {
        const uint32_t count = std::min(CHUNK, nframes - issued);
        for (int r = 0; r < world && rc == CLUMPY_OK; ++r) {
            Rank& k = ranks[r];
            rc = clumpy_cuda_advect_route_record(k.adv, count, k.pinned[buf], std::max<size_t>((size_t)k.nrows * width, 1));
            if (rc == CLUMPY_OK && (cudaSetDevice(k.ctx->device) != cudaSuccess || cudaEventRecord(k.landed[buf], k.ctx->copy_stream) != cudaSuccess)) {
                set_error("event record failed");
                rc = CLUMPY_ERR_CUDA;
            }
        }
        // while the GPUs run this chunk, hand out the one before
        if (rc == CLUMPY_OK && pending_count) rc = deliver(pending_first, pending_count, buf ^ 1);
        pending_first = issued; pending_count = count;
        issued += count;
        buf ^= 1;
    }
    if (rc == CLUMPY_OK && pending_count) rc = deliver(pending_first, pending_count, buf ^ 1);
    for (auto& k : ranks) // a fence time-out on any rank fails the command (the frames would be incomplete)
        if (rc == CLUMPY_OK && k.adv) rc = clumpy_cuda_advect_wait_frames(k.adv);
    std::string keep = rc ? clumpy_cuda_last_error() : "";
    cleanup();
    if (rc) set_error("%s", keep.c_str());
    return rc;
}
