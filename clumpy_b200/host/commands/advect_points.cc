// advect_points -- same command name, arguments, messages, frame naming and .npy output as the
// reference (commands/advect_points.cc:39-183 there).  Argument parsing, input validation and the file writes
// stay on the host here; the mt19937 reset offsets (:127-135) are drawn on a host thread inside the library
// while it uploads (clumpy_cuda_age_offsets: the same values as <random>'s); the 2*nframes simulation frames
// (Euler step, reset, u8 decay, splat) run in clumpy_cuda_advect_points, which hands back each
// recorded framebuffer from pinned memory while the GPU simulates the next one.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

extern bool advect_wrapx;

namespace {

struct AdvectPoints : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "create an animation by moving points according to a velocity field"; }
    string usage() const override {
        return "<input_pts> <velocities_img> <step_size> <kernel_size> <decay> <nframes> <suffix_img>";
    }
    string example() const override { return "coords.npy speeds.npy 1.0 3 0.5 240 anim.npy"; }
};

ClumpyCommand::Register registrar("advect_points", [] { return new AdvectPoints(); });

struct FrameSink {
    string suffix;
    uint32_t width, height, total, written = 0;
    bool failed = false;
};

void show_progress(uint32_t i, uint32_t count) {
    const int progress = count > 1 ? (int)(100ull * i / (count - 1)) : 100;
    char bar[104];
    bar[0] = '[';
    for (int c = 0; c < 100; ++c) bar[1 + c] = c < progress ? '=' : '-';
    bar[101] = ']'; bar[102] = '\r'; bar[103] = 0;
    fputs(bar, stdout);
    fflush(stdout);
}

int write_frame(void* user, uint32_t animframe, const uint8_t* pixels) {
    FrameSink* sink = static_cast<FrameSink*>(user);
    char name[32];
    snprintf(name, sizeof(name), "%03u", animframe);
    if (!npy::save_u8(name + sink->suffix, pixels, {sink->height, sink->width})) {
        sink->failed = true;
        return 1;
    }
    sink->written = animframe + 1;
    // the reference redraws its bar once per simulation frame; the recorded half is the second half
    show_progress(sink->total + animframe, 2 * sink->total);
    return 0;
}

bool AdvectPoints::exec(vector<string> args) {
    if (args.size() != 7) {
        printf("This command takes 7 arguments.\n");
        return false;
    }
    const float step_size = atof(args[2].c_str());
    const int kernel_size = atoi(args[3].c_str());
    float decay = atof(args[4].c_str());
    const uint32_t nframes = atoi(args[5].c_str());
    const string suffix = args[6];
    if (decay < 0) { // a negative decay selects x-wrap mode
        advect_wrapx = true;
        decay = -decay;
    }
    const npy::Array pts = npy::load(args[0]);
    if (pts.shape.size() != 2 || pts.shape[1] != 2) {
        printf("Input points have wrong shape.\n");
        return false;
    }
    if (pts.word_size != sizeof(float) || pts.type_code != 'f') {
        printf("Input points has wrong data type.\n");
        return false;
    }
    const npy::Array vel = npy::load(args[1]);
    if (vel.shape.size() != 3 || vel.shape[2] != 2) {
        printf("Velocities have wrong shape.\n");
        return false;
    }
    if (vel.word_size != sizeof(float) || vel.type_code != 'f') {
        printf("Velocities have wrong data type.\n");
        return false;
    }
    const uint32_t npts = pts.shape[0], height = vel.shape[0], width = vel.shape[1];
    if (nframes == 0) { // (the reference divides by zero in its progress bar here)
        printf("\nGenerated %03d%s through %03u%s.\n", 0, suffix.c_str(), 0u - 1u, suffix.c_str());
        return true;
    }
    if (0 == (kernel_size % 2)) {
        // the reference reaches this inside splat_disks, in its first simulation frame, and exits 1
        // (splat_points.cc:122-125): the first progress bar is already out by then
        show_progress(0, 2 * nframes);
        printf("Kernel size must be an odd integer.\n");
        return false;
    }

    const vector<int> devices = clumpy_devices();
    if (!devices.empty()) { // several GPUs of this box, driven from this process
        FrameSink sink{suffix, width, height, nframes};
        for (uint32_t s = 0; s < nframes; ++s) show_progress(s, 2 * nframes);
        const int rc = clumpy_cuda_advect_points_multi(devices.data(), (int)devices.size(), pts.data<float>(), npts, vel.data<float>(),
                                                       width, height, step_size, kernel_size, decay, advect_wrapx ? 1 : 0, nframes,
                                                       /*age_offset=*/nullptr, write_frame, &sink);
        if (sink.failed) {
            printf("\nCould not write frame %03u%s\n", sink.written, suffix.c_str());
            return false;
        }
        if (!Gpu::report(rc)) return false;
        printf("\nGenerated %03d%s through %03u%s.\n", 0, suffix.c_str(), nframes - 1, suffix.c_str());
        return true;
    }
    Gpu gpu;
    if (!gpu) return false;
    FrameSink sink{suffix, width, height, nframes};
    // The reference redraws its bar once per simulation frame (advect_points.cc:140).  The warm-up half runs
    // on the device in one go here (milliseconds), so its bars are printed up front; the recorded half's are
    // printed as the frames arrive (write_frame): stdout is byte-identical to the reference's.
    for (uint32_t s = 0; s < nframes; ++s) show_progress(s, 2 * nframes);
    const int rc = clumpy_cuda_advect_points(gpu.ctx, pts.data<float>(), npts, vel.data<float>(), width, height,
                                             step_size, kernel_size, decay, advect_wrapx ? 1 : 0, nframes,
                                             /*age_offset=*/nullptr, write_frame, &sink);
    if (sink.failed) {
        printf("\nCould not write frame %03u%s\n", sink.written, suffix.c_str());
        return false;
    }
    if (!Gpu::report(rc)) return false;
    printf("\nGenerated %03d%s through %03u%s.\n", 0, suffix.c_str(), nframes - 1, suffix.c_str());
    return true;
}

} // namespace
