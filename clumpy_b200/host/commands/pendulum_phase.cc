// pendulum_phase -- same command name, arguments, messages and .npy output as the reference
// (commands/pendulum_phase.cc:19-71 there); the field is filled by clumpy_cuda_pendulum_phase.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

namespace {

struct PendulumPhase : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "generate a θ-ω field of 2D vectors"; }
    // the reference advertises an <image_type> argument that exec() never took (:27 vs :39); kept verbatim
    string usage() const override { return "<image_type> <dims> <friction> <scale_x> <scale_y> <output_img>"; }
    string example() const override { return "500x500 0.01 1.0 5.0 field.npy"; }
};

ClumpyCommand::Register registrar("pendulum_phase", [] { return new PendulumPhase(); });

bool PendulumPhase::exec(vector<string> args) {
    if (args.size() != 5) {
        printf("The command takes 5 arguments.\n");
        return false;
    }
    const string dims = args[0];
    const float friction = atof(args[1].c_str());
    const float scalex = atof(args[2].c_str());
    const float scaley = atof(args[3].c_str());
    const uint32_t width = atoi(dims.c_str());
    const uint32_t height = atoi(dims.substr(dims.find('x') + 1).c_str());
    Gpu gpu;
    if (!gpu) return false;
    vector<float> field((size_t)width * height * 2);
    if (!Gpu::report(clumpy_cuda_pendulum_phase(gpu.ctx, width, height, friction, scalex, scaley, field.data())))
        return false;
    return npy::save_f32(args[4], field.data(), {height, width, 2});
}

} // namespace
