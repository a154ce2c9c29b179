// generate_simplex -- same command name, arguments, messages and .npy output as the reference
// (commands/generate_simplex.cc:9-87 there); the per-pixel loop runs in clumpy_cuda_generate_simplex.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

namespace {

struct GenerateSimplex : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "generate simplex noise"; }
    string usage() const override { return "<dims> <amplitude> <frequency> <seed> <output_img>"; }
    string example() const override { return "400x200 1.0 16.0 26 out.npy"; }
};

ClumpyCommand::Register registrar("generate_simplex", [] { return new GenerateSimplex(); });

bool GenerateSimplex::exec(vector<string> args) {
    if (args.size() != 5) {
        printf("The command takes 5 arguments.\n");
        return false;
    }
    const string dims = args[0];
    const uint32_t width = atoi(dims.c_str());
    const uint32_t height = atoi(dims.substr(dims.find('x') + 1).c_str());
    const double amplitude = atof(args[1].c_str());
    const double frequency = atof(args[2].c_str());
    const int64_t seed = atoi(args[3].c_str());
    const string output_file = args[4];

    Gpu gpu;
    if (!gpu) return false;
    vector<float> result((size_t)width * height);
    float minval = 0, maxval = 0;
    if (!Gpu::report(clumpy_cuda_generate_simplex(gpu.ctx, width, height, amplitude, frequency, seed,
                                                  result.data(), &minval, &maxval)))
        return false;
    printf("Noise range is %g to %g\n", minval, maxval);
    return npy::save_f32(output_file, result.data(), {height, width});
}

} // namespace
