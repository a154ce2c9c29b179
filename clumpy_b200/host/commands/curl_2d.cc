// curl_2d -- same command name, arguments, messages and .npy output as the reference
// (commands/curl_2d.cc:11-77 there); the stencil runs in clumpy_cuda_curl_2d.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

namespace {

struct Curl2d : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "apply the curl operator to a field of scalars"; }
    string usage() const override { return "<input_img> <output_img>"; }
    string example() const override { return "in.npy out.npy"; }
};

ClumpyCommand::Register registrar("curl_2d", [] { return new Curl2d(); });

bool Curl2d::exec(vector<string> args) {
    if (args.size() != 2) {
        printf("Wrong number of arguments.\n");
        return false;
    }
    const npy::Array arr = npy::load(args[0]);
    if (arr.shape.size() != 2) {
        printf("Input data has wrong shape.\n");
        return false;
    }
    if (arr.word_size != sizeof(float) || arr.type_code != 'f') {
        printf("Input data has wrong data type.\n");
        return false;
    }
    const uint32_t height = arr.shape[0], width = arr.shape[1];
    Gpu gpu;
    if (!gpu) return false;
    vector<float> result((size_t)width * height * 2);
    float minval = 0, maxval = 0;
    if (!Gpu::report(clumpy_cuda_curl_2d(gpu.ctx, arr.data<float>(), width, height, result.data(),
                                         &minval, &maxval)))
        return false;
    printf("Curl range is %g to %g\n", minval, maxval);
    printf("Curl shape is %ux%ux2\n", width, height);
    return npy::save_f32(args[1], result.data(), {height, width, 2});
}

} // namespace
