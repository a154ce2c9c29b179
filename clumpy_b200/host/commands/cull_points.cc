// cull_points -- same command name, arguments, messages and .npy output as the reference
// (commands/cull_points.cc:29-111 there); the filter runs in clumpy_cuda_cull_points.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

namespace {

struct CullPoints : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "remove points from areas with non-positive pixels"; }
    string usage() const override { return "<input_pts> <mask_img> <output_pts>"; }
    string example() const override { return "bridson.npy shapes.npy culled.npy"; }
};

ClumpyCommand::Register registrar("cull_points", [] { return new CullPoints(); });

bool CullPoints::exec(vector<string> args) {
    if (args.size() != 3) {
        printf("This command takes 3 arguments.\n");
        return false;
    }
    const npy::Array arr = npy::load(args[0]);
    if (arr.shape.size() != 2) {
        printf("Input data has wrong shape.\n");
        return false;
    }
    if (arr.shape[1] != 2) {
        printf("Input points must be 2D.\n");
        return false;
    }
    if (arr.word_size != sizeof(float) || arr.type_code != 'f') {
        printf("Input data has wrong data type.\n");
        return false;
    }
    const npy::Array img = npy::load(args[1]);
    if (img.shape.size() != 2) {
        printf("Input data has wrong shape.\n");
        return false;
    }
    if (img.word_size != sizeof(float) || img.type_code != 'f') {
        printf("Input data has wrong data type.\n");
        return false;
    }
    const uint32_t npts = (uint32_t)arr.shape[0];
    const uint32_t height = (uint32_t)img.shape[0], width = (uint32_t)img.shape[1];
    Gpu gpu;
    if (!gpu) return false;
    vector<float> result((size_t)std::max<uint32_t>(npts, 1u) * 2);
    uint32_t kept = 0;
    if (!Gpu::report(clumpy_cuda_cull_points(gpu.ctx, arr.data<float>(), npts, img.data<float>(), width, height,
                                             result.data(), &kept)))
        return false;
    printf("Removed %u points.\n", npts - kept);
    return npy::save_f32(args[2], result.data(), {kept, 2});
}

} // namespace
