// splat_points -- same command name, arguments, messages and .npy output as the reference
// (commands/splat_points.cc:24-104 there).  The two rasterisers it owns there, splat_points() and
// splat_disks(), are clumpy_cuda_splat_points / clumpy_cuda_splat_disks here; the process-global
// `advect_wrapx` (:20) is kept because advect_points sets it and other command TUs may read it.
#include "clumpy_command.hh"
#include "gpu.hh"
#include "npy.hh"

using std::string;
using std::vector;

bool advect_wrapx = false;

namespace {

struct SplatPoints : ClumpyCommand {
    bool exec(vector<string> args) override;
    string description() const override { return "consume a list of 2-tuples and create an image"; }
    string usage() const override { return "<input_pts> <dims> <kernel_type> <kernel_size> <alpha> <output_img>"; }
    string example() const override { return "bridson.npy 500x250 u8disk 5 1.0 splat.npy"; }
};

ClumpyCommand::Register registrar("splat_points", [] { return new SplatPoints(); });

bool SplatPoints::exec(vector<string> args) {
    if (args.size() != 6) {
        printf("The command takes 6 arguments.\n");
        return false;
    }
    const string dims = args[1], kernel_typestring = args[2];
    const int kernel_size = atoi(args[3].c_str());
    const float alpha = atof(args[4].c_str());
    const uint32_t width = atoi(dims.c_str());
    const uint32_t height = atoi(dims.substr(dims.find('x') + 1).c_str());
    if (0 == (kernel_size % 2)) {
        printf("Kernel size must be an odd integer.\n");
        return false;
    }
    const bool disk = kernel_typestring == "u8disk";
    if (!disk && kernel_typestring != "fp32sphere") {
        printf("Kernel type must be u8disk/fp32sphere.\n");
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
    const uint32_t npts = arr.shape[0];
    vector<uint8_t> dstimg((size_t)width * height);
    printf("Drawing %u points.\n", npts);
    const bool fast_path = alpha == 1 && kernel_size == 1; // splat_points.cc:93-94
    if (!fast_path && !disk) {
        printf("Only disks are implemented right now.\n");
        return false;
    }
    const vector<int> devices = clumpy_devices();
    if (!devices.empty()) { // several GPUs of this box, driven from this process
        if (!Gpu::report(clumpy_cuda_splat_multi(devices.data(), (int)devices.size(), arr.data<float>(), npts, width, height,
                                                 dstimg.data(), alpha, kernel_size, advect_wrapx ? 1 : 0)))
            return false;
        return npy::save_u8(args[5], dstimg.data(), {height, width});
    }
    Gpu gpu;
    if (!gpu) return false;
    int rc;
    if (fast_path) {
        rc = clumpy_cuda_splat_points(gpu.ctx, arr.data<float>(), npts, width, height, dstimg.data());
    } else {
        rc = clumpy_cuda_splat_disks(gpu.ctx, arr.data<float>(), npts, width, height, dstimg.data(), alpha,
                                     kernel_size, advect_wrapx ? 1 : 0);
    }
    if (!Gpu::report(rc)) return false;
    return npy::save_u8(args[5], dstimg.data(), {height, width});
}

} // namespace
