// make_ot_fixture.cpp -- writes tests/golden/tiny_model.ot exactly the way tch's `VarStore::save` does for a
// path without the ".safetensors" extension: tch's C shim `at_save_multi` is
//     torch::serialize::OutputArchive archive;
//     for each variable: archive.write(name, tensor, /*buffer=*/true);
//     archive.save_to(filename);
// (the reference calls it through `Network::save`, takzero/src/network/mod.rs:16-18).  The names are tch
// VarStore names of the reference network, including the `__K` collision suffix of the second SmallBlock
// (takzero/src/network/residual.rs:52-54).  Built against the libtorch that ships inside the Python torch wheel:
//   T=$(python -c 'import torch,os;print(os.path.dirname(torch.__file__))')
//   g++ -std=c++17 -O1 make_ot_fixture.cpp -o /tmp/make_ot_fixture -I$T/include \
//       -I$T/include/torch/csrc/api/include -L$T/lib -ltorch -ltorch_cpu -lc10 -Wl,-rpath,$T/lib
//   /tmp/make_ot_fixture tests/golden/tiny_model.ot
// Test tooling only (tests/test_model_file.py reads the committed output); nothing in the library uses libtorch.
#include <torch/serialize/archive.h>
#include <torch/torch.h>

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    torch::manual_seed(7);
    torch::serialize::OutputArchive archive;
    archive.write("core.input_conv2d.weight", torch::randn({4, 3, 3, 3}), true);
    archive.write("core.batch_norm.running_mean", torch::zeros({4}), true);
    archive.write("core.res_block_0.conv2d.weight", torch::randn({4, 4, 3, 3}), true);
    archive.write("core.res_block_0.conv2d.weight__7", torch::randn({4, 4, 3, 3}), true);
    archive.write("value.linear.bias", torch::randn({1}), true);
    archive.save_to(argv[1]);
    return 0;
}
