// Driver for tests/test_gpu_cpp_twin.py: runs an SVDQuant MLP (fc1, fc2) through the C++ twin
// nunchaku_b200::GEMM_W4A4 (include/nunchaku_b200_linear.hpp) on tensors read from a blob file and
// writes the outputs to another blob.  The Python test feeds the SAME bytes to the Python mirror
// (SVDQW4A4Linear) and demands bit-identical results.
//
// blob := u32 count, then per tensor: u32 name_len, name, u32 ndim, i64 dims[ndim], u32 elem_size, bytes
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>

#include "nunchaku_b200_linear.hpp"

using nunchaku_b200::DeviceTensor;
using nunchaku_b200::GEMM_W4A4;

struct HostTensor {
    std::vector<int64_t> shape;
    uint32_t elem = 0;
    std::vector<char> data;
};

static std::map<std::string, HostTensor> read_blob(const char *path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error(std::string("cannot open ") + path);
    std::map<std::string, HostTensor> m;
    uint32_t count = 0;
    f.read(reinterpret_cast<char *>(&count), 4);
    for (uint32_t i = 0; i < count; i++) {
        uint32_t nl = 0, nd = 0;
        f.read(reinterpret_cast<char *>(&nl), 4);
        std::string name(nl, '\0');
        f.read(name.data(), nl);
        f.read(reinterpret_cast<char *>(&nd), 4);
        HostTensor t;
        t.shape.resize(nd);
        f.read(reinterpret_cast<char *>(t.shape.data()), 8 * nd);
        f.read(reinterpret_cast<char *>(&t.elem), 4);
        size_t bytes = t.elem;
        for (int64_t d : t.shape) bytes *= static_cast<size_t>(d);
        t.data.resize(bytes);
        f.read(t.data.data(), static_cast<std::streamsize>(bytes));
        m[name] = std::move(t);
    }
    if (!f) throw std::runtime_error("truncated blob");
    return m;
}

static void write_tensor(std::ofstream &f, const std::string &name, const std::vector<int64_t> &shape, uint32_t elem, const void *dev) {
    size_t bytes = elem;
    for (int64_t d : shape) bytes *= static_cast<size_t>(d);
    std::vector<char> h(bytes);
    nunchaku_b200::cuda_check(cudaMemcpy(h.data(), dev, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H");
    const uint32_t nl = static_cast<uint32_t>(name.size()), nd = static_cast<uint32_t>(shape.size());
    f.write(reinterpret_cast<const char *>(&nl), 4);
    f.write(name.data(), nl);
    f.write(reinterpret_cast<const char *>(&nd), 4);
    f.write(reinterpret_cast<const char *>(shape.data()), 8 * nd);
    f.write(reinterpret_cast<const char *>(&elem), 4);
    f.write(h.data(), static_cast<std::streamsize>(bytes));
}

static void load_layer(GEMM_W4A4 &g, const std::map<std::string, HostTensor> &blob, const std::string &prefix) {
    for (const char *key : {"qweight", "wscales", "bias", "lora_down", "lora_up", "smooth", "wcscales", "wtscale"}) {
        auto it = blob.find(prefix + key);
        if (it == blob.end()) continue;
        g.load_param(key, it->second.data.data(), it->second.shape, it->second.elem);
    }
}

int main(int argc, char **argv) {
    if (argc != 3) {
        std::fprintf(stderr, "usage: %s <in.blob> <out.blob>\n", argv[0]);
        return 2;
    }
    try {
        auto blob = read_blob(argv[1]);
        const int64_t *meta = reinterpret_cast<const int64_t *>(blob.at("meta").data.data());  // M, D, H, fp4, dtype
        const int M = static_cast<int>(meta[0]), D = static_cast<int>(meta[1]), H = static_cast<int>(meta[2]);
        const bool fp4 = meta[3] != 0;
        const nb200_dtype dt = static_cast<nb200_dtype>(meta[4]);
        GEMM_W4A4 fc1(D, H, true, fp4, dt), fc2(H, D, true, fp4, dt);
        load_layer(fc1, blob, "fc1.");
        load_layer(fc2, blob, "fc2.");
        const HostTensor &xh = blob.at("x");
        DeviceTensor x({M, D}, 2);
        nunchaku_b200::cuda_check(cudaMemcpy(x.data(), xh.data.data(), xh.data.size(), cudaMemcpyHostToDevice), "cudaMemcpy H2D");

        DeviceTensor y_plain({M, H}, 2), y_silu({M, H}, 2), y_mlp({M, D}, 2);
        fc1.forward(x.data(), M, y_plain.data());
        fc1.forward_silu(x.data(), M, y_silu.data());
        GEMM_W4A4::QuantizedActivation q = fc1.forward_gelu_quant(x.data(), M, &fc2);
        fc2.forward_quant(q, GEMM_W4A4::FuseOptions::EMPTY, nullptr, y_mlp.data(), nullptr);
        nunchaku_b200::cuda_check(cudaDeviceSynchronize(), "cudaDeviceSynchronize");

        std::ofstream f(argv[2], std::ios::binary);
        const uint32_t count = 3;
        f.write(reinterpret_cast<const char *>(&count), 4);
        write_tensor(f, "y_plain", {M, H}, 2, y_plain.data());
        write_tensor(f, "y_silu", {M, H}, 2, y_silu.data());
        write_tensor(f, "y_mlp", {M, D}, 2, y_mlp.data());
        std::printf("ok rank=%d unsigned_next=%d\n", fc1.lora_rank, int(q.is_unsigned));
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "linear_twin_main: %s\n", e.what());
        return 1;
    }
}
