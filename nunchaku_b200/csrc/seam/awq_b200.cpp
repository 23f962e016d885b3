// Drop-in definition of the reference's AWQ W4A16 GEMV entry point (src/kernels/awq/gemv_awq.h:6-7) on libnunchaku_b200.so:
// a maintainer replaces src/kernels/awq/gemv_awq.cu (setup.py source list) by this file.  GEMV_AWQ::forward (src/Linear.cpp:56-86)
// and nunchaku::ops::gemv_awq (nunchaku/csrc/ops.h:130-152) compile and link unchanged.  The 4-bit weights, scales and zeros are
// read in the checkpoint layout, so there is nothing to convert and nothing to cache.
#include "Tensor.h"
#include "common.h"
#include "kernels/awq/gemv_awq.h"
#include "nunchaku_b200.h"

Tensor gemv_awq(Tensor _in_feats, Tensor _kernel, Tensor _scaling_factors, Tensor _zeros, int m, int n, int k, int group_size) {
    int dtype;
    if (_in_feats.scalar_type() == Tensor::BF16)
        dtype = NB200_BF16;
    else if (_in_feats.scalar_type() == Tensor::FP16)
        dtype = NB200_FP16;
    else
        throw std::invalid_argument("gemv_awq: fp16 / bf16 input expected");
    if (_scaling_factors.scalar_type() != _in_feats.scalar_type() || _zeros.scalar_type() != _in_feats.scalar_type())
        throw std::invalid_argument("gemv_awq: scales / zeros must have the input's dtype");
    // output shape as the reference builds it (gemv_awq.cu:253-259): the input's shape with the last dimension replaced by n
    std::vector<int> shape = _in_feats.shape.dataExtent;
    shape.back() = n;
    Tensor out = Tensor::empty(shape, _in_feats.scalar_type(), _in_feats.device());
    const int st = nb200_gemv_awq(dtype, _in_feats.data_ptr(), _kernel.data_ptr(), _scaling_factors.data_ptr(), _zeros.data_ptr(), out.data_ptr(), m, n, k,
                                  group_size, getCurrentCUDAStream());
    if (st != NB200_OK) throw std::runtime_error(std::string("nb200_gemv_awq: ") + nb200_last_error());
    return out;
}
