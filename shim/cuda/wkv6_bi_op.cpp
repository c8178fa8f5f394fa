// module "wkv6_bi": the fused bidirectional op with an int32 [B,T] mask and w = fp32 -exp(w_raw), as
// cuda/wkv6_bi_op.cpp:8-16 of the reference declares it.
#include "wkv6_b200_dl.h"
using namespace wkv6_b200_shim;

typedef int fwd_fn(int, int, int, int, const int *, const void *, const void *, const void *, const float *, const void *, void *, void *);
typedef int bwd_fn(int, int, int, int, const int *, const void *, const void *, const void *, const float *, const void *, const void *,
                   void *, void *, void *, void *, void *, void *, size_t, void *);

void forward(int64_t B, int64_t T, int64_t C, int64_t H, const torch::Tensor &mask, torch::Tensor &r, torch::Tensor &k,
             torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &y) {
    const at::cuda::OptionalCUDAGuard guard(device_of(r));
    static auto fn = sym<fwd_fn>("wkv6_bi_forward");
    check(fn(B, T, C, H, mask.data_ptr<int>(), r.data_ptr(), k.data_ptr(), v.data_ptr(), w.data_ptr<float>(), u.data_ptr(),
             y.data_ptr(), stream()),
          "wkv6_bi_forward");
}

void backward(int64_t B, int64_t T, int64_t C, int64_t H, const torch::Tensor &mask, torch::Tensor &r, torch::Tensor &k,
              torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &gy, torch::Tensor &gr, torch::Tensor &gk,
              torch::Tensor &gv, torch::Tensor &gw, torch::Tensor &gu) {
    const at::cuda::OptionalCUDAGuard guard(device_of(r));
    static auto fn = sym<bwd_fn>("wkv6_bi_backward");
    static auto ws_bytes = sym<ws_fn>("wkv6_backward_workspace_bytes");
    const size_t n = ws_bytes(B, T, C, H);
    torch::Tensor ws = workspace(r, n);
    check(fn(B, T, C, H, mask.data_ptr<int>(), r.data_ptr(), k.data_ptr(), v.data_ptr(), w.data_ptr<float>(), u.data_ptr(),
             gy.data_ptr(), gr.data_ptr(), gk.data_ptr(), gv.data_ptr(), gw.data_ptr(), gu.data_ptr(), ws.data_ptr(), n, stream()),
          "wkv6_bi_backward");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward", &forward, "wkv6 bi forward (libwkv6_b200)");
    m.def("backward", &backward, "wkv6 bi backward (libwkv6_b200)");
}
