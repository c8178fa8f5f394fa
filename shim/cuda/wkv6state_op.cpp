// module "wkv6state": forward / backward with raw bf16 logits w and a state s, as cuda/wkv6state_op.cpp:8-16 of the reference
// declares them (s = time_state [H,64,64], shared by the batch).
#include "wkv6_b200_dl.h"
using namespace wkv6_b200_shim;

typedef int fwd_fn(int, int, int, int, const void *, const void *, const void *, const void *, const void *, const void *, void *, void *);
typedef int bwd_fn(int, int, int, int, const void *, const void *, const void *, const void *, const void *, const void *, const void *,
                   void *, void *, void *, void *, void *, void *, void *, size_t, void *);

void forward(int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &r, torch::Tensor &k, torch::Tensor &v, torch::Tensor &w,
             torch::Tensor &u, torch::Tensor &s, torch::Tensor &y) {
    const at::cuda::OptionalCUDAGuard guard(device_of(r));
    static auto fn = sym<fwd_fn>("wkv6state_forward");
    check(fn(B, T, C, H, r.data_ptr(), k.data_ptr(), v.data_ptr(), w.data_ptr(), u.data_ptr(), s.data_ptr(), y.data_ptr(), stream()),
          "wkv6state_forward");
}

void backward(int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &r, torch::Tensor &k, torch::Tensor &v, torch::Tensor &w,
              torch::Tensor &u, torch::Tensor &s, torch::Tensor &gy, torch::Tensor &gr, torch::Tensor &gk, torch::Tensor &gv,
              torch::Tensor &gw, torch::Tensor &gu, torch::Tensor &gs) {
    const at::cuda::OptionalCUDAGuard guard(device_of(r));
    static auto fn = sym<bwd_fn>("wkv6state_backward");
    static auto ws_bytes = sym<ws_fn>("wkv6_backward_workspace_bytes");
    const size_t n = ws_bytes(B, T, C, H);
    torch::Tensor ws = workspace(r, n);
    check(fn(B, T, C, H, r.data_ptr(), k.data_ptr(), v.data_ptr(), w.data_ptr(), u.data_ptr(), s.data_ptr(), gy.data_ptr(),
             gr.data_ptr(), gk.data_ptr(), gv.data_ptr(), gw.data_ptr(), gu.data_ptr(), gs.data_ptr(), ws.data_ptr(), n, stream()),
          "wkv6state_backward");
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward", &forward, "wkv6state forward (libwkv6_b200)");
    m.def("backward", &backward, "wkv6state backward (libwkv6_b200)");
}
