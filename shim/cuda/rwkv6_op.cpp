// module "rwkv6": the inference op, fp32 state [H,64,64] ([B,H,64,64]) updated in place, w = fp32 exp(-exp(w_raw)),
// one entry per element type as cuda/rwkv6_op.cpp:12-34 of the reference declares them.
#include "wkv6_b200_dl.h"
using namespace wkv6_b200_shim;

typedef int fwd_fn(int, int, int, int, int, float *, const void *, const void *, const void *, const float *, const void *, void *, void *);

static void run(int dtype, int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &state, torch::Tensor &r, torch::Tensor &k,
                torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &y) {
    const at::cuda::OptionalCUDAGuard guard(device_of(state));
    static auto fn = sym<fwd_fn>("rwkv6_forward");
    check(fn(dtype, B, T, C, H, state.data_ptr<float>(), r.data_ptr(), k.data_ptr(), v.data_ptr(), w.data_ptr<float>(), u.data_ptr(),
             y.data_ptr(), stream()),
          "rwkv6_forward");
}
void forward_bf16(int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &state, torch::Tensor &r, torch::Tensor &k,
                  torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &y) { run(0, B, T, C, H, state, r, k, v, w, u, y); }
void forward_fp16(int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &state, torch::Tensor &r, torch::Tensor &k,
                  torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &y) { run(1, B, T, C, H, state, r, k, v, w, u, y); }
void forward_fp32(int64_t B, int64_t T, int64_t C, int64_t H, torch::Tensor &state, torch::Tensor &r, torch::Tensor &k,
                  torch::Tensor &v, torch::Tensor &w, torch::Tensor &u, torch::Tensor &y) { run(2, B, T, C, H, state, r, k, v, w, u, y); }

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("forward_bf16", &forward_bf16, "rwkv6 forward_bf16 (libwkv6_b200)");
    m.def("forward_fp16", &forward_fp16, "rwkv6 forward_fp16 (libwkv6_b200)");
    m.def("forward_fp32", &forward_fp32, "rwkv6 forward_fp32 (libwkv6_b200)");
}
