// Drop-in replacements for the reference's cuda/*_op.cpp (yynil/RWKV_LM_EXT): same module names, same pybind
// functions and argument lists, so the reference's unmodified
//     load(name="wkv6", sources=["cuda/wkv6_op.cpp", "cuda/wkv6_cuda.cu"], extra_cuda_cflags=[...])      src/model.py:188-189
// builds THESE files when its cuda/ directory is replaced by (or symlinked to) this one.  The kernels are not
// compiled here: libwkv6_b200.so (plain nvcc for sm_100a, include/wkv6_b200.h) is opened at run time, because
// load() passes no linker flags and torch's arch list cannot assemble tcgen05 anyway.  Library lookup order:
//   $WKV6_B200_LIB, <this directory>/libwkv6_b200.so, <this directory>/../../rwkv_lm_ext_b200/libwkv6_b200.so
// The matching *_cuda.cu / rwkv6.cu files next to this header are intentionally empty translation units.
#pragma once
#include <dlfcn.h>
#include <stdlib.h>

#include <string>

#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

namespace wkv6_b200_shim {

inline void *open_library() {
    std::string here(__FILE__);
    const size_t cut = here.find_last_of('/');
    here = cut == std::string::npos ? std::string(".") : here.substr(0, cut);
    std::string tried;
    const char *env = getenv("WKV6_B200_LIB");
    const std::string candidates[3] = {env ? env : "", here + "/libwkv6_b200.so", here + "/../../rwkv_lm_ext_b200/libwkv6_b200.so"};
    for (const std::string &c : candidates) {
        if (c.empty()) continue;
        if (void *h = dlopen(c.c_str(), RTLD_NOW | RTLD_LOCAL)) return h;
        tried += "\n  " + c + ": " + dlerror();
    }
    TORCH_CHECK(false, "libwkv6_b200.so not found (set WKV6_B200_LIB); tried:", tried);
    return nullptr;
}

template <typename Fn>
Fn *sym(const char *name) {
    static void *lib = open_library();
    void *p = dlsym(lib, name);
    TORCH_CHECK(p != nullptr, "libwkv6_b200.so has no symbol ", name);
    return reinterpret_cast<Fn *>(p);
}

inline void check(int rc, const char *what) {
    if (rc == 0) return;
    static auto last_error = sym<const char *()>("wkv6b200_last_error");
    TORCH_CHECK(false, what, " failed with code ", rc, ": ", last_error());
}

inline void *stream() { return at::cuda::getCurrentCUDAStream().stream(); }

// byte workspace owned by torch's caching allocator (stream-ordered like every other tensor of the call)
inline torch::Tensor workspace(const torch::Tensor &like, size_t bytes) {
    return torch::empty({(int64_t)(bytes ? bytes : 1)}, like.options().dtype(torch::kUInt8));
}

typedef size_t ws_fn(int, int, int, int);

}  // namespace wkv6_b200_shim
