"""Builds the drop-in modules of shim/cuda/ exactly the way the reference does (torch.utils.cpp_extension.load with
the source list and nvcc flags of src/model.py:188-189, src/model_run.py:46-47, cuda/wkv6_bi.py), into shim/_build/<name>
(git-ignored; travels to the GPU box, where load() then finds the module up to date and only imports it).
usage: python shim/build_shims.py [name ...]        -- works without a GPU (TORCH_CUDA_ARCH_LIST=10.0)"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
MODULES = {      # name -> sources, as the reference's load() calls list them
    "wkv6": ["cuda/wkv6_op.cpp", "cuda/wkv6_cuda.cu"],
    "wkv6state": ["cuda/wkv6state_op.cpp", "cuda/wkv6state_cuda.cu"],
    "wkv6infctx": ["cuda/wkv6infctx_op.cpp", "cuda/wkv6infctx_cuda.cu"],
    "wkv6_bi": ["cuda/wkv6_bi_op.cpp", "cuda/wkv6_bi_cuda.cu"],
    "rwkv6": ["cuda/rwkv6_op.cpp", "cuda/rwkv6.cu"],
}
REF_CUDA_CFLAGS = ["-res-usage", "--use_fast_math", "-O3", "-Xptxas -O3", "--extra-device-vectorization", "-D_N_=64", "-D_T_=4096"]


def load_shim(name, verbose=False, rebuild=False):
    """The reference's own call, with a fixed build directory.  A module already built there (by `python
    shim/build_shims.py`, e.g. in the build container) is imported as it is unless `rebuild` is set."""
    import torch  # noqa: F401  (the extension links against libtorch)
    bd = os.path.join(HERE, "_build", name)
    so = os.path.join(bd, name + ".so")
    if os.path.exists(so) and not rebuild:
        import importlib.util
        spec = importlib.util.spec_from_file_location(name, so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils.cpp_extension import load
    os.makedirs(bd, exist_ok=True)
    return load(name=name, sources=[os.path.join(HERE, s) for s in MODULES[name]], verbose=verbose,
                extra_cuda_cflags=REF_CUDA_CFLAGS, build_directory=bd)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(MODULES)):
        m = load_shim(n, verbose=True, rebuild=True)
        print("built", n, "->", m.__file__)
