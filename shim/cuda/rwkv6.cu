// Intentionally empty: the reference lists cuda/rwkv6.cu in its load() call (e.g. src/model.py:188), but the kernels
// live in libwkv6_b200.so, built by plain nvcc for sm_100a (rwkv_lm_ext_b200/build.py) and opened by the *_op.cpp next to this file.
