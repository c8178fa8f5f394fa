// tcgen05 / TMA chunked WKV6 forward (placeholder until the kernel lands; SIMT path serves all calls).
#include "common.cuh"

namespace wkv6 {
bool tc_forward_supported(const Args &) { return false; }
int tc_forward(const Args &) {
    set_error("tensor-core forward not built");
    return WKV6_EUNSUPPORTED;
}
}  // namespace wkv6
