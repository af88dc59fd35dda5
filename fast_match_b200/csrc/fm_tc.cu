// fm_tc.cu -- placeholder until the tcgen05 kernel lands (next commit).
#include "fm_common.cuh"
namespace fm {
bool tc_supported() { return false; }
size_t top2_tc_workspace_bytes(int64_t, int64_t) { return 0; }
int launch_top2_tc(const uint8_t *, int64_t, const uint8_t *, int64_t, int32_t, uint32_t *,
                   int32_t *, uint64_t *, void *, size_t, cudaStream_t) {
    set_error("tcgen05 kernel not built");
    return FM_EUNSUPPORTED;
}
}  // namespace fm
