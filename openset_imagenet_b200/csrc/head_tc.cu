// placeholder until the tcgen05 head lands (replaced in the next commit)
#include "osb_common.cuh"
extern "C" size_t osb_head_workspace_bytes(int64_t, int, int) { return 0; }
extern "C" int osb_head_forward(const void*, int64_t, const void*, const float*, const int64_t*, const float*, int64_t, int, int, int, double,
                                int64_t, const double*, int, double*, float*, float*, float*, float*, int32_t*, void*, int64_t, void*, size_t, void*) {
    osb::set_error("osb_head_forward: not built yet");
    return OSB_ERR_SHAPE;
}
extern "C" int osb_head_backward(const void*, int64_t, const void*, int64_t, const void*, int64_t, int, int, const float*, void*, int64_t,
                                 float*, float*, void*, size_t, void*) {
    osb::set_error("osb_head_backward: not built yet");
    return OSB_ERR_SHAPE;
}
