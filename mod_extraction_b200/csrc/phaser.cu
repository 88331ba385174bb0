// placeholder until the phaser kernel lands (next commit)
#include "common.cuh"
using namespace modfx;
extern "C" int64_t modfx_phaser_workspace_bytes(int32_t, int64_t) { return 0; }
extern "C" int modfx_phaser_f32(const float*, float*, int32_t, int64_t, float, const float*, const float*, const float*,
                                const float*, const float*, int32_t, const int32_t*, int32_t, void*, void*) {
    return fail(MODFX_ERR_UNSUPPORTED, "phaser kernel not built yet");
}
