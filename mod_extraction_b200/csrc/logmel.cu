// placeholder until the FFT kernel lands (next commit)
#include "common.cuh"
using namespace modfx;
extern "C" int modfx_logmel_f32(const float*, float*, int64_t, int64_t, int32_t, int32_t, int32_t, const float*,
                                const int32_t*, const int32_t*, const float*, int32_t, float, void*) {
    return fail(MODFX_ERR_UNSUPPORTED, "log-mel kernel not built yet");
}
