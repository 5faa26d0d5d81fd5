// placeholder, replaced below
#include "glg_common.cuh"
extern "C" int glg_pacman_step(int32_t*, int32_t*, const int32_t*, double*, int32_t, int32_t, int32_t, int32_t, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
extern "C" int glg_pacman_observe(const int32_t*, float*, int32_t, int32_t, int32_t, int32_t, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
