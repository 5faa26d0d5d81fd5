// placeholder, replaced below
#include "glg_common.cuh"
extern "C" int glg_collision(const float*, const float*, uint8_t*, int32_t, int32_t, int32_t, glg_stream_t) { glg::set_error("not built yet"); return GLG_ERR_UNSUPPORTED; }
extern "C" int glg_smallest_distance(const float*, const float*, float*, int32_t, int32_t, int32_t, glg_stream_t) { glg::set_error("not built yet"); return GLG_ERR_UNSUPPORTED; }
extern "C" int glg_is_valid(const float*, uint8_t*, int32_t, int32_t, glg_stream_t) { glg::set_error("not built yet"); return GLG_ERR_UNSUPPORTED; }
extern "C" int64_t glg_game_workspace_bytes(int32_t, int32_t, int32_t) { return 0; }
extern "C" int glg_game_create(glg_game**, void*, int64_t, const float*, const float*, int32_t, int32_t, int32_t, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
extern "C" void glg_game_destroy(glg_game*) {}
extern "C" int glg_game_validate_tracks(glg_game*, uint8_t*, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
extern "C" int glg_game_update_players(glg_game*, const int64_t*, const float*, int32_t, int32_t, uint8_t*, uint8_t*, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
extern "C" int glg_game_smallest_distance(glg_game*, const int64_t*, const float*, int32_t, int32_t, float*, glg_stream_t) { return GLG_ERR_UNSUPPORTED; }
