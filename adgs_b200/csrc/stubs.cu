// Temporary: entry points that are declared in include/adgs_b200.h and implemented later.
#include "api_internal.cuh"
extern "C" {
int adgs_trajectory_forward(const adgs_model*, const adgs_time_basis*, const adgs_deformed*, adgs_stream_t) { return ADGS_ERR_UNSUPPORTED; }
size_t adgs_render_saved_bytes(int32_t N) { return (size_t)N * 32 + 128; }
int adgs_render_forward(const adgs_camera*, const adgs_model*, const adgs_time_basis*, int32_t, const adgs_images*, const adgs_deformed*, char*, char*, int64_t, char*, char*, adgs_stream_t) { return ADGS_ERR_UNSUPPORTED; }
int adgs_render_backward(const adgs_camera*, const adgs_model*, const adgs_time_basis*, int32_t, const int32_t*, const char*, const char*, const char*, const char*, const float*, const adgs_image_grads*, const adgs_model*, float*, char*, adgs_stream_t) { return ADGS_ERR_UNSUPPORTED; }
size_t adgs_knn_workspace_bytes(int32_t) { return 0; }
int adgs_dist_cuda2(int32_t, const float*, float*, char*, adgs_stream_t) { return ADGS_ERR_UNSUPPORTED; }
}
