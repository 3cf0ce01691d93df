// Temporary: entry points that are declared in include/adgs_b200.h and implemented later.
#include "api_internal.cuh"
extern "C" {
size_t adgs_knn_workspace_bytes(int32_t) { return 0; }
int adgs_dist_cuda2(int32_t, const float*, float*, char*, adgs_stream_t) { return ADGS_ERR_UNSUPPORTED; }
}
