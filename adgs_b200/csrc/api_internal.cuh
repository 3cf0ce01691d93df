// Internal glue shared by the C-ABI translation units.
#pragma once
#include "common.cuh"
#include "gaussian_math.cuh"
#include "sort.cuh"
#include "raster.cuh"
#include "blend.cuh"

namespace adgs {

int record_cuda_error(cudaError_t e, const char* where);
int check_stage(const char* where, bool debug, cudaStream_t stream);
RasterParams make_raster_params(const adgs_camera* cam);

int bin_and_blend(const adgs_camera* cam, int P, int D_S, bool has_flow, const float* semantic,
                  const adgs_images* out, const int32_t* radii, GeometryState& gs, char* binning,
                  adgs_alloc_fn binning_alloc, void* alloc_user, int64_t capacity, ImageState& is,
                  bool sync_for_count, int* num_rendered, cudaStream_t stream, int stages = 3,
                  const float* mean_x = nullptr, const float* mean_y = nullptr);  // stages: 1 = bin, 2 = blend

// ---- per-stage profiling / launch counting (profile.cu) ------------------------------------------
enum Stage {
    kStagePerGaussianFwd = 0,
    kStageDepthSort,
    kStageScan,
    kStageEmit,
    kStageTileSort,
    kStageTileRanges,
    kStageBlendFwd,
    kStageBlendBwd,
    kStagePerGaussianBwd,
    kStageRotationBwd,
    kStageFills,
    kNumStages
};

void count_launch(int n);
int tune_variant(const char* env_name, int dflt);  // integer tuning knob from the environment (profile.cu)

struct StageScope {
    StageScope(int stage, cudaStream_t stream);
    ~StageScope();
    int stage_;
    cudaStream_t stream_;
    bool active_;
    int index_ = -1;
};

const uint32_t* sorted_point_list(const BinningState& bs, int num_tiles);
const uint32_t* sorted_tile_ids(const BinningState& bs, int num_tiles);

}  // namespace adgs
