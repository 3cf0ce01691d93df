// Optional per-stage CUDA-event timing and launch counting (bench.py's roofline leg).
// Disabled by default: when off, a stage scope costs one predictable branch.
#include <cstdlib>
#include <vector>
#include "api_internal.cuh"

namespace adgs {

static bool g_profile_on = false;
static unsigned long long g_launches = 0;
struct StageSample {
    int stage;
    cudaEvent_t start, stop;
};
static std::vector<StageSample> g_samples;
static std::vector<cudaEvent_t> g_pool;

static const char* kStageNames[kNumStages] = {"per_gaussian_forward", "depth_sort", "offset_scan", "emit_instances",
                                              "tile_sort", "tile_ranges", "blend_forward", "blend_backward",
                                              "per_gaussian_backward", "rotation_backward", "fills"};

void count_launch(int n)
{
    g_launches += (unsigned long long)n;
}

int tune_variant(const char* env_name, int dflt)
{
    const char* v = getenv(env_name);
    return (v && *v) ? atoi(v) : dflt;
}

static cudaEvent_t get_event()
{
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

StageScope::StageScope(int stage, cudaStream_t stream) : stage_(stage), stream_(stream), active_(g_profile_on)
{
    if (!active_) return;
    StageSample s;
    s.stage = stage;
    s.start = get_event();
    s.stop = get_event();
    cudaEventRecord(s.start, stream);
    g_samples.push_back(s);
    index_ = (int)g_samples.size() - 1;
}

StageScope::~StageScope()
{
    if (!active_) return;
    cudaEventRecord(g_samples[index_].stop, stream_);
}

}  // namespace adgs

using namespace adgs;

extern "C" {

unsigned long long adgs_launch_count(void)
{
    return g_launches;
}

int adgs_profile_begin(void)
{
    for (auto& s : g_samples) {
        g_pool.push_back(s.start);
        g_pool.push_back(s.stop);
    }
    g_samples.clear();
    g_profile_on = true;
    return ADGS_OK;
}

int adgs_profile_num_stages(void)
{
    return kNumStages;
}

const char* adgs_profile_stage_name(int stage)
{
    return (stage >= 0 && stage < kNumStages) ? kStageNames[stage] : "";
}

int adgs_profile_end(float* ms_per_stage, int32_t* scopes_per_stage)
{
    g_profile_on = false;
    if (!ms_per_stage || !scopes_per_stage) return ADGS_ERR_ARG;
    for (int i = 0; i < kNumStages; ++i) {
        ms_per_stage[i] = 0.f;
        scopes_per_stage[i] = 0;
    }
    for (auto& s : g_samples) {
        cudaError_t e = cudaEventSynchronize(s.stop);
        if (e != cudaSuccess) return record_cuda_error(e, "profile_end");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, s.start, s.stop);
        ms_per_stage[s.stage] += ms;
        scopes_per_stage[s.stage] += 1;
        g_pool.push_back(s.start);
        g_pool.push_back(s.stop);
    }
    g_samples.clear();
    return ADGS_OK;
}

}  // extern "C"
