// Library-level entry points: version, error text, launch counter and the optional per-stage
// CUDA-event timer that bench.py uses to time the dominant kernel inside a real step.
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <vector>

#include "nfe_common.cuh"

namespace nfe {

static thread_local char g_error[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// ---- stage timer: pairs of events recorded on the launching stream around a stage --------------
struct StageRecord { int stage; cudaEvent_t start, stop; };
static std::mutex g_timer_mutex;
static std::atomic<int> g_timer_on{0};
static std::vector<StageRecord> g_records;
static std::vector<cudaEvent_t> g_free_events;

static cudaEvent_t take_event()
{
    if (!g_free_events.empty()) { cudaEvent_t e = g_free_events.back(); g_free_events.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

int stage_begin(int stage, cudaStream_t stream)
{
    if (!g_timer_on.load(std::memory_order_relaxed)) return -1;
    std::lock_guard<std::mutex> lock(g_timer_mutex);
    StageRecord r{stage, take_event(), take_event()};
    cudaEventRecord(r.start, stream);
    g_records.push_back(r);
    return (int)g_records.size() - 1;
}

void stage_end(int token, cudaStream_t stream)
{
    if (token < 0) return;
    std::lock_guard<std::mutex> lock(g_timer_mutex);
    if (token < (int)g_records.size()) cudaEventRecord(g_records[token].stop, stream);
}

}  // namespace nfe

NFE_EXPORT int nfe_version(void) { return 1; }
NFE_EXPORT const char* nfe_last_error(void) { return nfe::g_error; }
NFE_EXPORT uint64_t nfe_launch_count(void) { return nfe::g_launches.load(std::memory_order_relaxed); }

NFE_EXPORT int nfe_timing_enable(int on)
{
    nfe::g_timer_on.store(on ? 1 : 0);
    return 0;
}

NFE_EXPORT int nfe_timing_read(double* ms_by_stage, int64_t* count_by_stage, int n_stages, int reset)
{
    using namespace nfe;
    NFE_REQUIRE(ms_by_stage && count_by_stage && n_stages > 0, "nfe_timing_read: bad arguments");
    std::lock_guard<std::mutex> lock(g_timer_mutex);
    for (int i = 0; i < n_stages; ++i) { ms_by_stage[i] = 0.0; count_by_stage[i] = 0; }
    for (const StageRecord& r : g_records) {
        if (cudaEventSynchronize(r.stop) != cudaSuccess) { set_error("nfe_timing_read: event synchronise failed"); return 2; }
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, r.start, r.stop) != cudaSuccess) { set_error("nfe_timing_read: elapsed time failed"); return 2; }
        if (r.stage >= 0 && r.stage < n_stages) { ms_by_stage[r.stage] += ms; count_by_stage[r.stage] += 1; }
    }
    if (reset) {
        for (const StageRecord& r : g_records) { g_free_events.push_back(r.start); g_free_events.push_back(r.stop); }
        g_records.clear();
    }
    return 0;
}
