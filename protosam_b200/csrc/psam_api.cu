// Library-level pieces of the C ABI (include/psam_b200.h): version, error text, launch count.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "psam_common.cuh"

namespace psam {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void max_carveout_once(const void* kernel, bool* done)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || done[dev]) return;
    done[dev] = true;
    static const bool off = getenv("PSAM_NO_CARVEOUT") != nullptr;      // experiment knob
    if (!off) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

static std::vector<void (*)(TraceRec*)>& trace_setters()
{
    static std::vector<void (*)(TraceRec*)> v;
    return v;
}

void trace_register(void (*setter)(TraceRec*)) { trace_setters().push_back(setter); }

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// ---- optional per-kernel timing: one CUDA event pair around every launch of this library ----
bool g_profiling = false;
struct ProfSpan {
    const char* what;
    cudaEvent_t e0, e1;
};
static std::mutex g_prof_mu;
static std::vector<ProfSpan> g_spans;
static thread_local cudaEvent_t t_pending = nullptr;

void prof_begin(cudaStream_t stream)
{
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, stream);
    t_pending = e;
}

void prof_end(const char* what, cudaStream_t stream)
{
    if (!t_pending) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e1) != cudaSuccess) return;
    cudaEventRecord(e1, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_spans.push_back({what, t_pending, e1});
    t_pending = nullptr;
}

}  // namespace psam

static_assert(sizeof(psam_prompt_rec) == 96, "psam_prompt_rec layout is part of the ABI");
static_assert(sizeof(psam_image_hdr) == 64, "psam_image_hdr layout is part of the ABI");

extern "C" int psam_abi_version(void) { return PSAM_ABI_VERSION; }
extern "C" const char* psam_last_error(void) { return psam::g_err; }
extern "C" uint64_t psam_launch_count(void) { return psam::g_launches.load(std::memory_order_relaxed); }

extern "C" int psam_trace_install(void* buffer, size_t bytes)
{
    using namespace psam;
    if (!kTraceCompiled) {
        set_error("psam_trace_install: the library was built without kernel tracing (make -C protosam_b200/csrc TRACE=1)");
        return PSAM_ERR_UNSUPPORTED;
    }
    TraceRec* buf = static_cast<TraceRec*>(buffer);
    if (buf) {
        if (bytes < 2 * sizeof(TraceRec)) { set_error("psam_trace_install: buffer too small"); return PSAM_ERR_ARG; }
        TraceRec hdr;
        memset(&hdr, 0, sizeof(hdr));
        hdr.t1 = bytes / sizeof(TraceRec);
        if (cudaMemcpy(buf, &hdr, sizeof(hdr), cudaMemcpyHostToDevice) != cudaSuccess) return PSAM_ERR_LAUNCH;
    }
    for (auto f : trace_setters()) f(buf);
    return cudaDeviceSynchronize() == cudaSuccess ? PSAM_OK : PSAM_ERR_LAUNCH;
}

extern "C" void psam_profile_enable(int on)
{
    psam::g_profiling = on != 0;
}

extern "C" int psam_profile_collect(char* names, size_t names_bytes, float* total_ms, int32_t* launches, int max_kernels)
{
    using namespace psam;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<std::string> order;
    std::map<std::string, std::pair<double, int>> agg;
    for (ProfSpan& s : g_spans) {
        float ms = 0.f;
        cudaEventSynchronize(s.e1);
        if (cudaEventElapsedTime(&ms, s.e0, s.e1) == cudaSuccess) {
            auto it = agg.find(s.what);
            if (it == agg.end()) { order.push_back(s.what); agg[s.what] = {ms, 1}; }
            else { it->second.first += ms; it->second.second += 1; }
        }
        cudaEventDestroy(s.e0);
        cudaEventDestroy(s.e1);
    }
    g_spans.clear();
    int n = 0;
    size_t off = 0;
    if (names && names_bytes) names[0] = 0;
    for (const std::string& k : order) {
        if (n >= max_kernels) break;
        if (names && off + k.size() + 2 < names_bytes) {
            memcpy(names + off, k.c_str(), k.size());
            off += k.size();
            names[off++] = '\n';
            names[off] = 0;
        }
        if (total_ms) total_ms[n] = (float)agg[k].first;
        if (launches) launches[n] = agg[k].second;
        ++n;
    }
    return n;
}
