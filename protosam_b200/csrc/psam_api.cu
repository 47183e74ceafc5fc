// Library-level pieces of the C ABI (include/psam_b200.h): version, error text, launch count.
#include <atomic>
#include <cstdarg>
#include <cstdio>

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

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace psam

static_assert(sizeof(psam_prompt_rec) == 96, "psam_prompt_rec layout is part of the ABI");
static_assert(sizeof(psam_image_hdr) == 64, "psam_image_hdr layout is part of the ABI");

extern "C" int psam_abi_version(void) { return PSAM_ABI_VERSION; }
extern "C" const char* psam_last_error(void) { return psam::g_err; }
extern "C" uint64_t psam_launch_count(void) { return psam::g_launches.load(std::memory_order_relaxed); }
