// Diagnostic host build (not a product path, not a test): the device functions compiled for the CPU with an event hook
// on FTL_COUNT, so that tools/scan_stats.py can histogram the exact green-zone scans per env-frame / per warp-step.
#include <vector>
#include <cstdint>
static std::vector<int64_t> g_events;   // (k << 32 | v)
void ftl_count_hook(int k, long long v) { g_events.push_back(((int64_t)k << 32) | (int64_t)(uint32_t)v); }
#include "../../tests/hostsim/hostsim.cpp"
extern "C" long long scan_events(int64_t* out, long long cap) {
    long long n = (long long)g_events.size();
    if (out) { for (long long i = 0; i < n && i < cap; i++) out[i] = g_events[i]; g_events.clear(); }
    return n;
}
