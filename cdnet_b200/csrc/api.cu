// api.cu -- library-level entry points: version, device check, launch counter and the optional
// per-launch CUDA-event profiler that bench.py uses for the live roofline numbers.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace cdnet {
unsigned long long g_launches = 0;
int g_prof_on = 0;
bool pdl_enabled() {
    static int on = -1;
    if (on < 0) on = getenv("CDNET_NO_PDL") ? 0 : 1;
    return on == 1;
}

struct ProfRec {
    const char* name;
    cudaEvent_t e0, e1;
};
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

void prof_begin(const char* name, cudaStream_t st) {
    ProfRec r{name, get_event(), get_event()};
    cudaEventRecord(r.e0, st);
    g_recs.push_back(r);
}

void prof_end(cudaStream_t st) {
    if (!g_recs.empty()) cudaEventRecord(g_recs.back().e1, st);
}
}  // namespace cdnet

extern "C" const char* cdnet_version(void) { return "cdnet_b200 0.1 (sm_100a)"; }

extern "C" unsigned long long cdnet_launch_count(void) { return cdnet::g_launches; }

extern "C" int cdnet_device_ok(int device) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
    return prop.major == 10 ? 1 : 0;
}

extern "C" void cdnet_profile_enable(int on) { cdnet::g_prof_on = on ? 1 : 0; }

// Synchronises the device, aggregates the recorded launches by kernel name and writes lines
// "name\tlaunches\ttotal_ms\n" into buf (truncated to cap-1 bytes).  Clears the records.
// Returns the number of distinct kernels.
extern "C" int cdnet_profile_report(char* buf, size_t cap) {
    using namespace cdnet;
    cudaDeviceSynchronize();
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto& a = agg[r.name];
            a.first += 1;
            a.second += ms;
        }
        g_pool.push_back(r.e0);
        g_pool.push_back(r.e1);
    }
    g_recs.clear();
    size_t off = 0;
    if (buf && cap) buf[0] = 0;
    for (auto& kv : agg) {
        char line[512];
        int n = snprintf(line, sizeof line, "%s\t%d\t%.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (buf && off + n + 1 < cap) {
            memcpy(buf + off, line, n + 1);
            off += n;
        }
    }
    return (int)agg.size();
}
