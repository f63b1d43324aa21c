// Per-launch CUDA-event profiler: bench.py runs one instrumented pass of the workload after the timed
// steps and reads back, per kernel class, launch count, summed device time and summed algorithmic work.
#include <cstdlib>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "../../include/cbx_b200.h"

namespace {
struct Rec { cudaEvent_t a, b; int cls; double work; };
std::mutex g_mu;
std::vector<Rec> g_recs;
bool g_on = false;
thread_local Rec t_cur;
thread_local int t_suspend = 0;
}  // namespace

bool prof_enabled() { return g_on && t_suspend == 0; }
bool prof_active() { return g_on; }
void prof_suspend(int d) { t_suspend += d; }
bool pdl_enabled() {
    static int v = -1;
    if (v < 0) { const char* d = getenv("CBX_DISABLE_PDL"); v = (d && d[0] == '1') ? 0 : 1; }
    return v == 1 && !prof_enabled();   // per-launch event timing needs plain stream order
}
void prof_begin_launch(int cls, double work, cudaStream_t st) {
    t_cur.cls = cls; t_cur.work = work;
    cudaEventCreate(&t_cur.a); cudaEventCreate(&t_cur.b);
    cudaEventRecord(t_cur.a, st);
}
void prof_end_launch(cudaStream_t st) {
    cudaEventRecord(t_cur.b, st);
    std::lock_guard<std::mutex> g(g_mu);
    g_recs.push_back(t_cur);
}

extern "C" {
int cbx_profile_begin(void) {
    std::lock_guard<std::mutex> g(g_mu);
    for (auto& r : g_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_recs.clear();
    g_on = true;
    return 0;
}
int cbx_profile_end(int64_t* counts, double* ms, double* work, int n_classes) {
    g_on = false;
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> g(g_mu);
    for (int i = 0; i < n_classes; i++) { counts[i] = 0; ms[i] = 0; work[i] = 0; }
    for (auto& r : g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && r.cls < n_classes) { counts[r.cls]++; ms[r.cls] += t; work[r.cls] += r.work; }
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
    }
    g_recs.clear();
    return 0;
}
}
