// extern "C" view of the drop-in classes (include/gpismap/GPisMap3.h, GPisMap.h) for callers that
// cannot link C++ directly: the Python tests / bench (ctypes) and the mex gateway pattern of the
// reference (one process-global map, string commands: mex/mexGPisMap3.cpp:28,43-169).
#include <cstring>
#include <vector>

#include "gpismap/GPisMap.h"
#include "gpismap/GPisMap3.h"
#include "map_core.hpp"

extern "C" {

// ------------------------------------------------------------------ GPisMap3
void* gm3_create(int device) { GPisMap3* m = new GPisMap3(); m->setDevice(device); return m; }
void* gm3_create_cam(int device, float fx, float fy, float cx, float cy, int w, int h) {
    GPisMap3Param p;
    camParam c(fx, fy, cx, cy, (float)w, (float)h);
    GPisMap3* m = new GPisMap3(p, c);
    m->setDevice(device);
    return m;
}
int gm3_set_tuning(void* m, float rtimes, float min_half, float max_half, float init_root_half, float cluster_half) {
    GPisMap3Tuning t;
    if (rtimes > 0) t.rtimes = rtimes;
    if (min_half > 0) t.tree_min_half = min_half;
    if (max_half > 0) t.tree_max_half = max_half;
    if (init_root_half > 0) t.tree_init_root_half = init_root_half;
    if (cluster_half > 0) t.tree_cluster_half = cluster_half;
    return ((GPisMap3*)m)->setTuning(t) ? 1 : 0;
}
void gm3_destroy(void* m) { delete (GPisMap3*)m; }
void gm3_reset(void* m) { ((GPisMap3*)m)->reset(); }
void gm3_set_cam(void* m, float fx, float fy, float cx, float cy, int w, int h) {
    ((GPisMap3*)m)->resetCam(camParam(fx, fy, cx, cy, (float)w, (float)h));
}
void gm3_update(void* m, float* depth, int N, const float* pose12) {
    std::vector<float> pose(pose12, pose12 + 12);
    ((GPisMap3*)m)->update(depth, N, pose);
}
int gm3_test(void* m, float* x, int n, float* res) { return ((GPisMap3*)m)->test(x, 3, n, res) ? 1 : 0; }
int gm3_get_all_points(void* m, float* out, int cap) {
    std::vector<float> pos;
    ((GPisMap3*)m)->getAllPoints(pos);
    const int n = (int)pos.size() / 3;
    if (out && cap >= n && n > 0) std::memcpy(out, pos.data(), sizeof(float) * pos.size());
    return n;
}
int gm3_all_samples(void* m, float* out, int cap) {
    std::vector<float> s;
    ((GPisMap3*)m)->getAllSamples(s);
    const int n = (int)s.size() / 9;
    if (out && cap >= n && n > 0) std::memcpy(out, s.data(), sizeof(float) * s.size());
    return n;
}
int gm3_leaves(void* m, float* centres, int* counts, int cap) {
    std::vector<float> c; std::vector<int> k;
    ((GPisMap3*)m)->getLeaves(c, k);
    const int n = (int)k.size();
    if (cap >= n && n > 0) {
        if (centres) std::memcpy(centres, c.data(), sizeof(float) * c.size());
        if (counts) std::memcpy(counts, k.data(), sizeof(int) * k.size());
    }
    return n;
}
int gm3_insert_samples(void* m, const float* s, int n) { return ((GPisMap3*)m)->insertSamples(s, n); }
int gm3_train_active(void* m) { return ((GPisMap3*)m)->trainActive(); }
int gm3_activate_all(void* m) { return ((GPisMap3*)m)->activateAll(); }
// development aid: coarse host profile (map_core.hpp), read and reset
void gm_profile(double* secs16, long long* calls16) {
    for (int i = 0; i < 16; ++i) { secs16[i] = gpismap_host::g_prof_s[i]; calls16[i] = gpismap_host::g_prof_n[i]; gpismap_host::g_prof_s[i] = 0; gpismap_host::g_prof_n[i] = 0; }
}
void gm3_timing(void* m, double* phases5, int* counts3, float* train_ms) {
    const GPisMap3Timing& t = ((GPisMap3*)m)->lastTiming();
    for (int i = 0; i < 5; ++i) phases5[i] = t.phase[i];
    counts3[0] = t.valid_pixels; counts3[1] = t.active_leaves; counts3[2] = t.trained_leaves;
    *train_ms = t.train_kernel_ms;
}
void* gm3_ctx(void* m) { return ((GPisMap3*)m)->cabiContext(); }

// ------------------------------------------------------------------ GPisMap
void* gm2_create(int device) { GPisMap* m = new GPisMap(); m->setDevice(device); return m; }
void gm2_destroy(void* m) { delete (GPisMap*)m; }
void gm2_reset(void* m) { ((GPisMap*)m)->reset(); }
void gm2_update(void* m, float* theta, float* range, int N, const float* pose6) {
    std::vector<float> pose(pose6, pose6 + 6);
    ((GPisMap*)m)->update(theta, range, N, pose);
}
int gm2_test(void* m, float* x, int n, float* res) { return ((GPisMap*)m)->test(x, 2, n, res) ? 1 : 0; }
int gm2_get_all_points(void* m, float* out, int cap) {
    std::vector<float> pos;
    ((GPisMap*)m)->getAllPoints(pos);
    const int n = (int)pos.size() / 2;
    if (out && cap >= n && n > 0) std::memcpy(out, pos.data(), sizeof(float) * pos.size());
    return n;
}
int gm2_all_samples(void* m, float* out, int cap) {
    std::vector<float> s;
    ((GPisMap*)m)->getAllSamples(s);
    const int n = (int)s.size() / 7;
    if (out && cap >= n && n > 0) std::memcpy(out, s.data(), sizeof(float) * s.size());
    return n;
}
int gm2_leaves(void* m, float* centres, int* counts, int cap) {
    std::vector<float> c; std::vector<int> k;
    ((GPisMap*)m)->getLeaves(c, k);
    const int n = (int)k.size();
    if (cap >= n && n > 0) {
        if (centres) std::memcpy(centres, c.data(), sizeof(float) * c.size());
        if (counts) std::memcpy(counts, k.data(), sizeof(int) * k.size());
    }
    return n;
}
int gm2_insert_samples(void* m, const float* s, int n) { return ((GPisMap*)m)->insertSamples(s, n); }
int gm2_train_active(void* m) { return ((GPisMap*)m)->trainActive(); }
int gm2_activate_all(void* m) { return ((GPisMap*)m)->activateAll(); }
void gm2_timing(void* m, double* phases5, int* counts3, float* train_ms) {
    const GPisMapTiming& t = ((GPisMap*)m)->lastTiming();
    for (int i = 0; i < 5; ++i) phases5[i] = t.phase[i];
    counts3[0] = t.valid_beams; counts3[1] = t.active_leaves; counts3[2] = t.trained_leaves;
    *train_ms = t.train_kernel_ms;
}
void* gm2_ctx(void* m) { return ((GPisMap*)m)->cabiContext(); }

}  // extern "C"
