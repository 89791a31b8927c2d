// Drop-in GPisMap3: same public class API, parameter structs and defaults as the reference
// (cpp/include/GPisMap3.h:29-140), with the data-parallel hot path running on a B200 through the
// C ABI in include/gpis_b200.h. Host side (sensor pre-processing, octree bookkeeping) is new code
// with the reference's semantics; see DESIGN.md.
#ifndef GPISMAP_B200_GPISMAP3_H
#define GPISMAP_B200_GPISMAP3_H

#include <vector>

#include "params.h"

typedef struct camParam_ {   // cpp/include/GPisMap3.h:29-46
    float fx, fy, cx, cy;
    int width, height;
    camParam_() : fx(568.0f), fy(568.0f), cx(310.f), cy(224.f), width(640), height(480) {}
    camParam_(float fx_, float fy_, float cx_, float cy_, float w_, float h_)
        : fx(fx_), fy(fy_), cx(cx_), cy(cy_), width((int)w_), height((int)h_) {}
} camParam;

typedef struct GPisMap3Param_ {   // cpp/include/GPisMap3.h:48-81
    float delx;                 // numerical step (surface normal sampling)
    float fbias;                // constant map bias (mean of the GP)
    float obs_var_thre;         // ObsGP variance above which a prediction is not trusted
    int obs_skip;               // use every skip-th pixel
    float min_position_noise;
    float min_grad_noise;
    float map_scale_param;
    float map_noise_param;
    GPisMap3Param_()
        : delx((float)gpismap_defaults::kDelx3), fbias((float)gpismap_defaults::kFbias3),
          obs_var_thre((float)gpismap_defaults::kObsVarThre3), obs_skip(gpismap_defaults::kObsSkip3),
          min_position_noise((float)gpismap_defaults::kMinPosNoise3),
          min_grad_noise((float)gpismap_defaults::kMinGradNoise3),
          map_scale_param((float)gpismap_defaults::kMapScale3), map_noise_param((float)gpismap_defaults::kMapNoise3) {}
} GPisMap3Param;

// Tree / training-ball constants that the reference fixes at compile time (cpp/include/params.h:40-44,
// cpp/src/GPisMap3.cpp:26-31); here they are run-time values so that BASELINE configs[3]/[4] (larger leaves, larger
// maps) need no rebuild. Defaults reproduce the reference. Not part of the reference API: see setTuning().
struct GPisMap3Tuning {
    float rtimes;                 // training-ball radius = rtimes * cluster half  (GPISMAP3_RTIMES)
    float tree_min_half;          // GPISMAP3_TREE_MIN_HALF_LENGTH
    float tree_max_half;          // GPISMAP3_TREE_MAX_HALF_LENGTH
    float tree_init_root_half;    // GPISMAP3_TREE_INIT_ROOT_HALF_LENGTH
    float tree_cluster_half;      // GPISMAP3_TREE_CLUSTER_HALF_LENGTH (the query's search box is 3x this, GPisMap3.cpp:811)
    GPisMap3Tuning()
        : rtimes((float)gpismap_defaults::kRtimes3), tree_min_half((float)gpismap_defaults::kTree3MinHalf),
          tree_max_half((float)gpismap_defaults::kTree3MaxHalf), tree_init_root_half((float)gpismap_defaults::kTree3InitRootHalf),
          tree_cluster_half((float)gpismap_defaults::kTree3ClusterHalf) {}
};

// Per-phase wall-clock of the last update() call, seconds (preproc, regressObs, updateMapPoints,
// addNewMeas, updateGPs) — the phases of GPisMap3::update (cpp/src/GPisMap3.cpp:218-237).
struct GPisMap3Timing {
    double phase[5];
    int valid_pixels, active_leaves, trained_leaves;
    float train_kernel_ms;
};

class GPisMap3 {
public:
    GPisMap3();
    explicit GPisMap3(GPisMap3Param par);
    GPisMap3(GPisMap3Param par, camParam c);
    ~GPisMap3();
    GPisMap3(const GPisMap3&) = delete;
    GPisMap3& operator=(const GPisMap3&) = delete;

    void reset();
    void resetMap() { reset(); }   // north_star name for reset()
    void getAllPoints(std::vector<float>& pos);
    // dataz: depth in metres, column-major (dataz[col*height + row]); pose = [t(3) | R col-major(9)]
    void update(float* dataz, int N, std::vector<float>& pose);
    // x: dim x leng interleaved; res: 8 x leng, read-modify-write [f, grad(3), var_f, var_grad(3)]
    bool test(float* x, int dim, int leng, float* res);
    void resetCam(camParam c);

    // ---- additions (not in the reference API)
    void setDevice(int cuda_device);            // before the first update(); default 0
    bool setTuning(const GPisMap3Tuning& t);    // before the first update() / after reset(); false once a map exists
    const GPisMap3Timing& lastTiming() const;
    void* cabiContext();                        // the gpis_ctx* underneath (tests / benches)
    // test hooks mirroring oracle/ref_harness.cpp: bulk-load samples (9 floats each) and train
    int insertSamples(const float* samples9, int n);
    int trainActive();
    int activateAll();                          // mark every non-empty leaf dirty (retrain the whole map with trainActive)
    int numLeaves();
    void getLeaves(std::vector<float>& centres, std::vector<int>& counts);
    void getAllSamples(std::vector<float>& samples9);

private:
    struct Impl;
    Impl* d;
};

#endif
