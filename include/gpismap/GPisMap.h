// Drop-in GPisMap (2D): same public class API, parameter struct and defaults as the reference
// (cpp/include/GPisMap.h:29-119), hot path on a B200 through include/gpis_b200.h.
#ifndef GPISMAP_B200_GPISMAP_H
#define GPISMAP_B200_GPISMAP_H

#include <vector>

#include "params.h"

typedef struct GPisMapParam_ {   // cpp/include/GPisMap.h:29-67
    float delx;
    float fbias;
    float sensor_offset[2];
    float angle_obs_limit[2];
    float obs_var_thre;
    float min_position_noise;
    float min_grad_noise;
    float map_scale_param;
    float map_noise_param;
    GPisMapParam_()
        : delx((float)gpismap_defaults::kDelx2), fbias((float)gpismap_defaults::kFbias2),
          obs_var_thre((float)gpismap_defaults::kObsVarThre2),
          min_position_noise((float)gpismap_defaults::kMinPosNoise2),
          min_grad_noise((float)gpismap_defaults::kMinGradNoise2),
          map_scale_param((float)gpismap_defaults::kMapScale2), map_noise_param((float)gpismap_defaults::kMapNoise2) {
        sensor_offset[0] = (float)gpismap_defaults::kSensorOffset0;
        sensor_offset[1] = (float)gpismap_defaults::kSensorOffset1;
        angle_obs_limit[0] = (float)gpismap_defaults::kAngleObsLimit0;
        angle_obs_limit[1] = (float)gpismap_defaults::kAngleObsLimit1;
    }
} GPisMapParam;

struct GPisMapTiming {
    double phase[5];
    int valid_beams, active_leaves, trained_leaves;
    float train_kernel_ms;
};

class GPisMap {
public:
    GPisMap();
    explicit GPisMap(GPisMapParam par);
    ~GPisMap();
    GPisMap(const GPisMap&) = delete;
    GPisMap& operator=(const GPisMap&) = delete;

    void reset();
    void resetMap() { reset(); }
    // datax: bearing (rad), dataf: range; pose = [t(2) | R col-major(4)]
    void update(float* datax, float* dataf, int N, std::vector<float>& pose);
    // x: 2 x leng interleaved; res: 6 x leng, read-modify-write [f, gx, gy, var_f, var_gx, var_gy]
    bool test(float* x, int dim, int leng, float* res);
    int getMapDimension() { return 2; }

    // ---- additions (not in the reference API)
    void setDevice(int cuda_device);
    const GPisMapTiming& lastTiming() const;
    void* cabiContext();
    void getAllPoints(std::vector<float>& pos);
    int insertSamples(const float* samples7, int n);
    int trainActive();
    int activateAll();                          // mark every non-empty leaf dirty (retrain the whole map with trainActive)
    int numLeaves();
    void getLeaves(std::vector<float>& centres, std::vector<int>& counts);
    void getAllSamples(std::vector<float>& samples7);

private:
    struct Impl;
    Impl* d;
};

#endif
