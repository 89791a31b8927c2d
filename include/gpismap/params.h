// Default hyper-parameters of the drop-in GPisMap / GPisMap3 classes. Values and meaning follow the
// reference's macros (cpp/include/params.h:27-110); here they are typed constants that seed the
// runtime parameter structs, so configurations with other leaf sizes do not need a rebuild.
#pragma once

namespace gpismap_defaults {

// ---- tree (params.h:34-44). Half-lengths must be (common base) * 2^k.
constexpr double kTree2ClusterHalf = 0.8, kTree2MinHalf = 0.2, kTree2MaxHalf = 102.4, kTree2InitRootHalf = 12.8;
constexpr double kTree3ClusterHalf = 0.025, kTree3MinHalf = 0.0125 / 2.0, kTree3MaxHalf = 1.6, kTree3InitRootHalf = 0.4;
constexpr double kRtimes3 = 2.0;   // training-ball radius = Rtimes * cluster half (params.h:40)
constexpr double kRtimes2 = 4.0;   // GPisMap.cpp:583,608

// ---- GPisMap (2D), params.h:64-74
constexpr double kDelx2 = 1e-2, kFbias2 = 0.2, kObsVarThre2 = 0.1;
constexpr double kSensorOffset0 = 0.08, kSensorOffset1 = 0.0;
constexpr double kAngleObsLimit0 = -135.0 * 3.14159265358979323846 / 180.0;
constexpr double kAngleObsLimit1 = 135.0 * 3.14159265358979323846 / 180.0;
constexpr double kMinPosNoise2 = 1e-2, kMinGradNoise2 = 1e-2, kMapScale2 = 1.2, kMapNoise2 = 1e-2;
constexpr double kMaxRange2 = 3e1, kMinRange2 = 2e-1;   // GPisMap.cpp:31-32

// ---- GPisMap3 (3D), params.h:77-93
constexpr double kMaxRange3 = 4e0, kMinRange3 = 4e-1;
constexpr double kDelx3 = 1e-3, kFbias3 = 0.2, kObsVarThre3 = 0.04;
constexpr int kObsSkip3 = 2;
constexpr double kMinPosNoise3 = 1e-3, kMinGradNoise3 = 1e-2, kMapScale3 = 0.04, kMapNoise3 = 5e-3;

}  // namespace gpismap_defaults
