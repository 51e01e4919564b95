// Newton-engine kernel instantiation: model set CDN_MS_BSIM3, kernel kind CDN_KIND_TILE16
// (one translation unit per instantiation: they compile in parallel).
#include "engine_launch.cuh"

CDN_DEFINE_LAUNCH(cdn_launch_bsim3_tile16, CDN_MS_BSIM3, CDN_KIND_TILE16)
