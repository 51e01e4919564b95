// Newton-engine kernel instantiation: model set CDN_MS_DIODE, kernel kind CDN_KIND_BLOCK
// (one translation unit per instantiation: they compile in parallel).
#include "engine_launch.cuh"

CDN_DEFINE_LAUNCH(cdn_launch_diode_block, CDN_MS_DIODE, CDN_KIND_BLOCK)
