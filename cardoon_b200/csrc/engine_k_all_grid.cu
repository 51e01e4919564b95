// Newton-engine kernel instantiation: model set CDN_MS_ALL, kernel kind CDN_KIND_GRID
// (one translation unit per instantiation: they compile in parallel).
#include "engine_launch.cuh"

CDN_DEFINE_LAUNCH(cdn_launch_all_grid, CDN_MS_ALL, CDN_KIND_GRID)
