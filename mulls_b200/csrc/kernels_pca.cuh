// PCA neighbourhood features (pca.hpp:294-354) — kernels added after the ICP path.
#pragma once
#include "device_types.cuh"
