// functors.cuh — the elementwise forward/backward maps live in include/fastad_b200/device_math.cuh
// so that user translation units compiled with nvcc (fused.cuh) instantiate the very same code.
#pragma once
#include "common.cuh"
#include "../../include/fastad_b200/device_math.cuh"
