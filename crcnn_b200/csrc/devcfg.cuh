// Function attributes (dynamic shared-memory limit, carve-out) are per DEVICE, while the launchers are per process: a
// process that drives contexts on several GPUs (BASELINE config 4 runs one CRT instance per GPU) has to configure every
// kernel once on each of them.  `static DeviceOnce once; if (once.first()) cudaFuncSetAttribute(...)`.
#pragma once
#include <cuda_runtime.h>

namespace crcnn {

struct DeviceOnce {
    bool seen[64] = {};
    // true the first time it is called with a given current device
    bool first() {
        int d = 0;
        if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
        if (seen[d]) return false;
        seen[d] = true;
        return true;
    }
};

}  // namespace crcnn
