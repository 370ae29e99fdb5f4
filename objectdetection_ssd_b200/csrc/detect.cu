#include "common.cuh"
namespace ssdhead { size_t detect_workspace_bytes(int B, int P, int C, int n) { return 16; } }
extern "C" {
int ssdhead_detect(const float*, const float*, const float*, int, int, int, float, float, int,
                   float*, float*, int32_t*, int32_t*, int32_t*, void*, size_t, void*) { return SSDHEAD_E_UNSUPPORTED; }
int ssdhead_detect_from_scores(const float*, const float*, int, int, int, float, float, int,
                   float*, float*, int32_t*, int32_t*, int32_t*, void*, size_t, void*) { return SSDHEAD_E_UNSUPPORTED; }
}
