#include "common.cuh"
extern "C" {
int  ssdhead_ctx_create(ssdhead_ctx**, int, int, int, int, int, int, const float*) { return SSDHEAD_E_UNSUPPORTED; }
void ssdhead_ctx_destroy(ssdhead_ctx*) {}
void* ssdhead_host_alloc(size_t bytes) { void* p = nullptr; return cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess ? p : nullptr; }
void  ssdhead_host_free(void* p) { if (p) cudaFreeHost(p); }
int ssdhead_ctx_multibox_loss_host(ssdhead_ctx*, const float*, const float*, const float*, const float*, const int32_t*,
                                   int, int, float, float*, float*, float*) { return SSDHEAD_E_UNSUPPORTED; }
int ssdhead_ctx_detect_host(ssdhead_ctx*, const float*, const float*, int, float, float,
                            float*, float*, int32_t*, int32_t*, int32_t*) { return SSDHEAD_E_UNSUPPORTED; }
}
