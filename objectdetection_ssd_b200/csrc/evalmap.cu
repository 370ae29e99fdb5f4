// VOC 11-point interpolated average precision on the GPU.  Reference: get_map, Util.py:783-885 (the direct consumer
// of inference()'s output; SURVEY.md section 8(f) "next" #1).
//
// One CTA per class.  (1) the class's detections are ranked by descending score (ties -> lower detection index, M1)
// with a bitonic sort of 64-bit keys; (2) a second sort by (image, rank) makes the detections of one image
// contiguous in rank order, and one thread per (class, image) group walks its detections greedily: best gt of the same
// image and class by IoU (ties -> first gt, M2), true positive iff IoU > thr and that gt is still unclaimed
// (Util.py:855-868); (3) cumulative TP over the rank order (block scan), precision = cumTP / (rank+1) and
// recall in fp64 exactly as the reference's numpy/torch mix computes them, and for each of the 11 recall levels the maximum precision among
// ranks whose recall reaches the level (Util.py:870-883).  IoU uses the reference's operation order (bit-exact).
#include <algorithm>
#include "common.cuh"

namespace ssdhead {

constexpr int AP_T = 256;

__device__ __forceinline__ void ap_bitonic_desc(unsigned long long* keys, int n_pad)
{
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ int ap_pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// workspace per class c (segment capacity capp = pow2ceil(N)): keyA[capp] u64, keyB[capp] u64, tp[capp] u8
__global__ void __launch_bounds__(AP_T)
voc_ap_kernel(const float4* __restrict__ det_box, const int* __restrict__ det_cls, const float* __restrict__ det_score,
              const int* __restrict__ det_off, int N, const float4* __restrict__ gt_box, const float* __restrict__ gt_cls,
              const int* __restrict__ gt_off, int M, int num_images, float iou_thr, const double* __restrict__ recall_levels,
              double* __restrict__ ap_out, unsigned long long* __restrict__ keyA_all, unsigned long long* __restrict__ keyB_all,
              unsigned char* __restrict__ tp_all, unsigned char* __restrict__ claimed /*[M], zero on entry*/, int capp)
{
    __shared__ int s_cnt, s_ngt;
    __shared__ unsigned int s_scan[AP_T / 32];
    __shared__ double s_best[11];
    const int c = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    unsigned long long* keyA = keyA_all + (size_t)c * capp;
    unsigned long long* keyB = keyB_all + (size_t)c * capp;
    unsigned char* tp = tp_all + (size_t)c * capp;
    if (t == 0) { s_cnt = 0; s_ngt = 0; }
    if (t < 11) s_best[t] = 0.0;
    __syncthreads();

    // ---- 0. this class's detections and ground-truth count ----
    int ngt = 0;
    for (int g = t; g < M; g += AP_T) ngt += ((int)gt_cls[g] == c) ? 1 : 0;
    if (ngt) atomicAdd(&s_ngt, ngt);
    for (int d = t; d < N; d += AP_T) {
        if (det_cls[d] == c) {
            const int slot = atomicAdd(&s_cnt, 1);
            keyA[slot] = ((unsigned long long)float_order_key(det_score[d]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)d);
        }
    }
    __syncthreads();
    const int n = s_cnt;
    if (n == 0) {                                    // Util.py:832-833: no detection of this class -> AP 0
        if (t == 0) ap_out[c] = 0.0;
        return;
    }
    const int n_pad = ap_pow2ceil(n);
    for (int i = n + t; i < n_pad; i += AP_T) keyA[i] = 0ull;
    __syncthreads();
    ap_bitonic_desc(keyA, n_pad);                    // rank order: descending score, ties -> lower detection index (M1)

    // ---- 1. group by image, keeping rank order inside a group: key = ~image << 32 | ~rank, sorted descending ----
    for (int r = t; r < n_pad; r += AP_T) {
        unsigned long long k = 0ull;
        if (r < n) {
            const int d = (int)(0xffffffffu - (unsigned)(keyA[r] & 0xffffffffull));
            int lo = 0, hi = num_images;             // image of detection d: last offset <= d
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (det_off[mid] <= d) lo = mid; else hi = mid; }
            k = ((unsigned long long)(0xffffffffu - (unsigned)lo) << 32) | (unsigned long long)(0xffffffffu - (unsigned)r);
        }
        keyB[r] = k;
        if (r < n) tp[r] = 0;
    }
    __syncthreads();
    ap_bitonic_desc(keyB, n_pad);

    // ---- 2. one thread per (class, image) group: greedy claim of the image's gts of this class (Util.py:835-868) ----
    for (int j = t; j < n; j += AP_T) {
        const unsigned img = 0xffffffffu - (unsigned)(keyB[j] >> 32);
        if (j > 0 && (0xffffffffu - (unsigned)(keyB[j - 1] >> 32)) == img) continue;       // not a group start
        const int g0 = gt_off[img], g1 = gt_off[img + 1];
        for (int q = j; q < n && (0xffffffffu - (unsigned)(keyB[q] >> 32)) == img; ++q) {
            const int r = (int)(0xffffffffu - (unsigned)(keyB[q] & 0xffffffffull));
            const int d = (int)(0xffffffffu - (unsigned)(keyA[r] & 0xffffffffull));
            const float4 db = det_box[d];
            const float da = box_area(db);
            float best = -INFINITY;
            int bestg = -1;
            for (int g = g0; g < g1; ++g) {
                if ((int)gt_cls[g] != c) continue;
                const float4 gb = gt_box[g];
                const float v = iou_xyxy(db, da, gb, box_area(gb));
                if (bestg < 0 || v > best) { best = v; bestg = g; }                        // M2: first maximal gt
            }
            if (bestg >= 0 && best > iou_thr && !claimed[bestg]) { claimed[bestg] = 1; tp[r] = 1; }
        }
    }
    __syncthreads();

    // ---- 3. cumulative TP in rank order, precision / recall in fp64, 11-point interpolation ----
    // Util.py:872 divides a numpy array by a 0-dim int64 tensor, which torch evaluates as reciprocal(tensor) * array
    // with a FLOAT32 reciprocal: recall = cumTP * float32(1 / #gt) (inf -> nan when the class has no gt)
    const double inv_nobj = (double)__fdiv_rn(1.0f, (float)s_ngt);
    double best[11];
#pragma unroll
    for (int q = 0; q < 11; ++q) best[q] = 0.0;
    unsigned running = 0u;
    for (int base = 0; base < n; base += AP_T) {
        const int r = base + t;
        const unsigned v = (r < n) ? (unsigned)tp[r] : 0u;
        unsigned inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned o = __shfl_up_sync(FULL, inc, d); if (lane >= d) inc += o; }
        __syncthreads();
        if (lane == 31) s_scan[warp] = inc;
        __syncthreads();
        unsigned wbase = 0u, total = 0u;
        for (int w = 0; w < AP_T / 32; ++w) { const unsigned cw = s_scan[w]; if (w < warp) wbase += cw; total += cw; }
        if (r < n) {
            const double ctp = (double)(running + wbase + inc);
            const double precision = ctp / (double)(r + 1);          // cum_TP / (cum_TP + cum_FP)
            const double recall = ctp * inv_nobj;                     // nan when the class has no gt: never >= a level
#pragma unroll
            for (int q = 0; q < 11; ++q) if (recall >= recall_levels[q]) best[q] = fmax(best[q], precision);
        }
        running += total;
    }
    __syncthreads();
    // block max per level (precision >= 0: the bit pattern of a non-negative double orders like the value)
#pragma unroll
    for (int q = 0; q < 11; ++q) {
        double m = best[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = fmax(m, __shfl_xor_sync(FULL, m, d));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned long long*>(&s_best[q]), (unsigned long long)__double_as_longlong(m));
    }
    __syncthreads();
    if (t == 0) {
        // numpy's mean of 11 values: pairwise over the first eight, then the last three in order
        const double* v = s_best;
        double s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        s += v[8]; s += v[9]; s += v[10];
        ap_out[c] = s / 11.0;
    }
}

size_t voc_ap_workspace_bytes(int N, int M, int num_fg)
{
    int capp = 1;
    while (capp < std::max(N, 1)) capp <<= 1;
    return (size_t)num_fg * capp * 17 + round_up((size_t)std::max(M, 1), 16) + 256;
}

}  // namespace ssdhead

using namespace ssdhead;

extern "C" {

size_t ssdhead_voc_ap_workspace_bytes(int N, int M, int num_fg) { return (N < 0 || M < 0 || num_fg <= 0) ? 0 : voc_ap_workspace_bytes(N, M, num_fg); }

int ssdhead_voc_ap(const float* det_boxes, const int32_t* det_cls, const float* det_score, const int32_t* det_off, int N,
                   const float* gt_boxes, const float* gt_cls, const int32_t* gt_off, int M, int num_images, int num_fg,
                   float iou_thr, const double* recall_levels, double* ap_out, void* ws, size_t ws_bytes, void* stream)
{
    if (N < 0 || M < 0 || num_images <= 0 || num_fg <= 0 || !det_off || !gt_off || !recall_levels || !ap_out || !ws) return SSDHEAD_E_BADARG;
    if ((N > 0 && (!det_boxes || !det_cls || !det_score)) || (M > 0 && (!gt_boxes || !gt_cls))) return SSDHEAD_E_BADARG;
    if ((N > 0 && !aligned16(det_boxes)) || (M > 0 && !aligned16(gt_boxes)) || !aligned16(ws)) return SSDHEAD_E_ALIGN;
    if (ws_bytes < voc_ap_workspace_bytes(N, M, num_fg)) return SSDHEAD_E_WORKSPACE;
    int capp = 1;
    while (capp < std::max(N, 1)) capp <<= 1;
    cudaStream_t st = (cudaStream_t)stream;
    char* w = (char*)ws;
    unsigned long long* keyA = (unsigned long long*)w;  w += (size_t)num_fg * capp * 8;
    unsigned long long* keyB = (unsigned long long*)w;  w += (size_t)num_fg * capp * 8;
    unsigned char* tp = (unsigned char*)w;              w += (size_t)num_fg * capp;
    unsigned char* claimed = (unsigned char*)w;
    SSD_CHECK_CUDA(cudaMemsetAsync(claimed, 0, (size_t)std::max(M, 1), st));
    voc_ap_kernel<<<num_fg, AP_T, 0, st>>>((const float4*)det_boxes, det_cls, det_score, det_off, N, (const float4*)gt_boxes, gt_cls,
                                           gt_off, M, num_images, iou_thr, recall_levels, ap_out, keyA, keyB, tp, claimed, capp);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
