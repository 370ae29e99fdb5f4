// Matching of ground-truth boxes against the prior set, dense IoU, and the box-format /
// offset elementwise ops.  Reference: Losses.py:150-171, Util.py:57-63, 86-102, 252-265,
// 288-301, 333-352.
//
// match_kernel: grid (tiles, B), 256 threads, 4 priors per thread (coalesced float4 reads of
// the L2-resident prior table).  The gt boxes of the image are staged in shared memory; each
// thread keeps the best gt of its priors (T1: first maximal gt), and per gt a warp-level
// redux.max + ballot gives the best prior of the tile (T2: lowest prior index), merged across
// warps and tiles with 64-bit atomicMax on (iou_key << 32 | ~prior).  The last tile of an image
// to finish applies the forced-match override (T3: highest gt index wins) and publishes
// best_prior / npos.  HBM traffic is negligible (gt + priors); the kernel exists so the loss
// kernel can know the batch-global positive count before it writes gradients.
#include "common.cuh"

namespace ssdhead {

constexpr int MT = 256;          // threads per CTA
constexpr int MPPT = 4;          // priors per thread
constexpr int MTILE = MT * MPPT; // priors per CTA
constexpr int MGC = 64;          // gt boxes staged per chunk

__global__ void __launch_bounds__(MT)
match_kernel(const float4* __restrict__ gt_xyxy, const float* __restrict__ gt_cls, const int* __restrict__ gt_off,
             const float4* __restrict__ pri_xyxy, int B, int P, int bg_class, float pos_iou,
             int* __restrict__ best_prior, int* __restrict__ npos, int* __restrict__ obj_idx, int* __restrict__ cls_out,
             unsigned long long* __restrict__ best_key, unsigned int* __restrict__ tile_counter)
{
    __shared__ float4 s_box[MGC];
    __shared__ float s_area[MGC];
    __shared__ unsigned long long s_key[MGC];
    __shared__ int s_red[MT / 32];
    __shared__ int s_last;

    const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int off0 = gt_off[b];
    const int G = gt_off[b + 1] - off0;

    float4 pb[MPPT];
    float pa[MPPT], best[MPPT];
    int bestg[MPPT];
    bool valid[MPPT];
#pragma unroll
    for (int i = 0; i < MPPT; ++i) {
        const int p = tile * MTILE + i * MT + t;
        valid[i] = p < P;
        pb[i] = valid[i] ? pri_xyxy[p] : make_float4(0.f, 0.f, 0.f, 0.f);
        pa[i] = box_area(pb[i]);
        best[i] = -INFINITY;
        bestg[i] = 0;
    }

    for (int g0 = 0; g0 < G; g0 += MGC) {
        const int gc = min(MGC, G - g0);
        __syncthreads();
        if (t < gc) {
            const float4 bx = gt_xyxy[off0 + g0 + t];
            s_box[t] = bx;
            s_area[t] = box_area(bx);
            s_key[t] = 0ull;
        }
        __syncthreads();
        for (int g = 0; g < gc; ++g) {
            const float4 gb = s_box[g];
            const float ga = s_area[g];
            uint32_t wbest = 0u, wprior = 0u;
#pragma unroll
            for (int i = 0; i < MPPT; ++i) {
                const float v = iou_xyxy(gb, ga, pb[i], pa[i]);
                if (v > best[i]) { best[i] = v; bestg[i] = g0 + g; }          // T1: strict > keeps the first
                const uint32_t k = valid[i] ? float_order_key(v) : 0u;
                const uint32_t m = __reduce_max_sync(FULL, k);
                if (m > wbest) {                                                // warp-uniform; strict > keeps lower i
                    const unsigned ball = __ballot_sync(FULL, k == m);
                    wbest = m;
                    wprior = (uint32_t)(tile * MTILE + i * MT + warp * 32 + (__ffs(ball) - 1));   // T2: lowest lane
                }
            }
            if (lane == 0 && wbest != 0u)
                atomicMax(&s_key[g], ((unsigned long long)wbest << 32) | (unsigned long long)(0xffffffffu - wprior));
        }
        __syncthreads();
        if (t < gc && s_key[t] != 0ull) atomicMax(&best_key[off0 + g0 + t], s_key[t]);
    }

    // natural (pre-override) match of this tile
    int cnt = 0;
#pragma unroll
    for (int i = 0; i < MPPT; ++i) {
        const int p = tile * MTILE + i * MT + t;
        if (!valid[i]) continue;
        const bool hit = (G > 0) && !(best[i] < pos_iou);                       // T6: matched <=> not (iou < thr)
        const int c = hit ? (int)gt_cls[off0 + bestg[i]] : bg_class;
        cnt += (c != bg_class) ? 1 : 0;                                         // positive <=> class != bg (Losses.py:179)
        if (obj_idx) obj_idx[(size_t)b * P + p] = off0 + bestg[i];
        if (cls_out) cls_out[(size_t)b * P + p] = c;
    }
    cnt = warp_sum(cnt);
    if (lane == 0) s_red[warp] = cnt;
    __syncthreads();
    if (t == 0) {
        int c = 0;
        for (int w = 0; w < MT / 32; ++w) c += s_red[w];
        if (c) atomicAdd(&npos[b], c);
        __threadfence();
        const unsigned done = atomicAdd(&tile_counter[b], 1u);
        s_last = (done == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last tile of image b: forced-match override (Losses.py:164-167) ----
    int extra = 0;
    for (int g = t; g < G; g += MT) {
        const uint32_t p = 0xffffffffu - (uint32_t)(ld_cg_u64(&best_key[off0 + g]) & 0xffffffffull);
        best_prior[off0 + g] = (int)p;
        bool winner = true;                                                     // T3: the highest gt index keeps the prior
        for (int g2 = g + 1; g2 < G; ++g2) {
            const uint32_t p2 = 0xffffffffu - (uint32_t)(ld_cg_u64(&best_key[off0 + g2]) & 0xffffffffull);
            if (p2 == p) { winner = false; break; }
        }
        if (!winner) continue;
        const float4 pbx = pri_xyxy[p];
        const float pax = box_area(pbx);
        float nb = -INFINITY;
        int ng = 0;
        for (int g2 = 0; g2 < G; ++g2) {
            const float4 gb = gt_xyxy[off0 + g2];
            const float v = iou_xyxy(gb, box_area(gb), pbx, pax);
            if (v > nb) { nb = v; ng = g2; }
        }
        const int c_nat = !(nb < pos_iou) ? (int)gt_cls[off0 + ng] : bg_class;  // what the tile pass counted
        const int c_new = (int)gt_cls[off0 + g];
        extra += (c_new != bg_class ? 1 : 0) - (c_nat != bg_class ? 1 : 0);
        if (obj_idx) obj_idx[(size_t)b * P + p] = off0 + g;
        if (cls_out) cls_out[(size_t)b * P + p] = c_new;
    }
    extra = warp_sum(extra);
    __syncthreads();
    if (lane == 0) s_red[warp] = extra;
    __syncthreads();
    if (t == 0) {
        int e = 0;
        for (int w = 0; w < MT / 32; ++w) e += s_red[w];
        const int total = atomicAdd(&npos[b], e) + e;
        atomicAdd(&npos[B], total);
    }
}

// ------------------------------------------------------------------------------- dense IoU
__global__ void __launch_bounds__(256)
iou_matrix_kernel(const float4* __restrict__ a, int n1, const float4* __restrict__ bxs, int n2, float* __restrict__ out)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * 8;
    if (j >= n2) return;
    const float4 pb = bxs[j];
    const float pa = box_area(pb);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = i0 + r;
        if (i >= n1) break;
        const float4 gb = a[i];
        out[(size_t)i * n2 + j] = iou_xyxy(gb, box_area(gb), pb, pa);
    }
}

// ------------------------------------------------------------------------------- elementwise box ops
template <int OP>
__global__ void __launch_bounds__(256)
box_op_kernel(const float4* __restrict__ in, const float4* __restrict__ pri, float4* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v = in[i];
    float4 r;
    if (OP == 0) r = cxcywh_to_xyxy(v);
    else if (OP == 1) r = xyxy_to_cxcywh(v);
    else if (OP == 2) r = encode_box(v, pri[i]);
    else r = decode_box(v, pri[i]);
    out[i] = r;
}

template <int OP>
static int launch_box_op(const float* in, const float* pri, float* out, int n, void* stream)
{
    if (n < 0 || (n > 0 && (!in || !out || (OP >= 2 && !pri)))) return SSDHEAD_E_BADARG;
    if (n == 0) return 0;
    if (!aligned16(in) || !aligned16(out) || (OP >= 2 && !aligned16(pri))) return SSDHEAD_E_ALIGN;
    box_op_kernel<OP><<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        (const float4*)in, (const float4*)pri, (float4*)out, n);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

}  // namespace ssdhead

using namespace ssdhead;

extern "C" {

int ssdhead_cxcywh_to_xyxy(const float* in, float* out, int n, void* stream) { return launch_box_op<0>(in, nullptr, out, n, stream); }
int ssdhead_xyxy_to_cxcywh(const float* in, float* out, int n, void* stream) { return launch_box_op<1>(in, nullptr, out, n, stream); }
int ssdhead_encode(const float* c, const float* p, float* out, int n, void* stream) { return launch_box_op<2>(c, p, out, n, stream); }
int ssdhead_decode(const float* g, const float* p, float* out, int n, void* stream) { return launch_box_op<3>(g, p, out, n, stream); }

int ssdhead_iou_matrix(const float* a, int n1, const float* b, int n2, float* out, void* stream)
{
    if (n1 < 0 || n2 < 0) return SSDHEAD_E_BADARG;
    if (n1 == 0 || n2 == 0) return 0;
    if (!a || !b || !out) return SSDHEAD_E_BADARG;
    if (!aligned16(a) || !aligned16(b)) return SSDHEAD_E_ALIGN;
    dim3 grid((n2 + 255) / 256, (n1 + 7) / 8);
    if (grid.y > 65535) return SSDHEAD_E_UNSUPPORTED;
    iou_matrix_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const float4*)a, n1, (const float4*)b, n2, out);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

int ssdhead_match(const float* gt_xyxy, const float* gt_cls, const int32_t* gt_off, const float* pri_xyxy,
                  int B, int P, int C, int sumG, float pos_iou,
                  int32_t* best_prior, int32_t* npos, int32_t* obj_idx, int32_t* cls,
                  void* ws, size_t ws_bytes, void* stream)
{
    if (B < 0 || P <= 0 || C < 2 || sumG < 0) return SSDHEAD_E_BADARG;
    if (B == 0) return 0;
    if (!gt_off || !pri_xyxy || !npos || !ws) return SSDHEAD_E_BADARG;
    if (sumG > 0 && (!gt_xyxy || !gt_cls || !best_prior)) return SSDHEAD_E_BADARG;
    if (B > 65535) return SSDHEAD_E_UNSUPPORTED;
    if (!aligned16(pri_xyxy) || (sumG > 0 && !aligned16(gt_xyxy)) || !aligned16(ws)) return SSDHEAD_E_ALIGN;
    const size_t need = ssdhead_workspace_bytes(SSDHEAD_WS_MATCH, B, P, C, sumG);
    if (ws_bytes < need) return SSDHEAD_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* best_key = (unsigned long long*)ws;
    unsigned int* counter = (unsigned int*)((char*)ws + round_up((size_t)sumG * 8, 16));
    SSD_CHECK_CUDA(cudaMemsetAsync(ws, 0, need, st));
    SSD_CHECK_CUDA(cudaMemsetAsync(npos, 0, (size_t)(B + 1) * sizeof(int), st));
    dim3 grid((P + MTILE - 1) / MTILE, B);
    match_kernel<<<grid, MT, 0, st>>>((const float4*)gt_xyxy, gt_cls, gt_off, (const float4*)pri_xyxy, B, P, C - 1, pos_iou,
                                      best_prior, npos, obj_idx, cls, best_key, counter);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
