// Fused multibox loss: cross entropy + L1 + hard-negative mining, forward and gradients in
// ONE pass over the head outputs.  Reference: ssd / ssd1_, Losses.py:119-199.
//
// Layout.  One thread-block CLUSTER per image (grid = (CS, B), cluster = (CS,1,1)); CTA r of the
// cluster owns the contiguous prior range [r*chunk, r*chunk+n).  Its slice of the conf tensor
// (n x C fp32, e.g. 1092 x 21 = 91 728 B for SSD300 with CS = 8) is brought into shared memory
// once by the TMA engine (1-D cp.async.bulk, mbarrier completion) and never re-read from HBM:
//   1. while the copy is in flight the CTA re-derives the match of its priors (same IoU code as
//      match_kernel + the best_prior list) -> class / positive flag per prior;
//   2. thread-per-row log-softmax from shared memory (row stride 21 words: conflict-free) -> CE;
//   3. cluster-wide exact radix select (4 x 8-bit passes, histograms all-reduced through
//      distributed shared memory) of the k = 3*npos-th largest background CE; ties go to the
//      lower prior index (T4), positives take part with value 0 (Losses.py:190);
//   4. the rows are overwritten IN PLACE with the gradient ((softmax - onehot)/N for positives
//      and mined negatives, 0 elsewhere) and leave through one bulk shared->global store.
// loc is only touched at positive priors; grad_loc is written densely with coalesced float4.
// HBM traffic per image = conf once in, grad_conf + grad_loc once out (the roofline minimum);
// the batch-global normaliser N is known beforehand from match_kernel, which is why the match
// is a separate, tiny kernel.  Partial sums are reduced in fp64 in a fixed order by the last
// CTA to finish, so the loss values are run-to-run deterministic.
#include <cooperative_groups.h>
#include "common.cuh"

namespace cg = cooperative_groups;

namespace ssdhead {

constexpr int LT = 256;            // threads per CTA
constexpr int LGC = 64;            // gt boxes staged per chunk
constexpr uint32_t PIECE = 16384;  // bytes per bulk copy

struct LossParams {
    const float* loc;
    const float* conf;
    const float4* gt_xyxy;
    const float* gt_cls;
    const int* gt_off;
    const float4* pri_xyxy;
    const float4* pri_cxcywh;
    const int* best_prior;
    const int* npos;
    const int* npos_norm;
    int B, P, chunk, neg_ratio, bg_class, use_tma;
    float pos_iou;
    double* sums;
    float* losses;
    float* grad_loc;
    float* grad_conf;
    uint32_t* mined_mask;
    float* ce_out;
    double* partials;            // [B*CS][2]
    unsigned int* done_counter;  // self-resetting
};

template <int C, int KPT, bool GRADS>
__global__ void __launch_bounds__(LT, 2)
multibox_loss_kernel(const LossParams p)
{
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* s_conf = reinterpret_cast<float*>(smem_raw);

    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_hist_sum[2][256];
    __shared__ uint32_t s_hist_local[256];
    __shared__ float4 s_gbox[LGC];
    __shared__ float s_garea[LGC];
    __shared__ int s_gbp[LGC];
    __shared__ double s_redd[2][LT / 32];
    __shared__ uint32_t s_sel[3];
    __shared__ uint32_t s_tie[16];
    __shared__ int s_wscan[LT / 32];
    __shared__ int s_is_last;

    const int b = blockIdx.y, r = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int CS = gridDim.x;
    const int start = r * p.chunk;
    const int n = max(0, min(p.chunk, p.P - start));
    const size_t row0 = (size_t)b * p.P + start;
    const float* gconf = p.conf + row0 * C;

    // ---- 0. arm the barrier and start the conf copy ----
    if (t == 0) { mbar_init(&s_bar, 1); mbar_fence_init(); }
    s_hist_sum[0][t] = 0u;
    s_hist_sum[1][t] = 0u;
    __syncthreads();
    if (p.use_tma) {
        if (t == 0 && n > 0) {
            const uint32_t bytes = (uint32_t)n * C * 4u;
            mbar_expect_tx(&s_bar, bytes);
            for (uint32_t o = 0; o < bytes; o += PIECE)
                bulk_g2s(reinterpret_cast<char*>(s_conf) + o, reinterpret_cast<const char*>(gconf) + o,
                         min(PIECE, bytes - o), &s_bar);
        }
    } else {
        for (int idx = t; idx < n * C; idx += LT) s_conf[idx] = gconf[idx];
    }

    // ---- 1. match of my priors (Losses.py:150-171), overlapped with the copy ----
    const int off0 = p.gt_off[b];
    const int G = p.gt_off[b + 1] - off0;
    float best[KPT];
    int bestg[KPT], forced[KPT];
#pragma unroll
    for (int i = 0; i < KPT; ++i) { best[i] = -INFINITY; bestg[i] = 0; forced[i] = -1; }
    for (int g0 = 0; g0 < G; g0 += LGC) {
        const int gc = min(LGC, G - g0);
        __syncthreads();
        if (t < gc) {
            const float4 bx = p.gt_xyxy[off0 + g0 + t];
            s_gbox[t] = bx;
            s_garea[t] = box_area(bx);
            s_gbp[t] = p.best_prior[off0 + g0 + t];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int j = i * LT + t;
            if (j < n) {
                const float4 pb = p.pri_xyxy[start + j];
                const float pa = box_area(pb);
                for (int g = 0; g < gc; ++g) {
                    const float v = iou_xyxy(s_gbox[g], s_garea[g], pb, pa);
                    if (v > best[i]) { best[i] = v; bestg[i] = g0 + g; }     // T1
                    if (s_gbp[g] == start + j) forced[i] = g0 + g;            // T3: ascending g, last wins
                }
            }
        }
    }
    // the first cluster barrier also orders the zeroing of s_hist_sum before any remote add
    cluster.sync();

    // ---- 2. wait for the conf slice ----
    if (p.use_tma) { if (n > 0) mbar_wait(&s_bar, 0u); }
    else __syncthreads();

    // ---- 3. cross entropy per prior, L1 on positives ----
    const float nrm = (float)(*p.npos_norm);
    const float gs_conf = __fdiv_rn(1.0f, nrm);
    const float gs_loc = __fdiv_rn(1.0f, __fmul_rn(4.0f, nrm));
    uint32_t key[KPT];
    int cls[KPT];          // class per prior; bg_class = negative
    double acc_l1 = 0.0, acc_ce = 0.0;
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int j = i * LT + t;
        key[i] = 0u;
        cls[i] = p.bg_class;
        if (j < n) {
            const int obj = forced[i] >= 0 ? forced[i] : bestg[i];
            const bool hit = forced[i] >= 0 || (G > 0 && !(best[i] < p.pos_iou));   // T6
            const int c = hit ? (int)p.gt_cls[off0 + obj] : p.bg_class;
            const bool pos = c != p.bg_class;                                       // Losses.py:179
            cls[i] = c;
            const float* row = s_conf + (size_t)j * C;
            float m = row[0];
#pragma unroll
            for (int q = 1; q < C; ++q) m = fmaxf(m, row[q]);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < C; ++q) s = __fadd_rn(s, expf(__fsub_rn(row[q], m)));
            // -(x_c - max - log(sum)), the order ATen's log_softmax uses
            float ce = __fsub_rn(logf(s), __fsub_rn(row[c], m));
            ce = __fadd_rn(ce, 0.0f);
            if (p.ce_out) p.ce_out[row0 + j] = ce;
            float4 gl = make_float4(0.f, 0.f, 0.f, 0.f);
            if (pos) {
                acc_ce += (double)ce;
                const float4 tgt = encode_box(xyxy_to_cxcywh(p.gt_xyxy[off0 + obj]), p.pri_cxcywh[start + j]);
                const float4 l = reinterpret_cast<const float4*>(p.loc)[row0 + j];
                const float dx = __fsub_rn(l.x, tgt.x), dy = __fsub_rn(l.y, tgt.y);
                const float dz = __fsub_rn(l.z, tgt.z), dw = __fsub_rn(l.w, tgt.w);
                acc_l1 += (double)fabsf(dx) + (double)fabsf(dy) + (double)fabsf(dz) + (double)fabsf(dw);
                if (GRADS) {
                    gl.x = dx > 0.f ? gs_loc : (dx < 0.f ? -gs_loc : 0.f);
                    gl.y = dy > 0.f ? gs_loc : (dy < 0.f ? -gs_loc : 0.f);
                    gl.z = dz > 0.f ? gs_loc : (dz < 0.f ? -gs_loc : 0.f);
                    gl.w = dw > 0.f ? gs_loc : (dw < 0.f ? -gs_loc : 0.f);
                }
            } else {
                key[i] = __float_as_uint(ce);          // CE >= 0: the bit pattern is order preserving
            }
            if (GRADS) reinterpret_cast<float4*>(p.grad_loc)[row0 + j] = gl;
        }
    }

    // ---- 4. hard-negative mining: exact k-th largest over the cluster (Losses.py:188-195) ----
    const long long kk = (long long)p.neg_ratio * (long long)p.npos[b];
    const uint32_t k = (uint32_t)min((long long)p.P, max(0ll, kk));
    uint32_t selmask = 0u;               // bit i: prior i of this thread is selected by the ranking
    if (k >= (uint32_t)p.P) {
        selmask = (1u << KPT) - 1u;
    } else if (k > 0u) {
        uint32_t prefix = 0u, mask = 0u, need = k, cnt_eq = 0u;
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            uint32_t* hs = s_hist_sum[pass & 1];
            s_hist_local[t] = 0u;
            s_hist_sum[(pass + 1) & 1][t] = 0u;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const bool in = (i * LT + t < n) && ((key[i] & mask) == prefix);
                const uint32_t bin = (key[i] >> shift) & 255u;
                const unsigned act = __ballot_sync(FULL, in);
                if (in) {
                    const unsigned same = __match_any_sync(act, bin);
                    if ((unsigned)lane == (unsigned)(__ffs(same) - 1)) atomicAdd(&s_hist_local[bin], (uint32_t)__popc(same));
                }
            }
            __syncthreads();
            const uint32_t v = s_hist_local[t];
            if (v) {
                for (int dst = 0; dst < CS; ++dst) atomicAdd(cluster.map_shared_rank(&hs[t], dst), v);
            }
            cluster.sync();
            if (warp == 0) {
                uint32_t c8[8], tot = 0u;
#pragma unroll
                for (int q = 0; q < 8; ++q) { c8[q] = hs[lane * 8 + q]; tot += c8[q]; }
                uint32_t suf = tot;                       // inclusive suffix sum: bins of lanes >= me
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const uint32_t o = __shfl_down_sync(FULL, suf, d);
                    if (lane + d < 32) suf += o;
                }
                uint32_t above = suf - tot;
                if (above < need && suf >= need) {
#pragma unroll
                    for (int q = 7; q >= 0; --q) {
                        if (above + c8[q] >= need) { s_sel[0] = (uint32_t)(lane * 8 + q); s_sel[1] = above; s_sel[2] = c8[q]; break; }
                        above += c8[q];
                    }
                }
            }
            __syncthreads();
            prefix |= s_sel[0] << shift;
            mask |= 255u << shift;
            need -= s_sel[1];
            cnt_eq = s_sel[2];
        }
        const uint32_t T = prefix;       // the k-th largest key; `need` of the cnt_eq entries equal to T are taken
        uint32_t tiemask = 0u;
        if (need >= cnt_eq) {
            tiemask = (1u << KPT) - 1u;  // all ties are inside the top k
        } else {
            // ties cross the boundary: hand them out in prior order (T4)
            int mine = 0;
#pragma unroll
            for (int i = 0; i < KPT; ++i) mine += __syncthreads_count((i * LT + t < n) && key[i] == T);
            if (t == 0)
                for (int dst = 0; dst < CS; ++dst) *cluster.map_shared_rank(&s_tie[r], dst) = (uint32_t)mine;
            cluster.sync();
            uint32_t running = 0u;
            for (int q = 0; q < r; ++q) running += s_tie[q];
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const bool f = (i * LT + t < n) && key[i] == T;
                const unsigned ball = __ballot_sync(FULL, f);
                if (lane == 0) s_wscan[warp] = __popc(ball);
                __syncthreads();
                uint32_t wbase = 0u, tot = 0u;
                for (int w = 0; w < LT / 32; ++w) { const uint32_t cw = (uint32_t)s_wscan[w]; if (w < warp) wbase += cw; tot += cw; }
                const uint32_t rank = running + wbase + (uint32_t)__popc(ball & ((1u << lane) - 1u));
                if (f && rank < need) tiemask |= 1u << i;
                running += tot;
                __syncthreads();
            }
        }
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            if (i * LT + t < n) {
                if (key[i] > T || (key[i] == T && ((tiemask >> i) & 1u))) selmask |= 1u << i;
            }
        }
    }

    // ---- 5. mined CE sum; gradients in place ----
#pragma unroll
    for (int i = 0; i < KPT; ++i) {
        const int j = i * LT + t;
        if (j < n) {
            const bool pos = cls[i] != p.bg_class;
            const bool mined = !pos && ((selmask >> i) & 1u);
            if (mined) acc_ce += (double)__uint_as_float(key[i]);
            if (GRADS) {
                float* row = s_conf + (size_t)j * C;
                if (pos || mined) {
                    float m = row[0];
#pragma unroll
                    for (int q = 1; q < C; ++q) m = fmaxf(m, row[q]);
                    float e[C];
                    float s = 0.0f;
#pragma unroll
                    for (int q = 0; q < C; ++q) { e[q] = expf(__fsub_rn(row[q], m)); s = __fadd_rn(s, e[q]); }
                    const float inv = __fdiv_rn(1.0f, s);
#pragma unroll
                    for (int q = 0; q < C; ++q)
                        row[q] = __fmul_rn(__fsub_rn(__fmul_rn(e[q], inv), q == cls[i] ? 1.0f : 0.0f), gs_conf);
                } else {
#pragma unroll
                    for (int q = 0; q < C; ++q) row[q] = 0.0f;
                }
            }
        }
    }
    if (p.mined_mask) {
        const int words = (p.P + 31) / 32;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int j = i * LT + t;
            if (j < n && cls[i] == p.bg_class && ((selmask >> i) & 1u)) {
                const int pr = start + j;
                atomicOr(&p.mined_mask[(size_t)b * words + (pr >> 5)], 1u << (pr & 31));
            }
        }
    }

    // ---- 6. gradient slice out ----
    if (GRADS) {
        float* gdst = p.grad_conf + row0 * C;
        if (p.use_tma) {
            fence_proxy_async_smem();
            __syncthreads();
            if (t == 0 && n > 0) {
                const uint32_t bytes = (uint32_t)n * C * 4u;
                for (uint32_t o = 0; o < bytes; o += PIECE)
                    bulk_s2g(reinterpret_cast<char*>(gdst) + o, reinterpret_cast<const char*>(s_conf) + o, min(PIECE, bytes - o));
                bulk_commit();
            }
        } else {
            __syncthreads();
            for (int idx = t; idx < n * C; idx += LT) gdst[idx] = s_conf[idx];
        }
    }

    // ---- 7. loss sums: CTA partial -> last CTA reduces all partials in a fixed order ----
    acc_l1 = warp_sum(acc_l1);
    acc_ce = warp_sum(acc_ce);
    if (lane == 0) { s_redd[0][warp] = acc_l1; s_redd[1][warp] = acc_ce; }
    __syncthreads();
    if (t == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < LT / 32; ++w) { a += s_redd[0][w]; c += s_redd[1][w]; }
        const size_t slot = (size_t)b * CS + r;
        p.partials[2 * slot] = a;
        p.partials[2 * slot + 1] = c;
        __threadfence();
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned done = atomicAdd(p.done_counter, 1u);
        s_is_last = (done == total - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        const int total = gridDim.x * gridDim.y;
        double a = 0.0, c = 0.0;
        for (int s = t; s < total; s += LT) {
            a += __ldcg(&p.partials[2 * s]);
            c += __ldcg(&p.partials[2 * s + 1]);
        }
        a = warp_sum(a);
        c = warp_sum(c);
        __syncthreads();
        if (lane == 0) { s_redd[0][warp] = a; s_redd[1][warp] = c; }
        __syncthreads();
        if (t == 0) {
            a = 0.0; c = 0.0;
            for (int w = 0; w < LT / 32; ++w) { a += s_redd[0][w]; c += s_redd[1][w]; }
            p.sums[0] = a;
            p.sums[1] = c;
            const double N = (double)(*p.npos_norm);
            p.losses[0] = (float)(a / (4.0 * N));
            p.losses[1] = (float)(c / N);
            *p.done_counter = 0u;
        }
    }

    // ---- 8. the shared-memory slice must outlive the bulk store's reads ----
    if (GRADS && p.use_tma && t == 0 && n > 0) bulk_wait_read_all();
}

__global__ void finish_loss_kernel(const double* __restrict__ sums, const int* __restrict__ npos_norm, float* __restrict__ losses)
{
    const double N = (double)(*npos_norm);
    losses[0] = (float)(sums[0] / (4.0 * N));
    losses[1] = (float)(sums[1] / N);
}

__global__ void __launch_bounds__(256)
scale_grads_kernel(float4* __restrict__ gl, size_t n4_loc, float* __restrict__ gl_tail, int tail_loc,
                   float4* __restrict__ gc, size_t n4_conf, float* __restrict__ gc_tail, int tail_conf,
                   const float* __restrict__ gout)
{
    const float a = gout[0], c = gout[1];
    if (a == 1.0f && c == 1.0f) return;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a != 1.0f) {
        for (size_t i = tid; i < n4_loc; i += stride) { float4 v = gl[i]; v.x *= a; v.y *= a; v.z *= a; v.w *= a; gl[i] = v; }
        if (tid < (size_t)tail_loc) gl_tail[tid] *= a;
    }
    if (c != 1.0f) {
        for (size_t i = tid; i < n4_conf; i += stride) { float4 v = gc[i]; v.x *= c; v.y *= c; v.z *= c; v.w *= c; gc[i] = v; }
        if (tid < (size_t)tail_conf) gc_tail[tid] *= c;
    }
}

struct LossPlan { int cs, chunk, kpt; size_t smem; };

static bool plan_loss(int P, int C, LossPlan* out)
{
    // smallest cluster whose per-CTA slice allows two CTAs per SM; otherwise the smallest that fits at all
    const size_t two_per_sm = 100 * 1024, one_per_sm = 200 * 1024;
    int pick = 0;
    for (int pass = 0; pass < 2 && !pick; ++pass) {
        for (int cs = 1; cs <= 16; cs <<= 1) {
            const int chunk = (int)round_up((size_t)(P + cs - 1) / cs, 4);
            const size_t bytes = (size_t)chunk * C * 4;
            if (chunk > LT * 8) continue;
            if (bytes <= (pass == 0 ? two_per_sm : one_per_sm) && (pass == 1 || cs <= 8)) { pick = cs; break; }
        }
    }
    if (!pick) return false;
    out->cs = pick;
    out->chunk = (int)round_up((size_t)(P + pick - 1) / pick, 4);
    out->kpt = out->chunk <= LT * 5 ? 5 : 8;
    out->smem = round_up((size_t)out->chunk * C * 4, 128);
    return true;
}

template <int C, int KPT, bool GRADS>
static int launch_loss(const LossParams& prm, const LossPlan& plan, cudaStream_t st)
{
    auto kern = multibox_loss_kernel<C, KPT, GRADS>;
    SSD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    if (plan.cs > 8) SSD_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan.cs, prm.B, 1);
    cfg.blockDim = dim3(LT, 1, 1);
    cfg.dynamicSmemBytes = plan.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = plan.cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    SSD_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, prm));
    count_launch();
    return 0;
}

size_t loss_workspace_bytes(int B, int P, int C)
{
    LossPlan plan;
    if (!plan_loss(P, C, &plan)) return 0;
    return round_up((size_t)B * plan.cs * 2 * sizeof(double), 16) + 16;
}

}  // namespace ssdhead

using namespace ssdhead;

extern "C" {

int ssdhead_multibox_loss(const float* loc, const float* conf,
                          const float* gt_xyxy, const float* gt_cls, const int32_t* gt_off,
                          const float* pri_xyxy, const float* pri_cxcywh,
                          const int32_t* best_prior, const int32_t* npos, const int32_t* npos_norm,
                          int B, int P, int C, int neg_ratio, float pos_iou,
                          double* sums, float* losses, float* grad_loc, float* grad_conf,
                          uint32_t* mined_mask, float* ce,
                          void* ws, size_t ws_bytes, void* stream)
{
    if (B < 0 || P <= 0 || neg_ratio < 0) return SSDHEAD_E_BADARG;
    if (!loc || !conf || !gt_off || !pri_xyxy || !pri_cxcywh || !npos || !npos_norm || !sums || !losses || !ws)
        return SSDHEAD_E_BADARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return SSDHEAD_E_BADARG;
    if (C != 21) return SSDHEAD_E_UNSUPPORTED;          // VOC head of the reference (Losses.py:184 hard-codes 21)
    if (B == 0) return 0;
    if (B > 65535) return SSDHEAD_E_UNSUPPORTED;
    if (!aligned16(loc) || !aligned16(pri_xyxy) || !aligned16(pri_cxcywh) || (gt_xyxy && !aligned16(gt_xyxy)) ||
        (grad_loc && !aligned16(grad_loc)) || !aligned16(ws))
        return SSDHEAD_E_ALIGN;
    LossPlan plan;
    if (!plan_loss(P, C, &plan)) return SSDHEAD_E_UNSUPPORTED;
    const size_t need = loss_workspace_bytes(B, P, C);
    if (ws_bytes < need) return SSDHEAD_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;

    LossParams prm;
    prm.loc = loc; prm.conf = conf; prm.gt_xyxy = (const float4*)gt_xyxy; prm.gt_cls = gt_cls; prm.gt_off = gt_off;
    prm.pri_xyxy = (const float4*)pri_xyxy; prm.pri_cxcywh = (const float4*)pri_cxcywh;
    prm.best_prior = best_prior; prm.npos = npos; prm.npos_norm = npos_norm;
    prm.B = B; prm.P = P; prm.chunk = plan.chunk; prm.neg_ratio = neg_ratio; prm.bg_class = C - 1;
    prm.pos_iou = pos_iou;
    // bulk copies need 16-byte aligned slices: the conf base, the per-image stride and every chunk start
    prm.use_tma = (aligned16(conf) && (grad_conf == nullptr || aligned16(grad_conf)) && (P % 4 == 0)) ? 1 : 0;
    prm.sums = sums; prm.losses = losses; prm.grad_loc = grad_loc; prm.grad_conf = grad_conf;
    prm.mined_mask = mined_mask; prm.ce_out = ce;
    prm.done_counter = (unsigned int*)ws;
    prm.partials = (double*)((char*)ws + 16);
    if (mined_mask) SSD_CHECK_CUDA(cudaMemsetAsync(mined_mask, 0, (size_t)B * ((P + 31) / 32) * sizeof(uint32_t), st));

    const bool grads = grad_loc != nullptr;
    if (plan.kpt == 5) return grads ? launch_loss<21, 5, true>(prm, plan, st) : launch_loss<21, 5, false>(prm, plan, st);
    return grads ? launch_loss<21, 8, true>(prm, plan, st) : launch_loss<21, 8, false>(prm, plan, st);
}

int ssdhead_finish_loss(const double* sums, const int32_t* npos_norm, float* losses, void* stream)
{
    if (!sums || !npos_norm || !losses) return SSDHEAD_E_BADARG;
    finish_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(sums, npos_norm, losses);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

int ssdhead_scale_grads(float* grad_loc, size_t n_loc, float* grad_conf, size_t n_conf, const float* gout, void* stream)
{
    if (!grad_loc || !grad_conf || !gout) return SSDHEAD_E_BADARG;
    if (!aligned16(grad_loc) || !aligned16(grad_conf)) return SSDHEAD_E_ALIGN;
    const size_t n4l = n_loc / 4, n4c = n_conf / 4;
    scale_grads_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(
        (float4*)grad_loc, n4l, grad_loc + n4l * 4, (int)(n_loc - n4l * 4),
        (float4*)grad_conf, n4c, grad_conf + n4c * 4, (int)(n_conf - n4c * 4), gout);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
