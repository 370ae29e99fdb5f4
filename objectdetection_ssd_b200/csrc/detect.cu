// Detection post-processing: decode + softmax + score threshold, per-class sort + greedy NMS, global top-k.
// Reference: inference(), Losses.py:11-98 (decode Util.py:86-91, corner form Util.py:93-96, IoU Util.py:252-301).
//
//   detect_score_kernel  grid (row tiles, B): the tile's conf rows are staged coalesced into shared memory,
//                        one thread per prior does softmax (or takes given probabilities), decodes the box to
//                        corner form once, and appends (prob, prior) keys to the candidate list of every
//                        foreground class whose prob >= min_score (one warp-aggregated atomic per class).
//   detect_nms_kernel    one CTA per (class, image).  The sweep stops once top_k boxes are kept (kept boxes come out
//                        in descending score order, so only the first top_k of a class can reach the global top-k:
//                        the result is unchanged and the O(n^2) tail of the reference's loop is never executed), so
//                        the candidate list is consumed in SLICES of descending score found with a histogram over
//                        linear probability bins; each slice is bitonic-sorted on the 64-bit keys (prob bits << 32 |
//                        ~prior: descending prob, ties -> lower prior, T5) and swept in blocks of 64: a block is first
//                        tested against the boxes kept so far (kept boxes live in shared memory), then resolved
//                        internally with a 64x64 suppression bit mask.  The IoU test avoids the division unless the
//                        ratio is within 2^-20 of the threshold, where the reference's exact `inter/union >= thr` is
//                        evaluated (bit-exact keep lists).
//   detect_topk_kernel   one CTA per image: class-major concatenation (Losses.py:71-73) or, if more than top_k
//                        survive, the top_k by descending prob with ties to the earlier class-major position (T7);
//                        two lower bounds on the top_k-th score prune the 20 sorted lists before the merge sort.
#include <algorithm>
#include "common.cuh"

namespace ssdhead {

constexpr int DT = 256;

__host__ __device__ inline int pow2ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct DetectWs {
    float4* boxes;                 // [B*P] decoded corner boxes
    unsigned long long* cand;      // [B*NF*CAPP] candidate keys; after NMS the kept keys sit at the front of each segment
    unsigned int* cand_cnt;        // [B*NF]  zero on entry, zero on exit
    unsigned int* kept_cnt;        // [B*NF]
    unsigned int* overflow;        // [B]     zero on entry, zero on exit
};

static size_t detect_ws_layout(int B, int P, int C, int n, DetectWs* w, void* base)
{
    const int NF = C - 1;
    const int capp = pow2ceil(n > 0 ? std::min(n, P) : P);
    size_t off = 0;
    char* b = (char*)base;
    auto take = [&](size_t bytes) { size_t o = off; off += round_up(bytes, 256); return b ? (void*)(b + o) : nullptr; };
    void* p0 = take((size_t)B * P * 16);
    void* p1 = take((size_t)B * NF * capp * 8);
    void* p2 = take((size_t)B * NF * 4);
    void* p3 = take((size_t)B * NF * 4);
    void* p4 = take((size_t)B * 4);
    if (w) { w->boxes = (float4*)p0; w->cand = (unsigned long long*)p1; w->cand_cnt = (unsigned int*)p2;
             w->kept_cnt = (unsigned int*)p3; w->overflow = (unsigned int*)p4; }
    return off;
}

size_t detect_workspace_bytes(int B, int P, int C, int n)
{
    if (B <= 0 || P <= 0 || C < 2) return 0;
    return detect_ws_layout(B, P, C, n, nullptr, nullptr);
}

// ------------------------------------------------------------------------------------------------
template <int C, bool FROM_SCORES>
__global__ void __launch_bounds__(DT)
detect_score_kernel(const float* __restrict__ loc, const float* __restrict__ conf, const float4* __restrict__ pri_cxcywh,
                    int P, float min_score, int cap, int capp,
                    float4* __restrict__ boxes_out, unsigned long long* __restrict__ cand,
                    unsigned int* __restrict__ cand_cnt, unsigned int* __restrict__ overflow)
{
    constexpr int NF = C - 1;
    __shared__ float s_conf[DT * C];
    const int b = blockIdx.y, tile = blockIdx.x, t = threadIdx.x, lane = t & 31;
    const int r0 = tile * DT;
    const int nrows = min(DT, P - r0);
    const size_t base = ((size_t)b * P + r0) * C;
    for (int i = t; i < nrows * C; i += DT) s_conf[i] = __ldg(conf + base + i);
    __syncthreads();

    const bool valid = t < nrows;
    const int row = r0 + t;
    float prob[NF];
#pragma unroll
    for (int q = 0; q < NF; ++q) prob[q] = -1.0f;
    if (valid) {
        const float* x = s_conf + t * C;
        float4 cx;
        if (FROM_SCORES) {
#pragma unroll
            for (int q = 0; q < NF; ++q) prob[q] = x[q];
            cx = reinterpret_cast<const float4*>(loc)[(size_t)b * P + row];
        } else {
            // softmax: exp(x - max) * (1 / sum)  (Losses.py:25)
            float e[C];
            float m = x[0];
#pragma unroll
            for (int q = 1; q < C; ++q) m = fmaxf(m, x[q]);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < C; ++q) { e[q] = __expf(__fsub_rn(x[q], m)); s = __fadd_rn(s, e[q]); }   // ex2.approx: rel. error ~2e-7
            const float inv = __fdiv_rn(1.0f, s);
#pragma unroll
            for (int q = 0; q < NF; ++q) prob[q] = __fmul_rn(e[q], inv);
            cx = decode_box(reinterpret_cast<const float4*>(loc)[(size_t)b * P + row], pri_cxcywh[row]);   // Losses.py:23
        }
        boxes_out[(size_t)b * P + row] = cxcywh_to_xyxy(cx);                                               // Losses.py:41,71
    }

    // candidates: prob >= min_score (Losses.py:32).  Lane c owns the counter of class c: one atomic instruction
    // reserves the warp's slots in all 20 lists at once.
    unsigned mine_ball = 0u;
    unsigned balls[NF];
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        balls[q] = __ballot_sync(FULL, valid && prob[q] >= min_score);
        if (lane == q) mine_ball = balls[q];
    }
    static_assert(NF <= 32, "one lane per foreground class");
    unsigned my_base = 0u;
    if (lane < NF && mine_ball) my_base = atomicAdd(&cand_cnt[(size_t)b * NF + lane], (unsigned)__popc(mine_ball));
#pragma unroll
    for (int q = 0; q < NF; ++q) {
        const unsigned bq = __shfl_sync(FULL, my_base, q);
        if ((balls[q] >> lane) & 1u) {
            const unsigned slot = bq + (unsigned)__popc(balls[q] & ((1u << lane) - 1u));
            if (slot < (unsigned)cap) {
                cand[((size_t)b * NF + q) * capp + slot] =
                    ((unsigned long long)__float_as_uint(prob[q]) << 32) | (unsigned long long)(0xffffffffu - (unsigned)row);
            } else {
                atomicOr(&overflow[b], 1u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// `a` suppresses `b`  <=>  inter / union >= thr, evaluated exactly as the reference does (Util.py:262-265,
// 294-301, Losses.py:51) but without the division when the answer is not within 2^-20 of the threshold.
__device__ __forceinline__ bool iou_ge(const float4 a, const float aa, const float4 b, const float ab,
                                       const float thr, const float thr_lo, const float thr_hi)
{
    const float dx = fmaxf(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 0.0f);
    const float dy = fmaxf(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 0.0f);
    const float inter = __fmul_rn(dx, dy);
    const float uni = __fsub_rn(__fadd_rn(aa, ab), inter);
    if (uni > 0.0f && thr > 0.0f) {
        if (inter < __fmul_rn(uni, thr_lo)) return false;
        if (inter > __fmul_rn(uni, thr_hi)) return true;
    }
    return __fdiv_rn(inter, uni) >= thr;
}

__device__ __forceinline__ void bitonic_sort_desc(unsigned long long* keys, int n_pad)
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

constexpr int SORT_SMEM = 4096;    // keys of one slice sorted in shared memory; larger tie groups sort in place in L2
constexpr int NMS_BINS = 2048;     // linear probability bins used to cut the candidate list into score slices

// One block of up to 64 sorted candidates against the kept list, then inside the block (Losses.py:44-55).
// Returns the new kept count.  s_kkey / s_kbox / s_karea hold the kept set (capacity top_k + 64).
struct NmsShared {
    float4 cbox[64];
    float carea[64];
    unsigned long long ckey[64];
    unsigned long long mask[64];
    unsigned int supp[2];
    unsigned long long alive;
};

__device__ __forceinline__ int nms_block(NmsShared& s, const unsigned long long* keys, int base, int m, const float4* bx,
                                         float4* s_kbox, float* s_karea, unsigned long long* s_kkey, int K, int kcap,
                                         float iou_thr, float thr_lo, float thr_hi)
{
    const int t = threadIdx.x;
    if (t < 64) {
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        unsigned long long key = 0ull;
        if (t < m) {
            key = keys[base + t];
            box = bx[0xffffffffu - (unsigned)(key & 0xffffffffull)];
        }
        s.cbox[t] = box;
        s.carea[t] = box_area(box);
        s.ckey[t] = key;
        s.mask[t] = 0ull;
    }
    if (t < 2) s.supp[t] = 0u;
    __syncthreads();
    {
        // (a) against the boxes kept so far: candidate = t & 63, the kept list is split four ways
        const int cnd = t & 63, part = t >> 6;
        if (cnd < m) {
            const float4 cb = s.cbox[cnd];
            const float ca = s.carea[cnd];
            for (int k = part; k < K; k += 4) {
                if (iou_ge(s_kbox[k], s_karea[k], cb, ca, iou_thr, thr_lo, thr_hi)) {
                    atomicOr(&s.supp[cnd >> 5], 1u << (cnd & 31));
                    break;
                }
            }
        }
        // (b) inside the block: row = t & 63 tests the 16 columns [16*part, 16*part+16) that come after it
        const int rowi = t & 63;
        if (rowi < m) {
            const float4 rb = s.cbox[rowi];
            const float ra = s.carea[rowi];
            unsigned long long bits = 0ull;
#pragma unroll 4
            for (int q = 0; q < 16; ++q) {
                const int col = part * 16 + q;
                if (col > rowi && col < m && iou_ge(rb, ra, s.cbox[col], s.carea[col], iou_thr, thr_lo, thr_hi))
                    bits |= 1ull << col;
            }
            if (bits) atomicOr(&s.mask[rowi], bits);
        }
    }
    __syncthreads();
    if (t < 32) {
        // (c) serial resolve: a box that is still alive suppresses the later boxes it overlaps.  The 64 mask rows are
        // pulled into registers first (two per lane, independent loads) so the dependent chain is pure ALU + shuffles.
        const unsigned long long m_lo = s.mask[t], m_hi = s.mask[t + 32];
        unsigned long long alive = (m == 64 ? ~0ull : ((1ull << m) - 1ull)) &
                                   ~((unsigned long long)s.supp[0] | ((unsigned long long)s.supp[1] << 32));
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned long long mk = __shfl_sync(FULL, m_lo, i);
            alive &= ((alive >> i) & 1ull) ? ~mk : ~0ull;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const unsigned long long mk = __shfl_sync(FULL, m_hi, i);
            alive &= ((alive >> (i + 32)) & 1ull) ? ~mk : ~0ull;
        }
        if (t == 0) s.alive = alive;
    }
    __syncthreads();
    const unsigned long long alive = s.alive;
    if (t < m && ((alive >> t) & 1ull)) {
        const int pos = K + __popcll(alive & ((1ull << t) - 1ull));
        if (pos < kcap) { s_kbox[pos] = s.cbox[t]; s_karea[pos] = s.carea[t]; s_kkey[pos] = s.ckey[t]; }
    }
    __syncthreads();
    return K + __popcll(alive);
}

__global__ void __launch_bounds__(DT)
detect_nms_kernel(const float4* __restrict__ boxes, unsigned long long* __restrict__ cand,
                  unsigned int* __restrict__ cand_cnt, unsigned int* __restrict__ kept_cnt,
                  int P, int NF, int cap, int capp, int top_k, float iou_thr)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: kept boxes float4[kcap] | kept keys u64[kcap] | kept areas float[kcap] | sort buffer u64[SORT_SMEM] | hist u32[NMS_BINS]
    const int kcap = top_k + 64;
    float4* s_kbox = reinterpret_cast<float4*>(smem_raw);
    unsigned long long* s_kkey = reinterpret_cast<unsigned long long*>(s_kbox + kcap);
    float* s_karea = reinterpret_cast<float*>(s_kkey + kcap);
    unsigned long long* s_sort = reinterpret_cast<unsigned long long*>(smem_raw + (((size_t)kcap * 28 + 15) & ~(size_t)15));
    unsigned int* s_hist = reinterpret_cast<unsigned int*>(s_sort + SORT_SMEM);
    __shared__ NmsShared s;
    __shared__ int s_lo, s_cnt;
    __shared__ unsigned int s_fill;
    __shared__ unsigned int s_wsum[DT / 32];

    const int c = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const size_t seg_id = (size_t)b * NF + c;
    unsigned long long* seg = cand + seg_id * capp;
    const int n = (int)min(cand_cnt[seg_id], (unsigned)cap);
    if (n == 0) {
        if (t == 0) { kept_cnt[seg_id] = 0u; cand_cnt[seg_id] = 0u; }
        return;
    }
    const float thr_lo = __fmul_rn(iou_thr, 1.0f - 9.5367431640625e-07f);   // thr * (1 - 2^-20)
    const float thr_hi = __fmul_rn(iou_thr, 1.0f + 9.5367431640625e-07f);
    const float4* bx = boxes + (size_t)b * P;
    int K = 0;
    bool full_sort = false;

    // The sweep stops once top_k boxes are kept, so usually only the top few hundred candidates matter.  The list
    // is therefore consumed in SLICES of descending score: a histogram over linear probability bins (monotone in
    // the key) finds bin ranges holding about 2*top_k candidates; each slice is compacted into shared memory,
    // bitonic-sorted (descending prob, ties -> lower prior, T5) and swept, continuing with the same kept set.
    const int slice_target = max(top_k + top_k / 2, 192);
    // bins in the LOG domain, anchored at probability 1: the float's exponent and top 8 mantissa bits give 256 bins
    // per octave, 8 octaves (2^-8 .. 1) in 2048 bins - a relative resolution of 0.4 % everywhere, so a slice lands
    // close to its target even where candidates crowd just above min_score; monotone in the key; smaller
    // probabilities share bin 0
    auto bin_of = [](unsigned long long key) {
        const int top = (int)(0x3f800000u >> 15);                     // bits of 1.0f
        const int b = NMS_BINS - 1 - (top - (int)((unsigned)(key >> 32) >> 15));
        return min(NMS_BINS - 1, max(0, b));
    };
    if (n <= slice_target * 2 && n <= SORT_SMEM) {
        // short list: one slice = everything
        const int n_pad = pow2ceil(n);
        for (int i = t; i < n_pad; i += DT) s_sort[i] = i < n ? seg[i] : 0ull;
        __syncthreads();
        bitonic_sort_desc(s_sort, n_pad);
        for (int base = 0; base < n && K < top_k; base += 64)
            K = nms_block(s, s_sort, base, min(64, n - base), bx, s_kbox, s_karea, s_kkey, K, kcap, iou_thr, thr_lo, thr_hi);
    } else {
        for (int i = t; i < NMS_BINS; i += DT) s_hist[i] = 0u;
        __syncthreads();
        for (int i = t; i < n; i += DT) atomicAdd(&s_hist[bin_of(seg[i])], 1u);
        __syncthreads();
        int hi = NMS_BINS;                       // bins >= hi are done
        int done = 0;
        while (K < top_k && done < n) {
            {
                // walk down from `hi` until the slice holds slice_target candidates (a bin is never split):
                // thread t owns 8 bins counted from the top, a block prefix sum finds where the target is crossed
                constexpr int per = NMS_BINS / DT;
                unsigned c8[per], mine = 0u;
#pragma unroll
                for (int q = 0; q < per; ++q) {
                    const int idx = NMS_BINS - 1 - (t * per + q);
                    c8[q] = idx < hi ? s_hist[idx] : 0u;
                    mine += c8[q];
                }
                unsigned inc = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const unsigned o = __shfl_up_sync(FULL, inc, d);
                    if ((t & 31) >= d) inc += o;
                }
                if ((t & 31) == 31) s_wsum[t >> 5] = inc;
                if (t == 0) { s_lo = 0; s_cnt = -1; s_fill = 0u; }
                __syncthreads();
                unsigned wbase = 0u, total = 0u;
#pragma unroll
                for (int w = 0; w < DT / 32; ++w) { const unsigned cw = s_wsum[w]; if (w < (t >> 5)) wbase += cw; total += cw; }
                unsigned above = wbase + inc - mine;
                if (above < (unsigned)slice_target && above + mine >= (unsigned)slice_target) {
#pragma unroll
                    for (int q = 0; q < per; ++q) {
                        above += c8[q];
                        if (above >= (unsigned)slice_target) { s_lo = NMS_BINS - 1 - (t * per + q); s_cnt = (int)above; break; }
                    }
                }
                __syncthreads();
                if (t == 0 && s_cnt < 0) s_cnt = (int)total;      // fewer than the target left: the rest is one slice (lo = 0)
            }
            __syncthreads();
            const int lo = s_lo, cnt = s_cnt;
            if (cnt > SORT_SMEM) { full_sort = true; break; }       // a crowd of (near-)equal scores: general path
            const int n_pad = pow2ceil(max(cnt, 1));
            for (int i = t; i < n_pad; i += DT) s_sort[i] = 0ull;
            __syncthreads();
            for (int i = t; i < n; i += DT) {
                const unsigned long long key = seg[i];
                const int bn = bin_of(key);
                if (bn >= lo && bn < hi) s_sort[atomicAdd(&s_fill, 1u)] = key;
            }
            __syncthreads();
            bitonic_sort_desc(s_sort, n_pad);
            for (int base = 0; base < cnt && K < top_k; base += 64)
                K = nms_block(s, s_sort, base, min(64, cnt - base), bx, s_kbox, s_karea, s_kkey, K, kcap, iou_thr, thr_lo, thr_hi);
            hi = lo;
            done += cnt;
        }
    }
    if (full_sort) {
        // general path: sort the whole list in place in global memory (L2) and sweep from the start
        const int n_pad = pow2ceil(n);
        for (int i = n + t; i < n_pad; i += DT) seg[i] = 0ull;
        __syncthreads();
        bitonic_sort_desc(seg, n_pad);
        K = 0;
        for (int base = 0; base < n && K < top_k; base += 64)
            K = nms_block(s, seg, base, min(64, n - base), bx, s_kbox, s_karea, s_kkey, K, kcap, iou_thr, thr_lo, thr_hi);
    }
    // kept keys, in descending score order, to the front of the segment
    for (int i = t; i < K; i += DT) seg[i] = s_kkey[i];
    if (t == 0) { kept_cnt[seg_id] = (unsigned)K; cand_cnt[seg_id] = 0u; }
}

// ------------------------------------------------------------------------------------------------
constexpr int TK_T = 1024;

__global__ void __launch_bounds__(TK_T)
detect_topk_kernel(const float4* __restrict__ boxes, const unsigned long long* __restrict__ cand,
                   const unsigned int* __restrict__ kept_cnt, unsigned int* __restrict__ overflow,
                   const float* __restrict__ img_wh, int P, int NF, int capp, int top_k,
                   float4* __restrict__ out_boxes, float* __restrict__ out_prob, int* __restrict__ out_cls,
                   int* __restrict__ out_prior, int* __restrict__ out_cnt)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_keys = reinterpret_cast<unsigned long long*>(smem_raw);    // [pow2ceil(NF*top_k)]
    __shared__ int s_pref[33];
    __shared__ int s_m[33];
    __shared__ unsigned int s_tlow;
    __shared__ int s_all_r;
    const int b = blockIdx.x, t = threadIdx.x;
    if (t < NF) s_m[t] = (int)kept_cnt[(size_t)b * NF + t];
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        for (int c = 0; c < NF; ++c) { s_pref[c] = acc; acc += s_m[c]; }
        s_pref[NF] = acc;
        s_tlow = 0u;
    }
    __syncthreads();
    const int total = s_pref[NF];
    float sx = 1.0f, sy = 1.0f;
    if (img_wh) { sx = img_wh[2 * b]; sy = img_wh[2 * b + 1]; }
    const float4* bx = boxes + (size_t)b * P;
    float4* ob = out_boxes + (size_t)b * top_k;
    float* op = out_prob + (size_t)b * top_k;
    int* oc = out_cls + (size_t)b * top_k;
    int* oi = out_prior ? out_prior + (size_t)b * top_k : nullptr;

    auto emit = [&](int slot, unsigned long long key, int c) {
        const unsigned prior = 0xffffffffu - (unsigned)(key & 0xffffffffull);
        const float4 v = bx[prior];
        ob[slot] = img_wh ? make_float4(__fmul_rn(v.x, sx), __fmul_rn(v.y, sy), __fmul_rn(v.z, sx), __fmul_rn(v.w, sy)) : v;   // Losses.py:89
        op[slot] = __uint_as_float((unsigned)(key >> 32));
        oc[slot] = c;
        if (oi) oi[slot] = (int)prior;
    };

    int nout;
    if (total <= top_k) {
        // class-major, each class in descending score order (Losses.py:71-73)
        nout = total;
        for (int c = 0; c < NF; ++c) {
            const int kc = s_pref[c + 1] - s_pref[c];
            const unsigned long long* seg = cand + ((size_t)b * NF + c) * capp;
            for (int r = t; r < kc; r += TK_T) emit(s_pref[c] + r, seg[r], c);
        }
    } else {
        // global top_k by descending prob, ties -> earlier class-major position (T7).  Only the first top_k kept
        // boxes of a class can qualify, and nothing scoring below the top_k-th score of any single class can either:
        // T_low = max over classes of that score prunes the merge to a few hundred keys.
        nout = top_k;
        // two lower bounds on the top_k-th score of the union: (a) the top_k-th score of any single class;
        // (b) with r = ceil(top_k / NF): if every class kept at least r boxes, the NF*r >= top_k boxes formed by the
        //     top r of each class all score >= the smallest r-th score, so the top_k-th score of the union does too.
        const int r = (top_k + NF - 1) / NF;
        if (t == 0) s_all_r = 1;
        __syncthreads();
        unsigned rth = 0xffffffffu;
        if (t < NF) {
            const int kc = s_pref[t + 1] - s_pref[t];
            const unsigned long long* seg = cand + ((size_t)b * NF + t) * capp;
            if (kc >= top_k) atomicMax(&s_tlow, (unsigned)(seg[top_k - 1] >> 32));
            if (kc >= r) rth = (unsigned)(seg[r - 1] >> 32); else s_all_r = 0;
        }
        if (t < 32) {
            rth = __reduce_min_sync(FULL, rth);
            __syncwarp();
            if (t == 0 && s_all_r) atomicMax(&s_tlow, rth);
        }
        __syncthreads();
        const unsigned tlow = s_tlow;
        if (t < NF) {
            // lists are sorted by descending prob: binary search for the first entry below T_low
            const unsigned long long* seg = cand + ((size_t)b * NF + t) * capp;
            int lo = 0, hi = min(s_pref[t + 1] - s_pref[t], top_k);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((unsigned)(seg[mid] >> 32) >= tlow) lo = mid + 1; else hi = mid;
            }
            s_m[t] = lo;
        }
        __syncthreads();
        int m_total = 0;
        for (int c = 0; c < NF; ++c) m_total += s_m[c];
        const int n_pad = pow2ceil(m_total);
        for (int i = t; i < n_pad; i += TK_T) s_keys[i] = 0ull;
        __syncthreads();
        int acc = 0;
        for (int c = 0; c < NF; ++c) {
            const int mc = s_m[c];
            const unsigned long long* seg = cand + ((size_t)b * NF + c) * capp;
            for (int r = t; r < mc; r += TK_T)
                s_keys[acc + r] = (seg[r] & 0xffffffff00000000ull) | (unsigned long long)(0xffffffffu - (unsigned)(c * top_k + r));
            acc += mc;
        }
        __syncthreads();
        bitonic_sort_desc(s_keys, n_pad);
        for (int s = t; s < top_k; s += TK_T) {
            const unsigned pos = 0xffffffffu - (unsigned)(s_keys[s] & 0xffffffffull);
            const int c = (int)(pos / (unsigned)top_k), r = (int)(pos % (unsigned)top_k);
            emit(s, cand[((size_t)b * NF + c) * capp + r], c);
        }
    }
    if (t == 0) {
        out_cnt[b] = overflow[b] ? -1 : nout;     // -1: a candidate list exceeded the caller's cap
        overflow[b] = 0u;
    }
}

template <bool FROM_SCORES>
static int run_detect(const float* loc, const float* conf, const float* pri_cxcywh, int B, int P, int C,
                      float min_score, float iou_thr, int top_k, const float* img_wh,
                      float* out_boxes, float* out_prob, int32_t* out_cls, int32_t* out_prior, int32_t* out_cnt,
                      void* ws, size_t ws_bytes, int n_cap, cudaStream_t st)
{
    if (B < 0 || P <= 0 || top_k <= 0) return SSDHEAD_E_BADARG;
    if (!loc || !conf || (!FROM_SCORES && !pri_cxcywh) || !out_boxes || !out_prob || !out_cls || !out_cnt || !ws) return SSDHEAD_E_BADARG;
    if (C != 21) return SSDHEAD_E_UNSUPPORTED;
    if (B == 0) return 0;
    if (B > 65535) return SSDHEAD_E_UNSUPPORTED;
    if (!aligned16(loc) || (!FROM_SCORES && !aligned16(pri_cxcywh)) || !aligned16(out_boxes) || !aligned16(ws)) return SSDHEAD_E_ALIGN;
    const int NF = C - 1;
    if ((size_t)pow2ceil(NF * top_k) * 8 > 200 * 1024) return SSDHEAD_E_UNSUPPORTED;     // top_k <= 1638 for 20 classes
    DetectWs w;
    const size_t need = detect_ws_layout(B, P, C, n_cap, &w, ws);
    if (ws_bytes < need) return SSDHEAD_E_WORKSPACE;
    const int cap = n_cap > 0 ? std::min(n_cap, P) : P;
    const int capp = pow2ceil(cap);

    dim3 g1((P + DT - 1) / DT, B);
    detect_score_kernel<21, FROM_SCORES><<<g1, DT, 0, st>>>(loc, conf, (const float4*)pri_cxcywh, P, min_score, cap, capp,
                                                           w.boxes, w.cand, w.cand_cnt, w.overflow);
    count_launch();
    SSD_LAUNCH_CHECK();

    const size_t smem_nms = (((size_t)(top_k + 64) * 28 + 15) & ~(size_t)15) + (size_t)SORT_SMEM * 8 + (size_t)NMS_BINS * 4;
    SSD_CHECK_CUDA(cudaFuncSetAttribute(detect_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_nms));
    detect_nms_kernel<<<dim3(NF, B), DT, smem_nms, st>>>(w.boxes, w.cand, w.cand_cnt, w.kept_cnt, P, NF, cap, capp, top_k, iou_thr);
    count_launch();
    SSD_LAUNCH_CHECK();

    const size_t smem_topk = (size_t)pow2ceil(NF * top_k) * 8;
    SSD_CHECK_CUDA(cudaFuncSetAttribute(detect_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_topk));
    detect_topk_kernel<<<B, TK_T, smem_topk, st>>>(w.boxes, w.cand, w.kept_cnt, w.overflow, img_wh, P, NF, capp, top_k,
                                                 (float4*)out_boxes, out_prob, out_cls, out_prior, out_cnt);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

}  // namespace ssdhead

using namespace ssdhead;

extern "C" {

int ssdhead_detect(const float* loc, const float* conf, const float* pri_cxcywh, int B, int P, int C,
                   float min_score, float iou_thr, int top_k, const float* img_wh, int max_candidates,
                   float* out_boxes, float* out_prob, int32_t* out_cls, int32_t* out_prior, int32_t* out_cnt,
                   void* ws, size_t ws_bytes, void* stream)
{
    return run_detect<false>(loc, conf, pri_cxcywh, B, P, C, min_score, iou_thr, top_k, img_wh,
                             out_boxes, out_prob, out_cls, out_prior, out_cnt, ws, ws_bytes, max_candidates, (cudaStream_t)stream);
}

int ssdhead_detect_from_scores(const float* boxes_cxcywh, const float* probs, int B, int P, int C,
                               float min_score, float iou_thr, int top_k, const float* img_wh, int max_candidates,
                               float* out_boxes, float* out_prob, int32_t* out_cls, int32_t* out_prior, int32_t* out_cnt,
                               void* ws, size_t ws_bytes, void* stream)
{
    return run_detect<true>(boxes_cxcywh, probs, nullptr, B, P, C, min_score, iou_thr, top_k, img_wh,
                            out_boxes, out_prob, out_cls, out_prior, out_cnt, ws, ws_bytes, max_candidates, (cudaStream_t)stream);
}

}  // extern "C"
