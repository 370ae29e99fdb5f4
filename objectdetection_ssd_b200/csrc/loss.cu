// Multibox loss: cross entropy + L1 + hard-negative mining, forward and gradients.
// Reference: ssd / ssd1_, Losses.py:119-199.
//
// The gradient of this loss is SPARSE: only positives and mined negatives (about 4*npos of the
// 8732 rows of an image) have a non-zero conf gradient, and only positives touch loc at all.
// The path is therefore split into one streaming kernel and one small per-image kernel:
//
//   ce_stream_kernel   persistent CTAs, one producer warp + 8 consumer warps.  The producer moves
//                      256-row tiles of conf (21 504 B) into a 4-stage shared-memory ring with 1-D
//                      TMA bulk copies (mbarrier full/empty pipeline) and, from a zeroed tile in
//                      shared memory, bulk-stores the zero background of grad_conf / grad_loc.
//                      Consumers do thread-per-row log-softmax out of shared memory (row stride 21
//                      words: bank-conflict free) and write one CE value per prior.  conf is read
//                      exactly once; nothing is re-read; the tile grid is flat over B*P rows so
//                      every bulk copy is 16-byte aligned whatever P is.
//   mine_kernel        one CTA per image: CE row + class bytes -> keys in shared memory, exact
//                      radix select (11/11/10 bits) of the k = 3*npos-th largest background CE,
//                      ties to the lower prior index (T4; positives rank with value 0,
//                      Losses.py:190), then thread-per-row over the ~4*npos selected rows only:
//                      re-read that conf row, write (softmax - onehot)/N into grad_conf, and for
//                      positives the L1 term and sign/(4N) into grad_loc.  Loss partials are reduced
//                      in fp64 in a fixed order by the last CTA -> run-to-run deterministic.
//
// HBM traffic per image: conf in once (733 KB) + CE out/in (2 x 35 KB, L2-resident between the two
// kernels) + dense gradients out once (873 KB) + ~4*npos sparse rows: the algorithmic minimum.
#include <algorithm>
#include "common.cuh"

namespace ssdhead {

// ------------------------------------------------------------------------------------------------
// per-row cross entropy, -(x_c - max - log(sum exp(x - max))): the order ATen's log_softmax uses
// ------------------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ float row_cross_entropy(const float* __restrict__ row, int c)
{
    float x[C];
#pragma unroll
    for (int q = 0; q < C; ++q) x[q] = row[q];
    float m = x[0];
#pragma unroll
    for (int q = 1; q < C; ++q) m = fmaxf(m, x[q]);
    float s = 0.0f, xc = 0.0f;
#pragma unroll
    for (int q = 0; q < C; ++q) {
        const float d = __fsub_rn(x[q], m);
        s = __fadd_rn(s, expf(d));
        if (q == c) xc = d;
    }
    const float ce = __fsub_rn(logf(s), xc);
    return __fadd_rn(ce, 0.0f);                  // -0.0 -> +0.0 so the bit pattern orders like the value
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

constexpr int CE_ROWS = 256;                      // rows per tile = consumer threads
constexpr int CE_STAGES = 4;
constexpr int CE_THREADS = CE_ROWS + 32;          // + one producer warp

template <int C, bool ZERO_FILL>
__global__ void __launch_bounds__(CE_THREADS, 2)
ce_stream_kernel(const float* __restrict__ conf, float* __restrict__ ce_out,
                 float* __restrict__ grad_conf, float* __restrict__ grad_loc, long long total_rows, int use_tma)
{
    constexpr uint32_t TILE_BYTES = CE_ROWS * C * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_full[CE_STAGES];
    __shared__ __align__(8) uint64_t s_empty[CE_STAGES];
    float* zero_tile = reinterpret_cast<float*>(smem_raw + (size_t)CE_STAGES * TILE_BYTES);

    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const long long full_tiles = use_tma ? total_rows / CE_ROWS : 0;

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < CE_STAGES; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], CE_ROWS / 32); }
        mbar_fence_init();
    }
    if (ZERO_FILL) {
        for (int i = t; i < (int)(TILE_BYTES / 16); i += CE_THREADS)
            reinterpret_cast<float4*>(zero_tile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        fence_proxy_async_smem();
    }
    __syncthreads();

    if (warp == CE_ROWS / 32) {
        // ---------------- producer warp: one elected lane drives the TMA engine ----------------
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (long long tile = blockIdx.x; tile < full_tiles; tile += gridDim.x) {
                mbar_wait(&s_empty[s], ph ^ 1u);
                mbar_expect_tx(&s_full[s], TILE_BYTES);
                bulk_g2s(smem_raw + (size_t)s * TILE_BYTES, conf + (size_t)tile * CE_ROWS * C, TILE_BYTES, &s_full[s]);
                if (ZERO_FILL) {
                    bulk_s2g(grad_conf + (size_t)tile * CE_ROWS * C, zero_tile, TILE_BYTES);
                    bulk_s2g(grad_loc + (size_t)tile * CE_ROWS * 4, zero_tile, CE_ROWS * 16);
                    bulk_commit();
                }
                if (++s == CE_STAGES) { s = 0; ph ^= 1u; }
            }
            if (ZERO_FILL) bulk_wait_read_all();
        }
    } else {
        // ---------------- consumer warps: thread per row ----------------
        int s = 0;
        uint32_t ph = 0;
        for (long long tile = blockIdx.x; tile < full_tiles; tile += gridDim.x) {
            const long long row = tile * CE_ROWS + t;
            mbar_wait(&s_full[s], ph);
            const float ce = row_cross_entropy<C>(reinterpret_cast<const float*>(smem_raw + (size_t)s * TILE_BYTES) + t * C, C - 1);
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
            ce_out[row] = ce;
            if (++s == CE_STAGES) { s = 0; ph ^= 1u; }
        }
        // rows past the last full tile (or every row when the pointers are not 16-byte aligned): plain loads
        const long long rest0 = full_tiles * CE_ROWS;
        for (long long row = rest0 + (long long)blockIdx.x * CE_ROWS + t; row < total_rows; row += (long long)gridDim.x * CE_ROWS) {
            ce_out[row] = row_cross_entropy<C>(conf + (size_t)row * C, C - 1);
            if (ZERO_FILL) {
#pragma unroll
                for (int q = 0; q < C; ++q) grad_conf[(size_t)row * C + q] = 0.0f;
#pragma unroll
                for (int q = 0; q < 4; ++q) grad_loc[(size_t)row * 4 + q] = 0.0f;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mine_kernel
// ------------------------------------------------------------------------------------------------
constexpr int MN_T = 512;
constexpr int MN_W = MN_T / 32;
constexpr int MN_BINS = 2048;

struct MineParams {
    const float* loc;
    const float* conf;
    const float* ce;             // CE of every row against the BACKGROUND class (ce_stream_kernel)
    float* ce_tap;               // nullable: receives the true CE of positive rows (debug tap)
    const uint8_t* cls_u8;
    const float4* gt_xyxy;
    const float* gt_cls;
    const int* gt_off;
    const float4* pri_xyxy;
    const float4* pri_cxcywh;
    const int* best_prior;
    const int* npos;
    const int* npos_norm;
    int B, P, neg_ratio, bg_class;
    float pos_iou;
    double* sums;
    float* losses;
    float* grad_loc;
    float* grad_conf;
    uint32_t* mined_mask;
    double* partials;            // [B][2]
    unsigned int* done_counter;  // self-resetting
};

// exclusive prefix sum over the MN_T threads of the CTA; *total receives the block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* s_warp /*[MN_W+1]*/, uint32_t* total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += o;
    }
    __syncthreads();                              // protect s_warp against the previous use
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t base = 0u, tot = 0u;
#pragma unroll
    for (int w = 0; w < MN_W; ++w) { const uint32_t c = s_warp[w]; if (w < warp) base += c; tot += c; }
    *total = tot;
    return base + inc - v;
}

constexpr int MN_GC = 64;        // gt boxes staged in shared memory
constexpr int MN_CAND = 512;     // boundary-bin candidates ranked directly (one per thread)

template <int C, bool GRADS>
__global__ void __launch_bounds__(MN_T)
mine_kernel(const MineParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* s_key = reinterpret_cast<uint32_t*>(smem_raw);            // [P]  CE bit pattern, 0 for positives; bit 31 = selected
    uint32_t* s_hist = s_key + p.P;                                      // [MN_BINS]
    uint16_t* s_list = reinterpret_cast<uint16_t*>(s_hist + MN_BINS);    // [P]  rows that carry a gradient (P < 65536)
    uint8_t* s_cls = reinterpret_cast<uint8_t*>(s_list + ((p.P + 7) & ~7)); // [P]  class bytes (16-byte aligned)
    __shared__ uint32_t s_warp[MN_W + 1];
    __shared__ uint32_t s_sel[3];
    __shared__ uint32_t s_nsel, s_ncand, s_max;
    __shared__ uint32_t s_ckey[MN_CAND], s_cidx[MN_CAND];
    __shared__ float4 s_gbox[MN_GC];
    __shared__ float s_garea[MN_GC];
    __shared__ int s_gbp[MN_GC];
    __shared__ double s_redd[2][MN_W];
    __shared__ int s_is_last;

    const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int P = p.P;
    const size_t row0 = (size_t)b * P;
    const int off0 = p.gt_off[b];
    const int G = p.gt_off[b + 1] - off0;
    double acc_l1 = 0.0, acc_ce = 0.0;

    // ---- 1. keys: CE bits for negatives, 0 for positives (Losses.py:188-190); class bytes; gts ----
    if (t == 0) { s_nsel = 0u; s_ncand = 0u; s_max = 0u; }
    if (t < min(G, MN_GC)) {
        const float4 bx = p.gt_xyxy[off0 + t];
        s_gbox[t] = bx;
        s_garea[t] = box_area(bx);
        s_gbp[t] = p.best_prior[off0 + t];
    }
    __syncthreads();
    uint32_t kmax = 0u;
    const bool vec_ok = ((P & 3) == 0) && (((row0 * 4) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p.ce) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(p.cls_u8) & 3) == 0);
    if (vec_ok) {
        // 16-byte CE loads + 4-byte class loads, several in flight per thread
        const float4* ce4 = reinterpret_cast<const float4*>(p.ce + row0);
        const uchar4* cl4 = reinterpret_cast<const uchar4*>(p.cls_u8 + row0);
        const int nvec = P >> 2;
#pragma unroll 4
        for (int v = t; v < nvec; v += MN_T) {
            const float4 e = ce4[v];
            const uchar4 c = cl4[v];
            const float ev[4] = {e.x, e.y, e.z, e.w};
            const uint8_t cv[4] = {c.x, c.y, c.z, c.w};
            uint32_t kv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const bool pos = (int)cv[q] != p.bg_class;
                kv[q] = pos ? 0u : (__float_as_uint(ev[q]) & 0x7fffffffu);
                kmax = max(kmax, kv[q]);
            }
            reinterpret_cast<uint4*>(s_key)[v] = make_uint4(kv[0], kv[1], kv[2], kv[3]);
            reinterpret_cast<uchar4*>(s_cls)[v] = c;
        }
    } else {
        for (int j = t; j < P; j += MN_T) {
            const float ce = p.ce[row0 + j];
            const uint8_t c = p.cls_u8[row0 + j];
            const bool pos = (int)c != p.bg_class;
            const uint32_t key = pos ? 0u : (__float_as_uint(ce) & 0x7fffffffu);
            s_key[j] = key;
            s_cls[j] = c;
            kmax = max(kmax, key);
        }
    }
    kmax = __reduce_max_sync(FULL, kmax);
    if (lane == 0) atomicMax(&s_max, kmax);
    __syncthreads();
    kmax = s_max;

    // ---- 2. the k largest keys, ties to the lower prior index (T4): mark them with bit 31 ----
    const long long kk = (long long)p.neg_ratio * (long long)p.npos[b];
    const uint32_t k = (uint32_t)min((long long)P, max(0ll, kk));
    bool listed = false;                         // the row list was already built by the fast path
    if (k >= (uint32_t)P) {
        for (int j = t; j < P; j += MN_T) s_key[j] |= 0x80000000u;
    } else if (k > 0u) {
        bool done = false;
        const float vmax = __uint_as_float(kmax);
        if (kmax > 0u && kmax < 0x7f800000u) {
            // fast path: one histogram over 2048 LINEAR bins of [0, max] (monotone in the key), then the few
            // keys of the boundary bin are ranked against each other directly.
            const float scale = __fdiv_rn(2047.0f, vmax);
            for (int i = t; i < MN_BINS; i += MN_T) s_hist[i] = 0u;
            __syncthreads();
            for (int j = t; j < P; j += MN_T)
                atomicAdd(&s_hist[(int)__fmul_rn(__uint_as_float(s_key[j]), scale)], 1u);
            __syncthreads();
            constexpr int per = MN_BINS / MN_T;
            uint32_t c4[per], mine = 0u;
#pragma unroll
            for (int q = 0; q < per; ++q) { c4[q] = s_hist[MN_BINS - 1 - (t * per + q)]; mine += c4[q]; }
            uint32_t tot;
            uint32_t above = block_exclusive_scan(mine, s_warp, &tot);
            if (above < k && above + mine >= k) {
#pragma unroll
                for (int q = 0; q < per; ++q) {
                    if (above + c4[q] >= k) { s_sel[0] = (uint32_t)(MN_BINS - 1 - (t * per + q)); s_sel[1] = above; s_sel[2] = c4[q]; break; }
                    above += c4[q];
                }
            }
            __syncthreads();
            const int bq = (int)s_sel[0];
            const uint32_t need = k - s_sel[1], cnt = s_sel[2];
            if (cnt <= (uint32_t)MN_CAND) {
                done = true;
                listed = true;
                // one pass: rows above the boundary bin are mined, rows in it become candidates, positives are listed
                for (int j = t; j < P; j += MN_T) {
                    const uint32_t key = s_key[j];
                    const bool pos = (int)s_cls[j] != p.bg_class;
                    const int bin = (int)__fmul_rn(__uint_as_float(key), scale);
                    if (pos) {
                        s_list[atomicAdd(&s_nsel, 1u)] = (uint16_t)j;
                    } else if (bin > bq) {
                        acc_ce += (double)__uint_as_float(key);
                        if (p.mined_mask) atomicOr(&p.mined_mask[(size_t)b * ((P + 31) / 32) + (j >> 5)], 1u << (j & 31));
                        if (GRADS) s_list[atomicAdd(&s_nsel, 1u)] = (uint16_t)j;
                    } else if (bin == bq) {
                        const uint32_t s = atomicAdd(&s_ncand, 1u);
                        s_ckey[s] = key;
                        s_cidx[s] = (uint32_t)j;
                    }
                }
                __syncthreads();
                // positives in the boundary bin (key 0, only when bq == 0) were listed above, not made candidates;
                // they still occupy ranks: `need` counts them, so rank them through the candidate count of bin 0
                const uint32_t ncand = s_ncand;
                if ((uint32_t)t < ncand) {
                    const uint32_t mk = s_ckey[t], mi = s_cidx[t];
                    uint32_t rank = 0u;
                    for (uint32_t q = 0; q < ncand; ++q) {
                        const uint32_t ok = s_ckey[q], oi = s_cidx[q];
                        rank += (ok > mk || (ok == mk && oi < mi)) ? 1u : 0u;
                    }
                    if (bq == 0) {
                        // positives (value 0) tie with zero-CE negatives: ties go to the lower prior index
                        if (mk == 0u) for (int j = 0; j < (int)mi; ++j) rank += ((int)s_cls[j] != p.bg_class) ? 1u : 0u;
                    }
                    if (rank < need) {
                        acc_ce += (double)__uint_as_float(mk);
                        if (p.mined_mask) atomicOr(&p.mined_mask[(size_t)b * ((P + 31) / 32) + (mi >> 5)], 1u << (mi & 31));
                        if (GRADS) s_list[atomicAdd(&s_nsel, 1u)] = (uint16_t)mi;
                    }
                }
            }
        }
        if (!done) {
            // general path (many equal keys, or non-finite CE): exact radix select, digits of 11 / 11 / 10 bits
            uint32_t prefix = 0u, mask = 0u, need = k;
#pragma unroll 1
            for (int pass = 0; pass < 3; ++pass) {
                const int shift = pass == 0 ? 21 : (pass == 1 ? 10 : 0);
                const int bins = pass == 2 ? 1024 : 2048;
                __syncthreads();
                for (int i = t; i < bins; i += MN_T) s_hist[i] = 0u;
                __syncthreads();
                for (int j = t; j < P; j += MN_T) {
                    const uint32_t key = s_key[j];
                    if ((key & mask) == prefix) atomicAdd(&s_hist[(key >> shift) & (uint32_t)(bins - 1)], 1u);
                }
                __syncthreads();
                const int per = bins / MN_T;          // 4 or 2
                uint32_t c4[4] = {0u, 0u, 0u, 0u}, mine = 0u;
#pragma unroll
                for (int q = 0; q < 4; ++q) if (q < per) { c4[q] = s_hist[bins - 1 - (t * per + q)]; mine += c4[q]; }
                uint32_t tot;
                uint32_t above = block_exclusive_scan(mine, s_warp, &tot);
                if (above < need && above + mine >= need) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (q < per) {
                        if (above + c4[q] >= need) { s_sel[0] = (uint32_t)(bins - 1 - (t * per + q)); s_sel[1] = above; s_sel[2] = c4[q]; break; }
                        above += c4[q];
                    }
                }
                __syncthreads();
                prefix |= s_sel[0] << shift;
                mask |= (uint32_t)(bins - 1) << shift;
                need -= s_sel[1];
            }
            const uint32_t T = prefix;               // take every key > T and the first `need` keys == T in prior order
            int per = (P + MN_T - 1) / MN_T;
            per |= 1;                                // odd stride: fewer shared-memory bank conflicts
            const int j0 = min(P, t * per), j1 = min(P, j0 + per);
            uint32_t ties = 0u;
            for (int j = j0; j < j1; ++j) ties += (s_key[j] == T) ? 1u : 0u;
            uint32_t tot_ties;
            uint32_t rank = block_exclusive_scan(ties, s_warp, &tot_ties);
            for (int j = j0; j < j1; ++j) {
                const uint32_t key = s_key[j];
                if (key > T) s_key[j] = key | 0x80000000u;
                else if (key == T) { if (rank < need) s_key[j] = key | 0x80000000u; ++rank; }
            }
        }
    }
    __syncthreads();

    // ---- 3. mined CE sum, list of rows that carry a gradient ----
    if (!listed) {
        for (int j = t; j < P; j += MN_T) {
            const uint32_t key = s_key[j];
            const bool pos = (int)s_cls[j] != p.bg_class;
            const bool mined = (key >> 31) && !pos;
            if (mined) {
                acc_ce += (double)__uint_as_float(key & 0x7fffffffu);
                if (p.mined_mask) atomicOr(&p.mined_mask[(size_t)b * ((P + 31) / 32) + (j >> 5)], 1u << (j & 31));
            }
            if (pos || (GRADS && mined)) s_list[atomicAdd(&s_nsel, 1u)] = (uint16_t)j;
        }
    }
    __syncthreads();

    // ---- 4. the selected rows only: conf gradient, and for positives the L1 term + loc gradient ----
    const uint32_t nsel = s_nsel;
    const float nrm = (float)(*p.npos_norm);
    const float gs_conf = __fdiv_rn(1.0f, nrm);
    const float gs_loc = __fdiv_rn(1.0f, __fmul_rn(4.0f, nrm));
    for (uint32_t idx = t; idx < nsel; idx += MN_T) {
        const int j = (int)s_list[idx];
        const int c = (int)s_cls[j];
        const bool pos = c != p.bg_class;
        float x[C];
        float4 pb, pc, l;
        if (GRADS || pos) {
            const float* row = p.conf + (row0 + j) * C;
#pragma unroll
            for (int q = 0; q < C; ++q) x[q] = __ldg(row + q);
        }
        if (pos) {                                   // issue every independent load before the first use
            pb = p.pri_xyxy[j];
            pc = p.pri_cxcywh[j];
            l = reinterpret_cast<const float4*>(p.loc)[row0 + j];
        }
        if (pos) {
            // the streaming kernel scored every row against the background class; a positive row gets its
            // true-class CE here, from the row it re-reads anyway (same code -> same bits as a direct evaluation)
            const float ce = row_cross_entropy<C>(x, c);
            acc_ce += (double)ce;
            if (p.ce_tap) p.ce_tap[row0 + j] = ce;
        }
        if (GRADS) {
            float m = x[0];
#pragma unroll
            for (int q = 1; q < C; ++q) m = fmaxf(m, x[q]);
            float s = 0.0f;
#pragma unroll
            for (int q = 0; q < C; ++q) { x[q] = expf(__fsub_rn(x[q], m)); s = __fadd_rn(s, x[q]); }
            const float inv = __fdiv_rn(1.0f, s);
            float* grow = p.grad_conf + (row0 + j) * C;
#pragma unroll
            for (int q = 0; q < C; ++q)
                grow[q] = __fmul_rn(__fsub_rn(__fmul_rn(x[q], inv), q == c ? 1.0f : 0.0f), gs_conf);
        }
        if (pos) {
            // which gt: the forced one (T3: highest index wins) or the natural argmax (T1)
            const float pa = box_area(pb);
            int obj = -1, ng = 0;
            float nb = 0.0f;
            for (int g = 0; g < G; ++g) {
                float4 gb; float ga; int bp;
                if (g < MN_GC) { gb = s_gbox[g]; ga = s_garea[g]; bp = s_gbp[g]; }
                else { gb = p.gt_xyxy[off0 + g]; ga = box_area(gb); bp = p.best_prior[off0 + g]; }
                if (bp == j) obj = g;
                const float v = iou_sparse(gb, ga, pb, pa);
                if (v > nb) { nb = v; ng = g; }
            }
            if (obj < 0) obj = ng;
            const float4 gbox = obj < MN_GC ? s_gbox[obj] : p.gt_xyxy[off0 + obj];
            const float4 tgt = encode_box(xyxy_to_cxcywh(gbox), pc);
            const float dx = __fsub_rn(l.x, tgt.x), dy = __fsub_rn(l.y, tgt.y);
            const float dz = __fsub_rn(l.z, tgt.z), dw = __fsub_rn(l.w, tgt.w);
            acc_l1 += (double)fabsf(dx) + (double)fabsf(dy) + (double)fabsf(dz) + (double)fabsf(dw);
            if (GRADS) {
                float4 gl;
                gl.x = dx > 0.f ? gs_loc : (dx < 0.f ? -gs_loc : 0.f);
                gl.y = dy > 0.f ? gs_loc : (dy < 0.f ? -gs_loc : 0.f);
                gl.z = dz > 0.f ? gs_loc : (dz < 0.f ? -gs_loc : 0.f);
                gl.w = dw > 0.f ? gs_loc : (dw < 0.f ? -gs_loc : 0.f);
                reinterpret_cast<float4*>(p.grad_loc)[row0 + j] = gl;
            }
        }
    }

    // ---- 5. loss sums: per-image partial -> the last CTA reduces all partials in a fixed order ----
    acc_l1 = warp_sum(acc_l1);
    acc_ce = warp_sum(acc_ce);
    if (lane == 0) { s_redd[0][warp] = acc_l1; s_redd[1][warp] = acc_ce; }
    __syncthreads();
    if (t == 0) {
        double a = 0.0, c = 0.0;
        for (int w = 0; w < MN_W; ++w) { a += s_redd[0][w]; c += s_redd[1][w]; }
        p.partials[2 * b] = a;
        p.partials[2 * b + 1] = c;
        __threadfence();
        const unsigned done = atomicAdd(p.done_counter, 1u);
        s_is_last = (done == gridDim.x - 1u) ? 1 : 0;
    }
    __syncthreads();
    if (s_is_last) {
        __threadfence();
        double a = 0.0, c = 0.0;
        for (int s = t; s < (int)gridDim.x; s += MN_T) {
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
            for (int w = 0; w < MN_W; ++w) { a += s_redd[0][w]; c += s_redd[1][w]; }
            p.sums[0] = a;
            p.sums[1] = c;
            const double N = (double)(*p.npos_norm);
            p.losses[0] = (float)(a / (4.0 * N));
            p.losses[1] = (float)(c / N);
            *p.done_counter = 0u;
        }
    }
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

static size_t mine_smem_bytes(int P) { return (size_t)P * 4 + MN_BINS * 4 + (size_t)((P + 7) & ~7) * 2 + round_up((size_t)P, 16); }

// workspace: [0,16) done counter | partials double[2B] | CE float[B*P] (when the caller passes no ce buffer)
size_t loss_workspace_bytes(int B, int P, int C)
{
    if (mine_smem_bytes(P) > 220 * 1024 || P >= 65536) return 0;   // P <= ~31 000 priors
    return 16 + round_up((size_t)B * 2 * sizeof(double), 16) + round_up((size_t)B * P * sizeof(float), 16);
}

static int g_num_sms = 0;
static int num_sms()
{
    if (g_num_sms == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        g_num_sms = n;
    }
    return g_num_sms;
}

template <int C, bool GRADS>
static int launch_ce_stream(const float* conf, float* ce, float* grad_conf, float* grad_loc, long long rows, cudaStream_t st)
{
    constexpr size_t tile = (size_t)CE_ROWS * C * 4;
    const size_t smem_ce = tile * CE_STAGES + (GRADS ? tile : 0);
    auto kce = ce_stream_kernel<C, GRADS>;
    SSD_CHECK_CUDA(cudaFuncSetAttribute(kce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_ce));
    const int use_tma = (aligned16(conf) && (!GRADS || (aligned16(grad_conf) && aligned16(grad_loc)))) ? 1 : 0;
    const long long tiles = (rows + CE_ROWS - 1) / CE_ROWS;
    const int grid_ce = (int)std::min<long long>(tiles, 2LL * num_sms());
    kce<<<grid_ce, CE_THREADS, smem_ce, st>>>(conf, ce, grad_conf, grad_loc, rows, use_tma);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

template <int C, bool GRADS>
static int launch_mine(const MineParams& prm, cudaStream_t st)
{
    auto kmn = mine_kernel<C, GRADS>;
    const size_t smem_mn = mine_smem_bytes(prm.P);
    SSD_CHECK_CUDA(cudaFuncSetAttribute(kmn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_mn));
    kmn<<<prm.B, MN_T, smem_mn, st>>>(prm);
    count_launch();
    SSD_LAUNCH_CHECK();
    return 0;
}

static float* ws_ce(void* ws, int B) { return (float*)((char*)ws + 16 + round_up((size_t)B * 2 * sizeof(double), 16)); }

}  // namespace ssdhead

using namespace ssdhead;

extern "C" {

int ssdhead_ce_stream(const float* conf, int B, int P, int C, float* ce, float* grad_loc, float* grad_conf,
                      void* ws, size_t ws_bytes, void* stream)
{
    if (B < 0 || P <= 0 || !conf || !ws) return SSDHEAD_E_BADARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return SSDHEAD_E_BADARG;
    if (C != 21) return SSDHEAD_E_UNSUPPORTED;          // VOC head of the reference (Losses.py:184 hard-codes 21)
    if (B == 0) return 0;
    if ((grad_loc && !aligned16(grad_loc)) || !aligned16(ws)) return SSDHEAD_E_ALIGN;
    const size_t need = loss_workspace_bytes(B, P, C);
    if (need == 0) return SSDHEAD_E_UNSUPPORTED;
    if (ws_bytes < need) return SSDHEAD_E_WORKSPACE;
    float* ce_buf = ce ? ce : ws_ce(ws, B);
    const long long rows = (long long)B * P;
    return grad_loc ? launch_ce_stream<21, true>(conf, ce_buf, grad_conf, grad_loc, rows, (cudaStream_t)stream)
                    : launch_ce_stream<21, false>(conf, ce_buf, nullptr, nullptr, rows, (cudaStream_t)stream);
}

int ssdhead_mine(const float* loc, const float* conf,
                 const float* gt_xyxy, const float* gt_cls, const int32_t* gt_off,
                 const float* pri_xyxy, const float* pri_cxcywh,
                 const int32_t* best_prior, const int32_t* npos, const int32_t* npos_norm, const uint8_t* cls_u8,
                 int B, int P, int C, int neg_ratio, float pos_iou,
                 double* sums, float* losses, float* grad_loc, float* grad_conf,
                 uint32_t* mined_mask, float* ce,
                 void* ws, size_t ws_bytes, void* stream)
{
    if (B < 0 || P <= 0 || neg_ratio < 0) return SSDHEAD_E_BADARG;
    if (!loc || !conf || !gt_off || !pri_xyxy || !pri_cxcywh || !npos || !npos_norm || !cls_u8 || !sums || !losses || !ws)
        return SSDHEAD_E_BADARG;
    if ((grad_loc == nullptr) != (grad_conf == nullptr)) return SSDHEAD_E_BADARG;
    if (C != 21) return SSDHEAD_E_UNSUPPORTED;
    if (B == 0) return 0;
    if (!aligned16(loc) || !aligned16(pri_xyxy) || !aligned16(pri_cxcywh) || (gt_xyxy && !aligned16(gt_xyxy)) ||
        (grad_loc && !aligned16(grad_loc)) || !aligned16(ws))
        return SSDHEAD_E_ALIGN;
    const size_t need = loss_workspace_bytes(B, P, C);
    if (need == 0) return SSDHEAD_E_UNSUPPORTED;
    if (ws_bytes < need) return SSDHEAD_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;

    MineParams prm;
    prm.loc = loc; prm.conf = conf; prm.cls_u8 = cls_u8;
    prm.gt_xyxy = (const float4*)gt_xyxy; prm.gt_cls = gt_cls; prm.gt_off = gt_off;
    prm.pri_xyxy = (const float4*)pri_xyxy; prm.pri_cxcywh = (const float4*)pri_cxcywh;
    prm.best_prior = best_prior; prm.npos = npos; prm.npos_norm = npos_norm;
    prm.B = B; prm.P = P; prm.neg_ratio = neg_ratio; prm.bg_class = C - 1; prm.pos_iou = pos_iou;
    prm.sums = sums; prm.losses = losses; prm.grad_loc = grad_loc; prm.grad_conf = grad_conf;
    prm.mined_mask = mined_mask;
    prm.done_counter = (unsigned int*)ws;
    prm.partials = (double*)((char*)ws + 16);
    prm.ce = ce ? ce : ws_ce(ws, B);
    prm.ce_tap = ce;
    if (mined_mask) SSD_CHECK_CUDA(cudaMemsetAsync(mined_mask, 0, (size_t)B * ((P + 31) / 32) * sizeof(uint32_t), st));
    return grad_loc ? launch_mine<21, true>(prm, st) : launch_mine<21, false>(prm, st);
}

int ssdhead_multibox_loss(const float* loc, const float* conf,
                          const float* gt_xyxy, const float* gt_cls, const int32_t* gt_off,
                          const float* pri_xyxy, const float* pri_cxcywh,
                          const int32_t* best_prior, const int32_t* npos, const int32_t* npos_norm, const uint8_t* cls_u8,
                          int B, int P, int C, int neg_ratio, float pos_iou,
                          double* sums, float* losses, float* grad_loc, float* grad_conf,
                          uint32_t* mined_mask, float* ce,
                          void* ws, size_t ws_bytes, void* stream)
{
    const int rc = ssdhead_ce_stream(conf, B, P, C, ce, grad_loc, grad_conf, ws, ws_bytes, stream);
    if (rc != 0) return rc;
    return ssdhead_mine(loc, conf, gt_xyxy, gt_cls, gt_off, pri_xyxy, pri_cxcywh, best_prior, npos, npos_norm, cls_u8,
                        B, P, C, neg_ratio, pos_iou, sums, losses, grad_loc, grad_conf, mined_mask, ce, ws, ws_bytes, stream);
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
