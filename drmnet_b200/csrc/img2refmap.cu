// K2: image -> refmap scatter as a deterministic sort-by-bin segmented select (sm_100a).
//
// Replaces refmap_mask_make (reference utils/img2refmap.py:6-37) and the xyz2thetaphi call inside it
// (utils/transform.py:84-89).  The reference tests every (cell, pixel) pair (O(res^2 n), [512,n,2] temporaries);
// here each pixel generates only the cells whose fp32 window predicate it can satisfy, the (cell, pixel) pairs are
// counting-sorted by cell (histogram -> exclusive scan -> scatter), and one warp per cell selects the lower median
// under the total order (channel-sum, pixel index).  Integer atomics only place pairs inside a cell's segment; the
// selection is order-independent, so every output is bit-exact and run-to-run deterministic.
//
// HBM-bound integer/compare work: no tensor cores.  Traffic per image: normals and colours are read once (the first
// pass leaves each pixel's cell id for the second), 8 B per pair written and read once, 4 int32 per cell.  Atomics are
// aggregated per warp (__match_any_sync): raster-ordered pixels of an image mostly share their neighbours' cell.
#include <math.h>

#include "common.cuh"

namespace drm {

static constexpr uint32_t KEY_NAN = 0xFFFFFFFFu;  // sentinel: colour sum is NaN (ignored by nanmedian, :30-31)
static constexpr uint32_t CELL_NONE = 0x7FFFFFFFu, CELL_MULTI = 0x80000000u;
static constexpr int SMALL_MAX = 48;  // cells up to this size are selected by one thread

struct I2RArgs {
    const float* colors;
    const float* geom;  // normals [n,3] or thetaphi [n,2]
    const int64_t* offsets;
    int64_t total_n;
    int B, C, res, res2;
    int is_thetaphi;
    float thr, stepf, inv_step;
    int R;  // candidate radius in cells
    int min_points, reduce_mode;
    int64_t pair_capacity;
    // workspace
    int32_t* bin_count;   // [M]
    int32_t* bin_offset;  // [M]
    int32_t* cursor;      // [M]
    int32_t* nan_count;   // [B]
    int32_t* nan_list;    // [total_n]  (image b owns the slice starting at offsets[b])
    int32_t* block_sums;  // scan scratch
    uint2* pairs;         // (key, global pixel index)
    uint32_t* cell0;      // [total_n] first cell of each pixel (CELL_NONE if none), bit 31 set if it has more cells
    int32_t* big_list;    // [M] cells left to the warp-per-cell select (large, NaN-angle members, mean mode)
    int32_t* big_count;   // [1]
    int32_t* status;      // bit 0: pair buffer overflow
    // outputs
    float* refmap;
    uint8_t* refmask;
    int32_t* counts;
    int32_t* sel_index;
};

// theta = acos(n . [0,1,0]), phi = atan2(n . [0,0,1], n . [-1,0,0]) with the dot products evaluated literally
// (utils/transform.py:87-88 with normal=[0,1,0], tangent=[-1,0,0], binormal=cross=[0,0,1]); IEEE acosf/atan2f.
__device__ __forceinline__ void normal_to_thetaphi(float x, float y, float z, float& th, float& ph) {
    float ny = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.f), __fmul_rn(y, 1.f)), __fmul_rn(z, 0.f));
    float nt = __fadd_rn(__fadd_rn(__fmul_rn(x, -1.f), __fmul_rn(y, 0.f)), __fmul_rn(z, 0.f));
    float nb = __fadd_rn(__fadd_rn(__fmul_rn(x, 0.f), __fmul_rn(y, 0.f)), __fmul_rn(z, 1.f));
    th = acosf(ny);
    ph = atan2f(nb, nt);
}

__global__ void normals_to_thetaphi_kernel(const float* __restrict__ normals, int64_t n, float* __restrict__ out) {
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    float th, ph;
    normal_to_thetaphi(normals[3 * p], normals[3 * p + 1], normals[3 * p + 2], th, ph);
    out[2 * p] = th;
    out[2 * p + 1] = ph;
}

__device__ __forceinline__ int image_of(const int64_t* __restrict__ offsets, int B, int64_t p) {
    int lo = 0, hi = B;  // find b with offsets[b] <= p < offsets[b+1]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (offsets[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// image of pixel p: one binary search per CTA (for its first pixel, by thread 0), then a short forward walk per thread
__device__ __forceinline__ int image_of_cta(const int64_t* __restrict__ offsets, int B, int64_t p, int64_t cta_first) {
    __shared__ int b0;
    if (threadIdx.x == 0) b0 = image_of(offsets, B, cta_first);
    __syncthreads();
    int b = b0;
    while (b + 1 < B && offsets[b + 1] <= p) ++b;
    return b;
}

// monotone uint key of the fp32 channel sum; -0.0 and +0.0 share a key (torch compares them equal)
__device__ __forceinline__ uint32_t sum_key(const float* __restrict__ c, int C) {
    float s = c[0];
    for (int k = 1; k < C; ++k) s = __fadd_rn(s, c[k]);  // ((c0 + c1) + c2), colors.sum(-1) at img2refmap.py:30
    if (isnan(s)) return KEY_NAN;
    s = __fadd_rn(s, 0.0f);  // -0.0 -> +0.0
    uint32_t b = __float_as_uint(s);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// One warp-wide increment per distinct cell: lanes that target the same cell elect a leader (raster-ordered pixels
// mostly share their neighbours' cell).  Returns the lane's rank inside its group and the group's base (PASS 1).
__device__ __forceinline__ int warp_claim(int32_t* counter, bool valid, int64_t gbin, bool want_base, int& rank) {
    const unsigned lane = threadIdx.x & 31;
    const unsigned long long key = valid ? (unsigned long long)gbin : (0x8000000000000000ull | lane);
    const unsigned grp = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(grp) - 1;
    rank = __popc(grp & ((1u << lane) - 1u));
    int base = 0;
    if (valid && (int)lane == leader) {
        if (want_base) base = atomicAdd(&counter[gbin], __popc(grp));
        else atomicAdd(&counter[gbin], __popc(grp));
    }
    if (want_base) base = __shfl_sync(0xffffffffu, base, leader);
    return base;
}

// PASS 0: angles, window predicate, histogram of the cells, first cell of each pixel left in cell0
__global__ void __launch_bounds__(256) i2r_hist_pass(I2RArgs a) {
    const int64_t cta_first = blockIdx.x * (int64_t)blockDim.x;
    const int64_t p = cta_first + threadIdx.x;
    const bool in_range = p < a.total_n;
    const int b = image_of_cta(a.offsets, a.B, in_range ? p : a.total_n - 1, cta_first);
    if (!in_range) return;
    const int64_t base = a.offsets[b];
    float th, ph;
    if (a.is_thetaphi) {
        th = a.geom[2 * p];
        ph = a.geom[2 * p + 1];
    } else {
        normal_to_thetaphi(a.geom[3 * p], a.geom[3 * p + 1], a.geom[3 * p + 2], th, ph);
    }
    if (isnan(th) || isnan(ph)) {
        // 'NaN > thr' is False (img2refmap.py:27): the pixel is a member of every cell of its image
        const int slot = atomicAdd(&a.nan_count[b], 1);
        a.nan_list[base + slot] = (int32_t)(p - base);
        a.cell0[p] = CELL_NONE;
        return;
    }
    bool have = false, multi = false;
    int64_t first = 0;
    // candidate cells: those whose centre can lie within thr of the angle, +-1 for rounding; the fp32 predicate below
    // alone decides membership
    const float lim = (float)a.res + (float)a.R + 1.f;
    const float flo_i = floorf((th - a.thr) * a.inv_step - 0.5f), fhi_i = floorf((th + a.thr) * a.inv_step - 0.5f);
    const float flo_j = floorf((ph - a.thr) * a.inv_step - 0.5f), fhi_j = floorf((ph + a.thr) * a.inv_step - 0.5f);
    if (fhi_i > -lim && flo_i < lim && fhi_j > -lim && flo_j < lim) {  // also rejects +-inf
        const int ilo = max((int)fmaxf(flo_i, -lim), 0), ihi = min((int)fminf(fhi_i, lim) + 1, a.res - 1);
        const int jlo = max((int)fmaxf(flo_j, -lim), 0), jhi = min((int)fminf(fhi_j, lim) + 1, a.res - 1);
        for (int i = ilo; i <= ihi; ++i) {
            const float ci = __fmul_rn((float)i + 0.5f, a.stepf);  // (arange + 0.5) * (pi / res), img2refmap.py:16
            if (fabsf(__fsub_rn(ci, th)) > a.thr) continue;
            for (int j = jlo; j <= jhi; ++j) {
                const float cj = __fmul_rn((float)j + 0.5f, a.stepf);
                if (fabsf(__fsub_rn(cj, ph)) > a.thr) continue;
                const int64_t gbin = (int64_t)b * a.res2 + (int64_t)i * a.res + j;
                if (!have) { have = true; first = gbin; }
                else multi = true;
                atomicAdd(&a.bin_count[gbin], 1);
            }
        }
    }
    a.cell0[p] = have ? ((uint32_t)first | (multi ? CELL_MULTI : 0u)) : CELL_NONE;
}

// PASS 1: scatter (sum key, pixel) pairs into the cell segments
__global__ void __launch_bounds__(256) i2r_scatter_pass(I2RArgs a) {
    const int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool in_range = p < a.total_n;
    const uint32_t c0 = in_range ? a.cell0[p] : CELL_NONE;
    const bool have = c0 != CELL_NONE;
    const int64_t first = (int64_t)(c0 & ~CELL_MULTI);
    uint32_t key = 0;
    if (have) key = sum_key(a.colors + p * a.C, a.C);
    int rank;
    const int base = warp_claim(a.cursor, have, first, true, rank);
    if (have) {
        const int64_t slot = (int64_t)a.bin_offset[first] + base + rank;
        if (slot < a.pair_capacity) a.pairs[slot] = make_uint2(key, (uint32_t)p);
        else atomicOr(a.status, 1);
    }
    if (have && (c0 & CELL_MULTI)) {
        // the pixel's other cells: recompute the window (same arithmetic as the histogram pass), skip the first one
        const int b = image_of(a.offsets, a.B, p);
        float th, ph;
        if (a.is_thetaphi) {
            th = a.geom[2 * p];
            ph = a.geom[2 * p + 1];
        } else {
            normal_to_thetaphi(a.geom[3 * p], a.geom[3 * p + 1], a.geom[3 * p + 2], th, ph);
        }
        const float lim = (float)a.res + (float)a.R + 1.f;
        const float flo_i = floorf((th - a.thr) * a.inv_step - 0.5f), fhi_i = floorf((th + a.thr) * a.inv_step - 0.5f);
        const float flo_j = floorf((ph - a.thr) * a.inv_step - 0.5f), fhi_j = floorf((ph + a.thr) * a.inv_step - 0.5f);
        const int ilo = max((int)fmaxf(flo_i, -lim), 0), ihi = min((int)fminf(fhi_i, lim) + 1, a.res - 1);
        const int jlo = max((int)fmaxf(flo_j, -lim), 0), jhi = min((int)fminf(fhi_j, lim) + 1, a.res - 1);
        for (int i = ilo; i <= ihi; ++i) {
            const float ci = __fmul_rn((float)i + 0.5f, a.stepf);
            if (fabsf(__fsub_rn(ci, th)) > a.thr) continue;
            for (int j = jlo; j <= jhi; ++j) {
                const float cj = __fmul_rn((float)j + 0.5f, a.stepf);
                if (fabsf(__fsub_rn(cj, ph)) > a.thr) continue;
                const int64_t gbin = (int64_t)b * a.res2 + (int64_t)i * a.res + j;
                if (gbin == first) continue;
                const int64_t slot = (int64_t)a.bin_offset[gbin] + atomicAdd(&a.cursor[gbin], 1);
                if (slot < a.pair_capacity) a.pairs[slot] = make_uint2(key, (uint32_t)p);
                else atomicOr(a.status, 1);
            }
        }
    }
}

// ---- exclusive scan of bin_count (3 small kernels; M <= B*res^2 ints) ----------------------------------------
static constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int& total) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_sums[w] = inc;
    __syncthreads();
    if (w == 0) {
        int s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int wprefix = w ? warp_sums[w - 1] : 0;
    total = warp_sums[SCAN_THREADS / 32 - 1];
    __syncthreads();
    return wprefix + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_tiles(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                           int32_t* __restrict__ block_sums, int64_t M) {
    const int64_t start = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (start + k < M) ? in[start + k] : 0;
        s += v[k];
    }
    int total;
    int ex = block_exclusive_scan(s, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (start + k < M) out[start + k] = ex;
        ex += v[k];
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(int32_t* block_sums, int nblocks) {
    int carry = 0;
    for (int start = 0; start < nblocks; start += SCAN_THREADS) {
        int i = start + threadIdx.x;
        int v = i < nblocks ? block_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, total);
        if (i < nblocks) block_sums[i] = carry + ex;
        carry += total;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_add(int32_t* __restrict__ out, const int32_t* __restrict__ block_sums,
                                                         int64_t M) {
    const int add = block_sums[blockIdx.x];
    const int64_t start = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (start + k < M) out[start + k] += add;
}

// ---- selection of the lower median under the total order (sum key, pixel index) ----------------------------------
struct CellView {
    const uint2* seg;
    int nreg;
    const int32_t* nan_list;  // image-local indices
    int nnan;
    int64_t base;             // first pixel of the image
    const float* colors;      // global
    int C;
    __device__ __forceinline__ uint64_t get(int e) const {  // (key << 32) | global pixel index
        if (e < nreg) {
            uint2 pr = seg[e];
            return ((uint64_t)pr.x << 32) | pr.y;
        }
        const int64_t idx = base + nan_list[e - nreg];
        return ((uint64_t)sum_key(colors + idx * C, C) << 32) | (uint32_t)idx;
    }
};

__device__ __forceinline__ void write_cell(const I2RArgs& a, int64_t gbin, int64_t base, bool filled, int cnt,
                                           int64_t winner) {
    for (int c = 0; c < a.C; ++c) a.refmap[gbin * a.C + c] = filled ? a.colors[winner * a.C + c] : 0.f;
    a.refmask[gbin] = filled ? 1 : 0;
    if (a.counts) a.counts[gbin] = cnt;
    if (a.sel_index) a.sel_index[gbin] = filled ? (int32_t)(winner - base) : -1;
}

// k-th smallest of arr[0..n) (distinct 64-bit values), Hoare quickselect in place
__device__ __forceinline__ uint64_t quickselect(uint64_t* arr, int n, int k) {
    int l = 0, r = n - 1;
    while (l < r) {
        const uint64_t pivot = arr[(l + r) >> 1];
        int i = l, j = r;
        while (i <= j) {
            while (arr[i] < pivot) ++i;
            while (arr[j] > pivot) --j;
            if (i <= j) {
                const uint64_t t = arr[i]; arr[i] = arr[j]; arr[j] = t;
                ++i; --j;
            }
        }
        if (k <= j) r = j;
        else if (k >= i) l = i;
        else break;
    }
    return arr[k];
}

// One thread per cell: cells with up to SMALL_MAX members and no NaN-angle members, median mode; the others are queued
// for the warp-per-cell kernel.  The consecutive cells of a CTA own one contiguous run of pairs: it is staged in
// shared memory with coalesced loads and each thread quickselects its own segment there.
static constexpr int SELECT_CAP = 6144;  // pairs staged per CTA (48 KB)
static constexpr int SELECT_THREADS = 128;  // cells per CTA: 128 x SMALL_MAX members always fit the staging buffer
__global__ void __launch_bounds__(SELECT_THREADS) i2r_select_small(I2RArgs a, int64_t M) {
    __shared__ uint64_t buf[SELECT_CAP];
    const int64_t g0 = blockIdx.x * (int64_t)blockDim.x;
    const int64_t gbin = g0 + threadIdx.x;
    const int64_t glast = min(g0 + (int64_t)blockDim.x, M) - 1;
    const int64_t run0 = a.bin_offset[g0];
    const int64_t run1 = (int64_t)a.bin_offset[glast] + a.bin_count[glast];
    const int run = (int)(run1 - run0);
    const bool staged = run <= SELECT_CAP;
    if (staged)
        for (int e = threadIdx.x; e < run; e += blockDim.x) {
            const uint2 pr = a.pairs[run0 + e];
            buf[e] = ((uint64_t)pr.x << 32) | pr.y;
        }
    __syncthreads();
    if (gbin >= M) return;
    const int b = (int)(gbin / a.res2);
    const int nreg = a.bin_count[gbin], nnan = a.nan_count[b];
    const int cnt = nreg + nnan;  // (~angle_mask).sum(-1), img2refmap.py:28
    const int64_t base = a.offsets[b];
    if (cnt == 0 || cnt < a.min_points) {
        write_cell(a, gbin, base, false, cnt, 0);
        return;
    }
    if (nnan > 0 || nreg > SMALL_MAX || a.reduce_mode != 0 || !staged) {
        a.big_list[atomicAdd(a.big_count, 1)] = (int32_t)gbin;
        return;
    }
    uint64_t* seg = buf + (a.bin_offset[gbin] - run0);
    int nvalid = 0;
    for (int e = 0; e < nreg; ++e) nvalid += (uint32_t)(seg[e] >> 32) != KEY_NAN;
    if (nvalid == 0) {
        write_cell(a, gbin, base, false, cnt, 0);
        return;
    }
    // NaN keys are the largest values, so rank (nvalid - 1) / 2 of the whole segment is the lower median of the valid
    // members (torch.nanmedian, :31)
    const uint64_t med = quickselect(seg, nreg, (nvalid - 1) >> 1);
    write_cell(a, gbin, base, true, cnt, (int64_t)(uint32_t)med);
}

// one warp per queued cell
__global__ void __launch_bounds__(256) i2r_select_big(I2RArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int nbig = *a.big_count;
    for (int64_t q = warp0; q < nbig; q += nwarps) {
        const int64_t gbin = a.big_list[q];
        const int b = (int)(gbin / a.res2);
        const int64_t base = a.offsets[b];
        CellView cv;
        cv.nreg = a.bin_count[gbin];
        cv.seg = a.pairs + a.bin_offset[gbin];
        cv.nnan = a.nan_count[b];
        cv.nan_list = a.nan_list + base;
        cv.base = base;
        cv.colors = a.colors;
        cv.C = a.C;
        const int cnt = cv.nreg + cv.nnan;

        int64_t winner = -1;
        float mean_c = 0.f;
        bool filled = false;
        // valid = members whose sum is not NaN
        int nvalid = 0;
        for (int e0 = 0; e0 < cnt; e0 += 32) {
            int e = e0 + lane;
            bool v = e < cnt && (uint32_t)(cv.get(e) >> 32) != KEY_NAN;
            nvalid += __popc(__ballot_sync(0xffffffffu, v));
        }
        if (nvalid > 0) {
            filled = true;
            if (a.reduce_mode == 0) {
                const int k = (nvalid - 1) >> 1;  // lower median, torch.nanmedian (:31)
                for (int e0 = 0; e0 < cnt && winner < 0; e0 += 32) {
                    const int e = e0 + lane;
                    const uint64_t mine = e < cnt ? cv.get(e) : ~0ull;
                    int rank = 0;
                    for (int f = 0; f < cnt; ++f) rank += cv.get(f) < mine;
                    unsigned hit = __ballot_sync(0xffffffffu, e < cnt && rank == k);
                    if (hit) winner = (int64_t)(uint32_t)__shfl_sync(0xffffffffu, mine, __ffs(hit) - 1);
                }
            } else {
                // mean: fp32 sum over valid members in ascending pixel order; lane c owns channel c
                int64_t prev = -1;
                float acc = 0.f;
                for (int r = 0; r < nvalid; ++r) {
                    long long best = 0x7fffffffffffffffll;
                    for (int e0 = 0; e0 < cnt; e0 += 32) {
                        int e = e0 + lane;
                        if (e < cnt) {
                            uint64_t el = cv.get(e);
                            long long idx = (long long)(uint32_t)el;
                            if ((uint32_t)(el >> 32) != KEY_NAN && idx > prev) best = min(best, idx);
                        }
                    }
                    for (int d = 16; d; d >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, d));
                    if (lane < a.C) acc = __fadd_rn(acc, cv.colors[best * a.C + lane]);
                    prev = best;
                }
                mean_c = __fdiv_rn(acc, (float)nvalid);
            }
        }
        if (lane < a.C) {
            float v = 0.f;
            if (filled) v = a.reduce_mode == 0 ? cv.colors[winner * a.C + lane] : mean_c;
            a.refmap[gbin * a.C + lane] = v;
        }
        if (lane == 0) {
            a.refmask[gbin] = filled ? 1 : 0;
            if (a.counts) a.counts[gbin] = cnt;
            if (a.sel_index) a.sel_index[gbin] = (filled && a.reduce_mode == 0) ? (int32_t)(winner - base) : -1;
        }
    }
}

static int64_t pairs_per_pixel_bound(int res, float thr) {
    const double step = M_PI / res;
    int64_t per_axis = (int64_t)floor(2.0 * (double)thr / step + 1e-3) + 1;
    if (per_axis > res) per_axis = res;
    if (per_axis < 1) per_axis = 1;
    return per_axis * per_axis;
}

static size_t i2r_carve(I2RArgs& a, void* ws, int64_t total_n, int B, int res, float thr) {
    const int64_t M = (int64_t)B * res * res;
    const int64_t nblocks = (M + SCAN_TILE - 1) / SCAN_TILE;
    Carver c(ws);
    a.bin_count = c.take<int32_t>(M);
    a.cursor = c.take<int32_t>(M);  // contiguous with bin_count + nan_count + status for one memset
    a.nan_count = c.take<int32_t>(B);
    a.big_count = c.take<int32_t>(1);
    a.status = c.take<int32_t>(1);
    const size_t zero_end = c.used();
    a.bin_offset = c.take<int32_t>(M);
    a.nan_list = c.take<int32_t>(total_n > 0 ? total_n : 1);
    a.cell0 = c.take<uint32_t>(total_n > 0 ? total_n : 1);
    a.big_list = c.take<int32_t>(M);
    a.block_sums = c.take<int32_t>(nblocks + 1);
    a.pair_capacity = total_n * pairs_per_pixel_bound(res, thr);
    a.pairs = c.take<uint2>(a.pair_capacity > 0 ? a.pair_capacity : 1);
    (void)zero_end;
    return c.used();
}

}  // namespace drm

using namespace drm;

extern "C" size_t drm_img2refmap_workspace_bytes(int64_t total_n, int B, int res, float thr) {
    if (total_n < 0 || B <= 0 || res <= 0 || !(thr >= 0.f)) return 0;
    I2RArgs a{};
    return i2r_carve(a, nullptr, total_n, B, res, thr);
}

extern "C" int drm_img2refmap(const float* colors, const float* geom, int input_is_thetaphi, const int64_t* offsets,
                              int64_t total_n, int B, int C, int res, float thr, int min_points, int reduce_mode,
                              float* refmap, uint8_t* refmask, int32_t* counts, int32_t* sel_index, void* workspace,
                              size_t workspace_bytes, void* cuda_stream) {
    DRM_REQUIRE(B > 0 && res > 0 && total_n >= 0, "img2refmap: B=%d res=%d total_n=%lld must be positive", B, res, (long long)total_n);
    DRM_REQUIRE(C >= 1 && C <= 4, "img2refmap: C=%d not in 1..4", C);
    DRM_REQUIRE(thr >= 0.f, "img2refmap: angle threshold must be a non-negative number");
    DRM_REQUIRE(reduce_mode == 0 || reduce_mode == 1, "img2refmap: reduce_mode %d (0 = median, 1 = mean)", reduce_mode);
    DRM_REQUIRE(offsets && refmap && refmask && (total_n == 0 || (colors && geom)), "img2refmap: null pointer");
    const int64_t M = (int64_t)B * res * res;
    DRM_REQUIRE(M < (1ll << 31), "img2refmap: B*res*res = %lld exceeds int32 cells", (long long)M);
    const int64_t per_px = pairs_per_pixel_bound(res, thr);
    if (total_n * per_px >= (1ll << 31)) {
        set_error("img2refmap: up to %lld (cell,pixel) pairs exceed the int32 segment offsets", (long long)(total_n * per_px));
        return DRM_EUNSUPPORTED;
    }
    I2RArgs a{};
    const size_t need = i2r_carve(a, workspace, total_n, B, res, thr);
    if (!workspace || workspace_bytes < need) {
        set_error("img2refmap: workspace of %zu bytes needed, %zu given", need, workspace_bytes);
        return DRM_EWORKSPACE;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    a.colors = colors; a.geom = geom; a.offsets = offsets; a.total_n = total_n;
    a.B = B; a.C = C; a.res = res; a.res2 = res * res; a.is_thetaphi = input_is_thetaphi;
    a.thr = thr;
    a.stepf = (float)(M_PI / res);  // python float pi/res rounded to fp32 once (img2refmap.py:16)
    a.inv_step = (float)(res / M_PI);
    a.R = (int)ceil((double)thr / (M_PI / res)) + 1;
    if (a.R > res) a.R = res;
    a.min_points = min_points; a.reduce_mode = reduce_mode;
    a.refmap = refmap; a.refmask = refmask; a.counts = counts; a.sel_index = sel_index;

    // bin_count, cursor, nan_count, big_count, status are carved first and contiguous (each 256-aligned)
    const size_t zero_bytes = (size_t)((char*)a.bin_offset - (char*)a.bin_count);
    DRM_CHECK_CUDA(cudaMemsetAsync(a.bin_count, 0, zero_bytes, st));
    const int64_t nblocks = (M + SCAN_TILE - 1) / SCAN_TILE;
    if (total_n > 0) {
        const unsigned grid = (unsigned)((total_n + 255) / 256);
        i2r_hist_pass<<<grid, 256, 0, st>>>(a);
    }
    scan_tiles<<<(unsigned)nblocks, SCAN_THREADS, 0, st>>>(a.bin_count, a.bin_offset, a.block_sums, M);
    scan_block_sums<<<1, SCAN_THREADS, 0, st>>>(a.block_sums, (int)nblocks);
    scan_add<<<(unsigned)nblocks, SCAN_THREADS, 0, st>>>(a.bin_offset, a.block_sums, M);
    if (total_n > 0) {
        const unsigned grid = (unsigned)((total_n + 255) / 256);
        i2r_scatter_pass<<<grid, 256, 0, st>>>(a);
    }
    i2r_select_small<<<(unsigned)((M + SELECT_THREADS - 1) / SELECT_THREADS), SELECT_THREADS, 0, st>>>(a, M);
    i2r_select_big<<<148 * 4, 256, 0, st>>>(a);
    DRM_CHECK_CUDA(cudaGetLastError());
    count_launches(total_n > 0 ? 7 : 5);
    return DRM_OK;
}

extern "C" int drm_normals_to_thetaphi(const float* normals, int64_t n, float* thetaphi, void* cuda_stream) {
    DRM_REQUIRE(n >= 0 && (n == 0 || (normals && thetaphi)), "normals_to_thetaphi: bad arguments");
    if (n == 0) return DRM_OK;
    normals_to_thetaphi_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(normals, n, thetaphi);
    DRM_CHECK_CUDA(cudaGetLastError());
    return DRM_OK;
}
