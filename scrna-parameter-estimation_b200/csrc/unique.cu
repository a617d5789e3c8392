// Compression of each (gene, group) segment to its distinct (count, size-factor bin) values.
//
// Replaces reference bootstrap.py:40-71 (_unique_expr: random-projection hash + np.unique sort of
// all cells of the group).  Here: one pass over the segment's NONZERO entries only, a shared-memory
// hash table keyed by (count << 8 | bin), then a bitonic sort of the U distinct keys so that the
// category order (and therefore the RNG stream consumption downstream) is deterministic.
// All zero-count cells are one implicit category: they contribute 0 to every moment, so merging
// them is exact for the multinomial resampling (marginalisation) -- SURVEY.md section 7 step 4.
//
// Output per segment, at pool offset seg_ptr[seg] - seg_ptr[seg_lo] (an upper bound on U needs no
// scan): prepared bootstrap entries {a = x/sf, b = (x^2 - (1-q) x)/sf^2, conditional probability,
// log(1-p), multiplicity, sampler mode} plus the raw (key, multiplicity) pairs for inspection.
//
// Tiers by segment nnz: <=768 one warp, table in that warp's shared memory; <=6144 one CTA, table in
// shared memory; larger one CTA, table in a global scratch region.  Tiers 2/3 are fed through a
// device-side list, so there is no host round trip.
#include "common.cuh"

namespace mm {

struct __align__(32) BootEntry {
    double a;     // contribution to M1 per resampled cell (before /n)
    double b;     // contribution to M2 per resampled cell (before /n)
    float p;      // conditional success probability given the remaining pool (<= 0.5 after flip)
    float lq;     // log(1 - p)
    int n;        // multiplicity in the data
    int mode;     // bit0: flipped (sample the complement), bit1: BTRS sampler, else inversion
};
static_assert(sizeof(BootEntry) == 32, "BootEntry layout");

constexpr uint32_t kEmpty = 0xFFFFFFFFu;
constexpr int kWarpCap = 1024, kWarpMaxNnz = 768;
constexpr int kCtaCap = 8192, kCtaMaxNnz = 6144, kCtaSmallCap = 2048;
constexpr int kUniqThreads = 256;
constexpr int kBtrsMin = 24;  // expected count above which the BTRS sampler is used

struct UniqueParams {
    const float* vals;
    const int* rows;
    const long long* seg_ptr;
    long long seg_lo;     // first segment of this tile (global segment index = gene * R + r)
    long long n_seg;      // segments in this tile
    int R;
    const unsigned char* cell_bin;   // per cell (group-sorted order)
    const double* bin_inv_sf;        // [n_bins]
    const double* group_q;           // [R]
    const int* group_ncells;         // [R]
    const int* group_nbins;          // [R] distinct bins present in the group
    int estimator;                   // 0 hyper_relative, 1 mean_only
    BootEntry* entries;              // pool
    uint32_t* raw_key;               // pool (may be null)
    int* raw_cnt;                    // pool (may be null)
    int* seg_U;                      // [n_seg] distinct nonzero categories; -1 => all-NaN bootstrap (U<=1 rule)
    int* big_list;                   // [0]=count, then segment indices (tile-relative)
    uint32_t* scratch_key;           // global tables for tier 3
    int* scratch_cnt;
};

__device__ __forceinline__ uint32_t hash32(uint32_t k) {
    k ^= k >> 16; k *= 0x7feb352du; k ^= k >> 15; k *= 0x846ca68bu; k ^= k >> 16;
    return k;
}

template <int NT>
__device__ __forceinline__ void group_sync() {
    if (NT == 32) __syncwarp(); else __syncthreads();
}

// Processes one segment with NT cooperating threads (t = thread index in the group).
// keys/cnts: hash table of `cap` slots (shared or global memory), zeroed/emptied here.  `limit` = most distinct keys
// the table may take: when more turn up the function gives up and returns false (nothing has been written to the
// outputs) and the caller retries with a larger table -- so a segment of 6000 nonzeros with its usual few hundred
// distinct (count, bin) values runs in a 2048-slot shared-memory table instead of a global one sized for 6000
// distinct keys.  Lanes of a warp that hold the same key are combined first (MATCH.ANY): in a dense segment a
// handful of keys (count 1-3 x the common bins) take most of the nonzeros, and 32 atomics on one shared-memory word
// serialise.
template <int NT>
__device__ bool unique_segment(const UniqueParams& P, long long seg_rel, uint32_t* keys, int* cnts, int cap, int limit,
                               int t, int* s_misc /* >= 4 ints of shared memory for this group */) {
    const long long seg = P.seg_lo + seg_rel;
    const long long lo = P.seg_ptr[seg], hi = P.seg_ptr[seg + 1];
    const long long pool = lo - P.seg_ptr[P.seg_lo];
    const int r = (int)(seg % P.R);
    const int mask = cap - 1;
    const int lane = t & 31;

    for (int i = t; i < cap; i += NT) { keys[i] = kEmpty; cnts[i] = 0; }
    if (t == 0) { s_misc[0] = 0; s_misc[1] = 0; s_misc[2] = 0; s_misc[3] = 0; }
    group_sync<NT>();

    // ---- hash insert of the nonzero entries
    int nz_local = 0;
    volatile int* overflow = s_misc + 3;
    constexpr int kUnroll = 4;                 // independent (value, row, bin) load chains in flight per thread
    for (long long base = lo; base < hi; base += kUnroll * NT) {
        float v[kUnroll];
        int row[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = base + u * NT + t;
            v[u] = i < hi ? ld_stream(P.vals + i) : 0.f;
            row[u] = i < hi ? ld_stream(P.rows + i) : 0;
        }
        uint32_t key[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u)
            key[u] = v[u] > 0.f ? (((uint32_t)v[u] << 8) | (uint32_t)__ldg(P.cell_bin + row[u])) : kEmpty;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const bool valid = key[u] != kEmpty;
            const unsigned vmask = __ballot_sync(kFull, valid);
            if (valid) {
                const unsigned peers = __match_any_sync(vmask, key[u]);
                if (lane == __ffs(peers) - 1) {                // one lane per distinct key of the warp
                    uint32_t h = hash32(key[u]) & mask;
                    while (!*overflow) {
                        const uint32_t prev = atomicCAS(keys + h, kEmpty, key[u]);
                        if (prev == kEmpty || prev == key[u]) {
                            atomicAdd(cnts + h, __popc(peers));
                            if (prev == kEmpty && atomicAdd(&s_misc[2], 1) >= limit) *overflow = 1;
                            break;
                        }
                        h = (h + 1) & mask;
                    }
                }
                ++nz_local;
            }
        }
    }
    if (nz_local) atomicAdd(&s_misc[1], nz_local);
    group_sync<NT>();
    if (*overflow) {
        group_sync<NT>();                                        // nobody resets s_misc while it is still being read
        return false;
    }

    // ---- in-place compaction to the front of the table (chunk reads precede chunk writes)
    int U = 0;
    for (int base = 0; base < cap; base += NT) {
        uint32_t k = keys[base + t];
        int c = cnts[base + t];
        bool occ = (k != kEmpty);
        int pos;
        if (NT == 32) {
            unsigned b = __ballot_sync(kFull, occ);
            pos = U + __popc(b & ((1u << t) - 1));
            U += __popc(b);
            __syncwarp();
        } else {
            // CTA: per-warp ballots combined through shared memory counter s_misc[0]
            unsigned b = __ballot_sync(kFull, occ);
            int lane = t & 31;
            int wbase = 0;
            if (lane == 0) wbase = atomicAdd(&s_misc[0], __popc(b));   // order among warps arbitrary; sorted later
            wbase = __shfl_sync(kFull, wbase, 0);
            pos = wbase + __popc(b & ((1u << lane) - 1));
            __syncthreads();
        }
        if (occ) { keys[pos] = k; cnts[pos] = c; }
        group_sync<NT>();
    }
    if (NT != 32) U = s_misc[0];
    // note (CTA path): writes go to pos < base + NT only if pos <= current chunk end; pos is bounded by the
    // number of occupied slots seen so far, which is <= base + NT, and all slots < base + NT were already read.

    // ---- bitonic sort of the U entries by key (pad to a power of two with kEmpty)
    int Ppow = 1;
    while (Ppow < U) Ppow <<= 1;
    for (int i = U + t; i < Ppow; i += NT) { keys[i] = kEmpty; cnts[i] = 0; }
    group_sync<NT>();
    for (int k = 2; k <= Ppow; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = t; i < Ppow; i += NT) {
                int l = i ^ j;
                if (l > i) {
                    uint32_t ki = keys[i], kl = keys[l];
                    bool up = ((i & k) == 0);
                    if ((ki > kl) == up) {
                        keys[i] = kl; keys[l] = ki;
                        int ci = cnts[i]; cnts[i] = cnts[l]; cnts[l] = ci;
                    }
                }
            }
            group_sync<NT>();
        }
    }

    // ---- prepared entries: conditional probabilities need the remaining pool size before each category
    const int n_cells = P.group_ncells[r];
    const int n_nonzero = s_misc[1];
    const int n_zero = n_cells - n_nonzero;
    const double q = P.group_q[r];
    // exclusive prefix sum of the multiplicities, in place into cnts' shadow: done serially per chunk
    int running = 0;  // identical in all threads
    for (int base = 0; base < U; base += NT) {
        int i = base + t;
        int c = (i < U) ? cnts[i] : 0;
        // inclusive scan across the group
        int incl = c;
        if (NT == 32) {
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(kFull, incl, o);
                if (t >= o) incl += y;
            }
        } else {
            int lane = t & 31, w = t >> 5;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int y = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += y;
            }
            __shared__ int wsum[kUniqThreads / 32];
            if (lane == 31) wsum[w] = incl;
            __syncthreads();
            int add = 0;
            for (int ww = 0; ww < w; ++ww) add += wsum[ww];
            incl += add;
            __syncthreads();
        }
        int before = running + incl - c;       // cells in earlier categories
        int chunk_total;
        if (NT == 32) chunk_total = __shfl_sync(kFull, incl, 31);
        else {
            __shared__ int tot;
            if (t == NT - 1) tot = incl;
            __syncthreads();
            chunk_total = tot;
            __syncthreads();
        }
        if (i < U) {
            uint32_t key = keys[i];
            double x = (double)(key >> 8);
            double w = P.bin_inv_sf[key & 0xFF];
            BootEntry e;
            e.a = x * w;
            e.b = (P.estimator == 0) ? (x * x - (1.0 - q) * x) * w * w : 0.0;
            e.n = c;
            long long rem = (long long)n_cells - before;   // pool the category is drawn from (includes the zeros)
            long long other = rem - c;
            int mode = 0;
            double pp;
            if (c > other) { mode |= 1; pp = (double)other / (double)rem; }
            else pp = (double)c / (double)rem;
            long long expect = c < other ? c : other;
            if (expect >= kBtrsMin) mode |= 2;
            e.p = (float)pp;
            e.lq = (float)log1p(-pp);
            e.mode = mode;
            P.entries[pool + i] = e;
            if (P.raw_key) { P.raw_key[pool + i] = key; P.raw_cnt[pool + i] = c; }
        }
        running += chunk_total;
    }
    if (t == 0) {
        // U<=1 rule of reference bootstrap.py:97-98, counting the zero cells' distinct bins
        bool all_nan = (U == 1 && n_zero == 0) || (U == 0 && P.group_nbins[r] <= 1);
        P.seg_U[seg_rel] = all_nan ? -1 : U;
    }
    group_sync<NT>();
    return true;
}

__global__ void __launch_bounds__(kUniqThreads)
unique_warp_kernel(UniqueParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kWarps = kUniqThreads / 32;
    uint32_t* keys = reinterpret_cast<uint32_t*>(smem) + warp * kWarpCap;
    int* cnts = reinterpret_cast<int*>(smem + kWarps * kWarpCap * 4) + warp * kWarpCap;
    int* misc = reinterpret_cast<int*>(smem + 2 * kWarps * kWarpCap * 4) + warp * 4;
    long long seg_rel = (long long)blockIdx.x * kWarps + warp;
    if (seg_rel >= P.n_seg) return;
    long long seg = P.seg_lo + seg_rel;
    long long nnz = P.seg_ptr[seg + 1] - P.seg_ptr[seg];
    if (nnz > kWarpMaxNnz) {
        if (lane == 0) P.big_list[1 + atomicAdd(P.big_list, 1)] = (int)seg_rel;
        return;
    }
    unique_segment<32>(P, seg_rel, keys, cnts, kWarpCap, kWarpCap, lane, misc);        // nnz <= 768: cannot overflow
}

__global__ void __launch_bounds__(kUniqThreads)
unique_cta_kernel(UniqueParams P) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* skeys = reinterpret_cast<uint32_t*>(smem);
    int* scnts = reinterpret_cast<int*>(smem + kCtaCap * 4);
    __shared__ int misc[4];
    const int n_big = P.big_list[0];
    for (int b = blockIdx.x; b < n_big; b += gridDim.x) {
        long long seg_rel = P.big_list[1 + b];
        long long seg = P.seg_lo + seg_rel;
        long long lo = P.seg_ptr[seg];
        long long nnz = P.seg_ptr[seg + 1] - lo;
        // small shared-memory table first (the distinct (count, bin) values of a segment are usually a few hundred),
        // then the full one, then -- only for segments that may really hold more distinct keys than that -- global memory
        bool done = unique_segment<kUniqThreads>(P, seg_rel, skeys, scnts, kCtaSmallCap, kCtaSmallCap * 3 / 4, threadIdx.x, misc);
        if (!done) done = unique_segment<kUniqThreads>(P, seg_rel, skeys, scnts, kCtaCap, kCtaMaxNnz, threadIdx.x, misc);
        if (!done) {
            int cap = 1;
            while (cap < nnz + (nnz >> 1)) cap <<= 1;           // <= 3 * nnz
            long long off = 3 * (lo - P.seg_ptr[P.seg_lo]);
            unique_segment<kUniqThreads>(P, seg_rel, P.scratch_key + off, P.scratch_cnt + off, cap, cap, threadIdx.x, misc);
        }
        __syncthreads();
    }
}

}  // namespace mm

using namespace mm;

// mm_seg_unique: see include/memento_b200.h
MM_EXPORT int mm_seg_unique(int device, void* stream, const float* vals, const int32_t* rows,
                            const int64_t* seg_ptr, int64_t seg_lo, int64_t n_seg, int32_t R,
                            const uint8_t* cell_bin, const double* bin_inv_sf, int32_t n_bins,
                            const double* group_q, const int32_t* group_ncells, const int32_t* group_nbins,
                            int32_t estimator, void* entries, uint32_t* raw_key, int32_t* raw_cnt,
                            int32_t* seg_U, int32_t* big_list, uint32_t* scratch_key, int32_t* scratch_cnt) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_seg >= 0 && R > 0, "n_seg/R");
    MM_REQUIRE(n_bins > 0 && n_bins <= 256, "n_bins must be in 1..256");
    MM_REQUIRE(estimator == 0 || estimator == 1, "estimator");
    if (n_seg == 0) return 0;
    MM_REQUIRE(vals && rows && seg_ptr && cell_bin && bin_inv_sf && group_q && group_ncells && group_nbins &&
               entries && seg_U && big_list && scratch_key && scratch_cnt, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    UniqueParams P;
    P.vals = vals; P.rows = rows; P.seg_ptr = (const long long*)seg_ptr; P.seg_lo = seg_lo; P.n_seg = n_seg;
    P.R = R; P.cell_bin = cell_bin; P.bin_inv_sf = bin_inv_sf; P.group_q = group_q;
    P.group_ncells = group_ncells; P.group_nbins = group_nbins; P.estimator = estimator;
    P.entries = (BootEntry*)entries; P.raw_key = raw_key; P.raw_cnt = raw_cnt; P.seg_U = seg_U;
    P.big_list = big_list; P.scratch_key = scratch_key; P.scratch_cnt = scratch_cnt;
    MM_CUDA(cudaMemsetAsync(big_list, 0, sizeof(int32_t), st));
    constexpr int kWarps = kUniqThreads / 32;
    size_t smem_warp = 2 * kWarps * kWarpCap * 4 + kWarps * 4 * 4;
    size_t smem_cta = 2 * kCtaCap * 4;
    static bool attr_done = false;
    if (!attr_done) {
        MM_CUDA(cudaFuncSetAttribute(unique_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_warp));
        MM_CUDA(cudaFuncSetAttribute(unique_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_cta));
        attr_done = true;
    }
    long long blocks = (n_seg + kWarps - 1) / kWarps;
    MM_REQUIRE(blocks < 2147483647LL, "too many segments for one launch");
    unique_warp_kernel<<<(unsigned)blocks, kUniqThreads, smem_warp, st>>>(P);
    if (int s = check_launch("unique_warp")) return s;
    unique_cta_kernel<<<148 * 3, kUniqThreads, smem_cta, st>>>(P);
    return check_launch("unique_cta");
}
