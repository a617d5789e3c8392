// Ingest re-layout: CSR (cells x genes, as uploaded) -> group-sorted CSC (segment gene * R + group holds the
// nonzeros of one gene in one group, cells renumbered so that a group is a contiguous row range, rows ascending
// inside a segment).
//
// Replaces reference main.py:115-132 + util.py:8-13 (create_groups: one boolean scan of the obs column and one
// X[mask].tocsc() copy per group) and, with a single group, the CSC view of all cells that setup_memento's moment
// passes read.  The first version of this library did it with a stable 64-bit radix sort of the nonzeros (torch);
// this is a stable counting transposition instead, three passes and no sort:
//
//   rows (in NEW order, i.e. group by group) are cut into chunks of kRowsPerChunk consecutive rows of one group;
//   1. count   one warp per chunk: cnt[chunk][gene] = nonzeros of the chunk in that gene        (global atomics,
//              only between the lanes of the owning warp)
//   2. scan    one thread per (group, gene): exclusive prefix over the group's chunks (in place), total ->
//              seg_len[gene * R + group]; the caller turns seg_len into seg_ptr (prefix sum over 64-bit offsets)
//   3. fill    one warp per chunk walks its rows IN ORDER; a row's nonzeros have distinct genes, so the lanes can
//              take and advance the chunk's per-gene cursors without atomics: position = seg_ptr[gene * R + group]
//              + cnt[chunk][gene]++.  Rows ascend inside every segment by construction.
//
// That generic path scatters lone 4-byte stores and has one warp per chunk (round 1: 0.16 TB/s, 2.5 % of HBM).  When
// the column indices of every CSR row ascend (scipy's canonical form) passes 1 and 3 run as a TILED transposition
// through shared memory instead (relayout_tile_kernel): a CTA owns a chunk of <= 256 rows (one per thread) and walks a
// range of 256-gene blocks.  Per block: every row's nonzeros of the block are found by walking the sorted row from where
// the previous block ended (no search), their (gene, row) incidence goes into a 256 x 256 bit matrix in shared memory,
// rank of an element inside its gene's run = popcount of the lower rows' bits -- so the block's nonzeros are placed
// in a shared staging buffer sorted by (gene, row) with no ordering constraint between warps, and go out as one
// coalesced run per gene (chunk rows x density elements: ~240 B on the 25k x 10k matrix) instead of 4-byte scatters.
// Bit-identical to the generic path and to the stable sort (tests/test_gpu_relayout.py).
#include "common.cuh"
#include <string.h>
#include <mutex>
#include <thread>
#include <vector>

namespace mm {

constexpr int kRelayoutThreads = 128;

struct RelayoutParams {
    const long long* indptr;     // CSR of the uploaded matrix (original cell order)
    const int* indices;
    const float* data;
    const int* order;            // [n_rows] original cell of every new row (nullable: identity)
    const int* chunk_row_lo;     // [n_chunks + 1] first new row of every chunk
    const int* chunk_group;      // [n_chunks]
    int n_chunks, n_genes, R;
    int* cnt;                    // [n_chunks][n_genes]
    int* err;                    // tiled path: set to 1 when a row's column indices do not ascend (nullable)
};

__global__ void __launch_bounds__(kRelayoutThreads)
relayout_count_kernel(RelayoutParams P) {
    const int lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * (kRelayoutThreads / 32) + (threadIdx.x >> 5);
    if (chunk >= P.n_chunks) return;
    int* cnt = P.cnt + (long long)chunk * P.n_genes;
    for (int r = P.chunk_row_lo[chunk]; r < P.chunk_row_lo[chunk + 1]; ++r) {
        const long long cell = P.order ? P.order[r] : r;
        const long long lo = P.indptr[cell], hi = P.indptr[cell + 1];
        for (long long e = lo + lane; e < hi; e += 32) atomicAdd(cnt + ld_stream(P.indices + e), 1);
    }
}

// chunks of group g: [group_chunk_lo[g], group_chunk_lo[g + 1])
__global__ void relayout_scan_kernel(int* __restrict__ cnt, const int* __restrict__ group_chunk_lo, int n_genes, int R,
                                     long long* __restrict__ seg_len) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)n_genes * R) return;
    const int g = (int)(i / n_genes), gene = (int)(i % n_genes);       // gene fastest: coalesced over cnt rows
    int run = 0;
    for (int c = group_chunk_lo[g]; c < group_chunk_lo[g + 1]; ++c) {
        int* p = cnt + (long long)c * n_genes + gene;
        const int v = *p;
        *p = run;
        run += v;
    }
    seg_len[(long long)gene * R + g] = run;
}

__global__ void __launch_bounds__(kRelayoutThreads)
relayout_fill_kernel(RelayoutParams P, const long long* __restrict__ seg_ptr, float* __restrict__ vals_out,
                     int* __restrict__ rows_out) {
    const int lane = threadIdx.x & 31;
    const int chunk = blockIdx.x * (kRelayoutThreads / 32) + (threadIdx.x >> 5);
    if (chunk >= P.n_chunks) return;
    int* cur = P.cnt + (long long)chunk * P.n_genes;
    const int g = P.chunk_group[chunk];
    for (int r = P.chunk_row_lo[chunk]; r < P.chunk_row_lo[chunk + 1]; ++r) {
        const long long cell = P.order ? P.order[r] : r;
        const long long lo = P.indptr[cell], hi = P.indptr[cell + 1];
        for (long long e = lo + lane; e < hi; e += 32) {
            const int gene = ld_stream(P.indices + e);
            const int k = cur[gene];                 // only this warp touches this cursor, and a row has the gene once
            cur[gene] = k + 1;
            const long long pos = __ldg(seg_ptr + (long long)gene * P.R + g) + k;
            vals_out[pos] = ld_stream(P.data + e);
            rows_out[pos] = r;
        }
        __syncwarp();                                // the next row may hit the same genes: order the cursor updates
    }
}

// ------------------------------------------------------------------ tiled path (sorted column indices)
constexpr int kTileRows = 256;       // rows per chunk (bit-matrix height); chunks may be shorter
constexpr int kTileGenes = 256;      // genes per block (bit-matrix width)
constexpr int kTileThreads = 256;    // thread t <-> row t of the chunk (phases 1, 3), gene t of the block (phase 2)
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kStageCap = 6144;      // staged nonzeros per pass (48 KB); a denser block takes several passes

struct TileSmem {
    unsigned mask[kTileRows / 32][kTileGenes];          // bit (row & 31) of word [row >> 5][gene]
    unsigned short wpre[kTileRows / 32][kTileGenes];    // nonzeros of the gene in the lower row words
    int off[kTileGenes + 1];                            // exclusive scan of the block's per-gene counts
    int wsum[kTileWarps];
    float sval[kStageCap];
    int srow[kStageCap];
};

// grid = (n_chunks, n_split): CTA (c, s) handles the gene blocks [s * per, (s + 1) * per) of chunk c.
// kFill == false: per-chunk per-gene counts -> cnt[c][gene];  kFill == true: cnt holds the exclusive prefix over the
// group's chunks (relayout_scan_kernel) and the nonzeros are written to their final positions.
//
// Every THREAD walks its own row (round 2: one warp per row with 32-index loads was a chain of dependent loads per
// warp -- 16 rows x 2 loads x ~1 us per block -- and ran at 0.4 TB/s; a lane per row keeps 256 independent loads in
// flight per CTA, the 32-byte sectors a lane walks through stay in L1 between its consecutive 4-byte loads).
template <bool kFill>
__global__ void __launch_bounds__(kTileThreads)
relayout_tile_kernel(RelayoutParams P, int blocks_per_cta, const long long* __restrict__ seg_ptr,
                     float* __restrict__ vals_out, int* __restrict__ rows_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& S = *reinterpret_cast<TileSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int chunk = blockIdx.x;
    const int r0 = P.chunk_row_lo[chunk], n_rows = P.chunk_row_lo[chunk + 1] - r0;
    const int n_blocks = (P.n_genes + kTileGenes - 1) / kTileGenes;
    const int b_lo = blockIdx.y * blocks_per_cta, b_hi = min(n_blocks, b_lo + blocks_per_cta);
    if (b_lo >= b_hi || n_rows <= 0) return;
    const int grp = P.chunk_group[chunk];
    int* cnt = P.cnt + (long long)chunk * P.n_genes;

    // this thread's row: extent, and the first block's start by binary search (the column indices of a row ascend)
    long long cur = 0, row_hi = 0;
    if (tid < n_rows) {
        const long long cell = P.order ? P.order[r0 + tid] : r0 + tid;
        cur = P.indptr[cell];
        row_hi = P.indptr[cell + 1];
        if (b_lo > 0) {
            const int g0 = b_lo * kTileGenes;
            long long a = cur, b = row_hi;
            while (a < b) {
                const long long m = (a + b) >> 1;
                if (__ldg(P.indices + m) < g0) a = m + 1; else b = m;
            }
            cur = a;
        }
    }
    for (int i = tid; i < (kTileRows / 32) * kTileGenes; i += kTileThreads) (&S.mask[0][0])[i] = 0u;
    __syncthreads();
    const unsigned bit = 1u << (tid & 31), below = bit - 1u;
    const int w_row = tid >> 5;

    for (int b = b_lo; b < b_hi; ++b) {
        const int g0 = b * kTileGenes, g1 = min(P.n_genes, g0 + kTileGenes);
        // ---- phase 1: incidence bits of the block, one row per thread
        // (128-bit loads once the position is 16-byte aligned: a lane's 4-byte loads would fetch every 32-byte
        // sector eight times over)
        long long nxt = cur;
        {
            auto visit = [&](int idx) -> bool {             // false: the row has left the block
                if (idx >= g1) return false;
                if (idx >= g0) atomicOr(&S.mask[w_row][idx - g0], bit);
                else if (P.err) *P.err = 1;                 // an earlier block's gene after a later one: not sorted
                ++nxt;
                return true;
            };
            bool in = true;
            while (in && nxt < row_hi && (nxt & 3)) in = visit(__ldg(P.indices + nxt));
            while (in && nxt + 4 <= row_hi) {
                const int4 q = __ldg(reinterpret_cast<const int4*>(P.indices + nxt));
                in = visit(q.x) && visit(q.y) && visit(q.z) && visit(q.w);
            }
            while (in && nxt < row_hi) in = visit(__ldg(P.indices + nxt));
        }
        __syncthreads();
        // ---- phase 2: per-gene counts, word prefixes, exclusive scan over the block's genes (thread = gene)
        int c = 0;
#pragma unroll
        for (int w = 0; w < kTileRows / 32; ++w) {
            S.wpre[w][tid] = (unsigned short)c;
            c += __popc(S.mask[w][tid]);
        }
        if (!kFill) {
            if (g0 + tid < g1) cnt[g0 + tid] = c;
        } else {
            int incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(kFull, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) S.wsum[warp] = incl;
            __syncthreads();
            int base = 0;
            for (int w = 0; w < warp; ++w) base += S.wsum[w];
            S.off[tid] = base + incl - c;
            if (tid == kTileGenes - 1) S.off[kTileGenes] = base + incl;
            __syncthreads();
            const int n_t = S.off[kTileGenes];
            for (int win = 0; win < n_t; win += kStageCap) {
                // ---- phase 3: stage the block's nonzeros sorted by (gene, row), one row per thread
                {
                    auto place = [&](int idx, float v) {
                        const int j = idx - g0;
                        if (j < 0) return;                  // unsorted input: reported by the count pass
                        const int slot = S.off[j] + S.wpre[w_row][j] + __popc(S.mask[w_row][j] & below) - win;
                        if (slot >= 0 && slot < kStageCap) {
                            S.sval[slot] = v;
                            S.srow[slot] = r0 + tid;
                        }
                    };
                    long long e = cur;
                    for (; e < nxt && (e & 3); ++e) place(__ldg(P.indices + e), __ldg(P.data + e));
                    for (; e + 4 <= nxt; e += 4) {
                        const int4 q = __ldg(reinterpret_cast<const int4*>(P.indices + e));
                        const float4 v = __ldg(reinterpret_cast<const float4*>(P.data + e));
                        place(q.x, v.x); place(q.y, v.y); place(q.z, v.z); place(q.w, v.w);
                    }
                    for (; e < nxt; ++e) place(__ldg(P.indices + e), __ldg(P.data + e));
                }
                __syncthreads();
                // ---- phase 4: one coalesced run per gene
                for (int j = warp; j < g1 - g0; j += kTileWarps) {
                    const int lo = max(S.off[j], win), hi = min(S.off[j + 1], win + kStageCap);
                    if (lo >= hi) continue;
                    const long long base_o = __ldg(seg_ptr + (long long)(g0 + j) * P.R + grp) + cnt[g0 + j] - S.off[j];
                    for (int i = lo + lane; i < hi; i += 32) {
                        vals_out[base_o + i] = S.sval[i - win];
                        rows_out[base_o + i] = S.srow[i - win];
                    }
                }
                __syncthreads();
            }
        }
        // ---- next block
        __syncthreads();
#pragma unroll
        for (int w = 0; w < kTileRows / 32; ++w) S.mask[w][tid] = 0u;
        cur = nxt;
        __syncthreads();
    }
}

// counts must be non-negative integers below 2^24: the compression keys of csrc/unique.cu (count << 8 | bin) and
// csrc/pairs.cu hold the count in 24 bits and the moment kernels use the same values, so anything else would make
// the bootstrap tables and the moments disagree silently.  flags: bit 0 negative or NaN, bit 1 fractional, bit 2 >= 2^24.
__global__ void validate_counts_kernel(const float* __restrict__ data, long long nnz, int* __restrict__ flags) {
    int f = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nnz; i += (long long)gridDim.x * blockDim.x) {
        const float v = ld_stream(data + i);
        if (!(v >= 0.0f)) f |= 1;
        else if (v >= 16777216.0f) f |= 4;
        else if (v != floorf(v)) f |= 2;
    }
    f = __reduce_or_sync(kFull, f);
    if ((threadIdx.x & 31) == 0 && f) atomicOr(flags, f);
}

// canonical-form check of an uploaded CSR: column indices strictly ascending inside every row (sorted, no duplicates).
// scipy's own has_canonical_format is a single-threaded host scan of the index array (0.2 s at 6e8 nonzeros).
__global__ void csr_check_sorted_kernel(const long long* __restrict__ indptr, const int* __restrict__ indices,
                                        long long n_rows, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const long long lo = indptr[row], hi = indptr[row + 1];
    bool bad = false;
    for (long long e = lo + lane; e + 1 < hi; e += 32) bad |= ld_stream(indices + e) >= ld_stream(indices + e + 1);
    if (__any_sync(kFull, bad) && lane == 0) *flag = 1;
}

}  // namespace mm

using namespace mm;

// ------------------------------------------------------------------ host -> device upload of large pageable buffers
// cudaMemcpy from pageable memory runs at ~6 GB/s (one driver thread copies through its bounce buffers): the upload of
// a rank's 4.8 GB CSR shard was the largest term of the north-star pipeline.  Here n_threads host threads copy 8 MB
// chunks into a pinned ring (two slots per thread) and queue cudaMemcpyAsync on their own streams, so the host copies
// run in parallel and overlap the DMA.  The ring (n_threads x 16 MB of pinned memory, streams, events) is created on
// first use and kept until mm_upload_release(): the one piece of persistent state this library owns.
namespace {
constexpr int kUpMaxThreads = 8;
constexpr size_t kUpSlot = 8u << 20;
struct UploadRing {
    std::mutex mu;
    int device = -1;
    void* slot[kUpMaxThreads][2] = {};
    cudaEvent_t ev[kUpMaxThreads][2] = {};
    cudaEvent_t done[kUpMaxThreads] = {};
    cudaEvent_t start = nullptr;
    cudaStream_t st[kUpMaxThreads] = {};
};
UploadRing g_up;

void upload_release_locked() {
    if (g_up.device < 0) return;
    cudaSetDevice(g_up.device);
    for (int t = 0; t < kUpMaxThreads; ++t) {
        if (g_up.st[t]) cudaStreamSynchronize(g_up.st[t]);
        for (int b = 0; b < 2; ++b) {
            if (g_up.slot[t][b]) cudaFreeHost(g_up.slot[t][b]);
            if (g_up.ev[t][b]) cudaEventDestroy(g_up.ev[t][b]);
            g_up.slot[t][b] = nullptr; g_up.ev[t][b] = nullptr;
        }
        if (g_up.done[t]) cudaEventDestroy(g_up.done[t]);
        if (g_up.st[t]) cudaStreamDestroy(g_up.st[t]);
        g_up.done[t] = nullptr; g_up.st[t] = nullptr;
    }
    if (g_up.start) cudaEventDestroy(g_up.start);
    g_up.start = nullptr;
    g_up.device = -1;
}
}  // namespace

MM_EXPORT int mm_upload_release(void) {
    std::lock_guard<std::mutex> lock(g_up.mu);
    upload_release_locked();
    return 0;
}

MM_EXPORT int mm_upload(int device, void* stream, void* dst, const void* src, int64_t bytes, int32_t n_threads) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(bytes >= 0 && (bytes == 0 || (dst && src)), "dst/src/bytes");
    if (bytes == 0) return 0;
    cudaStream_t user = (cudaStream_t)stream;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > kUpMaxThreads) n_threads = kUpMaxThreads;
    const long long n_chunks = (bytes + (long long)kUpSlot - 1) / (long long)kUpSlot;
    if (n_chunks < 4) {          // small: the driver's own staged copy, ordered on the caller's stream
        MM_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, user));
        return 0;
    }
    if (n_threads > n_chunks) n_threads = (int)n_chunks;
    std::lock_guard<std::mutex> lock(g_up.mu);
    if (g_up.device != device) {
        upload_release_locked();
        MM_CUDA(cudaSetDevice(device));
        g_up.device = device;
        MM_CUDA(cudaEventCreateWithFlags(&g_up.start, cudaEventDisableTiming));
    }
    for (int t = 0; t < n_threads; ++t) {
        if (g_up.st[t]) continue;
        MM_CUDA(cudaStreamCreateWithFlags(&g_up.st[t], cudaStreamNonBlocking));
        MM_CUDA(cudaEventCreateWithFlags(&g_up.done[t], cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b) {
            MM_CUDA(cudaHostAlloc(&g_up.slot[t][b], kUpSlot, cudaHostAllocPortable));
            MM_CUDA(cudaEventCreateWithFlags(&g_up.ev[t][b], cudaEventDisableTiming));
        }
    }
    // the destination may be a freshly recycled block with work still queued on the caller's stream
    MM_CUDA(cudaEventRecord(g_up.start, user));
    std::vector<std::thread> workers;
    std::vector<cudaError_t> errs(n_threads, cudaSuccess);
    for (int t = 0; t < n_threads; ++t) {
        workers.emplace_back([&, t]() {
            cudaError_t e = cudaSetDevice(device);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(g_up.st[t], g_up.start, 0);
            for (long long c = t; c < n_chunks && e == cudaSuccess; c += n_threads) {
                const int b = (int)((c / n_threads) & 1);
                const size_t off = (size_t)c * kUpSlot;
                const size_t sz = (size_t)bytes - off < kUpSlot ? (size_t)bytes - off : kUpSlot;
                e = cudaEventSynchronize(g_up.ev[t][b]);                 // the slot's previous copy has left it
                if (e != cudaSuccess) break;
                memcpy(g_up.slot[t][b], (const char*)src + off, sz);
                e = cudaMemcpyAsync((char*)dst + off, g_up.slot[t][b], sz, cudaMemcpyHostToDevice, g_up.st[t]);
                if (e == cudaSuccess) e = cudaEventRecord(g_up.ev[t][b], g_up.st[t]);
            }
            errs[t] = e;
        });
    }
    for (auto& w : workers) w.join();
    for (int t = 0; t < n_threads; ++t) {
        if (errs[t] != cudaSuccess) {
            set_error("mm_upload worker %d: %s", t, cudaGetErrorString(errs[t]));
            return 2;
        }
        MM_CUDA(cudaEventRecord(g_up.done[t], g_up.st[t]));
        MM_CUDA(cudaStreamWaitEvent(user, g_up.done[t], 0));             // later work on the caller's stream sees the data
    }
    return 0;
}

MM_EXPORT int mm_csr_check_sorted(int device, void* stream, const int64_t* indptr, const int32_t* indices,
                                  int64_t n_rows, int32_t* flag) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_rows >= 0 && flag && (n_rows == 0 || (indptr && indices)), "indptr/indices/flag");
    MM_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (n_rows == 0) return 0;
    const long long blocks = (n_rows + 7) / 8;
    MM_REQUIRE(blocks < 2147483647LL, "too many rows");
    csr_check_sorted_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const long long*)indptr, indices, n_rows, flag);
    return check_launch("mm_csr_check_sorted");
}

static int tile_launch_shape(int n_chunks, int n_genes, int* blocks_per_cta, dim3* grid) {
    const int n_blocks = (n_genes + kTileGenes - 1) / kTileGenes;
    // enough CTAs for two waves of 3 per SM; a CTA walks at least one gene block
    int split = (148 * 6 + n_chunks - 1) / (n_chunks > 0 ? n_chunks : 1);
    if (split < 1) split = 1;
    if (split > n_blocks) split = n_blocks;
    if (split > 65535) split = 65535;
    *blocks_per_cta = (n_blocks + split - 1) / split;
    *grid = dim3((unsigned)n_chunks, (unsigned)((n_blocks + *blocks_per_cta - 1) / *blocks_per_cta));
    return 0;
}

MM_EXPORT int mm_validate_counts(int device, void* stream, const float* data, int64_t nnz, int32_t* flags) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(nnz >= 0 && flags && (data || nnz == 0), "data/flags");
    MM_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), (cudaStream_t)stream));
    if (nnz == 0) return 0;
    long long blocks = (nnz + 256 * 8 - 1) / (256 * 8);
    if (blocks > 148 * 16) blocks = 148 * 16;
    validate_counts_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(data, nnz, flags);
    return check_launch("mm_validate_counts");
}

MM_EXPORT int mm_relayout_count(int device, void* stream, const int64_t* indptr, const int32_t* indices,
                                const int32_t* order, const int32_t* chunk_row_lo, const int32_t* chunk_group,
                                const int32_t* group_chunk_lo, int32_t n_chunks, int32_t n_genes, int32_t R,
                                int32_t* cnt, int64_t* seg_len, int32_t sorted_rows, int32_t* err_flag) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_chunks >= 0 && n_genes > 0 && R > 0, "n_chunks/n_genes/R");
    MM_REQUIRE(indptr && chunk_row_lo && chunk_group && group_chunk_lo && cnt && seg_len, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    MM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)(n_chunks > 0 ? n_chunks : 1) * n_genes, st));
    RelayoutParams P;
    P.indptr = (const long long*)indptr; P.indices = indices; P.data = nullptr; P.order = order;
    P.chunk_row_lo = chunk_row_lo; P.chunk_group = chunk_group; P.n_chunks = n_chunks; P.n_genes = n_genes; P.R = R;
    P.cnt = cnt; P.err = err_flag;
    if (err_flag) MM_CUDA(cudaMemsetAsync(err_flag, 0, sizeof(int32_t), st));
    if (n_chunks > 0 && sorted_rows) {      // tiled path: chunks of at most kTileRows rows (checked by the fill call too)
        MM_REQUIRE(indices, "null pointer");
        MM_REQUIRE(((uintptr_t)indices & 15) == 0, "tiled path: indices must be 16-byte aligned");
        int per; dim3 grid;
        tile_launch_shape(n_chunks, n_genes, &per, &grid);
        MM_CUDA(cudaFuncSetAttribute(relayout_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(TileSmem)));
        relayout_tile_kernel<false><<<grid, kTileThreads, sizeof(TileSmem), st>>>(P, per, nullptr, nullptr, nullptr);
        if (int s = check_launch("relayout_tile<count>")) return s;
    } else if (n_chunks > 0) {
        MM_REQUIRE(indices, "null pointer");
        const int per = kRelayoutThreads / 32;
        relayout_count_kernel<<<(n_chunks + per - 1) / per, kRelayoutThreads, 0, st>>>(P);
        if (int s = check_launch("relayout_count")) return s;
    }
    const long long items = (long long)n_genes * R;
    relayout_scan_kernel<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(cnt, group_chunk_lo, n_genes, R,
                                                                         (long long*)seg_len);
    return check_launch("relayout_scan");
}

MM_EXPORT int mm_relayout_fill(int device, void* stream, const int64_t* indptr, const int32_t* indices,
                               const float* data, const int32_t* order, const int32_t* chunk_row_lo,
                               const int32_t* chunk_group, int32_t n_chunks, int32_t n_genes, int32_t R, int32_t* cnt,
                               const int64_t* seg_ptr, float* vals_out, int32_t* rows_out, int32_t sorted_rows) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_chunks >= 0 && n_genes > 0 && R > 0, "n_chunks/n_genes/R");
    if (n_chunks == 0) return 0;
    MM_REQUIRE(indptr && indices && data && chunk_row_lo && chunk_group && cnt && seg_ptr && vals_out && rows_out,
               "null pointer");
    RelayoutParams P;
    P.indptr = (const long long*)indptr; P.indices = indices; P.data = data; P.order = order;
    P.chunk_row_lo = chunk_row_lo; P.chunk_group = chunk_group; P.n_chunks = n_chunks; P.n_genes = n_genes; P.R = R;
    P.cnt = cnt; P.err = nullptr;
    if (sorted_rows) {
        MM_REQUIRE((((uintptr_t)indices | (uintptr_t)data) & 15) == 0, "tiled path: indices / data must be 16-byte aligned");
        int per; dim3 grid;
        tile_launch_shape(n_chunks, n_genes, &per, &grid);
        MM_CUDA(cudaFuncSetAttribute(relayout_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)sizeof(TileSmem)));
        relayout_tile_kernel<true><<<grid, kTileThreads, sizeof(TileSmem), (cudaStream_t)stream>>>(
            P, per, (const long long*)seg_ptr, vals_out, rows_out);
        return check_launch("relayout_tile<fill>");
    }
    const int per = kRelayoutThreads / 32;
    relayout_fill_kernel<<<(n_chunks + per - 1) / per, kRelayoutThreads, 0, (cudaStream_t)stream>>>(
        P, (const long long*)seg_ptr, vals_out, rows_out);
    return check_launch("relayout_fill");
}
