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
#include "common.cuh"

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

}  // namespace mm

using namespace mm;

MM_EXPORT int mm_relayout_count(int device, void* stream, const int64_t* indptr, const int32_t* indices,
                                const int32_t* order, const int32_t* chunk_row_lo, const int32_t* chunk_group,
                                const int32_t* group_chunk_lo, int32_t n_chunks, int32_t n_genes, int32_t R,
                                int32_t* cnt, int64_t* seg_len) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_chunks >= 0 && n_genes > 0 && R > 0, "n_chunks/n_genes/R");
    MM_REQUIRE(indptr && chunk_row_lo && chunk_group && group_chunk_lo && cnt && seg_len, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    MM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (size_t)(n_chunks > 0 ? n_chunks : 1) * n_genes, st));
    RelayoutParams P;
    P.indptr = (const long long*)indptr; P.indices = indices; P.data = nullptr; P.order = order;
    P.chunk_row_lo = chunk_row_lo; P.chunk_group = chunk_group; P.n_chunks = n_chunks; P.n_genes = n_genes; P.R = R;
    P.cnt = cnt;
    if (n_chunks > 0) {
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
                               const int64_t* seg_ptr, float* vals_out, int32_t* rows_out) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_chunks >= 0 && n_genes > 0 && R > 0, "n_chunks/n_genes/R");
    if (n_chunks == 0) return 0;
    MM_REQUIRE(indptr && indices && data && chunk_row_lo && chunk_group && cnt && seg_ptr && vals_out && rows_out,
               "null pointer");
    RelayoutParams P;
    P.indptr = (const long long*)indptr; P.indices = indices; P.data = data; P.order = order;
    P.chunk_row_lo = chunk_row_lo; P.chunk_group = chunk_group; P.n_chunks = n_chunks; P.n_genes = n_genes; P.R = R;
    P.cnt = cnt;
    const int per = kRelayoutThreads / 32;
    relayout_fill_kernel<<<(n_chunks + per - 1) / per, kRelayoutThreads, 0, (cudaStream_t)stream>>>(
        P, (const long long*)seg_ptr, vals_out, rows_out);
    return check_launch("relayout_fill");
}
