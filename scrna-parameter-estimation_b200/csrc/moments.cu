// HBM-bound segmented reductions over the sparse count matrix.
//
//  * mm_csr_row_sums  -- per-cell UMI totals, optionally restricted to a gene mask
//                        (reference estimator.py:64-76, the two size-factor passes of setup_memento)
//  * mm_seg_moments   -- per (gene, group) segment of the group-sorted CSC matrix:
//                        sum x, max x, sum x/sf, sum x/sf^2, sum x^2/sf^2
//                        (reference estimator.py:175-185 sparse form of _hyper_1d_relative, and the
//                        obs_mean / obs_max filters of main.py:199-207, fused into one pass)
//
// Layout: values float32, row ids int32, both streamed once with 128-bit L1-bypassing loads;
// per-cell float64 1/size_factor gathered through L1/L2 (rows inside a segment are ascending and
// confined to the group's contiguous row range, so the gather footprint is small).
// Accumulation is float64.  One warp per segment; segments longer than kBigSeg nnz are deferred
// to a CTA-per-segment kernel through a device-side list (no host round trip).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace mm {

constexpr int kBigSeg = 32768;  // largest nnz a warp-level group handles; longer segments use a whole CTA
constexpr int kCtaThreads = 256;

// Calls f(value, index) for every element of [lo, hi) using `nthr` cooperating threads
// (thread id `t`): scalar head up to 16-byte alignment, float4/int4 body, scalar tail.
template <int UNROLL = 2, class F>
__device__ __forceinline__ void stream_pairs(const float* __restrict__ vals, const int* __restrict__ idx,
                                             long long lo, long long hi, int t, int nthr, F f) {
    long long body = (lo + 3) & ~3LL;
    if (body > hi) body = hi;
    for (long long i = lo + t; i < body; i += nthr) f(ld_stream(vals + i), ld_stream(idx + i));
    long long nvec = (hi - body) >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vals + body);
    const int4* i4 = reinterpret_cast<const int4*>(idx + body);
    long long k = t;
    if (UNROLL == 4) {   // four independent 128-bit load pairs in flight per thread
        for (; k + 3LL * nthr < nvec; k += 4LL * nthr) {
            float4 a = ld_stream4(v4 + k);
            int4 ai = ld_stream4(i4 + k);
            float4 b = ld_stream4(v4 + k + nthr);
            int4 bi = ld_stream4(i4 + k + nthr);
            float4 c = ld_stream4(v4 + k + 2LL * nthr);
            int4 ci = ld_stream4(i4 + k + 2LL * nthr);
            float4 d = ld_stream4(v4 + k + 3LL * nthr);
            int4 di = ld_stream4(i4 + k + 3LL * nthr);
            f(a.x, ai.x); f(a.y, ai.y); f(a.z, ai.z); f(a.w, ai.w);
            f(b.x, bi.x); f(b.y, bi.y); f(b.z, bi.z); f(b.w, bi.w);
            f(c.x, ci.x); f(c.y, ci.y); f(c.z, ci.z); f(c.w, ci.w);
            f(d.x, di.x); f(d.y, di.y); f(d.z, di.z); f(d.w, di.w);
        }
    }
    // two independent 128-bit load pairs in flight per thread
    for (; k + nthr < nvec; k += 2LL * nthr) {
        float4 a = ld_stream4(v4 + k);
        int4 ai = ld_stream4(i4 + k);
        float4 b = ld_stream4(v4 + k + nthr);
        int4 bi = ld_stream4(i4 + k + nthr);
        f(a.x, ai.x); f(a.y, ai.y); f(a.z, ai.z); f(a.w, ai.w);
        f(b.x, bi.x); f(b.y, bi.y); f(b.z, bi.z); f(b.w, bi.w);
    }
    for (; k < nvec; k += nthr) {
        float4 a = ld_stream4(v4 + k);
        int4 ai = ld_stream4(i4 + k);
        f(a.x, ai.x); f(a.y, ai.y); f(a.z, ai.z); f(a.w, ai.w);
    }
    for (long long i = body + (nvec << 2) + t; i < hi; i += nthr) f(ld_stream(vals + i), ld_stream(idx + i));
}

// Software-pipelined variant: the loads of the next DEPTH vector iterations are in flight while the
// current one is processed (gather + float64 math), which keeps HBM requests outstanding during the
// compute phase of short segments.
template <int DEPTH, class F>
__device__ __forceinline__ void stream_pairs_pipe(const float* __restrict__ vals, const int* __restrict__ idx,
                                                  long long lo, long long hi, int t, int nthr, F f) {
    long long body = (lo + 3) & ~3LL;
    if (body > hi) body = hi;
    for (long long i = lo + t; i < body; i += nthr) f(ld_stream(vals + i), ld_stream(idx + i));
    const long long nvec = (hi - body) >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vals + body);
    const int4* i4 = reinterpret_cast<const int4*>(idx + body);
    float4 bv[DEPTH];
    int4 bi[DEPTH];
    long long k = t;
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
        long long kk = k + (long long)d * nthr;
        if (kk < nvec) { bv[d] = ld_stream4(v4 + kk); bi[d] = ld_stream4(i4 + kk); }
    }
    while (k < nvec) {
        float4 a = bv[0];
        int4 ai = bi[0];
#pragma unroll
        for (int d = 0; d + 1 < DEPTH; ++d) { bv[d] = bv[d + 1]; bi[d] = bi[d + 1]; }
        long long kn = k + (long long)DEPTH * nthr;
        if (kn < nvec) { bv[DEPTH - 1] = ld_stream4(v4 + kn); bi[DEPTH - 1] = ld_stream4(i4 + kn); }
        f(a.x, ai.x); f(a.y, ai.y); f(a.z, ai.z); f(a.w, ai.w);
        k += nthr;
    }
    for (long long i = body + (nvec << 2) + t; i < hi; i += nthr) f(ld_stream(vals + i), ld_stream(idx + i));
}

// ------------------------------------------------------------------ CSR row sums
__global__ void __launch_bounds__(kCtaThreads)
csr_row_sums_kernel(const long long* __restrict__ indptr, const int* __restrict__ indices,
                    const float* __restrict__ data, long long n_rows,
                    const unsigned char* __restrict__ gene_mask, double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    long long row = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    long long lo = indptr[row], hi = indptr[row + 1];
    double acc = 0.0;
    if (gene_mask == nullptr) {
        stream_pairs(data, indices, lo, hi, lane, 32, [&](float v, int) { acc += (double)v; });
    } else {
        stream_pairs(data, indices, lo, hi, lane, 32,
                     [&](float v, int c) { acc += __ldg(gene_mask + c) ? (double)v : 0.0; });
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
}

// ------------------------------------------------------------------ segment moments
struct Mom {
    double sx = 0, s1 = 0, s2 = 0, s3 = 0;
    float mx = 0.f;
    __device__ __forceinline__ void add(float v, double w) {
        double x = (double)v;
        double xw = x * w;
        double xw2 = xw * w;
        sx += x;
        s1 += xw;
        s2 += xw2;
        s3 = fma(x, xw2, s3);
        mx = fmaxf(mx, v);
    }
};

// W lanes cooperate on one segment (W = 8, 16 or 32; 32 / W segments per warp): short segments keep
// every lane busy and need only log2(W) shuffle levels, while each group still reads whole 128-byte
// lines (8 lanes x 16 B).  Segments above `big_thresh` nonzeros go to the CTA kernel through big_list.
template <int W, bool kNoGather = false, int UNROLL = 2>
__global__ void __launch_bounds__(kCtaThreads)
seg_moments_group_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                         const long long* __restrict__ seg_ptr, long long n_seg,
                         const double* __restrict__ inv_sf, double* __restrict__ out,
                         int* __restrict__ big_list, int big_thresh) {
    constexpr int kGroups = 32 / W;
    const int lane = threadIdx.x & 31;
    const int sub = lane % W;
    const long long warp_id = (long long)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const long long seg = warp_id * kGroups + lane / W;
    const bool active = seg < n_seg;
    long long lo = 0, hi = 0;
    if (active) { lo = __ldg(seg_ptr + seg); hi = __ldg(seg_ptr + seg + 1); }
    const bool big = (hi - lo > big_thresh);
    if (big) {
        if (sub == 0) big_list[1 + atomicAdd(big_list, 1)] = (int)seg;
        hi = lo;
    }
    Mom m;
    if constexpr (kNoGather)   // tuning experiment only (MM_MOMENTS_NOGATHER): wrong results, streaming ceiling
        stream_pairs(vals, rows, lo, hi, sub, W, [&](float v, int r) { m.add(v, (double)(r & 7)); });
    else if constexpr (UNROLL >= 10)   // UNROLL = 10 + depth selects the software-pipelined streamer
        stream_pairs_pipe<UNROLL - 10>(vals, rows, lo, hi, sub, W, [&](float v, int r) { m.add(v, __ldg(inv_sf + r)); });
    else
        stream_pairs<UNROLL>(vals, rows, lo, hi, sub, W, [&](float v, int r) { m.add(v, __ldg(inv_sf + r)); });
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        m.sx += __shfl_xor_sync(kFull, m.sx, o);
        m.s1 += __shfl_xor_sync(kFull, m.s1, o);
        m.s2 += __shfl_xor_sync(kFull, m.s2, o);
        m.s3 += __shfl_xor_sync(kFull, m.s3, o);
        m.mx = fmaxf(m.mx, __shfl_xor_sync(kFull, m.mx, o));
    }
    if (active && !big && sub == 0) {
        out[seg] = m.sx;
        out[n_seg + seg] = (double)m.mx;
        out[2 * n_seg + seg] = m.s1;
        out[3 * n_seg + seg] = m.s2;
        out[4 * n_seg + seg] = m.s3;
    }
}

// ------------------------------------------------------------------ flat streaming variant
// The nonzero arrays are read as ONE contiguous stream: every warp owns kFlatWarpElems consecutive
// nonzeros (whole 512-byte rows of 128-bit loads, no per-segment head/tail, no per-segment latency
// chain) and walks the segment boundaries as it goes.  Lanes keep private float64 partials while the
// warp stays inside one segment; at a boundary the warp reduces them with shuffles and one lane adds
// the segment's partial to `out` (float64 atomics; a segment has one partial per warp it spans, so
// almost all segments have a single, deterministic writer).  `out` must be zero-initialised.
constexpr int kFlatWarpElems = 4096;
constexpr int kFlatThreads = 128;

__device__ __forceinline__ void mom_flush(Mom& m, long long seg, long long n_seg, double* __restrict__ out, int lane) {
    m.sx = warp_sum(m.sx); m.s1 = warp_sum(m.s1); m.s2 = warp_sum(m.s2); m.s3 = warp_sum(m.s3);
    m.mx = warp_max(m.mx);
    if (lane == 0) {
        if (m.sx != 0.0) atomicAdd(out + seg, m.sx);
        if (m.mx > 0.f)
            atomicMax(reinterpret_cast<unsigned long long*>(out + n_seg + seg),
                      (unsigned long long)__double_as_longlong((double)m.mx));
        if (m.s1 != 0.0) atomicAdd(out + 2 * n_seg + seg, m.s1);
        if (m.s2 != 0.0) atomicAdd(out + 3 * n_seg + seg, m.s2);
        if (m.s3 != 0.0) atomicAdd(out + 4 * n_seg + seg, m.s3);
    }
    m = Mom();
}

__global__ void __launch_bounds__(kFlatThreads)
seg_moments_flat_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                        const long long* __restrict__ seg_ptr, long long n_seg, long long nnz,
                        const int* __restrict__ chunk_seg, const double* __restrict__ inv_sf,
                        double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long wchunk = (long long)blockIdx.x * (kFlatThreads / 32) + (threadIdx.x >> 5);
    long long p = wchunk * kFlatWarpElems;               // multiple of 4: 16-byte aligned vector loads
    if (p >= nnz) return;
    long long p_end = p + kFlatWarpElems;
    if (p_end > nnz) p_end = nnz;
    long long cur = chunk_seg[wchunk];                   // segment containing p
    long long next_bnd = __ldg(seg_ptr + cur + 1);       // first index that is no longer in `cur`
    Mom m;
    const float4* v4 = reinterpret_cast<const float4*>(vals);
    const int4* i4 = reinterpret_cast<const int4*>(rows);
    for (; p < p_end; p += 256) {
        // two 128-element rows per trip: issue all streaming loads, then the gathers, then the math
        const long long q0 = p + 4 * lane, q1 = q0 + 128;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        int4 ai = make_int4(0, 0, 0, 0), bi = ai;
        const bool fa = q0 + 3 < p_end, fb = q1 + 3 < p_end;       // full vectors (p_end == nnz may be ragged)
        if (fa) { a = ld_stream4(v4 + (q0 >> 2)); ai = ld_stream4(i4 + (q0 >> 2)); }
        if (fb) { b = ld_stream4(v4 + (q1 >> 2)); bi = ld_stream4(i4 + (q1 >> 2)); }
        if (!fa) {   // ragged tail of the whole array: scalar loads
            if (q0 < p_end) { a.x = vals[q0]; ai.x = rows[q0]; }
            if (q0 + 1 < p_end) { a.y = vals[q0 + 1]; ai.y = rows[q0 + 1]; }
            if (q0 + 2 < p_end) { a.z = vals[q0 + 2]; ai.z = rows[q0 + 2]; }
        }
        if (!fb) {
            if (q1 < p_end) { b.x = vals[q1]; bi.x = rows[q1]; }
            if (q1 + 1 < p_end) { b.y = vals[q1 + 1]; bi.y = rows[q1 + 1]; }
            if (q1 + 2 < p_end) { b.z = vals[q1 + 2]; bi.z = rows[q1 + 2]; }
        }
        const float xv[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const int rv[8] = {ai.x, ai.y, ai.z, ai.w, bi.x, bi.y, bi.z, bi.w};
        double wv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const long long q = (j < 4 ? q0 : q1 - 4) + j;
            wv[j] = (q < p_end) ? __ldg(inv_sf + rv[j]) : 0.0;     // padded elements have x = 0 anyway
        }
        long long trip_end = p + 256;
        if (trip_end > p_end) trip_end = p_end;
        if (p >= next_bnd) {                                        // the previous trip ended exactly on a boundary
            mom_flush(m, cur, n_seg, out, lane);
            do { ++cur; next_bnd = __ldg(seg_ptr + cur + 1); } while (next_bnd <= p);
        }
        if (trip_end <= next_bnd) {                                 // whole trip inside the current segment
#pragma unroll
            for (int j = 0; j < 8; ++j) m.add(xv[j], wv[j]);
        } else {
            long long from = p;                                     // elements in [from, next_bnd) belong to `cur`
            while (true) {
                const long long upto = next_bnd < trip_end ? next_bnd : trip_end;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const long long q = (j < 4 ? q0 : q1 - 4) + j;
                    if (q >= from && q < upto) m.add(xv[j], wv[j]);
                }
                if (next_bnd >= trip_end) break;                    // `cur` reaches (at least) the end of the trip
                mom_flush(m, cur, n_seg, out, lane);                // boundary strictly inside the trip
                from = next_bnd;
                do { ++cur; next_bnd = __ldg(seg_ptr + cur + 1); } while (next_bnd <= from);
            }
        }
    }
    mom_flush(m, cur, n_seg, out, lane);
}

// Same reduction with the group's 1/size_factor window staged in shared memory: a block owns one group
// and a range of genes, so the per-nonzero gather (the L1 wavefront bottleneck of the kernel above:
// ~1 sector per nonzero) becomes a shared-memory read.  Used when the largest group fits.
template <int W>
__global__ void __launch_bounds__(kCtaThreads)
seg_moments_smem_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                        const long long* __restrict__ seg_ptr, long long n_genes, int R,
                        const long long* __restrict__ group_start, const double* __restrict__ inv_sf,
                        double* __restrict__ out, int* __restrict__ big_list, int big_thresh, int genes_per_block) {
    extern __shared__ double s_w[];
    constexpr int kGroups = 32 / W;
    const int r = blockIdx.y;
    const long long base = group_start[r];
    const int ncell = (int)(group_start[r + 1] - base);
    for (int i = threadIdx.x; i < ncell; i += kCtaThreads) s_w[i] = inv_sf[base + i];
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane % W;
    const long long n_seg = n_genes * R;
    const long long g_lo = (long long)blockIdx.x * genes_per_block;
    long long g_hi = g_lo + genes_per_block;
    if (g_hi > n_genes) g_hi = n_genes;
    const int slot = (threadIdx.x >> 5) * kGroups + lane / W;          // sub-group index inside the block
    for (long long g0 = g_lo; g0 < g_hi; g0 += (kCtaThreads / 32) * kGroups) {   // uniform trip count per block
        const long long g = g0 + slot;
        const bool active = g < g_hi;
        const long long seg = g * R + r;
        long long lo = 0, hi = 0;
        if (active) { lo = __ldg(seg_ptr + seg); hi = __ldg(seg_ptr + seg + 1); }
        const bool big = (hi - lo > big_thresh);
        if (big) {
            if (sub == 0) big_list[1 + atomicAdd(big_list, 1)] = (int)seg;
            hi = lo;
        }
        Mom m;
        stream_pairs(vals, rows, lo, hi, sub, W, [&](float v, int row) { m.add(v, s_w[row - (int)base]); });
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) {
            m.sx += __shfl_xor_sync(kFull, m.sx, o);
            m.s1 += __shfl_xor_sync(kFull, m.s1, o);
            m.s2 += __shfl_xor_sync(kFull, m.s2, o);
            m.s3 += __shfl_xor_sync(kFull, m.s3, o);
            m.mx = fmaxf(m.mx, __shfl_xor_sync(kFull, m.mx, o));
        }
        if (active && !big && sub == 0) {
            out[seg] = m.sx;
            out[n_seg + seg] = (double)m.mx;
            out[2 * n_seg + seg] = m.s1;
            out[3 * n_seg + seg] = m.s2;
            out[4 * n_seg + seg] = m.s3;
        }
    }
}

__global__ void __launch_bounds__(kCtaThreads)
seg_moments_cta_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                       const long long* __restrict__ seg_ptr, long long n_seg,
                       const double* __restrict__ inv_sf, double* __restrict__ out,
                       const int* __restrict__ big_list) {
    __shared__ double red[kCtaThreads / 32][4];
    __shared__ float redmax[kCtaThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_big = big_list[0];
    for (int b = blockIdx.x; b < n_big; b += gridDim.x) {
        long long seg = big_list[1 + b];
        long long lo = seg_ptr[seg], hi = seg_ptr[seg + 1];
        Mom m;
        stream_pairs(vals, rows, lo, hi, threadIdx.x, kCtaThreads,
                     [&](float v, int r) { m.add(v, __ldg(inv_sf + r)); });
        m.sx = warp_sum(m.sx); m.s1 = warp_sum(m.s1); m.s2 = warp_sum(m.s2); m.s3 = warp_sum(m.s3);
        m.mx = warp_max(m.mx);
        if (lane == 0) {
            red[warp][0] = m.sx; red[warp][1] = m.s1; red[warp][2] = m.s2; red[warp][3] = m.s3;
            redmax[warp] = m.mx;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0, b1 = 0, c = 0, d = 0;
            float mx = 0.f;
            for (int w = 0; w < kCtaThreads / 32; ++w) {  // fixed order: deterministic
                a += red[w][0]; b1 += red[w][1]; c += red[w][2]; d += red[w][3];
                mx = fmaxf(mx, redmax[w]);
            }
            out[seg] = a;
            out[n_seg + seg] = (double)mx;
            out[2 * n_seg + seg] = b1;
            out[3 * n_seg + seg] = c;
            out[4 * n_seg + seg] = d;
        }
        __syncthreads();
    }
}


// ------------------------------------------------------------------ pair products (2D moments)
// One warp per (pair, group): lanes walk the shorter of the two segments and binary-search each
// row id in the longer one (both row lists are ascending).
__global__ void __launch_bounds__(kCtaThreads)
pair_products_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                     const long long* __restrict__ seg_ptr, int R, const int* __restrict__ idx1,
                     const int* __restrict__ idx2, long long n_items, const double* __restrict__ inv_sf,
                     double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    long long item = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
    if (item >= n_items) return;
    long long k = item / R;
    int r = (int)(item % R);
    long long sa = (long long)idx1[k] * R + r, sb = (long long)idx2[k] * R + r;
    long long alo = seg_ptr[sa], ahi = seg_ptr[sa + 1], blo = seg_ptr[sb], bhi = seg_ptr[sb + 1];
    if (ahi - alo > bhi - blo) { long long t = alo; alo = blo; blo = t; t = ahi; ahi = bhi; bhi = t; }
    double acc = 0.0;
    for (long long i = alo + lane; i < ahi; i += 32) {
        int row = rows[i];
        long long lo = blo, hi = bhi;
        while (lo < hi) {
            long long mid = (lo + hi) >> 1;
            if (__ldg(rows + mid) < row) lo = mid + 1; else hi = mid;
        }
        if (lo < bhi && __ldg(rows + lo) == row) {
            double w = __ldg(inv_sf + row);
            acc += (double)vals[i] * (double)__ldg(vals + lo) * w * w;
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[item] = acc;
}

// ------------------------------------------------------------------ host-side error state
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}
int enter(int device) {
    g_err[0] = 0;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}
const char* last_error() { return g_err; }

}  // namespace mm

using namespace mm;

MM_EXPORT const char* mm_last_error(void) { return mm::last_error(); }
MM_EXPORT int mm_version(void) { return 100; }

MM_EXPORT int mm_csr_row_sums(int device, void* stream, const int64_t* indptr, const int32_t* indices,
                              const float* data, int64_t n_rows, const uint8_t* gene_mask, double* out) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_rows >= 0, "n_rows");
    if (n_rows == 0) return 0;
    MM_REQUIRE(indptr && out, "null pointer");
    long long blocks = (n_rows + (kCtaThreads / 32) - 1) / (kCtaThreads / 32);
    csr_row_sums_kernel<<<(unsigned)blocks, kCtaThreads, 0, (cudaStream_t)stream>>>(
        (const long long*)indptr, indices, data, n_rows, gene_mask, out);
    return check_launch("mm_csr_row_sums");
}

template <int W>
static void launch_smem(dim3 grid, size_t smem, cudaStream_t st, const float* vals, const int32_t* rows,
                        const long long* sp, long long n_genes, int R, const long long* gs, const double* inv_sf,
                        double* out, int* big_list, int big_thresh, int gpb) {
    cudaFuncSetAttribute(seg_moments_smem_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    seg_moments_smem_kernel<W><<<grid, kCtaThreads, smem, st>>>(vals, rows, sp, n_genes, R, gs, inv_sf, out, big_list,
                                                               big_thresh, gpb);
}

MM_EXPORT int mm_seg_moments(int device, void* stream, const float* vals, const int32_t* rows,
                             const int64_t* seg_ptr, int64_t n_seg, int64_t nnz, const double* inv_sf,
                             double* out, int32_t* big_list, const int64_t* group_start, int32_t R,
                             int64_t max_group_cells, const int32_t* chunk_seg) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_seg >= 0, "n_seg");
    if (n_seg == 0) return 0;
    MM_REQUIRE(seg_ptr && inv_sf && out && big_list, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    // flat streaming kernel (default when the caller supplies the per-warp-chunk segment index and the
    // segments are not tiny): one contiguous pass, boundaries walked on the fly
    if (chunk_seg && nnz > 0 && nnz / n_seg >= 48 && !getenv("MM_MOMENTS_NOFLAT")) {
        MM_CUDA(cudaMemsetAsync(out, 0, sizeof(double) * 5 * (size_t)n_seg, st));
        long long wchunks = (nnz + kFlatWarpElems - 1) / kFlatWarpElems;
        long long nb = (wchunks + (kFlatThreads / 32) - 1) / (kFlatThreads / 32);
        MM_REQUIRE(nb < 2147483647LL, "matrix too large for one launch");
        seg_moments_flat_kernel<<<(unsigned)nb, kFlatThreads, 0, st>>>(vals, rows, (const long long*)seg_ptr, n_seg, nnz,
                                                                      chunk_seg, inv_sf, out);
        return check_launch("seg_moments_flat");
    }
    MM_CUDA(cudaMemsetAsync(big_list, 0, sizeof(int32_t), st));
    // group width by mean segment length; segments far above the mean (or long enough that a single
    // warp would be the tail of the launch) are deferred to the CTA kernel
    const long long mean_len = nnz / n_seg;
    int W = mean_len < 512 ? 8 : (mean_len < 2048 ? 16 : 32);
    if (const char* ov = getenv("MM_MOMENTS_W")) { int w = atoi(ov); if (w == 8 || w == 16 || w == 32) W = w; }   // tuning hook
    long long thr = nnz / (148LL * 64);
    const int big_thresh = (int)(thr < 4096 ? 4096 : (thr > kBigSeg ? kBigSeg : thr));
    const long long segs_per_block = (kCtaThreads / 32) * (32 / W);
    long long blocks = (n_seg + segs_per_block - 1) / segs_per_block;
    MM_REQUIRE(blocks < 2147483647LL, "too many segments for one launch");
    const long long* sp = (const long long*)seg_ptr;
    const bool no_smem = getenv("MM_MOMENTS_NOSMEM") != nullptr;     // tuning hook
    if (group_start && R > 0 && R <= 65535 && max_group_cells > 0 && max_group_cells * 8 <= 96 * 1024 &&
        n_seg % R == 0 && !no_smem) {
        const long long n_genes = n_seg / R;
        const int per_iter = (kCtaThreads / 32) * (32 / W);
        long long gpb = (n_seg / 2400) / per_iter * per_iter;            // ~2400 blocks in total
        if (gpb < per_iter) gpb = per_iter;
        if (gpb > 4096) gpb = 4096;
        dim3 grid((unsigned)((n_genes + gpb - 1) / gpb), (unsigned)R);
        size_t smem = (size_t)max_group_cells * 8;
        const long long* gs = (const long long*)group_start;
        if (W == 8) launch_smem<8>(grid, smem, st, vals, rows, sp, n_genes, R, gs, inv_sf, out, big_list, big_thresh, (int)gpb);
        else if (W == 16) launch_smem<16>(grid, smem, st, vals, rows, sp, n_genes, R, gs, inv_sf, out, big_list, big_thresh, (int)gpb);
        else launch_smem<32>(grid, smem, st, vals, rows, sp, n_genes, R, gs, inv_sf, out, big_list, big_thresh, (int)gpb);
    } else if (getenv("MM_MOMENTS_NOGATHER"))
        seg_moments_group_kernel<16, true><<<(unsigned)blocks, kCtaThreads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
    else if (const char* v = getenv("MM_MOMENTS_VARIANT")) {     // tuning hook: threads per block / unroll
        int variant = atoi(v);
        int threads = 256;
        long long spb8 = (threads / 32) * 4, spb16 = (threads / 32) * 2;
        unsigned nb8 = (unsigned)((n_seg + spb8 - 1) / spb8), nb16 = (unsigned)((n_seg + spb16 - 1) / spb16);
        if (variant == 0) seg_moments_group_kernel<8, false, 12><<<nb8, threads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
        else if (variant == 1) seg_moments_group_kernel<8, false, 13><<<nb8, threads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
        else if (variant == 2) seg_moments_group_kernel<16, false, 12><<<nb16, threads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
        else if (variant == 3) seg_moments_group_kernel<16, false, 13><<<nb16, threads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
        else seg_moments_group_kernel<8, false, 14><<<nb8, threads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
    }
    else if (W == 8)
        seg_moments_group_kernel<8><<<(unsigned)blocks, kCtaThreads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
    else if (W == 16)
        seg_moments_group_kernel<16><<<(unsigned)blocks, kCtaThreads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
    else
        seg_moments_group_kernel<32><<<(unsigned)blocks, kCtaThreads, 0, st>>>(vals, rows, sp, n_seg, inv_sf, out, big_list, big_thresh);
    if (int s = check_launch("seg_moments_group")) return s;
    seg_moments_cta_kernel<<<148 * 4, kCtaThreads, 0, st>>>(
        vals, rows, (const long long*)seg_ptr, n_seg, inv_sf, out, big_list);
    return check_launch("seg_moments_cta");
}

MM_EXPORT int mm_pair_products(int device, void* stream, const float* vals, const int32_t* rows,
                               const int64_t* seg_ptr, int32_t R, const int32_t* idx1, const int32_t* idx2,
                               int64_t n_pairs, const double* inv_sf, double* out) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_pairs >= 0 && R > 0, "n_pairs/R");
    if (n_pairs == 0) return 0;
    MM_REQUIRE(vals && rows && seg_ptr && idx1 && idx2 && inv_sf && out, "null pointer");
    long long items = n_pairs * (long long)R;
    long long blocks = (items + (kCtaThreads / 32) - 1) / (kCtaThreads / 32);
    MM_REQUIRE(blocks < 2147483647LL, "too many pairs for one launch");
    pair_products_kernel<<<(unsigned)blocks, kCtaThreads, 0, (cudaStream_t)stream>>>(
        vals, rows, (const long long*)seg_ptr, R, idx1, idx2, items, inv_sf, out);
    return check_launch("mm_pair_products");
}
