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

namespace mm {

constexpr int kBigSeg = 32768;  // largest nnz a warp-level group handles; longer segments use a whole CTA
constexpr int kCtaThreads = 256;

// Calls f(value, index) for every element of [lo, hi) using `nthr` cooperating threads
// (thread id `t`): scalar head up to 16-byte alignment, float4/int4 body, scalar tail.
template <class F>
__device__ __forceinline__ void stream_pairs(const float* __restrict__ vals, const int* __restrict__ idx,
                                             long long lo, long long hi, int t, int nthr, F f) {
    long long body = (lo + 3) & ~3LL;
    if (body > hi) body = hi;
    for (long long i = lo + t; i < body; i += nthr) f(ld_stream(vals + i), ld_stream(idx + i));
    long long nvec = (hi - body) >> 2;
    const float4* v4 = reinterpret_cast<const float4*>(vals + body);
    const int4* i4 = reinterpret_cast<const int4*>(idx + body);
    long long k = t;
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
template <int W>
__global__ void __launch_bounds__(kCtaThreads)
seg_moments_group_kernel(const float* __restrict__ vals, const int* __restrict__ rows,
                         const long long* __restrict__ seg_ptr, long long n_seg,
                         const double* __restrict__ inv_sf, double* __restrict__ out,
                         int* __restrict__ big_list, int big_thresh) {
    constexpr int kGroups = 32 / W;
    const int lane = threadIdx.x & 31;
    const int sub = lane % W;
    const long long warp_id = (long long)blockIdx.x * (kCtaThreads / 32) + (threadIdx.x >> 5);
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
    stream_pairs(vals, rows, lo, hi, sub, W, [&](float v, int r) { m.add(v, __ldg(inv_sf + r)); });
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

MM_EXPORT int mm_seg_moments(int device, void* stream, const float* vals, const int32_t* rows,
                             const int64_t* seg_ptr, int64_t n_seg, int64_t nnz, const double* inv_sf,
                             double* out, int32_t* big_list) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_seg >= 0, "n_seg");
    if (n_seg == 0) return 0;
    MM_REQUIRE(seg_ptr && inv_sf && out && big_list, "null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    MM_CUDA(cudaMemsetAsync(big_list, 0, sizeof(int32_t), st));
    // group width by mean segment length; segments far above the mean (or long enough that a single
    // warp would be the tail of the launch) are deferred to the CTA kernel
    const long long mean_len = nnz / n_seg;
    const int W = mean_len < 512 ? 8 : (mean_len < 2048 ? 16 : 32);
    long long thr = nnz / (148LL * 64);
    const int big_thresh = (int)(thr < 4096 ? 4096 : (thr > kBigSeg ? kBigSeg : thr));
    const long long segs_per_block = (kCtaThreads / 32) * (32 / W);
    long long blocks = (n_seg + segs_per_block - 1) / segs_per_block;
    MM_REQUIRE(blocks < 2147483647LL, "too many segments for one launch");
    const long long* sp = (const long long*)seg_ptr;
    if (W == 8)
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
