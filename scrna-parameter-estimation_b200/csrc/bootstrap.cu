// Fused compressed bootstrap: multinomial resampling of each (gene, group) unique-value table and
// the bootstrapped moments, one thread per (segment, replicate).
//
// Replaces reference bootstrap.py:74-116 (_bootstrap_1d: Generator(PCG64(5)).multinomial into a
// U x B int64 matrix, then the tuple-form estimator estimator.py:171-174 over U x B temporaries)
// and the per-replicate residual variance (hypothesis_test.py:186 -> estimator.py:103-111).
//
// The multinomial is drawn exactly, as a chain of conditional binomials over the U nonzero
// categories (the zero-count category is the remainder and contributes nothing, so it is never
// drawn).  Binomial sampler: sequential inversion when the expected count is small, Hormann's BTRS
// transformed rejection otherwise (log-pmf ratio evaluated in a cancellation-free log1p form so
// that float32 is enough at n ~ 1e6).  Randomness: Philox4x32-10, counter = (replicate, segment,
// block, stream tag), key = seed: any replicate of any segment can be regenerated independently.
// Moments accumulate in float64.  No U x B matrix is ever materialised.
//
// mm_bootstrap_1d_replay evaluates the same moments from host-supplied resample counts (the
// deterministic parity mode: must match the reference's statistics to 1e-6).
#include "common.cuh"

namespace mm {

struct __align__(32) BootEntry {   // must match unique.cu
    double a, b;
    float p, lq;
    int n, mode;
};

constexpr int kBootThreads = 128;

__device__ __forceinline__ float stirling_tail(float k) {
    // log(k!) - [(k + 1/2) log(k + 1) - (k + 1) + 1/2 log(2 pi)]
    float kp1 = k + 1.f;
    float kp1sq = kp1 * kp1;
    float series = (0.08333333333f - (0.00277777778f - 0.00079365079f / kp1sq) / kp1sq) / kp1;
    if (k < 10.f) {
        const float tab[10] = {0.0810614667953272f, 0.0413406959554092f, 0.0276779256849983f,
                               0.02079067210376509f, 0.0166446911898211f, 0.0138761288230707f,
                               0.0118967099458917f, 0.0104112652619720f, 0.00925546218271273f,
                               0.00833056343336287f};
        int i = (int)k;
        float r = tab[0];
#pragma unroll
        for (int j = 1; j < 10; ++j) r = (i == j) ? tab[j] : r;
        return r;
    }
    return series;
}

// Binomial(n, p) by sequential inversion; intended for n*p below ~15 (p <= 0.5), lq = log(1-p).
__device__ __forceinline__ int binom_inversion(Philox& rng, int n, float p, float lq) {
    float f = __expf((float)n * lq);
    float s = __fdividef(p, 1.f - p);
    float u = rng.uniform();
    int k = 0;
    while (u > f) {
        u -= f;
        ++k;
        if (k > n) { k = n; break; }
        f *= s * __fdividef((float)(n - k + 1), (float)k);
        if (f < 1e-35f) break;   // tail mass below float resolution
    }
    return k;
}

// Binomial(n, p) by BTRS (Hormann 1993), valid for n*p >= 10, p <= 0.5.
__device__ __forceinline__ int binom_btrs(Philox& rng, int n, float p) {
    const float nf = (float)n;
    const float q = 1.f - p;
    const float spq = sqrtf(nf * p * q);
    const float b = 1.15f + 2.53f * spq;
    const float a = -0.0873f + 0.0248f * b + 0.01f * p;
    const float c = nf * p + 0.5f;
    const float vr = 0.92f - __fdividef(4.2f, b);
    const float alpha = (2.83f + __fdividef(5.1f, b)) * spq;
    const float m = floorf((nf + 1.f) * p);
    const float r = __fdividef(p, q);
    // lam = log(r (n-m+1) / (m+1)), |lam| = O(1/m)
    const float lam = log1pf(__fdividef(r * (nf - m + 1.f) - (m + 1.f), m + 1.f));
    const float fm = stirling_tail(m) + stirling_tail(nf - m);
    for (int it = 0; it < 64; ++it) {
        float u = rng.uniform() - 0.5f;
        float v = rng.uniform();
        float us = 0.5f - fabsf(u);
        float kf = floorf((2.f * __fdividef(a, us) + b) * u + c);
        if (us >= 0.07f && v <= vr) return (int)kf;
        if (kf < 0.f || kf > nf) continue;
        float lv = __logf(v * __fdividef(alpha, __fdividef(a, us * us) + b));
        float d = kf - m;
        // log(f(k)/f(m)) in cancellation-free form (see header comment)
        float bound = -(kf + 0.5f) * log1pf(__fdividef(d, m + 1.f))
                      - (nf - kf + 0.5f) * log1pf(__fdividef(-d, nf - m + 1.f))
                      + d * lam + fm - stirling_tail(kf) - stirling_tail(nf - kf);
        if (lv <= bound) return (int)kf;
    }
    return (int)m;  // unreachable in practice (acceptance probability > 0.8 per iteration)
}

struct BootParams {
    const BootEntry* entries;
    const long long* seg_ptr;
    long long seg_lo;
    long long n_seg;
    int R;
    const int* seg_U;           // [n_seg]; -1 => NaN row; rows with seg_skip != 0 are not computed
    const unsigned char* seg_skip;  // [n_seg] nullable
    const int* group_ncells;    // [R]
    const double* mv_fit;       // [R][3], highest power first
    int estimator;
    int B;
    unsigned long long seed;
    const long long* gene_id;   // [n_seg / R] global gene ids for the RNG counter (nullable: local index)
    double* out_mean;           // [n_seg][B]
    double* out_rv;             // [n_seg][B]
};

__device__ __forceinline__ void finish_replicate(double M1, double M2, double n, int estimator,
                                                 const double* fit, double& mean, double& rv) {
    double var;
    if (estimator == 0) { mean = M1 / n; var = M2 / n - mean * mean; }
    else { mean = M1 / n + 1.0; var = 10.0; }
    if (mean > 0.0 && var > 0.0) {
        double lm = log(mean);
        rv = exp(log(var) - ((fit[0] * lm + fit[1]) * lm + fit[2]));
    } else {
        rv = nan("");
    }
}

__global__ void __launch_bounds__(kBootThreads)
bootstrap_1d_kernel(BootParams P) {
    const long long seg_rel = blockIdx.y;
    const int b = blockIdx.x * kBootThreads + threadIdx.x;
    if (b >= P.B) return;
    if (P.seg_skip && P.seg_skip[seg_rel]) return;
    const long long seg = P.seg_lo + seg_rel;
    const int r = (int)(seg % P.R);
    const int U = P.seg_U[seg_rel];
    const long long o = seg_rel * (long long)P.B + b;
    if (U < 0) { P.out_mean[o] = nan(""); P.out_rv[o] = nan(""); return; }
    const BootEntry* tab = P.entries + (P.seg_ptr[seg] - P.seg_ptr[P.seg_lo]);
    const int n_cells = P.group_ncells[r];
    // RNG stream id: global (gene, group) so that results do not depend on tiling or gene sharding
    const long long sid = P.gene_id ? P.gene_id[seg_rel / P.R] * P.R + r : seg;
    Philox rng;
    rng.init(P.seed, (uint32_t)b, (uint32_t)sid, 0u, (uint32_t)(sid >> 32) ^ 0x1D1Du);
    int n_rem = n_cells;
    double M1 = 0.0, M2 = 0.0;
    for (int u = 0; u < U; ++u) {
        const int4 lo4 = __ldg(reinterpret_cast<const int4*>(tab + u));
        const int4 hi4 = __ldg(reinterpret_cast<const int4*>(tab + u) + 1);
        const double ea = __hiloint2double(lo4.y, lo4.x);
        const double eb = __hiloint2double(lo4.w, lo4.z);
        const float p = __int_as_float(hi4.x), lq = __int_as_float(hi4.y);
        const int mode = hi4.w;
        int k;
        if ((mode & 2) && (float)n_rem * p >= 10.f) k = binom_btrs(rng, n_rem, p);
        else k = binom_inversion(rng, n_rem, p, lq);
        if (mode & 1) k = n_rem - k;
        n_rem -= k;
        M1 = fma(ea, (double)k, M1);
        M2 = fma(eb, (double)k, M2);
        if (n_rem <= 0) break;
    }
    double mean, rv;
    finish_replicate(M1, M2, (double)n_cells, P.estimator, P.mv_fit + 3 * r, mean, rv);
    P.out_mean[o] = mean;
    P.out_rv[o] = rv;
}

// ---------------------------------------------------------------- deterministic replay
struct ReplayParams {
    const double* x;            // [sum U] distinct counts, reference table order
    const double* inv_sf;       // [sum U]
    const long long* W;         // per table t: B x U_t block at W + B * tab_ptr[t], row-major (replicate, category)
    const long long* tab_ptr;   // [n_tab + 1]
    const int* n_cells;         // [n_tab]
    const double* q;            // [n_tab]
    const double* mv_fit;       // [n_tab][3]
    int n_tab, B, estimator;
    double* out_mean;           // [n_tab][B]
    double* out_var;            // [n_tab][B]  plain variance
    double* out_rv;             // [n_tab][B]  residual variance
};

__global__ void __launch_bounds__(kBootThreads)
bootstrap_1d_replay_kernel(ReplayParams P) {
    const int t = blockIdx.y;
    const int b = blockIdx.x * kBootThreads + threadIdx.x;
    if (b >= P.B) return;
    const long long lo = P.tab_ptr[t], U = P.tab_ptr[t + 1] - lo;
    const long long o = (long long)t * P.B + b;
    if (U <= 1) {  // reference bootstrap.py:97-98
        P.out_mean[o] = nan(""); P.out_var[o] = nan(""); P.out_rv[o] = nan("");
        return;
    }
    const long long* Wb = P.W + (long long)P.B * lo + (long long)b * U;
    const double q = P.q[t];
    double M1 = 0.0, M2 = 0.0;
    for (long long u = 0; u < U; ++u) {
        double x = P.x[lo + u], w = P.inv_sf[lo + u];
        double k = (double)Wb[u];
        M1 = fma(x * w, k, M1);
        M2 = fma((x * x - (1.0 - q) * x) * w * w, k, M2);
    }
    double n = (double)P.n_cells[t];
    double mean, rv;
    finish_replicate(M1, M2, n, P.estimator, P.mv_fit + 3 * t, mean, rv);
    P.out_mean[o] = mean;
    P.out_var[o] = (P.estimator == 0) ? (M2 / n - (M1 / n) * (M1 / n)) : 10.0;
    P.out_rv[o] = rv;
}

}  // namespace mm

using namespace mm;

MM_EXPORT int mm_bootstrap_1d(int device, void* stream, const void* entries, const int64_t* seg_ptr,
                              int64_t seg_lo, int64_t n_seg, int32_t R, const int32_t* seg_U,
                              const uint8_t* seg_skip, const int32_t* group_ncells, const double* mv_fit,
                              int32_t estimator, int32_t num_boot, uint64_t seed, const int64_t* gene_id,
                              double* out_mean, double* out_rv) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_seg >= 0 && R > 0 && num_boot > 0, "n_seg/R/num_boot");
    MM_REQUIRE(n_seg <= 65535, "at most 65535 segments per launch (tile the genes)");
    MM_REQUIRE(estimator == 0 || estimator == 1, "estimator");
    if (n_seg == 0) return 0;
    MM_REQUIRE(entries && seg_ptr && seg_U && group_ncells && mv_fit && out_mean && out_rv, "null pointer");
    BootParams P;
    P.entries = (const BootEntry*)entries; P.seg_ptr = (const long long*)seg_ptr; P.seg_lo = seg_lo;
    P.n_seg = n_seg; P.R = R; P.seg_U = seg_U; P.seg_skip = seg_skip; P.group_ncells = group_ncells;
    P.mv_fit = mv_fit; P.estimator = estimator; P.B = num_boot; P.seed = seed;
    P.gene_id = (const long long*)gene_id; P.out_mean = out_mean; P.out_rv = out_rv;
    dim3 grid((num_boot + kBootThreads - 1) / kBootThreads, (unsigned)n_seg);
    bootstrap_1d_kernel<<<grid, kBootThreads, 0, (cudaStream_t)stream>>>(P);
    return check_launch("mm_bootstrap_1d");
}

MM_EXPORT int mm_bootstrap_1d_replay(int device, void* stream, const double* x, const double* inv_sf,
                                     const int64_t* W, const int64_t* tab_ptr, const int32_t* n_cells,
                                     const double* q, const double* mv_fit, int32_t n_tab, int32_t num_boot,
                                     int32_t estimator, double* out_mean, double* out_var, double* out_rv) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_tab >= 0 && num_boot > 0, "n_tab/num_boot");
    MM_REQUIRE(n_tab <= 65535, "at most 65535 tables per launch");
    if (n_tab == 0) return 0;
    MM_REQUIRE(x && inv_sf && W && tab_ptr && n_cells && q && mv_fit && out_mean && out_var && out_rv, "null pointer");
    ReplayParams P;
    P.x = x; P.inv_sf = inv_sf; P.W = (const long long*)W; P.tab_ptr = (const long long*)tab_ptr;
    P.n_cells = n_cells; P.q = q; P.mv_fit = mv_fit; P.n_tab = n_tab; P.B = num_boot; P.estimator = estimator;
    P.out_mean = out_mean; P.out_var = out_var; P.out_rv = out_rv;
    dim3 grid((num_boot + kBootThreads - 1) / kBootThreads, (unsigned)n_tab);
    bootstrap_1d_replay_kernel<<<grid, kBootThreads, 0, (cudaStream_t)stream>>>(P);
    return check_launch("mm_bootstrap_1d_replay");
}
