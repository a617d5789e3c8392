// GEV tail refinement of small achieved significance levels, on the device.
//
// Replaces reference hypothesis_test.py:94-141: for a test with <= 10 extreme bootstrap replicates the
// null distribution is sorted and each tail (N_exec = 300, 270, ..., 60 points) is fitted with a
// generalised extreme value law until a Kolmogorov-Smirnov check passes; the ASL is then
// N_exec/n * (cdf(-|stat|) + sf(|stat|)) of the fitted tails, else the empirical bound is kept.
//
// The reference does this with scipy (third party, unpinned; here scipy 1.18.1):
//   genextreme.fit  = Nelder-Mead (scipy.optimize.fmin defaults: xtol = ftol = 1e-4, maxiter = maxfun = 600,
//                     rho 1, chi 2, psi 0.5, sigma 0.5, initial simplex x0 * 1.05) on the penalised
//                     negative log-likelihood, started from the skewness-sign shape +-0.5 and the
//                     moment-matched location / scale (_fitstart / _fit_loc_scale_support);
//   kstest          = two-sided D against the fitted cdf, exact p-value > 0.05.
// This file restates that published algorithm: same start point, same simplex rules, same objective
// and penalty (100 * log(DBL_MAX) per point outside the support); the KS decision uses the exact
// critical values D_0.05(n) = scipy.stats.kstwo.isf(0.05, n) tabulated below for the nine tail sizes.
//
// One 2-warp CTA per flagged test: the <= 1024 most extreme values of each side are isolated by a few
// counting passes over the null (thresholds around mean -+ 1.5 sd, adjusted until 300..1024 values
// fall beyond them), sorted in shared memory (16 KB per test, so many tests share an SM), then warp 0
// walks the left-tail ladder and warp 1 the right-tail ladder, each evaluating the likelihood with a
// shuffle reduction.
#include "common.cuh"
#include <float.h>

namespace mm {

constexpr int kGevThreads = 64;    // two warps per test: left tail, right tail
constexpr int kLadder = 9;
__constant__ int c_tail_n[kLadder] = {300, 270, 240, 210, 180, 150, 120, 90, 60};
// scipy.stats.kstwo.isf(0.05, n) for n in c_tail_n (scipy 1.18.1)
__constant__ double c_ks_crit[kLadder] = {0.07783200514647708, 0.08200786390881423, 0.08693932047855427,
                                          0.09288606396320713, 0.100252942300571,   0.1097143944417969,
                                          0.12250018407843426, 0.1411693954054822,  0.1723049003305659};
constexpr double kPenalty = 709.782712893384 * 100.0;   // log(DBL_MAX) * 100

struct GevParams {
    const double* coef_rows;   // [n_rows][B + 1]
    const int* flagged;        // [n_flag] row ids; null: rows 0 .. n_flag - 1, filtered on the device by `extreme`
    const int* extreme;        // [n_rows] extreme counts of mm_regress_asl (used when flagged is null)
    int max_extreme;
    int n_flag, B, sort_cap;
    double* asl;               // [n_rows], updated in place where the GEV path succeeds
    int* status;               // [n_flag] 1 = GEV tails used, 0 = empirical bound kept
};

// negative penalised log-likelihood of GEV(c, loc, scale) (scipy sign convention) over x[0..n)
__device__ double gev_nnlf(const double* x, int n, double c, double loc, double scale, int lane) {
    if (!isfinite(c) || !(scale > 0.0)) return INFINITY;
    double acc = 0.0;
    int bad = 0;
    const double sup = (c > 0.0) ? 1.0 / fmax(c, DBL_MIN) : ((c < 0.0) ? 1.0 / fmin(c, -DBL_MIN) : 0.0);
    for (int i = lane; i < n; i += 32) {
        double z = (x[i] - loc) / scale;
        bool inside = (c > 0.0) ? (z < sup) : ((c < 0.0) ? (z > sup) : true);
        if (!inside || !(z == z)) { ++bad; continue; }
        double lp;
        if (c != 0.0) {
            double cx = c * z;
            double lex2 = log1p(-cx);
            double lpex2 = lex2 / c;
            lp = (cx == 1.0 || cx == -INFINITY) ? -INFINITY : (-exp(lpex2) + lpex2 - lex2);
        } else {
            lp = -exp(-z) - z;
        }
        if (isfinite(lp)) acc -= lp; else ++bad;
    }
    acc = warp_sum(acc);
    bad = warp_sum_int(bad);
    return acc + (bad > 0 ? bad * kPenalty : 0.0) + (double)n * log(scale);
}

__device__ double gev_logcdf_inner(double z, double c) {   // log cdf = -exp(log1p(-c z)/c)
    return (c != 0.0) ? -exp(log1p(-c * z) / c) : -exp(-z);
}
__device__ double gev_cdf(double xv, double c, double loc, double scale) {
    double z = (xv - loc) / scale;
    if (c > 0.0 && z >= 1.0 / c) return 1.0;
    if (c < 0.0 && z <= 1.0 / c) return 0.0;
    return exp(gev_logcdf_inner(z, c));
}
__device__ double gev_sf(double xv, double c, double loc, double scale) {
    double z = (xv - loc) / scale;
    if (c > 0.0 && z >= 1.0 / c) return 0.0;
    if (c < 0.0 && z <= 1.0 / c) return 1.0;
    return -expm1(gev_logcdf_inner(z, c));
}

// scipy genextreme._fitstart: shape from the sign of the skewness, loc / scale by moments, moved if the
// implied support does not contain the data.
__device__ void gev_fitstart(const double* x, int n, int lane, double th[3]) {
    double s = 0.0, mn = INFINITY, mx = -INFINITY;
    for (int i = lane; i < n; i += 32) { s += x[i]; mn = fmin(mn, x[i]); mx = fmax(mx, x[i]); }
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(kFull, mn, o));
        mx = fmax(mx, __shfl_xor_sync(kFull, mx, o));
    }
    const double mu = s / n;
    double m2 = 0.0, m3 = 0.0;
    for (int i = lane; i < n; i += 32) { double d = x[i] - mu; m2 += d * d; m3 += d * d * d; }
    m2 = warp_sum(m2) / n;
    m3 = warp_sum(m3) / n;
    const double c = (m3 / pow(m2, 1.5) < 0.0) ? 0.5 : -0.5;
    // distribution mean / variance at the start shape: c = 0.5 -> (0.227546149094484, 0.8584073464102069);
    // c = -0.5 -> variance infinite => scale estimate 0 -> replaced by 1, loc = sample mean
    double loc, scale;
    if (c > 0.0) {
        scale = sqrt(m2 / 0.8584073464102069);
        loc = mu - scale * 0.227546149094484;
        if (!isfinite(loc)) loc = 0.0;
        if (!(isfinite(scale) && scale > 0.0)) scale = 1.0;
        const double b = 2.0;                      // support z < 1/c
        if (!(mx < loc + b * scale)) { loc = (mx - b) + 0.1 * (mx - mn); scale = 1.0; }
    } else {
        loc = mu; scale = 1.0;
        const double a = -2.0;                     // support z > 1/c
        if (!(loc + a * scale < mn)) { loc = (mn - a) - 0.1 * (mx - mn); scale = 1.0; }
    }
    th[0] = c; th[1] = loc; th[2] = scale;
}

// scipy.optimize.fmin (Nelder-Mead) with its default settings; returns false on a FitError condition
__device__ bool gev_fit(const double* x, int n, int lane, double best[3]) {
    double sim[4][3], fsim[4];
    double x0[3];
    gev_fitstart(x, n, lane, x0);
    for (int k = 0; k < 3; ++k) sim[0][k] = x0[k];
    for (int j = 0; j < 3; ++j) {
        for (int k = 0; k < 3; ++k) sim[j + 1][k] = x0[k];
        sim[j + 1][j] = (x0[j] != 0.0) ? 1.05 * x0[j] : 0.00025;
    }
    int fcalls = 0;
    for (int j = 0; j < 4; ++j) { fsim[j] = gev_nnlf(x, n, sim[j][0], sim[j][1], sim[j][2], lane); ++fcalls; }
    auto sort4 = [&]() {   // stable insertion sort by fsim
        for (int i = 1; i < 4; ++i) {
            double f = fsim[i], v0 = sim[i][0], v1 = sim[i][1], v2 = sim[i][2];
            int j = i - 1;
            while (j >= 0 && fsim[j] > f) {
                fsim[j + 1] = fsim[j];
                sim[j + 1][0] = sim[j][0]; sim[j + 1][1] = sim[j][1]; sim[j + 1][2] = sim[j][2];
                --j;
            }
            fsim[j + 1] = f; sim[j + 1][0] = v0; sim[j + 1][1] = v1; sim[j + 1][2] = v2;
        }
    };
    sort4();
    const int maxit = 600, maxfun = 600;
    int it = 0;
    while (fcalls < maxfun && it < maxit) {
        double dx = 0.0, df = 0.0;
        for (int j = 1; j < 4; ++j) {
            for (int k = 0; k < 3; ++k) dx = fmax(dx, fabs(sim[j][k] - sim[0][k]));
            df = fmax(df, fabs(fsim[0] - fsim[j]));
        }
        if (dx <= 1e-4 && df <= 1e-4) break;
        double xbar[3], xr[3];
        for (int k = 0; k < 3; ++k) {
            xbar[k] = (sim[0][k] + sim[1][k] + sim[2][k]) / 3.0;
            xr[k] = 2.0 * xbar[k] - sim[3][k];
        }
        double fxr = gev_nnlf(x, n, xr[0], xr[1], xr[2], lane); ++fcalls;
        bool shrink = false;
        if (fxr < fsim[0]) {
            double xe[3];
            for (int k = 0; k < 3; ++k) xe[k] = 3.0 * xbar[k] - 2.0 * sim[3][k];
            double fxe = gev_nnlf(x, n, xe[0], xe[1], xe[2], lane); ++fcalls;
            if (fxe < fxr) { for (int k = 0; k < 3; ++k) sim[3][k] = xe[k]; fsim[3] = fxe; }
            else { for (int k = 0; k < 3; ++k) sim[3][k] = xr[k]; fsim[3] = fxr; }
        } else if (fxr < fsim[2]) {
            for (int k = 0; k < 3; ++k) sim[3][k] = xr[k];
            fsim[3] = fxr;
        } else if (fxr < fsim[3]) {
            double xc[3];
            for (int k = 0; k < 3; ++k) xc[k] = 1.5 * xbar[k] - 0.5 * sim[3][k];
            double fxc = gev_nnlf(x, n, xc[0], xc[1], xc[2], lane); ++fcalls;
            if (fxc <= fxr) { for (int k = 0; k < 3; ++k) sim[3][k] = xc[k]; fsim[3] = fxc; }
            else shrink = true;
        } else {
            double xcc[3];
            for (int k = 0; k < 3; ++k) xcc[k] = 0.5 * xbar[k] + 0.5 * sim[3][k];
            double fxcc = gev_nnlf(x, n, xcc[0], xcc[1], xcc[2], lane); ++fcalls;
            if (fxcc < fsim[3]) { for (int k = 0; k < 3; ++k) sim[3][k] = xcc[k]; fsim[3] = fxcc; }
            else shrink = true;
        }
        if (shrink) {
            for (int j = 1; j < 4; ++j) {
                for (int k = 0; k < 3; ++k) sim[j][k] = sim[0][k] + 0.5 * (sim[j][k] - sim[0][k]);
                fsim[j] = gev_nnlf(x, n, sim[j][0], sim[j][1], sim[j][2], lane); ++fcalls;
            }
        }
        sort4();
        ++it;
    }
    best[0] = sim[0][0]; best[1] = sim[0][1]; best[2] = sim[0][2];
    return isfinite(best[0]) && isfinite(best[1]) && (best[2] > 0.0);   // else scipy raises FitError
}

// two-sided KS statistic of sorted x[0..n) against the fitted cdf
__device__ double gev_ks(const double* x, int n, const double th[3], int lane) {
    double d = 0.0;
    for (int i = lane; i < n; i += 32) {
        double cdf = gev_cdf(x[i], th[0], th[1], th[2]);
        d = fmax(d, fmax((double)(i + 1) / n - cdf, cdf - (double)i / n));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d = fmax(d, __shfl_xor_sync(kFull, d, o));
    return d;
}

constexpr int kCand = 1024;     // candidate slots per tail (the 300 most extreme values are selected from them)

// bitonic sort (ascending) of kCand doubles in shared memory by the whole block
__device__ void sort_cand(double* a, int tid) {
    for (int k = 2; k <= kCand; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < kCand; i += kGevThreads) {
                int l = i ^ j;
                if (l > i) {
                    double x = a[i], y = a[l];
                    bool up = ((i & k) == 0);
                    if ((x > y) == up) { a[i] = y; a[l] = x; }
                }
            }
            __syncthreads();
        }
}

__global__ void __launch_bounds__(kGevThreads)
gev_tail_kernel(GevParams P) {
    __shared__ double s_lo[kCand];          // ascending; the c_lo smallest-side candidates first, +inf padding after
    __shared__ double s_hi[kCand];          // ascending; -inf padding first, the c_hi largest-side candidates last
    __shared__ double s_red[2][kGevThreads / 32];
    __shared__ int s_cnt[3];
    __shared__ int s_state[2][kLadder];    // 0 pending, 1 pass, 2 ks-fail, 3 fit error
    __shared__ double s_val[2][kLadder];
    const int row = P.flagged ? P.flagged[blockIdx.x] : (int)blockIdx.x;
    if (!P.flagged) {   // device-side selection (no host round trip): only tests with few extreme replicates
        const int e = P.extreme[row];
        if (e < 0 || e > P.max_extreme || !isfinite(P.asl[row])) {
            if (threadIdx.x == 0) P.status[blockIdx.x] = -1;
            return;
        }
    }
    const double* cr = P.coef_rows + (long long)row * (P.B + 1);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double stat = cr[0], astat = fabs(stat);
    if (tid < 2 * kLadder) { s_state[tid / kLadder][tid % kLadder] = 0; }
    // ---- size, mean and spread of the null (finite entries of coef[1:] - coef[0])
    double sum = 0.0, sq = 0.0;
    int cnt = 0;
    for (int i = tid; i < P.B; i += kGevThreads) {
        double d = cr[i + 1] - stat;
        if (isfinite(d)) { sum += d; sq = fma(d, d, sq); ++cnt; }
    }
    sum = warp_sum(sum); sq = warp_sum(sq); cnt = warp_sum_int(cnt);
    if (tid < 3) s_cnt[tid] = 0;
    if (lane == 0) { s_red[0][warp] = sum; s_red[1][warp] = sq; }
    __syncthreads();
    if (lane == 0) atomicAdd(&s_cnt[0], cnt);
    __syncthreads();
    const int n = s_cnt[0];
    double tsum = 0.0, tsq = 0.0;
    for (int w = 0; w < kGevThreads / 32; ++w) { tsum += s_red[0][w]; tsq += s_red[1][w]; }
    if (n < c_tail_n[0]) {     // fewer than 300 usable replicates: keep the empirical bound
        if (tid == 0) P.status[blockIdx.x] = 0;
        return;
    }
    const double mu = tsum / n;
    double var = tsq / n - mu * mu;
    const double sd = sqrt(var > 0.0 ? var : 0.0);
    // ---- thresholds that isolate between 300 and kCand values on each side (a few counting passes)
    double t_lo = mu - 1.5 * sd, t_hi = mu + 1.5 * sd;
    bool ok_lo = false, ok_hi = false;
    for (int it = 0; it < 24 && !(ok_lo && ok_hi); ++it) {
        int c_lo = 0, c_hi = 0;
        for (int i = tid; i < P.B; i += kGevThreads) {
            double d = cr[i + 1] - stat;
            if (isfinite(d)) { c_lo += (d <= t_lo); c_hi += (d >= t_hi); }
        }
        c_lo = warp_sum_int(c_lo); c_hi = warp_sum_int(c_hi);
        __syncthreads();
        if (tid < 3) s_cnt[tid] = 0;
        __syncthreads();
        if (lane == 0) { atomicAdd(&s_cnt[1], c_lo); atomicAdd(&s_cnt[2], c_hi); }
        __syncthreads();
        c_lo = s_cnt[1]; c_hi = s_cnt[2];
        ok_lo = c_lo >= c_tail_n[0] && c_lo <= kCand;
        ok_hi = c_hi >= c_tail_n[0] && c_hi <= kCand;
        const double step = sd * (it < 8 ? 0.2 : (it < 16 ? 0.05 : 0.0125));
        if (!ok_lo) t_lo += (c_lo < c_tail_n[0]) ? step : -0.5 * step;
        if (!ok_hi) t_hi -= (c_hi < c_tail_n[0]) ? step : -0.5 * step;
    }
    if (!(ok_lo && ok_hi)) {   // heavy ties around the cut: keep the empirical bound
        if (tid == 0) P.status[blockIdx.x] = 0;
        return;
    }
    // ---- gather the candidates and sort them
    for (int i = tid; i < kCand; i += kGevThreads) { s_lo[i] = INFINITY; s_hi[i] = -INFINITY; }
    __syncthreads();
    if (tid < 3) s_cnt[tid] = 0;
    __syncthreads();
    for (int i = tid; i < P.B; i += kGevThreads) {
        double d = cr[i + 1] - stat;
        if (!isfinite(d)) continue;
        if (d <= t_lo) s_lo[atomicAdd(&s_cnt[1], 1)] = d;
        if (d >= t_hi) s_hi[atomicAdd(&s_cnt[2], 1)] = d;
    }
    __syncthreads();
    sort_cand(s_lo, tid);
    sort_cand(s_hi, tid);
    // warp 0 walks the ladder of the left tail, warp 1 that of the right tail, each until a tail size
    // passes the KS check or a fit fails (no speculative fits)
    {
        const int side = warp;
        for (int li = 0; li < kLadder; ++li) {
            const int ne = c_tail_n[li];
            const double* x = side == 0 ? s_lo : s_hi + (kCand - ne);
            double th[3];
            int state;
            double val = 0.0;
            if (!gev_fit(x, ne, lane, th)) state = 3;
            else {
                double d = gev_ks(x, ne, th, lane);
                if (d < c_ks_crit[li]) {
                    state = 1;
                    double pr = side == 0 ? gev_cdf(-astat, th[0], th[1], th[2]) : gev_sf(astat, th[0], th[1], th[2]);
                    val = ((double)ne / (double)n) * pr;
                } else state = 2;
            }
            if (lane == 0) { s_state[side][li] = state; s_val[side][li] = val; }
            if (state != 2) break;
        }
    }
    __syncthreads();
    if (tid == 0) {
        double total = 0.0;
        bool ok = true;
        for (int side = 0; side < 2 && ok; ++side) {
            bool found = false;
            for (int li = 0; li < kLadder; ++li) {
                int s = s_state[side][li];
                if (s == 1) { total += s_val[side][li]; found = true; break; }
                if (s == 3 || s == 0) break;       // fit error -> the reference falls back
            }
            ok = found;
        }
        if (ok && isfinite(total)) { P.asl[row] = total; P.status[blockIdx.x] = 1; }
        else P.status[blockIdx.x] = 0;
    }
}

}  // namespace mm

using namespace mm;

MM_EXPORT int mm_gev_tail_asl(int device, void* stream, const double* coef_rows, const int32_t* flagged,
                              int32_t n_flag, int32_t num_boot, double* asl, int32_t* status,
                              const int32_t* extreme, int32_t max_extreme) {
    if (int s = enter(device)) return s;
    MM_REQUIRE(n_flag >= 0 && num_boot > 0, "n_flag/num_boot");
    if (n_flag == 0) return 0;
    MM_REQUIRE(coef_rows && (flagged || extreme) && asl && status, "null pointer");
    GevParams P;
    P.coef_rows = coef_rows; P.flagged = flagged; P.n_flag = n_flag; P.B = num_boot; P.sort_cap = 0;
    P.asl = asl; P.status = status; P.extreme = extreme; P.max_extreme = max_extreme;
    gev_tail_kernel<<<n_flag, kGevThreads, 0, (cudaStream_t)stream>>>(P);
    return check_launch("mm_gev_tail_asl");
}
