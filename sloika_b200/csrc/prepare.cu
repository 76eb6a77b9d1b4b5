// Raw-signal pre-processing on the device, batched, one CTA per read.
//
// Reference: what `basecall.raw_worker` does to a read before the network sees it (sloika/basecall.py:111-118):
//     signal = batch.trim_open_pore(signal, open_pore_fraction)      sloika/batch.py:194-220
//     signal = util.trim_array(signal, *trim)                        sloika/util.py:94-99
//     inMat  = (signal - np.median(signal)) / mad(signal)            sloika/maths.py:4-45 (float64), cast to float32
// On the host this is three np.median calls per read plus one per 100-sample window (~9 ms for a 60 k-sample read on
// one core, i.e. 6.5 M samples/s -- 75 cores to feed one GPU at the measured basecalling rate).
//
// Everything is float64 in the reference's operation order, medians are exact order statistics (for even counts
// the mean (a + b) / 2 of the two middle elements, as np.median), so the float32 output is bit-identical to the
// host path (tests/test_basecall_gpu.py).
//   * window MADs: one warp per window, bitonic sort of the (<= 128) values in shared memory, twice
//   * threshold = np.percentile(local_var, 100 * fraction), 'linear' interpolation: fraction 0 (the CLI default) is the
//     minimum; otherwise the two neighbouring order statistics are selected and blended as NumPy's _lerp does
//   * whole-read median / MAD: most-significant-digit-first radix select over the order-preserving 64-bit keys
//     (8 passes of 8 bits with a shared-memory histogram), plus one pass for the lower middle element
//   * output written straight into the time-major padded batch [Tmax, B] the network consumes, zero padded
#include <cfloat>
#include "common.cuh"

namespace sloika {
namespace prep {

constexpr int THREADS = 256;
constexpr double MAD_FACTOR = 1.4826;

__device__ __forceinline__ unsigned long long to_key(double x) {
    const long long b = __double_as_longlong(x);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}
__device__ __forceinline__ double from_key(unsigned long long k) {
    const long long b = (long long)k;
    return __longlong_as_double(b ^ ((~b >> 63) | (long long)0x8000000000000000ULL));
}

// sort s[0..127] ascending (one warp; unused tail padded with +inf by the caller)
__device__ void warp_sort128(double *s, int lane)
{
    for (int k = 2; k <= 128; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < 64; t += 32) {
                const int i = (t / j) * 2 * j + (t % j), l = i + j;
                const bool up = (i & k) == 0;
                const double a = s[i], b = s[l];
                if ((a > b) == up) { s[i] = b; s[l] = a; }
            }
            __syncwarp();
        }
    }
}
__device__ __forceinline__ double middle_of_sorted(const double *s, int n) {
    return (n & 1) ? s[n >> 1] : (s[(n >> 1) - 1] + s[n >> 1]) / 2.0;       // np.median: mean of the two middle ones
}

// k-th smallest (0-based) of f(0..n-1); all threads of the CTA call it, all get the result.
template <typename F>
__device__ double select_kth(F f, long n, long k, unsigned *hist, unsigned long long *bc)
{
    unsigned long long prefix = 0, mask = 0;
    long rank = k;
    for (int shift = 56; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += THREADS) hist[i] = 0;
        __syncthreads();
        for (long i = threadIdx.x; i < n; i += THREADS) {
            const unsigned long long key = to_key(f(i));
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long cum = 0;
            int bin = 0;
            for (; bin < 255; bin++) {
                if (cum + hist[bin] > rank) break;
                cum += hist[bin];
            }
            bc[0] = prefix | ((unsigned long long)bin << shift);
            bc[1] = (unsigned long long)(rank - cum);
        }
        __syncthreads();
        prefix = bc[0];
        rank = (long)bc[1];
        mask |= 0xffULL << shift;
        __syncthreads();
    }
    return from_key(prefix);
}

// largest element strictly below v, and how many there are (for the lower middle element of an even count)
template <typename F>
__device__ void below(F f, long n, double v, unsigned long long *bc, long *count_out, double *max_out)
{
    if (threadIdx.x == 0) { bc[0] = 0; bc[1] = 0; }
    __syncthreads();
    const unsigned long long kv = to_key(v);
    unsigned long long best = 0, cnt = 0;
    for (long i = threadIdx.x; i < n; i += THREADS) {
        const unsigned long long key = to_key(f(i));
        if (key < kv) { cnt++; best = key > best ? key : best; }
    }
    atomicMax(&bc[0], best);
    atomicAdd(&bc[1], cnt);
    __syncthreads();
    *count_out = (long)bc[1];
    *max_out = from_key(bc[0]);
    __syncthreads();
}

template <typename F>
__device__ double median_of(F f, long n, unsigned *hist, unsigned long long *bc)
{
    const long m = n >> 1;
    const double hi = select_kth(f, n, m, hist, bc);
    if (n & 1) return hi;
    long cnt;
    double lo;
    below(f, n, hi, bc, &cnt, &lo);
    if (cnt < m) lo = hi;                       // rank m-1 is another copy of the same value
    return (lo + hi) / 2.0;
}

__global__ void __launch_bounds__(THREADS)
prepare_signal_kernel(const double *__restrict__ sig, const long *__restrict__ off, int trim0, int trim1, double fraction,
                      int window, double *__restrict__ scratch, long scratch_ld, float *__restrict__ out, long ld_t,
                      int Tmax, int32_t *__restrict__ out_len)
{
    __shared__ double wbuf[THREADS / 32][128];
    __shared__ unsigned hist[256];
    __shared__ unsigned long long bc[2];
    __shared__ int s_first, s_last;
    const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double *x = sig + off[b];
    const long N = off[b + 1] - off[b];
    double *lv = scratch + (long)b * scratch_ld;            // local variation per window
    auto fail = [&](int code) {                              // -1: the reference raises (no window above the threshold)
        for (long t = tid; t < Tmax; t += THREADS) out[t * ld_t + b] = 0.0f;
        if (tid == 0) out_len[b] = code;
    };
    const long nwin = N / window;
    if (nwin < 1) { fail(-1); return; }

    // ---- trim_open_pore: MAD of every window (batch.py:207-214, maths.py:29-45) ----
    for (long w = warp; w < nwin; w += THREADS / 32) {
        double *s = wbuf[warp];
        const double *xw = x + w * window;
        for (int i = lane; i < 128; i += 32) s[i] = i < window ? xw[i] : DBL_MAX;
        __syncwarp();
        warp_sort128(s, lane);
        const double centre = middle_of_sorted(s, window);
        __syncwarp();
        for (int i = lane; i < 128; i += 32) s[i] = i < window ? fabs(xw[i] - centre) : DBL_MAX;
        __syncwarp();
        warp_sort128(s, lane);
        if (lane == 0) lv[w] = MAD_FACTOR * middle_of_sorted(s, window);
        __syncwarp();
    }
    __syncthreads();

    // ---- threshold = np.percentile(local_var, 100 * fraction) ('linear') ----
    auto lvf = [&](long i) { return lv[i]; };
    double thr;
    if (fraction <= 0.0) {
        thr = select_kth(lvf, nwin, 0, hist, bc);            // the minimum
    } else {
        // NumPy: quantile = q / 100; virtual index = n*quantile + (alpha + quantile*(1 - alpha - beta)) - 1, alpha = beta = 1
        const double q = (100.0 * fraction) / 100.0;
        const double vi = ((double)nwin * q + (1.0 + q * (1.0 - 1.0 - 1.0))) - 1.0;
        double prev = floor(vi);
        prev = prev < 0.0 ? 0.0 : (prev > (double)(nwin - 1) ? (double)(nwin - 1) : prev);
        const long ip = (long)prev, in = ip + 1 < nwin ? ip + 1 : nwin - 1;
        const double gamma = vi - prev;
        const double a = select_kth(lvf, nwin, ip, hist, bc), c = select_kth(lvf, nwin, in, hist, bc);
        const double diff = c - a;                           // _lerp (numpy/lib/_function_base_impl.py)
        thr = gamma >= 0.5 ? c - diff * (1.0 - gamma) : a + diff * gamma;
        if (gamma >= 1.0) thr = c;                           // (not reached for valid fractions; _lerp's where(t == 1))
    }

    // ---- first / last window above the threshold (batch.py:216-219) ----
    if (tid == 0) { s_first = 0x7fffffff; s_last = -1; }
    __syncthreads();
    int first = 0x7fffffff, last = -1;
    for (long w = tid; w < nwin; w += THREADS)
        if (lv[w] > thr) { first = min(first, (int)w); last = max(last, (int)w); }
    atomicMin(&s_first, first);
    atomicMax(&s_last, last);
    __syncthreads();
    if (s_last < 0) { fail(-1); return; }                    // np.flatnonzero(...) empty: .min() raises in the reference

    // ---- util.trim_array (util.py:94-99) ----
    const long lo = (long)s_first * window, hi_ = ((long)s_last + 1) * window;
    const long a0 = lo + trim0, a1 = hi_ - trim1;
    const long n = a1 - a0;
    if (n <= 0) { fail(0); return; }                         // "Read too short" (basecall.py:113-115)
    const double *xs = x + a0;

    // ---- (signal - median) / mad, float64 then cast (basecall.py:117-118) ----
    auto raw = [&](long i) { return xs[i]; };
    const double med = median_of(raw, n, hist, bc);
    auto dev = [&](long i) { return fabs(xs[i] - med); };
    const double spread = MAD_FACTOR * median_of(dev, n, hist, bc);
    for (long t = tid; t < Tmax; t += THREADS)
        out[t * ld_t + b] = t < n ? (float)((xs[t] - med) / spread) : 0.0f;
    if (tid == 0) out_len[b] = (int)(n < Tmax ? n : Tmax);
}

}  // namespace prep
}  // namespace sloika

using namespace sloika;

extern "C" size_t sloika_prepare_workspace_bytes(long max_len, int B, int window)
{
    if (max_len < 0 || B < 0 || window < 1) return 0;
    return sizeof(double) * (size_t)B * (size_t)(max_len / window + 1);
}

extern "C" int sloika_prepare_signal_fwd(const double *signals, const long *offsets, int B, long max_len, int trim_start,
                                         int trim_end, double open_pore_fraction, int window, void *ws, size_t ws_bytes,
                                         float *out, long ld_t, int Tmax, int32_t *out_len, void *stream)
{
    if (!signals || !offsets || !out || !out_len || B <= 0 || max_len < 0 || Tmax < 0 || trim_start < 0 || trim_end < 0)
        return SLOIKA_ERR_ARG;
    if (!(open_pore_fraction >= 0.0 && open_pore_fraction <= 1.0) || ld_t < B) return SLOIKA_ERR_ARG;
    if (window < 2 || window > 128) return SLOIKA_ERR_UNSUPPORTED;
    if (!ws || ws_bytes < sloika_prepare_workspace_bytes(max_len, B, window)) return SLOIKA_ERR_WORKSPACE;
    prep::prepare_signal_kernel<<<B, prep::THREADS, 0, (cudaStream_t)stream>>>(
        signals, offsets, trim_start, trim_end, open_pore_fraction, window, static_cast<double *>(ws),
        max_len / window + 1, out, ld_t, Tmax, out_len);
    SLOIKA_RETURN_LAUNCH_STATUS();
}
