// Host arithmetic of libpyascore_b200 that wants the host compiler's vector units: the m/z narrowing pass
// (see pa_lib.cu "host-side narrowing of the m/z array" for what it proves and why it is sound).
// AVX2 when the CPU has it (checked at run time), plain C++ otherwise; both give the same flags.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <immintrin.h>

namespace {

struct Extremes { double mn, mx; int bad; };

// out[i] = (float)m[i]; minimum, maximum, "some value is NaN or infinite"
Extremes convert_scalar(const double* m, int64_t P, float* out, int64_t i0, Extremes e) {
    for (int64_t i = i0; i < P; i++) {
        const double v = m[i];
        out[i] = (float)v;
        e.mn = v < e.mn ? v : e.mn;
        e.mx = v > e.mx ? v : e.mx;
        e.bad |= !(v - v == 0.);
    }
    return e;
}

// some (m[i] - dmin) / bin_size lies within eps of an integer
int near_scalar(const double* m, int64_t P, int64_t i0, double dmin, double inv, double eps) {
    int near = 0;
    for (int64_t i = i0; i < P; i++) {
        const double t = (m[i] - dmin) * inv;
        const double r = t - std::floor(t);
        near |= (r < eps) | (r > 1. - eps);
    }
    return near;
}

__attribute__((target("avx2"))) Extremes convert_avx2(const double* m, int64_t P, float* out) {
    Extremes e{m[0], m[0], 0};
    __m256d mn = _mm256_set1_pd(m[0]), mx = mn, bad = _mm256_setzero_pd();
    int64_t i = 0;
    for (; i + 4 <= P; i += 4) {
        const __m256d v = _mm256_loadu_pd(m + i);
        _mm_storeu_ps(out + i, _mm256_cvtpd_ps(v));
        mn = _mm256_min_pd(mn, v);
        mx = _mm256_max_pd(mx, v);
        const __m256d d = _mm256_sub_pd(v, v);                  // 0 for finite values, NaN otherwise
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(d, d, _CMP_UNORD_Q));
    }
    double a[4], b[4];
    _mm256_storeu_pd(a, mn); _mm256_storeu_pd(b, mx);
    for (int k = 0; k < 4; k++) { e.mn = a[k] < e.mn ? a[k] : e.mn; e.mx = b[k] > e.mx ? b[k] : e.mx; }
    e.bad = _mm256_movemask_pd(bad) != 0;
    return convert_scalar(m, P, out, i, e);
}

__attribute__((target("avx2"))) int near_avx2(const double* m, int64_t P, double dmin, double inv, double eps) {
    const __m256d vd = _mm256_set1_pd(dmin), vi = _mm256_set1_pd(inv), ve = _mm256_set1_pd(eps), v1 = _mm256_set1_pd(1. - eps);
    __m256d acc = _mm256_setzero_pd();
    int64_t i = 0;
    for (; i + 4 <= P; i += 4) {
        const __m256d t = _mm256_mul_pd(_mm256_sub_pd(_mm256_loadu_pd(m + i), vd), vi);
        const __m256d r = _mm256_sub_pd(t, _mm256_floor_pd(t));
        acc = _mm256_or_pd(acc, _mm256_or_pd(_mm256_cmp_pd(r, ve, _CMP_LT_OQ), _mm256_cmp_pd(r, v1, _CMP_GT_OQ)));
    }
    return (_mm256_movemask_pd(acc) != 0) | near_scalar(m, P, i, dmin, inv, eps);
}

// both passes in one, for spectra whose first / last value are the extremes (m/z-sorted spectra): the caller checks
// the returned extremes against that assumption
__attribute__((target("avx2"))) Extremes fused_avx2(const double* m, int64_t P, float* out, double dmin, double inv, double eps, int* near_out) {
    Extremes e{m[0], m[0], 0};
    const __m256d vd = _mm256_set1_pd(dmin), vi = _mm256_set1_pd(inv), ve = _mm256_set1_pd(eps), v1 = _mm256_set1_pd(1. - eps);
    __m256d mn = _mm256_set1_pd(m[0]), mx = mn, bad = _mm256_setzero_pd(), acc = _mm256_setzero_pd();
    int64_t i = 0;
    // The float32 copy is written once and next read by the copy engine: streaming stores from the first 16-byte aligned
    // element on keep it out of the caches and spare the read-for-ownership of the destination lines (the pass shares
    // the host's memory system with the DMA it feeds).  pa_narrow_spectra fences before it returns.
    const int64_t head = std::min<int64_t>(P, (int64_t)((32 - ((uintptr_t)out & 31)) & 31) / 4);
    if (head > 0) {
        Extremes h = convert_scalar(m, head, out, 0, e);
        *near_out = near_scalar(m, head, 0, dmin, inv, eps);
        mn = _mm256_set1_pd(h.mn); mx = _mm256_set1_pd(h.mx);
        if (h.bad) bad = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
        i = head;
    } else *near_out = 0;
#define PA_LANE4(v)                                                                                                       \
    do {                                                                                                                  \
        mn = _mm256_min_pd(mn, v);                                                                                        \
        mx = _mm256_max_pd(mx, v);                                                                                        \
        const __m256d d_ = _mm256_sub_pd(v, v);                                                                           \
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(d_, d_, _CMP_UNORD_Q));                                                     \
        const __m256d t_ = _mm256_mul_pd(_mm256_sub_pd(v, vd), vi);                                                       \
        const __m256d r_ = _mm256_sub_pd(t_, _mm256_floor_pd(t_));                                                        \
        acc = _mm256_or_pd(acc, _mm256_or_pd(_mm256_cmp_pd(r_, ve, _CMP_LT_OQ), _mm256_cmp_pd(r_, v1, _CMP_GT_OQ)));      \
    } while (0)
    for (; i + 8 <= P; i += 8) {
        const __m256d va = _mm256_loadu_pd(m + i), vb = _mm256_loadu_pd(m + i + 4);
        _mm256_stream_ps(out + i, _mm256_set_m128(_mm256_cvtpd_ps(vb), _mm256_cvtpd_ps(va)));
        PA_LANE4(va);
        PA_LANE4(vb);
    }
#undef PA_LANE4
    double a[4], b[4];
    _mm256_storeu_pd(a, mn); _mm256_storeu_pd(b, mx);
    for (int k = 0; k < 4; k++) { e.mn = a[k] < e.mn ? a[k] : e.mn; e.mx = b[k] > e.mx ? b[k] : e.mx; }
    e.bad = _mm256_movemask_pd(bad) != 0;
    *near_out |= (_mm256_movemask_pd(acc) != 0) | near_scalar(m, P, i, dmin, inv, eps);
    return convert_scalar(m, P, out, i, e);
}

}  // namespace

// spectra [sa, sb) of a CSR block: out32[i] = (float)mz[i]; flag[s] = 1 when spectrum s must keep its float64 values
void pa_narrow_spectra(const double* mz, const int64_t* spec_off, int64_t sa, int64_t sb, int64_t peak_base,
                       float bin_size, float* out32, uint8_t* flag) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    const double dbs = (double)bin_size, inv = 1.0 / dbs;
    for (int64_t s = sa; s < sb; s++) {
        const int64_t o = spec_off[s], P = spec_off[s + 1] - o;
        const double* m = mz + o;
        float* out = out32 + (o - peak_base);
        flag[s] = 0;
        if (P <= 0) continue;
        // m/z-sorted spectra (the rule): the ends are the extremes, so the bounds are known up front and one pass does both
        // the conversion and the near-boundary test; the extremes it finds are checked against that assumption
        Extremes e;
        int near_fused = -1;
        if (avx2 && m[0] <= m[P - 1] && m[0] > -1e30 && m[P - 1] < 1e30) {
            const double dmin0 = (double)(float)(std::floor(m[0] / 100.) * 100.);
            const double eps0 = std::fabs(m[P - 1]) * 0x1p-23 / dbs + 1e-9;
            e = fused_avx2(m, P, out, dmin0, inv, eps0, &near_fused);
            if (e.mn != m[0] || e.mx != m[P - 1]) near_fused = -1;
        } else e = avx2 ? convert_avx2(m, P, out) : convert_scalar(m, P, out, 0, Extremes{m[0], m[0], 0});
        const double mn = e.mn, mx = e.mx;
        int bad = e.bad;
        // cpp/Spectra.cpp:46-48 on the exact and on the rounded extremes (rounding is monotone: the rounded minimum
        // is the minimum of the rounded values)
        const float min_mz = (float)(std::floor(mn / 100.) * 100.), max_mz = (float)(std::ceil(mx / 100.) * 100.);
        const double mn32 = (double)(float)mn, mx32 = (double)(float)mx;
        if ((float)(std::floor(mn32 / 100.) * 100.) != min_mz || (float)(std::ceil(mx32 / 100.) * 100.) != max_mz) bad = 1;
        if (!(mx < 1e30) || !(mn > -1e30)) bad = 1;
        if (!bad) {
            const double dmin = (double)min_mz;
            // float32 rounding moves m/z by at most mx * 2^-24, i.e. the quotient by that over bin_size; twice that plus
            // the error of the reciprocal product is the margin inside which the exact formula decides
            const double eps = std::fabs(mx) * 0x1p-23 / dbs + 1e-9;
            const int near = near_fused >= 0 ? near_fused : (avx2 ? near_avx2(m, P, dmin, inv, eps) : near_scalar(m, P, 0, dmin, inv, eps));
            if (near)
                for (int64_t i = 0; i < P && !bad; i++) {
                    const double q1 = std::floor((m[i] - dmin) / dbs), q2 = std::floor(((double)out[i] - dmin) / dbs);
                    bad |= q1 != q2;
                }
        }
        flag[s] = (uint8_t)(bad != 0);
    }
    _mm_sfence();          // the streaming stores above are globally visible before the caller hands the buffer on
}
