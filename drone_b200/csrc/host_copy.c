/* host_copy.c -- the one host-side data movement of the NumPy-buffer step: the action batch goes into the
 * caller-visible (pinned) action buffer as clamp(action, -1, 1), which is what the reference leaves there
 * (clamp4 in place, dronelib.h:73-79,437).  Plain C with SIMD intrinsics so that the clamp rides on the copy at
 * memcpy speed: max(lo, x) / min(hi, .) with x as the SECOND operand keep a NaN a NaN (MAXPS / MINPS return their
 * second source when an operand is a NaN), like the reference's `if (v < min) ... if (v > max) ... return v`. */
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

static void clamp_copy_scalar(float *dst, const float *src, size_t n) {
    for (size_t k = 0; k < n; k++) {
        const float a = src[k];
        dst[k] = a < -1.0f ? -1.0f : (a > 1.0f ? 1.0f : a);
    }
}

__attribute__((target("avx2"))) static void clamp_copy_avx2(float *dst, const float *src, size_t n) {
    const __m256 lo = _mm256_set1_ps(-1.0f), hi = _mm256_set1_ps(1.0f);
    size_t k = 0;
    for (; k + 32 <= n; k += 32) {
        const __m256 a = _mm256_loadu_ps(src + k), b = _mm256_loadu_ps(src + k + 8), c = _mm256_loadu_ps(src + k + 16),
                     d = _mm256_loadu_ps(src + k + 24);
        _mm256_storeu_ps(dst + k, _mm256_min_ps(hi, _mm256_max_ps(lo, a)));
        _mm256_storeu_ps(dst + k + 8, _mm256_min_ps(hi, _mm256_max_ps(lo, b)));
        _mm256_storeu_ps(dst + k + 16, _mm256_min_ps(hi, _mm256_max_ps(lo, c)));
        _mm256_storeu_ps(dst + k + 24, _mm256_min_ps(hi, _mm256_max_ps(lo, d)));
    }
    clamp_copy_scalar(dst + k, src + k, n - k);
}

static void clamp_copy_sse2(float *dst, const float *src, size_t n) {
    const __m128 lo = _mm_set1_ps(-1.0f), hi = _mm_set1_ps(1.0f);
    size_t k = 0;
    for (; k + 8 <= n; k += 8) {
        const __m128 a = _mm_loadu_ps(src + k), b = _mm_loadu_ps(src + k + 4);
        _mm_storeu_ps(dst + k, _mm_min_ps(hi, _mm_max_ps(lo, a)));
        _mm_storeu_ps(dst + k + 4, _mm_min_ps(hi, _mm_max_ps(lo, b)));
    }
    clamp_copy_scalar(dst + k, src + k, n - k);
}

void b2d_clamp_copy(float *dst, const float *src, size_t n) {
    static int have_avx2 = -1;
    if (have_avx2 < 0) have_avx2 = __builtin_cpu_supports("avx2") ? 1 : 0;
    if (have_avx2) clamp_copy_avx2(dst, src, n);
    else clamp_copy_sse2(dst, src, n);
}
