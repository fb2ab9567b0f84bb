/* oracle/q4_fast.c -- vectorisable variant of oracle/q4_ref.c's Q4_0 matmul for the TIMED CPU baseline.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py); never linked into the product.
 *
 * Same arithmetic as src/gguf/shader.wgsl:41-133 ( y = sum_blocks d * sum_k (q - 8) x ), re-associated so a
 * compiler can keep 8 f32 lanes busy: per block the 32 products go to 8 partial sums (lane = element mod 8),
 * the block scale is applied to the partial sums, the 8 lanes are added at the end of the row.  The checker
 * (tests) keeps using the exact-order oracle_q4_matmul; this one is what bench.py times, because a scalar
 * strict-order loop would flatter the GPU (4.9 tok/s on 64 threads vs what the same cores can do).
 * Compiled with -O3 -mavx2 -mfma -ffp-contract=fast (q4.py build_c); agreement with the exact-order port is
 * tested to 2e-5 relative.
 */
#include <stdint.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float f16_bits_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu, bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else {
            int e = -1;
            do { e++; man <<= 1; } while ((man & 0x400u) == 0);
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3FFu) << 13);
        }
    } else if (exp == 31) {
        bits = sign | 0x7F800000u | (man << 13);
    } else {
        bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
    }
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#define ORACLE_FAST_AVX2 1
#endif

void oracle_q4_matmul_fast(const float *restrict x, const uint8_t *restrict raw, float *restrict y,
                           const float *restrict bias, int M, int N, int K, int threads) {
    const int bpr = K / 32;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#ifdef ORACLE_FAST_AVX2
    /* y = sum_b d_b * ( sum_k q_k x_k  -  8 * sum_k x_k ): the per-block sums of x are shared by all rows */
    float *sx = (float *)__builtin_alloca((size_t)M * bpr * sizeof(float));
    for (int m = 0; m < M; ++m)
        for (int b = 0; b < bpr; ++b) {
            const float *xb = x + (size_t)m * K + (size_t)b * 32;
            float t = 0.0f;
            for (int i = 0; i < 32; ++i) t += xb[i];
            sx[(size_t)m * bpr + b] = t;
        }
    const __m128i m4 = _mm_set1_epi8(0x0F);
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const uint8_t *row = raw + (size_t)n * bpr * 18;
        for (int m = 0; m < M; ++m) {
            const float *xr = x + (size_t)m * K;
            const float *sxr = sx + (size_t)m * bpr;
            __m256 acc = _mm256_setzero_ps();
            float accs = 0.0f;
            for (int b = 0; b < bpr; ++b) {
                const uint8_t *blk = row + (size_t)b * 18;
                const float d = f16_bits_to_f32((uint16_t)(blk[0] | (blk[1] << 8)));
                const __m128i qb = _mm_loadu_si128((const __m128i *)(blk + 2));
                const __m128i lo = _mm_and_si128(qb, m4);                      /* elements 0..15  */
                const __m128i hi = _mm_and_si128(_mm_srli_epi16(qb, 4), m4);   /* elements 16..31 */
                const float *xb = xr + (size_t)b * 32;
                __m256 s = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(lo)), _mm256_loadu_ps(xb));
                s = _mm256_fmadd_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(lo, 8))), _mm256_loadu_ps(xb + 8), s);
                s = _mm256_fmadd_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(hi)), _mm256_loadu_ps(xb + 16), s);
                s = _mm256_fmadd_ps(_mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_srli_si128(hi, 8))), _mm256_loadu_ps(xb + 24), s);
                acc = _mm256_fmadd_ps(_mm256_set1_ps(d), s, acc);
                accs += d * sxr[b];
            }
            float l[8];
            _mm256_storeu_ps(l, acc);
            const float t = (((l[0] + l[1]) + (l[2] + l[3])) + ((l[4] + l[5]) + (l[6] + l[7]))) - 8.0f * accs;
            y[(size_t)m * N + n] = bias ? t + bias[n] : t;
        }
    }
#else
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const uint8_t *row = raw + (size_t)n * bpr * 18;
        for (int m = 0; m < M; ++m) {
            const float *xr = x + (size_t)m * K;
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int b = 0; b < bpr; ++b) {
                const uint8_t *blk = row + (size_t)b * 18;
                const float d = f16_bits_to_f32((uint16_t)(blk[0] | (blk[1] << 8)));
                const uint8_t *q = blk + 2;
                const float *xb = xr + (size_t)b * 32;
                float w[32];
                for (int i = 0; i < 16; ++i) {
                    w[i] = (float)(q[i] & 0x0F) - 8.0f;
                    w[i + 16] = (float)(q[i] >> 4) - 8.0f;
                }
                float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (int j = 0; j < 4; ++j)
                    for (int l = 0; l < 8; ++l) s[l] += w[8 * j + l] * xb[8 * j + l];
                for (int l = 0; l < 8; ++l) acc[l] += d * s[l];
            }
            float t = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
            y[(size_t)m * N + n] = bias ? t + bias[n] : t;
        }
    }
#endif
}
