/* oracle/q4_ref.c -- CPU restatement of the reference's fused Q4_0 dequant+matmul.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Used by tests as the checker and
 * by bench.py as the timed CPU baseline ("port"); never linked into the product.
 *
 * Follows:
 *   src/gguf/shader.wgsl:41-133 / shader_naive.wgsl:31-98  (arithmetic + accumulation order)
 *   src/gguf/tensor.rs:83-113                               (block layout / dequant rule)
 *   src/gguf/linear.rs:34-40                                (bias added after the matmul)
 *
 * out[m,n] = sum over blocks b of row n, words wi=0..3:
 *     acc += dot4((lo(wi)-8)*d, x[b*32 + 4wi .. +3]);  acc += dot4((hi(wi)-8)*d, x[b*32+16+4wi ..])
 * all in IEEE f32, no FMA contraction (-ffp-contract=off), dot4 = ((p0+p1)+p2)+p3.
 */
#include <stdint.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline float f16_to_f32(uint16_t h) {
    uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1Fu;
    uint32_t man = h & 0x3FFu;
    uint32_t bits;
    if (exp == 0) {
        if (man == 0) {
            bits = sign;
        } else { /* subnormal */
            int e = -1;
            do { e++; man <<= 1; } while ((man & 0x400u) == 0);
            man &= 0x3FFu;
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | (man << 13);
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

void oracle_q4_dequant(const uint8_t *raw, float *out, long long n_blocks) {
#pragma omp parallel for schedule(static)
    for (long long b = 0; b < n_blocks; ++b) {
        const uint8_t *blk = raw + b * 18;
        float d = f16_to_f32((uint16_t)(blk[0] | (blk[1] << 8)));
        for (int i = 0; i < 16; ++i) {
            uint8_t q = blk[2 + i];
            out[b * 32 + i] = ((float)(q & 0x0F) - 8.0f) * d;
            out[b * 32 + i + 16] = ((float)((q >> 4) & 0x0F) - 8.0f) * d;
        }
    }
}

/* y[M,N] = x[M,K] . W[N,K]^T (+ bias[N]);  W in raw GGUF Q4_0 blocks, row-major by n. */
void oracle_q4_matmul(const float *x, const uint8_t *raw, float *y, const float *bias,
                      int M, int N, int K, int threads) {
    const int bpr = K / 32;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const uint8_t *row = raw + (size_t)n * bpr * 18;
        for (int m = 0; m < M; ++m) {
            const float *xr = x + (size_t)m * K;
            float acc = 0.0f;
            for (int b = 0; b < bpr; ++b) {
                const uint8_t *blk = row + b * 18;
                const float d = f16_to_f32((uint16_t)(blk[0] | (blk[1] << 8)));
                const float *xb = xr + b * 32;
                for (int wi = 0; wi < 4; ++wi) {
                    const uint8_t *q = blk + 2 + wi * 4;
                    float p0 = (((float)(q[0] & 0xF) - 8.0f) * d) * xb[wi * 4 + 0];
                    float p1 = (((float)(q[1] & 0xF) - 8.0f) * d) * xb[wi * 4 + 1];
                    float p2 = (((float)(q[2] & 0xF) - 8.0f) * d) * xb[wi * 4 + 2];
                    float p3 = (((float)(q[3] & 0xF) - 8.0f) * d) * xb[wi * 4 + 3];
                    acc += ((p0 + p1) + p2) + p3;
                    p0 = (((float)(q[0] >> 4) - 8.0f) * d) * xb[16 + wi * 4 + 0];
                    p1 = (((float)(q[1] >> 4) - 8.0f) * d) * xb[16 + wi * 4 + 1];
                    p2 = (((float)(q[2] >> 4) - 8.0f) * d) * xb[16 + wi * 4 + 2];
                    p3 = (((float)(q[3] >> 4) - 8.0f) * d) * xb[16 + wi * 4 + 3];
                    acc += ((p0 + p1) + p2) + p3;
                }
            }
            y[(size_t)m * N + n] = bias ? acc + bias[n] : acc;
        }
    }
}
