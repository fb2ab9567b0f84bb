// matvec_tc.cu -- K2-TC: Q4_0 dequant + matvec for M <= 8 tokens with the *dequant arithmetic* moved
// onto the tensor cores (mma.sync.m16n8k16, f16 x f16 -> f32), optionally fused with the RMSNorm
// (+ ADA scale) of its input.
//
// Why: the SIMT kernel (kernels.cu K2) needs ~4 issue slots per weight (SHF, LOP3, I2FP, FFMA);
// B200 streams 11.7 T weights/s from HBM (6.58 TB/s / 0.5625 B) but issues only ~36 T lane-instr/s,
// so a SIMT single-token decode can not be HBM-bound (profiles/README.md).  Here a Q4 nibble n
// reaches the MMA as the f16 *subnormal* n * 2^-24 (nibble in bits 0-3 of a 16-bit lane) or
// 16 n * 2^-24 (bits 4-7): one LOP3 isolates two weights, no int->float conversion, no multiply
// (scripts/mma_denorm_test.cu verifies on hardware that HMMA treats these inputs exactly).
// Same arithmetic as the reference (src/gguf/shader.wgsl:96-127), re-associated:
//
//   y[m,n] = sum_b d[n,b] * ( sum_{k in b} q[n,k] x[m,k]  -  8 sum_{k in b} x[m,k] )
//
// * inner sums over one 32-weight block = two MMAs (low nibbles, high nibbles) with f32 accumulate;
// * x[m,:] is scaled by a power of two (max|x| -> ~[2^7,2^8)) and split into two f16 pieces
//   hi = f16(x), mid = f16(x - hi): 22 mantissa bits, absolute error <= 2^-25 in scaled units --
//   below the f32 rounding of the dot product itself; products n*hi are exact in f32;
// * the per-block f16 scale d is applied to the f32 block sum in registers (it can not be folded into
//   the MMA: (q-8)*d needs 15 mantissa bits), exactly once per block like the reference.
// * NORM variant: x := ((x / sqrt(mean(x^2)+eps)) * gamma) * ada   (reference rms_norm.rs:42-47 +
//   model.rs:250-255) computed in the staging pass -- every CTA re-derives the row statistics from
//   the 12 KB input instead of a separate launch.
//
// Weight layout ("TC layout", built at load from the GGUF blocks, same 18 B / 32 weights):
//   qs_tc : uint4 [T = N/16 tiles][P = K/64 block pairs][32 lanes]
//           lane (g = lane/4, t = lane%4): .x = word t of (row 16T+g,   block 2P)
//                                          .y = word t of (row 16T+g+8, block 2P)
//                                          .z/.w = the same for block 2P+1
//           => a warp reads 512 contiguous bytes per request and each lane already holds its
//              m16n8k16 A-fragment source words.
//   d_tc  : uint2 [T][P][8]  = halves {d(g,2P), d(g+8,2P), d(g,2P+1), d(g+8,2P+1)}
// MMA K index <-> block element (low-nibble MMA; high-nibble MMA adds 16):
//   kcol 2t -> 4t, 2t+1 -> 4t+2, 2t+8 -> 4t+1, 2t+9 -> 4t+3   (B fragments are staged to match).
#include <cuda_fp16.h>

#include "common.h"
#include "kernels.h"

namespace vox {

void tc_count_launch(const char *name);
void set_tc_pdl(bool on);

namespace {

inline void cuda_check_tc(cudaError_t e, const char *what) {
    if (e != cudaSuccess) fail(VOX_ECUDA, fmt("CUDA error: %s: %s", what, cudaGetErrorString(e)));
}

constexpr int TC_WARPS = 8;
constexpr int TC_THREADS = TC_WARPS * 32;
constexpr int TC_UNROLL = 4;

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t a0, const uint32_t a1, const uint32_t a2,
                                         const uint32_t a3, const uint32_t b0, const uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D TMA bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Programmatic dependent launch: let the next kernel in the stream start its prologue (weight
// prefetch) now; block until the previous kernel's results are visible.  No-ops when the kernel was
// launched without the programmatic-serialization attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }

struct NormArgs {
    const float *gamma;  // nullptr => no normalisation
    const float *ada;    // optional extra elementwise scale
    float eps;
};

// Effective activation value (after the optional fused RMSNorm).
__device__ __forceinline__ float4 eff4(const float4 v, const float rinv_is_div, const float rms, const float *gamma,
                                       const float *ada, const int k) {
    (void)rinv_is_div;
    if (!gamma) return v;
    const float4 g = *reinterpret_cast<const float4 *>(gamma + k);
    float4 o = make_float4((v.x / rms) * g.x, (v.y / rms) * g.y, (v.z / rms) * g.z, (v.w / rms) * g.w);
    if (ada) {
        const float4 a = *reinterpret_cast<const float4 *>(ada + k);
        o.x *= a.x; o.y *= a.y; o.z *= a.z; o.w *= a.w;
    }
    return o;
}

// Shared-memory staging of the activation side for `nb` blocks starting at block b0:
//   bf  : uint2 [nb][2 (nibble half j)][2M cols][4 t]   B fragments {b0,b1} of lane (g = col, t)
//   off : float [nb][M]                                  -8 * sum_{k in block} x  (scaled units, * 2^-24)
// Column c = 2*token + split (0 = hi, 1 = mid).  One work item = (block, token, t): elements
// 4t..4t+3 and 16+4t..16+4t+3 of the block; the four t-items of a block sit in adjacent lanes so the
// block sum is a 2-step shuffle.
template <int M>
__device__ __forceinline__ void tc_stage(const float *__restrict__ x, const int K, const int b0, const int nb,
                                         const float *__restrict__ sx, const float *__restrict__ rms,
                                         const NormArgs na, uint2 *__restrict__ bf, float *__restrict__ off) {
    const int items = nb * M * 4;
    for (int base = 0; base < items; base += TC_THREADS) {
        const int i = base + threadIdx.x;
        const bool active = i < items;
        const int t = i & 3;
        const int m = active ? (i >> 2) % M : 0;
        const int bl = active ? (i >> 2) / M : 0;
        const int kb = (b0 + bl) * 32;
        float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
        if (active && kb < K) {
            lo = *reinterpret_cast<const float4 *>(x + (size_t)m * K + kb + 4 * t);
            hi = *reinterpret_cast<const float4 *>(x + (size_t)m * K + kb + 16 + 4 * t);
            lo = eff4(lo, 0.f, rms[m], na.gamma, na.ada, kb + 4 * t);
            hi = eff4(hi, 0.f, rms[m], na.gamma, na.ada, kb + 16 + 4 * t);
        }
        float bs = ((lo.x + lo.y) + (lo.z + lo.w)) + ((hi.x + hi.y) + (hi.z + hi.w));
        bs += __shfl_xor_sync(0xffffffffu, bs, 1);
        bs += __shfl_xor_sync(0xffffffffu, bs, 2);
        if (!active) continue;
        const float s = sx[m];
        const float e[8] = {lo.x * s, lo.y * s, lo.z * s, lo.w * s,
                            hi.x * s * 0.0625f, hi.y * s * 0.0625f, hi.z * s * 0.0625f, hi.w * s * 0.0625f};
        float h[8], md[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            h[q] = __half2float(__float2half_rn(e[q]));
            md[q] = e[q] - h[q];
        }
        // j = 0: b0 = {elem 4t, 4t+2}, b1 = {4t+1, 4t+3};  j = 1: same on the (x/16) high half
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int o = 4 * j;
            uint2 fh, fm;
            fh.x = pack_h2(h[o + 0], h[o + 2]);
            fh.y = pack_h2(h[o + 1], h[o + 3]);
            fm.x = pack_h2(md[o + 0], md[o + 2]);
            fm.y = pack_h2(md[o + 1], md[o + 3]);
            uint2 *dst = bf + ((size_t)(bl * 2 + j) * (2 * M)) * 4;
            dst[(2 * m + 0) * 4 + t] = fh;
            dst[(2 * m + 1) * 4 + t] = fm;
        }
        if (t == 0) off[bl * M + m] = -8.0f * bs * s * 5.9604644775390625e-08f;  // * 2^-24
    }
}

// EPI semantics as in kernels.h (Epi).
template <int M, int EPI>
__global__ void __launch_bounds__(TC_THREADS)
q4_matvec_tc_kernel(const uint4 *__restrict__ qs_tc, const uint2 *__restrict__ d_tc, const int N, const int K,
                    const int n_tiles, const int n_pairs, const float *__restrict__ x, float *__restrict__ y,
                    const int ldy, const float *__restrict__ bias, const float *__restrict__ res,
                    const int chunk_pairs, const NormArgs na, const int nbuf) {
    constexpr int CG = (M + 3) / 4;  // column groups of 8 (= 4 tokens x 2 splits)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *sx = reinterpret_cast<float *>(smem_raw);                  // [0..7] scale, [8..15] 1/scale, [16..23] rms
    float *stat = sx + 24;                                            // [TC_WARPS][M][2] partial (ssq, amax)
    float *red = stat + TC_WARPS * M * 2;                             // [TC_WARPS][16 rows][M]
    float *off = red + TC_WARPS * 16 * M;                             // [chunk blocks][M]
    uint2 *bf = reinterpret_cast<uint2 *>(off + (size_t)chunk_pairs * 2 * M);  // [chunk blocks][2][2M][4]
    // weight tiles arrive by TMA bulk copies: [nbuf][n_pairs*512 B nibbles | n_pairs*64 B scales]
    const uint32_t tile_q_bytes = (uint32_t)n_pairs * 512u, tile_d_bytes = (uint32_t)n_pairs * 64u;
    const uint32_t tile_bytes = tile_q_bytes + tile_d_bytes;
    size_t woff = (size_t)(reinterpret_cast<unsigned char *>(bf + (size_t)chunk_pairs * 2 * 2 * 2 * M * 4) - smem_raw);
    woff = (woff + 127) & ~(size_t)127;
    unsigned char *wbuf = smem_raw + woff;
    uint64_t *mbar = reinterpret_cast<uint64_t *>(wbuf + (size_t)nbuf * tile_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;

    auto issue_tile = [&](int tile, int buf) {  // one thread
        mbar_expect_tx(&mbar[buf], tile_bytes);
        bulk_g2s(wbuf + (size_t)buf * tile_bytes, qs_tc + (size_t)tile * n_pairs * 32, tile_q_bytes, &mbar[buf]);
        bulk_g2s(wbuf + (size_t)buf * tile_bytes + tile_q_bytes, d_tc + (size_t)tile * n_pairs * 8, tile_d_bytes, &mbar[buf]);
    };
    if (threadIdx.x == 0) {
        for (int i = 0; i < nbuf; ++i) mbar_init(&mbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    // the weights do not depend on the previous kernel: start streaming the first tile now, then let
    // the next kernel begin its own prefetch, and only then wait for our input activations
    if (threadIdx.x == 0 && (int)blockIdx.x < n_tiles) issue_tile(blockIdx.x, 0);
    pdl_trigger();
    pdl_wait();

    // ---- pass 1: per-token sum of squares (for the fused RMSNorm) and max |x * gamma * ada|
    {
        const int kq = K >> 2;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            float ssq = 0.0f, amax = 0.0f;
            for (int i = threadIdx.x; i < kq; i += TC_THREADS) {
                const float4 v = *reinterpret_cast<const float4 *>(x + (size_t)m * K + 4 * i);
                ssq = fmaf(v.x, v.x, ssq); ssq = fmaf(v.y, v.y, ssq); ssq = fmaf(v.z, v.z, ssq); ssq = fmaf(v.w, v.w, ssq);
                float4 w = v;
                if (na.gamma) {
                    const float4 gg = *reinterpret_cast<const float4 *>(na.gamma + 4 * i);
                    w.x *= gg.x; w.y *= gg.y; w.z *= gg.z; w.w *= gg.w;
                    if (na.ada) {
                        const float4 aa = *reinterpret_cast<const float4 *>(na.ada + 4 * i);
                        w.x *= aa.x; w.y *= aa.y; w.z *= aa.z; w.w *= aa.w;
                    }
                }
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(w.x), fabsf(w.y)), fmaxf(fabsf(w.z), fabsf(w.w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
                amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            }
            if (lane == 0) {
                stat[(warp * M + m) * 2 + 0] = ssq;
                stat[(warp * M + m) * 2 + 1] = amax;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < M) {
        const int m = threadIdx.x;
        float ssq = 0.0f, amax = 0.0f;
#pragma unroll
        for (int w = 0; w < TC_WARPS; ++w) {
            ssq += stat[(w * M + m) * 2 + 0];
            amax = fmaxf(amax, stat[(w * M + m) * 2 + 1]);
        }
        float rms = 1.0f;
        if (na.gamma) {
            rms = sqrtf(ssq / (float)K + na.eps);
            amax = amax / rms;
        }
        float s = 1.0f;
        if (amax > 0.0f && amax < 3.0e38f) {
            const int e = (int)((__float_as_uint(amax) >> 23) & 0xFF) - 127;  // floor(log2 amax) for normals
            int se = 7 - e;
            se = se > 100 ? 100 : (se < -100 ? -100 : se);
            s = __uint_as_float((uint32_t)(se + 127) << 23);
        }
        sx[m] = s;
        sx[8 + m] = 1.0f / s;  // exact: power of two
        sx[16 + m] = rms;
    }
    __syncthreads();

    const int n_chunks = (n_pairs + chunk_pairs - 1) / chunk_pairs;
    if (n_chunks == 1) {
        tc_stage<M>(x, K, 0, n_pairs * 2, sx, sx + 16, na, bf, off);
        __syncthreads();
    }

    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = nbuf == 2 ? (it & 1) : 0;
        const uint32_t parity = nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
        const int next_tile = tile + gridDim.x;
        if (nbuf == 2 && threadIdx.x == 0 && next_tile < n_tiles) {
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            issue_tile(next_tile, buf ^ 1);  // buffer buf^1 was released by the barrier ending iteration it-1
        }
        const uint4 *wq_s = reinterpret_cast<const uint4 *>(wbuf + (size_t)buf * tile_bytes);
        const uint2 *wd_s = reinterpret_cast<const uint2 *>(wbuf + (size_t)buf * tile_bytes + tile_q_bytes);
        mbar_wait(&mbar[buf], parity);
        float acc[CG][2];
#pragma unroll
        for (int c = 0; c < CG; ++c) acc[c][0] = acc[c][1] = 0.0f;

        for (int ch = 0; ch < n_chunks; ++ch) {
            const int p0 = ch * chunk_pairs;
            const int np = min(chunk_pairs, n_pairs - p0);
            if (n_chunks > 1) {
                __syncthreads();
                tc_stage<M>(x, K, p0 * 2, np * 2, sx, sx + 16, na, bf, off);
                __syncthreads();
            }
            const uint4 *qp = wq_s + (size_t)p0 * 32 + lane;
            const uint2 *dp = wd_s + (size_t)p0 * 8 + g;
            for (int pp0 = warp; pp0 < np; pp0 += TC_WARPS * TC_UNROLL) {
                uint4 wq[TC_UNROLL];
                uint2 wd[TC_UNROLL];
#pragma unroll
                for (int u = 0; u < TC_UNROLL; ++u) {
                    const int pp = pp0 + u * TC_WARPS;
                    if (pp < np) {
                        wq[u] = qp[(size_t)pp * 32];
                        wd[u] = dp[(size_t)pp * 8];
                    }
                }
#pragma unroll
                for (int u = 0; u < TC_UNROLL; ++u) {
                    const int pp = pp0 + u * TC_WARPS;
                    if (pp >= np) break;
                    const uint32_t words[2][2] = {{wq[u].x, wq[u].y}, {wq[u].z, wq[u].w}};
                    const __half2 dlo = *reinterpret_cast<const __half2 *>(&wd[u].x);
                    const __half2 dhi = *reinterpret_cast<const __half2 *>(&wd[u].y);
                    const float dsc[2][2] = {{__low2float(dlo), __high2float(dlo)}, {__low2float(dhi), __high2float(dhi)}};
#pragma unroll
                    for (int bb = 0; bb < 2; ++bb) {
                        const int bl = pp * 2 + bb;  // block index within the chunk
                        const uint32_t wg = words[bb][0], wg8 = words[bb][1];
                        const uint32_t sg = wg >> 8, sg8 = wg8 >> 8;
                        const uint32_t a_lo[4] = {wg & 0x000F000Fu, wg8 & 0x000F000Fu, sg & 0x000F000Fu, sg8 & 0x000F000Fu};
                        const uint32_t a_hi[4] = {wg & 0x00F000F0u, wg8 & 0x00F000F0u, sg & 0x00F000F0u, sg8 & 0x00F000F0u};
                        const uint2 *bfb = bf + (size_t)(bl * 2) * (2 * M) * 4;
#pragma unroll
                        for (int c = 0; c < CG; ++c) {
                            const int col = c * 8 + g;
                            uint2 blo = make_uint2(0u, 0u), bhi = blo;
                            if (col < 2 * M) {
                                blo = bfb[col * 4 + t];
                                bhi = bfb[(2 * M + col) * 4 + t];
                            }
                            float cc[4] = {0.f, 0.f, 0.f, 0.f};
                            mma16816(cc, a_lo[0], a_lo[1], a_lo[2], a_lo[3], blo.x, blo.y);
                            mma16816(cc, a_hi[0], a_hi[1], a_hi[2], a_hi[3], bhi.x, bhi.y);
                            // thread holds (row g | g+8) x (cols 2t, 2t+1) = token 4c+t, splits hi+mid
                            const int tok = c * 4 + t;
                            const float o = tok < M ? off[bl * M + tok] : 0.0f;
                            acc[c][0] = fmaf(dsc[bb][0], (cc[0] + cc[1]) + o, acc[c][0]);
                            acc[c][1] = fmaf(dsc[bb][1], (cc[2] + cc[3]) + o, acc[c][1]);
                        }
                    }
                }
            }
        }
        // ---- cross-warp reduction of the K split, then epilogue
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CG; ++c) {
            const int tok = c * 4 + t;
            if (tok < M) {
                red[(warp * 16 + g) * M + tok] = acc[c][0];
                red[(warp * 16 + g + 8) * M + tok] = acc[c][1];
            }
        }
        __syncthreads();
        if (EPI == EPI_SILU_MUL) {
            for (int i = threadIdx.x; i < 8 * M; i += TC_THREADS) {
                const int pr = i / M, tok = i - pr * M;
                float a = 0.0f, b = 0.0f;
#pragma unroll
                for (int w = 0; w < TC_WARPS; ++w) {
                    a += red[(w * 16 + 2 * pr) * M + tok];
                    b += red[(w * 16 + 2 * pr + 1) * M + tok];
                }
                const float unscale = 16777216.0f * sx[8 + tok];
                a *= unscale;
                b *= unscale;
                const int row = tile * 16 + 2 * pr;
                if (row + 1 < N) y[(size_t)tok * ldy + (row >> 1)] = (a / (1.0f + expf(-a))) * b;
            }
        } else {
            for (int i = threadIdx.x; i < 16 * M; i += TC_THREADS) {
                const int r = i / M, tok = i - r * M;
                float a = 0.0f;
#pragma unroll
                for (int w = 0; w < TC_WARPS; ++w) a += red[(w * 16 + r) * M + tok];
                a *= 16777216.0f * sx[8 + tok];
                const int row = tile * 16 + r;
                if (row < N) {
                    float v = a + (bias ? bias[row] : 0.0f);
                    if (EPI == EPI_RESIDUAL) v += res[(size_t)tok * ldy + row];
                    if (EPI == EPI_GELU) v = 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
                    y[(size_t)tok * ldy + row] = v;
                }
            }
        }
        if (nbuf == 1 && next_tile < n_tiles) {
            __syncthreads();  // everyone is done reading the single weight buffer
            if (threadIdx.x == 0) {
                asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
                issue_tile(next_tile, 0);
            }
        }
    }
}

// programmatic dependent launch between consecutive decode kernels (VOX_PDL=0 disables)
bool g_tc_pdl = !(getenv("VOX_PDL") && getenv("VOX_PDL")[0] == '0');

template <int M, int EPI>
void tc_launch_t(const Q4Weight &w, const float *x, float *y, int ldy, const float *bias, const float *res,
                 const NormArgs &na, cudaStream_t st) {
    constexpr size_t kSmemMax = 200 * 1024;
    const int n_tiles = (w.N + 15) / 16;
    const int n_pairs = (w.K / 32 + 1) / 2;
    const size_t tile_bytes = (size_t)n_pairs * 576;
    const size_t fixed = (24 + TC_WARPS * M * 2 + TC_WARPS * 16 * M) * sizeof(float);
    const size_t per_pair = 2 * ((size_t)M * sizeof(float) + (size_t)2 * 2 * M * 4 * sizeof(uint2));
    // activation staging budget: whatever is left beside one weight tile, capped at 100 KB
    size_t budget = kSmemMax - tile_bytes - 256 - fixed;
    if (budget > 100 * 1024) budget = 100 * 1024;
    int chunk_pairs = (int)(budget / per_pair);
    if (chunk_pairs >= n_pairs) chunk_pairs = n_pairs;
    else chunk_pairs = (chunk_pairs / (TC_WARPS * TC_UNROLL)) * (TC_WARPS * TC_UNROLL);
    VOX_CHECK(chunk_pairs > 0, VOX_EINVAL, "q4_matvec_tc: shared-memory budget too small (K=%d, M=%d)", w.K, M);
    const size_t act = fixed + per_pair * chunk_pairs;
    // CTAs per SM by shared memory (one tile buffer), then decide whether a CTA sees several tiles
    int per_sm = (int)((220 * 1024) / (act + tile_bytes + 256));
    per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
    int grid = n_tiles < 148 * per_sm ? n_tiles : 148 * per_sm;
    int nbuf = 1;
    if (n_tiles > grid && act + 2 * tile_bytes + 256 <= kSmemMax) nbuf = 2;
    const size_t smem = act + 128 + (size_t)nbuf * tile_bytes + 64;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        cuda_check_tc(cudaFuncSetAttribute(q4_matvec_tc_kernel<M, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)(kSmemMax + 8 * 1024)),
                      "cudaFuncSetAttribute(q4_matvec_tc)");
        attr_set = kSmemMax + 8 * 1024;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_tc_pdl ? 1 : 0;
    cuda_check_tc(cudaLaunchKernelEx(&cfg, q4_matvec_tc_kernel<M, EPI>, w.qs_tc, w.d_tc, w.N, w.K, n_tiles, n_pairs, x, y,
                                     ldy, bias, res, chunk_pairs, na, nbuf),
                  "cudaLaunchKernelEx(q4_matvec_tc)");
    tc_count_launch("q4_matvec_tc");
}

template <int M>
void tc_launch_m(const Q4Weight &w, const float *x, float *y, int ldy, const float *bias, const float *res, int epi,
                 const NormArgs &na, cudaStream_t st) {
    switch (epi) {
        case EPI_NONE: tc_launch_t<M, EPI_NONE>(w, x, y, ldy, bias, res, na, st); break;
        case EPI_RESIDUAL: tc_launch_t<M, EPI_RESIDUAL>(w, x, y, ldy, bias, res, na, st); break;
        case EPI_SILU_MUL: tc_launch_t<M, EPI_SILU_MUL>(w, x, y, ldy, bias, res, na, st); break;
        case EPI_GELU: tc_launch_t<M, EPI_GELU>(w, x, y, ldy, bias, res, na, st); break;
        default: fail(VOX_EINVAL, "bad epilogue");
    }
}

}  // namespace

void set_tc_pdl(bool on) { g_tc_pdl = on; }

void launch_q4_matvec_tc_norm(const Q4Weight &w, const float *x, int M, float *y, int ldy, const float *bias,
                              const float *res, int epi, const float *gamma, const float *ada, float eps,
                              cudaStream_t st) {
    VOX_CHECK(w.qs_tc != nullptr, VOX_EINVAL, "q4_matvec_tc: weight has no tensor-core layout");
    VOX_CHECK(M >= 1 && M <= 8, VOX_EINVAL, "q4_matvec_tc: M=%d out of range", M);
    VOX_CHECK(w.K % 32 == 0, VOX_EINVAL, "q4_matvec_tc: K=%d not a multiple of 32", w.K);
    const NormArgs na{gamma, ada, eps};
    switch (M) {
        case 1: tc_launch_m<1>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 2: tc_launch_m<2>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 3: tc_launch_m<3>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 4: tc_launch_m<4>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 5: tc_launch_m<5>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 6: tc_launch_m<6>(w, x, y, ldy, bias, res, epi, na, st); break;
        case 7: tc_launch_m<7>(w, x, y, ldy, bias, res, epi, na, st); break;
        default: tc_launch_m<8>(w, x, y, ldy, bias, res, epi, na, st); break;
    }
}

void launch_q4_matvec_tc(const Q4Weight &w, const float *x, int M, float *y, int ldy, const float *bias,
                         const float *res, int epi, cudaStream_t st) {
    launch_q4_matvec_tc_norm(w, x, M, y, ldy, bias, res, epi, nullptr, nullptr, 0.0f, st);
}

}  // namespace vox
