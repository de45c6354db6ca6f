// slab.cuh — building blocks of the narrow-band "slab" kernels (one CTA = one (b,f) slab = all T frames x channels).
//
// A slab [T<=256, C] lives in shared memory as a 16-bit UMMA operand tile in the chunk-column layout (umma.cuh):
// 8 channels per 16-byte chunk, rows linear at 16 B, chunk stride kCS.  The T axis is split in two M-tiles of 128
// rows; rows >= T are zero.  Accumulators: TMEM lane = row within the M-tile, column = output channel.
// Thread mapping of every epilogue: 256 threads, warp w -> M-tile (w>>2), TMEM lane quarter (w&3), thread = one
// frame: t = 128*(w>>2) + 32*(w&3) + lane.  Row reductions (LayerNorm, softmax) are therefore thread-local.
#pragma once
#include "common.cuh"
#include "layout.cuh"
#include "umma.cuh"

namespace nbss {

// ---------------------------------------------------------------- TMA bulk copy (global -> smem) of weight images
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// Call from ONE thread. The barrier must have been initialised with count 1.
__device__ __forceinline__ void load_image(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    mbar_expect_tx(bar, bytes);
    bulk_g2s(dst_smem, src, bytes, bar);
}

// ---------------------------------------------------------------- 16-bit intermediates in HBM: "slab tile" layout
// Every 16-bit tensor that only travels between these kernels (saved pre-activations, q|k|v, O, gradient operands) is
// stored as [slab][C/8 chunks][T rows][8 elements]: the byte image of the smem operand tile.  A thread-per-frame
// epilogue then writes 32 consecutive 16-byte pieces per warp instruction (fully coalesced), and a consumer brings a
// chunk column into shared memory with one TMA bulk copy of T*16 bytes.
__device__ __forceinline__ size_t tile_off(size_t slab, int nchunks, int T, int chunk, int t) {
    return (((size_t)slab * nchunks + chunk) * T + t) * 16;
}
// ONE thread: copy `nch` chunk columns (T rows each) of a slab-tile tensor into a smem tile; completion on `bar`.
__device__ __forceinline__ void bulk_load_chunks(unsigned char* tile, uint32_t cs, int row_off, const unsigned char* src,
                                                 int nch, int T, uint64_t* bar) {
    mbar_expect_tx(bar, (uint32_t)(nch * T * 16));
    for (int c = 0; c < nch; ++c) bulk_g2s(tile + (size_t)c * cs + row_off * 16, src + (size_t)c * T * 16, (uint32_t)(T * 16), bar);
}

// ---------------------------------------------------------------- MMA issue helpers (call from ONE thread)
// D[128 x N] (+)= A[128 x 16*ksteps] * B[N x 16*ksteps]^T, both K-major chunk-column tiles.
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t a_cs, uint32_t b_addr, uint32_t b_cs,
                                       int ksteps, uint32_t idesc, uint32_t acc) {
    // the start-address field is the low 14 bits of the descriptor: advancing K is a plain add of (bytes >> 4)
    uint64_t da = sdesc_kmajor(a_addr, a_cs), db = sdesc_kmajor(b_addr, b_cs);
    const uint64_t sa = (uint64_t)((2 * a_cs) >> 4), sb = (uint64_t)((2 * b_cs) >> 4);
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
        umma_f16(tmem_d, da, db, idesc, acc);
        da += sa;
        db += sb;
        acc = 1;
    }
}

// ---------------------------------------------------------------- packing
template <int FMT>
__device__ __forceinline__ uint4 pack8(const float* v) {
    return make_uint4(pack16<FMT>(v[0], v[1]), pack16<FMT>(v[2], v[3]), pack16<FMT>(v[4], v[5]), pack16<FMT>(v[6], v[7]));
}
__device__ __forceinline__ void unpack_f16x2(uint32_t p, float& lo, float& hi) {
    asm("{\n.reg .b16 l, h;\nmov.b32 {l, h}, %2;\ncvt.f32.f16 %0, l;\ncvt.f32.f16 %1, h;\n}" : "=f"(lo), "=f"(hi) : "r"(p));
}
template <int FMT>
__device__ __forceinline__ void unpack16(uint32_t p, float& lo, float& hi) {
    if constexpr (FMT == FMT_F16) unpack_f16x2(p, lo, hi);
    else { lo = bf16lo_to_f32(p); hi = bf16hi_to_f32(p); }
}

// ---------------------------------------------------------------- staging: fp32 rows -> (LayerNorm) -> 16-bit tile
// Warp-per-row, coalesced float4 loads (24 lanes x 16 B = one 96-channel row).  Writes rows [row_off, row_off+256)
// of the tile; rows t >= T are written as zeros.  gamma/beta in shared memory.
template <int FMT, bool LN>
__device__ __forceinline__ void stage_rows96(const float* __restrict__ xslab, int T, unsigned char* tile, int row_off,
                                             const float* s_gamma, const float* s_beta, int warp, int lane,
                                             float* stats_out = nullptr /* [T,2] (mean, rstd) of this slab */,
                                             int nwarps = 8) {
    const bool act = lane < 24;
    float4 g = make_float4(1.f, 1.f, 1.f, 1.f), be = make_float4(0.f, 0.f, 0.f, 0.f);
    if (LN && act) {
        g = *reinterpret_cast<const float4*>(s_gamma + 4 * lane);
        be = *reinterpret_cast<const float4*>(s_beta + 4 * lane);
    }
    const int iters = 256 / nwarps;
#pragma unroll 1
    for (int i = 0; i < iters; i += 4) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = warp + nwarps * (i + j);
            v[j] = (act && r < T) ? __ldg(reinterpret_cast<const float4*>(xslab + (size_t)r * kH) + lane)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = warp + nwarps * (i + j);
            float4 y = v[j];
            if (LN) {
                float s = warp_sum(y.x + y.y + y.z + y.w);
                const float mean = s * (1.f / kH);
                float4 d = act ? make_float4(y.x - mean, y.y - mean, y.z - mean, y.w - mean) : make_float4(0, 0, 0, 0);
                float q = warp_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w);
                const float rstd = rsqrtf(q * (1.f / kH) + 1e-5f);
                y = make_float4(d.x * rstd * g.x + be.x, d.y * rstd * g.y + be.y, d.z * rstd * g.z + be.z,
                                d.w * rstd * g.w + be.w);
                if (stats_out && lane == 0 && r < T) *reinterpret_cast<float2*>(stats_out + 2 * r) = make_float2(mean, rstd);
            }
            if (act) {
                uint2 p = (r < T) ? make_uint2(pack16<FMT>(y.x, y.y), pack16<FMT>(y.z, y.w)) : make_uint2(0u, 0u);
                *reinterpret_cast<uint2*>(tile + (lane >> 1) * kCS + (r + row_off) * 16 + (lane & 1) * 8) = p;
            }
        }
    }
}

// Column sums across the 32 lanes of a warp: on entry lane l holds v[0..31] = one row of a 32-column block; on return
// lane l gets the sum over all 32 rows of column l.  31 shuffles (recursive halving) instead of 32 x 5.
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float a = v[i], b = v[i + s];
            const float send = up ? a : b, keep = up ? b : a;
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return v[0];
}

// 16-column variant (15 shuffles): on return lane l holds the column sum of column (l >> 1) (both lanes of a pair).
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#define NBSS_HALVE(S, N)                                                          \
    {                                                                             \
        const bool up = (lane & S) != 0;                                          \
        _Pragma("unroll") for (int i = 0; i < N; ++i) {                           \
            const float a = v[i], b = v[i + N];                                   \
            v[i] = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, S);    \
        }                                                                         \
    }
    NBSS_HALVE(16, 8)
    NBSS_HALVE(8, 4)
    NBSS_HALVE(4, 2)
    NBSS_HALVE(2, 1)
#undef NBSS_HALVE
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// LayerNorm backward + residual, one warp per frame, fully coalesced.  `tile` holds d z (gradient wrt the LN output) as fp32
// staged by the thread-per-frame TMEM epilogue with 4-float chunks: addr = tile + chunk*cs + (r + row_off)*16.
// Lane l < 24 owns channels 4l..4l+3; it accumulates d gamma / d beta in registers across frames (and slabs).
__device__ __forceinline__ void ln_bwd_rows(const unsigned char* tile, uint32_t cs, int row_off, const float* __restrict__ xs,
                                            const float* __restrict__ dys, float* __restrict__ dxs,
                                            const float* __restrict__ stats, int T, const float4 g4, float4& dg4, float4& db4,
                                            int warp, int lane, int nwarps) {
    const bool act = lane < 24;
#pragma unroll 4
    for (int r = warp; r < T; r += nwarps) {
        float4 dz = make_float4(0, 0, 0, 0), xv = dz, dv = dz;
        if (act) {
            dz = *reinterpret_cast<const float4*>(tile + (size_t)lane * cs + (r + row_off) * 16);
            xv = __ldg(reinterpret_cast<const float4*>(xs + (size_t)r * kH) + lane);
            dv = __ldg(reinterpret_cast<const float4*>(dys + (size_t)r * kH) + lane);
        }
        const float2 st = __ldg(reinterpret_cast<const float2*>(stats + 2 * r));
        const float4 xh = act ? make_float4((xv.x - st.x) * st.y, (xv.y - st.x) * st.y, (xv.z - st.x) * st.y, (xv.w - st.x) * st.y)
                              : make_float4(0, 0, 0, 0);
        const float4 dzg = make_float4(dz.x * g4.x, dz.y * g4.y, dz.z * g4.z, dz.w * g4.w);
        const float m1 = warp_sum(dzg.x + dzg.y + dzg.z + dzg.w) * (1.f / kH);
        const float m2 = warp_sum(dzg.x * xh.x + dzg.y * xh.y + dzg.z * xh.z + dzg.w * xh.w) * (1.f / kH);
        dg4 = make_float4(dg4.x + dz.x * xh.x, dg4.y + dz.y * xh.y, dg4.z + dz.z * xh.z, dg4.w + dz.w * xh.w);
        db4 = make_float4(db4.x + dz.x, db4.y + dz.y, db4.z + dz.z, db4.w + dz.w);
        if (act)
            reinterpret_cast<float4*>(dxs + (size_t)r * kH)[lane] =
                make_float4(dv.x + st.y * (dzg.x - m1 - xh.x * m2), dv.y + st.y * (dzg.y - m1 - xh.y * m2),
                            dv.z + st.y * (dzg.z - m1 - xh.z * m2), dv.w + st.y * (dzg.w - m1 - xh.w * m2));
    }
}

// Block-wide sum of NV per-thread values (256 threads). `red` is smem scratch of 8*NV floats. Result broadcast.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* red, int warp, int lane) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();  // protect `red` from the previous use
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) red[warp * NV + i] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w * NV + i];
        v[i] = s;
    }
}

}  // namespace nbss
