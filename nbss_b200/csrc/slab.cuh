// slab.cuh — building blocks of the narrow-band "slab" kernels (one CTA = one (b,f) slab = all T frames x channels).
//
// A slab [T<=256, C] lives in shared memory as a 16-bit UMMA operand tile in the chunk-column layout (umma.cuh):
// 8 channels per 16-byte chunk, rows linear at 16 B, chunk stride kCS.  The T axis is split in two M-tiles of 128
// rows; rows >= T are zero.  Accumulators: TMEM lane = row within the M-tile, column = output channel.
// Thread mapping of every TMEM epilogue: 512 threads, warp w -> TMEM lane quarter (w&3), M-tile ((w>>2)&1), channel / key
// half (w>>3); thread = one frame: t = 128*m + 32*(w&3) + lane, so row reductions need at most a 2- or 4-thread exchange.
// The fp32 row phases around them (staging, LayerNorm backward, residual adds) use eight lanes per frame instead.
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

// ---------------------------------------------------------------- L2 prefetch of the next work item's inputs
// cp.async.bulk.prefetch.L2 (SASS UBLKPF.L2): one thread asks the memory system to pull `bytes` (multiple of 16) into L2.
// The persistent kernels issue it for the data their NEXT slab / row group will stage, so those latency-exposed loads
// hit L2 instead of HBM.  No architectural side effects: a wrong address range would only waste bandwidth.
__device__ __forceinline__ void l2_prefetch(const void* p, uint32_t bytes) {
#ifdef NBSS_NO_L2_PREFETCH  // A/B build (`make nopf`): measures what the prefetches are worth
    (void)p; (void)bytes;
    return;
#endif
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// fp32 slab [T,96] (T*384 bytes) in 6 pieces; call with i = 0..5 from six different threads
__device__ __forceinline__ void l2_prefetch_slab(const float* slab, int T, int i) {
    l2_prefetch(reinterpret_cast<const unsigned char*>(slab) + (size_t)i * T * 64, (uint32_t)(T * 64));
}

// ---------------------------------------------------------------- de-phasing the persistent CTAs
// All CTAs of a persistent kernel start together and do identical work, so they walk their phases in lockstep: every SM stages
// its inputs from HBM in the same few microseconds (a chip-wide burst at the full 6.5 TB/s) and then every SM computes while
// HBM idles.  Delaying CTA i by (i mod K) / K of one work item's duration makes the bursts of one part of the chip fall into the
// compute phases of the rest.  `period` = cycles per work item (from the phase profile); K = NBSS_STAGGER (0 = off).
// MEASURED (profiles/r02n_fconv_stage_order.txt): K = 2 and K = 4 make every slab kernel 1-4 % SLOWER - the CTAs are not limited
// by chip-wide bursts, the delay is pure cost - so the product build keeps it off; the hook stays for re-measuring.
#ifndef NBSS_STAGGER
#define NBSS_STAGGER 0
#endif
__device__ __forceinline__ void stagger_start(int period) {
#if NBSS_STAGGER > 1
    const int k = blockIdx.x % NBSS_STAGGER;
    if (k) {
        const long long t0 = clock64(), d = (long long)period * k / NBSS_STAGGER;
        while (clock64() - t0 < d) __nanosleep(256);
    }
#else
    (void)period;
#endif
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
// Call it from every lane of ONE warp with `leader` = elect_one(): the descriptor arithmetic is warp-uniform (uniform
// registers), only the tcgen05.mma itself is predicated.  (A lone `if (tid == 0)` caller may leave leader = true.)
__device__ __forceinline__ void mma_kk(uint32_t tmem_d, uint32_t a_addr, uint32_t a_cs, uint32_t b_addr, uint32_t b_cs,
                                       int ksteps, uint32_t idesc, uint32_t acc, bool leader = true) {
    // the start-address field is the low 14 bits of the descriptor: advancing K is a plain add of (bytes >> 4)
    uint64_t da = sdesc_kmajor(a_addr, a_cs), db = sdesc_kmajor(b_addr, b_cs);
    const uint64_t sa = (uint64_t)((2 * a_cs) >> 4), sb = (uint64_t)((2 * b_cs) >> 4);
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
        if (leader) umma_f16(tmem_d, da, db, idesc, acc);
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

// ---------------------------------------------------------------- 8-lanes-per-row helpers
// The fp32 row phases give every 96-channel row to EIGHT lanes (a warp works on 4 rows at once): lane l8 = lane & 7 of
// row-group sub = lane >> 3 owns the three float4 at channels 4*(l8 + 8j), j = 0..2.  A row's three loads are three
// coalesced 128-byte segments, a row reduction is 3 shuffles shared by 4 rows (instead of 5 for one row), and all 32
// lanes work (a warp-per-row walk only uses 24).
__device__ __forceinline__ float oct_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
}
__device__ __forceinline__ float f4_hsum(const float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float f4_dot(const float4 a, const float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// Per-lane channel-owned accumulators of the row phases (d gamma / d beta ...): 3 float4 at channels 4*(l8 + 8j).
struct Oct12 {
    float4 v[3];
    __device__ __forceinline__ void zero() { v[0] = v[1] = v[2] = make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ __forceinline__ void load(const float* p, int l8) {
#pragma unroll
        for (int j = 0; j < 3; ++j) v[j] = *reinterpret_cast<const float4*>(p + 4 * (l8 + 8 * j));
    }
    // sum over the 4 row-groups of the warp, then lanes 0..7 add their 12 channels into dst[96]
    __device__ __forceinline__ void flush_atomic(float* dst, int lane) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float c[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                c[e] += __shfl_xor_sync(0xffffffffu, c[e], 8);
                c[e] += __shfl_xor_sync(0xffffffffu, c[e], 16);
                if (lane < 8) atomicAdd(dst + 4 * (lane + 8 * j) + e, c[e]);
            }
        }
    }
};

// ---------------------------------------------------------------- staging: fp32 rows -> (LayerNorm) -> 16-bit tile
// Writes rows [row_off, row_off+256) of the tile; rows t >= T are written as zeros.  gamma/beta in shared memory.
// Each warp takes 16-row blocks; one load/store instruction covers rows R + 4*sub + u (u = 0..3 unrolled), which spreads
// the 8-byte tile stores of a warp over all 32 banks twice (the minimum for 256 bytes).
// EXT: the row statistics come from `ext` ([T] (mean, rstd) of this slab's frames) instead of the row itself — GroupBatchNorm of
// NBC2 (models/arch/NBC2.py:111-145), whose statistics span all frequencies of a frame and are reduced by another kernel.
template <int FMT, bool LN, int U = 4, bool EXT = false>
__device__ __forceinline__ void stage_rows96(const float* __restrict__ xslab, int T, unsigned char* tile, int row_off,
                                             const float* s_gamma, const float* s_beta, int warp, int lane,
                                             float* stats_out = nullptr /* [T,2] (mean, rstd) of this slab */,
                                             int nwarps = 8, const float2* __restrict__ ext = nullptr) {
    const int sub = lane >> 3, l8 = lane & 7;
    Oct12 g, be;
    if (LN) { g.load(s_gamma, l8); be.load(s_beta, l8); }
    unsigned char* tl = tile + (size_t)(l8 >> 1) * kCS + (l8 & 1) * 8 + (size_t)row_off * 16;
#pragma unroll 1
    for (int R = 4 * U * warp; R < 256; R += 4 * U * nwarps) {
        float4 v[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = R + U * sub + u;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                v[u][j] = (r < T) ? __ldg(reinterpret_cast<const float4*>(xslab + (size_t)r * kH) + l8 + 8 * j)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int r = R + U * sub + u;
            const bool ok = r < T;
            if (LN) {
                float mean, rstd;
                if constexpr (EXT) {
                    const float2 st = ok ? __ldg(ext + r) : make_float2(0.f, 0.f);
                    mean = st.x;
                    rstd = st.y;
#pragma unroll
                    for (int j = 0; j < 3; ++j) v[u][j] = make_float4(v[u][j].x - mean, v[u][j].y - mean, v[u][j].z - mean, v[u][j].w - mean);
                } else {
                    mean = oct_sum(f4_hsum(v[u][0]) + f4_hsum(v[u][1]) + f4_hsum(v[u][2])) * (1.f / kH);
                    float q = 0.f;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        v[u][j] = make_float4(v[u][j].x - mean, v[u][j].y - mean, v[u][j].z - mean, v[u][j].w - mean);
                        q += f4_dot(v[u][j], v[u][j]);
                    }
                    rstd = rsqrtf(oct_sum(q) * (1.f / kH) + 1e-5f);
                }
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    v[u][j] = make_float4(v[u][j].x * rstd * g.v[j].x + be.v[j].x, v[u][j].y * rstd * g.v[j].y + be.v[j].y,
                                          v[u][j].z * rstd * g.v[j].z + be.v[j].z, v[u][j].w * rstd * g.v[j].w + be.v[j].w);
                if (stats_out && l8 == 0 && ok) *reinterpret_cast<float2*>(stats_out + 2 * r) = make_float2(mean, rstd);
            }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const uint2 p = ok ? make_uint2(pack16<FMT>(v[u][j].x, v[u][j].y), pack16<FMT>(v[u][j].z, v[u][j].w)) : make_uint2(0u, 0u);
                *reinterpret_cast<uint2*>(tl + (size_t)(4 * j) * kCS + r * 16) = p;
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

// LayerNorm backward + residual, eight lanes per frame (see above), fully coalesced.  `tile` holds d z (gradient wrt the
// LN output) as fp32 staged by the thread-per-frame TMEM epilogue with 4-float chunks: addr = tile + chunk*cs + (r + row_off)*16.
// Every lane accumulates d gamma / d beta of its 12 channels in registers across frames (and slabs): Oct12::flush_atomic.
__device__ __forceinline__ void ln_bwd_rows(const unsigned char* tile, uint32_t cs, int row_off, const float* __restrict__ xs,
                                            const float* __restrict__ dys, float* __restrict__ dxs,
                                            const float* __restrict__ stats, int T, const float* s_gamma /* smem */, Oct12& dg,
                                            Oct12& db, int warp, int lane, int nwarps) {
    const int sub = lane >> 3, l8 = lane & 7;
    const unsigned char* tl = tile + (size_t)l8 * cs + (size_t)row_off * 16;
#pragma unroll 1
    for (int R = 8 * warp; R < T; R += 8 * nwarps) {
        float4 xv[2][3], dv[2][3];
        float2 st[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = R + 2 * sub + u;
            const bool ok = r < T;
            st[u] = ok ? __ldg(reinterpret_cast<const float2*>(stats + 2 * r)) : make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                xv[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(xs + (size_t)r * kH) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                dv[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(dys + (size_t)r * kH) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = R + 2 * sub + u;
            const bool ok = r < T;
            float4 dz[3];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dz[j] = ok ? *reinterpret_cast<const float4*>(tl + (size_t)(8 * j) * cs + r * 16) : make_float4(0, 0, 0, 0);
                xv[u][j] = make_float4((xv[u][j].x - st[u].x) * st[u].y, (xv[u][j].y - st[u].x) * st[u].y,
                                       (xv[u][j].z - st[u].x) * st[u].y, (xv[u][j].w - st[u].x) * st[u].y);  // x hat (0 if !ok)
                dg.v[j] = make_float4(dg.v[j].x + dz[j].x * xv[u][j].x, dg.v[j].y + dz[j].y * xv[u][j].y,
                                      dg.v[j].z + dz[j].z * xv[u][j].z, dg.v[j].w + dz[j].w * xv[u][j].w);
                db.v[j] = make_float4(db.v[j].x + dz[j].x, db.v[j].y + dz[j].y, db.v[j].z + dz[j].z, db.v[j].w + dz[j].w);
                const float4 gj = *reinterpret_cast<const float4*>(s_gamma + 4 * (l8 + 8 * j));
                dz[j] = make_float4(dz[j].x * gj.x, dz[j].y * gj.y, dz[j].z * gj.z, dz[j].w * gj.w);
                s1 += f4_hsum(dz[j]);
                s2 += f4_dot(dz[j], xv[u][j]);
            }
            const float m1 = oct_sum(s1) * (1.f / kH), m2 = oct_sum(s2) * (1.f / kH);
            if (ok) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    reinterpret_cast<float4*>(dxs + (size_t)r * kH)[l8 + 8 * j] =
                        make_float4(dv[u][j].x + st[u].y * (dz[j].x - m1 - xv[u][j].x * m2), dv[u][j].y + st[u].y * (dz[j].y - m1 - xv[u][j].y * m2),
                                    dv[u][j].z + st[u].y * (dz[j].z - m1 - xv[u][j].z * m2), dv[u][j].w + st[u].y * (dz[j].w - m1 - xv[u][j].w * m2));
            }
        }
    }
}

// Residual add, eight lanes per frame, coalesced: y[r] = x[r] + branch[r], the branch staged as fp32 in `tile` with 4-float
// chunks (addr = tile + chunk*cs + (r + row_off)*16) by a thread-per-frame TMEM epilogue.
// SKIP4: the staging area is a tile whose every 4th chunk among the first 16 must stay untouched (zero padding of the
// per-head K tile in mhsa_fwd): four-float chunk c lives at tile chunk skip4_chunk(c).
__device__ __forceinline__ int skip4_chunk(int c) { return c + (c < 12 ? c / 3 : 4); }
// row_part (nullable): [T][2] (sum, sum of squares) over the 96 channels of every OUTPUT row — the per-slab partials of NBC2's
// GroupBatchNorm, reduced over the frequencies of a frame by gbn_reduce (ffn_fwd.cu).
template <bool SKIP4 = false>
__device__ __forceinline__ void add_rows(const unsigned char* tile, uint32_t cs, int row_off, const float* __restrict__ xs,
                                         float* __restrict__ ys, int T, int warp, int lane, int nwarps, float* __restrict__ row_part = nullptr) {
    const int sub = lane >> 3, l8 = lane & 7;
    const unsigned char* tl = tile + (size_t)row_off * 16;
#pragma unroll 1
    for (int R = 16 * warp; R < T; R += 16 * nwarps) {
        float4 xv[4][3];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = R + 4 * sub + u;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                xv[u][j] = (r < T) ? __ldg(reinterpret_cast<const float4*>(xs + (size_t)r * kH) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int r = R + 4 * sub + u;
            float ps = 0.f, pq = 0.f;
            if (r < T) {
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int ch = SKIP4 ? skip4_chunk(l8 + 8 * j) : l8 + 8 * j;
                    const float4 v = *reinterpret_cast<const float4*>(tl + (size_t)ch * cs + r * 16);
                    const float4 o = make_float4(xv[u][j].x + v.x, xv[u][j].y + v.y, xv[u][j].z + v.z, xv[u][j].w + v.w);
                    reinterpret_cast<float4*>(ys + (size_t)r * kH)[l8 + 8 * j] = o;
                    ps += f4_hsum(o);
                    pq += f4_dot(o, o);
                }
            }
            if (row_part) {  // uniform across the block
                ps = oct_sum(ps);
                pq = oct_sum(pq);
                if (l8 == 0 && r < T) *reinterpret_cast<float2*>(row_part + 2 * r) = make_float2(ps, pq);
            }
        }
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
