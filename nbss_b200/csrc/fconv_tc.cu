// fconv_tc.cu — cross-band frequency-convolution sub-block on tensor cores (tcgen05), forward and backward.
//
// Replaces SpatialNetLayer._fconv + residual (models/arch/SpatialNet.py:85,87,116-127; modules :36-40,:49-53):
//     y = x + PReLU( Conv1d_F(k=5, groups=8, zero 'same' padding)( LayerNorm_H(x) ) )
// The convolution runs along F for a fixed frame (b,t).  A CTA stacks the F rows of a few consecutive frames into ONE
// UMMA operand tile with two zero rows between frames (row p = 2 + slot*(F+2) + f): a conv tap is then a row-shifted
// view of the tile (descriptor start + 16 B * tap) and the zero gaps ARE the zero padding.  Weights are block-diagonal
// 48x48 tiles (4 groups of 12 channels), one per (channel half, tap).  fp16 operands, fp32 accumulation in TMEM.
//   forward : stage LN(x) -> 5 taps x 2 halves MMAs per 128-row tile -> epilogue (+bias, PReLU, +x) thread = (frame,f) row
//   backward: recompute conv; dc = dy * PReLU'(c) -> 16-bit tile; weight grad = dc^T x shifted LN(x) (MN-major MMAs,
//             accumulated per thread in registers across the CTA's frame groups); data grad = transposed conv of dc;
//             LayerNorm backward + residual in the epilogue; d(bias), d(slope), d(gamma), d(beta) by warp column sums.
#include "slab.cuh"

namespace nbss {

constexpr int kFQ = 2;                                   // channel halves (48 channels = 4 groups)
constexpr int kFTaps = 5;
constexpr uint32_t FC_IMG_TILE = 6 * 48 * 16;            // one [48 x 48] 16-bit tile, chunk-column, 4608 B
constexpr uint32_t FC_IMG_BYTES = kFQ * kFTaps * FC_IMG_TILE;  // 46080: forward image; the transposed image follows

// ------------------------------------------------------------------------------------------------ weight images
// fwd tile (half q, tap k): elem(n, kk) = W[48q+n][kk%12][k] if n/12 == kk/12 (N index = out channel, K = in channel)
// bwd tile              : elem(n, kk) = W[48q+kk][n%12][k] if n/12 == kk/12 (N index = in channel,  K = out channel)
__global__ void fconv_pack_kernel(const float* __restrict__ W, unsigned char* img, int fmt) {
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    if (chunk * 16 >= 2 * FC_IMG_BYTES) return;
    const bool bwd = chunk * 16 >= FC_IMG_BYTES;
    const uint32_t i = (chunk * 16 - (bwd ? FC_IMG_BYTES : 0)) / 16;
    const int tile = i / (6 * 48), r = i % (6 * 48), c = r / 48, n = r % 48, q = tile / kFTaps, k = tile % kFTaps;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int kk = 8 * c + j;
        const bool same = (n / 12) == (kk / 12);
        v[j] = !same ? 0.f : (bwd ? W[((48 * q + kk) * 12 + n % 12) * 5 + k] : W[((48 * q + n) * 12 + kk % 12) * 5 + k]);
    }
    uint4 o = (fmt == FMT_F16) ? make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]))
                               : make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(img + (size_t)chunk * 16) = o;
}

struct FcGeom {
    int B, F, T, FP;   // FP = F + 2 (frame period in tile rows)
    int nfr, nt;       // frames per group, 128-row M-tiles per group
    int rows;          // tile rows (128*nt + 9)
    uint32_t cs;       // chunk stride = rows * 16
    int groups_per_b, ngroups;
};

// ------------------------------------------------------------------------------------------------ row phases
// The fp32 row phases walk the group's nfr*F (frame, f) rows with eight lanes per row (slab.cuh): a warp handles 4*U rows
// per pass, row index i = R + U*sub + u.  nfr <= 4.
__device__ __forceinline__ void fc_row(const FcGeom& g, int i, int& tt, int& f) {
    tt = (i >= g.F) + (i >= 2 * g.F) + (i >= 3 * g.F);
    f = i - tt * g.F;
}

// LN(x) of the group's frames into the tile (16-bit); rows of frames beyond T are zero; optional (mean, rstd) per tile row;
// optionally the upstream gradient dy of the same rows into a second tile.
template <int FMT, int U, int NW = 8>
__device__ __forceinline__ void fc_stage(const FcGeom& g, const float* __restrict__ x, int b, int t0, unsigned char* tile,
                                         const float* s_lnw, const float* s_lnb, float2* s_stats, int warp, int lane,
                                         const float* __restrict__ dy = nullptr, unsigned char* dytile = nullptr) {
    const int sub = lane >> 3, l8 = lane & 7, N = g.nfr * g.F;
    Oct12 gw, gb;
    gw.load(s_lnw, l8);
    gb.load(s_lnb, l8);
    const size_t loff = (size_t)(l8 >> 1) * g.cs + (l8 & 1) * 8;
#pragma unroll 1
    for (int R = 4 * U * warp; R < N; R += 4 * U * NW) {
        float4 v[U][3], w[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = R + U * sub + u;
            int tt, f;
            fc_row(g, i, tt, f);
            const bool ok = i < N && t0 + tt < g.T;
            const float4* px = reinterpret_cast<const float4*>(x + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH) + l8;
            const float4* pd = reinterpret_cast<const float4*>((dy ? dy : x) + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH) + l8;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                v[u][j] = ok ? __ldg(px + 8 * j) : make_float4(0, 0, 0, 0);
                w[u][j] = (ok && dy) ? __ldg(pd + 8 * j) : make_float4(0, 0, 0, 0);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = R + U * sub + u;
            int tt, f;
            fc_row(g, i, tt, f);
            const bool tok = t0 + tt < g.T;
            const int p = 2 + tt * g.FP + f;
            const float mean = oct_sum(f4_hsum(v[u][0]) + f4_hsum(v[u][1]) + f4_hsum(v[u][2])) * (1.f / kH);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                v[u][j] = make_float4(v[u][j].x - mean, v[u][j].y - mean, v[u][j].z - mean, v[u][j].w - mean);
                q += f4_dot(v[u][j], v[u][j]);
            }
            const float rstd = rsqrtf(oct_sum(q) * (1.f / kH) + 1e-5f);
            if (i < N) {
                if (s_stats && l8 == 0) s_stats[p] = make_float2(mean, rstd);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    uint2 pk = make_uint2(0u, 0u);
                    if (tok)
                        pk = make_uint2(pack16<FMT>(v[u][j].x * rstd * gw.v[j].x + gb.v[j].x, v[u][j].y * rstd * gw.v[j].y + gb.v[j].y),
                                        pack16<FMT>(v[u][j].z * rstd * gw.v[j].z + gb.v[j].z, v[u][j].w * rstd * gw.v[j].w + gb.v[j].w));
                    *reinterpret_cast<uint2*>(tile + loff + (size_t)(4 * j) * g.cs + p * 16) = pk;
                    if (dytile)  // upstream gradient of the same row, 16-bit, same slot of the second tile
                        *reinterpret_cast<uint2*>(dytile + loff + (size_t)(4 * j) * g.cs + p * 16) =
                            make_uint2(pack16<FMT>(w[u][j].x, w[u][j].y), pack16<FMT>(w[u][j].z, w[u][j].w));
                }
            }
        }
    }
}

// rows of a second stream tensor (the upstream gradient) -> 16-bit tile, same row slots, no LayerNorm
template <int FMT, int U, int NW>
__device__ __forceinline__ void fc_stage_plain(const FcGeom& g, const float* __restrict__ dy, int b, int t0, unsigned char* tile,
                                               int warp, int lane) {
    const int sub = lane >> 3, l8 = lane & 7, N = g.nfr * g.F;
    const size_t loff = (size_t)(l8 >> 1) * g.cs + (l8 & 1) * 8;
#pragma unroll 1
    for (int R = 4 * U * warp; R < N; R += 4 * U * NW) {
        float4 w[U][3];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = R + U * sub + u;
            int tt, f;
            fc_row(g, i, tt, f);
            const bool ok = i < N && t0 + tt < g.T;
            const float4* pd = reinterpret_cast<const float4*>(dy + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH) + l8;
#pragma unroll
            for (int j = 0; j < 3; ++j) w[u][j] = ok ? __ldg(pd + 8 * j) : make_float4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = R + U * sub + u;
            int tt, f;
            fc_row(g, i, tt, f);
            if (i >= N) continue;
            const int p = 2 + tt * g.FP + f;
#pragma unroll
            for (int j = 0; j < 3; ++j)
                *reinterpret_cast<uint2*>(tile + loff + (size_t)(4 * j) * g.cs + p * 16) =
                    make_uint2(pack16<FMT>(w[u][j].x, w[u][j].y), pack16<FMT>(w[u][j].z, w[u][j].w));
        }
    }
}

// L2 prefetch of the NEXT group's rows (384 B each, F-strided): one cp.async.bulk.prefetch.L2 per row, dealt to the threads,
// so that the next iteration's latency-exposed staging loads hit L2 instead of HBM
__device__ __forceinline__ void fc_prefetch_rows(const FcGeom& g, const float* __restrict__ p0, const float* __restrict__ p1, int grp, int tid, int nthreads) {
    if (grp >= g.ngroups) return;
    const int b = grp / g.groups_per_b, t0 = (grp % g.groups_per_b) * g.nfr, N = g.nfr * g.F, tot = p1 ? 2 * N : N;
    for (int i = tid; i < tot; i += nthreads) {
        const int r = i < N ? i : i - N;
        int tt, f;
        fc_row(g, r, tt, f);
        if (t0 + tt < g.T) l2_prefetch((i < N ? p0 : p1) + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH, kH * 4);
    }
}

// conv (or transposed conv) MMAs of one group: D[tile m][half q] = sum_tap A(rows shifted) * Wimg(q, tap)
__device__ __forceinline__ void fc_conv_mmas(const FcGeom& g, uint32_t tmem, uint32_t tile_addr, uint32_t w_addr, uint32_t idesc,
                                             bool transposed, bool leader) {
    for (int m = 0; m < g.nt; ++m)
        for (int q = 0; q < kFQ; ++q)
            for (int tap = 0; tap < kFTaps; ++tap) {
                const int row = 128 * m + (transposed ? 4 - tap : tap);  // tile row of output q0: 2 + q0 + (tap-2) resp. 2 + q0 - (tap-2)
                mma_kk(tmem + m * 96 + q * 48, tile_addr + 6 * q * g.cs + row * 16, g.cs, w_addr + (q * kFTaps + tap) * FC_IMG_TILE, 768, 3,
                       idesc, tap > 0, leader);
            }
}

struct FcFwdArgs {
    const float* x;
    float* y;
    FcGeom g;
    const float *lnw, *lnb, *bias, *slope;
    const unsigned char* img;
    int* err;
};

template <int FMT>
__global__ void __launch_bounds__(256, 2) fconv_tc_fwd_kernel(FcFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const FcGeom g = a.g;
    unsigned char* tile = smem;
    unsigned char* wimg = smem + (size_t)12 * g.cs;
    float* cst = reinterpret_cast<float*>(wimg + FC_IMG_BYTES);  // lnw, lnb, bias, slope
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(cst + 384);
    uint64_t* bar_w = bar_mma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const uint32_t ncols = g.nt * 96 <= 256 ? 256u : 512u;  // 256 columns let two CTAs share an SM (one frame per group)
    if (warp == 0) tmem_alloc(tmem_slot, ncols);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        fence_mbar_init();
        load_image(wimg, a.img, FC_IMG_BYTES, bar_w);
    }
    for (int i = tid; i < 96; i += 256) { cst[i] = a.lnw[i]; cst[96 + i] = a.lnb[i]; cst[192 + i] = a.bias[i]; cst[288 + i] = a.slope[i]; }
    for (int i = tid; i < (int)(12 * g.cs / 16); i += 256) reinterpret_cast<uint4*>(tile)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t ta = smem_u32(tile), wa = smem_u32(wimg);
    const uint32_t id48 = make_idesc(FMT, 128, 48, 0, 0);
    const uint32_t lane_off = (uint32_t)(32 * (warp & 3)) << 16;
    uint32_t ph = 0;
    bool wready = false;
    stagger_start(17000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    for (int grp = blockIdx.x; grp < g.ngroups; grp += gridDim.x, ++it_) {
        const int b = grp / g.groups_per_b, t0 = (grp % g.groups_per_b) * g.nfr;
        NBSS_TICK(0, 0, it_);
        fc_prefetch_rows(g, a.x, nullptr, grp + gridDim.x, tid, 256);
        fc_stage<FMT, 5>(g, a.x, b, t0, tile, cst, cst + 96, nullptr, warp, lane);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 1, it_);
        if (warp == 0) {
            tc_fence_after();
            if (!wready) mbar_wait(bar_w, 0, a.err);
            const bool leader = elect_one();
            fc_conv_mmas(g, tmem, ta, wa, id48, false, leader);
            if (leader) umma_commit(bar_mma);
        }
        wready = true;
        __syncwarp();
        mbar_wait(bar_mma, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        NBSS_TICK(0, 2, it_);
        // epilogue 1: thread = one tile row (TMEM lane): PReLU(D + bias) -> 16-bit, written over the operand tile (dead
        // once every MMA of the group has completed) at the row's own slot.  warps 0-3 even tiles, warps 4-7 odd tiles.
        for (int m = warp >> 2; m < g.nt; m += 2) {
            const int q = 128 * m + 32 * (warp & 3) + lane;
            // gap rows between frames (and rows past the last frame) must stay zero: they are the conv's zero padding
            const float rmask = ((q % g.FP) < g.F && (q / g.FP) < g.nfr) ? 1.f : 0.f;
#pragma unroll 1
            for (int c0 = 0; c0 < kH; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem + lane_off + m * 96 + c0, r);
                tmem_ld_wait();
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float u = __uint_as_float(r[j]) + cst[192 + c0 + j];
                    v[j] = (u >= 0.f ? u : cst[288 + c0 + j] * u) * rmask;
                }
#pragma unroll
                for (int cc = 0; cc < 4; ++cc)
                    *reinterpret_cast<uint4*>(tile + (size_t)(c0 / 8 + cc) * g.cs + (2 + q) * 16) = pack8<FMT>(v + 8 * cc);
            }
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 3, it_);
        // epilogue 2: eight lanes per (frame, f) row, coalesced: y = x + branch
        {
            constexpr int U = 5;
            const int sub = lane >> 3, l8 = lane & 7, N = g.nfr * g.F;
            const size_t loff = (size_t)(l8 >> 1) * g.cs + (l8 & 1) * 8;
#pragma unroll 1
            for (int R = 4 * U * warp; R < N; R += 32 * U) {
                float4 xv[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = R + U * sub + u;
                    int tt, f;
                    fc_row(g, i, tt, f);
                    const bool ok = i < N && t0 + tt < g.T;
                    const float4* px = reinterpret_cast<const float4*>(a.x + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH) + l8;
#pragma unroll
                    for (int j = 0; j < 3; ++j) xv[u][j] = ok ? __ldg(px + 8 * j) : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = R + U * sub + u;
                    int tt, f;
                    fc_row(g, i, tt, f);
                    if (!(i < N && t0 + tt < g.T)) continue;
                    float4* py = reinterpret_cast<float4*>(a.y + (((size_t)b * g.F + f) * g.T + t0 + tt) * kH) + l8;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        const uint2 pk = *reinterpret_cast<const uint2*>(tile + loff + (size_t)(4 * j) * g.cs + (2 + tt * g.FP + f) * 16);
                        float b0, b1, b2, b3;
                        unpack16<FMT>(pk.x, b0, b1);
                        unpack16<FMT>(pk.y, b2, b3);
                        py[8 * j] = make_float4(xv[u][j].x + b0, xv[u][j].y + b1, xv[u][j].z + b2, xv[u][j].w + b3);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 4, it_);
    }
    if (warp == 0) tmem_dealloc(tmem, ncols);
}

struct FcBwdArgs {
    const float* x;
    const float* dy;
    float* dx;
    FcGeom g;
    const float *lnw, *lnb, *bias, *slope;
    const unsigned char* img;  // forward image followed by the transposed image
    float *dW, *dbias, *dslope, *dlnw, *dlnb;
    int* err;
};

constexpr int kFcBwdThreads = 512;  // warp w -> TMEM lane quarter w & 3, work group w >> 2 (tiles x column blocks, taps)

template <int FMT>
__global__ void __launch_bounds__(kFcBwdThreads, 1) fconv_tc_bwd_kernel(FcBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NT = kFcBwdThreads, NW = NT / 32;
    const FcGeom g = a.g;
    unsigned char* htile = smem;                           // LN(x)
    unsigned char* gtile = smem + (size_t)12 * g.cs;       // dc
    unsigned char* wimg = gtile + (size_t)12 * g.cs;       // weight slot (fwd image, then transposed image); also the
                                                           // over-read area of the 128-feature MN-major window of gtile
    float* cst = reinterpret_cast<float*>(wimg + FC_IMG_BYTES);  // lnw, lnb, bias, slope (384) + acc (384): dbias, dslope, dlnw, dlnb
    float* acc = cst + 384;
    float* accdw = acc + 384;                               // [96][60] weight-gradient accumulators (thread-owned rows)
    float2* stats = reinterpret_cast<float2*>(accdw + 96 * 60);  // [rows] (mean, rstd)
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(stats + g.rows);
    uint64_t* bar_w = bar_mma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const int q4 = warp & 3, gq = warp >> 2;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 96; i += NT) { cst[i] = a.lnw[i]; cst[96 + i] = a.lnb[i]; cst[192 + i] = a.bias[i]; cst[288 + i] = a.slope[i]; }
    for (int i = tid; i < 384 + 96 * 60; i += NT) acc[i] = 0.f;
    for (int i = tid; i < (int)(24 * g.cs / 16); i += NT) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t ha = smem_u32(htile), ga = smem_u32(gtile), wa = smem_u32(wimg);
    const uint32_t id48 = make_idesc(FMT, 128, 48, 0, 0);
    const uint32_t id_wg = make_idesc(FMT, 128, 96, 1, 1);
    const uint32_t lane_off = (uint32_t)(32 * q4) << 16;
    uint32_t ph = 0, ph_w = 0;

    auto wait_mma = [&]() {
        __syncwarp();
        mbar_wait(bar_mma, ph, a.err);
        ph ^= 1;
        tc_fence_after();
    };
    stagger_start(60000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    for (int grp = blockIdx.x; grp < g.ngroups; grp += gridDim.x, ++it_) {
        const int b = grp / g.groups_per_b, t0 = (grp % g.groups_per_b) * g.nfr;
        NBSS_TICK(1, 0, it_);
        if (tid == 0) load_image(wimg, a.img, FC_IMG_BYTES, bar_w);
        // (staging dy first moves ~1.4 k cycles from the staging passes into E-A: the item's total is unchanged, see
        // profiles/r02n_fconv_stage_order.txt)
        fc_stage<FMT, 5, NW>(g, a.x, b, t0, htile, cst, cst + 96, stats, warp, lane);
        NBSS_TICK(1, 1, it_);
        fc_stage_plain<FMT, 5, NW>(g, a.dy, b, t0, gtile, warp, lane);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 2, it_);
        // ---- P1: recompute the conv
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            fc_conv_mmas(g, tmem, ha, wa, id48, false, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(1, 3, it_);
        if (tid == 0) load_image(wimg, a.img + FC_IMG_BYTES, FC_IMG_BYTES, bar_w);  // transposed image for the data gradient
        // next group's x / dy rows -> L2, issued after this group's own staging loads have landed (at the top of the group the
        // prefetch competed with them and cost 5 %)
        fc_prefetch_rows(g, a.x, a.dy, grp + gridDim.x, tid, NT);
        // ---- E-A: dc = dy * PReLU'(c), in place over the staged dy in gtile; column sums for dbias / dslope.
        //      work items (M-tile, 16-column block) are dealt to the four warp groups
#pragma unroll 1
        for (int it = gq; it < g.nt * 6; it += 4) {
            const int m = it / 6, c0 = 16 * (it - 6 * m);
            const int q = 128 * m + 32 * q4 + lane;
            unsigned char* grow_ = gtile + (size_t)(2 + q) * 16;
            uint32_t r[16];
            tmem_ld16(tmem + lane_off + m * 96 + c0, r);
            tmem_ld_wait();
            float dc[16], ds[16];
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const uint4 pk = *reinterpret_cast<const uint4*>(grow_ + (size_t)(c0 / 8 + cc) * g.cs);
                float dys[8];
                unpack16<FMT>(pk.x, dys[0], dys[1]);
                unpack16<FMT>(pk.y, dys[2], dys[3]);
                unpack16<FMT>(pk.z, dys[4], dys[5]);
                unpack16<FMT>(pk.w, dys[6], dys[7]);
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = c0 + 8 * cc + e;
                    const float cv = __uint_as_float(r[8 * cc + e]) + cst[192 + c];
                    dc[8 * cc + e] = dys[e] * (cv >= 0.f ? 1.f : cst[288 + c]);  // gap / missing rows hold zeros in gtile
                    ds[8 * cc + e] = cv < 0.f ? dys[e] * cv : 0.f;
                }
            }
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) *reinterpret_cast<uint4*>(grow_ + (size_t)(c0 / 8 + cc) * g.cs) = pack8<FMT>(dc + 8 * cc);
            const float sb = warp_colsum16(dc, lane), ss = warp_colsum16(ds, lane);  // lane pair (2k, 2k+1) holds column k
            if (!(lane & 1)) {
                atomicAdd(acc + c0 + (lane >> 1), sb);
                atomicAdd(acc + 96 + c0 + (lane >> 1), ss);
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 4, it_);
        // ---- P2: weight gradient  dW[co, ci, tap] += sum_q dc[q, co] * h[q + tap - 2, ci]   (MN-major x MN-major)
        if (warp == 0) {  // the whole warp runs the (uniform) descriptor arithmetic, one elected lane issues
            tc_fence_after();
            const bool leader = elect_one();
            const int nks = 8 * g.nt;  // 16 rows per k-step
            for (int tap = 0; tap < kFTaps; ++tap) {
                uint64_t da = sdesc_mnmajor(ga + 2 * 16, g.cs), db = sdesc_mnmajor(ha + tap * 16, g.cs);
                for (int ks = 0; ks < nks; ++ks) {
                    if (leader) umma_f16(tmem + 96 * tap, da, db, id_wg, ks ? 1u : 0u);
                    da += 16;  // 16 rows x 16 B >> 4 in the start-address field
                    db += 16;
                }
            }
            if (leader) umma_commit(bar_mma);
        }
        wait_mma();
        NBSS_TICK(1, 5, it_);
        if (q4 < 3) {
            // thread = out channel co (TMEM lane), taps dealt to the warp groups; it keeps the 12 in-channel columns
            // [12*(co/12), +12) of each tap.  tcgen05.ld takes ONE column address per warp, so every warp loads a uniform
            // 48-column window [24*q4, 24*q4+48) that contains the groups of its 32 channels and each lane selects its slice.
            const int co = 32 * q4 + lane;
            const int ws = 24 * q4;
            const int off = 12 * (co / 12) - ws;  // 0, 12, 24 or 36
#pragma unroll 1
            for (int tap = gq; tap < kFTaps; tap += 4) {
                uint32_t r[48];
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    uint32_t t8[8];
                    tmem_ld8(tmem + lane_off + 96 * tap + ws + 8 * k, t8);
#pragma unroll
                    for (int j = 0; j < 8; ++j) r[8 * k + j] = t8[j];
                }
                tmem_ld_wait();
#pragma unroll
                for (int ci = 0; ci < 12; ++ci) {
                    const uint32_t v = off == 0 ? r[ci] : (off == 12 ? r[12 + ci] : (off == 24 ? r[24 + ci] : r[36 + ci]));
                    accdw[co * 60 + ci * 5 + tap] += __uint_as_float(v);  // (co, tap) is owned by exactly one thread
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 6, it_);
        // ---- P3: data gradient of the conv
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            fc_conv_mmas(g, tmem, ga, wa, id48, true, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(1, 7, it_);
        // ---- E-B1: thread = tile row: d h (data gradient of the conv) -> 16-bit, over htile (dead after the weight
        //      gradient MMAs); gap rows are written as zeros so they keep acting as the next group's zero padding
#pragma unroll 1
        for (int it = gq; it < g.nt * 6; it += 4) {
            const int m = it / 6, c0 = 16 * (it - 6 * m);
            const int q = 128 * m + 32 * q4 + lane;
            const float rmask = ((q % g.FP) < g.F && (q / g.FP) < g.nfr) ? 1.f : 0.f;
            uint32_t r[16];
            tmem_ld16(tmem + lane_off + m * 96 + c0, r);
            tmem_ld_wait();
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]) * rmask;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc)
                *reinterpret_cast<uint4*>(htile + (size_t)(c0 / 8 + cc) * g.cs + (2 + q) * 16) = pack8<FMT>(v + 8 * cc);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 8, it_);
        // ---- E-B2: eight lanes per (frame, f) row, coalesced: LayerNorm backward + residual; d gamma / d beta per lane
        {
            constexpr int U = 3;
            const int sub = lane >> 3, l8 = lane & 7, N = g.nfr * g.F;
            const size_t loff = (size_t)(l8 >> 1) * g.cs + (l8 & 1) * 8;
            Oct12 dlg, dlb;  // this group's d gamma / d beta of the lane's 12 channels -> shared accumulators
            dlg.zero();
            dlb.zero();
#pragma unroll 1
            for (int R = 4 * U * warp; R < N; R += 4 * U * NW) {
                float4 xv[U][3];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = R + U * sub + u;
                    int tt, f;
                    fc_row(g, i, tt, f);
                    const bool ok = i < N && t0 + tt < g.T;
                    const size_t off = (((size_t)b * g.F + f) * g.T + t0 + tt) * kH;
#pragma unroll
                    for (int j = 0; j < 3; ++j)
                        xv[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(a.x + off) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int i = R + U * sub + u;
                    int tt, f;
                    fc_row(g, i, tt, f);
                    const bool ok = i < N && t0 + tt < g.T;
                    const int p = 2 + tt * g.FP + f;
                    const float2 st = ok ? stats[p] : make_float2(0.f, 0.f);
                    const size_t off = (((size_t)b * g.F + f) * g.T + t0 + tt) * kH;
                    float4 dh[3], dv[3];  // dy of this row was staged (and L2-prefetched) by this CTA: an L2 hit
#pragma unroll
                    for (int j = 0; j < 3; ++j) dv[j] = ok ? __ldg(reinterpret_cast<const float4*>(a.dy + off) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int j = 0; j < 3; ++j) {
                        dh[j] = make_float4(0, 0, 0, 0);
                        if (ok) {
                            const uint2 pk = *reinterpret_cast<const uint2*>(htile + loff + (size_t)(4 * j) * g.cs + p * 16);
                            unpack16<FMT>(pk.x, dh[j].x, dh[j].y);
                            unpack16<FMT>(pk.y, dh[j].z, dh[j].w);
                        }
                        xv[u][j] = make_float4((xv[u][j].x - st.x) * st.y, (xv[u][j].y - st.x) * st.y, (xv[u][j].z - st.x) * st.y,
                                               (xv[u][j].w - st.x) * st.y);  // x hat (0 for rows that are not ok)
                        dlg.v[j] = make_float4(dlg.v[j].x + dh[j].x * xv[u][j].x, dlg.v[j].y + dh[j].y * xv[u][j].y,
                                               dlg.v[j].z + dh[j].z * xv[u][j].z, dlg.v[j].w + dh[j].w * xv[u][j].w);
                        dlb.v[j] = make_float4(dlb.v[j].x + dh[j].x, dlb.v[j].y + dh[j].y, dlb.v[j].z + dh[j].z, dlb.v[j].w + dh[j].w);
                        const float4 gj = *reinterpret_cast<const float4*>(cst + 4 * (l8 + 8 * j));
                        dh[j] = make_float4(dh[j].x * gj.x, dh[j].y * gj.y, dh[j].z * gj.z, dh[j].w * gj.w);
                        s1 += f4_hsum(dh[j]);
                        s2 += f4_dot(dh[j], xv[u][j]);
                    }
                    const float m1 = oct_sum(s1) * (1.f / kH), m2 = oct_sum(s2) * (1.f / kH);
                    if (ok) {
                        float4* pdx = reinterpret_cast<float4*>(a.dx + off) + l8;
#pragma unroll
                        for (int j = 0; j < 3; ++j)
                            pdx[8 * j] = make_float4(dv[j].x + st.y * (dh[j].x - m1 - xv[u][j].x * m2), dv[j].y + st.y * (dh[j].y - m1 - xv[u][j].y * m2),
                                                     dv[j].z + st.y * (dh[j].z - m1 - xv[u][j].z * m2), dv[j].w + st.y * (dh[j].w - m1 - xv[u][j].w * m2));
                    }
                }
            }
            dlg.flush_atomic(acc + 192, lane);
            dlb.flush_atomic(acc + 288, lane);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 9, it_);
    }
    // ---- flush parameter gradients: one set of global atomics per CTA
    for (int i = tid; i < 96 * 60; i += NT) atomicAdd(a.dW + i, accdw[i]);
    for (int i = tid; i < 96; i += NT) {
        atomicAdd(a.dbias + i, acc[i]);
        atomicAdd(a.dslope + i, acc[96 + i]);
        atomicAdd(a.dlnw + i, acc[192 + i]);
        atomicAdd(a.dlnb + i, acc[288 + i]);
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static int fc_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

static bool fc_geom(FcGeom& g, int B, int F, int T, int max_tiles, int max_frames) {
    g.B = B; g.F = F; g.T = T; g.FP = F + 2;
    int nfr = (128 * max_tiles) / g.FP;
    if (nfr > max_frames) nfr = max_frames;
    if (nfr < 1) return false;
    g.nfr = nfr;
    g.nt = (nfr * g.FP + 127) / 128;
    g.rows = 128 * g.nt + 9;
    g.cs = (uint32_t)g.rows * 16;
    g.groups_per_b = (T + nfr - 1) / nfr;
    g.ngroups = B * g.groups_per_b;
    return true;
}

}  // namespace nbss

NBSS_PHASE_READER(nbss_debug_phases_fconv_tc)

using namespace nbss;

extern "C" unsigned int nbss_fconv_image_bytes() { return 2 * FC_IMG_BYTES; }

// W: fconv{1,2}.1.weight [96,12,5] -> forward + transposed UMMA images (nbss_fconv_image_bytes() bytes)
extern "C" int nbss_fconv_pack(const float* W, void* img, int fmt, void* stream) {
    if (!W || !img) return NBSS_ERR_NULL;
    const int n = 2 * FC_IMG_BYTES / 16;
    fconv_pack_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, (unsigned char*)img, fmt);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_fconv_tc_fwd(const float* x, float* y, int B, int F, int T, const float* lnw, const float* lnb,
                                 const float* bias, const float* slope, const void* img, int fmt, int* err, void* stream) {
    if (!x || !y || !lnw || !lnb || !bias || !slope || !img) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    FcFwdArgs a{x, y, {}, lnw, lnb, bias, slope, (const unsigned char*)img, err};
    // one frame (2 M-tiles) per group: 98 KB of shared memory and 256 TMEM columns, so two CTAs are resident per SM and
    // one stages / stores while the other issues MMAs
    if (!fc_geom(a.g, B, F, T, 2, 1) && !fc_geom(a.g, B, F, T, 5, 4)) return NBSS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)12 * a.g.cs + FC_IMG_BYTES + 384 * 4 + 64;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    auto kern = (fmt == FMT_F16) ? fconv_tc_fwd_kernel<FMT_F16> : fconv_tc_fwd_kernel<FMT_BF16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int per_sm = (smem <= 110 * 1024 && a.g.nt * 96 <= 256) ? 2 : 1;
    const int grid = a.g.ngroups < per_sm * fc_sms() ? a.g.ngroups : per_sm * fc_sms();
    kern<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_fconv_tc_bwd(const float* x, const float* dy, float* dx, int B, int F, int T, const float* lnw,
                                 const float* lnb, const float* bias, const float* slope, const void* img, float* dW,
                                 float* dbias, float* dslope, float* dlnw, float* dlnb, int fmt, int* err, void* stream) {
    if (!x || !dy || !dx || !lnw || !lnb || !bias || !slope || !img || !dW || !dbias || !dslope || !dlnw || !dlnb) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    FcBwdArgs a{x, dy, dx, {}, lnw, lnb, bias, slope, (const unsigned char*)img, dW, dbias, dslope, dlnw, dlnb, err};
    if (!fc_geom(a.g, B, F, T, 3, 2)) return NBSS_ERR_UNSUPPORTED;
    const size_t smem = (size_t)24 * a.g.cs + FC_IMG_BYTES + (768 + 96 * 60) * 4 + (size_t)a.g.rows * 8 + 64;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    auto kern = (fmt == FMT_F16) ? fconv_tc_bwd_kernel<FMT_F16> : fconv_tc_bwd_kernel<FMT_BF16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int grid = a.g.ngroups < fc_sms() ? a.g.ngroups : fc_sms();
    kern<<<grid, kFcBwdThreads, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
