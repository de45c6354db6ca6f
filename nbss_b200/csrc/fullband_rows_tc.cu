// fullband_rows_tc.cu — the per-point halves of the full-band sub-block on tensor cores (tcgen05):
//     squeeze   : s[b,t,g,f] = SiLU(Wsq LN(x[b,f,t,:]) + bsq)         (SpatialNet.py:129-137, modules :41-44)
//     unsqueeze : y[b,f,t,:] = x + SiLU(Wun u[b,t,:,f] + bun)          (SpatialNet.py:143-146, modules :45-48)
// and their backward passes.  (The F x F LinearGroup between them is fullband_tc.cu.)
//
// Work unit: a tile of 128 T-F points.  "Main" tiles are 128 consecutive frequencies of one (b,t) frame, so that the
// [B,T,8,F] tensors s / u / ds / du are read and written with f contiguous across the lanes of a warp; the F mod 128
// remaining frequencies (F = 129: one) of 128/nrem consecutive frames form "leftover" tiles.  A per-tile table in shared
// memory holds each row's offset into the [B,F,T,96] stream tensors and into the [B,T,8,F] tensors.
//
// Per tile: the 96-channel rows are staged as a 16-bit K-major operand tile by the eight-lanes-per-row phases of slab.cuh
// (LayerNorm fused); the 96<->8 maps are single UMMAs against tiny weight images built in shared memory (N = 16 or
// K = 16, zero padded); activations are applied by thread-per-row TMEM epilogues; the 96x8 weight gradients (and the bias
// gradients, through a ones feature) accumulate in TMEM across all tiles of the CTA as MN-major x MN-major UMMAs.
// Small footprints (45-65 KB of shared memory, 32-128 TMEM columns) keep 3-4 CTAs resident per SM, so one CTA's loads and
// epilogues overlap the others' MMAs.
#define NBSS_SILU_EXACT  // two-MUFU sigmoid (common.cuh): the unsqueeze SiLU output goes straight into the fp32 stream (no 16-bit rounding to hide behind)
#include "slab.cuh"

namespace nbss {

constexpr int kRtRows = 128;
constexpr uint32_t kRtCs = (kRtRows + 1) * 16;  // 2064: chunk stride of every row tile here (16 mod 128)
constexpr int kSq = 8;                          // dim_squeeze

struct RtGeom {
    int B, F, T, M;
    int nfb, nrem;      // full 128-frequency blocks per frame, remaining frequencies
    int n_main, ntiles;
};
static RtGeom rt_geom(int B, int F, int T) {
    RtGeom g;
    g.B = B; g.F = F; g.T = T; g.M = B * T;
    g.nfb = F / kRtRows; g.nrem = F % kRtRows;
    g.n_main = g.M * g.nfb;
    g.ntiles = g.n_main + (g.nrem ? (int)(((long long)g.M * g.nrem + kRtRows - 1) / kRtRows) : 0);
    return g;
}

// row r of tile -> offsets (elements) of its stream row [B,F,T,96] and of its (g = 0) element in a [B,T,8,F] tensor; -1: no row
__device__ __forceinline__ void rt_table(const RtGeom& g, int tile, long long* rowx, long long* rowq, int tid) {
    if (tid < kRtRows) {
        long long m;
        int f;
        if (tile < g.n_main) {
            m = tile / g.nfb;
            f = kRtRows * (tile % g.nfb) + tid;
        } else {
            const long long i = (long long)(tile - g.n_main) * kRtRows + tid;
            m = i / g.nrem;
            f = kRtRows * g.nfb + (int)(i % g.nrem);
        }
        const bool ok = m < g.M;
        const long long b = m / g.T, t = m % g.T;
        rowx[tid] = ok ? ((b * g.F + f) * g.T + t) * kH : -1;
        rowq[tid] = ok ? m * kSq * g.F + f : -1;
    }
}

// K-major weight image [N rows x Kp features] (chunk stride N*16) from a functor w(n, k)
template <int FMT, typename Wfn>
__device__ __forceinline__ void rt_build_img(unsigned char* img, int N, int Kp, Wfn w, int tid, int nthreads) {
    for (int u = tid; u < N * (Kp / 8); u += nthreads) {
        const int c = u / N, n = u - c * N;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = w(n, 8 * c + j);
        *reinterpret_cast<uint4*>(img + (size_t)c * N * 16 + n * 16) = pack8<FMT>(v);
    }
}

// 128 rows (8 warps x 16 rows, eight lanes per row) -> 16-bit tile chunks 0..11; optional LayerNorm (+ statistics)
template <int FMT, bool LN>
__device__ __forceinline__ void rt_stage(const float* __restrict__ x, const long long* rowx, unsigned char* tile,
                                         const float* s_gamma, const float* s_beta, float2* stats, int warp, int lane) {
    const int sub = lane >> 3, l8 = lane & 7, R = 16 * warp;
    Oct12 g, be;
    if (LN) { g.load(s_gamma, l8); be.load(s_beta, l8); }
    unsigned char* tl = tile + (size_t)(l8 >> 1) * kRtCs + (l8 & 1) * 8;
    float4 v[4][3];
    bool okv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long off = rowx[R + 4 * sub + u];
        okv[u] = off >= 0;
        const float4* p = reinterpret_cast<const float4*>(x + (okv[u] ? off : 0)) + l8;
#pragma unroll
        for (int j = 0; j < 3; ++j) v[u][j] = okv[u] ? __ldg(p + 8 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int r = R + 4 * sub + u;
        if (LN) {
            const float mean = oct_sum(f4_hsum(v[u][0]) + f4_hsum(v[u][1]) + f4_hsum(v[u][2])) * (1.f / kH);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                v[u][j] = make_float4(v[u][j].x - mean, v[u][j].y - mean, v[u][j].z - mean, v[u][j].w - mean);
                q += f4_dot(v[u][j], v[u][j]);
            }
            const float rstd = rsqrtf(oct_sum(q) * (1.f / kH) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 3; ++j)
                v[u][j] = make_float4(v[u][j].x * rstd * g.v[j].x + be.v[j].x, v[u][j].y * rstd * g.v[j].y + be.v[j].y,
                                      v[u][j].z * rstd * g.v[j].z + be.v[j].z, v[u][j].w * rstd * g.v[j].w + be.v[j].w);
            if (stats && l8 == 0) stats[r] = make_float2(mean, rstd);
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const uint2 p = okv[u] ? make_uint2(pack16<FMT>(v[u][j].x, v[u][j].y), pack16<FMT>(v[u][j].z, v[u][j].w)) : make_uint2(0u, 0u);
            *reinterpret_cast<uint2*>(tl + (size_t)(4 * j) * kRtCs + r * 16) = p;
        }
    }
}

// y[row] = x[row] + branch[row]; branch staged as fp32 4-float chunks: stage + chunk*kRtCs + r*16
__device__ __forceinline__ void rt_add(const unsigned char* stage, const long long* rowx, const float* __restrict__ x,
                                       float* __restrict__ y, int warp, int lane) {
    const int sub = lane >> 3, l8 = lane & 7, R = 16 * warp;
    float4 xv[4][3];
    long long off[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        off[u] = rowx[R + 4 * sub + u];
        const float4* p = reinterpret_cast<const float4*>(x + (off[u] >= 0 ? off[u] : 0)) + l8;
#pragma unroll
        for (int j = 0; j < 3; ++j) xv[u][j] = off[u] >= 0 ? __ldg(p + 8 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (off[u] < 0) continue;
        const int r = R + 4 * sub + u;
        float4* py = reinterpret_cast<float4*>(y + off[u]) + l8;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float4 v = *reinterpret_cast<const float4*>(stage + (size_t)(l8 + 8 * j) * kRtCs + r * 16);
            py[8 * j] = make_float4(xv[u][j].x + v.x, xv[u][j].y + v.y, xv[u][j].z + v.z, xv[u][j].w + v.w);
        }
    }
}

// LayerNorm backward + residual: dx[row] = dy[row] + LN'(d ln staged as fp32 4-float chunks); d gamma / d beta per lane
__device__ __forceinline__ void rt_ln_bwd(const unsigned char* stage, const long long* rowx, const float* __restrict__ x,
                                          const float* __restrict__ dy, float* __restrict__ dx, const float2* stats,
                                          const float* s_gamma, Oct12& dg, Oct12& db, int warp, int lane) {
    const int sub = lane >> 3, l8 = lane & 7;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        const int R = 16 * warp + 8 * pass;
        float4 xv[2][3], dv[2][3];
        long long off[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            off[u] = rowx[R + 2 * sub + u];
            const bool ok = off[u] >= 0;
            const float4* px = reinterpret_cast<const float4*>(x + (ok ? off[u] : 0)) + l8;
            const float4* pd = reinterpret_cast<const float4*>(dy + (ok ? off[u] : 0)) + l8;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                xv[u][j] = ok ? __ldg(px + 8 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                dv[u][j] = ok ? __ldg(pd + 8 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int r = R + 2 * sub + u;
            const bool ok = off[u] >= 0;
            const float2 st = ok ? stats[r] : make_float2(0.f, 0.f);
            float4 dz[3];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                dz[j] = ok ? *reinterpret_cast<const float4*>(stage + (size_t)(l8 + 8 * j) * kRtCs + r * 16) : make_float4(0.f, 0.f, 0.f, 0.f);
                xv[u][j] = make_float4((xv[u][j].x - st.x) * st.y, (xv[u][j].y - st.x) * st.y, (xv[u][j].z - st.x) * st.y,
                                       (xv[u][j].w - st.x) * st.y);  // x hat (0 for missing rows)
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
                float4* pdx = reinterpret_cast<float4*>(dx + off[u]) + l8;
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    pdx[8 * j] = make_float4(dv[u][j].x + st.y * (dz[j].x - m1 - xv[u][j].x * m2), dv[u][j].y + st.y * (dz[j].y - m1 - xv[u][j].y * m2),
                                             dv[u][j].z + st.y * (dz[j].z - m1 - xv[u][j].z * m2), dv[u][j].w + st.y * (dz[j].w - m1 - xv[u][j].w * m2));
            }
        }
    }
}

// D[128 x 96] (TMEM columns col0..col0+95) + bias -> (optional SiLU) -> fp32 staging; thread = (row, column half)
template <bool SILU>
__device__ __forceinline__ void rt_d96_to_stage(uint32_t tmem, uint32_t col0, unsigned char* stage, const float* s_bias, int warp, int lane) {
    const int q4 = warp & 3, hf = warp >> 2, r = 32 * q4 + lane;
    const uint32_t tacc = tmem + ((uint32_t)(32 * q4) << 16) + col0;
#pragma unroll 1
    for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tacc + c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float a = __uint_as_float(v[4 * j4 + e]) + (s_bias ? s_bias[c0 + 4 * j4 + e] : 0.f);
                o[e] = SILU ? silu(a) : a;
            }
            *reinterpret_cast<float4*>(stage + (size_t)(c0 / 4 + j4) * kRtCs + r * 16) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

struct RtArgs {
    RtGeom g;
    const float* x;     // stream [B,F,T,96]
    const float* dy;    // upstream gradient (backward kernels)
    float* y;           // forward: y; backward: dx
    const float* q_in;  // [B,T,8,F] input  (u for unsqueeze, ds for squeeze backward)
    float* q_out;       // [B,T,8,F] output (s for squeeze, du for unsqueeze backward)
    const float *lnw, *lnb, *W, *bias;  // LayerNorm affine (squeeze), W = Wsq [8,96] or Wun [96,8], its bias
    float *dlnw, *dlnb, *dW, *dbias;
    int* err;
};

constexpr uint32_t RT_IMG = 3072;  // one 16x96 or 96x16 image

// ------------------------------------------------------------------------------------------------ squeeze forward
constexpr uint32_t SQF_X = 0;                        // LN(x) tile, 12 chunks
constexpr uint32_t SQF_IMG = SQF_X + 12 * kRtCs;     // Wsq image [16 x 96]
constexpr uint32_t SQF_CST = SQF_IMG + RT_IMG;       // lnw 96, lnb 96, bsq 8
constexpr uint32_t SQF_TAB = SQF_CST + 208 * 4;      // rowx, rowq
constexpr uint32_t SQF_BAR = SQF_TAB + 2 * kRtRows * 8;
constexpr uint32_t SQF_SMEM = SQF_BAR + 32;

template <int FMT>
__global__ void __launch_bounds__(256, 3) squeeze_fwd_tc_kernel(RtArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const RtGeom g = a.g;
    unsigned char* xt = smem + SQF_X;
    unsigned char* img = smem + SQF_IMG;
    float* cst = reinterpret_cast<float*>(smem + SQF_CST);
    long long* rowx = reinterpret_cast<long long*>(smem + SQF_TAB);
    long long* rowq = rowx + kRtRows;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SQF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 32);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 96; i += 256) { cst[i] = a.lnw[i]; cst[96 + i] = a.lnb[i]; }
    if (tid < kSq) cst[192 + tid] = a.bias[tid];
    { const float* W = a.W; rt_build_img<FMT>(img, 16, 96, [W](int n, int k) { return n < kSq ? W[n * kH + k] : 0.f; }, tid, 256); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t xa = smem_u32(xt), ia = smem_u32(img);
    const uint32_t id16 = make_idesc(FMT, 128, 16, 0, 0);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        rt_table(g, tile, rowx, rowq, tid);
        __syncthreads();
        rt_stage<FMT, true>(a.x, rowx, xt, cst, cst + 96, nullptr, warp, lane);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem, xa, kRtCs, ia, 256, 6, id16, 0, leader);
            if (leader) umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        if (warp < 4) {
            uint32_t v[8];
            tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16), v);
            tmem_ld_wait();
            const long long q = rowq[32 * warp + lane];
            if (q >= 0) {
#pragma unroll
                for (int j = 0; j < kSq; ++j) a.q_out[q + (long long)j * g.F] = silu(__uint_as_float(v[j]) + cst[192 + j]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 32);
}

// ------------------------------------------------------------------------------------------------ unsqueeze forward
constexpr uint32_t UNF_U = 0;                        // u tile, 2 chunks
constexpr uint32_t UNF_ST = UNF_U + 2 * kRtCs;       // fp32 staging, 24 four-float chunks
constexpr uint32_t UNF_IMG = UNF_ST + 24 * kRtCs;    // Wun image [96 x 16]
constexpr uint32_t UNF_CST = UNF_IMG + RT_IMG;       // bun 96
constexpr uint32_t UNF_TAB = UNF_CST + 96 * 4;
constexpr uint32_t UNF_BAR = UNF_TAB + 2 * kRtRows * 8;
constexpr uint32_t UNF_SMEM = UNF_BAR + 32;

// u[b,t,0..7,f] of the tile's rows -> chunk 0 of a 16-bit tile; chunk 1 = (ones ? 1 : 0, 0, ...) for rows that exist
template <int FMT>
__device__ __forceinline__ void rt_gather8(const float* __restrict__ q_in, const long long* rowq, int F, unsigned char* tile, bool ones,
                                           int tid) {
    if (tid < kRtRows) {
        const long long q = rowq[tid];
        float v[8];
#pragma unroll
        for (int j = 0; j < kSq; ++j) v[j] = q >= 0 ? __ldg(q_in + q + (long long)j * F) : 0.f;
        *reinterpret_cast<uint4*>(tile + tid * 16) = pack8<FMT>(v);
        float o[8] = {(ones && q >= 0) ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4*>(tile + kRtCs + tid * 16) = pack8<FMT>(o);
    }
}

template <int FMT>
__global__ void __launch_bounds__(256, 3) unsqueeze_fwd_tc_kernel(RtArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const RtGeom g = a.g;
    unsigned char* ut = smem + UNF_U;
    unsigned char* stage = smem + UNF_ST;
    unsigned char* img = smem + UNF_IMG;
    float* cst = reinterpret_cast<float*>(smem + UNF_CST);
    long long* rowx = reinterpret_cast<long long*>(smem + UNF_TAB);
    long long* rowq = rowx + kRtRows;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + UNF_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 96; i += 256) cst[i] = a.bias[i];
    { const float* W = a.W; rt_build_img<FMT>(img, 96, 16, [W](int n, int k) { return k < kSq ? W[n * kSq + k] : 0.f; }, tid, 256); }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t ua = smem_u32(ut), ia = smem_u32(img);
    const uint32_t id96 = make_idesc(FMT, 128, 96, 0, 0);
    uint32_t ph = 0;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        rt_table(g, tile, rowx, rowq, tid);
        __syncthreads();
        rt_gather8<FMT>(a.q_in, rowq, g.F, ut, false, tid);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem, ua, kRtCs, ia, 96 * 16, 1, id96, 0, leader);
            if (leader) umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        rt_d96_to_stage<true>(tmem, 0, stage, cst, warp, lane);
        tc_fence_before();
        __syncthreads();
        rt_add(stage, rowx, a.x, a.y, warp, lane);
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------ unsqueeze backward
// TMEM: a = u Wun^T cols 0..95 | du = dv Wun cols 96..111 | d Wun^T (+ d bun in col 8) cols 112..127
constexpr uint32_t UNB_DY = 0;                       // dy -> dv tile, 16 chunks (12..15 zero: MN-major 128-feature window)
constexpr uint32_t UNB_U = UNB_DY + 16 * kRtCs;      // u tile (+ ones feature 8), 2 chunks
constexpr uint32_t UNB_IMG1 = UNB_U + 2 * kRtCs;     // Wun [96 x 16]
constexpr uint32_t UNB_IMG2 = UNB_IMG1 + RT_IMG;     // Wun^T [16 x 96]
constexpr uint32_t UNB_CST = UNB_IMG2 + RT_IMG;      // bun 96
constexpr uint32_t UNB_TAB = UNB_CST + 96 * 4;
constexpr uint32_t UNB_BAR = UNB_TAB + 2 * kRtRows * 8;
constexpr uint32_t UNB_SMEM = UNB_BAR + 32;

template <int FMT>
__global__ void __launch_bounds__(256, 4) unsqueeze_bwd_tc_kernel(RtArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const RtGeom g = a.g;
    unsigned char* dyt = smem + UNB_DY;
    unsigned char* ut = smem + UNB_U;
    unsigned char* img1 = smem + UNB_IMG1;
    unsigned char* img2 = smem + UNB_IMG2;
    float* cst = reinterpret_cast<float*>(smem + UNB_CST);
    long long* rowx = reinterpret_cast<long long*>(smem + UNB_TAB);
    long long* rowq = rowx + kRtRows;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + UNB_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 96; i += 256) cst[i] = a.bias[i];
    for (int i = tid; i < (int)(4 * kRtCs / 16); i += 256) reinterpret_cast<uint4*>(dyt + 12 * kRtCs)[i] = make_uint4(0, 0, 0, 0);
    {
        const float* W = a.W;  // Wun [96, 8]
        rt_build_img<FMT>(img1, 96, 16, [W](int n, int k) { return k < kSq ? W[n * kSq + k] : 0.f; }, tid, 256);
        rt_build_img<FMT>(img2, 16, 96, [W](int n, int k) { return n < kSq ? W[k * kSq + n] : 0.f; }, tid, 256);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t da = smem_u32(dyt), ua = smem_u32(ut), i1 = smem_u32(img1), i2 = smem_u32(img2);
    const uint32_t id96 = make_idesc(FMT, 128, 96, 0, 0), id16 = make_idesc(FMT, 128, 16, 0, 0), idw = make_idesc(FMT, 128, 16, 1, 1);
    uint32_t ph = 0;
    bool any = false;
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        rt_table(g, tile, rowx, rowq, tid);
        __syncthreads();
        rt_stage<FMT, false>(a.dy, rowx, dyt, nullptr, nullptr, nullptr, warp, lane);
        rt_gather8<FMT>(a.q_in, rowq, g.F, ut, true, tid);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem, ua, kRtCs, i1, 96 * 16, 1, id96, 0, leader);
            if (leader) umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        {
            // dv = dy * SiLU'(a), in place over the staged dy; thread = (row, column half)
            const int q4 = warp & 3, hf = warp >> 2, r = 32 * q4 + lane;
            const uint32_t tacc = tmem + ((uint32_t)(32 * q4) << 16);
            unsigned char* drow = dyt + r * 16;
#pragma unroll 1
            for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(tacc + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const uint4 pk = *reinterpret_cast<const uint4*>(drow + (size_t)(c0 / 8 + cc) * kRtCs);
                    float d[8];
                    unpack16<FMT>(pk.x, d[0], d[1]);
                    unpack16<FMT>(pk.y, d[2], d[3]);
                    unpack16<FMT>(pk.z, d[4], d[5]);
                    unpack16<FMT>(pk.w, d[6], d[7]);
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[e] *= silu_grad(__uint_as_float(v[8 * cc + e]) + cst[c0 + 8 * cc + e]);
                    *reinterpret_cast<uint4*>(drow + (size_t)(c0 / 8 + cc) * kRtCs) = pack8<FMT>(d);
                }
            }
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem + 96, da, kRtCs, i2, 256, 6, id16, 0, leader);
            for (int ks = 0; ks < kRtRows / 16; ++ks)
                if (leader) umma_f16(tmem + 112, sdesc_mnmajor(da + ks * 256, kRtCs), sdesc_mnmajor(ua + ks * 256, kRtCs), idw, (any || ks) ? 1u : 0u);
            if (leader) umma_commit(bar);
        }
        any = true;
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        if (warp < 4) {
            uint32_t v[8];
            tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + 96, v);
            tmem_ld_wait();
            const long long q = rowq[32 * warp + lane];
            if (q >= 0) {
#pragma unroll
                for (int j = 0; j < kSq; ++j) a.q_out[q + (long long)j * g.F] = __uint_as_float(v[j]);
            }
        }
        tc_fence_before();
        __syncthreads();
    }
    if (any && warp < 4) {
        // D3: lane = channel c, columns 0..7 = d Wun[c][g], column 8 = d bun[c]
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(32 * warp) << 16) + 112, v);
        tmem_ld_wait();
        const int c = 32 * warp + lane;
        if (c < kH) {
#pragma unroll
            for (int j = 0; j < kSq; ++j) atomicAdd(a.dW + c * kSq + j, __uint_as_float(v[j]));
            atomicAdd(a.dbias + c, __uint_as_float(v[8]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------ squeeze backward
// TMEM: z = LN(x) Wsq^T cols 0..15 | d ln = dz Wsq cols 16..111 | d Wsq^T (lane 96: d bsq) cols 112..127
constexpr uint32_t SQB_X = 0;                         // LN(x) tile, 16 chunks (feature 96 = ones, 97.. zero); later the
                                                      // fp32 d ln staging (24 chunks) aliases X and DZ's neighbour
constexpr uint32_t SQB_ST = 0;
constexpr uint32_t SQB_DZ = 24 * kRtCs;               // dz tile, 2 chunks
constexpr uint32_t SQB_IMG1 = SQB_DZ + 2 * kRtCs;     // Wsq [16 x 96]
constexpr uint32_t SQB_IMG2 = SQB_IMG1 + RT_IMG;      // Wsq^T [96 x 16]
constexpr uint32_t SQB_CST = SQB_IMG2 + RT_IMG;       // lnw 96, lnb 96, bsq 8, acc 192 (d lnw, d lnb)
constexpr uint32_t SQB_STATS = SQB_CST + 400 * 4;     // (mean, rstd) per row
constexpr uint32_t SQB_TAB = SQB_STATS + kRtRows * 8;
constexpr uint32_t SQB_BAR = SQB_TAB + 2 * kRtRows * 8;
constexpr uint32_t SQB_SMEM = SQB_BAR + 32;

template <int FMT>
__global__ void __launch_bounds__(256, 2) squeeze_bwd_tc_kernel(RtArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const RtGeom g = a.g;
    unsigned char* xt = smem + SQB_X;
    unsigned char* stage = smem + SQB_ST;
    unsigned char* dzt = smem + SQB_DZ;
    unsigned char* img1 = smem + SQB_IMG1;
    unsigned char* img2 = smem + SQB_IMG2;
    float* cst = reinterpret_cast<float*>(smem + SQB_CST);
    float* acc = cst + 208;
    float2* stats = reinterpret_cast<float2*>(smem + SQB_STATS);
    long long* rowx = reinterpret_cast<long long*>(smem + SQB_TAB);
    long long* rowq = rowx + kRtRows;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SQB_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    if (warp == 0) tmem_alloc(tmem_slot, 128);
    if (tid == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    for (int i = tid; i < 96; i += 256) { cst[i] = a.lnw[i]; cst[96 + i] = a.lnb[i]; }
    if (tid < kSq) cst[192 + tid] = a.bias[tid];
    for (int i = tid; i < 192; i += 256) acc[i] = 0.f;
    for (int i = tid; i < (int)(kRtCs / 16); i += 256) reinterpret_cast<uint4*>(dzt + kRtCs)[i] = make_uint4(0, 0, 0, 0);
    {
        const float* W = a.W;  // Wsq [8, 96]
        rt_build_img<FMT>(img1, 16, 96, [W](int n, int k) { return n < kSq ? W[n * kH + k] : 0.f; }, tid, 256);
        rt_build_img<FMT>(img2, 96, 16, [W](int n, int k) { return k < kSq ? W[k * kH + n] : 0.f; }, tid, 256);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t xa = smem_u32(xt), za = smem_u32(dzt), i1 = smem_u32(img1), i2 = smem_u32(img2);
    const uint32_t id96 = make_idesc(FMT, 128, 96, 0, 0), id16 = make_idesc(FMT, 128, 16, 0, 0), idw = make_idesc(FMT, 128, 16, 1, 1);
    uint32_t ph = 0;
    bool any = false;
    Oct12 dlg, dlb;
    dlg.zero();
    dlb.zero();
    for (int tile = blockIdx.x; tile < g.ntiles; tile += gridDim.x) {
        rt_table(g, tile, rowx, rowq, tid);
        __syncthreads();
        rt_stage<FMT, true>(a.x, rowx, xt, cst, cst + 96, stats, warp, lane);
        // chunk 12: ones feature (column sums of dz = d bsq) for rows that exist; chunks 13..15 zero (the previous tile's
        // fp32 staging overwrote them)
        if (tid < kRtRows) {
            float o[8] = {rowx[tid] >= 0 ? 1.f : 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            *reinterpret_cast<uint4*>(xt + 12 * kRtCs + tid * 16) = pack8<FMT>(o);
        }
        for (int i = tid; i < 3 * kRtRows; i += 256)
            *reinterpret_cast<uint4*>(xt + (size_t)(13 + i / kRtRows) * kRtCs + (i % kRtRows) * 16) = make_uint4(0, 0, 0, 0);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem, xa, kRtCs, i1, 256, 6, id16, 0, leader);
            if (leader) umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        if (warp < 4) {
            // dz = ds * SiLU'(z) -> 16-bit tile
            uint32_t v[8];
            tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16), v);
            tmem_ld_wait();
            const int r = 32 * warp + lane;
            const long long q = rowq[r];
            float d[8];
#pragma unroll
            for (int j = 0; j < kSq; ++j)
                d[j] = q >= 0 ? __ldg(a.q_in + q + (long long)j * g.F) * silu_grad(__uint_as_float(v[j]) + cst[192 + j]) : 0.f;
            *reinterpret_cast<uint4*>(dzt + r * 16) = pack8<FMT>(d);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            mma_kk(tmem + 16, za, kRtCs, i2, 96 * 16, 1, id96, 0, leader);
            for (int ks = 0; ks < kRtRows / 16; ++ks)
                if (leader) umma_f16(tmem + 112, sdesc_mnmajor(xa + ks * 256, kRtCs), sdesc_mnmajor(za + ks * 256, kRtCs), idw, (any || ks) ? 1u : 0u);
            if (leader) umma_commit(bar);
        }
        any = true;
        __syncwarp();
        mbar_wait(bar, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        rt_d96_to_stage<false>(tmem, 16, stage, nullptr, warp, lane);  // over the dead LN(x) tile
        tc_fence_before();
        __syncthreads();
        rt_ln_bwd(stage, rowx, a.x, a.dy, a.y, stats, cst, dlg, dlb, warp, lane);
        __syncthreads();
    }
    dlg.flush_atomic(acc, lane);
    dlb.flush_atomic(acc + 96, lane);
    if (any && warp < 4) {
        // D3: lane = channel c (lane 96 = the ones feature), column g: d Wsq[g][c] resp. d bsq[g]
        uint32_t v[8];
        tmem_ld8(tmem + ((uint32_t)(32 * warp) << 16) + 112, v);
        tmem_ld_wait();
        const int c = 32 * warp + lane;
        if (c < kH) {
#pragma unroll
            for (int j = 0; j < kSq; ++j) atomicAdd(a.dW + j * kH + c, __uint_as_float(v[j]));
        } else if (c == kH) {
#pragma unroll
            for (int j = 0; j < kSq; ++j) atomicAdd(a.dbias + j, __uint_as_float(v[j]));
        }
    }
    tc_fence_before();
    __syncthreads();
    for (int i = tid; i < 96; i += 256) { atomicAdd(a.dlnw + i, acc[i]); atomicAdd(a.dlnb + i, acc[96 + i]); }
    if (warp == 0) tmem_dealloc(tmem, 128);
}

static int rt_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

template <typename K>
static int rt_launch(K kern, RtArgs& a, uint32_t smem, int per_sm, void* stream) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int cap = per_sm * rt_sms();
    const int grid = a.g.ntiles < cap ? a.g.ntiles : cap;
    kern<<<grid, 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

}  // namespace nbss

using namespace nbss;

// s[b,t,g,f] = SiLU(Wsq LN(x) + bsq);  x [B,F,T,96], s [B,T,8,F]
extern "C" int nbss_squeeze_fwd_tc(const float* x, float* s, int B, int F, int T, const float* lnw, const float* lnb,
                                   const float* Wsq, const float* bsq, int fmt, int* err, void* stream) {
    if (!x || !s || !lnw || !lnb || !Wsq || !bsq) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    RtArgs a{rt_geom(B, F, T), x, nullptr, nullptr, nullptr, s, lnw, lnb, Wsq, bsq, nullptr, nullptr, nullptr, nullptr, err};
    return fmt == FMT_F16 ? rt_launch(squeeze_fwd_tc_kernel<FMT_F16>, a, SQF_SMEM, 3, stream)
                          : rt_launch(squeeze_fwd_tc_kernel<FMT_BF16>, a, SQF_SMEM, 3, stream);
}

// y = x + SiLU(Wun u + bun);  u [B,T,8,F]
extern "C" int nbss_unsqueeze_fwd_tc(const float* x, const float* u, float* y, int B, int F, int T, const float* Wun,
                                     const float* bun, int fmt, int* err, void* stream) {
    if (!x || !u || !y || !Wun || !bun) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    RtArgs a{rt_geom(B, F, T), x, nullptr, y, u, nullptr, nullptr, nullptr, Wun, bun, nullptr, nullptr, nullptr, nullptr, err};
    return fmt == FMT_F16 ? rt_launch(unsqueeze_fwd_tc_kernel<FMT_F16>, a, UNF_SMEM, 3, stream)
                          : rt_launch(unsqueeze_fwd_tc_kernel<FMT_BF16>, a, UNF_SMEM, 3, stream);
}

// Backward of the unsqueeze branch only: du [B,T,8,F], dWun += , dbun += .  (The residual dy is added by squeeze_bwd.)
extern "C" int nbss_unsqueeze_bwd_tc(const float* dy, const float* u, float* du, int B, int F, int T, const float* Wun,
                                     const float* bun, float* dWun, float* dbun, int fmt, int* err, void* stream) {
    if (!dy || !u || !du || !Wun || !bun || !dWun || !dbun) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    RtArgs a{rt_geom(B, F, T), nullptr, dy, nullptr, u, du, nullptr, nullptr, Wun, bun, nullptr, nullptr, dWun, dbun, err};
    return fmt == FMT_F16 ? rt_launch(unsqueeze_bwd_tc_kernel<FMT_F16>, a, UNB_SMEM, 4, stream)
                          : rt_launch(unsqueeze_bwd_tc_kernel<FMT_BF16>, a, UNB_SMEM, 4, stream);
}

// dx = dy + LN'(Wsq^T (ds * SiLU'(z)));  dWsq, dbsq, dlnw, dlnb += .
extern "C" int nbss_squeeze_bwd_tc(const float* x, const float* dy, const float* ds, float* dx, int B, int F, int T,
                                   const float* lnw, const float* lnb, const float* Wsq, const float* bsq, float* dWsq,
                                   float* dbsq, float* dlnw, float* dlnb, int fmt, int* err, void* stream) {
    if (!x || !dy || !ds || !dx || !lnw || !lnb || !Wsq || !bsq || !dWsq || !dbsq || !dlnw || !dlnb) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    RtArgs a{rt_geom(B, F, T), x, dy, dx, ds, nullptr, lnw, lnb, Wsq, bsq, dlnw, dlnb, dWsq, dbsq, err};
    return fmt == FMT_F16 ? rt_launch(squeeze_bwd_tc_kernel<FMT_F16>, a, SQB_SMEM, 2, stream)
                          : rt_launch(squeeze_bwd_tc_kernel<FMT_BF16>, a, SQB_SMEM, 2, stream);
}
