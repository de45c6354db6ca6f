// crossband.cu — cross-band block of SpatialNetLayer (models/arch/SpatialNet.py:85-87,116-146), fp32 CUDA-core
// kernels operating directly on the canonical stream layout [B,F,T,H=96] (no permute copies):
//   fconv_{fwd,bwd}    : y = x + PReLU(gconv_F(LN(x)))      one CTA per (b, 2 frames), whole F axis in smem
//   squeeze_{fwd,bwd}  : s[b,t,g,f] = SiLU(Wsq LN(x) + b)   row kernels, warp per T-F point
//   fullgemm           : u[b,t,g,:] = Wf[g] s[b,t,g,:] + bf  (LinearGroup, linear_group.py:29-34) + dgrad + wgrad
//   unsqueeze_{fwd,bwd}: y = x + SiLU(Wun u + b)
// These are 8 % of the layer FLOPs; they are HBM/L2-streaming kernels, fp32 end to end.
#define NBSS_SILU_EXACT  // two-MUFU sigmoid (common.cuh): the fp32 cross-check kernels are held to 1e-5 by the tests
#include <cstdlib>

#include "common.cuh"
#include "layout.cuh"

namespace nbss {

constexpr int kFK = 5;            // F-conv kernel size
constexpr int kFG = 12;           // channels per F-conv group (96/8)
constexpr int kHS = 8;            // dim_squeeze

__device__ __forceinline__ void ln_row(const float4 v, bool act, float& mean, float& rstd, float4& xh) {
    float s = warp_sum(act ? (v.x + v.y + v.z + v.w) : 0.f);
    mean = s * (1.f / kH);
    float4 d = act ? make_float4(v.x - mean, v.y - mean, v.z - mean, v.w - mean) : make_float4(0, 0, 0, 0);
    float q = warp_sum(d.x * d.x + d.y * d.y + d.z * d.z + d.w * d.w);
    rstd = rsqrtf(q * (1.f / kH) + 1e-5f);
    xh = make_float4(d.x * rstd, d.y * rstd, d.z * rstd, d.w * rstd);
}
__device__ __forceinline__ float4 f4_fma(float4 a, float4 b, float4 c) {
    return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_mul(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float f4_sum(float4 a) { return a.x + a.y + a.z + a.w; }

// Sum 8 per-lane values across the warp with 9 shuffles (recursive halving on lane bits 16, 8, 4, then 2 plain steps).
// On return every lane holds the total of value index g = lane >> 2.
__device__ __forceinline__ float warp_reduce8(float (&v)[8], int lane) {
    {
        const bool up = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float a = v[i], b = v[i + 4];
            v[i] = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 16);
        }
    }
    {
        const bool up = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float a = v[i], b = v[i + 2];
            v[i] = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 8);
        }
    }
    {
        const bool up = (lane & 4) != 0;
        const float a = v[0], b = v[1];
        v[0] = (up ? b : a) + __shfl_xor_sync(0xffffffffu, up ? a : b, 4);
    }
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 2);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
    return v[0];
}

// conv window dot product: 5 taps x 12 input channels of group g at frame slot tt
__device__ __forceinline__ float fconv_dot(const float* h, int f, int tt, int TT, int g, const float (&w)[60], float acc) {
#pragma unroll
    for (int k = 0; k < kFK; ++k) {
        const float4* row = reinterpret_cast<const float4*>(h + ((size_t)(f + k) * TT + tt) * kH + kFG * g);
        const float4 a0 = row[0], a1 = row[1], a2 = row[2];
        acc = fmaf(w[0 * 5 + k], a0.x, acc); acc = fmaf(w[1 * 5 + k], a0.y, acc);
        acc = fmaf(w[2 * 5 + k], a0.z, acc); acc = fmaf(w[3 * 5 + k], a0.w, acc);
        acc = fmaf(w[4 * 5 + k], a1.x, acc); acc = fmaf(w[5 * 5 + k], a1.y, acc);
        acc = fmaf(w[6 * 5 + k], a1.z, acc); acc = fmaf(w[7 * 5 + k], a1.w, acc);
        acc = fmaf(w[8 * 5 + k], a2.x, acc); acc = fmaf(w[9 * 5 + k], a2.y, acc);
        acc = fmaf(w[10 * 5 + k], a2.z, acc); acc = fmaf(w[11 * 5 + k], a2.w, acc);
    }
    return acc;
}

// LN(x) for the frames [t0, t0+TT) of utterance b into smem h[(F+4)][TT][96] (2 zero rows each side of F)
template <int TT>
__device__ __forceinline__ void fconv_stage(const float* __restrict__ x, float* h, int b, int F, int T, int t0,
                                            const float* lnw, const float* lnb, int tid, int nwarps) {
    const int warp = tid >> 5, lane = tid & 31;
    const bool act = lane < 24;
    for (int i = tid; i < 2 * TT * kH; i += nwarps * 32) { h[i] = 0.f; h[(size_t)(F + 2) * TT * kH + i] = 0.f; }
    float4 g4 = make_float4(0, 0, 0, 0), b4 = g4;
    if (act) { g4 = ld_f4(lnw + 4 * lane); b4 = ld_f4(lnb + 4 * lane); }
    for (int i = warp; i < F * TT; i += nwarps) {
        const int f = i / TT, tt = i % TT, t = t0 + tt;
        float4 v = make_float4(0, 0, 0, 0);
        if (act && t < T) v = ld_f4(x + (((size_t)b * F + f) * T + t) * kH + 4 * lane);
        float mean, rstd;
        float4 xh;
        ln_row(v, act, mean, rstd, xh);
        if (act) st_f4(h + ((size_t)(f + 2) * TT + tt) * kH + 4 * lane, t < T ? f4_fma(xh, g4, b4) : make_float4(0, 0, 0, 0));
    }
}

template <int TT>
__global__ void __launch_bounds__(96 * TT) fconv_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int F,
                                                            int T, const float* lnw, const float* lnb, const float* W,
                                                            const float* bias, const float* slope) {
    extern __shared__ __align__(16) float h[];
    const int tiles = (T + TT - 1) / TT, tid = threadIdx.x;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * TT;
    fconv_stage<TT>(x, h, b, F, T, t0, lnw, lnb, tid, 3 * TT);
    __syncthreads();
    const int tt = tid / kH, co = tid % kH, g = co / kFG, t = t0 + tt;
    float w[60];
#pragma unroll
    for (int i = 0; i < 60; ++i) w[i] = W[co * 60 + i];
    const float bi = bias[co], sl = slope[co];
    if (t < T) {
        for (int f = 0; f < F; ++f) {
            const float acc = fconv_dot(h, f, tt, TT, g, w, bi);
            const size_t idx = (((size_t)b * F + f) * T + t) * kH + co;
            y[idx] = x[idx] + (acc >= 0.f ? acc : sl * acc);
        }
    }
}

// Backward of y = x + PReLU(conv(LN(x))). Persistent CTAs; parameter gradients accumulate in registers and are
// flushed once with atomics into the fp32 gradient buffers (dW [96,12,5], dbias, dslope, dlnw, dlnb).
template <int TT>
__global__ void __launch_bounds__(96 * TT) fconv_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                            float* __restrict__ dx, int B, int F, int T, const float* lnw,
                                                            const float* lnb, const float* W, const float* bias,
                                                            const float* slope, float* dW, float* dbias, float* dslope,
                                                            float* dlnw, float* dlnb) {
    extern __shared__ __align__(16) float sm[];
    float* h = sm;                                  // LN(x), later dh
    float* dc = sm + (size_t)(F + 4) * TT * kH;     // grad wrt conv output (with PReLU derivative)
    const int tiles = (T + TT - 1) / TT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = 3 * TT;
    const int tt = tid / kH, co = tid % kH, g = co / kFG;
    const bool act = lane < 24;
    float w[60], wt[60], dw[60];
#pragma unroll
    for (int i = 0; i < 60; ++i) { w[i] = W[co * 60 + i]; dw[i] = 0.f; }
    // transposed weights for dgrad: thread owns input channel ci = co; wt[col*5+k] = W[12g+col][ci%12][k]
#pragma unroll
    for (int col = 0; col < kFG; ++col)
#pragma unroll
        for (int k = 0; k < kFK; ++k) wt[col * 5 + k] = W[(kFG * g + col) * 60 + (co % kFG) * 5 + k];
    const float bi = bias[co], sl = slope[co];
    float dbi = 0.f, dsl = 0.f;
    float4 g4 = make_float4(0, 0, 0, 0), dg4 = g4, db4 = g4;
    if (act) g4 = ld_f4(lnw + 4 * lane);

    for (int tile = blockIdx.x; tile < B * tiles; tile += gridDim.x) {
        const int b = tile / tiles, t0 = (tile % tiles) * TT, t = t0 + tt;
        __syncthreads();
        fconv_stage<TT>(x, h, b, F, T, t0, lnw, lnb, tid, nwarps);
        for (int i = tid; i < 2 * TT * kH; i += nwarps * 32) { dc[i] = 0.f; dc[(size_t)(F + 2) * TT * kH + i] = 0.f; }
        __syncthreads();
        // recompute conv, form dc, accumulate dW / dbias / dslope
        for (int f = 0; f < F; ++f) {
            float dcv = 0.f;
            if (t < T) {
                const float acc = fconv_dot(h, f, tt, TT, g, w, bi);
                const float dyv = dy[(((size_t)b * F + f) * T + t) * kH + co];
                dcv = dyv * (acc >= 0.f ? 1.f : sl);
                dsl += acc < 0.f ? dyv * acc : 0.f;
                dbi += dcv;
#pragma unroll
                for (int k = 0; k < kFK; ++k) {
                    const float4* row = reinterpret_cast<const float4*>(h + ((size_t)(f + k) * TT + tt) * kH + kFG * g);
                    const float4 a0 = row[0], a1 = row[1], a2 = row[2];
                    dw[0 * 5 + k] = fmaf(dcv, a0.x, dw[0 * 5 + k]); dw[1 * 5 + k] = fmaf(dcv, a0.y, dw[1 * 5 + k]);
                    dw[2 * 5 + k] = fmaf(dcv, a0.z, dw[2 * 5 + k]); dw[3 * 5 + k] = fmaf(dcv, a0.w, dw[3 * 5 + k]);
                    dw[4 * 5 + k] = fmaf(dcv, a1.x, dw[4 * 5 + k]); dw[5 * 5 + k] = fmaf(dcv, a1.y, dw[5 * 5 + k]);
                    dw[6 * 5 + k] = fmaf(dcv, a1.z, dw[6 * 5 + k]); dw[7 * 5 + k] = fmaf(dcv, a1.w, dw[7 * 5 + k]);
                    dw[8 * 5 + k] = fmaf(dcv, a2.x, dw[8 * 5 + k]); dw[9 * 5 + k] = fmaf(dcv, a2.y, dw[9 * 5 + k]);
                    dw[10 * 5 + k] = fmaf(dcv, a2.z, dw[10 * 5 + k]); dw[11 * 5 + k] = fmaf(dcv, a2.w, dw[11 * 5 + k]);
                }
            }
            dc[((size_t)(f + 2) * TT + tt) * kH + co] = dcv;
        }
        __syncthreads();
        // dgrad of the conv: dh[f, ci] = sum_k sum_col W[col, ci, k] * dc[f - k + 2, col]  -> overwrite h
        for (int f = 0; f < F; ++f) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < kFK; ++k) {
                const float4* row = reinterpret_cast<const float4*>(dc + ((size_t)(f + 4 - k) * TT + tt) * kH + kFG * g);
                const float4 a0 = row[0], a1 = row[1], a2 = row[2];
                acc = fmaf(wt[0 * 5 + k], a0.x, acc); acc = fmaf(wt[1 * 5 + k], a0.y, acc);
                acc = fmaf(wt[2 * 5 + k], a0.z, acc); acc = fmaf(wt[3 * 5 + k], a0.w, acc);
                acc = fmaf(wt[4 * 5 + k], a1.x, acc); acc = fmaf(wt[5 * 5 + k], a1.y, acc);
                acc = fmaf(wt[6 * 5 + k], a1.z, acc); acc = fmaf(wt[7 * 5 + k], a1.w, acc);
                acc = fmaf(wt[8 * 5 + k], a2.x, acc); acc = fmaf(wt[9 * 5 + k], a2.y, acc);
                acc = fmaf(wt[10 * 5 + k], a2.z, acc); acc = fmaf(wt[11 * 5 + k], a2.w, acc);
            }
            h[((size_t)(f + 2) * TT + tt) * kH + co] = acc;
        }
        __syncthreads();
        // LayerNorm backward, warp per row; dx = dy + dLN
        for (int i = warp; i < F * TT; i += nwarps) {
            const int f = i / TT, tr = i % TT, tq = t0 + tr;
            if (tq >= T) continue;
            const size_t base = (((size_t)b * F + f) * T + tq) * kH + 4 * lane;
            float4 v = make_float4(0, 0, 0, 0), dyv = v, dh = v;
            if (act) { v = ld_f4(x + base); dyv = ld_f4(dy + base); dh = ld_f4(h + ((size_t)(f + 2) * TT + tr) * kH + 4 * lane); }
            float mean, rstd;
            float4 xh;
            ln_row(v, act, mean, rstd, xh);
            const float4 dxh = f4_mul(dh, g4);
            const float m1 = warp_sum(f4_sum(dxh)) * (1.f / kH);
            const float m2 = warp_sum(f4_sum(f4_mul(dxh, xh))) * (1.f / kH);
            dg4 = f4_fma(dh, xh, dg4);
            db4 = f4_add(db4, dh);
            if (act) {
                float4 o;
                o.x = dyv.x + rstd * (dxh.x - m1 - xh.x * m2);
                o.y = dyv.y + rstd * (dxh.y - m1 - xh.y * m2);
                o.z = dyv.z + rstd * (dxh.z - m1 - xh.z * m2);
                o.w = dyv.w + rstd * (dxh.w - m1 - xh.w * m2);
                st_f4(dx + base, o);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 60; ++i) atomicAdd(dW + co * 60 + i, dw[i]);
    atomicAdd(dbias + co, dbi);
    atomicAdd(dslope + co, dsl);
    if (act) {
        atomicAdd(dlnw + 4 * lane + 0, dg4.x); atomicAdd(dlnw + 4 * lane + 1, dg4.y);
        atomicAdd(dlnw + 4 * lane + 2, dg4.z); atomicAdd(dlnw + 4 * lane + 3, dg4.w);
        atomicAdd(dlnb + 4 * lane + 0, db4.x); atomicAdd(dlnb + 4 * lane + 1, db4.y);
        atomicAdd(dlnb + 4 * lane + 2, db4.z); atomicAdd(dlnb + 4 * lane + 3, db4.w);
    }
}

// ------------------------------------------------------------------------------------------------ squeeze
// s[b,t,g,f] = SiLU(sum_c Wsq[g,c] LN(x)[b,f,t,c] + bsq[g]).  CTA per (b, 4 frames); results transposed through smem
// so the [B,T,8,F] tensor is written with f contiguous.
constexpr int kSQT = 4;

// Copies between the [kSQT][8][F] shared tile and a [B,T,8,F] tensor for frames t0..t0+3 and the range [f0, f0+FR):
// one warp per (tt, g) line, lanes along f (coalesced, no integer divisions).
__device__ __forceinline__ void sq_tile_load(float* tile, const float* __restrict__ src, int b, int t0, int T, int F, int f0,
                                             int FR, int warp, int lane) {
    for (int row = warp; row < kSQT * kHS; row += 8) {
        const int tt = row >> 3;
        const bool ok = t0 + tt < T;
        const float* sp = src + ((size_t)b * T + t0 + tt) * kHS * F + (size_t)(row & 7) * F + f0;
        float* tp = tile + row * F + f0;
        for (int fo = lane; fo < FR; fo += 32) tp[fo] = ok ? sp[fo] : 0.f;
    }
}
__device__ __forceinline__ void sq_tile_store(const float* tile, float* __restrict__ dst, int b, int t0, int T, int F, int f0,
                                              int FR, int warp, int lane) {
    for (int row = warp; row < kSQT * kHS; row += 8) {
        const int tt = row >> 3;
        if (t0 + tt >= T) continue;
        float* dp = dst + ((size_t)b * T + t0 + tt) * kHS * F + (size_t)(row & 7) * F + f0;
        const float* tp = tile + row * F + f0;
        for (int fo = lane; fo < FR; fo += 32) dp[fo] = tp[fo];
    }
}
__global__ void __launch_bounds__(256) squeeze_fwd_kernel(const float* __restrict__ x, float* __restrict__ s, int B, int F,
                                                          int T, const float* lnw, const float* lnb, const float* Wsq,
                                                          const float* bsq) {
    extern __shared__ __align__(16) float sm[];
    float* wsq = sm;                  // [8][96]
    float* st = sm + kHS * kH;        // [4][8][F]
    const int tiles = (T + kSQT - 1) / kSQT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kSQT;
    const int f0 = (int)blockIdx.y * F / (int)gridDim.y, FR = ((int)blockIdx.y + 1) * F / (int)gridDim.y - f0;
    const bool act = lane < 24;
    for (int i = tid; i < kHS * kH; i += 256) wsq[i] = Wsq[i];
    __syncthreads();
    float4 g4 = make_float4(0, 0, 0, 0), b4 = g4;
    if (act) { g4 = ld_f4(lnw + 4 * lane); b4 = ld_f4(lnb + 4 * lane); }
    for (int i = warp; i < FR * kSQT; i += 8) {
        const int f = f0 + i / kSQT, tt = i % kSQT, t = t0 + tt;
        if (t >= T) continue;
        float4 v = make_float4(0, 0, 0, 0);
        if (act) v = ld_f4(x + (((size_t)b * F + f) * T + t) * kH + 4 * lane);
        float mean, rstd;
        float4 xh;
        ln_row(v, act, mean, rstd, xh);
        const float4 ln = f4_fma(xh, g4, b4);
        float p[kHS];
#pragma unroll
        for (int g = 0; g < kHS; ++g) p[g] = act ? f4_sum(f4_mul(ln, ld_f4(wsq + g * kH + 4 * lane))) : 0.f;
        const float tot = warp_reduce8(p, lane);  // lane holds group lane >> 2
        if ((lane & 3) == 0) st[(tt * kHS + (lane >> 2)) * F + f] = silu(tot + bsq[lane >> 2]);
    }
    __syncthreads();
    sq_tile_store(st, s, b, t0, T, F, f0, FR, warp, lane);
}

// Backward of squeeze + its LayerNorm; also adds the residual: dx = dy + dLN.  ds: [B,T,8,F] gradient wrt s.
__global__ void __launch_bounds__(256) squeeze_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          const float* __restrict__ ds, float* __restrict__ dx, int B,
                                                          int F, int T, const float* lnw, const float* lnb,
                                                          const float* Wsq, const float* bsq, float* dWsq, float* dbsq,
                                                          float* dlnw, float* dlnb, int fsplit) {
    extern __shared__ __align__(16) float sm[];
    float* wsq = sm;                  // [8][96]
    float* st = sm + kHS * kH;        // [4][8][F] ds tile
    const int tiles = (T + kSQT - 1) / kSQT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool act = lane < 24;
    for (int i = tid; i < kHS * kH; i += 256) wsq[i] = Wsq[i];
    float4 g4 = make_float4(0, 0, 0, 0), b4 = g4, dg4 = g4, db4 = g4;
    if (act) { g4 = ld_f4(lnw + 4 * lane); b4 = ld_f4(lnb + 4 * lane); }
    float4 dwq[kHS];
#pragma unroll
    for (int g = 0; g < kHS; ++g) dwq[g] = make_float4(0, 0, 0, 0);
    float dbq = 0.f;  // lane g accumulates dbsq[g]
    for (int job = blockIdx.x; job < B * tiles * fsplit; job += gridDim.x) {
        const int fs = job % fsplit, tile = job / fsplit;
        const int f0 = fs * F / fsplit, FR = (fs + 1) * F / fsplit - f0;
        const int b = tile / tiles, t0 = (tile % tiles) * kSQT;
        __syncthreads();
        sq_tile_load(st, ds, b, t0, T, F, f0, FR, warp, lane);
        __syncthreads();
        for (int i = warp; i < FR * kSQT; i += 8) {
            const int f = f0 + i / kSQT, tt = i % kSQT, t = t0 + tt;
            if (t >= T) continue;
            const size_t base = (((size_t)b * F + f) * T + t) * kH + 4 * lane;
            float4 v = make_float4(0, 0, 0, 0), dyv = v;
            if (act) { v = ld_f4(x + base); dyv = ld_f4(dy + base); }
            float mean, rstd;
            float4 xh;
            ln_row(v, act, mean, rstd, xh);
            const float4 ln = f4_fma(xh, g4, b4);
            float4 dln = make_float4(0, 0, 0, 0);
            float pz[kHS];
#pragma unroll
            for (int g = 0; g < kHS; ++g) pz[g] = act ? f4_sum(f4_mul(ln, ld_f4(wsq + g * kH + 4 * lane))) : 0.f;
            const int gown = lane >> 2;
            const float zown = warp_reduce8(pz, lane) + bsq[gown];
            const float dz_own = st[(tt * kHS + gown) * F + f] * silu_grad(zown);
#pragma unroll
            for (int g = 0; g < kHS; ++g) {
                const float4 wv = act ? ld_f4(wsq + g * kH + 4 * lane) : make_float4(0, 0, 0, 0);
                const float dz = __shfl_sync(0xffffffffu, dz_own, 4 * g);
                dln = make_float4(fmaf(dz, wv.x, dln.x), fmaf(dz, wv.y, dln.y), fmaf(dz, wv.z, dln.z), fmaf(dz, wv.w, dln.w));
                dwq[g] = make_float4(fmaf(dz, ln.x, dwq[g].x), fmaf(dz, ln.y, dwq[g].y), fmaf(dz, ln.z, dwq[g].z), fmaf(dz, ln.w, dwq[g].w));
                if (lane == g) dbq += dz;
            }
            const float4 dxh = f4_mul(dln, g4);
            const float m1 = warp_sum(f4_sum(dxh)) * (1.f / kH);
            const float m2 = warp_sum(f4_sum(f4_mul(dxh, xh))) * (1.f / kH);
            dg4 = f4_fma(dln, xh, dg4);
            db4 = f4_add(db4, dln);
            if (act) {
                float4 o;
                o.x = dyv.x + rstd * (dxh.x - m1 - xh.x * m2);
                o.y = dyv.y + rstd * (dxh.y - m1 - xh.y * m2);
                o.z = dyv.z + rstd * (dxh.z - m1 - xh.z * m2);
                o.w = dyv.w + rstd * (dxh.w - m1 - xh.w * m2);
                st_f4(dx + base, o);
            }
        }
    }
    // flush: warps -> shared accumulators -> one set of global atomics per CTA
    float* red = st + kSQT * kHS * F;  // [768 dWsq | 96 dlnw | 96 dlnb | 8 dbsq]
    __syncthreads();
    for (int i = tid; i < 968; i += 256) red[i] = 0.f;
    __syncthreads();
    if (act) {
#pragma unroll
        for (int g = 0; g < kHS; ++g) {
            atomicAdd(red + g * kH + 4 * lane + 0, dwq[g].x); atomicAdd(red + g * kH + 4 * lane + 1, dwq[g].y);
            atomicAdd(red + g * kH + 4 * lane + 2, dwq[g].z); atomicAdd(red + g * kH + 4 * lane + 3, dwq[g].w);
        }
        atomicAdd(red + 768 + 4 * lane + 0, dg4.x); atomicAdd(red + 768 + 4 * lane + 1, dg4.y);
        atomicAdd(red + 768 + 4 * lane + 2, dg4.z); atomicAdd(red + 768 + 4 * lane + 3, dg4.w);
        atomicAdd(red + 864 + 4 * lane + 0, db4.x); atomicAdd(red + 864 + 4 * lane + 1, db4.y);
        atomicAdd(red + 864 + 4 * lane + 2, db4.z); atomicAdd(red + 864 + 4 * lane + 3, db4.w);
    }
    if (lane < kHS) atomicAdd(red + 960 + lane, dbq);
    __syncthreads();
    for (int i = tid; i < 968; i += 256) {
        float* dst = i < 768 ? dWsq + i : (i < 864 ? dlnw + (i - 768) : (i < 960 ? dlnb + (i - 864) : dbsq + (i - 960)));
        atomicAdd(dst, red[i]);
    }
}

// ------------------------------------------------------------------------------------------------ unsqueeze
// y[b,f,t,:] = x + SiLU(Wun u[b,t,:,f] + bun)
__global__ void __launch_bounds__(256) unsqueeze_fwd_kernel(const float* __restrict__ x, const float* __restrict__ u,
                                                            float* __restrict__ y, int B, int F, int T, const float* Wun,
                                                            const float* bun) {
    extern __shared__ __align__(16) float sm[];
    float* ut = sm;  // [4][8][F]
    const int tiles = (T + kSQT - 1) / kSQT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kSQT;
    const int f0 = (int)blockIdx.y * F / (int)gridDim.y, FR = ((int)blockIdx.y + 1) * F / (int)gridDim.y - f0;
    const bool act = lane < 24;
    sq_tile_load(ut, u, b, t0, T, F, f0, FR, warp, lane);
    float wun[4][kHS];
    float4 bu = make_float4(0, 0, 0, 0);
    if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int g = 0; g < kHS; ++g) wun[c][g] = Wun[(4 * lane + c) * kHS + g];
        bu = ld_f4(bun + 4 * lane);
    }
    __syncthreads();
    if (!act) return;
    for (int i = warp; i < FR * kSQT; i += 8) {
        const int f = f0 + i / kSQT, tt = i % kSQT, t = t0 + tt;
        if (t >= T) continue;
        float a[4] = {bu.x, bu.y, bu.z, bu.w};
#pragma unroll
        for (int g = 0; g < kHS; ++g) {
            const float uv = ut[(tt * kHS + g) * F + f];
#pragma unroll
            for (int c = 0; c < 4; ++c) a[c] = fmaf(wun[c][g], uv, a[c]);
        }
        const size_t base = (((size_t)b * F + f) * T + t) * kH + 4 * lane;
        const float4 xv = ld_f4(x + base);
        st_f4(y + base, make_float4(xv.x + silu(a[0]), xv.y + silu(a[1]), xv.z + silu(a[2]), xv.w + silu(a[3])));
    }
}

// Backward of the unsqueeze branch only: du[b,t,g,f] and dWun/dbun.  (The residual dy is added by squeeze_bwd.)
__global__ void __launch_bounds__(256) unsqueeze_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ u,
                                                            float* __restrict__ du, int B, int F, int T, const float* Wun,
                                                            const float* bun, float* dWun, float* dbun, int fsplit) {
    extern __shared__ __align__(16) float sm[];
    float* ut = sm;                      // [4][8][F]
    float* dut = sm + kSQT * kHS * F;    // [4][8][F]
    const int tiles = (T + kSQT - 1) / kSQT, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool act = lane < 24;
    float wun[4][kHS], dwun[4][kHS];
    float4 bu = make_float4(0, 0, 0, 0), dbu = bu;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int g = 0; g < kHS; ++g) { wun[c][g] = act ? Wun[(4 * lane + c) * kHS + g] : 0.f; dwun[c][g] = 0.f; }
    if (act) bu = ld_f4(bun + 4 * lane);
    for (int job = blockIdx.x; job < B * tiles * fsplit; job += gridDim.x) {
        const int fs = job % fsplit, tile = job / fsplit;
        const int f0 = fs * F / fsplit, FR = (fs + 1) * F / fsplit - f0;
        const int b = tile / tiles, t0 = (tile % tiles) * kSQT;
        __syncthreads();
        sq_tile_load(ut, u, b, t0, T, F, f0, FR, warp, lane);
        __syncthreads();
        for (int i = warp; i < FR * kSQT; i += 8) {
            const int f = f0 + i / kSQT, tt = i % kSQT, t = t0 + tt;
            if (t >= T) continue;
            float a[4] = {bu.x, bu.y, bu.z, bu.w}, uv[kHS];
#pragma unroll
            for (int g = 0; g < kHS; ++g) {
                uv[g] = ut[(tt * kHS + g) * F + f];
#pragma unroll
                for (int c = 0; c < 4; ++c) a[c] = fmaf(wun[c][g], uv[g], a[c]);
            }
            float4 dyv = make_float4(0, 0, 0, 0);
            if (act) dyv = ld_f4(dy + (((size_t)b * F + f) * T + t) * kH + 4 * lane);
            const float dv[4] = {dyv.x * silu_grad(a[0]), dyv.y * silu_grad(a[1]), dyv.z * silu_grad(a[2]), dyv.w * silu_grad(a[3])};
            dbu = make_float4(dbu.x + dv[0], dbu.y + dv[1], dbu.z + dv[2], dbu.w + dv[3]);
            float pg[kHS];
#pragma unroll
            for (int g = 0; g < kHS; ++g) {
                float p = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) { p = fmaf(dv[c], wun[c][g], p); dwun[c][g] = fmaf(dv[c], uv[g], dwun[c][g]); }
                pg[g] = p;
            }
            const float tot = warp_reduce8(pg, lane);  // lane holds group lane >> 2
            if ((lane & 3) == 0) dut[(tt * kHS + (lane >> 2)) * F + f] = tot;
        }
        __syncthreads();
        sq_tile_store(dut, du, b, t0, T, F, f0, FR, warp, lane);
    }
    // flush: warps -> shared accumulators -> one set of global atomics per CTA
    float* red = sm + 2 * kSQT * kHS * F;  // [768 dWun | 96 dbun]
    __syncthreads();
    for (int i = tid; i < 864; i += 256) red[i] = 0.f;
    __syncthreads();
    if (act) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int g = 0; g < kHS; ++g) atomicAdd(red + (4 * lane + c) * kHS + g, dwun[c][g]);
        atomicAdd(red + 768 + 4 * lane + 0, dbu.x); atomicAdd(red + 768 + 4 * lane + 1, dbu.y);
        atomicAdd(red + 768 + 4 * lane + 2, dbu.z); atomicAdd(red + 768 + 4 * lane + 3, dbu.w);
    }
    __syncthreads();
    for (int i = tid; i < 864; i += 256) atomicAdd(i < 768 ? dWun + i : dbun + (i - 768), red[i]);
}

// ------------------------------------------------------------------------------------------------ LinearGroup
// rows r = (frame m, group g) of a [M*8, F] matrix.  out[m,g,k] = sum_f in[m,g,f] * Wf[g,k,f] (+ bf[g,k])   (trans=0)
//                                                    out[m,g,f] = sum_k in[m,g,k] * Wf[g,k,f]                (trans=1)
// CTA: 64 frames x one group x one block of OB outputs (blockIdx.z); the block's OB x F weight rows sit in smem (padded rows),
// 256 threads x (1 frame, OB/4 outs).  OB = F for F <= 190 (one block); wider bands (16 kHz: F = 257) take two blocks.
__global__ void __launch_bounds__(256) fullgemm_kernel(const float* __restrict__ in, float* __restrict__ out, int M, int F,
                                                       const float* __restrict__ Wf, const float* __restrict__ bf, int trans, int OB) {
    extern __shared__ __align__(16) float sm[];
    const int FP = F + 1;
    float* ws = sm;               // [OB][FP]   ws[o][i] = weight(out o0 + o, in i)
    float* as = sm + OB * FP;     // [64][FP]
    const int g = blockIdx.y, m0 = blockIdx.x * 64, tid = threadIdx.x;
    const int o0 = blockIdx.z * OB, on = min(OB, F - o0);
    const float* wg = Wf + (size_t)g * F * F;
    for (int i = tid; i < on * F; i += 256) {
        const int r = i / F, c = i % F;  // W[k][f]: out = k (forward) or out = f (data gradient)
        ws[r * FP + c] = !trans ? wg[(size_t)(o0 + r) * F + c] : wg[(size_t)c * F + o0 + r];
    }
    for (int i = tid; i < 64 * F; i += 256) {
        const int fr = i / F, c = i % F;
        as[fr * FP + c] = (m0 + fr < M) ? in[((size_t)(m0 + fr) * kHS + g) * F + c] : 0.f;
    }
    __syncthreads();
    const int fr = tid >> 2, oq = tid & 3;
    constexpr int MAXO = 48;  // supports OB <= 192
    float acc[MAXO];
#pragma unroll
    for (int j = 0; j < MAXO; ++j) acc[j] = 0.f;
    const float* arow = as + fr * FP;
    for (int i = 0; i < F; ++i) {
        const float a = arow[i];
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int o = oq + 4 * j;
            if (o < on) acc[j] = fmaf(a, ws[o * FP + i], acc[j]);
        }
    }
    if (m0 + fr < M) {
#pragma unroll
        for (int j = 0; j < MAXO; ++j) {
            const int o = oq + 4 * j;
            if (o < on) out[((size_t)(m0 + fr) * kHS + g) * F + o0 + o] = acc[j] + ((bf && !trans) ? bf[g * F + o0 + o] : 0.f);
        }
    }
}
static inline int fullgemm_ob(int F) {  // outputs per CTA: everything when the whole weight fits, else the fewest equal blocks <= 192
    if ((size_t)(F + 64) * (F + 1) * 4 <= 200 * 1024 && F <= 192) return F;
    const int nb = (F + 131) / 132;
    return (F + nb - 1) / nb;
}

// dWf[g,k,f] += sum_m du[m,g,k] * s[m,g,f];  dbf[g,k] += sum_m du[m,g,k].   grid (chunks, 8 groups, kblk*fblk)
// Each CTA owns a 144 x 144 block of (k, f) outputs (9 x 9 per thread); column f == F is the bias gradient.
__global__ void __launch_bounds__(256) fullwgrad_kernel(const float* __restrict__ du, const float* __restrict__ s, int M,
                                                        int F, float* dWf, float* dbf) {
    extern __shared__ __align__(16) float sm[];
    constexpr int NJ = 9, BW = 16 * NJ;
    float* a = sm;              // du tile [32][BW]
    float* bt = sm + 32 * BW;   // s tile  [32][BW] (virtual column F = 1 -> bias gradient)
    const int nfb = (F + 1 + BW - 1) / BW;
    const int g = blockIdx.y, tid = threadIdx.x, kk = tid >> 4, ff = tid & 15;
    const int kbase = (blockIdx.z / nfb) * BW, fbase = (blockIdx.z % nfb) * BW;
    const int per = (M + gridDim.x - 1) / gridDim.x, mlo = blockIdx.x * per, mhi = min(M, mlo + per);
    float acc[NJ][NJ];
#pragma unroll
    for (int i = 0; i < NJ; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) acc[i][j] = 0.f;
    for (int m0 = mlo; m0 < mhi; m0 += 32) {
        __syncthreads();
        for (int i = tid; i < 32 * BW; i += 256) {
            const int r = i / BW, c = i % BW;
            const bool ok = m0 + r < mhi;
            const int k = kbase + c, f = fbase + c;
            a[i] = (ok && k < F) ? du[((size_t)(m0 + r) * kHS + g) * F + k] : 0.f;
            bt[i] = ok ? (f < F ? s[((size_t)(m0 + r) * kHS + g) * F + f] : (f == F ? 1.f : 0.f)) : 0.f;
        }
        __syncthreads();
        for (int r = 0; r < 32; ++r) {
            float av[NJ], bv[NJ];
#pragma unroll
            for (int i = 0; i < NJ; ++i) av[i] = a[r * BW + kk + 16 * i];
#pragma unroll
            for (int j = 0; j < NJ; ++j) bv[j] = bt[r * BW + ff + 16 * j];
#pragma unroll
            for (int i = 0; i < NJ; ++i)
#pragma unroll
                for (int j = 0; j < NJ; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
    }
#pragma unroll
    for (int i = 0; i < NJ; ++i)
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const int k = kbase + kk + 16 * i, f = fbase + ff + 16 * j;
            if (k < F && f < F) atomicAdd(dWf + ((size_t)g * F + k) * F + f, acc[i][j]);
            else if (k < F && f == F) atomicAdd(dbf + g * F + k, acc[i][j]);
        }
}

static int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

// Small batches: split every (b, 4-frame) tile of the squeeze/unsqueeze row kernels into frequency ranges so that the
// grid still fills 4 CTAs per SM.
static int f_split(int tiles, int F) {
    const int want = 4 * num_sms();
    int fs = 1;
    while (fs < 4 && tiles * fs < want && F / (2 * fs) >= 8) fs *= 2;
    return fs;
}

}  // namespace nbss

using namespace nbss;

extern "C" int nbss_fconv_fwd(const float* x, float* y, int B, int F, int T, const float* lnw, const float* lnb,
                              const float* W, const float* bias, const float* slope, void* stream) {
    if (!x || !y || !lnw || !lnb || !W || !bias || !slope) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    const size_t smem2 = (size_t)(F + 4) * 2 * kH * 4, smem1 = (size_t)(F + 4) * kH * 4;
    if (smem2 <= 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fconv_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        fconv_fwd_kernel<2><<<B * ((T + 1) / 2), 192, smem2, (cudaStream_t)stream>>>(x, y, B, F, T, lnw, lnb, W, bias, slope);
    } else if (smem1 <= 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fconv_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return (int)e;
        fconv_fwd_kernel<1><<<B * T, 96, smem1, (cudaStream_t)stream>>>(x, y, B, F, T, lnw, lnb, W, bias, slope);
    } else return NBSS_ERR_UNSUPPORTED;
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_fconv_bwd(const float* x, const float* dy, float* dx, int B, int F, int T, const float* lnw,
                              const float* lnb, const float* W, const float* bias, const float* slope, float* dW,
                              float* dbias, float* dslope, float* dlnw, float* dlnb, void* stream) {
    if (!x || !dy || !dx || !lnw || !lnb || !W || !bias || !slope || !dW || !dbias || !dslope || !dlnw || !dlnb) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    const size_t smem2 = (size_t)(F + 4) * 2 * kH * 4 * 2, smem1 = (size_t)(F + 4) * kH * 4 * 2;
    const int sms = num_sms();
    if (smem2 <= 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fconv_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return (int)e;
        const int tiles = B * ((T + 1) / 2);
        fconv_bwd_kernel<2><<<tiles < sms ? tiles : sms, 192, smem2, (cudaStream_t)stream>>>(x, dy, dx, B, F, T, lnw, lnb, W, bias, slope, dW, dbias, dslope, dlnw, dlnb);
    } else if (smem1 <= 227 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fconv_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
        if (e != cudaSuccess) return (int)e;
        const int tiles = B * T;
        fconv_bwd_kernel<1><<<tiles < 2 * sms ? tiles : 2 * sms, 96, smem1, (cudaStream_t)stream>>>(x, dy, dx, B, F, T, lnw, lnb, W, bias, slope, dW, dbias, dslope, dlnw, dlnb);
    } else return NBSS_ERR_UNSUPPORTED;
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// y = x + full-band branch.  s_out / u_out: [B,T,8,F] workspaces (kept by the caller for the backward pass).
extern "C" int nbss_full_fwd(const float* x, float* y, float* s_out, float* u_out, int B, int F, int T, const float* lnw,
                             const float* lnb, const float* Wsq, const float* bsq, const float* Wf, const float* bf,
                             const float* Wun, const float* bun, void* stream) {
    if (!x || !y || !s_out || !u_out || !lnw || !lnb || !Wsq || !bsq || !Wf || !bf || !Wun || !bun) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int OB = fullgemm_ob(F);
    if ((size_t)(OB + 64) * (F + 1) * 4 > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = B * ((T + kSQT - 1) / kSQT);
    const size_t sm_sq = (size_t)(kHS * kH + kSQT * kHS * F) * 4;
    const int fsplit = f_split(tiles, F);
    squeeze_fwd_kernel<<<dim3(tiles, fsplit), 256, sm_sq, st>>>(x, s_out, B, F, T, lnw, lnb, Wsq, bsq);
    NBSS_LAUNCH_CHECK();
    const size_t sm_g = (size_t)(OB + 64) * (F + 1) * 4;
    cudaError_t e = cudaFuncSetAttribute(fullgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_g);
    if (e != cudaSuccess) return (int)e;
    const int M = B * T;
    fullgemm_kernel<<<dim3((M + 63) / 64, kHS, (F + OB - 1) / OB), 256, sm_g, st>>>(s_out, u_out, M, F, Wf, bf, 0, OB);
    NBSS_LAUNCH_CHECK();
    unsqueeze_fwd_kernel<<<dim3(tiles, fsplit), 256, (size_t)kSQT * kHS * F * 4, st>>>(x, u_out, y, B, F, T, Wun, bun);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// Backward of y = x + full(x).  s, u: saved by nbss_full_fwd.  ws: workspace of 2 * B*T*8*F floats.
extern "C" int nbss_full_bwd(const float* x, const float* dy, float* dx, const float* s, const float* u, float* ws, int B,
                             int F, int T, const float* lnw, const float* lnb, const float* Wsq, const float* bsq,
                             const float* Wf, const float* Wun, const float* bun, float* dlnw, float* dlnb, float* dWsq,
                             float* dbsq, float* dWf, float* dbf, float* dWun, float* dbun, void* stream) {
    if (!x || !dy || !dx || !s || !u || !ws) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int OB = fullgemm_ob(F);
    if ((size_t)(OB + 64) * (F + 1) * 4 > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = num_sms();
    const int tiles = B * ((T + kSQT - 1) / kSQT), M = B * T;
    float* du = ws;
    float* ds = ws + (size_t)M * kHS * F;
    const int fsplit = f_split(tiles, F);
    const int pg = tiles * fsplit < 4 * sms ? tiles * fsplit : 4 * sms;  // persistent row kernels: 4 CTAs (32 warps) per SM
    cudaFuncSetAttribute(unsqueeze_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)(2 * kSQT * kHS * F + 1024) * 4));
    unsqueeze_bwd_kernel<<<pg, 256, (size_t)(2 * kSQT * kHS * F + 1024) * 4, st>>>(dy, u, du, B, F, T, Wun, bun, dWun, dbun, fsplit);
    NBSS_LAUNCH_CHECK();
    const size_t sm_g = (size_t)(OB + 64) * (F + 1) * 4;
    cudaError_t e = cudaFuncSetAttribute(fullgemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_g);
    if (e != cudaSuccess) return (int)e;
    fullgemm_kernel<<<dim3((M + 63) / 64, kHS, (F + OB - 1) / OB), 256, sm_g, st>>>(du, ds, M, F, Wf, nullptr, 1, OB);
    NBSS_LAUNCH_CHECK();
    const int chunks = M < 32 * 32 ? (M + 31) / 32 : 32;
    const int nkb = (F + 143) / 144, nfb = (F + 1 + 143) / 144;
    fullwgrad_kernel<<<dim3(chunks, kHS, nkb * nfb), 256, (size_t)2 * 32 * 144 * 4, st>>>(du, s, M, F, dWf, dbf);
    NBSS_LAUNCH_CHECK();
    const size_t sm_sq = (size_t)(kHS * kH + kSQT * kHS * F + 1024) * 4;
    squeeze_bwd_kernel<<<pg, 256, sm_sq, st>>>(x, dy, ds, dx, B, F, T, lnw, lnb, Wsq, bsq, dWsq, dbsq, dlnw, dlnb, fsplit);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// ---- the same block with the LinearGroup on tensor cores (fullband_tc.cu); img from nbss_lg_pack ---------------------
extern "C" int nbss_lg_tc_apply(const float* in, float* out, int M, int F, const void* img, const float* bias, int mode,
                                int fmt, int* err, void* stream);
extern "C" int nbss_lg_tc_wgrad(const float* du, const float* s, int M, int F, float* dW, float* db, int fmt, int* err,
                                void* stream);

extern "C" int nbss_squeeze_fwd_tc(const float* x, float* s, int B, int F, int T, const float* lnw, const float* lnb,
                                   const float* Wsq, const float* bsq, int fmt, int* err, void* stream);
extern "C" int nbss_unsqueeze_fwd_tc(const float* x, const float* u, float* y, int B, int F, int T, const float* Wun,
                                     const float* bun, int fmt, int* err, void* stream);
extern "C" int nbss_unsqueeze_bwd_tc(const float* dy, const float* u, float* du, int B, int F, int T, const float* Wun,
                                     const float* bun, float* dWun, float* dbun, int fmt, int* err, void* stream);
extern "C" int nbss_squeeze_bwd_tc(const float* x, const float* dy, const float* ds, float* dx, int B, int F, int T,
                                   const float* lnw, const float* lnb, const float* Wsq, const float* bsq, float* dWsq,
                                   float* dbsq, float* dlnw, float* dlnb, int fmt, int* err, void* stream);

// NBSS_ROWS_SIMT=1: keep the squeeze / unsqueeze halves on the fp32 CUDA-core kernels above (cross-check of
// fullband_rows_tc.cu); the LinearGroup stays on tensor cores.
static bool rows_simt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("NBSS_ROWS_SIMT");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}

extern "C" int nbss_full_fwd_tc(const float* x, float* y, float* s_out, float* u_out, int B, int F, int T, const float* lnw,
                                const float* lnb, const float* Wsq, const float* bsq, const float* bf, const float* Wun,
                                const float* bun, const void* img, int fmt, int* err, void* stream) {
    if (!x || !y || !s_out || !u_out || !lnw || !lnb || !Wsq || !bsq || !bf || !Wun || !bun || !img) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int tiles = B * ((T + kSQT - 1) / kSQT);
    const size_t sm_sq = (size_t)(kHS * kH + kSQT * kHS * F) * 4;
    if (sm_sq > 48 * 1024) return NBSS_ERR_UNSUPPORTED;
    const int fsplit = f_split(tiles, F);
    int rc;
    if (rows_simt()) {
        squeeze_fwd_kernel<<<dim3(tiles, fsplit), 256, sm_sq, st>>>(x, s_out, B, F, T, lnw, lnb, Wsq, bsq);
        NBSS_LAUNCH_CHECK();
    } else if ((rc = nbss_squeeze_fwd_tc(x, s_out, B, F, T, lnw, lnb, Wsq, bsq, fmt, err, stream)) != NBSS_OK) return rc;
    rc = nbss_lg_tc_apply(s_out, u_out, B * T, F, img, bf, 0, fmt, err, stream);
    if (rc != NBSS_OK) return rc;
    if (rows_simt()) {
        unsqueeze_fwd_kernel<<<dim3(tiles, fsplit), 256, (size_t)kSQT * kHS * F * 4, st>>>(x, u_out, y, B, F, T, Wun, bun);
        NBSS_LAUNCH_CHECK();
    } else if ((rc = nbss_unsqueeze_fwd_tc(x, u_out, y, B, F, T, Wun, bun, fmt, err, stream)) != NBSS_OK) return rc;
    return NBSS_OK;
}

extern "C" int nbss_full_bwd_tc(const float* x, const float* dy, float* dx, const float* s, const float* u, float* ws, int B,
                                int F, int T, const float* lnw, const float* lnb, const float* Wsq, const float* bsq,
                                const float* Wun, const float* bun, const void* img, float* dlnw, float* dlnb, float* dWsq,
                                float* dbsq, float* dWf, float* dbf, float* dWun, float* dbun, int fmt, int* err,
                                void* stream) {
    if (!x || !dy || !dx || !s || !u || !ws || !img) return NBSS_ERR_NULL;
    if (B < 1 || F < 1 || T < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int sms = num_sms();
    const int tiles = B * ((T + kSQT - 1) / kSQT), M = B * T;
    float* du = ws;
    float* ds = ws + (size_t)M * kHS * F;
    const int fsplit = f_split(tiles, F);
    const int pg = tiles * fsplit < 4 * sms ? tiles * fsplit : 4 * sms;
    const size_t sm_un = (size_t)(2 * kSQT * kHS * F + 1024) * 4, sm_sq = (size_t)(kHS * kH + kSQT * kHS * F + 1024) * 4;
    if (sm_un > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    int rc;
    if (rows_simt()) {
        cudaFuncSetAttribute(unsqueeze_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_un);
        unsqueeze_bwd_kernel<<<pg, 256, sm_un, st>>>(dy, u, du, B, F, T, Wun, bun, dWun, dbun, fsplit);
        NBSS_LAUNCH_CHECK();
    } else if ((rc = nbss_unsqueeze_bwd_tc(dy, u, du, B, F, T, Wun, bun, dWun, dbun, fmt, err, stream)) != NBSS_OK) return rc;
    rc = nbss_lg_tc_apply(du, ds, M, F, img, nullptr, 1, fmt, err, stream);
    if (rc != NBSS_OK) return rc;
    rc = nbss_lg_tc_wgrad(du, s, M, F, dWf, dbf, fmt, err, stream);
    if (rc != NBSS_OK) return rc;
    if (rows_simt()) {
        squeeze_bwd_kernel<<<pg, 256, sm_sq, st>>>(x, dy, ds, dx, B, F, T, lnw, lnb, Wsq, bsq, dWsq, dbsq, dlnw, dlnb, fsplit);
        NBSS_LAUNCH_CHECK();
    } else if ((rc = nbss_squeeze_bwd_tc(x, dy, ds, dx, B, F, T, lnw, lnb, Wsq, bsq, dWsq, dbsq, dlnw, dlnb, fmt, err, stream)) != NBSS_OK) return rc;
    return NBSS_OK;
}
