// online.cu — frame-by-frame (streaming) narrow-band block of the causal SpatialNet, fp32 CUDA-core kernels.
//
// BASELINE.json configs[4] / SURVEY.md §8 a15: models/arch/OnlineSpatialNet.py with attention='mhsa(N)'.  The reference has no
// chunked API (its inference=True loop re-walks the frames of a whole utterance, :156-168,203-221); this file is the
// state-in / state-out step the reference lacks: one call consumes ONE new frame of every (b,f) row and updates
//   * the attention key/value ring of the last `scope` frames  (CausalConv1d-free: order inside the ring is irrelevant to a softmax),
//   * the two previous frames of the three causal T-conv inputs (CausalConv1d(k=3), :22-60: left padding k-1),
//   * the four previous input frames of the causal encoder conv (k=5).
// The cross-band sub-blocks act on one frame by construction and reuse the tensor-core kernels of fconv_tc.cu / fullband_*.cu
// with T = 1.  One new frame is R = B*F rows x 96 channels — 355 kFLOP per row and layer: far too little for tcgen05 tiles, so
// these kernels are one CTA per row, weights streamed from L2 in [in][out] order (coalesced across the output threads), fp32
// throughout (parity with the oracle 1e-5 on the narrow-band part).  `pos` is a device-side frame counter so that a captured
// CUDA graph of the step can be replayed without host arguments changing.
//   nbss_online_encoder_step   Conv1d(k=5) over [x_{t-4} .. x_t]                                   OnlineSpatialNet.py:333,358
//   nbss_online_attn_step      LN -> q,k,v -> ring update -> softmax over the ring -> out-proj -> +x   :203-221
//   nbss_online_ffn_a_step     LN -> pw1 -> SiLU -> cconv -> SiLU -> cconv -> c2 (+ GroupNorm partial sums per row, group)   :223-240
//   nbss_online_gn_stats       (b, group) statistics over (24 channels x F) of the frame                :231-234
//   nbss_online_ffn_b_step     GN -> SiLU -> cconv -> SiLU -> pw2 -> +x
//   nbss_online_advance        pos += 1
#include "common.cuh"

namespace nbss {

constexpr int kOH = 96, kOHf = 192, kONH = 4, kODH = 24, kOG = 8, kOGC = 24;

__device__ __forceinline__ float block_reduce_sum(float v, float* red, int tid, int nthreads) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int w = 0; w < nthreads / 32; ++w) s += red[w];
    return s;
}

// LayerNorm of one 96-channel row held in smem row[96] -> ln[96]; all threads of the block call it
__device__ __forceinline__ void row_layernorm(const float* row, const float* __restrict__ g, const float* __restrict__ b, float* ln, float* red,
                                              int tid, int nthreads) {
    const float v = tid < kOH ? row[tid] : 0.f;
    const float mean = block_reduce_sum(v, red, tid, nthreads) * (1.f / kOH);
    const float d = tid < kOH ? v - mean : 0.f;
    const float var = block_reduce_sum(d * d, red, tid, nthreads) * (1.f / kOH);
    const float rstd = rsqrtf(var + 1e-5f);
    if (tid < kOH) ln[tid] = d * rstd * g[tid] + b[tid];
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------- encoder
// x_t [R, Cin], state [R, 4, Cin] (oldest first), Wt [5*Cin][96] (Wt[(k*Cin + c)][o] = W[o][c][k]) -> h [R, 96]; state shifts
__global__ void __launch_bounds__(96) online_encoder_kernel(const float* __restrict__ xt, float* state, const float* __restrict__ Wt,
                                                            const float* __restrict__ bias, float* __restrict__ h, int Cin) {
    __shared__ float win[5 * 16];
    const int r = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < 5 * Cin; i += 96) win[i] = i < 4 * Cin ? state[(size_t)r * 4 * Cin + i] : xt[(size_t)r * Cin + i - 4 * Cin];
    __syncthreads();
    float acc = bias[tid];
    for (int i = 0; i < 5 * Cin; ++i) acc = fmaf(win[i], Wt[i * kOH + tid], acc);
    h[(size_t)r * kOH + tid] = acc;
    for (int i = tid; i < 4 * Cin; i += 96) state[(size_t)r * 4 * Cin + i] = win[i + Cin];
}

// ---------------------------------------------------------------------------------------------------- attention
struct OnAttnArgs {
    float* x;               // [R, 96] in / out (residual added in place)
    const float *ln_w, *ln_b;
    const float* WinT;      // [96][288]  WinT[i][o] = in_proj_weight[o][i]
    const float* b_in;      // [288]
    const float* WoT;       // [96][96]   WoT[i][o] = out_proj.weight[o][i]
    const float* b_out;
    float *kcache, *vcache; // [R][scope][96]
    const int* pos;         // frames consumed so far
    int scope;
};
constexpr int kOnAttnThreads = 128;
constexpr int kOnMaxScope = 2048;  // 32 KB of scores in shared memory

__global__ void __launch_bounds__(kOnAttnThreads) online_attn_kernel(OnAttnArgs a) {
    __shared__ float row[kOH], ln[kOH], q[kOH], o[kOH], red[8];
    __shared__ float p[kONH][kOnMaxScope];
    __shared__ float hmax[kONH], hsum[kONH];
    const int r = blockIdx.x, tid = threadIdx.x, t = *a.pos;
    if (tid < kOH) row[tid] = a.x[(size_t)r * kOH + tid];
    __syncthreads();
    row_layernorm(row, a.ln_w, a.ln_b, ln, red, tid, kOnAttnThreads);
    // q | k | v = Win ln + b: 288 outputs over 128 threads
    const int slot = t % a.scope, n = min(t + 1, a.scope);
    float* kc = a.kcache + (size_t)r * a.scope * kOH;
    float* vc = a.vcache + (size_t)r * a.scope * kOH;
    for (int oi = tid; oi < 3 * kOH; oi += kOnAttnThreads) {
        float acc = a.b_in[oi];
#pragma unroll 8
        for (int i = 0; i < kOH; ++i) acc = fmaf(ln[i], a.WinT[i * 3 * kOH + oi], acc);
        if (oi < kOH) q[oi] = acc * rsqrtf((float)kODH);
        else if (oi < 2 * kOH) kc[(size_t)slot * kOH + oi - kOH] = acc;
        else vc[(size_t)slot * kOH + oi - 2 * kOH] = acc;
    }
    __syncthreads();  // also orders the ring writes of this block before its reads below (same block, global memory)
    // scores of the n cached keys, 4 heads each
    for (int j = tid; j < n; j += kOnAttnThreads) {
        const float* kj = kc + (size_t)j * kOH;
#pragma unroll
        for (int h = 0; h < kONH; ++h) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < kODH; ++c) s = fmaf(q[kODH * h + c], kj[kODH * h + c], s);
            p[h][j] = s;
        }
    }
    __syncthreads();
    // softmax per head: warp h reduces head h
    {
        const int h = tid >> 5, lane = tid & 31;
        float mx = -INFINITY;
        for (int j = lane; j < n; j += 32) mx = fmaxf(mx, p[h][j]);
        mx = warp_max(mx);
        float sm = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float e = __expf(p[h][j] - mx);
            p[h][j] = e;
            sm += e;
        }
        sm = warp_sum(sm);
        if (lane == 0) { hmax[h] = mx; hsum[h] = sm; }
    }
    __syncthreads();
    if (tid < kOH) {
        const int h = tid / kODH;
        float acc = 0.f;
        for (int j = 0; j < n; ++j) acc = fmaf(p[h][j], vc[(size_t)j * kOH + tid], acc);
        o[tid] = acc / hsum[h];
    }
    __syncthreads();
    if (tid < kOH) {
        float acc = a.b_out[tid];
#pragma unroll 8
        for (int i = 0; i < kOH; ++i) acc = fmaf(o[i], a.WoT[i * kOH + tid], acc);
        a.x[(size_t)r * kOH + tid] = row[tid] + acc;
    }
}

// ---------------------------------------------------------------------------------------------------- T-ConvFFN, part A
struct OnFfnAArgs {
    const float* x;          // [R, 96]
    const float *ln_w, *ln_b;
    const float* W1T;        // [96][192]
    const float* b1;
    const float *Wc1, *bc1, *Wc2, *bc2;  // grouped conv weights as stored by the reference: [192][24][3]
    float *st1, *st2;        // [R][2][192]: SiLU outputs of the two previous frames feeding conv1 / conv2 (older first)
    float* c2;               // [R][192] out
    float* part;             // [R][8][2] (sum, sum of squares) of c2 per conv group
};
__device__ __forceinline__ float cconv3(const float* prev2, const float* prev1, const float* cur, const float* __restrict__ W, int o) {
    // out[o] = sum_i sum_tap W[o][i][tap] * in_{t-2+tap}[24*(o/24) + i]
    const int g0 = kOGC * (o / kOGC);
    const float* w = W + (size_t)o * kOGC * 3;
    float acc = 0.f;
#pragma unroll 8
    for (int i = 0; i < kOGC; ++i)
        acc = fmaf(w[3 * i + 2], cur[g0 + i], fmaf(w[3 * i + 1], prev1[g0 + i], fmaf(w[3 * i], prev2[g0 + i], acc)));
    return acc;
}
__global__ void __launch_bounds__(192) online_ffn_a_kernel(OnFfnAArgs a) {
    __shared__ float row[kOH], ln[kOH], s1[3][kOHf], s2[3][kOHf], red[8];
    const int r = blockIdx.x, tid = threadIdx.x;
    if (tid < kOH) row[tid] = a.x[(size_t)r * kOH + tid];
    s1[0][tid] = a.st1[((size_t)r * 2 + 0) * kOHf + tid];
    s1[1][tid] = a.st1[((size_t)r * 2 + 1) * kOHf + tid];
    s2[0][tid] = a.st2[((size_t)r * 2 + 0) * kOHf + tid];
    s2[1][tid] = a.st2[((size_t)r * 2 + 1) * kOHf + tid];
    __syncthreads();
    row_layernorm(row, a.ln_w, a.ln_b, ln, red, tid, 192);
    {
        float acc = a.b1[tid];
#pragma unroll 8
        for (int i = 0; i < kOH; ++i) acc = fmaf(ln[i], a.W1T[i * kOHf + tid], acc);
        s1[2][tid] = acc / (1.f + __expf(-acc));
    }
    __syncthreads();
    {
        const float c1 = cconv3(s1[0], s1[1], s1[2], a.Wc1, tid) + a.bc1[tid];
        s2[2][tid] = c1 / (1.f + __expf(-c1));
    }
    __syncthreads();
    const float c2 = cconv3(s2[0], s2[1], s2[2], a.Wc2, tid) + a.bc2[tid];
    a.c2[(size_t)r * kOHf + tid] = c2;
    // GroupNorm partials: 24 consecutive threads form a group (not warp-aligned): smem tree per group
    __shared__ float gs[kOHf], gq[kOHf];
    gs[tid] = c2;
    gq[tid] = c2 * c2;
    __syncthreads();
    if (tid < kOG) {
        float s = 0.f, qq = 0.f;
        for (int i = 0; i < kOGC; ++i) { s += gs[kOGC * tid + i]; qq += gq[kOGC * tid + i]; }
        a.part[((size_t)r * kOG + tid) * 2] = s;
        a.part[((size_t)r * kOG + tid) * 2 + 1] = qq;
    }
    // shift the conv states
    a.st1[((size_t)r * 2 + 0) * kOHf + tid] = s1[1][tid];
    a.st1[((size_t)r * 2 + 1) * kOHf + tid] = s1[2][tid];
    a.st2[((size_t)r * 2 + 0) * kOHf + tid] = s2[1][tid];
    a.st2[((size_t)r * 2 + 1) * kOHf + tid] = s2[2][tid];
}

// (b, group) statistics over the F rows of the frame: part [B][F][8][2] -> stats [B][8][2] (mean, rstd), fp64 accumulation
__global__ void online_gn_stats_kernel(const float* __restrict__ part, int B, int F, float* stats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * kOG) return;
    const int b = i / kOG, g = i % kOG;
    double s = 0.0, q = 0.0;
    for (int f = 0; f < F; ++f) {
        s += (double)part[(((size_t)b * F + f) * kOG + g) * 2];
        q += (double)part[(((size_t)b * F + f) * kOG + g) * 2 + 1];
    }
    const double n = (double)F * kOGC, mean = s / n;
    double var = q / n - mean * mean;
    var = var > 0.0 ? var : 0.0;
    stats[2 * i] = (float)mean;
    stats[2 * i + 1] = (float)(1.0 / sqrt(var + 1e-5));
}

// ---------------------------------------------------------------------------------------------------- T-ConvFFN, part B
struct OnFfnBArgs {
    float* x;               // [R, 96] in / out
    const float* c2;        // [R][192]
    const float* stats;     // [B][8][2]
    const float *gn_w, *gn_b;
    const float *Wc3, *bc3;
    const float* W2T;       // [192][96]
    const float* b2;
    float* st3;             // [R][2][192]
    int F;
};
__global__ void __launch_bounds__(192) online_ffn_b_kernel(OnFfnBArgs a) {
    __shared__ float s3[3][kOHf], s4[kOHf];
    const int r = blockIdx.x, tid = threadIdx.x, b = r / a.F, g = tid / kOGC;
    s3[0][tid] = a.st3[((size_t)r * 2 + 0) * kOHf + tid];
    s3[1][tid] = a.st3[((size_t)r * 2 + 1) * kOHf + tid];
    const float mean = a.stats[((size_t)b * kOG + g) * 2], rstd = a.stats[((size_t)b * kOG + g) * 2 + 1];
    const float nrm = (a.c2[(size_t)r * kOHf + tid] - mean) * rstd * a.gn_w[tid] + a.gn_b[tid];
    s3[2][tid] = nrm / (1.f + __expf(-nrm));
    __syncthreads();
    const float c3 = cconv3(s3[0], s3[1], s3[2], a.Wc3, tid) + a.bc3[tid];
    s4[tid] = c3 / (1.f + __expf(-c3));
    __syncthreads();
    if (tid < kOH) {
        float acc = a.b2[tid];
#pragma unroll 8
        for (int i = 0; i < kOHf; ++i) acc = fmaf(s4[i], a.W2T[i * kOH + tid], acc);
        a.x[(size_t)r * kOH + tid] += acc;
    }
    a.st3[((size_t)r * 2 + 0) * kOHf + tid] = s3[1][tid];
    a.st3[((size_t)r * 2 + 1) * kOHf + tid] = s3[2][tid];
}

__global__ void online_advance_kernel(int* pos) { *pos += 1; }

// encoder.weight [96][Cin][5] -> Wt [(k*Cin + c)][96]
__global__ void online_pack_encoder_kernel(const float* __restrict__ W, float* __restrict__ Wt, int Cin) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 96 * Cin * 5) return;
    const int o = i / (Cin * 5), c = (i / 5) % Cin, kk = i % 5;
    Wt[(size_t)(kk * Cin + c) * 96 + o] = W[i];
}

// out[i][o] = in[o][i] for a [rows][cols] fp32 matrix (weight re-layout, once per weight update)
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int r = i / cols, c = i % cols;
    out[(size_t)c * rows + r] = in[i];
}

}  // namespace nbss

using namespace nbss;

// encoder.weight [96][Cin][5] -> Wt [5*Cin][96]; in_proj [288][96] -> [96][288]; out_proj / pw weights likewise
extern "C" int nbss_online_pack_encoder(const float* W, float* Wt, int Cin, void* stream) {
    if (!W || !Wt) return NBSS_ERR_NULL;
    if (Cin < 1 || Cin > 16) return NBSS_ERR_UNSUPPORTED;
    online_pack_encoder_kernel<<<(96 * Cin * 5 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(W, Wt, Cin);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
extern "C" int nbss_transpose(const float* in, float* out, int rows, int cols, void* stream) {
    if (!in || !out) return NBSS_ERR_NULL;
    if (rows < 1 || cols < 1) return NBSS_ERR_SHAPE;
    transpose_kernel<<<(rows * cols + 255) / 256, 256, 0, (cudaStream_t)stream>>>(in, out, rows, cols);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_encoder_step(const float* xt, float* state, const float* Wt, const float* bias, float* h, int R, int Cin,
                                        void* stream) {
    if (!xt || !state || !Wt || !bias || !h) return NBSS_ERR_NULL;
    if (R < 1 || Cin < 1 || Cin > 16) return NBSS_ERR_SHAPE;
    online_encoder_kernel<<<R, 96, 0, (cudaStream_t)stream>>>(xt, state, Wt, bias, h, Cin);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_attn_step(float* x, int R, const float* ln_w, const float* ln_b, const float* WinT, const float* b_in,
                                     const float* WoT, const float* b_out, float* kcache, float* vcache, const int* pos, int scope,
                                     void* stream) {
    if (!x || !ln_w || !ln_b || !WinT || !b_in || !WoT || !b_out || !kcache || !vcache || !pos) return NBSS_ERR_NULL;
    if (R < 1 || scope < 1 || scope > kOnMaxScope) return NBSS_ERR_SHAPE;
    OnAttnArgs a{x, ln_w, ln_b, WinT, b_in, WoT, b_out, kcache, vcache, pos, scope};
    online_attn_kernel<<<R, kOnAttnThreads, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_ffn_a_step(const float* x, int R, const float* ln_w, const float* ln_b, const float* W1T, const float* b1,
                                      const float* Wc1, const float* bc1, const float* Wc2, const float* bc2, float* st1, float* st2,
                                      float* c2, float* part, void* stream) {
    if (!x || !ln_w || !ln_b || !W1T || !b1 || !Wc1 || !bc1 || !Wc2 || !bc2 || !st1 || !st2 || !c2 || !part) return NBSS_ERR_NULL;
    if (R < 1) return NBSS_ERR_SHAPE;
    OnFfnAArgs a{x, ln_w, ln_b, W1T, b1, Wc1, bc1, Wc2, bc2, st1, st2, c2, part};
    online_ffn_a_kernel<<<R, 192, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_gn_stats(const float* part, int B, int F, float* stats, void* stream) {
    if (!part || !stats) return NBSS_ERR_NULL;
    if (B < 1 || F < 1) return NBSS_ERR_SHAPE;
    online_gn_stats_kernel<<<(B * kOG + 63) / 64, 64, 0, (cudaStream_t)stream>>>(part, B, F, stats);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_ffn_b_step(float* x, int R, int F, const float* c2, const float* stats, const float* gn_w, const float* gn_b,
                                      const float* Wc3, const float* bc3, const float* W2T, const float* b2, float* st3, void* stream) {
    if (!x || !c2 || !stats || !gn_w || !gn_b || !Wc3 || !bc3 || !W2T || !b2 || !st3) return NBSS_ERR_NULL;
    if (R < 1 || F < 1 || R % F) return NBSS_ERR_SHAPE;
    OnFfnBArgs a{x, c2, stats, gn_w, gn_b, Wc3, bc3, W2T, b2, st3, F};
    online_ffn_b_kernel<<<R, 192, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_advance(int* pos, void* stream) {
    if (!pos) return NBSS_ERR_NULL;
    online_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(pos);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
