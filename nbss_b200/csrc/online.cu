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
// these kernels are one CTA per row (one stream) or per 4 rows (many streams: each weight loaded from L2 feeds 4 rows), weights
// streamed from L2 in [in][out] order (coalesced across the output threads), fp32
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
constexpr int kOnFewRows = 600;  // up to ~2 CTAs per SM of single rows: below this, do not block rows

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
// Row blocking: a CTA owns RB consecutive rows and keeps RB accumulators per weight it loads, so the projection weights (147 KB
// per layer) are read from L2 once per RB rows; with many streams the kernel is then bound by the HBM read of the rings
// (2 x scope x 96 fp32 per row).  RB = 1 when there are too few rows to fill the GPU (one stream): shortest dependent chains.
struct OnAttnArgs {
    float* x;               // [R, 96] in / out (residual added in place)
    const float *ln_w, *ln_b;
    const float* WinT;      // [96][288]  WinT[i][o] = in_proj_weight[o][i]
    const float* b_in;      // [288]
    const float* WoT;       // [96][96]   WoT[i][o] = out_proj.weight[o][i]
    const float* b_out;
    float *kcache, *vcache; // [R][scope][96]
    const int* pos;         // frames consumed so far
    int scope, R;
};
constexpr int kOnAttnThreads = 384;
constexpr int kOnMaxScope = 2048;  // 32 KB of scores in shared memory (RB rows x 4 heads x scope <= 8192 floats)

// LayerNorm of row r (global, 96 channels) by one warp -> ln[96] (smem); rows beyond R give zeros
__device__ __forceinline__ void warp_layernorm(const float* __restrict__ x, int r, int R, const float* __restrict__ g, const float* __restrict__ b,
                                               float* ln, float* raw, int lane) {
    float v[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) v[k] = r < R ? x[(size_t)r * kOH + lane + 32 * k] : 0.f;
    const float mean = warp_sum(v[0] + v[1] + v[2]) * (1.f / kOH);
    float d[3], q = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) { d[k] = v[k] - mean; q = fmaf(d[k], d[k], q); }
    const float rstd = rsqrtf(warp_sum(q) * (1.f / kOH) + 1e-5f);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int c = lane + 32 * k;
        ln[c] = r < R ? d[k] * rstd * g[c] + b[c] : 0.f;
        if (raw) raw[c] = v[k];
    }
}

template <int RB>
__global__ void __launch_bounds__(kOnAttnThreads) online_attn_kernel(OnAttnArgs a) {
    constexpr int KS = 4 / RB;            // key splits of the P.V sum per row
    constexpr int PS = kOnMaxScope / RB;  // score slots per (row, head)
    __shared__ float row[RB][kOH], ln[RB][kOH], q[RB][kOH], o[RB][kOH];
    __shared__ float po[4][kOH];
    __shared__ float p[RB][kONH][PS];
    __shared__ float hsum[RB][kONH];
    const int r0 = blockIdx.x * RB, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, t = *a.pos;
    for (int j = warp; j < RB; j += kOnAttnThreads / 32) warp_layernorm(a.x, r0 + j, a.R, a.ln_w, a.ln_b, ln[j], row[j], lane);
    __syncthreads();
    const int slot = t % a.scope, n = min(t + 1, a.scope);
    const size_t ring = (size_t)a.scope * kOH;
    // q | k | v = Win ln + b: one output per thread, RB rows each
    if (tid < 3 * kOH) {
        float acc[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[j] = a.b_in[tid];
#pragma unroll 8
        for (int i = 0; i < kOH; ++i) {
            const float w = a.WinT[i * 3 * kOH + tid];
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[j] = fmaf(ln[j][i], w, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            if (tid < kOH) q[j][tid] = acc[j] * rsqrtf((float)kODH);
            else if (r0 + j < a.R) {
                if (tid < 2 * kOH) a.kcache[(size_t)(r0 + j) * ring + (size_t)slot * kOH + tid - kOH] = acc[j];
                else a.vcache[(size_t)(r0 + j) * ring + (size_t)slot * kOH + tid - 2 * kOH] = acc[j];
            }
        }
    }
    __syncthreads();  // also orders the ring writes of this block before its reads below (same block, global memory)
    // scores of the n cached keys of every row, 4 heads each
    for (int idx = tid; idx < RB * n; idx += kOnAttnThreads) {
        const int j = idx / n, key = idx - j * n;
        if (r0 + j >= a.R) { for (int h = 0; h < kONH; ++h) p[j][h][key] = 0.f; continue; }
        const float4* kj = reinterpret_cast<const float4*>(a.kcache + (size_t)(r0 + j) * ring + (size_t)key * kOH);
#pragma unroll
        for (int h = 0; h < kONH; ++h) {
            float sc = 0.f;
#pragma unroll
            for (int c = 0; c < kODH / 4; ++c) {
                const float4 kv = kj[(kODH / 4) * h + c];
                const float* qq = &q[j][kODH * h + 4 * c];
                sc = fmaf(qq[0], kv.x, sc); sc = fmaf(qq[1], kv.y, sc); sc = fmaf(qq[2], kv.z, sc); sc = fmaf(qq[3], kv.w, sc);
            }
            p[j][h][key] = sc;
        }
    }
    __syncthreads();
    // softmax per (row, head): one warp each
    for (int pr = warp; pr < RB * kONH; pr += kOnAttnThreads / 32) {
        float* ph = p[pr / kONH][pr % kONH];
        float mx = -INFINITY;
        for (int jj = lane; jj < n; jj += 32) mx = fmaxf(mx, ph[jj]);
        mx = warp_max(mx);
        float sm = 0.f;
        for (int jj = lane; jj < n; jj += 32) {
            const float e = __expf(ph[jj] - mx);
            ph[jj] = e;
            sm += e;
        }
        sm = warp_sum(sm);
        if (lane == 0) hsum[pr / kONH][pr % kONH] = sm;
    }
    __syncthreads();
    // P.V: thread (grp, c): row grp / KS, keys grp % KS, +KS, ...
    {
        const int grp = tid / kOH, c = tid % kOH, j = grp / KS, ks = grp % KS, h = c / kODH;
        float acc = 0.f;
        if (r0 + j < a.R) {
            const float* vc = a.vcache + (size_t)(r0 + j) * ring + c;
            const float* ph = p[j][h];
#pragma unroll 4
            for (int key = ks; key < n; key += KS) acc = fmaf(ph[key], vc[(size_t)key * kOH], acc);
        }
        po[grp][c] = acc;
    }
    __syncthreads();
    if (tid < RB * kOH) {
        const int j = tid / kOH, c = tid % kOH;
        float acc = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) acc += po[j * KS + ks][c];
        o[j][c] = acc / hsum[j][c / kODH];
    }
    __syncthreads();
    // out-proj: thread (kq, c) sums a quarter of the inputs for RB rows
    {
        const int kq = tid / kOH, c = tid % kOH;
        float acc[RB];
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[j] = 0.f;
#pragma unroll 8
        for (int i = 24 * kq; i < 24 * kq + 24; ++i) {
            const float w = a.WoT[i * kOH + c];
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[j] = fmaf(o[j][i], w, acc[j]);
        }
        __syncthreads();  // p is dead: reuse its first RB*4*96 floats for the partial sums
        float* part = &p[0][0][0];
#pragma unroll
        for (int j = 0; j < RB; ++j) part[(j * 4 + kq) * kOH + c] = acc[j];
    }
    __syncthreads();
    if (tid < RB * kOH) {
        const int j = tid / kOH, c = tid % kOH;
        const float* part = &p[0][0][0] + (size_t)j * 4 * kOH + c;
        if (r0 + j < a.R) a.x[(size_t)(r0 + j) * kOH + c] = row[j][c] + a.b_out[c] + ((part[0] + part[kOH]) + (part[2 * kOH] + part[3 * kOH]));
    }
}

// ---------------------------------------------------------------------------------------------------- T-ConvFFN, part A
struct OnFfnAArgs {
    const float* x;          // [R, 96]
    const float *ln_w, *ln_b;
    const float* W1T;        // [96][192]
    const float* b1;
    const float *Wc1T, *bc1, *Wc2T, *bc2;  // grouped conv weights transposed: WcT[(i*3 + tap)][o] = W[o][i][tap]  ([72][192])
    float *st1, *st2;        // [R][2][192]: SiLU outputs of the two previous frames feeding conv1 / conv2 (older first)
    float* c2;               // [R][192] out
    float* part;             // [R][8][2] (sum, sum of squares) of c2 per conv group
    int R;
};
// out[j][o] = bias[o] + sum_i sum_tap W[o][i][tap] * in_j[tap][24*(o/24) + i]   (tap 2 = the current frame)
template <int RB>
__device__ __forceinline__ void cconv3(const float (*s)[3][kOHf], const float* __restrict__ WT, float bias, int o, float* acc) {
    const int g0 = kOGC * (o / kOGC);
#pragma unroll
    for (int j = 0; j < RB; ++j) acc[j] = bias;
#pragma unroll 8
    for (int i = 0; i < kOGC; ++i) {
#pragma unroll
        for (int tap = 0; tap < 3; ++tap) {
            const float w = WT[(i * 3 + tap) * kOHf + o];
#pragma unroll
            for (int j = 0; j < RB; ++j) acc[j] = fmaf(w, s[j][tap][g0 + i], acc[j]);
        }
    }
}
__device__ __forceinline__ float silu_exact(float v) { return v / (1.f + __expf(-v)); }

template <int RB>
__global__ void __launch_bounds__(192) online_ffn_a_kernel(OnFfnAArgs a) {
    __shared__ float ln[RB][kOH], s1[RB][3][kOHf], s2[RB][3][kOHf];
    const int r0 = blockIdx.x * RB, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        const bool ok = r0 + j < a.R;
        const size_t base = (size_t)(r0 + j) * 2 * kOHf + tid;
        s1[j][0][tid] = ok ? a.st1[base] : 0.f;
        s1[j][1][tid] = ok ? a.st1[base + kOHf] : 0.f;
        s2[j][0][tid] = ok ? a.st2[base] : 0.f;
        s2[j][1][tid] = ok ? a.st2[base + kOHf] : 0.f;
    }
    for (int j = warp; j < RB; j += 6) warp_layernorm(a.x, r0 + j, a.R, a.ln_w, a.ln_b, ln[j], nullptr, lane);
    __syncthreads();
    float acc[RB];
#pragma unroll
    for (int j = 0; j < RB; ++j) acc[j] = a.b1[tid];
#pragma unroll 8
    for (int i = 0; i < kOH; ++i) {
        const float w = a.W1T[i * kOHf + tid];
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[j] = fmaf(ln[j][i], w, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) s1[j][2][tid] = silu_exact(acc[j]);
    __syncthreads();
    cconv3<RB>(s1, a.Wc1T, a.bc1[tid], tid, acc);
#pragma unroll
    for (int j = 0; j < RB; ++j) s2[j][2][tid] = silu_exact(acc[j]);
    __syncthreads();
    cconv3<RB>(s2, a.Wc2T, a.bc2[tid], tid, acc);
    // GroupNorm partials: 24 consecutive threads form a group (not warp-aligned): through smem
    __shared__ float gs[RB][kOHf];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        gs[j][tid] = acc[j];
        if (r0 + j < a.R) a.c2[(size_t)(r0 + j) * kOHf + tid] = acc[j];
    }
    __syncthreads();
    if (tid < RB * kOG) {
        const int j = tid / kOG, g = tid % kOG;
        float sm = 0.f, qq = 0.f;
        for (int i = 0; i < kOGC; ++i) { const float v = gs[j][kOGC * g + i]; sm += v; qq = fmaf(v, v, qq); }
        if (r0 + j < a.R) {
            a.part[((size_t)(r0 + j) * kOG + g) * 2] = sm;
            a.part[((size_t)(r0 + j) * kOG + g) * 2 + 1] = qq;
        }
    }
    // shift the conv states
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        if (r0 + j >= a.R) break;
        const size_t base = (size_t)(r0 + j) * 2 * kOHf + tid;
        a.st1[base] = s1[j][1][tid];
        a.st1[base + kOHf] = s1[j][2][tid];
        a.st2[base] = s2[j][1][tid];
        a.st2[base + kOHf] = s2[j][2][tid];
    }
}

// (b, group) statistics over the F rows of the frame: part [B][F][8][2] -> stats [B][8][2] (mean, rstd); one warp per (b, group),
// fp64 accumulation
__global__ void online_gn_stats_kernel(const float* __restrict__ part, int B, int F, float* stats) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= B * kOG) return;
    const int b = i / kOG, g = i % kOG;
    double s = 0.0, q = 0.0;
    for (int f = lane; f < F; f += 32) {
        const float2 v = *reinterpret_cast<const float2*>(part + (((size_t)b * F + f) * kOG + g) * 2);
        s += (double)v.x;
        q += (double)v.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        const double n = (double)F * kOGC, mean = s / n;
        double var = q / n - mean * mean;
        var = var > 0.0 ? var : 0.0;
        stats[2 * i] = (float)mean;
        stats[2 * i + 1] = (float)(1.0 / sqrt(var + 1e-5));
    }
}

// ---------------------------------------------------------------------------------------------------- T-ConvFFN, part B
struct OnFfnBArgs {
    float* x;               // [R, 96] in / out
    const float* c2;        // [R][192]
    const float* stats;     // [B][8][2]
    const float *gn_w, *gn_b;
    const float *Wc3T, *bc3;  // [72][192], see OnFfnAArgs
    const float* W2T;       // [192][96]
    const float* b2;
    float* st3;             // [R][2][192]
    int F, R;
};
template <int RB>
__global__ void __launch_bounds__(192) online_ffn_b_kernel(OnFfnBArgs a) {
    __shared__ float s3[RB][3][kOHf], s4[RB][kOHf], half[RB][kOH];
    const int r0 = blockIdx.x * RB, tid = threadIdx.x, g = tid / kOGC;
    const float gw = a.gn_w[tid], gb = a.gn_b[tid];
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        const int r = r0 + j;
        const bool ok = r < a.R;
        const size_t base = (size_t)r * 2 * kOHf + tid;
        s3[j][0][tid] = ok ? a.st3[base] : 0.f;
        s3[j][1][tid] = ok ? a.st3[base + kOHf] : 0.f;
        float v = 0.f;
        if (ok) {
            const int b = r / a.F;
            const float mean = a.stats[((size_t)b * kOG + g) * 2], rstd = a.stats[((size_t)b * kOG + g) * 2 + 1];
            v = silu_exact((a.c2[(size_t)r * kOHf + tid] - mean) * rstd * gw + gb);
        }
        s3[j][2][tid] = v;
    }
    __syncthreads();
    float acc[RB];
    cconv3<RB>(s3, a.Wc3T, a.bc3[tid], tid, acc);
#pragma unroll
    for (int j = 0; j < RB; ++j) s4[j][tid] = silu_exact(acc[j]);
    __syncthreads();
    // pw2: thread (kh, c) sums half of the 192 inputs
    const int kh = tid / kOH, c = tid % kOH;
#pragma unroll
    for (int j = 0; j < RB; ++j) acc[j] = 0.f;
#pragma unroll 8
    for (int i = kOH * kh; i < kOH * kh + kOH; ++i) {
        const float w = a.W2T[i * kOH + c];
#pragma unroll
        for (int j = 0; j < RB; ++j) acc[j] = fmaf(s4[j][i], w, acc[j]);
    }
    if (kh == 1) {
#pragma unroll
        for (int j = 0; j < RB; ++j) half[j][c] = acc[j];
    }
    __syncthreads();
    if (kh == 0) {
        const float b2 = a.b2[c];
#pragma unroll
        for (int j = 0; j < RB; ++j)
            if (r0 + j < a.R) a.x[(size_t)(r0 + j) * kOH + c] += b2 + (acc[j] + half[j][c]);
    }
#pragma unroll
    for (int j = 0; j < RB; ++j) {
        if (r0 + j >= a.R) break;
        const size_t base = (size_t)(r0 + j) * 2 * kOHf + tid;
        a.st3[base] = s3[j][1][tid];
        a.st3[base + kOHf] = s3[j][2][tid];
    }
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
    OnAttnArgs a{x, ln_w, ln_b, WinT, b_in, WoT, b_out, kcache, vcache, pos, scope, R};
    // few rows (one stream): one row per CTA, shortest chains; many rows: 4 (or 2 for long rings: RB * scope <= 2048) rows per CTA
    const int rb = R <= kOnFewRows ? 1 : (scope <= kOnMaxScope / 4 ? 4 : (scope <= kOnMaxScope / 2 ? 2 : 1));
    if (rb == 4) online_attn_kernel<4><<<(R + 3) / 4, kOnAttnThreads, 0, (cudaStream_t)stream>>>(a);
    else if (rb == 2) online_attn_kernel<2><<<(R + 1) / 2, kOnAttnThreads, 0, (cudaStream_t)stream>>>(a);
    else online_attn_kernel<1><<<R, kOnAttnThreads, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_ffn_a_step(const float* x, int R, const float* ln_w, const float* ln_b, const float* W1T, const float* b1,
                                      const float* Wc1, const float* bc1, const float* Wc2, const float* bc2, float* st1, float* st2,
                                      float* c2, float* part, void* stream) {
    if (!x || !ln_w || !ln_b || !W1T || !b1 || !Wc1 || !bc1 || !Wc2 || !bc2 || !st1 || !st2 || !c2 || !part) return NBSS_ERR_NULL;
    if (R < 1) return NBSS_ERR_SHAPE;
    OnFfnAArgs a{x, ln_w, ln_b, W1T, b1, Wc1, bc1, Wc2, bc2, st1, st2, c2, part, R};
    if (R <= kOnFewRows) online_ffn_a_kernel<1><<<R, 192, 0, (cudaStream_t)stream>>>(a);
    else online_ffn_a_kernel<4><<<(R + 3) / 4, 192, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_gn_stats(const float* part, int B, int F, float* stats, void* stream) {
    if (!part || !stats) return NBSS_ERR_NULL;
    if (B < 1 || F < 1) return NBSS_ERR_SHAPE;
    online_gn_stats_kernel<<<(B * kOG + 3) / 4, 128, 0, (cudaStream_t)stream>>>(part, B, F, stats);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_ffn_b_step(float* x, int R, int F, const float* c2, const float* stats, const float* gn_w, const float* gn_b,
                                      const float* Wc3, const float* bc3, const float* W2T, const float* b2, float* st3, void* stream) {
    if (!x || !c2 || !stats || !gn_w || !gn_b || !Wc3 || !bc3 || !W2T || !b2 || !st3) return NBSS_ERR_NULL;
    if (R < 1 || F < 1 || R % F) return NBSS_ERR_SHAPE;
    OnFfnBArgs a{x, c2, stats, gn_w, gn_b, Wc3, bc3, W2T, b2, st3, F, R};
    if (R <= kOnFewRows) online_ffn_b_kernel<1><<<R, 192, 0, (cudaStream_t)stream>>>(a);
    else online_ffn_b_kernel<4><<<(R + 3) / 4, 192, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_online_advance(int* pos, void* stream) {
    if (!pos) return NBSS_ERR_NULL;
    online_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(pos);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
