// loss.cu — negative SI-SDR with permutation-invariant training over 2 speakers (the step right after
// the hot path).  Replaces Loss.forward for loss_func = neg_si_sdr, pit = True (models/io/loss.py:21-29,95-118;
// configs/SpatialNet.yaml:33-37), i.e. torchmetrics.functional.audio scale_invariant_signal_distortion_ratio and
// permutation_invariant_training(mode="permutation-wise", eval_func="min") — third-party, absent from the reference tree
// and unpinned (requirements.txt:2); the published algorithm is restated in oracle/spatialnet_oracle.py:
//     alpha = (<p,t> + eps) / (<t,t> + eps);  val = 10 log10((|alpha t|^2 + eps) / (|alpha t - p|^2 + eps))
//     loss[b] = min over speaker permutations of  -mean_s val(est[b, perm(s)], ref[b, s]);   loss = mean_b loss[b]
// Everything follows from 12 inner products per utterance, so the forward is one streaming reduction + a tiny finalize
// kernel; the gradient wrt est is a per-utterance linear combination  c1 * ref[j] + c2 * est[i]  (one streaming kernel).
#include <algorithm>

#include "common.cuh"

namespace nbss {

constexpr int kLossS = 2;      // speakers
constexpr int kLossSums = 12;  // e0 e1 r0 r1 | e0e0 e1e1 r0r0 r1r1 | e0r0 e0r1 e1r0 e1r1
constexpr int kLossChunk = 4096;

// partial sums of one chunk of one utterance -> sums[b][12] (double atomics)
__global__ void __launch_bounds__(256) sisdr_sums_kernel(const float* __restrict__ est, const float* __restrict__ ref, int B,
                                                         long long Ts, double* sums) {
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n0 = (long long)blockIdx.x * kLossChunk, n1 = min(Ts, n0 + kLossChunk);
    const float* e0 = est + ((long long)b * kLossS + 0) * Ts;
    const float* e1 = est + ((long long)b * kLossS + 1) * Ts;
    const float* r0 = ref + ((long long)b * kLossS + 0) * Ts;
    const float* r1 = ref + ((long long)b * kLossS + 1) * Ts;
    float s[kLossSums];
#pragma unroll
    for (int i = 0; i < kLossSums; ++i) s[i] = 0.f;
    for (long long n = n0 + tid; n < n1; n += 256) {
        const float a0 = e0[n], a1 = e1[n], b0 = r0[n], b1 = r1[n];
        s[0] += a0; s[1] += a1; s[2] += b0; s[3] += b1;
        s[4] = fmaf(a0, a0, s[4]); s[5] = fmaf(a1, a1, s[5]); s[6] = fmaf(b0, b0, s[6]); s[7] = fmaf(b1, b1, s[7]);
        s[8] = fmaf(a0, b0, s[8]); s[9] = fmaf(a0, b1, s[9]); s[10] = fmaf(a1, b0, s[10]); s[11] = fmaf(a1, b1, s[11]);
    }
    __shared__ float red[8][kLossSums];
#pragma unroll
    for (int i = 0; i < kLossSums; ++i) {
        const float v = warp_sum(s[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (tid < kLossSums) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)red[w][tid];
        atomicAdd(sums + (size_t)b * kLossSums + tid, t);
    }
}

// per utterance: the four SI-SDR values, the better permutation, its loss; coef[b][i] = (c1, c2, ref index j, mean-e, mean-r)
// for the gradient  d loss / d est[b,i,n] = c1 * (ref[b,j,n] - mr) + c2 * (est[b,i,n] - me);  loss_out[0] = mean_b loss[b]
__global__ void sisdr_finalize_kernel(const double* __restrict__ sums, int B, long long Ts, int zero_mean, float eps,
                                      float* loss_out, float* loss_b, int* perm_out, float* coef) {
    __shared__ float s_loss[256];
    float acc = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const double* q = sums + (size_t)b * kLossSums;
        const double inv = 1.0 / (double)Ts;
        double P[2], T[2], C[2][2], me[2], mr[2];
        for (int i = 0; i < 2; ++i) {
            me[i] = zero_mean ? q[i] * inv : 0.0;
            mr[i] = zero_mean ? q[2 + i] * inv : 0.0;
            P[i] = q[4 + i] - (zero_mean ? q[i] * q[i] * inv : 0.0);
            T[i] = q[6 + i] - (zero_mean ? q[2 + i] * q[2 + i] * inv : 0.0);
        }
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) C[i][j] = q[8 + 2 * i + j] - (zero_mean ? q[i] * q[2 + j] * inv : 0.0);
        double val[2][2], c1[2][2], c2[2][2];
        const double K = 10.0 / log(10.0), e = (double)eps;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                const double alpha = (C[i][j] + e) / (T[j] + e);
                const double S = alpha * alpha * T[j];
                // |alpha t - p|^2 from the Gram sums; >= 0 in exact arithmetic, but an estimate that is an exact scaled copy
                // of the target cancels to -1e-12-ish, and log() of that would make the permutation compare with NaN
                const double N = fmax(S - 2.0 * alpha * C[i][j] + P[i], 0.0);
                val[i][j] = K * (log(S + e) - log(N + e));
                // d val / d p_n = K * (a1 t_n + a2 p_n)
                const double a1 = 2.0 * alpha * T[j] / ((T[j] + e) * (S + e)) - ((2.0 * alpha * T[j] - 2.0 * C[i][j]) / (T[j] + e) - 2.0 * alpha) / (N + e);
                const double a2 = -2.0 / (N + e);
                c1[i][j] = K * a1;
                c2[i][j] = K * a2;
            }
        const double l0 = -0.5 * (val[0][0] + val[1][1]);  // identity permutation: est i <-> ref i
        const double l1 = -0.5 * (val[0][1] + val[1][0]);  // swapped
        const int swap = l1 < l0;                          // torch.min keeps the first permutation on ties
        const double lb = swap ? l1 : l0;
        if (loss_b) loss_b[b] = (float)lb;
        if (perm_out) { perm_out[2 * b] = swap ? 1 : 0; perm_out[2 * b + 1] = swap ? 0 : 1; }  // ref index matched to est 0, 1
        acc += (float)lb;
        if (coef) {
            const double g = -0.5 / (double)B;  // d loss / d val of a chosen pair
            for (int i = 0; i < 2; ++i) {
                const int j = swap ? 1 - i : i;
                float* c = coef + ((size_t)b * 2 + i) * 5;
                c[0] = (float)(g * c1[i][j]);
                c[1] = (float)(g * c2[i][j]);
                c[2] = (float)j;
                c[3] = (float)me[i];
                c[4] = (float)mr[j];
            }
        }
    }
    s_loss[threadIdx.x] = acc;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s_loss[threadIdx.x] += s_loss[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) loss_out[0] = s_loss[0] / (float)B;
}

// dest[b,i,n] = gout * (c1 * (ref[b,j,n] - mr) + c2 * (est[b,i,n] - me))
__global__ void __launch_bounds__(256) sisdr_bwd_kernel(const float* __restrict__ est, const float* __restrict__ ref,
                                                        const float* __restrict__ coef, const float* __restrict__ gout, float* dest,
                                                        long long Ts) {
    const int bi = blockIdx.y, b = bi >> 1;
    const float* c = coef + (size_t)bi * 5;
    const float g = gout ? gout[0] : 1.f;
    const float c1 = g * c[0], c2 = g * c[1], me = c[3], mr = c[4];
    const int j = (int)c[2];
    const float* e = est + (long long)bi * Ts;
    const float* r = ref + ((long long)b * kLossS + j) * Ts;
    float* d = dest + (long long)bi * Ts;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < Ts; n += (long long)gridDim.x * blockDim.x)
        d[n] = c1 * (r[n] - mr) + c2 * (e[n] - me);
}

}  // namespace nbss

using namespace nbss;

// est, ref: [B,2,Ts] fp32.  sums: workspace of B*12 doubles.  Outputs: loss[1] (mean over the batch), loss_b[B] (nullable),
// perm[B,2] int (nullable; ref index matched to each estimate), coef[B,2,5] (nullable; for nbss_sisdr_pit_bwd).
extern "C" int nbss_sisdr_pit_fwd(const float* est, const float* ref, int B, int S, long long Ts, int zero_mean, double* sums,
                                  float* loss, float* loss_b, int* perm, float* coef, void* stream) {
    if (!est || !ref || !sums || !loss) return NBSS_ERR_NULL;
    if (B < 1 || Ts < 1) return NBSS_ERR_SHAPE;
    if (S != kLossS) return NBSS_ERR_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)B * kLossSums * sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    sisdr_sums_kernel<<<dim3((unsigned)((Ts + kLossChunk - 1) / kLossChunk), B), 256, 0, st>>>(est, ref, B, Ts, sums);
    NBSS_LAUNCH_CHECK();
    sisdr_finalize_kernel<<<1, 256, 0, st>>>(sums, B, Ts, zero_mean, 1.1920929e-07f /* torch.finfo(float32).eps */, loss, loss_b, perm, coef);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// dest [B,2,Ts] = gout[0] (device scalar, nullable = 1) * d loss / d est, from the coefficients of the forward call.
extern "C" int nbss_sisdr_pit_bwd(const float* est, const float* ref, const float* coef, const float* gout, float* dest, int B,
                                  int S, long long Ts, void* stream) {
    if (!est || !ref || !coef || !dest) return NBSS_ERR_NULL;
    if (B < 1 || Ts < 1) return NBSS_ERR_SHAPE;
    if (S != kLossS) return NBSS_ERR_UNSUPPORTED;
    const unsigned gx = (unsigned)min((long long)64, (Ts + 255) / 256);
    sisdr_bwd_kernel<<<dim3(gx, B * kLossS), 256, 0, (cudaStream_t)stream>>>(est, ref, coef, gout, dest, Ts);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// ------------------------------------------------------------------------------------------------ clip + Adam
// The optimiser tail of the step: torch.nn.utils.clip_grad_norm_(max_norm) followed by
// torch.optim.Adam (configs/SpatialNet.yaml:3-4,44; general_steps.py:243-271: Adam, lr 1e-3, gradient_clip_val 5) over the ONE
// flat fp32 gradient buffer of the network (spatialnet.py: make_flat_grads) in two launches instead of ~40.
namespace nbss {

__global__ void __launch_bounds__(256) gradnorm_kernel(const float* __restrict__ g, long long n, double* out, float* step) {
    float s = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) s = fmaf(g[i], g[i], s);
    s = warp_sum(s);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)red[w];
        atomicAdd(out, t);
        if (blockIdx.x == 0) step[0] += 1.f;  // the update kernel (next in the stream) reads the incremented step count
    }
}

struct AdamHyper { float max_norm, lr, beta1, beta2, eps; };

// element i of the flat buffer belongs to tensor t with off[t] <= i < off[t+1]; its parameter is ptrs[t][i - off[t]]
__global__ void __launch_bounds__(256) clip_adam_kernel(float* const* __restrict__ ptrs, const long long* __restrict__ off,
                                                        int ntensors, long long n, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const double* __restrict__ gnorm_sq, const float* __restrict__ step,
                                                        AdamHyper h) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    int lo = 0, hi = ntensors;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    const float total = (float)sqrt(gnorm_sq[0]);
    const float clip = fminf(1.f, h.max_norm / (total + 1e-6f));  // torch.nn.utils.clip_grad_norm_
    const float t = step[0];
    const float bc1 = 1.f - powf(h.beta1, t), bc2 = 1.f - powf(h.beta2, t);
    const float gi = g[i] * clip;
    const float mi = h.beta1 * m[i] + (1.f - h.beta1) * gi;
    const float vi = h.beta2 * v[i] + (1.f - h.beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + h.eps;
    float* p = ptrs[lo] + (i - off[lo]);
    *p -= (h.lr / bc1) * (mi / denom);
}

}  // namespace nbss

// params: device array of `ntensors` parameter pointers; offsets: device array of ntensors+1 cumulative element offsets into
// the flat buffers (offsets[ntensors] = n); flat_grad, exp_avg, exp_avg_sq: fp32 [n]; gnorm_sq: device double (workspace /
// output: squared total gradient norm BEFORE clipping); step: device float step counter (incremented by this call).
extern "C" int nbss_clip_adam(float* const* params, const long long* offsets, int ntensors, long long n, const float* flat_grad,
                              float* exp_avg, float* exp_avg_sq, double* gnorm_sq, float* step, float max_norm, float lr,
                              float beta1, float beta2, float eps, void* stream) {
    if (!params || !offsets || !flat_grad || !exp_avg || !exp_avg_sq || !gnorm_sq || !step) return NBSS_ERR_NULL;
    if (ntensors < 1 || n < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(gnorm_sq, 0, sizeof(double), st);
    if (e != cudaSuccess) return (int)e;
    const unsigned gb = (unsigned)std::min<long long>(1184, (n + 255) / 256);
    nbss::gradnorm_kernel<<<gb, 256, 0, st>>>(flat_grad, n, gnorm_sq, step);
    NBSS_LAUNCH_CHECK();
    nbss::clip_adam_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(params, offsets, ntensors, n, flat_grad, exp_avg, exp_avg_sq,
                                                                         gnorm_sq, step, nbss::AdamHyper{max_norm, lr, beta1, beta2, eps});
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// ------------------------------------------------------------------------------------------------ exact loss scaling of dy
// The backward kernels carry 16-bit gradient operands and are linear in the upstream gradient: dy is multiplied by the power of two
// that brings its largest element to ~1 and the flat gradient buffer is multiplied back (spatialnet.py _SpatialNetFn.backward).
// No host synchronisation: the scale lives on the device.   ws: one zeroed uint (bit pattern of max |dy|), scale: [2] = (s, 1/s).
namespace nbss {
__global__ void __launch_bounds__(256) absmax_kernel(const float* __restrict__ x, long long n, unsigned int* amax_bits) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) m = fmaxf(m, fabsf(x[i]));
    m = warp_max(m);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
        atomicMax(amax_bits, __float_as_uint(m));  // non-negative floats order like their bit patterns (NaN sorts above everything)
    }
}
__device__ __forceinline__ float pow2_scale(unsigned int amax_bits) {
    const float amax = __uint_as_float(amax_bits);
    return (amax > 0.f && amax < INFINITY) ? exp2f(-rintf(log2f(fmaxf(amax, 1e-30f)))) : 1.f;
}
__global__ void __launch_bounds__(256) prescale_kernel(const float* __restrict__ x, float* __restrict__ y, long long n,
                                                       const unsigned int* __restrict__ amax_bits, float* scale) {
    const float s = pow2_scale(*amax_bits);
    if (blockIdx.x == 0 && threadIdx.x == 0) { scale[0] = s; scale[1] = 1.f / s; }
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) y[i] = x[i] * s;
}
__global__ void __launch_bounds__(256) unscale_kernel(float* __restrict__ g, long long n, const float* __restrict__ scale) {
    const float r = scale[1];
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) g[i] *= r;
}
}  // namespace nbss

extern "C" int nbss_grad_prescale(const float* dy, long long n, float* dy_scaled, unsigned int* ws, float* scale, void* stream) {
    if (!dy || !dy_scaled || !ws || !scale) return NBSS_ERR_NULL;
    if (n < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(unsigned int), st);
    if (e != cudaSuccess) return (int)e;
    nbss::absmax_kernel<<<grid, 256, 0, st>>>(dy, n, ws);
    NBSS_LAUNCH_CHECK();
    nbss::prescale_kernel<<<grid, 256, 0, st>>>(dy, dy_scaled, n, ws, scale);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
extern "C" int nbss_grad_unscale(float* flat, long long n, const float* scale, void* stream) {
    if (!flat || !scale) return NBSS_ERR_NULL;
    if (n < 1) return NBSS_ERR_SHAPE;
    const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
    nbss::unscale_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(flat, n, scale);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
