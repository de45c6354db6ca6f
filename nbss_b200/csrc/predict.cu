// predict.cu — test/predict-time post-processing of the separated waveforms (the caller right after the
// hot path at inference).  Replaces, for a scale-invariant loss (configs/SpatialNet.yaml: neg_si_sdr),
//     recover_scale(preds, mixture, scale_src_together=False, norm_if_exceed_1=False)   models/utils/metrics.py:192-218
//         a = argmin || preds^T a - mixture ||  per utterance (torch.linalg.lstsq), preds *= a
//     peak normalisation   SharedTrainer.py:301-305:  preds /= max(1, max_n |preds|)  per (utterance, speaker)
// (the permutation re-ordering of SharedTrainer.py:297-299 is the PIT of csrc/loss.cu).  Three launches over [B,S,Ts], S <= 4:
// Gram / cross sums (fp64 atomics), a per-utterance S x S normal-equation solve in fp64 (the system is tiny and, for
// separated sources, well conditioned; lstsq's QR and the normal equations agree to fp32 rounding, tests/test_gpu_predict.py)
// fused with the scaling and the running peak, and the final division.
#include "common.cuh"

namespace nbss {

constexpr int kPrS = 4;                                  // max speakers
constexpr int kPrSums = kPrS * (kPrS + 1) / 2 + kPrS;    // Gram upper triangle + cross terms with the mixture
constexpr int kPrChunk = 4096;

__global__ void __launch_bounds__(256) predict_sums_kernel(const float* __restrict__ preds, const float* __restrict__ mix, int S,
                                                           long long Ts, double* sums) {
    const int b = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n0 = (long long)blockIdx.x * kPrChunk, n1 = min(Ts, n0 + kPrChunk);
    float s[kPrSums];
#pragma unroll
    for (int i = 0; i < kPrSums; ++i) s[i] = 0.f;
    for (long long n = n0 + tid; n < n1; n += 256) {
        float p[kPrS];
#pragma unroll
        for (int i = 0; i < kPrS; ++i) p[i] = i < S ? preds[((long long)b * S + i) * Ts + n] : 0.f;
        const float x = mix[(long long)b * Ts + n];
        int k = 0;
#pragma unroll
        for (int i = 0; i < kPrS; ++i)
#pragma unroll
            for (int j = i; j < kPrS; ++j) s[k] = fmaf(p[i], p[j], s[k]), ++k;
#pragma unroll
        for (int i = 0; i < kPrS; ++i) s[k + i] = fmaf(p[i], x, s[k + i]);
    }
    __shared__ float red[8][kPrSums];
#pragma unroll
    for (int i = 0; i < kPrSums; ++i) {
        const float v = warp_sum(s[i]);
        if (lane == 0) red[warp][i] = v;
    }
    __syncthreads();
    if (tid < kPrSums) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += (double)red[w][tid];
        atomicAdd(sums + (size_t)b * kPrSums + tid, t);
    }
}

// solve G a = c (S x S, Gaussian elimination with partial pivoting, fp64) per utterance; out = preds * a; peak[b,s] = max |out|
__global__ void __launch_bounds__(256) predict_scale_kernel(const float* __restrict__ preds, const double* __restrict__ sums, int S,
                                                            long long Ts, float* __restrict__ out, float* scale_out, unsigned int* peak) {
    const int b = blockIdx.y, tid = threadIdx.x;
    __shared__ double s_a[kPrS];
    if (tid == 0) {
        const double* q = sums + (size_t)b * kPrSums;
        double G[kPrS][kPrS + 1];
        int k = 0;
        for (int i = 0; i < kPrS; ++i)
            for (int j = i; j < kPrS; ++j) { G[i][j] = G[j][i] = q[k]; ++k; }
        for (int i = 0; i < kPrS; ++i) G[i][kPrS] = q[k + i];
        for (int c = 0; c < S; ++c) {
            int piv = c;
            for (int r = c + 1; r < S; ++r)
                if (fabs(G[r][c]) > fabs(G[piv][c])) piv = r;
            for (int j = 0; j <= kPrS; ++j) { const double t = G[c][j]; G[c][j] = G[piv][j]; G[piv][j] = t; }
            const double d = G[c][c];
            for (int r = c + 1; r < S; ++r) {
                const double f = d != 0.0 ? G[r][c] / d : 0.0;
                for (int j = c; j <= kPrS; ++j) G[r][j] -= f * G[c][j];
            }
        }
        for (int c = S - 1; c >= 0; --c) {
            double t = G[c][kPrS];
            for (int j = c + 1; j < S; ++j) t -= G[c][j] * s_a[j];
            s_a[c] = G[c][c] != 0.0 ? t / G[c][c] : 0.0;
        }
        if (scale_out && blockIdx.x == 0)
            for (int i = 0; i < S; ++i) scale_out[(size_t)b * S + i] = (float)s_a[i];
    }
    __syncthreads();
    const long long n0 = (long long)blockIdx.x * kPrChunk, n1 = min(Ts, n0 + kPrChunk);
    for (int i = 0; i < S; ++i) {
        const float a = (float)s_a[i];
        float mx = 0.f;
        for (long long n = n0 + tid; n < n1; n += 256) {
            const float v = preds[((long long)b * S + i) * Ts + n] * a;
            out[((long long)b * S + i) * Ts + n] = v;
            mx = fmaxf(mx, fabsf(v));
        }
        mx = warp_max(mx);
        if ((tid & 31) == 0) atomicMax(peak + (size_t)b * S + i, __float_as_uint(mx));  // non-negative floats order like their bits
    }
}

__global__ void __launch_bounds__(256) predict_peaknorm_kernel(float* __restrict__ out, const unsigned int* __restrict__ peak, long long Ts) {
    const int bs = blockIdx.y;
    const float mx = __uint_as_float(peak[bs]);
    if (!(mx > 1.f)) return;
    const float inv = 1.f / mx;
    const long long n0 = (long long)blockIdx.x * kPrChunk, n1 = min(Ts, n0 + kPrChunk);
    for (long long n = n0 + threadIdx.x; n < n1; n += 256) out[(long long)bs * Ts + n] *= inv;
}

}  // namespace nbss

// preds [B,S,Ts], mixture [B,Ts] -> out [B,S,Ts] (may alias preds); ws: B*(14 doubles) + B*S uints, zeroed by this call;
// scale_out (nullable) [B,S] receives the least-squares scales; recover != 0 applies them; norm_if_exceed_1 != 0 divides every
// (b,s) row by max(1, peak).
extern "C" int nbss_predict_post(const float* preds, const float* mixture, float* out, int B, int S, long long Ts, int recover,
                                 int norm_if_exceed_1, double* ws_sums, unsigned int* ws_peak, float* scale_out, void* stream) {
    using namespace nbss;
    if (!preds || !out || !ws_sums || !ws_peak || (recover && !mixture)) return NBSS_ERR_NULL;
    if (B < 1 || S < 1 || S > kPrS || Ts < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(ws_sums, 0, sizeof(double) * kPrSums * B, st);
    cudaMemsetAsync(ws_peak, 0, sizeof(unsigned int) * B * S, st);
    const dim3 grid((unsigned)((Ts + kPrChunk - 1) / kPrChunk), B);
    if (recover) {
        predict_sums_kernel<<<grid, 256, 0, st>>>(preds, mixture, S, Ts, ws_sums);
        NBSS_LAUNCH_CHECK();
    } else {
        // identity scales: G = I, c = 1
        return NBSS_ERR_UNSUPPORTED;
    }
    predict_scale_kernel<<<grid, 256, 0, st>>>(preds, ws_sums, S, Ts, out, scale_out, ws_peak);
    NBSS_LAUNCH_CHECK();
    if (norm_if_exceed_1) {
        predict_peaknorm_kernel<<<dim3(grid.x, B * S), 256, 0, st>>>(out, ws_peak, Ts);
        NBSS_LAUNCH_CHECK();
    }
    return NBSS_OK;
}
