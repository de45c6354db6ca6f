// io.cu — framing and the thin ends of the network, fp32 CUDA-core kernels (all HBM-bound, SURVEY.md §8 a1-a4,a10,a11):
//   stft (+ optional Norm 'frequency/online' + pack)   models/io/stft.py:49-66, models/io/norm.py:75-81,94,
//                                                      SharedTrainer.py:116-117
//   encoder  Conv1d(Cin->96, k=5, 'same') along T      models/arch/SpatialNet.py:175,205   (+ wgrad)
//   decoder  Linear(96->Cout)                          models/arch/SpatialNet.py:200,216   (+ dgrad, wgrad)
//   istft (+ optional inverse norm + unpack)           models/io/stft.py:68-97, models/io/norm.py:97-108,
//                                                      SharedTrainer.py:121-128            (+ backward)
// The reference loops torch.istft over B*S items in Python (stft.py:83-87); here it is one launch.
#include "common.cuh"
#include "layout.cuh"

namespace nbss {

constexpr int kFT = 8;  // frames per CTA in the framing kernels

__device__ __forceinline__ float hann(int n, int N) { return 0.5f - 0.5f * cospif(2.f * (float)n / (float)N); }

// ------------------------------------------------------------------------------------------------ STFT
struct StftArgs {
    const float* x;  // [B,C,Ts]
    int B, C, Ts, T, F, N, hop;
    int normalize, ref;  // Norm(mode='frequency', online=True) fused when normalize != 0
    float eps;
    float* out;  // re at out + b*ob + c*oc + f*of + t*ot, im at +1 (strides in floats)
    long long ob, oc, of, ot;
    float* xrmm;  // [B,F,T] or null
    float* xr;    // [B,F,T,2] (reference-channel STFT before normalisation) or null
};

__global__ void __launch_bounds__(256) stft_kernel(StftArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F, C = a.C;
    float2* tw = reinterpret_cast<float2*>(sm);             // [N] (cos, sin)(2 pi k / N)
    float* xw = sm + 2 * N;                                 // [kFT][C][N] windowed frames
    float2* X = reinterpret_cast<float2*>(xw + kFT * C * N);  // [kFT][C][F]
    const int tiles = (a.T + kFT - 1) / kFT, tid = threadIdx.x;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kFT;
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < kFT * C * N; i += 256) {
        const int n = i % N, c = (i / N) % C, tt = i / (N * C), t = t0 + tt;
        float v = 0.f;
        if (t < a.T) {
            int j = t * a.hop + n - N / 2;  // center=True, reflect padding
            if (j < 0) j = -j;
            if (j >= a.Ts) j = 2 * (a.Ts - 1) - j;
            v = a.x[((size_t)b * C + c) * a.Ts + j] * hann(n, N);
        }
        xw[i] = v;
    }
    __syncthreads();
    for (int i = tid; i < kFT * C * F; i += 256) {
        const int f = i % F, tc = i / F;
        const float* fr = xw + (size_t)tc * N;
        float re = 0.f, im = 0.f;
        for (int n = 0; n < N; ++n) {
            const float2 w = tw[(f * n) & (N - 1)];
            re = fmaf(fr[n], w.x, re);
            im = fmaf(-fr[n], w.y, im);
        }
        X[i] = make_float2(re, im);
    }
    __syncthreads();
    for (int i = tid; i < kFT * C * F; i += 256) {
        // order (f, tt, c): consecutive threads write consecutive channels / frames of one frequency
        const int c = i % C, tt = (i / C) % kFT, f = i / (C * kFT), t = t0 + tt;
        if (t >= a.T) continue;
        float2 v = X[((size_t)tt * C + c) * F + f];
        if (a.normalize) {
            const float2 r = X[((size_t)tt * C + a.ref) * F + f];
            const float mm = sqrtf(r.x * r.x + r.y * r.y) + a.eps;
            v.x /= mm;
            v.y /= mm;
            if (c == a.ref) {
                const size_t o = ((size_t)b * F + f) * a.T + t;
                if (a.xrmm) a.xrmm[o] = mm;
                if (a.xr) { a.xr[2 * o] = r.x; a.xr[2 * o + 1] = r.y; }
            }
        }
        float* o = a.out + b * a.ob + c * a.oc + f * a.of + t * a.ot;
        o[0] = v.x;
        o[1] = v.y;
    }
}

// ------------------------------------------------------------------------------------------------ iSTFT
struct IstftArgs {
    const float* in;  // re at in + b*ib + s*is + f*if_ + t*it, im at +1 (strides in floats)
    long long ib, is, if_, it;
    const float* scale;  // optional [B,F,T] multiplier (inverse normalisation), or null
    float* y;            // [B,S,Ts]
    int B, S, Ts, T, F, N, hop;
};

// squared-window envelope at padded position p (frames t with 0 <= p - t*hop < N, 0 <= t < T)
__device__ __forceinline__ float ola_env(int p, int N, int hop, int T) {
    float e = 0.f;
    const int thi = min(T - 1, p / hop);
    for (int t = thi; t >= 0 && p - t * hop < N; --t) {
        const float w = hann(p - t * hop, N);
        e += w * w;
    }
    return e;
}

__global__ void __launch_bounds__(256) istft_kernel(IstftArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F, R = N / a.hop, NFR = kFT + R - 1;
    float2* tw = reinterpret_cast<float2*>(sm);                 // [N]
    float2* X = reinterpret_cast<float2*>(sm + 2 * N);          // [NFR][F]
    float* fr = sm + 2 * N + 2 * NFR * F;                       // [NFR][N] windowed inverse DFT of each frame
    const int nseg = a.T + R - 1, tiles = (nseg + kFT - 1) / kFT, tid = threadIdx.x;
    const int bs = blockIdx.x / tiles, k0 = (blockIdx.x % tiles) * kFT;
    const int b = bs / a.S, s = bs % a.S;
    const int tfirst = k0 - (R - 1);
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < NFR * F; i += 256) {
        const int f = i % F, t = tfirst + i / F;
        float2 v = make_float2(0.f, 0.f);
        if (t >= 0 && t < a.T) {
            const float* p = a.in + b * a.ib + s * a.is + f * a.if_ + t * a.it;
            const float sc = a.scale ? a.scale[((size_t)b * F + f) * a.T + t] : 1.f;
            v = make_float2(p[0] * sc, p[1] * sc);
        }
        X[i] = v;
    }
    __syncthreads();
    const float invN = 1.f / N;
    for (int i = tid; i < NFR * N; i += 256) {
        const int n = i % N, fi = i / N;
        const float2* xf = X + (size_t)fi * F;
        float acc = xf[0].x + ((n & 1) ? -xf[F - 1].x : xf[F - 1].x);  // DC + Nyquist (imaginary parts ignored)
        float acc2 = 0.f;
        for (int f = 1; f < F - 1; ++f) {
            const float2 w = tw[(f * n) & (N - 1)];
            acc2 = fmaf(xf[f].x, w.x, acc2);
            acc2 = fmaf(-xf[f].y, w.y, acc2);
        }
        fr[i] = (acc + 2.f * acc2) * invN * hann(n, N);
    }
    __syncthreads();
    for (int i = tid; i < kFT * a.hop; i += 256) {
        const int k = k0 + i / a.hop, p = k * a.hop + i % a.hop, j = p - N / 2;
        if (k >= nseg || j < 0 || j >= a.Ts) continue;
        float v = 0.f;
        for (int t = max(0, k - R + 1); t <= min(k, a.T - 1); ++t) v += fr[(size_t)(t - tfirst) * N + p - t * a.hop];
        const float e = ola_env(p, N, a.hop, a.T);
        a.y[((size_t)b * a.S + s) * a.Ts + j] = e > 1e-11f ? v / e : 0.f;
    }
}

// Backward of iSTFT(+scale): d in[b,s,f,t] (re, im) from dy [B,S,Ts]:  a windowed, envelope-normalised DFT of dy.
struct IstftBwdArgs {
    const float* dy;  // [B,S,Ts]
    const float* scale;
    float* din;  // re at din + b*ib + s*is + f*if_ + t*it, im at +1
    long long ib, is, if_, it;
    int B, S, Ts, T, F, N, hop;
};

__global__ void __launch_bounds__(256) istft_bwd_kernel(IstftBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F;
    float2* tw = reinterpret_cast<float2*>(sm);  // [N]
    float* g = sm + 2 * N;                       // [kFT][N]  dy * w / env per frame
    const int tiles = (a.T + kFT - 1) / kFT, tid = threadIdx.x;
    const int bs = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kFT;
    const int b = bs / a.S, s = bs % a.S;
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < kFT * N; i += 256) {
        const int n = i % N, t = t0 + i / N, p = t * a.hop + n, j = p - N / 2;
        float v = 0.f;
        if (t < a.T && j >= 0 && j < a.Ts) {
            const float e = ola_env(p, N, a.hop, a.T);
            if (e > 1e-11f) v = a.dy[((size_t)b * a.S + s) * a.Ts + j] * hann(n, N) / e;
        }
        g[i] = v;
    }
    __syncthreads();
    const float invN = 1.f / N;
    for (int i = tid; i < kFT * F; i += 256) {
        const int tt = i % kFT, f = i / kFT, t = t0 + tt;
        if (t >= a.T) continue;
        const float* gr = g + (size_t)tt * N;
        float re = 0.f, im = 0.f;
        for (int n = 0; n < N; ++n) {
            const float2 w = tw[(f * n) & (N - 1)];
            re = fmaf(gr[n], w.x, re);
            im = fmaf(-gr[n], w.y, im);
        }
        const bool edge = (f == 0 || f == F - 1);
        const float c = (edge ? 1.f : 2.f) * invN * (a.scale ? a.scale[((size_t)b * F + f) * a.T + t] : 1.f);
        float* o = a.din + b * a.ib + s * a.is + f * a.if_ + t * a.it;
        o[0] = re * c;
        o[1] = edge ? 0.f : im * c;
    }
}

// ------------------------------------------------------------------------------------------------ FFT framing
// n_fft = 4^k (64, 256, 1024): radix-4 Stockham FFT in shared memory instead of the direct DFT above.  Two real frames ride
// in one complex transform (z = a + i b; A[f] = (Z[f] + conj Z[N-f]) / 2, B[f] = (Z[f] - conj Z[N-f]) / (2i)).
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// `nb` transforms of length N in `a` ([nb][N]); `b` is scratch of the same size.  tw[k] = (cos, sin)(2 pi k / N).
// Forward: X[k] = sum x[n] e^{-2 pi i k n / N}; INV: the conjugate kernel, not normalised.  Returns the result buffer.
template <bool INV>
__device__ float2* fft_r4_batch(float2* a, float2* b, int nb, int N, const float2* tw, int tid, int nthreads) {
    const int Q = N >> 2;
    for (int Ns = 1; Ns < N; Ns <<= 2) {
        const int tstep = N / (4 * Ns);
        for (int task = tid; task < nb * Q; task += nthreads) {
            const int fi = task / Q, j = task - fi * Q;
            const float2* src = a + (size_t)fi * N;
            float2* dst = b + (size_t)fi * N;
            const int k = j & (Ns - 1), m = k * tstep;
            float2 w1 = tw[m], w2 = tw[2 * m], w3 = tw[3 * m];
            if (!INV) { w1.y = -w1.y; w2.y = -w2.y; w3.y = -w3.y; }
            const float2 v0 = src[j], v1 = cmul(src[j + Q], w1), v2 = cmul(src[j + 2 * Q], w2), v3 = cmul(src[j + 3 * Q], w3);
            const float2 A = make_float2(v0.x + v2.x, v0.y + v2.y), B = make_float2(v0.x - v2.x, v0.y - v2.y);
            const float2 C = make_float2(v1.x + v3.x, v1.y + v3.y), D = make_float2(v1.x - v3.x, v1.y - v3.y);
            const int j0 = ((j - k) << 2) + k;  // (j / Ns) * Ns * 4 + k
            dst[j0] = make_float2(A.x + C.x, A.y + C.y);
            dst[j0 + 2 * Ns] = make_float2(A.x - C.x, A.y - C.y);
            // forward: y1 = B - i D, y3 = B + i D; inverse: the other way round
            const float2 m1 = make_float2(B.x + D.y, B.y - D.x), p1 = make_float2(B.x - D.y, B.y + D.x);
            dst[j0 + Ns] = INV ? p1 : m1;
            dst[j0 + 3 * Ns] = INV ? m1 : p1;
        }
        __syncthreads();
        float2* t = a; a = b; b = t;
    }
    return a;
}

constexpr int kFTF = 4;  // frames per CTA of the FFT STFT kernel

__global__ void __launch_bounds__(256) stft_fft_kernel(StftArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F, C = a.C, NF = kFTF * C, NB = (NF + 1) / 2;
    float2* tw = reinterpret_cast<float2*>(sm);   // [N]
    float2* za = tw + N;                           // [NB][N]
    float2* zb = za + (size_t)NB * N;              // [NB][N]
    float2* X = zb + (size_t)NB * N;               // [kFTF][C][F]
    const int tiles = (a.T + kFTF - 1) / kFTF, tid = threadIdx.x;
    const int b = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kFTF;
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < NB * N; i += 256) {
        const int p = i / N, n = i - p * N;
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int q = 2 * p + h, tt = q / C, c = q - tt * C, t = t0 + tt;
            if (q < NF && t < a.T) {
                int j = t * a.hop + n - N / 2;  // center=True, reflect padding
                if (j < 0) j = -j;
                if (j >= a.Ts) j = 2 * (a.Ts - 1) - j;
                v[h] = a.x[((size_t)b * C + c) * a.Ts + j];
            }
        }
        const float w = hann(n, N);
        za[i] = make_float2(v[0] * w, v[1] * w);
    }
    __syncthreads();
    const float2* Z = fft_r4_batch<false>(za, zb, NB, N, tw, tid, 256);
    for (int i = tid; i < NF * F; i += 256) {
        const int q = i / F, f = i - q * F;
        const float2 z1 = Z[(size_t)(q >> 1) * N + f], z2 = Z[(size_t)(q >> 1) * N + ((N - f) & (N - 1))];
        X[i] = (q & 1) ? make_float2(0.5f * (z1.y + z2.y), -0.5f * (z1.x - z2.x)) : make_float2(0.5f * (z1.x + z2.x), 0.5f * (z1.y - z2.y));
    }
    __syncthreads();
    for (int i = tid; i < kFTF * C * F; i += 256) {
        // order (f, tt, c): consecutive threads write consecutive channels / frames of one frequency
        const int c = i % C, tt = (i / C) % kFTF, f = i / (C * kFTF), t = t0 + tt;
        if (t >= a.T) continue;
        float2 v = X[((size_t)tt * C + c) * F + f];
        if (a.normalize) {
            const float2 r = X[((size_t)tt * C + a.ref) * F + f];
            const float mm = sqrtf(r.x * r.x + r.y * r.y) + a.eps;
            v.x /= mm;
            v.y /= mm;
            if (c == a.ref) {
                const size_t o = ((size_t)b * F + f) * a.T + t;
                if (a.xrmm) a.xrmm[o] = mm;
                if (a.xr) { a.xr[2 * o] = r.x; a.xr[2 * o + 1] = r.y; }
            }
        }
        float* o = a.out + b * a.ob + c * a.oc + f * a.of + t * a.ot;
        o[0] = v.x;
        o[1] = v.y;
    }
}

__global__ void __launch_bounds__(256) istft_fft_kernel(IstftArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F, R = N / a.hop, NFR = kFT + R - 1, NB = (NFR + 1) / 2;
    float2* tw = reinterpret_cast<float2*>(sm);                 // [N]
    float2* X = tw + N;                                         // [NFR][F]
    float2* za = X + (size_t)NFR * F;                           // [NB][N]
    float2* zb = za + (size_t)NB * N;                           // [NB][N]
    float* fr = reinterpret_cast<float*>(zb + (size_t)NB * N);  // [NFR][N] windowed inverse DFT of each frame
    const int nseg = a.T + R - 1, tiles = (nseg + kFT - 1) / kFT, tid = threadIdx.x;
    const int bs = blockIdx.x / tiles, k0 = (blockIdx.x % tiles) * kFT;
    const int b = bs / a.S, s = bs % a.S;
    const int tfirst = k0 - (R - 1);
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < NFR * F; i += 256) {
        const int f = i % F, t = tfirst + i / F;
        float2 v = make_float2(0.f, 0.f);
        if (t >= 0 && t < a.T) {
            const float* p = a.in + b * a.ib + s * a.is + f * a.if_ + t * a.it;
            const float sc = a.scale ? a.scale[((size_t)b * F + f) * a.T + t] : 1.f;
            v = make_float2(p[0] * sc, (f == 0 || f == F - 1) ? 0.f : p[1] * sc);  // DC / Nyquist: imaginary parts ignored
        }
        X[i] = v;
    }
    __syncthreads();
    // Hermitian extension of two frames packed as Ya + i Yb
    for (int i = tid; i < NB * N; i += 256) {
        const int p = i / N, k = i - p * N, kk = k <= N / 2 ? k : N - k;
        const float sg = k <= N / 2 ? 1.f : -1.f;
        const float2 ya = X[(size_t)(2 * p) * F + kk];
        const float2 yb = (2 * p + 1 < NFR) ? X[(size_t)(2 * p + 1) * F + kk] : make_float2(0.f, 0.f);
        za[i] = make_float2(ya.x - sg * yb.y, sg * ya.y + yb.x);
    }
    __syncthreads();
    const float2* Z = fft_r4_batch<true>(za, zb, NB, N, tw, tid, 256);
    const float invN = 1.f / N;
    for (int i = tid; i < NFR * N; i += 256) {
        const int n = i % N, fi = i / N;
        const float2 z = Z[(size_t)(fi >> 1) * N + n];
        fr[i] = ((fi & 1) ? z.y : z.x) * invN * hann(n, N);
    }
    __syncthreads();
    for (int i = tid; i < kFT * a.hop; i += 256) {
        const int k = k0 + i / a.hop, p = k * a.hop + i % a.hop, j = p - N / 2;
        if (k >= nseg || j < 0 || j >= a.Ts) continue;
        float v = 0.f;
        for (int t = max(0, k - R + 1); t <= min(k, a.T - 1); ++t) v += fr[(size_t)(t - tfirst) * N + p - t * a.hop];
        const float e = ola_env(p, N, a.hop, a.T);
        a.y[((size_t)b * a.S + s) * a.Ts + j] = e > 1e-11f ? v / e : 0.f;
    }
}

__global__ void __launch_bounds__(256) istft_bwd_fft_kernel(IstftBwdArgs a) {
    extern __shared__ __align__(16) float sm[];
    const int N = a.N, F = a.F, NB = kFT / 2;
    float2* tw = reinterpret_cast<float2*>(sm);  // [N]
    float2* za = tw + N;                          // [NB][N]: frames 2p (re) and 2p+1 (im): dy * w / env
    float2* zb = za + (size_t)NB * N;
    const int tiles = (a.T + kFT - 1) / kFT, tid = threadIdx.x;
    const int bs = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * kFT;
    const int b = bs / a.S, s = bs % a.S;
    for (int k = tid; k < N; k += 256) tw[k] = make_float2(cospif(2.f * k / N), sinpif(2.f * k / N));
    for (int i = tid; i < NB * N; i += 256) {
        const int p = i / N, n = i - p * N;
        float v[2] = {0.f, 0.f};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int t = t0 + 2 * p + h, pp = t * a.hop + n, j = pp - N / 2;
            if (t < a.T && j >= 0 && j < a.Ts) {
                const float e = ola_env(pp, N, a.hop, a.T);
                if (e > 1e-11f) v[h] = a.dy[((size_t)b * a.S + s) * a.Ts + j] * hann(n, N) / e;
            }
        }
        za[i] = make_float2(v[0], v[1]);
    }
    __syncthreads();
    const float2* Z = fft_r4_batch<false>(za, zb, NB, N, tw, tid, 256);
    const float invN = 1.f / N;
    for (int i = tid; i < kFT * F; i += 256) {
        const int tt = i % kFT, f = i / kFT, t = t0 + tt;
        if (t >= a.T) continue;
        const float2 z1 = Z[(size_t)(tt >> 1) * N + f], z2 = Z[(size_t)(tt >> 1) * N + ((N - f) & (N - 1))];
        const float re = (tt & 1) ? 0.5f * (z1.y + z2.y) : 0.5f * (z1.x + z2.x);
        const float im = (tt & 1) ? -0.5f * (z1.x - z2.x) : 0.5f * (z1.y - z2.y);
        const bool edge = (f == 0 || f == F - 1);
        const float c = (edge ? 1.f : 2.f) * invN * (a.scale ? a.scale[((size_t)b * F + f) * a.T + t] : 1.f);
        float* o = a.din + b * a.ib + s * a.is + f * a.if_ + t * a.it;
        o[0] = re * c;
        o[1] = edge ? 0.f : im * c;
    }
}

// ------------------------------------------------------------------------------------------------ encoder
// y[b,f,t,co] = bias[co] + sum_k sum_ci W[co,ci,k] * x[b,f,t+k-K/2,ci]     (K = 5, zero padding)
template <int CIN>
__global__ void __launch_bounds__(192) encoder_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int nslab, int T,
                                                          const float* W, const float* bias) {
    constexpr int K = 5, TT = 64;
    __shared__ __align__(16) float xs[(TT + K - 1) * CIN];
    const int tiles = (T + TT - 1) / TT, tid = threadIdx.x;
    const int slab = blockIdx.x / tiles, t0 = (blockIdx.x % tiles) * TT;
    for (int i = tid; i < (TT + K - 1) * CIN; i += 192) {
        const int t = t0 + i / CIN - K / 2;
        xs[i] = (t >= 0 && t < T) ? x[((size_t)slab * T + t) * CIN + i % CIN] : 0.f;
    }
    const int co = tid % kH, half = tid / kH;
    float w[CIN * K];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int k = 0; k < K; ++k) w[k * CIN + ci] = W[(co * CIN + ci) * K + k];
    const float bi = bias[co];
    __syncthreads();
    for (int r = half; r < TT; r += 2) {
        if (t0 + r >= T) break;
        float acc = bi;
        if constexpr (CIN % 4 == 0) {  // the K*CIN window of frame r is 16-byte aligned: one LDS.128 per four taps
            const float4* xr = reinterpret_cast<const float4*>(xs + r * CIN);
#pragma unroll
            for (int i = 0; i < K * CIN / 4; ++i) {
                const float4 v = xr[i];
                acc = fmaf(w[4 * i + 0], v.x, acc);
                acc = fmaf(w[4 * i + 1], v.y, acc);
                acc = fmaf(w[4 * i + 2], v.z, acc);
                acc = fmaf(w[4 * i + 3], v.w, acc);
            }
        } else {
#pragma unroll
            for (int i = 0; i < K * CIN; ++i) acc = fmaf(w[i], xs[r * CIN + i], acc);
        }
        y[((size_t)slab * T + t0 + r) * kH + co] = acc;
    }
}

// dW[co,ci,k] += sum dy[.,t,co] x[.,t+k-2,ci];  dbias[co] += sum dy.   (the network input needs no gradient)
template <int CIN>
__global__ void __launch_bounds__(192) encoder_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, int nslab,
                                                            int T, float* dW, float* dbias) {
    constexpr int K = 5, TT = 64;
    __shared__ __align__(16) float xs[(TT + K - 1) * CIN];
    const int tiles = (T + TT - 1) / TT, tid = threadIdx.x;
    const int co = tid % kH, half = tid / kH;
    float dw[CIN * K];
#pragma unroll
    for (int i = 0; i < CIN * K; ++i) dw[i] = 0.f;
    float db = 0.f;
    for (int tile = blockIdx.x; tile < nslab * tiles; tile += gridDim.x) {
        const int slab = tile / tiles, t0 = (tile % tiles) * TT;
        __syncthreads();
        for (int i = tid; i < (TT + K - 1) * CIN; i += 192) {
            const int t = t0 + i / CIN - K / 2;
            xs[i] = (t >= 0 && t < T) ? x[((size_t)slab * T + t) * CIN + i % CIN] : 0.f;
        }
        __syncthreads();
        // four frames in flight per thread: the dy loads of a batch are issued before their FMAs
        for (int r0 = half; r0 < TT; r0 += 8) {
            float g[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + 2 * u;
                g[u] = (r < TT && t0 + r < T) ? dy[((size_t)slab * T + t0 + r) * kH + co] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int r = r0 + 2 * u;
                if (r >= TT) break;
                db += g[u];
                if constexpr (CIN % 4 == 0) {
                    const float4* xr = reinterpret_cast<const float4*>(xs + r * CIN);
#pragma unroll
                    for (int i = 0; i < K * CIN / 4; ++i) {
                        const float4 v = xr[i];
                        dw[4 * i + 0] = fmaf(g[u], v.x, dw[4 * i + 0]);
                        dw[4 * i + 1] = fmaf(g[u], v.y, dw[4 * i + 1]);
                        dw[4 * i + 2] = fmaf(g[u], v.z, dw[4 * i + 2]);
                        dw[4 * i + 3] = fmaf(g[u], v.w, dw[4 * i + 3]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < K * CIN; ++i) dw[i] = fmaf(g[u], xs[r * CIN + i], dw[i]);
                }
            }
        }
    }
    // the two frame halves of a channel are summed through shared memory first: one atomic per (CTA, weight); the weights of
    // this layer sit in 183 L2 lines, so every lane-level atomic saved shortens the serialised tail of the launch
    __shared__ float s_half[kH * (CIN * K + 1)];
    __syncthreads();
    if (half == 1) {
#pragma unroll
        for (int i = 0; i < CIN * K; ++i) s_half[i * kH + co] = dw[i];
        s_half[CIN * K * kH + co] = db;
    }
    __syncthreads();
    if (half == 0) {
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
            for (int k = 0; k < K; ++k) atomicAdd(dW + (co * CIN + ci) * K + k, dw[k * CIN + ci] + s_half[(k * CIN + ci) * kH + co]);
        atomicAdd(dbias + co, db + s_half[CIN * K * kH + co]);
    }
}

// ------------------------------------------------------------------------------------------------ decoder
// y[n, o] = sum_c x[n,c] W[o,c] + b[o], o < COUT.  Warp per row.
template <int COUT>
__global__ void __launch_bounds__(256) decoder_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, size_t n,
                                                          const float* W, const float* bias) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * 256) >> 5;
    const bool act = lane < 24;
    float4 w[COUT];
#pragma unroll
    for (int o = 0; o < COUT; ++o) w[o] = act ? ld_f4(W + o * kH + 4 * lane) : make_float4(0, 0, 0, 0);
    for (size_t r = warp; r < n; r += nw) {
        const float4 v = act ? ld_f4(x + r * kH + 4 * lane) : make_float4(0, 0, 0, 0);
        float outv = 0.f;
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
            const float p = warp_sum(v.x * w[o].x + v.y * w[o].y + v.z * w[o].z + v.w * w[o].w);
            if (lane == o) outv = p + bias[o];
        }
        if (lane < COUT) y[r * COUT + lane] = outv;
    }
}

template <int COUT>
__global__ void __launch_bounds__(256) decoder_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                          float* __restrict__ dx, size_t n, const float* W, float* dW,
                                                          float* dbias) {
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * 256 + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * 256) >> 5;
    const bool act = lane < 24;
    float4 w[COUT], dw[COUT];
    float db = 0.f;
#pragma unroll
    for (int o = 0; o < COUT; ++o) {
        w[o] = act ? ld_f4(W + o * kH + 4 * lane) : make_float4(0, 0, 0, 0);
        dw[o] = make_float4(0, 0, 0, 0);
    }
    for (size_t r0 = warp; r0 < n; r0 += 4 * nw) {  // four rows in flight per warp
        float4 v[4];
        float gl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t r = r0 + u * nw;
            v[u] = (act && r < n) ? ld_f4(x + r * kH + 4 * lane) : make_float4(0, 0, 0, 0);
            gl[u] = (lane < COUT && r < n) ? dy[r * COUT + lane] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t r = r0 + u * nw;
            db += gl[u];
            float4 d = make_float4(0, 0, 0, 0);
#pragma unroll
            for (int o = 0; o < COUT; ++o) {
                const float g = __shfl_sync(0xffffffffu, gl[u], o);
                d = make_float4(fmaf(g, w[o].x, d.x), fmaf(g, w[o].y, d.y), fmaf(g, w[o].z, d.z), fmaf(g, w[o].w, d.w));
                dw[o] = make_float4(fmaf(g, v[u].x, dw[o].x), fmaf(g, v[u].y, dw[o].y), fmaf(g, v[u].z, dw[o].z), fmaf(g, v[u].w, dw[o].w));
            }
            if (act && r < n) st_f4(dx + r * kH + 4 * lane, d);
        }
    }
    // the CTA's eight warps are summed through shared memory: one atomic per (CTA, weight) instead of one per warp — the
    // whole layer is COUT*96 + COUT floats (13 L2 lines for COUT = 4), and lane-level atomics on so few lines serialise
    __shared__ float s_red[8][COUT * kH + COUT];
    const int w8 = threadIdx.x >> 5;
    if (act) {
#pragma unroll
        for (int o = 0; o < COUT; ++o) *reinterpret_cast<float4*>(&s_red[w8][o * kH + 4 * lane]) = dw[o];
    }
    if (lane < COUT) s_red[w8][COUT * kH + lane] = db;
    __syncthreads();
    for (int i = threadIdx.x; i < COUT * kH + COUT; i += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_red[w][i];
        atomicAdd(i < COUT * kH ? dW + i : dbias + (i - COUT * kH), t);
    }
}

// ------------------------------------------------------------------------------------------------ Norm (standalone)
// Norm.norm, mode='frequency', online=True (models/io/norm.py:75-81,94) on complex [B,C,F,T]: XrMM = |X[:,ref]| + eps,
// X /= XrMM (in place, like the reference), Xr = copy of the reference channel before normalisation.
__global__ void __launch_bounds__(256) norm_freq_kernel(float2* X, int B, int C, size_t FT, int ref, float eps, float* xrmm, float2* xr) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)B * FT) return;
    const size_t b = i / FT, r = i % FT;
    const float2 v = X[(b * C + ref) * FT + r];
    const float mm = sqrtf(v.x * v.x + v.y * v.y) + eps;
    xrmm[i] = mm;
    xr[i] = v;
    for (int c = 0; c < C; ++c) {
        float2 w = X[(b * C + c) * FT + r];
        w.x /= mm;
        w.y /= mm;
        X[(b * C + c) * FT + r] = w;
    }
}
// Norm.inorm (models/io/norm.py:97-108): Y[b,s,:] = X[b,s,:] * XrMM[b,:]
__global__ void __launch_bounds__(256) inorm_kernel(const float2* X, float2* Y, int B, int S, size_t FT, const float* xrmm) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= (size_t)B * S * FT) return;
    const size_t b = i / (S * FT), r = i % FT;
    const float m = xrmm[b * FT + r];
    const float2 v = X[i];
    Y[i] = make_float2(v.x * m, v.y * m);
}

static int io_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

}  // namespace nbss

using namespace nbss;

static bool pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
static bool pow4(int n) { return pow2(n) && (n & 0x55555555); }  // 4^k: the radix-4 FFT kernels

extern "C" int nbss_stft(const float* x, int B, int C, int Ts, int n_fft, int hop, int normalize, int ref_channel,
                         float eps, float* out, long long ob, long long oc, long long of, long long ot, float* xrmm,
                         float* xr, void* stream) {
    if (!x || !out) return NBSS_ERR_NULL;
    if (!pow2(n_fft) || n_fft > 1024 || hop < 1 || n_fft % hop || B < 1 || C < 1 || Ts <= n_fft / 2) return NBSS_ERR_SHAPE;
    if (normalize && (ref_channel < 0 || ref_channel >= C)) return NBSS_ERR_SHAPE;
    const int T = 1 + Ts / hop, F = n_fft / 2 + 1;
    const size_t smem = (size_t)(2 * n_fft + kFT * C * n_fft + 2 * kFT * C * F) * 4;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    StftArgs a{x, B, C, Ts, T, F, n_fft, hop, normalize, ref_channel, eps, out, ob, oc, of, ot, xrmm, xr};
    if (pow4(n_fft)) {
        const int NB = (kFTF * C + 1) / 2;
        const size_t sm_f = (size_t)(n_fft + 2 * NB * n_fft + kFTF * C * F) * 8;
        if (sm_f <= 227 * 1024) {
            cudaError_t ef = cudaFuncSetAttribute(stft_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f);
            if (ef != cudaSuccess) return (int)ef;
            stft_fft_kernel<<<B * ((T + kFTF - 1) / kFTF), 256, sm_f, (cudaStream_t)stream>>>(a);
            NBSS_LAUNCH_CHECK();
            return NBSS_OK;
        }
    }
    cudaError_t e = cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    stft_kernel<<<B * ((T + kFT - 1) / kFT), 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_istft(const float* in, long long ib, long long is, long long if_, long long it, const float* scale,
                          float* y, int B, int S, int Ts, int T, int n_fft, int hop, void* stream) {
    if (!in || !y) return NBSS_ERR_NULL;
    if (!pow2(n_fft) || n_fft > 1024 || hop < 1 || n_fft % hop || n_fft / hop > 8 || B < 1 || S < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int F = n_fft / 2 + 1, R = n_fft / hop, NFR = kFT + R - 1;
    const size_t smem = (size_t)(2 * n_fft + 2 * NFR * F + NFR * n_fft) * 4;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    IstftArgs a{in, ib, is, if_, it, scale, y, B, S, Ts, T, F, n_fft, hop};
    cudaError_t e = cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(y, 0, (size_t)B * S * Ts * 4, (cudaStream_t)stream);  // samples no frame covers stay 0
    if (e != cudaSuccess) return (int)e;
    const int nseg = T + R - 1;
    if (pow4(n_fft)) {
        const int NB = (NFR + 1) / 2;
        const size_t sm_f = (size_t)(n_fft + NFR * F + 2 * NB * n_fft) * 8 + (size_t)NFR * n_fft * 4;
        if (sm_f <= 227 * 1024) {
            cudaError_t ef = cudaFuncSetAttribute(istft_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f);
            if (ef != cudaSuccess) return (int)ef;
            istft_fft_kernel<<<B * S * ((nseg + kFT - 1) / kFT), 256, sm_f, (cudaStream_t)stream>>>(a);
            NBSS_LAUNCH_CHECK();
            return NBSS_OK;
        }
    }
    istft_kernel<<<B * S * ((nseg + kFT - 1) / kFT), 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_istft_bwd(const float* dy, const float* scale, float* din, long long ib, long long is, long long if_,
                              long long it, int B, int S, int Ts, int T, int n_fft, int hop, void* stream) {
    if (!dy || !din) return NBSS_ERR_NULL;
    if (!pow2(n_fft) || n_fft > 1024 || hop < 1 || n_fft % hop || B < 1 || S < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int F = n_fft / 2 + 1;
    const size_t smem = (size_t)(2 * n_fft + kFT * n_fft) * 4;
    IstftBwdArgs a{dy, scale, din, ib, is, if_, it, B, S, Ts, T, F, n_fft, hop};
    if (pow4(n_fft)) {
        const size_t sm_f = (size_t)(n_fft + kFT * n_fft) * 8;
        cudaError_t ef = cudaFuncSetAttribute(istft_bwd_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_f);
        if (ef != cudaSuccess) return (int)ef;
        istft_bwd_fft_kernel<<<B * S * ((T + kFT - 1) / kFT), 256, sm_f, (cudaStream_t)stream>>>(a);
        NBSS_LAUNCH_CHECK();
        return NBSS_OK;
    }
    cudaError_t e = cudaFuncSetAttribute(istft_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    istft_bwd_kernel<<<B * S * ((T + kFT - 1) / kFT), 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_encoder_fwd(const float* x, float* y, int nslab, int T, int cin, const float* W, const float* bias,
                                void* stream) {
    if (!x || !y || !W || !bias) return NBSS_ERR_NULL;
    if (nslab < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int grid = nslab * ((T + 63) / 64);
    cudaStream_t st = (cudaStream_t)stream;
    switch (cin) {
        case 2: encoder_fwd_kernel<2><<<grid, 192, 0, st>>>(x, y, nslab, T, W, bias); break;
        case 4: encoder_fwd_kernel<4><<<grid, 192, 0, st>>>(x, y, nslab, T, W, bias); break;
        case 8: encoder_fwd_kernel<8><<<grid, 192, 0, st>>>(x, y, nslab, T, W, bias); break;
        case 12: encoder_fwd_kernel<12><<<grid, 192, 0, st>>>(x, y, nslab, T, W, bias); break;
        case 16: encoder_fwd_kernel<16><<<grid, 192, 0, st>>>(x, y, nslab, T, W, bias); break;
        default: return NBSS_ERR_UNSUPPORTED;
    }
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_encoder_wgrad(const float* x, const float* dy, int nslab, int T, int cin, float* dW, float* dbias,
                                  void* stream) {
    if (!x || !dy || !dW || !dbias) return NBSS_ERR_NULL;
    if (nslab < 1 || T < 1) return NBSS_ERR_SHAPE;
    const int tiles = nslab * ((T + 63) / 64), cap = 4 * io_num_sms(), grid = tiles < cap ? tiles : cap;
    cudaStream_t st = (cudaStream_t)stream;
    switch (cin) {
        case 2: encoder_wgrad_kernel<2><<<grid, 192, 0, st>>>(x, dy, nslab, T, dW, dbias); break;
        case 4: encoder_wgrad_kernel<4><<<grid, 192, 0, st>>>(x, dy, nslab, T, dW, dbias); break;
        case 8: encoder_wgrad_kernel<8><<<grid, 192, 0, st>>>(x, dy, nslab, T, dW, dbias); break;
        case 12: encoder_wgrad_kernel<12><<<grid, 192, 0, st>>>(x, dy, nslab, T, dW, dbias); break;
        case 16: encoder_wgrad_kernel<16><<<grid, 192, 0, st>>>(x, dy, nslab, T, dW, dbias); break;
        default: return NBSS_ERR_UNSUPPORTED;
    }
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_decoder_fwd(const float* x, float* y, long long n, int cout, const float* W, const float* bias,
                                void* stream) {
    if (!x || !y || !W || !bias) return NBSS_ERR_NULL;
    if (n < 1) return NBSS_ERR_SHAPE;
    const int grid = 8 * io_num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    switch (cout) {
        case 2: decoder_fwd_kernel<2><<<grid, 256, 0, st>>>(x, y, (size_t)n, W, bias); break;
        case 4: decoder_fwd_kernel<4><<<grid, 256, 0, st>>>(x, y, (size_t)n, W, bias); break;
        case 6: decoder_fwd_kernel<6><<<grid, 256, 0, st>>>(x, y, (size_t)n, W, bias); break;
        case 8: decoder_fwd_kernel<8><<<grid, 256, 0, st>>>(x, y, (size_t)n, W, bias); break;
        default: return NBSS_ERR_UNSUPPORTED;
    }
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_decoder_bwd(const float* x, const float* dy, float* dx, long long n, int cout, const float* W, float* dW,
                                float* dbias, void* stream) {
    if (!x || !dy || !dx || !W || !dW || !dbias) return NBSS_ERR_NULL;
    if (n < 1) return NBSS_ERR_SHAPE;
    const int grid = 2 * io_num_sms();
    cudaStream_t st = (cudaStream_t)stream;
    switch (cout) {
        case 2: decoder_bwd_kernel<2><<<grid, 256, 0, st>>>(x, dy, dx, (size_t)n, W, dW, dbias); break;
        case 4: decoder_bwd_kernel<4><<<grid, 256, 0, st>>>(x, dy, dx, (size_t)n, W, dW, dbias); break;
        case 6: decoder_bwd_kernel<6><<<grid, 256, 0, st>>>(x, dy, dx, (size_t)n, W, dW, dbias); break;
        case 8: decoder_bwd_kernel<8><<<grid, 256, 0, st>>>(x, dy, dx, (size_t)n, W, dW, dbias); break;
        default: return NBSS_ERR_UNSUPPORTED;
    }
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_norm_freq_online(float* X, int B, int C, long long FT, int ref_channel, float eps, float* xrmm, float* xr,
                                     void* stream) {
    if (!X || !xrmm || !xr) return NBSS_ERR_NULL;
    if (B < 1 || C < 1 || FT < 1 || ref_channel < 0 || ref_channel >= C) return NBSS_ERR_SHAPE;
    const size_t n = (size_t)B * FT;
    norm_freq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((float2*)X, B, C, (size_t)FT, ref_channel, eps, xrmm, (float2*)xr);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_inorm(const float* X, float* Y, int B, int S, long long FT, const float* xrmm, void* stream) {
    if (!X || !Y || !xrmm) return NBSS_ERR_NULL;
    if (B < 1 || S < 1 || FT < 1) return NBSS_ERR_SHAPE;
    const size_t n = (size_t)B * S * FT;
    inorm_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float2*)X, (float2*)Y, B, S, (size_t)FT, xrmm);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
