// pack.cu — builds the per-layer UMMA weight images (layout.cuh) from the reference's fp32 parameters.
// Parameters follow the reference state_dict: tconvffn.{1,3,5,8,10}.weight, mhsa.in_proj_weight,
// mhsa.out_proj.weight.  Runs once per optimizer step (weights change), 39168 threads per layer.
#include "common.cuh"
#include "layout.cuh"
#include "umma.cuh"

namespace nbss {

struct PackArgs {
    const float* w1;   // [192,96]   tconvffn.1.weight[:, :, 0]
    const float* wc[3];  // [192,24,3] tconvffn.{3,5,8}.weight
    const float* w2;   // [96,192]   tconvffn.10.weight[:, :, 0]
    const float* w_in;  // [288,96]
    const float* w_out;  // [96,96]
    unsigned char* img;
    int fwd_fmt, bwd_fmt;
};

__device__ __forceinline__ float conv_fwd_elem(const float* wc, int p, int tap, int n, int kk) {
    return (n / kGC == kk / kGC) ? wc[(48 * p + n) * 72 + (kk % kGC) * 3 + tap] : 0.f;
}
__device__ __forceinline__ float conv_bwd_elem(const float* wc, int p, int tap, int n, int kk) {
    // dgrad: d_in[t, ci] = sum_tap sum_co d_out[t-(tap-1), co] * W[co, ci, tap];  N index = ci, K index = co
    return (n / kGC == kk / kGC) ? wc[(48 * p + kk) * 72 + (n % kGC) * 3 + tap] : 0.f;
}

__global__ void pack_layer_kernel(PackArgs a) {
    const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t byte = chunk * 16;
    if (byte >= IMG_LAYER_BYTES) return;
    float v[8];
    int fmt = a.fwd_fmt;
    auto plain = [&](uint32_t base, int rows, auto f) {
        uint32_t i = (byte - base) / 16;
        int c = i / rows, n = i % rows;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = f(n, c * 8 + j);
    };
    auto conv = [&](uint32_t base, const float* wc, bool bwd) {
        uint32_t i = (byte - base) / 16;
        int tile = i / (6 * 48), r = i % (6 * 48);
        int c = r / 48, n = r % 48, p = tile / 3, tap = tile % 3;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            v[j] = bwd ? conv_bwd_elem(wc, p, tap, n, c * 8 + j) : conv_fwd_elem(wc, p, tap, n, c * 8 + j);
    };
    if (byte < IMG_WC1) plain(IMG_W1, 192, [&](int n, int k) { return a.w1[n * 96 + k]; });
    else if (byte < IMG_WC2) conv(IMG_WC1, a.wc[0], false);
    else if (byte < IMG_WC3) conv(IMG_WC2, a.wc[1], false);
    else if (byte < IMG_W2) conv(IMG_WC3, a.wc[2], false);
    else if (byte < IMG_W2T) plain(IMG_W2, 96, [&](int n, int k) { return a.w2[n * 192 + k]; });
    else if (byte < IMG_WC3T) { fmt = a.bwd_fmt; plain(IMG_W2T, 192, [&](int n, int k) { return a.w2[k * 192 + n]; }); }
    else if (byte < IMG_WC2T) { fmt = a.bwd_fmt; conv(IMG_WC3T, a.wc[2], true); }
    else if (byte < IMG_WC1T) { fmt = a.bwd_fmt; conv(IMG_WC2T, a.wc[1], true); }
    else if (byte < IMG_W1T) { fmt = a.bwd_fmt; conv(IMG_WC1T, a.wc[0], true); }
    else if (byte < IMG_WKV) { fmt = a.bwd_fmt; plain(IMG_W1T, 96, [&](int n, int k) { return a.w1[k * 96 + n]; }); }
    else if (byte < IMG_WQ) plain(IMG_WKV, 192, [&](int n, int k) { return a.w_in[(96 + n) * 96 + k]; });
    else if (byte < IMG_WO) plain(IMG_WQ, 96, [&](int n, int k) { return a.w_in[n * 96 + k]; });
    else if (byte < IMG_WOT) plain(IMG_WO, 96, [&](int n, int k) { return a.w_out[n * 96 + k]; });
    else if (byte < IMG_WINT) { fmt = a.bwd_fmt; plain(IMG_WOT, 96, [&](int n, int k) { return a.w_out[k * 96 + n]; }); }
    else { fmt = a.bwd_fmt; plain(IMG_WINT, 96, [&](int n, int k) { return a.w_in[k * 96 + n]; }); }
    uint4 q;
    if (fmt == FMT_F16) q = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
    else q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(a.img + byte) = q;
}

}  // namespace nbss

extern "C" unsigned int nbss_layer_image_bytes() { return nbss::IMG_LAYER_BYTES; }

// ABI version of this library: 100 * round + revision (nbss_version()).
extern "C" int nbss_version() { return 200; }

// Bytes of caller-owned scratch one SpatialNetLayer needs around its kernels at [B,F,T] (SURVEY.md §8b:
// nbss_workspace_bytes()): training = activations saved by the forward for the backward kernels plus the transient
// gradient-operand tensors of the backward; inference = 0 (every sub-block updates the stream in place).
extern "C" long long nbss_workspace_bytes(int B, int F, int T, int training) {
    if (B < 1 || F < 1 || T < 1) return -1;
    if (!training) return 0;
    const long long n = (long long)B * F * T, nslab = (long long)B * F;
    const long long stream = n * 96 * 4;
    long long saved = 5 * stream;                         // the five sub-block inputs (fp32 stream)
    saved += n * (288 + 96) * 2 + nslab * 4 * T * 4;      // fp16 q|k|v, O, log2-sum-exp
    saved += 4 * n * 192 * 2 + nslab * 16 * 4;            // fp16 T-ConvFFN pre-activations, GroupNorm statistics
    saved += 2 * n * 2 * 4;                               // LayerNorm statistics of the two narrow-band sub-blocks
    saved += 2 * (long long)B * T * 8 * F * 4;            // squeeze / full-band outputs [B,T,8,F]
    const long long transient = 4 * n * 192 * 2 + n * 288 * 2 + 2 * stream;  // gradient operands + two stream gradients
    return saved + transient;
}

extern "C" int nbss_pack_layer_weights(const float* w1, const float* wc1, const float* wc2, const float* wc3,
                                       const float* w2, const float* w_in, const float* w_out, void* img, int fwd_fmt,
                                       int bwd_fmt, void* stream) {
    using namespace nbss;
    if (!w1 || !wc1 || !wc2 || !wc3 || !w2 || !w_in || !w_out || !img) return NBSS_ERR_NULL;
    PackArgs a{w1, {wc1, wc2, wc3}, w2, w_in, w_out, (unsigned char*)img, fwd_fmt, bwd_fmt};
    const int n = IMG_LAYER_BYTES / 16;
    pack_layer_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
