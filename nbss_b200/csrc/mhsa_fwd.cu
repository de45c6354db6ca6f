// mhsa_fwd.cu — narrow-band multi-head self-attention sub-block, forward, one CTA per (b,f) slab, tcgen05 + TMEM.
//
// Replaces SpatialNetLayer._tsa + residual (models/arch/SpatialNet.py:88-89,93-100) i.e. LN -> nn.MultiheadAttention
// (packed in-proj, q*dh^-0.5, softmax over T, out-proj; no mask, no dropout) for T <= 256:
//   P0 stage x -> LN -> fp16 A0 [256x96]
//   P1 K|V = A0 Wkv^T  (N=192)         E1: +bias -> K tile (per head padded 24->32, zero chunk) and V tile
//   P2 Q   = A0 Wq^T   (N=96, stays in TMEM)
//   per head h, query tile m:  Qs = (Q_h + b) * dh^-0.5 * log2(e)  -> smem [128x32]
//        S = Qs K_h^T (N=256 keys, TMEM)   softmax over keys in registers (thread = query row, 2 threads per row)
//        O_h = P V_h (P fp16 [128 x 128 keys] staged twice, V read MN-major)  -> O tile (normalised by the row sum)
//   P3 y = x + O Wo^T + bo
// HBM traffic: x in, y out (+ fp16 q|k|v, O and the log2-sum-exp when save != 0, for the backward kernels).
#include <cstdlib>

#include "slab.cuh"

namespace nbss {

struct MhsaFwdArgs {
    const float* x;
    float* y;
    int nslab, T;
    const float *ln_w, *ln_b, *b_in, *b_out;
    const unsigned char* img;
    unsigned char* save_qkv;  // fp16 slab-tile [nslab][36][T][8]: (scaled q | k | v) or null
    unsigned char* save_o;    // fp16 slab-tile [nslab][12][T][8] or null
    float* save_lse;          // [nslab, 4, T] log2-domain logsumexp or null
    float* ln_stats;          // [nslab*T, 2] (mean, rstd) of the LayerNorm or null
    int* err;
};

constexpr uint32_t kCSP = 129 * 16;  // chunk stride of the 128-row P / Qs tiles
constexpr uint32_t MH_AO = 0;
constexpr uint32_t MH_K = 13 * kCS;             // 55120
constexpr uint32_t MH_V = MH_K + 16 * kCS;      // 122960
constexpr uint32_t MH_W = MH_V + 13 * kCS;      // 178080
constexpr uint32_t MH_W_BYTES = 20 * kCSP;      // 41280: P (16 chunks) + Qs (4 chunks), or one weight image
constexpr uint32_t MH_CST = MH_W + MH_W_BYTES;  // 219360
constexpr uint32_t MH_XCH = MH_CST + 576 * 4;
constexpr uint32_t MH_BAR = MH_XCH + 4096;
constexpr int kMhThreads = 512;  // 16 warps
constexpr uint32_t MH_SMEM = MH_BAR + 64;
static_assert(IMG_W1_BYTES <= MH_W_BYTES, "weight image must fit the W region");

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// PTMEM: the softmax probabilities go back into tensor memory (over the score columns they came from) and feed the PV MMA
// as its A operand from TMEM (umma.cuh: umma_f16_ts): one MMA round per (head, query tile) instead of two, no P tile.
template <int FMT, bool PTMEM>
__global__ void __launch_bounds__(kMhThreads, 1) mhsa_fwd_kernel(MhsaFwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ao = smem + MH_AO;
    unsigned char* kt = smem + MH_K;
    unsigned char* vt = smem + MH_V;
    unsigned char* wr = smem + MH_W;
    unsigned char* pt = wr;               // P tile [128 x 128 keys]
    unsigned char* qs = wr + 16 * kCSP;   // Qs tile [128 x 32]
    float* cst = reinterpret_cast<float*>(smem + MH_CST);
    float *s_lng = cst, *s_lnb = cst + 96, *s_bin = cst + 192, *s_bout = cst + 480;
    float* xmax = reinterpret_cast<float*>(smem + MH_XCH);
    float* xsum = xmax + 512;  // [4 key quarters][128 rows] each
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + MH_BAR);
    uint64_t* bar_w = bar_mma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int T = a.T;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 96; i += kMhThreads) { s_lng[i] = a.ln_w[i]; s_lnb[i] = a.ln_b[i]; s_bout[i] = a.b_out[i]; }
    for (int i = tid; i < 288; i += kMhThreads) s_bin[i] = a.b_in[i];
    // zero everything that is read as padding: AO, K (pad chunks), V (13th chunk)
    for (int i = tid; i < (int)(MH_W / 16); i += kMhThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // 16 warps.  'Both tiles' epilogues: M-tile m, lane quarter q, column half hf.  Softmax: key quarter kq.
    const int m = (warp >> 2) & 1, q = warp & 3, hf = warp >> 3, kq = warp >> 2;
    const int rt = 32 * q + lane;       // row within an M-tile (TMEM lane)
    const int t = 128 * m + rt;         // frame handled in "both tiles" epilogues
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t aoa = smem_u32(ao), kta = smem_u32(kt), vta = smem_u32(vt), wra = smem_u32(wr), pta = smem_u32(pt),
                   qsa = smem_u32(qs);
    const uint32_t id192 = make_idesc(FMT, 128, 192, 0, 0), id96 = make_idesc(FMT, 128, 96, 0, 0),
                   id256 = make_idesc(FMT, 128, 256, 0, 0), idpv = make_idesc(FMT, 128, 32, 0, 1);
    const float qscale = rsqrtf((float)kDH) * 1.4426950408889634f;
    uint32_t ph_mma = 0, ph_w = 0;

    auto wait_mma = [&]() {
        __syncwarp();
        mbar_wait(bar_mma, ph_mma, a.err);
        ph_mma ^= 1;
        tc_fence_after();
    };
    auto end_epilogue = [&]() {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };

    int it_ = 0;
    for (int slab = blockIdx.x; slab < a.nslab; slab += gridDim.x, ++it_) {
        const float* xs = a.x + (size_t)slab * T * kH;
        NBSS_TICK(0, 0, it_);
        if (tid == 0) load_image(wr, a.img + IMG_WKV, IMG_W1_BYTES, bar_w);
        if (tid >= 32 && tid < 38 && slab + (int)gridDim.x < a.nslab)  // next slab's input rows -> L2 (after the weight copy)
            l2_prefetch_slab(a.x + (size_t)(slab + gridDim.x) * T * kH, T, tid - 32);
        stage_rows96<FMT, true>(xs, T, ao, 0, s_lng, s_lnb, warp, lane, a.ln_stats ? a.ln_stats + (size_t)slab * T * 2 : nullptr, kMhThreads / 32);
        end_epilogue();
        NBSS_TICK(0, 1, it_);
        // ---- P1: K|V
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + mm * 192, aoa + 128 * mm * 16, kCS, wra, 192 * 16, 6, id192, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(0, 2, it_);
        if (tid == 0) load_image(wr, a.img + IMG_WQ, IMG_WQ_BYTES, bar_w);
        {
            const bool valid = t < T;
            const uint32_t tacc = tmem + lane_off + m * 192;
            // K: cols 0..95 -> per-head padded chunks 4h..4h+2   (channel half 0 does K, half 1 does V)
#pragma unroll 1
            for (int h = 0; h < (hf == 0 ? kNH : 0); ++h) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t r[8];
                    tmem_ld8(tacc + kDH * h + 8 * k, r);
                    tmem_ld_wait();
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = valid ? __uint_as_float(r[j]) + s_bin[96 + kDH * h + 8 * k + j] : 0.f;
                    uint4 p = pack8<FMT>(v);
                    *reinterpret_cast<uint4*>(kt + (4 * h + k) * kCS + t * 16) = p;
                    if (a.save_qkv && valid) *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, 12 + 3 * h + k, t)) = pack8<FMT_F16>(v);
                }
            }
            // V: cols 96..191 -> compact chunks 0..11
#pragma unroll 1
            for (int c = 0; c < (hf == 1 ? 12 : 0); ++c) {
                uint32_t r[8];
                tmem_ld8(tacc + 96 + 8 * c, r);
                tmem_ld_wait();
                float v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = valid ? __uint_as_float(r[j]) + s_bin[192 + 8 * c + j] : 0.f;
                *reinterpret_cast<uint4*>(vt + c * kCS + t * 16) = pack8<FMT>(v);
                if (a.save_qkv && valid) *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, 24 + c, t)) = pack8<FMT_F16>(v);
            }
        }
        end_epilogue();
        NBSS_TICK(0, 3, it_);
        // ---- P2: Q (stays in TMEM cols 0..191)
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + mm * 96, aoa + 128 * mm * 16, kCS, wra, 96 * 16, 6, id96, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(0, 4, it_);
        // ---- heads x query tiles
#pragma unroll 1
        for (int hm = 0; hm < 2 * kNH; ++hm) {
            const int h = hm >> 1, mq = hm & 1;
            // EQ: warps 0..3 stage the scaled queries of tile mq
            if (warp < 4) {
                const int tq = 128 * mq + rt;
                const bool valid = tq < T;
                const uint32_t tq_acc = tmem + lane_off + mq * 96 + kDH * h;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t r[8];
                    tmem_ld8(tq_acc + 8 * k, r);
                    tmem_ld_wait();
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = valid ? (__uint_as_float(r[j]) + s_bin[kDH * h + 8 * k + j]) * qscale : 0.f;
                    *reinterpret_cast<uint4*>(qs + k * kCSP + rt * 16) = pack8<FMT>(v);
                    if (a.save_qkv && valid)
                        *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, 3 * h + k, tq)) = pack8<FMT_F16>(v);
                }
                *reinterpret_cast<uint4*>(qs + 3 * kCSP + rt * 16) = make_uint4(0, 0, 0, 0);
            }
            end_epilogue();
            NBSS_TICK(0, 8 + 5 * hm, it_);
            // S = Qs K_h^T
            if (warp == 0) {
                tc_fence_after();
                const bool leader = elect_one();
                mma_kk(tmem + 192, qsa, kCSP, kta + 4 * h * kCS, kCS, 2, id256, 0, leader);
                if (leader) umma_commit(bar_mma);
            }
            wait_mma();
            NBSS_TICK(0, 9 + 5 * hm, it_);
            // softmax: thread = (query row rt, key quarter kq): 64 of the 256 score columns
            const uint32_t ts = tmem + lane_off + 192 + 64 * kq;
            float mx = -INFINITY;
            {
                // both 32-column loads in flight before the single wait; four independent max chains
                uint32_t r0[32], r1[32];
                tmem_ld32(ts, r0);
                tmem_ld32(ts + 32, r1);
                tmem_ld_wait();
                float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    m4[j & 1] = fmaxf(m4[j & 1], (64 * kq + j < T) ? __uint_as_float(r0[j]) : -INFINITY);
                    m4[2 + (j & 1)] = fmaxf(m4[2 + (j & 1)], (64 * kq + 32 + j < T) ? __uint_as_float(r1[j]) : -INFINITY);
                }
                mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            }
            xmax[kq * 128 + rt] = mx;
            __syncthreads();
            const float rowmax = fmaxf(fmaxf(xmax[rt], xmax[128 + rt]), fmaxf(xmax[256 + rt], xmax[384 + rt]));
            uint32_t pk[32];
            float sum = 0.f;
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(ts + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 2) {
                    const int key = 64 * kq + c0 + j;
                    // masked keys get exponent -inf (ex2 -> 0): a select on the argument, no branch around the MUFU
                    float p0 = ex2(key < T ? __uint_as_float(r[j]) - rowmax : -INFINITY);
                    float p1 = ex2(key + 1 < T ? __uint_as_float(r[j + 1]) - rowmax : -INFINITY);
                    sum += p0 + p1;
                    pk[(c0 + j) >> 1] = pack16<FMT>(p0, p1);
                }
            }
            xsum[kq * 128 + rt] = sum;
            if constexpr (PTMEM) {
                // O_h = P V_h with P in tensor memory: key k of row rt -> column 192 + k/2 (two 16-bit values per column)
                __syncthreads();  // every thread has read its score columns: the P columns alias them
                const uint32_t tp = tmem + lane_off + 192 + 32 * kq;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t v8[8] = {pk[8 * c], pk[8 * c + 1], pk[8 * c + 2], pk[8 * c + 3], pk[8 * c + 4], pk[8 * c + 5], pk[8 * c + 6], pk[8 * c + 7]};
                    tmem_st8(tp + 8 * c, v8);
                }
                tmem_st_wait();
                end_epilogue();
                NBSS_TICK(0, 10 + 5 * hm, it_);
                if (warp == 0) {
                    tc_fence_after();
                    const bool leader = elect_one();
                    for (int ks = 0; ks < 16; ++ks)
                        if (leader) umma_f16_ts(tmem + 448, tmem + 192 + 8 * ks, sdesc_mnmajor(vta + 3 * h * kCS + 16 * ks * 16, kCS), idpv, ks ? 1u : 0u);
                    if (leader) umma_commit(bar_mma);
                }
                wait_mma();
                NBSS_TICK(0, 11 + 5 * hm, it_);
            } else {
            // O_h = P V_h, one key half at a time (the P tile holds 128 keys = two key quarters)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if ((kq >> 1) == half) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<uint4*>(pt + (8 * (kq & 1) + c) * kCSP + rt * 16) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
                end_epilogue();
                if (warp == 0) {
                    tc_fence_after();
                    const bool leader = elect_one();
                    for (int ks = 0; ks < 8; ++ks)
                        if (leader) umma_f16(tmem + 448, sdesc_kmajor(pta + 2 * ks * kCSP, kCSP),
                                 sdesc_mnmajor(vta + 3 * h * kCS + (128 * half + 16 * ks) * 16, kCS), idpv, (half | ks) ? 1u : 0u);
                    if (leader) umma_commit(bar_mma);
                }
                wait_mma();
            }
            }
            // EO: normalise and place O_h into the O tile (aliases A0, dead after P2)
            if (warp < 4) {
                const int tq = 128 * mq + rt;
                const float l = xsum[rt] + xsum[128 + rt] + xsum[256 + rt] + xsum[384 + rt];
                const float inv = 1.f / l;
                if (a.save_lse && tq < T) a.save_lse[((size_t)slab * kNH + h) * T + tq] = rowmax + log2f(l);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    uint32_t r[8];
                    tmem_ld8(tmem + lane_off + 448 + 8 * k, r);
                    tmem_ld_wait();
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) * inv;
                    *reinterpret_cast<uint4*>(ao + (3 * h + k) * kCS + tq * 16) = pack8<FMT>(v);
                    if (a.save_o && tq < T)
                        *reinterpret_cast<uint4*>(a.save_o + tile_off(slab, 12, T, 3 * h + k, tq)) = pack8<FMT_F16>(v);
                }
            }
            tc_fence_before();
            __syncthreads();
            NBSS_TICK(0, 12 + 5 * hm, it_);
        }
        // ---- P3: out-proj + residual
        if (tid == 0) load_image(wr, a.img + IMG_WO, IMG_WQ_BYTES, bar_w);
        end_epilogue();
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + 192 + mm * 96, aoa + 128 * mm * 16, kCS, wra, 96 * 16, 6, id96, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(0, 5, it_);
        {
            // thread = (frame, channel half): D + b_out -> fp32, staged into the dead K|V tiles at the frame's row slot (24
            // four-float chunks = the 12 K data chunks + V chunks 0..11; the zero pad chunks of K stay untouched:
            // slab.cuh skip4_chunk); then eight lanes per frame add the residual with coalesced traffic
            const uint32_t tacc = tmem + lane_off + 192 + m * 96;
#pragma unroll 1
            for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tacc + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                    float4 o;
                    o.x = __uint_as_float(r[4 * j4 + 0]) + s_bout[c0 + 4 * j4 + 0];
                    o.y = __uint_as_float(r[4 * j4 + 1]) + s_bout[c0 + 4 * j4 + 1];
                    o.z = __uint_as_float(r[4 * j4 + 2]) + s_bout[c0 + 4 * j4 + 2];
                    o.w = __uint_as_float(r[4 * j4 + 3]) + s_bout[c0 + 4 * j4 + 3];
                    *reinterpret_cast<float4*>(kt + (size_t)skip4_chunk(c0 / 4 + j4) * kCS + t * 16) = o;
                }
            }
            tc_fence_before();
            __syncthreads();
            NBSS_TICK(0, 6, it_);
            add_rows<true>(kt, kCS, 0, xs, a.y + (size_t)slab * T * kH, T, warp, lane, kMhThreads / 32);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 7, it_);
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace nbss

NBSS_PHASE_READER(nbss_debug_phases_mhsa_fwd)

extern "C" int nbss_mhsa_fwd(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b,
                             const float* b_in, const float* b_out, const void* layer_img, void* save_qkv, void* save_o,
                             float* save_lse, float* ln_stats, int fmt, int* err, void* stream) {
    using namespace nbss;
    if (!x || !y || !layer_img || !ln_w || !ln_b || !b_in || !b_out) return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    MhsaFwdArgs a{x, y, nslab, T, ln_w, ln_b, b_in, b_out, (const unsigned char*)layer_img, (unsigned char*)save_qkv,
                  (unsigned char*)save_o, save_lse, ln_stats, err};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = nslab < sms ? nslab : sms;
    static int ptmem = -1;  // NBSS_MHSA_PTMEM=0 keeps the softmax probabilities in a shared-memory tile (the older path)
    if (ptmem < 0) {
        const char* e = getenv("NBSS_MHSA_PTMEM");
        ptmem = (e && e[0] == '0') ? 0 : 1;
    }
    void (*kern)(MhsaFwdArgs) = ptmem ? ((fmt == FMT_F16) ? mhsa_fwd_kernel<FMT_F16, true> : mhsa_fwd_kernel<FMT_BF16, true>)
                                      : ((fmt == FMT_F16) ? mhsa_fwd_kernel<FMT_F16, false> : mhsa_fwd_kernel<FMT_BF16, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kMhThreads, MH_SMEM, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
