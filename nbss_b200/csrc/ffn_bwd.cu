// ffn_bwd.cu — backward (data gradient) of the narrow-band T-ConvFFN sub-block, one CTA per (b,f) slab, tcgen05.
//
// Mirrors ffn_fwd.cu in reverse with transposed weight images (layout.cuh: W2T, WC{3,2,1}T, W1T).  Inputs: block input
// x, upstream gradient dy, the fp16 pre-activations a1,c1,c2,c3 saved by ffn_fwd, GroupNorm / LayerNorm statistics.
//   B0 stage dy -> 16-bit tile G                      E1 g(c3) = D * SiLU'(c3)              -> G
//   B1 d s4 = dy  W2        (K=96,  N=192)            E2 GroupNorm backward (block sums)    -> G
//   B2 d s3 = conv3^T(g c3) (row-shifted views)       E3 g(c1) = D * SiLU'(c1)              -> G
//   B3 d s2 = conv2^T(g c2)                           E4 g(a1) = D * SiLU'(a1)              -> G
//   B4 d s1 = conv1^T(g c1)                           E5 LayerNorm backward, dx = dy + ...
//   B5 d ln = g(a1) W1      (K=192, N=96)
// Every E-phase also streams out, for the weight-gradient kernels (wgrad.cu), the gradient operand g(.) as a 16-bit [n,192]
// tensor; the activation operand s(.) = SiLU(.) is recomputed by wgrad from the forward's saved pre-activation (it used to be
// written here too: 1.6 GB of the kernel's 6.3 GB of HBM traffic per launch).  Column sums needed for the affine
// parameters of LN / GN are formed with warp transposing reductions and accumulated in shared memory.
#include "slab.cuh"

namespace nbss {

struct FfnBwdArgs {
    const float* x;
    const float* dy;
    float* dx;
    int nslab, T;
    const float *ln_w, *gn_w, *gn_b;
    const float* ln_stats;  // [n,2]
    const float* gn_stats;  // [nslab,8,2]
    const unsigned char* img;
    const unsigned char *a1, *c1, *c2, *c3;   // fp16 [n,192]
    unsigned char *g_a1, *g_c1, *g_c2, *g_c3;  // 16-bit (FMT) [n,192] gradients wrt the pre-activations
    float *d_lnw, *d_lnb, *d_gnw, *d_gnb;      // accumulated with atomics
    int* err;
};

constexpr uint32_t FB_HBUF = 0;
constexpr uint32_t FB_WS0 = 24 * kCS;
constexpr uint32_t FB_WS1 = FB_WS0 + IMG_WC_BYTES;
constexpr uint32_t FB_CST = FB_WS1 + IMG_WC_BYTES;  // ln_w 96, gn_w 192, gn_b 192
constexpr uint32_t FB_ACC = FB_CST + 480 * 4;       // column-sum accumulators: d_gnw 192, d_gnb 192, d_lnw 96, d_lnb 96
constexpr uint32_t FB_RED = FB_ACC + 576 * 4;       // [16 warps][8 groups][2] + group totals [8][2]
constexpr uint32_t FB_XCH = FB_RED + (256 + 16) * 4;  // LayerNorm partial sums of the two channel halves [512][2]
constexpr uint32_t FB_BAR = FB_XCH + 512 * 8;
constexpr int kFfnBwdThreads = 512;  // warp w -> M-tile (w>>2)&1, TMEM lane quarter w&3, channel half w>>3
constexpr uint32_t FB_SMEM = FB_BAR + 64;

// all 16-bit tensors here use the slab-tile layout [slab][24 chunks][T][8] (slab.cuh)
__device__ __forceinline__ void load_h16x8(const unsigned char* base, int slab, int T, int t, int c, float* v) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(base + tile_off(slab, 24, T, c / 8, t)));
    unpack_f16x2(q.x, v[0], v[1]);
    unpack_f16x2(q.y, v[2], v[3]);
    unpack_f16x2(q.z, v[4], v[5]);
    unpack_f16x2(q.w, v[6], v[7]);
}
__device__ __forceinline__ void load_h16x32(const unsigned char* base, int slab, int T, int t, int c0, float* v) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) load_h16x8(base, slab, T, t, c0 + 8 * cc, v + 8 * cc);
}
template <int FMT>
__device__ __forceinline__ void store16x32(unsigned char* base, int slab, int T, int t, int c0, const float* v) {
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) *reinterpret_cast<uint4*>(base + tile_off(slab, 24, T, c0 / 8 + cc, t)) = pack8<FMT>(v + 8 * cc);
}

template <int FMT>
__global__ void __launch_bounds__(kFfnBwdThreads, 1) ffn_bwd_kernel(FfnBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* hbuf = smem + FB_HBUF;
    unsigned char* ws0 = smem + FB_WS0;
    unsigned char* ws1 = smem + FB_WS1;
    float* cst = reinterpret_cast<float*>(smem + FB_CST);
    float *s_lng = cst, *s_gng = cst + 96, *s_gnb = cst + 288;
    float* acc = reinterpret_cast<float*>(smem + FB_ACC);  // [0,192) d_gnw, [192,384) d_gnb, [384,480) d_lnw, [480,576) d_lnb
    float* red = reinterpret_cast<float*>(smem + FB_RED);
    float* gtot = red + 256;  // [8][2] group totals S1, S2
    float2* xch = reinterpret_cast<float2*>(smem + FB_XCH);
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + FB_BAR);
    uint64_t* bar_mma1 = bar_mma + 1;  // one commit barrier per M-tile: tile 0's epilogue warps start while tile 1's MMAs run
    uint64_t* bar_w0 = bar_mma + 2;
    uint64_t* bar_w1 = bar_mma + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const int T = a.T;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_mma1, 1);
        mbar_init(bar_w0, 1);
        mbar_init(bar_w1, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 96; i += kFfnBwdThreads) s_lng[i] = a.ln_w[i];
    for (int i = tid; i < 192; i += kFfnBwdThreads) { s_gng[i] = a.gn_w[i]; s_gnb[i] = a.gn_b[i]; }
    for (int i = tid; i < 576; i += kFfnBwdThreads) acc[i] = 0.f;
    for (int i = tid; i < (int)(24 * kCS / 16); i += kFfnBwdThreads) reinterpret_cast<uint4*>(hbuf)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int m = (warp >> 2) & 1, q = warp & 3, hf = warp >> 3;  // two threads per frame: channel halves
    const int cb = 96 * hf;
    const int t = 128 * m + 32 * q + lane;
    const bool valid = t < T;
    const float vmask = valid ? 1.f : 0.f;
    const bool wfull = 128 * m + 32 * q + 31 < T;  // warp-uniform: every frame of this warp is valid, no masking needed
    const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16) + m * 192;
    unsigned char* hrow = hbuf + (t + 1) * 16;
    const uint32_t hb = smem_u32(hbuf), w0a = smem_u32(ws0), w1a = smem_u32(ws1);
    const uint32_t id192 = make_idesc(FMT, 128, 192, 0, 0), id48 = make_idesc(FMT, 128, 48, 0, 0),
                   id96 = make_idesc(FMT, 128, 96, 0, 0);
    uint32_t ph_mma = 0, ph_w0 = 0, ph_w1 = 0;
    const float inv_n = 1.f / (float)(kGC * T);

    auto wait_mma = [&]() {
        __syncwarp();
        mbar_wait(m ? bar_mma1 : bar_mma, ph_mma, a.err);
        // tile 1's conv MMAs read G row 128 (frame 127, the halo): the tile-0 warps that own frames 96..127 must not
        // overwrite it before tile 1 has finished too
        if (m == 0 && q == 3) mbar_wait(bar_mma1, ph_mma, a.err);
        ph_mma ^= 1;
        tc_fence_after();
    };
    // transposed conv: d_in[t] = sum_tap g[t - (tap-1)] Wt_tap  ->  A rows start at 128*mm + 2 - tap (frame t = row t+1)
    auto convT_phase = [&](uint32_t wsa, uint64_t* bar_w, uint32_t& ph_w) {
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) {
                for (int p = 0; p < kPairs; ++p)
                    for (int tap = 0; tap < 3; ++tap)
                        mma_kk(tmem + mm * 192 + p * 48, hb + 6 * p * kCS + (128 * mm + 2 - tap) * 16, kCS,
                               wsa + (p * 3 + tap) * 6 * 768, 768, 3, id48, tap > 0, leader);
                if (leader) umma_commit(mm ? bar_mma1 : bar_mma);
            }
        }
        ph_w ^= 1;
        wait_mma();
    };
    auto end_epilogue = [&]() {
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
    };
    // g = D * SiLU'(c), s = SiLU(c): the three plain activation epilogues
    // one 16-column block: g = D * SiLU'(c) -> G tile + global, s = SiLU(c) -> global
    auto silu_block = [&](const uint32_t (&r)[16], const uint4 (&cq)[2], int c0, unsigned char* gout, int slab) {
        float c[16], g[16];
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            unpack_f16x2(cq[cc].x, c[8 * cc + 0], c[8 * cc + 1]);
            unpack_f16x2(cq[cc].y, c[8 * cc + 2], c[8 * cc + 3]);
            unpack_f16x2(cq[cc].z, c[8 * cc + 4], c[8 * cc + 5]);
            unpack_f16x2(cq[cc].w, c[8 * cc + 6], c[8 * cc + 7]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float cv = c[j];
            const float sg = sigmoidf_(cv);
            g[j] = __uint_as_float(r[j]) * sg * fmaf(cv, 1.f - sg, 1.f);
        }
        if (!wfull) {  // frames >= T: zero gradient rows (they are the transposed conv's zero padding)
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] *= vmask;
        }
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const uint4 gp = pack8<FMT>(g + 8 * cc);
            if (valid) *reinterpret_cast<uint4*>(gout + tile_off(slab, 24, T, c0 / 8 + cc, t)) = gp;
            *reinterpret_cast<uint4*>(hrow + (c0 / 8 + cc) * kCS) = gp;
        }
    };
    // saved pre-activation chunks c0/8, c0/8+1 of this thread's frame (zeros for frames >= T)
    auto load_c = [&](const unsigned char* csave, int slab, int c0, uint4 (&cq)[2]) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc)
            cq[cc] = valid ? __ldg(reinterpret_cast<const uint4*>(csave + tile_off(slab, 24, T, c0 / 8 + cc, t))) : make_uint4(0, 0, 0, 0);
    };
    // the three plain activation epilogues: TMEM loads and the global loads of the saved pre-activations are both
    // software-pipelined over two register buffers (the global loads are issued one 16-column block ahead)
    auto silu_epilogue = [&](const unsigned char* csave, unsigned char* gout, int slab) {
        uint32_t ra[16], rb[16];
        uint4 ca[2], cb2[2];
        load_c(csave, slab, cb, ca);
        tmem_ld16(tacc + cb, ra);
        tmem_ld_wait();
#pragma unroll 1
        for (int c0 = cb; c0 < cb + 96; c0 += 32) {
            tmem_ld16(tacc + c0 + 16, rb);
            load_c(csave, slab, c0 + 16, cb2);
            silu_block(ra, ca, c0, gout, slab);
            tmem_ld_wait();
            if (c0 + 32 < cb + 96) {
                tmem_ld16(tacc + c0 + 32, ra);
                load_c(csave, slab, c0 + 32, ca);
            }
            silu_block(rb, cb2, c0 + 16, gout, slab);
            tmem_ld_wait();
        }
    };

    // L2 prefetch of one saved fp16 pre-activation tensor of a slab (slab-tile layout: 24 T 16 contiguous bytes), six threads of warp 1.
    // The activation epilogues keep only one 16-column block (32 B per thread, 16 KB per SM) of these loads in flight — the kernel
    // is at 128 registers — so out of HBM they run at ~11 B/clk; issued one phase ahead they are L2 hits.
    auto prefetch_saved = [&](const unsigned char* base, int slab) {
        if (tid >= 32 && tid < 38 && slab < a.nslab)
            l2_prefetch(base + tile_off(slab, 24, T, 0, 0) + (size_t)(tid - 32) * T * 64, (uint32_t)(T * 64));
    };
    stagger_start(112000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    for (int slab = blockIdx.x; slab < a.nslab; slab += gridDim.x, ++it_) {
        const size_t grow = (size_t)slab * T + t;
        const float* dys = a.dy + (size_t)slab * T * kH;
        NBSS_TICK(0, 0, it_);
        if (tid == 0) {
            load_image(ws0, a.img + IMG_W2T, IMG_W1_BYTES, bar_w0);
            load_image(ws1, a.img + IMG_WC3T, IMG_WC_BYTES, bar_w1);
        }
        // ---- B0: dy -> G (chunks 0..11)
        stage_rows96<FMT, false, 2>(dys, T, hbuf, 1, nullptr, nullptr, warp, lane, nullptr, kFfnBwdThreads / 32);
        end_epilogue();
        NBSS_TICK(0, 1, it_);
        // ---- B1: d s4 = dy W2
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w0, ph_w0, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) {
                mma_kk(tmem + mm * 192, hb + (128 * mm + 1) * 16, kCS, w0a, 192 * 16, 6, id192, 0, leader);
                if (leader) umma_commit(mm ? bar_mma1 : bar_mma);
            }
        }
        ph_w0 ^= 1;
        wait_mma();
        NBSS_TICK(0, 2, it_);
        if (tid == 0) load_image(ws0, a.img + IMG_WC2T, IMG_WC_BYTES, bar_w0);
        prefetch_saved(a.c2, slab);  // needed by E2, one MMA phase from now
        silu_epilogue(a.c3, a.g_c3, slab);
        end_epilogue();
        NBSS_TICK(0, 3, it_);
        // ---- B2: d s3 = conv3^T(g c3) ; E2: GroupNorm + SiLU backward
        convT_phase(w1a, bar_w1, ph_w1);
        NBSS_TICK(0, 4, it_);
        if (tid == 0) load_image(ws1, a.img + IMG_WC1T, IMG_WC_BYTES, bar_w1);
        prefetch_saved(a.c1, slab);  // E3
        {
            // GroupNorm + SiLU backward in two sweeps over the thread's 96 accumulator columns; the second sweep needs NO global
            // reads: sweep A parks the saved c2 bits in the G tile (dead after the conv^T MMAs) and writes
            // dn = d s3 * SiLU'(n) back over the accumulator columns in tensor memory (fp32, tcgen05.st).  Loops stay ROLLED:
            // this kernel is far larger than the instruction cache and every warp walks the code once per slab, so
            // straight-line unrolling costs more in instruction fetch than it saves (measured: 40k -> 57k cycles).
            const float* gst = a.gn_stats + (size_t)slab * 16;
            // sweep A, one group (24 channels) per iteration; the next group's saved c2 is requested one iteration ahead
            uint4 cq[3], cqn[3];
#pragma unroll
            for (int k = 0; k < 3; ++k)
                cq[k] = valid ? __ldg(reinterpret_cast<const uint4*>(a.c2 + tile_off(slab, 24, T, cb / 8 + k, t))) : make_uint4(0, 0, 0, 0);
#pragma unroll 1
            for (int gl = 0; gl < 4; ++gl) {
                const int g = 4 * hf + gl, c0 = kGC * g;
                uint32_t r[24];
#pragma unroll
                for (int k = 0; k < 3; ++k) tmem_ld8(tacc + c0 + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(r + 8 * k));
                if (gl < 3) {
#pragma unroll
                    for (int k = 0; k < 3; ++k)
                        cqn[k] = valid ? __ldg(reinterpret_cast<const uint4*>(a.c2 + tile_off(slab, 24, T, (c0 + kGC) / 8 + k, t))) : make_uint4(0, 0, 0, 0);
                }
                const float mean = __ldg(gst + 2 * g), rstd = __ldg(gst + 2 * g + 1);
                tmem_ld_wait();
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = c0 + 8 * k;
                    float cv[8];
                    uint32_t dn[8];
                    unpack_f16x2(cq[k].x, cv[0], cv[1]);
                    unpack_f16x2(cq[k].y, cv[2], cv[3]);
                    unpack_f16x2(cq[k].z, cv[4], cv[5]);
                    unpack_f16x2(cq[k].w, cv[6], cv[7]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float xh = (cv[j] - mean) * rstd;
                        const float n = fmaf(xh, s_gng[c + j], s_gnb[c + j]);
                        const float sg = sigmoidf_(n);
                        const float d = __uint_as_float(r[8 * k + j]) * sg * fmaf(n, 1.f - sg, 1.f) * vmask;  // frames >= T carry no gradient
                        const float dxh = d * s_gng[c + j];
                        s1 += dxh;
                        s2 = fmaf(dxh, xh, s2);
                        dn[j] = __float_as_uint(d);
                    }
                    tmem_st8(tacc + c, dn);
                    *reinterpret_cast<uint4*>(hrow + (c / 8) * kCS) = cq[k];
                }
                s1 = warp_sum(s1);
                s2 = warp_sum(s2);
                if (lane == 0) { red[(warp * 8 + g) * 2] = s1; red[(warp * 8 + g) * 2 + 1] = s2; }
#pragma unroll
                for (int k = 0; k < 3; ++k) cq[k] = cqn[k];
            }
            tmem_st_wait();
            NBSS_TICK(0, 20, it_);
            __syncthreads();
            if (tid < 16) {
                float s = 0.f;
#pragma unroll
                for (int w = 0; w < 8; ++w) s += red[((8 * ((tid >> 1) >> 2) + w) * 8 + (tid >> 1)) * 2 + (tid & 1)];
                gtot[tid] = s * inv_n;
            }
            __syncthreads();
            NBSS_TICK(0, 21, it_);
            // sweep B: g(c2) = rstd * (dn*gamma - S1/N - xhat*S2/N) -> G tile + global; column sums for d_gnw, d_gnb
#pragma unroll 1
            for (int c0 = cb; c0 < cb + 96; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tacc + c0, r);
                float dnv[16], dnx[16];
                uint4 pk[2];
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) pk[cc] = *reinterpret_cast<const uint4*>(hrow + (c0 / 8 + cc) * kCS);  // the c2 bits parked by sweep A
                tmem_ld_wait();
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int c = c0 + 8 * cc, g = c / kGC;
                    const float mean = __ldg(gst + 2 * g), rstd = __ldg(gst + 2 * g + 1), m1 = gtot[2 * g], m2 = gtot[2 * g + 1];
                    float cv[8], gv[8];
                    unpack_f16x2(pk[cc].x, cv[0], cv[1]);
                    unpack_f16x2(pk[cc].y, cv[2], cv[3]);
                    unpack_f16x2(pk[cc].z, cv[4], cv[5]);
                    unpack_f16x2(pk[cc].w, cv[6], cv[7]);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float xh = (cv[j] - mean) * rstd;
                        const float d = __uint_as_float(r[8 * cc + j]);  // 0 for frames >= T
                        dnv[8 * cc + j] = d;
                        dnx[8 * cc + j] = d * xh;
                        gv[j] = rstd * (d * s_gng[c + j] - m1 - xh * m2) * vmask;
                    }
                    const uint4 gp = pack8<FMT>(gv);
                    if (valid) *reinterpret_cast<uint4*>(a.g_c2 + tile_off(slab, 24, T, c / 8, t)) = gp;
                    *reinterpret_cast<uint4*>(hrow + (c / 8) * kCS) = gp;
                }
                const float sw = warp_colsum16(dnx, lane), sb = warp_colsum16(dnv, lane);
                if (!(lane & 1)) {
                    atomicAdd(acc + c0 + (lane >> 1), sw);
                    atomicAdd(acc + 192 + c0 + (lane >> 1), sb);
                }
            }
        }
        end_epilogue();
        NBSS_TICK(0, 5, it_);
        // ---- B3: d s2 = conv2^T(g c2)
        convT_phase(w0a, bar_w0, ph_w0);
        NBSS_TICK(0, 6, it_);
        prefetch_saved(a.a1, slab);  // E4
        if (tid == 0) load_image(ws0, a.img + IMG_W1T, IMG_W2_BYTES, bar_w0);
        silu_epilogue(a.c1, a.g_c1, slab);
        end_epilogue();
        NBSS_TICK(0, 7, it_);
        // ---- B4: d s1 = conv1^T(g c1)
        convT_phase(w1a, bar_w1, ph_w1);
        NBSS_TICK(0, 8, it_);
        // E5b of this slab reads x (fp32); the next slab starts with dy and its E1 reads c3
        if (tid >= 38 && tid < 44) l2_prefetch_slab(a.x + (size_t)slab * T * kH, T, tid - 38);
        if (tid >= 44 && tid < 50 && slab + (int)gridDim.x < a.nslab) l2_prefetch_slab(a.dy + (size_t)(slab + gridDim.x) * T * kH, T, tid - 44);
        prefetch_saved(a.c3, slab + gridDim.x);
        silu_epilogue(a.a1, a.g_a1, slab);
        end_epilogue();
        NBSS_TICK(0, 9, it_);
        // ---- B5: d ln = g(a1) W1 ; E5: LayerNorm backward + residual
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w0, ph_w0, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) {
                mma_kk(tmem + mm * 192, hb + (128 * mm + 1) * 16, kCS, w0a, 96 * 16, 12, id96, 0, leader);
                if (leader) umma_commit(mm ? bar_mma1 : bar_mma);
            }
        }
        ph_w0 ^= 1;
        wait_mma();
        NBSS_TICK(0, 10, it_);
        {
            // E5a: thread = (frame, channel half): d ln (fp32) staged into the dead G tile with 4-float chunks
#pragma unroll 1
            for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 16) {
                uint32_t r[16];
                tmem_ld16(tacc + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4)
                    *reinterpret_cast<float4*>(hrow + (c0 / 4 + j4) * kCS) =
                        make_float4(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                                    __uint_as_float(r[4 * j4 + 3]));
            }
            tc_fence_before();
            __syncthreads();
            NBSS_TICK(0, 11, it_);
            // E5b: eight lanes per frame: LayerNorm backward + residual, coalesced; d gamma / d beta -> smem accumulators
            Oct12 dlng, dlnb;
            dlng.zero();
            dlnb.zero();
            ln_bwd_rows(hbuf, kCS, 1, a.x + (size_t)slab * T * kH, dys, a.dx + (size_t)slab * T * kH, a.ln_stats + (size_t)slab * T * 2, T,
                        s_lng, dlng, dlnb, warp, lane, kFfnBwdThreads / 32);
            dlng.flush_atomic(acc + 384, lane);
            dlnb.flush_atomic(acc + 480, lane);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 12, it_);
    }
    // flush the affine-parameter gradients
    for (int i = tid; i < 192; i += kFfnBwdThreads) { atomicAdd(a.d_gnw + i, acc[i]); atomicAdd(a.d_gnb + i, acc[192 + i]); }
    for (int i = tid; i < 96; i += kFfnBwdThreads) { atomicAdd(a.d_lnw + i, acc[384 + i]); atomicAdd(a.d_lnb + i, acc[480 + i]); }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace nbss

NBSS_PHASE_READER(nbss_debug_phases_ffn_bwd)

extern "C" int nbss_ffn_bwd(const float* x, const float* dy, float* dx, int nslab, int T, const float* ln_w,
                            const float* gn_w, const float* gn_b, const float* ln_stats, const float* gn_stats,
                            const void* layer_img, const void* a1, const void* c1, const void* c2, const void* c3,
                            void* g_a1, void* g_c1, void* g_c2, void* g_c3,
                            float* d_lnw, float* d_lnb, float* d_gnw, float* d_gnb, int fmt, int* err, void* stream) {
    using namespace nbss;
    if (!x || !dy || !dx || !ln_w || !gn_w || !gn_b || !ln_stats || !gn_stats || !layer_img || !a1 || !c1 || !c2 || !c3 ||
        !g_a1 || !g_c1 || !g_c2 || !g_c3 || !d_lnw || !d_lnb || !d_gnw || !d_gnb)
        return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    FfnBwdArgs a{x, dy, dx, nslab, T, ln_w, gn_w, gn_b, ln_stats, gn_stats, (const unsigned char*)layer_img,
                 (const unsigned char*)a1, (const unsigned char*)c1, (const unsigned char*)c2, (const unsigned char*)c3,
                 (unsigned char*)g_a1, (unsigned char*)g_c1, (unsigned char*)g_c2, (unsigned char*)g_c3,
                 d_lnw, d_lnb, d_gnw, d_gnb, err};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = nslab < sms ? nslab : sms;
    auto kern = (fmt == FMT_F16) ? ffn_bwd_kernel<FMT_F16> : ffn_bwd_kernel<FMT_BF16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FB_SMEM);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kFfnBwdThreads, FB_SMEM, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
