// mhsa_bwd.cu — backward (data gradient) of the narrow-band MHSA sub-block, tcgen05 + TMEM, T <= 256.
//
//   mhsa_bwd_core : per (b,f) slab.  dO = dy Wo (tensor core), delta = rowsum(dO * O), then per head and per
//                   (query tile, key tile) block of 128 x 128:  S = Qs K^T and dP = dO V^T (TMEM), P = exp2(S - lse),
//                   dS = P (dP - delta) staged as 16-bit tiles, dQ += dS K, dK += dS^T Qs, dV += P^T dO with the
//                   accumulators living in TMEM across the blocks.  Writes dQKV [n,288] (gradient wrt the in-proj
//                   outputs) for the LN/in-proj backward and the weight-gradient kernel.
//   mhsa_bwd_ln   : per slab.  d ln = dQKV Win (K = 288), LayerNorm backward, dx = dy + ..., d gamma / d beta.
// Inputs saved by mhsa_fwd: fp16 (scaled q | k | v), fp16 O, log2-domain logsumexp, LN statistics.
// FMT_G is the 16-bit format of gradient operands (bf16); q,k,v,O stay fp16 (mixed-format kind::f16 MMAs).
#include "slab.cuh"

namespace nbss {

constexpr uint32_t kCSQ = 129 * 16;  // chunk stride of 128-row tiles (P, dS)
constexpr int kMbThreads = 512;      // 16 warps: M-tile (w>>2)&1, TMEM lane quarter w&3, half w>>3 / key quarter w>>2

struct MhsaBwdArgs {
    const float* dy;
    int nslab, T;
    const unsigned char* img;
    const unsigned char* qkv;  // fp16 [n,288] scaled q | k | v
    const unsigned char* o;    // fp16 [n,96]
    const float* lse;          // [nslab,4,T]
    unsigned char* dqkv;       // FMT_G [n,288]
    int* err;
};

// smem map (bytes)
constexpr uint32_t MB_DO = 0;                       // dO tile, 13 chunks x kCS (chunk 12 = zeros); first holds dy (12 chunks)
constexpr uint32_t MB_Q = 13 * kCS;                 // per-head Qs | K | V tiles [256 x 32] fp16, 4 chunks each, TWO sets: the next head's
constexpr uint32_t MB_QKV_SET = 12 * kCS;           // q,k,v arrive by TMA while the current head is processed
constexpr uint32_t MB_P = MB_Q + 2 * MB_QKV_SET;    // P tile [128 q x 128 keys], 16 chunks x kCSQ; aliases the WoT image
constexpr uint32_t MB_DS = MB_P + 16 * kCSQ;        // dS tile
constexpr uint32_t MB_DELTA = MB_DS + 16 * kCSQ;    // delta [4][256] floats
constexpr uint32_t MB_LSE = MB_DELTA + 4096;        // lse   [4][256] floats
constexpr uint32_t MB_BAR = MB_LSE + 4096;
constexpr uint32_t MB_SMEM = MB_BAR + 64;
static_assert(IMG_WQ_BYTES <= 16 * kCSQ, "WoT image must fit the P region");

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <int FMT_G>
__global__ void __launch_bounds__(kMbThreads, 1) mhsa_bwd_core_kernel(MhsaBwdArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* dot = smem + MB_DO;
    unsigned char* pt = smem + MB_P;
    unsigned char* dst = smem + MB_DS;
    float* s_delta = reinterpret_cast<float*>(smem + MB_DELTA);
    float* s_lse = reinterpret_cast<float*>(smem + MB_LSE);
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + MB_BAR);
    uint64_t* bar_w = bar_mma + 1;
    uint64_t* bar_q = bar_mma + 2;  // [2] q,k,v of tile set s have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 4);

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31, T = a.T;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_q, 1);
        mbar_init(bar_q + 1, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < (int)(MB_DELTA / 16); i += kMbThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int m = (warp >> 2) & 1, q = warp & 3, hf = warp >> 3, kq = warp >> 2, rt = 32 * q + lane, t = 128 * m + rt;
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t doa = smem_u32(dot), pa = smem_u32(pt), dsa = smem_u32(dst);
    // instruction descriptors: (A fmt, B fmt, A major, B major, N)
    auto idesc = [](uint32_t fa, uint32_t fb, uint32_t amn, uint32_t bmn, uint32_t n) {
        return (1u << 4) | (fa << 7) | (fb << 10) | (amn << 15) | (bmn << 16) | ((n >> 3) << 17) | (8u << 24);
    };
    const uint32_t id_do = idesc(FMT_G, FMT_G, 0, 0, 96);       // dO = dy WoT
    const uint32_t id_s = idesc(FMT_F16, FMT_F16, 0, 0, 128);   // S = Qs K^T
    const uint32_t id_dp = idesc(FMT_G, FMT_F16, 0, 0, 128);    // dP = dO V^T
    const uint32_t id_dq = idesc(FMT_G, FMT_F16, 0, 1, 32);     // dQ = dS K      (B MN-major)
    const uint32_t id_dk = idesc(FMT_G, FMT_F16, 1, 1, 32);     // dK = dS^T Qs   (A, B MN-major)
    const uint32_t id_dv = idesc(FMT_G, FMT_G, 1, 1, 32);       // dV = P^T dO
    // TMEM columns
    constexpr uint32_t C_S = 0, C_DP = 128, C_DQ = 256, C_DK = 320, C_DV = 384;  // dO (M0) uses 0..191
    const float kscale = 0.6931471805599453f;                   // dK was formed with log2e-scaled q
    const float qscale = rsqrtf((float)kDH);
    uint32_t ph_mma = 0, ph_w = 0, ph_q = 0;  // bit s of ph_q: phase of tile set s's barrier
    // ONE thread: the head's Qs, K, V (fp16, 24 -> 32 features: the 4th chunk of every tile stays zero): three chunk columns each,
    // straight TMA bulk copies out of the slab-tile tensor written by mhsa_fwd, into tile set h & 1
    auto load_qkv = [&](int slab, int h) {
        unsigned char* base = smem + MB_Q + (h & 1) * MB_QKV_SET;
        uint64_t* bar = bar_q + (h & 1);
        mbar_expect_tx(bar, (uint32_t)(9 * T * 16));
        for (int which = 0; which < 3; ++which)
            for (int c = 0; c < 3; ++c)
                bulk_g2s(base + (4 * which + c) * kCS, a.qkv + tile_off(slab, 36, T, 12 * which + 3 * h + c, 0), (uint32_t)(T * 16), bar);
    };

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

    stagger_start(78000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    for (int slab = blockIdx.x; slab < a.nslab; slab += gridDim.x, ++it_) {
        const size_t row0 = (size_t)slab * T;
        NBSS_TICK(0, 0, it_);
        if (tid == 0) {
            load_image(pt, a.img + IMG_WOT, IMG_WQ_BYTES, bar_w);
            load_qkv(slab, 0);  // every MMA of the previous slab has completed: tile set 0 is free
        }
        for (int i = tid; i < kNH * 256; i += kMbThreads) {
            const int h = i >> 8, tt = i & 255;
            s_lse[i] = tt < T ? a.lse[((size_t)slab * kNH + h) * T + tt] : 0.f;
        }
        stage_rows96<FMT_G, false>(a.dy + row0 * kH, T, dot, 0, nullptr, nullptr, warp, lane, nullptr, kMbThreads / 32);
        end_epilogue();
        NBSS_TICK(0, 1, it_);
        // ---- M0: dO = dy Wo
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + mm * 96, doa + 128 * mm * 16, kCS, pa, 96 * 16, 6, id_do, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(0, 2, it_);
        // ---- E0: dO -> 16-bit tile (in place of dy), delta_h = sum_c dO O
        {
            const bool valid = t < T;
            const uint32_t tacc = tmem + lane_off + m * 96;
#pragma unroll 1
            for (int h = 2 * hf; h < 2 * hf + 2; ++h) {  // each channel half owns two heads
                float dl = 0.f;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int c = kDH * h + 8 * k;
                    uint32_t r[8];
                    tmem_ld8(tacc + c, r);
                    tmem_ld_wait();
                    float v[8], ov[8];
                    if (valid) {
                        const uint4 oq = __ldg(reinterpret_cast<const uint4*>(a.o + tile_off(slab, 12, T, c / 8, t)));
                        unpack_f16x2(oq.x, ov[0], ov[1]);
                        unpack_f16x2(oq.y, ov[2], ov[3]);
                        unpack_f16x2(oq.z, ov[4], ov[5]);
                        unpack_f16x2(oq.w, ov[6], ov[7]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = valid ? __uint_as_float(r[j]) : 0.f;
                        dl += valid ? v[j] * ov[j] : 0.f;
                    }
                    *reinterpret_cast<uint4*>(dot + (c / 8) * kCS + t * 16) = pack8<FMT_G>(v);
                }
                s_delta[h * 256 + t] = dl;
            }
        }
        end_epilogue();
        NBSS_TICK(0, 3, it_);
        // next slab's upstream gradient -> L2, issued after this slab's own staging loads are done
        if (tid >= 32 && tid < 38 && slab + (int)gridDim.x < a.nslab) l2_prefetch_slab(a.dy + (size_t)(slab + gridDim.x) * T * kH, T, tid - 32);
        // ---- heads
#pragma unroll 1
        for (int h = 0; h < kNH; ++h) {
            // q,k,v of this head were requested one head ago (tile set h & 1); request the next head's into the other set,
            // whose last readers (the MMAs of head h - 1) completed before that head's read-out
            const uint32_t qa = smem_u32(smem + MB_Q + (h & 1) * MB_QKV_SET), ka = qa + 4 * kCS, va = qa + 8 * kCS;
            if (tid == 0 && h + 1 < kNH) load_qkv(slab, h + 1);
            mbar_wait(bar_q + (h & 1), (ph_q >> (h & 1)) & 1u, a.err);
            ph_q ^= 1u << (h & 1);
            NBSS_TICK(0, 4 + 12 * h, it_);
            // Software pipeline over the four 128x128 (query, key) blocks of the head: the S / dP MMAs of block b+1 are
            // issued together with the gradient MMAs of block b (different TMEM columns), so a block costs one
            // commit / wait round trip instead of two.
            auto issue_s_dp = [&](int blk, bool leader) {
                const int qb = blk >> 1, kb = blk & 1;
                mma_kk(tmem + C_S, qa + 128 * qb * 16, kCS, ka + 128 * kb * 16, kCS, 2, id_s, 0, leader);
                mma_kk(tmem + C_DP, doa + 3 * h * kCS + 128 * qb * 16, kCS, va + 128 * kb * 16, kCS, 2, id_dp, 0, leader);
            };
            if (warp == 0) {
                tc_fence_after();
                const bool leader = elect_one();
                issue_s_dp(0, leader);
                if (leader) umma_commit(bar_mma);
            }
#pragma unroll 1
            for (int blk = 0; blk < 4; ++blk) {
                const int qb = blk >> 1, kb = blk & 1;
                wait_mma();  // S, dP of this block (and the gradient MMAs of the previous one, which read the P / dS tiles)
                NBSS_TICK(0, 5 + 12 * h + 2 * blk, it_);
                // P and dS for this block: thread = (query row rt, key quarter kq: 32 of the 128 keys)
                {
                    const int tq = 128 * qb + rt, key0 = 128 * kb + 32 * kq;
                    // a query row beyond T gets lse = +inf: every P of the row is ex2(-inf) = 0 without a per-element select
                    const float lse = tq < T ? s_lse[h * 256 + tq] : INFINITY, dl = s_delta[h * 256 + tq];
                    uint32_t rs[32], rp[32];
                    tmem_ld32(tmem + lane_off + C_S + 32 * kq, rs);
                    tmem_ld32(tmem + lane_off + C_DP + 32 * kq, rp);
                    tmem_ld_wait();
                    float p[32], ds[32];
                    if (key0 + 32 <= T) {  // warp-uniform: all 32 keys of this quarter exist (every quarter but the slab's last)
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float pv = ex2f(__uint_as_float(rs[j]) - lse);
                            p[j] = pv;
                            ds[j] = pv * (__uint_as_float(rp[j]) - dl);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float pv = ex2f(key0 + j < T ? __uint_as_float(rs[j]) - lse : -INFINITY);
                            p[j] = pv;
                            ds[j] = pv * (__uint_as_float(rp[j]) - dl);
                        }
                    }
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int chunk = 4 * kq + cc;
                        *reinterpret_cast<uint4*>(pt + chunk * kCSQ + rt * 16) = pack8<FMT_G>(p + 8 * cc);
                        *reinterpret_cast<uint4*>(dst + chunk * kCSQ + rt * 16) = pack8<FMT_G>(ds + 8 * cc);
                    }
                }
                end_epilogue();
                NBSS_TICK(0, 6 + 12 * h + 2 * blk, it_);
                if (warp == 0) {
                    tc_fence_after();
                    const bool leader = elect_one();
                    for (int ks = 0; ks < 8; ++ks) {
                        // dQ[qb] += dS K[kb]: A = dS (K-major, K = keys), B = K tile MN-major (K = key rows)
                        if (leader) umma_f16(tmem + C_DQ + 32 * qb, sdesc_kmajor(dsa + 2 * ks * kCSQ, kCSQ),
                                 sdesc_mnmajor(ka + (128 * kb + 16 * ks) * 16, kCS), id_dq, (kb | ks) ? 1u : 0u);
                        // dK[kb] += dS^T Qs[qb]: A = dS MN-major (M = keys, K = query rows), B = Qs MN-major
                        if (leader) umma_f16(tmem + C_DK + 32 * kb, sdesc_mnmajor(dsa + 16 * ks * 16, kCSQ),
                                 sdesc_mnmajor(qa + (128 * qb + 16 * ks) * 16, kCS), id_dk, (qb | ks) ? 1u : 0u);
                        // dV[kb] += P^T dO[qb]
                        if (leader) umma_f16(tmem + C_DV + 32 * kb, sdesc_mnmajor(pa + 16 * ks * 16, kCSQ),
                                 sdesc_mnmajor(doa + 3 * h * kCS + (128 * qb + 16 * ks) * 16, kCS), id_dv, (qb | ks) ? 1u : 0u);
                    }
                    if (blk < 3) issue_s_dp(blk + 1, leader);
                    if (leader) umma_commit(bar_mma);
                }
            }
            wait_mma();  // the gradient MMAs of the last block
            NBSS_TICK(0, 13 + 12 * h, it_);
            // read out dQ_h, dK_h, dV_h (thread = frame t of tile m) -> dQKV
            {
                // tcgen05.ld is warp-collective: every lane issues it, only valid frames store
                const bool valid = t < T;
#pragma unroll
                for (int which = 0; which < 3; ++which) {
                    if ((which == 2) != (hf == 1)) continue;  // half 0 writes dQ and dK, half 1 writes dV (warp-uniform)
                    const uint32_t col = (which == 0 ? C_DQ : (which == 1 ? C_DK : C_DV)) + 32 * m;
                    const float sc = which == 0 ? qscale : (which == 1 ? kscale : 1.f);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        uint32_t r[8];
                        tmem_ld8(tmem + lane_off + col + 8 * k, r);
                        tmem_ld_wait();
                        float v[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]) * sc;
                        if (valid) *reinterpret_cast<uint4*>(a.dqkv + tile_off(slab, 36, T, 12 * which + 3 * h + k, t)) = pack8<FMT_G>(v);
                    }
                }
            }
            tc_fence_before();
            __syncthreads();
            NBSS_TICK(0, 14 + 12 * h, it_);
        }
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ LN / in-proj backward
struct MhsaLnArgs {
    const float* x;
    const float* dy;
    float* dx;
    int nslab, T;
    const float* ln_w;
    const float* ln_stats;
    const unsigned char* img;
    const unsigned char* dqkv;  // FMT_G [n,288]
    float *d_lnw, *d_lnb;
    int* err;
};
constexpr uint32_t ML_A = 0;                        // dQKV tile 36 chunks
constexpr uint32_t ML_W = 36 * kCS;                 // WinT image 55296
constexpr uint32_t ML_CST = ML_W + IMG_WINT_BYTES;  // ln_w 96 + acc 192 floats
constexpr uint32_t ML_XCH = ML_CST + 288 * 4;      // partial LayerNorm row sums of the two channel halves
constexpr uint32_t ML_BAR = ML_XCH + 512 * 8;
constexpr uint32_t ML_SMEM = ML_BAR + 64;

template <int FMT_G>
__global__ void __launch_bounds__(kMbThreads, 1) mhsa_bwd_ln_kernel(MhsaLnArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* at = smem + ML_A;
    unsigned char* wt = smem + ML_W;
    float* s_lng = reinterpret_cast<float*>(smem + ML_CST);
    float* acc = s_lng + 96;  // [0,96) d_lnw, [96,192) d_lnb
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + ML_BAR);
    uint64_t* bar_w = bar_mma + 1;
    uint64_t* bar_ld = bar_mma + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 3);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31, T = a.T;
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        mbar_init(bar_ld, 1);
        fence_mbar_init();
        load_image(wt, a.img + IMG_WINT, IMG_WINT_BYTES, bar_w);  // resident for the whole kernel
    }
    for (int i = tid; i < 96; i += kMbThreads) s_lng[i] = a.ln_w[i];
    for (int i = tid; i < 192; i += kMbThreads) acc[i] = 0.f;
    for (int i = tid; i < (int)(36 * kCS / 16); i += kMbThreads) reinterpret_cast<uint4*>(at)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int m = (warp >> 2) & 1, q = warp & 3, hf = warp >> 3, t = 128 * m + 32 * q + lane;
    float2* xch = reinterpret_cast<float2*>(smem + ML_XCH);
    const bool valid = t < T;
    const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16) + m * 96;
    const uint32_t ata = smem_u32(at), wta = smem_u32(wt);
    const uint32_t id96 = make_idesc(FMT_G, 128, 96, 0, 0);
    uint32_t ph = 0, ph_ld = 0;
    bool wready = false;
    stagger_start(26000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    for (int slab = blockIdx.x; slab < a.nslab; slab += gridDim.x, ++it_) {
        const size_t row0 = (size_t)slab * T, grow = row0 + t;
        NBSS_TICK(1, 0, it_);
        if (tid == 0) bulk_load_chunks(at, kCS, 0, a.dqkv + tile_off(slab, 36, T, 0, 0), 36, T, bar_ld);
        if (tid >= 32 && tid < 44) {  // this slab's x and dy rows (LayerNorm backward at the end of the iteration) -> L2
            const int i = tid - 32;
            l2_prefetch_slab((i < 6 ? a.x : a.dy) + row0 * kH, T, i % 6);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (!wready) mbar_wait(bar_w, 0, a.err);
            mbar_wait(bar_ld, ph_ld, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + mm * 96, ata + 128 * mm * 16, kCS, wta, 96 * 16, 18, id96, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        wready = true;
        ph_ld ^= 1;
        __syncwarp();
        mbar_wait(bar_mma, ph, a.err);
        ph ^= 1;
        tc_fence_after();
        NBSS_TICK(1, 1, it_);
        // thread = (frame, channel half): d ln (fp32) staged into the dead dQKV tile (4-float chunks), then one warp per
        // frame does the LayerNorm backward + residual with coalesced global traffic (slab.cuh: ln_bwd_rows)
#pragma unroll 1
        for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 16) {
            uint32_t r[16];
            tmem_ld16(tacc + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4)
                *reinterpret_cast<float4*>(at + (size_t)(c0 / 4 + j4) * kCS + t * 16) =
                    make_float4(__uint_as_float(r[4 * j4 + 0]), __uint_as_float(r[4 * j4 + 1]), __uint_as_float(r[4 * j4 + 2]),
                                __uint_as_float(r[4 * j4 + 3]));
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 2, it_);
        {
            Oct12 dlng, dlnb;
            dlng.zero();
            dlnb.zero();
            ln_bwd_rows(at, kCS, 0, a.x + row0 * kH, a.dy + row0 * kH, a.dx + row0 * kH, a.ln_stats + row0 * 2, T, s_lng, dlng, dlnb, warp, lane,
                        kMbThreads / 32);
            dlng.flush_atomic(acc, lane);
            dlnb.flush_atomic(acc + 96, lane);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(1, 3, it_);
    }
    for (int i = tid; i < 96; i += kMbThreads) { atomicAdd(a.d_lnw + i, acc[i]); atomicAdd(a.d_lnb + i, acc[96 + i]); }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace nbss

NBSS_PHASE_READER(nbss_debug_phases_mhsa_bwd)

// Backward of y = x + MHSA(LN(x)).  Writes dqkv (scratch, fmt_g [n,288], also consumed by nbss_mhsa_wgrad) and dx;
// accumulates d_lnw / d_lnb.
extern "C" int nbss_mhsa_bwd(const float* x, const float* dy, float* dx, int nslab, int T, const float* ln_w,
                             const float* ln_stats, const void* layer_img, const void* qkv, const void* o, const float* lse,
                             void* dqkv, float* d_lnw, float* d_lnb, int fmt_g, int* err, void* stream) {
    using namespace nbss;
    if (!x || !dy || !dx || !ln_w || !ln_stats || !layer_img || !qkv || !o || !lse || !dqkv || !d_lnw || !d_lnb) return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    if (fmt_g != FMT_BF16 && fmt_g != FMT_F16) return NBSS_ERR_UNSUPPORTED;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = nslab < sms ? nslab : sms;
    cudaStream_t st = (cudaStream_t)stream;
    {
        MhsaBwdArgs a{dy, nslab, T, (const unsigned char*)layer_img, (const unsigned char*)qkv, (const unsigned char*)o, lse,
                      (unsigned char*)dqkv, err};
        auto kern = (fmt_g == FMT_BF16) ? mhsa_bwd_core_kernel<FMT_BF16> : mhsa_bwd_core_kernel<FMT_F16>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MB_SMEM);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kMbThreads, MB_SMEM, st>>>(a);
        NBSS_LAUNCH_CHECK();
    }
    {
        MhsaLnArgs a{x, dy, dx, nslab, T, ln_w, ln_stats, (const unsigned char*)layer_img, (const unsigned char*)dqkv, d_lnw, d_lnb, err};
        auto kern = (fmt_g == FMT_BF16) ? mhsa_bwd_ln_kernel<FMT_BF16> : mhsa_bwd_ln_kernel<FMT_F16>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ML_SMEM);
        if (e != cudaSuccess) return (int)e;
        kern<<<grid, kMbThreads, ML_SMEM, st>>>(a);
        NBSS_LAUNCH_CHECK();
    }
    return NBSS_OK;
}
