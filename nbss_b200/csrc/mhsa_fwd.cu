// mhsa_fwd.cu — narrow-band multi-head self-attention sub-block, forward, one CTA per (b,f) slab, tcgen05 + TMEM.
//
// Replaces SpatialNetLayer._tsa + residual (models/arch/SpatialNet.py:88-89,93-100) i.e. LN -> nn.MultiheadAttention
// (packed in-proj, q*dh^-0.5, softmax over T, out-proj; no mask, no dropout) for T <= 256:
//   P0 stage x -> LN -> fp16 A0 [256x96]
//   P1 K|V = A0 Wkv^T  (N=192)         E1: +bias -> K tile (per head padded 24->32, zero chunk) and V tile
//   P2 Q   = A0 Wq^T   (N=96)          EQ: (Q + b) * dh^-0.5 * log2(e) -> fp16 Q tile (over the dead A0), all heads at once
//   per head h, query tiles a = (h,0), b = (h,1), two 256-column score buffers in tensor memory:
//        S = Q_h K_h^T (N = 256 keys)  issued one head ahead of its softmax
//        softmax WITHOUT a cross-thread exchange (split-K as in flash attention): thread = (query row, key quarter) takes
//        the max of ITS 64 scores, writes exp2(s - m_q) as fp16 pairs back over the first 32 of its own 64 score columns
//        and leaves (m_q, l_q) in shared memory; the P V_h MMA runs per key quarter (A operand from tensor memory) into the
//        last 32 columns of the quarter; the read-out combines the four partial outputs with exp2(m_q - m) and 1/l.
//        One __syncthreads per softmax and per read-out; the S / PV MMAs of one query tile run under the softmax / read-out
//        of the other.
//   P3 y = x + O Wo^T + bo   (O_h(tile) overwrites Q_h(tile) in the Q tile once its scores are done)
// HBM traffic: x in, y out (+ fp16 q|k|v, O and the log2-sum-exp when save != 0, for the backward kernels).
//
// T > 256 (inference; SURVEY.md §8 f4: validation / test utterances are longer than the 4 s training crops) runs as two passes of
// the same kernel over work items (slab, 256-frame chunk):
//   LONG = 1  P0, P1, E1 only: K | V of every chunk -> the fp16 slab-tile tensor [nslab][36][T][8] (the training path's save layout)
//   LONG = 2  P0, P2, EQ for the chunk's queries, then flash-style attention over the key blocks of the slab: step n = (head h, key
//             block kb) loads K_h[kb], V_h[kb] by TMA (two tile sets, one step ahead) and runs the SAME S / split-K softmax / P.V
//             machinery on the chunk's two query tiles; the read-out folds each block into a running (max, sum, output) per thread
//             and writes O_h over Q_h after the head's last block; P3 as before.
#include "slab.cuh"

namespace nbss {

struct MhsaFwdArgs {
    const float* x;
    float* y;
    int nslab, T;
    const float *ln_w, *ln_b, *b_in, *b_out;
    const unsigned char* img;
    unsigned char* save_qkv;  // fp16 slab-tile [nslab][36][T][8]: (scaled q | k | v) or null; LONG: the k | v exchange tensor
    unsigned char* save_o;    // fp16 slab-tile [nslab][12][T][8] or null
    float* save_lse;          // [nslab, 4, T] log2-domain logsumexp or null
    float* ln_stats;          // [nslab*T, 2] (mean, rstd) of the LayerNorm or null
    float* row_part;          // [nslab*T, 2] (sum, sum of squares) over the channels of every output row, or null (NBC2 GroupBatchNorm)
    int* err;
};

constexpr uint32_t MH_AO = 0;                       // A0, then the Q tile, then O: 13 chunks (the 13th stays zero)
constexpr uint32_t MH_K = 13 * kCS;                 // 55120: K tile, 16 chunks (4 heads: head h = chunks 4h..4h+2, 4h+3 = zero pad; 2 heads: 6 chunks each)
constexpr uint32_t MH_V = MH_K + 16 * kCS;          // 122960: V tile, 13 chunks (the 13th stays zero)
constexpr uint32_t MH_W = MH_V + 13 * kCS;          // 178080: one weight image at a time
constexpr uint32_t MH_W_BYTES = IMG_W1_BYTES;       // 36864
constexpr uint32_t MH_CST = MH_W + MH_W_BYTES;      // 214944
constexpr uint32_t MH_STAT = MH_CST + 576 * 4;      // (m_q, l_q) [2 buffers][4 key quarters][128 rows] float2 = 8192
constexpr uint32_t MH_BAR = MH_STAT + 8192;
constexpr int kMhThreads = 512;  // 16 warps
constexpr uint32_t MH_SMEM = MH_BAR + 96;  // 7 barriers + the TMEM slot (+ 2 K/V-set barriers of the long-sequence mode)
static_assert(IMG_WQ_BYTES <= MH_W_BYTES, "weight image must fit the W region");

// NHEADS = 4 (SpatialNet-small: head dim 24, padded to 32 in the K / O tiles) or 2 (NBC2: head dim 48, no padding).
// DBUF: the four partial-output accumulators of a query tile fit the spare columns of its score buffer (head dim <= 32), so two
// score buffers ping-pong; otherwise (head dim 48) one score buffer [0,256) and the partial outputs at [256, 256 + 4*48).
template <int FMT, int NHEADS, int LONG = 0>
__global__ void __launch_bounds__(kMhThreads, 1) mhsa_fwd_kernel(MhsaFwdArgs a) {
    static_assert(LONG == 0 || NHEADS == 4, "the long-sequence passes exist for the 4-head layer only");
    constexpr int DH = kH / NHEADS;                 // 24 | 48
    constexpr int KCH = DH / 8;                     // data chunks per head: 3 | 6
    constexpr int HS = (DH % 16) ? KCH + 1 : KCH;   // chunk stride of a head in the K / O tile (one zero pad chunk if DH % 16): 4 | 6
    constexpr int KKS = (DH + 15) / 16;             // k-steps of S = Q_h K_h^T: 2 | 3
    constexpr int NPV = 16 * KKS;                   // N of the PV MMA: 32 | 48
    constexpr bool DBUF = 4 * NPV <= 128;
    constexpr int CPT = NPV == 32 ? 8 : 16;         // output features per read-out thread (3 threads per query row)
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* ao = smem + MH_AO;
    unsigned char* kt = smem + MH_K;
    unsigned char* vt = smem + MH_V;
    unsigned char* wr = smem + MH_W;
    float* cst = reinterpret_cast<float*>(smem + MH_CST);
    float *s_lng = cst, *s_lnb = cst + 96, *s_bin = cst + 192, *s_bout = cst + 480;
    float2* stat = reinterpret_cast<float2*>(smem + MH_STAT);  // [buf][kq][row]
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + MH_BAR);  // projections
    uint64_t* bar_w = bar_mma + 1;
    uint64_t* bar_s = bar_mma + 2;   // [2] scores of buffer b are complete
    uint64_t* bar_pv = bar_mma + 4;  // [2] partial outputs of buffer b are complete
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 6);
    uint64_t* bar_kv = bar_mma + 7;  // [2] LONG = 2: K_h / V_h block of tile set s has landed

    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const int T = a.T;
    if (warp == 0) tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(bar_s + i, 1); mbar_init(bar_pv + i, 1); mbar_init(bar_kv + i, 1); }
        fence_mbar_init();
    }
    for (int i = tid; i < 96; i += kMhThreads) { s_lng[i] = a.ln_w[i]; s_lnb[i] = a.ln_b[i]; s_bout[i] = a.b_out[i]; }
    for (int i = tid; i < 288; i += kMhThreads) s_bin[i] = a.b_in[i];
    // zero everything that is read as padding: AO (13th chunk), K (pad chunks), V (13th chunk)
    for (int i = tid; i < (int)(MH_W / 16); i += kMhThreads) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // 16 warps.  'Both tiles' epilogues: M-tile m, lane quarter q, column half hf.  Softmax / read-out: key quarter kq.
    const int m = (warp >> 2) & 1, q = warp & 3, hf = warp >> 3, kq = warp >> 2;
    const int rt = 32 * q + lane;       // row within an M-tile (TMEM lane)
    const int t = 128 * m + rt;         // frame handled in "both tiles" epilogues
    const uint32_t lane_off = (uint32_t)(32 * q) << 16;
    const uint32_t aoa = smem_u32(ao), kta = smem_u32(kt), vta = smem_u32(vt), wra = smem_u32(wr);
    const uint32_t id192 = make_idesc(FMT, 128, 192, 0, 0), id96 = make_idesc(FMT, 128, 96, 0, 0),
                   id256 = make_idesc(FMT, 128, 256, 0, 0), idpv = make_idesc(FMT, 128, NPV, 0, 1);
    const float qscale = rsqrtf((float)DH) * 1.4426950408889634f;
    auto s_col = [&](int buf) -> uint32_t { return DBUF ? 256u * buf : 0u; };                          // score buffer
    auto o_col = [&](int buf, int kk) -> uint32_t { return DBUF ? 256u * buf + 64u * kk + 32u : 256u + NPV * kk; };  // partial outputs
    uint32_t ph_mma = 0, ph_w = 0, ph_s = 0, ph_pv = 0;  // bit b of ph_s / ph_pv: phase of buffer b's barrier
    int Tc = T, t0 = 0;  // LONG: frames of this work item's chunk, its first frame
    uint32_t ph_kv = 0;  // LONG = 2: bit s = phase of tile set s's barrier (only warp 0 waits on it)

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
    // warp 0: S(h, mq) = Q_h K_h^T into score buffer `buf`; completion on bar_s[buf]
    auto issue_s = [&](int h, int mq, int buf) {
        tc_fence_after();
        const bool leader = elect_one();
        mma_kk(tmem + s_col(buf), aoa + KCH * h * kCS + 128 * mq * 16, kCS, kta + HS * h * kCS, kCS, KKS, id256, 0, leader);
        if (leader) umma_commit(bar_s + buf);
        __syncwarp();
    };
    // softmax of one query tile over the thread's own 64 keys; P -> TMEM (first 32 of the thread's 64 score columns)
    auto softmax_local = [&](int b, int tk) {
        const bool kmask = 64 * kq + 63 >= tk;  // warp-uniform: this key quarter holds keys >= tk (they get probability 0)
        const uint32_t ts = tmem + lane_off + s_col(b) + 64 * kq;
        uint32_t r0[32], r1[32];
        tmem_ld32(ts, r0);
        tmem_ld32(ts + 32, r1);
        tmem_ld_wait();
        if (kmask) {  // keys >= T: -inf (only the last key quarter(s) of a short slab pay for this)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                if (64 * kq + j >= tk) r0[j] = 0xff800000u;
                if (64 * kq + 32 + j >= tk) r1[j] = 0xff800000u;
            }
        }
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            m4[j & 1] = fmaxf(m4[j & 1], __uint_as_float(r0[j]));
            m4[2 + (j & 1)] = fmaxf(m4[2 + (j & 1)], __uint_as_float(r1[j]));
        }
        const float mq_ = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        const float mref = mq_ == -INFINITY ? 0.f : mq_;  // a fully masked quarter: every exponent stays -inf -> p = 0
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_ftz(__uint_as_float(r0[j]) - mref), p1 = ex2_ftz(__uint_as_float(r0[j + 1]) - mref);
            s4[0] += p0;
            s4[1] += p1;
            pk[j >> 1] = pack16<FMT>(p0, p1);
        }
        tmem_st16(ts, pk);
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2_ftz(__uint_as_float(r1[j]) - mref), p1 = ex2_ftz(__uint_as_float(r1[j + 1]) - mref);
            s4[2] += p0;
            s4[3] += p1;
            pk[j >> 1] = pack16<FMT>(p0, p1);
        }
        tmem_st16(ts + 16, pk);
        stat[(b * 4 + kq) * 128 + rt] = make_float2(mq_, (s4[0] + s4[1]) + (s4[2] + s4[3]));
        tmem_st_wait();
    };
    // warp 0: O_q(h, mq) = P_q V_h for the four key quarters (A operand from tensor memory); completion on bar_pv[mq]
    auto issue_pv = [&](int h, int b) {
        tc_fence_after();
        const bool leader = elect_one();
#pragma unroll 1
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll 1
            for (int ks = 0; ks < 4; ++ks)
                if (leader)
                    umma_f16_ts(tmem + o_col(b, kk), tmem + s_col(b) + 64 * kk + 8 * ks,
                                sdesc_mnmajor(vta + KCH * h * kCS + (64 * kk + 16 * ks) * 16, kCS), idpv, ks ? 1u : 0u);
        if (leader) umma_commit(bar_pv + b);
        __syncwarp();
    };
    // read-out of query tile mq (score buffer b): combine the four partial outputs, normalise, saves; O_h(tile mq) goes into the Q tile
    // over Q_h(tile mq), which is dead once S(h, mq) has completed (the compact O tile is the out-proj's A operand)
    auto readout = [&](int h, int mq, int b, int slab) {
        const int tq = 128 * mq + rt;
        const float2 s0 = stat[(b * 4 + 0) * 128 + rt], s1 = stat[(b * 4 + 1) * 128 + rt], s2 = stat[(b * 4 + 2) * 128 + rt],
                     s3 = stat[(b * 4 + 3) * 128 + rt];
        const float mx = fmaxf(fmaxf(s0.x, s1.x), fmaxf(s2.x, s3.x));  // finite: key quarter 0 always holds key 0 < T
        const float f[4] = {ex2_ftz(s0.x - mx), ex2_ftz(s1.x - mx), ex2_ftz(s2.x - mx), ex2_ftz(s3.x - mx)};  // ex2(-inf) = 0
        const float l = f[0] * s0.y + f[1] * s1.y + f[2] * s2.y + f[3] * s3.y;
        const float inv = 1.f / l;
        if (kq < 3) {  // thread = (query row, CPT of the head's output features)
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
            if constexpr (CPT == 8) {
                uint32_t o[4][8];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) tmem_ld8(tmem + lane_off + o_col(b, kk) + 8 * kq, o[kk]);
                tmem_ld_wait();
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = fmaf(f[kk], __uint_as_float(o[kk][j]), acc[j]);
            } else {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint32_t o[16];
                    tmem_ld16(tmem + lane_off + o_col(b, kk) + 16 * kq, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[j] = fmaf(f[kk], __uint_as_float(o[j]), acc[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] *= inv;
#pragma unroll
            for (int cc = 0; cc < CPT / 8; ++cc) {
                *reinterpret_cast<uint4*>(ao + (KCH * h + (CPT / 8) * kq + cc) * kCS + tq * 16) = pack8<FMT>(acc + 8 * cc);
                if (a.save_o && tq < T)
                    *reinterpret_cast<uint4*>(a.save_o + tile_off(slab, 12, T, KCH * h + (CPT / 8) * kq + cc, tq)) = pack8<FMT_F16>(acc + 8 * cc);
            }
        } else if (a.save_lse && tq < T) {
            a.save_lse[((size_t)slab * NHEADS + h) * T + tq] = mx + log2f(l);
        }
    };

    stagger_start(59000);  // cycles per work item (profiles/r02e_phases.txt)
    int it_ = 0;
    const int nch = LONG ? (T + 255) / 256 : 1;  // work items per slab
    for (int item = blockIdx.x; item < a.nslab * nch; item += gridDim.x, ++it_) {
        const int slab = LONG ? item / nch : item;
        if constexpr (LONG != 0) { t0 = 256 * (item % nch); Tc = min(256, T - t0); }
        const float* xs = a.x + ((size_t)slab * T + t0) * kH;
        NBSS_TICK(0, 0, it_);
        if (tid == 0) load_image(wr, a.img + (LONG == 2 ? IMG_WQ : IMG_WKV), LONG == 2 ? IMG_WQ_BYTES : IMG_W1_BYTES, bar_w);
        if constexpr (LONG == 2) {
            // rows of the V tile sets beyond the slab's last (partial) key block must be finite: P = 0 there, but 0 x NaN is NaN, and
            // the previous item's fp32 out-proj staging left arbitrary bits.  (Full blocks overwrite them with real values later.)
            const int tl = T - 256 * ((T - 1) / 256);
            for (int i = tid; i < 8 * (256 - tl); i += kMhThreads) {
                const int c = i / (256 - tl), r = tl + i % (256 - tl);
                *reinterpret_cast<uint4*>(vt + (size_t)c * kCS + r * 16) = make_uint4(0, 0, 0, 0);
            }
        }
        stage_rows96<FMT, true>(xs, Tc, ao, 0, s_lng, s_lnb, warp, lane, (LONG == 0 && a.ln_stats) ? a.ln_stats + (size_t)slab * T * 2 : nullptr, kMhThreads / 32);
        end_epilogue();
        NBSS_TICK(0, 1, it_);
        // ---- P1: K|V
        if constexpr (LONG != 2) {
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
        if (LONG == 0 && tid == 0) load_image(wr, a.img + IMG_WQ, IMG_WQ_BYTES, bar_w);
        {
            const bool valid = t < Tc;
            const uint32_t tacc = tmem + lane_off + m * 192;
            // K: cols 0..95 -> per-head chunks HS*h .. HS*h+KCH-1   (channel half 0 does K, half 1 does V)
#pragma unroll 1
            for (int c3 = 0; c3 < (hf == 0 ? 12 : 0); c3 += 3) {  // three compact chunks at a time (never straddles a head)
                const int h = c3 / KCH, k0 = c3 % KCH;
                uint32_t r[24];
#pragma unroll
                for (int k = 0; k < 3; ++k) tmem_ld8(tacc + 8 * (c3 + k), *reinterpret_cast<uint32_t(*)[8]>(r + 8 * k));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = valid ? __uint_as_float(r[8 * k + j]) + s_bin[96 + 8 * (c3 + k) + j] : 0.f;
                    *reinterpret_cast<uint4*>(kt + (HS * h + k0 + k) * kCS + t * 16) = pack8<FMT>(v);
                    if (a.save_qkv && valid) *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, 12 + c3 + k, t0 + t)) = pack8<FMT_F16>(v);
                }
            }
            // V: cols 96..191 -> compact chunks 0..11
#pragma unroll 1
            for (int c = 0; c < (hf == 1 ? 12 : 0); c += 3) {
                uint32_t r[24];
#pragma unroll
                for (int k = 0; k < 3; ++k) tmem_ld8(tacc + 96 + 8 * (c + k), *reinterpret_cast<uint32_t(*)[8]>(r + 8 * k));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = valid ? __uint_as_float(r[8 * k + j]) + s_bin[192 + 8 * (c + k) + j] : 0.f;
                    *reinterpret_cast<uint4*>(vt + (c + k) * kCS + t * 16) = pack8<FMT>(v);
                    if (a.save_qkv && valid) *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, 24 + c + k, t0 + t)) = pack8<FMT_F16>(v);
                }
            }
        }
        end_epilogue();
        NBSS_TICK(0, 3, it_);
        if constexpr (LONG == 1) continue;  // K | V pass: TMEM and the tiles are reused by the next item
        }  // LONG != 2
        // ---- P2: Q
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
        if (tid == 0) load_image(wr, a.img + IMG_WO, IMG_WQ_BYTES, bar_w);  // out-proj image: needed only after the last head
        // next slab's input rows -> L2, issued HERE (not at the top of the slab, where it would compete with this slab's
        // latency-exposed staging loads): the attention heads below need no HBM traffic at all
        if (LONG == 0 && tid >= 32 && tid < 38 && slab + (int)gridDim.x < a.nslab) l2_prefetch_slab(a.x + (size_t)(slab + gridDim.x) * T * kH, T, tid - 32);
        // ---- EQ: scaled queries of all heads -> Q tile (over A0, dead now); thread = (frame, channel half)
        {
            const bool valid = t < Tc;
            const uint32_t tacc = tmem + lane_off + m * 96;
#pragma unroll 1
            for (int c0 = 48 * hf; c0 < 48 * hf + 48; c0 += 24) {
                uint32_t r[24];
#pragma unroll
                for (int k = 0; k < 3; ++k) tmem_ld8(tacc + c0 + 8 * k, *reinterpret_cast<uint32_t(*)[8]>(r + 8 * k));
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = valid ? (__uint_as_float(r[8 * k + j]) + s_bin[c0 + 8 * k + j]) * qscale : 0.f;
                    *reinterpret_cast<uint4*>(ao + (c0 / 8 + k) * kCS + t * 16) = pack8<FMT>(v);
                    if (LONG == 0 && a.save_qkv && valid) *reinterpret_cast<uint4*>(a.save_qkv + tile_off(slab, 36, T, c0 / 8 + k, t)) = pack8<FMT_F16>(v);
                }
            }
        }
        end_epilogue();
        NBSS_TICK(0, 8, it_);
        // ---- heads
        if constexpr (LONG == 2) {
            const int nkb = nch, nsteps = NHEADS * nkb;
            // ONE thread: K_h, V_h of key block kb (three chunk columns each) -> tile set `set` (K: chunks 4 set .. +2, the 4th stays the
            // zero pad of the S MMA's second k-step; V: chunks 4 set .. +2)
            auto load_kv = [&](int n, int set) {
                const int h = n / nkb, kb = n % nkb, tk = min(256, T - 256 * kb);
                mbar_expect_tx(bar_kv + set, (uint32_t)(6 * tk * 16));
                for (int c = 0; c < 3; ++c) {
                    bulk_g2s(kt + (size_t)(4 * set + c) * kCS, a.save_qkv + tile_off(slab, 36, T, 12 + 3 * h + c, 256 * kb), (uint32_t)(tk * 16), bar_kv + set);
                    bulk_g2s(vt + (size_t)(4 * set + c) * kCS, a.save_qkv + tile_off(slab, 36, T, 24 + 3 * h + c, 256 * kb), (uint32_t)(tk * 16), bar_kv + set);
                }
            };
            // warp 0: S of query tile b against the key block in tile set `set`
            auto issue_s_l = [&](int h, int b, int set) {
                tc_fence_after();
                const bool leader = elect_one();
                mma_kk(tmem + s_col(b), aoa + KCH * h * kCS + 128 * b * 16, kCS, kta + 4 * set * kCS, kCS, KKS, id256, 0, leader);
                if (leader) umma_commit(bar_s + b);
                __syncwarp();
            };
            auto issue_pv_l = [&](int set, int b) {
                tc_fence_after();
                const bool leader = elect_one();
#pragma unroll 1
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll 1
                    for (int ks = 0; ks < 4; ++ks)
                        if (leader)
                            umma_f16_ts(tmem + o_col(b, kk), tmem + s_col(b) + 64 * kk + 8 * ks,
                                        sdesc_mnmajor(vta + 4 * set * kCS + (64 * kk + 16 * ks) * 16, kCS), idpv, ks ? 1u : 0u);
                if (leader) umma_commit(bar_pv + b);
                __syncwarp();
            };
            if (tid == 0) {
                load_kv(0, 0);
                if (nsteps > 1) load_kv(1, 1);
            }
            if (warp == 0) {
                mbar_wait(bar_kv, ph_kv & 1u, a.err);
                ph_kv ^= 1u;
                issue_s_l(0, 0, 0);
                issue_s_l(0, 1, 0);
            }
            // running (max, sum, output) of the two query tiles; the read-out threads (kq < 3) own 8 output features each
            float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f}, acc_run[2][8];
#pragma unroll 1
            for (int n = 0; n < nsteps; ++n) {
                const int h = n / nkb, kb = n - h * nkb, set = n & 1, tk = min(256, T - 256 * kb);
                if (kb == 0) {
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        m_run[b] = -INFINITY;
                        l_run[b] = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc_run[b][j] = 0.f;
                    }
                }
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(bar_s + b, (ph_s >> b) & 1u, a.err);
                    ph_s ^= 1u << b;
                    tc_fence_after();
                    softmax_local(b, tk);
                    tc_fence_before();
                    __syncthreads();
                    if (warp == 0) issue_pv_l(set, b);
                }
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(bar_pv + b, (ph_pv >> b) & 1u, a.err);
                    ph_pv ^= 1u << b;
                    tc_fence_after();
                    {
                        const float2 s0 = stat[(b * 4 + 0) * 128 + rt], s1 = stat[(b * 4 + 1) * 128 + rt], s2 = stat[(b * 4 + 2) * 128 + rt],
                                     s3 = stat[(b * 4 + 3) * 128 + rt];
                        const float mx = fmaxf(fmaxf(s0.x, s1.x), fmaxf(s2.x, s3.x));  // finite: key quarter 0 holds the block's first key
                        const float mn = fmaxf(m_run[b], mx);
                        const float fr = ex2_ftz(m_run[b] - mn);                       // 0 on the first block (m_run = -inf)
                        const float f[4] = {ex2_ftz(s0.x - mn), ex2_ftz(s1.x - mn), ex2_ftz(s2.x - mn), ex2_ftz(s3.x - mn)};
                        l_run[b] = l_run[b] * fr + (f[0] * s0.y + f[1] * s1.y + f[2] * s2.y + f[3] * s3.y);
                        m_run[b] = mn;
                        if (kq < 3) {
                            uint32_t o[4][8];
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) tmem_ld8(tmem + lane_off + o_col(b, kk) + 8 * kq, o[kk]);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float v = acc_run[b][j] * fr;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) v = fmaf(f[kk], __uint_as_float(o[kk][j]), v);
                                acc_run[b][j] = v;
                            }
                            if (kb == nkb - 1) {  // the head's last key block: O_h(tile b) over Q_h(tile b), whose scores are all done
                                const float inv = 1.f / l_run[b];
                                float ov[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) ov[j] = acc_run[b][j] * inv;
                                *reinterpret_cast<uint4*>(ao + (KCH * h + kq) * kCS + (128 * b + rt) * 16) = pack8<FMT>(ov);
                            }
                        }
                    }
                    fence_async_smem();
                    tc_fence_before();
                    __syncthreads();
                    if (warp == 0 && n + 1 < nsteps) {
                        if (b == 0) {  // the next step's K_h / V_h block (requested one step ago)
                            mbar_wait(bar_kv + ((n + 1) & 1), (ph_kv >> ((n + 1) & 1)) & 1u, a.err);
                            ph_kv ^= 1u << ((n + 1) & 1);
                        }
                        issue_s_l((n + 1) / nkb, b, (n + 1) & 1);
                    }
                }
                // every MMA that read tile set n & 1 has completed (both bar_pv waits above): request step n + 2 into it
                if (tid == 0 && n + 2 < nsteps) load_kv(n + 2, set);
            }
        } else if constexpr (DBUF) {
            // the two query tiles of a head ping-pong between the two score buffers
            if (warp == 0) { issue_s(0, 0, 0); issue_s(0, 1, 1); }
#pragma unroll 1
            for (int h = 0; h < NHEADS; ++h) {
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(bar_s + b, (ph_s >> b) & 1u, a.err);
                    ph_s ^= 1u << b;
                    tc_fence_after();
                    softmax_local(b, T);
                    tc_fence_before();
                    __syncthreads();
                    if (warp == 0) issue_pv(h, b);
                    NBSS_TICK(0, 9 + 10 * h + 2 * b, it_);
                }
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    mbar_wait(bar_pv + b, (ph_pv >> b) & 1u, a.err);
                    ph_pv ^= 1u << b;
                    tc_fence_after();
                    readout(h, b, b, slab);
                    fence_async_smem();  // O -> Q tile is read by the out-proj MMA
                    tc_fence_before();
                    __syncthreads();
                    if (warp == 0 && h + 1 < NHEADS) issue_s(h + 1, b, b);  // buffer b is free again
                    NBSS_TICK(0, 10 + 10 * h + 2 * b + 3, it_);
                }
            }
        } else {
            // head dim 48: one score buffer; S -> softmax -> PV -> read-out in sequence (4 rounds per slab)
#pragma unroll 1
            for (int hm = 0; hm < 2 * NHEADS; ++hm) {
                const int h = hm >> 1, mq = hm & 1;
                if (warp == 0) issue_s(h, mq, 0);
                mbar_wait(bar_s, ph_s & 1u, a.err);
                ph_s ^= 1u;
                tc_fence_after();
                softmax_local(0, T);
                tc_fence_before();
                __syncthreads();
                if (warp == 0) issue_pv(h, 0);
                mbar_wait(bar_pv, ph_pv & 1u, a.err);
                ph_pv ^= 1u;
                tc_fence_after();
                readout(h, mq, 0, slab);
                fence_async_smem();
                tc_fence_before();
                __syncthreads();
            }
        }
        // ---- P3: out-proj + residual
        if (warp == 0) {
            tc_fence_after();
            mbar_wait(bar_w, ph_w, a.err);
            const bool leader = elect_one();
            for (int mm = 0; mm < 2; ++mm) mma_kk(tmem + mm * 96, aoa + 128 * mm * 16, kCS, wra, 96 * 16, 6, id96, 0, leader);
            if (leader) umma_commit(bar_mma);
        }
        ph_w ^= 1;
        wait_mma();
        NBSS_TICK(0, 5, it_);
        {
            // thread = (frame, channel half): D + b_out -> fp32, staged into the dead K|V tiles at the frame's row slot (24
            // four-float chunks; with per-head padding (HS != KCH) the zero pad chunks of K stay untouched: slab.cuh
            // skip4_chunk); then eight lanes per frame add the residual with coalesced traffic
            const uint32_t tacc = tmem + lane_off + m * 96;
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
                    *reinterpret_cast<float4*>(kt + (size_t)(HS != KCH ? skip4_chunk(c0 / 4 + j4) : c0 / 4 + j4) * kCS + t * 16) = o;
                }
            }
            tc_fence_before();
            __syncthreads();
            NBSS_TICK(0, 6, it_);
            add_rows<(HS != KCH)>(kt, kCS, 0, xs, a.y + ((size_t)slab * T + t0) * kH, Tc, warp, lane, kMhThreads / 32,
                                  (LONG == 0 && a.row_part) ? a.row_part + (size_t)slab * T * 2 : nullptr);
        }
        tc_fence_before();
        __syncthreads();
        NBSS_TICK(0, 7, it_);
    }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace nbss

NBSS_PHASE_READER(nbss_debug_phases_mhsa_fwd)

// num_heads = 4 (SpatialNet-small) or 2 (NBC2 small, models/arch/NBC2.py:294-311); row_part (nullable): per-row (sum, sum of
// squares) of the output over the 96 channels, the per-slab partials of NBC2's GroupBatchNorm (norm2, NBC2.py:170)
extern "C" int nbss_mhsa_fwd_nh(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b,
                                const float* b_in, const float* b_out, const void* layer_img, void* save_qkv, void* save_o,
                                float* save_lse, float* ln_stats, float* row_part, int num_heads, int fmt, int* err, void* stream) {
    using namespace nbss;
    if (!x || !y || !layer_img || !ln_w || !ln_b || !b_in || !b_out) return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    if ((fmt != FMT_F16 && fmt != FMT_BF16) || (num_heads != 4 && num_heads != 2)) return NBSS_ERR_UNSUPPORTED;
    MhsaFwdArgs a{x, y, nslab, T, ln_w, ln_b, b_in, b_out, (const unsigned char*)layer_img, (unsigned char*)save_qkv,
                  (unsigned char*)save_o, save_lse, ln_stats, row_part, err};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = nslab < sms ? nslab : sms;
    void (*kern)(MhsaFwdArgs) = num_heads == 4 ? ((fmt == FMT_F16) ? mhsa_fwd_kernel<FMT_F16, 4> : mhsa_fwd_kernel<FMT_BF16, 4>)
                                               : ((fmt == FMT_F16) ? mhsa_fwd_kernel<FMT_F16, 2> : mhsa_fwd_kernel<FMT_BF16, 2>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM);
    if (e != cudaSuccess) return (int)e;
    kern<<<grid, kMhThreads, MH_SMEM, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// T > 256, inference only (no saves): pass 1 writes k | v of all frames to kv_ws (fp16 [nslab][36][T][8], 576 T bytes per slab), pass 2
// attends chunk by chunk (see the header).  T <= 65536.
extern "C" int nbss_mhsa_fwd_long(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b_in,
                                  const float* b_out, const void* layer_img, void* kv_ws, int fmt, int* err, void* stream) {
    using namespace nbss;
    if (!x || !y || !layer_img || !ln_w || !ln_b || !b_in || !b_out || !kv_ws) return NBSS_ERR_NULL;
    if (T <= kTMax || T > 65536 || nslab < 1) return NBSS_ERR_SHAPE;
    if (fmt != FMT_F16) return NBSS_ERR_UNSUPPORTED;
    MhsaFwdArgs a{x, y, nslab, T, ln_w, ln_b, b_in, b_out, (const unsigned char*)layer_img, (unsigned char*)kv_ws, nullptr, nullptr, nullptr, nullptr, err};
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long items = (long long)nslab * ((T + 255) / 256);
    const int grid = items < sms ? (int)items : sms;
    void (*k1)(MhsaFwdArgs) = mhsa_fwd_kernel<FMT_F16, 4, 1>;
    void (*k2)(MhsaFwdArgs) = mhsa_fwd_kernel<FMT_F16, 4, 2>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MH_SMEM);
    if (e != cudaSuccess) return (int)e;
    k1<<<grid, kMhThreads, MH_SMEM, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    k2<<<grid, kMhThreads, MH_SMEM, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

extern "C" int nbss_mhsa_fwd(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b,
                             const float* b_in, const float* b_out, const void* layer_img, void* save_qkv, void* save_o,
                             float* save_lse, float* ln_stats, int fmt, int* err, void* stream) {
    return nbss_mhsa_fwd_nh(x, y, nslab, T, ln_w, ln_b, b_in, b_out, layer_img, save_qkv, save_o, save_lse, ln_stats, nullptr, 4, fmt, err, stream);
}
