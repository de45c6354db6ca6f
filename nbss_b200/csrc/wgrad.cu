// wgrad.cu — weight gradients of the narrow-band block on tensor cores (tcgen05, MN-major operands).
//
// dW[o, i] = sum over T-F points r of G[r, o] * ACT[r, i]: the contraction runs over ROWS, so both operands are the
// MN-major view of the same chunk-column smem tiles the forward kernels use (umma.cuh).  One persistent CTA walks
// its (b,f) slabs, stages G (gradient wrt a pre-activation, written by ffn_bwd / mhsa_bwd, or the fp32 upstream
// gradient) and ACT (activation operand, or LN(x) recomputed on the fly), and keeps accumulating into the same
// TMEM columns; there is ONE read-out per CTA at the end (fp32 atomics into the gradient buffers).
//   * conv taps: row-shifted view of ACT (start address + 16 B per frame)
//   * bias gradients: two extra ACT chunks filled with ones -> 16 extra accumulator columns (column sums of G)
//   * grouped conv weights: pairs of groups share one 96-wide MMA; the read-out keeps the diagonal 24x24 blocks
// A "job" (host-built descriptor) names the operands, the MMAs per slab and the read-out; several jobs run side by
// side in one launch (blockIdx.y).
#include "slab.cuh"

namespace nbss {

struct WgMma { int a_chunk, b_chunk, n, col, b_row; };
struct WgOut { float* dst; int col, lane0, nl, nc, mode, ld, tap; };  // mode 0 dense, 1 conv block-diagonal, 2 bias
struct WgJob {
    const void* g;     // gradient operand source
    const void* act;   // activation operand source
    const float *ln_w, *ln_b;
    int g_fp32, g_cols, g_c0, g_chunks, g_alloc;  // g_fp32: fp32 [n,96]; else 16-bit [n,g_cols], features [g_c0, g_c0+8*g_chunks)
    int a_ln, a_cols, a_c0, a_chunks, a_row_off;  // a_ln: fp32 x [n,96] through LayerNorm; ones chunks at [a_chunks, a_chunks+2)
    int dbl;                                      // both operands arrive by TMA: two tile sets, loads of slab i+1 overlap MMAs of slab i
    int a_silu;                                   // the activation operand is RECOMPUTED from the forward's saved fp16 pre-activation (act points at
                                                  // it): 1 = SiLU(.), 2 = SiLU(GroupNorm(.)); the tile is transformed in place after its TMA copy
    const float *gn_stats, *gn_w, *gn_b;          // a_silu == 2: [nslab][8][2] (mean, rstd), affine [192]
    int nmma, nout;
    WgMma mma[6];
    WgOut out[8];
};
struct WgArgs {
    WgJob job[3];
    int nslab, T;
    int* err;
};

constexpr int kWgThreads = 512;  // 16 warps: the fp32 operands of the pointwise jobs are staged by all of them

// In-place transform of a landed activation tile: saved pre-activation (fp16) -> s = SiLU(.) resp. SiLU(GroupNorm(.)) in FMT_A, with
// exactly the arithmetic ffn_bwd uses for its gradient (same sigmoid, same rounding points).  ffn_bwd therefore does not write the
// s1..s4 operand tensors any more (1.6 GB less HBM traffic per launch at batch 32).  Threads tidx of nthr walk 16-byte pieces, rows
// fastest (conflict-free shared-memory accesses).
template <int FMT_A>
__device__ __forceinline__ void wg_silu_tile(unsigned char* at, const WgJob& J, int slab, int T, int tidx, int nthr) {
    const int total = J.a_chunks << 8;  // pieces indexed (chunk, row) with 256 row slots per chunk: no division
    for (int i = tidx; i < total; i += nthr) {
        const int c = i >> 8, r = i & 255;
        if (r >= T) continue;
        uint4* p = reinterpret_cast<uint4*>(at + (size_t)c * kCS + (size_t)(J.a_row_off + r) * 16);
        const uint4 q = *p;
        float v[8];
        unpack_f16x2(q.x, v[0], v[1]);
        unpack_f16x2(q.y, v[2], v[3]);
        unpack_f16x2(q.z, v[4], v[5]);
        unpack_f16x2(q.w, v[6], v[7]);
        if (J.a_silu == 2) {
            const int ch = J.a_c0 + 8 * c, g = ch / kGC;  // a 16-byte piece never straddles a group (24 = 3 x 8)
            const float2 st = __ldg(reinterpret_cast<const float2*>(J.gn_stats + (size_t)slab * 16) + g);  // (mean, rstd)
            const float4 w0 = __ldg(reinterpret_cast<const float4*>(J.gn_w + ch)), w1 = __ldg(reinterpret_cast<const float4*>(J.gn_w + ch) + 1);
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(J.gn_b + ch)), b1 = __ldg(reinterpret_cast<const float4*>(J.gn_b + ch) + 1);
            const float gw[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, gb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf((v[j] - st.x) * st.y, gw[j], gb[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = v[j] * sigmoidf_(v[j]);
        *p = pack8<FMT_A>(v);
    }
}

template <int FMT_G, int FMT_A>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_kernel(WgArgs args) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar_mma, bar_ld;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_ln[192];
    const WgJob& J = args.job[blockIdx.y];
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31, T = args.T;
    unsigned char* gt = smem;
    unsigned char* at = smem + (size_t)J.g_alloc * kCS;
    const int a_alloc = J.a_chunks + 2;
    const int n_bulk = (J.g_fp32 ? 0 : 1) + (J.a_ln ? 0 : 1);  // operands arriving by TMA bulk copy (one arrive each)
    const uint32_t set_bytes = (uint32_t)(J.g_alloc + a_alloc) * kCS;
    const int nsets = J.dbl ? 2 : 1;
    __shared__ uint64_t bar_ld2;  // load barrier of the second tile set
    __shared__ int s_nslabs;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        mbar_init(&bar_ld, n_bulk > 0 ? n_bulk : 1);
        mbar_init(&bar_ld2, 2);
        fence_mbar_init();
    }
    if (J.a_ln) for (int i = tid; i < 96; i += kWgThreads) { s_ln[i] = J.ln_w[i]; s_ln[96 + i] = J.ln_b[i]; }
    for (int set = 0; set < nsets; ++set) {
        unsigned char* base = smem + (size_t)set * set_bytes;
        for (int i = tid; i < (int)((J.g_alloc + J.a_chunks) * kCS / 16); i += kWgThreads) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
        const uint32_t one2 = pack16<FMT_A>(1.f, 1.f);
        uint4* ones = reinterpret_cast<uint4*>(base + (size_t)(J.g_alloc + J.a_chunks) * kCS);
        for (int i = tid; i < (int)(2 * kCS / 16); i += kWgThreads) ones[i] = make_uint4(one2, one2, one2, one2);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t gta = smem_u32(gt), ata = smem_u32(at);
    uint32_t ph = 0, ph_ld = 0;
    bool first = true;
    (void)a_alloc;

    // called by every lane of warp 0 (uniform descriptor arithmetic); `leader` = the elected lane issues
    auto issue_mmas = [&](uint32_t ga, uint32_t aa, bool fresh, bool leader) {
        for (int i = 0; i < J.nmma; ++i) {
            const WgMma mm = J.mma[i];
            const uint32_t idesc = (1u << 4) | ((uint32_t)FMT_G << 7) | ((uint32_t)FMT_A << 10) | (1u << 15) | (1u << 16) |
                                   ((uint32_t)(mm.n >> 3) << 17) | (8u << 24);
            uint64_t da = sdesc_mnmajor(ga + mm.a_chunk * kCS, kCS), db = sdesc_mnmajor(aa + mm.b_chunk * kCS + mm.b_row * 16, kCS);
            for (int ks = 0; ks < 16; ++ks) {
                if (leader) umma_f16(tmem + mm.col, da, db, idesc, (fresh && ks == 0) ? 0u : 1u);
                da += 16;  // 16 rows x 16 B >> 4
                db += 16;
            }
        }
    };
    if (J.dbl && J.a_silu) {
        // ---- pipelined path with a recomputed activation operand: warp 0 drives loads and MMAs as below; the other 15 warps turn the
        //      landed pre-activation tile into the operand while the previous slab's MMAs and the next slab's copies are in flight
        const bool leader = warp == 0 && elect_one();
        auto issue_loads = [&](int slab, int set) {
            unsigned char* g0 = smem + (size_t)set * set_bytes;
            uint64_t* bar = set ? &bar_ld2 : &bar_ld;
            bulk_load_chunks(g0, kCS, 0, reinterpret_cast<const unsigned char*>(J.g) + tile_off(slab, J.g_cols / 8, T, J.g_c0 / 8, 0),
                             J.g_chunks, T, bar);
            bulk_load_chunks(g0 + (size_t)J.g_alloc * kCS, kCS, J.a_row_off,
                             reinterpret_cast<const unsigned char*>(J.act) + tile_off(slab, J.a_cols / 8, T, J.a_c0 / 8, 0), J.a_chunks, T, bar);
        };
        uint32_t phl = 0;  // bit s: phase of tile set s's load barrier
        int k = 0;
        if (leader && (int)blockIdx.x < args.nslab) issue_loads(blockIdx.x, 0);
        for (int slab = blockIdx.x; slab < args.nslab; slab += gridDim.x, ++k) {
            const int set = k & 1, nxt = slab + gridDim.x;
            if (warp == 0) {
                if (k >= 1) {  // MMAs of the previous slab (other tile set) are done: its tiles may be overwritten
                    mbar_wait(&bar_mma, ph, args.err);
                    ph ^= 1;
                }
                if (leader && nxt < args.nslab) issue_loads(nxt, set ^ 1);
            }
            // ONE warp (warp 1: another scheduler than the issuing warp 0) polls the copy's mbarrier and releases the other transform
            // warps through a named hardware barrier, so that 15 spinning warps do not compete with warp 0's long serial TMA / MMA
            // issue sequences.  (Measured: the all-warp loop costs +0.05 ms per launch against the warp-0-only loop even with an EMPTY
            // transform, the arithmetic another +0.045 ms; polling by one warp instead of fifteen recovers 0.005 ms of it.)
            if (warp == 1) mbar_wait(set ? &bar_ld2 : &bar_ld, (phl >> set) & 1u, args.err);
            phl ^= 1u << set;
            if (warp != 0) {
                asm volatile("bar.sync 1, %0;" ::"n"(kWgThreads - 32) : "memory");
                wg_silu_tile<FMT_A>(smem + (size_t)set * set_bytes + (size_t)J.g_alloc * kCS, J, slab, T, tid - 32, kWgThreads - 32);
            }
            fence_async_smem();
            tc_fence_before();
            __syncthreads();
            if (warp == 0) {
                tc_fence_after();
                const uint32_t ga = gta + set * set_bytes;
                issue_mmas(ga, ga + J.g_alloc * kCS, k == 0, leader);
                if (leader) umma_commit(&bar_mma);
                __syncwarp();
            }
        }
        if (warp == 0 && k >= 1) mbar_wait(&bar_mma, ph, args.err);
        first = k == 0;
        __syncthreads();
        tc_fence_after();
    } else if (J.dbl) {
        // ---- pipelined path (both operands by TMA): warp 0 drives loads and MMAs (one elected lane issues), two tile sets
        if (warp == 0) {
            const bool leader = elect_one();
            auto issue_loads = [&](int slab, int set) {
                unsigned char* g0 = smem + (size_t)set * set_bytes;
                uint64_t* bar = set ? &bar_ld2 : &bar_ld;
                bulk_load_chunks(g0, kCS, 0, reinterpret_cast<const unsigned char*>(J.g) + tile_off(slab, J.g_cols / 8, T, J.g_c0 / 8, 0),
                                 J.g_chunks, T, bar);
                bulk_load_chunks(g0 + (size_t)J.g_alloc * kCS, kCS, J.a_row_off,
                                 reinterpret_cast<const unsigned char*>(J.act) + tile_off(slab, J.a_cols / 8, T, J.a_c0 / 8, 0), J.a_chunks, T, bar);
            };
            uint32_t phl[2] = {0, 0};
            int k = 0;
            if (leader && (int)blockIdx.x < args.nslab) issue_loads(blockIdx.x, 0);
            for (int slab = blockIdx.x; slab < args.nslab; slab += gridDim.x, ++k) {
                const int set = k & 1, nxt = slab + gridDim.x;
                if (k >= 1) {  // MMAs of the previous slab (other tile set) are done: its tiles may be overwritten
                    mbar_wait(&bar_mma, ph, args.err);
                    ph ^= 1;
                }
                if (leader && nxt < args.nslab) issue_loads(nxt, set ^ 1);
                mbar_wait(set ? &bar_ld2 : &bar_ld, phl[set], args.err);
                phl[set] ^= 1;
                tc_fence_after();
                const uint32_t ga = gta + set * set_bytes;
                issue_mmas(ga, ga + J.g_alloc * kCS, k == 0, leader);
                if (leader) umma_commit(&bar_mma);
                __syncwarp();
            }
            if (k >= 1) mbar_wait(&bar_mma, ph, args.err);
            if (leader) s_nslabs = k;
        }
        __syncthreads();
        first = s_nslabs == 0;
        tc_fence_after();
    } else
    for (int slab = blockIdx.x; slab < args.nslab; slab += gridDim.x) {
        const size_t row0 = (size_t)slab * T;
        // ---- stage G and ACT: 16-bit slab-tile tensors come in by TMA bulk copies (one per chunk column), fp32
        //      sources (upstream gradient, x through LayerNorm) are converted by the warps
        if (tid == 0) {
            if (!J.g_fp32)
                bulk_load_chunks(gt, kCS, 0, reinterpret_cast<const unsigned char*>(J.g) + tile_off(slab, J.g_cols / 8, T, J.g_c0 / 8, 0),
                                 J.g_chunks, T, &bar_ld);
            if (!J.a_ln)
                bulk_load_chunks(at, kCS, J.a_row_off, reinterpret_cast<const unsigned char*>(J.act) + tile_off(slab, J.a_cols / 8, T, J.a_c0 / 8, 0),
                                 J.a_chunks, T, &bar_ld);
        }
        if (J.g_fp32) stage_rows96<FMT_G, false>(reinterpret_cast<const float*>(J.g) + row0 * kH, T, gt, 0, nullptr, nullptr, warp, lane, nullptr, kWgThreads / 32);
        if (J.a_ln) stage_rows96<FMT_A, true>(reinterpret_cast<const float*>(J.act) + row0 * kH, T, at, J.a_row_off, s_ln, s_ln + 96, warp, lane, nullptr, kWgThreads / 32);
        if (J.a_silu) {  // every thread waits for the copies, then the pre-activation tile becomes the operand in place
            mbar_wait(&bar_ld, ph_ld, args.err);
            ph_ld ^= 1;
            wg_silu_tile<FMT_A>(at, J, slab, T, tid, kWgThreads);
        }
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            if (n_bulk > 0 && !J.a_silu) { mbar_wait(&bar_ld, ph_ld, args.err); ph_ld ^= 1; }
            const bool leader = elect_one();
            issue_mmas(gta, ata, first, leader);
            if (leader) umma_commit(&bar_mma);
        }
        first = false;
        __syncwarp();
        mbar_wait(&bar_mma, ph, args.err);
        ph ^= 1;
        tc_fence_after();
    }
    // ---- read-out: warp w covers TMEM lanes 32*(w&3)..+31; warps 0-3 take even 8-column blocks, warps 4-7 odd ones
    if (!first) {
        const int L = 32 * (warp & 3) + lane;
        const uint32_t tl = tmem + ((uint32_t)(32 * (warp & 3)) << 16);
        for (int o = 0; o < J.nout; ++o) {
            const WgOut op = J.out[o];
            const int rl = L - op.lane0;
            const bool rowok = rl >= 0 && rl < op.nl;
            for (int cb = (warp >> 2); cb * 8 < op.nc; cb += kWgThreads / 128) {
                uint32_t r[8];
                tmem_ld8(tl + op.col + 8 * cb, r);
                tmem_ld_wait();
                if (!rowok) continue;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = 8 * cb + j;
                    if (c >= op.nc) continue;
                    const float v = __uint_as_float(r[j]);
                    if (op.mode == 0) atomicAdd(op.dst + (size_t)rl * op.ld + c, v);
                    else if (op.mode == 1) { if (rl / kGC == c / kGC) atomicAdd(op.dst + rl * 72 + (c % kGC) * 3 + op.tap, v); }
                    else if (c == 0) atomicAdd(op.dst + rl, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

static int wg_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return sms;
}

static int launch_wgrad(WgArgs& a, int njobs, int fmt_g, int fmt_a, cudaStream_t st) {
    int maxch = 0;
    for (int j = 0; j < njobs; ++j) {
        const int ch = (a.job[j].g_alloc + a.job[j].a_chunks + 2) * (a.job[j].dbl ? 2 : 1);
        maxch = ch > maxch ? ch : maxch;
    }
    const size_t smem = (size_t)maxch * kCS;
    if (smem > 227 * 1024 - 1024) return NBSS_ERR_UNSUPPORTED;
    void (*kern)(WgArgs) = nullptr;
    if (fmt_g == FMT_BF16 && fmt_a == FMT_BF16) kern = wgrad_kernel<FMT_BF16, FMT_BF16>;
    else if (fmt_g == FMT_BF16 && fmt_a == FMT_F16) kern = wgrad_kernel<FMT_BF16, FMT_F16>;
    else if (fmt_g == FMT_F16 && fmt_a == FMT_F16) kern = wgrad_kernel<FMT_F16, FMT_F16>;
    else return NBSS_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int per = wg_sms() / njobs;
    if (per < 1) per = 1;
    if (per > a.nslab) per = a.nslab;
    kern<<<dim3(per, njobs), kWgThreads, smem, st>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

static void add_mma(WgJob& j, int a_chunk, int b_chunk, int n, int col, int b_row) { j.mma[j.nmma++] = WgMma{a_chunk, b_chunk, n, col, b_row}; }
static void add_out(WgJob& j, float* dst, int col, int lane0, int nl, int nc, int mode, int ld, int tap) {
    j.out[j.nout++] = WgOut{dst, col, lane0, nl, nc, mode, ld, tap};
}

}  // namespace nbss

using namespace nbss;

// T-ConvFFN weight gradients.  g_*: 16-bit [n,192] gradient operands written by nbss_ffn_bwd (fmt_g); the activation operands
// s1 = SiLU(a1), s2 = SiLU(c1), s3 = SiLU(GroupNorm(c2)), s4 = SiLU(c3) are recomputed (fmt_a) from the forward's saved fp16
// pre-activations a1, c1, c2, c3 and the GroupNorm statistics / affine (wg_silu_tile).
// Accumulates into dW1 [192,96], db1, dWc{1,2,3} [192,24,3], dbc{1,2,3}, dW2 [96,192], db2 (all fp32, += semantics).
extern "C" int nbss_ffn_wgrad(const float* x, const float* dy, int nslab, int T, const float* ln_w, const float* ln_b,
                              const void* g_a1, const void* g_c1, const void* g_c2, const void* g_c3, const void* a1,
                              const void* c1, const void* c2, const void* c3, const float* gn_stats, const float* gn_w,
                              const float* gn_b, float* dW1, float* db1, float* dWc1,
                              float* dbc1, float* dWc2, float* dbc2, float* dWc3, float* dbc3, float* dW2, float* db2,
                              int fmt_g, int fmt_a, int* err, void* stream) {
    if (!x || !dy || !ln_w || !ln_b || !g_a1 || !g_c1 || !g_c2 || !g_c3 || !a1 || !c1 || !c2 || !c3 || !gn_stats || !gn_w || !gn_b ||
        !dW1 || !db1 || !dWc1 || !dbc1 || !dWc2 || !dbc2 || !dWc3 || !dbc3 || !dW2 || !db2)
        return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    // ---- launch 1: the two pointwise layers
    {
        WgArgs a{};
        a.nslab = nslab; a.T = T; a.err = err;
        // pw2: dW2[o<96, i<192] = dy^T s4
        WgJob& j = a.job[0];
        j.g = dy; j.g_fp32 = 1; j.g_alloc = 16;
        j.act = c3; j.a_silu = 1; j.a_cols = 192; j.a_c0 = 0; j.a_chunks = 24; j.a_row_off = 0;
        add_mma(j, 0, 0, 192, 0, 0);
        add_mma(j, 0, 24, 16, 192, 0);
        add_out(j, dW2, 0, 0, 96, 192, 0, 192, 0);
        add_out(j, db2, 192, 0, 96, 1, 2, 0, 0);
        // pw1: dW1[o<192, i<96] = g_a1^T LN(x); outputs 0..127 from the window at feature 0, 128..191 from the window at 64
        WgJob& k = a.job[1];
        k.g = g_a1; k.g_fp32 = 0; k.g_cols = 192; k.g_c0 = 0; k.g_chunks = 24; k.g_alloc = 24;
        k.act = x; k.a_ln = 1; k.ln_w = ln_w; k.ln_b = ln_b; k.a_chunks = 12; k.a_row_off = 0;
        add_mma(k, 0, 0, 96, 0, 0);
        add_mma(k, 0, 12, 16, 96, 0);
        add_mma(k, 8, 0, 96, 112, 0);
        add_mma(k, 8, 12, 16, 208, 0);
        add_out(k, dW1, 0, 0, 128, 96, 0, 96, 0);
        add_out(k, db1, 96, 0, 128, 1, 2, 0, 0);
        add_out(k, dW1 + 128 * 96, 112, 64, 64, 96, 0, 96, 0);
        add_out(k, db1 + 128, 208, 64, 64, 1, 2, 0, 0);
        int rc = launch_wgrad(a, 2, fmt_g, fmt_a, st);
        if (rc) return rc;
    }
    // ---- launches 2,3: the three grouped convs, two channel halves each (3 jobs per launch)
    const void* gs[3] = {g_c1, g_c2, g_c3};
    const void* as[3] = {a1, c1, c2};
    float* dWs[3] = {dWc1, dWc2, dWc3};
    float* dbs[3] = {dbc1, dbc2, dbc3};
    for (int half = 0; half < 2; ++half) {
        WgArgs a{};
        a.nslab = nslab; a.T = T; a.err = err;
        for (int c = 0; c < 3; ++c) {
            WgJob& j = a.job[c];
            j.g = gs[c]; j.g_fp32 = 0; j.g_cols = 192; j.g_c0 = 96 * half; j.g_chunks = 12; j.g_alloc = 12;  // M window rows 0..95
            j.dbl = 1;
            j.act = as[c]; j.a_cols = 192; j.a_c0 = 96 * half; j.a_chunks = 12; j.a_row_off = 1;
            j.a_silu = c == 2 ? 2 : 1;
            j.gn_stats = gn_stats; j.gn_w = gn_w; j.gn_b = gn_b;
            const int l0 = 0;  // TMEM lane of output channel 96*half (window starts at the half's first channel)
            for (int tap = 0; tap < 3; ++tap) {
                add_mma(j, 0, 0, 96, 96 * tap, tap);
                add_out(j, dWs[c] + (96 * half) * 72, 96 * tap, l0, 48, 48, 1, 0, tap);
                add_out(j, dWs[c] + (96 * half + 48) * 72, 96 * tap + 48, l0 + 48, 48, 48, 1, 0, tap);
            }
            add_mma(j, 0, 12, 16, 288, 1);
            add_out(j, dbs[c] + 96 * half, 288, l0, 96, 1, 2, 0, 0);
        }
        int rc = launch_wgrad(a, 3, fmt_g, fmt_a, st);
        if (rc) return rc;
    }
    return NBSS_OK;
}

// MHSA weight gradients: dWin [288,96] += dQKV^T LN(x), dbin += colsum(dQKV); dWo [96,96] += dy^T O, dbo += colsum(dy).
extern "C" int nbss_mhsa_wgrad(const float* x, const float* dy, int nslab, int T, const float* ln_w, const float* ln_b,
                               const void* dqkv, const void* o, float* dWin, float* dbin, float* dWo, float* dbo, int fmt_g,
                               int fmt_a, int* err, void* stream) {
    if (!x || !dy || !ln_w || !ln_b || !dqkv || !o || !dWin || !dbin || !dWo || !dbo) return NBSS_ERR_NULL;
    if (T < 1 || T > kTMax || nslab < 1) return NBSS_ERR_SHAPE;
    WgArgs a{};
    a.nslab = nslab; a.T = T; a.err = err;
    WgJob& j = a.job[0];  // out-proj
    j.g = dy; j.g_fp32 = 1; j.g_alloc = 16;
    j.act = o; j.a_cols = 96; j.a_c0 = 0; j.a_chunks = 12; j.a_row_off = 0;
    add_mma(j, 0, 0, 96, 0, 0);
    add_mma(j, 0, 12, 16, 96, 0);
    add_out(j, dWo, 0, 0, 96, 96, 0, 96, 0);
    add_out(j, dbo, 96, 0, 96, 1, 2, 0, 0);
    WgJob& k = a.job[1];  // in-proj: three 128-feature windows at 0, 128, 160
    k.g = dqkv; k.g_fp32 = 0; k.g_cols = 288; k.g_c0 = 0; k.g_chunks = 36; k.g_alloc = 36;
    k.act = x; k.a_ln = 1; k.ln_w = ln_w; k.ln_b = ln_b; k.a_chunks = 12; k.a_row_off = 0;
    const int win[3] = {0, 16, 20}, l0[3] = {0, 0, 96}, nl[3] = {128, 128, 32}, orow[3] = {0, 128, 256};
    for (int w = 0; w < 3; ++w) {
        add_mma(k, win[w], 0, 96, 112 * w, 0);
        add_mma(k, win[w], 12, 16, 112 * w + 96, 0);
        add_out(k, dWin + orow[w] * 96, 112 * w, l0[w], nl[w], 96, 0, 96, 0);
        add_out(k, dbin + orow[w], 112 * w + 96, l0[w], nl[w], 1, 2, 0, 0);
    }
    return launch_wgrad(a, 2, fmt_g, fmt_a, (cudaStream_t)stream);
}
