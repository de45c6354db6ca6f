// fullband_tc.cu — the full-band LinearGroup of the cross-band block on tensor cores (tcgen05).
//
// Replaces LinearGroup.forward (models/arch/base/linear_group.py:29-34) as used by SpatialNetLayer._full
// (models/arch/SpatialNet.py:129-146): for every (b,t) row m and squeeze channel g,
//     u[m,g,k] = sum_f Wf[g,k,f] * s[m,g,f] + bf[g,k]                       (forward)
//     ds[m,g,f] = sum_k Wf[g,k,f] * du[m,g,k]                               (data gradient)
//     dWf[g,k,f] += sum_m du[m,g,k] * s[m,g,f],  dbf[g,k] += sum_m du[m,g,k] (weight gradient)
// s, u, du, ds are the fp32 [B*T, 8, F] tensors of crossband.cu.  One CTA = (128-row tile, group g):
//   forward / dgrad: the rows are staged as a 16-bit K-major operand tile (chunk-column layout, umma.cuh), the group's
//                    F x F weight is a pre-packed image brought in by one TMA bulk copy; the SAME image serves both
//                    directions (K-major view: N = k, K = f; MN-major view: N = f, K = k).  D [128 x Fp] fp32 in TMEM,
//                    staged through shared memory for coalesced row writes.
//   wgrad          : both operands MN-major (K = rows), D [k<128, f] accumulates in TMEM over the CTA's row tiles;
//                    d bias and the rows k >= 128 of d W (F = 129: one row) are summed on CUDA cores from the same tiles
//                    while the MMAs run.
// Supports F <= 256 (Fp = F rounded up to 16 <= 256 = the UMMA N limit).
#include <algorithm>

#include "slab.cuh"

namespace nbss {

constexpr int kLgRows = 128;
constexpr int kLgG = 8;  // dim_squeeze: groups of the LinearGroup
constexpr uint32_t kLgCsA = (kLgRows + 1) * 16;  // 2064: chunk stride of a row tile (16 mod 128: conflict-free staging)

struct LgGeom {
    int M, F, Fp, nch;   // rows, features, padded features, 8-feature chunks
    uint32_t csb;        // image chunk stride (Fp + 1) * 16
    uint32_t img_bytes;  // per group
};
__host__ __device__ inline LgGeom lg_geom(int M, int F) {
    LgGeom g;
    g.M = M; g.F = F; g.Fp = (F + 15) / 16 * 16; g.nch = g.Fp / 8;
    g.csb = (uint32_t)(g.Fp + 1) * 16;
    g.img_bytes = (uint32_t)g.nch * g.csb;
    return g;
}

// image of group g: chunk c, row n (16 bytes) = Wf[g][n][8c .. 8c+7]; zero for n >= F or f >= F
__global__ void lg_pack_kernel(const float* __restrict__ Wf, unsigned char* img, LgGeom g, int fmt) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t per_group = (size_t)g.nch * (g.Fp + 1);
    if (idx >= per_group * kLgG) return;
    const int grp = (int)(idx / per_group), rem = (int)(idx % per_group), c = rem / (g.Fp + 1), n = rem % (g.Fp + 1);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int f = 8 * c + j;
        v[j] = (n < g.F && f < g.F) ? Wf[((size_t)grp * g.F + n) * g.F + f] : 0.f;
    }
    const uint4 o = (fmt == FMT_F16) ? make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]))
                                     : make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(img + idx * 16) = o;
}

// rows [m0, m0+128) of group grp of an fp32 [M, 8, F] tensor -> 16-bit row tile (nch chunks); rows >= M and features >= F zero
template <int FMT>
__device__ __forceinline__ void lg_stage(const float* __restrict__ in, const LgGeom& g, int m0, int grp, unsigned char* tile,
                                         int tid, int nthreads) {
    for (int u = tid; u < kLgRows * g.nch; u += nthreads) {
        const int r = u / g.nch, c = u - r * g.nch, m = m0 + r;
        float v[8];
        const float* src = in + ((size_t)m * kLgG + grp) * g.F + 8 * c;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (m < g.M && 8 * c + j < g.F) ? __ldg(src + j) : 0.f;
        *reinterpret_cast<uint4*>(tile + (size_t)c * kLgCsA + r * 16) = pack8<FMT>(v);
    }
}

struct LgArgs {
    const float* in;   // [M,8,F]
    float* out;        // [M,8,F]
    LgGeom g;
    const unsigned char* img;
    const float* bias;  // [8,F] or null
    int* err;
};

// MODE 0: out[m,g,k] = sum_f in[m,g,f] W[g,k,f] + bias[g,k];  MODE 1: out[m,g,f] = sum_k in[m,g,k] W[g,k,f]
template <int FMT, int MODE>
__global__ void __launch_bounds__(256, 2) lg_tc_kernel(LgArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LgGeom g = a.g;
    unsigned char* tile = smem;
    unsigned char* wimg = smem + (size_t)g.nch * kLgCsA;
    const int FS = g.F | 1;  // fp32 staging row stride (odd: conflict-free thread-per-row writes); aliases tile + image
    float* stage = reinterpret_cast<float*>(smem);
    const size_t body = max((size_t)g.nch * kLgCsA + g.img_bytes, (size_t)kLgRows * FS * 4);
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + ((body + 15) / 16) * 16);
    uint64_t* bar_w = bar_mma + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 2);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const int m0 = blockIdx.x * kLgRows, grp = blockIdx.y;
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        mbar_init(bar_w, 1);
        fence_mbar_init();
        load_image(wimg, a.img + (size_t)grp * g.img_bytes, g.img_bytes, bar_w);
    }
    lg_stage<FMT>(a.in, g, m0, grp, tile, tid, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (warp == 0) {
        mbar_wait(bar_w, 0, a.err);
        const bool leader = elect_one();
        const uint32_t ta = smem_u32(tile), wa = smem_u32(wimg);
        if (MODE == 0) {
            mma_kk(tmem, ta, kLgCsA, wa, g.csb, g.Fp / 16, make_idesc(FMT, 128, g.Fp, 0, 0), 0, leader);
        } else {
            const uint32_t id = make_idesc(FMT, 128, g.Fp, 0, 1);
            for (int ks = 0; ks < g.Fp / 16; ++ks)
                if (leader) umma_f16(tmem, sdesc_kmajor(ta + (uint32_t)(2 * ks) * kLgCsA, kLgCsA), sdesc_mnmajor(wa + ks * 256, g.csb), id, ks ? 1u : 0u);
        }
        if (leader) umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, 0, a.err);
    tc_fence_after();
    // epilogue 1: thread = (row, column half): D (+ bias) -> fp32 staging (the operand tiles are dead)
    {
        const int q = warp & 3, hf = warp >> 2, r = 32 * q + lane;
        const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16);
        const float* bias = (MODE == 0 && a.bias) ? a.bias + (size_t)grp * g.F : nullptr;
#pragma unroll 1
        for (int k = hf * (g.nch / 2); k < (hf + 1) * (g.nch / 2); ++k) {
            uint32_t v[8];
            tmem_ld8(tacc + 8 * k, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = 8 * k + j;
                if (c < g.F) stage[r * FS + c] = __uint_as_float(v[j]) + (bias ? bias[c] : 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    // epilogue 2: warp per row, coalesced
    for (int r = warp; r < kLgRows; r += 8) {
        const int m = m0 + r;
        if (m >= g.M) break;
        float* dst = a.out + ((size_t)m * kLgG + grp) * g.F;
        for (int c = lane; c < g.F; c += 32) dst[c] = stage[r * FS + c];
    }
    if (warp == 0) tmem_dealloc(tmem, 256);
}

struct LgWgArgs {
    const float* du;  // [M,8,F]
    const float* s;   // [M,8,F]
    LgGeom g;
    float* dW;        // [8,F,F]
    float* db;        // [8,F]
    int* err;
};

template <int FMT>
__global__ void __launch_bounds__(256, 2) lg_wgrad_kernel(LgWgArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const LgGeom g = a.g;
    const int ach = g.nch < 16 ? 16 : g.nch;  // the 128-feature MN-major window of du reads 16 chunks
    unsigned char* dut = smem;
    unsigned char* st = smem + (size_t)ach * kLgCsA;
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(st + (size_t)g.nch * kLgCsA);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + 1);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0) /* warp-uniform for ptxas: see umma.cuh elect_one */, lane = tid & 31;
    const int grp = blockIdx.y;
    if (warp == 0) tmem_alloc(tmem_slot, 256);
    if (tid == 0) {
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < (int)((size_t)ach * kLgCsA / 16); i += 256) reinterpret_cast<uint4*>(dut)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t da = smem_u32(dut), sa = smem_u32(st);
    const uint32_t id = make_idesc(FMT, 128, g.Fp, 1, 1);
    const int ntiles = (g.M + kLgRows - 1) / kLgRows;
    uint32_t ph = 0;
    bool any = false;
    // CUDA-core side sums: thread f < F accumulates d b[k = f] and the rows k >= 128 of d W (column f)
    constexpr int kMaxHi = 4;  // rows 128..131 in registers; more (F > 132) are handled by the generic loop below
    float dbias = 0.f, dhi[kMaxHi] = {0.f, 0.f, 0.f, 0.f};
    const int nhi = g.F > 128 ? g.F - 128 : 0;
    for (int mt = blockIdx.x; mt < ntiles; mt += gridDim.x) {
        if (any) {  // the previous tile's MMAs must have consumed the operand tiles
            __syncwarp();
            mbar_wait(bar_mma, ph, a.err);
            ph ^= 1;
            tc_fence_after();
            __syncthreads();
        }
        lg_stage<FMT>(a.du, g, mt * kLgRows, grp, dut, tid, 256);
        lg_stage<FMT>(a.s, g, mt * kLgRows, grp, st, tid, 256);
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (warp == 0) {
            tc_fence_after();
            const bool leader = elect_one();
            for (int ks = 0; ks < kLgRows / 16; ++ks)
                if (leader) umma_f16(tmem, sdesc_mnmajor(da + ks * 256, kLgCsA), sdesc_mnmajor(sa + ks * 256, kLgCsA), id, (any || ks) ? 1u : 0u);
            if (leader) umma_commit(bar_mma);
        }
        any = true;
        // while the tensor core works: column sums of du (bias gradient) and the k >= 128 rows of dW
        if (tid < g.F) {
            const unsigned char* dcol = dut + (size_t)(tid >> 3) * kLgCsA + (tid & 7) * 2;
            const unsigned char* scol = st + (size_t)(tid >> 3) * kLgCsA + (tid & 7) * 2;
            if (nhi == 0) {
#pragma unroll 4
                for (int r = 0; r < kLgRows; ++r) {
                    float lo, hi;
                    unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(dcol + r * 16), lo, hi);
                    dbias += lo;
                }
            } else if (nhi <= kMaxHi) {
                const unsigned char* hcol = dut + (size_t)16 * kLgCsA;  // features 128..135 of row r: 16 bytes
#pragma unroll 2
                for (int r = 0; r < kLgRows; ++r) {
                    float lo, hi, sv;
                    unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(dcol + r * 16), lo, hi);
                    dbias += lo;
                    unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(scol + r * 16), sv, hi);
                    const uint2 hk = *reinterpret_cast<const uint2*>(hcol + r * 16);
                    float h0, h1, h2, h3;
                    unpack16<FMT>(hk.x, h0, h1);
                    unpack16<FMT>(hk.y, h2, h3);
                    dhi[0] = fmaf(h0, sv, dhi[0]);
                    dhi[1] = fmaf(h1, sv, dhi[1]);
                    dhi[2] = fmaf(h2, sv, dhi[2]);
                    dhi[3] = fmaf(h3, sv, dhi[3]);
                }
            } else {
                // generic (slow) path for 132 < F <= 256: rows k >= 128 one at a time, straight into global memory
#pragma unroll 1
                for (int r = 0; r < kLgRows; ++r) {
                    float lo, hi;
                    unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(dcol + r * 16), lo, hi);
                    dbias += lo;
                }
#pragma unroll 1
                for (int k = 128; k < g.F; ++k) {
                    const unsigned char* kc = dut + (size_t)(k >> 3) * kLgCsA + (k & 7) * 2;
                    float acc = 0.f;
                    for (int r = 0; r < kLgRows; ++r) {
                        float dv, sv, hi;
                        unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(kc + r * 16), dv, hi);
                        unpack16<FMT>((uint32_t)*reinterpret_cast<const unsigned short*>(scol + r * 16), sv, hi);
                        acc = fmaf(dv, sv, acc);
                    }
                    atomicAdd(a.dW + ((size_t)grp * g.F + k) * g.F + tid, acc);
                }
            }
        }
    }
    if (any) {
        __syncwarp();
        mbar_wait(bar_mma, ph, a.err);
        tc_fence_after();
        // D[k (lane), f (column)] -> dW[g][k][f]
        const int q = warp & 3, hf = warp >> 2, k = 32 * q + lane;
        const uint32_t tacc = tmem + ((uint32_t)(32 * q) << 16);
        float* drow = a.dW + ((size_t)grp * g.F + (k < g.F ? k : 0)) * g.F;
#pragma unroll 1
        for (int c8 = hf * (g.nch / 2); c8 < (hf + 1) * (g.nch / 2); ++c8) {
            uint32_t v[8];
            tmem_ld8(tacc + 8 * c8, v);
            tmem_ld_wait();
            if (k < g.F) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (8 * c8 + j < g.F) atomicAdd(drow + 8 * c8 + j, __uint_as_float(v[j]));
            }
        }
        if (tid < g.F) {
            atomicAdd(a.db + (size_t)grp * g.F + tid, dbias);
            if (nhi > 0 && nhi <= kMaxHi) {
#pragma unroll
                for (int i = 0; i < kMaxHi; ++i)
                    if (i < nhi) atomicAdd(a.dW + ((size_t)grp * g.F + 128 + i) * g.F + tid, dhi[i]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

static int lg_sms() {
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

extern "C" unsigned int nbss_lg_image_bytes(int F) {
    if (F < 1 || F > 256) return 0;
    return lg_geom(1, F).img_bytes * kLgG;
}

// Wf = full.weight [8,F,F] (LinearGroup) -> 8 UMMA operand images
extern "C" int nbss_lg_pack(const float* Wf, void* img, int F, int fmt, void* stream) {
    if (!Wf || !img) return NBSS_ERR_NULL;
    if (F < 1 || F > 256) return NBSS_ERR_UNSUPPORTED;
    if (fmt != FMT_F16 && fmt != FMT_BF16) return NBSS_ERR_UNSUPPORTED;
    const LgGeom g = lg_geom(1, F);
    const size_t n = (size_t)g.nch * (g.Fp + 1) * kLgG;
    lg_pack_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(Wf, (unsigned char*)img, g, fmt);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// mode 0: out = in * W^T + bias (forward); mode 1: out = in * W (data gradient).  in/out: fp32 [M,8,F].
extern "C" int nbss_lg_tc_apply(const float* in, float* out, int M, int F, const void* img, const float* bias, int mode,
                                int fmt, int* err, void* stream) {
    if (!in || !out || !img) return NBSS_ERR_NULL;
    if (M < 1 || F < 1) return NBSS_ERR_SHAPE;
    if (F > 256 || (fmt != FMT_F16 && fmt != FMT_BF16) || (mode != 0 && mode != 1)) return NBSS_ERR_UNSUPPORTED;
    LgArgs a{in, out, lg_geom(M, F), (const unsigned char*)img, bias, err};
    const size_t body = std::max((size_t)a.g.nch * kLgCsA + a.g.img_bytes, (size_t)kLgRows * (F | 1) * 4);
    const size_t smem = (body + 15) / 16 * 16 + 64;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    void (*kern)(LgArgs) = (fmt == FMT_F16) ? (mode == 0 ? lg_tc_kernel<FMT_F16, 0> : lg_tc_kernel<FMT_F16, 1>)
                                            : (mode == 0 ? lg_tc_kernel<FMT_BF16, 0> : lg_tc_kernel<FMT_BF16, 1>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    kern<<<dim3((M + kLgRows - 1) / kLgRows, kLgG), 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}

// dW [8,F,F] += du^T s per group, db [8,F] += column sums of du.  du, s: fp32 [M,8,F].
extern "C" int nbss_lg_tc_wgrad(const float* du, const float* s, int M, int F, float* dW, float* db, int fmt, int* err,
                                void* stream) {
    if (!du || !s || !dW || !db) return NBSS_ERR_NULL;
    if (M < 1 || F < 1) return NBSS_ERR_SHAPE;
    if (F > 256 || (fmt != FMT_F16 && fmt != FMT_BF16)) return NBSS_ERR_UNSUPPORTED;
    LgWgArgs a{du, s, lg_geom(M, F), dW, db, err};
    const int ach = a.g.nch < 16 ? 16 : a.g.nch;
    const size_t smem = (size_t)(ach + a.g.nch) * kLgCsA + 64;
    if (smem > 227 * 1024) return NBSS_ERR_UNSUPPORTED;
    auto kern = (fmt == FMT_F16) ? lg_wgrad_kernel<FMT_F16> : lg_wgrad_kernel<FMT_BF16>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int ntiles = (M + kLgRows - 1) / kLgRows;
    int splits = (2 * lg_sms() + kLgG - 1) / kLgG;  // ~2 CTAs per SM over the 8 groups
    if (splits > ntiles) splits = ntiles;
    kern<<<dim3(splits, kLgG), 256, smem, (cudaStream_t)stream>>>(a);
    NBSS_LAUNCH_CHECK();
    return NBSS_OK;
}
