// umma_selftest.cu — one-CTA tcgen05 GEMM used by tests/test_umma_selftest.py to pin down every descriptor /
// layout convention of umma.cuh against a torch matmul: K-major and MN-major views of the chunk-column smem
// tile, row-shifted views (conv taps), bf16 / fp16 / tf32 kinds, N splits and accumulation.
#include "common.cuh"
#include "umma.cuh"

namespace nbss {

struct SelfTestArgs {
    const float* A;
    const float* B;
    float* D;
    int a_rows, a_feats, b_rows, b_feats;
    int N, Kdim;
    int a_mn, b_mn;      // 1 = MN-major view; a_mn = 2: the A operand is written to TENSOR MEMORY (columns 256..) by the threads
    int fmt;             // FMT_F16 / FMT_BF16 / FMT_TF32 (A operand)
    int fmt_b;           // B operand format (may differ from A for the 16-bit kinds)
    int a_shift, b_shift;  // row shift of the view (rows)
    int a_off, b_off;      // feature offset of the MN index (MN-major) or of the K index (K-major)
    int passes;            // issue the whole K loop `passes` times (accumulating)
    int tmem_col;          // accumulator column offset inside the allocation
    int* err;
};

__device__ __forceinline__ void fill_tile(unsigned char* dst, const float* src, int rows, int feats, int rtot, int fmt,
                                          int tid, int nthreads) {
    const int es = (fmt == FMT_TF32) ? 4 : 2;
    const int ce = 16 / es;
    const int nchunks = (feats + ce - 1) / ce;
    const int cs = rtot * 16;
    for (int idx = tid; idx < nchunks * rtot; idx += nthreads) {
        int c = idx / rtot, r = idx % rtot;
        unsigned char* p = dst + (size_t)c * cs + r * 16;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int f = c * ce + j;
            v[j] = (j < ce && r < rows && f < feats) ? src[(size_t)r * feats + f] : 0.f;
        }
        uint4 q;
        if (fmt == FMT_TF32) {
            q = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
        } else if (fmt == FMT_F16) {
            q = make_uint4(pack_f16(v[0], v[1]), pack_f16(v[2], v[3]), pack_f16(v[4], v[5]), pack_f16(v[6], v[7]));
        } else {
            q = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
        *reinterpret_cast<uint4*>(p) = q;
    }
}

__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(SelfTestArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int es = (a.fmt == FMT_TF32) ? 4 : 2;
    const int ce = 16 / es;
    const int a_rtot = a.a_rows + 8, b_rtot = a.b_rows + 8;
    const int a_cs = a_rtot * 16, b_cs = b_rtot * 16;
    const int a_chunks = (a.a_feats + ce - 1) / ce;
    unsigned char* sA = smem;
    unsigned char* sB = smem + (size_t)a_chunks * a_cs;

    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fill_tile(sA, a.A, a.a_rows, a.a_feats, a_rtot, a.fmt, tid, blockDim.x);
    fill_tile(sB, a.B, a.b_rows, a.b_feats, b_rtot, a.fmt_b, tid, blockDim.x);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tmem_base_s + a.tmem_col;
    const uint32_t ta_base = tmem_base_s + 256;  // A operand in tensor memory (a_mn == 2)
    if (a.a_mn == 2) {
        // thread = row: two 16-bit K elements per 32-bit column
        for (int c0 = 0; c0 < a.Kdim / 2; c0 += 8) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = 2 * (c0 + j);
                const float lo = (tid < a.a_rows && k < a.a_feats) ? a.A[(size_t)tid * a.a_feats + k] : 0.f;
                const float hi = (tid < a.a_rows && k + 1 < a.a_feats) ? a.A[(size_t)tid * a.a_feats + k + 1] : 0.f;
                v[j] = a.fmt == FMT_F16 ? pack_f16(lo, hi) : pack_bf16(lo, hi);
            }
            tmem_st8(tmem_addr(ta_base, (warp & 3) * 32, c0), v);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }

    if (tid == 0) {
        const uint32_t idesc = (make_idesc(a.fmt, 128, a.N, a.a_mn == 1, a.b_mn) & ~(7u << 10)) | ((uint32_t)a.fmt_b << 10);
        const int kstep = (a.fmt == FMT_TF32) ? 8 : 16;
        const int nk = a.Kdim / kstep;
        uint32_t acc = 0;
        for (int pass = 0; pass < a.passes; ++pass) {
            for (int k = 0; k < nk; ++k) {
                uint64_t da, db;
                if (a.a_mn != 1) {
                    uint32_t s = smem_u32(sA) + a.a_shift * 16 + (a.a_off / ce + 2 * k) * a_cs;
                    da = sdesc_kmajor(s, a_cs);
                } else {
                    uint32_t s = smem_u32(sA) + (a.a_off / ce) * a_cs + (a.a_shift + k * kstep) * 16;
                    da = sdesc_mnmajor(s, a_cs);
                }
                if (!a.b_mn) {
                    uint32_t s = smem_u32(sB) + a.b_shift * 16 + (a.b_off / ce + 2 * k) * b_cs;
                    db = sdesc_kmajor(s, b_cs);
                } else {
                    uint32_t s = smem_u32(sB) + (a.b_off / ce) * b_cs + (a.b_shift + k * kstep) * 16;
                    db = sdesc_mnmajor(s, b_cs);
                }
                if (a.a_mn == 2) umma_f16_ts(tbase, ta_base + 8 * k, db, idesc, acc);
                else if (a.fmt == FMT_TF32) umma_tf32(tbase, da, db, idesc, acc);
                else umma_f16(tbase, da, db, idesc, acc);
                acc = 1;
            }
        }
        umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0, a.err);
    tc_fence_after();

    // epilogue: thread = row
    const int row = tid;
    for (int c0 = 0; c0 < a.N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tmem_addr(tbase, (warp & 3) * 32, c0), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) a.D[(size_t)row * a.N + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

}  // namespace nbss

extern "C" int nbss_umma_selftest(const float* A, int a_rows, int a_feats, const float* B, int b_rows, int b_feats,
                                  float* D, int N, int Kdim, int a_mn, int b_mn, int fmt, int a_shift, int b_shift,
                                  int a_off, int b_off, int passes, int tmem_col, int fmt_b, int* err, void* stream) {
    using namespace nbss;
    if (N % 16 || N < 16 || N > 256) return NBSS_ERR_SHAPE;
    if (tmem_col + N > 512) return NBSS_ERR_SHAPE;
    if (a_mn == 2 && (tmem_col + N > 256 || Kdim / 2 > 256 || Kdim % 16 || fmt == FMT_TF32)) return NBSS_ERR_SHAPE;
    const int es = (fmt == FMT_TF32) ? 4 : 2, ce = 16 / es;
    size_t bytes = (size_t)((a_feats + ce - 1) / ce) * (a_rows + 8) * 16 + (size_t)((b_feats + ce - 1) / ce) * (b_rows + 8) * 16;
    if (bytes > 200 * 1024) return NBSS_ERR_SHAPE;
    SelfTestArgs a{A, B, D, a_rows, a_feats, b_rows, b_feats, N, Kdim, a_mn, b_mn, fmt, fmt_b < 0 ? fmt : fmt_b, a_shift, b_shift, a_off, b_off, passes, tmem_col, err};
    cudaError_t e = cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    umma_selftest_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(a);
    return (int)cudaGetLastError();
}
