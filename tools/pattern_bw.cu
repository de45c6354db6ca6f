// pattern_bw.cu — what HBM bandwidth do the two ACCESS PATTERNS of the [B,F,T,96] fp32 stream allow on this GPU, with no arithmetic at all?
//
//   slab pattern   (narrow-band kernels: mhsa_*, ffn_*):  item = (b,f): T x 96 floats, one contiguous 96 KB piece
//   frame pattern  (cross-band kernels: fconv_tc_*):      item = (b, t0..t0+1): for every f one 768-byte piece, pieces 96 KB apart
//   flat pattern   (reference point):                      grid-stride float4 copy over the whole tensor
//
// Every kernel does dx = x + dy (two reads, one write per element = the algorithmic traffic of a fused backward sub-block)
// with the row mapping of the product kernels (eight lanes per 384-byte row, U rows per lane group and pass: slab.cuh / fconv_tc.cu
// fc_stage), persistent CTAs of 512 threads, one item at a time per CTA, K CTAs per SM.  The numbers bound what ANY kernel with that
// access pattern and that much memory-level parallelism can reach; the product kernels add their serial MMA / epilogue phases on top.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o nbss_b200/csrc/build/pattern_bw tools/pattern_bw.cu
// run  : nbss_b200/csrc/build/pattern_bw [out.json]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

constexpr int kB = 32, kF = 129, kT = 250, kH = 96;

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// rows of one item: row r (0 <= r < nrows) lives at base + row_off(r) floats.  MODE 0: slab (rows contiguous), MODE 1: frame group.
template <int MODE>
__device__ __forceinline__ size_t row_offset(int item, int r) {
    if (MODE == 0) return ((size_t)item * kT + r) * kH;                       // item = b*F + f, r = t
    const int groups_per_b = kT / 2, b = item / groups_per_b, t0 = 2 * (item % groups_per_b);
    const int tt = r >= kF ? 1 : 0, f = r - tt * kF;                          // r = tt*F + f, two frames per group
    return (((size_t)b * kF + f) * kT + t0 + tt) * kH;
}

// SPLIT = 0: x and dy of a row are loaded together (2*3*U float4 in flight per thread); SPLIT = 1: all x rows of the item first, a
// barrier, then the dy rows (the product kernels' serial staging passes)
template <int MODE, int U, int SPLIT>
__global__ void __launch_bounds__(512) pattern_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dx,
                                                     int nitems, int nrows) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
        for (int R = 4 * U * warp; R < nrows; R += 4 * U * 16) {
            float4 v[U][3], w[U][3];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = R + U * sub + u;
                const bool ok = r < nrows;
                const size_t off = ok ? row_offset<MODE>(item, r) : 0;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    v[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(x + off) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                    if (!SPLIT) w[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(dy + off) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                }
            }
            if (SPLIT) {
                // the dy addresses depend on ALL x values of this thread having arrived (never true, but the compiler cannot know):
                // the second pass is issued only after the first one's data is there, like a staging pass that writes smem first
                float acc = 0.f;
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < 3; ++j) acc += v[u][j].x + v[u][j].w;
                const size_t bump = (acc == 12345.678f) ? 4 : 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int r = R + U * sub + u;
                    const bool ok = r < nrows;
                    const size_t off = (ok ? row_offset<MODE>(item, r) : 0) + bump;
#pragma unroll
                    for (int j = 0; j < 3; ++j) w[u][j] = ok ? __ldg(reinterpret_cast<const float4*>(dy + off) + l8 + 8 * j) : make_float4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = R + U * sub + u;
                if (r >= nrows) continue;
                float4* p = reinterpret_cast<float4*>(dx + row_offset<MODE>(item, r)) + l8;
#pragma unroll
                for (int j = 0; j < 3; ++j) p[8 * j] = add4(v[u][j], w[u][j]);
            }
        }
        if (SPLIT) __syncthreads();
    }
}

__global__ void __launch_bounds__(512) flat_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, float4* __restrict__ dx, size_t n4) {
    for (size_t i = (size_t)blockIdx.x * 512 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 512) dx[i] = add4(__ldg(x + i), __ldg(dy + i));
}

struct Result {
    std::string name;
    int ctas_per_sm, resident;
    double ms, gbs;
};

template <typename K>
static Result run(const char* name, K kern, int sms, int per_sm, const float* x, const float* dy, float* dx, int nitems, int nrows, double bytes) {
    int resident = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, kern, 512, 0));
    const int grid = sms * per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {  // rep 0 = warm-up; the three tensors (1.19 GB) exceed the 126 MB L2 many times over
        CK(cudaEventRecord(e0));
        kern<<<grid, 512>>>(x, dy, dx, nitems, nrows);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    Result r{name, per_sm, resident, best, bytes / (best * 1e-3) / 1e9};
    printf("%-44s CTAs/SM launched %d (resident limit %d)  %.4f ms  %.1f GB/s\n", name, per_sm, resident, r.ms, r.gbs);
    fflush(stdout);
    return r;
}

int main(int argc, char** argv) {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t n = (size_t)kB * kF * kT * kH;
    float *x, *dy, *dx;
    CK(cudaMalloc(&x, n * 4));
    CK(cudaMalloc(&dy, n * 4));
    CK(cudaMalloc(&dx, n * 4));
    CK(cudaMemset(x, 0, n * 4));
    CK(cudaMemset(dy, 0, n * 4));
    CK(cudaMemset(dx, 0, n * 4));
    const double bytes = 3.0 * n * 4;
    std::vector<Result> res;
    printf("stream [%d,%d,%d,%d] fp32 = %.1f MB per tensor; dx = x + dy moves %.2f GB; %d SMs\n", kB, kF, kT, kH, n * 4 / 1e6, bytes / 1e9, sms);
    {
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int per_sm : {1, 2, 4}) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                flat_kernel<<<sms * per_sm, 512>>>((const float4*)x, (const float4*)dy, (float4*)dx, n / 4);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep > 0 && ms < best) best = ms;
            }
            res.push_back({"flat grid-stride float4", per_sm, 4, best, bytes / (best * 1e-3) / 1e9});
            printf("%-44s CTAs/SM launched %d                      %.4f ms  %.1f GB/s\n", "flat grid-stride float4", per_sm, best, res.back().gbs);
            fflush(stdout);
        }
    }
    const int nslab = kB * kF, ngroups = kB * (kT / 2);
    res.push_back(run("slab  U=5 x,dy together", pattern_kernel<0, 5, 0>, sms, 1, x, dy, dx, nslab, kT, bytes));
    res.push_back(run("slab  U=5 x pass, then dy pass", pattern_kernel<0, 5, 1>, sms, 1, x, dy, dx, nslab, kT, bytes));
    res.push_back(run("frame U=5 x,dy together", pattern_kernel<1, 5, 0>, sms, 1, x, dy, dx, ngroups, 2 * kF, bytes));
    res.push_back(run("frame U=5 x pass, then dy pass", pattern_kernel<1, 5, 1>, sms, 1, x, dy, dx, ngroups, 2 * kF, bytes));
    for (int per_sm : {1, 2}) {
        res.push_back(run("slab  U=2 x,dy together", pattern_kernel<0, 2, 0>, sms, per_sm, x, dy, dx, nslab, kT, bytes));
        res.push_back(run("frame U=2 x,dy together", pattern_kernel<1, 2, 0>, sms, per_sm, x, dy, dx, ngroups, 2 * kF, bytes));
    }
    for (int per_sm : {1, 2, 4}) {
        res.push_back(run("slab  U=1 x,dy together", pattern_kernel<0, 1, 0>, sms, per_sm, x, dy, dx, nslab, kT, bytes));
        res.push_back(run("frame U=1 x,dy together", pattern_kernel<1, 1, 0>, sms, per_sm, x, dy, dx, ngroups, 2 * kF, bytes));
    }
    if (argc > 1) {
        FILE* f = fopen(argv[1], "w");
        if (f) {
            fprintf(f, "{\"what\": \"dx = x + dy over the fp32 stream [32,129,250,96] (1.19 GB of traffic) with the access patterns of the product kernels, no arithmetic; 512-thread persistent CTAs; tools/pattern_bw.cu\", \"sms\": %d, \"results\": [\n", sms);
            for (size_t i = 0; i < res.size(); ++i)
                fprintf(f, "  {\"pattern\": \"%s\", \"ctas_per_sm\": %d, \"resident_limit\": %d, \"ms\": %.4f, \"GBps\": %.1f}%s\n", res[i].name.c_str(),
                        res[i].ctas_per_sm, res[i].resident, res[i].ms, res[i].gbs, i + 1 < res.size() ? "," : "");
            fprintf(f, "]}\n");
            fclose(f);
        }
    }
    return 0;
}
