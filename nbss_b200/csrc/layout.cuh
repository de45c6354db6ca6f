// layout.cuh — compile-time geometry of the SpatialNet-small hot path and of the packed weight images.
//
// The tensor-core kernels are specialised for the configuration named by BASELINE.json / configs/SpatialNet.yaml
// (dim_hidden 96, dim_ffn 192, 4 heads, 8 conv groups, kernel sizes (5,3)).  Other shapes return
// NBSS_ERR_UNSUPPORTED from the C-ABI.
//
// "Weight image" = the exact bytes of a UMMA B-operand tile in the chunk-column smem layout of umma.cuh
// (addr(n, k) = (k/8)*rows*16 + n*16 + (k%8)*2, 16-bit elements), so a kernel brings a weight into shared memory
// with ONE cp.async.bulk (TMA bulk copy) and points a descriptor at it.
#pragma once
#include <cstdint>

namespace nbss {

constexpr int kH = 96;        // dim_hidden
constexpr int kHF = 192;      // dim_ffn
constexpr int kNH = 4;        // heads
constexpr int kDH = 24;       // head dim
constexpr int kGroups = 8;    // conv groups (both F- and T-conv)
constexpr int kGC = 24;       // channels per T-conv group (192/8)
constexpr int kPairs = 4;     // T-conv group pairs (48 channels) — the MMA N/K granule
constexpr int kTMax = 256;    // max frames per slab handled by the slab kernels (2 M-tiles of 128)
constexpr int kRT = 265;      // rows of a slab operand tile incl. halo/padding (chunk stride 4240 B: 16 mod 128)
constexpr uint32_t kCS = kRT * 16;

// ---- per-layer weight image offsets (bytes) ----
constexpr uint32_t IMG_W1 = 0;                       // pw1  [N=192 x K=96]   fwd
constexpr uint32_t IMG_W1_BYTES = 12 * 192 * 16;     // 36864
constexpr uint32_t IMG_WC_BYTES = 12 * 6 * 48 * 16;  // 55296: 4 pairs x 3 taps x [N=48 x K=48]
constexpr uint32_t IMG_WC1 = IMG_W1 + IMG_W1_BYTES;
constexpr uint32_t IMG_WC2 = IMG_WC1 + IMG_WC_BYTES;
constexpr uint32_t IMG_WC3 = IMG_WC2 + IMG_WC_BYTES;
constexpr uint32_t IMG_W2 = IMG_WC3 + IMG_WC_BYTES;  // pw2  [N=96 x K=192]   fwd
constexpr uint32_t IMG_W2_BYTES = 24 * 96 * 16;      // 36864
constexpr uint32_t IMG_W2T = IMG_W2 + IMG_W2_BYTES;  // dgrad of pw2: [N=192 x K=96], elem(n,k) = W2[k,n]
constexpr uint32_t IMG_WC3T = IMG_W2T + IMG_W1_BYTES;  // dgrad of convs: per (pair,tap) [N=ci x K=co]
constexpr uint32_t IMG_WC2T = IMG_WC3T + IMG_WC_BYTES;
constexpr uint32_t IMG_WC1T = IMG_WC2T + IMG_WC_BYTES;
constexpr uint32_t IMG_W1T = IMG_WC1T + IMG_WC_BYTES;  // dgrad of pw1: [N=96 x K=192], elem(n,k) = W1[k,n]
constexpr uint32_t IMG_WKV = IMG_W1T + IMG_W2_BYTES;   // in_proj rows 96..287: [N=192 x K=96]
constexpr uint32_t IMG_WQ = IMG_WKV + IMG_W1_BYTES;    // in_proj rows 0..95:   [N=96 x K=96]
constexpr uint32_t IMG_WQ_BYTES = 12 * 96 * 16;        // 18432
constexpr uint32_t IMG_WO = IMG_WQ + IMG_WQ_BYTES;     // out_proj [N=96 x K=96]
constexpr uint32_t IMG_WOT = IMG_WO + IMG_WQ_BYTES;    // dgrad out_proj: elem(n,k) = Wo[k,n]
constexpr uint32_t IMG_WINT = IMG_WOT + IMG_WQ_BYTES;  // dgrad in_proj: [N=96 x K=288], elem(n,k) = Win[k,n]
constexpr uint32_t IMG_WINT_BYTES = 36 * 96 * 16;      // 55296
constexpr uint32_t IMG_LAYER_BYTES = IMG_WINT + IMG_WINT_BYTES;  // 626688

}  // namespace nbss
