/* nbss_b200.h — C ABI of libnbss_b200.so: the B200 (sm_100a) SpatialNet hot path of Audio-WestlakeU/NBSS.
 *
 * The reference has no FFI: its hot path is a Python nn.Module injection point (TrainModule.__init__(arch, stft, norm),
 * SharedTrainer.py:38-63, populated from configs/SpatialNet.yaml:11-43).  This library is what the drop-in modules in
 * nbss_b200/{spatialnet,io}.py bind with ctypes; INTEGRATION.md shows the reference-side change (three class_path lines).
 *
 * Conventions
 *   - all pointers are DEVICE pointers owned by the caller (PyTorch's allocator); the library never allocates, frees or
 *     synchronises; `stream` is a cudaStream_t passed as void*; every call only enqueues work on that stream.
 *   - return value: 0 ok; < 0 argument error (-1 shape, -2 null pointer, -3 unsupported configuration, -4 workspace);
 *     > 0 the cudaError_t of the launch.  Nothing throws.
 *   - tensor-core kernels take `int* err`: a device int that is set to 0x7001 if an mbarrier wait timed out (bug guard).
 *   - stream tensor: fp32 [B,F,T,96] (H contiguous); "slab" = one (b,f) pair = T x 96 contiguous floats; nslab = B*F.
 *   - fmt / fmt_g / fmt_a: 16-bit tensor-core operand format, 0 = fp16, 1 = bf16.
 *   - parameter pointers use the reference's state_dict tensors unchanged, fp32 contiguous.
 *   - every *_bwd / *_wgrad entry ACCUMULATES (+=, fp32 atomics) into the parameter-gradient buffers it is given.
 */
#ifndef NBSS_B200_H
#define NBSS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- weight images (pack.cu): UMMA B-operand tiles of one SpatialNetLayer's narrow-band weights ------------------- */
unsigned int nbss_layer_image_bytes(void);
/* ABI version (100 * round + revision) and the bytes of caller-owned scratch one SpatialNetLayer needs at [B,F,T]
 * (training: activations saved for the backward + transient gradient operands; inference: 0). */
int nbss_version(void);
long long nbss_workspace_bytes(int B, int F, int T, int training);
/* w1 = tconvffn.1.weight [192,96,1], wc{1,2,3} = tconvffn.{3,5,8}.weight [192,24,3], w2 = tconvffn.10.weight [96,192,1],
 * w_in = mhsa.in_proj_weight [288,96], w_out = mhsa.out_proj.weight [96,96]  (models/arch/SpatialNet.py:58,61-73) */
int nbss_pack_layer_weights(const float* w1, const float* wc1, const float* wc2, const float* wc3, const float* w2,
                            const float* w_in, const float* w_out, void* img, int fwd_fmt, int bwd_fmt, void* stream);

/* ---- narrow-band block, forward (tcgen05) -------------------------------------------------------------------------- */
/* y = x + MHSA(LN(x)) over T per (b,f): SpatialNetLayer._tsa + residual, models/arch/SpatialNet.py:88-89,93-100.
 * save_* (nullable): fp16 (scaled q|k|v) [n,288], fp16 O [n,96], log2-domain logsumexp [nslab,4,T], LN (mean,rstd) [n,2]. */
int nbss_mhsa_fwd(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b_in,
                  const float* b_out, const void* layer_img, void* save_qkv, void* save_o, float* save_lse,
                  float* ln_stats, int fmt, int* err, void* stream);
/* y = x + tconvffn(x): SpatialNetLayer._tconvffn + residual, models/arch/SpatialNet.py:90,102-114 (modules :61-73).
 * save_* (nullable): fp16 pre-activations a1,c1,c2,c3 [n,192]; gn_stats [nslab,8,2]; ln_stats [n,2]. */
int nbss_ffn_fwd(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b1,
                 const float* bc1, const float* bc2, const float* bc3, const float* gn_w, const float* gn_b,
                 const float* b2, const void* layer_img, void* save_a1, void* save_c1, void* save_c2, void* save_c3,
                 float* gn_stats, float* ln_stats, int fmt, int* err, void* stream);

/* The same kernel for 4 heads (SpatialNet-small, head dim 24) or 2 heads (NBC2 small, head dim 48: models/arch/NBC2.py:294-311).
 * row_part (nullable): [n,2] per output row (sum, sum of squares) over the 96 channels — the per-slab partials of NBC2's
 * GroupBatchNorm norm2 (NBC2.py:111-145,170), reduced over the F slabs of an utterance by nbss_gbn_reduce. */
int nbss_mhsa_fwd_nh(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b_in,
                     const float* b_out, const void* layer_img, void* save_qkv, void* save_o, float* save_lse,
                     float* ln_stats, float* row_part, int num_heads, int fmt, int* err, void* stream);

/* ---- predict / test post-processing (predict.cu): recover_scale (models/utils/metrics.py:192-218, scale_src_together=False) and
 * the peak normalisation of SharedTrainer.py:301-305.  preds [B,S,Ts] (S <= 4), mixture [B,Ts] (reference channel), out may
 * alias preds; ws_sums: 14*B doubles, ws_peak: B*S uints (scratch, zeroed inside); scale_out (nullable) [B,S]. */
int nbss_predict_post(const float* preds, const float* mixture, float* out, int B, int S, long long Ts, int recover,
                      int norm_if_exceed_1, double* ws_sums, unsigned int* ws_peak, float* scale_out, void* stream);

/* ---- exact loss scaling of the upstream gradient (loss.cu): dy_scaled = dy * s with s = 2^-round(log2 max|dy|) (1 if dy == 0); the
 * backward kernels are linear in dy, nbss_grad_unscale multiplies the flat gradient buffer by 1/s.  ws: 1 uint, scale: 2 floats (s, 1/s);
 * everything stays on the device (no host synchronisation, graph-capturable). */
int nbss_grad_prescale(const float* dy, long long n, float* dy_scaled, unsigned int* ws, float* scale, void* stream);
int nbss_grad_unscale(float* flat, long long n, const float* scale, void* stream);

/* ---- T > 256, inference only (validation / test utterances are longer than the 4 s training crops; SharedTrainer.py:134-189) -----
 * Same layer image and parameters as nbss_mhsa_fwd / nbss_ffn_fwd; 256 < T <= 65536; fmt = 0 (fp16).  In place (y == x) is allowed.
 * nbss_mhsa_fwd_long: kv_ws = fp16 [nslab][36][T][8] (576*T bytes per slab): pass 1 writes k | v of all frames, pass 2 attends per
 *   256-frame query chunk over the slab's key blocks (flash-style running softmax), out-proj + residual.
 * nbss_ffn_fwd_long: the T-ConvFFN cut at its GroupNorm and tiled over T with halo rows; c2_ws fp16 [nslab][24][T][8] (384*T bytes
 *   per slab), part_ws [nslab][nbss_ffn_long_chunks(T,0)][8][2] floats, stats_ws [nslab][8][2] floats. */
int nbss_mhsa_fwd_long(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b_in,
                       const float* b_out, const void* layer_img, void* kv_ws, int fmt, int* err, void* stream);
long long nbss_ffn_long_chunks(int T, int part);
int nbss_ffn_fwd_long(const float* x, float* y, int nslab, int T, const float* ln_w, const float* ln_b, const float* b1,
                      const float* bc1, const float* bc2, const float* bc3, const float* gn_w, const float* gn_b, const float* b2,
                      const void* layer_img, void* c2_ws, float* part_ws, float* stats_ws, int fmt, int* err, void* stream);

/* ---- online / causal SpatialNet, one frame per call (online.cu; models/arch/OnlineSpatialNet.py with attention='mhsa(N)') ------
 * R = B*F rows of one frame, fp32 [R,96] stream updated in place; `pos` = device int, frames consumed so far (nbss_online_advance
 * increments it, so a captured CUDA graph of the whole step replays unchanged); weights marked T are [in][out] (nbss_transpose;
 * the grouped T-conv weights [192][24][3] are passed as their transpose [72][192]).  A CTA owns 1 row (R <= 600) or 4 rows. */
int nbss_transpose(const float* in, float* out, int rows, int cols, void* stream);
int nbss_online_pack_encoder(const float* W /*[96][Cin][5]*/, float* Wt /*[5*Cin][96]*/, int Cin, void* stream);
int nbss_online_encoder_step(const float* xt /*[R,Cin]*/, float* state /*[R,4,Cin]*/, const float* Wt, const float* bias, float* h,
                             int R, int Cin, void* stream);
int nbss_online_attn_step(float* x, int R, const float* ln_w, const float* ln_b, const float* WinT, const float* b_in,
                          const float* WoT, const float* b_out, float* kcache /*[R,scope,96]*/, float* vcache, const int* pos,
                          int scope, void* stream);
int nbss_online_ffn_a_step(const float* x, int R, const float* ln_w, const float* ln_b, const float* W1T, const float* b1,
                           const float* Wc1T /*[72][192]*/, const float* bc1, const float* Wc2T, const float* bc2, float* st1 /*[R,2,192]*/,
                           float* st2, float* c2 /*[R,192]*/, float* part /*[R,8,2]*/, void* stream);
int nbss_online_gn_stats(const float* part, int B, int F, float* stats /*[B,8,2]*/, void* stream);
int nbss_online_ffn_b_step(float* x, int R, int F, const float* c2, const float* stats, const float* gn_w, const float* gn_b,
                           const float* Wc3T, const float* bc3, const float* W2T, const float* b2, float* st3, void* stream);
int nbss_online_advance(int* pos, void* stream);

/* ---- NBC2 inference (BASELINE configs[3]; models/arch/NBC2.py:152-289) ------------------------------------------------ */
/* GroupBatchNorm statistics (NBC2.py:118-128): part [B][NP][T][NQ][2] (sum, sum of squares) -> stats [B][T][2] (mean, rstd)
 * over the NP*NQ partials of `count` elements in total (fp64 accumulation). */
int nbss_gbn_reduce(const float* part, int B, int NP, int T, int NQ, long long count, float eps, float* stats, void* stream);
/* NBC2Block._ff_block cut at its second GroupBatchNorm (conv.4), whose statistics span all F slabs of an utterance:
 *   part A: c2 = conv.3(SiLU(conv.1(SiLU(linear1(GBN_norm2(x)))))) -> fp16 slab tiles [nslab][24][T][8] + partial sums
 *           part [nslab][T][2][2]; row_stats = (mean, rstd) [B*T] of norm2, gbn_w/gbn_b = norm2.{weight,bias} [96]
 *   part B: y = x + linear2(SiLU(conv.6(SiLU(GBN_conv4(c2))))); row_stats of conv.4, gbn_w/gbn_b = conv.4.{weight,bias} [192]
 * layer_img: nbss_pack_layer_weights(linear1.weight, conv.1.weight, conv.3.weight, conv.6.weight, linear2.weight, in_proj, out_proj). */
int nbss_nbc2_ffn_a(const float* x, int nslab, int T, int F, const float* row_stats, const float* gbn_w, const float* gbn_b,
                    const float* b1, const float* bc1, const float* bc2, const void* layer_img, void* c2_out, float* part,
                    int fmt, int* err, void* stream);
int nbss_nbc2_ffn_b(const float* x, float* y, int nslab, int T, int F, const float* row_stats, const float* gbn_w,
                    const float* gbn_b, const float* bc3, const float* b2, const void* layer_img, const void* c2_in, int fmt,
                    int* err, void* stream);

/* ---- narrow-band block, backward (tcgen05); autograd of the above -------------------------------- */
int nbss_ffn_bwd(const float* x, const float* dy, float* dx, int nslab, int T, const float* ln_w, const float* gn_w,
                 const float* gn_b, const float* ln_stats, const float* gn_stats, const void* layer_img, const void* a1,
                 const void* c1, const void* c2, const void* c3, void* g_a1, void* g_c1, void* g_c2, void* g_c3,
                 float* d_lnw, float* d_lnb, float* d_gnw, float* d_gnb, int fmt, int* err, void* stream);
/* a1, c1, c2, c3: the forward's saved fp16 pre-activations; the activation operands SiLU(.) / SiLU(GroupNorm(c2)) of the weight
 * gradients are recomputed from them inside the kernel (gn_stats [nslab,8,2], gn_w / gn_b [192]) */
int nbss_ffn_wgrad(const float* x, const float* dy, int nslab, int T, const float* ln_w, const float* ln_b,
                   const void* g_a1, const void* g_c1, const void* g_c2, const void* g_c3, const void* a1, const void* c1,
                   const void* c2, const void* c3, const float* gn_stats, const float* gn_w, const float* gn_b, float* dW1,
                   float* db1, float* dWc1, float* dbc1, float* dWc2, float* dbc2, float* dWc3, float* dbc3, float* dW2,
                   float* db2, int fmt_g, int fmt_a, int* err, void* stream);
int nbss_mhsa_bwd(const float* x, const float* dy, float* dx, int nslab, int T, const float* ln_w, const float* ln_stats,
                  const void* layer_img, const void* qkv, const void* o, const float* lse, void* dqkv, float* d_lnw,
                  float* d_lnb, int fmt_g, int* err, void* stream);
int nbss_mhsa_wgrad(const float* x, const float* dy, int nslab, int T, const float* ln_w, const float* ln_b,
                    const void* dqkv, const void* o, float* dWin, float* dbin, float* dWo, float* dbo, int fmt_g,
                    int fmt_a, int* err, void* stream);

/* ---- cross-band block, fp32 (models/arch/SpatialNet.py:85-87,116-146; base/linear_group.py:29-34) ------------------ */
/* y = x + PReLU(gconv_F(LN(x))): _fconv with modules :36-40 / :49-53.  W [96,12,5]. */
int nbss_fconv_fwd(const float* x, float* y, int B, int F, int T, const float* lnw, const float* lnb, const float* W,
                   const float* bias, const float* slope, void* stream);
int nbss_fconv_bwd(const float* x, const float* dy, float* dx, int B, int F, int T, const float* lnw, const float* lnb,
                   const float* W, const float* bias, const float* slope, float* dW, float* dbias, float* dslope,
                   float* dlnw, float* dlnb, void* stream);
/* The same sub-block on tensor cores (fconv_tc.cu): frames stacked into one UMMA tile with zero gap rows, taps as
 * row-shifted views, block-diagonal 48x48 weight tiles; img from nbss_fconv_pack (nbss_fconv_image_bytes() bytes). */
unsigned int nbss_fconv_image_bytes(void);
int nbss_fconv_pack(const float* W, void* img, int fmt, void* stream);
int nbss_fconv_tc_fwd(const float* x, float* y, int B, int F, int T, const float* lnw, const float* lnb, const float* bias,
                      const float* slope, const void* img, int fmt, int* err, void* stream);
int nbss_fconv_tc_bwd(const float* x, const float* dy, float* dx, int B, int F, int T, const float* lnw, const float* lnb,
                      const float* bias, const float* slope, const void* img, float* dW, float* dbias, float* dslope,
                      float* dlnw, float* dlnb, int fmt, int* err, void* stream);
/* y = x + unsqueeze(full(squeeze(LN(x)))): _full :129-146.  s_out,u_out: [B,T,8,F] (kept for backward). */
int nbss_full_fwd(const float* x, float* y, float* s_out, float* u_out, int B, int F, int T, const float* lnw,
                  const float* lnb, const float* Wsq, const float* bsq, const float* Wf, const float* bf, const float* Wun,
                  const float* bun, void* stream);
/* ws: workspace of 2*B*T*8*F floats. */
int nbss_full_bwd(const float* x, const float* dy, float* dx, const float* s, const float* u, float* ws, int B, int F, int T,
                  const float* lnw, const float* lnb, const float* Wsq, const float* bsq, const float* Wf, const float* Wun,
                  const float* bun, float* dlnw, float* dlnb, float* dWsq, float* dbsq, float* dWf, float* dbf, float* dWun,
                  float* dbun, void* stream);

/* The same block with the LinearGroup (models/arch/base/linear_group.py:29-34) on tensor cores (fullband_tc.cu):
 * img = nbss_lg_pack(full.weight [8,F,F]) (nbss_lg_image_bytes(F) bytes, F <= 256; one image serves forward and data
 * gradient).  nbss_lg_tc_apply: mode 0 out = in W^T + bias, mode 1 out = in W; in/out fp32 [M,8,F].
 * nbss_lg_tc_wgrad: dW [8,F,F] += du^T s, db [8,F] += column sums of du. */
unsigned int nbss_lg_image_bytes(int F);
int nbss_lg_pack(const float* Wf, void* img, int F, int fmt, void* stream);
int nbss_lg_tc_apply(const float* in, float* out, int M, int F, const void* img, const float* bias, int mode, int fmt,
                     int* err, void* stream);
int nbss_lg_tc_wgrad(const float* du, const float* s, int M, int F, float* dW, float* db, int fmt, int* err, void* stream);
int nbss_full_fwd_tc(const float* x, float* y, float* s_out, float* u_out, int B, int F, int T, const float* lnw,
                     const float* lnb, const float* Wsq, const float* bsq, const float* bf, const float* Wun,
                     const float* bun, const void* img, int fmt, int* err, void* stream);
int nbss_full_bwd_tc(const float* x, const float* dy, float* dx, const float* s, const float* u, float* ws, int B, int F,
                     int T, const float* lnw, const float* lnb, const float* Wsq, const float* bsq, const float* Wun,
                     const float* bun, const void* img, float* dlnw, float* dlnb, float* dWsq, float* dbsq, float* dWf,
                     float* dbf, float* dWun, float* dbun, int fmt, int* err, void* stream);

/* The per-point halves of that block on tensor cores (fullband_rows_tc.cu), used by nbss_full_{fwd,bwd}_tc
 * (models/arch/SpatialNet.py:129-137 squeeze, :143-146 unsqueeze; modules :41-48).  s, u, ds, du: fp32 [B,T,8,F]. */
int nbss_squeeze_fwd_tc(const float* x, float* s, int B, int F, int T, const float* lnw, const float* lnb, const float* Wsq,
                        const float* bsq, int fmt, int* err, void* stream);
int nbss_unsqueeze_fwd_tc(const float* x, const float* u, float* y, int B, int F, int T, const float* Wun, const float* bun,
                          int fmt, int* err, void* stream);
int nbss_unsqueeze_bwd_tc(const float* dy, const float* u, float* du, int B, int F, int T, const float* Wun, const float* bun,
                          float* dWun, float* dbun, int fmt, int* err, void* stream);
int nbss_squeeze_bwd_tc(const float* x, const float* dy, const float* ds, float* dx, int B, int F, int T, const float* lnw,
                        const float* lnb, const float* Wsq, const float* bsq, float* dWsq, float* dbsq, float* dlnw,
                        float* dlnb, int fmt, int* err, void* stream);

/* ---- encoder / decoder, fp32 (models/arch/SpatialNet.py:175,205 and :200,216) --------------------------------------- */
int nbss_encoder_fwd(const float* x, float* y, int nslab, int T, int cin, const float* W, const float* bias, void* stream);
int nbss_encoder_wgrad(const float* x, const float* dy, int nslab, int T, int cin, float* dW, float* dbias, void* stream);
int nbss_decoder_fwd(const float* x, float* y, long long n, int cout, const float* W, const float* bias, void* stream);
int nbss_decoder_bwd(const float* x, const float* dy, float* dx, long long n, int cout, const float* W, float* dW,
                     float* dbias, void* stream);

/* ---- framing (models/io/stft.py:49-97, models/io/norm.py:61-108, SharedTrainer.py:113-131) ------------------------- */
/* STFT of x [B,C,Ts] (center, reflect, periodic Hann, onesided).  Output element (b,c,f,t): re at out + b*ob + c*oc +
 * f*of + t*ot (floats), im at +1 — so the same kernel writes complex [B,C,F,T] or the packed network input [B,F,T,2C].
 * normalize != 0 fuses Norm(mode='frequency', online=True): XrMM = |X[ref]| + eps -> xrmm [B,F,T]; xr (nullable). */
int nbss_stft(const float* x, int B, int C, int Ts, int n_fft, int hop, int normalize, int ref_channel, float eps,
              float* out, long long ob, long long oc, long long of, long long ot, float* xrmm, float* xr, void* stream);
/* iSTFT (window, overlap-add, envelope division, centre trim) of a complex tensor given by float strides (b,s,f,t);
 * scale (nullable) [B,F,T] multiplies every speaker's bin first (Norm.inorm).  y [B,S,Ts]. */
int nbss_istft(const float* in, long long ib, long long is, long long if_, long long it, const float* scale, float* y,
               int B, int S, int Ts, int T, int n_fft, int hop, void* stream);
int nbss_istft_bwd(const float* dy, const float* scale, float* din, long long ib, long long is, long long if_, long long it,
                   int B, int S, int Ts, int T, int n_fft, int hop, void* stream);
int nbss_norm_freq_online(float* X, int B, int C, long long FT, int ref_channel, float eps, float* xrmm, float* xr,
                          void* stream);
int nbss_inorm(const float* X, float* Y, int B, int S, long long FT, const float* xrmm, void* stream);

/* ---- loss next to the path: negative SI-SDR with 2-speaker PIT ------------------------------------
 * Loss.forward for loss_func = neg_si_sdr, pit = True (models/io/loss.py:21-29,95-118; configs/SpatialNet.yaml:33-37):
 * torchmetrics scale_invariant_signal_distortion_ratio (zero_mean as given; the reference uses the default, 0) and
 * permutation_invariant_training(mode="permutation-wise", eval_func="min").  est, ref: fp32 [B,2,Ts].
 * sums: workspace of B*12 doubles; loss[1] = batch mean; loss_b[B], perm[B,2] (int), coef[B,2,5] nullable; coef feeds the
 * backward, which writes dest = gout[0] * d loss / d est (gout: device scalar, NULL = 1). */
int nbss_sisdr_pit_fwd(const float* est, const float* ref, int B, int S, long long Ts, int zero_mean, double* sums,
                       float* loss, float* loss_b, int* perm, float* coef, void* stream);
int nbss_sisdr_pit_bwd(const float* est, const float* ref, const float* coef, const float* gout, float* dest, int B, int S,
                       long long Ts, void* stream);

/* Optimiser tail: torch.nn.utils.clip_grad_norm_(max_norm) + torch.optim.Adam (configs/SpatialNet.yaml:
 * 3-4,44: Adam lr 1e-3, gradient_clip_val 5) over the network's single flat gradient buffer, two launches.
 * params: device array of ntensors parameter pointers; offsets: device array of ntensors+1 cumulative element offsets;
 * flat_grad / exp_avg / exp_avg_sq: fp32 [n]; gnorm_sq: device double (out: squared gradient norm before clipping);
 * step: device float step counter, incremented by the call (bias correction uses the incremented value, as torch does). */
int nbss_clip_adam(float* const* params, const long long* offsets, int ntensors, long long n, const float* flat_grad,
                   float* exp_avg, float* exp_avg_sq, double* gnorm_sq, float* step, float max_norm, float lr, float beta1,
                   float beta2, float eps, void* stream);

/* ---- test hook: one-CTA tcgen05 GEMM that pins the descriptor conventions (tests/test_umma_selftest.py) ------------- */
int nbss_umma_selftest(const float* A, int a_rows, int a_feats, const float* B, int b_rows, int b_feats, float* D, int N,
                       int Kdim, int a_mn, int b_mn, int fmt, int a_shift, int b_shift, int a_off, int b_off, int passes,
                       int tmem_col, int fmt_b, int* err, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NBSS_B200_H */
