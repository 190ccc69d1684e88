/*
 * ldconv_b200.h -- C ABI of libldconv_b200.so: the B200 (sm_100a) implementation of DEAL-YOLO's LDConv hot path.
 *
 * This is the drop-in boundary.  The reference has no native code on this path: everything
 * below replaces eager-PyTorch lines of /root/reference/ultralytics/nn/modules/conv.py:350-503 (class LDConv), and
 * each entry point cites the lines it replaces.  The host-side nn.Module that keeps the reference's Python surface
 * (constructor, state_dict, YAML hook) lives in experiment_yolo_b200/ldconv.py and binds these symbols with ctypes;
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: raw DEVICE pointers, ints, and a cudaStream_t passed as void*; no torch types, no C++ types.
 *   - every call is asynchronous on `stream`, never synchronises the host, allocates no device memory, keeps no
 *     state between calls (deepcopy / pickle of the Python module stay trivial) and is CUDA-graph capturable.
 *   - return value: 0 on success, a negative LDCONV_E_* code otherwise; ldconv_last_error() returns a thread-local
 *     message.  There is no CPU fallback: on a machine without an sm_100 device every compute call fails loudly.
 *   - activations are NHWC ("channels_last"): x (B,H,W,C); dtype LDCONV_F32 or LDCONV_BF16 selects the element type of
 *     x / operand / weights of the (N,1) conv / outputs.  Sampling offsets, coordinates, bilinear weights, BatchNorm
 *     statistics and all accumulators are ALWAYS fp32 (fp64 for the cross-CTA statistics): the reference's own
 *     low-precision coordinate math is broken and is deliberately not reproduced.
 *   - h = (H-1)/s + 1, w = (W-1)/s + 1 (3x3, pad 1, stride s offset conv, conv.py:356); M = B*h*w; K = N*C.
 *   - offsets are (B,h,w,2N) fp32: channels [0,N) are ROW (H axis) offsets, [N,2N) COLUMN (W axis) offsets
 *     (conv.py:370, 463-467).
 *   - the resampled operand is (M, K) row-major with k = n*C + c.  The reference's (N,1)-conv weight (O,C,N,1)
 *     (conv.py:355) must therefore be passed permuted to (O, N, C) -> (O, K); the Python module does that.
 */
#ifndef LDCONV_B200_H
#define LDCONV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDCONV_ABI_VERSION 1

#define LDCONV_F32 0
#define LDCONV_BF16 1

#define LDCONV_ACT_NONE 0
#define LDCONV_ACT_SILU 1
#define LDCONV_ACT_LEAKY01 2   /* LeakyReLU(0.1): ScalSeq (nn/extra_modules/block.py:3424); tcgen05 epilogues only */

#define LDCONV_OK 0
#define LDCONV_E_ARG (-1)      /* bad shape / null pointer / unsupported value */
#define LDCONV_E_ALIGN (-2)    /* pointer not aligned for the vectorised path */
#define LDCONV_E_CUDA (-3)     /* a CUDA runtime / launch error; message holds cudaGetErrorString */
#define LDCONV_E_DEVICE (-4)   /* current device is not sm_100 (no fallback exists) */

/* Which implementation ldconv_gemm_fwd / ldconv_fused_fwd picked for the last call on this thread (for tests and for
 * bench.py's gpu_launches bookkeeping): 1 = CUDA-core FFMA tile kernel, 2 = tcgen05/TMEM kernel. */
#define LDCONV_IMPL_FFMA 1
#define LDCONV_IMPL_TCGEN05 2

int ldconv_version(void);
const char* ldconv_last_error(void);
/* 0 when the current CUDA device is compute capability 10.x, LDCONV_E_DEVICE otherwise. */
int ldconv_device_check(void);
int ldconv_last_impl(void);
/* Debug / A-B switches (thread-local).  LDCONV_FLAG_FORCE_FFMA = 1 routes bf16 GEMMs to the CUDA-core kernel too, so a
 * test can compare the two implementations; it is not a fallback (both are sm_100a CUDA). */
#define LDCONV_FLAG_FORCE_FFMA 1
/* LDCONV_FLAG_GATHER_DIRECT = 1 makes ldconv_gather_fwd use the direct-load kernel instead of the TMA-staged tile kernel
 * (same arithmetic, same results; A/B for profiles). */
#define LDCONV_FLAG_GATHER_DIRECT 2
int ldconv_set_flag(int flag, int value);
/* Optional device counter (uint64) that the TMA-tiled gather increments by the number of samples whose corners fell
 * outside the staged tile + halo and were served from L2 instead (halo miss rate); NULL disables it.  Thread-local. */
int ldconv_set_gather_miss_counter(void* device_u64);

/* conv.py:413-432 (_get_p_n): writes the 2N int32 base-grid table (rows then columns).  Host-side helper, no GPU. */
int ldconv_p_n(int N, int32_t* out_host);

/* conv.py:356,368  offset = p_conv(x): 3x3 / pad 1 / stride s, C -> 2N, + bias.
 *   x    (B,H,W,C) dtype        w (3,3,C,2N) fp32 (the reference's (2N,C,3,3) permuted)      bias (2N) fp32 or NULL
 *   off  (B,h,w,2N) fp32 */
int ldconv_offset_conv_fwd(const void* x, const float* w, const float* bias, float* off,
                           int B, int C, int H, int W, int N, int s, int dtype, void* stream);

/* The same offset conv on the tensor cores (bf16 x, C % 16 == 0): im2col in shared memory from a TMA-staged tile +
 * tcgen05.mma with N = 16/32, fp32 accumulation in TMEM.  w_bf16 is (2N, 9*C) bf16 with k = (ky*3+kx)*C + c (the
 * reference's (2N,C,3,3) permuted to (2N,3,3,C)); bias (2N) fp32 or NULL; off (B,h,w,2N) fp32.
 * ldconv_offset_conv_tc_supported returns 1 when the shape is covered. */
int ldconv_offset_conv_tc_supported(int C, int N, int stride, int dtype);
int ldconv_offset_conv_tc_fwd(const void* x, const void* w_bf16, const float* bias, float* off,
                              int B, int C, int H, int W, int N, int stride, int dtype, void* stream);

/* Stride 2 only: the same conv as a ZERO-COPY tcgen05 implicit GEMM on the space-to-depth view x'[i,j,(sy,sx,c)] =
 * x[2i+sy, 2j+sx, c] (never materialised: a 5-D TMA tensor map produces it), i.e. a 2x2 / stride-1 conv with 4C channels.
 * w_s2d (2N, 16*C) bf16: the reference's (2N,C,3,3) weight scattered to k = ((ty*2+tx)*2+sy)*2C + sx*C + c with
 * (ty,sy) = (0,1),(1,0),(1,1) for ky = 0,1,2 (likewise kx -> (tx,sx)), zeros elsewhere.  C in {16, 32}, H and W even. */
int ldconv_offset_conv_s2d_supported(int C, int N, int H, int W, int dtype);
int ldconv_offset_conv_s2d_fwd(const void* x, const void* w_s2d, const float* bias, float* off, int B, int C, int H, int W,
                               int N, int dtype, void* stream);

/* conv.py:369-407 + 413-503: sampling grid p = p_0 + p_n + offset, floor / independent clamps, four corner indices,
 * four bilinear weights, the four gathers, the bilinear sum and the 'b c h w n -> b c (h n) w' rearrange, fused.
 *   x (B,H,W,C) dtype; off (B,h,w,2N) fp32; p_n (2N) int32 device table
 *   operand (M, N*C) dtype
 *   dbg_idx   (M,N,4) int32 {r0,r1,k0,k1} or NULL;   dbg_coord (M,N,2) fp32 {clamped row, clamped col} or NULL
 * Indices and coordinates are bit-exact with the reference given the same `off`; the fp32 operand is bit-exact too
 * (same operation order, no FMA contraction), the bf16 operand is that value rounded to nearest-even. */
int ldconv_gather_fwd(const void* x, const float* off, const int32_t* p_n, void* operand, int32_t* dbg_idx,
                      float* dbg_coord, int B, int C, int H, int W, int N, int s, int dtype, void* stream);

/* conv.py:355,408: the (N,1)/(N,1) conv is the GEMM pre(M,O) = operand(M,K) . wt(O,K)^T, followed by the
 * per-channel affine (folded BatchNorm) and SiLU of the Sequential tail.
 *   a (M,K) dtype, wt (O,K) dtype
 *   out (M,O) dtype or NULL: act(acc*scale[o] + shift[o])   (scale/shift fp32 (O); NULL = identity)
 *   pre (M,O) dtype or NULL: the raw accumulator (training: saved for backward)
 *   stat_sum / stat_sqsum (O) fp64 or NULL: += sum_m acc, sum_m acc^2 (caller zero-initialises; BatchNorm batch stats)
 * Also used for the data gradient of that conv (a = grad_pre (M,O'), wt = W^T). */
int ldconv_gemm_fwd(const void* a, const void* wt, const float* scale, const float* shift, void* out, void* pre,
                    double* stat_sum, double* stat_sqsum, int M, int K, int O, int act, int dtype, void* stream);

/* Batch statistics of an existing (M, O) bf16 pre-activation (fp64 column sums / sums of squares, caller zero-inits): first pass of
 * the training forward of a `Conv` block (nn/modules/conv.py:41-59) whose convolution ran elsewhere. */
int ldconv_col_stats(const void* pre, double* stat_sum, double* stat_sqsum, long long M, int O, int dtype, void* stream);

/* torch.nn.BatchNorm2d bookkeeping of conv.py:355 (eps / momentum come from the module, never hard-coded).
 * training != 0: mean/var from stat_sum/stat_sqsum over `count` rows; running <- (1-momentum)*running +
 *   momentum*(mean, unbiased var) (running_* may be NULL); training == 0: mean/var = running_*.
 * Writes scale = gamma*invstd, shift = beta - mean*scale, save_mean, save_invstd (all (O) fp32). */
int ldconv_bn_finalize(const double* stat_sum, const double* stat_sqsum, long long count, const float* gamma,
                       const float* beta, float* running_mean, float* running_var, float eps, float momentum,
                       int training, float* scale, float* shift, float* save_mean, float* save_invstd, int O,
                       void* stream);

/* out = act(pre*scale + shift), (M,O) dtype -> (M,O) dtype: second pass of the training forward. */
int ldconv_bn_act_apply(const void* pre, const float* scale, const float* shift, void* out, long long M, int O,
                        int act, int dtype, void* stream);

/* Backward of act(BN(pre)), pass 1: red[0:O] += sum_m dz, red[O:2O] += sum_m dz*xhat   (fp64, caller zero-inits)
 * with z = pre*scale+shift, dz = grad_out * act'(z), xhat = (pre - mean)*invstd.  red[0:O] is grad beta, red[O:2O] is
 * grad gamma. */
int ldconv_bn_act_bwd_reduce(const void* pre, const void* grad_out, const float* scale, const float* shift,
                             const float* mean, const float* invstd, double* red, long long M, int O, int act,
                             int dtype, void* stream);
/* pass 2: grad_pre = scale * (dz - [training] (red0 + xhat*red1)/M)   (M,O) dtype */
int ldconv_bn_act_bwd_apply(const void* pre, const void* grad_out, const float* scale, const float* shift,
                            const float* mean, const float* invstd, const double* red, void* grad_pre, long long M,
                            int O, int act, int training, int dtype, void* stream);

/* Weight gradient of the (N,1) conv: grad_wt (O,K) fp32 += grad_pre(M,O)^T . operand(M,K)  (caller zero-inits). */
int ldconv_gemm_bwd_weight(const void* grad_pre, const void* operand, float* grad_wt, int M, int K, int O, int dtype,
                           void* stream);

/* Backward of the bilinear resampling (autograd of conv.py:386-405; closed form: SURVEY.md Appendix A).
 *   grad_operand (M, N*C) dtype;  x, off, p_n as in ldconv_gather_fwd
 *   grad_x   (B,H,W,C) fp32, ACCUMULATED with atomics (caller zero-inits): the four scatter_add_ of autograd;
 *            NULL when the input needs no gradient (the image, layer 0)
 *   grad_off (B,h,w,2N) fp32, written: clamp-backward indicator * sum_c (...) */
int ldconv_gather_bwd(const void* grad_operand, const void* x, const float* off, const int32_t* p_n, float* grad_x,
                      float* grad_off, int B, int C, int H, int W, int N, int s, int dtype, void* stream);

/* Backward of the offset conv (conv.py:356): grad_x (B,H,W,C) fp32 += conv_transpose(grad_off, w);
 * grad_w (3,3,C,2N) fp32 += ..., grad_b (2N) fp32 += ...  (caller zero-inits grad_w / grad_b; either may be NULL). */
int ldconv_offset_conv_bwd(const float* grad_off, const void* x, const float* w, float* grad_x, float* grad_w,
                           float* grad_b, int B, int C, int H, int W, int N, int s, int dtype, void* stream);

/* The same backward for bf16 activations with the weight gradient on the tensor cores: grad_off is rounded once to bf16, the
 * 3x3 neighbourhoods are laid out chunk by chunk as an (rows, 9C) matrix in `workspace` (sized to stay in L2) and reduced over
 * the output pixels by the MN-major tcgen05 kernel; the data gradient runs out of shared-memory weights.  Same outputs and
 * accumulate semantics as ldconv_offset_conv_bwd.  `workspace`: device memory of at least
 * ldconv_offset_conv_bwd_workspace_bytes(...) bytes, owned by the caller (the library never allocates). */
size_t ldconv_offset_conv_bwd_workspace_bytes(int B, int C, int H, int W, int N, int s, int dtype);
int ldconv_offset_conv_bwd_tc(const float* grad_off, const void* x, const float* w, float* grad_x, float* grad_w, float* grad_b,
                              void* workspace, size_t workspace_bytes, int B, int C, int H, int W, int N, int s, int dtype,
                              void* stream);

/* The same two backward steps with a 16-BIT accumulator for grad_x (bf16 activations only; C % 8 == 0): grad_x (B,H,W,C) bf16,
 * zero-initialised by the caller.  ldconv_gather_bwd_acc16 = ldconv_gather_bwd with eight channels per 16-byte reduction
 * (red.global.add.noftz.v4.bf16x2) instead of four -- half the L2 reduction requests, which bound the scatter -- and
 * ldconv_offset_conv_bwd_tc_acc16 adds the offset conv's data gradient onto the same bf16 tensor, so the fp32 buffer, half of its
 * memset and the fp32 -> bf16 cast pass of the fp32 route disappear.  The accumulator rounds to bf16 after every addition (what
 * autograd's scatter_add_ does for a reduced-precision model, conv.py:456-489 backward): parity bound rel-L2 <= 1e-2 against the
 * fp32 reference on bf16-rounded tensors in the benchmark regime (tests/test_gpu_parity.py); samples that pile onto one pixel
 * by the hundreds (a diverged offset conv) lose low-order bits -- use the fp32 entry points there.  Shapes outside the
 * kernels' range (ldconv_bwd_acc16_supported returns 0) return LDCONV_E_ARG. */
int ldconv_bwd_acc16_supported(int B, int C, int H, int W, int N, int s);
int ldconv_gather_bwd_acc16(const void* grad_operand, const void* x, const float* off, const int32_t* p_n, void* grad_x,
                            float* grad_off, int B, int C, int H, int W, int N, int s, void* stream);
int ldconv_offset_conv_bwd_tc_acc16(const float* grad_off, const void* x, const float* w, void* grad_x, float* grad_w,
                                    float* grad_b, void* workspace, size_t workspace_bytes, int B, int C, int H, int W, int N,
                                    int s, void* stream);

/* Inference forward of the whole module (conv.py:366-410, eval mode) in ONE kernel; the resampled operand never touches
 * HBM.  Two kernels behind it: C <= 4 (the first layer) runs one thread per output pixel on CUDA cores; C % 16 == 0 with
 * bf16 runs offset conv + grid + gather into shared memory in the tcgen05 operand layout + UMMA with TMEM accumulators +
 * folded BatchNorm + SiLU epilogue.  x (B,H,W,C), wt (O,K), out (B,h,w,O) all `dtype`; w_off (3,3,C,2N) / b_off fp32;
 * off_out (B,h,w,2N) fp32 may be NULL.  ldconv_fused_supported returns 1 when a shape is covered (else use the
 * offset_conv / gather / gemm entry points; ldconv_fused_fwd returns LDCONV_E_ARG). */
int ldconv_fused_supported(int B, int C, int H, int W, int N, int s, int O, int dtype);
int ldconv_fused_fwd(const void* x, const float* w_off, const float* b_off, const int32_t* p_n, const void* wt,
                     const float* scale, const float* shift, void* out, float* off_out, int B, int C, int H, int W,
                     int N, int s, int O, int act, int dtype, void* stream);

/* The WHOLE eval-mode forward of conv.py:366-410 in one persistent kernel, x read once (bf16): the 3x3 offset conv (:356,:368)
 * runs as a zero-copy tcgen05 implicit GEMM on the same TMA-staged tile the gather samples from (stride 2: on the
 * space-to-depth view), its 2N offsets go TMEM -> registers -> sampling grid, never to HBM; then as ldconv_gather_gemm_fwd.
 *   x (B,H,W,C) bf16 dense NHWC; w_offconv bf16: stride 1 (2N, 9C) with k = (ky*3+kx)*C + c, stride 2 (2N, 16C) in the
 *   space-to-depth order of ldconv_offset_conv_s2d_fwd; b_off (2N) fp32 or NULL; p_n (2N) int32; wt (O, N*C) bf16;
 *   scale / shift (O) fp32; out (B,h,w,O) bf16 with pixel stride ldo; off_out (B,h,w,2N) fp32 or NULL (debug / tests: the
 *   offsets the kernel used, bit-identical to ldconv_offset_conv_{tc,s2d}_fwd).
 * Covered: the yolov8-LD-P2 shapes (num_param 1 / stride 1 / C in {32,64,128}; num_param 3 / stride 2 / C in {16,32,64},
 * even H and W), O % 16 == 0, O <= 256; ldconv_onepass_supported returns 1 for them, else use the entry points below. */
int ldconv_debug_l0_variant(int v);                     /* A/B of the first-layer kernel: 0 = tensor cores (default), 1 = CUDA-core rows kernel */
int ldconv_debug_l0_trace(void* device_buf);            /* clock64 stamps of CTA 0 of the next first-layer launches (benchmarks/l0_ab.py --trace) */
int ldconv_debug_onepass_trace(void* device_buf);      /* debug timeline of the next ldconv_onepass_fwd call (benchmarks/trace_onepass.py) */
int ldconv_onepass_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype);
int ldconv_onepass_fwd(const void* x, const void* w_offconv, const float* b_off, const int32_t* p_n, const void* wt,
                       const float* scale, const float* shift, void* out, int ldo, float* off_out, int B, int C, int H, int W,
                       int N, int s, int O, int act, int dtype, void* stream);

/* Everything AFTER the offset conv in one persistent kernel (conv.py:369-408, eval mode, bf16): sampling grid, clamps,
 * bilinear gather (TMA-staged tile + halo, L2 beyond it), the rearrange, the (N,1) conv as a tcgen05 GEMM whose operand tile
 * is written by the gather warps straight into shared memory (never to HBM), folded BatchNorm + activation.
 *   x (B,H,W,C) bf16; off (B,h,w,2N) fp32 (from ldconv_offset_conv_*); p_n (2N) int32; wt (O, N*C) bf16; scale/shift (O) fp32
 *   out (B,h,w,O) bf16 with pixel stride ldo elements (ldo = O for a dense tensor; a channel slice of a concat buffer else)
 * C % 8 == 0, (N*C) % 16 == 0, O % 16 == 0, O <= 256; ldconv_gather_gemm_supported returns 1 when the shape (and its
 * shared-memory plan) is covered -- otherwise use ldconv_gather_fwd + ldconv_gemm_fwd. */
int ldconv_gather_gemm_supported(int B, int C, int H, int W, int N, int s, int O, int ldo, int dtype);
int ldconv_gather_gemm_fwd(const void* x, const float* off, const int32_t* p_n, const void* wt, const float* scale,
                           const float* shift, void* out, int ldo, int B, int C, int H, int W, int N, int s, int O, int act,
                           int dtype, void* stream);

/* ---- neighbours of LDConv in the DEAL-YOLO graph, same kernels, same ABI conventions ------------
 * `Conv` = Conv2d(k, no bias) + BatchNorm2d + SiLU (nn/modules/conv.py:41-59) with the BatchNorm folded to scale/shift.
 * x / out / residual may be channel slices of wider NHWC buffers: ld* are PIXEL strides in elements (multiples of 8), so
 * C2f / SPPF (nn/modules/block.py:151-232) can write their branches straight into the concatenated buffer (no torch.cat).
 * 3x3: x (B,H,W,Cin|ldx) bf16, wt (Cout, 9*Cin) bf16 with k = (ky*3+kx)*Cin + c, out (B,h,w,Cout|ldo) bf16,
 *      out = act(conv*scale+shift) (+ residual).  Cin % 16 == 0, Cout % 16 == 0, stride 1 or 2, pad 1.
 * 1x1: rows = B*H*W pixels; out(rows, Cout|ldo) = act(x(rows, Cin|ldx) . wt(Cout,Cin)^T * scale + shift) (+ residual). */
int ldconv_conv1x1_bn_act_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                              const void* residual, int ldr, void* out, int ldo, long long rows, int Cin, int Cout, int act,
                              int dtype, void* stream);
/* Same block with a second, dense destination out2 (rows, c2_n | ld2) for the output channels [c2_lo, c2_lo + c2_n) (multiples of
 * 16; Cout % 16 == 0): C2f's y = cv1(x).chunk(2) (nn/modules/block.py:222-226) -- the chunk the first Bottleneck reads is written
 * densely in the same pass, because reading a channel slice of the NHWC concat buffer costs the whole buffer in DRAM traffic. */
int ldconv_conv1x1_bn_act_fwd2(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                               const void* residual, int ldr, void* out, int ldo, void* out2, int ld2, int c2_lo, int c2_n,
                               long long rows, int Cin, int Cout, int act, int dtype, void* stream);
/* The point-wise block of the finest SSFF level (ScalSeq, nn/extra_modules/block.py:3414-3443: Conv3d(1x1x1) + BatchNorm3d +
 * LeakyReLU(0.1) per level, nearest up-sampling of the two coarser levels, MaxPool3d((3,1,1)) over the levels) fused with that
 * maximum and with the following Add layer (block.py:3479-3484): out = max(bf16(act(x . wt^T * scale + shift)), up(z1), up(z2))
 * (+ residual); z1 (B,H1,W1,Cout), z2 (B,H2,W2,Cout) dense bf16 = the coarser levels after the same block.  Cout % 16 == 0. */
int ldconv_conv1x1_bn_act_maxup_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                                    const void* z1, int H1, int W1, const void* z2, int H2, int W2, const void* residual, int ldr,
                                    void* out, int ldo, int B, int H, int W, int Cin, int Cout, int act, int dtype, void* stream);
/* The 1x1 `Conv` block (nn/modules/conv.py:41-59) with P in {2, 4} consecutive pixels packed into one GEMM row: x (rows, Cin) dense
 * bf16, wt_packed (P*Cout, P*Cin) = block_diag(wt, ..., wt), scale_rep / shift_rep = the folded BatchNorm repeated P times,
 * out (rows, Cout | ldo).  Same arithmetic per output as ldconv_conv1x1_bn_act_fwd (the extra products are exact zeros); for the
 * narrow layers whose cost is the number of 128-row tiles, not bytes.  Cout a power of two >= 16, P*Cout <= 256, rows % P == 0. */
int ldconv_conv1x1_bn_act_packed_fwd(const void* x, const void* wt_packed, const float* scale_rep, const float* shift_rep, void* out,
                                     int ldo, long long rows, int Cin, int Cout, int P, int act, int dtype, void* stream);
int ldconv_conv3x3_supported(int Cin, int Cout, int stride, int dtype);
int ldconv_conv3x3_bn_act_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift,
                              const void* residual, int ldr, void* out, int ldo, int B, int Cin, int H, int W, int Cout,
                              int stride, int act, int dtype, void* stream);

/* Detect head decode of one pyramid level (nn/modules/head.py:55-77, DFL nn/modules/block.py:37-56, dist2bbox
 * utils/tal.py:309-319): box (B,H,W,4*reg_max) bf16 logits, cls (B,H,W,nc) bf16 logits ->
 * y (B, 4+nc, total_anchors) bf16 at columns [anchor_offset, anchor_offset + H*W): xywh * stride, sigmoid(cls). */
int ldconv_detect_decode(const void* box, const void* cls, void* y, int B, int H, int W, int nc, int reg_max, float stride,
                         int anchor_offset, int total_anchors, int dtype, void* stream);

/* Post-processing of the predictor / validator on the device (utils/ops.py:292-427 `non_max_suppression`, single label, no
 * masks, not rotated, with the fork's own `soft_nms`, ops.py:260-290, that line 407 calls): confidence filter in anchor order,
 * class-offset boxes, sequential soft-NMS with in-place score decay, max_det cut.  One CTA per image, one launch.
 *   y (B, 4+nc, A) bf16 / fp32 decoded head output (xywh pixels, class scores); out (B, max_det, 6) fp32 rows
 *   (x1, y1, x2, y2, conf, cls) in keep order; out_count (B) int32, negative = -(candidates) when an image has more than
 *   max_nms candidates (the confidence-sorted truncation of ops.py:395-396 is not implemented);
 *   workspace: ldconv_nms_workspace_bytes(B, min(A, max_nms)) bytes, 16-byte aligned. */
size_t ldconv_nms_workspace_bytes(int B, int max_nms);
int ldconv_nms(const void* y, void* out, int32_t* out_count, void* workspace, size_t workspace_bytes, int B, int A, int nc,
               float conf_thres, float iou_thres, int agnostic, int max_det, int max_nms, float max_wh, int dtype, void* stream);

/* The LAST 1x1 conv of a Detect branch with the decode in its epilogue (head.py:55-77, DFL block.py:37-56, dist2bbox tal.py:309-319):
 * mode 1 = box branch (Cout = 64 = 4 sides x 16 bins) -> y rows 0..3 (xywh * stride); mode 2 = class branch (Cout = nc <= 16) ->
 * y rows 4..4+nc (sigmoid).  x (B*H*W, Cin | ldx) bf16; y (B, 4+nc, total) bf16, this level at columns [a0, a0 + H*W).  Bit-identical
 * to ldconv_conv1x1_bn_act_fwd (act none) + ldconv_detect_decode; the logits never reach HBM. */
int ldconv_conv1x1_detect_fwd(const void* x, int ldx, const void* wt, const float* scale, const float* shift, void* y, int mode,
                              int B, int H, int W, int Cin, int Cout, int nc, float stride, int a0, int total, int dtype,
                              void* stream);

/* Task-aligned assigner of the training criterion (utils/tal.py:13-290), dense part, fp32:
 * ldconv_tal_metric: scores (B,na,nc) in [0,1], boxes (B,na,4) xyxy px, anchors (na,2) px, gt_labels (B,n) int32, gt_boxes (B,n,4)
 *   xyxy px, gt_valid (B,n) bytes -> align = score[label]^alpha * max(CIoU, 0)^beta and overlaps = max(CIoU, 0) on anchors
 *   strictly inside the gt box (tal.py:98-122, :226-243, utils/metrics.py:103-128), 0 elsewhere; both (B,n,na).
 * ldconv_tal_assign: topk_idx (B,n,k) int64 = torch.topk(align, k).indices -> foreground flags fg (B,na) bytes, assigned gt
 *   gt_idx (B,na) int64, align_sel (B,na) = metric of the assigned pair, pos_align / pos_over (B,n) = per-gt maxima over its
 *   positives (tal.py:124-157, :245-272, :83-88).  mask_ws: B*n*na bytes of workspace (zeroed by the call). */
int ldconv_tal_metric(const float* scores, const float* boxes, const float* anchors, const int32_t* gt_labels,
                      const float* gt_boxes, const unsigned char* gt_valid, float* align, float* overlaps, int B, int na, int n,
                      int nc, float alpha, float beta, float eps, void* stream);
int ldconv_tal_assign(const long long* topk_idx, const float* anchors, const float* gt_boxes, const unsigned char* gt_valid,
                      const float* align, const float* overlaps, unsigned char* mask_ws, unsigned char* fg, long long* gt_idx,
                      float* align_sel, float* pos_align, float* pos_over, int B, int na, int n, int k, float eps, void* stream);

/* Dense gradient-free decode of the raw head rows for the assigner (utils/loss.py:347-354 bbox_decode + pred_scores.sigmoid()):
 * x (rows = b*na, 64 + nc) bf16 / fp32; anc (na, 2) fp32 grid units -> boxes (rows, 4) fp32 xyxy grid units, scores (rows, nc) fp32. */
int ldconv_head_decode_rows(const void* x, const float* anc, float* boxes, float* scores, long long rows, int na, int nc,
                            int reg_max, int dtype, void* stream);

/* Glue ops of the graph (bf16 NHWC, channel-slice aware through the pixel strides ld*):
 * nearest up-sampling by an integer factor (yolov8-LD-P2.yaml:26,33); the SSFF tail = max over the three pyramid levels,
 * coarser levels indexed like torch's nearest interpolation, + optional residual (nn/extra_modules/block.py:3432-3443,
 * :3479-3484); SPPF's three chained k x k max-pools (nn/modules/block.py:166-171) as the k, 2k-1, 3k-2 window maxima. */
int ldconv_upsample_nearest(const void* x, int ldx, void* out, int ldo, int B, int H, int W, int C, int factor, int dtype,
                            void* stream);
/* backward of ldconv_upsample_nearest for the training graph: grad_x (B,H,W,C) = sum over each factor x factor block of grad_out
 * (B,H*factor,W*factor,C); fp32 sum, one rounding; ldg / ldx = pixel strides (grad_out may be a channel slice of the Concat's gradient). */
/* SSFF tail of the training graph (nn/extra_modules/block.py:3438-3443): `pre` (3 M, C) = the three depth slices of the Conv3d
 * output as rows, scale / shift (C) fp32 = the batch-statistics BatchNorm3d folded.  fwd: out (M, C) = max over the slices of
 * LeakyReLU(0.1)(pre * scale + shift).  bwd: dz (3 M, C) = the gradient w.r.t. the BatchNorm output: grad_out (M, C) * LeakyReLU'
 * on the first slice attaining the maximum (MaxPool3d's tie rule), zero on the others; feed it to ldconv_bn_act_bwd_* with act NONE. */
int ldconv_ssff_max_fwd(const void* pre, const float* scale, const float* shift, void* out, long long M, int C, int dtype, void* stream);
int ldconv_ssff_max_bwd(const void* pre, const float* scale, const float* shift, const void* grad_out, void* dz, long long M, int C,
                        int dtype, void* stream);
int ldconv_upsample_nearest_bwd(const void* grad_out, int ldg, void* grad_x, int ldx, int B, int H, int W, int C, int factor,
                                int dtype, void* stream);
/* `Add` rows (nn/extra_modules/block.py:3479-3484, torch.sum(torch.stack(x), 0)): out = sum of n <= 4 NHWC tensors / channel
 * slices (host arrays of device pointers and pixel strides), fp32 accumulation, one rounding. */
int ldconv_add_nhwc(const void* const* srcs, const int* lds, int n, void* out, int ldo, long long pixels, int C, int dtype,
                    void* stream);
int ldconv_scalseq_tail(const void* z0, const void* z1, const void* z2, const void* addend, int ld_add, void* out, int ldo,
                        int B, int H, int W, int H1, int W1, int H2, int W2, int C, int dtype, void* stream);
/* uint8 NCHW image batch -> `scale`-normalised bf16 NHWC (the predictor's `im.half(); im /= 255`, engine/predictor.py:120-131,
 * plus the layout change), one pass. */
int ldconv_image_u8_to_nhwc(const void* x_u8, void* out, int B, int C, int H, int W, float scale, int dtype, void* stream);
int ldconv_sppf_pools(const void* x, void* o1, void* o2, void* o3, int ld, int B, int H, int W, int C, int k, int dtype,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDCONV_B200_H */
