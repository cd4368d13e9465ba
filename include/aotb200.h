/* libaotb200.so -- C ABI of the B200-native AOT/DeAOT mask-propagation hot path.
 *
 * The reference (yoxu515/aot-benchmark @601c138) is pure Python/PyTorch and has no FFI of its
 * own; this is the boundary the drop-in Python engines bind with ctypes, and the one
 * a reference maintainer would bind (INTEGRATION.md shows the stub).  Each entry point replaces
 * the reference code cited beside it (paths relative to the reference root).
 *
 * Conventions: every pointer is a caller-owned DEVICE pointer (fp32 unless noted) obtained from
 * torch tensors via data_ptr(); no allocation, no host sync, no exceptions; `stream` is a
 * cudaStream_t (launch is capturable in a CUDA graph); returns 0 or a negative error code,
 * message via aotb_last_error_string().  Activations are NHWC / [rows][ld] row-major; an `ld*`
 * argument is the row (pixel) stride in elements so kernels can address channel slices.
 * Activation codes: 0 none, 1 ReLU, 2 GELU(erf), 3 SiLU, 4 ReLU6.
 */
#ifndef AOTB200_H
#define AOTB200_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

int aotb_version(void);
const char* aotb_arch(void);
const char* aotb_last_error_string(void);
/* kernels launched by this library in this process so far (bench.py reports the delta). */
unsigned long long aotb_launch_count(void);
/* Launch every kernel with programmatic dependent launch (prologues overlap the previous kernel's tail). */
void aotb_set_pdl(int on);
/* Tuning / diagnostic mask of aotb_conv2d_nhwc_tc (default 0); results are identical up to fp32 summation order.
 *   bit 0: the pre-cost-model heuristic (N = 64 tiles unless a wider tile fills the GPU on its own) instead of the
 *          fitted cost model over N tile x split-K cluster size;
 *   bit 1: mbarrier waits spin without the suspend hint;
 *   bit 2: every CTA writes clock64 stamps (0 start, 1 prologue done, 2 first A stage stored, 3 first stage
 *          consumable, 4 last MMA issued, 5 accumulator complete, 6 tile staged, 7 exit, 8 tile visible to the
 *          finish (cluster barrier), 9 finish stored, 10-11 unused) to `workspace` as long long[ctas][12];
 *   bits 4-7: force the N tile (1 = 64, 2 = 128, 3 = 256; 0 = policy); bits 8-11: force the split-K factor
 *          (1, 2, 4, 8; 0 = policy).  Forced values that do not divide the problem are an argument error. */
int aotb_set_conv_tiling(int mode);

/* nn.Conv2d (+ folded FrozenBatchNorm2d, + residual, + activation) as im2col-free implicit GEMM.
 * networks/encoders/resnet.py:34-54,140-157; networks/layers/normalization.py:30-43;
 * networks/models/aot.py:19-21,83; networks/decoders/fpn.py:34-58.
 * in [B][H][W][ldin], w [KH*KW*Cin][Cout], out [B][Ho][Wo][ldout], res like out with ldres. */
int aotb_conv2d_nhwc_f32(const float* in, const float* w, const float* bias, const float* res, float* out,
                         int B, int H, int W, int Cin, int ldin, int Cout, int ldout, int ldres,
                         int KH, int KW, int stride, int pad, int dil, int act, void* stream);

/* Same contract as aotb_conv2d_nhwc_f32 (dilation 1) on the tcgen05 tensor cores, fp32-faithful through split-fp16
 * operands: wh / wl are the weights pre-split as hi = fp16(w), lo = fp16(w - hi), laid out [Cout][KH*KW*Cin] (K-major,
 * K ordered (ky,kx,ci), zero-padded to a multiple of 64); activations are split on the fly.
 * Requires Cin % 4 == 0 and Cout % 64 == 0.  Few-tile deep-K layers run split-K: the 2 / 4 / 8 CTAs of one output
 * tile form a thread-block cluster and sum their partial tiles over distributed shared memory in rank order
 * (deterministic).  `workspace` / `workspace_bytes` are only used by the diagnostic mode of aotb_set_conv_tiling
 * (may be NULL / 0 otherwise). */
int aotb_conv2d_nhwc_tc(const float* in, const void* wh, const void* wl, const float* bias, const float* res,
                        float* out, int B, int H, int W, int Cin, int ldin, int Cout, int ldout, int ldres,
                        int KH, int KW, int stride, int pad, int act, void* workspace, size_t workspace_bytes,
                        void* stream);

/* nn.Linear on tokens: out[M][N] = act(in[M][K] @ wt[K][N] + bias + res).
 * networks/layers/transformer.py:321-367,582-665; networks/layers/attention.py:76-79,119,710,858. */
int aotb_linear_f32(const float* in, const float* wt, const float* bias, const float* res, float* out,
                    int M, int K, int ldin, int N, int ldout, int ldres, int act, void* stream);

/* Layout changes at the API edge (callers hand NCHW images, read NCHW features). */
int aotb_nchw_to_nhwc_f32(const float* in, float* out, int B, int C, int HW, void* stream);
int aotb_nhwc_to_nchw_f32(const float* in, float* out, int B, int C, int HW, void* stream);
/* [3][HW] image -> [HW][4] NHWC with a zero 4th channel (16-byte pixels for the stem convolution). */
int aotb_image_to_nhwc4_f32(const float* in, float* out, int HW, void* stream);

/* nn.MaxPool2d(3, 2, 1): networks/encoders/resnet.py:79,146. */
int aotb_maxpool3x3s2_nhwc_f32(const float* in, float* out, int B, int H, int W, int C, void* stream);

/* Depthwise conv, w [KH*KW][C]: networks/layers/basic.py:15-57 (5x5 of the FFN / gated propagation),
 * networks/encoders/mobilenetv2.py:93-101 (3x3 + folded BN + ReLU6). */
int aotb_dwconv_nhwc_f32(const float* in, const float* w, const float* bias, float* out, int B, int H, int W,
                         int C, int ldin, int ldout, int KH, int KW, int stride, int pad, int dil, int act,
                         void* stream);

/* F.interpolate(mode="bilinear", align_corners=...): networks/decoders/fpn.py:45-54. */
int aotb_bilinear_nhwc_f32(const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo,
                           int align_corners, void* stream);

/* Strided element-wise: op 0 copy, 1 a+b, 2 a*b, 3 silu(a), 4 silu(a)*b, 5 fill(scalar).
 * networks/layers/attention.py:585-586,707,855; networks/layers/transformer.py:602-611,625-626. */
int aotb_eltwise_f32(int op, const float* a, int lda, const float* b, int ldb, float* out, int ldo,
                     int rows, int cols, float scalar, void* stream);

/* nn.LayerNorm(C) per row; if out2 != NULL also out2 = LN(x) + add (with_pos_embed,
 * networks/layers/transformer.py:305-310,321-322). */
int aotb_layernorm_f32(const float* x, int ldx, const float* gamma, const float* beta, const float* add,
                       int ldadd, float* out, int ldo, float* out2, int ldo2, int rows, int C, void* stream);

/* Swin (shifted-)window multi-head self-attention core, batch 1 (BASELINE config 4 encoder):
 * WindowAttention.forward networks/encoders/swin/swin_transformer.py:158-196 fused with the zero padding, cyclic
 * shift, window partition / reverse and crop of SwinTransformerBlock.forward :273-316 and the shifted-window mask of
 * BasicLayer.forward :416-438.  qkv [H*W][ldqkv] = the qkv Linear applied to the UN-padded norm1 output, columns
 * [q(C) | k(C) | v(C)], head h = columns h*32..h*32+31 of each; qkv_bias [3C] stands in for padded positions (the
 * reference pads after norm1, so a padded token's q/k/v are the bias); rel_bias [heads][49][49] =
 * relative_position_bias_table[relative_position_index] permuted (:176-183); out [H*W][ldo] = softmax(q k^T / sqrt(32)
 * + rel_bias + mask) v per head, heads concatenated, before `proj`.  window must be 7, C == heads * 32. */
int aotb_window_attention_f32(const float* qkv, int ldqkv, const float* qkv_bias, const float* rel_bias, float* out,
                              int ldo, int H, int W, int C, int heads, int window, int shift, void* stream);

/* PatchMerging gather (swin_transformer.py:339-360): x [H*W][ldx] (C channels) -> out [ceil(H/2)*ceil(W/2)][ldo] with
 * 4C channels ordered [(0,0) | (1,0) | (0,1) | (1,1)] of each 2x2 block, zeros outside H x W. */
int aotb_patch_merge_f32(const float* x, int ldx, float* out, int ldo, int H, int W, int C, void* stream);

/* nn.GroupNorm(G, C) over [B][P pixels][C] + activation: networks/layers/basic.py:6-12,18,30-32,75-85.  The workspace
 * (aotb_groupnorm_workspace_bytes(B, G) bytes, reusable for any smaller G) must be zero-filled before its first use: it holds
 * the launch counter with which the last block of the statistics kernel finalises (mean, rstd) once, in a fixed order. */
size_t aotb_groupnorm_workspace_bytes(int B, int G);
int aotb_groupnorm_nhwc_f32(const float* x, int ldx, const float* gamma, const float* beta, float* out, int ldo,
                            int B, int P, int C, int G, int act, void* workspace, void* stream);

/* softmax((Q/T) K^T) V per head, scores never materialised (fp32 reference-precision path).
 * networks/layers/attention.py:82-117 (MultiheadAttention) and :672-704 (GatedPropagation).
 * Tk_dev (optional) is a device int holding the live key count.  With Mout/Lout the kernel writes
 * the un-normalised split-KV partial (row max, row sum, O) for aotb_attn_merge_f32. */
int aotb_attention_f32(const float* Q, int ldq, const float* K, int ldk, const float* V, int ldv, float* O,
                       int ldo, int N, int Tk, const int* Tk_dev, int H, int d_qk, int d_v, float* Mout,
                       float* Lout, void* stream);
int aotb_attn_merge_f32(const float* Opart, const float* Mpart, const float* Lpart, float* O, int R, int N,
                        int H, int d_v, int ldo, void* stream);

/* The same merge with every rank's partials read in place over peer memory (sharded long-term bank, BASELINE configs[3]):
 * Oparts / Mparts / Lparts are HOST arrays of `ranks` (<= 8) device pointers -- the local buffer and the NVLink peer mappings
 * of a symmetric-memory allocation -- to Opart_r [splits][N][H*d_v] and Mpart_r / Lpart_r [splits][H][N]; the exchange step of
 * the split-KV attention (otherwise three NCCL all-gathers per layer) is the P2P loads of this kernel.  Partials are merged in
 * (rank, split) order on every rank: outputs are bit-identical across ranks. */
int aotb_attn_merge_peers_f32(const void* const* Oparts, const void* const* Mparts, const void* const* Lparts, int ranks,
                              int splits, float* O, int N, int H, int d_v, int ldo, void* stream);

/* 15x15 local-window attention with relative_emb_k / relative_emb_v:
 * networks/layers/attention.py:308-428 (MultiheadLocalAttentionV2) and :789-914 (LocalGatedPropagation);
 * replaces the third-party spatial_correlation_sampler call sites :341,:828. */
int aotb_local_attention_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                             const float* relk_w, const float* relk_b, const float* relv, float* out, int ldo,
                             int h, int w, int H, int d_att, int d_v, void* stream);

/* Same computation for the AOT head shape (d_att = d_v = 32) with the K / V window halos staged in shared
 * memory per 8x8 query tile; relv_t is relative_emb_v transposed to [H][225][32]. */
int aotb_local_attention_tile_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                                  const float* relk_w, const float* relk_b, const float* relv_t, float* out, int ldo,
                                  int h, int w, int H, void* stream);

/* Same computation for the DeAOT head shape (one head, d_att 128, d_v 1024, no relative_emb_v; attention.py:789-861) with the
 * window halos staged in shared memory per 8x6 query tile and the channels walked in chunks of 32. */
int aotb_local_gated_tile_f32(const float* q, int ldq, const float* k, int ldk, const float* v, int ldv,
                              const float* relk_w, const float* relk_b, float* out, int ldo, int h, int w, void* stream);

/* one_hot_mask + patch_wise_id_bank conv as a gather-sum (+ LayerNorm for DeAOT):
 * utils/image.py:69-74; networks/models/aot.py:50-63,76-79; networks/models/deaot.py:51-55.
 * mask [Hm][Wm] float ids; wt [(ky*KS+kx)*nid + id][C]. */
int aotb_id_embed_f32(const float* mask, int Hm, int Wm, const float* wt, const float* bias,
                      const float* ln_gamma, const float* ln_beta, float* out, int ldo, int C, int nid,
                      int ksize, int stride, int pad, void* stream);

/* Same result through run lengths: wp = exclusive prefix sums of the table along kx, [ksize][ksize+1][nid][C=256]. */
int aotb_id_embed_runs_f32(const float* mask, int Hm, int Wm, const float* wp, const float* bias,
                           const float* ln_gamma, const float* ln_beta, float* out, int ldo, int C, int nid,
                           int ksize, int stride, int pad, void* stream);

/* networks/engines/aot_engine.py:367-378: mask ids > obj_num with -1e10, bilinear upsample to NCHW. */
int aotb_logits_postproc_f32(const float* logits_nhwc, float* lowres_nchw, float* out_nchw, int h, int w,
                             int NC, int obj_num, int Ho, int Wo, int align_corners, void* stream);
/* fused upsample + argmax (networks/managers/evaluator.py:339-361 for one engine, no TTA). */
int aotb_logits_argmax_f32(const float* lowres_nchw, float* label, int h, int w, int NC, int Ho, int Wo,
                           int align_corners, void* stream);
/* networks/engines/aot_engine.py:565-582 (AOTInferEngine.soft_logit_aggregation) for n_engines sub-engines of max_obj (= 10)
 * objects each, fused into one pass: logits[e] -> NCHW fp32 [1 + max_obj][HW] on the device (the array of pointers itself is
 * host memory); out [1 + n_engines * max_obj][HW] = logit(clamp([prod_e softmax_e[0], softmax_0[1:], softmax_1[1:], ...])). */
int aotb_soft_logit_aggregation_f32(const float* const* logits, int n_engines, int max_obj, float* out, int HW,
                                    void* stream);
/* networks/engines/aot_engine.py:515-533 (AOTInferEngine.separate_mask, label-map form): out[e][i] = mask[i] - e*max_obj if
 * e*max_obj < mask[i] <= (e+1)*max_obj else 0, for e in [0, n_engines). */
int aotb_separate_labels_f32(const float* mask, int n_engines, int max_obj, float* out, int HW, void* stream);
/* A chain of tensor-core convolutions as ONE persistent kernel with tile-level dataflow (csrc/conv_chain.cu): replaces the
 * per-layer launches of networks/encoders/resnet.py:34-54,140-157 (Bottleneck stacks, FrozenBN folded) by a program of 128-pixel
 * tiles that start as soon as the tiles of the producing layer they read are complete.  Every layer follows the contract of
 * aotb_conv2d_nhwc_tc (fp32 NHWC in / out, pre-split fp16 weights [Cout][KH*KW*Cin], bias, optional residual, activation) with
 * Cin % 64 == 0 and Cout % 64 == 0; in_layer / res_layer name the chain layer that produces `in` / `res` (-1: complete before
 * the launch).  Every output buffer must be written by exactly one layer of the chain (the SMs' L1 caches are not coherent:
 * no address may change its value twice inside one launch).  aotb_conv_chain_plan sizes the device-resident program,
 * aotb_conv_chain_build writes it (once per geometry, outside stream capture), aotb_conv_chain_run clears the dependency
 * counters and launches the kernel (capturable); aotb_conv_chain_dump exposes the tile program to host-side tests.  Layers with
 * few tiles and a long K loop are cut into up to 4 split-K work items; the item that holds the last K range adds the others'
 * partial tiles (kept in a scratch area of the program buffer) in split order, so results are deterministic. */
typedef struct aotb_chain_layer {
    const float* in;
    const void* wh;
    const void* wl;
    const float* bias;
    const float* res;
    float* out;
    int H, W, Cin, ldin, Cout, ldout, ldres, KH, KW, stride, pad, act, in_layer, res_layer;
} aotb_chain_layer;
int aotb_conv_chain_plan(const void* layers, int nlayers, size_t* program_bytes, int* ntiles, int* ncounters);
int aotb_conv_chain_dump(const void* layers, int nlayers, int* tiles8, int max_tiles, int* layers10);
int aotb_conv_chain_build(const void* layers, int nlayers, void* program, size_t program_bytes, void* stream);
int aotb_conv_chain_run(void* program, int nlayers, int ntiles, int ncounters, void* stream);
size_t aotb_conv_chain_prof_offset(int nlayers, int ntiles, int ncounters);
/* Frame input side: dataloaders/eval_datasets.py:60-61 + dataloaders/video_transforms.py:594-715 (MultiRestrictSize's
 * cv2.resize(INTER_CUBIC) of the float image, MultiToTensor's / 255, - mean, / std, HWC -> CHW) on the uint8 frame in one pass.
 * img uint8 [H][W][3]; ix / cx [Wo][4] and iy / cy [Ho][4] = clamped tap indices and Keys-cubic (A = -0.75) weights per output
 * column / row (all four null when Ho == H and Wo == W); flip != 0 mirrors horizontally after the resize; out fp32 [3][Ho][Wo]. */
int aotb_preprocess_bgr_u8(const void* img, int H, int W, const int* ix, const float* cx, const int* iy, const float* cy,
                           float* out, int Ho, int Wo, int flip, void* stream);
/* Mask output side: utils/image.py:103-105 (`mask_tensor.cpu().numpy().astype('uint8')`): the float label map leaves the
 * device as uint8 (1 byte per pixel over PCIe instead of 4 or 8). */
int aotb_label_to_u8(const float* label, void* out_u8, int n, void* stream);
/* F.interpolate(mode="nearest") of a label map: networks/managers/evaluator.py:418-421. */
int aotb_nearest_resize_f32(const float* in, float* out, int H, int W, int Ho, int Wo, void* stream);

/* Tensor-core long-term attention (tcgen05 + TMEM + TMA), AOT head shape H x 32, split-fp16 ("fp16x2")
 * operands: every fp32 value x is stored as hi = fp16(x), lo = fp16(x - hi) in rows [hi(32) | lo(32)].
 * networks/layers/attention.py:82-117 called from networks/layers/transformer.py:346 (and :324, Tk = N).
 *   aotb_tc_pack_rows_f16x2: fp32 [rows][ld] -> packed [H][cap][64] at a row offset, values / div first
 *                            (div = T for Q, attention.py:82; 1 for K and V).  Buffers must be zero-filled
 *                            beyond the live rows.
 *   aotb_lt_attn_tc_f16x2  : exact bit 0 set -> S = QhKh + QlKh + QhKl, O = (Ph + Pl)[Vh|Vl] (fp32-faithful);
 *                            clear -> S = QhKh, O = Ph[Vh|Vl].  exact bit 1 selects the softmax layout: clear = all
 *                            16 softmax warps on one 128x128 score tile at a time (4 threads per query row); set =
 *                            two groups of 8 warps, one per query tile, out of phase (2 threads per row, one TMEM
 *                            read per tile).  Same arithmetic; row sums are associated differently (~1e-7 relative).
 *                            exact bit 3: "ahead" layout -- three score buffers in TMEM (P_hi and P_lo both alias the
 *                            owning thread's score columns), S issued one tile ahead, the TMEM read of tile n+1 under the
 *                            ex2 pass of tile n (experimental: built at the end of round 1, not yet run on a GPU).
 *                            exact bit 2: the mbarrier waits between the softmax warps and the MMA issuer poll instead
 *                            of sleeping (latency experiment; results unchanged).  splits > 1 writes split-KV partials
 *                            for aotb_attn_merge_f32.  dbg (optional) receives S and O' of CTA 0. */
int aotb_tc_pack_rows_f16x2(const float* src, int ld, void* dst, int cap, int rows, int H, int row_off,
                            const int* row_off_dev, float div, void* stream);
size_t aotb_lt_attn_tc_smem_bytes(void);
int aotb_lt_attn_tc_f16x2(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N, int Tk,
                          const int* Tk_dev, int H, float* O, int ldo, float* Opart, float* Mpart, float* Lpart,
                          int splits, int exact, float* dbg, void* stream);

/* DeAOT long-term attention as GEMM -> row softmax -> GEMM on the tensor cores (AOTB_DEAOT_LT=gemm):
 * GatedPropagation.forward networks/layers/attention.py:672-704 with 1 head, d_qk = 128, d_v = 1024.  The two GEMMs are
 * aotb_conv2d_nhwc_tc with the bank as pre-split weights; these entry points maintain those copies and do the softmax.
 *   aotb_split_rows_f16x2: src fp32 [rows][lds] (C columns) -> hi / lo fp16 [.][ldw] rows [row_off, row_off + rows)
 *                          (hi = fp16(x), lo = fp16(x - hi)); the keys [Tk_cap][128] = weights of S = Q K^T.
 *   aotb_split_cols_f16x2: the same values written as COLUMNS [col_off, col_off + rows) of hiT / loT [C][ldt]: the
 *                          transposed value bank [1024][Tk_cap] = weights of O = P V.
 *   aotb_row_softmax_f32 : S [N][ld] in place: columns [0, live) <- softmax(scale * S[r][0:live]) (attention.py:686-693,
 *                          scale = 1 / T applied to the scores instead of to Q), columns [live, cols) <- 0; live = *Tk_dev
 *                          if given, else Tk.  Offsets / counts may be device-resident so a captured graph can be replayed. */
int aotb_split_rows_f16x2(const float* src, int lds, void* hi, void* lo, int ldw, int rows, int C, int row_off,
                          const int* row_off_dev, void* stream);
int aotb_split_cols_f16x2(const float* src, int lds, void* hiT, void* loT, int ldt, int rows, int C, int col_off,
                          const int* col_off_dev, void* stream);
int aotb_row_softmax_f32(float* S, int ld, int N, int cols, int Tk, const int* Tk_dev, float scale, void* stream);

/* DeAOT long-term attention fused on the tensor cores (EXPERIMENTAL, AOTB_DEAOT_LT=tc; gp_attn_tc.cu): GatedPropagation.forward
 * networks/layers/attention.py:672-704 with 1 head, d_qk = 128, d_v = dv (1024).  Operands in the split-fp16 row format of
 * aotb_tc_pack_rows_f16x2 with one "head" per 32 channels: Qp [4][Nq_cap][64] (Q / T, T = sqrt(128), Nq_cap a multiple of 128),
 * Kp [4][kv_cap][64], Vp [dv/32][kv_cap][64].  O [N][ldo] = softmax((Q / T) K^T) V.  exact bit 0 / bit 2 as in
 * aotb_lt_attn_tc_f16x2.  splits > 1 writes un-normalised partials Opart [splits][N][dv], Mpart / Lpart [splits][1][N] for
 * aotb_attn_merge_f32 (H = 1, d_v = dv). */
int aotb_gp_attn_tc_f16x2(const void* Qp, int Nq_cap, const void* Kp, const void* Vp, int kv_cap, int N, int Tk,
                          const int* Tk_dev, int dv, float* O, int ldo, float* Opart, float* Mpart, float* Lpart,
                          int splits, int exact, void* stream);

/* Long-term memory append in place (replaces torch.cat, networks/engines/aot_engine.py:291-305). */
int aotb_bank_append_f32(const float* src, int lds, float* bank, int ldb, int rows, int cols, int offset,
                         const int* offset_dev, void* stream);
int aotb_counter_add(int* counter, int delta, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AOTB200_H */
