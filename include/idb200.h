/*
 * idb200.h -- C ABI of libidb200.so: the B200 (sm_100a) implementation of the batched
 * generation-and-corruption hot path of EquilibriaW/Interpolated_Diffusion.
 *
 * The reference has no FFI layer (it is pure PyTorch); every entry point below
 * names the reference Python function (file:line under the reference root) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer of the reference would add.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes; every pointer is a DEVICE pointer unless marked "host";
 *   - the caller owns and allocates every buffer; the library never allocates device memory,
 *     never synchronises the device and only enqueues work on `stream` (so every call is
 *     CUDA-graph capturable);
 *   - tensors are dense, row-major, in the reference's own layouts ([B,T,D] trajectories,
 *     [B,K] int64 indices, uint8 0/1 for torch.bool);
 *   - return value: 0 on success, a negative IDB200_E* code otherwise; idb200_last_error()
 *     returns a thread-local human readable message for the last failure.
 */
#ifndef IDB200_H
#define IDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* idb200_stream_t; /* a cudaStream_t */

#define IDB200_OK 0
#define IDB200_EINVAL (-1)      /* bad shape / size / flag */
#define IDB200_EALIGN (-2)      /* pointer not aligned as documented */
#define IDB200_EUNSUPPORTED (-3)/* valid but outside what the kernels cover (e.g. T > 256) */
#define IDB200_ECUDA (-4)       /* cudaGetLastError() after launch */

int idb200_version(void);
const char* idb200_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * K1  nested anchor masks + Interp(x0 | M_s)
 *     replaces src/corruptions/keyframes.py:172-209 (build_nested_masks_batch: argsort of the
 *     rand scores, per-level cat/sort/scatter_) fused with :348-380 (interpolate_from_indices) as
 *     used by src/train/train_interp_levels.py:227-383.  Also serves :42-81
 *     (sample_fixed_k_indices_batch; n_levels = 1) and :260-294 (build_nested_masks_from_logits;
 *     IDB200_F_DESCENDING with scores = logits + 1, score_stride = T).
 *
 *   scores   [B, n] fp32 with row stride `score_stride` elements; n = T-2 (interior positions
 *            1..T-2) or n = T with IDB200_F_NO_ENDPOINTS.  Finite values only.
 *   K_list   host, n_levels ints: anchors per level (level 0 finest).  Level s marks t in {0,T-1}
 *            plus the K_s-2 interior positions of lowest stable rank
 *            rank(j) = #{u : s_u < s_j or (s_u == s_j and u < j)}   (ties: lower index first).
 *   masks    [B, n_levels, T] uint8 (0/1) or NULL.
 *   idx_out  NULL, or int64 buffer holding for level s a dense [B, W_s] block at element offset
 *            B * (W_0 + ... + W_{s-1}), W_s = (K_s <= 2 || T <= 2) ? 2 : K_s  (endpoint mode),
 *            W_s = K_s (IDB200_F_NO_ENDPOINTS): ascending anchor positions.
 *   x0       [B, T, D] fp32, 16-byte aligned, D in {2, 4}; NULL = masks / idx only.
 *   x_levels output for levels s_lo..s_hi: level s is a dense [B, T, D] block at element offset
 *            (s - s_lo) * level_stride.  Interp(x0 | M_s): anchors copied exactly, interior
 *            y = v_l + ((t - i_l) / max(i_r - i_l, 1)) * (v_r - v_l), every fp32 op rounded
 *            separately (bit-identical to the reference's eager ops).
 *   IDB200_F_RECOMPUTE_VELOCITY (D == 4): v[t] = (p[t+1] - p[t]) / (1/T), v[T-1] = 0
 *            (keyframes.py:373-379).
 * ---------------------------------------------------------------------------------------------- */
#define IDB200_F_DESCENDING 1
#define IDB200_F_RECOMPUTE_VELOCITY 2
#define IDB200_F_NO_ENDPOINTS 4

int idb200_nested_masks_interp(const float* x0, const float* scores, int64_t score_stride, int64_t B, int T,
                               int D, int n_levels, const int* K_list, uint8_t* masks, int64_t* idx_out,
                               float* x_levels, int64_t level_stride, int s_lo, int s_hi, int flags,
                               idb200_stream_t stream);

/* interpolate_from_indices, src/corruptions/keyframes.py:348-380, general form:
 *   idx [B,K] int64 ascending (duplicates allowed, endpoints not required: linear extrapolation
 *   outside [idx[0], idx[K-1]] exactly as the reference's clamp of `seg` to [0, K-2] does),
 *   vals [B,K,D] fp32, y [B,T,D] fp32.  Any D >= 1, T <= 4096, 2 <= K <= T.  */
int idb200_interpolate_from_indices(const int64_t* idx, const float* vals, int64_t B, int K, int T, int D,
                                    int recompute_velocity, float* y, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K1c  training corruption, src/train/train_interp_levels.py:458-510 (_corrupt_from_anchors) with
 *      :444-455 (_distance_alpha): gather anchors (optionally index-jittered), add anchor noise on
 *      the position dims (not on endpoints), interpolate, add tent-weighted noise, optionally
 *      recompute velocity.  Noise tensors are INPUTS (drawn by the caller with the reference's
 *      generator calls) so results are bit-identical.
 *   source [B,T,D]; idx [B,K] int64; idx_gather [B,K] int64 or NULL (= idx);
 *   anchor_noise [B,K,2] or NULL; path_noise [B,T,2] or NULL; mode_dist: 1 = "dist" tent weight.
 *   row_index  NULL, or int64 [n]: row r of idx/noise corresponds to source / out row row_index[r]
 *              (the boolean-mask gather/scatter of :328-382 without materialising it); then B = n.
 * ---------------------------------------------------------------------------------------------- */
int idb200_corrupt_from_anchors(const float* source, const int64_t* idx, const int64_t* idx_gather,
                                const float* anchor_noise, const float* path_noise, const int64_t* row_index,
                                int64_t B, int K, int T, int D, float sigma, float anchor_sigma, int mode_dist,
                                int clamp_endpoints, int recompute_velocity, float* out,
                                idb200_stream_t stream);

/* K1c, single launch for the whole batch: build_interp_adjacent_batch / build_interp_level_batch,
 * src/train/train_interp_levels.py:227-383.  Row b uses level s = s_idx[b]: x_s = corrupt(Interp(source | M_s)), and when
 * x_prev != NULL x_prev = corrupt(Interp(source | M_{s-1})); mask_s / mask_prev (uint8 [B,T], may be NULL) are the rows
 * M_s / M_{s-1}.  Rows with s outside [1, n_levels) are zero-filled (the reference's loop skips them).  The anchor lists
 * are read off masks_levels [B, n_levels, T] (ascending set bits == idx_levels[s]).  sigma_levels / anchor_sigma_levels:
 * HOST arrays [n_levels] (per-level sigma of :386-401 and sigma * corrupt_anchor_frac).  Noise:
 *   parity mode  anchor_noise [B,2,Kmax,2] and path_noise [B,2,T,2] fp32 (slot 0 = x_s, slot 1 = x_prev), drawn by the caller;
 *   Philox mode  both NULL: N(0,1) from Philox4x32-10 + Box-Muller keyed by (seed, offset, row, slot, kind, position);
 *                anchor_noise_out / path_noise_out (same layouts, may be NULL) export what was drawn.
 * Index jitter (:471-483) is not handled here (callers with corrupt_index_jitter_max > 0 use idb200_corrupt_from_anchors). */
int idb200_corrupt_adjacent(const float* source, const uint8_t* masks_levels, const int64_t* s_idx, int64_t B, int T, int D,
                            int n_levels, const float* sigma_levels, const float* anchor_sigma_levels,
                            const float* anchor_noise, const float* path_noise, int Kmax, uint64_t seed, uint64_t offset,
                            float* anchor_noise_out, float* path_noise_out, int mode_dist, int clamp_endpoints,
                            int recompute_velocity, float* x_s, float* x_prev, uint8_t* mask_s, uint8_t* mask_prev,
                            idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2  DDIM update (+ known-value clamp), src/diffusion/ddpm.py:37-48 (eta = 0) fused with the
 *     torch.where at src/sample/sample_generate.py:397-399:
 *       x0 = (z - sqrt(1-ab_t)*eps) / sqrt(ab_t);  z' = sqrt(ab_p)*x0 + sqrt(1-ab_p)*eps
 *       z' = known_mask ? known_values : z' ;  optional clip of dims 0:2 (:382-386)
 *     every op rounded separately, IEEE sqrt/div.  n_rows rows of `row_len` elements; the
 *     timestep of row i is t[i] / t_prev[i] (int64, gathered from alpha_bar[n_train]) or, when
 *     t == NULL, the two scalars ab_t / ab_prev.  known_mask (uint8) / known_values may be NULL.
 *     D is the innermost (feature) size used by pos_clip.  z_out may alias z.
 * ---------------------------------------------------------------------------------------------- */
int idb200_ddim_step(const float* z, const float* eps, const int64_t* t, const int64_t* t_prev,
                     const float* alpha_bar, int n_train, float ab_t, float ab_prev, int64_t n_rows,
                     int64_t row_len, int D, const uint8_t* known_mask, const float* known_values,
                     int pos_clip, float clip_min, float clip_max, float* z_out, idb200_stream_t stream);

/* q_sample, src/diffusion/ddpm.py:15-24: out = sqrt_ab[t]*r0 + sqrt_1m_ab[t]*noise (per-row t). */
int idb200_q_sample(const float* r0, const float* noise, const int64_t* t, const float* sqrt_ab,
                    const float* sqrt_1m_ab, int n_train, int64_t n_rows, int64_t row_len, float* out,
                    idb200_stream_t stream);

/* _build_known_mask_values, src/sample/sample_generate.py:260-280 (+ logit_pos of
 * src/utils/normalize.py:4-11 when logit_space != 0, as at sample_generate.py:1021-1022).
 *   idx [B,K] int64, start_goal [B,4]; known_mask [B,K,D] uint8, known_values [B,K,D] fp32. */
int idb200_known_mask_values(const int64_t* idx, const float* start_goal, int64_t B, int K, int D, int T,
                             int clamp_endpoints, int logit_space, float logit_eps, uint8_t* known_mask,
                             float* known_values, idb200_stream_t stream);

/* logit_pos / sigmoid_pos, src/utils/normalize.py:4-20: dims 0:2 of the innermost axis. */
int idb200_logit_pos(const float* x, int64_t n_rows, int D, float eps, float* out, idb200_stream_t stream);
int idb200_sigmoid_pos(const float* x, int64_t n_rows, int D, float* out, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K2b Stage-2 epilogue, src/sample/sample_generate.py:1252-1285 / :1160-1204 with
 *     src/utils/clamp.py:4-32:
 *       x = x_in (+ delta if delta != NULL)
 *       soft:  x[dims] += (conf*lam) * (x_ref[dims] - x[dims])      (conf != NULL and lam > 0)
 *       hard:  x[dims]  = clamp ? x_ref[dims] : x[dims]             (policy)
 *       clip:  x[0:2]   = min(max(x, lo), hi)                       (pos_clip)
 *     dims = 0:2 ("pos", clamp_dims_all == 0) or all D.  policy: 0 none, 1 endpoints (t in
 *     {0,T-1}), 2 mask (clamp_mask [B,T] uint8).  out may alias x_in.
 * ---------------------------------------------------------------------------------------------- */
#define IDB200_CLAMP_NONE 0
#define IDB200_CLAMP_ENDPOINTS 1
#define IDB200_CLAMP_MASK 2

int idb200_stage2_epilogue(const float* x_in, const float* delta, const float* x_ref, const float* conf,
                           float lam, int policy, const uint8_t* clamp_mask, int clamp_dims_all, int pos_clip,
                           float clip_min, float clip_max, int64_t B, int T, int D, float* out,
                           idb200_stream_t stream);

/* _build_anchor_conf + _anneal_conf, src/sample/sample_generate.py:319-360 (train twin
 * src/train/train_interp_levels.py:546-576), and the mask_in stack at :1163-1178 / :1255-1256.
 *   mask_s [B,T] uint8; student [B,T] uint8 or NULL; s_row int64 [B] or NULL (then scalar s);
 *   anneal_mode 0 none, 1 linear, 2 cosine.  conf [B,T] fp32 (may be NULL when only mask_in is
 *   wanted).  mask_in [B,T,C] fp32 or NULL: C == 2 -> [mask_s, conf]; C == 3 -> [mask_s,
 *   mask_prev, conf] (mask_prev [B,T] uint8 required). */
int idb200_anchor_conf(const uint8_t* mask_s, const uint8_t* student, const uint8_t* mask_prev,
                       const int64_t* s_row, int s_scalar, int levels, int anneal_mode, float conf_teacher,
                       float conf_student, float conf_endpoints, float conf_missing, int clamp_endpoints,
                       int64_t B, int T, int C, float* conf, float* mask_in, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * K3a  token GEMM on tcgen05 / TMEM fed by TMA:  out[M,N] = epi(A[M,K] * W[N,K]^T + bias[N])
 *      The dense contractions of src/models/transformer.py:35-46 (packed QKV, out_proj, ff.0, ff.2)
 *      and the in/out projections of the two denoisers (nn.Linear weight layout [out, in] = [N, K]).
 *   A, W: bf16 row-major, 16-byte aligned; K % 64 == 0, N % 32 == 0; M arbitrary.
 *   epilogue: 0 bf16 store, 1 SiLU then bf16 store (ff.0), 2 fp32 residual accumulate
 *             out[M,N] += acc + bias (out_proj / ff.2 into the residual stream), 3 fp32 store;
 *             0 | IDB200_EPI_OUT_F16: the 16-bit store is IEEE half instead of bf16 (the FiLM table of idb200_encoder_fused).
 * ---------------------------------------------------------------------------------------------- */
#define IDB200_EPI_BF16 0
#define IDB200_EPI_SILU_BF16 1
#define IDB200_EPI_RESID_F32 2
#define IDB200_EPI_F32 3
#define IDB200_EPI_OUT_F16 0x100

int idb200_gemm_bf16(const void* A, const void* W, const float* bias, void* out, int64_t M, int N, int K, int epilogue,
                     idb200_stream_t stream);
/* Training forms of the same GEMM with a second bf16 [M,N] tensor (N % 64 == 0), replacing the separate SiLU passes of the MLP
 * (transformer.py:43-45) in the forward / backward of the training step, bit-identical to them:
 *   epilogue 4: out = bf16(acc + bias) (pre-activation u), aux = bf16(SiLU(out))            -- ff.0 forward, both kept
 *   epilogue 5: out = bf16(bf16(acc) * SiLU'(aux)), aux = u (read through TMA)               -- dU = (dY W2) . SiLU'(u) */
#define IDB200_EPI_BF16_SILU_DUAL 4
#define IDB200_EPI_BF16_DSILU 5
int idb200_gemm_bf16_aux(const void* A, const void* W, const float* bias, void* out, void* aux, int64_t M, int N, int K, int epilogue,
                         idb200_stream_t stream);
/* Epilogue 5 that also leaves colpart [8 * ceil(M / 256), N] fp32 (4 rows per 128-row block, blocks rounded up to whole CTA pairs; every
 * row is written): per 128-row block and epilogue warp, the column sums of the bf16-rounded dU rows of that warp (zeros for rows >= M).
 * Column sums of colpart = column sums of dU = the ff.0 bias gradient (transformer.py:23). */
int idb200_gemm_bf16_dsilu_sums(const void* A, const void* W, void* out, void* aux, float* colpart, int64_t M, int N, int K,
                                idb200_stream_t stream);

/* Implicit-GEMM 3x3 convolution stack (MazeEncoder, src/models/encoders.py:15-25) on zero-bordered NHWC activations
 * act [B, (H+2)*(W+2), C] bf16 (border positions are zeros = the conv's zero padding): the input pixel of tap (ky, kx) for the
 * output position m is the ROW m + (ky-1)*(W+2) + (kx-1), so each k-block of the tcgen05 GEMM is a TMA box at a shifted row --
 * no im2col matrix.  idb200_conv_first_nhwc: first layer (C_in <= 2) + SiLU on CUDA cores from NCHW fp32 input;
 * idb200_conv3x3_gemm: one further layer, out = epi(conv(act) + bias) (epilogue 0: bf16, 1: SiLU + bf16), C_in == 32 or a
 * multiple of 64 (<= 256), C_out % 64 == 0, Wm bf16 [C_out, nkb*64] in the k-block order documented in csrc/gemm.cu;
 * idb200_pool_bordered: pooled[B, C] = mean over the H*W interior positions. */
int idb200_conv_first_nhwc(const float* x, int64_t B, int Cin, int H, int W, const float* w, const float* bias, int C1, void* out,
                           idb200_stream_t stream);
int idb200_conv3x3_gemm(const void* act, int C, const void* Wm, const float* bias, void* out, int N, int64_t B, int H, int W,
                        int epilogue, idb200_stream_t stream);
int idb200_pool_bordered(const void* act, int64_t B, int H, int W, int C, float* pooled, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Denoiser building blocks (KeypointDenoiser / InterpLevelDenoiser forwards,
 * src/models/denoiser_keypoints.py:82-113, src/models/denoiser_interp_levels.py:64-84,
 * src/models/transformer.py:35-46, src/models/encoders.py:8-71).
 * ---------------------------------------------------------------------------------------------- */

/* fp32 SIMT GEMM out[M,N](ldo) = act(A[M,K](lda) * W[N,K]^T + bias) (+ out when accumulate != 0):
 * the per-trajectory linears (cond_proj, FiLM, t_embed, level_proj, maze.fc, sg.mlp; M = B, K <= 256) and
 * the fp32 check mode of the token GEMMs.  A fp32 or bf16 (a_is_bf16); act: 0 none, 1 SiLU. */
int idb200_sgemm(const void* A, int a_is_bf16, int64_t lda, const float* W, const float* bias, float* out,
                 int64_t ldo, int64_t M, int N, int K, int act, int accumulate, idb200_stream_t stream);

/* MazeEncoder conv stack, src/models/encoders.py:15-24: [conv3x3 pad 1 + SiLU] x n_layers then the mean
 * over H x W.  occ / sdf [B,1,H,W] fp32 (sdf iff channels[0] == 2); channels host [n_layers+1];
 * weights / biases host arrays of device pointers ([C_out,C_in,3,3], [C_out]); pooled [B, C_last];
 * scratch: 2 * B * max(channels[1..n_layers-1]) * H * W floats for the intermediate planes (NULL if 1 layer). */
int idb200_conv_encoder(const float* occ, const float* sdf, int64_t B, int H, int W, int n_layers, const int* channels,
                        const float* const* weights, const float* const* biases, float* scratch, float* pooled,
                        idb200_stream_t stream);

/* sinusoid table out[rows, dim] = [sin(a f_i) | cos(a f_i)], f_i = exp(-ln(1e4) i / (dim/2)):
 * mode 0: a = r / max(1, rows-1) (continuous_time_embedding of idx/(T-1), denoiser_keypoints.py:24-34);
 * mode 1: a = args[r] (timestep_embedding :11-21; _positional_embedding of denoiser_interp_levels.py:54-62). */
int idb200_sinusoid(const float* args, int rows, int dim, int mode, float* out, idb200_stream_t stream);

/* token assembly + in_proj (denoiser_keypoints.py:102-111, denoiser_interp_levels.py:71-82):
 *   h[m,:] = [src0 | src1 | src2][m,:] . Wf + tab[tab_idx ? tab_idx[m] : m % L] + row_a[(m/L) * row_a_stride]
 *            + row_b[(m/L) * d]
 * src0 fp32 [M,n0], src1 fp32 [M,n1] or NULL, src2 uint8 [M,n2] or NULL (n0+n1+n2 <= 16), Wf [n0+n1+n2, d]. */
int idb200_embed_tokens(const float* src0, int n0, const float* src1, int n1, const uint8_t* src2, int n2,
                        const float* Wf, const float* tab, const int64_t* tab_idx, const float* row_a,
                        int64_t row_a_stride, const float* row_b, float* h, int64_t M, int L, int d,
                        idb200_stream_t stream);

/* LayerNorm(eps 1e-5) + FiLM, src/models/transformer.py:28-41: out = LN(h) * (1 + gamma[b]) + beta[b];
 * gamma_beta row b = [gamma (d) | beta (d)] at stride gb_stride (NULL: no FiLM); out bf16 or fp32 [M,d]. */
int idb200_ln_film(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                   void* out, int out_is_bf16, int64_t M, int L, int d, idb200_stream_t stream);
/* Training-forward form: the same, and the residual rows it read are also written to h_copy [M,d] fp32 (16-byte aligned, != h) --
 * what loss.backward() keeps of the LayerNorm input (train_interp_levels.py:1142-1161) without a second read of h. */
int idb200_ln_film_save(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                        void* out, int out_is_bf16, float* h_copy, int64_t M, int L, int d, idb200_stream_t stream);

/* output head y[M,D] = h[M,d] . W[D,d]^T + bias (D <= 4; the `out` Linear of both denoisers). */
int idb200_out_head(const float* h, const float* W, const float* bias, float* y, int64_t M, int d, int D,
                    idb200_stream_t stream);

/* multi-head self-attention of nn.MultiheadAttention (transformer.py:11,39) on a packed qkv [B*L, 3*H*32]:
 * out[B*L, H*32] = softmax(q k^T / sqrt(32) [+ causal mask]) v per trajectory and head, L <= 256, bf16 or fp32 in/out.
 * force_simt = 0: bf16 runs on tcgen05 (csrc/attention_tc5.cu: S = Q K^T and O = P V as tcgen05.mma with the score tile and
 *                the output accumulator in tensor memory, any L, H even; H odd falls to the legacy path); fp32 on CUDA cores.
 * force_simt = 1: fp32-arithmetic CUDA-core kernel (check mode).   force_simt = 2: legacy mma.sync kernel (L % 16 == 0 or
 *                L in {2,4,8}; otherwise the CUDA-core kernel) -- kept as a cross-check. */
int idb200_attention(const void* qkv, void* out, int is_bf16, int64_t B, int L, int H, int causal, int force_simt,
                     idb200_stream_t stream);

/* K3d  fused transformer MLP (d_model = 256): h[M,256] += W2 . SiLU(W1 . a + b1) + b2, the `ff` branch of
 * src/models/transformer.py:43-45 in one tcgen05 kernel; the [M, d_ff] hidden activation stays in TMEM / shared
 * memory.  a bf16 [M,256] (LN+FiLM output), W1 bf16 [ff,256], W2 bf16 [256,ff], ff % 128 == 0, ff <= 2048. */
int idb200_mlp_fused(const void* a, const void* W1, const float* b1, const void* W2, const float* b2, float* h,
                     int64_t M, int d, int ff, idb200_stream_t stream);

/* K3e  fused attention half of a transformer block (d_model = 256, 8 heads), src/models/transformer.py:35-41:
 *        h[M,256] += out_proj(MHA(LayerNorm(h) * (1 + gamma) + beta))
 * in one tcgen05 kernel per 128-token tile (LayerNorm+FiLM prologue, QKV projection, in-tile attention over the
 * 128 / L trajectories of the tile, out projection, TMA reduce-add residual).  L | 128, M % L == 0.
 *   gamma_beta: FiLM rows [gamma (256) | beta (256)] of trajectory m / L at stride gb_stride floats, or NULL.
 *   wqkv_packed bf16 [768,256] / bqkv_packed fp32 [768]: in_proj rows re-ordered head-group-major, group g (heads
 *   2g, 2g+1) = rows [Wq[64g:64g+64]; Wk[64g:64g+64]; Wv[64g:64g+64]].  wo bf16 [256,256], bo fp32 [256].
 *   causal != 0: the -inf upper-triangular mask of transformer.py:68-71. */
int idb200_attn_block(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                      const void* wqkv_packed, const float* bqkv_packed, const void* wo, const float* bo, int64_t M,
                      int L, int d, int H, int causal, idb200_stream_t stream);

/* K3d' fused MLP half, src/models/transformer.py:42-45:  h[M,256] += ff.2(SiLU(ff.0(LayerNorm(h) * (1 + gamma) + beta)))
 * = idb200_mlp_fused with the LayerNorm + FiLM prologue computed in shared memory (no [M,256] operand in HBM).
 * L | 8 or 8 | L (trajectory length, selects the FiLM row m / L). */
int idb200_mlp_block(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                     const void* W1, const float* b1, const void* W2, const float* b2, int64_t M, int L, int d, int ff,
                     idb200_stream_t stream);

/* K3g  in_proj + multi-head self attention of one block in ONE kernel (nn.MultiheadAttention inside src/models/transformer.py:11,39
 * without the out_proj): o[M, d] = MHA(a), a = the LayerNorm + FiLM output (bf16).  The packed projections q|k|v never leave the SM
 * (accumulator -> bf16 staging tile -> in-tile attention core); for the models the whole-encoder kernel does not take
 * (d_model = 384 / 12 heads, the trainer defaults of src/train/train_interp_levels.py:57-62; also d_model = 256).
 *   wqkv_packed bf16 [3d, d], bqkv_packed fp32 [3d]: head-group-major (group g of two heads = rows [Wq[64g..]; Wk[64g..]; Wv[64g..]],
 *   as for idb200_attn_block); L | 128, M % L == 0, head_dim 32; o may alias a (a tile's rows are read before they are written). */
int idb200_qkv_attention(const void* a, const void* wqkv_packed, const float* bqkv_packed, void* o, int64_t M, int L, int d, int H,
                         int causal, idb200_stream_t stream);
/* ... with the LayerNorm + FiLM prologue of transformer.py:35-41 computed in shared memory from the fp32 residual stream h (read only):
 * o[M, d] = MHA(LayerNorm(h) * (1 + gamma) + beta); gamma_beta: raw rows [gamma | beta] per trajectory (row m / L), or NULL. */
int idb200_ln_qkv_attention(const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride,
                            const void* wqkv_packed, const float* bqkv_packed, void* o, int64_t M, int L, int d, int H, int causal,
                            idb200_stream_t stream);

/* K3h  the MLP of one block in ONE pair-mode kernel for d_model = 384 (and 256), src/models/transformer.py:43-45:
 *   h[M, d] += W2 . SiLU(W1 . a + b1) + b2,  a = the LayerNorm + FiLM output (bf16); the hidden activation [M, d_ff] never leaves the SM.
 *   W1 bf16 [ff, d] (nn.Linear layout), W2_packed bf16 [d, ff] = ff.2.weight with its ROWS permuted by idb200_mlp_pair_w2_order
 *   (order[i] = source row of packed row i: each CTA of a pair holds the rows its half of the cta_group::2 MMAs produces);
 *   d_ff % 128 == 0, 128 <= d_ff <= 2048; h fp32, 16-byte aligned (updated by TMA reduce-add). */
int idb200_mlp_pair_w2_order(int d, int* order);
int idb200_mlp_pair(const void* a, const void* W1, const float* b1, const void* W2_packed, const float* b2, float* h, int64_t M, int d, int ff,
                    idb200_stream_t stream);
/* ... with the LayerNorm + FiLM prologue of transformer.py:42-43 computed in shared memory from h itself:
 * h += W2 . SiLU(W1 . (LayerNorm(h) * (1 + gamma) + beta) + b1) + b2; gamma_beta: raw rows [gamma | beta] per trajectory (row m / L) or NULL. */
int idb200_ln_mlp_pair(float* h, const float* ln_w, const float* ln_b, const float* gamma_beta, int64_t gb_stride, int L, const void* W1,
                       const float* b1, const void* W2_packed, const float* b2, int64_t M, int d, int ff, idb200_stream_t stream);

/* K3f  the whole TransformerEncoder (every layer of src/models/transformer.py:73-82) in ONE persistent tcgen05 kernel,
 * d_model = 256, 8 heads, d_ff % 128 == 0 (<= 1024), L | 128, M % L == 0.  A CTA (pair) carries a 128-token tile through
 * all layers with the fp32 residual stream resident in tensor memory; h is read and written once.
 *   layer_params fp32, per layer: [ln1_w 256 | ln1_b 256 | pend1 256 | bqkv_packed 768 | ln2_w 256 | ln2_b 256 | pend2 256 |
 *     0.5 * b1 (ff)].  pend1 / pend2 = the bias of the GEMM that accumulated into h just before that LayerNorm (ff.2 bias of
 *     the previous layer, zero for layer 0 / out_proj bias of this layer): the accumulating GEMMs never add their own bias,
 *     the LayerNorm that follows does.  bias_last [256] = ff.2 bias of the last layer (added when h is written back).
 *     Of bqkv_packed (per head group [bq 64 | bk 64 | bv 64]) the kernel adds ONLY the q part: a k bias shifts every score of a
 *     query equally (softmax-invariant: dropped) and a v bias passes through the attention unchanged (rows of softmax sum to 1),
 *     so the caller folds it into pend2 = out_proj.bias + out_proj.weight . bv (transformer.py:39 semantics are unchanged).
 *   film: FiLM table or NULL: the row (512 floats) of trajectory b = m / L for LayerNorm slot j (2l = film1 of layer l,
 *     2l+1 = film2) is at film + b * film_stride + j * film_ln_stride.  [B][2 * n_layers][512] is (8192-ish, 512);
 *     the LayerNorm-major form [2 * n_layers][B][512] (film_stride = 512, film_ln_stride = B * 512) makes the rows of a
 *     tile contiguous, which the kernel stages with ONE bulk copy per LayerNorm (16 separate copies cost 3 k cycles).  film_folded == 0: rows are [gamma | beta] (a = LN(h) * (1 + gamma) + beta); != 0: rows are
 *     [scale | shift] with the LayerNorm affine folded in, scale = ln_w * (1 + gamma), shift = ln_b * (1 + gamma) + beta
 *     (a = n * scale + shift, n = the normalised row) -- both are linear in cond_vec, so the host folds them into the FiLM GEMM.
 *     film_folded == 2: the same [scale | shift] rows stored as IEEE half (`film` points at fp16 data, both strides count
 *     16-bit elements and must be multiples of 8; needs L >= 8): halves the shared-memory bytes the LayerNorm reads per column.
 *   wqkv_packed bf16 [n_layers*768, 256] (per layer head-group-major, see idb200_attn_block), wo bf16 [n_layers*256, 256],
 *   w1 bf16 [n_layers*ff, 256], w2 bf16 [n_layers*256, ff]. */
int idb200_encoder_fused(float* h, const float* layer_params, const float* bias_last, const float* film, int64_t film_stride,
                         int64_t film_ln_stride, int film_folded, const void* wqkv_packed, const void* wo, const void* w1, const void* w2,
                         int64_t M, int L, int d, int H, int ff, int n_layers, int causal, idb200_stream_t stream);

/* The whole denoiser forward around the encoder in the same launch (denoiser_keypoints.py:102-113,
 * denoiser_interp_levels.py:71-84): token assembly + in_proj as the tile prologue, the `out` Linear as the tile epilogue.
 *   embed != NULL: h is not read; the tile's rows are built as in idb200_embed_tokens (same fp32 operation order).
 *   head  != NULL: h is not written; y[M, D] = (h + bias_last) . W^T + bias, D <= 4.
 *   h may be NULL when both are given.  Everything else as idb200_encoder_fused. */
typedef struct {
    const float* src0; int n0;                    /* [M, n0] fp32 token features */
    const float* src1; int n1;                    /* [M, n1] fp32 or NULL */
    const uint8_t* src2; int n2;                  /* [M, n2] uint8 (0/1) or NULL;  n0 + n1 + n2 <= 16 */
    const float* Wf;                              /* [n0+n1+n2, 256] fp32 */
    const float* tab;                             /* [rows, 256] fp32 */
    const int64_t* tab_idx;                       /* [M] row of tab per token, or NULL: the token's position in its trajectory */
    const float* row_a; int64_t row_a_stride;     /* [B or 1, 256] fp32 (stride 0 broadcasts one row) */
    const float* row_b;                           /* [B, 256] fp32 */
    int tab_rows;                                 /* rows of tab; 1..64 (and n0 + n1 + n2 <= 8) selects the staged-table prologue
                                                     (the table is copied into shared memory by TMA once per tile and gathered
                                                     there); 0 = unknown: rows are gathered from global memory */
} idb200_embed_t;
typedef struct {
    const float* W;                               /* [D, 256] fp32 */
    const float* bias;                            /* [D] */
    float* y;                                     /* [M, D] */
    int D;
} idb200_head_t;
int idb200_denoiser_fused(const idb200_embed_t* embed, const idb200_head_t* head, float* h, const float* layer_params,
                          const float* bias_last, const float* film, int64_t film_stride, int64_t film_ln_stride, int film_folded,
                          const void* wqkv_packed, const void* wo, const void* w1, const void* w2, int64_t M, int L, int d,
                          int H, int ff, int n_layers, int causal, idb200_stream_t stream);

/* K4 (tensor-core path)  two-layer MazeEncoder conv stack of src/models/encoders.py:15-24 in one launch:
 * conv3x3(cin->c1)+SiLU on CUDA cores into a shared-memory bf16 channels-last tile, conv3x3(c1->c2)+SiLU as an
 * implicit GEMM (mma.sync bf16, fp32 accumulate), spatial mean -> pooled [B, c2].
 * w0 [c1,cin,3,3] fp32; w1 packed bf16 [c2, 9*c1] with k = (ky*3+kx)*c1 + c.  c1 in {16,32,48,64}, c2 in {32,64}. */
int idb200_conv_encoder_tc(const float* occ, const float* sdf, int64_t B, int H, int W, int cin, int c1, int c2,
                           const float* w0, const float* b0, const void* w1_packed_bf16, const float* b1, float* pooled,
                           idb200_stream_t stream);

/* K4' (tcgen05 path)  the same two-layer stack as idb200_conv_encoder_tc with the second conv on the 5th-generation
 * tensor cores (TMEM accumulators, SWIZZLE_64B shifted activation copies instead of im2col; csrc/conv_tc5.cu).
 * Specialised for maze_channels = (32, 64) and H * roundup8(W + 2) <= 512 (the 21 x 21 particle maze). */
int idb200_conv_encoder_tc5(const float* occ, const float* sdf, int64_t B, int H, int W, int cin, int c1, int c2,
                            const float* w0, const float* b0, const void* w1_packed_bf16, const float* b1,
                            float* pooled, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Tail of the Stage-2 training step, src/train/train_interp_levels.py:1142-1173 (the network's backward pass between the
 * loss gradient and the optimiser is built from the "Backward" entry points below).
 * ------------------------------------------------------------------------------------------------ */
/* loss = sum_bt(w * ||delta_hat - target||^2) / (sum_bt(w) * D + 1e-8) / grad_accum, w = w_missing + (w_anchor - w_missing) *
 * conf[b,t] (anchor_conf branch, conf fp32 [B,T]) or w_anchor / w_missing by mask[b,t] (uint8) -- exactly one of conf / mask.
 * loss_scal[0] = loss, loss_scal[1] = the common factor of the gradient; grad_out (or NULL) = d loss / d delta_hat [B,T,D].
 * scratch: idb200_tail_scratch_doubles() doubles (per-block partial sums, reduced in a fixed order: deterministic). */
int64_t idb200_tail_scratch_doubles(void);   /* capacity both reduction entry points below need (their grids are clamped to it) */
int idb200_stage2_loss(const float* delta_hat, const float* target, const float* conf, const uint8_t* mask, float w_anchor,
                       float w_missing, float grad_accum, int64_t B, int T, int D, double* scratch, float* loss_scal,
                       float* grad_out, idb200_stream_t stream);

/* torch.nn.utils.clip_grad_norm_ over ONE flat gradient buffer: norm_coef[0] = ||grad||_2, norm_coef[1] = min(1, max_norm /
 * (norm + 1e-6)).  scratch: idb200_tail_scratch_doubles() doubles.  The coefficient is consumed on the device by idb200_adamw_ema_step. */
int idb200_grad_clip_coef(const float* grad, int64_t n, float max_norm, double* scratch, float* norm_coef, idb200_stream_t stream);

/* torch.optim.AdamW step (decoupled weight decay, bias-corrected, torch's operation order) on grad * norm_coef[1] (norm_coef
 * NULL: unclipped), followed by EMA.update (src/utils/ema.py:11-17; ema NULL: skipped).  step is 1-based.  36 B per parameter. */
int idb200_adamw_ema_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n, float lr,
                          float beta1, float beta2, float eps, float weight_decay, int64_t step, float ema_decay,
                          const float* norm_coef, idb200_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Backward of the denoisers (what loss.backward() does in src/train/train_interp_levels.py:1159 and
 * src/train/train_keypoints.py:540 for the modules of src/models/{transformer,encoders,denoiser_*}.py).
 * Dense contractions reuse idb200_gemm_bf16: dX = dY * W is the same GEMM with W^T as the weight operand; dW = dY^T * X is
 * idb200_gemm_bf16_nn_splitk (reduction over the tokens, operands read in place) followed by idb200_reduce_rows.
 * ---------------------------------------------------------------------------------------------- */
/* partial[s] [M,N] fp32 = A[:, K_s] * W[:, K_s]^T for the s-th of `splits` equal slices of K ((K / 64) % splits == 0). */
int idb200_gemm_bf16_splitk(const void* A, const void* W, float* partial, int64_t M, int N, int K, int splits,
                            idb200_stream_t stream);
/* partial[s] [M,N] fp32 = A[K_s, M]^T * W[K_s, N]: both operands row-major with the REDUCTION index as rows (the weight
 * gradient dW = dY^T X straight from the token-major activations, no transposes: MN-major UMMA operands).  M % 8 == 0,
 * N % 64 == 0, K % 64 == 0, (K / 64) % splits == 0. */
int idb200_gemm_bf16_nn_splitk(const void* A, const void* W, float* partial, int64_t M, int N, int64_t K, int splits,
                               idb200_stream_t stream);
/* dst bf16 [N,M] = src[M,N]^T (src fp32 or bf16). */
int idb200_transpose_bf16(const void* src, int src_is_f32, int64_t M, int N, void* dst, idb200_stream_t stream);
/* out[N] (+)= scale * column sums of src[M,N] (src_kind 0 fp32, 1 bf16); scratch: idb200_colsum_scratch_floats(M, N) floats. */
int idb200_colsum_scratch_floats(int64_t M, int N);
int idb200_colsum(const void* src, int src_kind, int64_t M, int N, float* scratch, float scale, int accumulate, float* out,
                  idb200_stream_t stream);
/* The same for `segs` stacked matrices: src [segs, M, N] -> out [segs, N] in two launches (the per-LayerNorm [B, 3d] partials of a
 * whole backward pass are reduced at once); scratch: segs * idb200_colsum_scratch_floats(M, N) floats. */
int idb200_colsum_segments(const void* src, int src_kind, int segs, int64_t M, int N, float* scratch, float scale, int accumulate,
                           float* out, idb200_stream_t stream);
/* out[W] (+)= scale * sum_r partial[r, W] (fixed order). */
/* Batched small operations of a training step (pointer tables are HOST arrays, passed on as kernel parameters: graph capturable):
 * n contiguous fp32 copies dsts[i][0..counts[i]) = srcs[i][...] in one launch per 96 segments (gradient slices into the flat arena); */
int idb200_multi_copy_f32(const float* const* srcs, float* const* dsts, const int64_t* counts, int n, idb200_stream_t stream);
/* and, for n fp32 matrices [rows[i], cols[i]]: dsts[i] = bf16 copy, dsts_t[i] = bf16 transpose [cols, rows] (either may be NULL) -- the
 * per-step operand copies of the token GEMMs' weights (autocast's bf16 casts, train_interp_levels.py:1142-1161). */
int idb200_cast_weights_bf16(const float* const* srcs, void* const* dsts, void* const* dsts_t, const int* rows, const int* cols, int n,
                             idb200_stream_t stream);
int idb200_reduce_rows(const float* partial, int R, int64_t W, float scale, int accumulate, float* out, idb200_stream_t stream);
/* bf16 elementwise, n even: mode 0 y = silu(u); mode 1 y = g * silu'(u). */
int idb200_silu_bf16(const void* u, const void* g, int64_t n, int mode, void* y, idb200_stream_t stream);
int idb200_silu_f32(const float* u, const float* g, int64_t n, int mode, float* y, idb200_stream_t stream);   /* fp32 variant */
/* Backward of a = LayerNorm(h) * (1 + gamma) + beta (transformer.py:35-46): dh += ..., dh_bf16 (or NULL) = bf16 copy of the
 * updated dh, dgb[b] = [dgamma | dbeta] (NULL iff gamma_beta is NULL), dwb_part [B, 2d] = per-trajectory [dw | db] partials. */
int idb200_ln_film_bwd(const float* da, const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta,
                       int64_t gb_stride, int64_t B, int L, int d, float* dh, void* dh_bf16, float* dgb, int64_t dgb_stride,
                       float* dwb_part, float* stats_scratch, idb200_stream_t stream);
/* Two-pass form with (i) da in fp32 or bf16 (da_is_bf16) and (ii) with_dh_sum != 0: dwb_part is [B, 3d] = [dw | db | sum_t of the
 * UPDATED dh]: the last third, summed over the trajectories, is the bias gradient of the GEMM that accumulated into this point of the
 * residual stream (transformer.py:39-45: out_proj / the previous layer's ff.2), which otherwise costs a full extra read of dh.
 * stats_scratch fp32 [B*L, 4], 16-byte aligned (required). */
int idb200_ln_film_bwd2(const void* da, int da_is_bf16, const float* h, const float* ln_w, const float* ln_b, const float* gamma_beta,
                        int64_t gb_stride, int64_t B, int L, int d, float* dh, void* dh_bf16, float* dgb, int64_t dgb_stride,
                        float* dwb_part, int with_dh_sum, float* stats_scratch, idb200_stream_t stream);
/* stats_scratch: fp32 [B*L, 4] (16-byte aligned) selects the two-pass form (row scalars by a warp per token, then a thread per
 * column over the trajectory's tokens: coalesced, low register count); NULL runs the one-block-per-trajectory kernel. */
/* Backward of the packed-QKV multi-head attention (head_dim 32, L <= 64): qkv, dqkv bf16 [B*L, 3d]; dO bf16 [B*L, d].
 * 32 < L <= 64 runs on mma.sync tensor cores (P and dS rounded to bf16 for the second products, as the forward rounds P);
 * force_simt != 0 selects the fp32 shared-memory kernel that serves L <= 32. */
int idb200_attention_bwd(const void* qkv, const void* dO, void* dqkv, int64_t B, int L, int H, int causal, int force_simt,
                         idb200_stream_t stream);
/* The same (tensor-core kernel, 32 < L <= 64) + traj_colsum [B, 3d] fp32: per-trajectory sums over the tokens of the bf16 dqkv rows.
 * Their sum over B is the in_proj bias gradient (nn.MultiheadAttention inside transformer.py:39) without re-reading dqkv. */
int idb200_attention_bwd_sums(const void* qkv, const void* dO, void* dqkv, float* traj_colsum, int64_t B, int L, int H, int causal,
                              idb200_stream_t stream);
/* dh[M,d] = dy[M,D] * W[D,d] (out head backward, D <= 8), plus an optional bf16 copy. */
int idb200_head_bwd(const float* dy, const float* W, int64_t M, int d, int D, float* dh, void* dh_bf16, idb200_stream_t stream);
/* out[n,K] (+)= A[M,n]^T * X[M,K], n <= 8 (out-head / in_proj weight gradients); scratch: ..._scratch_floats(M, n, K). */
int idb200_narrow_outer_scratch_floats(int64_t M, int n, int K);
int idb200_narrow_outer(const float* A, int n, const float* X, int64_t M, int K, float* scratch, int accumulate, float* out,
                        idb200_stream_t stream);
/* out[B,d] = sum over the L tokens of src[B,L,d]. */
int idb200_token_sum(const float* src, int64_t B, int L, int d, float* out, idb200_stream_t stream);
/* out[i,j] (+)= sum_k A[i*sa0 + k*sa1] * B[j*sb0 + k*sb1]: strided fp32 GEMM for the per-trajectory linears' backward. */
int idb200_sgemm_strided(const float* A, int64_t sa0, int64_t sa1, const float* Bm, int64_t sb0, int64_t sb1, float* out,
                         int64_t ldo, int M, int N, int K, int accumulate, idb200_stream_t stream);
/* Conv encoder in training form (encoders.py:8-25), activations NHWC bf16 [B, H*W, C] holding pre-activations:
 * col bf16 [B*H*W, Kpad], col[., (ky*3+kx)*C + c] = act(src[., y+ky-1, x+kx-1, c]) (act 1: SiLU), zero padding. */
int idb200_im2col3x3(const void* src, int64_t B, int H, int W, int C, int Kpad, int act, void* col, idb200_stream_t stream);
/* pooled[B,C] = mean_p silu(u[B,P,C]);  du = dpooled / P * silu'(u). */
int idb200_pool_silu(const void* u, int64_t B, int P, int C, float* pooled, idb200_stream_t stream);
int idb200_pool_silu_bwd(const void* u, const float* dpooled, int64_t B, int P, int C, void* du, idb200_stream_t stream);

/* KeypointSelector pieces (src/models/keypoint_selector.py:22-188; the producer of anchor logits for kp_index_mode = selector).
 * Cross attention of nn.MultiheadAttention(q, memory, memory), head_dim 32: q bf16 [B,Lq,d] (projected queries, Lq % 16 == 0),
 * kv_a bf16 [B,La,2d] and kv_b bf16 [B,Lb,2d] (or NULL, Lb = 0): [K | V] projections of the memory tokens in two buffers
 * (spatial tokens, extra tokens; key order does not matter), out bf16 [B,Lq,d]. */
int idb200_cross_attention(const void* q, const void* kv_a, const void* kv_b, void* out, int64_t B, int Lq, int La, int Lb, int H,
                           idb200_stream_t stream);
/* Start / goal maps (:113-146): out fp32 [B,2,H,W] = exp(-((x - cx)^2 + (y - cy)^2) / (2 sigma^2)) for sigma > 0, one-hot at the
 * rounded (half-to-even) cell for sigma <= 0 (:129-139). */
int idb200_sg_map(const float* start_goal, int64_t B, int H, int W, float sigma, float* out, idb200_stream_t stream);

/* DP anchor placement, src/selection/epiplexity_dp.py:171-228 (dp_select_indices_batch): the producer of idx for
 * kp_index_mode = dp.  C fp32 [B,T,T] segment costs (inf = no segment); idx int64 [B,K], 2 <= K <= T, idx[:,0] = 0,
 * idx[:,K-1] = T-1, minimising sum C[idx[k-1], idx[k]] with torch.argmin's first-minimum tie rule (bit-identical indices).
 * status int32 [B]: 0 ok, 1 no finite path to T-1, 2 backtrack failed (the reference raises RuntimeError for both). */
int idb200_dp_select(const float* C, int64_t B, int T, int K, int64_t* idx, int* status, idb200_stream_t stream);

/* Segment costs for the DP placement, src/selection/epiplexity_dp.py:120-147 (compute_segment_costs_batch): x_pos fp32 [B,T,D]
 * (D >= 2), segments s = (seg_i[s], seg_j[s]) with n sample points t_idx [S,n] (int64) at chord fractions alpha [S,n] and
 * weight [S]; out fp32 [B,S] = weight * weight_scale * sum_n ||x[t] - (x_i + alpha (x_j - x_i))||^2 over the 2 position dims. */
int idb200_segment_costs(const float* x_pos, int64_t B, int T, int D, const int64_t* seg_i, const int64_t* seg_j, const int64_t* t_idx,
                         const float* alpha, const float* weight, int S, int n, float weight_scale, float* out, idb200_stream_t stream);

/* Batched trajectory metrics, src/eval/metrics.py:68-128 (compute_metrics_batch; _pos_to_cell :13-24): the step right after
 * the generation path (the reference loops over samples on the host, sample_generate.py:1323-1398).
 *   occ fp32 [B,H,W] (occ_stride = H*W, or 0 to broadcast one map); traj fp32 [B,T,D], dims 0:2 = (x, y) in [0,1];
 *   goal fp32 [B,D] (goal_stride = D or 0); gt fp32 [B,T,D] or NULL (gt_stride = T*D or 0); success_thr = 1 / W as fp32.
 *   out fp32 [B, n_out]: collision_rate, goal_dist, success, path_length, smoothness (0 if T < 3) [, mse_to_gt if n_out == 6]. */
int idb200_traj_metrics(const float* occ, int64_t occ_stride, const float* traj, const float* goal, int64_t goal_stride,
                        const float* gt, int64_t gt_stride, int64_t B, int T, int D, int H, int W, float success_thr,
                        float* out, int n_out, idb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* IDB200_H */
