/* idee_b200 -- C ABI of the B200-native IDEE hot path (libidee_b200.so, sm_100a only).
 *
 * The reference (HakamShams/IDEE) ships no native code and no FFI: its hot path is plain PyTorch modules
 * (SURVEY.md section 2).  These entry points are therefore what a binding for that path binds; each one names the
 * reference code it replaces (paths relative to the reference repo root).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; idee_last_error() returns the message (thread-local);
 *   - all pointers are DEVICE pointers into caller-owned buffers (including workspaces, sized by the *_workspace_bytes
 *     queries); the library never allocates, frees or synchronises; every launch goes to `stream` (a cudaStream_t);
 *   - activations are fp32, channel-last: element (n,v,t,h,w,c) of a token tensor lives at ((((n*V+v)*T+t)*H+h)*W+w)*C+c
 *     unless explicit strides (in elements) are passed;
 *   - parameters keep the reference (PyTorch state_dict) layouts; gradients are written in the same layouts;
 *   - there is no CPU fallback and no silent dispatch: an unsupported configuration is an error.
 */
#ifndef IDEE_B200_H
#define IDEE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IDEE_B200_VERSION 205

const char* idee_last_error(void);
int idee_version(void);
int idee_check_device(void);

/* ---- PatchEmbed3D, patch (1,1,1): Conv3d(Cin->E,k=1,bias) + LayerNorm(E, no affine)   Swin_3D.py:449-491, :434 ----
 * x: [N,V,Cin,T,H,W] with strides x_strides[6] = (n,v,c,t,h,w); w: [V][E][Cin]; b: [V][E]; y: [N,V,T,H,W,E]. */
int idee_embed_ln_fwd(const float* x, const int64_t* x_strides, const float* w, const float* b, float* y,
                      int N, int V, int Cin, int T, int H, int W, int E, void* stream);
size_t idee_embed_ln_bwd_workspace_bytes(int V);
int idee_embed_ln_bwd(const float* x, const int64_t* x_strides, const float* w, const float* b, const float* gy,
                      float* gw, float* gb, int N, int V, int Cin, int T, int H, int W, int E,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---- SwinTransformerBlock3D (LN1, cyclic shift, window partition, W-MSA/SW-MSA with relative position bias and shift
 *      mask, proj, residual, LN2, MLP with erf-GELU, residual)                    Swin_3D.py:24-42, 45-74, 93-287, 340-352 ----
 * One call runs the block of ALL V variables (V independent weight sets).
 * params: [V][param_stride] floats; per variable, in the reference state_dict order of one block:
 *   relative_position_bias_table[rpb_rows][heads] | qkv.weight[3C][C] | qkv.bias[3C] | proj.weight[C][C] | proj.bias[C] |
 *   fc1.weight[hidden][C] | fc1.bias[hidden] | fc2.weight[C][hidden] | fc2.bias[C]      (idee_swin_block_packed_floats)
 * rel_index: int32 [G][G], G = wd*wh*ww, the [:G,:G] slice of the module's relative_position_index buffer (Swin_3D.py:158-160).
 * (wd,wh,ww) / (st,sh,sw) are the window and shift ALREADY clamped by get_window_size (Swin_3D.py:77-90). */
typedef struct {
    int N, V, T, H, W;      /* tokens [N,V,T,H,W,C] */
    int C, heads, hidden;   /* built: 16, 2, 64 */
    int wd, wh, ww;         /* window */
    int st, sh, sw;         /* cyclic shift of this block (0,0,0 = W-MSA) */
    int rpb_rows;           /* rows of the bias table, (2Wd-1)(2Wh-1)(2Ww-1) of the CONFIGURED window */
    float scale;            /* qk scale, head_dim^-0.5 unless overridden */
    int64_t param_stride;   /* floats between consecutive variables' packed parameter blocks */
    int precision;          /* 0: fp32 CUDA-core exact path; 1: bf16 tensor-core operands, fp32 accumulate / softmax / LayerNorm */
    /* Fused patch embedding (optional, precision 1, in_chans == 1): when embed_x != NULL the block's input tokens are
     * LayerNorm(embed_w[v][c] * embed_x[n,v,t,h,w] + embed_b[v][c]) (PatchEmbed3D, Swin_3D.py:473-491) evaluated on the fly from
     * the contiguous raw input embed_x [N,V,1,T,H,W]; the token tensor `x` of idee_swin_block_fwd/bwd is then not read (may be
     * NULL) and never has to exist in HBM.  gx of idee_swin_block_bwd is still the gradient w.r.t. those tokens
     * (feed it to idee_embed_ln_bwd). */
    const float* embed_x; const float* embed_w; const float* embed_b;
    /* backward, optional: when embed_gw / embed_gb ([V][16] each) are given, the call also runs the embedding's backward
     * (LayerNorm backward + weight / bias reduction) and writes the two gradients here.  act_dtype 1: gx receives the gradient
     * w.r.t. the embedded tokens and one streaming pass over it produces the two gradients; act_dtype 0: the attention half does
     * it on the token gradient in registers and the final content of gx is unspecified. */
    float* embed_gw; float* embed_gb;
    /* act_dtype 1 (precision 1 only) selects the tcgen05 / TMEM kernels (swin_umma.cuh): the saved activation ymid and the token
     * gradients gout / gx are then bf16 (__nv_bfloat16* passed through the float* parameters), out_bf16 must be NULL, gx must not
     * alias gout, and the block input x / output out are float (x_dtype / out_dtype 0) or bf16 (1).  act_dtype 0: every token
     * tensor is float and x_dtype / out_dtype must be 0. */
    int act_dtype, x_dtype, out_dtype;
} idee_swin_desc;

int idee_swin_block_packed_floats(int rpb_rows);
/* ymid (optional, may be NULL): the mid-block residual x + attn(...) saved for the backward pass.
 * out_bf16 (optional, may be NULL; precision 1 only): a bf16 copy of `out` in the same layout, written by the same kernel for a
 * consumer that rounds its input to bf16 anyway (the proj_var conv, see idee_conv_desc.x_dtype); `out` may then be NULL
 * (the fp32 tensor is not written at all) */
int idee_swin_block_fwd(const idee_swin_desc* d, const float* x, float* out, float* ymid, void* out_bf16, const float* params,
                        const int32_t* rel_index, void* stream);
size_t idee_swin_block_bwd_workspace_bytes(const idee_swin_desc* d);
/* gx may alias gout.  gparams: [V][param_stride], every packed element is overwritten. */
int idee_swin_block_bwd(const idee_swin_desc* d, const float* x, const float* ymid, const float* gout, float* gx,
                        const float* params, const int32_t* rel_index, float* gparams,
                        void* workspace, size_t workspace_bytes, void* stream);

/* ---- channel-last Conv3d of the path ----
 * proj=1: Conv3d(Cin,Cout,k=3,s=1,p=1,padding_mode='replicate')           Swin_3D.py:586-592   (encoder proj_var)
 * proj=0: Conv3d(Cin,Cout,k=(2,3,3),s=(2,1,1),p=(0,1,1)), zero padding     classifier/CNN_3D.py:36-38, 83-85
 * Images are indexed (n,v), n<N, v<V; image (n,v) uses weight set v when Vw==V, weight set 0 when Vw==1.
 * Channels come in 16-wide chunks; chunk k of the input sits at (k / in_cpg) * x_sg + (k % in_cpg) * 16 from the pixel
 * address (this lets the joint classifier head read the V per-variable planes of z_q as one 16*V-channel image);
 * same for the output with out_cpg / y_sg.  w: [Vw][Cout][Cin][kt][3][3] (reference layout), b: [Vw][Cout]. */
typedef struct {
    int N, V, Vw;
    int Cin, Cout;
    int Ti, Hi, Wi, To, Ho, Wo;
    int proj;               /* 1: 3x3x3 replicate, 0: (2,3,3)/(2,1,1) zero-pad */
    int relu;               /* forward: fuse ReLU into the epilogue */
    int precision;          /* 0: fp32 CUDA-core exact path; 1: bf16 tensor-core operands, fp32 accumulate;
                               2: as 1, and the 96->96 classifier conv (forward / data gradient) runs on tcgen05 + TMEM */
    int64_t x_sn, x_sv, x_st, x_sh, x_sw, x_sg; int in_cpg;
    int64_t y_sn, y_sv, y_st, y_sh, y_sw, y_sg; int out_cpg;
    /* element type of the activation tensors, 0: float (default), 1: bf16.  bf16 storage is accepted by the precision >= 1
     * 16 -> 16 proj conv only: those kernels round their operands to bf16 when they load them, so keeping the bf16 copy in
     * HBM gives bit-identical results with half the traffic and no conversion pass.  Strides stay in elements.
     *   x_dtype: x (forward / weight gradient) and relu_src (data gradient);  y_dtype: y and gy;  gx_dtype: gx */
    int x_dtype, y_dtype, gx_dtype;
    /* 1: run the bf16-input 16 -> 16 proj conv (forward and data gradient) on tcgen05 + TMEM (conv16_umma.cu) instead of
     * mma.sync; same operands and rounding points, different accumulation order */
    int umma16;
    /* 1: run the 96 -> 96 classifier conv (forward and data gradient, fp32 I/O) on the warp-specialised tcgen05 + TMEM kernel
     * (conv96_umma.cu) instead of mma.sync */
    int umma96;
    /* > 0: only the first cin_real input channels carry data (the rest are structurally zero, e.g. the padded plane image of the
     * rank-1 joint conv); the data gradient then computes and writes only those channels (rounded up to 8) */
    int cin_real;
} idee_conv_desc;

size_t idee_conv3d_fwd_workspace_bytes(const idee_conv_desc* d);
int idee_conv3d_fwd(const idee_conv_desc* d, const void* x, const float* w, const float* b, void* y,
                    void* workspace, size_t workspace_bytes, void* stream);
/* gx = conv^T(gy); when relu_src != NULL (same layout as gx) gx is multiplied by (relu_src > 0): the ReLU that
 * produced this conv's input is folded into the data gradient */
size_t idee_conv3d_dgrad_workspace_bytes(const idee_conv_desc* d);
int idee_conv3d_dgrad(const idee_conv_desc* d, const void* gy, const float* w, const void* relu_src, void* gx,
                      void* workspace, size_t workspace_bytes, void* stream);
size_t idee_conv3d_wgrad_workspace_bytes(const idee_conv_desc* d);
int idee_conv3d_wgrad(const idee_conv_desc* d, const void* x, const void* gy, float* gw, float* gb,
                      void* workspace, size_t workspace_bytes, void* stream);
/* Both gradients of one conv (what autograd asks of torch.nn.Conv3d.backward, Swin_3D.py:586-592 / classifier/CNN_3D.py:83-85):
 * same arguments as the two calls above.  The 16 -> 1 proj conv on bf16 storage (x_dtype 1, relu_src == x or NULL) runs ONE
 * kernel that reads x and gy once for both results; other geometries run idee_conv3d_wgrad then idee_conv3d_dgrad. */
size_t idee_conv3d_bwd_workspace_bytes(const idee_conv_desc* d);
int idee_conv3d_bwd(const idee_conv_desc* d, const void* x, const void* gy, const float* w, const void* relu_src, void* gx,
                    float* gw, float* gb, void* workspace, size_t workspace_bytes, void* stream);

/* ---- LFQ quantiser, dim=16, codebook_size=2                                      models/codebook/LFQ.py:183-307 ----
 * z,zq,gz,gzq: [ntok][16]; indices: int64 [ntok]; xq (optional, may be NULL): float [ntok], the quantised scalar x (+-1) with
 * z_q = x * w_out + b_out, gxq (optional): gradient w.r.t. x from consumers that use the rank-1 form of z_q; stats: float[8] = {aux, per_sample_entropy, codebook_entropy,
 * commit, mean p0, mean p1, ntok, 0} (written only when training); grads: float[49] = g_w_in[16] | g_b_in | g_w_out[16] | g_b_out[16]
 * Pre-projected form: dim == 1 (forward) / w_in == NULL (backward): z is the scalar s = project_in(z) itself ([ntok]), computed by
 * the producer (the encoder's last conv folded with project_in into one 16 -> 1 conv); gz is then [ntok] and g_w_in / g_b_in are 0. */
size_t idee_lfq_workspace_bytes(int64_t ntok);
int idee_lfq_fwd(const float* z, const float* w_in, const float* b_in, const float* w_out, const float* b_out,
                 float* zq, int64_t* indices, float* xq, float* stats, int64_t ntok, int dim, int codebook_size, int training,
                 float inv_temperature, float lambda_commit, float lambda_entropy, float diversity_gamma,
                 void* workspace, size_t workspace_bytes, void* zq_bf16 /* optional bf16 copy of zq, may be NULL */, void* stream);
int idee_lfq_bwd(const float* z, const float* gzq, const float* gxq, const float* g_aux, const float* stats, const float* w_in,
                 const float* b_in, const float* w_out, float* gz, float* grads, int64_t ntok, float inv_temperature,
                 float lambda_commit, float lambda_entropy, float diversity_gamma,
                 void* workspace, size_t workspace_bytes, void* stream);

/* General LFQ, codebook_size = 2^K for K = 2..4 (K = 1 is idee_lfq_fwd): project_in Linear(16,K), K-bit sign codes (argmin of the
 * 2^K code distances, ties -> lowest index), straight-through estimator, project_out Linear(K,16), entropy / commitment losses.
 * w_in [K][16], b_in [K], w_out [16][K], b_out [16] (reference layouts).  stats: float[4 + 2^K] = aux | H_tok | H_cb | commit | mean p.
 * grads (backward): g_w_in [K][16] | g_b_in [K] | g_w_out [16][K] | g_b_out [16].                       LFQ.py:92-101,134-146,183-307 */
size_t idee_lfqk_workspace_bytes(int codebook_bits);
int idee_lfqk_fwd(const float* z, const float* w_in, const float* b_in, const float* w_out, const float* b_out, float* zq,
                  int64_t* indices, float* stats, int64_t ntok, int dim, int codebook_bits, int training, float inv_temperature,
                  float lambda_commit, float lambda_entropy, float diversity_gamma, void* workspace, size_t workspace_bytes, void* stream);
int idee_lfqk_bwd(const float* z, const float* gzq, const float* g_aux, const float* stats, const float* w_in, const float* b_in,
                  const float* w_out, float* gz, float* grads, int64_t ntok, int codebook_bits, int training, float inv_temperature,
                  float lambda_commit, float lambda_entropy, float diversity_gamma, void* workspace, size_t workspace_bytes, void* stream);

/* ---- losses                                                                         models/losses.py:98-168 ----
 * BCE_loss_synthetic over K logit maps sharing one target: element (k,n,i) of pred at k*stride_k + n*stride_n + i, i<HW;
 * target: [N][HW]; wts: float[2] out (class weights); loss: float[K]; dpred (optional, pred's layout): d loss[k] / d pred. */
size_t idee_bce_loss_workspace_bytes(int K);
int idee_bce_loss_fwd(const float* pred, int64_t stride_k, int64_t stride_n, int K, int N, int64_t HW, const float* target,
                      float* wts, float* loss, float* dpred, void* workspace, size_t workspace_bytes, void* stream);
/* Anomaly_L1_loss_synthetic: zq [N,V,T,HW,16], mask [N][HW], vq0 [16]; out: float[2] = {loss, total weight} */
size_t idee_anomaly_l1_workspace_bytes(int64_t ntok);
int idee_anomaly_l1_fwd(const float* zq, const float* mask, const float* vq0, int N, int V, int T, int64_t HW, int C, float* out,
                        void* workspace, size_t workspace_bytes, void* stream);
int idee_anomaly_l1_bwd(const float* zq, const float* mask, const float* vq0, int N, int V, int T, int64_t HW, int C,
                        const float* out, const float* g_loss, float* gzq, void* stream);

/* The same loss on the rank-1 form of z_q (z_q[c] = x * w_out[c] + b_out[c], x = +-1 the quantised scalar of idee_lfq_fwd):
 * reads the scalar plane xq [N,V,T,HW] instead of the 16-channel z_q; out: float[4] = {loss, total weight, count(+1), count(-1)};
 * backward writes gxq [N,V,T,HW] (gradient w.r.t. x, fed to idee_lfq_bwd) and g_w_out[16], g_b_out[16] */
size_t idee_anomaly_rank1_workspace_bytes(int64_t ntok);
int idee_anomaly_rank1_fwd(const float* xq, const float* mask, const float* w_out, const float* b_out, const float* vq0, int N, int V,
                           int T, int64_t HW, int C, float* out, void* workspace, size_t workspace_bytes, void* stream);
int idee_anomaly_rank1_bwd(const float* xq, const float* mask, const float* w_out, const float* b_out, const float* vq0, int N, int V,
                           int T, int64_t HW, int C, const float* out, const float* g_loss, float* gxq, float* gw, float* gb,
                           void* stream);

/* The V scalar planes xq [N,V,THW] of the rank-1 form as one 16-channel channel-last image planes [N,THW,16]: channel v < V = xq_v,
 * channel V = 1 (carries the project_out bias through the zero-padded border), channels above = 0 -- the input of the joint
 * classifier's first conv (classifier/CNN_3D.py:77,118) once project_out (LFQ.py:284) is folded into its weights.
 * The backward gathers channels 0..V-1 of the image gradient into gxq [N,V,THW].  1 <= V <= 15. */
int idee_rank1_planes_fwd(const float* xq, float* planes, int N, int V, int64_t THW, void* stream);
int idee_rank1_planes_bwd(const float* gplanes, float* gxq, int N, int V, int64_t THW, void* stream);

/* ---- CNN_3D encoder: per-token tail of a residual conv block                     models/encoder/CNN_3D.py:74-147 ----
 * out = shortcut + ReLU(LayerNorm_16(y) * gamma[v] + beta[v]) on channel-last tokens [N,V,THW,16] (y = conv3^3 output of
 * idee_conv3d_fwd); gamma / beta [V][16].  out_bf16 (optional) receives a bf16 copy of out for the next conv (x_dtype 1).
 * Backward: gy = d/dy, dgamma / dbeta [V][16]; the gradient w.r.t. the shortcut is gout itself. */
int idee_ln_act_res_fwd(const float* y, const float* shortcut, const float* gamma, const float* beta, float* out, void* out_bf16,
                        int N, int V, int64_t THW, int C, void* stream);
size_t idee_ln_act_res_bwd_workspace_bytes(int V);
int idee_ln_act_res_bwd(const float* y, const float* gamma, const float* beta, const float* gout, float* gy, float* dgamma,
                        float* dbeta, int N, int V, int64_t THW, int C, void* workspace, size_t workspace_bytes, void* stream);

/* ---- optimiser: torch.optim.Adam(lr, betas, eps, weight_decay) on one flat buffer        train_synthetic.py:127-129 ---- */
int idee_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int step, void* stream);
/* the same update with the step counter and the learning rate in device memory (state[0] = steps taken so far as float, advanced
 * by the call; state[1] = lr): one captured CUDA-graph launch sequence serves every replay, schedulers write state[1] */
int idee_adam_step_state(float* p, const float* g, float* m, float* v, int64_t n, float* state, float beta1, float beta2, float eps,
                         float weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IDEE_B200_H */
