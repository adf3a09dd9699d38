// One encoder layer per ABI call in the bf16-STORAGE mode: the SST self-attention layer (pcdet/models/model_utils/
// sst_basic_block.py:58-84) or the WCA cross-attention layer (wca_block.py:70-103), forward and backward.
// Forward = 6 launches (7 / 9 for a cross layer's second projection), the [W | table^T] operand of the projection coming from the
// caller's per-step cache (tmae_bf16_weights.in_wcat) or, without it, from one more launch:
//   [packed q/k/v GEMM with the position term as a one-hot second operand + per-head L2 normalisation] -> window attention (warp
//   kernels for <= 16-token windows, tcgen05 tiles above) -> [out_proj GEMM + bias + residual + LayerNorm1]
//   -> [linear1 GEMM + bias + GELU, GELU' saved for the backward] -> [linear2 GEMM + bias + residual + LayerNorm2]
// Backward: LayerNorm-backward passes (with the bias gradients), data-gradient GEMMs on the caller's stream, weight-gradient GEMMs on the
// library's auxiliary stream (fork / join by events inside the call).
// The reference runs ~25 kernels per drop level per layer for the same arithmetic.  Activations (x, q/k/v, o, x1, h, y) and
// their gradients are bf16 in HBM; row statistics, softmax log-sum-exp, normalisation factors, the position table, master
// weights and every parameter gradient are fp32.  The attention core runs on the tcgen05 window kernel (attention_tc.cu);
// `attn_impl = 0` bridges to the fp32-I/O mma.sync kernels of attention.cu through cast passes (kept as the checker).
#include <cuda_bf16.h>

#include <map>
#include <mutex>

#include "common.cuh"

using namespace tmae;

extern "C" {
int tmae_bf16_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* preact, int64_t m, int64_t n, int64_t k, int32_t act,
                         int32_t accumulate, void* stream);
int tmae_bf16_qkv_fwd(const void* x, const void* w, const float* table, const uint8_t* posidx, void* y, float* inv, int64_t m, int64_t n,
                      int64_t k, int32_t norm_cols, int32_t hd, void* stream);
int tmae_bf16_linear_ln_fwd(const void* a, const void* w, const float* bias, const void* res, const uint8_t* rowmask, const float* gamma,
                            const float* beta, float eps, void* v, void* y, float* mean, float* rstd, int64_t m, int64_t n, int64_t k,
                            void* stream);
int tmae_bf16_linear_bwd_data(const void* dy, const void* w, const void* gelu_pre, void* dx, int64_t m, int64_t n, int64_t k, int32_t accumulate,
                              void* stream);
int tmae_bf16_linear_bwd_weight(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n, int64_t k,
                                void* stream);
int tmae_bf16_layernorm_bwd(const void* dy, const void* v, const uint8_t* rowmask, const float* gamma, const float* mean, const float* rstd,
                            void* dv, void* dres, float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int32_t c, void* stream);
int tmae_bf16_colsum(const void* x, float* out, int64_t rows, int32_t cols, void* stream);
int tmae_bf16_binned_colsum(const void* dy, const uint8_t* rowidx, float* dtable, int64_t rows, int32_t n, void* stream);
int tmae_cast_bf16_f32(const void* src, float* dst, int64_t n, void* stream);
int tmae_cast_f32_bf16(const float* src, void* dst, int64_t n, void* stream);
int tmae_scale_cast_bf16(const float* src, const float* scale, void* dst, int64_t rows, int32_t n, int32_t norm_cols, int32_t hd, void* stream);
}

namespace tmae {
int bf16_layernorm_bwd_impl(const void* dy, const void* v, const uint8_t* rowmask, const float* gamma, const float* mean, const float* rstd,
                            void* dv, void* dres, float* dgamma, float* dbeta, float* dcolsum, int64_t rows, int32_t c, bool zero, void* stream);
int bf16_colsum_impl(const void* x, float* out, int64_t rows, int32_t cols, bool zero, void* stream);
int bf16_linear_bwd_weight_impl(const void* dy, const void* x, float* dw, const void* onehot, float* dtab_t, int64_t m, int64_t n, int64_t k,
                                bool zero, void* stream);
// attention_tc.cu
int attn_tc_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s);
int attn_tc_bwd(const void* dout, const void* q, const void* k, const void* v, const void* o, const float* lse, const float* inv_q, int ld_inv_q,
                const float* inv_k, int ld_inv_k, void* dq, void* dk, void* dv, float* dtau, const tmae_layer_tables* T, const float* tau, float tau_min,
                int64_t m_q, int64_t m_kv, int c, int heads, int ldq, int ldk, int ldv, cudaStream_t s);
bool attn_tc_available();
}  // namespace tmae

namespace tmae { int g_bf16_wgrad_stream = 1; }   // tmae_set_option "wgrad_stream": weight-gradient GEMMs of the layer backward on an auxiliary stream

namespace {

typedef __nv_bfloat16 bf16;

// The weight-gradient GEMMs of a layer's backward are OFF its critical path (nothing in the layer reads dW) and, like every persistent
// GEMM here, end with a tail in which the CTAs that drew 3 tiles idle next to those that drew 4.  They are therefore enqueued on an
// auxiliary stream of the device (created on first use, kept for the life of the process; the only stream the library owns): their
// CTAs fill the tails of the data-gradient kernels on the caller's stream and vice versa.  Events order the two streams inside one
// call -- fork after the producer of an operand, join before the kernel that overwrites it and at the end of the call -- so nothing
// of the auxiliary stream is in flight when the call returns to its stream's later work.
struct Aux {
  cudaStream_t s2 = nullptr;
  cudaEvent_t fork[6] = {}, done[4] = {};
};
std::mutex g_aux_mu;
std::map<int, Aux> g_aux;
Aux* aux_stream() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(g_aux_mu);
  Aux& a = g_aux[dev];
  if (!a.s2) {
    if (cudaStreamCreateWithFlags(&a.s2, cudaStreamNonBlocking) != cudaSuccess) { a.s2 = nullptr; return nullptr; }
    for (auto& e : a.fork) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& e : a.done) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  }
  return &a;
}

struct Carve {
  char* p;
  size_t used = 0, cap;
  Carve(void* base, size_t bytes) : p((char*)base), cap(bytes) {}
  void* take(size_t bytes) {
    size_t a = (bytes + 255) / 256 * 256;
    void* r = p + used;
    used += a;
    return used <= cap ? r : nullptr;
  }
};
inline size_t csz(size_t bytes) { return (bytes + 255) / 256 * 256; }

// fp32 (64, 3C) position table of the epilogue form, or the bf16 (3C, C + 64) [W | table^T] operand of the one-hot form
inline size_t tab_bytes(int c) {
  const size_t a = (size_t)64 * 3 * c * 4, b = (size_t)3 * c * (c + 64) * 2;
  return a > b ? a : b;
}

struct Saved {
  bf16 *qkv, *kv, *o, *v1, *x1, *h, *hpre, *v2;
  float *inv_q, *inv_k, *tab, *lse, *m1, *r1, *m2, *r2;
};

size_t saved_bytes(int64_t mq, int64_t mkv, int c, int ff, int heads, bool cross) {
  size_t b = 0;
  b += csz((size_t)mq * c * (cross ? 1 : 3) * 2) + (cross ? csz((size_t)mkv * c * 2 * 2) : 0);        // q[kv], kv
  b += csz((size_t)mq * heads * (cross ? 1 : 2) * 4) + (cross ? csz((size_t)mkv * heads * 4) : 0);    // inv
  b += csz(tab_bytes(c));                                                                               // position table / [W | table] operand
  b += csz((size_t)mq * c * 2) * 4;                                                                     // o, v1, x1, v2
  b += csz((size_t)mq * heads * 4) + csz((size_t)mq * 4) * 4;                                           // lse, m1, r1, m2, r2
  b += csz((size_t)mq * ff * 2) * 2;                                                                    // h, hpre
  return b + 256;
}

bool carve_saved(Saved& s, void* buf, size_t bytes, int64_t mq, int64_t mkv, int c, int ff, int heads, bool cross) {
  Carve cv(buf, bytes);
  s.qkv = (bf16*)cv.take((size_t)mq * c * (cross ? 1 : 3) * 2);
  s.kv = cross ? (bf16*)cv.take((size_t)mkv * c * 2 * 2) : nullptr;
  s.inv_q = (float*)cv.take((size_t)mq * heads * (cross ? 1 : 2) * 4);
  s.inv_k = cross ? (float*)cv.take((size_t)mkv * heads * 4) : s.inv_q;   // self: (m, 2H) holds q heads then k heads per row
  s.tab = (float*)cv.take(tab_bytes(c));
  s.o = (bf16*)cv.take((size_t)mq * c * 2);
  s.v1 = (bf16*)cv.take((size_t)mq * c * 2);
  s.x1 = (bf16*)cv.take((size_t)mq * c * 2);
  s.v2 = (bf16*)cv.take((size_t)mq * c * 2);
  s.lse = (float*)cv.take((size_t)mq * heads * 4);
  s.m1 = (float*)cv.take((size_t)mq * 4);
  s.r1 = (float*)cv.take((size_t)mq * 4);
  s.m2 = (float*)cv.take((size_t)mq * 4);
  s.r2 = (float*)cv.take((size_t)mq * 4);
  s.h = (bf16*)cv.take((size_t)mq * ff * 2);
  s.hpre = (bf16*)cv.take((size_t)mq * ff * 2);
  return s.hpre != nullptr && cv.used <= cv.cap;
}

// fp32 temporaries of the bridge to the fp32-I/O attention kernels (attn_impl 0): q/k/v, o (forward); + dO, dq/dk/dv, dsum (backward)
size_t bridge_bytes(int64_t mq, int64_t mkv, int c, int heads, bool cross, bool bwd) {
  size_t qkv = csz((size_t)mq * c * (cross ? 1 : 3) * 4) + (cross ? csz((size_t)mkv * c * 2 * 4) : 0);
  size_t b = qkv + csz((size_t)mq * c * 4);
  if (bwd) b += qkv + csz((size_t)mq * c * 4) + csz((size_t)mq * heads * 4);
  return b;
}

size_t scratch_bytes(int64_t mq, int64_t mkv, int c, int ff, int heads, bool cross) {
  size_t b = csz((size_t)mq * c * 2) * 3;                                                               // dx1, da, dob
  b += csz((size_t)mq * c * (cross ? 1 : 3) * 2) + (cross ? csz((size_t)mkv * c * 2 * 2) : 0);          // dqkv
  b += csz((size_t)mq * ff * 2);                                                                        // dh
  b += csz((size_t)64 * 3 * c * 4);                                                                     // dtab
  b += bridge_bytes(mq, mkv, c, heads, cross, true);
  return b + 256;
}

#define TRY(call)               \
  do {                          \
    int rc__ = (call);          \
    if (rc__ != 0) return rc__; \
  } while (0)

int g_attn_impl = 1;   // 1: tcgen05 window kernel (attention_tc.cu); 0: bridge to the fp32-I/O kernels (checker / A-B runs)

}  // namespace

extern "C" {

int tmae_bf16_set_attention_impl(int32_t impl) {
  g_attn_impl = impl;   // 0 bridge, 1 tcgen05; debug: 2 = tcgen05 forward only, 3 = tcgen05 backward only
  return 0;
}
int tmae_bf16_attention_tc_available(void) { return attn_tc_available() ? 1 : 0; }

size_t tmae_bf16_encoder_layer_saved_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross) {
  const int64_t mkv = cross ? m_kv : m_q;
  return saved_bytes(m_q, mkv, c, ff, heads, cross != 0) + bridge_bytes(m_q, mkv, c, heads, cross != 0, false);
}
size_t tmae_bf16_encoder_layer_scratch_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross) {
  return scratch_bytes(m_q, cross ? m_kv : m_q, c, ff, heads, cross != 0);
}

int tmae_bf16_encoder_layer_fwd(const void* x, const void* x_kv, const tmae_layer_params* P, const tmae_bf16_weights* W, const tmae_layer_tables* T,
                                const float* pos_lut, float tau_min, float eps, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads,
                                int32_t need_backward, void* y, void* saved, size_t saved_size, void* stream) {
  const bool cross = x_kv != nullptr;
  if (!cross) { m_kv = m_q; x_kv = x; }
  cudaStream_t st = (cudaStream_t)stream;
  TMAE_CHECK_ARG(c % 128 == 0 && (c / heads == 16 || c / heads == 32) && ff % 32 == 0, "channels must be a multiple of 128, head_dim 16 or 32");
  const size_t sb = saved_bytes(m_q, m_kv, c, ff, heads, cross);
  TMAE_CHECK_ARG(saved_size >= sb + bridge_bytes(m_q, m_kv, c, heads, cross, false), "saved buffer too small");
  if (m_q <= 0) return 0;
  Saved s;
  TMAE_CHECK_ARG(carve_saved(s, saved, sb, m_q, m_kv, c, ff, heads, cross), "saved buffer carve failed");
  const int hd = c / heads;
  const int64_t cc = (int64_t)c * c;
  const bf16* in_w = (const bf16*)W->in_w;
  float* table = s.tab;
  const int ldq = cross ? c : 3 * c, ldkv = cross ? 2 * c : 3 * c;
  // q = (x + pos) Wq^T + bq, k = (x_kv + pos) Wk^T + bk, v = x_kv Wv^T + bv as ONE packed projection per source tensor with the
  // position term as a 64-row table (fp32, from the fp32 master weights) added in the TMEM epilogue; q, k leave as unit vectors
  const bool onehot = T->onehot_q != nullptr && (!cross || T->onehot_kv != nullptr) && c % 64 == 0;
  if (onehot) {
    // ... as a 64-column one-hot second A operand: [x | onehot(cell)] [W | table^T]^T, K = C + 64 (the operand the weight-gradient GEMM
    // already reads); [W | table^T] is rebuilt from the fp32 masters per call (147 KB at C = 128)
    // rows [0, C) q, [C, 2C) k (position term applied), [2C, 3C) v (bias only): one operand for the self AND the cross form
    const bf16* wcat = (const bf16*)W->in_wcat;
    if (!wcat) {
      TRY(tmae_bf16_qkv_wcat(pos_lut, P->in_w, P->in_b, s.tab, 3 * c, 2 * c, c, stream));
      wcat = (const bf16*)s.tab;
    }
    if (!cross) {
      TRY(tmae_bf16_qkv_fwd_onehot(x, T->onehot_q, wcat, s.qkv, s.inv_q, m_q, 3 * c, c, 2 * c, hd, stream));
    } else {
      const bf16* wkv = wcat + (size_t)c * (c + 64);
      TRY(tmae_bf16_qkv_fwd_onehot(x, T->onehot_q, wcat, s.qkv, s.inv_q, m_q, c, c, c, hd, stream));
      TRY(tmae_bf16_qkv_fwd_onehot(x_kv, T->onehot_kv, wkv, s.kv, s.inv_k, m_kv, 2 * c, c, c, hd, stream));
      TMAE_CUDA(cudaMemsetAsync(s.o, 0, (size_t)m_q * c * 2, st));   // rows outside paired windows
    }
  } else if (!cross) {
    TRY(tmae_pos_table(pos_lut, P->in_w, P->in_b, table, nullptr, 3 * c, 2 * c, c, stream));
    TRY(tmae_bf16_qkv_fwd(x, in_w, table, T->posidx_q, s.qkv, s.inv_q, m_q, 3 * c, c, 2 * c, hd, stream));
  } else {
    float* tkv = table + 64 * c;
    TRY(tmae_pos_table(pos_lut, P->in_w, P->in_b, table, nullptr, c, c, c, stream));
    TRY(tmae_pos_table(pos_lut, P->in_w + cc, P->in_b + c, tkv, nullptr, 2 * c, c, c, stream));
    TRY(tmae_bf16_qkv_fwd(x, in_w, table, T->posidx_q, s.qkv, s.inv_q, m_q, c, c, c, hd, stream));
    TRY(tmae_bf16_qkv_fwd(x_kv, in_w + cc, tkv, T->posidx_kv, s.kv, s.inv_k, m_kv, 2 * c, c, c, hd, stream));
    TMAE_CUDA(cudaMemsetAsync(s.o, 0, (size_t)m_q * c * 2, st));   // rows outside paired windows
  }
  const bf16* qp = s.qkv;
  const bf16* kp = cross ? s.kv : s.qkv + c;
  const bf16* vp = cross ? s.kv + c : s.qkv + 2 * c;
  if (g_attn_impl == 1 || g_attn_impl == 2) {
    TRY(attn_tc_fwd(qp, kp, vp, s.o, s.lse, T, P->tau, tau_min, m_q, m_kv, c, heads, ldq, ldkv, ldkv, st));
  } else {
    Carve br((char*)saved + sb, saved_size - sb);
    float* q32 = (float*)br.take((size_t)m_q * ldq * 4);
    float* kv32 = cross ? (float*)br.take((size_t)m_kv * ldkv * 4) : nullptr;
    float* o32 = (float*)br.take((size_t)m_q * c * 4);
    TMAE_CHECK_ARG(o32 != nullptr, "bridge carve failed");
    TRY(tmae_cast_bf16_f32(s.qkv, q32, m_q * ldq, stream));
    if (cross) TRY(tmae_cast_bf16_f32(s.kv, kv32, m_kv * ldkv, stream));
    if (cross) TMAE_CUDA(cudaMemsetAsync(o32, 0, (size_t)m_q * c * 4, st));
    const float* q_ = q32;
    const float* k_ = cross ? kv32 : q32 + c;
    const float* v_ = cross ? kv32 + c : q32 + 2 * c;
    TRY(tmae_window_attention_fwd(q_, k_, v_, o32, s.lse, T->qtok, T->qcnt, T->ktok, T->kcnt, T->n_win, T->small_end, T->mid_end, T->max_windows,
                                  P->tau, tau_min, c, heads, ldq, ldkv, ldkv, m_q, m_kv, TMAE_PREC_TF32, stream));
    TRY(tmae_cast_f32_bf16(o32, s.o, m_q * c, stream));
  }
  TRY(tmae_bf16_linear_ln_fwd(s.o, W->out_w, P->out_b, x, T->rowmask, P->ln1_g, P->ln1_b, eps, need_backward ? s.v1 : nullptr, s.x1, s.m1, s.r1,
                              m_q, c, c, stream));
  TRY(tmae_bf16_linear_fwd(s.x1, W->w1, P->b1, s.h, need_backward ? s.hpre : nullptr, m_q, ff, c, need_backward ? TMAE_ACT_GELU_DERIV : TMAE_ACT_GELU, 0,
                           stream));   // hpre receives gelu'(pre-activation): the only thing the backward needs it for
  TRY(tmae_bf16_linear_ln_fwd(s.h, W->w2, P->b2, s.x1, nullptr, P->ln2_g, P->ln2_b, eps, need_backward ? s.v2 : nullptr, y, s.m2, s.r2, m_q, c, ff,
                              stream));
  return 0;
}

int tmae_bf16_encoder_layer_bwd(const void* dy, const void* x, const void* x_kv, const tmae_layer_params* P, const tmae_bf16_weights* W,
                                const tmae_layer_tables* T, const float* pos_lut, float tau_min, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff,
                                int32_t heads, const void* saved, size_t saved_size, void* dx, void* dx_kv, const tmae_layer_params* G,
                                void* g_base, size_t g_bytes, void* scratch, size_t scratch_size, void* stream) {
  const bool cross = x_kv != nullptr;
  if (!cross) { m_kv = m_q; x_kv = x; }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t cc = (int64_t)c * c;
  if (m_q <= 0) {   // an empty query set contributes nothing: zero gradients, not whatever the caller's buffers held
    const struct { const float* p; size_t n; } z[13] = {{G->in_w, (size_t)3 * cc}, {G->in_b, (size_t)3 * c}, {G->out_w, (size_t)cc}, {G->out_b, (size_t)c},
        {G->tau, 1}, {G->ln1_g, (size_t)c}, {G->ln1_b, (size_t)c}, {G->w1, (size_t)ff * c}, {G->b1, (size_t)ff}, {G->w2, (size_t)ff * c},
        {G->b2, (size_t)c}, {G->ln2_g, (size_t)c}, {G->ln2_b, (size_t)c}};
    for (auto& e : z) if (e.p) TMAE_CUDA(cudaMemsetAsync((void*)e.p, 0, e.n * sizeof(float), st));
    if (cross && dx_kv && m_kv > 0) TMAE_CUDA(cudaMemsetAsync(dx_kv, 0, (size_t)m_kv * c * 2, st));
    return 0;
  }
  const size_t sb = saved_bytes(m_q, m_kv, c, ff, heads, cross);
  Saved s;
  TMAE_CHECK_ARG(saved_size >= sb && carve_saved(s, (void*)saved, sb, m_q, m_kv, c, ff, heads, cross), "saved buffer carve failed");
  TMAE_CHECK_ARG(scratch_size >= scratch_bytes(m_q, m_kv, c, ff, heads, cross), "scratch too small");
  Carve cv(scratch, scratch_size);
  bf16* dx1 = (bf16*)cv.take((size_t)m_q * c * 2);
  bf16* da = (bf16*)cv.take((size_t)m_q * c * 2);
  bf16* dob = (bf16*)cv.take((size_t)m_q * c * 2);
  bf16* dqkv = (bf16*)cv.take((size_t)m_q * c * (cross ? 1 : 3) * 2);
  bf16* dkv = cross ? (bf16*)cv.take((size_t)m_kv * c * 2 * 2) : nullptr;
  bf16* dh = (bf16*)cv.take((size_t)m_q * ff * 2);
  float* dtab = (float*)cv.take((size_t)64 * 3 * c * 4);
  TMAE_CHECK_ARG(dtab != nullptr, "scratch carve failed");
  // every fp32 accumulation target of this call (the 13 parameter gradients, when the caller says they are one buffer, and the
  // position-table gradient) is cleared up front: 2 memsets per layer instead of ~14
  const bool z = g_base == nullptr;   // true: each kernel entry clears its own targets
  if (!z) TMAE_CUDA(cudaMemsetAsync(g_base, 0, g_bytes, st));
  TMAE_CUDA(cudaMemsetAsync(dtab, 0, (size_t)64 * 3 * c * 4, st));
  const int hd = c / heads;
  const bf16* in_w = (const bf16*)W->in_w;
  float* g_in_w = (float*)G->in_w; float* g_in_b = (float*)G->in_b; float* g_out_w = (float*)G->out_w; float* g_out_b = (float*)G->out_b;
  float* g_tau = (float*)G->tau; float* g_ln1_g = (float*)G->ln1_g; float* g_ln1_b = (float*)G->ln1_b; float* g_w1 = (float*)G->w1;
  float* g_b1 = (float*)G->b1; float* g_w2 = (float*)G->w2; float* g_b2 = (float*)G->b2; float* g_ln2_g = (float*)G->ln2_g;
  float* g_ln2_b = (float*)G->ln2_b;

  // the weight-gradient GEMMs go to the auxiliary stream (see Aux above); z (per-kernel clears) keeps everything on one stream
  Aux* ax = (g_bf16_wgrad_stream && !z) ? aux_stream() : nullptr;
  void* wstream = ax ? (void*)ax->s2 : stream;
  auto fork = [&](int i) {        // the auxiliary stream continues behind what the caller's stream has enqueued so far
    if (ax) { cudaEventRecord(ax->fork[i], st); cudaStreamWaitEvent(ax->s2, ax->fork[i], 0); }
  };
  auto mark = [&](int i) { if (ax) cudaEventRecord(ax->done[i], ax->s2); };          // a point of the auxiliary stream ...
  auto await = [&](int i) { if (ax) cudaStreamWaitEvent(st, ax->done[i], 0); };     // ... the caller's stream waits for
  struct Join {                   // whatever path leaves the call: nothing of the auxiliary stream outlives it
    Aux* ax; cudaStream_t st;
    ~Join() { if (ax) { cudaEventRecord(ax->done[3], ax->s2); cudaStreamWaitEvent(st, ax->done[3], 0); } }
  } join_at_exit{ax, st};

  // LN2 -> FFN -> LN1 (the bias gradients of linear2 and out_proj are column sums of what the LayerNorm backward passes write)
  TRY(bf16_layernorm_bwd_impl(dy, s.v2, nullptr, P->ln2_g, s.m2, s.r2, dx1, nullptr, g_ln2_g, g_ln2_b, g_b2, m_q, c, z, stream));
  fork(0);
  TRY(bf16_linear_bwd_weight_impl(dx1, s.h, g_w2, nullptr, nullptr, m_q, c, ff, z, wstream));
  mark(0);                                                                                                    // dx1 has been read
  TRY(tmae_bf16_linear_bwd_data(dx1, W->w2, s.hpre, dh, m_q, c, ff, TMAE_BWD_PRE_IS_DERIVATIVE, stream));      // dh = (dx1 W2) * gelu'(hpre)
  fork(1);
  TRY(bf16_linear_bwd_weight_impl(dh, s.x1, g_w1, nullptr, nullptr, m_q, ff, c, z, wstream));
  TRY(bf16_colsum_impl(dh, g_b1, m_q, ff, z, stream));
  await(0);                                                                                                   // ... before it is overwritten
  TRY(tmae_bf16_linear_bwd_data(dh, W->w1, nullptr, dx1, m_q, ff, c, 1, stream));     // dx1 += dh W1: grad wrt x1 (both branches)
  TRY(bf16_layernorm_bwd_impl(dx1, s.v1, T->rowmask, P->ln1_g, s.m1, s.r1, dx, T->rowmask ? da : nullptr, g_ln1_g, g_ln1_b, g_out_b, m_q, c, z, stream));
  const bf16* dap = T->rowmask ? da : (const bf16*)dx;
  fork(2);
  TRY(bf16_linear_bwd_weight_impl(dap, s.o, g_out_w, nullptr, nullptr, m_q, c, c, z, wstream));
  mark(1);                                                                                                    // dap (= dx without a row mask) has been read
  TRY(tmae_bf16_linear_bwd_data(dap, W->out_w, nullptr, dob, m_q, c, c, 0, stream));
  // attention core: gradients land in the packed layout of the projections, already taken back through the normalisation
  const int ldq = cross ? c : 3 * c, ldkv = cross ? 2 * c : 3 * c;
  const bf16* qp = s.qkv;
  const bf16* kp = cross ? s.kv : s.qkv + c;
  const bf16* vp = cross ? s.kv + c : s.qkv + 2 * c;
  bf16* dqp = dqkv;
  bf16* dkp = cross ? dkv : dqkv + c;
  bf16* dvp = cross ? dkv + c : dqkv + 2 * c;
  if (z) TMAE_CUDA(cudaMemsetAsync(g_tau, 0, sizeof(float), st));
  if (cross) {   // rows outside paired windows get no gradient
    TMAE_CUDA(cudaMemsetAsync(dqkv, 0, (size_t)m_q * c * 2, st));
    TMAE_CUDA(cudaMemsetAsync(dkv, 0, (size_t)m_kv * 2 * c * 2, st));
  }
  const float* inv_q = s.inv_q;
  const float* inv_k = cross ? s.inv_k : s.inv_q + heads;   // self: row r holds [q heads | k heads] at r * 2H
  if (g_attn_impl == 1 || g_attn_impl == 3) {
    const int ld_inv = cross ? heads : 2 * heads;
    TRY(attn_tc_bwd(dob, qp, kp, vp, s.o, s.lse, inv_q, ld_inv, inv_k, ld_inv, dqp, dkp, dvp, g_tau, T, P->tau, tau_min, m_q, m_kv, c, heads, ldq, ldkv,
                    ldkv, st));
  } else {
    float* q32 = (float*)cv.take((size_t)m_q * ldq * 4);
    float* kv32 = cross ? (float*)cv.take((size_t)m_kv * ldkv * 4) : nullptr;
    float* o32 = (float*)cv.take((size_t)m_q * c * 4);
    float* dq32 = (float*)cv.take((size_t)m_q * ldq * 4);
    float* dkv32 = cross ? (float*)cv.take((size_t)m_kv * ldkv * 4) : nullptr;
    float* do32 = (float*)cv.take((size_t)m_q * c * 4);
    float* dsum = (float*)cv.take((size_t)m_q * heads * 4);
    TMAE_CHECK_ARG(dsum != nullptr, "bridge carve failed");
    TRY(tmae_cast_bf16_f32(s.qkv, q32, m_q * ldq, stream));
    if (cross) TRY(tmae_cast_bf16_f32(s.kv, kv32, m_kv * ldkv, stream));
    TRY(tmae_cast_bf16_f32(s.o, o32, m_q * c, stream));
    TRY(tmae_cast_bf16_f32(dob, do32, m_q * c, stream));
    if (cross) {
      TMAE_CUDA(cudaMemsetAsync(dq32, 0, (size_t)m_q * ldq * 4, st));
      TMAE_CUDA(cudaMemsetAsync(dkv32, 0, (size_t)m_kv * ldkv * 4, st));
    }
    const float* q_ = q32;
    const float* k_ = cross ? kv32 : q32 + c;
    const float* v_ = cross ? kv32 + c : q32 + 2 * c;
    float* dq_ = dq32;
    float* dk_ = cross ? dkv32 : dq32 + c;
    float* dv_ = cross ? dkv32 + c : dq32 + 2 * c;
    TRY(tmae_window_attention_bwd(do32, q_, k_, v_, o32, s.lse, dsum, dq_, dk_, dv_, g_tau, T->qtok, T->qcnt, T->ktok, T->kcnt, T->n_win,
                                  T->small_end, T->mid_end, T->max_windows, P->tau, tau_min, c, heads, ldq, ldkv, ldkv, m_q, m_kv, TMAE_PREC_TF32,
                                  stream));
    // the fp32 kernels re-normalise the (already unit) q / k rows: their dq / dk are (I - qq^T) dq_hat, still to be scaled by 1 / |q|
    if (!cross) {
      TRY(tmae_scale_cast_bf16(dq32, s.inv_q, dqkv, m_q, 3 * c, 2 * c, hd, stream));   // inv_q rows are [q heads | k heads] = column / hd
    } else {
      TRY(tmae_scale_cast_bf16(dq32, s.inv_q, dqkv, m_q, c, c, hd, stream));
      TRY(tmae_scale_cast_bf16(dkv32, s.inv_k, dkv, m_kv, 2 * c, c, hd, stream));
    }
  }
  // packed in-projection: dW = dqkv^T x, the bias and the position term from ONE binned column sum of dqkv over the 64 window
  // cells, dx += dqkv W
  // (one-hot cell indices from the plan: the binned sum is extra columns of the SAME weight-gradient GEMM, dy^T [x | onehot];
  //  without them a separate binned column-sum pass)
  // dtab_ = this projection's own (pre-zeroed) slice of the table-gradient scratch
  auto in_proj_grads = [&](const bf16* dy_, const void* xin, const uint8_t* pidx, const float* onehot, int64_t rows, int n, int n_pos, float* gw,
                           float* gb, float* dtab_) -> int {
    if (onehot) {
      TRY(bf16_linear_bwd_weight_impl(dy_, xin, gw, onehot, dtab_, rows, n, c, z, wstream));
      return tmae_pos_table_bwd(dtab_, 1, pos_lut, gw, gb, n, n_pos, c, wstream);
    }
    TRY(bf16_linear_bwd_weight_impl(dy_, xin, gw, nullptr, nullptr, rows, n, c, z, wstream));
    TRY(tmae_bf16_binned_colsum(dy_, pidx, dtab_, rows, n, wstream));
    return tmae_pos_table_bwd(dtab_, 0, pos_lut, gw, gb, n, n_pos, c, wstream);
  };
  fork(3);          // behind the attention backward: dqkv / dkv are complete
  if (!cross) {
    TRY(in_proj_grads(dqkv, x, T->posidx_q, T->onehot_q, m_q, 3 * c, 2 * c, g_in_w, g_in_b, dtab));
    await(1);       // the out_proj weight gradient has read dx (= dap) before dx += dqkv W
    TRY(tmae_bf16_linear_bwd_data(dqkv, in_w, nullptr, dx, m_q, 3 * c, c, 1, stream));
  } else {
    TRY(in_proj_grads(dqkv, x, T->posidx_q, T->onehot_q, m_q, c, c, g_in_w, g_in_b, dtab));
    await(1);
    TRY(tmae_bf16_linear_bwd_data(dqkv, in_w, nullptr, dx, m_q, c, c, 1, stream));
    TRY(in_proj_grads(dkv, x_kv, T->posidx_kv, T->onehot_kv, m_kv, 2 * c, c, g_in_w + cc, g_in_b + c, dtab + 64 * c));
    if (dx_kv) TRY(tmae_bf16_linear_bwd_data(dkv, in_w + cc, nullptr, dx_kv, m_kv, 2 * c, c, 0, stream));
  }
  return 0;
}

}  // extern "C"
