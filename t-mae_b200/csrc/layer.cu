// One encoder layer per ABI call: the host-side orchestration of an SST self-attention layer
// (pcdet/models/model_utils/sst_basic_block.py:58-84) or a WCA cross-attention layer (wca_block.py:70-103),
// forward and backward.  The reference runs ~25 small kernels per level per layer from Python; here one call
// enqueues the whole layer's kernel sequence on the stream, with every intermediate carved out of one caller-owned
// buffer that is handed back for the backward pass.
#include "common.cuh"

using namespace tmae;

namespace {

struct Carve {
  float* p;
  size_t used = 0, cap;
  Carve(void* base, size_t bytes) : p((float*)base), cap(bytes / sizeof(float)) {}
  float* take(int64_t n) {
    size_t a = (size_t)align_up(n, 64);
    float* r = p + used;
    used += a;
    return used <= cap ? r : nullptr;
  }
};
inline size_t carve_sz(int64_t n) { return (size_t)align_up(n, 64) * sizeof(float); }

struct Saved {  // forward intermediates kept for backward
  // self layer: qkv = (m, 3C) packed [q | k | v]; cross layer: q = (mq, C) and kv = (mkv, 2C) packed [k | v]
  float *qkv, *kv, *tab, *o, *lse, *a, *x1, *m1, *r1, *h, *hpre, *f, *m2, *r2;
};

bool carve_saved(Saved& s, void* buf, size_t bytes, int64_t mq, int64_t mkv, int c, int ff, int heads, bool cross) {
  Carve cv(buf, bytes);
  s.qkv = cv.take(mq * c * (cross ? 1 : 3));
  s.kv = cross ? cv.take(mkv * c * 2) : nullptr;
  s.tab = cv.take(64 * 3 * c);   // (pos_lut W^T + b) tables of the packed projections
  s.o = cv.take(mq * c);
  s.lse = cv.take(mq * heads);
  s.a = cv.take(mq * c);
  s.x1 = cv.take(mq * c);
  s.m1 = cv.take(mq);
  s.r1 = cv.take(mq);
  s.h = cv.take(mq * ff);
  s.hpre = cv.take(mq * ff);
  s.f = cv.take(mq * c);
  s.m2 = cv.take(mq);
  s.r2 = cv.take(mq);
  return s.r2 != nullptr && cv.used <= cv.cap;
}

size_t saved_bytes(int64_t mq, int64_t mkv, int c, int ff, int heads, bool cross) {
  size_t b = 0;
  b += carve_sz(mq * c * (cross ? 1 : 3)) + (cross ? carve_sz(mkv * c * 2) : 0) + carve_sz(64 * 3 * c);   // packed projections, tables
  b += carve_sz(mq * c) * 4;                       // o, a, x1, f
  b += carve_sz(mq * heads) + carve_sz(mq) * 4;    // lse, m1, r1, m2, r2
  b += carve_sz(mq * ff) * 2;                      // h, hpre
  return b + 256;
}

size_t scratch_bytes(int64_t mq, int64_t mkv, int c, int ff, int heads) {
  // dx1, da, do (mq*c each) ; dqkv (mq*3c, or mq*c + mkv*2c for a cross layer: bounded by both) ; dh (mq*ff) ;
  // dsum (mq*heads) ; position tables and their gradients (2 x 64 x 3c)
  return carve_sz(mq * c) * 3 + carve_sz(mq * c * 3) + carve_sz(mkv * c * 2) + carve_sz(mq * ff) + carve_sz(mq * heads) +
         carve_sz(64 * 3 * c) * 2 + 256;
}

#define TRY(call)              \
  do {                         \
    int rc__ = (call);         \
    if (rc__ != 0) return rc__; \
  } while (0)

}  // namespace

extern "C" {

size_t tmae_encoder_layer_saved_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross) {
  return saved_bytes(m_q, cross ? m_kv : m_q, c, ff, heads, cross != 0);
}
size_t tmae_encoder_layer_scratch_bytes(int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t cross) {
  return scratch_bytes(m_q, cross ? m_kv : m_q, c, ff, heads);
}

int tmae_encoder_layer_fwd(const float* x, const float* x_kv, const tmae_layer_params* P, const tmae_layer_tables* T, const float* pos_lut,
                           float tau_min, float eps, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads, int32_t precision,
                           int32_t need_backward, float* y, void* saved, size_t saved_size, void* stream) {
  const bool cross = x_kv != nullptr;
  if (!cross) { m_kv = m_q; x_kv = x; }
  TMAE_CHECK_ARG(saved_size >= saved_bytes(m_q, m_kv, c, ff, heads, cross), "saved buffer too small");
  if (m_q <= 0) return 0;
  Saved s;
  TMAE_CHECK_ARG(carve_saved(s, saved, saved_size, m_q, m_kv, c, ff, heads, cross), "saved buffer carve failed");
  // q = (x + pos) Wq^T + bq, k = (x_kv + pos) Wk^T + bk, v = x_kv Wv^T + bv  (sst_basic_block.py:44 ; wca_block.py:52-56;
  // cosine_msa.py:57-62) as ONE packed projection per source tensor: the position embedding depends only on the voxel's
  // cell in its window, so (pos W^T + b) is a 64-row table added per row in the GEMM epilogue -- x + pos is never
  // materialised and x is read once instead of three times.
  const int64_t cc = (int64_t)c * c;
  float* table = s.tab;
  const int ldq = cross ? c : 3 * c, ldkv = cross ? 2 * c : 3 * c;
  // tensor-core mode with one-hot cell indices from the plan: the table term is a second (one-hot, table^T) source pair of
  // the same GEMM; otherwise (fp32 parity mode) the table is added per row in the SIMT epilogue.  Same arithmetic.
  const bool dual = precision == TMAE_PREC_TF32 && T->onehot_q && (!cross || T->onehot_kv);
  if (!cross) {
    TRY(tmae_pos_table(pos_lut, P->in_w, P->in_b, dual ? nullptr : table, dual ? table : nullptr, 3 * c, 2 * c, c, stream));
    if (dual) TRY(tmae_linear_fwd_dual(x, P->in_w, T->onehot_q, table, s.qkv, m_q, 3 * c, c, 64, precision, stream));
    else TRY(tmae_linear_fwd_lut(x, P->in_w, table, T->posidx_q, s.qkv, m_q, 3 * c, c, precision, stream));
  } else {
    float* tkv = table + 64 * c;
    TRY(tmae_pos_table(pos_lut, P->in_w, P->in_b, dual ? nullptr : table, dual ? table : nullptr, c, c, c, stream));
    TRY(tmae_pos_table(pos_lut, P->in_w + cc, P->in_b + c, dual ? nullptr : tkv, dual ? tkv : nullptr, 2 * c, c, c, stream));
    if (dual) {
      TRY(tmae_linear_fwd_dual(x, P->in_w, T->onehot_q, table, s.qkv, m_q, c, c, 64, precision, stream));
      TRY(tmae_linear_fwd_dual(x_kv, P->in_w + cc, T->onehot_kv, tkv, s.kv, m_kv, 2 * c, c, 64, precision, stream));
    } else {
      TRY(tmae_linear_fwd_lut(x, P->in_w, table, T->posidx_q, s.qkv, m_q, c, c, precision, stream));
      TRY(tmae_linear_fwd_lut(x_kv, P->in_w + cc, tkv, T->posidx_kv, s.kv, m_kv, 2 * c, c, precision, stream));
    }
    TMAE_CUDA(cudaMemsetAsync(s.o, 0, (size_t)m_q * c * sizeof(float), (cudaStream_t)stream));  // rows outside paired windows
  }
  const float* qp = s.qkv;
  const float* kp = cross ? s.kv : s.qkv + c;
  const float* vp = cross ? s.kv + c : s.qkv + 2 * c;
  TRY(tmae_window_attention_fwd(qp, kp, vp, s.o, s.lse, T->qtok, T->qcnt, T->ktok, T->kcnt, T->n_win, T->small_end, T->mid_end, T->max_windows,
                                P->tau, tau_min, c, heads, ldq, ldkv, ldkv, m_q, m_kv, precision, stream));
  TRY(tmae_linear_fwd(s.o, P->out_w, P->out_b, nullptr, s.a, nullptr, m_q, c, c, TMAE_ACT_NONE, precision, stream));
  TRY(tmae_add_layernorm_fwd(x, s.a, T->rowmask, P->ln1_g, P->ln1_b, s.x1, s.m1, s.r1, m_q, c, eps, stream));
  // the pre-activation copy exists only for the GELU backward: inference skips that write (2 FF-wide rows per voxel)
  TRY(tmae_linear_fwd(s.x1, P->w1, P->b1, nullptr, s.h, need_backward ? s.hpre : nullptr, m_q, ff, c, TMAE_ACT_GELU, precision, stream));
  TRY(tmae_linear_fwd(s.h, P->w2, P->b2, nullptr, s.f, nullptr, m_q, c, ff, TMAE_ACT_NONE, precision, stream));
  TRY(tmae_add_layernorm_fwd(s.x1, s.f, nullptr, P->ln2_g, P->ln2_b, y, s.m2, s.r2, m_q, c, eps, stream));
  return 0;
}

int tmae_encoder_layer_bwd(const float* dy, const float* x, const float* x_kv, const tmae_layer_params* P, const tmae_layer_tables* T,
                           const float* pos_lut, float tau_min, int64_t m_q, int64_t m_kv, int32_t c, int32_t ff, int32_t heads,
                           int32_t precision, const void* saved, size_t saved_size, float* dx, float* dx_kv, const tmae_layer_params* G,
                           void* scratch, size_t scratch_size, void* stream) {
  const bool cross = x_kv != nullptr;
  if (!cross) { m_kv = m_q; x_kv = x; }
  cudaStream_t st = (cudaStream_t)stream;
  if (m_q <= 0) {
    // an empty query set (e.g. a sample with < 4 voxels under a 75 % mask, or an empty stage) contributes nothing:
    // every parameter gradient and the key/value input gradient are ZERO, not whatever the caller's buffers held
    const size_t cc_ = (size_t)c * c;
    const struct { const float* p; size_t n; } z[13] = {{G->in_w, 3 * cc_}, {G->in_b, (size_t)3 * c}, {G->out_w, cc_}, {G->out_b, (size_t)c},
        {G->tau, 1}, {G->ln1_g, (size_t)c}, {G->ln1_b, (size_t)c}, {G->w1, (size_t)ff * c}, {G->b1, (size_t)ff}, {G->w2, (size_t)ff * c},
        {G->b2, (size_t)c}, {G->ln2_g, (size_t)c}, {G->ln2_b, (size_t)c}};
    for (auto& e : z) if (e.p) TMAE_CUDA(cudaMemsetAsync((void*)e.p, 0, e.n * sizeof(float), st));
    if (cross && dx_kv && m_kv > 0) TMAE_CUDA(cudaMemsetAsync(dx_kv, 0, (size_t)m_kv * c * sizeof(float), st));
    return 0;
  }
  Saved s;
  TMAE_CHECK_ARG(carve_saved(s, (void*)saved, saved_size, m_q, m_kv, c, ff, heads, cross), "saved buffer carve failed");
  TMAE_CHECK_ARG(scratch_size >= scratch_bytes(m_q, m_kv, c, ff, heads), "scratch too small");
  Carve cv(scratch, scratch_size);
  float* dx1 = cv.take(m_q * c);
  float* da = cv.take(m_q * c);
  float* dob = cv.take(m_q * c);
  float* dqkv = cv.take(m_q * c * (cross ? 1 : 3));   // self: (m, 3C) packed [dq | dk | dv] ; cross: dq (mq, C)
  float* dkv = cross ? cv.take(m_kv * c * 2) : nullptr;  // cross: (mkv, 2C) packed [dk | dv]
  float* dh = cv.take(m_q * ff);
  float* dsum = cv.take(m_q * heads);
  float* dtab = cv.take(64 * 3 * c);
  TMAE_CHECK_ARG(dtab != nullptr, "scratch carve failed");
  const int64_t cc = (int64_t)c * c;
  // G holds the gradient buffers with the same field meaning as P (const-cast: the struct type is shared)
  float* g_in_w = (float*)G->in_w; float* g_in_b = (float*)G->in_b; float* g_out_w = (float*)G->out_w; float* g_out_b = (float*)G->out_b;
  float* g_tau = (float*)G->tau; float* g_ln1_g = (float*)G->ln1_g; float* g_ln1_b = (float*)G->ln1_b; float* g_w1 = (float*)G->w1;
  float* g_b1 = (float*)G->b1; float* g_w2 = (float*)G->w2; float* g_b2 = (float*)G->b2; float* g_ln2_g = (float*)G->ln2_g;
  float* g_ln2_b = (float*)G->ln2_b;

  // LN2 -> FFN -> LN1
  // the bias gradients of linear2 and out_proj are column sums of what the two LayerNorm backward passes write: they
  // accumulate them on the way instead of a separate pass over the (m, c) gradient
  TRY(tmae_add_layernorm_bwd_colsum(dy, s.x1, s.f, nullptr, P->ln2_g, s.m2, s.r2, dx1, nullptr, g_ln2_g, g_ln2_b, g_b2, m_q, c, stream));
  TRY(tmae_linear_bwd_weight(dx1, s.h, g_w2, nullptr, m_q, c, ff, precision, stream));
  TRY(tmae_linear_bwd_data_gelu(dx1, P->w2, s.hpre, dh, m_q, c, ff, precision, stream));
  TRY(tmae_linear_bwd_weight(dh, s.x1, g_w1, g_b1, m_q, ff, c, precision, stream));
  TRY(tmae_linear_bwd_data(dh, P->w1, dx1, m_q, ff, c, 1, precision, stream));  // dx1 = grad wrt x1 (both branches)
  TRY(tmae_add_layernorm_bwd_colsum(dx1, x, s.a, T->rowmask, P->ln1_g, s.m1, s.r1, dx, T->rowmask ? da : nullptr, g_ln1_g, g_ln1_b, g_out_b,
                                    m_q, c, stream));
  const float* dap = T->rowmask ? da : dx;  // grad wrt the attention branch (masked rows contribute nothing)
  // out projection
  TRY(tmae_linear_bwd_weight(dap, s.o, g_out_w, nullptr, m_q, c, c, precision, stream));
  TRY(tmae_linear_bwd_data(dap, P->out_w, dob, m_q, c, c, 0, precision, stream));
  // attention core: gradients land in the packed layout of the projections
  const int ldq = cross ? c : 3 * c, ldkv = cross ? 2 * c : 3 * c;
  const float* qp = s.qkv;
  const float* kp = cross ? s.kv : s.qkv + c;
  const float* vp = cross ? s.kv + c : s.qkv + 2 * c;
  float* dqp = dqkv;
  float* dkp = cross ? dkv : dqkv + c;
  float* dvp = cross ? dkv + c : dqkv + 2 * c;
  TMAE_CUDA(cudaMemsetAsync(g_tau, 0, sizeof(float), st));
  if (cross) {  // rows outside paired windows get no gradient
    TMAE_CUDA(cudaMemsetAsync(dqkv, 0, (size_t)m_q * c * sizeof(float), st));
    TMAE_CUDA(cudaMemsetAsync(dkv, 0, (size_t)m_kv * 2 * c * sizeof(float), st));
  }
  TRY(tmae_window_attention_bwd(dob, qp, kp, vp, s.o, s.lse, dsum, dqp, dkp, dvp, g_tau, T->qtok, T->qcnt, T->ktok, T->kcnt, T->n_win,
                                T->small_end, T->mid_end, T->max_windows, P->tau, tau_min, c, heads, ldq, ldkv, ldkv, m_q, m_kv, precision, stream));
  // packed in-projection: dW = dqkv^T x (+ the position term), db and the position term from ONE binned column sum of
  // dqkv over the 64 window cells, dx += dqkv W
  // (dual: the binned sum is the weight-gradient GEMM against the one-hot matrix, dtab = dy^T onehot, (n, 64))
  const bool dual = precision == TMAE_PREC_TF32 && T->onehot_q && (!cross || T->onehot_kv);
  auto table_grad = [&](const float* dy_, const uint8_t* pidx, const float* onehot, int64_t rows, int n, int n_pos, float* gw, float* gb) -> int {
    if (dual) TRY(tmae_linear_bwd_weight(dy_, onehot, dtab, nullptr, rows, n, 64, precision, stream));
    else TRY(tmae_binned_colsum(dy_, pidx, dtab, rows, n, stream));
    return tmae_pos_table_bwd(dtab, dual ? 1 : 0, pos_lut, gw, gb, n, n_pos, c, stream);
  };
  if (!cross) {
    TRY(tmae_linear_bwd_weight(dqkv, x, g_in_w, nullptr, m_q, 3 * c, c, precision, stream));
    TRY(table_grad(dqkv, T->posidx_q, T->onehot_q, m_q, 3 * c, 2 * c, g_in_w, g_in_b));
    TRY(tmae_linear_bwd_data(dqkv, P->in_w, dx, m_q, 3 * c, c, 1, precision, stream));
  } else {
    TRY(tmae_linear_bwd_weight(dqkv, x, g_in_w, nullptr, m_q, c, c, precision, stream));
    TRY(table_grad(dqkv, T->posidx_q, T->onehot_q, m_q, c, c, g_in_w, g_in_b));
    TRY(tmae_linear_bwd_data(dqkv, P->in_w, dx, m_q, c, c, 1, precision, stream));
    TRY(tmae_linear_bwd_weight(dkv, x_kv, g_in_w + cc, nullptr, m_kv, 2 * c, c, precision, stream));
    TRY(table_grad(dkv, T->posidx_kv, T->onehot_kv, m_kv, 2 * c, c, g_in_w + cc, g_in_b + c));
    if (dx_kv) TRY(tmae_linear_bwd_data(dkv, P->in_w + cc, dx_kv, m_kv, 2 * c, c, 0, precision, stream));
  }
  return 0;
}

}  // extern "C"
