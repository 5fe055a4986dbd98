// Host-side orchestration of the actor / critic forward and backward passes, the T-step rollout and the A2C
// gradient computation. Every arithmetic op is a kernel from gemm_tc.cu / attention.cu / kernels_simt.cu / env.cu;
// this file only sequences launches on one stream and carves the caller-provided workspace.
//
// Reference call graph restated here:
//   Encoder.call            common/encoder.py:67-99        -> encoder_forward / encoder_backward
//   EncoderLayer.call       common/encoder_layer.py:15-23  -> enc_layer_forward / enc_layer_backward
//   Decoder.call            actor/decoder.py:62-101        -> decoder_forward / decoder_backward
//   CriticTransformer.call  critic/model.py:55-85          -> critic_forward / critic_backward
//   trainer loop            agents/trainer.py:49-156       -> tpc_rollout / tpc_a2c_grads
#include <cstring>
#include <vector>

#include "model.cuh"
#include "kernels.cuh"
#include "env.cuh"
#include "decide.cuh"

using namespace tpc;

namespace {

struct Arena {
  char* base = nullptr;
  size_t size = 0, off = 0, peak = 0;
  bool dry = false;
  float* f(size_t nfloats) {
    const size_t bytes = (nfloats * sizeof(float) + 1023) & ~size_t(1023);
    char* p = base + off;
    off += bytes;
    if (off > peak) peak = off;
    return reinterpret_cast<float*>(p);
  }
  bool fits() const { return dry || peak <= size; }
};

struct Ctx {
  tpc_model* m;
  const NetLayout* L;
  const float* P;   // canonical parameters
  const float* W;   // weight image
  float* G;         // canonical gradient accumulator (nullptr in forward-only calls)
  Arena* ar;
  cudaStream_t st;
  const struct SideStream* side = nullptr;   // rollout only: second stream for the two independent encoder stacks
  bool dry() const { return ar->dry; }
  TmapCache* tc() const { return m->h->tmaps; }
  int sms() const { return m->h->num_sms; }
};

inline int rollout_fork_mode() {   // read per call: tests switch it inside one process
  const char* v = getenv("TPC_ROLLOUT_FORK");
  return v ? atoi(v) : 1;
}
// fork / join helper: work issued on `stream` between fork() and join() runs beside the caller's stream (also while
// that stream is being captured into a graph: the event pair becomes two graph edges)
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int fork(cudaStream_t main) const {
    TPC_CHECK_CUDA(cudaEventRecord(ev_fork, main));
    TPC_CHECK_CUDA(cudaStreamWaitEvent(stream, ev_fork, 0));
    return TPC_OK;
  }
  int join(cudaStream_t main) const {
    TPC_CHECK_CUDA(cudaEventRecord(ev_join, stream));
    TPC_CHECK_CUDA(cudaStreamWaitEvent(main, ev_join, 0));
    return TPC_OK;
  }
};

struct MaskRef { const float* mask; int mS; int m0; int split = 1 << 30; int gap = 0; };  // mask[c * mS + m0 + (j < split ? j : j + gap)]

struct EncLayerActs {
  float *qkv, *lse, *att, *out1, *rstd1, *h, *out2, *rstd2;
  uint32_t* hmask;        // ReLU bit image of h: rows x dff / 32 words
  bool has_mask = false;  // written by the forward pass (weight-resident FFN1 GEMM of a training context)
};
struct EncActs {
  float *emb_b = nullptr, *emb_r = nullptr, *cat = nullptr;
  std::vector<EncLayerActs> bins, res, joint;
  float* enc_out = nullptr;
};
struct StateRef { const float* state; int ld; const float* extra; };  // (C,S,ld) [+ is_empty (C,S)]

#define RUN(expr)                         \
  do {                                    \
    if (!cx.dry()) { TPC_TRY(expr); }     \
  } while (0)

int gemm(Ctx& cx, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N, int K,
         const GemmEpi& e, int ks = 1) {
  if (cx.dry()) return TPC_OK;
  return launch_gemm(cx.tc(), A, lda, Bt, ldb, D, ldd, M, N, K, e, ks, cx.sms(), cx.st);
}

// forward GEMM: fp32-accurate FP16 split product; Bt must be a "_t" entry of the weight image, outputs are not rounded
int gemm_fwd(Ctx& cx, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N, int K,
             GemmEpi e, int ks = 1) {
  if (cx.dry()) return TPC_OK;
  e.round_out = 0;
  return launch_gemm(cx.tc(), A, lda, Bt, ldb, D, ldd, M, N, K, e, ks, cx.sms(), cx.st, Bt + cx.L->shadow_half);
}

// dW (+ db) of a Dense layer whose output gradient is dY[M, Nout]; W canonical [Kin][Nout]
int dense_grads(Ctx& cx, const float* X, int ldx, const float* dY, int ldy, int M, const DenseP& d) {
  if (cx.dry()) return TPC_OK;
  WgradDst dst;
  const int nj = ceil_div(d.out, 128);
  for (int j = 0; j < nj; ++j) { dst.ptr[j] = cx.G + d.w + j * 128; dst.ld[j] = d.out; dst.bias[j] = cx.G + d.b + j * 128; }
  return launch_wgrad(cx.tc(), X, ldx, dY, ldy, M, d.in, d.out, dst, cx.sms(), cx.st);
}

// fused output gradient dY[M, 128 * n] feeding n Dense layers of 128 outputs each (q|k|v or k2|v2|W2)
int fused_grads(Ctx& cx, const float* X, int ldx, const float* dY, int ldy, int M, const DenseP* const* ds, int n) {
  if (cx.dry()) return TPC_OK;
  WgradDst dst;
  for (int j = 0; j < n; ++j) { dst.ptr[j] = cx.G + ds[j]->w; dst.ld[j] = 128; dst.bias[j] = cx.G + ds[j]->b; }
  return launch_wgrad(cx.tc(), X, ldx, dY, ldy, M, 128, 128 * n, dst, cx.sms(), cx.st);
}

GemmEpi epi_bias(const float* b) { GemmEpi e; e.bias = b; return e; }
GemmEpi epi_ln(const Ctx& cx, const float* b, const float* resid, const LnP& ln, float* rstd) {
  GemmEpi e;
  e.bias = b; e.aux = resid; e.ldaux = 128; e.aux_mode = 1; e.ln = 1;
  e.gamma = cx.P + ln.g; e.beta = cx.P + ln.b; e.rstd_out = rstd;
  return e;
}
GemmEpi epi_add(const float* aux, int ld) { GemmEpi e; e.aux = aux; e.ldaux = ld; e.aux_mode = 1; return e; }

// ------------------------------------------------------------------------------------------------------
// encoder
// ------------------------------------------------------------------------------------------------------
void alloc_layer(Arena& ar, EncLayerActs& a, int rows, int dff) {
  a.qkv = ar.f((size_t)rows * 384);
  a.lse = ar.f((size_t)rows * 8);
  a.att = ar.f((size_t)rows * 128);
  a.out1 = ar.f((size_t)rows * 128);
  a.rstd1 = ar.f(rows);
  a.h = ar.f((size_t)rows * dff);
  a.hmask = reinterpret_cast<uint32_t*>(ar.f((size_t)((rows + 127) / 128 * 128) * (dff / 32)));   // [dff / 128][rows padded][4]
  a.out2 = ar.f((size_t)rows * 128);
  a.rstd2 = ar.f(rows);
}

int enc_layer_forward(Ctx& cx, const EncLayerP& lp, const float* x, int C, int Lt, MaskRef mk, EncLayerActs& a) {
  const int rows = C * Lt, dff = cx.L->dff;
  alloc_layer(*cx.ar, a, rows, dff);
  TPC_TRY(gemm_fwd(cx, x, 128, cx.W + lp.s.wqkv_t, 128, a.qkv, 384, rows, 384, 128, epi_bias(cx.W + lp.s.bqkv)));
  RUN(attention_fwd(a.qkv, mk.mask, mk.mS, mk.m0, C, Lt, a.att, a.lse, cx.st, mk.split, mk.gap));
  TPC_TRY(gemm_fwd(cx, a.att, 128, cx.W + lp.s.wo_t, 128, a.out1, 128, rows, 128, 128,
               epi_ln(cx, cx.P + lp.mha.o.b, x, lp.ln1, a.rstd1)));
  GemmEpi e1 = epi_bias(cx.P + lp.f1.b);
  e1.relu = 1;
  // training context: the FFN1 epilogue also leaves the sign of h as a bit image for the ReLU backward
  a.has_mask = cx.G != nullptr && dff % 128 == 0 && gemm_weight_resident(rows, dff, 128, e1, true, cx.sms());
  if (a.has_mask) { e1.mask_out = a.hmask; e1.mask_ld = (rows + 127) / 128 * 128; }
  TPC_TRY(gemm_fwd(cx, a.out1, 128, cx.W + lp.s.w1_t, 128, a.h, dff, rows, dff, 128, e1));
  TPC_TRY(gemm_fwd(cx, a.h, dff, cx.W + lp.s.w2_t, dff, a.out2, 128, rows, 128, dff,
               epi_ln(cx, cx.P + lp.f2.b, a.out1, lp.ln2, a.rstd2)));
  return TPC_OK;
}

struct BwdScratch { float *dA, *dB, *dQKV; };

// LayerNorm backward fused into the epilogue of the GEMM that produces the gradient of the LayerNorm output
// (GemmEpi::lnb_y, gemm_tc.cu): TPC_FUSE_LNB=0 falls back to the stand-alone ln_bwd launches (A / B runs, tests).
bool fuse_lnb() {
  static const bool on = [] { const char* v = getenv("TPC_FUSE_LNB"); return !(v && v[0] == '0'); }();
  return on;
}
GemmEpi epi_lnb(const Ctx& cx, const float* aux, const float* y, const float* rstd, const LnP& ln) {
  GemmEpi e;
  e.aux = aux; e.ldaux = 128; e.aux_mode = 1;
  e.gamma = cx.P + ln.g; e.beta = cx.P + ln.b;
  e.lnb_y = y; e.lnb_rstd = rstd; e.lnb_dgamma = cx.G + ln.g; e.lnb_dbeta = cx.G + ln.b;
  return e;
}
struct PrevLn { const LnP* ln; const float* rstd; };   // the LayerNorm that produced this layer's input x

// dout: gradient w.r.t. out2, rows addressed through dmap -- or, with dout_is_dpre2, already the gradient w.r.t. the
// input of LayerNorm 2 (the layer above fused that LayerNorm backward into its dx GEMM). dx receives the gradient
// w.r.t. x, or with prev != nullptr the gradient w.r.t. the input of the LayerNorm that produced x.
int enc_layer_backward(Ctx& cx, const EncLayerP& lp, const float* x, int C, int Lt, MaskRef mk, EncLayerActs& a,
                       const float* dout, RowMap dmap, float* dx, BwdScratch& s, bool dout_is_dpre2 = false,
                       const PrevLn* prev = nullptr) {
  const int rows = C * Lt, dff = cx.L->dff;
  const RowMap ident{Lt, Lt, 0};
  const float* dpre2 = dout;
  if (!dout_is_dpre2) {
    RUN(ln_bwd(dout, dmap, a.out2, a.rstd2, cx.P + lp.ln2.g, cx.P + lp.ln2.b, rows, s.dA, cx.G + lp.ln2.g,
               cx.G + lp.ln2.b, cx.st));
    dpre2 = s.dA;
  }
  TPC_TRY(dense_grads(cx, a.h, dff, dpre2, 128, rows, lp.f2));
  {  // dH = (dpre2 @ W2^T) * (h > 0), written over h
    GemmEpi e;
    if (a.has_mask && gemm_weight_resident(rows, dff, 128, e, false, cx.sms())) {
      e.mask_in = a.hmask; e.mask_ld = (rows + 127) / 128 * 128;   // 16 B per row tile instead of re-reading 512 B of h
    } else {
      e.aux = a.h; e.ldaux = dff; e.aux_mode = 2;
    }
    TPC_TRY(gemm(cx, dpre2, 128, cx.W + lp.s.w2_c, 128, a.h, dff, rows, dff, 128, e));
  }
  TPC_TRY(dense_grads(cx, a.out1, 128, a.h, dff, rows, lp.f1));
  const float* dpre1;
  float* tmp;
  if (fuse_lnb()) {
    // dout1 = dH @ W1^T + dpre2 never reaches HBM: the epilogue turns it into dpre1
    TPC_TRY(gemm(cx, a.h, dff, cx.W + lp.s.w1_c, dff, s.dB, 128, rows, 128, dff,
                 epi_lnb(cx, dpre2, a.out1, a.rstd1, lp.ln1)));
    dpre1 = s.dB; tmp = s.dA;
  } else {
    TPC_TRY(gemm(cx, a.h, dff, cx.W + lp.s.w1_c, dff, s.dB, 128, rows, 128, dff, epi_add(dpre2, 128)));  // dout1
    RUN(ln_bwd(s.dB, ident, a.out1, a.rstd1, cx.P + lp.ln1.g, cx.P + lp.ln1.b, rows, s.dA, cx.G + lp.ln1.g,
               cx.G + lp.ln1.b, cx.st));
    dpre1 = s.dA; tmp = s.dB;
  }
  TPC_TRY(dense_grads(cx, a.att, 128, dpre1, 128, rows, lp.mha.o));
  TPC_TRY(gemm(cx, dpre1, 128, cx.W + lp.s.wo_c, 128, tmp, 128, rows, 128, 128, GemmEpi()));          // datt
  RUN(attention_bwd(a.qkv, a.att, a.lse, tmp, mk.mask, mk.mS, mk.m0, C, Lt, s.dQKV, cx.st, mk.split, mk.gap));
  const DenseP* qkv[3] = {&lp.mha.q, &lp.mha.k, &lp.mha.v};
  TPC_TRY(fused_grads(cx, x, 128, s.dQKV, 384, rows, qkv, 3));
  TPC_TRY(gemm(cx, s.dQKV, 384, cx.W + lp.s.wqkv_c, 384, dx, 128, rows, 128, 384,
               prev ? epi_lnb(cx, dpre1, x, prev->rstd, *prev->ln) : epi_add(dpre1, 128)));  // dx (or dpre2 of the layer below)
  return TPC_OK;
}

// backward through one stack of encoder layers (top layer first). d_top: gradient w.r.t. the stack's output, rows through
// dm; x0: the stack's input; bufs: two ping-pong buffers. Returns the buffer holding the gradient w.r.t. x0.
int enc_stack_backward(Ctx& cx, const std::vector<EncLayerP>& lps, std::vector<EncLayerActs>& acts, const float* x0,
                       int C, int Lt, MaskRef mk, const float* d_top, RowMap dm, float* const bufs[2], BwdScratch& s,
                       const float** d_x0) {
  const int nl = (int)lps.size();
  const float* dc = d_top;
  bool is_dpre2 = false;
  int bi = 0;
  for (int i = nl - 1; i >= 0; --i) {
    const float* x = i == 0 ? x0 : acts[i - 1].out2;
    PrevLn pv;
    const bool fuse_prev = i > 0 && fuse_lnb();
    if (fuse_prev) { pv.ln = &lps[i - 1].ln2; pv.rstd = acts[i - 1].rstd2; }
    TPC_TRY(enc_layer_backward(cx, lps[i], x, C, Lt, mk, acts[i], dc, dm, bufs[bi], s, is_dpre2, fuse_prev ? &pv : nullptr));
    dc = bufs[bi];
    dm = RowMap{Lt, Lt, 0};
    is_dpre2 = fuse_prev;
    bi ^= 1;
  }
  *d_x0 = dc;
  return TPC_OK;
}

// r0 > 0 (actor only): the first r0 rule tokens of every state are left out of the compact buffers (TokWin, kernels.cuh)
int encoder_forward(Ctx& cx, StateRef sr, const float* mha_mask, int C, int N, int R, EncActs& a, int r0 = 0) {
  const NetLayout& L = *cx.L;
  const int S = N + R, nl = L.nl, Re = R - r0, Se = N + Re;
  Arena& ar = *cx.ar;
  a.emb_b = ar.f((size_t)C * N * 128);
  a.emb_r = ar.f((size_t)C * Re * 128);
  const DenseP& eb = L.enc.emb1;
  const DenseP& er = L.enc.common ? L.enc.emb1 : L.enc.emb2;
  // The nodes stack and the rules stack are independent until the concatenation: the rollout runs the rules stack on a
  // second stream beside the nodes stack (TPC_ROLLOUT_FORK=0: one stream). Small batches leave most SMs idle in every
  // kernel of a stack (config #1: 11.98 -> 11.38 ms per iteration); at batch 1024 the two stacks fill each other's tails
  // (greedy decode of the tester cell 79.3 -> 77.4 ms).
  const bool fork = cx.side && !cx.dry() && rollout_fork_mode() != 0;
  Ctx cr = cx;                       // context of the rules stack
  if (fork) {
    cr.st = cx.side->stream;
    TPC_TRY(cx.side->fork(cx.st));
  }
  a.bins.resize(nl); a.res.resize(nl); a.joint.resize(nl);
  a.cat = ar.f((size_t)C * Se * 128);
  RUN(small_dense_fwd(sr.state, sr.ld, sr.extra, C * N, RowMap{N, S, 0}, cx.P + eb.w, cx.P + eb.b, eb.in, a.emb_b, cx.st));
  const float* xb = a.emb_b;
  for (int i = 0; i < nl; ++i) {
    TPC_TRY(enc_layer_forward(cx, L.enc.bins[i], xb, C, N, MaskRef{mha_mask, S, 0}, a.bins[i]));
    xb = a.bins[i].out2;
  }
  RUN(scatter_rows(xb, C * N, RowMap{N, Se, 0}, a.cat, cx.st));
  RUN(small_dense_fwd(sr.state, sr.ld, sr.extra, C * Re, RowMap{Re, S, N + r0}, cx.P + er.w, cx.P + er.b, er.in, a.emb_r, cr.st));
  const float* xr = a.emb_r;
  for (int i = 0; i < nl; ++i) {
    TPC_TRY(enc_layer_forward(cr, L.enc.res[i], xr, C, Re, MaskRef{mha_mask, S, N + r0}, a.res[i]));
    xr = a.res[i].out2;
  }
  RUN(scatter_rows(xr, C * Re, RowMap{Re, Se, N}, a.cat, cr.st));
  if (fork) TPC_TRY(cx.side->join(cx.st));
  const float* xj = a.cat;
  for (int i = 0; i < nl; ++i) {
    TPC_TRY(enc_layer_forward(cx, L.enc.joint[i], xj, C, Se, MaskRef{mha_mask, S, 0, N, r0}, a.joint[i]));
    xj = a.joint[i].out2;
  }
  a.enc_out = const_cast<float*>(xj);
  return TPC_OK;
}

// d_enc: gradient w.r.t. the encoder output [C*S,128] (overwritten as scratch)
int encoder_backward(Ctx& cx, StateRef sr, const float* mha_mask, int C, int N, int R, EncActs& a, float* d_enc,
                     int r0 = 0) {
  const NetLayout& L = *cx.L;
  const int S = N + R, nl = L.nl, Re = R - r0, Se = N + Re;
  Arena& ar = *cx.ar;
  BwdScratch s;
  s.dA = ar.f((size_t)C * Se * 128);
  s.dB = ar.f((size_t)C * Se * 128);
  s.dQKV = ar.f((size_t)C * Se * 384);
  float* dx0 = ar.f((size_t)C * Se * 128);
  float* dx1 = ar.f((size_t)C * Se * 128);
  float* pp[2] = {dx0, dx1};
  const float* dcat = nullptr;       // [C*Se,128]
  TPC_TRY(enc_stack_backward(cx, L.enc.joint, a.joint, a.cat, C, Se, MaskRef{mha_mask, S, 0, N, r0}, d_enc, RowMap{Se, Se, 0},
                             pp, s, &dcat));
  float* other = dcat == dx0 ? dx1 : dx0;      // free ping-pong buffer
  float* third = const_cast<float*>(d_enc);    // d_enc is consumed: reuse as a third buffer
  float* bufs[2] = {other, third};
  const float* dc = nullptr;
  // bins stack
  TPC_TRY(enc_stack_backward(cx, L.enc.bins, a.bins, a.emb_b, C, N, MaskRef{mha_mask, S, 0}, dcat, RowMap{N, Se, 0}, bufs, s, &dc));
  const DenseP& eb = L.enc.emb1;
  RUN(small_dense_bwd(sr.state, sr.ld, sr.extra, C * N, RowMap{N, S, 0}, dc, eb.in, cx.G + eb.w, cx.G + eb.b, cx.st));
  // resources stack
  TPC_TRY(enc_stack_backward(cx, L.enc.res, a.res, a.emb_r, C, Re, MaskRef{mha_mask, S, N + r0}, dcat, RowMap{Re, Se, N}, bufs, s,
                             &dc));
  const DenseP& er = L.enc.common ? L.enc.emb1 : L.enc.emb2;
  RUN(small_dense_bwd(sr.state, sr.ld, sr.extra, C * Re, RowMap{Re, S, N + r0}, dc, er.in, cx.G + er.w, cx.G + er.b, cx.st));
  return TPC_OK;
}

// ------------------------------------------------------------------------------------------------------
// actor decoder + pointer
// ------------------------------------------------------------------------------------------------------
struct DecLayerActs { float *v1, *o1, *rstd1, *q2, *kvw, *pc, *ctx, *o2, *rstd2, *f, *o3, *rstd3; };
struct ActorActs {
  EncActs enc;
  float* y0 = nullptr;
  std::vector<DecLayerActs> dec;
  float *w1d = nullptr, *score = nullptr;
};

// fuse != nullptr (rollout): the pointer head runs inside the fused decision kernel (decide.cu) together with the
// pick, the environment step and the staging of the next step's inputs
int decoder_forward(Ctx& cx, const float* dec_input, const float* ptr_mask, const float* mha_mask, int C, int N, int R,
                    int r0, ActorActs& a, float* logits, float* probs, int32_t* argmax, DecideArgs* fuse = nullptr) {
  const NetLayout& L = *cx.L;
  Arena& ar = *cx.ar;
  const int dff = L.dff, nl = L.nl;
  const int S = N + R - r0;                 // compact tokens per state
  TokWin win;
  win.N = N; win.gap = r0; win.Sfull = N + R;
  a.y0 = ar.f((size_t)C * 128);
  a.dec.resize(nl);
  a.w1d = ar.f((size_t)C * 128);
  a.score = ar.f((size_t)C * S);
  const float* y = nullptr;                 // first layer: the chain kernel embeds dec_input itself
  for (int i = 0; i < nl; ++i) {
    const DecLayerP& lp = L.dec[i];
    DecLayerActs& d = a.dec[i];
    d.v1 = ar.f((size_t)C * 128); d.o1 = ar.f((size_t)C * 128); d.rstd1 = ar.f(C);
    d.q2 = ar.f((size_t)C * 128); d.kvw = ar.f((size_t)C * S * 384); d.pc = ar.f((size_t)C * 8 * S);
    d.ctx = ar.f((size_t)C * 128); d.o2 = ar.f((size_t)C * 128); d.rstd2 = ar.f(C);
    d.f = ar.f((size_t)C * dff); d.o3 = ar.f((size_t)C * 128); d.rstd3 = ar.f(C);
    // K2 | V2 | W2e of every encoder token: the one GEMM-shaped product of the decoder (tensor cores, fp32-accurate split product)
    TPC_TRY(gemm_fwd(cx, a.enc.enc_out, 128, cx.W + lp.s.wkvp_t, 128, d.kvw, 384, C * S, 384, 128,
                 epi_bias(cx.W + lp.s.bkvp)));
    // everything on the single decoder token: one fused kernel (csrc/kernels.cuh, DecChainArgs)
    DecChainArgs ca{};
    ca.dec_input = dec_input; ca.F = L.demb.in; ca.y_in = y;
    ca.Wemb = cx.P + L.demb.w; ca.bemb = cx.P + L.demb.b;
    ca.Wv1 = cx.P + lp.mha1.v.w; ca.bv1 = cx.P + lp.mha1.v.b; ca.Wo1 = cx.P + lp.mha1.o.w; ca.bo1 = cx.P + lp.mha1.o.b;
    ca.g1 = cx.P + lp.ln1.g; ca.be1 = cx.P + lp.ln1.b;
    ca.Wq2 = cx.P + lp.mha2.q.w; ca.bq2 = cx.P + lp.mha2.q.b; ca.Wo2 = cx.P + lp.mha2.o.w; ca.bo2 = cx.P + lp.mha2.o.b;
    ca.g2 = cx.P + lp.ln2.g; ca.be2 = cx.P + lp.ln2.b;
    ca.Wf1 = cx.P + lp.f1.w; ca.bf1 = cx.P + lp.f1.b; ca.Wf2 = cx.P + lp.f2.w; ca.bf2 = cx.P + lp.f2.b; ca.dff = dff;
    ca.g3 = cx.P + lp.ln3.g; ca.be3 = cx.P + lp.ln3.b;
    const bool last = i == nl - 1;
    ca.Wp1 = last ? cx.P + L.W1.w : nullptr; ca.bp1 = last ? cx.P + L.W1.b : nullptr;
    ca.kvw = d.kvw; ca.mask = mha_mask; ca.S = S; ca.win = win;
    ca.y0 = a.y0; ca.v1 = d.v1; ca.o1 = d.o1; ca.rstd1 = d.rstd1; ca.q2 = d.q2; ca.pc = d.pc; ca.ctx = d.ctx;
    ca.o2 = d.o2; ca.rstd2 = d.rstd2; ca.f = d.f; ca.o3 = d.o3; ca.rstd3 = d.rstd3; ca.w1d = a.w1d;
    ca.C = C; ca.ln_eps = 1e-6f;
    RUN(decoder_chain_fwd(ca, cx.st));
    y = d.o3;
  }
  if (fuse) {
    fuse->w1d = a.w1d; fuse->kvw = a.dec[nl - 1].kvw; fuse->V = cx.P + L.V.w; fuse->bV = cx.P + L.V.b;
    fuse->clipC = cx.m->dims.clip_C; fuse->has_clip = cx.m->dims.has_clip; fuse->ptr_mask = ptr_mask;
    fuse->score = nullptr; fuse->logits = logits; fuse->win = win;
    RUN(decide_step(*fuse, C, cx.st));
    return TPC_OK;
  }
  RUN(pointer_fwd(a.w1d, a.dec[nl - 1].kvw, cx.P + L.V.w, cx.P + L.V.b, cx.m->dims.clip_C, cx.m->dims.has_clip, ptr_mask,
                  C, S, a.score, logits, probs, argmax, cx.st, win));
  return TPC_OK;
}

// dlogits [C,S] -> parameter gradients of the decoder; d_enc [C*S,128] receives the encoder-output gradient
int decoder_backward(Ctx& cx, const float* dec_input, int C, int N, int R, int r0, ActorActs& a, const float* dlogits,
                     float* d_enc) {
  const NetLayout& L = *cx.L;
  Arena& ar = *cx.ar;
  const int dff = L.dff, nl = L.nl;
  const int S = N + R - r0;                 // compact tokens per state
  TokWin win;
  win.N = N; win.gap = r0; win.Sfull = N + R;
  float* dA = ar.f((size_t)C * 128);
  float* dB = ar.f((size_t)C * 128);
  float* dC = ar.f((size_t)C * 128);
  float* dY = ar.f((size_t)C * 128);       // gradient w.r.t. the current layer's output o3
  float* dq2 = ar.f((size_t)C * 128);
  float* dw1d = ar.f((size_t)C * 128);
  float* dkvw = ar.f((size_t)C * S * 384);
  const RowMap id1{1, 1, 0};
  DecLayerActs& last = a.dec[nl - 1];
  if (!cx.dry()) TPC_CHECK_CUDA(cudaMemsetAsync(dkvw, 0, (size_t)C * S * 384 * sizeof(float), cx.st));
  RUN(pointer_bwd(dlogits, a.score, a.w1d, last.kvw, cx.P + L.V.w, cx.m->dims.clip_C, cx.m->dims.has_clip, C, S, dw1d,
                  dkvw, cx.G + L.V.w, cx.G + L.V.b, cx.st, win));
  TPC_TRY(dense_grads(cx, last.o3, 128, dw1d, 128, C, L.W1));
  TPC_TRY(gemm(cx, dw1d, 128, cx.W + L.w1p_c, 128, dY, 128, C, 128, 128, GemmEpi()));
  bool enc_written = false;
  for (int i = nl - 1; i >= 0; --i) {
    const DecLayerP& lp = L.dec[i];
    DecLayerActs& d = a.dec[i];
    const float* y = i == 0 ? a.y0 : a.dec[i - 1].o3;
    RUN(ln_bwd(dY, id1, d.o3, d.rstd3, cx.P + lp.ln3.g, cx.P + lp.ln3.b, C, dA, cx.G + lp.ln3.g, cx.G + lp.ln3.b, cx.st));
    TPC_TRY(dense_grads(cx, d.f, dff, dA, 128, C, lp.f2));
    {
      GemmEpi e;
      e.aux = d.f; e.ldaux = dff; e.aux_mode = 2;
      TPC_TRY(gemm(cx, dA, 128, cx.W + lp.s.w2_c, 128, d.f, dff, C, dff, 128, e));
    }
    TPC_TRY(dense_grads(cx, d.o2, 128, d.f, dff, C, lp.f1));
    TPC_TRY(gemm(cx, d.f, dff, cx.W + lp.s.w1_c, dff, dB, 128, C, 128, dff, epi_add(dA, 128)));        // do2
    RUN(ln_bwd(dB, id1, d.o2, d.rstd2, cx.P + lp.ln2.g, cx.P + lp.ln2.b, C, dA, cx.G + lp.ln2.g, cx.G + lp.ln2.b, cx.st));
    TPC_TRY(dense_grads(cx, d.ctx, 128, dA, 128, C, lp.mha2.o));
    TPC_TRY(gemm(cx, dA, 128, cx.W + lp.s.wo2_c, 128, dC, 128, C, 128, 128, GemmEpi()));               // dctx
    if (i != nl - 1 && !cx.dry()) TPC_CHECK_CUDA(cudaMemsetAsync(dkvw, 0, (size_t)C * S * 384 * sizeof(float), cx.st));
    RUN(cross_attn_bwd(dC, d.q2, d.kvw, d.pc, C, S, dq2, dkvw, cx.st));
    TPC_TRY(dense_grads(cx, d.o1, 128, dq2, 128, C, lp.mha2.q));
    TPC_TRY(gemm(cx, dq2, 128, cx.W + lp.s.wq2_c, 128, dB, 128, C, 128, 128, epi_add(dA, 128)));       // do1
    RUN(ln_bwd(dB, id1, d.o1, d.rstd1, cx.P + lp.ln1.g, cx.P + lp.ln1.b, C, dA, cx.G + lp.ln1.g, cx.G + lp.ln1.b, cx.st));
    TPC_TRY(dense_grads(cx, d.v1, 128, dA, 128, C, lp.mha1.o));
    TPC_TRY(gemm(cx, dA, 128, cx.W + lp.s.wo1_c, 128, dC, 128, C, 128, 128, GemmEpi()));               // dv1
    TPC_TRY(dense_grads(cx, y, 128, dC, 128, C, lp.mha1.v));
    TPC_TRY(gemm(cx, dC, 128, cx.W + lp.s.wv1_c, 128, dY, 128, C, 128, 128, epi_add(dA, 128)));        // dy
    // encoder-output gradient through [K2 | V2 | W2] and their weight gradients
    const int nf = (i == nl - 1) ? 3 : 2;
    const DenseP* kvp[3] = {&lp.mha2.k, &lp.mha2.v, &L.W2};
    TPC_TRY(fused_grads(cx, a.enc.enc_out, 128, dkvw, 384, C * S, kvp, nf));
    TPC_TRY(gemm(cx, dkvw, 384, cx.W + lp.s.wkvp_c, 384, d_enc, 128, C * S, 128, 128 * nf,
                 enc_written ? epi_add(d_enc, 128) : GemmEpi()));
    enc_written = true;
  }
  RUN(small_dense_bwd(dec_input, L.demb.in, nullptr, C, id1, dY, L.demb.in, cx.G + L.demb.w, cx.G + L.demb.b, cx.st));
  return TPC_OK;
}

// r0: rule tokens [0, r0) of every state are left out (TokWin); the caller guarantees they are masked in mha_mask AND
// ptr_mask for all C states (rollout: rules placed at earlier steps; update: checked on the device, lead_masked_kernel)
int actor_forward(Ctx& cx, StateRef sr, const float* dec_input, const float* ptr_mask, const float* mha_mask, int C,
                  int N, int R, ActorActs& a, float* logits, float* probs, int32_t* argmax, DecideArgs* fuse = nullptr,
                  int r0 = 0) {
  TPC_TRY(encoder_forward(cx, sr, mha_mask, C, N, R, a.enc, r0));
  return decoder_forward(cx, dec_input, ptr_mask, mha_mask, C, N, R, r0, a, logits, probs, argmax, fuse);
}

int actor_backward(Ctx& cx, StateRef sr, const float* dec_input, const float* mha_mask, int C, int N, int R,
                   ActorActs& a, const float* dlogits, int r0 = 0) {
  float* d_enc = cx.ar->f((size_t)C * (N + R - r0) * 128);
  TPC_TRY(decoder_backward(cx, dec_input, C, N, R, r0, a, dlogits, d_enc));
  return encoder_backward(cx, sr, mha_mask, C, N, R, a.enc, d_enc, r0);
}

// Token dropping is exact only when a dropped token's logit is exactly -1e9, i.e. |C tanh(score)| < ulp(1e9) / 2 = 32
bool can_drop_tokens(const tpc_model* m) { return m->dims.has_clip && fabsf(m->dims.clip_C) <= 30.f; }

// lead[k] = min over the states of chunk k of the number of leading rule tokens that are hidden BOTH as attention keys
// (mha_mask) and as pointer targets (ptr_mask): the window the actor may leave out for that chunk.
__global__ void lead_masked_kernel(const float* __restrict__ mha, const float* __restrict__ ptr, long long n, int S, int N,
                                   int chunk, int* __restrict__ lead) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* m = mha + i * S + N;
  const float* p = ptr + i * S + N;
  int k = 0;
  while (k < S - N && m[k] != 0.f && p[k] != 0.f) ++k;
  atomicMin(lead + (int)(i / chunk), k);
}

// ------------------------------------------------------------------------------------------------------
// critic
// ------------------------------------------------------------------------------------------------------
struct CriticActs { EncActs enc; float *h0 = nullptr, *h1 = nullptr; };

int critic_forward(Ctx& cx, StateRef sr, const float* mha_mask, int C, int N, int R, CriticActs& a, float* values) {
  const NetLayout& L = *cx.L;
  const int S = N + R;
  if (S != L.S_critic) {
    fprintf(stderr, "[tpc] critic is tied to the training shape S=%d (critic/model.py:65), got S=%d\n", L.S_critic, S);
    return TPC_ERR_SHAPE;
  }
  TPC_TRY(encoder_forward(cx, sr, mha_mask, C, N, R, a.enc));
  a.h0 = cx.ar->f((size_t)C * 128);
  a.h1 = cx.ar->f((size_t)C * 128);
  const int K = S * 128;
  RUN(fill_rows_bias(a.h0, cx.P + L.h0.b, C, cx.st));
  GemmEpi e;
  e.atomic = 1; e.round_out = 0;
  int ks = ceil_div(cx.sms(), ceil_div(C, 128));
  if (ks > K / 32 / 4) ks = K / 32 / 4;
  if (ks < 1) ks = 1;
  TPC_TRY(gemm_fwd(cx, a.enc.enc_out, K, cx.W + L.h0_t, K, a.h0, 128, C, 128, K, e, ks));   // Flatten -> Dense(128)
  TPC_TRY(gemm_fwd(cx, a.h0, 128, cx.W + L.h1_t, 128, a.h1, 128, C, 128, 128, epi_bias(cx.P + L.h1.b)));
  RUN(value_head_fwd(a.h1, cx.P + L.hv.w, cx.P + L.hv.b, C, values, cx.st));
  return TPC_OK;
}

int critic_backward(Ctx& cx, StateRef sr, const float* mha_mask, int C, int N, int R, CriticActs& a, const float* dv) {
  const NetLayout& L = *cx.L;
  const int S = N + R, K = S * 128;
  Arena& ar = *cx.ar;
  float* dh1 = ar.f((size_t)C * 128);
  float* dh0 = ar.f((size_t)C * 128);
  float* d_enc = ar.f((size_t)C * S * 128);
  RUN(value_head_bwd(dv, a.h1, cx.P + L.hv.w, C, dh1, cx.G + L.hv.w, cx.G + L.hv.b, cx.st));
  TPC_TRY(dense_grads(cx, a.h0, 128, dh1, 128, C, L.h1));
  TPC_TRY(gemm(cx, dh1, 128, cx.W + L.h1_c, 128, dh0, 128, C, 128, 128, GemmEpi()));
  TPC_TRY(dense_grads(cx, a.enc.enc_out, K, dh0, 128, C, L.h0));
  TPC_TRY(gemm(cx, dh0, 128, cx.W + L.h0_c, 128, d_enc, K, C, K, 128, GemmEpi()));
  return encoder_backward(cx, sr, mha_mask, C, N, R, a.enc, d_enc);
}

__global__ void gather_dec_kernel(const float* __restrict__ batch, int B, int S, int F, int tok, float* __restrict__ dec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * F) return;
  const int b = i / F, f = i % F;
  dec[i] = batch[((size_t)b * S + tok) * F + f];
}

__global__ void scatter_reward_kernel(const float* __restrict__ r, int B, int T, int t, float* __restrict__ rewards) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b < B) rewards[(size_t)b * T + t] = r[b];
}

Ctx make_ctx(tpc_model* m, int which, const float* params, const float* shadow, float* grads, Arena* ar, void* stream) {
  Ctx cx;
  cx.m = m;
  cx.L = which ? &m->critic : &m->actor;
  cx.P = params; cx.W = shadow; cx.G = grads; cx.ar = ar; cx.st = (cudaStream_t)stream;
  return cx;
}

int a2c_grads_impl(tpc_model* m, const float* actor_shadow, const float* actor_params, const float* critic_shadow,
                   const float* critic_params, const tpc_rollout_buffers* buf, int B, int N, int R,
                   const tpc_a2c_hyper* hp, float* actor_grads, float* critic_grads, float* values, float* advantages,
                   float* losses, Arena& ar, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int S = N + R, T = R, F = m->cfg.dec_features;
  const long long n = (long long)T * B;
  // default: ~10 k states per chunk, all chunks equal (fewer, fuller launches: 1.68 s -> 1.52 s per update at the
  // headline shape vs 2 k chunks); ~5.8 MB of saved activations per state for the 9-layer critic
  int chunk = hp->chunk_states;
  if (chunk <= 0) {
    const long long pieces = (n + 10239) / 10240;
    chunk = (int)((n + pieces - 1) / pieces);
  }
  if (chunk > n) chunk = (int)n;
  const float inv_total = 1.0f / (float)(hp->total_states > 0 ? hp->total_states : n);
  float* Rt = ar.f((size_t)n);
  float* dv = ar.f((size_t)chunk);
  float* dlogits = ar.f((size_t)chunk * S);
  float* logits = ar.f((size_t)chunk * S);
  const size_t mark = ar.off;
  if (!ar.dry) TPC_TRY(discounted_returns(buf->rewards, B, T, hp->gamma, Rt, st));
  const bool cost = buf->is_empties != nullptr;
  const bool do_critic = hp->phases == 0 || (hp->phases & TPC_A2C_CRITIC) || ar.dry;
  const bool do_actor = hp->phases == 0 || (hp->phases & TPC_A2C_ACTOR) || ar.dry;
  // ---- critic: value loss + gradients (agents/agent.py:126-166, trainer.py:115-123)
  for (long long s0 = 0; do_critic && s0 < n; s0 += chunk) {
    const int C = (int)((n - s0) < chunk ? (n - s0) : chunk);
    ar.off = mark;
    Ctx cx = make_ctx(m, 1, critic_params, critic_shadow, critic_grads, &ar, stream);
    StateRef sr{buf->states + (size_t)s0 * S * F, F, cost ? buf->is_empties + (size_t)s0 * S : nullptr};
    const float* mha = buf->mha_masks + (size_t)s0 * S;
    CriticActs a;
    TPC_TRY(critic_forward(cx, sr, mha, C, N, R, a, values + s0));
    RUN(critic_loss_bwd(Rt + s0, values + s0, C, hp->values_loss_coefficient, inv_total, advantages + s0, dv, losses, st));
    TPC_TRY(critic_backward(cx, sr, mha, C, N, R, a, dv));
    if (!ar.fits()) return TPC_ERR_WORKSPACE;
    if (ar.dry) break;
  }
  // ---- actor: policy loss (+ zero-gradient entropy term) and gradients (agents/agent.py:168-242, trainer.py:125-141)
  // rule tokens hidden in every state of a chunk (placed at earlier steps) are left out of the actor's buffers: their
  // count per chunk is measured on the device from the stored masks (one small read-back per update)
  const int n_chunks = (int)((n + chunk - 1) / chunk);
  std::vector<int> lead(n_chunks, 0);
  if (do_actor && !ar.dry && can_drop_tokens(m)) {
    int* d_lead = reinterpret_cast<int*>(dlogits);   // free until the first actor chunk writes it
    TPC_CHECK_CUDA(cudaMemsetAsync(d_lead, 0x7f, n_chunks * sizeof(int), st));
    lead_masked_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(buf->mha_masks, buf->ptr_masks, n, S, N, chunk, d_lead);
    TPC_CHECK_LAUNCH();
    TPC_CHECK_CUDA(cudaMemcpyAsync(lead.data(), d_lead, n_chunks * sizeof(int), cudaMemcpyDeviceToHost, st));
    TPC_CHECK_CUDA(cudaStreamSynchronize(st));
    for (int& l : lead) l = l < 0 ? 0 : (l > R - 1 ? R - 1 : l);
  }
  for (long long s0 = 0; do_actor && s0 < n; s0 += chunk) {
    const int C = (int)((n - s0) < chunk ? (n - s0) : chunk);
    const int r0 = lead[(int)(s0 / chunk)];
    ar.off = mark;
    Ctx cx = make_ctx(m, 0, actor_params, actor_shadow, actor_grads, &ar, stream);
    StateRef sr{buf->states + (size_t)s0 * S * F, F, cost ? buf->is_empties + (size_t)s0 * S : nullptr};
    const float* mha = buf->mha_masks + (size_t)s0 * S;
    const float* dec = buf->dec_inputs + (size_t)s0 * F;
    ActorActs a;
    TPC_TRY(actor_forward(cx, sr, dec, buf->ptr_masks + (size_t)s0 * S, mha, C, N, R, a, logits, nullptr, nullptr, nullptr, r0));
    RUN(actor_loss_bwd(logits, buf->actions + s0, advantages + s0, C, S, hp->entropy_coefficient, inv_total, dlogits,
                       losses, st));
    TPC_TRY(actor_backward(cx, sr, dec, mha, C, N, R, a, dlogits, r0));
    if (!ar.fits()) return TPC_ERR_WORKSPACE;
    if (ar.dry) break;
  }
  return TPC_OK;
}

// ---- cached CUDA graphs of the rollout's step loop (see rollout_impl)
struct RolloutKey {
  tpc_env_config env;
  const void* p[15];
  size_t ws_bytes;
  uint64_t seed;
  int64_t id0;
  int B, N, R, greedy, drop, fork_mode;
};
struct RolloutGraphs {
  struct Slot { RolloutKey key; cudaGraphExec_t exec = nullptr; int seen = 0; unsigned long long stamp = 0, nodes = 0; };
  std::vector<Slot> slots;
  cudaStream_t cap_stream = nullptr;
  int64_t* d_draw0 = nullptr;
  unsigned long long clock = 0, replays = 0;
  bool broken = false;   // a capture failed once: stream launches from then on
  bool unusable = false; // streams / events could not be created
  SideStream side_cap, side_run;   // second stream of the encoder fork / join: under capture, on the caller's stream
  // the slot of this argument set (created, or recycled from the least recently used one, when it is new)
  Slot* find(const RolloutKey& k) {
    ++clock;
    for (Slot& s : slots)
      if (memcmp(&s.key, &k, sizeof(k)) == 0) { s.stamp = clock; return &s; }
    if (slots.size() < 8) {
      slots.emplace_back();
    } else {
      size_t lru = 0;
      for (size_t i = 1; i < slots.size(); ++i) if (slots[i].stamp < slots[lru].stamp) lru = i;
      if (slots[lru].exec) cudaGraphExecDestroy(slots[lru].exec);
      std::swap(slots[lru], slots.back());
      slots.back() = Slot();
    }
    Slot& s = slots.back();
    s.key = k; s.stamp = clock;
    return &s;
  }
};
// per-model rollout state: graph cache, capture stream, the side streams of encoder_forward's fork / join
RolloutGraphs* rollout_graphs(tpc_model* m) {
  if (!m->rollout_graphs) {
    RolloutGraphs* rg = new RolloutGraphs();
    bool ok = cudaStreamCreateWithFlags(&rg->cap_stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaMalloc(&rg->d_draw0, sizeof(int64_t)) == cudaSuccess;
    for (SideStream* sd : {&rg->side_cap, &rg->side_run})
      ok = ok && cudaStreamCreateWithFlags(&sd->stream, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&sd->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&sd->ev_join, cudaEventDisableTiming) == cudaSuccess;
    m->rollout_graphs = rg;
    m->rollout_graphs_free = [](void* p) {
      RolloutGraphs* g = static_cast<RolloutGraphs*>(p);
      for (auto& s : g->slots) if (s.exec) cudaGraphExecDestroy(s.exec);
      for (SideStream* sd : {&g->side_cap, &g->side_run}) {
        if (sd->ev_fork) cudaEventDestroy(sd->ev_fork);
        if (sd->ev_join) cudaEventDestroy(sd->ev_join);
        if (sd->stream) cudaStreamDestroy(sd->stream);
      }
      if (g->cap_stream) cudaStreamDestroy(g->cap_stream);
      cudaFree(g->d_draw0);
      delete g;
    };
    if (!ok) {
      cudaGetLastError();
      rg->unusable = true;
    }
  }
  RolloutGraphs* rg = static_cast<RolloutGraphs*>(m->rollout_graphs);
  return rg->unusable ? nullptr : rg;
}
// TPC_ROLLOUT_GRAPH=0 keeps the stream launches (A / B runs, debugging); read per call: tests switch it inside one process
// Under a CUDA injection tool (Nsight Compute, compute-sanitizer: CUDA_INJECTION64_PATH / NV_COMPUTE_PROFILER_* in the
// environment) the loop stays on the stream as well: ncu cannot profile launches that are being captured and shuts the
// process down (seen with `ncu --metrics gpu__time_duration.sum python bench.py`), and a profile should list the kernels.
bool rollout_graph_wanted() {
  const char* v = getenv("TPC_ROLLOUT_GRAPH");
  if (v && v[0] == '0') return false;
  if (v && v[0] == '2') return true;   // forced (debugging the capture path under a tool)
  return !getenv("CUDA_INJECTION64_PATH") && !getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR");
}

int rollout_impl(tpc_model* m, const tpc_env_config* e, const float* actor_shadow, const float* actor_params,
                 const tpc_rollout_buffers* buf, int B, int N, int R, int greedy, uint64_t seed, int64_t id0,
                 int64_t epoch, Arena& ar, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool cvrp = e->kind == TPC_ENV_CVRP;
  const int S = N + R, F = e->num_features;
  const int T = cvrp ? N - 1 + R : R;                     // CVRP adapter: N - 1 node visits + R returns to the depot
  auto req_at = [&](int t) { return cvrp ? (t < N - 1 ? t + 1 : 0) : N + t; };   // CVRP: oracle/cvrp.py request_node
  float* dec_tmp = ar.f((size_t)B * F);
  float* mask_tmp = ar.f((size_t)B * S);
  const size_t mark = ar.off;
  auto dec_at = [&](int t) { return buf->dec_inputs ? buf->dec_inputs + (size_t)t * B * F : dec_tmp; };
  auto pmask_at = [&](int t) { return buf->ptr_masks ? buf->ptr_masks + (size_t)t * B * S : mask_tmp; };
  if (!ar.dry) {
    // inputs of step 0: decoder input = first rule, Agent.store copies (agents/trainer.py:68-76), feasible mask.
    // For every later step the fused decision kernel of step t - 1 stages them.
    gather_dec_kernel<<<ceil_div(B * F, 256), 256, 0, st>>>(buf->batch, B, S, F, req_at(0), dec_at(0));
    TPC_CHECK_LAUNCH();
    if (buf->states)
      TPC_CHECK_CUDA(cudaMemcpyAsync(buf->states, buf->batch, (size_t)B * S * F * 4, cudaMemcpyDeviceToDevice, st));
    if (buf->is_empties && buf->is_empty)
      TPC_CHECK_CUDA(cudaMemcpyAsync(buf->is_empties, buf->is_empty, (size_t)B * S * 4, cudaMemcpyDeviceToDevice, st));
    if (buf->mha_masks)
      TPC_CHECK_CUDA(cudaMemcpyAsync(buf->mha_masks, buf->mha_mask, (size_t)B * S * 4, cudaMemcpyDeviceToDevice, st));
    TPC_TRY(env_feasible_mask(*e, buf->batch, F, dec_at(0), buf->bin_net_mask, B, S, pmask_at(0), st));
  }
  // At step t the rules 0 .. t-1 are placed: hidden as attention keys by the env step itself and, as long as the rule
  // columns of bin_net_mask are set (checked once here), never a pointer target -> the actor leaves them out.
  bool drop = false;
  if (!ar.dry && !cvrp && can_drop_tokens(m)) {
    int* d_flag = reinterpret_cast<int*>(mask_tmp);
    int h_flag = 0;
    TPC_CHECK_CUDA(cudaMemsetAsync(d_flag, 0x7f, sizeof(int), st));
    lead_masked_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(buf->bin_net_mask, buf->bin_net_mask, B, S, N, B, d_flag);
    TPC_CHECK_LAUNCH();
    TPC_CHECK_CUDA(cudaMemcpyAsync(&h_flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    TPC_CHECK_CUDA(cudaStreamSynchronize(st));
    drop = h_flag >= R;
  }
  // The T decoding steps. draw0 == nullptr: the Philox draw index epoch * T + t is a kernel argument; otherwise the
  // kernels read epoch * T from *draw0 (graph replay: everything else about a step is the same from call to call).
  RolloutGraphs* rg = ar.dry ? nullptr : rollout_graphs(m);
  auto run_steps = [&](cudaStream_t rs, const int64_t* draw0) -> int {
    for (int t = 0; t < T; ++t) {
      ar.off = mark;
      Ctx cx = make_ctx(m, 0, actor_params, actor_shadow, nullptr, &ar, rs);
      if (rg) cx.side = draw0 ? &rg->side_cap : &rg->side_run;
      DecideArgs da{};
      if (!ar.dry) {
        da.greedy = greedy; da.seed = seed; da.id0 = id0;
        da.draw = draw0 ? t : epoch * T + t;
        da.draw0 = draw0;
        da.actions = buf->actions + (size_t)t * B;
        da.env = *e;
        da.batch = buf->batch; da.bin_mask = buf->bin_net_mask; da.mha_mask = buf->mha_mask; da.is_empty = buf->is_empty;
        da.N = N; da.R = R; da.t = t; da.T = T; da.req = req_at(t); da.req_next = req_at(t + 1 < T ? t + 1 : t);
        da.rewards = buf->rewards;
        if (t + 1 < T) {
          da.dec_next = dec_at(t + 1);
          da.pmask_next = pmask_at(t + 1);
          da.states_next = buf->states ? buf->states + (size_t)(t + 1) * B * S * F : nullptr;
          da.mha_next = buf->mha_masks ? buf->mha_masks + (size_t)(t + 1) * B * S : nullptr;
          da.is_empties_next = (buf->is_empties && buf->is_empty) ? buf->is_empties + (size_t)(t + 1) * B * S : nullptr;
        }
      }
      ActorActs a;
      StateRef sr{buf->batch, F, buf->is_empty};
      TPC_TRY(actor_forward(cx, sr, ar.dry ? nullptr : dec_at(t), ar.dry ? nullptr : pmask_at(t), buf->mha_mask, B, N, R, a,
                            buf->logits ? buf->logits + (size_t)t * B * S : nullptr, nullptr, nullptr, &da, drop ? t : 0));
      if (!ar.fits()) return TPC_ERR_WORKSPACE;
      if (ar.dry) break;
    }
    return TPC_OK;
  };
  if (ar.dry) return run_steps(st, nullptr);

  // ---- CUDA-graph replay of the step loop (T x ~22 launches; launch-bound for small batches). Every kernel argument
  // of the loop is a function of the call's arguments alone (workspace offsets, trajectory slots, tensor maps), except
  // the Philox draw index, which the graph reads from device memory. A call whose arguments match a cached graph
  // replays it; a graph is built the second time an argument set is seen (one-off callers never pay for a capture).
  if (!rg || !rollout_graph_wanted()) return run_steps(st, nullptr);
  RolloutKey key;
  memset(&key, 0, sizeof(key));   // padding bytes take part in the comparison
  key.env = *e;
  key.p[0] = actor_shadow; key.p[1] = actor_params; key.p[2] = ar.base; key.p[3] = buf->batch; key.p[4] = buf->bin_net_mask;
  key.p[5] = buf->mha_mask; key.p[6] = buf->is_empty; key.p[7] = buf->states; key.p[8] = buf->dec_inputs;
  key.p[9] = buf->ptr_masks; key.p[10] = buf->mha_masks; key.p[11] = buf->is_empties; key.p[12] = buf->actions;
  key.p[13] = buf->rewards; key.p[14] = buf->logits;
  key.ws_bytes = ar.size; key.seed = seed; key.id0 = id0;
  key.B = B; key.N = N; key.R = R; key.greedy = greedy; key.drop = drop ? 1 : 0; key.fork_mode = rollout_fork_mode();
  RolloutGraphs::Slot* slot = rg->find(key);
  if (slot && !slot->exec && slot->seen >= 1 && !rg->broken) {
    // capture on the cache's own non-blocking stream (the caller's may be the legacy default stream, which cannot be
    // captured); nothing executes during capture, so no ordering with `st` is needed until the launch below
    TPC_TRY(gemm_tc_init());
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(rg->cap_stream, cudaStreamCaptureModeThreadLocal);
    int rc = TPC_OK;
    if (ce == cudaSuccess) {
      const unsigned long long l0 = g_tpc_launches;
      rc = run_steps(rg->cap_stream, rg->d_draw0);
      slot->nodes = g_tpc_launches - l0;   // recorded, not launched: counted per replay below (tpc_launch_count)
      g_tpc_launches = l0;
      ce = cudaStreamEndCapture(rg->cap_stream, &graph);
    }
    if (ce == cudaSuccess && rc == TPC_OK && graph) ce = cudaGraphInstantiate(&slot->exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != TPC_OK || !slot->exec) {
      fprintf(stderr, "[tpc] rollout graph capture failed (%s, rc %d): falling back to stream launches\n",
              cudaGetErrorString(ce), rc);
      cudaGetLastError();
      slot->exec = nullptr;
      rg->broken = true;
    }
  }
  if (slot && slot->exec) {
    const int64_t draw_base = epoch * T;
    TPC_CHECK_CUDA(cudaMemcpyAsync(rg->d_draw0, &draw_base, sizeof(draw_base), cudaMemcpyHostToDevice, st));
    TPC_CHECK_CUDA(cudaGraphLaunch(slot->exec, st));
    g_tpc_launches += slot->nodes;
    ++rg->replays;
    return TPC_OK;
  }
  if (slot) ++slot->seen;
  return run_steps(st, nullptr);
}

}  // namespace

extern "C" {

size_t tpc_workspace_bytes(const tpc_model* mc, int mode, int nstates, int num_bins, int num_resources) {
  tpc_model* m = const_cast<tpc_model*>(mc);
  if (!m || nstates <= 0) return 0;
  Arena ar;
  ar.dry = true;
  const int N = num_bins, R = num_resources, C = nstates;
  int rc = TPC_OK;
  if (mode == 0 || mode == 2) {
    Ctx cx = make_ctx(m, 0, nullptr, nullptr, nullptr, &ar, nullptr);
    ActorActs a;
    rc = actor_forward(cx, StateRef{nullptr, 3, nullptr}, nullptr, nullptr, nullptr, C, N, R, a, nullptr, nullptr, nullptr);
    if (rc == TPC_OK && mode == 2) rc = actor_backward(cx, StateRef{nullptr, 3, nullptr}, nullptr, nullptr, C, N, R, a, nullptr);
  } else if (mode == 1 || mode == 3) {
    Ctx cx = make_ctx(m, 1, nullptr, nullptr, nullptr, &ar, nullptr);
    CriticActs a;
    rc = critic_forward(cx, StateRef{nullptr, 3, nullptr}, nullptr, C, N, R, a, nullptr);
    if (rc == TPC_OK && mode == 3) rc = critic_backward(cx, StateRef{nullptr, 3, nullptr}, nullptr, C, N, R, a, nullptr);
  } else if (mode == 4) {   // rollout of nstates episodes
    tpc_env_config e{};
    e.num_features = m->cfg.dec_features;
    tpc_rollout_buffers buf{};
    rc = rollout_impl(m, &e, nullptr, nullptr, &buf, C, N, R, 1, 0, 0, 0, ar, nullptr);
  }
  if (rc != TPC_OK) return 0;
  return ar.peak + 4096;
}

// workspace of tpc_a2c_grads for B episodes and the given chunk size
size_t tpc_a2c_workspace_bytes(const tpc_model* mc, int B, int num_bins, int num_resources, int chunk_states) {
  tpc_model* m = const_cast<tpc_model*>(mc);
  if (!m) return 0;
  Arena ar;
  ar.dry = true;
  tpc_a2c_hyper hp{};
  hp.chunk_states = chunk_states;
  tpc_rollout_buffers buf{};
  int rc = a2c_grads_impl(m, nullptr, nullptr, nullptr, nullptr, &buf, B, num_bins, num_resources, &hp, nullptr, nullptr,
                          nullptr, nullptr, nullptr, ar, nullptr);
  if (rc != TPC_OK) return 0;
  const size_t mx = ar.peak;
  return mx + 4096;
}

int tpc_actor_forward(tpc_model* m, const float* shadow, const float* params, const float* state,
                      const float* dec_input, const float* ptr_mask, const float* mha_mask, int B, int num_bins,
                      int num_resources, float* logits, float* probs, int32_t* argmax, void* workspace,
                      size_t workspace_bytes, void* stream) {
  if (!m || !shadow || !params || !state || !dec_input || !ptr_mask || !mha_mask || !workspace) return TPC_ERR_ARG;
  Arena ar;
  ar.base = (char*)workspace; ar.size = workspace_bytes;
  if (tpc_workspace_bytes(m, 0, B, num_bins, num_resources) > workspace_bytes) return TPC_ERR_WORKSPACE;
  Ctx cx = make_ctx(m, 0, params, shadow, nullptr, &ar, stream);
  ActorActs a;
  const int ld = m->cfg.common_embedding ? m->cfg.dec_features : m->cfg.num_bin_features;
  return actor_forward(cx, StateRef{state, ld, nullptr}, dec_input, ptr_mask, mha_mask, B, num_bins, num_resources, a,
                       logits, probs, argmax);
}

int tpc_critic_forward(tpc_model* m, const float* shadow, const float* params, const float* state,
                       const float* mha_mask, int nstates, int num_bins, int num_resources, float* values,
                       void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !shadow || !params || !state || !mha_mask || !values || !workspace) return TPC_ERR_ARG;
  Arena ar;
  ar.base = (char*)workspace; ar.size = workspace_bytes;
  if (tpc_workspace_bytes(m, 1, nstates, num_bins, num_resources) > workspace_bytes) return TPC_ERR_WORKSPACE;
  Ctx cx = make_ctx(m, 1, params, shadow, nullptr, &ar, stream);
  CriticActs a;
  const int ld = m->cfg.common_embedding ? m->cfg.dec_features : m->cfg.num_bin_features;
  return critic_forward(cx, StateRef{state, ld, nullptr}, mha_mask, nstates, num_bins, num_resources, a, values);
}

int tpc_rollout(tpc_model* m, const tpc_env_config* e, const float* actor_shadow, const float* actor_params,
                const tpc_rollout_buffers* buf, int B, int num_bins, int num_resources, int greedy, uint64_t seed,
                int64_t episode_id0, int64_t epoch, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !e || !actor_shadow || !actor_params || !buf || !workspace) return TPC_ERR_ARG;
  if (!buf->batch || !buf->bin_net_mask || !buf->mha_mask || !buf->actions || !buf->rewards) return TPC_ERR_ARG;
  if (e->num_features != m->cfg.dec_features) return TPC_ERR_SHAPE;
  if (tpc_workspace_bytes(m, 4, B, num_bins, num_resources) > workspace_bytes) return TPC_ERR_WORKSPACE;
  Arena ar;
  ar.base = (char*)workspace; ar.size = workspace_bytes;
  return rollout_impl(m, e, actor_shadow, actor_params, buf, B, num_bins, num_resources, greedy, seed, episode_id0,
                      epoch, ar, stream);
}

unsigned long long tpc_rollout_graph_replays(const tpc_model* m) {
  return (m && m->rollout_graphs) ? static_cast<const RolloutGraphs*>(m->rollout_graphs)->replays : 0;
}

int tpc_a2c_grads(tpc_model* m, const float* actor_shadow, const float* actor_params, const float* critic_shadow,
                  const float* critic_params, const tpc_rollout_buffers* buf, int B, int num_bins, int num_resources,
                  const tpc_a2c_hyper* hp, float* actor_grads, float* critic_grads, float* values, float* advantages,
                  float* losses, void* workspace, size_t workspace_bytes, void* stream) {
  if (!m || !actor_shadow || !actor_params || !critic_shadow || !critic_params || !buf || !hp || !actor_grads ||
      !critic_grads || !values || !advantages || !losses || !workspace)
    return TPC_ERR_ARG;
  if (!buf->states || !buf->dec_inputs || !buf->ptr_masks || !buf->mha_masks || !buf->actions || !buf->rewards)
    return TPC_ERR_ARG;
  if (tpc_a2c_workspace_bytes(m, B, num_bins, num_resources, hp->chunk_states) > workspace_bytes) return TPC_ERR_WORKSPACE;
  Arena ar;
  ar.base = (char*)workspace; ar.size = workspace_bytes;
  return a2c_grads_impl(m, actor_shadow, actor_params, critic_shadow, critic_params, buf, B, num_bins, num_resources, hp,
                        actor_grads, critic_grads, values, advantages, losses, ar, stream);
}

}  // extern "C"
