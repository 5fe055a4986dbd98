// Parameter inventory of the actor / critic in the reference's canonical (Keras attribute-tracking) order,
// plus the device-side "shadow" copies the tensor-core kernels read (transposed / fused / TF32-rounded).
//   canonical order: agents/models/transformer/actor/model.py:22-39, common/encoder.py:31-63,
//   common/encoder_layer.py:9-13, common/attention.py:17-21, common/utils.py:8-11,
//   actor/decoder.py:33-60, actor/decoder_layer.py:9-16, actor/pointer_attention.py:15-17,
//   critic/model.py:31-53  (SURVEY appendix B).
#pragma once
#include <vector>
#include <string>
#include "common.cuh"
#include "handle.cuh"

namespace tpc {

constexpr int D_MODEL = 128;   // every shipped config: dim_model 128, 8 heads x 16, attention/last units 128
constexpr int N_HEADS = 8;
constexpr int HEAD_DIM = 16;

struct DenseP { size_t w = 0, b = 0; int in = 0, out = 0; };
struct LnP { size_t g = 0, b = 0; };
struct MhaP { DenseP q, k, v, o; };
// shadow offsets (in floats, inside the shadow buffer)
struct EncLayerS {
  size_t wqkv_t;   // [384][128] rows: q outputs, k outputs, v outputs  (forward B operand)
  size_t bqkv;     // [384]
  size_t wqkv_c;   // [128][384] canonical (in, fused out)             (dX B operand)
  size_t wo_t, wo_c;   // [128][128]
  size_t w1_t, w1_c;   // [dff][128], [128][dff]
  size_t w2_t, w2_c;   // [128][dff], [dff][128]
};
struct EncLayerP { MhaP mha; DenseP f1, f2; LnP ln1, ln2; EncLayerS s; };
struct DecLayerS {
  size_t wv1_t, wv1_c, wo1_t, wo1_c, wq2_t, wq2_c, wo2_t, wo2_c, w1_t, w1_c, w2_t, w2_c;
  size_t wkvp_t;  // [384][128]: mha2.wk | mha2.wv | pointer.W2 outputs (only the LAST decoder layer fuses W2)
  size_t bkvp;    // [384]
  size_t wkvp_c;  // [128][384]
};
struct DecLayerP { MhaP mha1, mha2; DenseP f1, f2; LnP ln1, ln2, ln3; DecLayerS s; };
struct EncoderP {
  DenseP emb1, emb2;
  bool common = true;
  std::vector<EncLayerP> joint, bins, res;  // enc_layers, enc_layers_bins, enc_layers_resources
};
struct TensorInfo { size_t off, n; std::string name; int rows = 0; };   // rows > 0: kernel [rows][n / rows]

struct NetLayout {
  bool is_actor = true;
  int nl = 1, dff = 128;
  EncoderP enc;
  // actor
  DenseP demb;
  std::vector<DecLayerP> dec;
  DenseP W1, W2, V;
  size_t w1p_t = 0, w1p_c = 0;  // pointer W1 shadows
  // critic
  DenseP h0, h1, hv;
  size_t h0_t = 0, h1_t = 0, h1_c = 0;  // h0_t: [128][S*128]; canonical h0 [S*128][128] is used directly for dX
  size_t h0_c = 0;
  int S_critic = 0;
  size_t total = 0;         // canonical parameter count
  size_t shadow_total = 0;  // floats in the shadow buffer
  size_t shadow_half = 0;   // forward weight W_t at offset o has its fp16 split image at o + shadow_half
  std::vector<TensorInfo> tensors;  // canonical tensors in order (Adam / clipnorm granularity)
};

// copy job of the shadow builder: dst[r*ld_dst + c (+col0)] = tf32_rn(src[...]) ; transpose selects src index order
struct ShadowJob {
  unsigned long long src, dst;   // float offsets
  int rows, cols;                // shape of the DESTINATION block
  int ld_dst;                    // leading dimension of destination
  int ld_src;                    // leading dimension of source
  int transpose;                 // 1: dst[r][c] = src[c][r] ; 0: dst[r][c] = src[r][c]
  int round;                     // 0 raw copy | 1 RN-round to TF32 | 2 fp16 split image (see below)
  // round == 2: dst is the float offset of the whole [mat_rows][ld_dst] matrix in the mirror half of the image, read
  // as fp16 [2][mat_rows][ld_dst]: plane 0 = x_h = fp16(x), plane 1 = fp16(x - x_h); this block starts at row0
  int row0 = 0, mat_rows = 0;
};

struct ModelDims {
  int nl_actor, dff_actor, nl_critic, dff_critic;
  int common_emb, Fb, Fr, Fdec;
  int num_bins_train, S_train;  // critic head width
  float clip_C; int has_clip;
};

int build_layout(const ModelDims& md, bool actor, NetLayout* out, std::vector<ShadowJob>* jobs);

}  // namespace tpc

// opaque model object of the C-ABI
struct tpc_model {
  tpc_handle* h = nullptr;
  tpc_model_config cfg;
  tpc::ModelDims dims;
  tpc::NetLayout actor, critic;
  std::vector<tpc::ShadowJob> actor_jobs, critic_jobs;
  tpc::ShadowJob* d_actor_jobs = nullptr;
  tpc::ShadowJob* d_critic_jobs = nullptr;
  // Adam chunk tables (device): chunk -> (tensor id, begin, end)
  int* d_chunks[2] = {nullptr, nullptr};
  int n_chunks[2] = {0, 0};
  int n_tensors[2] = {0, 0};
  float* d_norms[2] = {nullptr, nullptr};
  // cached CUDA graphs of the rollout's step loop (net.cu: RolloutGraphs), released through the function beside it
  void* rollout_graphs = nullptr;
  void (*rollout_graphs_free)(void*) = nullptr;
};
