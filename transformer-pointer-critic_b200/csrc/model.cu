// Parameter layouts (canonical Keras order), kernel-side weight image ("shadow") builder, Adam + clipnorm.
#include "handle.cuh"
#include "model.cuh"

namespace tpc {

namespace {

struct Builder {
  NetLayout* L;
  std::vector<ShadowJob>* jobs;
  size_t off = 0;   // canonical cursor
  size_t soff = 0;  // shadow cursor

  DenseP dense(const std::string& name, int in, int out) {
    DenseP p;
    p.in = in; p.out = out;
    p.w = off; L->tensors.push_back({off, (size_t)in * out, name + ".kernel", in}); off += (size_t)in * out;
    p.b = off; L->tensors.push_back({off, (size_t)out, name + ".bias"}); off += out;
    return p;
  }
  LnP ln(const std::string& name) {
    LnP p;
    p.g = off; L->tensors.push_back({off, (size_t)D_MODEL, name + ".gamma"}); off += D_MODEL;
    p.b = off; L->tensors.push_back({off, (size_t)D_MODEL, name + ".beta"}); off += D_MODEL;
    return p;
  }
  MhaP mha(const std::string& name) {
    MhaP m;
    m.q = dense(name + ".wq", D_MODEL, D_MODEL);
    m.k = dense(name + ".wk", D_MODEL, D_MODEL);
    m.v = dense(name + ".wv", D_MODEL, D_MODEL);
    m.o = dense(name + ".dense", D_MODEL, D_MODEL);
    return m;
  }
  size_t salloc(size_t n) {
    size_t o = soff;
    soff += (n + 255) & ~size_t(255);  // 1 KB alignment keeps every TMA base 16-byte aligned
    return o;
  }
  // transposed RAW copy of W[in][out] into dst[out][in] (ld = ld_dst) at row offset row0 of a matrix with mat_rows
  // rows: forward B operand of the split GEMM (the tensor core truncates it); the matching bf16 correction image goes to
  // the mirror half of the shadow buffer (patched to dst + shadow_half once the layout is complete: round == 2)
  void job_t(const DenseP& d, size_t dst, int ld_dst, int row0, int mat_rows = 0) {
    jobs->push_back({d.w, dst + (size_t)row0 * ld_dst, d.out, d.in, ld_dst, d.out, 1, 0});
    jobs->push_back({d.w, dst, d.out, d.in, ld_dst, d.out, 1, 2, row0, mat_rows ? mat_rows : d.out});
  }
  // rounded copy of W[in][out] into dst[in][ld_dst] at column offset col0
  void job_c(const DenseP& d, size_t dst, int ld_dst, int col0) {
    jobs->push_back({d.w, dst + col0, d.in, d.out, ld_dst, d.out, 0, 1});
  }
  void job_bias(const DenseP& d, size_t dst, int col0) {
    jobs->push_back({d.b, dst + col0, 1, d.out, d.out, d.out, 0, 0});
  }
  void pair(const DenseP& d, size_t* t, size_t* c) {
    *t = salloc((size_t)d.in * d.out);
    *c = salloc((size_t)d.in * d.out);
    job_t(d, *t, d.in, 0);
    job_c(d, *c, d.out, 0);
  }

  EncLayerP enc_layer(const std::string& name, int dff) {
    EncLayerP e;
    e.mha = mha(name + ".mha");
    e.f1 = dense(name + ".ffn.0", D_MODEL, dff);
    e.f2 = dense(name + ".ffn.1", dff, D_MODEL);
    e.ln1 = ln(name + ".layernorm1");
    e.ln2 = ln(name + ".layernorm2");
    e.s.wqkv_t = salloc(3 * D_MODEL * D_MODEL);
    e.s.bqkv = salloc(3 * D_MODEL);
    e.s.wqkv_c = salloc(3 * D_MODEL * D_MODEL);
    const DenseP* qkv[3] = {&e.mha.q, &e.mha.k, &e.mha.v};
    for (int j = 0; j < 3; ++j) {
      job_t(*qkv[j], e.s.wqkv_t, D_MODEL, j * D_MODEL, 3 * D_MODEL);
      job_c(*qkv[j], e.s.wqkv_c, 3 * D_MODEL, j * D_MODEL);
      job_bias(*qkv[j], e.s.bqkv, j * D_MODEL);
    }
    pair(e.mha.o, &e.s.wo_t, &e.s.wo_c);
    pair(e.f1, &e.s.w1_t, &e.s.w1_c);
    pair(e.f2, &e.s.w2_t, &e.s.w2_c);
    return e;
  }
};

__global__ void shadow_kernel(const float* __restrict__ params, float* __restrict__ shadow,
                              const ShadowJob* __restrict__ jobs) {
  const ShadowJob j = jobs[blockIdx.x];
  const long long n = (long long)j.rows * j.cols;
  for (long long i = (long long)blockIdx.y * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.y * blockDim.x) {
    const int r = (int)(i / j.cols), c = (int)(i % j.cols);
    float v = j.transpose ? params[j.src + (size_t)c * j.ld_src + r] : params[j.src + (size_t)r * j.ld_src + c];
    if (j.round == 2) {
      __half* img = reinterpret_cast<__half*>(shadow + j.dst);
      const __half hv = f16_hi(v);
      img[(size_t)(j.row0 + r) * j.ld_dst + c] = hv;
      img[(size_t)(j.mat_rows + j.row0 + r) * j.ld_dst + c] = f16_hi(v - __half2float(hv));
      continue;
    }
    if (j.round == 1) v = tf32_rn(v);
    shadow[j.dst + (size_t)r * j.ld_dst + c] = v;
  }
}

// ---- Adam (Keras semantics) -------------------------------------------------------------------------
constexpr int ADAM_CHUNK = 2048;

__global__ void sumsq_kernel(const float* __restrict__ g, const int* __restrict__ chunks, float* __restrict__ norms,
                             float scale) {
  const int t = chunks[blockIdx.x * 3 + 0];
  const int b = chunks[blockIdx.x * 3 + 1], e = chunks[blockIdx.x * 3 + 2];
  float s = 0.f;
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    const float v = g[i] * scale;
    s = fmaf(v, v, s);
  }
  s = warp_sum(s);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    atomicAdd(&norms[t], tot);
  }
}

// theta -= lr_t * m / (sqrt(v) + eps) with g clipped per tensor: g *= clipnorm / max(norm, clipnorm)
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, const int* __restrict__ chunks, const float* __restrict__ norms,
                            float lr_t, float beta1, float beta2, float eps, float clipnorm, float scale) {
  const int t = chunks[blockIdx.x * 3 + 0];
  const int b = chunks[blockIdx.x * 3 + 1], e = chunks[blockIdx.x * 3 + 2];
  float mult = scale;
  if (clipnorm > 0.f) {
    const float norm = sqrtf(norms[t]);
    if (norm > clipnorm) mult *= clipnorm / norm;   // tf.clip_by_norm
  }
  for (int i = b + threadIdx.x; i < e; i += blockDim.x) {
    const float gi = g[i] * mult;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] -= lr_t * mi / (sqrtf(vi) + eps);
  }
}

}  // namespace

int build_layout(const ModelDims& md, bool actor, NetLayout* L, std::vector<ShadowJob>* jobs) {
  Builder b{L, jobs};
  L->is_actor = actor;
  L->nl = actor ? md.nl_actor : md.nl_critic;
  L->dff = actor ? md.dff_actor : md.dff_critic;
  if (L->dff % 128 != 0 || L->dff <= 0 || L->nl < 1) return TPC_ERR_SHAPE;
  L->enc.common = md.common_emb != 0;
  const int fb = md.common_emb ? md.Fdec : md.Fb;
  L->enc.emb1 = b.dense("encoder.embedding1", fb, D_MODEL);
  if (!md.common_emb) L->enc.emb2 = b.dense("encoder.embedding2", md.Fr, D_MODEL);
  for (int i = 0; i < L->nl; ++i) L->enc.joint.push_back(b.enc_layer("encoder.enc_layers." + std::to_string(i), L->dff));
  for (int i = 0; i < L->nl; ++i) L->enc.bins.push_back(b.enc_layer("encoder.enc_layers_bins." + std::to_string(i), L->dff));
  for (int i = 0; i < L->nl; ++i) L->enc.res.push_back(b.enc_layer("encoder.enc_layers_resources." + std::to_string(i), L->dff));
  if (actor) {
    L->demb = b.dense("decoder.embedding", md.Fdec, D_MODEL);
    for (int i = 0; i < L->nl; ++i) {
      const std::string name = "decoder.dec_layers." + std::to_string(i);
      DecLayerP d;
      d.mha1 = b.mha(name + ".mha1");
      d.mha2 = b.mha(name + ".mha2");
      d.f1 = b.dense(name + ".ffn.0", D_MODEL, L->dff);
      d.f2 = b.dense(name + ".ffn.1", L->dff, D_MODEL);
      d.ln1 = b.ln(name + ".layernorm1");
      d.ln2 = b.ln(name + ".layernorm2");
      d.ln3 = b.ln(name + ".layernorm3");
      L->dec.push_back(d);
    }
    L->W1 = b.dense("decoder.pointer.W1", D_MODEL, D_MODEL);
    L->W2 = b.dense("decoder.pointer.W2", D_MODEL, D_MODEL);
    L->V = b.dense("decoder.pointer.V", D_MODEL, 1);
    for (int i = 0; i < L->nl; ++i) {
      DecLayerP& d = L->dec[i];
      b.pair(d.mha1.v, &d.s.wv1_t, &d.s.wv1_c);
      b.pair(d.mha1.o, &d.s.wo1_t, &d.s.wo1_c);
      b.pair(d.mha2.q, &d.s.wq2_t, &d.s.wq2_c);
      b.pair(d.mha2.o, &d.s.wo2_t, &d.s.wo2_c);
      b.pair(d.f1, &d.s.w1_t, &d.s.w1_c);
      b.pair(d.f2, &d.s.w2_t, &d.s.w2_c);
      // enc_out -> [K2 | V2 | (pointer W2 on the last layer, zeros otherwise)] : always 384 wide
      d.s.wkvp_t = b.salloc(3 * D_MODEL * D_MODEL);
      d.s.bkvp = b.salloc(3 * D_MODEL);
      d.s.wkvp_c = b.salloc(3 * D_MODEL * D_MODEL);
      const bool last = (i == L->nl - 1);
      const DenseP* kvp[3] = {&d.mha2.k, &d.mha2.v, &L->W2};
      for (int j = 0; j < (last ? 3 : 2); ++j) {
        b.job_t(*kvp[j], d.s.wkvp_t, D_MODEL, j * D_MODEL, 3 * D_MODEL);
        b.job_c(*kvp[j], d.s.wkvp_c, 3 * D_MODEL, j * D_MODEL);
        b.job_bias(*kvp[j], d.s.bkvp, j * D_MODEL);
      }
    }
    b.pair(L->W1, &L->w1p_t, &L->w1p_c);
  } else {
    L->S_critic = md.S_train;
    L->h0 = b.dense("final_layer0", md.S_train * D_MODEL, D_MODEL);
    L->h1 = b.dense("final_layer1", D_MODEL, D_MODEL);
    L->hv = b.dense("final_layer", D_MODEL, 1);
    b.pair(L->h0, &L->h0_t, &L->h0_c);
    b.pair(L->h1, &L->h1_t, &L->h1_c);
  }
  L->total = b.off;
  L->shadow_half = b.soff;
  L->shadow_total = 2 * b.soff;   // [main image | bf16 correction images of the forward weights at the same offsets]
  for (ShadowJob& j : *jobs)
    if (j.round == 2) j.dst += L->shadow_half;
  return TPC_OK;
}

}  // namespace tpc

using namespace tpc;

extern "C" {

int tpc_model_create(tpc_handle* h, const tpc_model_config* cfg, tpc_model** out) {
  if (!h || !cfg || !out) return TPC_ERR_ARG;
  if (cfg->dim_model != D_MODEL || cfg->num_heads != N_HEADS || cfg->attention_dense_units != D_MODEL ||
      cfg->last_layer_units != D_MODEL) {
    fprintf(stderr, "[tpc] unsupported model dims: dim_model/units must be 128 and num_heads 8\n");
    return TPC_ERR_SHAPE;
  }
  if (cfg->dec_features < 1 || cfg->dec_features > 4 || cfg->num_bins < 1 || cfg->num_resources < 1) return TPC_ERR_SHAPE;
  if (!cfg->common_embedding && (cfg->num_bin_features < 1 || cfg->num_bin_features > 4 ||
                                 cfg->num_resource_features < 1 || cfg->num_resource_features > 4))
    return TPC_ERR_SHAPE;
  tpc_model* m = new tpc_model();
  m->h = h;
  m->cfg = *cfg;
  ModelDims& d = m->dims;
  d.nl_actor = cfg->actor_num_layers; d.dff_actor = cfg->actor_inner_layer_dim;
  d.nl_critic = cfg->critic_num_layers; d.dff_critic = cfg->critic_inner_layer_dim;
  d.common_emb = cfg->common_embedding; d.Fb = cfg->num_bin_features; d.Fr = cfg->num_resource_features;
  d.Fdec = cfg->dec_features;
  d.num_bins_train = cfg->num_bins; d.S_train = cfg->num_bins + cfg->num_resources;
  d.clip_C = cfg->logit_clipping_C; d.has_clip = cfg->has_logit_clipping;
  int rc = build_layout(d, true, &m->actor, &m->actor_jobs);
  if (rc == TPC_OK) rc = build_layout(d, false, &m->critic, &m->critic_jobs);
  if (rc != TPC_OK) { delete m; return rc; }
  for (int which = 0; which < 2; ++which) {
    const NetLayout& L = which ? m->critic : m->actor;
    const std::vector<ShadowJob>& jobs = which ? m->critic_jobs : m->actor_jobs;
    ShadowJob** dj = which ? &m->d_critic_jobs : &m->d_actor_jobs;
    TPC_CHECK_CUDA(cudaMalloc(dj, jobs.size() * sizeof(ShadowJob)));
    TPC_CHECK_CUDA(cudaMemcpy(*dj, jobs.data(), jobs.size() * sizeof(ShadowJob), cudaMemcpyHostToDevice));
    std::vector<int> chunks;
    for (size_t t = 0; t < L.tensors.size(); ++t) {
      for (size_t b0 = 0; b0 < L.tensors[t].n; b0 += ADAM_CHUNK) {
        const size_t e0 = b0 + ADAM_CHUNK < L.tensors[t].n ? b0 + ADAM_CHUNK : L.tensors[t].n;
        chunks.push_back((int)t);
        chunks.push_back((int)(L.tensors[t].off + b0));
        chunks.push_back((int)(L.tensors[t].off + e0));
      }
    }
    m->n_chunks[which] = (int)chunks.size() / 3;
    m->n_tensors[which] = (int)L.tensors.size();
    TPC_CHECK_CUDA(cudaMalloc(&m->d_chunks[which], chunks.size() * sizeof(int)));
    TPC_CHECK_CUDA(cudaMemcpy(m->d_chunks[which], chunks.data(), chunks.size() * sizeof(int), cudaMemcpyHostToDevice));
    TPC_CHECK_CUDA(cudaMalloc(&m->d_norms[which], L.tensors.size() * sizeof(float)));
  }
  *out = m;
  return TPC_OK;
}

int tpc_model_destroy(tpc_model* m) {
  if (!m) return TPC_OK;
  if (m->rollout_graphs && m->rollout_graphs_free) m->rollout_graphs_free(m->rollout_graphs);
  cudaFree(m->d_actor_jobs);
  cudaFree(m->d_critic_jobs);
  for (int w = 0; w < 2; ++w) {
    cudaFree(m->d_chunks[w]);
    cudaFree(m->d_norms[w]);
  }
  delete m;
  return TPC_OK;
}

size_t tpc_param_count(const tpc_model* m, int which) { return m ? (which ? m->critic.total : m->actor.total) : 0; }
size_t tpc_shadow_count(const tpc_model* m, int which) {
  return m ? (which ? m->critic.shadow_total : m->actor.shadow_total) : 0;
}
int tpc_num_tensors(const tpc_model* m, int which) {
  return m ? (int)(which ? m->critic.tensors.size() : m->actor.tensors.size()) : 0;
}
int tpc_param_layout(const tpc_model* m, int which, int64_t* offsets, int64_t* sizes) {
  if (!m || !offsets || !sizes) return TPC_ERR_ARG;
  const NetLayout& L = which ? m->critic : m->actor;
  for (size_t i = 0; i < L.tensors.size(); ++i) {
    offsets[i] = (int64_t)L.tensors[i].off;
    sizes[i] = (int64_t)L.tensors[i].n;
  }
  return TPC_OK;
}

int tpc_param_info(const tpc_model* m, int which, int index, char* name, int name_capacity, int64_t* rows,
                   int64_t* cols) {
  if (!m || !name || name_capacity < 1 || !rows || !cols) return TPC_ERR_ARG;
  const NetLayout& L = which ? m->critic : m->actor;
  if (index < 0 || index >= (int)L.tensors.size()) return TPC_ERR_SHAPE;
  const TensorInfo& t = L.tensors[index];
  snprintf(name, (size_t)name_capacity, "%s", t.name.c_str());
  *rows = t.rows > 0 ? t.rows : (int64_t)t.n;
  *cols = t.rows > 0 ? (int64_t)(t.n / t.rows) : 0;
  return TPC_OK;
}

int tpc_prepare_weights(tpc_model* m, int which, const float* params, float* shadow, void* stream) {
  if (!m || !params || !shadow) return TPC_ERR_ARG;
  const std::vector<ShadowJob>& jobs = which ? m->critic_jobs : m->actor_jobs;
  const ShadowJob* dj = which ? m->d_critic_jobs : m->d_actor_jobs;
  const NetLayout& L = which ? m->critic : m->actor;
  cudaStream_t st = (cudaStream_t)stream;
  TPC_CHECK_CUDA(cudaMemsetAsync(shadow, 0, L.shadow_total * sizeof(float), st));
  dim3 grid((unsigned)jobs.size(), 16);
  shadow_kernel<<<grid, 256, 0, st>>>(params, shadow, dj);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int tpc_adam_apply(tpc_model* m, int which, float* params, const float* grads, float* adam_m, float* adam_v, int step,
                   float lr, float clipnorm, float grad_scale, void* stream) {
  if (!m || !params || !grads || !adam_m || !adam_v || step < 1) return TPC_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const float beta1 = 0.9f, beta2 = 0.999f, eps = 1e-7f;
  const double lr_t = (double)lr * sqrt(1.0 - pow((double)beta2, step)) / (1.0 - pow((double)beta1, step));
  if (clipnorm > 0.f) {
    TPC_CHECK_CUDA(cudaMemsetAsync(m->d_norms[which], 0, m->n_tensors[which] * sizeof(float), st));
    sumsq_kernel<<<m->n_chunks[which], 256, 0, st>>>(grads, m->d_chunks[which], m->d_norms[which], grad_scale);
    TPC_CHECK_LAUNCH();
  }
  adam_kernel<<<m->n_chunks[which], 256, 0, st>>>(params, grads, adam_m, adam_v, m->d_chunks[which], m->d_norms[which],
                                                  (float)lr_t, beta1, beta2, eps, clipnorm, grad_scale);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // extern "C"
