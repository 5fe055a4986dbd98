// tcgen05 (UMMA) TF32 GEMM kernels for sm_100a: TMA-fed, TMEM accumulators, warp-specialised, persistent.
//
//   gemm_tc_kernel  : D[M,N]  = epi(A[M,K] @ Bt[N,K]^T)       both operands K-major   (forward + dX)
//   wgrad_tc_kernel : dW[Kin,Nout] += X[M,Kin]^T @ dY[M,Nout]  both operands MN-major  (weight gradients)
//
// These carry every Dense layer of the reference model:
//   agents/models/transformer/common/attention.py:33-35,50 (wq/wk/wv/dense),
//   common/utils.py:8-11 (FFN), actor/decoder_layer.py:22-30, actor/pointer_attention.py:40-41 (W1/W2),
//   critic/model.py:65-83 (final_layer0/1) and their backward passes (agents/trainer.py:115-141).
//
// Two numeric modes:
//   single pass (backward dX GEMMs, wgrad): tcgen05.mma kind::tf32 on the raw fp32 tiles (SS mode, M=128 N=128 K=8)
//   split (every forward GEMM; fp32-accurate, ~2^-21 relative): x = x_h + x_r with x_h = fp16(x), x_r = fp16(x - x_h)
//     and D = A_h.B_h + A_h.B_r + A_r.B_h as three kind::f16 products (K=16) with fp32 accumulation in TMEM. The weight
//     halves come from a prebuilt fp16 image (build_split_image / the shadow builder), the activation halves are made on
//     chip by four splitter warps that park them in TMEM (TS-mode MMAs). This replaced 3xTF32 (three kind::tf32 products
//     on fp32 tiles) in round 2: the B200 runs these launches at its 1000 W power cap (nvidia-smi: sw_power_cap, SM clock
//     1.3-1.5 GHz under 3xTF32 with random operands, 1.97 GHz with zeros), so tensor-pipe cycles and operand bytes are
//     paid for in clock rate; the fp16 split needs half the tensor cycles and half the weight bytes per k-block.
//
// Roles in gemm_tc_kernel (one persistent CTA of 352 threads per SM; gemm_bres_kernel: 320, no warp 10):
//   warps 0-3: splitters (split mode): activation k-block smem -> registers -> fp16 halves -> TMEM slot
//   warps 4-7: epilogue (thread == accumulator row: tcgen05.ld, bias/ReLU/residual/LayerNorm in registers,
//              swizzled smem staging, TMA store)
//   warp 8   : TMA producer of the activation slots (single pass: of the A + B stages), TMEM allocator
//   warp 9   : MMA issuer (whole warp in uniform control flow, one elected lane issues), barrier init
//   warp 10  : TMA producer of the weight slots (split mode)
#include <cuda_fp16.h>
#include "gemm_tc.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace tpc {

namespace {

constexpr int BM = 128;
constexpr int BN = 128;
constexpr int BK = 32;                        // fp32 columns per k-block == one 128-byte swizzle span
// Ring space of the generic kernel: 160 KB. The HBM latency is hidden by activation bytes in flight, so
//   split mode:  NA activation slots (16 KB; freed by the splitter warps as soon as the k-block sits in TMEM, NOT by
//                the MMAs) + NB weight slots (fp16 B_h + B_r, 16 KB, served by L2; freed by the MMA commit)
//   single-pass: 5 stages of A + B (32 KB), freed by the MMA commit (SS-mode MMAs read A from shared memory)
#ifndef GEMM_NA
#define GEMM_NA 5
#endif
#ifndef GEMM_NB
#define GEMM_NB 4
#endif
#ifndef BRES_SPLIT_ASTAGES
#define BRES_SPLIT_ASTAGES 6
#endif
#ifndef BRES_SPLIT_EPI_COLS
#define BRES_SPLIT_EPI_COLS 128
#endif
constexpr int NA = GEMM_NA;
constexpr int NB = GEMM_NB;
constexpr int NT = 8;                         // TMEM slots of a split k-block (16 + 16 columns of packed halves)
constexpr int STAGES_1P = 5;
constexpr int MAX_STAGES = 6;
constexpr int RING_BYTES = 160 * 1024;
constexpr int LN_SCRATCH_OFF = (NA + NB) * 16 * 1024;   // split mode leaves the tail of the ring free: LayerNorm row partials
static_assert(LN_SCRATCH_OFF + 2048 <= RING_BYTES && (STAGES_1P - 1) * 32 * 1024 <= LN_SCRATCH_OFF, "LayerNorm scratch must not overlap a live stage");
constexpr int TILE_BYTES = BM * BK * 4;       // 16 KB: one operand k-block
constexpr int BSLOT_BYTES = TILE_BYTES;       // split mode: fp16 B_h + B_r tiles of a k-block (2 x 8 KB)
constexpr int EPI_BYTES = BM * BN * 4;        // 64 KB staging tile (4 boxes of 128 rows x 32 cols)
constexpr int GEMM_THREADS = 320;             // warps 0-3 A_lo splitters, 4-7 epilogue, 8 TMA, 9 MMA
constexpr int BRES_THREADS = 448;             // weight-resident kernel: + warps 10-13 = second epilogue warpgroup
constexpr int GEMM_TC_THREADS = 480;          // generic kernel: + warp 10 = TMA producer of the weight slots (split mode),
                                              // warps 11-14 = second half of the LayerNorm epilogue
constexpr int WG_THREADS = 256;
constexpr int GEMM_SMEM = RING_BYTES + EPI_BYTES + 512 + 1024;  // + barriers + alignment slack
constexpr int TMEM_COLS = 512;                // 2 x 128 accumulators + NT x 32 columns of split k-blocks
constexpr int TMEM_SPLIT = 256;               // first column of the split k-block slots

constexpr int WG_TB = 32;                     // tokens per wgrad stage (4 MMAs of K=8)
constexpr int WG_STAGES = 6;
constexpr int WG_BOX_BYTES = WG_TB * BK * 4;  // 4 KB: 32 tokens x 32 features
constexpr int WG_STAGE_BYTES = 8 * WG_BOX_BYTES;  // X: 4 boxes (128 features) + dY: 4 boxes
constexpr int WG_ONES_BYTES = 64 * 128;          // 64 rows x 128 B of 1.0f: A operand of the bias-gradient MMA
constexpr int WG_SMEM = WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES + 256 + 1024;
constexpr int WG_TMEM_COLS = 512;                // 2 x 128 (dW accumulators) + 2 x 128 (column sums)

#ifdef GEMM_PROFILE
// phase timing: every role thread accumulates its clock64 intervals in registers (a global read-modify-write per
// stamp would add its own latency to the next interval) and CTA 0 flushes them once at the end of the kernel
__device__ unsigned long long gemm_prof[16];
#define GP_DECL() long long gp_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define GP_T0() long long gp_t = clock64()
#define GP_ADD(i)                        \
  do {                                   \
    const long long gp_n = clock64();    \
    gp_acc[i] += gp_n - gp_t;            \
    gp_t = gp_n;                         \
  } while (0)
#define GP_FLUSH()                                                                                      \
  do {                                                                                                  \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) {                                                    \
      _Pragma("unroll") for (int gp_i = 0; gp_i < 16; ++gp_i)                                           \
        if (gp_acc[gp_i]) atomicAdd(&gemm_prof[gp_i], (unsigned long long)gp_acc[gp_i]);                 \
    }                                                                                                   \
  } while (0)
#else
#define GP_DECL()
#define GP_T0()
#define GP_ADD(i)
#define GP_FLUSH()
#endif

#ifdef GEMM_CLOCKS
__device__ long long gemm_clk[2];   // SM clocks and nanoseconds of CTA 0's MMA issuer (clock-rate probe)
#endif

struct GemmArgs {
  int M, N, K;
  int n_tiles, k_splits, kb_per_split, num_items;
  const float* bias;
  const float* gamma;
  const float* beta;
  float* rstd_out;
  float* atomic_out;
  int ld_out;
  int aux_mode, relu, ln, round_out, atomic;
  int aux_pf;  // LayerNorm epilogues: L2 prefetch of the next tile's residual boxes one tile ahead (TPC_AUX_PREFETCH=0: off)
  int split;   // fp32-accurate FP16 split product (see the file header): activation halves built on chip, weight halves from the image
  float ln_eps;
  uint32_t* mask_out;         // ReLU bit image (weight-resident kernel, MASK template variants)
  const uint32_t* mask_in;
  int mask_ld;
  const float* lnb_y;         // LayerNorm-backward epilogue (see GemmEpi)
  const float* lnb_rstd;
  float* lnb_dgamma;
  float* lnb_dbeta;
};

// byte offset of (row, col) inside the staging tile: 4 boxes [128 rows][32 cols], rows of 128 B, 16-byte chunks
// XOR-swizzled with (row & 7) -- the layout TMA produces/consumes for CU_TENSOR_MAP_SWIZZLE_128B.
__device__ __forceinline__ uint32_t epi_off(int row, int col4 /* col / 4 */) {
  const int box = col4 >> 3;
  const int chunk = (col4 & 7) ^ (row & 7);
  return box * TILE_BYTES + row * 128 + chunk * 16;
}


// Splitter body (split mode), thread == tile row: the row's 32 fp32 of a k-block -> 32 packed words for its TMEM slot
//   columns [TMEM_SPLIT + 32 ts, +16):      fp16(a)            } packed pairs (element 2c low half, 2c + 1 high half):
//   columns [TMEM_SPLIT + 32 ts + 16, +16): fp16(a - fp16(a))  } A operands of the kind::f16 MMAs
__device__ __forceinline__ void split_row_load(const uint8_t* sa, int row, float* out) {
  uint32_t* pk = reinterpret_cast<uint32_t*>(out);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 a = *reinterpret_cast<const float4*>(sa + ((c ^ (row & 7)) << 4));
    f16_split2(a.x, a.y, pk[c * 2 + 0], pk[16 + c * 2 + 0]);
    f16_split2(a.z, a.w, pk[c * 2 + 1], pk[16 + c * 2 + 1]);
  }
}
// The 6 MMAs (kind::f16, K = 16) of a split k-block: D (+)= A_h.B_h + A_h.B_r + A_r.B_h with x = x_h + x_r the two-term
// FP16 split (f16_split2): fp32-accurate at half the tensor-pipe time of three TF32 products, and on 16-bit weight
// tiles. sb = shared address of the weight k-block: fp16(B) then fp16(B_r), 8 KB each (128 rows x 64 B, 64B swizzle).
__device__ __forceinline__ void issue_split_kblock(uint32_t tmem_d, uint32_t tmem_base, int ts, uint32_t sb, bool first) {
  constexpr uint32_t idesc16 = umma_idesc_f16(BM, BN, 0, 0);
  const uint64_t dbh = umma_smem_desc(sb, 16, 512, 4);
  const uint64_t dbr = umma_smem_desc(sb + TILE_BYTES / 2, 16, 512, 4);
  const uint32_t tah = tmem_base + TMEM_SPLIT + ts * 32;
  const uint32_t tar = tah + 16;
  // advance along K inside the 64-byte swizzle span: 16 halves = 32 B -> +2 in the (addr >> 4) field
#pragma unroll
  for (int j = 0; j < BK / 16; ++j) {
    umma_f16_ts(tmem_d, tah + 8 * j, dbh + 2 * j, idesc16, (!first || j > 0) ? 1u : 0u);   // A_h . B_h
    umma_f16_ts(tmem_d, tah + 8 * j, dbr + 2 * j, idesc16, 1u);                            // A_h . B_r
    umma_f16_ts(tmem_d, tar + 8 * j, dbh + 2 * j, idesc16, 1u);                            // A_r . B_h
  }
}

// LNB: the LayerNorm-BACKWARD epilogue (GemmEpi::lnb_y) as its own instantiation, so that the common one keeps its
// registers (every epilogue variant of this kernel sits at the 128-register cap of 480 threads).
template <bool LNB>
__global__ void __launch_bounds__(GEMM_TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmAux,
               const __grid_constant__ CUtensorMap tmD, const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* epi = smem + RING_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi + EPI_BYTES);
  uint64_t* full = bars;                 // [MAX_STAGES] stage landed (split mode: activation slot landed)
  uint64_t* empty = bars + MAX_STAGES;   // [MAX_STAGES] stage free (split mode: activation slot read by the 4 splitter warps)
  uint64_t* tfull = bars + 2 * MAX_STAGES;  // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint64_t* auxfull = tempty + 2;        // [1]
  uint64_t* lofull = auxfull + 1;        // [NT] split k-block is in its TMEM slot
  uint64_t* auxbox = lofull + NT;        // [4] streaming epilogue: aux chunk of a staging box landed
  uint64_t* tfree = auxbox + 4;          // [NT] TMEM slot consumed by the MMAs
  uint64_t* bfull = tfree + NT;          // [NB] weight slot landed
  uint64_t* bempty = bfull + NB;         // [NB] weight slot consumed by the MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty + NB);
  // single-pass ring (one stage fewer under a LayerNorm epilogue: its row partials live in the tail of the ring)
  const int nstages = g.ln ? STAGES_1P - 1 : STAGES_1P;
  constexpr int stage_bytes = 2 * TILE_BYTES;
  uint8_t* bring = ring + NA * TILE_BYTES;      // split mode: weight slots behind the activation slots

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  GP_DECL();

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (g.split) tma_prefetch_desc(&tmBlo);
    if (g.aux_mode) tma_prefetch_desc(&tmAux);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], g.split ? 4 : 1);
    }
    for (int s = 0; s < NT; ++s) {
      mbar_init(&lofull[s], 4);
      mbar_init(&tfree[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&bfull[s], 1);
      mbar_init(&bempty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], (g.ln || LNB) ? 8 : 4);   // one arrival per epilogue warp (LayerNorm forward / backward: two warpgroups)
    }
    mbar_init(auxfull, 1);
    for (int a = 0; a < 4; ++a) mbar_init(&auxbox[a], 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x;
  const int step = gridDim.x;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer (split mode: activation slots only)
    if (lane == 0) {
      const int ns = g.split ? NA : nstages;
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first; item < g.num_items; item += step) {
        const int n_t = item % g.n_tiles;
        const int rest = item / g.n_tiles;
        const int ks = rest % g.k_splits;
        const int m_t = rest / g.k_splits;
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.K / BK);
        for (int kb = kb0; kb < kb1; ++kb) {
          GP_T0();
          mbar_wait(&empty[stage], phase ^ 1);
          GP_ADD(0);
          if (g.split) {
            mbar_expect_tx(&full[stage], TILE_BYTES);
            tma_load_2d(ring + stage * TILE_BYTES, &tmA, &full[stage], kb * BK, m_t * BM);
          } else {
            mbar_expect_tx(&full[stage], 2 * TILE_BYTES);
            uint8_t* sa = ring + stage * stage_bytes;
            tma_load_2d(sa, &tmA, &full[stage], kb * BK, m_t * BM);
            tma_load_2d(sa + TILE_BYTES, &tmB, &full[stage], kb * BK, n_t * BN);
          }
          if (++stage == ns) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 10) {
    // ------------------------------------------------------------------ TMA producer of the weight slots (split mode)
    // B and B_lo k-blocks come from L2 (every CTA reads the same weight image); their slots are held until the MMAs
    // that read them retire, so they cycle on their own barriers and never hold an activation slot back.
    if (lane == 0 && g.split) {
      int bs = 0;
      uint32_t bphase = 0;
      for (int item = first; item < g.num_items; item += step) {
        const int n_t = item % g.n_tiles;
        const int ks = (item / g.n_tiles) % g.k_splits;
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.K / BK);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bempty[bs], bphase ^ 1);
          mbar_expect_tx(&bfull[bs], BSLOT_BYTES);
          uint8_t* sb = bring + bs * BSLOT_BYTES;
          tma_load_2d(sb, &tmBlo, &bfull[bs], kb * BK, n_t * BN);                          // fp16(B)
          tma_load_2d(sb + BSLOT_BYTES / 2, &tmBlo, &bfull[bs], kb * BK, g.N + n_t * BN);  // fp16(B_r)
          if (++bs == NB) { bs = 0; bphase ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    // ------------------------------------------------------------------ splitters (split mode only)
    // thread == tile row: read its 32 fp32 of the k-block from the swizzled smem slot, hand the slot straight back to
    // the producer (the k-block now lives in registers), then park its two fp16 halves in a TMEM slot as the A operands
    // of the TS-mode MMAs.
    if (g.split) {
      const int row = threadIdx.x;  // 0..127 == TMEM lane
      int stage = 0, ts = 0;
      uint32_t phase = 0, tphase = 0;
      for (int item = first; item < g.num_items; item += step) {
        const int ks = (item / g.n_tiles) % g.k_splits;
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.K / BK);
        for (int kb = kb0; kb < kb1; ++kb) {
          GP_T0();
          mbar_wait(&full[stage], phase);
          if (threadIdx.x == 0) GP_ADD(1);
          float sp[32];
          split_row_load(ring + stage * TILE_BYTES + row * 128, row, sp);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);   // slot read by this warp's 32 rows
          if (threadIdx.x == 0) GP_ADD(2);
          mbar_wait(&tfree[ts], tphase ^ 1);
          tc_fence_after();
          tmem_st_32x32(tmem_base + (uint32_t(warp * 32) << 16) + TMEM_SPLIT + ts * 32, sp);
          if (threadIdx.x == 0) GP_ADD(3);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&lofull[ts]);
          if (threadIdx.x == 0) GP_ADD(4);
          if (++stage == NA) { stage = 0; phase ^= 1; }
          if (++ts == NT) { ts = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 0);
#ifdef GEMM_CLOCKS
      const long long gc0 = clock64();
      unsigned long long gt0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
      int stage = 0, ts = 0, bs = 0;
      uint32_t phase = 0, tphase = 0, bphase = 0;
      int it = 0;
      for (int item = first; item < g.num_items; item += step, ++it) {
        const int rest = item / g.n_tiles;
        const int ks = rest % g.k_splits;
        const int kb0 = ks * g.kb_per_split;
        const int kb1 = min(kb0 + g.kb_per_split, g.K / BK);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        {
          GP_T0();
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          GP_ADD(8);
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          GP_T0();
          if (g.split) {
            // fp32-accurate split product: all MMAs take A from TMEM (parked there by the splitter warps), see
            // issue_split_kblock
            mbar_wait(&bfull[bs], bphase);
            GP_ADD(5);
            mbar_wait(&lofull[ts], tphase);
            GP_ADD(6);
            tc_fence_after();
            if (elect_one()) {
              issue_split_kblock(tmem_d, tmem_base, ts, smem_u32(bring + bs * BSLOT_BYTES), kb == kb0);
              umma_commit(&bempty[bs]);  // weight slot and TMEM slot are free when these MMAs are done
              umma_commit(&tfree[ts]);
            }
            __syncwarp();
            if (++bs == NB) { bs = 0; bphase ^= 1; }
            if (++ts == NT) { ts = 0; tphase ^= 1; }
          } else {
            mbar_wait(&full[stage], phase);
            GP_ADD(5);
            tc_fence_after();
            const uint32_t sa = smem_u32(ring + stage * stage_bytes);
            const uint64_t da = umma_smem_desc_sw128(sa, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(sa + TILE_BYTES, 16, 1024);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 8; ++k)
                umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              umma_commit(&empty[stage]);
            }
            __syncwarp();
            if (++stage == nstages) { stage = 0; phase ^= 1; }
          }
          GP_ADD(7);
        }
        if (elect_one()) umma_commit(&tfull[acc]);      // accumulator complete
        __syncwarp();
      }
#ifdef GEMM_CLOCKS
      if (blockIdx.x == 0 && lane == 0) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        gemm_clk[0] = clock64() - gc0;
        gemm_clk[1] = (long long)(gt1 - gt0);
      }
#endif
    }
  } else if (!LNB && g.ln && ((warp >= 4 && warp < 8) || warp >= 11)) {
    // ------------------------------------------------------------------ LayerNorm epilogue (256 threads)
    // Two warpgroups share a tile: thread (row, half) owns 64 columns of accumulator row `et` (TMEM lane quarter
    // warp % 4 either way), so a row lives in 64 registers instead of 128 -- the one-thread-per-row version could not
    // keep its loads in flight (168 registers, spills) and its 9.6 k clocks per tile bound the K = 128 launches.
    // Row sums / sums of squared deviations are exchanged through two floats per row in the free tail of the ring.
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const int half = warp >= 11 ? 1 : 0;
    const int et = q * 32 + lane;              // accumulator row inside the tile == TMEM lane
    const bool leader = (warp == 4 && lane == 0);
    float* red_sum = reinterpret_cast<float*>(ring + LN_SCRATCH_OFF);   // [2][128]
    float* red_ssq = red_sum + 256;                                      // [2][128]
    uint32_t aux_phase = 0;
    int it = 0;
    if (g.aux_mode && leader && first < g.num_items) {
      const int n_t = first % g.n_tiles;
      const int m_t = (first / g.n_tiles) / g.k_splits;
      mbar_expect_tx(auxfull, EPI_BYTES);
      for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, n_t * BN + j * BK, m_t * BM);
    }
    for (int item = first; item < g.num_items; item += step, ++it) {
      const int n_t = item % g.n_tiles;
      const int m_t = (item / g.n_tiles) / g.k_splits;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_t * BM + et;
      // The staging tile serves as residual input AND as output, so the residual of the next tile can only be requested
      // once this tile's store has read it out: that load's latency is exposed (and binds the K = 128 launches, whose
      // main loop is shorter than the epilogue). An L2 prefetch one tile ahead takes the HBM part out of it.
      if (g.aux_mode && g.aux_pf && leader && item + step < g.num_items) {
        const int nn = (item + step) % g.n_tiles, nm = ((item + step) / g.n_tiles) / g.k_splits;
        for (int j = 0; j < 4; ++j) tma_prefetch_2d(&tmAux, nn * BN + j * BK, nm * BM);
      }
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (g.aux_mode) {
        mbar_wait(auxfull, aux_phase);
        aux_phase ^= 1;
      }
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * 64;
      float v[64];
      tmem_ld_32x32(taddr, v);
      tmem_ld_32x32(taddr + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);  // accumulator drained: MMA may start the next-but-one tile
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g.aux_mode == 1) a = *reinterpret_cast<const float4*>(epi + epi_off(et, half * 16 + c4));
        const float4 b = g.bias ? __ldg(reinterpret_cast<const float4*>(g.bias) + half * 16 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        v[c4 * 4 + 0] += a.x + b.x;
        v[c4 * 4 + 1] += a.y + b.y;
        v[c4 * 4 + 2] += a.z + b.z;
        v[c4 * 4 + 3] += a.w + b.w;
        s0 += v[c4 * 4 + 0] + v[c4 * 4 + 1];
        s1 += v[c4 * 4 + 2] + v[c4 * 4 + 3];
      }
      const float my_sum = s0 + s1;
      red_sum[half * 128 + et] = my_sum;
      named_bar_sync(2, 256);
      // both halves add the two partials in the same order: identical mean in both threads of a row
      const float mean = (red_sum[et] + red_sum[128 + et]) * (1.0f / 128.0f);
      float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j += 4) {
        const float d0 = v[j] - mean, d1 = v[j + 1] - mean, d2 = v[j + 2] - mean, d3 = v[j + 3] - mean;
        q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
      }
      red_ssq[half * 128 + et] = (q0 + q1) + (q2 + q3);
      named_bar_sync(2, 256);
      const float rstd = 1.0f / sqrtf((red_ssq[et] + red_ssq[128 + et]) * (1.0f / 128.0f) + g.ln_eps);
      if (half == 0 && g.rstd_out && row < g.M) g.rstd_out[row] = rstd;
#pragma unroll
      for (int c4 = 0; c4 < 16; ++c4) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(g.gamma) + half * 16 + c4);
        const float4 bt = __ldg(reinterpret_cast<const float4*>(g.beta) + half * 16 + c4);
        float4 o;
        o.x = (v[c4 * 4 + 0] - mean) * rstd * gm.x + bt.x;
        o.y = (v[c4 * 4 + 1] - mean) * rstd * gm.y + bt.y;
        o.z = (v[c4 * 4 + 2] - mean) * rstd * gm.z + bt.z;
        o.w = (v[c4 * 4 + 3] - mean) * rstd * gm.w + bt.w;
        if (g.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
        *reinterpret_cast<float4*>(epi + epi_off(et, half * 16 + c4)) = o;
      }
      // staging tile complete -> TMA store (rows beyond M are clipped by the tensor map)
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (leader) {
        for (int j = 0; j < 4; ++j) tma_store_2d(&tmD, epi + j * TILE_BYTES, n_t * BN + j * BK, m_t * BM);
        tma_store_commit();
        tma_store_wait_read0();  // staging tile may be overwritten
        const int next = item + step;
        if (g.aux_mode && next < g.num_items) {
          const int nn = next % g.n_tiles;
          const int nm = (next / g.n_tiles) / g.k_splits;
          mbar_expect_tx(auxfull, EPI_BYTES);
          for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, nn * BN + j * BK, nm * BM);
        }
      }
      named_bar_sync(1, 256);
    }
    if (leader) tma_store_wait_all0();
  } else if (LNB && ((warp >= 4 && warp < 8) || warp >= 11)) {
    // ------------------------------------------------------------------ LayerNorm-BACKWARD epilogue (256 threads)
    // Phase 1 (thread = accumulator row x 64-column half, as TMEM dictates): dy = accumulator + aux, written back over
    // the aux tile in the staging buffer. Phase 2 (warp = row, lane = 4 columns, the layout of the standalone ln_bwd
    // kernel): row means by warp shuffles, dpre written over dy, column sums for dgamma / dbeta kept in registers
    // across all tiles of the CTA and reduced once at the end. The rows of y for the NEXT tile are requested before
    // the accumulator wait, so their latency hides behind the main loop.
    const int q = warp & 3;
    const int half = warp >= 11 ? 1 : 0;
    const int et = q * 32 + lane;
    const int ew = half * 4 + q;                // epilogue warp 0..7: rows ew * 16 .. + 15 of the tile in phase 2
    const bool leader = (warp == 4 && lane == 0);
    const float4 gm = __ldg(reinterpret_cast<const float4*>(g.gamma) + lane);
    const float4 bt = __ldg(reinterpret_cast<const float4*>(g.beta) + lane);
    float4 ig;
    ig.x = gm.x != 0.f ? 1.f / gm.x : 0.f; ig.y = gm.y != 0.f ? 1.f / gm.y : 0.f;
    ig.z = gm.z != 0.f ? 1.f / gm.z : 0.f; ig.w = gm.w != 0.f ? 1.f / gm.w : 0.f;
    float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = make_float4(0.f, 0.f, 0.f, 0.f);
    uint32_t aux_phase = 0;
    int it = 0;
    if (g.aux_mode && leader && first < g.num_items) {
      const int m_t = first / g.n_tiles;
      mbar_expect_tx(auxfull, EPI_BYTES);
      for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, j * BK, m_t * BM);
    }
    for (int item = first; item < g.num_items; item += step, ++it) {
      const int m_t = item / g.n_tiles;         // n_tiles == 1, k_splits == 1 in this mode
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      // the saved LayerNorm output rows of this warp's 16 phase-2 rows (and their 1/sigma), in flight during the wait
      float4 yv[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int r = m_t * BM + ew * 16 + i;
        yv[i] = r < g.M ? __ldg(reinterpret_cast<const float4*>(g.lnb_y + (size_t)r * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (g.aux_mode && g.aux_pf && leader && item + step < g.num_items) {   // see the LayerNorm forward epilogue
        const int nm = (item + step) / g.n_tiles;
        for (int j = 0; j < 4; ++j) tma_prefetch_2d(&tmAux, j * BK, nm * BM);
      }
      const int rl = m_t * BM + ew * 16 + (lane & 15);          // lane i (and i + 16) keeps 1/sigma of row i
      const float rs_lane = rl < g.M ? __ldg(g.lnb_rstd + rl) : 0.f;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (g.aux_mode) {
        mbar_wait(auxfull, aux_phase);
        aux_phase ^= 1;
      }
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN + half * 64;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        if (c == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[acc]);
        }
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          float4* slot = reinterpret_cast<float4*>(epi + epi_off(et, half * 16 + c * 8 + j4));
          float4 o = make_float4(v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
          if (g.aux_mode == 1) { const float4 a = *slot; o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
          *slot = o;
        }
      }
      named_bar_sync(2, 256);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int rt = ew * 16 + i;
        float4* slot = reinterpret_cast<float4*>(epi + epi_off(rt, lane));
        const float4 d = *slot;
        const float4 o = yv[i];
        float4 xh, gg;
        xh.x = (o.x - bt.x) * ig.x; xh.y = (o.y - bt.y) * ig.y; xh.z = (o.z - bt.z) * ig.z; xh.w = (o.w - bt.w) * ig.w;
        gg.x = d.x * gm.x; gg.y = d.y * gm.y; gg.z = d.z * gm.z; gg.w = d.w * gm.w;
        const float mg = warp_sum((gg.x + gg.y) + (gg.z + gg.w)) * (1.f / 128.f);
        const float mgx = warp_sum((gg.x * xh.x + gg.y * xh.y) + (gg.z * xh.z + gg.w * xh.w)) * (1.f / 128.f);
        const float rsi = __shfl_sync(0xffffffffu, rs_lane, i);
        float4 dp;
        dp.x = rsi * (gg.x - mg - xh.x * mgx); dp.y = rsi * (gg.y - mg - xh.y * mgx);
        dp.z = rsi * (gg.z - mg - xh.z * mgx); dp.w = rsi * (gg.w - mg - xh.w * mgx);
        if (g.round_out) { dp.x = tf32_rn(dp.x); dp.y = tf32_rn(dp.y); dp.z = tf32_rn(dp.z); dp.w = tf32_rn(dp.w); }
        *slot = dp;
        ag.x = fmaf(d.x, xh.x, ag.x); ag.y = fmaf(d.y, xh.y, ag.y); ag.z = fmaf(d.z, xh.z, ag.z); ag.w = fmaf(d.w, xh.w, ag.w);
        ab.x += d.x; ab.y += d.y; ab.z += d.z; ab.w += d.w;
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 256);
      if (leader) {
        for (int j = 0; j < 4; ++j) tma_store_2d(&tmD, epi + j * TILE_BYTES, j * BK, m_t * BM);
        tma_store_commit();
        tma_store_wait_read0();  // staging tile may be overwritten
        const int next = item + step;
        if (g.aux_mode && next < g.num_items) {
          const int nm = next / g.n_tiles;
          mbar_expect_tx(auxfull, EPI_BYTES);
          for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, j * BK, nm * BM);
        }
      }
      named_bar_sync(1, 256);
    }
    if (leader) tma_store_wait_all0();
    // column sums: the 8 epilogue warps meet in the (now free) staging tile, one warp adds them to global memory
    named_bar_sync(1, 256);
    float4* sg = reinterpret_cast<float4*>(epi);
    float4* sb = sg + 8 * 32;
    sg[ew * 32 + lane] = ag;
    sb[ew * 32 + lane] = ab;
    named_bar_sync(1, 256);
    if (ew == 0) {
      for (int w = 1; w < 8; ++w) {
        const float4 x = sg[w * 32 + lane], y = sb[w * 32 + lane];
        ag.x += x.x; ag.y += x.y; ag.z += x.z; ag.w += x.w;
        ab.x += y.x; ab.y += y.y; ab.z += y.z; ab.w += y.w;
      }
      atomicAdd(g.lnb_dgamma + lane * 4 + 0, ag.x); atomicAdd(g.lnb_dgamma + lane * 4 + 1, ag.y);
      atomicAdd(g.lnb_dgamma + lane * 4 + 2, ag.z); atomicAdd(g.lnb_dgamma + lane * 4 + 3, ag.w);
      atomicAdd(g.lnb_dbeta + lane * 4 + 0, ab.x); atomicAdd(g.lnb_dbeta + lane * 4 + 1, ab.y);
      atomicAdd(g.lnb_dbeta + lane * 4 + 2, ab.z); atomicAdd(g.lnb_dbeta + lane * 4 + 3, ab.w);
    }
  } else if (!LNB && warp >= 4 && warp < 8) {
    // ------------------------------------------------------------------ epilogue (128 threads)
    const int et = threadIdx.x - 128;          // 0..127 == accumulator row inside the tile == TMEM lane
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    const bool leader = (et == 0);
    uint32_t aux_phase = 0;
    int it = 0;
    if (!g.ln && !g.atomic) {
      // Streaming epilogue (bias / ReLU / aux add / aux gate / rounding): each 32-column chunk has its own 16 KB
      // staging box; its aux chunk is prefetched by TMA three chunks ahead, the result is stored right away.
      const int my_items = first < g.num_items ? (g.num_items - first + step - 1) / step : 0;
      const int nchunks = my_items * 4;
      auto issue_aux = [&](int ci) {
        const int item = first + (ci >> 2) * step;
        const int n_t = item % g.n_tiles, m_t = (item / g.n_tiles) / g.k_splits, c = ci & 3;
        mbar_expect_tx(&auxbox[c], TILE_BYTES);
        tma_load_2d(epi + c * TILE_BYTES, &tmAux, &auxbox[c], n_t * BN + c * BK, m_t * BM);
      };
      if (g.aux_mode && leader)
        for (int ci = 0; ci < 3 && ci < nchunks; ++ci) issue_aux(ci);
      int cidx = 0;
      for (int item = first; item < g.num_items; item += step, ++it) {
        const int n_t = item % g.n_tiles;
        const int m_t = (item / g.n_tiles) / g.k_splits;
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = 0; c < 4; ++c, ++cidx) {
          uint8_t* box = epi + c * TILE_BYTES;
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          if (c == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          if (g.aux_mode) mbar_wait(&auxbox[c], it & 1);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4* slot = reinterpret_cast<float4*>(box + et * 128 + ((j4 ^ (et & 7)) << 4));
            float4 o = make_float4(v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (g.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n_t * BN) + c * 8 + j4);
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (g.aux_mode) {
              const float4 a = *slot;
              if (g.aux_mode == 1) { o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
              else { o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f; }
            }
            if (g.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
            *slot = o;
          }
          // without aux the box about to be rewritten three chunks later only needs its old store to be done
          if (!g.aux_mode && leader) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (leader) {
            tma_store_2d(&tmD, box, n_t * BN + c * BK, m_t * BM);
            tma_store_commit();
            if (g.aux_mode) {
              asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
              if (cidx + 3 < nchunks) issue_aux(cidx + 3);
            }
          }
        }
      }
      if (leader) tma_store_wait_all0();
    } else {
    if (g.aux_mode && leader && first < g.num_items) {
      const int n_t = first % g.n_tiles;
      const int m_t = (first / g.n_tiles) / g.k_splits;
      mbar_expect_tx(auxfull, EPI_BYTES);
      for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, n_t * BN + j * BK, m_t * BM);
    }
    for (int item = first; item < g.num_items; item += step, ++it) {
      const int n_t = item % g.n_tiles;
      const int m_t = (item / g.n_tiles) / g.k_splits;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row = m_t * BM + et;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      if (g.aux_mode) {
        mbar_wait(auxfull, aux_phase);
        aux_phase ^= 1;
      }
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;

      if (g.atomic) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          if (row < g.M) {
            float* dst = g.atomic_out + (size_t)row * g.ld_out + n_t * BN + c * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        continue;
      }

      {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const int c4 = c * 8 + j4;
            float4 o = make_float4(v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (g.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n_t * BN) + c4);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (g.aux_mode) {
              const float4 a = *reinterpret_cast<const float4*>(epi + epi_off(et, c4));
              if (g.aux_mode == 1) { o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
              else { o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f; }
            }
            if (g.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
            *reinterpret_cast<float4*>(epi + epi_off(et, c4)) = o;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
      }

      // staging tile complete -> TMA store (rows beyond M are clipped by the tensor map)
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (leader) {
        for (int j = 0; j < 4; ++j) tma_store_2d(&tmD, epi + j * TILE_BYTES, n_t * BN + j * BK, m_t * BM);
        tma_store_commit();
        tma_store_wait_read0();  // staging tile may be overwritten
        const int next = item + step;
        if (g.aux_mode && next < g.num_items) {
          const int nn = next % g.n_tiles;
          const int nm = (next / g.n_tiles) / g.k_splits;
          mbar_expect_tx(auxfull, EPI_BYTES);
          for (int j = 0; j < 4; ++j) tma_load_2d(epi + j * TILE_BYTES, &tmAux, auxfull, nn * BN + j * BK, nm * BM);
        }
      }
      named_bar_sync(1, 128);
    }
    if (leader) tma_store_wait_all0();
    }
  }

  GP_FLUSH();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// Weight-resident variant for K == 128 (every 128-wide projection): a CTA owns ONE 128-column tile of the output,
// loads that slice of the weights (64 KB: fp32 tiles, or the fp16 B_h + B_r tiles in split mode) into shared memory once and then only
// streams activation tiles. The generic kernel re-reads the weight tile from L2 for every 128-row tile, which at
// these shapes is 1/3 to 1/2 of all L2->SM traffic (ncu: L2-bound at ~60 % of the HBM rate). CTAs that own
// different column tiles of the same rows run side by side, so the activation tile comes from HBM once.
// Epilogue: bias / ReLU / aux add / aux ReLU-gate / TF32 rounding (no LayerNorm, no split-K: generic kernel).
// ---------------------------------------------------------------------------------------------------------
template <bool SPLIT, int MASK = 0>
struct Bres {
  static constexpr int NKB = 4;                                  // K = 128 = 4 k-blocks of 32
  static constexpr int B_KB_BYTES = TILE_BYTES;                  // fp32 k-block, or its fp16 B_h + B_r tiles (2 x 8 KB)
  static constexpr int B_BYTES = NKB * B_KB_BYTES;               // 64 KB
  static constexpr int ASTAGES = SPLIT ? BRES_SPLIT_ASTAGES : (MASK == 2 ? 5 : 6);  // 16 KB activation k-blocks in flight
  static constexpr int EPI_COLS = SPLIT ? BRES_SPLIT_EPI_COLS : 128;  // staging tile: boxes of 32 columns
  static constexpr int EPI_B = BM * EPI_COLS * 4;
  static constexpr int GATE_BOX = BM * 16;                       // MASK == 2: 4 gate words per row of a tile
  static constexpr int GATE_B = MASK == 2 ? 4 * GATE_BOX : 0;    // ring of 4 tiles
  static constexpr int SMEM = B_BYTES + ASTAGES * TILE_BYTES + EPI_B + GATE_B + 512 + 1024;
};

// MASK: 0 none | 1 the relu epilogue also writes the sign of its output as a bit image | 2 the output is gated by a
// bit image (ReLU backward). Compile-time so that the plain epilogue -- a latency-critical per-chunk chain -- stays as
// it is.
template <bool SPLIT, int MASK>
__global__ void __launch_bounds__(BRES_THREADS, 1)
gemm_bres_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmAux,
                 const __grid_constant__ CUtensorMap tmD, const GemmArgs g) {
  using C = Bres<SPLIT, MASK>;
#ifdef GEMM_PROFILE
  __shared__ volatile long long gp_issue_t[8];
  __shared__ volatile long long gp_commit_t[8];
#endif
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* bsm = smem;
  uint8_t* ring = smem + C::B_BYTES;
  uint8_t* epi = ring + C::ASTAGES * TILE_BYTES;
  uint8_t* gbox = epi + C::EPI_B;               // MASK == 2: gate-word boxes (2 KB each, contiguous in global memory)
  uint64_t* bars = reinterpret_cast<uint64_t*>(gbox + C::GATE_B);
  uint64_t* full = bars;                        // [ASTAGES]
  uint64_t* empty = full + C::ASTAGES;          // [ASTAGES]
  uint64_t* lofull = empty + C::ASTAGES;        // [NT] split k-block is in its TMEM slot
  uint64_t* tfull = lofull + NT;                // [2]
  uint64_t* tempty = tfull + 2;                 // [2]
  uint64_t* auxfull = tempty + 2;               // [4] one per staging box (aux chunk landed)
  uint64_t* bfull = auxfull + 4;                // [1] weights resident
  uint64_t* tfree = bfull + 1;                  // [NT] split mode: TMEM slot of a split k-block consumed by the MMAs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfree + NT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  GP_DECL();
  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmD);
    if (SPLIT) tma_prefetch_desc(&tmBlo);
    if (g.aux_mode) tma_prefetch_desc(&tmAux);
  }
  if (warp == 9 && lane == 0) {
    for (int s = 0; s < C::ASTAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], SPLIT ? 4 : 1);   // split mode: the four splitter warps free the slot, else the MMA commit
    }
    for (int s = 0; s < NT; ++s) {
      mbar_init(&lofull[s], 4);
      mbar_init(&tfree[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], g.aux_mode == 0 ? 8 : 4);   // one arrival per epilogue warp (no aux input: two warpgroups)
    }
    for (int a = 0; a < 4; ++a) mbar_init(&auxfull[a], 1);
    mbar_init(bfull, 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // activation slots in use. Measured (1.55 M rows, split mode): with ONE column tile (every CTA streams its own rows) four
  // slots beat six by 10 % (258 vs 284 us); with three or four column tiles sharing each row tile six beat four by 4 %.
  const int nst = (SPLIT && g.n_tiles == 1) ? 4 : C::ASTAGES;
  // this CTA: column tile n_t, row tiles m0, m0 + groups, ...
  const int n_t = blockIdx.x % g.n_tiles;
  const int groups = gridDim.x / g.n_tiles;
  const int m0 = blockIdx.x / g.n_tiles;
  const int m_tiles = (g.M + BM - 1) / BM;

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(bfull, C::B_BYTES);
      for (int kb = 0; kb < C::NKB; ++kb) {
        if (SPLIT) {   // fp16 B_h and B_r tiles (see issue_split_kblock)
          tma_load_2d(bsm + kb * C::B_KB_BYTES, &tmBlo, bfull, kb * BK, n_t * BN);
          tma_load_2d(bsm + kb * C::B_KB_BYTES + TILE_BYTES / 2, &tmBlo, bfull, kb * BK, g.N + n_t * BN);
        } else {
          tma_load_2d(bsm + kb * C::B_KB_BYTES, &tmB, bfull, kb * BK, n_t * BN);
        }
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int m_t = m0; m_t < m_tiles; m_t += groups) {
        for (int kb = 0; kb < C::NKB; ++kb) {
          GP_T0();
          mbar_wait(&empty[stage], phase ^ 1);
          GP_ADD(0);
          mbar_expect_tx(&full[stage], TILE_BYTES);
#ifdef GEMM_PROFILE
          if (!SPLIT && (phase || m_t != m0)) { gp_acc[15] += clock64() - gp_commit_t[stage]; }
          gp_issue_t[stage] = clock64();
#endif
#ifdef GEMM_DBG_L2_A
          tma_load_2d(ring + stage * TILE_BYTES, &tmA, &full[stage], kb * BK, 0);
#else
          tma_load_2d(ring + stage * TILE_BYTES, &tmA, &full[stage], kb * BK, m_t * BM);
#endif
          if (++stage == nst) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 4) {
    if (SPLIT) {
      const int row = threadIdx.x;
      int stage = 0, ts = 0;
      uint32_t phase = 0, tphase = 0;
      for (int m_t = m0; m_t < m_tiles; m_t += groups) {
        for (int kb = 0; kb < C::NKB; ++kb) {
          GP_T0();
          mbar_wait(&full[stage], phase);
          if (threadIdx.x == 0) GP_ADD(1);
#ifdef GEMM_PROFILE
          if (threadIdx.x == 0) { gp_acc[13] += clock64() - gp_issue_t[stage]; gp_acc[14] += 1; }
#endif
          float sp[32];
          split_row_load(ring + stage * TILE_BYTES + row * 128, row, sp);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[stage]);   // the k-block is in registers: the slot goes back to the producer
          mbar_wait(&tfree[ts], tphase ^ 1);
          tc_fence_after();
          tmem_st_32x32(tmem_base + (uint32_t(warp * 32) << 16) + TMEM_SPLIT + ts * 32, sp);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&lofull[ts]);
          if (threadIdx.x == 0) GP_ADD(2);
          if (++stage == nst) { stage = 0; phase ^= 1; }
          if (++ts == NT) { ts = 0; tphase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    {   // whole warp in uniform control flow, one elected lane issues (see elect_one)
      constexpr uint32_t idesc = umma_idesc_tf32(BM, BN, 0, 0);
      mbar_wait(bfull, 0);
#ifdef GEMM_CLOCKS
      const long long gc0 = clock64();
      unsigned long long gt0;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
#endif
      int stage = 0, ts = 0;
      uint32_t phase = 0, tphase = 0;
      int it = 0;
      for (int m_t = m0; m_t < m_tiles; m_t += groups, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        {
          GP_T0();
          mbar_wait(&tempty[acc], acc_phase ^ 1);
          GP_ADD(8);
        }
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < C::NKB; ++kb) {
          GP_T0();
          if (SPLIT) {   // all products with A from TMEM (see issue_split_kblock); the activation slot is already free
            mbar_wait(&lofull[ts], tphase);
            GP_ADD(6);
            tc_fence_after();
            if (elect_one()) {
              issue_split_kblock(tmem_d, tmem_base, ts, smem_u32(bsm + kb * C::B_KB_BYTES), kb == 0);
              umma_commit(&tfree[ts]);
            }
            __syncwarp();
            if (++ts == NT) { ts = 0; tphase ^= 1; }
          } else {
#ifdef GEMM_PROFILE
            gp_acc[12] += clock64() - gp_issue_t[stage];   // age of this stage's TMA when the issuer gets to it
#endif
            mbar_wait(&full[stage], phase);
            GP_ADD(5);
            tc_fence_after();
            const uint64_t da = umma_smem_desc_sw128(smem_u32(ring + stage * TILE_BYTES), 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(smem_u32(bsm + kb * C::B_KB_BYTES), 16, 1024);
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < BK / 8; ++k)
                umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              umma_commit(&empty[stage]);
            }
            __syncwarp();
#ifdef GEMM_PROFILE
            gp_commit_t[stage] = clock64();
#endif
            if (++stage == nst) { stage = 0; phase ^= 1; }
          }
          GP_ADD(7);
        }
        if (elect_one()) umma_commit(&tfull[acc]);
        __syncwarp();
      }
#ifdef GEMM_CLOCKS
      if (blockIdx.x == 0) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        gemm_clk[0] = clock64() - gc0;
        gemm_clk[1] = (long long)(gt1 - gt0);
      }
#endif
    }
  } else if (g.aux_mode == 0 && ((warp >= 4 && warp < 8) || warp >= 10)) {
    // Streaming epilogue on two warpgroups: group 0 (warps 4-7) takes the 32-column chunks 0 and 1 of a tile, group 1
    // (warps 10-13) chunks 2 and 3; a thread owns accumulator row `et` (TMEM lane quarter warp % 4 in both groups).
    // Every chunk goes through a 16 KB staging box and is stored by TMA right away; each group has its own pair of
    // boxes, its own named barrier and its own store-issuing leader, so the two chains (TMEM load -> math -> staging ->
    // store), which were the critical path of this kernel on one warpgroup, run side by side.
    const int grp = warp >= 10 ? 1 : 0;
    const int q = warp & 3;
    const int et = q * 32 + lane;
    const bool leader = lane == 0 && q == 0;
    const int bar_id = 1 + grp;
    int it = 0;
    {
      int k = 0;   // chunks this group has staged
      // MASK == 2: the gate words of (this column tile, 128 rows) are 2 KB contiguous in global memory (image layout
      // [column tile][row][4 words]); the leader of group 0 brings them in two tiles ahead with a bulk copy into a ring
      // of four boxes (per-thread loads of these words cost one exposed DRAM latency per tile, whatever the prefetch
      // distance). Box (it + 2) & 3 held the words of tile it - 2: once tfull[it] has completed, every epilogue warp
      // has drained tile it - 2 (the MMAs of tile it waited for that) and read its words long before.
      const int my_tiles_g = (m_tiles - m0 + groups - 1) / groups;
      auto issue_gate = [&](int ti) {
        const int mt = m0 + ti * groups;
        mbar_expect_tx(&auxfull[ti & 3], C::GATE_BOX);
        bulk_load(gbox + (ti & 3) * C::GATE_BOX, g.mask_in + ((size_t)n_t * g.mask_ld + (size_t)mt * BM) * 4, C::GATE_BOX,
                  &auxfull[ti & 3]);
      };
      if (MASK == 2 && leader && grp == 0)
        for (int ti = 0; ti < 2 && ti < my_tiles_g; ++ti) issue_gate(ti);
      for (int m_t = m0; m_t < m_tiles; m_t += groups, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        GP_T0();
        mbar_wait(&tfull[acc], acc_phase);
        if (leader) GP_ADD(9);
        tc_fence_after();
        uint2 gate_w = make_uint2(0u, 0u), bits_out = make_uint2(0u, 0u);
        if (MASK == 2) {
          if (leader && grp == 0 && it + 2 < my_tiles_g) issue_gate(it + 2);
          mbar_wait(&auxfull[it & 3], (it >> 2) & 1);
          gate_w = *reinterpret_cast<const uint2*>(gbox + (it & 3) * C::GATE_BOX + et * 16 + grp * 8);
        }
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc, ++k) {
          const int c = grp * 2 + cc;
          uint8_t* box = epi + (grp * 2 + (k & 1)) * TILE_BYTES;
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          if (cc == 1) {                                  // this warp's share of the accumulator is drained
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          const uint32_t gw = cc == 0 ? gate_w.x : gate_w.y;
          uint32_t part[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4 o = make_float4(v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (g.bias) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n_t * BN) + c * 8 + j4);
              o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (MASK == 2) {
              o.x = (gw & (1u << (j4 * 4 + 0))) ? o.x : 0.f; o.y = (gw & (1u << (j4 * 4 + 1))) ? o.y : 0.f;
              o.z = (gw & (1u << (j4 * 4 + 2))) ? o.z : 0.f; o.w = (gw & (1u << (j4 * 4 + 3))) ? o.w : 0.f;
            }
            if (MASK == 1)   // eight independent 4-bit groups (one OR chain over 32 elements would serialise the chunk)
              part[j4] = (o.x > 0.f ? 1u << (j4 * 4 + 0) : 0u) | (o.y > 0.f ? 1u << (j4 * 4 + 1) : 0u) |
                         (o.z > 0.f ? 1u << (j4 * 4 + 2) : 0u) | (o.w > 0.f ? 1u << (j4 * 4 + 3) : 0u);
            if (g.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
            *reinterpret_cast<float4*>(box + et * 128 + ((j4 ^ (et & 7)) << 4)) = o;
          }
          if (MASK == 1) {   // one 8-byte store per row, group and tile (image rows are padded to the tile)
            const uint32_t wbits = ((part[0] | part[1]) | (part[2] | part[3])) | ((part[4] | part[5]) | (part[6] | part[7]));
            if (cc == 0) bits_out.x = wbits;
            else {
              bits_out.y = wbits;
              *reinterpret_cast<uint2*>(g.mask_out + ((size_t)n_t * g.mask_ld + m_t * BM + et) * 4 + grp * 2) = bits_out;
            }
          }
          // the box the NEXT chunk of this group writes is the one its previous store read: that store must be done
          // reading before the threads are released
          if (leader) GP_ADD(10);
          if (leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          if (leader) GP_ADD(11);
          fence_proxy_async_smem();
          named_bar_sync(bar_id, 128);
          if (leader) {
            tma_store_2d(&tmD, box, n_t * BN + c * BK, m_t * BM);
            tma_store_commit();
          }
          if (leader) GP_ADD(11);
        }
      }
      if (leader) tma_store_wait_all0();
    }
  } else if (g.aux_mode != 0 && warp >= 4 && warp < 8) {
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    const bool leader = (et == 0);
    int it = 0;
    if (!SPLIT) {
      // Streaming epilogue with a second input (residual add / ReLU gate): the aux chunk of a 32-column box is
      // prefetched by TMA three chunks ahead into the ring of 4 boxes, combined in place and stored right away.
      const int my_tiles = (m_tiles - m0 + groups - 1) / groups;
      const int nchunks = my_tiles * 4;
      auto issue_aux = [&](int ci) {
        const int m_t = m0 + (ci >> 2) * groups, c = ci & 3, b = ci & 3;
        mbar_expect_tx(&auxfull[b], TILE_BYTES);
        tma_load_2d(epi + b * TILE_BYTES, &tmAux, &auxfull[b], n_t * BN + c * BK, m_t * BM);
      };
      if (leader)
        for (int ci = 0; ci < 3 && ci < nchunks; ++ci) issue_aux(ci);
      int cidx = 0;
      for (int m_t = m0; m_t < m_tiles; m_t += groups, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tfull[acc], acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
#pragma unroll 1
        for (int c = 0; c < 4; ++c, ++cidx) {
          uint8_t* box = epi + c * TILE_BYTES;                 // box index == cidx & 3 == c
          float v[32];
          tmem_ld_32x32(taddr + c * 32, v);
          if (c == 3) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
          }
          mbar_wait(&auxfull[c], it & 1);                       // box c is filled once per tile
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            float4* slot = reinterpret_cast<float4*>(box + et * 128 + ((j4 ^ (et & 7)) << 4));
            const float4 a = *slot;
            float4 o = make_float4(v[j4 * 4 + 0], v[j4 * 4 + 1], v[j4 * 4 + 2], v[j4 * 4 + 3]);
            if (g.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(g.bias + n_t * BN) + c * 8 + j4);
              o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
            }
            if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            if (g.aux_mode == 1) { o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
            else { o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f; o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f; }
            if (g.round_out) { o.x = tf32_rn(o.x); o.y = tf32_rn(o.y); o.z = tf32_rn(o.z); o.w = tf32_rn(o.w); }
            *slot = o;
          }
          fence_proxy_async_smem();
          named_bar_sync(1, 128);
          if (leader) {
            tma_store_2d(&tmD, box, n_t * BN + c * BK, m_t * BM);
            tma_store_commit();
            // the box of chunk cidx+3 (== box of chunk cidx-1) is free once store(cidx-1) has read it
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (cidx + 3 < nchunks) issue_aux(cidx + 3);
          }
        }
      }
      if (leader) tma_store_wait_all0();
    }
  }

  GP_FLUSH();
  tc_fence_before();
  __syncthreads();
  if (warp == 8) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// Weight gradient: D[Kin-tile 128][Nout-tile 128] += sum over tokens of X[t][ki*128+i] * dY[t][nj*128+j]
// Operands are the raw row-major activation tiles (token rows of 128 B per 32-feature box) consumed as
// MN-major UMMA operands: LBO = distance between 32-feature boxes, 8 tokens (two 4-token swizzle atoms) per MMA.
// ---------------------------------------------------------------------------------------------------------
struct WgradArgs {
  int M, Kin, Nout;
  int ki_tiles, nj_tiles, splits, tok_per_split, num_items;
  float* dst[4];
  int ld[4];
  float* bias[4];   // db destination of column tile nj (or nullptr): db += colsum(dY tile), computed by ki == 0 items
};

__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmY, const WgradArgs g) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  float* ones = reinterpret_cast<float*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES + WG_ONES_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + WG_STAGES;
  uint64_t* tfull = bars + 2 * WG_STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  for (int i = threadIdx.x; i < WG_ONES_BYTES / 4; i += blockDim.x) ones[i] = 1.0f;
  fence_proxy_async_smem();   // generic-proxy writes -> visible to the UMMA operand reads

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) tmem_alloc<WG_TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int first = blockIdx.x;
  const int step = gridDim.x;

  // item -> (ki, nj, split); nj fastest so CTAs running side by side share the X token block in L2
  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = first; item < g.num_items; item += step) {
        const int nj = item % g.nj_tiles;
        const int rest = item / g.nj_tiles;
        const int ki = rest % g.ki_tiles;
        const int sp = rest / g.ki_tiles;
        const int t0 = sp * g.tok_per_split;
        const int t1 = min(t0 + g.tok_per_split, g.M);
        for (int t = t0; t < t1; t += WG_TB) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], WG_STAGE_BYTES);
          uint8_t* sx = ring + stage * WG_STAGE_BYTES;
          for (int j = 0; j < 4; ++j) tma_load_2d(sx + j * WG_BOX_BYTES, &tmX, &full[stage], ki * 128 + j * BK, t);
          for (int j = 0; j < 4; ++j)
            tma_load_2d(sx + (4 + j) * WG_BOX_BYTES, &tmY, &full[stage], nj * 128 + j * BK, t);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(128, 128, 1, 1);
      // bias gradient: D2[64][128] += Ones[64][8 tokens] (K-major) * dY tile (MN-major): every row = column sums
      constexpr uint32_t idesc_ones = umma_idesc_tf32(64, 128, 0, 1);
      const uint64_t dones = umma_smem_desc_sw128(smem_u32(ones), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int item = first; item < g.num_items; item += step, ++it) {
        const int nj_i = item % g.nj_tiles;
        const int ki_i = (item / g.nj_tiles) % g.ki_tiles;
        const bool do_bias = (ki_i == 0) && (g.bias[nj_i] != nullptr);
        const int sp = (item / g.nj_tiles) / g.ki_tiles;
        const int t0 = sp * g.tok_per_split;
        const int t1 = min(t0 + g.tok_per_split, g.M);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 128;
        for (int t = t0; t < t1; t += WG_TB) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sx = smem_u32(ring + stage * WG_STAGE_BYTES);
          // MN-major TF32 must use the 128B swizzle with 32-byte atoms (4-token K atom):
          // LBO = next 32-feature box (4 KB), SBO = next 4-token group (512 B)
          const uint64_t da = umma_smem_desc(sx, WG_BOX_BYTES, 512, 1);
          const uint64_t db = umma_smem_desc(sx + 4 * WG_BOX_BYTES, WG_BOX_BYTES, 512, 1);
#pragma unroll
          for (int k = 0; k < WG_TB / 8; ++k) {
            // next 8 tokens: +1024 B -> +64 in the (addr >> 4) field
            umma_tf32_ss(tmem_d, da + 64 * k, db + 64 * k, idesc, (t > t0 || k > 0) ? 1u : 0u);
            if (do_bias) umma_tf32_ss(tmem_d + 256, dones, db + 64 * k, idesc_ones, (t > t0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);
          if (++stage == WG_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else if (warp >= 4) {
    const int et = threadIdx.x - 128;
    const int q = warp & 3;
    int it = 0;
    for (int item = first; item < g.num_items; item += step, ++it) {
      const int nj = item % g.nj_tiles;
      const int ki = (item / g.nj_tiles) % g.ki_tiles;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * 128;
      const int row = ki * 128 + et;
      float* dst = g.dst[nj] + (size_t)row * g.ld[nj];
      const int ncols = min(128, g.Nout - nj * 128);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        if (row < g.Kin) {
          if (c * 32 + 32 <= ncols && (g.ld[nj] & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)   // 16-byte vector reduction (sm_90+): 4x fewer L2 atomic ops
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + c * 32 + j), "f"(v[j]),
                           "f"(v[j + 1]), "f"(v[j + 2]), "f"(v[j + 3])
                           : "memory");
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < ncols) atomicAdd(dst + c * 32 + j, v[j]);
          }
        }
      }
      if (ki == 0 && g.bias[nj] != nullptr && q == 0) {
        // rows of the M=64 accumulator are identical; TMEM lane 0 always holds row 0
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float v[32];
          tmem_ld_32x32(tmem_base + acc * 128 + 256 + c * 32, v);
          if (lane == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < ncols) atomicAdd(g.bias[nj] + c * 32 + j, v[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<WG_TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor maps + launchers
// ---------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_init_once;
int g_init_rc = TPC_OK;

}  // namespace

struct TmapCache {
  std::map<std::tuple<const void*, int, int, int, int, int>, CUtensorMap> maps;
  std::mutex mu;
};
TmapCache* tmap_cache_create() { return new TmapCache(); }
void tmap_cache_destroy(TmapCache* c) { delete c; }

int gemm_tc_init() {
  std::call_once(g_init_once, []() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || fn == nullptr || qres != cudaDriverEntryPointSuccess) {
      fprintf(stderr, "[tpc] cuTensorMapEncodeTiled not available (%s)\n", cudaGetErrorString(e));
      g_init_rc = TPC_ERR_DRIVER;
      return;
    }
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_bres_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bres<true>::SMEM);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_bres_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bres<true>::SMEM);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gemm_bres_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bres<false, 2>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gemm_bres_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Bres<false>::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
    if (e != cudaSuccess) {
      fprintf(stderr, "[tpc] cudaFuncSetAttribute failed: %s\n", cudaGetErrorString(e));
      g_init_rc = TPC_ERR_CUDA;
    }
  });
  return g_init_rc;
}

// 2-D fp32 row-major [rows][cols] (leading dimension ld floats), box = [box_rows][32 cols], 128B swizzle.
// atom32 selects CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (MN-major TF32 operands of the wgrad kernel).
static int get_tmap(TmapCache* tc, const float* ptr, int rows, int cols, int ld, int box_rows, int atom32,
                    CUtensorMap* out) {
  auto key = std::make_tuple((const void*)ptr, rows, cols, ld, box_rows, atom32);
  std::lock_guard<std::mutex> lk(tc->mu);
  auto it = tc->maps.find(key);
  if (it != tc->maps.end()) {
    *out = it->second;
    return TPC_OK;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 4) != 0) {
    fprintf(stderr, "[tpc] tensor map: pointer/ld not 16-byte aligned (ptr=%p ld=%d)\n", (const void*)ptr, ld);
    return TPC_ERR_ARG;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[tpc] cuTensorMapEncodeTiled failed: %d (rows=%d cols=%d ld=%d box_rows=%d)\n", (int)r, rows,
            cols, ld, box_rows);
    return TPC_ERR_DRIVER;
  }
  if (tc->maps.size() > 4096) tc->maps.clear();
  tc->maps[key] = m;
  *out = m;
  return TPC_OK;
}

// split image of a forward weight: [2 N][K] fp16 (plane 0 = B_h = fp16(B), plane 1 = B_r = fp16(B - B_h)), ld elements
// per row; boxes of 128 rows x 32 columns (64 B) in the 64-byte swizzle
static int get_tmap_corr(TmapCache* tc, const void* ptr, int rows2, int cols, int ld, CUtensorMap* out) {
  auto key = std::make_tuple(ptr, rows2, cols, ld, BN, 2);
  std::lock_guard<std::mutex> lk(tc->mu);
  auto it = tc->maps.find(key);
  if (it != tc->maps.end()) {
    *out = it->second;
    return TPC_OK;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0 || (ld % 8) != 0) {
    fprintf(stderr, "[tpc] tensor map: correction image pointer/ld not 16-byte aligned (ptr=%p ld=%d)\n", ptr, ld);
    return TPC_ERR_ARG;
  }
  CUtensorMap m;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows2};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[tpc] cuTensorMapEncodeTiled (correction image) failed: %d (rows=%d cols=%d ld=%d)\n", (int)r, rows2,
            cols, ld);
    return TPC_ERR_DRIVER;
  }
  if (tc->maps.size() > 4096) tc->maps.clear();
  tc->maps[key] = m;
  *out = m;
  return TPC_OK;
}

__global__ void split_image_kernel(const float* __restrict__ Bt, int ldb, int N, int K, __half* __restrict__ img) {
  const long long n = (long long)N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / K), c = (int)(i % K);
    const float v = Bt[(size_t)r * ldb + c];
    const __half hv = f16_hi(v);
    img[(size_t)r * ldb + c] = hv;
    img[((size_t)N + r) * ldb + c] = f16_hi(v - __half2float(hv));
  }
}

int build_split_image(const float* Bt, int ldb, int N, int K, void* image, cudaStream_t stream) {
  if (!Bt || !image || N <= 0 || K <= 0 || ldb < K) return TPC_ERR_ARG;
  const long long n = (long long)N * K;
  const int grid = (int)std::min<long long>((n + 255) / 256, 1184);
  split_image_kernel<<<grid, 256, 0, stream>>>(Bt, ldb, N, K, static_cast<__half*>(image));
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

bool gemm_weight_resident(int M, int N, int K, const GemmEpi& epi, bool split, int num_sms) {
  const int m_tiles = ceil_div(M, BM), n_tiles = N / BN;
  return (K == 128) && N % BN == 0 && !epi.ln && !epi.atomic && !(split && epi.aux_mode) && n_tiles <= num_sms &&
         m_tiles * n_tiles >= 2 * num_sms;
}

int launch_gemm(TmapCache* tc, const float* A, int lda, const float* Bt, int ldb, float* D, int ldd, int M, int N,
                int K, const GemmEpi& epi, int k_splits, int num_sms, cudaStream_t stream, const void* Bt_corr) {
  TPC_TRY(gemm_tc_init());
  if (M <= 0) return TPC_OK;
  if (N % BN != 0 || K % BK != 0 || K <= 0) return TPC_ERR_SHAPE;
  if (epi.ln && (N != BN || !epi.gamma || !epi.beta)) return TPC_ERR_SHAPE;
  if (k_splits < 1) k_splits = 1;
  if (k_splits > 1 && !epi.atomic) return TPC_ERR_ARG;
  if (epi.atomic && (epi.aux_mode || epi.ln || epi.relu || epi.bias)) return TPC_ERR_ARG;
  if (epi.lnb_y && (N != BN || epi.ln || epi.atomic || epi.relu || epi.bias || epi.aux_mode == 2 || Bt_corr || !epi.lnb_rstd ||
                    !epi.lnb_dgamma || !epi.lnb_dbeta || !epi.gamma || !epi.beta || epi.mask_in || epi.mask_out))
    return TPC_ERR_ARG;
  GemmArgs g;
  g.M = M; g.N = N; g.K = K;
  g.n_tiles = N / BN;
  const int kblocks = K / BK;
  g.kb_per_split = ceil_div(kblocks, k_splits);
  g.k_splits = ceil_div(kblocks, g.kb_per_split);
  g.num_items = ceil_div(M, BM) * g.n_tiles * g.k_splits;
  g.bias = epi.bias; g.gamma = epi.gamma; g.beta = epi.beta; g.rstd_out = epi.rstd_out;
  g.atomic_out = D; g.ld_out = ldd;
  g.aux_mode = epi.aux_mode; g.relu = epi.relu; g.ln = epi.ln; g.round_out = epi.round_out; g.atomic = epi.atomic;
  g.ln_eps = epi.ln_eps;
  g.lnb_y = epi.lnb_y; g.lnb_rstd = epi.lnb_rstd; g.lnb_dgamma = epi.lnb_dgamma; g.lnb_dbeta = epi.lnb_dbeta;
  g.split = Bt_corr ? 1 : 0;
  static const int aux_pf = [] { const char* v = getenv("TPC_AUX_PREFETCH"); return (v && v[0] == '0') ? 0 : 1; }();
  g.aux_pf = aux_pf;
  CUtensorMap tA, tB, tBlo, tAux, tD;
  TPC_TRY(get_tmap(tc, A, M, K, lda, BM, 0, &tA));
  TPC_TRY(get_tmap(tc, Bt, N, K, ldb, BN, 0, &tB));
  if (Bt_corr) { TPC_TRY(get_tmap_corr(tc, Bt_corr, 2 * N, K, ldb, &tBlo)); } else { tBlo = tB; }
  TPC_TRY(get_tmap(tc, D, M, N, ldd, BM, 0, &tD));
  if (epi.aux_mode) {
    if (!epi.aux) return TPC_ERR_ARG;
    TPC_TRY(get_tmap(tc, epi.aux, M, N, epi.ldaux, BM, 0, &tAux));
  } else {
    tAux = tD;
  }
  // weight-resident kernel: K == 128, no LayerNorm / split-K, (split mode: no aux), enough row tiles per CTA
  const bool bres = !epi.lnb_y && gemm_weight_resident(M, N, K, epi, g.split != 0, num_sms);
  g.mask_out = epi.mask_out; g.mask_in = epi.mask_in; g.mask_ld = epi.mask_ld;
  if (epi.mask_out || epi.mask_in) {
    // bit images: 16-byte aligned rows, mask_out from the split-mode relu epilogue, mask_in into a single-pass GEMM
    // image layout [N / 128][mask_ld rows][4 words], mask_ld = M rounded up to a multiple of 128
    if (!bres || epi.aux_mode || epi.mask_ld < M || epi.mask_ld % BM != 0 || (epi.mask_out && epi.mask_in)) return TPC_ERR_ARG;
    if ((epi.mask_out && !(g.split && epi.relu)) || (epi.mask_in && g.split)) return TPC_ERR_ARG;
  }
  if (bres) {
    int grid = (num_sms / g.n_tiles) * g.n_tiles;
    if (epi.mask_out) gemm_bres_kernel<true, 1><<<grid, BRES_THREADS, Bres<true>::SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
    else if (epi.mask_in) gemm_bres_kernel<false, 2><<<grid, BRES_THREADS, Bres<false, 2>::SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
    else if (g.split) gemm_bres_kernel<true, 0><<<grid, BRES_THREADS, Bres<true>::SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
    else gemm_bres_kernel<false, 0><<<grid, BRES_THREADS, Bres<false>::SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
    TPC_CHECK_LAUNCH();
    return TPC_OK;
  }
  const int grid = g.num_items < num_sms ? g.num_items : num_sms;
  if (epi.lnb_y) gemm_tc_kernel<true><<<grid, GEMM_TC_THREADS, GEMM_SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
  else gemm_tc_kernel<false><<<grid, GEMM_TC_THREADS, GEMM_SMEM, stream>>>(tA, tB, tBlo, tAux, tD, g);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

int launch_wgrad(TmapCache* tc, const float* X, int ldx, const float* dY, int ldy, int M, int Kin, int Nout,
                 const WgradDst& dst, int num_sms, cudaStream_t stream) {
  TPC_TRY(gemm_tc_init());
  if (M <= 0) return TPC_OK;
  if (Kin % BK != 0 || Nout % BK != 0) return TPC_ERR_SHAPE;
  WgradArgs g;
  g.M = M; g.Kin = Kin; g.Nout = Nout;
  g.ki_tiles = ceil_div(Kin, 128);
  g.nj_tiles = ceil_div(Nout, 128);
  if (g.nj_tiles > 4) return TPC_ERR_SHAPE;
  const int out_tiles = g.ki_tiles * g.nj_tiles;
  // one work item per SM (each item ends with 16 K atomic adds onto the same addresses: more splits only add
  // contention), at least 256 tokens (8 stages) per split
  int splits = num_sms / out_tiles;
  const int max_splits = ceil_div(M, 256);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  g.tok_per_split = ceil_div(ceil_div(M, splits), WG_TB) * WG_TB;
  g.splits = ceil_div(M, g.tok_per_split);
  g.num_items = out_tiles * g.splits;
  for (int j = 0; j < 4; ++j) { g.dst[j] = dst.ptr[j]; g.ld[j] = dst.ld[j]; g.bias[j] = dst.bias[j]; }
  CUtensorMap tX, tY;
  TPC_TRY(get_tmap(tc, X, M, Kin, ldx, WG_TB, 1, &tX));
  TPC_TRY(get_tmap(tc, dY, M, Nout, ldy, WG_TB, 1, &tY));
  const int grid = g.num_items < num_sms ? g.num_items : num_sms;
  wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM, stream>>>(tX, tY, g);
  TPC_CHECK_LAUNCH();
  return TPC_OK;
}

}  // namespace tpc
