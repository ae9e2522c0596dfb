// b2q_mlp.cu — K3: fused 3-layer MLP forward (in<=64 -> 256 -> 256 -> out<=32) on tcgen05 tensor cores.
//
// One CTA (256 threads: two warps per TMEM lane quarter, each taking half of the columns in the epilogues) per 128-row tile of the batch:
//   * weights live in HBM as ready-made shared-memory images (bf16, K-major, 128-byte swizzle, 64-column panels) and are
//     brought in by bulk async copies (cp.async.bulk -> UBLKCP) that complete on mbarriers;
//   * the input tile is converted f32 -> bf16 by the CTA's threads straight into the swizzled A-operand layout;
//   * each layer is a chain of tcgen05.mma (M=128, N=256|32, K=16) issued by ONE thread, accumulating in TMEM
//     (layer 1 -> columns 0..255, layer 2 -> 256..511, layer 3 -> 0..31); tcgen05.commit signals an mbarrier;
//   * the epilogue warps read the accumulator with tcgen05.ld (each thread owns one row = one TMEM lane), apply
//     bias+ReLU in f32, and write the bf16 activations back into the A-operand region for the next layer, so
//     activations never leave the SM; the last epilogue applies tanh / clamp / sampling / log-prob and stores f32.
// Shared memory: A 64 KB | W2 128 KB | W1 then W3 32 KB | biases 2.1 KB | barriers  = 226.2 KB (of 227 KB).
// Reference: Actor/Critic.forward (model/mujoco_model.py:44-89), SAC.predict/sample (alg/sac.py:60-75).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <new>
#include <string>
#include "../../include/b2q_mlp.h"
#include "b2q_tc.cuh"
#include "b2q_mlp_internal.h"
#include "b2q_philox.cuh"

using namespace b2q_tc;

namespace {

constexpr int HID = B2Q_MLP_HIDDEN;
constexpr int TILE_M = 128;
using b2q_mlp_img::IMG_W1; using b2q_mlp_img::IMG_W2; using b2q_mlp_img::IMG_W3; using b2q_mlp_img::IMG_BIAS; using b2q_mlp_img::IMG_BYTES;
using b2q_mlp_img::IMG_W2T; using b2q_mlp_img::IMG_W1A;
constexpr uint32_t SZ_W1A = (uint32_t)b2q_mlp_img::SZ_W1A, OFF_W1A_IN_W13 = 16384;   // the W1A image sits behind W3 in the (32 KB) W1/W3 region
constexpr uint32_t SZ_A = 65536, SZ_W2 = (uint32_t)b2q_mlp_img::SZ_W2, SZ_W13 = 32768, SZ_W1 = (uint32_t)b2q_mlp_img::SZ_W1, SZ_W3 = (uint32_t)b2q_mlp_img::SZ_W3,
                   SZ_BIAS = (uint32_t)b2q_mlp_img::SZ_BIAS;
constexpr uint32_t OFF_A = 0, OFF_W2 = OFF_A + SZ_A, OFF_W13 = OFF_W2 + SZ_W2, OFF_BIAS = OFF_W13 + SZ_W13, OFF_BAR = OFF_BIAS + SZ_BIAS;
constexpr uint32_t SMEM_BYTES = OFF_BAR + 64 + 512;   // + the head's log-prob exchange [128] f32
static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory per CTA");

using b2q_philox::philox_normal;

struct FwdArgs {
  const float* in1; const float* in2; int in1_dim, in_dim, out_dim, M, mode; uint64_t seed; const float* eps;
  float* out; float* logp; float* raw; const uint8_t* img; size_t img_stride;
  B2QMlpSaves sv; int save;
  float* da;   // input-gradient pass (see b2q_mlp_internal.h) or null
  const int* seed_ctr;   // device-side counter folded into the sampling key (b2q_philox.cuh) or null
};

// actor head of one row, actions j with (j & 1) == chalf: tanh(mean) or the rsample() + tanh-Gaussian log-prob; returns the row's log-prob share.
// AC > 0: the action dimension as a compile-time constant (the TMEM register array stays statically indexed); AC == 0: runtime dimension.
template <int AC>
__device__ __forceinline__ float head_actions(const uint32_t (&r)[32], const float* b3, const FwdArgs& a, int row, size_t orow, int chalf) {
  const int A = AC > 0 ? AC : (a.out_dim >> 1);
  const uint64_t seed_eff = b2q_philox::effective_seed(a.seed, a.seed_ctr);
  float y[32];
#pragma unroll
  for (int j = 0; j < 32; j++) y[j] = __uint_as_float(r[j]) + b3[j];
  float lp = 0.f;
#pragma unroll
  for (int j = 0; j < (AC > 0 ? AC : 16); j++) {
    if ((j & 1) != chalf || j >= A) continue;
    const float mean = y[j];
    float act;
    if (a.mode == B2Q_MLP_PREDICT) {
      act = tanhf(mean);                                                   // sac.py:60-63
    } else {
      const float ls = fminf(fmaxf(y[A + j], -20.f), 2.f), sd = expf(ls);  // mujoco_model.py:21-22,59
      const float e = a.eps ? a.eps[(size_t)row * A + j] : philox_normal(seed_eff, (uint32_t)row, (uint32_t)j);
      const float x = mean + sd * e;                                       // rsample
      act = tanhf(x);
      lp += -0.5f * e * e - ls - 0.9189385332046727f;                      // Normal.log_prob(x)
      lp -= logf((1.f - act * act) + 1e-6f);                               // sac.py:72
    }
    a.out[orow * A + j] = act;
  }
  return lp;
}

constexpr int NTHR = 256;
__global__ void __launch_bounds__(NTHR, 1) b2q_mlp_fwd_kernel(FwdArgs a) { pdl_sync();
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, net = blockIdx.y;
  const int trow = tid & (TILE_M - 1), chalf = tid >> 7;   // tile row owned in the epilogues; column half (0: cols 0..127, 1: 128..255)
  const int row0 = blockIdx.x * TILE_M, row = row0 + trow;
  const uint32_t sbase = smem_u32(smem);
  if ((sbase & 1023u) != 0) __trap();   // SWIZZLE_128B operands need a 1024-byte aligned base
  const uint32_t sA = sbase + OFF_A, sW2 = sbase + OFF_W2, sW13 = sbase + OFF_W13;
  const uint32_t bar_w1 = sbase + OFF_BAR, bar_w2 = bar_w1 + 8, bar_w3 = bar_w1 + 16, bar_mma = bar_w1 + 24, bar_w2t = bar_w1 + 32;
  volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + OFF_BAR + 40);
  const float* bias = reinterpret_cast<const float*>(smem + OFF_BIAS);
  const uint8_t* img = a.img + (size_t)net * a.img_stride;

  if (tid == 0) {
    mbar_init(bar_w1, 1); mbar_init(bar_w2, 1); mbar_init(bar_w3, 1); mbar_init(bar_mma, 1); mbar_init(bar_w2t, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // weights: W1 (+biases) first, then W2 in 32 KB pieces
    mbar_expect_tx(bar_w1, SZ_W1 + SZ_BIAS);
    bulk_g2s(sW13, img + IMG_W1, SZ_W1, bar_w1);
    bulk_g2s(sbase + OFF_BIAS, img + IMG_BIAS, SZ_BIAS, bar_w1);
    mbar_expect_tx(bar_w2, SZ_W2);
#pragma unroll
    for (int i = 0; i < 4; i++) bulk_g2s(sW2 + i * 32768u, img + IMG_W2 + (size_t)i * 32768u, 32768u, bar_w2);
  }
  if (warp == 1) {  // TMEM: all 512 columns (one CTA per SM by construction: 226 KB of shared memory)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<const uint32_t*>(tmem_slot))), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // input tile: f32 [rows, in_dim] (two sources concatenated) -> bf16 swizzled panel 0 (K padded to 64 with zeros)
  {
    const int in2_dim = a.in_dim - a.in1_dim;
    // 32 elements per thread (one column k = tid & 63, rows (tid >> 6) + 4 j), ALL loads in flight before the first use
    constexpr int PER = 64 * TILE_M / NTHR;
    const int k = tid & 63;
    const float* src = k < a.in1_dim ? a.in1 + k : (k < a.in_dim ? a.in2 + (k - a.in1_dim) : nullptr);
    const int ld = k < a.in1_dim ? a.in1_dim : in2_dim;
    float v[PER];
#pragma unroll
    for (int j = 0; j < PER; j++) {
      const int gr = row0 + (tid >> 6) + 4 * j;
      v[j] = (src && gr < a.M) ? __ldg(src + (size_t)gr * ld) : 0.f;
    }
#pragma unroll
    for (int j = 0; j < PER; j++)
      *reinterpret_cast<__nv_bfloat16*>(smem + OFF_A + sw128_offset((tid >> 6) + 4 * j, k, TILE_M)) = __float2bfloat16(v[j]);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t lane_addr = tmem + ((uint32_t)((warp & 3) * 32) << 16);   // a warp may touch TMEM lanes 32*(warp%4)..+31

  // Activation dumps for the backward pass, written from the shared-memory tile the epilogue just produced (while the next layer's MMAs
  // read the same tile): row-major [batch][256] with one full 512-byte row per warp instruction, and the [256][batch] copy as 16-byte
  // runs of eight consecutive batch rows per column (the epilogue's own registers hold one ROW per thread: its stores would be 2-byte
  // scatters for the transposed copy and half-used sectors for the row-major one).
  auto dump_tile = [&](__nv_bfloat16* d_rm, __nv_bfloat16* d_t, int ncols /*256: hidden activations (per net), 64: the input tile (shared by the nets)*/) {
    const int nrows = min(TILE_M, a.M - row0);
    const size_t nbase = ncols == HID ? (size_t)net : 0;
    if (d_rm) {
      const int lane = tid & 31;
      if (lane * 8 < ncols)
        for (int r = warp; r < nrows; r += NTHR / 32) {
          const uint4 v = *reinterpret_cast<const uint4*>(smem + OFF_A + sw128_offset(r, lane * 8, TILE_M));
          *reinterpret_cast<uint4*>(d_rm + (nbase * a.M + row0 + r) * ncols + lane * 8) = v;
        }
    }
    if (d_t && tid < ncols) {
      const int c = tid;                                   // NTHR == HID: one column per thread
      __nv_bfloat16* dst = d_t + (nbase * ncols + c) * a.M + row0;
#pragma unroll 4
      for (int r0 = 0; r0 < TILE_M; r0 += 8) {
        uint32_t w[4];                                     // packed in registers (a local bf16[8] would live in local memory)
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const uint32_t lo = *reinterpret_cast<const uint16_t*>(smem + OFF_A + sw128_offset(r0 + 2 * i, c, TILE_M));
          const uint32_t hi = *reinterpret_cast<const uint16_t*>(smem + OFF_A + sw128_offset(r0 + 2 * i + 1, c, TILE_M));
          w[i] = lo | (hi << 16);
        }
        if (r0 + 8 <= nrows) *reinterpret_cast<uint4*>(dst + r0) = make_uint4(w[0], w[1], w[2], w[3]);
        else {
          uint16_t* d16 = reinterpret_cast<uint16_t*>(dst + r0);
#pragma unroll
          for (int i = 0; i < 8; i++) if (r0 + i < nrows) d16[i] = (uint16_t)(w[i >> 1] >> (16 * (i & 1)));
        }
      }
    }
  };
  // ---- layer 1: [128 x 64] x [256 x 64]^T -> TMEM cols 0..255
  if (tid == 0) {
    mbar_wait(bar_w1, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(TILE_M, HID);
#pragma unroll
    for (int ks = 0; ks < 4; ks++) umma_f16(tmem, umma_desc(sA + ks * 32), umma_desc(sW13 + ks * 32), idesc, ks > 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  if (a.save && net == 0 && (a.sv.x_rm || a.sv.x_t)) { dump_tile(a.sv.x_rm, a.sv.x_t, 64); __syncthreads(); }   // the input tile (panel 0), before epilogue 1 overwrites it
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  if (tid == 0) {  // W1 is consumed: reuse its region for W3
    mbar_expect_tx(bar_w3, SZ_W3 + (a.da ? SZ_W1A : 0u));
    bulk_g2s(sW13, img + IMG_W3, SZ_W3, bar_w3);
    if (a.da) bulk_g2s(sW13 + OFF_W1A_IN_W13, img + IMG_W1A, SZ_W1A, bar_w3);
  }
  // relu'(h1) of the thread's 128 columns (bit j of word cc: column 32 (4 chalf + cc) + j), kept for the input-gradient pass
  uint32_t m1[4] = {0u, 0u, 0u, 0u};
  auto epilogue_hidden = [&](uint32_t col_base, const float* b, bool keep_mask) {
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++) {
      const int cc = 4 * chalf + c4;
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(lane_addr + col_base + cc * 32, r);
      uint32_t mk = 0u;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          float v0 = fmaxf(__uint_as_float(r[j0 + j]) + b[cc * 32 + j0 + j], 0.f);
          float v1 = fmaxf(__uint_as_float(r[j0 + j + 1]) + b[cc * 32 + j0 + j + 1], 0.f);
          mk |= (v0 > 0.f ? 1u : 0u) << (j0 + j);
          mk |= (v1 > 0.f ? 1u : 0u) << (j0 + j + 1);
          __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(smem + OFF_A + sw128_offset(trow, cc * 32 + j0, TILE_M)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      if (keep_mask) m1[c4] = mk;
    }
  };
  epilogue_hidden(0, bias, a.da != nullptr);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- layer 2: [128 x 256] x [256 x 256]^T -> TMEM cols 256..511
  if (tid == 0) {
    mbar_wait(bar_w2, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(TILE_M, HID);
#pragma unroll
    for (int ks = 0; ks < 16; ks++)
      umma_f16(tmem + 256, umma_desc(sA + (ks >> 2) * (TILE_M * 128) + (ks & 3) * 32), umma_desc(sW2 + (ks >> 2) * (HID * 128) + (ks & 3) * 32), idesc, ks > 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  if (a.save) { dump_tile(a.sv.h1_rm, a.sv.h1_t, HID); __syncthreads(); }   // reads of the h1 tile end before any thread's next epilogue overwrites it
  mbar_wait(bar_mma, 1);
  tc_fence_after();
  if (a.da && tid == 0) {   // layer 2 has consumed W2: its region takes the W2^T image for the input-gradient pass
    mbar_expect_tx(bar_w2t, SZ_W2);
#pragma unroll
    for (int i = 0; i < 4; i++) bulk_g2s(sW2 + i * 32768u, img + IMG_W2T + (size_t)i * 32768u, 32768u, bar_w2t);
  }
  epilogue_hidden(256, bias + HID, false);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- layer 3: [128 x 256] x [32 x 256]^T -> TMEM cols 0..31
  if (tid == 0) {
    mbar_wait(bar_w3, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc(TILE_M, 32);
#pragma unroll
    for (int ks = 0; ks < 16; ks++)
      umma_f16(tmem, umma_desc(sA + (ks >> 2) * (TILE_M * 128) + (ks & 3) * 32), umma_desc(sW13 + (ks >> 2) * (32 * 128) + (ks & 3) * 32), idesc, ks > 0);
    umma_commit(bar_mma);
  }
  __syncwarp();
  if (a.save) dump_tile(a.sv.h2_rm, a.sv.h2_t, HID);   // the head epilogue does not write the tile: no barrier needed
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  // Head epilogue.  RAW (critics, BC): the row's thread of column half 0 stores its outputs.  PREDICT / SAMPLE (actor): BOTH threads of a row
  // (the two column halves read the same 32 TMEM columns) take every second action — the per-action tanh / exp / log / counter-RNG work is the
  // longest stretch of the actor forward — and the log-prob halves meet in shared memory.  All indices are compile-time (no local arrays).
  {
    uint32_t r[32];
    __syncwarp();
    tmem_ld32(lane_addr, r);
    const float* b3 = bias + 2 * HID;
    float* slp = reinterpret_cast<float*>(smem + OFF_BAR + 64);            // [128] log-prob share of column half 1
    const int od = a.out_dim, A = od >> 1;
    const size_t orow = (size_t)net * a.M + row;
    if (chalf == 0 && row < a.M && (a.raw || a.mode == B2Q_MLP_RAW)) {
#pragma unroll
      for (int j = 0; j < 32; j++) {
        if (j < od) {
          const float y = __uint_as_float(r[j]) + b3[j];
          if (a.raw) a.raw[orow * od + j] = y;
          if (a.mode == B2Q_MLP_RAW) a.out[orow * od + j] = y;
        }
      }
    }
    if (a.mode != B2Q_MLP_RAW) {
      float lp = 0.f;
      if (row < a.M) {
        if (A == 12) lp = head_actions<12>(r, b3, a, row, orow, chalf);        // the A1's action dimension: every index compile-time
        else lp = head_actions<0>(r, b3, a, row, orow, chalf);
      }
      if (a.mode == B2Q_MLP_SAMPLE && a.logp) {
        if (chalf == 1) slp[trow] = lp;
        __syncthreads();
        if (chalf == 0 && row < a.M) a.logp[orow] = lp + slp[trow];
      }
    }
  }
  if (a.da) {
    // ---- input-gradient pass (out_dim == 1): unit output gradient back to the action columns of the input, on the same tile.
    // (a) dh2 = W3 . relu'(h2), in place over the h2 tile (layer 3's MMAs have completed: every thread waited on bar_mma above).  Row 0 of the
    //     W3 operand image is W3 itself: 128 contiguous bytes per 64-column panel.
#pragma unroll 4
    for (int c0 = 128 * chalf; c0 < 128 * chalf + 128; c0 += 8) {
      uint8_t* hp = smem + OFF_A + sw128_offset(trow, c0, TILE_M);
      const uint4 hv = *reinterpret_cast<const uint4*>(hp);
      const uint4 wv = *reinterpret_cast<const uint4*>(smem + OFF_W13 + (c0 >> 6) * (32 * 128) + (c0 & 63) * 2);
      const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) o[j] = ((hw[j] & 0x7fffu) ? (ww[j] & 0xffffu) : 0u) | ((hw[j] & 0x7fff0000u) ? (ww[j] & 0xffff0000u) : 0u);   // h2 = relu(.) >= 0: nonzero <=> positive
      *reinterpret_cast<uint4*>(hp) = make_uint4(o[0], o[1], o[2], o[3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // (b) dh1 pre-mask = dh2 . W2  -> TMEM cols 256..511 (B operand: the W2^T image)
    if (tid == 0) {
      mbar_wait(bar_w2t, 0);
      tc_fence_after();
      const uint32_t idesc = umma_idesc(TILE_M, HID);
#pragma unroll
      for (int ks = 0; ks < 16; ks++)
        umma_f16(tmem + 256, umma_desc(sA + (ks >> 2) * (TILE_M * 128) + (ks & 3) * 32), umma_desc(sW2 + (ks >> 2) * (HID * 128) + (ks & 3) * 32), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, 1);
    tc_fence_after();
    // (c) dh1 = . relu'(h1) (bits kept from epilogue 1) -> bf16 over the tile
#pragma unroll
    for (int c4 = 0; c4 < 4; c4++) {
      const int cc = 4 * chalf + c4;
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(lane_addr + 256 + cc * 32, r);
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        uint32_t pk[4];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const float v0 = ((m1[c4] >> (j0 + j)) & 1u) ? __uint_as_float(r[j0 + j]) : 0.f, v1 = ((m1[c4] >> (j0 + j + 1)) & 1u) ? __uint_as_float(r[j0 + j + 1]) : 0.f;
          __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
          pk[j >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(smem + OFF_A + sw128_offset(trow, cc * 32 + j0, TILE_M)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // (d) da = dh1 . W1[:, action]  -> TMEM cols 0..15 (B operand: the W1A image, N = 16)
    if (tid == 0) {
      const uint32_t idesc = umma_idesc(TILE_M, 16);
#pragma unroll
      for (int ks = 0; ks < 16; ks++)
        umma_f16(tmem, umma_desc(sA + (ks >> 2) * (TILE_M * 128) + (ks & 3) * 32), umma_desc(sW13 + OFF_W1A_IN_W13 + (ks >> 2) * (16 * 128) + (ks & 3) * 32), idesc, ks > 0);
      umma_commit(bar_mma);
    }
    __syncwarp();
    mbar_wait(bar_mma, 0);
    tc_fence_after();
    if (chalf == 0) {
      uint32_t r[32];
      __syncwarp();
      tmem_ld32(lane_addr, r);          // columns 16..31 are stale head outputs: not stored
      if (row < a.M) {
        float4* dst = reinterpret_cast<float4*>(a.da + ((size_t)net * a.M + row) * 16);
#pragma unroll
        for (int j = 0; j < 4; j++) dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// f32 nn.Linear weights -> bf16 swizzled operand images (+ f32 biases) in the per-net image
__global__ void b2q_mlp_pack_kernel(uint8_t* img, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3,
                                    int in_dim, int out_dim, int a_off, int a_dim) { pdl_sync();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < HID * HID) {   // W2^T image: B operand [N = in][K = out] of the input-gradient pass
    int n = i >> 8, k = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(img + IMG_W2T + sw128_offset(n, k, HID)) = __float2bfloat16(w2[(size_t)k * HID + n]);
  }
  if (i < 16 * HID) {    // W1A image: [N = 16 action slots][K = hidden]
    int n = i >> 8, k = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(img + IMG_W1A + sw128_offset(n, k, 16)) = __float2bfloat16(n < a_dim ? w1[(size_t)k * in_dim + a_off + n] : 0.f);
  }
  if (i < HID * 64) {  // W1 [256 x 64 padded]
    int n = i >> 6, k = i & 63;
    *reinterpret_cast<__nv_bfloat16*>(img + IMG_W1 + sw128_offset(n, k, HID)) = __float2bfloat16(k < in_dim ? w1[(size_t)n * in_dim + k] : 0.f);
  }
  if (i < HID * HID) {
    int n = i >> 8, k = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(img + IMG_W2 + sw128_offset(n, k, HID)) = __float2bfloat16(w2[(size_t)n * HID + k]);
  }
  if (i < 32 * HID) {
    int n = i >> 8, k = i & 255;
    *reinterpret_cast<__nv_bfloat16*>(img + IMG_W3 + sw128_offset(n, k, 32)) = __float2bfloat16(n < out_dim ? w3[(size_t)n * HID + k] : 0.f);
  }
  float* bias = reinterpret_cast<float*>(img + IMG_BIAS);
  if (i < HID) { bias[i] = b1[i]; bias[HID + i] = b2[i]; }
  if (i < 32) bias[2 * HID + i] = i < out_dim ? b3[i] : 0.f;
}

}  // namespace

struct B2QMlp {
  int device, in_dim, out_dim, nets;
  int a_off = 0, a_dim = 0;
  uint8_t* img = nullptr;
  std::string err;
  int64_t launches = 0;
};

extern "C" {

int b2q_mlp_create(int device, int in_dim, int out_dim, int nets, B2QMlpHandle* out) {
  if (!out || in_dim < 1 || in_dim > B2Q_MLP_MAX_IN || out_dim < 1 || out_dim > B2Q_MLP_MAX_OUT || nets < 1 || nets > 8) return -1;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return -2;
  B2QMlp* h = new (std::nothrow) B2QMlp();
  if (!h) return -3;
  h->device = device; h->in_dim = in_dim; h->out_dim = out_dim; h->nets = nets;
  cudaSetDevice(device);
  if (cudaMalloc(&h->img, IMG_BYTES * nets) != cudaSuccess) { delete h; return -3; }
  cudaMemset(h->img, 0, IMG_BYTES * nets);
  if (cudaFuncSetAttribute(b2q_mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES) != cudaSuccess) { cudaFree(h->img); delete h; return -2; }
  *out = h;
  return 0;
}
int b2q_mlp_destroy(B2QMlpHandle h) {
  if (!h) return -1;
  cudaSetDevice(h->device);
  cudaFree(h->img);
  delete h;
  return 0;
}
const char* b2q_mlp_last_error(B2QMlpHandle h) { return h ? h->err.c_str() : "null handle / create failed"; }
int64_t b2q_mlp_launch_count(B2QMlpHandle h) { return h ? h->launches : 0; }
int b2q_mlp_set_action_slice(B2QMlpHandle h, int a_off, int a_dim) {
  if (!h || a_off < 0 || a_dim < 0 || a_dim > 16 || a_off + a_dim > h->in_dim) return -1;
  h->a_off = a_off; h->a_dim = a_dim;
  return 0;
}
uint8_t* b2q_mlp_image(B2QMlpHandle h, int net) { return (h && net >= 0 && net < h->nets) ? h->img + (size_t)net * IMG_BYTES : nullptr; }

int b2q_mlp_set_weights(B2QMlpHandle h, int net, const float* w1, const float* b1, const float* w2, const float* b2, const float* w3, const float* b3, void* stream) {
  if (!h || net < 0 || net >= h->nets || !w1 || !b1 || !w2 || !b2 || !w3 || !b3) { if (h) h->err = "b2q_mlp_set_weights: bad argument"; return -1; }
  pdl_launch(b2q_mlp_pack_kernel, dim3((HID * HID + 255) / 256), dim3(256), 0, (cudaStream_t)stream, h->img + (size_t)net * IMG_BYTES, w1, b1, w2, b2, w3, b3, h->in_dim, h->out_dim, h->a_off, h->a_dim);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return -2; }
  return 0;
}

int b2q_mlp_forward_ex(B2QMlpHandle h, const float* in1, int in1_dim, const float* in2, int M, int mode, uint64_t seed, const float* eps, float* out,
                       float* logp, float* raw, const B2QMlpSaves* saves, float* da, const int* seed_ctr, void* stream) {
  if (!h || !in1 || !out || M < 1 || in1_dim < 1 || in1_dim > h->in_dim || (in1_dim < h->in_dim && !in2) || mode < 0 || mode > 2 ||
      (mode != B2Q_MLP_RAW && (h->out_dim & 1)) || (da && (h->out_dim != 1 || h->a_dim < 1 || saves))) { if (h) h->err = "b2q_mlp_forward: bad argument"; return -1; }
  { int cur = -1; if (cudaGetDevice(&cur) != cudaSuccess || cur != h->device) cudaSetDevice(h->device); }   // handles are per GPU
  FwdArgs a{in1, in2, in1_dim, h->in_dim, h->out_dim, M, mode, seed, eps, out, logp, raw, h->img, IMG_BYTES, B2QMlpSaves{}, 0, da, seed_ctr};
  if (saves) { a.sv = *saves; a.save = 1; }
  dim3 grid((M + TILE_M - 1) / TILE_M, h->nets);
  pdl_launch(b2q_mlp_fwd_kernel, dim3(grid), dim3(NTHR), SMEM_BYTES, (cudaStream_t)stream, a);
  h->launches++;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = cudaGetErrorString(e); return -2; }
  return 0;
}
int b2q_mlp_forward(B2QMlpHandle h, const float* in1, int in1_dim, const float* in2, int M, int mode, uint64_t seed, const float* eps, float* out,
                    float* logp, float* raw, void* stream) {
  return b2q_mlp_forward_ex(h, in1, in1_dim, in2, M, mode, seed, eps, out, logp, raw, nullptr, nullptr, nullptr, stream);
}

}  // extern "C"
