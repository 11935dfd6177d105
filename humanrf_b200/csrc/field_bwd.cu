// Fused backward of the HumanRF radiance field for sm_100a (autograd of humanrf.py:158-208).
// Per 128-sample tile (thread r = sample r = operand row r = TMEM lane r):
//   1. forward recompute of both MLPs from the saved composed features (or a re-encode),
//      keeping every activation tile in shared memory (bf16, UMMA K-major core-matrix layout);
//   2. MLP backward on the tensor cores.  The SAME shared-memory tiles serve three roles by
//      switching descriptors only:  K-major A operand of the forward / dgrad GEMMs, and
//      MN-major A / B operands of the wgrad GEMMs (contraction over the 128 samples).  The
//      packed forward weight blob is re-used as the MN-major B operand of the dgrad GEMMs.
//      Weight-gradient accumulators live in TMEM for the whole persistent CTA (M=64 tiles)
//      and are flushed once with fp32 atomics;
//   3. scatter: per level re-gather the 4x8 corners (needed for the vector gradients),
//      red.global.add.v2.f32 into the fp32 table gradients, warp-combined adds for the time axis.
#include "field_common.cuh"

namespace hrf {

struct __align__(1024) BwdSmem {
  unsigned char w[kWBlobBytes];
  unsigned char feat[kTile * 32 * 2];  // composed features            (A32)
  unsigned char cin[kTile * 32 * 2];   // colour-net input              (A32)
  unsigned char hs[kTile * 64 * 2];    // sigma hidden, then d(hidden)  (A64)
  unsigned char h1[kTile * 64 * 2];    // colour hidden 1, then its gradient, then dFeat staging
  unsigned char h2[kTile * 64 * 2];    // colour hidden 2, then its gradient
  unsigned char g3[kTile * 16 * 2];    // d(colour pre-activation)  [128,16]
  unsigned char gs[kTile * 16 * 2];    // d(sigma-net output)       [128,16]
  uint64_t bar_w, bar_mma;
  uint32_t tmem_base;
};

// TMEM column map (256 columns allocated): work area + persistent weight-gradient accumulators
constexpr uint32_t kColWork = 0;     // 64 cols: layer outputs / dgrad results
constexpr uint32_t kColW1s = 64;     // dW1s   [64 out, 32 in]
constexpr uint32_t kColW2s = 96;     // dW2s^T [64 in, 16 out]
constexpr uint32_t kColW1c = 112;    // dW1c   [64 out, 32 in]
constexpr uint32_t kColW2c = 144;    // dW2c   [64 out, 64 in]
constexpr uint32_t kColW3c = 208;    // dW3c^T [64 in, 16 out]
constexpr uint32_t kTmemCols = 256;

struct BwdArgs {
  hrf_field f;
  hrf_samples s;
  const hrf_segment_grads* seg_grads;
  const float* d_sigma;
  const float* d_rgb;
  const uint4* feat_in;  // bf16 [N,32] saved by the forward, or NULL (re-encode)
  float* d_mlp;
};

// D[128,Nin] = G[128,Kout] * W[Kout,Nin]  : A = gradient tile (K-major), B = forward blob read MN-major
__device__ __forceinline__ void issue_dgrad(uint32_t tmem_d, uint32_t g_addr, uint32_t w_addr, int Nin, int Kout) {
  const uint32_t idesc = make_idesc_bf16(kTile, Nin, 0, 1);
  const uint32_t b_sbo = (uint32_t)(Kout >> 3) * 128u;
  for (int k = 0; k < Kout / 16; ++k) {
    const uint64_t ad = make_smem_desc(g_addr + (uint32_t)k * 2u * kAChunk, kAChunk, 128u);
    const uint64_t bd = make_smem_desc(w_addr + (uint32_t)k * 256u, 128u, b_sbo);
    umma_bf16(tmem_d, ad, bd, idesc, k > 0 ? 1u : 0u);
  }
}
// D[64,N] (+)= X^T[64,128] * Y[128,N] : both tiles read MN-major, contraction over the 128 samples
__device__ __forceinline__ void issue_wgrad(uint32_t tmem_d, uint32_t x_addr, uint32_t y_addr, int N, bool acc) {
  const uint32_t idesc = make_idesc_bf16(64, N, 1, 1);
  for (int k = 0; k < kTile / 16; ++k) {
    const uint64_t ad = make_smem_desc(x_addr + (uint32_t)k * 256u, 128u, kAChunk);
    const uint64_t bd = make_smem_desc(y_addr + (uint32_t)k * 256u, 128u, kAChunk);
    umma_bf16(tmem_d, ad, bd, idesc, (acc || k > 0) ? 1u : 0u);
  }
}

template <class F>
__device__ __forceinline__ void mma_round(BwdSmem& sm, uint32_t& phase, F&& issue) {
  tc_fence_before();
  fence_proxy_async_smem();
  __syncthreads();
  if (threadIdx.x == 0) {
    tc_fence_after();
    issue();
    umma_commit(&sm.bar_mma);
  }
  mbar_wait(&sm.bar_mma, phase);
  phase ^= 1u;
  tc_fence_after();
}

// store relu(v) as a K=64 bf16 tile row
__device__ __forceinline__ void store_relu64(unsigned char* buf, uint32_t roff, const float* v) {
#pragma unroll
  for (int kg = 0; kg < 8; ++kg) {
    uint4 q;
    q.x = pack_bf16x2(fmaxf(v[kg * 8 + 0], 0.f), fmaxf(v[kg * 8 + 1], 0.f));
    q.y = pack_bf16x2(fmaxf(v[kg * 8 + 2], 0.f), fmaxf(v[kg * 8 + 3], 0.f));
    q.z = pack_bf16x2(fmaxf(v[kg * 8 + 4], 0.f), fmaxf(v[kg * 8 + 5], 0.f));
    q.w = pack_bf16x2(fmaxf(v[kg * 8 + 6], 0.f), fmaxf(v[kg * 8 + 7], 0.f));
    *reinterpret_cast<uint4*>(buf + kg * kAChunk + roff) = q;
  }
}
// in place: tile row holds relu(h) (bf16); replace by g * (h > 0) (bf16)
__device__ __forceinline__ void relu_backward_inplace64(unsigned char* buf, uint32_t roff, const float* g) {
#pragma unroll
  for (int kg = 0; kg < 8; ++kg) {
    uint4* p = reinterpret_cast<uint4*>(buf + kg * kAChunk + roff);
    const uint4 a = *p;
    const uint32_t av[4] = {a.x, a.y, a.z, a.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float lo = (av[j] & 0x7fffu) != 0u && !(av[j] & 0x8000u) ? g[kg * 8 + 2 * j] : 0.f;
      const float hi = (av[j] & 0x7fff0000u) != 0u && !(av[j] & 0x80000000u) ? g[kg * 8 + 2 * j + 1] : 0.f;
      o[j] = pack_bf16x2(lo, hi);
    }
    *p = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
__device__ __forceinline__ void red_add2(float* addr, float a, float b) {
  atomicAdd(reinterpret_cast<float2*>(addr), make_float2(a, b));  // red.global.add.v2.f32
}

// Scatter d(features) of one sample into the table / vector gradients of its segment.
__device__ __forceinline__ void scatter_sample(const hrf_field& f, const Sample& s, const hrf_segment_grads* sgs,
                                               const float2* dfeat /* smem, [level][row] */, int row) {
  const bool valid = s.seg != nullptr;
  const hrf_segment* sg = valid ? s.seg : f.segments;
  const hrf_segment_grads gr = sgs[valid ? (int)(s.seg - f.segments) : 0];
  const uint32_t hmask = sg->hashed_mask;
  const float* vec = sg->vectors;
  const VecTap tx = make_tap(s.x, f.vec_res, 0), ty = make_tap(s.y, f.vec_res, 1), tz = make_tap(s.z, f.vec_res, 2),
               tt = make_tap(s.t, f.vec_res, 3);
  // time-axis taps are usually identical across the warp (one ray = one frame): combine first
  const uint32_t full = 0xffffffffu;
  // (shuffles are executed unconditionally by all 32 lanes; only the comparison is predicated)
  const uint32_t o0_first = __shfl_sync(full, tt.o0, 0), o1_first = __shfl_sync(full, tt.o1, 0);
  const unsigned long long vec_first = __shfl_sync(full, (unsigned long long)gr.vectors, 0);
  const bool same = valid && tt.o0 == o0_first && tt.o1 == o1_first && (unsigned long long)gr.vectors == vec_first;
  const bool t_uniform = __all_sync(full, same);
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int l = 0; l < HRF_N_LEVELS; ++l) {
    const float2 dO = valid ? dfeat[l * kTile + row] : make_float2(0.f, 0.f);
    const float scale = f.level_scale[l];
    const uint32_t res = f.level_res[l];
    const uint32_t off = sg->level_offset[l];
    const uint32_t size = sg->level_size[l];
    const bool hashed = (hmask >> l) & 1u;
    const Cell cx = to_cell(scale, s.x), cy = to_cell(scale, s.y), cz = to_cell(scale, s.z), ct = to_cell(scale, s.t);
    const float2 vx = lerp_tap(vec, tx, 2 * l), vy = lerp_tap(vec, ty, 2 * l), vz = lerp_tap(vec, tz, 2 * l),
                 vt = lerp_tap(vec, tt, 2 * l);
    float2 e[4];
    // grid k pairs with vector: xyz<->t, xyt<->z, yzt<->x, xzt<->y (tensor_composition.cu:49-52)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const Cell a = (k == 2) ? cy : cx;
      const Cell b = (k == 0 || k == 1) ? cy : cz;
      const Cell c = (k == 0) ? cz : ct;
      const float2 v = (k == 0) ? vt : (k == 1) ? vz : (k == 2) ? vx : vy;
      uint32_t idx[8];
      float w[8];
      corner_indices(hashed, res, size, a, b, c, idx);
      corner_weights(a, b, c, w);
      const uint32_t* tab = sg->grid[k] + off;
      float* gtab = gr.grid[k] + 2 * (size_t)off;
      float2 acc = make_float2(0.f, 0.f);
      const float gx = v.x * dO.x, gy = v.y * dO.y;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t raw = __ldg(tab + idx[q]);
        acc.x = __fmaf_rn(w[q], bf16_lo(raw), acc.x);
        acc.y = __fmaf_rn(w[q], bf16_hi(raw), acc.y);
        if (valid && (gx != 0.f || gy != 0.f)) red_add2(gtab + 2 * (size_t)idx[q], w[q] * gx, w[q] * gy);
      }
      e[k] = acc;
    }
    if (valid) {
      // d vectors[axis][i0/i1][2l..2l+1] (tensor_composition.cu:109-111)
      const float2 dvx = make_float2(e[2].x * dO.x, e[2].y * dO.y), dvy = make_float2(e[3].x * dO.x, e[3].y * dO.y),
                   dvz = make_float2(e[1].x * dO.x, e[1].y * dO.y);
      red_add2(gr.vectors + tx.o0 + 2 * l, dvx.x * (1.f - tx.frac), dvx.y * (1.f - tx.frac));
      red_add2(gr.vectors + tx.o1 + 2 * l, dvx.x * tx.frac, dvx.y * tx.frac);
      red_add2(gr.vectors + ty.o0 + 2 * l, dvy.x * (1.f - ty.frac), dvy.y * (1.f - ty.frac));
      red_add2(gr.vectors + ty.o1 + 2 * l, dvy.x * ty.frac, dvy.y * ty.frac);
      red_add2(gr.vectors + tz.o0 + 2 * l, dvz.x * (1.f - tz.frac), dvz.y * (1.f - tz.frac));
      red_add2(gr.vectors + tz.o1 + 2 * l, dvz.x * tz.frac, dvz.y * tz.frac);
    }
    float2 dvt = make_float2(e[0].x * dO.x, e[0].y * dO.y);
    if (t_uniform) {
      const float a0 = warp_sum_f(dvt.x * (1.f - tt.frac)), a1 = warp_sum_f(dvt.y * (1.f - tt.frac));
      const float b0 = warp_sum_f(dvt.x * tt.frac), b1 = warp_sum_f(dvt.y * tt.frac);
      if (lane == 0) {
        red_add2(gr.vectors + tt.o0 + 2 * l, a0, a1);
        red_add2(gr.vectors + tt.o1 + 2 * l, b0, b1);
      }
    } else if (valid) {
      red_add2(gr.vectors + tt.o0 + 2 * l, dvt.x * (1.f - tt.frac), dvt.y * (1.f - tt.frac));
      red_add2(gr.vectors + tt.o1 + 2 * l, dvt.x * tt.frac, dvt.y * tt.frac);
    }
  }
}

__global__ void __launch_bounds__(kTile, 2) field_backward_kernel(const __grid_constant__ BwdArgs args) {
  extern __shared__ unsigned char smem_raw[];
  BwdSmem& sm = *reinterpret_cast<BwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int tid = threadIdx.x;
  const hrf_field& f = args.f;
  const uint32_t roff = a_row_off(tid);

  if (tid == 0) {
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_mma, 1);
    fence_mbar_init();
  }
  if (tid < 32) {
    tmem_alloc(&sm.tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    mbar_arrive_expect_tx(&sm.bar_w, kWBlobBytes);
    tma_load_1d(sm.w, f.mlp_blob, kWBlobBytes, &sm.bar_w);
  }
  const uint32_t tm = sm.tmem_base;
  const uint32_t trow = tm + ((uint32_t)(tid & ~31) << 16);  // this warp's TMEM lanes
  const uint32_t wbase = smem_u32(sm.w);
  const uint32_t a_feat = smem_u32(sm.feat), a_cin = smem_u32(sm.cin), a_hs = smem_u32(sm.hs),
                 a_h1 = smem_u32(sm.h1), a_h2 = smem_u32(sm.h2), a_g3 = smem_u32(sm.g3), a_gs = smem_u32(sm.gs);
  bool weights_ready = false, have_acc = false;
  uint32_t phase = 0;

  const int64_t n = args.s.num_samples;
  const int64_t num_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t i = tile * kTile + tid;
    const bool valid = i < n;
    const Sample s = load_sample(f, args.s, i, true);
    if (args.feat_in != nullptr) {
#pragma unroll
      for (int kg = 0; kg < 4; ++kg)
        *reinterpret_cast<uint4*>(sm.feat + kg * kAChunk + roff) =
            (valid && s.seg != nullptr) ? __ldg(args.feat_in + i * 4 + kg) : make_uint4(0, 0, 0, 0);
    } else {
      encode_to_smem(f, s, sm.feat, tid);
    }
    if (!weights_ready) {
      mbar_wait(&sm.bar_w, 0);
      weights_ready = true;
    }
    float v[64], o[16];
    // ---------------- forward recompute ----------------
    mma_round(sm, phase, [&] { issue_layer(tm + kColWork, a_feat, wbase + kWSig1, 64, 32); });
    tmem_ld64(trow + kColWork, v);
    store_relu64(sm.hs, roff, v);
    mma_round(sm, phase, [&] { issue_layer(tm + kColWork, a_hs, wbase + kWSig2, 16, 64); });
    tmem_ld16(trow + kColWork, o);
    const float h0 = o[0];
    {
      float sh[16];
      sh4(s.dx, s.dy, s.dz, sh);
      *reinterpret_cast<uint4*>(sm.cin + 0 * kAChunk + roff) = make_uint4(
          pack_bf16x2(sh[0], sh[1]), pack_bf16x2(sh[2], sh[3]), pack_bf16x2(sh[4], sh[5]), pack_bf16x2(sh[6], sh[7]));
      *reinterpret_cast<uint4*>(sm.cin + 1 * kAChunk + roff) =
          make_uint4(pack_bf16x2(sh[8], sh[9]), pack_bf16x2(sh[10], sh[11]), pack_bf16x2(sh[12], sh[13]),
                     pack_bf16x2(sh[14], sh[15]));
      *reinterpret_cast<uint4*>(sm.cin + 2 * kAChunk + roff) = make_uint4(
          pack_bf16x2(o[1], o[2]), pack_bf16x2(o[3], o[4]), pack_bf16x2(o[5], o[6]), pack_bf16x2(o[7], o[8]));
      *reinterpret_cast<uint4*>(sm.cin + 3 * kAChunk + roff) =
          make_uint4(pack_bf16x2(o[9], o[10]), pack_bf16x2(o[11], o[12]), pack_bf16x2(o[13], o[14]),
                     pack_bf16x2(o[15], 1.0f));
    }
    mma_round(sm, phase, [&] { issue_layer(tm + kColWork, a_cin, wbase + kWCol1, 64, 32); });
    tmem_ld64(trow + kColWork, v);
    store_relu64(sm.h1, roff, v);
    mma_round(sm, phase, [&] { issue_layer(tm + kColWork, a_h1, wbase + kWCol2, 64, 64); });
    tmem_ld64(trow + kColWork, v);
    store_relu64(sm.h2, roff, v);
    mma_round(sm, phase, [&] { issue_layer(tm + kColWork, a_h2, wbase + kWCol3, 16, 64); });
    tmem_ld16(trow + kColWork, o);

    // ---------------- backward ----------------
    {  // d(colour pre-activation) = d_rgb * rgb * (1 - rgb), cols 3..15 = 0
      float d3[3] = {0.f, 0.f, 0.f};
      if (valid && args.d_rgb != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float r = 1.f / (1.f + __expf(-o[c]));
          d3[c] = args.d_rgb[3 * i + c] * r * (1.f - r);
        }
      }
      *reinterpret_cast<uint4*>(sm.g3 + 0 * kAChunk + roff) =
          make_uint4(pack_bf16x2(d3[0], d3[1]), pack_bf16x2(d3[2], 0.f), 0u, 0u);
      *reinterpret_cast<uint4*>(sm.g3 + 1 * kAChunk + roff) = make_uint4(0u, 0u, 0u, 0u);
    }
    mma_round(sm, phase, [&] {
      issue_wgrad(tm + kColW3c, a_h2, a_g3, 16, have_acc);                   // dW3c^T += H2^T dO3
      issue_dgrad(tm + kColWork, a_g3, wbase + kWCol3, 64, 16);              // dH2 = dO3 W3c
    });
    tmem_ld64(trow + kColWork, v);
    relu_backward_inplace64(sm.h2, roff, v);
    mma_round(sm, phase, [&] {
      issue_wgrad(tm + kColW2c, a_h2, a_h1, 64, have_acc);                   // dW2c += dH2^T H1
      issue_dgrad(tm + kColWork, a_h2, wbase + kWCol2, 64, 64);              // dH1 = dH2 W2c
    });
    tmem_ld64(trow + kColWork, v);
    relu_backward_inplace64(sm.h1, roff, v);
    mma_round(sm, phase, [&] {
      issue_wgrad(tm + kColW1c, a_h1, a_cin, 32, have_acc);                  // dW1c += dH1^T Cin
      issue_dgrad(tm + kColWork, a_h1, wbase + kWCol1, 32, 64);              // dCin = dH1 W1c
    });
    {
      float dc[32];
      tmem_ld32(trow + kColWork, dc);
      // d(sigma-net output): col 0 from the density (truncated_exp backward, activation.py:21), 1..15 = d geo
      float dh0 = 0.f;
      if (valid && args.d_sigma != nullptr)
        dh0 = args.d_sigma[i] * f.density_scale * __expf(fminf(fmaxf(h0, -15.f), 15.f));
      *reinterpret_cast<uint4*>(sm.gs + 0 * kAChunk + roff) = make_uint4(
          pack_bf16x2(dh0, dc[16]), pack_bf16x2(dc[17], dc[18]), pack_bf16x2(dc[19], dc[20]), pack_bf16x2(dc[21], dc[22]));
      *reinterpret_cast<uint4*>(sm.gs + 1 * kAChunk + roff) = make_uint4(
          pack_bf16x2(dc[23], dc[24]), pack_bf16x2(dc[25], dc[26]), pack_bf16x2(dc[27], dc[28]), pack_bf16x2(dc[29], dc[30]));
    }
    mma_round(sm, phase, [&] {
      issue_wgrad(tm + kColW2s, a_hs, a_gs, 16, have_acc);                   // dW2s^T += Hs^T dOs
      issue_dgrad(tm + kColWork, a_gs, wbase + kWSig2, 64, 16);              // dHs = dOs W2s
    });
    tmem_ld64(trow + kColWork, v);
    relu_backward_inplace64(sm.hs, roff, v);
    mma_round(sm, phase, [&] {
      issue_wgrad(tm + kColW1s, a_hs, a_feat, 32, have_acc);                 // dW1s += dHs^T feat
      issue_dgrad(tm + kColWork, a_hs, wbase + kWSig1, 32, 64);              // dFeat = dHs W1s
    });
    have_acc = true;
    {
      float df[32];
      tmem_ld32(trow + kColWork, df);

      float2* st = reinterpret_cast<float2*>(sm.h1);  // h1 is dead: stage d(features) as [level][row]
#pragma unroll
      for (int l = 0; l < 16; ++l) st[l * kTile + tid] = make_float2(df[2 * l], df[2 * l + 1]);
    }
    scatter_sample(f, s, args.seg_grads, reinterpret_cast<const float2*>(sm.h1), tid);
  }

  if (!weights_ready) mbar_wait(&sm.bar_w, 0);
  // ---------------- flush the weight-gradient accumulators ----------------
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (have_acc && args.d_mlp != nullptr) {
    const int lane = tid & 31, warp = tid >> 5;
    const int m = warp * 16 + lane;  // M=64 accumulators: row m lives in lane m%16 of sub-partition m/16
    float acc[64];
    tmem_ld32(trow + kColW1s, acc);
    if (lane < 16)
      for (int c = 0; c < 32; ++c) atomicAdd(args.d_mlp + kGSig1 + m * 32 + c, acc[c]);
    tmem_ld16(trow + kColW2s, acc);
    if (lane < 16)
      for (int c = 0; c < 16; ++c) atomicAdd(args.d_mlp + kGSig2 + c * 64 + m, acc[c]);
    tmem_ld32(trow + kColW1c, acc);
    if (lane < 16)
      for (int c = 0; c < 32; ++c) atomicAdd(args.d_mlp + kGCol1 + m * 32 + c, acc[c]);
    tmem_ld64(trow + kColW2c, acc);
    if (lane < 16)
      for (int c = 0; c < 64; ++c) atomicAdd(args.d_mlp + kGCol2 + m * 64 + c, acc[c]);
    tmem_ld16(trow + kColW3c, acc);
    if (lane < 16)
      for (int c = 0; c < 16; ++c) atomicAdd(args.d_mlp + kGCol3 + c * 64 + m, acc[c]);
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(sm.tmem_base, kTmemCols);
}

}  // namespace hrf

using namespace hrf;

extern "C" int hrf_field_backward(const hrf_field* f, const hrf_samples* s, const hrf_segment_grads* seg_grads,
                                  const float* d_sigma, const float* d_rgb, const void* feat_bf16, float* d_mlp,
                                  void* stream) {
  HRF_REQUIRE(f != nullptr && s != nullptr && seg_grads != nullptr, "null argument");
  if (s->num_samples == 0) return 0;
  HRF_REQUIRE(d_sigma != nullptr || d_rgb != nullptr, "no upstream gradient given");
  if (s->ray_origins == nullptr) {
    HRF_REQUIRE(s->positions && s->frame_numbers, "query form needs positions and frame numbers");
    HRF_REQUIRE(d_rgb == nullptr || s->directions, "radiance gradients need directions");
  }
  if (s->num_samples == 0) return 0;
  BwdArgs a;
  a.f = *f;
  a.s = *s;
  a.seg_grads = seg_grads;
  a.d_sigma = d_sigma;
  a.d_rgb = d_rgb;
  a.feat_in = reinterpret_cast<const uint4*>(feat_bf16);
  a.d_mlp = d_mlp;
  const int64_t tiles = (s->num_samples + kTile - 1) / kTile;
  const int smem = (int)sizeof(BwdSmem) + 1024;
  const int64_t max_ctas = (int64_t)sm_count() * 2;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  HRF_CUDA(cudaFuncSetAttribute(field_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  field_backward_kernel<<<grid, kTile, smem, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  HRF_CHECK_LAUNCH();
  return 0;
}
