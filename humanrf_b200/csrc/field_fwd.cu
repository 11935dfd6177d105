// Fused forward of the HumanRF radiance field for sm_100a:
//   sample load -> 4 hash grids x 16 levels gather -> vector composition -> sigma MLP (tcgen05)
//   -> truncated_exp density -> SH + colour MLP (tcgen05) -> sigmoid radiance.
// One persistent CTA of 128 threads owns one 128-sample tile at a time; thread r owns sample r
// (= row r of every MMA operand and TMEM lane r of every accumulator), so the whole
// encode -> MLP chain stays in registers / shared memory / TMEM.
// Reference semantics: humanrf/scene_representation/humanrf.py:158-208,
// decomposition4d.py:124-135, native/tensor_composition.cu:9-55.
#include <stdlib.h>

#include "field_common.cuh"

namespace hrf {


// One activation buffer serves every layer: a layer's output tile is written over its input tile, which is safe
// because each thread writes only after the tcgen05.commit barrier of the MMA that read the old tile.
struct __align__(128) FwdSmem {
  unsigned char w[kWBlobBytes];       // packed weights (TMA bulk copy, once per CTA)
  unsigned char a[kTile * 64 * 2];    // A tile: K = 32 layout in the first 8 KB, or K = 64 layout (hidden activations)
  uint64_t bar_w;                     // weights landed
  uint64_t bar_mma;                   // tcgen05.commit arrival
  uint32_t tmem_base;
};

template <int N>
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float* v) {
  if constexpr (N == 64) tmem_ld64(taddr, v);
  else if constexpr (N == 32) tmem_ld32(taddr, v);
  else tmem_ld16(taddr, v);
}

// One dense layer for the CTA's tile; every thread ends up with its row's N outputs in v[].
// kSimt == true is the fp32 CUDA-core debug path (tests only).
template <bool kSimt, int N, int K>
__device__ __forceinline__ void run_layer(FwdSmem& sm, const unsigned char* abuf, uint32_t woff, uint32_t& phase,
                                          float* v) {
  const int tid = threadIdx.x;
  if constexpr (!kSimt) {
    tc_fence_before();
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer(sm.tmem_base, smem_u32(abuf), smem_u32(sm.w + woff), N, K);
      umma_commit(&sm.bar_mma);
    }
    mbar_wait(&sm.bar_mma, phase);
    phase ^= 1u;
    tc_fence_after();
    const uint32_t taddr = sm.tmem_base + ((uint32_t)(tid & ~31) << 16);
    tmem_load_row<N>(taddr, v);
  } else {
    __syncthreads();
    const uint32_t roff = a_row_off(tid);
    float in[K];
#pragma unroll
    for (int k = 0; k < K; k += 2) {
      const uint32_t u = *reinterpret_cast<const uint32_t*>(abuf + (k >> 3) * kAChunk + roff + (k & 7) * 2);
      in[k] = bf16_lo(u), in[k + 1] = bf16_hi(u);
    }
#pragma unroll 4
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < K; k += 2) {
        const uint32_t u = *reinterpret_cast<const uint32_t*>(sm.w + woff + w_off(n, k, N));
        acc = __fmaf_rn(in[k], bf16_lo(u), acc);
        acc = __fmaf_rn(in[k + 1], bf16_hi(u), acc);
      }
      v[n] = acc;
    }
    __syncthreads();
  }
}

// relu -> bf16 -> K-major A tile with K = 64
__device__ __forceinline__ void store_hidden(unsigned char* hbuf, int row, const float* v) {
  const uint32_t roff = a_row_off(row);
#pragma unroll
  for (int kg = 0; kg < 8; ++kg) {
    uint4 q;
    q.x = pack_bf16x2(fmaxf(v[kg * 8 + 0], 0.f), fmaxf(v[kg * 8 + 1], 0.f));
    q.y = pack_bf16x2(fmaxf(v[kg * 8 + 2], 0.f), fmaxf(v[kg * 8 + 3], 0.f));
    q.z = pack_bf16x2(fmaxf(v[kg * 8 + 4], 0.f), fmaxf(v[kg * 8 + 5], 0.f));
    q.w = pack_bf16x2(fmaxf(v[kg * 8 + 6], 0.f), fmaxf(v[kg * 8 + 7], 0.f));
    *reinterpret_cast<uint4*>(hbuf + kg * kAChunk + roff) = q;
  }
}

template <bool kSimt, int kCtasPerSm>
__global__ void __launch_bounds__(kTile, kCtasPerSm) field_forward_kernel(const __grid_constant__ FieldArgs args) {
  extern __shared__ unsigned char smem_raw[];
  FwdSmem& sm = *reinterpret_cast<FwdSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  const int tid = threadIdx.x;
  const hrf_field& f = args.f;

  // ---- one-time CTA setup: barriers, TMEM, weights via TMA bulk copy -------------------
  if (tid == 0) {
    mbar_init(&sm.bar_w, 1);
    mbar_init(&sm.bar_mma, 1);
    fence_mbar_init();
  }
  if constexpr (!kSimt) {
    if (tid < 32) {
      tmem_alloc(&sm.tmem_base, 64);
      tmem_relinquish();
    }
    tc_fence_before();
  }
  __syncthreads();
  if constexpr (!kSimt) tc_fence_after();
  if (tid == 0) {
    mbar_arrive_expect_tx(&sm.bar_w, kWBlobBytes);
    tma_load_1d(sm.w, f.mlp_blob, kWBlobBytes, &sm.bar_w);
  }
  bool weights_ready = false;
  uint32_t phase = 0;

  const int64_t n = args.s.num_samples;
  const int64_t num_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t i = tile * kTile + tid;
    const Sample s = load_sample(f, args.s, i, args.mode != 0);
    encode_to_smem(f, s, sm.a, tid);
    if (args.feat != nullptr && i < args.s.num_samples) {
      const uint32_t ro = a_row_off(tid);
#pragma unroll
      for (int kg = 0; kg < 4; ++kg) args.feat[i * 4 + kg] = *reinterpret_cast<const uint4*>(sm.a + kg * kAChunk + ro);
    }
    if (!weights_ready) {
      mbar_wait(&sm.bar_w, 0);
      weights_ready = true;
    }

    // ---- sigma net: 32 -> 64 (ReLU) -> 16 ------------------------------------------------
    float v[64];
    run_layer<kSimt, 64, 32>(sm, sm.a, kWSig1, phase, v);
    store_hidden(sm.a, tid, v);
    float o[16];
    run_layer<kSimt, 16, 64>(sm, sm.a, kWSig2, phase, o);
    const bool valid = i < n;
    // humanrf.py:184 : density = truncated_exp(h[...,0]) * density_scale  (exp in fp32)
    const float sigma = __expf(o[0]) * f.density_scale;
    if (valid) {
      if (args.sigma != nullptr) args.sigma[i] = sigma;
      if (args.geo != nullptr) {
        uint4* gp = reinterpret_cast<uint4*>(args.geo + i * 8);
        gp[0] = make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]),
                           pack_bf16x2(o[6], o[7]));
        gp[1] = make_uint4(pack_bf16x2(o[8], o[9]), pack_bf16x2(o[10], o[11]), pack_bf16x2(o[12], o[13]),
                           pack_bf16x2(o[14], o[15]));
      }
    }
    if (args.mode == 0) continue;

    // ---- colour net: [SH16 | geo15 | 1.0] -> 64 (ReLU) -> 64 (ReLU) -> 16 -> sigmoid[:3] ----
    {
      float sh[16];
      sh4(s.dx, s.dy, s.dz, sh);
      const uint32_t roff = a_row_off(tid);
      *reinterpret_cast<uint4*>(sm.a + 0 * kAChunk + roff) = make_uint4(
          pack_bf16x2(sh[0], sh[1]), pack_bf16x2(sh[2], sh[3]), pack_bf16x2(sh[4], sh[5]), pack_bf16x2(sh[6], sh[7]));
      *reinterpret_cast<uint4*>(sm.a + 1 * kAChunk + roff) =
          make_uint4(pack_bf16x2(sh[8], sh[9]), pack_bf16x2(sh[10], sh[11]), pack_bf16x2(sh[12], sh[13]),
                     pack_bf16x2(sh[14], sh[15]));
      *reinterpret_cast<uint4*>(sm.a + 2 * kAChunk + roff) = make_uint4(
          pack_bf16x2(o[1], o[2]), pack_bf16x2(o[3], o[4]), pack_bf16x2(o[5], o[6]), pack_bf16x2(o[7], o[8]));
      *reinterpret_cast<uint4*>(sm.a + 3 * kAChunk + roff) =
          make_uint4(pack_bf16x2(o[9], o[10]), pack_bf16x2(o[11], o[12]), pack_bf16x2(o[13], o[14]),
                     pack_bf16x2(o[15], 1.0f));
    }
    run_layer<kSimt, 64, 32>(sm, sm.a, kWCol1, phase, v);
    store_hidden(sm.a, tid, v);
    run_layer<kSimt, 64, 64>(sm, sm.a, kWCol2, phase, v);
    store_hidden(sm.a, tid, v);
    run_layer<kSimt, 16, 64>(sm, sm.a, kWCol3, phase, o);
    if (valid && args.rgb != nullptr) {
      float* rp = args.rgb + 3 * i;
      rp[0] = 1.f / (1.f + __expf(-o[0]));
      rp[1] = 1.f / (1.f + __expf(-o[1]));
      rp[2] = 1.f / (1.f + __expf(-o[2]));
    }
  }

  // ---- teardown ---------------------------------------------------------------------------
  if (!weights_ready) mbar_wait(&sm.bar_w, 0);  // never leave a bulk copy in flight
  if constexpr (!kSimt) {
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(sm.tmem_base, 64);
  }
}

}  // namespace hrf

using namespace hrf;

extern "C" int hrf_field_forward(const hrf_field* f, const hrf_samples* s, int mode, int mlp_impl, float* sigma,
                                 void* geo_bf16, float* rgb, void* feat_bf16, void* stream) {
  HRF_REQUIRE(f != nullptr && s != nullptr, "null field/samples");
  HRF_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (density) or 1 (density+radiance)");
  HRF_REQUIRE(s->num_samples >= 0, "negative sample count");
  HRF_REQUIRE(f->mlp_blob != nullptr && f->segments != nullptr, "field not initialised");
  if (s->num_samples == 0) return 0;
  if (s->ray_origins != nullptr) {
    HRF_REQUIRE(s->ray_directions && s->ray_frame_numbers && s->sample_distances && s->ray_indices,
                "ray-batch form needs origins, directions, frame numbers, distances and ray indices");
  } else {
    HRF_REQUIRE(s->positions && s->frame_numbers, "query form needs positions and frame numbers");
    HRF_REQUIRE(mode == 0 || s->directions, "radiance queries need directions");
  }
  if (s->num_samples == 0) return 0;
  FieldArgs a;
  a.f = *f;
  a.s = *s;
  a.sigma = sigma;
  a.geo = reinterpret_cast<uint32_t*>(geo_bf16);
  a.rgb = rgb;
  a.feat = reinterpret_cast<uint4*>(feat_bf16);
  a.mode = mode;
  const int64_t tiles = (s->num_samples + kTile - 1) / kTile;
  const int smem = (int)sizeof(FwdSmem) + 128;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // CTAs per SM, measured on B200 on the bench batch: 4 (118 regs) 1.291 ms, 5 (96 regs) 1.265 ms, 6 (80 regs,
  // spills, less gather ILP per thread) 1.475 ms.  HRF_FWD_CTAS overrides for experiments.
  static const int ctas_per_sm = [] {
    const char* e = getenv("HRF_FWD_CTAS");
    const int v = e ? atoi(e) : 5;
    return (v == 4 || v == 6) ? v : 5;
  }();
  const int64_t max_ctas = (int64_t)sm_count() * ctas_per_sm;
  const int grid = (int)(tiles < max_ctas ? tiles : max_ctas);
  auto launch = [&](auto kernel) -> int {
    HRF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    kernel<<<grid, kTile, smem, st>>>(a);
    return 0;
  };
  int rc;
  if (mlp_impl != 0) rc = launch(field_forward_kernel<true, 4>);
  else if (ctas_per_sm == 6) rc = launch(field_forward_kernel<false, 6>);
  else if (ctas_per_sm == 5) rc = launch(field_forward_kernel<false, 5>);
  else rc = launch(field_forward_kernel<false, 4>);
  if (rc) return rc;
  HRF_CHECK_LAUNCH();
  return 0;
}
