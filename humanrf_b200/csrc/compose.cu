// Stand-alone tensor composition  out = xyz*v_t + xyt*v_z + yzt*v_x + xzt*v_y  on half features (the op the reference
// exposes as tensor_composition_native, humanrf/scene_representation/native/tensor_composition.cu:120-219; semantics of
// its kernels :30-54,85-117).  The fused field kernels do this internally; this entry exists for code written against
// the reference extension (decomposition4d.py:8-39) and as a differential target for the reference build.
//
// Layout of the work (written for this op's access pattern, not the reference's thread-per-element mapping):
//   * a thread owns ONE feature pair (half2 / float2 accesses) and walks kWalk consecutive samples; the 16 lanes of a
//     half-warp cover the 32 features of a sample, so every feature row is read / written as one 64-byte segment;
//   * the four lerp taps of a sample are computed once per thread and sample (they do not depend on the feature);
//   * backward: consecutive samples of a batch are consecutive steps along a ray, 4e-4 apart, i.e. they hit the same two
//     rows of `vectors` for many samples (always, on the time axis).  The vector gradient is therefore accumulated in
//     registers while the tap pair stays the same and flushed with one red.global.add.v2.f32 per row and run -- instead
//     of two fp32 atomics per (sample, feature, axis), which serialise on the handful of time rows a batch touches.
#include "common.cuh"

namespace hrf {

constexpr int kCmpLanes = 16, kCmpRows = 16, kCmpWalk = 8;   // block = 16 x 16 threads, 128 samples

struct Tap4 {
  int o0[4], o1[4];   // row offsets (elements) of the two taps of each axis
  float fr[4];
};
__device__ __forceinline__ Tap4 make_taps(const float* __restrict__ coords, int64_t si, int VR, int F) {
  Tap4 t;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float c = coords[si * 4 + i] * VR - 0.5f;    // tensor_composition.cu:37-45
    const float fl = floorf(c);
    t.fr[i] = c - fl;
    const int c0 = (int)fmaxf(fl, 0.f), c1 = (int)fminf(fl + 1.f, (float)(VR - 1));
    t.o0[i] = (i * VR + c0) * F, t.o1[i] = (i * VR + c1) * F;
  }
  return t;
}
// two neighbouring features of a row; `pair` tells whether feature f+1 exists (odd feature counts)
__device__ __forceinline__ float2 ld_h2(const __half* p, int64_t e, bool vec, bool pair) {
  if (vec) return __half22float2(*reinterpret_cast<const __half2*>(p + e));
  return make_float2(__half2float(p[e]), pair ? __half2float(p[e + 1]) : 0.f);
}
__device__ __forceinline__ void st_h2(__half* p, int64_t e, float2 v, bool vec, bool pair) {
  if (vec) {
    *reinterpret_cast<__half2*>(p + e) = __floats2half2_rn(v.x, v.y);
  } else {
    p[e] = __float2half(v.x);
    if (pair) p[e + 1] = __float2half(v.y);
  }
}
__device__ __forceinline__ float2 ld_f2(const float* p, int64_t e, bool vec, bool pair) {
  if (vec) return *reinterpret_cast<const float2*>(p + e);
  return make_float2(p[e], pair ? p[e + 1] : 0.f);
}
__device__ __forceinline__ void red_f2(float* p, int64_t e, float2 v, bool vec, bool pair) {
  if (v.x == 0.f && v.y == 0.f) return;
  if (vec) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + e), "f"(v.x), "f"(v.y) : "memory");
  } else {
    atomicAdd(p + e, v.x);
    if (pair) atomicAdd(p + e + 1, v.y);
  }
}

template <bool kBackward>
__global__ void __launch_bounds__(kCmpLanes* kCmpRows) compose_kernel(
    const __half* __restrict__ xyz, const __half* __restrict__ xyt, const __half* __restrict__ yzt,
    const __half* __restrict__ xzt, const float* __restrict__ vec, const float* __restrict__ coords,
    const __half* __restrict__ dout, int64_t n, int F, int VR, __half* __restrict__ o_xyz /* out | d_xyz */,
    __half* __restrict__ o_xyt, __half* __restrict__ o_yzt, __half* __restrict__ o_xzt, float* __restrict__ dvec) {
  const int lane = threadIdx.x, rowt = threadIdx.y;
  const int64_t s0 = ((int64_t)blockIdx.x * kCmpRows + rowt) * kCmpWalk;
  const bool vec2 = (F & 1) == 0;   // rows start on half2 / float2 boundaries
  for (int f = 2 * lane; f < F; f += 2 * kCmpLanes) {
    const bool pair = f + 1 < F;
    // run state of the vector gradient, per axis: the current tap pair and what has been accumulated for it
    int ro0[4] = {-1, -1, -1, -1}, ro1[4] = {-1, -1, -1, -1};
    float2 a0[4], a1[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a0[i] = a1[i] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int w = 0; w < kCmpWalk; ++w) {
      const int64_t si = s0 + w;
      if (si >= n) break;
      const Tap4 t = make_taps(coords, si, VR, F);
      const int64_t e = si * F + f;
      float2 sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 v0 = ld_f2(vec, t.o0[i] + f, vec2, pair), v1 = ld_f2(vec, t.o1[i] + f, vec2, pair);
        sv[i] = make_float2(v0.x + t.fr[i] * (v1.x - v0.x), v0.y + t.fr[i] * (v1.y - v0.y));
      }
      // pairing: xyz <-> v[3] = t, xyt <-> v[2] = z, yzt <-> v[0] = x, xzt <-> v[1] = y   (tensor_composition.cu:49-52)
      const float2 fx = ld_h2(xyz, e, vec2, pair), fy = ld_h2(xyt, e, vec2, pair), fz = ld_h2(yzt, e, vec2, pair),
                   fw = ld_h2(xzt, e, vec2, pair);
      if (!kBackward) {
        st_h2(o_xyz, e, make_float2(fx.x * sv[3].x + fy.x * sv[2].x + fz.x * sv[0].x + fw.x * sv[1].x,
                                    fx.y * sv[3].y + fy.y * sv[2].y + fz.y * sv[0].y + fw.y * sv[1].y), vec2, pair);
      } else {
        const float2 d = ld_h2(dout, e, vec2, pair);
        st_h2(o_xyz, e, make_float2(sv[3].x * d.x, sv[3].y * d.y), vec2, pair);
        st_h2(o_xyt, e, make_float2(sv[2].x * d.x, sv[2].y * d.y), vec2, pair);
        st_h2(o_yzt, e, make_float2(sv[0].x * d.x, sv[0].y * d.y), vec2, pair);
        st_h2(o_xzt, e, make_float2(sv[1].x * d.x, sv[1].y * d.y), vec2, pair);
        const float2 fe[4] = {fz, fw, fy, fx};   // the feature that multiplies vector axis i
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (t.o0[i] != ro0[i] || t.o1[i] != ro1[i]) {   // new tap pair on this axis: flush the run
            if (ro0[i] >= 0) {
              red_f2(dvec, ro0[i] + f, a0[i], vec2, pair);
              red_f2(dvec, ro1[i] + f, a1[i], vec2, pair);
            }
            a0[i] = a1[i] = make_float2(0.f, 0.f);
            ro0[i] = t.o0[i], ro1[i] = t.o1[i];
          }
          const float gx = fe[i].x * d.x, gy = fe[i].y * d.y;     // :109-111  dV[i0] += g (1-frac), dV[i1] += g frac
          a0[i].x += gx * (1.f - t.fr[i]), a0[i].y += gy * (1.f - t.fr[i]);
          a1[i].x += gx * t.fr[i], a1[i].y += gy * t.fr[i];
        }
      }
    }
    if (kBackward) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (ro0[i] >= 0) {
          red_f2(dvec, ro0[i] + f, a0[i], vec2, pair);
          red_f2(dvec, ro1[i] + f, a1[i], vec2, pair);
        }
    }
  }
}

}  // namespace hrf

using namespace hrf;

extern "C" int hrf_compose_tensors_forward(const void* xyz, const void* xyt, const void* yzt, const void* xzt,
                                           const float* vectors, const float* coords, int64_t n, int feature_dim,
                                           int vec_res, void* out, void* stream) {
  HRF_REQUIRE(feature_dim >= 1 && vec_res >= 1, "bad feature / vector size");
  if (n == 0) return 0;
  const int64_t per_block = (int64_t)kCmpRows * kCmpWalk;
  compose_kernel<false><<<(unsigned)((n + per_block - 1) / per_block), dim3(kCmpLanes, kCmpRows), 0,
                          reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __half*)xyz, (const __half*)xyt, (const __half*)yzt, (const __half*)xzt, vectors, coords, nullptr, n, feature_dim,
      vec_res, (__half*)out, nullptr, nullptr, nullptr, nullptr);
  HRF_CHECK_LAUNCH();
  return 0;
}

extern "C" int hrf_compose_tensors_backward(const void* xyz, const void* xyt, const void* yzt, const void* xzt,
                                            const float* vectors, const float* coords, const void* d_out, int64_t n,
                                            int feature_dim, int vec_res, void* d_xyz, void* d_xyt, void* d_yzt,
                                            void* d_xzt, float* d_vectors, void* stream) {
  HRF_REQUIRE(feature_dim >= 1 && vec_res >= 1, "bad feature / vector size");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  HRF_CUDA(cudaMemsetAsync(d_vectors, 0, sizeof(float) * 4 * (size_t)vec_res * feature_dim, st));  // :188 zeros_like
  if (n == 0) return 0;
  const int64_t per_block = (int64_t)kCmpRows * kCmpWalk;
  compose_kernel<true><<<(unsigned)((n + per_block - 1) / per_block), dim3(kCmpLanes, kCmpRows), 0, st>>>(
      (const __half*)xyz, (const __half*)xyt, (const __half*)yzt, (const __half*)xzt, vectors, coords, (const __half*)d_out, n,
      feature_dim, vec_res, (__half*)d_xyz, (__half*)d_xyt, (__half*)d_yzt, (__half*)d_xzt, d_vectors);
  HRF_CHECK_LAUNCH();
  return 0;
}
