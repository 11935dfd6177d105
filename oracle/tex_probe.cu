// TEST INFRASTRUCTURE ONLY: samples a real CUDA 3-D texture configured exactly like the reference's
// occupancy texture (actorshq/dataset/native/occupancy_grid.cu:17-38: uint8 channel, clamp
// addressing, linear filter, normalised-float read mode, normalised coordinates) so the bit-packed
// integer emulation in csrc/sampler.cu and oracle/sampler.py can be pinned against the hardware.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

__global__ void probe_kernel(cudaTextureObject_t tex, const float* p, int64_t n, float* out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = tex3D<float>(tex, p[3 * i], p[3 * i + 1], p[3 * i + 2]);
}

extern "C" int tex_probe(const uint8_t* grid_dev, int G, const float* points_dev, int64_t n, float* out_dev) {
  cudaChannelFormatDesc cd = cudaCreateChannelDesc(8, 0, 0, 0, cudaChannelFormatKindUnsigned);
  cudaArray_t arr = nullptr;
  cudaExtent ext = make_cudaExtent(G, G, G);
  if (cudaMalloc3DArray(&arr, &cd, ext) != cudaSuccess) return 1;
  cudaMemcpy3DParms cp;
  memset(&cp, 0, sizeof(cp));
  cp.srcPtr = make_cudaPitchedPtr((void*)grid_dev, G, G, G);
  cp.dstArray = arr;
  cp.extent = ext;
  cp.kind = cudaMemcpyDefault;
  if (cudaMemcpy3D(&cp) != cudaSuccess) return 2;
  cudaResourceDesc rd;
  memset(&rd, 0, sizeof(rd));
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = arr;
  cudaTextureDesc td;
  memset(&td, 0, sizeof(td));
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModeLinear;
  td.readMode = cudaReadModeNormalizedFloat;
  td.normalizedCoords = 1;
  cudaTextureObject_t tex = 0;
  if (cudaCreateTextureObject(&tex, &rd, &td, nullptr) != cudaSuccess) return 3;
  probe_kernel<<<(unsigned)((n + 255) / 256), 256>>>(tex, points_dev, n, out_dev);
  cudaError_t e = cudaDeviceSynchronize();
  cudaDestroyTextureObject(tex);
  cudaFreeArray(arr);
  return e == cudaSuccess ? 0 : 4;
}
