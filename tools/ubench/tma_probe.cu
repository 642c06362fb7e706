// TMA probe: which way of handing a CUtensorMap to cp.async.bulk.tensor works on this driver / toolkit.
//   ./tma_probe A   descriptor = direct __grid_constant__ kernel parameter
//   ./tma_probe B   descriptor = element of an array inside a __grid_constant__ struct, runtime index
//   ./tma_probe C   descriptor in global memory
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

struct Maps { CUtensorMap lv[8]; };

__device__ __forceinline__ void tile_load(const void* map, uint8_t* dst, uint64_t* bar, int x, int y, int f, int bytes) {
  const uint32_t b = (uint32_t)__cvta_generic_to_shared(bar), d = (uint32_t)__cvta_generic_to_shared(dst);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(d), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(f), "r"(b) : "memory");
  }
  __syncwarp();
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(b) : "memory");
  } while (!ok);
}

__global__ void kA(const __grid_constant__ CUtensorMap m, uint8_t* out, int x, int y, int f, int nb) {
  extern __shared__ __align__(128) uint8_t sm[];
  tile_load(&m, sm + 128, reinterpret_cast<uint64_t*>(sm), x, y, f, nb);
  for (int i = threadIdx.x; i < nb; i += 32) out[i] = sm[128 + i];
}
__global__ void kB(const __grid_constant__ Maps m, int l, uint8_t* out, int x, int y, int f, int nb) {
  extern __shared__ __align__(128) uint8_t sm[];
  tile_load(&m.lv[l], sm + 128, reinterpret_cast<uint64_t*>(sm), x, y, f, nb);
  for (int i = threadIdx.x; i < nb; i += 32) out[i] = sm[128 + i];
}
__global__ void kC(const CUtensorMap* m, int l, uint8_t* out, int x, int y, int f, int nb) {
  extern __shared__ __align__(128) uint8_t sm[];
  tile_load(m + l, sm + 128, reinterpret_cast<uint64_t*>(sm), x, y, f, nb);
  for (int i = threadIdx.x; i < nb; i += 32) out[i] = sm[128 + i];
}

typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const char mode = argc > 1 ? argv[1][0] : 'A';
  const int x = argc > 2 ? atoi(argv[2]) : 123, y = argc > 3 ? atoi(argv[3]) : 457;
  const int BW = argc > 4 ? atoi(argv[4]) : 48, BH = argc > 5 ? atoi(argv[5]) : 44;
  const int l2 = argc > 6 ? atoi(argv[6]) : 2;
  const int W = 640, H = 480, F = 3, pitch = 640;
  std::vector<uint8_t> h((size_t)pitch * H * F);
  for (size_t i = 0; i < h.size(); i++) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *out;
  cudaMalloc(&d, h.size());
  cudaMalloc(&out, 256 * 256);
  cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  printf("entry point: %s q=%d p=%p\n", cudaGetErrorString(e), (int)q, p);
  Maps M;
  memset(&M, 0, sizeof(M));
  const cuuint64_t dims[3] = {W, H, F};
  const cuuint64_t strides[2] = {pitch, (cuuint64_t)pitch * H};
  const cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}, es[3] = {1, 1, 1};
  for (int l = 0; l < 8; l++) {
    CUresult r = ((Enc)p)(&M.lv[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)l2,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) printf("encode %d failed %d\n", l, (int)r);
  }
  const int f = 2, l = 5;
  const int nb = BW * BH;
  if (mode == 'A') kA<<<1, 32, 128 + nb>>>(M.lv[l], out, x, y, f, nb);
  if (mode == 'B') kB<<<1, 32, 128 + nb>>>(M, l, out, x, y, f, nb);
  if (mode == 'C') {
    CUtensorMap* dm;
    cudaMalloc(&dm, sizeof(M));
    cudaMemcpy(dm, &M, sizeof(M), cudaMemcpyHostToDevice);
    kC<<<1, 32, 128 + nb>>>(dm, l, out, x, y, f, nb);
  }
  e = cudaDeviceSynchronize();
  printf("mode %c x=%d y=%d box=%dx%d l2=%d: %s\n", mode, x, y, BW, BH, l2, cudaGetErrorString(e));
  if (e != cudaSuccess) return 1;
  std::vector<uint8_t> o(nb);
  cudaMemcpy(o.data(), out, o.size(), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < BH; r++)
    for (int c = 0; c < BW; c++) {
      const int yy = y + r, xx = x + c;
      const uint8_t want = (yy < H && xx < W) ? h[(size_t)f * pitch * H + (size_t)yy * pitch + xx] : 0;
      bad += o[r * BW + c] != want;
    }
  printf("mode %c: mismatches %d (rows past the image bottom must read 0)\n", mode, bad);
  return bad != 0;
}
