// Ad-hoc probe (not product code): minimal TMA tiled load with zero OOB fill, to find what the blur H-pass
// tile load needs on sm_100a. Usage: tma_probe <rank 2|3> <box_w> <xs> <proxy_fence 0|1>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
constexpr int kRows = 16, kMaxBox = 192;
template <int RANK>
__global__ void k(const __grid_constant__ CUtensorMap tmap, int box_w, int xs, int ys, int z, int fence_kind, float* out) {
  __shared__ __align__(128) float s[kRows * kMaxBox];
  __shared__ __align__(8) unsigned long long mbar;
  const int tid = threadIdx.x;
  const unsigned mbar_addr = static_cast<unsigned>(__cvta_generic_to_shared(&mbar));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
    if (fence_kind) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    else asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned bytes = static_cast<unsigned>(box_w) * kRows * sizeof(float);
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(mbar_addr), "r"(xs), "r"(ys) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                   ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(mbar_addr), "r"(xs), "r"(ys), "r"(z) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(mbar_addr) : "memory");
  for (int i = tid; i < box_w * kRows; i += blockDim.x) out[i] = s[i];
}
int main(int argc, char** argv) {
  const int rank = argc > 1 ? atoi(argv[1]) : 2, box_w = argc > 2 ? atoi(argv[2]) : 32, xs = argc > 3 ? atoi(argv[3]) : 0;
  const int fence_kind = argc > 4 ? atoi(argv[4]) : 0;
  const int W = argc > 5 ? atoi(argv[5]) : 96, H = 80, P = (W + 31) / 32 * 32, planes = 3, ys = 72, z = 1;
  std::vector<float> h(static_cast<size_t>(P) * H * planes);
  for (size_t i = 0; i < h.size(); ++i) h[i] = static_cast<float>(i % 1000) + 1.0f;
  float *d = nullptr, *o = nullptr;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, kRows * kMaxBox * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no entry point\n"); return 2; }
  CUtensorMap tm;
  const cuuint64_t gdim[3] = {W, H, planes};
  const cuuint64_t gstr[2] = {P * 4ull, static_cast<cuuint64_t>(P) * H * 4ull};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_w), kRows, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = reinterpret_cast<EncodeFn>(fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d + (rank == 2 ? static_cast<size_t>(z) * P * H : 0), gdim, gstr, box, es,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("W %d encode rank %d box %d xs %d fence %d -> %d\n", W, rank, box_w, xs, fence_kind, static_cast<int>(r));
  if (r != CUDA_SUCCESS) return 3;
  if (rank == 2) k<2><<<1, 256>>>(tm, box_w, xs, ys, z, fence_kind, o); else k<3><<<1, 256>>>(tm, box_w, xs, ys, z, fence_kind, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  if (e != cudaSuccess) return 4;
  std::vector<float> got(static_cast<size_t>(box_w) * kRows);
  cudaMemcpy(got.data(), o, got.size() * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < kRows; ++y)
    for (int x = 0; x < box_w; ++x) {
      const int gx = xs + x, gy = ys + y;
      const float want = (gx >= 0 && gx < W && gy >= 0 && gy < H) ? h[static_cast<size_t>(z) * P * H + static_cast<size_t>(gy) * P + gx] : 0.0f;
      bad += got[static_cast<size_t>(y) * box_w + x] != want;
    }
  printf("mismatches: %d\n", bad);
  return bad ? 5 : 0;
}
