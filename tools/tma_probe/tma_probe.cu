// Development probe: which way of handing a tensor map / box shape to cp.async.bulk.tensor works on this GPU.
// usage: tma_probe <variant>   (one variant per process: a faulting variant kills the context)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda/barrier>

struct Maps { CUtensorMap m[8]; };

template <int RANK>
__device__ void tma_load(unsigned dst, const void* desc, unsigned bar, int x, int y, int z) {
  if (RANK == 3)
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(desc)), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
  else
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(reinterpret_cast<unsigned long long>(desc)), "r"(x), "r"(y), "r"(bar) : "memory");
}

template <int RANK>
__device__ void body(const void* desc, int bytes, int x, int y, int z, unsigned* out, int nout) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar_s;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_s), dst = (unsigned)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    tma_load<RANK>(dst, desc, bar, x, y, z);
  }
  __syncwarp();
  unsigned ok = 0;
  for (int s = 0; s < (1 << 22) && !ok; ++s)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
  if (threadIdx.x == 0) out[0] = ok;
  for (int i = threadIdx.x; i < nout; i += 32) out[1 + i] = reinterpret_cast<unsigned*>(sm)[i];
}

__global__ void k_single3(const __grid_constant__ CUtensorMap m, int bytes, int x, int y, int z, unsigned* out, int nout) { body<3>(&m, bytes, x, y, z, out, nout); }
__global__ void k_single2(const __grid_constant__ CUtensorMap m, int bytes, int x, int y, int z, unsigned* out, int nout) { body<2>(&m, bytes, x, y, z, out, nout); }
__global__ void k_array3(const __grid_constant__ Maps m, int l, int bytes, int x, int y, int z, unsigned* out, int nout) { body<3>(&m.m[l], bytes, x, y, z, out, nout); }
__global__ void k_global3(const CUtensorMap* m, int bytes, int x, int y, int z, unsigned* out, int nout) { body<3>(m, bytes, x, y, z, out, nout); }

// 1-D bulk copy (no tensor map): rows of `rb` bytes, one cp.async.bulk per row
__global__ void k_bulk1d(const unsigned char* src, int pitch, int rb, int rows, unsigned* out, int nout) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar_s;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_s), dst = (unsigned)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rb * rows) : "memory");
  }
  __syncwarp();
  for (int r = threadIdx.x; r < rows; r += 32)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + r * rb),
                 "l"(src + (size_t)r * pitch), "r"(rb), "r"(bar) : "memory");
  unsigned ok = 0;
  for (int s = 0; s < (1 << 22) && !ok; ++s)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar) : "memory");
  if (threadIdx.x == 0) out[0] = ok;
  for (int i = threadIdx.x; i < nout; i += 32) out[1 + i] = reinterpret_cast<unsigned*>(sm)[i];
}

// libcu++ reference: cuda::barrier + cp_async_bulk_tensor_2d
namespace cde = cuda::device::experimental;
__global__ void k_libcu2(const __grid_constant__ CUtensorMap m, int bytes, int x, int y, unsigned* out, int nout) {
  extern __shared__ __align__(128) unsigned char sm[];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ cuda::barrier<cuda::thread_scope_block> bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  cuda::barrier<cuda::thread_scope_block>::arrival_token tok;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(sm, &m, x, y, bar);
    tok = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else tok = bar.arrive();
  bar.wait(std::move(tok));
  if (threadIdx.x == 0) out[0] = 1;
  for (int i = threadIdx.x; i < nout; i += 32) out[1 + i] = reinterpret_cast<unsigned*>(sm)[i];
}

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int W = 640, H = 480, F = 3, bw = (variant & 8) ? 128 : 96, bh = (variant & 16) ? 32 : 39;
  std::vector<unsigned char> img((size_t)W * H * F);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (unsigned char)(i * 7 + (i >> 9));
  unsigned char* d; cudaMalloc(&d, img.size()); cudaMemcpy(d, img.data(), img.size(), cudaMemcpyHostToDevice);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("no encoder\n"); return 2; }
  auto enc = (PFN_cuTensorMapEncodeTiled)fn;
  Maps maps; memset(&maps, 0, sizeof(maps));
  const int rank = (((variant & 3) == 1) || (variant & 64)) ? 2 : 3;
  const int es_bytes = (variant & 256) ? 4 : 1;   // element type: u8 or u32 (x is in elements)
  cuuint64_t dims[3] = {(cuuint64_t)(W / es_bytes), H, F}, strides[2] = {W, (cuuint64_t)W * H};
  cuuint32_t box[3] = {(cuuint32_t)(bw / es_bytes), (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
  for (int l = 0; l < 8; ++l) {
    CUresult r = enc(&maps.m[l], (variant & 256) ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 3; }
  }
  if (variant & 128) {
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(&maps.m[0]);
    printf("ptr %p desc:", (void*)d);
    for (int i = 0; i < 16; ++i) printf(" %016llx", q[i]);
    printf("\n");
  }
  const int bytes = bw * bh, nout = bytes / 4, x = argc > 2 ? atoi(argv[2]) : 36, y = argc > 3 ? atoi(argv[3]) : 100, z = (rank == 3) ? 1 : 0;
  unsigned* out; cudaMalloc(&out, 4 * (nout + 1)); cudaMemset(out, 0xee, 4 * (nout + 1));
  CUtensorMap* dm; cudaMalloc(&dm, sizeof(CUtensorMap)); cudaMemcpy(dm, &maps.m[0], sizeof(CUtensorMap), cudaMemcpyHostToDevice);
  if (variant & 32) {
    k_bulk1d<<<1, 32, bytes>>>(d + (size_t)z * W * H + (size_t)y * W + 32, W, bw, bh, out, nout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d (1-D bulk): %s\n", variant, cudaGetErrorString(e)); return 1; }
    std::vector<unsigned> h(nout + 1); cudaMemcpy(h.data(), out, 4 * (nout + 1), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; ++r) for (int c = 0; c < bw; ++c)
      bad += reinterpret_cast<unsigned char*>(h.data() + 1)[r * bw + c] != img[(size_t)z * W * H + (size_t)(y + r) * W + (32 + c)];
    printf("variant %d (1-D bulk rows %dx%d): landed %u, %d wrong bytes\n", variant, bw, bh, h[0], bad);
    return 0;
  }
  if (variant & 64) { k_libcu2<<<1, 32, bytes>>>(maps.m[0], bytes, x, y, out, nout); goto done; }
  switch (variant & 3) {
    case 0: k_single3<<<1, 32, bytes>>>(maps.m[0], bytes, x, y, z, out, nout); break;
    case 1: k_single2<<<1, 32, bytes>>>(maps.m[0], bytes, x, y, z, out, nout); break;
    case 2: k_array3<<<1, 32, bytes>>>(maps, (variant & 4) ? 5 : 0, bytes, x, y, z, out, nout); break;
    case 3: k_global3<<<1, 32, bytes>>>(dm, bytes, x, y, z, out, nout); break;
  }
done:
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
  std::vector<unsigned> h(nout + 1); cudaMemcpy(h.data(), out, 4 * (nout + 1), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int r = 0; r < bh; ++r)
    for (int c = 0; c < bw; ++c) {
      const unsigned char got = reinterpret_cast<unsigned char*>(h.data() + 1)[r * bw + c];
      const unsigned char want = img[(size_t)z * W * H + (size_t)(y + r) * W + (x * es_bytes + c)];
      bad += got != want;
    }
  printf("variant %d (rank %d, box %dx%d): landed %u, %d wrong bytes\n", variant, rank, bw, bh, h[0], bad);
  return 0;
}
