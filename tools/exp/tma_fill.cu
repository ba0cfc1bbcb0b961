// Experiment: how fast can cp.async.bulk (TMA) zero-fill the two dense observation arrays from a shared zero page,
// and what do sparse 16-byte "ones" granules cost afterwards?  nvcc -arch=sm_100a -O3 -o tma_fill tma_fill.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store(void* g, const void* s, unsigned bytes) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(s);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }

// persistent: CTA c fills tiles c, c + grid, ...; a tile = tile_bytes contiguous; ops of op_bytes from the zero page;
// `issuers` threads (one per warp) issue in parallel; mode 1: after the tile's fill completes, every issuer's warp
// writes `ones` 16-byte granules (strided) into it
__global__ void fill_kernel(uint8_t* dst, size_t total, int tile_bytes, int op_bytes, int ones, int mode) {
  extern __shared__ __align__(128) uint8_t zero[];
  for (int i = threadIdx.x * 16; i < op_bytes; i += blockDim.x * 16) *reinterpret_cast<uint4*>(zero + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const size_t ntiles = total / tile_bytes;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    uint8_t* base = dst + t * (size_t)tile_bytes;
    const int nops = tile_bytes / op_bytes;
    if (lane == 0) {
      for (int o = warp; o < nops; o += nw) bulk_store(base + (size_t)o * op_bytes, zero, op_bytes);
      bulk_commit();
    }
    if (mode == 1) {
      if (lane == 0) bulk_wait_all();
      __syncthreads();  // every issuer's ops are complete
      // sparse granules: `ones` per 7000-byte env, spread
      const int envs = tile_bytes / 7000;
      for (int i = threadIdx.x; i < envs * ones; i += blockDim.x) {
        const int e = i / ones, k = i - e * ones;
        uint4* g = reinterpret_cast<uint4*>(base + (size_t)e * 7000 / 16 * 16 + (size_t)((k * 197 + e * 13) % 430) * 16);
        __stcs(g, make_uint4(0x3f800000u, 0, 0, 0));
      }
    }
  }
  if (lane == 0) bulk_wait_read_all();
}

template <int D> __device__ __forceinline__ void bulk_wait_pending() { asm volatile("cp.async.bulk.wait_group %0;\n" ::"n"(D) : "memory"); }
template <int D> __device__ __forceinline__ void bulk_wait_read_pending() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(D) : "memory"); }

// the candidate design: envs of 7000 bytes = node_features 5600 (zero page + 7 one-granules after the fill of that env
// completed, pipelined D envs deep per warp) + action_mask 1400 (per-warp image of a PAIR of envs = 2800 bytes with
// ones, one bulk store, wait .read before reuse).  Layout here: [nf of all envs | mask of all envs].
template <int D>
__global__ void pipe_kernel(uint8_t* dst, int nenv, int with_mask, int with_ones) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* zero = sm;                      // 5600
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  uint8_t* img = sm + 5632 + warp * 2816;  // 2800 per warp
  for (int i = threadIdx.x * 16; i < 5632 + nw * 2816; i += blockDim.x * 16) *reinterpret_cast<uint4*>(sm + i) = make_uint4(0, 0, 0, 0);
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  uint8_t* nf = dst;
  uint8_t* mask = dst + (size_t)nenv * 5600;
  const int gw = blockIdx.x * nw + warp, tw = gridDim.x * nw;
  int it = 0;
  for (int e = gw * 2; e < nenv; e += tw * 2, ++it) {  // a warp takes pairs of envs
    if (lane == 0) {
      bulk_store(nf + (size_t)e * 5600, zero, 5600);
      bulk_store(nf + (size_t)(e + 1) * 5600, zero, 5600);
      bulk_commit();
    }
    if (with_mask) {
      // image of the pair's masks: ~56 ones
      if (lane == 0 && it > 0) bulk_wait_read_pending<D>();  // conservative: the image group is older than the last D fill groups
      __syncwarp();
      for (int k = lane; k < 56; k += 32) img[(k * 53 + e) % 2800] = 1;
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        bulk_store(mask + (size_t)e * 1400, img, 2800);
        bulk_commit();
        bulk_wait_read_pending<0>();
      }
      __syncwarp();
      for (int k = lane; k < 56; k += 32) img[(k * 53 + e) % 2800] = 0;
      __syncwarp();
    }
    if (with_ones && it >= D) {
      if (lane == 0) bulk_wait_pending<D * 1>();  // fills of pair (it - D) are complete (mask groups in between only make this stricter)
      __syncwarp();
      const int eo = e - D * tw * 2;
      if (lane < 14) {
        uint4* g = reinterpret_cast<uint4*>(nf + (size_t)(eo + lane / 7) * 5600 + (size_t)(((lane % 7) * 47 + eo) % 350) * 16);
        __stcs(g, make_uint4(0x3f800000u, 0, 0, 0));
      }
    }
  }
  if (lane == 0) bulk_wait_pending<0>();
}

// reference: plain STG.128 streaming zero fill
__global__ void stg_fill(uint4* dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) __stcs(dst + i, make_uint4(0, 0, 0, 0));
}

int main() {
  const size_t total = (size_t)65536 * 7000;  // action_mask + node_features of config 3
  uint8_t* d;
  CK(cudaMalloc(&d, total + 4096));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  CK(cudaFuncSetAttribute(fill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  auto run = [&](const char* name, auto launch) {
    for (int i = 0; i < 3; ++i) launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) launch();
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 20;
    printf("%-60s %8.2f us  %7.1f GB/s\n", name, ms * 1e3, total / ms / 1e6);
  };
  run("cudaMemsetAsync", [&] { cudaMemsetAsync(d, 0, total, 0); });
  run("STG.128 .cs grid 148*8 x 512", [&] { stg_fill<<<148 * 8, 512>>>((uint4*)d, total / 16); });
  const int tile = 32 * 7000;  // 224000 bytes per 32-env tile
  char buf[128];
  for (int op : {2000, 8000, 32000}) {  // divisors of 224000 / multiples of 16: 224000 = 2^6 * 3500
    if (tile % op) continue;
    for (int ctas : {1, 2, 4}) {
      for (int warps : {1, 4}) {
        snprintf(buf, sizeof buf, "TMA fill op=%5d B  %d CTA/SM  %d issuer warps", op, ctas, warps);
        run(buf, [&] { fill_kernel<<<148 * ctas, warps * 32, op>>>(d, total, tile, op, 0, 0); });
      }
    }
  }
  for (int ones : {8, 35}) {
    snprintf(buf, sizeof buf, "TMA fill op=8000 2 CTA/SM 4 warps + %d granules/env after completion", ones);
    run(buf, [&] { fill_kernel<<<148 * 2, 128, 8000>>>(d, total, tile, 8000, ones, 1); });
    snprintf(buf, sizeof buf, "TMA fill op=8000 4 CTA/SM 4 warps + %d granules/env after completion", ones);
    run(buf, [&] { fill_kernel<<<148 * 4, 128, 8000>>>(d, total, tile, 8000, ones, 1); });
  }
  CK(cudaFuncSetAttribute(pipe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(pipe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  CK(cudaFuncSetAttribute(pipe_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
  for (int warps : {2, 4, 8}) {
    for (int ctas : {1, 2}) {
      const int smem = 5632 + warps * 2816;
      for (int cfg = 0; cfg < 4; ++cfg) {
        const int wm = cfg & 1, wo = cfg >> 1;
        snprintf(buf, sizeof buf, "pipe D=2 %d warps %d CTA/SM mask=%d ones=%d", warps, ctas, wm, wo);
        run(buf, [&] { pipe_kernel<2><<<148 * ctas, warps * 32, smem>>>(d, 65536, wm, wo); });
      }
      snprintf(buf, sizeof buf, "pipe D=1 %d warps %d CTA/SM mask=1 ones=1", warps, ctas);
      run(buf, [&] { pipe_kernel<1><<<148 * ctas, warps * 32, smem>>>(d, 65536, 1, 1); });
      snprintf(buf, sizeof buf, "pipe D=4 %d warps %d CTA/SM mask=1 ones=1", warps, ctas);
      run(buf, [&] { pipe_kernel<4><<<148 * ctas, warps * 32, smem>>>(d, 65536, 1, 1); });
    }
  }
  return 0;
}
