// Micro-benchmark of the per-SM resources gram2_kernel shares: tcgen05.ld bandwidth, the issue-to-completion rate of the Gram
// MMA shape (M=128, N, K=16, both operands the same K-major SWIZZLE_128B tile), and the two (plus converter-like shared-memory
// stores) running at the same time.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I gplasdi_b200/csrc
// -I include tools/ubench/tmem_mma.cu -o build/ubench_tmem_mma.   Output: clocks per operation, one line per experiment.
#include "common.cuh"
#include "decode_gram.cuh"
#include <cstdio>
using namespace glasdi;
using namespace glasdi::gr;

struct Args { int ld_warps, ld_iters, ld_x16, mma_iters, mma_n, sts_warps, sts_iters, tf32; long long* out; };

__global__ void __launch_bounds__(1024, 1) ub_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  long long t0 = clock64(), t1 = t0;
  if (warp == 0) {
    if (a.mma_iters > 0) {
      const uint32_t idesc = a.tf32 ? umma_idesc_tf32_m128((uint32_t)a.mma_n) : umma_idesc_f16_m128((uint32_t)a.mma_n);
      const uint64_t desc0 = umma_desc_kmajor(smem_u32(smem), 16u, 1024u) | (2ull << 61);
      if (elect_one()) {
        for (int it = 0; it < a.mma_iters; ++it) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t dk = desc0 + (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4) + (uint64_t)((it & 1) * 2048);
            if (a.tf32) tc_mma_tf32(tb + (it & 1) * 128, dk, dk, idesc, k != 0);
            else tc_mma_f16(tb + (it & 1) * 128, dk, dk, idesc, k != 0);
          }
        }
        tc_commit(&bar);
      }
      __syncwarp();
      mbar_wait(&bar, 0);
      t1 = clock64();
    }
  } else if (warp >= 4 && warp < 4 + a.ld_warps) {
    const uint32_t addr = tb + ((uint32_t)((warp & 3) * 32) << 16) + 256 + ((warp >> 2) & 3) * 32;
    uint32_t acc = 0;
    if (a.ld_x16) {
      uint32_t r[16];
      for (int it = 0; it < a.ld_iters; ++it) {
        tmem_ld_32x16(addr, r); tmem_ld_wait_bind(r); acc += r[3];
        tmem_ld_32x16(addr + 16, r); tmem_ld_wait_bind(r); acc += r[5];
      }
    } else {
      uint32_t r[32];
      for (int it = 0; it < a.ld_iters; ++it) {
        tmem_ld_32x32(addr, r); tmem_ld_wait(); acc += r[3] + r[17];
      }
    }
    t1 = clock64();
    if (acc == 0x12345) a.out[100] = acc;
  } else if (warp >= 20 && warp < 20 + a.sts_warps) {
    uint8_t* dst = smem + 65536 + (warp - 20) * 2048 + lane * 16;
    for (int it = 0; it < a.sts_iters; ++it) {
#pragma unroll
      for (int m = 0; m < 4; ++m)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(dst + m * 512)), "r"(it), "r"(m), "r"(lane), "r"(warp) : "memory");
    }
    t1 = clock64();
  }
  if (blockIdx.x == 0 && lane == 0) a.out[warp] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 1024);
  const size_t smem = 65536 + 12 * 2048 + 1024;
  cudaFuncSetAttribute(ub_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  struct Exp { const char* name; Args a; };
  const int LI = 2000, MI = 500, SI = 4000;
  Exp exps[] = {
    {"ld x32, 4 warps", {4, LI, 0, 0, 112, 0, 0, 0, out}},
    {"ld x32, 8 warps", {8, LI, 0, 0, 112, 0, 0, 0, out}},
    {"ld x32, 16 warps", {16, LI, 0, 0, 112, 0, 0, 0, out}},
    {"ld x16, 16 warps", {16, LI, 1, 0, 112, 0, 0, 0, out}},
    {"mma f16 N=112", {0, 0, 0, MI, 112, 0, 0, 0, out}},
    {"mma f16 N=128", {0, 0, 0, MI, 128, 0, 0, 0, out}},
    {"mma f16 N=64", {0, 0, 0, MI, 64, 0, 0, 0, out}},
    {"mma f16 N=256", {0, 0, 0, MI, 256, 0, 0, 0, out}},
    {"mma tf32 N=128", {0, 0, 0, MI, 128, 0, 0, 1, out}},
    {"mma f16 N=112 + ld x32 16 warps", {16, LI, 0, MI, 112, 0, 0, 0, out}},
    {"mma f16 N=112 + sts 12 warps", {0, 0, 0, MI, 112, 12, SI, 0, out}},
    {"mma f16 N=112 + ld 16 + sts 12", {16, LI, 0, MI, 112, 12, SI, 0, out}},
    {"sts 12 warps", {0, 0, 0, 0, 112, 12, SI, 0, out}},
  };
  for (auto& e : exps) {
    cudaMemset(out, 0, 1024);
    ub_kernel<<<148, 1024, smem>>>(e.a);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[32]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-36s err=%d", e.name, (int)err);
    if (e.a.mma_iters) printf("  clk/MMA %.1f", (double)h[0] / (e.a.mma_iters * 8));
    if (e.a.ld_warps) { long long mx = 0; for (int w = 4; w < 4 + e.a.ld_warps; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("  ld: %.1f clk per 4KB/warp, %.1f B/clk/SM", (double)mx / e.a.ld_iters, 4096.0 * e.a.ld_warps * e.a.ld_iters / mx); }
    if (e.a.sts_warps) { long long mx = 0; for (int w = 20; w < 20 + e.a.sts_warps; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("  sts: %.1f B/clk/SM", 2048.0 * e.a.sts_warps * e.a.sts_iters / mx); }
    printf("\n");
  }
  return 0;
}
