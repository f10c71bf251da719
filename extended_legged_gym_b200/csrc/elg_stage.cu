// elg_stage.cu -- SM-issued block copy between pinned (mapped) host memory and device memory: a measurement, not the product path.
//
// The end-to-end step moves ONE packed block of simulator state in (5.26 MB at 4096 anymal_c envs) and one packed block of
// observations / rewards / reset flags out.  Question: does a kernel that issues the transfer from the SMs -- thousands of 16-byte
// (mode 0) or 16 KB TMA (mode 1) reads outstanding over PCIe at once -- beat one copy engine's cudaMemcpyAsync?  Measured on B200
// (scripts/stage_probe.py, profiles/README.md r2j): copy engine 52.7 GB/s alone / 42.7 GB/s while the copy-out runs; this kernel
// 49.9 GB/s at every grid size, both modes, and 36-39 GB/s next to the copy-out.  No: the link sets the rate, bench.py keeps the
// copy engine (and pipelines the copy-in under the previous step instead).  Kept as a probe.  No reference counterpart.
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

__device__ __forceinline__ uint4 ld_stream16(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// mode 0: grid-stride loop of 16-byte loads, kUnroll independent loads per thread in flight
template <int kUnroll>
__global__ void __launch_bounds__(256) elg_stage_ldst_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, const int64_t n16) {
  pdl_launch_dependents();
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (kUnroll - 1) * stride < n16; i += kUnroll * stride) {
    uint4 v[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) v[k] = ld_stream16(src + i + k * stride);
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) dst[i + k * stride] = v[k];
  }
  for (; i < n16; i += stride) dst[i] = ld_stream16(src + i);
}

// mode 1: every CTA walks over `tile`-byte pieces: cp.async.bulk global -> shared (mbarrier), cp.async.bulk shared -> global;
// kStages pieces in flight per CTA
constexpr int kStageDepth = 4;
__global__ void __launch_bounds__(32) elg_stage_bulk_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, const int64_t bytes, const int tile) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar[kStageDepth];
  if (threadIdx.x == 0)
    for (int s = 0; s < kStageDepth; ++s) mbar_init(&bar[s], 1);
  pdl_launch_dependents();
  __syncwarp();
  pdl_wait();
  if (threadIdx.x != 0) return;
  const int64_t ntiles = (bytes + tile - 1) / tile;
  auto tile_bytes = [&](int64_t t) { return (uint32_t)min((int64_t)tile, bytes - t * tile); };
  // prologue: the first kStageDepth pieces of this CTA
  int64_t t_load = blockIdx.x;
  for (int s = 0; s < kStageDepth && t_load < ntiles; ++s, t_load += gridDim.x) {
    mbar_expect_tx(&bar[s], tile_bytes(t_load));
    bulk_g2s(smem_raw + (size_t)s * tile, src + t_load * tile, tile_bytes(t_load), &bar[s]);
  }
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % kStageDepth;
    mbar_wait(&bar[s], (uint32_t)(it / kStageDepth) & 1u);
    bulk_s2g(dst + t * tile, smem_raw + (size_t)s * tile, tile_bytes(t));
    bulk_commit();
    if (t_load < ntiles) {
      bulk_wait_read_all();   // the slot has been read out before it is overwritten
      mbar_expect_tx(&bar[s], tile_bytes(t_load));
      bulk_g2s(smem_raw + (size_t)s * tile, src + t_load * tile, tile_bytes(t_load), &bar[s]);
      t_load += gridDim.x;
    }
  }
  bulk_wait_read_all();
}

}  // namespace elg

extern "C" {

// dst / src: 16-byte aligned, one of them typically pinned host memory (cudaHostAlloc / torch pin_memory: mapped under unified
// addressing); bytes: multiple of 16.  mode 0: load / store kernel, mode 1: TMA bulk pieces through shared memory.
// grid <= 0: default (4 CTAs per SM for mode 0, one per SM for mode 1).
int elg_stage_block(void* dst, const void* src, int64_t bytes, int mode, int grid, void* stream) {
  if (bytes == 0) return ELG_OK;
  if (!dst || !src) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_stage_block: dst/src is NULL");
  if (bytes < 0 || (bytes & 15) || (reinterpret_cast<uintptr_t>(dst) & 15) || (reinterpret_cast<uintptr_t>(src) & 15))
    return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_stage_block: pointers and byte count must be multiples of 16");
  const int sms = elg::sm_count();
  if (sms <= 0) return elg::set_error(ELG_ERR_CUDA, "elg_stage_block: no device");
  cudaLaunchConfig_t cfg{};
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (mode == 1) {
    const int tile = 16 * 1024;
    static elg::SmemCache smem_cache = {};
    size_t& have = elg::smem_slot(smem_cache);
    if (have == 0) {
      if (cudaFuncSetAttribute(elg::elg_stage_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, elg::kStageDepth * tile) != cudaSuccess)
        return elg::set_error(ELG_ERR_CUDA, "elg_stage_block: cannot reserve shared memory");
      have = (size_t)elg::kStageDepth * tile;
    }
    const int64_t ntiles = (bytes + tile - 1) / tile;
    int g = grid > 0 ? grid : sms;
    if (g > ntiles) g = (int)ntiles;
    cfg.gridDim = dim3((unsigned)g);
    cfg.blockDim = dim3(32u);
    cfg.dynamicSmemBytes = (size_t)elg::kStageDepth * tile;
    cudaLaunchKernelEx(&cfg, elg::elg_stage_bulk_kernel, (uint8_t*)dst, (const uint8_t*)src, bytes, tile);
  } else {
    const int64_t n16 = bytes / 16;
    int g = grid > 0 ? grid : 4 * sms;
    const int64_t need = (n16 + 255) / 256;
    if (g > need) g = (int)need;
    cfg.gridDim = dim3((unsigned)g);
    cfg.blockDim = dim3(256u);
    cudaLaunchKernelEx(&cfg, elg::elg_stage_ldst_kernel<8>, (uint4*)dst, (const uint4*)src, n16);
  }
  return elg::check_launch("elg_stage_block");
}

}  // extern "C"
