// elg_probe.cu -- launch-floor probes for the step kernel's roofline argument (diagnostics, no reference counterpart).
//
// The fused step at 4096 envs is one 148-CTA launch whose duration is dominated by fixed latencies, not by bytes.  These two
// kernels measure the fixed part on the same box, with the same launch shape (148 CTAs x 1024 threads, ~90 KB dynamic shared
// memory, programmatic dependent launch, CUDA-graph replay) and none of the arithmetic:
//   elg_probe_empty      griddepcontrol.launch_dependents / .wait and nothing else: the launch-to-launch floor of a PDL chain
//   elg_probe_roundtrip  per CTA: one cp.async.bulk of `bytes_in` global -> shared on an mbarrier, then one cp.async.bulk of
//                        `bytes_out` shared -> global: the step's bulk data movement (its algorithmic bytes, cold or warm L2)
//                        with zero instructions in between -- what "the bytes alone" cost at this launch shape
#include <cuda_runtime.h>
#include <stdint.h>

#include "elg_common.cuh"
#include "elg_async.cuh"

namespace elg {

__global__ void __launch_bounds__(1024, 1) elg_probe_empty_kernel(int* sink) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  pdl_launch_dependents();
  pdl_wait();
  if (sink && threadIdx.x == 0 && blockIdx.x == 0x7fffffff) *sink = smem_raw[0];
}

__global__ void __launch_bounds__(1024, 1)
elg_probe_roundtrip_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, uint32_t bytes_in, uint32_t bytes_out) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  pdl_launch_dependents();
  __syncthreads();
  pdl_wait();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, bytes_in);
    bulk_g2s(smem_raw, src + (size_t)blockIdx.x * bytes_in, bytes_in, &bar);
  }
  mbar_wait(&bar, 0);
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_async_smem();
    bulk_s2g(dst + (size_t)blockIdx.x * bytes_out, smem_raw, bytes_out);
    bulk_commit();
    bulk_wait_read_all();
  }
}

}  // namespace elg

extern "C" {

int elg_probe_empty(int grid, int threads, int smem_bytes, int pdl, void* stream) {
  if (grid < 1 || threads < 32 || threads > 1024 || smem_bytes < 0 || smem_bytes > 227 * 1024)
    return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_probe_empty: bad launch shape");
  if (cudaFuncSetAttribute(elg::elg_probe_empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess)
    return elg::set_error(ELG_ERR_CUDA, "elg_probe_empty: cannot reserve shared memory");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  int* sink = nullptr;
  if (cudaLaunchKernelEx(&cfg, elg::elg_probe_empty_kernel, sink) != cudaSuccess) return elg::check_launch("elg_probe_empty");
  return elg::check_launch("elg_probe_empty");
}

int elg_probe_roundtrip(const void* src, void* dst, int64_t bytes_in_per_cta, int64_t bytes_out_per_cta, int grid, int pdl, void* stream) {
  if (!src || !dst) return elg::set_error(ELG_ERR_NULL_POINTER, "elg_probe_roundtrip: src/dst is NULL");
  const int64_t smem = bytes_in_per_cta > bytes_out_per_cta ? bytes_in_per_cta : bytes_out_per_cta;
  if (grid < 1 || bytes_in_per_cta < 16 || bytes_out_per_cta < 16 || (bytes_in_per_cta & 15) || (bytes_out_per_cta & 15) || smem > 200 * 1024)
    return elg::set_error(ELG_ERR_INVALID_ARGUMENT, "elg_probe_roundtrip: byte counts must be multiples of 16 and fit shared memory");
  if (cudaFuncSetAttribute(elg::elg_probe_roundtrip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return elg::set_error(ELG_ERR_CUDA, "elg_probe_roundtrip: cannot reserve shared memory");
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(1024u);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  if (cudaLaunchKernelEx(&cfg, elg::elg_probe_roundtrip_kernel, (const uint8_t*)src, (uint8_t*)dst, (uint32_t)bytes_in_per_cta,
                         (uint32_t)bytes_out_per_cta) != cudaSuccess)
    return elg::check_launch("elg_probe_roundtrip");
  return elg::check_launch("elg_probe_roundtrip");
}

}  // extern "C"
