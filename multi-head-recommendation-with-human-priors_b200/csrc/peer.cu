// Peer access to the row-sharded item table inside one NVLink / NVSwitch node.
// Every rank exports its table shard and its per-step gradient-row buffer with CUDA IPC; the peers map them once and
// from then on the lookup / gradient exchange is plain 16-byte loads over NVLink inside our own kernels:
//   gather_rows_sharded   out[i] = shard[id % W][id / W]          (replaces ids all-to-all + gather + rows all-to-all)
//   scatter_add_sorted_peer (embed.cu) the owner reads the gradient rows of its ids straight out of every rank's buffer
//                         while it reduces them (replaces the gradient-row all-to-all + local reduction)
// No host synchronisation, no data-dependent sizes: both run on fixed-capacity id lists (ids <= 0 = no row).
#include <cuda.h>

#include "common.cuh"

static CUresult (*p_cuMemGetAddressRange)(CUdeviceptr*, size_t*, CUdeviceptr) = nullptr;

static int load_driver_fn() {
  if (p_cuMemGetAddressRange) return 0;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  B200_CUDA_OK(cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &q));
  B200_CHECK_ARG(q == cudaDriverEntryPointSuccess && f != nullptr, "cuMemGetAddressRange not available");
  p_cuMemGetAddressRange = (CUresult(*)(CUdeviceptr*, size_t*, CUdeviceptr))f;
  return 0;
}

// handle_out: 64 bytes (cudaIpcMemHandle_t) of the cudaMalloc allocation that contains ptr; offset_out: ptr - base.
extern "C" int b200rec_ipc_export(const void* ptr, void* handle_out, int64_t* offset_out) {
  if (load_driver_fn()) return 1;
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = p_cuMemGetAddressRange(&base, &size, (CUdeviceptr)ptr);
  B200_CHECK_ARG(r == CUDA_SUCCESS, "ipc_export: cuMemGetAddressRange failed (%d)", (int)r);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
  if (e != cudaSuccess) {
    b200rec_set_error("ipc_export: cudaIpcGetMemHandle -> %s (the buffer must come from cudaMalloc: disable "
                      "PYTORCH_CUDA_ALLOC_CONF=expandable_segments)", cudaGetErrorString(e));
    cudaGetLastError();
    return 2;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  memcpy(handle_out, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)ptr - base);
  return 0;
}

// Maps a peer's allocation into this process (peer access is enabled lazily by the driver); returns its base.
extern "C" int b200rec_ipc_import(const void* handle, void** base_out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    b200rec_set_error("ipc_import: cudaIpcOpenMemHandle -> %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 2;
  }
  *base_out = p;
  return 0;
}

extern "C" int b200rec_ipc_close(void* base) {
  B200_CUDA_OK(cudaIpcCloseMemHandle(base));
  return 0;
}

// out[i, :] = shard[id % W][id / W, :]   (ids < 0: zero row).  16-byte vectors, 4 rows in flight per thread: remote
// rows arrive over NVLink with ~2 us latency, so bytes in flight per SM, not issue rate, set the speed.
template <int U>
__global__ void __launch_bounds__(256) gather_rows_sharded_kernel(const float* const* __restrict__ shards, int W,
                                                                  int D4, const int64_t* __restrict__ ids,
                                                                  int64_t n_vec, float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i < n_vec; i += U * stride) {
    float v[U][4];
    int64_t idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = i + u * stride;
      v[u][0] = v[u][1] = v[u][2] = v[u][3] = 0.f;
      if (idx[u] < n_vec) {
        const int64_t r = idx[u] / D4;
        const int c = (int)(idx[u] - r * D4);
        const int64_t id = __ldg(ids + r);
        if (id >= 0) {
          const float* base = shards[(int)(id % W)];
          load4<float>(base + ((id / W) * D4 + c) * 4, v[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (idx[u] < n_vec) store4<float>(out + idx[u] * 4, v[u]);
  }
}

extern "C" int b200rec_gather_rows_sharded(const void* shard_ptrs_dev, int W, int D, const int64_t* ids, int64_t n_ids,
                                           float* out, void* stream) {
  B200_CHECK_ARG(D % 4 == 0 && W >= 1, "gather_rows_sharded: bad D / W");
  if (n_ids == 0) return 0;
  const int64_t n_vec = n_ids * (D / 4);
  const int blocks = (int)std::min<int64_t>((n_vec + 256 * 8 - 1) / (256 * 8), 148 * 8);
  gather_rows_sharded_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float* const*)shard_ptrs_dev, W, D / 4,
                                                                          ids, n_vec, out);
  B200_LAUNCH_OK();
  return 0;
}
