// Peer-mapped device memory for the bank-sharded matching: one process per GPU, each exports one allocation
// through CUDA IPC and maps the other ranks' allocations, so that kernels can store into a peer's HBM over NVLink.
// The 64-byte handles travel over whatever the host already has (torch.distributed all_gather_object in aocb200/shard.py).
#include <string.h>

#include "common.cuh"

using namespace aoc;

extern "C" int aoc_peer_alloc(size_t bytes, void** ptr_out) {
    AOC_CHECK_ARG(bytes > 0 && ptr_out, "bad args");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) {
        set_error("aoc_peer_alloc: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return AOC_ELAUNCH;
    }
    *ptr_out = p;
    return AOC_OK;
}

extern "C" int aoc_peer_free(void* ptr) {
    if (ptr && cudaFree(ptr) != cudaSuccess) {
        cudaGetLastError();
        set_error("aoc_peer_free: cudaFree failed");
        return AOC_ELAUNCH;
    }
    return AOC_OK;
}

// handle_out: AOC_PEER_HANDLE_BYTES (64) host bytes
extern "C" int aoc_peer_export(void* ptr, void* handle_out) {
    AOC_CHECK_ARG(ptr && handle_out, "bad args");
    static_assert(sizeof(cudaIpcMemHandle_t) == AOC_PEER_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        set_error("aoc_peer_export: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return AOC_ELAUNCH;
    }
    memcpy(handle_out, &h, sizeof(h));
    return AOC_OK;
}

// maps another process's allocation into this one (peer access is enabled lazily by the driver)
extern "C" int aoc_peer_open(const void* handle, void** ptr_out) {
    AOC_CHECK_ARG(handle && ptr_out, "bad args");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        set_error("aoc_peer_open: %s", cudaGetErrorString(e));
        cudaGetLastError();
        return AOC_ELAUNCH;
    }
    *ptr_out = p;
    return AOC_OK;
}

extern "C" int aoc_peer_close(void* ptr) {
    if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) {
        cudaGetLastError();
        set_error("aoc_peer_close: cudaIpcCloseMemHandle failed");
        return AOC_ELAUNCH;
    }
    return AOC_OK;
}
