// Staging of pageable host buffers for the batch *_host entries (host_stage.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>

namespace osb {

bool host_is_pageable(const void* p);
void host_parallel_copy(void* dst, const void* src, size_t bytes);

struct StageIn {
    bool on = false;
    void* slot[2] = {nullptr, nullptr};
    int open(bool pageable, size_t max_group_bytes, int device);
    int src(int g, const void* user, size_t bytes, const cudaEvent_t* h2d_done, const void** out);
};

struct StageOut {
    bool on = false;
    void* slot[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t* ev = nullptr;  // [groups] D2H-landed events, owned by the caller
    void* user[32];
    size_t bytes[32];
    int n = 0, handed = 0;
    int open(bool pageable, size_t max_group_bytes, int device, cudaEvent_t* events);
    int dst(int g, void* user_dst, void** out);
    int done(int g, void* user_dst, size_t nbytes, cudaStream_t s_out);
    int finish();
    int hand_over(int g);
};

}  // namespace osb
