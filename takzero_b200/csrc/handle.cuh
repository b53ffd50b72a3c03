// handle.cuh -- the opaque tz_handle behind the C ABI.
#pragma once
#include <vector>

#include "../../include/takzero_b200.h"
#include "common.cuh"

struct NnState;  // nn.cu

struct tz_handle {
    int device = 0;
    TzDev d;
    cudaStream_t stream = nullptr;
    std::vector<void*> allocs;
    // agent
    int agent_kind = 0;
    tz_agent_fn agent_fn = nullptr;
    void* agent_ctx = nullptr;
    NnState* nn = nullptr;
    // device staging of host-facing arguments / results
    float* betas = nullptr;
    float* gumbel = nullptr;
    int gumbel_stride = 0;
    uint16_t* moves = nullptr;
    int* sym = nullptr;
    int* adj = nullptr;
    uint8_t* mask = nullptr;
    int* terminal = nullptr;
    unsigned long long* randoms = nullptr;
    TzState* fin_start = nullptr;
    uint16_t* fin_replay = nullptr;
    int* fin_len = nullptr;
    int* tbl_n = nullptr;
    uint16_t* tbl_moves = nullptr;
    uint32_t *tbl_u32a = nullptr, *tbl_u32b = nullptr, *tbl_u32c = nullptr;
    float *tbl_f32a = nullptr, *tbl_f32b = nullptr, *tbl_f32c = nullptr;
    uint32_t* root_stats = nullptr;
    float* ube = nullptr;
    // pinned host staging
    unsigned char* pin_small = nullptr;
    TzState* pin_states = nullptr;
    uint16_t* pin_actions = nullptr;
    int* pin_nact = nullptr;
    float* pin_logits = nullptr;
    float* pin_value = nullptr;
    float* pin_variance = nullptr;
    // bookkeeping
    unsigned long long move_counter = 0;
    unsigned long long opening_counter = 0;
    unsigned long long launches = 0;  // kernels launched by this handle
};
