// comm.cuh -- NCCL over NVLink / NVSwitch for the two exchanges the path has (comm.cu): the weight broadcast of a
// generation and the sums of the counters.  NCCL is bound at run time (dlopen), so the library loads without it.
#pragma once
#include <stddef.h>

#include "handle.cuh"

int comm_unique_id(void* out128);
int comm_init(tz_handle* h, const void* id128, int nranks, int rank);
void comm_destroy(tz_handle* h);
int comm_nranks(const tz_handle* h);  // 1 without a communicator
int comm_rank(const tz_handle* h);
int comm_broadcast(tz_handle* h, void* dev_buf, size_t bytes, int root, cudaStream_t st);
int comm_allreduce_sum_u64(tz_handle* h, unsigned long long* dev_buf, int count, cudaStream_t st);
const char* comm_last_error();
