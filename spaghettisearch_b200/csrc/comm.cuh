// Cross-GPU exchange used by the sharded paths; no-ops for
// world == 1.  All calls enqueue on the engine stream.
#pragma once
#include <stddef.h>

#include <cuda_runtime.h>

#include "spaghetti.h"

int comm_rank(const ss_engine* e);
int comm_world(const ss_engine* e);
// in-place sum over ranks of a small device vector (normaliser / residual)
int comm_allreduce_sum_f64(ss_engine* e, double* dev_buf, size_t count);
// every rank r contributes dev_buf[byte_off[r] .. +byte_cnt[r]) of a buffer that
// has the same layout on all ranks (rank blocks of the PageRank state)
int comm_allgatherv_bytes(ss_engine* e, void* dev_buf, const size_t* byte_off, const size_t* byte_cnt);
// every rank contributes `bytes` bytes; out receives world * bytes (host buffers, small payloads:
// staged through device memory and NCCL)
int comm_allgather_host_bytes(ss_engine* e, const void* in, size_t bytes, void* out);
// n equally sized device blocks per rank: out[i] receives world * bytes[i] (rank-major), one NCCL group
int comm_allgather_dev(ss_engine* e, int n, const void* const* in, void* const* out, const size_t* bytes);
// the same two on an explicit stream (the PageRank sweep loop issues its NCCL calls on an exchange stream)
int comm_allreduce_sum_f64_on(ss_engine* e, cudaStream_t st, double* dev_buf, size_t count);
int comm_allgatherv_bytes_on(ss_engine* e, cudaStream_t st, void* dev_buf, const size_t* byte_off, const size_t* byte_cnt);
// in-place sum over ranks of a device vector of 32-bit counters (in-degree histogram of a sharded load)
int comm_allreduce_sum_u32(ss_engine* e, uint32_t* dev_buf, size_t count);
// n parallel all-to-alls with the same layout: rank r receives send[i][send_off[r] .. +send_cnt[r]) of every
// rank into recv[i][recv_off[src] .. +recv_cnt[src]) (edge exchange of a sharded load)
int comm_alltoallv_u32(ss_engine* e, int n, const uint32_t* const* send, const size_t* send_off, const size_t* send_cnt,
                       uint32_t* const* recv, const size_t* recv_off, const size_t* recv_cnt);
