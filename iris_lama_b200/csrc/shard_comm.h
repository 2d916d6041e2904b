// shard_comm.h -- the NCCL side of a sharded PFSlam2D: one communicator per rank, a stream of its own, an all-gather of the match
// results and the point-to-point exchange of migrating particle maps.  NCCL is loaded at run time (dlopen "libnccl.so.2"): the
// library has no link-time dependency on it and single-GPU users never need it.
//
// The reference has no distributed layer (its only parallelism is the thread pool of src/pf_slam2d.cpp:254-266,292-302); SURVEY 8(e)
// derives this one from the data flow of PFSlam2D::update: particles interact only through normalize / resample (:274-287,:511-574).
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace lama_b200 {

struct ShardComm;

// rank 0 creates the id every rank needs to connect (ncclGetUniqueId: 128 bytes)
int shard_unique_id(uint8_t id[128], std::string& err);
ShardComm* shard_comm_create(const uint8_t id[128], int rank, int world, int device, std::string& err);
void shard_comm_destroy(ShardComm* c);
void* shard_stream(ShardComm* c);   // cudaStream_t the collectives run on
int shard_world(const ShardComm* c);
int shard_rank(const ShardComm* c);
const std::string& shard_error(const ShardComm* c);
uint64_t shard_collectives(const ShardComm* c);

// all-gather of `bytes` per rank between device buffers, enqueued on the communicator's stream
int shard_allgather(ShardComm* c, const void* d_send, void* d_recv, size_t bytes);
// one grouped exchange: every (peer, device pointer, bytes) of `sends` goes out, every one of `recvs` comes in
struct ShardXfer { int peer; void* dptr; size_t bytes; };
int shard_exchange(ShardComm* c, const std::vector<ShardXfer>& sends, const std::vector<ShardXfer>& recvs);
int shard_sync(ShardComm* c);       // host waits for the communicator's stream

}  // namespace lama_b200
