// match_kernels.h — launch interface of the Hamming best-2 kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/eorb_b200.h"

namespace eorb {

int hamming_chunks(long long ndb, int sms, long long* chunkRows);
int hamming_query_tiles(int nq);
cudaError_t launch_hamming_best2(const uint8_t* d_q, int nq, const uint8_t* d_db, long long ndb, long long indexOffset,
                                 long long chunkRows, int nchunks, eorb_best2* d_partial, cudaStream_t st);
cudaError_t launch_merge_best2(const eorb_best2* d_parts, int nparts, int nq, eorb_best2* d_merged, eorb_match* d_out,
                               int th, float ratio, cudaStream_t st);
// tensor-core engine (hamming_tc.cu): same partial format, hamming_tc_parts_per_chunk() partial arrays per chunk
int hamming_tc_chunks(long long ndb, int nq, int sms, long long* chunkRows);
int hamming_tc_parts_per_chunk();
cudaError_t launch_hamming_best2_tc(const uint8_t* d_q, int nq, const uint8_t* d_db, long long ndb, long long indexOffset, long long chunkRows,
                                    int nchunks, eorb_best2* d_partial, cudaStream_t st);
cudaError_t launch_popc_probe(unsigned* d_out, int blocks, int iters, cudaStream_t st);

}  // namespace eorb
