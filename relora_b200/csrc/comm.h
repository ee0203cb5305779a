// NVLink peer-memory collectives over symmetric buffers (implemented in comm.cu).
//
// Every rank owns one allocation of identical size (torch symmetric memory / cuMem VMM); `peers.ptr[p]` is rank p's
// allocation mapped into this process, `mc` (optional) is the NVLS multicast alias of the same allocation.
// Synchronisation between ranks uses a flag area inside a second symmetric buffer: slot[set][src] on rank dst.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rb {

constexpr int kMaxPeers = 8;

struct PeerPtrs {
  void* ptr[kMaxPeers];
};

struct CommCtx {
  int rank = 0, world = 1;
  PeerPtrs flags;        // uint32 flag area of every rank (>= 4 * kMaxPeers words) + float norm slots
  uint32_t* local_go;    // device word in local (non-symmetric) memory: intra-kernel broadcast of "barrier passed"
};

// Cross-GPU barrier: every rank calls it with the same monotonically increasing `epoch` (>0).
void xgpu_barrier(const CommCtx& c, int set, uint32_t epoch, cudaStream_t s);

// In-place all-reduce (sum) of `n` bf16 elements living at byte offset `off` of the symmetric buffers.
// two-shot: rank r reduces chunk r over all peers (fp32 accumulate) and writes the result to every peer.
// `mc` != nullptr uses multimem.ld_reduce / multimem.st (in-switch reduction + broadcast).
void allreduce_bf16(const CommCtx& c, const PeerPtrs& bufs, void* mc, long long off_elems, long long n, uint32_t epoch0, int max_blocks,
                    cudaStream_t s);

// Fused data-parallel update on flat buffers (ZeRO-1 dataflow, result identical on every rank):
//   1. grads_f32 (local) -> bf16 into the symmetric gradient buffer              [cast]
//   2. barrier; rank r reduces chunk r over peers (fp32), writes `gred` (fp32, local, chunk-sized), accumulates sum(g^2)
//   3. sum(g^2) partials exchanged through the flag area -> clip coefficient     [norm]
//   4. AdamW on chunk r (moments are chunk-sized, local) and broadcast of the updated bf16 parameters to every peer
//   5. barrier
struct FusedUpdateArgs {
  const float* grads_f32;   // local [n]; nullptr: the bf16 gradients are already in grad_bufs[rank]
  PeerPtrs grad_bufs;       // symmetric bf16 [n]
  void* grad_mc;            // multicast alias or nullptr
  float* gred;              // local fp32 [n / world]
  PeerPtrs param_bufs;      // symmetric bf16 [n] (the flat parameter buffer of every rank)
  void* param_mc;
  void* exp_avg;            // local bf16 [n / world]
  void* exp_avg_sq;
  long long n;              // multiple of 8 * world
  float lr, beta1, beta2, eps, weight_decay;
  int step;
  const float* step_dev;    // device-resident step count (nullable): replaces `step` in the bias corrections
  float max_norm;           // <= 0: no clipping
  float inv_world;          // gradients are averaged over ranks
  const float* skip;        // device flag (nullable): != 0 on ANY rank -> no update on every rank (combined in the norm exchange)
  const float* loss_in;     // this rank's mean loss (nullable); the mean over ranks lands in loss_out  [replaces the NCCL all-reduce
  float* loss_out;          //   of the reference's loss_info, torchrun_main.py:810]
  float* norm_out;          // device float: total gradient norm (averaged gradient)
  float* sq_accum;          // device float[3] scratch: sum(g^2) (zeroed by the call), grad scale, combined skip flag
  int max_blocks;
};
void fused_update(const CommCtx& c, const FusedUpdateArgs& a, uint32_t epoch0, cudaStream_t s);

}  // namespace rb
