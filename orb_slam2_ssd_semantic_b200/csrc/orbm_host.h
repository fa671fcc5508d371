// orbm_host.h -- matcher handle: a stream + a grow-only device arena for the host-buffer entry points.
#pragma once
#include "orbm_match.cuh"

struct orbm {
  int device = 0;
  cudaStream_t stream = nullptr;
  long long launches = 0;
  void* d_arena = nullptr;
  size_t arena_bytes = 0;
  ~orbm();
  int reserve(size_t bytes);
};

// shared with the stream pipeline (orbs.cu)
int launch_match_last(const b200::MatchBatch& mb, const b200::MatchCam& cam, int npairs, int cmax, cudaStream_t stream);
