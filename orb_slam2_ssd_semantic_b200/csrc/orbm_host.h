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

// shared with the stream pipeline (orbs.cu): grid build + candidate lists + order-exact resolve of `npairs` pairs
int launch_match_last(const b200::CurView& cv, const b200::LastView& lv, const b200::MatchCam& cam, int npairs, int* d_goff,
                      int* d_gidx, unsigned* d_list, int* d_count, int* d_accepted, int* d_cur2last, int* d_nmatch,
                      int cmax, int lmax, cudaStream_t stream, long long* launches);
