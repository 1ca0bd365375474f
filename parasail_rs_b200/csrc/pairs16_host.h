// pairs16_host.h -- launch interface of the packed 16-bit many-pairs kernels (pairs16.cu), kept in its own
// translation unit so that its ~50 template instantiations compile beside engine.cu, not after it.
#pragma once
#include <cuda_runtime.h>

#include <string>

#include "kern_pairs16.cuh"

namespace psb {

struct P16Class { int G, K; };
int p16_num_classes();
P16Class p16_class(int idx);
// smallest class whose G*K rows hold a query of lq residues; prefer_wide chooses the 32-lane groups (half
// the profile bytes per warp, two pairs per warp) when both group widths could hold it.  -1: none.
int p16_pick_class(int lq, bool prefer_wide);
// grid = min(work, resident capacity) persistent warps; returns PSB_OK or a PSB_E* code with *err filled
int p16_launch(int cls, bool sw, bool trace, const Pairs16Params &p, int sms, cudaStream_t stream, std::string *err,
               int *warps_per_sm_out = nullptr);
int p16_launch_walk(const Walk16Params &w, bool stats, cudaStream_t stream, std::string *err);

}  // namespace psb
