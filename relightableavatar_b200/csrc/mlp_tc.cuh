// placeholder, replaced below
#pragma once
#include "common.cuh"
#include <string>
struct ra_weights;
struct TcWeights { int dummy; };
static int tc_init(TcWeights&, std::string&) { return 0; }
static void tc_free(TcWeights&) {}
static int tc_upload(TcWeights&, const ra_weights*, std::string&, cudaStream_t) { return 0; }
static void tc_set_frame(TcWeights&, const FrameConst*, cudaStream_t, int64_t&) {}
static void tc_distance(TcWeights&, const float*, float*, const int*, float, int, cudaStream_t, int64_t&) {}
