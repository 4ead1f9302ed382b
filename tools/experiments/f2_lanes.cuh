// f2_lanes.cuh -- launch interface of the lane-stream kernel (f2_lanes.cu): the decimated-output
// hot path with warp-uniform coefficients (lane = stream, warp = channel).
#pragma once
#include "f2_common.cuh"
#include "f2_edge.cuh"

#ifndef F2_LANE_WARPS
#define F2_LANE_WARPS 4
#endif

namespace f2 {

constexpr int kLaneWarps = F2_LANE_WARPS;  // channels per CTA
constexpr int kMaxConstChan = 1024;        // channels that fit the __constant__ parameter block

// One lane's work: output samples [t0, t1) of utterance `utt` (t0 a multiple of 32).
struct LaneStream {
    int utt;
    int t0;
    int t1;
    int pad;
};

// Per group of 32 streams: warm-up schedule in tiles of 32 samples.  Tiles [0, mA) run the
// cascade only, [mA, mB) add magnitude + low-pass, [mB, ...) produce output.
struct LaneGroup {
    int mA;
    int mB;
};

struct LaneParams {
    const UttDesc* utts;
    const LaneStream* streams;
    const LaneGroup* groups;
    int n_streams;
    const float2* xz;
    const float* G;
    const float* edge;  // [utt][C][8]: injection coefficients for even t (0..3) and odd t (4..7)
    float* dec;         // decimated envelope frames [frame][C]
    int C;
    int step;
    int phase;
    int lpf;
    float lp_k;
    float lp_b0;
};

cudaError_t upload_lane_constants(const float* host_par, int c_pad, cudaStream_t stream);
cudaError_t launch_lanes(const LaneParams& p, int n_groups, cudaStream_t stream);
int lane_tile_samples();

}  // namespace f2
