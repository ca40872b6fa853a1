// Rollout kernel instantiations for the planar quadrotor tracking a time-varying reference (HJB_CTL_TRACK);
// see rollout_kernel.cuh and systems.cuh::TrackCtl.
#include "rollout_kernel.cuh"

namespace hjb {
#define SYS_QUAD2D(F) Quad2DSys<F>
HJB_DEFINE_PROBLEM(quad2d_track, SYS_QUAD2D, TrackCtl, false)
}  // namespace hjb
