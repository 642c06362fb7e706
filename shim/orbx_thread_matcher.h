// shim/orbx_thread_matcher.h — the matcher context of the calling thread, shared by the reference-side bodies
// (ORBmatcher_orbx.cc, ORBmatcher_next_orbx.cc, ORBmatcher_sim3_orbx.cc, FrameStereo_orbx.cc, Tracking_orbx.cc,
// FrameBoW_orbx.cc).
//
// ORBmatcher objects are per-call stack temporaries used from the Tracking, LocalMapping and LoopClosing threads
// (src/Tracking.cc:2784,3303, src/LocalMapping.cc:435, src/LoopClosing.cc:729), so the context cannot live in the
// object and must not be shared: one orbm_matcher per thread, created on first use ON THE DEVICE THE EXTRACTORS LIVE ON
// (ORBextractor::SetDevice; orbm_stereo_match checks that both sides agree) and destroyed when the thread exits — its
// stream, pinned staging block and device scratch go with it.
#ifndef ORBX_THREAD_MATCHER_H_
#define ORBX_THREAD_MATCHER_H_

#include <stdexcept>

#include "ORBextractor.h"
#include "orbm.h"

namespace ORB_SLAM3 {

inline orbm_matcher* OrbxThreadMatcher() {
  struct Holder {
    orbm_matcher* m = nullptr;
    int device = -1;
    ~Holder() {
      if (m) orbm_destroy(m);
    }
  };
  thread_local Holder h;
  const int device = ORBextractor::GetDevice();
  if (h.m && h.device != device) {  // the application moved to another GPU: the old context is of no use
    orbm_destroy(h.m);
    h.m = nullptr;
  }
  if (!h.m) {
    if (orbm_create(&h.m, device) != ORBX_OK) throw std::runtime_error(orbm_last_error(nullptr));
    h.device = device;
  }
  return h.m;
}

}  // namespace ORB_SLAM3

#endif
