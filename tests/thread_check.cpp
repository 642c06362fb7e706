// thread_check — the threading contract of SURVEY.md §8(b) on the CUDA library, through the C ABI, against the CPU
// oracle's stored outputs (tools/ubench/data/hotpath_case.bin: 16 stereo pairs with keypoints, descriptors, mvuRight,
// mvDepth):
//   1. the stereo Frame constructor's pattern (src/Frame.cc:200-203): the left and right extractor OBJECTS run on two
//      std::threads at the same time, every frame; then Frame::ComputeStereoMatches on the caller's thread;
//   2. ORBmatcher temporaries on three threads at once (Tracking / LocalMapping / LoopClosing, include/ORBmatcher.h:38-133):
//      three threads, each with its own thread-local matcher context (created on first use, destroyed at thread exit, as
//      shim/orbx_thread_matcher.h does) and its own extractor pair, hammer extract + stereo match + knnMatch concurrently.
// Every result must equal the oracle's (and therefore the serial run's). Exit code 0 = all equal.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "orbm.h"
#include "orbx.h"

struct Eye {
  int32_t n, mono;
  std::vector<uint8_t> kps, desc;
};
struct Pair {
  Eye eye[2];
  int32_t n_matched;
  std::vector<float> u_right, depth;
};
static bool rd(FILE* f, void* p, size_t n) { return fread(p, 1, n, f) == n; }

struct MatcherHolder {  // shim/orbx_thread_matcher.h's pattern
  orbm_matcher* m = nullptr;
  ~MatcherHolder() {
    if (m) orbm_destroy(m);
  }
};
static orbm_matcher* thread_matcher() {
  thread_local MatcherHolder h;
  if (!h.m && orbm_create(&h.m, 0) != ORBX_OK) {
    printf("orbm_create: %s\n", orbm_last_error(nullptr));
    exit(3);
  }
  return h.m;
}

static int W, H, NFEAT, CAP;
static float MBF, MB;
static std::vector<uint8_t> g_img[2];
static std::vector<Pair> g_want;
static std::atomic<long> g_bad{0};

struct Out {
  std::vector<orbx_kp> kps;
  std::vector<uint8_t> desc;
  int32_t n = 0, mono = 0;
};

static void extract_eye(orbx_extractor* ex, int pair, int eye, Out* o) {
  o->kps.resize(CAP);
  o->desc.resize((size_t)CAP * 32);
  const int rc = orbx_extract(ex, g_img[eye].data() + (size_t)pair * W * H, W, H, W, 0, 0, o->kps.data(), o->desc.data(), CAP,
                              &o->n, &o->mono);
  if (rc != ORBX_OK) {
    printf("orbx_extract: rc %d: %s\n", rc, orbx_last_error(ex));
    g_bad++;
  }
}

static void check_eye(const Out& o, int pair, int eye, const char* who) {
  const Eye& y = g_want[pair].eye[eye];
  if (o.n != y.n || o.mono != y.mono || memcmp(o.kps.data(), y.kps.data(), (size_t)y.n * 28) ||
      memcmp(o.desc.data(), y.desc.data(), (size_t)y.n * 32)) {
    if (g_bad++ < 8) printf("%s: pair %d eye %d differs from the oracle (n %d vs %d)\n", who, pair, eye, o.n, y.n);
  }
}

static void stereo_and_check(orbx_extractor* l, orbx_extractor* r, const Out& ol, const Out& orr, int pair, const char* who) {
  std::vector<float> ur(ol.n > 0 ? ol.n : 1), dp(ol.n > 0 ? ol.n : 1);
  int32_t nm = 0;
  orbm_matcher* m = thread_matcher();
  const int rc = orbm_stereo_match(m, l, r, 0, ol.kps.data(), ol.desc.data(), ol.n, orr.kps.data(), orr.desc.data(), orr.n, MBF, MB,
                                   ur.data(), dp.data(), &nm);
  const Pair& w = g_want[pair];
  if (rc != ORBX_OK || nm != w.n_matched || memcmp(ur.data(), w.u_right.data(), (size_t)ol.n * 4) ||
      memcmp(dp.data(), w.depth.data(), (size_t)ol.n * 4)) {
    if (g_bad++ < 8) printf("%s: pair %d stereo differs (rc %d, matched %d vs %d)\n", who, pair, rc, nm, w.n_matched);
  }
}

int main(int argc, char** argv) {
  std::string path = argc > 1 ? argv[1] : "tools/ubench/data/hotpath_case.bin";
  const int rounds = argc > 2 ? atoi(argv[2]) : 3;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return printf("cannot open %s\n", path.c_str()), 1;
  int32_t hdr[5];
  if (!rd(f, hdr, 20) || hdr[0] != 0x4F524258 || !rd(f, &MBF, 4) || !rd(f, &MB, 4)) return printf("bad case file\n"), 1;
  const int D = hdr[1];
  W = hdr[2];
  H = hdr[3];
  NFEAT = hdr[4];
  const size_t fb = (size_t)W * H;
  for (int e = 0; e < 2; e++) {
    g_img[e].resize(D * fb);
    if (!rd(f, g_img[e].data(), D * fb)) return printf("short case file\n"), 1;
  }
  g_want.resize(D);
  for (int i = 0; i < D; i++) {
    for (int e = 0; e < 2; e++) {
      Eye& y = g_want[i].eye[e];
      if (!rd(f, &y.n, 4) || !rd(f, &y.mono, 4)) return printf("short case file\n"), 1;
      y.kps.resize((size_t)y.n * 28);
      y.desc.resize((size_t)y.n * 32);
      if (!rd(f, y.kps.data(), y.kps.size()) || !rd(f, y.desc.data(), y.desc.size())) return printf("short case file\n"), 1;
    }
    const int nl = g_want[i].eye[0].n;
    g_want[i].u_right.resize(nl);
    g_want[i].depth.resize(nl);
    if (!rd(f, &g_want[i].n_matched, 4) || !rd(f, g_want[i].u_right.data(), (size_t)nl * 4) ||
        !rd(f, g_want[i].depth.data(), (size_t)nl * 4))
      return printf("short case file\n"), 1;
  }
  fclose(f);

  // ---- 1. the stereo Frame constructor: two extractor objects on two threads, per frame ----
  {
    orbx_extractor *l = nullptr, *r = nullptr;
    if (orbx_extractor_create(&l, 0, NFEAT, 1.2f, 8, 20, 7, 1) || orbx_extractor_create(&r, 0, NFEAT, 1.2f, 8, 20, 7, 1))
      return printf("create: %s\n", orbx_last_error(nullptr)), 1;
    CAP = orbx_extractor_capacity(l);
    for (int round = 0; round < rounds; round++)
      for (int p = 0; p < D; p++) {
        Out ol, orr;
        std::thread tl(extract_eye, l, p, 0, &ol), tr(extract_eye, r, p, 1, &orr);  // src/Frame.cc:200-203
        tl.join();
        tr.join();
        check_eye(ol, p, 0, "frame-ctor");
        check_eye(orr, p, 1, "frame-ctor");
        stereo_and_check(l, r, ol, orr, p, "frame-ctor");
      }
    orbx_extractor_destroy(l);
    orbx_extractor_destroy(r);
    printf("two extractor objects on two threads, %d frames: %s\n", rounds * D, g_bad ? "FAILED" : "equal to the oracle");
  }

  // ---- 2. three threads, each with a thread-local matcher and its own extractor pair, all at once ----
  {
    // knn2 reference: the serial answer of the same library for the first pair's descriptors (oracle-checked elsewhere)
    std::vector<std::vector<int32_t>> knn_want(D);
    auto worker = [&](int t, bool record) {
      orbx_extractor *l = nullptr, *r = nullptr;
      if (orbx_extractor_create(&l, 0, NFEAT, 1.2f, 8, 20, 7, 1) || orbx_extractor_create(&r, 0, NFEAT, 1.2f, 8, 20, 7, 1)) {
        g_bad++;
        return;
      }
      for (int round = 0; round < rounds; round++)
        for (int p = record ? 0 : t; p < D; p += record ? 1 : 3) {
          Out ol, orr;
          extract_eye(l, p, 0, &ol);
          extract_eye(r, p, 1, &orr);
          check_eye(ol, p, 0, "3-threads");
          check_eye(orr, p, 1, "3-threads");
          stereo_and_check(l, r, ol, orr, p, "3-threads");
          // cv::BFMatcher::knnMatch of the left descriptors against the right ones (src/Frame.cc:1293)
          std::vector<int32_t> k((size_t)4 * (ol.n > 0 ? ol.n : 1));
          const int n = ol.n;
          const int rc = orbm_knn2(thread_matcher(), ol.desc.data(), n, orr.desc.data(), orr.n, k.data(), k.data() + n,
                                   k.data() + 2 * (size_t)n, k.data() + 3 * (size_t)n);
          k.resize((size_t)4 * n);
          if (record) knn_want[p] = k;
          else if (rc != ORBX_OK || k != knn_want[p]) {
            if (g_bad++ < 8) printf("3-threads: pair %d knn2 differs from the serial run (rc %d)\n", p, rc);
          }
        }
      orbx_extractor_destroy(l);
      orbx_extractor_destroy(r);
    };
    const int rounds_keep = rounds;
    (void)rounds_keep;
    worker(0, true);  // serial pass: records the knn2 answers (and checks everything else against the oracle)
    std::thread a(worker, 0, false), b(worker, 1, false), c(worker, 2, false);
    a.join();
    b.join();
    c.join();
    printf("three threads with thread-local matchers, %d frames each way: %s\n", rounds * D, g_bad ? "FAILED" : "equal to the oracle / serial run");
  }
  return g_bad ? 2 : 0;
}
