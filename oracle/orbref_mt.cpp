// orbref_mt.cpp — frame-parallel driver around the CPU oracle (test infrastructure; see orbref.h).
// Used only by bench.py's cpu_baseline / --impl reference legs: runs independent frames on all host cores, the most
// favourable CPU scaling for the reference (frames are independent).
#include <atomic>
#include <thread>
#include <vector>

#include "orbref.h"

extern "C" {

// Extracts n_frames images (each w x h, contiguous, frame_stride bytes apart) with `threads` workers, one private
// extractor per worker. counts[n_frames] receives the keypoint count per frame; if kps/desc are non-NULL they receive
// cap entries per frame. Returns 0.
int orbref_extract_many(const uint8_t* imgs, int n_frames, int w, int h, long frame_stride, int nfeatures,
                        float scale_factor, int nlevels, int ini_th, int min_th, int lap0, int lap1, int threads,
                        orbx_kp* kps, uint8_t* desc, int cap, int* counts) {
  std::atomic<int> next(0);
  auto worker = [&]() {
    orbref_extractor* ex = orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th);
    std::vector<orbx_kp> k(cap);
    std::vector<uint8_t> d((size_t)cap * 32);
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= n_frames) break;
      int n = 0, mono = 0;
      orbx_kp* ko = kps ? kps + (size_t)i * cap : k.data();
      uint8_t* dd = desc ? desc + (size_t)i * cap * 32 : d.data();
      orbref_extract(ex, imgs + (size_t)i * frame_stride, w, h, w, lap0, lap1, ko, dd, cap, &n, &mono);
      if (counts) counts[i] = n;
    }
    orbref_extractor_destroy(ex);
  };
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  return 0;
}

// Stereo pairs: extract left + right and run ComputeStereoMatches, pair-parallel over `threads` workers.
// imgs_l / imgs_r hold n_pairs images each. matched[n_pairs] receives the surviving match count.
int orbref_stereo_many(const uint8_t* imgs_l, const uint8_t* imgs_r, int n_pairs, int w, int h, long frame_stride,
                       int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, float mbf, float mb,
                       int threads, int* counts_l, int* counts_r, int* matched) {
  std::atomic<int> next(0);
  const int cap = nfeatures + 4 * nlevels + 64;
  auto worker = [&]() {
    orbref_extractor* el = orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th);
    orbref_extractor* er = orbref_extractor_create(nfeatures, scale_factor, nlevels, ini_th, min_th);
    std::vector<orbx_kp> kl(cap), kr(cap);
    std::vector<uint8_t> dl((size_t)cap * 32), dr((size_t)cap * 32);
    std::vector<float> ur(cap), dep(cap);
    for (;;) {
      int i = next.fetch_add(1);
      if (i >= n_pairs) break;
      int nl = 0, nr = 0, mono = 0;
      orbref_extract(el, imgs_l + (size_t)i * frame_stride, w, h, w, 0, 0, kl.data(), dl.data(), cap, &nl, &mono);
      orbref_extract(er, imgs_r + (size_t)i * frame_stride, w, h, w, 0, 0, kr.data(), dr.data(), cap, &nr, &mono);
      int m = orbref_stereo_match(el, er, kl.data(), dl.data(), nl, kr.data(), dr.data(), nr, mbf, mb, ur.data(),
                                  dep.data());
      if (counts_l) counts_l[i] = nl;
      if (counts_r) counts_r[i] = nr;
      if (matched) matched[i] = m;
    }
    orbref_extractor_destroy(el);
    orbref_extractor_destroy(er);
  };
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) pool.emplace_back(worker);
  for (auto& t : pool) t.join();
  return 0;
}

}  // extern "C"
