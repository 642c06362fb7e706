// hotpath_check — the bench's hot path (extract x2 + ComputeStereoMatches, device resident) through the C ABI without
// Python: (1) parity of every output word against the CPU oracle's outputs stored in data/hotpath_case.bin
// (tools/ubench/make_hotpath_case.py writes it here, before gpurun), (2) per-stage and whole-step timing like bench.py's
// device-resident leg. A run costs seconds of GPU box time instead of the minutes a pytest + bench.py call takes, which
// is what a kernel-tuning iteration needs; the full `pytest -m gpu` stays the gate before a commit.
//
//   nvcc -O2 -std=c++17 -o hotpath_check hotpath_check.cu -I../../include -L../../orb_slam3_fast_b200 -lorbx \
//        -Xlinker -rpath -Xlinker '$ORIGIN/../../orb_slam3_fast_b200'
//   ./hotpath_check [pairs per step = 1024] [steps = 10] [e2e steps = steps, 0 = skip] [case file]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <string>
#include <vector>

#include "orbm.h"
#include "orbx.h"

#define CK(x)                                                                   \
  do {                                                                          \
    cudaError_t e_ = (x);                                                       \
    if (e_ != cudaSuccess) {                                                    \
      printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return 1;                                                                 \
    }                                                                           \
  } while (0)
#define OX(ex, x)                                                                   \
  do {                                                                              \
    int rc_ = (x);                                                                  \
    if (rc_ != 0) {                                                                 \
      printf("%s:%d %s: rc %d: %s\n", __FILE__, __LINE__, #x, rc_, orbx_last_error(ex)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

struct Eye {
  int32_t n, mono;
  std::vector<uint8_t> kps, desc;
};
struct Pair {
  Eye eye[2];
  int32_t n_matched;
  std::vector<float> u_right, depth;
};

struct DevOut {
  orbx_kp* kps;
  uint8_t* desc;
  int32_t *n, *mono, *status;
};

static int alloc_out(DevOut* o, int P, int cap) {
  CK(cudaMalloc(&o->kps, (size_t)P * cap * sizeof(orbx_kp)));
  CK(cudaMalloc(&o->desc, (size_t)P * cap * 32));
  CK(cudaMalloc(&o->n, (size_t)P * 4));
  CK(cudaMalloc(&o->mono, (size_t)P * 4));
  CK(cudaMalloc(&o->status, (size_t)P * 4));
  return 0;
}

int main(int argc, char** argv) {
  const int P = argc > 1 ? atoi(argv[1]) : 1024, steps = argc > 2 ? atoi(argv[2]) : 10;
  const int e2e_steps = argc > 3 ? atoi(argv[3]) : steps;
  std::string path = argc > 4 ? argv[4] : "";
  if (path.empty()) {
    path = argv[0];
    const size_t s = path.find_last_of('/');
    path = (s == std::string::npos ? std::string(".") : path.substr(0, s)) + "/data/hotpath_case.bin";
  }
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return printf("cannot open %s (run tools/ubench/make_hotpath_case.py first)\n", path.c_str()), 1;
  int32_t hdr[5];
  float mbf, mb;
  if (fread(hdr, 4, 5, f) != 5 || hdr[0] != 0x4F524258 || fread(&mbf, 4, 1, f) != 1 || fread(&mb, 4, 1, f) != 1)
    return printf("bad case file\n"), 1;
  const int D = hdr[1], W = hdr[2], H = hdr[3], nfeat = hdr[4];
  const size_t fbytes = (size_t)W * H;
  std::vector<uint8_t> imgs[2] = {std::vector<uint8_t>(D * fbytes), std::vector<uint8_t>(D * fbytes)};
  for (int e = 0; e < 2; e++)
    if (fread(imgs[e].data(), 1, D * fbytes, f) != D * fbytes) return printf("short case file\n"), 1;
  std::vector<Pair> want(D);
  for (int i = 0; i < D; i++) {
    for (int e = 0; e < 2; e++) {
      Eye& y = want[i].eye[e];
      if (fread(&y.n, 4, 1, f) != 1 || fread(&y.mono, 4, 1, f) != 1) return printf("short case file\n"), 1;
      y.kps.resize((size_t)y.n * 28);
      y.desc.resize((size_t)y.n * 32);
      if (fread(y.kps.data(), 1, y.kps.size(), f) != y.kps.size() ||
          fread(y.desc.data(), 1, y.desc.size(), f) != y.desc.size())
        return printf("short case file\n"), 1;
    }
    const int nl = want[i].eye[0].n;
    want[i].u_right.resize(nl);
    want[i].depth.resize(nl);
    if (fread(&want[i].n_matched, 4, 1, f) != 1 || fread(want[i].u_right.data(), 4, nl, f) != (size_t)nl ||
        fread(want[i].depth.data(), 4, nl, f) != (size_t)nl)
      return printf("short case file\n"), 1;
  }
  fclose(f);
  static_assert(sizeof(orbx_kp) == 28, "cv::KeyPoint layout");

  orbx_extractor* ex[2] = {nullptr, nullptr};
  orbm_matcher* mt = nullptr;
  for (int e = 0; e < 2; e++) OX(nullptr, orbx_extractor_create(&ex[e], 0, nfeat, 1.2f, 8, 20, 7, P));
  if (orbm_create(&mt, 0) != 0) return printf("orbm_create: %s\n", orbm_last_error(nullptr)), 1;
  const int cap = orbx_extractor_capacity(ex[0]);
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));

  // two rotating input batches of P pairs (pair p of batch k = distinct pair (p + k) % D), like bench.py
  uint8_t* d_img[2][2];
  for (int k = 0; k < 2; k++)
    for (int e = 0; e < 2; e++) {
      CK(cudaMalloc(&d_img[k][e], (size_t)P * fbytes));
      for (int p = 0; p < P; p++)
        CK(cudaMemcpyAsync(d_img[k][e] + (size_t)p * fbytes, imgs[e].data() + (size_t)((p + k) % D) * fbytes, fbytes,
                           cudaMemcpyHostToDevice, st));
    }
  DevOut o[2];
  for (int e = 0; e < 2; e++)
    if (alloc_out(&o[e], P, cap)) return 1;
  float *d_ur, *d_dp;
  int32_t* d_nm;
  CK(cudaMalloc(&d_ur, (size_t)P * cap * 4));
  CK(cudaMalloc(&d_dp, (size_t)P * cap * 4));
  CK(cudaMalloc(&d_nm, (size_t)P * 4));
  CK(cudaStreamSynchronize(st));

  auto step = [&](int k) -> int {
    for (int e = 0; e < 2; e++)
      OX(ex[e], orbx_extract_batch_device(ex[e], P, d_img[k & 1][e], W, H, W, (int64_t)fbytes, 0, 0, o[e].kps, o[e].desc,
                                          cap, o[e].n, o[e].mono, o[e].status, st));
    const int rc = orbm_stereo_match_batch_device(mt, ex[0], ex[1], P, o[0].kps, o[0].desc, o[0].n, o[1].kps, o[1].desc,
                                                  o[1].n, cap, mbf, mb, d_ur, d_dp, d_nm, st);
    if (rc != 0) return printf("stereo match: rc %d: %s\n", rc, orbm_last_error(mt)), 1;
    return 0;
  };

  // ---- parity: the first min(P, D) pairs of batch 0 are the case's pairs in order ----
  if (step(0)) return 1;
  CK(cudaStreamSynchronize(st));
  const int nchk = P < D ? P : D;
  long bad = 0;
  for (int i = 0; i < nchk; i++) {
    int32_t n[2], mono[2], status[2], nm;
    for (int e = 0; e < 2; e++) {
      CK(cudaMemcpy(&n[e], o[e].n + i, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&mono[e], o[e].mono + i, 4, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(&status[e], o[e].status + i, 4, cudaMemcpyDeviceToHost));
      const Eye& y = want[i].eye[e];
      bool ok = n[e] == y.n && mono[e] == y.mono && status[e] == 0;
      if (ok) {
        std::vector<uint8_t> k((size_t)y.n * 28), d((size_t)y.n * 32);
        CK(cudaMemcpy(k.data(), o[e].kps + (size_t)i * cap, k.size(), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(d.data(), o[e].desc + (size_t)i * cap * 32, d.size(), cudaMemcpyDeviceToHost));
        ok = memcmp(k.data(), y.kps.data(), k.size()) == 0 && memcmp(d.data(), y.desc.data(), d.size()) == 0;
      }
      if (!ok && bad++ < 8) printf("pair %d eye %d differs (n %d vs %d, status %d)\n", i, e, n[e], y.n, status[e]);
    }
    CK(cudaMemcpy(&nm, d_nm + i, 4, cudaMemcpyDeviceToHost));
    const int nl = want[i].eye[0].n;
    std::vector<float> ur(nl), dp(nl);
    CK(cudaMemcpy(ur.data(), d_ur + (size_t)i * cap, (size_t)nl * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dp.data(), d_dp + (size_t)i * cap, (size_t)nl * 4, cudaMemcpyDeviceToHost));
    if ((nm != want[i].n_matched || memcmp(ur.data(), want[i].u_right.data(), (size_t)nl * 4) != 0 ||
         memcmp(dp.data(), want[i].depth.data(), (size_t)nl * 4) != 0) && bad++ < 8)
      printf("pair %d stereo differs (n_matched %d vs %d)\n", i, nm, want[i].n_matched);
  }
  printf("parity vs oracle: %s (%d pairs: keypoints, descriptors, uRight, depth, counts)\n",
         bad ? "FAILED" : "bit-exact", nchk);

  // ---- timing ----
  for (int k = 0; k < 3; k++)
    if (step(k)) return 1;
  CK(cudaStreamSynchronize(st));
  for (int e = 0; e < 2; e++) {
    orbx_profile_enable(ex[e], 1);
    orbx_profile_read(ex[e], nullptr, nullptr, 1);
  }
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0, st));
  for (int k = 0; k < steps; k++)
    if (step(k)) return 1;
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  float stage[5] = {0, 0, 0, 0, 0};
  for (int e = 0; e < 2; e++) {
    float s[5];
    int32_t c[5];
    if (orbx_profile_read(ex[e], s, c, 1) == 0)
      for (int k = 0; k < 5; k++) stage[k] += s[k];
    orbx_profile_enable(ex[e], 0);
  }
  printf("%d pairs/step x %d steps: %.3f ms/step = %.0f frames/s device-resident\n", P, steps, ms / steps,
         2.0 * P * steps / (ms * 1e-3));
  printf("stage ms/step: pyramid %.3f fast %.3f quadtree %.3f blur %.3f describe %.3f\n", stage[0] / steps,
         stage[1] / steps, stage[2] / steps, stage[3] / steps, stage[4] / steps);

  // ---- end to end: orbm_stereo_frames_batch, pinned host buffers in and out (bench.py's e2e leg) ----
  if (e2e_steps > 0) {
    const int G = P / 8 < 8 ? 8 : (P / 8 > 64 ? 64 : P / 8);
    orbx_extractor* ex2[2] = {nullptr, nullptr};
    for (int e = 0; e < 2; e++) OX(nullptr, orbx_extractor_create(&ex2[e], 0, nfeat, 1.2f, 8, 20, 7, G));
    uint8_t* h_img[2][2];
    for (int k = 0; k < 2; k++)
      for (int e = 0; e < 2; e++) {
        h_img[k][e] = static_cast<uint8_t*>(orbx_host_alloc((int64_t)P * fbytes));
        if (!h_img[k][e]) return printf("orbx_host_alloc failed\n"), 1;
        for (int p = 0; p < P; p++)
          memcpy(h_img[k][e] + (size_t)p * fbytes, imgs[e].data() + (size_t)((p + k) % D) * fbytes, fbytes);
      }
    orbx_kp* h_kps[2];
    uint8_t* h_desc[2];
    int32_t* h_n[2];
    for (int e = 0; e < 2; e++) {
      h_kps[e] = static_cast<orbx_kp*>(orbx_host_alloc((int64_t)P * cap * sizeof(orbx_kp)));
      h_desc[e] = static_cast<uint8_t*>(orbx_host_alloc((int64_t)P * cap * 32));
      h_n[e] = static_cast<int32_t*>(orbx_host_alloc((int64_t)P * 4));
    }
    float* h_ur = static_cast<float*>(orbx_host_alloc((int64_t)P * cap * 4));
    float* h_dp = static_cast<float*>(orbx_host_alloc((int64_t)P * cap * 4));
    int32_t* h_nm = static_cast<int32_t*>(orbx_host_alloc((int64_t)P * 4));
    auto e2e = [&](int k) -> int {
      const int rc = orbm_stereo_frames_batch(mt, ex2[0], ex2[1], P, h_img[k & 1][0], h_img[k & 1][1], W, H, W,
                                              (int64_t)fbytes, mbf, mb, h_kps[0], h_desc[0], h_n[0], h_kps[1], h_desc[1],
                                              h_n[1], cap, h_ur, h_dp, h_nm);
      if (rc != 0) return printf("orbm_stereo_frames_batch: rc %d: %s\n", rc, orbm_last_error(mt)), 1;
      return 0;
    };
    for (int k = 0; k < 3; k++)
      if (e2e(k)) return 1;
    // the host-facing call must give the oracle's words too (batch 0 = the case's pairs in order)
    long bad2 = 0;
    if (e2e(0)) return 1;
    for (int i = 0; i < nchk; i++) {
      const Pair& w = want[i];
      const int nl = w.eye[0].n;
      bool ok = h_n[0][i] == nl && h_n[1][i] == w.eye[1].n && h_nm[i] == w.n_matched;
      ok = ok && memcmp(h_kps[0] + (size_t)i * cap, w.eye[0].kps.data(), (size_t)nl * 28) == 0 &&
           memcmp(h_desc[1] + (size_t)i * cap * 32, w.eye[1].desc.data(), (size_t)w.eye[1].n * 32) == 0 &&
           memcmp(h_ur + (size_t)i * cap, w.u_right.data(), (size_t)nl * 4) == 0 &&
           memcmp(h_dp + (size_t)i * cap, w.depth.data(), (size_t)nl * 4) == 0;
      if (!ok && bad2++ < 8) printf("e2e pair %d differs\n", i);
    }
    printf("e2e parity vs oracle: %s\n", bad2 ? "FAILED" : "bit-exact");
    bad += bad2;
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (int k = 0; k < e2e_steps; k++)
      if (e2e(k)) return 1;
    CK(cudaDeviceSynchronize());
    clock_gettime(CLOCK_MONOTONIC, &t1);
    const double dt = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
    printf("e2e (host buffers, groups of %d pairs): %.3f ms/step = %.0f frames/s\n", G, 1e3 * dt / e2e_steps,
           2.0 * P * e2e_steps / dt);
    for (int e = 0; e < 2; e++) orbx_extractor_destroy(ex2[e]);
  }
  orbm_destroy(mt);
  for (int e = 0; e < 2; e++) orbx_extractor_destroy(ex[e]);
  return bad ? 2 : 0;
}
