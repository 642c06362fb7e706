// tests/shim_smoke.cpp — TEST INFRASTRUCTURE. Drives shim/ORBextractor.{h,cc} exactly the way the reference's
// Frame::ExtractORB does (src/Frame.cc:549-560) and dumps the outputs for comparison with the oracle.
// usage: shim_smoke <in.raw> <w> <h> <nfeatures> <lap0> <lap1> <out.bin>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ORBextractor.h"

int main(int argc, char** argv) {
  if (argc < 8) return 2;
  const int w = atoi(argv[2]), h = atoi(argv[3]), nf = atoi(argv[4]);
  cv::Mat im(h, w, CV_8UC1);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(im.data, 1, (size_t)w * h, f) != (size_t)w * h) return 3;
  fclose(f);
  ORB_SLAM3::ORBextractor ex(nf, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> keys;
  cv::Mat desc;
  std::vector<int> lap = {atoi(argv[5]), atoi(argv[6])};
  int mono = 0;
  for (int rep = 0; rep < 2; rep++) mono = ex(im, cv::Mat(), keys, desc, lap);  // second call reuses the handle
  FILE* o = fopen(argv[7], "wb");
  const int n = (int)keys.size();
  fwrite(&mono, 4, 1, o);
  fwrite(&n, 4, 1, o);
  fwrite(keys.data(), sizeof(cv::KeyPoint), n, o);
  for (int i = 0; i < n; i++) fwrite(desc.ptr(i), 1, 32, o);
  // mvImagePyramid[3]: ROI of the bordered mirror; dump the bordered buffer through the ROI's negative offsets
  const cv::Mat& l3 = ex.mvImagePyramid[3];
  int lw = l3.cols, lh = l3.rows, st = (int)l3.step;
  fwrite(&lw, 4, 1, o);
  fwrite(&lh, 4, 1, o);
  for (int y = -19; y < lh + 19; y++) fwrite(l3.data + (ptrdiff_t)y * st - 19, 1, lw + 38, o);
  fclose(o);
  std::vector<float> sf = ex.GetScaleFactors();
  printf("mono=%d n=%d levels=%d sf1=%.9g\n", mono, n, ex.GetLevels(), sf[1]);
  return 0;
}
