// Host-side driver used by tests/test_gpu_host_classes.py: exercises the reference's call pattern
// (Frame::ExtractKeyPoints -> SPextractor::operator(); Tracking -> SPmatcher::MatchingPoints_onnx) through the
// C++ class surface and dumps the results for comparison with the C-ABI path by the Python test.
//   host_driver <h> <w> <imgA.raw> <imgB.raw> <out.bin>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "Extractors/SPextractor.h"
#include "Matchers/SPmatcher.h"

static bool read_raw(const char* path, cv::Mat& m) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  const size_t n = fread(m.data, 1, static_cast<size_t>(m.rows) * m.cols, f);
  fclose(f);
  return n == static_cast<size_t>(m.rows) * m.cols;
}

template <typename T>
static void put(FILE* f, const T* p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int h = atoi(argv[1]), w = atoi(argv[2]);
  ORB_SLAM3::Frame fa, fb;
  fa.imgLeft.create(h, w, CV_8UC1);
  fb.imgLeft.create(h, w, CV_8UC1);
  if (!read_raw(argv[3], fa.imgLeft) || !read_raw(argv[4], fb.imgLeft)) return 3;
  ORB_SLAM3::SPextractor ext(1000, 1.2f, 1, 20, 7);           // Tracking.cc:645-651 style construction, nLevels = 1
  const int na = ext(fa.imgLeft, fa.mvKeys, fa.mDescriptors);
  const int nb = ext(fb.imgLeft, fb.mvKeys, fb.mDescriptors);
  ORB_SLAM3::SPextractor multi(1000, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> dummy;
  cv::Mat dd;
  const int nmulti = multi(fa.imgLeft, dummy, dd);            // reference yields 0 for nLevels != 1
  ORB_SLAM3::SPmatcher matcher(0.0f);
  std::vector<int> vn_frame, vn_kp;
  const int m_frame = matcher.MatchingPoints_onnx(fa, fb, vn_frame);
  const int m_kp = matcher.MatchingPoints_onnx(fa.mvKeys, fb.mvKeys, fa.mDescriptors, fb.mDescriptors, vn_kp);
  // the reference's compiled-out adaptive score rule, switched on (LocalMapping.cc:951-952 writes lastmatchnum)
  ORB_SLAM3::SPextractor ada(1000, 1.2f, 1, 20, 7);
  ada.featureExtractor->adaptive_threshold = true;
  ada.lastmatchnum = 150.0f;
  std::vector<cv::KeyPoint> ka;
  cv::Mat da;
  const int n_ada = ada(fa.imgLeft, ka, da);
  FILE* f = fopen(argv[5], "wb");
  if (!f) return 4;
  const int hdr[6] = {na, nb, nmulti, m_frame, m_kp, n_ada};
  put(f, hdr, 6);
  for (const auto* fr : {&fa, &fb}) {
    for (const cv::KeyPoint& k : fr->mvKeys) {
      const float v[3] = {k.pt.x, k.pt.y, k.response};
      put(f, v, 3);
    }
    for (int i = 0; i < fr->mDescriptors.rows; ++i) put(f, fr->mDescriptors.ptr<float>(i), 256);
  }
  put(f, vn_frame.data(), vn_frame.size());
  put(f, vn_kp.data(), vn_kp.size());
  for (const cv::KeyPoint& k : ka) {
    const float v[3] = {k.pt.x, k.pt.y, k.response};
    put(f, v, 3);
  }
  for (int i = 0; i < da.rows; ++i) put(f, da.ptr<float>(i), 256);
  fclose(f);
  printf("host_driver: %d / %d keypoints, multi=%d, matches frame=%d kp=%d\n", na, nb, nmulti, m_frame, m_kp);
  return 0;
}
