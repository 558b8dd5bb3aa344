// Batch-1 latency of the reference's per-frame call pattern through the C++ class surface:
//   Frame::ExtractKeyPoints -> SPextractor::operator()            (Frame.cc:544-559, once per incoming frame)
//   Tracking::TrackWithMotionModel -> SPmatcher::MatchingPoints_onnx(Frame, Frame)   (Tracking.cc:3465 -> SPmatcher.cc:457-542)
// i.e. per frame ONE extraction of the new frame and ONE match against the previous frame, host vectors / cv::Mat in and
// out, every host<->device copy inside the timed region.  Prints one JSON object (milliseconds).
//   latency_driver <h> <w> <imgA.raw> <imgB.raw> <iterations>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <chrono>
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
static double pct(std::vector<double> v, double p) {
  std::sort(v.begin(), v.end());
  return v[std::min(v.size() - 1, static_cast<size_t>(p * (v.size() - 1) + 0.5))];
}

int main(int argc, char** argv) {
  if (argc < 6) return 2;
  const int h = atoi(argv[1]), w = atoi(argv[2]), iters = atoi(argv[5]);
  cv::Mat img[2];
  img[0].create(h, w, CV_8UC1);
  img[1].create(h, w, CV_8UC1);
  if (!read_raw(argv[3], img[0]) || !read_raw(argv[4], img[1])) return 3;
  ORB_SLAM3::SPextractor ext(1000, 1.2f, 1, 20, 7);
  ORB_SLAM3::SPmatcher matcher(0.0f);
  ORB_SLAM3::Frame fr[2];
  fr[0].imgLeft = img[0];
  fr[1].imgLeft = img[1];
  ext(fr[0].imgLeft, fr[0].mvKeys, fr[0].mDescriptors);          // "previous frame"
  std::vector<double> t_ext, t_match, t_total;
  int nk = 0, nm = 0;
  for (int it = -5; it < iters; ++it) {                           // 5 warm-up frames
    ORB_SLAM3::Frame& cur = fr[(it + 5 + 1) & 1];
    ORB_SLAM3::Frame& last = fr[(it + 5) & 1];
    cur.mvKeys.clear();
    const auto t0 = std::chrono::high_resolution_clock::now();
    nk = ext(cur.imgLeft, cur.mvKeys, cur.mDescriptors);
    const auto t1 = std::chrono::high_resolution_clock::now();
    std::vector<int> vn;
    nm = matcher.MatchingPoints_onnx(cur, last, vn);
    const auto t2 = std::chrono::high_resolution_clock::now();
    if (it >= 0) {
      t_ext.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
      t_match.push_back(std::chrono::duration<double, std::milli>(t2 - t1).count());
      t_total.push_back(std::chrono::duration<double, std::milli>(t2 - t0).count());
    }
  }
  if (t_total.empty() || nk == 0) {
    printf("{\"error\": \"no keypoints / no device\"}\n");
    return 0;
  }
  printf("{\"frames\": %d, \"keypoints\": %d, \"matches\": %d, \"extract_ms_p50\": %.3f, \"extract_ms_p99\": %.3f, "
         "\"match_ms_p50\": %.3f, \"match_ms_p99\": %.3f, \"frame_ms_p50\": %.3f, \"frame_ms_p99\": %.3f}\n",
         iters, nk, nm, pct(t_ext, 0.5), pct(t_ext, 0.99), pct(t_match, 0.5), pct(t_match, 0.99), pct(t_total, 0.5), pct(t_total, 0.99));
  return 0;
}
