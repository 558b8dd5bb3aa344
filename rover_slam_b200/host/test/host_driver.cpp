// Host-side driver used by tests/test_gpu_host_classes.py: exercises the reference's call pattern
// (Frame::ExtractKeyPoints -> SPextractor::operator(); Tracking -> SPmatcher::MatchingPoints_onnx) through the
// C++ class surface and dumps the results for comparison with the C-ABI path by the Python test.
//   host_driver <h> <w> <imgA.raw> <imgB.raw> <out.bin>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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
  // ---- the remaining overloads and the container semantics the reference's callers rely on --------------------------------
  std::vector<cv::Point2f> pfa, pfb;
  for (const cv::KeyPoint& k : fa.mvKeys) pfa.emplace_back(k.pt);
  for (const cv::KeyPoint& k : fb.mvKeys) pfb.emplace_back(k.pt);
  // Point2f + Mat overload (SPmatcher.cc:374-410) on a NON-EMPTY vnMatches12: resize(n, -1) keeps the stale entries
  std::vector<int> vn_pf(5, 77);
  const int m_pf = matcher.MatchingPoints_onnx(pfa, pfb, fa.mDescriptors, fb.mDescriptors, vn_pf);
  // float* overload (SPmatcher.cc:359-372): only the count comes back
  std::vector<float> da_flat(static_cast<size_t>(na) * 256), db_flat(static_cast<size_t>(nb) * 256);
  for (int i = 0; i < na; ++i) memcpy(da_flat.data() + static_cast<size_t>(i) * 256, fa.mDescriptors.ptr<float>(i), 1024);
  for (int i = 0; i < nb; ++i) memcpy(db_flat.data() + static_cast<size_t>(i) * 256, fb.mDescriptors.ptr<float>(i), 1024);
  const int m_fp = matcher.MatchingPoints_onnx(pfa, pfb, da_flat.data(), db_flat.data());
  // operator() APPENDS to the keypoint vector and returns its size (superpoint_onnx.cc:230, SPextractor.cc:616)
  std::vector<cv::KeyPoint> pre(3, cv::KeyPoint(1.f, 2.f, 3.f));
  cv::Mat dpre;
  const int n_app_ret = ext(fa.imgLeft, pre, dpre);
  const int n_app_ok = (na > 0 && pre.size() == static_cast<size_t>(na) + 3 && pre[0].pt.x == 1.f && pre[2].size == 3.f &&
                        pre[3].pt.x == fa.mvKeys[0].pt.x && dpre.rows == na) ? 1 : 0;
  // Matcher_Inference(KeyPoint...) takes pt AS IS (lightglue_onnx.cpp:268-275): fed with normalised coordinates it must equal the
  // Point2f path; an extra, unrelated Matcher_PreProcess call in between must change nothing (the runner keeps no state)
  LightGlueDecoupleOnnxRunner* fm = matcher.featureMatcher;
  std::vector<cv::Point2f> nka = fm->Matcher_PreProcess(pfa, 300, 400);
  (void)fm->Matcher_PreProcess(pfb, 123, 457);              // odd number of PreProcess calls before the inference
  std::vector<cv::Point2f> nkb = fm->Matcher_PreProcess(fb.mvKeys, 300, 400);
  std::vector<cv::KeyPoint> kka, kkb;
  for (const cv::Point2f& q : nka) kka.emplace_back(q.x, q.y, 1.f);
  for (const cv::Point2f& q : nkb) kkb.emplace_back(q.x, q.y, 1.f);
  LightGlueResult r_kp = fm->Matcher_Inference(kka, kkb, da_flat.data(), db_flat.data());
  std::vector<int> vn_kpinf(na, -1);
  const int m_kpinf = fm->Matcher_PostProcess_fused(r_kp, pfa, pfb, vn_kpinf);
  // 3-channel NormalizeImage (BGR -> RGB, 1/255) and RGB2Grayscale (transform.cpp:3-17, :85-89)
  cv::Mat bgr(2, 2, CV_8UC3);
  for (int i = 0; i < 12; ++i) bgr.data[i] = static_cast<uint8_t>(20 * i + 5);
  cv::Mat rgbf = NormalizeImage(bgr);
  cv::Mat grayf = RGB2Grayscale(rgbf);
  const float px0[4] = {rgbf.ptr<float>(0)[0], rgbf.ptr<float>(0)[1], rgbf.ptr<float>(0)[2], grayf.ptr<float>(0)[0]};
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
  const int hdr2[6] = {m_pf, m_fp, n_app_ret, n_app_ok, m_kpinf, static_cast<int>(vn_pf.size())};
  put(f, hdr2, 6);
  put(f, vn_pf.data(), vn_pf.size());
  put(f, vn_kpinf.data(), vn_kpinf.size());
  put(f, px0, 4);
  fclose(f);
  printf("host_driver: %d / %d keypoints, multi=%d, matches frame=%d kp=%d\n", na, nb, nmulti, m_frame, m_kp);
  return 0;
}
