// The LightGlue entry points of ORB_SLAM3::SPmatcher on the B200 front end.
// Replaces SPmatcher.cc:17-27 (ctor) and :359-542 (the four MatchingPoints_onnx overloads) of the reference; the
// quirks that change results are reproduced: three overloads normalise with the HARD-CODED image size 300x400
// (SPmatcher.cc:360-361, :376-377, :414-415), the Frame overload with f2.imgLeft's size (:463-464);
// vnMatches12.resize(n, -1) keeps stale entries of a non-empty vector (:375, :413, :460).
#include "Matchers/SPmatcher.h"

#include <cmath>
#include <vector>

#ifndef ROVER_FE_STANDALONE
#include "Frame.h"
#endif

namespace ORB_SLAM3 {

const float SPmatcher::TH_HIGH = 1.4f;   // SPmatcher.cc:13
const float SPmatcher::TH_LOW = 1.2f;    // SPmatcher.cc:14
const int SPmatcher::HISTO_LENGTH = 30;

SPmatcher::SPmatcher(float thre) {
  Configuration cfg;
  cfg.device = "cuda";
  cfg.lightgluePath = "onnxmodel/lightglue_sim.onnx";
  featureMatcher = new LightGlueDecoupleOnnxRunner();
  featureMatcher->InitOrtEnv(cfg);
  featureMatcher->SetMatchThresh(thre);
}

namespace {
// contiguous [rows][cols] copy of a CV_32F descriptor matrix (the reference copies into a leaked new float[])
std::vector<float> flatten(const cv::Mat& m) {
  std::vector<float> v(static_cast<size_t>(m.rows) * m.cols);
  for (int i = 0; i < m.rows; ++i) {
    const float* r = m.ptr<float>(i);
    for (int j = 0; j < m.cols; ++j) v[static_cast<size_t>(i) * m.cols + j] = r[j];
  }
  return v;
}
}  // namespace

int SPmatcher::MatchingPoints_onnx(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1, float* desc0, float* desc1) {
  const int rows = 300, cols = 400;
  auto n0 = featureMatcher->Matcher_PreProcess(kpts0, rows, cols);
  auto n1 = featureMatcher->Matcher_PreProcess(kpts1, rows, cols);
  LightGlueResult out = featureMatcher->Matcher_Inference(n0, n1, desc0, desc1);
  std::vector<int> vnMatches12(n0.size(), -1);
  return featureMatcher->Matcher_PostProcess_fused(out, kpts0, kpts1, vnMatches12);
}

int SPmatcher::MatchingPoints_onnx(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1, cv::Mat desc0, cv::Mat desc1,
                                   std::vector<int>& vnMatches12) {
  vnMatches12.resize(kpts0.size(), -1);
  const int rows = 300, cols = 400;
  auto n0 = featureMatcher->Matcher_PreProcess(kpts0, rows, cols);
  auto n1 = featureMatcher->Matcher_PreProcess(kpts1, rows, cols);
  std::vector<float> d0 = flatten(desc0), d1 = flatten(desc1);
  LightGlueResult out = featureMatcher->Matcher_Inference(n0, n1, d0.data(), d1.data());
  return featureMatcher->Matcher_PostProcess_fused(out, kpts0, kpts1, vnMatches12);
}

int SPmatcher::MatchingPoints_onnx(std::vector<cv::KeyPoint> kpts0, const std::vector<cv::KeyPoint> kpts1, cv::Mat desc0,
                                   const cv::Mat desc1, std::vector<int>& vnMatches12) {
  vnMatches12.resize(kpts0.size(), -1);
  const int rows = 300, cols = 400;
  std::vector<cv::Point2f> pf0, pf1;
  for (const cv::KeyPoint& k : kpts0) pf0.emplace_back(k.pt);
  for (const cv::KeyPoint& k : kpts1) pf1.emplace_back(k.pt);
  auto n0 = featureMatcher->Matcher_PreProcess(kpts0, rows, cols);
  auto n1 = featureMatcher->Matcher_PreProcess(kpts1, rows, cols);
  std::vector<float> d0 = flatten(desc0), d1 = flatten(desc1);
  LightGlueResult out = featureMatcher->Matcher_Inference(n0, n1, d0.data(), d1.data());
  return featureMatcher->Matcher_PostProcess_fused(out, pf0, pf1, vnMatches12);
}

int SPmatcher::MatchingPoints_onnx(Frame& f1, Frame& f2, std::vector<int>& vnMatches12) {
  vnMatches12.resize(f1.mvKeys.size(), -1);
  const int rows = f2.imgLeft.rows, cols = f2.imgLeft.cols;
  std::vector<cv::Point2f> kpts1, kpts2;
  for (const cv::KeyPoint& k : f1.mvKeys) kpts1.emplace_back(k.pt);
  for (const cv::KeyPoint& k : f2.mvKeys) kpts2.emplace_back(k.pt);
  auto n1 = featureMatcher->Matcher_PreProcess(kpts1, rows, cols);
  auto n2 = featureMatcher->Matcher_PreProcess(kpts2, rows, cols);
  std::vector<float> d1 = flatten(f1.mDescriptors), d2 = flatten(f2.mDescriptors);
  LightGlueResult out = featureMatcher->Matcher_Inference(n1, n2, d1.data(), d2.data());
  return featureMatcher->Matcher_PostProcess_fused(out, kpts1, kpts2, vnMatches12);
}

float SPmatcher::DescriptorDistance_sp(const cv::Mat& a, const cv::Mat& b) {
  const float* pa = a.ptr<float>(0);
  const float* pb = b.ptr<float>(0);
  float s = 0.0f;
  for (int i = 0; i < a.cols; ++i) {
    const float d = pa[i] - pb[i];
    s += d * d;
  }
  return std::sqrt(s);
}

}  // namespace ORB_SLAM3
