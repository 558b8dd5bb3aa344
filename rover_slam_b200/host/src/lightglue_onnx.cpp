// LightGlueDecoupleOnnxRunner on librover_fe.so (reference: src/Matchers/lightglue_onnx.cpp).
#include "Matchers/lightglue_onnx.h"

#include <stdlib.h>

#include <algorithm>
#include <chrono>
#include <iostream>

LightGlueDecoupleOnnxRunner::LightGlueDecoupleOnnxRunner(unsigned int threads) : num_threads(threads) {}

LightGlueDecoupleOnnxRunner::~LightGlueDecoupleOnnxRunner() {
  if (ctx_) rfe_destroy(ctx_);
}

int LightGlueDecoupleOnnxRunner::InitOrtEnv(Configuration cfg) {
  std::cout << "< - * -------- INITIAL ROVER_FE (B200) MATCHER START -------- * ->" << std::endl;
  rfe_config rc = {};
  rc.device = getenv("ROVER_FE_DEVICE") ? atoi(getenv("ROVER_FE_DEVICE")) : 0;
  const std::string& p = cfg.lightgluePath;
  rc.weights_path = (p.size() > 4 && p.substr(p.size() - 4) == ".rfw") ? p.c_str() : nullptr;
  rc.max_batch = 1;
  rc.max_height = 8;     // the matcher ctx does not extract: keep the SuperPoint buffers minimal
  rc.max_width = 8;
  rc.max_keypoints = cap_;
  rc.flags = RFE_FLAG_NO_EXTRACTOR;       // a matcher never extracts: no SuperPoint weights or activation buffers
  if (rfe_create(&rc, &ctx_) != RFE_OK) {
    std::cerr << "[ERROR] rover_fe matcher init failed : " << rfe_last_error() << std::endl;
    ctx_ = nullptr;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

// Matcher_PreProcess is the reference's pure function (lightglue_onnx.cpp:140-159): it only normalises.  Nothing is remembered
// between it and Matcher_Inference: the runner hands the NORMALISED keypoints to the device (rfe_lg_match_normalized), which
// is exactly the tensor the reference feeds its session, so any interleaving of the two calls behaves like the reference.
std::vector<cv::Point2f> LightGlueDecoupleOnnxRunner::Matcher_PreProcess(std::vector<cv::Point2f> kpts, int h, int w) {
  return NormalizeKeypoints(kpts, h, w);
}

std::vector<cv::Point2f> LightGlueDecoupleOnnxRunner::Matcher_PreProcess(std::vector<cv::KeyPoint> kpts, int h, int w) {
  std::vector<cv::Point2f> pf;
  pf.reserve(kpts.size());
  for (const cv::KeyPoint& k : kpts) pf.emplace_back(k.pt);
  return NormalizeKeypoints(pf, h, w);
}

LightGlueResult LightGlueDecoupleOnnxRunner::RunNormalized(const std::vector<float>& k0, const std::vector<float>& k1,
                                                           float* desc0, float* desc1) {
  LightGlueResult out;
  if (!ctx_) {
    std::cerr << "[ERROR] LightGlueDecoupleOnnxRunner Matcher inference failed : not initialised" << std::endl;
    return out;
  }
  const int n0 = static_cast<int>(k0.size() / 2), n1 = static_cast<int>(k1.size() / 2);
  out.matches.resize(static_cast<size_t>(std::max(n0, 1)) * 2);
  out.mscores.resize(std::max(n0, 1));
  int k = 0;
  auto t0 = std::chrono::high_resolution_clock::now();
  // threshold 0 here: the graph's own 0.1 filter applies; matchThresh is applied in Matcher_PostProcess_fused,
  // exactly where the reference applies it (lightglue_onnx.cpp:437-453)
  const int rc = rfe_lg_match_normalized(ctx_, k0.data(), n0, k1.data(), n1, desc0, desc1, 0.0f, out.matches.data(),
                                         out.mscores.data(), &k);
  auto t1 = std::chrono::high_resolution_clock::now();
  matcher_timer += std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
  if (rc != RFE_OK) {
    std::cerr << "[ERROR] LightGlueDecoupleOnnxRunner Matcher inference failed : " << rfe_last_error() << std::endl;
    return LightGlueResult();
  }
  out.count = k;
  out.ok = true;
  return out;
}

// kpts: the output of Matcher_PreProcess (normalised), as in the reference (lightglue_onnx.cpp:162-240)
LightGlueResult LightGlueDecoupleOnnxRunner::Matcher_Inference(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1,
                                                               float* desc0, float* desc1) {
  std::vector<float> k0(kpts0.size() * 2), k1(kpts1.size() * 2);
  for (size_t i = 0; i < kpts0.size(); ++i) { k0[2 * i] = kpts0[i].x; k0[2 * i + 1] = kpts0[i].y; }
  for (size_t i = 0; i < kpts1.size(); ++i) { k1[2 * i] = kpts1[i].x; k1[2 * i + 1] = kpts1[i].y; }
  return RunNormalized(k0, k1, desc0, desc1);
}

// The KeyPoint twin feeds kpts[i].pt to the session AS IS (lightglue_onnx.cpp:241-330, :268-275): no normalisation
// happens inside; the caller is expected to have done it.  Reproduced.
LightGlueResult LightGlueDecoupleOnnxRunner::Matcher_Inference(std::vector<cv::KeyPoint> kpts0, std::vector<cv::KeyPoint> kpts1,
                                                               float* desc0, float* desc1) {
  std::vector<float> k0(kpts0.size() * 2), k1(kpts1.size() * 2);
  for (size_t i = 0; i < kpts0.size(); ++i) { k0[2 * i] = kpts0[i].pt.x; k0[2 * i + 1] = kpts0[i].pt.y; }
  for (size_t i = 0; i < kpts1.size(); ++i) { k1[2 * i] = kpts1[i].pt.x; k1[2 * i + 1] = kpts1[i].pt.y; }
  return RunNormalized(k0, k1, desc0, desc1);
}

int LightGlueDecoupleOnnxRunner::Matcher_PostProcess_fused(LightGlueResult& output, std::vector<cv::Point2f> kpts0,
                                                           std::vector<cv::Point2f> kpts1, std::vector<int>& vnMatches12) {
  (void)kpts0; (void)kpts1;
  int size = 0;
  if (!output.ok) return 0;        // the reference would index an empty vector here (lightglue_onnx.cpp:404)
  for (int i = 0; i < output.count; ++i) {
    if (output.mscores[i] > this->matchThresh) {
      const int q = output.matches[2 * i];
      if (q >= 0 && q < static_cast<int>(vnMatches12.size())) {
        size++;
        vnMatches12[q] = output.matches[2 * i + 1];
      }
    }
  }
  return size;
}

float LightGlueDecoupleOnnxRunner::GetMatchThresh() { return matchThresh; }
void LightGlueDecoupleOnnxRunner::SetMatchThresh(float thresh) { matchThresh = thresh; }
std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> LightGlueDecoupleOnnxRunner::GetKeypointsResult() {
  return keypoints_result;                 // lightglue_onnx.cpp:505-508
}
double LightGlueDecoupleOnnxRunner::GetTimer(std::string name) {
  if (name == "extractor") return static_cast<double>(extractor_timer);
  return static_cast<double>(matcher_timer);
}
