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
  if (rfe_create(&rc, &ctx_) != RFE_OK) {
    std::cerr << "[ERROR] rover_fe matcher init failed : " << rfe_last_error() << std::endl;
    ctx_ = nullptr;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

namespace {
int g_norm_h = 0, g_norm_w = 0;   // remembered per thread: Matcher_PreProcess(h, w) precedes Matcher_Inference
thread_local int t_norm_h = 0, t_norm_w = 0;
thread_local std::vector<cv::Point2f> t_px0, t_px1;
thread_local int t_which = 0;
}  // namespace

std::vector<cv::Point2f> LightGlueDecoupleOnnxRunner::Matcher_PreProcess(std::vector<cv::Point2f> kpts, int h, int w) {
  t_norm_h = h;
  t_norm_w = w;
  (t_which++ % 2 == 0 ? t_px0 : t_px1) = kpts;     // keep the pixel coordinates: the device normalises them itself
  return NormalizeKeypoints(kpts, h, w);
}

std::vector<cv::Point2f> LightGlueDecoupleOnnxRunner::Matcher_PreProcess(std::vector<cv::KeyPoint> kpts, int h, int w) {
  std::vector<cv::Point2f> pf;
  pf.reserve(kpts.size());
  for (const cv::KeyPoint& k : kpts) pf.emplace_back(k.pt);
  return Matcher_PreProcess(pf, h, w);
}

LightGlueResult LightGlueDecoupleOnnxRunner::Matcher_Inference(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1,
                                                               float* desc0, float* desc1) {
  LightGlueResult out;
  (void)g_norm_h; (void)g_norm_w;
  if (!ctx_ || t_norm_h <= 0 || t_norm_w <= 0) {
    std::cerr << "[ERROR] LightGlueDecoupleOnnxRunner Matcher inference failed : not initialised" << std::endl;
    return out;
  }
  // Recover pixel coordinates: prefer the ones remembered by Matcher_PreProcess (exact); otherwise invert
  // (kpt - shift) / scale.
  auto to_px = [&](const std::vector<cv::Point2f>& nk, const std::vector<cv::Point2f>& px) {
    std::vector<float> v(nk.size() * 2);
    if (px.size() == nk.size()) {
      for (size_t i = 0; i < nk.size(); ++i) { v[2 * i] = px[i].x; v[2 * i + 1] = px[i].y; }
    } else {
      const float sx = static_cast<float>(t_norm_w) / 2, sy = static_cast<float>(t_norm_h) / 2;
      const float sc = static_cast<float>(std::max(t_norm_w, t_norm_h)) / 2;
      for (size_t i = 0; i < nk.size(); ++i) { v[2 * i] = nk[i].x * sc + sx; v[2 * i + 1] = nk[i].y * sc + sy; }
    }
    return v;
  };
  const std::vector<float> p0 = to_px(kpts0, t_px0), p1 = to_px(kpts1, t_px1);
  const int n0 = static_cast<int>(kpts0.size()), n1 = static_cast<int>(kpts1.size());
  out.matches.resize(static_cast<size_t>(std::max(n0, 1)) * 2);
  out.mscores.resize(std::max(n0, 1));
  int k = 0;
  auto t0 = std::chrono::high_resolution_clock::now();
  // threshold 0 here: the graph's own 0.1 filter applies; matchThresh is applied in Matcher_PostProcess_fused,
  // exactly where the reference applies it (lightglue_onnx.cpp:437-453)
  const int rc = rfe_lg_match(ctx_, p0.data(), n0, p1.data(), n1, desc0, desc1, t_norm_h, t_norm_w, 0.0f,
                              out.matches.data(), out.mscores.data(), &k);
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
double LightGlueDecoupleOnnxRunner::GetTimer(std::string name) {
  if (name == "extractor") return static_cast<double>(extractor_timer);
  return static_cast<double>(matcher_timer);
}
