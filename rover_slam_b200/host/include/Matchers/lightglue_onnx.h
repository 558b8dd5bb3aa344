// LightGlueDecoupleOnnxRunner without ONNXRuntime (reference: include/Matchers/lightglue_onnx.h:10-67).
#pragma once
#include <opencv2/opencv.hpp>
#include <string>
#include <utility>
#include <vector>

#include "Matchers/Configuration.h"
#include "Matchers/transform.h"
#include "rover_fe.h"

struct LightGlueResult {               // the two ONNX outputs matches0 [K,2] / mscores0 [K] (lightglue_onnx.cpp:210-214)
  std::vector<int32_t> matches;
  std::vector<float> mscores;
  int count = 0;
  bool ok = false;
};

class LightGlueDecoupleOnnxRunner {
 public:
  const unsigned int num_threads;
  float matchThresh = 0.0f;
  long long extractor_timer = 0;
  long long matcher_timer = 0;
  std::vector<float> scales = {1.0f, 1.0f};
  LightGlueResult matcher_outputtensors;
  std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> keypoints_result;

  explicit LightGlueDecoupleOnnxRunner(unsigned int num_threads = 1);
  ~LightGlueDecoupleOnnxRunner();

  int InitOrtEnv(Configuration cfg);                                                        // lightglue_onnx.cpp:4-98
  std::vector<cv::Point2f> Matcher_PreProcess(std::vector<cv::KeyPoint> kpts, int h, int w);  // lightglue_onnx.cpp:140-159
  std::vector<cv::Point2f> Matcher_PreProcess(std::vector<cv::Point2f> kpts, int h, int w);
  // kpts are the NORMALISED keypoints produced by Matcher_PreProcess, exactly as in the reference; the runner keeps no state
  // between the two calls (the device consumes the normalised coordinates directly: rfe_lg_match_normalized)
  LightGlueResult Matcher_Inference(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1, float* desc0,
                                    float* desc1);                                          // lightglue_onnx.cpp:162-240
  LightGlueResult Matcher_Inference(std::vector<cv::KeyPoint> kpts0, std::vector<cv::KeyPoint> kpts1, float* desc0,
                                    float* desc1);                                          // lightglue_onnx.cpp:241-330 (pt used as is)
  int Matcher_PostProcess_fused(LightGlueResult& output, std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1,
                                std::vector<int>& vnMatches12);                             // lightglue_onnx.cpp:396-482
  float GetMatchThresh();
  void SetMatchThresh(float thresh);
  double GetTimer(std::string name);
  std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> GetKeypointsResult();        // lightglue_onnx.cpp:505-508
  rfe_ctx* context() { return ctx_; }

 private:
  LightGlueResult RunNormalized(const std::vector<float>& k0, const std::vector<float>& k1, float* desc0, float* desc1);
  rfe_ctx* ctx_ = nullptr;
  int cap_ = 8192;
};
