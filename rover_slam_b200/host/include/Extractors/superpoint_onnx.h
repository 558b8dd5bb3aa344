// SuperPointOnnxRunner without ONNXRuntime: same class name and ORT-free members as the reference's
// include/Extractors/superpoint_onnx.h:10-68; the Ort::Value carriers are replaced by a POD result struct and
// the session by an rfe_ctx (include/rover_fe.h).  All arithmetic runs in librover_fe.so on the GPU.
#pragma once
#include <opencv2/opencv.hpp>
#include <string>
#include <utility>
#include <vector>

#include "Matchers/Configuration.h"
#include "Matchers/transform.h"
#include "rover_fe.h"

struct SuperPointResult {              // what the three ONNX outputs carried (superpoint_onnx.cc:150)
  std::vector<int32_t> keypoints;      // [N][2] (x, y)
  std::vector<float> scores;           // [N]
  std::vector<float> descriptors;      // [N][256]
  int count = 0;
};

class SuperPointOnnxRunner {
 public:
  const unsigned int num_threads;
  float matchThresh = 0.0f;
  long long extractor_timer = 0;       // milliseconds, like the reference (superpoint_onnx.cc:138-140)
  long long matcher_timer = 0;
  float lastmatch = 0;
  // the reference carries a score filter it compiles out (`bool adaptivethresold = false;`,
  // superpoint_onnx.cc:192-210): threshold = mean - 0.6 sigma - 0.02 / (1 + exp(-0.02 (lastmatch - 270))) over the scores of
  // the frame; keypoints below it are dropped in Extractor_PostProcess.  Off by default (= the reference's behaviour).
  bool adaptive_threshold = false;
  static float AdaptiveThreshold(const float* scores, int n, float lastmatch);
  // the `nfeatures` cap.  SPextractor stores nfeatures and never applies it (SPextractor.cc:84-146; the graph has
  // no top-K).  > 0: keep that many best-scoring keypoints per frame, in the graph's row-major order (rfe_sp_set_topk).
  // 0 (default) = the reference's behaviour: every keypoint.
  int max_keypoints_topk = 0;
  std::vector<float> scales = {1.0f, 1.0f};
  std::vector<SuperPointResult> extractor_outputtensors;
  std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> keypoints_result;           // superpoint_onnx.h:38

  explicit SuperPointOnnxRunner(unsigned int num_threads = 1);
  ~SuperPointOnnxRunner();

  int InitOrtEnv(Configuration cfg);                                               // superpoint_onnx.cc:4-66
  cv::Mat Extractor_PreProcess(Configuration cfg, const cv::Mat& Image, float& scale);    // superpoint_onnx.cc:68-86 (dead in the reference)
  int Extractor_Inference(Configuration cfg, const cv::Mat& image);                // superpoint_onnx.cc:88-162 (CV_32F [0,1] or CV_8UC1)
  void Extractor_PostProcess(Configuration cfg, SuperPointResult tensor, std::vector<cv::KeyPoint>& vKeyPoints,
                             cv::Mat& Descriptors);                                // superpoint_onnx.cc:165-255
  float GetMatchThresh();
  void SetMatchThresh(float thresh);
  double GetTimer(std::string name);                                               // superpoint_onnx.cc:268-277
  std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> GetKeypointsResult();   // superpoint_onnx.cc:279-282
  // SURVEY.md 8(f).1 -- the DBoW3 feed.  Frame::binarize_descriptors (Frame.cc:1034-1043) thresholds mDescriptors at 0
  // into a CV_8UC1 N x 256 matrix on every ComputeBoW3; the extractor already produced that matrix on the GPU.
  int BinarizeLast(cv::Mat& bin);                             // descriptors of the last Extractor_Inference
  int BinarizeDescriptors(const cv::Mat& desc, cv::Mat& bin); // any N x 256 CV_32F matrix (KeyFrame.cc:113-123)
  rfe_ctx* context() { return ctx_; }

 private:
  rfe_ctx* ctx_ = nullptr;
  int cap_ = 8192;
};
