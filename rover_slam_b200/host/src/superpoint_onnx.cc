// SuperPointOnnxRunner on librover_fe.so.  Mirrors src/Extractors/superpoint_onnx.cc of the reference:
// same call sequence (InitOrtEnv -> Extractor_Inference -> Extractor_PostProcess), same return codes
// (EXIT_SUCCESS / EXIT_FAILURE, never throws), same millisecond timers.
#include "Extractors/superpoint_onnx.h"

#include <stdlib.h>

#include <chrono>
#include <cmath>
#include <iostream>

SuperPointOnnxRunner::SuperPointOnnxRunner(unsigned int threads) : num_threads(threads) {}

SuperPointOnnxRunner::~SuperPointOnnxRunner() {
  if (ctx_) rfe_destroy(ctx_);
}

int SuperPointOnnxRunner::InitOrtEnv(Configuration cfg) {
  std::cout << "< - * -------- INITIAL ROVER_FE (B200) EXTRACTOR START -------- * ->" << std::endl;
  rfe_config rc = {};
  rc.device = getenv("ROVER_FE_DEVICE") ? atoi(getenv("ROVER_FE_DEVICE")) : 0;
  // the reference opens "onnxmodel/superpoint.onnx" relative to the CWD (SPextractor.cc:93); a path ending in
  // ".rfw" selects an explicit weight blob, anything else falls back to $ROVER_FE_WEIGHTS / weights/rover_fe.rfw
  const std::string& p = cfg.extractorPath;
  rc.weights_path = (p.size() > 4 && p.substr(p.size() - 4) == ".rfw") ? p.c_str() : nullptr;
  rc.max_batch = 1;
  rc.max_height = 1024;
  rc.max_width = 1280;
  rc.max_keypoints = cap_;
  rc.flags = RFE_FLAG_NO_MATCHER;         // an extractor never matches: no LightGlue weights or state buffers
  if (rfe_create(&rc, &ctx_) != RFE_OK) {
    std::cerr << "[ERROR] rover_fe extractor init failed : " << rfe_last_error() << std::endl;
    ctx_ = nullptr;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

// Public but never called by the reference (superpoint_onnx.cc:68-86): normalise, then RGB -> gray for "superpoint".
// Like the reference it only works on a 3-channel image (cvtColor RGB2GRAY rejects 1 channel, transform.cpp:87).
cv::Mat SuperPointOnnxRunner::Extractor_PreProcess(Configuration cfg, const cv::Mat& Image, float& scale) {
  (void)scale;
  cv::Mat tempImage = Image.clone();
  cv::Mat resultImage = NormalizeImage(tempImage);
  if (cfg.extractorType == "superpoint") resultImage = RGB2Grayscale(resultImage);
  return resultImage;
}

int SuperPointOnnxRunner::Extractor_Inference(Configuration cfg, const cv::Mat& image) {
  (void)cfg;
  extractor_outputtensors.clear();
  if (!ctx_ || image.empty()) {
    std::cerr << "[ERROR] SuperPointOnnxRunner Extractor inference failed : not initialised / empty image" << std::endl;
    return EXIT_FAILURE;
  }
  // The reference hands a CV_32F image already scaled by 1/255 (SPextractor.cc:596-599).  The kernels take the
  // original bytes and apply the same float multiply on the device, so recover them exactly (v * 255 rounds back).
  cv::Mat gray;
  if (image.depth() == CV_8U && image.channels() == 1) {
    gray = image;
  } else if (image.depth() == CV_32F && image.channels() == 1) {
    gray.create(image.rows, image.cols, CV_8UC1);
    for (int r = 0; r < image.rows; ++r) {
      const float* s = image.ptr<float>(r);
      uint8_t* d = gray.ptr<uint8_t>(r);
      for (int c = 0; c < image.cols; ++c) d[c] = static_cast<uint8_t>(std::lround(s[c] * 255.0f));
    }
  } else {
    std::cerr << "[ERROR] SuperPointOnnxRunner Extractor inference failed : expected a 1-channel image" << std::endl;
    return EXIT_FAILURE;
  }
  if (rfe_sp_set_topk(ctx_, max_keypoints_topk < cap_ ? max_keypoints_topk : cap_) != RFE_OK) return EXIT_FAILURE;
  SuperPointResult res;
  res.keypoints.resize(static_cast<size_t>(cap_) * 2);
  res.scores.resize(cap_);
  res.descriptors.resize(static_cast<size_t>(cap_) * RFE_DESC_DIM);
  int32_t count = 0;
  auto t0 = std::chrono::high_resolution_clock::now();
  const int rc = rfe_sp_extract_u8(ctx_, gray.ptr<uint8_t>(0), gray.rows, gray.cols, static_cast<int>(gray.step), 1,
                                   res.keypoints.data(), res.scores.data(), res.descriptors.data(), &count, cap_);
  auto t1 = std::chrono::high_resolution_clock::now();
  extractor_timer += std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
  if (rc != RFE_OK && rc != RFE_ERR_CAPACITY) {
    std::cerr << "[ERROR] SuperPointOnnxRunner Extractor inference failed : " << rfe_last_error() << std::endl;
    return EXIT_FAILURE;
  }
  if (rc == RFE_ERR_CAPACITY)     // the reference returns every keypoint; say so when the fixed capacity truncates
    std::cerr << "[WARN] SuperPointOnnxRunner : " << count << " keypoints exceed the capacity " << cap_
              << "; the first " << cap_ << " in row-major order are returned" << std::endl;
  res.count = count < cap_ ? count : cap_;
  res.keypoints.resize(static_cast<size_t>(res.count) * 2);
  res.scores.resize(res.count);
  res.descriptors.resize(static_cast<size_t>(res.count) * RFE_DESC_DIM);
  extractor_outputtensors.emplace_back(std::move(res));
  return EXIT_SUCCESS;
}

void SuperPointOnnxRunner::Extractor_PostProcess(Configuration cfg, SuperPointResult tensor,
                                                 std::vector<cv::KeyPoint>& vKeyPoints, cv::Mat& Descriptors) {
  (void)cfg;
  // reference: threshold 0 with adaptivethresold=false keeps every keypoint (superpoint_onnx.cc:190-210); size=10,
  // octave=0 (:222-236).  The reference reads response from scores[2*idx] (out-of-bounds bug, :227) -- waived:
  // response = scores[idx].
  const int n = tensor.count;
  const float threshold = adaptive_threshold ? AdaptiveThreshold(tensor.scores.data(), n, lastmatch) : 0.0f;
  int keep = 0;
  for (int i = 0; i < n; ++i) keep += !(tensor.scores[i] < threshold);     // superpoint_onnx.cc:212-217
  Descriptors.create(keep, RFE_DESC_DIM, CV_32F);
  int row = 0;
  for (int i = 0; i < n; ++i) {
    if (tensor.scores[i] < threshold) continue;                            // superpoint_onnx.cc:226
    cv::KeyPoint kp;
    kp.pt = cv::Point2f(static_cast<float>(tensor.keypoints[2 * i]), static_cast<float>(tensor.keypoints[2 * i + 1]));
    kp.size = 10;
    kp.octave = 0;
    kp.response = tensor.scores[i];
    vKeyPoints.emplace_back(kp);
    memcpy(Descriptors.ptr<float>(row++), tensor.descriptors.data() + static_cast<size_t>(i) * RFE_DESC_DIM,
           sizeof(float) * RFE_DESC_DIM);
  }
}

// The reference's disabled rule, operation for operation (superpoint_onnx.cc:194-209): float sum / mean / variance
// accumulated in index order, the final expression in double, narrowed to float.
float SuperPointOnnxRunner::AdaptiveThreshold(const float* scores, int n, float lastmatch) {
  float sum = 0;
  for (int i = 0; i < n; i++) sum = sum + scores[i];
  float mean = sum / n;
  float variance = 0.0;
  for (int i = 0; i < n; i++) variance += (scores[i] - mean) * (scores[i] - mean);
  variance /= n;
  return static_cast<float>(mean - 0.6 * std::sqrt(variance) - 0.02 / (1.0 + std::exp(-0.02 * (lastmatch - 270))));
}

int SuperPointOnnxRunner::BinarizeLast(cv::Mat& bin) {
  if (!ctx_ || extractor_outputtensors.empty()) return EXIT_FAILURE;
  const int n = extractor_outputtensors.back().count;
  bin.create(n, RFE_DESC_DIM, CV_8UC1);
  if (n == 0) return EXIT_SUCCESS;
  int32_t count = 0;
  const int rc = rfe_sp_read_slot_bin(ctx_, 0, bin.ptr<uint8_t>(0), &count, n);
  if (rc != RFE_OK && rc != RFE_ERR_CAPACITY) {
    std::cerr << "[ERROR] SuperPointOnnxRunner BinarizeLast failed : " << rfe_last_error() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

int SuperPointOnnxRunner::BinarizeDescriptors(const cv::Mat& desc, cv::Mat& bin) {
  if (!ctx_ || desc.cols != RFE_DESC_DIM || desc.depth() != CV_32F) return EXIT_FAILURE;
  bin.create(desc.rows, RFE_DESC_DIM, CV_8UC1);
  if (desc.rows == 0) return EXIT_SUCCESS;
  if (rfe_binarize_descriptors(ctx_, desc.ptr<float>(0), desc.rows, bin.ptr<uint8_t>(0), nullptr) != RFE_OK) {
    std::cerr << "[ERROR] SuperPointOnnxRunner BinarizeDescriptors failed : " << rfe_last_error() << std::endl;
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}

std::pair<std::vector<cv::Point2f>, std::vector<cv::Point2f>> SuperPointOnnxRunner::GetKeypointsResult() {
  return keypoints_result;                 // superpoint_onnx.cc:279-282
}
float SuperPointOnnxRunner::GetMatchThresh() { return matchThresh; }
void SuperPointOnnxRunner::SetMatchThresh(float thresh) { matchThresh = thresh; }
double SuperPointOnnxRunner::GetTimer(std::string name) {
  if (name == "extractor") return static_cast<double>(extractor_timer);
  return static_cast<double>(matcher_timer);
}
