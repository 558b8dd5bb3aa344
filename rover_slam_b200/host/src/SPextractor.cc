// ORB_SLAM3::SPextractor on the B200 front end (reference: src/Extractors/SPextractor.cc:84-146, :516-617).
#include "Extractors/SPextractor.h"

#include <cassert>
#include <cmath>
#include <iostream>

namespace ORB_SLAM3 {

SPextractor::SPextractor(int _nfeatures, float _scaleFactor, int _nlevels, float _iniThFAST, float _minThFAST)
    : nfeatures(_nfeatures), scaleFactor(_scaleFactor), nlevels(_nlevels), iniThFAST(_iniThFAST), minThFAST(_minThFAST) {
  if (mModelstr == "onnx") {
    Configuration cfg;
    cfg.device = "cuda";
    cfg.extractorPath = "onnxmodel/superpoint.onnx";
    cfg.extractorType = "superpoint";
    featureExtractor = new SuperPointOnnxRunner();
    featureExtractor->InitOrtEnv(cfg);
  }
  // ORB-style scale tables, as the reference builds them (SPextractor.cc:109-144)
  mvScaleFactor.resize(nlevels);
  mvLevelSigma2.resize(nlevels);
  mvScaleFactor[0] = 1.0f;
  mvLevelSigma2[0] = 1.0f;
  for (int i = 1; i < nlevels; i++) {
    mvScaleFactor[i] = static_cast<float>(mvScaleFactor[i - 1] * scaleFactor);
    mvLevelSigma2[i] = mvScaleFactor[i] * mvScaleFactor[i];
  }
  mvInvScaleFactor.resize(nlevels);
  mvInvLevelSigma2.resize(nlevels);
  for (int i = 0; i < nlevels; i++) {
    mvInvScaleFactor[i] = 1.0f / mvScaleFactor[i];
    mvInvLevelSigma2[i] = 1.0f / mvLevelSigma2[i];
  }
  mvImagePyramid.resize(nlevels);
  mnFeaturesPerLevel.resize(nlevels);
  const float factor = 1.0f / static_cast<float>(scaleFactor);
  float nDesiredFeaturesPerScale = nfeatures * (1 - factor) / (1 - static_cast<float>(std::pow((double)factor, (double)nlevels)));
  int sumFeatures = 0;
  for (int level = 0; level < nlevels - 1; level++) {
    mnFeaturesPerLevel[level] = static_cast<int>(std::lround(nDesiredFeaturesPerScale));
    sumFeatures += mnFeaturesPerLevel[level];
    nDesiredFeaturesPerScale *= factor;
  }
  mnFeaturesPerLevel[nlevels - 1] = std::max(nfeatures - sumFeatures, 0);
}

SPextractor::~SPextractor() { delete featureExtractor; }

int SPextractor::operator()(cv::InputArray _image, std::vector<cv::KeyPoint>& _keypoints, cv::Mat& _descriptors) {
  if (_image.empty()) return 0;
  cv::Mat image = _image.getMat();
  assert(image.type() == CV_8UC1);
  if (nlevels == 1) return ExtractSingleLayer(image, _keypoints, _descriptors);
  return ExtractMultiLayers(image, _keypoints, _descriptors);
}

int SPextractor::ExtractSingleLayer(const cv::Mat& image, std::vector<cv::KeyPoint>& vKeyPoints, cv::Mat& Descriptors) {
  if (mModelstr == "onnx") {
    Configuration cfg;
    featureExtractor->lastmatch = lastmatchnum;
    // the reference clones, NormalizeImage()s and copies the float image (SPextractor.cc:595-599); the u8 image goes
    // straight to the device here and the 1/255 multiply happens in conv1a_kernel
    if (featureExtractor->Extractor_Inference(cfg, image) != EXIT_SUCCESS) return 0;
    featureExtractor->Extractor_PostProcess(cfg, std::move(featureExtractor->extractor_outputtensors[0]), vKeyPoints,
                                            Descriptors);
  }
  return static_cast<int>(vKeyPoints.size());
}

int SPextractor::ExtractMultiLayers(const cv::Mat& image, std::vector<cv::KeyPoint>& vKeyPoints, cv::Mat& Descriptors) {
  // The reference's multi-level path has its inference calls commented out and yields no keypoints
  // (SPextractor.cc:619-653); reproduced: only nLevels == 1 extracts.
  (void)image; (void)vKeyPoints; (void)Descriptors;
  return 0;
}

}  // namespace ORB_SLAM3
