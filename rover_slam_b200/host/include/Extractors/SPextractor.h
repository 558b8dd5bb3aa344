// ORB_SLAM3::SPextractor with the reference's public surface (include/Extractors/SPextractor.h:55-136), minus the
// commented-out octree code.  Frame.cc / Tracking.cc / LocalMapping.cc call it unchanged.
#ifndef SPEXTRACTOR_H
#define SPEXTRACTOR_H

#include <opencv2/opencv.hpp>
#include <string>
#include <vector>

#include "Extractors/superpoint_onnx.h"

namespace ORB_SLAM3 {

class SPextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  SPextractor(int nfeatures, float scaleFactor, int nlevels, float iniThFAST, float minThFAST);
  ~SPextractor();

  // Returns the number of keypoints; appends to `keypoints` (the reference does not clear it,
  // superpoint_onnx.cc:230); `descriptors` becomes N x 256 CV_32F.
  int operator()(cv::InputArray image, std::vector<cv::KeyPoint>& keypoints, cv::Mat& descriptors);

  int inline GetLevels() { return nlevels; }
  float inline GetScaleFactor() { return scaleFactor; }
  std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
  std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
  std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
  std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }

  // DBoW3 feed: the CV_8UC1 N x 256 sign-binarised descriptors of the last operator() call, i.e. what
  // Frame::binarize_descriptors (Frame.cc:1034-1043) would compute from the returned descriptors.
  int GetBinaryDescriptors(cv::Mat& bin) { return featureExtractor ? featureExtractor->BinarizeLast(bin) : 1; }

  std::vector<cv::Mat> mvImagePyramid;
  SuperPointOnnxRunner* featureExtractor;
  std::string mModelstr = "onnx";
  float lastmatchnum = 0;

 protected:
  int ExtractSingleLayer(const cv::Mat& image, std::vector<cv::KeyPoint>& vKeyPoints, cv::Mat& localDescriptors);
  int ExtractMultiLayers(const cv::Mat& image, std::vector<cv::KeyPoint>& vKeyPoints, cv::Mat& Descriptors);

  int nfeatures;
  double scaleFactor;
  int nlevels;
  float iniThFAST;
  float minThFAST;
  std::vector<int> mnFeaturesPerLevel;
  std::vector<int> umax;
  std::vector<float> mvScaleFactor;
  std::vector<float> mvInvScaleFactor;
  std::vector<float> mvLevelSigma2;
  std::vector<float> mvInvLevelSigma2;
};

}  // namespace ORB_SLAM3
#endif
