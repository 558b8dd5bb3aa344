// ORB_SLAM3::SPmatcher -- the LightGlue-facing part of the reference's class (include/Matchers/SPmatcher.h:44-142):
// constructor, the four MatchingPoints_onnx overloads, the distance thresholds and DescriptorDistance_sp.
// Inside the Rover-SLAM tree the reference's own SPmatcher.h is kept (its Search*/Fuse members touch MapPoint
// pointers and stay host code); only the bodies in src/SPmatcher_onnx.cc replace SPmatcher.cc:17-27 and :359-542.
// ROVER_FE_STANDALONE supplies the minimal Frame this file needs when built outside that tree.
#ifndef SPMATCHER_H
#define SPMATCHER_H

#include <opencv2/opencv.hpp>
#include <vector>

#include "Matchers/lightglue_onnx.h"

namespace ORB_SLAM3 {

#ifdef ROVER_FE_STANDALONE
struct Frame {                       // the three members MatchingPoints_onnx(Frame&, Frame&, ...) reads (SPmatcher.cc:457-542)
  std::vector<cv::KeyPoint> mvKeys;
  cv::Mat mDescriptors;
  cv::Mat imgLeft;
};
#else
class Frame;
#endif

class SPmatcher {
 public:
  SPmatcher(float thre);        // no destructor, like the reference's header (SPmatcher.h:44-142): featureMatcher lives as long as the process
  int MatchingPoints_onnx(Frame& f1, Frame& f2, std::vector<int>& vnMatches12);
  int MatchingPoints_onnx(std::vector<cv::KeyPoint> kpts0, const std::vector<cv::KeyPoint> kpts1, cv::Mat desc0,
                          const cv::Mat desc1, std::vector<int>& vnMatches12);
  int MatchingPoints_onnx(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1, cv::Mat desc0, cv::Mat desc1,
                          std::vector<int>& vnMatches12);
  int MatchingPoints_onnx(std::vector<cv::Point2f> kpts0, std::vector<cv::Point2f> kpts1, float* desc0, float* desc1);
  static float DescriptorDistance_sp(const cv::Mat& a, const cv::Mat& b);   // SPmatcher.cc:2184-2189 (L2)

  static const float TH_LOW;
  static const float TH_HIGH;
  static const int HISTO_LENGTH;
  LightGlueDecoupleOnnxRunner* featureMatcher;
};

}  // namespace ORB_SLAM3
#endif
