// Host-side pre-processing twins of the reference's src/Matchers/transform.cpp.  In the product path the u8 -> 1/255
// scaling and the keypoint normalisation happen inside the CUDA kernels (conv1a_kernel, posenc_kernel); these
// functions exist because callers of the reference may call them directly.
#include "Matchers/transform.h"

#include <algorithm>
#include <stdexcept>

cv::Mat NormalizeImage(cv::Mat& Image) {
  cv::Mat normalizedImage;
  if (Image.channels() == 1) {
    Image.convertTo(normalizedImage, CV_32F, 1.0 / 255.0);
  } else {
    throw std::invalid_argument("[ERROR] NormalizeImage: the SuperPoint front end takes 1-channel images");
  }
  return normalizedImage;
}

std::vector<cv::Point2f> NormalizeKeypoints(std::vector<cv::Point2f> kpts, int h, int w) {
  const cv::Point2f shift(static_cast<float>(w) / 2, static_cast<float>(h) / 2);
  const float scale = static_cast<float>((std::max)(w, h)) / 2;
  std::vector<cv::Point2f> out;
  out.reserve(kpts.size());
  for (const cv::Point2f& k : kpts) out.push_back((k - shift) / scale);
  return out;
}
