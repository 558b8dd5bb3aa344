// Host-side pre-processing twins of the reference's src/Matchers/transform.cpp.  In the product path the u8 -> 1/255
// scaling and the keypoint normalisation happen inside the CUDA kernels (conv1a_kernel, posenc_kernel); these
// functions exist because callers of the reference may call them directly.
#include "Matchers/transform.h"

#include <algorithm>
#include <stdexcept>

// Same branches and the same exception as the reference (transform.cpp:3-17 throws std::invalid_argument for anything
// but 1 or 3 channels; SPextractor::operator() asserts CV_8UC1 before it gets here, SPextractor.cc:525).
cv::Mat NormalizeImage(cv::Mat& Image) {
  cv::Mat normalizedImage = Image.clone();
  if (Image.channels() == 3) {
    cv::cvtColor(normalizedImage, normalizedImage, cv::COLOR_BGR2RGB);
    normalizedImage.convertTo(normalizedImage, CV_32F, 1.0 / 255.0);
  } else if (Image.channels() == 1) {
    Image.convertTo(normalizedImage, CV_32F, 1.0 / 255.0);
  } else {
    throw std::invalid_argument("[ERROR] Not an image");
  }
  return normalizedImage;
}

cv::Mat RGB2Grayscale(cv::Mat& Image) {            // transform.cpp:85-89
  cv::Mat resultImage;
  cv::cvtColor(Image, resultImage, cv::COLOR_RGB2GRAY);
  return resultImage;
}

std::vector<cv::Point2f> NormalizeKeypoints(std::vector<cv::Point2f> kpts, int h, int w) {
  const cv::Point2f shift(static_cast<float>(w) / 2, static_cast<float>(h) / 2);
  const float scale = static_cast<float>((std::max)(w, h)) / 2;
  std::vector<cv::Point2f> out;
  out.reserve(kpts.size());
  for (const cv::Point2f& k : kpts) out.push_back((k - shift) / scale);
  return out;
}
