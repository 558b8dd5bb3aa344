// ORT-free twin of the reference's include/Matchers/transform.h.
#pragma once
#include <opencv2/opencv.hpp>
#include <vector>

cv::Mat NormalizeImage(cv::Mat& Image);                                                   // transform.cpp:3-17
cv::Mat RGB2Grayscale(cv::Mat& Image);                                                    // transform.cpp:85-89
std::vector<cv::Point2f> NormalizeKeypoints(std::vector<cv::Point2f> kpts, int h, int w);  // transform.cpp:19-32
