// Same fields as the reference's include/Matchers/Configuration.h:6-20 (kept so callers compile unchanged).
#ifndef CONFIGURATION_H
#define CONFIGURATION_H

#include <string>

struct Configuration {
  std::string lightgluePath;   // reference: "onnxmodel/lightglue_sim.onnx"; here: optional RFW1 blob path
  std::string extractorPath;   // reference: "onnxmodel/superpoint.onnx";    here: optional RFW1 blob path
  std::string extractorType;
  bool isEndtoEnd = true;
  bool grayScale = false;
  unsigned int image_size = 512;
  float threshold = 0.0f;
  std::string device;          // "cuda" (the reference hard-codes it: SPextractor.cc:92, SPmatcher.cc:20)
  bool viz = false;
};
#endif  // CONFIGURATION_H
