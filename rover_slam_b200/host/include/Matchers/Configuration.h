// Runner configuration record of the learned front end.
//
// The reference fills one of these by member assignment before InitOrtEnv (SPextractor.cc:90-95, SPmatcher.cc:19-24) and passes
// default-constructed ones to the per-call entry points (SPmatcher.cc:364, 402, 447; SPextractor.cc:595), so the member NAMES
// and TYPES are the compatibility surface; nothing in the reference aggregate-initialises it, which leaves the order free.
// What each member means to this implementation is stated per member.
#pragma once

#include <string>

struct Configuration {
  // --- where the networks come from ------------------------------------------------------------------------------
  // The reference points these at onnxmodel/superpoint.onnx and onnxmodel/lightglue_sim.onnx (relative to the CWD).  Here
  // a path ending in ".rfw" selects a packed weight blob (tools/pack_weights.py); any other value falls through to
  // $ROVER_FE_WEIGHTS and then weights/rover_fe.rfw relative to the CWD.  One blob holds both networks.
  std::string extractorPath;
  std::string lightgluePath;

  // --- where they run --------------------------------------------------------------------------------------------------
  // "cuda" at both of the reference's call sites.  There is no CPU execution provider behind this class, so the string is not
  // interpreted: the runners always create a GPU context, on ordinal $ROVER_FE_DEVICE (default 0), and InitOrtEnv returns
  // EXIT_FAILURE when no sm_100 device is present.
  std::string device;

  // --- carried for source compatibility, not interpreted -----------------------------------------------------------------
  std::string extractorType;          // "superpoint"; only the reference's uncalled Extractor_PreProcess branches on it (superpoint_onnx.cc:78)
  unsigned int image_size = 512;      // resize hint; the reference's only use is commented out (superpoint_onnx.cc:76): frames keep sensor size
  float threshold = 0.0f;             // initial match threshold; SPmatcher overrides it through SetMatchThresh (SPmatcher.cc:25)
  bool grayScale = false;             // NormalizeImage decides by channel count (transform.cpp:5), not by this flag
  bool isEndtoEnd = true;             // selects the fused-vs-decoupled ONNX export upstream; one implementation here
  bool viz = false;
};
