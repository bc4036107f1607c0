// Drop-in for the reference's src/gaze_view_points.h (gaze_view_points.h:9-22, .cc:3-37): same
// class, same public members, same file format - one record per line that contains
//
//     frame,<uint>,forward,<float>,<float>,eye,<float>,<float>
//
// anywhere in it (the reference runs std::regex_search per line); `forward` is the head/view
// centre, `eye` the gaze point, both normalised to [0,1] (x right, y down) and fed to
// SampleFrameRectGPU / InterpolateFrameRectGPU as center_x / center_y
// (run_satlogrectilinear.cc:519-525).  pred_* of a record are the previous record's measured
// points (a one-frame-latency predictor), its own for the first record.
//
// The grammar of a <float> is the reference's regex  [-+]?\d*\.?\d+(?:[eE][-+]?\d+)?  and the
// text is converted with std::stof / std::stoul like the reference does; this header scans the
// line by hand instead of instantiating std::regex (which costs ~1 ms per line at -O0 and
// dominates the start-up of a 64-stream server).
#pragma once
#include <cctype>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

class GazeViewPoints {
 public:
  struct GazeViewPoint {
    unsigned int frame = 0;
    float view_point[2];
    float gaze_point[2];
    float pred_view_point[2];
    float pred_gaze_point[2];
  };

  std::vector<GazeViewPoint> points;
  GazeViewPoints() = default;
  explicit GazeViewPoints(std::string file_path) {
    std::ifstream file(file_path);
    if (!file.good()) {
      std::cerr << "Cannot open file: " << file_path << std::endl;  // gaze_view_points.cc:35
      return;
    }
    std::string line;
    while (std::getline(file, line)) AddLine(line);
  }

  // Parses one line; returns false when it holds no record (such lines are skipped).
  bool AddLine(const std::string &line) {
    for (size_t at = line.find("frame,"); at != std::string::npos; at = line.find("frame,", at + 1)) {
      GazeViewPoint p;
      if (!MatchRecord(line, at, &p)) continue;
      for (int k = 0; k < 2; ++k) {
        const GazeViewPoint &prev = points.empty() ? p : points.back();
        p.pred_view_point[k] = prev.view_point[k];
        p.pred_gaze_point[k] = prev.gaze_point[k];
      }
      points.push_back(p);
      return true;
    }
    return false;
  }

 private:
  static bool IsDigit(char c) { return std::isdigit(static_cast<unsigned char>(c)) != 0; }

  // Longest prefix of s[i..] matching [-+]?\d*\.?\d+(?:[eE][-+]?\d+)? that is followed by `next`
  // (0 = anything).  The regex engine backtracks to shorter matches when the literal after the
  // group fails; only two lengths can differ in what follows: with and without the exponent.
  static size_t MatchFloat(const std::string &s, size_t i, char next) {
    size_t j = i;
    if (j < s.size() && (s[j] == '-' || s[j] == '+')) ++j;
    size_t d0 = j;
    while (j < s.size() && IsDigit(s[j])) ++j;
    size_t mant_end = 0;
    if (j + 1 < s.size() && s[j] == '.' && IsDigit(s[j + 1])) {
      ++j;
      while (j < s.size() && IsDigit(s[j])) ++j;
      mant_end = j;
    } else if (j > d0) {
      mant_end = j;  // digits only
    } else {
      return 0;
    }
    size_t e = mant_end;
    if (e < s.size() && (s[e] == 'e' || s[e] == 'E')) {
      size_t k = e + 1;
      if (k < s.size() && (s[k] == '-' || s[k] == '+')) ++k;
      const size_t x0 = k;
      while (k < s.size() && IsDigit(s[k])) ++k;
      if (k > x0 && (next == 0 || (k < s.size() && s[k] == next))) return k - i;
    }
    if (next == 0 || (mant_end < s.size() && s[mant_end] == next)) return mant_end - i;
    return 0;
  }

  static bool Literal(const std::string &s, size_t *i, const char *lit) {
    const size_t n = std::char_traits<char>::length(lit);
    if (s.compare(*i, n, lit) != 0) return false;
    *i += n;
    return true;
  }

  static bool MatchRecord(const std::string &s, size_t i, GazeViewPoint *p) {
    if (!Literal(s, &i, "frame,")) return false;
    size_t j = i;
    while (j < s.size() && IsDigit(s[j])) ++j;
    if (j == i) return false;
    const std::string frame = s.substr(i, j - i);
    i = j;
    if (!Literal(s, &i, ",forward,")) return false;
    std::string f[4];
    const char *after[4] = {",", ",eye,", ",", nullptr};
    for (int k = 0; k < 4; ++k) {
      const size_t n = MatchFloat(s, i, after[k] ? ',' : 0);
      if (n == 0) return false;
      f[k] = s.substr(i, n);
      i += n;
      if (after[k] && !Literal(s, &i, after[k])) return false;
    }
    p->frame = static_cast<unsigned int>(std::stoul(frame));
    p->view_point[0] = std::stof(f[0]);
    p->view_point[1] = std::stof(f[1]);
    p->gaze_point[0] = std::stof(f[2]);
    p->gaze_point[1] = std::stof(f[3]);
    return true;
  }
};
