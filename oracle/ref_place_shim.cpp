// C-ABI shim over the REFERENCE's own retrieval code, compiled from the source where it lies:
//   /root/reference/src/PlaceRecognizer.cc   CosineDescriptorIndex::add / query (:21-52), TemporalConsistencyVoter::vote (:54-68)
// cv::Mat and the three arithmetic calls it makes (cv::norm, Mat / double, Mat * Mat) come from the functional stand-in
// in oracle/stubs_cv/ - the control flow (recency window, score gate, ordering, top-K, vote streaks) is the
// reference's, the last ulps of a score are the stand-in's.  Built by oracle/Makefile into oracle/_ref/libref_place.so;
// tests/test_oracle_ref_place.py holds oracle/eigenplaces.py (and through it the device index) to it.  TEST INFRASTRUCTURE.
#include <vector>

#include "PlaceRecognizer.h"

extern "C" {
void* ref_index_new() { return new superslam::CosineDescriptorIndex; }
void ref_index_delete(void* h) { delete static_cast<superslam::CosineDescriptorIndex*>(h); }
int ref_index_size(void* h) { return static_cast<int>(static_cast<superslam::CosineDescriptorIndex*>(h)->size()); }
// descriptor as a [1, dim] row, or as a [dim, 1] column when as_column != 0 (normalizedRow reshapes either)
void ref_index_add(void* h, size_t id, float* desc, int dim, int as_column) {
  const cv::Mat d = as_column ? cv::Mat(dim, 1, CV_32F, desc) : cv::Mat(1, dim, CV_32F, desc);
  static_cast<superslam::CosineDescriptorIndex*>(h)->add(id, d);
}
int ref_index_query(void* h, float* desc, int dim, size_t exclude_recent, int top_k, float min_score, size_t* ids,
                    float* scores, int cap) {
  const cv::Mat d(1, dim, CV_32F, desc);
  const std::vector<superslam::LoopCandidate> c =
      static_cast<const superslam::CosineDescriptorIndex*>(h)->query(d, exclude_recent, top_k, min_score);
  for (size_t i = 0; i < c.size() && static_cast<int>(i) < cap; ++i) ids[i] = c[i].keyframe_id, scores[i] = c[i].score;
  return static_cast<int>(c.size());
}
void* ref_voter_new(int required_votes, size_t id_tolerance) {
  return new superslam::TemporalConsistencyVoter(required_votes, id_tolerance);
}
void ref_voter_delete(void* h) { delete static_cast<superslam::TemporalConsistencyVoter*>(h); }
// has_best == 0 -> vote(nullptr)
int ref_voter_vote(void* h, int has_best, size_t id, float score) {
  superslam::LoopCandidate c;
  c.keyframe_id = id, c.score = score;
  return static_cast<superslam::TemporalConsistencyVoter*>(h)->vote(has_best ? &c : nullptr) ? 1 : 0;
}
}
