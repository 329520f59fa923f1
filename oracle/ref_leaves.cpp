// Thin C wrappers around the reference's own header-only leaves, compiled VERBATIM from
// /root/reference (see Makefile target `ref`; output oracle/_ref/libref_leaves.so, git-ignored).
// Used by tests/ to pin the oracle restatement against the real reference code.
#include <cstdint>
#include <vector>
#include "openvslam/orb_point_pairs.h"
#include "openvslam/trigonometric.h"
#include "openvslam/match_base.h"
#include "openvslam/match_angle_checker.h"

extern "C" {
unsigned ref_hamming(const uint32_t *a, const uint32_t *b) {
    return openvslam::match::compute_descriptor_distance_32(a, b);
}
float ref_cos(float v) { return openvslam::util::cos(v); }
float ref_sin(float v) { return openvslam::util::sin(v); }
void ref_pattern(float *out1024) {
    for (unsigned i = 0; i < openvslam::feature::orb_point_pairs_size; ++i) out1024[i] = openvslam::feature::orb_point_pairs[i];
}
int ref_angle_invalid(const float *deltas, const int *ids, int n, int *invalid_out) {
    openvslam::match::angle_checker<int> checker;
    for (int i = 0; i < n; ++i) checker.append_delta_angle(deltas[i], ids[i]);
    const std::vector<int> inv = checker.get_invalid_matches();
    for (size_t i = 0; i < inv.size(); ++i) invalid_out[i] = inv[i];
    return (int)inv.size();
}
int ref_angle_valid(const float *deltas, const int *ids, int n, int *valid_out) {
    openvslam::match::angle_checker<int> checker;
    for (int i = 0; i < n; ++i) checker.append_delta_angle(deltas[i], ids[i]);
    const std::vector<int> v = checker.get_valid_matches();
    for (size_t i = 0; i < v.size(); ++i) valid_out[i] = v[i];
    return (int)v.size();
}
unsigned ref_thr_low(void) { return openvslam::match::HAMMING_DIST_THR_LOW; }
unsigned ref_thr_high(void) { return openvslam::match::HAMMING_DIST_THR_HIGH; }
unsigned ref_max_dist(void) { return openvslam::match::MAX_HAMMING_DIST; }
}
