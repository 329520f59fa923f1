// Stand-in for DBoW2/FORB.h: 32-byte ORB descriptors as 1 x 32 CV_8U matrices, Hamming distance.
#pragma once
#include <opencv2/core.hpp>
namespace DBoW2 {
struct FORB {
    typedef cv::Mat TDescriptor;
    static const int L = 32;
    static int distance(const TDescriptor &a, const TDescriptor &b) {
        int d = 0;
        for (int i = 0; i < L; ++i) d += __builtin_popcount((unsigned)(a.data[i] ^ b.data[i]));
        return d;
    }
};
}  // namespace DBoW2
