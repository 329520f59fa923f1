#pragma once
#include <functional>
#include <vector>
#include "image.hpp"
namespace accelerated { namespace operations {
using Function = std::function<Future(Image &, Image &)>;
inline Future callUnary(Function &f, Image &in, Image &out) { return f(in, out); }
struct StandardFactory {
    virtual ~StandardFactory() = default;
    struct Builder {
        Builder &setInterpolation(Image::Interpolation) { return *this; }
        Builder &setBorder(Image::Border) { return *this; }
        Function build(const Image &) { return Function(); }
    };
    virtual Builder rescale() { return Builder(); }
    virtual Builder fixedConvolution2D(const std::vector<std::vector<double>> &) { return Builder(); }
};
} }  // namespace accelerated::operations
